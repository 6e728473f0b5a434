#!/usr/bin/env python
"""The UNMODIFIED reference on a synthetic BINARY world of the kind the benchmark uses (rooms and
corridors generator, waypoints on a grid, 8-sector graph): dense reference tables are still feasible
at this size (768^2 map, N ~ 800), so a real Mappy / PathMaker is built on the synthetic map and
waypoints, Mappy.computeFrustums fills the edge tables, and getCoverage / getLoopClosures / calcObj
evaluate seeded random-walk populations (6 and 3 agents) plus the hand-made edge cases.

  -> tests/golden/world_synth.npz, eval_synth.npz, eval_synth_A3.npz, edge_synth.npz

This pins the binary-map path (bit-exact -coverage) and the synthetic generators to the
reference itself, not only to the oracle.

  python tests/golden/make_golden_synth.py      (this container only: needs /root/reference)
"""
import os
import sys
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
sys.path.insert(0, ROOT)
sys.path.insert(0, HERE)
import ref_shim  # noqa: E402
from world import World  # noqa: E402
import make_golden as mg  # noqa: E402

SIZE, SEED = 768, 3


def main():
    from genetic_coverage_planning_b200 import synth
    ga, mappy_m, pm_m, got = ref_shim.load()
    sw = synth.make_world(SIZE, seed=SEED)                  # map, waypoints, graph (+ host frustums)
    mp, pp, op, gp = ref_shim.driver_params(6)
    m = mappy_m.Mappy(sw.img.copy(), mp)
    p = pm_m.PathMaker(m, pp)
    p._XY = np.asarray(sw.xy).astype(int)
    g = np.zeros((sw.N, sw.N))
    src = np.repeat(np.arange(sw.N), np.diff(sw.row_ptr))
    g[src, sw.col] = 1
    p._graph = g
    p.findCrossoverCand()
    m.setWaypoints(p.waypoint_locs)
    t0 = time.time()
    m.computeFrustums(p._graph)                             # the reference's own table builder
    print("reference computeFrustums: N=%d E=%d in %.1fs" % (sw.N, sw.E, time.time() - t0))
    assert np.array_equal(m._traverse_frustum[src, sw.col], sw.edge_poly), "windowed restatement differs"
    starts = [int(v) for v in synth.spread_starts(sw.xy, 6, SIZE)]
    op['start_idx'] = starts
    w = World.from_reference(m, p, op, name="synth")
    w.save_npz(os.path.join(HERE, "world_synth.npz"))
    dnas = synth.random_walk_population(sw.xy, sw.row_ptr, sw.col, starts, 24, 150, op['max_dna_len'], seed=1)
    rec = mg.eval_records(ga, m, p, op, dnas)
    np.savez_compressed(os.path.join(HERE, "eval_synth.npz"), dna=dnas.astype(np.int16), versions=mg.VERSIONS, **rec)
    print("eval_synth: %d organisms, feasible %d, -coverage range [%.4f, %.4f]" %
          (len(dnas), int(np.sum(rec['obj'][:, 0] != 0)), rec['neg_cov'].min(), rec['neg_cov'].max()))
    ed = mg.edge_case_dnas(m, p, op)
    rec = mg.eval_records(ga, m, p, op, ed)
    np.savez_compressed(os.path.join(HERE, "edge_synth.npz"), dna=ed.astype(np.int16), versions=mg.VERSIONS, **rec)
    # 3 agents with the driver's 3-agent parameters
    mp3, _, op3, _ = ref_shim.driver_params(3)
    op3['start_idx'] = starts[:3]
    m._solo_sep_thresh = mp3['solo_sep_thresh']
    d3 = synth.random_walk_population(sw.xy, sw.row_ptr, sw.col, starts[:3], 16, 100, op3['max_dna_len'], seed=2)
    rec = mg.eval_records(ga, m, p, op3, d3)
    np.savez_compressed(os.path.join(HERE, "eval_synth_A3.npz"), dna=d3.astype(np.int16), versions=mg.VERSIONS,
                        solo_sep_thresh=m._solo_sep_thresh, min_solo_lcs=op3['min_solo_lcs'],
                        min_comb_lcs=op3['min_comb_lcs'], max_dna_len=op3['max_dna_len'], **rec)
    for f in ("world_synth.npz", "eval_synth.npz", "edge_synth.npz", "eval_synth_A3.npz"):
        print(f, "%.2f MB" % (os.path.getsize(os.path.join(HERE, f)) / 1e6))


if __name__ == "__main__":
    main()
