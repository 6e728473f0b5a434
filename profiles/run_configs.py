#!/usr/bin/env python
"""Parity + timing of the BASELINE.json configs other than the bench workload (C3):

  C2     reference big map (gray, 662x462, N=555), 6 agents, P=4096: evaluation + maximin over 2P
  C4-*   agent sweep 2/6/16/32 agents, L=640 (walks of 512), P=16384 on the 4096^2 world
  C5     16384^2 world, ~200k waypoints, P=32768 (the per-GPU shard of 262,144 / 8)

  python profiles/run_configs.py C2 C4-2 C4-6 ...        (on a GPU box; one JSON line per config)

Each config: build the world, draw a seeded population, evaluate on the GPU (timed with CUDA
events, 3 warm-ups + 5 timed), and check `--check` organisms against the oracle (bit-exact)."""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))

import numpy as np  # noqa: E402
import torch  # noqa: E402


def timed(fn, reps=5, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / reps


def check(eng, ow, dna, n_check, tol=0.0):
    import fitness_oracle as fo
    r = eng.evaluate(dna[:n_check], detail=False)
    obj, flags = r.obj.cpu().numpy(), r.flags.cpu().numpy()
    bad = 0
    for k in range(n_check):
        det = fo.calc_obj(ow, dna[k].cpu().numpy().astype(int), detail=True)
        ok = det['obj'][1] == obj[k, 1] and abs(det['obj'][0] - obj[k, 0]) <= tol and \
            det['feasible'] == flags[k, 0] and det['n_components'] == flags[k, 2]
        bad += not ok
    return bad


def run_c2(n_check):
    from world import World
    import fitness_oracle as fo
    from genetic_coverage_planning_b200 import packing, synth
    from genetic_coverage_planning_b200.engine import FitnessEngine
    w = World.load_npz(os.path.join(ROOT, "tests", "golden", "world_big.npz"))
    pw = packing.pack_world(w.img, w.xy, w.row_ptr, w.col, w.edge_poly, w.edge_theta, w.edge_dist,
                            w.map_params)
    op = w.org_params
    eng = FitnessEngine(pw, op)
    eng.set_vary_params(op, {'path_memory': 10, 'log_scale_weight': 0.7})
    P = 4096
    dna = eng.random_walks(P, op['start_idx'], 150, seed=0)
    eng.mutate(dna, seed=0, gen=0, salt=7)                 # constructor variation shapes
    ms = timed(lambda: eng.evaluate(dna, detail=False))
    obj = eng.evaluate(dna, detail=False).obj
    glad = torch.cat([obj, obj.flip(0) * 1.0001])
    ms_mm = timed(lambda: eng.maximin(glad, -0.3))
    fit = eng.maximin(glad, -0.3)
    ms_ord = timed(lambda: eng.arena_order(fit))
    want = fo.rack_n_stack(glad.cpu().numpy(), 0, dict(coverage_aging=80, coverage_constr_0=0.3,
                                                       coverage_constr_f=0.95))
    rank_ok = bool(np.array_equal(fit.cpu().numpy(), want) and
                   np.array_equal(eng.arena_order(fit)[0].cpu().numpy(), fo.arena_order(want)))
    t0 = time.time()
    for k in range(8):
        fo.calc_obj(w, dna[k].cpu().numpy().astype(int))
    cpu = 8 / (time.time() - t0)
    return {"config": "C2", "map": "big_map_scaled 662x462 (gray)", "N": pw.N, "E": pw.E, "A": 6, "L": 200,
            "P": P, "eval_ms": ms, "organisms_per_s": P / ms * 1e3, "kernels": "k_eval_fused<256, HAS_GRAY> (objectives-only path; gray map)",
            "maximin_ms_2P": ms_mm, "arena_order_ms_2P": ms_ord, "rank_bit_exact_vs_oracle": rank_ok,
            "oracle_mismatches": check(eng, w, dna, n_check, tol=1e-12), "checked": n_check,
            "cpu_port_organisms_per_s_1core": cpu}


def run_synth(name, n_check):
    import fitness_oracle as fo  # noqa: F401
    from world import World
    from genetic_coverage_planning_b200 import synth, worlds
    from genetic_coverage_planning_b200.engine import FitnessEngine
    cfg = worlds.CONFIGS[name]
    t0 = time.time()
    w, pw = worlds.build_synth_world(cfg['size'], seed=0, verbose=True)
    t_build = time.time() - t0
    A, L, P = cfg['agents'], cfg['L'], min(cfg['pop'], 32768 if name == 'C5' else cfg['pop'])
    starts = synth.spread_starts(w.xy, A, cfg['size'])
    op = worlds.org_params_for(cfg, starts)
    eng = FitnessEngine(pw, op, num_agents=A)
    dna = synth.random_walk_population_torch(w.xy, w.row_ptr, w.col, starts, P, cfg['walk'], L, seed=1,
                                             device=eng.device)
    ms = timed(lambda: eng.evaluate(dna, detail=False))
    S = float(worlds.executed_steps(dna[:1024].cpu().numpy()).mean())
    B_org = S * 544 + A * L * 4 + 64
    ow = World(w.img, w.xy, w.row_ptr, w.col, w.edge_poly, w.edge_theta, w.edge_dist, w.map_params,
               op, name=name)
    return {"config": name, "map": "%d^2 synthetic binary" % cfg['size'], "N": w.N, "E": w.E, "A": A,
            "L": L, "P": P, "eval_ms": ms, "organisms_per_s": P / ms * 1e3,
            "mean_steps": S, "algorithmic_GBps": B_org * P / (ms * 1e-3) / 1e9,
            "frac_of_6552": B_org * P / (ms * 1e-3) / 1e9 / 6552.3,
            "oracle_mismatches": check(eng, ow, dna, n_check), "checked": n_check,
            "world_build_s": t_build, "footprint_store_MB": pw.n_quarters * 128 / 1e6,
            "masked_store_MB": pw.m_n_quarters * 128 / 1e6}


if __name__ == "__main__":
    names = [a for a in sys.argv[1:] if not a.startswith("--")] or ["C2"]
    n_check = 4
    for a in sys.argv[1:]:
        if a.startswith("--check="):
            n_check = int(a.split("=")[1])
    for n in names:
        try:
            out = run_c2(n_check) if n == "C2" else run_synth(n, n_check)
        except Exception as e:                                   # report and keep going
            out = {"config": n, "error": "%s: %s" % (type(e).__name__, e)}
        print(json.dumps(out), flush=True)
