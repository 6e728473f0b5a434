#!/usr/bin/env python
"""General evaluation path (any coverage_blend, all six pixel counters): the fused kernel over the
unmasked store against the staged general kernels (path_costs -> k_coverage -> loop_closures) - equality
of every output and time per launch, on C3 (binary) and on the reference's big gray map.
    GCP_LIB=<tag> python profiles/exp_general.py [P]"""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle"))
import numpy as np
import torch
from genetic_coverage_planning_b200 import packing, synth, worlds, _lib
from genetic_coverage_planning_b200.engine import FitnessEngine

P = int(sys.argv[1]) if len(sys.argv) > 1 else 65536


def timed(fn, reps=5):
    for _ in range(2):
        r = fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        r = fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / reps, r


def run(name, eng, dna, blend):
    out = {"world": name, "P": int(dna.shape[0]), "lib": os.path.basename(_lib.LIB_PATH), "blend": blend}
    ms_fast, fast = timed(lambda: eng.evaluate(dna, detail=False))
    out["objectives_only_ms"] = ms_fast
    ms_g, g = timed(lambda: eng.evaluate(dna, detail=True, pixel_counts=True))
    eng.set_option("fused_eval", 0)
    ms_s, s = timed(lambda: eng.evaluate(dna, detail=True, pixel_counts=True), 3)
    eng.set_option("fused_eval", 1)
    out["general_fused_ms"] = ms_g
    out["general_staged_ms"] = ms_s
    for k in ("obj", "neg_cov", "dist", "turn", "lc_mat", "flags", "cov"):
        out["same_" + k] = bool(torch.equal(getattr(g, k), getattr(s, k)))
    out["same_obj_as_objectives_only"] = bool(torch.equal(g.obj, fast.obj))
    out["error_flags"] = int((g.flags[:, 3] != 0).sum().item())
    out["mean_cov"] = [float(x) for x in g.cov.double().mean(0).cpu().numpy()[:6]]
    print(json.dumps(out), flush=True)


cfg = worlds.CONFIGS["C3"]
sw, pw = worlds.build_synth_world(cfg["size"], verbose=True)
starts = synth.spread_starts(sw.xy, cfg["agents"], cfg["size"] // 4)
op = worlds.org_params_for(cfg, starts)
for blend in (1.0, 0.1):
    pw.blend = blend
    eng = FitnessEngine(pw, op, num_agents=cfg["agents"])
    eng.set_vary_params(op, {'path_memory': 10, 'log_scale_weight': 0.7})
    dna = eng.random_walks(P, starts, cfg["walk"], seed=0)
    run("C3", eng, dna, blend)
    del eng

from world import World
w = World.load_npz(os.path.join(ROOT, "tests", "golden", "world_big.npz"))
for blend in (1.0, 0.1):
    w.map_params['coverage_blend'] = blend
    pwg = packing.pack_world(w.img, w.xy, w.row_ptr, w.col, w.edge_poly, w.edge_theta, w.edge_dist, w.map_params)
    eng = FitnessEngine(pwg, w.org_params)
    eng.set_vary_params(w.org_params, {'path_memory': 10, 'log_scale_weight': 0.7})
    dna = eng.random_walks(min(P, 16384), w.org_params['start_idx'], 150, seed=0)
    run("big_map_gray", eng, dna, blend)
    del eng
