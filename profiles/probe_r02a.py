#!/usr/bin/env python
"""Round-2 planning probe (one B200): how the stages of a generation scale with the
population, for the strong-scaling design (global P = 65,536 split over 1/2/4/8 GPUs) and the
replicated ranking over up to 1 M gladiators; lowVarSample index mismatch rate at M = 65,536.

  python profiles/probe_r02a.py > gpurun_out/probe_r02a.jsonl
"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))

import numpy as np  # noqa: E402
import torch  # noqa: E402


def timed(fn, reps=10, warm=3):
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


def main():
    from genetic_coverage_planning_b200 import synth, worlds
    from genetic_coverage_planning_b200.engine import FitnessEngine
    import fitness_oracle as fo
    cfg = dict(worlds.CONFIGS["C3"])
    w, pw = worlds.build_synth_world(cfg['size'], seed=0, verbose=True)
    starts = synth.spread_starts(w.xy, cfg['agents'], cfg['size'])
    op = worlds.org_params_for(cfg, starts)
    eng = FitnessEngine(pw, op, num_agents=cfg['agents'])
    eng.set_vary_params(op, w.path_params)
    dev = torch.device("cuda", 0)
    dna_all = synth.random_walk_population_torch(w.xy, w.row_ptr, w.col, starts, 65536, cfg['walk'],
                                                 cfg['L'], seed=1, device=dev)
    gp = worlds.GEN_PARAMS_6
    for P in (4096, 8192, 16384, 32768, 65536):
        dna = dna_all[:P].contiguous()
        ms_eval = timed(lambda: eng.evaluate(dna, detail=False))
        obj = eng.evaluate(dna, detail=False).obj
        fit = eng.maximin(obj, -0.3)
        ms_sel = timed(lambda: eng.low_var_select(fit, 0.5, 0.1 / P))
        strong = eng.low_var_select(fit, 0.5, 0.1 / P)
        ms_x = timed(lambda: eng.crossover(dna, strong, 7, 1))
        kids = eng.crossover(dna, strong, 7, 1)
        base = kids.clone()

        def mut():
            kids.copy_(base)
            eng.mutate(kids, 7, 1)
        ms_copy = timed(lambda: kids.copy_(base))
        ms_mut = timed(mut) - ms_copy
        glad = torch.cat([obj, obj.flip(0) * 1.0001])
        ms_mm = timed(lambda: eng.maximin(glad, -0.3))
        gfit = eng.maximin(glad, -0.3)
        ms_ord = timed(lambda: eng.arena_order(gfit))
        order, _ = eng.arena_order(gfit)
        ms_gat = timed(lambda: eng.gather_survivors(order, dna, kids, glad, gfit))
        print(json.dumps({"probe": "stages_vs_P", "P": P, "eval_ms": ms_eval, "select_ms": ms_sel,
                          "crossover_ms": ms_x, "mutate_ms": ms_mut, "maximin_2P_ms": ms_mm,
                          "order_2P_ms": ms_ord, "gather_ms": ms_gat}), flush=True)
    # ranking over many gladiators (replicated ranking of a weak-scaled population)
    rs = np.random.RandomState(0)
    for n in (131072, 262144, 524288, 1048576):
        o = np.stack([-rs.rand(n) * (rs.rand(n) < 0.5), rs.rand(n) * 3.0], 1)
        glad = torch.from_numpy(o).to(dev)
        ms_mm = timed(lambda: eng.maximin(glad, -0.3), reps=5)
        gfit = eng.maximin(glad, -0.3)
        ms_ord = timed(lambda: eng.arena_order(gfit), reps=5)
        print(json.dumps({"probe": "ranking_vs_N", "n_gladiators": n, "maximin_ms": ms_mm,
                          "order_ms": ms_ord}), flush=True)
    # lowVarSample: index mismatch rate against the oracle's sequential cumulate
    for M in (5000, 65536):
        for kind in ("randn", "ties"):
            rs = np.random.RandomState(M)
            fit = rs.randn(M) * 0.3
            if kind == "ties":
                fit[rs.rand(M) < 0.7] = 0.25
            r0 = 0.37 / M
            got = eng.low_var_select(torch.from_numpy(fit), 0.5, r0).cpu().numpy()
            want = fo.low_var_sample_indices(fit, 0.5, r0)
            print(json.dumps({"probe": "low_var", "M": M, "kind": kind,
                              "mismatch_rate": float((got != want).mean()),
                              "max_abs_diff": int(np.abs(got - want).max())}), flush=True)


if __name__ == "__main__":
    main()
