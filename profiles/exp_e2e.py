#!/usr/bin/env python
"""End-to-end gcp_eval_host from pinned int32 host chromosomes: int32 on the wire against the host-side
narrowing to int16 with 1..N worker threads:  python profiles/exp_e2e.py [P]"""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402


def main():
    from genetic_coverage_planning_b200 import synth, worlds
    from genetic_coverage_planning_b200.engine import FitnessEngine
    P = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
    cfg = dict(worlds.CONFIGS["C3"])
    w, pw = worlds.build_synth_world(cfg['size'], seed=0)
    starts = synth.spread_starts(w.xy, 6, cfg['size'])
    op = worlds.org_params_for(cfg, starts)
    eng = FitnessEngine(pw, op, num_agents=6)
    dna = torch.cat([synth.random_walk_population_torch(w.xy, w.row_ptr, w.col, starts, 8192, 150, 200, seed=1000 + c,
                                                        device=eng.device) for c in range(P // 8192)]).contiguous()
    want = eng.evaluate(dna, detail=False).obj.cpu()
    dna_h = dna.cpu().pin_memory()
    obj_h = torch.empty((P, 2), dtype=torch.float64).pin_memory()
    flg_h = torch.empty((P, 4), dtype=torch.uint8).pin_memory()
    print(json.dumps({"host_cores": os.cpu_count()}), flush=True)
    for narrow, threads in [(0, 0), (1, 1), (1, 2), (1, 4), (1, 6), (1, 8), (1, 12), (1, 15), (1, 0)]:
        eng.set_option("host_narrow", narrow)
        eng.set_option("host_threads", threads)
        for _ in range(3):
            eng.evaluate_host(dna_h, obj_h, flg_h)
        t0 = time.perf_counter()
        n = 10
        for _ in range(n):
            eng.evaluate_host(dna_h, obj_h, flg_h)
        dt = (time.perf_counter() - t0) / n
        print(json.dumps({"P": P, "host_narrow": narrow, "host_threads": threads, "ms": dt * 1e3,
                          "organisms_per_s": P / dt, "same": bool(torch.equal(obj_h, want))}), flush=True)


if __name__ == "__main__":
    main()
