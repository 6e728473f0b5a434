#!/usr/bin/env python
"""A/B of k_eval_fused variants on one GPU (options of gcp_set_option), results checked equal:
  python profiles/exp_fused.py C3|C5 [spread|diverse] [P] opt=a,b opt2=c,d ...
every combination of the listed option values is timed (CUDA events, 20 launches after 5)."""
import hashlib
import itertools
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
import torch  # noqa: E402


def main():
    from genetic_coverage_planning_b200 import synth, worlds
    from genetic_coverage_planning_b200.engine import FitnessEngine
    name = sys.argv[1]
    starts_mode = sys.argv[2]
    P = int(sys.argv[3])
    opts = [(a.split("=")[0], [int(v) for v in a.split("=")[1].split(",")]) for a in sys.argv[4:]]
    cfg = dict(worlds.CONFIGS[name])
    w, pw = worlds.build_synth_world(cfg['size'], seed=0)
    starts = synth.spread_starts(w.xy, 6, cfg['size'])
    op = worlds.org_params_for(cfg, starts)
    eng = FitnessEngine(pw, op, num_agents=6)
    parts = []
    for c in range(P // 8192):
        st = starts if starts_mode == "spread" else np.random.RandomState(1000 + c).randint(0, w.N, size=(8192, 6))
        parts.append(synth.random_walk_population_torch(w.xy, w.row_ptr, w.col, st, 8192, 150, 200, seed=1000 + c,
                                                        device=eng.device))
    dna = torch.cat(parts).contiguous()
    dna16 = dna.to(torch.int16) if w.N <= 32767 else None
    ref = None
    for combo in itertools.product(*[v for _, v in opts]):
        for (k, _), v in zip(opts, combo):
            eng.set_option(k, v)
        for label, fn in (("int32", lambda: eng.evaluate(dna, detail=False)),
                          ("int16", (lambda: eng.evaluate16(dna16)) if dna16 is not None else None)):
            if fn is None:
                continue
            for _ in range(5):
                r = fn()
            torch.cuda.synchronize()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            for _ in range(20):
                r = fn()
            b.record()
            torch.cuda.synchronize()
            if ref is None:
                ref = r.obj.clone()
            same = bool(torch.equal(ref, r.obj)) and not bool(r.flags[:, 3].any())
            print(json.dumps({"config": name, "starts": starts_mode, "P": P, "genes": label,
                              "options": dict(zip([k for k, _ in opts], combo)),
                              "ms": a.elapsed_time(b) / 20, "same_results": same,
                              "lib": os.environ.get("GCP_LIB", "release"),
                              "obj_sha1": hashlib.sha1(r.obj.cpu().numpy().tobytes()).hexdigest()[:12]}), flush=True)


if __name__ == "__main__":
    main()
