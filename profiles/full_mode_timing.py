#!/usr/bin/env python
"""C3-scale timing of the evaluation variants: objectives-only wall coverage (fused kernel on the
pre-masked store) vs the general path (unmasked footprints against the map planes: blend < 1 or
all six pixel counters; staged sub-tile kernels of gcp_eval.cu).

Experiment r01 (not kept): running the general path through the quarter-granular machinery of the
fused kernel was SLOWER (17.8 ms fused, 14.4 ms staged) than the sub-tile kernels (10.8 ms): full
footprints fill most quarters of their blocks, so quarter granularity only multiplies the items."""
import sys, os, json, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from genetic_coverage_planning_b200 import synth, worlds
from genetic_coverage_planning_b200.engine import FitnessEngine
w, pw = worlds.build_synth_world(4096, seed=0)
A, L, P = 6, 200, 65536
starts = synth.spread_starts(w.xy, A, 4096)
op = worlds.org_params_for(dict(L=L), starts)
eng = FitnessEngine(pw, op, num_agents=A)
dna = synth.random_walk_population_torch(w.xy, w.row_ptr, w.col, starts, P, 150, L, seed=1, device=eng.device)
def timed(fn, reps=5):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / reps
out = {}
out["wall_fused_ms"] = timed(lambda: eng.evaluate(dna, detail=False))
out["all_counters_general_ms"] = timed(lambda: eng.evaluate(dna, detail=True, pixel_counts=True))
eng.set_params(op, blend=0.1)
out["blend0.1_general_ms"] = timed(lambda: eng.evaluate(dna, detail=False))
eng.set_params(op, blend=1.0)
eng.set_option("fused_eval", 0)
out["wall_staged_ms"] = timed(lambda: eng.evaluate(dna, detail=False))
eng.set_option("fused_eval", 1)
print(json.dumps(out))
