#!/usr/bin/env python
"""CPU time to enqueue each stage of a generation (no device sync inside the timed loops)."""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from genetic_coverage_planning_b200 import synth, worlds
from genetic_coverage_planning_b200.engine import FitnessEngine
cfg = dict(worlds.CONFIGS["C3"])
w, pw = worlds.build_synth_world(cfg['size'], seed=0)
cst = synth.clustered_starts(w.xy, w.row_ptr, 6, cfg['size'])
op = worlds.org_params_for(cfg, cst)
eng = FitnessEngine(pw, op, num_agents=6)
eng.set_vary_params(op, w.path_params)
P = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
dna = synth.random_walk_population_torch(w.xy, w.row_ptr, w.col, cst, P, 150, 200, seed=5, device=eng.device)
d16 = dna.to(torch.int16)
obj = eng.evaluate(dna, detail=False).obj
glad = torch.cat([obj, obj * 1.0001])
fit = eng.maximin(glad, -0.3)
strong = eng.low_var_select(fit[:P], 0.5, 0.1 / P)
tab = eng.shard_table([d16.data_ptr()], P)
kids = torch.empty_like(d16)
order, _ = eng.arena_order(fit)
out = torch.empty_like(d16); oo = torch.empty((P, 2), dtype=torch.float64, device=eng.device); of = torch.empty(P, dtype=torch.float64, device=eng.device)
def t(name, fn, launches_of=None, n=50):
    torch.cuda.synchronize()
    l0 = eng.launch_count()
    t0 = time.perf_counter()
    for _ in range(n):
        fn()
    dt = time.perf_counter() - t0
    torch.cuda.synchronize()
    dw = time.perf_counter() - t0
    print(json.dumps({"stage": name, "enqueue_us": 1e6 * dt / n, "wall_us": 1e6 * dw / n, "launches": (eng.launch_count() - l0) / n}))
t("empty_kernel_torch", lambda: torch.empty(1, device=eng.device).zero_())
t("select", lambda: eng.low_var_select(fit[:P], 0.5, 0.1 / P))
t("crossover", lambda: eng.crossover_shards(tab, 2, strong, P, 0, P // 2, 1, 1, kids))
t("mutate", lambda: eng.mutate_genes(kids, 0, 1, 1))
t("eval16", lambda: eng.evaluate_genes(kids))
t("maximin", lambda: eng.maximin(glad, -0.3))
t("order", lambda: eng.arena_order(fit))
t("gather", lambda: eng.gather_shards(order, P, 0, P, tab, tab, 2, glad, fit, out, oo, of))
