#!/usr/bin/env python
"""Small end-to-end pass over every kernel family for compute-sanitizer (memcheck / racecheck /
initcheck):  compute-sanitizer --tool racecheck python profiles/sanitize_smoke.py"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
import torch  # noqa: E402
from genetic_coverage_planning_b200 import packing, synth, worlds  # noqa: E402
from genetic_coverage_planning_b200.engine import FitnessEngine  # noqa: E402

sw = synth.make_world(512, seed=2)
pw = packing.pack_world(sw.img, sw.xy, sw.row_ptr, sw.col, sw.edge_poly, sw.edge_theta, sw.edge_dist,
                        sw.map_params)
starts = synth.spread_starts(sw.xy, 4, 512)
op = worlds.org_params_for(dict(L=96), starts)
eng = FitnessEngine(pw, op, num_agents=4)
eng.set_vary_params(op, sw.path_params)
P = 64
dna = eng.random_walks(P, starts, 70, seed=1)
eng.mutate(dna, seed=1, gen=0, salt=3)
fused = eng.evaluate(dna, detail=True, pixel_counts=False)          # k_eval_fused
full = eng.evaluate(dna, detail=True, pixel_counts=True)            # staged general kernels
eng.set_option("fused_eval", 0)
staged = eng.evaluate(dna, detail=False)                            # staged masked kernels
eng.set_option("fused_eval", 1)
eng.set_option("force_lc_big", 1)
big = eng.evaluate(dna, detail=True, pixel_counts=True)             # sort-based loop closures
eng.set_option("force_lc_big", 0)
assert torch.equal(fused.obj, full.obj) and torch.equal(fused.obj, staged.obj) and torch.equal(big.obj, full.obj)
e16 = eng.evaluate16(dna.to(torch.int16))
assert torch.equal(e16.obj, fused.obj)
for algo in (1, 2):
    eng.set_option("rank_algo", algo)
    fit = eng.maximin(torch.cat([fused.obj, fused.obj * 1.01]), -0.3)
    order, _ = eng.arena_order(fit)
eng.set_option("rank_algo", 0)
fit = eng.maximin(fused.obj, -0.3)
kids = eng.select_vary(dna, fit, 0.5, seed=7, gen=1)
ko = eng.evaluate(kids, detail=False).obj
go = torch.cat([fused.obj, ko])
gf = eng.maximin(go, -0.3)
order, _ = eng.arena_order(gf)
eng.gather_survivors(order, dna, kids, go, gf)
obj_h, _ = eng.evaluate_host(dna.cpu().pin_memory())
assert torch.equal(obj_h, fused.obj.cpu())
torch.cuda.synchronize()
print("sanitize smoke ok, launches", eng.launch_count())
