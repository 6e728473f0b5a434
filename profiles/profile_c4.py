"""Workload for ncu captures of the long-genome kernels: C4-32 (32 agents x 640 genes), P = 2048.
  ncu --set full --clock-control none --import-source on -k regex:k_cov_quarters -s 2 -c 1 -o out python profiles/profile_c4.py"""
import sys, torch
sys.path.insert(0, "/root/repo")
from genetic_coverage_planning_b200 import synth, worlds
from genetic_coverage_planning_b200.engine import FitnessEngine
w, pw = worlds.build_synth_world(4096, seed=0)
cfg = worlds.CONFIGS["C4-32"]
A, L, P = cfg['agents'], cfg['L'], 2048
starts = synth.spread_starts(w.xy, A, 4096)
op = worlds.org_params_for(cfg, starts)
eng = FitnessEngine(pw, op, num_agents=A)
dna = synth.random_walk_population_torch(w.xy, w.row_ptr, w.col, starts, P, cfg['walk'], L, seed=1, device=eng.device)
for _ in range(4):
    eng.evaluate(dna, detail=False)
torch.cuda.synchronize()
print("ok")
