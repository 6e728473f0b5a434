#!/usr/bin/env python
"""Gray-map evaluation (the reference's big map, config C2 shape) at several population sizes:
GCP_LIB=<tag> python profiles/exp_gray.py"""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle"))
import torch
from world import World
from genetic_coverage_planning_b200 import packing, _lib
from genetic_coverage_planning_b200.engine import FitnessEngine
w = World.load_npz(os.path.join(ROOT, "tests", "golden", "world_big.npz"))
pw = packing.pack_world(w.img, w.xy, w.row_ptr, w.col, w.edge_poly, w.edge_theta, w.edge_dist, w.map_params)
op = w.org_params
eng = FitnessEngine(pw, op)
eng.set_vary_params(op, {'path_memory': 10, 'log_scale_weight': 0.7})
for P in (4096, 65536):
    dna = eng.random_walks(P, op['start_idx'], 150, seed=0)
    eng.mutate(dna, seed=0, gen=0, salt=7)
    for _ in range(5):
        r = eng.evaluate(dna, detail=False)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(20):
        r = eng.evaluate(dna, detail=False)
    b.record(); torch.cuda.synchronize()
    print(json.dumps({"lib": os.path.basename(_lib.LIB_PATH), "P": P, "ms": a.elapsed_time(b) / 20,
                      "checksum": float(r.obj.sum().item())}))
