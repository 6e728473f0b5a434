"""GPU, two processes: the sharded generation (distributed.DistributedEvolution) must be
bit-identical to the single-process generation with the same seed.  Runs on a ONE-GPU box too:
both ranks then share cuda:0 and the collectives go through gloo (NCCL refuses two ranks on one
device); with >= 2 GPUs each rank takes its own device and the collectives are NCCL."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

WORKER = r'''
import os, sys
sys.path.insert(0, %(root)r); sys.path.insert(0, os.path.join(%(root)r, "oracle"))
import numpy as np, torch, torch.distributed as dist
from genetic_coverage_planning_b200 import packing, synth, worlds
from genetic_coverage_planning_b200.engine import FitnessEngine
from genetic_coverage_planning_b200.distributed import DistributedEvolution
rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
multi = torch.cuda.device_count() >= world
devi = rank if multi else 0
torch.cuda.set_device(devi)
dist.init_process_group("nccl" if multi else "gloo")
sw = synth.make_world(768, seed=3)
pw = packing.pack_world(sw.img, sw.xy, sw.row_ptr, sw.col, sw.edge_poly, sw.edge_theta, sw.edge_dist,
                        sw.map_params)
starts = synth.spread_starts(sw.xy, 3, 768)
op = worlds.org_params_for(dict(L=120), starts)
op.update(min_solo_lcs=0, min_comb_lcs=0)
gp = dict(worlds.GEN_PARAMS_6)
eng = FitnessEngine(pw, op, num_agents=3, device=devi)
eng.set_vary_params(op, sw.path_params)
P = 256
dna = eng.random_walks(P, starts, 80, seed=5)
obj = eng.evaluate(dna, detail=False).obj
fit = eng.maximin(obj, -gp["coverage_constr_0"])
ev = DistributedEvolution(eng, gp)
d, o, f = dna, obj, fit
for g in range(3):
    d, o, f = ev.step(d, o, f, seed=99, gen=g)
# single-process reference of the same three generations (every rank computes it itself)
d1, o1, f1 = dna, obj, fit
for g in range(3):
    kids = eng.select_vary(d1, f1, gp["gamma"], 99, g)
    ko = eng.evaluate(kids, detail=False).obj
    go = torch.cat([o1, ko])
    gf = eng.maximin(go, ev.constraint(g))
    order, _ = eng.arena_order(gf)
    d1, o1, f1 = eng.gather_survivors(order, d1, kids, go, gf)
assert torch.equal(d, d1) and torch.equal(o, o1) and torch.equal(f, f1), "sharded != single"
dist.barrier()
if rank == 0:
    print("DIST_OK backend=%%s" %% dist.get_backend())
dist.destroy_process_group()
'''


def test_sharded_generation_equals_single_process(tmp_path):
    script = os.path.join(tmp_path, "worker.py")
    with open(script, "w") as f:
        f.write(WORKER % {"root": ROOT})
    env = dict(os.environ)
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1",
                        "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
                        "--master-port", "29533", script], capture_output=True, text=True,
                       env=env, timeout=600)
    assert r.returncode == 0 and "DIST_OK" in r.stdout, r.stdout[-2000:] + r.stderr[-4000:]
