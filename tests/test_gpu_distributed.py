"""GPU: the sharded generation loop (distributed.ShardedEvolution) must be bit-identical to the
plain single-GPU sequence of C-ABI calls with the same seed - for int16 and int32 chromosomes,
for one rank, and for two processes whose shards are mapped into each other with CUDA IPC.
The two-process test runs on a ONE-GPU box too: both ranks then share cuda:0 and the
collectives go through gloo (NCCL refuses two ranks on one device); with >= 2 GPUs each rank
takes its own device, the collectives are NCCL and the rendezvous is the peer-memory barrier."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

COMMON = r'''
import os, sys
sys.path.insert(0, %(root)r); sys.path.insert(0, os.path.join(%(root)r, "oracle"))
import numpy as np, torch, torch.distributed as dist
from genetic_coverage_planning_b200 import packing, synth, worlds
from genetic_coverage_planning_b200.engine import FitnessEngine
from genetic_coverage_planning_b200.distributed import ShardedEvolution

def setup(devi):
    sw = synth.make_world(768, seed=3)
    pw = packing.pack_world(sw.img, sw.xy, sw.row_ptr, sw.col, sw.edge_poly, sw.edge_theta, sw.edge_dist,
                            sw.map_params)
    starts = synth.spread_starts(sw.xy, 3, 768)
    op = worlds.org_params_for(dict(L=120), starts)
    op.update(min_solo_lcs=0, min_comb_lcs=0)
    gp = dict(worlds.GEN_PARAMS_6)
    eng = FitnessEngine(pw, op, num_agents=3, device=devi)
    eng.set_vary_params(op, sw.path_params)
    return eng, gp, starts

def reference_generations(eng, gp, dna, n_gen, seed):
    """the single-GPU sequence of calls (int32 chromosomes)"""
    d1 = dna
    o1 = eng.evaluate(d1, detail=False).obj
    f1 = eng.maximin(o1, -gp["coverage_constr_0"])
    for g in range(n_gen):
        kids = eng.select_vary(d1, f1, gp["gamma"], seed, g)
        ko = eng.evaluate(kids, detail=False).obj
        go = torch.cat([o1, ko])
        alpha = min(1, g / gp["coverage_aging"])
        c = (1 - alpha) * -gp["coverage_constr_0"] + alpha * -gp["coverage_constr_f"]
        gf = eng.maximin(go, c)
        order, _ = eng.arena_order(gf)
        d1, o1, f1 = eng.gather_survivors(order, d1, kids, go, gf)
    return d1, o1, f1
'''

WORKER = COMMON + r'''
rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
multi = torch.cuda.device_count() >= world
devi = rank if multi else 0
torch.cuda.set_device(devi)
dist.init_process_group("nccl" if multi else "gloo")
eng, gp, starts = setup(devi)
P = 256
dna = eng.random_walks(P, starts, 80, seed=5)
want = reference_generations(eng, gp, dna, 3, 99)
for gene_dtype in (torch.int16, torch.int32):
    ev = ShardedEvolution(eng, gp, P, gene_dtype=gene_dtype)
    assert ev.world == world and ev.per == P // world
    ev.load(dna, 0)
    for g in range(3):
        ev.step(99, g)
    ev.check()
    d = ev.gather_parents()
    assert torch.equal(d, want[0]), "sharded dna != single (%%s)" %% gene_dtype
    assert torch.equal(ev.obj, want[1]) and torch.equal(ev.fit, want[2]), "sharded obj/fit != single"
    torch.cuda.synchronize()
    dist.barrier()
    del ev
dist.barrier()
if rank == 0:
    print("DIST_OK backend=%%s barrier=%%s" %% (dist.get_backend(), "peer" if multi else "collective"))
dist.destroy_process_group()
'''

SINGLE = COMMON + r'''
eng, gp, starts = setup(0)
P = 512
dna = eng.random_walks(P, starts, 80, seed=6)
want = reference_generations(eng, gp, dna, 4, 7)
for gene_dtype in (torch.int16, torch.int32):
    ev = ShardedEvolution(eng, gp, P, gene_dtype=gene_dtype)
    ev.load(dna, 0)
    for g in range(4):
        ev.step(7, g)
    assert torch.equal(ev.gather_parents(), want[0])
    assert torch.equal(ev.obj, want[1]) and torch.equal(ev.fit, want[2])
print("SINGLE_OK")
'''


def test_one_rank_int16_and_int32_equal_plain_calls(tmp_path):
    script = os.path.join(tmp_path, "single.py")
    with open(script, "w") as f:
        f.write(SINGLE % {"root": ROOT})
    r = subprocess.run([sys.executable, script], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0 and "SINGLE_OK" in r.stdout, r.stdout[-2000:] + r.stderr[-4000:]


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
