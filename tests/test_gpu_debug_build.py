"""GPU: the bounds-checked build (libgcp_debug.so, -DGCP_BOUNDS_CHECK: a device-side assert in
front of every shared-memory / scratch index the kernels form) runs every kernel family without
tripping an assert and reproduces the release library's results.  compute-sanitizer is not
available on the GPU pool; this is its stand-in."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

SCRIPT = r'''
import os, sys, hashlib
sys.path.insert(0, %(root)r)
import numpy as np, torch
from genetic_coverage_planning_b200 import _lib, packing, synth, worlds
from genetic_coverage_planning_b200.engine import FitnessEngine
from genetic_coverage_planning_b200.distributed import ShardedEvolution
assert os.path.basename(_lib.LIB_PATH) == %(lib)r
h = hashlib.sha1()
def add(t):
    h.update(t.detach().cpu().numpy().tobytes())
sw = synth.make_world(512, seed=2)
pw = packing.pack_world(sw.img, sw.xy, sw.row_ptr, sw.col, sw.edge_poly, sw.edge_theta, sw.edge_dist, sw.map_params)
for A, L, walk in ((4, 96, 70), (6, 200, 150), (8, 640, 500)):          # fused 256 / static shape / long genome
    starts = synth.spread_starts(sw.xy, A, 512)
    op = worlds.org_params_for(dict(L=L), starts)
    op.update(min_solo_lcs=0, min_comb_lcs=0)
    eng = FitnessEngine(pw, op, num_agents=A)
    eng.set_vary_params(op, sw.path_params)
    P = 64
    dna = eng.random_walks(P, starts, walk, seed=1)
    eng.mutate(dna, seed=1, gen=0, salt=3)
    for kw in (dict(detail=True, pixel_counts=False), dict(detail=True, pixel_counts=True), dict(detail=False)):
        r = eng.evaluate(dna, **kw); add(r.obj); add(r.flags)
    for opt, vals in (("fused_eval", (0, 1)), ("force_lc_big", (1, 2, 0)), ("force_general", (1, 0))):
        for v in vals:
            eng.set_option(opt, v)
            r = eng.evaluate(dna, detail=True, pixel_counts=True); add(r.obj); add(r.lc_mat)
    obj = eng.evaluate(dna, detail=False).obj
    for algo in (1, 2, 0):
        eng.set_option("rank_algo", algo)
        fit = eng.maximin(torch.cat([obj, obj * 1.0001]), -0.3); add(fit)
        order, nd = eng.arena_order(fit); add(order)
    for gene_dtype in ((torch.int16, torch.int32) if eng.gene16_available() else (torch.int32,)):
        ev = ShardedEvolution(eng, worlds.GEN_PARAMS_6, P, gene_dtype=gene_dtype)
        ev.load(dna, 0)
        for g in range(4):
            ev.step(3, g)
        add(ev.gather_parents()); add(ev.obj); add(ev.fit)
    torch.cuda.synchronize()
    eng.close()
print("DIGEST", h.hexdigest())
'''


def _run(tmp_path, lib, env_lib):
    script = os.path.join(tmp_path, "dbg_%s.py" % env_lib)
    with open(script, "w") as f:
        f.write(SCRIPT % {"root": ROOT, "lib": lib})
    env = dict(os.environ, GCP_LIB=env_lib)
    r = subprocess.run([sys.executable, script], capture_output=True, text=True, env=env, timeout=900)
    assert r.returncode == 0 and "Assertion" not in r.stderr, r.stdout[-1500:] + r.stderr[-3000:]
    return [l for l in r.stdout.splitlines() if l.startswith("DIGEST")][0]


def test_bounds_checked_build_is_clean_and_equal(tmp_path):
    if not os.path.exists(os.path.join(ROOT, "genetic_coverage_planning_b200", "libgcp_debug.so")):
        pytest.skip("libgcp_debug.so not built (python -m genetic_coverage_planning_b200.build --debug)")
    assert _run(tmp_path, "libgcp_debug.so", "debug") == _run(tmp_path, "libgcp.so", "release")
