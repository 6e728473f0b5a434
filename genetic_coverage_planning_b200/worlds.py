"""Named benchmark worlds (BASELINE.json configs / SURVEY.md section 8d) with an on-disk
cache so that back-to-back bench runs and the ranks of one torchrun job build them once."""
import hashlib
import os
import pickle
import time

import numpy as np

from . import packing, synth

# reference pather_driver.py:122-158, 6-agent column
ORG_PARAMS_6 = {'max_dna_len': 200, 'min_dna_len': 30, 'crossover_prob': 0.7,
                'crossover_time_thresh': 70, 'mutation_prob': 0.3, 'muterpolate_prob': 0.2,
                'num_muterpolations': 20, 'muterpolation_srch_dist': 5,
                'muterpolation_sub_prob': 0.8, 'min_solo_lcs': 2, 'min_comb_lcs': 5,
                'flight_time_scale': 0.0001}
GEN_PARAMS_6 = {'gamma': 0.5, 'coverage_constr_0': 0.3, 'coverage_constr_f': 0.95,
                'coverage_aging': 80}

CONFIGS = {
    # name: map size, agents, max_dna_len, walk steps, population (the WHOLE population: bench.py
    # shards it over the ranks under --scaling strong)
    "C3": dict(size=4096, agents=6, L=200, walk=150, pop=65536),
    "C3-small": dict(size=1024, agents=6, L=200, walk=150, pop=2048),
    "C4-2": dict(size=4096, agents=2, L=640, walk=512, pop=16384),
    "C4-6": dict(size=4096, agents=6, L=640, walk=512, pop=16384),
    "C4-16": dict(size=4096, agents=16, L=640, walk=512, pop=16384),
    "C4-32": dict(size=4096, agents=32, L=640, walk=512, pop=16384),
    "C5": dict(size=16384, agents=6, L=200, walk=150, pop=262144),
}


def cache_dir():
    d = os.environ.get("GCP_CACHE", "/tmp/gcp_cache")
    os.makedirs(d, exist_ok=True)
    return d


def build_synth_world(size, seed=0, use_cache=True, verbose=False):
    """(SynthWorld, PackedWorld) for a size x size synthetic map; cached by (size, seed)
    and by a hash of the generator sources."""
    h = hashlib.sha1()
    for mod in (synth, packing):
        h.update(open(mod.__file__.replace(".pyc", ".py"), "rb").read())
    import torch
    mode = "device" if torch.cuda.is_available() else "host"     # footprint polygons: CUDA kernel when there is a GPU
    path = os.path.join(cache_dir(), "world_%d_%d_%s_%s.pkl" % (size, seed, h.hexdigest()[:10], mode))
    if use_cache and os.path.exists(path):
        with open(path, "rb") as f:
            return pickle.load(f)
    t0 = time.time()
    w = synth.make_world(size, seed, frustums=mode)
    t1 = time.time()
    pw = packing.pack_world(w.img, w.xy, w.row_ptr, w.col, w.edge_poly, w.edge_theta,
                            w.edge_dist, w.map_params)
    if verbose:
        import sys
        print("[worlds] size %d: N=%d E=%d world %.1fs (%s frustums%s) pack %.1fs" %
              (size, w.N, w.E, t1 - t0, mode,
               ", kernel %.1f ms, %d ambiguous" % (w.frustum_info["kernel_ms"], w.frustum_info["ambiguous_edges"])
               if mode == "device" else "", time.time() - t1), file=sys.stderr, flush=True)
    if use_cache:
        tmp = path + ".%d.tmp" % os.getpid()
        with open(tmp, "wb") as f:
            pickle.dump((w, pw), f, protocol=4)
        os.replace(tmp, path)
    return w, pw


def org_params_for(cfg, starts):
    op = dict(ORG_PARAMS_6)
    op['max_dna_len'] = cfg['L']
    op['start_idx'] = list(starts)
    return op


def executed_steps(dna):
    """S per organism: number of steps mappy.py:343-355 executes (first -1 / idx==L-1 stop)."""
    dna = np.asarray(dna)
    L = dna.shape[-1]
    neg = dna == -1
    first = np.where(neg.any(-1), neg.argmax(-1), L)
    return np.minimum(first, L - 1).sum(-1)
