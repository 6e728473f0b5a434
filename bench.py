#!/usr/bin/env python
"""Benchmark of the fitness hot path (BASELINE.json metric: organisms evaluated/sec and
generations/sec at population 65,536, 6 agents, synthetic 4096^2 map = config C3, the population
sharded over 1/2/4/8 B200).

  python bench.py --gpus N --steps K --warmup W             # this repo's CUDA path
  python bench.py --impl reference --gpus N --steps K ...    # CPU restatement of the
        reference's own loop (oracle port, all host cores, no GPU) on bounded samples

A "step" = one evaluation pass (Organism.calcObj for every organism: path costs + coverage union
+ loop closures / constraints / objectives) over the population, DNA already resident in HBM.

--scaling strong (default, BASELINE configs[2]): ONE population of 65,536 organisms, rank r
evaluates organisms [r P/N, (r+1) P/N); tables replicated; the evaluation has no data-path
collective (organisms are independent).  --scaling weak: 65,536 per GPU.
Extra legs on the same JSON line: whole generations with the population sharded over the ranks
(strong and weak), ranking, per-kernel times, end to end from host memory, and - at N = 1 - the
CPU baselines (evaluation, rackNstack, one reference generation, config C1).
"""
import argparse
import json
import math
import os
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

METRIC = "organisms_evaluated_per_s"
UNIT = "organisms/s"
L2_BYTES = 126e6
CHUNK = 8192                 # organisms per seeded chunk of the synthetic population


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", default=os.environ.get("GCP_BENCH_CONFIG", "C3"))
    ap.add_argument("--scaling", default=os.environ.get("GCP_BENCH_SCALING", "strong"),
                    choices=["strong", "weak"])
    ap.add_argument("--pop", type=int, default=0, help="override the population of the config")
    ap.add_argument("--starts", default=os.environ.get("GCP_BENCH_STARTS", "spread"), choices=["spread", "diverse"],
                    help="spread: the config's fixed start waypoints (SURVEY 8d); diverse: every organism "
                         "starts from its own random waypoints, so a launch touches the WHOLE footprint "
                         "store (HBM-bound stress when the store outgrows the L2, config C5)")
    ap.add_argument("--cpu-sample", type=int, default=128)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extras", action="store_true", help="skip e2e / per-kernel / ranking / generation legs")
    return ap.parse_args()


def dist_env():
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    return rank, world, local


def measured_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def known_traffic(workload):
    """per-launch counters of the evaluation kernel from the committed ncu --set full capture
    (profiles/roofline_traffic.json): dram bytes, or a dict with more counters; or None."""
    p = os.path.join(ROOT, "profiles", "roofline_traffic.json")
    if os.path.exists(p):
        try:
            return json.load(open(p)).get(workload)
        except Exception:
            return None
    return None


class ClockSampler(object):
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        self.p = None
        try:
            self.p = subprocess.Popen(
                ["nvidia-smi", "-i", str(gpu_index), "--query-gpu=" + self.Q,
                 "--format=csv,noheader,nounits", "-lms", "100"], stdout=self.f,
                stderr=subprocess.DEVNULL)
        except Exception:
            self.p = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        if self.p is None:
            return out
        time.sleep(0.15)
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except Exception:
            self.p.kill()
        self.f.flush()
        rows = [l.strip().split(",") for l in open(self.f.name) if l.strip()]
        os.unlink(self.f.name)
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in rows:
            try:
                sm.append(float(r[0]))
                mx.append(float(r[1]))
                for k, n in enumerate(names):
                    if r[3 + k].strip().lower().startswith("active"):
                        reasons.add(n)
            except Exception:
                continue
        if sm:
            out.update(sm_mhz=float(np.median(sm)), sm_max_mhz=float(max(mx)),
                       reasons=sorted(reasons), samples=len(sm))
        return out


def config_of(name, pop_override):
    from genetic_coverage_planning_b200 import worlds
    cfg = dict(worlds.CONFIGS[name])
    if pop_override:
        cfg['pop'] = pop_override
    return cfg


def populations(cfg, scaling, world_size):
    """(organisms per rank, global population) of the timed step."""
    P = cfg['pop']
    if scaling == "weak":
        return P, P * world_size
    if P % (2 * world_size):
        raise SystemExit("population %d does not split over %d ranks" % (P, world_size))
    return P // world_size, P


def workload_string(name, cfg, N, E, scaling, world_size):
    per, glob = populations(cfg, scaling, world_size)
    return ("%s: synthetic %d^2 binary map, N=%d waypoints, E=%d edges, %d agents x %d genes "
            "(walks of %d steps), population %d sharded over %d GPU(s) (%s scaling)"
            % (name, cfg['size'], N, E, cfg['agents'], cfg['L'], cfg['walk'], glob, world_size, scaling))


def oracle_world(w, op):
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    from world import World
    ow = World(w.img, w.xy, w.row_ptr, w.col, w.edge_poly, w.edge_theta, w.edge_dist,
               w.map_params, op, name="synth")
    ow.path_params = dict(getattr(w, "path_params", {}) or {})
    return ow


_OW = None


def _cpu_eval_slice(dnas):
    import fitness_oracle as fo
    return [fo.calc_obj(_OW, d.astype(int)) for d in dnas]


def cpu_port_rate(w, op, dna_sample, procs):
    """organisms/s of the oracle port (restatement of the reference loop) on `procs`
    host processes.  Returns (rate, seconds)."""
    global _OW
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import fitness_oracle as fo
    _OW = oracle_world(w, op)
    _OW.safety_img  # build once, shared with forked workers
    _OW.edge_id(0, 0)
    if procs <= 1:
        t0 = time.time()
        for d in dna_sample:
            fo.calc_obj(_OW, d.astype(int))
        dt = time.time() - t0
        return len(dna_sample) / dt, dt
    import multiprocessing as mp
    ctx = mp.get_context("fork")
    chunks = [c for c in np.array_split(dna_sample, procs) if len(c)]
    with ctx.Pool(len(chunks)) as pool:
        pool.map(_cpu_eval_slice, [c[:1] for c in chunks])        # warm the workers
        t0 = time.time()
        pool.map(_cpu_eval_slice, chunks)
        dt = time.time() - t0
    return len(dna_sample) / dt, dt


# --------------------------------------------------------------------------------------------
# reference arm: the CPU restatement of the reference's own loop, all host cores, NO GPU:
# nothing below imports torch, the engine or libgcp.so (the world is built with the host
# restatement of computeFrustums, synth.make_world(frustums="host"))
# --------------------------------------------------------------------------------------------
def host_world(cfg):
    import pickle
    from genetic_coverage_planning_b200 import synth
    d = os.environ.get("GCP_CACHE", "/tmp/gcp_cache")
    os.makedirs(d, exist_ok=True)
    path = os.path.join(d, "hostworld_%d_0.pkl" % cfg['size'])
    if os.path.exists(path):
        with open(path, "rb") as f:
            return pickle.load(f)
    w = synth.make_world(cfg['size'], 0, frustums="host")
    tmp = path + ".%d.tmp" % os.getpid()
    with open(tmp, "wb") as f:
        pickle.dump(w, f, protocol=4)
    os.replace(tmp, path)
    return w


def run_reference(args):
    rank, world_size, _ = dist_env()
    if rank != 0:
        return 0
    from genetic_coverage_planning_b200 import synth, worlds      # numpy-only modules
    cfg = config_of(args.config, args.pop)
    w = host_world(cfg)
    starts = synth.spread_starts(w.xy, cfg['agents'], cfg['size'])
    op = worlds.org_params_for(cfg, starts)
    cores = os.cpu_count() or 1
    procs = max(1, min(cores, 64))
    per_step = max(procs * 6, 32)            # ~1-2 s of host work per step
    n = per_step * (args.steps + args.warmup)
    dna = synth.random_walk_population(w.xy, w.row_ptr, w.col, starts, n, cfg['walk'], cfg['L'], seed=1)
    secs = []
    for s in range(args.steps + args.warmup):
        r, dt = cpu_port_rate(w, op, dna[s * per_step:(s + 1) * per_step], procs)
        if s >= args.warmup:
            secs.append(dt)
    value = per_step * len(secs) / sum(secs)
    per, glob = populations(cfg, args.scaling, args.gpus)
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT,
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": 1000.0 * sum(secs) / len(secs), "higher_is_better": True,
        "scaling": args.scaling, "vs_baseline": None, "dtype": "u64 bitsets + f64 costs",
        "data": "synthetic",
        "config": {"workload": workload_string(args.config, cfg, w.N, w.E, args.scaling, args.gpus),
                   "organisms_per_gpu": per, "global_population": glob},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": procs, "kind": "port",
                         "sample": "%d organisms of the same kind of seeded population per step (the "
                                   "full population would take hours), Organism.calcObj restated in "
                                   "numpy/cv2 (oracle/fitness_oracle.py), %d forked processes; rate = "
                                   "sample / time; no GPU, no libgcp.so in this process"
                                   % (per_step, procs)},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    assert "torch" not in sys.modules and "genetic_coverage_planning_b200._lib" not in sys.modules, \
        "the reference arm must stay a pure CPU process"
    print(json.dumps(line), flush=True)
    return 0


# --------------------------------------------------------------------------------------------
# CPU baselines printed beside the GPU legs (rank 0, N = 1): BASELINE.md section 3 items 1-4
# --------------------------------------------------------------------------------------------
def cpu_ranking_baseline():
    """GeneticalGorithm.rackNstack's literal double loop (geneticalgorithm.py:109-120) at
    N = 256 / 512 / 1024, seconds per ordered pair, extrapolated by N^2."""
    import fitness_oracle as fo
    rs = np.random.RandomState(0)
    gp = dict(coverage_aging=80, coverage_constr_0=0.3, coverage_constr_f=0.95)
    out = {}
    per_pair = []
    for n in (256, 512, 1024):
        o = np.stack([-rs.rand(n) * (rs.rand(n) < 0.5), rs.rand(n) * 3.0], 1)
        t0 = time.time()
        fo.rack_n_stack_loop(o, 0, gp)
        dt = time.time() - t0
        out["N=%d_s" % n] = dt
        per_pair.append(dt / (n * (n - 1)))
    out["us_per_ordered_pair"] = float(np.mean(per_pair) * 1e6)
    return out


def c1_world():
    """BASELINE configs[0] (reference map_scaled.png + wilk_3 graph, 6 agents) from the committed
    fixture of the reference's own tables (tests/golden/world_small.npz)."""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    from world import World
    w = World.load_npz(os.path.join(ROOT, "tests", "golden", "world_small.npz"))
    w.path_params = {'path_memory': 10, 'log_scale_weight': 0.7}
    op = dict(w.org_params)
    gp = {'gen_size': 100, 'starting_path_len': 150, 'num_agents': 6, 'gamma': 0.5,
          'coverage_constr_0': 0.0, 'coverage_constr_f': 0.95, 'coverage_aging': 80, 'org_params': op}
    return w, gp


def cpu_generation_baseline(n_gen=2):
    """One reference generation at P = 100 on config C1 (pather_driver.py:235-240): the variation
    oracle replays the reference's loop call for call (pinned to it by tests/test_oracle_variation.py).
    coverage_constr_0 = 0 skips the first generation's rejection loop (466 s)."""
    import variation_oracle as vo
    w, gp = c1_world()
    np.random.seed(0)
    t0 = time.time()
    pop = vo.Population(w, gp)
    t_first = time.time() - t0
    t0 = time.time()
    pop.run_evolution(n_gen)
    dt = (time.time() - t0) / n_gen
    return {"population": gp['gen_size'], "s_per_generation": dt, "generations_per_s": 1.0 / dt,
            "organisms_varied_and_evaluated_per_s": gp['gen_size'] / dt,
            "first_generation_s": t_first, "generations_timed": n_gen, "cores": 1, "kind": "port",
            "what": "oracle/variation_oracle.Population.run_evolution: lowVarSample + crossover + "
                    "Organism constructor (mutation, muterpolate, pruneUTurns, calcObj) + rackNstack "
                    "over 2P + argsort, single thread like the reference"}


def gpu_c1_run(device, n_gen=100):
    """BASELINE configs[0] through the drop-in facade: GeneticalGorithm(...).runEvolution(100) at
    P = 100 (pather_driver.py:235-240)."""
    import torch
    from genetic_coverage_planning_b200 import synth
    from genetic_coverage_planning_b200.geneticalgorithm import GeneticalGorithm
    w, gp = c1_world()
    mappy, pather = synth.CsrMappy(w), synth.CsrPather(w)
    t0 = time.time()
    pop = GeneticalGorithm(mappy, pather, gp, device=device, seed=0)
    torch.cuda.synchronize()
    t_first = time.time() - t0
    pop.runEvolution(2)
    torch.cuda.synchronize()
    t0 = time.time()
    hist = pop.runEvolution(n_gen)
    torch.cuda.synchronize()
    dt = time.time() - t0
    last = np.array(hist[-1])
    return {"population": gp['gen_size'], "generations": n_gen, "s_total": dt,
            "generations_per_s": n_gen / dt, "first_generation_s": t_first,
            "final_feasible_fraction": float((last[:, 0] < 0).mean()),
            "final_best_coverage": float(-last[:, 0].min()),
            "api": "genetic_coverage_planning_b200.geneticalgorithm.GeneticalGorithm(mappy, pather, "
                   "gen_params).runEvolution(100), wall clock incl. the per-generation history the "
                   "reference returns"}


def main():
    args = parse_args()
    if args.impl == "reference":
        return run_reference(args)

    import torch
    import torch.distributed as dist
    rank, world_size, local = dist_env()
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (no CPU fallback)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    distributed = world_size > 1
    # one process per GPU: keep the process (and the pinned staging buffers it allocates from here
    # on) on the GPU's NUMA node, otherwise the e2e leg's H2D copies cross the socket interconnect
    from genetic_coverage_planning_b200.distributed import ShardedEvolution, bind_to_gpu_numa_node
    numa_node = bind_to_gpu_numa_node(local)
    if distributed:
        dist.init_process_group("nccl", device_id=dev)

    def barrier():
        if distributed:
            dist.barrier()

    def max_over_ranks(x):
        if not distributed:
            return float(x)
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    from genetic_coverage_planning_b200 import synth, worlds
    from genetic_coverage_planning_b200.engine import FitnessEngine

    cfg = config_of(args.config, args.pop)
    if rank == 0:
        w, pw = worlds.build_synth_world(cfg['size'], seed=0, verbose=True)
    barrier()
    if rank != 0:
        w, pw = worlds.build_synth_world(cfg['size'], seed=0)
    starts = synth.spread_starts(w.xy, cfg['agents'], cfg['size'])
    op = worlds.org_params_for(cfg, starts)
    A, L = cfg['agents'], cfg['L']
    P, P_global = populations(cfg, args.scaling, world_size)
    eng = FitnessEngine(pw, op, num_agents=A, device=local)
    eng.set_vary_params(op, w.path_params)
    for kv in filter(None, os.environ.get("GCP_OPTS", "").split(",")):     # tuning experiments
        k, v = kv.split("=")
        eng.set_option(k, int(v))

    def population_slice(lo, hi, start_nodes, seed_base):
        """organisms [lo, hi) of the seeded synthetic population (chunks of CHUNK organisms, each
        with its own seed, so the global population does not depend on the number of ranks)."""
        parts = []
        for c in range(lo // CHUNK, (hi + CHUNK - 1) // CHUNK):
            st = start_nodes
            if st is None:                                   # diverse: per-organism random starts
                st = np.random.RandomState(seed_base + c).randint(0, w.N, size=(CHUNK, A))
            d = synth.random_walk_population_torch(w.xy, w.row_ptr, w.col, st, CHUNK,
                                                   cfg['walk'], L, seed=seed_base + c, device=dev)
            parts.append(d[max(lo, c * CHUNK) - c * CHUNK:min(hi, (c + 1) * CHUNK) - c * CHUNK])
        return torch.cat(parts).contiguous()

    t0 = time.time()
    # the timed step must read its input from HBM: one population if it is larger than the L2,
    # otherwise a rotation of distinct populations whose total is (steps are launched back to back)
    dna_bytes = P * A * L * 4
    n_rot = 1 if dna_bytes > 1.2 * L2_BYTES else int(math.ceil(2.5 * L2_BYTES / dna_bytes))
    lo = rank * P
    dnas = [population_slice(lo, lo + P, starts if args.starts == "spread" else None, 1000 + 100000 * k)
            for k in range(n_rot)]
    dna = dnas[0]
    torch.cuda.synchronize()
    if rank == 0:
        print("[bench] %d population(s) of %d x %d x %d generated in %.1fs" % (n_rot, P, A, L, time.time() - t0),
              file=sys.stderr, flush=True)
    S_mean = float(worlds.executed_steps(dna[:4096].cpu().numpy()).mean())
    B_org = S_mean * 544 + A * L * 4 + 64

    # ---------------- timed region: K evaluation steps ----------------
    # clocks are sampled (nvidia-smi, 100 ms period) from before the warm-up to the end of the
    # timed region: same kernel, same load, enough samples even when K steps take < 100 ms
    sampler = ClockSampler(local) if rank == 0 else None
    n_warm = max(args.warmup, 3) + 30
    for s in range(n_warm):
        res = eng.evaluate(dnas[s % n_rot], detail=False)
    torch.cuda.synchronize()
    barrier()
    l0 = eng.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    for s in range(args.steps):
        res = eng.evaluate(dnas[s % n_rot], detail=False)
    e1.record()
    torch.cuda.synchronize()
    barrier()
    launches = eng.launch_count() - l0
    ms = max_over_ranks(e0.elapsed_time(e1))
    clocks = sampler.stop() if sampler else None
    ms_per_step = ms / args.steps
    value = P_global / (ms_per_step / 1000.0)
    res = eng.evaluate(dna, detail=False)

    extras = {}
    roofline = None
    e2e = None
    cpu_baseline = None
    if not args.no_extras:
        # ---------------- per-kernel timing (CUDA events on the launch stream) ----------
        fast = eng.has_masked and pw.is_binary and pw.blend == 1.0
        if fast:
            edge_ids, step_info, d_dist, d_turn, flags = eng.path_costs(dna, with_step_info=True)
        else:
            edge_ids, d_dist, d_turn, flags = eng.path_costs(dna)
            step_info = None
        cov, _ = eng.coverage(edge_ids)
        torch.cuda.synchronize()

        def time_stage(fn, reps):
            for _ in range(3):
                fn()
            torch.cuda.synchronize()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            for _ in range(reps):
                fn()
            b.record()
            torch.cuda.synchronize()
            return a.elapsed_time(b) / reps
        reps = max(args.steps, 5)
        ms_k2g = time_stage(lambda: eng.coverage(edge_ids, cov, flags), reps)
        ms_k2m = time_stage(lambda: eng.coverage_masked(step_info, cov, flags), reps) if fast else None
        ms_k1 = time_stage(lambda: eng.lib.gcp_path_costs(
            eng.ctx, dna.data_ptr(), P, edge_ids.data_ptr(),
            step_info.data_ptr() if fast else None, d_dist.data_ptr(), d_turn.data_ptr(),
            flags.data_ptr(), eng._stream()), reps)
        cov, _ = eng.coverage(edge_ids)
        ms_k3 = time_stage(lambda: eng.loop_closures(dna, d_dist, d_turn, cov, flags), reps)
        ms_general = time_stage(lambda: eng.evaluate(dna, detail=True, pixel_counts=True), 3)
        del edge_ids, step_info, cov
        peak, peak_src = measured_peak()
        ms_k2 = ms_k2m if fast else ms_k2g
        staged = {"path_costs": ms_k1, "coverage_masked": ms_k2m,
                  "coverage_general_all_counters": ms_k2g, "loop_closures": ms_k3,
                  "general_path_whole_evaluation": ms_general}
        tr = known_traffic(args.config + ("" if args.starts == "spread" else "-" + args.starts))
        tr = tr if isinstance(tr, dict) else {"dram_bytes": tr}
        # per launch of the profiled shape (65,536 organisms): scale to this launch's organisms
        scale = P / float(tr.get("organisms", 65536) or 65536)
        dram = tr.get("dram_bytes")
        dram = dram * scale if dram else None
        ms_kernel = ms_per_step if fast else ms_k2
        achieved = B_org * P / (ms_kernel * 1e-3) / 1e9
        roofline = {
            "bound": "issue/lsu" if fast else "hbm",
            "bound_source": "ncu --set full (profiles/): issue slots and the LSU data pipe are the busiest "
                            "units, DRAM ~1 % of peak; the byte model below is SURVEY 8d's HBM model",
            "ncu_unit_utilisation": {k: tr[k] for k in ("issue_active_pct", "l1tex_throughput_pct", "warps_active_pct",
                                                        "warp_instructions_per_organism") if k in tr} or None,
            "model_bound": "hbm",
            "kernel": ("k_eval_fused (path-cost gather + footprint bitset union/popcount + loop closures + "
                       "objectives, one CTA per organism; the only kernel of the timed step)") if fast else
                      "k_coverage (bitset union + popcount against map planes)",
            "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
            "traffic": dram, "peak_source": peak_src,
            "real_dram_frac": (dram / (ms_kernel * 1e-3) / 1e9 / peak) if dram else None,
            "algorithmic_bytes_per_organism": B_org, "organisms_per_launch": P,
            "ms_per_launch": ms_kernel,
            "l2_to_sm": {"bytes_per_organism": tr.get("quarters_per_organism", 1964) * 128,
                         "achieved_GBps": tr.get("quarters_per_organism", 1964) * 128 * P / (ms_kernel * 1e-3) / 1e9,
                         "what": "distinct 128-B footprint quarters gathered per organism (after the "
                                 "revisited-edge dedup) x organisms / launch time: the L2 -> SM stream "
                                 "the kernel really moves"},
            "note": "achieved = SURVEY 8d algorithmic bytes (one 512-B footprint tile + 32-B edge record per "
                    "executed step + DNA + result) / launch time; frac > 1 is possible because the model "
                    "charges HBM for tiles that the kernel gathers from the L2-resident store; `traffic` "
                    "is the measured DRAM bytes (ncu), `real_dram_frac` their share of the HBM peak",
            "staged_kernels_ms": staged,
            "general_path_frac": (B_org * P / (ms_general * 1e-3) / 1e9) / peak}
        # ---------------- ranking leg: maximin + arena order over 2P gladiators ---------
        obj2 = torch.cat([res.obj, res.obj.flip(0) * 1.0001])
        constr = -worlds.GEN_PARAMS_6['coverage_constr_0']
        ms_mm = time_stage(lambda: eng.maximin(obj2, constr), 3)
        fit = eng.maximin(obj2, constr)
        ms_ord = time_stage(lambda: eng.arena_order(fit), 3)
        n_gl = 2 * P
        extras["ranking"] = {"n_gladiators": n_gl, "maximin_ms": ms_mm, "arena_order_ms": ms_ord,
                             "algorithm": "exact O(N log N) sweep (DESIGN section 6) + stable radix order",
                             "n2_equivalent_pairs_per_s": n_gl * (n_gl - 1) / (ms_mm * 1e-3),
                             "note": "pairs/s = N (N - 1) ordered pairs of the reference's double loop "
                                     "divided by the time of the sweep that yields the same bits; not a "
                                     "flop rate"}
        del obj2, fit
        # ---------------- whole generations, the population sharded over the ranks -------
        # select -> crossover (parents read from peer shards over NVLink) -> mutate -> evaluate own
        # children -> all-gather of objective records -> maximin + order -> survivors fetched from
        # their owners.  Six clustered starts: agents that start next to each other meet, so the
        # population is not 100 % infeasible and ranking / selection see real objectives.
        gp = worlds.GEN_PARAMS_6
        cstarts = synth.clustered_starts(w.xy, w.row_ptr, A, cfg['size'])
        opc = worlds.org_params_for(cfg, cstarts)
        # the loop-closure minima of the reference's 6-agent column (2 solo per agent, 5 combined;
        # pather_driver.py:150-158) are tuned to its 391 x 261 building and are met by ~0.1 % of the
        # organisms on this map - with or without evolution (profiles/probe_r02b.jsonl).  The generation
        # legs therefore keep the connectivity constraint and one combined closure, which ~96 % of the
        # clustered walks and ~100 % of their descendants meet: ranking and selection see real objectives
        opc.update(min_solo_lcs=0, min_comb_lcs=1)

        def generation_leg(P_gen, n_gen=5, n_pre=40):
            per = P_gen // world_size
            evo = ShardedEvolution(eng, gp, P_gen)
            evo.load(population_slice(rank * per, (rank + 1) * per, cstarts, 5000), 0)
            feas_walks = float((evo.obj[:, 0] != 0).double().mean().item())
            # random walks rarely meet the loop-closure constraints (geneticalgorithm.py:197-200); the
            # timed generations run on a population that has evolved for n_pre generations, where the
            # penalty has made feasible organisms the majority - the state a real run spends its time in
            for g in range(n_pre):
                evo.step(4321, g)
            feas0 = float((evo.obj[:, 0] != 0).double().mean().item())
            torch.cuda.synchronize()
            barrier()
            g0, g1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            g0.record()
            for g in range(n_pre, n_pre + n_gen):
                evo.step(4321, g)
            g1.record()
            torch.cuda.synchronize()
            ms_gen = max_over_ranks(g0.elapsed_time(g1) / n_gen)
            eng.set_timing(True)
            evo.step(4321, n_pre + n_gen)
            torch.cuda.synchronize()
            names = ["path_costs", "eval_fused_or_coverage", "loop_closures", "maximin",
                     "arena_order", "select", "crossover", "mutate"]
            last = {n: eng.stage_ms(k) for k, n in enumerate(names)}
            eng.set_timing(False)
            evo.check()
            out = {"global_population": P_gen, "organisms_per_gpu": per, "n_gpus": world_size,
                   "ms_per_generation": ms_gen, "generations_per_s": 1000.0 / ms_gen,
                   "organisms_varied_and_evaluated_per_s": P_gen / ms_gen * 1e3,
                   "generations_timed": n_gen, "gene_bits": 8 * evo.gb,
                   "rank0_stage_ms_one_generation": {k: v for k, v in last.items() if v >= 0},
                   "generations_before_timing": n_pre,
                   "constraints": {"min_solo_lcs": opc['min_solo_lcs'], "min_comb_lcs": opc['min_comb_lcs'],
                                   "start_waypoints": [int(x) for x in cstarts]},
                   "mean_valid_genes_per_row_after": float((evo.parents != -1).double().sum(-1).mean().item()),
                   "feasible_fraction_random_walks": feas_walks,
                   "feasible_fraction_at_timing_start": feas0,
                   "feasible_fraction_after": float((evo.obj[:, 0] != 0).double().mean().item()),
                   "mean_constrained_coverage_after": float((-evo.obj[:, 0]).mean().item()),
                   "collectives_per_generation": "none" if world_size == 1 else
                   "1 NCCL all-gather of the children's objective records (%d B per rank) + 1 device "
                   "rendezvous through peer-memory flags; chromosomes: peer loads over NVLink only for "
                   "crossover parents and survivors owned by another rank" % (per * 16)}
            del evo
            return out
        eng.set_params(opc)
        extras["generation"] = generation_leg(cfg['pop'])
        extras["generation"]["includes"] = ("lowVarSample + crossover + mutation/muterpolate/prune + "
                                            "evaluation of P children + maximin over 2P + stable order + "
                                            "survivor truncation; no host round trip")
        if distributed:
            extras["generation_weak"] = generation_leg(cfg['pop'] * world_size)
        eng.set_params(op)
        # ---------------- end to end: host DNA in, host objectives out ------------------
        def e2e_leg(dna_h):
            obj_h = torch.empty((P, 2), dtype=torch.float64).pin_memory()
            flg_h = torch.empty((P, 4), dtype=torch.uint8).pin_memory()
            for _ in range(2):
                eng.evaluate_host(dna_h, obj_h, flg_h)
            barrier()
            t0 = time.perf_counter()
            n_e2e = max(3, min(args.steps, 10))
            for _ in range(n_e2e):
                eng.evaluate_host(dna_h, obj_h, flg_h)
            dt = max_over_ranks((time.perf_counter() - t0) / n_e2e)
            assert torch.equal(obj_h, res.obj.cpu()), "host path result differs from device path"
            return dt
        use16 = fast and w.N <= 32767
        dna_h32 = dna.cpu().pin_memory()
        # the library narrows int32 genes to int16 on the host where that pays (one rank per host, >= 6 worker threads)
        narrowed = bool(use16 and eng.lib.gcp_host_narrow_active(eng.ctx))
        dt32 = e2e_leg(dna_h32)                      # the call as a user makes it (library defaults)
        dt16 = e2e_leg(dna.to(torch.int16).cpu().pin_memory()) if use16 else None
        dt32_other = None
        if use16:                                    # the same call with the host-side narrowing forced the other way
            eng.set_option("host_narrow", 0 if narrowed else 1)
            dt32_other = e2e_leg(dna_h32)
            eng.set_option("host_narrow", -1)
        del dna_h32
        host_cores = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
        e2e = {"value": P_global / dt32, "unit": UNIT,
               "h2d_bytes_per_step": int(P * A * L * (2 if narrowed else 4)), "d2h_bytes_per_step": int(P * 20),
               "host_bytes_read_per_step": int(P * A * L * 4),
               "ms_per_step": dt32 * 1e3,
               "api": "FitnessEngine.evaluate_host -> gcp_eval_host (pinned host int32 chromosomes in, "
                      "objectives + flags out; chunked H2D / kernel / D2H on three streams), per rank "
                      "its own shard of the population" +
                      ("; the waypoint ids fit int16, so worker threads of the library narrow the int32 genes "
                       "on the way (inside the timed call): half the PCIe bytes; %d host cores for %d rank(s)"
                       % (host_cores, world_size) if narrowed else
                       ("; int32 genes on the wire: with %d ranks on this host (%d cores) the library does not "
                        "narrow them on the host (the GPUs' DMA engines together read host memory faster than "
                        "the cores convert it)" % (world_size, host_cores) if use16 else ""))}
        if dt32_other is not None:
            if narrowed:
                e2e["int32_on_the_wire"] = {"value": P_global / dt32_other, "ms_per_step": dt32_other * 1e3,
                                            "h2d_bytes_per_step": int(P * A * L * 4),
                                            "note": "gcp_set_option(host_narrow, 0): the int32 genes cross PCIe as they are"}
            else:
                e2e["narrowed_on_host"] = {"value": P_global / dt32_other, "ms_per_step": dt32_other * 1e3,
                                           "h2d_bytes_per_step": int(P * A * L * 2),
                                           "note": "gcp_set_option(host_narrow, 1): worker threads narrow the genes to int16 "
                                                   "inside the call (what a single rank per host does by default)"}
        if dt16 is not None:
            e2e["int16_genes"] = {"value": P_global / dt16, "ms_per_step": dt16 * 1e3,
                                  "h2d_bytes_per_step": int(P * A * L * 2),
                                  "note": "gcp_eval_host16: host chromosomes already narrowed to int16 "
                                          "(valid for N <= 32767 waypoints); the narrowing is NOT timed"}
        # ---------------- CPU baselines (rank 0, N=1 only) --------------------------------
        if rank == 0 and world_size == 1 and not args.no_cpu_baseline:
            # bounded sample: ~20 s of single-thread work (one organism takes 0.13 s at 4096^2 and
            # seconds at 16384^2, where the reference copies and reduces a 2 GB map per organism)
            t0 = time.time()
            cpu_port_rate(w, op, dna[:1].cpu().numpy(), 1)
            n_cpu = int(min(args.cpu_sample, max(4, 20.0 / max(time.time() - t0, 1e-3))))
            sample = dna[:n_cpu].cpu().numpy()
            rate, secs = cpu_port_rate(w, op, sample, 1)
            cpu_baseline = {"value": rate, "unit": UNIT, "cores": 1, "kind": "port",
                            "sample": "first %d organisms of the same population, "
                                      "Organism.calcObj restated in numpy/cv2 (oracle/"
                                      "fitness_oracle.py), single thread like the reference, "
                                      "%.1fs" % (len(sample), secs),
                            "host_cores_available": os.cpu_count()}
            # parity spot check of the timed population against the oracle
            import fitness_oracle as fo
            ow = oracle_world(w, op)
            full = eng.evaluate(dna[:4])
            for k in range(4):
                det = fo.calc_obj(ow, sample[k].astype(int), detail=True)
                o = full.obj[k].cpu().numpy()
                assert det['obj'][0] == o[0] and det['obj'][1] == o[1], "bench parity check failed"
            rk = cpu_ranking_baseline()
            n_gl = extras["ranking"]["n_gladiators"]
            rk["extrapolated_s_at_n_gladiators"] = rk["us_per_ordered_pair"] * 1e-6 * n_gl * (n_gl - 1)
            rk["n_gladiators"] = n_gl
            rk["gpu_speedup_vs_extrapolation"] = rk["extrapolated_s_at_n_gladiators"] / \
                ((extras["ranking"]["maximin_ms"] + extras["ranking"]["arena_order_ms"]) * 1e-3)
            rk["what"] = ("the reference's rackNstack double loop (oracle rack_n_stack_loop) timed at N = 256 / "
                          "512 / 1024 on one core and extrapolated by N^2 (the full N takes hours)")
            cpu_baseline["ranking"] = rk
            cg = cpu_generation_baseline()
            cpu_baseline["generation_P100_C1"] = cg
            c1 = gpu_c1_run(local)
            c1["cpu_port_generations_per_s"] = cg["generations_per_s"]
            c1["speedup_vs_cpu_port"] = c1["generations_per_s"] / cg["generations_per_s"]
            extras["config_C1"] = c1
            extras["generation"]["cpu_port_organisms_per_s_P100"] = cg["organisms_varied_and_evaluated_per_s"]

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world_size,
            "steps": args.steps, "warmup": n_warm, "ms_per_step": ms_per_step,
            "higher_is_better": True, "scaling": args.scaling, "vs_baseline": None,
            "dtype": "u64 bitsets + f64 costs", "data": "synthetic",
            "config": {"workload": workload_string(args.config, cfg, w.N, w.E, args.scaling, world_size) +
                       ("" if args.starts == "spread" else "; DIVERSE starts (every organism from its own random waypoints)"),
                       "organisms_per_gpu": P, "global_population": P_global,
                       "mean_executed_steps_per_organism": S_mean,
                       "l2": ("per-step input (dna, %.0f MB) exceeds the 126 MB L2; no flush" % (dna_bytes / 1e6))
                       if n_rot == 1 else
                       ("per-step input is %.0f MB < L2: the timed steps rotate over %d distinct populations "
                        "(%.0f MB in total) so that every step reads its chromosomes from HBM; no flush"
                        % (dna_bytes / 1e6, n_rot, n_rot * dna_bytes / 1e6)),
                       "l2_tables": ("the %.0f MB pre-masked footprint store is reused inside a launch and stays "
                                     "L2-resident by design" if pw.m_n_quarters * 128 < 0.8 * L2_BYTES else
                                     "the %.0f MB pre-masked footprint store does not fit the 126 MB L2: footprint "
                                     "gathers that leave the neighbourhood of the start waypoints come from HBM")
                       % (pw.m_n_quarters * 128 / 1e6),
                       "parallelism": "population sharded over the ranks, tables replicated, no collective "
                                      "in the evaluation step",
                       "host_numa_node_rank0": numa_node},
            "gpu_launches": int(launches), "clocks": clocks,
        }
        if roofline:
            line["roofline"] = roofline
        if cpu_baseline:
            line["cpu_baseline"] = cpu_baseline
        if e2e:
            line["e2e"] = e2e
        line.update(extras)
        print(json.dumps(line), flush=True)
    if distributed:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
