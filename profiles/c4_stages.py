import os, sys, torch
sys.path.insert(0, "/root/repo")
from genetic_coverage_planning_b200 import synth, worlds
from genetic_coverage_planning_b200.engine import FitnessEngine
w, pw = worlds.build_synth_world(4096, seed=0)
for name in ([a for a in sys.argv[1:]] or ["C4-16", "C4-32"]):
    cfg = worlds.CONFIGS[name]
    A, L, P = cfg['agents'], cfg['L'], 4096
    starts = synth.spread_starts(w.xy, A, 4096)
    op = worlds.org_params_for(cfg, starts)
    eng = FitnessEngine(pw, op, num_agents=A)
    for kv in filter(None, os.environ.get("GCP_OPTS", "").split(",")):
        k, v = kv.split("=")
        eng.set_option(k, int(v))
    dna = synth.random_walk_population_torch(w.xy, w.row_ptr, w.col, starts, P, cfg['walk'], L, seed=1, device=eng.device)
    for _ in range(2): eng.evaluate(dna, detail=False)
    eng.set_timing(True)
    eng.evaluate(dna, detail=False)
    torch.cuda.synchronize()
    print(os.environ.get("GCP_OPTS", ""), name, "P", P, {n: round(eng.stage_ms(k), 3) for k, n in enumerate(["path", "cov", "lc"])})
    eng.set_timing(False)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(5): eng.evaluate(dna, detail=False)
    b.record(); torch.cuda.synchronize()
    print("   whole evaluation %.3f ms" % (a.elapsed_time(b) / 5))
