#!/bin/bash
# C5 with DIVERSE starts (footprint store beyond the L2): evaluation step under the prefetch modes
# (GCP_OPTS=cov_prefetch=0 none, 1 while sorting (B1), 2 look-ahead in B2, 3 both) and build variants
# (GCP_LIB=<tag>: look-ahead distance, A1 unroll).  bash profiles/exp_c5_prefetch.sh
run() {  # lib pf starts
  GCP_LIB=$1 GCP_OPTS=cov_prefetch=$2 python bench.py --config C5 --starts $3 --pop 65536 --steps 5 --warmup 3 --no-extras --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print(json.dumps({'lib':'$1','cov_prefetch':$2,'starts':'$3','ms_per_step':d['ms_per_step'],'organisms_per_s':d['value']}))
"
}
run "" 2 diverse
run pf32 2 diverse
run pf8 2 diverse
run u2 2 diverse
run u2pf32 2 diverse
run u2 0 diverse
run u2 2 spread
