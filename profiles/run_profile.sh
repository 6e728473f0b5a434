#!/bin/bash
# Run on the GPU box (under gpurun) from the repo root:  bash profiles/run_profile.sh [tag]
#   1) plain bench (the number; never taken under a profiler)
#   2) ncu launch list of the TIMED STEP only (--no-extras: warm-up + steps of gcp_eval)
#   3) ncu --set full of each kernel in $KERNELS (one launch each, after warm-up)
set -u
TAG=${1:-r01}
OUT=gpurun_out
mkdir -p $OUT
STEP="python bench.py --steps 4 --warmup 3 --no-extras"
FULL="python bench.py --steps 2 --warmup 3 --no-cpu-baseline"
python bench.py --steps 10 --warmup 3 > $OUT/bench_${TAG}.json 2> $OUT/bench_${TAG}.err || { echo "bench failed"; tail -20 $OUT/bench_${TAG}.err; exit 1; }
tail -c 600 $OUT/bench_${TAG}.json
$STEP > $OUT/plain_${TAG}.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:^k_ -c 400 --csv \
    --log-file $OUT/launches_${TAG}.csv $STEP > $OUT/ncu_launches_${TAG}.log 2>&1
echo "launch list rc=$?"
for K in ${KERNELS:-k_eval_fused}; do
$FULL > $OUT/plain2_${TAG}.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:$K -s ${SKIP:-3} -c 1 \
    -o $OUT/prof_${K}_${TAG} -f $FULL > $OUT/ncu_full_${K}_${TAG}.log 2>&1
echo "full capture $K rc=$?"
done
ls -la $OUT | tail -20
