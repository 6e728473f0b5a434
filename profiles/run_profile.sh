#!/bin/bash
# Run on the GPU box (under gpurun) from the repo root:  bash profiles/run_profile.sh [tag]
# 1) plain bench (the number), 2) ncu launch list of our kernels, 3) ncu --set full of the
# coverage kernel.  Profiler runs never produce bench numbers.
set -u
TAG=${1:-r01}
OUT=gpurun_out
mkdir -p $OUT
CMD="python bench.py --steps 2 --warmup 3 --no-cpu-baseline"
python bench.py --steps 10 --warmup 3 > $OUT/bench_${TAG}.json 2> $OUT/bench_${TAG}.err || { echo "bench failed"; tail -20 $OUT/bench_${TAG}.err; exit 1; }
tail -1 $OUT/bench_${TAG}.json
$CMD > $OUT/plain_${TAG}.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:^k_ -c 200 --csv \
    --log-file $OUT/launches_${TAG}.csv $CMD > $OUT/ncu_launches_${TAG}.log 2>&1
echo "launch list rc=$?"
for K in ${KERNELS:-k_coverage_masked}; do
$CMD > $OUT/plain2_${TAG}.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:$K -s 3 -c 1 \
    -o $OUT/prof_${K}_${TAG} -f $CMD > $OUT/ncu_full_${K}_${TAG}.log 2>&1
echo "full capture $K rc=$?"
done
ls -la $OUT
