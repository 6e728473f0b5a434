#!/bin/bash
# Round 2, final kernels, on the GPU box (under gpurun) from the repo root: bash profiles/run_profile_r02z.sh
#   1) default bench line (1 GPU), then the ncu launch list of the same command's timed step
#   2) ncu --set full of the general fused kernel (all pixel counters, C3)
#   3) C5 with DIVERSE starts (footprint store beyond the L2): bench line, then ncu --set full
set -u
TAG=${1:-r02z}
OUT=gpurun_out
mkdir -p $OUT
python bench.py > $OUT/bench_${TAG}_1gpu.json 2> $OUT/bench_${TAG}_1gpu.err
echo "bench rc=$?"
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:^k_ -c 400 --csv \
    --log-file $OUT/launches_${TAG}.csv python bench.py --steps 2 --warmup 3 --no-extras --no-cpu-baseline > $OUT/ncu_${TAG}.log 2>&1
echo "launch list rc=$?"
python profiles/general_step.py 16384 3 > $OUT/general_step_${TAG}.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:k_eval_fused -s 2 -c 1 \
    -o $OUT/prof_k_eval_fused_general_${TAG} -f python profiles/general_step.py 16384 3 > $OUT/ncu_general_${TAG}.log 2>&1
echo "general capture rc=$?"
C5="python bench.py --config C5 --starts diverse --pop 65536 --steps 5 --warmup 3"
$C5 --no-cpu-baseline > $OUT/bench_c5_diverse_${TAG}.json 2> $OUT/bench_c5_diverse_${TAG}.err
echo "C5 diverse bench rc=$?"
ncu --set full --clock-control none --import-source on -k regex:k_eval_fused -s 4 -c 1 \
    -o $OUT/prof_k_eval_fused_c5_diverse_${TAG} -f $C5 --no-extras --no-cpu-baseline > $OUT/ncu_full_c5_diverse_${TAG}.log 2>&1
echo "C5 diverse full capture rc=$?"
ls -la $OUT | tail -8
