#!/bin/bash
# Round 2, on the GPU box (under gpurun) from the repo root: bash profiles/run_profile_r02.sh [tag]
#   1) ncu launch lists of whole generations at P = 65,536 and P = 8,192 (strong-scaling shard)
#   2) config C5 with DIVERSE starts (the footprint store outgrows the L2): bench line, then
#      ncu --set full of the evaluation kernel on that workload (real DRAM traffic / throughput)
set -u
TAG=${1:-r02}
OUT=gpurun_out
mkdir -p $OUT
for P in 65536 8192; do
  python profiles/gen_step.py $P 3 40 > $OUT/gen_plain_${P}_${TAG}.log 2>&1 &&
  ncu --metrics gpu__time_duration.sum --clock-control none -k regex:^k_ --csv \
      --log-file $OUT/launches_gen_${P}_${TAG}.csv python profiles/gen_step.py $P 2 3 > $OUT/ncu_gen_${P}_${TAG}.log 2>&1
  echo "generation launch list P=$P rc=$?"; cat $OUT/gen_plain_${P}_${TAG}.log | tail -1
done
C5="python bench.py --config C5 --starts diverse --pop 65536 --steps 5 --warmup 3"
$C5 --no-cpu-baseline > $OUT/bench_c5_diverse_${TAG}.json 2> $OUT/bench_c5_diverse_${TAG}.err
echo "C5 diverse bench rc=$?"; tail -3 $OUT/bench_c5_diverse_${TAG}.err
python bench.py --config C5 --pop 65536 --steps 5 --warmup 3 --no-extras > $OUT/bench_c5_spread_${TAG}.json 2> $OUT/bench_c5_spread_${TAG}.err
echo "C5 spread bench rc=$?"
ncu --set full --clock-control none --import-source on -k regex:k_eval_fused -s 4 -c 1 \
    -o $OUT/prof_k_eval_fused_c5_diverse_${TAG} -f $C5 --no-extras > $OUT/ncu_full_c5_diverse_${TAG}.log 2>&1
echo "C5 diverse full capture rc=$?"
ls -la $OUT | tail -12
