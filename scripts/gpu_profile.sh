#!/bin/bash
# Runs on the GPU box (under gpurun): GPU parity tests, the default bench, the ncu launch list
# and one `ncu --set full` capture of the dominant kernels.  Outputs land in gpurun_out/.
# usage: scripts/gpu_profile.sh <tag> [kernel-regex]
set -u
TAG=${1:-r1}
KRE=${2:-"k_sweep|k_finish_rows|k_keys|k_mtf_small|k_huff"}
O=gpurun_out
mkdir -p $O
python -m pytest tests -m gpu -x -q > $O/pytest_$TAG.log 2>&1; echo "pytest rc=$?"
tail -3 $O/pytest_$TAG.log
python bench.py > $O/bench_$TAG.log 2>&1; echo "bench rc=$?"
python bench.py --impl reference --steps 1 --warmup 0 > $O/bench_ref_$TAG.log 2>&1; echo "ref rc=$?"
CMD="python bench.py --lines 10000000 --steps 1 --warmup 1 --no-cpu-baseline"
$CMD > $O/plain_$TAG.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/launches_$TAG.csv $CMD > $O/ncu_l_$TAG.log 2>&1
echo "launch list rc=$?"
CMD2="python bench.py --lines 10000000 --steps 1 --warmup 1 --no-cpu-baseline"
$CMD2 > $O/plain2_$TAG.log 2>&1 &&
timeout 600 ncu --set full --clock-control none --import-source on -k "regex:$KRE" -s 9 -c 9 -f -o $O/prof_$TAG $CMD2 > $O/ncu_f_$TAG.log 2>&1
echo "ncu full rc=$?"
