#!/bin/bash
# Runs on the GPU box (under gpurun): GPU parity tests, the default bench, the ncu launch list
# and one `ncu --set full` capture of the dominant kernels.  Outputs land in gpurun_out/.
# usage: scripts/gpu_profile.sh <tag> [kernel-regex]
set -u
TAG=${1:-r1}
KRE=${2:-"k_sweep|k_finish_rows|k_keys|k_mtf_small|k_huff|k_front|k_rle_write"}
O=gpurun_out
mkdir -p $O
python -m pytest tests -m gpu -x -q > $O/pytest_$TAG.log 2>&1; echo "pytest rc=$?"
tail -3 $O/pytest_$TAG.log
python bench.py > $O/bench_$TAG.log 2>&1; echo "bench rc=$?"
python bench.py --impl reference --steps 1 --warmup 0 > $O/bench_ref_$TAG.log 2>&1; echo "ref rc=$?"
for c in 1 3 4; do python bench.py --cfg $c --no-cpu-baseline > $O/bench_${TAG}_cfg$c.log 2>&1; echo "cfg$c rc=$?"; done
CMD="python bench.py --lines 10000000 --steps 1 --warmup 1 --no-cpu-baseline --no-parity"
$CMD > $O/plain_$TAG.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/launches_$TAG.csv $CMD > $O/ncu_l_$TAG.log 2>&1
echo "launch list rc=$?"
$CMD > $O/plain2_$TAG.log 2>&1 &&
timeout 900 ncu --set full --clock-control none --import-source on -k "regex:$KRE" -s 12 -c 14 -f -o $O/prof_$TAG $CMD > $O/ncu_f_$TAG.log 2>&1
echo "ncu full rc=$?"
