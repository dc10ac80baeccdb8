#!/bin/bash
# Run on the GPU box (under gpurun): one `ncu --set full` capture of a kernel of the config-2 step, plus its SASS-level
# source page as CSV (read back here with scripts/ncu_by_line.py against `nvdisasm -g` of the shipped cubin).
# usage: profile_kernel.sh <kernel-regex> <tag> [env assignments for scripts/prof_step.py ...]
set -x
K=$1; TAG=$2; shift 2
mkdir -p gpurun_out
env "$@" ncu --set full --clock-control none --import-source on -k regex:$K -s 2 -c 1 -f -o gpurun_out/prof_${TAG} \
    python scripts/prof_step.py > gpurun_out/prof_${TAG}.log 2>&1
ncu -i gpurun_out/prof_${TAG}.ncu-rep --page source --csv > gpurun_out/prof_${TAG}_source.csv 2>/dev/null
ncu -i gpurun_out/prof_${TAG}.ncu-rep --page raw --csv > gpurun_out/prof_${TAG}_raw.csv 2>/dev/null
tail -2 gpurun_out/prof_${TAG}.log
