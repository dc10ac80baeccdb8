#!/bin/bash
# Run on the GPU box (under gpurun): collects the evidence that scripts/summarise_profiles.py turns into profiles/.
# usage: collect_profiles.sh [a|b]   a = tests, bench, launch list, configs, batch; b = the two `ncu --set full` captures
# (gpurun merges at most 64 MiB back per call, the two reports alone are ~60 MB)
set -x
mkdir -p gpurun_out
PART=${1:-ab}
if [[ $PART == *a* ]]; then
python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu_r2.log 2>&1
python bench.py --steps 20 --warmup 3 > gpurun_out/bench_r2.json 2> gpurun_out/bench_r2.err
python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/bench_reference_r2.json 2>> gpurun_out/bench_r2.err
# launch list of the bench command (per-launch times under ncu are cold-cache and serialised: compare shares)
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_r2.csv \
    python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-lbfgs > gpurun_out/bench_under_ncu_r2.log 2>&1
fi
if [[ $PART == *b* ]]; then
# one full capture of the dominant kernel and of the interior-branch kernel
ncu --set full --clock-control none --import-source on -k regex:k_outer -s 2 -c 1 -f -o gpurun_out/prof_outer_r2 \
    python scripts/prof_step.py > gpurun_out/prof_r2.log 2>&1
ncu -i gpurun_out/prof_outer_r2.ncu-rep --page raw --csv > gpurun_out/prof_outer_r2_raw.csv 2>/dev/null
ncu -i gpurun_out/prof_outer_r2.ncu-rep --page source --csv > gpurun_out/prof_outer_r2_source.csv 2>/dev/null
ncu --set full --clock-control none -k regex:k_gsip -s 2 -c 1 -f -o gpurun_out/prof_gsip_r2 \
    python scripts/prof_step.py >> gpurun_out/prof_r2.log 2>&1
ncu -i gpurun_out/prof_gsip_r2.ncu-rep --page raw --csv > gpurun_out/prof_gsip_r2_raw.csv 2>/dev/null
# the mesh functor's k_outer (config 4m scene at 20 000 points)
SVSDF_SHAPE=sdHorseshoe SVSDF_N=16 SVSDF_MESH=1 SVSDF_P=20000 ncu --set full --clock-control none -k regex:k_outer -s 2 -c 1 -f \
    -o gpurun_out/prof_outer_mesh_r2 python scripts/prof_step.py >> gpurun_out/prof_r2.log 2>&1
ncu -i gpurun_out/prof_outer_mesh_r2.ncu-rep --page raw --csv > gpurun_out/prof_outer_mesh_r2_raw.csv 2>/dev/null
fi
if [[ $PART == *b* ]]; then
# gpurun merges at most 64 MiB back: keep the CSV exports (raw metrics, SASS source page, details), drop the 40 MB reports
for t in outer_r2 gsip_r2 outer_mesh_r2; do
  ncu -i gpurun_out/prof_$t.ncu-rep --page details --csv > gpurun_out/prof_${t}_details.csv 2>/dev/null
  rm -f gpurun_out/prof_$t.ncu-rep
done
fi
if [[ $PART == *a* ]]; then
python tests/tools/run_configs.py 1 2 3 4 4m > gpurun_out/configs_r2.jsonl 2> gpurun_out/configs_r2.err
python scripts/run_batch.py --problems 4096 --first 96 > gpurun_out/batch_1gpu_r2.json 2> gpurun_out/batch_r2.err
tail -3 gpurun_out/pytest_gpu_r2.log
fi
