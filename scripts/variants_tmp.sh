set -x
python -m pytest tests/test_gpu_parity.py -x -q -k "batched or full_size or bit_identical" > gpurun_out/pytest_v1.log 2>&1; tail -2 gpurun_out/pytest_v1.log
b() { python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-lbfgs 2>/dev/null | python -c "
import json,sys;d=json.loads(sys.stdin.read());print('$1', d['value'], d['kernel_ms'])"; }
b ilp2_cta3
SVSDF_EXTRA_NVCC_FLAGS="-DSVSDF_OUTER_MIN_CTAS=2" python -m implicit_svsdf_planner_b200.build --force > /dev/null 2>&1
b ilp2_cta2
SVSDF_EXTRA_NVCC_FLAGS="-DSVSDF_OUTER_MIN_CTAS=2 -DSVSDF_ENGINE_ILP=1" python -m implicit_svsdf_planner_b200.build --force > /dev/null 2>&1
b ilp1_cta2
SVSDF_EXTRA_NVCC_FLAGS="-DSVSDF_OUTER_MIN_CTAS=4" python -m implicit_svsdf_planner_b200.build --force > /dev/null 2>&1
b ilp2_cta4
