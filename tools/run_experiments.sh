#!/bin/bash
# One gpurun call that validates and times the opt-in experiments (everything is wrapped in its own timeout):
#   /usr/local/graft/bin/gpurun --timeout 900 -- 'bash tools/run_experiments.sh'
# Outputs land in gpurun_out/exp_*.log.
cd "$(dirname "$0")/.." || exit 1
mkdir -p gpurun_out
echo "== parity of the gated experimental kernels"
IGCN_EXPERIMENTAL=1 timeout 300 python -m pytest tests/test_gpu_step_kernels.py tests/test_gpu_tc.py -x -q \
    -k "hot_row_staging or threshold_in_mma" > gpurun_out/exp_parity.log 2>&1
echo "rc=$?" >> gpurun_out/exp_parity.log
tail -4 gpurun_out/exp_parity.log
echo "== full-size property tests"
timeout 300 python -m pytest tests/test_zz_gpu_fullsize.py -x -q > gpurun_out/exp_fullsize.log 2>&1
echo "rc=$?" >> gpurun_out/exp_fullsize.log
tail -4 gpurun_out/exp_fullsize.log
echo "== scoring kernel: production vs threshold-in-MMA (variant 7)"
timeout 200 python tools/tc_floor.py yelp-lightgcn 0 7 0 7 2>&1 | grep variant | tee gpurun_out/exp_tc7.log
echo "== training step: production vs hot rows staged in shared memory"
for hot in 0 768; do
    IGCN_SPMM_HOT=$hot timeout 200 python bench.py --steps 300 --warmup 20 --no-cpu-baseline 2> gpurun_out/exp_hot_$hot.err |
        python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('IGCN_SPMM_HOT=$hot ms/step', d['ms_per_step'], 'spmm frac', d['roofline']['frac'], 'eval ms', d['eval']['ms'])" |
        tee -a gpurun_out/exp_hot.log
done
