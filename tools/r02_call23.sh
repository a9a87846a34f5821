mkdir -p gpurun_out
for l in 8 4; do for w in yelp-lightgcn gowalla-igcn amazon-igcn; do IGCN_SPMM_LANES=$l timeout 300 python tools/time_step.py $w 2>/dev/null | tail -1 | sed "s/^/lanes=$l /" | tee -a gpurun_out/r02_lanes.log; done; done
IGCN_SPMM_LANES=4 timeout 300 python -m pytest tests/test_gpu_step_kernels.py tests/test_gpu_parity.py -q -x > gpurun_out/r02_gputests_13.log 2>&1; echo "rc=$?" >> gpurun_out/r02_gputests_13.log; tail -3 gpurun_out/r02_gputests_13.log | cut -c1-200
