mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_step_kernels.py tests/test_gpu_parity.py tests/test_gpu_oracle_small.py tests/test_zz_gpu_fullsize.py -q -x > gpurun_out/r02_gputests_15.log 2>&1; echo "rc=$?" >> gpurun_out/r02_gputests_15.log; tail -4 gpurun_out/r02_gputests_15.log | cut -c1-250
for l in 2 4; do IGCN_SPMM_L8=$l timeout 200 python tools/time_narrow.py yelp 2>/dev/null | grep "us per layer" | tee -a gpurun_out/r02_narrow.log; done
timeout 200 python tools/time_narrow.py amazon 2>/dev/null | grep "us per layer" | tee -a gpurun_out/r02_narrow.log
for w in yelp-lightgcn gowalla-igcn amazon-igcn; do timeout 300 python tools/time_step.py $w 2>/dev/null | tail -1 | tee -a gpurun_out/r02_narrow.log; done
