mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_step_kernels.py -q -x -k narrow > gpurun_out/r02_gputests_16.log 2>&1; echo "rc=$?" >> gpurun_out/r02_gputests_16.log; tail -3 gpurun_out/r02_gputests_16.log | cut -c1-200
timeout 200 python tools/time_narrow.py yelp 2>/dev/null | grep "us per layer" | tee gpurun_out/r02_narrow3.log
timeout 200 python tools/time_narrow.py amazon 2>/dev/null | grep "us per layer" | tee -a gpurun_out/r02_narrow3.log
