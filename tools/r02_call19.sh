mkdir -p gpurun_out
timeout 200 python -m pytest tests/test_gpu_tc.py -q -x > gpurun_out/r02_gputests_11.log 2>&1; echo "rc=$?" >> gpurun_out/r02_gputests_11.log; tail -3 gpurun_out/r02_gputests_11.log | cut -c1-200
timeout 120 python tools/tc_floor.py yelp-lightgcn 5 3 0 2>/dev/null | grep -E "popularity.*variant" | tee gpurun_out/r02_tc_floor_i.log
