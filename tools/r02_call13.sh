mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_tc.py tests/test_gpu_topk.py tests/test_gpu_parity.py tests/test_gpu_oracle_small.py tests/test_zz_gpu_fullsize.py -q -x > gpurun_out/r02_gputests_9.log 2>&1; echo "rc=$?" >> gpurun_out/r02_gputests_9.log
tail -12 gpurun_out/r02_gputests_9.log | cut -c1-300
for w in yelp-lightgcn amazon-igcn gowalla-igcn; do timeout 300 python tools/tc_floor.py $w 0 3 2 5 2>/dev/null | grep -E "popularity" | tee -a gpurun_out/r02_tc_floor_e.log; done
