mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_tc.py tests/test_gpu_topk.py -q > gpurun_out/r02_gputests_4.log 2>&1; echo "rc=$?" >> gpurun_out/r02_gputests_4.log
tail -4 gpurun_out/r02_gputests_4.log
for w in yelp-lightgcn amazon-igcn gowalla-igcn; do timeout 300 python tools/tc_floor.py $w 0 3 2 5 2>/dev/null | grep -E "popularity" | tee -a gpurun_out/r02_tc_floor_c.log; done
