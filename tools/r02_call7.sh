mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_tc.py tests/test_gpu_topk.py tests/test_gpu_parity.py tests/test_gpu_step_kernels.py -q -x > gpurun_out/r02_gputests_5.log 2>&1; echo "rc=$?" >> gpurun_out/r02_gputests_5.log
tail -4 gpurun_out/r02_gputests_5.log | cut -c1-300
for w in yelp-lightgcn amazon-igcn gowalla-igcn; do timeout 300 python tools/tc_floor.py $w 0 3 2 5 2>/dev/null | grep -E "popularity" | tee -a gpurun_out/r02_tc_floor_d.log; done
IGCN_TC_CLUSTER=1 timeout 300 python tools/tc_floor.py yelp-lightgcn 0 5 2>/dev/null | grep -E "popularity" | sed 's/^/unpaired /' | tee -a gpurun_out/r02_tc_floor_d.log
IGCN_PEER_TIMEOUT_S=30 timeout 900 python -m pytest tests/test_dist.py -q -x > gpurun_out/r02_dist2.log 2>&1; echo "rc=$?" >> gpurun_out/r02_dist2.log
tail -5 gpurun_out/r02_dist2.log | cut -c1-400
