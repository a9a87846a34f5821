mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r02_gputests_1.log 2>&1; echo "rc=$?" >> gpurun_out/r02_gputests_1.log
tail -5 gpurun_out/r02_gputests_1.log
for hot in 0 768 1600; do IGCN_SPMM_HOT=$hot timeout 300 python tools/time_step.py yelp-lightgcn 2>/dev/null | tail -1 | tee -a gpurun_out/r02_hot.log; done
for hot in 0 768; do IGCN_SPMM_HOT=$hot timeout 300 python tools/time_step.py amazon-igcn 2>/dev/null | tail -1 | tee -a gpurun_out/r02_hot.log; done
