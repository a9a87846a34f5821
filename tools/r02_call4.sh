mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_tc.py tests/test_gpu_topk.py tests/test_gpu_parity.py -q > gpurun_out/r02_gputests_3.log 2>&1; echo "rc=$?" >> gpurun_out/r02_gputests_3.log
tail -4 gpurun_out/r02_gputests_3.log
timeout 300 python tools/tc_floor.py yelp-lightgcn 2>/dev/null | grep -E "variant|stats" | tee gpurun_out/r02_tc_floor_b.log
IGCN_PEER_TIMEOUT_S=30 timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/r02_bench2.json 2> gpurun_out/r02_bench2.err; echo "bench2 rc=$?"
tail -c 400 gpurun_out/r02_bench2.json; tail -5 gpurun_out/r02_bench2.err
