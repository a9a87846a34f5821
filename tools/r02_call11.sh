mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_step_kernels.py tests/test_gpu_tc.py tests/test_gpu_topk.py tests/test_gpu_parity.py tests/test_gpu_oracle_small.py -q -x > gpurun_out/r02_gputests_8.log 2>&1; echo "rc=$?" >> gpurun_out/r02_gputests_8.log
tail -6 gpurun_out/r02_gputests_8.log | cut -c1-300
timeout 900 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r02_bench1d.json 2> gpurun_out/r02_bench1d.err; echo "bench1 rc=$?"
IGCN_PEER_TIMEOUT_S=30 timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 4 --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r02_bench4b.json 2> gpurun_out/r02_bench4b.err; echo "bench4 rc=$?"
