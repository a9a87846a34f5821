mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_step_kernels.py -q -x -k narrow > gpurun_out/r02_gputests_14.log 2>&1; echo "rc=$?" >> gpurun_out/r02_gputests_14.log; tail -3 gpurun_out/r02_gputests_14.log | cut -c1-200
IGCN_PEER_TIMEOUT_S=30 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 4 --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r02_bench4e.json 2> gpurun_out/r02_bench4e.err; echo "bench4 rc=$?"
