mkdir -p gpurun_out
IGCN_PEER_TIMEOUT_S=30 timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 8 --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r02_bench8.json 2> gpurun_out/r02_bench8.err; echo "bench8 rc=$?"
tail -c 300 gpurun_out/r02_bench8.json
IGCN_PEER_TIMEOUT_S=30 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29512 tests/dist_worker.py gpu > gpurun_out/r02_dist8.log 2>&1; echo "dist8 rc=$?"
tail -3 gpurun_out/r02_dist8.log | cut -c1-300
