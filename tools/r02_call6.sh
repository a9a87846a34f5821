mkdir -p gpurun_out
IGCN_PEER_TIMEOUT_S=30 timeout 900 python -m pytest tests/test_dist.py -q -x > gpurun_out/r02_dist.log 2>&1; echo "rc=$?" >> gpurun_out/r02_dist.log
tail -15 gpurun_out/r02_dist.log | cut -c1-400
IGCN_PEER_TIMEOUT_S=30 timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r02_bench2b.json 2> gpurun_out/r02_bench2b.err; echo "bench2 rc=$?"
tail -c 300 gpurun_out/r02_bench2b.json; grep -v "^\*\|OMP_NUM\|^$" gpurun_out/r02_bench2b.err | tail -8 | cut -c1-300
