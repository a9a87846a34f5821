# usage: bash tools/r02_finalN.sh N   (run under gpurun --gpus N)
N=$1
mkdir -p gpurun_out
IGCN_PEER_TIMEOUT_S=30 timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/r02_bench_${N}gpu.json 2> gpurun_out/r02_bench_${N}gpu.err; echo "bench$N rc=$?"
tail -c 200 gpurun_out/r02_bench_${N}gpu.json
