mkdir -p gpurun_out
for order in degree typed; do IGCN_ROW_ORDER=$order timeout 300 python tools/time_step.py yelp-lightgcn 2>/dev/null | tail -1 | sed "s/^/order=$order /" | tee -a gpurun_out/r02_roworder.log; done
for order in degree typed; do IGCN_ROW_ORDER=$order timeout 300 python tools/time_step.py gowalla-igcn 2>/dev/null | tail -1 | sed "s/^/order=$order /" | tee -a gpurun_out/r02_roworder.log; done
IGCN_PEER_TIMEOUT_S=30 timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 4 --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r02_bench4.json 2> gpurun_out/r02_bench4.err; echo "bench4 rc=$?"
tail -c 300 gpurun_out/r02_bench4.json
