mkdir -p gpurun_out
IGCN_PEER_TIMEOUT_S=30 timeout 1800 python -m pytest tests -m gpu -q > gpurun_out/r02_gputests_7.log 2>&1; echo "rc=$?" >> gpurun_out/r02_gputests_7.log
tail -6 gpurun_out/r02_gputests_7.log | cut -c1-300
timeout 900 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r02_bench1c.json 2> gpurun_out/r02_bench1c.err; echo "bench1 rc=$?"
IGCN_PEER_TIMEOUT_S=30 timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r02_bench2c.json 2> gpurun_out/r02_bench2c.err; echo "bench2 rc=$?"
python tools/prof_eval.py yelp-lightgcn > gpurun_out/r02_plain_eval.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:"score_tc|tc_finalize" -s 2 -c 2 -o gpurun_out/r02_prof_eval_yelp python tools/prof_eval.py yelp-lightgcn > gpurun_out/r02_ncu3.log 2>&1
echo "eval set rc=$?"
