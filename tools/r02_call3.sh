mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/r02_gputests_2.log 2>&1; echo "rc=$?" >> gpurun_out/r02_gputests_2.log
tail -5 gpurun_out/r02_gputests_2.log
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/r02_bench1.json 2> gpurun_out/r02_bench1.err; echo "bench rc=$?"
tail -c 600 gpurun_out/r02_bench1.json
for w in yelp-lightgcn gowalla-igcn; do timeout 300 python tools/tc_floor.py $w 2>/dev/null | grep -E "variant|stats" | tee -a gpurun_out/r02_tc_floor.log; done
