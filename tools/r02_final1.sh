# final single-GPU validation of the round: tests, smoke, bench (both arms), the other workloads, ncu traffic of the default workload
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -q > gpurun_out/r02_final_gputests.log 2>&1; echo "rc=$?" >> gpurun_out/r02_final_gputests.log
tail -4 gpurun_out/r02_final_gputests.log | cut -c1-300
timeout 300 python __graft_entry__.py smoke > gpurun_out/r02_final_smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/r02_final_smoke.log | cut -c1-300
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/r02_bench_1gpu.json 2> gpurun_out/r02_bench_1gpu.err; echo "bench rc=$?"
timeout 900 python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/r02_bench_reference.json 2> gpurun_out/r02_bench_reference.err; echo "reference rc=$?"
timeout 600 python bench.py --workload dropui-gowalla > gpurun_out/r02_dropui_gowalla.json 2> gpurun_out/r02_dropui.err; echo "dropui rc=$?"; tail -c 400 gpurun_out/r02_dropui_gowalla.json
timeout 600 python bench.py --workload scaleout-mid --steps 3 --warmup 1 > gpurun_out/r02_scaleout_mid.json 2> gpurun_out/r02_scaleout_mid.err; echo "scaleout-mid rc=$?"; tail -c 600 gpurun_out/r02_scaleout_mid.json
python tools/prof_step.py amazon-igcn 3 > gpurun_out/r02_plain_step_amazon.log 2>&1 &&
ncu --set full --clock-control none -k regex:"prop_kernel" -s 9 -c 9 -o gpurun_out/r02_prof_step_amazon python tools/prof_step.py amazon-igcn 3 > gpurun_out/r02_ncu4.log 2>&1
echo "amazon step set rc=$?"
python tools/prof_eval.py amazon-igcn > gpurun_out/r02_plain_eval_amazon.log 2>&1 &&
ncu --set full --clock-control none -k regex:"score_tc" -s 1 -c 1 -o gpurun_out/r02_prof_eval_amazon python tools/prof_eval.py amazon-igcn > gpurun_out/r02_ncu5.log 2>&1
echo "amazon eval set rc=$?"
