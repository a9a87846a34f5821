mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/r02_gputests_6.log 2>&1; echo "rc=$?" >> gpurun_out/r02_gputests_6.log
tail -6 gpurun_out/r02_gputests_6.log | cut -c1-300
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/r02_bench1b.json 2> gpurun_out/r02_bench1b.err; echo "bench rc=$?"
# ---- ncu: launch list of 3 eager IGCN steps + one evaluation (gowalla shape), then full sets of one step's kernels and of the eval kernels
python tools/prof_step.py gowalla-igcn 3 > gpurun_out/r02_plain_step.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_launches_step_gowalla-igcn.csv python tools/prof_step.py gowalla-igcn 3 > gpurun_out/r02_ncu1.log 2>&1
echo "launch list rc=$?"
ncu --set full --clock-control none -k regex:"prop_kernel|colsum|plan_fast|bpr_|adam|sample_kernel|loss_fin|dw_stage" -s 40 -c 32 -o gpurun_out/r02_prof_step_gowalla python tools/prof_step.py gowalla-igcn 3 > gpurun_out/r02_ncu2.log 2>&1
echo "step set rc=$?"
python tools/prof_eval.py yelp-lightgcn > gpurun_out/r02_plain_eval.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:"score_tc|tc_finalize|tc_pack|maxabs|user_metrics" -s 5 -c 5 -o gpurun_out/r02_prof_eval_yelp python tools/prof_eval.py yelp-lightgcn > gpurun_out/r02_ncu3.log 2>&1
echo "eval set rc=$?"
ls -la gpurun_out/*.ncu-rep | tail -5
