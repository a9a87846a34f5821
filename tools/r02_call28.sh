mkdir -p gpurun_out
timeout 900 python bench.py --workload scaleout --steps 2 --warmup 1 > gpurun_out/r02_scaleout_1gpu.json 2> gpurun_out/r02_scaleout_1gpu.err; echo "scaleout rc=$?"
tail -c 1500 gpurun_out/r02_scaleout_1gpu.json
