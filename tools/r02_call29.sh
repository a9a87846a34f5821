mkdir -p gpurun_out
timeout 200 python tools/time_narrow.py yelp 2>/dev/null | grep "us per layer" | tee gpurun_out/r02_narrow2.log
timeout 200 python tools/time_narrow.py amazon 2>/dev/null | grep "us per layer" | tee -a gpurun_out/r02_narrow2.log
