mkdir -p gpurun_out
timeout 200 python tools/time_narrow.py yelp 2>/dev/null | grep "us per layer" | tee gpurun_out/r02_narrow4.log
timeout 200 python tools/time_narrow.py amazon 16 8 2>/dev/null | grep "us per layer" | tee -a gpurun_out/r02_narrow4.log
