mkdir -p gpurun_out
for l in 8 4; do for w in yelp-lightgcn gowalla-igcn; do IGCN_SPMM_LANES=$l timeout 300 python tools/time_step.py $w 2>/dev/null | tail -1 | sed "s/^/lanes=$l (3 CTAs\/SM for 4x4) /" | tee -a gpurun_out/r02_lanes2.log; done; done
