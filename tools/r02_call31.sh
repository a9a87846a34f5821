mkdir -p gpurun_out
python tools/time_narrow.py yelp 16 64 > gpurun_out/r02_plain_narrow.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:prop_kernel -s 20 -c 1 -o gpurun_out/r02_prof_narrow16 python tools/time_narrow.py yelp 16 > gpurun_out/r02_ncu6.log 2>&1
echo "rc=$?"
ncu --set full --clock-control none --import-source on -k regex:prop_kernel -s 20 -c 1 -o gpurun_out/r02_prof_narrow64 python tools/time_narrow.py yelp 64 > gpurun_out/r02_ncu7.log 2>&1
echo "rc=$?"
