# final validation of the round on a 2-GPU box: full GPU suite (incl. the 2-rank parity worker), smoke, CPU suite
mkdir -p gpurun_out
IGCN_PEER_TIMEOUT_S=60 timeout 1800 python -m pytest tests -m gpu -q > gpurun_out/r02_final2_gputests.log 2>&1; echo "rc=$?" >> gpurun_out/r02_final2_gputests.log
tail -5 gpurun_out/r02_final2_gputests.log | cut -c1-300
timeout 300 python __graft_entry__.py smoke > gpurun_out/r02_final2_smoke.log 2>&1; echo "smoke rc=$?"
