"""Host-side profile of trainer.eval (D2H, metrics): python tools/prof_eval_host.py [workload]"""
import cProfile
import os
import pstats
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import bench  # noqa: E402

workload = sys.argv[1] if len(sys.argv) > 1 else 'yelp-lightgcn'
shape, kind, l2_reg, dropout = bench.WORKLOADS[workload]
dev = torch.device('cuda:0')
ds = bench.build_dataset(shape, dev)
model, trainer = bench.build_model(ds, kind, dropout, l2_reg, dev, use_graph=False)
for _ in range(3):
    model._bump()
    trainer.eval('val')
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(5):
    model._bump()
    trainer.eval('val')
torch.cuda.synchronize()
print('eval e2e ms', (time.perf_counter() - t0) / 5 * 1e3)
pr = cProfile.Profile()
pr.enable()
for _ in range(5):
    model._bump()
    trainer.eval('val')
torch.cuda.synchronize()
pr.disable()
pstats.Stats(pr).sort_stats('tottime').print_stats(18)
