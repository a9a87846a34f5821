"""Device time of the graph-replayed training step of one workload (CUDA events, median of 10 x K steps).
    python tools/time_step.py yelp-lightgcn [K]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import bench  # noqa: E402

name = sys.argv[1] if len(sys.argv) > 1 else 'yelp-lightgcn'
K = int(sys.argv[2]) if len(sys.argv) > 2 else 200
shape, kind, l2_reg, dropout = bench.WORKLOADS[name]
dev = torch.device('cuda:0')
ds = bench.build_dataset(shape, dev)
model, trainer = bench.build_model(ds, kind, dropout, l2_reg, dev)
model.train()
for _ in range(20):
    trainer.step.run()
torch.cuda.synchronize()
ts = []
for _ in range(10):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(K):
        trainer.step.run()
    e1.record()
    torch.cuda.synchronize()
    ts.append(e0.elapsed_time(e1) / K)
ts.sort()
print('%s IGCN_SPMM_HOT=%s ms/step median %.4f min %.4f' % (name, os.environ.get('IGCN_SPMM_HOT', '0'), ts[len(ts) // 2], ts[0]))
