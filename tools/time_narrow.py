"""Device time of one full propagation layer (igcn_spmm) on column slices of width 64 / 32 / 16 / 8 of the same table
(what a rank of the column-sharded training step runs on 1 / 2 / 4 / 8 GPUs).  python tools/time_narrow.py [shape] [widths ...]"""
import ctypes as C
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402
import torch  # noqa: E402

from igcn_cf_b200 import graph, synth  # noqa: E402
from igcn_cf_b200._lib import call, ptr, stream_ptr  # noqa: E402

shape = sys.argv[1] if len(sys.argv) > 1 else 'yelp'
widths = [int(x) for x in sys.argv[2:]] or [64, 32, 16, 8]
dev = torch.device('cuda:0')
split = synth.gen_named(shape, seed=2021)
ptr_, items = split.csr('train')
users = np.repeat(np.arange(split.n_users, dtype=np.int64), np.diff(ptr_))
adj = graph.NormAdj(split.n_users, split.n_items, np.stack([users, items], axis=1), dev)
n = split.n_users + split.n_items
x = torch.randn(n, 64, device=dev)
none = (C.c_void_p * 1)()
for D in widths:
    xs = x[:, :D].contiguous()
    y = torch.empty_like(xs)
    run = lambda: call('igcn_spmm', adj.csr.struct(D), ptr(xs), ptr(y), D, none, 0, None, 1.0, None, 0, stream_ptr())
    for _ in range(5):
        run()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(50):
        run()
    e1.record()
    torch.cuda.synchronize()
    print('%s D=%d L8=%s: %.1f us per layer' % (shape, D, os.environ.get('IGCN_SPMM_L8', '2'), e0.elapsed_time(e1) / 50 * 1e3), flush=True)
