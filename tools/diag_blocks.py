"""Diagnostic (1 GPU): time igcn_spmm on the whole graph and on the row blocks a 2/4/8-rank shard would own,
local stores only -- separates kernel-size effects from NVLink effects."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import bench  # noqa: E402
from igcn_cf_b200 import graph  # noqa: E402
from igcn_cf_b200._lib import call, ptr, stream_ptr  # noqa: E402
import ctypes as C  # noqa: E402


def timeit(fn, n=30):
    """GPU time per call: the calls are captured into a CUDA graph so host launch cost drops out."""
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(n):
            fn()
    g.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    g.replay()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e3


def main():
    shape = sys.argv[1] if len(sys.argv) > 1 else 'yelp'
    dev = torch.device('cuda:0')
    ds = bench.build_dataset(shape, dev)
    n, D = ds.n_users + ds.n_items, 64
    x = torch.randn(n, D, device=dev)
    y = torch.empty(n, D, device=dev)
    none = (C.c_void_p * 1)()
    full = graph.NormAdj(ds.n_users, ds.n_items, ds.train_pairs, dev)
    rp, colh, valh = full.rowptr_full, full.col_full, full.val_full
    for (r0, r1) in ((0, n), (0, 9470), (75173, 80959)):
        lo, hi = int(rp[r0]), int(rp[r1])
        for thr, ch in ((256, 128), (128, 64)):
            csr = graph.CsrDevice(rp[r0:r1 + 1] - lo, colh[lo:hi], valh[lo:hi], n, dev, threshold=thr, chunk=ch)
            off = r0 * D * 4
            fn = lambda: call('igcn_spmm', csr.struct(D), ptr(x), ptr(y) + off, D, none, 0, None, 1.0, None, 0, stream_ptr())
            print('world - rows %d-%d threshold %d chunk %d n_chunks %d: %.1f us' % (r0, r1, thr, ch, csr.n_chunks, timeit(fn)), flush=True)
    for world in (1, 2):
        for rank in range(world):
            adj = graph.NormAdj(ds.n_users, ds.n_items, ds.train_pairs, dev, shard=None if world == 1 else (rank, world))
            line = []
            for blk in adj.blocks:
                off = blk.row0 * D * 4
                fn = lambda: call('igcn_spmm', blk.csr.struct(D), ptr(x), ptr(y) + off, D, none, 0, None, 1.0, None, 0, stream_ptr())
                line.append('rows %d-%d nnz %d chunks %d: %.1f us' % (blk.row0, blk.row1, blk.csr.nnz, blk.csr.n_chunks, timeit(fn)))
            print('world %d rank %d | ' % (world, rank) + ' | '.join(line), flush=True)
            if world > 2 and rank >= 1:
                break


if __name__ == '__main__':
    main()
