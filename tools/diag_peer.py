"""Diagnostic (torchrun, >= 2 GPUs): cost of the fused peer-store all-gather per SpMM layer.
Times one layer of the row-sharded SpMM (a) storing locally only, (b) storing to every rank, (c) the
barrier alone, (d) a plain peer copy of the same row block for comparison."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

import bench  # noqa: E402
from igcn_cf_b200 import dist as idist  # noqa: E402
from igcn_cf_b200._lib import call, ptr, stream_ptr  # noqa: E402


def timeit(fn, peers, n=20):
    for _ in range(3):
        fn()
    peers.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e3


def main():
    workload = sys.argv[1] if len(sys.argv) > 1 else 'yelp-lightgcn'
    local = int(os.environ['LOCAL_RANK'])
    torch.cuda.set_device(local)
    dev = torch.device('cuda', local)
    dist.init_process_group('nccl', device_id=dev)
    peers = idist.init_peers()
    shape, kind, l2_reg, dropout = bench.WORKLOADS[workload]
    ds = bench.build_dataset(shape, dev)
    model, trainer = bench.build_model(ds, kind, dropout, l2_reg, dev, use_graph=False)
    prop = model._propagator()
    adj = model.norm_adj
    sh = prop.shard
    D = prop.dim
    x = prop.layers[0]
    y = prop.layers[1]
    x.normal_()
    off = sh.row0 * D * 4
    rows = sh.row1 - sh.row0

    def local_only():
        call('igcn_spmm', adj.csr.struct(D), ptr(x), ptr(y) + off, D, prop._adds((), off), 0, None, 1.0, None, 0, stream_ptr())

    def with_peers():
        arr, n = sh.peers(y, off)
        call('igcn_spmm', adj.csr.struct(D), ptr(x), ptr(y) + off, D, prop._adds((), off), 0, None, 1.0, arr, n, stream_ptr())

    def with_peers_barrier():
        with_peers()
        peers.barrier()

    def barrier_only():
        peers.barrier()

    buf = sh._bufs[y.data_ptr()]
    other = (peers.rank + 1) % peers.world

    def memcpy_block():
        import ctypes
        torch.cuda.cudart().cudaMemcpyAsync(buf.ptrs[other] + off, ptr(y) + off, rows * D * 4, 3, stream_ptr())

    res = {'rows': rows, 'nnz': adj.csr.nnz, 'block_MB': rows * D * 4 / 1e6,
           'local_only_us': timeit(local_only, peers), 'peer_stores_us': timeit(with_peers, peers),
           'peer_stores_barrier_us': timeit(with_peers_barrier, peers), 'barrier_us': timeit(barrier_only, peers)}
    try:
        res['memcpy_block_us'] = timeit(memcpy_block, peers)
    except Exception as e:
        res['memcpy_block_us'] = repr(e)
    print('rank', peers.rank, res, flush=True)
    peers.check()
    dist.barrier()
    idist.shutdown()
    dist.destroy_process_group()


if __name__ == '__main__':
    main()
