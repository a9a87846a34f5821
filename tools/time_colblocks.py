"""Scale-out graph (BASELINE.json configs[4]): time of ONE adjacency layer, user rows and item rows apart, with the item
rows unblocked and cut into column ranges of several widths (graph.column_blocks).
    python tools/time_colblocks.py [block MB ...]"""
import os
import sys
import types

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from igcn_cf_b200 import engine, graph, synth  # noqa: E402


def timed(fn, reps=3):
    fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


def main():
    sizes = [int(a) for a in sys.argv[1:]] or [0, 24, 48, 96]
    shape = os.environ.get('SHAPE', 'scaleout')
    dev = torch.device('cuda:0')
    u, i, e = {'scaleout': (10_000_000, 1_000_000, 500_000_000), 'scaleout-mid': (1_000_000, 200_000, 50_000_000)}[shape]
    dg = synth.gen_device(u, i, e, dev, seed=2021)
    n = dg.n_users + dg.n_items
    x = torch.randn(n, 64, device=dev)
    y = torch.empty_like(x)
    prop = engine.Propagator(n, 64, 1, dev)
    graph.COL_BLOCK_MIN_TABLE = 0
    for mb in sizes:
        graph.COL_BLOCK_BYTES = mb << 20
        adj = graph.NormAdj.from_device(dg)
        if len(adj.blocks) == 1:          # unblocked: still time the two halves apart
            adj2 = graph.NormAdj.from_device(dg, shard=(0, 1))
            adj = adj2 if len(adj2.blocks) == 2 else adj
        for b in adj.blocks:
            one = types.SimpleNamespace(blocks=[b])
            ms = timed(lambda: prop.spmm(one, x, y))
            nb = len(b.col_blocks) if b.col_blocks else 1
            print('%s block %d MB: rows [%d, %d) nnz %d in %d column ranges: %.3f ms' % (shape, mb, b.row0, b.row1, b.csr.nnz, nb, ms),
                  flush=True)
        del adj
        torch.cuda.empty_cache()


if __name__ == '__main__':
    main()
