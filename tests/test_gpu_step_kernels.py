"""Kernels the fused training step adds on top of the plain layers: the scatter plan (register/shuffle bitonic
sort), the row-list SpMM (last forward layer on the batch's rows only) and the column-filter SpMM (first
backward layer).  Integer outputs are checked bit-exactly against numpy; the partial layers must reproduce the
full layer's values on the rows they compute."""
import ctypes as C
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
DEV = torch.device('cuda:0')


def _plan(triples, item_off, n_rows, want_bits=True):
    from igcn_cf_b200._lib import call, ptr, stream_ptr
    B = triples.shape[0]
    t = torch.from_numpy(triples).to(DEV)
    order = torch.full((3 * B,), -1, dtype=torch.int32, device=DEV)
    seg_start = torch.full((3 * B + 1,), -1, dtype=torch.int32, device=DEV)
    seg_row = torch.full((3 * B,), -1, dtype=torch.int64, device=DEV)
    n_seg = torch.zeros(1, dtype=torch.int32, device=DEV)
    bits = torch.full(((n_rows + 31) // 32 + 1,), -1, dtype=torch.int32, device=DEV) if want_bits else None
    call('igcn_bpr_plan', ptr(t), B, item_off, n_rows, ptr(order), ptr(seg_start), ptr(seg_row), ptr(n_seg), ptr(bits), stream_ptr())
    torch.cuda.synchronize()
    n = int(n_seg.item())
    return order.cpu().numpy(), seg_start.cpu().numpy()[:n + 1], seg_row.cpu().numpy()[:n], bits


def _plan_numpy(triples, item_off):
    B = triples.shape[0]
    ids = np.concatenate([triples[:, 0], triples[:, 1] + item_off, triples[:, 2] + item_off])
    order = np.lexsort((np.arange(3 * B), ids))                       # by id, then slot
    sorted_ids = ids[order]
    heads = np.flatnonzero(np.r_[True, sorted_ids[1:] != sorted_ids[:-1]])
    return order.astype(np.int32), np.r_[heads, 3 * B].astype(np.int32), sorted_ids[heads]


@pytest.mark.parametrize('B,n_users,n_items', [(1, 5, 7), (300, 40, 50), (341, 1000, 3000), (700, 90, 120),
                                                 (1365, 20000, 30000), (2048, 75173, 42706), (2730, 200000, 300000),
                                                 (2048, 400000, 400000), (4000, 5000, 5000)])
def test_scatter_plan_matches_numpy(B, n_users, n_items):
    """Fast path (1/2/4/8 keys per thread, ids < 2^19, 3B <= 8192) and the 64-bit fallback."""
    rng = np.random.default_rng(B)
    tri = np.stack([rng.integers(n_users, size=B), rng.integers(min(n_items, 37), size=B),       # hot positives: long segments
                    rng.integers(n_items, size=B)], axis=1).astype(np.int64)
    n_rows = n_users + n_items
    order, seg_start, seg_row, bits = _plan(tri, n_users, n_rows)
    o, s, r = _plan_numpy(tri, n_users)
    assert np.array_equal(order, o)
    assert np.array_equal(seg_start, s)
    assert np.array_equal(seg_row, r)
    want = np.zeros(((n_rows + 31) // 32) * 32, dtype=bool)
    want[r] = True
    got = np.unpackbits(bits.cpu().numpy()[:(n_rows + 31) // 32].view(np.uint8), bitorder='little').astype(bool)
    assert np.array_equal(got, want)
    _plan(tri, n_users, n_rows, want_bits=False)                      # the bitmap is optional


def _graph(shape='small'):
    from igcn_cf_b200 import graph, synth
    split = synth.gen_named(shape, seed=11)
    ptr_, items = split.csr('train')
    users = np.repeat(np.arange(split.n_users, dtype=np.int64), np.diff(ptr_))
    return split, graph.NormAdj(split.n_users, split.n_items, np.stack([users, items], axis=1), DEV)


def _spmm(name, adj, x, y, adds, alpha, *extra):
    from igcn_cf_b200._lib import call, ptr, stream_ptr
    arr = (C.c_void_p * max(1, len(adds)))(*[a.data_ptr() for a in adds])
    call(name, adj.csr.struct(64), ptr(x), ptr(y), 64, arr, len(adds), None, alpha, *extra, None, 0, stream_ptr())


def test_row_list_layer_equals_full_layer_on_its_rows():
    split, adj = _graph()
    n = split.n_users + split.n_items
    assert adj.csr.n_chunks > 0 and adj.csr.n_medium > 0            # all three row classes are exercised
    g = torch.Generator(device=DEV).manual_seed(0)
    x = torch.randn(n, 64, device=DEV, generator=g)
    add = torch.randn(n, 64, device=DEV, generator=g)
    full = torch.empty_like(x)
    _spmm('igcn_spmm', adj, x, full, [add, x], 0.25)
    deg = np.diff(adj.rowptr_full)
    rows = np.unique(np.r_[np.argsort(-deg)[:40], np.random.default_rng(1).integers(n, size=1500)]).astype(np.int64)
    row_list = torch.from_numpy(np.r_[rows, np.zeros(100, dtype=np.int64)]).to(DEV)    # padded: only n_list entries count
    n_list = torch.tensor([len(rows)], dtype=torch.int32, device=DEV)
    part = torch.full_like(x, float('nan'))
    _spmm('igcn_spmm_rows', adj, x, part, [add, x], 0.25, row_list.data_ptr(), n_list.data_ptr(), len(rows) + 100, 0)
    torch.cuda.synchronize()
    assert torch.equal(part[rows], full[rows])                        # bit-identical on the listed rows
    mask = torch.ones(n, dtype=torch.bool, device=DEV)
    mask[rows] = False
    assert bool(torch.isnan(part[mask]).all())                        # nothing else is written


def test_column_filter_layer_equals_full_layer_on_sparse_input():
    split, adj = _graph()
    n = split.n_users + split.n_items
    rng = np.random.default_rng(2)
    deg = np.diff(adj.rowptr_full)
    touched = np.unique(np.r_[np.argsort(-deg)[:30], rng.integers(n, size=2000)])
    x = torch.zeros(n, 64, device=DEV)
    x[touched] = torch.randn(len(touched), 64, device=DEV)
    flags = np.zeros(((n + 31) // 32) * 32, dtype=np.uint8)
    flags[touched] = 1
    bits = torch.from_numpy(np.packbits(flags, bitorder='little').view(np.int32).copy()).to(DEV)
    full, part = torch.empty_like(x), torch.empty_like(x)
    _spmm('igcn_spmm', adj, x, full, [x], 1.0)
    _spmm('igcn_spmm_cols', adj, x, part, [x], 1.0, bits.data_ptr())
    torch.cuda.synchronize()
    # skipped terms are exact zeros, so the sums agree exactly (up to the sign of zero)
    assert torch.equal(part + 0.0, full + 0.0)


def _spmm_w(name, adj, x, y, adds, alpha, *extra):
    """_spmm for any width: D = x.shape[1]."""
    from igcn_cf_b200._lib import call, ptr, stream_ptr
    D = int(x.shape[1])
    arr = (C.c_void_p * max(1, len(adds)))(*[a.data_ptr() for a in adds])
    call(name, adj.csr.struct(D), ptr(x), ptr(y), D, arr, len(adds), None, alpha, *extra, None, 0, stream_ptr())


@pytest.mark.parametrize('width', [32, 16, 8])
def test_narrow_tables_reproduce_the_column_slices_of_the_full_width_layer(width):
    """The column-sharded training step runs every propagation kernel on D / ranks columns (32, 16, 8 for 2, 4, 8
    GPUs; 4 or 2 lanes per row below 32).  Each kernel variant on a contiguous column slice must give the bits of the
    same slice of the D = 64 result: full layer with add operands, row-list layer, column-filter layer, INMO layer
    forward (hash dropout) and its transpose, masked column sums."""
    from igcn_cf_b200 import engine, graph
    from igcn_cf_b200._lib import call, ptr, stream_ptr
    split, adj = _graph()
    n = split.n_users + split.n_items
    assert adj.csr.n_chunks > 0 and adj.csr.n_medium > 0
    g = torch.Generator(device=DEV).manual_seed(width)
    x = torch.randn(n, 64, device=DEV, generator=g)
    add = torch.randn(n, 64, device=DEV, generator=g)
    rng = np.random.default_rng(width)
    deg = np.diff(adj.rowptr_full)
    rows = np.unique(np.r_[np.argsort(-deg)[:40], rng.integers(n, size=1500)]).astype(np.int64)
    row_list = torch.from_numpy(rows).to(DEV)
    n_list = torch.tensor([len(rows)], dtype=torch.int32, device=DEV)
    touched = np.unique(np.r_[np.argsort(-deg)[:30], rng.integers(n, size=2000)])
    flags = np.zeros(((n + 31) // 32) * 32, dtype=np.uint8)
    flags[touched] = 1
    bits = torch.from_numpy(np.packbits(flags, bitorder='little').view(np.int32).copy()).to(DEV)
    xs = torch.zeros(n, 64, device=DEV)
    xs[touched] = x[touched]

    def run_all(xx, aa, xxs):
        out = {}
        y = torch.empty_like(xx)
        _spmm_w('igcn_spmm', adj, xx, y, [aa, xx], 0.25)
        out['full'] = y
        y = torch.zeros_like(xx)
        _spmm_w('igcn_spmm_rows', adj, xx, y, [aa, xx], 0.25, row_list.data_ptr(), n_list.data_ptr(), len(rows), 0)
        out['rows'] = y
        y = torch.empty_like(xx)
        _spmm_w('igcn_spmm_cols', adj, xxs, y, [xxs], 1.0, bits.data_ptr())
        out['cols'] = y
        return out

    want = run_all(x, add, xs)
    for c0 in range(0, 64, width):
        sl = slice(c0, c0 + width)
        got = run_all(x[:, sl].contiguous(), add[:, sl].contiguous(), xs[:, sl].contiguous())
        torch.cuda.synchronize()
        for k in want:
            assert torch.equal(got[k], want[k][:, sl]), (k, width, c0)

    # INMO layer (identity templates, hash dropout) and its transpose + the masked column sums
    pairs = np.stack([np.repeat(np.arange(split.n_users, dtype=np.int64), np.diff(split.csr('train')[0])), split.csr('train')[1]], axis=1)
    feat = graph.TemplateFeat(split.n_users, split.n_items, pairs, np.arange(split.n_users), np.arange(split.n_items),
                              split.n_users, split.n_items, DEV)
    feat.set_alpha(0.9)
    emb = torch.randn(n + 2, 64, device=DEV, generator=g)
    gsc = torch.randn(n, 64, device=DEV, generator=g)
    drop = {'mode': 1, 'p': 0.3, 'seed': 12345}

    def inmo(e, gs):
        D = int(e.shape[1])
        x0 = torch.empty((n, D), device=DEV)
        engine.inmo_forward(feat, e, x0, drop, D)
        d_emb = torch.zeros_like(e)
        engine.inmo_backward(feat, gs, d_emb, drop, D, engine.colsum_scratch(n, D, DEV))
        return x0, d_emb

    x0_w, de_w = inmo(emb, gsc)
    for c0 in range(0, 64, width):
        sl = slice(c0, c0 + width)
        x0, de = inmo(emb[:, sl].contiguous(), gsc[:, sl].contiguous())
        torch.cuda.synchronize()
        assert torch.equal(x0, x0_w[:, sl]) and torch.equal(de, de_w[:, sl]), (width, c0)
