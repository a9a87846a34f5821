"""CPU-side checks: the C-ABI library loads and exports every declared symbol, and the host-built
graph structures equal the reference's tensors (golden vectors).  No kernel is launched here."""
import ctypes
import os
import re

import numpy as np
import pytest
import torch

from conftest import ROOT, load_golden
from igcn_cf_b200 import _lib, graph
from igcn_cf_b200.dataset import AuxiliaryDataset, ListDataset, get_dataset


def _declared_functions():
    text = open(os.path.join(ROOT, 'include', 'igcn_b200.h')).read()
    text = re.sub(r'/\*.*?\*/', '', text, flags=re.S)
    return sorted(set(re.findall(r'\b(igcn_[a-z0-9_]+)\s*\(', text)))


def test_library_exports_every_declared_symbol():
    if not os.path.exists(_lib.LIB_PATH):
        from igcn_cf_b200 import build
        build.build()
    lib = ctypes.CDLL(_lib.LIB_PATH)
    declared = _declared_functions()
    assert len(declared) >= 16
    for name in declared:
        assert hasattr(lib, name), name
    assert sorted(_lib.EXPORTS) == declared          # the ctypes binding covers the whole header
    lib.igcn_abi_version.restype = ctypes.c_int
    assert lib.igcn_abi_version() == _lib.ABI_VERSION


def test_struct_layout_matches_header():
    # sizes implied by include/igcn_b200.h on LP64
    assert ctypes.sizeof(_lib.CsrStruct) == 3 * 8 + 3 * 8 + 2 * 4 + 8 * 8 + 2 * 4
    assert ctypes.sizeof(_lib.DropoutStruct) == 4 + 4 + 8 + 4 * 8


def test_missing_cuda_fails_loudly(tiny):
    from igcn_cf_b200.model import get_model
    ds = _dataset(tiny)
    with pytest.raises(RuntimeError, match='no CPU fallback'):
        get_model({'name': 'LightGCN', 'embedding_size': 64, 'n_layers': 3, 'device': torch.device('cpu')}, ds)
    with pytest.raises(RuntimeError, match='CUDA'):
        _lib.require_cuda(torch.zeros(4), torch.float32, 'x')


def _dataset(tiny):
    return get_dataset({'name': 'ListDataset', 'train': tiny['train'], 'val': tiny['val'], 'test': tiny['test'],
                        'n_items': tiny['n_items'], 'device': 'cpu'})


def test_dataset_mirror(tiny):
    ds = _dataset(tiny)
    assert (ds.n_users, ds.n_items) == (tiny['n_users'], tiny['n_items'])
    assert ds.train_array == tiny['pairs'].tolist()
    assert len(ds) == len(tiny['pairs'])
    t = ds[0]
    assert t.shape == (1, 3) and t.dtype == np.int64
    u, p, n = t[0]
    assert p in ds.train_data[u] and n not in ds.train_data[u]
    aux = AuxiliaryDataset(ds, {u: u for u in range(ds.n_users)}, {i: i for i in range(ds.n_items)})
    assert aux.train_data == ds.train_data and len(aux) == len(ds)


def test_norm_adj_matches_reference(tiny):
    g = load_golden('tiny_lightgcn')
    adj = graph.NormAdj(tiny['n_users'], tiny['n_items'], tiny['pairs'], 'cpu')
    assert np.array_equal(adj.indices().numpy(), g['adj_idx'])
    assert np.array_equal(adj.values().numpy(), g['adj_val'])      # bit-exact fp32 values
    assert adj._nnz() == g['adj_val'].shape[0] and tuple(adj.shape) == (700, 700)


def _feat(tiny, user_tmpl, item_tmpl, pairs=None, n_users=None, n_items=None):
    n_users = tiny['n_users'] if n_users is None else n_users
    n_items = tiny['n_items'] if n_items is None else n_items
    pairs = tiny['pairs'] if pairs is None else pairs
    return graph.TemplateFeat(n_users, n_items, pairs, user_tmpl, item_tmpl,
                              int((user_tmpl >= 0).sum()), int((item_tmpl >= 0).sum()), 'cpu')


def test_template_feat_identity_matches_reference(tiny):
    g = load_golden('tiny_igcn')
    f = _feat(tiny, np.arange(tiny['n_users']), np.arange(tiny['n_items']))
    f.set_alpha(1.)
    assert f.tmpl is None                                           # identity map needs no indirection
    assert list(f.shape) == g['feat_shape'].tolist()
    assert np.array_equal(f.indices().numpy(), g['feat_idx'])
    assert np.array_equal(f.row_sum.numpy(), g['row_sum'])
    np.testing.assert_allclose(f.values().numpy(), g['feat_val'], rtol=1e-6)
    f.set_alpha(float(g['alpha1']))
    np.testing.assert_allclose(f.values().numpy(), g['feat_val1'], rtol=1e-6)


def test_template_feat_ratio_matches_reference(tiny):
    g = load_golden('tiny_igcn_ratio')
    f = _feat(tiny, g['user_map'], g['item_map'])
    f.set_alpha(1.)
    assert f.tmpl is not None
    assert list(f.shape) == g['feat_shape'].tolist()
    assert np.array_equal(f.indices().numpy(), g['feat_idx'])
    assert np.array_equal(f.row_sum.numpy(), g['row_sum'])
    np.testing.assert_allclose(f.values().numpy(), g['feat_val'], rtol=1e-6)


def test_keep_bits_and_tperm(tiny):
    f = _feat(tiny, np.arange(tiny['n_users']), np.arange(tiny['n_items']))
    rng = np.random.default_rng(0)
    keep = rng.random(f._nnz()) < 0.7
    ek, sk = f.keep_bits(keep)
    edge_pos, self_pos, _ = f.coo_order()
    words = ek.numpy().view(np.uint32)
    for e in (0, 1, 31, 32, 33, len(edge_pos) - 1):
        assert bool((words[e >> 5] >> (e & 31)) & 1) == bool(keep[edge_pos[e]])
    words = sk.numpy().view(np.uint32)
    for r in (0, 5, 299, 300, 699):
        assert bool((words[r >> 5] >> (r & 31)) & 1) == bool(keep[self_pos[r]])
    tp = f.tperm().numpy()
    rp, col = f.csr.rowptr_host, f.csr.col_host
    rows = np.repeat(np.arange(len(rp) - 1), np.diff(rp))
    assert np.array_equal(col[tp], rows) and np.array_equal(rows[tp], col)


def test_chunk_plan_covers_long_rows():
    rowptr = np.array([0, 3, 3, 1003, 1010, 1600], dtype=np.int64)
    row, begin, length, first, count = graph.chunk_plan(rowptr, threshold=512, chunk=256)
    assert row.tolist() == [2, 2, 2, 2, 4, 4, 4]
    assert begin.tolist() == [3, 259, 515, 771, 1010, 1266, 1522]
    assert length.tolist() == [256, 256, 256, 232, 256, 256, 78]
    assert first.tolist() == [0, 0, 0, 0, 4, 4, 4] and count.tolist() == [4, 4, 4, 4, 3, 3, 3]
    assert graph.chunk_plan(np.array([0, 5, 9], dtype=np.int64))[0].size == 0


@pytest.mark.parametrize('dups', [False, True])
def test_device_graph_builder_matches_scipy_builder(dups):
    """the edge list -> CSR / values / template row sums built with device ops
    (DeviceGraph.from_pairs + from_device; torch ops, so they also run on the CPU here) are bit-identical to
    the numpy/scipy constructors that restate utils.py:41-49 and model.py:386-421 -- including duplicated
    interactions, isolated nodes and partial template sets."""
    rng = np.random.default_rng(7)
    n_users, n_items = 90, 70
    pairs = np.stack([rng.integers(0, n_users - 5, 900), rng.integers(3, n_items, 900)], axis=1).astype(np.int64)
    if not dups:
        pairs = np.unique(pairs, axis=0)
        pairs = pairs[rng.permutation(len(pairs))]
    dev = torch.device('cpu')
    dg = graph.DeviceGraph.from_pairs(n_users, n_items, pairs, dev)
    assert (dg.mult is not None) == dups
    host = graph.NormAdj(n_users, n_items, pairs, dev)
    devb = graph.NormAdj.from_device(dg)
    assert np.array_equal(host.rowptr_full, devb.rowptr_full)
    assert np.array_equal(host.col_full, devb.col_full) and np.array_equal(host.csr.col_host, devb.csr.col_host)
    assert np.array_equal(host.val_full, devb.val_full)                    # bit-identical, not allclose
    assert torch.equal(host.indices(), devb.indices()) and torch.equal(host.values(), devb.values())
    sr, sc = devb.sampler_csr()
    assert np.array_equal(sr.numpy(), host.rowptr_full[:n_users + 1]) and np.array_equal(sc.numpy(), host.col_full[:sr[-1]])
    # templates: identity, and a partial set (every third user / the first 50 items)
    ut = np.where(np.arange(n_users) % 3 == 0, np.arange(n_users) // 3, -1)
    it = np.where(np.arange(n_items) < 50, np.arange(n_items), -1)
    cases = [(np.arange(n_users), np.arange(n_items), n_users, n_items), (ut, it, int((ut >= 0).sum()), 50)]
    for u_t, i_t, tu, ti in cases:
        fh = graph.TemplateFeat(n_users, n_items, pairs, u_t, i_t, tu, ti, dev)
        fd = graph.TemplateFeat.from_device(dg, adj=devb, user_tmpl=u_t, item_tmpl=i_t, t_users=tu, t_items=ti)
        assert fh.shape == fd.shape and (fh.tmpl is None) == (fd.tmpl is None)
        assert fh.tmpl is None or torch.equal(fh.tmpl, fd.tmpl)
        assert torch.equal(fh.row_sum, fd.row_sum)
        assert (fh.glob_user, fh.glob_item) == (fd.glob_user, fd.glob_item)
        fh.set_alpha(0.7)
        fd.set_alpha(0.7)
        assert torch.equal(fh.indices(), fd.indices()) and torch.equal(fh.values(), fd.values())
        assert torch.equal(fh.tperm(), fd.tperm())
    fi = graph.TemplateFeat.from_device(dg, adj=devb)                      # feature_ratio == 1 shortcut
    assert fi.tmpl is None and torch.equal(fi.row_sum, graph.TemplateFeat(n_users, n_items, pairs, *cases[0], dev).row_sum)


def test_processed_dataset_fast_parser(tmp_path):
    """ProcessedDataset.read_data (dataset.py:154-164): the numpy tokeniser gives the lists, n_items, train
    pairs and CSR arrays of the reference's per-token loop, keeps users without items, and falls back to that
    loop for files with irregular separators."""
    import warnings
    from igcn_cf_b200.dataset import ProcessedDataset, output_data
    rng = np.random.default_rng(3)
    lists = {w: [sorted(rng.choice(60, size=int(rng.integers(0, 6)), replace=False).tolist()) for _ in range(40)]
             for w in ('train', 'val', 'test')}
    lists['train'][7] = []
    lists['test'][39] = []
    for w in lists:
        output_data(str(tmp_path / (w + '.txt')), lists[w])
    with warnings.catch_warnings():
        warnings.simplefilter('error')                      # the tokeniser must not lean on deprecated numpy paths
        ds = ProcessedDataset({'name': 'ProcessedDataset', 'path': str(tmp_path), 'device': torch.device('cpu')})
    assert ds.train_data == lists['train'] and ds.val_data == lists['val'] and ds.test_data == lists['test']
    assert ds.n_users == 40 and ds.n_items == 1 + max(max(x) for w in lists for x in lists[w] if x)
    assert ds.train_array == [[u, i] for u, x in enumerate(lists['train']) for i in x]
    for w in lists:
        ptr, items = ds.csr(w)
        assert ptr.tolist() == np.cumsum([0] + [len(x) for x in lists[w]]).tolist()
        assert items.tolist() == [i for x in lists[w] for i in x]
    assert '_parsed' in ds.__dict__                          # the arrays came from the parser, not from a list walk
    # irregular separators (double space): same result through the reference's loop
    with open(tmp_path / 'val.txt', 'w') as f:
        for u, x in enumerate(lists['val']):
            f.write(' '.join([str(u)] + [str(i) for i in x]) + '\n')
    raw = open(tmp_path / 'train.txt').read().replace('\n', ' \n', 1)
    open(tmp_path / 'train.txt', 'w').write(raw)
    with pytest.raises(ValueError):
        ProcessedDataset({'name': 'ProcessedDataset', 'path': str(tmp_path), 'device': torch.device('cpu')})


def test_scoring_cta_plan(monkeypatch):
    """TcScorer.plan_ctas: whole waves of user tiles stay unsplit, only the tail that would run as a partial
    last wave is split (at most 8 ways, never more CTAs than one wave); fewer tiles than SMs: uniform splits."""
    from igcn_cf_b200.engine import TcScorer
    monkeypatch.delenv('IGCN_TC_SPLITS', raising=False)
    for n in list(range(1, 700)) + [1175, 5000, 78125]:
        n_head, n_splits = TcScorer.plan_ctas(n, 167)
        assert 0 <= n_head <= n and 1 <= n_splits <= 8
        if n < 148:
            assert n_head == 0 and n_splits == TcScorer.pick_splits(n)
            continue
        tail = n - n_head
        assert n_head % 148 == 0 or tail == 0                      # the unsplit part is whole waves (or everything)
        assert tail * n_splits <= 148 or n_splits == 1             # the split tail fits one wave
        if tail and n_splits > 1:
            assert tail * (n_splits + 1) > 148 or n_splits == 8    # and uses as much of it as 8 splits allow
    assert TcScorer.plan_ctas(588) == (588, 1) and TcScorer.plan_ctas(297) == (296, 8) and TcScorer.plan_ctas(234) == (234, 1)
    monkeypatch.setenv('IGCN_TC_SPLITS', '3')
    assert TcScorer.plan_ctas(588) == (0, 3)


def test_device_graph_builder_row_sharded_blocks_match():
    """Row-sharded construction (one user slice + one item slice per rank): the device builder cuts the same
    blocks with the same values as the scipy builder, for every rank of a 3-rank world."""
    rng = np.random.default_rng(9)
    n_users, n_items = 120, 80
    pairs = np.unique(np.stack([rng.integers(0, n_users, 1500), rng.integers(0, n_items, 1500)], axis=1), axis=0).astype(np.int64)
    dev = torch.device('cpu')
    dg = graph.DeviceGraph.from_pairs(n_users, n_items, pairs, dev)
    for rank in range(3):
        host = graph.NormAdj(n_users, n_items, pairs, dev, shard=(rank, 3))
        devb = graph.NormAdj.from_device(dg, shard=(rank, 3))
        assert host.block_key() == devb.block_key() and len(devb.blocks) == 2
        for bh, bd in zip(host.blocks, devb.blocks):
            assert np.array_equal(bh.csr.rowptr_host, bd.csr.rowptr_host)
            assert torch.equal(bh.csr.col, bd.csr.col) and torch.equal(bh.csr.val, bd.csr.val)
        fh = graph.TemplateFeat(n_users, n_items, pairs, np.arange(n_users), np.arange(n_items), n_users, n_items, dev, shard=(rank, 3))
        fd = graph.TemplateFeat.from_device(dg, adj=devb, shard=(rank, 3))
        assert fh.block_key() == fd.block_key() and torch.equal(fh.row_sum, fd.row_sum)
        for bh, bd in zip(fh.blocks, fd.blocks):
            assert torch.equal(bh.csr.col, bd.csr.col) and bd.csr.val is None


def test_row_normalised_adjacency_of_ngcf_matches_reference(tiny):
    """siblings.RowNormAdj = normalize(A + I, 'l1') of NGCF.generate_graph (reference model.py:255-261): indices and
    values bit-equal to the reference's coalesced COO, the stored transpose values are those of A^T, and a restatement
    of NGCF.get_rep (model.py:277-291, eval mode) on that matrix reproduces the reference's representation."""
    from igcn_cf_b200.siblings import RowNormAdj
    g = load_golden('tiny_ngcf_imcgae')
    dg = graph.DeviceGraph.from_pairs(tiny['n_users'], tiny['n_items'], tiny['pairs'], 'cpu')
    adj = RowNormAdj(dg)
    n = tiny['n_users'] + tiny['n_items']
    assert tuple(adj.shape) == (n, n) and adj._nnz() == g['ngcf_adj_idx'].shape[1]
    assert np.array_equal(adj.indices().numpy(), g['ngcf_adj_idx'])
    assert np.array_equal(adj.values().numpy(), g['ngcf_adj_val'])
    r, c = adj.indices()
    dense, dense_t = torch.zeros(n, n), torch.zeros(n, n)
    dense[r, c] = adj.csr_fwd.val
    dense_t[r, c] = adj.csr_bwd.val
    assert torch.equal(dense_t, dense.t()) and torch.equal(dense.sum(1), dense.sum(1))
    assert float((dense.sum(1) - 1).abs().max()) < 1e-6
    fwd, bwd = adj.pair(torch.arange(adj.nnz) % 3 != 0, 0.25)                      # edge dropout keeps the pair consistent
    dense[r, c], dense_t[r, c] = fwd.val, bwd.val
    assert torch.equal(dense_t, dense.t())
    # restated eval-mode forward on the golden parameters
    dense[r, c] = adj.csr_fwd.val
    p = lambda k: torch.from_numpy(g['ngcf_p0_' + k])
    rep = p('embedding.weight')
    hops = [rep]
    for l in range(3):
        m0 = dense @ rep
        m1 = rep * m0
        rep = torch.nn.functional.leaky_relu(m0 @ p('gc_layers.%d.weight' % l).t() + p('gc_layers.%d.bias' % l)
                                             + m1 @ p('bi_layers.%d.weight' % l).t() + p('bi_layers.%d.bias' % l), 0.2)
        hops.append(torch.nn.functional.normalize(rep, p=2, dim=1))
    got = torch.cat(hops, 1).numpy()[::5]
    assert np.abs(got - g['ngcf_rep0_eval_every5']).max() < 1e-5


def test_column_blocks_partition_a_csr():
    """graph.column_blocks: every entry lands in exactly one range, rows keep their column order, empty ranges and
    empty rows survive (pure index work: runs on the CPU)."""
    rng = np.random.default_rng(4)
    n_rows, n_cols, width = 50, 1000, 300
    rows = []
    for r in range(n_rows):
        k = 0 if r % 7 == 0 else int(rng.integers(1, 40))
        rows.append(np.sort(rng.choice(np.arange(100, 900), size=k, replace=False)))     # columns 900.. stay empty
    rowptr = np.zeros(n_rows + 1, dtype=np.int64)
    np.cumsum([len(x) for x in rows], out=rowptr[1:])
    col = torch.from_numpy(np.concatenate(rows).astype(np.int32))
    val = torch.arange(len(col), dtype=torch.float32)
    blocks = graph.column_blocks(rowptr, col, val, 0, n_cols, width, n_cols, 'cpu')
    assert len(blocks) == 4 and blocks[3].nnz == 0
    seen = [[] for _ in range(n_rows)]
    for b, c in enumerate(blocks):
        assert c.n_rows == n_rows and c.n_cols == n_cols
        for r in range(n_rows):
            lo, hi = c.rowptr_host[r], c.rowptr_host[r + 1]
            cc = c.col[lo:hi].numpy()
            assert ((cc >= b * width) & (cc < (b + 1) * width)).all() and (np.diff(cc) > 0).all()
            seen[r] += list(zip(cc.tolist(), c.val[lo:hi].tolist()))
    for r in range(n_rows):
        lo, hi = rowptr[r], rowptr[r + 1]
        assert seen[r] == list(zip(col[lo:hi].tolist(), val[lo:hi].tolist()))
    assert graph.column_block_width(10_000_000) == graph.COL_BLOCK_BYTES // 256 and graph.column_block_width(100_000) == 0
