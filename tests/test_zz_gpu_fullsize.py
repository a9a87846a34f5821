"""Parity at BASELINE.json's full size (configs[1]: Yelp-shaped, 75,173 users x 42,706 items, 1.35 M train
interactions), where the Python oracle would take minutes: the CUDA path is checked through properties that do
not depend on the size -- a full propagation layer against the same product in fp64 (torch.sparse on the device),
linearity and symmetry of the layer, the partial layers against the full one (bit-exact), the fused scoring /
mask / top-k against its own invariants (sorted, unmasked, distinct, scores that recompute, nothing better left
out) with the tensor-core and the exact kernels agreeing bit for bit, the sampler's membership rules, and run-to-run
determinism of the whole training step.  Tolerance for fp32 results: 1e-5 relative (BASELINE.json north_star).
(The file name sorts last on purpose: the small-size parity tests run first.)"""
import ctypes as C

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
DEV = torch.device('cuda:0')
TOL = 1e-5
SHAPE = 'yelp'


@pytest.fixture(scope='module')
def split():
    from igcn_cf_b200 import synth
    return synth.gen_named(SHAPE, seed=2021)


@pytest.fixture(scope='module')
def adj(split):
    from igcn_cf_b200 import graph
    ptr_, items = split.csr('train')
    users = np.repeat(np.arange(split.n_users, dtype=np.int64), np.diff(ptr_))
    return graph.NormAdj(split.n_users, split.n_items, np.stack([users, items], axis=1), DEV)


def _spmm(name, adj, x, y, adds, alpha, *extra):
    from igcn_cf_b200._lib import call, ptr, stream_ptr
    arr = (C.c_void_p * max(1, len(adds)))(*[a.data_ptr() for a in adds])
    call(name, adj.csr.struct(64), ptr(x), ptr(y), 64, arr, len(adds), None, alpha, *extra, None, 0, stream_ptr())


def _layer(adj, x):
    y = torch.empty_like(x)
    _spmm('igcn_spmm', adj, x, y, [], 1.0)
    return y


def _rel(a, b):
    return float((a.double() - b.double()).abs().max() / b.double().abs().max())


def test_full_layer_matches_fp64_product_and_is_linear_and_symmetric(split, adj):
    n = split.n_users + split.n_items
    assert adj.shape[0] == n and adj.csr.n_chunks > 0 and adj.csr.n_medium > 0     # all three row classes present
    g = torch.Generator(device=DEV).manual_seed(3)
    x = torch.rand(n, 64, device=DEV, generator=g)
    z = torch.rand(n, 64, device=DEV, generator=g)
    ax, az = _layer(adj, x), _layer(adj, z)
    a64 = adj.to_sparse_coo().to(torch.float64)
    ref = torch.sparse.mm(a64, x.double())
    assert _rel(ax, ref) < TOL
    # linearity: A (2x + 3z) = 2 Ax + 3 Az
    assert _rel(_layer(adj, 2 * x + 3 * z), 2 * ax.double() + 3 * az.double()) < TOL
    # symmetry of D^-1/2 A D^-1/2: <Ax, z> = <x, Az>  (non-negative vectors: no cancellation in the sums)
    lhs, rhs = (ax.double() * z.double()).sum(), (x.double() * az.double()).sum()
    assert abs(float(lhs - rhs)) / float(rhs) < TOL
    # the layer mean fused into the last layer: alpha * (A x + x + z)
    y = torch.empty_like(x)
    _spmm('igcn_spmm', adj, x, y, [x, z], 1.0 / 3)
    assert _rel(y, (ref + x.double() + z.double()) / 3) < TOL
    torch.cuda.synchronize()


def test_partial_layers_equal_the_full_layer_bit_for_bit(split, adj):
    """igcn_spmm_rows on a training step's row list and igcn_spmm_cols on its touched columns (6,144 of 117,879)."""
    n = split.n_users + split.n_items
    rng = np.random.default_rng(5)
    g = torch.Generator(device=DEV).manual_seed(4)
    x = torch.randn(n, 64, device=DEV, generator=g)
    add = torch.randn(n, 64, device=DEV, generator=g)
    full = torch.empty_like(x)
    _spmm('igcn_spmm', adj, x, full, [add, x], 0.25)
    deg = np.diff(adj.rowptr_full)
    rows = np.unique(np.r_[np.argsort(-deg)[:64], rng.integers(n, size=6000)]).astype(np.int64)
    row_list = torch.from_numpy(np.r_[rows, np.zeros(6144 - len(rows), dtype=np.int64)]).to(DEV)
    n_list = torch.tensor([len(rows)], dtype=torch.int32, device=DEV)
    part = torch.full_like(x, float('nan'))
    _spmm('igcn_spmm_rows', adj, x, part, [add, x], 0.25, row_list.data_ptr(), n_list.data_ptr(), 6144, 0)
    torch.cuda.synchronize()
    assert torch.equal(part[rows], full[rows])
    untouched = torch.ones(n, dtype=torch.bool, device=DEV)
    untouched[rows] = False
    assert bool(torch.isnan(part[untouched]).all())
    # column filter: X is non-zero on the touched rows only
    xs = torch.zeros(n, 64, device=DEV)
    xs[rows] = x[rows]
    flags = np.zeros(((n + 31) // 32) * 32, dtype=np.uint8)
    flags[rows] = 1
    bits = torch.from_numpy(np.packbits(flags, bitorder='little').view(np.int32).copy()).to(DEV)
    a, b = torch.empty_like(x), torch.empty_like(x)
    _spmm('igcn_spmm', adj, xs, a, [xs], 1.0)
    _spmm('igcn_spmm_cols', adj, xs, b, [xs], 1.0, bits.data_ptr())
    torch.cuda.synchronize()
    assert torch.equal(a + 0.0, b + 0.0)


def test_scoring_topk_invariants_and_both_kernels_agree(split):
    from igcn_cf_b200 import engine
    n_users, n_items, k = split.n_users, split.n_items, 20
    g = torch.Generator(device=DEV).manual_seed(6)
    rep = torch.randn(n_users + n_items, 64, device=DEV, generator=g) * 0.1
    ptr_, items = split.csr('train')
    mask = engine.ListCSR.from_arrays(ptr_, items, DEV)
    u = torch.arange(n_users, device=DEV)
    ex_i, ex_s = engine.score_topk(rep, u, n_users, n_items, k, mask, impl='exact')
    tc_i, tc_s = engine.score_topk(rep, u, n_users, n_items, k, mask, impl='tc', users_host='identity')
    torch.cuda.synchronize()
    assert torch.equal(ex_i, tc_i) and torch.equal(ex_s, tc_s)                   # tcgen05 path == CUDA-core path
    rec, sc = tc_i.long(), tc_s
    assert int(rec.min()) >= 0 and int(rec.max()) < n_items                      # every user has >= k unseen items
    assert bool((sc[:, :-1] >= sc[:, 1:]).all())                                 # sorted by score ...
    tie = sc[:, :-1] == sc[:, 1:]
    assert bool((rec[:, :-1][tie] < rec[:, 1:][tie]).all())                      # ... ties by item id
    srt = torch.sort(rec, dim=1)[0]
    assert bool((srt[:, :-1] != srt[:, 1:]).all())                               # distinct items
    seen = np.repeat(np.arange(n_users, dtype=np.int64), np.diff(ptr_)) * n_items + items
    got = (u[:, None] * n_items + rec).cpu().numpy().ravel()
    assert not np.isin(got, seen).any()                                          # no train item is recommended
    # the listed scores are the dot products
    again = (rep[:n_users].double()[:, None, :] * rep[n_users + rec].double()).sum(-1)
    assert float((sc.double() - again).abs().max()) < TOL * float(again.abs().max())
    # nothing better was left out: dense fp32 scores of a sample of users, seen and listed items removed
    pick = torch.from_numpy(np.random.default_rng(7).choice(n_users, size=2048, replace=False)).to(DEV)
    dense = rep[pick] @ rep[n_users:].T
    pc = pick.cpu().numpy()
    rows = np.repeat(np.arange(len(pc)), np.diff(ptr_)[pc])
    cols = np.concatenate([items[ptr_[x]:ptr_[x + 1]] for x in pc])
    dense[torch.from_numpy(rows).to(DEV), torch.from_numpy(cols).to(DEV)] = -float('inf')
    dense.scatter_(1, rec[pick], -float('inf'))
    slack = TOL * float(sc.abs().max())
    assert bool((dense.max(dim=1)[0] <= sc[pick, -1] + slack).all())


def test_sampler_membership_at_full_size(split, adj):
    from igcn_cf_b200._lib import call, ptr, stream_ptr
    n_users, n_items, B = split.n_users, split.n_items, 200000
    rowptr, col = adj.sampler_csr()
    out = torch.empty((B, 3), dtype=torch.int64, device=DEV)
    call('igcn_sample_triples', ptr(rowptr), ptr(col), n_users, n_users, n_items, B, 11, 0, None, ptr(out), stream_ptr())
    torch.cuda.synchronize()
    t = out.cpu().numpy()
    assert t[:, 0].min() >= 0 and t[:, 0].max() < n_users and t[:, 1:].min() >= 0 and t[:, 1:].max() < n_items
    ptr_, items = split.csr('train')
    seen = np.repeat(np.arange(n_users, dtype=np.int64), np.diff(ptr_)) * n_items + items
    assert np.isin(t[:, 0] * n_items + t[:, 1], seen).all()                      # positives are train items of the user
    assert not np.isin(t[:, 0] * n_items + t[:, 2], seen).any()                  # negatives are not


def test_training_steps_are_deterministic_at_full_size():
    import bench
    ds = bench.build_dataset(SHAPE, DEV)
    outs = []
    for _ in range(2):
        model, trainer = bench.build_model(ds, 'LightGCN', None, 1e-4, DEV, use_graph=True)
        model.train()
        start = model.embedding.weight.detach().clone()
        for _ in range(12):
            trainer.step.run()
        torch.cuda.synchronize()
        outs.append((model.embedding.weight.detach().clone(), trainer.step.meter_avg()))
        assert not torch.equal(start, outs[-1][0]) and bool(torch.isfinite(outs[-1][0]).all())
    assert torch.equal(outs[0][0], outs[1][0]) and outs[0][1] == outs[1][1]
    assert 0.0 < outs[0][1] < 1.0                    # mean BPR loss of the first steps: close to ln 2, never above 1
