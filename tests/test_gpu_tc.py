"""tcgen05 scoring path: (1) the tensor core really produces an UPPER BOUND of the exact score
(validates operand packing, UMMA descriptors and the error-bound K block), (2) its final top-k is
bit-identical to the exact CUDA-core kernel (items and scores), including masks, banned ranges and
bitmaps, odd sizes and the fallback path for users whose bound cannot be verified."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
DEV = torch.device('cuda:0')


def _case(n_users, n_items, D, seed, mask_deg=6, scale=0.1):
    g = torch.Generator().manual_seed(seed)
    rep = (torch.randn(n_users + n_items, D, generator=g) * scale).to(DEV)
    rng = np.random.default_rng(seed)
    lists = [sorted(rng.choice(n_items, size=int(rng.integers(0, mask_deg * 2 + 1)), replace=False).tolist())
             for _ in range(n_users)]
    return rep, lists


def _both(rep, n_users, k, lists=None, lo=0, hi=None, banned=None, users=None):
    from igcn_cf_b200 import engine
    from igcn_cf_b200.graph import _pack_bits
    n_items = rep.shape[0] - n_users
    mask = None if lists is None else engine.lists_to_csr(lists, DEV)
    bits = None
    if banned is not None:
        flags = np.zeros(n_items, dtype=bool)
        flags[list(banned)] = True
        bits = _pack_bits(flags, DEV)
    u = torch.arange(n_users, device=DEV) if users is None else torch.tensor(users, dtype=torch.int64, device=DEV)
    ex = engine.score_topk(rep, u, n_users, n_items, k, mask, lo, hi, bits, impl='exact')
    tc = engine.score_topk(rep, u, n_users, n_items, k, mask, lo, hi, bits, impl='tc')
    torch.cuda.synchronize()
    return ex, tc, int(engine._tc_scorer.last_fallback.item())


def test_tensor_core_scores_are_upper_bounds():
    from igcn_cf_b200 import engine
    n_users, n_items, D = 200, 700, 64
    rep, _ = _case(n_users, n_items, D, seed=5)
    u = torch.arange(n_users, device=DEV)
    scorer = engine.TcScorer()
    _, _, dump, ws = scorer.topk(rep, u, n_users, n_items, 20, dump=True, n_splits=2)
    torch.cuda.synchronize()
    maxabs = ws['maxabs'].view(torch.float32).item()
    assert maxabs == rep.abs().max().item()
    scale = 2.0 ** (9 - int(np.floor(np.log2(maxabs))))
    items = rep[n_users:].double()
    mean = torch.from_numpy(ws['center'].cpu().numpy().astype(np.float64) * np.float64(np.float32(1.0) / np.float32(n_items))).to(DEV)
    assert torch.allclose(mean, items.mean(dim=0), atol=1e-7)
    centered = items - mean                                    # the tensor core scores u . (i - mean item)
    exact = (rep[:n_users].double() @ centered.t()).cpu().numpy()
    s_hat = dump[:n_users, :n_items].double().cpu().numpy() / scale ** 2
    nu = rep[:n_users].double().norm(dim=1).cpu().numpy()[:, None]
    ni = centered.norm(dim=1).cpu().numpy()[None, :]
    slack = s_hat - exact
    assert slack.min() >= 0.0, slack.min()                      # never below the exact score
    assert (slack <= 2.2e-3 * nu * ni + 1e-6).all()             # and not wastefully loose (c = 1e-3)
    assert (slack >= 0.2e-3 * nu * ni).mean() > 0.99            # the bound block is really in the MMA


@pytest.mark.parametrize('n_users,n_items,D,k', [(300, 1000, 64, 20), (129, 257, 64, 5), (64, 3000, 32, 24),
                                                 (1000, 5000, 64, 20), (5, 40, 64, 20)])
def test_tc_equals_exact(n_users, n_items, D, k):
    rep, lists = _case(n_users, n_items, D, seed=n_items + 1)
    ex, tc, fb = _both(rep, n_users, k, lists)
    assert torch.equal(ex[0], tc[0]) and torch.equal(ex[1], tc[1])
    assert fb <= max(2, n_users // 50)


def test_tc_unmasked_and_user_subset():
    rep, lists = _case(400, 900, 64, seed=9)
    ex, tc, _ = _both(rep, 400, 20)
    assert torch.equal(ex[0], tc[0]) and torch.equal(ex[1], tc[1])
    users = [3, 399, 0, 77, 77, 200]
    ex, tc, _ = _both(rep, 400, 20, lists, users=users)
    assert torch.equal(ex[0], tc[0]) and torch.equal(ex[1], tc[1])


@pytest.mark.parametrize('lo,hi', [(0, 700), (300, 1000), (250, 260), (256, 512)])
def test_tc_item_ranges(lo, hi):
    rep, lists = _case(150, 1000, 64, seed=2)
    ex, tc, _ = _both(rep, 150, 20, lists, lo=lo, hi=hi)
    assert torch.equal(ex[0], tc[0]) and torch.equal(ex[1], tc[1])
    ok = tc[0][tc[0] >= 0]
    assert int(ok.min()) >= lo and int(ok.max()) < hi


def test_tc_banned_bitmap_and_exhaustion():
    rep, lists = _case(70, 300, 64, seed=3, mask_deg=148)
    ex, tc, _ = _both(rep, 70, 24, lists, banned=range(0, 300, 2))
    assert torch.equal(ex[0], tc[0]) and torch.equal(ex[1], tc[1])
    assert bool((tc[0] == -1).any())


def test_tc_fallback_on_ties():
    """Duplicated item rows -> more than 32 exactly tied scores: the bound cannot separate them, the
    user must go through the exact kernel, and the answer is still identical."""
    rep, lists = _case(130, 600, 64, seed=4)
    rep[130 + 100:130 + 200] = rep[130 + 7]
    ex, tc, fb = _both(rep, 130, 20, lists)
    assert torch.equal(ex[0], tc[0]) and torch.equal(ex[1], tc[1])
    assert fb > 0


def test_tc_large_dynamic_range():
    rep, lists = _case(256, 2048, 64, seed=6, scale=30.0)
    rep[256 + 5] *= 40.0          # one huge-norm item
    rep[10] *= 1e-4               # one tiny-norm user
    ex, tc, _ = _both(rep, 256, 20, lists)
    assert torch.equal(ex[0], tc[0]) and torch.equal(ex[1], tc[1])


@pytest.mark.parametrize('n_tiles', [150, 300])
def test_tc_split_tail_plan(n_tiles):
    """More user tiles than SMs: whole waves are left unsplit (list slot 0 only) and the tail is split to fill
    the last wave; same answer as the exact kernel."""
    from igcn_cf_b200 import engine
    n_users, n_items, D, k = (n_tiles - 1) * 128 + 3, 700, 32, 20
    rep, lists = _case(n_users, n_items, D, seed=11, mask_deg=3)
    n_head, n_splits = engine.TcScorer.plan_ctas(n_tiles, (n_items + 255) // 256)
    assert 0 < n_head < n_tiles and n_splits > 1 and n_head % 148 == 0
    mask = engine.lists_to_csr(lists, DEV)
    u = torch.arange(n_users, device=DEV)
    ex = engine.score_topk(rep, u, n_users, n_items, k, mask, impl='exact')
    tc = engine.score_topk(rep, u, n_users, n_items, k, mask, impl='tc')
    torch.cuda.synchronize()
    assert torch.equal(ex[0], tc[0]) and torch.equal(ex[1], tc[1])


def _random_order(n_items, seed):
    from igcn_cf_b200 import engine
    return engine.ItemOrder(np.random.default_rng(seed).permutation(n_items), DEV)


@pytest.mark.parametrize('n_users,n_items,D,k', [(300, 1000, 64, 20), (129, 257, 64, 5), (1000, 5000, 64, 20), (700, 20000, 32, 20)])
def test_tc_scan_order_does_not_change_the_result(n_users, n_items, D, k):
    """The items are scanned in a caller-supplied order (engine.ItemOrder; the trainers pass train popularity):
    mask buckets, candidate lists and thresholds live in position space, the final lists must still be the exact
    kernel's, bit for bit."""
    from igcn_cf_b200 import engine
    rep, lists = _case(n_users, n_items, D, seed=n_items + 3)
    mask = engine.lists_to_csr(lists, DEV)
    u = torch.arange(n_users, device=DEV)
    ex = engine.score_topk(rep, u, n_users, n_items, k, mask, impl='exact')
    for order in (_random_order(n_items, 1), engine.ItemOrder.by_score(rep[n_users:].norm(dim=1).cpu().numpy(), DEV)):
        tc = engine.score_topk(rep, u, n_users, n_items, k, mask, impl='tc', order=order)
        torch.cuda.synchronize()
        assert torch.equal(ex[0], tc[0]) and torch.equal(ex[1], tc[1])


@pytest.mark.parametrize('lo,hi,banned', [(300, 1000, None), (0, 700, range(0, 1000, 3)), (250, 260, None), (0, 1000, range(5, 900))])
def test_tc_scan_order_with_ranges_and_banned_items(lo, hi, banned):
    from igcn_cf_b200 import engine
    from igcn_cf_b200.graph import _pack_bits
    n_users, n_items = 150, 1000
    rep, lists = _case(n_users, n_items, 64, seed=12)
    mask = engine.lists_to_csr(lists, DEV)
    bits = None
    if banned is not None:
        flags = np.zeros(n_items, dtype=bool)
        flags[list(banned)] = True
        bits = _pack_bits(flags, DEV)
    u = torch.arange(n_users, device=DEV)
    ex = engine.score_topk(rep, u, n_users, n_items, 20, mask, lo, hi, bits, impl='exact')
    tc = engine.score_topk(rep, u, n_users, n_items, 20, mask, lo, hi, bits, impl='tc', order=_random_order(n_items, 2))
    torch.cuda.synchronize()
    assert torch.equal(ex[0], tc[0]) and torch.equal(ex[1], tc[1])


def test_tc_filter_statistics_and_popular_first_order():
    """Structured scores (a few items every user likes): scanning them first must keep far more chunks on the
    compare-free path than the natural order, with identical results; the counters are consistent."""
    from igcn_cf_b200 import engine
    n_users, n_items, D, k = 1024, 8192, 64, 20
    g = torch.Generator().manual_seed(3)
    rep = (torch.randn(n_users + n_items, D, generator=g) * 0.1)
    direction = torch.randn(D, generator=g)
    direction /= direction.norm()
    pop = torch.rand(n_items, generator=g) ** 8                       # a handful of very popular items, spread over the ids
    rep[:n_users] += 0.3 * direction
    rep[n_users:] += pop[:, None] * 1.5 * direction
    rep = rep.to(DEV)
    u = torch.arange(n_users, device=DEV)
    scorer = engine.TcScorer()
    out = {}
    for name, order in (('natural', None), ('popular', engine.ItemOrder.by_score(pop.numpy(), DEV))):
        stats = torch.zeros(5, dtype=torch.int64, device=DEV)
        items, scores = scorer.topk(rep, u, n_users, n_items, k, order=order, stats=stats, n_splits=1)
        torch.cuda.synchronize()
        out[name] = (items.clone(), scores.clone(), stats.tolist())
    ex = engine.score_topk(rep, u, n_users, n_items, k, impl='exact')
    for name in out:
        assert torch.equal(out[name][0], ex[0]) and torch.equal(out[name][1], ex[1])
        chunks, slow, groups, hits, comp = out[name][2]
        assert chunks == (n_users // 32) * (n_items // 32) and 0 < slow <= chunks and slow <= groups <= 4 * slow
        assert hits >= n_users * k and comp <= slow
    assert out['popular'][2][1] < 0.5 * out['natural'][2][1]          # far fewer chunks leave the compare-free path
    # item splits are strided over the scan order (every list sees the popular head first): still the same answer
    stats = torch.zeros(5, dtype=torch.int64, device=DEV)
    items, scores = scorer.topk(rep, u, n_users, n_items, k, order=engine.ItemOrder.by_score(pop.numpy(), DEV), stats=stats, n_splits=4)
    torch.cuda.synchronize()
    assert torch.equal(items, ex[0]) and torch.equal(scores, ex[1])
    assert stats.tolist()[1] < out['natural'][2][1]
