"""Fused score + mask + top-k kernel against torch.mm / index_put / topk on the CPU oracle side
(reference model.py:122, trainer.py:149-164), including the edge cases: ragged / empty mask rows,
banned ranges and bitmaps, k larger than the number of candidates, odd sizes, D in {32, 64, 128}."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
DEV = torch.device('cuda:0')


def _case(n_users, n_items, D, seed, mask_deg=5):
    g = torch.Generator().manual_seed(seed)
    rep = torch.randn(n_users + n_items, D, generator=g) * 0.1
    rng = np.random.default_rng(seed)
    lists = [sorted(rng.choice(n_items, size=int(rng.integers(0, mask_deg * 2 + 1)), replace=False).tolist())
             for _ in range(n_users)]
    return rep, lists


def _oracle(rep, n_users, users, lists, k, lo=0, hi=None, banned=None):
    from oracle import restate as R
    scores = torch.mm(rep[users, :], rep[n_users:, :].t())
    n_items = scores.shape[1]
    ban = []
    if lo > 0:
        ban += list(range(lo))
    if hi is not None and hi < n_items:
        ban += list(range(hi, n_items))
    if banned is not None:
        ban += list(banned)
    vals, items = R.masked_topk(scores, users, min(k, n_items), lists, np.array(ban, dtype=np.int64) if ban else None)
    return vals, items, scores.numpy()


def _run(rep, n_users, users, lists, k, lo=0, hi=None, banned=None):
    from igcn_cf_b200 import engine
    from igcn_cf_b200.graph import _pack_bits
    n_items = rep.shape[0] - n_users
    mask = None if lists is None else engine.lists_to_csr(lists, DEV)[:2]
    bits = None
    if banned is not None:
        flags = np.zeros(n_items, dtype=bool)
        flags[list(banned)] = True
        bits = _pack_bits(flags, DEV)
    items, vals = engine.score_topk(rep.to(DEV).contiguous(), torch.tensor(users, dtype=torch.int64, device=DEV), n_users,
                                    n_items, k, mask=mask, item_lo=lo, item_hi=hi, banned_bits=bits)
    return items.cpu().numpy(), vals.cpu().numpy()


def _compare(items, vals, o_vals, o_items, scores, tol=1e-5):
    """Identical lists except where the oracle's neighbouring scores are tied within fp32 noise."""
    k = o_items.shape[1]
    scale = np.abs(scores).max()
    for r in range(items.shape[0]):
        finite = np.isfinite(o_vals[r])
        nf = int(finite.sum())
        assert (items[r, nf:k] == -1).all()                        # masked slots are reported as -1
        if np.array_equal(items[r, :nf], o_items[r, :nf]):
            np.testing.assert_allclose(vals[r, :nf], o_vals[r, :nf], atol=tol * scale)
            continue
        for a, b in zip(items[r, :nf], o_items[r, :nf]):           # a swap is only legal inside a tie
            if a != b:
                assert abs(scores[r, a] - scores[r, b]) <= tol * scale, (r, a, b)


@pytest.mark.parametrize('n_users,n_items,D,k', [(130, 1000, 64, 20), (64, 257, 32, 5), (7, 3000, 128, 50),
                                                 (300, 400, 64, 128)])
def test_topk_matches_oracle(n_users, n_items, D, k):
    rep, lists = _case(n_users, n_items, D, seed=n_items)
    users = list(range(n_users))
    items, vals = _run(rep, n_users, users, lists, k)
    o_vals, o_items, scores = _oracle(rep, n_users, users, lists, k)
    _compare(items, vals, o_vals, o_items, scores)


def test_topk_unmasked_subset_of_users():
    rep, _ = _case(200, 900, 64, seed=1)
    users = [3, 199, 0, 77, 77]
    items, vals = _run(rep, 200, users, None, 20)
    o_vals, o_items, scores = _oracle(rep, 200, users, None, 20)
    _compare(items, vals, o_vals, o_items, scores)


@pytest.mark.parametrize('lo,hi', [(0, 700), (300, 1000), (128, 129)])
def test_topk_banned_ranges(lo, hi):
    rep, lists = _case(100, 1000, 64, seed=2)
    users = list(range(100))
    items, vals = _run(rep, 100, users, lists, 20, lo=lo, hi=hi)
    o_vals, o_items, scores = _oracle(rep, 100, users, lists, 20, lo=lo, hi=hi)
    _compare(items, vals, o_vals, o_items, scores)
    ok = items[items >= 0]
    assert ok.min() >= lo and ok.max() < hi


def test_topk_banned_bitmap_and_exhaustion():
    rep, lists = _case(50, 300, 64, seed=3, mask_deg=100)
    banned = list(range(0, 300, 2))
    users = list(range(50))
    items, vals = _run(rep, 50, users, lists, 128, banned=banned)
    o_vals, o_items, scores = _oracle(rep, 50, users, lists, 128, banned=banned)
    _compare(items, vals, o_vals, o_items, scores)
    assert (items == -1).any()                                      # some users run out of candidates


def test_topk_all_scores_equal_breaks_ties_by_item_id():
    rep = torch.ones(10 + 500, 64) * 0.5
    items, vals = _run(rep, 10, list(range(10)), None, 20)
    assert (items == np.arange(20)[None, :]).all() and np.allclose(vals, 16.0)


def test_hits_kernel():
    from igcn_cf_b200 import engine
    rec = torch.tensor([[1, 2, 3], [4, -1, 6], [7, 8, 9]], dtype=torch.int32, device=DEV)
    csr = engine.lists_to_csr([[9, 1], [], [9]], DEV)
    hit = engine.hit_matrix(rec, csr).cpu().numpy()
    assert hit.tolist() == [[1, 0, 0], [0, 0, 0], [0, 0, 1]] and hit.dtype == np.float32


@pytest.mark.parametrize('n_list', [1, 9, 700, 2500])
def test_exact_kernel_item_split_form_equals_plain(n_list):
    """The fallback form (catalogue cut into ranges + merge, used while the user list is short) returns exactly
    the plain kernel's lists; above the cap (2048 entries) it silently runs unsplit."""
    from igcn_cf_b200 import engine
    g = torch.Generator().manual_seed(n_list)
    n_users, n_items, k = 3000, 5000, 20
    rep = (torch.randn(n_users + n_items, 64, generator=g) * 0.1).to(DEV)
    rep[n_users + 100:n_users + 140] = rep[n_users + 7]            # exact ties across range boundaries
    rng = np.random.default_rng(n_list)
    lists = [sorted(rng.choice(n_items, size=int(rng.integers(0, 30)), replace=False).tolist()) for _ in range(n_users)]
    mask = engine.lists_to_csr(lists, DEV)
    users = torch.from_numpy(rng.choice(n_users, size=n_list, replace=n_list > n_users)).to(DEV)
    count = torch.tensor([n_list], dtype=torch.int32, device=DEV)
    rows = torch.arange(n_list, dtype=torch.int32, device=DEV)
    plain = engine.score_topk_exact(rep, users, n_users, n_items, k, mask, 50, 4900)
    scratch = torch.zeros(engine.FALLBACK_SPLIT_CAP * engine.FALLBACK_SPLITS * k, dtype=torch.int64, device=DEV)
    out = (torch.full((n_list, k), -7, dtype=torch.int32, device=DEV), torch.zeros((n_list, k), device=DEV))
    engine.score_topk_exact(rep, users, n_users, n_items, k, mask, 50, 4900, out=out, out_rows=rows, n_eval_dev=count,
                            split_keys=scratch)
    torch.cuda.synchronize()
    assert torch.equal(out[0], plain[0]) and torch.equal(out[1], plain[1])
