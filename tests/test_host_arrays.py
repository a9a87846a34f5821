"""CPU tests of the array-level host helpers (no kernel is launched): CSR conversions used by the evaluation
path, the inductive_eval restriction, the dataset CSR cache, the row-class / chunk plan and the per-rank row
ranges of the sharded propagation."""
import numpy as np
import pytest
import torch

from igcn_cf_b200 import engine, graph
from igcn_cf_b200.dataset import get_dataset


def _lists(rng, n_users, n_items, max_len=12):
    return [rng.choice(n_items, size=int(rng.integers(0, max_len)), replace=False).tolist() for _ in range(n_users)]


def test_lists_to_arrays_merge_and_restrict_match_list_surgery():
    rng = np.random.default_rng(0)
    a, b = _lists(rng, 50, 40), _lists(rng, 50, 40)
    pa, ia = engine.lists_to_arrays(a)
    assert pa[-1] == len(ia) == sum(map(len, a)) and ia.tolist() == [i for x in a for i in x]
    pm, im = engine.merge_csr((pa, ia), engine.lists_to_arrays(b))
    assert [im[pm[u]:pm[u + 1]].tolist() for u in range(50)] == [x + y for x, y in zip(a, b)]
    # the reference's inductive_eval restriction (trainer.py:185-217) on lists vs on arrays
    n_old_users, n_old_items = 30, 25
    for users, keep in ((range(50), None), (range(n_old_users), None), (range(n_old_users, 50), None),
                        (range(50), lambda it: it < n_old_items), (range(50), lambda it: it >= n_old_items),
                        (range(n_old_users), lambda it: it < n_old_items)):
        want = []
        for u in range(50):
            if u not in users:
                want.append([])
            elif keep is None:
                want.append(list(a[u]))
            else:
                items = np.array(a[u], dtype=np.int64)
                want.append(items[keep(items)].tolist())
        lo_i, hi_i = (0, 40) if keep is None else ((0, n_old_items) if keep(np.int64(0)) else (n_old_items, 40))
        pr, ir = engine.restrict_csr((pa, ia), users[0], users[-1] + 1, lo_i, hi_i)
        assert [ir[pr[u]:pr[u + 1]].tolist() for u in range(50)] == want


def test_list_csr_from_arrays_equals_from_lists_and_tiles_cover_all_pairs():
    rng = np.random.default_rng(1)
    lists = _lists(rng, 300, 700, max_len=30)
    x = engine.ListCSR(lists, 'cpu')
    y = engine.ListCSR.from_arrays(*engine.lists_to_arrays(lists), 'cpu')
    assert np.array_equal(x.ptr_host, y.ptr_host) and np.array_equal(x.items_host, y.items_host)
    assert all(sorted(l) == x.items_host[x.ptr_host[u]:x.ptr_host[u + 1]].tolist() for u, l in enumerate(lists))
    tile_ptr, entries = x.tiles(700)
    tile_ptr, entries = tile_ptr.numpy(), entries.numpy().view(np.uint16)
    n_it = (700 + 255) // 256
    assert tile_ptr.shape == ((300 + 127) // 128, n_it + 1) and tile_ptr[-1, -1] == len(entries) == len(x.items_host)
    got = set()
    for ut in range(tile_ptr.shape[0]):
        for it in range(n_it):
            for e in entries[tile_ptr[ut, it]:tile_ptr[ut, it + 1]]:
                got.add((ut * 128 + (int(e) >> 8), it * 256 + (int(e) & 255)))
    assert got == {(u, i) for u, l in enumerate(lists) for i in l}
    sub = np.array([5, 299, 0, 77], dtype=np.int64)                    # user subset: rows are positions in the subset
    tp, en = x.tiles(700, sub)
    en = en.numpy().view(np.uint16)
    assert sorted((int(e) >> 8) for e in en) == sorted(r for r, u in enumerate(sub) for _ in lists[u])


def test_dataset_csr_cache_follows_list_identity():
    rng = np.random.default_rng(2)
    tr, va, te = _lists(rng, 20, 30), _lists(rng, 20, 30), _lists(rng, 20, 30)
    ds = get_dataset({'name': 'ListDataset', 'train': tr, 'val': va, 'test': te, 'n_items': 30, 'device': 'cpu'})
    first = ds.csr('test')
    assert ds.csr('test') is first                                       # cached
    assert first[1].tolist() == [i for x in ds.test_data for i in x]
    ds.test_data = [list(x) for x in ds.test_data]                       # the reference's inductive_eval replaces the list
    ds.test_data[3] = []
    second = ds.csr('test')
    assert second is not first and second[0][4] - second[0][3] == 0


def test_row_classes_and_chunk_plan():
    deg = np.array([0, 1, 64, 65, 256, 257, 700, 3, 128, 129], dtype=np.int64)
    rowptr = np.concatenate([[0], np.cumsum(deg)])
    col = np.zeros(int(rowptr[-1]), dtype=np.int32)
    csr = graph.CsrDevice(rowptr, col, None, 10, 'cpu')
    order = csr.row_order.numpy()
    assert (np.diff(deg[order]) <= 0).all()                              # longest first
    assert csr.n_long == 2 and csr.n_medium == 4                         # > 256: {257, 700}; (64, 256]: {65, 256, 128, 129}
    assert set(order[:2].tolist()) == {5, 6} and set(order[2:6].tolist()) == {3, 4, 8, 9}
    cr, cb, cl, cf, cc = (t.numpy() for t in csr._plan)
    assert csr.n_chunks == 3 + 6 and cl.sum() == 257 + 700
    for row in (5, 6):
        sel = cr == row
        assert cc[sel][0] == sel.sum() and (cb[sel] == rowptr[row] + 128 * np.arange(sel.sum())).all()
        assert cl[sel][:-1].tolist() == [128] * (sel.sum() - 1) and cf[sel].min() == np.flatnonzero(sel)[0]


def test_row_ranges_cover_every_row_once_and_balance_each_half():
    rng = np.random.default_rng(3)
    n_users, n_items = 500, 200
    deg = np.concatenate([rng.integers(1, 40, size=n_users), rng.integers(1, 400, size=n_items)]).astype(np.int64)
    rowptr = np.concatenate([[0], np.cumsum(deg)])
    assert graph._row_ranges(rowptr, n_users, None) == [(0, n_users + n_items)]
    for world in (2, 3, 8):
        seen = np.zeros(n_users + n_items, dtype=int)
        user_nnz, item_nnz = [], []
        for rank in range(world):
            (u0, u1), (i0, i1) = graph._row_ranges(rowptr, n_users, (rank, world))
            assert 0 <= u0 <= u1 <= n_users <= i0 <= i1 <= n_users + n_items
            seen[u0:u1] += 1
            seen[i0:i1] += 1
            user_nnz.append(rowptr[u1] - rowptr[u0])
            item_nnz.append(rowptr[i1] - rowptr[i0])
        assert (seen == 1).all()
        assert max(user_nnz) - min(user_nnz) <= 2 * 40 + 8 and max(item_nnz) - min(item_nnz) <= 2 * 400 + 8


def test_popularity_and_identity_map_host_side():
    from igcn_cf_b200.model import IdentityMap
    m = IdentityMap(5)
    assert len(m) == 5 and 4 in m and 5 not in m and m[3] == 3 and list(m.keys()) == [0, 1, 2, 3, 4]
    assert dict(m.items()) == {k: k for k in range(5)}
    try:
        m[7]
        assert False
    except KeyError:
        pass
    assert torch.equal(torch.arange(3), torch.arange(3))


def test_item_order_positions_mask_tiles_and_position_bitmap():
    """engine.ItemOrder: perm / pos are inverse permutations; ListCSR.tiles buckets the seen pairs by the item's
    POSITION in the scan order; ranges and banned items become one bitmap over positions."""
    import torch
    from igcn_cf_b200 import engine
    from igcn_cf_b200.graph import _pack_bits
    rng = np.random.default_rng(0)
    n_users, n_items = 300, 700
    score = rng.integers(0, 50, size=n_items)
    order = engine.ItemOrder.by_score(score, 'cpu')
    assert np.array_equal(order.pos_host[order.perm_host], np.arange(n_items))
    assert (np.diff(score[order.perm_host]) <= 0).all()                       # descending, ties by id (stable)
    ties = np.nonzero(np.diff(score[order.perm_host]) == 0)[0]
    assert (order.perm_host[ties] < order.perm_host[ties + 1]).all()
    with pytest.raises(ValueError):
        engine.ItemOrder(np.array([0, 0, 1]), 'cpu')
    lists = [sorted(rng.choice(n_items, size=int(rng.integers(0, 9)), replace=False).tolist()) for _ in range(n_users)]
    csr = engine.lists_to_csr(lists, 'cpu')
    tile_ptr, ent = csr.tiles(n_items, None, order)
    tile_ptr, ent = tile_ptr.numpy(), ent.numpy().view(np.uint16)
    n_it = (n_items + 255) // 256
    got = set()
    for ut in range(tile_ptr.shape[0]):
        for t in range(n_it):
            for e in ent[tile_ptr[ut, t]:tile_ptr[ut, t + 1]]:
                got.add((ut * 128 + (int(e) >> 8), int(order.perm_host[t * 256 + (int(e) & 255)])))
    assert got == {(u, i) for u, l in enumerate(lists) for i in l}
    flags = np.zeros(n_items, dtype=bool)
    flags[::7] = True
    bits = order.position_bits(n_items, 100, 650, _pack_bits(flags, 'cpu')).numpy().view(np.uint32)
    for p in range(n_items):
        item = order.perm_host[p]
        want = item < 100 or item >= 650 or item % 7 == 0
        assert bool((bits[p >> 5] >> (p & 31)) & 1) == want
    assert order.position_bits(n_items, 0, n_items, None) is None
