"""Synthetic bipartite interaction graphs shaped like the paper's datasets.

The real Gowalla / Yelp / Amazon-book dumps are not available offline, so every
benchmark and parity test runs on graphs produced here:

* user degrees ~ lognormal(sigma=1), rescaled so they sum to the requested
  number of interactions and clamped to >= 10 (the reference keeps only
  10-core users/items, /root/reference/run/process_dataset.py:7);
* item popularity proportional to rank^-0.8 under a random permutation;
* items are unique per user;
* each user's list is split 70/10/20 into train/val/test in generated order with
  the same integer arithmetic as /root/reference/dataset.py:107-111.

The on-disk format is the reference's (`train.txt`/`val.txt`/`test.txt`, one
line per user: "<user> <item> <item> ...", /root/reference/dataset.py:40-44).
"""
import os

import numpy as np

# name -> (n_users, n_items, n_interactions); SURVEY.md §8 table.
SHAPES = {
    'tiny': (300, 400, 6000),
    'small': (3000, 4000, 90000),
    'gowalla': (29858, 40981, 1027370),
    'yelp': (75173, 42706, 1931173),
    'amazon': (109730, 96421, 3181759),
}


class SynthSplit:
    """CSR-style per-user lists for the three splits (all int64 numpy arrays)."""

    def __init__(self, n_users, n_items, ptr, items, n_train, n_val):
        self.n_users = int(n_users)
        self.n_items = int(n_items)
        self.ptr = ptr          # [U+1] offsets into items
        self.items = items      # [E] item ids, per-user in generated order
        self.n_train = n_train  # [U] how many leading entries are train
        self.n_val = n_val      # [U] how many following entries are val

    def lists(self, which):
        """Python list-of-lists for one split ('train' | 'val' | 'test')."""
        out = []
        ptr, items = self.ptr, self.items
        for u in range(self.n_users):
            lo, hi = int(ptr[u]), int(ptr[u + 1])
            a = lo + int(self.n_train[u])
            b = a + int(self.n_val[u])
            if which == 'train':
                seg = items[lo:a]
            elif which == 'val':
                seg = items[a:b]
            else:
                seg = items[b:hi]
            out.append(seg.tolist())
        return out

    def csr(self, which):
        """(ptr[U+1], items) of one split as int64 numpy arrays (generated order kept)."""
        deg = np.diff(self.ptr)
        lo = self.ptr[:-1]
        if which == 'train':
            start, cnt = lo, self.n_train
        elif which == 'val':
            start, cnt = lo + self.n_train, self.n_val
        else:
            start, cnt = lo + self.n_train + self.n_val, deg - self.n_train - self.n_val
        ptr = np.zeros(self.n_users + 1, dtype=np.int64)
        np.cumsum(cnt, out=ptr[1:])
        # gather the segments
        idx = np.repeat(start - ptr[:-1], cnt) + np.arange(ptr[-1], dtype=np.int64)
        return ptr, self.items[idx]


def gen_synth(n_users, n_items, n_inter, seed=2021, min_degree=10):
    """Generate a power-law bipartite interaction set; returns a SynthSplit."""
    rng = np.random.default_rng(seed)
    raw = rng.lognormal(mean=0.0, sigma=1.0, size=n_users)
    deg = raw / raw.sum() * n_inter
    deg = np.maximum(min_degree, np.rint(deg)).astype(np.int64)
    deg = np.minimum(deg, max(min_degree, n_items // 4))

    rank = np.arange(1, n_items + 1, dtype=np.float64)
    pop = rank ** -0.8
    pop = pop[rng.permutation(n_items)]
    cdf = np.cumsum(pop / pop.sum())
    cdf[-1] = 1.0

    # oversample with replacement, drop duplicate (user, item) pairs, then keep
    # the first deg[u] survivors of each user in a random order.
    over = (deg * 1.6).astype(np.int64) + 8
    users = np.repeat(np.arange(n_users, dtype=np.int64), over)
    items = np.searchsorted(cdf, rng.random(users.shape[0]), side='right').astype(np.int64)
    items = np.minimum(items, n_items - 1)
    key = np.unique(users * n_items + items)
    users, items = key // n_items, key % n_items
    order = np.lexsort((rng.random(users.shape[0]), users))
    users, items = users[order], items[order]
    cnt = np.bincount(users, minlength=n_users)
    start = np.zeros(n_users + 1, dtype=np.int64)
    np.cumsum(cnt, out=start[1:])
    pos_in_user = np.arange(users.shape[0], dtype=np.int64) - start[users]
    keep = pos_in_user < deg[users]
    users, items = users[keep], items[keep]

    cnt = np.bincount(users, minlength=n_users).astype(np.int64)
    ptr = np.zeros(n_users + 1, dtype=np.int64)
    np.cumsum(cnt, out=ptr[1:])
    # same integer arithmetic as the reference split (dataset.py:107-111)
    n_train = (cnt * 0.7).astype(np.int64)
    n_test = (cnt * 0.2).astype(np.int64)
    # reference slices [n_train:-n_test]; with n_test == 0 that slice is empty and
    # test takes the whole list ([-0:]) -- min_degree >= 10 keeps n_test >= 2 here.
    n_val = cnt - n_train - n_test
    return SynthSplit(n_users, n_items, ptr, items, n_train, n_val)


def gen_named(name, seed=2021):
    u, i, e = SHAPES[name]
    return gen_synth(u, i, e, seed=seed)


def write_split(split, path):
    """Write train/val/test.txt in the reference's text format."""
    os.makedirs(path, exist_ok=True)
    for which in ('train', 'val', 'test'):
        ptr, items = split.csr(which)
        with open(os.path.join(path, which + '.txt'), 'w') as f:
            for u in range(split.n_users):
                seg = items[ptr[u]:ptr[u + 1]]
                f.write(' '.join([str(u)] + [str(int(i)) for i in seg]) + '\n')


def dropui(split, ratio=0.8):
    """Reduced split with the first ratio*U users and items < ratio*I.

    Follows /root/reference/run/dropui/dataset_dropui.py:7-27 (filter every split
    by item id, keep the leading users)."""
    n_users = int(split.n_users * ratio)
    n_items = int(split.n_items * ratio)
    ptrs, chunks, n_tr, n_va = [0], [], [], []
    tr, va, te = split.csr('train'), split.csr('val'), split.csr('test')
    for u in range(n_users):
        parts = []
        for ptr, items in (tr, va, te):
            seg = items[ptr[u]:ptr[u + 1]]
            parts.append(seg[seg < n_items])
        n_tr.append(len(parts[0]))
        n_va.append(len(parts[1]))
        chunks.extend(parts)
        ptrs.append(ptrs[-1] + sum(len(p) for p in parts))
    items = np.concatenate(chunks) if chunks else np.zeros(0, dtype=np.int64)
    return SynthSplit(n_users, n_items, np.array(ptrs, dtype=np.int64), items,
                      np.array(n_tr, dtype=np.int64), np.array(n_va, dtype=np.int64))


def dropit(split, ratio=0.8):
    """Reduced split keeping the first ratio of every user's train items.

    Follows /root/reference/run/dropit/dataset_dropit.py:6-9."""
    tr, va, te = split.csr('train'), split.csr('val'), split.csr('test')
    ptrs, chunks, n_tr, n_va = [0], [], [], []
    for u in range(split.n_users):
        t = tr[1][tr[0][u]:tr[0][u + 1]]
        t = t[:int(len(t) * ratio)]
        v = va[1][va[0][u]:va[0][u + 1]]
        e = te[1][te[0][u]:te[0][u + 1]]
        n_tr.append(len(t))
        n_va.append(len(v))
        chunks.extend([t, v, e])
        ptrs.append(ptrs[-1] + len(t) + len(v) + len(e))
    return SynthSplit(split.n_users, split.n_items, np.array(ptrs, dtype=np.int64),
                      np.concatenate(chunks), np.array(n_tr, dtype=np.int64),
                      np.array(n_va, dtype=np.int64))


def gen_device(n_users, n_items, n_inter, device, seed=2021, min_degree=10, chunk=1 << 26):
    """The same generative model as gen_synth, evaluated with torch ops ON THE DEVICE and returned as a
    symmetric CSR (igcn_cf_b200.graph.DeviceGraph) -- for graphs that are too large for Python lists
    (BASELINE.json config 5).  Every interaction is a train interaction (that config only asks for
    propagation + scoring).  Duplicate (user, item) draws are dropped, so the interaction count comes out a
    few per cent under n_inter."""
    import torch
    from .graph import DeviceGraph
    dev = torch.device(device)
    g = torch.Generator(device=dev).manual_seed(int(seed))
    raw = torch.empty(n_users, dtype=torch.float64, device=dev).log_normal_(0.0, 1.0, generator=g)
    deg = torch.clamp(torch.round(raw / raw.sum() * n_inter), min=min_degree, max=max(min_degree, n_items // 4)).long()
    del raw
    pop = torch.arange(1, n_items + 1, dtype=torch.float64, device=dev) ** -0.8
    pop = pop[torch.randperm(n_items, generator=g, device=dev)]
    cdf = torch.cumsum(pop / pop.sum(), 0)
    cdf[-1] = 1.0
    del pop
    users = torch.repeat_interleave(torch.arange(n_users, device=dev), deg)
    total = int(users.shape[0])
    keys = torch.empty(total, dtype=torch.int64, device=dev)
    for lo in range(0, total, chunk):
        hi = min(total, lo + chunk)
        r = torch.rand(hi - lo, dtype=torch.float64, device=dev, generator=g)
        items = torch.searchsorted(cdf, r, right=True).clamp_(max=n_items - 1)
        keys[lo:hi] = users[lo:hi] * n_items + items
        del r, items
    del users, cdf, deg
    keys = torch.unique(keys)                                     # sorted by (user, item), duplicates dropped
    u = torch.div(keys, n_items, rounding_mode='floor')
    it = keys - u * n_items
    del keys
    deg_u = torch.bincount(u, minlength=n_users)
    deg_i = torch.bincount(it, minlength=n_items)
    # item rows: sort the transposed pairs by (item, user)
    keys_t = torch.sort(it * n_users + u).values
    col_u = (it + n_users).to(torch.int32)
    del it, u
    col_i = (keys_t - torch.div(keys_t, n_users, rounding_mode='floor') * n_users).to(torch.int32)
    del keys_t
    rowptr = torch.zeros(n_users + n_items + 1, dtype=torch.int64, device=dev)
    torch.cumsum(torch.cat([deg_u, deg_i]), 0, out=rowptr[1:])
    col = torch.cat([col_u, col_i])
    return DeviceGraph(n_users, n_items, rowptr, col)
