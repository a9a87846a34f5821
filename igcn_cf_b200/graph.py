"""Device-resident graph structures behind `norm_adj` and `feat_mat`.

`NormAdj` is what `LightGCN.generate_graph` returns here and `TemplateFeat` what
`IGCN.generate_feat` returns.  Both keep the reference's observable surface (`shape`,
`indices()`, `values()`, `_nnz()`, `to_sparse_coo()`; reference model.py:85-94, 386-421) but store
one CSR per owned row range on the GPU (int64 rowptr, int32 col, fp32 val; one range on a single GPU,
a user slice + an item slice per rank when rows are sharded) plus the row-class / long-row chunk plan
the kernels use (include/igcn_b200.h, `igcn_csr`).  They are built ONCE per generate_* call, not once per
`get_rep` as the reference's `dgl.graph(...)` is (model.py:99-100, 439-440).

HBM layout: rowptr[N+1] int64 | col[nnz] int32 | val[nnz] fp32 | chunk plan (5 small int arrays)
| partial[n_chunks, D] fp32 | counters[2 * n_chunks] int32.  The INMO layer reuses rowptr/col (its
pattern is the adjacency pattern filtered by template membership) and adds
tmpl[N] int32 (only when some node is not a template) and rowscale[N] fp32.
"""
import ctypes as C
import itertools
import os

import numpy as np
import scipy.sparse as sp
import torch

from . import _lib

LONG_THRESHOLD = 256     # rows with more non-zeros than this are split ...
CHUNK = 128              # ... into chunks of this many non-zeros
MEDIUM_NNZ = 64          # IGCN_MEDIUM_NNZ: rows above it (and not long) get a whole warp


def train_pairs_of(dataset):
    """[E, 2] int64 array of (user, item) train interactions of a reference-style dataset."""
    pairs = getattr(dataset, 'train_pairs', None)
    if pairs is None:
        pairs = np.asarray(dataset.train_array, dtype=np.int64).reshape(-1, 2)
    return pairs


def build_adjacency(n_users, n_items, pairs):
    """scipy CSR of the symmetric bipartite adjacency, duplicates summed (utils.py:41-49)."""
    n = n_users + n_items
    u, i = pairs[:, 0], pairs[:, 1] + n_users
    rows = np.concatenate([u, i])
    cols = np.concatenate([i, u])
    adj = sp.csr_matrix((np.ones(rows.shape[0], dtype=np.float32), (rows, cols)), shape=(n, n))
    adj.sum_duplicates()
    adj.sort_indices()
    return adj


def chunk_plan(rowptr, threshold=LONG_THRESHOLD, chunk=CHUNK):
    """Split rows longer than `threshold` into chunks of `chunk` non-zeros (host, numpy)."""
    deg = np.diff(rowptr)
    long_rows = np.nonzero(deg > threshold)[0]
    n_ch = (deg[long_rows] + chunk - 1) // chunk
    total = int(n_ch.sum())
    first = np.zeros(len(long_rows) + 1, dtype=np.int64)
    np.cumsum(n_ch, out=first[1:])
    chunk_row = np.repeat(long_rows, n_ch).astype(np.int32)
    chunk_first = np.repeat(first[:-1], n_ch).astype(np.int32)
    chunk_count = np.repeat(n_ch, n_ch).astype(np.int32)
    k = np.arange(total, dtype=np.int64) - chunk_first
    chunk_begin = rowptr[chunk_row] + k * chunk
    chunk_len = np.minimum(chunk, rowptr[chunk_row.astype(np.int64) + 1] - chunk_begin).astype(np.int32)
    return chunk_row, chunk_begin.astype(np.int64), chunk_len, chunk_first, chunk_count


class CsrDevice:
    """One CSR block on the GPU + chunk plan + per-D scratch; produces the `igcn_csr` struct."""

    def __init__(self, rowptr, col, val, n_cols, device, threshold=LONG_THRESHOLD, chunk=CHUNK, split_at=None):
        """rowptr: host int64 array; col / val: host arrays, or tensors already on `device` (scale-out
        graphs are generated on the GPU and never visit the host).  split_at: local index of the first ITEM row of
        this block (kept for callers; the visiting order does not use it)."""
        self.device = torch.device(device)   # kernels need CUDA; a CPU device only supports the views
        self.n_rows = int(len(rowptr) - 1)
        self.n_cols = int(n_cols)
        self.nnz = int(rowptr[-1])
        self.rowptr_host = np.ascontiguousarray(rowptr, dtype=np.int64)
        self.rowptr = torch.from_numpy(self.rowptr_host).to(self.device)
        if torch.is_tensor(col):
            self._col_host = None
            self.col = col.to(device=self.device, dtype=torch.int32).contiguous()
            self.val = None if val is None else val.to(device=self.device, dtype=torch.float32).contiguous()
        else:
            self._col_host = np.ascontiguousarray(col, dtype=np.int32)
            self.col = torch.from_numpy(self._col_host).to(self.device)
            self.val = None if val is None else torch.from_numpy(np.ascontiguousarray(val, dtype=np.float32)).to(self.device)
        self.threshold = int(threshold)
        plan = chunk_plan(self.rowptr_host, threshold, chunk)
        self.n_chunks = int(len(plan[0]))
        self._plan = [torch.from_numpy(np.ascontiguousarray(a)).to(self.device) for a in plan]
        self.counters = torch.zeros(max(1, 2 * self.n_chunks), dtype=torch.int32, device=self.device)   # chunk groups + rows
        # visiting order: longest rows first, ties by row id (stable) -> rows sharing a warp are alike.
        # (Keeping the two halves of the bipartite graph apart inside each row class -- item rows, then user rows, so
        # that the cold user lines do not flush the popular item lines out of L1 -- was measured in round 2:
        # Yelp-shaped step 0.405 -> 0.399 ms, Gowalla-shaped 0.326 -> 0.349 ms.  Not used.)
        deg = np.diff(self.rowptr_host)
        self.n_long = int((deg > self.threshold).sum())
        self.n_medium = max(0, int((deg > MEDIUM_NNZ).sum()) - self.n_long)
        self.row_order = torch.from_numpy(np.argsort(-deg, kind='stable').astype(np.int32)).to(self.device)
        self._partial = {}
        self._structs = {}

    @property
    def col_host(self):
        """Host copy of the column indices (fetched once from the device for CSRs that were built there)."""
        if self._col_host is None:
            self._col_host = self.col.cpu().numpy()
        return self._col_host

    def with_values(self, val):
        """Same pattern/plan, different (or no) value array; shares index memory."""
        other = object.__new__(CsrDevice)
        other.__dict__.update(self.__dict__)
        other.val = val
        other.counters = torch.zeros_like(self.counters)
        other._partial, other._structs = {}, {}
        return other

    def struct(self, D):
        s = self._structs.get(D)
        if s is None:
            partial = torch.empty((max(1, self.n_chunks), D), dtype=torch.float32, device=self.device)
            self._partial[D] = partial
            cr, cb, cl, cf, cc = self._plan
            s = _lib.CsrStruct(self.n_rows, self.n_cols, self.nnz, self.rowptr.data_ptr(), self.col.data_ptr(),
                               None if self.val is None else self.val.data_ptr(), self.threshold, self.n_chunks,
                               cr.data_ptr(), cb.data_ptr(), cl.data_ptr(), cf.data_ptr(), cc.data_ptr(),
                               partial.data_ptr(), self.counters.data_ptr(), self.row_order.data_ptr(),
                               self.n_long, self.n_medium)
            self._structs[D] = s
        return C.byref(s)


class _SparseView:
    """Reference-compatible read-only view (`shape`, `indices()`, `values()`, `_nnz()`)."""
    shape = None

    def _coo(self):
        raise NotImplementedError

    def indices(self):
        return self._coo()[0]

    def values(self):
        return self._coo()[1]

    def _nnz(self):
        return int(self._coo()[0].shape[1])

    def size(self, dim=None):
        return self.shape if dim is None else self.shape[dim]

    def to_sparse_coo(self):
        idx, val = self._coo()
        return torch.sparse_coo_tensor(idx, val, self.shape, is_coalesced=True)


class RowBlock:
    """Rows [row0, row1) of a graph as their own CSR.  col_blocks: the same rows cut into column ranges (one CsrDevice
    each, ascending), present only when the table these rows gather from is far larger than L2 (column_blocks)."""

    def __init__(self, row0, row1, csr, col_blocks=None):
        self.row0, self.row1, self.csr = int(row0), int(row1), csr
        self.col_blocks = col_blocks


# Column blocking (BASELINE.json configs[4], the HBM-resident graphs): the item rows of the scale-out graph gather
# 496 M rows of 256 bytes out of a 2.6 GB user table, far more than L2 holds.  Cutting the columns into ranges whose
# slice of the table stays in L2 and running the layer once per range, each pass adding to the previous partial result,
# turns the DRAM gathers into one streaming read of the table plus a read-modify-write of the (10x smaller) output per
# pass.  Measured (profiles/r02_colblocks.log, one layer, item rows): unblocked 19.4 ms; 26 ranges of 96 MB 17.5 ms,
# 51 x 48 MB 20.8 ms, 102 x 24 MB 30.8 ms, 204 x 12 MB 50.4 ms = 11 ms + 0.19 ms per pass: with the table slice in L2
# the gathers still move nnz x 256 B = 127 GB through the L2 -> SM fabric, whose ~11.5 TB/s is the floor (the user
# rows, which gather from a 256 MB item table, run at 9.3 TB/s without any blocking).  So blocking buys 10 %, not
# the 3x the DRAM traffic ratio suggests; 96 MB ranges are the default.
# Ranges are global column intervals of a fixed width, so a row's partial sums are the same whatever the row
# sharding: results stay GPU-count independent (they differ in the last bits from the unblocked order, which is why
# blocking is decided by table size alone, never by rank count).
COL_BLOCK_BYTES = int(os.environ.get('IGCN_COL_BLOCK_MB', 96)) << 20       # slice of the gathered table per pass
COL_BLOCK_MIN_TABLE = int(os.environ.get('IGCN_COL_BLOCK_MIN_TABLE_MB', 1024)) << 20   # tables below this are not blocked


def column_block_width(n_table_rows, row_bytes=256):
    """Rows of the gathered table per column block, or 0 = do not block."""
    if COL_BLOCK_BYTES <= 0 or n_table_rows * row_bytes < COL_BLOCK_MIN_TABLE:
        return 0
    return max(1, COL_BLOCK_BYTES // row_bytes)


def column_blocks(rowptr_host, col, val, col_lo, col_hi, width, n_cols, device):
    """Cut a CSR (host rowptr int64 [rows + 1], device col int32 / val fp32, columns sorted inside a row) into the
    column ranges [col_lo + b * width, +width): one CsrDevice per range over the SAME rows, entries in row-major,
    column-ascending order.  One stable sort by range id, one histogram, one scan per range."""
    n_rows = len(rowptr_host) - 1
    nb = max(1, -(-(col_hi - col_lo) // width))
    bid = torch.div(col - col_lo, width, rounding_mode='floor').to(torch.int16 if nb < 32768 else torch.int32)
    order = torch.sort(bid, stable=True)[1]
    counts = torch.diff(torch.from_numpy(np.ascontiguousarray(rowptr_host)).to(device))
    rows = torch.repeat_interleave(torch.arange(n_rows, device=device, dtype=torch.int32), counts)
    per = torch.bincount(bid.long() * n_rows + rows.long(), minlength=nb * n_rows).view(nb, n_rows)
    del rows, bid
    sizes = per.sum(dim=1).cpu().numpy()
    starts = np.concatenate([[0], np.cumsum(sizes)])
    out = []
    for b in range(nb):
        rp = np.zeros(n_rows + 1, dtype=np.int64)
        np.cumsum(per[b].cpu().numpy(), out=rp[1:])
        idx = order[int(starts[b]):int(starts[b + 1])]
        out.append(CsrDevice(rp, col[idx].contiguous(), None if val is None else val[idx].contiguous(), n_cols, device))
    return out


def _row_ranges(rowptr, n_users, shard):
    """Row ranges this rank computes; shard = (rank, world) or None (everything).  A rank gets one slice of
    the USER rows and one slice of the ITEM rows, each balanced by non-zeros: the two halves of the
    bipartite adjacency cost differently per non-zero (user rows gather from the item table, item rows from
    the usually much larger user table), so balancing them separately balances the ranks."""
    n = len(rowptr) - 1
    if shard is None:
        return [(0, n)]
    from .dist import shard_bounds
    rank, world = shard
    ub = shard_bounds(rowptr[:n_users + 1], world)
    ib = shard_bounds(rowptr[n_users:] - rowptr[n_users], world) + n_users
    return [(int(ub[rank]), int(ub[rank + 1])), (int(ib[rank]), int(ib[rank + 1]))]


def _block_csr(rowptr, col, val, row0, row1, n_cols, device, n_users=None):
    lo, hi = int(rowptr[row0]), int(rowptr[row1])
    return CsrDevice(rowptr[row0:row1 + 1] - lo, col[lo:hi], None if val is None else val[lo:hi], n_cols, device,
                     split_at=None if n_users is None else n_users - row0)


_UID = itertools.count(1)


class _Blocked:
    """Mixin: `blocks` (list of RowBlock) + single-block compatibility attributes."""
    blocks = ()
    uid = 0

    def _set_blocks(self, blocks):
        # a process-wide serial number identifies the graph object in cache keys (id() values are reused by CPython
        # once an object is freed; a captured CUDA graph keyed on one would replay dangling pointers)
        self.uid = next(_UID)
        self.blocks = blocks
        self.csr = blocks[0].csr                     # single-GPU view (tests, sampler, chunk statistics)
        self.row0, self.row1 = blocks[0].row0, blocks[0].row1
        self.device = self.csr.device

    @property
    def local_rows(self):
        return sum(b.csr.n_rows for b in self.blocks)

    @property
    def local_nnz(self):
        return sum(b.csr.nnz for b in self.blocks)

    def block_key(self):
        return tuple((b.row0, b.row1) for b in self.blocks)


class DeviceGraph:
    """A bipartite train graph that already is a symmetric CSR ON THE DEVICE: rowptr int64 [N+1] (also kept
    on the host for the chunk plan and the shard bounds), col int32 [nnz] sorted inside each row, users
    first.  Produced by igcn_cf_b200.synth.gen_device for graphs too large to pass through Python lists
    (BASELINE.json config 5: 10 M users, 1 M items, 500 M interactions)."""

    def __init__(self, n_users, n_items, rowptr, col, mult=None):
        self.n_users, self.n_items = int(n_users), int(n_items)
        self.rowptr, self.col = rowptr, col
        self.mult = mult                      # fp32 [nnz] multiplicity of duplicated interactions, None = all 1
        self.rowptr_host = rowptr.cpu().numpy()
        self.device = col.device
        self._rows = None

    @classmethod
    def from_pairs(cls, n_users, n_items, pairs, device):
        """Edge list -> symmetric CSR on the device (replaces the scipy build of
        utils.py:41-49 / model.py:85-94 on the inductive-update path): [E, 2] (user, item) pairs, duplicates
        counted like scipy's sum_duplicates.  Two sorts, two histograms and one scan; the only host traffic is
        the upload of the pairs and the read-back of rowptr for the row plan."""
        n_users, n_items = int(n_users), int(n_items)
        p = torch.as_tensor(np.ascontiguousarray(pairs, dtype=np.int64) if not torch.is_tensor(pairs) else pairs)
        p = p.to(device=device, dtype=torch.int64).reshape(-1, 2)
        key = torch.sort(p[:, 0] * n_items + p[:, 1])[0]                 # (user, item) order = user rows of the CSR
        key, cnt = torch.unique_consecutive(key, return_counts=True)
        u = torch.div(key, n_items, rounding_mode='floor')
        i = key - u * n_items
        key2, perm = torch.sort(i * n_users + u)                         # (item, user) order = item rows
        i2 = torch.div(key2, n_users, rounding_mode='floor')
        u2 = key2 - i2 * n_users
        n = n_users + n_items
        rowptr = torch.zeros(n + 1, dtype=torch.int64, device=p.device)
        torch.cumsum(torch.cat([torch.bincount(u, minlength=n_users), torch.bincount(i2, minlength=n_items)]), 0, out=rowptr[1:])
        col = torch.cat([i + n_users, u2]).to(torch.int32)
        mult = None
        if key.shape[0] != p.shape[0]:                                   # some pair occurs more than once
            mult = torch.cat([cnt, cnt[perm]]).to(torch.float32)
        return cls(n_users, n_items, rowptr, col, mult)

    def rows(self):
        """Row id of every non-zero (int64 [nnz], cached)."""
        if self._rows is None:
            n = self.n_users + self.n_items
            self._rows = torch.repeat_interleave(torch.arange(n, device=self.device), self.rowptr[1:] - self.rowptr[:-1])
        return self._rows

    def degrees_host(self):
        """fp32 row sums of the adjacency (duplicates counted) as a host array: np.diff(rowptr) without
        duplicates.  d^-1/2 is taken on the host with numpy like the reference does (utils.py:44-46) so that
        the values are bit-identical to the host-built graph whatever the GPU's pow rounds to."""
        if self.mult is None:
            return np.diff(self.rowptr_host).astype(np.float32)
        n = self.n_users + self.n_items
        deg = torch.zeros(n, dtype=torch.float64, device=self.device).index_add_(0, self.rows(), self.mult.double())
        return deg.cpu().numpy().astype(np.float32)

    @property
    def n_interactions(self):
        return int(self.rowptr_host[self.n_users])

    def block(self, row0, row1):
        lo, hi = int(self.rowptr_host[row0]), int(self.rowptr_host[row1])
        return self.rowptr_host[row0:row1 + 1] - lo, self.col[lo:hi], lo, hi


class NormAdj(_SparseView, _Blocked):
    """D^-1/2 A D^-1/2 (deg clamped to >= 1) as a device CSR; reference model.py:85-94."""

    @classmethod
    def from_device(cls, dg, shard=None):
        """Same object from a DeviceGraph: values computed on the GPU with the reference's rounding order
        ((d_r * 1) * d_c in fp32), only this rank's row block is kept.  When the user table is far larger than L2
        the item rows are additionally cut into column blocks (RowBlock.col_blocks, see column_blocks)."""
        self = object.__new__(cls)
        n = dg.n_users + dg.n_items
        self.n_users, self.n_items = dg.n_users, dg.n_items
        self.shape = torch.Size([n, n])
        self.rowptr_full = dg.rowptr_host
        self.col_full = self.val_full = self.multiplicity_host = None
        self.nnz = int(self.rowptr_full[-1])
        self._dg = dg
        deg = np.maximum(np.float32(1.), dg.degrees_host())
        d_inv = torch.from_numpy(np.power(deg, np.float32(-0.5)).astype(np.float32)).to(dg.device)
        blocks = []
        width = column_block_width(dg.n_users)
        ranges = _row_ranges(self.rowptr_full, dg.n_users, shard)
        if width and shard is None:
            ranges = [(0, dg.n_users), (dg.n_users, n)]      # the item rows become a block of their own
        for row0, row1 in ranges:
            rp, col, lo, hi = dg.block(row0, row1)
            rows = torch.repeat_interleave(torch.arange(row0, row1, device=dg.device),
                                           dg.rowptr[row0 + 1:row1 + 1] - dg.rowptr[row0:row1])
            # same rounding sequence as d_mat.dot(adj).dot(d_mat) in fp32: (d_r * a) * d_c
            left = d_inv[rows] if dg.mult is None else d_inv[rows] * dg.mult[lo:hi]
            val = left * d_inv[col.long()]
            del rows, left
            cbs = None
            if width and row0 >= dg.n_users and row1 > row0:
                cbs = column_blocks(rp, col, val, 0, dg.n_users, width, n, dg.device)
            blocks.append(RowBlock(row0, row1, CsrDevice(rp, col, val, n, dg.device, split_at=dg.n_users - row0), cbs))
        self._set_blocks(blocks)
        self._coo_cache = None
        self._sampler = (dg.rowptr[:dg.n_users + 1], dg.col[:dg.n_interactions])
        return self

    # host copies of a device-built graph, fetched on first use (compat views, tools)
    @property
    def col_full(self):
        if self._col_full is None and getattr(self, '_dg', None) is not None:
            self._col_full = self._dg.col.cpu().numpy()
        return self._col_full

    @col_full.setter
    def col_full(self, v):
        self._col_full = v

    @property
    def val_full(self):
        if self._val_full is None and getattr(self, '_dg', None) is not None:
            if len(self.blocks) != 1:
                raise RuntimeError('val_full of a row-sharded device-built graph is not available')
            self._val_full = self.csr.val.cpu().numpy()
        return self._val_full

    @val_full.setter
    def val_full(self, v):
        self._val_full = v

    def __init__(self, n_users, n_items, pairs, device, shard=None):
        adj = build_adjacency(n_users, n_items, pairs)
        deg = np.maximum(1., np.asarray(adj.sum(axis=1)).squeeze()).astype(np.float32)
        d_inv = np.power(deg, np.float32(-0.5)).astype(np.float32)
        rows = np.repeat(np.arange(adj.shape[0], dtype=np.int64), np.diff(adj.indptr))
        # same rounding sequence as d_mat.dot(adj).dot(d_mat) in fp32: (d_r * a) * d_c
        val = (d_inv[rows] * adj.data.astype(np.float32)) * d_inv[adj.indices]
        self.n_users, self.n_items = int(n_users), int(n_items)
        self.shape = torch.Size([adj.shape[0], adj.shape[1]])
        self.multiplicity_host = adj.data.astype(np.float32)
        self.rowptr_full = adj.indptr.astype(np.int64)
        self.col_full = adj.indices.astype(np.int32)
        self.val_full = val
        self.nnz = int(self.rowptr_full[-1])
        self._set_blocks([RowBlock(r0, r1, _block_csr(self.rowptr_full, self.col_full, val, r0, r1, adj.shape[0], device, n_users))
                          for r0, r1 in _row_ranges(self.rowptr_full, n_users, shard)])
        self._coo_cache = None
        self._sampler = None

    def sampler_csr(self):
        """(rowptr, col) on the device covering ALL user rows: the triple sampler is replicated on every
        rank, so a row-sharded adjacency keeps a separate copy of the user half of the pattern."""
        if self._sampler is None:
            if len(self.blocks) == 1 and self.row0 == 0 and self.row1 >= self.n_users:
                self._sampler = (self.csr.rowptr, self.csr.col)
            else:
                rp = self.rowptr_full[:self.n_users + 1]
                self._sampler = (torch.from_numpy(rp.copy()).to(self.device),
                                 torch.from_numpy(self.col_full[:rp[-1]].copy()).to(self.device))
        return self._sampler

    def _coo(self):
        if self._coo_cache is None:
            rows = np.repeat(np.arange(self.shape[0], dtype=np.int64), np.diff(self.rowptr_full))
            idx = torch.from_numpy(np.stack([rows, self.col_full.astype(np.int64)])).to(self.device)
            self._coo_cache = (idx, torch.from_numpy(self.val_full).to(self.device))
        return self._coo_cache


class TemplateFeat(_SparseView, _Blocked):
    """The INMO template incidence matrix `feat_mat` (reference model.py:386-421, 374-377).

    Row r holds one entry per neighbour c of r whose node is a template (column tmpl[c]) plus one
    global-template entry (column T_u+T_i for users, T_u+T_i+1 for items); every entry of row r
    has the value row_sum[r] ** ((alpha-1)/2 - 1/2).  Stored as: the adjacency pattern CSR (shared
    layout with NormAdj), tmpl[N] (None when every node is a template and tmpl is the identity),
    row_sum[N], rowscale[N]."""

    @classmethod
    def from_device(cls, dg, adj=None, shard=None, user_tmpl=None, item_tmpl=None, t_users=None, t_items=None):
        """Template structure of a DeviceGraph; shares the index arrays of `adj` (a NormAdj built from the same
        graph and shard) when given.  user_tmpl / item_tmpl: host int arrays, template id of every user / item
        or -1 (model.py:392-412 as arrays); None = every node is its own template (feature_ratio == 1).  The
        membership count behind row_sum (model.py:419) is one gather + one segmented integer sum on the device."""
        self = object.__new__(cls)
        n = dg.n_users + dg.n_items
        self.n_users, self.n_items = dg.n_users, dg.n_items
        self._dg = dg
        self.rowptr_full, self.col_full = dg.rowptr_host, None
        if user_tmpl is not None:
            self.t_users, self.t_items = int(t_users), int(t_items)
            tmpl = np.concatenate([user_tmpl, np.where(item_tmpl >= 0, item_tmpl + self.t_users, -1)]).astype(np.int32)
            identity = (self.t_users == dg.n_users and self.t_items == dg.n_items
                        and np.array_equal(tmpl, np.arange(n, dtype=np.int32)))
            if not identity:
                return self._finish_partial(dg, tmpl, shard, adj)
        self.t_users, self.t_items = dg.n_users, dg.n_items
        self.shape = torch.Size([n, n + 2])
        self.tmpl_host = None if user_tmpl is None else tmpl
        ranges = _row_ranges(self.rowptr_full, dg.n_users, shard)
        if adj is not None and list(adj.block_key()) == ranges:
            blocks = [RowBlock(b.row0, b.row1, b.csr.with_values(None)) for b in adj.blocks]
        else:
            blocks = [RowBlock(r0, r1, CsrDevice(dg.block(r0, r1)[0], dg.block(r0, r1)[1], None, n, dg.device,
                                                 split_at=dg.n_users - r0))
                      for r0, r1 in ranges]
        self._set_blocks(blocks)
        self.tmpl = None
        self.row_sum = torch.from_numpy(dg.degrees_host() + np.float32(1.)).to(self.device)
        self.rowscale = torch.ones(n, dtype=torch.float32, device=self.device)
        self.glob_user, self.glob_item = n, n + 1
        self._order = self._tperm = None
        return self

    def _finish_partial(self, dg, tmpl, shard, adj):
        """from_device when some node is not a template: keep the whole adjacency pattern, tmpl[] filters."""
        n = dg.n_users + dg.n_items
        self.shape = torch.Size([n, self.t_users + self.t_items + 2])
        self.tmpl_host = tmpl
        ranges = _row_ranges(self.rowptr_full, dg.n_users, shard)
        if adj is not None and list(adj.block_key()) == ranges:
            blocks = [RowBlock(b.row0, b.row1, b.csr.with_values(None)) for b in adj.blocks]
        else:
            blocks = [RowBlock(r0, r1, CsrDevice(dg.block(r0, r1)[0], dg.block(r0, r1)[1], None, n, dg.device,
                                                 split_at=dg.n_users - r0))
                      for r0, r1 in ranges]
        self._set_blocks(blocks)
        self.tmpl = torch.from_numpy(tmpl).to(self.device)
        member = (self.tmpl[dg.col.long()] >= 0).to(torch.float64)
        if dg.mult is not None:
            member = member * dg.mult.double()
        row_sum = torch.zeros(n, dtype=torch.float64, device=self.device).index_add_(0, dg.rows(), member)
        self.row_sum = row_sum.to(torch.float32) + 1.0
        self.rowscale = torch.ones(n, dtype=torch.float32, device=self.device)
        self.glob_user = self.t_users + self.t_items
        self.glob_item = self.t_users + self.t_items + 1
        self._order = self._tperm = None
        return self

    @property
    def col_full(self):
        if self._col_full is None and getattr(self, '_dg', None) is not None:
            self._col_full = self._dg.col.cpu().numpy()
        return self._col_full

    @col_full.setter
    def col_full(self, v):
        self._col_full = v

    @property
    def tmpl_host(self):
        """Template id of every node on the host (the identity when every node is a template)."""
        if self._tmpl_host is None:
            self._tmpl_host = np.arange(self.n_users + self.n_items, dtype=np.int32)
        return self._tmpl_host

    @tmpl_host.setter
    def tmpl_host(self, v):
        self._tmpl_host = v

    def __init__(self, n_users, n_items, pairs, user_tmpl, item_tmpl, t_users, t_items, device, shard=None):
        adj = build_adjacency(n_users, n_items, pairs)
        n = adj.shape[0]
        self.n_users, self.n_items = int(n_users), int(n_items)
        self.t_users, self.t_items = int(t_users), int(t_items)
        self.shape = torch.Size([n, self.t_users + self.t_items + 2])
        tmpl = np.concatenate([user_tmpl, np.where(item_tmpl >= 0, item_tmpl + self.t_users, -1)]).astype(np.int32)
        identity = (self.t_users == n_users and self.t_items == n_items
                    and np.array_equal(tmpl, np.arange(n, dtype=np.int32)))
        self.tmpl_host = tmpl
        # row_sum counts duplicates like feat.sum(axis=1) does (model.py:419); +1 = global template
        member = (tmpl[adj.indices] >= 0) * adj.data.astype(np.float64)
        rows = np.repeat(np.arange(n, dtype=np.int64), np.diff(adj.indptr))
        row_sum = np.bincount(rows, weights=member, minlength=n).astype(np.float32) + np.float32(1.)
        self.rowptr_full = adj.indptr.astype(np.int64)
        self.col_full = adj.indices.astype(np.int32)
        self._set_blocks([RowBlock(r0, r1, _block_csr(self.rowptr_full, self.col_full, None, r0, r1, n, device, n_users))
                          for r0, r1 in _row_ranges(self.rowptr_full, n_users, shard)])
        self.tmpl = None if identity else torch.from_numpy(tmpl).to(self.device)
        self.row_sum = torch.from_numpy(row_sum.astype(np.float32)).to(self.device)
        self.rowscale = torch.ones(n, dtype=torch.float32, device=self.device)
        self.glob_user = self.t_users + self.t_items
        self.glob_item = self.t_users + self.t_items + 1
        self._order = None
        self._tperm = None

    def set_alpha(self, alpha, row_sum=None):
        """update_feat_mat (model.py:374-377): value of every entry of row r.  Written in place so
        captured CUDA graphs keep pointing at live memory across the per-epoch anneal."""
        if row_sum is not None:
            self.row_sum = row_sum
        new = torch.pow(self.row_sum, (alpha - 1.) / 2. - 0.5)
        if new.shape == self.rowscale.shape and new.device == self.rowscale.device:
            self.rowscale.copy_(new)
        else:
            self.rowscale = new.contiguous()

    # ---- reference-order bookkeeping (compat views and replay of recorded dropout draws)
    def coo_order(self):
        """Positions of my entries inside the reference's coalesced `feat_mat`.

        Returns (edge_pos[nnz_adj] int64, self_pos[N] int64, idx[2, nnz_feat]): edge_pos[e] is the
        index, in the reference's (row, column)-sorted nnz order, of adjacency edge e (or -1 when
        its column node is not a template); self_pos[r] is the index of row r's global entry."""
        if self._order is None:
            rp, col, tm = self.rowptr_full, self.col_full, self.tmpl_host
            n = len(rp) - 1
            rows = np.repeat(np.arange(n, dtype=np.int64), np.diff(rp))
            fcol = tm[col].astype(np.int64)
            keep = fcol >= 0
            e_rows, e_cols = rows[keep], fcol[keep]
            s_rows = np.arange(n, dtype=np.int64)
            s_cols = np.where(s_rows < self.n_users, self.glob_user, self.glob_item).astype(np.int64)
            all_rows = np.concatenate([e_rows, s_rows])
            all_cols = np.concatenate([e_cols, s_cols])
            order = np.lexsort((all_cols, all_rows))
            pos = np.empty(len(order), dtype=np.int64)
            pos[order] = np.arange(len(order), dtype=np.int64)
            edge_pos = np.full(len(col), -1, dtype=np.int64)
            edge_pos[keep] = pos[:len(e_rows)]
            self_pos = pos[len(e_rows):]
            idx = np.stack([all_rows[order], all_cols[order]])
            self._order = (edge_pos, self_pos, idx)
        return self._order

    def _coo(self):
        edge_pos, self_pos, idx = self.coo_order()
        idx_t = torch.from_numpy(idx).to(self.device)
        return idx_t, self.rowscale[idx_t[0]]

    def keep_bits(self, keep_ref):
        """Pack a keep vector given in the reference's nnz order into (edge_keep, self_keep) bit
        arrays in CSR order -- the `mode 2` dropout of include/igcn_b200.h."""
        keep_ref = np.asarray(keep_ref).astype(bool)
        edge_pos, self_pos, _ = self.coo_order()
        ek = np.zeros(len(edge_pos), dtype=bool)
        ok = edge_pos >= 0
        ek[ok] = keep_ref[edge_pos[ok]]
        sk = keep_ref[self_pos]
        return _pack_bits(ek, self.device), _pack_bits(sk, self.device)

    def tperm(self):
        """tperm[e] = CSR position of the reverse edge (for the mode-2 backward pass)."""
        if self._tperm is None:
            rp, col = self.rowptr_full, self.col_full.astype(np.int64)
            n = len(rp) - 1
            rows = np.repeat(np.arange(n, dtype=np.int64), np.diff(rp))
            fwd = rows * n + col        # ascending (CSR order, sorted columns)
            self._tperm = torch.from_numpy(np.searchsorted(fwd, col * n + rows).astype(np.int64)).to(self.device)
        return self._tperm


def _pack_bits(flags, device):
    n = len(flags)
    padded = np.zeros((n + 31) // 32 * 32 + 32, dtype=np.uint8)
    padded[:n] = flags
    words = np.packbits(padded.reshape(-1, 32)[:, ::-1], axis=1).view('>u4').astype(np.uint32).ravel()
    return torch.from_numpy(words.view(np.int32).copy()).to(device)
