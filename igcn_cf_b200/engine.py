"""Host-side drivers of the sm_100a kernels: propagation forward/backward, the fused BPR training
step and full-ranking evaluation.  Every arithmetic step is a call into libigcn_b200.so through
igcn_cf_b200._lib; PyTorch only owns the buffers, the stream and (optionally) the CUDA graph.

Reference call sites are cited on each method.
"""
import ctypes as C
import os

import numpy as np
import torch

from . import _lib
from ._lib import call, ptr, stream_ptr

BETA1, BETA2, ADAM_EPS = 0.9, 0.999, 1e-8      # torch.optim.Adam defaults (trainer.py:43-45)
MAX_PLAN = 16384                                # igcn_bpr_plan: 3 * B <= 16384


def _drop_struct(drop):
    """drop: None | dict(mode, p, seed, seed_dev, edge_keep, self_keep, tperm) -> C struct pointer."""
    if drop is None:
        return None
    s = _lib.DropoutStruct(int(drop.get('mode', 0)), float(drop.get('p', 0.)), int(drop.get('seed', 0)),
                           ptr(drop.get('seed_dev')), ptr(drop.get('edge_keep')), ptr(drop.get('self_keep')),
                           ptr(drop.get('tperm')))
    return C.byref(s)


class Shard:
    """Row-sharded propagation: the peer context plus the registry of symmetric (peer-mapped) buffers the
    kernels may store into.  Which rows this rank computes is a property of the graph objects
    (`adj.blocks`).  `None` stands for the single-GPU case."""

    # row blocks of at least this many bytes leave through igcn_peer_push (bulk, after the kernel) instead of
    # the kernel epilogue's peer stores: measured on the 2.8 GB layers of the scale-out graph the scattered
    # in-kernel stores reach ~100 GB/s, the bulk push the link rate
    PUSH_BYTES = int(os.environ.get('IGCN_PEER_PUSH_BYTES', 32 << 20))

    def __init__(self, ctx):
        self.ctx = ctx
        self._bufs = {}
        self._side = None
        self._pending = False

    def push_behind(self, t, row0, n_rows, dim):
        """Bulk push of a finished row block on a side stream: it leaves over NVLink while the main stream already
        computes the rank's next row block (user slice out while the item slice is computed).  join() before the
        barrier."""
        if self._side is None:
            self._side = torch.cuda.Stream()
        main = torch.cuda.current_stream()
        self._side.wait_stream(main)
        with torch.cuda.stream(self._side):
            self.push(t, row0, n_rows, dim)
        self._pending = True

    def join(self):
        # only when a push was forked since the last join (inside a CUDA-graph capture the side stream must have been
        # forked from the capturing stream)
        if self._pending:
            torch.cuda.current_stream().wait_stream(self._side)
            self._pending = False

    def bulk(self, n_rows, dim):
        return n_rows * dim * 4 >= self.PUSH_BYTES

    def push(self, t, row0, n_rows, dim):
        """Stream rows [row0, +n_rows) of symmetric buffer `t` to every other rank."""
        arr, world = self.peers(t, 0)
        call('igcn_peer_push', arr, world, self.ctx.rank, row0 * dim, n_rows * dim, stream_ptr())

    def new_buffer(self, rows, dim):
        buf = self.ctx.alloc((rows, dim), torch.float32)
        self._bufs[buf.tensor.data_ptr()] = buf
        return buf.tensor

    def peers(self, t, byte_offset):
        buf = self._bufs.get(t.data_ptr())
        if buf is None:
            raise RuntimeError('row-sharded propagation writes into symmetric buffers only '
                               '(allocate the output with Propagator.new_buffer)')
        return buf.peer_array(byte_offset), self.ctx.world


def _peer_args(shard, t, byte_offset):
    return (None, 0) if shard is None else shard.peers(t, byte_offset)


class Propagator:
    """L-layer propagation with the layer mean fused into the last SpMM, and its backward.

    forward:  X_{l+1} = A X_l, rep = mean(X_0..X_L)           (model.py:96-106, 434-446)
    backward: h_L = g', h_l = A h_{l+1} + g', dX_0 = h_0 with g' = d_rep / (L+1)  (A symmetric)
    Intermediate layers live in persistent buffers that the backward pass reuses as ping-pong
    space (the propagation is linear, no activation has to be saved).

    With a `shard` every rank computes its own row block and the kernel epilogue stores each row into
    all ranks' buffers (fused all-gather over NVLink); a device-side barrier follows every layer."""

    def __init__(self, n_nodes, dim, n_layers, device, shard=None):
        if n_layers > _lib.MAX_ADD:
            raise RuntimeError('n_layers > %d is not supported by the fused layer-mean epilogue' % _lib.MAX_ADD)
        self.n, self.dim, self.n_layers = int(n_nodes), int(dim), int(n_layers)
        self.device = torch.device(device)
        self.shard = shard
        self.layers = [self.new_buffer(self.n) for _ in range(max(2, self.n_layers - 1))]   # X_1.. / backward ping-pong
        self.x0 = None          # allocated on demand by the INMO layer
        self.rep = self.new_buffer(self.n)
        self._out = None
        self._add_arrays = {}

    def new_buffer(self, rows):
        """[rows, dim] fp32 buffer a propagation kernel may write: symmetric when row-sharded."""
        if self.shard is not None:
            return self.shard.new_buffer(rows, self.dim)
        return torch.empty((rows, self.dim), dtype=torch.float32, device=self.device)

    def x0_buffer(self):
        if self.x0 is None:
            self.x0 = self.new_buffer(self.n)
        return self.x0

    def out_buffer(self):
        """Scratch output for the autograd bridges (the fused TrainStep owns its own)."""
        if self._out is None:
            self._out = self.new_buffer(self.n)
        return self._out

    def _adds(self, tensors, off):
        key = tuple(t.data_ptr() + off for t in tensors)
        arr = self._add_arrays.get(key)
        if arr is None:
            arr = (C.c_void_p * max(1, len(key)))(*key)
            self._add_arrays[key] = arr
        return arr

    def _partials(self, n_rows):
        """Two [n_rows, dim] ping-pong buffers for the partial results of a column-blocked layer."""
        key = (n_rows, self.dim)
        if getattr(self, '_partial_key', None) != key:
            self._partial_key = key
            self._partial_bufs = [torch.empty((n_rows, self.dim), dtype=torch.float32, device=self.device) for _ in range(2)]
        return self._partial_bufs

    def spmm(self, adj, x, y, adds=(), rowscale=None, alpha=1.0, rows=None, cols=None):
        """One layer.  rows = (row_list int64, n_list int32[1], max_list): compute the listed rows only
        (igcn_spmm_rows); cols = bitmap of the columns whose X row is non-zero (igcn_spmm_cols).
        Row blocks that carry column blocks (graph.column_blocks: tables far larger than L2) run once per column
        range, every pass adding the previous partial result; the last pass applies adds / rowscale / alpha and the
        exchange."""
        sh = self.shard
        for blk in adj.blocks:
            row0 = blk.row0
            off = row0 * self.dim * 4
            bulk = sh is not None and sh.bulk(blk.csr.n_rows, self.dim) and rows is None
            if bulk:
                sh.peers(y, off)                                   # must be a symmetric buffer all the same
                peers, n_peers = None, 0
            else:
                peers, n_peers = _peer_args(sh, y, off)
            rs = None if rowscale is None else ptr(rowscale) + row0 * 4
            cbs = blk.col_blocks if (rows is None and cols is None) else None
            if cbs and len(cbs) > 1:
                bufs = self._partials(blk.csr.n_rows)
                prev = None
                for b, csr in enumerate(cbs[:-1]):
                    dst = bufs[b & 1]
                    padd = (C.c_void_p * 1)(*([] if prev is None else [prev.data_ptr()]))
                    call('igcn_spmm', csr.struct(self.dim), ptr(x), ptr(dst), self.dim, padd, 0 if prev is None else 1, None, 1.0,
                         None, 0, stream_ptr())
                    prev = dst
                last = (C.c_void_p * (len(adds) + 1))(prev.data_ptr(), *[t.data_ptr() + off for t in adds])
                call('igcn_spmm', cbs[-1].struct(self.dim), ptr(x), ptr(y) + off, self.dim, last, len(adds) + 1, rs, float(alpha),
                     peers, n_peers, stream_ptr())
                if bulk:
                    sh.push_behind(y, row0, blk.csr.n_rows, self.dim)
                continue
            head = (blk.csr.struct(self.dim), ptr(x), ptr(y) + off, self.dim, self._adds(adds, off), len(adds), rs, float(alpha))
            if rows is not None:
                call('igcn_spmm_rows', *head, ptr(rows[0]), ptr(rows[1]), int(rows[2]), row0, peers, n_peers, stream_ptr())
            elif cols is not None:
                call('igcn_spmm_cols', *head, ptr(cols), peers, n_peers, stream_ptr())
            else:
                call('igcn_spmm', *head, peers, n_peers, stream_ptr())
            if bulk:
                sh.push_behind(y, row0, blk.csr.n_rows, self.dim)      # overlaps the next row block's kernels
        if sh is not None:
            sh.join()
            sh.ctx.barrier()

    def forward(self, adj, x0, out=None, rows=None, before_last=None):
        """rep = mean_l A^l x0.  `out` defaults to the persistent rep buffer.  With `rows` the last layer
        (and with it the layer mean) is only evaluated on the listed rows -- a training step reads
        nothing else; `before_last` runs right before that layer is enqueued (stream join)."""
        L = self.n_layers
        rep = self.rep if out is None else out
        if L == 0:
            rep.copy_(x0)
            return rep
        xs = [x0]
        for l in range(1, L):
            self.spmm(adj, xs[-1], self.layers[l - 1])
            xs.append(self.layers[l - 1])
        if before_last is not None:
            before_last()
        self.spmm(adj, xs[-1], rep, adds=xs, alpha=1.0 / (L + 1), rows=rows)
        return rep

    def backward(self, adj, gprime, out, rowscale=None, alpha=1.0, gprime_rows=None):
        """out = alpha * rowscale .* h_0, h from the Horner recurrence above; gprime = d_rep/(L+1).
        gprime_rows: bitmap of the rows where gprime is non-zero (first layer skips the other columns)."""
        L = self.n_layers
        if L == 0:
            if rowscale is None:
                out.copy_(gprime)
                if alpha != 1.0:
                    out.mul_(alpha)
            else:
                torch.mul(gprime, rowscale[:, None] * alpha, out=out)
            return out
        h = gprime
        cols = gprime_rows
        for l in range(L - 1, 0, -1):
            dst = self.layers[l % 2]
            self.spmm(adj, h, dst, adds=(gprime,), cols=cols)
            h, cols = dst, None
        self.spmm(adj, h, out, adds=(gprime,), rowscale=rowscale, alpha=alpha, cols=cols)
        return out


def inmo_forward(feat, emb, x0, drop, dim, shard=None):
    """X0 = F~ E with dropout fused (model.py:423-432 after model.py:435)."""
    for blk in feat.blocks:
        row0 = blk.row0
        off = row0 * dim * 4
        bulk = shard is not None and shard.bulk(blk.csr.n_rows, dim)
        if bulk:
            shard.peers(x0, off)
            peers, n_peers = None, 0
        else:
            peers, n_peers = _peer_args(shard, x0, off)
        call('igcn_inmo_fwd', blk.csr.struct(dim), ptr(feat.tmpl), ptr(feat.rowscale) + row0 * 4, _drop_struct(drop),
             ptr(emb), ptr(x0) + off, dim, row0, feat.n_users, feat.glob_user, feat.glob_item, peers, n_peers, stream_ptr())
        if bulk:
            shard.push_behind(x0, row0, blk.csr.n_rows, dim)          # leaves while the next row block is computed
    if shard is not None:
        shard.join()
        shard.ctx.barrier()


def inmo_backward(feat, g_scaled, d_emb, drop, dim, scratch, shard=None, side=None):
    """dE = F~^T dX0 given g_scaled = rowscale/(1-p) .* dX0 (autograd backward of model.py:430).
    Row-sharded: each rank produces the template rows of its own node block and stores them into every
    rank's d_emb; the two global-template rows are column sums every rank computes for itself.
    side: a stream for the two column sums (small, latency-bound, independent of the transposed product: they
    write the two global-template rows only); the caller joins it before reading d_emb."""
    n, u = feat.shape[0], feat.n_users
    peers, n_peers = _peer_args(shard, d_emb, 0)
    d = _drop_struct(drop)

    def colsums():
        call('igcn_colsum_masked', ptr(g_scaled), 0, u, dim, d, ptr(scratch), ptr(d_emb[feat.glob_user]), stream_ptr())
        call('igcn_colsum_masked', ptr(g_scaled), u, n, dim, d, ptr(scratch), ptr(d_emb[feat.glob_item]), stream_ptr())

    if side is not None:
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            colsums()
    for blk in feat.blocks:
        call('igcn_inmo_bwd', blk.csr.struct(dim), ptr(feat.tmpl), _drop_struct(drop), ptr(g_scaled), ptr(d_emb), dim,
             blk.row0, peers, n_peers, stream_ptr())
    if side is None:
        colsums()
    if shard is not None:
        shard.ctx.barrier()


def colsum_scratch(n_rows, dim, device):
    return torch.empty(((n_rows + 255) // 256 + 1, dim), dtype=torch.float32, device=device)


# --------------------------------------------------------------------------- autograd bridges
class LightGCNRep(torch.autograd.Function):
    """get_rep for LightGCN as one autograd node (generic API path; model.py:96-106)."""

    @staticmethod
    def forward(ctx, emb, model):
        prop = model._propagator()
        rep = prop.forward(model.norm_adj, emb.detach().contiguous()).clone()
        ctx.model = model
        return rep

    @staticmethod
    def backward(ctx, g):
        model = ctx.model
        prop = model._propagator()
        gprime = (g * (1.0 / (prop.n_layers + 1))).contiguous()
        d_emb = prop.backward(model.norm_adj, gprime, prop.out_buffer())
        return d_emb.clone(), None


class IGCNRep(torch.autograd.Function):
    """get_rep for IGCN as one autograd node (model.py:434-446); drop is the mask description."""

    @staticmethod
    def forward(ctx, emb, model, drop):
        prop = model._propagator()
        feat = model.feat_mat
        x0 = prop.x0_buffer()
        inmo_forward(feat, emb.detach().contiguous(), x0, drop, prop.dim, prop.shard)
        rep = prop.forward(model.norm_adj, x0).clone()
        ctx.model, ctx.drop, ctx.emb_shape = model, drop, emb.shape
        return rep

    @staticmethod
    def backward(ctx, g):
        model, drop = ctx.model, ctx.drop
        prop = model._propagator()
        feat = model.feat_mat
        gprime = (g * (1.0 / (prop.n_layers + 1))).contiguous()
        g_scaled = prop.x0_buffer()
        inv_keep = 1.0 if (drop is None or drop.get('mode', 0) == 0) else 1.0 / (1.0 - drop['p'])
        prop.backward(model.norm_adj, gprime, g_scaled, rowscale=feat.rowscale, alpha=inv_keep)
        if prop.shard is None:
            d_emb = torch.zeros(ctx.emb_shape, dtype=torch.float32, device=g.device)
        else:
            d_emb = model._grad_buffer(ctx.emb_shape[0])
            d_emb.zero_()
            prop.shard.ctx.barrier()          # nobody stores template rows before every copy is zeroed
        if drop is not None and drop.get('mode', 0) == 2 and drop.get('tperm') is None:
            drop = dict(drop, tperm=feat.tperm())
        inmo_backward(feat, g_scaled, d_emb, drop, prop.dim, colsum_scratch(prop.n, prop.dim, g.device), prop.shard)
        return (d_emb if prop.shard is None else d_emb.clone()), None, None


class ColumnRep:
    """Evaluation-mode get_rep of a column-sharded model (model_config['shard'] = 'dims' / 'auto'): every rank runs the
    template layer and the L propagation layers on ITS embedding_size / world columns only -- the propagation is linear
    and acts on each column independently, so nothing is exchanged between layers -- then stores its slice of the
    layer mean into the column range of every rank's full-width copy over NVLink (igcn_peer_push_cols) and one
    device-side barrier closes the all-gather.  Same kernels, same per-row summation order as the full-width
    propagation: bit-identical to one GPU (bench.py `shard_parity`, tests/dist_worker.py)."""

    def __init__(self, model):
        rank, world = model._dim_shard
        self.peers = model._peers
        self.world, self.D_full = world, model.embedding_size
        self.D = self.D_full // world
        self.col0 = rank * self.D
        self.n = model.n_users + model.n_items
        self.rows = model.embedding.weight.shape[0]
        dev = model.embedding.weight.device
        self.prop = Propagator(self.n, self.D, model.n_layers, dev, None)
        self.emb_s = torch.empty((self.rows, self.D), dtype=torch.float32, device=dev)
        self.full = self.peers.alloc((self.n, self.D_full), torch.float32)
        self.key = self.key_of(model)

    @staticmethod
    def key_of(model):
        return (model.n_users + model.n_items, model.embedding.weight.shape[0], model.embedding_size, model.n_layers)

    def run(self, model):
        """Collective: every rank calls it for the same parameters and graph."""
        self.emb_s.copy_(model.embedding.weight.data[:, self.col0:self.col0 + self.D])
        feat = getattr(model, 'feat_mat', None)
        if feat is None:
            x0 = self.emb_s
        else:
            x0 = self.prop.x0_buffer()
            inmo_forward(feat, self.emb_s, x0, None, self.D)
        rep = self.prop.forward(model.norm_adj, x0)
        self.peers.barrier()                             # every rank is done reading the copy of the previous evaluation
        call('igcn_peer_push_cols', self.full.peer_array(0), self.world, ptr(rep), self.n, self.D, self.D_full, self.col0,
             stream_ptr())
        self.peers.barrier()
        return self.full.tensor.clone()                  # callers keep the representation; the symmetric copy is reused


# --------------------------------------------------------------------------- fused training step
class TrainStep:
    """One BPR training step with no autograd tape (trainer.py:233-247 and 296-318).

    sample (or take injected) triples -> plan -> propagate -> fused BPR forward -> loss on device
    -> deterministic gradient rows -> propagate backward -> (INMO transpose) -> Adam.
    With `use_graph=True` the whole step is captured once into a CUDA graph and replayed; step
    counter, sampler stream, dropout seed and Adam bias corrections advance on the device."""

    def __init__(self, model, opt, l2_reg, aux_reg=None, batch_size=2048, seed=0, use_graph=False):
        self.model = model
        self.opt = opt
        self.is_igcn = hasattr(model, 'feat_mat')
        self.lr, self.l2_reg, self.aux_reg = float(opt.param_groups[0]['lr']), float(l2_reg), aux_reg
        self.B = int(batch_size)
        if 3 * self.B > MAX_PLAN:
            raise RuntimeError('batch_size %d too large for the single-CTA scatter plan (3*B <= %d)' % (self.B, MAX_PLAN))
        self.seed = int(seed)
        self.use_graph = bool(use_graph)
        dev = model.embedding.weight.device
        self.device = dev
        # column-sharded training (model_config['shard'] = 'dims'): this rank owns D / world columns of the parameters
        # and of every layer buffer; the step runs the same kernels on the narrow tables
        self.dims = getattr(model, '_dim_shard', None)
        D = model.embedding_size
        if self.dims is not None:
            rank, world = self.dims
            if world not in (2, 4, 8) or D % (4 * world):
                raise RuntimeError('column sharding needs 2, 4 or 8 ranks and embedding_size %% (4 * ranks) == 0')
            self.D_full, D = D, D // world
            self.col0 = rank * D
        self.D = D
        i32 = lambda *s: torch.empty(s, dtype=torch.int32, device=dev)
        f32 = lambda *s: torch.empty(s, dtype=torch.float32, device=dev)
        self.triples = torch.zeros((self.B, 3), dtype=torch.int64, device=dev)
        self.order, self.seg_start = i32(3 * self.B), i32(3 * self.B + 1)
        self.seg_row = torch.empty(3 * self.B, dtype=torch.int64, device=dev)
        self.n_seg = torch.zeros(1, dtype=torch.int32, device=dev)
        n_nodes = model.n_users + model.n_items
        self.touched = torch.zeros((n_nodes + 31) // 32 + 1, dtype=torch.int32, device=dev)   # bitmap of the batch's rows
        self.sp, self.sig, self.l2 = f32(self.B), f32(self.B), f32(self.B)
        self.loss = torch.zeros(1, dtype=torch.float32, device=dev)
        self.acc = torch.zeros(2, dtype=torch.float64, device=dev)
        self.state = opt.device_state(dev)                                # igcn_step_state
        n = model.n_users + model.n_items
        self.gprime = torch.zeros((n, D), dtype=torch.float32, device=dev)
        if self.dims is None:
            prop = model._propagator()
            self.prop = None                                                  # looked up per step: the model may rebuild it
            self.shard = prop.shard
            self.d_emb = prop.new_buffer(model.embedding.weight.shape[0])     # symmetric when row-sharded
            self.d_emb.zero_()
            self.emb_m, self.emb_v = opt.moments(model.embedding.weight)
        else:
            self._init_dims(model, n, D, dev)
        if self.is_igcn:
            self.a_triples = torch.zeros((self.B, 3), dtype=torch.int64, device=dev)
            self.a_order, self.a_seg_start = i32(3 * self.B), i32(3 * self.B + 1)
            self.a_seg_row = torch.empty(3 * self.B, dtype=torch.int64, device=dev)
            self.a_n_seg = torch.zeros(1, dtype=torch.int32, device=dev)
            self.a_sp, self.a_sig = f32(self.B), f32(self.B)
            self.d_w = torch.zeros(D, dtype=torch.float32, device=dev)
            self.dw_scratch = f32((self.B + 63) // 64, D)
            if self.dims is None:
                self.w_m, self.w_v = opt.moments(model.w)
            else:
                self.w_s = torch.empty(D, dtype=torch.float32, device=dev)
                self.w_m, self.w_v = torch.zeros_like(self.w_s), torch.zeros_like(self.w_s)
            self.colsum_scratch = colsum_scratch(n, D, dev)
        self._graphs = {}
        self._side = torch.cuda.Stream(device=dev)       # main scatter plan: joins before the last forward layer
        self._side2 = torch.cuda.Stream(device=dev)      # zeroed gradient buffers + auxiliary plan: joins before the gradient kernels
        self._side3 = torch.cuda.Stream(device=dev)      # loss value + running meter: nothing downstream reads them; joins at the end
        self._side4 = torch.cuda.Stream(device=dev)      # gradient of the auxiliary weight vector: needs sig only; joins before Adam
        self._side5 = torch.cuda.Stream(device=dev)      # column sums of the global-template gradient rows, beside the transposed INMO product

    # -- column-sharded training
    def _init_dims(self, model, n, D, dev):
        """Buffers of the column-sharded step: narrow parameter / moment / gradient tables, a private propagator of
        width D / world (no exchange: the propagation acts on every column independently), the exchange buffer of
        the per-triple partial sums and a symmetric staging copy for the parameter all-gather."""
        rank, world = self.dims
        peers = model._peers
        rows = model.embedding.weight.shape[0]
        self.prop = Propagator(n, D, model.n_layers, dev, None)
        self.shard = None
        z = lambda *sh: torch.zeros(sh, dtype=torch.float32, device=dev)
        self.emb_s, self.emb_m, self.emb_v, self.d_emb = z(rows, D), z(rows, D), z(rows, D), z(rows, D)
        self.parts = peers.alloc((2 * world * self.B * 8,), torch.float32)
        self.gather = peers.alloc((rows + 1, self.D_full), torch.float32)     # last row: w
        self._seen_epoch = None
        self._dirty = False
        model._param_sync = self.sync_params

    def _reslice(self):
        """Take this rank's columns of the model's (full-width) parameters; Adam moments restart only when the
        parameters were replaced from outside (load / assignment), not after our own steps."""
        m = self.model
        c0, c1 = self.col0, self.col0 + self.D
        self.emb_s.copy_(m.embedding.weight.data[:, c0:c1])
        if self.is_igcn:
            self.w_s.copy_(m.w.data[c0:c1])

    def sync_params(self):
        """All-gather of the column slices into model.embedding.weight (and w) on every rank: each rank writes its
        columns into every rank's staging copy over NVLink (igcn_peer_push_cols), barrier, local copy.  Called by the
        model before it needs full-width parameters (get_rep / save) and at the end of an epoch; collective."""
        if self.dims is None or not self._dirty:
            return
        m, peers = self.model, self.model._peers
        rank, world = self.dims
        rows = m.embedding.weight.shape[0]
        peers.barrier()                                  # nobody still reads the staging copy of the previous gather
        arr = self.gather.peer_array(0)
        call('igcn_peer_push_cols', arr, world, ptr(self.emb_s), rows, self.D, self.D_full, self.col0, stream_ptr())
        if self.is_igcn:
            arr_w = self.gather.peer_array(rows * self.D_full * 4)
            call('igcn_peer_push_cols', arr_w, world, ptr(self.w_s), 1, self.D, self.D_full, self.col0, stream_ptr())
        peers.barrier()
        full = self.gather.tensor
        m.embedding.weight.data.copy_(full[:rows])
        if self.is_igcn:
            m.w.data.copy_(full[rows])
        self._dirty = False
        m._bump()
        self._seen_epoch = m._param_epoch

    # -- pieces
    def _sample_main(self, B):
        m = self.model
        rowptr, col = m.norm_adj.sampler_csr()
        call('igcn_sample_triples', ptr(rowptr), ptr(col), m.n_users, m.n_users, m.n_items, B, self.seed,
             0, ptr(self.state), ptr(self.triples), stream_ptr())

    def _sample_aux(self, B):
        a = self.model.aux_csr()
        call('igcn_sample_triples', ptr(a['rowptr']), ptr(a['col']), a['col_offset'], a['n_users'], a['n_items'], B,
             self.seed ^ 0x5bd1e995, 0, ptr(self.state), ptr(self.a_triples), stream_ptr())

    def _body(self, B, sample, drop):
        m, D, st = self.model, self.D, stream_ptr
        dims = self.dims
        prop = m._propagator() if dims is None else self.prop
        L = prop.n_layers
        emb = m.embedding.weight.data if dims is None else self.emb_s
        w = (m.w.data if dims is None else self.w_s) if self.is_igcn else None
        call('igcn_step_tick', ptr(self.state), self.lr, BETA1, BETA2, st())
        # the triples and their scatter plans are needed by the LAST forward layer at the earliest: sample and sort
        # them on side streams while the forward propagation runs (one or a few CTAs each)
        main = torch.cuda.current_stream()
        self._side.wait_stream(main)
        with torch.cuda.stream(self._side):
            if sample:
                self._sample_main(B)
            call('igcn_bpr_plan', ptr(self.triples), B, m.n_users, m.n_users + m.n_items, ptr(self.order),
                 ptr(self.seg_start), ptr(self.seg_row), ptr(self.n_seg), ptr(self.touched), st())
        # second side stream: everything else the gradient kernels need that does not depend on the forward pass --
        # the zeroed gradient buffers and (IGCN) the auxiliary scatter plan; joins before the gradient kernels
        self._side2.wait_stream(main)
        with torch.cuda.stream(self._side2):
            self.gprime.zero_()
            if self.is_igcn:
                self.d_w.zero_()
                if m.feat_mat.tmpl is not None:
                    # template rows without a node in this graph get no gradient; zeroed long before any rank
                    # stores template rows into this copy (several barriers later)
                    self.d_emb.zero_()
                if sample:
                    self._sample_aux(B)
                call('igcn_bpr_plan', ptr(self.a_triples), B, m.feat_mat.t_users, m.embedding.weight.shape[0],
                     ptr(self.a_order), ptr(self.a_seg_start), ptr(self.a_seg_row), ptr(self.a_n_seg), None, st())
        # forward: full layers 1..L-1, then the last layer + layer mean on the batch's rows only (the plan's
        # sorted unique row list)
        rows = (self.seg_row, self.n_seg, 3 * B)
        join = lambda: main.wait_stream(self._side)
        if self.is_igcn:
            x0 = prop.x0_buffer()
            inmo_forward(m.feat_mat, emb, x0, drop, D, self.shard)
            rep = prop.forward(m.norm_adj, x0, rows=rows, before_last=join)
            l2_table = rep
        else:
            rep = prop.forward(m.norm_adj, emb, rows=rows, before_last=join)
            l2_table = emb
        if L == 0:
            join()
        main.wait_stream(self._side2)            # auxiliary triples, zeroed gradient buffers
        if dims is None:
            call('igcn_bpr_fwd', ptr(rep), ptr(l2_table), None, ptr(self.triples), B, m.n_users, D, ptr(self.sp),
                 ptr(self.sig), ptr(self.l2), st())
            if self.is_igcn:
                call('igcn_bpr_fwd', ptr(emb), None, ptr(w), ptr(self.a_triples), B, m.feat_mat.t_users, D, ptr(self.a_sp),
                     ptr(self.a_sig), None, st())
        else:
            # the only cross-column quantities of the step: every rank's partial dot products / squared norms go to
            # every rank (20 bytes per triple and peer), one device barrier, then the tree-ordered combine
            rank, world = dims
            arr = self.parts.peer_array(0)
            call('igcn_bpr_partial', ptr(rep), ptr(l2_table), None, ptr(self.triples), B, m.n_users, D, arr, world, rank, 0,
                 self.B, ptr(self.state), st())
            if self.is_igcn:
                call('igcn_bpr_partial', ptr(emb), None, ptr(w), ptr(self.a_triples), B, m.feat_mat.t_users, D, arr, world,
                     rank, 5, self.B, ptr(self.state), st())
            m._peers.barrier()
            parts = self.parts.tensor
            call('igcn_bpr_combine', ptr(parts), B, self.B, world, 0, 1, ptr(self.state), ptr(self.sp), ptr(self.sig),
                 ptr(self.l2), st())
            if self.is_igcn:
                call('igcn_bpr_combine', ptr(parts), B, self.B, world, 5, 0, ptr(self.state), ptr(self.a_sp), ptr(self.a_sig),
                     None, st())
        self._side3.wait_stream(main)
        with torch.cuda.stream(self._side3):
            if self.is_igcn:
                call('igcn_loss_finalize', ptr(self.sp), ptr(self.l2), ptr(self.a_sp), B, B, self.l2_reg, self.aux_reg,
                     ptr(self.loss), ptr(self.acc), st())
            else:
                call('igcn_loss_finalize', ptr(self.sp), ptr(self.l2), None, B, 0, self.l2_reg, 0.0, ptr(self.loss),
                     ptr(self.acc), st())
        if self.is_igcn:
            # d_w depends on sig and the parameters only: beside the whole backward propagation
            self._side4.wait_stream(main)
            with torch.cuda.stream(self._side4):
                call('igcn_bpr_dw', ptr(emb), ptr(self.a_triples), B, m.feat_mat.t_users, D, ptr(self.a_sig), float(self.aux_reg),
                     ptr(self.d_w), ptr(self.dw_scratch), st())
        # backward
        main.wait_stream(self._side)
        call('igcn_bpr_bwd', ptr(rep), None, ptr(self.triples), B, m.n_users, D, ptr(self.sig), 1.0 / (L + 1),
             self.l2_reg if self.is_igcn else 0.0, 1 if self.is_igcn else 0, ptr(self.order), ptr(self.seg_start),
             ptr(self.seg_row), ptr(self.n_seg), ptr(self.gprime), 0, None, None, st())
        if self.is_igcn:
            feat = m.feat_mat
            g_scaled = prop.x0_buffer()
            inv_keep = 1.0 if (drop is None or drop.get('mode', 0) == 0) else 1.0 / (1.0 - drop['p'])
            prop.backward(m.norm_adj, self.gprime, g_scaled, rowscale=feat.rowscale, alpha=inv_keep,
                          gprime_rows=self.touched)
            inmo_backward(feat, g_scaled, self.d_emb, drop, D, self.colsum_scratch, self.shard, side=self._side5)
            call('igcn_bpr_bwd', ptr(emb), ptr(w), ptr(self.a_triples), B, feat.t_users, D, ptr(self.a_sig),
                 float(self.aux_reg), 0.0, 0, ptr(self.a_order), ptr(self.a_seg_start), ptr(self.a_seg_row),
                 ptr(self.a_n_seg), ptr(self.d_emb), 1, None, None, st())
            main.wait_stream(self._side5)
            main.wait_stream(self._side4)
        else:
            prop.backward(m.norm_adj, self.gprime, self.d_emb, gprime_rows=self.touched)
            if self.l2_reg != 0.0:
                call('igcn_l2_rows_bwd', ptr(emb), ptr(self.d_emb), D, 2.0 * self.l2_reg / B, ptr(self.seg_start),
                     ptr(self.seg_row), ptr(self.n_seg), 3 * B, st())
        # Adam
        call('igcn_adam', ptr(emb), ptr(self.d_emb), ptr(self.emb_m), ptr(self.emb_v), emb.numel(), self.lr,
             BETA1, BETA2, ADAM_EPS, 0, ptr(self.state), st())
        if self.is_igcn:
            call('igcn_adam', ptr(w), ptr(self.d_w), ptr(self.w_m), ptr(self.w_v), w.numel(), self.lr, BETA1,
                 BETA2, ADAM_EPS, 0, ptr(self.state), st())
        main.wait_stream(self._side3)

    def _production_drop(self):
        m = self.model
        if not self.is_igcn or not m.training or m.dropout <= 0.0:
            return None
        return {'mode': 1, 'p': m.dropout, 'seed': self.seed * 0x9e3779b97f4a7c15 % (1 << 64), 'seed_dev': self.state}

    # -- public
    def run(self, triples=None, aux_triples=None, drop='auto', batch=None):
        """One step.  triples/aux_triples: injected int64 [b,3] device tensors (parity mode) or None
        to sample `batch` (default batch_size) triples on the device; drop: 'auto' (hash dropout
        when the model is in training mode) or an explicit dict (see _drop_struct)."""
        B = self.B if batch is None else int(batch)
        if B > self.B:
            raise RuntimeError('batch %d exceeds the step\'s capacity %d' % (B, self.B))
        sample = triples is None
        if not sample:
            B = int(triples.shape[0])
            self.triples[:B].copy_(triples)
            if self.is_igcn:
                self.a_triples[:B].copy_(aux_triples)
        if self.dims is not None and self._seen_epoch != self.model._param_epoch:
            self.sync_params()                 # (no-op unless our own steps are pending)
            self._reslice()                    # the full-width parameters were set from outside: take our columns again
        # the learning rate is a launch argument of igcn_step_tick / igcn_adam: follow the optimizer's param group
        # (schedulers, manual decay) like Adam.step() does; a new value re-captures the CUDA graph
        self.lr = float(self.opt.param_groups[0]['lr'])
        d = self._production_drop() if isinstance(drop, str) else drop
        if d is not None and d.get('mode', 0) == 2 and d.get('tperm') is None:
            d = dict(d, tperm=self.model.feat_mat.tperm())
        graphable = self.use_graph and (d is None or d.get('mode', 0) != 2)
        if not graphable:
            self._body(B, sample, d)
        else:
            gv = self.model.graph_version()
            key = (B, sample, None if d is None else (d['mode'], d['p'], d['seed']), gv, self.lr)
            g = self._graphs.get(key)
            if g is None:
                # captured graphs of a replaced norm_adj / feat_mat hold pointers into buffers that may be freed
                self._graphs = {k: v for k, v in self._graphs.items() if k[3] == gv}
                if len(self._graphs) >= 8:
                    self._graphs.clear()
                g = self._capture(B, sample, d)
                self._graphs[key] = g
            g.replay()
        self.opt.t += 1
        self.model._bump()
        if self.dims is not None:
            self._dirty = True
            self._seen_epoch = self.model._param_epoch
        return self.loss

    def _capture(self, B, sample, d):
        torch.cuda.synchronize()
        state_backup = self.state.clone()
        snap = self._snapshot()
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            self._body(B, sample, d)       # warm-up outside capture (lazy allocations, func attributes)
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        self._restore(snap, state_backup)
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            self._body(B, sample, d)
        return g

    def _snapshot(self):
        dims = self.dims is not None
        ts = [self.emb_s if dims else self.model.embedding.weight.data, self.emb_m, self.emb_v, self.acc]
        if self.is_igcn:
            ts += [self.w_s if dims else self.model.w.data, self.w_m, self.w_v]
        return [(t, t.clone()) for t in ts]

    def _restore(self, snap, state_backup):
        for t, c in snap:
            t.copy_(c)
        self.state.copy_(state_backup)

    def reset_meter(self):
        self.acc.zero_()

    def meter_avg(self):
        """losses.avg of the reference's AverageMeter (utils.py:126-135); one D2H read per epoch."""
        s, c = self.acc.tolist()
        return s / c if c else 0.0


# --------------------------------------------------------------------------- evaluation
class ListCSR:
    """Per-user item lists as a sorted CSR on the device (+ host copies for the tile bucketing)."""

    def __init__(self, lists, device, sort=True):
        ptr_, flat = lists_to_arrays(lists)
        self._init_arrays(ptr_, flat, device, sort)

    @classmethod
    def from_arrays(cls, ptr_, items, device, sort=True):
        """Same object from numpy CSR arrays (ptr int64 [n+1], items [nnz]) -- no Python-level iteration."""
        self = object.__new__(cls)
        self._init_arrays(np.ascontiguousarray(ptr_, dtype=np.int64), np.asarray(items, dtype=np.int64), device, sort)
        return self

    def _init_arrays(self, ptr_, flat, device, sort):
        lens = np.diff(ptr_)
        if sort and len(flat):
            rows = np.repeat(np.arange(len(lens), dtype=np.int64), lens)
            flat = flat[np.lexsort((flat, rows))]
        self.lens = lens
        self.ptr_host, self.items_host = ptr_, flat.astype(np.int32)
        self.ptr = torch.from_numpy(ptr_).to(device)
        self.items = torch.from_numpy(self.items_host).to(device)
        self.device = device
        self._tiles = {}

    def __getitem__(self, i):          # (ptr, items, lens) tuple view used by older call sites
        return (self.ptr, self.items, self.lens)[i]

    def tiles(self, n_items, users_host=None, order=None):
        """Seen-item pairs bucketed by (128-user tile, 256-item tile) for the tensor-core kernel:
        (tile_ptr int32 [n_utiles, n_itiles + 1], entries uint16 ((row << 8) | col)).  With an ItemOrder the item
        coordinate is the item's POSITION in the scan order."""
        key = (int(n_items), None if users_host is None else users_host.tobytes(), None if order is None else order.uid)
        hit = self._tiles.get(key)
        if hit is None:
            ptr_, items = self.ptr_host, self.items_host.astype(np.int64)
            if order is not None:
                items = order.pos_host[items]
            if users_host is None:
                n_eval = len(ptr_) - 1
                pos = np.repeat(np.arange(n_eval, dtype=np.int64), np.diff(ptr_))
            else:
                n_eval = len(users_host)
                lens = ptr_[users_host + 1] - ptr_[users_host]
                pos = np.repeat(np.arange(n_eval, dtype=np.int64), lens)
                start = np.repeat(ptr_[users_host], lens)
                off = np.arange(len(pos), dtype=np.int64) - np.repeat(np.cumsum(lens) - lens, lens)
                items = items[start + off]
            n_ut, n_it = (n_eval + 127) // 128, (n_items + 255) // 256
            keyv = (pos // 128) * n_it + items // 256
            order = np.argsort(keyv, kind='stable')
            ent = (((pos % 128) << 8) | (items % 256)).astype(np.uint16)[order]
            counts = np.bincount(keyv, minlength=n_ut * n_it)
            excl = np.zeros(n_ut * n_it + 1, dtype=np.int64)
            np.cumsum(counts, out=excl[1:])
            tile_ptr = np.empty((n_ut, n_it + 1), dtype=np.int32)
            tile_ptr[:, :n_it] = excl[:-1].reshape(n_ut, n_it)
            tile_ptr[:, n_it] = excl[np.arange(1, n_ut + 1) * n_it]
            hit = (torch.from_numpy(tile_ptr).to(self.device),
                   torch.from_numpy(ent.view(np.int16).copy()).to(self.device))
            self._tiles = {key: hit}
        return hit


def lists_to_arrays(lists):
    """list-of-lists -> (ptr int64 [n+1], items int64 [nnz]) numpy arrays (order inside a list kept)."""
    lens = np.fromiter(map(len, lists), dtype=np.int64, count=len(lists))
    ptr_ = np.zeros(len(lists) + 1, dtype=np.int64)
    np.cumsum(lens, out=ptr_[1:])
    flat = np.fromiter((i for x in lists for i in x), dtype=np.int64, count=int(ptr_[-1]))
    return ptr_, flat


def merge_csr(a, b):
    """Row-wise concatenation of two CSR array pairs with the same number of rows."""
    (pa, ia), (pb, ib) = a, b
    la, lb = np.diff(pa), np.diff(pb)
    ptr_ = np.zeros(len(pa), dtype=np.int64)
    np.cumsum(la + lb, out=ptr_[1:])
    out = np.empty(int(ptr_[-1]), dtype=np.int64)
    ra = np.repeat(ptr_[:-1], la) + (np.arange(len(ia), dtype=np.int64) - np.repeat(pa[:-1], la))
    rb = np.repeat(ptr_[:-1] + la, lb) + (np.arange(len(ib), dtype=np.int64) - np.repeat(pb[:-1], lb))
    out[ra], out[rb] = ia, ib
    return ptr_, out


def restrict_csr(csr, user_lo, user_hi, item_lo, item_hi):
    """Rows outside [user_lo, user_hi) emptied, items outside [item_lo, item_hi) dropped (the list surgery of
    BasicTrainer.inductive_eval, trainer.py:185-217, on arrays)."""
    ptr_, items = csr
    rows = np.repeat(np.arange(len(ptr_) - 1, dtype=np.int64), np.diff(ptr_))
    keep = (rows >= user_lo) & (rows < user_hi) & (items >= item_lo) & (items < item_hi)
    new_ptr = np.zeros(len(ptr_), dtype=np.int64)
    np.cumsum(np.bincount(rows[keep], minlength=len(ptr_) - 1), out=new_ptr[1:])
    return new_ptr, items[keep]


def lists_to_csr(lists, device, sort=True):
    """list-of-lists -> ListCSR (indexable as (ptr, items, lens))."""
    return ListCSR(lists, device, sort)


class ItemOrder:
    """Scan order of the items for the tensor-core scoring path: perm[p] = item at position p, pos[i] = position of
    item i.  A static permutation (the trainers pass train popularity, most popular first): the result does not
    depend on it, the cost does -- items that rank high for most users come first, the per-user thresholds tighten
    within the first tiles and the filter's compare-free path takes almost every later chunk."""
    _uids = iter(range(1, 1 << 62))

    def __init__(self, perm_host, device):
        perm_host = np.ascontiguousarray(perm_host, dtype=np.int64)
        n = len(perm_host)
        pos = np.empty(n, dtype=np.int64)
        pos[perm_host] = np.arange(n, dtype=np.int64)
        if not np.array_equal(np.sort(perm_host), np.arange(n, dtype=np.int64)):
            raise ValueError('item order must be a permutation of range(n_items)')
        self.uid = next(ItemOrder._uids)
        self.perm_host, self.pos_host = perm_host, pos
        self.perm = torch.from_numpy(perm_host.astype(np.int32)).to(device)
        self._bits = {}

    @classmethod
    def by_score(cls, item_score, device):
        """Descending `item_score` (e.g. train degree), ties by item id."""
        item_score = np.asarray(item_score)
        return cls(np.argsort(-item_score, kind='stable'), device)

    def position_bits(self, n_items, item_lo, item_hi, banned_bits):
        """Bitmap over POSITIONS of the items outside [item_lo, item_hi) or banned (None when nothing is excluded)."""
        if item_lo <= 0 and item_hi >= n_items and banned_bits is None:
            return None
        key = (int(item_lo), int(item_hi), None if banned_bits is None else banned_bits.data_ptr())
        hit = self._bits.get(key)
        if hit is None or hit[0] is not banned_bits:
            # host-side (tiny, cached): only the inductive evaluation passes restrict or ban items
            ids = np.arange(n_items, dtype=np.int64)
            bad = (ids < item_lo) | (ids >= item_hi)
            if banned_bits is not None:
                words = banned_bits.cpu().numpy().view(np.uint32)
                bad |= ((words[ids >> 5] >> (ids & 31).astype(np.uint32)) & 1).astype(bool)
            from .graph import _pack_bits
            hit = (banned_bits, _pack_bits(bad[self.perm_host], self.perm.device))
            self._bits = {key: hit}
        return hit[1]


FALLBACK_SPLITS, FALLBACK_SPLIT_CAP = 64, 2048      # item ranges / max users for the split form of the exact kernel


def score_topk_exact(rep, user_ids, n_users, n_items, k, mask=None, item_lo=0, item_hi=None, banned_bits=None,
                     out=None, out_rows=None, n_eval_dev=None, split_keys=None):
    """Exact CUDA-core kernel (also the tensor-core path's fallback; `split_keys` = scratch of
    FALLBACK_SPLIT_CAP * FALLBACK_SPLITS * k uint64 enables the item-split form for short user lists)."""
    n = int(user_ids.shape[0])
    if out is None:
        out = (torch.empty((n, k), dtype=torch.int32, device=rep.device),
               torch.empty((n, k), dtype=torch.float32, device=rep.device))
    mptr, mitems = (None, None) if mask is None else (mask[0], mask[1])
    call('igcn_score_topk_exact', ptr(rep), ptr(user_ids), n, n_users, n_items, rep.shape[1], ptr(mptr), ptr(mitems),
         int(item_lo), int(n_items if item_hi is None else item_hi), ptr(banned_bits), int(k), ptr(out[0]), ptr(out[1]),
         ptr(out_rows), ptr(n_eval_dev), 1 if split_keys is None else FALLBACK_SPLITS, ptr(split_keys),
         0 if split_keys is None else FALLBACK_SPLIT_CAP, stream_ptr())
    return out


class TcScorer:
    """tcgen05 scoring pipeline: pack -> candidates -> finalize -> exact fallback (engine-owned buffers)."""

    def __init__(self):
        self._ws = {}
        self.last_fallback = None       # device int32[1]: users sent to the exact kernel by the last call

    @staticmethod
    def pick_splits(n_groups, n_sm=148):
        """Uniform item-range splits per user tile when there are fewer user tiles than SMs: as many as fit in ONE
        wave, at most 8 (147 tiles -> 1, 74 -> 2, 30 -> 4).  Round 1 filled two waves (better tail balance); with
        the cheap epilogue of round 2 the lists are what costs -- every split is one more list per user for
        igcn_tc_finalize to re-score and rank, and at N GPUs every rank is in this regime (4-GPU Yelp shape: 147 tiles
        per rank in 3 splits spent as long in finalize, 0.17 ms, as in the candidate kernel)."""
        return int(min(8, max(1, n_sm // max(1, int(n_groups)))))

    @classmethod
    def plan_ctas(cls, n_groups, n_itiles=None, n_sm=148):
        """(n_head, n_splits) of igcn_tc_candidates.  Every extra list per user costs a threshold warm-up in the
        epilogue and a longer igcn_tc_finalize, so with at least one wave of user tiles only the TAIL (the tiles left
        over after whole waves of n_sm CTAs) is split, just enough to fill the last wave: 297 tiles -> 296 unsplit + 1
        tile in 8 splits.  Fewer user tiles than SMs: uniform splits (pick_splits).
        (A makespan model that also split whole waves -- Gowalla shape: 234 tiles as 148 + 86 x 5 instead of 234
        unsplit in 1.58 waves -- was measured in round 2: igcn_tc_candidates 0.254 -> 0.272 ms and igcn_tc_finalize
        0.091 -> 0.253 ms; five lists per user cost more than the fuller last wave gains.  Not used.)"""
        forced = int(os.environ.get('IGCN_TC_SPLITS', 0))
        if forced:
            return 0, forced
        if n_groups < n_sm:
            return 0, cls.pick_splits(n_groups, n_sm)
        tail = n_groups % n_sm
        splits = 1 if tail == 0 else int(min(8, n_sm // tail))
        return (n_groups, 1) if splits == 1 else (n_groups - tail, splits)

    def _workspace(self, n_eval, n_items, D, n_splits, k, device):
        key = (n_eval, n_items, D, n_splits, k, str(device))
        ws = self._ws.get(key)
        if ws is None:
            a_b, b_b, slots = C.c_int64(), C.c_int64(), C.c_int64()
            call('igcn_tc_workspace', n_eval, n_items, D, n_splits, C.byref(a_b), C.byref(b_b), C.byref(slots))
            z = lambda n, dt: torch.zeros(n, dtype=dt, device=device)
            ws = {'a_img': z(a_b.value, torch.uint8), 'b_img': z(b_b.value, torch.uint8),
                  'cand_items': z(slots.value, torch.int32), 'cand_cnt': z(n_eval * n_splits, torch.int32),
                  'cand_thr': z(n_eval * n_splits, torch.float32), 'maxabs': z(1, torch.int32),
                  'fb_count': z(1, torch.int32), 'fb_users': z(n_eval, torch.int64), 'fb_rows': z(n_eval, torch.int32),
                  'split_keys': z(FALLBACK_SPLIT_CAP * FALLBACK_SPLITS * k, torch.int64),
                  'center': z(D, torch.float32), 'center_scratch': z(((n_items + 255) // 256 + 1) * D, torch.float32)}
            self._ws = {key: ws}
        return ws

    def topk(self, rep, user_ids, n_users, n_items, k, mask=None, item_lo=0, item_hi=None, banned_bits=None,
             users_host=None, n_splits=None, dump=False, order=None, stats=None):
        """order: ItemOrder (scan order of the items) or None; stats: uint64 [5] device tensor (instrumented kernel)."""
        n_eval, D = int(user_ids.shape[0]), int(rep.shape[1])
        item_hi = n_items if item_hi is None else item_hi
        # the candidate kernel works in position space: with a scan order, ranges / banned items become one bitmap
        # over positions, and the exact fallback keeps the caller's item-id arguments
        c_lo, c_hi, c_bits, perm = item_lo, item_hi, banned_bits, None
        if order is not None:
            if len(order.perm_host) != n_items:
                raise RuntimeError('item order has %d entries, the catalogue %d' % (len(order.perm_host), n_items))
            perm = order.perm
            c_bits = order.position_bits(n_items, item_lo, item_hi, banned_bits)
            c_lo, c_hi = 0, n_items
        n_head = 0
        if n_splits is None:
            n_head, n_splits = self.plan_ctas((n_eval + 127) // 128, (n_items + 255) // 256)
        ws = self._workspace(n_eval, n_items, D, n_splits, k, rep.device)
        st = stream_ptr
        call('igcn_tc_pack', ptr(rep), rep.numel(), ptr(user_ids), n_eval, n_users, n_items, D, ptr(perm), ptr(ws['maxabs']),
             ptr(ws['a_img']), ptr(ws['b_img']), ptr(ws['center']), ptr(ws['center_scratch']), st())
        tile_ptr, entries = (None, None)
        if mask is not None:
            tile_ptr, entries = mask.tiles(n_items, users_host, order)
        dump_t = None
        if dump:
            dump_t = torch.zeros(((n_eval + 127) // 128 * 128, (n_items + 255) // 256 * 256), dtype=torch.float32,
                                 device=rep.device)
        call('igcn_tc_candidates', ptr(ws['a_img']), ptr(ws['b_img']), n_eval, n_items, D, n_splits, n_head,
             int(c_lo), int(c_hi), ptr(c_bits), ptr(tile_ptr), ptr(entries), ptr(ws['cand_items']), ptr(ws['cand_cnt']),
             ptr(ws['cand_thr']), ptr(dump_t), ptr(stats), st())
        out_i = torch.empty((n_eval, k), dtype=torch.int32, device=rep.device)
        out_s = torch.empty((n_eval, k), dtype=torch.float32, device=rep.device)
        call('igcn_tc_finalize', ptr(rep), ptr(user_ids), n_eval, n_users, D, n_splits, ptr(ws['cand_items']),
             ptr(ws['cand_cnt']), ptr(ws['cand_thr']), ptr(ws['maxabs']), ptr(ws['center']), n_items, ptr(perm), int(k), ptr(out_i), ptr(out_s),
             ptr(ws['fb_count']), ptr(ws['fb_users']), ptr(ws['fb_rows']), st())
        # users whose bound did not verify: exact kernel on the device-side list (no host sync)
        score_topk_exact(rep, ws['fb_users'], n_users, n_items, k, mask, item_lo, item_hi, banned_bits,
                         out=(out_i, out_s), out_rows=ws['fb_rows'], n_eval_dev=ws['fb_count'], split_keys=ws['split_keys'])
        self.last_fallback = ws['fb_count']
        if dump:
            return out_i, out_s, dump_t, ws
        return out_i, out_s


_tc_scorer = TcScorer()


def score_topk(rep, user_ids, n_users, n_items, k, mask=None, item_lo=0, item_hi=None, banned_bits=None,
               users_host=None, impl='auto', order=None):
    """Fused scoring + seen-item mask + top-k for `user_ids` (model.py:118-123 + trainer.py:149-164).
    Returns (items int32 [n, k], scores fp32 [n, k]), sorted by (score desc, item asc); slots without
    a candidate hold -1 / -inf.  impl: 'tc' (tcgen05 path, D <= 64 and k <= 24), 'exact', or 'auto'.  order: an
    ItemOrder, the order in which the tensor-core path scans the catalogue (cost only, never the result)."""
    _lib.require_cuda(rep, torch.float32, 'rep')
    tc_ok = rep.shape[1] <= 64 and k <= 24 and (mask is None or isinstance(mask, ListCSR))
    if impl == 'tc' and not tc_ok:
        raise RuntimeError('tensor-core scoring needs D <= 64, k <= 24 and a ListCSR mask')
    if impl == 'exact' or not tc_ok:
        return score_topk_exact(rep, user_ids, n_users, n_items, k, mask, item_lo, item_hi, banned_bits)
    if isinstance(users_host, str):          # 'identity': the caller vouches for user_ids == arange(len(mask))
        users_host = None
    elif mask is not None and users_host is None:
        n = int(user_ids.shape[0])
        if n != len(mask.ptr_host) - 1 or not bool((user_ids == torch.arange(n, device=user_ids.device)).all()):
            users_host = user_ids.cpu().numpy()
    return _tc_scorer.topk(rep, user_ids, n_users, n_items, k, mask, item_lo, item_hi, banned_bits, users_host, order=order)


def hit_matrix(rec, eval_csr):
    """hit[u, j] = rec[u, j] in eval_data[u] (trainer.py:111-115) as fp32 on the device."""
    n, k = rec.shape
    hit = torch.empty((n, k), dtype=torch.float32, device=rec.device)
    call('igcn_hits', ptr(rec), n, k, ptr(eval_csr[0]), ptr(eval_csr[1]), ptr(hit), stream_ptr())
    return hit
