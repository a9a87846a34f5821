"""Sibling models of the reference that reuse the propagation and ranking kernels:

    NGCF    model.py:232-299   row-normalised (A + I), edge dropout, two dense layers per hop, concatenated hops
    IMCGAE  model.py:546-585   LightGCN's adjacency on [personal | general | identical] embeddings, node dropout

Their propagation is `dgl.ops.gspmm(g, 'mul', 'sum', X, A.values())` (model.py:281, 576) like LightGCN's; here it is
one autograd node per hop on `igcn_spmm` (forward: the CSR of A, backward: the CSR of A^T), the dense layers /
activations between the hops stay torch modules (plain library GEMMs).  The trainer runs them through the generic
autograd step (trainer.AutogradStep): no fused step, no sharded training -- at N > 1 the evaluation users are sharded
like everywhere else.  Evaluation goes through the same fused score + mask + top-k kernel (exact CUDA-core form: the
representations are 256 / 192 columns wide, the tensor-core form stops at 64).

Random draws: train-mode dropout masks come from torch's CUDA generator -- at N > 1 the training replicas stay identical
only if every rank seeds it alike (the reference launchers call set_seed(2021) first); `model.injected` (a test
facility) replays masks recorded from the reference -- {'edge': bool [nnz] in the reference's coalesced COO order, 'dense': [bool
tensors, one per F.dropout call of a forward pass]}."""
import ctypes as C

import torch
import torch.nn as nn
import torch.nn.functional as F
from torch.nn.init import kaiming_uniform_, normal_, zeros_

from . import graph
from ._lib import call, ptr, stream_ptr, require_cuda
from .model import _GraphModel

_BLOCK = 64          # columns per igcn_spmm call for representations wider than the kernel's 128


def spmm_raw(csr, x):
    """Y = A X for a graph.CsrDevice A and a dense fp32 X [n_cols, D] (any D % 4 == 0; wide X in 64-column blocks)."""
    x = x.contiguous()
    D = x.shape[1]
    y = torch.empty((csr.n_rows, D), dtype=torch.float32, device=x.device)
    no_adds = (C.c_void_p * 1)()
    if D <= 128:
        call('igcn_spmm', csr.struct(D), ptr(x), ptr(y), D, no_adds, 0, None, 1.0, None, 0, stream_ptr())
        return y
    if D % _BLOCK:
        raise RuntimeError('wide representations must be a multiple of %d columns, got %d' % (_BLOCK, D))
    for c0 in range(0, D, _BLOCK):
        xb = x[:, c0:c0 + _BLOCK].contiguous()
        yb = torch.empty((csr.n_rows, _BLOCK), dtype=torch.float32, device=x.device)
        call('igcn_spmm', csr.struct(_BLOCK), ptr(xb), ptr(yb), _BLOCK, no_adds, 0, None, 1.0, None, 0, stream_ptr())
        y[:, c0:c0 + _BLOCK] = yb
    return y


class SpMM(torch.autograd.Function):
    """One propagation hop as an autograd node: forward A X, backward A^T G (both igcn_spmm)."""

    @staticmethod
    def forward(ctx, x, fwd_csr, bwd_csr):
        ctx.bwd_csr = bwd_csr
        return spmm_raw(fwd_csr, x.detach())

    @staticmethod
    def backward(ctx, g):
        return spmm_raw(ctx.bwd_csr, g), None, None


class RowNormAdj(graph._SparseView, graph._Blocked):
    """normalize(A + I, norm='l1', axis=1) as a device CSR (NGCF.generate_graph, model.py:255-261) together with its
    transpose: the pattern is symmetric, so A^T shares the index arrays; its values are those of A permuted (`perm`:
    position in A of the entry (c, r) for every entry (r, c) in CSR order), which is also how an edge-dropout mask
    drawn in A's order reaches the backward product."""

    def __init__(self, dg):
        n = dg.n_users + dg.n_items
        dev = dg.device
        self.n_users, self.n_items = dg.n_users, dg.n_items
        self.shape = torch.Size([n, n])
        rows = torch.cat([dg.rows(), torch.arange(n, device=dev)])
        cols = torch.cat([dg.col.long(), torch.arange(n, device=dev)])
        mult = torch.ones(rows.shape[0], dtype=torch.float64, device=dev)
        if dg.mult is not None:
            mult[:dg.col.shape[0]] = dg.mult.double()
        key, order = torch.sort(rows * n + cols)                      # row-major, the diagonal at its sorted place
        rows, cols, mult = rows[order], cols[order], mult[order]
        deg = torch.zeros(n, dtype=torch.float64, device=dev).index_add_(0, rows, mult)
        # sp.eye is float64, so the reference normalises in float64 and rounds to float32 at the very end
        val = (mult / deg[rows]).to(torch.float32)
        rowptr = torch.zeros(n + 1, dtype=torch.int64, device=dev)
        torch.cumsum(torch.bincount(rows, minlength=n), 0, out=rowptr[1:])
        self.perm = torch.argsort(cols * n + rows)
        self.rows_idx, self.cols_idx = rows, cols
        self.csr_fwd = graph.CsrDevice(rowptr.cpu().numpy(), cols.to(torch.int32), val, n, dev)
        self.csr_bwd = self.csr_fwd.with_values(val[self.perm].contiguous())
        self.nnz = int(rows.shape[0])
        self._set_blocks([graph.RowBlock(0, n, self.csr_fwd)])
        self._sampler = (dg.rowptr[:dg.n_users + 1], dg.col[:dg.n_interactions])

    def sampler_csr(self):
        return self._sampler

    def _coo(self):
        return torch.stack([self.rows_idx, self.cols_idx]), self.csr_fwd.val

    def pair(self, keep=None, p=0.):
        """(forward CSR, backward CSR) with the edges where keep is False removed and the others scaled by
        1 / (1 - p) (NGCF.dropout_sp_mat, model.py:263-275); a dropped edge stays in the pattern with value 0."""
        if keep is None:
            return self.csr_fwd, self.csr_bwd
        val = self.csr_fwd.val * keep.to(torch.float32) / (1. - p)
        return self.csr_fwd.with_values(val), self.csr_bwd.with_values(val[self.perm].contiguous())


class _Sibling(_GraphModel):
    fused_step = False          # trained by trainer.AutogradStep

    def _no_training_shard(self):
        self._shard_rows, self._shard_auto, self._dim_shard = False, False, None

    def _take(self, kind):
        inj = getattr(self, 'injected', None)
        if not inj:
            return None
        if kind == 'edge':
            return inj.pop('edge', None)
        dense = inj.get('dense')
        return dense.pop(0) if dense else None

    def _dropout(self, x, p):
        """F.dropout(x, p, training) with an optional recorded keep mask."""
        if not self.training:
            return x
        keep = self._take('dense')
        if keep is None:
            return F.dropout(x, p=p, training=True)
        return x * keep.to(device=x.device, dtype=x.dtype) / (1. - p)

    def bpr_forward(self, users, pos_items, neg_items):
        """NGCF.bpr_forward (model.py:293-299): L2 over the propagated rows."""
        rep = self.get_rep()
        users_r = rep[users, :]
        pos_r, neg_r = rep[self.n_users + pos_items, :], rep[self.n_users + neg_items, :]
        l2_norm_sq = torch.norm(users_r, p=2, dim=1) ** 2 + torch.norm(pos_r, p=2, dim=1) ** 2 \
            + torch.norm(neg_r, p=2, dim=1) ** 2
        return users_r, pos_r, neg_r, l2_norm_sq

    def _cache_key(self):
        # every parameter, not just the embedding table (the dense layers are trained too)
        return (self.graph_version(), self._param_epoch) + tuple((id(p), p._version) for p in self.parameters())

    def get_rep(self):
        self._check_graph()
        require_cuda(self.embedding.weight, torch.float32, 'embedding.weight')
        return self._cached_rep(self._compute_rep)


def _dense_layer(n_in, n_out):
    layer = nn.Linear(n_in, n_out)          # model.py:24-28
    kaiming_uniform_(layer.weight)
    zeros_(layer.bias)
    return layer


class NGCF(_Sibling):
    """model.py:232-299."""

    def __init__(self, model_config):
        super().__init__(model_config)
        self._no_training_shard()
        self.dropout = model_config['dropout']
        self.embedding_size = model_config['embedding_size']
        self.layer_sizes = list(model_config['layer_sizes'])
        self.embedding = nn.Embedding(self.n_users + self.n_items, self.embedding_size)
        kaiming_uniform_(self.embedding.weight)
        self.n_layers = len(self.layer_sizes)
        self.layer_sizes.insert(0, self.embedding_size)
        gc, bi = [], []
        for l in range(1, self.n_layers + 1):           # same construction order as the reference (generator draws)
            gc.append(_dense_layer(self.layer_sizes[l - 1], self.layer_sizes[l]))
            bi.append(_dense_layer(self.layer_sizes[l - 1], self.layer_sizes[l]))
        self.gc_layers, self.bi_layers = nn.ModuleList(gc), nn.ModuleList(bi)
        self.norm_adj = self.generate_graph(model_config['dataset'])
        self.to(device=self.device)

    def generate_graph(self, dataset):
        dg = getattr(dataset, 'device_graph', None)
        if dg is None:
            dg = graph.DeviceGraph.from_pairs(dataset.n_users, dataset.n_items, graph.train_pairs_of(dataset), self.device)
        return RowNormAdj(dg)

    def _edge_pair(self):
        if not self.training:
            return self.norm_adj.pair()
        keep = self._take('edge')
        if keep is None:
            if self.dropout <= 0.:
                return self.norm_adj.pair()
            keep = torch.floor(1. - self.dropout + torch.rand(self.norm_adj.nnz, device=self.device)).to(torch.bool)
        return self.norm_adj.pair(torch.as_tensor(keep, device=self.device), self.dropout)

    def _compute_rep(self):
        rep = self.embedding.weight
        hops = [rep]
        fwd, bwd = self._edge_pair()
        for l in range(self.n_layers):
            m0 = SpMM.apply(rep, fwd, bwd)
            m1 = rep * m0
            rep = F.leaky_relu(self.gc_layers[l](m0) + self.bi_layers[l](m1), negative_slope=0.2)
            rep = self._dropout(rep, self.dropout)
            hops.append(F.normalize(rep, p=2, dim=1))
        return torch.cat(hops, dim=1)


class IMCGAE(_Sibling):
    """model.py:546-585."""

    def __init__(self, model_config):
        super().__init__(model_config)
        self._no_training_shard()
        self.embedding_size = model_config['embedding_size']
        self.n_layers = model_config['n_layers']
        self.dropout = model_config['dropout']
        self.embedding = nn.Embedding(self.n_users + self.n_items + 3, self.embedding_size)
        self.norm_adj = self.generate_graph(model_config['dataset'])
        normal_(self.embedding.weight, std=0.1)
        self.to(device=self.device)

    def generate_graph(self, dataset):
        """LightGCN.generate_graph (model.py:553-554); the matrix is symmetric, so it is its own transpose."""
        dg = self._device_graph_of(dataset)
        if dg is not None:
            return graph.NormAdj.from_device(dg)
        return graph.NormAdj(dataset.n_users, dataset.n_items, graph.train_pairs_of(dataset), self.device)

    def _compute_rep(self):
        w, U, I = self.embedding.weight, self.n_users, self.n_items
        ident, gen_u, gen_i = w[U + I], w[U + I + 1], w[U + I + 2]
        u_rep = torch.cat([w[:U], gen_u[None, :].expand(U, -1), ident[None, :].expand(U, -1)], dim=1)
        i_rep = torch.cat([w[U:U + I], gen_i[None, :].expand(I, -1), ident[None, :].expand(I, -1)], dim=1)
        rep = torch.cat([u_rep, i_rep], dim=0)
        total = rep
        csr = self.norm_adj.csr
        for l in range(self.n_layers):
            mask = self._dropout(torch.ones(U + I, dtype=torch.float32, device=rep.device), self.dropout - 0.1 * l)
            rep = SpMM.apply(rep * mask[:, None], csr, csr)
            total = total + rep / float(l + 2)
        return total
