"""Drop-in model layer: `LightGCN` and `IGCN` (INMO-LightGCN) with the reference's public surface
(reference model.py:16-49, 75-123, 354-466) on top of the sm_100a kernels.

What is the same: class names, constructor configs, attribute names (`embedding`, `norm_adj`,
`feat_mat`, `row_sum`, `user_map`, `item_map`, `alpha`, `w`, ...), method names and return values,
the parameter-initialisation order (so a given torch seed yields the same initial weights), the
checkpoint format (including the 'sate_dict' key of model.py:455) and the external mutation
sequence of run/dropui/igcn_dropui.py:28-32.

What differs: `norm_adj` / `feat_mat` are device CSR objects (igcn_cf_b200.graph) built once per
generate_* call; `get_rep` is a single autograd node backed by libigcn_b200.so; in eval mode the
representation is cached until a parameter or graph changes (the reference recomputes the full
propagation for every 512-user batch, model.py:119); there is no CPU path.
"""
import os
import sys

import numpy as np
import scipy.sparse as sp
import torch
import torch.nn as nn
from torch.nn.init import normal_

from . import dist, engine, graph
from ._lib import require_cuda


def get_model(config, dataset):
    """Name-dispatched factory (model.py:16-21)."""
    config = config.copy()
    config['dataset'] = dataset
    cls = getattr(sys.modules[__name__], config['name'], None)
    if cls is None or config.get('out_of_scope'):
        raise NotImplementedError('model %r is a baseline outside the B200 hot path (SURVEY.md 2.1); implemented: '
                                  'LightGCN, IGCN, IMF, MF, NGCF, IMCGAE, Popularity' % config['name'])
    return cls(config)


def _device_of(cfg):
    dev = torch.device(cfg['device'])
    if dev.type != 'cuda':
        raise RuntimeError('igcn_cf_b200 models run on CUDA devices only (got %s): there is no CPU fallback' % dev)
    return dev


class BasicModel(nn.Module):
    """model.py:31-49."""

    def __init__(self, model_config):
        super().__init__()
        print(model_config)
        self.config = model_config
        self.name = model_config['name']
        self.device = _device_of(model_config)
        self.n_users = model_config['dataset'].n_users
        self.n_items = model_config['dataset'].n_items
        self.trainable = True
        # multi-GPU (igcn_cf_b200.dist.init_peers() was called, world > 1).  model_config['shard']:
        #   False   single-GPU behaviour on every rank
        #   'users' training step replicated, evaluation users sharded (no exchange anywhere on the training path)
        #   True    propagation rows sharded over the ranks (fused NVLink all-gather), eval users sharded
        #   'dims'  training with the embedding COLUMNS sharded: every rank owns embedding_size / world columns of the
        #           parameters, the optimizer state and every layer buffer.  The propagation is linear and acts on each
        #           column independently, so no layer embedding ever crosses NVLink; the only exchange of a step is 20-28
        #           bytes per triple (partial dot products), and the parameters are all-gathered once per epoch /
        #           before an evaluation.  Bit-identical to one GPU.  Eval users sharded, eval propagation column-sharded too.
        #   'auto'  (default) eval users always sharded; rows sharded only when every rank keeps at least
        #           SHARD_MIN_NNZ_PER_RANK non-zeros (graphs whose layer tables live in HBM: the scale-out class, where
        #           one rank's propagation is tens of milliseconds); otherwise columns when 2, 4 or 8 ranks divide the
        #           embedding size into multiples of 4 -- measured fastest on every paper-sized graph --
        #           else replicated training
        mode = model_config.get('shard', 'auto')
        self._peers = dist.current() if mode else None
        self._shard_rows = mode is True
        self._shard_auto = mode == 'auto'
        D = model_config.get('embedding_size', 0)
        self._dims_ok = (self._peers is not None and self._peers.world in (2, 4, 8) and D and D % (4 * self._peers.world) == 0)
        self._dim_shard = (self._peers.rank, self._peers.world) if (mode == 'dims' and self._peers is not None) else None
        self._param_sync = None         # set by engine.TrainStep in 'dims' mode: all-gathers the column slices

    def predict(self, users):
        raise NotImplementedError

    def save(self, path):
        sync = getattr(self, '_sync_params', None)
        if callable(sync):
            sync()
        torch.save(self.state_dict(), path)

    def load(self, path):
        self.load_state_dict(torch.load(path, map_location=self.device))


class _GraphModel(BasicModel):
    """Machinery shared by LightGCN and IGCN: propagator buffers, eval-mode cache, predict."""
    _prop = None
    _rep_cache = None
    _param_epoch = 0

    def _bump(self):
        """Tell the eval-mode cache that parameters were changed behind autograd's back."""
        self._param_epoch += 1

    def _sync_params(self):
        """Column-sharded training keeps the up-to-date parameters as per-rank column slices: gather them into
        embedding.weight (collective, a no-op when nothing is pending) before anything reads full-width parameters."""
        if self._param_sync is not None:
            self._param_sync()

    def graph_version(self):
        """Identity of the graph objects the kernels read: their construction serial numbers (graph._Blocked.uid),
        never id() -- CPython reuses ids of freed objects."""
        feat = getattr(self, 'feat_mat', None)
        return (getattr(self.norm_adj, 'uid', None), None if feat is None else getattr(feat, 'uid', None),
                self.n_users, self.n_items)

    SHARD_MIN_NNZ_PER_RANK = int(os.environ.get('IGCN_SHARD_MIN_NNZ_PER_RANK', 50_000_000))

    def _rows_sharded(self, dataset=None):
        if self._peers is None:
            return False
        if self._shard_auto and dataset is not None:
            dg = getattr(dataset, 'device_graph', None)
            nnz = 2 * (dg.n_interactions if dg is not None else len(graph.train_pairs_of(dataset)))
            self._shard_rows = nnz >= self.SHARD_MIN_NNZ_PER_RANK * self._peers.world
            self._shard_auto = False                     # decided once per model: all graphs of a model agree
            if not self._shard_rows and self._dims_ok:
                self._dim_shard = (self._peers.rank, self._peers.world)
        return self._shard_rows

    def _shard_arg(self, dataset=None):
        return (self._peers.rank, self._peers.world) if self._rows_sharded(dataset) else None

    _dg_cache = None

    def _device_graph_of(self, dataset):
        """The train graph of `dataset` as a graph.DeviceGraph: the dataset's own (scale-out datasets), or the CSR
        built ON THE DEVICE from its train pairs (graph.DeviceGraph.from_pairs; SURVEY.md 8f rank 1 -- the
        inductive update of run/dropui rebuilds graph and templates in milliseconds instead of a scipy pass
        each).  generate_graph and generate_feat of the same dataset share one build.  model_config
        'graph_builder': 'host' selects the numpy/scipy constructors instead (bit-identical structures)."""
        dg = getattr(dataset, 'device_graph', None)
        if dg is not None:
            return dg
        if self.config.get('graph_builder', 'device') == 'host':
            return None
        pairs = graph.train_pairs_of(dataset)
        # the cache holds references to the dataset and its pair array, so their ids cannot be recycled while the
        # entry is alive; `is` comparisons instead of id() keys
        c = self._dg_cache
        if c is None or c[0] is not dataset or c[1] is not pairs or c[2] != (len(pairs), dataset.n_users, dataset.n_items):
            dg = graph.DeviceGraph.from_pairs(dataset.n_users, dataset.n_items, pairs, self.device)
            self._dg_cache = c = (dataset, pairs, (len(pairs), dataset.n_users, dataset.n_items), dg)
        return c[3]

    def _propagator(self):
        n = self.n_users + self.n_items
        block = self.norm_adj.block_key()
        p = self._prop
        if p is None or (p.n, p.dim, p.n_layers, p.block) != (n, self.embedding_size, self.n_layers, block):
            shard = engine.Shard(self._peers) if self._rows_sharded() else None
            p = engine.Propagator(n, self.embedding_size, self.n_layers, self.device, shard)
            p.block = block
            self._prop = p
            self._grad_buf = None
        return p

    def _grad_buffer(self, rows):
        """Symmetric [rows, D] gradient buffer for the autograd bridge of the row-sharded INMO layer."""
        if getattr(self, '_grad_buf', None) is None or self._grad_buf.shape[0] != rows:
            self._grad_buf = self._propagator().new_buffer(rows)
        return self._grad_buf

    def _cache_key(self):
        w = self.embedding.weight
        return (self.graph_version(), id(w), w._version, self._param_epoch, getattr(self, 'alpha', None))

    def _check_graph(self):
        n = self.n_users + self.n_items
        if self.norm_adj.shape[0] != n:
            raise RuntimeError('norm_adj has %d rows but the model has %d nodes; regenerate the graph'
                               % (self.norm_adj.shape[0], n))

    _col_rep = None

    def _column_rep(self):
        """Column-sharded models evaluate column-sharded too (engine.ColumnRep): each rank propagates its
        embedding_size / world columns and the slices are all-gathered over NVLink.  Collective."""
        c = self._col_rep
        if c is None or c.key != engine.ColumnRep.key_of(self):
            c = self._col_rep = engine.ColumnRep(self)
        return c.run(self)

    def _cached_rep(self, compute):
        """Eval mode: the representation is a pure function of parameters and graph (model.py:264-265
        makes dropout the identity), so compute it once."""
        if self.training or torch.is_grad_enabled() and self.embedding.weight.requires_grad:
            return compute()
        key = self._cache_key()
        if self._rep_cache is None or self._rep_cache[0] != key:
            sharded = self._dim_shard is not None and (self.n_layers > 0 or hasattr(self, 'feat_mat'))
            self._rep_cache = (key, self._column_rep() if sharded else compute())
        return self._rep_cache[1]

    def load(self, path):
        super().load(path)
        self._bump()

    def predict(self, users):
        """LightGCN.predict (model.py:118-123): dense scores [len(users), n_items] from the igcn_predict_scores
        kernel.  The trainer's eval() does not come through here, it uses the fused score + mask + top-k kernels
        on the cached representation; gradients do not flow through predict (the reference only calls it under
        eval)."""
        from ._lib import call, ptr, stream_ptr
        rep = self.get_rep().detach().contiguous()
        users = torch.as_tensor(users, dtype=torch.int64, device=rep.device).contiguous()
        out = torch.empty((users.shape[0], self.n_items), dtype=torch.float32, device=rep.device)
        for lo in range(0, users.shape[0], 65535):
            chunk = users[lo:lo + 65535]
            call('igcn_predict_scores', ptr(rep), ptr(chunk), chunk.shape[0], self.n_users, self.n_items, rep.shape[1],
                 ptr(out[lo:lo + 65535]), stream_ptr())
        return out


class LightGCN(_GraphModel):
    """model.py:75-123."""

    def __init__(self, model_config):
        super().__init__(model_config)
        self.embedding_size = model_config['embedding_size']
        self.n_layers = model_config['n_layers']
        self.embedding = nn.Embedding(self.n_users + self.n_items, self.embedding_size)
        self.norm_adj = self.generate_graph(model_config['dataset'])
        normal_(self.embedding.weight, std=0.1)
        self.to(device=self.device)

    def generate_graph(self, dataset):
        """D^-1/2 A D^-1/2 as a device CSR (model.py:85-94)."""
        dg = self._device_graph_of(dataset)
        if dg is not None:
            return graph.NormAdj.from_device(dg, shard=self._shard_arg(dataset))
        return graph.NormAdj(dataset.n_users, dataset.n_items, graph.train_pairs_of(dataset), self.device,
                             shard=self._shard_arg(dataset))

    def get_rep(self):
        self._check_graph()
        self._sync_params()
        require_cuda(self.embedding.weight, torch.float32, 'embedding.weight')
        return self._cached_rep(lambda: engine.LightGCNRep.apply(self.embedding.weight, self))

    def bpr_forward(self, users, pos_items, neg_items):
        """model.py:108-116: the L2 term is over RAW embedding rows."""
        rep = self.get_rep()
        users_e = self.embedding(users)
        pos_e, neg_e = self.embedding(self.n_users + pos_items), self.embedding(self.n_users + neg_items)
        l2_norm_sq = (users_e ** 2).sum(dim=1) + (pos_e ** 2).sum(dim=1) + (neg_e ** 2).sum(dim=1)
        return rep[users, :], rep[self.n_users + pos_items, :], rep[self.n_users + neg_items, :], l2_norm_sq


class MF(LightGCN):
    """Matrix factorisation (model.py:52-72) as the zero-layer member of the family: get_rep is the embedding table
    itself, bpr_forward is LightGCN's (raw rows, L2 over raw rows -- exactly MF.bpr_forward), so the fused BPR step and
    the fused ranking kernels run it unchanged.  The reference keeps two tables; here they are the user and the item
    half of ONE [U + I, D] table (what the kernels index), created with the reference's draw order (user table first)
    and exposed under the reference's names: `user_embedding` / `item_embedding` (read-only views sharing storage with
    `embedding`), and state_dict / load_state_dict use the reference's keys, so checkpoints are interchangeable."""

    def __init__(self, model_config):
        cfg = dict(model_config, n_layers=0)
        _GraphModel.__init__(self, cfg)
        self.config = model_config
        self.embedding_size = model_config['embedding_size']
        self.n_layers = 0
        # same generator consumption as the reference constructor: both nn.Embedding inits, then both normal_ calls
        user = nn.Embedding(self.n_users, self.embedding_size)
        item = nn.Embedding(self.n_items, self.embedding_size)
        normal_(user.weight, std=0.1)
        normal_(item.weight, std=0.1)
        self.embedding = nn.Embedding(self.n_users + self.n_items, self.embedding_size)
        with torch.no_grad():
            self.embedding.weight.copy_(torch.cat([user.weight, item.weight], dim=0))
        self.norm_adj = self.generate_graph(model_config['dataset'])     # sampler CSR + touched-row plan of the step
        self.to(device=self.device)

    class _Half:
        """`model.user_embedding.weight` / `model.user_embedding(ids)` of the reference, as a view of one half."""

        def __init__(self, model, lo, hi):
            self._m, self._lo, self._hi = model, lo, hi

        @property
        def weight(self):
            return self._m.embedding.weight[self._lo:self._hi]

        def __call__(self, ids):
            return self._m.embedding(ids + self._lo)

    @property
    def user_embedding(self):
        return MF._Half(self, 0, self.n_users)

    @property
    def item_embedding(self):
        return MF._Half(self, self.n_users, self.n_users + self.n_items)

    def bpr_forward(self, users, pos_items, neg_items):
        """model.py:62-67."""
        self._sync_params()
        users_e = self.embedding(users)
        pos_e, neg_e = self.embedding(self.n_users + pos_items), self.embedding(self.n_users + neg_items)
        l2_norm_sq = (users_e ** 2).sum(dim=1) + (pos_e ** 2).sum(dim=1) + (neg_e ** 2).sum(dim=1)
        return users_e, pos_e, neg_e, l2_norm_sq

    def state_dict(self, *args, **kwargs):
        self._sync_params()
        w = self.embedding.weight.detach()
        return {'user_embedding.weight': w[:self.n_users].clone(), 'item_embedding.weight': w[self.n_users:].clone()}

    def load_state_dict(self, state, strict=True):
        with torch.no_grad():
            self.embedding.weight[:self.n_users].copy_(state['user_embedding.weight'])
            self.embedding.weight[self.n_users:].copy_(state['item_embedding.weight'])
        self._bump()


class IdentityMap:
    """user_map / item_map of feature_ratio == 1 (model.py:392-401) without materialising n dict entries."""

    def __init__(self, n):
        self.n = int(n)

    def __len__(self):
        return self.n

    def __contains__(self, k):
        return 0 <= k < self.n

    def __getitem__(self, k):
        if not 0 <= k < self.n:
            raise KeyError(k)
        return k

    def keys(self):
        return range(self.n)

    values = keys

    def items(self):
        return ((k, k) for k in range(self.n))


class Popularity(BasicModel):
    """Item-popularity ranker (model.py:338-351; run/dropui/igcn_dropui.py:43-48 evaluates it next to IGCN).
    Expressed as a 4-wide representation -- user rows (1, 0, 0, 0), item rows (degree, 0, 0, 0) -- so that the
    fused score + mask + top-k kernels rank it like any other model: score(u, i) = degree(i)."""

    def __init__(self, model_config):
        super().__init__(model_config)
        self.item_degree = self.calculate_degree(model_config['dataset'])
        self.trainable = False

    def calculate_degree(self, dataset):
        pairs = graph.train_pairs_of(dataset)
        deg = np.bincount(pairs[:, 1], minlength=self.n_items).astype(np.float32)
        return torch.tensor(deg, dtype=torch.float32, device=self.device)

    def get_rep(self):
        rep = torch.zeros((self.n_users + self.n_items, 4), dtype=torch.float32, device=self.device)
        rep[:self.n_users, 0] = 1.
        rep[self.n_users:, 0] = self.item_degree
        return rep

    def predict(self, users):
        return self.item_degree[None, :].repeat(users.shape[0], 1)


def graph_rank_nodes(dataset, ranking_metric):
    """Template ("core") node ranking for feature_ratio < 1 (utils.py:94-123).  Host-side, one-off."""
    adj = graph.build_adjacency(dataset.n_users, dataset.n_items, graph.train_pairs_of(dataset))
    n_u = dataset.n_users
    if ranking_metric == 'degree':
        deg = np.asarray(adj.sum(axis=1)).squeeze()
        user_metrics, item_metrics = deg[:n_u], deg[n_u:]
    elif ranking_metric in ('greedy', 'sort'):
        from sklearn.preprocessing import normalize
        walk = normalize(adj, axis=1, norm='l1')
        user_metrics = np.asarray(walk[:, :n_u].sum(axis=0)).squeeze()
        item_metrics = np.asarray(walk[:, n_u:].sum(axis=0)).squeeze()
    elif ranking_metric == 'page_rank':
        import networkx as nx
        g = nx.Graph()
        g.add_edges_from(np.array(adj.nonzero()).T)
        pr = nx.pagerank(g)
        pr = np.array([pr[i] for i in range(adj.shape[0])])
        user_metrics, item_metrics = pr[:n_u], pr[n_u:]
    else:
        return None
    return np.argsort(user_metrics)[::-1].copy(), np.argsort(item_metrics)[::-1].copy()


class IGCN(_GraphModel):
    """INMO-LightGCN (model.py:354-466)."""

    def __init__(self, model_config):
        super().__init__(model_config)
        self.embedding_size = model_config['embedding_size']
        self.n_layers = model_config['n_layers']
        self.dropout = model_config['dropout']
        self.feature_ratio = model_config['feature_ratio']
        self.norm_adj = self.generate_graph(model_config['dataset'])
        self.alpha = 1.
        self.delta = model_config.get('delta', 0.99)
        self.feat_mat, self.user_map, self.item_map, self.row_sum = \
            self.generate_feat(model_config['dataset'], ranking_metric=model_config.get('ranking_metric', 'sort'))
        self.update_feat_mat()

        self.embedding = nn.Embedding(self.feat_mat.shape[1], self.embedding_size)
        self.w = nn.Parameter(torch.ones([self.embedding_size], dtype=torch.float32, device=self.device))
        normal_(self.embedding.weight, std=0.1)
        self.to(device=self.device)
        self._drop_calls = 0
        self.drop_seed = int(model_config.get('dropout_seed', 0))
        self.injected_keep = None      # tests: keep vector in the reference's nnz order for the next get_rep
        self._aux = None

    # ---- graph / template construction
    def generate_graph(self, dataset):
        return LightGCN.generate_graph(self, dataset)

    def _maps_to_arrays(self, user_map, item_map):
        ut = np.full(self.n_users, -1, dtype=np.int64)
        it = np.full(self.n_items, -1, dtype=np.int64)
        for arr, mp in ((ut, user_map), (it, item_map)):
            if len(mp):
                keys = np.fromiter(mp.keys(), dtype=np.int64, count=len(mp))
                vals = np.fromiter(mp.values(), dtype=np.int64, count=len(mp))
                ok = keys < len(arr)
                arr[keys[ok]] = vals[ok]
        return ut, it

    def generate_feat(self, dataset, is_updating=False, ranking_metric=None):
        """Template incidence structure (model.py:386-421).  Returns (feat, user_map, item_map,
        row_sum) like the reference; `feat` is a graph.TemplateFeat."""
        dg = getattr(dataset, 'device_graph', None)
        if dg is not None:
            if is_updating or self.feature_ratio < 1.:
                raise RuntimeError('device-resident graphs support feature_ratio == 1 without template updates')
            adj = self.norm_adj if self.norm_adj.shape[0] == dg.n_users + dg.n_items else None
            feat = graph.TemplateFeat.from_device(dg, adj=adj, shard=self._shard_arg(dataset))
            self._aux = None
            return feat, IdentityMap(dg.n_users), IdentityMap(dg.n_items), feat.row_sum
        dg = self._device_graph_of(dataset)
        if not is_updating:
            if self.feature_ratio < 1.:
                ranked_users, ranked_items = graph_rank_nodes(dataset, ranking_metric)
                core_users = ranked_users[:int(self.n_users * self.feature_ratio)]
                core_items = ranked_items[:int(self.n_items * self.feature_ratio)]
            else:
                core_users = np.arange(self.n_users, dtype=np.int64)
                core_items = np.arange(self.n_items, dtype=np.int64)
            user_map = {int(u): t for t, u in enumerate(core_users.tolist())}
            item_map = {int(i): t for t, i in enumerate(core_items.tolist())}
        else:
            user_map, item_map = self.user_map, self.item_map
        ut, it = self._maps_to_arrays(user_map, item_map)
        if dg is not None:
            adj = getattr(self, 'norm_adj', None)
            if adj is None or getattr(adj, '_dg', None) is not dg:
                adj = None                         # only a NormAdj of the same build shares its index arrays
            feat = graph.TemplateFeat.from_device(dg, adj=adj, shard=self._shard_arg(dataset), user_tmpl=ut, item_tmpl=it,
                                                  t_users=len(user_map), t_items=len(item_map))
        else:
            feat = graph.TemplateFeat(self.n_users, self.n_items, graph.train_pairs_of(dataset), ut, it,
                                      len(user_map), len(item_map), self.device, shard=self._shard_arg(dataset))
        self._aux = None
        return feat, user_map, item_map, feat.row_sum

    def update_feat_mat(self):
        """Entries of row r become row_sum[r] ** ((alpha-1)/2 - 1/2) (model.py:374-377)."""
        self.feat_mat.set_alpha(self.alpha, self.row_sum)

    def feat_mat_anneal(self):
        """model.py:379-381."""
        self.alpha *= self.delta
        self.update_feat_mat()

    def aux_csr(self):
        """User-by-item train CSR in TEMPLATE id space for the auxiliary sampler (dataset.py:258-273)."""
        if self._aux is None:
            feat = self.feat_mat
            if feat.tmpl is None:
                rowptr, col = self.norm_adj.sampler_csr()
                self._aux = {'rowptr': rowptr, 'col': col, 'col_offset': self.n_users,
                             'n_users': self.n_users, 'n_items': self.n_items}
            else:
                pairs = graph.train_pairs_of(self.config['dataset'])
                ut, it = self._maps_to_arrays(self.user_map, self.item_map)
                tu, ti = ut[pairs[:, 0]], it[pairs[:, 1]]
                ok = (tu >= 0) & (ti >= 0)
                m = sp.csr_matrix((np.ones(int(ok.sum()), dtype=np.float32), (tu[ok], ti[ok])),
                                  shape=(feat.t_users, feat.t_items))
                m.sum_duplicates()
                m.sort_indices()
                self._aux = {'rowptr': torch.from_numpy(m.indptr.astype(np.int64)).to(self.device),
                             'col': torch.from_numpy(m.indices.astype(np.int32)).to(self.device),
                             'col_offset': 0, 'n_users': feat.t_users, 'n_items': feat.t_items}
        return self._aux

    # ---- forward
    def _next_drop(self):
        if not self.training:
            return None                                  # model.py:264-265
        feat = self.feat_mat
        if self.injected_keep is not None:
            if self._rows_sharded():
                raise RuntimeError('replaying an explicit dropout mask is a single-GPU test facility')
            ek, sk = feat.keep_bits(self.injected_keep)
            self.injected_keep = None
            return {'mode': 2, 'p': self.dropout, 'edge_keep': ek, 'self_keep': sk, 'tperm': feat.tperm()}
        if self.dropout <= 0.:
            return None
        self._drop_calls += 1
        seed = (self.drop_seed * 0x9e3779b97f4a7c15 + self._drop_calls * 0xd1342543de82ef95) % (1 << 64)
        return {'mode': 1, 'p': self.dropout, 'seed': seed}

    def inductive_rep_layer(self, feat_mat):
        """feat_mat @ embedding.weight (model.py:423-432); differentiable through get_rep only."""
        prop = self._propagator()
        x0 = torch.empty((feat_mat.shape[0], self.embedding_size), dtype=torch.float32, device=self.device)
        if prop.shard is not None:
            x0 = prop.x0_buffer()
        engine.inmo_forward(feat_mat, self.embedding.weight.detach().contiguous(), x0, None, prop.dim, prop.shard)
        return x0 if prop.shard is None else x0.clone()

    def get_rep(self):
        """model.py:434-446."""
        self._check_graph()
        self._sync_params()
        feat = self.feat_mat
        if feat.shape[0] != self.n_users + self.n_items or feat.shape[1] != self.embedding.weight.shape[0]:
            raise RuntimeError('feat_mat %s does not match nodes=%d / templates=%d'
                               % (tuple(feat.shape), self.n_users + self.n_items, self.embedding.weight.shape[0]))
        require_cuda(self.embedding.weight, torch.float32, 'embedding.weight')
        drop = self._next_drop()
        return self._cached_rep(lambda: engine.IGCNRep.apply(self.embedding.weight, self, drop))

    def bpr_forward(self, users, pos_items, neg_items):
        """NGCF.bpr_forward (model.py:293-299): the L2 term is over PROPAGATED rows."""
        rep = self.get_rep()
        users_r = rep[users, :]
        pos_r, neg_r = rep[self.n_users + pos_items, :], rep[self.n_users + neg_items, :]
        l2_norm_sq = (users_r ** 2).sum(dim=1) + (pos_r ** 2).sum(dim=1) + (neg_r ** 2).sum(dim=1)
        return users_r, pos_r, neg_r, l2_norm_sq

    # ---- checkpoint (model.py:454-466; key spelling kept for file compatibility)
    def save(self, path):
        self._sync_params()
        params = {'sate_dict': self.state_dict(), 'user_map': self.user_map,
                  'item_map': self.item_map, 'alpha': self.alpha}
        torch.save(params, path)

    def load(self, path):
        params = torch.load(path, map_location=self.device, weights_only=False)
        self.load_state_dict(params['sate_dict'])
        self.user_map = params['user_map']
        self.item_map = params['item_map']
        self.alpha = params['alpha']
        self.feat_mat, _, _, self.row_sum = self.generate_feat(self.config['dataset'], is_updating=True)
        self.update_feat_mat()
        self._bump()


class IMF(IGCN):
    """INMO-MF (model.py:536-543): the template layer alone, no propagation -- IGCN with zero layers, whatever
    n_layers the config carries (config.py:44 passes 0).  Runs on the same fused step and ranking kernels."""

    def __init__(self, model_config):
        super().__init__(model_config)
        self.n_layers = 0


from .siblings import IMCGAE, NGCF  # noqa: E402,F401  (get_model dispatches on this module's namespace; siblings imports _GraphModel)
