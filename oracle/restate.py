"""CPU restatement of the reference's hot path.  TEST INFRASTRUCTURE ONLY.

This file is the checker the GPU parity tests compare against on the GPU box, where
/root/reference does not exist.  It restates -- in plain torch / numpy on CPU, with the same
library calls the reference reaches (torch.sparse.mm, torch.mm, torch.topk, F.softplus,
torch.optim.Adam) -- what these reference functions compute:

  utils.py:32-49      generate_daj_mat / get_sparse_tensor      -> bipartite_adjacency, to_coalesced
  model.py:85-94      LightGCN.generate_graph                    -> normalized_adjacency
  model.py:386-421    IGCN.generate_feat                         -> template_incidence
  model.py:374-381    IGCN.update_feat_mat / feat_mat_anneal     -> feat_values
  model.py:263-275    NGCF.dropout_sp_mat                        -> dropout_sparse
  model.py:423-432    IGCN.inductive_rep_layer                   -> (feat @ E) in igcn_rep
  model.py:96-106     LightGCN.get_rep                           -> lightgcn_rep
  model.py:434-446    IGCN.get_rep                               -> igcn_rep
  model.py:108-116    LightGCN.bpr_forward                       -> lightgcn_bpr_forward
  model.py:293-299    NGCF.bpr_forward (used by IGCN)            -> igcn_bpr_forward
  trainer.py:231-248  BPRTrainer.train_one_epoch (one step)      -> OracleLightGCN.train_step
  trainer.py:294-320  IGCNTrainer.train_one_epoch (one step)     -> OracleIGCN.train_step
  model.py:118-123    LightGCN.predict                           -> predict_scores
  trainer.py:140-167  BasicTrainer.eval                          -> masked_topk / evaluate
  trainer.py:109-138  BasicTrainer.calculate_metrics             -> calculate_metrics
  dataset.py:119-131  BasicDataset.__getitem__                   -> sample_triples
  model.py:255-261    NGCF.generate_graph                        -> row_normalized_adjacency
  model.py:277-291    NGCF.get_rep                               -> ngcf_rep
  model.py:560-581    IMCGAE.get_rep                             -> imcgae_rep

It is pinned against outputs of the reference itself (tests/golden/*.npz, produced by
tests/golden/make_golden.py in the build container through oracle/ref_loader.py); see
tests/test_oracle_golden.py.  The product never imports it.
"""
import random

import numpy as np
import scipy.sparse as sp
import torch
import torch.nn.functional as F


# --------------------------------------------------------------------------- graph building
def bipartite_adjacency(n_users, n_items, train_pairs):
    """Symmetric (U+I)x(U+I) 0/1 adjacency, users first (utils.py:41-49)."""
    pairs = np.asarray(train_pairs, dtype=np.int64).reshape(-1, 2)
    u, i = pairs[:, 0], pairs[:, 1] + n_users
    rows = np.concatenate([u, i])
    cols = np.concatenate([i, u])
    n = n_users + n_items
    return sp.coo_matrix((np.ones(rows.shape[0]), (rows, cols)), shape=(n, n), dtype=np.float32).tocsr()


def to_coalesced(mat):
    """scipy matrix -> coalesced torch COO, int64 indices / fp32 values (utils.py:32-38)."""
    coo = mat.tocoo()
    idx = torch.from_numpy(np.stack([coo.row, coo.col]).astype(np.int64))
    val = torch.from_numpy(coo.data.astype(np.float32))
    return torch.sparse_coo_tensor(idx, val, coo.shape).coalesce()


def normalized_adjacency(n_users, n_items, train_pairs):
    """D^-1/2 A D^-1/2 with deg clamped to >= 1, all in fp32 (model.py:85-94)."""
    adj = bipartite_adjacency(n_users, n_items, train_pairs)
    deg = np.maximum(1., np.asarray(adj.sum(axis=1)).squeeze())
    d_inv = sp.diags(np.power(deg, -0.5), format='csr', dtype=np.float32)
    return to_coalesced(d_inv.dot(adj).dot(d_inv))


def identity_maps(n_users, n_items):
    """feature_ratio == 1: every user/item is a template (model.py:392-401)."""
    return {u: u for u in range(n_users)}, {i: i for i in range(n_items)}


def template_incidence(n_users, n_items, train_pairs, user_map, item_map):
    """0/1 template incidence F and its row sums (model.py:402-421).

    Row r of F lists the templates of r's neighbours plus one global template column
    (users: T_u+T_i, items: T_u+T_i+1)."""
    t_u, t_i = len(user_map), len(item_map)
    rows, cols = [], []
    for u, i in np.asarray(train_pairs, dtype=np.int64).reshape(-1, 2).tolist():
        if i in item_map:
            rows.append(u)
            cols.append(t_u + item_map[i])
        if u in user_map:
            rows.append(n_users + i)
            cols.append(user_map[u])
    rows.extend(range(n_users))
    cols.extend([t_u + t_i] * n_users)
    rows.extend(range(n_users, n_users + n_items))
    cols.extend([t_u + t_i + 1] * n_items)
    feat = sp.coo_matrix((np.ones(len(rows)), (np.array(rows), np.array(cols))),
                         shape=(n_users + n_items, t_u + t_i + 2), dtype=np.float32).tocsr()
    row_sum = torch.from_numpy(np.asarray(feat.sum(axis=1)).squeeze().astype(np.float32))
    return to_coalesced(feat), row_sum


def feat_values(feat, row_sum, alpha):
    """Every nnz of row r gets row_sum[r] ** ((alpha-1)/2 - 1/2) (model.py:374-377)."""
    rows = feat.indices()[0]
    vals = torch.pow(row_sum[rows], (alpha - 1.) / 2. - 0.5)
    return torch.sparse_coo_tensor(feat.indices(), vals, feat.shape).coalesce()


def dropout_sparse(mat, p, rand):
    """Edge dropout with an injected U[0,1) vector (model.py:263-275).

    keep_e = floor(1 - p + rand_e) as bool; survivors are divided by (1 - p)."""
    keep = torch.floor((1 - p) + rand).type(torch.bool)
    idx = mat.indices()[:, keep]
    val = mat.values()[keep] / (1. - p)
    return torch.sparse_coo_tensor(idx, val, mat.shape).coalesce()


def row_normalized_adjacency(n_users, n_items, train_pairs):
    """normalize(A + I, norm='l1', axis=1) as a coalesced torch COO (NGCF.generate_graph, model.py:255-261).  sp.eye is
    float64, so the division happens in float64 and the values are rounded to float32 at the very end."""
    adj = bipartite_adjacency(n_users, n_items, train_pairs)
    adj = (adj + sp.eye(adj.shape[0], format='csr')).tocsr()
    inv = 1.0 / np.asarray(adj.sum(axis=1)).reshape(-1)
    return to_coalesced(sp.diags(inv).dot(adj))


# --------------------------------------------------------------------------- sibling models
def ngcf_rep(adj, emb, gc, bi, p=0., edge_keep=None, dense_keep=None):
    """NGCF.get_rep (model.py:277-291).  adj: row_normalized_adjacency; gc / bi: lists of (weight, bias) of the two
    dense layers per hop.  Train mode: edge_keep (bool per non-zero, model.py:263-275) and dense_keep (one bool
    [N, size] mask per hop, the F.dropout draw of model.py:287) with rate p; eval mode: both None."""
    if edge_keep is not None:
        adj = torch.sparse_coo_tensor(adj.indices()[:, edge_keep], adj.values()[edge_keep] / (1. - p), adj.shape).coalesce()
    rep = emb
    hops = [rep]
    for l in range(len(gc)):
        m0 = torch.sparse.mm(adj, rep)
        m1 = rep * m0
        rep = F.leaky_relu(F.linear(m0, *gc[l]) + F.linear(m1, *bi[l]), negative_slope=0.2)
        if dense_keep is not None:
            rep = rep * dense_keep[l].to(rep.dtype) / (1. - p)
        hops.append(F.normalize(rep, p=2, dim=1))
    return torch.cat(hops, dim=1)


def imcgae_rep(norm_adj, emb, n_users, n_items, n_layers, p=0., node_keep=None):
    """IMCGAE.get_rep (model.py:560-581).  emb: [U + I + 3, D] (personal rows, then identical / general-user /
    general-item); node_keep: per hop a bool [U + I] mask drawn with rate p - 0.1 * hop (train mode) or None."""
    U, I = n_users, n_items
    ident, gen_u, gen_i = emb[U + I], emb[U + I + 1], emb[U + I + 2]
    u_rep = torch.cat([emb[:U], gen_u[None, :].expand(U, -1), ident[None, :].expand(U, -1)], dim=1)
    i_rep = torch.cat([emb[U:U + I], gen_i[None, :].expand(I, -1), ident[None, :].expand(I, -1)], dim=1)
    rep = torch.cat([u_rep, i_rep], dim=0)
    layers = [rep]
    for l in range(n_layers):
        if node_keep is not None:
            rep = rep * (node_keep[l].to(rep.dtype) / (1. - (p - 0.1 * l)))[:, None]
        rep = torch.sparse.mm(norm_adj, rep)
        layers.append(rep / float(l + 2))
    return torch.stack(layers, dim=0).sum(dim=0)


def rep_bpr_loss(rep, n_users, users, pos, neg, l2_reg):
    """BPRTrainer's loss on a propagated representation with NGCF.bpr_forward's L2 term (model.py:293-299;
    trainer.py:238-243)."""
    u, p_, n_ = rep[users], rep[n_users + pos], rep[n_users + neg]
    l2 = torch.norm(u, p=2, dim=1) ** 2 + torch.norm(p_, p=2, dim=1) ** 2 + torch.norm(n_, p=2, dim=1) ** 2
    return F.softplus((u * n_).sum(1) - (u * p_).sum(1)).mean() + l2_reg * l2.mean()


# --------------------------------------------------------------------------- propagation
def _propagate_mean(norm_adj, x0, n_layers):
    layers = [x0]
    x = x0
    for _ in range(n_layers):
        x = torch.sparse.mm(norm_adj, x)
        layers.append(x)
    return torch.stack(layers, dim=0).mean(dim=0)


def lightgcn_rep(norm_adj, emb, n_layers):
    """model.py:96-106."""
    return _propagate_mean(norm_adj, emb, n_layers)


def igcn_rep(norm_adj, feat, emb, n_layers):
    """model.py:434-446 with `feat` already dropped/rescaled (or the eval-mode matrix)."""
    x0 = torch.sparse.mm(feat, emb)
    return _propagate_mean(norm_adj, x0, n_layers)


def _sq_norm(x):
    return torch.norm(x, p=2, dim=1) ** 2


def lightgcn_bpr_forward(rep, emb, n_users, users, pos, neg):
    """model.py:108-116: L2 term over RAW embedding rows."""
    l2 = _sq_norm(emb[users]) + _sq_norm(emb[n_users + pos]) + _sq_norm(emb[n_users + neg])
    return rep[users, :], rep[n_users + pos, :], rep[n_users + neg, :], l2


def igcn_bpr_forward(rep, n_users, users, pos, neg):
    """model.py:293-299: L2 term over PROPAGATED rows."""
    u, p, n = rep[users, :], rep[n_users + pos, :], rep[n_users + neg, :]
    return u, p, n, _sq_norm(u) + _sq_norm(p) + _sq_norm(n)


def bpr_loss(users_r, pos_r, neg_r):
    """trainer.py:238-241."""
    pos_s = torch.sum(users_r * pos_r, dim=1)
    neg_s = torch.sum(users_r * neg_r, dim=1)
    return F.softplus(neg_s - pos_s).mean()


def aux_loss(emb, w, t_u, users, pos, neg):
    """trainer.py:304-311: BPR on raw template rows, weighted by w."""
    u, p, n = emb[users], emb[pos + t_u], emb[neg + t_u]
    pos_s = torch.sum(u * p * w[None, :], dim=1)
    neg_s = torch.sum(u * n * w[None, :], dim=1)
    return F.softplus(neg_s - pos_s).mean()


# --------------------------------------------------------------------------- stateful wrappers
class _OracleBase:
    def __init__(self, n_users, n_items, train_pairs, n_layers, emb_init, lr):
        self.n_users, self.n_items, self.n_layers = n_users, n_items, n_layers
        self.norm_adj = normalized_adjacency(n_users, n_items, train_pairs)
        self.emb = torch.nn.Parameter(torch.as_tensor(emb_init, dtype=torch.float32).clone())
        self.lr = lr
        self.opt = None

    def _make_opt(self, params):
        self.opt = torch.optim.Adam(params, lr=self.lr)

    def predict(self, users):
        """model.py:118-123."""
        with torch.no_grad():
            rep = self.get_rep(train=False)
            return torch.mm(rep[users, :], rep[self.n_users:, :].t())


class OracleLightGCN(_OracleBase):
    def __init__(self, n_users, n_items, train_pairs, n_layers, emb_init, lr=1e-3, l2_reg=1e-4):
        super().__init__(n_users, n_items, train_pairs, n_layers, emb_init, lr)
        self.l2_reg = l2_reg
        self._make_opt([self.emb])

    def get_rep(self, train=False):
        return lightgcn_rep(self.norm_adj, self.emb, self.n_layers)

    def loss(self, users, pos, neg):
        rep = self.get_rep()
        u, p, n, l2 = lightgcn_bpr_forward(rep, self.emb, self.n_users, users, pos, neg)
        return bpr_loss(u, p, n) + self.l2_reg * l2.mean()

    def train_step(self, users, pos, neg):
        """One iteration of trainer.py:233-247; returns loss.item()."""
        loss = self.loss(users, pos, neg)
        self.opt.zero_grad()
        loss.backward()
        self.opt.step()
        return loss.item()


class OracleIGCN(_OracleBase):
    def __init__(self, n_users, n_items, train_pairs, n_layers, emb_init, dropout, lr=1e-3,
                 l2_reg=0., aux_reg=0.01, user_map=None, item_map=None, delta=0.99):
        if user_map is None:
            user_map, item_map = identity_maps(n_users, n_items)
        self.user_map, self.item_map = user_map, item_map
        super().__init__(n_users, n_items, train_pairs, n_layers, emb_init, lr)
        self.dropout, self.l2_reg, self.aux_reg = dropout, l2_reg, aux_reg
        self.alpha, self.delta = 1., delta
        self.feat_pattern, self.row_sum = template_incidence(n_users, n_items, train_pairs,
                                                             user_map, item_map)
        self.feat = feat_values(self.feat_pattern, self.row_sum, self.alpha)
        self.w = torch.nn.Parameter(torch.ones(self.emb.shape[1], dtype=torch.float32))
        self._make_opt([self.emb, self.w])   # parameter order of IGCN.__init__: embedding, w

    def anneal(self):
        """model.py:379-381."""
        self.alpha *= self.delta
        self.feat = feat_values(self.feat_pattern, self.row_sum, self.alpha)

    def regraph(self, n_users, n_items, train_pairs):
        """The inductive update of run/dropui/igcn_dropui.py:28-32 (templates unchanged)."""
        self.n_users, self.n_items = n_users, n_items
        self.norm_adj = normalized_adjacency(n_users, n_items, train_pairs)
        self.feat_pattern, self.row_sum = template_incidence(n_users, n_items, train_pairs,
                                                             self.user_map, self.item_map)
        self.feat = feat_values(self.feat_pattern, self.row_sum, self.alpha)

    def get_rep(self, train=False, rand=None):
        feat = self.feat
        if train:
            if rand is None:
                rand = torch.rand(feat._nnz())
            feat = dropout_sparse(feat, self.dropout, rand)
        return igcn_rep(self.norm_adj, feat, self.emb, self.n_layers)

    def loss(self, users, pos, neg, a_users, a_pos, a_neg, rand=None, train=True):
        rep = self.get_rep(train=train, rand=rand)
        u, p, n, l2 = igcn_bpr_forward(rep, self.n_users, users, pos, neg)
        aux = aux_loss(self.emb, self.w, len(self.user_map), a_users, a_pos, a_neg)
        return bpr_loss(u, p, n) + (self.l2_reg * l2.mean() + self.aux_reg * aux)

    def train_step(self, users, pos, neg, a_users, a_pos, a_neg, rand=None):
        """One iteration of trainer.py:296-318; returns loss.item()."""
        loss = self.loss(users, pos, neg, a_users, a_pos, a_neg, rand=rand)
        self.opt.zero_grad()
        loss.backward()
        self.opt.step()
        return loss.item()


# --------------------------------------------------------------------------- evaluation
def masked_topk(scores, users, k, exclude_lists=None, banned_items=None):
    """trainer.py:149-164: -inf the excluded (user, item) pairs and banned columns, then topk."""
    scores = scores.clone()
    if exclude_lists is not None:
        rows, cols = [], []
        for r, u in enumerate(users):
            items = exclude_lists[u]
            rows.extend([r] * len(items))
            cols.extend(items)
        scores[rows, cols] = -np.inf
    if banned_items is not None:
        scores[:, banned_items] = -np.inf
    vals, items = torch.topk(scores, k=k)
    return vals.numpy(), items.numpy()


def calculate_metrics(eval_data, rec_items, topks):
    """trainer.py:109-138 with the same dtypes (fp32 hit matrix and log2 table, int32 lengths)."""
    results = {'Precision': {}, 'Recall': {}, 'NDCG': {}}
    hit = np.zeros_like(rec_items, dtype=np.float32)
    for u in range(rec_items.shape[0]):
        truth = set(eval_data[u])
        if truth:
            hit[u] = [1. if it in truth else 0. for it in rec_items[u]]
    lens = np.array([len(items) for items in eval_data], dtype=np.int32)
    for k in topks:
        hit_num = np.sum(hit[:, :k], axis=1)
        precisions = hit_num / k
        with np.errstate(invalid='ignore', divide='ignore'):
            recalls = hit_num / lens
        max_hit = np.minimum(lens, k)
        ideal = (np.arange(k)[None, :] < max_hit[:, None]).astype(np.float32)
        denom = np.log2(np.arange(2, k + 2, dtype=np.float32))[None, :]
        dcgs = np.sum(hit[:, :k] / denom, axis=1)
        idcgs = np.sum(ideal / denom, axis=1)
        with np.errstate(invalid='ignore', divide='ignore'):
            ndcgs = dcgs / idcgs
        mask = max_hit > 0
        results['Precision'][k] = precisions[mask].mean()
        results['Recall'][k] = recalls[mask].mean()
        results['NDCG'][k] = ndcgs[mask].mean()
    return results


def evaluate(model, which, train_data, val_data, eval_data, topks, batch=512, banned_items=None):
    """trainer.py:140-167: batched predict -> mask -> topk -> metrics.  Returns (metrics, rec)."""
    k = max(topks)
    rec = []
    for lo in range(0, model.n_users, batch):
        users = list(range(lo, min(model.n_users, lo + batch)))
        scores = model.predict(torch.tensor(users, dtype=torch.int64))
        excl = None
        if which != 'train':
            excl = train_data if which == 'val' else [a + b for a, b in zip(train_data, val_data)]
        _, items = masked_topk(scores, users, k, excl, banned_items)
        rec.append(items)
    rec = np.concatenate(rec, axis=0)
    return calculate_metrics(eval_data, rec, topks), rec


# --------------------------------------------------------------------------- sampling
def sample_triples(train_data, n_users, n_items, count):
    """dataset.py:119-131 with neg_ratio 1: uniform user (retry if empty), uniform positive,
    rejection-sampled negative.  Uses the same global `random` / `np.random` streams."""
    out = np.empty((count, 3), dtype=np.int64)
    for t in range(count):
        user = random.randint(0, n_users - 1)
        while not train_data[user]:
            user = random.randint(0, n_users - 1)
        pos = np.random.choice(train_data[user])
        neg = random.randint(0, n_items - 1)
        while neg in train_data[user]:
            neg = random.randint(0, n_items - 1)
        out[t] = (user, pos, neg)
    return out
