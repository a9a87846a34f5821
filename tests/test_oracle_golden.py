"""Pin the CPU restatement (oracle/restate.py) to outputs of the reference itself.

The golden vectors were produced by tests/golden/make_golden.py running the unmodified
reference on CPU.  Tolerance 1e-6 relative-to-scale: both sides are fp32 torch CPU code.
"""
import numpy as np
import torch

from conftest import load_golden, rel_err
from oracle import restate as R

TOL = 1e-6


def _t(x):
    return torch.from_numpy(np.asarray(x))


def test_graph_matches_reference(tiny):
    g = load_golden('tiny_lightgcn')
    adj = R.normalized_adjacency(tiny['n_users'], tiny['n_items'], tiny['pairs'])
    assert np.array_equal(adj.indices().numpy(), g['adj_idx'])
    assert np.array_equal(adj.values().numpy(), g['adj_val'])          # bit-exact fp32


def test_lightgcn_rep_fwd_bwd(tiny):
    g = load_golden('tiny_lightgcn')
    m = R.OracleLightGCN(tiny['n_users'], tiny['n_items'], tiny['pairs'], 3, g['emb0'], l2_reg=1e-4)
    assert rel_err(m.get_rep().detach().numpy(), g['rep0']) < TOL
    assert rel_err(m.predict(_t(g['scores0_users'])).numpy(), g['scores0']) < TOL
    tri = _t(g['fb_triples'])
    loss = m.loss(tri[:, 0], tri[:, 1], tri[:, 2])
    loss.backward()
    assert abs(loss.item() - float(g['fb_loss'])) < TOL
    assert rel_err(m.emb.grad.numpy(), g['fb_grad_emb']) < TOL


def test_lightgcn_epoch_and_eval(tiny):
    g = load_golden('tiny_lightgcn')
    m = R.OracleLightGCN(tiny['n_users'], tiny['n_items'], tiny['pairs'], 3, g['emb0'], l2_reg=1e-4)
    tri = g['epoch_triples']
    tot, cnt = 0., 0
    for lo in range(0, len(tri), 2048):
        b = _t(tri[lo:lo + 2048])
        tot += m.train_step(b[:, 0], b[:, 1], b[:, 2]) * len(b)
        cnt += len(b)
    assert abs(tot / cnt - float(g['epoch_loss'])) < 1e-6
    assert rel_err(m.emb.detach().numpy(), g['emb1']) < 1e-5
    for which in ('train', 'val', 'test'):
        metrics, rec = R.evaluate(m, which, tiny['train'], tiny['val'], tiny[which], [5, 20])
        assert np.array_equal(rec, g['e1_%s_rec' % which])
        for name in ('Precision', 'Recall', 'NDCG'):
            for k in (5, 20):
                assert metrics[name][k] == g['e1_%s_%s@%d' % (which, name, k)]


def test_igcn_feat_matches_reference(tiny):
    g = load_golden('tiny_igcn')
    um, im = R.identity_maps(tiny['n_users'], tiny['n_items'])
    feat, row_sum = R.template_incidence(tiny['n_users'], tiny['n_items'], tiny['pairs'], um, im)
    feat = R.feat_values(feat, row_sum, 1.)
    assert list(feat.shape) == g['feat_shape'].tolist()
    assert np.array_equal(feat.indices().numpy(), g['feat_idx'])
    assert np.array_equal(feat.values().numpy(), g['feat_val'])
    assert np.array_equal(row_sum.numpy(), g['row_sum'])


def _igcn(tiny, g, **kw):
    return R.OracleIGCN(tiny['n_users'], tiny['n_items'], tiny['pairs'], 3, g['emb0'], 0.3, **kw)


def test_igcn_rep_eval_and_train(tiny):
    g = load_golden('tiny_igcn')
    m = _igcn(tiny, g)
    assert rel_err(m.get_rep(train=False).detach().numpy(), g['rep0_eval']) < TOL
    rep = m.get_rep(train=True, rand=_t(g['rep0_train_rand']))
    assert rel_err(rep.detach().numpy(), g['rep0_train']) < TOL


def test_igcn_fwd_bwd(tiny):
    g = load_golden('tiny_igcn')
    for tag, l2_reg in (('', 0.), ('_l2', 1e-3)):
        m = _igcn(tiny, g, l2_reg=l2_reg)
        t, a = _t(g['fb_triples']), _t(g['fb_aux_triples'])
        loss = m.loss(t[:, 0], t[:, 1], t[:, 2], a[:, 0], a[:, 1], a[:, 2], rand=_t(g['fb_rand']))
        loss.backward()
        assert abs(loss.item() - float(g['fb_loss' + tag])) < TOL
        assert rel_err(m.emb.grad.numpy(), g['fb_grad_emb' + tag]) < TOL
        assert rel_err(m.w.grad.numpy(), g['fb_grad_w' + tag]) < TOL


def test_igcn_epoch_anneal_eval(tiny):
    g = load_golden('tiny_igcn')
    m = _igcn(tiny, g)
    tri, atri = g['epoch_triples'], g['epoch_aux_triples']
    tot, cnt = 0., 0
    for s, lo in enumerate(range(0, len(tri), 2048)):
        b, a = _t(tri[lo:lo + 2048]), _t(atri[lo:lo + 2048])
        tot += m.train_step(b[:, 0], b[:, 1], b[:, 2], a[:, 0], a[:, 1], a[:, 2],
                            rand=_t(g['epoch_rand_%d' % s])) * len(b)
        cnt += len(b)
    m.anneal()
    assert s + 1 == int(g['epoch_n_steps'])
    assert abs(tot / cnt - float(g['epoch_loss'])) < 1e-6
    assert rel_err(m.emb.detach().numpy(), g['emb1']) < 1e-5
    assert rel_err(m.w.detach().numpy(), g['w1']) < 1e-5
    assert m.alpha == float(g['alpha1'])
    assert rel_err(m.feat.values().numpy(), g['feat_val1']) < TOL
    assert rel_err(m.get_rep().detach().numpy(), g['rep1_eval']) < 1e-5
    assert rel_err(m.predict(_t(g['scores1_users'])).numpy(), g['scores1']) < 1e-5
    for which in ('train', 'val', 'test'):
        metrics, rec = R.evaluate(m, which, tiny['train'], tiny['val'], tiny[which], [5, 20])
        assert np.array_equal(rec, g['e1_%s_rec' % which])
        for name in ('Precision', 'Recall', 'NDCG'):
            for k in (5, 20):
                assert metrics[name][k] == g['e1_%s_%s@%d' % (which, name, k)]


def test_igcn_feature_ratio(tiny):
    g = load_golden('tiny_igcn_ratio')
    um = {int(u): int(t) for u, t in enumerate(g['user_map']) if t >= 0}
    im = {int(i): int(t) for i, t in enumerate(g['item_map']) if t >= 0}
    m = R.OracleIGCN(tiny['n_users'], tiny['n_items'], tiny['pairs'], 3, g['emb0'], 0.3,
                     user_map=um, item_map=im)
    assert list(m.feat.shape) == g['feat_shape'].tolist()
    assert np.array_equal(m.feat.indices().numpy(), g['feat_idx'])
    assert np.array_equal(m.feat.values().numpy(), g['feat_val'])
    assert rel_err(m.get_rep().detach().numpy(), g['rep0_eval']) < TOL
    rep = m.get_rep(train=True, rand=_t(g['rep0_train_rand']))
    assert rel_err(rep.detach().numpy(), g['rep0_train']) < TOL


def test_metrics_edge_cases():
    # users with empty eval lists are excluded from the means (trainer.py:134-137)
    rec = np.array([[1, 2, 3], [4, 5, 6], [7, 8, 9]])
    res = R.calculate_metrics([[1, 9], [], [9]], rec, [1, 3])
    assert res['Precision'][1] == np.float32(0.5)
    assert abs(res['Recall'][3] - 0.75) < 1e-12
    assert res['NDCG'][3].dtype == np.float32


def test_imf_is_igcn_without_layers(tiny):
    """IMF (model.py:536-543) = the oracle's IGCN with n_layers = 0: rep, one recorded epoch, evals."""
    g = load_golden('tiny_siblings')
    m = R.OracleIGCN(tiny['n_users'], tiny['n_items'], tiny['pairs'], 0, g['imf_emb0'], 0.1, l2_reg=1e-5, aux_reg=0.1)
    assert rel_err(m.get_rep(train=False).detach().numpy(), g['imf_rep0_eval']) < TOL
    tri, atri = g['imf_epoch_triples'], g['imf_epoch_aux_triples']
    tot, cnt = 0., 0
    for s, lo in enumerate(range(0, len(tri), 2048)):
        b, a = _t(tri[lo:lo + 2048]), _t(atri[lo:lo + 2048])
        tot += m.train_step(b[:, 0], b[:, 1], b[:, 2], a[:, 0], a[:, 1], a[:, 2],
                            rand=_t(g['imf_epoch_rand_%d' % s])) * len(b)
        cnt += len(b)
    m.anneal()
    assert abs(tot / cnt - float(g['imf_epoch_loss'])) < 1e-6
    assert rel_err(m.emb.detach().numpy(), g['imf_emb1']) < 1e-5
    assert rel_err(m.get_rep().detach().numpy(), g['imf_rep1_eval']) < 1e-5
    for which in ('train', 'val', 'test'):
        metrics, rec = R.evaluate(m, which, tiny['train'], tiny['val'], tiny[which], [5, 20])
        assert np.array_equal(rec, g['imf_e1_%s_rec' % which])


def _unpack(g, key):
    shape = tuple(g[key + '_shape'])
    return _t(np.unpackbits(g[key])[:int(np.prod(shape))].reshape(shape).astype(bool))


def test_ngcf_restatement(tiny):
    """NGCF (model.py:232-299): graph bit-equal, eval representation, and the train-mode forward/backward with the
    reference's recorded edge and feature dropout draws -- loss and the gradient of every parameter."""
    g = load_golden('tiny_ngcf_imcgae')
    U, I = tiny['n_users'], tiny['n_items']
    adj = R.row_normalized_adjacency(U, I, tiny['pairs'])
    assert np.array_equal(adj.indices().numpy(), g['ngcf_adj_idx'])
    assert np.array_equal(adj.values().numpy(), g['ngcf_adj_val'])
    P = {k: _t(g['ngcf_p0_' + k]).clone().requires_grad_(True) for k in g['ngcf_param_names']}
    gc = [(P['gc_layers.%d.weight' % l], P['gc_layers.%d.bias' % l]) for l in range(3)]
    bi = [(P['bi_layers.%d.weight' % l], P['bi_layers.%d.bias' % l]) for l in range(3)]
    with torch.no_grad():
        rep = R.ngcf_rep(adj, P['embedding.weight'], gc, bi)
    assert rel_err(rep.numpy()[::5], g['ngcf_rep0_eval_every5']) < TOL
    assert rel_err((rep[:64] @ rep[U:].t()).numpy(), g['ngcf_scores0']) < TOL
    keep = torch.floor(np.float32(0.9) + _t(g['ngcf_fb_rand_0'])).bool()
    rep = R.ngcf_rep(adj, P['embedding.weight'], gc, bi, p=0.1, edge_keep=keep,
                     dense_keep=[_unpack(g, 'ngcf_fb_dense_%d' % l) for l in range(3)])
    tri = _t(g['ngcf_fb_triples'])
    loss = R.rep_bpr_loss(rep, U, tri[:, 0], tri[:, 1], tri[:, 2], 1e-3)
    loss.backward()
    assert abs(loss.item() - float(g['ngcf_fb_loss'])) < TOL
    for k, v in P.items():
        assert rel_err(v.grad.numpy(), g['ngcf_fb_grad_' + k]) < TOL, k


def test_imcgae_restatement(tiny):
    """IMCGAE (model.py:546-585): eval representation / scores and the train-mode forward/backward with the recorded
    node-dropout masks (rates 0.3, 0.2, 0.1)."""
    g = load_golden('tiny_ngcf_imcgae')
    U, I = tiny['n_users'], tiny['n_items']
    adj = R.normalized_adjacency(U, I, tiny['pairs'])
    emb = _t(g['imcgae_p0_embedding.weight']).clone().requires_grad_(True)
    with torch.no_grad():
        rep = R.imcgae_rep(adj, emb, U, I, 3)
    assert rep.shape == (U + I, 192)
    assert rel_err(rep.numpy()[::5], g['imcgae_rep0_eval_every5']) < TOL
    assert rel_err((rep[:64] @ rep[U:].t()).numpy(), g['imcgae_scores0']) < TOL
    rep = R.imcgae_rep(adj, emb, U, I, 3, p=0.3, node_keep=[_unpack(g, 'imcgae_fb_dense_%d' % l) for l in range(3)])
    tri = _t(g['imcgae_fb_triples'])
    loss = R.rep_bpr_loss(rep, U, tri[:, 0], tri[:, 1], tri[:, 2], 0.)
    loss.backward()
    assert abs(loss.item() - float(g['imcgae_fb_loss'])) < TOL
    assert rel_err(emb.grad.numpy(), g['imcgae_fb_grad_embedding.weight']) < TOL
