"""GPU parity tests: the CUDA path (through the C ABI) against the golden vectors produced by the
reference itself, and against the CPU oracle (oracle/restate.py) on seeded inputs.

Tolerances (north_star): embeddings / loss 1e-5 relative to scale (fp32); top-k lists identical
wherever the oracle's scores are not tied; metrics identical.
"""
import numpy as np
import pytest
import torch

from conftest import check_metrics, check_topk_lists, load_golden, rel_err

pytestmark = pytest.mark.gpu
TOL = 1e-5
DEV = torch.device('cuda:0')


def _dataset(tiny, **over):
    from igcn_cf_b200.dataset import get_dataset
    cfg = {'name': 'ListDataset', 'train': tiny['train'], 'val': tiny['val'], 'test': tiny['test'],
           'n_items': tiny['n_items'], 'device': DEV}
    cfg.update(over)
    return get_dataset(cfg)


def _lgcn(tiny, g, l2_reg=1e-4, **tr):
    from igcn_cf_b200.model import get_model
    from igcn_cf_b200.trainer import get_trainer
    ds = _dataset(tiny)
    model = get_model({'name': 'LightGCN', 'embedding_size': 64, 'n_layers': 3, 'device': DEV}, ds)
    with torch.no_grad():
        model.embedding.weight.copy_(torch.from_numpy(g['emb0']))
    cfg = {'name': 'BPRTrainer', 'optimizer': 'Adam', 'lr': 1e-3, 'l2_reg': l2_reg, 'device': DEV, 'n_epochs': 1,
           'batch_size': 2048, 'dataloader_num_workers': 0, 'test_batch_size': 512, 'topks': [5, 20]}
    cfg.update(tr)
    return ds, model, get_trainer(cfg, ds, model)


def _igcn(tiny, g, l2_reg=0., ratio=1., ds=None, **tr):
    from igcn_cf_b200.model import get_model
    from igcn_cf_b200.trainer import get_trainer
    ds = _dataset(tiny) if ds is None else ds
    model = get_model({'name': 'IGCN', 'embedding_size': 64, 'n_layers': 3, 'device': DEV, 'dropout': 0.3,
                       'feature_ratio': ratio}, ds)
    with torch.no_grad():
        model.embedding.weight.copy_(torch.from_numpy(g['emb0']))
    cfg = {'name': 'IGCNTrainer', 'optimizer': 'Adam', 'lr': 1e-3, 'l2_reg': l2_reg, 'aux_reg': 0.01, 'device': DEV,
           'n_epochs': 1, 'batch_size': 2048, 'dataloader_num_workers': 0, 'test_batch_size': 512, 'topks': [5, 20]}
    cfg.update(tr)
    return ds, model, get_trainer(cfg, ds, model)


def _keep(rand, p=0.3):
    """The reference's mask: floor(1 - p + rand) as bool (model.py:266-268), in fp32 like torch."""
    return np.floor(np.float32(1 - p) + rand.astype(np.float32)).astype(bool)


def _dev(a, dtype=torch.int64):
    return torch.as_tensor(np.ascontiguousarray(a), dtype=dtype, device=DEV)


def _check_evals(trainer, g, prefix, rep_key):
    """eval('train'|'val'|'test') against the reference's lists and metrics: lists identical except where the
    reference's own scores (recomputed from ITS representation g[rep_key]) are tied; metrics always asserted."""
    ds, model = trainer.dataset, trainer.model
    model.eval()
    with torch.no_grad():
        rep_mine = model.get_rep().cpu().numpy()
    for which in ('train', 'val', 'test'):
        _, metrics = trainer.eval(which)
        rec, _ = trainer.recommend(which)
        n_diff = check_topk_lists(rec.cpu().numpy(), g['%s_%s_rec' % (prefix, which)], g[rep_key], ds.n_users,
                                  rep_mine=rep_mine, scale_tol=TOL)
        n_valid = sum(1 for x in getattr(ds, which + '_data') if len(x) > 0)
        check_metrics(metrics, [(name, k, g['%s_%s_%s@%d' % (prefix, which, name, k)])
                                for name in ('Precision', 'Recall', 'NDCG') for k in (5, 20)], n_diff, n_valid)


# ------------------------------------------------------------------------------ LightGCN
def test_lightgcn_rep_and_scores(tiny):
    g = load_golden('tiny_lightgcn')
    _, model, _ = _lgcn(tiny, g)
    model.eval()
    with torch.no_grad():
        rep = model.get_rep()
        scores = model.predict(_dev(g['scores0_users']))
    assert rel_err(rep.cpu().numpy(), g['rep0']) < TOL
    assert rel_err(scores.cpu().numpy(), g['scores0']) < TOL


def test_lightgcn_autograd_path(tiny):
    """Generic API: bpr_forward + torch loss + backward through the custom autograd node."""
    g = load_golden('tiny_lightgcn')
    _, model, _ = _lgcn(tiny, g)
    model.train()
    t = _dev(g['fb_triples'])
    u_r, p_r, n_r, l2 = model.bpr_forward(t[:, 0], t[:, 1], t[:, 2])
    loss = torch.nn.functional.softplus((u_r * n_r).sum(1) - (u_r * p_r).sum(1)).mean() + 1e-4 * l2.mean()
    loss.backward()
    assert rel_err(u_r.detach().cpu().numpy(), g['fb_users_r']) < TOL
    assert rel_err(l2.detach().cpu().numpy(), g['fb_l2']) < TOL
    assert abs(loss.item() - float(g['fb_loss'])) < TOL
    assert rel_err(model.embedding.weight.grad.cpu().numpy(), g['fb_grad_emb']) < TOL


def test_lightgcn_fused_gradient(tiny):
    """Fused step internals: loss and dE of one step against the reference's autograd."""
    g = load_golden('tiny_lightgcn')
    _, model, trainer = _lgcn(tiny, g, cuda_graph=False)
    model.train()
    step = trainer.step
    step.run(_dev(g['fb_triples']))
    assert abs(step.loss.item() - float(g['fb_loss'])) < TOL
    assert rel_err(step.d_emb.cpu().numpy(), g['fb_grad_emb']) < TOL


@pytest.mark.parametrize('graph_mode', [False, True])
def test_lightgcn_epoch_and_eval(tiny, graph_mode):
    g = load_golden('tiny_lightgcn')
    _, model, trainer = _lgcn(tiny, g, cuda_graph=graph_mode)
    model.train()
    tri = g['epoch_triples']
    trainer.step.reset_meter()
    for lo in range(0, len(tri), 2048):
        trainer.step.run(_dev(tri[lo:lo + 2048]))
    assert abs(trainer.step.meter_avg() - float(g['epoch_loss'])) < TOL
    assert rel_err(model.embedding.weight.detach().cpu().numpy(), g['emb1']) < TOL
    model.eval()
    with torch.no_grad():
        assert rel_err(model.get_rep().cpu().numpy(), g['rep1']) < TOL
    _check_evals(trainer, g, 'e1', 'rep1')


# ------------------------------------------------------------------------------ IGCN
def test_igcn_structures_on_device(tiny):
    g = load_golden('tiny_igcn')
    _, model, _ = _igcn(tiny, g)
    assert np.array_equal(model.norm_adj.indices().cpu().numpy(), g['adj_idx'])
    assert np.array_equal(model.norm_adj.values().cpu().numpy(), g['adj_val'])
    assert np.array_equal(model.feat_mat.indices().cpu().numpy(), g['feat_idx'])
    assert rel_err(model.feat_mat.values().cpu().numpy(), g['feat_val']) < 1e-6
    assert np.array_equal(model.row_sum.cpu().numpy(), g['row_sum'])
    assert model.embedding.weight.shape == (702, 64)


def test_igcn_rep_eval_and_train(tiny):
    g = load_golden('tiny_igcn')
    _, model, _ = _igcn(tiny, g)
    model.eval()
    with torch.no_grad():
        assert rel_err(model.get_rep().cpu().numpy(), g['rep0_eval']) < TOL
    model.train()
    model.injected_keep = _keep(g['rep0_train_rand'])
    with torch.no_grad():
        assert rel_err(model.get_rep().cpu().numpy(), g['rep0_train']) < TOL


@pytest.mark.parametrize('tag,l2_reg', [('', 0.), ('_l2', 1e-3)])
def test_igcn_autograd_path(tiny, tag, l2_reg):
    g = load_golden('tiny_igcn')
    _, model, _ = _igcn(tiny, g)
    model.train()
    model.injected_keep = _keep(g['fb_rand'])
    t, a = _dev(g['fb_triples']), _dev(g['fb_aux_triples'])
    sp = torch.nn.functional.softplus
    u_r, p_r, n_r, l2 = model.bpr_forward(t[:, 0], t[:, 1], t[:, 2])
    bpr = sp((u_r * n_r).sum(1) - (u_r * p_r).sum(1)).mean()
    t_u = len(model.user_map)
    au, ap, an = model.embedding(a[:, 0]), model.embedding(a[:, 1] + t_u), model.embedding(a[:, 2] + t_u)
    aux = sp((au * an * model.w[None, :]).sum(1) - (au * ap * model.w[None, :]).sum(1)).mean()
    loss = bpr + (l2_reg * l2.mean() + 0.01 * aux)
    loss.backward()
    assert rel_err(u_r.detach().cpu().numpy(), g['fb_users_r']) < TOL
    assert abs(loss.item() - float(g['fb_loss' + tag])) < TOL
    assert rel_err(model.embedding.weight.grad.cpu().numpy(), g['fb_grad_emb' + tag]) < TOL
    assert rel_err(model.w.grad.cpu().numpy(), g['fb_grad_w' + tag]) < TOL


@pytest.mark.parametrize('tag,l2_reg', [('', 0.), ('_l2', 1e-3)])
def test_igcn_fused_gradient(tiny, tag, l2_reg):
    g = load_golden('tiny_igcn')
    _, model, trainer = _igcn(tiny, g, l2_reg=l2_reg, cuda_graph=False)
    model.train()
    ek, sk = model.feat_mat.keep_bits(_keep(g['fb_rand']))
    drop = {'mode': 2, 'p': 0.3, 'edge_keep': ek, 'self_keep': sk}
    step = trainer.step
    step.run(_dev(g['fb_triples']), _dev(g['fb_aux_triples']), drop=drop)
    assert abs(step.loss.item() - float(g['fb_loss' + tag])) < TOL
    assert rel_err(step.d_emb.cpu().numpy(), g['fb_grad_emb' + tag]) < TOL
    assert rel_err(step.d_w.cpu().numpy(), g['fb_grad_w' + tag]) < TOL


def test_igcn_epoch_anneal_eval(tiny):
    g = load_golden('tiny_igcn')
    _, model, trainer = _igcn(tiny, g, cuda_graph=False)
    model.train()
    tri, atri = g['epoch_triples'], g['epoch_aux_triples']
    trainer.step.reset_meter()
    for s, lo in enumerate(range(0, len(tri), 2048)):
        ek, sk = model.feat_mat.keep_bits(_keep(g['epoch_rand_%d' % s]))
        trainer.step.run(_dev(tri[lo:lo + 2048]), _dev(atri[lo:lo + 2048]),
                         drop={'mode': 2, 'p': 0.3, 'edge_keep': ek, 'self_keep': sk})
    model.feat_mat_anneal()
    assert abs(trainer.step.meter_avg() - float(g['epoch_loss'])) < TOL
    assert rel_err(model.embedding.weight.detach().cpu().numpy(), g['emb1']) < TOL
    assert rel_err(model.w.detach().cpu().numpy(), g['w1']) < TOL
    assert model.alpha == float(g['alpha1'])
    assert rel_err(model.feat_mat.values().cpu().numpy(), g['feat_val1']) < 1e-6
    model.eval()
    with torch.no_grad():
        assert rel_err(model.get_rep().cpu().numpy(), g['rep1_eval']) < TOL
        assert rel_err(model.predict(_dev(g['scores1_users'])).cpu().numpy(), g['scores1']) < TOL
    _check_evals(trainer, g, 'e1', 'rep1_eval')


def test_igcn_feature_ratio(tiny):
    g = load_golden('tiny_igcn_ratio')
    _, model, _ = _igcn(tiny, g, ratio=0.5)
    um = np.full(tiny['n_users'], -1, dtype=np.int64)
    for k, v in model.user_map.items():
        um[k] = v
    im = np.full(tiny['n_items'], -1, dtype=np.int64)
    for k, v in model.item_map.items():
        im[k] = v
    assert np.array_equal(um, g['user_map']) and np.array_equal(im, g['item_map'])
    assert np.array_equal(model.feat_mat.indices().cpu().numpy(), g['feat_idx'])
    model.eval()
    with torch.no_grad():
        assert rel_err(model.get_rep().cpu().numpy(), g['rep0_eval']) < TOL
    model.train()
    model.injected_keep = _keep(g['rep0_train_rand'])
    with torch.no_grad():
        assert rel_err(model.get_rep().cpu().numpy(), g['rep0_train']) < TOL


def test_igcn_dropui_sequence(tiny):
    """run/dropui/igcn_dropui.py:17-35: train on the reduced split, re-aggregate on the full graph
    without touching parameters, then the six inductive_eval passes."""
    from igcn_cf_b200 import synth
    from igcn_cf_b200.dataset import get_dataset
    from igcn_cf_b200.trainer import get_trainer
    g = load_golden('tiny_igcn_dropui')
    full = synth.gen_named('tiny', seed=2021)
    ds_small = get_dataset({'name': 'SyntheticDataset', 'split': full, 'variant': 'dropui', 'device': DEV})
    ds_full = get_dataset({'name': 'SyntheticDataset', 'split': full, 'device': DEV})
    assert (ds_small.n_users, ds_small.n_items) == (int(g['n_old_users']), int(g['n_old_items']))
    _, model, trainer = _igcn(tiny, g, ds=ds_small, cuda_graph=False)
    model.train()
    tri, atri = g['epoch_triples'], g['epoch_aux_triples']
    for s, lo in enumerate(range(0, len(tri), 2048)):
        ek, sk = model.feat_mat.keep_bits(_keep(g['epoch_rand_%d' % s]))
        trainer.step.run(_dev(tri[lo:lo + 2048]), _dev(atri[lo:lo + 2048]),
                         drop={'mode': 2, 'p': 0.3, 'edge_keep': ek, 'self_keep': sk})
    model.feat_mat_anneal()
    assert rel_err(model.embedding.weight.detach().cpu().numpy(), g['emb1']) < TOL

    model.config['dataset'] = ds_full
    model.n_users, model.n_items = ds_full.n_users, ds_full.n_items
    model.norm_adj = model.generate_graph(ds_full)
    model.feat_mat, _, _, model.row_sum = model.generate_feat(ds_full, is_updating=True)
    model.update_feat_mat()
    assert list(model.feat_mat.shape) == g['feat_shape'].tolist()
    assert np.array_equal(model.feat_mat.indices().cpu().numpy(), g['feat_idx'])
    assert rel_err(model.feat_mat.values().cpu().numpy(), g['feat_val']) < 1e-6
    model.eval()
    with torch.no_grad():
        rep_mine = model.get_rep().cpu().numpy()
    assert rel_err(rep_mine, g['rep_full']) < TOL
    cfg = dict(trainer.config)
    cfg.pop('dataset'), cfg.pop('model')
    trainer = get_trainer(cfg, ds_full, model)
    seen = []
    orig = trainer.calculate_metrics

    def spy(eval_data, rec_items):
        res = orig(eval_data, rec_items)
        seen.append((np.array(rec_items), res, int((np.asarray(eval_data.lens) > 0).sum())))
        return res

    trainer.calculate_metrics = spy
    trainer.inductive_eval(ds_small.n_users, ds_small.n_items)
    assert len(seen) == int(g['n_ind']) == 6
    for c, (rec, res, n_valid) in enumerate(seen):
        # lists identical except at proven ties of the reference's own scores; metrics always asserted
        n_diff = check_topk_lists(rec, g['ind%d_rec' % c], g['rep_full'], ds_full.n_users, rep_mine=rep_mine, scale_tol=TOL)
        check_metrics(res, [(name, k, g['ind%d_%s@%d' % (c, name, k)]) for name in res for k in res[name]], n_diff, n_valid)


# ------------------------------------------------------------------------------ properties
def test_hash_dropout_is_adjoint_and_reproducible(tiny):
    """Production dropout (mode 1): forward and transposed backward regenerate the same mask:
    <F~ E, G> == <E, F~^T G>; the keep rate matches 1-p; same seed -> same bits."""
    from igcn_cf_b200 import engine
    g = load_golden('tiny_igcn')
    _, model, _ = _igcn(tiny, g)
    feat, D = model.feat_mat, 64
    n, t = feat.shape
    gen = torch.Generator(device='cpu').manual_seed(3)
    E = torch.randn(t, D, generator=gen).to(DEV)
    G = torch.randn(n, D, generator=gen).to(DEV)
    drop = {'mode': 1, 'p': 0.3, 'seed': 12345}
    x0 = torch.empty(n, D, device=DEV)
    engine.inmo_forward(feat, E, x0, drop, D)
    x0b = torch.empty_like(x0)
    engine.inmo_forward(feat, E, x0b, drop, D)
    assert torch.equal(x0, x0b)
    scaled = (G * (feat.rowscale / 0.7)[:, None]).contiguous()
    dE = torch.zeros(t, D, device=DEV)
    engine.inmo_backward(feat, scaled, dE, drop, D, engine.colsum_scratch(n, D, DEV))
    lhs = (x0.double() * G.double()).sum().item()
    rhs = (E.double() * dE.double()).sum().item()
    assert abs(lhs - rhs) < 1e-4 * max(1.0, abs(lhs))
    ones = torch.ones(t, D, device=DEV)
    feat_plain = torch.empty(n, D, device=DEV)
    engine.inmo_forward(feat, ones, feat_plain, None, D)
    engine.inmo_forward(feat, ones, x0, drop, D)
    kept = (x0[:, 0] / feat.rowscale * 0.7).sum().item()
    total = (feat_plain[:, 0] / feat.rowscale).sum().item()
    assert abs(kept / total - 0.7) < 0.02


def test_long_row_chunks_match_unsplit(tiny):
    """Rows cut into chunks (power-law heads) give the same sums as whole rows, deterministically."""
    from igcn_cf_b200 import engine, graph
    g = load_golden('tiny_lightgcn')
    whole = graph.NormAdj(tiny['n_users'], tiny['n_items'], tiny['pairs'], DEV)
    old = graph.LONG_THRESHOLD, graph.CHUNK
    graph.LONG_THRESHOLD, graph.CHUNK = 16, 8
    try:
        # CsrDevice reads the module constants through its defaults at call time
        cut = graph.NormAdj(tiny['n_users'], tiny['n_items'], tiny['pairs'], DEV)
        cut.csr = graph.CsrDevice(cut.csr.rowptr_host, cut.csr.col_host, cut.csr.val.cpu().numpy(), 700, DEV,
                                  threshold=16, chunk=8)
    finally:
        graph.LONG_THRESHOLD, graph.CHUNK = old
    assert cut.csr.n_chunks > 50 and whole.csr.n_chunks == 0
    x = torch.from_numpy(g['emb0']).to(DEV)
    prop = engine.Propagator(700, 64, 3, DEV)
    a = prop.forward(whole, x, out=torch.empty_like(x)).clone()
    b = prop.forward(cut, x, out=torch.empty_like(x)).clone()
    c = prop.forward(cut, x, out=torch.empty_like(x)).clone()
    assert torch.equal(b, c)                                   # run-to-run deterministic
    assert rel_err(b.cpu().numpy(), a.cpu().numpy()) < 1e-6
    assert rel_err(a.cpu().numpy(), g['rep0']) < TOL


def test_device_sampler_is_valid_and_uniform(tiny):
    from igcn_cf_b200._lib import call, ptr, stream_ptr
    g = load_golden('tiny_lightgcn')
    _, model, _ = _lgcn(tiny, g)
    csr = model.norm_adj.csr
    B = 60000
    out = torch.empty((B, 3), dtype=torch.int64, device=DEV)
    call('igcn_sample_triples', ptr(csr.rowptr), ptr(csr.col), 300, 300, 400, B, 7, 0, None, ptr(out), stream_ptr())
    out2 = torch.empty_like(out)
    call('igcn_sample_triples', ptr(csr.rowptr), ptr(csr.col), 300, 300, 400, B, 7, 0, None, ptr(out2), stream_ptr())
    assert torch.equal(out, out2)
    call('igcn_sample_triples', ptr(csr.rowptr), ptr(csr.col), 300, 300, 400, B, 7, 1, None, ptr(out2), stream_ptr())
    assert not torch.equal(out, out2)
    t = out.cpu().numpy()
    train = [set(x) for x in tiny['train']]
    assert all(p in train[u] and n not in train[u] for u, p, n in t[:5000])
    counts = np.bincount(t[:, 0], minlength=300)
    assert counts.min() > 0 and abs(counts.std() / counts.mean() - np.sqrt(300 / B)) < 0.03   # uniform users
    u0 = t[t[:, 0] == 0][:, 1]
    pc = np.bincount(u0, minlength=400)[sorted(train[0])]
    assert pc.min() > 0 and pc.max() / pc.mean() < 2.5                                          # uniform positives


def test_fused_step_is_deterministic(tiny):
    g = load_golden('tiny_igcn')
    outs = []
    for _ in range(2):
        _, model, trainer = _igcn(tiny, g, cuda_graph=True, seed=5)
        model.train()
        for _ in range(4):
            trainer.step.run()
        outs.append((model.embedding.weight.detach().clone(), model.w.detach().clone(), trainer.step.meter_avg()))
    assert torch.equal(outs[0][0], outs[1][0]) and torch.equal(outs[0][1], outs[1][1]) and outs[0][2] == outs[1][2]
    assert not torch.equal(outs[0][0].cpu(), torch.from_numpy(g['emb0']))


def test_train_api_end_to_end(tiny, tmp_path, monkeypatch):
    """trainer.train(): epochs, validation, checkpoint save/load in the reference's format."""
    monkeypatch.chdir(tmp_path)
    g = load_golden('tiny_igcn')
    _, model, trainer = _igcn(tiny, g, n_epochs=3)
    best = trainer.train(verbose=False)
    assert best > 0.05 and trainer.save_path and trainer.save_path.startswith('checkpoints/IGCN_IGCNTrainer_')
    params = torch.load(trainer.save_path, weights_only=False)
    assert set(params) == {'sate_dict', 'user_map', 'item_map', 'alpha'}
    assert abs(model.alpha - params['alpha']) < 1e-12
    assert trainer.eval('val')[1]['NDCG'][5] == pytest.approx(best)      # best_ndcg keys on topks[0]


# ------------------------------------------------------------------------------ sibling models
def test_imf_epoch_and_eval(tiny):
    """IMF = the INMO layer without propagation (model.py:536-543) through the same fused step."""
    from igcn_cf_b200.model import get_model
    from igcn_cf_b200.trainer import get_trainer
    g = load_golden('tiny_siblings')
    ds = _dataset(tiny)
    model = get_model({'name': 'IMF', 'embedding_size': 64, 'n_layers': 0, 'device': DEV, 'dropout': 0.1,
                       'feature_ratio': 1.}, ds)
    with torch.no_grad():
        model.embedding.weight.copy_(torch.from_numpy(g['imf_emb0']))
    trainer = get_trainer({'name': 'IGCNTrainer', 'optimizer': 'Adam', 'lr': 1e-3, 'l2_reg': 1e-5, 'aux_reg': 0.1,
                           'device': DEV, 'n_epochs': 1, 'batch_size': 2048, 'dataloader_num_workers': 0,
                           'test_batch_size': 512, 'topks': [5, 20], 'cuda_graph': False}, ds, model)
    model.eval()
    with torch.no_grad():
        assert rel_err(model.get_rep().cpu().numpy(), g['imf_rep0_eval']) < TOL
    model.train()
    tri, atri = g['imf_epoch_triples'], g['imf_epoch_aux_triples']
    trainer.step.reset_meter()
    for s, lo in enumerate(range(0, len(tri), 2048)):
        ek, sk = model.feat_mat.keep_bits(_keep(g['imf_epoch_rand_%d' % s], p=0.1))
        trainer.step.run(_dev(tri[lo:lo + 2048]), _dev(atri[lo:lo + 2048]),
                         drop={'mode': 2, 'p': 0.1, 'edge_keep': ek, 'self_keep': sk})
    model.feat_mat_anneal()
    assert s + 1 == int(g['imf_epoch_n_steps'])
    assert abs(trainer.step.meter_avg() - float(g['imf_epoch_loss'])) < TOL
    assert rel_err(model.embedding.weight.detach().cpu().numpy(), g['imf_emb1']) < TOL
    assert rel_err(model.w.detach().cpu().numpy(), g['imf_w1']) < TOL
    model.eval()
    with torch.no_grad():
        assert rel_err(model.get_rep().cpu().numpy(), g['imf_rep1_eval']) < TOL
    _check_evals(trainer, g, 'imf_e1', 'imf_rep1_eval')


def test_mf_epoch_and_eval(tiny, tmp_path):
    """MF (model.py:52-72; config.py:6-10 hyper-parameters) on the fused step and ranking kernels: predict, one
    epoch on the reference's recorded triples, evals, and checkpoints with the reference's state_dict keys."""
    from igcn_cf_b200.model import get_model
    from igcn_cf_b200.trainer import get_trainer
    g = load_golden('tiny_mf')
    ds = _dataset(tiny)
    model = get_model({'name': 'MF', 'embedding_size': 64, 'device': DEV}, ds)
    with torch.no_grad():
        model.embedding.weight.copy_(torch.from_numpy(np.concatenate([g['mf_user0'], g['mf_item0']])))
    trainer = get_trainer({'name': 'BPRTrainer', 'optimizer': 'Adam', 'lr': 1e-4, 'l2_reg': 1e-3, 'device': DEV, 'n_epochs': 1,
                           'batch_size': 2048, 'dataloader_num_workers': 0, 'test_batch_size': 512, 'topks': [5, 20],
                           'cuda_graph': False}, ds, model)
    assert model.user_embedding.weight.shape == (ds.n_users, 64) and model.item_embedding.weight.shape == (ds.n_items, 64)
    model.eval()
    with torch.no_grad():
        scores = model.predict(_dev(g['mf_scores0_users']))
    assert rel_err(scores.cpu().numpy(), g['mf_scores0']) < TOL
    model.train()
    tri = g['mf_epoch_triples']
    trainer.step.reset_meter()
    for lo in range(0, len(tri), 2048):
        trainer.step.run(_dev(tri[lo:lo + 2048]))
    assert abs(trainer.step.meter_avg() - float(g['mf_epoch_loss'])) < TOL
    sd = model.state_dict()
    assert sorted(sd) == ['item_embedding.weight', 'user_embedding.weight']            # the reference's keys
    assert rel_err(sd['user_embedding.weight'].cpu().numpy(), g['mf_user1']) < TOL
    assert rel_err(sd['item_embedding.weight'].cpu().numpy(), g['mf_item1']) < TOL
    g['mf_rep1'] = np.concatenate([g['mf_user1'], g['mf_item1']])
    _check_evals(trainer, g, 'mf_e1', 'mf_rep1')
    path = str(tmp_path / 'mf.pth')
    model.save(path)
    other = get_model({'name': 'MF', 'embedding_size': 64, 'device': DEV}, ds)
    other.load(path)
    assert torch.equal(other.embedding.weight, model.embedding.weight)


def test_mf_initialisation_follows_the_reference_draw_order(tiny):
    """Same torch seed -> same initial tables as the reference constructor (user table first, model.py:56-60)."""
    from igcn_cf_b200.model import get_model
    from igcn_cf_b200.utils import set_seed
    g = load_golden('tiny_mf')
    set_seed(2021)
    model = get_model({'name': 'MF', 'embedding_size': 64, 'device': DEV}, _dataset(tiny))
    assert np.array_equal(model.user_embedding.weight.detach().cpu().numpy(), g['mf_user0'])
    assert np.array_equal(model.item_embedding.weight.detach().cpu().numpy(), g['mf_item0'])


def test_popularity_ranking(tiny):
    """Popularity (model.py:338-351) through BasicTrainer, as run/dropui/igcn_dropui.py:43-48 uses it."""
    from igcn_cf_b200.model import get_model
    from igcn_cf_b200.trainer import get_trainer
    g = load_golden('tiny_siblings')
    ds = _dataset(tiny)
    model = get_model({'name': 'Popularity', 'device': DEV}, ds)
    assert np.array_equal(model.item_degree.cpu().numpy(), g['pop_item_degree']) and not model.trainable
    trainer = get_trainer({'name': 'BasicTrainer', 'device': DEV, 'n_epochs': 0, 'topks': [5, 20],
                           'test_batch_size': 512}, ds, model)
    assert abs(trainer.train(verbose=False) - float(g['pop_train_return'])) < 0.02      # ties: see below
    # degrees tie massively: the reference's torch.topk breaks ties its own way, so compare what is well defined --
    # the SCORES of the recommended items and every metric that only depends on them being the right set
    for which in ('train', 'val', 'test'):
        rec, scores = trainer.recommend(which)
        ref_items = g['pop_%s_rec' % which]
        ref_scores = g['pop_item_degree'][ref_items]
        assert np.array_equal(scores.cpu().numpy(), ref_scores)
