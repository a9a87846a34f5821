"""NGCF (reference model.py:232-299) and IMCGAE (model.py:546-585) -- SURVEY.md 8f rank 4, the sibling models whose
propagation is the reference's gspmm -- on igcn_spmm / the fused ranking kernel, against golden vectors produced by the
unmodified reference (tests/golden/make_golden.py wide).  Every random draw of the reference's train-mode passes was
recorded (torch.rand of dropout_sp_mat, the keep mask of each F.dropout call) and is replayed here through
`model.injected`, so train-mode losses, gradients and one whole epoch are compared number for number.

Tolerances: representations, scores, losses, gradients, weights 1e-5 relative to scale (fp32, different summation
orders); top-k lists identical wherever the reference's own scores are not tied; metrics identical."""
import numpy as np
import pytest
import torch

from conftest import load_golden, rel_err
from test_gpu_parity import DEV, TOL, _check_evals, _dataset, _dev, _keep

pytestmark = pytest.mark.gpu

CFG = {'ngcf': ({'name': 'NGCF', 'embedding_size': 64, 'layer_sizes': [64, 64, 64], 'device': DEV, 'dropout': 0.1}, 1e-3),
       'imcgae': ({'name': 'IMCGAE', 'embedding_size': 64, 'n_layers': 3, 'device': DEV, 'dropout': 0.3}, 0.)}


def _build(tiny, g, px, **tr):
    from igcn_cf_b200.model import get_model
    from igcn_cf_b200.trainer import get_trainer
    ds = _dataset(tiny)
    mcfg, l2 = CFG[px]
    model = get_model(mcfg, ds)
    names = [k for k, _ in model.named_parameters()]
    assert names == list(g[px + '_param_names'])                 # same modules, same state_dict keys as the reference
    with torch.no_grad():
        for k, p in model.named_parameters():
            p.copy_(torch.from_numpy(g['%s_p0_%s' % (px, k)]))
    model._bump()
    cfg = {'name': 'BPRTrainer', 'optimizer': 'Adam', 'lr': 1e-3, 'l2_reg': l2, 'device': DEV, 'n_epochs': 1,
           'batch_size': 2048, 'dataloader_num_workers': 0, 'test_batch_size': 512, 'topks': [5, 20]}
    cfg.update(tr)
    return ds, model, get_trainer(cfg, ds, model), l2


def _injected(g, px, key, first_rand, first_dense, n_dense, p_edge):
    """The recorded draws of one forward pass in the form `model.injected` takes."""
    inj = {'dense': []}
    if px == 'ngcf':
        inj['edge'] = torch.from_numpy(_keep(g['%s_%s_rand_%d' % (px, key, first_rand)], p_edge)).to(DEV)
    for s in range(first_dense, first_dense + n_dense):
        shape = tuple(g['%s_%s_dense_%d_shape' % (px, key, s)])
        bits = np.unpackbits(g['%s_%s_dense_%d' % (px, key, s)])[:int(np.prod(shape))].reshape(shape)
        inj['dense'].append(torch.from_numpy(bits.astype(bool)).to(DEV))
    return inj


@pytest.mark.parametrize('px', ['ngcf', 'imcgae'])
def test_rep_and_scores(tiny, px):
    g = load_golden('tiny_ngcf_imcgae')
    ds, model, _, _ = _build(tiny, g, px)
    model.eval()
    with torch.no_grad():
        rep = model.get_rep()
        scores = model.predict(_dev(np.arange(64)))
    assert rep.shape == (ds.n_users + ds.n_items, 256 if px == 'ngcf' else 192)
    assert rel_err(rep.cpu().numpy()[::5], g[px + '_rep0_eval_every5']) < TOL
    assert rel_err(scores.cpu().numpy(), g[px + '_scores0']) < TOL


def test_ngcf_graph_is_the_references(tiny):
    """normalize(A + I, 'l1') (model.py:255-261): indices identical, values bit-equal; the stored transpose is one."""
    g = load_golden('tiny_ngcf_imcgae')
    _, model, _, _ = _build(tiny, g, 'ngcf')
    adj = model.norm_adj
    assert tuple(adj.shape) == (700, 700) and adj._nnz() == g['ngcf_adj_idx'].shape[1]
    assert np.array_equal(adj.indices().cpu().numpy(), g['ngcf_adj_idx'])
    assert np.array_equal(adj.values().cpu().numpy(), g['ngcf_adj_val'])
    dense = torch.zeros(700, 700, device=DEV)
    r, c = adj.indices()
    dense[r, c] = adj.csr_fwd.val
    dense_t = torch.zeros(700, 700, device=DEV)
    dense_t[r, c] = adj.csr_bwd.val
    assert torch.equal(dense_t, dense.t())
    x = torch.randn(700, 64, device=DEV, generator=torch.Generator(device=DEV).manual_seed(3))
    from igcn_cf_b200.siblings import spmm_raw
    assert rel_err(spmm_raw(adj.csr_fwd, x).cpu().numpy(), (dense.double() @ x.double()).cpu().numpy()) < TOL
    assert rel_err(spmm_raw(adj.csr_bwd, x).cpu().numpy(), (dense.t().double() @ x.double()).cpu().numpy()) < TOL


@pytest.mark.parametrize('px', ['ngcf', 'imcgae'])
def test_train_mode_forward_backward(tiny, px):
    """bpr_forward + the trainer's loss + backward through the SpMM autograd nodes, with the reference's recorded masks:
    loss, the L2 term and the gradient of EVERY parameter."""
    g = load_golden('tiny_ngcf_imcgae')
    _, model, _, l2 = _build(tiny, g, px)
    model.train()
    assert int(g[px + '_fb_n_dense']) == 3 and int(g[px + '_fb_n_rand']) == (1 if px == 'ngcf' else 0)
    model.injected = _injected(g, px, 'fb', 0, 0, 3, 0.1)
    t = _dev(g[px + '_fb_triples'])
    u_r, p_r, n_r, l2n = model.bpr_forward(t[:, 0], t[:, 1], t[:, 2])
    assert not model.injected['dense'] and 'edge' not in model.injected          # every recorded draw was consumed
    loss = torch.nn.functional.softplus((u_r * n_r).sum(1) - (u_r * p_r).sum(1)).mean() + l2 * l2n.mean()
    model.zero_grad()
    loss.backward()
    assert abs(loss.item() - float(g[px + '_fb_loss'])) < TOL
    assert rel_err(l2n.detach().cpu().numpy(), g[px + '_fb_l2_norm_sq']) < TOL
    for k, p in model.named_parameters():
        assert rel_err(p.grad.cpu().numpy(), g['%s_fb_grad_%s' % (px, k)]) < TOL, k


@pytest.mark.parametrize('px', ['ngcf', 'imcgae'])
def test_epoch_and_eval(tiny, px, tmp_path):
    """One epoch of BPRTrainer on the reference's triples and masks (trainer.AutogradStep + igcn_adam), then the three
    evaluations through the fused ranking kernel (exact form: the representation is wider than 64 columns)."""
    g = load_golden('tiny_ngcf_imcgae')
    ds, model, trainer, _ = _build(tiny, g, px)
    from igcn_cf_b200.trainer import AutogradStep
    assert isinstance(trainer.step, AutogradStep)
    model.train()
    tri = g[px + '_epoch_triples']
    trainer.step.reset_meter()
    n_steps = (len(tri) + 2047) // 2048
    assert int(g[px + '_epoch_n_dense']) == 3 * n_steps
    for s in range(n_steps):
        model.injected = _injected(g, px, 'epoch', s, 3 * s, 3, 0.1)
        trainer.step.run(_dev(tri[2048 * s:2048 * (s + 1)]))
    assert abs(trainer.step.meter_avg() - float(g[px + '_epoch_loss'])) < TOL
    for k, p in model.named_parameters():
        assert rel_err(p.detach().cpu().numpy(), g['%s_p1_%s' % (px, k)]) < TOL, k
    model.eval()
    with torch.no_grad():
        assert rel_err(model.get_rep().cpu().numpy(), g[px + '_rep1_eval']) < TOL
    _check_evals(trainer, g, px + '_e1', px + '_rep1_eval')
    # checkpoint round trip (BasicModel.save / load, model.py:45-49)
    path = str(tmp_path / (px + '.pth'))
    model.save(path)
    from igcn_cf_b200.model import get_model
    other = get_model(CFG[px][0], ds)
    other.load(path)
    other.eval()
    with torch.no_grad():
        assert torch.equal(other.get_rep(), model.get_rep())


@pytest.mark.parametrize('px', ['ngcf', 'imcgae'])
def test_public_epoch_loop_with_device_sampler(tiny, px):
    """train_one_epoch through the public loop: device sampler, torch-generated dropout; the loss is finite, every
    parameter moves and (NGCF; IMCGAE's node dropout at 3 steps per epoch is too noisy to say) falls over a few epochs."""
    g = load_golden('tiny_ngcf_imcgae')
    _, model, trainer, _ = _build(tiny, g, px, seed=5)
    before = {k: p.detach().clone() for k, p in model.named_parameters()}
    losses = [trainer.train_one_epoch() for _ in range(4)]
    assert all(np.isfinite(losses)) and (px != 'ngcf' or losses[-1] < losses[0])
    assert all(not torch.equal(before[k], p.detach()) for k, p in model.named_parameters())
    _, metrics = trainer.eval('val')
    assert 0. <= metrics['NDCG'][20] <= 1.


def test_initialisation_follows_the_reference_draw_order(tiny):
    """Same torch seed -> same initial parameters as the reference constructors (kaiming / normal draws in order)."""
    from igcn_cf_b200.model import get_model
    from igcn_cf_b200.utils import set_seed
    g = load_golden('tiny_ngcf_imcgae')
    for px in ('ngcf', 'imcgae'):
        set_seed(2021)
        model = get_model(CFG[px][0], _dataset(tiny))
        for k, p in model.named_parameters():
            assert np.array_equal(p.detach().cpu().numpy(), g['%s_p0_%s' % (px, k)]), (px, k)


@pytest.mark.parametrize('px', ['ngcf', 'imcgae'])
def test_against_the_oracle_on_a_larger_graph(px):
    """'small' shape (3000 x 4000, ~90 K interactions, rows long enough for the chunked-row path): eval representation,
    one train-mode forward/backward with seeded masks injected into both sides, and the top-20 of 512 users against
    oracle/restate.py (pinned to the reference by tests/test_oracle_golden.py)."""
    from igcn_cf_b200 import synth
    from igcn_cf_b200.dataset import get_dataset
    from igcn_cf_b200.model import get_model
    from igcn_cf_b200 import engine
    from conftest import check_topk_lists
    from oracle import restate as R
    ds = get_dataset({'name': 'SyntheticDataset', 'split': synth.gen_named('small', seed=11), 'device': DEV})
    U, I = ds.n_users, ds.n_items
    torch.manual_seed(4)
    mcfg, l2 = CFG[px]
    model = get_model(mcfg, ds)
    P = {k: v.detach().cpu().clone().requires_grad_(True) for k, v in model.named_parameters()}
    gen = torch.Generator().manual_seed(9)
    tri = torch.from_numpy(np.stack([np.arange(512) % U, [ds.train_data[u % U][0] for u in range(512)],
                                     (np.arange(512) * 7919) % I], axis=1).astype(np.int64))
    if px == 'ngcf':
        adj = R.row_normalized_adjacency(U, I, ds.train_pairs)
        gc = [(P['gc_layers.%d.weight' % l], P['gc_layers.%d.bias' % l]) for l in range(3)]
        bi = [(P['bi_layers.%d.weight' % l], P['bi_layers.%d.bias' % l]) for l in range(3)]
        edge = torch.rand(adj._nnz(), generator=gen) >= 0.1
        dense = [torch.rand(U + I, 64, generator=gen) >= 0.1 for _ in range(3)]
        oracle_rep = lambda train: R.ngcf_rep(adj, P['embedding.weight'], gc, bi, p=0.1, edge_keep=edge if train else None,
                                              dense_keep=dense if train else None)
        injected = {'edge': edge.to(DEV), 'dense': [d.to(DEV) for d in dense]}
    else:
        adj = R.normalized_adjacency(U, I, ds.train_pairs)
        node = [torch.rand(U + I, generator=gen) >= 0.3 - 0.1 * l for l in range(3)]
        oracle_rep = lambda train: R.imcgae_rep(adj, P['embedding.weight'], U, I, 3, p=0.3, node_keep=node if train else None)
        injected = {'dense': [d.to(DEV) for d in node]}
    model.eval()
    with torch.no_grad():
        mine = model.get_rep()
        want = oracle_rep(False)
    assert rel_err(mine.cpu().numpy(), want.numpy()) < TOL
    users = torch.arange(512, device=DEV)
    rec, _ = engine.score_topk(mine.contiguous(), users, U, I, 20)
    ref = torch.topk(want[:512] @ want[U:].t(), 20, dim=1)[1].numpy()
    check_topk_lists(rec.cpu().numpy(), ref, want.numpy(), U, rep_mine=mine.cpu().numpy(), scale_tol=TOL)
    model.train()
    model.injected = injected
    t = tri.to(DEV)
    u_r, p_r, n_r, l2n = model.bpr_forward(t[:, 0], t[:, 1], t[:, 2])
    loss = torch.nn.functional.softplus((u_r * n_r).sum(1) - (u_r * p_r).sum(1)).mean() + l2 * l2n.mean()
    model.zero_grad()
    loss.backward()
    ref_loss = R.rep_bpr_loss(oracle_rep(True), U, tri[:, 0], tri[:, 1], tri[:, 2], l2)
    ref_loss.backward()
    assert abs(loss.item() - ref_loss.item()) < TOL
    for k, p in model.named_parameters():
        assert rel_err(p.grad.cpu().numpy(), P[k].grad.numpy()) < TOL, k
