"""CUDA path vs the CPU oracle (oracle/restate.py) at BASELINE.json's FULL sizes: Gowalla-shaped IGCN (configs[0],
the north-star target) and Yelp-shaped LightGCN (configs[1]).  Per workload: the evaluation-mode representation, one
training step on injected triples (loss, updated weights -- i.e. the gradient through Adam's first step, whose
update is lr * sign-like and therefore the strictest check of dE), and the masked top-20 of one 512-user batch
(the reference's own eval batch, trainer.py:145-164).  A few seconds of CPU work each (the CPU port does a
Yelp-shaped step in ~0.6 s on 16 cores).  Tolerances: 1e-5 relative (fp32, north_star); lists identical except at
proven ties of the oracle's scores."""
import numpy as np
import pytest
import torch

from conftest import check_topk_lists, rel_err

pytestmark = pytest.mark.gpu
DEV = torch.device('cuda:0')
TOL = 1e-5


def _triples(ds, n, seed):
    rng = np.random.default_rng(seed)
    pairs = ds.train_pairs
    sel = rng.integers(len(pairs), size=n)
    return np.stack([pairs[sel, 0], pairs[sel, 1], rng.integers(ds.n_items, size=n)], axis=1).astype(np.int64)


@pytest.mark.parametrize('shape,kind', [('gowalla', 'IGCN'), ('yelp', 'LightGCN')])
def test_full_size_rep_step_and_topk_against_the_oracle(shape, kind):
    from igcn_cf_b200.dataset import get_dataset
    from igcn_cf_b200.model import get_model
    from igcn_cf_b200.trainer import get_trainer
    from oracle import restate as R
    ds = get_dataset({'name': 'SyntheticDataset', 'shape': shape, 'seed': 2021, 'device': DEV})
    torch.manual_seed(2021)
    mcfg = {'name': kind, 'embedding_size': 64, 'n_layers': 3, 'device': DEV}
    tcfg = {'optimizer': 'Adam', 'lr': 1e-3, 'device': DEV, 'n_epochs': 1, 'batch_size': 2048, 'dataloader_num_workers': 0,
            'test_batch_size': 512, 'topks': [20], 'cuda_graph': False}
    if kind == 'IGCN':
        mcfg.update(dropout=0.0, feature_ratio=1.)          # train-mode parity without an injected mask (Amazon's setting)
        tcfg.update(name='IGCNTrainer', l2_reg=0., aux_reg=0.01)
    else:
        tcfg.update(name='BPRTrainer', l2_reg=1e-4)
    model = get_model(mcfg, ds)
    trainer = get_trainer(tcfg, ds, model)
    emb0 = model.embedding.weight.detach().cpu().numpy()
    orc = (R.OracleIGCN(ds.n_users, ds.n_items, ds.train_pairs, 3, emb0, 0.0) if kind == 'IGCN'
           else R.OracleLightGCN(ds.n_users, ds.n_items, ds.train_pairs, 3, emb0, l2_reg=1e-4))

    # ---- representation (eval mode)
    model.eval()
    with torch.no_grad():
        rep = model.get_rep().cpu().numpy()
        rep_ref = orc.get_rep().detach().numpy()
    assert rel_err(rep, rep_ref) < TOL

    # ---- masked top-20 of one 512-user batch (the reference's eval('val') loop body on its first batch)
    users = list(range(512))
    with torch.no_grad():
        scores = orc.predict(torch.tensor(users, dtype=torch.int64))
    _, ref_items = R.masked_topk(scores, users, 20, ds.train_data)
    rec, _ = trainer.recommend('val', users=torch.arange(512, device=DEV), users_host=np.arange(512, dtype=np.int64))
    n_diff = check_topk_lists(rec.cpu().numpy(), ref_items, rep_ref, ds.n_users, rep_mine=rep, scale_tol=TOL)
    assert n_diff <= 5                                        # ties are rare

    # ---- one training step on injected triples: loss and the weights after Adam's first update
    t = torch.from_numpy(_triples(ds, 2048, 1))
    a = torch.from_numpy(_triples(ds, 2048, 2))
    model.train()
    if kind == 'IGCN':
        ref_loss = orc.train_step(t[:, 0], t[:, 1], t[:, 2], a[:, 0], a[:, 1], a[:, 2])
        loss = trainer.step.run(t.to(DEV), a.to(DEV)).item()
    else:
        ref_loss = orc.train_step(t[:, 0], t[:, 1], t[:, 2])
        loss = trainer.step.run(t.to(DEV)).item()
    assert abs(loss - ref_loss) < TOL * max(1.0, abs(ref_loss))
    # gradient itself: d_emb of the fused step against the oracle's autograd gradient, relative to its scale
    assert rel_err(trainer.step.d_emb.cpu().numpy(), orc.emb.grad.numpy()) < TOL      # opt.step() leaves .grad in place
    assert rel_err(model.embedding.weight.detach().cpu().numpy(), orc.emb.detach().numpy()) < TOL
    if kind == 'IGCN':
        assert rel_err(model.w.detach().cpu().numpy(), orc.w.detach().numpy()) < TOL
