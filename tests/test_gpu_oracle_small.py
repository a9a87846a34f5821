"""CUDA path vs the CPU oracle on a larger seeded graph ('small': 3000 x 4000, ~90K interactions,
power-law heads long enough to hit the chunked-row path)."""
import numpy as np
import pytest
import torch

from conftest import check_metrics, check_topk_lists, rel_err

pytestmark = pytest.mark.gpu
DEV = torch.device('cuda:0')
TOL = 1e-5


@pytest.fixture(scope='module')
def small():
    from igcn_cf_b200 import synth
    from igcn_cf_b200.dataset import get_dataset
    split = synth.gen_named('small', seed=11)
    ds = get_dataset({'name': 'SyntheticDataset', 'split': split, 'device': DEV})
    return ds


def _rand_triples(ds, n, seed):
    rng = np.random.default_rng(seed)
    out = np.zeros((n, 3), dtype=np.int64)
    for t in range(n):
        u = int(rng.integers(ds.n_users))
        out[t] = (u, int(rng.choice(ds.train_data[u])), int(rng.integers(ds.n_items)))
    return out


def test_lightgcn_three_steps_and_topk(small):
    from igcn_cf_b200.model import get_model
    from igcn_cf_b200.trainer import get_trainer
    from oracle import restate as R
    ds = small
    torch.manual_seed(0)
    model = get_model({'name': 'LightGCN', 'embedding_size': 64, 'n_layers': 3, 'device': DEV}, ds)
    assert model.norm_adj.csr.n_chunks > 0                          # long rows present
    trainer = get_trainer({'name': 'BPRTrainer', 'optimizer': 'Adam', 'lr': 1e-3, 'l2_reg': 1e-4, 'device': DEV,
                           'n_epochs': 1, 'batch_size': 2048, 'dataloader_num_workers': 0, 'test_batch_size': 512,
                           'topks': [20], 'cuda_graph': False}, ds, model)
    emb0 = model.embedding.weight.detach().cpu().numpy()
    orc = R.OracleLightGCN(ds.n_users, ds.n_items, ds.train_pairs, 3, emb0, l2_reg=1e-4)
    assert rel_err(model.get_rep().detach().cpu().numpy(), orc.get_rep().detach().numpy()) < TOL
    model.train()
    for s in range(3):
        tri = _rand_triples(ds, 2048, s)
        t = torch.from_numpy(tri)
        ref_loss = orc.train_step(t[:, 0], t[:, 1], t[:, 2])
        loss = trainer.step.run(t.to(DEV)).item()
        assert abs(loss - ref_loss) < TOL
    assert rel_err(model.embedding.weight.detach().cpu().numpy(), orc.emb.detach().numpy()) < TOL
    metrics, rec = R.evaluate(orc, 'test', ds.train_data, ds.val_data, ds.test_data, [20])
    _, mine = trainer.eval('test')
    rec_dev, _ = trainer.recommend('test')
    model.eval()
    with torch.no_grad():
        rep_mine = model.get_rep().cpu().numpy()
    n_diff = check_topk_lists(rec_dev.cpu().numpy(), rec, orc.get_rep().detach().numpy(), ds.n_users, rep_mine=rep_mine,
                              scale_tol=TOL)
    check_metrics(mine, [(name, 20, metrics[name][20]) for name in metrics], n_diff,
                  sum(1 for x in ds.test_data if len(x) > 0))


def test_igcn_step_with_zero_dropout(small):
    """Amazon-style config (dropout 0.0, reference config.py:169): train-mode parity needs no mask."""
    from igcn_cf_b200.model import get_model
    from igcn_cf_b200.trainer import get_trainer
    from oracle import restate as R
    ds = small
    torch.manual_seed(1)
    model = get_model({'name': 'IGCN', 'embedding_size': 64, 'n_layers': 3, 'device': DEV, 'dropout': 0.0,
                       'feature_ratio': 1.}, ds)
    trainer = get_trainer({'name': 'IGCNTrainer', 'optimizer': 'Adam', 'lr': 1e-3, 'l2_reg': 0., 'aux_reg': 0.01,
                           'device': DEV, 'n_epochs': 1, 'batch_size': 2048, 'dataloader_num_workers': 0,
                           'test_batch_size': 512, 'topks': [20], 'cuda_graph': False}, ds, model)
    emb0 = model.embedding.weight.detach().cpu().numpy()
    orc = R.OracleIGCN(ds.n_users, ds.n_items, ds.train_pairs, 3, emb0, 0.0)
    model.train()
    for s in range(2):
        t, a = torch.from_numpy(_rand_triples(ds, 2048, s)), torch.from_numpy(_rand_triples(ds, 2048, 100 + s))
        ref_loss = orc.loss(t[:, 0], t[:, 1], t[:, 2], a[:, 0], a[:, 1], a[:, 2], train=False)
        orc.opt.zero_grad()
        ref_loss.backward()
        orc.opt.step()
        loss = trainer.step.run(t.to(DEV), a.to(DEV)).item()
        assert abs(loss - ref_loss.item()) < TOL
    assert rel_err(model.embedding.weight.detach().cpu().numpy(), orc.emb.detach().numpy()) < TOL
    assert rel_err(model.w.detach().cpu().numpy(), orc.w.detach().numpy()) < TOL


def test_igcn_partial_templates_training_and_device_sampler(small):
    """feature_ratio 0.5 (model.py:386-421 with core users/items only): the fused step works in template-id
    space for the auxiliary loss, template rows without a node get no gradient, and the device sampler of the
    auxiliary stream only draws template users / template items of their train lists (dataset.py:258-273)."""
    from igcn_cf_b200.model import get_model
    from igcn_cf_b200.trainer import get_trainer
    from oracle import restate as R
    ds = small
    torch.manual_seed(2)
    model = get_model({'name': 'IGCN', 'embedding_size': 64, 'n_layers': 3, 'device': DEV, 'dropout': 0.0,
                       'feature_ratio': 0.5, 'ranking_metric': 'degree'}, ds)
    t_u, t_i = len(model.user_map), len(model.item_map)
    assert (t_u, t_i) == (ds.n_users // 2, ds.n_items // 2) and model.embedding.weight.shape[0] == t_u + t_i + 2
    trainer = get_trainer({'name': 'IGCNTrainer', 'optimizer': 'Adam', 'lr': 1e-3, 'l2_reg': 1e-4, 'aux_reg': 0.05,
                           'device': DEV, 'n_epochs': 1, 'batch_size': 2048, 'dataloader_num_workers': 0,
                           'test_batch_size': 512, 'topks': [20], 'cuda_graph': False, 'seed': 9}, ds, model)
    emb0 = model.embedding.weight.detach().cpu().numpy()
    orc = R.OracleIGCN(ds.n_users, ds.n_items, ds.train_pairs, 3, emb0, 0.0, l2_reg=1e-4, aux_reg=0.05,
                       user_map=dict(model.user_map), item_map=dict(model.item_map))
    model.eval()
    with torch.no_grad():
        assert rel_err(model.get_rep().cpu().numpy(), orc.get_rep().detach().numpy()) < TOL
    model.train()
    rng = np.random.default_rng(5)
    for s in range(2):
        t = torch.from_numpy(_rand_triples(ds, 2048, 10 + s))
        a = torch.from_numpy(np.stack([rng.integers(t_u, size=2048), rng.integers(t_i, size=2048),
                                       rng.integers(t_i, size=2048)], axis=1).astype(np.int64))
        ref_loss = orc.loss(t[:, 0], t[:, 1], t[:, 2], a[:, 0], a[:, 1], a[:, 2], train=False)
        orc.opt.zero_grad()
        ref_loss.backward()
        orc.opt.step()
        loss = trainer.step.run(t.to(DEV), a.to(DEV)).item()
        assert abs(loss - ref_loss.item()) < TOL
    assert rel_err(model.embedding.weight.detach().cpu().numpy(), orc.emb.detach().numpy()) < TOL
    assert rel_err(model.w.detach().cpu().numpy(), orc.w.detach().numpy()) < TOL
    # device sampler in template space: valid ids, positives from the user's template items, negatives outside
    inv_u = {v: k for k, v in model.user_map.items()}
    trainer.step.run()
    aux = trainer.step.a_triples.cpu().numpy()
    assert aux[:, 0].max() < t_u and aux[:, 1:].max() < t_i and aux.min() >= 0
    for u, p, n in aux[:300]:
        tmpl_items = {model.item_map[i] for i in ds.train_data[inv_u[int(u)]] if i in model.item_map}
        assert int(p) in tmpl_items and int(n) not in tmpl_items


def test_duplicate_interactions_and_empty_users():
    """Collisions and holes the reference tolerates: a (user, item) pair listed twice becomes an adjacency value
    of 2 (utils.py:46-48 sums duplicates, degrees count it twice), users / items without any train interaction
    keep degree clamp 1 (model.py:87-88) and are skipped by the sampler (dataset.py:120-122)."""
    from igcn_cf_b200.dataset import get_dataset
    from igcn_cf_b200.model import get_model
    from igcn_cf_b200.trainer import get_trainer
    from oracle import restate as R
    rng = np.random.default_rng(3)
    n_users, n_items = 90, 120
    train = [sorted(rng.choice(n_items - 5, size=int(rng.integers(3, 15)), replace=False).tolist()) for _ in range(n_users)]
    train[7] = []                                   # user without train items
    train[20] = train[20] + train[20][:2]           # duplicated interactions
    train[33] = [5, 5, 5, 9]
    empty = [[] for _ in range(n_users)]
    test = [[int(rng.integers(n_items))] for _ in range(n_users)]
    ds = get_dataset({'name': 'ListDataset', 'train': train, 'val': empty, 'test': test, 'n_items': n_items, 'device': DEV})
    for kind in ('LightGCN', 'IGCN'):
        torch.manual_seed(4)
        cfg = {'name': kind, 'embedding_size': 64, 'n_layers': 2, 'device': DEV}
        if kind == 'IGCN':
            cfg.update(dropout=0.0, feature_ratio=1.)
        model = get_model(cfg, ds)
        emb0 = model.embedding.weight.detach().cpu().numpy()
        orc = (R.OracleLightGCN(n_users, n_items, ds.train_pairs, 2, emb0) if kind == 'LightGCN'
               else R.OracleIGCN(n_users, n_items, ds.train_pairs, 2, emb0, 0.0))
        model.eval()
        with torch.no_grad():
            assert rel_err(model.get_rep().cpu().numpy(), orc.get_rep().detach().numpy()) < TOL, kind
    trainer = get_trainer({'name': 'IGCNTrainer', 'optimizer': 'Adam', 'lr': 1e-3, 'l2_reg': 0., 'aux_reg': 0.01,
                           'device': DEV, 'n_epochs': 1, 'batch_size': 512, 'dataloader_num_workers': 0,
                           'test_batch_size': 512, 'topks': [5], 'cuda_graph': False, 'seed': 1}, ds, model)
    model.train()
    trainer.step.run()
    tri = trainer.step.triples[:512].cpu().numpy()
    assert 7 not in set(tri[:, 0].tolist())          # the empty user is never drawn
    seen = [set(x) for x in train]
    assert all(p in seen[u] and n not in seen[u] for u, p, n in tri)
    _, metrics = trainer.eval('test')                # items 115..119 have no interaction at all: still rankable
    assert 0.0 <= metrics['Recall'][5] <= 1.0
