"""GPU tests of the drop-in surface that round 1 left uncovered: the launcher flow of run/run.py:10-26 and
run/dropui/igcn_dropui.py:10-48 executed through the `dropin/` aliases (reference module names), the host
'reference' sampler (trainer.py:226-227, 285-289) replaying the reference's own draw sequence, the generic
optimizer step followed by eval, the learning rate following the optimizer's param group, and
IGCN.inductive_rep_layer (model.py:423-432)."""
import os
import subprocess
import sys

import numpy as np
import pytest
import torch

from conftest import ROOT, load_golden, rel_err

pytestmark = pytest.mark.gpu
DEV = torch.device('cuda:0')
TOL = 1e-5

LAUNCHER = r'''
import os, sys
sys.path.insert(0, sys.argv[1]); sys.path.insert(0, sys.argv[2])
os.chdir(sys.argv[3])
# ---- what run/run.py:1-7 and run/dropui/igcn_dropui.py:1-7 import (tensorboardX is not installed here)
from dataset import get_dataset
from model import get_model
from trainer import get_trainer
import torch
from utils import init_run
from config import get_gowalla_config
import model as model_module, igcn_cf_b200.model
assert model_module.IGCN is igcn_cf_b200.model.IGCN

init_run(os.path.join(sys.argv[3], 'log'), 2021)
device = torch.device('cuda')
config = get_gowalla_config(device)
dataset_config, model_config, trainer_config = config[2]
dataset_config['path'] = os.path.join(sys.argv[3], 'time_0_dropui')
trainer_config = dict(trainer_config, n_epochs=2, dataloader_num_workers=0)

dataset = get_dataset(dataset_config)
model = get_model(model_config, dataset)
trainer = get_trainer(trainer_config, dataset, model)
best = trainer.train(verbose=True, writer=None)
results, _ = trainer.eval('test')
print('Test result. {:s}'.format(results))
saved = trainer.save_path

dataset_config['path'] = dataset_config['path'][:-7]
new_dataset = get_dataset(dataset_config)
model.config['dataset'] = new_dataset
model.n_users, model.n_items = new_dataset.n_users, new_dataset.n_items
model.norm_adj = model.generate_graph(new_dataset)
model.feat_mat, _, _, model.row_sum = model.generate_feat(new_dataset, is_updating=True)
model.update_feat_mat()
trainer = get_trainer(trainer_config, new_dataset, model)
print('Inductive results.')
trainer.inductive_eval(dataset.n_users, dataset.n_items)

model_config['name'] = 'Popularity'
trainer_config['name'] = 'BasicTrainer'
model = get_model(model_config, new_dataset)
trainer = get_trainer(trainer_config, new_dataset, model)
print('Popularity model results.')
trainer.inductive_eval(dataset.n_users, dataset.n_items)
print('SAVED', saved, os.path.exists(saved), 'BEST', best)
'''


def test_reference_launcher_flow_through_the_aliases(tmp_path):
    from igcn_cf_b200 import synth
    full = synth.gen_named('tiny', seed=2021)
    synth.write_split(full, str(tmp_path / 'time_0'))          # igcn_dropui.py:26 strips '_dropui' from the path
    synth.write_split(synth.dropui(full), str(tmp_path / 'time_0_dropui'))
    out = subprocess.run([sys.executable, '-c', LAUNCHER, ROOT, os.path.join(ROOT, 'dropin'), str(tmp_path)],
                         capture_output=True, text=True, timeout=600)
    log = tmp_path / 'log' / 'log.txt'              # init_run redirects stdout / stderr there (utils.py:23-29)
    text = log.read_text() if log.exists() else ''
    assert out.returncode == 0, (out.stderr + text)[-3000:]
    assert text.count('Epoch ') == 2 and 'Validation result. Precision: ' in text and 'Best NDCG, save model to' in text
    assert 'Test result. Precision: ' in text
    for title in ('All users and all items', 'Old users and all items', 'New users and all items',
                  'All users and old items', 'All users and new items', 'Old users and old items'):
        assert text.count(title + ' result. Precision: ') == 2, title          # IGCN pass + Popularity pass
    last = text.strip().splitlines()[-1].split()
    assert last[0] == 'SAVED' and last[2] == 'True' and float(last[4]) > 0.0
    assert os.path.basename(last[1]).startswith('IGCN_IGCNTrainer_ProcessedDataset_')


def _trainer(kind, tiny, g, **tr):
    from igcn_cf_b200.dataset import get_dataset
    from igcn_cf_b200.model import get_model
    from igcn_cf_b200.trainer import get_trainer
    ds = get_dataset({'name': 'ListDataset', 'train': tiny['train'], 'val': tiny['val'], 'test': tiny['test'],
                      'n_items': tiny['n_items'], 'device': DEV})
    mcfg = {'name': kind, 'embedding_size': 64, 'n_layers': 3, 'device': DEV}
    tcfg = {'optimizer': 'Adam', 'lr': 1e-3, 'device': DEV, 'n_epochs': 1, 'batch_size': 2048,
            'dataloader_num_workers': 0, 'test_batch_size': 512, 'topks': [5, 20]}
    if kind == 'IGCN':
        mcfg.update(dropout=0.3, feature_ratio=1.)
        tcfg.update(name='IGCNTrainer', l2_reg=0., aux_reg=0.01)
    else:
        tcfg.update(name='BPRTrainer', l2_reg=1e-4)
    tcfg.update(tr)
    model = get_model(mcfg, ds)
    with torch.no_grad():
        model.embedding.weight.copy_(torch.from_numpy(g['emb0']))
    return ds, model, get_trainer(tcfg, ds, model)


def _spy_steps(trainer):
    seen = []
    run = trainer.step.run

    def spy(*a, **k):
        seen.append([None if t is None else t.cpu().numpy().copy() for t in a])
        return run(*a, **k)

    trainer.step.run = spy
    return seen


def test_reference_sampler_replays_the_reference_epoch_lightgcn(tiny):
    """'sampler': 'reference' = the reference's DataLoader over BasicDataset.__getitem__ (dataset.py:119-131): seeded
    like tests/golden/make_golden.py seeded the reference, it draws the SAME triples, so one epoch reproduces the
    reference's epoch loss and weights."""
    from igcn_cf_b200.utils import set_seed
    g = load_golden('tiny_lightgcn')
    _, model, trainer = _trainer('LightGCN', tiny, g, sampler='reference', cuda_graph=False)
    seen = _spy_steps(trainer)
    set_seed(2021 + 1)
    model.train()
    loss = trainer.train_one_epoch()
    assert np.array_equal(np.concatenate([s[0] for s in seen]), g['epoch_triples'])
    assert abs(loss - float(g['epoch_loss'])) < TOL
    assert rel_err(model.embedding.weight.detach().cpu().numpy(), g['emb1']) < TOL


def test_reference_sampler_draws_the_reference_streams_igcn(tiny):
    """IGCN zips the main and the auxiliary loader (trainer.py:296): both streams equal the reference's recording
    (the dropout draws use torch's generator, not python's / numpy's, so they do not disturb the triple streams)."""
    from igcn_cf_b200.utils import set_seed
    g = load_golden('tiny_igcn')
    _, model, trainer = _trainer('IGCN', tiny, g, sampler='reference', cuda_graph=False)
    seen = _spy_steps(trainer)
    set_seed(2021 + 1)
    model.train()
    alpha0 = model.alpha
    trainer.train_one_epoch()
    assert np.array_equal(np.concatenate([s[0] for s in seen]), g['epoch_triples'])
    assert np.array_equal(np.concatenate([s[1] for s in seen]), g['epoch_aux_triples'])
    assert len(seen) == int(g['epoch_n_steps']) and model.alpha == alpha0 * model.delta == float(g['alpha1'])


def test_generic_optimizer_step_invalidates_the_cached_representation(tiny):
    """loss.backward(); trainer.opt.step(); trainer.eval() must see the updated parameters (igcn_adam
    writes through raw pointers, so autograd's version counter does not move)."""
    g = load_golden('tiny_lightgcn')
    _, model, trainer = _trainer('LightGCN', tiny, g, cuda_graph=False)
    model.eval()
    with torch.no_grad():
        rep_before = model.get_rep().clone()
    _, m0 = trainer.eval('val')
    model.train()
    t = torch.as_tensor(g['fb_triples'], device=DEV)
    for _ in range(20):
        trainer.opt.zero_grad()
        u_r, p_r, n_r, l2 = model.bpr_forward(t[:, 0], t[:, 1], t[:, 2])
        loss = torch.nn.functional.softplus((u_r * n_r).sum(1) - (u_r * p_r).sum(1)).mean() + 1e-4 * l2.mean()
        loss.backward()
        trainer.opt.step()
    model.eval()
    with torch.no_grad():
        rep_after = model.get_rep()
    assert float((rep_after - rep_before).abs().max()) > 1e-4
    # and the new representation is the propagation of the NEW weights
    from oracle import restate as R
    orc = R.OracleLightGCN(model.n_users, model.n_items, tiny['pairs'], 3, model.embedding.weight.detach().cpu().numpy())
    assert rel_err(rep_after.cpu().numpy(), orc.get_rep().detach().numpy()) < TOL


@pytest.mark.parametrize('graph_mode', [False, True])
def test_fused_step_follows_the_param_group_learning_rate(tiny, graph_mode):
    """the fused step used to freeze lr at construction (also inside captured graphs)."""
    g = load_golden('tiny_lightgcn')
    _, model, trainer = _trainer('LightGCN', tiny, g, cuda_graph=graph_mode)
    model.train()
    t = torch.as_tensor(g['fb_triples'], device=DEV)
    trainer.step.run(t)
    w1 = model.embedding.weight.detach().clone()
    assert float((w1.cpu() - torch.from_numpy(g['emb0'])).abs().max()) > 1e-5
    trainer.opt.param_groups[0]['lr'] = 0.0
    trainer.step.run(t)
    assert torch.equal(model.embedding.weight.detach(), w1)
    trainer.opt.param_groups[0]['lr'] = 1e-3
    trainer.step.run(t)
    assert not torch.equal(model.embedding.weight.detach(), w1)


def test_inductive_rep_layer_is_feat_times_embedding(tiny):
    """IGCN.inductive_rep_layer (model.py:423-432) = feat_mat @ embedding.weight, against the reference's own
    feat_mat tensors (golden) multiplied on the host in fp64."""
    g = load_golden('tiny_igcn')
    _, model, _ = _trainer('IGCN', tiny, g)
    model.eval()
    with torch.no_grad():
        x0 = model.inductive_rep_layer(model.feat_mat).cpu().numpy()
    import scipy.sparse as sp
    shape = tuple(int(x) for x in g['feat_shape'])
    F = sp.csr_matrix((g['feat_val'].astype(np.float64), (g['feat_idx'][0], g['feat_idx'][1])), shape=shape)
    want = F @ g['emb0'].astype(np.float64)
    assert x0.shape == want.shape and rel_err(x0, want) < TOL
