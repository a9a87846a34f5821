"""The drop-in boundary on the host (no GPU): the `dropin/` directory put first on sys.path makes the reference's
top-level module names (`model`, `trainer`, `dataset`, `utils`, `config` -- what run/run.py:1-7 and
run/dropui/igcn_dropui.py:1-7 import) resolve to this package, the configuration lists keep the reference's slot
positions, the checkpoint policy of the epoch loop is safe with several ranks, and the tie-proving list checker
used by the GPU parity tests accepts ties and nothing else."""
import os
import subprocess
import sys

import numpy as np
import pytest

from conftest import ROOT, check_metrics, check_topk_lists


def test_reference_module_names_resolve_to_this_package():
    code = r'''
import sys
sys.path.insert(0, sys.argv[1]); sys.path.insert(0, sys.argv[2])
import model, trainer, dataset, utils, config
import igcn_cf_b200.model as M, igcn_cf_b200.trainer as T, igcn_cf_b200.dataset as D, igcn_cf_b200.utils as U
assert model.IGCN is M.IGCN and model.LightGCN is M.LightGCN and model.get_model is M.get_model
assert trainer.IGCNTrainer is T.IGCNTrainer and trainer.BPRTrainer is T.BPRTrainer and trainer.get_trainer is T.get_trainer
assert dataset.get_dataset is D.get_dataset and dataset.ProcessedDataset is D.ProcessedDataset
assert utils.init_run is U.init_run and utils.set_seed is U.set_seed and utils.AverageMeter is U.AverageMeter
# the statements at the top of the reference launchers, verbatim
from dataset import get_dataset
from model import get_model
from trainer import get_trainer
from utils import init_run
from config import get_gowalla_config, get_yelp_config, get_amazon_config
cfg = get_gowalla_config('cuda')
assert cfg[2][1]['name'] == 'IGCN' and cfg[2][2]['name'] == 'IGCNTrainer'      # run/run.py:15
assert cfg[1][1]['name'] == 'LightGCN' and cfg[0][1]['name'] == 'MF' and cfg[6][1]['name'] == 'IMF'
print('ok')
'''
    out = subprocess.run([sys.executable, '-c', code, ROOT, os.path.join(ROOT, 'dropin')], capture_output=True, text=True,
                         timeout=300)
    assert out.returncode == 0 and out.stdout.strip().endswith('ok'), out.stderr[-2000:]


def test_config_slots_and_hyperparameters_match_the_reference_table():
    """reference config.py:6-72, 81-147, 156-222: positions and the hot-path hyper-parameters."""
    from igcn_cf_b200 import config as C
    names = ['MF', 'LightGCN', 'IGCN', 'ItemKNN', 'NGCF', 'MultiVAE', 'IMF', 'IMCGAE', 'IDCF_LGCN', 'NeuMF']
    want = {'gowalla': ((1e-4, 1e-3), 1e-4, 0.3, (0.1, 0.1), 'data/Gowalla/time', (0.1, 1e-3), 0.3),
            'yelp': ((1e-3, 1e-3), 1e-4, 0.3, (0.5, 0.01), 'data/Yelp/time', (0.3, 1e-3), 0.3),
            'amazon': ((1e-3, 1e-4), 1e-5, 0.0, (0.3, 0.1), 'data/Amazon/time', (0.3, 1e-4), 0.9)}
    for key, fn in (('gowalla', C.get_gowalla_config), ('yelp', C.get_yelp_config), ('amazon', C.get_amazon_config)):
        cfg = fn('cuda')
        assert [m['name'] for _, m, _ in cfg] == names
        mf, l2, drop, imf, path, ngcf, imcgae = want[key]
        assert all(d == {'name': 'ProcessedDataset', 'path': path, 'device': 'cuda'} for d, _, _ in cfg)
        assert (cfg[0][2]['lr'], cfg[0][2]['l2_reg'], cfg[0][2]['name']) == (mf[0], mf[1], 'BPRTrainer')
        assert cfg[1][1] == {'name': 'LightGCN', 'embedding_size': 64, 'n_layers': 3, 'device': 'cuda'}
        assert (cfg[1][2]['l2_reg'], cfg[1][2]['lr'], cfg[1][2]['batch_size']) == (l2, 1e-3, 2048)
        assert cfg[2][1] == {'name': 'IGCN', 'embedding_size': 64, 'n_layers': 3, 'device': 'cuda', 'dropout': drop,
                             'feature_ratio': 1.}
        assert (cfg[2][2]['l2_reg'], cfg[2][2]['aux_reg'], cfg[2][2]['topks']) == (0., 0.01, [20])
        assert (cfg[6][1]['dropout'], cfg[6][2]['aux_reg'], cfg[6][2]['l2_reg'], cfg[6][1]['n_layers']) == (imf[0], imf[1], 1e-5, 0)
        # reference config.py:30-34, 51-55 (and the Yelp / Amazon counterparts): the sibling models on the same kernels
        assert cfg[4][1] == {'name': 'NGCF', 'embedding_size': 64, 'layer_sizes': [64, 64, 64], 'device': 'cuda', 'dropout': ngcf[0]}
        assert (cfg[4][2]['name'], cfg[4][2]['l2_reg'], cfg[4][2]['lr']) == ('BPRTrainer', ngcf[1], 1e-3)
        assert cfg[7][1] == {'name': 'IMCGAE', 'embedding_size': 64, 'n_layers': 3, 'device': 'cuda', 'dropout': imcgae}
        assert (cfg[7][2]['name'], cfg[7][2]['l2_reg']) == ('BPRTrainer', 0.)
        assert all(cfg[i][1].get('out_of_scope') for i in (3, 5, 8, 9))
    from igcn_cf_b200.model import get_model
    with pytest.raises(NotImplementedError, match='outside the B200 hot path'):
        get_model(C.get_gowalla_config('cuda')[3][1], None)


class _FakeModel:
    name = 'IGCN'

    def __init__(self):
        self.saved, self.loaded = [], []

    def save(self, path):
        with open(path, 'w') as f:
            f.write('ckpt')
        self.saved.append(path)

    def load(self, path):
        assert os.path.exists(path)
        self.loaded.append(path)


class _FakeTrainer:
    name, max_patience, val_interval, best_ndcg, save_path = 'IGCNTrainer', 3, 1, -np.inf, None

    def __init__(self):
        self.model = _FakeModel()
        self.dataset = type('D', (), {'name': 'ProcessedDataset'})()


def test_best_checkpoint_policy(tmp_path):
    """trainer.py:90-106: better NDCG -> new file named by ndcg*100 with 3 decimals, previous best removed, patience
    reset; otherwise patience shrinks by val_interval; the file is written atomically (no .tmp left behind)."""
    from igcn_cf_b200.trainer import BestCheckpoint
    tr = _FakeTrainer()
    keeper = BestCheckpoint(tr, str(tmp_path / 'checkpoints'))
    assert keeper.offer(0.05) and tr.best_ndcg == 0.05
    first = tr.save_path
    assert os.path.basename(first) == 'IGCN_IGCNTrainer_ProcessedDataset_5.000.pth' and os.path.exists(first)
    assert keeper.offer(0.04) and keeper.patience == 2 and tr.save_path == first
    assert keeper.offer(0.0612345) and not os.path.exists(first) and keeper.patience == 3
    assert os.path.basename(tr.save_path) == 'IGCN_IGCNTrainer_ProcessedDataset_6.123.pth'
    assert keeper.offer(0.01) and keeper.offer(0.01) and not keeper.offer(0.01)          # patience 3 -> 0: stop
    assert sorted(os.listdir(tmp_path / 'checkpoints')) == ['IGCN_IGCNTrainer_ProcessedDataset_6.123.pth']
    keeper.restore()
    assert tr.model.loaded == [tr.save_path]


def test_best_checkpoint_non_writer_rank_touches_no_file(tmp_path, monkeypatch):
    """With one process per GPU only rank 0 writes / removes (every rank used to os.remove the same file)."""
    from igcn_cf_b200 import trainer as T
    monkeypatch.setattr(T.BestCheckpoint, '_rank', staticmethod(lambda: 1))
    tr = _FakeTrainer()
    keeper = T.BestCheckpoint(tr, str(tmp_path / 'ck'))
    assert keeper.offer(0.5) and tr.model.saved == [] and not os.path.exists(tmp_path / 'ck')
    assert tr.save_path.endswith('50.000.pth') and tr.best_ndcg == 0.5


def test_list_checker_accepts_ties_only():
    rng = np.random.default_rng(0)
    U, I, D = 6, 40, 8
    rep = rng.standard_normal((U + I, D)).astype(np.float32)
    rep[U + 7] = rep[U + 3]                                   # items 3 and 7 tie exactly for every user
    scores = rep[:U].astype(np.float64) @ rep[U:].astype(np.float64).T
    ref = np.argsort(-scores, axis=1, kind='stable')[:, :5]
    assert check_topk_lists(ref, ref, rep, U) == 0
    swapped = ref.copy()
    for u in range(U):
        row = swapped[u].tolist()
        if 3 in row and 7 in row:
            a, b = row.index(3), row.index(7)
            swapped[u, a], swapped[u, b] = 7, 3
    n = int((swapped != ref).any(axis=1).sum())
    assert check_topk_lists(swapped, ref, rep, U, rep_mine=rep) == n
    wrong = ref.copy()
    wrong[0, 0], wrong[0, 1] = ref[0, 1], ref[0, 0]           # swaps two items whose scores differ
    if scores[0, ref[0, 0]] != scores[0, ref[0, 1]]:
        with pytest.raises(AssertionError, match='no tie'):
            check_topk_lists(wrong, ref, rep, U)
    m = {'Recall': {20: 0.25}}
    check_metrics(m, [('Recall', 20, 0.25)], 0, 10)
    with pytest.raises(AssertionError):
        check_metrics(m, [('Recall', 20, 0.2500001)], 0, 10)
    check_metrics(m, [('Recall', 20, 0.30)], 1, 10)           # one tied user of ten: bound 0.1
    with pytest.raises(AssertionError):
        check_metrics(m, [('Recall', 20, 0.40)], 1, 10)
