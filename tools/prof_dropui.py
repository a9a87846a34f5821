"""Host-side profile of the inductive (dropui) sequence: python tools/prof_dropui.py [shape]"""
import cProfile
import contextlib
import io
import os
import pstats
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from igcn_cf_b200 import synth  # noqa: E402
from igcn_cf_b200.dataset import get_dataset  # noqa: E402
from igcn_cf_b200.model import get_model  # noqa: E402
from igcn_cf_b200.trainer import get_trainer  # noqa: E402

shape = sys.argv[1] if len(sys.argv) > 1 else 'gowalla'
dev = torch.device('cuda:0')
full = synth.gen_named(shape, seed=2021)
with contextlib.redirect_stdout(io.StringIO()):
    ds_small = get_dataset({'name': 'SyntheticDataset', 'split': full, 'variant': 'dropui', 'device': dev})
    ds_full = get_dataset({'name': 'SyntheticDataset', 'split': full, 'device': dev})
    mcfg = {'name': 'IGCN', 'embedding_size': 64, 'n_layers': 3, 'device': dev, 'dropout': 0.3, 'feature_ratio': 1.}
    tcfg = {'name': 'IGCNTrainer', 'optimizer': 'Adam', 'lr': 1e-3, 'l2_reg': 0., 'aux_reg': 0.01, 'device': dev, 'n_epochs': 1,
            'batch_size': 2048, 'dataloader_num_workers': 0, 'test_batch_size': 512, 'topks': [20], 'seed': 2021}
    model = get_model(mcfg, ds_small)
    trainer = get_trainer(tcfg, ds_small, model)


def inductive():
    model.config['dataset'] = ds_full
    model.n_users, model.n_items = ds_full.n_users, ds_full.n_items
    model.norm_adj = model.generate_graph(ds_full)
    model.feat_mat, _, _, model.row_sum = model.generate_feat(ds_full, is_updating=True)
    model.update_feat_mat()
    tr = get_trainer(tcfg, ds_full, model)
    tr.inductive_eval(ds_small.n_users, ds_small.n_items)
    torch.cuda.synchronize()


with contextlib.redirect_stdout(io.StringIO()):
    inductive()
    pr = cProfile.Profile()
    pr.enable()
    inductive()
    pr.disable()
pstats.Stats(pr).sort_stats('cumulative').print_stats(28)
