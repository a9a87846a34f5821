"""Configuration tables with the reference's shape: `get_<dataset>_config(device)` returns a LIST of
(dataset_config, model_config, trainer_config) triples and the reference launchers index it by position
(`config[2]` is IGCN in run/run.py:15 and run/dropui/igcn_dropui.py:16), so every slot of the reference's list
keeps its index here (reference config.py:6-72, 81-147, 156-222):

    0 MF   1 LightGCN   2 IGCN   3 ItemKNN   4 NGCF   5 MultiVAE   6 IMF   7 IMCGAE   8 IDCF_LGCN   9 NeuMF

Slots 0, 1, 2 and 6 are the models on the B200 hot path; slots 4 and 7 (NGCF, IMCGAE) are the sibling models that
reuse its propagation and ranking kernels (igcn_cf_b200/siblings.py).  All six carry the reference's hyper-parameters
per dataset.  The other slots name baseline models that are out of scope; they are present so that
the positions stay right, and `get_model` answers them with a clear error instead of an IndexError.
"""

OUT_OF_SCOPE = ('ItemKNN', 'MultiVAE', 'IDCF_LGCN', 'NeuMF')
_SLOTS = ('MF', 'LightGCN', 'IGCN', 'ItemKNN', 'NGCF', 'MultiVAE', 'IMF', 'IMCGAE', 'IDCF_LGCN', 'NeuMF')

# per dataset: MF (lr, l2), LightGCN l2, IGCN dropout, IMF (dropout, aux_reg), NGCF (dropout, l2), IMCGAE dropout
_HYPER = {
    'gowalla': {'path': 'data/Gowalla/time', 'mf': (1.e-4, 1.e-3), 'lgcn_l2': 1.e-4, 'igcn_dropout': 0.3, 'imf': (0.1, 0.1),
                'ngcf': (0.1, 1.e-3), 'imcgae': 0.3},
    'yelp': {'path': 'data/Yelp/time', 'mf': (1.e-3, 1.e-3), 'lgcn_l2': 1.e-4, 'igcn_dropout': 0.3, 'imf': (0.5, 0.01),
             'ngcf': (0.3, 1.e-3), 'imcgae': 0.3},
    'amazon': {'path': 'data/Amazon/time', 'mf': (1.e-3, 1.e-4), 'lgcn_l2': 1.e-5, 'igcn_dropout': 0.0, 'imf': (0.3, 0.1),
               'ngcf': (0.3, 1.e-4), 'imcgae': 0.9},
}


def _table(device, dataset_config, h):
    def trainer(name, **kw):
        cfg = {'name': name, 'optimizer': 'Adam', 'lr': 1.e-3, 'device': device, 'n_epochs': 1000, 'batch_size': 2048,
               'dataloader_num_workers': 6, 'test_batch_size': 512, 'topks': [20]}
        cfg.update(kw)
        return cfg

    emb = {'embedding_size': 64, 'device': device}
    entries = {
        'MF': (dict(emb, name='MF'), trainer('BPRTrainer', lr=h['mf'][0], l2_reg=h['mf'][1])),
        'LightGCN': (dict(emb, name='LightGCN', n_layers=3), trainer('BPRTrainer', l2_reg=h['lgcn_l2'])),
        'IGCN': (dict(emb, name='IGCN', n_layers=3, dropout=h['igcn_dropout'], feature_ratio=1.),
                 trainer('IGCNTrainer', l2_reg=0., aux_reg=0.01)),
        'IMF': (dict(emb, name='IMF', n_layers=0, dropout=h['imf'][0], feature_ratio=1.),
                trainer('IGCNTrainer', l2_reg=1.e-5, aux_reg=h['imf'][1])),
        'NGCF': (dict(emb, name='NGCF', layer_sizes=[64, 64, 64], dropout=h['ngcf'][0]), trainer('BPRTrainer', l2_reg=h['ngcf'][1])),
        'IMCGAE': (dict(emb, name='IMCGAE', n_layers=3, dropout=h['imcgae']), trainer('BPRTrainer', l2_reg=0.)),
    }
    out = []
    for name in _SLOTS:
        model_config, trainer_config = entries.get(name, ({'name': name, 'device': device, 'out_of_scope': True},
                                                          {'name': 'BasicTrainer', 'device': device, 'n_epochs': 0,
                                                           'test_batch_size': 512, 'topks': [20]}))
        out.append((dataset_config, model_config, trainer_config))
    return out


def _named(device, key):
    h = _HYPER[key]
    return _table(device, {'name': 'ProcessedDataset', 'path': h['path'], 'device': device}, h)


def get_gowalla_config(device):
    """reference config.py:1-73."""
    return _named(device, 'gowalla')


def get_yelp_config(device):
    """reference config.py:76-148."""
    return _named(device, 'yelp')


def get_amazon_config(device):
    """reference config.py:151-223."""
    return _named(device, 'amazon')


def get_synthetic_config(device, shape='gowalla', seed=2021):
    """Same slots and hyper-parameters on an in-memory synthetic graph of the named shape (the real datasets are
    not available offline)."""
    h = _HYPER.get(shape, _HYPER['gowalla'])
    return _table(device, {'name': 'SyntheticDataset', 'shape': shape, 'seed': seed, 'device': device}, h)
