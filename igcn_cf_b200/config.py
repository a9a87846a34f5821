"""Hot-path configurations of the reference (config.py entries [1] LightGCN and [2] IGCN of each
dataset: reference config.py:12-23, 87-98, 162-173) plus synthetic-graph variants of the same
hyper-parameters, since the real datasets are not available offline."""


def _pair(device, path, lgcn_l2, igcn_dropout):
    dataset_config = {'name': 'ProcessedDataset', 'path': path, 'device': device}
    common = {'optimizer': 'Adam', 'lr': 1.e-3, 'device': device, 'n_epochs': 1000, 'batch_size': 2048,
              'dataloader_num_workers': 6, 'test_batch_size': 512, 'topks': [20]}
    lgcn = ({'name': 'LightGCN', 'embedding_size': 64, 'n_layers': 3, 'device': device},
            dict(common, name='BPRTrainer', l2_reg=lgcn_l2))
    igcn = ({'name': 'IGCN', 'embedding_size': 64, 'n_layers': 3, 'device': device, 'dropout': igcn_dropout,
             'feature_ratio': 1.},
            dict(common, name='IGCNTrainer', l2_reg=0., aux_reg=0.01))
    return [(dataset_config, lgcn[0], lgcn[1]), (dataset_config, igcn[0], igcn[1])]


def get_gowalla_config(device):
    """[LightGCN, IGCN] triples (reference config.py:12-23)."""
    return _pair(device, 'data/Gowalla/time', 1.e-4, 0.3)


def get_yelp_config(device):
    """reference config.py:87-98."""
    return _pair(device, 'data/Yelp/time', 1.e-4, 0.3)


def get_amazon_config(device):
    """reference config.py:162-173 (IGCN dropout 0.0, LightGCN l2 1e-5)."""
    return _pair(device, 'data/Amazon/time', 1.e-5, 0.0)


def get_synthetic_config(device, shape='gowalla', seed=2021):
    """Same hyper-parameters on an in-memory synthetic graph of the named shape."""
    base = {'gowalla': get_gowalla_config, 'yelp': get_yelp_config, 'amazon': get_amazon_config}.get(
        shape, get_gowalla_config)(device)
    ds = {'name': 'SyntheticDataset', 'shape': shape, 'seed': seed, 'device': device}
    return [(ds, m, t) for _, m, t in base]
