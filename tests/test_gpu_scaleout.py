"""Device-resident graph path (BASELINE.json config 5 at test size): the CSR generated on the GPU must be the
same graph the host path builds from its pairs, and IGCN propagation + unmasked full ranking on it must equal
the list-based path on identical weights."""
import numpy as np
import pytest
import torch

from conftest import rel_err

pytestmark = pytest.mark.gpu
DEV = torch.device('cuda:0')


@pytest.fixture(scope='module')
def device_ds():
    from igcn_cf_b200.dataset import get_dataset
    return get_dataset({'name': 'DeviceSyntheticDataset', 'shape': (3000, 4000, 90000), 'seed': 5, 'device': DEV})


def _pairs(dg):
    rp = dg.rowptr_host
    users = np.repeat(np.arange(dg.n_users, dtype=np.int64), np.diff(rp[:dg.n_users + 1]))
    items = dg.col[:dg.n_interactions].cpu().numpy().astype(np.int64) - dg.n_users
    return np.stack([users, items], axis=1)


def test_device_graph_is_the_host_graph(device_ds):
    from igcn_cf_b200 import graph
    dg = device_ds.device_graph
    pairs = _pairs(dg)
    assert len(np.unique(pairs[:, 0] * dg.n_items + pairs[:, 1])) == len(pairs)          # no duplicate interactions
    assert np.diff(dg.rowptr_host[:dg.n_users + 1]).min() >= 1 and 0.9 * 90000 < len(pairs) <= 90000 * 1.05
    host = graph.NormAdj(dg.n_users, dg.n_items, pairs, DEV)
    dev = graph.NormAdj.from_device(dg)
    assert np.array_equal(host.csr.rowptr_host, dev.csr.rowptr_host)
    assert torch.equal(host.csr.col, dev.csr.col)
    assert rel_err(dev.csr.val.cpu().numpy(), host.csr.val.cpu().numpy()) < 1e-6
    # a row block of the sharded constructor is a slice of the whole
    part = graph.NormAdj.from_device(dg, shard=(1, 3))
    assert len(part.blocks) == 2 and part.blocks[0].row1 <= dg.n_users <= part.blocks[1].row0
    for b in part.blocks:
        lo, hi = int(dev.csr.rowptr_host[b.row0]), int(dev.csr.rowptr_host[b.row1])
        assert torch.equal(b.csr.col, dev.csr.col[lo:hi]) and torch.equal(b.csr.val, dev.csr.val[lo:hi])


def test_igcn_on_device_graph_matches_list_path(device_ds):
    from igcn_cf_b200.dataset import get_dataset
    from igcn_cf_b200.model import get_model
    from igcn_cf_b200.trainer import BasicTrainer
    dg = device_ds.device_graph
    pairs = _pairs(dg)
    lists = [[] for _ in range(dg.n_users)]
    for u, i in pairs.tolist():
        lists[u].append(i)
    empty = [[] for _ in range(dg.n_users)]
    list_ds = get_dataset({'name': 'ListDataset', 'train': lists, 'val': empty, 'test': empty, 'n_items': dg.n_items,
                           'device': DEV})
    cfg = {'name': 'IGCN', 'embedding_size': 64, 'n_layers': 3, 'device': DEV, 'dropout': 0.3, 'feature_ratio': 1.}
    torch.manual_seed(3)
    a = get_model(cfg, device_ds)
    torch.manual_seed(3)
    b = get_model(cfg, list_ds)
    assert torch.equal(a.embedding.weight, b.embedding.weight)
    a.eval(), b.eval()
    with torch.no_grad():
        ra, rb = a.get_rep(), b.get_rep()
    assert rel_err(ra.cpu().numpy(), rb.cpu().numpy()) < 1e-5
    tcfg = {'name': 'BasicTrainer', 'device': DEV, 'n_epochs': 0, 'topks': [20], 'test_batch_size': 512}
    ta = BasicTrainer(dict(tcfg, dataset=device_ds, model=a))
    tb = BasicTrainer(dict(tcfg, dataset=list_ds, model=b))
    ia, sa = ta.recommend('train')
    ib, sb = tb.recommend('train')
    same = (ia == ib).float().mean().item()
    assert same > 0.999, same                      # rep differs by ~1e-7: only exact ties may swap
    assert rel_err(sa.cpu().numpy(), sb.cpu().numpy()) < 1e-5
