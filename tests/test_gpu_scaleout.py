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


def test_column_blocked_item_rows_match_the_unblocked_layer(device_ds, monkeypatch):
    """graph.column_blocks (tables far larger than L2; forced here on a small graph): the item rows run once per
    column range, each pass adding the previous partial sums -- same representation as the unblocked kernels to fp32
    summation-order noise, identical whatever the row sharding (ranges are global column intervals), and the fused
    ranking on top returns the same lists."""
    from igcn_cf_b200 import engine, graph
    from igcn_cf_b200.model import get_model
    dg = device_ds.device_graph
    cfg = {'name': 'IGCN', 'embedding_size': 64, 'n_layers': 3, 'device': DEV, 'dropout': 0.3, 'feature_ratio': 1.}
    torch.manual_seed(3)
    plain = get_model(cfg, device_ds)
    assert all(b.col_blocks is None for b in plain.norm_adj.blocks)
    monkeypatch.setattr(graph, 'COL_BLOCK_BYTES', 700 * 256)            # 700 users per range: 5 ranges over 3000 users
    monkeypatch.setattr(graph, 'COL_BLOCK_MIN_TABLE', 0)
    torch.manual_seed(3)
    blocked = get_model(cfg, device_ds)
    adj = blocked.norm_adj
    assert [(b.row0, b.row1) for b in adj.blocks] == [(0, dg.n_users), (dg.n_users, dg.n_users + dg.n_items)]
    cbs = adj.blocks[1].col_blocks
    assert adj.blocks[0].col_blocks is None and len(cbs) == 5
    # the ranges partition the item rows' entries: same multiset per row, columns inside the range, ascending
    full = adj.blocks[1].csr
    assert sum(c.nnz for c in cbs) == full.nnz
    deg_sum = np.zeros(full.n_rows, dtype=np.int64)
    for b, c in enumerate(cbs):
        assert int(c.col.min()) >= b * 700 and int(c.col.max()) < (b + 1) * 700
        deg_sum += np.diff(c.rowptr_host)
    assert np.array_equal(deg_sum, np.diff(full.rowptr_host))
    assert torch.equal(blocked.embedding.weight, plain.embedding.weight)
    plain.eval(); blocked.eval()
    with torch.no_grad():
        ra, rb = plain.get_rep(), blocked.get_rep()
    assert rel_err(rb.cpu().numpy(), ra.cpu().numpy()) < 1e-6
    # GPU-count independence: the item slice of a 3-way row sharding cut at the same global ranges gives the same rows
    part = graph.NormAdj.from_device(dg, shard=(1, 3))
    blk = part.blocks[1]
    assert blk.col_blocks is not None and len(blk.col_blocks) == 5
    prop = engine.Propagator(dg.n_users + dg.n_items, 64, 1, DEV)
    x = ra.contiguous()
    y_full = torch.zeros_like(x)
    y_part = torch.zeros_like(x)
    prop.spmm(adj, x, y_full)
    prop.spmm(part, x, y_part)
    assert torch.equal(y_part[blk.row0:blk.row1], y_full[blk.row0:blk.row1])
    assert rel_err(y_full.cpu().numpy(), torch.sparse.mm(plain.norm_adj.to_sparse_coo().double(), x.double()).cpu().numpy()) < 1e-6
