"""Worker for the multi-rank tests; launched as
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
        tests/dist_worker.py {cpu|gpu}

cpu: gloo, world_size N -- host-side sharding logic only (row bounds, user split, list gather).
gpu: nccl, one rank per GPU -- the row-sharded propagation with fused peer-store all-gather, the
     replicated BPR step and the user-sharded evaluation must be BIT-IDENTICAL to the single-GPU path
     run in the same process (`'shard': False`); so must the column-sharded training step (`'shard': 'dims'`:
     every rank owns embedding_size / world columns, per-triple partial sums exchanged, parameters all-gathered)."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from igcn_cf_b200 import dist as idist  # noqa: E402
from igcn_cf_b200 import synth  # noqa: E402


def check_host_logic(rank, world):
    split = synth.gen_named('tiny', seed=7)
    ptr, items = split.csr('train')
    deg_u = np.diff(ptr)
    deg_i = np.bincount(items, minlength=split.n_items)
    rowptr = np.concatenate([[0], np.cumsum(np.concatenate([deg_u, deg_i]))])
    bounds = idist.shard_bounds(rowptr, world)
    assert bounds[0] == 0 and bounds[-1] == len(rowptr) - 1 and np.all(np.diff(bounds) >= 0)
    cost = np.diff(rowptr[bounds]) + 4 * np.diff(bounds)
    assert cost.max() <= cost.sum() / world + rowptr[1:].max() - 0 + 8, cost      # balanced up to one row
    # every rank computes the same bounds
    got = [None] * world
    dist.all_gather_object(got, bounds.tolist())
    assert all(g == got[0] for g in got)
    # user split covers [0, n) exactly once
    n = 1001
    pieces = [idist.split_range(n, r, world) for r in range(world)]
    assert pieces[0][0] == 0 and pieces[-1][1] == n
    assert all(pieces[r][1] == pieces[r + 1][0] for r in range(world - 1))
    # gather_rows reassembles per-rank blocks in order
    full = torch.arange(n * 3, dtype=torch.int32).reshape(n, 3)
    lo, hi = pieces[rank]
    out = idist.gather_rows(full[lo:hi].clone(), n)
    assert torch.equal(out, full)


def check_gpu(rank, world):
    from igcn_cf_b200.dataset import get_dataset
    from igcn_cf_b200.model import get_model
    from igcn_cf_b200.trainer import get_trainer
    dev = torch.device('cuda', int(os.environ.get('LOCAL_RANK', rank)))
    peers = idist.init_peers()
    assert peers is not None and peers.world == world
    split = synth.gen_named('small', seed=2021)
    ds = get_dataset({'name': 'SyntheticDataset', 'split': split, 'device': dev})
    for kind in ('LightGCN', 'IGCN'):
        models, trainers = [], []
        for shard in (True, False):
            torch.manual_seed(5)
            mcfg = {'name': kind, 'embedding_size': 64, 'n_layers': 3, 'device': dev, 'shard': shard}
            tcfg = {'optimizer': 'Adam', 'lr': 1e-3, 'l2_reg': 1e-4, 'device': dev, 'n_epochs': 1, 'batch_size': 2048,
                    'dataloader_num_workers': 0, 'test_batch_size': 512, 'topks': [20], 'cuda_graph': True, 'seed': 11,
                    'name': 'BPRTrainer'}
            if kind == 'IGCN':
                mcfg.update(dropout=0.3, feature_ratio=1.)
                tcfg.update(name='IGCNTrainer', aux_reg=0.01, l2_reg=0.)
            m = get_model(mcfg, ds)
            models.append(m)
            trainers.append(get_trainer(tcfg, ds, m))
        sharded, single = models
        assert sharded._peers is peers and single._peers is None
        assert len(sharded.norm_adj.blocks) == 2 and len(single.norm_adj.blocks) == 1
        assert torch.equal(sharded.embedding.weight, single.embedding.weight)
        # eval-mode representation
        for m in models:
            m.eval()
        with torch.no_grad():
            assert torch.equal(sharded.get_rep(), single.get_rep()), kind + ' rep differs'
        # bulk form of the exchange (what the scale-out graph uses): every finished row block leaves through
        # igcn_peer_push on a side stream while the next block is computed
        from igcn_cf_b200 import engine as _engine
        keep_bytes, _engine.Shard.PUSH_BYTES = _engine.Shard.PUSH_BYTES, 0
        sharded._bump()
        with torch.no_grad():
            assert torch.equal(sharded.get_rep(), single.get_rep()), kind + ' rep differs (bulk push)'
        _engine.Shard.PUSH_BYTES = keep_bytes
        # autograd bridge
        for m in models:
            m.train()
        grads = []
        for m in models:
            m.zero_grad()
            if kind == 'IGCN':
                m._drop_calls = 0
            rep = m.get_rep()
            (rep * rep).sum().backward()
            grads.append(m.embedding.weight.grad.clone())
        assert torch.equal(grads[0], grads[1]), kind + ' autograd gradient differs'
        # fused training steps (CUDA graph, device sampler, hash dropout) then evaluation
        losses = []
        for t in trainers:
            t.model.train()
            ls = [t.step.run().item() for _ in range(5)]
            losses.append(ls)
        assert losses[0] == losses[1], (kind, losses)
        assert torch.equal(sharded.embedding.weight, single.embedding.weight), kind + ' weights differ after 5 steps'
        r0, m0 = trainers[0].eval('val')
        r1, m1 = trainers[1].eval('val')
        assert r0 == r1 and m0 == m1, (r0, r1)
        # column-sharded training against the same single-GPU run: 5 more steps on both, then weights / loss / eval
        torch.manual_seed(5)
        mcfg = dict(models[1].config, shard='dims')
        mcfg.pop('dataset')
        dm = get_model(mcfg, ds)
        tcfg = {k: v for k, v in trainers[1].config.items() if k not in ('dataset', 'model')}
        dt = get_trainer(tcfg, ds, dm)
        assert dt.step.dims == (rank, world) and dt.step.D == 64 // world
        with torch.no_grad():
            dm.embedding.weight.copy_(single.embedding.weight)
            if kind == 'IGCN':
                dm.w.copy_(single.w)
                dm.alpha = single.alpha
                dm.update_feat_mat()
        dm._bump()
        # same optimizer state and step counter as the single-GPU trainer (moments in column slices)
        c0, c1 = dt.step.col0, dt.step.col0 + dt.step.D
        m_s, v_s = trainers[1].opt.moments(single.embedding.weight)
        dt.step.emb_m.copy_(m_s[:, c0:c1]); dt.step.emb_v.copy_(v_s[:, c0:c1])
        if kind == 'IGCN':
            wm, wv = trainers[1].opt.moments(single.w)
            dt.step.w_m.copy_(wm[c0:c1]); dt.step.w_v.copy_(wv[c0:c1])
        dt.step.state.copy_(trainers[1].step.state)
        dt.opt.t = trainers[1].opt.t
        dm.train(); single.train()
        l_d = [dt.step.run().item() for _ in range(5)]
        l_s = [trainers[1].step.run().item() for _ in range(5)]
        assert l_d == l_s, (kind, 'dims losses', l_d, l_s)
        dt.step.sync_params()
        assert torch.equal(dm.embedding.weight, single.embedding.weight), kind + ' weights differ (dims)'
        if kind == 'IGCN':
            assert torch.equal(dm.w, single.w)
        # evaluation-mode representation of the column-sharded model: propagated per column slice, all-gathered
        dm.eval(); single.eval()
        with torch.no_grad():
            rep_d = dm.get_rep()
            assert dm._col_rep is not None and dm._col_rep.D == 64 // world
            assert torch.equal(rep_d, single.get_rep()), kind + ' column-sharded eval rep differs'
        r2, m2 = dt.eval('val')
        r3, m3 = trainers[1].eval('val')
        assert r2 == r3 and m2 == m3, (r2, r3)
        # the default ('auto') picks the column mode on paper-sized graphs and trains through the public epoch loop
        torch.manual_seed(5)
        am = get_model({k: v for k, v in models[1].config.items() if k not in ('dataset', 'shard')}, ds)
        at = get_trainer(tcfg, ds, am)
        assert am._dim_shard == (rank, world) and not am._rows_sharded() and at.step.dims == (rank, world)
        am.train()
        loss_auto = at.train_one_epoch()
        assert np.isfinite(loss_auto) and not at.step._dirty                 # the epoch ends with the parameter all-gather
        gathered = [None] * world
        dist.all_gather_object(gathered, float(am.embedding.weight.detach().double().sum().item()))
        assert all(g == gathered[0] for g in gathered), gathered            # every rank holds the same full-width weights
        peers.check()
        if rank == 0:
            print('dist_worker: %s ok, world %d, %d barriers, %s' % (kind, world, peers.n_barriers, r0))
    torch.cuda.synchronize()


def main():
    mode = sys.argv[1]
    rank, world = int(os.environ['RANK']), int(os.environ['WORLD_SIZE'])
    if mode == 'cpu':
        dist.init_process_group('gloo')
        check_host_logic(rank, world)
    else:
        local = int(os.environ.get('LOCAL_RANK', rank))
        torch.cuda.set_device(local)
        dist.init_process_group('nccl', device_id=torch.device('cuda', local))
        check_gpu(rank, world)
        idist.shutdown()
    dist.barrier()
    dist.destroy_process_group()
    if rank == 0:
        print('dist_worker: all ok')


if __name__ == '__main__':
    main()
