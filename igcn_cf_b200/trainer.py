"""Drop-in trainer layer: `BasicTrainer`, `BPRTrainer`, `IGCNTrainer` with the reference's public
surface (reference trainer.py:14-219, 222-248, 281-320).

Same: class names and config keys, `train / train_one_epoch / eval / inductive_eval /
calculate_metrics / record / initialize_optimizer`, the returned result string and metric dict,
checkpoint naming and early stopping.

Different (B200-native): `train_one_epoch` drives igcn_cf_b200.engine.TrainStep -- sampler,
propagation, fused BPR loss, deterministic gradient, backward propagation and Adam are sm_100a
kernels, optionally replayed as one CUDA graph, and the running loss stays on the device until the
epoch ends (the reference syncs on loss.item() every step, trainer.py:247/318).  `eval` computes the
representation once and runs the fused score + mask + top-k kernel over all users (the reference
re-propagates per 512-user batch and round-trips an [512, n_items] score matrix, trainer.py:145-164).

Trainer config extras (all optional): 'sampler': 'device' (default) | 'reference' (host
DataLoader, same draw sequence as the reference), 'cuda_graph': bool (default True), 'seed': int.
"""
import os
import sys
import time

import numpy as np
import torch
import torch.nn.functional as F
from torch.utils.data import DataLoader

from . import dist, engine
from ._lib import call, ptr, stream_ptr
from .dataset import AuxiliaryDataset
from .utils import AverageMeter  # noqa: F401  (re-exported like the reference module does)


def get_trainer(config, dataset, model):
    """Name-dispatched factory (trainer.py:14-20)."""
    config = config.copy()
    config['dataset'] = dataset
    config['model'] = model
    cls = getattr(sys.modules[__name__], config['name'])
    return cls(config)


class Adam(torch.optim.Optimizer):
    """torch.optim.Adam semantics (default betas/eps, no weight decay) on the igcn_adam kernel.

    Selected by name from config['optimizer'] like the reference does (trainer.py:43-45).  The
    moments and the device-resident step counter are shared with the fused TrainStep, so mixing
    `opt.step()` (generic autograd path) and fused steps keeps one optimizer state."""

    def __init__(self, params, lr=1e-3):
        super().__init__(params, dict(lr=lr, betas=(engine.BETA1, engine.BETA2), eps=engine.ADAM_EPS))
        self.t = 0
        self.step_state = None      # igcn_step_state on the device, created lazily
        self.on_step = None         # called after every update (the trainer hooks the model's eval-cache invalidation)

    def moments(self, p):
        st = self.state[p]
        if 'exp_avg' not in st:
            st['exp_avg'] = torch.zeros_like(p.data)
            st['exp_avg_sq'] = torch.zeros_like(p.data)
        return st['exp_avg'], st['exp_avg_sq']

    def device_state(self, device):
        if self.step_state is None:
            self.step_state = torch.zeros(16, dtype=torch.uint8, device=device)
        return self.step_state

    @torch.no_grad()
    def step(self, closure=None):
        self.t += 1
        for group in self.param_groups:
            b1, b2 = group['betas']
            for p in group['params']:
                if p.grad is None:
                    continue
                m, v = self.moments(p)
                g = p.grad.contiguous()
                call('igcn_adam', ptr(p.data), ptr(g), ptr(m), ptr(v), p.numel(), group['lr'], b1, b2, group['eps'],
                     self.t, None, stream_ptr())
        if self.step_state is not None:
            self._sync_device_step()
        if self.on_step is not None:
            # igcn_adam wrote through raw device pointers: autograd's tensor versions did not move, so whoever
            # caches values derived from the parameters (the eval-mode representation) has to be told
            self.on_step()

    def _sync_device_step(self):
        self.step_state[:8].copy_(torch.tensor([self.t], dtype=torch.int64).view(torch.uint8))


class BestCheckpoint:
    """Best-validation checkpoint + patience policy of the reference's epoch loop (trainer.py:66, 90-106): a better
    NDCG replaces the previous best file `checkpoints/<model>_<trainer>_<dataset>_<ndcg*100:.3f>.pth`, anything
    else costs `val_interval` epochs of patience.

    With one process per GPU every rank sees the same gathered metrics (replicas are bit-identical), so only rank 0
    touches the file system -- through a temporary file and os.replace, so a reader never sees a torn file -- and
    all ranks meet at a barrier before anyone reloads the best file."""

    def __init__(self, trainer, directory):
        self.tr, self.dir = trainer, directory
        self.patience = trainer.max_patience
        self.writer_rank = self._rank() == 0
        if self.writer_rank:
            os.makedirs(directory, exist_ok=True)

    @staticmethod
    def _rank():
        import torch.distributed as td
        return td.get_rank() if td.is_available() and td.is_initialized() else 0

    @staticmethod
    def _barrier():
        import torch.distributed as td
        if td.is_available() and td.is_initialized() and td.get_world_size() > 1:
            td.barrier()

    def offer(self, ndcg):
        """Record one validation result; False = patience exhausted."""
        tr = self.tr
        if ndcg > tr.best_ndcg:
            old = tr.save_path
            tr.save_path = os.path.join(self.dir, '{:s}_{:s}_{:s}_{:.3f}.pth'.format(
                tr.model.name, tr.name, tr.dataset.name, ndcg * 100))
            tr.best_ndcg = ndcg
            if self.writer_rank:
                tmp = tr.save_path + '.tmp'
                tr.model.save(tmp)
                os.replace(tmp, tr.save_path)
                if old and old != tr.save_path and os.path.exists(old):
                    os.remove(old)
            self.patience = tr.max_patience
            print('Best NDCG, save model to {:s}'.format(tr.save_path))
            return True
        self.patience -= tr.val_interval
        return self.patience > 0

    def restore(self):
        self._barrier()                 # the writer has finished before anyone reads
        self.tr.model.load(self.tr.save_path)
        self._barrier()                 # nobody removes / rewrites the file while another rank still reads it


class BasicTrainer:
    """trainer.py:23-219."""

    def __init__(self, trainer_config):
        print(trainer_config)
        self.config = trainer_config
        self.name = trainer_config['name']
        self.dataset = trainer_config['dataset']
        self.model = trainer_config['model']
        self.topks = trainer_config['topks']
        self.device = trainer_config['device']
        self.n_epochs = trainer_config['n_epochs']
        self.max_patience = trainer_config.get('max_patience', 50)
        self.val_interval = trainer_config.get('val_interval', 1)
        self.test_batch_size = trainer_config.get('test_batch_size', 512)
        self.epoch = 0
        self.best_ndcg = -np.inf
        self.save_path = None
        self.opt = None
        self._mask_cache = {}
        self._eval_cache = {}
        self._last_rec = None
        self._item_order = None
        self._metric_tables = {}
        self.test_users = torch.arange(self.dataset.n_users, dtype=torch.int64, device=self.device)

    def initialize_optimizer(self):
        name = self.config['optimizer']
        if name != 'Adam':
            raise NotImplementedError('the B200 hot path implements Adam (every shipped config uses it); got ' + name)
        self.opt = Adam(self.model.parameters(), lr=self.config['lr'])
        bump = getattr(self.model, '_bump', None)
        if callable(bump):
            self.opt.on_step = bump

    def train_one_epoch(self):
        raise NotImplementedError

    def record(self, writer, stage, metrics):
        """tensorboard scalars, tag layout of trainer.py:50-55."""
        prefix = '{:s}_{:s}/{:s}'.format(self.model.name, self.name, stage)
        for metric, per_k in metrics.items():
            for k in self.topks:
                writer.add_scalar('{:s}_{:s}@{:d}'.format(prefix, metric, k), per_k[k], self.epoch)

    def _timed(self, fn):
        t0 = time.time()
        out = fn()
        return out, time.time() - t0

    def _check_peers(self):
        """Row-sharded runs: a device barrier that timed out must stop the run, not be trained through."""
        peers = getattr(self.model, '_peers', None)
        if peers is not None:
            peers.check()

    def train(self, verbose=True, writer=None):
        """Epoch loop of trainer.py:57-107: one training epoch, eval('train') every epoch, eval('val') every
        `val_interval` epochs, keep the checkpoint of the best validation NDCG@topks[0], stop after `max_patience`
        epochs without improvement, reload the best checkpoint.  The printed lines, the tensorboard tags and the
        checkpoint file names are the reference's (they are what downstream scripts parse); the policy lives in
        BestCheckpoint, which is also what makes the loop safe with one process per GPU."""
        say = print if verbose else (lambda *a, **k: None)
        if not self.model.trainable:
            results, metrics = self.eval('val')
            say('Validation result. {:s}'.format(results))
            return metrics['NDCG'][self.topks[0]]

        keeper = BestCheckpoint(self, 'checkpoints')
        tag = '{:s}_{:s}'.format(self.model.name, self.name)
        for self.epoch in range(self.n_epochs):
            def one_epoch():
                self.model.train()
                loss = self.train_one_epoch()
                return loss, self.eval('train')[1]
            (loss, metrics), seconds = self._timed(one_epoch)
            self._check_peers()
            say('Epoch {:d}/{:d}, Loss: {:.6f}, Time: {:.3f}s'.format(self.epoch, self.n_epochs, loss, seconds))
            if writer:
                writer.add_scalar(tag + '/train_loss', loss, self.epoch)
                self.record(writer, 'train', metrics)
            if (self.epoch + 1) % self.val_interval:
                continue
            (results, metrics), seconds = self._timed(lambda: self.eval('val'))
            say('Validation result. {:s}Time: {:.3f}s'.format(results, seconds))
            if writer:
                self.record(writer, 'validation', metrics)
            if not keeper.offer(metrics['NDCG'][self.topks[0]]):
                print('Early stopping!')
                break
        keeper.restore()
        return self.best_ndcg

    # ---- metrics (trainer.py:109-138)
    def _eval_csr(self, eval_data):
        """Sorted CSR of an eval list-of-lists; rebuilt when any inner list object or length changed
        (the reference's inductive_eval swaps inner lists in place, trainer.py:185-217).  An engine.ListCSR is
        passed through (our inductive_eval restricts on arrays)."""
        if isinstance(eval_data, engine.ListCSR):
            return eval_data
        sig = (tuple(map(id, eval_data)), tuple(map(len, eval_data)))
        hit = self._eval_cache.get(id(eval_data))
        if hit is None or hit[0] != sig:
            hit = (sig, engine.lists_to_csr(eval_data, self.device))
            self._eval_cache[id(eval_data)] = hit
        return hit[1]

    def _dataset_csr(self, which):
        """numpy CSR of dataset.<which>_data: the dataset's cached arrays when it offers them."""
        get = getattr(self.dataset, 'csr', None)
        if callable(get):
            return get(which)
        return engine.lists_to_arrays(getattr(self.dataset, which + '_data'))

    def calculate_metrics(self, eval_data, rec_items):
        """Precision / Recall / NDCG @k with the reference's dtypes and bits (trainer.py:109-138): fp32 hits and log2
        table, int32 list lengths, users without eval items excluded from the means.

        The [U, k] part -- membership of every recommended item, per-user hit counts and DCGs -- runs on the device
        (igcn_user_metrics: numpy's own float32 row-sum order, so the per-user values are the reference's bit for
        bit); only 2 x len(topks) x U floats come back, and the remaining vector expressions are the reference's,
        evaluated by numpy on [U] arrays.  rec_items: numpy array or device tensor [U, max(topks)]."""
        results = {'Precision': {}, 'Recall': {}, 'NDCG': {}}
        csr = self._eval_csr(eval_data)
        if torch.is_tensor(rec_items) and rec_items.is_cuda:
            rec_dev = rec_items.to(torch.int32).contiguous()
        elif self._last_rec is not None and self._last_rec[0] is rec_items:
            rec_dev = self._last_rec[1]
        else:
            rec_dev = torch.as_tensor(np.ascontiguousarray(rec_items), device=self.device).to(torch.int32).contiguous()
        n, k_rec = int(rec_dev.shape[0]), int(rec_dev.shape[1])
        eval_data_len = csr.lens.astype(np.int32)
        topks = list(self.topks)
        if k_rec <= 32 and len(topks) <= 8 and max(topks) <= k_rec:
            per_user = self._device_user_metrics(rec_dev, csr, topks, k_rec)
        else:                                           # wide lists: the [U, k] hit matrix comes to the host
            hit_matrix = engine.hit_matrix(rec_dev, csr).cpu().numpy()
            per_user = {}
            for k in topks:
                denominator = np.log2(np.arange(2, k + 2, dtype=np.float32))[None, :]
                per_user[k] = (np.sum(hit_matrix[:, :k], axis=1), np.sum(hit_matrix[:, :k] / denominator, axis=1))
        for k in topks:
            hit_num, dcgs = per_user[k]
            precisions = hit_num / k
            with np.errstate(invalid='ignore', divide='ignore'):
                recalls = hit_num / eval_data_len
            max_hit_num = np.minimum(eval_data_len, k)
            denominator = np.log2(np.arange(2, k + 2, dtype=np.float32))[None, :]
            # the ideal DCG of a user only depends on min(len, k): evaluate the reference's expression
            # (trainer.py:126-130) once per possible value -- same numpy row reduction, so the same bits -- and
            # look it up instead of building a second [U, k] matrix
            ideal_rows = (np.arange(k)[None, :] < np.arange(k + 1)[:, None]).astype(np.float32)
            idcgs = np.sum(ideal_rows / denominator, axis=1)[max_hit_num]
            with np.errstate(invalid='ignore', divide='ignore'):
                ndcgs = dcgs / idcgs
            user_masks = (max_hit_num > 0)
            results['Precision'][k] = precisions[user_masks].mean()
            results['Recall'][k] = recalls[user_masks].mean()
            results['NDCG'][k] = ndcgs[user_masks].mean()
        return results

    def _device_user_metrics(self, rec_dev, csr, topks, k_rec):
        """{k: (hit_num fp32 [U], dcg fp32 [U])} from igcn_user_metrics; one D2H copy of 2 x len(topks) x U floats."""
        import ctypes as C
        n = int(rec_dev.shape[0])
        table = self._metric_tables.get(k_rec)
        if table is None:
            # the reference's fp32 table, computed by numpy on the host exactly as trainer.py:128 does
            table = torch.from_numpy(np.log2(np.arange(2, k_rec + 2, dtype=np.float32))).to(self.device)
            self._metric_tables[k_rec] = table
        out = torch.empty((2, len(topks), n), dtype=torch.float32, device=self.device)
        ks = (C.c_int32 * len(topks))(*topks)
        call('igcn_user_metrics', ptr(rec_dev), n, k_rec, ptr(csr.ptr), ptr(csr.items), ptr(table), ks, len(topks),
             ptr(out[0]), ptr(out[1]), stream_ptr())
        host = out.cpu().numpy()
        return {k: (host[0, t], host[1, t]) for t, k in enumerate(topks)}

    # ---- evaluation (trainer.py:140-177)
    def _mask_csr(self, val_or_test):
        if val_or_test == 'train':
            return None
        ds = self.dataset
        key = (val_or_test, id(ds.train_data), id(ds.val_data), len(ds.train_data))
        hit = self._mask_cache.get(val_or_test)
        if hit is None or hit[0] != key:
            arrays = self._dataset_csr('train')
            if val_or_test == 'test':
                arrays = engine.merge_csr(arrays, self._dataset_csr('val'))
            hit = (key, engine.ListCSR.from_arrays(arrays[0], arrays[1], self.device))
            self._mask_cache[val_or_test] = hit
        return hit[1]

    def _banned(self, banned_items):
        """banned_items (index array) -> (item_lo, item_hi, bitmap or None)."""
        n = self.dataset.n_items
        if banned_items is None:
            return 0, n, None
        b = np.unique(np.asarray(torch.as_tensor(banned_items).cpu()).astype(np.int64))
        if len(b) == 0:
            return 0, n, None
        if b[-1] - b[0] + 1 == len(b):              # contiguous: trim the scored range instead
            if b[0] == 0:
                return int(b[-1]) + 1, n, None
            if b[-1] == n - 1:
                return 0, int(b[0]), None
        flags = np.zeros(n, dtype=bool)
        flags[b[(b >= 0) & (b < n)]] = True
        from .graph import _pack_bits
        return 0, n, _pack_bits(flags, self.device)

    def item_order(self):
        """Scan order of the catalogue for the tensor-core ranking kernels: train popularity, most popular first
        (engine.ItemOrder).  Static per dataset; trainer_config['item_order'] = 'natural' disables it."""
        if self.config.get('item_order', 'popularity') != 'popularity':
            return None
        n_items = self.dataset.n_items
        if self._item_order is None or len(self._item_order.perm_host) != n_items:
            dg = getattr(self.dataset, 'device_graph', None)
            if dg is not None:
                deg = np.diff(dg.rowptr_host[dg.n_users:])
            else:
                from .graph import train_pairs_of
                deg = np.bincount(train_pairs_of(self.dataset)[:, 1], minlength=n_items)
            self._item_order = engine.ItemOrder.by_score(deg, self.device)
        return self._item_order

    def recommend(self, val_or_test, banned_items=None, users=None, users_host=None):
        """Top-max(topks) item ids (device int32 [n, k]) and scores for `users` (default: all).
        users_host: the same ids as a numpy array when the caller has them (saves a D2H sync)."""
        self.model.eval()
        with torch.no_grad():
            rep = self.model.get_rep().contiguous()
        if users is None:
            users, users_host = self.test_users, 'identity'
        lo, hi, bits = self._banned(banned_items)
        return engine.score_topk(rep, users, self.model.n_users, self.model.n_items, max(self.topks),
                                 mask=self._mask_csr(val_or_test), item_lo=lo, item_hi=hi, banned_bits=bits,
                                 users_host=users_host, impl=self.config.get('score_impl', 'auto'), order=self.item_order())

    def recommend_local(self, val_or_test, banned_items=None):
        """Top-k lists of the users this rank evaluates: all of them on one GPU, an even contiguous
        slice per rank when row-sharded (users are independent: no communication while scoring)."""
        peers = getattr(self.model, '_peers', None)
        if peers is None:
            return self.recommend(val_or_test, banned_items)[0]
        lo, hi = dist.split_range(self.dataset.n_users, peers.rank, peers.world)
        return self.recommend(val_or_test, banned_items, users=self.test_users[lo:hi],
                              users_host=np.arange(lo, hi, dtype=np.int64))[0]

    def eval(self, val_or_test, banned_items=None, eval_data=None):
        """trainer.py:140-177.  eval_data (extension): score against this engine.ListCSR / list-of-lists instead
        of dataset.<val_or_test>_data (masking still follows val_or_test)."""
        if eval_data is None:
            eval_data = getattr(self.dataset, val_or_test + '_data')
            if callable(getattr(self.dataset, 'csr', None)):
                # the dataset keeps numpy CSR arrays per list OBJECT (replacing dataset.<x>_data, as the
                # reference's inductive_eval does, is noticed; editing the inner lists in place is not)
                arrays = self.dataset.csr(val_or_test)
                hit = self._eval_cache.get(val_or_test)
                if hit is None or hit[0] is not arrays:
                    hit = (arrays, engine.ListCSR.from_arrays(arrays[0], arrays[1], self.device))
                    self._eval_cache[val_or_test] = hit
                eval_data = hit[1]
        peers = getattr(self.model, '_peers', None)
        rec_dev = self.recommend_local(val_or_test, banned_items)
        if peers is not None:
            # the top-k lists are gathered so that every rank reports the same metrics
            rec_dev = dist.gather_rows(rec_dev, self.dataset.n_users, peers.group)
        if 'calculate_metrics' in self.__dict__:
            # somebody wrapped calculate_metrics (the reference's signature takes the lists as a numpy array)
            rec_items = rec_dev.cpu().numpy()
            self._last_rec = (rec_items, rec_dev)
            metrics = self.calculate_metrics(eval_data, rec_items)
            self._last_rec = None
        else:
            metrics = self.calculate_metrics(eval_data, rec_dev)       # lists stay on the device

        return self.format_results(metrics), metrics

    def format_results(self, metrics):
        """'Precision: ..%@k, Recall: ..%@k, NDCG: ..%@k, ' -- the result line of trainer.py:169-176 (launchers and
        log parsers read it; the field order and the '{:.3f}%@{:d}, ' cells are the interface)."""
        cells = lambda name: ''.join('{:.3f}%@{:d}, '.format(metrics[name][k] * 100., k) for k in self.topks)
        return 'Precision: {:s}Recall: {:s}NDCG: {:s}'.format(cells('Precision'), cells('Recall'), cells('NDCG'))

    def inductive_eval(self, n_old_users, n_old_items):
        """Six test passes over user/item subsets (trainer.py:179-219).  The reference edits copies of
        dataset.test_data list by list; the same restrictions are applied here to the CSR arrays of the test
        lists and handed to eval() directly -- dataset.test_data itself is never touched."""
        ds = self.dataset
        n_u, n_i = ds.n_users, ds.n_items
        full = self._dataset_csr('test')
        ban_new, ban_old = np.arange(n_old_items, n_i), np.arange(n_old_items)
        plan = [('All users and all items', (0, n_u), (0, n_i), None),
                ('Old users and all items', (0, n_old_users), (0, n_i), None),
                ('New users and all items', (n_old_users, n_u), (0, n_i), None),
                ('All users and old items', (0, n_u), (0, n_old_items), ban_new),
                ('All users and new items', (0, n_u), (n_old_items, n_i), ban_old),
                ('Old users and old items', (0, n_old_users), (0, n_old_items), ban_new)]
        for title, users, items, banned in plan:
            ptr_, flat = engine.restrict_csr(full, users[0], users[1], items[0], items[1])
            results, _ = self.eval('test', banned_items=banned,
                                   eval_data=engine.ListCSR.from_arrays(ptr_, flat, self.device))
            print('{:s} result. {:s}'.format(title, results))


class AutogradStep:
    """One BPR step of BPRTrainer.train_one_epoch (trainer.py:233-247) for the sibling models that have no fused step
    (model.fused_step is False: NGCF, IMCGAE): the model's own bpr_forward under autograd -- every propagation hop is
    an autograd node on igcn_spmm (siblings.SpMM) --, the loss as the reference writes it, igcn_adam through
    Adam.step.  Triples come from the device sampler (igcn_sample_triples, same counter-based stream as the fused
    step) unless the caller passes them.  Same interface as engine.TrainStep as far as the epoch loop uses it; the
    running loss stays on the device (one read per epoch instead of the reference's loss.item() per step)."""

    dims = None

    def __init__(self, model, opt, l2_reg, batch_size=2048, seed=0):
        self.model, self.opt, self.l2_reg = model, opt, float(l2_reg)
        self.B, self.seed = int(batch_size), int(seed)
        dev = model.embedding.weight.device
        self.triples = torch.zeros((self.B, 3), dtype=torch.int64, device=dev)
        self.acc = torch.zeros(2, dtype=torch.float64, device=dev)
        self.loss = torch.zeros((), dtype=torch.float32, device=dev)

    def run(self, triples=None, aux_triples=None, drop='auto', batch=None):
        m = self.model
        if triples is None:
            B = self.B if batch is None else int(batch)
            rowptr, col = m.norm_adj.sampler_csr()
            call('igcn_sample_triples', ptr(rowptr), ptr(col), m.n_users, m.n_users, m.n_items, B, self.seed,
                 self.opt.t + 1, None, ptr(self.triples), stream_ptr())
            triples = self.triples[:B]
        users, pos_items, neg_items = triples[:, 0], triples[:, 1], triples[:, 2]
        users_r, pos_r, neg_r, l2_norm_sq = m.bpr_forward(users, pos_items, neg_items)
        pos_scores = torch.sum(users_r * pos_r, dim=1)
        neg_scores = torch.sum(users_r * neg_r, dim=1)
        loss = F.softplus(neg_scores - pos_scores).mean() + self.l2_reg * l2_norm_sq.mean()
        self.opt.zero_grad()
        loss.backward()
        self.opt.step()
        self.loss = loss.detach()
        n = triples.shape[0]
        self.acc[0] += self.loss.double() * n
        self.acc[1] += n
        return self.loss

    def sync_params(self):
        pass

    def reset_meter(self):
        self.acc.zero_()

    def meter_avg(self):
        s, n = self.acc.tolist()
        return s / max(n, 1.)


class _FusedBPRMixin:
    """Epoch loop shared by BPRTrainer and IGCNTrainer on top of engine.TrainStep."""

    def _init_fused(self, trainer_config, aux_reg=None):
        self.l2_reg = trainer_config['l2_reg']
        self.batch_size = trainer_config['batch_size']
        self.sampler = trainer_config.get('sampler', 'device')
        self.initialize_optimizer()
        if not getattr(self.model, 'fused_step', True):
            self.step = AutogradStep(self.model, self.opt, self.l2_reg, self.batch_size, trainer_config.get('seed', 0))
            return
        self.step = engine.TrainStep(self.model, self.opt, self.l2_reg, aux_reg=aux_reg, batch_size=self.batch_size,
                                     seed=trainer_config.get('seed', 0),
                                     use_graph=trainer_config.get('cuda_graph', True))

    def _host_batches(self):
        loaders = [self.dataloader] + ([self.aux_dataloader] if isinstance(self, IGCNTrainer) else [])
        for batches in zip(*loaders):
            yield [b[:, 0, :].to(device=self.device, dtype=torch.int64) for b in batches]

    def _run_epoch(self):
        self.step.reset_meter()
        if self.sampler == 'reference':
            for batches in self._host_batches():
                self.step.run(*batches)
        else:
            total, B = len(self.dataset), self.batch_size
            for lo in range(0, total, B):
                self.step.run(batch=min(B, total - lo))
        self.step.sync_params()             # column-sharded training: full-width parameters on every rank again
        return self.step.meter_avg()


class BPRTrainer(BasicTrainer, _FusedBPRMixin):
    """trainer.py:222-248."""

    def __init__(self, trainer_config):
        super().__init__(trainer_config)
        self.dataloader = DataLoader(self.dataset, batch_size=trainer_config['batch_size'],
                                     num_workers=trainer_config['dataloader_num_workers'])
        self._init_fused(trainer_config)

    def train_one_epoch(self):
        return self._run_epoch()


class IGCNTrainer(BasicTrainer, _FusedBPRMixin):
    """trainer.py:281-320."""

    def __init__(self, trainer_config):
        super().__init__(trainer_config)
        self.dataloader = DataLoader(self.dataset, batch_size=trainer_config['batch_size'],
                                     num_workers=trainer_config['dataloader_num_workers'])
        self._aux_dataloader = None
        self.aux_reg = trainer_config['aux_reg']
        self._init_fused(trainer_config, aux_reg=self.aux_reg)

    @property
    def aux_dataloader(self):
        """Host sampler in template-id space (trainer.py:287-289); built on first use because only the
        'reference' sampler mode needs it (the device sampler reads model.aux_csr())."""
        if self._aux_dataloader is None:
            aux = AuxiliaryDataset(self.dataset, self.model.user_map, self.model.item_map)
            self._aux_dataloader = DataLoader(aux, batch_size=self.config['batch_size'],
                                              num_workers=self.config['dataloader_num_workers'])
        return self._aux_dataloader

    def train_one_epoch(self):
        loss = self._run_epoch()
        self.model.feat_mat_anneal()
        return loss
