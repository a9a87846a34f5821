"""Drop-in data layer (reference dataset.py:10-14, 47-65, 116-164, 258-273).

`ProcessedDataset` reads the reference's `train.txt` / `val.txt` / `test.txt` and exposes the same
attributes (`n_users`, `n_items`, `train_data`, `val_data`, `test_data`, `train_array`, `name`,
`device`, `__len__`, `__getitem__`).  It additionally keeps `train_pairs` (an [E, 2] int64 numpy
array) so graph construction does not have to walk Python lists.  `SyntheticDataset` builds the
same object straight from igcn_cf_b200.synth (no text files): the real datasets are not available
offline.  The raw-dump preprocessors (dataset.py:167-255) are out of scope.
"""
import os
import random
import sys

import numpy as np
from torch.utils.data import Dataset

from . import synth


def get_dataset(config):
    """Name-dispatched factory (dataset.py:10-14)."""
    config = config.copy()
    cls = getattr(sys.modules[__name__], config['name'])
    return cls(config)


def output_data(file_path, data):
    """One line per user: "<user> <item> ..." (dataset.py:40-44)."""
    with open(file_path, 'w') as f:
        for user, items in enumerate(data):
            f.write(' '.join([str(user)] + [str(i) for i in items]) + '\n')


class BasicDataset(Dataset):
    """dataset.py:47-65, 116-137."""
    _train_array = None          # class-level defaults: subclasses that skip __init__ still answer train_array
    train_pairs = None

    def __init__(self, dataset_config):
        print(dataset_config)
        self.config = dataset_config
        self.name = dataset_config['name']
        self.min_interactions = dataset_config.get('min_inter')
        self.split_ratio = dataset_config.get('split_ratio')
        self.device = dataset_config['device']
        self.negative_sample_ratio = dataset_config.get('neg_ratio', 1)
        self.shuffle = dataset_config.get('shuffle', False)
        self.n_users = 0
        self.n_items = 0
        self.user_inter_lists = None
        self.train_data = None
        self.val_data = None
        self.test_data = None
        self._train_array = None
        self.train_pairs = None
        print('init dataset ' + dataset_config['name'])

    @property
    def train_array(self):
        """[[user, item], ...] (dataset.py:150-152).  Nothing in this package reads it (the kernels take
        `train_pairs`), so the E two-element Python lists are only materialised when a caller asks for them."""
        if self._train_array is None and self.train_pairs is not None:
            self._train_array = self.train_pairs.tolist()
        return self._train_array

    @train_array.setter
    def train_array(self, value):
        self._train_array = value
        if value is not None:
            self.train_pairs = None            # an array assigned from outside wins (graph.train_pairs_of re-derives)

    def __len__(self):
        return len(self.train_pairs) if self.train_pairs is not None else len(self.train_array)

    def __getitem__(self, index):
        """Host triple sampler, same draw sequence as dataset.py:119-131 (the index is ignored).
        The trainers use the device sampler (igcn_sample_triples) unless asked otherwise."""
        user = random.randint(0, self.n_users - 1)
        while not self.train_data[user]:
            user = random.randint(0, self.n_users - 1)
        pos_item = np.random.choice(self.train_data[user])
        rows = []
        for _ in range(self.negative_sample_ratio):
            neg_item = random.randint(0, self.n_items - 1)
            while neg_item in self.train_data[user]:
                neg_item = random.randint(0, self.n_items - 1)
            rows.append([user, pos_item, neg_item])
        return np.array(rows, dtype=np.int64)

    def output_dataset(self, path):
        if not os.path.exists(path):
            os.mkdir(path)
        for which in ('train', 'val', 'test'):
            output_data(os.path.join(path, which + '.txt'), getattr(self, which + '_data'))

    def csr(self, which):
        """(ptr int64 [U+1], items int64) numpy arrays of the 'train' | 'val' | 'test' lists, list order kept.
        Built once per list object (the trainers use it instead of re-walking the Python lists on every
        evaluation); replacing `self.<which>_data` by another list object invalidates the entry."""
        lists = getattr(self, which + '_data')
        cache = self.__dict__.setdefault('_csr_cache', {})
        hit = cache.get(which)
        if hit is None or hit[0] is not lists or hit[1] != len(lists):
            parsed = self.__dict__.get('_parsed', {}).get(id(lists))
            if parsed is not None and parsed[0] is lists and sum(map(len, lists)) == len(parsed[2]):
                arrays = (parsed[1], parsed[2])                    # what the file parser already produced
            else:
                from .engine import lists_to_arrays
                arrays = lists_to_arrays(lists)
            hit = (lists, len(lists), arrays)
            cache[which] = hit
        return hit[2]

    def _finish(self):
        """train_array / train_pairs from train_data (dataset.py:150-152)."""
        ptr, items = self.csr('train')
        users = np.repeat(np.arange(len(self.train_data), dtype=np.int64), np.diff(ptr))
        self._train_array = None
        self.train_pairs = np.stack([users, items], axis=1) if len(users) else np.zeros((0, 2), dtype=np.int64)


class ProcessedDataset(BasicDataset):
    """dataset.py:140-164."""

    def __init__(self, dataset_config):
        super().__init__(dataset_config)
        self.train_data = self.read_data(os.path.join(dataset_config['path'], 'train.txt'))
        self.val_data = self.read_data(os.path.join(dataset_config['path'], 'val.txt'))
        self.test_data = self.read_data(os.path.join(dataset_config['path'], 'test.txt'))
        assert len(self.train_data) == len(self.val_data)
        assert len(self.train_data) == len(self.test_data)
        self.n_users = len(self.train_data)
        self._finish()

    def read_data(self, file_path):
        """One line per user, "<user> <item> <item> ..." (dataset.py:154-164).  The whole file is tokenised by
        numpy in one call and cut into rows with the per-line token counts; the CSR arrays
        are kept for `csr()` so the trainers never walk the lists again.  Files that are not plain
        single-space separated integers go through the reference's per-token Python loop."""
        with open(file_path, 'rb') as f:
            raw = f.read().strip()
        lines = raw.split(b'\n') if raw else []
        try:
            counts = np.fromiter((ln.count(b' ') for ln in lines), dtype=np.int64, count=len(lines))
            tokens = np.fromstring(raw.decode('ascii'), dtype=np.int64, sep=' ') if raw else np.zeros(0, dtype=np.int64)
            if tokens.size != int(counts.sum()) + len(lines):
                raise ValueError('irregular separators')
        except (ValueError, UnicodeDecodeError, DeprecationWarning):
            data = []
            for line in raw.decode().split('\n') if raw else []:
                items = [int(tok) for tok in line.split(' ')[1:]]
                if items:
                    self.n_items = max(self.n_items, max(items) + 1)
                data.append(items)
            return data
        ptr = np.zeros(len(lines) + 1, dtype=np.int64)
        np.cumsum(counts, out=ptr[1:])
        keep = np.ones(tokens.size, dtype=bool)
        keep[ptr[:-1] + np.arange(len(lines))] = False            # the leading user id of every line
        items = tokens[keep]
        if items.size:
            self.n_items = max(self.n_items, int(items.max()) + 1)
        flat = items.tolist()
        bounds = ptr.tolist()
        data = [flat[bounds[u]:bounds[u + 1]] for u in range(len(lines))]
        self.__dict__.setdefault('_parsed', {})[id(data)] = (data, ptr, items)
        return data


class SyntheticDataset(BasicDataset):
    """Reference-shaped dataset generated in memory.  config: {'name': 'SyntheticDataset',
    'shape': 'gowalla'|'yelp'|'amazon'|'tiny'|'small' or (U, I, E), 'seed': 2021, 'device': ...,
    optional 'variant': 'dropui'|'dropit' (run/dropui/dataset_dropui.py, run/dropit/dataset_dropit.py)}."""

    def __init__(self, dataset_config):
        super().__init__(dataset_config)
        shape = dataset_config.get('shape', 'tiny')
        u, i, e = synth.SHAPES[shape] if isinstance(shape, str) else shape
        split = dataset_config.get('split')
        if split is None:
            split = synth.gen_synth(u, i, e, seed=dataset_config.get('seed', 2021))
        variant = dataset_config.get('variant')
        if variant == 'dropui':
            split = synth.dropui(split, dataset_config.get('ratio', 0.8))
        elif variant == 'dropit':
            split = synth.dropit(split, dataset_config.get('ratio', 0.8))
        self.split = split
        self.n_users, self.n_items = split.n_users, split.n_items
        self.train_data = split.lists('train')
        self.val_data = split.lists('val')
        self.test_data = split.lists('test')
        self._finish()


class DeviceSyntheticDataset(BasicDataset):
    """Scale-out synthetic graph generated on the GPU and kept there as a CSR (`device_graph`); there are no
    Python per-user lists, so only propagation and unmasked full-ranking run on it (BASELINE.json config 5).
    config: {'name': 'DeviceSyntheticDataset', 'shape': (U, I, E) or a synth.SHAPES name, 'seed', 'device'}."""

    def __init__(self, dataset_config):
        super().__init__(dataset_config)
        shape = dataset_config['shape']
        u, i, e = synth.SHAPES[shape] if isinstance(shape, str) else shape
        self.device_graph = synth.gen_device(u, i, e, dataset_config['device'], seed=dataset_config.get('seed', 2021))
        self.n_users, self.n_items = self.device_graph.n_users, self.device_graph.n_items

    def __len__(self):
        return self.device_graph.n_interactions

    def __getitem__(self, index):
        raise RuntimeError('DeviceSyntheticDataset has no host-side sampler; use the device sampler')


class ListDataset(BasicDataset):
    """Dataset from in-memory per-user lists.  config: {'name': 'ListDataset', 'train': [[...], ...],
    'val': [...], 'test': [...], 'n_items': int (optional), 'device': ...}."""

    def __init__(self, dataset_config):
        cfg = dict(dataset_config)
        lists = {k: cfg.pop(k) for k in ('train', 'val', 'test')}
        super().__init__(cfg)
        self.train_data = [list(map(int, x)) for x in lists['train']]
        self.val_data = [list(map(int, x)) for x in lists['val']]
        self.test_data = [list(map(int, x)) for x in lists['test']]
        self.n_users = len(self.train_data)
        seen = [max(x) + 1 for d in (self.train_data, self.val_data, self.test_data) for x in d if x]
        self.n_items = int(cfg.get('n_items') or (max(seen) if seen else 0))
        self._finish()


class AuxiliaryDataset(BasicDataset):
    """Triple sampler in template-id space for the self-enhanced loss (dataset.py:258-273)."""

    def __init__(self, dataset, user_map, item_map):
        self.n_users = len(user_map)
        self.n_items = len(item_map)
        self.device = dataset.device
        self.negative_sample_ratio = 1
        self.train_data = [[] for _ in range(self.n_users)]
        self.length = len(dataset)
        for o_user, items in enumerate(dataset.train_data):
            if o_user in user_map:
                row = self.train_data[user_map[o_user]]
                row.extend(item_map[o_item] for o_item in items if o_item in item_map)

    def __len__(self):
        return self.length
