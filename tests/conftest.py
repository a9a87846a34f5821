import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, 'tests', 'golden')


def pytest_configure(config):
    config.addinivalue_line('markers', 'gpu: needs a CUDA device (run on the B200 box)')


def load_golden(name):
    return dict(np.load(os.path.join(GOLDEN, name + '.npz')))


def lists_of(ptr, items):
    return [items[ptr[u]:ptr[u + 1]].tolist() for u in range(len(ptr) - 1)]


@pytest.fixture(scope='session')
def tiny():
    """The tiny synthetic split as python lists + the train pair array (reference order)."""
    d = load_golden('tiny_data')
    out = {'n_users': int(d['n_users']), 'n_items': int(d['n_items'])}
    for which in ('train', 'val', 'test'):
        out[which] = lists_of(d[which + '_ptr'], d[which + '_items'])
    out['pairs'] = np.array([[u, i] for u in range(out['n_users']) for i in out['train'][u]], dtype=np.int64)
    return out


def rel_err(a, b):
    """max |a-b| relative to the scale of b (the 1e-5 'relative' bar of north_star)."""
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    return float(np.max(np.abs(a - b)) / max(np.max(np.abs(b)), 1e-30))
