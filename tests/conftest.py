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


def check_topk_lists(rec, ref, rep_oracle, n_users, rep_mine=None, scale_tol=1e-5):
    """Top-k lists of the CUDA path (`rec` [U, k]) against the oracle's (`ref`): identical wherever the ORACLE's own
    scores are not tied.  For every position where the two lists name different items a and b, the oracle's scores
    s(u, a) and s(u, b) -- recomputed here in fp64 from the oracle's representation -- must agree to within
        4 ulp(fp32) of the score  +  2 x (the largest difference between the two paths' scores of this user's listed items)
    i.e. be a tie at the precision to which the two representations agree (they are only required to match to
    `scale_tol` relative, north_star); that second term is itself asserted to stay below scale_tol x score scale.
    Returns the number of users whose lists differ (all of them proven ties)."""
    rec, ref = np.asarray(rec, dtype=np.int64), np.asarray(ref, dtype=np.int64)
    assert rec.shape == ref.shape
    rows = np.nonzero((rec != ref).any(axis=1))[0]
    if len(rows) == 0:
        return 0
    R = np.asarray(rep_oracle, dtype=np.float64)
    M = None if rep_mine is None else np.asarray(rep_mine, dtype=np.float64)
    scale = float(np.abs(R[:n_users] @ R[n_users:n_users + 64].T).max()) if len(R) > n_users else 1.0
    eps = float(np.finfo(np.float32).eps)
    for u in rows:
        a, b = rec[u], ref[u]
        pos = np.nonzero(a != b)[0]
        ok = (a >= 0) & (b >= 0)
        assert ok[pos].all(), ('list lengths differ', int(u))
        sa, sb = R[n_users + a] @ R[u], R[n_users + b] @ R[u]
        delta = 0.0
        if M is not None:
            delta = float(max(np.abs(M[n_users + a] @ M[u] - sa).max(), np.abs(M[n_users + b] @ M[u] - sb).max()))
            assert delta <= scale_tol * max(scale, np.abs(sa).max()), ('scores drifted', int(u), delta)
        tol = 4 * eps * np.maximum(np.abs(sa), np.abs(sb)) + 2 * delta
        bad = np.abs(sa - sb)[pos] > tol[pos]
        assert not bad.any(), ('lists differ where the oracle has no tie', int(u), a[pos][bad], b[pos][bad],
                               sa[pos][bad], sb[pos][bad])
    return len(rows)


def check_metrics(mine, ref, n_diff_rows, n_valid_users, keys=None):
    """Metrics identical when no list differs; otherwise each differing (tied) user can move a per-user ratio by at
    most 1, so a mean over n_valid_users moves by at most n_diff_rows / n_valid_users."""
    bound = 0.0 if n_diff_rows == 0 else n_diff_rows / max(1, n_valid_users)
    for name, k, want in ref:
        got = float(mine[name][k])
        if bound == 0.0:
            assert got == float(want), (name, k, got, float(want))
        else:
            assert abs(got - float(want)) <= bound, (name, k, got, float(want), bound)
