"""Soundness of the tensor-core scoring bound, checked on the CPU in exact arithmetic.

igcn_tc_pack (csrc/eval_tc.cu: tc_pack_kernel, tc_scale) rounds the scaled operands to fp16 and appends one K
column holding b_u = fp16_up(c * |u| * 1.000001 + 2^-21) for users and b_i = fp16_up(|i| * 1.000001 + 2^-11) for
(mean-centred) items, so that the tensor core's  s_hat = <h_u, h_i> + b_u * b_i  is an UPPER bound of the exact
score <v_u, v_i> of the fp32 vectors it was packed from.  Everything the candidate filter and the proof in
igcn_tc_finalize do rests on that inequality.  This test restates the packing in numpy (same fp32 operations) and
verifies the inequality with rational arithmetic -- including the worst-case fp32 accumulation error of the MMA
(80 products, one rounding each) -- on random, tiny-norm, huge-dynamic-range and fp16-subnormal inputs."""
from fractions import Fraction

import numpy as np
import pytest

TC_C = np.float32(1.0e-3)
EPS_U = np.float32(2.0 ** -21)
EPS_I = np.float32(2.0 ** -11)
F32 = np.float32


def tc_scale(maxabs):
    """tc_common.cuh tc_scale: 2^(10 - e) with maxabs = f * 2^e, f in [0.5, 1)."""
    if not (maxabs > 0) or not np.isfinite(maxabs):
        return F32(1.0)
    _, e = np.frexp(F32(maxabs))
    return F32(np.ldexp(1.0, 10 - int(e)))


def half_up(x):
    h = np.float16(x)
    if float(h) < float(x):
        h = np.nextafter(h, np.float16(np.inf))
    return h


def row_norm(v):
    """sqrtf(sum of squares) * 1.000001f with the kernel's reduction order: 4 elements per lane, xor tree."""
    v = v.astype(F32)
    lanes = np.zeros(16, dtype=F32)
    for l in range(len(v) // 4):
        x, y, z, w = v[4 * l:4 * l + 4]
        lanes[l] = F32(F32(F32(x * x) + F32(y * y)) + F32(z * z)) + F32(w * w)
    for o in (8, 4, 2, 1):
        lanes = (lanes + lanes[np.arange(16) ^ o]).astype(F32)
    return F32(np.sqrt(lanes[0], dtype=F32) * F32(1.000001))


def pack(rep, n_users):
    rep = rep.astype(F32)
    scale = tc_scale(np.abs(rep).max())
    items = rep[n_users:]
    # igcn_colsum_masked sums in a fixed order; any fp32 mean works for the inequality as long as the SAME
    # mean is used for the packed item and for the exact score it is compared with
    mean = (items.sum(axis=0, dtype=F32) * F32(1.0 / len(items))).astype(F32)
    out = []
    for r in range(len(rep)):
        is_user = r < n_users
        v = rep[r] if is_user else (rep[r] - mean).astype(F32)
        v = (v * scale).astype(F32)
        norm = row_norm(v)
        b = half_up(F32(TC_C * norm + EPS_U)) if is_user else half_up(F32(norm + EPS_I))
        out.append((v, v.astype(np.float16), b))
    return out


def frac_dot(a, b):
    return sum(Fraction(float(x)) * Fraction(float(y)) for x, y in zip(a, b))


def check(rep, n_users, pairs):
    rows = pack(rep, n_users)
    worst = None
    for u, i in pairs:
        vu, hu, bu = rows[u]
        vi, hi, bi = rows[n_users + i]
        exact = frac_dot(vu, vi)
        s_hat = frac_dot(hu, hi) + Fraction(float(bu)) * Fraction(float(bi))
        # fp32 accumulation in TMEM: at most one rounding per product, each relative 2^-24 of a partial sum
        # that never exceeds the sum of magnitudes
        mag = sum(abs(Fraction(float(x)) * Fraction(float(y))) for x, y in zip(hu, hi)) + Fraction(float(bu)) * Fraction(float(bi))
        acc_err = 80 * Fraction(1, 2 ** 24) * mag
        margin = s_hat - acc_err - exact
        assert margin >= 0, (u, i, float(exact), float(s_hat), float(acc_err))
        rel = float(margin / mag) if mag else 0.0
        worst = rel if worst is None else min(worst, rel)
    return worst


def _pairs(rng, n_users, n_items, n):
    return list(zip(rng.integers(n_users, size=n).tolist(), rng.integers(n_items, size=n).tolist()))


@pytest.mark.parametrize('D', [64, 32, 4])
def test_upper_bound_random_embeddings(D):
    rng = np.random.default_rng(D)
    rep = (rng.standard_normal((40 + 60, D)) * 0.1).astype(F32)
    check(rep, 40, _pairs(rng, 40, 60, 400))


def test_upper_bound_aligned_and_opposed_vectors():
    """Cauchy-Schwarz is tight when the rounding errors line up with the other operand: items that are
    multiples of the user vector, in both directions."""
    rng = np.random.default_rng(1)
    users = (rng.standard_normal((20, 64)) * 0.1).astype(F32)
    items = np.concatenate([users * F32(1.7), -users * F32(0.9), np.sign(users) * F32(0.05)]).astype(F32)
    rep = np.concatenate([users, items])
    check(rep, 20, [(u, i) for u in range(20) for i in (u, 20 + u, 40 + u, (u + 7) % 60)])


def test_upper_bound_dynamic_range_and_subnormals():
    """One huge element fixes the scale; the other rows then land in fp16's subnormal range (absolute rounding
    error 2^-25 per element) or flush to zero -- the 2^-21 / 2^-11 entries pay for that."""
    rng = np.random.default_rng(2)
    rep = (rng.standard_normal((30 + 50, 64)) * 0.1).astype(F32)
    rep[3] *= F32(1e-4)                    # tiny user
    rep[4] *= F32(1e-7)                    # user entirely below fp16's subnormal step after scaling
    rep[5] = 0                             # zero user
    rep[30 + 2] *= F32(40.0)               # huge item: sets maxabs
    rep[30 + 3] *= F32(1e-5)               # tiny item (relative to the mean: still the mean's size)
    rep[30 + 4] = rep[30:].mean(axis=0)    # item equal to the mean: centred row ~ 0
    pairs = [(u, i) for u in (0, 3, 4, 5, 9) for i in (0, 2, 3, 4, 11)] + _pairs(rng, 30, 50, 200)
    check(rep, 30, pairs)
    # every row tiny except one element: all dims of most rows are subnormal in fp16
    rep2 = (rng.standard_normal((10 + 10, 64)) * 1e-6).astype(F32)
    rep2[0, 0] = F32(3.0)
    check(rep2, 10, [(u, i) for u in range(10) for i in range(10)])


def test_bound_is_not_wastefully_loose():
    """The slack stays near c |u| |i| (c = 1e-3): the candidate lists depend on it being tight."""
    rng = np.random.default_rng(3)
    rep = (rng.standard_normal((30 + 30, 64)) * 0.1).astype(F32)
    rows = pack(rep, 30)
    for u, i in _pairs(rng, 30, 30, 100):
        vu, hu, bu = rows[u]
        vi, hi, bi = rows[30 + i]
        slack = float(frac_dot(hu, hi) + Fraction(float(bu)) * Fraction(float(bi)) - frac_dot(vu, vi))
        nu, ni = float(np.linalg.norm(vu.astype(np.float64))), float(np.linalg.norm(vi.astype(np.float64)))
        assert 0.0 <= slack <= 2.2e-3 * nu * ni + 1.0
