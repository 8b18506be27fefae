"""Host-side check of the K2b filter's error bound (csrc/gemm_filter.cu, "error bound"; runs without a GPU).

The kernel prunes a row when `s~ + eps < L`; that is only exact if |s~ - s| <= eps for EVERY row and query.
eps is derived from the actual bf16 rounding residuals: |x.q - x_b.q_b| <= |r||q_b| + |x||t| with r = x - x_b,
t = q - q_b.  This file emulates the operand rounding with the oracle's bf16 helper and fp64 arithmetic and
checks the inequality on adversarial (coherently rounding, few-level, same-sign) and random inputs, and pins
the counter-example on which the round-1 bound (2^-9 per operand) failed."""
import zlib

import numpy as np
import pytest

from oracle import exact_search as oracle


def _operand_bound(X, q):
    """(|s~ - s|, eps_operands) per row for raw inner products, fp64."""
    X64, q64 = X.astype(np.float64), q.astype(np.float64)
    Xb, qb = oracle.bf16_round(X).astype(np.float64), oracle.bf16_round(q).astype(np.float64)
    err = np.abs(Xb @ qb - X64 @ q64)
    eps = np.linalg.norm(X64 - Xb, axis=1) * np.linalg.norm(qb) + np.linalg.norm(X64, axis=1) * np.linalg.norm(q64 - qb)
    return err, eps


def test_round1_counter_example_needs_the_data_derived_bound():
    D = 608
    q = np.zeros(D, np.float32)
    q[:301], q[301:601] = 1.0039, 1.00391
    A = np.zeros(D, np.float32)
    A[:301] = 1.0039
    B = np.zeros(D, np.float32)
    B[301:601] = 1.00391
    X = np.stack([A, B])
    err, eps = _operand_bound(X, q)
    scale = np.linalg.norm(X.astype(np.float64), axis=1) * np.linalg.norm(q.astype(np.float64))
    old = 2.0 ** -8 * 1.002 + D * 1.2e-7 + 1e-6            # round 1: "2^-9 per rounded operand"
    assert np.all(err / scale > old), "the round-1 bound must be violated here (that was the bug)"
    assert np.all(err <= eps)
    assert np.all(eps / scale < 2.0 ** -7 * 1.001)          # never looser than the worst case 2 * 2^-8


@pytest.mark.parametrize("seed", range(8))
def test_operand_bound_holds_on_adversarial_and_random_data(seed):
    rng = np.random.default_rng(seed)
    dim = (4, 64, 200, 608, 768, 1536, 5, 96)[seed]
    n = 400
    levels = np.array([1.0039, 1.00391, 0.50195, 2.0078, 0.25098, 1.9922, 0.99609, 3.0117], np.float32)
    Xs = [
        (rng.random((n, dim)) < 0.6) * levels[rng.integers(0, 8, size=(n, 1))],     # few-level, same sign, coherent
        rng.standard_normal((n, dim)),                                              # ordinary
        np.abs(rng.standard_normal((n, dim))) * (1.0 + 2.0 ** -9),                  # same sign
        rng.standard_normal((n, dim)) * 10.0 ** rng.uniform(-20, 20, size=(n, 1)),  # wild scales
    ]
    qs = [levels[rng.integers(0, 8)] * np.ones(dim), rng.standard_normal(dim), np.abs(rng.standard_normal(dim))]
    for X in Xs:
        for q in qs:
            err, eps = _operand_bound(np.asarray(X, np.float32), np.asarray(q, np.float32))
            assert np.all(err <= eps * (1.0 + 1e-12) + 1e-300)


def test_typical_bound_is_tighter_than_worst_case():
    """On ordinary data the data-derived eps is ~3-4x below the worst case 2 * 2^-8 |x||q| (fewer candidates)."""
    rng = np.random.default_rng(0)
    X = rng.standard_normal((2000, 768)).astype(np.float32)
    q = rng.standard_normal(768).astype(np.float32)
    _, eps = _operand_bound(X, q)
    rel = eps / (np.linalg.norm(X.astype(np.float64), axis=1) * np.linalg.norm(q.astype(np.float64)))
    assert rel.max() < 0.0040 and rel.mean() < 0.0036


# ------------------------------------------------------------------------------------------ int8 operands
# Host restatement of shadow8_rows_kernel / prep_queries_i8_kernel / the int8 epilogue (csrc/gemm_filter.cu): the
# quantisation is emulated in fp32 exactly as the kernels do it, the inequality is checked in fp64.
_F = np.float32


def _fma32(a, b, c):
    return (a.astype(np.float64) * b.astype(np.float64) + c.astype(np.float64)).astype(_F)


def _sum32(v, axis=-1):
    """an fp32 sum in SOME order (the bound must not depend on the order: norms carry an inflation factor)"""
    return np.add.reduce(v.astype(_F), axis=axis, dtype=_F)


def _i8_rows(X):
    X = X.astype(_F)
    mx = np.abs(X).max(axis=1, keepdims=True)
    with np.errstate(divide="ignore", invalid="ignore"):
        sx = np.where(mx > 0, (mx / _F(127)).astype(_F), _F(0))
        rinv = np.where(mx > 0, (_F(127) / mx).astype(_F), _F(0))
    xi = np.clip(np.rint((X * rinv).astype(_F)), -127, 127).astype(_F)
    t = _fma32(-sx * np.ones_like(X), xi, X)
    rres = (np.sqrt(_sum32(t * t)).astype(_F) * _F(1.001)).astype(_F)
    return xi, sx[:, 0], rres


def _i8_query(q):
    q = q.astype(_F)
    mx = np.abs(q).max()
    ok = mx > 0 and np.isfinite(mx)
    s1 = _F(mx / _F(127)) if ok else _F(0)
    r1 = _F(_F(127) / mx) if ok else _F(0)
    s2 = _F(s1 / _F(254))
    r2 = _F(_F(254) * r1)
    d1 = np.clip(np.rint((q * r1).astype(_F)), -127, 127).astype(_F)
    e1 = _fma32(-s1 * np.ones_like(q), d1, q)
    d2 = np.clip(np.rint((e1 * r2).astype(_F)), -127, 127).astype(_F)
    if not ok:
        d1[:] = 0
        d2[:] = 0
    qt = _fma32(s2 * np.ones_like(q), d2, (s1 * d1).astype(_F))
    t = (q - qt).astype(_F)
    ss, sb, st = _sum32(q * q), _sum32(qt * qt), _sum32(t * t)
    q_bn = _F(np.sqrt(sb) * _F(1.0001))
    q_tn = _F(_F(np.sqrt(st)) * _F(1.01) + _F(2e-7) * _F(np.sqrt(ss)))
    return d1, d2, s1, s2, _F(np.sqrt(ss)), q_bn, q_tn


def _acc_rel(dim, dpad):
    return _F(dim * 1.21e-7 + (dpad / 32.0 + 40.0) * 6e-8 + 1e-6)       # filter_acc_rel


def _i8_case(kind, rng, n, dim):
    if kind == "gauss":
        X = rng.standard_normal((n, dim))
    elif kind == "outlier":      # one huge element eats the row's scale: everything else quantises to 0
        X = rng.standard_normal((n, dim))
        X[np.arange(n), rng.integers(0, dim, n)] *= 10.0 ** rng.uniform(2, 6, n)
    elif kind == "cauchy":
        X = rng.standard_cauchy((n, dim))
    elif kind == "few_level":    # values halfway between two grid points, all rounding the same way
        X = (rng.random((n, dim)) < 0.7) * (rng.integers(1, 254, size=(n, dim)) + 0.5) / 254.0
        X[:, 0] = 127.5 / 127.0
    elif kind == "scales":
        X = rng.standard_normal((n, dim)) * 10.0 ** rng.uniform(-15, 15, size=(n, 1))
    elif kind == "exact_grid":   # rows that quantise exactly: residual ~ 2^-24 |x| from the rounding of the scale alone
        X = rng.integers(-127, 128, size=(n, dim)).astype(np.float64) * rng.uniform(0.3, 3.0, size=(n, 1))
        X[:, 0] = 127.0 * np.abs(X).max(axis=1) / 127.0
    else:
        raise AssertionError(kind)
    return X.astype(_F)


@pytest.mark.parametrize("kind", ["gauss", "outlier", "cauchy", "few_level", "scales", "exact_grid"])
@pytest.mark.parametrize("dim", [5, 96, 768])
def test_int8_operand_bound_holds(kind, dim):
    rng = np.random.default_rng(zlib.crc32(f"{kind}{dim}".encode()))
    n = 300
    X = _i8_case(kind, rng, n, dim)
    dpad = (dim + 3) // 4 * 4
    queries = [rng.standard_normal(dim), np.abs(rng.standard_normal(dim)), _i8_case(kind, rng, 1, dim)[0],
               np.ones(dim) * 0.37, np.eye(1, dim, 0)[0] * 3.0]
    xi, sx, rres = _i8_rows(X)
    X64 = X.astype(np.float64)
    xn = np.linalg.norm(X64, axis=1)
    for q in queries:
        q = np.asarray(q, _F)
        d1, d2, s1, s2, qnrm, q_bn, q_tn = _i8_query(q)
        q64 = q.astype(np.float64)
        # pure mathematics: x.q = x~.q~ + r.q~ + x.t
        xt = sx.astype(np.float64)[:, None] * xi.astype(np.float64)
        qt = float(s1) * d1.astype(np.float64) + float(s2) * d2.astype(np.float64)
        exact = X64 @ q64
        op_err = np.abs(xt @ qt - exact)
        op_eps = np.linalg.norm(X64 - xt, axis=1) * np.linalg.norm(qt) + xn * np.linalg.norm(q64 - qt)
        # (the last term: fp64 cancellation in `xt @ qt - exact` itself, relative to the size of the products)
        assert np.all(op_err <= op_eps * (1 + 1e-12) + 1e-14 * xn * np.linalg.norm(q64) + 1e-300)
        # the kernel's own quantities: d = fp32(sx * fma(A2, s2, A1 * s1)) from exact integer dots, eps (ip form) =
        # |r|~ 1.001 |q~|~ + |x| 1.00001 * 1.001 (acc_rel |q| + |t|~)   (make_query_bound / bound_eval)
        A1 = xi.astype(np.float64) @ d1.astype(np.float64)
        A2 = xi.astype(np.float64) @ d2.astype(np.float64)
        assert np.abs(A1).max() < 2 ** 31 and np.abs(A2).max() < 2 ** 31
        a1, a2 = A1.astype(_F), A2.astype(_F)
        d = (sx * _fma32(a2, np.full(n, s2, _F), (a1 * s1).astype(_F))).astype(_F)
        A = _F(1.001) * q_bn
        C = _F(1.001) * _F(_acc_rel(dim, dpad) * qnrm + q_tn)
        eps = rres.astype(np.float64) * float(A) + xn * 1.00001 * float(C)
        err = np.abs(d.astype(np.float64) - exact)
        k1_slack = (dpad / 32.0 + 40.0) * 2.0 ** -24 * xn * np.linalg.norm(q64)    # |s_K1 - s_exact|, part of acc_rel
        assert np.all(err + k1_slack <= eps), float(np.max((err + k1_slack) / np.maximum(eps, 1e-300)))


def test_int8_typical_bound_is_a_fraction_of_sigma():
    """Gaussian rows at D = 768: eps ~ 1 % of |x||q| = a few tenths of the score distribution's sigma (1/sqrt(D))."""
    rng = np.random.default_rng(3)
    X = rng.standard_normal((2000, 768)).astype(_F)
    q = rng.standard_normal(768).astype(_F)
    _, _, rres = _i8_rows(X)
    d1, d2, s1, s2, qnrm, q_bn, q_tn = _i8_query(q)
    xn = np.linalg.norm(X.astype(np.float64), axis=1)
    rel = (rres * float(q_bn) + xn * float(q_tn)) / (xn * float(qnrm))
    assert rel.mean() < 0.012 and rel.mean() * np.sqrt(768) < 0.35
