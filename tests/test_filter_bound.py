"""Host-side check of the K2b filter's error bound (csrc/gemm_filter.cu, "error bound"; runs without a GPU).

The kernel prunes a row when `s~ + eps < L`; that is only exact if |s~ - s| <= eps for EVERY row and query.
eps is derived from the actual bf16 rounding residuals: |x.q - x_b.q_b| <= |r||q_b| + |x||t| with r = x - x_b,
t = q - q_b.  This file emulates the operand rounding with the oracle's bf16 helper and fp64 arithmetic and
checks the inequality on adversarial (coherently rounding, few-level, same-sign) and random inputs, and pins
the counter-example on which the round-1 bound (2^-9 per operand) failed."""
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
