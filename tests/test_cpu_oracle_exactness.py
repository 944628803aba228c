"""The oracle against the MATHEMATICS of collapsed Gibbs sampling for LDA (Griffiths & Steyvers 2004), which is what
``lda._lda._sample_topics`` implements -- an anchor that does not depend on the ``lda``
package being installable (the oracle's parity with the package itself stays unpinned until
``oracle/pin_against_pypi.py`` can run).

1. The conditional the oracle samples from, recovered EXACTLY from its own sweep function (the topic it picks is a
   step function of the uniform draw; the steps are found by bisection), equals the closed form
   p(z_i = k | z_-i, w) ~ (n_kw + eta) / (n_k + W eta) * (n_dk + alpha) to 1e-12.
2. Long chains of the oracle visit the K^N assignments of a tiny corpus with the frequencies of the exact posterior
   (all assignments enumerated); deliberately wrong models of the same corpus are far outside the tolerance.
3. The log-likelihood the oracle reports is the log of that joint, assignment by assignment.
"""
import itertools
import math

import numpy as np
import pytest

from oracle import lda_oracle as lo

_i32p, _f64p = lo._i32p, lo._f64p


def _counts(WS, DS, ZS, D, W, K):
    nzw, ndz, nz = np.zeros((K, W), np.int32), np.zeros((D, K), np.int32), np.zeros(K, np.int32)
    np.add.at(nzw, (ZS, WS), 1)
    np.add.at(ndz, (DS, ZS), 1)
    np.add.at(nz, ZS, 1)
    return nzw, ndz, nz


def _sweep(WS, DS, ZS, nzw, ndz, nz, alpha, eta, rands):
    K, W = nzw.shape
    a, e = np.full(K, alpha, np.float64), np.full(W, eta, np.float64)
    rc = lo.lib().oracle_sample_topics(lo._p(WS, _i32p), lo._p(DS, _i32p), lo._p(ZS, _i32p), len(WS), K, W,
                                       lo._p(nzw, _i32p), lo._p(ndz, _i32p), lo._p(nz, _i32p),
                                       lo._p(a, _f64p), lo._p(e, _f64p), lo._p(rands, _f64p), len(rands))
    assert rc == 0


def _binary_corpus(rng, D, W, density):
    while True:
        m = rng.random((D, W)) < density
        if m.any(axis=1).all() and m.any(axis=0).all():
            DS, WS = np.nonzero(m)
            return WS.astype(np.int32), DS.astype(np.int32)


@pytest.mark.parametrize("seed,K,alpha,eta", [(1, 2, 0.5, 0.1), (2, 4, 12.5, 0.1), (3, 7, 50 / 7, 0.01), (4, 5, 0.05, 2.0)])
def test_oracle_conditional_is_the_closed_form(seed, K, alpha, eta):
    rng = np.random.default_rng(seed)
    D, W = 5, 9
    WS0, DS0 = _binary_corpus(rng, D, W, 0.45)
    N = len(WS0)
    Z0 = rng.integers(0, K, N).astype(np.int32)
    for token in rng.choice(N, 6, replace=False):
        order = np.r_[token, np.delete(np.arange(N), token)]  # the token under test is sampled first
        WS, DS, Z = WS0[order].copy(), DS0[order].copy(), Z0[order].copy()
        nzw, ndz, nz = _counts(WS, DS, Z, D, W, K)

        def pick(u):
            z, a, b, c = Z.copy(), nzw.copy(), ndz.copy(), nz.copy()
            rands = np.full(N, 0.5)
            rands[0] = u
            _sweep(WS, DS, z, a, b, c, alpha, eta, rands)
            return int(z[0])

        # boundaries of the step function u -> topic: the largest u that still gives a topic <= k
        bounds = []
        for k in range(K - 1):
            lo_u, hi_u = 0.0, 1.0
            if pick(1.0) <= k:
                bounds.append(1.0)
                continue
            for _ in range(80):
                mid = 0.5 * (lo_u + hi_u)
                if pick(mid) <= k:
                    lo_u = mid
                else:
                    hi_u = mid
            bounds.append(lo_u)
        got = np.diff(np.r_[0.0, bounds, 1.0])
        w, d, z = WS[0], DS[0], Z[0]
        a, b, c = nzw.astype(np.float64), ndz.astype(np.float64), nz.astype(np.float64)
        a[z, w] -= 1
        b[d, z] -= 1
        c[z] -= 1
        want = (a[:, w] + eta) / (c + W * eta) * (b[d] + alpha)
        want /= want.sum()
        np.testing.assert_allclose(got, want, atol=1e-12, rtol=0)
        assert pick(0.0) == 0 or want[0] == 0  # left bisection: u = 0 gives the first topic


def _log_joint(WS, DS, ZS, D, W, K, alpha, eta):
    """log p(w, z | alpha, eta) of Griffiths & Steyvers (their eqs. 2 and 3)."""
    nzw, ndz, nz = _counts(WS, DS, np.asarray(ZS, np.int32), D, W, K)
    nd = ndz.sum(axis=1)
    lg = math.lgamma
    ll = 0.0
    for k in range(K):
        ll += lg(W * eta) - W * lg(eta) + sum(lg(nzw[k, w] + eta) for w in range(W)) - lg(nz[k] + W * eta)
    for d in range(D):
        ll += lg(K * alpha) - K * lg(alpha) + sum(lg(ndz[d, k] + alpha) for k in range(K)) - lg(nd[d] + K * alpha)
    return ll


def _exact_posterior(WS, DS, D, W, K, alpha, eta):
    states = list(itertools.product(range(K), repeat=len(WS)))
    logp = np.array([_log_joint(WS, DS, s, D, W, K, alpha, eta) for s in states])
    p = np.exp(logp - logp.max())
    return p / p.sum(), logp


CORPORA = {
    # (cell, region) pairs of a binary matrix, K, alpha, eta
    "K2": ([(0, 0), (0, 1), (0, 2), (1, 1), (1, 2), (1, 3), (2, 0), (2, 3)], 2, 0.7, 0.4),
    "K3": ([(0, 0), (0, 1), (1, 1), (1, 2), (2, 0), (2, 2)], 3, 0.3, 0.2),
}


@pytest.mark.parametrize("name", sorted(CORPORA))
def test_oracle_chain_visits_assignments_with_their_posterior_probability(name):
    pairs, K, alpha, eta = CORPORA[name]
    DS = np.array([p[0] for p in pairs], np.int32)
    WS = np.array([p[1] for p in pairs], np.int32)
    D, W, N = int(DS.max()) + 1, int(WS.max()) + 1, len(pairs)
    exact, logp = _exact_posterior(WS, DS, D, W, K, alpha, eta)
    rng = np.random.default_rng(11)
    n_sweeps, burn = 300_000, 1000
    rands = rng.random((n_sweeps + burn, N))
    Z = (np.arange(N) % K).astype(np.int32)  # the package's initialisation
    nzw, ndz, nz = _counts(WS, DS, Z, D, W, K)
    weights = K ** np.arange(N - 1, -1, -1)
    visits = np.zeros(K ** N, np.int64)
    for s in range(n_sweeps + burn):
        _sweep(WS, DS, Z, nzw, ndz, nz, alpha, eta, rands[s])
        if s >= burn:
            visits[int(Z @ weights)] += 1
    # the counts the oracle maintains are those of its assignment
    a, b, c = _counts(WS, DS, Z, D, W, K)
    assert np.array_equal(a, nzw) and np.array_equal(b, ndz) and np.array_equal(c, nz)
    freq = visits / visits.sum()
    tv = 0.5 * np.abs(freq - exact).sum()
    # wrong models of the same corpus, for scale: the exact posteriors under other hyper-parameters
    tv_alpha = 0.5 * np.abs(_exact_posterior(WS, DS, D, W, K, alpha * 1.5, eta)[0] - exact).sum()
    tv_eta = 0.5 * np.abs(_exact_posterior(WS, DS, D, W, K, alpha, eta * 1.5)[0] - exact).sum()
    # what n independent draws from the exact posterior would leave (mean total variation of a multinomial sample);
    # consecutive sweeps are correlated, hence the factor
    tv_sampling = 0.5 * math.sqrt(2 / (math.pi * n_sweeps)) * np.sqrt(exact * (1 - exact)).sum()
    print(f"{name}: TV {tv:.4f}, sampling noise {tv_sampling:.4f}, wrong alpha {tv_alpha:.3f}, wrong eta {tv_eta:.3f}")
    assert tv < 2.0 * tv_sampling, (tv, tv_sampling)
    assert tv_alpha > 4 * tv and tv_eta > 4 * tv, (tv, tv_alpha, tv_eta)
    # every assignment with non-negligible mass is visited in proportion (relative error of the 20 most likely ones)
    top = np.argsort(exact)[::-1][:20]
    np.testing.assert_allclose(freq[top], exact[top], rtol=0.06)
    # the oracle's log-likelihood (lda._lda._loglikelihood, A7) is the log of the same joint
    for s in (0, int(top[0]), int(top[5]), K ** N - 1):
        z = np.array([(s // int(wt)) % K for wt in weights], np.int32)
        n1, n2, n3 = _counts(WS, DS, z, D, W, K)
        assert lo.loglikelihood_true(n1, n2, n3, alpha, eta) == pytest.approx(logp[s], abs=1e-10)
