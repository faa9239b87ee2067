"""GPU parity tests of the host-buffer C-ABI entry points (mfk_kmf_sgd_host / mfk_bias_sgd_host / mfk_bias_als_host):
the calls a binding without torch makes and the one bench.py's `e2e` number goes through.  Plain
numpy host arrays in, results compared with the fp64 oracle replaying the order the call reports (`h_order`)."""
import ctypes as C

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _rel(a, b):
    return float(np.max(np.abs(a - b)) / max(1e-12, np.max(np.abs(b))))


def _data(seed, U, I, N, hot=0.0):
    rng = np.random.default_rng(seed)
    keys = rng.choice(U * I, N, replace=False)
    u, i = (keys // I).astype(np.int32), (keys % I).astype(np.int32)
    if hot:
        m = rng.random(N) < hot
        i[m] = rng.integers(0, 3, m.sum())
        keep = np.unique(u.astype(np.int64) * I + i, return_index=True)[1]
        rng.shuffle(keep)
        u, i = np.ascontiguousarray(u[keep]), np.ascontiguousarray(i[keep])
    r = rng.integers(1, 6, len(u)).astype(np.float32)
    return u, i, r, rng


@pytest.mark.parametrize("kname,kid", [("linear", 0), ("sigmoid", 1), ("rbf", 2)])
@pytest.mark.parametrize("flags", [(1, 1), (1, 0)])
def test_kmf_sgd_host_matches_oracle_replay(kname, kid, flags):
    from matrix_factorization_b200 import _lib
    from oracle import oracle as orc

    L = _lib.lib()
    U, I, N, F, E = 700, 300, 30_000, 96, 3
    u, i, r, rng = _data(kid + 10 * flags[1], U, I, N, hot=0.4 if kname == "linear" else 0.0)
    n = len(u)
    ld = F
    P = rng.normal(0, 0.1, (U, ld)).astype(np.float32)
    Q = rng.normal(0, 0.1, (I, ld)).astype(np.float32)
    bu = (rng.normal(0, 0.05, U) * (kname != "rbf")).astype(np.float32)
    bi = (rng.normal(0, 0.05, I) * (kname != "rbf")).astype(np.float32)
    P0, Q0, bu0, bi0 = (x.astype(np.float64) for x in (P, Q, bu, bi))
    mu, lr, reg, gamma = float(r.mean()), {"linear": 0.01, "sigmoid": 0.05, "rbf": 0.3}[kname], 0.02, 0.05
    rmse = np.zeros(E, dtype=np.float64)
    order = np.zeros(n, dtype=np.int64)
    opts = _lib.PlanOpts(0, 0, F, 150 if kname == "linear" else 0xFFFFFFFF, 0, 0, 0)
    _lib.check(L.mfk_kmf_sgd_host(kid, _lib.ptr(u), _lib.ptr(i), _lib.ptr(r), n, U, I, _lib.ptr(P), _lib.ptr(Q), _lib.ptr(bu),
                                  _lib.ptr(bi), F, ld, mu, E, lr, reg, gamma, 0.0, 5.0, flags[0], flags[1], C.byref(opts),
                                  _lib.ptr(rmse), _lib.ptr(order)))
    assert np.array_equal(np.sort(order), np.arange(n))
    Po, Qo, buo, bio = P0, Q0, bu0, bi0
    r64 = r.astype(np.float64)
    for e in range(E):
        Po, Qo, buo, bio = orc.kmf_replay(kname, u, i, r64, order, mu, buo, bio, Po, Qo, lr, reg, gamma, 0.0, 5.0,
                                          bool(flags[0]), bool(flags[1]))
        assert abs(rmse[e] - orc.kmf_rmse(kname, u, i, r64, mu, buo, bio, Po, Qo, gamma)) < 5e-5
    assert _rel(P.astype(np.float64), Po) < 1e-4 and _rel(Q.astype(np.float64), Qo) < 1e-4
    assert np.max(np.abs(bu - buo)) < 5e-5 and np.max(np.abs(bi - bio)) < 5e-5
    if not flags[1]:
        assert np.array_equal(Q.astype(np.float64), Q0) and np.array_equal(bi.astype(np.float64), bi0)


def test_kmf_sgd_host_argument_errors():
    from matrix_factorization_b200 import _lib

    L = _lib.lib()
    u = np.zeros(1, np.int32)
    r = np.ones(1, np.float32)
    P = np.zeros((1, 4), np.float32)
    b = np.zeros(1, np.float32)
    rm = np.zeros(1)
    with pytest.raises(_lib.MfkError):  # bad kernel id
        _lib.check(L.mfk_kmf_sgd_host(7, _lib.ptr(u), _lib.ptr(u), _lib.ptr(r), 1, 1, 1, _lib.ptr(P), _lib.ptr(P), _lib.ptr(b),
                                      _lib.ptr(b), 4, 4, 3.0, 1, 0.1, 0.1, 0.1, 0.0, 5.0, 1, 1, None, _lib.ptr(rm), None))
    with pytest.raises(_lib.MfkError):  # ld not a multiple of 4
        _lib.check(L.mfk_kmf_sgd_host(0, _lib.ptr(u), _lib.ptr(u), _lib.ptr(r), 1, 1, 1, _lib.ptr(P), _lib.ptr(P), _lib.ptr(b),
                                      _lib.ptr(b), 3, 3, 3.0, 1, 0.1, 0.1, 0.1, 0.0, 5.0, 1, 1, None, _lib.ptr(rm), None))
    # empty rating set: parameters come back unchanged, RMSE is NaN like the reference's mean of an empty array
    P[:] = 0.25
    _lib.check(L.mfk_kmf_sgd_host(0, None, None, None, 0, 1, 1, _lib.ptr(P), _lib.ptr(P.copy()), _lib.ptr(b), _lib.ptr(b), 4, 4,
                                  3.0, 1, 0.1, 0.1, 0.1, 0.0, 5.0, 1, 1, None, _lib.ptr(rm), None))
    assert np.all(P == 0.25) and np.isnan(rm[0])


@pytest.mark.parametrize("flags", [(1, 1), (1, 0)])
def test_bias_sgd_host_matches_oracle_replay(flags):
    from matrix_factorization_b200 import _lib
    from oracle import oracle as orc

    L = _lib.lib()
    U, I, N, E = 500, 250, 20_000, 3
    u, i, r, rng = _data(5 + flags[1], U, I, N)
    bu = rng.normal(0, 0.05, U).astype(np.float32)
    bi = rng.normal(0, 0.05, I).astype(np.float32)
    bu0, bi0 = bu.astype(np.float64), bi.astype(np.float64)
    mu, lr, reg = float(r.mean()), 0.01, 0.02
    rmse = np.zeros(E)
    order = np.zeros(N, dtype=np.int64)
    _lib.check(L.mfk_bias_sgd_host(_lib.ptr(u), _lib.ptr(i), _lib.ptr(r), N, U, I, _lib.ptr(bu), _lib.ptr(bi), mu, E, lr, reg,
                                   flags[0], flags[1], None, _lib.ptr(rmse), _lib.ptr(order)))
    buo, bio = bu0, bi0
    r64 = r.astype(np.float64)
    for e in range(E):
        buo, bio = orc.bias_replay(u, i, r64, order, mu, buo, bio, lr, reg, bool(flags[0]), bool(flags[1]))
        assert abs(rmse[e] - orc.bias_rmse(u, i, r64, mu, buo, bio)) < 2e-5
    assert np.max(np.abs(bu - buo)) < 2e-5 and np.max(np.abs(bi - bio)) < 2e-5
    if not flags[1]:
        assert np.array_equal(bi.astype(np.float64), bi0)


def test_bias_als_host_matches_oracle():
    from matrix_factorization_b200 import _lib
    from oracle import oracle as orc

    L = _lib.lib()
    U, I, N, E = 400, 300, 25_000, 4
    u, i, r, rng = _data(9, U, I, N)
    bu, bi = np.zeros(U, np.float32), np.zeros(I, np.float32)
    mu, reg = float(r.mean()), 0.5
    rmse = np.zeros(E)
    _lib.check(L.mfk_bias_als_host(_lib.ptr(u), _lib.ptr(i), _lib.ptr(r), N, U, I, _lib.ptr(bu), _lib.ptr(bi), mu, E, reg,
                                   _lib.ptr(rmse)))
    buo, bio, rmo = orc.bias_als(u, i, r.astype(np.float64), mu, U, I, E, reg)
    np.testing.assert_allclose(bu, buo, atol=2e-6)
    np.testing.assert_allclose(bi, bio, atol=2e-6)
    np.testing.assert_allclose(rmse, rmo, atol=2e-6)
