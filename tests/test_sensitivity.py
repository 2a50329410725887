'''
Sobol' sensitivity analysis (mod16_b200/sensitivity.py) against the restatement of SALib's
sampling and estimators (oracle/salib_oracle.py) and, for the restatement itself, against the
analytic indices of the Ishigami function -- SALib is absent, see the oracle's header.
'''
import numpy as np
import pytest
import torch

from mod16_b200 import sensitivity as S
from mod16_b200 import synthetic
from oracle import mod16_oracle as orc
from oracle import salib_oracle as sal

NAMES = list(synthetic.DRIVER_NAMES)
BPLUT7 = [-8., 8.8, 650., 4400., .02, .02, 1e-5, .0055, 60., 95., 250.]


def _problem(bounds):
    return {'num_vars': len(bounds), 'bounds': [list(b) for b in bounds]}


def test_oracle_sample_layout():
    D, N = 4, 8
    prob = _problem([(0, 1)] * D)
    x = sal.sample(prob, N, scramble=False).reshape(N, 2 * D + 2, D)
    A, B = x[:, 0], x[:, -1]
    for k in range(D):
        ab, ba = x[:, 1 + k], x[:, 1 + D + k]
        for j in range(D):
            assert np.array_equal(ab[:, j], B[:, j] if j == k else A[:, j])
            assert np.array_equal(ba[:, j], A[:, j] if j == k else B[:, j])
    x1 = sal.sample(prob, N, calc_second_order=False, scramble=False).reshape(N, D + 2, D)
    assert np.array_equal(x1[:, :D + 1], x[:, :D + 1]) and np.array_equal(x1[:, -1], B)
    # bounds are applied per column
    y = sal.sample(_problem([(-1, 1), (10, 20), (0, 1), (5, 5.5)]), N, scramble=False)
    assert y[:, 1].min() >= 10 and y[:, 1].max() <= 20 and y[:, 0].min() >= -1


def test_oracle_estimators_recover_the_ishigami_indices():
    prob = _problem([(-np.pi, np.pi)] * 3)
    x = sal.sample(prob, 8192, seed=7)
    got = sal.analyze(prob, sal.ishigami(x), num_resamples=50, seed=1)
    s1, st = sal.ishigami_indices()
    assert np.abs(got['S1'] - s1).max() < 0.02, got['S1']
    assert np.abs(got['ST'] - st).max() < 0.02, got['ST']
    assert abs(got['S2'][0, 2] - (st[0] - s1[0])) < 0.03      # the x1-x3 interaction
    assert abs(got['S2'][0, 1]) < 0.03 and abs(got['S2'][1, 2]) < 0.03
    assert (got['S1_conf'] < 0.05).all() and (got['ST_conf'] < 0.05).all()
    assert np.isnan(got['S2'][1, 0]) and np.isnan(got['S2'][0, 0])


@pytest.mark.parametrize('second', [True, False])
def test_product_sample_and_estimators_equal_the_oracle_on_cpu_tensors(second):
    bounds = list(synthetic.PARAM_BOUNDS.values())
    prob = _problem(bounds)
    for scramble, seed in ((False, None), (True, 11)):
        want = sal.sample(prob, 64, calc_second_order=second, scramble=scramble, seed=seed)
        got = S.saltelli_sample(bounds, 64, calc_second_order=second, scramble=scramble,
                                seed=seed, device='cpu').numpy()
        assert got.shape == want.shape and np.array_equal(got, want)
    want = sal.sample(prob, 32, scramble=False, skip_values=64)
    got = S.saltelli_sample(synthetic.PARAM_BOUNDS, 32, scramble=False, skip_values=64,
                            device='cpu').numpy()
    assert np.array_equal(got, want)
    rng = np.random.default_rng(3)
    D, N = 11, 64
    Y = rng.normal(size=N * ((2 * D + 2) if second else (D + 2))) ** 3
    r = rng.integers(N, size=(N, 40))
    want = sal.analyze(prob, Y, calc_second_order=second, resample_index=r)
    got = S.sobol_indices(torch.from_numpy(Y), D, calc_second_order=second, resample_index=r)
    assert list(got) == list(want)
    for k in want:
        assert np.allclose(got[k], want[k], rtol=1e-9, atol=1e-12, equal_nan=True), k
    with pytest.raises(RuntimeError):
        S.sobol_indices(Y[:-1], D, calc_second_order=second)


@pytest.mark.gpu
def test_parameter_sensitivity_like_the_reference_loop(cuda_lib):
    T, N = 40, 24
    d = synthetic.host_drivers(T, N, 301, dtype=np.float32)
    drv = [d[k] for k in NAMES]
    d64 = [v.astype(np.float64) if isinstance(v, np.ndarray) else v for v in drv]
    rng = np.random.default_rng(302)
    with np.errstate(all='ignore'):
        obs = (orc.static_et(BPLUT7, *d64) + rng.normal(0, 10, (T, N))).astype(np.float32)
    Si, X, Y = S.parameter_sensitivity(drv, obs, N=16, seed=5, num_resamples=20,
                                       return_output=True)
    X, Y = X.cpu().numpy(), Y.cpu().numpy()
    assert X.shape == (16 * 24, 11) and Y.shape == (16 * 24,)
    prob = _problem(list(synthetic.PARAM_BOUNDS.values()))
    assert np.array_equal(X, sal.sample(prob, 16, seed=5))
    o64 = obs.astype(np.float64)
    for i in rng.integers(0, len(X), 12):       # sensitivity.py:94-96, one call per row
        with np.errstate(all='ignore'):
            yhat = orc.static_et(list(X[i]), *d64)
        sse, _ = orc.objective_stats(yhat, o64)
        want = orc.nse_from_stats(sse, o64, norm=True)
        assert abs(Y[i] - want) <= 2e-5 * abs(want), (i, Y[i], want)
    want = sal.analyze(prob, Y, num_resamples=20, seed=0)
    for k in ('S1', 'ST', 'S2'):
        assert np.allclose(Si[k], want[k], rtol=1e-9, atol=1e-12, equal_nan=True), k


@pytest.mark.gpu
def test_driver_sensitivity_like_the_reference_loop(cuda_lib):
    Si, X, Y = S.driver_sensitivity(BPLUT7, N=32, seed=9, num_resamples=10, dtype=np.float64,
                                    return_output=True)
    X, Y = X.cpu().numpy(), Y.cpu().numpy()
    assert X.shape == (32 * 30, 14) and Y.shape == (32 * 30,)
    for i in np.random.default_rng(1).integers(0, len(X), 40):   # sensitivity.py:113-114
        with np.errstate(all='ignore'):
            want = orc.static_et(BPLUT7, *[float(v) for v in X[i]])
        assert abs(Y[i] - want) <= 1e-10 * max(abs(want), 1.0), (i, Y[i], want)
    prob = _problem(list(synthetic.BOUNDS.values()))
    want = sal.analyze(prob, Y, num_resamples=10, seed=0)
    for k in ('S1', 'ST'):
        assert np.allclose(Si[k], want[k], rtol=1e-9, atol=1e-12), k
    assert np.isfinite(Si['ST']).all() and Si['ST'].max() > 0.05
