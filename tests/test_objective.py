'''
K2 -- objective batch (many parameter draws on the same drivers).

CPU part (`-m "not gpu"`): the draw/element factorisation of the kernel math
(mod16_b200/csrc/mod16_objective.cuh, compiled for the host by tests/hostmath) against the
numpy oracle's `_et`, one oracle call per draw -- exactly the loop the reference runs
(mod16/sensitivity.py:94-96).
GPU part: `mod16_objective_batch_{f32,f64}` through the C ABI against
`oracle.objective_stats(oracle.static_et(draw, *drivers), obs, w)` per draw.

Criteria.  float64: sse within 1e-10 relative, count exact.  float32: RMSD is the L2 norm of
the weighted residual / sqrt(n), so a per-element prediction error of at most
delta = max(1e-5 |pred|, 1e-3 W m-2) (the north-star tolerance) moves it by at most
max(w delta): |rmsd - rmsd_ref| <= max over elements of w * delta.  Counts are exact.
'''
import numpy as np
import pytest

import hostmath
import parity
from mod16_b200 import synthetic, utils
from oracle import mod16_oracle as orc

NAMES = list(synthetic.DRIVER_NAMES)


def _column(pft=7):
    table = utils.bplut_table(utils.restore_bplut(utils.BPLUT_C51), orc.PARAM_ORDER,
                              default_beta=250.0)
    return np.array(table[:, pft], np.float64)


def _problem(T, N, D, seed, kind='physical', nan_frac=0.0, fixed=True, obs_nan=0.3,
             dtype=np.float32):
    d = synthetic.host_drivers(T, N, seed, kind=kind, dtype=dtype, nan_frac=nan_frac,
                               fill_frac=nan_frac)
    drv = [d[k] for k in NAMES]
    d64 = [v.astype(np.float64) if isinstance(v, np.ndarray) else v for v in drv]
    col = _column()
    draws = synthetic.host_param_draws(
        D, seed + 1, bplut_column=col,
        fixed=('tmin_close', 'tmin_open', 'vpd_open') if fixed else ())
    rng = np.random.default_rng(seed + 2)
    with np.errstate(all='ignore'):
        truth = orc.static_et(list(col), *d64)
    obs = truth + rng.normal(0, 15, truth.shape)
    obs[obs < 0] = np.nan                       # calibration.py:631
    obs[rng.random(obs.shape) < obs_nan] = np.nan
    w = rng.uniform(0.2, 1.0, obs.shape)
    return drv, d64, draws, obs.astype(dtype), w.astype(dtype)


def _reference_stats(d64, draws, obs, w, gate='global', semantics='static'):
    'One oracle call per draw, as the reference loops do.'
    sse = np.zeros(len(draws))
    cnt = np.zeros(len(draws))
    tol = np.zeros(len(draws))
    o64, w64 = np.asarray(obs, np.float64), np.asarray(w, np.float64)
    for i, x in enumerate(draws):
        with np.errstate(all='ignore'):
            if semantics == 'static':
                pred = orc.static_et(list(x), *d64, gate=gate)
            else:
                day, night = orc.InstanceModel(dict(zip(orc.PARAM_ORDER, x))).evapotranspiration(
                    *d64)
                pred = day + night
        sse[i], cnt[i] = orc.objective_stats(pred, o64, w64)
        ok = ~np.isnan((pred - o64) * w64)
        if ok.any():
            tol[i] = np.max((w64 * np.maximum(1e-5 * np.abs(pred), 1e-3))[ok])
    return sse, cnt, tol


# ---- CPU: the factorised math -------------------------------------------------------------
@pytest.mark.parametrize('pair', [False, True])
@pytest.mark.parametrize('kind,fixed', [('physical', True), ('physical', False),
                                        ('stress', False)])
def test_factorised_math_matches_oracle_per_draw(kind, fixed, pair):
    T, N, D = 12, 700, 9
    drv, d64, draws, _, _ = _problem(T, N, D, 77, kind=kind, fixed=fixed)
    r = hostmath.objective(draws, drv, pair=pair)
    assert r['draw_ok'].all()
    assert r['fast'].mean() > 0.95
    for i, x in enumerate(draws):
        with np.errstate(all='ignore'):
            day, night = orc.static_evapotranspiration(list(x), *d64, gate='open')
            closed = orc.static_et(list(x), *d64, gate='closed')
        ref = day + night
        m = r['fast']
        parity.assert_parity(r['tot'][i][m], ref[m], np.float32, what='K2 pred draw %d' % i)
        parity.assert_parity((r['tot'][i] - r['tr'][i])[m], closed[m], np.float32,
                             what='K2 gate-closed pred draw %d' % i)


def test_bare_and_saturated_elements_stay_on_the_fast_path():
    'lai == 0 and rh == 1 switch g_canopy to the constant 1e-7 (:336-337): handled, not exact.'
    T, N = 6, 400
    d = synthetic.host_drivers(T, N, 6, dtype=np.float32)
    d['lai'][0] = 0.0
    d['lai'][1, ::3] = 0.0
    d['fpar'][0] = 0.0
    d['vpd_day'][2] = 0.0       # rh == 1 by day
    d['vpd_night'][3] = 0.0
    drv = [d[k] for k in NAMES]
    d64 = [v.astype(np.float64) if isinstance(v, np.ndarray) else v for v in drv]
    draws = synthetic.host_param_draws(5, 3, bplut_column=_column(), fixed=())
    for pair in (False, True):
        r = hostmath.objective(draws, drv, pair=pair)
        assert r['fast'][:4].mean() > 0.99
        for i, x in enumerate(draws):
            with np.errstate(all='ignore'):
                ref = orc.static_et(list(x), *d64, gate='open')
                closed = orc.static_et(list(x), *d64, gate='closed')
            m = r['fast']
            parity.assert_parity(r['tot'][i][m], ref[m], np.float32, what='bare/saturated %d' % i)
            parity.assert_parity((r['tot'][i] - r['tr'][i])[m], closed[m], np.float32,
                                 what='bare/saturated closed %d' % i)


def test_elements_outside_the_fast_domain_get_no_record():
    T, N = 6, 400
    d = synthetic.host_drivers(T, N, 5, dtype=np.float32, nan_frac=0.01, fill_frac=0.01)
    drv = [d[k] for k in NAMES]
    r = hostmath.objective(_column()[None], drv)
    bad = np.zeros((T, N), bool)
    for k in NAMES:
        # (a NaN / fill annual temperature only fails the window test, :230-233: finite result)
        if isinstance(d[k], np.ndarray) and k != 'temp_annual':
            bad |= np.broadcast_to(np.isnan(d[k]) | (d[k] == -9999), (T, N))
    assert not r['fast'][bad].any()
    assert r['fast'].mean() > 0.7


def test_degenerate_draws_are_flagged():
    col = _column()
    draws = np.stack([col, col, col, col])
    draws[1, 10] = np.nan          # beta
    draws[2, 1] = draws[2, 0]      # zero-width tmin ramp
    draws[3, 4] = 0.0              # gl_sh
    drv = [synthetic.host_drivers(2, 8, 1, dtype=np.float32)[k] for k in NAMES]
    r = hostmath.objective(draws, drv)
    assert list(r['draw_ok']) == [True, False, False, False]


# ---- GPU: through the C ABI ---------------------------------------------------------------
def _check(sse, cnt, ref_sse, ref_cnt, tol, dtype):
    assert np.array_equal(cnt, ref_cnt), 'valid-pair counts differ'
    if np.dtype(dtype) == np.float64:
        assert np.all(np.abs(sse - ref_sse) <= 1e-10 * np.maximum(ref_sse, 1.0))
    else:
        with np.errstate(invalid='ignore', divide='ignore'):
            r0, r1 = np.sqrt(sse / cnt), np.sqrt(ref_sse / ref_cnt)
        ok = ref_cnt > 0
        assert np.all(np.abs(r0 - r1)[ok] <= tol[ok]), (np.abs(r0 - r1)[ok] / tol[ok]).max()


@pytest.mark.gpu
@pytest.mark.parametrize('dtype', [np.float32, np.float64])
@pytest.mark.parametrize('kind,nan,fixed', [('physical', 0.0, True), ('physical', 0.002, False),
                                            ('stress', 0.002, False)])
def test_objective_batch_matches_oracle_loop(cuda_lib, dtype, kind, nan, fixed):
    from mod16_b200.objective import objective_batch
    T, N, D = 37, 301, 70
    drv, d64, draws, obs, w = _problem(T, N, D, 404, kind=kind, nan_frac=nan, fixed=fixed,
                                       dtype=dtype)
    ref_sse, ref_cnt, tol = _reference_stats(d64, draws, obs, w)
    sse, cnt = objective_batch(draws, drv, obs, w, dtype=dtype)
    _check(sse, cnt, ref_sse, ref_cnt, tol, dtype)
    # unweighted, observations complete
    obs2 = np.where(np.isnan(obs), 50.0, obs).astype(dtype)
    ref_sse, ref_cnt, tol = _reference_stats(d64, draws, obs2, 1.0)
    sse, cnt = objective_batch(draws, drv, obs2, None, dtype=dtype)
    _check(sse, cnt, ref_sse, ref_cnt, tol, dtype)


@pytest.mark.gpu
@pytest.mark.parametrize('dtype', [np.float32, np.float64])
def test_objective_batch_flat_calibration_layout(cuda_lib, dtype):
    'The calibration passes 1-D drivers with the NaN observations already removed.'
    from mod16_b200.objective import ObjectiveBackend
    T, N, D = 50, 120, 200
    drv, d64, draws, obs, w = _problem(T, N, D, 11, dtype=dtype)
    keep = ~np.isnan(obs)
    flat = [v[keep] if isinstance(v, np.ndarray) and v.shape == obs.shape else
            (np.broadcast_to(v, obs.shape)[keep] if isinstance(v, np.ndarray) else v)
            for v in drv]
    f64 = [v.astype(np.float64) if isinstance(v, np.ndarray) else v for v in flat]
    ref_sse, ref_cnt, tol = _reference_stats(f64, draws[:40], obs[keep], w[keep])
    backend = ObjectiveBackend(flat, obs[keep], w[keep], dtype=dtype)
    sse, cnt = backend(draws)
    _check(sse[:40], cnt[:40], ref_sse, ref_cnt, tol, dtype)
    # repeatable bit for bit (deterministic reductions), and independent of the batch split
    sse2, cnt2 = backend(draws)
    assert np.array_equal(sse, sse2) and np.array_equal(cnt, cnt2)
    a, _ = backend(draws[:64])
    b, _ = backend(draws[64:])
    if np.dtype(dtype) == np.float64:
        assert np.array_equal(np.concatenate([a, b]), sse)
    else:
        assert np.allclose(np.concatenate([a, b]), sse, rtol=1e-6)
    nr = backend.neg_rmsd(draws[:3])
    assert np.allclose(nr, -np.sqrt(sse[:3] / cnt[:3]))


@pytest.mark.gpu
@pytest.mark.parametrize('dtype', [np.float32, np.float64])
def test_objective_batch_per_draw_gate(cuda_lib, dtype):
    '''
    `if np.any(g_surf > 0)` is evaluated per reference call, i.e. per draw: a draw with
    csl == 0 (or a Tmin ramp that never opens) loses its cuticular transpiration too.
    '''
    from mod16_b200.objective import objective_batch
    T, N, D = 9, 257, 40
    drv, d64, draws, obs, w = _problem(T, N, D, 31, dtype=dtype, fixed=False)
    draws[3, 7] = 0.0                         # csl = 0: g_surf == 0 everywhere
    draws[5, 0], draws[5, 1] = 60.0, 80.0     # Tmin ramp above every Tmin
    for gate in ('global', 'open', 'closed'):
        ref_sse, ref_cnt, tol = _reference_stats(d64, draws, obs, w, gate=gate)
        sse, cnt, g = objective_batch(draws, drv, obs, w, dtype=dtype, gate=gate,
                                      return_gate=True)
        _check(sse, cnt, ref_sse, ref_cnt, tol, dtype)
        assert not g[3] and not g[5] and g[[0, 1, 2, 4]].all()


@pytest.mark.gpu
@pytest.mark.parametrize('mask', ['observation', 'weight'])
@pytest.mark.parametrize('dtype', [np.float32, np.float64])
def test_objective_batch_gate_sees_elements_without_an_observation(cuda_lib, dtype, mask):
    '''
    `np.any(g_surf > 0)` (mod16/__init__.py:343) runs over every element of the draw's `_et`
    call; the observation mask is applied afterwards, in the score (calibration.py:224-227).
    So an element WITHOUT a valid observation (or weight) still opens the gate.  Constructed
    case: two draws whose Tmin ramp opens on the five warmest-Tmin elements only, and exactly
    those five have no observation (second configuration: a valid observation but no weight).
    Every valid element then has g_surf == 0, yet the reference keeps its cuticular
    transpiration: its score is the open gate's, which differs from the closed gate's.
    '''
    from mod16_b200.objective import ObjectiveBackend, finish_raw
    T, N, D = 9, 257, 24
    drv, d64, draws, obs, w = _problem(T, N, D, 61, dtype=dtype, fixed=False, obs_nan=0.2)
    obs, w = obs.copy(), w.copy()
    tmin_c = np.asarray(d64[NAMES.index('tmin')], np.float64) - 273.15
    order = np.argsort(tmin_c, axis=None)
    top = np.unravel_index(order[-5:], tmin_c.shape)
    thr = 0.5 * (tmin_c.flat[order[-6]] + tmin_c.flat[order[-5]])
    assert tmin_c.flat[order[-5]] - tmin_c.flat[order[-6]] > 1e-3      # a clean threshold
    k_a, k_b, k_closed = 4, 9, 13
    for k in (k_a, k_b):
        draws[k, 0], draws[k, 1] = thr, thr + 10.0      # tmin_close, tmin_open [deg C]
        draws[k, 2], draws[k, 3] = 650.0, 1.0e5         # vpd_open, vpd_close: never closes
        assert draws[k, 7] > 0                          # csl
    draws[k_closed, 7] = 0.0                            # csl = 0: closed on every element
    if mask == 'observation':
        obs[top] = np.nan
        assert np.isfinite(w[top]).all()
    else:
        obs[top] = 100.0                                # valid: the weight alone drops them
        w[top] = np.nan
    ref = {}
    for gate in ('global', 'open', 'closed'):
        ref[gate] = _reference_stats(d64, draws, obs, w, gate=gate)
    for k in (k_a, k_b):
        # the case discriminates: open and closed gates give different scores, and the
        # reference's (global) gate is the open one
        assert ref['global'][0][k] == ref['open'][0][k]
        assert abs(ref['open'][0][k] - ref['closed'][0][k]) > 1e-3 * ref['open'][0][k]
    assert ref['global'][0][k_closed] == ref['closed'][0][k_closed]
    backend = ObjectiveBackend(drv, obs, w, dtype=dtype)
    sse, cnt, g = backend(draws, return_gate=True)
    _check(sse, cnt, ref['global'][0], ref['global'][1], ref['global'][2], dtype)
    assert g[k_a] and g[k_b] and not g[k_closed]
    # the un-gated statistics (the site-sharded form) carry the same gate
    s2, c2, g2 = finish_raw(backend(draws, raw=True))
    _check(s2.cpu().numpy(), c2.cpu().numpy(), ref['global'][0], ref['global'][1],
           ref['global'][2], dtype)
    assert np.array_equal(g2.cpu().numpy().astype(bool), np.asarray(g, bool))


@pytest.mark.gpu
def test_objective_batch_degenerate_draws_and_small_batches(cuda_lib):
    from mod16_b200.objective import objective_batch
    T, N, D = 11, 200, 48
    drv, d64, draws, obs, w = _problem(T, N, D, 8, fixed=False)
    draws[1, 10] = np.nan          # beta NaN: every prediction NaN, nothing is valid
    draws[2, 1] = draws[2, 0]      # zero-width Tmin ramp: double code decides
    draws[7, 3] = draws[7, 2]      # zero-width VPD ramp
    ref_sse, ref_cnt, tol = _reference_stats(d64, draws, obs, w)
    sse, cnt = objective_batch(draws, drv, obs, w, dtype=np.float32)
    assert cnt[1] == 0 and ref_cnt[1] == 0
    _check(sse, cnt, ref_sse, ref_cnt, tol, np.float32)
    # D below the fast-path threshold, D == 1 (a sequential MCMC proposal), D == 0
    for k in (1, 5):
        s, c = objective_batch(draws[3:3 + k], drv, obs, w, dtype=np.float32)
        _check(s, c, ref_sse[3:3 + k], ref_cnt[3:3 + k], tol[3:3 + k], np.float32)
    s, c = objective_batch(np.zeros((0, 11)), drv, obs, w, dtype=np.float32)
    assert s.shape == (0,) and c.shape == (0,)


@pytest.mark.gpu
@pytest.mark.parametrize('dtype', [np.float32, np.float64])
def test_objective_batch_instance_semantics(cuda_lib, dtype):
    from mod16_b200.objective import objective_batch
    T, N, D = 7, 150, 6
    drv, d64, draws, obs, w = _problem(T, N, D, 3, dtype=dtype)
    obs = (obs / 2.45e6).astype(dtype)
    ref_sse, ref_cnt, _ = _reference_stats(d64, draws, obs, w, semantics='instance')
    sse, cnt = objective_batch(draws, drv, obs, w, dtype=dtype, semantics='instance')
    assert np.array_equal(cnt, ref_cnt)
    assert np.all(np.abs(sse - ref_sse) <= 1e-10 * ref_sse)


@pytest.mark.gpu
def test_objective_batch_sobol_sized_slice(cuda_lib):
    '''
    A slice of BASELINE config 5 (parameter rows x 1000 pixels x 365 days): size-independent
    properties at a size the oracle cannot loop over -- the fast and the double path agree
    within the RMSD bound, the statistics of a draw do not depend on its batch.
    '''
    from mod16_b200.objective import ObjectiveBackend
    T, N, D = 365, 1000, 2048
    drv, d64, draws, obs, w = _problem(T, N, D, 16005, fixed=False)
    fast = ObjectiveBackend(drv, obs, w, dtype=np.float32)
    s32, c32 = fast(draws)
    exact = ObjectiveBackend([v.astype(np.float64) if isinstance(v, np.ndarray) else v
                              for v in drv], obs.astype(np.float64), w.astype(np.float64),
                             dtype=np.float64)
    pick = np.arange(0, D, 97)
    s64, c64 = exact(draws[pick])
    assert np.array_equal(c32[pick], c64)
    r32, r64 = np.sqrt(s32[pick] / c32[pick]), np.sqrt(s64 / c64)
    assert np.all(np.abs(r32 - r64) <= np.maximum(1e-5 * r64, 1e-3)), np.abs(r32 - r64).max()
    ref_sse, ref_cnt, tol = _reference_stats(d64, draws[:2], obs, w)
    _check(s32[:2], c32[:2], ref_sse, ref_cnt, tol, np.float32)


@pytest.mark.gpu
@pytest.mark.parametrize('dtype', [np.float32, np.float64])
def test_raw_statistics_and_site_sharding_on_one_gpu(cuda_lib, dtype):
    '''
    The un-gated (D, 5) statistics: finishing them reproduces (sse, count), and summing the
    statistics of two site shards before the gate is applied equals the unsharded call --
    including a draw whose gate is open on one shard only.
    '''
    import torch
    from mod16_b200.objective import ObjectiveBackend, finish_raw
    T, N, D = 9, 256, 40
    drv, d64, draws, obs, w = _problem(T, N, D, 52, dtype=dtype, fixed=False)
    drv[8] = drv[8].copy()
    drv[8][:, :N // 2] = 230.0                 # shard 0 never opens the Tmin ramp (< -35 C)
    draws[3, 7] = 0.0                          # csl = 0: closed everywhere
    whole = ObjectiveBackend(drv, obs, w, dtype=dtype)
    sse, cnt, gate = whole(draws, return_gate=True)
    raw = whole(draws, raw=True)
    s2, c2, g2 = finish_raw(raw)
    # (a draw no valid element opens is redone by the double kernel's pass over all elements
    # in both calls -- only that pass sees the elements without an observation, which are
    # part of the gate; the tolerance is kept for the other draws of the float32 raw call)
    assert np.allclose(s2.cpu().numpy(), sse, rtol=1e-6 if dtype == np.float32 else 0, atol=0)
    assert np.array_equal(c2.cpu().numpy(), cnt)
    assert np.array_equal(g2.cpu().numpy(), gate) and not gate[3]
    parts = []
    for lo, hi in ((0, N // 2), (N // 2, N)):
        sl = [v[:, lo:hi] if isinstance(v, np.ndarray) and v.ndim == 2 else
              (v[lo:hi] if isinstance(v, np.ndarray) and v.ndim == 1 else v) for v in drv]
        parts.append(ObjectiveBackend(sl, obs[:, lo:hi], w[:, lo:hi], dtype=dtype)(draws, raw=True))
    assert float(parts[0][:, 4].max()) == 0.0 and float(parts[1][:, 4].max()) == 1.0
    s3, c3, _ = finish_raw(parts[0] + parts[1])
    assert np.array_equal(c3.cpu().numpy(), cnt)
    assert np.allclose(s3.cpu().numpy(), sse, rtol=1e-6 if dtype == np.float32 else 1e-12)
    # gating each shard on its own would be wrong for shard 0 (it would drop transpiration)
    s0, _, _ = finish_raw(parts[0])
    assert not np.allclose(s0.cpu().numpy() + finish_raw(parts[1])[0].cpu().numpy(), sse,
                           rtol=1e-6)


# ---- the objectives' definitions, pinned by contract ------------------------------------------
# mod17 (EXTERNAL: PyPI `mod17>=0.1.1`, pyproject.toml:22; not under /root/reference) documents
#   BlackBoxLikelihood.loglik(params, observed, x):  predicted = model(params, *x);
#       -sqrt(nanmean(((predicted - observed) * weights) ** 2))        [objective "rmsd"]
#   science.nash_sutcliffe(predicted, observed, norm = False):
#       nse = 1 - nansum((observed - predicted)**2) / nansum((observed - nanmean(observed))**2)
#       norm -> 1 / (2 - nse)
# The call sites are mod16/calibration.py:224-227 and mod16/sensitivity.py:94-96.  The vectors
# below are worked by hand from those formulas, so that the definition this repository
# implements is written down and checked in one place.
def test_objective_definitions_known_answers():
    from mod16_b200 import objective as obj
    obs = np.array([1.0, 2.0, 3.0, 4.0])
    pred = np.array([1.0, 3.0, 2.0, 6.0])
    sse, n = orc.objective_stats(pred, obs)
    assert (sse, n) == (6.0, 4)                                   # 0 + 1 + 1 + 4
    assert obj.rmsd(sse, n) == np.sqrt(1.5)
    assert orc.rmsd_from_stats(sse, n) == np.sqrt(1.5)
    # sum((obs - 2.5)^2) = 5  ->  NSE = 1 - 6 / 5 = -0.2,  normalised 1 / 2.2
    assert abs(obj.nash_sutcliffe(sse, obs) - (-0.2)) < 1e-15
    assert abs(obj.nash_sutcliffe(sse, obs, norm=True) - 1 / 2.2) < 1e-15
    assert abs(orc.nse_from_stats(sse, obs, norm=True) - 1 / 2.2) < 1e-15
    # missing values: a pair counts when both are present (nanmean / nansum of the squared
    # error); the NSE denominator runs over every observation that is present
    obs = np.array([1.0, np.nan, 3.0])
    pred = np.array([2.0, 5.0, np.nan])
    sse, n = orc.objective_stats(pred, obs)
    assert (sse, n) == (1.0, 1) and obj.rmsd(sse, n) == 1.0
    assert obj.nash_sutcliffe(sse, obs) == 0.5                    # 1 - 1 / ((1-2)^2 + (3-2)^2)
    # weights multiply the error before it is squared
    sse, n = orc.objective_stats(np.array([1.0, 2.0]), np.array([0.0, 0.0]), np.array([2.0, 1.0]))
    assert (sse, n) == (8.0, 2) and obj.rmsd(sse, n) == 2.0
    # no valid pair: NaN, not an exception
    assert np.isnan(obj.rmsd(0.0, 0.0))
    # a perfect model: NSE = 1 either way
    assert obj.nash_sutcliffe(0.0, np.array([1.0, 2.0])) == 1.0
    assert obj.nash_sutcliffe(0.0, np.array([1.0, 2.0]), norm=True) == 1.0


@pytest.mark.gpu
def test_calibration_config_at_full_size(cuda_lib):
    '''
    BASELINE config 4 at its full size: 10^5 parameter draws x 300 tower sites x 3652 days
    (10 years of daily data), 30 % of the observations missing -- the batch the calibration's
    likelihood loop amounts to (calibration.py:224-227, one `_et` call per MCMC proposal).
    The oracle cannot loop over 10^5 draws: a sample of draws is checked against one oracle
    call each (valid-pair counts exactly, RMSD within the float32 bound), and size-independent
    properties hold for all of them -- every draw sees the same valid pairs, the statistics
    of a draw do not depend on the batch it is evaluated in, a repeated call is bit-identical.
    '''
    from mod16_b200.objective import ObjectiveBackend
    T, N, D = 3652, 300, 100000
    drv, d64, draws, obs, w = _problem(T, N, D, 16004)
    backend = ObjectiveBackend(drv, obs, w, dtype=np.float32)
    sse, cnt = backend(draws)
    assert sse.shape == (D,) and np.isfinite(sse).all()
    valid = int((~np.isnan(obs)).sum())
    assert 0.6 * T * N < valid < 0.75 * T * N
    assert (cnt == valid).all()                 # physical drivers: no NaN predictions
    pick = np.array([0, 1, 4999, 31337, 65536, 99999])
    ref_sse, ref_cnt, tol = _reference_stats(d64, draws[pick], obs, w)
    _check(sse[pick], cnt[pick], ref_sse, ref_cnt, tol, np.float32)
    s2, c2 = backend(draws)                     # no floating-point atomics anywhere
    assert np.array_equal(s2, sse) and np.array_equal(c2, cnt)
    s3, c3 = backend(draws[pick])               # a batch of six takes the double-precision kernel
    r_big, r_small = np.sqrt(sse[pick] / cnt[pick]), np.sqrt(s3 / c3)
    assert np.array_equal(c3, cnt[pick]) and np.all(np.abs(r_big - r_small) <= tol)
    rm = np.sqrt(sse / cnt)
    assert 10.0 < rm.min() < rm.max() < 1000.0  # plausible RMSD range, W m-2


@pytest.mark.parametrize('pair', [False, True])
def test_factorised_math_at_the_discontinuities(pair):
    '''
    K2's own arithmetic on the elements of tests/switch_cases.py (float32, 1e-8 ... 1e-3 on
    both sides of every switch of the model).  Draw 0 is the parameter column the elements
    were placed for, so the draw-dependent bound of the annual-temperature window (an exact
    float comparison, mod16_objective.cuh: tann_lo) and the ends of the draw's ramps fall ON
    elements; the other draws move them.  Every element that keeps a fast record must match
    the oracle for every draw -- gate open and gate closed -- and the elements within rounding
    of a draw-independent switch must have lost theirs.
    '''
    import switch_cases
    drv, groups, grp, off = switch_cases.build('static', np.float32)
    d64 = [np.asarray(v, np.float64) for v in drv]
    col = np.array(switch_cases.BPLUT7)
    draws = np.vstack([col[None, :], synthetic.host_param_draws(7, 321, bplut_column=col)])
    r = hostmath.objective(draws, drv, pair=pair)
    assert r['draw_ok'].all()
    m = r['fast']
    for i, x in enumerate(draws):
        with np.errstate(all='ignore'):
            day, night = orc.static_evapotranspiration(list(x), *d64, gate='open')
            closed = orc.static_et(list(x), *d64, gate='closed')
        parity.assert_parity(r['tot'][i][m], (day + night)[m], np.float32,
                             what='K2 pred at the switches, draw %d' % i)
        parity.assert_parity((r['tot'][i] - r['tr'][i])[m], closed[m], np.float32,
                             what='K2 gate-closed pred at the switches, draw %d' % i)
    # the draw-independent switches: within rounding -> no fast record; far away -> kept
    fast = m.reshape(-1)
    near, far = np.abs(off) <= 1e-7, np.abs(off) >= 1e-4
    for name in ('rh07', 'g_day', 'g_night', 'a_zero'):
        k = grp == groups.index(name)
        assert fast[k & near].mean() < 0.5 and fast[k & far].mean() > 0.9, (
            name, fast[k & near].mean(), fast[k & far].mean())
    print('fast records kept at the switches: %.3f' % fast.mean())
