'''
Checks the kernels' per-element math (mod16_b200/csrc/*.cuh, compiled for the host by
tests/hostmath) against the oracle, where no GPU exists.  The GPU parity tests proper are
in test_forward_parity.py; this file keeps the device math honest between GPU runs.
'''
import os

import numpy as np
import pytest

import hostmath
import parity
from mod16_b200 import synthetic
from oracle import mod16_oracle as orc

NAMES = list(synthetic.DRIVER_NAMES)
BPLUT7 = [-8., 8.8, 650., 4400., .02, .02, 1e-5, .0055, 60., 95., 250.]


def totals(r, sem):
    d, n = r['day'], r['night']
    if sem == 'static':
        return (d[2] + d[0]) + d[1], (n[2] + n[0]) + n[1]
    return (d[0] + d[1]) + d[2], (n[0] + n[1]) + n[2]


@pytest.mark.parametrize('kind,nan,fill', [('physical', 0, 0), ('physical', .01, .01),
                                           ('stress', .01, .01)])
def test_mirror_f64_matches_oracle(kind, nan, fill):
    d = synthetic.host_drivers(16, 200, 77, kind=kind, nan_frac=nan, fill_frac=fill)
    drv = [d[k] for k in NAMES]
    rng = np.random.default_rng(3)
    par = [np.asarray(v) * rng.uniform(0.8, 1.2, (1, 200)) for v in BPLUT7]
    with np.errstate(all='ignore'):
        od, on = orc.static_evapotranspiration(par, *drv)
    r = hostmath.forward(par, drv, 'static', 'global')
    hd, hn = totals(r, 'static')
    parity.assert_parity(hd, od, np.float64, what='static day')
    parity.assert_parity(hn, on, np.float64, what='static night')
    P = dict(zip(orc.PARAM_ORDER, BPLUT7))
    with np.errstate(all='ignore'):
        sep = orc.InstanceModel(P).evapotranspiration(*drv, separate=True)
    r = hostmath.forward(BPLUT7, drv, 'instance')
    for h, hn_ in enumerate(('day', 'night')):
        for c in range(3):
            parity.assert_parity(r[hn_][c], sep[h][c], np.float64, parity.MASS_SCALE,
                                 what='instance %s %d' % (hn_, c))


def test_mirror_f64_tables(golden_dir):
    z = np.load(os.path.join(golden_dir, 'tables.npz'))
    drv = [z['drivers'][:, i] for i in range(14)]
    par = list(z['params'])
    r = hostmath.forward(par, drv, 'static', 'element')
    hd, hn = totals(r, 'static')
    parity.assert_parity(hd, z['static_day'], np.float64, what='table static day')
    parity.assert_parity(hn, z['static_night'], np.float64, what='table static night')
    r = hostmath.forward(par, drv, 'instance')
    for hn_ in ('day', 'night'):
        for c, cn in enumerate(('canopy', 'soil', 'trans')):
            parity.assert_parity(r[hn_][c], z['inst_%s_%s' % (hn_, cn)], np.float64,
                                 parity.MASS_SCALE, what='table inst %s %s' % (hn_, cn))


def test_mirror_gate_modes():
    d = synthetic.host_drivers(6, 50, 9, kind='stress')
    drv = [d[k] for k in NAMES]
    for gate in ('element', 'open'):
        with np.errstate(all='ignore'):
            od, on = orc.static_evapotranspiration(BPLUT7, *drv, gate=gate)
        r = hostmath.forward(BPLUT7, drv, 'static', gate)
        hd, hn = totals(r, 'static')
        parity.assert_parity(hd, od, np.float64, what=gate)
        parity.assert_parity(hn, on, np.float64, what=gate)
    # whole batch closed: tmin far below tmin_close
    drv[NAMES.index('tmin')] = np.full((6, 50), 240.0)
    with np.errstate(all='ignore'):
        od, _ = orc.static_evapotranspiration(BPLUT7, *drv)
    r = hostmath.forward(BPLUT7, drv, 'static', 'global')
    assert not r['g_pos'].any()
    parity.assert_parity(totals(r, 'static')[0], od, np.float64, what='closed batch')


def test_mirror_r_corr_list():
    d = synthetic.host_drivers(4, 40, 11)
    drv = [d[k] for k in NAMES]
    rc = [orc.r_correction(d['pressure'], d['temp_day']) * 1.1,
          orc.r_correction(d['pressure'], d['temp_night']) * 0.9]
    with np.errstate(all='ignore'):
        od, on = orc.static_evapotranspiration(BPLUT7, *drv, r_corr_list=rc)
    r = hostmath.forward(BPLUT7, drv, 'static', 'global', r_corr_list=rc)
    hd, hn = totals(r, 'static')
    parity.assert_parity(hd, od, np.float64, what='r_corr day')
    parity.assert_parity(hn, on, np.float64, what='r_corr night')


# ---- float32 fast path (mod16_fast.cuh): algebra, guards and NaN masks ------------------
# On the host the MUFU approximations are replaced by libm calls, so this checks the
# rearranged formulas and the three guards; the approximations' accuracy is covered by the
# GPU tests.

def _table():
    from mod16_b200 import utils
    return utils.bplut_table(utils.restore_bplut(utils.BPLUT_C51), orc.PARAM_ORDER,
                             default_beta=250.0)


FAST_MODES = [True, 'pair']   # scalar (float) and packed-pair (F2) instantiations


@pytest.mark.parametrize('fast', FAST_MODES)
@pytest.mark.parametrize('kind,nan', [('physical', 0.0), ('stress', 0.0),
                                      ('physical', 0.003), ('stress', 0.003)])
def test_fast_f32_matches_f64_oracle(kind, nan, fast):
    T, N = 40, 5000
    d = synthetic.host_drivers(T, N, 321, kind=kind, dtype=np.float32, nan_frac=nan,
                               fill_frac=nan)
    drv = [d[k] for k in NAMES]
    d64 = [v.astype(np.float64) if isinstance(v, np.ndarray) else v for v in drv]
    table = _table()
    pft = synthetic.host_pft_map(N, 5)
    vec = [table[i][pft].reshape(1, N) for i in range(11)]
    with np.errstate(all='ignore'):
        od, on = orc.static_evapotranspiration(vec, *d64)
    r = hostmath.forward(vec, drv, 'static', 'global', dtype=np.float32, fast=fast)
    parity.assert_parity(r['tot_day'], od, np.float32, what='fast static day')
    parity.assert_parity(r['tot_night'], on, np.float32, what='fast static night')
    parity.assert_parity(r['tot'], od + on, np.float32, what='fast static sum')
    col = list(table[:, 7])
    with np.errstate(all='ignore'):
        sep = orc.InstanceModel(dict(zip(orc.PARAM_ORDER, col))).evapotranspiration(
            *d64, separate=True)
    r = hostmath.forward(col, drv, 'instance', dtype=np.float32, fast=fast)
    for h, hn_ in enumerate(('day', 'night')):
        for c in range(3):
            parity.assert_parity(r[hn_][c], sep[h][c], np.float32, parity.MASS_SCALE,
                                 what='fast instance %s %d' % (hn_, c))


@pytest.mark.parametrize('fast', FAST_MODES)
def test_fast_f32_tables(golden_dir, fast):
    z = np.load(os.path.join(golden_dir, 'tables.npz'))
    vec = list(z['params'])
    drv = [z['drivers'][:, i].astype(np.float32) for i in range(14)]
    d64 = [v.astype(np.float64) for v in drv]
    with np.errstate(all='ignore'):
        od, on = orc.static_evapotranspiration(vec, *d64, gate='element')
        sep = orc.InstanceModel(dict(zip(orc.PARAM_ORDER, vec))).evapotranspiration(
            *d64, separate=True)
    r = hostmath.forward(vec, drv, 'static', 'element', dtype=np.float32, fast=fast)
    parity.assert_parity(r['tot_day'], od, np.float32, what='fast table day')
    parity.assert_parity(r['tot_night'], on, np.float32, what='fast table night')
    r = hostmath.forward(vec, drv, 'instance', dtype=np.float32, fast=fast)
    for h, hn_ in enumerate(('day', 'night')):
        for c in range(3):
            parity.assert_parity(r[hn_][c], sep[h][c], np.float32, parity.MASS_SCALE,
                                 what='fast table inst %s %d' % (hn_, c))


@pytest.mark.parametrize('fast', FAST_MODES)
def test_fast_f32_guards_near_thresholds(fast):
    'Elements placed within rounding of rh = 0.7, rh = 0+ and |G| = 0.39|A| (opposite sign).'
    rng = np.random.default_rng(8)
    n = 20000
    t = rng.uniform(260, 305, n).astype(np.float32)
    es = orc.svp(t.astype(np.float64))
    rh_target = np.where(rng.random(n) < 0.5, 0.7, 0.0) + rng.normal(0, 2e-7, n)
    rh_target[n // 2:] = np.abs(rng.normal(0, 5e-4, n - n // 2))
    vpd = (es * (1 - rh_target)).astype(np.float32)
    d = synthetic.host_drivers(1, n, 5, dtype=np.float32)
    d['temp_day'] = t[None]
    d['temp_night'] = (t - 6)[None]
    d['vpd_day'] = vpd[None]
    # soil heat flux switch: cold days with positive A so that G < 0 < 0.39 A, |G| ~ 0.39|A|
    tc = d['temp_day'].astype(np.float64) - 273.15
    g = 4.73 * tc - 20.87
    a_needed = np.abs(g) / 0.39 * (1 + rng.normal(0, 1e-7, g.shape))
    d['sw_rad_day'] = np.where(g < 0, (a_needed + 40) / (1 - d['sw_albedo'].astype(np.float64)),
                               d['sw_rad_day']).astype(np.float32)
    d['lw_net_day'] = np.where(g < 0, -40.0, d['lw_net_day']).astype(np.float32)
    d['temp_annual'] = np.full(n, 285.0, np.float32)
    drv = [d[k] for k in NAMES]
    d64 = [v.astype(np.float64) if isinstance(v, np.ndarray) else v for v in drv]
    with np.errstate(all='ignore'):
        od, on = orc.static_evapotranspiration(BPLUT7, *d64)
    r = hostmath.forward(BPLUT7, drv, 'static', 'global', dtype=np.float32, fast=fast)
    parity.assert_parity(r['tot_day'], od, np.float32, what='guards day')
    parity.assert_parity(r['tot_night'], on, np.float32, what='guards night')


def test_fast_f32_guard_rate_is_small():
    'The double-precision detour must stay rare on clean data (it costs ~20x a fast element).'
    T, N = 50, 4000
    table = _table()
    pft = synthetic.host_pft_map(N, 5)
    vec = [table[i][pft].reshape(1, N) for i in range(11)]
    for kind, limit in (('physical', 2e-4), ('stress', 2e-3)):
        d = synthetic.host_drivers(T, N, 99, kind=kind, dtype=np.float32)
        r = hostmath.forward(vec, [d[k] for k in NAMES], 'static', 'global', dtype=np.float32,
                             fast=True)
        rate = r['guarded'].mean()
        print('guard rate', kind, rate)
        assert rate < limit, (kind, rate)


@pytest.mark.parametrize('semantics', ['static', 'instance'])
def test_masked_driver_rule_is_sound(semantics):
    '''
    Where the rule says a half-day total is NaN, the float64 oracle's is NaN; it decides for
    (almost) every element whose only defect is a NaN in fPAR / LAI / albedo ..., i.e. masked
    pixels never reach the double-precision code; its gate predicate is the oracle's.
    '''
    T, N = 30, 2000
    rng = np.random.default_rng(8)
    d = synthetic.host_drivers(T, N, 55, dtype=np.float32)
    truth = dict((k, v.copy()) for k, v in d.items() if isinstance(v, np.ndarray))
    names2d = [k for k in NAMES if isinstance(d[k], np.ndarray) and d[k].ndim == 2]
    for k in names2d:                       # single NaNs in every per-step driver
        d[k][rng.random((T, N)) < 0.004] = np.nan
    d['pressure'][rng.random(N) < 0.01] = np.nan
    masked = rng.random(N) < 0.2            # whole masked pixels, as on a real tile
    for k in ('fpar', 'lai', 'sw_albedo'):
        d[k][:, masked] = np.nan
    fill_td = rng.random((T, N)) < 0.002
    d['temp_day'][fill_td] = -9999.0     # a negative temperature: r_corr is NaN (:292)
    drv = [d[k] for k in NAMES]
    d64 = [v.astype(np.float64) if isinstance(v, np.ndarray) else v for v in drv]
    table = _table()
    col = list(table[:, 7])
    h, g = hostmath.nan_halves(col, drv, semantics)
    with np.errstate(all='ignore'):
        if semantics == 'static':
            day, night = orc.static_evapotranspiration(col, *d64, gate='open')
        else:
            day, night = orc.InstanceModel(dict(zip(orc.PARAM_ORDER, col))).evapotranspiration(*d64)
    assert np.isnan(day[(h & 1) > 0]).all() and np.isnan(night[(h & 2) > 0]).all()
    # a negative (fill) temperature makes its half-day NaN like a NaN one does
    assert ((h[fill_td] & 1) > 0).all() and np.isnan(day[fill_td]).all()
    only_masked = np.broadcast_to(masked, (T, N)).copy()
    for k in names2d:
        if k not in ('fpar', 'lai', 'sw_albedo'):
            only_masked &= ~np.isnan(d[k])
    only_masked &= ~np.isnan(np.broadcast_to(d['pressure'], (T, N))) & (d['temp_day'] != -9999.0)
    assert (h[only_masked] == 3).all()
    if semantics == 'static':
        with np.errstate(all='ignore'):
            g_surf = (col[7] * orc.ramp_up(d64[8] - 273.15, col[0], col[1]) *
                      orc.ramp_down(d64[9], col[2], col[3])) / orc.r_correction(d64[11], d64[5])
        assert np.array_equal(g[h > 0], (g_surf > 0)[h > 0])


# ---- the float64 entry's fast path (forward_fast64_kernel) -----------------------------------
@pytest.mark.parametrize('semantics', ['static', 'instance'])
@pytest.mark.parametrize('kind,nan', [('physical', 0.0), ('stress', 0.0), ('physical', 0.01),
                                      ('stress', 0.01)])
def test_float64_fast_path_within_1e10_of_the_oracle(semantics, kind, nan):
    'The rearranged algebra in double: 1e-10 like the op-for-op code, and rarely guarded.'
    T, N = 60, 400
    d = synthetic.host_drivers(T, N, 4242, kind=kind, dtype=np.float64, nan_frac=nan,
                               fill_frac=nan)
    rng = np.random.default_rng(5)
    d['vpd_day'][rng.random((T, N)) < 0.03] *= -0.1        # supersaturated air
    drv = [d[k] for k in NAMES]
    table = _table()
    pft = synthetic.host_pft_map(N, 4243)
    vec = [table[i][pft].reshape(1, N) for i in range(11)]
    r = hostmath.forward(vec, drv, semantics, 'element', dtype=np.float64, fast=True)
    with np.errstate(all='ignore'):
        if semantics == 'static':
            day, night = orc.static_evapotranspiration(vec, *drv, gate='element', components=True)
            scale = 1.0
        else:
            day = [np.empty((T, N)) for _ in range(3)]
            night = [np.empty((T, N)) for _ in range(3)]
            for code in np.unique(pft):
                m = pft == code
                sub = [v[:, m] if isinstance(v, np.ndarray) and v.ndim == 2 else
                       (v[m] if isinstance(v, np.ndarray) else v) for v in drv]
                out = orc.InstanceModel(dict(zip(orc.PARAM_ORDER, table[:, code]))) \
                    .evapotranspiration(*sub, separate=True)
                for c in range(3):
                    day[c][:, m], night[c][:, m] = out[0][c], out[1][c]
            scale = parity.MASS_SCALE
    for c in range(3):
        parity.assert_parity(r['day'][c], day[c], np.float64, scale, what='day %d' % c)
        parity.assert_parity(r['night'][c], night[c], np.float64, scale, what='night %d' % c)
    tot = (day[0] + day[1] + day[2]) + (night[0] + night[1] + night[2])
    parity.assert_parity(r['tot'], tot, np.float64, scale, what='total')
    if nan == 0.0:
        assert r['guarded'].mean() < 2e-3, r['guarded'].mean()


def test_float64_fast_path_equals_the_op_for_op_code_on_a_plain_element():
    drv = [-30., -20., 300., 0., 0.15, 295., 285., 283., 280., 1200., 300., 95000., 0.5, 2.0]
    a = hostmath.forward(BPLUT7, drv, 'static', 'element', dtype=np.float64, fast=True)
    b = hostmath.forward(BPLUT7, drv, 'static', 'element', dtype=np.float64)
    for h in ('day', 'night'):
        for c in range(3):
            assert abs(a[h][c] - b[h][c]) <= 1e-12 * max(abs(b[h][c]), 1.0)
    assert not a['guarded'].any()


def test_float64_exp2_log2_against_libm():
    '''
    The float64 entry evaluates exp2 / log2 with its own restricted-range functions
    (csrc/mod16_fast.cuh: d_exp2, d_log2).  Over the arguments the fast algebra can produce
    they are 2-3 ulp functions (measured maxima against libm: 4.7e-16 for exp2, 5.8e-16 for
    log2; held below 1e-15 here) -- five orders inside the 1e-10 the entry promises -- and
    the special values the algebra relies on are exact: log2(0) = -inf (rh clamped to 0,
    mod16/__init__.py:280-281), exp2(-inf) = 0, exp2(0) = 1, log2(1) = 0.
    '''
    rng = np.random.default_rng(77)
    # exp2: the saturation-pressure exponent (|x| < 10) and the exponent of rh^(vpd / beta),
    # which is <= 0 and reaches far below the underflow threshold
    x = np.concatenate([rng.uniform(-10.0, 10.0, 200000), -rng.exponential(30.0, 200000),
                        rng.uniform(-1021.0, -900.0, 20000), np.arange(-1021.0, 11.0),
                        np.arange(-1021.0, 10.0) + 0.5, np.arange(-1020.0, 10.0) - 0.5])
    ex, _ = hostmath.exp2_log2(x)
    want = np.exp2(x)
    rel = np.abs(ex - want) / want
    assert rel.max() < 1e-15, rel.max()
    # log2: rh in [2e-5, 3e3] (0 < rh < 2e-5 is guarded, negative VPD reaches ~3e3), dense
    # around 1 and around sqrt(2) / 2, sqrt(2) where the reduction switches
    r = np.concatenate([np.exp(rng.uniform(np.log(2e-5), np.log(3e3), 300000)),
                        1.0 + rng.uniform(-1e-3, 1e-3, 50000), 1.0 + rng.uniform(-1e-9, 1e-9, 5000),
                        np.sqrt(2.0) * (1.0 + rng.uniform(-1e-6, 1e-6, 5000)),
                        np.sqrt(0.5) * (1.0 + rng.uniform(-1e-6, 1e-6, 5000)),
                        2.0 ** np.arange(-15, 12), np.nextafter(1.0, 0.0) * np.ones(1),
                        np.nextafter(1.0, 2.0) * np.ones(1)])
    _, lg = hostmath.exp2_log2(r)
    want = np.log2(r)
    # relative where log2 is away from 0, absolute (in units of |r - 1|) next to r = 1
    err = np.abs(lg - want) / np.maximum(np.abs(want), 1e-300)
    assert err[want != 0].max() < 1e-15, err[want != 0].max()
    # special values
    ex, lg = hostmath.exp2_log2(np.array([0.0, -np.inf, -1021.0, -1021.5, -5000.0, 1.0]))
    assert ex[0] == 1.0 and ex[1] == 0.0 and ex[2] == 2.0 ** -1021 and ex[3] == 0.0 and ex[4] == 0.0
    assert lg[0] == -np.inf and lg[5] == 0.0
    _, lg = hostmath.exp2_log2(2.0 ** np.arange(-14, 12))
    assert np.array_equal(lg, np.arange(-14.0, 12.0))


@pytest.mark.parametrize('semantics', ['static', 'instance'])
def test_float64_fast_path_at_the_discontinuities(semantics):
    '''
    Elements placed on both sides of every switch of the model, 1e-13 ... 1e-7 away
    (tests/switch_cases.py lists them).  The float64 fast path -- its guards, and since round
    2 its own exp2 / log2 -- must stay within 1e-10 of the oracle on all of them.
    '''
    import switch_cases
    drv, groups, grp, off = switch_cases.build(semantics)
    r = hostmath.forward(BPLUT7, drv, semantics, 'element', dtype=np.float64, fast=True)
    day, night, scale = switch_cases.reference(semantics, drv)
    assert np.isfinite(day[0] + day[1] + day[2]).mean() > 0.95     # a test of values, not NaNs
    for c in range(3):
        parity.assert_parity(r['day'][c], day[c], np.float64, scale, what='day %d' % c)
        parity.assert_parity(r['night'][c], night[c], np.float64, scale, what='night %d' % c)
    tot = (day[0] + day[1] + day[2]) + (night[0] + night[1] + night[2])
    parity.assert_parity(r['tot'], tot, np.float64, scale, what='total')
    # the guards are narrow: close to a guarded switch most elements take the op-for-op
    # detour, 1e-8 away none does; exact comparisons and continuous ramps need none
    g = r['guarded'].reshape(-1)
    near, far = np.abs(off) <= 3e-10, np.abs(off) >= 1e-8
    for name in ('rh07', 'g_day', 'g_night', 'a_zero'):
        m = grp == groups.index(name)
        assert g[m & near].mean() > 0.4 and g[m & far].mean() < 0.05, name
    for name in ('rh1', 'soil_n', 'dt5', 'tann_hi', 'tann_lo', 'tmin_hi', 'vpd_lo'):
        assert g[grp == groups.index(name)].mean() == 0.0, name
    # the zero of a ramp (Tmin at tmin_close, VPD at vpd_close): under the per-element gate of
    # the static path g_surf = 0+ switches the cuticular transpiration on, so it is guarded
    # (1e-12 of the ramp's terms in double); the instance path has no gate and needs none
    for name in ('tmin_lo', 'vpd_hi'):
        m = grp == groups.index(name)
        if semantics == 'static':
            # (the VPD ramp's guard ends at |off| = 1e-12 exactly: asserted one decade inside)
            assert g[m & (np.abs(off) <= 1e-13)].all() and not g[m & (np.abs(off) >= 1e-9)].any(), name
        elif name == 'tmin_lo':
            assert not g[m].any(), name


@pytest.mark.parametrize('fast', [True, 'pair'])
@pytest.mark.parametrize('semantics', ['static', 'instance'])
def test_float32_fast_path_at_the_discontinuities(semantics, fast):
    '''
    The same switches for the float32 kernels' math (scalar and packed-pair instantiations),
    offsets 1e-8 ... 1e-3 rounded to float32: 1e-5 relative / 1e-3 W m-2 against the float64
    oracle on the same float32 inputs, components and totals.
    '''
    import switch_cases
    drv, groups, grp, off = switch_cases.build(semantics, np.float32)
    r = hostmath.forward(BPLUT7, drv, semantics, 'element', dtype=np.float32, fast=fast)
    day, night, scale = switch_cases.reference(semantics, drv)
    for c in range(3):
        parity.assert_parity(r['day'][c], day[c], np.float32, scale, what='day %d' % c)
        parity.assert_parity(r['night'][c], night[c], np.float32, scale, what='night %d' % c)
    td, tn = day[0] + day[1] + day[2], night[0] + night[1] + night[2]
    parity.assert_parity(r['tot_day'], td, np.float32, scale, what='day total')
    parity.assert_parity(r['tot_night'], tn, np.float32, scale, what='night total')
    g = r['guarded'].reshape(-1)
    print('guarded fraction at the discontinuities, float32 (%s, %s): %.3f' % (
        semantics, fast, g.mean()))
