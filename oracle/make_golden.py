'''
TEST INFRASTRUCTURE ONLY -- freezes golden vectors from the UNMODIFIED reference.

Run in the build container (where /root/reference exists):

    python oracle/make_golden.py

It imports the reference package from /root/reference (with oracle/refshim/mod17 standing
in for the single external function ``mod17.linear_constraint``), evaluates it on

  1. the reference's own unit-test fixture and all 15 known answers (tests/tests.py),
  2. the inline 3-element arrays of tests/verification/verify.py:48-67 (PFT 7),
  3. the notebook scalar (notebooks/MOD16_Forward_Run_example.ipynb:222),
  4. the static-vs-instance divergence cases and the NaN-mask cases of SURVEY.md 8(a),
  5. seeded random (T, N) drivers from both synthetic generators, with NaN and -9999
     fill injection, per-pixel parameters gathered by PFT, in float64 and float32,

asserts that oracle/mod16_oracle.py reproduces every one BIT-FOR-BIT (float64), and writes
inputs + reference outputs to tests/golden/*.npz.  Those files travel to the GPU box; the
reference does not.
'''
import os
import sys
import json
import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, os.path.join(HERE, 'refshim'))
sys.path.insert(0, '/root/reference')
sys.path.insert(0, ROOT)

import mod16 as ref                                   # noqa: E402  (the reference)
from mod16.utils import restore_bplut as ref_restore  # noqa: E402
from mod16.models import MOD16Collection61 as RefC61  # noqa: E402
from oracle import mod16_oracle as orc                # noqa: E402
from mod16_b200 import synthetic, utils               # noqa: E402

GOLDEN = os.path.join(ROOT, 'tests', 'golden')
NAMES = list(synthetic.DRIVER_NAMES)
ORDER = list(ref.MOD16.required_parameters)
assert tuple(ORDER) == orc.PARAM_ORDER


def same(a, b):
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    return a.shape == b.shape and np.array_equal(a, b, equal_nan=True)


def check(tag, a, b):
    if not same(a, b):
        raise SystemExit('oracle != reference for %s:\n%r\n%r' % (tag, a, b))


def bplut_c51():
    d = ref_restore(os.path.join(
        os.path.dirname(ref.__file__),
        'data/MOD16_BPLUT_C5.1_05deg_MCD43B_Albedo_MERRA_GMAO.csv'))
    mine = utils.restore_bplut(utils.BPLUT_C51)
    for k in d:
        check('bplut ' + k, d[k], mine[k])
    return d


def params_for(bplut, pft, beta=250.0):
    p = dict((k, bplut[k][pft]) for k in ORDER)
    if np.isnan(p['beta']):
        p['beta'] = beta
    return p


def both_paths(p, drv, tag):
    'Evaluate the reference and the oracle on one driver set; return reference outputs.'
    vec = [p[k] for k in ORDER]
    with np.errstate(all='ignore'):
        r_day, r_night = ref.MOD16._evapotranspiration(vec, *drv)
        r_et = ref.MOD16._et(vec, *drv)
        o_day, o_night = orc.static_evapotranspiration(vec, *drv)
        o_et = orc.static_et(vec, *drv)
        m = ref.MOD16(p)
        i_day, i_night = m.evapotranspiration(*drv)
        i_sep = m.evapotranspiration(*drv, separate=True)
        om = orc.InstanceModel(p)
        j_day, j_night = om.evapotranspiration(*drv)
        j_sep = om.evapotranspiration(*drv, separate=True)
    check(tag + ' static day', r_day, o_day)
    check(tag + ' static night', r_night, o_night)
    check(tag + ' static et', r_et, o_et)
    check(tag + ' inst day', i_day, j_day)
    check(tag + ' inst night', i_night, j_night)
    for h in (0, 1):
        for c in (0, 1, 2):
            check(tag + ' inst sep %d %d' % (h, c), i_sep[h][c], j_sep[h][c])
    bc = lambda x: np.asarray(x, np.float64)
    return dict(
        static_day=bc(r_day), static_night=bc(r_night), static_et=bc(r_et),
        inst_day=bc(i_day), inst_night=bc(i_night),
        inst_day_canopy=bc(i_sep[0][0]), inst_day_soil=bc(i_sep[0][1]),
        inst_day_trans=bc(i_sep[0][2]), inst_night_canopy=bc(i_sep[1][0]),
        inst_night_soil=bc(i_sep[1][1]), inst_night_trans=bc(i_sep[1][2]))


def unit_tests_fixture():
    'tests/tests.py: the setUp fixture and all 15 known answers, recomputed here.'
    P = dict(gl_sh=0.01, gl_wv=0.01, g_cuticular=1e-5, tmin_close=-8, tmin_open=8,
             vpd_open=650, vpd_close=3000, rbl_min=60, rbl_max=90, csl=2.4e-3,
             beta=250)
    pressure = 100e3
    temp_k = 273.15 + 30
    vpd = 1000
    lai = 1.5
    fpar = 0.5
    rad_canopy = 5000
    rad_soil = 5000
    r_corr = (101300 / pressure) * (temp_k / 293.15)**1.75
    drv = [-50, -30, 150, 0, 0.3, 293, 290, 285, 285, 1000, 500, 100e3, 0.5, 1.5]
    _pressure = np.arange(98e3, 103e3, 1e3)
    _temp_k = 273.15 + np.array([0, 10, 20, 30, 40])
    _vpd = np.arange(0, 5000, 1000)
    _lai = np.arange(0.5, 3, 0.5)
    _fpar = np.array([0.1, 0.3, 0.5, 0.7, 0.9])
    _rad_canopy = np.arange(3e3, 8e3, 1e3)
    tmin = 285
    m = ref.MOD16(P)
    o = orc.InstanceModel(P)
    out = {}
    cases = [
        ('evaporation_soil', (pressure, temp_k, vpd, fpar, rad_soil, r_corr)),
        ('evaporation_soil_by_fpar', (pressure, temp_k, vpd, _fpar, rad_soil, r_corr)),
        ('transpiration_day', (pressure, temp_k, vpd, lai, fpar, rad_canopy, tmin, r_corr)),
        ('transpiration_night', (pressure, temp_k, vpd, lai, fpar, rad_canopy, tmin, r_corr)),
        ('wet_canopy', (pressure, temp_k, vpd, lai, fpar, rad_canopy)),
        ('wet_canopy_by_pressure', (_pressure, temp_k, vpd, lai, fpar, rad_canopy)),
        ('wet_canopy_by_temperature', (pressure, _temp_k, vpd, lai, fpar, rad_canopy)),
        ('wet_canopy_by_vpd', (pressure, temp_k, _vpd, lai, fpar, rad_canopy)),
        ('wet_canopy_by_lai', (pressure, temp_k, vpd, _lai, fpar, rad_canopy)),
        ('wet_canopy_by_fpar', (pressure, temp_k, vpd, lai, _fpar, rad_canopy)),
        ('wet_canopy_by_rad', (pressure, temp_k, vpd, lai, fpar, _rad_canopy)),
    ]
    for name, args in cases:
        if name.startswith('evaporation_soil'):
            r = m.evaporation_soil(*args)
            q = o.evaporation_soil(*args)
        elif name == 'transpiration_day':
            r = m.transpiration(*args, daytime=True)
            q = o.transpiration(*args, daytime=True)
        elif name == 'transpiration_night':
            r = m.transpiration(*args, daytime=False)
            q = o.transpiration(*args, daytime=False)
        else:
            r = m.evaporation_wet_canopy(*args)
            q = o.evaporation_wet_canopy(*args)
        check('unit ' + name, r, q)
        out[name] = np.asarray(r, np.float64)
    # helpers
    pp = np.array((100e3, 80e3, 100e3, 80e3, 81.8e3))
    tt = 273.15 + np.array((10, 10, 25, 25, 25))
    r = ref.psychrometric_constant(pp, tt)
    check('psy', r, orc.psychrometric_constant(pp, tt))
    out['psychrometric_in_p'] = pp
    out['psychrometric_in_t'] = tt
    out['psychrometric'] = r
    ts = 273.15 + np.array((10., 20., 25.))
    r = ref.svp_slope(ts)
    check('slope', r, orc.svp_slope(ts))
    out['svp_slope_in'] = ts
    out['svp_slope'] = r
    out['svp'] = ref.svp(ts)
    check('svp', out['svp'], orc.svp(ts))
    out['lhv'] = ref.latent_heat_vaporization(ts)
    check('lhv', out['lhv'], orc.lhv(ts))
    rr = np.array([0.1, 0.5, 0.9])
    out['air_density'] = ref.MOD16.air_density(ts, 95e3, rr)
    check('rho', out['air_density'], orc.air_density(ts, 95e3, rr))
    vv = np.array([-50., 800., 5000.])
    out['rhumidity_in_vpd'] = vv
    out['rhumidity'] = ref.MOD16.rhumidity(ts, vv)
    check('rh', out['rhumidity'], orc.rhumidity_clamped(ts, vv))
    qv = np.array([0.002, 0.008, 0.015])
    out['vpd_in_qv'] = qv
    out['vpd'] = ref.MOD16.vpd(qv, 95e3, ts)
    check('vpd', out['vpd'], orc.vpd_from_qv(qv, 95e3, ts))
    el = np.array([0., 1500., 4000.])
    out['air_pressure_in'] = el
    out['air_pressure'] = ref.MOD16.air_pressure(el)
    check('pa', out['air_pressure'], orc.air_pressure(el))
    # radiation_soil / soil_heat_flux / surface_conductance
    rs = m.radiation_soil(-50, -30, 150, 0, 0.3, 293, 285, 285, _fpar)
    qs = o.radiation_soil(-50, -30, 150, 0, 0.3, 293, 285, 285, _fpar)
    check('rad_soil d', rs[0], qs[0])
    check('rad_soil n', rs[1], qs[1])
    out['radiation_soil_day'] = np.asarray(rs[0], np.float64)
    out['radiation_soil_night'] = np.asarray(rs[1], np.float64)
    tm = np.array([260., 270., 280., 290.])
    vd = np.array([300., 900., 2000., 3500.])
    out['surface_conductance_in_tmin'] = tm
    out['surface_conductance_in_vpd'] = vd
    out['surface_conductance'] = m.surface_conductance(tm, vd)
    check('gs', out['surface_conductance'], o.surface_conductance(tm, vd))
    # the whole-model consistency test (tests.py:64-90)
    res = both_paths(P, drv, 'unit et')
    out.update(dict(('et_' + k, v) for k, v in res.items()))
    assert round(float(res['static_et']), 2) == 40.94
    lam_d = ref.latent_heat_vaporization(293)
    lam_n = ref.latent_heat_vaporization(290)
    assert round(float(res['inst_day'] * lam_d + res['inst_night'] * lam_n), 1) == 41.0
    out['params'] = np.array([P[k] for k in ORDER], np.float64)
    out['drivers'] = np.array(drv, np.float64)
    np.savez(os.path.join(GOLDEN, 'unit_tests.npz'), **out)
    return len(out)


def verify_fixture(bplut):
    'tests/verification/verify.py:48-85 -- inline arrays, PFT 7, beta = 250.'
    p = params_for(bplut, 7)
    Tday = np.array((286.20189, 292.52667, 298.3286))
    pressure = np.array((92753.47, 92753.47, 92753.47))
    VPD_day = np.array((710.9, 1249.4, 1979.))
    Tmin = np.array((278.92, 284.43, 289.88))
    LAI = np.array((0.3, 0.6, 1.0))
    Fwet_day = np.array((0, 0.4, 0.8))
    out = {}
    m = ref.MOD16(p)
    o = orc.InstanceModel(p)
    a = (-117, -65, 419, 0, 0.116, Tday, Tday, 289.74402, Tmin, VPD_day, VPD_day,
         pressure, 0.35839, LAI)
    r = m.evapotranspiration(*a, f_wet=Fwet_day, separate=True)
    q = o.evapotranspiration(*a, f_wet=Fwet_day, separate=True)
    for h in (0, 1):
        for c, cn in enumerate(('canopy', 'soil', 'trans')):
            check('verify inst', r[h][c], q[h][c])
            out['inst_%s_%s' % (('day', 'night')[h], cn)] = np.asarray(r[h][c], np.float64)
    b = (-117, -65, 419, 0, 0.116, Tday, Tday, 289.74402, Tmin, VPD_day, 512,
         pressure, 0.35839, LAI)
    vec = [p[k] for k in ORDER]
    r = ref.MOD16._evapotranspiration(vec, *b, f_wet=Fwet_day)
    q = orc.static_evapotranspiration(vec, *b, f_wet=Fwet_day)
    check('verify static d', r[0], q[0])
    check('verify static n', r[1], q[1])
    out['static_day'] = r[0]
    out['static_night'] = r[1]
    out['static_et'] = ref.MOD16._et(vec, *b)
    out['params'] = np.array(vec, np.float64)
    out['temp'] = Tday
    out['pressure'] = pressure
    out['vpd_day'] = VPD_day
    out['tmin'] = Tmin
    out['lai'] = LAI
    out['scalars'] = np.array([-117, -65, 419, 0, 0.116, 289.74402, 0.35839, 512.])
    # intermediates quoted in SURVEY.md 8(c)
    out['lhv'] = ref.latent_heat_vaporization(Tday)
    out['svp'] = ref.svp(Tday)
    # notebook scalar (…Forward_Run_example.ipynb:222)
    nb = RefC61(12).evaporation_soil(101e3, 293.15, 1000, 0.5, 100)
    assert float(nb) == 5.083295223395212e-06, nb
    p12 = params_for(bplut, 12)
    check('notebook', nb, orc.InstanceModel(p12).evaporation_soil(101e3, 293.15, 1000, 0.5, 100))
    out['notebook_scalar'] = np.float64(nb)
    out['notebook_params'] = np.array([p12[k] for k in ORDER], np.float64)
    np.savez(os.path.join(GOLDEN, 'verify.npz'), **out)
    return len(out)


def table_fixture(bplut):
    'SURVEY.md 8(a): divergence cases and NaN-mask cases around one base point.'
    p = params_for(bplut, 7)
    base = dict(zip(NAMES, [-50., -30., 150., 0., 0.3, 293., 285., 285., 285.,
                            1000., 500., 1e5, 0.5, 1.5]))
    cases = []

    def add(tag, **kw):
        d = dict(base)
        d.update(kw)
        cases.append((tag, d))

    add('base')
    add('tann_eq_close', temp_annual=273.15 + p['tmin_close'])
    add('tann_eq_open', temp_annual=273.15 + p['tmin_open'])
    add('tann_cold', temp_annual=270.0)
    add('tann_hot', temp_annual=299.0)
    add('vpd_day_neg', vpd_day=-200.)
    add('tmin_closed', tmin=260.)
    add('lai_zero', lai=0.)
    add('fpar_one', fpar=1.)
    add('vpd_zero', vpd_day=0., vpd_night=0.)
    add('sw0_lwneg', sw_rad_day=0., lw_net_day=-100.)
    add('wet_day', vpd_day=300.)
    add('wet_both', vpd_day=300., vpd_night=100.)
    add('dT_small', temp_night=290.)
    add('g_limited', temp_day=305., temp_night=295., lw_net_day=-95., sw_rad_day=20.)
    add('cold_g_negative', temp_day=276., temp_night=268., temp_annual=282.)
    add('beta_zero_like', vpd_day=0.)
    for name in NAMES:
        add('nan_' + name, **{name: np.nan})
    for name in ('temp_day', 'temp_night', 'vpd_day', 'lai', 'fpar', 'tmin', 'pressure'):
        add('fill_' + name, **{name: -9999.})
    out = {'params': np.array([p[k] for k in ORDER], np.float64),
           'case_names': np.array([c[0] for c in cases])}
    rows = []
    res_all = {}
    for tag, d in cases:
        # np.float64 scalars: python floats would turn (negative)**1.75 complex
        drv = [np.float64(d[k]) for k in NAMES]
        rows.append(drv)
        res = both_paths(p, drv, 'table ' + tag)
        for k, v in res.items():
            res_all.setdefault(k, []).append(float(v))
    out['drivers'] = np.array(rows, np.float64)
    for k, v in res_all.items():
        out[k] = np.array(v, np.float64)
    # batch-level gate: the whole batch closed vs one open element (static only)
    vec = [p[k] for k in ORDER]
    tm = np.array([260., 262.])
    drv = [np.float64(base[k]) for k in NAMES]
    drv[NAMES.index('tmin')] = tm
    out['gate_closed_tmin'] = tm
    out['gate_closed_et'] = ref.MOD16._et(vec, *drv)
    check('gate closed', out['gate_closed_et'], orc.static_et(vec, *drv))
    tm2 = np.array([260., 285.])
    drv[NAMES.index('tmin')] = tm2
    out['gate_open_tmin'] = tm2
    out['gate_open_et'] = ref.MOD16._et(vec, *drv)
    check('gate open', out['gate_open_et'], orc.static_et(vec, *drv))
    tm3 = np.array([np.nan, np.nan])
    drv[NAMES.index('tmin')] = tm3
    out['gate_nan_et'] = ref.MOD16._et(vec, *drv)
    check('gate nan', out['gate_nan_et'], orc.static_et(vec, *drv))
    np.savez(os.path.join(GOLDEN, 'tables.npz'), **out)
    return len(cases)


def random_fixture(bplut, tag, T, N, seed, kind, dtype, nan_frac, fill_frac):
    '''
    Seeded (T, N) drivers; static path with per-pixel parameters gathered by PFT, instance
    path for one PFT.  float32 cases store the float32 inputs and the reference evaluated
    in float64 on those same values (the fp32 kernel's target).
    '''
    d = synthetic.host_drivers(T, N, seed, kind=kind, dtype=dtype,
                               nan_frac=nan_frac, fill_frac=fill_frac)
    pft = synthetic.host_pft_map(N, seed + 1)
    table = utils.bplut_table(bplut, ORDER, default_beta=250.0)
    up = lambda v: v.astype(np.float64) if isinstance(v, np.ndarray) else v
    drv = [up(d[k]) for k in NAMES]
    vec = [table[i][pft].reshape(1, N) for i in range(11)]
    with np.errstate(all='ignore'):
        r_day, r_night = ref.MOD16._evapotranspiration(vec, *drv)
        r_et = ref.MOD16._et(vec, *drv)
        o_day, o_night = orc.static_evapotranspiration(vec, *drv)
        check(tag + ' static day', r_day, o_day)
        check(tag + ' static night', r_night, o_night)
        check(tag + ' static et', r_et, orc.static_et(vec, *drv))
        p = params_for(bplut, 4)
        i_sep = ref.MOD16(p).evapotranspiration(*drv, separate=True)
        j_sep = orc.InstanceModel(p).evapotranspiration(*drv, separate=True)
    out = dict(pft=pft, table=table, inst_params=np.array([p[k] for k in ORDER]),
               static_day=r_day, static_night=r_night, static_et=r_et)
    for h in (0, 1):
        for c, cn in enumerate(('canopy', 'soil', 'trans')):
            check(tag + ' inst', i_sep[h][c], j_sep[h][c])
            out['inst_%s_%s' % (('day', 'night')[h], cn)] = np.asarray(i_sep[h][c])
    for k in NAMES:
        if isinstance(d[k], np.ndarray):
            out['drv_' + k] = d[k]
    np.savez_compressed(os.path.join(GOLDEN, tag + '.npz'), **out)
    return T * N


def components_fixture(bplut):
    '''
    Every component method / helper of the public API on seeded random inputs, with and
    without the optional arguments, NaN injected -- evaluated by the REFERENCE; the oracle
    must agree bit for bit.
    '''
    rng = np.random.default_rng(16110)
    n = 64
    P = params_for(bplut, 7)
    m, o = ref.MOD16(P), orc.InstanceModel(P)
    u = lambda lo, hi: rng.uniform(lo, hi, n)
    x = dict(pressure=u(7e4, 1.0134e5), temp_k=u(255, 305), vpd=u(0, 3400), fpar=u(0.02, 0.89),
             lai=u(0, 5.3), rad_soil=u(-50, 300), rad_canopy=u(-50, 300), tmin=u(252, 298),
             r_corr=u(0.8, 1.4), lhv=u(2.4e6, 2.55e6), rh=u(0, 1), fw=u(0, 1),
             lw_d=u(-100, 0), lw_n=u(-50, 0), sw_d=u(0, 360), sw_n=u(0, 1e-5), alb=u(0.1, 0.22),
             t_d=u(255, 305), t_ann=u(265, 300), qv=u(0.001, 0.02), elev=u(-100, 5000),
             sw=u(0, 5000))
    x['t_n'] = x['t_d'] - u(0, 15)
    x['lai'][:4] = 0.0
    x['fw'][4:8] = 0.0
    x['vpd'][8:10] = 0.0
    for k in ('temp_k', 'vpd', 'lai', 'rad_canopy', 'tmin', 'pressure', 'fpar'):
        x[k][rng.integers(10, n)] = np.nan
    out = dict(('in_' + k, v) for k, v in x.items())
    out['params'] = np.array([P[k] for k in ORDER], np.float64)

    def both(tag, rfn, ofn, *args, **kw):
        with np.errstate(all='ignore'):
            r, q = rfn(*args, **kw), ofn(*args, **kw)
        if isinstance(r, (tuple, list)):
            for i, (a, b) in enumerate(zip(r, q)):
                check('component %s[%d]' % (tag, i), a, b)
                out['%s_%d' % (tag, i)] = np.asarray(a, np.float64)
        else:
            check('component ' + tag, r, q)
            out[tag] = np.asarray(r, np.float64)

    a5 = (x['pressure'], x['temp_k'], x['vpd'], x['fpar'], x['rad_soil'])
    both('evaporation_soil', m.evaporation_soil, o.evaporation_soil, *a5)
    both('evaporation_soil_opt', m.evaporation_soil, o.evaporation_soil, *a5, x['r_corr'],
         x['lhv'], x['rh'], x['fw'])
    bp = (P['vpd_open'], P['vpd_close'], P['rbl_min'], P['rbl_max'])
    both('potential_soil_evaporation', ref.MOD16.potential_soil_evaporation,
         orc.InstanceModel.potential_soil_evaporation, *a5, None, None, None, None, *bp)
    both('potential_soil_evaporation_opt', ref.MOD16.potential_soil_evaporation,
         orc.InstanceModel.potential_soil_evaporation, *a5, x['r_corr'], x['lhv'], x['rh'],
         x['fw'], *bp)
    a6 = (x['pressure'], x['temp_k'], x['vpd'], x['lai'], x['fpar'], x['rad_canopy'])
    both('evaporation_wet_canopy', m.evaporation_wet_canopy, o.evaporation_wet_canopy, *a6)
    both('evaporation_wet_canopy_opt', m.evaporation_wet_canopy, o.evaporation_wet_canopy, *a6,
         x['lhv'], x['rh'], x['fw'])
    for day in (True, False):
        tag = 'transpiration_%s' % ('day' if day else 'night')
        both(tag, m.transpiration, o.transpiration, *a6, x['tmin'], daytime=day)
        both(tag + '_opt', m.transpiration, o.transpiration, *a6, x['tmin'], x['r_corr'],
             x['lhv'], x['rh'], x['fw'], daytime=day)
    both('surface_conductance', m.surface_conductance, o.surface_conductance, x['tmin'], x['vpd'])
    a9 = (x['lw_d'], x['lw_n'], x['sw_d'], x['sw_n'], x['alb'], x['t_d'], x['t_n'], x['t_ann'],
          x['fpar'])
    both('radiation_soil', m.radiation_soil, o.radiation_soil, *a9)
    rad_d = x['sw_d'] * (1 - x['alb']) + x['lw_d']
    both('soil_heat_flux', m.soil_heat_flux, o.soil_heat_flux, rad_d, x['lw_n'], x['t_d'],
         x['t_n'], x['t_ann'])
    a7 = (x['lw_d'], x['sw_d'], x['alb'], x['pressure'], x['temp_k'], x['vpd'], x['fpar'])
    both('potential_transpiration', ref.MOD16.potential_transpiration,
         orc.InstanceModel.potential_transpiration, *a7)
    both('potential_transpiration_opt', ref.MOD16.potential_transpiration,
         orc.InstanceModel.potential_transpiration, *a7, x['rh'], x['fw'], 1.1)
    both('svp', ref.svp, orc.svp, x['temp_k'])
    both('svp_slope', ref.svp_slope, orc.svp_slope, x['temp_k'])
    both('svp_slope_opt', ref.svp_slope, orc.svp_slope, x['temp_k'], x['lhv'] * 1e-3)
    both('lhv', ref.latent_heat_vaporization, orc.lhv, x['temp_k'])
    both('psychrometric', ref.psychrometric_constant, orc.psychrometric_constant, x['pressure'],
         x['temp_k'])
    both('air_density', ref.MOD16.air_density, orc.air_density, x['temp_k'], x['pressure'],
         x['rh'])
    both('rhumidity', ref.MOD16.rhumidity, orc.rhumidity_clamped, x['temp_k'], x['vpd'])
    both('vpd', ref.MOD16.vpd, orc.vpd_from_qv, x['qv'], x['pressure'], x['temp_k'])
    both('air_pressure', ref.MOD16.air_pressure, orc.air_pressure, x['elev'])
    both('radiation_net', ref.radiation_net, orc.radiation_net, x['sw'], x['alb'], x['temp_k'])
    # the reference's own known answers for radiation_net (tests/tests.py:217-226)
    swrad = np.array((500, 5000, 500, 5000, 500, 5000, 500, 5000.))
    albedo = np.array((0.4, 0.4, 0.8, 0.8, 0.4, 0.4, 0.8, 0.8))
    tk = 273.15 + np.array((10, 10, 10, 10, 25, 25, 25, 25.))
    r = ref.radiation_net(swrad, albedo, tk)
    assert list(r.round(1)) == [223.3, 2923.3, 23.3, 923.3, 241.8, 2941.8, 41.8, 941.8]
    check('radiation_net known', r, orc.radiation_net(swrad, albedo, tk))
    out['radiation_net_known_in'] = np.stack([swrad, albedo, tk])
    out['radiation_net_known'] = r
    np.savez_compressed(os.path.join(GOLDEN, 'components.npz'), **out)
    return len(out)


def constraint_fixture():
    '''
    `constrain_by_map` is a function nested in CalibrationAPI.tune (mod16/calibration.py:775-
    796); calibration.py itself cannot be imported here (pymc, h5py ... are absent).  Its
    source text is cut out of the unmodified file with `ast` and executed on its own -- it
    only needs numpy -- so the golden answers come from the reference's own statements.
    '''
    import ast
    import textwrap
    path = '/root/reference/mod16/calibration.py'
    src = open(path).read()
    node = [n for n in ast.walk(ast.parse(src)) if isinstance(n, ast.FunctionDef)
            and n.name == 'constrain_by_map']
    assert len(node) == 1
    lines = src.split('\n')[node[0].lineno - 1:node[0].end_lineno]
    scope = {'np': np}
    exec(compile(textwrap.dedent('\n'.join(lines)), path, 'exec'), scope)
    ref_fn = scope['constrain_by_map']
    rng = np.random.default_rng(16201)
    out = {'lines': np.array([node[0].lineno, node[0].end_lineno])}
    n = 0
    for tag, T, N, Y in (('a', 730, 17, 2), ('b', 1461, 5, 4), ('c', 40, 3, 3)):
        years = np.sort(rng.integers(0, Y, T)) + 2001
        le = rng.normal(60, 50, (T, N))                # some negative fluxes (:786)
        temp = rng.uniform(255, 305, (T, N))
        lhv = ref.latent_heat_vaporization(temp)
        precip = rng.uniform(200, 1500, (np.unique(years).size, N))
        if tag == 'c':
            precip[:] = 1e9                            # never exceeded: the objective is -0
        want = ref_fn(le.copy(), years, lhv, precip)
        check('constrain_by_map ' + tag, want, orc.constrain_by_map(le.copy(), years, lhv, precip))
        for k, v in (('le', le), ('years', years), ('lhv', lhv), ('precip', precip),
                     ('neg_nrmsd', np.float64(want)),
                     ('annual', orc.annual_mass_sums(le.copy(), years, lhv))):
            out['%s_%s' % (tag, k)] = v
        n += 1
    np.savez_compressed(os.path.join(GOLDEN, 'constraint.npz'), **out)
    return n


def main():
    os.makedirs(GOLDEN, exist_ok=True)
    bplut = bplut_c51()
    report = {}
    report['unit_tests'] = unit_tests_fixture()
    report['verify'] = verify_fixture(bplut)
    report['tables'] = table_fixture(bplut)
    report['components'] = components_fixture(bplut)
    report['constraint'] = constraint_fixture()
    report['rand_phys_f64'] = random_fixture(
        bplut, 'rand_phys_f64', 24, 96, 16101, 'physical', np.float64, 0.0, 0.0)
    report['rand_phys_fill_f64'] = random_fixture(
        bplut, 'rand_phys_fill_f64', 24, 96, 16102, 'physical', np.float64, 0.01, 0.01)
    report['rand_stress_fill_f64'] = random_fixture(
        bplut, 'rand_stress_fill_f64', 24, 96, 16103, 'stress', np.float64, 0.01, 0.01)
    report['rand_phys_f32'] = random_fixture(
        bplut, 'rand_phys_f32', 24, 96, 16104, 'physical', np.float32, 0.0, 0.0)
    report['rand_stress_fill_f32'] = random_fixture(
        bplut, 'rand_stress_fill_f32', 24, 96, 16105, 'stress', np.float32, 0.01, 0.01)
    report['numpy'] = np.__version__
    with open(os.path.join(GOLDEN, 'MANIFEST.json'), 'w') as f:
        json.dump(report, f, indent=1)
    print('golden vectors written:', report)


if __name__ == '__main__':
    main()
