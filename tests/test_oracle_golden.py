'''
Pins the oracle (oracle/mod16_oracle.py) to the reference: replays the golden vectors that
oracle/make_golden.py froze from the UNMODIFIED reference and requires bit-identical
float64 output, and checks the reference's own published known answers (the rounded values
asserted in /root/reference/tests/tests.py).  Runs without a GPU and without the reference.
'''
import glob
import os

import numpy as np
import pytest

from mod16_b200 import synthetic
from oracle import mod16_oracle as orc

NAMES = list(synthetic.DRIVER_NAMES)


def same(a, b):
    a = np.asarray(a, np.float64)
    b = np.asarray(b, np.float64)
    return a.shape == b.shape and np.array_equal(a, b, equal_nan=True)


def load(golden_dir, name):
    return np.load(os.path.join(golden_dir, name + '.npz'))


def test_manifest_lists_all_fixtures(golden_dir):
    import json
    man = json.load(open(os.path.join(golden_dir, 'MANIFEST.json')))
    for k in man:
        if k != 'numpy':
            assert os.path.exists(os.path.join(golden_dir, k + '.npz')), k


def test_reference_unit_test_known_answers(golden_dir):
    'The rounded answers are the ones /root/reference/tests/tests.py asserts.'
    z = load(golden_dir, 'unit_tests')
    P = dict(zip(orc.PARAM_ORDER, z['params']))
    m = orc.InstanceModel(P)
    pressure, temp_k, vpd, lai, fpar = 100e3, 273.15 + 30, 1000, 1.5, 0.5
    rad = 5000
    r_corr = (101300 / pressure) * (temp_k / 293.15)**1.75
    h = 60 * 60
    assert (h * m.evaporation_soil(pressure, temp_k, vpd, fpar, rad, r_corr)).round(3) == 3.102
    f = np.array([0.1, 0.3, 0.5, 0.7, 0.9])
    assert np.array_equal(
        (h * m.evaporation_soil(pressure, temp_k, vpd, f, rad, r_corr)).round(3),
        [3.128, 3.115, 3.102, 3.089, 3.076])
    assert (h * m.transpiration(pressure, temp_k, vpd, lai, fpar, rad, 285, r_corr,
                                daytime=True)).round(3) == 1.248
    assert (h * m.transpiration(pressure, temp_k, vpd, lai, fpar, rad, 285, r_corr,
                                daytime=False)).round(3) == 0.011
    assert m.evaporation_wet_canopy(pressure, temp_k, vpd, lai, fpar, rad).round(6) == 4.49e-4
    cases = [
        ((np.arange(98e3, 103e3, 1e3), temp_k, vpd, lai, fpar, rad),
         [1.623, 1.62, 1.618, 1.615, 1.612]),
        ((pressure, 273.15 + np.array([0, 10, 20, 30, 40]), vpd, lai, fpar, rad),
         [0., 0., 0., 1.618, 3.222]),
        ((pressure, temp_k, np.arange(0, 5000, 1000), lai, fpar, rad),
         [5.382, 1.618, 0, 0, 0]),
        ((pressure, temp_k, vpd, np.arange(0.5, 3, 0.5), fpar, rad),
         [1.174, 1.478, 1.618, 1.699, 1.752]),
        ((pressure, temp_k, vpd, lai, f, rad), [1.611, 1.615, 1.618, 1.621, 1.624]),
        ((pressure, temp_k, vpd, lai, fpar, np.arange(3e3, 8e3, 1e3)),
         [0.974, 1.296, 1.618, 1.94, 2.262]),
    ]
    for args, want in cases:
        assert np.array_equal((h * m.evaporation_wet_canopy(*args)).round(3), want)
    pp = np.array((100e3, 80e3, 100e3, 80e3, 81.8e3))
    tt = 273.15 + np.array((10, 10, 25, 25, 25))
    assert np.array_equal(orc.psychrometric_constant(pp, tt).round(2),
                          [65.74, 52.59, 66.69, 53.35, 54.55])
    assert np.array_equal(orc.svp_slope(273.15 + np.array([10., 20., 25.])).round(1),
                          [82.3, 144.8, 188.8])
    # tests.py:64-90 -- the cross-interface consistency test
    drv = list(z['drivers'])
    vec = list(z['params'])
    assert round(float(orc.static_et(vec, *drv)), 2) == 40.94
    d, n = orc.static_evapotranspiration(vec, *drv)
    assert round(float(d + n), 2) == 40.94
    d2, n2 = m.evapotranspiration(*drv)
    assert round(float(d2 * orc.lhv(293) + n2 * orc.lhv(290)), 1) == 41.0


def test_unit_fixture_bit_exact(golden_dir):
    z = load(golden_dir, 'unit_tests')
    P = dict(zip(orc.PARAM_ORDER, z['params']))
    m = orc.InstanceModel(P)
    drv = list(z['drivers'])
    vec = list(z['params'])
    assert same(orc.static_et(vec, *drv), z['et_static_et'])
    sep = m.evapotranspiration(*drv, separate=True)
    assert same(sep[0][2], z['et_inst_day_trans'])
    assert same(sep[1][1], z['et_inst_night_soil'])
    assert same(orc.svp(z['svp_slope_in']), z['svp'])
    assert same(orc.lhv(z['svp_slope_in']), z['lhv'])
    assert same(orc.air_density(z['svp_slope_in'], 95e3, np.array([0.1, 0.5, 0.9])),
                z['air_density'])
    assert same(orc.rhumidity_clamped(z['svp_slope_in'], z['rhumidity_in_vpd']), z['rhumidity'])
    assert same(orc.vpd_from_qv(z['vpd_in_qv'], 95e3, z['svp_slope_in']), z['vpd'])
    assert same(orc.air_pressure(z['air_pressure_in']), z['air_pressure'])
    assert same(m.surface_conductance(z['surface_conductance_in_tmin'],
                                      z['surface_conductance_in_vpd']),
                z['surface_conductance'])


def test_verify_inline_arrays(golden_dir):
    'tests/verification/verify.py:48-85 inline arrays, PFT 7.'
    z = load(golden_dir, 'verify')
    vec = list(z['params'])
    lw_d, lw_n, sw_d, sw_n, alb, tann, fc, vpd_n = z['scalars']
    a = (lw_d, lw_n, sw_d, sw_n, alb, z['temp'], z['temp'], tann, z['tmin'], z['vpd_day'],
         z['vpd_day'], z['pressure'], fc, z['lai'])
    m = orc.InstanceModel(dict(zip(orc.PARAM_ORDER, vec)))
    sep = m.evapotranspiration(*a, separate=True)
    for h, hn in enumerate(('day', 'night')):
        for c, cn in enumerate(('canopy', 'soil', 'trans')):
            assert same(sep[h][c], z['inst_%s_%s' % (hn, cn)]), (hn, cn)
    b = a[:10] + (vpd_n,) + a[11:]
    d, n = orc.static_evapotranspiration(vec, *b)
    assert same(d, z['static_day']) and same(n, z['static_night'])
    # values quoted in SURVEY.md 8(c)
    assert np.allclose(z['static_et'], [56.37724313202413, 46.91395550198356,
                                        44.65166377561862], rtol=0, atol=1e-12)
    assert np.allclose(z['lhv'], [2470184.48771, 2455251.68213, 2441553.3254], atol=1e-4)
    # notebook scalar, …Forward_Run_example.ipynb:222
    nb = orc.InstanceModel(dict(zip(orc.PARAM_ORDER, z['notebook_params']))) \
        .evaporation_soil(101e3, 293.15, 1000, 0.5, 100)
    assert float(nb) == 5.083295223395212e-06 == float(z['notebook_scalar'])


def test_divergence_and_nan_tables(golden_dir):
    z = load(golden_dir, 'tables')
    vec = list(z['params'])
    m = orc.InstanceModel(dict(zip(orc.PARAM_ORDER, vec)))
    for i, name in enumerate(z['case_names']):
        drv = [np.float64(v) for v in z['drivers'][i]]
        with np.errstate(all='ignore'):
            d, n = orc.static_evapotranspiration(vec, *drv)
            sep = m.evapotranspiration(*drv, separate=True)
        assert same(d, z['static_day'][i]) and same(n, z['static_night'][i]), name
        for h, hn in enumerate(('day', 'night')):
            for c, cn in enumerate(('canopy', 'soil', 'trans')):
                assert same(sep[h][c], z['inst_%s_%s' % (hn, cn)][i]), (name, hn, cn)
    base = [np.float64(v) for v in z['drivers'][0]]
    for key in ('closed', 'open'):
        drv = list(base)
        drv[NAMES.index('tmin')] = z['gate_%s_tmin' % key]
        assert same(orc.static_et(vec, *drv), z['gate_%s_et' % key])
    # documented facts: base case and the closed gate
    assert z['static_day'][0] == 38.19216280296294
    # (with the gate closed the reference's result does not even broadcast to tmin's shape)
    assert z['gate_closed_et'].shape == ()
    assert abs(float(z['gate_closed_et']) - (6.7454 + 12.738963243349616)) < 1e-3


@pytest.mark.parametrize('name', ['rand_phys_f64', 'rand_phys_fill_f64',
                                  'rand_stress_fill_f64', 'rand_phys_f32',
                                  'rand_stress_fill_f32'])
def test_random_fixtures_bit_exact(golden_dir, name):
    z = load(golden_dir, name)
    up = lambda v: v.astype(np.float64)
    drv = [up(z['drv_' + k]) if 'drv_' + k in z else 0 for k in NAMES]
    vec = [z['table'][i][z['pft']].reshape(1, -1) for i in range(11)]
    with np.errstate(all='ignore'):
        d, n = orc.static_evapotranspiration(vec, *drv)
        assert same(d, z['static_day']) and same(n, z['static_night'])
        assert same(orc.static_et(vec, *drv), z['static_et'])
        m = orc.InstanceModel(dict(zip(orc.PARAM_ORDER, z['inst_params'])))
        sep = m.evapotranspiration(*drv, separate=True)
    for h, hn in enumerate(('day', 'night')):
        for c, cn in enumerate(('canopy', 'soil', 'trans')):
            assert same(sep[h][c], z['inst_%s_%s' % (hn, cn)])


def test_gate_modes_of_the_oracle():
    'element / open gates are oracle-side extensions; they must bracket the reference gate.'
    d = synthetic.host_drivers(8, 16, 5)
    vec = [-8., 8., 650., 3000., .01, .01, 1e-5, .0024, 60., 95., 250.]
    drv = [d[k] for k in NAMES]
    g = orc.static_et(vec, *drv, gate='global')
    o = orc.static_et(vec, *drv, gate='open')
    assert same(g, o)  # some element is open in this batch
    e = orc.static_et(vec, *drv, gate='element')
    day = orc.static_evapotranspiration(vec, *drv, components=True)[0]
    assert np.all((e == g) | (day[2] != 0))


def test_constraint_golden(golden_dir):
    'oracle.constrain_by_map against the answers of the reference\'s own nested function.'
    g = np.load(os.path.join(golden_dir, 'constraint.npz'))
    for tag in 'abc':
        le, years, lhv, precip = (g['%s_%s' % (tag, k)] for k in ('le', 'years', 'lhv', 'precip'))
        got = orc.constrain_by_map(le.copy(), years, lhv, precip)
        assert got == float(g[tag + '_neg_nrmsd'])
        assert np.array_equal(orc.annual_mass_sums(le.copy(), years, lhv), g[tag + '_annual'])
    assert float(g['c_neg_nrmsd']) == 0.0
