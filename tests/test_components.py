'''
K4 -- the component methods and physics helpers of the reference's public API.

CPU part: the oracle restatement against tests/golden/components.npz (outputs of the
UNMODIFIED reference on seeded inputs, oracle/make_golden.py) and against the reference's
own known answers (tests/tests.py:92-232).
GPU part: `mod16_b200.MOD16` / module functions through `mod16_elementwise_*` against the same
golden vectors, and the reference's unit tests transcribed call for call.
Criteria: float64 within 1e-10 relative (parity.py); float32 inputs are evaluated in double
on the device and rounded once, so they must match the float64 reference of the same float32
inputs to float32 rounding (2e-7 relative); identical NaN masks.
'''
import os

import numpy as np
import pytest

import parity
from oracle import mod16_oracle as orc

PARAMS = dict(gl_sh=0.01, gl_wv=0.01, g_cuticular=1e-5, tmin_close=-8, tmin_open=8,
              vpd_open=650, vpd_close=3000, rbl_min=60, rbl_max=90, csl=2.4e-3, beta=250)


@pytest.fixture(scope='module')
def gold(golden_dir):
    return np.load(os.path.join(golden_dir, 'components.npz'))


def _calls(z, model, mod):
    'Every golden entry as (tag, callable returning the value(s)); `model`/`mod` supply the API.'
    x = dict((k[3:], z[k]) for k in z.files if k.startswith('in_'))
    P = dict(zip(orc.PARAM_ORDER, z['params']))
    a5 = (x['pressure'], x['temp_k'], x['vpd'], x['fpar'], x['rad_soil'])
    a6 = (x['pressure'], x['temp_k'], x['vpd'], x['lai'], x['fpar'], x['rad_canopy'])
    a7 = (x['lw_d'], x['sw_d'], x['alb'], x['pressure'], x['temp_k'], x['vpd'], x['fpar'])
    a9 = (x['lw_d'], x['lw_n'], x['sw_d'], x['sw_n'], x['alb'], x['t_d'], x['t_n'], x['t_ann'],
          x['fpar'])
    opt = (x['lhv'], x['rh'], x['fw'])
    bp = (P['vpd_open'], P['vpd_close'], P['rbl_min'], P['rbl_max'])
    rad_d = x['sw_d'] * (1 - x['alb']) + x['lw_d']
    m = model(P)
    return [
        ('evaporation_soil', lambda: m.evaporation_soil(*a5)),
        ('evaporation_soil_opt', lambda: m.evaporation_soil(*a5, x['r_corr'], *opt)),
        ('potential_soil_evaporation',
         lambda: m.potential_soil_evaporation(*a5, None, None, None, None, *bp)),
        ('potential_soil_evaporation_opt',
         lambda: m.potential_soil_evaporation(*a5, x['r_corr'], *opt, *bp)),
        ('evaporation_wet_canopy', lambda: m.evaporation_wet_canopy(*a6)),
        ('evaporation_wet_canopy_opt', lambda: m.evaporation_wet_canopy(*a6, *opt)),
        ('transpiration_day', lambda: m.transpiration(*a6, x['tmin'], daytime=True)),
        ('transpiration_day_opt',
         lambda: m.transpiration(*a6, x['tmin'], x['r_corr'], *opt, daytime=True)),
        ('transpiration_night', lambda: m.transpiration(*a6, x['tmin'], daytime=False)),
        ('transpiration_night_opt',
         lambda: m.transpiration(*a6, x['tmin'], x['r_corr'], *opt, daytime=False)),
        ('surface_conductance', lambda: m.surface_conductance(x['tmin'], x['vpd'])),
        ('radiation_soil', lambda: m.radiation_soil(*a9)),
        ('soil_heat_flux',
         lambda: m.soil_heat_flux(rad_d, x['lw_n'], x['t_d'], x['t_n'], x['t_ann'])),
        ('potential_transpiration', lambda: m.potential_transpiration(*a7)),
        ('potential_transpiration_opt',
         lambda: m.potential_transpiration(*a7, x['rh'], x['fw'], 1.1)),
        ('svp', lambda: mod['svp'](x['temp_k'])),
        ('svp_slope', lambda: mod['svp_slope'](x['temp_k'])),
        ('svp_slope_opt', lambda: mod['svp_slope'](x['temp_k'], x['lhv'] * 1e-3)),
        ('lhv', lambda: mod['lhv'](x['temp_k'])),
        ('psychrometric', lambda: mod['psychrometric'](x['pressure'], x['temp_k'])),
        ('air_density', lambda: m.air_density(x['temp_k'], x['pressure'], x['rh'])),
        ('rhumidity', lambda: m.rhumidity(x['temp_k'], x['vpd'])),
        ('vpd', lambda: m.vpd(x['qv'], x['pressure'], x['temp_k'])),
        ('air_pressure', lambda: m.air_pressure(x['elev'])),
        ('radiation_net', lambda: mod['radiation_net'](x['sw'], x['alb'], x['temp_k'])),
    ]


def _golden(z, tag):
    if tag in z.files:
        return [z[tag]]
    return [z[tag + '_0'], z[tag + '_1']]


class _OracleModel(orc.InstanceModel):
    air_density = staticmethod(orc.air_density)
    rhumidity = staticmethod(orc.rhumidity_clamped)
    vpd = staticmethod(orc.vpd_from_qv)
    air_pressure = staticmethod(orc.air_pressure)


def test_oracle_components_match_the_frozen_reference_outputs(gold):
    mod = dict(svp=orc.svp, svp_slope=orc.svp_slope, lhv=orc.lhv,
               psychrometric=orc.psychrometric_constant, radiation_net=orc.radiation_net)
    n = 0
    for tag, fn in _calls(gold, _OracleModel, mod):
        with np.errstate(all='ignore'):
            got = fn()
        got = list(got) if isinstance(got, (tuple, list)) else [got]
        for g, r in zip(got, _golden(gold, tag)):
            assert np.array_equal(np.asarray(g, np.float64), r, equal_nan=True), tag
            n += 1
    assert n == 29
    swrad, albedo, tk = gold['radiation_net_known_in']
    assert list(orc.radiation_net(swrad, albedo, tk).round(1)) == \
        [223.3, 2923.3, 23.3, 923.3, 241.8, 2941.8, 41.8, 941.8]


def _product_api():
    import mod16_b200
    mod = dict(svp=mod16_b200.svp, svp_slope=mod16_b200.svp_slope,
               lhv=mod16_b200.latent_heat_vaporization,
               psychrometric=mod16_b200.psychrometric_constant,
               radiation_net=mod16_b200.radiation_net)
    return mod16_b200.MOD16, mod


@pytest.mark.gpu
def test_components_match_the_frozen_reference_outputs(cuda_lib, gold):
    model, mod = _product_api()
    for tag, fn in _calls(gold, model, mod):
        got = fn()
        got = list(got) if isinstance(got, (tuple, list)) else [got]
        ref = _golden(gold, tag)
        assert len(got) == len(ref), tag
        for g, r in zip(got, ref):
            scale = max(1e-300, float(np.nanmax(np.abs(r[np.isfinite(r)]))) * 1e-3) \
                if np.isfinite(r).any() else 1.0
            parity.assert_parity(g, r, np.float64, scale, what=tag)


@pytest.mark.gpu
def test_components_float32_inputs(cuda_lib, gold):
    'float32 in, float32 out; evaluated in double against the float64 oracle on the same inputs.'
    model, mod = _product_api()
    z32 = dict((k, gold[k].astype(np.float32) if k.startswith('in_') else gold[k])
               for k in gold.files)

    class Z(dict):
        files = list(z32.keys())
    z32 = Z(z32)
    z64 = Z(dict((k, v.astype(np.float64)) for k, v in z32.items()))
    omod = dict(svp=orc.svp, svp_slope=orc.svp_slope, lhv=orc.lhv,
                psychrometric=orc.psychrometric_constant, radiation_net=orc.radiation_net)
    for (tag, fn), (_, ofn) in zip(_calls(z32, model, mod), _calls(z64, _OracleModel, omod)):
        got = fn()
        with np.errstate(all='ignore'):
            ref = ofn()
        got = list(got) if isinstance(got, (tuple, list)) else [got]
        ref = list(ref) if isinstance(ref, (tuple, list)) else [ref]
        for g, r in zip(got, ref):
            g = np.asarray(g)
            assert g.dtype == np.float32, tag
            r = np.asarray(r, np.float64)
            assert np.array_equal(np.isnan(g), np.isnan(r)), tag
            ok = np.isfinite(r)
            assert np.all(np.abs(g[ok] - r[ok]) <= 2e-7 * np.abs(r[ok]) + 1e-30), tag


@pytest.mark.gpu
def test_reads_like_the_reference_unit_tests(cuda_lib):
    'tests/tests.py:92-232 of the reference, call for call, with its known answers.'
    from mod16_b200 import MOD16, psychrometric_constant, radiation_net, svp_slope
    pressure, temp_k, vpd, lai, fpar = 100e3, 273.15 + 30, 1000, 1.5, 0.5
    rad_canopy = rad_soil = 5000
    tmin = 285
    r_corr = (101300 / pressure) * (temp_k / 293.15)**1.75
    _pressure = np.arange(98e3, 103e3, 1e3)
    _temp_k = 273.15 + np.array([0, 10, 20, 30, 40])
    _vpd = np.arange(0, 5000, 1000)
    _lai = np.arange(0.5, 3, 0.5)
    _fpar = np.array([0.1, 0.3, 0.5, 0.7, 0.9])
    _rad_canopy = np.arange(3e3, 8e3, 1e3)
    model = MOD16(PARAMS)
    evap = 60 * 60 * model.evaporation_soil(pressure, temp_k, vpd, fpar, rad_soil, r_corr)
    assert evap.round(3) == 3.102
    evap = 60 * 60 * model.evaporation_soil(pressure, temp_k, vpd, _fpar, rad_soil, r_corr)
    assert np.equal(evap.round(3), np.array([3.128, 3.115, 3.102, 3.089, 3.076])).all()
    trans = 60 * 60 * model.transpiration(
        pressure, temp_k, vpd, lai, fpar, rad_canopy, tmin, r_corr, daytime=True)
    assert trans.round(3) == 1.248
    trans = 60 * 60 * model.transpiration(
        pressure, temp_k, vpd, lai, fpar, rad_canopy, tmin, r_corr, daytime=False)
    assert trans.round(3) == 0.011
    evap = model.evaporation_wet_canopy(pressure, temp_k, vpd, lai, fpar, rad_canopy).round(6)
    assert evap == 4.49e-4
    wc = lambda *a: (60 * 60 * model.evaporation_wet_canopy(*a)).round(3)
    assert np.equal(wc(_pressure, temp_k, vpd, lai, fpar, rad_canopy),
                    np.array([1.623, 1.62, 1.618, 1.615, 1.612])).all()
    assert np.equal(wc(pressure, _temp_k, vpd, lai, fpar, rad_canopy),
                    np.array([0., 0., 0., 1.618, 3.222])).all()
    assert np.equal(wc(pressure, temp_k, _vpd, lai, fpar, rad_canopy),
                    np.array([5.382, 1.618, 0, 0, 0])).all()
    assert np.equal(wc(pressure, temp_k, vpd, _lai, fpar, rad_canopy),
                    np.array([1.174, 1.478, 1.618, 1.699, 1.752])).all()
    assert np.equal(wc(pressure, temp_k, vpd, lai, _fpar, rad_canopy),
                    np.array([1.611, 1.615, 1.618, 1.621, 1.624])).all()
    assert np.equal(wc(pressure, temp_k, vpd, lai, fpar, _rad_canopy),
                    np.array([0.974, 1.296, 1.618, 1.94, 2.262])).all()
    p = np.array((100e3, 80e3, 100e3, 80e3))
    t = 273.15 + np.array((10, 10, 25, 25))
    answer = [65.74, 52.59, 66.69, 53.35]
    for i in range(4):
        assert answer[i] == psychrometric_constant(p[i], t[i]).round(2)
    assert 54.55 == np.round(psychrometric_constant(81.8e3, 25 + 273.15), 2)
    swrad = np.array((500, 5000, 500, 5000, 500, 5000, 500, 5000))
    albedo = np.array((0.4, 0.4, 0.8, 0.8, 0.4, 0.4, 0.8, 0.8))
    tk = 273.15 + np.array((10, 10, 10, 10, 25, 25, 25, 25))
    answer = [223.3, 2923.3, 23.3, 923.3, 241.8, 2941.8, 41.8, 941.8]
    for i in range(8):
        assert answer[i] == radiation_net(swrad[i], albedo[i], tk[i]).round(1)
    assert 82.3 == svp_slope(273.15 + 10).round(1)
    assert 144.8 == svp_slope(273.15 + 20).round(1)
    assert 188.8 == svp_slope(273.15 + 25).round(1)


@pytest.mark.gpu
def test_daylength_composite(cuda_lib):
    'verify2.py:101-115: instance day / night rates to kg m-2 per 8 days.'
    from mod16_b200 import MOD16, synthetic
    from mod16_b200.components import daylength_composite
    T, N = 46, 33
    d = synthetic.host_drivers(T, N, 12)
    drv = [d[k] for k in synthetic.DRIVER_NAMES]
    P = dict(zip(orc.PARAM_ORDER, [-8., 8.8, 650., 4400., .02, .02, 1e-5, .0055, 60., 95., 250.]))
    day, night = MOD16(P).evapotranspiration(*drv)
    with np.errstate(all='ignore'):
        rd, rn = orc.InstanceModel(P).evapotranspiration(*drv)
    rng = np.random.default_rng(4)
    updown = rng.uniform(4, 20, (T, 1)) + np.zeros((T, N))
    updown[3] = np.nan                                  # polar night: sunrise_sunset gave none
    got = daylength_composite(day, night, updown)
    ref = orc.daylength_composite(rd, rn, updown)
    parity.assert_parity(got, ref, np.float64, 1e-3, what='8-day composite')
    assert np.isnan(got[3]).all()


@pytest.mark.gpu
def test_component_errors_and_shapes(cuda_lib):
    import torch
    from mod16_b200 import MOD16, _abi, components, svp
    lib = _abi.load()
    ew = _abi.Elementwise()
    ew.op = 99
    assert lib.mod16_elementwise_f64(ew, 1, 1, None) == -3
    ew.op, ew.present = _abi.OPS.index('air_density'), 0b011   # a required argument missing
    assert lib.mod16_elementwise_f64(ew, 1, 1, None) == -1
    # broadcasting (T, N) x (N,) x scalar, CUDA tensors in -> CUDA tensor out
    t = torch.full((3, 5), 293.15, device='cuda', dtype=torch.float64)
    rho = MOD16.air_density(t, np.full(5, 1e5), 0.5)
    assert isinstance(rho, torch.Tensor) and tuple(rho.shape) == (3, 5)
    assert np.allclose(rho.cpu().numpy(), orc.air_density(293.15, 1e5, 0.5), rtol=1e-12)
    assert svp(np.zeros((0,))).shape == (0,)
    with pytest.raises(KeyError):
        components.elementwise('no_such_op', [1.0])
