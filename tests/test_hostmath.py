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
