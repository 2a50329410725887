'''
Elements placed on both sides of every switch of the model (shared by the host-math test and
the GPU test of the float64 fast path): see `build`.
'''
import numpy as np

from mod16_b200 import synthetic
from oracle import mod16_oracle as orc

NAMES = list(synthetic.DRIVER_NAMES)
BPLUT7 = [-8., 8.8, 650., 4400., .02, .02, 1e-5, .0055, 60., 95., 250.]


def build(semantics, dtype=np.float64):
    '''
    Returns (drivers as fourteen (1, n) float64 arrays, group names, group index per element,
    offset per element).  Offsets are 1e-13 ... 1e-7 on both sides (below 1e-13 the
    reference's own rounding picks the side); with ``dtype=np.float32`` they are 1e-8 ... 1e-3
    and every array is rounded to float32 afterwards -- the smallest offsets then fall within
    rounding of the switch, which is where the float32 path must hand over to its double code.  Groups: the f_wet jump at rh = 0.7 (:282),
    rh -> 0+ and rh == 0 (:281), rh == 1 and rh > 1 (negative VPD), |G| = 0.39 |A| on both
    half-days (:240), A_day = 0, the night condition of the soil radiation (:248-252; the day
    condition :245-247 cannot be reached once |G| <= 0.39 |A|), the 5 K day-night switch and
    the two ends of the annual-temperature window (:230-234), the ends of the Tmin and VPD
    ramps (:325-326) and of the r_tot ramp (:357-361).  Parameters: BPLUT7 for every element.
    '''
    rng = np.random.default_rng(2024)
    if np.dtype(dtype) == np.float64:
        eps = np.array([1e-13, 1e-12, 1e-11, 1e-10, 3e-10, 1e-9, 3e-9, 1e-8, 1e-7])
    else:
        eps = np.array([1e-8, 1e-7, 3e-7, 1e-6, 4e-6, 1e-5, 3e-5, 1e-4, 1e-3])
    offs = np.concatenate([eps, -eps])
    per = 40                                            # elements per (group, offset)
    groups = ['rh07', 'rh0', 'rh1', 'g_day', 'g_night', 'a_zero', 'soil_n', 'dt5', 'tann_hi',
              'tann_lo', 'tmin_lo', 'tmin_hi', 'vpd_lo', 'vpd_hi']
    n = len(groups) * len(offs) * per
    d = synthetic.host_drivers(1, n, 909, dtype=np.float64)
    d = dict((k, (np.array(v[0] if v.ndim == 2 else v, np.float64) if isinstance(v, np.ndarray)
                  else np.full(n, float(v)))) for k, v in d.items())
    # a common, well-conditioned base: warm days, 7 K cooler nights, window open
    d['temp_day'] = rng.uniform(283.0, 303.0, n)
    d['temp_night'] = d['temp_day'] - 7.0
    d['temp_annual'] = np.full(n, 288.0)
    d['tmin'] = d['temp_night'] - 1.0
    grp = np.repeat(np.arange(len(groups)), len(offs) * per)
    off = np.tile(np.repeat(offs, per), len(groups))
    p = dict(zip(orc.PARAM_ORDER, BPLUT7))

    def sel(name):
        return grp == groups.index(name)

    es_d = orc.svp(d['temp_day'])
    es_n = orc.svp(d['temp_night'])
    m = sel('rh07')
    d['vpd_day'][m] = (es_d * (1.0 - (0.7 + off)))[m]
    d['vpd_night'][m] = (es_n * (1.0 - (0.7 - off)))[m]
    m = sel('rh0')                                      # off < 0: vpd > svp, rh clamps to 0
    d['vpd_day'][m] = (es_d * (1.0 - off))[m]
    d['vpd_night'][m] = (es_n * (1.0 - 30.0 * off))[m]  # reaches past the 2e-5 guard's end
    m = sel('rh1')                                      # off < 0: negative VPD, rh > 1
    d['vpd_day'][m] = (es_d * off)[m]
    d['vpd_night'][m] = np.where(off > 0, 0.0, es_n * off)[m]
    a_day = d['sw_rad_day'] * (1.0 - d['sw_albedo']) + d['lw_net_day']      # :225
    for name, t_key, lw_key in (('g_day', 'temp_day', 'lw_net_day'),
                                ('g_night', 'temp_night', 'lw_net_night')):
        m = sel(name)
        g = 4.73 * (d[t_key] - 273.15) - 20.87
        want = np.abs(g) / 0.39 * (1.0 + off) * np.where(rng.random(n) < 0.5, 1.0, -1.0)
        if name == 'g_day':
            d[lw_key][m] = (want - d['sw_rad_day'] * (1.0 - d['sw_albedo']))[m]
        else:
            d[lw_key][m] = want[m]
    m = sel('a_zero')
    d['lw_net_day'][m] = (off * 10.0 - d['sw_rad_day'] * (1.0 - d['sw_albedo']))[m]
    # rad_night - G_night = -0.5 A_day (1 + off) with G_night = 20 W m-2 inside its limit
    # (|G_n| <= 0.39 * 60) and G_day = 53.1 inside its own (0.39 * 160 = 62.4)
    m = sel('soil_n')
    d['temp_night'][m] = 273.15 + (20.0 + 20.87) / 4.73
    d['temp_day'][m] = d['temp_night'][m] + 7.0
    d['tmin'][m] = d['temp_night'][m] - 1.0
    d['lw_net_night'][m] = -60.0
    d['lw_net_day'][m] = (160.0 / (1.0 + off) - d['sw_rad_day'] * (1.0 - d['sw_albedo']))[m]
    m = sel('dt5')
    d['temp_night'][m] = (d['temp_day'] - 5.0 + off * 100.0)[m]
    m = sel('tann_hi')
    d['temp_annual'][m] = (273.15 + 25.0) + (off * 100.0)[m]
    m = sel('tann_lo')
    lo = 273.15 + (p['tmin_open'] if semantics == 'static' else p['tmin_close'])
    d['temp_annual'][m] = lo + (off * 100.0)[m]
    for name, base in (('tmin_lo', p['tmin_close']), ('tmin_hi', p['tmin_open'])):
        m = sel(name)
        d['tmin'][m] = (273.15 + base) + (off * 100.0)[m]
    for name, base in (('vpd_lo', p['vpd_open']), ('vpd_hi', p['vpd_close'])):
        m = sel(name)
        d['vpd_day'][m] = base * (1.0 + off[m])
        d['vpd_night'][m] = base * (1.0 - off[m])
    drv = [d[k].reshape(1, n).astype(dtype) for k in NAMES]
    return drv, groups, grp, off


def reference(semantics, drv):
    '''
    Oracle components (day, night) of `drv` with the per-element gate, and the mass scale;
    float32 drivers are evaluated on their float64 values.
    '''
    import parity
    drv = [np.asarray(v, np.float64) for v in drv]
    p = dict(zip(orc.PARAM_ORDER, BPLUT7))
    with np.errstate(all='ignore'):
        if semantics == 'static':
            day, night = orc.static_evapotranspiration(BPLUT7, *drv, gate='element',
                                                       components=True)
            return day, night, 1.0
        day, night = orc.InstanceModel(p).evapotranspiration(*drv, separate=True)
        return day, night, parity.MASS_SCALE
