'''
TEST INFRASTRUCTURE ONLY -- CPU restatement (numpy) of the reference algorithm.

This file is the parity oracle for the CUDA path in ``mod16_b200/csrc``.  Only
``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline`` / ``--impl
reference`` legs of ``bench.py`` may import it.  The product package
``mod16_b200`` never imports anything under ``oracle/``.

It restates, operation for operation (same association order, so that float64
results are bit-identical with the reference on the same inputs), the forward
model of arthur-e/MOD16.  All citations are ``file:line`` under
``/root/reference``.

Parity is PINNED: ``oracle/make_golden.py`` imports the unmodified reference
(with ``oracle/refshim/mod17`` standing in for the one external function it
needs) in the build container, checks this restatement bit-for-bit against it on
the reference's own known answers (tests/tests.py, tests/verification/verify.py
inline arrays, the notebook scalar) and on randomised inputs with NaN / fill
injection, and freezes the vectors in ``tests/golden/*.npz``.
``tests/test_oracle_golden.py`` replays them without the reference.

Two semantics are restated, because the reference has two forward paths that
differ:

* ``static``   -- ``MOD16._evapotranspiration`` / ``MOD16._et``
                  (mod16/__init__.py:163-382): W m-2, per-pixel parameter
                  vectors, no clamps, batch-global transpiration gate.
* ``instance`` -- ``MOD16.evapotranspiration`` and the component methods
                  (mod16/__init__.py:675-1258): kg m-2 s-1, one parameter set,
                  with clamps, night cuticular transpiration.
'''
import warnings
import numpy as np

# mod16/__init__.py:106-110, :152-155
PFT_VALID = (1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 12)
SIGMA = 5.67e-8
CP_AIR = 1013
EPSILON = 0.622
PARAM_ORDER = (
    'tmin_close', 'tmin_open', 'vpd_open', 'vpd_close', 'gl_sh', 'gl_wv',
    'g_cuticular', 'csl', 'rbl_min', 'rbl_max', 'beta')
TINY = 1e-7


# --------------------------------------------------------------------------
# Physics helpers
# --------------------------------------------------------------------------

def lhv(temp_k):
    'Latent heat of vaporisation, J kg-1 (mod16/__init__.py:121).'
    return (2.501 - 0.002361 * (temp_k - 273.15)) * 1e6


def svp(temp_k):
    'Saturation vapour pressure, Pa (mod16/__init__.py:1365-1367).'
    tc = temp_k - 273.15
    return 1e3 * 0.6108 * np.exp((17.27 * tc) / (tc + 237.3))


def svp_slope(temp_k, s=None):
    'Slope of the SVP curve, Pa K-1 (mod16/__init__.py:1395-1397).'
    if s is None:
        s = svp(temp_k)
    return 17.38 * 239.0 * s / (239.0 + temp_k - 273.15)**2


def psychrometric_constant(pressure, temp_k):
    'Psychrometric constant, Pa K-1 (mod16/__init__.py:1288-1290).'
    return (CP_AIR * pressure) / (lhv(temp_k) * EPSILON)


def air_density(temp_k, pressure, rh):
    'NIST simplified moist-air density, kg m-3 (mod16/__init__.py:409-412).'
    return np.divide(
        0.348444 * (pressure / 100) - (rh * 100) *
        (0.00252 * (temp_k - 273.15) - 0.020582), temp_k)


def rhumidity_clamped(temp_k, vpd):
    'RH clamped to [0, 1]; instance path only (mod16/__init__.py:670-673).'
    esat = svp(temp_k)
    avp = esat - vpd
    rh = avp / esat
    return np.where(avp < 0, 0, np.where(rh > 1, 1, rh))


def vpd_from_qv(qv10m, pressure, tmean):
    'Driver prep: VPD from specific humidity (mod16/__init__.py:639-644).'
    tc = tmean - 273.15
    avp = (qv10m * pressure) / (0.622 + (0.379 * qv10m))
    es = 610.7 * np.exp((17.38 * tc) / (239 + tc))
    return es - avp


def air_pressure(elevation_m):
    'Driver prep: pressure from elevation (mod16/__init__.py:111-118,446-447).'
    rate = 9.80665 / (0.0065 * (8.3143 / 28.9644e-3))
    ratio = 1 - ((0.0065 * elevation_m) / 288.15)
    return 101325.0 * np.power(ratio, rate)


def radiation_net(sw_rad, sw_albedo, temp_k):
    'DEPRECATED net radiation of Cleugh et al. 2007 (mod16/__init__.py:1333-1337).'
    emis_surface = 0.97
    emis_atmos = 1 - 0.26 * np.exp(-7.77e-4 * np.power(temp_k - 273.15, 2))
    return sw_rad * (1 - sw_albedo) + \
        SIGMA * (emis_atmos - emis_surface) * np.power(temp_k, 4)


def ramp_up(x, xmin, xmax):
    '''
    ``mod17.linear_constraint(xmin, xmax)`` -- EXTERNAL to /root/reference
    (imported mod16/__init__.py:104; used :325, :1148).  1 at/above xmax, 0
    below xmin, linear between.  NaN in, NaN out.
    '''
    return np.where(x >= xmax, 1, np.where(
        x < xmin, 0, np.divide(np.subtract(x, xmin), xmax - xmin)))


def ramp_down(x, xmin, xmax):
    "``mod17.linear_constraint(xmin, xmax, 'reversed')`` (used :326, :1149)."
    return np.where(x >= xmax, 0, np.where(
        x < xmin, 1, 1 - np.divide(np.subtract(x, xmin), xmax - xmin)))


def r_correction(pressure, temp_k):
    'Equation 13 correction factor (mod16/__init__.py:292, :512, :771).'
    return (101300 / pressure) * (temp_k / 293.15)**1.75


def f_wet_of(rh):
    'Wet surface fraction (mod16/__init__.py:282, :764).'
    return np.where(rh < 0.7, 0, np.power(rh, 4))


def r_radiative(rho, temp_k):
    'Resistance to radiative heat transfer (mod16/__init__.py:298-299).'
    return (rho * CP_AIR) / (4 * SIGMA * temp_k**3)


def r_total_soil(vpd, vpd_open, vpd_close, rbl_min, rbl_max):
    'Piecewise boundary-layer resistance (mod16/__init__.py:357-361, :527-531).'
    return np.where(vpd <= vpd_open, rbl_min, np.where(
        vpd >= vpd_close, rbl_max, rbl_max - (
            (rbl_max - rbl_min) * (vpd_close - vpd)) / (vpd_close - vpd_open)))


# --------------------------------------------------------------------------
# Soil heat flux / radiation reaching the soil -- shared skeleton, the two
# paths differ only in the annual-temperature test (see below)
# --------------------------------------------------------------------------

def _ground_heat(rad_day, rad_night, temp_day, temp_night, cond):
    'mod16/__init__.py:235-241 (static) == :1108-1119 (instance).'
    out = []
    for rad_i, temp_i in ((rad_day, temp_day), (rad_night, temp_night)):
        g = np.where(cond, (4.73 * (temp_i - 273.15)) - 20.87, 0)
        g = np.where(np.abs(g) > (0.39 * np.abs(rad_i)), 0.39 * rad_i, g)
        out.append(g)
    return out


def _soil_radiation(rad_day, rad_night, g_day, g_night, fpar):
    'mod16/__init__.py:245-254 (static) == :1039-1053 (instance).'
    g_day = np.where(
        np.logical_and(rad_day - g_day < 0, rad_day > 0), rad_day, g_day)
    g_night = np.where(
        np.logical_and(rad_day > 0, (rad_night - g_night) < (-0.5 * rad_day)),
        rad_night + (0.5 * rad_day), g_night)
    return ((1 - fpar) * (rad_day - g_day), (1 - fpar) * (rad_night - g_night))


# --------------------------------------------------------------------------
# STATIC path: MOD16._evapotranspiration / MOD16._et
# --------------------------------------------------------------------------

def static_evapotranspiration(
        params, lw_net_day, lw_net_night, sw_rad_day, sw_rad_night,
        sw_albedo, temp_day, temp_night, temp_annual, tmin, vpd_day,
        vpd_night, pressure, fpar, lai, f_wet=None, tiny=1e-7,
        r_corr_list=None, gate='global', components=False):
    '''
    Restates mod16/__init__.py:196-382.  Returns ``[day, night]`` in W m-2.

    ``gate`` is an oracle-side extension used to check the CUDA gate modes:
    ``'global'`` is the reference behaviour (`np.any(g_surf > 0)` over the whole
    call, :343); ``'element'`` evaluates the same predicate per element (what a
    loop of scalar calls does, e.g. sensitivity.py:114); ``'open'`` always
    computes transpiration.  ``components=True`` returns, per half, the tuple
    (e_canopy, e_soil, transpiration) instead of the sum.  ``f_wet`` and
    ``tiny`` are accepted and ignored exactly as the reference does (:191,
    :282).
    '''
    p = params
    tiny = 1e-7
    rad_day = sw_rad_day * (1 - sw_albedo) + lw_net_day           # :225
    rad_night = lw_net_night                                      # :226
    # NOTE the static path tests temp_annual against params[1] (tmin_open),
    #   strictly greater (:233); the instance path uses tmin_close, >= (:1106)
    cond = np.logical_and(np.logical_and(
        temp_annual < (273.15 + 25), temp_annual > (273.15 + p[1])),
        (temp_day - temp_night) >= 5)                             # :230-234
    g_day, g_night = _ground_heat(
        rad_day, rad_night, temp_day, temp_night, cond)
    rad_soil = _soil_radiation(rad_day, rad_night, g_day, g_night, fpar)

    halves = []
    for i in (0, 1):
        temp_k = (temp_day, temp_night)[i]
        vpd = (vpd_day, vpd_night)[i]
        sw_rad = (sw_rad_day, sw_rad_night)[i]
        lw_net = (lw_net_day, lw_net_night)[i]
        a_soil = rad_soil[i]
        a_net = sw_rad * (1 - sw_albedo) + lw_net                 # :272
        a_canopy = fpar * a_net                                   # :274
        es = svp(temp_k)                                          # :279
        rh = (es - vpd) / es                                      # :280
        rh = np.where(rh < 0, 0, rh)                              # :281
        fw = np.where(rh < 0.7, 0, np.power(rh, 4))               # :282
        s = svp_slope(temp_k, es)                                 # :284
        lam = lhv(temp_k)                                         # :286
        gamma = psychrometric_constant(pressure, temp_k)          # :288
        if r_corr_list is None:
            r_corr = (101300 / pressure) * (temp_k / 293.15)**1.75  # :292
        else:
            r_corr = r_corr_list[i]                               # :294
        rho = air_density(temp_k, pressure, rh)                   # :296
        r_r = (rho * CP_AIR) / (4 * SIGMA * temp_k**3)            # :298-299
        with warnings.catch_warnings():
            warnings.simplefilter('ignore')
            with np.errstate(all='ignore'):
                r_h = 1 / (p[4] * lai * fw)                       # :305
                r_e = 1 / (p[5] * lai * fw)                       # :308
                r_aw = np.divide(r_h * r_r, r_h + r_r)            # :311
                e = np.divide(
                    fw * ((s * a_canopy) + (
                        rho * CP_AIR * fpar * vpd * 1 / r_aw)),
                    s + ((pressure * CP_AIR * r_e) *
                         1 / (lam * EPSILON * r_aw)))             # :314-319
        e_canopy = np.where(lai * fw <= tiny, 0, e)               # :320

        g_surf = 0                                                # :323
        if i == 0:
            g_surf = (p[7] * ramp_up(tmin - 273.15, p[0], p[1]) *
                      ramp_down(vpd, p[2], p[3]))                 # :325-327
        g_surf = g_surf / r_corr                                  # :328
        g_cut = p[6] / r_corr                                     # :329
        gl_sh = p[4] * lai * (1 - fw)                             # :333
        with np.errstate(all='ignore'):
            g = ((gl_sh * (g_surf + g_cut)) / (gl_sh + g_surf + g_cut))  # :334
        g_canopy = np.where(
            np.logical_and(lai > 0, (1 - fw) > 0), g, tiny)       # :336-337
        r_ad = (1 / p[4] * r_r) / (1 / p[4] + r_r)                # :340
        with np.errstate(all='ignore'):
            t_open = (1 - fw) * ((s * a_canopy) + (
                rho * CP_AIR * fpar * (vpd / r_ad)))              # :344-345
            t_open = t_open / (s + gamma * (1 + (1 / g_canopy) / r_ad))  # :346
        if gate == 'global' or i == 1:
            # at night g_surf == 0, so the reference's gate is never open there;
            #   the extension modes below only concern the day half
            t = t_open if np.any(g_surf > 0) else 0               # :343-348
        elif gate == 'element':
            t = np.where(g_surf > 0, t_open, 0)
        elif gate == 'open':
            t = t_open
        elif gate == 'closed':
            t = 0
        else:
            raise ValueError(gate)

        r_tot = r_total_soil(vpd, p[2], p[3], p[8], p[9])         # :357-361
        r_tot = r_tot / r_corr                                    # :363
        with np.errstate(all='ignore'):
            r_as = (r_tot * r_r) / (r_tot + r_r)                  # :365
            numer = (s * a_soil) + (
                rho * CP_AIR * (1 - fpar) * (vpd / r_as))         # :367-368
            denom = (s + gamma * (r_tot / r_as))                  # :369
            e_sat = (numer * fw) / denom                          # :371
            e_unsat = (numer * (1 - fw)) / denom                  # :373
            e_soil = e_sat + e_unsat * rh**(vpd / p[10])          # :376
        if components:
            halves.append((e_canopy, e_soil, t))
        else:
            halves.append(t + e_canopy + e_soil)                  # :380
    return halves


def static_et(params, *drivers, r_corr_list=None, gate='global', **kw):
    'Restates MOD16._et (mod16/__init__.py:188-193): day + night.'
    day, night = static_evapotranspiration(
        params, *drivers[:14], r_corr_list=r_corr_list, gate=gate)
    return np.add(day, night)


# --------------------------------------------------------------------------
# INSTANCE path: MOD16.evapotranspiration and its component methods
# --------------------------------------------------------------------------

class InstanceModel(object):
    '''
    Restates the instance interface of ``MOD16`` (mod16/__init__.py:157-160,
    :675-1258) for ONE parameter set (dict keyed by ``PARAM_ORDER``; a missing
    key raises KeyError like the reference constructor).
    '''

    def __init__(self, params):
        self.params = params
        for key in PARAM_ORDER:
            setattr(self, key, params[key])

    # :1055-1119
    def soil_heat_flux(self, rad_day, rad_night, temp_day, temp_night,
                       temp_annual):
        cond = np.logical_and(np.logical_and(
            temp_annual < (273.15 + 25),
            temp_annual >= (273.15 + self.tmin_close)),
            (temp_day - temp_night) >= 5)
        return _ground_heat(rad_day, rad_night, temp_day, temp_night, cond)

    # :963-1053
    def radiation_soil(self, lw_net_day, lw_net_night, sw_rad_day,
                       sw_rad_night, sw_albedo, temp_day, temp_night,
                       temp_annual, fpar):
        rad_day = sw_rad_day * (1 - sw_albedo) + lw_net_day
        rad_night = lw_net_night
        g_day, g_night = self.soil_heat_flux(
            rad_day, rad_night, temp_day, temp_night, temp_annual)
        return _soil_radiation(rad_day, rad_night, g_day, g_night, fpar)

    # :450-544
    @staticmethod
    def potential_soil_evaporation(
            pressure, temp_k, vpd, fpar, rad_soil, r_corr=None, lam=None,
            rh=None, fw=None, vpd_open=None, vpd_close=None, rbl_min=None,
            rbl_max=None):
        if lam is None:
            lam = lhv(temp_k)
        if rh is None:
            rh = rhumidity_clamped(temp_k, vpd)
        if fw is None:
            fw = f_wet_of(rh)
        if r_corr is None:
            r_corr = r_correction(pressure, temp_k)
        s = svp_slope(temp_k)
        rho = air_density(temp_k, pressure, rh)
        gamma = psychrometric_constant(pressure, temp_k)
        r_r = r_radiative(rho, temp_k)
        r_tot = r_total_soil(vpd, vpd_open, vpd_close, rbl_min, rbl_max)
        r_tot = r_tot / r_corr
        with np.errstate(all='ignore'):
            r_as = (r_tot * r_r) / (r_tot + r_r)
            numer = (s * rad_soil) + (
                rho * CP_AIR * (1 - fpar) * (vpd / r_as))
            denom = (s + gamma * (r_tot / r_as))
            return ((numer * fw) / denom, (numer * (1 - fw)) / denom)

    # :547-602
    @staticmethod
    def potential_transpiration(lw_net, sw_rad, sw_albedo, pressure, temp_k,
                                vpd, fpar, rh=None, fw=None, alpha=1.26):
        if rh is None:
            rh = rhumidity_clamped(temp_k, vpd)
        if fw is None:
            fw = f_wet_of(rh)
        rad_net = sw_rad * (1 - sw_albedo) + lw_net
        rad_canopy = fpar * rad_net
        s = svp_slope(temp_k)
        gamma = psychrometric_constant(pressure, temp_k)
        return (alpha * (s * rad_canopy) * (1 - fw)) / (s + gamma)

    # :795-864
    def evaporation_soil(self, pressure, temp_k, vpd, fpar, rad_soil,
                         r_corr=None, lam=None, rh=None, fw=None):
        if lam is None:
            lam = lhv(temp_k)
        if rh is None:
            rh = rhumidity_clamped(temp_k, vpd)
        e_sat, e_unsat = self.potential_soil_evaporation(
            pressure, temp_k, vpd, fpar, rad_soil, r_corr, lam, rh, fw,
            self.vpd_open, self.vpd_close, self.rbl_min, self.rbl_max)
        with np.errstate(all='ignore'):
            e = np.where(e_sat < 0, 0, e_sat)
            e = e + np.where(
                e_unsat < 0, 0, e_unsat * np.power(rh, vpd / self.beta))
        return e / lam

    # :866-961
    def evaporation_wet_canopy(self, pressure, temp_k, vpd, lai, fpar,
                               rad_canopy, lam=None, rh=None, fw=None,
                               tiny=1e-7):
        if lam is None:
            lam = lhv(temp_k)
        if rh is None:
            rh = rhumidity_clamped(temp_k, vpd)
        if fw is None:
            fw = f_wet_of(rh)
        fw = np.where(fw == 0, fw + tiny, fw)                     # :934
        lai = np.where(lai == 0, lai + tiny, lai)                 # :935
        s = svp_slope(temp_k)
        rho = air_density(temp_k, pressure, rh)
        with np.errstate(all='ignore'):
            r_h = 1 / (self.gl_sh * lai * fw)
            r_e = 1 / (self.gl_wv * lai * fw)
            r_r = r_radiative(rho, temp_k)
            r_aw = np.divide(r_h * r_r, r_h + r_r)
            numer = fw * ((s * rad_canopy) + (
                rho * CP_AIR * fpar * vpd * 1 / r_aw))
            denom = s + ((pressure * CP_AIR * r_e) *
                         1 / (lam * EPSILON * r_aw))
            evap = np.where(numer < 0, 0, (numer / denom) / lam)  # :959
        return np.where(np.logical_or(fw <= tiny, lai <= tiny), 0, evap)

    # :1121-1150
    def surface_conductance(self, tmin, vpd_day):
        return (self.csl * ramp_up(
            tmin - 273.15, self.tmin_close, self.tmin_open) * ramp_down(
                vpd_day, self.vpd_open, self.vpd_close))

    # :1152-1258
    def transpiration(self, pressure, temp_k, vpd, lai, fpar, rad_canopy,
                      tmin, r_corr=None, lam=None, rh=None, fw=None,
                      daytime=True, tiny=1e-7):
        if lam is None:
            lam = lhv(temp_k)
        if rh is None:
            rh = rhumidity_clamped(temp_k, vpd)
        if fw is None:
            fw = f_wet_of(rh)
        if r_corr is None:
            r_corr = r_correction(pressure, temp_k)
        s = svp_slope(temp_k)
        rho = air_density(temp_k, pressure, rh)
        gamma = psychrometric_constant(pressure, temp_k)
        r_r = r_radiative(rho, temp_k)
        g_surf = 0
        if daytime:
            g_surf = self.surface_conductance(tmin, vpd) / r_corr
        g_cut = self.g_cuticular / r_corr
        gl_sh = self.gl_sh * lai * (1 - fw)
        with np.errstate(all='ignore'):
            g = ((gl_sh * (g_surf + g_cut)) / (gl_sh + g_surf + g_cut))
        g_canopy = np.where(np.logical_and(lai > 0, (1 - fw) > 0), g, tiny)
        r_ad = (1 / self.gl_sh * r_r) / (1 / self.gl_sh + r_r)
        rad_canopy = np.where(rad_canopy < 0, 0, rad_canopy)      # :1251
        with np.errstate(all='ignore'):
            t = (1 - fw) * ((s * rad_canopy) + (
                rho * CP_AIR * fpar * (vpd / r_ad)))
            t = t / (s + gamma * (1 + (1 / g_canopy) / r_ad))
        return np.where(g_canopy <= tiny, 0, t / lam)             # :1258

    # :675-793
    def evapotranspiration(
            self, lw_net_day, lw_net_night, sw_rad_day, sw_rad_night,
            sw_albedo, temp_day, temp_night, temp_annual, tmin, vpd_day,
            vpd_night, pressure, fpar, lai, f_wet=None, separate=False):
        rad_soil = self.radiation_soil(
            lw_net_day, lw_net_night, sw_rad_day, sw_rad_night, sw_albedo,
            temp_day, temp_night, temp_annual, fpar)
        out = []
        for i in (0, 1):
            temp_k = (temp_day, temp_night)[i]
            vpd = (vpd_day, vpd_night)[i]
            sw_rad = (sw_rad_day, sw_rad_night)[i]
            lw_net = (lw_net_day, lw_net_night)[i]
            a_net = sw_rad * (1 - sw_albedo) + lw_net             # :756
            a_canopy = fpar * a_net                               # :758
            rh = rhumidity_clamped(temp_k, vpd)                   # :763
            fw = f_wet_of(rh)                                     # :764
            lam = lhv(temp_k)                                     # :768
            r_corr = r_correction(pressure, temp_k)               # :771
            e_c = self.evaporation_wet_canopy(
                pressure, temp_k, vpd, lai, fpar, a_canopy, lam, rh, fw)
            e_s = self.evaporation_soil(
                pressure, temp_k, vpd, fpar, rad_soil[i], r_corr, lam, rh, fw)
            tr = self.transpiration(
                pressure, temp_k, vpd, lai, fpar, a_canopy, tmin, r_corr,
                lam, rh, fw, daytime=(i == 0))
            if separate:
                out.append((e_c, e_s, tr))                        # :790
            else:
                out.append(e_c + e_s + tr)                        # :792
        return tuple(out)


# --------------------------------------------------------------------------
# Objective statistics used by the calibration / sensitivity callers
# --------------------------------------------------------------------------

def objective_stats(pred, obs, weights=1.0):
    '''
    Sufficient statistics of the callers' objectives (SURVEY.md section 8(c):
    `mod17.calibration.BlackBoxLikelihood` RMSD and `mod17.science.
    nash_sutcliffe` are EXTERNAL and absent -- parity for them is UNPINNED;
    they are specified by contract, not by source).  Returns
    (sum((w*(pred-obs))**2 over valid pairs), number of valid pairs), where a
    pair is valid when neither the prediction nor the observation is NaN.
    '''
    err = (np.asarray(pred, np.float64) - obs) * weights
    ok = ~np.isnan(err)
    return float(np.sum(np.square(err[ok]))), int(ok.sum())


def rmsd_from_stats(sse, n):
    return np.sqrt(sse / n) if n > 0 else np.nan


def nse_from_stats(sse, obs, norm=True):
    '1 - SSE / sum((obs-mean)^2); `norm` gives 1/(2-NSE) (sensitivity.py:96).'
    o = obs[~np.isnan(obs)]
    nse = 1 - sse / np.sum((o - o.mean())**2)
    return 1 / (2 - nse) if norm else nse


def daylength_composite(pred_day, pred_night, updown, days=8):
    '''
    Mass of water per composite period from day / night rates and day length in hours
    (tests/verification/verify2.py:112-115).
    '''
    return ((pred_day * updown * days * 60 * 60) +
            (pred_night * (24 - updown) * days * 60 * 60))


# --------------------------------------------------------------------------
# Annual sums of the calibration's precipitation constraint
# --------------------------------------------------------------------------

def annual_mass_sums(pred_le, years, lhv_arr):
    '''
    (Y, N) annual evapotranspiration [mm] from (T, N) latent heat flux [W m-2]:
    the first half of `constrain_by_map` (mod16/calibration.py:783-790).
    '''
    mass_rate = (pred_le * 60 * 60 * 24) / lhv_arr
    mass_rate[mass_rate < 0] = 0
    annual = []
    for y in np.unique(years):
        annual.append(np.apply_along_axis(lambda x: x.sum(), 0, mass_rate[years == y]))
    return np.stack(annual, axis=0)


def constrain_by_map(pred_le, years, lhv_arr, annual_precip):
    '''
    The calibration's precipitation constraint, both halves (mod16/calibration.py:775-796, a
    function nested in `CalibrationAPI.tune`): annual evapotranspiration must not exceed
    annual precipitation.  Returns the NEGATIVE normalised RMSD
    `-100 * mean(max(annual_ET - annual_precip, 0)^2) / sum(annual_precip)`.
    '''
    pred_precip = annual_mass_sums(pred_le, years, lhv_arr)       # :783-790
    diff = pred_precip - annual_precip                            # :791
    diff = np.where(diff < 0, 0, diff)                            # :792
    nrmsd = 100 * ((diff**2).mean() / annual_precip.sum())        # :795
    return -nrmsd                                                 # :796
