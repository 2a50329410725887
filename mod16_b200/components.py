'''
Host side of K4: the component methods and physics helpers of the reference's public API
(``mod16_elementwise_*`` of ``include/mod16_b200.h``), one fused elementwise launch each.

The functions keep the reference's names, argument order, optional arguments and return
conventions (reference mod16/__init__.py:121, :385-673, :795-1258, :1261-1397), so its own
unit tests (tests/tests.py:92-232) read unchanged against this package.  Arithmetic is in
double precision on the device for both float32 and float64 inputs.
'''
import ctypes
import numpy as np
import torch

from . import _abi
from .forward import (_Staged, _shape, _view_2d, hand_over, launch_stream, pick_device,
                      result_dtype)

_OP = dict((name, i) for i, name in enumerate(_abi.OPS))


def elementwise(op, args, params=None, opt=(0.0, 0.0), n_out=1, dtype=None, device=None,
                stream=None, as_torch=None):
    '''
    Evaluates component ``op`` on the broadcast of ``args`` (the positional arguments of the
    reference signature; ``None`` for an optional argument that is not supplied).
    Returns one array, or a tuple of ``n_out`` arrays (numpy unless an input is a CUDA tensor).
    '''
    lib = _abi.load()
    given = [a for a in args if a is not None]
    np_dtype = result_dtype(given, dtype)
    dev = pick_device(given, device)
    if as_torch is None:
        as_torch = any(isinstance(v, torch.Tensor) and v.device.type == 'cuda' for v in given)
    full_shape = tuple(np.broadcast_shapes(*[_shape(v) for v in given]))
    T, N = _view_2d(full_shape)
    with torch.cuda.device(dev), launch_stream(dev, stream) as (s, caller_stream):
        st = _Staged(dev, np_dtype)
        st.full_shape = full_shape
        ew = _abi.Elementwise()
        ew.op = _OP[op]
        present = 0
        for k, v in enumerate(args):
            if v is None:
                continue
            present |= 1 << k
            st.fill(ew.inp[k], v, T, N, len(full_shape))
        ew.present = present
        for k in range(_abi.N_PARAMS):
            ew.params[k] = float(params[k]) if params is not None else float('nan')
        ew.opt[0], ew.opt[1] = float(opt[0]), float(opt[1])
        outs = [torch.empty((T, N), device=dev, dtype=st.t_dtype) for _ in range(n_out)]
        for k, o in enumerate(outs):
            ew.out[k] = o.data_ptr()
        fn = lib.mod16_elementwise_f32 if np_dtype == np.dtype('float32') \
            else lib.mod16_elementwise_f64
        _abi.check(fn(ctypes.byref(ew), T, N, ctypes.c_void_p(s.cuda_stream)))
        outs = [o.reshape(full_shape) for o in outs]
        if as_torch:
            hand_over(outs, caller_stream)
            res = outs
        else:
            res = [o.cpu().numpy() for o in outs]
            if full_shape == ():
                res = [r[()] for r in res]
    return res[0] if n_out == 1 else tuple(res)


# ---- module-level functions of the reference (mod16/__init__.py) -----------------------------
def latent_heat_vaporization(temp_k, **kw):
    'Latent heat of vaporization [J kg-1] (reference :121).'
    return elementwise('lhv', [temp_k], **kw)


def psychrometric_constant(pressure, temp_k, **kw):
    'Psychrometric constant [Pa K-1] (reference :1261-1290).'
    return elementwise('psychrometric', [pressure, temp_k], **kw)


def radiation_net(sw_rad, sw_albedo, temp_k, **kw):
    'DEPRECATED in the reference (:1293-1337), kept because its unit tests call it.'
    return elementwise('radiation_net', [sw_rad, sw_albedo, temp_k], **kw)


def svp(temp_k, **kw):
    'Saturation vapour pressure [Pa] (reference :1340-1367).'
    return elementwise('svp', [temp_k], **kw)


def svp_slope(temp_k, s=None, **kw):
    'Slope of the saturation vapour pressure curve [Pa K-1] (reference :1370-1397).'
    return elementwise('svp_slope', [temp_k, s], **kw)


def daylength_composite(pred_day, pred_night, daylength_hours, days=8, **kw):
    '''
    Day and night rates [kg m-2 s-1] to the water mass of a ``days``-day composite period
    [kg m-2], weighting each by its share of the day, as the reference's product comparison
    does (tests/verification/verify2.py:112-115):
    ``day * dl * days * 3600 + night * (24 - dl) * days * 3600``.  NaN day length (polar
    day / night) propagates.  Sums over true composite periods: ``aggregate.time_aggregate``.
    '''
    return elementwise('daylength_composite', [pred_day, pred_night, daylength_hours],
                       opt=(days, 0.0), **kw)


class ComponentMixin(object):
    'The static and instance component methods of the reference ``MOD16`` class.'

    # -- static helpers ------------------------------------------------------------------
    @staticmethod
    def air_density(temp_k, pressure, rhumidity, **kw):
        'NIST simplified air density [kg m-3] (reference :385-412).'
        return elementwise('air_density', [temp_k, pressure, rhumidity], **kw)

    @staticmethod
    def air_pressure(elevation_m, **kw):
        'Atmospheric pressure [Pa] from elevation [m] (reference :415-447).'
        return elementwise('air_pressure', [elevation_m], **kw)

    @staticmethod
    def vpd(qv10m, pressure, tmean, **kw):
        'Vapour pressure deficit [Pa] from specific humidity (reference :605-644).'
        return elementwise('vpd', [qv10m, pressure, tmean], **kw)

    @staticmethod
    def rhumidity(temp_k, vpd, **kw):
        'Relative humidity clamped to [0, 1] (reference :647-673).'
        return elementwise('rhumidity', [temp_k, vpd], **kw)

    @staticmethod
    def potential_soil_evaporation(
            pressure, temp_k, vpd, fpar, rad_soil, r_corr=None, lhv=None, rhumidity=None,
            f_wet=None, vpd_open=None, vpd_close=None, rbl_min=None, rbl_max=None, **kw):
        '(evap_sat, evap_unsat) [W m-2] (reference :450-544).'
        par = [float('nan')] * _abi.N_PARAMS
        par[2], par[3], par[8], par[9] = vpd_open, vpd_close, rbl_min, rbl_max
        return elementwise('potential_soil_evaporation',
                           [pressure, temp_k, vpd, fpar, rad_soil, r_corr, lhv, rhumidity, f_wet],
                           params=par, n_out=2, **kw)

    @staticmethod
    def potential_transpiration(
            lw_net, sw_rad, sw_albedo, pressure, temp_k, vpd, fpar, rhumidity=None,
            f_wet=None, alpha=1.26, **kw):
        'Potential (daytime) transpiration [W m-2] (reference :547-602).'
        return elementwise('potential_transpiration',
                           [lw_net, sw_rad, sw_albedo, pressure, temp_k, vpd, fpar, rhumidity,
                            f_wet], opt=(alpha, 0.0), **kw)

    # -- instance methods ----------------------------------------------------------------
    def evaporation_soil(self, pressure, temp_k, vpd, fpar, rad_soil, r_corr=None, lhv=None,
                         rhumidity=None, f_wet=None, **kw):
        'Bare soil evaporation [kg m-2 s-1] (reference :795-864).'
        return elementwise('evaporation_soil',
                           [pressure, temp_k, vpd, fpar, rad_soil, r_corr, lhv, rhumidity, f_wet],
                           params=self._vector(), **kw)

    def evaporation_wet_canopy(self, pressure, temp_k, vpd, lai, fpar, rad_canopy, lhv=None,
                               rhumidity=None, f_wet=None, tiny=1e-7, **kw):
        'Wet canopy evaporation [kg m-2 s-1] (reference :866-961).'
        return elementwise('evaporation_wet_canopy',
                           [pressure, temp_k, vpd, lai, fpar, rad_canopy, lhv, rhumidity, f_wet],
                           params=self._vector(), opt=(0.0, tiny), **kw)

    def radiation_soil(self, lw_net_day, lw_net_night, sw_rad_day, sw_rad_night, sw_albedo,
                       temp_day, temp_night, temp_annual, fpar, **kw):
        '(day, night) net radiation received by the soil [W m-2] (reference :963-1053).'
        return elementwise('radiation_soil',
                           [lw_net_day, lw_net_night, sw_rad_day, sw_rad_night, sw_albedo,
                            temp_day, temp_night, temp_annual, fpar],
                           params=self._vector(), n_out=2, **kw)

    def soil_heat_flux(self, rad_net_day, rad_net_night, temp_day, temp_night, temp_annual,
                       **kw):
        '[day, night] soil heat flux [W m-2] (reference :1055-1119).'
        return list(elementwise('soil_heat_flux',
                                [rad_net_day, rad_net_night, temp_day, temp_night, temp_annual],
                                params=self._vector(), n_out=2, **kw))

    def surface_conductance(self, tmin, vpd_day, **kw):
        'Surface conductance before the r_corr correction (reference :1121-1150).'
        return elementwise('surface_conductance', [tmin, vpd_day], params=self._vector(), **kw)

    def transpiration(self, pressure, temp_k, vpd, lai, fpar, rad_canopy, tmin, r_corr=None,
                      lhv=None, rhumidity=None, f_wet=None, daytime=True, tiny=1e-7, **kw):
        'Plant transpiration [kg m-2 s-1] (reference :1152-1258).'
        return elementwise('transpiration',
                           [pressure, temp_k, vpd, lai, fpar, rad_canopy, tmin, r_corr, lhv,
                            rhumidity, f_wet],
                           params=self._vector(), opt=(1.0 if daytime else 0.0, tiny), **kw)
