'''
Drop-in host mirror of the reference's model class for the forward path.

``MOD16`` keeps the reference's constructor, ``required_parameters`` and the signatures of
``evapotranspiration()``, ``_evapotranspiration()`` and ``_et()`` (reference
mod16/__init__.py:124-201, :675-682).  Every call runs the fused CUDA kernel through the C
ABI (``include/mod16_b200.h``); there is no numpy implementation behind it.

Extensions, all keyword-only and optional, never change the reference's results:
``dtype`` (np.float32 selects the fast path; default follows numpy promotion of the
inputs), ``gate`` (see ``run_forward``), ``device`` / ``stream``.
'''
import numpy as np

from .components import ComponentMixin
from .forward import run_forward

PFT_VALID = (1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 12)


class MOD16(ComponentMixin):
    '''
    The MODIS MxD16 evapotranspiration model, B200 backend.

    Parameters
    ----------
    params : dict
        Model parameters; must hold every name in ``required_parameters`` (KeyError
        otherwise, as in the reference), extra keys are kept and ignored.
    '''
    required_parameters = [
        'tmin_close', 'tmin_open', 'vpd_open', 'vpd_close', 'gl_sh', 'gl_wv',
        'g_cuticular', 'csl', 'rbl_min', 'rbl_max', 'beta'
    ]

    def __init__(self, params):
        self.params = params
        for key in self.required_parameters:
            setattr(self, key, params[key])

    # -- static, batched interface (calibration / sensitivity callers) --------------------
    @staticmethod
    def _evapotranspiration(
            params, lw_net_day, lw_net_night, sw_rad_day, sw_rad_night,
            sw_albedo, temp_day, temp_night, temp_annual, tmin, vpd_day,
            vpd_night, pressure, fpar, lai, f_wet=None, tiny=1e-7,
            r_corr_list=None, **kw):
        '''
        Day and night latent heat flux totals [W m-2] (reference :196-382).  ``params`` is
        a sequence in ``required_parameters`` order, each a scalar or a (1 x N) array;
        drivers are scalars, (N,) or (T x N).  ``f_wet`` and ``tiny`` are accepted and
        ignored, exactly as the reference does (:191, :282).  Returns ``[day, night]``.
        '''
        return run_forward(
            (lw_net_day, lw_net_night, sw_rad_day, sw_rad_night, sw_albedo, temp_day,
             temp_night, temp_annual, tmin, vpd_day, vpd_night, pressure, fpar, lai),
            params=params, semantics='static', out_mode='day_night',
            r_corr_list=r_corr_list, **kw)

    @staticmethod
    def _et(
            params, lw_net_day, lw_net_night, sw_rad_day, sw_rad_night,
            sw_albedo, temp_day, temp_night, temp_annual, tmin, vpd_day,
            vpd_night, pressure, fpar, lai, f_wet=None, tiny=1e-7,
            r_corr_list=None, **kw):
        'Combined day + night latent heat flux [W m-2] (reference :163-193).'
        return run_forward(
            (lw_net_day, lw_net_night, sw_rad_day, sw_rad_night, sw_albedo, temp_day,
             temp_night, temp_annual, tmin, vpd_day, vpd_night, pressure, fpar, lai),
            params=params, semantics='static', out_mode='sum',
            r_corr_list=r_corr_list, **kw)[0]

    @staticmethod
    def _et_by_pft(table, pft, *drivers, **kw):
        '''
        ``_et`` with the parameters gathered on the device from the dense 11 x 13 BPLUT
        ``table`` by the per-pixel land-cover code ``pft`` -- equivalent to
        ``_et([table[k][pft].reshape(1, N) for k in range(11)], *drivers)`` without ever
        building the eleven per-pixel arrays.
        '''
        return run_forward(drivers, pft=pft, table=table, semantics='static',
                           out_mode='sum', **kw)[0]

    @staticmethod
    def _et_potential(params, *drivers, **kw):
        '''
        ``_et`` together with the potential terms the MOD16 product's PET / PLE layers are
        built from, all from ONE pass over the drivers [W m-2]: returns the dict ``et`` (day +
        night, as ``_et``), ``potential_transpiration`` (daytime, reference
        ``MOD16.potential_transpiration``, mod16/__init__.py:547-602), ``evap_sat_day`` /
        ``evap_unsat_day`` / ``evap_sat_night`` / ``evap_unsat_night`` (the two results of
        ``MOD16.potential_soil_evaporation``, :450-545, with the static path's radiation
        to the soil).  float32, (T, N) drivers.  ``params`` may be the pair ``(table, pft)``.
        '''
        if isinstance(params, tuple) and len(params) == 2:
            kw.update(table=params[0], pft=params[1])
            params = None
        kw.setdefault('semantics', 'static')
        out = run_forward(drivers, params=params, out_mode='potential', **kw)
        return dict(zip(('et', 'potential_transpiration', 'evap_sat_day', 'evap_unsat_day',
                         'evap_sat_night', 'evap_unsat_night'), out))

    @staticmethod
    def _annual_et(params, years, *drivers, **kw):
        '''
        Annual evapotranspiration [mm] per pixel of ``_et(params, *drivers)`` -- the model-side
        half of the calibration's precipitation constraint (``constrain_by_map``, reference
        mod16/calibration.py:775-790 with ``lhv = latent_heat_vaporization((temp_day +
        temp_night) / 2)``, :838-842) -- accumulated inside the fused kernel: the (T, N)
        latent heat flux is never written.  ``years`` labels every time step; rows of the
        (Y, N) float64 result follow ``np.unique(years)``.  ``params`` may also be the pair
        ``(table, pft)`` for a land-cover gather.
        '''
        import torch
        yr = np.asarray(years.cpu() if isinstance(years, torch.Tensor) else years)
        labels, idx = np.unique(yr, return_inverse=True)
        if isinstance(params, tuple) and len(params) == 2:
            kw.update(table=params[0], pft=params[1])
            params = None
        return run_forward(drivers, params=params, semantics='static', out_mode='annual',
                           groups=idx.astype(np.int32), n_groups=len(labels), **kw)[0]

    def evapotranspiration_composite(self, updown, *drivers, days=8, **kw):
        '''
        The day-length weighted composite the verification scripts form from
        ``evapotranspiration()`` (reference tests/verification/verify2.py:100-115):
        ``pred_day * updown * days * 3600 + pred_night * (24 - updown) * days * 3600``
        [kg m-2 per composite period], with ``updown`` the (T,) day length in hours, computed
        in the same pass as the model (the day and night rates are never written).
        float32, (T, N) drivers.
        '''
        u = np.asarray(updown, np.float64).reshape(-1)
        w = ((u * days * 3600.0).astype(np.float32),
             ((24.0 - u) * days * 3600.0).astype(np.float32))
        return run_forward(drivers, params=self._vector(), semantics='instance',
                           out_mode='composite', weights=w, **kw)[0]

    # -- instance interface ----------------------------------------------------------------
    def _vector(self):
        return [getattr(self, k) for k in self.required_parameters]

    def evapotranspiration(
            self, lw_net_day, lw_net_night, sw_rad_day, sw_rad_night, sw_albedo,
            temp_day, temp_night, temp_annual, tmin, vpd_day, vpd_night, pressure,
            fpar, lai, f_wet=None, separate=False, **kw):
        '''
        Day and night evapotranspiration [kg m-2 s-1] (reference :675-793).  Returns
        ``(day, night)``; with ``separate=True`` each is the tuple
        ``(e_canopy, e_soil, transpiration)``.  ``f_wet`` is ignored (:764).
        '''
        out = run_forward(
            (lw_net_day, lw_net_night, sw_rad_day, sw_rad_night, sw_albedo, temp_day,
             temp_night, temp_annual, tmin, vpd_day, vpd_night, pressure, fpar, lai),
            params=self._vector(), semantics='instance',
            out_mode='components' if separate else 'day_night', **kw)
        if separate:
            return (tuple(out[0:3]), tuple(out[3:6]))
        return tuple(out)
