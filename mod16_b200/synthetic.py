'''
Synthetic driver generators for tests and benchmarks.

Value ranges are the ones the reference's sensitivity module adopts as the [2, 98]
percentiles of its Cal-Val data (``BOUNDS``, /root/reference/mod16/sensitivity.py:31-46)
and its parameter bounds (``PARAM_BOUNDS``, :47-59).  Two generators:

* ``physical`` -- day/night/minimum temperatures are ordered, VPD is derived from a
  Beta(2, 2) relative humidity so that VPD <= SVP and about a fifth of the elements are
  wet (RH >= 0.7);
* ``stress``   -- all 14 drivers independent and uniform in ``BOUNDS``, as
  sensitivity.py:107-114 samples them (most elements then have VPD > SVP).

Host (numpy) generation is used by the parity tests; ``device_drivers`` builds the same
distributions directly in HBM with torch's generator for the benchmark configurations,
which are too large to stage through the host.
'''
from collections import OrderedDict
import numpy as np

DRIVER_NAMES = (
    'lw_net_day', 'lw_net_night', 'sw_rad_day', 'sw_rad_night', 'sw_albedo',
    'temp_day', 'temp_night', 'temp_annual', 'tmin', 'vpd_day', 'vpd_night',
    'pressure', 'fpar', 'lai')
# drivers that vary per pixel only in every reference caller (calibration.py:390)
PER_PIXEL = ('temp_annual', 'pressure')

BOUNDS = OrderedDict([
    ('lw_net_day', (-100.0, 0.0)), ('lw_net_night', (-50.0, 0.0)),
    ('sw_rad_day', (0.0, 360.0)), ('sw_rad_night', (0.0, 0.00001)),
    ('sw_albedo', (0.1, 0.22)), ('temp_day', (255.0, 305.0)),
    ('temp_night', (250.0, 300.0)), ('temp_annual', (265.0, 300.0)),
    ('tmin', (252.0, 298.0)), ('vpd_day', (10.0, 3400.0)),
    ('vpd_night', (10.0, 2400.0)), ('pressure', (70000.0, 101340.0)),
    ('fpar', (0.02, 0.89)), ('lai', (0.13, 5.34))])

PARAM_BOUNDS = OrderedDict([
    ('tmin_close', (-35.0, 0.0)), ('tmin_open', (0.0, 25.0)),
    ('vpd_open', (0.0, 1000.0)), ('vpd_close', (1000.0, 8000.0)),
    ('gl_sh', (0.001, 0.2)), ('gl_wv', (0.001, 0.2)),
    ('g_cuticular', (1e-7, 1e-3)), ('csl', (0.0001, 0.2)),
    ('rbl_min', (10.0, 1000.0)), ('rbl_max', (100.0, 2000.0)),
    ('beta', (0.0, 2000.0))])


def _svp_host(temp_k):
    tc = temp_k - 273.15
    return 610.8 * np.exp(17.27 * tc / (tc + 237.3))


def host_drivers(T, N, seed, kind='physical', dtype=np.float64,
                 nan_frac=0.0, fill_frac=0.0, sw_rad_night='zero'):
    '''
    Returns an OrderedDict of the 14 drivers in call order.  Per-step drivers are (T, N);
    ``temp_annual`` and ``pressure`` are (N,); ``sw_rad_night`` is the python scalar 0
    (as in every reference caller) unless ``sw_rad_night='array'``.
    '''
    rng = np.random.default_rng(seed)
    u = lambda name, shape: rng.uniform(*BOUNDS[name], size=shape)
    d = OrderedDict()
    if kind == 'physical':
        t_day = u('temp_day', (T, N))
        t_night = t_day - rng.uniform(0, 15, (T, N))
        tmin = t_night - rng.uniform(0, 5, (T, N))
        vpd_day = _svp_host(t_day) * (1 - rng.beta(2, 2, (T, N)))
        vpd_night = _svp_host(t_night) * (1 - rng.beta(2, 2, (T, N)))
    elif kind == 'stress':
        t_day = u('temp_day', (T, N))
        t_night = u('temp_night', (T, N))
        tmin = u('tmin', (T, N))
        vpd_day = u('vpd_day', (T, N))
        vpd_night = u('vpd_night', (T, N))
    else:
        raise ValueError(kind)
    d['lw_net_day'] = u('lw_net_day', (T, N))
    d['lw_net_night'] = u('lw_net_night', (T, N))
    d['sw_rad_day'] = u('sw_rad_day', (T, N))
    d['sw_rad_night'] = 0 if sw_rad_night == 'zero' else u('sw_rad_night', (T, N))
    d['sw_albedo'] = u('sw_albedo', (T, N))
    d['temp_day'] = t_day
    d['temp_night'] = t_night
    d['temp_annual'] = u('temp_annual', (N,))
    d['tmin'] = tmin
    d['vpd_day'] = vpd_day
    d['vpd_night'] = vpd_night
    d['pressure'] = u('pressure', (N,))
    d['fpar'] = u('fpar', (T, N))
    d['lai'] = u('lai', (T, N))
    for k, v in d.items():
        if not isinstance(v, np.ndarray):
            continue
        v = v.astype(dtype)
        if nan_frac > 0:
            v[rng.random(v.shape) < nan_frac] = np.nan
        if fill_frac > 0:
            v[rng.random(v.shape) < fill_frac] = -9999
        d[k] = v
    return d


def host_pft_map(N, seed, valid=(1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 12)):
    'uint8[N] land-cover codes, uniform over the valid PFT codes.'
    rng = np.random.default_rng(seed)
    return np.asarray(valid, np.uint8)[rng.integers(0, len(valid), N)]


def host_param_draws(D, seed, bplut_column=None, fixed=('tmin_close', 'tmin_open', 'vpd_open')):
    '''
    (D, 11) float64 parameter draws, uniform in ``PARAM_BOUNDS``; the names in ``fixed``
    take the BPLUT value of ``bplut_column`` (an 11-vector in ``required_parameters``
    order) as the calibration does (/root/reference/mod16/calibration.py:233-235).
    '''
    rng = np.random.default_rng(seed)
    names = list(PARAM_BOUNDS.keys())
    out = np.empty((D, 11), np.float64)
    for j, k in enumerate(names):
        out[:, j] = rng.uniform(*PARAM_BOUNDS[k], size=D)
        if bplut_column is not None and k in fixed:
            out[:, j] = bplut_column[j]
    return out


def device_drivers(T, N, seed, device, kind='physical', dtype=None):
    '''
    The ``physical`` / ``stress`` distributions generated in HBM (torch CUDA generator).
    Returns an OrderedDict like ``host_drivers`` holding torch tensors (and the scalar 0
    for ``sw_rad_night``).
    '''
    import torch
    dtype = dtype or torch.float32
    g = torch.Generator(device=device)
    g.manual_seed(seed)

    def u(name, shape):
        lo, hi = BOUNDS[name]
        return torch.empty(shape, device=device, dtype=dtype).uniform_(lo, hi, generator=g)

    def rh_like(shape):
        # Beta(2, 2) as the median of three uniforms
        a, b, c = (torch.empty(tuple(shape), device=device, dtype=dtype).uniform_(0, 1, generator=g)
                   for _ in range(3))
        return torch.maximum(torch.minimum(a, b), torch.minimum(torch.maximum(a, b), c))

    def svp_t(t):
        tc = t - 273.15
        return 610.8 * torch.exp(17.27 * tc / (tc + 237.3))

    d = OrderedDict()
    t_day = u('temp_day', (T, N))
    if kind == 'physical':
        t_night = t_day - torch.empty_like(t_day).uniform_(0, 15, generator=g)
        tmin = t_night - torch.empty_like(t_day).uniform_(0, 5, generator=g)
        vpd_day = svp_t(t_day) * (1 - rh_like((T, N)))
        vpd_night = svp_t(t_night) * (1 - rh_like((T, N)))
    else:
        t_night = u('temp_night', (T, N))
        tmin = u('tmin', (T, N))
        vpd_day = u('vpd_day', (T, N))
        vpd_night = u('vpd_night', (T, N))
    d['lw_net_day'] = u('lw_net_day', (T, N))
    d['lw_net_night'] = u('lw_net_night', (T, N))
    d['sw_rad_day'] = u('sw_rad_day', (T, N))
    d['sw_rad_night'] = 0
    d['sw_albedo'] = u('sw_albedo', (T, N))
    d['temp_day'] = t_day
    d['temp_night'] = t_night
    d['temp_annual'] = u('temp_annual', (N,))
    d['tmin'] = tmin
    d['vpd_day'] = vpd_day
    d['vpd_night'] = vpd_night
    d['pressure'] = u('pressure', (N,))
    d['fpar'] = u('fpar', (T, N))
    d['lai'] = u('lai', (T, N))
    return d
