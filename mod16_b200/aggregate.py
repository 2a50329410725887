'''
Host side of K3: per-pixel time aggregates of a (T, N) field on the GPU
(``mod16_time_aggregate_*`` of ``include/mod16_b200.h``).

``annual_mass_sums`` is the reduction the calibration's precipitation constraint performs on
the model output (reference mod16/calibration.py:783-790): latent heat flux [W m-2] to
[mm day-1] with the latent heat of vaporisation, negative rates dropped, summed per year.
'''
import ctypes
import numpy as np
import torch

from . import _abi
from .forward import hand_over, launch_stream, pick_device

_TORCH_DT = {np.dtype('float32'): torch.float32, np.dtype('float64'): torch.float64}


def _dev(x, dev, dtype):
    if isinstance(x, torch.Tensor):
        return x.detach().to(device=dev, dtype=dtype).contiguous()
    return torch.from_numpy(np.ascontiguousarray(x)).to(device=dev, dtype=dtype)


def time_aggregate(x, group=None, n_groups=None, divisor=None, scale=1.0, clamp_neg=False,
                   skip_nan=False, total=False, device=None, stream=None):
    '''
    out[g, n] = sum over the steps t with group[t] == g of  x[t, n] * scale / divisor[t, n].

    x, divisor : (T, N) numpy arrays or CUDA tensors, float32 or float64 (both the same)
    group      : (T,) integer group index per step (steps outside [0, n_groups) are skipped);
                 None: one group
    Returns ``out`` (n_groups, N) float64, or ``(out, total)`` with total[g] = out[g].sum()
    when ``total`` is set.  numpy in, numpy out; CUDA tensors in, CUDA tensors out.
    '''
    lib = _abi.load()
    as_torch = isinstance(x, torch.Tensor) and x.device.type == 'cuda'
    dev = pick_device([x], device)
    np_dt = np.dtype(str(x.dtype).replace('torch.', '')) if isinstance(x, torch.Tensor) \
        else np.asarray(x).dtype
    if np_dt not in _TORCH_DT:
        np_dt = np.dtype('float64')
    if np.ndim(x) != 2:
        raise ValueError('x must be (T, N)')
    with torch.cuda.device(dev), launch_stream(dev, stream) as (s, caller_stream):
        xt = _dev(x, dev, _TORCH_DT[np_dt])
        T, N = xt.shape
        keep = [xt]
        ag = _abi.Aggregate()
        ag.x = xt.data_ptr()
        if divisor is not None:
            dt = _dev(divisor, dev, _TORCH_DT[np_dt])
            if tuple(dt.shape) != (T, N):
                raise ValueError('divisor must have the shape of x')
            keep.append(dt)
            ag.divisor = dt.data_ptr()
        ag.scale = float(scale)
        if group is not None:
            gt = _dev(np.asarray(group.cpu() if isinstance(group, torch.Tensor) else group),
                      dev, torch.int32)
            if gt.numel() != T:
                raise ValueError('group must hold one index per time step')
            keep.append(gt)
            ag.group = gt.data_ptr()
            if n_groups is None:
                n_groups = int(gt.max().item()) + 1 if T else 1
        ag.n_groups = int(n_groups or 1)
        out = torch.empty((ag.n_groups, N), device=dev, dtype=torch.float64)
        ag.out = out.data_ptr()
        tot = None
        if total:
            tot = torch.empty((ag.n_groups,), device=dev, dtype=torch.float64)
            ag.total = tot.data_ptr()
        flags = (_abi.AGG_CLAMP_NEG if clamp_neg else 0) | (_abi.AGG_SKIP_NAN if skip_nan else 0)
        fn = lib.mod16_time_aggregate_f32 if np_dt == np.dtype('float32') \
            else lib.mod16_time_aggregate_f64
        _abi.check(fn(ctypes.byref(ag), T, N, flags, ctypes.c_void_p(s.cuda_stream)))
        if as_torch:
            hand_over((out, tot), caller_stream)
            return (out, tot) if total else out
        res = out.cpu().numpy()
        return (res, tot.cpu().numpy()) if total else res


def annual_mass_sums(pred_le, years, lhv, **kw):
    '''
    Annual evapotranspiration [mm] per pixel from latent heat flux [W m-2]: the first half of
    ``constrain_by_map`` (reference mod16/calibration.py:783-790).  ``years`` is the (T,)
    array of year labels; rows of the result follow ``np.unique(years)``.
    '''
    yr = np.asarray(years.cpu() if isinstance(years, torch.Tensor) else years)
    labels, idx = np.unique(yr, return_inverse=True)
    return time_aggregate(pred_le, group=idx.astype(np.int32), n_groups=len(labels),
                          divisor=lhv, scale=60 * 60 * 24, clamp_neg=True, **kw)


def constrain_by_map(pred_le, years, lhv, annual_precip, **kw):
    '''
    The calibration's precipitation constraint (reference mod16/calibration.py:775-796):
    annual evapotranspiration [mm] per pixel from latent heat flux (K3, ``annual_mass_sums``)
    against (Y, N) annual precipitation; returns the NEGATIVE normalised RMSD
    ``-100 * mean(max(annual_ET - annual_precip, 0) ** 2) / sum(annual_precip)`` -- the
    objective the configuration calls ``nRMSD`` (data/MOD16_calibration_config.yaml:49).
    A python float, or a 0-d CUDA tensor when ``pred_le`` is a CUDA tensor.
    '''
    as_torch = isinstance(pred_le, torch.Tensor) and pred_le.device.type == 'cuda'
    dev = pick_device([pred_le], kw.get('device'))
    le = pred_le if as_torch else _dev(pred_le, dev, _TORCH_DT.get(
        np.asarray(pred_le).dtype, torch.float64))
    lh = lhv if isinstance(lhv, torch.Tensor) else _dev(lhv, dev, le.dtype)
    annual = annual_mass_sums(le, years, lh.to(le.dtype), **kw)       # (Y, N) float64, on device
    precip = _dev(annual_precip, dev, torch.float64)
    diff = torch.clamp(annual - precip, min=0.0)                      # :791-792
    val = -100.0 * ((diff * diff).mean() / precip.sum())              # :795-796
    return val if as_torch else float(val.item())


def constrain_by_map_fused(params, drivers, years, annual_precip, **kw):
    '''
    ``constrain_by_map(MOD16._et(params, *drivers), years, lhv, annual_precip)`` with ``lhv``
    from the mean of day and night temperature (reference mod16/calibration.py:775-796,
    :838-849) in ONE kernel pass: the annual sums are an epilogue of the forward model
    (``MOD16._annual_et``), only the (Y, N) remainder runs as torch operations.
    '''
    from .model import MOD16
    annual = MOD16._annual_et(params, years, *drivers, as_torch=True, **kw)
    precip = _dev(annual_precip, annual.device, torch.float64)
    diff = torch.clamp(annual - precip, min=0.0)
    return float((-100.0 * ((diff * diff).mean() / precip.sum())).item())
