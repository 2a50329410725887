'''
Host side of K1: turns the reference's loosely typed call arguments (python scalars,
0-d / (N,) / (1, N) / (T, N) numpy arrays or torch tensors) into the broadcast descriptors
of ``include/mod16_b200.h`` and launches the fused CUDA forward kernel.

The broadcasting contract mirrors numpy's for the shapes the reference's callers use
(SURVEY.md H6): the result has the broadcast shape of all drivers and parameters; an
input that is constant along time or along pixels is never materialised -- it is passed
with a zero stride.  Inputs already resident on the GPU (torch CUDA tensors) are used in
place; host inputs are copied to the device for the call and results are copied back.
'''
import contextlib
import ctypes
import numpy as np
import torch

from . import _abi, hostcopy
from .rawio import RawDriver, DTYPE_CODE, dequantize, prep_word

_TORCH_DT = {np.dtype('float32'): torch.float32, np.dtype('float64'): torch.float64}


def _is_scalar(x):
    if isinstance(x, (int, float, np.integer, np.floating, bool)):
        return True
    if isinstance(x, np.ndarray) and x.ndim == 0:
        return True
    if isinstance(x, torch.Tensor) and x.ndim == 0 and x.device.type == 'cpu':
        return True
    return False


def _shape(x):
    if _is_scalar(x):
        return ()
    return tuple(x.shape) if hasattr(x, 'shape') else tuple(np.shape(x))


def _arrayish(x):
    'python sequences (lists, tuples) and array-likes without .shape become numpy arrays.'
    if _is_scalar(x) or isinstance(x, (np.ndarray, torch.Tensor, RawDriver)):
        return x
    return np.asarray(x)


def result_dtype(values, dtype=None):
    '''
    float32 only when numpy's own promotion of the inputs would give float32 (python
    scalars are weak, NEP 50); otherwise float64.  ``dtype`` overrides.
    '''
    if dtype is not None:
        dt = np.dtype(dtype)
        if dt not in _TORCH_DT:
            raise TypeError('dtype must be float32 or float64, not %s' % dt)
        return dt
    kinds = []
    for v in values:
        if isinstance(v, (int, float, bool)):
            continue  # weak python scalar
        if isinstance(v, RawDriver):
            kinds.append(np.dtype('float32'))
        elif isinstance(v, torch.Tensor):
            kinds.append(np.dtype(str(v.dtype).replace('torch.', '')))
        elif isinstance(v, (np.ndarray, np.generic)):
            kinds.append(v.dtype)
        else:
            kinds.append(np.asarray(v).dtype)
    if not kinds:
        return np.dtype('float64')
    dt = np.result_type(*kinds)
    if dt == np.dtype('float32') or dt == np.dtype('float16'):
        return np.dtype('float32')
    return np.dtype('float64')


def pick_device(values, device=None):
    if device is not None:
        return torch.device(device)
    for v in values:
        if isinstance(v, RawDriver):
            v = v.raw
        if isinstance(v, torch.Tensor) and v.device.type == 'cuda':
            return v.device
    if not torch.cuda.is_available():
        raise _abi.Mod16Error(
            'mod16_b200 needs a CUDA device (B200, sm_100a); there is no CPU fallback')
    return torch.device('cuda', torch.cuda.current_device())


@contextlib.contextmanager
def launch_stream(dev, stream):
    '''
    Makes ``stream`` (None: the current stream; an int: a raw cudaStream_t) the current stream
    for the body, after it has waited for the work already queued on the caller's current
    stream.  Staging copies, dtype conversions, scratch allocation, the launches and the
    read-back of a call are all issued inside, so they are ordered with the kernel (ADVICE r1:
    with a foreign stream the kernel could otherwise start before its inputs had landed, and
    a read-back could overtake it).  Yields ``(s, caller_stream)``; ``caller_stream`` is None
    unless ``s`` is foreign.
    '''
    cur = torch.cuda.current_stream(dev)
    if stream is None:
        s = cur
    elif isinstance(stream, int):
        s = torch.cuda.ExternalStream(stream, device=dev)
    else:
        s = stream
    if s != cur:
        s.wait_stream(cur)
    with torch.cuda.stream(s):
        yield s, (cur if s != cur else None)


def hand_over(tensors, caller_stream):
    'Results of a call on a foreign stream will be read on the caller\'s: tell the allocator.'
    if caller_stream is not None:
        for t in tensors:
            if t is not None:
                t.record_stream(caller_stream)


class _Staged(object):
    'Keeps device copies alive for the duration of a call and fills Driver structs.'

    def __init__(self, device, dtype):
        self.device = device
        self.np_dtype = dtype
        self.t_dtype = _TORCH_DT[dtype]
        self.keep = []

    def tensor(self, x, dtype=None):
        dtype = dtype or self.t_dtype
        if isinstance(x, torch.Tensor):
            t = x.detach()
            if t.device != self.device or t.dtype != dtype:
                t = t.to(device=self.device, dtype=dtype, non_blocking=True)
        else:
            a = np.asarray(x)
            t = torch.from_numpy(np.ascontiguousarray(a)).to(
                device=self.device, dtype=dtype, non_blocking=True)
        self.keep.append(t)
        return t

    def fill(self, slot, x, T, N, ndim, keep_dtype=False):
        '''
        Describes input ``x`` (broadcastable to the call's (T, N) view) in ``slot``.
        ``keep_dtype``: stage the array in its own element type (storage formats).
        '''
        if _is_scalar(x):
            slot.ptr = None
            slot.t_stride = 0
            slot.n_stride = 0
            slot.scalar = float(x)
            return
        if keep_dtype:
            own = x.dtype if isinstance(x, torch.Tensor) else \
                getattr(torch, np.asarray(x).dtype.name)
            t = self.tensor(x, own)
        else:
            t = self.tensor(x)
        shp = tuple(t.shape)
        if len(self.full_shape) > 2 and len(shp) >= 2:
            # more than two axes: leading axes fold into time; materialise the broadcast
            t = torch.broadcast_to(t, self.full_shape).reshape(T, N)
            self.keep.append(t)
            shp = tuple(t.shape)
        if len(shp) == 1:
            shp = (1,) + shp
            strides = (0, t.stride(0))
        else:
            strides = (t.stride(0), t.stride(1))
        if any(s < 0 for s in strides):
            t = t.contiguous()
            self.keep.append(t)
            strides = (t.stride(0), t.stride(1)) if t.ndim == 2 else (0, t.stride(0))
        ts = 0 if shp[0] == 1 else strides[0]
        ns = 0 if shp[1] == 1 else strides[1]
        if shp[0] not in (1, T) or shp[1] not in (1, N):
            raise ValueError('input of shape %r does not broadcast to (%d, %d)' % (
                tuple(x.shape), T, N))
        slot.ptr = t.data_ptr()
        slot.t_stride = ts
        slot.n_stride = ns
        slot.scalar = 0.0


_WORKSPACES = {}          # (device index, stream handle) -> int64 scratch tensor
_WS_MAX_BYTES = 256 << 20   # per stream; a full queue only costs speed (marker + sweep)
_WS_MAX_STREAMS = 8         # least recently used entries beyond this are dropped


def _workspace(dev, stream_handle, lib, units=0):
    '''
    Per-device, per-stream scratch of the float32 path (the queue of elements re-evaluated in
    double precision).  Cached: calls on one stream are ordered, so they can share it.  Sized
    for 1/64 of the call's elements (the measured rates are 5e-5 on physical and 1e-3 on
    all-uniform drivers; unmasked fill values raise it), at least the library's
    recommendation, at most 256 MiB; a full queue only costs speed, never results.  At most
    eight streams keep a buffer (least recently used first out); ``release_workspaces()``
    frees them all.
    '''
    key = (dev.index, stream_handle)
    want = max(lib.mod16_workspace_bytes() // 8, min(units // 64 + 1024, _WS_MAX_BYTES // 8))
    ws = _WORKSPACES.pop(key, None)
    if ws is None or ws.numel() < want:
        ws = torch.empty(want, device=dev, dtype=torch.int64)
    _WORKSPACES[key] = ws          # (re-inserted: dicts keep insertion order = LRU order)
    while len(_WORKSPACES) > _WS_MAX_STREAMS:
        _WORKSPACES.pop(next(iter(_WORKSPACES)))
    return ws


def release_workspaces():
    'Drops every cached scratch buffer (they are re-created on demand).'
    _WORKSPACES.clear()


def _view_2d(shape):
    'Maps a broadcast result shape onto the kernels\' (T, N) view.'
    if len(shape) == 0:
        return 1, 1
    if len(shape) == 1:
        return 1, shape[0]
    T = 1
    for s in shape[:-1]:
        T *= s
    return T, shape[-1]


def run_forward(drivers, *, params=None, pft=None, table=None, semantics='static',
                out_mode='sum', gate='global', dtype=None, r_corr_list=None, device=None,
                stream=None, force=0, as_torch=None, gate_flag=None, out=None,
                workspace=None, prep=None, groups=None, n_groups=None, weights=None,
                surface_pressure=None, out_scale=None):
    '''
    Launches K1 once.

    drivers : the 14 positional drivers of the reference signatures
    params  : sequence of 11 values in ``required_parameters`` order, each a scalar or a
              per-pixel array ((N,), (1, N)); or None when ``pft``/``table`` are given
    pft, table : uint8 land-cover code per pixel and the dense 11 x 13 BPLUT table
    out_mode : 'sum' | 'day_night' | 'components' | 'annual' (float32 static path: no
               per-step output, one (n_groups, N) float64 array of evapotranspiration sums
               [mm] per pixel and group of time steps -- ``groups`` holds the group (year)
               index of every step; the reduction of the calibration's precipitation
               constraint, mod16/calibration.py:783-790, fused into the kernel)
               | 'potential' (float32, (T, N) drivers: the total and the potential terms
               [W m-2] -- total, potential transpiration (day, :547-602), evaporation of
               the saturated / potential evaporation of the unsaturated soil fraction
               (day), the same two at night (:450-545) -- from the same pass)
               | 'composite' (float32, (T, N) drivers: ``day * weights[0][t] + night *
               weights[1][t]``, the day-length weighted composite of verify2.py:112-115;
               ``weights`` = two (T,) arrays)
               | 'sum_i16' (float32, (T, N) drivers: the total as int16 counts of
               ``out_scale``, -32768 where it is NaN -- the way the MOD16 product stores its
               layers; returns a ``rawio.RawDriver`` whose ``physical()`` gives the floats)
    gate     : 'global' (reference) | 'element' | 'open' | 'closed' | 'open_track' | 'fixup'

    drivers may hold ``rawio.RawDriver`` objects (storage formats: uint8 / int16 with scale,
    offset and fill value); ``prep`` names loader conversions applied on the device:
    'vpd_night_clamp' (calibration.py:401), 'elevation' (the ``pressure`` driver holds
    elevation in metres, calibration.py:405-408).  ``surface_pressure=(ps_day, ps_night)``:
    the ``vpd_day`` / ``vpd_night`` drivers hold the water vapour mixing ratio QV10M and the
    kernel forms ``VPD = MOD16.vpd(qv10m, ps, temp)`` itself (mod16/__init__.py:605-644, the
    loaders' other way to the VPD, calibration.py:393-400).  All select the float32 path.

    out      : optional list of preallocated CUDA tensors ((T, N), contiguous) to write into
    workspace : optional int64 CUDA tensor, scratch of the float32 path (default: a cached
               per-device, per-stream buffer of ``mod16_workspace_bytes()``)

    Returns a list of 1, 2 or 6 arrays of the broadcast shape (numpy, unless an input was
    a CUDA tensor or ``as_torch`` is set).
    '''
    lib = _abi.load()
    drivers = [_arrayish(v) for v in drivers]
    if len(drivers) != _abi.N_DRIVERS:
        raise TypeError('expected 14 drivers, got %d' % len(drivers))
    extra = [_arrayish(v) for v in r_corr_list] if r_corr_list is not None else []
    plist = [_arrayish(v) for v in params] if params is not None else []
    if params is not None and len(plist) != _abi.N_PARAMS:
        raise TypeError('expected 11 parameters, got %d' % len(plist))
    everything = drivers + extra + plist
    np_dtype = result_dtype(everything, dtype)
    dev = pick_device(everything + ([pft] if pft is not None else []), device)
    if as_torch is None:
        as_torch = any((isinstance(v, torch.Tensor) and v.device.type == 'cuda') or
                       (isinstance(v, RawDriver) and v.is_cuda()) for v in everything)
    prep = prep_word(prep)
    if surface_pressure is not None:
        if len(surface_pressure) != 2:
            raise TypeError('surface_pressure = (ps_day, ps_night)')
        prep |= _abi.PREP_VPD_FROM_QV
        surface_pressure = [_arrayish(v) for v in surface_pressure]
    raw = prep != 0 or any(isinstance(v, RawDriver) for v in drivers)
    if raw and (np_dtype != np.dtype('float32') or r_corr_list is not None or force):
        # storage formats belong to the float32 kernels: other call forms see plain arrays
        drivers, prep, raw = _plain_drivers(drivers, prep, dev, np_dtype, surface_pressure), 0, False

    shapes = [_shape(v) for v in everything]
    if pft is not None:
        shapes.append(tuple(np.shape(pft)) if not isinstance(pft, torch.Tensor)
                      else tuple(pft.shape))
    full_shape = tuple(np.broadcast_shapes(*shapes))
    T, N = _view_2d(full_shape)

    if out_mode == 'annual':
        if groups is None:
            groups = np.zeros(T, np.int32)
        if np_dtype != np.dtype('float32') or semantics != 'static' or len(full_shape) != 2:
            raise _abi.LayoutError('out_mode="annual" belongs to the float32 static path on '
                                   '(T, N) drivers')
    with torch.cuda.device(dev), launch_stream(dev, stream) as (s, caller_stream):
        if out_mode == 'annual':
            try:
                return _run_annual(lib, s, dev, np_dtype, drivers, plist, pft, table, gate,
                                   as_torch, full_shape, T, N, caller_stream,
                                   prep if raw else None, groups, n_groups, surface_pressure)
            except _abi.LayoutError:
                if not raw:
                    raise
                drivers = _plain_drivers(drivers, prep, dev, np_dtype, surface_pressure)
                return _run_annual(lib, s, dev, np_dtype, drivers, plist, pft, table, gate,
                                   as_torch, full_shape, T, N, caller_stream, None, groups,
                                   n_groups)
        try:
            return _run_on_stream(lib, s, dev, np_dtype, drivers, extra, plist, pft, table,
                                  semantics, out_mode, gate, r_corr_list, force, as_torch,
                                  gate_flag, out, workspace, full_shape, T, N, caller_stream,
                                  prep if raw else None, weights=weights, ps=surface_pressure,
                                  out_scale=out_scale)
        except _abi.LayoutError:
            if not raw:          # plain float32 drivers: the mode itself needs a streamable layout
                raise
            # a storage-format layout the fused kernel cannot stream: same arithmetic as
            # torch operations on the device first, then the plain float32 entry
            drivers = _plain_drivers(drivers, prep, dev, np_dtype, surface_pressure)
            return _run_on_stream(lib, s, dev, np_dtype, drivers, extra, plist, pft, table,
                                  semantics, out_mode, gate, r_corr_list, force, as_torch,
                                  gate_flag, out, workspace, full_shape, T, N, caller_stream,
                                  None, weights=weights, out_scale=out_scale)


def _run_annual(lib, s, dev, np_dtype, drivers, plist, pft, table, gate, as_torch, full_shape,
                T, N, caller_stream, prep, groups, n_groups, ps=None):
    '''
    out_mode='annual': integer accumulators on the device (2^-32 mm, order-independent), the
    double-precision queue sized for an eighth of the call (no per-step output exists that
    could carry overflow markers), the result as float64 millimetres (NaN where a step of
    the pixel's group was NaN, like the reference's ``x.sum()``).
    '''
    g = torch.as_tensor(np.asarray(groups.cpu() if isinstance(groups, torch.Tensor) else groups))
    if g.numel() != T:
        raise ValueError('groups must hold one index per time step')
    g = g.to(device=dev, dtype=torch.int32)
    G = int(n_groups) if n_groups is not None else (int(g.max().item()) + 1 if T else 1)
    acc = torch.zeros((G, N), device=dev, dtype=torch.int64)
    ws = torch.empty(max(1 << 20, T * N // 8 + 1024), device=dev, dtype=torch.int64)
    flag = torch.zeros(1, device=dev, dtype=torch.int32)
    _run_on_stream(lib, s, dev, np_dtype, drivers, [], plist, pft, table, 'static', 'annual',
                   gate, None, 0, True, flag, [acc], ws, full_shape, T, N, caller_stream, prep,
                   groups=(g, G), ps=ps)
    if int(flag.item()) & _abi.ANNUAL_OVERFLOW:
        raise _abi.Mod16Error(
            'out_mode="annual": more than an eighth of the elements need the double-precision '
            'pass (unmasked fill values?); mask them, or use out_mode="sum" with '
            'aggregate.annual_mass_sums')
    out = acc.to(torch.float64) * (1.0 / 4294967296.0)
    out = torch.where(acc >= (1 << 54), torch.full_like(out, float('nan')), out)
    hand_over((out,), caller_stream)
    return [out if as_torch else out.cpu().numpy()]


def _plain_drivers(drivers, prep, dev, np_dtype, ps=None):
    'Storage-format drivers and loader conversions as plain device tensors of `np_dtype`.'
    t_dtype = _TORCH_DT[np_dtype]
    out = []
    if prep & _abi.PREP_VPD_FROM_QV:          # MOD16.vpd (:639-644), in double
        drivers = list(drivers)
        for h, (k_vpd, k_t) in enumerate(((9, 5), (10, 6))):
            qv, p_sfc, temp = (torch.as_tensor(v.physical() if isinstance(v, RawDriver) else v)
                               .to(dev).double() for v in (drivers[k_vpd], ps[h], drivers[k_t]))
            tc = temp - 273.15
            drivers[k_vpd] = (610.7 * torch.exp((17.38 * tc) / (239 + tc)) -
                              (qv * p_sfc) / (0.622 + (0.379 * qv))).to(t_dtype)
    for k, v in enumerate(drivers):
        clamp = k == 10 and (prep & _abi.PREP_VPD_NIGHT_CLAMP)
        if isinstance(v, RawDriver):
            v = dequantize(v, dev, prep_clamp=clamp).to(t_dtype)
        elif clamp:
            v = torch.as_tensor(v).to(dev)
            v = torch.where(v < 0, torch.zeros_like(v), v)
        if k == 11 and (prep & _abi.PREP_ELEVATION):       # MOD16.air_pressure (:415-447)
            z = torch.as_tensor(v).to(dev).double()
            v = (101325.0 * torch.pow(1.0 - (0.0065 * z) / 288.15, 5.256)).to(t_dtype)
        out.append(v)
    return out


def _run_on_stream(lib, s, dev, np_dtype, drivers, extra, plist, pft, table, semantics,
                   out_mode, gate, r_corr_list, force, as_torch, gate_flag, out, workspace,
                   full_shape, T, N, caller_stream, prep=None, groups=None, weights=None,
                   ps=None, out_scale=None):
    'Body of run_forward; the current stream is `s`.  `prep` not None: the storage-format entry.'
    params = plist if plist else None
    st = _Staged(dev, np_dtype)
    st.full_shape = full_shape
    # large pageable numpy drivers: chunked, threaded staging through pinned buffers
    # (a float64 array headed for the float32 path is converted on the device)
    big = [k for k, v in enumerate(drivers) if hostcopy.worthwhile(v)]
    if big:
        for k, t in zip(big, hostcopy.upload([drivers[k] for k in big], dev, s)):
            # (a float32 array headed for the float64 entry stays float32: widened in the kernel)
            if t.dtype != st.t_dtype and not (t.dtype == torch.float32 and
                                              st.t_dtype == torch.float64 and prep is None):
                t = t.to(st.t_dtype)
            drivers[k] = t
            st.keep.append(t)
    if prep is not None:
        cd = _abi.RawDrivers()
        cd.prep = prep
        for k, v in enumerate(drivers):
            slot = cd.drv[k]
            slot.dtype, slot.has_fill, slot.scale, slot.offset, slot.fill = _abi.T_F32, 0, 1.0, 0.0, 0
            if isinstance(v, RawDriver):
                st.fill(slot, v.raw, T, N, len(full_shape), keep_dtype=True)
                slot.dtype = DTYPE_CODE[v.np_dtype]
                slot.scale, slot.offset = v.scale, v.offset
                if v.fill is not None and slot.dtype != _abi.T_F32:
                    slot.has_fill, slot.fill = 1, int(v.fill)
            else:
                st.fill(slot, v, T, N, len(full_shape))
        if prep & _abi.PREP_VPD_FROM_QV:
            for h in range(2):
                slot = cd.ps[h]
                slot.dtype, slot.has_fill, slot.scale, slot.offset, slot.fill = _abi.T_F32, 0, 1.0, 0.0, 0
                st.fill(slot, ps[h], T, N, len(full_shape))
    else:
        cd = _abi.Drivers()
        mask = 0
        for k, v in enumerate(drivers):
            # float64 entry: a float32 driver stays float32 (widened exactly in the kernel) --
            # the reference's stock call is float32 drivers + the float64 BPLUT
            is32 = (np_dtype == np.dtype('float64') and not _is_scalar(v) and
                    str(getattr(v, 'dtype', '')).replace('torch.', '') == 'float32')
            st.fill(cd.drv[k], v, T, N, len(full_shape), keep_dtype=is32)
            if is32 and cd.drv[k].ptr:
                mask |= 1 << k
        cd.f32_mask = mask
        cd.has_r_corr = 1 if r_corr_list is not None else 0
        for k, v in enumerate(extra[:2]):
            st.fill(cd.r_corr[k], v, T, N, len(full_shape))

    cp = _abi.Params()
    if pft is not None:
        if table is None:
            raise TypeError('`pft` needs the 11 x 13 `table`')
        tab = np.asarray(table, np.float64)
        if tab.shape != (_abi.N_PARAMS, _abi.N_PFT):
            raise ValueError('table must be 11 x 13')
        pt = st.tensor(pft, torch.uint8).reshape(-1)
        if pt.numel() != N:
            raise ValueError('pft must hold one code per pixel (N = %d)' % N)
        pt = pt.contiguous()
        st.keep.append(pt)
        cp.pft = pt.data_ptr()
    else:
        tab = np.full((_abi.N_PARAMS, _abi.N_PFT), np.nan)
        if all(_is_scalar(v) for v in plist):
            tab[:, 0] = [float(v) for v in plist]
        else:
            for k, v in enumerate(plist):
                if _is_scalar(v):
                    v = np.full((1,), float(v))
                t = st.tensor(v).reshape(-1)
                if t.numel() not in (1, N):
                    raise ValueError(
                        'parameter %d has %d values, expected 1 or N = %d' % (
                            k, t.numel(), N))
                t = t.contiguous()
                st.keep.append(t)
                cp.arrays[k] = t.data_ptr()
                cp.array_stride[k] = 0 if t.numel() == 1 else 1
    for i in range(_abi.N_PARAMS):
        for j in range(_abi.N_PFT):
            cp.table[i][j] = tab[i, j]

    n_out = {'sum': 1, 'day_night': 2, 'components': 6, 'annual': 1, 'potential': 6,
             'composite': 1, 'sum_i16': 1}[out_mode]
    flags = {'static': _abi.SEM_STATIC, 'instance': _abi.SEM_INSTANCE}[semantics]
    flags |= {'sum': _abi.OUT_SUM, 'day_night': _abi.OUT_DAY_NIGHT,
              'components': _abi.OUT_COMPONENTS, 'annual': _abi.OUT_ANNUAL,
              'potential': _abi.OUT_POTENTIAL, 'composite': _abi.OUT_COMPOSITE,
              'sum_i16': _abi.OUT_SUM_I16}[out_mode]
    flags |= _abi.GATE_FLAGS[gate] | force
    if out_mode == 'annual':
        outs = list(out)                    # (n_groups, N) int64 accumulators, zeroed
    elif out_mode == 'sum_i16':
        if not out_scale or out_scale <= 0:
            raise ValueError('out_mode="sum_i16" needs out_scale > 0')
        outs = list(out) if out is not None else [
            torch.empty((T, N), device=dev, dtype=torch.int16)]
    elif out is not None:
        outs = list(out)
        if len(outs) != n_out or any(
                (not o.is_cuda) or o.dtype != st.t_dtype or o.numel() != T * N
                or not o.is_contiguous() for o in outs):
            raise ValueError('`out` must be %d contiguous CUDA tensors of %d elements' % (
                n_out, T * N))
        as_torch = True
    else:
        outs = [torch.empty((T, N), device=dev, dtype=st.t_dtype) for _ in range(n_out)]
    co = _abi.Out()
    if groups is not None:
        co.group, co.n_groups = groups[0].data_ptr(), groups[1]
        st.keep.append(groups[0])
    if out_mode == 'sum_i16':
        co.out_scale = float(out_scale)
        # (no float output to carry overflow markers: a generous queue, like 'annual')
        workspace = torch.empty(max(1 << 20, T * N // 8 + 1024), device=dev, dtype=torch.int64)
    if out_mode == 'composite':
        if weights is None or len(weights) != 2:
            raise TypeError('out_mode="composite" needs weights=(day_seconds, night_seconds)')
        wt = [st.tensor(np.asarray(w.cpu() if isinstance(w, torch.Tensor) else w, np.float32),
                        torch.float32).reshape(-1).contiguous() for w in weights]
        if any(w.numel() != T for w in wt):
            raise ValueError('composite weights must hold one value per time step')
        st.keep.extend(wt)
        co.day_weight, co.night_weight = wt[0].data_ptr(), wt[1].data_ptr()
    for k, o in enumerate(outs):
        co.out[k] = o.data_ptr()
    if gate_flag is None:
        gate_flag = torch.zeros(1, device=dev, dtype=torch.int32)
    co.gate_flag = gate_flag.data_ptr()
    handle = s.cuda_stream
    if np_dtype == np.dtype('float32'):
        ws = workspace if workspace is not None else _workspace(dev, handle, lib, T * N)
        co.workspace = ws.data_ptr()
        co.workspace_bytes = ws.numel() * ws.element_size()
        st.keep.append(ws)
    if prep is not None:
        fn = lib.mod16_forward_raw_f32
    else:
        fn = lib.mod16_forward_f32 if np_dtype == np.dtype('float32') else lib.mod16_forward_f64
    _abi.check(fn(ctypes.byref(cd), ctypes.byref(cp), ctypes.byref(co), T, N, flags,
                  ctypes.c_void_p(handle)))
    if out_mode == 'annual':
        return outs
    if out_mode == 'sum_i16':
        if int(gate_flag.item()) & _abi.ANNUAL_OVERFLOW:
            raise _abi.Mod16Error('out_mode="sum_i16": more than an eighth of the elements need '
                                  'the double-precision pass; mask fill values first')
        o = outs[0].reshape(full_shape)
        hand_over((o,), caller_stream)
        return [RawDriver(o if as_torch else o.cpu().numpy(), float(out_scale), 0.0, -32768)]
    outs = [o.reshape(full_shape) for o in outs]
    if as_torch:
        # Staged copies were allocated and used on `s` alone: the caching allocator's stream
        # ordering keeps them alive.  The results belong to `s` too; the caller orders its
        # consumers after `s` (and will most likely read them on its own stream).
        hand_over(outs, caller_stream)
        return outs
    res = [hostcopy.download(o, s) if (o.numel() * o.element_size() >=
                                       hostcopy.THRESHOLD_BYTES)
           else o.cpu().numpy() for o in outs]  # issued on `s`; synchronises
    if full_shape == ():
        res = [r[()] for r in res]
    return res
