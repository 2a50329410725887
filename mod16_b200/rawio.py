'''
Drivers in their storage formats.

The reference's loaders convert the archive's integer products to floating point on the host
before they call the model -- ``fpar /= 100``, ``lai /= 10`` (mod16/calibration.py:422-423,
mod16/sensitivity.py:167-168: MOD15 stores fPAR in percent and LAI in tenths, one byte each),
``vpd_night = np.where(vpd_night < 0, 0, vpd_night)`` (calibration.py:401), ``pressure =
MOD16.air_pressure(elevation)`` (calibration.py:405-408) -- and hand over 48 bytes per
pixel-timestep.  ``RawDriver`` passes the stored integers instead; the fused kernel applies
``raw * scale + offset`` (fill value -> NaN) while it reads them, so a pixel-timestep costs
36 bytes of input (MODIS layout: albedo int16, fPAR / LAI uint8) or 20 (PACKED layout: the
meteorological drivers as int16 too), which is what a PCIe- or HBM-bound run pays for.

    drivers[12] = RawDriver(fpar_percent_u8, scale=0.01, fill=255)
    MOD16._et(params, *drivers, prep=('vpd_night_clamp',))

Layouts the fused kernel cannot stream (see include/mod16_b200.h) are converted to float32
on the device by ``dequantize`` first -- same arithmetic, one extra pass.
'''
import numpy as np
import torch

from . import _abi

DTYPE_CODE = {np.dtype('float32'): _abi.T_F32, np.dtype('int16'): _abi.T_I16,
              np.dtype('uint8'): _abi.T_U8}
PREP_FLAGS = {'vpd_night_clamp': _abi.PREP_VPD_NIGHT_CLAMP, 'elevation': _abi.PREP_ELEVATION}


class RawDriver(object):
    '''
    One driver as stored: ``physical = raw * scale + offset`` (float32 arithmetic, one fused
    multiply-add); elements equal to ``fill`` are missing (NaN).  ``raw`` is a numpy array
    or torch tensor of uint8, int16 or float32 (float32: scale / offset / fill are ignored).
    '''

    def __init__(self, raw, scale=1.0, offset=0.0, fill=None):
        self.raw = raw
        self.scale, self.offset, self.fill = float(scale), float(offset), fill
        dt = np.dtype(str(raw.dtype).replace('torch.', ''))
        if dt not in DTYPE_CODE:
            raise TypeError('RawDriver holds uint8, int16 or float32, not %s' % dt)
        self.np_dtype = dt

    @property
    def shape(self):
        return tuple(self.raw.shape)

    @property
    def ndim(self):
        return len(self.raw.shape)

    @property
    def dtype(self):                       # (numpy promotion: every storage format computes in f32)
        return np.dtype('float32')

    def is_cuda(self):
        return isinstance(self.raw, torch.Tensor) and self.raw.device.type == 'cuda'

    def __getitem__(self, idx):
        return RawDriver(self.raw[idx], self.scale, self.offset, self.fill)

    def physical(self, dtype=np.float64):
        '''
        The values the model sees, for checks on the host: exactly the float32 numbers the
        kernel's fused multiply-add ``raw * float32(scale) + float32(offset)`` produces
        (formed in double, rounded once), fill -> NaN; returned as ``dtype``.
        '''
        raw = self.raw.cpu().numpy() if isinstance(self.raw, torch.Tensor) else np.asarray(self.raw)
        if self.np_dtype == np.dtype('float32'):
            return raw.astype(dtype)
        out = (raw.astype(np.float64) * np.float64(np.float32(self.scale)) +
               np.float64(np.float32(self.offset))).astype(np.float32)
        if self.fill is not None:
            out[raw == self.fill] = np.nan
        return out.astype(dtype)


def quantize(x, dtype, scale=1.0, offset=0.0, fill=None):
    '''
    Stores physical values ``x`` (numpy or torch) as ``dtype``: round((x - offset) / scale),
    clipped to the type's range; NaN becomes ``fill``.  Returns a ``RawDriver``.
    '''
    dt = np.dtype(dtype)
    info = np.iinfo(dt)
    if isinstance(x, torch.Tensor):
        q = torch.round((x.double() - offset) / scale).clamp(info.min, info.max)
        if fill is not None:
            q = torch.where(torch.isnan(x), torch.full_like(q, float(fill)), q)
        raw = q.to(getattr(torch, dt.name))
    else:
        x = np.asarray(x, np.float64)
        with np.errstate(invalid='ignore'):
            q = np.clip(np.round((x - offset) / scale), info.min, info.max)
        if fill is not None:
            q = np.where(np.isnan(x), float(fill), q)
        raw = np.nan_to_num(q).astype(dt)
    return RawDriver(raw, scale, offset, fill)


def dequantize(d, device, prep_clamp=False):
    'float32 device tensor of a RawDriver: the kernel\'s arithmetic as torch operations.'
    t = d.raw if isinstance(d.raw, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(d.raw))
    t = t.to(device, non_blocking=True)
    if d.np_dtype == np.dtype('float32'):
        out = t
    else:
        # (torch.addcmul rounds twice where the kernel's FMA rounds once: <= 1 ulp apart)
        out = t.to(torch.float32) * np.float32(d.scale) + np.float32(d.offset)
        if d.fill is not None:
            out = torch.where(t == d.fill, torch.full_like(out, float('nan')), out)
    if prep_clamp:
        out = torch.where(out < 0, torch.zeros_like(out), out)
    return out


def prep_word(prep):
    word = 0
    for name in ([prep] if isinstance(prep, str) else (prep or ())):
        if name not in PREP_FLAGS:
            raise ValueError('unknown prep %r (known: %s)' % (name, ', '.join(PREP_FLAGS)))
        word |= PREP_FLAGS[name]
    return word
