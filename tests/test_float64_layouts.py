'''
GPU tests of the float64 entry's two fast kernels (csrc/mod16_forward_fast64.cu) and of the
dispatch between them (bit 8 of mod16_last_forward_kernel(): canonical-layout kernel):

* the canonical-layout kernel (every time-varying source a (T, N) array of one width and one
  time stride, fetched a step ahead with cp.async) and the general kernel (one pointer per
  source) give the same bits on a canonical call -- for both widths, both semantics, every
  output mode, a ragged pixel count, a time chunk with an odd tail, NaN-masked pixels and
  fill values; the static sum is also held to 1e-10 of the oracle here (the other modes and
  the instance semantics of this algebra are pinned to the oracle in test_forward_parity.py);
* a pixel range of a wider array stays canonical, a (T, 1) broadcast driver does not;
* a scalar driver that is not a float32 value keeps float32 arrays off the all-float32 kernel;
* `r_corr_list` (two more sources) on both kernels;
* both kernels on elements placed 1e-13 ... 1e-7 on either side of every switch of the model
  (tests/switch_cases.py), against the oracle.
'''
import os

import numpy as np
import pytest
import torch

import parity
from mod16_b200 import MOD16, _abi, synthetic, utils
from mod16_b200.forward import run_forward
from oracle import mod16_oracle as orc

pytestmark = pytest.mark.gpu
NAMES = list(synthetic.DRIVER_NAMES)
CANON = 1   # mod16_last_forward_kernel() >> 8 for the canonical-layout kernel


def _table():
    return utils.bplut_table(utils.restore_bplut(utils.BPLUT_C51), MOD16.required_parameters,
                             default_beta=250.0)


def _d64(drv):
    return [v.astype(np.float64) if isinstance(v, np.ndarray) else v for v in drv]


def _kernel(lib):
    k = lib.mod16_last_forward_kernel()
    return k & 0xff, k >> 8


class _general_kernel(object):
    'MOD16_F64_GENERAL=1 (read at every launch) keeps a canonical call on the general kernel.'
    def __enter__(self):
        os.environ['MOD16_F64_GENERAL'] = '1'

    def __exit__(self, *exc):
        os.environ.pop('MOD16_F64_GENERAL', None)


def _same_bits(a, b):
    return all(np.array_equal(x.view(np.int64), y.view(np.int64)) for x, y in zip(a, b))


@pytest.mark.parametrize('width', [np.float64, np.float32])
@pytest.mark.parametrize('semantics', ['static', 'instance'])
@pytest.mark.parametrize('out_mode', ['sum', 'day_night', 'components'])
def test_canonical_and_general_kernels_agree(cuda_lib, width, semantics, out_mode):
    T, N = 70, 1000                       # 64-step chunks: a 6-step tail; 1000 = 7 * 128 + 104
    d = synthetic.host_drivers(T, N, 311, dtype=width, nan_frac=0.02, fill_frac=0.01)
    drv = [d[k] for k in NAMES]
    table, pft = _table(), synthetic.host_pft_map(N, 312)
    kw = dict(pft=pft, table=table, semantics=semantics, out_mode=out_mode, gate='global',
              dtype=np.float64)
    canon = run_forward(drv, **kw)
    assert _kernel(cuda_lib) == (_abi.KERNEL_FAST64, CANON)
    with _general_kernel():
        general = run_forward(drv, **kw)
        assert _kernel(cuda_lib) == (_abi.KERNEL_FAST64, 0)
    assert _same_bits(canon, general)
    # and against the oracle, on the float64 values of the same inputs (the instance
    # semantics of this algebra are pinned in test_forward_parity.py; here: the same bits)
    if semantics == 'static' and out_mode == 'sum':
        vec = [table[i][pft].reshape(1, N) for i in range(11)]
        with np.errstate(all='ignore'):
            ref = orc.static_et(vec, *_d64(drv))
        parity.assert_parity(canon[0], ref, np.float64, what='canonical sum')


def test_pixel_range_is_canonical_and_a_broadcast_driver_is_not(cuda_lib):
    dev = torch.device('cuda', 0)
    T, N = 9, 4000
    d = synthetic.device_drivers(T, N, 321, dev)
    lst = [d[k].double() if isinstance(d[k], torch.Tensor) else d[k] for k in NAMES]
    table = _table()
    pft = torch.from_numpy(synthetic.host_pft_map(N, 322)).to(dev)
    kw = dict(table=table, gate='open', dtype=np.float64, device=dev, as_torch=True)
    full = run_forward(lst, pft=pft, **kw)[0]
    assert _kernel(cuda_lib) == (_abi.KERNEL_FAST64, CANON)
    lo, hi = 131, 3777                    # time stride N, not the range's width
    sl = [v[:, lo:hi] if isinstance(v, torch.Tensor) and v.ndim == 2 else
          (v[lo:hi] if isinstance(v, torch.Tensor) else v) for v in lst]
    part = run_forward(sl, pft=pft[lo:hi], **kw)[0]
    assert _kernel(cuda_lib) == (_abi.KERNEL_FAST64, CANON)
    assert torch.equal(part.view(torch.int64), full[:, lo:hi].contiguous().view(torch.int64))
    # one driver the same at every pixel: a (T, 1) array (pixel stride 0) -> general kernel;
    # the same values spelled out as a (T, N) array -> canonical kernel; same bits
    k = NAMES.index('sw_rad_night')
    col = lst[NAMES.index('sw_rad_day')][:, :1].contiguous() * 0.01
    bro = list(lst); bro[k] = col
    spelled = list(lst); spelled[k] = col.expand(T, N).contiguous()
    a = run_forward(bro, pft=pft, **kw)[0]
    assert _kernel(cuda_lib) == (_abi.KERNEL_FAST64, 0)
    b = run_forward(spelled, pft=pft, **kw)[0]
    assert _kernel(cuda_lib) == (_abi.KERNEL_FAST64, CANON)
    assert torch.equal(a.view(torch.int64), b.view(torch.int64))


def test_scalar_driver_that_is_not_a_float32_value(cuda_lib):
    '''
    The all-float32 kernel reads scalar drivers through the same 4-byte slots as the arrays:
    only a scalar that IS a float32 value may take it.  288.123456789 is not, 288.5 is.
    '''
    T, N = 6, 777
    d = synthetic.host_drivers(T, N, 331, dtype=np.float32)
    table, pft = _table(), synthetic.host_pft_map(N, 332)
    vec = [table[i][pft].reshape(1, N) for i in range(11)]
    for t_ann, want in ((288.5, CANON), (288.123456789, 0)):
        drv = [d[k] for k in NAMES]
        drv[NAMES.index('temp_annual')] = t_ann
        got = run_forward(drv, pft=pft, table=table, dtype=np.float64)[0]
        assert _kernel(cuda_lib) == (_abi.KERNEL_FAST64, want), t_ann
        with np.errstate(all='ignore'):
            ref = orc.static_et(vec, *_d64(drv))
        parity.assert_parity(got, ref, np.float64, what='scalar temp_annual %r' % t_ann)


def test_r_corr_list_on_both_kernels(cuda_lib):
    T, N = 7, 900
    d = synthetic.host_drivers(T, N, 341, dtype=np.float64)
    drv = [d[k] for k in NAMES]
    rc = [np.broadcast_to(orc.r_correction(d['pressure'], d[t]) * f, (T, N)).copy()
          for t, f in (('temp_day', 1.1), ('temp_night', 0.9))]
    table, pft = _table(), synthetic.host_pft_map(N, 342)
    vec = [table[i][pft].reshape(1, N) for i in range(11)]
    kw = dict(pft=pft, table=table, dtype=np.float64, r_corr_list=rc)
    canon = run_forward(drv, **kw)
    assert _kernel(cuda_lib) == (_abi.KERNEL_FAST64, CANON)
    with _general_kernel():
        general = run_forward(drv, **kw)
        assert _kernel(cuda_lib) == (_abi.KERNEL_FAST64, 0)
    assert _same_bits(canon, general)
    with np.errstate(all='ignore'):
        ref = orc.static_et(vec, *drv, r_corr_list=rc)
    parity.assert_parity(canon[0], ref, np.float64, what='r_corr_list')
    # float32 drivers: r_corr_list is widened by the host (f32_mask covers the 14 drivers
    # only), so the launch holds both widths and takes the general kernel
    d32 = [v.astype(np.float32) if isinstance(v, np.ndarray) else v for v in drv]
    got = run_forward(d32, **kw)
    assert _kernel(cuda_lib) == (_abi.KERNEL_FAST64, 0)
    with np.errstate(all='ignore'):
        ref = orc.static_et(vec, *_d64(d32), r_corr_list=rc)
    parity.assert_parity(got[0], ref, np.float64, what='r_corr_list, float32 drivers')


@pytest.mark.parametrize('semantics', ['static', 'instance'])
def test_float64_kernels_at_the_discontinuities(cuda_lib, semantics):
    '''
    tests/switch_cases.py: elements 1e-13 ... 1e-7 on both sides of every switch of the model
    (rh = 0.7, rh -> 0+, rh >= 1, |G| = 0.39 |A|, A_day = 0, the night soil-radiation
    condition, the 5 K day-night switch, the annual-temperature window, the ramp ends).  The
    host build of the same algebra is held to the oracle in tests/test_hostmath.py; here the
    CUDA kernels are -- the canonical-layout kernel and the general one, whose reciprocals
    are the hardware-seeded ones -- on the same elements laid out as a (T, N) block, so that
    the time loop and the input ring take part.  Per-element gate, all six components.
    '''
    import switch_cases
    drv, groups, grp, off = switch_cases.build(semantics)
    T = 40
    N = drv[0].size // T
    assert T * N == drv[0].size
    blk = [np.ascontiguousarray(v.reshape(T, N)) for v in drv]
    day, night, scale = switch_cases.reference(semantics, blk)
    kw = dict(params=switch_cases.BPLUT7, semantics=semantics, out_mode='components',
              gate='element', dtype=np.float64)
    canon = run_forward(blk, **kw)
    assert _kernel(cuda_lib) == (_abi.KERNEL_FAST64, CANON)
    with _general_kernel():
        general = run_forward(blk, **kw)
        assert _kernel(cuda_lib) == (_abi.KERNEL_FAST64, 0)
    assert _same_bits(canon, general)
    for c in range(3):
        parity.assert_parity(canon[c], day[c], np.float64, scale, what='day %d' % c)
        parity.assert_parity(canon[3 + c], night[c], np.float64, scale, what='night %d' % c)
    # and the summed outputs of the same block (float64 arrays only: float32 cannot hold
    # offsets of 1e-13)
    tot = (day[0] + day[1] + day[2]) + (night[0] + night[1] + night[2])
    kw['out_mode'] = 'sum' if semantics == 'static' else 'day_night'
    got = run_forward(blk, **kw)
    got = got[0] if semantics == 'static' else got[0] + got[1]
    parity.assert_parity(got, tot, np.float64, scale, what='total')
