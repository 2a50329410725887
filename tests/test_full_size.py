'''
GPU tests at the BASELINE.json sizes, where the oracle cannot loop over everything:
size-independent properties plus an oracle check on a random sample of elements.

* config 1: a single tower site, 20-year daily series (T = 7305, N = 1, seed 16001), all
  three entry points, float64 and float32, whole series against the oracle;
* config 2: the MODIS tile-year (T = 365, N = 2400 x 2400, float32, 100.9 GB resident):
    - a random sample of 40 000 elements (plus every element the kernel routed to the exact
      path in a slab) against the float64 oracle on the same inputs;
    - determinism: a second run is bit-identical;
    - partition invariance (the multi-GPU contract): any pixel range evaluated on its own
      equals the slice of the full result, for ranges that do and do not align with the
      kernel's 480-pixel tiles, through both the bulk-copy and the strided kernel;
    - checksum of checksums: the per-step sums of the output equal the sums over the pixel
      ranges of a 7-way split.
'''
import numpy as np
import pytest
import torch

import parity
from mod16_b200 import MOD16, _abi, synthetic, utils
from mod16_b200.forward import run_forward
from oracle import mod16_oracle as orc

pytestmark = pytest.mark.gpu
NAMES = list(synthetic.DRIVER_NAMES)


def _table():
    return utils.bplut_table(utils.restore_bplut(utils.BPLUT_C51), MOD16.required_parameters,
                             default_beta=250.0)


@pytest.mark.parametrize('dtype', [np.float64, np.float32])
def test_single_site_20_year_series(cuda_lib, dtype):
    'BASELINE configs[0] / SURVEY 8(d) config 1 (notebook Forward_Run_example.ipynb:423).'
    T = 7305
    d = synthetic.host_drivers(T, 1, 16001, dtype=dtype)
    drv = [d[k][:, 0] if isinstance(d[k], np.ndarray) and d[k].ndim == 2 else
           (float(d[k][0]) if isinstance(d[k], np.ndarray) else d[k]) for k in NAMES]
    d64 = [v.astype(np.float64) if isinstance(v, np.ndarray) else v for v in drv]
    col = list(_table()[:, 7])
    with np.errstate(all='ignore'):
        ref = orc.static_et(col, *d64)
        rd, rn = orc.static_evapotranspiration(col, *d64)
        inst = orc.InstanceModel(dict(zip(orc.PARAM_ORDER, col))).evapotranspiration(
            *d64, separate=True)
    got = MOD16._et(col, *drv, dtype=dtype)
    assert got.shape == (T,)
    parity.assert_parity(got, ref, dtype, what='_et series')
    gd, gn = MOD16._evapotranspiration(col, *drv, dtype=dtype)
    parity.assert_parity(gd, rd, dtype, what='day series')
    parity.assert_parity(gn, rn, dtype, what='night series')
    day, night = MOD16(dict(zip(orc.PARAM_ORDER, col))).evapotranspiration(
        *drv, separate=True, dtype=dtype)
    for h, got_h in enumerate((day, night)):
        for c in range(3):
            parity.assert_parity(got_h[c], inst[h][c], dtype, parity.MASS_SCALE,
                                 what='instance series %d %d' % (h, c))


def _resident(T, N, seed, device):
    names = [k for k in NAMES if k not in synthetic.PER_PIXEL and k != 'sw_rad_night']
    big = dict((k, torch.empty((T, N), device=device, dtype=torch.float32)) for k in names)
    slab = max(1, (1 << 27) // N)
    first = None
    for t0 in range(0, T, slab):
        t1 = min(T, t0 + slab)
        d = synthetic.device_drivers(t1 - t0, N, seed + t0, device)
        first = first or d
        for k in names:
            big[k][t0:t1].copy_(d[k])
        del d
    big['temp_annual'], big['pressure'], big['sw_rad_night'] = \
        first['temp_annual'], first['pressure'], 0
    return big


def test_tile_year_properties(cuda_lib):
    dev = torch.device('cuda', 0)
    free, _ = torch.cuda.mem_get_info(dev)
    T, N = 365, 2400 * 2400
    if free < 125e9:   # a smaller resident block on a smaller device; same properties
        T = int(free * 0.6 / (48.0 * N))
    drv = _resident(T, N, 16002, dev)
    # a few fills, so that the exact path and the queue take part at full size
    g = torch.Generator(device=dev)
    g.manual_seed(5)
    for k, val in (('lai', float('nan')), ('temp_day', -9999.0), ('vpd_night', float('nan'))):
        idx = torch.randint(0, T * N, (4000,), generator=g, device=dev)
        drv[k].view(-1)[idx] = val
    table = _table()
    pft = torch.from_numpy(synthetic.host_pft_map(N, 16003)).to(dev)
    lst = [drv[k] for k in NAMES]
    out = torch.empty((T, N), device=dev, dtype=torch.float32)
    flag = torch.zeros(1, device=dev, dtype=torch.int32)
    kw = dict(pft=pft, table=table, semantics='static', out_mode='sum', dtype=np.float32,
              device=dev)
    run_forward(lst, gate='global', gate_flag=flag, out=[out], **kw)
    torch.cuda.synchronize()

    # (1) random sample + the poisoned elements against the float64 oracle
    rng = np.random.default_rng(1)
    ts = torch.from_numpy(rng.integers(0, T, 40000)).to(dev)
    ns = torch.from_numpy(rng.integers(0, N, 40000)).to(dev)
    bad = torch.nonzero(torch.isnan(drv['lai'].view(-1)) | (drv['temp_day'].view(-1) == -9999.0)
                        | torch.isnan(drv['vpd_night'].view(-1)))[:, 0][:6000]
    ts = torch.cat([ts, bad // N])
    ns = torch.cat([ns, bad % N])
    samp = []
    for k in NAMES:
        v = drv[k]
        if not isinstance(v, torch.Tensor):
            samp.append(v)
        elif v.ndim == 2:
            samp.append(v[ts, ns].double().cpu().numpy())
        else:
            samp.append(v[ns].double().cpu().numpy())
    codes = pft[ns].cpu().numpy()
    vec = [table[i][codes] for i in range(11)]
    with np.errstate(all='ignore'):
        ref = orc.static_et(vec, *samp, gate='open')   # (the block's gate is open)
    assert int(flag.item()) == 1
    got = out[ts, ns].cpu().numpy()
    m = parity.assert_parity(got, ref, np.float32, what='tile-year sample')
    assert np.isnan(ref).sum() > 1000, 'the poisoned elements must be in the sample'
    print('tile-year sample:', m)

    # (2) determinism
    out2 = torch.empty_like(out[:32])
    first = out[:32].clone()
    run_forward([v[:32] if isinstance(v, torch.Tensor) and v.ndim == 2 else v for v in lst],
                gate='open', out=[out2], **kw)
    assert torch.equal(out2.view(torch.int32), first.view(torch.int32))

    # (3) partition invariance, aligned and ragged ranges, both kernels
    Ts = 48
    for lo, hi, force in ((0, 480 * 100, 0), (480 * 7 + 4, 480 * 900 + 8, 0),
                          (1234567, 1234567 + 99999, 0), (N - 77777, N, 0),
                          (2400 * 1000, 2400 * 1000 + 50000, _abi.FORCE_GENERIC)):
        sl = [v[:Ts, lo:hi] if isinstance(v, torch.Tensor) and v.ndim == 2 else
              (v[lo:hi] if isinstance(v, torch.Tensor) else v) for v in lst]
        part = run_forward(sl, pft=pft[lo:hi], table=table, semantics='static', out_mode='sum',
                           dtype=np.float32, device=dev, gate='open', force=force,
                           as_torch=True)[0]
        a, b = part.view(torch.int32), out[:Ts, lo:hi].contiguous().view(torch.int32)
        nan_ok = torch.isnan(part) & torch.isnan(out[:Ts, lo:hi])
        assert bool(((a == b) | nan_ok).all()), (lo, hi, force)

    # (4) checksum of checksums over a 7-way pixel split (float64 sums, NaN skipped)
    total = torch.nansum(out.double(), dim=1)
    edges = [N * i // 7 for i in range(8)]
    parts = sum(torch.nansum(out[:, a:b].double(), dim=1) for a, b in zip(edges, edges[1:]))
    assert torch.allclose(total, parts, rtol=1e-12)
    assert 1.0 < float(total.mean()) / N < 200.0   # plausible mean ET, W m-2
