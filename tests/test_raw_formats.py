'''
Drivers in their storage formats (mod16_forward_raw_f32): the conversions the
reference's loaders apply on the host -- `fpar /= 100`, `lai /= 10` (calibration.py:422-423),
`vpd_night < 0 -> 0` (:401), `pressure = MOD16.air_pressure(elevation)` (:405-408) -- fused
into K1.  Parity: the oracle evaluated on the physical values of the same stored integers.
'''
import numpy as np
import pytest
import torch

import parity
from mod16_b200 import MOD16, _abi, rawio, synthetic, utils
from mod16_b200.forward import run_forward
from oracle import mod16_oracle as orc

NAMES = list(synthetic.DRIVER_NAMES)
I16_MET = {   # name -> (scale, offset): int16 packings of the meteorological drivers
    'lw_net_day': (0.01, 0.0), 'lw_net_night': (0.01, 0.0), 'sw_rad_day': (0.02, 0.0),
    'temp_day': (0.01, 273.15), 'temp_night': (0.01, 273.15), 'tmin': (0.01, 273.15),
    'vpd_day': (0.25, 0.0), 'vpd_night': (0.25, 0.0)}


def _table():
    return utils.bplut_table(utils.restore_bplut(utils.BPLUT_C51), MOD16.required_parameters,
                             default_beta=250.0)


def _stored(d, layout, rng, fills=True):
    'The drivers of `host_drivers` as a loader would find them in the archive.'
    out = dict(d)
    out['fpar'] = rawio.quantize(d['fpar'], np.uint8, 0.01, 0.0, fill=255)  # percent
    out['lai'] = rawio.quantize(d['lai'], np.uint8, 0.1, 0.0, fill=255)     # tenths
    out['sw_albedo'] = rawio.quantize(d['sw_albedo'], np.int16, 0.001, 0.0, fill=32767)
    if layout == 'packed':
        for k, (sc, off) in I16_MET.items():
            out[k] = rawio.quantize(d[k], np.int16, sc, off, fill=-32768)
    if fills:
        for k in ('fpar', 'sw_albedo') + (('temp_day', 'vpd_night') if layout == 'packed' else ()):
            raw = out[k].raw
            raw[rng.random(raw.shape) < 0.01] = out[k].fill
    return out


def test_quantize_roundtrip_and_physical_values():
    x = np.array([[0.0, 0.5, 1.0, np.nan], [0.004, 0.996, 2.6, -1.0]])
    r = rawio.quantize(x, np.uint8, 0.01, 0.0, fill=255)
    assert r.raw.dtype == np.uint8 and r.raw.tolist() == [[0, 50, 100, 255], [0, 100, 255, 0]]
    p = r.physical()
    assert np.isnan(p[0, 3]) and np.isnan(p[1, 2]) and p[0, 1] == 0.5
    t = rawio.quantize(np.array([250.0, 273.15, 320.0]), np.int16, 0.01, 273.15, fill=-32768)
    assert np.allclose(t.physical(), [250.0, 273.15, 320.0], atol=0.006)
    assert rawio.prep_word(('vpd_night_clamp', 'elevation')) == 3 and rawio.prep_word(None) == 0
    with pytest.raises(ValueError):
        rawio.prep_word('nope')
    with pytest.raises(TypeError):
        rawio.RawDriver(np.zeros(3, np.int32))
    assert r[0].shape == (4,) and r.dtype == np.float32


def _reference(stored, prep, col_or_vec, gate='global'):
    phys = []
    for k in NAMES:
        v = stored[k]
        v = v.physical() if isinstance(v, rawio.RawDriver) else (
            v.astype(np.float64) if isinstance(v, np.ndarray) else v)
        if k == 'vpd_night' and 'vpd_night_clamp' in prep:
            v = np.where(v < 0, 0, v)                                     # calibration.py:401
        if k == 'pressure' and 'elevation' in prep:
            v = orc.air_pressure(v).astype(np.float32).astype(np.float64)  # :405-408
        phys.append(v)
    with np.errstate(all='ignore'):
        return orc.static_evapotranspiration(col_or_vec, *phys, gate=gate), phys


@pytest.mark.gpu
@pytest.mark.parametrize('layout,fmt', [('modis', 2), ('packed', 3)])
@pytest.mark.parametrize('prep', [(), ('vpd_night_clamp', 'elevation')])
def test_storage_formats_match_the_oracle_on_the_same_stored_values(cuda_lib, layout, fmt, prep):
    T, N = 19, 4816
    rng = np.random.default_rng(7)
    d = synthetic.host_drivers(T, N, 91, dtype=np.float32)
    d['vpd_night'][rng.random((T, N)) < 0.05] *= -1.0          # the clamp has work to do
    if 'elevation' in prep:
        d['pressure'] = rng.uniform(0, 3500, N).astype(np.float32)     # metres
    stored = _stored(d, layout, rng)
    drv = [stored[k] for k in NAMES]
    table = _table()
    pft = synthetic.host_pft_map(N, 92)
    vec = [table[i][pft].reshape(1, N) for i in range(11)]
    (rd, rn), _ = _reference(stored, prep, vec)
    got = MOD16._et_by_pft(table, pft, *drv, prep=prep)
    k, v = cuda_lib.mod16_last_forward_kernel() & 0xff, cuda_lib.mod16_last_forward_kernel() >> 12
    assert (k, v) == (_abi.KERNEL_PIPE, fmt)
    m = parity.assert_parity(got, rd + rn, np.float32, what='%s sum' % layout)
    assert 0.015 < np.isnan(got).mean() < 0.06                 # the fill values came out as NaN
    gd, gn = run_forward(drv, pft=pft, table=table, out_mode='day_night', prep=prep)
    parity.assert_parity(gd, rd, np.float32, what='%s day' % layout)
    parity.assert_parity(gn, rn, np.float32, what='%s night' % layout)
    # instance semantics through the same formats
    col = list(table[:, 7])
    inst = MOD16(dict(zip(orc.PARAM_ORDER, col)))
    _, phys = _reference(stored, prep, col)
    with np.errstate(all='ignore'):
        want = orc.InstanceModel(dict(zip(orc.PARAM_ORDER, col))).evapotranspiration(*phys)
    day, night = inst.evapotranspiration(*drv, prep=prep)
    assert cuda_lib.mod16_last_forward_kernel() >> 12 == fmt
    parity.assert_parity(day, want[0], np.float32, parity.MASS_SCALE, what='instance day')
    parity.assert_parity(night, want[1], np.float32, parity.MASS_SCALE, what='instance night')


@pytest.mark.gpu
def test_float32_drivers_with_loader_conversions(cuda_lib):
    T, N = 11, 2400
    rng = np.random.default_rng(8)
    d = synthetic.host_drivers(T, N, 93, dtype=np.float32)
    d['vpd_night'][rng.random((T, N)) < 0.1] *= -1.0
    d['vpd_night'][rng.random((T, N)) < 0.01] = np.nan                   # a NaN stays a NaN
    d['pressure'] = rng.uniform(0, 3500, N).astype(np.float32)
    prep = ('vpd_night_clamp', 'elevation')
    col = list(_table()[:, 2])
    (rd, rn), _ = _reference(d, prep, col)
    got = MOD16._et(col, *[d[k] for k in NAMES], prep=prep)
    assert cuda_lib.mod16_last_forward_kernel() >> 12 == 1
    parity.assert_parity(got, rd + rn, np.float32, what='float32 + prep')
    # the conversions are the library's own K4 operations as well
    p = MOD16.air_pressure(d['pressure'].astype(np.float64))
    assert np.allclose(p, orc.air_pressure(d['pressure'].astype(np.float64)), rtol=1e-12)


@pytest.mark.gpu
def test_layouts_the_kernel_cannot_stream_are_converted_on_the_device(cuda_lib):
    'N % 16 != 0, components output, float64: same results through the plain entries.'
    T, N = 7, 1004
    rng = np.random.default_rng(9)
    d = synthetic.host_drivers(T, N, 94, dtype=np.float32)
    stored = _stored(d, 'packed', rng)
    drv = [stored[k] for k in NAMES]
    col = list(_table()[:, 5])
    (rd, rn), phys = _reference(stored, (), col)
    got = MOD16._et(col, *drv)
    assert cuda_lib.mod16_last_forward_kernel() >> 12 == 0          # plain kernel after torch
    parity.assert_parity(got, rd + rn, np.float32, what='ragged N')
    with np.errstate(all='ignore'):
        comps = orc.static_evapotranspiration(col, *phys, components=True)
    six = run_forward(drv, params=col, out_mode='components')
    for h in range(2):
        for c in range(3):
            parity.assert_parity(six[3 * h + c], comps[h][c], np.float32, what='components')
    g64 = MOD16._et(col, *drv, dtype=np.float64)
    # (float64 arithmetic on the float32 conversions of the stored values)
    assert np.allclose(np.nan_to_num(g64), np.nan_to_num(rd + rn), rtol=2e-5, atol=1e-3)
    # the raw C entry itself refuses instead of taking a slow path
    from mod16_b200.forward import _run_on_stream  # noqa: F401  (the entry is reached below)
    with pytest.raises(_abi.LayoutError):
        import ctypes
        rdv = _abi.RawDrivers()
        t = torch.zeros((T, N), dtype=torch.uint8, device='cuda')
        for k in range(14):
            rdv.drv[k].ptr, rdv.drv[k].t_stride, rdv.drv[k].n_stride = t.data_ptr(), N, 1
            rdv.drv[k].dtype = _abi.T_U8
        par, out = _abi.Params(), _abi.Out()
        o = torch.zeros((T, N), device='cuda')
        flag = torch.zeros(1, dtype=torch.int32, device='cuda')
        out.out[0], out.gate_flag = o.data_ptr(), flag.data_ptr()
        _abi.check(cuda_lib.mod16_forward_raw_f32(ctypes.byref(rdv), ctypes.byref(par),
                                                  ctypes.byref(out), T, N, 0, None))


@pytest.mark.gpu
def test_streamed_forward_with_stored_slabs(cuda_lib):
    from mod16_b200.streaming import StreamedForward
    T, N = 24, 4800
    rng = np.random.default_rng(10)
    d = synthetic.host_drivers(T, N, 95, dtype=np.float32)
    stored = _stored(d, 'packed', rng, fills=False)
    table = _table()
    pft = synthetic.host_pft_map(N, 96)
    host, formats = {}, {}
    for k in NAMES:
        v = stored[k]
        if isinstance(v, rawio.RawDriver):
            host[k] = rawio.RawDriver(torch.from_numpy(v.raw).pin_memory(), v.scale, v.offset, v.fill)
            formats[k] = (v.np_dtype, v.scale, v.offset, v.fill)
        elif isinstance(v, np.ndarray):
            host[k] = torch.from_numpy(v).pin_memory()
        else:
            host[k] = v
    out_host = torch.empty((T, N), dtype=torch.float32).pin_memory()
    runner = StreamedForward(N, 5, table, pft, formats=formats)
    runner.run(host, out_host)
    assert runner.h2d_bytes == T * N * 20 + 2 * N * 4
    vec = [table[i][pft].reshape(1, N) for i in range(11)]
    (rd, rn), _ = _reference(stored, (), vec, gate='open')
    parity.assert_parity(out_host.numpy(), rd + rn, np.float32, what='streamed stored slabs')


@pytest.mark.gpu
def test_stored_drivers_with_the_other_output_modes(cuda_lib):
    'Annual sums and potential terms from stored drivers: converted on the device, then fused.'
    T, N = 20, 2400
    rng = np.random.default_rng(12)
    d = synthetic.host_drivers(T, N, 97, dtype=np.float32)
    stored = _stored(d, 'modis', rng, fills=False)
    drv = [stored[k] for k in NAMES]
    col = list(_table()[:, 7])
    (rd, rn), phys = _reference(stored, (), col)
    years = np.repeat([0, 1], T // 2)
    with np.errstate(all='ignore'):
        want = orc.annual_mass_sums(rd + rn, years, orc.lhv(0.5 * (phys[5] + phys[6])))
    got = MOD16._annual_et(col, years, *drv)
    assert np.all(np.abs(got - want) <= 1e-5 * np.maximum(np.abs(want), 1.0))
    pot = MOD16._et_potential(col, *drv)
    parity.assert_parity(pot['et'], rd + rn, np.float32, what='total from stored drivers')
    with np.errstate(all='ignore'):
        pt = orc.InstanceModel.potential_transpiration(phys[0], phys[2], phys[4], phys[11], phys[5],
                                                       phys[9], phys[12])
    parity.assert_parity(pot['potential_transpiration'], pt, np.float32, what='PT')


@pytest.mark.gpu
@pytest.mark.parametrize('N', [4800, 1003])
def test_vpd_from_mixing_ratio_and_surface_pressure(cuda_lib, N):
    '''
    `MOD16.vpd(qv10m, ps, temp)` (mod16/__init__.py:605-644; calibration.py:393-400) formed
    inside K1 from QV10M and surface pressure, against the oracle's `vpd_from_qv` in double
    followed by `_et` (N = 1003: the same through the conversion on the device).
    '''
    T = 17
    rng = np.random.default_rng(13)
    d = synthetic.host_drivers(T, N, 98, dtype=np.float32)
    ps_d = rng.uniform(7.0e4, 1.02e5, (T, N)).astype(np.float32)
    ps_n = (ps_d * rng.uniform(0.995, 1.005, (T, N))).astype(np.float32)
    # mixing ratios that give the synthetic VPDs back: qv = 0.622 avp / (ps - 0.379 avp)
    def qv_of(vpd, temp, ps):
        tc = temp.astype(np.float64) - 273.15
        avp = 610.7 * np.exp(17.38 * tc / (239 + tc)) - vpd
        return (0.622 * avp / (ps - 0.379 * avp)).astype(np.float32)
    qv_d, qv_n = qv_of(d['vpd_day'], d['temp_day'], ps_d), qv_of(d['vpd_night'], d['temp_night'], ps_n)
    drv = [d[k] for k in NAMES]
    drv[9], drv[10] = qv_d, qv_n
    d64 = _d64 = [v.astype(np.float64) if isinstance(v, np.ndarray) else v for v in [d[k] for k in NAMES]]
    vd = orc.vpd_from_qv(qv_d.astype(np.float64), ps_d.astype(np.float64), d64[5])
    vn = orc.vpd_from_qv(qv_n.astype(np.float64), ps_n.astype(np.float64), d64[6])
    vn = np.where(vn < 0, 0, vn)                                   # calibration.py:401
    ref_in = list(d64)
    ref_in[9], ref_in[10] = vd, vn
    col = list(_table()[:, 7])
    with np.errstate(all='ignore'):
        rd, rn = orc.static_evapotranspiration(col, *ref_in)
    got = MOD16._et(col, *drv, surface_pressure=(ps_d, ps_n), prep=('vpd_night_clamp',))
    k = cuda_lib.mod16_last_forward_kernel()
    if N % 4 == 0:
        assert (k & 0xff, (k >> 8) & 0xf, k >> 12) == (_abi.KERNEL_PIPE, 4, 1)
    # the VPD is a difference of two O(1e3) Pa numbers formed in float32: 1e-2 Pa of noise
    m = parity.mismatch(got, rd + rn, np.float32)
    assert m['nan_mismatch'] == 0 and m['max_abs'] < 5e-3 and m['n_fail'] <= 2e-4 * m['n'], m
    gd, gn = run_forward(drv, params=col, out_mode='day_night', surface_pressure=(ps_d, ps_n),
                         prep=('vpd_night_clamp',))
    assert np.abs(gd - rd).max() < 5e-3 and np.abs(gn - rn).max() < 5e-3


@pytest.mark.gpu
def test_per_element_gate_with_stored_drivers_and_pipeline_only_outputs(cuda_lib):
    '''
    The per-element gate of the static float32 path needs the guard at the zeros of the Tmin /
    VPD ramps, which the pipelined kernel carries in separate instantiations
    for plain float32 drivers and the three classic outputs only:
    * stored drivers are converted on the device and evaluated by the plain entry (on that
      guard-enabled pipeline) -- same values as the oracle's per-element gate on the stored
      values;
    * an epilogue output (here the potential terms) answers MOD16_ERR_LAYOUT;
    * the same calls with the batch-global gate stay on the fused pipeline.
    '''
    T, N = 6, 1024
    rng = np.random.default_rng(19)
    d = synthetic.host_drivers(T, N, 95, dtype=np.float32)
    stored = _stored(d, 'packed', rng)
    drv = [stored[k] for k in NAMES]
    col = list(_table()[:, 5])
    (rd, rn), _ = _reference(stored, (), col, gate='element')
    got = run_forward(drv, params=col, gate='element')[0]
    k = cuda_lib.mod16_last_forward_kernel()
    assert k & 0xff == _abi.KERNEL_PIPE and k >> 12 == 0          # plain float32 after torch
    parity.assert_parity(got, rd + rn, np.float32, what='stored drivers, per-element gate')
    run_forward(drv, params=col, gate='global')
    k = cuda_lib.mod16_last_forward_kernel()
    assert k & 0xff == _abi.KERNEL_PIPE and k >> 12 == 3            # fused, PACKED format
    plain = [d[n] for n in NAMES]
    with pytest.raises(_abi.LayoutError):
        run_forward(plain, params=col, gate='element', out_mode='potential')
    six = run_forward(plain, params=col, gate='global', out_mode='potential')
    assert cuda_lib.mod16_last_forward_kernel() & 0xff == _abi.KERNEL_PIPE and len(six) == 6
