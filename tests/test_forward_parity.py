'''
GPU parity tests of K1 (the fused forward kernel) through the drop-in Python API, which
calls the C ABI.  Compared with the oracle on the same seeded inputs and with the golden
vectors frozen from the reference.
'''
import os

import numpy as np
import pytest
import torch

import parity
from mod16_b200 import MOD16, _abi, synthetic, utils
from mod16_b200.models import MOD16Collection61
from oracle import mod16_oracle as orc

pytestmark = pytest.mark.gpu
NAMES = list(synthetic.DRIVER_NAMES)
BPLUT7 = [-8., 8.8, 650., 4400., .02, .02, 1e-5, .0055, 60., 95., 250.]


def as_f64(drv):
    return [v.astype(np.float64) if isinstance(v, np.ndarray) else v for v in drv]


def test_unit_test_scalars_like_the_reference(cuda_lib):
    'Reads like /root/reference/tests/tests.py:64-90.'
    params = dict(gl_sh=0.01, gl_wv=0.01, g_cuticular=1e-5, tmin_close=-8, tmin_open=8,
                  vpd_open=650, vpd_close=3000, rbl_min=60, rbl_max=90, csl=2.4e-3,
                  beta=250)
    vec = [params[p] for p in MOD16.required_parameters]
    drv = [-50, -30, 150, 0, 0.3, 293, 290, 285, 285, 1000, 500, 100e3, 0.5, 1.5]
    _et = MOD16._et(vec, *drv)
    _day, _night = MOD16._evapotranspiration(vec, *drv)
    model = MOD16(params)
    _day2, _night2 = model.evapotranspiration(*drv)
    assert round(float(_et), 2) == 40.94
    assert round(float(_day + _night), 2) == 40.94
    assert round(float(_day2 * orc.lhv(293) + _night2 * orc.lhv(290)), 1) == 41.0
    with pytest.raises(KeyError):
        MOD16({'gl_sh': 1})


def test_verify_inline_arrays(cuda_lib, golden_dir):
    z = np.load(os.path.join(golden_dir, 'verify.npz'))
    lw_d, lw_n, sw_d, sw_n, alb, tann, fc, vpd_n = z['scalars']
    a = (lw_d, lw_n, sw_d, sw_n, alb, z['temp'], z['temp'], tann, z['tmin'], z['vpd_day'],
         z['vpd_day'], z['pressure'], fc, z['lai'])
    model = MOD16Collection61(7)
    assert [getattr(model, k) for k in MOD16.required_parameters] == list(z['params'])
    day, night = model.evapotranspiration(*a, f_wet=np.array((0, 0.4, 0.8)), separate=True)
    for h, (hn, got) in enumerate((('day', day), ('night', night))):
        for c, cn in enumerate(('canopy', 'soil', 'trans')):
            parity.assert_parity(got[c], z['inst_%s_%s' % (hn, cn)], np.float64,
                                 parity.MASS_SCALE, what='%s %s' % (hn, cn))
    b = a[:10] + (vpd_n,) + a[11:]
    d, n = MOD16._evapotranspiration(list(z['params']), *b, f_wet=np.array((0, 0.4, 0.8)))
    parity.assert_parity(d, z['static_day'], np.float64, what='static day')
    parity.assert_parity(n, z['static_night'], np.float64, what='static night')
    nb = MOD16Collection61(12).evapotranspiration(  # same model the notebook scalar uses
        -50, -30, 150, 0, 0.3, 293, 285, 285, 285, 1000, 500, 1e5, 0.5, 1.5)
    ref = orc.InstanceModel(dict(zip(orc.PARAM_ORDER, z['notebook_params']))) \
        .evapotranspiration(-50, -30, 150, 0, 0.3, 293, 285, 285, 285, 1000, 500, 1e5, 0.5, 1.5)
    parity.assert_parity(nb[0], ref[0], np.float64, parity.MASS_SCALE)


@pytest.mark.parametrize('dtype', [np.float64, np.float32])
def test_divergence_and_nan_tables(cuda_lib, golden_dir, dtype):
    'SURVEY.md 8(a): every case as its own call (scalar semantics => element gate).'
    z = np.load(os.path.join(golden_dir, 'tables.npz'))
    vec = list(z['params'])
    drv = [z['drivers'][:, i].astype(dtype) for i in range(14)]
    ref = z
    if dtype == np.float32:  # the reference evaluated on the float32-rounded inputs
        with np.errstate(all='ignore'):
            d, n = orc.static_evapotranspiration(vec, *as_f64(drv), gate='element')
            sep = orc.InstanceModel(dict(zip(orc.PARAM_ORDER, vec))) \
                .evapotranspiration(*as_f64(drv), separate=True)
        ref = dict(static_day=d, static_night=n)
        for h, hn in enumerate(('day', 'night')):
            for c, cn in enumerate(('canopy', 'soil', 'trans')):
                ref['inst_%s_%s' % (hn, cn)] = sep[h][c]
    d, n = MOD16._evapotranspiration(vec, *drv, gate='element', dtype=dtype)
    parity.assert_parity(d, ref['static_day'], dtype, what='static day')
    parity.assert_parity(n, ref['static_night'], dtype, what='static night')
    day, night = MOD16(dict(zip(MOD16.required_parameters, vec))).evapotranspiration(
        *drv, separate=True, dtype=dtype)
    for hn, got in (('day', day), ('night', night)):
        for c, cn in enumerate(('canopy', 'soil', 'trans')):
            parity.assert_parity(got[c], ref['inst_%s_%s' % (hn, cn)], dtype,
                                 parity.MASS_SCALE, what='inst %s %s' % (hn, cn))


def test_global_gate_is_call_exact(cuda_lib, golden_dir):
    z = np.load(os.path.join(golden_dir, 'tables.npz'))
    vec = list(z['params'])
    base = [np.float64(v) for v in z['drivers'][0]]
    for key in ('closed', 'open', ):
        drv = list(base)
        drv[NAMES.index('tmin')] = z['gate_%s_tmin' % key]
        got = MOD16._et(vec, *drv)
        parity.assert_parity(got, np.broadcast_to(z['gate_%s_et' % key], got.shape),
                             np.float64, what='gate ' + key)
    drv[NAMES.index('tmin')] = np.array([np.nan, np.nan])
    got = MOD16._et(vec, *drv)
    parity.assert_parity(got, np.broadcast_to(z['gate_nan_et'], got.shape), np.float64,
                         what='gate nan')


@pytest.mark.parametrize('name', ['rand_phys_f64', 'rand_phys_fill_f64',
                                  'rand_stress_fill_f64', 'rand_phys_f32',
                                  'rand_stress_fill_f32'])
def test_random_golden_fixtures(cuda_lib, golden_dir, name):
    z = np.load(os.path.join(golden_dir, name + '.npz'))
    dtype = np.float32 if name.endswith('f32') else np.float64
    drv = [z['drv_' + k] if 'drv_' + k in z else 0 for k in NAMES]
    pft = z['pft']
    vec = [z['table'][i][pft].reshape(1, -1).astype(dtype) for i in range(11)]
    d, n = MOD16._evapotranspiration(vec, *drv, dtype=dtype)
    parity.assert_parity(d, z['static_day'], dtype, what='static day')
    parity.assert_parity(n, z['static_night'], dtype, what='static night')
    et = MOD16._et_by_pft(z['table'], pft, *drv, dtype=dtype)
    parity.assert_parity(et, z['static_et'], dtype, what='_et by pft')
    model = MOD16(dict(zip(MOD16.required_parameters, z['inst_params'])))
    day, night = model.evapotranspiration(*drv, separate=True, dtype=dtype)
    for hn, got in (('day', day), ('night', night)):
        for c, cn in enumerate(('canopy', 'soil', 'trans')):
            parity.assert_parity(got[c], z['inst_%s_%s' % (hn, cn)], dtype,
                                 parity.MASS_SCALE, what='inst %s %s' % (hn, cn))
    d2, n2 = model.evapotranspiration(*drv, dtype=dtype)
    parity.assert_parity(d2, (z['inst_day_canopy'] + z['inst_day_soil']) + z['inst_day_trans'],
                         dtype, parity.MASS_SCALE, what='inst day total')


@pytest.mark.parametrize('dtype,kind,nan', [
    (np.float64, 'physical', 0.0), (np.float64, 'stress', 0.002),
    (np.float32, 'physical', 0.0), (np.float32, 'stress', 0.002),
    (np.float32, 'physical', 0.002)])
def test_medium_grid_against_oracle(cuda_lib, dtype, kind, nan):
    'A (T, N) block big enough to exercise the vectorised layout, tails and many blocks.'
    T, N = 37, 4 * 1031
    d = synthetic.host_drivers(T, N, 4242, kind=kind, dtype=dtype, nan_frac=nan,
                               fill_frac=nan)
    drv = [d[k] for k in NAMES]
    pft = synthetic.host_pft_map(N, 7)
    table = utils.bplut_table(utils.restore_bplut(utils.BPLUT_C51),
                              MOD16.required_parameters, default_beta=250.0)
    vec = [table[i][pft].reshape(1, N) for i in range(11)]
    with np.errstate(all='ignore'):
        ref = orc.static_et(vec, *as_f64(drv))
    got = MOD16._et_by_pft(table, pft, *drv, dtype=dtype)
    m = parity.assert_parity(got, ref, dtype, what='%s %s' % (kind, np.dtype(dtype).name),
                             allow_fail=0)
    print('medium grid', np.dtype(dtype).name, kind, m)


@pytest.mark.parametrize('dtype', [np.float32, np.float64])
def test_masked_pixels(cuda_lib, dtype):
    '''
    A tile with masked pixels (NaN fPAR / LAI / albedo at every step, NaN pressure on a few,
    single NaNs elsewhere): the float32 kernels answer NaN totals from the masked-driver rule
    without the double code; masks and values must still be the reference's, for every
    entry point and output mode.
    '''
    T, N = 23, 4 * 977
    rng = np.random.default_rng(77)
    d = synthetic.host_drivers(T, N, 606, dtype=dtype)
    masked = rng.random(N) < 0.3
    for k in ('fpar', 'lai', 'sw_albedo'):
        d[k][:, masked] = np.nan
    d['pressure'][rng.random(N) < 0.02] = np.nan
    for k in ('lw_net_night', 'temp_day', 'vpd_night', 'tmin', 'lai'):
        d[k][rng.random((T, N)) < 0.003] = np.nan
    d['temp_night'][rng.random((T, N)) < 0.002] = -9999.0
    drv = [d[k] for k in NAMES]
    d64 = as_f64(drv)
    pft = synthetic.host_pft_map(N, 8)
    table = utils.bplut_table(utils.restore_bplut(utils.BPLUT_C51),
                              MOD16.required_parameters, default_beta=250.0)
    vec = [table[i][pft].reshape(1, N) for i in range(11)]
    with np.errstate(all='ignore'):
        rd, rn = orc.static_evapotranspiration(vec, *d64)
    got = MOD16._et_by_pft(table, pft, *drv, dtype=dtype)
    m = parity.assert_parity(got, rd + rn, dtype, what='masked _et')
    assert 0.3 < np.isnan(got).mean() < 0.4
    gd, gn = MOD16._evapotranspiration(vec, *drv, dtype=dtype)
    parity.assert_parity(gd, rd, dtype, what='masked day')
    parity.assert_parity(gn, rn, dtype, what='masked night')
    col = list(table[:, 7])
    model = MOD16(dict(zip(MOD16.required_parameters, col)))
    with np.errstate(all='ignore'):
        ref = orc.InstanceModel(dict(zip(orc.PARAM_ORDER, col))).evapotranspiration(*d64)
        sep = orc.InstanceModel(dict(zip(orc.PARAM_ORDER, col))).evapotranspiration(
            *d64, separate=True)
    day, night = model.evapotranspiration(*drv, dtype=dtype)
    parity.assert_parity(day, ref[0], dtype, parity.MASS_SCALE, what='masked inst day')
    parity.assert_parity(night, ref[1], dtype, parity.MASS_SCALE, what='masked inst night')
    cd, cn = model.evapotranspiration(*drv, separate=True, dtype=dtype)
    for c in range(3):
        parity.assert_parity(cd[c], sep[0][c], dtype, parity.MASS_SCALE, what='masked comp d')
        parity.assert_parity(cn[c], sep[1][c], dtype, parity.MASS_SCALE, what='masked comp n')
    # ragged N / strided kernel
    sl = [v[:, :1001] if isinstance(v, np.ndarray) and v.ndim == 2 else
          (v[:1001] if isinstance(v, np.ndarray) else v) for v in drv]
    got = MOD16._et([v[:, :1001] for v in vec], *sl, dtype=dtype)
    parity.assert_parity(got, (rd + rn)[:, :1001], dtype, what='masked strided')


def test_input_kinds_and_broadcasting(cuda_lib):
    'python scalars, (N,), (1, N), (T, 1), (T, N), int16 zeros, torch CUDA tensors.'
    T, N = 5, 64
    d = synthetic.host_drivers(T, N, 99)
    drv = [d[k] for k in NAMES]
    drv[NAMES.index('sw_rad_night')] = np.zeros((T, N), np.int16)   # calibration.py:383
    drv[NAMES.index('sw_albedo')] = d['sw_albedo'][:, :1]            # (T, 1)
    drv[NAMES.index('pressure')] = d['pressure'].reshape(1, N)       # (1, N)
    drv[NAMES.index('fpar')] = 0.4                                   # python scalar
    vec = list(BPLUT7)
    vec[4] = np.linspace(0.01, 0.04, N).reshape(1, N)                # mixed scalar/(1, N)
    with np.errstate(all='ignore'):
        ref = orc.static_et(vec, *drv)
    got = MOD16._et(vec, *drv)
    assert isinstance(got, np.ndarray) and got.shape == (T, N) and got.dtype == np.float64
    parity.assert_parity(got, ref, np.float64, what='mixed shapes')
    # 1-D masked vectors as sensitivity.py:179-185 builds them
    flat = [v.ravel() if isinstance(v, np.ndarray) and v.shape == (T, N) else v for v in
            [d[k] for k in NAMES]]
    flat[NAMES.index('temp_annual')] = np.repeat(d['temp_annual'][None], T, 0).ravel()
    flat[NAMES.index('pressure')] = np.repeat(d['pressure'][None], T, 0).ravel()
    with np.errstate(all='ignore'):
        ref1 = orc.static_et(BPLUT7, *flat)
    got1 = MOD16._et(BPLUT7, *flat)
    assert got1.shape == (T * N,)
    parity.assert_parity(got1, ref1, np.float64, what='1-D')
    # CUDA tensors in, CUDA tensor out, no host round trip
    tdrv = [torch.from_numpy(v).cuda() if isinstance(v, np.ndarray) else v
            for v in [d[k] for k in NAMES]]
    got2 = MOD16._et(BPLUT7, *tdrv)
    assert isinstance(got2, torch.Tensor) and got2.is_cuda and got2.dtype == torch.float64
    with np.errstate(all='ignore'):
        ref2 = orc.static_et(BPLUT7, *[d[k] for k in NAMES])
    parity.assert_parity(got2.cpu().numpy(), ref2, np.float64, what='cuda tensors')
    # float32 inputs + python-float params => float32 compute, like numpy promotion
    f32 = [v.astype(np.float32) if isinstance(v, np.ndarray) else v for v in [d[k] for k in NAMES]]
    got3 = MOD16._et(BPLUT7, *f32)
    assert got3.dtype == np.float32
    parity.assert_parity(got3, orc.static_et(BPLUT7, *as_f64(f32)), np.float32, what='f32')
    # non-contiguous (transposed) storage
    tr = [np.asfortranarray(v) if isinstance(v, np.ndarray) and v.ndim == 2 else v
          for v in [d[k] for k in NAMES]]
    parity.assert_parity(MOD16._et(BPLUT7, *tr), ref2, np.float64, what='fortran order')


def test_python_sequences_as_inputs(cuda_lib):
    'Lists and tuples broadcast like the arrays numpy would make of them.'
    p7 = [-8., 8.8, 650., 4400., .02, .02, 1e-5, .0055, 60., 95., 250.]
    drv = [-50, -30, [150, 160, 170], 0, 0.3, (293, 294, 295), 285, 285, 285, 1000, 500, 1e5,
           0.5, [1.5, 1.0, 0.5]]
    arr = [np.asarray(v, np.float64) if isinstance(v, (list, tuple)) else v for v in drv]
    with np.errstate(all='ignore'):
        ref = orc.static_et(p7, *arr)
    got = MOD16._et(p7, *drv)
    assert got.shape == (3,)
    parity.assert_parity(got, ref, np.float64, what='list inputs')


def test_empty_and_ragged(cuda_lib):
    d = synthetic.host_drivers(3, 0, 1)
    got = MOD16._et(BPLUT7, *[d[k] for k in NAMES])
    assert got.shape == (3, 0)
    for N in (1, 3, 5, 127, 129):   # not multiples of the vector width
        d = synthetic.host_drivers(2, N, N, dtype=np.float32)
        drv = [d[k] for k in NAMES]
        got = MOD16._et(BPLUT7, *drv, dtype=np.float32)
        parity.assert_parity(got, orc.static_et(BPLUT7, *as_f64(drv)), np.float32,
                             what='N=%d' % N)


def test_r_corr_list(cuda_lib):
    d = synthetic.host_drivers(4, 40, 11)
    drv = [d[k] for k in NAMES]
    rc = [orc.r_correction(d['pressure'], d['temp_day']) * 1.1,
          orc.r_correction(d['pressure'], d['temp_night']) * 0.9]
    with np.errstate(all='ignore'):
        ref = orc.static_et(BPLUT7, *drv, r_corr_list=rc)
    parity.assert_parity(MOD16._et(BPLUT7, *drv, r_corr_list=rc), ref, np.float64)
    f32 = [v.astype(np.float32) if isinstance(v, np.ndarray) else v for v in drv]
    rc32 = [r.astype(np.float32) for r in rc]
    with np.errstate(all='ignore'):
        ref32 = orc.static_et(BPLUT7, *as_f64(f32), r_corr_list=as_f64(rc32))
    parity.assert_parity(MOD16._et(BPLUT7, *f32, r_corr_list=rc32), ref32, np.float32)


def test_errors_are_loud(cuda_lib):
    with pytest.raises(TypeError):
        MOD16._et(BPLUT7[:5], *([1.0] * 14))
    with pytest.raises(ValueError):
        MOD16._et(BPLUT7, *([np.zeros((2, 3))] + [np.zeros((2, 4))] * 13))
    with pytest.raises(_abi.Mod16Error):
        _abi.load('/nonexistent/libmod16_b200.so')


def test_large_host_arrays_take_the_threaded_staging_path(cuda_lib):
    'Plain numpy inputs above the staging threshold: same bits as the device-resident call.'
    from mod16_b200 import hostcopy
    T, N = 5, 4 * 1024 * 1024 + 4          # 80 MB per driver, ragged chunk at the end
    dev = torch.device('cuda', 0)
    d = synthetic.device_drivers(T, N, 31, dev)
    d['lai'][2, 1000:2000] = float('nan')
    drv_dev = [d[k] for k in NAMES]
    host = [v.cpu().numpy() if isinstance(v, torch.Tensor) else v for v in drv_dev]
    assert hostcopy.worthwhile(host[0])
    p7 = [-8., 8.8, 650., 4400., .02, .02, 1e-5, .0055, 60., 95., 250.]
    ref = MOD16._et(p7, *drv_dev, dtype=np.float32).cpu().numpy()
    got = MOD16._et(p7, *host, dtype=np.float32)
    assert isinstance(got, np.ndarray) and got.shape == (T, N)
    assert np.array_equal(got.view(np.int32), ref.view(np.int32))
    # float64 arrays headed for the float32 path are converted on the device
    host64 = [v.astype(np.float64) if isinstance(v, np.ndarray) else v for v in host]
    got64 = MOD16._et(p7, *host64, dtype=np.float32)
    assert got64.dtype == np.float32
    assert np.array_equal(got64.view(np.int32), ref.view(np.int32))
    # round trip of the staging helpers alone, float64, non-multiple of the chunk size
    a = np.random.default_rng(0).normal(size=(3, 3_000_001))
    s = torch.cuda.current_stream(dev)
    t = hostcopy.upload([a], dev, s)[0]
    assert np.array_equal(hostcopy.download(t, s), a)


@pytest.mark.parametrize('N', [480 * 3, 480 * 3 + 4, 1001, 4])
def test_outputs_stay_inside_their_buffers(cuda_lib, N):
    'Canaries around every output of K1 (bulk-copy and strided kernels, ragged last tile).'
    from mod16_b200.forward import run_forward
    T = 7
    dev = torch.device('cuda', 0)
    d = synthetic.device_drivers(T, N, 9, dev)
    d['lai'][3, N // 2] = float('nan')
    drv = [d[k] for k in NAMES]
    p7 = [-8., 8.8, 650., 4400., .02, .02, 1e-5, .0055, 60., 95., 250.]
    pad = 1024
    for sem, mode, n_out in (('static', 'sum', 1), ('static', 'day_night', 2),
                             ('instance', 'components', 6)):
        big = torch.full((n_out, T * N + 2 * pad), -12345.0, device=dev)
        outs = [big[k, pad:pad + T * N].view(T, N) for k in range(n_out)]
        run_forward(drv, params=p7, semantics=sem, out_mode=mode, dtype=np.float32, device=dev,
                    out=outs)
        torch.cuda.synchronize()
        assert bool((big[:, :pad] == -12345.0).all()) and bool((big[:, pad + T * N:] == -12345.0).all())
        assert not bool((big[:, pad:pad + T * N] == -12345.0).any())


@pytest.mark.parametrize('dtype', [np.float32, np.float64])
def test_dense_sw_rad_night(cuda_lib, dtype):
    '''
    Every reference caller passes sw_rad_night = 0, but the signature takes an array: the
    twelfth stream of the bulk-copy kernel (one stage fewer), odd T, masked values.
    '''
    T, N = 11, 4 * 700
    rng = np.random.default_rng(12)
    d = synthetic.host_drivers(T, N, 808, dtype=dtype, sw_rad_night='array')
    d['sw_rad_night'] = rng.uniform(0, 40, (T, N)).astype(dtype)   # twilight: make it matter
    d['sw_rad_night'][rng.random((T, N)) < 0.01] = np.nan
    d['fpar'][:, ::7] = np.nan
    drv = [d[k] for k in NAMES]
    d64 = as_f64(drv)
    with np.errstate(all='ignore'):
        rd, rn = orc.static_evapotranspiration(BPLUT7, *d64)
        inst = orc.InstanceModel(dict(zip(orc.PARAM_ORDER, BPLUT7))).evapotranspiration(*d64)
    gd, gn = MOD16._evapotranspiration(BPLUT7, *drv, dtype=dtype)
    parity.assert_parity(gd, rd, dtype, what='swn day')
    parity.assert_parity(gn, rn, dtype, what='swn night')
    parity.assert_parity(MOD16._et(BPLUT7, *drv, dtype=dtype), rd + rn, dtype, what='swn sum')
    day, night = MOD16(dict(zip(MOD16.required_parameters, BPLUT7))).evapotranspiration(
        *drv, dtype=dtype)
    parity.assert_parity(day, inst[0], dtype, parity.MASS_SCALE, what='swn inst day')
    parity.assert_parity(night, inst[1], dtype, parity.MASS_SCALE, what='swn inst night')


def test_streamed_forward_from_pinned_host_buffers(cuda_lib):
    'The end-to-end path of bench.py: double-buffered slabs, ragged last slab, bit-identical.'
    from mod16_b200.streaming import StreamedForward
    T, N = 11, 4 * 5000
    dev = torch.device('cuda', 0)
    d = synthetic.host_drivers(T, N, 515, dtype=np.float32, nan_frac=0.002)
    table = utils.bplut_table(utils.restore_bplut(utils.BPLUT_C51),
                              MOD16.required_parameters, default_beta=250.0)
    pft = synthetic.host_pft_map(N, 9)
    host = dict((k, torch.from_numpy(v).pin_memory() if isinstance(v, np.ndarray) else v)
                for k, v in d.items())
    out_host = torch.empty((T, N), dtype=torch.float32).pin_memory()
    runner = StreamedForward(N, 4, table, pft, device=dev, dtype=np.float32, gate='open')
    runner.run(host, out_host)
    ref = MOD16._et_by_pft(table, pft, *[d[k] for k in NAMES], dtype=np.float32, gate='open')
    assert np.array_equal(out_host.numpy().view(np.int32), ref.view(np.int32))
    assert runner.h2d_bytes == 11 * T * N * 4 + 2 * N * 4 and runner.d2h_bytes == T * N * 4
    runner.run(host, out_host)          # buffers are reusable
    assert np.array_equal(out_host.numpy().view(np.int32), ref.view(np.int32))


def test_global_gate_helper_and_mirror_flag(cuda_lib):
    from mod16_b200 import distributed as D
    T, N = 6, 257
    d = synthetic.host_drivers(T, N, 21)
    d['tmin'] = np.full((T, N), 230.0)           # nobody transpires: the gate is closed
    drv = [d[k] for k in NAMES]
    with np.errstate(all='ignore'):
        ref = orc.static_et(BPLUT7, *drv)
        ref_open = orc.static_et(BPLUT7, *drv, gate='open')
    assert np.abs(ref - ref_open).max() > 1e-3
    out = D.forward_global_gate(drv, params=BPLUT7)[0].cpu().numpy()   # world size 1
    parity.assert_parity(out, ref, np.float64, what='open_track + fixup')
    # the float32 op-for-op kernel (testing flag): float32 arithmetic, so only close
    d32 = [v.astype(np.float32) if isinstance(v, np.ndarray) else v for v in drv]
    got = MOD16._et(BPLUT7, *d32, dtype=np.float32, force=_abi.FORCE_MIRROR)
    assert np.allclose(got, ref, rtol=2e-3, atol=2e-2)


def test_double_queue_overflow_takes_the_sweep(cuda_lib):
    '''
    More out-of-domain elements than the queue holds (unmasked fill values, negative VPD): the
    main kernel leaves markers, the sweep kernel evaluates them; results are the oracle's.
    '''
    from mod16_b200.forward import run_forward
    T, N = 12, 4 * 480 * 3
    d = synthetic.host_drivers(T, N, 99, dtype=np.float32)
    rng = np.random.default_rng(5)
    d['vpd_day'][rng.random((T, N)) < 0.05] = -40.0
    d['lai'][:, ::11] = -9999.0
    d['tmin'][rng.random((T, N)) < 0.01] = np.nan
    drv = [d[k] for k in NAMES]
    with np.errstate(all='ignore'):
        rd, rn = orc.static_evapotranspiration(BPLUT7, *as_f64(drv))
    tiny = torch.zeros(1 + 64, dtype=torch.int64, device='cuda')     # room for 64 entries
    for mode, ref in (('sum', [rd + rn]), ('day_night', [rd, rn])):
        got = run_forward(drv, params=BPLUT7, semantics='static', out_mode=mode,
                          dtype=np.float32, workspace=tiny)
        for g, r in zip(got, ref):
            parity.assert_parity(g, r, np.float32, what='overflow ' + mode)
    with np.errstate(all='ignore'):
        ro = orc.static_et(BPLUT7, *as_f64(drv), gate='open')
    got = run_forward(drv, params=BPLUT7, semantics='static', out_mode='sum', gate='open',
                      dtype=np.float32, workspace=tiny)[0]
    assert int(tiny[0].item()) > 64                                   # it did overflow
    parity.assert_parity(got, ro, np.float32, what='overflow open')
    # (no workspace at all: block-local queue, in-place evaluation beyond it)
    none = torch.zeros(1, dtype=torch.int64, device='cuda')
    got = run_forward(drv, params=BPLUT7, semantics='static', out_mode='sum', dtype=np.float32,
                      workspace=none)[0]
    parity.assert_parity(got, rd + rn, np.float32, what='no workspace')


@pytest.mark.gpu
@pytest.mark.parametrize('dtype', [np.float32, np.float64])
def test_land_cover_codes_outside_the_table(cuda_lib, dtype):
    '''
    A land-cover code above 12 (254 = water, 255 = fill in MCD12Q1): numpy's fancy index
    `params[key][pft_map]` raises IndexError in the reference's callers
    (Forward_Run_example.ipynb cell 32); the kernels cannot raise per pixel and answer NaN --
    the documented deviation.  Codes 0 and 11 gather the BPLUT's own NaN columns.
    '''
    T, N = 6, 960
    d = synthetic.host_drivers(T, N, 171, dtype=dtype)
    drv = [d[k] for k in NAMES]
    table = utils.bplut_table(utils.restore_bplut(utils.BPLUT_C51), MOD16.required_parameters,
                              default_beta=250.0)
    pft = synthetic.host_pft_map(N, 172)
    pft[5], pft[6], pft[100], pft[481], pft[959] = 13, 254, 255, 0, 11
    got = MOD16._et_by_pft(table, pft, *drv, dtype=dtype)
    bad = np.zeros(N, bool)
    bad[[5, 6, 100, 481, 959]] = True
    assert np.isnan(got[:, bad]).all() and not np.isnan(got[:, ~bad]).any()
    ok = np.flatnonzero(~bad)
    vec = [table[i][pft[ok]].reshape(1, -1) for i in range(11)]
    d64 = [v[..., ok].astype(np.float64) if isinstance(v, np.ndarray) else v for v in drv]
    with np.errstate(all='ignore'):
        ref = orc.static_et(vec, *d64)
    parity.assert_parity(got[:, ok], ref, dtype, what='codes inside the table')
    with pytest.raises(IndexError):
        [table[i][pft] for i in range(11)]


@pytest.mark.gpu
def test_stock_promotion_float32_drivers_with_float64_parameters(cuda_lib):
    '''
    The reference's stock call: float32 drivers from the archive + the float64 BPLUT, which
    numpy promotes to float64.  The float64 entry reads the float32 arrays as they are
    (`mod16_drivers_t.f32_mask`); results equal the call on float64 copies bit for bit.
    '''
    T, N = 9, 2052
    d = synthetic.host_drivers(T, N, 181, dtype=np.float32, nan_frac=0.01, fill_frac=0.005)
    drv = [d[k] for k in NAMES]
    table = utils.bplut_table(utils.restore_bplut(utils.BPLUT_C51), MOD16.required_parameters,
                              default_beta=250.0)
    pft = synthetic.host_pft_map(N, 182)
    vec = [table[i][pft].reshape(1, N) for i in range(11)]          # float64
    got = MOD16._et(vec, *drv)                                      # promotes: float64
    assert got.dtype == np.float64
    assert cuda_lib.mod16_last_forward_kernel() & 0xff == _abi.KERNEL_FAST64
    d64 = [v.astype(np.float64) if isinstance(v, np.ndarray) else v for v in drv]
    same = MOD16._et(vec, *d64)
    assert np.array_equal(got, same, equal_nan=True)
    with np.errstate(all='ignore'):
        ref = orc.static_et(vec, *d64)
    parity.assert_parity(got, ref, np.float64, what='stock promotion')
    # device tensors of mixed precision, instance semantics, the op-for-op kernel as well
    dev = torch.device('cuda', 0)
    mixed = [torch.from_numpy(v).to(dev) if isinstance(v, np.ndarray) else v for v in drv]
    mixed[4] = mixed[4].double()
    col = dict(zip(orc.PARAM_ORDER, table[:, 4]))
    day, night = MOD16(col).evapotranspiration(*mixed, dtype=np.float64)
    with np.errstate(all='ignore'):
        rd, rn = orc.InstanceModel(col).evapotranspiration(*d64)
    parity.assert_parity(day.cpu().numpy(), rd, np.float64, parity.MASS_SCALE, what='mixed day')
    parity.assert_parity(night.cpu().numpy(), rn, np.float64, parity.MASS_SCALE, what='mixed night')
    from mod16_b200.forward import run_forward
    mirror = run_forward(drv, params=vec, dtype=np.float64, force=_abi.FORCE_MIRROR)[0]
    parity.assert_parity(mirror, ref, np.float64, what='op-for-op kernel on float32 arrays')
