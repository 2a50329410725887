'''
GPU tests of the call forms the reference's own callers use, and of which kernel they reach
(mod16_last_forward_kernel()):

* `MOD16._et([params[k][pft_map].reshape((1, N)) ...], *drivers)` -- per-pixel parameter
  arrays, notebooks/MOD16_Forward_Run_example.ipynb cells 32 / 35 -- must run on the
  bulk-copy pipeline like the land-cover gather does;
* (T, N) `temp_annual` / `pressure` (the annual means of mod16/calibration.py:556-560
  broadcast to the drivers' shape);
* a pixel range of a wider array (what a rank of a pixel-sharded run holds);
* negative VPD and negative (fill) temperatures, which the fast path now decides itself;
* the T-chunked plan of mod16_b200/chunked.py against one call on the whole block;
* the NCCL code paths of mod16_b200/distributed.py in a world of one rank.
'''
import os
import socket

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


def _d64(drv):
    return [v.astype(np.float64) if isinstance(v, np.ndarray) else v for v in drv]


def _kernel(lib):
    k = lib.mod16_last_forward_kernel()
    return k & 0xff, k >> 8


@pytest.mark.parametrize('N', [4096, 4100, 484])
def test_reference_call_form_reaches_the_pipeline(cuda_lib, N):
    T = 21
    d = synthetic.host_drivers(T, N, 31, dtype=np.float32)
    drv = [d[k] for k in NAMES]
    table = _table()
    pft = synthetic.host_pft_map(N, 32)
    vec = [table[i][pft].reshape(1, N) for i in range(11)]
    with np.errstate(all='ignore'):
        ref = orc.static_et(vec, *_d64(drv))
    # float64 parameter arrays + float32 drivers promote like numpy: ask for float32
    got = MOD16._et(vec, *drv, dtype=np.float32)
    assert _kernel(cuda_lib) == (_abi.KERNEL_PIPE, 0)
    parity.assert_parity(got, ref, np.float32, what='per-pixel parameter arrays')
    by_pft = MOD16._et_by_pft(table, pft, *drv, dtype=np.float32)
    assert _kernel(cuda_lib) == (_abi.KERNEL_PIPE, 0)
    # the float32 copies of the parameters differ from the table's doubles in the last bit
    assert np.allclose(got, by_pft, rtol=2e-5, atol=1e-3)
    # instance semantics, day / night
    vec32 = [v.astype(np.float32) for v in vec]
    run_forward(drv, params=vec32, semantics='instance', out_mode='day_night')
    assert _kernel(cuda_lib) == (_abi.KERNEL_PIPE, 0)


def test_time_varying_annual_temperature_and_pressure(cuda_lib):
    T, N = 37, 2400
    d = synthetic.host_drivers(T, N, 41, dtype=np.float32)
    rng = np.random.default_rng(42)
    # (T, N) annual means: a few distinct years, as calibration.py:556-560 builds them
    year = np.repeat(np.arange(4), 10)[:T]
    d['temp_annual'] = (d['temp_annual'][None, :] + year[:, None]).astype(np.float32)
    d['pressure'] = (d['pressure'][None, :] * (1 + 1e-3 * year[:, None])).astype(np.float32)
    d['temp_annual'][rng.random((T, N)) < 0.01] = np.nan    # masked-away NaN (:230-233)
    drv = [d[k] for k in NAMES]
    col = list(_table()[:, 4])
    with np.errstate(all='ignore'):
        ref = orc.static_et(col, *_d64(drv))
        rd, rn = orc.static_evapotranspiration(col, *_d64(drv))
    got = MOD16._et(col, *drv, dtype=np.float32)
    assert _kernel(cuda_lib) == (_abi.KERNEL_PIPE, 2)
    parity.assert_parity(got, ref, np.float32, what='(T, N) annual temperature / pressure')
    gd, gn = MOD16._evapotranspiration(col, *drv, dtype=np.float32)
    assert _kernel(cuda_lib) == (_abi.KERNEL_PIPE, 2)
    parity.assert_parity(gd, rd, np.float32, what='day')
    parity.assert_parity(gn, rn, np.float32, what='night')
    # with a dense sw_rad_night as well: all fourteen arrays streamed
    d['sw_rad_night'] = rng.uniform(0, 1e-5, (T, N)).astype(np.float32)
    drv = [d[k] for k in NAMES]
    with np.errstate(all='ignore'):
        ref = orc.static_et(col, *_d64(drv))
    got = MOD16._et(col, *drv, dtype=np.float32)
    assert _kernel(cuda_lib) == (_abi.KERNEL_PIPE, 3)
    parity.assert_parity(got, ref, np.float32, what='fourteen streamed arrays')


def test_pixel_range_of_a_wider_array_is_streamed(cuda_lib):
    dev = torch.device('cuda', 0)
    T, N = 12, 9600
    d = synthetic.device_drivers(T, N, 51, dev)
    lst = [d[k] for k in NAMES]
    table = _table()
    pft = torch.from_numpy(synthetic.host_pft_map(N, 52)).to(dev)
    full = run_forward(lst, pft=pft, table=table, gate='open', dtype=np.float32, device=dev,
                       as_torch=True)[0]
    for lo, hi, want in ((0, 4800, _abi.KERNEL_PIPE), (484, 9600, _abi.KERNEL_PIPE),
                         (486, 9590, _abi.KERNEL_STRIDED)):
        sl = [v[:, lo:hi] if isinstance(v, torch.Tensor) and v.ndim == 2 else
              (v[lo:hi] if isinstance(v, torch.Tensor) else v) for v in lst]
        part = run_forward(sl, pft=pft[lo:hi], table=table, gate='open', dtype=np.float32,
                           device=dev, as_torch=True)[0]
        assert _kernel(cuda_lib)[0] == want, (lo, hi)
        assert torch.equal(part.view(torch.int32), full[:, lo:hi].contiguous().view(torch.int32))


@pytest.mark.parametrize('semantics', ['static', 'instance'])
@pytest.mark.parametrize('out_mode', ['sum', 'day_night', 'components'])
def test_negative_vpd_and_fill_temperatures(cuda_lib, semantics, out_mode):
    T, N = 16, 4800
    d = synthetic.host_drivers(T, N, 61, dtype=np.float32)
    rng = np.random.default_rng(62)
    for k in ('vpd_day', 'vpd_night'):           # reanalysis dew point above air temperature
        m = rng.random((T, N)) < 0.05
        d[k][m] = -rng.uniform(0, 300, m.sum()).astype(np.float32)
    for k, fill in (('temp_day', -9999.0), ('temp_night', -32768.0)):
        d[k][rng.random((T, N)) < 0.01] = fill
    d['tmin'][rng.random((T, N)) < 0.01] = -9999.0
    drv = [d[k] for k in NAMES]
    col = list(_table()[:, 7])
    with np.errstate(all='ignore'):
        if semantics == 'static':
            day, night = orc.static_evapotranspiration(col, *_d64(drv))
            comps = orc.static_evapotranspiration(col, *_d64(drv), components=True)
        else:
            comps = orc.InstanceModel(dict(zip(orc.PARAM_ORDER, col))).evapotranspiration(
                *_d64(drv), separate=True)
            day = comps[0][0] + comps[0][1] + comps[0][2]
            night = comps[1][0] + comps[1][1] + comps[1][2]
    scale = 1.0 if semantics == 'static' else parity.MASS_SCALE
    got = run_forward(drv, params=col, semantics=semantics, out_mode=out_mode, dtype=np.float32)
    assert _kernel(cuda_lib)[0] == _abi.KERNEL_PIPE
    if out_mode == 'sum':
        parity.assert_parity(got[0], day + night, np.float32, scale, what='sum')
    elif out_mode == 'day_night':
        parity.assert_parity(got[0], day, np.float32, scale, what='day')
        parity.assert_parity(got[1], night, np.float32, scale, what='night')
    else:
        for h in range(2):
            for c in range(3):
                parity.assert_parity(got[3 * h + c], comps[h][c], np.float32, scale,
                                     what='component %d %d' % (h, c))
    if out_mode != 'components':   # the fill temperatures came out as NaN: 1 % of the days
        assert np.isnan(got[0]).mean() > 0.008


def test_chunked_plan_equals_one_call(cuda_lib):
    from mod16_b200.chunked import ChunkedForward, PER_STEP
    dev = torch.device('cuda', 0)
    T, N = 23, 4804
    d = synthetic.device_drivers(T, N, 71, dev)
    table = _table()
    pft = torch.from_numpy(synthetic.host_pft_map(N, 72)).to(dev)
    lst = [d[k] for k in NAMES]
    for closed in (False, True):
        if closed:
            d['tmin'].fill_(230.0)       # nobody transpires: the whole call's gate is closed
        flag = torch.zeros(1, device=dev, dtype=torch.int32)
        one = run_forward(lst, pft=pft, table=table, gate='global', dtype=np.float32,
                          device=dev, gate_flag=flag, as_torch=True)[0]
        cf = ChunkedForward(T, N, table, pft, d['temp_annual'], d['pressure'], device=dev,
                            budget_bytes=2 * len(PER_STEP) * N * 4 * 6)
        assert len(cf.chunks) == 4 and len(cf.bufs) == 2
        calls = []

        def fill(k, t0, t1, bufs):
            calls.append(k)
            for name in PER_STEP:
                bufs[name].copy_(d[name][t0:t1])
        out = torch.full((T, N), -1.0, device=dev)
        cf.run(fill, out)
        assert torch.equal(out.view(torch.int32), one.view(torch.int32))
        assert len(calls) == (8 if closed else 4) and cf.kernel_ms > 0


def test_nccl_world_of_one(cuda_lib):
    'The NCCL branches of distributed.py / chunked.py (all-reduce, all-gather) on one rank.'
    import torch.distributed as dist
    from mod16_b200 import distributed as D
    from mod16_b200.objective import ObjectiveBackend
    dev = torch.device('cuda', 0)
    with socket.socket() as s:
        s.bind(('127.0.0.1', 0))
        port = s.getsockname()[1]
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('nccl', rank=0, world_size=1, device_id=dev)
    try:
        T, N = 9, 960
        d = synthetic.device_drivers(T, N, 81, dev)
        lst = [d[k] for k in NAMES]
        col = list(_table()[:, 7])
        flag = torch.zeros(1, device=dev, dtype=torch.int32)
        one = run_forward(lst, params=col, gate='global', dtype=np.float32, device=dev,
                          gate_flag=flag, as_torch=True)[0]
        # D.world() sees an initialised group of size 1: force the collective branch
        size = D.world
        D.world = lambda group=None: (0, 2)
        try:
            got = D.forward_global_gate(lst, params=col, dtype=np.float32, device=dev)[0]
            assert torch.equal(got.view(torch.int32), one.view(torch.int32))
            obs = one + 1.0
            backend = ObjectiveBackend(lst, obs, None, dtype=np.float32, device=dev)
            draws = torch.from_numpy(synthetic.host_param_draws(
                40, 5, bplut_column=np.array(col))).to(dev)
            s0, c0 = backend(draws, as_torch=True)
            s1, c1 = D.objective_site_sharded(backend, draws)
            assert torch.equal(s0, s1) and torch.equal(c0, c1)
            t = torch.arange(6, dtype=torch.float64, device=dev)
            a, b = D.reduce_objective(t.clone(), t.clone() + 1)
            assert torch.equal(a, t) and torch.equal(b, t + 1)
        finally:
            D.world = size
        s2, c2 = D.objective_draw_sharded(backend, draws)
        assert torch.equal(s0, s2) and torch.equal(c0, c2)
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize('semantics', ['static', 'instance'])
def test_potential_terms_from_the_same_pass(cuda_lib, semantics):
    '''
    MOD16_OUT_POTENTIAL: the total and the potential terms of
    `MOD16.potential_transpiration` (:547-602) and `MOD16.potential_soil_evaporation`
    (:450-545) as outputs of the fused kernel, against the reference's own functions
    (restated in the oracle, golden-pinned in components.npz) on the same inputs.
    '''
    T, N = 18, 4800
    d = synthetic.host_drivers(T, N, 181, dtype=np.float32)
    rng = np.random.default_rng(182)
    d['lai'][rng.random((T, N)) < 0.01] = np.nan           # the queue takes part
    drv = [d[k] for k in NAMES]
    d64 = dict((k, v.astype(np.float64) if isinstance(v, np.ndarray) else v) for k, v in d.items())
    col = list(_table()[:, 7])
    p = dict(zip(orc.PARAM_ORDER, col))
    M = orc.InstanceModel
    with np.errstate(all='ignore'):
        rad_day = d64['sw_rad_day'] * (1 - d64['sw_albedo']) + d64['lw_net_day']
        rad_night = d64['lw_net_night']
        lo = 273.15 + (p['tmin_open'] if semantics == 'static' else p['tmin_close'])
        ta = d64['temp_annual']
        cond = (ta < 273.15 + 25) & ((ta > lo) if semantics == 'static' else (ta >= lo)) & \
            ((d64['temp_day'] - d64['temp_night']) >= 5)
        g = orc._ground_heat(rad_day, rad_night, d64['temp_day'], d64['temp_night'], cond)
        rs = orc._soil_radiation(rad_day, rad_night, g[0], g[1], d64['fpar'])
        kw = dict(vpd_open=p['vpd_open'], vpd_close=p['vpd_close'], rbl_min=p['rbl_min'],
                  rbl_max=p['rbl_max'])
        sat_d, unsat_d = M.potential_soil_evaporation(
            d64['pressure'], d64['temp_day'], d64['vpd_day'], d64['fpar'], rs[0], **kw)
        sat_n, unsat_n = M.potential_soil_evaporation(
            d64['pressure'], d64['temp_night'], d64['vpd_night'], d64['fpar'], rs[1], **kw)
        pt = M.potential_transpiration(d64['lw_net_day'], d64['sw_rad_day'], d64['sw_albedo'],
                                       d64['pressure'], d64['temp_day'], d64['vpd_day'],
                                       d64['fpar'])
        if semantics == 'static':
            tot = orc.static_et(col, *[d64[k] for k in NAMES])
        else:
            day, night = M(p).evapotranspiration(*[d64[k] for k in NAMES])
            tot = day + night
    got = MOD16._et_potential(col, *drv, semantics=semantics)
    assert _kernel(cuda_lib) == (_abi.KERNEL_PIPE, 0)
    scale = 1.0 if semantics == 'static' else parity.MASS_SCALE
    parity.assert_parity(got['et'], tot, np.float32, scale, what='total')
    ok = ~np.isnan(d['lai'])       # (LAI does not enter the potential terms: finite there too)
    for name, want in (('potential_transpiration', pt), ('evap_sat_day', sat_d),
                       ('evap_unsat_day', unsat_d), ('evap_sat_night', sat_n),
                       ('evap_unsat_night', unsat_n)):
        parity.assert_parity(got[name][ok], want[ok], np.float32, what=name)
        parity.assert_parity(got[name][~ok], want[~ok], np.float32, what=name + ' (queued)')
    assert np.isnan(got['et'][~ok]).all() == (semantics == 'static')


def test_day_length_composite_from_the_same_pass(cuda_lib):
    'MOD16_OUT_COMPOSITE against verify2.py:112-115 applied to the oracle\'s day / night rates.'
    T, N = 23, 2400
    d = synthetic.host_drivers(T, N, 191, dtype=np.float32)
    rng = np.random.default_rng(192)
    d['fpar'][rng.random((T, N)) < 0.01] = np.nan
    drv = [d[k] for k in NAMES]
    col = list(_table()[:, 4])
    p = dict(zip(orc.PARAM_ORDER, col))
    updown = rng.uniform(8.0, 16.0, T)
    with np.errstate(all='ignore'):
        day, night = orc.InstanceModel(p).evapotranspiration(*_d64(drv))
        want = orc.daylength_composite(day, night, updown[:, None], days=8)
    got = MOD16(p).evapotranspiration_composite(updown, *drv)
    assert _kernel(cuda_lib) == (_abi.KERNEL_PIPE, 0)
    # kg m-2 (8 days)-1: rates of ~1e-5 kg m-2 s-1 times ~3e5 s
    parity.assert_parity(got, want, np.float32, parity.MASS_SCALE * 8 * 86400,
                         what='8-day composite')
    assert 0.005 < np.isnan(got).mean() < 0.02
    with pytest.raises(_abi.LayoutError):       # (ragged N: no strided kernel for this mode)
        MOD16(p).evapotranspiration_composite(updown, *[v[..., :1001] if isinstance(
            v, np.ndarray) else v for v in drv])


@pytest.mark.parametrize('out_mode', ['sum', 'day_night', 'components'])
def test_without_a_workspace_the_block_queues_and_its_own_sweep_take_over(cuda_lib, out_mode):
    '''
    The C ABI allows `workspace = NULL`: deferred elements then go to a 512-entry queue per
    block, evaluated by the block after its loop; when that overflows, markers in the output
    are found again by the block's own sweep.  10 % finite garbage (values the double code
    must evaluate as the reference does) overflows every block's queue.
    '''
    dev = torch.device('cuda', 0)
    T, N = 64, 4800
    d = synthetic.host_drivers(T, N, 201, dtype=np.float32)
    rng = np.random.default_rng(202)
    bad = rng.random((T, N)) < 0.10
    d['lai'][bad] = -9999.0            # finite fill: garbage in, the same garbage out
    d['fpar'][rng.random((T, N)) < 0.02] = 2.55
    drv = [d[k] for k in NAMES]
    col = list(_table()[:, 7])
    with np.errstate(all='ignore'):
        if out_mode == 'components':
            want = [c for half in orc.static_evapotranspiration(col, *_d64(drv), components=True)
                    for c in half]
        else:
            day, night = orc.static_evapotranspiration(col, *_d64(drv))
            want = [day + night] if out_mode == 'sum' else [day, night]
    empty = torch.empty(0, dtype=torch.int64, device=dev)
    got = run_forward(drv, params=col, out_mode=out_mode, dtype=np.float32, workspace=empty)
    assert _kernel(cuda_lib)[0] == _abi.KERNEL_PIPE
    for g, w in zip(got, want):
        parity.assert_parity(g, w, np.float32, what='no workspace, ' + out_mode)
    # and a tiny workspace (queue of 15 entries): the marker sweep kernel
    tiny = torch.empty(16, dtype=torch.int64, device=dev)
    got = run_forward(drv, params=col, out_mode=out_mode, dtype=np.float32, workspace=tiny)
    for g, w in zip(got, want):
        parity.assert_parity(g, w, np.float32, what='tiny workspace, ' + out_mode)


@pytest.mark.parametrize('semantics', ['static', 'instance'])
def test_scaled_int16_output(cuda_lib, semantics):
    'MOD16_OUT_SUM_I16: the total as the product stores it -- int16 counts, -32768 for NaN.'
    T, N = 21, 4800
    d = synthetic.host_drivers(T, N, 211, dtype=np.float32)
    rng = np.random.default_rng(212)
    d['lai'][rng.random((T, N)) < 0.01] = np.nan
    d['fpar'][rng.random((T, N)) < 0.005] = -1.0          # finite garbage: the double code decides
    drv = [d[k] for k in NAMES]
    col = list(_table()[:, 7])
    with np.errstate(all='ignore'):
        if semantics == 'static':
            want = orc.static_et(col, *_d64(drv))
            scale = 0.1                                    # W m-2 per count
        else:
            day, night = orc.InstanceModel(dict(zip(orc.PARAM_ORDER, col))).evapotranspiration(
                *_d64(drv))
            want = day + night
            scale = 1e-8                                   # kg m-2 s-1 per count
    got = run_forward(drv, params=col, semantics=semantics, out_mode='sum_i16', out_scale=scale)[0]
    assert _kernel(cuda_lib)[0] == _abi.KERNEL_PIPE
    raw = got.raw
    assert raw.dtype == np.int16 and raw.shape == (T, N)
    assert np.array_equal(raw == -32768, np.isnan(want))
    ok = ~np.isnan(want)
    counts = np.clip(want[ok] / scale, -32767, 32767)
    # half a count of rounding + the float32 tolerance of the total itself
    tol = 0.5 + np.maximum(1e-5 * np.abs(want[ok]), 1e-3 * (1.0 if semantics == 'static'
                                                            else parity.MASS_SCALE)) / scale
    assert np.all(np.abs(raw[ok] - counts) <= tol + 1e-6)
    assert np.allclose(got.physical()[ok], want[ok], atol=0.51 * scale + 1e-3 * (
        1.0 if semantics == 'static' else parity.MASS_SCALE), rtol=1e-5)
    with pytest.raises(ValueError):
        run_forward(drv, params=col, out_mode='sum_i16')


@pytest.mark.parametrize('dtype', [np.float32, np.float64])
def test_a_stream_that_is_not_the_current_one(cuda_lib, dtype):
    '''
    `stream=` with numpy inputs: the staging copies, the launch and the read-back of one call
    must all be ordered on THAT stream (they ran on the current stream
    while the kernel ran on the given one, so the kernel could start before its inputs had
    landed and the read-back could overtake it).  The current stream is kept busy while the
    call is made on a side stream; results must equal the default-stream call bit for bit --
    as numpy arrays, as device tensors, and with the stream given as a raw handle.
    '''
    dev = torch.device('cuda', 0)
    T, N = 12, 300000                      # ~170 MB of pageable host arrays to stage
    d = synthetic.host_drivers(T, N, 71, dtype=dtype, nan_frac=0.001, fill_frac=0.001)
    drv = [d[k] for k in NAMES]
    table, pft = _table(), synthetic.host_pft_map(N, 72)
    kw = dict(pft=pft, table=table, dtype=dtype, gate='global')
    want = run_forward(drv, **kw)[0]
    view = np.int32 if np.dtype(dtype) == np.float32 else np.int64

    def busy():
        'Queues ~100 ms of work on the current stream.'
        a = torch.randn(4096, 4096, device=dev)
        for _ in range(40):
            a = a @ a
            a = a / a.abs().max()
        return a

    side = torch.cuda.Stream(device=dev)
    for handle in (side, side.cuda_stream):
        keep = busy()
        got = run_forward(drv, stream=handle, **kw)[0]          # numpy in, numpy out
        assert np.array_equal(got.view(view), want.view(view))
        keep = busy()
        t = run_forward(drv, stream=handle, as_torch=True, **kw)[0]
        side.synchronize()                                      # the caller's contract
        assert np.array_equal(t.cpu().numpy().view(view), want.view(view))
        del keep
    torch.cuda.synchronize()
    # and the reference-facing call (keyword passed through MOD16._et_by_pft)
    keep = busy()
    got = MOD16._et_by_pft(table, pft, *drv, dtype=dtype, stream=side)
    assert np.array_equal(got.view(view), want.view(view))
    del keep


@pytest.mark.parametrize('semantics', ['static', 'instance'])
def test_float32_kernels_at_the_discontinuities(cuda_lib, semantics):
    '''
    tests/switch_cases.py in float32: elements 1e-8 ... 1e-3 on both sides of every switch of
    the model, rounded to float32, laid out as a (T, N) block.

    * Per-element gate: with it the zeros of the Tmin and VPD ramps are switches too (g_surf
      = 0+ turns the cuticular transpiration on, mod16/__init__.py:343-348).  The static path
      then runs on instantiations of the pipelined kernel that carry a guard for them (and on
      the strided kernel, which has it too, where the layout is not streamable); the instance
      path has no gate.
    * Batch-global gate (the reference's `_et`): the pipelined kernel, every component.
    All against the float64 oracle on the same float32 inputs, 1e-5 relative / 1e-3 W m-2.
    '''
    import switch_cases
    drv, groups, grp, off = switch_cases.build(semantics, np.float32)
    T = 40
    N = drv[0].size // T                  # 252: a multiple of 4, streamable
    blk = [np.ascontiguousarray(v.reshape(T, N)) for v in drv]
    day, night, scale = switch_cases.reference(semantics, blk)
    kw = dict(params=switch_cases.BPLUT7, semantics=semantics, dtype=np.float32)
    # all six components (with time-varying annual temperature and pressure that output is
    # the strided kernel's in any case)
    got = run_forward(blk, gate='element', out_mode='components', **kw)
    assert _kernel(cuda_lib)[0] == _abi.KERNEL_STRIDED
    for c in range(3):
        parity.assert_parity(got[c], day[c], np.float32, scale, what='element gate, day %d' % c)
        parity.assert_parity(got[3 + c], night[c], np.float32, scale,
                             what='element gate, night %d' % c)
    # day / night totals: a layout the pipeline streams (static path: the guard-enabled
    # instantiation) -- the strided kernel above and this one agree with the same oracle
    td, tn = day[0] + day[1] + day[2], night[0] + night[1] + night[2]
    got = run_forward(blk, gate='element', out_mode='day_night', **kw)
    assert _kernel(cuda_lib)[0] == _abi.KERNEL_PIPE
    parity.assert_parity(got[0], td, np.float32, scale, what='element gate, day total')
    parity.assert_parity(got[1], tn, np.float32, scale, what='element gate, night total')
    if semantics == 'static':
        # the reference's own gate: open for the whole block (some element transpires)
        with np.errstate(all='ignore'):
            gday, gnight = orc.static_evapotranspiration(
                switch_cases.BPLUT7, *[np.asarray(v, np.float64) for v in blk], gate='global',
                components=True)
        got = run_forward(blk, gate='global', out_mode='day_night', **kw)
        assert _kernel(cuda_lib)[0] == _abi.KERNEL_PIPE
        parity.assert_parity(got[0], gday[0] + gday[1] + gday[2], np.float32, scale,
                             what='global gate, day total')
        parity.assert_parity(got[1], gnight[0] + gnight[1] + gnight[2], np.float32, scale,
                             what='global gate, night total')
