'''
GPU parity tests of K3 (per-pixel time aggregates) against the oracle's restatement of the
calibration's annual sums (reference mod16/calibration.py:783-790).
'''
import numpy as np
import pytest
import torch

from mod16_b200 import aggregate, synthetic
from oracle import mod16_oracle as orc

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize('dtype', [np.float64, np.float32])
@pytest.mark.parametrize('T,N', [(730, 317), (366, 4096 + 3), (5, 1)])
def test_annual_mass_sums_like_constrain_by_map(cuda_lib, dtype, T, N):
    rng = np.random.default_rng(T + N)
    le = rng.normal(60, 50, (T, N)).astype(dtype)          # negative rates occur
    temp = rng.uniform(255, 305, (T, N)).astype(dtype)
    lhv = orc.lhv(temp).astype(dtype)
    years = 2000 + (np.arange(T) // 365)
    ref = orc.annual_mass_sums(le.astype(np.float64), years, lhv.astype(np.float64))
    got = aggregate.annual_mass_sums(le, years, lhv)
    assert got.shape == ref.shape and got.dtype == np.float64
    tol = 1e-10 if dtype == np.float64 else 2e-6   # f32: the division is done in double
    np.testing.assert_allclose(got, ref, rtol=tol, atol=tol)


def test_groups_nan_and_total(cuda_lib):
    rng = np.random.default_rng(3)
    T, N = 50, 777
    x = rng.normal(0, 1, (T, N))
    x[rng.random((T, N)) < 0.01] = np.nan
    group = rng.integers(-1, 4, T)                 # -1: skipped steps; unsorted labels
    out, tot = aggregate.time_aggregate(x, group=group, n_groups=4, total=True)
    for g in range(4):
        ref = x[group == g].sum(axis=0)
        assert np.array_equal(np.isnan(out[g]), np.isnan(ref))
        np.testing.assert_allclose(out[g], ref, rtol=1e-12, atol=1e-12)
        assert np.isnan(tot[g]) == np.isnan(ref.sum())
    out2, tot2 = aggregate.time_aggregate(x, group=group, n_groups=4, skip_nan=True, total=True)
    for g in range(4):
        np.testing.assert_allclose(out2[g], np.nansum(x[group == g], axis=0), rtol=1e-12,
                                   atol=1e-12)
        np.testing.assert_allclose(tot2[g], np.nansum(x[group == g]), rtol=1e-10)
    # one group, CUDA tensors in and out
    xt = torch.from_numpy(np.nan_to_num(x)).cuda()
    o = aggregate.time_aggregate(xt)
    assert isinstance(o, torch.Tensor) and o.shape == (1, N)
    np.testing.assert_allclose(o.cpu().numpy()[0], np.nan_to_num(x).sum(axis=0), rtol=1e-12,
                               atol=1e-12)


def test_aggregate_of_forward_output_is_the_annual_et(cuda_lib):
    'K1 -> K3 on the device: a year of `_et` summed per pixel without leaving the GPU.'
    from mod16_b200 import MOD16
    T, N = 365, 512
    d = synthetic.host_drivers(T, N, 5, dtype=np.float32)
    names = list(synthetic.DRIVER_NAMES)
    p7 = [-8., 8.8, 650., 4400., .02, .02, 1e-5, .0055, 60., 95., 250.]
    drv = [torch.from_numpy(d[k]).cuda() if isinstance(d[k], np.ndarray) else d[k] for k in names]
    et = MOD16._et(p7, *drv, dtype=np.float32)
    lhv = torch.from_numpy(orc.lhv(d['temp_day'].astype(np.float64)).astype(np.float32)).cuda()
    got = aggregate.annual_mass_sums(et, np.zeros(T, np.int32), lhv).cpu().numpy()
    with np.errstate(all='ignore'):
        ref_et = orc.static_et(p7, *[d[k].astype(np.float64) if isinstance(d[k], np.ndarray)
                                     else d[k] for k in names])
    ref = orc.annual_mass_sums(ref_et, np.zeros(T, np.int32),
                               orc.lhv(d['temp_day'].astype(np.float64)))
    np.testing.assert_allclose(got, ref, rtol=2e-5)


def test_aggregate_outputs_stay_inside_their_buffers(cuda_lib):
    import ctypes
    from mod16_b200 import _abi
    lib = _abi.load()
    for T, N in ((9, 1000), (9, 1001), (3, 4)):
        x = torch.rand((T, N), device='cuda')
        pad = 256
        big = torch.full((2 * N + 2 * pad,), -5.0, device='cuda', dtype=torch.float64)
        grp = torch.tensor([0, 1] * T, dtype=torch.int32, device='cuda')[:T]
        ag = _abi.Aggregate()
        ag.x, ag.scale, ag.group, ag.n_groups = x.data_ptr(), 1.0, grp.data_ptr(), 2
        ag.out = big.data_ptr() + 8 * pad
        _abi.check(lib.mod16_time_aggregate_f32(ctypes.byref(ag), T, N, 0, None))
        torch.cuda.synchronize()
        assert bool((big[:pad] == -5.0).all()) and bool((big[pad + 2 * N:] == -5.0).all())
        ref = torch.stack([x[0::2].double().sum(0), x[1::2].double().sum(0)]).reshape(-1)
        assert torch.allclose(big[pad:pad + 2 * N], ref, rtol=1e-12)


def test_constrain_by_map_golden(cuda_lib, golden_dir):
    'Both halves of the precipitation constraint against the reference\'s own statements.'
    import os
    g = np.load(os.path.join(golden_dir, 'constraint.npz'))
    for tag in 'abc':
        le, years, lhv, precip = (g['%s_%s' % (tag, k)] for k in ('le', 'years', 'lhv', 'precip'))
        want = float(g[tag + '_neg_nrmsd'])
        got = aggregate.constrain_by_map(le, years, lhv, precip)
        assert abs(got - want) <= 1e-10 * max(abs(want), 1e-12), (tag, got, want)
        annual = aggregate.annual_mass_sums(le, years, lhv)
        assert np.allclose(annual, g[tag + '_annual'], rtol=1e-12, atol=1e-9)
        # float32 fluxes (what K1 writes): the sums are still accumulated in double
        got32 = aggregate.constrain_by_map(le.astype(np.float32), years, lhv.astype(np.float32),
                                           precip)
        assert abs(got32 - want) <= 1e-5 * max(abs(want), 1e-12)


@pytest.mark.parametrize('T,N,by_pft', [(730, 4800, True), (366, 964, False), (41, 480, False)])
def test_annual_sums_fused_into_the_forward_kernel(cuda_lib, T, N, by_pft):
    '''
    out_mode='annual' (MOD16._annual_et): the annual ET of `_et` without ever writing the
    (T, N) flux, against the oracle's `constrain_by_map` ingredients on the same inputs.
    '''
    from mod16_b200 import MOD16, utils
    from mod16_b200.forward import run_forward
    d = synthetic.host_drivers(T, N, 1601, dtype=np.float32)
    rng = np.random.default_rng(1602)
    d['lai'][rng.random((T, N)) < 0.001] = np.nan          # a NaN step poisons its pixel-year
    d['temp_day'][rng.random((T, N)) < 0.0005] = -9999.0
    d['sw_rad_day'][:, :7] = 0.0                           # negative fluxes are dropped (:786)
    names = list(synthetic.DRIVER_NAMES)
    drv = [d[k] for k in names]
    d64 = [v.astype(np.float64) if isinstance(v, np.ndarray) else v for v in drv]
    years = 2001 + (np.arange(T) * 3) // max(T, 1)         # three unequal "years"
    table = utils.bplut_table(utils.restore_bplut(utils.BPLUT_C51), MOD16.required_parameters,
                              default_beta=250.0)
    if by_pft:
        pft = synthetic.host_pft_map(N, 1603)
        params = (table, pft)
        vec = [table[i][pft].reshape(1, N) for i in range(11)]
    else:
        params = vec = list(table[:, 7])
    with np.errstate(all='ignore'):
        le = orc.static_et(vec, *d64)
        lhv = orc.lhv(0.5 * (d64[5] + d64[6]))             # calibration.py:838-842
        want = orc.annual_mass_sums(le.copy(), years, lhv)
    got = MOD16._annual_et(params, years, *drv)
    assert got.shape == want.shape == (3, N) and got.dtype == np.float64
    assert np.array_equal(np.isnan(got), np.isnan(want))
    ok = ~np.isnan(want)
    assert np.isnan(want).mean() > 0.01
    assert np.all(np.abs(got - want)[ok] <= 1e-5 * np.maximum(np.abs(want[ok]), 1.0)), \
        np.abs(got - want)[ok].max()
    # integer accumulation: bit-identical again, and whatever the launch geometry
    again = MOD16._annual_et(params, years, *drv)
    assert np.array_equal(again, got, equal_nan=True)
    h = (N // 2) // 4 * 4
    half = MOD16._annual_et(params if not by_pft else (table, pft[:h]), years,
                            *[v[..., :h] if isinstance(v, np.ndarray) else v for v in drv])
    assert np.array_equal(half, got[:, :h], equal_nan=True)
    # the objective itself (nRMSD), and the same from the unfused pieces
    precip = rng.uniform(100, 900, (3, N))
    keep = ~np.isnan(want).any(axis=0)
    cols = np.flatnonzero(keep)[: (keep.sum() // 4) * 4]
    sub = [v[..., cols] if isinstance(v, np.ndarray) else v for v in drv]
    sub64 = [v.astype(np.float64) if isinstance(v, np.ndarray) else v for v in sub]
    p_sub = (table, pft[cols]) if by_pft else params
    v_sub = [table[i][pft[cols]].reshape(1, -1) for i in range(11)] if by_pft else vec
    with np.errstate(all='ignore'):
        ref = orc.constrain_by_map(orc.static_et(v_sub, *sub64), years,
                                   orc.lhv(0.5 * (sub64[5] + sub64[6])), precip[:, cols])
    val = aggregate.constrain_by_map_fused(p_sub, sub, years, precip[:, cols])
    assert abs(val - ref) <= 1e-5 * abs(ref), (val, ref)
