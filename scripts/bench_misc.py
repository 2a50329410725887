#!/usr/bin/env python
'''Throughput of the K1 variants next to the headline one on a 64-step slab of the tile:
float64 fast path, annual-sum epilogue, storage formats (MODIS / PACKED), K4 (one op).
python scripts/bench_misc.py'''
import json, os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mod16_b200 import MOD16, synthetic, utils, rawio
from mod16_b200.forward import run_forward
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
T, N = 64, 2400 * 2400
dev = torch.device('cuda', 0)
d = synthetic.device_drivers(T, N, 1, dev)
table = utils.bplut_table(utils.restore_bplut(utils.BPLUT_C51), MOD16.required_parameters, default_beta=250.0)
pft = torch.from_numpy(synthetic.host_pft_map(N, 2)).to(dev)
names = list(synthetic.DRIVER_NAMES)


def timed(name, fn, units, nbytes, reps=5):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    print(json.dumps({'case': name, 'ms': ms, 'G units/s': units / ms / 1e6,
                      'GB/s (algorithmic)': nbytes * units / ms / 1e6}))


drv = [d[k] for k in names]
out = torch.empty((T, N), device=dev)
flag = torch.zeros(1, device=dev, dtype=torch.int32)
timed('float32 sum (headline kernel)', lambda: run_forward(drv, pft=pft, table=table, gate='global', dtype=np.float32, device=dev, gate_flag=flag, out=[out]), T * N, 48)
years = np.zeros(T, np.int32)
timed('annual sums epilogue (no per-step output)', lambda: MOD16._annual_et((table, pft), years, *drv, as_torch=True), T * N, 44)
for layout in ('modis', 'packed'):
    fmt = dict(bench.MODIS_NATIVE)
    if layout == 'packed':
        fmt.update(dict((k, (np.int16,) + v) for k, v in bench.PACKED_MET.items()))
    raw = [rawio.quantize(d[k], fmt[k][0], *fmt[k][1:]) if k in fmt else d[k] for k in names]
    nb = sum(np.dtype(fmt[k][0]).itemsize if k in fmt else 4 for k in names if isinstance(d[k], torch.Tensor) and d[k].ndim == 2) + 4
    timed('storage formats: %s' % layout, lambda: run_forward(raw, pft=pft, table=table, gate='global', device=dev, gate_flag=flag, out=[out]), T * N, nb)
    del raw
Ts = 16
d64 = [v[:Ts].double() if isinstance(v, torch.Tensor) and v.ndim == 2 else (v.double() if isinstance(v, torch.Tensor) else v) for v in drv]
out64 = torch.empty((Ts, N), device=dev, dtype=torch.float64)
timed('float64 fast path', lambda: run_forward(d64, pft=pft, table=table, gate='global', dtype=np.float64, device=dev, gate_flag=flag, out=[out64]), Ts * N, 96, reps=3)
from mod16_b200 import _abi
timed('float64 op-for-op (MOD16_FORCE_MIRROR)', lambda: run_forward(d64, pft=pft, table=table, gate='global', dtype=np.float64, device=dev, gate_flag=flag, out=[out64], force=_abi.FORCE_MIRROR), Ts * N, 96, reps=2)
from mod16_b200 import components
timed('K4 svp (one helper, float64)', lambda: components.svp(d64[5]), Ts * N, 16, reps=3)
