#!/usr/bin/env python
'''Throughput of the float64 entry on a slab of the tile: float64 drivers, float32 drivers
(the reference's stock call: float32 archive + float64 BPLUT), 30 % masked pixels.
python scripts/bench_f64.py [T]      (MOD16_B200_LIB selects a library variant)'''
import json, os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mod16_b200 import MOD16, synthetic, utils, _abi
from mod16_b200.forward import run_forward
T = int(sys.argv[1]) if len(sys.argv) > 1 else 32
N = 2400 * 2400
dev = torch.device('cuda', 0)
d = synthetic.device_drivers(T, N, 1, dev)
table = utils.bplut_table(utils.restore_bplut(utils.BPLUT_C51), MOD16.required_parameters,
                          default_beta=250.0)
pft = torch.from_numpy(synthetic.host_pft_map(N, 2)).to(dev)
names = list(synthetic.DRIVER_NAMES)
drv32 = [d[k] for k in names]
drv64 = [v.double() if isinstance(v, torch.Tensor) else v for v in drv32]
out = torch.empty((T, N), device=dev, dtype=torch.float64)
flag = torch.zeros(1, device=dev, dtype=torch.int32)
lib = _abi.load()


def timed(name, drv, nbytes, reps=5, **kw):
    fn = lambda: run_forward(drv, pft=pft, table=table, gate='global', dtype=np.float64,
                             device=dev, gate_flag=flag, out=[out], **kw)
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    print(json.dumps({'case': name, 'ms': round(ms, 3), 'G units/s': round(T * N / ms / 1e6, 2),
                      'GB/s (algorithmic)': round(nbytes * T * N / ms / 1e6, 1),
                      'kernel': lib.mod16_last_forward_kernel() & 0xff}))
    return out.clone()


a = timed('float64 drivers', drv64, 96)
b = timed('float32 drivers, float64 parameters (stock promotion)', drv32, 52)
print(json.dumps({'float32-driver results equal the float64-copy results':
                  bool(torch.equal(torch.nan_to_num(a, nan=-1.0), torch.nan_to_num(b, nan=-1.0)))}))
mixed = list(drv32); mixed[5] = drv64[5]; mixed[12] = drv64[12]
timed('mixed widths', mixed, 60)
mask = torch.rand(N, device=dev) < 0.3
m32 = list(drv32)
for k in (12, 13):
    m32[k] = torch.where(mask[None, :], torch.full_like(m32[k], float('nan')), m32[k])
timed('float32 drivers, 30 % masked pixels', m32, 52)
