#!/usr/bin/env python
'''K1 on a tile where a fraction of the PIXELS is masked (NaN fPAR / LAI at every step, as
ocean / fill pixels of a MODIS tile are after masking): python scripts/bench_fill.py [frac]'''
import json, os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mod16_b200 import MOD16, synthetic, utils
from mod16_b200.forward import run_forward
frac = float(sys.argv[1]) if len(sys.argv) > 1 else 0.3
mode = sys.argv[2] if len(sys.argv) > 2 else 'masked'   # masked | negvpd | negvpd2 | filltemp | fill
dtype = np.float64 if (len(sys.argv) > 3 and sys.argv[3] == 'f64') else np.float32
T, N = (16 if dtype == np.float64 else 64), 2400 * 2400
dev = torch.device('cuda', 0)
d = synthetic.device_drivers(T, N, 1, dev, dtype=torch.float64 if dtype == np.float64 else torch.float32)
if mode == 'masked':       # whole pixels masked (NaN fPAR / LAI at every step)
    mask = torch.rand(N, device=dev) < frac
    for k in ('fpar', 'lai'):
        d[k][:, mask] = float('nan')
elif mode == 'negvpd':     # scattered supersaturated elements (negative VPD): double code
    d['vpd_day'][torch.rand((T, N), device=dev) < frac] = -50.0
elif mode == 'negvpd2':    # both VPDs negative at random elements
    for k in ('vpd_day', 'vpd_night'):
        d[k][torch.rand((T, N), device=dev) < frac] *= -0.1
elif mode == 'filltemp':   # unmasked fill temperatures on whole pixels (-9999): NaN from registers
    mask = torch.rand(N, device=dev) < frac
    for k in ('temp_day', 'temp_night', 'tmin'):
        d[k][:, mask] = -9999.0
else:                      # unmasked fill values on whole pixels: double code, garbage out
    mask = torch.rand(N, device=dev) < frac
    d['fpar'][:, mask] = 2.55
    d['lai'][:, mask] = -9999.0
table = utils.bplut_table(utils.restore_bplut(utils.BPLUT_C51), MOD16.required_parameters, default_beta=250.0)
pft = torch.from_numpy(synthetic.host_pft_map(N, 2)).to(dev)
drv = [d[k] for k in synthetic.DRIVER_NAMES]
out = torch.empty((T, N), device=dev, dtype=torch.float64 if dtype == np.float64 else torch.float32)
flag = torch.zeros(1, device=dev, dtype=torch.int32)
f = lambda: run_forward(drv, pft=pft, table=table, semantics='static', out_mode='sum', gate='global',
                        dtype=dtype, device=dev, gate_flag=flag, out=[out])
for _ in range(2): f()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(3): f()
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 3
nan_out = float(torch.isnan(out).float().mean())
print(json.dumps({'mode': mode, 'fraction': frac, 'ms': ms, 'G units/s': T * N / ms / 1e6, 'nan_fraction_out': nan_out}))
