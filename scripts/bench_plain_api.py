#!/usr/bin/env python
'''End-to-end time of the plain drop-in call with pageable numpy inputs (no pinned buffers, no
StreamedForward): python scripts/bench_plain_api.py [T]'''
import json, os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mod16_b200 import MOD16, synthetic, utils
T = int(sys.argv[1]) if len(sys.argv) > 1 else 16
N = 2400 * 2400
d = synthetic.device_drivers(T, N, 3, torch.device('cuda', 0))
host = [v.cpu().numpy() if isinstance(v, torch.Tensor) else v for v in (d[k] for k in synthetic.DRIVER_NAMES)]
del d
torch.cuda.empty_cache()
table = utils.bplut_table(utils.restore_bplut(utils.BPLUT_C51), MOD16.required_parameters, default_beta=250.0)
pft = synthetic.host_pft_map(N, 2)
for _ in range(2):
    out = MOD16._et_by_pft(table, pft, *host, dtype=np.float32)
t0 = time.perf_counter()
reps = 3
for _ in range(reps):
    out = MOD16._et_by_pft(table, pft, *host, dtype=np.float32)
dt = (time.perf_counter() - t0) / reps
print(json.dumps({'call': 'MOD16._et_by_pft(numpy)', 'T': T, 'N': N, 's': dt, 'G units/s': T * N / dt / 1e9,
                  'GB/s in+out': 48.0 * T * N / dt / 1e9, 'finite': bool(np.isfinite(out).all())}))
