#!/usr/bin/env python
'''End-to-end (host to host, wall clock) time of the plain drop-in calls with pageable numpy
inputs -- no pinned buffers, no StreamedForward:
* MOD16._et_by_pft(..., dtype=float32);
* the reference's stock call MOD16._et(eleven (1, N) float64 parameter arrays, *float32 drivers),
  which numpy promotes to float64 (float64 entry, float32 drivers read as they are);
* the same call on float64 copies of the drivers (what the upload looked like before f32_mask).
python scripts/bench_plain_api.py [T]'''
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


def timed(name, fn, nbytes, reps=3):
    for _ in range(2):
        out = fn()
    t0 = time.perf_counter()
    for _ in range(reps):
        out = fn()
    dt = (time.perf_counter() - t0) / reps
    print(json.dumps({'call': name, 'T': T, 'N': N, 's': round(dt, 4),
                      'G units/s': round(T * N / dt / 1e9, 3),
                      'GB/s in+out': round(nbytes * T * N / dt / 1e9, 2),
                      'dtype': str(out.dtype), 'finite': bool(np.isfinite(out).all())}))
    return out


vec = [table[i][pft].reshape(1, N) for i in range(11)]          # float64, the notebook's form
a = timed('MOD16._et(float64 (1, N) parameters, float32 numpy drivers)  [stock call]',
          lambda: MOD16._et(vec, *host), 11 * 4 + 8)       # eleven (T, N) float32 arrays in, float64 out
host64 = [v.astype(np.float64) if isinstance(v, np.ndarray) else v for v in host]
b = timed('MOD16._et(float64 (1, N) parameters, float64 numpy drivers)',
          lambda: MOD16._et(vec, *host64), 11 * 8 + 8)
print(json.dumps({'same bits': bool(np.array_equal(a.view(np.int64), b.view(np.int64)))}))
