#!/usr/bin/env python
'''Throughput of K3 (annual sums) on the tile-year: python scripts/bench_aggregate.py'''
import json, os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mod16_b200.aggregate import time_aggregate
T, N = 365, 2400 * 2400
dev = torch.device('cuda', 0)
x = torch.rand((T, N), device=dev) * 100
lam = 2.4e6 + torch.rand((T, N), device=dev) * 1e5
grp = torch.zeros(T, dtype=torch.int32, device=dev)
for name, kw, nbytes in (('sum', {}, 4), ('mass (x * 86400 / lhv, clamp)', dict(divisor=lam, scale=86400.0, clamp_neg=True), 8)):
    for _ in range(3):
        out = time_aggregate(x, group=grp, n_groups=1, **kw)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        out = time_aggregate(x, group=grp, n_groups=1, **kw)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 10
    print(json.dumps({'kernel': 'time_aggregate_kernel<float>', 'case': name, 'ms': ms,
                      'GB/s': nbytes * T * N / ms / 1e6, 'G elements/s': T * N / ms / 1e6}))
