#!/usr/bin/env python
'''K1 throughput of every semantics / output mode on a resident block: python scripts/bench_modes.py [T]'''
import json, os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mod16_b200 import MOD16, synthetic, utils
from mod16_b200.forward import run_forward
T = int(sys.argv[1]) if len(sys.argv) > 1 else 64
N = 2400 * 2400
dev = torch.device('cuda', 0)
d = synthetic.device_drivers(T, N, 1, dev)
table = utils.bplut_table(utils.restore_bplut(utils.BPLUT_C51), MOD16.required_parameters, default_beta=250.0)
pft = torch.from_numpy(synthetic.host_pft_map(N, 2)).to(dev)
drv = [d[k] for k in synthetic.DRIVER_NAMES]
flag = torch.zeros(1, device=dev, dtype=torch.int32)
for sem in ('static', 'instance'):
    for mode, n_out in (('sum', 1), ('day_night', 2), ('components', 6)):
        outs = [torch.empty((T, N), device=dev) for _ in range(n_out)]
        f = lambda: run_forward(drv, pft=pft, table=table, semantics=sem, out_mode=mode, gate='global',
                                dtype=np.float32, device=dev, gate_flag=flag, out=outs)
        for _ in range(2): f()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(5): f()
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 5
        b = 44 + 4 * n_out
        print(json.dumps({'semantics': sem, 'out': mode, 'ms': ms, 'G units/s': T * N / ms / 1e6,
                          'bytes_per_unit': b, 'GB/s': b * T * N / ms / 1e6}))
        del outs
