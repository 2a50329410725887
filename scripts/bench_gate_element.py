#!/usr/bin/env python
'''What the per-element gate costs on the float32 static path: a 64-step slab of the tile with
the batch-global gate on the pipelined kernel, the same forced onto the strided kernel, and the
per-element gate (routed to the strided kernel, which guards the zeros of the Tmin / VPD ramps).
python scripts/bench_gate_element.py'''
import json, os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mod16_b200 import MOD16, synthetic, utils, _abi
from mod16_b200.forward import run_forward
T, N = 64, 2400 * 2400
dev = torch.device('cuda', 0)
d = synthetic.device_drivers(T, N, 1, dev)
table = utils.bplut_table(utils.restore_bplut(utils.BPLUT_C51), MOD16.required_parameters,
                          default_beta=250.0)
pft = torch.from_numpy(synthetic.host_pft_map(N, 2)).to(dev)
drv = [d[k] for k in synthetic.DRIVER_NAMES]
out = torch.empty((T, N), device=dev)
flag = torch.zeros(1, device=dev, dtype=torch.int32)
lib = _abi.load()
names = {_abi.KERNEL_PIPE: 'forward_pipe_kernel', _abi.KERNEL_STRIDED: 'forward_strided_fast_kernel'}


def timed(name, reps=5, **kw):
    fn = lambda: run_forward(drv, pft=pft, table=table, dtype=np.float32, device=dev,
                             gate_flag=flag, out=[out], **kw)
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
                      'kernel': names.get(lib.mod16_last_forward_kernel() & 0xff)}))


timed('batch-global gate', gate='global')
timed('batch-global gate, strided kernel forced', gate='global', force=_abi.FORCE_GENERIC)
timed('per-element gate', gate='element')
timed('gate open', gate='open')
