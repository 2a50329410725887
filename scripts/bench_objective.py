#!/usr/bin/env python
'''
Throughput of K2 (objective batch) on the BASELINE.json configs 4 and 5:

    python scripts/bench_objective.py --config calibration [--draws D] [--steps K]
    python scripts/bench_objective.py --config sobol

calibration: D = 10^5 draws x 300 sites x 3652 days (flattened, 30 % of the observations
missing and removed by the compaction), tmin_close / tmin_open / vpd_open fixed;
sobol: D = 2^16 rows x 1000 pixels x 365 days, all eleven parameters free.
One step = one `mod16_objective_batch_f32` call (all launches of the call); drivers,
observations and draws are resident in HBM.  Prints one JSON line; the unit is
draw-elements/s over the VALID elements (those that enter the sums).
'''
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--config', default='calibration', choices=['calibration', 'sobol'])
    ap.add_argument('--draws', type=int, default=None)
    ap.add_argument('--steps', type=int, default=3)
    ap.add_argument('--warmup', type=int, default=1)
    ap.add_argument('--dtype', default='float32')
    ap.add_argument('--cpu-draws', type=int, default=2)
    args = ap.parse_args()

    import torch
    from mod16_b200 import synthetic, utils, MOD16
    from mod16_b200.objective import ObjectiveBackend
    from mod16_b200.forward import run_forward
    dev = torch.device('cuda', 0)
    torch.cuda.set_device(dev)
    if args.config == 'calibration':
        T, N, D, seed, fixed = 3652, 300, 100000, 16004, ('tmin_close', 'tmin_open', 'vpd_open')
    else:
        T, N, D, seed, fixed = 365, 1000, 65536, 16005, ()
    D = args.draws or D
    t_dtype = torch.float32 if args.dtype == 'float32' else torch.float64
    table = utils.bplut_table(utils.restore_bplut(utils.BPLUT_C51), MOD16.required_parameters,
                              default_beta=250.0)
    col = np.array(table[:, 7])
    d = synthetic.device_drivers(T, N, seed, dev, dtype=t_dtype)
    drv = [d[k] for k in synthetic.DRIVER_NAMES]
    truth = run_forward(drv, params=list(col), dtype=np.dtype(args.dtype), device=dev,
                        as_torch=True)[0]
    g = torch.Generator(device=dev)
    g.manual_seed(seed)
    obs = truth + 15.0 * torch.randn(truth.shape, generator=g, device=dev, dtype=t_dtype)
    obs[obs < 0] = float('nan')
    obs[torch.rand(truth.shape, generator=g, device=dev) < 0.3] = float('nan')
    w = 0.2 + 0.8 * torch.rand(truth.shape, generator=g, device=dev, dtype=t_dtype)
    draws = torch.from_numpy(synthetic.host_param_draws(D, seed, bplut_column=col,
                                                        fixed=fixed)).to(dev)
    backend = ObjectiveBackend(drv, obs, w, dtype=np.dtype(args.dtype), device=dev)
    for _ in range(args.warmup):
        sse, cnt = backend(draws, as_torch=True)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        sse, cnt = backend(draws, as_torch=True)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / args.steps
    valid = float(cnt.max().item())
    units = valid * D
    rm = torch.sqrt(sse / cnt)

    # the reference loop on the host: one numpy `_et` + nanmean per draw
    cpu = None
    if args.cpu_draws > 0:
        from oracle import mod16_oracle as orc
        h = [v.cpu().numpy().astype(np.float64) if isinstance(v, torch.Tensor) else v for v in drv]
        ho, hw = obs.cpu().numpy().astype(np.float64), w.cpu().numpy().astype(np.float64)
        hd = draws[:args.cpu_draws].cpu().numpy()
        t0 = time.perf_counter()
        ref = []
        with np.errstate(all='ignore'):
            for x in hd:
                s, c = orc.objective_stats(orc.static_et(list(x), *h), ho, hw)
                ref.append(np.sqrt(s / c))
        dt = time.perf_counter() - t0
        got = rm[:args.cpu_draws].cpu().numpy()
        cpu = {'value': args.cpu_draws * T * N / dt, 'unit': 'draw-elements/s (all elements)',
               'cores': 1, 'kind': 'port', 'sample': '%d draws' % args.cpu_draws,
               'rmsd_max_abs_diff': float(np.max(np.abs(got - np.array(ref))))}
    print(json.dumps({
        'metric': 'draw-elements/s (objective batch)', 'value': units / (ms * 1e-3),
        'unit': 'draw-elements/s', 'ms_per_step': ms, 'dtype': args.dtype,
        'config': {'workload': args.config, 'draws': D, 'T': T, 'N': N,
                   'valid_elements': valid, 'all_elements': T * N},
        'rmsd_range': [float(rm.min().item()), float(rm.max().item())],
        'cpu_baseline': cpu}), flush=True)


if __name__ == '__main__':
    main()
