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

Under torchrun (one rank per GPU, NCCL) `--shard draw` gives every rank all sites and
D / world draws (all-gather of the results, no data-path collective); `--shard site` gives
every rank N / world sites and all draws (ONE all-reduce of the (D, 5) un-gated statistics,
then each draw's gate over all shards).  Strong scaling: the job is the same at every world.
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
    ap.add_argument('--shard', default='draw', choices=['draw', 'site'])
    args = ap.parse_args()

    import torch
    from mod16_b200 import synthetic, utils, MOD16
    from mod16_b200.objective import ObjectiveBackend
    from mod16_b200.forward import run_forward
    import torch.distributed as dist
    from mod16_b200 import distributed as mdist
    world = int(os.environ.get('WORLD_SIZE', '1'))
    rank = int(os.environ.get('RANK', '0'))
    dev = torch.device('cuda', int(os.environ.get('LOCAL_RANK', '0')))
    torch.cuda.set_device(dev)
    if world > 1:
        dist.init_process_group('nccl', device_id=dev)
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
    # (every rank generated the same data: same seeds)
    if world > 1 and args.shard == 'site':
        lo, hi = mdist.pixel_range(N, rank, world, align=1)
        local = mdist.shard_pixels(drv, lo, hi)
        backend = ObjectiveBackend(local, obs[:, lo:hi], w[:, lo:hi], dtype=np.dtype(args.dtype),
                                   device=dev)
        run = lambda: mdist.objective_site_sharded(backend, draws)
    elif world > 1:
        backend = ObjectiveBackend(drv, obs, w, dtype=np.dtype(args.dtype), device=dev)
        run = lambda: mdist.objective_draw_sharded(backend, draws)
    else:
        backend = ObjectiveBackend(drv, obs, w, dtype=np.dtype(args.dtype), device=dev)
        run = lambda: backend(draws, as_torch=True)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
    for _ in range(args.warmup):
        sse, cnt = run()
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        sse, cnt = run()
    e1.record()
    barrier()
    tm = torch.tensor([e0.elapsed_time(e1) / args.steps], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(tm, op=dist.ReduceOp.MAX)
    ms = float(tm.item())
    valid = float(cnt.max().item())
    units = valid * D
    rm = torch.sqrt(sse / cnt)

    # the reference loop on the host: one numpy `_et` + nanmean per draw
    cpu = None
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
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
        'unit': 'draw-elements/s', 'n_gpus': world, 'shard': args.shard if world > 1 else None,
        'scaling': 'strong', 'ms_per_step': ms, 'dtype': args.dtype,
        'config': {'workload': args.config, 'draws': D, 'T': T, 'N': N,
                   'valid_elements': valid, 'all_elements': T * N},
        'rmsd_range': [float(rm.min().item()), float(rm.max().item())],
        'cpu_baseline': cpu}), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == '__main__':
    main()
