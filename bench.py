#!/usr/bin/env python
'''
Benchmark of the MOD16 forward hot path (BASELINE.json metric: pixel-timesteps/s of ET).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

One "step" = one `_et` pass over one resident block of synthetic drivers: at N = 1 the
MODIS 500 m tile-year (2400 x 2400 pixels x 365 daily steps, float32, 100.9 GB in HBM --
BASELINE.json configs[1]); with N > 1 every rank owns one such tile-year (independent pixel
ranges, no data-path collective; "scaling": "weak").  Inputs are far larger than L2
(126 MB), so no flush is needed between timed iterations.

Printed by rank 0 as ONE JSON line: whole-job throughput with inputs resident in HBM
(`value`), the same metric end to end from pinned host buffers through the public API
(`e2e`), the HBM roofline of the fused kernel (`roofline`), the numpy port of the reference
timed on this box's host cores (`cpu_baseline`), SM clocks under load (`clocks`).

`--impl reference` times the CPU implementation of the path (the numpy restatement of the
reference in oracle/, all host cores) on a bounded sample of the same workload.
'''
import argparse
import json
import os
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

TILE_N = 2400 * 2400
YEAR_T = 365
BYTES_PER_UNIT_F32 = 48  # 11 per-step driver reads x 4 B + 1 ET write x 4 B
METRIC = 'pixel-timesteps/s (ET)'


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=10)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='ours', choices=['ours', 'reference'])
    ap.add_argument('--pixels', type=int, default=TILE_N, help='pixels per GPU')
    ap.add_argument('--draws', type=int, default=None, help='objective batch: parameter draws')
    ap.add_argument('--cpu-draws', type=int, default=2,
                    help='objective batch: draws of the CPU baseline loop')
    ap.add_argument('--shard', default='draw', choices=['draw', 'site'],
                    help='objective batch under torchrun: shard by draw or by site')
    ap.add_argument('--workload', default='tile',
                    choices=['tile', 'global', 'calibration', 'sobol'],
                    help="tile: one MODIS tile-year per GPU (BASELINE configs[1], weak scaling; "
                         "the default and the headline); global: the 0.05 deg CMG year "
                         "(7200 x 3600, configs[2]) pixel-tiled across the ranks (strong "
                         "scaling; needs >= 4 GPUs to be resident); calibration / sobol: the "
                         "objective batch K2 on configs[3] / configs[4] (secondary lines)")
    ap.add_argument('--T', type=int, default=YEAR_T, help='time steps resident per GPU')
    ap.add_argument('--dtype', default='float32', choices=['float32', 'float64'])
    ap.add_argument('--kind', default='physical', choices=['physical', 'stress'])
    ap.add_argument('--e2e-T', type=int, default=16, help='time steps of the end-to-end leg')
    ap.add_argument('--cpu-pixels', type=int, default=20000)
    ap.add_argument('--no-cpu', action='store_true')
    ap.add_argument('--no-e2e', action='store_true')
    ap.add_argument('--gate', default='global')
    ap.add_argument('--kernel', default='pipe', choices=['pipe', 'generic', 'mirror'],
                    help='f32 kernel variant (profiling): bulk-copy pipeline (default), strided '
                         'fast math, op-for-op')
    return ap.parse_args()


def bplut_table():
    from mod16_b200 import MOD16, utils
    return utils.bplut_table(utils.restore_bplut(utils.BPLUT_C51),
                             MOD16.required_parameters, default_beta=250.0)


# ---- clocks ------------------------------------------------------------------------------
class ClockSampler(threading.Thread):
    'Samples SM clock and throttle reasons of one GPU through NVML while the bench runs.'

    def __init__(self, index, period=0.01):
        super().__init__(daemon=True)
        self.index, self.period = index, period
        self.samples, self.reasons = [], set()
        self.max_mhz = None
        self._stop_evt = threading.Event()
        self.active = threading.Event()

    def run(self):
        try:
            import pynvml
            pynvml.nvmlInit()
            h = pynvml.nvmlDeviceGetHandleByIndex(self.index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(h, pynvml.NVML_CLOCK_SM)
            names = {
                pynvml.nvmlClocksThrottleReasonHwSlowdown: 'hw_slowdown',
                pynvml.nvmlClocksThrottleReasonHwThermalSlowdown: 'hw_thermal_slowdown',
                pynvml.nvmlClocksThrottleReasonSwThermalSlowdown: 'sw_thermal_slowdown',
                pynvml.nvmlClocksThrottleReasonSwPowerCap: 'sw_power_cap',
                pynvml.nvmlClocksThrottleReasonHwPowerBrakeSlowdown: 'hw_power_brake',
            }
            while not self._stop_evt.is_set():
                if self.active.is_set():
                    self.samples.append(pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM))
                    r = pynvml.nvmlDeviceGetCurrentClocksThrottleReasons(h)
                    for bit, nm in names.items():
                        if r & bit:
                            self.reasons.add(nm)
                time.sleep(self.period)
        except Exception as exc:  # NVML missing: report nothing rather than die
            self.reasons.add('nvml_unavailable: %s' % exc)

    def stop(self):
        self._stop_evt.set()

    def summary(self):
        return {'sm_mhz': float(np.median(self.samples)) if self.samples else None,
                'sm_max_mhz': self.max_mhz, 'samples': len(self.samples),
                'reasons': sorted(self.reasons)}


# ---- CPU legs ----------------------------------------------------------------------------
_CPU_DATA = {}


def _cpu_prepare(key, seed, T, n, kind):
    'Builds (once per process) the inputs of the numpy port for one pixel chunk.'
    if key not in _CPU_DATA:
        from mod16_b200 import synthetic
        d = synthetic.host_drivers(T, n, seed, kind=kind, dtype=np.float32)
        table = bplut_table()
        pft = synthetic.host_pft_map(n, seed + 1)
        vec = [table[i][pft].reshape(1, n) for i in range(11)]
        _CPU_DATA[key] = (vec, [d[k] for k in synthetic.DRIVER_NAMES])
    return _CPU_DATA[key]


def _cpu_chunk(args):
    """
    One worker: the numpy port of the reference `_et` on its pixel chunk.  With run=False
    only the inputs are built, so that the timed region holds the model evaluation alone.
    """
    seed, T, n, kind, run = args
    from oracle import mod16_oracle as orc
    vec, drv = _cpu_prepare((os.getpid(), T, n, kind), seed + os.getpid() % 9973, T, n, kind)
    if not run:
        return 0.0, 0.0
    t0 = time.perf_counter()
    with np.errstate(all='ignore'):
        out = orc.static_et(vec, *drv)
    dt = time.perf_counter() - t0
    return dt, float(np.nansum(out))


def cpu_single(T, n, kind):
    _cpu_chunk((16002, T, n, kind, False))
    dt, _ = _cpu_chunk((16002, T, n, kind, True))
    return {'value': T * n / dt, 'unit': 'pixel-timesteps/s', 'cores': 1, 'kind': 'port',
            'sample': 'numpy port of MOD16._et (oracle/mod16_oracle.py), float32 drivers + '
                      'float64 per-pixel BPLUT parameters, %d pixels x %d steps of the '
                      'workload, 1 process, numpy %s' % (n, T, np.__version__)}


def run_reference(args):
    'The reference arm: the CPU implementation of the path on all host cores.'
    import multiprocessing as mp
    cores = os.cpu_count() or 1
    T, n = args.T, args.cpu_pixels
    try:  # the numpy path holds ~42 full-size temporaries: bound the sample by free memory
        import psutil
        budget = 0.4 * psutil.virtual_memory().available / cores
        n = int(max(256, min(n, budget / (400.0 * T))))
    except Exception:
        pass
    ctx = mp.get_context('spawn')
    with ctx.Pool(cores) as pool:
        jobs = lambda run: [(16002, T, n, args.kind, run)] * cores
        pool.map(_cpu_chunk, jobs(False), chunksize=1)      # imports, inputs, page-in
        for _ in range(args.warmup):
            pool.map(_cpu_chunk, jobs(True), chunksize=1)
        t0 = time.perf_counter()
        for _ in range(args.steps):
            pool.map(_cpu_chunk, jobs(True), chunksize=1)
        wall = time.perf_counter() - t0
    units = args.steps * cores * T * n
    value = units / wall
    sample = ('numpy port of MOD16._et (oracle/mod16_oracle.py; the Python reference itself '
              'cannot travel to the GPU box), %d processes x %d pixels x %d steps per bench '
              'step, float32 drivers + float64 per-pixel parameters, numpy %s' % (
                  cores, n, T, np.__version__))
    line = {
        'impl': 'reference', 'metric': METRIC, 'value': value, 'unit': 'pixel-timesteps/s',
        'n_gpus': args.gpus, 'steps': args.steps, 'warmup': args.warmup,
        'ms_per_step': 1e3 * wall / args.steps, 'higher_is_better': True, 'scaling': 'weak',
        'vs_baseline': None, 'dtype': 'f64', 'data': 'synthetic',
        'config': {'workload': workload_name(args), 'sample_units_per_step': cores * T * n},
        'cpu_baseline': {'value': value, 'unit': 'pixel-timesteps/s', 'cores': cores,
                         'kind': 'port', 'sample': sample},
        'e2e': {'value': value, 'unit': 'pixel-timesteps/s', 'h2d_bytes_per_step': 0,
                'd2h_bytes_per_step': 0},
    }
    emit(json.dumps(line))


def workload_name(args):
    if args.workload == 'global':
        return ('global 0.05 deg CMG grid 7200x3600 x %d daily steps pixel-tiled across the '
                'ranks, synthetic %s drivers' % (args.T, args.kind))
    if args.pixels == TILE_N and args.T == YEAR_T:
        return 'MODIS 500 m tile 2400x2400 x 365 daily steps, synthetic %s drivers' % args.kind
    return '%d pixels x %d steps per GPU, synthetic %s drivers' % (args.pixels, args.T, args.kind)



# ---- K2: objective batch (BASELINE configs 4 and 5) ---------------------------------------
def run_objective_bench(args):
    '''
    `--workload calibration`: D = 10^5 draws x 300 sites x 3652 days (30 % of the observations
    missing and removed by the compaction), tmin_close / tmin_open / vpd_open fixed;
    `--workload sobol`: D = 2^16 rows x 1000 pixels x 365 days, all eleven parameters free.
    One step = one `mod16_objective_batch_f32` call (all its launches) with drivers,
    observations and draws resident in HBM; the unit is draw-elements/s over the VALID
    elements.  Under torchrun `--shard draw` gives every rank all sites and D / world draws
    (all-gather of the results, no data-path collective), `--shard site` gives every rank
    N / world sites and all draws (ONE all-reduce of the (D, 5) un-gated statistics, then each
    draw's gate over all shards); the job is the same at every world size (strong scaling).
    The bound is FP32 issue, not HBM: no `roofline` object is claimed for this line.
    '''
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
    if args.workload == 'calibration':
        T, N, D, seed, fixed = 3652, 300, 100000, 16004, ('tmin_close', 'tmin_open', 'vpd_open')
    else:
        T, N, D, seed, fixed = 365, 1000, 65536, 16005, ()
    D = args.draws or D
    t_dtype = torch.float32 if args.dtype == 'float32' else torch.float64
    steps, warm = min(args.steps, 5), min(max(args.warmup, 1), 3)
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
    for _ in range(warm):
        sse, cnt = run()
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        sse, cnt = run()
    e1.record()
    barrier()
    tm = torch.tensor([e0.elapsed_time(e1) / steps], device=dev, dtype=torch.float64)
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
    emit(json.dumps({
        'metric': 'draw-elements/s (objective batch)', 'value': units / (ms * 1e-3),
        'unit': 'draw-elements/s', 'n_gpus': world, 'shard': args.shard if world > 1 else None,
        'scaling': 'strong', 'steps': steps, 'warmup': warm, 'ms_per_step': ms,
        'higher_is_better': True, 'vs_baseline': None, 'data': 'synthetic',
        'dtype': 'f32' if args.dtype == 'float32' else 'f64',
        'config': {'workload': args.workload, 'draws': D, 'T': T, 'N': N,
                   'valid_elements': valid, 'all_elements': T * N},
        'rmsd_range': [float(rm.min().item()), float(rm.max().item())],
        'cpu_baseline': cpu}))
    if world > 1:
        dist.destroy_process_group()


# ---- GPU arm -----------------------------------------------------------------------------
def resident_drivers(T, N, seed, device, torch_dtype, kind):
    'Fills the (T, N) driver block in HBM slab by slab (bounded generator temporaries).'
    import torch
    from mod16_b200 import synthetic
    names = [k for k in synthetic.DRIVER_NAMES
             if k not in synthetic.PER_PIXEL and k != 'sw_rad_night']
    big = dict((k, torch.empty((T, N), device=device, dtype=torch_dtype)) for k in names)
    slab = max(1, min(T, (1 << 27) // max(N, 1)))
    first = None
    for t0 in range(0, T, slab):
        t1 = min(T, t0 + slab)
        d = synthetic.device_drivers(t1 - t0, N, seed + t0, device, kind=kind, dtype=torch_dtype)
        if first is None:
            first = d
        for k in names:
            big[k][t0:t1].copy_(d[k])
        del d
    out = dict(big)
    out['temp_annual'] = first['temp_annual']
    out['pressure'] = first['pressure']
    out['sw_rad_night'] = 0
    return out


class _QuietStdout(object):
    '''
    Keeps stdout for the ONE JSON line: everything libraries print to file descriptor 1 in
    between (NCCL's version banner, for one) goes to stderr.
    '''

    def __init__(self):
        sys.stdout.flush()
        self.fd = os.dup(1)
        os.dup2(2, 1)

    def emit(self, line):
        sys.stdout.flush()
        os.write(self.fd, (line + '\n').encode())


_OUT = None


def emit(line):
    if _OUT is not None:
        _OUT.emit(line)
    else:
        print(line, flush=True)


def main():
    global _OUT
    args = parse()
    if int(os.environ.get('WORLD_SIZE', '1')) > 1:
        _OUT = _QuietStdout()
    if args.impl == 'reference':
        if int(os.environ.get('RANK', '0')) == 0:
            run_reference(args)
        return

    if args.workload in ('calibration', 'sobol'):
        return run_objective_bench(args)
    import torch
    import torch.distributed as dist
    from mod16_b200 import MOD16, synthetic, _abi
    from mod16_b200.forward import run_forward
    from mod16_b200.streaming import StreamedForward

    world = int(os.environ.get('WORLD_SIZE', '1'))
    rank = int(os.environ.get('RANK', '0'))
    local = int(os.environ.get('LOCAL_RANK', '0'))
    if not torch.cuda.is_available():
        raise SystemExit('bench.py needs a CUDA device (no CPU fallback); use --impl reference '
                         'for the CPU arm')
    torch.cuda.set_device(local)
    device = torch.device('cuda', local)
    from mod16_b200.streaming import bind_to_gpu_numa_node
    numa = bind_to_gpu_numa_node(local) if world > 1 else None
    if world > 1:
        dist.init_process_group('nccl', device_id=device)
    _abi.load()

    np_dtype = np.dtype(args.dtype)
    t_dtype = torch.float32 if np_dtype == np.float32 else torch.float64
    T, N = args.T, args.pixels
    if args.workload == 'global':
        from mod16_b200.distributed import pixel_range
        lo, hi = pixel_range(7200 * 3600, rank, world)
        N = hi - lo
        need = 48.0 * T * N / 1e9
        if need > 170.0:
            raise SystemExit('global workload: %.0f GB per GPU do not fit; use more GPUs or '
                             'a smaller --T' % need)
    table = bplut_table()
    pft = torch.from_numpy(synthetic.host_pft_map(N, 16002 + rank)).to(device)
    drv = resident_drivers(T, N, 16002 + 1000 * rank, device, t_dtype, args.kind)
    drv_list = [drv[k] for k in synthetic.DRIVER_NAMES]
    out = torch.empty((T, N), device=device, dtype=t_dtype)
    flag = torch.zeros(1, device=device, dtype=torch.int32)
    # per `_et` call: forward kernel + redo-queue kernel + overflow sweep (f32 pipeline), twice
    # when the call-exact global gate adds its fix-up pass
    launches_per_step = ((3 if (np_dtype == np.float32 and args.kernel == 'pipe') else 1) *
                         (2 if args.gate == 'global' else 1))

    force = {'pipe': 0, 'generic': _abi.FORCE_GENERIC,
             'mirror': _abi.FORCE_MIRROR}[args.kernel]

    def step():
        if args.workload == 'global' and args.gate == 'global' and world > 1:
            # one reference call over the WHOLE grid: the transpiration gate is evaluated
            # across the ranks (open pass, all-reduce(max) of one int, fix-up pass)
            run_forward(drv_list, pft=pft, table=table, semantics='static', out_mode='sum',
                        gate='open_track', dtype=np_dtype, device=device, gate_flag=flag,
                        out=[out], force=force)
            dist.all_reduce(flag, op=dist.ReduceOp.MAX)
            run_forward(drv_list, pft=pft, table=table, semantics='static', out_mode='sum',
                        gate='fixup', dtype=np_dtype, device=device, gate_flag=flag,
                        out=[out], force=force)
            flag.zero_()
            return
        run_forward(drv_list, pft=pft, table=table, semantics='static', out_mode='sum',
                    gate=args.gate, dtype=np_dtype, device=device, gate_flag=flag, out=[out],
                    force=force)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    sampler = ClockSampler(local)
    sampler.start()
    for _ in range(max(args.warmup, 3)):
        step()
    barrier()
    evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True))
           for _ in range(args.steps)]
    sampler.active.set()
    t_wall = time.perf_counter()
    for a, b in evs:
        a.record()
        step()
        b.record()
    barrier()
    t_wall = time.perf_counter() - t_wall
    sampler.active.clear()
    total_ms = evs[0][0].elapsed_time(evs[-1][1])
    step_ms = [a.elapsed_time(b) for a, b in evs]
    tm = torch.tensor([total_ms], device=device, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(tm, op=dist.ReduceOp.MAX)
    total_ms = float(tm.item())
    units_per_rank = T * N
    tu = torch.tensor([units_per_rank], device=device, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(tu, op=dist.ReduceOp.SUM)
    units_total = float(tu.item())
    value = units_total * args.steps / (total_ms * 1e-3)
    bytes_per_unit = BYTES_PER_UNIT_F32 * (2 if np_dtype == np.float64 else 1)
    kernel_ms = float(np.mean(step_ms))
    achieved = bytes_per_unit * units_per_rank / (kernel_ms * 1e-3) / 1e9

    # DRAM traffic of the dominant kernel from the committed `ncu --set full` capture
    # (bytes per pixel-timestep x the units of one launch), see profiles/traffic.json
    traffic = None
    tpath = os.path.join(ROOT, 'profiles', 'traffic.json')
    kname = ('forward_pipe_kernel' if args.kernel == 'pipe' else args.kernel) \
        if np_dtype == np.float32 else 'forward_strided_kernel<double>'
    if os.path.exists(tpath):
        rec = json.load(open(tpath)).get(kname)
        if rec:
            traffic = rec['dram_bytes_per_unit'] * units_per_rank
    peaks_path = os.path.join(ROOT, 'MEASURED_PEAKS.json')
    if os.path.exists(peaks_path):
        peak, peak_src = json.load(open(peaks_path))['hbm_gbs'], 'measured (MEASURED_PEAKS.json)'
    else:
        peak, peak_src = 6650.0, 'fallback (B200_PROFILING.md)'

    # a cheap sanity check that real work happened: ET of the resident block is finite
    # and of plausible size (the parity tests are the real check)
    probe = out[:1, :4096].float().cpu().numpy()
    assert np.isfinite(probe).all() and 0 < np.abs(probe).mean() < 1e3, 'implausible ET'

    e2e = None
    if not args.no_e2e:
        Te = min(args.e2e_T, T)
        names = [k for k in synthetic.DRIVER_NAMES
                 if k not in synthetic.PER_PIXEL and k != 'sw_rad_night']
        host = dict((k, drv[k][:Te].cpu().pin_memory()) for k in names)
        for k in synthetic.PER_PIXEL:
            host[k] = drv[k].cpu().pin_memory()
        host['sw_rad_night'] = 0
        out_host = torch.empty((Te, N), dtype=t_dtype).pin_memory()
        runner = StreamedForward(N, max(1, Te // 4), table, pft.cpu().numpy(), device=device,
                                 dtype=np_dtype, gate=args.gate)
        for _ in range(2):
            runner.run(host, out_host)
        barrier()
        runner.h2d_bytes = runner.d2h_bytes = runner.launches = 0
        reps = max(3, min(args.steps, 10))
        t0 = time.perf_counter()
        for _ in range(reps):
            runner.run(host, out_host)
        barrier()
        dt = time.perf_counter() - t0
        tt = torch.tensor([dt], device=device, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        dt = float(tt.item())
        e2e = {'value': world * reps * Te * N / dt, 'unit': 'pixel-timesteps/s',
               'h2d_bytes_per_step': runner.h2d_bytes // reps,
               'd2h_bytes_per_step': runner.d2h_bytes // reps,
               'sample': '%d steps x %d pixels per GPU from pinned host memory through '
                         'mod16_b200.streaming.StreamedForward (4 slabs, copies overlapped '
                         'with the kernel)%s' % (Te, N, '' if numa is None else
                                                 '; rank bound to NUMA node %d' % numa)}
        assert np.isfinite(out_host[:1, :4096].numpy()).all()

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu:   # (reported at N = 1 only)
        cpu = cpu_single(min(T, YEAR_T), args.cpu_pixels, args.kind)
    sampler.stop()
    clocks = sampler.summary()

    if rank == 0:
        line = {
            'metric': METRIC, 'value': value, 'unit': 'pixel-timesteps/s', 'n_gpus': world,
            'steps': args.steps, 'warmup': max(args.warmup, 3),
            'ms_per_step': total_ms / args.steps, 'higher_is_better': True,
            'scaling': 'strong' if args.workload == 'global' else 'weak', 'vs_baseline': None,
            'dtype': 'f32' if np_dtype == np.float32 else 'f64', 'data': 'synthetic',
            'config': {'workload': workload_name(args), 'pixels_per_gpu': N, 'steps_per_gpu': T,
                       'units_per_step': units_total, 'gate': args.gate,
                       'l2': 'inputs (%.1f GB per GPU) exceed L2 (126 MB); no flush needed' % (
                           bytes_per_unit * units_per_rank / 1e9)},
            'roofline': {'bound': 'hbm', 'achieved': achieved, 'peak': peak, 'unit': 'GB/s',
                         'frac': achieved / peak, 'traffic': traffic, 'peak_source': peak_src,
                         'frac_of_spec_8tbs': achieved / 8000.0,
                         'kernel': {'pipe': 'forward_pipe_kernel',
                                    'generic': 'forward_strided_fast_kernel',
                                    'mirror': 'forward_strided_kernel<float>'}[args.kernel]
                                   if np_dtype == np.float32 else 'forward_strided_kernel<double>',
                         'kernel_ms': kernel_ms, 'bytes_per_unit': bytes_per_unit},
            'e2e': e2e, 'gpu_launches': launches_per_step * args.steps,
            'clocks': clocks, 'cpu_baseline': cpu,
            'wall_s_timed_region': t_wall,
        }
        emit(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == '__main__':
    main()
