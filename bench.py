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

The same line carries `secondary`: bounded runs of the other BASELINE.json configurations on
the same box and ranks -- the reference's own call form (per-pixel parameter arrays), the
float64 entry, the global 0.05 degree year pixel-tiled across the ranks with the
transpiration gate all-reduced inside the timed region (T-chunked where it does not fit),
and the objective batches K2 of the calibration and Sobol configurations (sharded by draw
and by site under torchrun).

`--impl reference` times the CPU implementation of the path -- the UNMODIFIED reference
`MOD16._et` from oracle/_ref when it has been materialised (oracle/make_ref.py), else the
numpy restatement in oracle/ -- on all host cores, on a bounded sample of the same workload,
in the reference's stock precision (float32 drivers + float64 BPLUT: numpy promotes to
float64) and with float32 parameters (float32 throughout); `value` is the faster leg.
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
    ap.add_argument('--e2e-T', type=int, default=YEAR_T,
                    help='time steps of the end-to-end leg (bounded by free host memory)')
    ap.add_argument('--cpu-pixels', type=int, default=20000)
    ap.add_argument('--no-cpu', action='store_true')
    ap.add_argument('--no-e2e', action='store_true')
    ap.add_argument('--no-secondary', action='store_true',
                    help='skip the bounded runs of the other configurations')
    ap.add_argument('--gate', default='global')
    ap.add_argument('--format', default='float32', choices=['float32', 'modis', 'packed'],
                    help='storage format of the RESIDENT drivers of the headline loop '
                         '(profiling the storage-format kernels; implies --no-secondary)')
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


def _cpu_model():
    '''
    The CPU implementation that is timed: the unmodified reference (oracle/_ref, made by
    oracle/make_ref.py in the build container) or, when that is absent, the numpy port.
    Returns (callable `_et(params, *drivers)`, kind).
    '''
    from oracle import make_ref
    mod = make_ref.load()
    if mod is not None:
        return mod.MOD16._et, 'reference'
    from oracle import mod16_oracle as orc
    return orc.static_et, 'port'


def _cpu_prepare(key, seed, T, n, kind, par_dtype):
    'Builds (once per process) the inputs of the CPU model for one pixel chunk.'
    if key not in _CPU_DATA:
        from mod16_b200 import synthetic
        d = synthetic.host_drivers(T, n, seed, kind=kind, dtype=np.float32)
        table = bplut_table()
        pft = synthetic.host_pft_map(n, seed + 1)
        # the reference callers' form: params[key][pft_map].reshape((1, N))
        vec = [table[i][pft].reshape(1, n).astype(par_dtype) for i in range(11)]
        _CPU_DATA[key] = (vec, [d[k] for k in synthetic.DRIVER_NAMES])
    return _CPU_DATA[key]


def _cpu_chunk(args):
    """
    One worker: `_et` on its pixel chunk.  With run=False only the inputs are built, so that
    the timed region holds the model evaluation alone.  par_dtype float64 is the reference's
    stock path (numpy promotes the float32 drivers), float32 computes in float32 throughout.
    """
    seed, T, n, kind, run, par_dtype = args
    fn, _ = _cpu_model()
    vec, drv = _cpu_prepare((os.getpid(), T, n, kind, par_dtype), seed + os.getpid() % 9973, T, n,
                            kind, np.dtype(par_dtype))
    if not run:
        return 0.0, 0.0
    t0 = time.perf_counter()
    with np.errstate(all='ignore'):
        import warnings
        with warnings.catch_warnings():
            warnings.simplefilter('ignore')
            out = fn(vec, *drv)
    dt = time.perf_counter() - t0
    return dt, float(np.nansum(out))


def cpu_single(T, n, kind):
    _, impl = _cpu_model()
    legs = {}
    for par_dtype in ('float64', 'float32'):
        _cpu_chunk((16002, T, n, kind, False, par_dtype))
        dt, _ = _cpu_chunk((16002, T, n, kind, True, par_dtype))
        legs[par_dtype] = T * n / dt
    best = max(legs, key=legs.get)
    return {'value': legs[best], 'unit': 'pixel-timesteps/s', 'cores': 1, 'kind': impl,
            'legs': {'f64_stock_promotion': legs['float64'], 'f32_throughout': legs['float32']},
            'sample': '%s MOD16._et, per-pixel BPLUT parameters (1, N), %d pixels x %d steps of '
                      'the workload, 1 process, numpy %s; value = the faster of the two legs '
                      '(%s parameters)' % (
                          'unmodified reference (oracle/_ref)' if impl == 'reference' else
                          'numpy port (oracle/mod16_oracle.py) of', n, T, np.__version__, best)}


def run_reference(args):
    'The reference arm: the CPU implementation of the path on all host cores.'
    import multiprocessing as mp
    cores = os.cpu_count() or 1
    try:
        cores = len(os.sched_getaffinity(0))
    except Exception:
        pass
    T, n = args.T, args.cpu_pixels
    try:  # the numpy path holds ~42 full-size temporaries: bound the sample by free memory
        import psutil
        budget = 0.4 * psutil.virtual_memory().available / cores
        n = int(max(256, min(n, budget / (400.0 * T))))
    except Exception:
        pass
    _, impl = _cpu_model()
    ctx = mp.get_context('spawn')
    legs = {}
    with ctx.Pool(cores) as pool:
        for par_dtype in ('float64', 'float32'):
            jobs = lambda run: [(16002, T, n, args.kind, run, par_dtype)] * cores
            pool.map(_cpu_chunk, jobs(False), chunksize=1)      # imports, inputs, page-in
            for _ in range(args.warmup):
                pool.map(_cpu_chunk, jobs(True), chunksize=1)
            t0 = time.perf_counter()
            for _ in range(args.steps):
                pool.map(_cpu_chunk, jobs(True), chunksize=1)
            legs[par_dtype] = time.perf_counter() - t0
    units = args.steps * cores * T * n
    best = min(legs, key=legs.get)
    wall = legs[best]
    value = units / wall
    sample = ('%s MOD16._et through its public call form (per-pixel parameter arrays (1, N)), '
              '%d processes x %d pixels x %d steps per bench step; float32 drivers; value = the '
              'faster leg (%s parameters); numpy %s' % (
                  'unmodified reference (oracle/_ref/mod16)' if impl == 'reference' else
                  'numpy port (oracle/mod16_oracle.py; oracle/_ref absent) of', cores, n, T,
                  best, np.__version__))
    line = {
        'impl': 'reference', 'metric': METRIC, 'value': value, 'unit': 'pixel-timesteps/s',
        'n_gpus': args.gpus, 'steps': args.steps, 'warmup': args.warmup,
        'ms_per_step': 1e3 * wall / args.steps, 'higher_is_better': True, 'scaling': 'weak',
        'vs_baseline': None, 'dtype': 'f32' if best == 'float32' else 'f64', 'data': 'synthetic',
        'config': bench_config(args, args.gpus, args.pixels, args.T),
        'cpu_baseline': {'value': value, 'unit': 'pixel-timesteps/s', 'cores': cores,
                         'kind': impl, 'sample': sample,
                         'sample_units_per_step': cores * T * n,
                         'legs': {'f64_stock_promotion': units / legs['float64'],
                                  'f32_throughout': units / legs['float32']}},
        'e2e': {'value': value, 'unit': 'pixel-timesteps/s', 'h2d_bytes_per_step': 0,
                'd2h_bytes_per_step': 0},
    }
    emit(json.dumps(line))


def bench_config(args, world, N, T):
    'The `config` object of the headline line -- the same for both arms.'
    bpu = BYTES_PER_UNIT_F32 * (2 if args.dtype == 'float64' else 1)
    return {'workload': workload_name(args), 'pixels_per_gpu': N, 'steps_per_gpu': T,
            'units_per_step': float(T) * N * (world if args.workload != 'global' else 1),
            'gate': args.gate,
            'l2': 'inputs (%.1f GB per GPU) exceed L2 (126 MB); no flush needed' % (
                bpu * T * N / 1e9)}


def workload_name(args):
    if args.workload == 'global':
        return ('global 0.05 deg CMG grid 7200x3600 x %d daily steps pixel-tiled across the '
                'ranks, synthetic %s drivers' % (args.T, args.kind))
    if args.pixels == TILE_N and args.T == YEAR_T:
        return 'MODIS 500 m tile 2400x2400 x 365 daily steps, synthetic %s drivers' % args.kind
    return '%d pixels x %d steps per GPU, synthetic %s drivers' % (args.pixels, args.T, args.kind)



# ---- K2: objective batch (BASELINE configs 4 and 5) ---------------------------------------
def objective_bench(workload, *, dev, world, rank, shard='draw', draws=None, dtype='float32',
                    steps=3, warm=1, cpu_draws=2):
    """
    `calibration`: D = 10^5 draws x 300 sites x 3652 days (30 % of the observations missing
    and removed by the compaction), tmin_close / tmin_open / vpd_open fixed;
    `sobol`: D = 2^16 rows x 1000 pixels x 365 days, all eleven parameters free.
    One step = one `mod16_objective_batch_f32` call (all its launches) with drivers,
    observations and draws resident in HBM; the unit is draw-elements/s over the VALID
    elements.  With world > 1 `shard='draw'` gives every rank all sites and D / world draws
    (all-gather of the results, no data-path collective), `shard='site'` gives every rank
    N / world sites and all draws (ONE all-reduce of the (D, 5) un-gated statistics, then each
    draw's gate over all shards); the job is the same at every world size (strong scaling).
    The bound is the FP32 pipe (packed instructions) together with the special-function
    unit, not HBM: `pipe` reports the fraction of the packed-FP32 issue rate the kernel's
    51 packed instructions per pair of draws and element amount to.
    Returns the result dict on rank 0, None elsewhere.
    """
    import torch
    import torch.distributed as dist
    from mod16_b200 import synthetic, utils, MOD16
    from mod16_b200.objective import ObjectiveBackend
    from mod16_b200.forward import run_forward
    from mod16_b200 import distributed as mdist
    if workload == 'calibration':
        T, N, D, seed, fixed = 3652, 300, 100000, 16004, ('tmin_close', 'tmin_open', 'vpd_open')
    else:
        T, N, D, seed, fixed = 365, 1000, 65536, 16005, ()
    D = draws or D
    t_dtype = torch.float32 if dtype == 'float32' else torch.float64
    table = utils.bplut_table(utils.restore_bplut(utils.BPLUT_C51), MOD16.required_parameters,
                              default_beta=250.0)
    col = np.array(table[:, 7])
    d = synthetic.device_drivers(T, N, seed, dev, dtype=t_dtype)
    drv = [d[k] for k in synthetic.DRIVER_NAMES]
    truth = run_forward(drv, params=list(col), dtype=np.dtype(dtype), device=dev,
                        as_torch=True)[0]
    g = torch.Generator(device=dev)
    g.manual_seed(seed)
    obs = truth + 15.0 * torch.randn(truth.shape, generator=g, device=dev, dtype=t_dtype)
    obs[obs < 0] = float('nan')
    obs[torch.rand(truth.shape, generator=g, device=dev) < 0.3] = float('nan')
    w = 0.2 + 0.8 * torch.rand(truth.shape, generator=g, device=dev, dtype=t_dtype)
    pdraws = torch.from_numpy(synthetic.host_param_draws(D, seed, bplut_column=col,
                                                         fixed=fixed)).to(dev)
    # (every rank generated the same data: same seeds)
    if world > 1 and shard == 'site':
        lo, hi = mdist.pixel_range(N, rank, world, align=1)
        local = mdist.shard_pixels(drv, lo, hi)
        backend = ObjectiveBackend(local, obs[:, lo:hi], w[:, lo:hi], dtype=np.dtype(dtype),
                                   device=dev)
        run = lambda: mdist.objective_site_sharded(backend, pdraws)
    elif world > 1:
        backend = ObjectiveBackend(drv, obs, w, dtype=np.dtype(dtype), device=dev)
        run = lambda: mdist.objective_draw_sharded(backend, pdraws)
    else:
        backend = ObjectiveBackend(drv, obs, w, dtype=np.dtype(dtype), device=dev)
        run = lambda: backend(pdraws, as_torch=True)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
    sampler = ClockSampler(dev.index)
    sampler.start()
    for _ in range(warm):
        sse, cnt = run()
    barrier()
    sampler.active.set()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        sse, cnt = run()
    e1.record()
    barrier()
    sampler.active.clear()
    sampler.stop()
    tm = torch.tensor([e0.elapsed_time(e1) / steps], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(tm, op=dist.ReduceOp.MAX)
    ms = float(tm.item())
    valid = float(cnt.max().item())
    units = valid * D
    rm = torch.sqrt(sse / cnt)
    if rank != 0:
        return None
    # the reference loop on the host: one numpy `_et` + nanmean per draw
    cpu = None
    if cpu_draws > 0:
        from oracle import mod16_oracle as orc
        fn, impl = _cpu_model()
        h = [v.cpu().numpy().astype(np.float64) if isinstance(v, torch.Tensor) else v for v in drv]
        ho, hw = obs.cpu().numpy().astype(np.float64), w.cpu().numpy().astype(np.float64)
        hd = pdraws[:cpu_draws].cpu().numpy()
        t0 = time.perf_counter()
        ref = []
        with np.errstate(all='ignore'):
            for x in hd:
                sr, cr = orc.objective_stats(fn(list(x), *h), ho, hw)
                ref.append(np.sqrt(sr / cr))
        dt = time.perf_counter() - t0
        got = rm[:cpu_draws].cpu().numpy()
        cpu = {'value': cpu_draws * T * N / dt, 'unit': 'draw-elements/s (all elements)',
               'cores': 1, 'kind': impl, 'sample': '%d draws' % cpu_draws,
               'rmsd_max_abs_diff': float(np.max(np.abs(got - np.array(ref))))}
    clocks = sampler.summary()
    # packed-FP32 pipe: one FFMA2-class warp instruction per two clocks per SM sub-partition
    # = 32 lanes x 2 draws... per pair of draws and element the fast kernel issues 51 packed
    # instructions: rate bound = 148 SMs x 4 x clk / 2 x 32 pairs x 2 draws / 51
    clk = (clocks['sm_mhz'] or 1965.0) * 1e6
    pipe_peak = 148 * 4 * clk / 2.0 * 64.0 / 51.0 * world
    value = units / (ms * 1e-3)
    return {
        'metric': 'draw-elements/s (objective batch)', 'value': value,
        'unit': 'draw-elements/s', 'n_gpus': world, 'shard': shard if world > 1 else None,
        'scaling': 'strong', 'steps': steps, 'warmup': warm, 'ms_per_step': ms,
        'higher_is_better': True, 'vs_baseline': None, 'data': 'synthetic',
        'dtype': 'f32' if dtype == 'float32' else 'f64',
        'config': {'workload': workload, 'draws': D, 'T': T, 'N': N,
                   'valid_elements': valid, 'all_elements': T * N},
        'rmsd_range': [float(rm.min().item()), float(rm.max().item())],
        'pipe': {'bound': 'fp32 packed issue (FFMA2 class, 1 per 2 clk per SM sub-partition)',
                 'achieved': value, 'peak': pipe_peak, 'unit': 'draw-elements/s',
                 'frac': value / pipe_peak, 'packed_instructions_per_pair_element': 51},
        'clocks': clocks, 'cpu_baseline': cpu}


def run_objective_bench(args):
    'Stand-alone `--workload calibration|sobol` line.'
    import torch
    import torch.distributed as dist
    world = int(os.environ.get('WORLD_SIZE', '1'))
    rank = int(os.environ.get('RANK', '0'))
    dev = torch.device('cuda', int(os.environ.get('LOCAL_RANK', '0')))
    torch.cuda.set_device(dev)
    if world > 1:
        dist.init_process_group('nccl', device_id=dev)
    res = objective_bench(args.workload, dev=dev, world=world, rank=rank, shard=args.shard,
                          draws=args.draws, dtype=args.dtype, steps=min(args.steps, 5),
                          warm=min(max(args.warmup, 1), 3), cpu_draws=args.cpu_draws)
    if res is not None:
        emit(json.dumps(res))
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


def fill_synthetic(bufs, rows, N, seed, device, torch_dtype, kind):
    'Writes `rows` time steps of synthetic per-step drivers into the buffers, slab by slab.'
    from mod16_b200 import synthetic
    slab = max(1, min(rows, (1 << 27) // max(N, 1)))
    for t0 in range(0, rows, slab):
        t1 = min(rows, t0 + slab)
        d = synthetic.device_drivers(t1 - t0, N, seed + t0, device, kind=kind, dtype=torch_dtype)
        for k, v in bufs.items():
            v[t0:t1].copy_(d[k])
        del d


def all_ok(ok, world, device):
    'True when every rank says so (one int all-reduce): ranks skip a leg together or not at all.'
    import torch
    import torch.distributed as dist
    if world == 1:
        return bool(ok)
    t = torch.tensor([1 if ok else 0], device=device, dtype=torch.int32)
    dist.all_reduce(t, op=dist.ReduceOp.MIN)
    return bool(t.item())


def hbm_peak():
    peaks_path = os.path.join(ROOT, 'MEASURED_PEAKS.json')
    if os.path.exists(peaks_path):
        return json.load(open(peaks_path))['hbm_gbs'], 'measured (MEASURED_PEAKS.json)'
    return 6650.0, 'fallback (B200_PROFILING.md)'


def global_year(args, device, world, rank, steps, warm):
    """
    BASELINE configs[2]: the global 0.05 degree CMG grid (7200 x 3600 pixels) x 365 daily
    steps as ONE `_et` call, pixel-tiled across the ranks (strong scaling).  A rank's shard
    that does not fit in HBM (one or two GPUs) is walked in T-chunks through two alternating
    buffer sets (mod16_b200/chunked.py); the transpiration gate is evaluated
    over all chunks and ranks -- its all-reduce is inside the timed region.  Throughput is
    over the device time of the kernels and the gate exchange; refilling a buffer set with
    the next chunk's synthetic drivers is not in it ("kernel time over
    resident chunks").
    """
    import torch
    import torch.distributed as dist
    from mod16_b200 import synthetic
    from mod16_b200.chunked import ChunkedForward, PER_STEP
    from mod16_b200.distributed import pixel_range
    np_dtype = np.dtype(args.dtype)
    t_dtype = torch.float32 if np_dtype == np.float32 else torch.float64
    T = args.T
    lo, hi = pixel_range(7200 * 3600, rank, world)
    N = hi - lo
    free, total = torch.cuda.mem_get_info(device)
    out_bytes = T * N * np_dtype.itemsize
    budget = int(min(free, 170e9) - out_bytes - 6e9)
    resident = len(PER_STEP) * T * N * np_dtype.itemsize <= budget
    err = None
    cf = out = None
    try:
        table = bplut_table()
        pft = torch.from_numpy(synthetic.host_pft_map(N, 16002 + rank)).to(device)
        g = torch.Generator(device=device)
        g.manual_seed(16002 + rank)
        lo_b, hi_b = synthetic.BOUNDS['temp_annual']
        t_ann = torch.empty(N, device=device, dtype=t_dtype).uniform_(lo_b, hi_b, generator=g)
        lo_b, hi_b = synthetic.BOUNDS['pressure']
        prs = torch.empty(N, device=device, dtype=t_dtype).uniform_(lo_b, hi_b, generator=g)
        cf = ChunkedForward(T, N, table, pft, t_ann, prs, device=device, dtype=np_dtype,
                            budget_bytes=budget, n_buffers=1 if resident else 2)
        out = torch.empty((T, N), device=device, dtype=t_dtype)
    except Exception as exc:  # (out of memory on a small GPU ...)
        err = '%s: %s' % (type(exc).__name__, exc)
    if not all_ok(err is None, world, device):
        return {'name': 'global', 'error': err or 'another rank failed to set up'}
    filled = set()

    def fill(k, t0, t1, bufs):
        if resident and k in filled:
            return          # the whole shard stays resident between steps
        fill_synthetic(bufs, t1 - t0, N, 16002 + 1000 * rank + 7 * t0, device, t_dtype, args.kind)
        filled.add(k)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
    sampler = ClockSampler(device.index)
    sampler.start()
    for _ in range(warm):
        cf.run(fill, out)
    barrier()
    sampler.active.set()
    ms = 0.0
    for _ in range(steps):
        cf.run(fill, out)
        ms += cf.kernel_ms
    barrier()
    sampler.active.clear()
    sampler.stop()
    tm = torch.tensor([ms / steps], device=device, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(tm, op=dist.ReduceOp.MAX)
    ms = float(tm.item())
    probe = out[:1, :4096].float().cpu().numpy()
    assert np.isfinite(probe).all() and 0 < np.abs(probe).mean() < 1e3, 'implausible ET'
    units = float(T) * 7200 * 3600
    bpu = BYTES_PER_UNIT_F32 * (2 if np_dtype == np.float64 else 1)
    peak, peak_src = hbm_peak()
    achieved = bpu * units / (ms * 1e-3) / 1e9
    res = {
        'name': 'global', 'metric': METRIC, 'value': units / (ms * 1e-3),
        'unit': 'pixel-timesteps/s', 'n_gpus': world, 'scaling': 'strong', 'steps': steps,
        'warmup': warm, 'ms_per_step': ms, 'dtype': 'f32' if np_dtype == np.float32 else 'f64',
        'config': {'workload': 'global 0.05 deg CMG grid 7200x3600 x %d daily steps, one `_et` '
                               'call pixel-tiled across %d rank(s), synthetic %s drivers' % (
                                   T, world, args.kind),
                   'pixels_per_gpu': N, 'units_per_step': units,
                   'time_chunks_per_rank': [list(c) for c in cf.chunks],
                   'resident': resident,
                   'timed': 'device time of the chunk kernels + the cross-rank gate all-reduce '
                            '(max over ranks)%s' % ('' if resident else
                                                     '; refills between chunks not included'),
                   'gate': 'global (open pass per chunk, all-reduce(max) of one int over %d '
                           'rank(s), closed pass only if nobody transpired)' % world},
        'roofline': {'bound': 'hbm', 'achieved': achieved, 'peak': peak * world, 'unit': 'GB/s',
                     'frac': achieved / (peak * world), 'peak_source': peak_src,
                     'frac_of_spec_8tbs': achieved / (8000.0 * world),
                     'kernel': 'forward_pipe_kernel', 'bytes_per_unit': bpu},
        'gpu_launches': cf.launches * steps, 'comm': {'collective': 'all_reduce(max) int32[1]',
                                                       'per_step': 1, 'nranks': world},
        'clocks': sampler.summary()}
    del cf, out
    torch.cuda.empty_cache()
    return res if rank == 0 else None


# int16 packings of the meteorological drivers for the PACKED storage layout (netCDF-style
# scale / offset): name -> (scale, offset, fill)
PACKED_MET = {'lw_net_day': (0.01, 0.0, -32768), 'lw_net_night': (0.01, 0.0, -32768),
              'sw_rad_day': (0.02, 0.0, -32768), 'temp_day': (0.01, 273.15, -32768),
              'temp_night': (0.01, 273.15, -32768), 'tmin': (0.01, 273.15, -32768),
              'vpd_day': (0.25, 0.0, -32768), 'vpd_night': (0.25, 0.0, -32768)}
MODIS_NATIVE = {'fpar': (np.uint8, 0.01, 0.0, 255), 'lai': (np.uint8, 0.1, 0.0, 255),
                'sw_albedo': (np.int16, 0.001, 0.0, 32767)}


def end_to_end(args, device, world, rank, numa, drv, pft, table, T, N, np_dtype, t_dtype,
               barrier):
    """
    The same metric through the public API with HOST buffers: every step's drivers are copied
    from pinned host memory to the device and its ET back inside the timed region
    (mod16_b200.streaming.StreamedForward: slabs of time steps, copies overlapped with the
    kernel).  Three input formats, same physical values:
      packed  -- the drivers as an archive stores them: fPAR / LAI uint8, albedo and the
                 meteorological fields int16 with scale / offset (20 B per pixel-timestep);
                 the fused kernel converts while it reads (mod16_forward_raw_f32)
      modis   -- fPAR / LAI uint8, albedo int16, meteorology float32 (36 B)
      float32 -- every driver a float32 array, the reference's in-memory form (44 B)
    `e2e` reports the packed leg (host -> ET is PCIe-bound: bytes per unit are what counts);
    the other two are listed under `legs`.
    """
    import torch
    import torch.distributed as dist
    from mod16_b200 import rawio, synthetic
    from mod16_b200.streaming import StreamedForward
    names = [k for k in synthetic.DRIVER_NAMES
             if k not in synthetic.PER_PIXEL and k != 'sw_rad_night']
    def drop_pinned():
        # pinned blocks go back to torch's host cache, not to the system: without this the
        # three legs' buffers add up (and, times the ranks, exhaust the host)
        import gc
        gc.collect()
        torch.cuda.synchronize()
        if hasattr(torch._C, '_host_emptyCache'):
            torch._C._host_emptyCache()

    try:  # a third of the free host memory may be pinned at a time, over all ranks
        import psutil
        room = psutil.virtual_memory().available / 3.0 / max(world, 1)
        if not hasattr(torch._C, '_host_emptyCache'):
            room /= 3.0        # (the legs' buffers cannot be returned: share the budget)
    except Exception:
        room = 16 * 12 * N * np_dtype.itemsize
    tr = torch.tensor([room], device=device, dtype=torch.float64)
    if world > 1:             # (every rank measured at a different moment: agree on the least)
        dist.all_reduce(tr, op=dist.ReduceOp.MIN)
    room = float(tr.item())
    per_pixel = dict((k, drv[k].cpu().pin_memory()) for k in synthetic.PER_PIXEL)
    legs = {}
    layouts = ['float32'] + (['modis', 'packed'] if np_dtype == np.float32 else [])
    for layout in layouts:
        fmt = {}
        if layout in ('modis', 'packed'):
            fmt.update(MODIS_NATIVE)
        if layout == 'packed':
            fmt.update(dict((k, (np.int16,) + v) for k, v in PACKED_MET.items()))
        # as many time steps of the resident block as fit the pinned budget in this format
        # (inputs + the float output): the whole year where the host has the memory
        step_bytes = N * (sum(np.dtype(fmt[k][0]).itemsize if k in fmt else np_dtype.itemsize
                              for k in names) + np_dtype.itemsize)
        Te = max(4, min(args.e2e_T, T, int(room // step_bytes)))
        tq = torch.tensor([Te], device=device, dtype=torch.int64)
        if world > 1:
            dist.all_reduce(tq, op=dist.ReduceOp.MIN)
        Te = int(tq.item())
        slab = max(1, min(8, Te // 4))
        host = dict(per_pixel)
        host['sw_rad_night'] = 0
        for k in names:
            if k in fmt:
                q = rawio.quantize(drv[k][:Te], fmt[k][0], *fmt[k][1:])
                pinned = torch.empty(q.raw.shape, dtype=q.raw.dtype, pin_memory=True)
                pinned.copy_(q.raw)
                host[k] = rawio.RawDriver(pinned, q.scale, q.offset, q.fill)
                del q
            else:
                host[k] = torch.empty((Te, N), dtype=t_dtype, pin_memory=True)
                host[k].copy_(drv[k][:Te])
        out_host = torch.empty((Te, N), dtype=t_dtype, pin_memory=True)
        runner = StreamedForward(N, slab, table, pft.cpu().numpy(), device=device,
                                 dtype=np_dtype, gate=args.gate, formats=fmt)
        runner.run(host, out_host)
        barrier()
        runner.h2d_bytes = runner.d2h_bytes = runner.launches = 0
        reps = 3 if Te > 64 else max(3, min(args.steps, 10))
        t0 = time.perf_counter()
        for _ in range(reps):
            runner.run(host, out_host)
        barrier()
        dt = time.perf_counter() - t0
        tt = torch.tensor([dt], device=device, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        dt = float(tt.item())
        ok = bool(np.isfinite(out_host[:1, :4096].numpy()).all())
        assert ok, 'implausible ET from the end-to-end leg'
        legs[layout] = {'value': world * reps * Te * N / dt, 'unit': 'pixel-timesteps/s',
                        'steps_of_the_block': Te,
                        'h2d_bytes_per_step': runner.h2d_bytes // reps,
                        'd2h_bytes_per_step': runner.d2h_bytes // reps,
                        'h2d_bytes_per_unit': runner.h2d_bytes / reps / (Te * N)}
        del host, out_host, runner
        torch.cuda.empty_cache()
        drop_pinned()
    head = 'packed' if 'packed' in legs else 'float32'
    e2e = dict(legs[head])
    e2e['input_format'] = {
        'packed': 'storage formats: fPAR / LAI uint8, albedo + meteorology int16 (scale, '
                  'offset, fill), converted inside the fused kernel; float32 arithmetic and '
                  'output; parity: tests/test_raw_formats.py',
        'float32': 'float32 arrays'}[head]
    e2e['legs'] = legs
    e2e['sample'] = ('%d of the %d steps x %d pixels per GPU from pinned host memory through '
                     'mod16_b200.streaming.StreamedForward (slabs of %d steps, copies '
                     'overlapped with the kernel)%s' % (
                         legs[head]['steps_of_the_block'], T, N, slab,
                         '' if numa is None else '; rank bound to NUMA node %d' % numa))
    return e2e


def secondary_lines(args, device, world, rank, drv_list, drv, pft, table, out):
    """
    Bounded runs of the other BASELINE.json configurations (and of the other entry points of
    the headline one) on the same box and ranks; each entry carries its own timing, roofline
    or pipe fraction and clocks.  `drv_list` is the resident float32 tile-year of the
    headline; it is released before the legs that need the memory.
    """
    import torch
    import torch.distributed as dist
    from mod16_b200 import synthetic
    from mod16_b200.forward import run_forward
    lines = []
    T, N = out.shape
    peak, peak_src = hbm_peak()

    def timed(fn, steps, warm=2):
        for _ in range(warm):
            fn()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            fn()
        e1.record()
        torch.cuda.synchronize()
        tm = torch.tensor([e0.elapsed_time(e1) / steps], device=device, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(tm, op=dist.ReduceOp.MAX)
        return float(tm.item())

    # (1) the reference's own call form: MOD16._et(params, *drivers) with per-pixel parameter
    #     arrays params[k] = table[k][pft_map].reshape(1, N) (Forward_Run_example.ipynb cell 32)
    err = None
    try:
        tab_t = torch.from_numpy(np.asarray(table, np.float32)).to(device)
        vec = [tab_t[k][pft.long()].reshape(1, N).contiguous() for k in range(11)]
        flag = torch.zeros(1, device=device, dtype=torch.int32)
        ref_out = out[:2].clone()
        fn = lambda: run_forward(drv_list, params=vec, semantics='static', out_mode='sum',
                                 gate='global', dtype=np.float32, device=device,
                                 gate_flag=flag, out=[out])
    except Exception as exc:
        err = '%s: %s' % (type(exc).__name__, exc)
    if all_ok(err is None, world, device):
        ms = timed(fn, 5)
        diff = float((torch.nan_to_num(out[:2]) - torch.nan_to_num(ref_out)).abs().max().item())
        ach = BYTES_PER_UNIT_F32 * T * N / (ms * 1e-3) / 1e9
        if rank == 0:
            lines.append({
                'name': 'tile-year, reference call form (per-pixel parameter arrays)',
                'metric': METRIC, 'value': world * T * N / (ms * 1e-3),
                'unit': 'pixel-timesteps/s', 'n_gpus': world, 'ms_per_step': ms, 'steps': 5,
                'dtype': 'f32', 'max_abs_diff_vs_pft_gather_w_m2': diff,
                'roofline': {'bound': 'hbm', 'achieved': ach, 'peak': peak, 'unit': 'GB/s',
                             'frac': ach / peak, 'kernel': 'forward_pipe_kernel',
                             'peak_source': peak_src}})
        del vec
    else:
        lines.append({'name': 'tile-year, reference call form', 'error': err})

    # (2) the float64 entry (1e-10 path) on a slab of the same tile
    err = None
    try:
        Ts = 16
        d64 = [v[:Ts].double() if isinstance(v, torch.Tensor) and v.ndim == 2 else
               (v.double() if isinstance(v, torch.Tensor) else v) for v in drv_list]
        out64 = torch.empty((Ts, N), device=device, dtype=torch.float64)
        flag = torch.zeros(1, device=device, dtype=torch.int32)
        fn = lambda: run_forward(d64, pft=pft, table=table, semantics='static', out_mode='sum',
                                 gate='global', dtype=np.float64, device=device,
                                 gate_flag=flag, out=[out64])
    except Exception as exc:
        err = '%s: %s' % (type(exc).__name__, exc)
    if all_ok(err is None, world, device):
        ms = timed(fn, 3, warm=1)
        ach = 2 * BYTES_PER_UNIT_F32 * Ts * N / (ms * 1e-3) / 1e9
        if rank == 0:
            lines.append({
                'name': 'float64 entry, %d-step slab of the tile' % Ts, 'metric': METRIC,
                'value': world * Ts * N / (ms * 1e-3), 'unit': 'pixel-timesteps/s',
                'n_gpus': world, 'ms_per_step': ms, 'steps': 3, 'dtype': 'f64',
                'roofline': {'bound': 'hbm', 'achieved': ach, 'peak': peak, 'unit': 'GB/s',
                             'frac': ach / peak, 'kernel': f64_kernel_name(),
                             'peak_source': peak_src, 'bytes_per_unit': 96}})
        del d64, out64
    else:
        lines.append({'name': 'float64 entry', 'error': err})
    return lines


def _named(name, line):
    'A secondary line with a `name` key like the others (None stays None: non-zero ranks).'
    if isinstance(line, dict) and 'name' not in line:
        line = dict([('name', name)] + list(line.items()))
    return line


def f64_kernel_name():
    'The float64 kernel the LAST launch of this process used (mod16_last_forward_kernel()).'
    from mod16_b200 import _abi
    k = _abi.load().mod16_last_forward_kernel()
    if k & 0xff == _abi.KERNEL_EXACT:
        return 'forward_strided_kernel'
    return 'forward_canon64_kernel' if (k >> 8) & 1 else 'forward_fast64_kernel'


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

    if args.workload == 'global':
        res = global_year(args, device, world, rank, args.steps, max(args.warmup, 3))
        if rank == 0:
            res.update({'higher_is_better': True, 'vs_baseline': None, 'data': 'synthetic'})
            emit(json.dumps(res))
        if world > 1:
            dist.destroy_process_group()
        return

    np_dtype = np.dtype(args.dtype)
    t_dtype = torch.float32 if np_dtype == np.float32 else torch.float64
    T, N = args.T, args.pixels
    table = bplut_table()
    pft = torch.from_numpy(synthetic.host_pft_map(N, 16002 + rank)).to(device)
    drv = resident_drivers(T, N, 16002 + 1000 * rank, device, t_dtype, args.kind)
    drv_list = [drv[k] for k in synthetic.DRIVER_NAMES]
    bytes_in_per_unit = 44 if np_dtype == np.float32 else 88
    if args.format != 'float32':
        from mod16_b200 import rawio
        fmt = dict(MODIS_NATIVE)
        if args.format == 'packed':
            fmt.update(dict((k, (np.int16,) + v) for k, v in PACKED_MET.items()))
        for i, k in enumerate(synthetic.DRIVER_NAMES):
            if k in fmt:
                drv_list[i] = rawio.quantize(drv[k], fmt[k][0], *fmt[k][1:])
                drv[k] = None
        bytes_in_per_unit = {'modis': 36, 'packed': 20}[args.format]
        args.no_secondary = args.no_e2e = True
        torch.cuda.empty_cache()
    out = torch.empty((T, N), device=device, dtype=t_dtype)
    flag = torch.zeros(1, device=device, dtype=torch.int32)
    # per `_et` call: forward kernel + redo-queue kernel + overflow sweep (f32 pipeline), twice
    # when the call-exact global gate adds its fix-up pass
    launches_per_step = ((3 if (np_dtype == np.float32 and args.kernel == 'pipe') else 1) *
                         (2 if args.gate == 'global' else 1))

    force = {'pipe': 0, 'generic': _abi.FORCE_GENERIC,
             'mirror': _abi.FORCE_MIRROR}[args.kernel]

    def step():
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
    units_total = float(units_per_rank) * world
    value = units_total * args.steps / (total_ms * 1e-3)
    bytes_per_unit = bytes_in_per_unit + np_dtype.itemsize
    kernel_ms = float(np.mean(step_ms))
    achieved = bytes_per_unit * units_per_rank / (kernel_ms * 1e-3) / 1e9

    # DRAM traffic of the dominant kernel from the committed `ncu --set full` capture
    # (bytes per pixel-timestep x the units of one launch), see profiles/traffic.json; the
    # entry names the commit of the kernel source it was captured from
    traffic = traffic_src = None
    tpath = os.path.join(ROOT, 'profiles', 'traffic.json')
    kname = ('forward_pipe_kernel' if args.kernel == 'pipe' else args.kernel) \
        if np_dtype == np.float32 else f64_kernel_name()
    if os.path.exists(tpath):
        rec = json.load(open(tpath)).get(kname)
        if rec:
            traffic = rec['dram_bytes_per_unit'] * units_per_rank
            traffic_src = '%s; kernel source at commit %s' % (
                rec.get('source', ''), rec.get('kernel_source_commit', '?'))
    peak, peak_src = hbm_peak()

    # a cheap sanity check that real work happened: ET of the resident block is finite
    # and of plausible size (the parity tests are the real check)
    probe = out[:1, :4096].float().cpu().numpy()
    assert np.isfinite(probe).all() and 0 < np.abs(probe).mean() < 1e3, 'implausible ET'

    e2e = None
    if not args.no_e2e:
        e2e = end_to_end(args, device, world, rank, numa, drv, pft, table, T, N, np_dtype,
                         t_dtype, barrier)

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu:   # (reported at N = 1 only)
        cpu = cpu_single(min(T, YEAR_T), args.cpu_pixels, args.kind)
    sampler.stop()
    clocks = sampler.summary()

    secondary = None
    if not args.no_secondary and args.kernel == 'pipe' and np_dtype == np.float32:
        secondary = secondary_lines(args, device, world, rank, drv_list, drv, pft, table, out)
        del drv, drv_list, out
        torch.cuda.empty_cache()
        g_args = argparse.Namespace(**vars(args))
        g_args.T = YEAR_T
        legs = [lambda: global_year(g_args, device, world, rank, 2, 1)]
        for wl in ('calibration', 'sobol'):
            for shard in (('draw', 'site') if (world > 1 and wl == 'calibration') else ('draw',)):
                legs.append(lambda wl=wl, shard=shard: _named(
                    'objective batch (K2), %s workload, sharded by %s' % (wl, shard),
                    objective_bench(wl, dev=device, world=world, rank=rank, shard=shard,
                                    steps=3, warm=2, cpu_draws=1 if world == 1 else 0)))
        for leg in legs:
            res = leg()
            torch.cuda.empty_cache()
            if rank == 0 and res is not None:
                secondary.append(res)

    if rank == 0:
        line = {
            'metric': METRIC, 'value': value, 'unit': 'pixel-timesteps/s', 'n_gpus': world,
            'steps': args.steps, 'warmup': max(args.warmup, 3),
            'ms_per_step': total_ms / args.steps, 'higher_is_better': True,
            'scaling': 'weak', 'vs_baseline': None,
            'dtype': 'f32' if np_dtype == np.float32 else 'f64', 'data': 'synthetic',
            'config': bench_config(args, world, N, T),
            'roofline': {'bound': 'hbm', 'achieved': achieved, 'peak': peak, 'unit': 'GB/s',
                         'frac': achieved / peak, 'traffic': traffic,
                         'traffic_source': traffic_src, 'peak_source': peak_src,
                         'frac_of_spec_8tbs': achieved / 8000.0,
                         'kernel': {'pipe': 'forward_pipe_kernel',
                                    'generic': 'forward_strided_fast_kernel',
                                    'mirror': 'forward_strided_kernel<float>'}[args.kernel]
                                   if np_dtype == np.float32 else kname,
                         'kernel_ms': kernel_ms, 'bytes_per_unit': bytes_per_unit},
            'e2e': e2e, 'gpu_launches': launches_per_step * args.steps,
            'clocks': clocks, 'cpu_baseline': cpu,
            'wall_s_timed_region': t_wall,
            'secondary': secondary,
        }
        emit(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == '__main__':
    main()
