'''
T-chunked, double-buffered forward runs of a pixel shard whose year of drivers does not fit
in HBM at once (a MODIS tile-year in float64 is 202 GB, the global 0.05 degree
year on two GPUs 227 GB per GPU in float32).

The shard's (T, N) domain is split along time into chunks that fit a memory budget; two
device buffer sets alternate, so that the caller's ``fill`` of chunk k + 1 (a host->device
stream, a decoder, a generator ...) runs on its own stream while the fused kernel evaluates
chunk k.  ONE reference call covers the whole domain -- ``MOD16._et(params, *drivers)`` over
all pixels and days, notebooks/MOD16_Forward_Run_example.ipynb cell 35 -- so the static
path's transpiration gate ``np.any(g_surf > 0)`` (mod16/__init__.py:343) spans all chunks
AND all ranks: every chunk is evaluated with the gate open while the predicate is OR-ed into
one device int; after the last chunk the int is all-reduced (max) across the ranks; only if
it stayed 0 everywhere (no element of the whole call had g_surf > 0) are the chunks
evaluated again with transpiration = 0.
'''
import numpy as np
import torch
import torch.distributed as dist

from .forward import run_forward
from .synthetic import DRIVER_NAMES, PER_PIXEL

PER_STEP = [k for k in DRIVER_NAMES if k not in PER_PIXEL and k != 'sw_rad_night']


def plan_chunks(T, N, itemsize, budget_bytes, n_buffers=2, n_arrays=len(PER_STEP), even=True):
    '''
    Splits T time steps into the fewest equal chunks such that ``n_buffers`` sets of
    ``n_arrays`` driver arrays of one chunk fit ``budget_bytes``.  Returns [(t0, t1), ...].
    Chunks are even-sized (the kernel pairs time steps) except the last.
    '''
    per_step = n_buffers * n_arrays * N * itemsize
    max_t = int(budget_bytes // per_step)
    if max_t < 1:
        raise ValueError('not even one time step of %d pixels fits %d bytes' % (N, budget_bytes))
    n_chunks = -(-T // max_t)
    size = -(-T // n_chunks)
    if even and size % 2 and size < max_t:
        size += 1
    return [(t0, min(T, t0 + size)) for t0 in range(0, T, size)]


class ChunkedForward(object):
    '''
    Parameters
    ----------
    T, N : time steps of the call and pixels of THIS rank's shard
    table, pft : dense 11 x 13 BPLUT and the shard's uint8 land-cover codes
    temp_annual, pressure : the shard's per-pixel drivers ((N,) device tensors or scalars)
    budget_bytes : device memory the driver buffers may use (both sets together)
    group : process group of the ranks that share the call (None: default group, or a
            single process)
    '''

    def __init__(self, T, N, table, pft, temp_annual, pressure, *, device, dtype=np.float32,
                 budget_bytes=120 << 30, n_buffers=2, sw_rad_night=0.0, group=None,
                 forward=run_forward):
        self.T, self.N = int(T), int(N)
        self.device = torch.device(device)
        self.np_dtype = np.dtype(dtype)
        self.t_dtype = torch.float32 if self.np_dtype == np.float32 else torch.float64
        self.table = np.asarray(table, np.float64)
        self.pft = pft
        self.per_pixel = {'temp_annual': temp_annual, 'pressure': pressure}
        self.sw_rad_night = sw_rad_night
        self.group = group
        self.forward = forward
        self.chunks = plan_chunks(self.T, self.N, self.np_dtype.itemsize, budget_bytes, n_buffers)
        rows = max(t1 - t0 for t0, t1 in self.chunks)
        on_gpu = self.device.type == 'cuda'
        self.bufs = [dict((k, torch.empty((rows, self.N), device=self.device, dtype=self.t_dtype))
                          for k in PER_STEP) for _ in range(min(n_buffers, len(self.chunks)))]
        self.flag = torch.zeros(1, device=self.device, dtype=torch.int32)
        self.s_fill = torch.cuda.Stream(self.device) if on_gpu else None
        self.kernel_ms = 0.0       # device time of the kernels + the gate exchange, last run()
        self.launches = 0

    def _size(self):
        if dist.is_available() and dist.is_initialized():
            return dist.get_world_size(self.group)
        return 1

    def _pass(self, fill, out, gate, timed):
        on_gpu = self.device.type == 'cuda'
        cur = torch.cuda.current_stream(self.device) if on_gpu else None
        nb = len(self.bufs)
        filled = [None] * nb     # event: fill of the set finished
        freed = [None] * nb      # event: kernel that read the set finished

        def start_fill(k):
            b = k % nb
            t0, t1 = self.chunks[k]
            view = dict((n, v[:t1 - t0]) for n, v in self.bufs[b].items())
            if not on_gpu:
                fill(k, t0, t1, view)
                return
            with torch.cuda.stream(self.s_fill):
                if freed[b] is not None:
                    self.s_fill.wait_event(freed[b])
                fill(k, t0, t1, view)
                filled[b] = torch.cuda.Event()
                filled[b].record(self.s_fill)

        if on_gpu:
            self.s_fill.wait_stream(cur)
        start_fill(0)
        for k, (t0, t1) in enumerate(self.chunks):
            b = k % nb
            if k + 1 < len(self.chunks) and nb > 1:
                start_fill(k + 1)
            if on_gpu:
                cur.wait_event(filled[b])
            drv = [self.bufs[b][n][:t1 - t0] if n in self.bufs[b] else
                   (self.per_pixel[n] if n in self.per_pixel else self.sw_rad_night)
                   for n in DRIVER_NAMES]
            if on_gpu and timed is not None:
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record(cur)
            self.forward(drv, pft=self.pft, table=self.table, semantics='static', out_mode='sum',
                         gate=gate, dtype=self.np_dtype, device=self.device, gate_flag=self.flag,
                         out=[out[t0:t1]], as_torch=True)
            self.launches += 3 if self.np_dtype == np.float32 else 1
            if on_gpu:
                if timed is not None:
                    e1.record(cur)
                    timed.append((e0, e1))
                freed[b] = torch.cuda.Event()
                freed[b].record(cur)
            if k + 1 < len(self.chunks) and nb == 1:
                start_fill(k + 1)

    def run(self, fill, out):
        '''
        fill(k, t0, t1, bufs) : writes the per-step drivers of time steps [t0, t1) of this
            rank's pixels into ``bufs[name][: t1 - t0]`` (device tensors); it is called on a
            side stream (current inside the call) one chunk ahead of the kernel.
        out : (T, N) device tensor receiving ET [W m-2].
        Returns ``out``; ``self.kernel_ms`` holds the device time of the kernels and of the
        cross-rank gate exchange (the fills are not in it).
        '''
        on_gpu = self.device.type == 'cuda'
        timed = [] if on_gpu else None
        self.launches = 0
        self.flag.zero_()
        self._pass(fill, out, 'open_track', timed)
        if on_gpu:
            g0, g1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            g0.record()
        if self._size() > 1:
            dist.all_reduce(self.flag, op=dist.ReduceOp.MAX, group=self.group)
        if on_gpu:
            g1.record()
            timed.append((g0, g1))
        # (the one host read of the call; a closed gate over a whole call does not occur with
        # real daytime data, so the second pass is the rare case)
        if int(self.flag.item()) == 0:
            self._pass(fill, out, 'closed', timed)
        if on_gpu:
            torch.cuda.synchronize(self.device)
            self.kernel_ms = float(sum(a.elapsed_time(b) for a, b in timed))
        return out
