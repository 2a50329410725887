'''
Multi-GPU host logic of the forward path: one process per GPU (``torch.distributed``), work
partitioned by pixel range (forward runs) or by parameter draw / by site (objective
batches).  Every (pixel, day, draw) is independent, so the data path needs NO collective;
the only exchanges are

* one int, all-reduce(max): the static path's batch-global transpiration gate
  ``np.any(g_surf > 0)`` (reference mod16/__init__.py:343) made call-exact across shards;
* ``2 x D`` float64, all-reduce(sum): the per-draw objective partials when an objective
  batch is sharded by site (sharding by draw needs none: ``gather_draws``).

The functions take the process group explicitly and work with any backend (NCCL on the
GPUs; the CPU tests run them over gloo with the oracle standing in for the kernel).
'''
import numpy as np
import torch
import torch.distributed as dist

from .forward import run_forward


def world(group=None):
    'Returns (rank, world_size); (0, 1) when torch.distributed is not initialised.'
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(group), dist.get_world_size(group)
    return 0, 1


def pixel_range(n_pixels, rank, world_size, align=4):
    '''
    Contiguous pixel range [lo, hi) of ``rank``: ``align``-pixel granules dealt out as evenly
    as possible (4 pixels = 16 bytes keeps every shard eligible for the bulk-copy kernel),
    the remainder going to the last rank.
    '''
    granules = n_pixels // align
    base, extra = divmod(granules, world_size)
    lo = (rank * base + min(rank, extra)) * align
    hi = lo + (base + (1 if rank < extra else 0)) * align
    if rank == world_size - 1:
        hi = n_pixels
    return lo, hi


def draw_range(n_draws, rank, world_size):
    'Contiguous range of parameter draws [lo, hi) of ``rank``.'
    base, extra = divmod(n_draws, world_size)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def shard_pixels(arrays, lo, hi):
    '''
    Slices the pixel (last) axis of every array-like driver / parameter to [lo, hi); scalars
    and arrays of pixel extent 1 (broadcast) pass through.
    '''
    out = []
    for a in arrays:
        if isinstance(a, (np.ndarray, torch.Tensor)) and a.ndim >= 1 and a.shape[-1] > 1:
            out.append(a[..., lo:hi])
        else:
            out.append(a)
    return out


def forward_global_gate(drivers, *, group=None, forward=run_forward, device=None, **kw):
    '''
    ``_et`` / ``_evapotranspiration`` of this rank's pixel shard with the transpiration gate
    evaluated over ALL shards, i.e. exactly as one reference call on the whole domain would:
    pass 1 computes with the gate open and ORs the predicate into a device int, the int is
    all-reduced (max), pass 2 recomputes with transpiration = 0 only if no shard saw
    ``g_surf > 0``.  ``kw`` goes to ``run_forward`` (params / pft + table, out_mode, dtype ...).
    '''
    rank, size = world(group)
    if device is None:
        device = torch.device('cuda', torch.cuda.current_device()) \
            if torch.cuda.is_available() else torch.device('cpu')
    flag = torch.zeros(1, device=device, dtype=torch.int32)
    outs = forward(drivers, gate='open_track', gate_flag=flag, as_torch=True, **kw)
    if size > 1:
        dist.all_reduce(flag, op=dist.ReduceOp.MAX, group=group)
    forward(drivers, gate='fixup', gate_flag=flag, out=outs, as_torch=True, **kw)
    return outs


def objective_site_sharded(backend, draws, group=None):
    '''
    Objective batch with the SITES (elements) sharded across ranks: ``backend`` is this rank's
    ``ObjectiveBackend`` over its own sites, every rank passes the same (D, 11) ``draws``.
    One all-reduce(sum) of the (D, 5) un-gated statistics; each draw's gate is then taken
    over all shards, exactly as one reference call per draw on the whole data set would
    (``np.any(g_surf > 0)``, reference mod16/__init__.py:343).  Returns (sse, count) tensors.
    '''
    from .objective import finish_raw
    raw = backend(draws, raw=True)
    rank, size = world(group)
    if size > 1:
        dist.all_reduce(raw, op=dist.ReduceOp.SUM, group=group)
    sse, cnt, _ = finish_raw(raw, backend.semantics, backend.gate)
    return sse, cnt


def objective_draw_sharded(backend, draws, group=None):
    '''
    Objective batch with the DRAWS sharded across ranks (every rank holds all sites): rank r
    evaluates ``draw_range(D, r, world)``; the per-draw results are all-gathered.  No
    data-path collective.  Returns (sse, count) tensors of D entries on every rank.
    '''
    rank, size = world(group)
    D = len(draws)
    lo, hi = draw_range(D, rank, size)
    sse, cnt = backend(draws[lo:hi], as_torch=True)
    both = gather_draws(torch.stack([sse, cnt], dim=1), D, group)
    return both[:, 0], both[:, 1]


def reduce_objective(sse, count, group=None):
    '''
    Site-sharded objective batch: sums the per-draw partials (float64 tensors of D entries)
    over the ranks with ONE all-reduce of 2 x D values.  Returns (sse, count) in place.
    '''
    rank, size = world(group)
    if size > 1:
        both = torch.stack([sse, count])
        dist.all_reduce(both, op=dist.ReduceOp.SUM, group=group)
        sse.copy_(both[0])
        count.copy_(both[1])
    return sse, count


def gather_draws(local, n_draws, group=None):
    '''
    Draw-sharded objective batch: every rank evaluated ``draw_range(n_draws, rank, world)``;
    returns the full (n_draws, ...) tensor on every rank (all-gather, padded to equal chunks).
    '''
    rank, size = world(group)
    if size == 1:
        return local
    chunk = (n_draws + size - 1) // size
    pad = torch.zeros((chunk,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    pad[:local.shape[0]] = local
    parts = [torch.empty_like(pad) for _ in range(size)]
    dist.all_gather(parts, pad, group=group)
    out = []
    for r, p in enumerate(parts):
        lo, hi = draw_range(n_draws, r, size)
        out.append(p[:hi - lo])
    return torch.cat(out, dim=0)
