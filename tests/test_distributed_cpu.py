'''
CPU tests (gloo, world_size 2) of the multi-GPU host logic in mod16_b200/distributed.py.
The oracle stands in for the CUDA kernel (`forward=` hook), so what is exercised is the
partitioning, the two-pass gate protocol with its one-int all-reduce and the objective
reductions -- the parts that do not need a GPU.
'''
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from mod16_b200 import distributed as D
from mod16_b200 import synthetic
from oracle import mod16_oracle as orc

NAMES = list(synthetic.DRIVER_NAMES)
BPLUT7 = [-8., 8.8, 650., 4400., .02, .02, 1e-5, .0055, 60., 95., 250.]


def test_ranges_cover_and_align():
    for n, w in ((5760000, 8), (25920000, 8), (1001, 3), (7, 4), (0, 2), (300, 7)):
        edges = [D.pixel_range(n, r, w) for r in range(w)]
        assert edges[0][0] == 0 and edges[-1][1] == n
        for (a, b), (c, d) in zip(edges[:-1], edges[1:]):
            assert b == c and a <= b and a % 4 == 0
        assert max(b - a for a, b in edges) - min(b - a for a, b in edges) <= 4 + n % 4
        de = [D.draw_range(n, r, w) for r in range(w)]
        assert de[0][0] == 0 and de[-1][1] == n and all(b == c for (_, b), (c, _) in zip(de, de[1:]))


def _oracle_forward(drivers, *, params, gate, gate_flag, as_torch=True, out=None, **kw):
    'Stands in for run_forward: implements the open_track / fixup halves of the gate protocol.'
    drv = [d.numpy() if isinstance(d, torch.Tensor) else d for d in drivers]
    with np.errstate(all='ignore'):
        if gate == 'open_track':
            day, night = orc.static_evapotranspiration(params, *drv, gate='open')
            # the gate predicate on this shard, as the kernel evaluates it: g_surf > 0 (:343)
            g_surf = (params[7] * orc.ramp_up(drv[8] - 273.15, params[0], params[1]) *
                      orc.ramp_down(drv[9], params[2], params[3]))
            if np.any(g_surf / orc.r_correction(drv[11], drv[5]) > 0):
                gate_flag.fill_(1)
            return [torch.from_numpy(np.asarray(day + night))]
        assert gate == 'fixup'
        if int(gate_flag.item()) == 0:
            day, night = orc.static_evapotranspiration(params, *drv, gate='closed')
            out[0].copy_(torch.from_numpy(np.asarray(day + night)))
        return out


def _worker(rank, size, port, case, q):
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=size)
    try:
        T, N = 6, 64
        d = synthetic.host_drivers(T, N, 77)
        if case == 'closed':
            d['tmin'] = np.full((T, N), 240.0)          # nobody transpires: gate closed
        elif case == 'one_shard_open':
            d['tmin'] = np.full((T, N), 240.0)
            d['tmin'][:, N - 3:] = 290.0                # only the last rank sees g_surf > 0
        drv = [d[k] for k in NAMES]
        with np.errstate(all='ignore'):
            ref = orc.static_et(BPLUT7, *drv)           # one reference call on the whole domain
        lo, hi = D.pixel_range(N, rank, size)
        local = D.shard_pixels(drv, lo, hi)
        out = D.forward_global_gate(local, forward=_oracle_forward, params=BPLUT7,
                                    device=torch.device('cpu'))[0].numpy()
        ok = np.allclose(out, ref[:, lo:hi], rtol=1e-12, atol=1e-12, equal_nan=True)
        # objective partials sharded by site: all-reduce(sum) of 2 x D values
        obs = ref + 1.0
        sse = torch.tensor([float(np.nansum((out - obs[:, lo:hi])**2)), 2.0 * rank],
                           dtype=torch.float64)
        cnt = torch.tensor([float(out.size), 1.0], dtype=torch.float64)
        D.reduce_objective(sse, cnt)
        ok = ok and abs(sse[0].item() - T * N) < 1e-6 and cnt[0].item() == T * N \
            and sse[1].item() == 2.0 * sum(range(size))
        # draw-sharded results: all-gather back to the full vector
        dl, dh = D.draw_range(11, rank, size)
        full = D.gather_draws(torch.arange(dl, dh, dtype=torch.float64).reshape(-1, 1), 11)
        ok = ok and torch.equal(full.ravel(), torch.arange(11, dtype=torch.float64))
        q.put((rank, bool(ok)))
    finally:
        dist.destroy_process_group()


class _OracleBackend(object):
    'Stands in for ObjectiveBackend: un-gated (D, 5) statistics from the oracle, one call per draw.'
    semantics, gate = 'static', 'global'

    def __init__(self, drv, obs):
        self.drv, self.obs = drv, obs

    def _raw(self, draws):
        rows = []
        for x in draws:
            x = list(np.asarray(x, np.float64))
            with np.errstate(all='ignore'):
                so, no = orc.objective_stats(orc.static_et(x, *self.drv, gate='open'), self.obs)
                sc, nc = orc.objective_stats(orc.static_et(x, *self.drv, gate='closed'), self.obs)
                g = (x[7] * orc.ramp_up(self.drv[8] - 273.15, x[0], x[1]) *
                     orc.ramp_down(self.drv[9], x[2], x[3])) / orc.r_correction(
                         self.drv[11], self.drv[5])
            rows.append([so, no, sc, nc, float(np.any(g > 0))])
        return torch.tensor(rows, dtype=torch.float64)

    def __call__(self, draws, raw=False, as_torch=False):
        r = self._raw(draws)
        if raw:
            return r
        from mod16_b200.objective import finish_raw
        sse, cnt, _ = finish_raw(r)
        return sse, cnt


def _objective_worker(rank, size, port, q):
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=size)
    try:
        T, N, nd = 5, 48, 7
        d = synthetic.host_drivers(T, N, 91)
        d['tmin'] = d['tmin'].copy()
        d['tmin'][:, :N // 2] = 230.0            # rank 0's sites never open the Tmin ramp
        drv = [d[k] for k in NAMES]
        draws = synthetic.host_param_draws(nd, 5, bplut_column=np.array(BPLUT7))
        draws[2, 7] = 0.0                        # csl = 0: closed on every shard
        with np.errstate(all='ignore'):
            obs = orc.static_et(BPLUT7, *drv) + 3.0
        ref = [orc.objective_stats(orc.static_et(list(x), *drv), obs) for x in draws]
        ref_sse = np.array([r[0] for r in ref])
        ref_cnt = np.array([r[1] for r in ref])
        lo, hi = D.pixel_range(N, rank, size)
        mine = _OracleBackend(D.shard_pixels(drv, lo, hi), obs[:, lo:hi])
        sse, cnt = D.objective_site_sharded(mine, draws)
        ok = np.allclose(sse.numpy(), ref_sse, rtol=1e-12) and np.array_equal(cnt.numpy(), ref_cnt)
        everything = _OracleBackend(drv, obs)
        sse, cnt = D.objective_draw_sharded(everything, draws)
        ok = ok and np.allclose(sse.numpy(), ref_sse, rtol=1e-12) and \
            np.array_equal(cnt.numpy(), ref_cnt)
        q.put((rank, bool(ok)))
    finally:
        dist.destroy_process_group()


def test_objective_sharded_by_site_and_by_draw_world2():
    with socket.socket() as s:
        s.bind(('127.0.0.1', 0))
        port = s.getsockname()[1]
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    procs = [ctx.Process(target=_objective_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=180) for _ in procs)
    for p in procs:
        p.join(timeout=60)
    assert res == [(0, True), (1, True)], res


@pytest.mark.parametrize('case', ['open', 'closed', 'one_shard_open'])
def test_gate_protocol_and_reductions_world2(case):
    with socket.socket() as s:
        s.bind(('127.0.0.1', 0))
        port = s.getsockname()[1]
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, case, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(timeout=60)
    assert res == [(0, True), (1, True)], res


# ---- T-chunked execution (mod16_b200/chunked.py): config 3 on two GPUs ---------------------
def test_chunk_plan_covers_and_fits():
    from mod16_b200.chunked import plan_chunks
    # global 0.05 degree year, two ranks, float32, 120 GiB for two buffer sets
    N = 7200 * 3600 // 2
    ch = plan_chunks(365, N, 4, 120 << 30, n_buffers=2)
    assert ch[0][0] == 0 and ch[-1][1] == 365
    assert all(b == c for (_, b), (c, _) in zip(ch, ch[1:]))
    assert max(b - a for a, b in ch) * 2 * 11 * N * 4 <= 120 << 30
    assert all((b - a) % 2 == 0 for a, b in ch[:-1])
    assert len(plan_chunks(365, 5760000, 4, 170 << 30, n_buffers=1)) == 1   # a tile-year fits
    assert len(plan_chunks(365, 5760000, 8, 170 << 30, n_buffers=1)) == 2   # ... not in float64
    with pytest.raises(ValueError):
        plan_chunks(10, 1 << 30, 4, 1 << 20)


def _oracle_forward_chunk(drivers, *, gate, gate_flag, out, pft=None, table=None, **kw):
    'Stands in for run_forward in ChunkedForward: open_track / closed on one chunk.'
    params = BPLUT7
    drv = [d.numpy() if isinstance(d, torch.Tensor) else d for d in drivers]
    with np.errstate(all='ignore'):
        day, night = orc.static_evapotranspiration(
            params, *drv, gate='open' if gate == 'open_track' else 'closed')
        if gate == 'open_track':
            g_surf = (params[7] * orc.ramp_up(drv[8] - 273.15, params[0], params[1]) *
                      orc.ramp_down(drv[9], params[2], params[3]))
            if np.any(g_surf / orc.r_correction(drv[11], drv[5]) > 0):
                gate_flag.fill_(1)
    out[0].copy_(torch.from_numpy(np.asarray(day + night)))
    return out


def _chunk_worker(rank, size, port, case, q):
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=size)
    try:
        from mod16_b200.chunked import ChunkedForward, PER_STEP
        T, N = 11, 32
        d = synthetic.host_drivers(T, N, 78)
        if case != 'open':
            d['tmin'] = np.full((T, N), 240.0)
        if case == 'last_chunk_of_last_rank_open':
            d['tmin'][T - 1, N - 1] = 290.0
        drv = [d[k] for k in NAMES]
        with np.errstate(all='ignore'):
            ref = orc.static_et(BPLUT7, *drv)           # ONE reference call: all days, all pixels
        lo, hi = D.pixel_range(N, rank, size)
        n = hi - lo
        budget = 2 * len(PER_STEP) * n * 8 * 4          # two buffer sets of four steps
        cf = ChunkedForward(T, n, None, None, torch.from_numpy(d['temp_annual'][lo:hi]),
                            torch.from_numpy(d['pressure'][lo:hi]), device='cpu',
                            dtype=np.float64, budget_bytes=budget, forward=_oracle_forward_chunk)
        calls = []

        def fill(k, t0, t1, bufs):
            calls.append((k, t0, t1))
            for name in PER_STEP:
                bufs[name].copy_(torch.from_numpy(d[name][t0:t1, lo:hi]))
        out = torch.empty((T, n), dtype=torch.float64)
        cf.run(fill, out)
        ok = len(cf.chunks) == 3 and cf.chunks[-1] == (8, 11)
        ok = ok and np.allclose(out.numpy(), ref[:, lo:hi], rtol=1e-12, atol=1e-12, equal_nan=True)
        # closed over the whole call: every chunk is filled and evaluated twice
        ok = ok and len(calls) == (6 if case == 'closed' else 3)
        q.put((rank, bool(ok)))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize('case', ['open', 'closed', 'last_chunk_of_last_rank_open'])
def test_chunked_forward_gate_spans_chunks_and_ranks_world2(case):
    with socket.socket() as s:
        s.bind(('127.0.0.1', 0))
        port = s.getsockname()[1]
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    procs = [ctx.Process(target=_chunk_worker, args=(r, 2, port, case, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(timeout=60)
    assert res == [(0, True), (1, True)], res
