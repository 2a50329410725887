'''
TEST HARNESS ONLY: builds tests/hostmath/hostmath.cpp (the kernels' per-element math,
compiled for the host) and exposes it through ctypes.  Used by the `-m "not gpu"` tests to
check the device math against the oracle where no GPU exists.  Never imported by the
product package.
'''
import ctypes
import os
import subprocess
import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
SO = os.path.join(HERE, '_build', 'libhostmath.so')
SRC = os.path.join(HERE, 'hostmath.cpp')
CSRC = os.path.join(os.path.dirname(os.path.dirname(HERE)), 'mod16_b200', 'csrc')


def build():
    deps = [SRC] + [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith('.cuh')]
    if os.path.exists(SO) and all(os.path.getmtime(d) <= os.path.getmtime(SO) for d in deps):
        return SO
    os.makedirs(os.path.dirname(SO), exist_ok=True)
    fast = os.path.exists(os.path.join(CSRC, 'mod16_fast.cuh'))
    cmd = ['g++', '-O2', '-std=c++17', '-fPIC', '-shared', '-ffp-contract=off',
           '-o', SO, SRC] + (['-DHM_WITH_FAST'] if fast else [])
    subprocess.run(cmd, check=True)
    return SO


_lib = None


def lib():
    global _lib
    if _lib is None:
        _lib = ctypes.CDLL(build())
    return _lib


def _flat(arrs, dtype):
    b = np.broadcast_arrays(*[np.asarray(a, dtype=dtype) for a in arrs])
    shape = b[0].shape
    return np.ascontiguousarray(np.stack([x.ravel() for x in b])), shape


def forward(params, drivers, semantics='static', gate='global', dtype=np.float64,
            r_corr_list=None, fast=False):
    '''
    Evaluates the kernel math on the host.  Returns dict(day=(canopy, soil, trans),
    night=(...), g_pos=bool array) with arrays of the broadcast shape.
    '''
    allv, shape = _flat(list(drivers) + list(params) +
                        (list(r_corr_list) if r_corr_list is not None else []), dtype)
    n = allv.shape[1]
    drv = np.ascontiguousarray(allv[:14])
    par = np.ascontiguousarray(allv[14:25])
    if fast:
        # the fast path receives the BPLUT columns in double, as the kernels do
        par = np.ascontiguousarray(np.stack([
            np.broadcast_to(np.asarray(p, np.float64), shape).ravel() for p in params]))
    rc = np.ascontiguousarray(allv[25:27]) if r_corr_list is not None else np.zeros((2, 1), dtype)
    out = np.empty((9 if fast else 6, n), dtype)
    gpos = np.empty(n, np.uint8)
    sem = {'static': 0, 'instance': 1}[semantics]
    g = {'open': 0, 'global': 0, 'element': 1, 'closed': 2}[gate]
    suffix = 'f64' if dtype == np.float64 else 'f32'
    name = (('hm_fastpair_' if fast == 'pair' else 'hm_fast_') if fast else 'hm_forward_') + suffix
    fn = getattr(lib(), name)
    fn.restype = None
    fn(ctypes.c_int(sem), ctypes.c_int(g), drv.ctypes.data_as(ctypes.c_void_p),
       par.ctypes.data_as(ctypes.c_void_p), rc.ctypes.data_as(ctypes.c_void_p),
       ctypes.c_int(1 if r_corr_list is not None else 0), ctypes.c_int64(n),
       out.ctypes.data_as(ctypes.c_void_p), gpos.ctypes.data_as(ctypes.c_void_p))
    out = out.reshape((out.shape[0],) + shape)
    guarded = (gpos.reshape(shape) & 2).astype(bool)
    gpos = (gpos.reshape(shape) & 1).astype(bool)
    closed = semantics == 'static' and gate == 'global' and not gpos.any()
    if closed:
        out[2] = 0  # the batch-global gate is closed (:343-348)
    res = dict(day=(out[0], out[1], out[2]), night=(out[3], out[4], out[5]), g_pos=gpos,
               guarded=guarded)
    if fast and not closed:
        res.update(tot_day=out[6], tot_night=out[7], tot=out[8])
    return res


def objective(draws, drivers, pair=False):
    '''
    K2 math on the host: for every (draw, element) the gate-open prediction, the day
    transpiration and the gate quantity.  Returns dict(tot, tr, gs: (D, n) float32 arrays
    (NaN where the element has no fast record), fast: bool (n,), draw_ok: bool (D,)).
    '''
    allv, shape = _flat(list(drivers), np.float32)
    n = allv.shape[1]
    drv = np.ascontiguousarray(allv[:14])
    dr = np.ascontiguousarray(np.asarray(draws, np.float64).reshape(-1, 11))
    D = dr.shape[0]
    tot = np.full((D, n), np.nan, np.float32)
    tr = np.full((D, n), np.nan, np.float32)
    gs = np.full((D, n), np.nan, np.float32)
    cls = np.zeros(n, np.uint8)
    ok = np.zeros(D, np.uint8)
    fn = lib().hm_objective_f32
    fn.restype = None
    p = lambda a: a.ctypes.data_as(ctypes.c_void_p)
    fn(p(drv), p(dr), ctypes.c_int64(D), ctypes.c_int64(n), ctypes.c_int(1 if pair else 0),
       p(tot), p(tr), p(gs), p(cls), p(ok))
    rs = lambda a: a.reshape((D,) + shape)
    return dict(tot=rs(tot), tr=rs(tr), gs=rs(gs), fast=(cls == 1).reshape(shape),
                draw_ok=ok.astype(bool))


def nan_halves(params, drivers, semantics='static'):
    '''
    The masked-driver rule of the fast path: returns (h, g_pos) with h bit 0 / 1 set where the
    day / night total is NaN by the rule (0: the rule does not decide).
    '''
    allv, shape = _flat(list(drivers), np.float32)
    n = allv.shape[1]
    drv = np.ascontiguousarray(allv[:14])
    par = np.ascontiguousarray(np.stack([
        np.broadcast_to(np.asarray(p, np.float64), shape).ravel() for p in params]))
    h = np.zeros(n, np.uint8)
    g = np.zeros(n, np.uint8)
    fn = lib().hm_nan_halves
    fn.restype = None
    p = lambda a: a.ctypes.data_as(ctypes.c_void_p)
    fn(ctypes.c_int({'static': 0, 'instance': 1}[semantics]), p(drv), p(par), ctypes.c_int64(n),
       p(h), p(g))
    return h.reshape(shape), g.reshape(shape).astype(bool)


def exp2_log2(x):
    '''d_exp2(x), d_log2(x) of csrc/mod16_fast.cuh (the float64 entry's own functions).'''
    x = np.ascontiguousarray(x, np.float64).ravel()
    ex, lg = np.empty_like(x), np.empty_like(x)
    fn = lib().hm_exp2_log2
    fn.restype = None
    fn.argtypes = [ctypes.c_void_p, ctypes.c_int64, ctypes.c_void_p, ctypes.c_void_p]
    fn(x.ctypes.data, x.size, ex.ctypes.data, lg.ctypes.data)
    return ex, lg
