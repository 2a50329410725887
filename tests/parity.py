'''
Shared parity criteria (BASELINE.json north_star):
  float64 path: within 1e-10 relative of the reference (absolute floor 1e-10 x unit scale,
                because ET is a signed sum that passes through zero -- SURVEY.md H8);
  float32 path: within 1e-5 relative or 1e-3 W m-2 absolute, whichever is looser;
  identical NaN masks, identical infinities.
The unit scale is 1 for W m-2 (static path) and 1 / 2.45e6 for kg m-2 s-1 (instance path).
'''
import numpy as np

W_SCALE = 1.0
MASS_SCALE = 1.0 / 2.45e6


def mismatch(got, ref, dtype, scale=W_SCALE):
    '''
    Returns a dict: n, nan_mismatch, inf_mismatch, n_fail, max_rel, max_abs, worst index.
    '''
    got = np.asarray(got, np.float64)
    ref = np.broadcast_to(np.asarray(ref, np.float64), got.shape)
    nan_mis = int((np.isnan(got) != np.isnan(ref)).sum())
    inf = np.isinf(got) | np.isinf(ref)
    with np.errstate(all='ignore'):
        inf_mis = int((inf & (got != ref)).sum())
        ok = np.isfinite(got) & np.isfinite(ref)
        err = np.where(ok, np.abs(got - ref), 0.0)
        if np.dtype(dtype) == np.float64:
            tol = 1e-10 * np.maximum(np.abs(ref), scale)
        else:
            tol = np.maximum(1e-5 * np.abs(ref), 1e-3 * scale)
        tol = np.where(ok, tol, np.inf)
        fail = err > tol
        rel = np.where(ok, err / np.maximum(np.abs(ref), 1e-300), 0.0)
    worst = np.unravel_index(np.argmax(err / np.where(ok, tol, 1.0)), got.shape) if got.size else ()
    return dict(n=int(got.size), nan_mismatch=nan_mis, inf_mismatch=inf_mis,
                n_fail=int(fail.sum()), max_rel=float(rel.max()) if got.size else 0.0,
                max_abs=float(err.max()) if got.size else 0.0, worst=worst)


def assert_parity(got, ref, dtype, scale=W_SCALE, what='', allow_fail=0):
    m = mismatch(got, ref, dtype, scale)
    assert np.shape(got) == np.shape(ref) or np.broadcast_shapes(np.shape(got), np.shape(ref)) == np.shape(got), \
        '%s: shape %r vs %r' % (what, np.shape(got), np.shape(ref))
    assert m['nan_mismatch'] == 0, '%s: NaN masks differ in %d of %d elements' % (
        what, m['nan_mismatch'], m['n'])
    assert m['inf_mismatch'] == 0, '%s: infinities differ (%d)' % (what, m['inf_mismatch'])
    assert m['n_fail'] <= allow_fail, '%s: %d of %d elements out of tolerance (max rel %.3e, ' \
        'max abs %.3e, worst at %r)' % (what, m['n_fail'], m['n'], m['max_rel'], m['max_abs'],
                                        m['worst'])
    return m
