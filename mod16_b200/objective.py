'''
Host side of K2: many parameter draws evaluated on the SAME drivers and reduced on the GPU
to the sufficient statistics of the callers' objectives (``mod16_objective_batch_*`` of
``include/mod16_b200.h``).

The reference's objective loops call the model once per draw and reduce on the host:

* ``for X in param_sweep: yhat = MOD16._et(X, *drivers); nash_sutcliffe(yhat, obs, norm=True)``
  (reference mod16/sensitivity.py:94-96);
* mod17's ``BlackBoxLikelihood(model=MOD16._et, observed, x=drivers, weights, objective=
  'rmsd')`` evaluating ``-sqrt(nanmean(((model(params, *x) - observed) * weights) ** 2))``
  once per MCMC proposal (reference mod16/calibration.py:224-227, 907-909; EXTERNAL module,
  parity unpinned -- SURVEY.md 8(c)).

``objective_batch`` evaluates all draws in one call and returns, per draw,
``sse = sum(((pred - obs) * w) ** 2)`` and ``count`` over the pairs where obs, w and pred
are not NaN; ``rmsd`` / ``nash_sutcliffe`` form the objectives from them.
``ObjectiveBackend`` keeps the drivers resident on the device between calls (the MCMC use).
'''
import ctypes
import numpy as np
import torch

from . import _abi
from .forward import (_Staged, _is_scalar, _shape, _view_2d, hand_over, launch_stream,
                      pick_device, result_dtype)


def _draw_matrix(draws, dev):
    if isinstance(draws, torch.Tensor):
        d = draws.detach().to(device=dev, dtype=torch.float64)
    else:
        d = torch.from_numpy(np.ascontiguousarray(np.asarray(draws, np.float64))).to(dev)
    if d.ndim == 1:
        d = d.reshape(1, -1)
    if d.ndim != 2 or d.shape[1] != _abi.N_PARAMS:
        raise ValueError('draws must be (D, 11) in MOD16.required_parameters order')
    return d.contiguous()


class ObjectiveBackend(object):
    '''
    Drivers, observations and weights staged once on the device; ``__call__(draws)`` returns
    ``(sse, count)`` for a (D, 11) matrix of parameter draws.

    Parameters
    ----------
    drivers : sequence of the 14 positional drivers of ``MOD16._et`` (scalars, (N,), (T, N);
        numpy or CUDA tensors)
    observed : array broadcastable to the drivers' shape; NaN marks a missing observation
    weights : None (unweighted), scalar or array like ``observed``
    semantics : 'static' (``_et``, W m-2) or 'instance' (day + night of
        ``evapotranspiration()``, kg m-2 s-1)
    gate : 'global' -- each draw gets the batch-global transpiration gate of ITS call, as in
        the reference (:343) --, 'open' or 'closed'.  The gate of a draw is taken over EVERY
        element handed in, with or without an observation or weight: the reference evaluates
        ``np.any(g_surf > 0)`` inside ``_et`` on whatever drivers it is given and masks the
        missing observations afterwards, in the score (calibration.py:224-227).  Its default
        calibration run hands ``_et`` the drivers with NaN observations in place
        (``_load_data``, calibration.py:355-370: validation days are set to NaN, nothing is
        removed); the k-fold branch (calibration.py:895-901) and the sensitivity loader
        (sensitivity.py:179-186) remove those elements first, in which case there is nothing
        to tell apart.  Elements without a valid pair cost nothing on the fast path: only a
        draw that no valid element opens (csl == 0, a Tmin ramp above every Tmin) is
        re-evaluated over all of them.
    dtype : np.float32 (fast path) or np.float64; default follows numpy promotion
    '''

    def __init__(self, drivers, observed, weights=None, semantics='static', gate='global',
                 dtype=None, device=None, force=0):
        self.lib = _abi.load()
        drivers = list(drivers)
        if len(drivers) != _abi.N_DRIVERS:
            raise TypeError('expected 14 drivers, got %d' % len(drivers))
        if gate not in ('global', 'open', 'closed'):
            raise ValueError('gate must be global, open or closed')
        everything = drivers + [observed] + ([weights] if weights is not None else [])
        self.np_dtype = result_dtype(drivers, dtype)
        self.device = pick_device(everything, device)
        shapes = [_shape(v) for v in everything]
        self.full_shape = tuple(np.broadcast_shapes(*shapes))
        self.T, self.N = _view_2d(self.full_shape)
        self.flags = ({'static': _abi.SEM_STATIC, 'instance': _abi.SEM_INSTANCE}[semantics] |
                      _abi.GATE_FLAGS[gate] | force)
        with torch.cuda.device(self.device):
            st = _Staged(self.device, self.np_dtype)
            st.full_shape = self.full_shape
            self.cd = _abi.Drivers()
            for k, v in enumerate(drivers):
                st.fill(self.cd.drv[k], v, self.T, self.N, len(self.full_shape))
            self.co = _abi.Objective()
            st.fill(self.co.observed, observed, self.T, self.N, len(self.full_shape))
            st.fill(self.co.weights, 1.0 if weights is None else weights, self.T, self.N,
                    len(self.full_shape))
            self._staged = st  # keeps the device copies alive
        self._ws = None
        self.semantics, self.gate = semantics, gate

    def __call__(self, draws, stream=None, return_gate=False, as_torch=False, raw=False):
        '''
        Returns ``(sse, count)`` (+ ``gate`` with ``return_gate``) for the (D, 11) ``draws``.
        ``raw=True`` returns instead the (D, 5) CUDA tensor of un-gated statistics
        [sse_open, n_open, sse_closed, n_closed, any(g_surf > 0)] of this backend's elements
        (see ``finish_raw`` and ``distributed.objective_site_sharded``).
        '''
        dev = self.device
        with torch.cuda.device(dev), launch_stream(dev, stream) as (s, caller_stream):
            d = _draw_matrix(draws, dev)
            D = d.shape[0]
            is_f32 = 1 if self.np_dtype == np.dtype('float32') else 0
            need = self.lib.mod16_objective_workspace_bytes(self.T, self.N, D, is_f32)
            if self._ws is None or self._ws.numel() * 8 < need:
                self._ws = torch.empty((need + 7) // 8 + 32, device=dev, dtype=torch.int64)
            ws = self._ws
            off = (-ws.data_ptr()) % 256
            sse = torch.empty(D, device=dev, dtype=torch.float64)
            cnt = torch.empty(D, device=dev, dtype=torch.float64)
            gate = torch.empty(D, device=dev, dtype=torch.int32) if return_gate else None
            raw_t = torch.empty((D, 5), device=dev, dtype=torch.float64) if raw else None
            co = self.co
            co.raw = raw_t.data_ptr() if raw else None
            co.draws = d.data_ptr()
            co.sse = sse.data_ptr()
            co.count = cnt.data_ptr()
            co.gate = gate.data_ptr() if gate is not None else None
            co.workspace = ws.data_ptr() + off
            co.workspace_bytes = ws.numel() * 8 - off
            fn = self.lib.mod16_objective_batch_f32 if is_f32 else \
                self.lib.mod16_objective_batch_f64
            _abi.check(fn(ctypes.byref(self.cd), ctypes.byref(co), self.T, self.N, D,
                          self.flags, ctypes.c_void_p(s.cuda_stream)))
            hand_over((sse, cnt, gate, raw_t), caller_stream)
            if raw:
                return raw_t
            if as_torch:
                return (sse, cnt, gate) if return_gate else (sse, cnt)
            out = (sse.cpu().numpy(), cnt.cpu().numpy())
            return out + (gate.cpu().numpy().astype(bool),) if return_gate else out

    # objectives of the reference's callers, formed from the sufficient statistics
    def neg_rmsd(self, draws, **kw):
        'What mod17 BlackBoxLikelihood(objective="rmsd") returns, for every draw.'
        sse, cnt = self(draws, **kw)[:2]
        return -rmsd(sse, cnt)


def finish_raw(raw, semantics='static', gate='global'):
    '''
    Applies each draw's transpiration gate to (summed) un-gated statistics ``raw`` (D, 5):
    returns (sse, count, gate_open) tensors.  The same selection the reduce kernel makes.
    '''
    any_pos = raw[:, 4] > 0
    if semantics != 'static' or gate == 'open':
        use_open = torch.ones_like(any_pos)
    elif gate == 'closed':
        use_open = torch.zeros_like(any_pos)
    else:
        use_open = any_pos
    sse = torch.where(use_open, raw[:, 0], raw[:, 2])
    cnt = torch.where(use_open, raw[:, 1], raw[:, 3])
    return sse, cnt, any_pos


def objective_batch(draws, drivers, observed, weights=None, **kw):
    'One-shot form of ``ObjectiveBackend``: returns (sse, count) per draw.'
    call_kw = dict((k, kw.pop(k)) for k in ('stream', 'return_gate', 'as_torch', 'raw')
                   if k in kw)
    return ObjectiveBackend(drivers, observed, weights, **kw)(draws, **call_kw)


def rmsd(sse, count):
    '''
    sqrt(nanmean(((pred - obs) w)^2)) from the statistics; NaN where no pair was valid.
    mod17's ``BlackBoxLikelihood.loglik`` (objective "rmsd", the reference's configured
    ``RMSD``, data/MOD16_calibration_config.yaml:49) returns its negative:
    ``-np.sqrt(np.nanmean(((predicted - observed) * weights) ** 2))``.  Known answers:
    tests/test_objective.py::test_objective_definitions_known_answers.
    '''
    with np.errstate(invalid='ignore', divide='ignore'):
        return np.sqrt(np.asarray(sse) / np.asarray(count))


def nash_sutcliffe(sse, observed, norm=False):
    '''
    Nash-Sutcliffe efficiency from the unweighted statistics: 1 - sse / sum((obs - mean(obs))^2)
    over the non-NaN observations; ``norm`` gives 1 / (2 - NSE) as sensitivity.py:96 asks.
    Restates ``mod17.science.nash_sutcliffe(predicted, observed, norm)``:
    ``nse = 1 - nansum((observed - predicted) ** 2) / nansum((observed - nanmean(observed)) ** 2)``
    (the numerator skips a pair when either value is NaN -- that is ``sse`` --, the
    denominator runs over every observation present); ``norm`` returns ``1 / (2 - nse)``.
    Known answers: tests/test_objective.py::test_objective_definitions_known_answers.
    '''
    obs = np.asarray(observed, np.float64)
    obs = obs[~np.isnan(obs)]
    nse = 1.0 - np.asarray(sse) / np.sum((obs - obs.mean()) ** 2)
    return 1.0 / (2.0 - nse) if norm else nse
