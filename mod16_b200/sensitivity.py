'''
Sobol' sensitivity analysis with the model evaluations on the GPU -- the B200 counterpart of
the reference's mod16/sensitivity.py, whose two loops

    param_sweep = sobol_sample(problem, 512)            # SALib.sample.sobol.sample  (:92)
    for i, X in enumerate(param_sweep):
        Y[i] = nash_sutcliffe(MOD16._et(X, *drivers), tower_obs, norm = True)       (:94-96)
    metrics = sobol.analyze(problem, Y)                 # SALib.analyze.sobol        (:116)

and (`analysis="drivers"`, :98-114) `Y[i] = MOD16._et(params_vector, *X)` call the forward
model once per row of a Saltelli sample.  Here the whole sample is ONE objective batch (K2,
`mod16_objective_batch_*`) or ONE forward call (K1), and both the Saltelli expansion of the
base Sobol' sequence and the index estimators run on the device (torch as plumbing: they are
O(rows x D) work next to O(rows x elements) model evaluations).

SALib is an external package (not under /root/reference, not installed): the sample layout
[A, AB_1 .. AB_D, BA_1 .. BA_D, B] per base point, the output normalisation, the estimators
(Saltelli et al. 2010: first order `mean(B (AB_j - A)) / var`, total order
`mean((A - AB_j)^2) / 2 var`, second order `mean(BA_j AB_k - A B) / var - S_j - S_k`) and the
bootstrap confidence half-widths (`Z * std(ddof = 1)` over `num_resamples` resamples of the
base points) follow its published algorithm; oracle/salib_oracle.py is the numpy restatement
the tests compare with.  The base sequence comes from `scipy.stats.qmc.Sobol` in 2 D
dimensions, which is what SALib itself draws from.
'''
from collections import OrderedDict

import numpy as np
import torch

from .forward import pick_device, run_forward
from .objective import ObjectiveBackend
from .synthetic import BOUNDS, PARAM_BOUNDS   # sensitivity.py:31-59

__all__ = ['BOUNDS', 'PARAM_BOUNDS', 'saltelli_sample', 'sobol_indices',
           'parameter_sensitivity', 'driver_sensitivity']


def saltelli_sample(bounds, N, calc_second_order=True, scramble=True, skip_values=0, seed=None,
                    device=None):
    '''
    The sample of `SALib.sample.sobol.sample(problem, N)`: ``N * (2 D + 2)`` rows (``D + 2``
    without second order) of D parameters as a float64 tensor on ``device``.

    bounds : sequence of D (lower, upper) pairs -- ``problem['bounds']``
    N      : base points; a power of two keeps the Sobol' sequence balanced (sensitivity.py:91)
    '''
    from scipy.stats import qmc
    b = np.asarray(list(bounds.values()) if isinstance(bounds, dict) else bounds, np.float64)
    D = b.shape[0]
    dev = torch.device(device) if device is not None else pick_device([])
    qrng = qmc.Sobol(d=2 * D, scramble=scramble, seed=seed)
    if skip_values > 0:
        qrng.fast_forward(int(skip_values))
    base = torch.from_numpy(qrng.random(N)).to(dev)              # (N, 2 D)
    a_part, b_part = base[:, :D], base[:, D:]
    eye = torch.eye(D, dtype=torch.bool, device=dev)
    blocks = [a_part[:, None, :]]
    # AB_k: A with column k from B;  BA_k: B with column k from A
    blocks.append(torch.where(eye[None], b_part[:, None, :], a_part[:, None, :]))
    if calc_second_order:
        blocks.append(torch.where(eye[None], a_part[:, None, :], b_part[:, None, :]))
    blocks.append(b_part[:, None, :])
    rows = torch.cat(blocks, dim=1).reshape(-1, D)
    lo = torch.from_numpy(b[:, 0]).to(dev)
    hi = torch.from_numpy(b[:, 1]).to(dev)
    return rows * (hi - lo) + lo


def sobol_indices(Y, D, calc_second_order=True, num_resamples=100, conf_level=0.95, seed=None,
                  resample_index=None):
    '''
    `SALib.analyze.sobol.analyze(problem, Y)` on the device.  ``Y`` holds the model output of
    the rows of ``saltelli_sample`` in order (numpy array or tensor); returns an OrderedDict of
    numpy arrays S1, S1_conf, ST, ST_conf (D,) and, with second order, S2, S2_conf (D, D; NaN
    on and below the diagonal, as SALib leaves them).
    '''
    from scipy.stats import norm
    y = Y.detach() if isinstance(Y, torch.Tensor) else torch.as_tensor(np.asarray(Y, np.float64))
    y = y.to(torch.float64).reshape(-1)
    dev = y.device
    step = 2 * D + 2 if calc_second_order else D + 2
    if y.numel() % step:
        raise RuntimeError('incorrect number of samples in the model output')
    N = y.numel() // step
    y = (y - y.mean()) / y.std(unbiased=False)
    y = y.reshape(N, step)
    A, B = y[:, 0], y[:, step - 1]
    AB = y[:, 1:D + 1]
    BA = y[:, D + 1:2 * D + 1] if calc_second_order else None
    if resample_index is None:
        g = torch.Generator(device=dev)
        if seed is not None:
            g.manual_seed(int(seed))
        r = torch.randint(0, N, (N, num_resamples), generator=g, device=dev)
    else:
        r = torch.as_tensor(np.asarray(resample_index)).to(dev)
    Z = float(norm.ppf(0.5 + conf_level / 2))

    def var_ab(a, b):                  # np.var(np.r_[A, B], axis=0)
        return torch.cat([a, b], dim=0).var(dim=0, unbiased=False)

    def first(a, ab, b):               # a, b: (N, ...) ; ab: (N, ..., D)
        return (b[..., None] * (ab - a[..., None])).mean(dim=0) / var_ab(a, b)[..., None]

    def total(a, ab, b):
        return 0.5 * ((a[..., None] - ab) ** 2).mean(dim=0) / var_ab(a, b)[..., None]

    Ar, Br, ABr = A[r], B[r], AB[r]    # (N, R), (N, R), (N, R, D)
    out = OrderedDict()
    out['S1'] = first(A, AB, B)
    out['S1_conf'] = Z * first(Ar, ABr, Br).std(dim=0, unbiased=True)
    out['ST'] = total(A, AB, B)
    out['ST_conf'] = Z * total(Ar, ABr, Br).std(dim=0, unbiased=True)
    if calc_second_order:
        BAr = BA[r]

        def second(a, ab, ba, b):      # -> (..., D, D): [j, k]
            v = var_ab(a, b)
            vjk = (ba[..., :, None] * ab[..., None, :]).mean(dim=0) - (a * b).mean(dim=0)[..., None, None]
            s = first(a, ab, b)
            return vjk / v[..., None, None] - s[..., :, None] - s[..., None, :]
        upper = torch.triu(torch.ones(D, D, dtype=torch.bool, device=dev), diagonal=1)
        nan = torch.full((D, D), float('nan'), dtype=torch.float64, device=dev)
        out['S2'] = torch.where(upper, second(A, AB, BA, B), nan)
        out['S2_conf'] = torch.where(upper, Z * second(Ar, ABr, BAr, Br).std(dim=0, unbiased=True),
                                     nan)
    return OrderedDict((k, v.cpu().numpy()) for k, v in out.items())


def parameter_sensitivity(drivers, tower_obs, N=512, bounds=None, calc_second_order=True,
                          seed=None, num_resamples=100, dtype=np.float32, device=None,
                          batch=65536, return_output=False, **sample_kw):
    '''
    `sensitivity.main(analysis="parameters")` (:82-96, :116): the Nash-Sutcliffe efficiency
    (normalised, `1 / (2 - NSE)`) of `MOD16._et(X, *drivers)` against ``tower_obs`` for every
    row X of the Saltelli sample of the eleven parameters, then the Sobol' indices of that
    output.  All rows go through the objective batch K2, ``batch`` rows per call.
    '''
    bounds = bounds if bounds is not None else PARAM_BOUNDS
    D = len(bounds)
    backend = ObjectiveBackend(drivers, tower_obs, None, dtype=dtype, device=device)
    X = saltelli_sample(bounds, N, calc_second_order, seed=seed, device=backend.device,
                        **sample_kw)
    sse = torch.cat([backend(X[i:i + batch], as_torch=True)[0]
                     for i in range(0, X.shape[0], batch)])
    obs = tower_obs.detach().to(backend.device, torch.float64) if isinstance(
        tower_obs, torch.Tensor) else torch.as_tensor(
            np.asarray(tower_obs, np.float64)).to(backend.device)
    obs = obs[~torch.isnan(obs)]
    nse = 1.0 - sse / ((obs - obs.mean()) ** 2).sum()       # mod17.science.nash_sutcliffe
    Y = 1.0 / (2.0 - nse)                                   # norm = True (:96)
    S = sobol_indices(Y, D, calc_second_order, num_resamples, seed=seed)
    return (S, X, Y) if return_output else S


def driver_sensitivity(params_vector, N=2048, bounds=None, calc_second_order=True, seed=None,
                       num_resamples=100, dtype=np.float32, device=None, return_output=False,
                       **sample_kw):
    '''
    `sensitivity.main(analysis="drivers")` (:98-114, :116): the output is `MOD16._et` itself for
    a fixed parameter vector, one row of the Saltelli sample of the fourteen drivers per
    evaluation -- here ONE forward call over all rows.
    '''
    bounds = bounds if bounds is not None else BOUNDS
    D = len(bounds)
    X = saltelli_sample(bounds, N, calc_second_order, seed=seed, device=device, **sample_kw)
    t_dtype = torch.float32 if np.dtype(dtype) == np.float32 else torch.float64
    cols = [X[:, j].to(t_dtype).contiguous() for j in range(D)]
    Y = run_forward(cols, params=[float(v) for v in params_vector], gate='element',
                    dtype=np.dtype(dtype), device=X.device, as_torch=True)[0]
    S = sobol_indices(Y, D, calc_second_order, num_resamples, seed=seed)
    return (S, X, Y) if return_output else S
