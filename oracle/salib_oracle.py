'''
TEST INFRASTRUCTURE ONLY -- CPU restatement of the two SALib functions the reference's
sensitivity analysis calls (mod16/sensitivity.py:24-25, :92, :110, :116):

    from SALib.sample.sobol import sample as sobol_sample
    from SALib.analyze import sobol            # sobol.analyze(problem, Y)

SALib is an EXTERNAL dependency that is absent from /root/reference and from this image (the
reference does not even pin it in pyproject.toml; `SALib.sample.sobol` exists from SALib
1.4.6 on).  What follows restates its published algorithm -- Saltelli's extension of a Sobol'
sequence (Saltelli 2002; Campolongo et al. 2011) and the estimators of Saltelli et al. 2010,
Table 2 (first order: eq. (b), total order: eq. (f)) exactly as SALib's source arranges them
(loop order, output normalisation, `ddof`, bootstrap) -- in plain numpy loops.
PARITY UNPINNED against SALib itself (it cannot be run here); anchored instead on the analytic
Sobol' indices of the Ishigami function (tests/test_sensitivity.py) and on the structure of
the sample matrix the analysis relies on.
'''
import numpy as np
from scipy.stats import norm, qmc


def sample(problem, N, calc_second_order=True, scramble=True, skip_values=0, seed=None):
    'SALib.sample.sobol.sample: (N * (2 D + 2), D) rows [A, AB_1..AB_D, BA_1..BA_D, B] per base point.'
    D = problem['num_vars']
    qrng = qmc.Sobol(d=2 * D, scramble=scramble, seed=seed)
    if skip_values > 0:
        qrng.fast_forward(skip_values)
    base = qrng.random(N)
    rows = (2 * D + 2) if calc_second_order else (D + 2)
    out = np.zeros((rows * N, D))
    index = 0
    for i in range(N):
        for j in range(D):                      # A
            out[index, j] = base[i, j]
        index += 1
        for k in range(D):                      # AB_k: A with column k from B
            for j in range(D):
                out[index, j] = base[i, j + D] if j == k else base[i, j]
            index += 1
        if calc_second_order:
            for k in range(D):                  # BA_k: B with column k from A
                for j in range(D):
                    out[index, j] = base[i, j] if j == k else base[i, j + D]
                index += 1
        for j in range(D):                      # B
            out[index, j] = base[i, j + D]
        index += 1
    bounds = np.asarray(problem['bounds'], np.float64)
    return out * (bounds[:, 1] - bounds[:, 0]) + bounds[:, 0]


def separate_output_values(Y, D, N, calc_second_order):
    AB = np.zeros((N, D))
    BA = np.zeros((N, D)) if calc_second_order else None
    step = 2 * D + 2 if calc_second_order else D + 2
    A = Y[0:Y.size:step]
    B = Y[(step - 1):Y.size:step]
    for j in range(D):
        AB[:, j] = Y[(j + 1):Y.size:step]
        if calc_second_order:
            BA[:, j] = Y[(j + 1 + D):Y.size:step]
    return A, B, AB, BA


def first_order(A, AB, B):
    return np.mean(B * (AB - A), axis=0) / np.var(np.r_[A, B], axis=0)


def total_order(A, AB, B):
    return 0.5 * np.mean((A - AB) ** 2, axis=0) / np.var(np.r_[A, B], axis=0)


def second_order(A, ABj, ABk, BAj, B):
    Vjk = np.mean(BAj * ABk - A * B, axis=0) / np.var(np.r_[A, B], axis=0)
    return Vjk - first_order(A, ABj, B) - first_order(A, ABk, B)


def analyze(problem, Y, calc_second_order=True, num_resamples=100, conf_level=0.95, seed=None,
            resample_index=None):
    '''
    SALib.analyze.sobol.analyze.  `resample_index` (N, num_resamples) replaces the random
    bootstrap draw, so that an implementation with another generator can be compared.
    '''
    D = problem['num_vars']
    Y = np.asarray(Y, np.float64)
    step = 2 * D + 2 if calc_second_order else D + 2
    assert Y.size % step == 0, 'incorrect number of samples in the model output'
    N = Y.size // step
    Y = (Y - Y.mean()) / Y.std()
    A, B, AB, BA = separate_output_values(Y, D, N, calc_second_order)
    if resample_index is None:
        resample_index = np.random.default_rng(seed).integers(N, size=(N, num_resamples))
    r = np.asarray(resample_index)
    Z = norm.ppf(0.5 + conf_level / 2)
    S = {'S1': np.zeros(D), 'S1_conf': np.zeros(D), 'ST': np.zeros(D), 'ST_conf': np.zeros(D)}
    for j in range(D):
        S['S1'][j] = first_order(A, AB[:, j], B)
        S['S1_conf'][j] = Z * first_order(A[r], AB[r, j], B[r]).std(ddof=1)
        S['ST'][j] = total_order(A, AB[:, j], B)
        S['ST_conf'][j] = Z * total_order(A[r], AB[r, j], B[r]).std(ddof=1)
    if calc_second_order:
        S['S2'] = np.full((D, D), np.nan)
        S['S2_conf'] = np.full((D, D), np.nan)
        for j in range(D):
            for k in range(j + 1, D):
                S['S2'][j, k] = second_order(A, AB[:, j], AB[:, k], BA[:, j], B)
                S['S2_conf'][j, k] = Z * second_order(
                    A[r], AB[r, j], AB[r, k], BA[r, j], B[r]).std(ddof=1)
    return S


def ishigami(x, a=7.0, b=0.1):
    return np.sin(x[:, 0]) + a * np.sin(x[:, 1]) ** 2 + b * x[:, 2] ** 4 * np.sin(x[:, 0])


def ishigami_indices(a=7.0, b=0.1):
    'Analytic first- and total-order indices of the Ishigami function on [-pi, pi]^3.'
    pi = np.pi
    V = a ** 2 / 8 + b * pi ** 4 / 5 + b ** 2 * pi ** 8 / 18 + 0.5
    V1 = 0.5 * (1 + b * pi ** 4 / 5) ** 2
    V2 = a ** 2 / 8
    V13 = b ** 2 * pi ** 8 * (1.0 / 18 - 1.0 / 50)
    return np.array([V1, V2, 0.0]) / V, np.array([V1 + V13, V2, V13]) / V
