'''
The device-resident DE-MC population sampler (mod16_b200/sampler.py).  The reference's
sampler is EXTERNAL (mod17 + pymc, absent), so there are no draws to compare with:

CPU: the prior densities (YAML layout of mod16/data/MOD16_BPLUT_prior_*.yaml) normalise to 1
     and their samplers match them; the transition kernel, driven by a stub backend with a
     known Gaussian likelihood, reproduces that posterior's mean and spread;
GPU: a synthetic calibration (observations generated from known parameters) moves the
     population from the prior towards fits as good as the generating parameters'.
'''
import numpy as np
import pytest
import torch

from mod16_b200 import MOD16, synthetic, utils
from mod16_b200.sampler import Prior, PopulationSampler

SPEC = {
    'vpd_close': {'lower': [None, 1000.] * 6 + [1000.], 'upper': [None, 8000.] * 6 + [8000.]},
    'gl_sh': {'mu': -3.91, 'sigma': 0.1},
    'gl_wv': {'mu': -3.91, 'sigma': 0.1},
    'g_cuticular': {'mu': -10.42, 'sigma': 1.04},
    'csl': {'mu': -5.5, 'sigma': 0.8},
    'rbl_min': {'lower': 10., 'upper': 1000., 'c': 10.},
    'rbl_max': {'lower': 100., 'upper': 1000., 'c': 1000.},
    'beta': {'lower': 0., 'upper': 1000.},
}


def test_prior_densities_and_samplers():
    p = Prior(SPEC, pft=7)
    assert p.names == ['vpd_close', 'gl_sh', 'gl_wv', 'g_cuticular', 'csl', 'rbl_min',
                       'rbl_max', 'beta']
    g = torch.Generator().manual_seed(1)
    s = p.sample(200000, g)
    # moments of the samplers
    assert abs(s[:, 0].mean().item() - 4500) < 30                       # Uniform(1000, 8000)
    assert abs(s[:, 1].mean().item() / np.exp(-3.91 + 0.005) - 1) < 5e-3  # LogNormal
    assert abs(s[:, 5].mean().item() - (10 + 1000 + 10) / 3) < 3        # Triangular, mode at lower
    assert abs(s[:, 6].mean().item() - (100 + 1000 + 1000) / 3) < 3     # mode at upper
    assert torch.isfinite(p.log_prob(s)).all()
    # each marginal density integrates to 1
    for j, (lo, hi) in enumerate([(1000, 8000), (1e-4, 0.1), (1e-4, 0.1), (1e-8, 2e-2),
                                  (1e-5, 0.5), (10, 1000), (100, 1000), (0, 1000)]):
        x = s[:1].repeat(400001, 1)
        grid = torch.linspace(lo, hi, 400001, dtype=torch.float64)
        x[:, j] = grid
        ref = p.log_prob(s[:1])
        # density of coordinate j up to the (constant) other factors
        dens = torch.exp(p.log_prob(x) - ref)
        one = Prior({p.names[j]: SPEC[p.names[j]]}, pft=7)
        dj = torch.exp(one.log_prob(grid[:, None]))
        assert abs(torch.trapezoid(dj, grid).item() - 1) < 2e-2, p.names[j]
        assert torch.allclose(dens / dens.max(), dj / dj.max(), atol=1e-9)
    # outside the support
    bad = s[:1].clone()
    bad[0, 0] = 999.0
    assert p.log_prob(bad).item() == -np.inf


def test_transition_kernel_samples_a_known_posterior():
    'Stub backend: Gaussian log-likelihood in (vpd_close, beta); wide uniform priors.'
    spec = {'vpd_close': {'lower': 1000., 'upper': 8000.}, 'beta': {'lower': 0., 'upper': 1000.}}
    prior = Prior(spec)
    mean = torch.tensor([4000., 300.], dtype=torch.float64)
    sd = torch.tensor([200., 40.], dtype=torch.float64)
    cols = [MOD16.required_parameters.index(k) for k in ('vpd_close', 'beta')]

    def backend(draws):
        z = (draws[:, cols] - mean) / sd
        return -0.5 * (z ** 2).sum(dim=1)
    base = [-8., 8.8, 650., 4400., .02, .02, 1e-5, .0055, 60., 95., 250.]
    smp = PopulationSampler(backend, base, prior, chains=64, scaling=1e-3, seed=3)
    out = smp.run(draws=400, tune=300, thin=2)
    tr = out['trace'][:, :, cols].reshape(-1, 2)
    assert out['free'] == ['vpd_close', 'beta'] and 0.1 < out['accept_rate'] < 0.7
    assert torch.allclose(tr.mean(dim=0), mean, rtol=0.02)
    assert torch.allclose(tr.std(dim=0), sd, rtol=0.15)
    # fixed parameters never move
    other = [i for i in range(11) if i not in cols]
    assert torch.equal(out['trace'][:, :, other],
                       torch.tensor(base, dtype=torch.float64)[other].expand(400, 64, 9))
    with pytest.raises(ValueError):
        PopulationSampler(backend, base, prior, chains=4)


@pytest.mark.gpu
def test_synthetic_calibration_recovers_the_generating_fit(cuda_lib):
    from mod16_b200.objective import ObjectiveBackend
    T, N = 365, 24
    d = synthetic.host_drivers(T, N, 2024, dtype=np.float32)
    drv = [d[k] for k in synthetic.DRIVER_NAMES]
    table = utils.bplut_table(utils.restore_bplut(utils.BPLUT_C51), MOD16.required_parameters,
                              default_beta=250.0)
    truth = np.array(table[:, 7])
    pred = MOD16._et(list(truth), *drv, dtype=np.float32)
    rng = np.random.default_rng(0)
    obs = (pred + rng.normal(0, 8, pred.shape)).astype(np.float32)
    backend = ObjectiveBackend(drv, obs, None, dtype=np.float32)
    rmsd_truth = float(-backend.neg_rmsd(truth[None])[0])
    smp = PopulationSampler.from_backend(backend, truth, Prior(SPEC, pft=7), chains=96,
                                         scaling=1e-3, seed=11, fixed={'beta': 250})
    assert smp.free == ['vpd_close', 'gl_sh', 'gl_wv', 'g_cuticular', 'csl', 'rbl_min', 'rbl_max']
    start = float(-smp.logp[torch.isfinite(smp.logp)].mean())
    out = smp.run(draws=100, tune=400)
    final_rmsd = -backend.neg_rmsd(out['trace'][-1].cpu().numpy())
    assert out['trace'].is_cuda and out['trace'].shape == (100, 96, 11)
    # The target is prior x exp(-RMSD) (a Potential of -RMSD, as the reference builds it): a
    # broad distribution, not an optimiser -- the population moves towards good fits, its
    # best members fit as well as the generating parameters, and it stays a spread.
    assert np.min(final_rmsd) < rmsd_truth * 1.05 + 0.3
    assert float(np.median(final_rmsd)) < 0.75 * start
    assert np.std(final_rmsd) > 0.05
    assert np.all(out['trace'][:, :, 10].cpu().numpy() == 250.0)
    assert np.all(out['trace'][:, :, 0].cpu().numpy() == truth[0])
