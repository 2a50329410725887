'''
Population MCMC on the device, driving the objective-batch kernel (K2).

The reference calibrates one PFT with pymc's ``DEMetropolisZ`` through mod17's
``StochasticSampler`` (EXTERNAL; reference mod16/calibration.py:152-267, 956-958): 3 chains,
each proposal one ``MOD16._et`` call on the host.  A batch backend only pays off when many
proposals are evaluated at once, so this module runs the population form of the same
algorithm -- Differential Evolution Markov Chain (ter Braak 2006; the reference's docstring
cites ter Braak 2008) -- with every chain's proposal of an iteration evaluated in ONE
``mod16_objective_batch`` call:

    x*_i = x_i + gamma (x_a - x_b) + e,   a != b != i drawn from the current population,
    gamma = 2.38 / sqrt(2 d)  (1 every 10th iteration),   e ~ N(0, (scaling * width)^2)
    accept with probability min(1, p(x*_i) / p(x_i))

The target is what the reference builds in ``compile_et_model`` (:224-262): the prior of each
free parameter (Uniform / LogNormal / Triangular, the YAML layout of
mod16/data/MOD16_BPLUT_prior_*.yaml) times ``exp(objective)`` with the objective -RMSD
(``optimization: objective: RMSD`` of mod16/data/MOD16_calibration_config.yaml:49, added as
a pymc Potential).  Parameters shared with MOD17 (tmin_close, tmin_open, vpd_open) and any
``fixed`` ones stay at their BPLUT values (:231-235, :246-259).

Everything -- proposals, priors, accept/reject, the trace -- lives in torch tensors on the
model's device; ``torch`` is plumbing here, the likelihoods come from the CUDA kernel.  There
is no reference implementation of the sampler in /root/reference to compare draws with: the
tests check the transition kernel on a target with a known answer (a stub backend) and that
a synthetic calibration recovers the generating parameters.
'''
import math

import numpy as np
import torch

from .model import MOD16

SHARED_WITH_MOD17 = ('tmin_close', 'tmin_open', 'vpd_open')


class Prior(object):
    '''
    Independent priors of the free parameters.  ``spec`` maps a parameter name to a dict in
    the reference's YAML layout: ``{lower, upper}`` Uniform, ``{mu, sigma}`` LogNormal,
    ``{lower, upper, c}`` Triangular; values may be 13-element per-PFT lists (``pft`` picks).
    '''

    def __init__(self, spec, pft=None, device='cpu'):
        self.names = [k for k in MOD16.required_parameters if k in spec]
        self.kind, rows = [], []
        for k in self.names:
            d = dict((a, (b[pft] if isinstance(b, (list, tuple)) else b)) for a, b in spec[k].items())
            if 'mu' in d:
                self.kind.append(1)
                rows.append([d['mu'], d['sigma'], 0.0])
            elif 'c' in d:
                self.kind.append(2)
                rows.append([d['lower'], d['upper'], d['c']])
            else:
                self.kind.append(0)
                rows.append([d['lower'], d['upper'], 0.0])
        self.par = torch.tensor(rows, dtype=torch.float64, device=device).reshape(-1, 3)
        self.kind_t = torch.tensor(self.kind, device=device)

    def to(self, device):
        self.par, self.kind_t = self.par.to(device), self.kind_t.to(device)
        return self

    def log_prob(self, x):
        'x: (C, d) values of the free parameters; returns (C,) log density (-inf outside).'
        a, b, c = self.par[:, 0], self.par[:, 1], self.par[:, 2]
        ninf = torch.full_like(x, -math.inf)
        uni = torch.where((x >= a) & (x <= b), -torch.log(b - a).expand_as(x), ninf)
        xs = torch.clamp(x, min=1e-300)
        logn = torch.where(x > 0, -torch.log(xs * b * math.sqrt(2 * math.pi)) -
                           (torch.log(xs) - a) ** 2 / (2 * b ** 2), ninf)
        up = 2 * (x - a) / ((b - a) * torch.clamp(c - a, min=1e-300))
        dn = 2 * (b - x) / ((b - a) * torch.clamp(b - c, min=1e-300))
        tri = torch.where((x >= a) & (x <= b), torch.where(x < c, up, dn), torch.zeros_like(x))
        tri = torch.where(x == c, (2 / (b - a)).expand_as(x), tri)
        tri = torch.where(tri > 0, torch.log(torch.clamp(tri, min=1e-300)), ninf)
        k = self.kind_t
        return torch.where(k == 1, logn, torch.where(k == 2, tri, uni)).sum(dim=1)

    def sample(self, n, generator=None):
        '(n, d) draws from the prior (initial population).'
        a, b, c = self.par[:, 0], self.par[:, 1], self.par[:, 2]
        u = torch.rand((n, len(self.names)), dtype=torch.float64, device=self.par.device,
                       generator=generator)
        z = torch.randn((n, len(self.names)), dtype=torch.float64, device=self.par.device,
                        generator=generator)
        uni = a + (b - a) * u
        logn = torch.exp(a + b * z)
        fc = (c - a) / (b - a)
        tri = torch.where(u < fc, a + torch.sqrt(u * (b - a) * (c - a)),
                          b - torch.sqrt((1 - u) * (b - a) * (b - c)))
        k = self.kind_t
        return torch.where(k == 1, logn, torch.where(k == 2, tri, uni))

    def width(self):
        'A scale per parameter for the jitter term: prior range, or the LogNormal sd.'
        a, b = self.par[:, 0], self.par[:, 1]
        sd = torch.sqrt((torch.exp(b ** 2) - 1) * torch.exp(2 * a + b ** 2))
        return torch.where(self.kind_t == 1, sd, b - a)


class PopulationSampler(object):
    '''
    DE-MC over the free parameters of one PFT.

    backend : callable ``neg_rmsd(draws (C, 11) CUDA/CPU float64 tensor) -> (C,) tensor`` --
        normally ``lambda d: -torch.sqrt(sse / cnt)`` of an ``ObjectiveBackend`` (see
        ``from_backend``); any log-likelihood works.
    params : the 11 BPLUT values of the PFT (fixed parameters are taken from here)
    prior : ``Prior`` over the free parameters
    chains : population size (>= 2 d + 1; the kernel's fast path needs >= 32)
    '''

    def __init__(self, backend, params, prior, chains=128, scaling=1e-3, seed=0, fixed=None):
        self.backend = backend
        self.prior = prior
        dev = prior.par.device
        self.device = dev
        self.base = torch.as_tensor(np.asarray(params, np.float64), device=dev)
        names = list(MOD16.required_parameters)
        fixed = dict(fixed or {})
        self.free = [k for k in prior.names if k not in SHARED_WITH_MOD17 and
                     fixed.get(k) is None]
        for k, v in fixed.items():
            if v is not None:
                self.base[names.index(k)] = float(v)
        self.cols = torch.tensor([names.index(k) for k in self.free], device=dev)
        self.pidx = torch.tensor([prior.names.index(k) for k in self.free], device=dev)
        self.chains = int(chains)
        self.d = len(self.free)
        if self.chains < 2 * self.d + 1:
            raise ValueError('DE-MC needs at least 2 d + 1 = %d chains' % (2 * self.d + 1))
        self.scaling = float(scaling)
        self.gen = torch.Generator(device=dev)
        self.gen.manual_seed(seed)
        self.x = prior.sample(self.chains, self.gen)[:, self.pidx]
        self.logp = self._log_target(self.x)
        self.accepted = 0
        self.proposed = 0

    @classmethod
    def from_backend(cls, objective_backend, params, prior, **kw):
        'Wraps a ``mod16_b200.objective.ObjectiveBackend``: objective = -RMSD per draw.'
        def neg_rmsd(draws):
            sse, cnt = objective_backend(draws, as_torch=True)
            return -torch.sqrt(sse / cnt)
        return cls(neg_rmsd, params, prior.to(objective_backend.device), **kw)

    def full(self, x):
        'Free-parameter matrix (C, d) -> full (C, 11) draw matrix.'
        out = self.base.repeat(x.shape[0], 1)
        out[:, self.cols] = x
        return out

    def _log_target(self, x):
        lp = self._free_log_prior(x)
        ok = torch.isfinite(lp)
        ll = torch.full_like(lp, -math.inf)
        if bool(ok.any()):
            val = self.backend(self.full(x))
            ll = torch.where(ok & torch.isfinite(val), val.to(torch.float64), ll)
        return torch.where(ok, lp + ll, lp)

    def _free_log_prior(self, x):
        p = self.prior
        sub = Prior.__new__(Prior)
        sub.names = self.free
        sub.par = p.par[self.pidx]
        sub.kind_t = p.kind_t[self.pidx]
        return sub.log_prob(x)

    def step(self, gamma=None):
        C, d = self.chains, self.d
        g = self.gen
        if gamma is None:
            gamma = 2.38 / math.sqrt(2 * d)
        # two distinct partners, both different from the chain itself
        r = torch.rand((C, C), device=self.device, generator=g)
        r.fill_diagonal_(2.0)
        ab = torch.topk(r, 2, dim=1, largest=False).indices
        width = self.prior.width()[self.pidx]
        eps = torch.randn((C, d), dtype=torch.float64, device=self.device, generator=g) * \
            (self.scaling * width)
        prop = self.x + gamma * (self.x[ab[:, 0]] - self.x[ab[:, 1]]) + eps
        logp = self._log_target(prop)
        u = torch.rand(C, dtype=torch.float64, device=self.device, generator=g)
        acc = torch.log(u) < (logp - self.logp)
        self.x = torch.where(acc[:, None], prop, self.x)
        self.logp = torch.where(acc, logp, self.logp)
        self.accepted += int(acc.sum().item())
        self.proposed += C
        return acc

    def run(self, draws, tune=0, thin=1):
        '''
        ``tune`` burn-in iterations, then ``draws`` kept iterations (every ``thin``-th).
        Returns dict(trace (draws, chains, 11) float64 tensor, logp (draws, chains),
        accept_rate, names).
        '''
        for it in range(tune):
            self.step(1.0 if it % 10 == 9 else None)
        self.accepted = self.proposed = 0
        trace = torch.empty((draws, self.chains, 11), dtype=torch.float64, device=self.device)
        logp = torch.empty((draws, self.chains), dtype=torch.float64, device=self.device)
        for k in range(draws):
            for t in range(thin):
                self.step(1.0 if (k * thin + t) % 10 == 9 else None)
            trace[k] = self.full(self.x)
            logp[k] = self.logp
        return dict(trace=trace, logp=logp, accept_rate=self.accepted / max(self.proposed, 1),
                    names=list(MOD16.required_parameters), free=self.free)
