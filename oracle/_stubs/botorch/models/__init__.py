import torch

from botorch.models.model import Model
from botorch.posteriors import GPyTorchPosterior


class SingleTaskGP(Model):
    """Wraps one ``oracle.gp.OracleObjective``."""

    num_outputs = 1

    def __init__(self, objective):
        self.objective = objective

    def posterior(self, X, observation_noise=False):
        from oracle import gp as ogp

        mean, cov = ogp.posterior(self.objective, X, observation_noise=observation_noise)
        return GPyTorchPosterior(mean.unsqueeze(-1), [cov])


class ModelListGP(Model):
    def __init__(self, *models):
        self.models = list(models)

    @property
    def num_outputs(self):
        return len(self.models)

    def posterior(self, X, observation_noise=False):
        ps = [m.posterior(X, observation_noise=observation_noise) for m in self.models]
        mean = torch.cat([p.mean for p in ps], dim=-1)
        return GPyTorchPosterior(mean, [p._covs[0] for p in ps])
