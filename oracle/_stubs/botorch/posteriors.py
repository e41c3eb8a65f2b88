class _MVN:
    def __init__(self, cov):
        self.covariance_matrix = cov


class GPyTorchPosterior:
    """mean: (q, M); per-output dense covariances (block diagonal across outputs)."""

    def __init__(self, mean, covs):
        self.mean = mean
        self._covs = covs
        self.mvn = _MVN(covs[0]) if len(covs) == 1 else None
