class _MVN:
    def __init__(self, cov):
        self.covariance_matrix = cov


class ScalarizedPosteriorTransform:
    """mean -> sum_m w_m mean_m ; covariance -> sum_m w_m^2 Cov_m (independent outputs)."""

    def __init__(self, weights):
        self.weights = weights

    def __call__(self, posterior):
        from botorch.posteriors import GPyTorchPosterior

        w = self.weights
        mean = (posterior.mean * w).sum(-1, keepdim=True)
        cov = sum(w[m] ** 2 * c for m, c in enumerate(posterior._covs))
        return GPyTorchPosterior(mean, [cov])
