"""Likelihoods (reference ``src/likelihoods.py``)."""
from __future__ import annotations

import math

import torch
import torch.nn as nn


class Likelihood:
    def mnll_batch_estimate(self, *args, **kwargs):
        return 0.0


class GaussianLikelihood(nn.Module, Likelihood):
    def __init__(self, sigma: float = 1.0):
        super().__init__()
        self.sigma = nn.Parameter(torch.tensor(sigma))

    def mnll_batch_estimate(self, y: torch.Tensor, y_hat: torch.Tensor, n: int) -> torch.Tensor:
        """-n/(m*n_mc) * sum_{b,i,s} log N(y[b,i] | y_hat[b,i,s], sigma)
        (``src/likelihoods.py:18-29``), as one closed-form reduction instead of a Python loop
        over outputs building ``Normal`` objects.

        :param y: targets (m, n_out);  :param y_hat: predictions (m, n_out, n_mc);
        :param int n: data-set size.
        """
        m, n_out, n_mc = y_hat.size()
        sq = (y.reshape(m, n_out, 1) - y_hat).square().sum()
        return self.mnll_from_sq_error(sq, m, n_out, n_mc, n)

    def mnll_from_sq_error(self, sq: torch.Tensor, m: int, n_out: int, n_mc: int, n: int) -> torch.Tensor:
        """The same estimator given sum_{b,i,s} (y - y_hat)^2 (which the fused last-layer
        kernel reduces on the fly)."""
        count = m * n_out * n_mc
        log_prob_sum = -count * (torch.log(self.sigma) + 0.5 * math.log(2.0 * math.pi)) - 0.5 * sq / self.sigma ** 2
        return -n / (m * n_mc) * log_prob_sum
