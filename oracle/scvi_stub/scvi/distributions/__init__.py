"""scvi.distributions.NegativeBinomialMixture / log_mixture_nb of scvi-tools 0.20.0,
restated (TEST INFRASTRUCTURE ONLY).  Used by the reference at
src/spVIPES/module/spVIPESmodule.py:759 and :823-824.
"""
import torch
import torch.nn.functional as F
from torch.distributions import Distribution, constraints
from torch.distributions.utils import broadcast_all


def log_mixture_nb(x, mu_1, mu_2, theta_1, theta_2, pi_logits, eps=1e-8):
    if theta_2 is not None:
        log_nb_1 = log_nb_positive(x, mu_1, theta_1)
        log_nb_2 = log_nb_positive(x, mu_2, theta_2)
    else:
        theta = theta_1
        if theta.ndimension() == 1:
            theta = theta.view(1, theta.size(0))
        log_theta_mu_1_eps = torch.log(theta + mu_1 + eps)
        log_theta_mu_2_eps = torch.log(theta + mu_2 + eps)
        lgamma_x_theta = torch.lgamma(x + theta)
        lgamma_theta = torch.lgamma(theta)
        lgamma_x_plus_1 = torch.lgamma(x + 1)
        log_nb_1 = (
            theta * (torch.log(theta + eps) - log_theta_mu_1_eps)
            + x * (torch.log(mu_1 + eps) - log_theta_mu_1_eps)
            + lgamma_x_theta
            - lgamma_theta
            - lgamma_x_plus_1
        )
        log_nb_2 = (
            theta * (torch.log(theta + eps) - log_theta_mu_2_eps)
            + x * (torch.log(mu_2 + eps) - log_theta_mu_2_eps)
            + lgamma_x_theta
            - lgamma_theta
            - lgamma_x_plus_1
        )
    logsumexp = torch.logsumexp(torch.stack((log_nb_1, log_nb_2 - pi_logits)), dim=0)
    softplus_pi = F.softplus(-pi_logits)
    return logsumexp - softplus_pi


def log_nb_positive(x, mu, theta, eps=1e-8):
    log_theta_mu_eps = torch.log(theta + mu + eps)
    return (
        theta * (torch.log(theta + eps) - log_theta_mu_eps)
        + x * (torch.log(mu + eps) - log_theta_mu_eps)
        + torch.lgamma(x + theta)
        - torch.lgamma(theta)
        - torch.lgamma(x + 1)
    )


class NegativeBinomialMixture(Distribution):
    arg_constraints = {
        "mu1": constraints.greater_than_eq(0),
        "mu2": constraints.greater_than_eq(0),
        "theta1": constraints.greater_than_eq(0),
        "mixture_probs": constraints.half_open_interval(0.0, 1.0),
        "mixture_logits": constraints.real,
    }
    support = constraints.nonnegative_integer

    def __init__(self, mu1, mu2, theta1, mixture_logits, theta2=None, validate_args=False):
        self.mu1, self.theta1, self.mu2, self.mixture_logits = broadcast_all(mu1, theta1, mu2, mixture_logits)
        super().__init__(validate_args=validate_args)
        self.theta2 = None if theta2 is None else broadcast_all(mu1, theta2)[1]

    @property
    def mean(self):
        pi = torch.sigmoid(self.mixture_logits)
        return pi * self.mu1 + (1 - pi) * self.mu2

    def log_prob(self, value):
        return log_mixture_nb(value, self.mu1, self.mu2, self.theta1, self.theta2, self.mixture_logits, eps=1e-8)
