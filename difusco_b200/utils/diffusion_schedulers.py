"""Noise schedules and the inference time schedule - host side, float64 numpy.

Drop-in for the reference's `utils.diffusion_schedulers` (difusco/utils/diffusion_schedulers.py):
same class names, constructor arguments and attributes
  CategoricalDiffusion(T, schedule) -> .T .beta .Qs (T,2,2) .Q_bar (T+1,2,2) .sample(x0_onehot, t)   :46-82
  GaussianDiffusion(T, schedule)    -> .T .beta .betabar .alpha (T+1) .alphabar (T+1) .sample(x0, t)  :9-43
  InferenceSchedule(inference_schedule, T, inference_T)(i) -> (t1, t2)                                :85-111
These tables are tiny and stay on the host exactly as in the reference; the device only ever
sees the four per-step posterior constants derived from them (pl_meta_model.posterior_consts).
"""
import math

import numpy as np
import torch


def _beta_schedule(T, schedule):
  if schedule == "linear":
    return np.linspace(1e-4, 2e-2, T)
  if schedule == "cosine":
    steps = np.arange(0, T + 1, 1)

    def f(t):
      return np.cos(math.pi * 0.5 * (t / T + 0.008) / (1 + 0.008)) ** 2
    abar = f(steps) / f(0)
    return np.clip(1 - (abar[1:] / abar[:-1]), None, 0.999)
  raise ValueError(f"Unknown diffusion schedule: {schedule}")


class GaussianDiffusion(object):
  """Variance-preserving Gaussian forward process."""

  def __init__(self, T, schedule):
    self.T = T
    self.beta = _beta_schedule(T, schedule)
    self.betabar = np.cumprod(self.beta)
    self.alpha = np.concatenate((np.array([1.0]), 1 - self.beta))
    self.alphabar = np.cumprod(self.alpha)

  def sample(self, x0, t):
    shape = (x0.shape[0],) + (1,) * (x0.dim() - 1)
    ab = torch.from_numpy(self.alphabar[t]).view(shape).to(x0.device)
    eps = torch.randn_like(x0)
    return torch.sqrt(ab) * x0 + torch.sqrt(1.0 - ab) * eps, eps


class CategoricalDiffusion(object):
  """Two-state D3PM with uniform transition kernels Q_t = (1 - beta_t) I + beta_t/2 11^T."""

  def __init__(self, T, schedule):
    self.T = T
    self.beta = _beta_schedule(T, schedule)
    b = self.beta.reshape((-1, 1, 1))
    self.Qs = (1 - b) * np.eye(2).reshape((1, 2, 2)) + (b / 2) * np.ones((1, 2, 2))
    acc = np.eye(2)
    chain = [acc]
    for q in self.Qs:          # same left-to-right product order as the reference (:69-72): bit-exact
      acc = acc @ q
      chain.append(acc)
    self.Q_bar = np.stack(chain, axis=0)

  def sample(self, x0_onehot, t):
    qb = torch.from_numpy(self.Q_bar[t]).float().to(x0_onehot.device)
    xt = torch.matmul(x0_onehot, qb.reshape((qb.shape[0], 1, 2, 2)))
    return torch.bernoulli(xt[..., 1].clamp(0, 1))


class InferenceSchedule(object):
  def __init__(self, inference_schedule="linear", T=1000, inference_T=1000):
    self.inference_schedule = inference_schedule
    self.T = T
    self.inference_T = inference_T

  def _at(self, j):
    frac = float(j) / self.inference_T
    if self.inference_schedule == "linear":
      return self.T - int(frac * self.T)
    if self.inference_schedule == "cosine":
      return self.T - int(np.sin(frac * np.pi / 2) * self.T)
    raise ValueError("Unknown inference schedule: {}".format(self.inference_schedule))

  def __call__(self, i):
    assert 0 <= i < self.inference_T
    t1 = np.clip(self._at(i), 1, self.T)
    t2 = np.clip(self._at(i + 1), 0, self.T - 1)
    return t1, t2
