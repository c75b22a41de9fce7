"""Noise schedules and the inference time schedule - host side, float64 numpy.

Drop-in for the reference's `utils.diffusion_schedulers` (difusco/utils/diffusion_schedulers.py): same class names,
constructor arguments and public attributes
  CategoricalDiffusion(T, schedule) -> .T .beta .Qs (T,2,2) .Q_bar (T+1,2,2) .sample(x0_onehot, t)   (:46-82)
  GaussianDiffusion(T, schedule)    -> .T .beta .betabar .alpha (T+1) .alphabar (T+1) .sample(x0, t)  (:9-43)
  InferenceSchedule(inference_schedule, T, inference_T)(i) -> (t1, t2)                                (:85-111)
These tables are tiny and stay on the host exactly as in the reference; the device only ever sees the four per-step
posterior constants derived from them (pl_meta_model.COMetaModel.posterior_consts).  tests/test_host_cpu.py checks
every table bit-for-bit against tables produced by the reference itself (tests/golden/schedules.npz).
"""
import math

import numpy as np
import torch


class _NoiseSchedule(object):
  """Shared beta schedule: 'linear' = linspace(1e-4, 2e-2, T); 'cosine' = Nichol & Dhariwal (offset 0.008, clip .999)."""

  def __init__(self, T, schedule):
    self.T = T
    if schedule == "linear":
      self.beta = np.linspace(1e-4, 2e-2, T)
    elif schedule == "cosine":
      grid = np.arange(0, T + 1, 1)
      curve = np.cos(math.pi * 0.5 * (grid / T + 0.008) / (1 + 0.008)) ** 2
      ratio = curve / (np.cos(math.pi * 0.5 * (0 / T + 0.008) / (1 + 0.008)) ** 2)
      self.beta = np.clip(1 - (ratio[1:] / ratio[:-1]), None, 0.999)
    else:
      raise ValueError(f"Unknown diffusion schedule: {schedule}")


class GaussianDiffusion(_NoiseSchedule):
  """Variance-preserving Gaussian forward process: alphabar[t] = prod_{s<=t} (1 - beta_s), alphabar[0] = 1."""

  def __init__(self, T, schedule):
    super().__init__(T, schedule)
    self.betabar = np.cumprod(self.beta)
    self.alpha = np.concatenate((np.array([1.0]), 1 - self.beta))
    self.alphabar = np.cumprod(self.alpha)

  def sample(self, x0, t):
    """q(x_t | x_0): returns (x_t, epsilon) with one alphabar per leading-dim element."""
    lead = (x0.shape[0],) + (1,) * (x0.dim() - 1)
    keep = torch.from_numpy(self.alphabar[t]).view(lead).to(x0.device)
    noise = torch.randn_like(x0)
    return keep.sqrt() * x0 + (1.0 - keep).sqrt() * noise, noise


class CategoricalDiffusion(_NoiseSchedule):
  """Two-state D3PM with uniform transition kernels Q_t = (1 - beta_t) I + (beta_t / 2) 11^T; Q_bar[t] = Q_1 ... Q_t."""

  def __init__(self, T, schedule):
    super().__init__(T, schedule)
    b = self.beta.reshape((-1, 1, 1))
    self.Qs = (1 - b) * np.eye(2).reshape((1, 2, 2)) + (b / 2) * np.ones((2, 2)).reshape((1, 2, 2))
    running = np.eye(2)
    chain = [running]
    for step in self.Qs:       # left-to-right products in the reference's order (:69-72) -> identical float64 bits
      running = running @ step
      chain.append(running)
    self.Q_bar = np.stack(chain, axis=0)

  def sample(self, x0_onehot, t):
    """q(x_t | x_0) for one-hot x0 of shape (B, ..., 2): Bernoulli draw of the state-1 probability."""
    kernel = torch.from_numpy(self.Q_bar[t]).float().to(x0_onehot.device)
    probs = torch.matmul(x0_onehot, kernel.reshape((kernel.shape[0], 1, 2, 2)))
    return torch.bernoulli(probs[..., 1].clamp(0, 1))


class InferenceSchedule(object):
  """Maps inference step i in [0, inference_T) to the (source, target) training timesteps (t1, t2)."""

  def __init__(self, inference_schedule="linear", T=1000, inference_T=1000):
    self.inference_schedule = inference_schedule
    self.T = T
    self.inference_T = inference_T

  def _timestep(self, j):
    progress = float(j) / self.inference_T
    if self.inference_schedule == "linear":
      return self.T - int(progress * self.T)
    if self.inference_schedule == "cosine":
      return self.T - int(np.sin(progress * np.pi / 2) * self.T)
    raise ValueError("Unknown inference schedule: {}".format(self.inference_schedule))

  def __call__(self, i):
    assert 0 <= i < self.inference_T
    return np.clip(self._timestep(i), 1, self.T), np.clip(self._timestep(i + 1), 0, self.T - 1)
