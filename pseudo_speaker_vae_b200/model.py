"""``VAEModel`` -- drop-in for ps_vae/model.py:7-69 whose arithmetic runs in the sm_100a CUDA library.

Same constructor (``input_dim=512, latent_dim=64, normalize_decoder=False``), same sub-module names and
state-dict keys (``encoder_mu.{0,2,4}.{weight,bias}``, ``encoder_sigma.*``, ``decoder.*``), same default
``nn.Linear`` initialisation in the same construction order (so a given ``torch.manual_seed`` yields the same
initial weights as the reference), same ``forward(x) -> (x_hat, mu, log_sigma)`` and ``decode(z) -> x_hat``.

Additive keyword arguments (reference values are the defaults): ``hidden_dim=512``, ``num_hidden_layers=2``
(the reference hard-codes both, model.py:14-36; BASELINE config 5 widens them) and ``precision='fp32'|'bf16'``.

Differences a caller can observe:
  * the reparameterisation noise comes from the library's counter-based Philox generator (seeded from
    ``torch.initial_seed()``), not from torch's global generator; pass ``eps=`` to inject the draw;
  * with autograd on, the outputs carry a graph whose backward pass is one fused library call (``psvae_vae_backward``: the
    activations are recomputed, nothing is kept between the two calls); the gradient with respect to ``x`` itself is not
    computed.  ``PseudoSpeakerVAE.training_step`` (lightning.py) does not use this route: it computes its loss and all
    gradients in one fused call;
  * the module must live on a CUDA (B200) device: there is no CPU path.
"""
from __future__ import annotations

from typing import List, Optional, Tuple

import torch
import torch.nn as nn


def _mlp(d_in: int, hidden: int, n_hidden: int, d_out: int) -> nn.Sequential:
    dims = [d_in] + [hidden] * n_hidden + [d_out]
    layers: List[nn.Module] = []
    for j in range(len(dims) - 1):
        layers.append(nn.Linear(dims[j], dims[j + 1]))
        if j + 2 < len(dims):
            layers.append(nn.ReLU())
    return nn.Sequential(*layers)


class VAEModel(nn.Module):
    def __init__(self, input_dim: int = 512, latent_dim: int = 64, normalize_decoder: bool = False, *, hidden_dim: int = 512,
                 num_hidden_layers: int = 2, precision: str = "fp32"):
        super().__init__()
        self.input_dim = int(input_dim)
        self.latent_dim = int(latent_dim)
        self.hidden_dim = int(hidden_dim)
        self.num_hidden_layers = int(num_hidden_layers)
        self.normalize_decoder = bool(normalize_decoder)
        self.precision = precision
        # construction order = the reference's (model.py:14-36): same RNG consumption, same initial weights
        self.encoder_mu = _mlp(self.input_dim, self.hidden_dim, self.num_hidden_layers, self.latent_dim)
        self.encoder_sigma = _mlp(self.input_dim, self.hidden_dim, self.num_hidden_layers, self.latent_dim)
        self.decoder = _mlp(self.latent_dim, self.hidden_dim, self.num_hidden_layers, self.input_dim)
        self._hot = None          # HotPath; built lazily, or installed by the owning PseudoSpeakerVAE

    # -- used by engine.HotPath to lay the parameters out in the flat arena
    def linears(self, name: str) -> List[nn.Linear]:
        return [m for m in getattr(self, name) if isinstance(m, nn.Linear)]

    def hot_path(self):
        if self._hot is None:
            from .engine import HotPath

            object.__setattr__(self, "_hot", HotPath(self, None, self.precision))
        return self._hot

    def _adopt(self, hot) -> None:
        object.__setattr__(self, "_hot", hot)

    def forward(self, x: torch.Tensor, eps: Optional[torch.Tensor] = None) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor]:
        """x [batch, input_dim] -> (x_hat [batch, input_dim], mu [batch, latent], log_sigma [batch, latent])  (model.py:38-63)."""
        return self.hot_path().forward(x, eps)

    def decode(self, z: torch.Tensor) -> torch.Tensor:
        """model.py:65-69."""
        return self.hot_path().decode(z)


if __name__ == "__main__":  # the reference's shape smoke (model.py:71-75), on the GPU
    model = VAEModel(784, 20).to("cuda")
    z = torch.randn(32, 784, device="cuda")
    x_hat, mu, sigma = model(z)
    print(x_hat.shape, mu.shape, sigma.shape)
