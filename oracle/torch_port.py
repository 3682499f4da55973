"""TEST / BASELINE INFRASTRUCTURE ONLY -- a plain-PyTorch restatement of the reference's CPU path.

The reference (pure Python on torch; /root/reference does not travel to the GPU box) runs its step as
``training_step -> loss.backward() -> Adam.step()`` on stock ``nn.Linear`` / autograd / ``torch.optim.Adam``
(ps_vae/model.py:14-63, ps_vae/lightning.py:67-131,204-205) and samples with ``decode(randn)`` and an autograd
Langevin loop (ps_vae/inference.py:10-110).  This file restates exactly that with the same torch library calls, so
timing it on the GPU box's host cores measures what the reference's own CPU path would cost there
(bench.py ``cpu_baseline`` and ``--impl reference``; kind = "port").  It is pinned like the numpy oracle:
tests/test_oracle_golden.py::test_torch_port_matches_golden checks it against the fixtures the unmodified reference
produced.  The product package never imports it.
"""
from __future__ import annotations

from typing import Optional

import torch
import torch.nn as nn
import torch.nn.functional as F


def _mlp(d_in, hidden, n_hidden, d_out):
    dims = [d_in] + [hidden] * n_hidden + [d_out]
    layers = []
    for j in range(len(dims) - 1):
        layers.append(nn.Linear(dims[j], dims[j + 1]))
        if j + 2 < len(dims):
            layers.append(nn.ReLU())
    return nn.Sequential(*layers)


class TorchVAE(nn.Module):
    """model.py:7-69 (hidden width/depth generalised like the product's VAEModel)."""

    def __init__(self, input_dim=512, latent_dim=64, normalize_decoder=False, hidden_dim=512, num_hidden_layers=2):
        super().__init__()
        self.normalize_decoder = normalize_decoder
        self.encoder_mu = _mlp(input_dim, hidden_dim, num_hidden_layers, latent_dim)
        self.encoder_sigma = _mlp(input_dim, hidden_dim, num_hidden_layers, latent_dim)
        self.decoder = _mlp(latent_dim, hidden_dim, num_hidden_layers, input_dim)

    def forward(self, x, eps: Optional[torch.Tensor] = None):
        mu = self.encoder_mu(x)
        log_sigma = self.encoder_sigma(x)
        sigma = torch.exp(0.5 * log_sigma)
        z = mu + sigma * (torch.randn_like(sigma) if eps is None else eps)
        return self.decode(z), mu, log_sigma

    def decode(self, z):
        x_hat = self.decoder(z)
        if self.normalize_decoder:
            x_hat = F.normalize(x_hat, p=2, dim=1)
        return x_hat


class TorchStep(nn.Module):
    """lightning.py:10-131 for a single-label (or absent) 1..n-layer latent classifier."""

    def __init__(self, input_dim=256, latent_dim=64, num_classes: Optional[int] = 2, hidden_dim=512, num_hidden_layers=2, kl_w=1.0, clf_w=1.0,
                 normalize_decoder=False, use_cos_loss=False):
        super().__init__()
        self.model = TorchVAE(input_dim, latent_dim, normalize_decoder, hidden_dim, num_hidden_layers)
        self.classifier = nn.Linear(latent_dim, num_classes) if num_classes else None
        self.kl_w, self.clf_w, self.use_cos_loss = kl_w, clf_w, use_cos_loss

    def loss(self, x, y, eps=None):
        x_hat, mu, log_sigma = self.model(x, eps)
        clf = F.cross_entropy(self.classifier(mu), y) if self.classifier is not None else 0
        if self.use_cos_loss:
            recon = F.cosine_embedding_loss(x_hat, x, torch.ones(x.size(0)))
        else:
            recon = F.mse_loss(x_hat, x, reduction="mean") / 10
        kl = -0.5 * torch.mean(torch.sum(1 + log_sigma - mu.pow(2) - log_sigma.exp(), dim=-1))
        return recon + self.kl_w * kl + self.clf_w * clf


def train_steps(step_mod: TorchStep, opt: torch.optim.Optimizer, x, y, n_steps: int):
    """The automatic-optimisation body Lightning runs per batch."""
    loss = None
    for _ in range(n_steps):
        opt.zero_grad()
        loss = step_mod.loss(x, y)
        loss.backward()
        opt.step()
    return float(loss.detach())


def unconditional(step_mod: TorchStep, n: int, latent_dim: int):
    """inference.py:22-25."""
    return step_mod.model.decode(torch.randn((n, latent_dim))).detach()


def conditional(step_mod: TorchStep, n: int, latent_dim: int, target: int, step_size=0.01, num_steps=100, noise_weight=1.0):
    """inference.py:72-105 (incl. the per-step history append and the two .item() calls the reference pays for)."""
    z = torch.randn((n, latent_dim), requires_grad=True)
    history = []
    for _ in range(num_steps):
        logits = step_mod.classifier(z)
        log_p_y = F.log_softmax(logits, dim=-1)[:, target]
        log_p_z = -0.5 * (z ** 2).sum(dim=1)
        tot = log_p_y + log_p_z
        grad = torch.autograd.grad(tot.sum(), z)[0]
        noise = torch.randn_like(z)
        z = z + 0.5 * (step_size ** 2) * grad + step_size * noise_weight * noise
        history.append(z.detach().cpu().numpy())
        z.requires_grad_()
        _ = (tot.mean().item(), log_p_y.exp().mean().item())
    return step_mod.model.decode(z).detach()
