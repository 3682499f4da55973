"""``LatentClassifier`` -- drop-in for ps_vae/latent_classifier.py:5-70 (same ctor, attributes, state-dict keys).

The module structure (``layers`` ModuleList of Linear/activation, ``output_layers`` ModuleDict of heads) is the
reference's, so checkpoints load unchanged.  ``forward`` runs through the CUDA library when the classifier is
owned by a ``PseudoSpeakerVAE`` on a GPU -- it is used inside the fused train step (lightning.py) and inside the
Langevin kernel (inference.py); called on its own it evaluates the small MLP with torch ops (it is a few hundred
FLOPs per row and not on the measured path).
"""
from __future__ import annotations

from typing import Dict, List, Union

import torch
import torch.nn as nn

_ACTIVATIONS = {"relu": nn.ReLU, "tanh": nn.Tanh, "sigmoid": nn.Sigmoid, "leaky_relu": nn.LeakyReLU}


class LatentClassifier(nn.Module):
    def __init__(self, input_dim: int, num_classes: Union[dict, int], num_layers: int = 1, hidden_dim: int = 128, activation: str = "relu"):
        super().__init__()
        self.single_label_mode = isinstance(num_classes, int)
        self.num_classes = num_classes
        self.input_dim = int(input_dim)
        self.hidden_dim = int(hidden_dim)
        self.num_layers = int(num_layers)
        self.activation = activation
        self.layers = nn.ModuleList()
        self.output_layers = nn.ModuleDict()
        if activation not in _ACTIVATIONS:
            raise ValueError(f"Unsupported activation: {activation}")
        if num_layers < 1:
            raise ValueError("num_layers must be >= 1")
        act = _ACTIVATIONS[activation]
        if self.single_label_mode:
            if num_layers == 1:
                self.layers.append(nn.Linear(input_dim, num_classes))
            else:
                self.layers.append(nn.Linear(input_dim, hidden_dim))
                for _ in range(num_layers - 2):
                    self.layers.append(act())
                    self.layers.append(nn.Linear(hidden_dim, hidden_dim))
                self.layers.append(act())
                self.layers.append(nn.Linear(hidden_dim, num_classes))
        else:
            if num_layers == 1:
                for label, c in num_classes.items():
                    self.output_layers[label] = nn.Linear(input_dim, c)
            else:
                self.layers.append(nn.Linear(input_dim, hidden_dim))
                for _ in range(num_layers - 2):
                    self.layers.append(act())
                    self.layers.append(nn.Linear(hidden_dim, hidden_dim))
                self.layers.append(act())
                for label, c in num_classes.items():
                    self.output_layers[label] = nn.Linear(hidden_dim, c)

    # -- layout helpers for engine.HotPath ---------------------------------------------------------
    @property
    def num_trunk_linears(self) -> int:
        return self.num_layers - 1

    def trunk_linears(self) -> List[nn.Linear]:
        lin = [m for m in self.layers if isinstance(m, nn.Linear)]
        return lin[:-1] if self.single_label_mode else lin

    def head_linears(self) -> List[nn.Linear]:
        if self.single_label_mode:
            return [[m for m in self.layers if isinstance(m, nn.Linear)][-1]]
        return list(self.output_layers.values())

    @property
    def head_names(self) -> List:
        return [None] if self.single_label_mode else list(self.output_layers.keys())

    @property
    def head_classes(self) -> List[int]:
        return [lin.out_features for lin in self.head_linears()]

    @property
    def label_classes(self) -> Dict[str, int]:
        """What ps_vae/lightning.py:95 reads (and the reference never defines)."""
        return dict(self.num_classes) if not self.single_label_mode else {"label": self.num_classes}

    def forward(self, x: torch.Tensor):
        for layer in self.layers:
            x = layer(x)
        if self.single_label_mode:
            return x
        return {label: head(x) for label, head in self.output_layers.items()}


if __name__ == "__main__":  # the reference's shape smoke (latent_classifier.py:72-82)
    clf = LatentClassifier(64, {"age": 3, "gender": 2}, num_layers=2)
    out = clf(torch.randn(5, 64))
    print({k: v.shape for k, v in out.items()})
