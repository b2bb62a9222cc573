"""Encoder / decoder / MLP blocks.  These stay stock PyTorch (cuDNN) — BASELINE.json north_star:
"The encoder and decoder CNNs stay in PyTorch" — and exist here only so that the model and the
training-step benchmark run without the reference tree.  Layer names, shapes and forward
semantics follow models/vael_networks.py (MNIST :15-158, Mario :165-317) so that reference
checkpoints ({'model': state_dict}, utils/early_stopping.py:53-65) load unchanged.
"""
from __future__ import annotations

import torch
from torch import nn
from torch.nn import functional as F


def _conv_stack(module: nn.Module, prefix: str, channels, kernel, transpose=False, extra=None):
    """Registers `{prefix}_1..n` on `module` (names matter for state_dict compatibility)."""
    extra = extra or {}
    for i, (cin, cout) in enumerate(zip(channels[:-1], channels[1:]), 1):
        cls = nn.ConvTranspose2d if transpose else nn.Conv2d
        k = kernel[i - 1] if isinstance(kernel, (list, tuple)) and isinstance(kernel[0], (list, tuple)) else kernel
        module.add_module(f"{prefix}_{i}", cls(cin, cout, kernel_size=k, stride=2, padding=1, **extra.get(i, {})))


class MNISTPairsEncoder(nn.Module):
    """[B,1,28,56] -> (mu, logvar) [B, latent_dim]; 3x Conv2d(k4,s2,p1)+ReLU (+dropout after the first two)."""

    def __init__(self, img_channels=1, hidden_channels=32, latent_dim=8, dropout=0.5):
        super().__init__()
        self.img_channels, self.hidden_channels, self.latent_dim = img_channels, hidden_channels, latent_dim
        self.label_dim, self.unflatten_dim = 20, (3, 7)
        h = hidden_channels
        _conv_stack(self, "enc_block", [img_channels, h, 2 * h, 4 * h], 4)
        flat = 4 * h * self.unflatten_dim[0] * self.unflatten_dim[1]
        self.dense_mu = nn.Linear(flat, latent_dim)
        self.dense_logvar = nn.Linear(flat, latent_dim)
        self.dropout = nn.Dropout(p=dropout)

    def forward(self, x):
        x = self.dropout(F.relu(self.enc_block_1(x)))
        x = self.dropout(F.relu(self.enc_block_2(x)))
        x = F.relu(self.enc_block_3(x)).flatten(1)
        return self.dense_mu(x), self.dense_logvar(x)


class MNISTPairsDecoder(nn.Module):
    """[B, latent_dim + label_dim] -> [B,1,28,56] in (0,1)."""

    def __init__(self, img_channels=1, hidden_channels=32, latent_dim=32, label_dim=9, dropout=0.5, **params):
        super().__init__()
        self.img_channels, self.hidden_channels, self.latent_dim = img_channels, hidden_channels, latent_dim
        self.unflatten_dim, self.label_dim = (3, 7), label_dim
        h = hidden_channels
        self.dense = nn.Linear(latent_dim + label_dim, 4 * h * self.unflatten_dim[0] * self.unflatten_dim[1])
        _conv_stack(self, "dec_block", [4 * h, 2 * h, h, img_channels], [(5, 4), (4, 4), (4, 4)], transpose=True)
        self.dropout = nn.Dropout(p=dropout)

    def forward(self, x):
        x = self.dense(x).reshape(x.size(0), 4 * self.hidden_channels, *self.unflatten_dim)
        x = self.dropout(F.relu(self.dec_block_1(x)))
        x = self.dropout(F.relu(self.dec_block_2(x)))
        return torch.sigmoid(self.dec_block_3(x))


class MNISTPairsMLP(nn.Module):
    """z_sym -> logits of the two digit disjunctions, returned as a pair (models/vael_networks.py:144-158)."""

    def __init__(self, in_features=20, n_facts=10, hidden_channels=20):
        super().__init__()
        self.n_facts = n_facts
        self.hidden_layer = nn.Linear(in_features, hidden_channels)
        self.dense = nn.Linear(hidden_channels, n_facts)

    def logits(self, x):
        return self.dense(F.relu(self.hidden_layer(x)))

    def forward(self, x):
        return torch.split(self.logits(x), self.n_facts // 2, dim=1)


class MarioEncoder(nn.Module):
    """[B,3,100,100] -> (mu, logvar); 4x Conv2d(k5,s2,p1)+SELU+dropout."""

    def __init__(self, img_channels=3, hidden_channels=32, latent_dim=8, dropout=0.0):
        super().__init__()
        self.img_channels, self.hidden_channels, self.latent_dim = img_channels, hidden_channels, latent_dim
        h = hidden_channels
        _conv_stack(self, "enc_block", [img_channels, h, 2 * h, 4 * h, 8 * h], 5)
        self.dense_mu = nn.Linear(8 * h * 25, latent_dim)
        self.dense_logvar = nn.Linear(8 * h * 25, latent_dim)
        self.dropout = nn.Dropout(p=dropout)

    def forward(self, x):
        for i in range(1, 5):
            x = self.dropout(F.selu(getattr(self, f"enc_block_{i}")(x)))
        x = x.flatten(1)
        return self.dense_mu(x), self.dense_logvar(x)


class MarioDecoder(nn.Module):
    """[B, latent_dim_sub + label_dim] -> [B,3,100,100] in (0,1)."""

    def __init__(self, img_channels=3, hidden_channels=32, latent_dim_sub=32, label_dim=9, dropout=0.5):
        super().__init__()
        self.img_channels, self.hidden_channels, self.latent_dim = img_channels, hidden_channels, latent_dim_sub
        self.image_dim, self.label_dim = (3, 3), label_dim
        h = hidden_channels
        self.dense = nn.Linear(latent_dim_sub + label_dim, 8 * h * 25)
        _conv_stack(self, "dec_block", [8 * h, 4 * h, 2 * h, h, img_channels], 5, transpose=True,
                    extra={2: dict(output_padding=1), 4: dict(output_padding=1)})
        self.dropout = nn.Dropout(p=dropout)

    def forward(self, x):
        x = self.dense(x).reshape(x.size(0), 8 * self.hidden_channels, 5, 5)
        for i in range(1, 4):
            x = self.dropout(F.selu(getattr(self, f"dec_block_{i}")(x)))
        return torch.sigmoid(self.dec_block_4(x))


class MarioMLP(nn.Module):
    """z_sym -> 9 position logits of one frame (n_facts counts both frames)."""

    def __init__(self, in_features=20, n_facts=9, hidden_channels=32):
        super().__init__()
        self.n_facts = n_facts
        self.hidden_layer = nn.Linear(in_features, hidden_channels)
        self.dense = nn.Linear(hidden_channels, n_facts // 2)

    def forward(self, x):
        return self.dense(F.selu(self.hidden_layer(x)))
