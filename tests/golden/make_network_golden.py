"""Golden fixture for the encoder / decoder / MLP mirrors: run in the build container only.

Imports the REFERENCE modules (/root/reference/models/vael_networks.py, unmodified), builds them at the
shipped channel widths to record the state_dict layout (names + shapes: what a reference checkpoint holds,
utils/early_stopping.py:53-65), and at a small width to record weights, a seeded input and the outputs,
so that tests/test_host_cpu.py can check that this repo's modules load a reference state_dict and
reproduce the reference forward.  /root/reference does not exist on the GPU box."""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.environ.get("VAEL_REFERENCE", "/root/reference"))
from models import vael_networks as ref  # noqa: E402


def layout(m):
    return {k: list(v.shape) for k, v in m.state_dict().items()}


def main():
    torch.manual_seed(0)
    out = {}
    full = {
        "MNISTPairsEncoder": ref.MNISTPairsEncoder(hidden_channels=64, latent_dim=23),
        "MNISTPairsDecoder": ref.MNISTPairsDecoder(label_dim=20, hidden_channels=64, latent_dim=8),
        "MNISTPairsMLP": ref.MNISTPairsMLP(in_features=15, n_facts=20),
        "MarioEncoder": ref.MarioEncoder(img_channels=3, hidden_channels=64, latent_dim=48),
        "MarioDecoder": ref.MarioDecoder(img_channels=3, hidden_channels=64, latent_dim_sub=30, label_dim=9),
        "MarioMLP": ref.MarioMLP(in_features=18, n_facts=18, hidden_channels=32),
    }
    import json
    json.dump({k: layout(m) for k, m in full.items()}, open(os.path.join(HERE, "network_layouts.json"), "w"), indent=1,
              sort_keys=True)

    small = {
        "MNISTPairsEncoder": (ref.MNISTPairsEncoder(hidden_channels=4, latent_dim=23), torch.randn(3, 1, 28, 56)),
        "MNISTPairsDecoder": (ref.MNISTPairsDecoder(label_dim=20, hidden_channels=4, latent_dim=8), torch.randn(3, 28)),
        "MNISTPairsMLP": (ref.MNISTPairsMLP(in_features=15, n_facts=20), torch.randn(3, 15)),
        "MarioEncoder": (ref.MarioEncoder(img_channels=3, hidden_channels=2, latent_dim=48), torch.rand(2, 3, 100, 100)),
        "MarioDecoder": (ref.MarioDecoder(img_channels=3, hidden_channels=2, latent_dim_sub=30, label_dim=9), torch.randn(2, 39)),
        "MarioMLP": (ref.MarioMLP(in_features=18, n_facts=18, hidden_channels=32), torch.randn(3, 18)),
    }
    for name, (m, x) in small.items():
        m.eval()
        with torch.no_grad():
            y = m(x)
        ys = y if isinstance(y, (tuple, list)) else (y,)
        out[f"{name}/x"] = x.numpy()
        for i, t in enumerate(ys):
            out[f"{name}/y{i}"] = t.numpy()
        for k, v in m.state_dict().items():
            out[f"{name}/w/{k}"] = v.numpy()
    np.savez_compressed(os.path.join(HERE, "networks.npz"), **out)
    print("wrote network_layouts.json, networks.npz", os.path.getsize(os.path.join(HERE, "networks.npz")))


if __name__ == "__main__":
    main()
