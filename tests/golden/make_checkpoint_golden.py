"""Golden checkpoint written by the REFERENCE classes (build container only).

Builds the reference's MNISTPairsVAELModel and MarioVAELModel (models/vael.py, models/vael_networks.py,
unmodified; problog stubbed as in make_golden.py) at small channel widths, takes one Adam step on a
synthetic loss so the optimizer has state, and saves them the way the reference does
(utils/early_stopping.py:53-65): torch.save({'model': state_dict, 'optimizer': state_dict}).
tests/test_host_cpu.py then loads these files into this repo's models.  /root/reference does not exist on
the GPU box, which is why the (small) files are committed."""
import json
import os
import sys

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
from make_golden import REF, install_problog_stub  # noqa: E402

install_problog_stub()
from models import vael as ref_vael  # noqa: E402
from models import vael_networks as ref_net  # noqa: E402


def one_step(model, params_loss):
    opt = torch.optim.Adam(model.parameters(), lr=1e-3)
    params_loss().backward()
    opt.step()
    return opt


def main():
    torch.manual_seed(0)
    dev = torch.device("cpu")
    w_q = torch.zeros(100, 20)
    m = ref_vael.MNISTPairsVAELModel(ref_net.MNISTPairsEncoder(hidden_channels=4, latent_dim=23),
                                     ref_net.MNISTPairsDecoder(label_dim=20, hidden_channels=4, latent_dim=8),
                                     ref_net.MNISTPairsMLP(in_features=15, n_facts=20), (15, 8), None, w_q, dropout=0.5,
                                     is_train=True, device=dev)
    opt = one_step(m, lambda: sum((p ** 2).sum() for p in m.parameters()))
    torch.save({"model": m.state_dict(), "optimizer": opt.state_dict()}, os.path.join(HERE, "ref_checkpoint_mnist.pt"))
    keys = {"MNISTPairsVAELModel": {k: list(v.shape) for k, v in m.state_dict().items()}}

    # W_all / W_adm are stored as nested lists (the task script wraps them: mario_task_VAEL.py does torch.tensor)
    tabs = [torch.as_tensor(torch.load(os.path.join(REF, "utils/mario_utils/problog_matrices/3x3", n + ".pt")))
            for n in ("WQ_all", "W_all", "WQ_adm", "W_adm")]
    mm = ref_vael.MarioVAELModel(ref_net.MarioEncoder(img_channels=3, hidden_channels=2, latent_dim=48),
                                 ref_net.MarioDecoder(img_channels=3, hidden_channels=2, latent_dim_sub=30, label_dim=9),
                                 ref_net.MarioMLP(in_features=18, n_facts=18, hidden_channels=32), (18, 30), None,
                                 *tabs, (3, 3), dropout=0.5, is_train=True, device=dev)
    opt = one_step(mm, lambda: sum((p ** 2).sum() for p in mm.parameters()))
    torch.save({"model": mm.state_dict(), "optimizer": opt.state_dict()}, os.path.join(HERE, "ref_checkpoint_mario.pt"))
    keys["MarioVAELModel"] = {k: list(v.shape) for k, v in mm.state_dict().items()}
    json.dump(keys, open(os.path.join(HERE, "checkpoint_layouts.json"), "w"), indent=1, sort_keys=True)
    for n in ("ref_checkpoint_mnist.pt", "ref_checkpoint_mario.pt"):
        print(n, os.path.getsize(os.path.join(HERE, n)))


if __name__ == "__main__":
    main()
