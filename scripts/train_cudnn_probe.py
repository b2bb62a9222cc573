"""B = 4096 graphed VAEL train step with cuDNN's heuristic algorithm choice against its benchmark (autotune) mode."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from vael_b200.mnist_task import build_model_dict, build_worlds_queries_matrix
from vael_b200.networks import MNISTPairsDecoder, MNISTPairsEncoder, MNISTPairsMLP
from vael_b200.train import GraphedTrainStep, as_channels_last
from vael_b200.vael import MNISTPairsVAELModel
dev = torch.device("cuda", 0)
B = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
md, wq = build_model_dict(2, 10, device=dev), build_worlds_queries_matrix(2, 10).to(dev)
def run(bench):
    torch.backends.cudnn.benchmark = bench
    torch.manual_seed(0)
    model = MNISTPairsVAELModel(MNISTPairsEncoder(hidden_channels=64, latent_dim=23), MNISTPairsDecoder(label_dim=20, hidden_channels=64, latent_dim=8),
                                MNISTPairsMLP(in_features=15, n_facts=20), (15, 8), md, wq, dropout=0.5, is_train=True, device=dev,
                                query_per_row=True).to(dev).to(memory_format=torch.channels_last)
    opt = torch.optim.Adam(model.parameters(), lr=1e-3, capturable=True, fused=True)
    x = as_channels_last(torch.randn(B, 1, 28, 56, device=dev))
    d = torch.randint(0, 10, (B, 2), device=dev)
    labels = torch.cat([d, d.sum(1, keepdim=True), -torch.ones(B, 1, dtype=torch.long, device=dev)], 1)
    step = GraphedTrainStep(model, opt, tuple(x.shape), tuple(labels.shape), device=dev, memory_format=torch.channels_last)
    for _ in range(5): step(x, labels)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20): t = step(x, labels)
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / 20, float(t[0])
for bench in (False, True, False, True):
    ms, loss = run(bench)
    print(f"cudnn.benchmark={bench}: {ms:.3f} ms/step  {B / ms * 1e3:.0f} samples/s  loss {loss:.3f}", flush=True)
