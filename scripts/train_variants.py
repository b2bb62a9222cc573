"""Compare plumbing variants of the B=4096 VAEL train step (encoder/decoder stay stock PyTorch)."""
import os, sys, itertools
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from vael_b200.mnist_task import build_model_dict, build_worlds_queries_matrix
from vael_b200.networks import MNISTPairsDecoder, MNISTPairsEncoder, MNISTPairsMLP
from vael_b200.train import GraphedTrainStep, train_step
from vael_b200.vael import MNISTPairsVAELModel
dev = torch.device("cuda", 0)
B = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
md, wq = build_model_dict(2, 10, device=dev), build_worlds_queries_matrix(2, 10).to(dev)
def run(channels_last, bench, fused):
    torch.backends.cudnn.benchmark = bench
    torch.manual_seed(0)
    model = MNISTPairsVAELModel(MNISTPairsEncoder(hidden_channels=64, latent_dim=23), MNISTPairsDecoder(label_dim=20, hidden_channels=64, latent_dim=8),
                                MNISTPairsMLP(in_features=15, n_facts=20), (15, 8), md, wq, dropout=0.5, is_train=True, device=dev, query_per_row=True).to(dev)
    if channels_last:
        model = model.to(memory_format=torch.channels_last)
    opt = torch.optim.Adam(model.parameters(), lr=1e-3, capturable=True, fused=fused)
    x = torch.randn(B, 1, 28, 56, device=dev)
    if channels_last:
        x = x.contiguous(memory_format=torch.channels_last)
    d = torch.randint(0, 10, (B, 2), device=dev)
    labels = torch.cat([d, d.sum(1, keepdim=True), -torch.ones(B, 1, dtype=torch.long, device=dev)], 1)
    step = GraphedTrainStep(model, opt, tuple(x.shape), tuple(labels.shape), device=dev)
    if channels_last:
        step.data = step.data.contiguous(memory_format=torch.channels_last)
    for _ in range(5): step(x, labels)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20): t = step(x, labels)
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / 20, float(t[0])
for cl, bench, fused in itertools.product((False, True), (False, True), (False, True)):
    try:
        ms, loss = run(cl, bench, fused)
        print(f"channels_last={cl} cudnn.benchmark={bench} fused_adam={fused}: {ms:.3f} ms/step  {B / ms * 1e3:.0f} samples/s  loss {loss:.3f}")
    except Exception as e:
        print(cl, bench, fused, "failed", repr(e)[:200])
