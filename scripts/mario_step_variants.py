"""Mario B = 32 train step: plumbing variants around the stock-PyTorch CNNs (memory format, fused Adam), graphed."""
import os, sys, itertools
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from vael_b200.mario_task import compile_mario_family, mario_tables
from vael_b200.networks import MarioDecoder, MarioEncoder, MarioMLP
from vael_b200.train import GraphedTrainStep, mario_train_step
from vael_b200.vael import MarioVAELModel
dev = torch.device("cuda", 0)
B = int(sys.argv[1]) if len(sys.argv) > 1 else 32
tabs = [torch.from_numpy(t.astype("float32")).to(dev) for t in mario_tables(compile_mario_family((3, 3)))]
W_all, WQ_all, W_adm, WQ_adm = tabs
def run(cl, fused):
    torch.manual_seed(0)
    model = MarioVAELModel(MarioEncoder(hidden_channels=64, latent_dim=48, dropout=0.0),
                           MarioDecoder(hidden_channels=64, latent_dim_sub=30, label_dim=9, dropout=0.5),
                           MarioMLP(18, 18, 32), (18, 30), None, WQ_all, W_all, WQ_adm, W_adm, (3, 3), dropout=0.5,
                           is_train=True, device=dev).to(dev)
    if cl:
        model = model.to(memory_format=torch.channels_last)
    opt = torch.optim.Adam(model.parameters(), lr=1e-4, capturable=True, fused=fused)
    x = torch.rand(B, 3, 200, 100, device=dev)
    if cl:
        x = x.contiguous(memory_format=torch.channels_last)
    moves = torch.eye(4, dtype=torch.long, device=dev)[torch.randint(0, 4, (B,), device=dev)]
    kw = dict(memory_format=torch.channels_last) if cl else {}
    stepper = GraphedTrainStep(model, opt, tuple(x.shape), tuple(moves.shape), device=dev, step_fn=mario_train_step, **kw)
    for _ in range(3): stepper(x, moves)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(30): t = stepper(x, moves)
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / 30, float(t.item() if t.dim() == 0 else t[0])
for cl, fused in itertools.product((False, True), (False, True)):
    try:
        ms, loss = run(cl, fused)
        print(f"B={B} channels_last={cl} fused_adam={fused}: {ms:.3f} ms/step {B / ms * 1e3:.0f} samples/s loss {loss:.2f}", flush=True)
    except Exception as e:
        print(cl, fused, "failed", repr(e)[:300], flush=True)
