"""torch.profiler view of one eager B=4096 train step: which aten ops own the copy / layout-conversion kernels."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from torch.profiler import profile, ProfilerActivity
from vael_b200.mnist_task import build_model_dict, build_worlds_queries_matrix
from vael_b200.networks import MNISTPairsDecoder, MNISTPairsEncoder, MNISTPairsMLP
from vael_b200.train import as_channels_last, train_step
from vael_b200.vael import MNISTPairsVAELModel
B = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
dev = torch.device("cuda", 0)
torch.manual_seed(0)
model = MNISTPairsVAELModel(MNISTPairsEncoder(hidden_channels=64, latent_dim=23), MNISTPairsDecoder(label_dim=20, hidden_channels=64, latent_dim=8),
                            MNISTPairsMLP(in_features=15, n_facts=20), (15, 8), build_model_dict(2, 10, device=dev),
                            build_worlds_queries_matrix(2, 10).to(dev), dropout=0.5, is_train=True, device=dev, query_per_row=True).to(dev)
model = model.to(memory_format=torch.channels_last)
opt = torch.optim.Adam(model.parameters(), lr=1e-3, fused=True)
x = as_channels_last(torch.randn(B, 1, 28, 56, device=dev))
d = torch.randint(0, 10, (B, 2), device=dev)
labels = torch.cat([d, d.sum(1, keepdim=True), -torch.ones(B, 1, dtype=torch.long, device=dev)], 1)
for _ in range(4):
    train_step(model, opt, x, labels)
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA], record_shapes=True) as prof:
    train_step(model, opt, x, labels)
    torch.cuda.synchronize()
print(prof.key_averages(group_by_input_shape=True).table(sort_by="cuda_time_total", row_limit=45, max_name_column_width=55, max_shapes_column_width=70))
