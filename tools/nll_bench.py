import sys; sys.path.insert(0, '/root/repo')
import torch, bench
dev = torch.device('cuda', 0)
torch.cuda.set_device(0)
for g in (False, True):
    print(bench.time_nll_training(dev, 'bf16x3', use_graph=g))
print(bench.time_nll_training(dev, 'bf16', use_graph=True))
