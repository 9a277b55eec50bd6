"""NLL training steps (AD-22, batch 256: BASELINE configs[1]) for ncu launch lists."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, bench
dev = torch.device("cuda", 0)
print(bench.time_nll_training(dev, sys.argv[1] if len(sys.argv) > 1 else "bf16x3", steps=1, warmup=2, use_graph=False))
