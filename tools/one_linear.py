import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from pangnn_b200 import ops
n, k = int(sys.argv[1]), int(sys.argv[2])
x = torch.randn(1_000_000, k, device="cuda:0"); w = torch.randn(n, k, device="cuda:0") / 8; b = torch.randn(n, device="cuda:0")
for _ in range(3): y = ops.node_linear(x, w, None, ops.ACT_NONE)
torch.cuda.synchronize(); print("ok", float(y.abs().mean()))
