import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.nn as nn
import streamflow_b200 as sfb
torch.set_grad_enabled(False)
torch.manual_seed(0)
C, P, h, w = 324, int(sys.argv[1]) if len(sys.argv) > 1 else 3, 55, 128
ffn1 = nn.Sequential(nn.Conv2d(C, 486, 1), nn.GELU(), nn.Conv2d(486, C, 1)).cuda().eval()
x = torch.randn(P, C, h, w, device="cuda")
for _ in range(3):
    sfb.pcblock_ffn1(x, ffn1)
torch.cuda.synchronize()
