"""Phase timeline of one CTA of sf_pcblock_ffn1 (%globaltimer stamps of the first worker thread)."""
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
tr = torch.zeros(32, dtype=torch.int64, device="cuda")
sfb.lib().sf_debug_ffn1_trace(tr.data_ptr())
sfb.pcblock_ffn1(x, ffn1)
torch.cuda.synchronize()
sfb.lib().sf_debug_ffn1_trace(None)
t = tr.cpu().tolist()
t0 = t[0]
print(f"staging            {t[1] - t0:7d} ns")
for hh in range(4):
    a, b, c = t[2 + 3 * hh], t[3 + 3 * hh], t[4 + 3 * hh]
    prev = t[1] if hh == 0 else t[4 + 3 * (hh - 1)]
    print(f"chunk {hh}: wait D1 {a - prev:6d}  GELU {b - a:6d}  wait G free {c - b:6d}   (t = {c - t0} ns)")
print(f"write G + wait Y   {t[27] - t[26]:7d} ns (after last chunk {t[26] - t[13]} ns)")
print(f"final pass         {t[28] - t[27]:7d} ns")
print(f"total              {t[28] - t0:7d} ns")
