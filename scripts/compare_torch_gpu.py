"""Like-for-like bar: the reference's own torch op sequence for the hot path, run on the SAME B200
(cuBLAS / ATen kernels; TF32 off = the reference's default, and on), against the streamflow_b200 operators.
Sintel size, one clip (3 pairs / 3 maps), CUDA events, GPU time per call with the CPU launch cost included for
both sides (eager), plus a CUDA-graph replay of the full 12-iteration hot path for both."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
torch.set_grad_enabled(False)   # inference-only operators
import bench
import streamflow_b200 as sfb
from oracle import torch_port as tp

dev = torch.device("cuda", 0)
host = bench.make_inputs(0)
t = {k: host[k].to(dev) for k in ("fm_nhwc", "inps", "mfs", "coords")}
fmaps = t["fm_nhwc"].permute(0, 1, 4, 2, 3)
w_qk, w_v = host["w_qk"].to(dev), host["w_v"].to(dev)
class _A: pass
att = sfb.Attention(args=_A(), dim=128, heads=1, max_pos_size=160, dim_head=128).to(dev)
agg = sfb.Aggregate(args=_A(), dim=128, heads=1, dim_head=128).to(dev)
with torch.no_grad():
    att.to_qk.weight.copy_(w_qk.view(256, 128, 1, 1)); agg.to_v.weight.copy_(w_v.view(128, 128, 1, 1)); agg.gamma.fill_(0.8)


def timeit(fn, n=10):
    for _ in range(2): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e3


@torch.no_grad()
def ref_step(autocast):
    pyrs = [tp.CpuCorrPyramid(fmaps[:, i], fmaps[:, i + 1]) for i in range(3)]
    with torch.autocast("cuda", dtype=torch.float16, enabled=autocast):
        attn = tp.cpu_attention(t["inps"], w_qk)
    for it in range(12):
        feats = torch.stack([pyrs[i](t["coords"][it, i]) for i in range(3)], 0)
        with torch.autocast("cuda", dtype=torch.float16, enabled=autocast):
            out = tp.cpu_aggregate(attn, t["mfs"], w_v, 0.8)
    return feats, out


@torch.no_grad()
def our_step():
    blocks = [sfb.CorrBlock(fmaps[:, i], fmaps[:, i + 1], radius=4) for i in range(3)]
    group = sfb.CorrGroup(blocks)
    handle = att(t["inps"])
    for it in range(12):
        feats = group([t["coords"][it, i] for i in range(3)])
        out = agg(handle, t["mfs"])
    return feats, out


rows = []
for tf32 in (False, True):
    torch.backends.cuda.matmul.allow_tf32 = tf32
    torch.backends.cudnn.allow_tf32 = tf32
    tag = "TF32 on" if tf32 else "TF32 off (reference default)"
    pyr = tp.CpuCorrPyramid(fmaps[:, 0], fmaps[:, 1])
    with torch.autocast("cuda", dtype=torch.float16):
        attn = tp.cpu_attention(t["inps"], w_qk)
    rows.append((f"torch build, 1 pair, {tag}", timeit(lambda: tp.CpuCorrPyramid(fmaps[:, 0], fmaps[:, 1]))))
    rows.append((f"torch lookup, 1 pair, {tag}", timeit(lambda: pyr(t["coords"][0, 0]))))
    rows.append((f"torch attention fp16 autocast, 3 maps, {tag}", timeit(lambda: torch.autocast("cuda", dtype=torch.float16).__enter__() and None or tp.cpu_attention(t["inps"], w_qk), n=4)))
    def _agg():
        with torch.autocast("cuda", dtype=torch.float16):
            return tp.cpu_aggregate(attn, t["mfs"], w_v, 0.8)
    rows.append((f"torch aggregate fp16 autocast, 3 maps, {tag}", timeit(_agg)))
    rows.append((f"torch full hot path (12 iters, autocast GMA), {tag}", timeit(lambda: ref_step(True), n=3)))
    del pyr, attn
blk = sfb.CorrBlock(fmaps[:, 0], fmaps[:, 1]); hd = att(t["inps"])
rows.append(("ours build, 1 pair (f16 operands)", timeit(lambda: sfb.CorrBlock(fmaps[:, 0], fmaps[:, 1]))))
rows.append(("ours build, 1 pair (f16x2, fp32-faithful)", timeit(lambda: sfb.CorrBlock(fmaps[:, 0], fmaps[:, 1], precision="f16x2"))))
rows.append(("ours lookup, 1 pair", timeit(lambda: blk(t["coords"][0, 0]))))
rows.append(("ours attention, 3 maps", timeit(lambda: att(t["inps"]), n=4)))
rows.append(("ours aggregate, 3 maps", timeit(lambda: agg(hd, t["mfs"]))))
rows.append(("ours full hot path (12 iters)", timeit(our_step, n=5)))
for name, us in rows:
    print(f"{name:62s} {us:10.1f} us")
