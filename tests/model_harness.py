"""Test-side restatement of the StreamFlow refinement loop around the hot path (TEST INFRASTRUCTURE ONLY).

The reference model cannot run on the GPU box (its sources do not travel and it needs `timm`), so the end-to-end
criterion of BASELINE.json -- "final flow after 12 iterations within 0.01 px mean EPE of the reference, using
identical random-init weights and synthetic frames" -- is checked on this harness: the SAME surrounding network
is run twice, once on the reference's L1 operators (torch restatement, oracle/torch_port.py, on the same device)
and once on the B200 operators, and the two final flows are compared.

What is restated from the reference, with its structure and hyper-parameters:
  * forward loop ........... core/models/streamflow.py:95-147 (T-1 CorrBlocks, Attention once, per iteration:
                             lookups -> update block -> coords += delta; convex 8x upsampling of the last flow)
  * update block ........... SKUpdateBlock_TAM_v3, core/update.py:739-782
  * motion encoder ......... SKMotionEncoder6_Deep_nopool_res, core/update.py:313-339
  * PCBlock ................ PCBlock4_Deep_nopool_res, core/update.py:12-36 (k_conv [1,15], updater [1,7])
  * temporal transformer ... TemporalLayer2 / TransformerBlock, core/update.py:459-513 -- re-randomised instead of
                             zero-initialised, and gamma ~ U(0.5, 1.5) instead of 0, otherwise the GMA path would
                             contribute nothing at init
What is NOT restated: the Twins-SVT encoder (needs timm); a small strided conv encoder with the same output
contract ([B, T, 256, H/8, W/8], fnet fp32 after `.float()`, cnet under autocast) stands in for it -- the hot
path only sees its outputs.
"""
from __future__ import annotations

import torch
import torch.nn as nn
import torch.nn.functional as F


class PCBlock(nn.Module):
    def __init__(self, c_in, c_out, k_conv):
        super().__init__()
        self.conv_list = nn.ModuleList(
            [nn.Conv2d(c_in, c_in, k, stride=1, padding=k // 2, groups=c_in) for k in k_conv])
        hid = int(1.5 * c_in)
        self.ffn1 = nn.Sequential(nn.Conv2d(c_in, hid, 1), nn.GELU(), nn.Conv2d(hid, c_in, 1))
        self.pw = nn.Conv2d(c_in, c_in, 1)
        self.ffn2 = nn.Sequential(nn.Conv2d(c_in, hid, 1), nn.GELU(), nn.Conv2d(hid, c_out, 1))

    def forward(self, x):
        x = F.gelu(x + self.ffn1(x))
        for conv in self.conv_list:
            x = F.gelu(x + conv(x))
        x = F.gelu(x + self.pw(x))
        return self.ffn2(x)


class MotionEncoder(nn.Module):
    def __init__(self, k_conv=(1, 15), out_dim=128):
        super().__init__()
        self.convc1 = PCBlock(324, 256, k_conv)
        self.convc2 = PCBlock(256, 192, k_conv)
        self.convf1 = nn.Conv2d(2, 128, 1)
        self.convf2 = PCBlock(128, 64, k_conv)
        self.conv = PCBlock(64 + 192, out_dim - 2, k_conv)

    def forward(self, flow, corr):
        cor = self.convc2(F.gelu(self.convc1(corr)))
        flo = self.convf2(self.convf1(flow))
        out = self.conv(torch.cat([cor, flo], dim=1))
        return torch.cat([out, flow], dim=1)


class TemporalBlock(nn.Module):
    """LayerNorm -> single-head attention over the T-1 tokens of a pixel -> MLP (ratio 2), residual."""

    def __init__(self, dim):
        super().__init__()
        self.norm1, self.norm2 = nn.LayerNorm(dim), nn.LayerNorm(dim)
        self.qkv = nn.Linear(dim, 3 * dim, bias=False)
        self.proj = nn.Linear(dim, dim)
        self.fc1, self.fc2 = nn.Linear(dim, 2 * dim), nn.Linear(2 * dim, dim)

    def forward(self, x):                      # [(B H W), T, C]
        q, k, v = self.qkv(self.norm1(x)).chunk(3, dim=-1)
        a = torch.softmax(q @ k.transpose(1, 2) * q.shape[-1] ** -0.5, dim=-1)
        x = x + self.proj(a @ v)
        return x + self.fc2(F.gelu(self.fc1(self.norm2(x))))


class UpdateBlock(nn.Module):
    def __init__(self, aggregator, T, dim=128):
        super().__init__()
        self.encoder = MotionEncoder()
        self.aggregator = aggregator
        self.gru = PCBlock(dim * 5, dim, (1, 7))
        self.mask = nn.Sequential(nn.Conv2d(dim, dim * 2, 3, padding=1), nn.ReLU(inplace=True),
                                  nn.Conv2d(dim * 2, 64 * 9, 1))
        self.temporal = TemporalBlock(dim)
        self.flow_head = PCBlock(dim * (T - 1), 2 * (T - 1), (1, 15))

    def forward(self, nets, inps, corrs, flows, attentions, T):
        BT, _, H, W = nets.shape
        B = BT // T
        mf = self.encoder(flows, corrs)
        mf_global = self.aggregator(attentions, mf)
        tok = mf.view(B, T, -1, H, W).permute(0, 3, 4, 1, 2).reshape(B * H * W, T, -1)
        mf_temporal = self.temporal(tok).view(B, H, W, T, -1).permute(0, 3, 4, 1, 2).reshape(BT, -1, H, W)
        nets = self.gru(torch.cat([nets, inps, mf, mf_global, mf_temporal.to(mf.dtype)], dim=1))
        delta = self.flow_head(nets.view(B, T * nets.shape[1], H, W))
        masks = 0.25 * self.mask(nets)
        return nets, masks.view(B, T, -1, H, W), delta.view(B, T, 2, H, W)


class StandInEncoder(nn.Module):
    """[B, T, 3, H, W] -> [B, T, 256, H/8, W/8]; channels-last output like the Twins encoder's token layout."""

    def __init__(self, norm):
        super().__init__()
        n = (lambda c: nn.InstanceNorm2d(c)) if norm == "instance" else (lambda c: nn.BatchNorm2d(c))
        self.net = nn.Sequential(nn.Conv2d(3, 64, 7, stride=2, padding=3), n(64), nn.ReLU(),
                                 nn.Conv2d(64, 128, 3, stride=2, padding=1), n(128), nn.ReLU(),
                                 nn.Conv2d(128, 192, 3, stride=2, padding=1), n(192), nn.ReLU(),
                                 nn.Conv2d(192, 256, 1))

    def forward(self, x):
        B, T = x.shape[:2]
        y = self.net(x.flatten(0, 1)).contiguous(memory_format=torch.channels_last)
        return y.view(B, T, *y.shape[1:])


def coords_grid(b, h, w, device):
    ys, xs = torch.meshgrid(torch.arange(h, device=device), torch.arange(w, device=device), indexing="ij")
    return torch.stack((xs, ys), 0).float()[None].repeat(b, 1, 1, 1)


class FlowModel(nn.Module):
    """SKFlow_MF8.forward with pluggable hot-path operators (corr_cls, attention module, aggregate module)."""

    def __init__(self, corr_cls, attention, aggregate, T=4):
        super().__init__()
        self.T = T
        self.fnet, self.cnet = StandInEncoder("instance"), StandInEncoder("batch")
        self.att = attention
        self.update_block = UpdateBlock(aggregate, T)
        self.corr_cls = corr_cls

    @staticmethod
    def upsample_flow(flow, mask, ratio=8):
        n, _, h, w = flow.shape
        mask = torch.softmax(mask.view(n, 1, 9, ratio, ratio, h, w), dim=2)
        up = F.unfold(ratio * flow, [3, 3], padding=1).view(n, 2, 9, 1, 1, h, w)
        up = torch.sum(mask * up, dim=2).permute(0, 1, 4, 2, 5, 3)
        return up.reshape(n, 2, ratio * h, ratio * w)

    @torch.no_grad()
    def forward(self, images, iters=12, mixed_precision=True):
        T = len(images)
        x = torch.stack(images, dim=1)
        B, _, _, H, W = x.shape
        x = 2 * (x / 255.0) - 1.0
        ac = dict(device_type="cuda", dtype=torch.float16, enabled=mixed_precision)
        with torch.autocast(**ac):
            fmaps = self.fnet(x).float()
            cnets = self.cnet(x[:, :-1])
        corr_fns = [self.corr_cls(fmaps[:, i], fmaps[:, i + 1], radius=4) for i in range(T - 1)]
        h, w = H // 8, W // 8
        coord0 = [coords_grid(B, h, w, x.device) for _ in range(T - 1)]
        coord1 = [c.clone() for c in coord0]
        with torch.autocast(**ac):
            nets, inps = torch.split(cnets, [128, 128], dim=2)
            nets = torch.tanh(nets.flatten(0, 1))
            inps = torch.relu(inps).flatten(0, 1)
            attentions = self.att(inps)
        masks = None
        for _ in range(iters):
            corrs = torch.stack([corr_fns[i](coord1[i]) for i in range(T - 1)], dim=1).flatten(0, 1)
            flows = torch.stack([coord1[i] - coord0[i] for i in range(T - 1)], dim=1).flatten(0, 1)
            with torch.autocast(**ac):
                nets, masks, delta = self.update_block(nets, inps, corrs, flows, attentions, T - 1)
            coord1 = [coord1[i] + delta[:, i].float() for i in range(T - 1)]
        low = [coord1[i] - coord0[i] for i in range(T - 1)]
        up = [self.upsample_flow(low[i], masks[:, i].float()) for i in range(T - 1)]
        return up, low


def randomise(model, seed=0, flow_gain=0.05):
    """Random init everywhere (incl. the blocks the reference zero-initialises); small flow-head gain keeps the
    12-iteration flow in a realistic +-20 px range instead of running off the image."""
    g = torch.Generator().manual_seed(seed)
    for name, p in model.named_parameters():
        if p.dim() > 1:
            fan_in = p[0].numel()
            p.data.copy_(torch.randn(p.shape, generator=g) * (1.0 / fan_in) ** 0.5)
        elif "gamma" in name:
            p.data.copy_(0.5 + torch.rand(p.shape, generator=g))
        elif name.endswith("weight"):                       # norm scales
            p.data.fill_(1.0)
        else:
            p.data.copy_(torch.randn(p.shape, generator=g) * 0.02)
    last = model.update_block.flow_head.ffn2[2]
    last.weight.data.mul_(flow_gain)
    last.bias.data.mul_(flow_gain)
    return model


def synthetic_clip(T, H, W, seed=0, device="cuda"):
    """Smooth random texture translated by a few pixels per frame (so the correlation volume has real structure)."""
    g = torch.Generator().manual_seed(seed)
    base = torch.rand(1, 3, H // 4 + 16, W // 4 + 16, generator=g)
    base = F.interpolate(base, scale_factor=4, mode="bicubic", align_corners=False)
    base = (base - base.min()) / (base.max() - base.min()) * 255.0
    frames = []
    for t in range(T):
        dx, dy = 3 * t + 2, 2 * t + 1
        frames.append(base[:, :, 16 + dy:16 + dy + H, 16 + dx:16 + dx + W].contiguous().to(device))
    return frames
