"""Clip / streaming-window sharding across the GPUs of one box with ONE collective: the gather of output flows.

The hot path has no exchange step between clips: every T-frame clip or window is independent
(``demo.py:518-532`` passes no ``flow_init``), so ranks take contiguous blocks of windows, run the whole model
on their own GPU and only the resulting flow fields cross NVLink (``all_gather`` over NCCL; gloo in CPU tests).

Window semantics follow the reference's streaming loop exactly (``demo.py:515-532``,
``core/mf_datasets.py:1125-1150``): windows of T frames advance by T-1 (consecutive windows share one frame);
if the last window would run past the sequence it is re-anchored to the final T frames and the flows that an
earlier window already produced are dropped.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Callable, List, Sequence

import torch
import torch.distributed as dist


@dataclass(frozen=True)
class Window:
    start: int                 # index of the first frame
    keep: tuple                # keep[k] -> emit the flow of pair (start+k, start+k+1)

    def frames(self, T: int) -> range:
        return range(self.start, self.start + T)


def window_schedule(n_frames: int, T: int = 4) -> List[Window]:
    """Windows covering ``n_frames`` frames; concatenating the kept flows gives exactly n_frames-1 flows."""
    if T < 2:
        raise ValueError("T must be at least 2")
    if n_frames < T:
        raise ValueError(f"need at least T={T} frames, got {n_frames}")
    out, i = [], 0
    while True:
        if i + T <= n_frames:
            out.append(Window(i, tuple(True for _ in range(T - 1))))
        else:
            s = n_frames - T
            out.append(Window(s, tuple((s + k) >= i for k in range(T - 1))))
        if i + T >= n_frames:
            break
        i += T - 1
    return out


def partition(n_items: int, world: int) -> List[range]:
    """Contiguous block partition; the first ``n_items % world`` ranks get one extra item."""
    base, extra = divmod(n_items, world)
    out, lo = [], 0
    for r in range(world):
        n = base + (1 if r < extra else 0)
        out.append(range(lo, lo + n))
        lo += n
    return out


def gather_flows(local: torch.Tensor, counts: Sequence[int], group=None, timings: dict | None = None) -> torch.Tensor:
    """All-gather per-rank flow stacks ``[n_r, 2, H, W]`` (n_r = counts[rank]) into ``[sum(counts), 2, H, W]``.

    One ``all_gather`` of equally sized (padded) buffers -- the only collective on the path.  ``timings`` (optional
    dict) receives ``gather_events`` = a (start, stop) pair of CUDA events bracketing the collective on the current
    stream and ``gather_bytes`` = the bytes every rank receives."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return local
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    if len(counts) != world or local.shape[0] != counts[rank]:
        raise ValueError("counts must list every rank's number of flows")
    cap = max(counts)       # `counts` is identical on every rank, so all ranks take the same branch
    if cap == 0:
        return local
    padded = local.new_zeros((cap,) + tuple(local.shape[1:]))
    padded[: local.shape[0]] = local
    # one output tensor, rank-major: all_gather_into_tensor is a single NCCL all-gather (no per-rank list copies)
    out = padded.new_empty((world * cap,) + tuple(local.shape[1:]))
    ev = None
    if timings is not None and local.is_cuda:
        ev = (torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True))
        ev[0].record()
    dist.all_gather_into_tensor(out, padded.contiguous(), group=group)
    if ev is not None:
        ev[1].record()
        timings["gather_events"] = ev
        timings["gather_bytes"] = out.numel() * out.element_size()
    if all(c == cap for c in counts):
        return out
    return torch.cat([out[r * cap: r * cap + c] for r, c in enumerate(counts)], dim=0)


def run_windows(frames: Sequence[torch.Tensor], flow_fn: Callable[[List[torch.Tensor]], List[torch.Tensor]],
                T: int = 4, group=None, timings: dict | None = None) -> torch.Tensor:
    """Streaming inference over a frame sequence, windows sharded over the ranks of ``group``.

    ``flow_fn(window_frames)`` must return the T-1 flows ``[2, H, W]`` of one window (e.g. the StreamFlow model in
    test mode).  Returns all ``len(frames) - 1`` flows, in temporal order, on every rank."""
    sched = window_schedule(len(frames), T)
    world = dist.get_world_size(group) if dist.is_available() and dist.is_initialized() else 1
    rank = dist.get_rank(group) if world > 1 else 0
    parts = partition(len(sched), world)
    counts = [sum(sum(sched[i].keep) for i in parts[r]) for r in range(world)]
    mine: List[torch.Tensor] = []
    for i in parts[rank]:
        w = sched[i]
        flows = flow_fn([frames[j] for j in w.frames(T)])
        if len(flows) != T - 1:
            raise ValueError(f"flow_fn returned {len(flows)} flows for a window of {T} frames")
        mine += [f for f, k in zip(flows, w.keep) if k]
    if mine:
        local = torch.stack(mine, 0)
    else:
        ref = frames[0]
        local = ref.new_zeros((0, 2) + tuple(ref.shape[-2:]))
    if timings is not None:
        timings["windows_per_rank"] = [len(p) for p in parts]
    return gather_flows(local, counts, group, timings)


def run_clips(clips: Sequence, flow_fn: Callable, group=None, timings: dict | None = None) -> torch.Tensor:
    """Independent clips sharded over ranks; ``flow_fn(clip) -> Tensor[T-1, 2, H, W]``; gathered in clip order."""
    world = dist.get_world_size(group) if dist.is_available() and dist.is_initialized() else 1
    rank = dist.get_rank(group) if world > 1 else 0
    # validated BEFORE any collective, from arguments every rank shares, so all ranks raise identically (a rank
    # that bailed out after its peers entered a collective would leave them hanging)
    if len(clips) < world:
        raise ValueError(f"run_clips: {len(clips)} clips for {world} ranks (every rank needs at least one clip)")
    parts = partition(len(clips), world)
    outs = [flow_fn(clips[i]) for i in parts[rank]]
    per = outs[0].shape[0]
    if any(o.shape[0] != per for o in outs):
        raise ValueError("run_clips: flow_fn must return the same number of flows for every clip")
    local = torch.cat(outs, 0)
    counts = [len(parts[r]) * per for r in range(world)]
    if timings is not None:
        timings["clips_per_rank"] = [len(p) for p in parts]
    return gather_flows(local, counts, group, timings)
