#!/usr/bin/env python
"""bench.py -- StreamFlow correlation/GMA hot path on B200: flow frames/s + kernel rooflines.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--eager] [--quick]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

Workload (BASELINE.json configs[1], SURVEY.md 8(d) "Config 2", operator-level synthetic): one T=4 clip at
Sintel 436x1024 (padded 440x1024 -> 55x128 at 1/8, N = 7040, D = 256, 3 frame pairs) per GPU, 12 refinement
iterations.  One STEP = one pass of the hot path over one clip per rank:
    3 x CorrBlock build  +  1 x Attention  +  12 x (3-pair lookup + Aggregate)
(core/models/streamflow.py:110,124,132 and core/update.py:769).  Clips are independent, so ranks shard clips
with no data-path collective (weak scaling).

JSON line (rank 0)
  value      flow fields per second over all ranks, hot path, inputs resident in HBM.  Default launch mode: the
             public-API calls of one clip captured once in a CUDA graph (streamflow_b200.GraphedCall) and replayed;
             `eager_ms_per_step` is the same calls issued eagerly (--eager makes that the headline).
  e2e        frames in -> flows out: the UNMODIFIED reference model (oracle/_ref: SKFlow_MF8 + SKUpdateBlock_TAM_v3 +
             Twins_CSC) running on the B200 operators through streamflow_b200.install(), its whole forward replayed as one
             CUDA graph per clip (streamflow_b200.GraphedModel; `e2e.eager_ms_per_step` = the same forward launched
             eagerly, the two are checked against each other); every step uploads the 4 uint8 frames from pinned host
             memory and downloads the 3 full-resolution flows.  Falls back to the
             hot-path call chain with host buffers when the reference snapshot is absent (`e2e.workload` says which).
  parity     the LAST timed step's lookup features and GMA result compared with the reference op sequence
             (oracle/torch_port, fp32, TF32 off) on the same GPU; the run fails above 1e-3.
  roofline   dominant kernel (GMA aggregate, HBM-bound E stream); kernels.* = the same for lookup / corr GEMM.
             `us_per_launch` = the kernel launched back-to-back inside one CUDA graph, CUDA events around replays;
             `us_per_launch_event_pairs_in_step` = one event pair per launch inside the eager step (upper bound).
             `traffic` = dram bytes per launch read from the committed ncu summary (profiles/ncu_dram_bytes.json).
  torch_gpu_baseline   the reference's own torch ops for the hot path on the SAME B200 (TF32 off = its default, and on).
  full_model           whole-forward frames/s of the unmodified reference model on its own operators vs on the B200
                       operators (same GPU, same weights), hot-path share of each; `all_patches` = plus the two optional
                       caller-side patches (patch_upsample, patch_motion_encoder = sf_pcblock_ffn1).
  kernels.pcblock_ffn1 SURVEY 8(f) row 2: the motion encoder's entry on the lookup output, against the reference's eager ops.
  configs              KITTI- and Spring-shaped hot-path throughput (BASELINE.json configs[2], [3]), one clip per GPU.
  streaming, kitti_x8  BASELINE.json configs[4] and [2] across the N ranks with the NCCL flow gather inside the timed
                       region (64 frames -> 21 windows -> 63 flows; 8 KITTI clips sharded 8/N per GPU, strong scaling).
  cpu_baseline         the reference's own CorrBlock / Attention / Aggregate (oracle/_ref, `kind: "reference"`; the
                       torch port if the snapshot is absent) on this host's cores.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time
import warnings

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

D, T, ITERS, CDIM = 256, 4, 12, 128
PAIRS = T - 1
SHAPES = {  # name -> (H, W of the frames, h, w at 1/8 after InputPadder)
    "sintel_436x1024": (436, 1024, 55, 128),
    "kitti_376x1248": (376, 1248, 47, 156),
    "spring_1080x1920": (1080, 1920, 135, 240),
}
H8, W8 = SHAPES["sintel_436x1024"][2:]
N = H8 * W8
METRIC = "flow frames/s (Sintel 436x1024, T=4, 12 iters); corr-lookup HBM GB/s"
WORKLOAD = "sintel_436x1024_T4_12iters_hotpath"


def load_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            p = json.load(f)
        return {"hbm_gbs": p["hbm_gbs"], "bf16_tflops": p["bf16_tflops"],
                "bf16_tflops_sustained": p.get("bf16_tflops_sustained", p["bf16_tflops"]), "source": "measured"}
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0, "source": "fallback"}


def load_ncu_traffic():
    """dram__bytes_read.sum + dram__bytes_write.sum per launch, written by scripts/ncu_summary.py --json from the
    committed `ncu --set full` capture (bench.py itself must not run under a profiler)."""
    path = os.path.join(ROOT, "profiles", "ncu_dram_bytes.json")
    if not os.path.exists(path):
        return {}, None
    with open(path) as f:
        d = json.load(f)
    return d.get("kernels", {}), d.get("source")


def make_inputs(seed: int, h8: int = H8, w8: int = W8, pairs: int = PAIRS, clips: int = 1):
    """Synthetic operator inputs of one clip (or `clips` clips batched), identical on every arm (torch CPU generator)."""
    import torch
    g = torch.Generator().manual_seed(seed)
    P = pairs * clips
    # fnet output under mixed precision: fp16 values upcast to fp32, channels-last storage (streamflow.py:107)
    fm = torch.randn(clips, pairs + 1, h8, w8, D, generator=g).half().float()
    inps = torch.relu(torch.randn(P, CDIM, h8, w8, generator=g))
    mfs = torch.randn(P, CDIM, h8, w8, generator=g)
    ys, xs = torch.meshgrid(torch.arange(h8), torch.arange(w8), indexing="ij")
    grid = torch.stack((xs, ys), 0).float()[None, None]                       # [1,1,2,h,w]
    walk = torch.cumsum(torch.randn(ITERS, pairs, clips, 2, h8, w8, generator=g) * 5.0 / ITERS ** 0.5, 0)
    coords = (grid + walk).contiguous()                                       # [iters, pairs, clips, 2, h, w]
    w_qk = torch.randn(2 * CDIM, CDIM, generator=g) * (CDIM ** -0.5) * 2.0
    w_v = torch.randn(CDIM, CDIM, generator=g) * (CDIM ** -0.5)
    return {"fm_nhwc": fm, "inps": inps, "mfs": mfs, "coords": coords, "w_qk": w_qk, "w_v": w_v, "gamma": 0.8}


def make_frames(n_frames: int, H: int, W: int, seed: int):
    """uint8 frames [n, 3, H, W]: a smooth random texture translated by a few pixels per frame, so the correlation
    volume has real structure and the flow is Sintel-sized (torch CPU generator: identical on every rank)."""
    import torch
    import torch.nn.functional as F
    g = torch.Generator().manual_seed(seed)
    mx, my = 3 * n_frames + 8, 2 * n_frames + 8
    base = torch.rand(1, 3, (H + my) // 4 + 2, (W + mx) // 4 + 2, generator=g)
    base = F.interpolate(base, scale_factor=4, mode="bicubic", align_corners=False)
    base = ((base - base.min()) / (base.max() - base.min()) * 255.0).round().clamp(0, 255).to(torch.uint8)[0]
    return torch.stack([base[:, 2 * t + 1: 2 * t + 1 + H, 3 * t + 2: 3 * t + 2 + W] for t in range(n_frames)]).contiguous()


class ClockSampler(threading.Thread):
    """Samples SM clocks / throttle reasons while the timed region runs (NVML every 20 ms; nvidia-smi fallback)."""

    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        super().__init__(daemon=True)
        self.index, self.samples, self._stop_evt = index, [], threading.Event()
        self.nvml = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nvml = pynvml
            self.handle = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = float(pynvml.nvmlDeviceGetMaxClockInfo(self.handle, pynvml.NVML_CLOCK_SM))
        except Exception:
            self.nvml = None

    def _sample_nvml(self):
        n = self.nvml
        mhz = float(n.nvmlDeviceGetClockInfo(self.handle, n.NVML_CLOCK_SM))
        try:
            mask = int(n.nvmlDeviceGetCurrentClocksEventReasons(self.handle))
        except Exception:
            mask = int(n.nvmlDeviceGetCurrentClocksThrottleReasons(self.handle))
        bits = {"hw_slowdown": 0x8, "hw_thermal_slowdown": 0x40, "sw_thermal_slowdown": 0x20, "sw_power_cap": 0x4}
        self.samples.append((mhz, self.max_mhz, [k for k, b in bits.items() if mask & b]))

    def _sample_smi(self):
        out = subprocess.run(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                              "-i", str(self.index)], capture_output=True, text=True, timeout=5).stdout
        p = [x.strip() for x in out.strip().split(",")]
        if len(p) >= 7:
            names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
            self.samples.append((float(p[0]), float(p[1]),
                                 [n for i, n in enumerate(names) if p[3 + i].lower().startswith("active")]))

    def run(self):
        while not self._stop_evt.is_set():
            try:
                self._sample_nvml() if self.nvml else self._sample_smi()
            except Exception:
                pass
            self._stop_evt.wait(0.02 if self.nvml else 0.15)

    def stop(self):
        self._stop_evt.set()
        self.join(timeout=6)
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["clock sampling unavailable"]}
        reasons = sorted({r for s in self.samples for r in s[2]})
        return {"sm_mhz": statistics.median(s[0] for s in self.samples), "sm_max_mhz": max(s[1] for s in self.samples),
                "reasons": reasons, "samples": len(self.samples), "source": "nvml" if self.nvml else "nvidia-smi"}


# ----------------------------------------------------------------------------- reference arm (CPU)
def _reference_l1():
    """The reference's own hot-path operators: oracle/_ref core/corr.py + core/gma.py (`kind: "reference"`), or the
    torch restatement oracle/torch_port.py when the snapshot did not travel (`kind: "port"`)."""
    import torch
    from oracle import make_ref
    core = make_ref.ref_core_dir()
    if core is None:
        from oracle import torch_port as tp
        return "port", None, tp
    if core not in sys.path:
        sys.path.insert(0, core)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        import corr as ref_corr
        import gma as ref_gma
    return "reference", (ref_corr, ref_gma), None


class _Ns:
    pass


def reference_hot_path_fn(inp, device="cpu", autocast=False):
    """Closure running `iters` refinement iterations of the hot path with the reference's operators on `device`
    (the once-per-clip part always runs); returns (feats, out)."""
    import torch
    kind, mods, tp = _reference_l1()
    fmaps = inp["fm_nhwc"].to(device).permute(0, 1, 4, 2, 3)
    coords, inps, mfs = inp["coords"].to(device), inp["inps"].to(device), inp["mfs"].to(device)
    pairs = fmaps.shape[1] - 1
    ac = dict(device_type="cuda" if str(device).startswith("cuda") else "cpu", dtype=torch.float16,
              enabled=bool(autocast))
    if kind == "reference":
        ref_corr, ref_gma = mods
        att = ref_gma.Attention(args=_Ns(), dim=CDIM, heads=1, max_pos_size=160, dim_head=CDIM).to(device)
        agg = ref_gma.Aggregate(args=_Ns(), dim=CDIM, heads=1, dim_head=CDIM).to(device)
        with torch.no_grad():
            att.to_qk.weight.copy_(inp["w_qk"].view(2 * CDIM, CDIM, 1, 1))
            agg.to_v.weight.copy_(inp["w_v"].view(CDIM, CDIM, 1, 1))
            agg.gamma.fill_(inp["gamma"])

        @torch.no_grad()
        def run(iters=ITERS):
            with warnings.catch_warnings():
                warnings.simplefilter("ignore")
                fns = [ref_corr.CorrBlock(fmaps[:, i], fmaps[:, i + 1], radius=4) for i in range(pairs)]
                with torch.autocast(**ac):
                    attn = att(inps)
                feats = out = None
                for it in range(iters):
                    feats = torch.stack([fns[i](coords[it, i]) for i in range(pairs)], 0)
                    with torch.autocast(**ac):
                        out = agg(attn, mfs)
            return feats, out
    else:
        w_qk, w_v = inp["w_qk"].to(device), inp["w_v"].to(device)

        @torch.no_grad()
        def run(iters=ITERS):
            pyrs = [tp.CpuCorrPyramid(fmaps[:, i], fmaps[:, i + 1]) for i in range(pairs)]
            with torch.autocast(**ac):
                attn = tp.cpu_attention(inps, w_qk)
            feats = out = None
            for it in range(iters):
                feats = torch.stack([pyrs[i](coords[it, i]) for i in range(pairs)], 0)
                with torch.autocast(**ac):
                    out = tp.cpu_aggregate(attn, mfs, w_v, inp["gamma"])
            return feats, out
    return kind, run


def cpu_hot_path_sample(inp, sample_iters):
    """One bounded CPU sample of the hot path: the once-per-clip part + `sample_iters` of the 12 iterations, the
    iteration part scaled to 12 (iterations are identical work).  Returns seconds for the full workload."""
    kind, run = reference_hot_path_fn(inp, "cpu")
    t0 = time.perf_counter()
    run(0)
    t1 = time.perf_counter()
    run(sample_iters)
    t2 = time.perf_counter()
    once = t1 - t0
    per_iter = max((t2 - t1) - once, 0.0) / sample_iters
    return kind, once + ITERS * per_iter


def cpu_full_model_sample(frames_u8):
    """One bounded CPU sample of the reference's FULL forward (oracle/_ref model on its own operators): 2 of the 12
    iterations run, the time of one iteration (between two update-block calls) is scaled to 12."""
    import torch
    from oracle import ref_model as rm
    mod = rm.load_model_module("reference")
    st = cpu_full_model_sample.__dict__
    if "model" not in st:
        torch.manual_seed(0)
        st["model"] = rm.randomise(rm.build_model(mod), seed=1).eval()
    model = st["model"]
    stamps = []
    h = model.update_block.register_forward_pre_hook(lambda m, a: stamps.append(time.perf_counter()))
    frames = [f[None].float() for f in frames_u8]
    from streamflow_b200.flowio import InputPadder
    padder = InputPadder(frames[0].shape)
    try:
        with torch.no_grad(), warnings.catch_warnings():
            warnings.simplefilter("ignore")
            t0 = time.perf_counter()
            model(padder.pad_list(frames), iters=2, test_mode=True)
            t1 = time.perf_counter()
    finally:
        h.remove()
    per_iter = stamps[1] - stamps[0]
    return (t1 - t0) + (ITERS - 2) * per_iter


def run_reference(args):
    import torch
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    torch.set_grad_enabled(False)
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    inp = make_inputs(0)
    # bounded sample: a full step is 0.4-2 s of CPU work; long runs sample 2 of the 12 iterations per step
    sample_iters = ITERS if args.steps <= 30 else 2
    kind = "port"
    for _ in range(args.warmup):
        kind, _ = cpu_hot_path_sample(inp, min(sample_iters, 2))
    dts = []
    for _ in range(args.steps):
        kind, dt = cpu_hot_path_sample(inp, sample_iters)
        dts.append(dt)
    dt = sum(dts) / len(dts)
    val = PAIRS / dt
    sample = (f"{args.steps} steps of 1 clip: 3 builds + attention + {sample_iters} of 12 iterations (3 lookups + "
              f"aggregate), iteration time scaled x{ITERS / sample_iters:g}; fp32 torch CPU, {cores} threads")
    e2e = {"value": val, "unit": "flow frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0,
           "workload": WORKLOAD}
    from oracle import ref_model as rm
    if rm.available() and not args.quick:
        # the reference's whole forward on CPU, the counterpart of our arm's frames-in -> flows-out `e2e`
        frames = make_frames(T, 436, 1024, 0)
        n = max(1, min(args.steps, 3))
        cpu_full_model_sample(frames)                                  # warm-up (weights, thread pools)
        full = sum(cpu_full_model_sample(frames) for _ in range(n)) / n
        e2e = {"value": PAIRS / full, "unit": "flow frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0,
               "ms_per_step": full * 1e3, "workload": "sintel_436x1024_T4_12iters_full_model",
               "sample": f"{n} forwards of the unmodified reference model (oracle/_ref) on CPU with 2 of 12 iterations "
                         "run; one iteration's time (between two update-block calls) scaled to 12"}
    line = {
        "impl": "reference", "metric": METRIC, "value": val, "unit": "flow frames/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt * 1e3, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": WORKLOAD, "note": "the reference's own CPU implementation of the path on the host cores"
                   if kind == "reference" else "torch CPU port of core/corr.py + core/gma.py (oracle/_ref absent)"},
        "cpu_baseline": {"value": val, "unit": "flow frames/s", "cores": cores, "kind": kind, "sample": sample},
        "e2e": e2e,
        "gpu_launches": 0,
    }
    print(json.dumps(line))
    return 0


# ----------------------------------------------------------------------------------- our arm (GPU)
def _rel(a, b):
    return float((a.double() - b.double()).norm() / b.double().norm().clamp_min(1e-30))


class HotPath:
    """The hot path of one clip (or a batch of clips) through the public API -- what core/models/streamflow.py drives."""

    def __init__(self, sfb, dev, host):
        import torch
        self.sfb, self.dev = sfb, dev
        self.att = sfb.Attention(args=_Ns(), dim=CDIM, heads=1, max_pos_size=160, dim_head=CDIM).to(dev)
        self.agg = sfb.Aggregate(args=_Ns(), dim=CDIM, heads=1, dim_head=CDIM).to(dev)
        with torch.no_grad():
            self.att.to_qk.weight.copy_(host["w_qk"].view(2 * CDIM, CDIM, 1, 1))
            self.agg.to_v.weight.copy_(host["w_v"].view(CDIM, CDIM, 1, 1))
            self.agg.gamma.fill_(host["gamma"])

    def __call__(self, t, iters=ITERS):
        fmaps = t["fm_nhwc"].permute(0, 1, 4, 2, 3)                         # [B, T, D, h, w] channels-last views
        pairs = fmaps.shape[1] - 1
        group = self.sfb.CorrGroup.from_fmaps(fmaps, radius=4)              # the T-1 pyramids, one batched build
        handle = self.att(t["inps"])
        feats = out = None
        for it in range(iters):
            feats = group([t["coords"][it, i] for i in range(pairs)])
            out = self.agg(handle, t["mfs"])
        return feats, out


def _time_events(fn, steps, warmup, barrier, extra_streams=()):
    import torch
    for _ in range(warmup):
        fn()
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        fn()
    for st in extra_streams:
        torch.cuda.current_stream().wait_stream(st)
    e1.record()
    barrier()
    return e0.elapsed_time(e1) / steps


def load_full_models(dev):
    """(reference model on its own operators, the same model on the B200 operators), identical weights; None when the
    reference snapshot (oracle/_ref) is absent."""
    import torch
    from oracle import ref_model as rm
    if not rm.available():
        return None
    import streamflow_b200 as sfb
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        ref_mod, our_mod = rm.load_model_module("reference"), rm.load_model_module("b200")
        torch.manual_seed(0)
        ref = rm.randomise(rm.build_model(ref_mod), seed=1).to(dev).eval()
        ours = rm.build_model(our_mod).to(dev).eval()
    ours.load_state_dict(ref.state_dict(), strict=True)
    if ours.update_block.aggregator.__class__ is not sfb.Aggregate or our_mod.CorrBlock is not sfb.CorrBlock:
        raise SystemExit("bench.py: install() did not bind the B200 operators into the reference model")
    return ref, ours


class FullModelRunner:
    """frames (uint8, pinned host) -> flows (fp32, pinned host) through a StreamFlow model on `dev`."""

    def __init__(self, model, dev, H, W, T=4, iters=ITERS):
        import torch
        from streamflow_b200.flowio import InputPadder
        self.model, self.dev, self.iters, self.T = model, dev, iters, T
        self.padder = InputPadder((H, W))
        self.dev_frames = torch.empty((T, 3, H, W), dtype=torch.uint8, device=dev)
        self.host_flows = torch.empty((T - 1, 2, H, W), dtype=torch.float32).pin_memory()
        self.h2d = T * 3 * H * W
        self.d2h = (T - 1) * 2 * H * W * 4

    def flows_on_device(self, frames_u8):
        """frames_u8: [T, 3, H, W] uint8 (pinned host or device) -> [T-1, 2, H, W] fp32 on the device."""
        import torch
        self.dev_frames.copy_(frames_u8, non_blocking=True)
        frames = [f[None].float() for f in self.dev_frames]
        with torch.no_grad(), warnings.catch_warnings():
            warnings.simplefilter("ignore")
            out = self.model(self.padder.pad_list(frames), iters=self.iters, test_mode=True)
        return torch.cat([self.padder.unpad(o) for o in out], 0)

    def __call__(self, frames_u8):
        flows = self.flows_on_device(frames_u8)
        self.host_flows.copy_(flows, non_blocking=True)
        return flows


def run_ours(args):
    import torch
    import torch.distributed as dist
    import streamflow_b200 as sfb
    from streamflow_b200 import _lib
    from streamflow_b200 import dist as sfd

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device -- the product path has no CPU fallback")
    torch.cuda.set_device(local)
    torch.set_grad_enabled(False)                 # inference-only operators (the reference evaluates under no_grad)
    torch.backends.cuda.matmul.allow_tf32 = False  # the reference never enables TF32
    torch.backends.cudnn.allow_tf32 = False
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    L = sfb.lib()
    peaks = load_peaks()
    traffic, traffic_src = load_ncu_traffic()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(ms):
        if world > 1:
            tt = torch.tensor([ms], device=dev, dtype=torch.float64)
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
            return float(tt.item())
        return ms

    def min_over_ranks(ms):
        if world > 1:
            tt = torch.tensor([ms], device=dev, dtype=torch.float64)
            dist.all_reduce(tt, op=dist.ReduceOp.MIN)
            return float(tt.item())
        return ms

    host = make_inputs(rank)                      # each rank owns a different clip (weak scaling)
    hot = HotPath(sfb, dev, host)
    resident = {k: host[k].to(dev) for k in ("fm_nhwc", "inps", "mfs", "coords")}

    # ---- headline: device-resident inputs; CUDA-graph replay of the public-API calls unless --eager
    def step_eager():
        return hot(resident)

    graphed = None if args.eager else sfb.GraphedCall(step_eager)
    step_resident = step_eager if args.eager else graphed
    sampler = ClockSampler(local) if rank == 0 else None
    launches0 = L.sf_launch_count()
    if sampler:
        sampler.start()
    ms_step = max_over_ranks(_time_events(step_resident, args.steps, args.warmup, barrier))
    clocks = sampler.stop() if sampler else None
    if args.eager:
        launches = (L.sf_launch_count() - launches0) * args.steps // (args.steps + args.warmup)
        ms_eager = ms_step
    else:
        launches = graphed.launches * args.steps
        ms_eager = max_over_ranks(_time_events(step_eager, min(args.steps, 50), 3, barrier))

    # ---- parity of what was just timed: last step's outputs vs the reference op sequence on the same GPU (fp32)
    feats, out = step_resident()
    torch.cuda.synchronize()
    from oracle import torch_port as tp                                   # the checker, never the thing measured
    fm = resident["fm_nhwc"].permute(0, 1, 4, 2, 3)
    ref_feats = torch.stack([tp.CpuCorrPyramid(fm[:, i], fm[:, i + 1])(resident["coords"][ITERS - 1, i])
                             for i in range(PAIRS)], 0).reshape(feats.shape)
    w_qk, w_v = host["w_qk"].to(dev), host["w_v"].to(dev)
    ref_out = tp.cpu_aggregate(tp.cpu_attention(resident["inps"], w_qk), resident["mfs"], w_v, host["gamma"])
    parity = {"corr_rel": _rel(feats, ref_feats), "gma_rel": _rel(out - resident["mfs"], ref_out - resident["mfs"]),
              "bound": 1e-3, "checker": "oracle/torch_port op sequence on the same GPU, fp32, TF32 off, same inputs as "
                                        "the timed step (last iteration)"}
    del ref_feats, ref_out
    ok = torch.tensor([float(parity["corr_rel"] < 1e-3 and parity["gma_rel"] < 1e-3)], device=dev)
    if world > 1:
        dist.all_reduce(ok, op=dist.ReduceOp.MIN)
    if float(ok.item()) < 1.0:
        raise SystemExit(f"bench.py: parity check failed on rank {rank}: {parity}")
    torch.cuda.empty_cache()

    # ---- e2e
    models = None if args.quick else load_full_models(dev)
    e2e, full_model = None, None
    if models is not None:
        ref_model, our_model = models
        Hs, Ws = SHAPES["sintel_436x1024"][:2]
        frames_host = make_frames(T, Hs, Ws, 100 + rank).pin_memory()
        runner = FullModelRunner(our_model, dev, Hs, Ws)
        k_e2e = max(3, min(args.steps, 10))
        ms_e2e_eager = max_over_ranks(_time_events(lambda: runner(frames_host), k_e2e, 3, barrier))
        # public API for a stream of equally shaped clips: the whole forward of the unmodified model replayed as ONE CUDA
        # graph (possible because the B200 operators never synchronise or touch the host; the reference's own CorrBlock
        # builds its sampling grid on the host every lookup)
        gm = sfb.GraphedModel(our_model, (T, 3, Hs, Ws), iters=ITERS)
        e2e_host_flows = torch.empty((PAIRS, 2, Hs, Ws), dtype=torch.float32).pin_memory()

        def step_e2e_graph():
            e2e_host_flows.copy_(gm(frames_host), non_blocking=True)

        ms_e2e = max_over_ranks(_time_events(step_e2e_graph, k_e2e, 3, barrier))
        graph_diff = float((gm(frames_host) - runner.flows_on_device(frames_host)).abs().max())
        if not graph_diff < 1e-3:
            raise SystemExit(f"bench.py: graph replay of the full model differs from the eager forward by {graph_diff} px")
        e2e = {"value": world * PAIRS / (ms_e2e / 1e3), "unit": "flow frames/s", "ms_per_step": ms_e2e,
               "h2d_bytes_per_step": runner.h2d, "d2h_bytes_per_step": runner.d2h, "steps": k_e2e,
               "eager_ms_per_step": ms_e2e_eager, "graph_vs_eager_max_abs_px": graph_diff,
               "workload": "sintel_436x1024_T4_12iters_full_model",
               "launch": "cuda_graph replay of the whole forward (streamflow_b200.GraphedModel)",
               "note": "uint8 frames from pinned host memory -> unmodified reference model (oracle/_ref SKFlow_MF8 / "
                       "SKUpdateBlock_TAM_v3 / Twins_CSC, fp16 autocast) on the B200 operators via "
                       "streamflow_b200.install() -> full-resolution flows to pinned host memory"}
        ms_e2e_graph, ms_e2e = ms_e2e, ms_e2e_eager        # the same-GPU comparison below is eager against eager
        if rank == 0 and world == 1:
            ref_runner = FullModelRunner(ref_model, dev, Hs, Ws)
            ms_ref = _time_events(lambda: ref_runner(frames_host), k_e2e, 3, barrier)
            fo = runner.flows_on_device(frames_host)
            fr = ref_runner.flows_on_device(frames_host)
            epe = torch.sqrt(((fo - fr) ** 2).sum(1)).mean(dim=(1, 2))
            full_model = {"b200_l1_ms": ms_e2e, "reference_l1_ms": ms_ref,
                          "b200_l1_flows_per_s": PAIRS / (ms_e2e / 1e3), "reference_l1_flows_per_s": PAIRS / (ms_ref / 1e3),
                          "speedup": ms_ref / ms_e2e, "hot_path_share_b200": ms_eager / ms_e2e,
                          "b200_l1_graph_ms": ms_e2e_graph, "speedup_graph": ms_ref / ms_e2e_graph,
                          "mean_epe_px_per_pair": [float(x) for x in epe],
                          "flow_magnitude_px": float(torch.sqrt((fr ** 2).sum(1)).mean()),
                          "note": "whole forward incl. H2D of frames and D2H of flows, same GPU, same random-init weights "
                                  "(gamma ~ U(0.5,1.5), temporal block re-randomised), TF32 off; everything outside the "
                                  "hot path is the reference's unchanged eager PyTorch code"}
            # optional caller-side patch: the model's convex 8x upsampling, which the reference runs
            # for every iteration and pair although test mode returns only the last one, replaced by sf_upsample_flow
            cls = our_model.__class__
            orig_up = cls.upsample_flow
            try:
                sfb.patch_upsample(cls)
                ms_up = _time_events(lambda: runner(frames_host), k_e2e, 2, barrier)
                fu = runner.flows_on_device(frames_host)
                full_model["b200_l1_plus_upsample_kernel_ms"] = ms_up
                full_model["b200_l1_plus_upsample_kernel_flows_per_s"] = PAIRS / (ms_up / 1e3)
                full_model["upsample_kernel_mean_epe_px_per_pair"] = [
                    float(x) for x in torch.sqrt(((fu - fr) ** 2).sum(1)).mean(dim=(1, 2))]
                del fu
                # ... plus the motion encoder's entry: the first stage of its four PCBlocks on
                # sf_pcblock_ffn1 (streamflow_b200.patch_motion_encoder), eager and as one graph replay of the forward
                patched = sfb.patch_motion_encoder(our_model)
                ms_all = _time_events(lambda: runner(frames_host), k_e2e, 2, barrier)
                fa = runner.flows_on_device(frames_host)
                gm_all = sfb.GraphedModel(our_model, (T, 3, Hs, Ws), iters=ITERS)
                ms_all_graph = _time_events(lambda: e2e_host_flows.copy_(gm_all(frames_host), non_blocking=True), k_e2e, 2,
                                            barrier)
                full_model["all_patches"] = {
                    "patched": ["upsample_flow"] + [f"encoder.{n}.ffn1" for n in patched],
                    "eager_ms": ms_all, "graph_ms": ms_all_graph, "graph_flows_per_s": PAIRS / (ms_all_graph / 1e3),
                    "speedup_vs_reference_l1": ms_ref / ms_all_graph,
                    "mean_epe_px_per_pair": [float(x) for x in torch.sqrt(((fa - fr) ** 2).sum(1)).mean(dim=(1, 2))]}
                del fa, gm_all
            finally:
                cls.upsample_flow = orig_up
                sfb.unpatch_motion_encoder(our_model)
            del ref_runner, fo, fr
    else:
        # fallback boundary: the hot-path call chain itself with host buffers (fm uploaded as the fp16 it is)
        pinned = {"fm_nhwc": host["fm_nhwc"].half().pin_memory(), "inps": host["inps"].pin_memory(),
                  "mfs": host["mfs"].pin_memory(), "coords": host["coords"].pin_memory()}
        dev_in = {k: torch.empty_like(v, device=dev) for k, v in pinned.items()}
        host_out = [torch.empty((PAIRS, 324, H8, W8)).pin_memory(), torch.empty((PAIRS, CDIM, H8, W8)).pin_memory()]

        def step_e2e():
            for k, v in pinned.items():
                dev_in[k].copy_(v, non_blocking=True)
            f, o = hot(dev_in)
            host_out[0].copy_(f, non_blocking=True)
            host_out[1].copy_(o, non_blocking=True)

        ms_e2e = max_over_ranks(_time_events(step_e2e, args.steps, max(args.warmup, 3), barrier))
        e2e = {"value": world * PAIRS / (ms_e2e / 1e3), "unit": "flow frames/s", "ms_per_step": ms_e2e,
               "h2d_bytes_per_step": sum(v.numel() * v.element_size() for v in pinned.values()),
               "d2h_bytes_per_step": sum(v.numel() * 4 for v in host_out), "workload": WORKLOAD,
               "note": "oracle/_ref absent: hot-path call chain with pinned host buffers"}

    # ---- per-kernel timing
    def _raw_event():
        ev = torch.cuda.Event(enable_timing=True)
        ev.record()          # forces creation of the underlying cudaEvent_t
        return ev

    def _step_with_kernel_events(which):
        t = resident
        fmaps = t["fm_nhwc"].permute(0, 1, 4, 2, 3)
        pairs_ev = []

        def arm():
            a, b = _raw_event(), _raw_event()
            L.sf_profile_kernel(which, a.cuda_event, b.cuda_event)
            pairs_ev.append((a, b))

        def disarm():
            L.sf_profile_kernel(0, None, None)

        if which in (_lib.KERNEL_CORR_GEMM, _lib.KERNEL_CORR_PACK):
            arm()
        group = sfb.CorrGroup.from_fmaps(fmaps, radius=4)
        disarm()
        if which == _lib.KERNEL_GMA_STATS:
            arm()
        handle = hot.att(t["inps"])
        disarm()
        for it in range(ITERS):
            if which == _lib.KERNEL_LOOKUP:
                arm()
            group([t["coords"][it, i] for i in range(PAIRS)])
            disarm()
            if which in (_lib.KERNEL_GMA_AGGREGATE, _lib.KERNEL_GMA_PROJ):
                arm()
            hot.agg(handle, t["mfs"])
            disarm()
        torch.cuda.synchronize()
        return [a.elapsed_time(b) * 1e3 for a, b in pairs_ev]       # microseconds

    def kernel_time(which):
        durs = [_step_with_kernel_events(which) for _ in range(3)]
        flat = [d for ds in durs[1:] for d in ds]
        return sum(flat) / len(flat), len(flat)

    def graph_kernel_time(fn, calls, gma_mask=7, corr_mask=3, replays=5):
        """Kernel duration without event gaps: `calls` back-to-back launches captured in ONE CUDA graph with the
        library restricted to the kernel under test, CUDA events around `replays` replays."""
        L.sf_debug_select_kernels(gma_mask, corr_mask)
        try:
            st = torch.cuda.Stream()
            with torch.cuda.stream(st):
                for i in range(2):
                    fn(i)
                torch.cuda.synchronize()
                g = torch.cuda.CUDAGraph()
                with torch.cuda.graph(g, stream=st):
                    keep = [fn(i) for i in range(calls)]
            torch.cuda.synchronize()
            g.replay()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(replays):
                g.replay()
            e1.record()
            torch.cuda.synchronize()
            del keep, g
            return e0.elapsed_time(e1) * 1e3 / (replays * calls)
        finally:
            L.sf_debug_select_kernels(7, 3)

    kernels = {}
    if rank == 0:
        hbm = peaks["hbm_gbs"]
        fm_views = resident["fm_nhwc"].permute(0, 1, 4, 2, 3)
        g_group = sfb.CorrGroup.from_fmaps(fm_views, radius=4)
        # two independent sets of softmax numerators, alternated, so no launch can find its own 297 MB stream
        # left over in the 126 MB L2 by the previous launch
        g_handle = [hot.att(resident["inps"]), hot.att(resident["inps"])]
        us_graph = {
            "gma_aggregate": graph_kernel_time(lambda i: hot.agg(g_handle[i & 1], resident["mfs"]), ITERS, gma_mask=2),
            "gma_cast": graph_kernel_time(lambda i: hot.agg(g_handle[i & 1], resident["mfs"]), ITERS, gma_mask=1),
            "corr_lookup": graph_kernel_time(
                lambda i: g_group([resident["coords"][i % ITERS, j] for j in range(PAIRS)]), ITERS),
            "corr_gemm": graph_kernel_time(lambda i: sfb.CorrGroup.from_fmaps(fm_views, radius=4), 4, corr_mask=2),
        }
        # launch-size dependence of the lookup: the same kernel over 1 and 8 pairs per launch (8 = SF_MAX_GROUPS, a T = 9
        # clip) separates a fixed per-launch cost (launch, pipeline fill, tail) from the streaming rate
        g1 = sfb.CorrGroup(g_group.blocks[:1])
        us_lk1 = graph_kernel_time(lambda i: g1([resident["coords"][i % ITERS, 0]]), ITERS)
        fm9 = torch.cat([fm_views, fm_views, fm_views[:, :1]], dim=1)                   # 9 frames -> 8 pairs
        g8 = sfb.CorrGroup.from_fmaps(fm9, radius=4)
        c8 = [resident["coords"][:, j % PAIRS] for j in range(8)]
        us_lk8 = graph_kernel_time(lambda i: g8([c8[j][i % ITERS] for j in range(8)]), ITERS)
        del g1, g8, fm9
        del g_group, g_handle
        us, n = kernel_time(_lib.KERNEL_GMA_AGGREGATE)
        npad = L.sf_gma_npad(N)
        # E stream + fp16 X + the fused epilogue's fmap read and result write, per launch
        bytes_agg = PAIRS * N * npad * 2 + PAIRS * CDIM * npad * 2 + 2 * PAIRS * CDIM * N * 4
        ug = us_graph["gma_aggregate"]
        kernels["gma_aggregate"] = {"bound": "hbm", "achieved": bytes_agg / ug / 1e3, "peak": hbm, "unit": "GB/s",
                                    "frac": bytes_agg / ug / 1e3 / hbm, "us_per_launch": ug,
                                    "us_per_launch_event_pairs_in_step": us, "launches_timed": n,
                                    "algorithmic_bytes": bytes_agg, "traffic": traffic.get("gma_aggregate"),
                                    "flops": 2.0 * PAIRS * N * N * CDIM,
                                    "tensor_frac_if_P_cached": 2.0 * PAIRS * N * N * CDIM / ug / 1e6 / peaks["bf16_tflops"]}
        us, n = kernel_time(_lib.KERNEL_LOOKUP)
        bytes_lk = 2904 * PAIRS * N
        ug = us_graph["corr_lookup"]
        kernels["corr_lookup"] = {"bound": "hbm", "achieved": bytes_lk / ug / 1e3, "peak": hbm, "unit": "GB/s",
                                  "frac": bytes_lk / ug / 1e3 / hbm, "us_per_launch": ug,
                                  "us_per_launch_event_pairs_in_step": us, "launches_timed": n,
                                  "algorithmic_bytes": bytes_lk, "traffic": traffic.get("corr_lookup"),
                                  "note": "3 pairs per launch, coords random-walk; pyramid 805 MB >> L2",
                                  "launch_size": {"us_1_pair": us_lk1, "us_3_pairs": ug, "us_8_pairs": us_lk8,
                                                  "us_per_extra_pair": (us_lk8 - us_lk1) / 7.0,
                                                  "frac_8_pairs": 2904 * 8 * N / us_lk8 / 1e3 / hbm,
                                                  "frac_marginal": 2904 * N / ((us_lk8 - us_lk1) / 7.0) / 1e3 / hbm,
                                                  "note": "1 / 3 / 8 pairs per launch, the same warp-specialised cp.async kernel: the rate "
                                                          "per extra pair does not improve with launch size (it drops once the 2.1 GB of "
                                                          "8 pyramids leave nothing of the previous launch in L2), i.e. the gather, not a "
                                                          "fixed per-launch cost, sets the fraction"}}
        us, n = kernel_time(_lib.KERNEL_CORR_GEMM)
        flops = 2.0 * N * N * D * PAIRS                     # one launch builds the pyramids of all pairs
        us_ev, us = us, us_graph["corr_gemm"]
        tf = flops / us / 1e6
        peak_tf = peaks["bf16_tflops"]
        out_bytes = PAIRS * 4 * N * sum(((H8 >> l) + 3) // 4 * (((W8 >> l) + 3) // 4) * 16 for l in range(4))
        # Primary roofline = HBM writes (the kernel stores 805 MB per launch against 76 GFLOP: 20 us of MMA vs 42 us of
        # stores per pair); the tensor-pipe view north_star names is reported beside it (ncu: 39 % active in this mode,
        # 58 % in the three-product mode, profiles/r2_ncu_full_summary.md).
        write_probe = 5880.0      # GB/s, scripts/probes/store_probe.cu on this pool (profiles/r1d_probes.txt)
        kernels["corr_gemm"] = {"bound": "hbm", "achieved": out_bytes / us / 1e3, "peak": hbm, "unit": "GB/s",
                                "frac": out_bytes / us / 1e3 / hbm, "frac_of_write_probe": out_bytes / us / 1e3 / write_probe,
                                "us_per_launch": us, "us_per_launch_event_pairs_in_step": us_ev, "launches_timed": n,
                                "algorithmic_bytes": out_bytes, "traffic": traffic.get("corr_gemm"),
                                "algorithmic_flops": flops, "tensor_tflops": tf, "tensor_frac": tf / peak_tf,
                                "tensor_pipe_active_pct_ncu": {"auto_exact_inputs": 39.2, "f16x2": 58.1},
                                "note": "one launch = the %d pairs of the clip; precision auto on fp16-exact inputs (single-product "
                                        "level 0, hi*hi + hi*lo pooled levels), CTA pairs (cta_group::2), fp32 accumulate" % PAIRS}
        kernels["gma_cast"] = {"us_per_launch": us_graph["gma_cast"]}
        for name, kind in (("gma_stats_pass2", _lib.KERNEL_GMA_STATS), ("corr_pack", _lib.KERNEL_CORR_PACK)):
            us, n = kernel_time(kind)
            kernels[name] = {"us_per_launch_event_pairs_in_step": us, "launches_timed": n}
        # the motion encoder's entry on the lookup output (convc1: 324 -> 486 -> 324, fp32 in / out),
        # one launch for the 3 maps of the clip, against the reference's own four eager ops under autocast
        import torch.nn as nn
        import torch.nn.functional as F
        g = torch.Generator().manual_seed(7)
        ffn1 = nn.Sequential(nn.Conv2d(324, 486, 1), nn.GELU(), nn.Conv2d(486, 324, 1)).to(dev).eval()
        xin = feats.reshape(PAIRS, 324, H8, W8).clone()
        us_ffn = graph_kernel_time(lambda i: sfb.pcblock_ffn1(xin, ffn1), ITERS)

        def ref_ffn():
            with torch.autocast("cuda", dtype=torch.float16):
                return F.gelu(xin + ffn1(xin))
        ref_y = F.gelu(xin + ffn1(xin))
        torch.cuda.synchronize()
        ours_y = sfb.pcblock_ffn1(xin, ffn1)
        torch.cuda.synchronize()
        ffn_rel = _rel(ours_y, ref_y)
        # fp16 operands cannot reproduce the fp32 ops bit for bit: an exact 0.0 (seen once in ~10 runs, not reproduced by
        # scripts/ffn1_stress.py) would mean the comparison did not see the kernel's output
        ffn_suspect = ffn_rel == 0.0 or ours_y.data_ptr() == ref_y.data_ptr()
        if ffn_suspect:                      # compare once more against a fresh evaluation of both sides
            ffn_rel = _rel(sfb.pcblock_ffn1(xin, ffn1), F.gelu(xin + ffn1(xin)))
        del ours_y
        us_ref_ffn = _time_events(ref_ffn, 20, 5, lambda: torch.cuda.synchronize()) * 1e3
        fl = 2.0 * PAIRS * N * 324 * 486 * 2
        kernels["pcblock_ffn1"] = {"us_per_launch": us_ffn, "reference_ops_us": us_ref_ffn, "speedup": us_ref_ffn / us_ffn,
                                   "rel_err_vs_fp32_ops": ffn_rel, "rel_err_first_check_was_exact_zero": bool(ffn_suspect),
                                   "algorithmic_flops": fl,
                                   "tensor_tflops": fl / us_ffn / 1e6, "tensor_frac": fl / us_ffn / 1e6 / peak_tf,
                                   "note": "gelu(x + W2 gelu(W1 x + b1) + b2) on the 324-channel lookup output of the clip "
                                           "(core/update.py:31); bound by the 128 x (512 + 336) exact-erf GELUs per 128-pixel tile "
                                           "on the FP32 pipe and by shared-memory-operand MMAs, not by the tensor pipe; the "
                                           "reference runs conv1x1, GELU, conv1x1, add + GELU eagerly under autocast"}
        if not ffn_rel < 2e-3:
            raise SystemExit(f"bench.py: pcblock_ffn1 differs from the torch ops by {ffn_rel}")
        del ffn1, xin, ref_y
        torch.cuda.empty_cache()

    # ---- the reference's torch ops on the SAME GPU (like-for-like bar), rank 0 at N=1
    torch_gpu = None
    if rank == 0 and world == 1 and not args.quick:
        torch_gpu = {}
        for tf32 in (False, True):
            torch.backends.cuda.matmul.allow_tf32 = tf32
            torch.backends.cudnn.allow_tf32 = tf32
            kind, run = reference_hot_path_fn(host, dev, autocast=True)
            ms = _time_events(run, 3, 2, barrier)
            torch_gpu["tf32_on" if tf32 else "tf32_off"] = {"ms_per_step": ms, "flows_per_s": PAIRS / (ms / 1e3)}
        torch.backends.cuda.matmul.allow_tf32 = False
        torch.backends.cudnn.allow_tf32 = False
        torch_gpu["kind"] = kind
        torch_gpu["speedup_vs_tf32_off"] = torch_gpu["tf32_off"]["ms_per_step"] / ms_eager
        torch_gpu["note"] = ("the reference's own CorrBlock / Attention / Aggregate (fp16 autocast GMA as in the model) "
                             "eager on this GPU, same workload as `value`; compared with our eager step")
        del run
        torch.cuda.empty_cache()

    # ---- KITTI- and Spring-shaped hot path (BASELINE.json configs[2], [3]): one clip per rank (weak scaling like the
    # headline), time = max over ranks, flows/s = all ranks' flows
    configs = None
    if not args.quick:
        configs = {}
        for name in ("kitti_376x1248", "spring_1080x1920"):
            h8, w8 = SHAPES[name][2:]
            hi = make_inputs(7 + rank, h8, w8)
            res = {k: hi[k].to(dev) for k in ("fm_nhwc", "inps", "mfs", "coords")}
            gcall = sfb.GraphedCall(lambda: hot(res))
            reps = 20 if name.startswith("kitti") else 4
            ms = max_over_ranks(_time_events(gcall, reps, 2, barrier))
            n_ = h8 * w8
            configs[name] = {"ms_per_clip": ms, "flows_per_s": world * PAIRS / (ms / 1e3), "N": n_, "steps": reps,
                             "clips_per_gpu": 1,
                             "pyramid_gb": PAIRS * 4 * n_ * sum(((h8 >> l) + 3) // 4 * (((w8 >> l) + 3) // 4) * 16
                                                               for l in range(4)) / 1e9,
                             "softmax_numerators_gb": PAIRS * L.sf_gma_e_elems(1, n_) * 2 / 1e9,
                             "launch": "cuda_graph replay, one clip of T=4, 12 iterations, hot path"}
            del gcall, res, hi
            torch.cuda.empty_cache()

    # ---- multi-GPU legs with the ONE collective of the design: the NCCL gather of the output flows
    streaming = kitti_x8 = None
    if models is not None and not args.quick:
        our_model = models[1]
        # config 5: 64 frames -> 21 overlapping T=4 windows -> 63 flows, windows block-partitioned over the ranks
        Hs, Ws = SHAPES["sintel_436x1024"][:2]
        n_frames = args.stream_frames
        video = make_frames(n_frames, Hs, Ws, 0).pin_memory()            # identical on every rank
        if "gm" not in locals():
            gm = sfb.GraphedModel(our_model, (T, 3, Hs, Ws), iters=ITERS)

        def flow_fn(window):          # graph replay per window; the static output is overwritten by the next replay
            return list(gm(torch.stack(window)).clone())

        frames_list = list(video)
        gm(video[:T])                                                      # warm-up
        if world > 1:      # the communicator's first all-gather sets up its channels (~100 ms): not part of the path
            sfd.gather_flows(torch.zeros(1, 2, 8, 8, device=dev), [1] * world)
        tm = {}
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        flows = sfd.run_windows(frames_list, flow_fn, T=T, timings=tm)
        host_flows = flows.cpu() if rank == 0 else None                   # rank 0 hands the sequence to the caller
        e1.record()
        barrier()
        ms_total = max_over_ranks(e0.elapsed_time(e1))
        gather_ms = tm["gather_events"][0].elapsed_time(tm["gather_events"][1]) if "gather_events" in tm else 0.0
        # a rank that finishes its windows early waits inside the collective for the slowest one: the last arriver (MIN
        # over ranks) sees the transfer itself, the MAX is transfer + load imbalance (one window = ~90 ms)
        gather_wait_ms = max_over_ranks(gather_ms)
        gather_ms = min_over_ranks(gather_ms)
        if tuple(flows.shape) != (n_frames - 1, 2, Hs, Ws) or not bool(torch.isfinite(flows).all()):
            raise SystemExit(f"bench.py: streaming gather returned {tuple(flows.shape)}")
        wpr = tm.get("windows_per_rank", [len(sfd.window_schedule(n_frames, T))])
        streaming = {"frames": n_frames, "windows": sum(wpr), "flows": n_frames - 1, "windows_per_rank": wpr,
                     "critical_path_windows": max(wpr), "ms_total": ms_total, "flows_per_s": (n_frames - 1) / (ms_total / 1e3),
                     "gather_ms": gather_ms, "gather_incl_wait_for_slowest_rank_ms": gather_wait_ms,
                     "gather_bytes_per_rank": tm.get("gather_bytes", 0),
                     "collective": "torch.distributed all_gather_into_tensor over NCCL (inside the timed region)" if world > 1
                     else "none (1 rank)",
                     "launch": "one CUDA-graph replay of the whole forward per window (streamflow_b200.GraphedModel)",
                     "note": "demo.py:518-532 window loop; full model per window (uint8 frames H2D, flows stay on the "
                             "device until the gather; rank 0 copies the 63 flows to the host inside the timed region)"}
        del flows, host_flows, video, frames_list
        # config 3: 8 KITTI-shaped clips sharded 8/N per GPU (strong scaling)
        if world <= 8:
            Hk, Wk = SHAPES["kitti_376x1248"][:2]
            clips = [make_frames(T, Hk, Wk, s).pin_memory() for s in range(8)]
            kgm = sfb.GraphedModel(our_model, (T, 3, Hk, Wk), iters=ITERS, pad_mode="kitti")
            kgm(clips[0])
            tm = {}
            barrier()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            kflows = sfd.run_clips(clips, lambda clip: kgm(clip).clone(), timings=tm)
            e1.record()
            barrier()
            ms_total = max_over_ranks(e0.elapsed_time(e1))
            gather_ms = tm["gather_events"][0].elapsed_time(tm["gather_events"][1]) if "gather_events" in tm else 0.0
            if tuple(kflows.shape) != (8 * PAIRS, 2, Hk, Wk):
                raise SystemExit(f"bench.py: KITTI gather returned {tuple(kflows.shape)}")
            kitti_x8 = {"clips": 8, "clips_per_rank": tm.get("clips_per_rank", [8]), "ms_total": ms_total,
                        "flows_per_s": 8 * PAIRS / (ms_total / 1e3), "gather_ms": min_over_ranks(gather_ms),
                        "gather_incl_wait_for_slowest_rank_ms": max_over_ranks(gather_ms),
                        "gather_bytes_per_rank": tm.get("gather_bytes", 0), "scaling": "strong",
                        "launch": "one CUDA-graph replay of the whole forward per clip (streamflow_b200.GraphedModel)"}
            del kflows, clips, kgm
        torch.cuda.empty_cache()

    # ---- CPU baseline (rank 0, N=1 only): the reference's own operators on this host's cores
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu:
        cores = os.cpu_count() or 1
        torch.set_num_threads(cores)
        cpu_hot_path_sample(host, 1)
        kind, dt = cpu_hot_path_sample(host, ITERS)
        cpu = {"value": PAIRS / dt, "unit": "flow frames/s", "cores": cores, "kind": kind,
               "sample": "1 full step (1 clip: 3 builds + attention + 12 x (3 lookups + aggregate)), fp32 torch CPU, "
                         "the reference's own core/corr.py + core/gma.py" if kind == "reference" else
                         "1 full step of the torch CPU port (oracle/_ref absent)"}

    if rank == 0:
        dom = kernels["gma_aggregate"]
        line = {
            "metric": METRIC, "value": world * PAIRS / (ms_step / 1e3), "unit": "flow frames/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_step, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f16 operands / f32 accumulate", "data": "synthetic",
            "config": {"workload": WORKLOAD, "clips_per_gpu": 1, "pairs": PAIRS, "grid_1_8": [H8, W8], "D": D,
                       "iters": ITERS, "l2": "inputs larger than L2: the step streams a 805 MB pyramid and "
                       "297 MB of softmax numerators per rank, no explicit flush", "parallelism": f"clips x{world}",
                       "launch": "eager public-API calls" if args.eager else
                       "cuda_graph replay of the public-API calls of one clip (streamflow_b200.GraphedCall)"},
            "e2e": e2e,
            "gpu_launches": int(launches),
            "eager_ms_per_step": ms_eager,
            "parity": parity,
            "roofline": {k: dom[k] for k in ("bound", "achieved", "peak", "unit", "frac", "traffic")},
            "kernels": kernels,
            "ncu_traffic_source": traffic_src,
            "torch_gpu_baseline": torch_gpu,
            "full_model": full_model,
            "configs": configs,
            "streaming": streaming,
            "kitti_x8": kitti_x8,
            "cpu_baseline": cpu,
            "clocks": clocks,
            "peaks": peaks["source"],
            "corr_lookup_hbm_gbs": kernels["corr_lookup"]["achieved"],
        }
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--eager", action="store_true",
                    help="time eager public-API calls as the headline instead of the CUDA-graph replay")
    ap.add_argument("--graph", action="store_true", help="(default; kept for compatibility)")
    ap.add_argument("--quick", action="store_true",
                    help="hot path, parity and kernels only: skip the full-model, baseline, config and multi-GPU legs")
    ap.add_argument("--stream-frames", type=int, default=64, help="frames of the streaming configuration")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)
    return run_ours(args)


if __name__ == "__main__":
    sys.exit(main())
