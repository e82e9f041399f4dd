#!/usr/bin/env python
"""bench.py - spectrogram-frames/sec of the FlowSE reverse-ODE sampler (N=5 Euler) on B200.

Contract (driver): python bench.py --gpus N --steps K --warmup W [--impl reference]
  * a "step" = one pass of the hot path over one batch: prior sample + 5 Euler steps (5 NCSN++ evaluations) on one
    synthetic noisy spectrogram batch [B=1, 1, 256, 512] per GPU (BASELINE.json configs[1]);
  * value  = frames/s with inputs resident in HBM (CUDA events, max over ranks, whole job);
  * e2e    = the same through the reference-facing API (get_white_box_solver) with pinned HOST buffers,
             H2D + D2H inside the timed region;
  * roofline / cpu_baseline objects as described in DESIGN.md section "Measurement";
  * --impl reference times the CPU oracle port (torch fp32, all host threads) on a bounded sample of the same workload.
One JSON line on stdout from rank 0.
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

# NCCL prints its version banner on stdout when NCCL_DEBUG=VERSION (this image's default); stdout carries the JSON line only
if os.environ.get("NCCL_DEBUG", "VERSION").upper() == "VERSION":
    os.environ["NCCL_DEBUG"] = "WARN"

import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "spectrogram-frames/sec at N=5 Euler"
UNIT = "frames/s"
N_STEPS_ODE = 5
F_BINS = 256
T_FRAMES = 512
T_REV, T_EPS, SIGMA_MAX = 1.0, 0.03, 0.487
# Algorithmic work: 2.0795 GFLOP per frame per NFE, 24 B per T-F bin per Euler update.
GFLOP_PER_FRAME_NFE = 2.0795


def workload_string(B: int = 1) -> str:
    """config.workload, identical in both arms (the driver compares the strings)."""
    return f"batch={B} complex-STFT 2x{F_BINS}x{T_FRAMES}, N={N_STEPS_ODE} Euler (BASELINE.json configs[1])"


def synth_input(seed: int, B: int = 1, T: int = T_FRAMES) -> torch.Tensor:
    """config 2 input: Y = 0.3 * randn([B,1,256,T]) complex64, seeded."""
    g = torch.Generator().manual_seed(seed)
    return torch.view_as_complex(0.3 * torch.randn(B, 1, F_BINS, T, 2, generator=g))


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return dict(hbm_gbs=d["hbm_gbs"], tf_burst=d["bf16_tflops"], tf_sustained=d["bf16_tflops_sustained"],
                    source="MEASURED_PEAKS.json")
    return dict(hbm_gbs=6650.0, tf_burst=1590.0, tf_sustained=1400.0, source="fallback (B200_PROFILING.md)")


class ClockSampler:
    """nvidia-smi clock / throttle-reason sampling during the timed region."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.gpu = gpu_index
        self.lines = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "100", "-i", str(self.gpu)], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            self.th = threading.Thread(target=self._read, daemon=True)
            self.th.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append((time.time(), line.strip()))

    def stop(self, t0: float, t1: float):
        if self.proc is None:
            return dict(sm_mhz=None, sm_max_mhz=None, reasons=["nvidia-smi unavailable"])
        time.sleep(0.15)
        self.proc.terminate()
        sm, smax, reasons = [], [], set()
        for ts, line in self.lines:
            parts = [p.strip() for p in line.split(",")]
            if len(parts) < 9:
                continue
            try:
                smax.append(float(parts[2]))
                if t0 - 0.05 <= ts <= t1 + 0.05:
                    sm.append(float(parts[1]))
                    for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"),
                                       parts[5:9]):
                        if v.lower().startswith("active"):
                            reasons.add(name)
            except ValueError:
                continue
        return dict(sm_mhz=statistics.median(sm) if sm else None, sm_max_mhz=max(smax) if smax else None,
                    reasons=sorted(reasons), samples=len(sm))


# --------------------------------------------------------------------------------------------------------------
# reference arm / cpu baseline: the oracle port on the host cores
# --------------------------------------------------------------------------------------------------------------
def cpu_sampler_seconds(sd, T: int, n_ode: int, seed: int = 0) -> float:
    from oracle import ncsnpp_oracle as orc
    Y = synth_input(seed, 1, T)
    z = torch.view_as_complex(torch.randn(1, 1, F_BINS, T, 2, generator=torch.Generator().manual_seed(1234)) * (0.5 ** 0.5))
    t0 = time.perf_counter()
    orc.sample(sd, Y, z, n_ode, "euler", T_REV, T_EPS, 0.0, SIGMA_MAX)
    return time.perf_counter() - t0


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from flowmse_b200.checkpoint import synthetic_state_dict
    torch.set_grad_enabled(False)
    torch.set_num_threads(os.cpu_count() or 1)      # torchrun exports OMP_NUM_THREADS=1; the CPU arm uses every host core
    sd = synthetic_state_dict(0)
    cores = torch.get_num_threads()
    # each step = the bench workload ITSELF (one N=5 Euler sampler call on [1,1,256,512]).  Only when the host is so slow
    # that K + W steps would take more than 15 minutes is a slice of it used, and then config.reference_T says so.
    t_cal = cpu_sampler_seconds(sd, 64, 1)
    t_cal = min(t_cal, cpu_sampler_seconds(sd, 64, 1))
    T_s = T_FRAMES
    while T_s > 64 and (args.steps + args.warmup) * t_cal * N_STEPS_ODE * (T_s / 64.0) > 900.0:
        T_s //= 2
    for _ in range(args.warmup):
        cpu_sampler_seconds(sd, T_s, N_STEPS_ODE)
    t0 = time.perf_counter()
    for k in range(args.steps):
        cpu_sampler_seconds(sd, T_s, N_STEPS_ODE, seed=k)
    dt = time.perf_counter() - t0
    val = T_s * args.steps / dt
    sample = (f"oracle port (torch CPU fp32 restatement of the reference path), {cores} threads of {os.cpu_count()} "
              f"host cores; each step = one full N=5 Euler sampler call on [1,1,256,{T_s}]"
              + ("" if T_s == T_FRAMES else f" (the first {T_s} of the workload's {T_FRAMES} frames: bounded sample)"))
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": workload_string(1), "reference_T": T_s,
                   "reference_T_note": "frames per step the CPU arm actually ran (== the workload's 512 unless stated)"},
        "cpu_baseline": {"value": val, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }))


# --------------------------------------------------------------------------------------------------------------
# our arm
# --------------------------------------------------------------------------------------------------------------
def torch_cuda_reference(dev, value, e2e_val):
    """N=5 Euler on [1,1,256,512] through torch-CUDA (oracle restatement on cuda:0), TF32 off and on."""
    from oracle import ncsnpp_oracle as orc
    from flowmse_b200.checkpoint import synthetic_state_dict
    sd_cuda = {k: v.to(dev) for k, v in synthetic_state_dict(0).items()}
    Y = synth_input(1000, 1).to(dev)
    z = torch.view_as_complex(torch.randn(1, 1, F_BINS, T_FRAMES, 2, generator=torch.Generator().manual_seed(1234)) * (0.5 ** 0.5)).to(dev)
    old = (torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32)
    out = {"measured_in_this_run": True, "workload": workload_string(1),
           "what": "oracle restatement of the reference path evaluated by torch on cuda:0 (cuDNN / ATen), eager, CUDA events, "
                   "1 warm-up + 3 timed sampler calls per setting"}
    xs = {}
    try:
        for name, tf32 in (("fp32", False), ("tf32", True)):
            torch.backends.cudnn.allow_tf32 = tf32
            torch.backends.cuda.matmul.allow_tf32 = tf32
            with torch.device(dev):
                orc.sample(sd_cuda, Y, z, N_STEPS_ODE)
                torch.cuda.synchronize()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                for _ in range(3):
                    xs[name] = orc.sample(sd_cuda, Y, z, N_STEPS_ODE)
                e1.record()
                torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / 3
            out[f"{name}_ms_per_step"] = ms
            out[f"{name}_frames_per_s"] = T_FRAMES / (ms * 1e-3)
    finally:
        torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32 = old
    a, b = torch.view_as_real(xs["tf32"]), torch.view_as_real(xs["fp32"])
    out["tf32_frac_outside_tolerance_vs_fp32"] = ((a - b).abs() > 1e-4 + 1e-3 * b.abs()).float().mean().item()
    out["ours_over_fp32"] = value / out["fp32_frames_per_s"]
    out["ours_over_tf32"] = value / out["tf32_frames_per_s"]
    out["ours_e2e_over_tf32"] = e2e_val / out["tf32_frames_per_s"]
    return out


def run_config4(args, model, ctx, dev, world, rank, dist):
    """BASELINE.json configs[3]: 512 synthetic utterances x [1,1,256,512], N=5 Euler, sharded over the ranks by
    flowmse_b200.sharding.enhance_sharded (length-aware LPT assignment, batches of 16, ragged all-gather of the enhanced
    spectrograms over NCCL) - STRONG scaling: the total work is fixed, the window holds sampling + gather."""
    from flowmse_b200 import sharding
    n_utts = args.config4 if args.config4 > 0 else 512
    # every rank holds the same utterance list (a fixed block of 16 distinct spectrograms, reused: HBM-resident inputs)
    base = synth_input(4242, 16).to(dev)
    specs = [base[i % 16] for i in range(n_utts)]

    def enhance(Yb):
        return model.enhance_spec(Yb, N=N_STEPS_ODE)

    def once():
        return sharding.enhance_sharded(specs, enhance, dev, max_batch=16)

    torch.manual_seed(99 + rank)
    warm = [specs[i] for i in range(min(n_utts, 16 * world))]
    sharding.enhance_sharded(warm, enhance, dev, max_batch=16)        # plan + graphs for the (16, 512) shape, warm gather
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    out = once()
    e1.record()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    assert len(out) == n_utts and all(o.shape[-1] == T_FRAMES for o in out)
    return {"workload": f"{n_utts} utterances x [1,1,256,{T_FRAMES}], N=5 Euler, batches of 16, sharded over {world} GPU(s) "
                        f"+ ragged all-gather (BASELINE.json configs[3])",
            "scaling": "strong", "n_gpus": world, "ms": ms.item(), "value": n_utts * T_FRAMES / (ms.item() * 1e-3),
            "unit": UNIT, "gathered_bytes": n_utts * F_BINS * T_FRAMES * 8}


def run_ours(args):
    import torch.distributed as dist
    from flowmse_b200.checkpoint import synthetic_state_dict, flatten_state_dict, unflatten_state_dict
    from flowmse_b200.model import VFModel
    from flowmse_b200.sampling import get_white_box_solver
    from flowmse_b200 import sharding

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py (our arm) needs a CUDA device; there is no CPU fallback")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    torch.set_grad_enabled(False)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    # weights: rank 0 builds the synthetic checkpoint, ONE broadcast of the flat fp32 blob over NCCL
    n_params = None
    if rank == 0:
        blob = flatten_state_dict(synthetic_state_dict(0)).to(dev)
    else:
        from flowmse_b200 import ncsnpp_spec
        blob = torch.empty(ncsnpp_spec.num_params(), dtype=torch.float32, device=dev)
    sharding.broadcast_weights(blob)
    model = VFModel(backbone="ncsnpp", ode="flowmatching", t_eps=T_EPS, T_rev=T_REV)
    model.dnn.load_state_dict(unflatten_state_dict(blob.cpu()), strict=True)
    model.eval()
    ctx = model.flowse_context(dev)
    del blob

    B = args.batch
    Y_host = synth_input(1000 + rank, B).pin_memory()       # one utterance (batch) per rank: weak scaling
    Y = Y_host.to(dev)
    out_host = torch.empty_like(Y_host).pin_memory()
    timesteps = torch.linspace(T_REV, T_EPS, N_STEPS_ODE)

    def step_device():
        z = torch.randn_like(Y)
        return ctx.sample(Y, z, timesteps, solver=0, sigma=model.ode.prior_std())

    def step_e2e():
        Yd = Y_host.to(dev, non_blocking=True)
        sampler = get_white_box_solver("euler", model.ode, model, Y=Yd, Y_prior=Yd, T_rev=T_REV, t_eps=T_EPS,
                                       N=N_STEPS_ODE)
        x, _ = sampler()
        out_host.copy_(x, non_blocking=True)
        return x

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def gather_result(last):
        """the job's single gather of enhanced spectrograms (1 MiB per rank here)"""
        if world > 1:
            gathered = [torch.empty_like(torch.view_as_real(last)) for _ in range(world)]
            dist.all_gather(gathered, torch.view_as_real(last).contiguous())

    def timed(fn, steps, warmup):
        last = None
        for _ in range(warmup):
            last = fn()
        gather_result(last)              # the collective is warmed like everything else in the window
        barrier()
        l0 = ctx.kernel_launches()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        w0 = time.time()
        e0.record()
        for _ in range(steps):
            last = fn()
        gather_result(last)
        e1.record()
        barrier()
        w1 = time.time()
        mine = e0.elapsed_time(e1)
        ms = torch.tensor([mine], device=dev)
        per_rank = [mine]
        if world > 1:
            allms = [torch.empty_like(ms) for _ in range(world)]
            dist.all_gather(allms, ms)
            per_rank = [float(v.item()) for v in allms]
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return ms.item(), ctx.kernel_launches() - l0, (w0, w1), per_rank

    clocks = ClockSampler(local_rank)
    clocks.start()
    time.sleep(0.3)
    ms_dev, launches, (w0, w1), per_rank_ms = timed(step_device, args.steps, args.warmup)
    clk = clocks.stop(w0, w1)
    ms_e2e, _, _, _ = timed(step_e2e, args.steps, args.warmup)

    frames_per_step = B * T_FRAMES * world
    value = frames_per_step * args.steps / (ms_dev * 1e-3)
    e2e_val = frames_per_step * args.steps / (ms_e2e * 1e-3)
    peaks = measured_peaks()

    result = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms_dev / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": {"workload": workload_string(B),
                   "per_gpu": "one such batch per GPU per step (weak scaling), sigma_max=0.487, synthetic seeded weights "
                              "(65.6 M params); one NCCL weight broadcast before, one all-gather of the outputs inside the window",
                   "l2": "working set > L2: 250 MiB packed weights + ~1.3 GiB activations per NFE vs 126 MB L2",
                   "arithmetic": "fp32 parity via fp16 hi/lo split on tcgen05 (3 MMAs per product), fp32 accumulate"},
        "clocks": clk,
        "e2e": {"value": e2e_val, "unit": UNIT, "h2d_bytes_per_step": Y_host.numel() * 8,
                "d2h_bytes_per_step": out_host.numel() * 8, "ms_per_step": ms_e2e / args.steps,
                "api": "flowmse_b200.sampling.get_white_box_solver(...)() with pinned host Y, result copied to host"},
        "gpu_launches": int(launches),
        "per_rank_ms_per_step": [round(v / args.steps, 4) for v in per_rank_ms],
    }

    if rank == 0:
        # ---- roofline of the dominant kernel: per-op CUDA-event timing of one NFE (same plan, same buffers) --------
        ctx.profile_forward()
        ops = ctx.profile_forward()
        tot_ms = sum(o["ms"] for o in ops)
        by_kind = {}
        for o in ops:
            k = by_kind.setdefault(o["kind"], dict(ms=0.0, n=0, flops=0.0))
            k["ms"] += o["ms"]; k["n"] += 1; k["flops"] += o["flops"]
        empty = dict(ms=0.0, n=0, flops=0.0)
        conv = {k: by_kind.get("conv_gemm", empty)[k] + by_kind.get("conv_halo", empty)[k] for k in ("ms", "n", "flops")}
        # dominant kernel = conv_halo_kernel<128,1,0>: the 3x3 convs of the three highest resolutions (the engine reports
        # which conv kernel each op uses); the low-resolution launches use the per-tap kernel with cluster split-K
        halo = [o for o in ops if o["kind"] == "conv_halo"]
        ncu_traffic = {}
        try:    # DRAM bytes / tensor-pipe activity of ONE launch of this kernel from the committed `ncu --set full` capture
            summ = json.load(open(os.path.join(ROOT, "profiles", "r2_conv_halo_full_summary.json")))
            # the 256 -> 128 layers at 256x512 (K = 2304) are the longest launches of the capture without shortcut chunks
            recs = [r for r in summ if "conv_halo_kernel<128, 1, 0, 1>" in r["kernel"] and 130 < r.get("dram_read_MB", 0) < 140
                    and r.get("duration_us_under_ncu", 0) > 170]
            rec = recs[0]
            ncu_traffic = {"bytes": (rec["dram_read_MB"] + rec["dram_write_MB"]) * 1e6,
                           "tensor_pipe_active_pct_of_elapsed": rec["tensor_pipe_active_pct_of_elapsed"],
                           "note": "dram__bytes_read.sum + dram__bytes_write.sum of ONE launch (256x512, K = 2304, operands prepared "
                                   "in the kernel; algorithmic 202.6 MB: 128 MiB fp32 input + 64 MiB fp32 output + 1.2 MB weights) "
                                   "from profiles/r2_conv_halo_full_summary.json (ncu --set full, round 2)"}
        except Exception:
            pass
        h_ms = sum(o["ms"] for o in halo); h_fl = sum(o["flops"] for o in halo)
        achieved = h_fl / (h_ms * 1e-3) / 1e12
        all_conv = conv["flops"] / (conv["ms"] * 1e-3) / 1e12
        result["roofline"] = {
            "bound": "tensor", "kernel": "conv_halo_kernel<128,1,0,1> (3x3 ResBlock convolutions at 256xT, 128xT/2, 64xT/4; GroupNorm + SiLU + fp16 split of the operand fused in; 4 of its 38 launches per NFE run the TMA-operand variant <128,1,0,0>)",
            "achieved": achieved, "peak": peaks["tf_sustained"], "unit": "TFLOP/s", "frac": achieved / peaks["tf_sustained"],
            "traffic": ncu_traffic.get("bytes"), "traffic_note": ncu_traffic.get("note"),
            "ncu_tensor_pipe_active_pct_of_elapsed": ncu_traffic.get("tensor_pipe_active_pct_of_elapsed"),
            "frac_of_burst_peak": achieved / peaks["tf_burst"],
            "issued_mma_tflops": 3 * achieved, "issued_frac": 3 * achieved / peaks["tf_sustained"],
            "issued_frac_of_burst_peak": 3 * achieved / peaks["tf_burst"],
            "avg_launch_ms": h_ms / max(1, len(halo)), "launches_per_nfe": len(halo),
            "algorithmic_gflop_per_launch_avg": h_fl / max(1, len(halo)) / 1e9,
            "share_of_nfe_time": h_ms / tot_ms,
            "peak_source": peaks["source"] + " bf16 dense sustained (the kernel is timed inside a step; fp16 issues at the "
                           "same rate); parity costs 3 issued MMAs per algorithmic product, so frac <= 1/3",
            "how": "CUDA events after every op of one NFE queued behind a spin kernel on a private stream, so the kernels run "
                   "back to back as in the graph replay (flowse_profile_forward), after the timed region",
            "all_conv_launches": {"launches_per_nfe": conv["n"], "achieved": all_conv, "frac": all_conv / peaks["tf_sustained"],
                                  "algorithmic_gflop_per_nfe": conv["flops"] / 1e9, "share_of_nfe_time": conv["ms"] / tot_ms},
            "nfe_ms_by_kernel_family": {k: round(v["ms"], 4) for k, v in by_kind.items()},
        }
        # the same layers with the standalone operand-prep pass (option fuse_prep = 0), measured in this run: what the
        # fusion trades - the conv kernel alone is faster, conv + prep together are slower
        try:
            ctx.set_option("fuse_prep", 0)
            step_device(); step_device()
            ctx.profile_forward()
            ops0 = ctx.profile_forward()
            halo0 = [o for o in ops0 if o["kind"] == "conv_halo"]
            h0_ms = sum(o["ms"] for o in halo0); h0_fl = sum(o["flops"] for o in halo0)
            prep0 = sum(o["ms"] for o in ops0 if o["kind"] == "gn_prep")
            prep1 = sum(o["ms"] for o in ops if o["kind"] == "gn_prep")
            a0 = h0_fl / (h0_ms * 1e-3) / 1e12
            result["roofline"]["standalone_prep_variant"] = {
                "achieved": a0, "frac": a0 / peaks["tf_sustained"], "issued_frac": 3 * a0 / peaks["tf_sustained"],
                "conv_halo_ms_per_nfe": h0_ms, "gn_prep_ms_per_nfe": prep0,
                "fused": {"conv_halo_ms_per_nfe": h_ms, "gn_prep_ms_per_nfe": prep1},
                "note": "fuse_prep = 0: the halo kernel takes ready-made fp16 hi/lo operands by TMA (round-1 design); its own "
                        "fraction is higher, but conv + prep per NFE cost more than the fused kernel + the remaining prep"}
        except Exception as e:
            result["roofline"]["standalone_prep_variant"] = {"error": str(e)}
        finally:
            ctx.set_option("fuse_prep", 1)
            step_device()
        # the dominant layer (256 -> 128 at 256x512, K = 2304: 3 launches per NFE, 22 % of the FLOPs) timed ALONE: the same
        # kernel through the op-level C ABI, 20 launches back to back between two events (no per-op event overhead; the
        # in-NFE figure above carries ~5 us of event serialisation per op) -> against the burst peak
        try:
            import numpy as np
            gk = torch.Generator(device=dev).manual_seed(0)
            a32 = torch.randn(1, F_BINS, T_FRAMES, 256, device=dev, generator=gk)
            hi = a32.half(); A_op = torch.stack([hi, (a32 - hi.float()).half()]).contiguous(); del a32, hi
            w = torch.randn(128, 256, 3, 3) / np.sqrt(256 * 9)
            Wp, wexp = ctx.pack_conv_weights(w, None, 128)
            bias0 = torch.zeros(1, 128, device=dev)
            out_op = torch.empty(1, F_BINS, T_FRAMES, 128, device=dev)
            for _ in range(3):
                ctx.op_conv_gemm(A_op, Wp, wexp, bias0, 128, out=out_op, impl=2)
            ka, kb = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            iters = 20
            ka.record()
            for _ in range(iters):
                ctx.op_conv_gemm(A_op, Wp, wexp, bias0, 128, out=out_op, impl=2)
            kb.record(); torch.cuda.synchronize()
            us = ka.elapsed_time(kb) / iters * 1e3
            fl = 2.0 * F_BINS * T_FRAMES * 128 * 2304
            alone = fl / (us * 1e-6) / 1e12
            result["roofline"]["kernel_alone"] = {
                "layer": "256x512, Cin 256 -> Cout 128, K = 2304 (ResBlock Conv_0 of the top-level up path), operands given "
                         "(TMA variant of the kernel, as the op-level C ABI exposes it)",
                "us_per_launch": us, "achieved": alone, "peak": peaks["tf_burst"], "unit": "TFLOP/s", "frac": alone / peaks["tf_burst"],
                "issued_mma_tflops": 3 * alone, "issued_frac": 3 * alone / peaks["tf_burst"],
                "algorithmic_bytes": 201.3e6, "peak_source": peaks["source"] + " bf16 dense burst (kernel timed alone)"}
            del A_op, Wp, out_op
        except Exception as e:      # the headline numbers do not depend on this extra measurement
            result["roofline"]["kernel_alone"] = {"error": str(e)}
        # Euler-update kernel against the HBM roofline, on a buffer larger than L2 (B=1 moves only 3 MiB per launch)
        nbig = 32 * 1024 * 1024                      # 32 Mi complex = 256 MiB per tensor
        xa = torch.view_as_complex(torch.randn(nbig, 2, device=dev)); va = torch.view_as_complex(torch.randn(nbig, 2, device=dev))
        for _ in range(3):
            ctx.euler_step(xa, va, 0.2425)
        ea, eb = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ea.record()
        for _ in range(10):
            ctx.euler_step(xa, va, 0.2425)
        eb.record(); torch.cuda.synchronize()
        gbs = 24.0 * nbig * 10 / (ea.elapsed_time(eb) * 1e-3) / 1e9
        result["roofline_euler"] = {"bound": "hbm", "kernel": "axpy_kernel (x + v*dt, complex64)", "achieved": gbs,
                                    "peak": peaks["hbm_gbs"], "unit": "GB/s", "frac": gbs / peaks["hbm_gbs"],
                                    "bytes_per_bin": 24, "sample": "32 Mi bins (768 MiB traffic per launch, > L2)"}
        del xa, va
        # ---- CPU baseline (oracle port) on a bounded sample ------------------------------------------------------
        if not args.no_cpu_baseline and world == 1:
            sd = synthetic_state_dict(0)
            torch.set_num_threads(os.cpu_count() or 1)
            cores = torch.get_num_threads()
            cpu_sampler_seconds(sd, 64, 1)                       # warm-up (thread pool, oneDNN primitives)
            reps, secs = 0, 0.0
            while reps < 2 or (secs < 10.0 and reps < 6):        # ~10-30 s of CPU work
                secs += cpu_sampler_seconds(sd, T_FRAMES, N_STEPS_ODE, seed=reps)
                reps += 1
            result["cpu_baseline"] = {
                "value": reps * T_FRAMES / secs, "unit": UNIT, "cores": cores, "kind": "port",
                "sample": f"oracle port (torch CPU fp32 restatement of the reference path), {reps} full N=5 Euler sampler calls "
                          f"on the bench workload [1,1,256,{T_FRAMES}] ({secs:.1f} s), {cores} threads of {os.cpu_count()} "
                          f"host cores"}
        # ---- the same path evaluated by PyTorch on this GPU, measured in this run (BASELINE.json configs[1]: "vs reference
        # torch-cuda").  The reference package cannot travel to the GPU box, so this is its restatement (oracle/) run with
        # every tensor on cuda:0 - the ATen ops the reference dispatches: cuDNN convolutions, native_group_norm, bmm,
        # softmax - timed with CUDA events after a warm-up call, TF32 off (fp32: the parity setting) and on (PyTorch's
        # default for convolutions: the speed baseline).  It is a checker timed beside the product, never part of it.
        if not args.no_torch_reference and world == 1:
            try:
                result["gpu_torch_reference"] = torch_cuda_reference(dev, value, e2e_val)
            except Exception as e:
                result["gpu_torch_reference"] = {"error": f"{type(e).__name__}: {e}"}
    # ---- BASELINE.json configs[3], strong scaling (every rank takes part; reported beside the weak-scaling headline).
    # Runs last: it re-plans the context for batches of 16.
    if args.config4 != 0:
        try:
            c4 = run_config4(args, model, ctx, dev, world, rank, dist)
        except Exception as e:
            if world > 1:
                raise
            c4 = {"error": f"{type(e).__name__}: {e}"}
        result["extra"] = {"config4_strong_scaling": c4}
    if rank == 0:
        print(json.dumps(result))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--batch", type=int, default=1)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-torch-reference", action="store_true", help="skip the live torch-CUDA measurement (N=1 only)")
    ap.add_argument("--config4", type=int, default=-1,
                    help="utterances of the configs[3] strong-scaling leg reported in extra (-1 = 512, 0 = skip)")
    args = ap.parse_args()
    if args.warmup < 3 and args.impl == "ours":
        args.warmup = 3
    if args.impl == "reference":
        run_reference_arm(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
