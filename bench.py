#!/usr/bin/env python
"""bench.py — BASELINE.json metric: clips/sec (10 s @ 32 kHz) of a passt_s p16_128 TRAIN step.

Workload (BASELINE.json configs[1]): passt_s, s_patchout_t=40, s_patchout_f=4 (N=474 tokens),
64 clips per GPU, bf16 tensor-core arithmetic with fp32 master weights / residual stream, synthetic AudioSet-shaped
data (0.1*randn waveforms, multi-hot 527-class targets), random-init weights.  One step = waveform -> fused mel
kernel (band augmentation + SpecAugment on) -> patchout-ViT forward -> BCE-with-logits -> hand-written backward
-> (N>1: NCCL gradient all-reduce) -> AdamW step.  Weak scaling: 64 clips per GPU at every N.

  python bench.py --gpus N --steps K --warmup W          # candidate (sm_100a kernels)
  python bench.py --impl reference ...                   # the reference algorithm on the host CPU cores (oracle port)

Prints ONE JSON line on rank 0.
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

import torch
import torch.nn.functional as F

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

CLIP_LEN = 320000
N_CLASSES = 527
BATCH_PER_GPU = 64
NET_KW = dict(s_patchout_t=40, s_patchout_f=4)
# all host threads the oracle can use productively: torch's intra-op pool stops scaling (and thrashes) well before the
# 100+ hardware threads of the GPU box on these matrix sizes, so cap it; the count actually used is reported as `cores`
CPU_THREADS = max(1, min(os.cpu_count() or 1, int(os.environ.get("PASST_CPU_THREADS", "32"))))
WORKLOAD = "passt_s p16_128 s_patchout_t=40 s_patchout_f=4, batch=64/GPU, 10s@32kHz, train step bf16"


def load_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.isfile(path):
        with open(path) as f:
            p = json.load(f)
        return dict(hbm_gbs=p["hbm_gbs"], bf16_tflops=p["bf16_tflops"], bf16_sustained=p["bf16_tflops_sustained"],
                    source="measured")
    return dict(hbm_gbs=6650.0, bf16_tflops=1590.0, bf16_sustained=1400.0, source="fallback")


def train_flops_per_clip(ntok, depth=12, n_classes=N_CLASSES):
    fwd = 2 * (ntok - 2) * 256 * 768 + depth * (2 * ntok * 768 * 9216 + 4 * ntok * ntok * 768) + 2 * 768 * n_classes
    return 3 * fwd


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""

    def __init__(self, index):
        self.index = index
        self.proc = None
        self.lines = []

    def start(self):
        q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
             "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={q}", "--format=csv,noheader,nounits", "-lms", "100",
                                          "-i", str(self.index)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL,
                                         text=True)
            self.t = threading.Thread(target=self._pump, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], 0.0, set()
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); mx = max(mx, float(f[2]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        busy = [s for s in sm if s > 0.5 * mx] or sm
        return {"sm_mhz": statistics.median(busy) if busy else None, "sm_max_mhz": mx or None,
                "reasons": sorted(reasons), "samples": len(sm)}


# ======================================================================================================
# reference arm: the reference algorithm (CPU oracle port, validated bit-exact against /root/reference) on host cores
# ======================================================================================================
def cpu_train_step_rate(sample_clips=2, steps=1, warmup=0):
    """clips/s of one full train step (mel train-mode + net fwd/bwd + AdamW) of the oracle port on the CPU."""
    from oracle import passt_oracle as O
    torch.set_num_threads(CPU_THREADS)
    mcfg, ncfg = O.MelCfg(), O.NetCfg(**NET_KW)
    params = {k: v.clone().requires_grad_(not k.startswith("head_dist")) for k, v in O.synth_params(ncfg, 0).items()}
    opt = torch.optim.AdamW([p for p in params.values() if p.requires_grad], lr=2e-5, weight_decay=1e-4)
    torch.manual_seed(0)
    wave = 0.1 * torch.randn(sample_clips, CLIP_LEN)
    y = (torch.rand(sample_clips, N_CLASSES) < 0.005).float()

    def step():
        d = O.draw_mel(mcfg, True, sample_clips)
        with torch.no_grad():
            spec = O.mel_frontend(wave, mcfg, d, True).unsqueeze(1)
        dp = O.draw_patchout(ncfg, 12, 99, True)
        logits, _ = O.passt_forward(params, spec, ncfg, dp)
        loss = F.binary_cross_entropy_with_logits(logits, y)
        opt.zero_grad(set_to_none=True)
        loss.backward()
        opt.step()
        return float(loss.detach())

    for _ in range(warmup):
        step()
    t0 = time.perf_counter()
    for _ in range(steps):
        step()
    dt = time.perf_counter() - t0
    return sample_clips * steps / dt, dt / steps


REFERENCE_BUDGET_S = float(os.environ.get("PASST_REF_BUDGET_S", "150"))


def run_reference(args, rank, world):
    """Reference arm: the reference algorithm (oracle port, bit-exact with /root/reference on CPU) on the host cores.
    One step = one full train step on a bounded 1-clip sample of the workload; at most --steps steps are timed, fewer
    if the time budget (a few minutes) would be exceeded — the number actually timed is reported as `steps`."""
    if rank != 0:
        return
    from oracle import passt_oracle as O
    torch.set_num_threads(CPU_THREADS)
    sample = 1
    mcfg, ncfg = O.MelCfg(), O.NetCfg(**NET_KW)
    params = {k: v.clone().requires_grad_(not k.startswith("head_dist")) for k, v in O.synth_params(ncfg, 0).items()}
    opt = torch.optim.AdamW([p for p in params.values() if p.requires_grad], lr=2e-5, weight_decay=1e-4)
    torch.manual_seed(0)
    wave = 0.1 * torch.randn(sample, CLIP_LEN)
    y = (torch.rand(sample, N_CLASSES) < 0.005).float()

    def step():
        d = O.draw_mel(mcfg, True, sample)
        with torch.no_grad():
            spec = O.mel_frontend(wave, mcfg, d, True).unsqueeze(1)
        dp = O.draw_patchout(ncfg, 12, 99, True)
        logits, _ = O.passt_forward(params, spec, ncfg, dp)
        loss = F.binary_cross_entropy_with_logits(logits, y)
        opt.zero_grad(set_to_none=True)
        loss.backward()
        opt.step()
        return float(loss.detach())

    t_start = time.perf_counter()
    warm = 1 if args.warmup > 0 else 0
    for _ in range(warm):
        step()
    warm_s = time.perf_counter() - t_start
    done, t0 = 0, time.perf_counter()
    while done < max(1, args.steps):
        step()
        done += 1
        elapsed = time.perf_counter() - t0
        if (time.perf_counter() - t_start) + elapsed / done > REFERENCE_BUDGET_S:
            break
    dt = time.perf_counter() - t0
    rate = sample * done / dt
    line = {
        "impl": "reference", "metric": "clips/sec (10s@32kHz) passt_s p16_128 train step", "value": rate,
        "unit": "clips/s", "n_gpus": args.gpus, "steps": done, "warmup": warm,
        "ms_per_step": dt / done * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
        "data": "synthetic",
        "config": {"workload": WORKLOAD, "sample": f"{sample} clip per step on the host CPU ({CPU_THREADS} threads)",
                   "requested_steps": args.steps, "time_budget_s": REFERENCE_BUDGET_S, "warmup_s": warm_s},
        "cpu_baseline": {"value": rate, "unit": "clips/s", "cores": CPU_THREADS, "kind": "port",
                         "sample": f"{done} x {sample}-clip train step (mel train + fwd + bwd + AdamW), oracle port, fp32"},
        "e2e": {"value": rate, "unit": "clips/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    _emit(line)


# ======================================================================================================
# candidate arm
# ======================================================================================================
def run_candidate(args, rank, local_rank, world):
    import torch.distributed as dist
    from passt_b200 import _lib as L
    from passt_b200 import engine
    from passt_b200.passt import get_model
    from passt_b200.preprocess import AugmentMelSTFT
    from passt_b200.ddp import GradAllReducer

    if not torch.cuda.is_available():
        raise RuntimeError("bench.py candidate arm needs a CUDA device (sm_100a); there is no CPU fallback")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    L.load()
    peaks = load_peaks()
    B = BATCH_PER_GPU
    torch.manual_seed(rank)
    import contextlib
    with contextlib.redirect_stdout(sys.stderr):     # the module prints the reference's "FMAX is None" notice
        mel = AugmentMelSTFT(n_mels=128, sr=32000, win_length=800, hopsize=320, n_fft=1024, freqm=48, timem=192,
                             fmin=0.0, fmax=None, fmin_aug_range=10, fmax_aug_range=2000).to(dev).train()
    torch.manual_seed(0)   # identical initial weights on every rank
    net = get_model(arch="passt_s_swa_p16_128_ap476", pretrained=False, n_classes=N_CLASSES, **NET_KW).to(dev).train()
    use_graph = bool(args.graph)
    opt_params = [p for n, p in net.named_parameters() if not n.startswith("head_dist")]
    if args.optim == "own":
        from passt_b200.optim import FusedAdamW
        opt = FusedAdamW(opt_params, lr=2e-5, weight_decay=1e-4).attach(net)
    else:
        opt = torch.optim.AdamW(opt_params, lr=2e-5, weight_decay=1e-4, fused=True, capturable=use_graph)
    reducer = GradAllReducer(net) if world > 1 else None
    torch.manual_seed(1000 + rank)
    n_batches = 4
    host_waves = [(0.1 * torch.randn(B, CLIP_LEN)).pin_memory() for _ in range(n_batches)]
    dev_waves = [w.to(dev) for w in host_waves]
    y = (torch.rand(B, N_CLASSES, device=dev) < 0.005).float()

    def eager_step(wave_dev):
        with torch.no_grad():
            spec = mel(wave_dev).unsqueeze(1)
        logits, _ = net(spec)
        loss = F.binary_cross_entropy_with_logits(logits, y)
        opt.zero_grad(set_to_none=True)
        loss.backward()
        if reducer is not None:
            reducer.all_reduce()
        opt.step()
        return loss

    graphed = None
    if use_graph:
        from passt_b200.graphed import GraphedTrainStep
        graphed = GraphedTrainStep(mel, net, opt, F.binary_cross_entropy_with_logits, dev_waves[0], y,
                                   reducer=reducer, warmup=3)

    def train_step(wave_dev, consumed=None):
        # public API: either the eager modules (mel -> net -> loss.backward -> opt.step) or the same step replayed
        # as one CUDA graph (passt_b200.graphed.GraphedTrainStep; inputs are copied into its static buffers).
        # consumed: event recorded once the input buffer may be refilled (graph: right after the staging copy)
        if graphed is not None:
            return graphed(wave_dev, None, consumed=consumed)
        loss = eager_step(wave_dev)
        if consumed is not None:
            consumed.record(torch.cuda.current_stream())
        return loss

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(steps):
            fn(i)
        e1.record()
        barrier()
        ms = e0.elapsed_time(e1)
        if world > 1:
            t = torch.tensor([ms], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms

    # ---- warm-up
    for i in range(max(3, args.warmup)):
        train_step(dev_waves[i % n_batches])
    barrier()
    ntok = net.last_plan.ntok
    if world > 1:
        # replicas must stay bit-identical: same initial weights + averaged gradients => same parameters on every rank
        chk = torch.stack([p.detach().double().sum() for p in net.parameters()]).sum().reshape(1)
        lo, hi = chk.clone(), chk.clone()
        dist.all_reduce(lo, op=dist.ReduceOp.MIN)
        dist.all_reduce(hi, op=dist.ReduceOp.MAX)
        if float((hi - lo).abs()) != 0.0:
            raise RuntimeError(f"DDP replicas diverged after warm-up: parameter checksum spread {float(hi - lo)}")

    # ---- (1) device-resident inputs
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    L.reset_launch_count()
    ms_dev = timed(lambda i: train_step(dev_waves[i % n_batches]), args.steps)
    launches = L.launch_count()
    if graphed is not None:
        # kernels are replayed by the graph; count the launches of one eager step and scale
        L.reset_launch_count()
        eager_step(dev_waves[0])
        launches = L.launch_count() * args.steps
        net._wcache.invalidate()
    clocks = sampler.stop() if rank == 0 else None

    # ---- (2) end to end through the public API with HOST buffers: every step's waveform batch is copied from pinned
    #          host memory (on a copy stream, one batch ahead of the compute, like a prefetching loader) and every
    #          step's loss is read back to the host (asynchronously, consumed one step later, like a logging hook).
    copy_stream = torch.cuda.Stream(device=dev)
    dev_bufs = [torch.empty(B, CLIP_LEN, device=dev) for _ in range(2)]
    h2d_done = [torch.cuda.Event() for _ in range(2)]
    buf_free = [torch.cuda.Event() for _ in range(2)]
    loss_hosts = [torch.zeros(1).pin_memory() for _ in range(2)]
    loss_done = [torch.cuda.Event() for _ in range(2)]
    losses = []

    def upload(i):
        k = i & 1
        with torch.cuda.stream(copy_stream):
            copy_stream.wait_event(buf_free[k])
            dev_bufs[k].copy_(host_waves[i % n_batches], non_blocking=True)
            h2d_done[k].record(copy_stream)

    def e2e_run(steps):
        cur = torch.cuda.current_stream()
        for k in range(2):
            buf_free[k].record(cur)
        upload(0)
        for i in range(steps):
            k = i & 1
            if i + 1 < steps:
                upload(i + 1)
            cur.wait_event(h2d_done[k])
            loss = train_step(dev_bufs[k], consumed=buf_free[k])
            loss_hosts[k].copy_(loss.detach().reshape(1), non_blocking=True)
            loss_done[k].record(cur)
            if i > 0:
                loss_done[k ^ 1].synchronize()
                losses.append(float(loss_hosts[k ^ 1]))
        loss_done[(steps - 1) & 1].synchronize()
        losses.append(float(loss_hosts[(steps - 1) & 1]))

    e2e_run(2)

    def timed_once(fn):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        barrier()
        ms = e0.elapsed_time(e1)
        if world > 1:
            t = torch.tensor([ms], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms

    ms_e2e = timed_once(lambda: e2e_run(args.steps))
    if os.environ.get("PASST_BENCH_E2E_DIAG") and rank == 0:
        # where the end-to-end overhead comes from (stderr only): same loop without the H2D copies / the loss read-back
        def variant(h2d, d2h):
            def run():
                cur = torch.cuda.current_stream()
                for k in range(2):
                    buf_free[k].record(cur)
                if h2d:
                    upload(0)
                for i in range(args.steps):
                    k = i & 1
                    if h2d:
                        if i + 1 < args.steps:
                            upload(i + 1)
                        cur.wait_event(h2d_done[k])
                    loss = train_step(dev_bufs[k])
                    buf_free[k].record(cur)
                    if d2h:
                        loss_hosts[k].copy_(loss.detach().reshape(1), non_blocking=True)
                        loss_done[k].record(cur)
                        if i > 0:
                            loss_done[k ^ 1].synchronize()
                torch.cuda.synchronize()
            return timed_once(run) / args.steps
        for h2d, d2h in ((True, True), (False, True), (True, False), (False, False)):
            print(f"[e2e diag] h2d={h2d} d2h={d2h}: {variant(h2d, d2h):.3f} ms/step", file=sys.stderr)
        # same event structure, but only 4 KB cross the bus: separates "bytes" from "stream/event structure"
        tiny_host = torch.zeros(1024).pin_memory()
        tiny_dev = torch.zeros(1024, device=dev)

        def upload_tiny(i):
            k = i & 1
            with torch.cuda.stream(copy_stream):
                copy_stream.wait_event(buf_free[k])
                tiny_dev.copy_(tiny_host, non_blocking=True)
                h2d_done[k].record(copy_stream)

        def run_tiny():
            cur = torch.cuda.current_stream()
            for k in range(2):
                buf_free[k].record(cur)
            upload_tiny(0)
            for i in range(args.steps):
                k = i & 1
                if i + 1 < args.steps:
                    upload_tiny(i + 1)
                cur.wait_event(h2d_done[k])
                train_step(dev_bufs[k])
                buf_free[k].record(cur)
            torch.cuda.synchronize()
        print(f"[e2e diag] 4 KB copies, same events: {timed_once(run_tiny) / args.steps:.3f} ms/step", file=sys.stderr)

        # full copies, but issued right AFTER the step's launch instead of before it
        def run_late():
            cur = torch.cuda.current_stream()
            for k in range(2):
                buf_free[k].record(cur)
            upload(0)
            for i in range(args.steps):
                k = i & 1
                cur.wait_event(h2d_done[k])
                train_step(dev_bufs[k])
                buf_free[k].record(cur)
                if i + 1 < args.steps:
                    upload(i + 1)
            torch.cuda.synchronize()
        print(f"[e2e diag] full copies issued after the launch: {timed_once(run_late) / args.steps:.3f} ms/step", file=sys.stderr)

    # ---- (3) roofline of the dominant kernel family (tcgen05 GEMM): CUDA events around every GEMM launch in a
    #          repeat of the timed steps (kept out of the headline timing so the events do not perturb it)
    engine.GEMM_TRACE = []
    ms_instr = timed(lambda i: eager_step(dev_waves[i % n_batches]), args.steps)
    torch.cuda.synchronize()
    trace, engine.GEMM_TRACE = engine.GEMM_TRACE, None
    gemm_ms = sum(a.elapsed_time(b) for a, b, _ in trace)
    gemm_flops = sum(f for _, _, f in trace)
    n_gemm = max(1, len(trace))
    achieved_tf = gemm_flops / (gemm_ms * 1e-3) / 1e12 if gemm_ms > 0 else 0.0
    peak_tf = peaks["bf16_sustained"]

    # ---- (4) the two other kernels the contract names, timed alone with CUDA events
    other = None
    if rank == 0 and world == 1:
        other = _other_kernel_rooflines(mel, net, dev_waves[0], B, ntok, peaks)

    total_clips = B * world * args.steps
    value = total_clips / (ms_dev * 1e-3)
    e2e_value = total_clips / (ms_e2e * 1e-3)
    step_flops = train_flops_per_clip(ntok) * B

    line = None
    if rank == 0:
        cpu_rate, cpu_sec = cpu_train_step_rate(sample_clips=1, steps=3, warmup=1) if world == 1 else (None, None)
        line = {
            "metric": "clips/sec (10s@32kHz) passt_s p16_128 train step", "value": value, "unit": "clips/s",
            "n_gpus": world, "steps": args.steps, "warmup": max(3, args.warmup), "ms_per_step": ms_dev / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
            "config": {"workload": WORKLOAD, "tokens": ntok, "global_batch": B * world, "parallelism": f"dp{world}",
                       "optimizer": ("passt_b200.FusedAdamW (one launch, refreshes the bf16 weight copies), fp32 master weights"
                                     if args.optim == "own" else "torch.optim.AdamW(fused), fp32 master weights"),
                       "loss": "BCE-with-logits, 527 classes",
                       "cuda_graph": bool(use_graph),
                       "l2": "4 rotating input batches; per-step working set (~10 GB of activations) >> 126 MB L2",
                       "model_flops_per_step": step_flops,
                       "model_tflops": step_flops * world / (ms_dev / args.steps * 1e-3) / 1e12},
            "e2e": {"value": e2e_value, "unit": "clips/s", "h2d_bytes_per_step": B * CLIP_LEN * 4,
                    "d2h_bytes_per_step": 4, "ms_per_step": ms_e2e / args.steps},
            "gpu_launches": launches,
            "clocks": clocks,
            "roofline": {"kernel": "gemm_kernel<BN,MODE> (tcgen05 GEMM family: fwd, dgrad, wgrad)", "bound": "tensor",
                         "achieved": achieved_tf, "peak": peak_tf, "unit": "TFLOP/s",
                         "frac": achieved_tf / peak_tf if peak_tf else None,
                         # dram__bytes_read+write of the qkv-forward launch (107.4 GFLOP, 190 MB algorithmic) from
                         # profiles/r1_ncu_gemm2_full_v2.txt (ncu --set full): 50.3 MB read + 89.5 MB written
                         "traffic": 139.7e6, "traffic_launch": "qkv forward M=30336 N=2304 K=768",
                         "peak_source": f"MEASURED_PEAKS.json bf16_tflops_sustained ({peaks['source']})",
                         "launches_per_step": n_gemm / args.steps, "avg_launch_ms": gemm_ms / n_gemm,
                         "share_of_step": gemm_ms / ms_instr if ms_instr else None,
                         "how": "CUDA events around each GEMM launch in an eager (non-graph) repeat of the timed steps"},
        }
        if other is not None:
            line["other_kernels"] = other
        if cpu_rate is not None:
            line["cpu_baseline"] = {"value": cpu_rate, "unit": "clips/s", "cores": CPU_THREADS, "kind": "port",
                                    "sample": "3 x 1-clip train step after 1 warm-up (mel + fwd + bwd + AdamW) of the CPU oracle port, fp32"}
        _emit(line)
    return line


def _other_kernel_rooflines(mel, net, wave, B, ntok, peaks):
    """mel_kernel against the HBM roofline (algorithmic 1.792 MB per clip) and the attention kernels against the
    tensor roofline (4 / 10 N^2 d flops per clip and head), each timed alone: 3 warm-up + 20 launches."""
    from passt_b200 import _lib as L

    def avg_ms(fn, n=20):
        for _ in range(3):
            fn()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        e0.record()
        for _ in range(n):
            fn()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / n

    out = {}
    was_training = mel.training
    mel.eval()                       # fixed band, no SpecAugment draws: the launch is the mel kernel alone
    with torch.no_grad():
        ms = avg_ms(lambda: mel(wave))
    mel.train(was_training)
    gbs = B * 1.792e6 / (ms * 1e-3) / 1e9
    out["mel_kernel"] = {"bound": "hbm", "ms": ms, "achieved": gbs, "peak": peaks["hbm_gbs"], "unit": "GB/s",
                         "frac": gbs / peaks["hbm_gbs"],
                         "note": "fp32 FFT on CUDA cores: issue-bound (~4000 instructions per frame), DESIGN.md 5.3"}
    H, hd = net.num_heads, net.embed_dim // net.num_heads
    C = H * hd
    dev = wave.device
    qkv = torch.randn(B, ntok, 3 * C, device=dev).bfloat16()
    o = torch.empty(B, ntok, C, device=dev, dtype=torch.bfloat16)
    npad = ((ntok + 127) // 128) * 128
    lse = torch.empty(B, H, npad, device=dev)
    dO = torch.randn(B, ntok, C, device=dev).bfloat16()
    dqkv = torch.empty_like(qkv)
    ws = torch.empty(L.load().passt_attn_bwd_workspace_bytes(B, ntok, H), dtype=torch.uint8, device=dev)
    scale = hd ** -0.5
    ms_f = avg_ms(lambda: L.call("passt_attn_fwd", L.ptr(qkv), L.ptr(o), L.ptr(lse), B, ntok, H, scale, L.stream_ptr()))
    ms_b = avg_ms(lambda: L.call("passt_attn_bwd", L.ptr(qkv), L.ptr(o), L.ptr(dO), L.ptr(lse), L.ptr(dqkv), None,
                                 L.ptr(ws), B, ntok, H, scale, L.stream_ptr()))
    peak = peaks["bf16_sustained"]
    for name, ms, k in (("attn_fwd_kernel", ms_f, 4.0), ("attn_bwd (D pre-pass + kernel + dQ pack)", ms_b, 10.0)):
        tf = k * B * H * ntok * ntok * hd / (ms * 1e-3) / 1e12
        out[name] = {"bound": "tensor", "ms": ms, "achieved": tf, "peak": peak, "unit": "TFLOP/s", "frac": tf / peak,
                     "note": "hd=64: one exp2 per score caps the tensor pipe at 50 % (16 ex2/clk/SM), DESIGN.md 5.2"}
    return out


_REAL_STDOUT_FD = None


def _quiet_stdout():
    """Everything except the one JSON line goes to stderr -- including what C libraries write to fd 1 (NCCL prints
    its version banner there, the reference frontend prints its FMAX notice)."""
    global _REAL_STDOUT_FD
    sys.stdout.flush()
    _REAL_STDOUT_FD = os.dup(1)
    os.dup2(2, 1)


def _emit(line):
    sys.stdout.flush()
    if _REAL_STDOUT_FD is not None:
        os.write(_REAL_STDOUT_FD, (json.dumps(line) + "\n").encode())
    else:
        print(json.dumps(line), flush=True)


def main():
    _quiet_stdout()
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=30)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="candidate", choices=["candidate", "reference"])
    ap.add_argument("--optim", default="own", choices=["own", "torch"],
                    help="own: passt_b200.optim.FusedAdamW (default); torch: torch.optim.AdamW(fused=True)")
    ap.add_argument("--graph", type=int, default=1, help="1: replay the train step as one CUDA graph (default); 0: eager")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.impl == "reference":
        run_reference(args, rank, world)
        return
    if world > 1:
        import torch.distributed as dist
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        torch.cuda.set_device(local_rank)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    try:
        run_candidate(args, rank, local_rank, world)
    finally:
        if world > 1:
            import torch.distributed as dist
            dist.destroy_process_group()


if __name__ == "__main__":
    main()
