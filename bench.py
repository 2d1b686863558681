#!/usr/bin/env python
"""bench.py — audio-seconds/s of the Conformer-ASR encoder hot path (mel front-end → subsampling → 18 Conformer
layers, forward + backward + gradient all-reduce + AdamW) on N B200s.

Workload = BASELINE.json configs[1]: Conformer-CTC-Large (18L, d=512, 8 heads, conv k=31, striding 4x), fwd+bwd,
32 x 30 s synthetic 16 kHz utterances PER GPU (weak scaling), bf16 tensor-core arithmetic with fp32 master weights,
dropout active (0.1 / 0.1 / att 0.1 as in examples/asr/conf/conformer/conformer_ctc_bpe.yaml), random-init weights.

    python bench.py [--gpus N] [--steps K] [--warmup W]            # ours (N>1: launched under torchrun)
    python bench.py --impl reference [--steps K] [--warmup W]      # the reference algorithm on the host CPU cores

Prints ONE JSON line (contract in the task statement): value = whole-job audio-s/s with inputs resident in HBM;
e2e = same metric through the public module API with pinned-host inputs (H2D each step) and a D2H loss read;
roofline = achieved TFLOP/s of the dominant kernel family (tcgen05 GEMM) from CUDA events, vs the measured bf16 peak;
cpu_baseline = the fp32 oracle port timed on this box's host cores on a bounded sample.
"""
from __future__ import annotations

import argparse
import gc
import json
import os
import subprocess
import sys
import threading
import time

import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

LARGE = dict(n_layers=18, d_model=512, n_heads=8, conv_kernel_size=31, subsampling_kind="striding", subsampling_factor=4)
SECS, SR, PER_GPU_BATCH = 30.0, 16000, 32
METRIC = "audio-sec/s Conformer-L fwd+bwd"


def flops_fwd_bwd(B, T, d=512, L=18, h=8, dff=2048, k=31, C=512, F1=40, F2=20, T1=1501):
    """dense-contraction MACs of the forward x 6 (fwd + bwd = 3 x 2 x MAC)."""
    dk = d // h
    sub = B * T1 * F1 * C * 9 + B * T * F2 * C * C * 9 + B * T * F2 * C * d
    per_layer = (4 * B * T * d * dff + 4 * B * T * d * d + (2 * T - 1) * d * d + B * h * T * T * dk
                 + B * h * T * (2 * T - 1) * dk + B * h * T * T * dk + 2 * B * T * d * d + B * T * d * k + B * T * d * d)
    return 6.0 * (sub + L * per_layer)


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            j = json.load(open(p))
            return dict(tflops=float(j["bf16_tflops_sustained"]), hbm=float(j["hbm_gbs"]), src="measured")
        except Exception:
            pass
    return dict(tflops=1590.0, hbm=6650.0, src="fallback")


class ClockSampler:
    """SM clock + throttle reasons DURING the timed region (B200_PROFILING.md), read through NVML from the main thread
    at the one moment it cannot perturb the measurement: right after the last timed step has been ENQUEUED, while the
    GPU is still executing the queued work of the region.  (NVML queries hold a driver lock that blocks kernel
    submission — up to ~150 ms per query on some pool hosts — so the recipe's `nvidia-smi -lms 200`, or a 5 Hz sampler
    thread, stretched single steps from 87 ms to 245 ms.)"""

    REASONS = {0x8: "hw_slowdown", 0x40: "hw_thermal_slowdown", 0x20: "sw_thermal_slowdown", 0x4: "sw_power_cap"}

    def __init__(self, index=0):
        self.index, self.rows, self.cost, self.err, self.h, self.nv = index, [], [], None, None, None
        try:
            import pynvml

            pynvml.nvmlInit()
            vis = os.environ.get("CUDA_VISIBLE_DEVICES")
            idx = int(vis.split(",")[index]) if vis and index < len(vis.split(",")) and vis.split(",")[index].isdigit() else index
            self.h = pynvml.nvmlDeviceGetHandleByIndex(idx)
            self.nv = pynvml
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception as e:  # noqa: BLE001
            self.err = repr(e)

    def sample(self, n=3, gap=0.01):
        if self.nv is None:
            return
        for _ in range(n):
            try:
                t0 = time.perf_counter()
                mhz = self.nv.nvmlDeviceGetClockInfo(self.h, self.nv.NVML_CLOCK_SM)
                try:
                    rs = self.nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                except Exception:
                    rs = self.nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                self.rows.append((mhz, rs))
                self.cost.append((time.perf_counter() - t0) * 1e3)
            except Exception as e:  # noqa: BLE001
                self.err = repr(e)
            time.sleep(gap)

    def result(self):
        if not self.rows:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [f"nvml unavailable: {self.err}"]}
        sm = sorted(r[0] for r in self.rows)
        reasons = set()
        for _, rs in self.rows:
            for bit, name in self.REASONS.items():
                if rs & bit:
                    reasons.add(name)
        return {"sm_mhz": float(sm[len(sm) // 2]), "sm_max_mhz": float(self.max_mhz), "reasons": sorted(reasons),
                "samples": len(sm), "nvml_ms": round(max(self.cost), 2) if self.cost else None,
                "when": "after the last timed step was enqueued, GPU still executing the region"}


# ------------------------------------------------------------------------------------------------ reference / CPU arm
def cpu_reference_step(O, sd, wav, wlen, cfg):
    """One fwd+bwd of the reference algorithm (fp32 oracle port, PyTorch eager on the host cores)."""
    params = {k: (v.clone().requires_grad_(True) if (v.is_floating_point() and "running" not in k) else v) for k, v in sd.items()}
    mel, mlen = O.mel_frontend(wav, wlen)
    y, _ = O.encoder_forward(params, mel, mlen, n_layers=cfg["n_layers"], d_model=cfg["d_model"], n_heads=cfg["n_heads"],
                             subsampling_kind=cfg["subsampling_kind"], subsampling_factor=cfg["subsampling_factor"],
                             training=True)
    (y.float().pow(2).mean()).backward()


def host_threads():
    """Threads the CPU arm uses: the cores this process may run on, capped at 32 — PyTorch-eager on this path stops
    scaling well before that (128 threads measured 40x SLOWER than 8 on the pool's hosts: 0.13 vs 4.8 audio-s/s)."""
    try:
        n = len(os.sched_getaffinity(0))
    except Exception:
        n = os.cpu_count() or 1
    return max(1, min(n, 32))


def cpu_baseline(steps=2, warmup=1, batch=1, secs=10.0):
    from oracle import conformer_oracle as O
    torch.set_num_threads(host_threads())
    sd = O.init_state_dict(seed=0, **LARGE)
    wav, wlen = O.synthetic_waveforms(batch, secs)
    for _ in range(warmup):
        cpu_reference_step(O, sd, wav, wlen, LARGE)
    t0 = time.perf_counter()
    for _ in range(steps):
        cpu_reference_step(O, sd, wav, wlen, LARGE)
    dt = (time.perf_counter() - t0) / steps
    return {"value": batch * secs / dt, "unit": "audio-sec/s", "cores": torch.get_num_threads(), "kind": "port",
            "sample": f"{steps} steps of {batch}x{int(secs)} s mel+fwd+bwd, Conformer-CTC-Large (fp32 oracle port of the reference "
                      f"modules, PyTorch eager, {torch.get_num_threads()} threads)",
            "ms_per_step": dt * 1e3}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cb = cpu_baseline(steps=max(1, args.steps), warmup=max(1, min(args.warmup, 1)), batch=1)
    out = {
        "impl": "reference", "metric": METRIC, "value": cb["value"], "unit": "audio-sec/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": cb["ms_per_step"], "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": "Conformer-CTC-Large 18L d512 h8 k31 striding4x fwd+bwd, 30 s utterances",
                   "sample": "1x10 s utterance per step on host CPU cores"},
        "cpu_baseline": {k: cb[k] for k in ("value", "unit", "cores", "kind", "sample")},
        "e2e": {"value": cb["value"], "unit": "audio-sec/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(out), flush=True)


# ------------------------------------------------------------------------------------------------ ours
def run_ours(args):
    import torch.distributed as dist

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (no CPU fallback); use --impl reference for the CPU arm")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)

    from nemo_b200 import _lib as L
    from nemo_b200.encoder import B200ConformerEncoder
    from nemo_b200.parallel import FlatGradients
    from nemo_b200.preprocessor import B200MelPreprocessor

    B = args.batch
    torch.manual_seed(0)  # identical replicas
    enc = B200ConformerEncoder(feat_in=80, n_layers=LARGE["n_layers"], d_model=LARGE["d_model"], n_heads=LARGE["n_heads"],
                               conv_kernel_size=LARGE["conv_kernel_size"], subsampling="striding", subsampling_factor=4,
                               dropout=0.1, dropout_pre_encoder=0.1, dropout_emb=0.0, dropout_att=0.1).to(dev).train()
    with torch.no_grad():
        for lyr in enc.layers:
            lyr.self_attn.pos_bias_u.normal_(0, 0.02)
            lyr.self_attn.pos_bias_v.normal_(0, 0.02)
    pre = B200MelPreprocessor(sample_rate=SR, window_size=0.025, window_stride=0.01, features=80, n_fft=512, dither=1e-5,
                              pad_to=0).to(dev).train()
    fg = FlatGradients(enc)
    opt = torch.optim.AdamW(enc.parameters(), lr=1e-4, betas=(0.9, 0.98), weight_decay=1e-3, fused=True)
    S = int(SECS * SR)
    g = torch.Generator().manual_seed(1234 + rank)  # every rank its own shard of the global batch
    wav_host = (torch.rand(B, S, generator=g) * 0.2 - 0.1).pin_memory()
    len_host = torch.full((B,), S, dtype=torch.int64).pin_memory()
    wav_dev, len_dev = wav_host.to(dev), len_host.to(dev)
    T_enc = ((S // 160 + 1 - 1) // 2 + 1 - 1) // 2 + 1
    proj = torch.randn(B, LARGE["d_model"], T_enc, device=dev) / (B * T_enc * LARGE["d_model"])
    loss_scale = fg.loss_scale()

    gemm_events = []
    orig_gemm = L.gemm

    def timed_gemm(A, Bm, M, N, K, *a, **kw):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        orig_gemm(A, Bm, M, N, K, *a, **kw)
        e1.record()
        gemm_events.append((2.0 * M * N * K, e0, e1))

    def step(wav, wlen, from_host):
        if from_host:
            wav, wlen = wav.to(dev, non_blocking=True), wlen.to(dev, non_blocking=True)
        fg.zero()
        mel, mlen = pre(input_signal=wav, length=wlen)
        y, _ = enc(audio_signal=mel, length=mlen)
        loss = (y * proj).sum() * loss_scale
        loss.backward()
        fg.finish()
        opt.step()
        # detached: a live `loss` keeps the autograd node — and with it this step's ~17 GB activation arena — alive into
        # the next step.  That doubled the resident activations, and the first time it happened (timed step 2: the warm-up
        # loop discards the return value) the caching allocator spent ~250 ms in device-synchronising cudaMallocs.
        return loss.detach()

    per_step = []
    host_step = []
    in_flight = []

    def throttle(ev):
        """The host may run at most two steps ahead of the GPU (the queue never drains, so nothing is lost); it bounds the
        launch queue and the memory held by in-flight steps."""
        in_flight.append(ev)
        if len(in_flight) > 2:
            in_flight.pop(0).synchronize()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(n, from_host, profile_gemm=False, sampler=None):
        # manual garbage collection (cf. nemo/lightning/pytorch/callbacks/garbage_collection.py): a generational sweep in the
        # middle of a step stalls kernel submission (one timed step in ~10 showed a >100 ms host-side gap without this)
        gc.collect()
        gc.disable()
        try:
            return _timed(n, from_host, profile_gemm, sampler)
        finally:
            gc.enable()

    def _timed(n, from_host, profile_gemm, sampler):
        barrier()
        if profile_gemm:
            import nemo_b200.encoder as E
            L.gemm = timed_gemm
            E.L.gemm = timed_gemm
        launches0 = L.lib.b2c_launch_count()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        marks = []
        host_marks = [time.perf_counter()]
        e0.record()
        for _ in range(n):
            loss = step(wav_host if from_host else wav_dev, len_host if from_host else len_dev, from_host)
            if from_host:
                _ = loss.item()  # device→host read of the step's result
            marks.append(torch.cuda.Event(enable_timing=True))
            marks[-1].record()
            host_marks.append(time.perf_counter())
            throttle(marks[-1])
        e1.record()
        if sampler is not None:
            sampler.sample()  # GPU is still busy with the enqueued steps; nothing is left to submit
        barrier()
        per_step.clear()
        host_step.clear()
        host_step.extend(round((b_ - a_) * 1e3, 1) for a_, b_ in zip(host_marks[:-1], host_marks[1:]))
        prev = e0
        for mk in marks:
            per_step.append(round(prev.elapsed_time(mk), 2))
            prev = mk
        if profile_gemm:
            import nemo_b200.encoder as E
            L.gemm = orig_gemm
            E.L.gemm = orig_gemm
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return ms.item(), L.lib.b2c_launch_count() - launches0

    # NVML is initialised (and queried once) BEFORE the warm-up: on some pool hosts the first NVML calls of a process
    # stall kernel submission for 100-300 ms a moment later — inside the first timed steps if they come after warm-up
    clocks = ClockSampler(local) if rank == 0 else None
    if clocks is not None:
        clocks.sample(n=1)
        clocks.rows.clear()
        clocks.cost.clear()
    n_warm = max(args.warmup, 3)
    for _ in range(n_warm):
        step(wav_dev, len_dev, False)
        ev = torch.cuda.Event()
        ev.record()
        throttle(ev)
    ms, launches = timed(args.steps, False, sampler=clocks)
    clk = clocks.result() if rank == 0 else None
    step_ms_value = list(per_step)
    host_ms_value = list(host_step)
    ms_e2e, _ = timed(args.steps, True)
    step_ms_e2e = list(per_step)
    # dominant-kernel roofline: every tcgen05 GEMM launch of one extra step bracketed by CUDA events
    gemm_events.clear()
    timed(1, False, profile_gemm=True)
    torch.cuda.synchronize()
    gf = sum(f for f, _, _ in gemm_events)
    gms = sum(a.elapsed_time(b) for _, a, b in gemm_events)

    if rank == 0:
        peaks = measured_peaks()
        audio_s = world * B * SECS
        value = audio_s * args.steps / (ms / 1e3)
        e2e = audio_s * args.steps / (ms_e2e / 1e3)
        achieved = gf / (gms / 1e3) / 1e12 if gms > 0 else 0.0
        step_flops = flops_fwd_bwd(B, T_enc)
        out = {
            "metric": METRIC, "value": value, "unit": "audio-sec/s", "n_gpus": world, "steps": args.steps,
            "warmup": n_warm, "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
            "config": {"workload": f"Conformer-CTC-Large 18L d512 h8 k31 striding4x, {B}x30 s per GPU, mel+fwd+bwd+allreduce+AdamW",
                       "global_batch": world * B, "parallelism": f"dp{world}", "dropout": "0.1/0.1/att0.1",
                       "l2": "inputs+activations per step (>17 GB) exceed the 126 MB L2; no flush needed",
                       "step_ms": step_ms_value, "host_enqueue_ms": host_ms_value, "step_ms_e2e": step_ms_e2e,
                       "model_tflop_per_step_per_gpu": step_flops / 1e12,
                       "model_tflops_achieved_per_gpu": step_flops / (ms / args.steps / 1e3) / 1e12},
            "e2e": {"value": e2e, "unit": "audio-sec/s", "h2d_bytes_per_step": world * (wav_host.numel() * 4 + B * 8),
                    "d2h_bytes_per_step": world * 4},
            "gpu_launches": launches,
            "clocks": clk,
            "roofline": {"bound": "tensor", "kernel": "gemm_tc_kernel (tcgen05 GEMM family, all launches of one step)",
                         "achieved": achieved, "peak": peaks["tflops"], "unit": "TFLOP/s", "frac": achieved / peaks["tflops"],
                         "traffic": None, "peak_source": peaks["src"] + " bf16_tflops_sustained",
                         "gemm_ms_per_step": gms, "gemm_share_of_step": gms / (ms / args.steps)},
        }
        if world == 1 and not args.no_cpu_baseline:
            cb = cpu_baseline(steps=2, warmup=1, batch=1)
            out["cpu_baseline"] = {k: cb[k] for k in ("value", "unit", "cores", "kind", "sample")}
        print(json.dumps(out), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--batch", type=int, default=PER_GPU_BATCH, help="utterances per GPU (default 32 = BASELINE config 2)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
