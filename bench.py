#!/usr/bin/env python
"""bench.py — audio-seconds/s of the Conformer-ASR encoder hot path (mel front-end → subsampling → Conformer layers,
forward + backward + gradient all-reduce + AdamW/Noam) on N B200s, for every BASELINE.json configuration.

    python bench.py [--gpus N] [--steps K] [--warmup W]              # configs[1], the driver's headline (default)
    python bench.py --config 3|4 [--scaling weak|strong] ...         # configs[2] (FastConformer-L) / configs[3] (Transducer-XL)
    python bench.py --config 5                                       # configs[4]: mel preprocessor sweep (HBM roofline)
    python bench.py --impl reference [--config C] [--steps K]        # the reference algorithm on the host CPU cores

  --config 2  Conformer-CTC-Large   18L d512  h8 k31 striding 4x  C=512   32 x 30 s per GPU (1 GPU: the quoted metric)
  --config 3  FastConformer-Large   17L d512  h8 k9  dw_striding 8x C=256 16 x 30 s per GPU (64 x 30 s over 4 GPUs)
  --config 4  Conformer-Transd.-XL  24L d1024 h8 k5  striding 4x  C=1024   2 x 30 s per GPU (16 x 30 s over 8 GPUs)
  --config 5  AudioToMelSpectrogramPreprocessor sweep: 1-60 s x batch 1-256, forward only

Prints ONE JSON line (contract in the task statement): value = whole-job audio-s/s with inputs resident in HBM;
e2e = same metric through the public module API with pinned-host inputs (H2D each step) and a D2H result read;
roofline = achieved TFLOP/s of the dominant kernel family (tcgen05 GEMM) from CUDA events vs the measured bf16 peak
(config 5: achieved algorithmic GB/s of the mel kernels vs the measured HBM peak); cpu_baseline = the fp32 oracle port on
this box's host cores on a bounded sample; gpu_eager_baseline = the reference op sequence (oracle port) on this B200 under
torch.autocast(bf16), materialised attention and `use_pytorch_sdpa` variants (SURVEY.md §8d "the number to beat").
"""
from __future__ import annotations

import argparse
import gc
import json
import os
import platform
import sys
import time

import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

SECS, SR = 30.0, 16000
CONFIGS = {
    2: dict(name="Conformer-CTC-Large", tag="18L d512 h8 k31 striding4x C512", per_gpu_batch=32, ref_gpus=1,
            model=dict(n_layers=18, d_model=512, n_heads=8, conv_kernel_size=31, subsampling_kind="striding",
                       subsampling_factor=4, subsampling_conv_channels=512)),
    3: dict(name="FastConformer-Large", tag="17L d512 h8 k9 dw_striding8x C256", per_gpu_batch=16, ref_gpus=4,
            model=dict(n_layers=17, d_model=512, n_heads=8, conv_kernel_size=9, subsampling_kind="dw_striding",
                       subsampling_factor=8, subsampling_conv_channels=256)),
    4: dict(name="Conformer-Transducer-XL", tag="24L d1024 h8 k5 striding4x C1024", per_gpu_batch=2, ref_gpus=8,
            model=dict(n_layers=24, d_model=1024, n_heads=8, conv_kernel_size=5, subsampling_kind="striding",
                       subsampling_factor=4, subsampling_conv_channels=1024)),
}
LARGE = CONFIGS[2]["model"]
METRIC = "audio-sec/s Conformer-L fwd+bwd"
METRICS = {2: METRIC, 3: "audio-sec/s FastConformer-L fwd+bwd", 4: "audio-sec/s Conformer-Transducer-XL encoder fwd+bwd",
           5: "audio-sec/s AudioToMelSpectrogramPreprocessor fwd"}
SWEEP_SECS = (1, 5, 10, 30, 60)
SWEEP_BATCH = (1, 8, 32, 128, 256)


def enc_frames(S, factor):
    t = S // 160 + 1
    for _ in range({4: 2, 8: 3}[factor]):
        t = (t - 1) // 2 + 1
    return t


def flops_fwd_bwd(B, T=None, d=512, L=18, h=8, dff=None, k=31, C=None, kind="striding", factor=4, S=int(SECS * SR), feat=80, **_):
    """dense-contraction MACs of the forward x 6 (fwd + bwd = 3 x 2 x MAC), any BASELINE config."""
    dff = dff or 4 * d
    C = C or d
    dk = d // h
    Ts, Fs = [S // 160 + 1], [feat]
    for _ in range({4: 2, 8: 3}[factor]):
        Ts.append((Ts[-1] - 1) // 2 + 1)
        Fs.append((Fs[-1] - 1) // 2 + 1)
    T = T or Ts[-1]
    if kind == "striding":
        sub = B * Ts[1] * Fs[1] * C * 9 + B * Ts[2] * Fs[2] * C * C * 9 + B * Ts[2] * Fs[2] * C * d
    else:
        sub = B * Ts[1] * Fs[1] * C * 9
        for i in (2, 3):
            sub += B * Ts[i] * Fs[i] * C * 9 + B * Ts[i] * Fs[i] * C * C
        sub += B * Ts[3] * Fs[3] * C * d
    per_layer = (4 * B * T * d * dff + 4 * B * T * d * d + (2 * T - 1) * d * d + B * h * T * T * dk
                 + B * h * T * (2 * T - 1) * dk + B * h * T * T * dk + 2 * B * T * d * d + B * T * d * k + B * T * d * d)
    return 6.0 * (sub + L * per_layer)


def model_flops(cfg_id, B):
    m = CONFIGS[cfg_id]["model"]
    return flops_fwd_bwd(B, d=m["d_model"], L=m["n_layers"], h=m["n_heads"], k=m["conv_kernel_size"],
                         C=m["subsampling_conv_channels"], kind=m["subsampling_kind"], factor=m["subsampling_factor"])


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            j = json.load(open(p))
            return dict(tflops=float(j["bf16_tflops_sustained"]), hbm=float(j["hbm_gbs"]), src="measured")
        except Exception:
            pass
    return dict(tflops=1590.0, hbm=6650.0, src="fallback")


def ncu_traffic(key):
    """DRAM bytes per launch (dram__bytes_read.sum + dram__bytes_write.sum) of the dominant kernel from the committed
    `ncu --set full` capture (profiles/dram_traffic.json, written by tools/summarize_profiles.py); None if absent."""
    p = os.path.join(ROOT, "profiles", "dram_traffic.json")
    try:
        return json.load(open(p)).get(key)
    except Exception:
        return None


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
def oracle_kwargs(m):
    return dict(n_layers=m["n_layers"], d_model=m["d_model"], n_heads=m["n_heads"], subsampling_kind=m["subsampling_kind"],
                subsampling_factor=m["subsampling_factor"])


def cpu_reference_step(O, sd, wav, wlen, m):
    """One fwd+bwd of the reference algorithm (fp32 oracle port, PyTorch eager on the host cores)."""
    params = {k: (v.clone().requires_grad_(True) if (v.is_floating_point() and "running" not in k) else v) for k, v in sd.items()}
    mel, mlen = O.mel_frontend(wav, wlen)
    y, _ = O.encoder_forward(params, mel, mlen, training=True, **oracle_kwargs(m))
    (y.float().pow(2).mean()).backward()


def host_threads():
    """Threads the CPU arm uses: the cores this process may run on, capped at 32 — PyTorch-eager on this path stops
    scaling well before that (128 threads measured 40x SLOWER than 8 on the pool's hosts: 0.13 vs 4.8 audio-s/s)."""
    try:
        n = len(os.sched_getaffinity(0))
    except Exception:
        n = os.cpu_count() or 1
    return max(1, min(n, 32))


def host_info():
    model = platform.processor() or ""
    try:
        for line in open("/proc/cpuinfo"):
            if line.startswith("model name"):
                model = line.split(":", 1)[1].strip()
                break
    except Exception:
        pass
    return {"cpu": model, "logical_cpus": os.cpu_count(), "threads_used": host_threads()}


def cpu_baseline(cfg_id=2, steps=2, warmup=1, batch=1, secs=SECS):
    """The reference algorithm on the host cores, 1 x 30 s utterance per step (B=1 x 30 s, scale linearly)."""
    from oracle import conformer_oracle as O
    torch.set_num_threads(host_threads())
    if cfg_id == 5:
        wav, wlen = O.synthetic_waveforms(32, secs)
        for _ in range(warmup):
            O.mel_frontend(wav, wlen)
        t0 = time.perf_counter()
        for _ in range(steps):
            O.mel_frontend(wav, wlen)
        dt = (time.perf_counter() - t0) / steps
        return {"value": 32 * secs / dt, "unit": "audio-sec/s", "cores": torch.get_num_threads(), "kind": "port",
                "sample": f"{steps} passes of 32x{int(secs)} s through the fp32 oracle port of FilterbankFeatures.forward "
                          f"(torch.stft + mel + log + normalise, PyTorch eager, {torch.get_num_threads()} threads)",
                "ms_per_step": dt * 1e3, "host": host_info()}
    m = CONFIGS[cfg_id]["model"]
    sd = O.init_state_dict(seed=0, **{k: v for k, v in m.items()})
    wav, wlen = O.synthetic_waveforms(batch, secs)
    for _ in range(warmup):
        cpu_reference_step(O, sd, wav, wlen, m)
    t0 = time.perf_counter()
    for _ in range(steps):
        cpu_reference_step(O, sd, wav, wlen, m)
    dt = (time.perf_counter() - t0) / steps
    return {"value": batch * secs / dt, "unit": "audio-sec/s", "cores": torch.get_num_threads(), "kind": "port",
            "sample": f"{steps} steps of {batch}x{int(secs)} s mel+fwd+bwd, {CONFIGS[cfg_id]['name']} (fp32 oracle port of the "
                      f"reference modules, PyTorch eager, {torch.get_num_threads()} threads)",
            "ms_per_step": dt * 1e3, "host": host_info()}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    c = args.config
    # bounded: each step is one 1x30 s utterance (≈2-4 s of CPU work for Large, ≈15 s for XL)
    steps = max(1, min(args.steps, 5 if c != 4 else 2))
    cb = cpu_baseline(cfg_id=c, steps=steps, warmup=1, batch=1)
    name = "AudioToMelSpectrogramPreprocessor" if c == 5 else f"{CONFIGS[c]['name']} {CONFIGS[c]['tag']}"
    out = {
        "impl": "reference", "metric": METRICS[c], "value": cb["value"], "unit": "audio-sec/s", "n_gpus": args.gpus,
        "steps": steps, "warmup": 1, "ms_per_step": cb["ms_per_step"], "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": f"{name} {'fwd' if c == 5 else 'fwd+bwd'}, 30 s utterances",
                   "sample": ("32x30 s per step" if c == 5 else "1x30 s utterance per step") + " on host CPU cores",
                   "host": cb["host"]},
        "cpu_baseline": {k: cb[k] for k in ("value", "unit", "cores", "kind", "sample")},
        "e2e": {"value": cb["value"], "unit": "audio-sec/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(out), flush=True)


# ------------------------------------------------------------------------------------------------ GPU eager baseline
def gpu_eager_baseline(cfg_id, dev, batch_try=(16, 8, 4, 2, 1), steps=2):
    """SURVEY.md §8(d) secondary baseline: the reference op sequence (oracle port: F.conv2d / F.linear / matmul /
    softmax / layer_norm ... exactly the reference's torch calls) on THIS GPU under torch.autocast(bf16), fwd+bwd, default
    materialised attention and the `use_pytorch_sdpa=True` branch.  Batch is the largest of batch_try that fits (the
    materialised [B,h,T,2T-1] score tensors of 18-24 layers do not fit 32 x 30 s in 180 GB); audio-s/s is per-GPU."""
    from oracle import conformer_oracle as O
    m = CONFIGS[cfg_id]["model"]
    sd = {k: v.to(dev) for k, v in O.init_state_dict(seed=0, **m).items()}
    out = {"unit": "audio-sec/s", "what": "oracle port of the reference modules, torch eager on this GPU, autocast bf16, fwd+bwd"}

    def one(B, sdpa):
        wav, wlen = O.synthetic_waveforms(B, SECS)
        wav, wlen = wav.to(dev), wlen.to(dev)

        def step():
            params = {k: (v.detach().requires_grad_(True) if (v.is_floating_point() and "running" not in k) else v) for k, v in sd.items()}
            with torch.no_grad():
                mel, mlen = O.mel_frontend(wav, wlen)
            with torch.autocast("cuda", dtype=torch.bfloat16):
                y, _ = O.encoder_forward(params, mel, mlen, training=True, use_sdpa=sdpa, **oracle_kwargs(m))
            y.float().pow(2).mean().backward()

        step()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            step()
        e1.record()
        torch.cuda.synchronize()
        return B * SECS * steps / (e0.elapsed_time(e1) / 1e3)

    for sdpa in (False, True):
        key = "sdpa" if sdpa else "materialised"
        for B in batch_try:
            try:
                out[key] = {"value": one(B, sdpa), "batch": B}
                break
            except torch.OutOfMemoryError:
                gc.collect()
                torch.cuda.empty_cache()
            except Exception as e:  # noqa: BLE001
                out[key] = {"error": repr(e)[:200]}
                break
        gc.collect()
        torch.cuda.empty_cache()
    vals = [v["value"] for v in out.values() if isinstance(v, dict) and "value" in v]
    out["value"] = max(vals) if vals else None
    return out


# ------------------------------------------------------------------------------------------------ config 5: mel sweep
def run_mel_sweep(args):
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (no CPU fallback); use --impl reference for the CPU arm")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    import torch.distributed as dist
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    from nemo_b200 import _lib as L
    from nemo_b200.preprocessor import B200MelPreprocessor

    pre = B200MelPreprocessor(sample_rate=SR, window_size=0.025, window_stride=0.01, features=80, n_fft=512, dither=1e-5,
                              pad_to=0).to(dev).eval()
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)  # > 126 MB L2
    peaks = measured_peaks()
    g = torch.Generator().manual_seed(1234 + rank)

    def point(B, secs, iters):
        S = int(secs * SR)
        wav = (torch.rand(B, S, generator=g) * 0.2 - 0.1).to(dev)
        ln = torch.full((B,), S, dtype=torch.int64, device=dev)
        for _ in range(3):
            pre(input_signal=wav, length=ln)
        ts = []
        for _ in range(iters):
            flush.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            pre(input_signal=wav, length=ln)
            e1.record()
            e1.synchronize()
            ts.append(e0.elapsed_time(e1))
        ts.sort()
        ms = ts[len(ts) // 2]
        T = S // 160 + 1
        nbytes = B * (4 * S + 4 * 80 * T)  # 4·S in + 4·80·T_mel out per utterance = 96 kB / audio-s
        return ms, nbytes

    table = []
    for secs in SWEEP_SECS:
        for B in SWEEP_BATCH:
            ms, nbytes = point(B, secs, 7)
            table.append({"batch": B, "secs": secs, "ms": round(ms, 4), "audio_s_per_s": B * secs / (ms / 1e3),
                          "gbs": nbytes / (ms / 1e3) / 1e9, "frac": nbytes / (ms / 1e3) / 1e9 / peaks["hbm"]})
    # headline point = the batch of config 2 (32 x 30 s), timed like every other config: W warm-up, K steps, one region
    B, S = 32, int(SECS * SR)
    wav_host = (torch.rand(B, S, generator=g) * 0.2 - 0.1).pin_memory()
    len_host = torch.full((B,), S, dtype=torch.int64).pin_memory()
    wav_dev, len_dev = wav_host.to(dev), len_host.to(dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(n, from_host):
        barrier()
        n0 = L.lib.b2c_launch_count()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(n):
            if from_host:
                w, l_ = wav_host.to(dev, non_blocking=True), len_host.to(dev, non_blocking=True)
                mel, mlen = pre(input_signal=w, length=l_)
                _ = (mlen.sum() + mel[:, :, 0].sum().long()).item()  # D2H read of the result lengths + a checksum word
            else:
                flush.zero_()
                pre(input_signal=wav_dev, length=len_dev)
        e1.record()
        barrier()
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return ms.item(), L.lib.b2c_launch_count() - n0

    clocks = ClockSampler(local) if rank == 0 else None
    n_warm = max(args.warmup, 3)
    steps = max(args.steps, 20)
    timed(n_warm, False)
    # the flush write is inside the loop: time it alone and subtract (it is not part of the path)
    barrier()
    f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    f0.record()
    for _ in range(steps):
        flush.zero_()
    f1.record()
    torch.cuda.synchronize()
    flush_ms = f0.elapsed_time(f1)
    ms, launches = timed(steps, False)
    if clocks is not None:
        clocks.sample()
    ms = max(ms - flush_ms, 1e-6)
    ms_e2e, _ = timed(steps, True)
    if rank == 0:
        T = S // 160 + 1
        nbytes = B * (4 * S + 4 * 80 * T)
        audio_s = world * B * SECS
        gbs = nbytes / (ms / steps / 1e3) / 1e9
        out = {
            "metric": METRICS[5], "value": audio_s * steps / (ms / 1e3), "unit": "audio-sec/s", "n_gpus": world, "steps": steps,
            "warmup": n_warm, "ms_per_step": ms / steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": {"workload": "AudioToMelSpectrogramPreprocessor (dither-free eval: preemph, STFT 512/400/160, power, 80 mel, "
                                   "log, per-feature normalise, mask), 32x30 s per GPU; sweep 1-60 s x batch 1-256 in `sweep`",
                       "global_batch": world * B, "parallelism": f"dp{world}",
                       "l2": "256 MB flush write between iterations (its own time measured and subtracted)", "sweep": table},
            "e2e": {"value": audio_s * steps / (ms_e2e / 1e3), "unit": "audio-sec/s",
                    "h2d_bytes_per_step": world * (wav_host.numel() * 4 + B * 8), "d2h_bytes_per_step": world * 8},
            "gpu_launches": launches, "clocks": clocks.result(),
            "roofline": {"bound": "hbm", "kernel": "mel_frames_kernel + mel_normalize_kernel (one pass of the front-end)",
                         "achieved": gbs, "peak": peaks["hbm"], "unit": "GB/s", "frac": gbs / peaks["hbm"],
                         "traffic": ncu_traffic("mel_frontend_32x30s"), "peak_source": peaks["src"] + " hbm_gbs",
                         "algorithmic_bytes_per_launch": nbytes},
        }
        if world == 1 and not args.no_cpu_baseline:
            cb = cpu_baseline(cfg_id=5, steps=2, warmup=1)
            out["cpu_baseline"] = {k: cb[k] for k in ("value", "unit", "cores", "kind", "sample")}
        print(json.dumps(out), flush=True)
    if world > 1:
        _teardown(dist)


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

    C = CONFIGS[args.config]
    m = C["model"]
    if args.batch is not None:
        B = args.batch
    elif args.scaling == "strong":
        gb = C["per_gpu_batch"] * C["ref_gpus"]  # BASELINE's global batch, split over however many GPUs there are
        if gb % world:
            raise SystemExit(f"strong scaling: global batch {gb} is not divisible by {world} GPUs")
        B = gb // world
    else:
        B = C["per_gpu_batch"]
    torch.manual_seed(0)  # identical replicas
    enc = B200ConformerEncoder(feat_in=80, n_layers=m["n_layers"], d_model=m["d_model"], n_heads=m["n_heads"],
                               conv_kernel_size=m["conv_kernel_size"], subsampling=m["subsampling_kind"],
                               subsampling_factor=m["subsampling_factor"], subsampling_conv_channels=m["subsampling_conv_channels"],
                               dropout=0.1, dropout_pre_encoder=0.1, dropout_emb=0.0, dropout_att=0.1).to(dev).train()
    with torch.no_grad():
        for lyr in enc.layers:
            lyr.self_attn.pos_bias_u.normal_(0, 0.02)
            lyr.self_attn.pos_bias_v.normal_(0, 0.02)
    pre = B200MelPreprocessor(sample_rate=SR, window_size=0.025, window_stride=0.01, features=80, n_fft=512, dither=1e-5,
                              pad_to=0).to(dev).train()
    from nemo_b200.graph import CapturedStep
    from nemo_b200.optim import B200FusedAdamW
    fg = FlatGradients(enc)
    # the recipe's optimizer (conformer_ctc_bpe.yaml:175-192): AdamW lr 2.0, betas (0.9, 0.98), wd 1e-3 + NoamAnnealing
    opt = B200FusedAdamW(fg, lr=2.0, betas=(0.9, 0.98), weight_decay=1e-3, noam_d_model=m["d_model"], noam_warmup_steps=10000,
                         noam_min_lr=1e-6)
    S = int(SECS * SR)
    g = torch.Generator().manual_seed(1234 + rank)  # every rank its own shard of the global batch
    wav_host = (torch.rand(B, S, generator=g) * 0.2 - 0.1).pin_memory()
    len_host = torch.full((B,), S, dtype=torch.int64).pin_memory()
    wav_dev, len_dev = wav_host.to(dev), len_host.to(dev)
    T_enc = enc_frames(S, m["subsampling_factor"])
    proj = torch.randn(B, m["d_model"], T_enc, device=dev) / (B * T_enc * m["d_model"])
    loss_scale = fg.loss_scale()

    gemm_events = []
    orig_gemm = L.gemm

    def timed_gemm(A, Bm, M, N, K, *a, **kw):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        orig_gemm(A, Bm, M, N, K, *a, **kw)
        e1.record()
        gemm_events.append((2.0 * M * N * K, e0, e1))

    def core_step(wav, wlen):
        opt.zero_grad()  # free after the first step: B200FusedAdamW zeroes each gradient element as it consumes it
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

    # The whole step as ONE CUDA graph (nemo_b200.graph.CapturedStep): static device input buffers, device-side RNG counter
    # and optimizer step count.  --no-graph (or a failed capture) falls back to eager enqueueing of the same kernels.
    captured, graph_note = None, "off (--no-graph)"
    wav_static, len_static = wav_dev.clone(), len_dev.clone()
    if not args.no_graph:
        try:
            captured = CapturedStep(lambda: core_step(wav_static, len_static), warmup=3)
            graph_note = "on"
        except Exception as e:  # noqa: BLE001
            graph_note = f"capture failed, eager fallback: {repr(e)[:160]}"
            captured = None
            torch.cuda.synchronize()

    def step(wav, wlen, from_host, eager=False):
        if captured is not None and not eager:
            if from_host or wav is not wav_static:
                wav_static.copy_(wav, non_blocking=True)  # H2D from pinned memory (e2e) or D2D refresh of the static inputs
                len_static.copy_(wlen, non_blocking=True)
            return captured()
        if from_host:
            wav, wlen = wav.to(dev, non_blocking=True), wlen.to(dev, non_blocking=True)
        return core_step(wav, wlen)

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
        if profile_gemm:
            # the per-GEMM events must bracket GPU execution only: park the GPU for ~100 ms so that the host enqueues the whole
            # (eager) step ahead of it — at configs 3 / 4 the host is slower than the GPU and every event pair would otherwise
            # include the host's tensor-map encoding + launch latency of the call it brackets
            torch.cuda._sleep(int(0.10 * 1.9e9))
        e0.record()
        for _ in range(n):
            loss = step(wav_host if from_host else wav_static, len_host if from_host else len_static, from_host, eager=profile_gemm)
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
        launched = L.lib.b2c_launch_count() - launches0
        if captured is not None and not profile_gemm:
            launched += n * captured.launches_per_replay  # kernels inside the replayed graph (counted once, at capture)
        return ms.item(), launched

    # NVML is initialised (and queried once) BEFORE the warm-up: on some pool hosts the first NVML calls of a process
    # stall kernel submission for 100-300 ms a moment later — inside the first timed steps if they come after warm-up
    clocks = ClockSampler(local) if rank == 0 else None
    if clocks is not None:
        clocks.sample(n=1)
        clocks.rows.clear()
        clocks.cost.clear()
    n_warm = max(args.warmup, 3)
    for _ in range(n_warm):
        step(wav_static, len_static, False)
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
    timed(1, False, profile_gemm=True)  # eager enqueue of the same kernels (events cannot bracket nodes of a replayed graph)
    torch.cuda.synchronize()
    gf = sum(f for f, _, _ in gemm_events)
    gms = sum(a.elapsed_time(b) for _, a, b in gemm_events)

    if rank == 0:
        peaks = measured_peaks()
        audio_s = world * B * SECS
        value = audio_s * args.steps / (ms / 1e3)
        e2e = audio_s * args.steps / (ms_e2e / 1e3)
        achieved = gf / (gms / 1e3) / 1e12 if gms > 0 else 0.0
        step_flops = model_flops(args.config, B)
        out = {
            "metric": METRICS[args.config], "value": value, "unit": "audio-sec/s", "n_gpus": world, "steps": args.steps,
            "warmup": n_warm, "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": args.scaling,
            "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
            "config": {"workload": f"{C['name']} {C['tag']}, {B}x30 s per GPU, mel+fwd+bwd+allreduce+AdamW",
                       "global_batch": world * B, "parallelism": f"dp{world}", "dropout": "0.1/0.1/att0.1",
                       "cuda_graph": graph_note, "optimizer": "B200FusedAdamW lr 2.0 betas (0.9,0.98) wd 1e-3 + Noam(d_model, 10000, 1e-6)",
                       "l2": "inputs+activations per step (GBs) exceed the 126 MB L2; no flush needed",
                       "step_ms": step_ms_value, "host_enqueue_ms": host_ms_value, "step_ms_e2e": step_ms_e2e,
                       "model_tflop_per_step_per_gpu": step_flops / 1e12,
                       "model_tflops_achieved_per_gpu": step_flops / (ms / args.steps / 1e3) / 1e12,
                       "model_flop_utilisation": step_flops / (ms / args.steps / 1e3) / 1e12 / peaks["tflops"]},
            "e2e": {"value": e2e, "unit": "audio-sec/s", "h2d_bytes_per_step": world * (wav_host.numel() * 4 + B * 8),
                    "d2h_bytes_per_step": world * 4},
            "gpu_launches": launches,
            "clocks": clk,
            "roofline": {"bound": "tensor", "kernel": "gemm_tc_kernel (tcgen05 GEMM family, all launches of one step)",
                         "achieved": achieved, "peak": peaks["tflops"], "unit": "TFLOP/s", "frac": achieved / peaks["tflops"],
                         "traffic": ncu_traffic(f"gemm_family_cfg{args.config}"), "peak_source": peaks["src"] + " bf16_tflops_sustained",
                         "gemm_ms_per_step": gms, "gemm_share_of_step": gms / (ms / args.steps),
                         "gemm_flop_per_step": gf, "gemm_launches_per_step": len(gemm_events)},
        }
        if world == 1 and not args.no_cpu_baseline:
            cb = cpu_baseline(cfg_id=args.config, steps=1 if args.config == 4 else 2, warmup=1, batch=1)
            out["cpu_baseline"] = {k: cb[k] for k in ("value", "unit", "cores", "kind", "sample")}
            out["config"]["host"] = cb["host"]
        if world == 1 and not args.no_eager_baseline:
            del enc, opt, fg, captured
            gc.collect()
            torch.cuda.empty_cache()
            out["gpu_eager_baseline"] = gpu_eager_baseline(args.config, dev)
        print(json.dumps(out), flush=True)
    if world > 1:
        _teardown(dist)


def _teardown(dist):
    """Leave a multi-rank run without the NCCL / CUDA-graph teardown hang: a process group that was used inside a captured graph
    can block in destroy_process_group() or at interpreter exit (observed: every rank idle until the launcher's timeout).  All
    ranks meet at a barrier, flush, and exit hard; a watchdog covers the barrier itself."""
    import threading

    threading.Timer(30.0, lambda: os._exit(0)).start()
    try:
        dist.barrier()
        torch.cuda.synchronize()
    except Exception:  # noqa: BLE001
        pass
    sys.stdout.flush()
    sys.stderr.flush()
    os._exit(0)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", type=int, default=2, choices=[2, 3, 4, 5], help="BASELINE.json configs[config-1] (default 2: Conformer-CTC-Large)")
    ap.add_argument("--scaling", default="weak", choices=["weak", "strong"],
                    help="weak: the config's per-GPU batch on every GPU; strong: BASELINE's global batch split over the GPUs")
    ap.add_argument("--batch", type=int, default=None, help="utterances per GPU (overrides the config's)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-eager-baseline", action="store_true")
    ap.add_argument("--no-graph", action="store_true", help="enqueue the step's kernels eagerly instead of replaying a CUDA graph")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    elif args.config == 5:
        run_mel_sweep(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
