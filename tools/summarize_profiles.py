"""Turns the scratch ncu outputs under gpurun_out/ into the tracked summaries under profiles/ (named per round).
  python tools/summarize_profiles.py r01"""
import collections
import csv
import json
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tag = sys.argv[1] if len(sys.argv) > 1 else "r01"
out_dir = os.path.join(ROOT, "profiles")
os.makedirs(out_dir, exist_ok=True)
G = os.path.join(ROOT, "gpurun_out")


def launches():
    p = os.path.join(G, "launches.csv")
    if not os.path.exists(p):
        return
    lines = [l for l in open(p) if l.startswith('"')]
    r = csv.reader(lines)
    hdr = next(r)
    ix = {h: i for i, h in enumerate(hdr)}
    tot, cnt = collections.defaultdict(float), collections.Counter()
    for row in r:
        try:
            v = float(row[ix["Metric Value"]].replace(",", ""))
        except Exception:
            continue
        u = row[ix["Metric Unit"]]
        v = v / 1e3 if u == "ns" else (v * 1e3 if u == "ms" else v)
        name = row[ix["Kernel Name"]]
        m = re.search(r"gemm_tc_kernel<(.*?)>", name)
        key = f"gemm_tc_kernel<{m.group(1)}>" if m else re.sub(r"<.*", "", name).replace("void ", "")[:80]
        tot[key] += v
        cnt[key] += 1
    T = sum(tot.values())
    with open(os.path.join(out_dir, f"{tag}_launches.md"), "w") as f:
        f.write(f"# {tag}: ncu launch list of `bench.py --steps 1 --warmup 1` (first 6000 launches ≈ 3 fwd + 2.7 bwd steps, "
                "32x30 s, Conformer-CTC-Large)\n\n")
        f.write("`ncu --metrics gpu__time_duration.sum --clock-control none -c 6000` — per-launch times are cold-cache and "
                "serialised: compare SHARES.  gemm_tc_kernel<BN, A_MN, B_MN, EPILOGUE, CONV>.\n\n")
        f.write("| kernel | launches | total ms | share | avg us |\n|---|---:|---:|---:|---:|\n")
        for k, v in sorted(tot.items(), key=lambda x: -x[1]):
            if v / T < 0.001:
                continue
            f.write(f"| `{k}` | {cnt[k]} | {v/1e3:.2f} | {100*v/T:.1f}% | {v/cnt[k]:.1f} |\n")
        f.write(f"\ntotal {T/1e3:.1f} ms over {sum(cnt.values())} launches\n")


def ncu_raw(rep, name, pattern):
    p = os.path.join(G, rep)
    if not os.path.exists(p):
        return
    raw = subprocess.run(["ncu", "-i", p, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    if len(rows) < 3:
        return
    hdr, units = rows[0], rows[1]
    want = ["Kernel Name", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
            "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
            "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
            "launch__registers_per_thread", "launch__grid_size", "launch__block_size"]
    idx = [hdr.index(w) for w in want if w in hdr]
    with open(os.path.join(out_dir, f"{tag}_{name}.csv"), "w") as f:
        w = csv.writer(f)
        w.writerow([hdr[i] for i in idx])
        w.writerow([units[i] for i in idx])
        for row in rows[2:]:
            out = [row[i] for i in idx]
            out[0] = re.sub(r"void |b2c::|\(.*", "", out[0])[:70]
            w.writerow(out)


def ncu_kernels():
    """gpurun_out/prof_kernels_raw.csv (text export of `ncu --set full` over tools/prof_kernels.py, written on the GPU box
    by tools/gpu_prof3.sh — the .ncu-rep itself exceeds the return limit) → profiles/<tag>_ncu_kernels.md"""
    p = os.path.join(G, "prof_kernels_raw.csv")
    if not os.path.exists(p):
        return
    rows = list(csv.reader(open(p)))
    hdr = rows[0]
    col = {h: i for i, h in enumerate(hdr)}
    stall = [h for h in hdr if "smsp__average_warps_issue_stalled" in h and h.endswith("_per_issue_active.ratio")]
    sel = [("gpu__time_duration.sum", "time us"), ("sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_active", "tensor pipe %"),
           ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue %"), ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "L2 %"),
           ("l1tex__data_pipe_lsu_wavefronts.sum.pct_of_peak_sustained_elapsed", "LSU %"),
           ("dram__bytes_read.sum", "DRAM rd MB"), ("dram__bytes_write.sum", "DRAM wr MB"), ("launch__registers_per_thread", "regs")]
    with open(os.path.join(out_dir, f"{tag}_ncu_kernels.md"), "w") as f:
        f.write(f"# {tag}: `ncu --set full --clock-control none` over `tools/prof_kernels.py` (hot kernels alone at the bench shapes: "
                "M = 24000 tokens, T = 750, 8 heads x 64, dropout 0.1)\n\n")
        f.write("Times under ncu are serialised and cache-cold; `tools/prof_kernels.py --time` (CUDA events) gives the timings quoted "
                "in DESIGN.md.  DRAM MB = `dram__bytes_read/write.sum` per launch.\n\n")
        f.write("| kernel | " + " | ".join(n for _, n in sel) + " | top stalls (warp-cycles per issued instruction) |\n")
        f.write("|---|" + "---:|" * len(sel) + "---|\n")
        for r in rows[2:]:
            name = re.sub(r"void |b2c::|\(.*", "", r[col["Kernel Name"]])[:60]
            vals = []
            for k, _ in sel:
                v = r[col[k]] if k in col else ""
                try:
                    vals.append(f"{float(v):.1f}")
                except ValueError:
                    vals.append(v)
            st = sorted(((float(r[col[h]] or 0), h.replace("smsp__average_warps_issue_stalled_", "").replace("_per_issue_active.ratio", ""))
                         for h in stall), reverse=True)[:4]
            f.write(f"| `{name}` | " + " | ".join(vals) + " | " + ", ".join(f"{b} {a:.2f}" for a, b in st) + " |\n")
    kt = os.path.join(G, "kernel_times.log")
    if os.path.exists(kt):
        with open(os.path.join(out_dir, f"{tag}_kernel_times.txt"), "w") as f:
            f.write("# tools/prof_kernels.py --time on one B200 (CUDA events, 20 iterations, 256 MB L2 flush between iterations)\n")
            f.write(open(kt).read())


def dram_traffic():
    """profiles/dram_traffic.json: DRAM bytes per launch (dram__bytes_read.sum + dram__bytes_write.sum) from the `ncu --set full`
    capture of tools/prof_kernels.py — what bench.py puts into `roofline.traffic`.  `gemm_family_cfg2` = mean over the eight
    representative GEMM launches of a Conformer-L layer (FFN fwd / dgrad / wgrad, qkv); `mel_frontend_32x30s` = mel_frames +
    mel_normalize of one 32 x 30 s pass."""
    p = os.path.join(G, "prof_kernels_raw.csv")
    if not os.path.exists(p):
        return
    rows = list(csv.reader(open(p)))
    hdr = rows[0]
    col = {h: i for i, h in enumerate(hdr)}
    unit = {h: rows[1][i] for h, i in col.items()}

    def to_bytes(r, key):
        v = float((r[col[key]] or "0").replace(",", ""))
        u = unit.get(key, "byte").lower()
        return v * {"byte": 1, "kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9}.get(u, 1)

    per = []
    for r in rows[2:]:
        name = re.sub(r"void |b2c::|\(.*", "", r[col["Kernel Name"]])[:70]
        per.append((name, to_bytes(r, "dram__bytes_read.sum") + to_bytes(r, "dram__bytes_write.sum")))
    gem = [b for n, b in per if n.startswith("gemm_tc_kernel")]
    mel = [b for n, b in per if n.startswith("mel_")]
    out = {"source": f"ncu --set full --clock-control none over tools/prof_kernels.py ({tag}); bytes per launch",
           "gemm_family_cfg2": sum(gem) / len(gem) if gem else None,
           "gemm_family_cfg2_detail": {f"{i}:{n}": b for i, (n, b) in enumerate(per) if n.startswith("gemm_tc_kernel")},
           "mel_frontend_32x30s": sum(mel[-2:]) if len(mel) >= 2 else None,
           "all": {f"{i}:{n}": b for i, (n, b) in enumerate(per)}}
    json.dump(out, open(os.path.join(out_dir, "dram_traffic.json"), "w"), indent=1)


launches()
ncu_kernels()
dram_traffic()
for f in ("bench.log",):
    p = os.path.join(G, f)
    if os.path.exists(p):
        for line in open(p):
            if line.startswith("{"):
                json.dump(json.loads(line), open(os.path.join(out_dir, f"{tag}_bench.json"), "w"), indent=1)
print(sorted(os.listdir(out_dir)))
