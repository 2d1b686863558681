"""Summarise an ncu launch list (`ncu --metrics gpu__time_duration.sum --clock-control none --csv`) into a per-kernel table.
  python tools/launch_table.py gpurun_out/launches_cfg3.csv "title" > profiles/r02_launches_cfg3.md"""
import collections
import csv
import re
import sys

path, title = sys.argv[1], (sys.argv[2] if len(sys.argv) > 2 else "")
lines = [l for l in open(path) if l.startswith('"')]
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
    key = f"gemm_tc_kernel<{m.group(1)}>" if m else re.sub(r"\(.*", "", re.sub(r"<.*", "", name)).replace("void ", "")[:70]
    tot[key] += v
    cnt[key] += 1
T = sum(tot.values())
print(f"# {title}\n")
print("`ncu --metrics gpu__time_duration.sum --clock-control none` — per-launch times are cold-cache and serialised: compare SHARES.  "
      "gemm_tc_kernel<BN, A_MN, B_MN, EPILOGUE, CONV>.\n")
print("| kernel | launches | total ms | share | avg us |\n|---|---:|---:|---:|---:|")
for k, v in sorted(tot.items(), key=lambda x: -x[1]):
    if v / T < 0.002:
        continue
    print(f"| `{k}` | {cnt[k]} | {v/1e3:.2f} | {100*v/T:.1f}% | {v/cnt[k]:.1f} |")
print(f"\ntotal {T/1e3:.1f} ms over {sum(cnt.values())} launches")
