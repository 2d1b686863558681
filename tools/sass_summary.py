"""profiles/<tag>_sass_summary.md: per-kernel counts of the SASS mnemonics that prove tcgen05 / TMEM / TMA use
(B200_PROFILING.md: UTCHMMA = tcgen05.mma, LDTM / STTM = tcgen05.ld / st, UTMALDG / UTMASTG = TMA load / store, UTCBAR = tcgen05.commit,
HMMA = mma.sync) from `cuobjdump -sass` of the in-tree library.   python tools/sass_summary.py r02"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tag = sys.argv[1] if len(sys.argv) > 1 else "r02"
lib = os.path.join(ROOT, "nemo_b200", "csrc", "libb200conformer.so")
out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
names = subprocess.run(["c++filt"], input="\n".join(re.findall(r"Function : (\S+)", out)), capture_output=True, text=True).stdout.splitlines()
MN = ["UTCHMMA", "UTCQMMA", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UTCBAR", "HMMA", "SYNCS", "MUFU", "FFMA"]
rows, cur, i = [], None, -1
for line in out.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        i += 1
        cur = collections.Counter()
        rows.append((names[i] if i < len(names) else m.group(1), cur))
        continue
    if cur is None:
        continue
    m = re.search(r"^\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
    if m:
        op = m.group(1).split(".")[0]
        cur["total"] += 1
        for k in MN:
            if op == k:
                cur[k] += 1
with open(os.path.join(ROOT, "profiles", f"{tag}_sass_summary.md"), "w") as f:
    f.write(f"# {tag}: SASS instruction counts per kernel (`cuobjdump -sass nemo_b200/csrc/libb200conformer.so`, sm_100a)\n\n")
    f.write("UTCHMMA = `tcgen05.mma` (kind::f16), LDTM/STTM = `tcgen05.ld/st` (TMEM), UTMALDG/UTMASTG = TMA tensor load/store, UTCBAR = "
            "`tcgen05.commit`, HMMA = `mma.sync` (head-dim-128 attention and the position-side backward only).\n\n")
    f.write("| kernel | SASS instrs | " + " | ".join(MN) + " |\n|---|---:|" + "---:|" * len(MN) + "\n")
    tot = collections.Counter()
    for name, c in sorted(rows, key=lambda r: -sum(r[1][k] for k in MN[:7])):
        short = re.sub(r"\(.*", "", name.replace("b2c::", "").replace("void ", ""))[:78]
        f.write(f"| `{short}` | {c['total']} | " + " | ".join(str(c[k]) for k in MN) + " |\n")
        tot.update(c)
    f.write(f"| **all {len(rows)} kernels** | {tot['total']} | " + " | ".join(str(tot[k]) for k in MN) + " |\n")
print(open(os.path.join(ROOT, "profiles", f"{tag}_sass_summary.md")).read()[:3000])
