"""Summarise an ncu launch list (csv, gpu__time_duration.sum) of scripts/vae_profile.py: the LAST encode and decode
passes, time per kernel family.  A pass starts at its conv_thin_to_wide (conv_in) launch and ends at conv_wide_to_thin."""
import csv
import re
import sys
from collections import OrderedDict

rows = []
with open(sys.argv[1], newline="") as f:
    lines = [l for l in f if l.startswith('"')]
for r in csv.DictReader(lines):
    if r.get("Metric Name") == "gpu__time_duration.sum":
        v = float(r["Metric Value"].replace(",", ""))
        unit = r.get("Metric Unit", "ns")
        v *= {"ns": 1e-3, "us": 1.0, "ms": 1e3}.get(unit, 1e-3)
        rows.append((r["Kernel Name"], v))
starts = [i for i, (n, _) in enumerate(rows) if "conv_thin_to_wide" in n]
ends = [i for i, (n, _) in enumerate(rows) if "conv_wide_to_thin" in n]
passes = list(zip(starts, ends))[-2:]
for label, (a, b) in zip(("encode", "decode"), passes):
    fam = OrderedDict()
    for n, v in rows[a:b + 1]:
        key = n.replace("void ", "").replace("<unnamed>::", "").replace("(anonymous namespace)::", "")
        key = re.sub(r"\(.*", "", key).strip()
        if "gemm" in key:
            m = re.search(r"gemm_kernel<(\d+), *(\d+), *\(?\w*\)?(\d+|true|false)", n)
            key = f"vn_gemm BN{m.group(1)} split{m.group(3)}" if m else key
        c = fam.setdefault(key, [0, 0.0])
        c[0] += 1; c[1] += v
    tot = sum(v for _, v in rows[a:b + 1])
    print(f"{label}: {b - a + 1} launches, {tot:.1f} us serialized (cold-cache, ncu)")
    for k, (c, v) in sorted(fam.items(), key=lambda kv: -kv[1][1]):
        print(f"   {k:40s} {c:4d} launches {v:9.1f} us  {100 * v / tot:5.1f} %")
    big = sorted(rows[a:b + 1], key=lambda r: -r[1])[:6]
    print("   longest:", ", ".join(f"{v:.0f}" for _, v in big), "us")
