"""Top source lines by warp-stall samples from an .ncu-rep captured with --import-source on (needs -lineinfo).
   python tools/ncu_lines.py rep.ncu-rep [kernel_index] [top_n]"""
import csv, subprocess, sys, io

rep = sys.argv[1]
kidx = int(sys.argv[2]) if len(sys.argv) > 2 else 0
top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"],
                     capture_output=True, text=True).stdout
kernels, cur, hdr, fpath = [], None, None, None
for row in csv.reader(io.StringIO(out)):
    if not row:
        continue
    if row[0] == "File Path":
        fpath = row[1]
        continue
    if row[0] == "Function Name":
        if not kernels or kernels[-1]["name"] != row[1] or kernels[-1].get("closed"):
            kernels.append({"name": row[1], "lines": {}, "sass": []})
        cur = kernels[-1]
        continue
    if row[0] == "Line No":
        hdr = row
        continue
    if cur is None or hdr is None:
        continue
    d = dict(zip(hdr, row))
    try:
        ns = int(d.get("# Samples", "0") or 0)
    except ValueError:
        continue
    if row[0]:      # source-level row
        key = (fpath.split("/")[-1], int(row[0]))
        e = cur["lines"].setdefault(key, [0, row[1], {}])
        e[0] += ns
        for k, v in d.items():
            if k.startswith("stall_") and "Not Issued" not in k and v not in ("", "0"):
                e[2][k] = e[2].get(k, 0) + int(v)
# kernels may repeat per file; merge by consecutive name
merged = []
for k in kernels:
    if merged and merged[-1]["name"] == k["name"] and not merged[-1].get("done"):
        for key, e in k["lines"].items():
            m = merged[-1]["lines"].setdefault(key, [0, e[1], {}])
            m[0] += e[0]
            for a, b in e[2].items():
                m[2][a] = m[2].get(a, 0) + b
    else:
        merged.append(k)
print(f"{len(merged)} kernel(s) in report")
for i, k in enumerate(merged):
    print(i, k["name"][:100])
k = merged[kidx]
tot = sum(e[0] for e in k["lines"].values())
print(f"== kernel {kidx}: {tot} samples")
for (f, ln), e in sorted(k["lines"].items(), key=lambda kv: -kv[1][0])[:top]:
    st = ", ".join(f"{a[6:]}={b}" for a, b in sorted(e[2].items(), key=lambda ab: -ab[1])[:3])
    print(f"{100 * e[0] / max(tot, 1):5.1f}% {f}:{ln:<4d} {e[1].strip()[:90]:90s} | {st}")
