"""SASS instructions with the most stall samples (with the source line each belongs to), in address order.
   python tools/ncu_sass.py rep.ncu-rep kernel_index [min_samples]"""
import csv, subprocess, sys, io
rep, kidx = sys.argv[1], int(sys.argv[2])
mins = int(sys.argv[3]) if len(sys.argv) > 3 else 100
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
k = -1
hdr = None
rows = []
for row in csv.reader(io.StringIO(out)):
    if not row:
        continue
    if row[0] == "Kernel Name":
        k += 1
        continue
    if row[0] in ("Address",):
        hdr = row
        continue
    if k == kidx and hdr and row[0].startswith("0x"):
        rows.append(dict(zip(hdr, row)))
tot = sum(int(r["# Samples"] or 0) for r in rows)
print("total samples", tot, "instructions", len(rows))
for i, r in enumerate(rows):
    n = int(r["# Samples"] or 0)
    if n >= mins:
        st = {a: int(b) for a, b in r.items() if a.startswith("stall_") and "Not Issued" not in a and b not in ("", "0")}
        s3 = ", ".join(f"{a[6:]}={b}" for a, b in sorted(st.items(), key=lambda ab: -ab[1])[:3])
        print(f"{i:5d} {100 * n / tot:5.1f}% {r['Source'][:70]:70s} | {s3}")
