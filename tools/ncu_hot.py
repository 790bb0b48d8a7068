"""Rank CUDA source lines of a kernel by warp-stall samples from an .ncu-rep (dev tool).
usage: python tools/ncu_hot.py <rep> <kernel-regex> [top]"""
import csv, subprocess, sys, io
rep, kre = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass",
                      "--kernel-name", "regex:" + kre], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
fpath, hdr, agg = None, None, {}
for r in rows:
    if not r:
        continue
    if r[0] == "File Path":
        fpath = r[1].split("/")[-1]; continue
    if r[0] == "Function Name":
        continue
    if r[0] == "Line No":
        hdr = r; continue
    if hdr is None or r[0] == "":
        continue  # SASS rows have empty line no
    d = dict(zip(hdr[4:], r[4:]))
    try:
        n = int(d["# Samples"])
    except Exception:
        continue
    st = {k: int(v) for k, v in d.items() if k.startswith("stall_") and "Not Issued" not in k and v not in ("", "0")}
    key = (fpath, r[0], r[1].strip()[:100])
    a = agg.setdefault(key, [0, 0, {}])
    a[0] += n
    a[1] += int(d.get("Instructions Executed", "0") or 0)
    for k, v in st.items():
        a[2][k] = a[2].get(k, 0) + v
tot = sum(a[0] for a in agg.values())
print("total samples", tot)
for (f, l, src), (n, ins, st) in sorted(agg.items(), key=lambda x: -x[1][0])[:top]:
    tops = sorted(st.items(), key=lambda x: -x[1])[:3]
    print(f"{100*n/tot:5.1f}% {f}:{l:>4} inst={ins:>10}  {src}\n         {tops}")
allst = {}
for (n, ins, st) in agg.values():
    for k, v in st.items():
        allst[k] = allst.get(k, 0) + v
ts = sum(allst.values())
print("stall reasons over the kernel:", ", ".join(f"{k[6:]} {100*v/ts:.1f}%" for k, v in sorted(allst.items(), key=lambda x: -x[1])[:10]))
