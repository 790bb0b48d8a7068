"""Warp-instructions executed per CUDA source line / per function of a kernel in an .ncu-rep (dev tool).
usage: python tools/ncu_inst.py <rep> <kernel-regex> [top]"""
import csv, io, re, subprocess, sys
rep, kre = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 25
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass",
                      "--kernel-name", "regex:" + kre], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
fpath, hdr, per_line = None, None, {}
for r in rows:
    if not r:
        continue
    if r[0] == "File Path":
        fpath = r[1]; continue
    if r[0] == "Function Name":
        continue
    if r[0] == "Line No":
        hdr = r; continue
    if hdr is None or r[0] == "":
        continue
    d = dict(zip(hdr[4:], r[4:]))
    try:
        n = int(d.get("Instructions Executed", "0") or 0)
    except ValueError:
        continue
    k = (fpath, int(r[0]), r[1].strip()[:90])
    per_line[k] = per_line.get(k, 0) + n
tot = sum(per_line.values())
print(f"total warp-instructions {tot / 1e6:.1f} M")
# function of a line: nearest preceding '__device__ ... name(' / '__global__' definition in the file
funcs = {}
for f in {k[0] for k in per_line}:
    try:
        src = open(f).read().splitlines()
    except OSError:
        continue
    cur, table = "?", []
    for i, ln in enumerate(src, 1):
        m = re.search(r"(?:__device__|__global__)[^;]*?\b([A-Za-z_][A-Za-z0-9_]*)\s*\(", ln)
        if m and not ln.strip().startswith("//"):
            cur = m.group(1)
        table.append(cur)
    funcs[f] = table
agg = {}
for (f, l, _), n in per_line.items():
    fn = funcs.get(f, ["?"] * (l + 1))[l - 1] if l - 1 < len(funcs.get(f, [])) else "?"
    agg[(f.split("/")[-1], fn)] = agg.get((f.split("/")[-1], fn), 0) + n
for (f, fn), n in sorted(agg.items(), key=lambda x: -x[1])[:top]:
    print(f"{100 * n / tot:5.1f}%  {n / 1e6:8.1f} M  {f}:{fn}")
