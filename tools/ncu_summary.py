"""One table row per profiled launch out of `ncu -i X.ncu-rep --page raw --csv` exports (dev tool).
usage: python tools/ncu_summary.py gpurun_out/a_raw.csv [b_raw.csv ...]  > profiles/kernels_rNN.md
Columns: device time, DRAM bytes (read + write), achieved DRAM GB/s, FMA-pipe and issue-slot utilisation,
resident warps, registers, grid x block, dynamic shared memory, and the three largest warp-stall reasons
(cycles stalled per issued instruction)."""
import csv
import sys

KEYS = {
    "time": "gpu__time_duration.sum",
    "rd": "dram__bytes_read.sum",
    "wr": "dram__bytes_write.sum",
    "fma": "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
    "issue": "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "warps": "sm__warps_active.avg.pct_of_peak_sustained_active",
    "regs": "launch__registers_per_thread",
    "grid": "launch__grid_size",
    "block": "launch__block_size",
    "smem": "launch__shared_mem_per_block_dynamic",
}
UNIT = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12,
        "ns": 1e-9, "us": 1e-6, "ms": 1e-3, "s": 1.0, "usecond": 1e-6, "msecond": 1e-3, "nsecond": 1e-9, "second": 1.0}


def val(row, units, col, name):
    i = col.get(name)
    if i is None or row[i] in ("", "n/a"):
        return None
    return float(row[i].replace(",", "")) * UNIT.get(units[i].split("/")[0], 1.0)


def main():
    print("| kernel | time | DRAM rd+wr | DRAM GB/s | fma pipe % | issue % | warps % | regs | grid x block | dyn smem | top stalls (cycles / issue) |")
    print("|---|---:|---:|---:|---:|---:|---:|---:|---|---:|---|")
    for path in sys.argv[1:]:
        rows = list(csv.reader(open(path)))
        hdr, units = rows[0], rows[1]
        col = {n: i for i, n in enumerate(hdr)}
        kn = col["Kernel Name"]
        stall = [(n, i) for n, i in col.items()
                 if n.startswith("smsp__average_warps_issue_stalled_") and n.endswith("_per_issue_active.ratio")
                 and "not_issued" not in n and "selected" not in n.replace("not_selected", "")]
        for r in rows[2:]:
            if len(r) <= kn:
                continue
            g = {k: val(r, units, col, v) for k, v in KEYS.items()}
            st = sorted(((float(r[i]), n.split("stalled_")[1].split("_per_issue")[0]) for n, i in stall if r[i] not in ("", "n/a")),
                        reverse=True)[:3]
            t = g["time"] or 0.0
            dram = (g["rd"] or 0.0) + (g["wr"] or 0.0)
            name = r[kn].split("(")[0]
            print(f"| `{name}` ({path.split('/')[-1].replace('_raw.csv', '')}) | {t * 1e3:.3f} ms | {dram / 1e6:.2f} MB | "
                  f"{dram / t / 1e9 if t else 0:.0f} | {g['fma'] or 0:.1f} | {g['issue'] or 0:.1f} | "
                  f"{g['warps'] or 0:.1f} | {int(g['regs'] or 0)} | {int(g['grid'] or 0)} x {int(g['block'] or 0)} | "
                  f"{(g['smem'] or 0) / 1e3:.1f} KB | " + ", ".join(f"{n} {v:.2f}" for v, n in st) + " |")


if __name__ == "__main__":
    main()
