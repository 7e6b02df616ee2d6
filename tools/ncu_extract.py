"""Condense `ncu -i X.ncu-rep --page raw --csv` into one line per captured launch: duration, DRAM bytes read + written,
DRAM / tensor-pipe / issue utilisation, registers, achieved occupancy.  usage: python tools/ncu_extract.py raw.csv"""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
hdr = rows[0]
col = {n: i for i, n in enumerate(hdr)}


def find(*names):
    for n in names:
        for h in hdr:
            if h == n or h.startswith(n):
                return col[h]
    return None


want = {
    "dur_us": find("gpu__time_duration.sum"),
    "dram_rd": find("dram__bytes_read.sum"),
    "dram_wr": find("dram__bytes_write.sum"),
    "dram_pct": find("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"),
    "tensor_pct": find("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_tensor_op"),
    "issue_pct": find("sm__inst_issued.avg.pct_of_peak_sustained_active", "sm__issue_active.avg.pct"),
    "warps_pct": find("sm__warps_active.avg.pct_of_peak_sustained_active"),
    "regs": find("launch__registers_per_thread"),
    "l2_hit": find("lts__t_sector_hit_rate.pct"),
}
units = rows[1]
ki, gi = col["Kernel Name"], col.get("Grid Size")


def num(r, i):
    if i is None:
        return float("nan")
    try:
        return float(r[i].replace(",", ""))
    except ValueError:
        return float("nan")


def scale_bytes(v, unit):
    return v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(unit, 1)


def scale_time(v, unit):
    return v * {"ns": 1e-3, "us": 1, "ms": 1e3, "s": 1e6, "usecond": 1, "nsecond": 1e-3, "msecond": 1e3, "second": 1e6}.get(unit, 1)


print(f"{'kernel':58s} {'grid':>14s} {'us':>8s} {'DRAM rd MB':>10s} {'wr MB':>8s} {'GB/s':>7s} {'dram%':>6s} {'tens%':>6s} {'issue%':>6s} {'warps%':>6s} {'regs':>5s} {'L2hit%':>6s}")
for r in rows[2:]:
    if len(r) < len(hdr):
        continue
    d = scale_time(num(r, want["dur_us"]), units[want["dur_us"]]) if want["dur_us"] is not None else float("nan")
    rd = scale_bytes(num(r, want["dram_rd"]), units[want["dram_rd"]]) if want["dram_rd"] is not None else float("nan")
    wr = scale_bytes(num(r, want["dram_wr"]), units[want["dram_wr"]]) if want["dram_wr"] is not None else float("nan")
    name = r[ki].split("(")[0][-58:]
    print(f"{name:58s} {r[gi] if gi is not None else '':>14s} {d:8.1f} {rd / 1e6:10.1f} {wr / 1e6:8.1f} {(rd + wr) / d / 1e3 if d else 0:7.0f} "
          f"{num(r, want['dram_pct']):6.1f} {num(r, want['tensor_pct']):6.1f} {num(r, want['issue_pct']):6.1f} {num(r, want['warps_pct']):6.1f} "
          f"{num(r, want['regs']):5.0f} {num(r, want['l2_hit']):6.1f}")
