"""Summarise ncu output brought back in gpurun_out/ into small tracked files under profiles/.

  python tools/ncu_summary.py launches gpurun_out/launches_r1.csv profiles/r1_launches.md [tail_launches]
  python tools/ncu_summary.py full     gpurun_out/prof_r1.ncu-rep  profiles/r1_ncu_full.md

`launches`: per-kernel count / total / share from a `--metrics gpu__time_duration.sum` launch list
(cold-cache, serialised: compare SHARES, not absolutes).  `full`: the handful of `--set full`
metrics the roofline discussion in DESIGN.md uses, one block per captured launch.
"""
import collections
import csv
import re
import subprocess
import sys

FULL_METRICS = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
    "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__instruction_throughput.avg.pct_of_peak_sustained_active", "sm__inst_executed.avg.per_cycle_elapsed",
    "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
    "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
    "launch__occupancy_limit_registers", "launch__grid_size", "launch__block_size", "smsp__inst_executed.sum",
]


def short(name):
    m = re.match(r"(?:void )?([\w:]+)(<[^(]*>)?", name)
    return (m.group(1) + (m.group(2) or "")) if m else name


def launches(src, dst, tail=None):
    rows = [r for r in csv.reader(open(src, errors="replace")) if len(r) > 10 and r[0].isdigit()]
    if tail:
        rows = rows[-int(tail):]
    tot, cnt = collections.Counter(), collections.Counter()
    for r in rows:
        k = short(r[4])
        tot[k] += float(r[-1])
        cnt[k] += 1
    T = sum(tot.values())
    with open(dst, "w") as f:
        f.write(f"# ncu launch list summary of `{src}` ({len(rows)} launches"
                f"{', last ' + str(tail) if tail else ''}; gpu__time_duration.sum, --clock-control none)\n\n")
        f.write("Per-launch times under ncu are cold-cache and serialised: the SHARE column is what is comparable\n"
                "with the live CUDA-event numbers in bench.py, not the absolute microseconds.\n\n")
        f.write("| kernel | launches | total us | share | avg us |\n|---|---:|---:|---:|---:|\n")
        for k, v in tot.most_common():
            f.write(f"| `{k}` | {cnt[k]} | {v / 1e3:.1f} | {v / T:.3f} | {v / cnt[k] / 1e3:.1f} |\n")
    print(open(dst).read())


def full(src, dst):
    raw = subprocess.run(["ncu", "-i", src, "--page", "raw", "--csv"], stdout=subprocess.PIPE, text=True,
                         stderr=subprocess.DEVNULL).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]
    ki = hdr.index("Kernel Name")
    with open(dst, "w") as f:
        f.write(f"# ncu --set full summary of `{src}` (--clock-control none --import-source on)\n\n")
        for r in rows[2:]:
            f.write(f"## `{short(r[ki])}`\n\n| metric | value | unit |\n|---|---:|---|\n")
            for m in FULL_METRICS:
                if m in hdr:
                    i = hdr.index(m)
                    f.write(f"| {m} | {r[i]} | {units[i]} |\n")
            f.write("\n")
    print(open(dst).read()[:3000])


if __name__ == "__main__":
    {"launches": launches, "full": full}[sys.argv[1]](*sys.argv[2:])
