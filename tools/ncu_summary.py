#!/usr/bin/env python
"""ncu_summary.py report.ncu-rep [kernel-substring] - the handful of numbers DESIGN.md / profiles/*.md quote, per captured launch,
as a markdown table: duration, DRAM bytes read / written (and per unit if --units N is given), DRAM and SM throughput as % of peak,
achieved occupancy, active threads per instruction, L2 hit rate, registers.  Reads the report with `ncu -i ... --page raw --csv`
(ncu is in the image; no GPU needed).  Usage after a capture:  python tools/ncu_summary.py gpurun_out/x.ncu-rep k1_fused --units 239500000"""
import csv
import io
import subprocess
import sys

WANT = [
    ("gpu__time_duration.sum", "duration"),
    ("dram__bytes_read.sum", "DRAM read"),
    ("dram__bytes_write.sum", "DRAM written"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "DRAM throughput % of peak"),
    ("dram__throughput.avg.pct_of_peak_sustained_elapsed", "DRAM throughput % of peak"),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "SM throughput %"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "achieved occupancy %"),
    ("smsp__thread_inst_executed_per_inst_executed.ratio", "active threads / instruction"),
    ("lts__t_sector_hit_rate.pct", "L2 sector hit rate %"),
    ("lts__t_sectors_srcunit_tex_op_write.sum", "L2 write sectors from SMs"),
    ("lts__t_sectors_srcunit_tex_op_atom.sum", "L2 atomic sectors"),
    ("launch__registers_per_thread", "registers / thread"),
    ("launch__occupancy_limit_registers", "CTAs/SM by registers"),
    ("launch__occupancy_limit_shared_mem", "CTAs/SM by shared memory"),
]
SCALE = {"Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12, "byte": 1.0, "ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3,
         "usecond": 1e-3, "msecond": 1.0, "nsecond": 1e-6, "second": 1e3}


def main():
    args = [a for a in sys.argv[1:] if not a.startswith("--")]
    units = None
    if "--units" in sys.argv:
        units = float(sys.argv[sys.argv.index("--units") + 1])
        args = [a for a in args if a != sys.argv[sys.argv.index("--units") + 1]]
    if not args:
        print(__doc__)
        return 1
    rep, pat = args[0], (args[1] if len(args) > 1 else "")
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True)
    if out.returncode != 0:
        print(out.stderr, file=sys.stderr)
        return 1
    rows = list(csv.reader(io.StringIO(out.stdout)))
    head, unit, data = rows[0], rows[1], rows[2:]
    col = {}
    for i, h in enumerate(head):
        for name, _ in WANT:
            if h == name or h.endswith("." + name):
                col.setdefault(name, i)
    kcol = head.index("Kernel Name")
    for r in data:
        if pat and pat not in r[kcol]:
            continue
        print(f"### `{r[kcol].split('(')[0]}`  (launch id {r[0]}, grid {r[head.index('Grid Size')]}, block {r[head.index('Block Size')]})\n")
        print("| metric | value |\n|---|---|")
        vals = {}
        for name, label in WANT:
            if name not in col or r[col[name]] in ("", "no data"):
                continue
            v, u = float(r[col[name]].replace(",", "")), unit[col[name]]
            if u in SCALE and ("byte" in u or "second" in u or u in ("ns", "us", "ms", "s")):
                v *= SCALE[u]
                u = "B" if "byte" in u else "ms"
            vals[name] = v
            txt = f"{v / 1e9:.3f} GB" if u == "B" else f"{v:.3f} ms" if u == "ms" else f"{v:.4g} {u}".rstrip()
            if units and u == "B":
                txt += f" ({v / units:.1f} B per unit)"
            print(f"| {name} ({label}) | {txt} |")
        if units and "gpu__time_duration.sum" in vals:
            print(f"| units / s | {units / (vals['gpu__time_duration.sum'] / 1e3) / 1e9:.2f} G/s |")
        print()
    return 0


if __name__ == "__main__":
    raise SystemExit(main())
