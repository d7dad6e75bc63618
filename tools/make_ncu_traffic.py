#!/usr/bin/env python
"""make_ncu_traffic.py out.json kernel=report.ncu-rep:units[:note] ... - DRAM bytes per unit of work from `ncu --set full` captures, for
bench.py's roofline.traffic (profiles/r02_ncu_traffic.json).  units = the events / pending events the captured launch processed."""
import csv
import io
import json
import subprocess
import sys


def metrics(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    head, unit, data = rows[0], rows[1], rows[2]
    scale = {"Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12, "byte": 1.0, "ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3,
             "usecond": 1e-3, "msecond": 1.0, "nsecond": 1e-6, "second": 1e3}
    got = {}
    for name in ("dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__time_duration.sum"):
        i = head.index(name)
        got[name] = float(data[i].replace(",", "")) * scale.get(unit[i], 1.0)
    got["kernel"] = data[head.index("Kernel Name")]
    return got


def main():
    out = {}
    for spec in sys.argv[2:]:
        kern, rest = spec.split("=", 1)
        parts = rest.split(":")
        rep, units = parts[0], float(parts[1])
        m = metrics(rep)
        tot = m["dram__bytes_read.sum"] + m["dram__bytes_write.sum"]
        out[kern] = {"dram_bytes_per_unit": tot / units, "dram_bytes_read": m["dram__bytes_read.sum"], "dram_bytes_written": m["dram__bytes_write.sum"],
                     "units_in_captured_launch": units, "duration_ms_under_ncu": m["gpu__time_duration.sum"], "source": "profiles/" + rep.split("/")[-1].replace(".ncu-rep", ".md"),
                     "note": parts[2] if len(parts) > 2 else ""}
    json.dump(out, open(sys.argv[1], "w"), indent=1)
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
