#!/usr/bin/env python
"""Record the DRAM traffic of one hot-kernel launch from an `ncu --set full` report into profiles/r02_traffic.json
(read by bench.py as `roofline.traffic`, only when the kernel label of the bench run matches the recorded one).

    python tools/ncu_traffic.py REPORT.ncu-rep BENCH_LINE.json [--flow dense]

BENCH_LINE.json is the JSON line of a `bench.py` run made with the same build and knobs (it carries the kernel label
and the workload); the report must hold one launch of that kernel at the full workload size.
"""
import csv
import io
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def main():
    rep, line_path = sys.argv[1], sys.argv[2]
    line = json.loads(open(line_path).read().strip().splitlines()[-1])
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    h, units = rows[0], rows[1]
    best = None
    for r in rows[2:]:
        rd, wr = float(r[h.index("dram__bytes_read.sum")]), float(r[h.index("dram__bytes_write.sum")])
        scale = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0, "Tbyte": 1e12}
        tot = rd * scale[units[h.index("dram__bytes_read.sum")]] + wr * scale[units[h.index("dram__bytes_write.sum")]]
        if best is None or tot > best[0]:
            best = (tot, r[h.index("Kernel Name")], float(r[h.index("gpu__time_duration.sum")]), units[h.index("gpu__time_duration.sum")])
    key = f"{line['config']['workload']}|{line['config'].get('flow', 'dense')}"
    out_path = os.path.join(ROOT, "profiles", "r02_traffic.json")
    data = json.load(open(out_path)) if os.path.exists(out_path) else {}
    data[key] = {"kernel": line["roofline"]["kernel"], "ncu_kernel_name": best[1], "dram_bytes": int(best[0]),
                 "algorithmic_bytes": line["roofline"]["algorithmic_bytes_per_launch"],
                 "ratio": round(best[0] / line["roofline"]["algorithmic_bytes_per_launch"], 4),
                 "duration_under_ncu": f"{best[2]} {best[3]}",
                 "source": f"ncu --set full --clock-control none, one launch ({os.path.basename(rep)}); dram__bytes_read.sum + dram__bytes_write.sum"}
    json.dump(data, open(out_path, "w"), indent=1, sort_keys=True)
    print(key, data[key])


if __name__ == "__main__":
    main()
