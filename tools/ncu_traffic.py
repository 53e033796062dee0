"""ncu report -> per-launch DRAM traffic of the kernels matching a name: writes {"kernel", "launches", "dram_read_mb": [...],
"dram_write_mb": [...], "traffic_bytes": mean(read + write)} (bench.py's roofline.traffic reads it)."""
import csv
import json
import subprocess
import sys


def main(rep, match, out):
    txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(txt.splitlines()))
    hdr = rows[0]
    rd, wr, us, tc = [], [], [], []
    unit = dict(zip(hdr, rows[1]))
    scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
    for r in rows[2:]:
        d = dict(zip(hdr, r))
        if match not in d.get("Kernel Name", ""):
            continue
        rd.append(float(d["dram__bytes_read.sum"]) * scale.get(unit["dram__bytes_read.sum"], 1.0) / 1e6)
        wr.append(float(d["dram__bytes_write.sum"]) * scale.get(unit["dram__bytes_write.sum"], 1.0) / 1e6)
        us.append(float(d["gpu__time_duration.sum"]))
        tc.append(float(d.get("sm__pipe_tc_cycles_active.avg.pct_of_peak_sustained_elapsed", "nan")))
    res = {"kernel": match, "launches": len(rd), "dram_read_mb": rd, "dram_write_mb": wr, "us": us, "tensor_pipe_active_pct": tc,
           "traffic_bytes": (sum(rd) + sum(wr)) / max(len(rd), 1) * 1e6}
    json.dump(res, open(out, "w"))
    print(res)


if __name__ == "__main__":
    main(*sys.argv[1:4])
