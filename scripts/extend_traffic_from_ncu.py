#!/usr/bin/env python
"""profiles/extend_traffic.json from an ncu `--set full` capture of the extend launches of ONE timed step of the default bench command.

Capture (one GPU, scripts/gpu_traffic.sh):
  ncu --set full --clock-control none -k regex:k_extend -s <launches before the timed step> -c <launches per step> -o gpurun_out/X \\
      python bench.py --steps 1 --warmup 3 --no-cpu-baseline
With the defaults (1080p, 64 spp = 2 passes of 32 spp, max path length 5) a step has 2 x 4 extend launches.
This script reads the .ncu-rep HERE (no GPU needed), sums dram__bytes_read.sum + dram__bytes_write.sum per launch and writes their
mean = `roofline.traffic` of bench.py (per launch, like `roofline.achieved`), with the per-launch list next to it.

  python scripts/extend_traffic_from_ncu.py gpurun_out/X.ncu-rep [--out profiles/extend_traffic.json]
"""
import csv
import io
import json
import subprocess
import sys

UNIT = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}


def main():
    rep = sys.argv[1]
    out = sys.argv[sys.argv.index("--out") + 1] if "--out" in sys.argv else "profiles/extend_traffic.json"
    txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(txt)))
    hdr, units = rows[0], rows[1]
    iname, ird, iwr, idur = hdr.index("Kernel Name"), hdr.index("dram__bytes_read.sum"), hdr.index("dram__bytes_write.sum"), hdr.index("gpu__time_duration.sum")
    launches = []
    for r in rows[2:]:
        if "k_extend" not in r[iname]:
            continue
        rd = float(r[ird].replace(",", "")) * UNIT[units[ird]]
        wr = float(r[iwr].replace(",", "")) * UNIT[units[iwr]]
        launches.append({"kernel": r[iname][:60], "dram_read_bytes": rd, "dram_write_bytes": wr, "duration": r[idur] + " " + units[idur]})
    if not launches:
        raise SystemExit("no k_extend launches in " + rep)
    mean = sum(l["dram_read_bytes"] + l["dram_write_bytes"] for l in launches) / len(launches)
    json.dump({"dram_bytes_per_launch": mean, "launches": launches, "source": rep,
               "command": "python bench.py --steps 1 --warmup 3 --no-cpu-baseline (the extend launches of the timed step)"}, open(out, "w"), indent=1)
    print(f"{len(launches)} launches, mean DRAM traffic {mean / 1e9:.3f} GB per launch -> {out}")


if __name__ == "__main__":
    main()
