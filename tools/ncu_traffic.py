#!/usr/bin/env python3
"""profiles/ncu_traffic.json from `ncu --set full` captures: DRAM bytes read + written per launch of the dominant kernel of
each bench stage, for the workload shape the capture was taken at.  bench.py reports `roofline.traffic` from this file when
the run has the same number of frames (never a constant in the source).

    python tools/ncu_traffic.py <workload> <frames> <report.ncu-rep> [<report2.ncu-rep> ...]
"""
import csv
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
STAGE_OF = (("mfcc_frames_kernel", "mfcc"), ("gmm_tc16_kernel", "gmm"), ("gmm_tc_kernel", "gmm"), ("gmm_loglik_mma", "gmm"),
            ("gmm_loglik_exact_kernel", "gmm"), ("trellis_band2_kernel", "viterbi"), ("trellis_band_kernel", "viterbi"),
            ("trellis_small_kernel", "viterbi"), ("trellis_reg_kernel", "viterbi"))
UNIT = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}


def main():
    workload, frames, reps = sys.argv[1], int(sys.argv[2]), sys.argv[3:]
    path = os.path.join(ROOT, "profiles", "ncu_traffic.json")
    tab = json.load(open(path)) if os.path.exists(path) else {}
    ent = {"frames": frames, "source": ", ".join("profiles/" + os.path.basename(r).replace(".ncu-rep", "_ncu_full.txt") for r in reps), "stages": {}}
    for rep in reps:
        out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
        rows = list(csv.reader(out.splitlines()))
        hdr, units = rows[0], rows[1]
        col = {h: i for i, h in enumerate(hdr)}
        for r in rows[2:]:
            name = r[col["Kernel Name"]]
            stage = next((s for k, s in STAGE_OF if k in name), None)
            if stage is None or stage in ent["stages"]:
                continue
            rd = float(r[col["dram__bytes_read.sum"]]) * UNIT[units[col["dram__bytes_read.sum"]]]
            wr = float(r[col["dram__bytes_write.sum"]]) * UNIT[units[col["dram__bytes_write.sum"]]]
            ent["stages"][stage] = {"kernel": name.split("(")[0][:80], "dram_read_bytes": rd, "dram_write_bytes": wr,
                                    "duration_us_under_ncu": float(r[col["gpu__time_duration.sum"]]) *
                                    {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}.get(units[col["gpu__time_duration.sum"]], 1.0)}
    tab[workload] = ent
    json.dump(tab, open(path, "w"), indent=1)
    print(json.dumps(ent, indent=1))


if __name__ == "__main__":
    main()
