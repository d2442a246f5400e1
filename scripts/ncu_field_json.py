#!/usr/bin/env python
"""ncu report -> profiles/r02_field_ncu.json (the tracked file bench.py reads `roofline.traffic` and the ncu counters from).

    ncu -i gpurun_out/prof_fp16.ncu-rep --page raw --csv > /tmp/raw.csv        (or pass the .ncu-rep: this script calls ncu itself)
    python scripts/ncu_field_json.py gpurun_out/prof_fp16.ncu-rep profiles/r02_field_ncu.json

The capture must hold one full round of the field pair (k_tc_amb + k_tc_sigcol over 8,388,608 samples) of
`python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-ref-cuda --no-may --eager` (scripts/gpu_profile.sh).  Per kernel: duration,
DRAM bytes read / written, L2 bytes, tensor-pipe and issue utilisation, executed warp instructions, registers; `dram_bytes_per_round` =
sum over both kernels of dram read + write = roofline.traffic (per launch pair, like roofline.algorithmic_bytes_per_launch)."""
import csv
import io
import json
import subprocess
import sys

WANT = {
    "gpu__time_duration.sum": "duration",
    "dram__bytes_read.sum": "dram_read",
    "dram__bytes_write.sum": "dram_write",
    "lts__t_bytes.sum": "l2_bytes",
    "lts__t_sector_hit_rate.pct": "l2_hit_pct",
    "l1tex__t_sector_hit_rate.pct": "l1_hit_pct",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active": "tensor_pipe_pct",
    "smsp__issue_active.avg.pct_of_peak_sustained_active": "issue_active_pct",
    "smsp__inst_executed.sum": "warp_inst",
    "launch__registers_per_thread": "registers",
    "launch__grid_size": "grid",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed": "sm_throughput_pct",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed": "dram_throughput_pct",
}
SCALE = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0, "ms": 1e-3, "us": 1e-6, "ns": 1e-9, "s": 1.0, "msecond": 1e-3, "usecond": 1e-6, "nsecond": 1e-9, "second": 1.0}


def main(rep, out):
    txt = open(rep).read() if rep.endswith(".csv") else subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(io.StringIO(txt)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    col = {h: i for i, h in enumerate(hdr)}
    kernels = []
    for r in data:
        k = {"kernel": r[col["Kernel Name"]]}
        for metric, name in WANT.items():
            if metric in col and r[col[metric]] not in ("", "n/a"):
                v = float(r[col[metric]].replace(",", ""))
                k[name] = v * SCALE.get(units[col[metric]], 1.0)
        kernels.append(k)
    # one FULL round per field kernel: the capture window may also contain the (empty) extra-round launches -> keep the longest launch of each
    field = []
    for name in ("k_tc_amb", "k_tc_sigcol"):
        cand = [k for k in kernels if name in k["kernel"]]
        if cand:
            field.append(max(cand, key=lambda k: k.get("duration", 0.0)))
    kernels = field or kernels
    res = {"source": rep, "kernels": kernels,
           "dram_bytes_per_round": int(sum(k.get("dram_read", 0) + k.get("dram_write", 0) for k in field)) if field else None,
           "l2_bytes_per_round": int(sum(k.get("l2_bytes", 0) for k in field)) if field else None,
           "note": "one ncu --set full --clock-control none capture per kernel (cold cache, serialised): durations are NOT bench values"}
    json.dump(res, open(out, "w"), indent=1)
    print(json.dumps(res)[:600])


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2])
