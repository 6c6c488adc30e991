"""Reads dram__bytes_read.sum + dram__bytes_write.sum (and the duration) of one kernel out of an `ncu --page raw --csv`
export and writes profiles/zgemm_capture.json, the file bench.py's roofline.traffic comes from.

    ncu -i gpurun_out/x.ncu-rep --page raw --csv > raw.csv
    python tools/ncu_traffic.py raw.csv zgemm_dmma_persistent --flops 1.108e12 --bytes 2.15e9 --round 2 --launch "..."
"""
import argparse
import csv
import json
import os

UNIT = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("csv")
    ap.add_argument("kernel")
    ap.add_argument("--flops", type=float, required=True)
    ap.add_argument("--bytes", type=float, required=True)
    ap.add_argument("--round", type=int, default=2)
    ap.add_argument("--launch", default="")
    ap.add_argument("--out", default=os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "profiles",
                                                   "zgemm_capture.json"))
    args = ap.parse_args()
    rows = list(csv.reader(ln for ln in open(args.csv) if not ln.startswith("==")))
    head, units = rows[0], rows[1]
    col = {h: i for i, h in enumerate(head)}
    best = None
    for r in rows[2:]:
        if args.kernel not in r[col["Kernel Name"]]:
            continue
        def val(name):
            i = col[name]
            return float(r[i].replace(",", "")) * UNIT.get(units[i], 1.0)
        dram = val("dram__bytes_read.sum") + val("dram__bytes_write.sum")
        dur = float(r[col["gpu__time_duration.sum"]].replace(",", ""))
        dur_ms = dur * {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3, "nsecond": 1e-6, "usecond": 1e-3, "msecond": 1.0,
                        "second": 1e3}[units[col["gpu__time_duration.sum"]]]
        if best is None or dur_ms > best["duration_ms_under_ncu"]:
            best = {"kernel": r[col["Kernel Name"]][:80], "dram_bytes": dram, "duration_ms_under_ncu": dur_ms}
            for k in ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
                      "sm__inst_executed_pipe_tensor.sum", "lts__t_sector_hit_rate.pct"):
                if k in col:
                    best[k] = float(r[col[k]].replace(",", ""))
    if best is None:
        raise SystemExit("kernel not found in " + args.csv)
    best.update({"launch": args.launch, "flops": args.flops, "algorithmic_bytes": args.bytes, "round": args.round,
                 "source": os.path.basename(args.csv)})
    with open(args.out, "w") as f:
        json.dump(best, f, indent=1)
    print(json.dumps(best))


if __name__ == "__main__":
    main()
