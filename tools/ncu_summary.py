"""Turn `ncu --page raw --csv` exports (made on the GPU box, see profiles/r02_notes.md for the commands) into what gets
committed: a per-kernel table of the metrics the notes cite, and profiles/ncu_traffic.json -- the per-launch DRAM bytes
bench.py reports as `roofline.traffic`, stamped with the digest of the library that was profiled.

    python tools/ncu_summary.py gpurun_out/r02_ncu_full_tile_kernels.csv gpurun_out/r02_ncu_full_rowwise_kernels.csv \
        --out profiles/r02_ncu_summary.md
"""
import argparse
import csv
import json
import os
import sys

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
METRICS = [
    ("gpu__time_duration.sum", "duration"),
    ("dram__bytes_read.sum", "dram read"),
    ("dram__bytes_write.sum", "dram write"),
    ("lts__t_bytes.sum", "L2 bytes"),
    ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor pipe % of active"),
    ("TPC.TriageCompute.sm__pipe_tensor_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed", "tensor pipe % of elapsed"),
    ("sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "XU (MUFU) % of active"),
    ("sm__issue_active.avg.pct_of_peak_sustained_elapsed", "issue active % of elapsed"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps active %"),
    ("launch__registers_per_thread", "registers / thread"),
    ("launch__grid_size", "grid"),
    ("launch__block_size", "block"),
    ("smsp__inst_executed.sum", "warp instructions"),
    ("sm__inst_executed_pipe_uniform.sum", "uniform-pipe instructions"),
]
UNIT_BYTES = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}


def short_name(full):
    n = full.replace("void ", "").replace("simclr::", "")
    return n.split("(")[0]


def key_of(name):
    if "contrastive_tile_kernel" in name:
        args = name[name.index("<") + 1:name.index(">")].replace(" ", "").split(",")
        return "backward_tile" if args[2] in ("1", "true") else "forward_tile"
    for k in ("forward_finalize", "backward_finalize", "prepare", "backward_prepare"):
        if k + "_kernel" in name:
            return k
    return None


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("csvs", nargs="+")
    ap.add_argument("--out", default=os.path.join(REPO, "profiles", "ncu_summary.md"))
    ap.add_argument("--traffic", default=os.path.join(REPO, "profiles", "ncu_traffic.json"))
    ap.add_argument("--source", default="ncu --set full --clock-control none, tools/profile_step.py --steps 3 (cold, serialised)")
    args = ap.parse_args()
    lines = ["| kernel | " + " | ".join(label for _, label in METRICS) + " |", "|---|" + "---|" * len(METRICS)]
    traffic = {}
    for path in args.csvs:
        rows = list(csv.reader(open(path)))
        hdr, units = rows[0], rows[1]
        col = {h: i for i, h in enumerate(hdr)}
        for r in rows[2:]:
            name = r[col["Kernel Name"]]
            cells = []
            for m, _ in METRICS:
                if m in col and r[col[m]] != "":
                    cells.append(f"{r[col[m]]} {units[col[m]]}".strip())
                else:
                    cells.append("n/a")
            lines.append(f"| `{short_name(name)}` | " + " | ".join(cells) + " |")
            k = key_of(name)
            if k and k not in traffic and "dram__bytes_read.sum" in col:
                rd = float(r[col["dram__bytes_read.sum"]]) * UNIT_BYTES.get(units[col["dram__bytes_read.sum"]], 1.0)
                wr = float(r[col["dram__bytes_write.sum"]]) * UNIT_BYTES.get(units[col["dram__bytes_write.sum"]], 1.0)
                dur = float(r[col["gpu__time_duration.sum"]])
                traffic[k] = {"dram_bytes": int(rd + wr), "dram_read_bytes": int(rd), "dram_write_bytes": int(wr),
                              "duration": f"{dur} {units[col['gpu__time_duration.sum']]}", "kernel": short_name(name)}
    stamp = None
    try:
        stamp = open(os.path.join(REPO, "pytorch-simclr_b200", "lib", "libsimclr_b200.so.stamp")).read().strip()
    except OSError:
        pass
    with open(args.out, "w") as f:
        f.write("# ncu per-kernel summary\n\nSource: " + args.source + f"\n\nLibrary stamp: `{stamp}`\n\n" + "\n".join(lines) + "\n")
    with open(args.traffic, "w") as f:
        json.dump({"library_stamp": stamp, "source": args.source, "kernels": traffic}, f, indent=1)
    print("\n".join(lines))
    print(json.dumps(traffic, indent=1))


if __name__ == "__main__":
    sys.exit(main())
