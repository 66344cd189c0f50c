"""Condense `ncu --page raw --csv` exports (gpurun_out/<tag>_<kernel>_raw.csv) into a small per-launch table.

    python profiles/summarize_ncu.py gpurun_out/r1b_conv_raw.csv [...] > profiles/<tag>_ncu_summary.md

Columns: duration, grid, regs, DRAM read+write per launch (the `traffic` of bench.py's roofline), achieved
occupancy, issue-slot utilisation, FMA-pipe utilisation, shared-memory wavefronts, top warp-stall reasons."""
import csv
import json
import sys

COLS = [
    ("gpu__time_duration.sum", "us"),
    ("launch__grid_size", "grid"),
    ("launch__registers_per_thread", "regs"),
    ("launch__shared_mem_per_block_dynamic", "dyn smem KB"),
    ("launch__waves_per_multiprocessor", "waves"),
    ("dram__bytes_read.sum", "DRAM rd MB"),
    ("dram__bytes_write.sum", "DRAM wr MB"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "occ %"),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue %"),
    ("sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "fma pipe %"),
    ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor pipe %"),
    ("sm__pipe_tensor_subpipe_utc_cycles_active.avg.pct_of_peak_sustained_active", "tcgen05 (UTC) pipe %"),
    ("sm__inst_executed_pipe_tensor.sum", "tensor insts"),
    ("l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed", "smem wavefront %"),
    ("l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "bank conflicts"),
    ("lts__t_sector_hit_rate.pct", "L2 hit %"),
]
STALLS = ["long_scoreboard", "short_scoreboard", "barrier", "mio_throttle", "math_pipe_throttle", "not_selected", "wait",
          "dispatch_stall", "lg_throttle", "no_instruction", "imc_miss"]


def num(v):
    try:
        return float(v.replace(",", ""))
    except ValueError:
        return None


def main(paths):
    traffic = {}
    for path in paths:
        rows = list(csv.reader(open(path)))
        hdr, units, data = rows[0], rows[1], rows[2:]
        ix = {h: i for i, h in enumerate(hdr)}
        print("### %s (%d launches)\n" % (path.split("/")[-1], len(data)))
        names = [c for c, _ in COLS if c in ix]
        print("| # | kernel | " + " | ".join(lbl for c, lbl in COLS if c in ix) + " | top stalls (warps per issue) |")
        print("|" + "---|" * (len(names) + 3))
        for n, r in enumerate(data):
            kn = r[ix["Kernel Name"]].replace("void ", "").split("(")[0]
            cells = []
            for c in names:
                v = num(r[ix[c]])
                u = units[ix[c]]
                if v is None:
                    cells.append(r[ix[c]])
                    continue
                if c.startswith("dram__bytes"):
                    v = v * {"Kbyte": 1e-3, "Mbyte": 1.0, "Gbyte": 1e3, "byte": 1e-6}.get(u, 1.0)
                if c == "gpu__time_duration.sum":
                    v = v * {"ns": 1e-3, "us": 1.0, "ms": 1e3}.get(u, 1.0)   # table is in microseconds
                cells.append("%.4g" % v)
            st = []
            for s in STALLS:
                k = "smsp__average_warps_issue_stalled_%s_per_issue_active.ratio" % s
                if k in ix and num(r[ix[k]]) is not None:
                    st.append((num(r[ix[k]]), s))
            st.sort(reverse=True)
            print("| %d | %s | %s | %s |" % (n, kn, " | ".join(cells), ", ".join("%s %.2f" % (s, v) for v, s in st[:3])))
            rd, wr = num(r[ix["dram__bytes_read.sum"]]), num(r[ix["dram__bytes_write.sum"]])
            sc = {"Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "byte": 1.0}
            if rd is not None and wr is not None:
                b = rd * sc.get(units[ix["dram__bytes_read.sum"]], 1) + wr * sc.get(units[ix["dram__bytes_write.sum"]], 1)
                traffic.setdefault(kn.split("<")[0].replace("paig::", ""), []).append(b)
        print()
    print("<!-- traffic (DRAM bytes, sum over the captured launches of each kernel): %s -->" %
          json.dumps({k: sum(v) for k, v in traffic.items()}))


if __name__ == "__main__":
    main(sys.argv[1:])
