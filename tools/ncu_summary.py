#!/usr/bin/env python
"""Turn `ncu --set full` reports into the markdown tables kept under profiles/.

    python tools/ncu_summary.py gpurun_out/x.ncu-rep [more.ncu-rep ...] > profiles/rNN_ncu_full.md

One row per profiled launch: duration, DRAM bytes read / written (the `roofline.traffic` figure of
bench.py), DRAM % of ncu's own peak, occupancy, registers, issue-slot and FP64-pipe utilisation,
L1 / L2 sector hit rates, sectors per L1 request (scatter of the gathers) and the top warp stall.
Needs `ncu` on PATH (reads the report with `--page raw --csv`); no GPU.
"""
import csv
import io
import subprocess
import sys

COLS = [
    ("Kernel Name", "kernel", None),
    ("Block Size", "block", None),
    ("Grid Size", "grid", None),
    ("gpu__time_duration.sum", "ms", "ms"),
    ("dram__bytes_read.sum", "DRAM read GB", "GB"),
    ("dram__bytes_write.sum", "DRAM write GB", "GB"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "DRAM %", "f1"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps active %", "f1"),
    ("launch__registers_per_thread", "regs", "i"),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue active %", "f1"),
    ("sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "fp64 pipe %", "f1"),
    ("l1tex__t_sector_hit_rate.pct", "L1 hit %", "f1"),
    ("lts__t_sector_hit_rate.pct", "L2 hit %", "f1"),
    (("l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum", "l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum"),
     "sectors / global-load request", "ratio"),
]

UNIT_SCALE = {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3,
              "byte": 1e-9, "Kbyte": 1e-6, "Mbyte": 1e-3, "Gbyte": 1.0, "Tbyte": 1e3}


def raw_rows(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True,
                         check=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    return rows[0], rows[1], rows[2:]


def stall_columns(hdr):
    """Per-warp stall reasons (pc-sampling ratios), whatever this ncu version calls them."""
    cols = {}
    for i, h in enumerate(hdr):
        if h.startswith("smsp__average_warps_issue_stalled_") and h.endswith("_per_issue_active.ratio"):
            name = h[len("smsp__average_warps_issue_stalled_"):-len("_per_issue_active.ratio")]
            if not name.endswith("_not_issued"):
                cols[name] = i
    return cols


def fmt(val, unit, kind):
    if kind is None:
        return val.replace("void ", "").replace("(SpmvArgs)", "").replace("(OrthoArgs)", "") \
                  .replace("(RestartArgs)", "").strip()
    try:
        v = float(val.replace(",", ""))
    except ValueError:
        return val
    if kind in ("ms", "GB"):
        v *= UNIT_SCALE.get(unit, 1.0)
        return f"{v:.4f}" if kind == "ms" else f"{v:.3f}"
    if kind == "i":
        return str(int(v))
    return f"{v:.1f}"


def main():
    for path in sys.argv[1:]:
        hdr, units, rows = raw_rows(path)
        idx = []
        for c, title, kind in COLS:
            if kind == "ratio":
                i = (hdr.index(c[0]), hdr.index(c[1])) if c[0] in hdr and c[1] in hdr else -1
            else:
                i = hdr.index(c) if c in hdr else -1
            idx.append((i, title, kind))
        stalls = stall_columns(hdr)
        print(f"### `{path.split('/')[-1]}`\n")
        print("| " + " | ".join(t for _, t, _ in idx) + " | top stalls (warps per issue) |")
        print("|" + "---|" * (len(idx) + 1))
        for r in rows:
            cells = []
            for i, _, k in idx:
                if i == -1:
                    cells.append("-")
                elif k == "ratio":
                    den = float(r[i[1]].replace(",", "") or 0)
                    cells.append(f"{float(r[i[0]].replace(',', '')) / den:.1f}" if den else "-")
                else:
                    cells.append(fmt(r[i], units[i], k))
            cells[0] = "`" + cells[0] + "`"
            top = sorted(((float(r[i]) if r[i] else 0.0, n) for n, i in stalls.items()), reverse=True)[:3]
            cells.append(", ".join(f"{n} {v:.2f}" for v, n in top))
            print("| " + " | ".join(cells) + " |")
        print()


if __name__ == "__main__":
    main()
