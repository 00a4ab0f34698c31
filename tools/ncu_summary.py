#!/usr/bin/env python
"""Condense ncu reports into a few lines per kernel launch (run where `ncu` is installed: the GPU box or the build container).

    python tools/ncu_summary.py gpurun_out/r2_*.ncu-rep > profiles/round2_ncu_summary.txt

For every launch: duration, DRAM bytes read+written (the `traffic` of bench.py's roofline entries), DRAM throughput (% of peak),
L2 hit rate, achieved occupancy, registers/thread, grid/block, issue-slot utilisation, the top warp-stall reasons.
"""
import csv
import io
import subprocess
import sys

KEYS = [("gpu__time_duration.sum", "dur"), ("dram__bytes_read.sum", "dram_rd"), ("dram__bytes_write.sum", "dram_wr"),
        ("dram__throughput.avg.pct_of_peak_sustained_elapsed", "dram_pct"), ("lts__t_sector_hit_rate.pct", "l2_hit_pct"),
        ("sm__warps_active.avg.pct_of_peak_sustained_active", "occ_pct"), ("launch__registers_per_thread", "regs"),
        ("launch__grid_size", "grid"), ("launch__block_size", "block"), ("sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "fp64_pct"),
        ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue_pct"), ("launch__shared_mem_per_block_dynamic", "dsmem"),
        ("l1tex__t_sector_hit_rate.pct", "l1_hit_pct")]


def num(v):
    try:
        return float(str(v).replace(",", ""))
    except Exception:
        return None


def main():
    for rep in sys.argv[1:]:
        out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
        rows = list(csv.reader(io.StringIO(out)))
        if len(rows) < 3:
            print("# %s: no data" % rep)
            continue
        hdr, units = rows[0], rows[1]
        col = {h: i for i, h in enumerate(hdr)}
        print("# %s" % rep)
        for r in rows[2:]:
            name = r[col["Kernel Name"]] if "Kernel Name" in col else "?"
            d = {}
            for k, short in KEYS:
                if k in col:
                    d[short] = (num(r[col[k]]), units[col[k]])
            stalls = []
            for h, i in col.items():
                if h.startswith("smsp__average_warps_issue_stalled_") and h.endswith("_per_issue_active.ratio"):
                    v = num(r[i])
                    if v:
                        stalls.append((v, h[len("smsp__average_warps_issue_stalled_"):-len("_per_issue_active.ratio")]))
            stalls.sort(reverse=True)

            def g(k):
                v = d.get(k)
                return v[0] if v and v[0] is not None else float("nan")
            dur, du = d.get("dur", (None, ""))
            dur_ms = (dur / 1e6 if "ns" in du else (dur / 1e3 if "us" in du else dur)) if dur is not None else float("nan")

            def byt(k):
                v, u = d.get(k, (None, ""))
                if v is None:
                    return float("nan")
                return v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(u, 1)
            traffic = byt("dram_rd") + byt("dram_wr")
            print("%-70s %9.4f ms  dram %8.3f GB (rd %.3f wr %.3f) = %6.0f GB/s  dram%% %5.1f  L2hit %5.1f  L1hit %5.1f  occ%% %5.1f  issue%% %5.1f  regs %3.0f  grid %6.0f x %4.0f  smem %6.0f  stalls: %s"
                  % (name[:70], dur_ms, traffic / 1e9, byt("dram_rd") / 1e9, byt("dram_wr") / 1e9, traffic / 1e9 / (dur_ms * 1e-3) if dur_ms == dur_ms and dur_ms > 0 else float("nan"), g("dram_pct"), g("l2_hit_pct"), g("l1_hit_pct"),
                     g("occ_pct"), g("issue_pct"), g("regs"), g("grid"), g("block"), g("dsmem"), ", ".join("%s %.1f" % (s, v) for v, s in stalls[:3])))


if __name__ == "__main__":
    main()
