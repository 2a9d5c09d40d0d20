"""Turn ncu outputs into the markdown / json summaries kept under profiles/.

    python profiles/summarize.py launches <launches.csv> <out.md>
    python profiles/summarize.py kernel <report.ncu-rep> <kernel-substring> <pixels> <out.md> [traffic.json]
"""
import csv
import json
import subprocess
import sys


def launches(path, out):
    rows = [r for r in csv.reader(open(path, errors="replace")) if len(r) > 5]
    hdr = rows[0]
    ik, iv = hdr.index("Kernel Name"), hdr.index("Metric Value")
    agg = {}
    for r in rows[1:]:
        try:
            v = float(r[iv].replace(",", ""))
        except ValueError:
            continue
        agg.setdefault(r[ik].split("(")[0], []).append(v)
    tot = sum(sum(v) for v in agg.values())
    with open(out, "w") as fh:
        fh.write("| kernel | launches | total ms | share | mean ms |\n|---|---:|---:|---:|---:|\n")
        for k, v in sorted(agg.items(), key=lambda kv: -sum(kv[1])):
            fh.write(f"| `{k}` | {len(v)} | {sum(v)/1e6:.3f} | {sum(v)/tot*100:.1f}% | {sum(v)/len(v)/1e6:.4f} |\n")
        fh.write(f"\ntotal device time {tot/1e6:.2f} ms over {sum(len(v) for v in agg.values())} launches "
                 "(ncu serialises launches and runs them cold: compare SHARES, not absolutes)\n")


def kernel(rep, name, pixels, out, traffic_json=None):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr = rows[0]
    pick = None
    for r in rows[2:]:
        if name in r[hdr.index("Kernel Name")]:
            pick = r
            break
    want = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "smsp__inst_executed.sum",
            "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic", "launch__grid_size", "launch__block_size",
            "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
            "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_bytes.sum",
            "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
            "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
            "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
            "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
            "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed",
            "smsp__warps_active.avg.per_cycle_active", "smsp__warps_eligible.avg.per_cycle_active",
            "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
            "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
            "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
            "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
            "smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio",
            "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio",
            "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
            "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio",
            "smsp__average_warps_issue_stalled_dispatch_stall_per_issue_active.ratio",
            "sm__cycles_active.avg", "sm__cycles_elapsed.avg"]
    units = rows[1]
    vals = {}
    with open(out, "w") as fh:
        fh.write(f"kernel `{pick[hdr.index('Kernel Name')]}` ({rep})\n\n| metric | value | unit |\n|---|---:|---|\n")
        for w in want:
            if w in hdr:
                i = hdr.index(w)
                vals[w] = pick[i]
                fh.write(f"| {w} | {pick[i]} | {units[i]} |\n")
        def num(k):
            return float(vals[k].replace(",", ""))
        scale = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0}
        rd = num("dram__bytes_read.sum") * scale[units[hdr.index("dram__bytes_read.sum")]]
        wr = num("dram__bytes_write.sum") * scale[units[hdr.index("dram__bytes_write.sum")]]
        fh.write(f"\nper pixel ({pixels} px): DRAM read {rd/pixels:.3f} B, write {wr/pixels:.3f} B, "
                 f"thread-instructions {num('smsp__inst_executed.sum')*32/pixels:.1f}\n")
    if traffic_json:
        json.dump({"kernel": name, "pixels": pixels, "dram_bytes_per_px": (rd + wr) / pixels,
                   "dram_read_bytes_per_px": rd / pixels, "dram_write_bytes_per_px": wr / pixels,
                   "thread_instructions_per_px": num("smsp__inst_executed.sum") * 32 / pixels, "source": rep},
                  open(traffic_json, "w"), indent=1)


if __name__ == "__main__":
    if sys.argv[1] == "launches":
        launches(sys.argv[2], sys.argv[3])
    else:
        kernel(sys.argv[2], sys.argv[3], int(sys.argv[4]), sys.argv[5], sys.argv[6] if len(sys.argv) > 6 else None)
