"""Turns ncu outputs into the committed summaries under profiles/.

  python scripts/ncu_summary.py rep   gpurun_out/prof.ncu-rep   profiles/name.md   ["title"]
  python scripts/ncu_summary.py list  gpurun_out/launches.csv   profiles/name.md   ["title"]
"""
import csv
import io
import subprocess
import sys
from collections import OrderedDict

KEYS = [
    ("gpu__time_duration.sum", "duration"),
    ("launch__grid_size", "grid"), ("launch__block_size", "block"), ("launch__registers_per_thread", "registers / thread"),
    ("launch__shared_mem_per_block_dynamic", "dynamic smem / block"), ("launch__occupancy_limit_registers", "occupancy limit: registers (blocks)"),
    ("launch__occupancy_limit_shared_mem", "occupancy limit: shared mem (blocks)"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "achieved occupancy % of peak warps"),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "SM issue utilisation (issue slots busy) %"),
    ("smsp__inst_executed.sum", "warp instructions executed"),
    ("smsp__thread_inst_executed_per_inst_executed.ratio", "active threads per warp instruction"),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "SM throughput % of peak"),
    ("dram__bytes_read.sum", "DRAM read"), ("dram__bytes_write.sum", "DRAM write"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "DRAM throughput % of peak"),
    ("dram__throughput.avg.pct_of_peak_sustained_elapsed", "DRAM throughput % of peak (dram__)"),
    ("lts__t_bytes.sum", "L2 bytes"), ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "L2 throughput % of peak"),
    ("lts__t_sector_hit_rate.pct", "L2 hit rate %"), ("l1tex__t_sector_hit_rate.pct", "L1 hit rate %"),
    ("l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "L1/TEX throughput % of peak"),
    ("l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "shared-memory bank conflicts"),
    ("smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "stall: long scoreboard (global/L2 loads)"),
    ("smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio", "stall: short scoreboard (shared mem)"),
    ("smsp__average_warps_issue_stalled_wait_per_issue_active.ratio", "stall: wait (fixed latency)"),
    ("smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio", "stall: not selected"),
    ("smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio", "stall: branch resolving"),
    ("smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio", "stall: math pipe throttle"),
    ("smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio", "stall: barrier"),
    ("smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio", "stall: no instruction"),
]


def rep(path, out, title):
    raw = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    col = {h: i for i, h in enumerate(hdr)}
    with open(out, "w") as f:
        f.write("# %s\n\nSource: `ncu --set full --clock-control none --import-source on` capture `%s` (%d launches captured; "
                "replayed, cold cache: absolute times are not bench values).\n\n" % (title, path.split("/")[-1], len(data)))
        names = [r[col["Kernel Name"]] for r in data]
        f.write("Kernel: `%s`\n\n| metric | " % names[0] + " | ".join("launch %d" % i for i in range(len(data))) + " | unit |\n|---|" + "---|" * (len(data) + 1) + "\n")
        for k, label in KEYS:
            if k in col:
                f.write("| %s (`%s`) | " % (label, k) + " | ".join(r[col[k]] for r in data) + " | %s |\n" % units[col[k]])
        # derived absolute rates against the B200 peaks (HBM 6544.3 GB/s measured copy; SM issue: 4 schedulers x 148 SMs)
        scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1e-9, "us": 1e-6, "ms": 1e-3, "s": 1.0}
        def val(r, k):
            return float(r[col[k]]) * scale.get(units[col[k]], 1.0)
        if all(k in col for k in ("gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum")):
            f.write("\n| derived | " + " | ".join("launch %d" % i for i in range(len(data))) + " |\n|---|" + "---|" * len(data) + "\n")
            f.write("| achieved HBM GB/s (of 6544.3 measured peak) | " + " | ".join("%.1f (%.1f %%)" % ((val(r, "dram__bytes_read.sum") + val(r, "dram__bytes_write.sum")) / val(r, "gpu__time_duration.sum") / 1e9,
                    100 * (val(r, "dram__bytes_read.sum") + val(r, "dram__bytes_write.sum")) / val(r, "gpu__time_duration.sum") / 6544.3e9) for r in data) + " |\n")
            if "lts__t_bytes.sum" in col:
                f.write("| achieved L2 GB/s | " + " | ".join("%.1f" % (val(r, "lts__t_bytes.sum") / val(r, "gpu__time_duration.sum") / 1e9) for r in data) + " |\n")
            elif "lts__t_sectors.sum" in col:      # 32-byte sectors through the L2 tag stage
                f.write("| achieved L2 GB/s (`lts__t_sectors.sum` x 32 B) | " + " | ".join("%.1f" % (float(r[col["lts__t_sectors.sum"]]) * 32 / val(r, "gpu__time_duration.sum") / 1e9) for r in data) + " |\n")
            if "smsp__inst_executed.sum" in col:
                f.write("| warp instructions issued per second (peak: 4 schedulers x 148 SMs x 1.965 GHz = 1163 G/s) | " + " | ".join("%.1f G/s" % (float(r[col["smsp__inst_executed.sum"]]) / val(r, "gpu__time_duration.sum") / 1e9) for r in data) + " |\n")

def launches(path, out, title):
    rows = [r for r in csv.reader(open(path)) if len(r) > 10 and r[0].isdigit()]
    agg = OrderedDict()
    for r in rows:
        name = r[4].split("(")[0]
        ns = float(r[-1])
        a = agg.setdefault(name, [0, 0.0])
        a[0] += 1
        a[1] += ns
    tot = sum(a[1] for a in agg.values())
    with open(out, "w") as f:
        f.write("# %s\n\nSource: `ncu --metrics gpu__time_duration.sum --clock-control none` launch list `%s` (%d launches; serialised, cold "
                "cache: compare SHARES, not absolutes).\n\n| kernel | launches | total ms | share | avg us |\n|---|---|---|---|---|\n" % (title, path.split("/")[-1], len(rows)))
        for name, (n, ns) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            f.write("| `%s` | %d | %.3f | %.1f %% | %.1f |\n" % (name, n, ns / 1e6, 100 * ns / tot, ns / n / 1e3))


if __name__ == "__main__":
    mode, src, dst = sys.argv[1:4]
    title = sys.argv[4] if len(sys.argv) > 4 else src
    (rep if mode == "rep" else launches)(src, dst, title)
