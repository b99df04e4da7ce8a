"""Summarise ncu reports (read on the CPU box): python scripts/summarize_ncu.py out.md rep1.ncu-rep [rep2 ...]
Prints one block per captured launch with the metrics the roofline discussion uses."""
import csv
import io
import subprocess
import sys

KEYS = [
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue slots busy %"),
    ("l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed", "L1 LSU data pipe %"),
    ("gpu__time_duration.sum", "duration"),
    ("dram__bytes_read.sum", "dram read"),
    ("dram__bytes_write.sum", "dram write"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram % of peak"),
    ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "L2 (lts) % of peak"),
    ("lts__t_sector_hit_rate.pct", "L2 hit rate %"),
    ("l1tex__m_xbar2l1tex_read_bytes.sum", "L2->SM bytes"),
    ("l1tex__m_xbar2l1tex_read_bytes.sum.per_second", "L2->SM rate"),
    ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed", "tensor pipe active % (elapsed)"),
    ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor pipe active % (SM-active)"),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "SM throughput %"),
    ("sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "XU (MUFU) pipe %"),
    ("sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "FMA pipe %"),
    ("sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "ALU pipe %"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "achieved occupancy %"),
    ("launch__registers_per_thread", "registers/thread"),
    ("launch__grid_size", "grid"),
    ("launch__block_size", "block"),
    ("sm__cycles_elapsed.max", "SM cycles"),
    ("sm__cycles_elapsed.max.per_second", "SM clock"),
]


def rows_of(rep):
    """``rep``: an .ncu-rep, or the ``--page raw --csv`` export of one (what scripts/gpu_profile.sh brings back for the big GEMM reports)."""
    if rep.endswith(".csv"):
        out = open(rep, errors="ignore").read()
    else:
        out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    return rows[0], rows[1], rows[2:]


def main():
    out_path, reps = sys.argv[1], sys.argv[2:]
    lines = []
    for rep in reps:
        hdr, units, rows = rows_of(rep)
        idx = {h: i for i, h in enumerate(hdr)}
        lines.append(f"## {rep.split('/')[-1]}\n")
        for r in rows:
            name = r[idx["Kernel Name"]]
            lines.append(f"### launch {r[idx['ID']]}: `{name[:110]}`\n")
            lines.append("| metric | value |\n|---|---|")
            for k, label in KEYS:
                if k in idx and r[idx[k]] != "":
                    lines.append(f"| {label} (`{k}`) | {r[idx[k]]} {units[idx[k]]} |")
            lines.append("")
    open(out_path, "w").write("\n".join(lines) + "\n")
    print("\n".join(lines))


if __name__ == "__main__":
    main()
