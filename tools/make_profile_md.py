"""Summarise an .ncu-rep (`ncu --set full`) as a markdown table: python tools/make_profile_md.py rep.ncu-rep "title" > out.md"""
import csv
import subprocess
import sys

rep, title = sys.argv[1], (sys.argv[2] if len(sys.argv) > 2 else "")
if rep.endswith(".csv"):           # `ncu -i rep --page raw --csv` already run on the GPU box (the .ncu-rep stays there)
    with open(rep) as f:
        out = f.read()
else:
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr, units = rows[0], rows[1]
COLS = [("gpu__time_duration.sum", "time"), ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed", "tensor pipe %"),
        ("sm__pipe_tc_cycles_active.avg.pct_of_peak_sustained_elapsed", "TC unit busy %"),
        ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue active %"),
        ("l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed", "smem LSU %"),
        ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "DRAM %"), ("dram__bytes_read.sum", "DRAM read"),
        ("dram__bytes_write.sum", "DRAM write"), ("launch__registers_per_thread", "regs"),
        ("launch__shared_mem_per_block_dynamic", "dyn smem"), ("launch__grid_size", "grid"),
        ("sm__warps_active.avg.pct_of_peak_sustained_active", "occupancy %")]
have = [(hdr.index(m), n, units[hdr.index(m)]) for m, n in COLS if m in hdr]
print(f"# {title}\n")
print(f"Source: `{rep.split('/')[-1]}` (`ncu --set full --clock-control none --import-source on`, one GPU, cold-cache replays: compare")
print("shares and pipe utilisations, not absolute times; bench.py's CUDA-event times are the reported numbers).\n")
print("| kernel | " + " | ".join(f"{n} [{u}]" if u else n for _, n, u in have) + " |")
print("|---|" + "---|" * len(have))
ki = hdr.index("Kernel Name")
for r in rows[2:]:
    name = r[ki].split("(")[0].replace("void ", "")
    vals = []
    for i, _, _ in have:
        v = r[i]
        try:
            f = float(v.replace(",", ""))
            v = f"{f:.1f}" if abs(f) < 1e6 else f"{f:.3g}"
        except ValueError:
            pass
        vals.append(v)
    print(f"| `{name[:60]}` | " + " | ".join(vals) + " |")
