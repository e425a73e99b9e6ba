"""Markdown table of the metrics the roofline discussion uses, from one or more .ncu-rep files (read with `ncu -i`)."""
import csv
import subprocess
import sys

WANT = ["gpu__time_duration.sum", "sm__cycles_elapsed.max", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_tc_cycles_active.avg.pct_of_peak_sustained_active",
        "l1tex__data_pipe_tc_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct",
        "smsp__issue_inst0.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size",
        "launch__block_size", "launch__shared_mem_per_block_dynamic"]

for rep in sys.argv[1:]:
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units, body = rows[0], rows[1], rows[2:]
    name_i = hdr.index("Kernel Name")
    print(f"\n### {rep}\n")
    print("| metric | unit | " + " | ".join(r[name_i].split("(")[0].replace("void ", "")[:60] for r in body) + " |")
    print("|---|---|" + "---:|" * len(body))
    for w in WANT:
        if w in hdr:
            i = hdr.index(w)
            print(f"| `{w}` | {units[i]} | " + " | ".join(r[i] for r in body) + " |")
