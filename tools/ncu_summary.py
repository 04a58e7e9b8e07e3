"""Summarise an .ncu-rep (one `ncu --set full` capture) into the handful of numbers DESIGN.md / bench.py cite.

    python tools/ncu_summary.py gpurun_out/prof_tds.ncu-rep > profiles/r1_tds_batch_kernel.txt
"""
import csv
import re
import subprocess
import sys

KEYS = [
    r"^gpu__time_duration\.sum$", r"^launch__grid_size$", r"^launch__block_size$", r"^launch__registers_per_thread$",
    r"^launch__shared_mem_per_block_dynamic$", r"^launch__shared_mem_per_block_static$",
    r"^launch__occupancy_limit_(registers|shared_mem|warps)$",
    r"^dram__bytes_read\.sum$", r"^dram__bytes_write\.sum$", r"^dram__bytes_read\.sum\.per_second$",
    r"^gpu__dram_throughput\.avg\.pct_of_peak_sustained_elapsed$",
    r"^sm__warps_active\.avg\.pct_of_peak_sustained_active$", r"^sm__cycles_elapsed\.avg$",
    r"^smsp__inst_executed\.sum$", r"^smsp__issue_active\.avg\.pct_of_peak_sustained_active$",
    r"^smsp__thread_inst_executed_per_inst_executed\.ratio$",
    r"^sm__inst_executed_pipe_(fp64|lsu|alu|fma|xu|cbu|adu|uniform)\.avg\.pct_of_peak_sustained_active$",
    r"^sm__pipe_fp64_cycles_active\.avg\.pct_of_peak_sustained_(active|elapsed)$",
    r"^sm__pipe_tensor_cycles_active", r"^l1tex__data_pipe_lsu_wavefronts_mem_shared\.sum(\.pct_of_peak_sustained_elapsed)?$",
    r"^sass__inst_executed_local_(loads|stores)$", r"^sass__inst_executed_shared_(loads|stores)$",
    r"^smsp__average_warps_issue_stalled_.*_per_issue_active\.ratio$",
]


def main(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    for r in rows[2:]:
        d = dict(zip(hdr, r))
        print(f"# kernel: {d.get('Kernel Name')}  grid={d.get('Grid Size')} block={d.get('Block Size')}")
        for h, u in zip(hdr, units):
            if any(re.search(k, h) for k in KEYS):
                print(f"{h:95s} {d[h]:>20s} {u}")


if __name__ == "__main__":
    main(sys.argv[1])
