"""Summarise an .ncu-rep (one `ncu --set full` capture) into the handful of numbers DESIGN.md / bench.py cite.

    python tools/ncu_summary.py gpurun_out/prof_tds.ncu-rep > profiles/r1_tds_batch_kernel.txt
    python tools/ncu_summary.py gpurun_out/prof_tds.ncu-rep --traffic-key tds_noise_on --note "4096 x 44100 steps"
        additionally records dram__bytes_read.sum + dram__bytes_write.sum of the capture's LAST kernel under that key
        in profiles/traffic.json (what bench.py reports as roofline.traffic).
"""
import json
import os
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


def to_bytes(value, unit):
    mult = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}
    return float(value.replace(",", "")) * mult[unit]


def main(path, traffic_key=None, note=""):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    for r in rows[2:]:
        d = dict(zip(hdr, r))
        print(f"# kernel: {d.get('Kernel Name')}  grid={d.get('Grid Size')} block={d.get('Block Size')}")
        for h, u in zip(hdr, units):
            if any(re.search(k, h) for k in KEYS):
                print(f"{h:95s} {d[h]:>20s} {u}")
    if traffic_key:
        d = dict(zip(hdr, rows[-1]))
        u = dict(zip(hdr, units))
        rd = to_bytes(d["dram__bytes_read.sum"], u["dram__bytes_read.sum"])
        wr = to_bytes(d["dram__bytes_write.sum"], u["dram__bytes_write.sum"])
        tj = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "profiles", "traffic.json")
        try:
            t = json.load(open(tj))
        except Exception:
            t = {}
        t[traffic_key] = {"bytes": int(rd + wr), "read": int(rd), "write": int(wr), "kernel": d.get("Kernel Name"),
                          "source": f"dram__bytes_read.sum + dram__bytes_write.sum, ncu --set full, {os.path.basename(path)} ({note})"}
        json.dump(t, open(tj, "w"), indent=1, sort_keys=True)


if __name__ == "__main__":
    key = sys.argv[sys.argv.index("--traffic-key") + 1] if "--traffic-key" in sys.argv else None
    note = sys.argv[sys.argv.index("--note") + 1] if "--note" in sys.argv else ""
    main(sys.argv[1], key, note)
