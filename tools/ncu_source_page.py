"""What the ncu source page says about a warp-specialised kernel (tds_batch_kernel): hot code, who waits where.

    python tools/ncu_source_page.py gpurun_out/r2i_src_noise0.ncu-rep --steps 4411 --ctas 147 [--lib path/to/libafs_b200.so]

The capture: `ncu --section SourceCounters --import-source on -k regex:tds_batch -o <rep> python tools/prof_tds.py`
(T = 4410, i.e. 4411 loop passes).  Every SASS instruction of the report carries its executed count and the stall
samples taken at it; dividing the executed count by (CTAs x loop passes) gives the number of WARPS of a CTA that
execute the instruction once per step -- 8 for the dynamic chain warps, 6 for the static ones, 1 for X1 / X2 / the
code of one reduced-system chain, a loop count for loop bodies.  From that:
  * hot: the instructions executed at least every other step -- the code a step streams through the instruction
    caches -- in instructions, KB and 128-byte lines, per multiplicity class and (with --lib, which supplies the line
    table through nvdisasm) per source function;
  * stalls: stall reasons per multiplicity class;
  * barriers: the stall_barrier samples sit at the instruction BEHIND each BAR.SYNC; per barrier site, the share of
    a warp's time spent waiting there.  The warps that wait least at a barrier are the ones the others wait for.
This is how the longest paths of the round were found.
"""
import argparse
import collections
import csv
import io
import os
import re
import subprocess
import tempfile


def source_rows(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    k = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
    return rows[k - 1][1] if k else "?", rows[k], rows[k + 1:]


def line_table(lib, kernel_mangled_part, src_path):
    """function name of every SASS instruction of the kernel (innermost inlined function), from nvdisasm -g"""
    src = open(src_path).read().splitlines()
    starts = []
    for i, l in enumerate(src, 1):
        m = re.match(r"^(?:template.*>\s*)?(?:__device__|__global__|static|inline).*?\b([A-Za-z_0-9]+)\s*\(", l)
        if m and not l.strip().endswith(";"):
            starts.append((i, m.group(1)))

    def fn_of(line):
        name = "?"
        for s, n in starts:
            if s <= line:
                name = n
            else:
                break
        return name

    with tempfile.TemporaryDirectory() as d:
        subprocess.run(["cuobjdump", "-xelf", "tds_kernels", os.path.abspath(lib)], cwd=d, capture_output=True)
        cubin = [f for f in os.listdir(d) if f.endswith(".cubin")][0]
        sass = subprocess.run(["nvdisasm", "-g", "-c", os.path.join(d, cubin)], capture_output=True, text=True).stdout
    fns, inside, cur = [], False, "?"
    for l in sass.splitlines():
        if l.startswith(".text."):
            inside = kernel_mangled_part in l
            continue
        if not inside:
            continue
        m = re.search(r'//## File "([^"]+)", line (\d+)', l)
        if m:
            f = os.path.basename(m.group(1))
            cur = fn_of(int(m.group(2))) if f == os.path.basename(src_path) else f
            continue
        if re.match(r"\s+/\*[0-9a-f]{4,}\*/", l):
            fns.append(cur)
    return fns


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("rep")
    ap.add_argument("--steps", type=int, required=True, help="loop passes of the capture (T + 1)")
    ap.add_argument("--ctas", type=int, default=147)
    ap.add_argument("--lib", help="the library the capture ran (adds the per-function table)")
    ap.add_argument("--src", default=os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))),
                                                  "areafunctionsynthesis_b200", "csrc", "tds_kernels.cu"))
    a = ap.parse_args()
    kernel, hdr, data = source_rows(a.rep)
    ie, ib = hdr.index("Instructions Executed"), hdr.index("stall_barrier")
    names = hdr[hdr.index("stall_barrier"):hdr.index("stall_wait") + 1]
    i0 = hdr.index("stall_barrier")
    unit = a.ctas * a.steps
    mult = [int(r[ie]) / unit for r in data]
    total = sum(int(r[2]) for r in data)
    print(f"# {kernel}: {len(data)} instructions, {total} stall samples")

    def cls(m):
        return "cold" if m < 0.4 else "x1" if round(m) == 1 else "x8 (dynamic chain warps)" if round(m) == 8 else \
            "x5-7 (static chain warps)" if round(m) in (5, 6, 7) else "x2-4" if round(m) in (2, 3, 4) else "loops (> 8)"

    hot = [k for k, m in enumerate(mult) if m >= 0.5]
    print(f"hot code: {len(hot)} instructions = {len(hot) * 16 / 1024:.1f} KB, {len(set(k // 8 for k in hot))} lines of 128 B")
    agg = {}
    for k, r in enumerate(data):
        c = cls(mult[k])
        v = agg.setdefault(c, [0] * (len(names) + 2))
        for i in range(len(names)):
            v[i] += int(r[i0 + i])
        v[-2] += int(r[ie]); v[-1] += 1
    for c, v in sorted(agg.items()):
        s = sum(v[:-2])
        top = sorted(zip(names, v[:-2]), key=lambda t: -t[1])[:6]
        print(f"  {c:28s} {v[-1]:5d} instr, {v[-2] / unit:8.0f} warp-instr per step, {100 * s / max(total, 1):4.1f} % of the samples: "
              + ", ".join(f"{n[6:]} {100 * x / max(s, 1):.0f}%" for n, x in top))
    if a.lib:
        part = "tds_batch_kernelILb%dELb%dE" % (1 if "<1" in kernel.replace("(bool)", "") else 0, 1 if ", 1>" in kernel.replace("(bool)", "") else 0)
        fns = line_table(a.lib, part, a.src)
        if len(fns) == len(data):
            cnt, ex = collections.Counter(), collections.Counter()
            for k in hot:
                cnt[fns[k]] += 1; ex[fns[k]] += int(data[k][ie])
            print("hot instructions by (innermost) source function:")
            for f, v in cnt.most_common(40):
                print(f"  {v:5d} instr {ex[f] / unit:8.1f} warp-instr per step  {f}")
        else:
            print(f"(line table has {len(fns)} instructions, the report {len(data)}: not the library of the capture)")
    print("barrier waits (samples at the instruction behind a barrier; share of ONE such warp's time):")
    per_warp = total / 16
    for k, r in enumerate(data):
        if k and "BAR.SYNC" in data[k - 1][1] and mult[k] >= 0.5 and int(r[ib]) > 0:
            print(f"  @{k:5d} x{mult[k]:4.1f}  {100 * int(r[ib]) / per_warp / mult[k]:5.1f} %   after {data[k - 1][1].strip()[:40]}")
    # (a conditional branch around the barrier puts the samples one instruction further on)
    for k in sorted(sorted(range(len(data)), key=lambda k: -int(data[k][ib]))[:24]):
        if not (k and "BAR.SYNC" in data[k - 1][1]) and mult[k] >= 0.5:
            print(f"  @{k:5d} x{mult[k]:4.1f}  {100 * int(data[k][ib]) / per_warp / mult[k]:5.1f} %   at {data[k][1].strip()[:40]} (behind {data[k - 1][1].strip()[:30]})")


if __name__ == "__main__":
    main()
