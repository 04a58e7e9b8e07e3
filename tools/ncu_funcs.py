"""Instructions / stall samples per device function from an ncu --page source csv of tds_kernels.cu."""
import csv, collections, re, sys
rows = list(csv.reader(open(sys.argv[1])))
steps = float(sys.argv[2]) if len(sys.argv) > 2 else 1.0
src = open(sys.argv[3] if len(sys.argv) > 3 else 'areafunctionsynthesis_b200/csrc/tds_kernels.cu').read().split('\n')
func = {}; cur = '?'
for i, l in enumerate(src, 1):
    if l.startswith('__device__') or l.startswith('__global__') or l.startswith('tds_batch_kernel'):
        m = re.search(r'(\w+)\(', l)
        if m: cur = m.group(1)
    func[i] = cur
hdr = None; tot = collections.Counter(); stall = collections.Counter()
for r in rows:
    if r and r[0] == "Line No": hdr = r; continue
    if hdr is None or len(r) < len(hdr): continue
    d = dict(zip(hdr, r))
    try: ln = int(d["Line No"])
    except ValueError: continue
    def num(k):
        try: return int(float(d.get(k, "0") or "0"))
        except ValueError: return 0
    tot[func.get(ln, '?')] += num("Instructions Executed"); stall[func.get(ln, '?')] += num("Warp Stall Sampling (All Samples)")
T = sum(tot.values()); S = sum(stall.values())
for k, v in tot.most_common(28):
    print(f"{k:28s} {v/steps:9.0f} inst/step {100*v/T:5.1f}%  stall {100*stall[k]/max(S,1):5.1f}%")
