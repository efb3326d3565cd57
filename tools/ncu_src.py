"""Summarises the ncu source page: top stall lines, instruction mix executed, shared-memory excess wavefronts."""
import csv, subprocess, sys, collections
rep = sys.argv[1]; topn = int(sys.argv[2]) if len(sys.argv) > 2 else 25
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr = rows[1]; col = {h: i for i, h in enumerate(hdr)}
body = [r for r in rows[2:] if len(r) == len(hdr)]
def num(r, k):
    try: return float(r[col[k]])
    except: return 0.0
tot_inst = sum(num(r, 'Instructions Executed') for r in body)
tot_samp = sum(num(r, '# Samples') for r in body)
print(f"total warp instructions executed {tot_inst:.0f}; samples {tot_samp:.0f}")
mix = collections.Counter()
for r in body:
    op = r[col['Source']].split()
    op = [o for o in op if not o.startswith('@')]
    name = op[0].split('.')[0] if op else '?'
    mix[name] += num(r, 'Instructions Executed')
print("executed mix:", ", ".join(f"{k}:{v/tot_inst*100:.1f}%" for k, v in mix.most_common(18)))
exc = [(num(r, 'L1 Wavefronts Shared Excessive'), num(r, 'L1 Wavefronts Shared'), r[col['Source']].strip()) for r in body]
exc = [e for e in exc if e[0] > 0]
print("smem excessive wavefronts:", sum(e[0] for e in exc), "of", sum(num(r, 'L1 Wavefronts Shared') for r in body))
for e in sorted(exc, reverse=True)[:10]: print("   ", e)
print("top stall lines (samples, instr, top reasons):")
stall_cols = [h for h in hdr if h.startswith('stall_')]
for r in sorted(body, key=lambda r: -num(r, '# Samples'))[:topn]:
    reasons = sorted(((num(r, s), s) for s in stall_cols), reverse=True)[:3]
    print(f"  {num(r,'# Samples'):7.0f} {r[col['Source']].strip()[:70]:70s} " + " ".join(f"{s[6:]}={v:.0f}" for v, s in reasons if v > 0))
agg = collections.Counter()
for r in body:
    for s in stall_cols: agg[s] += num(r, s)
print("stall totals:", ", ".join(f"{k[6:]}:{v/tot_samp*100:.1f}%" for k, v in agg.most_common(10)))
