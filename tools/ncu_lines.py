"""Per CUDA-source-line totals (warp instructions executed, stall samples, shared wavefronts) from an
.ncu-rep captured with --import-source on (run here, no GPU needed).  usage: ncu_lines.py rep [topn]"""
import csv, subprocess, sys
rep = sys.argv[1]; topn = int(sys.argv[2]) if len(sys.argv) > 2 else 40
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
fname = "?"; hdr = None; lines = {}
for r in rows:
    if len(r) == 2 and r[0] == "File Path": fname = r[1].split("/")[-1]; continue
    if r and r[0] == "Line No": hdr = r; ci = hdr.index("Instructions Executed"); cs = hdr.index("# Samples"); cw = hdr.index("L1 Wavefronts Shared"); continue
    if hdr is None or len(r) != len(hdr): continue
    if r[0] != "":  # a source line header: its own totals are given on this row
        key = (fname, int(r[0])); src = r[1].strip()
        try: lines[key] = [float(r[ci]), float(r[cs]), float(r[cw]), src]
        except ValueError: pass
tot = sum(v[0] for v in lines.values()); ts = sum(v[1] for v in lines.values())
print(f"total warp instructions {tot:.0f}, samples {ts:.0f}")
print(f"{'file:line':28s} {'inst%':>6s} {'samp%':>6s} {'smem wf':>11s}  source")
for k, v in sorted(lines.items(), key=lambda kv: -kv[1][0])[:topn]:
    print(f"{k[0]+':'+str(k[1]):28s} {v[0]/tot*100:6.2f} {v[1]/max(ts,1)*100:6.2f} {v[2]:11.0f}  {v[3][:90]}")
