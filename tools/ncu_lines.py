"""aggregate an ncu --page source (cuda,sass) CSV by source line: stall samples, instructions, barrier stalls
usage: ncu_lines.py report.ncu-rep [top] [kernel-substring]"""
import csv, subprocess, sys
rep = sys.argv[1]; top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
kfilter = sys.argv[3] if len(sys.argv) > 3 else None
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
cur = None; agg = {}; hdr = None; fn = None
for r in rows:
    if len(r) >= 2 and r[0] == 'File Path': cur = r[1].split('/')[-1]; continue
    if len(r) >= 2 and r[0] == 'Function Name': fn = r[1]; continue
    if len(r) >= 2 and r[0] == 'Line No': hdr = r; continue
    if hdr is None or len(r) < 8 or r[2] != '-': continue
    if kfilter and (fn is None or kfilter not in fn): continue
    try: ln = int(r[0]); samples = int(r[6]); inst = int(r[7])
    except ValueError: continue
    a = agg.get((cur, ln), (0, 0, r[1][:110], 0))
    agg[(cur, ln)] = (a[0] + samples, a[1] + inst, r[1][:110], a[3] + int(r[hdr.index('stall_barrier')]))
tot = sum(v[0] for v in agg.values())
print('total samples', tot)
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
    print(f"{k[0]}:{k[1]:<4} {v[0]:5d} {100*v[0]/max(tot,1):5.1f}%  inst {v[1]:8d}  barrier {v[3]:5d} | {v[2]}")
