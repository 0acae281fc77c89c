"""Summarise an ncu report per CUDA source line: samples, instructions executed, avg active threads.
usage: python tools_ncu_lines.py report.ncu-rep [top_n]"""
import csv, subprocess, sys
rep = sys.argv[1]; top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr = None; lines = []; cur_file = ""
for r in rows:
    if len(r) == 2 and r[0] == "File Path": cur_file = r[1].split("/")[-1]; continue
    if r and r[0] == "Line No": hdr = r; continue
    if hdr and r and r[0].isdigit() and len(r) >= 9:
        d = dict(zip(hdr[4:], r[4:]))
        if not d["# Samples"].isdigit(): continue
        lines.append((cur_file, int(r[0]), r[1].strip()[:90], int(d["# Samples"]), int(d["Instructions Executed"]), int(d["Thread Instructions Executed"]), d))
tot_s = sum(l[3] for l in lines); tot_i = sum(l[4] for l in lines)
print(f"total samples {tot_s}  total warp-inst {tot_i}")
for l in sorted(lines, key=lambda l: -l[3])[:top]:
    d = l[6]
    stalls = {k[6:]: int(v) for k, v in d.items() if k.startswith("stall_") and "Not Issued" not in k and v.isdigit() and int(v) > 0}
    topst = ",".join(f"{k}:{v}" for k, v in sorted(stalls.items(), key=lambda kv: -kv[1])[:3])
    print(f"{l[0]}:{l[1]:4d} smp {100*l[3]/tot_s:5.1f}% inst {100*l[4]/tot_i:5.1f}% thr/inst {l[5]/max(l[4],1):4.1f} [{topst}] {l[2]}")
