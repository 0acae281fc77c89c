"""Aggregate an ncu source-level profile by enclosing function (found by scanning the CUDA sources for
`__device__` / `__global__` / struct headers).  usage: python profiles/ncu_funcs.py report.ncu-rep [kernel name regex]"""
import csv, os, re, subprocess, sys
rep = sys.argv[1]
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"] +
                     (["--kernel-name", f"regex:{sys.argv[2]}"] if len(sys.argv) > 2 else []), capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
marks = {}
def funcs(path):
    if path in marks: return marks[path]
    m = []
    try:
        src = open(path).read().splitlines()
    except OSError:
        src = []
    for i, l in enumerate(src, 1):
        if re.match(r"^\s*(template.*)?(static )?(__device__|__global__)", l) or re.match(r"^(template <.*> )?struct \w+", l):
            name = re.search(r"(\w+)\s*\(", l)
            sname = re.search(r"struct (\w+)", l)
            m.append((i, (name.group(1) if name and "struct" not in l else (sname.group(1) if sname else l.strip()[:30]))))
    marks[path] = m
    return m
def where(path, line):
    local = os.path.join(ROOT, "gudni_b200", "csrc", os.path.basename(path))
    name = "?"
    for ln, n in funcs(local if os.path.exists(local) else path):
        if ln <= line: name = n
        else: break
    return os.path.basename(path).split(".")[0][:12] + ":" + name
hdr = None; cur = ""; agg = {}
for r in rows:
    if len(r) == 2 and r[0] == "File Path": cur = r[1]; continue
    if r and r[0] == "Line No": hdr = r; continue
    if hdr and r and r[0].isdigit() and len(r) >= 9:
        d = dict(zip(hdr[4:], r[4:]))
        if not d["# Samples"].isdigit(): continue
        a = agg.setdefault(where(cur, int(r[0])), [0, 0, 0])
        a[0] += int(d["# Samples"]); a[1] += int(d["Instructions Executed"]); a[2] += int(d["Thread Instructions Executed"])
ts = sum(a[0] for a in agg.values()); ti = sum(a[1] for a in agg.values())
print(f"{'function':44s} samples%  inst%   thr/inst")
for k, a in sorted(agg.items(), key=lambda kv: -kv[1][0]):
    if a[0] * 200 < ts: continue
    print(f"{k:44s} {100*a[0]/ts:6.1f}  {100*a[1]/ti:6.1f}  {a[2]/max(a[1],1):6.1f}")
