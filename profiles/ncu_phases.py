"""Group an ncu source-level profile into phases of the raster kernel by source line ranges."""
import csv, subprocess, sys, re
rep = sys.argv[1]
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
# phase map from function starts in the current source
import os
src = open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "gudni_b200/csrc/raster_device.cuh")).read().splitlines()
marks = []
names = {"threadGeom":"geom","addLineSegment":"gen.addLine","strandThresholds":"gen.strand","intersectCurve":"gen.bisect","buildThresholds":"gen.build","sortQueue":"sort",
 "compositeOverPremul":"color.composite","determineColor":"color.lane","insertSorted":"sweep.insert","splitNext":"sweep.split","sweepStep":"sweep.step","sweepColumn":"sweep.lane","rasterThread":"thread",
 "struct ChipQueue":"queue","struct WarpQueue":"queue","struct HbmQueue":"queue","struct ShapeStack":"stack","invSlope":"thr.math","intersectX":"thr.math","isBelow":"thr.isBelow","hPositive":"hdr","pixelWord":"pixel","tagMeta":"table","premultiply":"table","fillTileTable":"table","readPicture":"picture","nextPixel":"pixel"}
for i,l in enumerate(src,1):
    for k,v in names.items():
        if (("__device__" in l or l.startswith("struct") or l.startswith("static")) and k in l):
            marks.append((i,v)); break
marks.sort()
def phase(fname, line):
    if fname.endswith("raster_warp.cuh"):
        # anchors found by scanning the source for the section comments of sweepWarp
        wsrc = open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "gudni_b200/csrc/raster_warp.cuh")).read().splitlines()
        anchors = []
        for i, l in enumerate(wsrc, 1):
            for key, name in (("float4 denseColor(", "warp.color"), ("uint32_t stackHash(", "warp.hash"), ("int generateWarp(", "warp.generate"),
                              ("int sweepWarp(", "warp.setup"), ("// ---- flush", "warp.flush"), ("// ---- (A)", "warp.A-boundary"),
                              ("// ---- (B)", "warp.B-sections"), ("// ---- (C)", "warp.C-resolve"), ("// ---- (D)", "warp.D-accumulate")):
                if key in l: anchors.append((i, name))
        ph = "warp.other"
        for ln, name in anchors:
            if ln <= line: ph = name
        return ph
    if fname.endswith("raster_kernels.cu"): return "kernel"
    p="other"
    for ln,v in marks:
        if ln<=line: p=v
        else: break
    return p
hdr=None; cur=""; agg={}
for r in rows:
    if len(r)==2 and r[0]=="File Path": cur=r[1]; continue
    if r and r[0]=="Line No": hdr=r; continue
    if hdr and r and r[0].isdigit() and len(r)==len(hdr):
        d=dict(zip(hdr[4:],r[4:]))
        if not d["# Samples"].isdigit(): continue
        ph=phase(cur,int(r[0]))
        a=agg.setdefault(ph,[0,0,0])
        a[0]+=int(d["# Samples"]); a[1]+=int(d["Instructions Executed"]); a[2]+=int(d["Thread Instructions Executed"])
ts=sum(a[0] for a in agg.values()); ti=sum(a[1] for a in agg.values())
print(f"{'phase':18s} samples%  inst%   thr/inst")
for k,a in sorted(agg.items(), key=lambda kv:-kv[1][0]):
    print(f"{k:18s} {100*a[0]/ts:6.1f}  {100*a[1]/ti:6.1f}  {a[2]/max(a[1],1):6.1f}")
