"""Builds oracle/_ref/libgudni_ref.so: the reference's OWN kernel source compiled for the host CPU.

TEST INFRASTRUCTURE, NOT PRODUCT.  The reference's rasterizer is one OpenCL C file,
src/Graphics/Gudni/OpenCL/Kernels.cl, JIT-compiled by its Haskell host layer (OpenCL/Setup.hs:102-147).
Neither GHC nor an OpenCL runtime exists in this image, but the kernel file is plain C apart from
OpenCL's vector types and a handful of built-ins, so g++ can compile it against a small compatibility
header (cl_compat.hpp) and a harness that plays the NDRange (ref_harness.cpp).  The source is read where
it lies under /root/reference; nothing of it is copied into the repository — the rewritten text lives in
a temporary directory for the duration of the compile and only the shared object is kept, under
oracle/_ref/ (git-ignored; it travels to the GPU box with the other built .so files).

The text g++ sees differs from the file in exactly these mechanical ways:
  1. the first 40 lines are replaced by the #define block, as the reference itself does before
     handing the source to the OpenCL compiler (OpenCL/CppDefines.hs:66-70 appendCppDefines with
     sOURCEfILEpADDING = 40, Raster/Constants.hs:66; the list is OpenCL/Setup.hs:45-64 cppDefines).
     MAXTHRESHOLDS is bound to a variable of the harness instead of a literal so that one build serves
     every RasterSpec the tests use; the other values are the reference's constants;
  2. OpenCL vector literals `(float4)(a,b,c,d)` — in C++ a cast of a comma expression — become
     calls `mk_float4(a,b,c,d)` (also through the source's own aliases COLOR, THRESHOLD, SPACE2);
  3. `pos2 (x, y, width)` (Kernels.cl:813, dead code, parameters declared without types — not valid
     OpenCL C either) gets `int` parameters.
"""
import os
import re
import subprocess
import sys
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
ORACLE = os.path.dirname(HERE)
OUT_DIR = os.path.join(ORACLE, "_ref")
OUT = os.path.join(OUT_DIR, "libgudni_ref.so")
REFERENCE_ROOT = os.environ.get("GUDNI_REFERENCE_ROOT", "/root/reference")
KERNELS_CL = os.path.join(REFERENCE_ROOT, "src", "Graphics", "Gudni", "OpenCL", "Kernels.cl")
SOURCE_FILE_PADDING = 40          # Raster/Constants.hs:66

# OpenCL/Setup.hs:45-64 with the values of Raster/Constants.hs:47-89
DEFINES = [
    ("STOCHASTIC_FACTOR", "0.0"),
    ("RANDOMFIELDSIZE", "4096"),
    ("MAXTHRESHOLDS", "(cl_max_thresholds)"),   # spec value, supplied per call by the harness
    ("MAXSHAPE", "127"),
    ("SHAPESTACKSECTIONS", "8"),
    ("SHAPETAG_SUBSTANCETYPE_BITMASK", "0XC000000000000000"),
    ("SHAPETAG_SUBSTANCETYPE_SOLIDCOLOR", "0X8000000000000000"),
    ("SHAPETAG_SUBSTANCETYPE_PICTURE", "0X4000000000000000"),
    ("SHAPETAG_SUBSTANCETYPE_SHIFT", "30"),
    ("SHAPETAG_COMPOUNDTYPE_BITMASK", "0X3000000000000000"),
    ("SHAPETAG_COMPOUNDTYPE_CONTINUE", "0X1000000000000000"),
    ("SHAPETAG_COMPOUNDTYPE_ADD", "0X2000000000000000"),
    ("SHAPETAG_COMPOUNDTYPE_SUBTRACT", "0X3000000000000000"),
    ("SHAPETAG_COMPOUNDTYPE_SHIFT", "28"),
    ("SHAPETAG_SUBSTANCEID_BITMASK", "0XFFFFFFFFFFFFFFF"),
]

VECTOR_ALIASES = {"float2": "float2", "float4": "float4", "float8": "float8", "int2": "int2",
                  "uchar4": "uchar4", "COLOR": "float4", "THRESHOLD": "float4", "SPACE2": "float2",
                  "SPACE4": "float4"}
LITERAL = re.compile(r"\((%s)\)\s*\(" % "|".join(VECTOR_ALIASES))


def rewritten_source():
    with open(KERNELS_CL, "r", encoding="utf-8", errors="replace") as f:
        lines = f.read().split("\n")
    head = ["#define %s %s" % d for d in DEFINES]
    head += ["// Padding line "] * (SOURCE_FILE_PADDING - len(head))
    text = "\n".join(head + lines[SOURCE_FILE_PADDING:])
    text = LITERAL.sub(lambda m: "mk_%s(" % VECTOR_ALIASES[m.group(1)], text)
    text, n = re.subn(r"inline int pos2 \(x, y, width\)", "inline int pos2 (int x, int y, int width)", text)
    assert n == 1, "pos2 not found: reference source differs from the surveyed revision"
    return text


def opencl_source(max_thresholds=1024, strict=False):
    """The text the reference hands to the OpenCL compiler (OpenCL/Setup.hs:132 addDefinesToSource):
    Kernels.cl with its first 40 lines replaced by the #define block, MAXTHRESHOLDS a literal, and the
    two places a current OpenCL compiler rejects patched (pos2's untyped parameters, :813, and an
    address-space qualifier on a struct field, :420).  `strict` uses one of the padding lines
    for `#pragma OPENCL FP_CONTRACT OFF`."""
    with open(KERNELS_CL, "r", encoding="utf-8", errors="replace") as f:
        lines = f.read().split("\n")
    defines = [(k, str(int(max_thresholds)) if k == "MAXTHRESHOLDS" else v) for k, v in DEFINES]
    head = ["#define %s %s" % d for d in defines]
    if strict:
        head.append("#pragma OPENCL FP_CONTRACT OFF")
    head += ["// Padding line "] * (SOURCE_FILE_PADDING - len(head))
    text = "\n".join(head + lines[SOURCE_FILE_PADDING:])
    text, n = re.subn(r"inline int pos2 \(x, y, width\)", "inline int pos2 (int x, int y, int width)", text)
    assert n == 1
    # Kernels.cl:420 qualifies a struct FIELD with __private; NVIDIA's OpenCL 3.0 compiler rejects
    # that ("field may not be qualified with an address space"), so the qualifier goes
    text, n = re.subn(r"PMEM SHAPESTACK shapeStack\[SHAPESTACKSECTIONS\];", "SHAPESTACK shapeStack[SHAPESTACKSECTIONS];", text)
    assert n == 1
    return text


def write_opencl_blob(path=None, max_thresholds=(1024, 256)):
    """zlib-compressed program text for oracle/refbuild/ocl_run.py (the OpenCL runtime compiles from
    source; there is no offline compiler in this image).  Lives under oracle/_ref/ (git-ignored)."""
    import json
    import zlib
    path = path or os.path.join(OUT_DIR, "ocl_program.bin")
    os.makedirs(OUT_DIR, exist_ok=True)
    blob = json.dumps({str(m): {"reference": opencl_source(m, strict=False), "strict": opencl_source(m, strict=True)}
                       for m in max_thresholds}).encode()
    with open(path, "wb") as f:
        f.write(zlib.compress(blob, 9))
    return path


def compiler():
    return "/usr/bin/g++" if os.access("/usr/bin/g++", os.X_OK) else "g++"


def build(force=False, verbose=False):
    """Returns the path of the built library, or None when the reference tree is not present
    (the GPU box: only the prebuilt file is used there)."""
    if not os.path.exists(KERNELS_CL):
        return OUT if os.path.exists(OUT) else None
    deps = [KERNELS_CL, __file__, os.path.join(HERE, "cl_compat.hpp"), os.path.join(HERE, "ref_harness.cpp")]
    if not force and os.path.exists(OUT) and all(os.path.getmtime(OUT) >= os.path.getmtime(d) for d in deps):
        return OUT
    os.makedirs(OUT_DIR, exist_ok=True)
    with tempfile.TemporaryDirectory(prefix="gudni_ref_") as tmp:
        with open(os.path.join(tmp, "kernels_cl.inc"), "w") as f:
            f.write(rewritten_source())
        cmd = [compiler(), "-O3", "-std=gnu++17", "-fPIC", "-shared", "-fopenmp", "-ffp-contract=off",
               "-fno-fast-math", "-fno-strict-aliasing", "-w", "-fpermissive",
               "-I", tmp, "-I", HERE, "-I", os.path.join(os.path.dirname(ORACLE), "include"),
               "-o", OUT, os.path.join(HERE, "ref_harness.cpp")]
        if verbose:
            print(" ".join(cmd))
        subprocess.run(cmd, check=True)
    return OUT


if __name__ == "__main__":
    path = build(force="-B" in sys.argv, verbose=True)
    print(path if path else "reference tree not present and no prebuilt library")
