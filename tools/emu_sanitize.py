"""The kernels under sanitizers, CPU only: builds the host SIMT emulation of the CUDA kernels
(tests/native/raster_emu.cpp) with -fsanitize=undefined or -fsanitize=address and runs the emulated-kernel
scenarios of tests/test_kernels_emulated.py through it (catalogue, circles, pictures, both spill paths, the wide
slice pass, launches in batches, levels 2 and 3, refused frames).  Global and shared memory are host heap / static arrays there, so an out-of-bounds access or a
misaligned vector load in a kernel is reported like any host bug.

    python tools/emu_sanitize.py ubsan|asan
    python tools/emu_sanitize.py coverage      # gcov line / branch coverage of the kernel sources by those scenarios
"""
import ctypes
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
MODES = {"ubsan": (["-fsanitize=undefined,float-cast-overflow", "-fno-sanitize-recover=undefined,float-cast-overflow"], "libubsan.so", {}),
         "asan": (["-fsanitize=address"], "libasan.so", {"ASAN_OPTIONS": "detect_leaks=0:detect_stack_use_after_return=0"})}


def scenario(lib_path):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from gudni_b200 import scenes
    from gudni_b200.formats import CSpec, RasterSpec
    import test_kernels_emulated as T
    c = ctypes
    L = ctypes.CDLL(lib_path)
    vp, i32, i64, sz = c.c_void_p, c.c_int, c.c_int64, c.c_size_t
    L.raster_emu_frame.argtypes = [vp, sz, vp, vp, vp, vp, i32, i32, c.POINTER(CSpec), vp, i64, vp, vp, i32, i64, vp, vp, vp, vp]
    L.raster_emu_scene.argtypes = [vp, sz, vp, i32, vp, i32, vp, vp, vp, vp, vp, vp, vp, i32, i32, c.POINTER(CSpec), vp,
                                   vp, sz, vp, i64, vp, i64, vp, i64, vp, vp, i64, vp, vp]
    for make in T.CATALOGUE:
        T.run(L, make())
    T.run(L, scenes.fuzzy_circles(150, 200, 150, 4, 40, 6))
    T.run(L, scenes.picture_scene(320, 300, flowers_size=(350, 200)))
    T.run(L, scenes.thin_rectangles(150, width=256, height=256, spacing=1.5, thickness=0.7, one_shape=True))
    T.run(L, scenes.fuzzy_circles(1200, 96, 96, 5, 50, 77))
    # runs past the slice kernel's scratch (raster_slice_wide_kernel: its list, its 72-entry scratch, the overflow beyond)
    for n in (13, 50, 72, 73):
        T.run(L, T.identical_shapes(n))
    # a launch in batches: own work cursors, own regions of the stack table and of the wide list
    for batches in (2, 3, 8):
        L.raster_emu_set_batches(batches)
        T.run(L, scenes.fuzzy_circles(150, 200, 150, 4, 40, 6))
        T.run(L, T.identical_shapes(40, width=300, height=40))
        T.run(L, scenes.picture_scene(320, 300, flowers_size=(350, 200)))
    L.raster_emu_set_batches(1)
    for level in (2, 3):
        T.run_scene(L, scenes.mixed_bag(100, 300, 200, 7003), level)
        T.run_scene(L, scenes.fuzzy_circles(400, 150, 130, 5, 40, 0x1234), level, RasterSpec(64, 64, 64, 256, 254, 127))
    import numpy as np
    # a point at infinity: strand_bounds_kernel's flag, tile_order_kernel emptying the launch's shape lists
    T.test_infinite_coordinate_is_refused_not_rasterized(L, np.inf, 0)
    T.test_infinite_coordinate_is_refused_not_rasterized(L, -np.inf, 1)
    for level in (2, 3):
        T.test_infinite_coordinate_is_refused_at_levels_2_and_3(L, level)
        T.run_scene(L, scenes.huge_boxes(), level)       # boxes beyond int32 root tiles
        for seed in range(12):                           # shapes of 1 to 1e30 pixels: no out-of-range float -> int anywhere
            T.run_scene(L, scenes.far_shapes(12, 150, 110, 0xFA50 + seed), level)
    print("sanitized run complete: no reports")


def coverage():
    import shutil
    work = "/tmp/gudni_emu_cov"
    shutil.rmtree(work, ignore_errors=True)
    os.makedirs(work)
    cxx = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++"
    lib = os.path.join(work, "libraster_emu_cov.so")
    subprocess.run([cxx, "-O0", "-g", "--coverage", "-std=c++17", "-fPIC", "-shared", "-ffp-contract=off", "-fno-fast-math", "-w",
                    "-I", os.path.join(ROOT, "tests", "native", "emu"), "-I", os.path.join(ROOT, "include"), "-o", lib,
                    os.path.join(ROOT, "tests", "native", "raster_emu.cpp")], check=True, cwd=work)
    subprocess.run([sys.executable, os.path.abspath(__file__), "--run", lib], check=True, cwd=work)
    out = subprocess.run(["gcov", "-b", "libraster_emu_cov.so-raster_emu.gcda"], cwd=work, capture_output=True, text=True).stdout
    keep = False
    for line in out.splitlines():
        if line.startswith("File "):
            keep = "/gudni_b200/csrc/" in line
        if keep and (line.startswith("File ") or line.startswith("Lines executed") or line.startswith("Taken at least once")):
            print(line)
    print("annotated sources:", work, "(*.gcov)")


def main():
    if len(sys.argv) == 3 and sys.argv[1] == "--run":
        return scenario(sys.argv[2])
    mode = sys.argv[1] if len(sys.argv) > 1 else "ubsan"
    if mode == "coverage":
        return coverage()
    flags, runtime, env = MODES[mode]
    lib = f"/tmp/libraster_emu_{mode}.so"
    cxx = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++"
    subprocess.run([cxx, "-O1", "-g", "-std=c++17", "-fPIC", "-shared", "-ffp-contract=off", "-fno-fast-math", "-w"] + flags +
                   ["-I", os.path.join(ROOT, "tests", "native", "emu"), "-I", os.path.join(ROOT, "include"), "-o", lib,
                    os.path.join(ROOT, "tests", "native", "raster_emu.cpp")], check=True)
    preload = subprocess.run([cxx, "-print-file-name=" + runtime], capture_output=True, text=True, check=True).stdout.strip()
    e = dict(os.environ, LD_PRELOAD=preload, **env)
    sys.exit(subprocess.run([sys.executable, os.path.abspath(__file__), "--run", lib], env=e).returncode)


if __name__ == "__main__":
    main()
