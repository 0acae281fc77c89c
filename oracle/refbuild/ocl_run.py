"""Runs the reference's kernels VERBATIM through an OpenCL runtime (NVIDIA's, on the B200 box) and
compares them with the restated oracle.  TEST INFRASTRUCTURE / MEASUREMENT TOOL, NOT PRODUCT.

  python oracle/refbuild/ocl_run.py --out gpurun_out/ocl_reference.json [--scenes small,s4b,s4]

The image has an ICD loader (CUDA's libOpenCL.so.1) but no vendor file; if the driver mount carries
libnvidia-opencl.so.1 the loader is pointed at it with OCL_ICD_FILENAMES.  The program text comes from
oracle/_ref/ocl_program.bin (build_ref.write_opencl_blob, made where /root/reference exists).  Two
builds: "reference" with the reference's own options (-cl-fast-relaxed-math -cl-strict-aliasing,
OpenCL/Setup.hs:126-129) and "strict" (FP_CONTRACT OFF, correctly rounded divide, no fast math).
Host side follows generateCall (OpenCL/CallKernels.hs:88-179): per job four scratch buffers, three
launches over Work2D numTiles threadsPerTile with work-group [1, threadsPerTile].
Every step is logged into the output JSON so a failure half-way still tells how far it got.
"""
import argparse
import ctypes as C
import glob
import json
import os
import sys
import time
import zlib

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

CL_DEVICE_TYPE_ALL = 0xFFFFFFFF
CL_MEM_READ_WRITE, CL_MEM_READ_ONLY, CL_MEM_COPY_HOST_PTR = 1, 4, 32
CL_PROGRAM_BUILD_LOG = 0x1183
CL_DEVICE_NAME, CL_DEVICE_VERSION, CL_DRIVER_VERSION = 0x102B, 0x102F, 0x102D
CL_DEVICE_MAX_WORK_GROUP_SIZE, CL_DEVICE_MAX_MEM_ALLOC_SIZE, CL_DEVICE_MAX_COMPUTE_UNITS = 0x1004, 0x1010, 0x1002
CL_PLATFORM_NAME = 0x0902

LOG = {"steps": []}


def step(msg, **kw):
    LOG["steps"].append({"t": round(time.time() - T0, 3), "msg": msg, **kw})
    print(f"[ocl {time.time() - T0:7.2f}s] {msg} {kw if kw else ''}", flush=True)


def flush_log(path):
    with open(path, "w") as f:
        json.dump(LOG, f, indent=1)


def find_vendor_library():
    pats = ["/usr/lib/x86_64-linux-gnu/libnvidia-opencl.so*", "/usr/lib64/libnvidia-opencl.so*",
            "/usr/local/nvidia/lib64/libnvidia-opencl.so*", "/usr/lib/libnvidia-opencl.so*",
            "/usr/local/cuda/compat/libnvidia-opencl.so*"]
    found = []
    for p in pats:
        found += glob.glob(p)
    return sorted(set(found))


def find_loader():
    for p in ["/usr/local/cuda/lib64/libOpenCL.so.1", "/usr/local/cuda/targets/x86_64-linux/lib/libOpenCL.so.1",
              "/usr/lib/x86_64-linux-gnu/libOpenCL.so.1"]:
        if os.path.exists(p):
            return p
    return None


class CLError(RuntimeError):
    pass


def chk(rc, what):
    if rc != 0:
        raise CLError(f"{what} failed with {rc}")


class CL:
    def __init__(self, loader):
        L = self.L = C.CDLL(loader)
        vp, u32, u64, sz = C.c_void_p, C.c_uint32, C.c_uint64, C.c_size_t
        pi32 = C.POINTER(C.c_int32)
        L.clGetPlatformIDs.argtypes = [u32, C.POINTER(vp), C.POINTER(u32)]
        L.clGetPlatformInfo.argtypes = [vp, u32, sz, vp, C.POINTER(sz)]
        L.clGetDeviceIDs.argtypes = [vp, u64, u32, C.POINTER(vp), C.POINTER(u32)]
        L.clGetDeviceInfo.argtypes = [vp, u32, sz, vp, C.POINTER(sz)]
        L.clCreateContext.restype = vp
        L.clCreateContext.argtypes = [vp, u32, C.POINTER(vp), vp, vp, pi32]
        L.clCreateCommandQueue.restype = vp
        L.clCreateCommandQueue.argtypes = [vp, vp, u64, pi32]
        L.clCreateBuffer.restype = vp
        L.clCreateBuffer.argtypes = [vp, u64, sz, vp, pi32]
        L.clReleaseMemObject.argtypes = [vp]
        L.clCreateProgramWithSource.restype = vp
        L.clCreateProgramWithSource.argtypes = [vp, u32, C.POINTER(C.c_char_p), C.POINTER(sz), pi32]
        L.clBuildProgram.argtypes = [vp, u32, C.POINTER(vp), C.c_char_p, vp, vp]
        L.clGetProgramBuildInfo.argtypes = [vp, vp, u32, sz, vp, C.POINTER(sz)]
        L.clCreateKernel.restype = vp
        L.clCreateKernel.argtypes = [vp, C.c_char_p, pi32]
        L.clSetKernelArg.argtypes = [vp, u32, sz, vp]
        L.clEnqueueNDRangeKernel.argtypes = [vp, vp, u32, C.POINTER(sz), C.POINTER(sz), C.POINTER(sz), u32, vp, vp]
        L.clEnqueueReadBuffer.argtypes = [vp, vp, u32, sz, sz, vp, u32, vp, vp]
        L.clEnqueueWriteBuffer.argtypes = [vp, vp, u32, sz, sz, vp, u32, vp, vp]
        L.clFinish.argtypes = [vp]
        for f in ("clGetPlatformIDs", "clGetPlatformInfo", "clGetDeviceIDs", "clGetDeviceInfo", "clReleaseMemObject",
                  "clBuildProgram", "clGetProgramBuildInfo", "clSetKernelArg", "clEnqueueNDRangeKernel",
                  "clEnqueueReadBuffer", "clEnqueueWriteBuffer", "clFinish"):
            getattr(L, f).restype = C.c_int32

    def info_str(self, fn, obj, param):
        n = C.c_size_t()
        chk(fn(obj, param, 0, None, C.byref(n)), "info size")
        buf = C.create_string_buffer(n.value + 1)
        chk(fn(obj, param, n.value, buf, None), "info")
        return buf.value.decode(errors="replace")

    def open(self):
        L = self.L
        n = C.c_uint32()
        chk(L.clGetPlatformIDs(0, None, C.byref(n)), "clGetPlatformIDs(count)")
        plats = (C.c_void_p * n.value)()
        chk(L.clGetPlatformIDs(n.value, plats, None), "clGetPlatformIDs")
        step("platforms", names=[self.info_str(L.clGetPlatformInfo, p, CL_PLATFORM_NAME) for p in plats])
        self.platform = plats[0]
        chk(L.clGetDeviceIDs(self.platform, CL_DEVICE_TYPE_ALL, 0, None, C.byref(n)), "clGetDeviceIDs(count)")
        devs = (C.c_void_p * n.value)()
        chk(L.clGetDeviceIDs(self.platform, CL_DEVICE_TYPE_ALL, n.value, devs, None), "clGetDeviceIDs")
        self.device = C.c_void_p(devs[0])
        wg, alloc, cu = C.c_size_t(), C.c_uint64(), C.c_uint32()
        L.clGetDeviceInfo(self.device, CL_DEVICE_MAX_WORK_GROUP_SIZE, 8, C.byref(wg), None)
        L.clGetDeviceInfo(self.device, CL_DEVICE_MAX_MEM_ALLOC_SIZE, 8, C.byref(alloc), None)
        L.clGetDeviceInfo(self.device, CL_DEVICE_MAX_COMPUTE_UNITS, 4, C.byref(cu), None)
        self.device_info = {"name": self.info_str(L.clGetDeviceInfo, self.device, CL_DEVICE_NAME),
                            "version": self.info_str(L.clGetDeviceInfo, self.device, CL_DEVICE_VERSION),
                            "driver": self.info_str(L.clGetDeviceInfo, self.device, CL_DRIVER_VERSION),
                            "max_work_group_size": wg.value, "max_mem_alloc": alloc.value, "compute_units": cu.value}
        step("device", **self.device_info)
        err = C.c_int32()
        devarr = (C.c_void_p * 1)(self.device.value)
        self.ctx = L.clCreateContext(None, 1, devarr, None, None, C.byref(err))
        chk(err.value, "clCreateContext")
        self.queue = L.clCreateCommandQueue(self.ctx, self.device, 0, C.byref(err))
        chk(err.value, "clCreateCommandQueue")

    def build(self, source, options):
        L = self.L
        err = C.c_int32()
        src = source.encode()
        arr = (C.c_char_p * 1)(src)
        lens = (C.c_size_t * 1)(len(src))
        prog = L.clCreateProgramWithSource(self.ctx, 1, arr, lens, C.byref(err))
        chk(err.value, "clCreateProgramWithSource")
        devarr = (C.c_void_p * 1)(self.device.value)
        t0 = time.time()
        rc = L.clBuildProgram(prog, 1, devarr, options.encode(), None, None)
        n = C.c_size_t()
        L.clGetProgramBuildInfo(prog, self.device, CL_PROGRAM_BUILD_LOG, 0, None, C.byref(n))
        buf = C.create_string_buffer(n.value + 1)
        L.clGetProgramBuildInfo(prog, self.device, CL_PROGRAM_BUILD_LOG, n.value, buf, None)
        log = buf.value.decode(errors="replace")
        step("clBuildProgram", rc=rc, seconds=round(time.time() - t0, 2), options=options, log_tail=log[-1500:])
        chk(rc, "clBuildProgram")
        kernels = {}
        for name in ("generateThresholds", "sortThresholds", "renderThresholds"):
            kernels[name] = L.clCreateKernel(prog, name.encode(), C.byref(err))
            chk(err.value, "clCreateKernel " + name)
        return kernels

    def buffer(self, nbytes, host=None, flags=CL_MEM_READ_WRITE):
        err = C.c_int32()
        nbytes = max(int(nbytes), 16)
        if host is not None:
            host = np.ascontiguousarray(host)
            if host.nbytes < nbytes:                      # empty / tiny inputs: padded to 16 bytes
                host = np.concatenate([np.frombuffer(host.tobytes(), np.uint8), np.zeros(nbytes - host.nbytes, np.uint8)])
            m = self.L.clCreateBuffer(self.ctx, flags | CL_MEM_COPY_HOST_PTR, host.nbytes, host.ctypes.data, C.byref(err))
        else:
            m = self.L.clCreateBuffer(self.ctx, flags, nbytes, None, C.byref(err))
        chk(err.value, f"clCreateBuffer({nbytes})")
        return C.c_void_p(m)

    def release(self, m):
        self.L.clReleaseMemObject(m)

    def set_args(self, kernel, args):
        for i, a in enumerate(args):
            if isinstance(a, C.c_void_p):               # cl_mem
                chk(self.L.clSetKernelArg(kernel, i, 8, C.byref(a)), f"clSetKernelArg {i}")
            else:
                a = np.ascontiguousarray(a)
                chk(self.L.clSetKernelArg(kernel, i, a.nbytes, a.ctypes.data), f"clSetKernelArg {i}")

    def launch(self, kernel, n_tiles, threads):
        gws = (C.c_size_t * 2)(n_tiles, threads)
        lws = (C.c_size_t * 2)(1, threads)
        chk(self.L.clEnqueueNDRangeKernel(self.queue, kernel, 2, None, gws, lws, 0, None, None), "clEnqueueNDRangeKernel")

    def finish(self):
        chk(self.L.clFinish(self.queue), "clFinish")

    def read(self, m, arr):
        chk(self.L.clEnqueueReadBuffer(self.queue, m, 1, 0, arr.nbytes, arr.ctypes.data, 0, None, None), "clEnqueueReadBuffer")


def render_opencl(cl, kernels, scene, jobs, spec, taps=True):
    """queueRasterJobs + raster + generateCall.  Returns image, per-job counts/bits, seconds in kernels."""
    i32 = np.int32
    geo = cl.buffer(0, scene.geometry, CL_MEM_READ_ONLY)
    sub = cl.buffer(0, np.ascontiguousarray(scene.substances, dtype=np.float32), CL_MEM_READ_ONLY)
    pict = cl.buffer(0, scene.picture_bytes, CL_MEM_READ_ONLY)
    uses = cl.buffer(0, scene.picture_uses, CL_MEM_READ_ONLY)
    rnd = cl.buffer(0, np.zeros(4096, np.float32), CL_MEM_READ_ONLY)
    out_host = np.zeros((scene.height, scene.width), np.uint32)
    out = cl.buffer(0, out_host)
    bitmap = np.array([scene.width, scene.height], i32)
    depth = np.array([int(np.log2(spec.threads_per_tile))], i32)
    zero = np.array([0], i32)
    bg = np.ascontiguousarray(scene.background, dtype=np.float32)
    G, maxT = spec.threads_per_tile, spec.max_thresholds
    counts, bits = [], []
    t_kernels = 0.0
    for index, job in enumerate(jobs):
        cols = job.columns
        thr = cl.buffer(cols * maxT * 16)
        hdr = cl.buffer(cols * maxT * 4)
        shs = cl.buffer(cols * 1088)
        qsl = cl.buffer(cols * 8)
        shapes = cl.buffer(0, job.shapes, CL_MEM_READ_ONLY)
        tiles = cl.buffer(0, job.tiles, CL_MEM_READ_ONLY)
        jobi = np.array([index], i32)
        nt = len(job.tiles)
        cl.finish()
        t0 = time.perf_counter()
        cl.set_args(kernels["generateThresholds"], [geo, shapes, tiles, bitmap, depth, zero, jobi, thr, hdr, shs, qsl])
        cl.launch(kernels["generateThresholds"], nt, G)
        if taps:
            cl.finish()
            t_kernels += time.perf_counter() - t0
            q = np.zeros((cols, 2), i32)
            s = np.zeros((cols, 1088 // 4), np.uint32)
            cl.read(qsl, q)
            cl.read(shs, s)
            counts.append((q, s[:, 0].astype(i32)))
            t0 = time.perf_counter()
        cl.set_args(kernels["sortThresholds"], [thr, hdr, qsl, tiles, bitmap, depth, zero, jobi])
        cl.launch(kernels["sortThresholds"], nt, G)
        cl.set_args(kernels["renderThresholds"], [thr, hdr, shs, qsl, sub, pict, uses, rnd, shapes, tiles, bg, bitmap,
                                                   depth, zero, jobi, out])
        cl.launch(kernels["renderThresholds"], nt, G)
        cl.finish()
        t_kernels += time.perf_counter() - t0
        for m in (thr, hdr, shs, qsl, shapes, tiles):
            cl.release(m)
    cl.read(out, out_host)
    for m in (geo, sub, pict, uses, rnd, out):
        cl.release(m)
    return out_host, counts, t_kernels


def channel_diff(a, b):
    d = np.zeros(a.shape, dtype=np.int32)
    for shift in (0, 8, 16, 24):
        d = np.maximum(d, np.abs(((a >> shift) & 0xFF).astype(np.int32) - ((b >> shift) & 0xFF).astype(np.int32)))
    return d


def compare(scene, spec, ref, img, taps):
    """ref: the restated oracle's RenderResult (taps on)."""
    res = {}
    if taps:
        cnt_eq = bits_eq = True
        n_cnt_diff = 0
        for (q, b), rc, rb in zip(taps, ref.n_thresholds, ref.shape_bits):
            active = rc >= 0
            n_cnt_diff += int((q[active, 1] != rc[active]).sum())
            for t in np.flatnonzero(active & (q[:, 1] != rc))[:16]:
                res.setdefault("count_differences", []).append([int(t), int(q[t, 1]), int(rc[t])])
            cnt_eq &= bool(np.array_equal(q[active, 1], rc[active]))
            bits_eq &= bool(np.array_equal(b[active], rb[active]))
        res.update(threshold_counts_equal=cnt_eq, threads_with_other_count=n_cnt_diff, shape_bits_equal=bits_eq)
    d = channel_diff(img, ref.image)
    import hashlib
    res.update(sha256=hashlib.sha256(img.astype("<u4").tobytes()).hexdigest(),
               oracle_sha256=hashlib.sha256(ref.image.astype("<u4").tobytes()).hexdigest())
    bad = np.argwhere(d > 0)
    if 0 < len(bad) <= 5000:
        res["differing"] = [[int(y), int(x), int(img[y, x]), int(ref.image[y, x])] for y, x in bad]
    res.update(pixels=int(d.size), pixels_differing=int((d > 0).sum()), max_channel_diff=int(d.max()),
               exact_rate=float((d == 0).mean()), within_1=bool(d.max() <= 1))
    return res


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default="gpurun_out/ocl_reference.json")
    ap.add_argument("--scenes", default="small,s2,s3,s4b,s4")
    ap.add_argument("--variants", default="reference,strict")
    ap.add_argument("--frames", type=int, default=1)
    args = ap.parse_args()
    os.makedirs(os.path.dirname(os.path.abspath(args.out)), exist_ok=True)
    try:
        run(args)
    except Exception as e:  # noqa: BLE001 - everything is logged
        import traceback
        step("FAILED", error=repr(e), trace=traceback.format_exc()[-2000:])
    flush_log(args.out)


def run(args):
    vend = find_vendor_library()
    loader = find_loader()
    step("probe", vendor_libraries=vend, loader=loader, icd_dir=os.path.isdir("/etc/OpenCL/vendors"))
    flush_log(args.out)
    if not loader:
        raise CLError("no ICD loader")
    if vend and not os.path.isdir("/etc/OpenCL/vendors"):
        os.environ["OCL_ICD_FILENAMES"] = vend[0]
    elif not vend:
        os.environ["OCL_ICD_FILENAMES"] = "libnvidia-opencl.so.1"
    cl = CL(loader)
    cl.open()
    LOG["device"] = cl.device_info
    flush_log(args.out)
    blob = json.loads(zlib.decompress(open(os.path.join(os.path.dirname(HERE), "_ref", "ocl_program.bin"), "rb").read()))
    from gudni_b200 import scenes
    from gudni_b200.formats import CANONICAL_SPEC
    from oracle import oracle
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from golden.make_golden import SCENES, digest
    options = {"reference": "-cl-fast-relaxed-math -cl-strict-aliasing",
               "strict": "-cl-fp32-correctly-rounded-divide-sqrt"}
    small = {name: (make, spec or CANONICAL_SPEC) for name, (make, spec) in SCENES.items()}
    small["fuzzy_circles_2000"] = (lambda: scenes.fuzzy_circles(2000, 640, 480, 5, 50, 0x5EED), CANONICAL_SPEC)
    small["random_rectangles_300"] = (lambda: scenes.random_rectangles(300, 640, 480, 5), CANONICAL_SPEC)
    big = {"s2": scenes.s2, "s3": scenes.s3, "s4b": scenes.s4b, "s4": scenes.s4}
    wanted = args.scenes.split(",")
    todo = {}
    if "small" in wanted:
        todo.update(small)
    todo.update({k: (v, CANONICAL_SPEC) for k, v in big.items() if k in wanted})
    prepared = {}
    for name, (make, spec) in todo.items():
        sc = make()
        ref = oracle.render(sc, spec, taps=True)
        prepared[name] = (sc, ref, spec)
        step("oracle ready", scene=name, thresholds=int(ref.total_thresholds), jobs=len(ref.jobs))
    LOG["results"] = {}
    LOG["hashes"] = {}
    programs = {}

    def program(variant, max_t):
        key = (variant, max_t)
        if key not in programs:
            src = blob[str(max_t)][variant]
            try:
                programs[key] = cl.build(src, options[variant])
            except CLError:
                # C99 inline semantics: a plain `inline` function that is not inlined has no definition
                # to link against; one of the padding lines can carry the usual cure
                step("retrying with '#define inline static inline' on a padding line", variant=variant)
                programs[key] = cl.build(src.replace("// Padding line ", "#define inline static inline", 1), options[variant])
                LOG.setdefault("static_inline_needed", []).append(variant)
            flush_log(args.out)
        return programs[key]

    class _Taps:
        pass

    for variant in args.variants.split(","):
        for name, (sc, ref, spec) in prepared.items():
            if variant != "reference" and name in ("s2", "s3"):
                continue
            try:
                kernels = program(variant, spec.max_thresholds)
            except CLError as e:
                step("variant skipped", variant=variant, error=repr(e))
                break
            img, taps, _ = render_opencl(cl, kernels, sc, ref.jobs, spec, taps=True)
            res = compare(sc, spec, ref, img, taps)
            # the same digest tests/golden/make_golden.py takes (inactive threads read back as -1)
            t = _Taps()
            t.image = img
            t.total_thresholds = sum(int(q[rc >= 0, 1].sum()) for (q, b), rc in zip(taps, ref.n_thresholds))
            t.n_thresholds = [np.where(rc >= 0, q[:, 1], -1) for (q, b), rc in zip(taps, ref.n_thresholds)]
            t.shape_bits = [np.where(rc >= 0, b, -1) for (q, b), rc in zip(taps, ref.n_thresholds)]
            LOG["hashes"].setdefault(variant, {})[name] = digest(t)
            res["digest_equals_oracle"] = digest(t) == digest(ref)
            if name in big:                       # timing: whole frames, kernels only and with the scratch churn
                times, ktimes = [], []
                for _ in range(args.frames):
                    t0 = time.perf_counter()
                    _, _, tk = render_opencl(cl, kernels, sc, ref.jobs, spec, taps=False)
                    times.append(time.perf_counter() - t0)
                    ktimes.append(tk)
                res.update(frame_seconds=min(times), kernel_seconds=min(ktimes), frames_timed=args.frames)
            LOG["results"].setdefault(variant, {})[name] = res
            step("compared", variant=variant, scene=name, **{k: v for k, v in res.items() if k != "differing"})
            flush_log(args.out)


if __name__ == "__main__":
    T0 = time.time()
    main()
