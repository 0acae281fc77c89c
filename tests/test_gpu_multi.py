"""Several devices in one process through the C ABI (gudni_b200_multi_*): the assembled canvas — in the caller's
host bitmap and on the presenting device — must be the single-device frame bit for bit, whatever the strips.
On a one-GPU box the devices are several contexts on the same GPU (the ABI allows naming a device twice); with more
GPUs visible they are distinct."""
import ctypes

import numpy as np
import pytest

from gudni_b200 import scenes
from gudni_b200.multi import MultiRasterizer

pytestmark = pytest.mark.gpu


def device_list(n):
    import torch
    have = torch.cuda.device_count()
    return [d % have for d in range(n)]


def read_device(ptr, device, shape):
    rt = ctypes.CDLL("libcudart.so.12")
    out = np.empty(shape, np.uint32)
    assert rt.cudaSetDevice(device) == 0
    assert rt.cudaDeviceSynchronize() == 0
    assert rt.cudaMemcpy(ctypes.c_void_p(out.ctypes.data), ctypes.c_void_p(ptr), ctypes.c_size_t(out.nbytes), 2) == 0
    return out


@pytest.mark.parametrize("n", [1, 2, 3])
def test_multi_frame_equals_single_device_frame(rasterizer, n):
    scene = scenes.fuzzy_circles(6000, 1100, 1500, 5, 60, 0x3171 + n)
    full, st = rasterizer.raster_scene(0, scene)
    m = MultiRasterizer(device_list(n))
    try:
        for frame in range(3):           # frame 0: strips from the boxes; 1, 2: re-cut from the measured times
            img, ms = m.frame(frame, scene)
            assert np.array_equal(img, full), f"{n} devices, frame {frame}: canvas differs from the single-device frame"
            assert ms.n_devices == n and ms.rows[0][0] == 0 and ms.rows[-1][1] == scene.height
            assert all(a[1] == b[0] for a, b in zip(ms.rows, ms.rows[1:]))
            assert ms.total["n_thresholds"] == st.n_thresholds
        m.set_presenting(0)
        img, ms = m.frame(3, scene)
        ptr, dev = m.canvas()
        assert np.array_equal(img, full)
        assert np.array_equal(read_device(ptr, dev, full.shape), full), "device-side canvas differs"
        m.set_presenting(n - 1)          # the presenter need not be the first device
        _, ms = m.frame(4, scene, want_image=False)
        ptr, dev = m.canvas()
        assert np.array_equal(read_device(ptr, dev, full.shape), full)
    finally:
        m.close()


def test_multi_frame_input_cache_and_pictures(rasterizer):
    scene = scenes.picture_scene(640, 700, flowers_size=(700, 375))
    full, _ = rasterizer.raster_scene(0, scene)
    m = MultiRasterizer(device_list(2))
    try:
        gens = [3, 3, 3, 3, 3]
        a, s1 = m.frame(0, scene, generations=gens)
        b, s2 = m.frame(1, scene, generations=gens)
        assert np.array_equal(a, full) and np.array_equal(b, full)
        assert s2.total["ms_upload"] <= s1.total["ms_upload"]
    finally:
        m.close()


def test_multi_frame_reports_a_refused_frame(rasterizer):
    from gudni_b200.raster import GudniError
    scene = scenes.medium_square(size=600)
    scene.geometry = scene.geometry.copy()
    scene.geometry.view(np.float32)[2 * 2] = np.inf
    m = MultiRasterizer(device_list(2))
    try:
        with pytest.raises(GudniError) as e:
            m.frame(0, scene)
        assert "infinity" in str(e.value)
        good = scenes.medium_square(size=600)
        img, _ = m.frame(1, good)
        ref, _ = rasterizer.raster_scene(0, good)
        assert np.array_equal(img, ref)
    finally:
        m.close()
