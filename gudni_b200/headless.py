"""Headless output target: the reference's only `handleOutput`s present the DrawTarget through
SDL (Application.hs:84-101), so its frames cannot be looked at on a box without a display.  Here the BGRA words the
rasterizer writes (B | G<<8 | R<<16 | 0xFF<<24, Kernels.cl:842-844) go to a binary PPM file; the Haskell twin is
haskell/Graphics/Gudni/CUDA/Headless.hs."""
import numpy as np


def bgra_to_rgb_bytes(image):
    """(H, W) uint32 BGRA words -> (H, W, 3) uint8 RGB."""
    image = np.asarray(image, dtype=np.uint32)
    return np.stack([(image >> 16) & 0xFF, (image >> 8) & 0xFF, image & 0xFF], axis=-1).astype(np.uint8)


def write_ppm(path, image):
    """Binary P6, top row first."""
    rgb = bgra_to_rgb_bytes(image)
    h, w = rgb.shape[:2]
    with open(path, "wb") as f:
        f.write(f"P6\n{w} {h}\n255\n".encode("ascii"))
        f.write(rgb.tobytes())


def read_ppm(path):
    """The inverse of write_ppm: (H, W) uint32 BGRA words with alpha 0xFF."""
    with open(path, "rb") as f:
        data = f.read()
    fields, pos = [], 0
    while len(fields) < 4:          # magic, width, height, maxval: whitespace separated
        while data[pos:pos + 1].isspace():
            pos += 1
        end = pos
        while not data[end:end + 1].isspace():
            end += 1
        fields.append(data[pos:end])
        pos = end
    assert fields[0] == b"P6" and fields[3] == b"255"
    w, h = int(fields[1]), int(fields[2])
    rgb = np.frombuffer(data, np.uint8, count=w * h * 3, offset=pos + 1).reshape(h, w, 3).astype(np.uint32)
    return (rgb[..., 0] << 16) | (rgb[..., 1] << 8) | rgb[..., 2] | np.uint32(0xFF000000)


def handle_output_ppm(prefix, frame, image):
    """`Model.handleOutput` for an application without a window: one numbered file per frame."""
    path = f"{prefix}-{frame:04d}.ppm"
    write_ppm(path, image)
    return path
