"""Named scenes — HARNESS, not product.

Restates the BASELINE.json configs (S1, S4, S4b, S5, S5b) and the hand-checkable
scenes of the reference's scene catalogue (benchmarks/GudniTests.hs) as inputs for tests and
bench.py.  Shapes are added top-most first, which is the order `traverseShapeTree` emits them
(Raster/TraverseShapeTree.hs:59-62; `cSubtract a b` emits b, flagged subtract, before a:
Layout/Draw.hs:62-64).
"""
import numpy as np

from .scene import SceneBuilder

RED = (1.0, 0.0, 0.0)
GREEN = (0.0, 1.0, 0.0)   # pure channels keep the known answers exact
BLUE = (0.0, 0.0, 1.0)
YELLOW = (1.0, 1.0, 0.0)
ORANGE = (1.0, 0.5, 0.0)
WHITE = (1.0, 1.0, 1.0)
BLUISH_BACKGROUND = (0.35, 0.45, 0.95, 1.0)   # stands in for `light . greenish $ blue` (Square.hs:49)


def square(size=100, theta_turn=0.4, scale=50.0):
    """S1 — examples/Square.hs:44-53: tTranslate (100,100) . tScale s . tRotate θ . solid yellow $ unitSquare."""
    b = SceneBuilder(size, size, BLUISH_BACKGROUND, name=f"S1-square-{size}-{theta_turn}")
    y = b.solid(*YELLOW, 1.0)
    b.rectangle(y, 1.0, 1.0, [("translate", 100.0, 100.0), ("scale", scale), ("rotate", theta_turn)])
    return b.freeze()


def fuzzy_circles(n, width, height, min_rad, max_rad, seed, background=(1.0, 1.0, 1.0, 1.0), name=""):
    """fuzzyCircles (benchmarks/GudniTests.hs:143-149) over the canvas."""
    b = SceneBuilder(width, height, background, name=name)
    b.fuzzy_circles(n, float(width), float(height), float(min_rad), float(max_rad), seed)
    return b.freeze()


def s4(n=100_000, width=3840, height=2160):
    return fuzzy_circles(n, width, height, 5, 50, 0x5EED0004, name=f"S4-{n}-circles-{width}x{height}")


def s4b():
    return fuzzy_circles(6250, 3840, 2160, 5, 50, 0x5EED004B, name="S4b-6250-circles-3840x2160")


def s5(n=62_500, side=16384):
    return fuzzy_circles(n, side, side, 20, 200, 0x5EED0005, name=f"S5-{n}-circles-{side}x{side}")


def s5b(n=1_000_000, side=16384):
    return fuzzy_circles(n, side, side, 5, 10, 0x5EED005B, name=f"S5b-{n}-circles-{side}x{side}")


# ---- hand-checkable scenes (benchmarks/GudniTests.hs) ------------------------------------------------

def tiny_square(size=16, background=(0.0, 0.0, 1.0, 1.0)):
    """tinySquare :360-365 — 2x2 red square at (0.1, 0.1)."""
    b = SceneBuilder(size, size, background, name="tinySquare")
    r = b.solid(*RED, 1.0)
    b.rectangle(r, 2, 2, [("translate", 0.1, 0.1)])
    return b.freeze()


def medium_square(size=16, background=(0.0, 0.0, 1.0, 1.0)):
    """mediumSquare :368-373."""
    b = SceneBuilder(size, size, background, name="mediumSquare")
    r = b.solid(*RED, 1.0)
    b.rectangle(r, 10, 10, [("translate", 0.1, 0.1)])
    return b.freeze()


def full_rectangle(width=64, height=48, background=(0.0, 0.0, 1.0, 1.0)):
    """fullRectangle :375-380 — a rectangle larger than the canvas."""
    b = SceneBuilder(width, height, background, name="fullRectangle")
    r = b.solid(*RED, 1.0)
    b.rectangle(r, 2880, 1800, [("translate", 0.0, 0.0)])
    return b.freeze()


def stack_of_squares(size=16, background=(0.0, 0.0, 1.0, 1.0)):
    """stackOfSquares :284-291."""
    b = SceneBuilder(size, size, background, name="stackOfSquares")
    r = b.solid(*RED, 1.0)
    b.rectangle(r, 4, 4, [("translate", 0, 0)])
    g = b.solid(*GREEN, 1.0)
    b.rectangle(g, 4, 4, [("translate", 0, 4)])
    return b.freeze()


def open_square(size=16, alpha=0.5, background=(0.0, 0.0, 1.0, 1.0)):
    """openSquare :294-298 — 5x5 minus 3x3 at (1,1), one translucent substance."""
    b = SceneBuilder(size, size, background, name="openSquare")
    o = b.solid(*ORANGE, alpha)
    b.rectangle(o, 3, 3, [("translate", 1, 1)], subtract=True)
    b.rectangle(o, 5, 5, [])
    return b.freeze()


def concentric_squares2(size=16, background=(0.0, 0.0, 0.0, 1.0)):
    """concentricSquares2 :320-324 — abutting edges."""
    b = SceneBuilder(size, size, background, name="concentricSquares2")
    r = b.solid(*RED, 1.0)
    b.rectangle(r, 3, 3, [("translate", 0, 0), ("translate", 1, 1)], subtract=True)
    b.rectangle(r, 5, 5, [("translate", 0, 0)])
    bl = b.solid(*BLUE, 1.0)
    b.rectangle(bl, 1, 1, [("translate", 1, 1), ("translate", 1, 1)], subtract=True)
    b.rectangle(bl, 3, 3, [("translate", 1, 1)])
    return b.freeze()


def concentric_squares3(size=16, background=(0.0, 0.0, 0.0, 1.0)):
    """concentricSquares3 :327-332."""
    b = SceneBuilder(size, size, background, name="concentricSquares3")
    r = b.solid(*RED, 1.0)
    b.rectangle(r, 6, 6, [("translate", 0, 0), ("translate", 2, 2)], subtract=True)
    b.rectangle(r, 10, 10, [("translate", 0, 0)])
    g = b.solid(*GREEN, 1.0)
    b.rectangle(g, 2, 2, [("translate", 2, 2), ("translate", 2, 2)], subtract=True)
    b.rectangle(g, 6, 6, [("translate", 2, 2)])
    bl = b.solid(*BLUE, 1.0)
    b.rectangle(bl, 2, 2, [("translate", 4, 4)])
    return b.freeze()


def six_point_rectangle(size=16, background=(0.0, 0.0, 1.0, 1.0)):
    """sixPointRectangle :352-357 — colinear points on straight edges."""
    b = SceneBuilder(size, size, background, name="sixPointRectangle")
    r = b.solid(*RED, 1.0)
    pts = [(0, 0), (1, 0), (2, 0), (2, 1), (1, 1), (0, 1)]
    b.shape(r, [_straight_outline(pts)])
    return b.freeze()


def hour_glass(size=16, scale=8.0, background=(0.0, 0.0, 1.0, 1.0)):
    """hourGlass :258-268 — self-intersecting outline."""
    b = SceneBuilder(size, size, background, name="hourGlass")
    r = b.solid(*RED, 1.0)
    pts = [(0, 0), (scale, scale), (scale, 0), (0, scale)]
    b.shape(r, [_straight_outline(pts)])
    return b.freeze()


def translucent_stack(size=32, layers=5, background=(1.0, 1.0, 1.0, 1.0)):
    """Axis-aligned translucent rectangles on integer coordinates: exact `composite` known answer."""
    b = SceneBuilder(size, size, background, name="translucentStack")
    colors = [RED, GREEN, BLUE, YELLOW, ORANGE]
    for i in range(layers):
        s = b.solid(*colors[i % len(colors)], 0.5)
        b.rectangle(s, size - 2 * i - 2, size - 2 * i - 2, [("translate", i + 1, i + 1)])
    return b.freeze()


def _straight_outline(points):
    """segmentsToCurvePairs for straight segments: control = midpoint to the next anchor
    (Figure/Outline.hs:95-101)."""
    pts = np.asarray(points, dtype=np.float32)
    nxt = np.roll(pts, -1, axis=0)
    mid = np.float32(0.5) * pts + np.float32(0.5) * nxt
    return np.concatenate([pts, mid], axis=1).astype(np.float32)


def random_rectangles(n, width, height, seed, max_size=40.0, alpha=(0.2, 1.0)):
    """Seeded mix of rotated rectangles and circles with add / subtract pairs — parity fodder."""
    rng = np.random.default_rng(seed)
    b = SceneBuilder(width, height, (0.9, 0.9, 0.9, 1.0), name=f"randomRects-{n}-{seed}")
    for _ in range(n):
        col = rng.uniform(0, 1, 3)
        a = 1.0 if rng.uniform() < 0.3 else rng.uniform(*alpha)
        s = b.solid(float(col[0]), float(col[1]), float(col[2]), float(a))
        x, y = rng.uniform(-10, width), rng.uniform(-10, height)
        w, h = rng.uniform(0.3, max_size, 2)
        rot = rng.uniform(0, 1)
        kind = rng.integers(0, 4)
        if kind == 0:
            b.rectangle(s, float(w), float(h), [("translate", float(x), float(y))])
        elif kind == 1:
            b.rectangle(s, float(w), float(h), [("translate", float(x), float(y)), ("rotate", float(rot))])
        elif kind == 2:
            b.circle(s, [("translate", float(x), float(y)), ("scale", float(w) / 2)])
        else:
            b.circle(s, [("translate", float(x), float(y)), ("scale", float(w) / 4)], subtract=True)
            b.circle(s, [("translate", float(x), float(y)), ("scale", float(w) / 2)])
    return b.freeze()
