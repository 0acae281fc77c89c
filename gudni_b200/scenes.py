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


# ---- textured / glyph configs -----------------------------------------------

def _circle_outline(transforms):
    """Unit circle (Layout/Draw.hs:153-154) through a transformer stack given outermost first,
    returned as curve pairs; f32 arithmetic like the harness' C++ (apply innermost first)."""
    from .scene import unit_circle_pairs
    pts = unit_circle_pairs().reshape(-1, 2).astype(np.float32)
    for t in reversed(list(transforms)):
        if t[0] == "translate":
            pts = pts + np.array([t[1], t[2]], np.float32)
        elif t[0] == "scale":
            pts = pts * np.float32(t[1])
    return pts.reshape(-1, 4).astype(np.float32)


def synthetic_picture(width, height, seed):
    """Stand-in for image/hero-yellow-flowers.jpg (1400x750): seeded smooth RGBA8 field with opaque
    and translucent regions.  (Reference assets are not copied into this repo.)"""
    rng = np.random.default_rng(seed)
    y, x = np.mgrid[0:height, 0:width].astype(np.float32)
    img = np.zeros((height, width, 4), np.uint8)
    for c in range(3):
        fx, fy, ph = rng.uniform(0.005, 0.05), rng.uniform(0.005, 0.05), rng.uniform(0, 6.28)
        img[..., c] = (127.5 + 127.5 * np.sin(fx * x + fy * y + ph)).astype(np.uint8)
    img[..., 3] = np.where((x // 37 + y // 29) % 5 == 0, 128, 255).astype(np.uint8)
    return img


def hsl_gradient_picture(w=200, h=200):
    """testPict's PictureFunction (benchmarks/GudniTests.hs:272-276): hsl 0 (x/w) (y/h), RGBA8."""
    y, x = np.mgrid[0:h, 0:w].astype(np.float32)
    s, l = x / np.float32(w), y / np.float32(h)
    q = np.where(l < 0.5, l * (1 + s), l + s - l * s)
    p = 2 * l - q

    def comp(t):
        t = t - np.floor(t)
        return np.where(t < 1 / 6, p + (q - p) * 6 * t, np.where(t < 0.5, q, np.where(t < 2 / 3, p + (q - p) * 6 * (2 / 3 - t), p)))

    rgb = np.stack([comp(np.full_like(l, 1 / 3)), comp(np.zeros_like(l)), comp(np.full_like(l, -1 / 3))], axis=-1)
    out = np.empty((h, w, 4), np.uint8)
    out[..., :3] = np.clip(np.floor(rgb * 255.0), 0, 255).astype(np.uint8)   # colorToRGBA8 truncates
    out[..., 3] = 255
    return out


def picture_scene(width=640, height=480, scale=1.0, background=(0.2, 0.2, 0.2, 1.0), flowers_size=(1400, 750)):
    """testPict (benchmarks/GudniTests.hs:271-281): a picture on `circle200 - (circle100@(100,100) +
    circle100)`, an HSL gradient PictureFunction on a radius-200 circle, a translucent bar; the
    whole thing optionally scaled (S3 uses x2).  Picture usages take the translation but only the
    scale factor of an enclosing tScale (Figure/Picture.hs:113-116)."""
    b = SceneBuilder(width, height, background, name=f"testPict-{width}x{height}-x{scale}")
    sc = ("scale", float(scale))
    outer = [sc, ("translate", 100.0, 50.0)]
    flowers = b.picture(synthetic_picture(flowers_size[0], flowers_size[1], 0xF10E5))
    s0 = b.picture_substance(flowers, (100.0, 50.0), float(scale))
    b.shape(s0, [_circle_outline(outer + [("translate", 100.0, 100.0), ("scale", 100.0)])], subtract=True, is_picture=True)
    b.shape(s0, [_circle_outline(outer + [("scale", 100.0)])], subtract=True, is_picture=True)
    b.shape(s0, [_circle_outline(outer + [("scale", 200.0)])], is_picture=True)
    gradient = b.picture(hsl_gradient_picture())
    s1 = b.picture_substance(gradient, (100.0, 50.0), float(scale))
    b.shape(s1, [_circle_outline(outer + [("scale", 200.0)])], is_picture=True)
    s2 = b.solid(0.0, 0.0, 1.0, 0.2)
    b.rectangle(s2, 40.0, 2000.0, outer)
    return b.freeze()


def s3(width=3840, height=2160):
    """S3 — plots + textures at 4K: testPict x2 plus a row of solid closed curves standing in for the
    turtle plots of examples/Plot.hs (arcs, rounded boxes, circles)."""
    b = SceneBuilder(width, height, (0.2, 0.2, 0.2, 1.0), name=f"S3-plots-textures-{width}x{height}")
    outer = [("scale", 2.0), ("translate", 100.0, 50.0)]
    flowers = b.picture(synthetic_picture(1400, 750, 0xF10E5))
    s0 = b.picture_substance(flowers, (100.0, 50.0), 2.0)
    b.shape(s0, [_circle_outline(outer + [("translate", 100.0, 100.0), ("scale", 100.0)])], subtract=True, is_picture=True)
    b.shape(s0, [_circle_outline(outer + [("scale", 100.0)])], subtract=True, is_picture=True)
    b.shape(s0, [_circle_outline(outer + [("scale", 200.0)])], is_picture=True)
    gradient = b.picture(hsl_gradient_picture())
    s1 = b.picture_substance(gradient, (100.0, 50.0), 2.0)
    b.shape(s1, [_circle_outline(outer + [("scale", 200.0)])], is_picture=True)
    s2 = b.solid(0.0, 0.0, 1.0, 0.2)
    b.rectangle(s2, 40.0, 2000.0, outer)
    yellow = b.solid(*YELLOW, 1.0)
    for i in range(16):   # 16-wide row of plot-like closed curves, scale 30 -> x4
        x, y = 1000.0 + 170.0 * i, 300.0 + 90.0 * (i % 5)
        if i % 3 == 0:
            b.circle(yellow, [("translate", x, y), ("scale", 60.0)])
        elif i % 3 == 1:
            b.rectangle(yellow, 120.0, 80.0, [("translate", x, y), ("rotate", 0.03 * i)])
        else:
            b.circle(yellow, [("translate", x, y), ("scale", 30.0)], subtract=True)
            b.circle(yellow, [("translate", x, y), ("scale", 70.0)])
    return b.freeze()


def _synthetic_glyph(rng):
    """A glyph-like shape in the unit em box: an outer contour of 10-20 quadratic curves around the
    centre and, half of the time, an inner counter (a hole drawn as a second outline)."""
    outlines = []
    for k, (r0, r1) in enumerate([(0.28, 0.45)] + ([(0.08, 0.16)] if rng.uniform() < 0.5 else [])):
        n = int(rng.integers(5, 11))
        ang = np.sort(rng.uniform(0, 2 * np.pi, 2 * n))
        rad = rng.uniform(r0, r1, 2 * n)
        pts = np.stack([0.5 + rad * np.cos(ang) * 0.8, 0.5 + rad * np.sin(ang)], axis=1).astype(np.float32)
        if k == 1:
            pts = pts[::-1].copy()
        outlines.append(pts.reshape(n, 4).astype(np.float32))   # (on, off) pairs: Outline.hs:74-77 pairPoints
    return outlines


def s2(width=1920, height=1080, lines=30, em=30.0, seed=0x5EED0002):
    """S2 — examples/Paragraph.hs restated: 30 lines of glyph outlines, em = 30 px, advance 0.8 em +
    0.2 em gaps, ONE opaque black substance for every glyph and for a radius-10-em circle to the right
    of the text, light gray background.  No usable font ships with the image's Python stack, so the
    glyphs are seeded synthetic outlines (SURVEY.md §8(d) fallback); 76 glyphs per line ~ 2,280."""
    rng = np.random.default_rng(seed)
    b = SceneBuilder(width, height, (0.83, 0.83, 0.83, 1.0), name=f"S2-paragraph-{width}x{height}")
    black = b.solid(0.0, 0.0, 0.0, 1.0)
    alphabet = [_synthetic_glyph(rng) for _ in range(48)]
    per_line = 76
    for line in range(lines):
        for col in range(per_line):
            if rng.uniform() < 0.15:
                continue   # a space
            g = alphabet[int(rng.integers(0, len(alphabet)))]
            ox, oy = np.float32(10.0 + col * em * 0.8), np.float32(10.0 + line * em * 1.15)
            outs = []
            for o in g:
                p = o.reshape(-1, 2) * np.float32(em) + np.array([ox, oy], np.float32)
                outs.append(p.reshape(-1, 4).astype(np.float32))
            b.shape(black, outs)
    b.circle(black, [("translate", 10.0 + per_line * em * 0.8 + 10 * em, 10 * em + 10.0), ("scale", 10 * em)])
    return b.freeze()


def thin_rectangles(n, width=64, height=None, spacing=1.0, thickness=0.5, skew=0.3, background=(1.0, 1.0, 1.0, 1.0),
                    one_shape=False):
    """maxThresholdTest / maxShapeTest flavour (benchmarks/GudniTests.hs:100-127): a stack of wide, thin,
    slightly skewed translucent rectangles — hundreds of thresholds per pixel column.  Each rectangle is
    a shape with its own substance, so more than MAXSHAPE of them make the tiles split; `one_shape`
    puts all the outlines into a single shape instead (one substance, tiles stay whole, every column of
    a tall tile crosses 2 n thresholds)."""
    height = int(n * spacing + 8) if height is None else height
    b = SceneBuilder(width, height, background, name=f"thinRects-{n}" + ("-one" if one_shape else ""))
    outlines = []
    for i in range(n):
        y = np.float32(2.0 + i * spacing)
        pts = [(-1.0, float(y)), (width + 1.0, float(y + skew)), (width + 1.0, float(y + skew + thickness)), (-1.0, float(y + thickness))]
        if one_shape:
            outlines.append(_straight_outline(pts))
        else:
            s = b.solid(0.9 * ((i * 37) % 11) / 10.0, 0.9 * ((i * 53) % 7) / 6.0, 0.5, 0.4)
            b.shape(s, [_straight_outline(pts)])
    if one_shape:
        b.shape(b.solid(0.1, 0.3, 0.6, 0.6), outlines)
    return b.freeze()


def mixed_bag(n, width, height, seed, name=""):
    """Seeded mix of everything the path distinguishes, for differential tests: free-form closed curves
    with random control points (knobs to split, strands of every length, self-intersections), thin
    slivers, rotated rectangles, circles, add/subtract pairs sharing a substance, opaque and translucent
    solids, and picture substances with random placement and scale."""
    rng = np.random.default_rng(seed)
    b = SceneBuilder(width, height, tuple(float(v) for v in rng.uniform(0, 1, 3)) + (1.0,),
                     name=name or f"mixedBag-{n}-{width}x{height}-{seed}")
    pictures = [b.picture(synthetic_picture(int(rng.integers(8, 90)), int(rng.integers(8, 70)), int(seed) + k))
                for k in range(2)]

    def blob(cx, cy, radius, points):
        ang = np.sort(rng.uniform(0, 2 * np.pi, points))
        rad = radius * rng.uniform(0.35, 1.0, points)
        on = np.stack([cx + rad * np.cos(ang), cy + rad * np.sin(ang)], axis=1)
        mid = 0.5 * (on + np.roll(on, -1, axis=0))
        off = mid + rng.normal(0, 0.45 * radius, (points, 2)) * (rng.uniform(size=(points, 1)) < 0.7)
        return np.concatenate([on, off], axis=1).astype(np.float32)

    for _ in range(n):
        x, y = float(rng.uniform(-0.1 * width, 1.1 * width)), float(rng.uniform(-0.1 * height, 1.1 * height))
        size = float(np.exp(rng.uniform(np.log(0.4), np.log(0.45 * max(width, height)))))
        use_picture = rng.uniform() < 0.15
        if use_picture:
            sub = b.picture_substance(pictures[int(rng.integers(0, 2))], (x - size, y - size),
                                      float(rng.choice([1.0, 1.0, 2.0, 0.5, 3.3])))
        else:
            alpha = 1.0 if rng.uniform() < 0.3 else float(rng.uniform(0.05, 0.95))
            sub = b.solid(*(float(v) for v in rng.uniform(0, 1, 3)), alpha)
        kind = int(rng.integers(0, 6))
        if kind == 0:
            b.shape(sub, [blob(x, y, size, int(rng.integers(3, 24)))], is_picture=use_picture)
        elif kind == 1:      # a hole cut by a second, subtracting shape of the same substance (listed first = on top)
            b.shape(sub, [blob(x, y, 0.5 * size, int(rng.integers(3, 9)))], subtract=True, is_picture=use_picture)
            b.shape(sub, [blob(x, y, size, int(rng.integers(3, 12)))], is_picture=use_picture)
        elif kind == 2:      # two outlines in one shape (even-odd between them)
            b.shape(sub, [blob(x, y, size, int(rng.integers(3, 10))), blob(x, y, 0.6 * size, int(rng.integers(3, 10)))],
                    is_picture=use_picture)
        elif kind == 3 and not use_picture:
            b.rectangle(sub, float(rng.uniform(0.05, 2.0) * size), float(rng.uniform(0.05, 2.0) * size),
                        [("translate", x, y), ("rotate", float(rng.uniform(0, 1)))])
        elif kind == 4 and not use_picture:
            b.circle(sub, [("translate", x, y), ("scale", size)])
        else:                # sliver: a long triangle thinner than a pixel
            ang = float(rng.uniform(0, 2 * np.pi))
            d = np.array([np.cos(ang), np.sin(ang)]) * size * 3
            nrm = np.array([-np.sin(ang), np.cos(ang)]) * float(rng.uniform(0.05, 0.8))
            b.shape(sub, [_straight_outline([(x, y), (x + d[0], y + d[1]), (x + d[0] + nrm[0], y + d[1] + nrm[1])])],
                    is_picture=use_picture)
    return b.freeze()


def huge_boxes(width=300, height=200):
    """Translucent rectangles up to 4e12 pixels across, behind and in front of some circles: their boxes, divided by
    the root tile size, are far outside int32, and the tile tree (Raster/TileTree.hs:120-139) compares floats."""
    b = SceneBuilder(width, height, (0.9, 0.9, 0.9, 1.0), name="hugeBoxes")
    b.rectangle(b.solid(0.1, 0.3, 0.8, 0.4), 4.0e12, 4.0e12, [("translate", -2.0e12, -2.0e12)])
    b.fuzzy_circles(40, width, height, 5, 40, 0xB16)
    b.rectangle(b.solid(0.8, 0.3, 0.1, 0.3), 3.0e12, 50.0, [("translate", -1.0e12, 70.0)])
    return b.freeze()


def far_shapes(n, width, height, seed):
    """Seeded rectangles and circles whose size is log-uniform between 1 and 1e30 pixels, placed so that most of them
    cross or cover the canvas — coordinates far beyond what an int or a pixel grid holds (fuzz fodder for the binning's
    float compares and the kernels' arithmetic at the far end of float32)."""
    rng = np.random.default_rng(seed)
    b = SceneBuilder(width, height, (0.9, 0.9, 0.9, 1.0), name=f"farShapes-{n}-{seed}")
    for _ in range(n):
        col = rng.uniform(0, 1, 3)
        s = b.solid(float(col[0]), float(col[1]), float(col[2]), float(rng.uniform(0.2, 1.0)))
        size = float(10.0 ** rng.uniform(0.0, 30.0))
        x = float(rng.uniform(0, width) - size * rng.uniform(0, 1))
        y = float(rng.uniform(0, height) - size * rng.uniform(0, 1))
        kind = rng.integers(0, 3)
        if kind == 0:
            b.rectangle(s, size, float(size * rng.uniform(0.01, 1.0)), [("translate", x, y)])
        elif kind == 1:
            b.rectangle(s, size, float(size * rng.uniform(0.01, 1.0)), [("translate", x, y), ("rotate", float(rng.uniform(0, 1)))])
        else:
            b.circle(s, [("translate", float(x + size / 2), float(y + size / 2)), ("scale", size / 2)])
    return b.freeze()
