// scene.cpp — harness-side scene serialisation: what Raster/Serialize.hs does in the reference
// (substance ids, shape tags, canvas culling, geometry pile, picture heap), restated so tests and
// bench.py can produce the byte buffers that cross the drop-in boundary.
// Built into libgudni_host.so and driven from Python over ctypes.  NOT the rasterizer product and
// NOT the oracle: it only produces inputs.  Paths relative to /root/reference/src/Graphics/Gudni/.
#include <cstdint>
#include <cstring>
#include <string>
#include <vector>

#include "../../../include/gudni_b200.h"
#include "figure.hpp"
#include "strand.hpp"

using namespace gudni;

namespace {

// SplitMix64 — the harness PRNG (Haskell's StdGen streams are not reproducible).
struct SplitMix64 {
    uint64_t s;
    explicit SplitMix64(uint64_t seed) : s(seed) {}
    uint64_t next() {
        uint64_t z = (s += 0x9E3779B97F4A7C15ull);
        z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
        z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
        return z ^ (z >> 31);
    }
    float uniform(float lo, float hi) {  // 24 random mantissa bits, [lo, hi)
        float u = (float)(next() >> 40) * (1.0f / 16777216.0f);
        return lo + (hi - lo) * u;
    }
};

// Data.Colour.RGBSpace.HSL.hsl, as used by hslColor (Figure/Color.hs:103-104).
void hslToRgb(float h, float s, float l, float rgb[3]) {
    auto mod1 = [](float x) { return x - std::floor(x); };
    float hk = h / 360.f;
    float q = (l < 0.5f) ? l * (1.f + s) : l + s - l * s;
    float p = 2.f * l - q;
    float t[3] = {mod1(hk + 1.f / 3.f), mod1(hk), mod1(hk - 1.f / 3.f)};
    for (int i = 0; i < 3; i++) {
        float c;
        if (t[i] < 1.f / 6.f) c = p + ((q - p) * 6.f * t[i]);
        else if (t[i] < 0.5f) c = q;
        else if (t[i] < 2.f / 3.f) c = p + ((q - p) * 6.f * (2.f / 3.f - t[i]));
        else c = p;
        rgb[i] = c;
    }
}

struct Scene {
    int width = 0, height = 0;
    float background[4] = {0, 0, 0, 1};
    std::vector<uint8_t> geometry;            // geoGeometryPile (Raster/Serialize.hs:111)
    std::vector<gudni_shape_entry> entries;   // shapes in scene order (first = top-most)
    std::vector<float> substances;            // suSubstancePile, 4 floats each
    std::vector<uint8_t> pictureBytes;        // makePictData (Figure/Picture.hs:159)
    std::vector<gudni_picture_use> pictureUses;
    std::vector<int> pictureOffsets, pictureW, pictureH;
    int64_t culled = 0, curves = 0;

    // The same scene BEFORE serialisation, as level 3 of the ABI takes it (gudni_b200_raster_outlines):
    // every shape with its untransformed outlines and its transformer stack, culled ones included.
    std::vector<gudni_outline_shape> rawShapes;
    std::vector<gudni_outline> rawOutlines;
    std::vector<gudni_curve_pair> rawPairs;
    std::vector<gudni_transform> rawTransforms;
    int rawUnitCircle = -1;   // outline index shared by every circle

    int rawAddOutline(const Outline& o) {
        gudni_outline r{(uint32_t)rawPairs.size(), (uint32_t)o.size()};
        for (const CurvePair& p : o) rawPairs.push_back({p.on.x, p.on.y, p.off.x, p.off.y});
        rawOutlines.push_back(r);
        return (int)rawOutlines.size() - 1;
    }
    static uint64_t makeTag(int substance, bool isPicture, bool subtract) {
        return (isPicture ? GUDNI_TAG_SUBSTANCE_PICTURE : GUDNI_TAG_SUBSTANCE_SOLID) |
               (subtract ? GUDNI_TAG_COMPOUND_SUBTRACT : GUDNI_TAG_COMPOUND_ADD) |
               ((uint64_t)substance & GUDNI_TAG_SUBSTANCEID_MASK);
    }
    void rawAddShape(int substance, bool isPicture, bool subtract, int firstOutline, int nOutlines,
                     const std::vector<Transform>& stack) {
        gudni_outline_shape r{};
        r.tag = makeTag(substance, isPicture, subtract);
        r.first_outline = (uint32_t)firstOutline;
        r.n_outlines = (uint32_t)nOutlines;
        r.first_transform = (uint32_t)rawTransforms.size();
        r.n_transforms = (uint32_t)stack.size();
        for (const Transform& t : stack) {
            gudni_transform g{};
            if (t.kind == Transform::Translate) { g.kind = GUDNI_TRANSFORM_TRANSLATE; g.a = t.delta.x; g.b = t.delta.y; }
            else if (t.kind == Transform::Scale) { g.kind = GUDNI_TRANSFORM_SCALE; g.a = t.factor; }
            else { g.kind = GUDNI_TRANSFORM_ROTATE; g.a = std::cos(t.factor); g.b = std::sin(t.factor); }   // as rotatePoint
            rawTransforms.push_back(g);
        }
        rawShapes.push_back(r);
    }

    // onSubstance (Raster/Serialize.hs:216-262), Solid branch.
    int addSolid(float r, float g, float b, float a) {
        substances.insert(substances.end(), {r, g, b, a});
        return (int)(substances.size() / 4) - 1;
    }
    // makePictData appends a picture the first time a usage refers to it.
    int addPicture(const uint8_t* rgba, int w, int h) {
        pictureOffsets.push_back((int)pictureBytes.size());
        pictureW.push_back(w);
        pictureH.push_back(h);
        pictureBytes.insert(pictureBytes.end(), rgba, rgba + (size_t)w * h * 4);
        return (int)pictureOffsets.size() - 1;
    }
    // onSubstance, Texture branch: the usage carries translate/scale; the substance record holds
    // the usage index in its first word (Raster/ShapeInfo.hs:143-148).
    int addPictureSubstance(int picture, float tx, float ty, float scale) {
        gudni_picture_use u{};
        u.translate_x = tx;
        u.translate_y = ty;
        u.width = pictureW[picture];
        u.height = pictureH[picture];
        u.mem_offset = (uint32_t)pictureOffsets[picture];
        u.scale = scale;
        pictureUses.push_back(u);
        uint32_t idx = (uint32_t)pictureUses.size() - 1;
        float rec[4] = {0, 0, 0, 0};
        std::memcpy(&rec[0], &idx, 4);
        substances.insert(substances.end(), rec, rec + 4);
        return (int)(substances.size() / 4) - 1;
    }

    // onShape (Raster/Serialize.hs:148-177) for outlines that are already transformed.
    void addShape(int substance, bool isPicture, bool subtract, const std::vector<Outline>& outlines) {
        // boxOf: min/max over on- and off-curve points (Figure/Outline.hs:123-133)
        float l = INFINITY, t = INFINITY, r = -INFINITY, b = -INFINITY;
        for (const Outline& o : outlines)
            for (const CurvePair& p : o) {
                l = std::fmin(l, std::fmin(p.on.x, p.off.x));
                r = std::fmax(r, std::fmax(p.on.x, p.off.x));
                t = std::fmin(t, std::fmin(p.on.y, p.off.y));
                b = std::fmax(b, std::fmax(p.on.y, p.off.y));
            }
        // excludeBox (:97-104)
        if (l >= (float)width || t >= (float)height || r <= 0.f || b <= 0.f) {
            culled++;
            return;
        }
        // enclose (Raster/Enclosure.hs:62-73)
        std::vector<StrandPoints> strands;
        for (const Outline& o : outlines) outlineToStrands(o, strands);
        // appendGeoRef (:76-85): start measured in 16-byte units
        gudni_shape_entry e{};
        e.tag = makeTag(substance, isPicture, subtract);
        e.geo_start = (uint32_t)(geometry.size() / 16);
        e.num_strands = (uint32_t)strands.size();
        e.left = l; e.top = t; e.right = r; e.bottom = b;
        for (const StrandPoints& sp : strands) {
            appendStrandBytes(sp, geometry);
            curves += (int64_t)(sp.size() - 1) / 2;
        }
        entries.push_back(e);
    }
};

}  // namespace

extern "C" {

void* gs_scene_new(int width, int height, const float* background_rgba) {
    Scene* s = new Scene();
    s->width = width;
    s->height = height;
    if (background_rgba) std::memcpy(s->background, background_rgba, 16);
    return s;
}
void gs_scene_free(void* h) { delete static_cast<Scene*>(h); }

int gs_add_solid(void* h, float r, float g, float b, float a) { return static_cast<Scene*>(h)->addSolid(r, g, b, a); }
int gs_add_picture(void* h, const uint8_t* rgba, int w, int hgt) { return static_cast<Scene*>(h)->addPicture(rgba, w, hgt); }
int gs_add_picture_substance(void* h, int picture, float tx, float ty, float scale) {
    return static_cast<Scene*>(h)->addPictureSubstance(picture, tx, ty, scale);
}

// pairs = 4 floats per curve pair (on.x, on.y, off.x, off.y); outline_sizes = pairs per outline.
void gs_add_shape(void* h, int substance, int is_picture, int subtract, const float* pairs,
                  const int* outline_sizes, int n_outlines) {
    std::vector<Outline> outlines(n_outlines);
    const float* p = pairs;
    for (int i = 0; i < n_outlines; i++) {
        outlines[i].resize(outline_sizes[i]);
        for (int k = 0; k < outline_sizes[i]; k++, p += 4) outlines[i][k] = {{p[0], p[1]}, {p[2], p[3]}};
    }
    Scene* s = static_cast<Scene*>(h);
    const int first = (int)s->rawOutlines.size();
    for (const Outline& o : outlines) s->rawAddOutline(o);
    s->rawAddShape(substance, is_picture != 0, subtract != 0, first, n_outlines, {});   // already transformed
    s->addShape(substance, is_picture != 0, subtract != 0, outlines);
}

// Layout/Draw.hs rectangle / circle run through a transformer stack given outermost first as
// (kind, a, b) triples: kind 0 = translate(a,b), 1 = scale(a), 2 = rotate(a turns).
static std::vector<Transform> parseStack(const float* stack, int n) {
    std::vector<Transform> t;
    for (int i = 0; i < n; i++) {
        int kind = (int)stack[3 * i];
        if (kind == 0) t.push_back(Transform::translate(stack[3 * i + 1], stack[3 * i + 2]));
        else if (kind == 1) t.push_back(Transform::scale(stack[3 * i + 1]));
        else t.push_back(Transform::rotateTurn(stack[3 * i + 1]));
    }
    return t;
}
void gs_add_rectangle(void* h, int substance, int subtract, float w, float hgt, const float* stack, int n_stack) {
    Scene* s = static_cast<Scene*>(h);
    const std::vector<Transform> st = parseStack(stack, n_stack);
    s->rawAddShape(substance, false, subtract != 0, s->rawAddOutline(rectangleOutline(w, hgt)), 1, st);
    Outline o = transformOutline(st, rectangleOutline(w, hgt));
    s->addShape(substance, false, subtract != 0, {o});
}
void gs_add_circle(void* h, int substance, int is_picture, int subtract, const float* stack, int n_stack) {
    Scene* s = static_cast<Scene*>(h);
    const std::vector<Transform> st = parseStack(stack, n_stack);
    if (s->rawUnitCircle < 0) s->rawUnitCircle = s->rawAddOutline(circleOutline());
    s->rawAddShape(substance, is_picture != 0, subtract != 0, s->rawUnitCircle, 1, st);
    Outline o = transformOutline(st, circleOutline());
    s->addShape(substance, is_picture != 0, subtract != 0, {o});
}
// Number of curve pairs of the unit circle outline followed by the pairs themselves (for Python).
int gs_unit_circle(float* pairs, int capacity) {
    Outline o = circleOutline();
    if ((int)o.size() <= capacity)
        for (size_t i = 0; i < o.size(); i++) {
            pairs[4 * i] = o[i].on.x; pairs[4 * i + 1] = o[i].on.y;
            pairs[4 * i + 2] = o[i].off.x; pairs[4 * i + 3] = o[i].off.y;
        }
    return (int)o.size();
}
int gs_arc(float turns, float* pairs, int capacity) {
    Outline o = arcOutline(turns * 6.283185307179586f);
    if ((int)o.size() <= capacity)
        for (size_t i = 0; i < o.size(); i++) {
            pairs[4 * i] = o[i].on.x; pairs[4 * i + 1] = o[i].on.y;
            pairs[4 * i + 2] = o[i].off.x; pairs[4 * i + 3] = o[i].off.y;
        }
    return (int)o.size();
}

// fuzzyCircles / millionFuzzyCircles (benchmarks/GudniTests.hs:143-166; Util/Fuzzy.hs:130-137,
// colour distribution :38-44): n random translucent circles, each its own solid substance,
// `tTranslate point . tScale radius $ circle`; first generated = top-most.
void gs_add_fuzzy_circles(void* h, int n, float range_w, float range_h, float min_rad, float max_rad,
                          uint64_t seed) {
    Scene* s = static_cast<Scene*>(h);
    SplitMix64 rng(seed);
    const Outline unit = circleOutline();
    for (int i = 0; i < n; i++) {
        float hue = rng.uniform(0.f, 360.f), sat = rng.uniform(0.3f, 1.f);
        float light = rng.uniform(0.4f, 0.9f), alpha = rng.uniform(0.2f, 0.5f);
        float radius = rng.uniform(min_rad, max_rad);
        float px = rng.uniform(0.f, range_w), py = rng.uniform(0.f, range_h);
        float rgb[3];
        hslToRgb(hue, sat, light, rgb);
        int sub = s->addSolid(rgb[0], rgb[1], rgb[2], alpha);
        const std::vector<Transform> st{Transform::translate(px, py), Transform::scale(radius)};
        if (s->rawUnitCircle < 0) s->rawUnitCircle = s->rawAddOutline(unit);
        s->rawAddShape(sub, false, false, s->rawUnitCircle, 1, st);
        Outline o = transformOutline(st, unit);
        s->addShape(sub, false, false, {o});
    }
}

// --- getters -------------------------------------------------------------------------------------
const void* gs_geometry(void* h, size_t* bytes) {
    Scene* s = static_cast<Scene*>(h);
    *bytes = s->geometry.size();
    return s->geometry.data();
}
const void* gs_entries(void* h, int* n) {
    Scene* s = static_cast<Scene*>(h);
    *n = (int)s->entries.size();
    return s->entries.data();
}
const float* gs_substances(void* h, int* n) {
    Scene* s = static_cast<Scene*>(h);
    *n = (int)(s->substances.size() / 4);
    return s->substances.data();
}
const uint8_t* gs_picture_bytes(void* h, size_t* bytes) {
    Scene* s = static_cast<Scene*>(h);
    *bytes = s->pictureBytes.size();
    return s->pictureBytes.data();
}
const void* gs_picture_uses(void* h, int* n) {
    Scene* s = static_cast<Scene*>(h);
    *n = (int)s->pictureUses.size();
    return s->pictureUses.data();
}
// the scene before serialisation (level 3 inputs)
const void* gs_raw_shapes(void* h, int* n) { Scene* s = static_cast<Scene*>(h); *n = (int)s->rawShapes.size(); return s->rawShapes.data(); }
const void* gs_raw_outlines(void* h, int* n) { Scene* s = static_cast<Scene*>(h); *n = (int)s->rawOutlines.size(); return s->rawOutlines.data(); }
const void* gs_raw_pairs(void* h, int* n) { Scene* s = static_cast<Scene*>(h); *n = (int)s->rawPairs.size(); return s->rawPairs.data(); }
const void* gs_raw_transforms(void* h, int* n) { Scene* s = static_cast<Scene*>(h); *n = (int)s->rawTransforms.size(); return s->rawTransforms.data(); }
void gs_info(void* h, int* width, int* height, float* background, int64_t* culled, int64_t* curves) {
    Scene* s = static_cast<Scene*>(h);
    *width = s->width;
    *height = s->height;
    std::memcpy(background, s->background, 16);
    *culled = s->culled;
    *curves = s->curves;
}

}  // extern "C"
