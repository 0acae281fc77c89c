// strand_build.cuh — level 3 of the ABI: raw outlines + transformer stacks -> geometry heap + shape
// entries, on the GPU.  In the reference this is Haskell, run on one core per
// frame: onShape (Raster/Serialize.hs:148-177: transform, bounding box, canvas culling, geometry pile),
// enclose (Raster/Enclosure.hs:62-73), outlineToStrands / splitShape (Raster/Strand.hs:153-178),
// replaceKnobs (Raster/Deknob.hs:36-108) and the reorder table (Raster/ReorderTable.hs:44-110).
// Paths relative to /root/reference/src/Graphics/Gudni/.
//
// The functions below are plain C++ over pointers (no CUDA types) so that the same text also compiles
// for the host in tests/strand_host_check.cpp, where it is held byte for byte to the harness's own
// restatement (csrc/host/strand.hpp) before it ever runs on a GPU.  One logical worker per shape:
//   measure  — walks the shape's outlines once: bounding box, strand count, 16-byte units of geometry;
//   emit     — walks them again and writes the strands (header + points in tree order) at the shape's
//              offset, which an exclusive scan over the kept shapes produced in between.
// Arithmetic is IEEE f32 in the reference's operation order (no FMA: the library is built -fmad=false).
#pragma once
#include <math.h>
#include <stdint.h>

#include "../../include/gudni_b200.h"

#if defined(__CUDACC__)
#define GUDNI_HD __host__ __device__ __forceinline__
#else
#define GUDNI_HD inline
#endif

namespace gudni_strands {

constexpr int kMaxBeziersPerStrand = 16;   // mAXsECTIONsIZE `div` 2, Raster/Constants.hs:67, Strand.hs:178
constexpr int kMaxStrandPoints = 2 * kMaxBeziersPerStrand + 1;
constexpr float kIota = 0.0001f;           // Figure/Space.hs:102

struct Pt {
    float x, y;
};
struct Bez {
    Pt a, c, b;   // start, control, end
};

// ---- ReorderTable.hs:44-110: breadth-first order of a left-complete binary tree over the interior
// points, after [last, first, first control].  row(n)[k] = index into the strand's 2n+1 points.
struct ReorderTable {
    uint8_t row[kMaxBeziersPerStrand + 1][kMaxStrandPoints + 3];
};
inline int perfectTreePartition(int n) {   // :48-58
    int x = 1;
    while (x <= n / 2) x *= 2;
    return ((x / 2) - 1 <= n - x) ? x - 1 : n - (x / 2);
}
inline void buildReorderTable(ReorderTable& t) {
    for (int n = 0; n <= kMaxBeziersPerStrand; n++) {
        uint8_t* row = t.row[n];
        const int size = 2 * n + 1;
        for (int k = 0; k < kMaxStrandPoints + 3; k++) row[k] = 0;
        if (size < 3) continue;
        int out = 0;
        row[out++] = (uint8_t)(size - 1);
        row[out++] = 0;
        row[out++] = 1;
        // buildITree over [0, n-1) interior nodes, emitted level by level (breadth, :72-80)
        struct Range { int lo, len; };
        Range level[64], next[64];
        int nLevel = 0;
        level[nLevel++] = {0, n - 1};
        while (nLevel > 0) {
            int nNext = 0;
            for (int i = 0; i < nLevel; i++) {
                if (level[i].len <= 0) continue;
                const int half = perfectTreePartition(level[i].len);
                const int centre = level[i].lo + half;
                row[out++] = (uint8_t)(2 * centre + 2);      // makeTreeRow: (*2), [x, x+1], then (+2)
                row[out++] = (uint8_t)(2 * centre + 3);
                next[nNext++] = {level[i].lo, half};
                next[nNext++] = {centre + 1, level[i].len - half - 1};
            }
            nLevel = nNext;
            for (int i = 0; i < nNext; i++) level[i] = next[i];
        }
    }
}

// ---- Figure/Transformer.hs:94-105 applied innermost (last listed) first -----------------------------
GUDNI_HD Pt applyStack(const gudni_transform* tr, uint32_t first, uint32_t count, Pt p) {
    for (uint32_t i = count; i-- > 0;) {
        const gudni_transform t = tr[first + i];
        if (t.kind == GUDNI_TRANSFORM_TRANSLATE) { p.x = p.x + t.a; p.y = p.y + t.b; }
        else if (t.kind == GUDNI_TRANSFORM_SCALE) { p.x = p.x * t.a; p.y = p.y * t.a; }
        else { const float x = p.x * t.a - p.y * t.b, y = p.y * t.a + p.x * t.b; p.x = x; p.y = y; }   // Angle.hs:52-53
    }
    return p;
}

// ---- Deknob.hs:40-101 -----------------------------------------------------------------------------------
GUDNI_HD Pt between(float t, Pt v0, Pt v1) {
    Pt r;
    r.x = v0.x * (1.f - t) + v1.x * t;
    r.y = v0.y * (1.f - t) + v1.y * t;
    return r;
}
// findSplit: `leftward` selects isLeftOf, else isRightOf.  Returns mid0, onCurve, mid1.
GUDNI_HD void findSplit(bool leftward, const Bez& z, Pt& mid0, Pt& onCurve, Pt& mid1) {
    float bottom = 0.f, top = 1.f, t = 0.5f;
    for (;;) {
        mid0 = between(t, z.a, z.c);
        mid1 = between(t, z.c, z.b);
        onCurve = between(t, mid0, mid1);
        if (top - bottom <= kIota) return;
        const bool stays1 = leftward ? (mid1.x < onCurve.x) : (mid1.x > onCurve.x);
        const bool stays0 = leftward ? (mid0.x < onCurve.x) : (mid0.x > onCurve.x);
        if (stays1) { const float nt = t + ((top - t) / 2.f); bottom = t; t = nt; }
        else if (stays0) { const float nt = bottom + ((t - bottom) / 2.f); top = t; t = nt; }
        else return;
    }
}
// fixKnob: 1 or 2 knob-free Béziers for one curve section.
GUDNI_HD int fixKnob(const Bez& z, Bez out[2]) {
    const bool left = z.c.x < z.a.x && z.c.x < z.b.x;
    const bool right = z.c.x > z.a.x && z.c.x > z.b.x;
    if (!left && !right) { out[0] = z; return 1; }
    Pt m0, on, m1;
    findSplit(left, z, m0, on, m1);
    out[0].a = z.a; out[0].c = m0; out[0].b = on;
    out[1].a = on; out[1].c = m1; out[1].b = z.b;
    return 2;
}

GUDNI_HD int compareHorizontal(const Bez& z) { return (z.a.x < z.b.x) ? -1 : (z.a.x > z.b.x ? 1 : 0); }   // Strand.hs:68-69

// What a walk over one outline reports to its visitor: every strand (a chunk of <= 16 connectable
// Béziers) as its 2*len+1 points in walking order, with `inLastRun` telling whether it belongs to the run
// still being accumulated when the outline ends — splitIntoStrands (Strand.hs:98-102) puts that run FIRST.
template <class Visitor>
GUDNI_HD void walkOutline(const gudni_curve_pair* pairs, const gudni_outline ol, const gudni_transform* tr, uint32_t trFirst,
                          uint32_t trCount, int lastRunStart, Visitor& visit) {
    const uint32_t n = ol.n_pairs;
    if (n < 2) return;                                              // Strand.hs:175-177
    Pt pts[kMaxStrandPoints];
    int len = 0;            // Béziers in the open chunk
    int prevDir = 0;
    int index = 0;          // knob-free Béziers seen so far
    bool open = false;
    const gudni_curve_pair p0 = pairs[ol.first_pair];
    Pt firstOn = applyStack(tr, trFirst, trCount, Pt{p0.on_x, p0.on_y});
    Pt on = firstOn;
    for (uint32_t i = 0; i < n; i++) {
        const gudni_curve_pair cur = pairs[ol.first_pair + i];
        Bez z;
        z.a = on;
        z.c = applyStack(tr, trFirst, trCount, Pt{cur.off_x, cur.off_y});
        if (i + 1 < n) {
            const gudni_curve_pair nx = pairs[ol.first_pair + i + 1];
            z.b = applyStack(tr, trFirst, trCount, Pt{nx.on_x, nx.on_y});
        } else {
            z.b = firstOn;                                          // overNeighbors wraps around, Strand.hs:116-124
        }
        on = z.b;
        visit.section(z);
        Bez fixed[2];
        const int nf = fixKnob(z, fixed);
        for (int f = 0; f < nf; f++) {
            const Bez& q = fixed[f];
            const int dir = compareHorizontal(q);
            const bool joins = open && dir == prevDir && dir != 0;   // connectable, Strand.hs:77-81
            if (open && (!joins || len == kMaxBeziersPerStrand)) {   // run ends, or splitTooLarge (:105-109)
                visit.strand(pts, len, lastRunStart >= 0 && index - len >= lastRunStart);
                len = 0;
            }
            if (!joins) visit.runStartsAt(index);
            if (len == 0) pts[0] = q.a;
            pts[2 * len + 1] = q.c;
            pts[2 * len + 2] = q.b;
            len++;
            prevDir = dir;
            open = true;
            index++;
        }
    }
    if (open) visit.strand(pts, len, lastRunStart >= 0 && index - len >= lastRunStart);
}

// ---- pass 1: measure ------------------------------------------------------------------------------------
struct ShapeMeasure {
    float left, top, right, bottom;
    uint32_t units;      // 16-byte units of geometry
    uint32_t strands;
};
struct MeasureVisitor {
    ShapeMeasure m;
    int lastRun;
    GUDNI_HD void section(const Bez& z) {
        // boxOf over on- and off-curve points of the transformed outline (Figure/Outline.hs:123-133)
        m.left = fminf(m.left, fminf(z.a.x, z.c.x));
        m.right = fmaxf(m.right, fmaxf(z.a.x, z.c.x));
        m.top = fminf(m.top, fminf(z.a.y, z.c.y));
        m.bottom = fmaxf(m.bottom, fmaxf(z.a.y, z.c.y));
    }
    GUDNI_HD void runStartsAt(int index) { lastRun = index; }
    GUDNI_HD void strand(const Pt*, int len, bool) { m.units += (uint32_t)len + 1u; m.strands += 1u; }
};

// Bounding box of a shape whose outlines are too short to produce strands still counts its points
// (onShape boxes the outlines before enclose looks at their length).
GUDNI_HD ShapeMeasure measureShape(const gudni_outline_shape& s, const gudni_outline* outlines, const gudni_curve_pair* pairs,
                                   const gudni_transform* tr) {
    MeasureVisitor v;
    v.m.left = v.m.top = __builtin_huge_valf();
    v.m.right = v.m.bottom = -__builtin_huge_valf();
    v.m.units = v.m.strands = 0;
    for (uint32_t o = 0; o < s.n_outlines; o++) {
        const gudni_outline ol = outlines[s.first_outline + o];
        v.lastRun = -1;
        if (ol.n_pairs < 2) {
            for (uint32_t i = 0; i < ol.n_pairs; i++) {
                const gudni_curve_pair p = pairs[ol.first_pair + i];
                Bez z;
                z.a = applyStack(tr, s.first_transform, s.n_transforms, Pt{p.on_x, p.on_y});
                z.c = applyStack(tr, s.first_transform, s.n_transforms, Pt{p.off_x, p.off_y});
                z.b = z.a;
                v.section(z);
            }
            continue;
        }
        walkOutline(pairs, ol, tr, s.first_transform, s.n_transforms, -1, v);
    }
    return v.m;
}

// excludeBox, Raster/Serialize.hs:97-104
GUDNI_HD bool culled(const ShapeMeasure& m, int width, int height) {
    return m.left >= (float)width || m.top >= (float)height || m.right <= 0.f || m.bottom <= 0.f;
}

// ---- pass 2: emit ---------------------------------------------------------------------------------------
struct RunVisitor {      // first walk of an outline in the emit pass: where does the last run start, how big is it
    int lastRun;
    uint32_t lastRunUnits, units;
    GUDNI_HD void section(const Bez&) {}
    GUDNI_HD void runStartsAt(int index) { lastRun = index; lastRunUnits = 0; }
    GUDNI_HD void strand(const Pt*, int len, bool) { units += (uint32_t)len + 1u; lastRunUnits += (uint32_t)len + 1u; }
};
struct EmitVisitor {
    uint8_t* heap;               // geometry heap
    const ReorderTable* table;
    uint64_t cursorLast, cursorRest;   // byte offsets: the last run goes first, the others behind it
    GUDNI_HD void section(const Bez&) {}
    GUDNI_HD void runStartsAt(int) {}
    GUDNI_HD void strand(const Pt* pts, int len, bool inLastRun) {
        uint64_t& cursor = inLastRun ? cursorLast : cursorRest;
        const int size = 2 * len + 1;
        const bool backwards = pts[0].x > pts[size - 1].x;          // reverseIfBackwards, Strand.hs:137-143
        // StorableM Strand (Strand.hs:190-206): u16 size in 8-byte units (header + points), u16 0, u32 0
        uint32_t* hdr = reinterpret_cast<uint32_t*>(heap + cursor);
        hdr[0] = (uint32_t)(size + 1);
        hdr[1] = 0u;
        Pt* out = reinterpret_cast<Pt*>(heap + cursor + 8);
        const uint8_t* row = table->row[len];
        for (int k = 0; k < size; k++) {
            const int j = row[k];
            out[k] = backwards ? pts[size - 1 - j] : pts[j];        // beziersToPoints + reorder (:129-150)
        }
        cursor += 16ull * (uint64_t)(len + 1);
    }
};

GUDNI_HD void emitShape(const gudni_outline_shape& s, const gudni_outline* outlines, const gudni_curve_pair* pairs,
                        const gudni_transform* tr, const ReorderTable* table, uint8_t* heap, uint64_t byteOffset) {
    for (uint32_t o = 0; o < s.n_outlines; o++) {
        const gudni_outline ol = outlines[s.first_outline + o];
        if (ol.n_pairs < 2) continue;
        RunVisitor rv;
        rv.lastRun = -1;
        rv.lastRunUnits = rv.units = 0;
        walkOutline(pairs, ol, tr, s.first_transform, s.n_transforms, -1, rv);
        EmitVisitor ev;
        ev.heap = heap;
        ev.table = table;
        ev.cursorLast = byteOffset;
        ev.cursorRest = byteOffset + 16ull * rv.lastRunUnits;
        walkOutline(pairs, ol, tr, s.first_transform, s.n_transforms, rv.lastRun, ev);
        byteOffset += 16ull * rv.units;
    }
}

}  // namespace gudni_strands
