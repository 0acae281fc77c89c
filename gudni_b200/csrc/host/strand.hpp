// strand.hpp — outline -> knob-free Béziers -> x-monotone strands -> implicit-binary-tree point
// order -> geometry-heap bytes.  This is the wire-format PRODUCER of the hot path (SURVEY.md §2
// row 5): in the reference it is Haskell (Raster/Strand.hs, Deknob.hs, ReorderTable.hs,
// Enclosure.hs); the harness restates it so that tests and bench.py can generate the exact byte
// layout `Kernels.cl:1365-1389` parses.  Paths relative to /root/reference/src/Graphics/Gudni/.
#pragma once
#include <algorithm>
#include <cstring>
#include <vector>

#include "figure.hpp"

namespace gudni {

struct Bezier {
    P2 a, c, b;  // start, control, end
};

constexpr int kMaxSectionSize = 32;        // mAXsECTIONsIZE, Raster/Constants.hs:67
constexpr float kIota = 0.0001f;           // Figure/Space.hs:101-102

// ---- Deknob.hs ----------------------------------------------------------------------------------
inline P2 between(float t, P2 v0, P2 v1) {  // :40-41
    return {v0.x * (1.f - t) + v1.x * t, v0.y * (1.f - t) + v1.y * t};
}
struct SplitPoints {
    P2 mid0, onCurve, mid1;
};
inline SplitPoints curvePoint(float t, P2 v0, P2 c, P2 v1) {  // :45-50
    P2 m0 = between(t, v0, c), m1 = between(t, c, v1);
    return {m0, between(t, m0, m1), m1};
}
// :61-79 — bisection on the curve parameter until the vertical tangent is bracketed within iota.
template <class Rel>
inline SplitPoints findSplit(Rel staysRelativeTo, const Bezier& z) {
    float bottom = 0.f, top = 1.f, t = 0.5f;
    for (;;) {
        SplitPoints s = curvePoint(t, z.a, z.c, z.b);
        if (top - bottom <= kIota) return s;
        if (staysRelativeTo(s.mid1, s.onCurve)) {
            float nt = t + ((top - t) / 2.f);
            bottom = t;
            t = nt;
        } else if (staysRelativeTo(s.mid0, s.onCurve)) {
            float nt = bottom + ((t - bottom) / 2.f);
            top = t;
            t = nt;
        } else {
            return s;
        }
    }
}
inline void fixKnob(const Bezier& z, std::vector<Bezier>& out) {  // :85-101
    auto leftOf = [](P2 p, P2 q) { return p.x < q.x; };
    auto rightOf = [](P2 p, P2 q) { return p.x > q.x; };
    if (leftOf(z.c, z.a) && leftOf(z.c, z.b)) {
        SplitPoints s = findSplit(leftOf, z);
        out.push_back({z.a, s.mid0, s.onCurve});
        out.push_back({s.onCurve, s.mid1, z.b});
    } else if (rightOf(z.c, z.a) && rightOf(z.c, z.b)) {
        SplitPoints s = findSplit(rightOf, z);
        out.push_back({z.a, s.mid0, s.onCurve});
        out.push_back({s.onCurve, s.mid1, z.b});
    } else {
        out.push_back(z);
    }
}

// ---- ReorderTable.hs ----------------------------------------------------------------------------
inline int perfectTreePartition(int n) {  // :48-62
    int x = 1;
    while (x <= n / 2) x *= 2;
    return ((x / 2) - 1 <= n - x) ? x - 1 : n - (x / 2);
}
// breadth-first list of the centres of the left-complete tree over [lo, lo+len)  (:64-87)
inline std::vector<int> breadthOrder(int n) {
    struct Range {
        int lo, len;
    };
    std::vector<int> order;
    std::vector<Range> level{{0, n}};
    while (!level.empty()) {
        std::vector<Range> next;
        for (const Range& r : level) {
            if (r.len <= 0) continue;
            int half = perfectTreePartition(r.len);
            order.push_back(r.lo + half);
            next.push_back({r.lo, half});
            next.push_back({r.lo + half + 1, r.len - half - 1});
        }
        level.swap(next);
    }
    return order;
}
// :96-104 — index list for a strand of `size` points (size = 2n+1)
inline std::vector<int> reorderForExtents(int size) {
    std::vector<int> idx;
    if (size < 3) return idx;
    idx = {size - 1, 0, 1};
    if (size > 3) {
        for (int c : breadthOrder((size - 3) / 2)) {
            idx.push_back(2 * c + 2);
            idx.push_back(2 * c + 3);
        }
    }
    return idx;
}

// ---- Strand.hs ----------------------------------------------------------------------------------
inline int compareHorizontal(const Bezier& z) { return (z.a.x < z.b.x) ? -1 : (z.a.x > z.b.x ? 1 : 0); }
inline bool connectable(const Bezier& p, const Bezier& q) {  // :77-81
    int hp = compareHorizontal(p);
    return hp == compareHorizontal(q) && hp != 0;
}

// One strand's points, already in tree order.
using StrandPoints = std::vector<P2>;

// outlineToStrands (:171-178) = splitShape (:153-168)
inline void outlineToStrands(const Outline& outline, std::vector<StrandPoints>& strands) {
    const size_t n = outline.size();
    if (n < 2) return;
    // pairsToBeziers (:124-126): each pair with its next neighbour, wrapping around.
    std::vector<Bezier> beziers;
    for (size_t i = 0; i < n; i++) {
        Bezier z{outline[i].on, outline[i].off, outline[(i + 1) % n].on};
        fixKnob(z, beziers);  // replaceKnobs (Deknob.hs:104-108)
    }
    // splitIntoStrands (:98-102): fold left; the run still being accumulated at the end is put FIRST.
    std::vector<std::vector<Bezier>> runs;
    std::vector<Bezier> acc;
    for (const Bezier& z : beziers) {
        if (acc.empty() || connectable(acc.back(), z)) {
            acc.push_back(z);
        } else {
            runs.push_back(acc);
            acc.assign(1, z);
        }
    }
    runs.insert(runs.begin(), acc);
    const size_t maxBeziers = kMaxSectionSize / 2;  // sectionSize `div` 2 (:178)
    for (const auto& run : runs) {
        // splitTooLarge (:105-109) cuts off maxSize-long prefixes while more than maxSize remain,
        // which is plain chunking; runs are never empty.
        for (size_t start = 0; start < run.size(); start += maxBeziers) {
            size_t len = std::min(maxBeziers, run.size() - start);
            std::vector<Bezier> part(run.begin() + start, run.begin() + start + len);
            // reverseIfBackwards (:137-143)
            if (part.front().a.x > part.back().b.x) {
                std::vector<Bezier> rev;
                for (size_t i = part.size(); i-- > 0;) rev.push_back({part[i].b, part[i].c, part[i].a});
                part.swap(rev);
            }
            // beziersToPoints (:129-134)
            std::vector<P2> pts;
            for (const Bezier& z : part) {
                pts.push_back(z.a);
                pts.push_back(z.c);
            }
            pts.push_back(part.back().b);
            // reorder (:147-150)
            std::vector<int> idx = reorderForExtents((int)pts.size());
            StrandPoints sp(pts.size());
            for (size_t i = 0; i < pts.size(); i++) sp[i] = pts[idx[i]];
            strands.push_back(std::move(sp));
        }
    }
}

// StorableM Strand (:190-206): u16 size in 8-byte units, u16 0, u32 0, then the points.
inline void appendStrandBytes(const StrandPoints& sp, std::vector<uint8_t>& heap) {
    uint16_t size = (uint16_t)(1 + sp.size());
    uint8_t hdr[8] = {0};
    std::memcpy(hdr, &size, 2);
    heap.insert(heap.end(), hdr, hdr + 8);
    const uint8_t* p = reinterpret_cast<const uint8_t*>(sp.data());
    heap.insert(heap.end(), p, p + sp.size() * sizeof(P2));
}

}  // namespace gudni
