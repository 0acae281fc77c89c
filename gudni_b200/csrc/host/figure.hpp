// figure.hpp — the slice of Gudni's front end (Graphics.Gudni.Figure / Layout) that the harness
// needs in order to PRODUCE hot-path inputs: points, segments, open curves, outlines, transforms.
// All arithmetic is IEEE f32 like the reference's `SubSpace` (Figure/Space.hs:73-75).
//
// This is not part of the rasterizer product; in a real deployment the Haskell front end produces
// these bytes.  Paths cited are relative to /root/reference/src/Graphics/Gudni/.
#pragma once
#include <cmath>
#include <cstdint>
#include <vector>

namespace gudni {

struct P2 {
    float x = 0.f, y = 0.f;
};
inline P2 operator+(P2 a, P2 b) { return {a.x + b.x, a.y + b.y}; }
inline P2 operator-(P2 a, P2 b) { return {a.x - b.x, a.y - b.y}; }
inline P2 operator*(P2 a, float s) { return {a.x * s, a.y * s}; }
inline bool operator==(P2 a, P2 b) { return a.x == b.x && a.y == b.y; }

// Figure/Segment.hs:39-42 — an on-curve anchor and an optional control point.
struct Segment {
    P2 anchor;
    bool curved = false;
    P2 control;
};
inline Segment straight(float x, float y) { return {{x, y}, false, {}}; }
inline Segment curvedSeg(float x, float y, float cx, float cy) { return {{x, y}, true, {cx, cy}}; }

// Figure/OpenCurve.hs:55-59
struct OpenCurve {
    std::vector<Segment> segments;
    P2 terminator;
};

// Figure/Outline.hs:55-57 — (onCurve, offCurve)
struct CurvePair {
    P2 on, off;
};
using Outline = std::vector<CurvePair>;

// Figure/Transformer.hs:94-105.  Angles are kept in radians.
struct Transform {
    enum Kind { Translate, Scale, Rotate } kind;
    P2 delta{};
    float factor = 1.f;  // scale factor or angle in radians
    static Transform translate(float x, float y) { return {Translate, {x, y}, 1.f}; }
    static Transform scale(float s) { return {Scale, {}, s}; }
    static Transform rotateRad(float a) { return {Rotate, {}, a}; }
    static Transform rotateTurn(float t) { return {Rotate, {}, t * 6.283185307179586f}; }
};

// Figure/Angle.hs:52-53
inline P2 rotatePoint(float a, P2 p) {
    float c = std::cos(a), s = std::sin(a);
    return {p.x * c - p.y * s, p.y * c + p.x * s};
}

inline P2 applyTransform(const Transform& t, P2 p) {
    switch (t.kind) {
        case Transform::Translate: return p + t.delta;            // tTranslate = (^+^)
        case Transform::Scale: return p * t.factor;               // tScale = flip (^*)
        default: return rotatePoint(t.factor, p);
    }
}

// A transformer stack is applied innermost first (Raster/TraverseShapeTree.hs:63,
// Figure/Transformer.hs:105: CombineTransform a b = b . a).  `stack` lists transforms in the
// order the client wrote them, outermost first — `tTranslate p . tScale s . tRotate a $ shape` is
// {translate p, scale s, rotate a} — so application walks it backwards.
inline P2 applyStack(const std::vector<Transform>& stack, P2 p) {
    for (size_t i = stack.size(); i-- > 0;) p = applyTransform(stack[i], p);
    return p;
}

inline Outline transformOutline(const std::vector<Transform>& stack, const Outline& o) {
    Outline r(o.size());
    for (size_t i = 0; i < o.size(); i++) r[i] = {applyStack(stack, o[i].on), applyStack(stack, o[i].off)};
    return r;
}

inline OpenCurve overCurve(const OpenCurve& c, const Transform& t) {
    OpenCurve r;
    r.segments.reserve(c.segments.size());
    for (const Segment& s : c.segments) {
        Segment q = s;
        q.anchor = applyTransform(t, s.anchor);
        if (s.curved) q.control = applyTransform(t, s.control);
        r.segments.push_back(q);
    }
    r.terminator = applyTransform(t, c.terminator);
    return r;
}

// Figure/OpenCurve.hs:91-95 — (<^>): translate c1 so it starts where c0 ends, then append.
inline OpenCurve joinCurves(const OpenCurve& c0, const OpenCurve& c1) {
    P2 delta = c0.terminator - c1.segments.front().anchor;
    OpenCurve moved = overCurve(c1, Transform::translate(delta.x, delta.y));
    OpenCurve r;
    r.segments = c0.segments;
    r.segments.insert(r.segments.end(), moved.segments.begin(), moved.segments.end());
    r.terminator = moved.terminator;
    return r;
}

// Figure/OpenCurve.hs:81-88
inline OpenCurve reverseCurve(const OpenCurve& c) {
    std::vector<Segment> ext = c.segments;
    ext.push_back({c.terminator, false, {}});
    std::vector<Segment> rev;
    for (size_t i = 0; i + 1 < ext.size(); i++) {  // pullSegments (Seg o0 c0) (Seg o1 c1) = Seg o1 c0
        Segment s;
        s.anchor = ext[i + 1].anchor;
        s.curved = ext[i].curved;
        s.control = ext[i].control;
        rev.push_back(s);
    }
    OpenCurve r;
    r.segments.assign(rev.rbegin(), rev.rend());
    r.terminator = c.segments.front().anchor;
    return r;
}

// Util/Plot.hs:48-55 — arcs are built from quadratic pieces of < 45 degrees.
inline OpenCurve makeArcSegment(float rad) {
    OpenCurve c;
    c.segments.push_back(curvedSeg(1.f, 0.f, 1.f, std::tan(rad / 2.f)));
    c.terminator = {std::cos(rad), std::sin(rad)};
    return c;
}
inline OpenCurve makeArc(float rad) {
    float deg = rad * (180.f / 3.14159265358979323846f);
    if (std::fabs(deg) < 45.f) return makeArcSegment(rad);
    OpenCurve half = makeArc(rad / 2.f);
    return joinCurves(half, overCurve(half, Transform::rotateRad(rad / 2.f)));
}

// Figure/Outline.hs:108-114
inline std::vector<Segment> closeOpenCurve(const OpenCurve& c) {
    std::vector<Segment> segs;
    if (!(c.terminator == c.segments.front().anchor)) segs.push_back({c.terminator, false, {}});
    segs.insert(segs.end(), c.segments.begin(), c.segments.end());
    return segs;
}

// Figure/Outline.hs:90-101 — straight segments get a control point at the midpoint.
inline P2 lerpHalf(P2 a, P2 b) {  // Linear.lerp 0.5 a b = 0.5 *^ a ^+^ 0.5 *^ b
    return {0.5f * a.x + 0.5f * b.x, 0.5f * a.y + 0.5f * b.y};
}
inline Outline segmentsToOutline(const std::vector<Segment>& segs) {
    Outline o;
    if (segs.empty()) return o;
    P2 first = segs.front().anchor;
    for (size_t i = 0; i < segs.size(); i++) {
        const Segment& s = segs[i];
        if (s.curved) {
            o.push_back({s.anchor, s.control});
        } else {
            P2 next = (i + 1 < segs.size()) ? segs[i + 1].anchor : first;
            o.push_back({s.anchor, lerpHalf(s.anchor, next)});
        }
    }
    return o;
}

// Layout/Draw.hs:72-78, 134, 153-154
inline Outline rectangleOutline(float w, float h) {
    return segmentsToOutline({straight(0, 0), straight(w, 0), straight(w, h), straight(0, h)});
}
inline Outline arcOutline(float rad) { return segmentsToOutline(closeOpenCurve(makeArc(rad))); }
inline Outline circleOutline() { return arcOutline(6.283185307179586f); }

}  // namespace gudni
