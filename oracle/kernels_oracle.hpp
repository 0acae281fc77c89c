// kernels_oracle.hpp — TEST INFRASTRUCTURE, NOT PRODUCT.
//
// CPU restatement of the reference's three OpenCL kernels, function for function, in IEEE f32
// with no FMA contraction (build with -ffp-contract=off) and no fast-math.  Every function cites
// the lines of /root/reference/src/Graphics/Gudni/OpenCL/Kernels.cl ("K.cl") it follows.
//
// PINNED TO THE REFERENCE'S OWN KERNEL SOURCE.  The reference has no automated tests, golden vectors or
// fixtures for this path, and its Haskell + OpenCL host cannot run in this image (no GHC, no OpenCL
// runtime: SURVEY.md §4, §8(c)).  Its kernel file, however, compiles for the host: oracle/refbuild/
// builds K.cl with g++ against an OpenCL-C compatibility header into oracle/_ref/libgudni_ref.so, and
// tests/test_reference_pin.py requires this restatement to reproduce its per-thread threshold counts,
// shape-bit counts and BGRA words exactly (every catalogue scene, seeded random scenes under three
// RasterSpecs, pictures, glyphs, and S4b / S4 at full size: 14,114,195 thresholds, 8.3 M pixels, 0
// differences).  tests/golden/ holds vectors that library produced.  What stays outside the pin:
// -cl-fast-relaxed-math on a real OpenCL device (OpenCL/Setup.hs:129), not reproducible by definition.
// Hand-derived known answers (tests/test_oracle_known_answers.py) and an independent exact-area renderer
// (tests/test_oracle_exact_area.py) check that the algorithm itself computes what it claims.
//
// Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
// use this file.  Dead code of K.cl is deliberately not restated.
#pragma once
#include <cfloat>
#include <cmath>
#include <cstdint>
#include <cstring>

#include "../include/gudni_b200.h"

namespace oracle {

struct float2 { float x, y; };
struct float4 { float x, y, z, w; };  // also THRESHOLD: (top, bottom, left, right)  K.cl:218-225
typedef uint32_t HEADER;

// ---- constants, K.cl:45-47, 83-91, 168-185, 216, 249-250 -----------------------------------------
static const float MINCROP = 0.2f;
static const float LEFTBORDER = 0.0f;
static const float RIGHTBORDER = 1.0f;
static const float MAXCHANNELFLOAT = 255.0f;
static const HEADER POSITIVE_SLOPE_MASK = 0x80000000u;
static const HEADER PERSIST_AND_SLOPE_MASK = 0xC0000000u;
static const HEADER PERSIST_TOP = 0xC0000000u;
static const HEADER PERSIST_BOTTOM = 0x40000000u;
static const HEADER POSITIVE_SLOPE = 0x80000000u;
static const HEADER NEGATIVE_SLOPE = 0x00000000u;
static const HEADER PERSIST = 0x40000000u;
static const HEADER NONPERSIST = 0x00000000u;
static const HEADER UNPERSISTMASK = 0xBFFFFFFFu;
static const HEADER SHAPEBIT_MASK = 0x0FFFFFFFu;
static const float VERTICALSLOPE = FLT_MAX;
static const float TAXICAB_FLATNESS = 0.25f;
static const int SHAPESTACKSECTIONS = 8;   // Raster/Constants.hs:51
static const int MAXSHAPE_LIMIT = 127;     // Raster/Constants.hs:55-58

// ---- header / threshold accessors, K.cl:188-246 --------------------------------------------------
static inline bool headerPositiveSlope(HEADER h) { return (h & POSITIVE_SLOPE_MASK) != 0; }
static inline bool headerPersistTop(HEADER h) { return (h & PERSIST_AND_SLOPE_MASK) == PERSIST_TOP; }
static inline bool headerPersistBottom(HEADER h) { return (h & PERSIST_AND_SLOPE_MASK) == PERSIST_BOTTOM; }
static inline bool headerPersistEither(HEADER h) { return (h & PERSIST) != 0; }
static inline uint32_t headerShapeBit(HEADER h) { return h & SHAPEBIT_MASK; }
static inline HEADER unPersist(HEADER h) { return h & UNPERSISTMASK; }

static inline float4 makeThreshold(float top, float bottom, float left, float right) { return {top, bottom, left, right}; }
static inline float tTop(float4 t) { return t.x; }
static inline float tBottom(float4 t) { return t.y; }
static inline float tLeft(float4 t) { return t.z; }
static inline float tRight(float4 t) { return t.w; }
static inline float2 tStart(HEADER h, float4 t) { return {tLeft(t), headerPositiveSlope(h) ? tTop(t) : tBottom(t)}; }
static inline float tTopX(HEADER h, float4 t) { return headerPositiveSlope(h) ? tLeft(t) : tRight(t); }
static inline float tHeight(float4 t) { return tBottom(t) - tTop(t); }
static inline bool tIsHorizontal(float4 t) { return tTop(t) == tBottom(t); }
static inline bool tKeep(HEADER h, float4 t) { return headerPersistEither(h) || (tHeight(t) >= MINCROP); }

static inline float thresholdInvertedSlope(HEADER header, float4 t) {  // K.cl:240-246
    float slopeSign = headerPositiveSlope(header) ? 1.0f : -1.0f;
    return tIsHorizontal(t) ? VERTICALSLOPE : (tRight(t) - tLeft(t)) / (tBottom(t) - tTop(t)) * slopeSign;
}

// ---- structures, K.cl:312-341, 349-354, 417-458 --------------------------------------------------
struct Slice { int32_t sStart, sLength; };

struct ShapeState {  // K.cl:417-421 (1,088 bytes in the reference)
    uint32_t shapeBits;
    uint64_t shapeIndices[MAXSHAPE_LIMIT];
    uint64_t shapeStack[SHAPESTACKSECTIONS];
};

struct TileState {  // K.cl:438-451 (tileIndex / threadUnique dropped: never read on a live path)
    uint32_t tileShapeStart;
    int tileNumShapes;
    int threadId;
    int tileSizeX, tileSizeY;
    int bitmapW, bitmapH;
    int internalDeltaX, internalDeltaY;
    int threadDeltaX, threadDeltaY;
    int intHeight;
    float floatHeight;
    int column;
};

struct Traversal {  // K.cl:453-468
    float4 travLeftControl;  // left.x left.y control.x control.y
    float2 travRight;
    float travXPos;
    int travIndex;
};

struct ThresholdQueue {  // K.cl:369-373
    HEADER* thresholdHeaders;
    float4* thresholds;
    Slice qSlice;
    int capacity;     // MAXTHRESHOLDS
    bool overflow;    // set instead of the reference's out-of-bounds write
};

// ---- queue, K.cl:375-411.  cycleLocation never wraps while sLength <= MAXTHRESHOLDS because
// sStart + sLength is invariant (= MAXTHRESHOLDS), so locations are sStart + i. ----------------------
static inline int cycleLocation(const ThresholdQueue* tQ, int i) { return i > tQ->capacity ? i - tQ->capacity : i; }
static inline int tSLocation(const ThresholdQueue* tQ, int i) { return cycleLocation(tQ, tQ->qSlice.sStart + i); }
static inline float4 getThreshold(const ThresholdQueue* tQ, int i) { return tQ->thresholds[tSLocation(tQ, i)]; }
static inline void setThreshold(ThresholdQueue* tQ, int i, float4 v) { tQ->thresholds[tSLocation(tQ, i)] = v; }
static inline HEADER getHeader(const ThresholdQueue* tQ, int i) { return tQ->thresholdHeaders[tSLocation(tQ, i)]; }
static inline void setHeader(ThresholdQueue* tQ, int i, HEADER v) { tQ->thresholdHeaders[tSLocation(tQ, i)] = v; }
// pushTopSlot, K.cl:402-405.  The reference does not check capacity; the oracle flags it and
// refuses the push so that it never writes outside the thread's slice.
static inline bool pushTopSlot(ThresholdQueue* tQ) {
    if (tQ->qSlice.sLength >= tQ->capacity) { tQ->overflow = true; return false; }
    tQ->qSlice.sStart = cycleLocation(tQ, tQ->qSlice.sStart - 1);
    tQ->qSlice.sLength += 1;
    return true;
}
static inline void popTop(ThresholdQueue* tQ) {  // K.cl:407-411
    tQ->qSlice.sStart = cycleLocation(tQ, tQ->qSlice.sStart + 1);
    tQ->qSlice.sLength -= 1;
}

// ---- shape stack, K.cl:265-304 -------------------------------------------------------------------
static inline int clz64(uint64_t x) { return x == 0 ? 64 : __builtin_clzll(x); }
static inline uint64_t ignoreStack(uint64_t section, int ignoreBits) {  // K.cl:271-274
    return ignoreBits >= 64 ? section : (~(0xFFFFFFFFFFFFFFFFull << ignoreBits)) & section;
}
static inline int findTop(const uint64_t* shapeStack, int ignoreAbove) {  // K.cl:281-292
    int ignoreSection = ignoreAbove >> 6;
    int ignoreBits = ignoreAbove & 0x3F;
    uint64_t section = ignoreStack(shapeStack[ignoreSection], ignoreBits);
    while (section == 0 && ignoreSection > 0) {
        ignoreSection -= 1;
        section = shapeStack[ignoreSection];
    }
    int sectionBits = 64 - clz64(section);
    return (ignoreSection << 6) + sectionBits - 1;
}
static inline void flipBit(uint32_t shapeBit, uint64_t* shapeStack) {  // K.cl:298-304
    int section = shapeBit >> 6;
    int bit = shapeBit & 0x3F;
    shapeStack[section] ^= ((uint64_t)1 << bit);
}
static inline void passHeader(ShapeState* shS, HEADER h) { flipBit(headerShapeBit(h), shS->shapeStack); }  // K.cl:1515
static inline void passHeaderTop(ShapeState* shS, HEADER h) { if (headerPersistTop(h)) passHeader(shS, h); }
static inline void passHeaderBottom(ShapeState* shS, HEADER h) { if (headerPersistBottom(h)) passHeader(shS, h); }

// ---- small maths, K.cl:818-887 -------------------------------------------------------------------
static inline float yIntercept(float2 e0, float2 e1, float x) {  // K.cl:818-820
    return (((e1.y - e0.y) / (e1.x - e0.x)) * (x - e0.x)) + e0.y;
}
static inline float xInterceptInvertedSlope(float2 e, float invertedSlope, float y) {  // K.cl:832-834
    return ((y - e.y) * invertedSlope) + e.x;
}
static inline float2 midPoint(float2 v0, float2 v1) {  // K.cl:864-866, T = 0.5
    return {((1.0f - 0.5f) * v0.x) + (0.5f * v1.x), ((1.0f - 0.5f) * v0.y) + (0.5f * v1.y)};
}
static inline float taxiDistance(float2 v0, float2 v1) { return std::fabs(v1.x - v0.x) + std::fabs(v1.y - v0.y); }

static inline float4 composite(float4 fg, float4 bg) {  // K.cl:878-887
    float alphaOut = fg.w + bg.w * (1.0f - fg.w);
    if (alphaOut > 0) {
        float4 c;
        c.x = ((fg.x * fg.w) + (bg.x * bg.w * (1.0f - fg.w))) / alphaOut;
        c.y = ((fg.y * fg.w) + (bg.y * bg.w * (1.0f - fg.w))) / alphaOut;
        c.z = ((fg.z * fg.w) + (bg.z * bg.w * (1.0f - fg.w))) / alphaOut;
        c.w = alphaOut;
        return c;
    }
    return {0, 0, 0, 0};
}

// ---- threshold geometry, K.cl:895-1005 -----------------------------------------------------------
static inline float thresholdIntersectX(HEADER header, float4 t, float y) {  // K.cl:900-913
    if (tLeft(t) == tRight(t)) return tLeft(t);
    return xInterceptInvertedSlope(tStart(header, t), thresholdInvertedSlope(header, t), y);
}
// K.cl:916-926 — the definition's parameter order (clampLow, clampHigh) wins over the prototype's.
static inline float thresholdMidXLow(float4 t, HEADER h, float yTop, float yBottom, float clampLow, float clampHigh) {
    float yMid = yTop + ((yBottom - yTop) * 0.5f);
    float x = thresholdIntersectX(h, t, yMid);
    return x >= clampHigh ? clampLow : (clampLow < x ? x : clampLow);  // max(clampLow, x)
}
static inline void divideThreshold(HEADER* headerTop, float4* thresholdTop, HEADER* headerBottom,
                                   float4* thresholdBottom, float splitX, float splitY) {  // K.cl:928-957
    if (headerPositiveSlope(*headerTop)) {
        *thresholdBottom = makeThreshold(splitY, tBottom(*thresholdTop), splitX, tRight(*thresholdTop));
        *thresholdTop = makeThreshold(tTop(*thresholdTop), splitY, tLeft(*thresholdTop), splitX);
        *headerBottom = unPersist(*headerTop);
    } else {
        *thresholdBottom = makeThreshold(splitY, tBottom(*thresholdTop), tLeft(*thresholdTop), splitX);
        *thresholdTop = makeThreshold(tTop(*thresholdTop), splitY, splitX, tRight(*thresholdTop));
        *headerBottom = *headerTop;
        *headerTop = unPersist(*headerTop);
    }
}
static inline void splitThreshold(HEADER* topHeader, float4* top, HEADER* bottomHeader, float4* bottom, float splitY) {
    float splitX = thresholdIntersectX(*topHeader, *top, splitY);  // K.cl:959-979
    divideThreshold(topHeader, top, bottomHeader, bottom, splitX, splitY);
}
static inline void trimThresholdTop(HEADER* header, float4* threshold, float splitY) {  // K.cl:982-1005
    float splitX = thresholdIntersectX(*header, *threshold, splitY);
    if (headerPositiveSlope(*header)) {
        *threshold = makeThreshold(splitY, tBottom(*threshold), splitX, tRight(*threshold));
        *header = unPersist(*header);
    } else {
        *threshold = makeThreshold(splitY, tBottom(*threshold), tLeft(*threshold), splitX);
    }
}

// ---- ordering and insertion, K.cl:1079-1124 ------------------------------------------------------
static inline bool thresholdIsBelow(HEADER aHeader, float4 a, HEADER bHeader, float4 b) {  // K.cl:1079-1094
    return (tTop(a) > tTop(b)) ||
           ((tTop(a) == tTop(b)) &&
            ((tTopX(aHeader, a) > tTopX(bHeader, b)) ||
             ((tTopX(aHeader, a) == tTopX(bHeader, b)) &&
              (thresholdInvertedSlope(aHeader, a) > thresholdInvertedSlope(bHeader, b)))));
}
static inline void pushThreshold(ThresholdQueue* tQ, HEADER h, float4 t) {  // K.cl:1096-1103
    if (!pushTopSlot(tQ)) return;
    setHeader(tQ, 0, h);
    setThreshold(tQ, 0, t);
}
static inline void insertThreshold(ThresholdQueue* tQ, HEADER newHeader, float4 nw) {  // K.cl:1105-1124
    if (!pushTopSlot(tQ)) return;
    int cursor = 0;
    bool isBelow = true;
    while (cursor < (tQ->qSlice.sLength - 1) && isBelow) {
        HEADER oldHeader = getHeader(tQ, cursor + 1);
        float4 old = getThreshold(tQ, cursor + 1);
        isBelow = thresholdIsBelow(newHeader, nw, oldHeader, old);
        if (isBelow) {
            setHeader(tQ, cursor, oldHeader);
            setThreshold(tQ, cursor, old);
            cursor += 1;
        }
    }
    setHeader(tQ, cursor, newHeader);
    setThreshold(tQ, cursor, nw);
}

// ---- active-set slicing, K.cl:1007-1077 ----------------------------------------------------------
static inline int countActive(ThresholdQueue* tQ, float* nextTop) {  // K.cl:1007-1024
    float top = tTop(getThreshold(tQ, 0));
    bool notDone = true;
    int numActive = 1;
    while (notDone && numActive < tQ->qSlice.sLength) {
        float4 next = getThreshold(tQ, numActive);
        if (tTop(next) > top) {
            notDone = false;
            *nextTop = tTop(next);
        } else {
            numActive += 1;
        }
    }
    return numActive;
}
static inline float nextSlicePoint(ThresholdQueue* tQ, float slicePoint, int numActive) {  // K.cl:1026-1038
    float top = tTop(getThreshold(tQ, 0));
    for (int i = 0; i < numActive; i++) {
        float bottom = tBottom(getThreshold(tQ, i));
        if (top < bottom) slicePoint = std::fmin(slicePoint, bottom);
    }
    return slicePoint;
}
static inline void sliceActive(ThresholdQueue* tQ, float slicePoint, int numActive) {  // K.cl:1040-1067
    for (int cursor = 0; cursor < numActive; cursor++) {
        HEADER currentHeader = getHeader(tQ, cursor);
        float4 current = getThreshold(tQ, cursor);
        if (tTop(current) < slicePoint && slicePoint < tBottom(current)) {
            HEADER splitHeader;
            float4 split;
            splitThreshold(&currentHeader, &current, &splitHeader, &split, slicePoint);
            setHeader(tQ, cursor, currentHeader);
            setThreshold(tQ, cursor, current);
            if (tKeep(splitHeader, split)) insertThreshold(tQ, splitHeader, split);
        }
    }
}

}  // namespace oracle
