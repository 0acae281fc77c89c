// kernels_oracle.cpp — TEST INFRASTRUCTURE, NOT PRODUCT.  See kernels_oracle.hpp for the header
// comment (scope, how it is pinned to the reference, who may call this).
//
// Second half of the restatement of Kernels.cl ("K.cl"): curve traversal and threshold
// generation, colour determination, the sweep, the three kernel bodies, and the host loop that
// stands in for OpenCL/CallKernels.hs:88-179 (generateCall) — one logical work-item per
// (tile, column), three passes with the same global scratch layout, OpenMP over the NDRange.
#include "kernels_oracle.hpp"

#include <omp.h>

#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <vector>

namespace oracle {

struct FrameInputs {
    const uint8_t* geometry;            // geometryHeap
    const gudni_shape* shapes;          // shapeHeap of the job
    const gudni_tile* tiles;            // tileHeap of the job
    const float4* substances;
    const uint8_t* pictureData;
    const gudni_picture_use* pictureRefs;
    float4 backgroundColor;
    int bitmapW, bitmapH;
    int computeDepth;
    int maxThresholds;                  // MAXTHRESHOLDS
    int maxShape;                       // MAXSHAPE
};

// ---- lineToThreshold / lineToHeader, K.cl:1131-1150 ----------------------------------------------
static inline float4 lineToThreshold(float2 left, float2 right) {
    return makeThreshold(std::fmin(left.y, right.y), std::fmax(left.y, right.y), left.x, right.x);
}
static inline HEADER lineToHeader(uint32_t shapeBit, float2 left, float2 right) {
    bool positiveSlope = left.y <= right.y;
    bool notVertical = left.x != right.x;
    bool touchingLeftBorder = left.x == LEFTBORDER;
    bool isPersistent = notVertical && touchingLeftBorder;
    return (positiveSlope ? POSITIVE_SLOPE : NEGATIVE_SLOPE) | (isPersistent ? PERSIST : NONPERSIST) | shapeBit;
}

// ---- addThreshold / addLineSegment, K.cl:1155-1220 -----------------------------------------------
static void addThreshold(ThresholdQueue* tQ, const TileState* tileS, HEADER newHeader, float4 newThreshold,
                         bool* thresholdWasAdded, bool* enclosedByStrand) {
    const float RENDERSTART = 0.0f, RENDEREND = tileS->floatHeight;
    if (tKeep(newHeader, newThreshold)) {
        *enclosedByStrand = *enclosedByStrand ||
                            ((tTop(newThreshold) <= RENDERSTART) && headerPersistTop(newHeader)) ||
                            ((tBottom(newThreshold) <= RENDERSTART) && headerPersistBottom(newHeader));
        if ((tTop(newThreshold) < RENDEREND) && (tBottom(newThreshold) > RENDERSTART) && tLeft(newThreshold) < RIGHTBORDER) {
            if (tTop(newThreshold) <= RENDERSTART) trimThresholdTop(&newHeader, &newThreshold, RENDERSTART);
            if (tRight(newThreshold) <= LEFTBORDER) {
                *enclosedByStrand = true;
            } else {
                *thresholdWasAdded = true;
                pushThreshold(tQ, newHeader, newThreshold);
            }
        }
    }
}
static void addLineSegment(ThresholdQueue* tQ, const TileState* tileS, float2 left, float2 right, uint32_t shapeBit,
                           bool* thresholdWasAdded, bool* enclosedByStrand) {
    addThreshold(tQ, tileS, lineToHeader(shapeBit, left, right), lineToThreshold(left, right), thresholdWasAdded,
                 enclosedByStrand);
}

// ---- curve bisection, K.cl:1226-1258 -------------------------------------------------------------
static void bifurcateCurve(Traversal* t) {
    float2 left = {t->travLeftControl.x, t->travLeftControl.y};
    float2 control = {t->travLeftControl.z, t->travLeftControl.w};
    float2 right = t->travRight;
    if (!(left.x == control.x && left.y == control.y)) {
        for (;;) {
            float2 leftMid = midPoint(left, control);
            float2 rightMid = midPoint(control, right);
            float2 onCurve = midPoint(leftMid, rightMid);
            float flatness = taxiDistance(control, onCurve);
            if (flatness > TAXICAB_FLATNESS) {
                if (t->travXPos < onCurve.x) {
                    control = leftMid;
                    right = onCurve;
                } else {
                    left = onCurve;
                    control = rightMid;
                }
            } else {
                break;
            }
        }
    }
    t->travLeftControl = {left.x, left.y, control.x, control.y};
    t->travRight = right;
}
static float intersectCurve(Traversal t) {  // by value: K.cl:1226-1230
    bifurcateCurve(&t);
    float2 left = {t.travLeftControl.x, t.travLeftControl.y};
    return (left.x == t.travRight.x) ? std::fmin(left.y, t.travRight.y) : yIntercept(left, t.travRight, t.travXPos);
}

// ---- spawnThresholds, K.cl:1264-1333 -------------------------------------------------------------
static void spawnThresholds(ThresholdQueue* tQ, const TileState* tileS, uint32_t shapeBit, Traversal* l, Traversal* r,
                            bool* thresholdWasAdded, bool* enclosedByStrand) {
    const float lLeftX = l->travLeftControl.x, lLeftY = l->travLeftControl.y;
    const float rLeftX = r->travLeftControl.x, rLeftY = r->travLeftControl.y;
    float y_L = (lLeftX >= LEFTBORDER) ? lLeftY : intersectCurve(*l);
    bool leftWing;
    if ((l->travRight.x < RIGHTBORDER) && (l->travRight.x > LEFTBORDER)) {
        addLineSegment(tQ, tileS, {l->travXPos, y_L}, l->travRight, shapeBit, thresholdWasAdded, enclosedByStrand);
        leftWing = true;
    } else {
        leftWing = false;
    }
    float y_R = (r->travRight.x <= RIGHTBORDER) ? r->travRight.y : intersectCurve(*r);
    bool rightWing;
    if ((rLeftX > LEFTBORDER) && (rLeftX < RIGHTBORDER) && (l->travIndex != r->travIndex)) {
        addLineSegment(tQ, tileS, {rLeftX, rLeftY}, {r->travXPos, y_R}, shapeBit, thresholdWasAdded, enclosedByStrand);
        rightWing = true;
    } else {
        rightWing = false;
    }
    if (l->travRight.x < rLeftX || (!leftWing && !rightWing)) {
        float2 bridge_L = (leftWing || (lLeftX == l->travRight.x)) ? l->travRight : float2{l->travXPos, y_L};
        float2 bridge_R = (rightWing || (rLeftX == r->travRight.x)) ? float2{rLeftX, rLeftY} : float2{r->travXPos, y_R};
        addLineSegment(tQ, tileS, bridge_L, bridge_R, shapeBit, thresholdWasAdded, enclosedByStrand);
    }
}

// ---- searchTree / traverseTree, K.cl:1337-1408 ---------------------------------------------------
static inline float4 loadF4(const uint8_t* p) { float4 v; std::memcpy(&v, p, 16); return v; }
static inline float2 loadF2(const uint8_t* p) { float2 v; std::memcpy(&v, p, 8); return v; }

static void searchTree(Traversal* trav, const uint8_t* tree, int treeSize, float4 threadDelta4, bool isLeft) {
    trav->travIndex = 0;
    while (trav->travIndex < treeSize) {
        float4 n = loadF4(tree + 16 * (size_t)trav->travIndex);
        float4 currentTree = {n.x - threadDelta4.x, n.y - threadDelta4.y, n.z - threadDelta4.z, n.w - threadDelta4.w};
        if ((trav->travXPos < currentTree.x) || (isLeft && trav->travXPos == currentTree.x)) {
            trav->travRight = {currentTree.x, currentTree.y};
            trav->travIndex = (trav->travIndex << 1) + 1;
        } else {
            trav->travLeftControl = currentTree;
            trav->travIndex = (trav->travIndex << 1) + 2;
        }
    }
}
static bool traverseTree(const uint8_t* strandHeap /* float2 units */, int currentSize, float2 threadDelta,
                         Traversal* l, Traversal* r) {
    int treeSize = (currentSize - 4) / 2;
    float4 threadDelta4 = {threadDelta.x, threadDelta.y, threadDelta.x, threadDelta.y};
    float2 right = loadF2(strandHeap + 8);
    l->travRight = {right.x - threadDelta.x, right.y - threadDelta.y};
    float4 lc = loadF4(strandHeap + 16);
    l->travLeftControl = {lc.x - threadDelta4.x, lc.y - threadDelta4.y, lc.z - threadDelta4.z, lc.w - threadDelta4.w};
    const uint8_t* tree = strandHeap + 32;
    bool inRange = (l->travLeftControl.x <= RIGHTBORDER && l->travRight.x > LEFTBORDER);  // checkInRange :1360
    if (inRange) {
        *r = *l;
        l->travXPos = std::fmax(LEFTBORDER, l->travLeftControl.x);
        r->travXPos = std::fmin(RIGHTBORDER, l->travRight.x);
        searchTree(l, tree, treeSize, threadDelta4, true);
        searchTree(r, tree, treeSize, threadDelta4, false);
    }
    return inRange;
}

// ---- buildThresholdArray, K.cl:1540-1595 ---------------------------------------------------------
static void buildThresholdArray(const FrameInputs& in, const TileState* tileS, ThresholdQueue* tQ, ShapeState* shS,
                                uint32_t shapeStart, uint32_t numShapes, float2 threadDelta) {
    for (uint32_t n = 0; n < numShapes && shS->shapeBits < (uint32_t)in.maxShape; n++) {
        uint32_t shapeIndex = shapeStart + n;
        bool thresholdWasAdded = false;
        gudni_shape shape = in.shapes[shapeIndex];
        const uint8_t* strandHeap = in.geometry + 16 * (size_t)shape.geo_start;
        bool enclosedByShape = false;
        for (uint32_t currentStrand = 0; currentStrand < shape.num_strands; currentStrand++) {
            uint16_t currentSize;
            std::memcpy(&currentSize, strandHeap, 2);
            bool enclosedByStrand = false;
            Traversal left, right;
            bool inRange = traverseTree(strandHeap, currentSize, threadDelta, &left, &right);
            if (inRange) spawnThresholds(tQ, tileS, shS->shapeBits, &left, &right, &thresholdWasAdded, &enclosedByStrand);
            strandHeap += 8 * (size_t)currentSize;
            enclosedByShape = enclosedByShape != enclosedByStrand;
        }
        if (enclosedByShape) passHeader(shS, shS->shapeBits);
        if (thresholdWasAdded || enclosedByShape) {
            if (shS->shapeBits < (uint32_t)in.maxShape) shS->shapeIndices[shS->shapeBits] = shapeIndex;
            shS->shapeBits += 1;
        }
    }
}

// ---- initTileState / isActiveThread, K.cl:1690-1722 ----------------------------------------------
static void initTileState(TileState* tileS, const gudni_tile* tileInfo, int bitmapW, int bitmapH, int column,
                          int computeDepth) {
    tileS->tileShapeStart = tileInfo->shape_start;
    tileS->tileNumShapes = (int)tileInfo->shape_count;
    tileS->bitmapW = bitmapW;
    tileS->bitmapH = bitmapH;
    tileS->threadId = tileInfo->column_allocation + column;
    tileS->column = column;
    int hDepth = (int)tileInfo->h_depth;
    int vDepth = (int)tileInfo->v_depth;
    int diffDepth = std::max(0, vDepth - (computeDepth - hDepth));
    int desiredHeight = 1 << diffDepth;
    int internalX = ((1 << hDepth) - 1) & column;
    int internalY = (column >> hDepth) << diffDepth;
    tileS->internalDeltaX = internalX;
    tileS->internalDeltaY = internalY;
    tileS->threadDeltaX = internalX + tileInfo->left;
    tileS->threadDeltaY = internalY + tileInfo->top;
    tileS->intHeight = std::min(desiredHeight, bitmapH - tileS->threadDeltaY);
    tileS->floatHeight = (float)tileS->intHeight;
    tileS->tileSizeX = tileInfo->right - tileInfo->left;
    tileS->tileSizeY = tileInfo->bottom - tileInfo->top;
}
static bool isActiveThread(const TileState* tileS) {
    return (tileS->internalDeltaY < tileS->tileSizeY) && (tileS->threadDeltaX < tileS->bitmapW) &&
           (tileS->threadDeltaY < tileS->bitmapH);
}

// ---- colour, K.cl:852-860, 1411-1513 -------------------------------------------------------------
struct ColorState {  // K.cl:357-363
    float4 csBackgroundColor;
    const uint8_t* csPictureData;
    const gudni_picture_use* csPictureRefs;
    int absX, absY;
};
static float4 getPicturePixel(const uint8_t* pictData, int w, int x, int y) {
    const uint8_t* p = pictData + 4 * ((size_t)y * w + x);
    return {(float)p[0] / MAXCHANNELFLOAT, (float)p[1] / MAXCHANNELFLOAT, (float)p[2] / MAXCHANNELFLOAT,
            (float)p[3] / MAXCHANNELFLOAT};
}
static float4 readColor(const ColorState* cS, const float4* substances, uint64_t substanceId, bool isSolidColor) {
    float4 substance = substances[substanceId];
    if (isSolidColor) return substance;
    uint32_t pictId;
    std::memcpy(&pictId, &substance.x, 4);
    gudni_picture_use pRef = cS->csPictureRefs[pictId];
    float scale = pRef.scale;
    scale = scale < 0.0000001f ? 0.0000001f : scale;
    int rx = (int)(((float)cS->absX / scale) - pRef.translate_x);  // convert_int2: round toward zero
    int ry = (int)(((float)cS->absY / scale) - pRef.translate_y);
    if (rx >= 0 && ry >= 0 && rx < pRef.width && ry < pRef.height)
        return getPicturePixel(cS->csPictureData + pRef.mem_offset, pRef.width, rx, ry);
    return {0, 0, 0, 0};
}
static float4 determineColor(const FrameInputs& in, const ShapeState* shS, const ColorState* cS) {  // K.cl:1447-1513
    int topBit = in.maxShape;
    float4 baseColor = {0, 0, 0, 0};
    float4 nextColor = {0, 0, 0, 0};
    bool done = false;
    uint64_t lastId = 0xFFFFFFFFFFFFFFFFull;
    bool lastIsContinue = true;
    bool lastIsSet = false;
    while (!done) {
        uint64_t substanceId = 0xFFFFFFFFFFFFFFFFull;
        bool shouldComposite = true;
        topBit = findTop(shS->shapeStack, topBit);
        if (topBit < 0) {
            nextColor = cS->csBackgroundColor;
            done = true;
            shouldComposite = true;
            lastIsSet = true;
        } else {
            int referenceFromBit = (int)shS->shapeIndices[topBit];
            uint64_t tag = in.shapes[referenceFromBit].tag;
            substanceId = tag & GUDNI_TAG_SUBSTANCEID_MASK;
            bool isContinue = (tag & GUDNI_TAG_COMPOUND_MASK) == GUDNI_TAG_COMPOUND_CONTINUE;
            bool isAdd = (tag & GUDNI_TAG_COMPOUND_MASK) == GUDNI_TAG_COMPOUND_ADD;
            if (substanceId == lastId) {
                if (lastIsContinue) {
                    if (!isContinue) {
                        lastIsSet = isAdd;
                        lastIsContinue = false;
                    } else {
                        lastIsSet = !lastIsSet;
                    }
                }
                shouldComposite = false;
            }
            if (substanceId != lastId) {
                nextColor = readColor(cS, in.substances, substanceId,
                                      (tag & GUDNI_TAG_SUBSTANCETYPE_MASK) == GUDNI_TAG_SUBSTANCE_SOLID);
                shouldComposite = true;
                lastIsSet = isAdd || isContinue;
            }
            lastId = substanceId;
        }
        if (shouldComposite && lastIsSet) {
            baseColor = composite(baseColor, nextColor);
            if (baseColor.w == 1.0f) done = true;
        }
    }
    return baseColor;
}

// ---- the sweep, K.cl:1744-1916 -------------------------------------------------------------------
struct ParseState {  // K.cl:423-436 (debug counters and the inert random field dropped)
    int currentThreshold;
    int numActive;
    float2 sectionStart, sectionEnd;
    float pixelY;
    float acc[8];
};
static float splitNext(ThresholdQueue* tQ, ParseState* pS) {  // K.cl:1069-1077
    float slicePoint = FLT_MAX;
    pS->numActive = countActive(tQ, &slicePoint);
    slicePoint = std::fmin(slicePoint, nextSlicePoint(tQ, slicePoint, pS->numActive));
    sliceActive(tQ, slicePoint, pS->numActive);
    return slicePoint;
}
static void verticalAdvance(ThresholdQueue* tQ, const TileState* tileS, ParseState* pS, ShapeState* shS) {
    if (pS->sectionEnd.x == RIGHTBORDER) {
        for (int i = 0; i < pS->numActive; i++) passHeader(shS, getHeader(tQ, i));
        float nextBreak = std::fmin(tileS->floatHeight, pS->pixelY);
        float activeBottom = tQ->qSlice.sLength > 0 ? tBottom(getThreshold(tQ, 0)) : FLT_MAX;
        if (activeBottom == pS->sectionEnd.y) {
            while (pS->numActive > 0) {
                passHeaderBottom(shS, getHeader(tQ, 0));
                popTop(tQ);
                pS->numActive -= 1;
            }
        }
        float nextBottom;
        if (pS->numActive > 0) {
            nextBottom = std::fmin(activeBottom, nextBreak);
        } else {
            float nextTop = pS->numActive < tQ->qSlice.sLength ? tTop(getThreshold(tQ, pS->numActive)) : FLT_MAX;
            if (nextTop > pS->sectionEnd.y) {
                nextBottom = std::fmin(nextBreak, nextTop);
            } else {
                nextBottom = std::fmin(nextBreak, splitNext(tQ, pS));
                while (pS->numActive > 0 && tIsHorizontal(getThreshold(tQ, 0))) {
                    passHeaderTop(shS, getHeader(tQ, 0));
                    popTop(tQ);
                    pS->numActive -= 1;
                }
                for (int i = 0; i < pS->numActive; i++)
                    if (tTop(getThreshold(tQ, i)) > 0.0f) passHeaderTop(shS, getHeader(tQ, i));
            }
        }
        pS->sectionStart.y = pS->sectionEnd.y;
        pS->sectionEnd.y = nextBottom;
        pS->sectionStart.x = pS->sectionEnd.x = LEFTBORDER;
        pS->currentThreshold = 0;
    }
}
static void horizontalAdvance(ThresholdQueue* tQ, ParseState* pS) {  // K.cl:1826-1851
    float nextX;
    if (pS->currentThreshold < pS->numActive)
        nextX = thresholdMidXLow(getThreshold(tQ, pS->currentThreshold), getHeader(tQ, pS->currentThreshold),
                                 pS->sectionStart.y, pS->sectionEnd.y, LEFTBORDER, RIGHTBORDER);
    else
        nextX = RIGHTBORDER;
    pS->sectionStart.x = pS->sectionEnd.x;
    pS->sectionEnd.x = nextX;
}
static void calculatePixel(const FrameInputs& in, const TileState* tileS, ThresholdQueue* tQ, ShapeState* shS,
                           ParseState* pS, ColorState* cS) {  // K.cl:1881-1916
    while ((pS->sectionEnd.x < RIGHTBORDER) || (pS->sectionEnd.y < pS->pixelY)) {
        verticalAdvance(tQ, tileS, pS, shS);
        horizontalAdvance(tQ, pS);
        // sectionColor, K.cl:1724-1742 with STOCHASTIC_FACTOR = 0: adjustedArea = area + area*random*0
        float4 color = determineColor(in, shS, cS);
        float area = (pS->sectionEnd.x - pS->sectionStart.x) * (pS->sectionEnd.y - pS->sectionStart.y);
        pS->acc[0] += color.x * area;
        pS->acc[1] += color.y * area;
        pS->acc[2] += color.z * area;
        pS->acc[3] += color.w * area;
        pS->acc[4] += area; pS->acc[5] += area; pS->acc[6] += area; pS->acc[7] += area;
        if (pS->currentThreshold < pS->numActive) passHeader(shS, getHeader(tQ, pS->currentThreshold));
        pS->currentThreshold += 1;
    }
}
// convert_uchar4 (K.cl:843): round toward zero.  Out-of-range input (the signed section areas make a
// channel slightly negative or above 255 on a few pixels per million) is undefined in OpenCL C without
// _sat; this follows what the reference's kernels do on the one real OpenCL device they could be run on
// (NVIDIA OpenCL 3.0 on the B200, profiles/r1_opencl_reference.json): clamp to [0,255], NaN -> 0.
static inline uint32_t toByte(float v) { return v >= 255.0f ? 255u : (v > 0.0f ? (uint32_t)(int32_t)v : 0u); }
static void writePixelGlobal(const TileState* tileS, float4 color, uint32_t* out, int y) {  // K.cl:842-844, 1853-1862
    uint32_t word = toByte(color.z * MAXCHANNELFLOAT) | (toByte(color.y * MAXCHANNELFLOAT) << 8) |
                    (toByte(color.x * MAXCHANNELFLOAT) << 16) | (toByte(1.0f * MAXCHANNELFLOAT) << 24);
    size_t outPos = (size_t)(tileS->threadDeltaY + y) * tileS->bitmapW + tileS->threadDeltaX;
    out[outPos] = word;
}
static void renderThresholdArray(const FrameInputs& in, const TileState* tileS, ThresholdQueue* tQ, ShapeState* shS,
                                 uint32_t* out) {  // K.cl:1978-2028
    ParseState pS;
    pS.currentThreshold = 0;
    pS.numActive = 0;
    for (int i = 0; i < 8; i++) pS.acc[i] = 0.0f;
    pS.sectionStart = {LEFTBORDER, 0.0f};
    pS.sectionEnd = {RIGHTBORDER, 0.0f};
    ColorState cS = {in.backgroundColor, in.pictureData, in.pictureRefs, tileS->threadDeltaX, tileS->threadDeltaY};
    int yInt = -1;
    for (pS.pixelY = 1.0f; pS.pixelY <= tileS->floatHeight; pS.pixelY += 1.0f) {
        yInt += 1;
        calculatePixel(in, tileS, tQ, shS, &pS, &cS);
        if (tQ->overflow) return;
        float4 color = {pS.acc[0] / pS.acc[4], pS.acc[1] / pS.acc[5], pS.acc[2] / pS.acc[6], pS.acc[3] / pS.acc[7]};
        writePixelGlobal(tileS, color, out, yInt);
        for (int i = 0; i < 8; i++) pS.acc[i] = 0.0f;
        pS.sectionStart = {LEFTBORDER, pS.pixelY};
        cS.absY += 1;
    }
}

// ---- sortThresholdArray, K.cl:1932-1976 (bubble sort, stable) ------------------------------------
static bool swapIfAbove(ThresholdQueue* tQ, int i, bool done) {
    HEADER aHeader = getHeader(tQ, i);
    float4 a = getThreshold(tQ, i);
    HEADER bHeader = getHeader(tQ, i + 1);
    float4 b = getThreshold(tQ, i + 1);
    if (thresholdIsBelow(aHeader, a, bHeader, b)) {  // same predicate, K.cl:1940-1951
        setHeader(tQ, i, bHeader);
        setThreshold(tQ, i, b);
        setHeader(tQ, i + 1, aHeader);
        setThreshold(tQ, i + 1, a);
        done = false;
    }
    return done;
}
static void sortThresholdArray(ThresholdQueue* tQ) {
    bool done = false;
    int k = tQ->qSlice.sLength;
    while (!done) {
        done = true;
        for (int i = 0; i < k - 1; i++) done = swapIfAbove(tQ, i, done);
        k--;
    }
}

// ---- the job: scratch layout of generateCall (OpenCL/CallKernels.hs:124-127) ---------------------
struct JobScratch {
    std::vector<float4> thresholdHeap;
    std::vector<HEADER> headerHeap;
    std::vector<ShapeState> shapeStateHeap;
    std::vector<Slice> qSliceHeap;
    std::vector<uint8_t> overflow;
    void ensure(size_t columns, int maxThresholds) {
        size_t n = columns * (size_t)maxThresholds;
        if (thresholdHeap.size() < n) { thresholdHeap.resize(n); headerHeap.resize(n); }
        if (shapeStateHeap.size() < columns) { shapeStateHeap.resize(columns); qSliceHeap.resize(columns); }
        overflow.assign(columns, 0);
    }
};

static void initThresholdQueue(ThresholdQueue* tQ, const TileState* tileS, JobScratch& s, int maxThresholds, Slice q) {
    tQ->thresholdHeaders = s.headerHeap.data() + (size_t)tileS->threadId * maxThresholds;  // K.cl:1622-1631
    tQ->thresholds = s.thresholdHeap.data() + (size_t)tileS->threadId * maxThresholds;
    tQ->qSlice = q;
    tQ->capacity = maxThresholds;
    tQ->overflow = false;
}

// generateThresholds, K.cl:2030-2082
static void kernelGenerate(const FrameInputs& in, JobScratch& s, int tileIndex, int column) {
    TileState tileS;
    initTileState(&tileS, &in.tiles[tileIndex], in.bitmapW, in.bitmapH, column, in.computeDepth);
    if (!isActiveThread(&tileS)) return;
    ThresholdQueue tQ;
    initThresholdQueue(&tQ, &tileS, s, in.maxThresholds, Slice{in.maxThresholds, 0});
    ShapeState& shS = s.shapeStateHeap[tileS.threadId];
    shS.shapeBits = 0;
    for (int i = 0; i < SHAPESTACKSECTIONS; i++) shS.shapeStack[i] = 0;
    buildThresholdArray(in, &tileS, &tQ, &shS, tileS.tileShapeStart, (uint32_t)tileS.tileNumShapes,
                        float2{(float)tileS.threadDeltaX, (float)tileS.threadDeltaY});
    s.qSliceHeap[tileS.threadId] = tQ.qSlice;
    if (tQ.overflow) s.overflow[tileS.threadId] = 1;
}
// sortThresholds, K.cl:2084-2115
static void kernelSort(const FrameInputs& in, JobScratch& s, int tileIndex, int column) {
    TileState tileS;
    initTileState(&tileS, &in.tiles[tileIndex], in.bitmapW, in.bitmapH, column, in.computeDepth);
    if (!isActiveThread(&tileS)) return;
    if (s.overflow[tileS.threadId]) return;
    ThresholdQueue tQ;
    initThresholdQueue(&tQ, &tileS, s, in.maxThresholds, s.qSliceHeap[tileS.threadId]);
    sortThresholdArray(&tQ);
}
// renderThresholds, K.cl:2117-2167
static void kernelRender(const FrameInputs& in, JobScratch& s, int tileIndex, int column, uint32_t* out) {
    TileState tileS;
    initTileState(&tileS, &in.tiles[tileIndex], in.bitmapW, in.bitmapH, column, in.computeDepth);
    if (!isActiveThread(&tileS)) return;
    if (s.overflow[tileS.threadId]) return;
    ThresholdQueue tQ;
    initThresholdQueue(&tQ, &tileS, s, in.maxThresholds, s.qSliceHeap[tileS.threadId]);
    ShapeState shS = s.shapeStateHeap[tileS.threadId];
    renderThresholdArray(in, &tileS, &tQ, &shS, out);
    if (tQ.overflow) s.overflow[tileS.threadId] = 1;
}

}  // namespace oracle

// =================================================================================================
// C API (ctypes).  One call = one RasterJob = generateCall (OpenCL/CallKernels.hs:114-179).
// =================================================================================================
extern "C" {

static oracle::JobScratch g_scratch;  // reused across jobs (the reference allocates per job)

int gudni_oracle_threads(void) { return omp_get_max_threads(); }
void gudni_oracle_set_threads(int n) { if (n > 0) omp_set_num_threads(n); }

// Rasterizes one job into `out` (bitmapW*bitmapH words).  Optional taps, indexed by
// threadId = column_allocation + column: n_thresholds (qSlice.sLength after generate, -1 for
// inactive threads), shape_bits (ShapeState.shapeBits).  Returns the number of threads that
// overflowed max_thresholds (their pixels are not written).
int64_t gudni_oracle_raster_job(const void* geometry, const float* substances, const uint8_t* picture_bytes,
                                const gudni_picture_use* picture_uses, const float* background_rgba, int width,
                                int height, const gudni_spec* spec, const gudni_shape* shapes, const gudni_tile* tiles,
                                int n_tiles, int columns_allocated, uint32_t* out, int32_t* n_thresholds,
                                int32_t* shape_bits, int64_t* total_thresholds) {
    using namespace oracle;
    FrameInputs in;
    in.geometry = static_cast<const uint8_t*>(geometry);
    in.shapes = shapes;
    in.tiles = tiles;
    in.substances = reinterpret_cast<const float4*>(substances);
    in.pictureData = picture_bytes;
    in.pictureRefs = picture_uses;
    in.backgroundColor = {background_rgba[0], background_rgba[1], background_rgba[2], background_rgba[3]};
    in.bitmapW = width;
    in.bitmapH = height;
    int threadsPerTile = spec->threads_per_tile;
    int computeDepth = 0;
    while ((1 << computeDepth) < threadsPerTile) computeDepth++;  // adjustedLog, Raster/TileTree.hs:74-75
    in.computeDepth = computeDepth;
    in.maxThresholds = spec->max_thresholds;
    in.maxShape = spec->max_shapes;
    g_scratch.ensure((size_t)columns_allocated, spec->max_thresholds);
    if (n_thresholds) for (int i = 0; i < columns_allocated; i++) n_thresholds[i] = -1;
    if (shape_bits) for (int i = 0; i < columns_allocated; i++) shape_bits[i] = -1;
    const long total = (long)n_tiles * threadsPerTile;
#pragma omp parallel for schedule(dynamic, 64)
    for (long g = 0; g < total; g++) kernelGenerate(in, g_scratch, (int)(g / threadsPerTile), (int)(g % threadsPerTile));
    int64_t sum = 0;
    if (n_thresholds || shape_bits || total_thresholds) {
        for (long g = 0; g < total; g++) {
            TileState tileS;
            initTileState(&tileS, &tiles[g / threadsPerTile], width, height, (int)(g % threadsPerTile), computeDepth);
            if (!isActiveThread(&tileS)) continue;
            sum += g_scratch.qSliceHeap[tileS.threadId].sLength;
            if (n_thresholds) n_thresholds[tileS.threadId] = g_scratch.qSliceHeap[tileS.threadId].sLength;
            if (shape_bits) shape_bits[tileS.threadId] = (int32_t)g_scratch.shapeStateHeap[tileS.threadId].shapeBits;
        }
    }
    if (total_thresholds) *total_thresholds = sum;
#pragma omp parallel for schedule(dynamic, 64)
    for (long g = 0; g < total; g++) kernelSort(in, g_scratch, (int)(g / threadsPerTile), (int)(g % threadsPerTile));
#pragma omp parallel for schedule(dynamic, 64)
    for (long g = 0; g < total; g++) kernelRender(in, g_scratch, (int)(g / threadsPerTile), (int)(g % threadsPerTile), out);
    int64_t overflowed = 0;
    for (int i = 0; i < columns_allocated; i++) overflowed += g_scratch.overflow[i];
    return overflowed;
}

void gudni_oracle_release_scratch(void) { g_scratch = oracle::JobScratch(); }

}  // extern "C"
