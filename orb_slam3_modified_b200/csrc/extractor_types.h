// Internal layout of the batched extractor (not part of the C-ABI).
#pragma once
#include <stdint.h>

namespace orbx {

constexpr int kMaxLevels = 16;
constexpr int EDGE_THRESHOLD = 19;   // reference src/ORBextractor.cc:73
constexpr int HALF_PATCH = 15;       // :72
constexpr int PATCH_SIZE = 31;       // :71

// Geometry of one pyramid level for the currently configured image size.
struct LevelGeom {
    int w, h, pitch;          // plane size; pitch is a multiple of 32 bytes
    uint32_t planeOff;        // byte offset of the plane inside one frame's pyramid buffer
    int minBX, minBY, maxBX, maxBY;   // detection window (reference :789-792)
    int nCols, nRows, wCell, hCell;   // FAST cell grid (:797-803)
    int cellBase, nCells;     // active cells of this level inside the frame's cell table
    int cellCap;              // list capacity of one cell (max NMS survivors)
    uint32_t candOff, candCap;        // candidate scratch of this level inside one frame
    int nDesired;             // mnFeaturesPerLevel[level]
    int selOff, selCap;       // selected-keypoint slots of this level inside one frame
    int nIni; float hX;       // quadtree roots (:559-561)
    float scale; float sizeScaled;    // mvScaleFactor[level], (float)(int)(31*scale)
    int area2x;               // level is an exact 2x decimation of the previous (cv::resize -> INTER_AREA)
    uint32_t xtabOff, ytabOff;        // resize coefficient tables (entries, not bytes)
    int blurTileBase, blurTilesX, blurTilesY;   // flattened tile index range of the blur launch
    int fastBoxW, fastBoxH;   // TMA box of the FAST cell ROI: width (multiple of 16 bytes) x height
    int pyrBoxW, pyrBoxH;     // TMA box (in level l-1) that covers the sources of one 128 x 8 destination tile of this level
};

// One active FAST cell (cells skipped by the reference's `continue`s are not listed).
struct CellDesc {
    uint16_t level;
    uint16_t x0, y0;          // ROI origin in level coordinates
    uint16_t rw, rh;          // ROI size (detection interior is [3,rw-3) x [3,rh-3))
    uint16_t offX, offY;      // j*wCell, i*hCell: added to ROI-relative corner coordinates (:866-867)
    uint16_t pad;
    uint32_t listOff;         // offset of this cell's list inside one frame's cell-list buffer
};

// Selected keypoint before output assembly.
struct SelKp {
    int16_t x, y;             // level coordinates (border offset already added)
    int32_t response;
    float angle;
};

struct ExtractParams {
    int nlevels, batch;
    int iniTh, minTh;
    int lap0, lap1;
    LevelGeom lv[kMaxLevels];
    // device pointers (per-frame strides in elements of the pointed type)
    const uint8_t* src; size_t srcStep, srcFrameStride; int rows, cols;
    uint8_t* pyr;  uint8_t* blur; size_t pyrFrameStride;
    // level 0 may alias the caller's device image instead of a copy inside `pyr`
    const uint8_t* lv0; size_t lv0Pitch, lv0FrameStride;
    int blurTilesTotal;
    const CellDesc* cells; int nCellsTotal;
    int* cellCount;        size_t cellCountStride;
    uint32_t* cellList;    size_t cellListStride;
    uint32_t* cand;        size_t candStride;       // gathered candidates (packed x|y|score)
    uint16_t* nodeOf;                               // same stride as cand
    SelKp* sel;            size_t selStride;
    int* selCount;                                  // [batch][kMaxLevels]
    int* dstIndex;                                  // [batch][selStride] output row of each selected keypoint
    const short4* xtab; const short4* ytab;
    // outputs
    void* outKp; uint8_t* outDesc; int outCap; int* outN; int* outMono;
    int* status;                                    // [batch] sticky error flags
    int maxNodes, maxCellsPerLevel;
    int dbg;
};

// candidate packing: x (12 bits) | y (12 bits) << 12 | score << 24
__host__ __device__ inline uint32_t pack_cand(int x, int y, int s) { return (uint32_t)x | ((uint32_t)y << 12) | ((uint32_t)s << 24); }
__host__ __device__ inline int cand_x(uint32_t c) { return c & 0xFFF; }
__host__ __device__ inline int cand_y(uint32_t c) { return (c >> 12) & 0xFFF; }
__host__ __device__ inline int cand_s(uint32_t c) { return c >> 24; }

}  // namespace orbx
