// Hand-written sm_100a kernels of the batched ORB extractor.  Every kernel carries a
// batch (frame) dimension in blockIdx.y / blockIdx.z so that one launch covers all streams.
//
// Reference semantics (file:line in /root/reference) are cited per kernel; the
// bit-exact scalar pieces live in exact_math.h.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include "exact_math.h"
#include "device_utils.cuh"
#include "extractor_types.h"

namespace orbx {

__constant__ int8_t c_pattern[1024];      // 256 pairs x (x0,y0,x1,y1), reference src/ORBextractor.cc:149-407
__constant__ int c_umax[HALF_PATCH + 1];  // :453-468

// ------------------------------------------------------------------------------------------
// plane addressing: level 0 may alias the caller's image
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ const uint8_t* plane_ptr(const ExtractParams& P, int f, int l, int& pitch) {
    if (l == 0) { pitch = (int)P.lv0Pitch; return P.lv0 + (size_t)f * P.lv0FrameStride; }
    pitch = P.lv[l].pitch;
    return P.pyr + (size_t)f * P.pyrFrameStride + P.lv[l].planeOff;
}
__device__ __forceinline__ const uint8_t* blur_ptr(const ExtractParams& P, int f, int l, int& pitch) {
    pitch = P.lv[l].pitch;
    return P.blur + (size_t)f * P.pyrFrameStride + P.lv[l].planeOff;
}

// ------------------------------------------------------------------------------------------
// K0: copy the caller's image into the level-0 plane (only when level 0 cannot alias it).
// ------------------------------------------------------------------------------------------
__global__ void copy_level0_kernel(ExtractParams P) {
    const int f = blockIdx.z;
    const int y = blockIdx.y;
    const uint8_t* s = P.src + (size_t)f * P.srcFrameStride + (size_t)y * P.srcStep;
    uint8_t* d = P.pyr + (size_t)f * P.pyrFrameStride + P.lv[0].planeOff + (size_t)y * P.lv[0].pitch;
    for (int x = blockIdx.x * blockDim.x + threadIdx.x; x < P.cols; x += gridDim.x * blockDim.x) d[x] = s[x];
}

// ------------------------------------------------------------------------------------------
// K1: pyramid level l from level l-1.  cv::resize(INTER_LINEAR) restated in integer
// fixed point (reference call site src/ORBextractor.cc:1183).  Each thread
// produces 4 horizontally adjacent pixels and stores them as one 32-bit word.
// The 19-px reflected border the reference also writes (:1185-1191) is never read on this
// path (SURVEY.md 8a reach analysis) and is not materialised.
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) pyr_resize_kernel(ExtractParams P, int l) {
    const LevelGeom& G = P.lv[l];
    const int f = blockIdx.z;
    const int dy = blockIdx.y * blockDim.y + threadIdx.y;
    const int dx0 = (blockIdx.x * blockDim.x + threadIdx.x) * 4;
    if (dy >= G.h || dx0 >= G.w) return;
    int spitch;
    const uint8_t* src = plane_ptr(P, f, l - 1, spitch);
    uint8_t* dst = P.pyr + (size_t)f * P.pyrFrameStride + G.planeOff + (size_t)dy * G.pitch;
    uint32_t out = 0;
    if (G.area2x) {
        const uint8_t* r0 = src + (size_t)(2 * dy) * spitch;
        const uint8_t* r1 = r0 + spitch;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const int dx = dx0 + k;
            if (dx < G.w) {
                int v = (r0[2 * dx] + r0[2 * dx + 1] + r1[2 * dx] + r1[2 * dx + 1] + 2) >> 2;
                out |= (uint32_t)v << (8 * k);
            }
        }
    } else {
        const short4 yt = P.ytab[G.ytabOff + dy];  // {sy0, sy1, b0, b1}
        const uint8_t* r0 = src + (size_t)yt.x * spitch;
        const uint8_t* r1 = src + (size_t)yt.y * spitch;
        const int b0 = yt.z, b1 = yt.w;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const int dx = dx0 + k;
            if (dx < G.w) {
                const short4 xt = __ldg(&P.xtab[G.xtabOff + dx]);  // {sx0, sx1, a0, a1}
                const int s0 = r0[xt.x] * xt.z + r0[xt.y] * xt.w;
                const int s1 = r1[xt.x] * xt.z + r1[xt.y] * xt.w;
                const int v = (((b0 * (s0 >> 4)) >> 16) + ((b1 * (s1 >> 4)) >> 16) + 2) >> 2;
                out |= (uint32_t)(v & 0xFF) << (8 * k);
            }
        }
    }
    *reinterpret_cast<uint32_t*>(dst + dx0) = out;   // pitch is a multiple of 32: in-bounds even for the tail
}

// ------------------------------------------------------------------------------------------
// K2: FAST-9/16 per cell with the two-threshold fallback and cell-local NMS.
// Reference: ORBextractor::ComputeKeyPointsOctTree src/ORBextractor.cc:805-873 calling
// cv::FAST(cellROI, th, nonmax=true).  One CTA per (cell, frame).
//   score = max over the 16 nine-pixel arcs of max(min(v-p), min(p-v)) - 1; corner at T <=> score >= T;
//   NMS: strict 8-neighbour maximum, neighbours outside the ROI interior count as 0.
// A strict local maximum of the th=minTh score map with score >= T is exactly what
// cv::FAST(T) keeps, so one score map serves both passes.
// Output: per cell a count and a row-major (y, x) list of packed (x+offX, y+offY, score).
// ------------------------------------------------------------------------------------------
#ifndef ORBX_FAST_NT
#define ORBX_FAST_NT 64    /* measured: 256 -> 1.21 ms, 128 -> 0.93, 64 -> 0.91 per 240 frames (a cell has ~270 pixel quads) */
#endif
constexpr int FAST_NT = ORBX_FAST_NT;   // threads per (cell, frame) CTA

__device__ __forceinline__ int fast_score_at(const uint8_t* t, const int* off) {
    // d_k = v - p_k (darker ring), e_k = p_k - v (brighter ring) for the 16 ring pixels.
    // NOTE: both polarities are written as min-trees over separately subtracted arrays on purpose.
    // The shorter form max(min9(d), -max9(d)) is miscompiled by ptxas 12.9 for sm_100a (the negation is
    // dropped when it is folded into a VIMNMX3 operand); see DESIGN.md "toolchain findings".
    const int v = t[0];
    int d[16], e[16];
#pragma unroll
    for (int k = 0; k < 16; ++k) { const int p = t[off[k]]; d[k] = v - p; e[k] = p - v; }
    // min over every window of 9 consecutive (cyclic) ring positions via two levels of 3-input minima
    int a3[16], b3[16];
#pragma unroll
    for (int k = 0; k < 16; ++k) {
        a3[k] = min(d[k], min(d[(k + 1) & 15], d[(k + 2) & 15]));
        b3[k] = min(e[k], min(e[(k + 1) & 15], e[(k + 2) & 15]));
    }
    int best = -512;
#pragma unroll
    for (int k = 0; k < 16; ++k) {
        const int a9 = min(a3[k], min(a3[(k + 3) & 15], a3[(k + 6) & 15]));
        const int b9 = min(b3[k], min(b3[(k + 3) & 15], b3[(k + 6) & 15]));
        best = max(best, max(a9, b9));
    }
    return best;  // corner at T <=> best > T, score = best - 1
}

// TMA descriptors: one 3-D tensor {x, y, frame} per pyramid level.  Levels >= 1 live in a global-memory table; level 0 (which
// may alias the caller's frames and so changes per call) travels as a __grid_constant__ kernel parameter.
struct alignas(64) TmaMaps { CUtensorMap lv[kMaxLevels]; };

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tWAIT_%=:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra DONE_%=;\n\tbra WAIT_%=;\n\tDONE_%=:\n\t}" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
// cp.async.bulk.tensor (TMA) 3-D tile load: box {bw, bh, 1} at (x, y, frame) -> shared memory, completion on an mbarrier
__device__ __forceinline__ void tma_load_3d(void* dst, const CUtensorMap* map, int x, int y, int z, uint64_t* bar) {
    asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
                 ::"r"(smem_u32(dst)), "l"(map), "r"(x), "r"(y), "r"(z), "r"(smem_u32(bar)) : "memory");
}

__global__ void __launch_bounds__(FAST_NT) fast_cells_kernel(ExtractParams P, const __grid_constant__ CUtensorMap map0, const CUtensorMap* __restrict__ maps) {
    extern __shared__ __align__(128) uint8_t smem_raw[];
    __shared__ __align__(8) uint64_t s_bar;
    const int f = blockIdx.y;
    const CellDesc cd = P.cells[blockIdx.x];
    const int rw = cd.rw, rh = cd.rh;
    int* outCount = P.cellCount + (size_t)f * P.cellCountStride + blockIdx.x;
    if (rw < 7 || rh < 7) { if (threadIdx.x == 0) *outCount = 0; return; }
    const LevelGeom& G = P.lv[cd.level];
    const int tp = G.fastBoxW;                  // tile pitch = TMA box width (multiple of 16)
    const int iw = rw - 6, ih = rh - 6;         // detection interior (<= 70 x 70)
    const int sp = (iw + 2 + 3) & ~3;           // score-map pitch (1-px zero ring), word aligned
    // TMA tile mode needs a 16-byte aligned box start (unaligned inner coordinates fault: measured), so the box
    // starts at x0 & ~15 and the ROI begins `xoff` bytes into every tile row
    const int xoff = cd.x0 & 15;
    uint8_t* tileBase = smem_raw;                                               // fastBoxH * tp, 128-byte aligned
    const uint8_t* tile = tileBase + xoff;
    uint8_t* score = tileBase + ((G.fastBoxH * tp + 127) & ~127);               // (ih+2) * sp
    uint16_t* list = reinterpret_cast<uint16_t*>(score + (((ih + 2) * sp + 15) & ~15));  // iw*ih candidates (y << 7 | x)
    __shared__ int s_nCand, s_cntHi, s_warp[33];
    __shared__ uint32_t s_hi[72 * 3], s_lo[72 * 3];             // per-row keep bitmaps (<= 70 rows x 96 columns)
    __shared__ int s_rowOff[96];
    const int tid = threadIdx.x, lane = tid & 31;

    // ---- tile: one TMA bulk-tensor load of the whole ROI (image pyramid staged into shared memory) ----
    if (tid == 0) {
        mbar_init(&s_bar, 1);
        mbar_expect_tx(&s_bar, (uint32_t)(tp * G.fastBoxH));
        tma_load_3d(tileBase, cd.level == 0 ? &map0 : maps + cd.level, cd.x0 - xoff, cd.y0, f, &s_bar);
        s_nCand = 0; s_cntHi = 0;
    }
    for (int i = tid; i < ((ih + 2) * sp) >> 2; i += FAST_NT) reinterpret_cast<uint32_t*>(score)[i] = 0;
    for (int i = tid; i < 72 * 3; i += FAST_NT) { s_hi[i] = 0; s_lo[i] = 0; }
    __syncthreads();                                            // barrier init visible + score zeroed
    mbar_wait(&s_bar, 0);

    // phase 1: cheap rejection (any 9-arc contains ring pixel k or k+8 for every k) + compaction of the survivors.
    // One thread tests four horizontally adjacent pixels; pixels travel in pairs as 2 x 16-bit lanes so that one
    // VIADD.16x2 / LOP3 serves two pixels: with NLO = T - v and NHI1 = -(v + T + 1) per lane,
    //   p < v - T  <=>  sign(p + NLO) set,      p > v + T  <=>  sign(p + NHI1) clear.
    const int th = P.minTh;
    {
        const int nq = (iw + 3) >> 2, nQ = nq * ih;
        const unsigned qmagic = 0xFFFFFFFFu / (unsigned)nq + 1u;
        const uint32_t TT = (uint32_t)th * 0x10001u, TT1 = (uint32_t)(th + 1) * 0x10001u;
        const int shC = ((xoff) & 3) * 8;            // byte phase of column (4q + xoff) inside an aligned word
        for (int base = 0; base < nQ; base += FAST_NT) {
            const int qi = base + tid;
            uint32_t c01 = 0, c23 = 0;
            int y = 0, q = 0;
            if (qi < nQ) {
                y = (int)__umulhi((unsigned)qi, qmagic); q = qi - y * nq;
                // aligned word address of tile byte (row y+3, column 4q + xoff) = first of the 10 bytes c0-3 .. c0+6, c0 = 4q + 3 + xoff
                const uint32_t* rc = reinterpret_cast<const uint32_t*>(tileBase + (y + 3) * tp + ((4 * q + xoff) & ~3));
                const int tpw = tp >> 2;
                // three 4-byte windows per row: bytes [c0-3, c0], [c0+1, c0+4], [c0+5, c0+8]
#define ORBX_WIN3(r, A, B, C) { const uint32_t w0 = (r)[0], w1 = (r)[1], w2 = (r)[2], w3 = (r)[3]; \
                                A = __funnelshift_r(w0, w1, shC); B = __funnelshift_r(w1, w2, shC); C = __funnelshift_r(w2, w3, shC); }
                uint32_t A0, B0, C0, Au, Bu, Cu, Ad, Bd, Cd;
                ORBX_WIN3(rc, A0, B0, C0);                         // centre row
                ORBX_WIN3(rc + 2 * tpw, Au, Bu, Cu);               // row +2
                ORBX_WIN3(rc - 2 * tpw, Ad, Bd, Cd);               // row -2
                uint32_t t3a, t3b;                                 // rows +-3: bytes c0 .. c0+3 = window offset 3..6
                { const uint32_t* r3 = rc + 3 * tpw; const uint32_t w0 = r3[0], w1 = r3[1], w2 = r3[2];
                  t3a = __funnelshift_r(__funnelshift_r(w0, w1, shC), __funnelshift_r(w1, w2, shC), 24); }
                { const uint32_t* r3 = rc - 3 * tpw; const uint32_t w0 = r3[0], w1 = r3[1], w2 = r3[2];
                  t3b = __funnelshift_r(__funnelshift_r(w0, w1, shC), __funnelshift_r(w1, w2, shC), 24); }
#undef ORBX_WIN3
                // window byte offsets (relative to c0-3): centre = 3..6, p12 (x-3) = 0..3, p4 (x+3) = 6..9, x-2 = 1..4, x+2 = 5..8
                const uint32_t ctr = __funnelshift_r(A0, B0, 24);                 // bytes c0 .. c0+3
                const uint32_t p12 = A0;                                           // bytes c0-3 .. c0
                const uint32_t p4 = __funnelshift_r(B0, C0, 16);                  // bytes c0+3 .. c0+6
                const uint32_t p2 = __funnelshift_r(Bu, Cu, 8), p14 = __funnelshift_r(Au, Bu, 8);    // row +2: x+2 (5..8), x-2 (1..4)
                const uint32_t p6 = __funnelshift_r(Bd, Cd, 8), p10 = __funnelshift_r(Ad, Bd, 8);    // row -2: x+2, x-2
                const uint32_t p0 = t3a, p8 = t3b;
#define ORBX_LO(w) __byte_perm((w), 0u, 0x4140)   /* bytes 0,1 -> two 16-bit lanes */
#define ORBX_HI(w) __byte_perm((w), 0u, 0x4342)   /* bytes 2,3 */
                const uint32_t V01 = ORBX_LO(ctr), V23 = ORBX_HI(ctr);
                const uint32_t NLO01 = __vsub2(TT, V01), NLO23 = __vsub2(TT, V23);                 // T - v
                const uint32_t NHI01 = __vsub2(0u, __vadd2(V01, TT1)), NHI23 = __vsub2(0u, __vadd2(V23, TT1));   // -(v + T + 1)
                uint32_t dk01 = 0xFFFFFFFFu, dk23 = 0xFFFFFFFFu, br01 = 0u, br23 = 0u;
#define ORBX_OPP(pa, pb) { \
                    const uint32_t a01 = ORBX_LO(pa), a23 = ORBX_HI(pa), b01 = ORBX_LO(pb), b23 = ORBX_HI(pb); \
                    dk01 &= __vadd2(a01, NLO01) | __vadd2(b01, NLO01); dk23 &= __vadd2(a23, NLO23) | __vadd2(b23, NLO23); \
                    br01 |= __vadd2(a01, NHI01) & __vadd2(b01, NHI01); br23 |= __vadd2(a23, NHI23) & __vadd2(b23, NHI23); }
                ORBX_OPP(p0, p8); ORBX_OPP(p4, p12); ORBX_OPP(p2, p10); ORBX_OPP(p6, p14);
#undef ORBX_OPP
#undef ORBX_LO
#undef ORBX_HI
                c01 = (dk01 | ~br01) & 0x80008000u;
                c23 = (dk23 | ~br23) & 0x80008000u;
                const int xr = iw - 4 * q;     // pixels of this quad inside the interior
                if (xr < 4) { c23 = xr <= 2 ? 0u : (c23 & 0x8000u); if (xr < 2) c01 &= 0x8000u; }
            }
            // compaction: up to four candidates per lane, in (y, x) packed form
            const int n4 = ((c01 >> 15) & 1) + (c01 >> 31) + ((c23 >> 15) & 1) + (c23 >> 31);
            int incl = n4;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) { const int t = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += t; }
            const int wtot = __shfl_sync(0xffffffffu, incl, 31);
            if (wtot) {
                int wbase = 0;
                if (lane == 31) wbase = atomicAdd(&s_nCand, wtot);
                wbase = __shfl_sync(0xffffffffu, wbase, 31);
                int o = wbase + incl - n4;
                const int yx = (y << 7) | (4 * q);
                if (c01 & 0x8000u) list[o++] = (uint16_t)yx;
                if (c01 & 0x80000000u) list[o++] = (uint16_t)(yx + 1);
                if (c23 & 0x8000u) list[o++] = (uint16_t)(yx + 2);
                if (c23 & 0x80000000u) list[o++] = (uint16_t)(yx + 3);
            }
        }
    }
    __syncthreads();

    // phase 2: exact score for the survivors
    {
        int off[16];
        const int dxs[16] = {0, 1, 2, 3, 3, 3, 2, 1, 0, -1, -2, -3, -3, -3, -2, -1};
        const int dys[16] = {3, 3, 2, 1, 0, -1, -2, -3, -3, -3, -2, -1, 0, 1, 2, 3};
#pragma unroll
        for (int k = 0; k < 16; ++k) off[k] = dys[k] * tp + dxs[k];
        const int nc = s_nCand;
        for (int c = tid; c < nc; c += FAST_NT) {
            const int yx = list[c];
            const int y = yx >> 7, x = yx & 127;
            const int b = fast_score_at(tile + (y + 3) * tp + (x + 3), off);
            if (b > th) score[(y + 1) * sp + (x + 1)] = (uint8_t)(b - 1);
        }
    }
    __syncthreads();

    // phase 3: strict 8-neighbour maxima among the scored candidates only (corners are sparse) -> per-row keep bitmaps
    // for both thresholds via shared-memory atomicOr
    {
        const int nc = s_nCand;
        int cntHi = 0;
        for (int c = tid; c < nc; c += FAST_NT) {
            const int yx = list[c];
            const int y = yx >> 7, x = yx & 127;
            const uint8_t* sc = score + (y + 1) * sp + (x + 1);
            const int v = sc[0];
            if (v == 0) continue;
            const bool keep = v > sc[-1] && v > sc[1] && v > sc[-sp - 1] && v > sc[-sp] && v > sc[-sp + 1] &&
                              v > sc[sp - 1] && v > sc[sp] && v > sc[sp + 1];          // score >= minTh by construction
            if (keep) {
                atomicOr(&s_lo[y * 3 + (x >> 5)], 1u << (x & 31));
                if (v >= P.iniTh) { atomicOr(&s_hi[y * 3 + (x >> 5)], 1u << (x & 31)); ++cntHi; }
            }
        }
        if (cntHi) atomicAdd(&s_cntHi, cntHi);
    }
    __syncthreads();
    // fallback to minTh only if the cell is empty at iniTh (:843-859); ordered (row-major) emission, one thread per row
    const uint32_t* bits = s_cntHi > 0 ? s_hi : s_lo;
    const int nwr = (iw + 31) >> 5;
    for (int y = tid; y < 96; y += FAST_NT) {
        int cnt = 0;
        if (y < ih) for (int w = 0; w < nwr; ++w) cnt += __popc(bits[y * 3 + w]);
        s_rowOff[y] = cnt;
    }
    __syncthreads();
    const int total = block_excl_scan(s_rowOff, 96, s_warp);
    for (int y = tid; y < ih; y += FAST_NT) {
        uint32_t* out = P.cellList + (size_t)f * P.cellListStride + cd.listOff;
        int o = s_rowOff[y];
        for (int w = 0; w < nwr; ++w)
            for (unsigned m = bits[y * 3 + w]; m; m &= m - 1) {
                const int x = (w << 5) + __ffs(m) - 1;
                out[o++] = pack_cand(x + 3 + cd.offX, y + 3 + cd.offY, score[(y + 1) * sp + (x + 1)]);
            }
    }
    if (tid == 0) *outCount = total;
}


// ------------------------------------------------------------------------------------------
// K1 (TMA form): pyramid level l from level l-1 with the source rows of a 128 x 8 destination tile staged into shared memory by ONE
// TMA bulk-tensor load (box pyrBoxW x pyrBoxH of level l-1, start aligned to 16 bytes); the fixed-point bilinear taps then read
// shared memory instead of gathering bytes from global memory.  Same integer arithmetic as pyr_resize_kernel (cv::resize INTER_LINEAR).
// ------------------------------------------------------------------------------------------
constexpr int PYR_TILE_W = 128, PYR_TILE_H = 8;
__global__ void __launch_bounds__(256) pyr_resize_tma_kernel(ExtractParams P, int l, const __grid_constant__ CUtensorMap srcMap) {
    extern __shared__ __align__(128) uint8_t smem_raw[];
    __shared__ __align__(8) uint64_t s_bar;
    const LevelGeom& G = P.lv[l];
    const int f = blockIdx.z;
    const int tid = threadIdx.y * 32 + threadIdx.x;
    const int dxT = blockIdx.x * PYR_TILE_W, dyT = blockIdx.y * PYR_TILE_H;
    const int cx0 = (int)P.xtab[G.xtabOff + dxT].x & ~15;                 // first source column of the tile, aligned down
    const int ry0 = (int)P.ytab[G.ytabOff + dyT].x;
    if (tid == 0) {
        mbar_init(&s_bar, 1);
        mbar_expect_tx(&s_bar, (uint32_t)(G.pyrBoxW * G.pyrBoxH));
        tma_load_3d(smem_raw, &srcMap, cx0, ry0, f, &s_bar);
    }
    __syncthreads();
    mbar_wait(&s_bar, 0);
    const int dy = dyT + threadIdx.y;
    const int dx0 = dxT + threadIdx.x * 4;
    if (dy >= G.h || dx0 >= G.w) return;
    const short4 yt = P.ytab[G.ytabOff + dy];  // {sy0, sy1, b0, b1}
    const uint8_t* r0 = smem_raw + (yt.x - ry0) * G.pyrBoxW - cx0;
    const uint8_t* r1 = smem_raw + (yt.y - ry0) * G.pyrBoxW - cx0;
    const int b0 = yt.z, b1 = yt.w;
    uint32_t out = 0;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const int dx = dx0 + k;
        if (dx < G.w) {
            const short4 xt = __ldg(&P.xtab[G.xtabOff + dx]);  // {sx0, sx1, a0, a1}
            const int s0 = r0[xt.x] * xt.z + r0[xt.y] * xt.w;
            const int s1 = r1[xt.x] * xt.z + r1[xt.y] * xt.w;
            const int v = (((b0 * (s0 >> 4)) >> 16) + ((b1 * (s1 >> 4)) >> 16) + 2) >> 2;
            out |= (uint32_t)(v & 0xFF) << (8 * k);
        }
    }
    uint8_t* dst = P.pyr + (size_t)f * P.pyrFrameStride + G.planeOff + (size_t)dy * G.pitch;
    *reinterpret_cast<uint32_t*>(dst + dx0) = out;   // pitch is a multiple of 32: in-bounds even for the tail
}

// ------------------------------------------------------------------------------------------
// K3: quadtree distribution + orientation.  One CTA per (level, frame).
// Reference: ORBextractor::DistributeOctTree src/ORBextractor.cc:555-779,
// ExtractorNode::DivideNode :480-536, compareNodes :538-553, IC_Angle :76-103.
//
// The std::list / push_front choreography of the reference is restated as a rebuild of the
// whole node list per pass ("split a set of nodes in processing order, children blocks go to the
// front in reverse processing order, unsplit nodes keep their relative order"); the only
// sequential piece is the libstdc++ std::sort emulation that decides ties in the final phase.
// ------------------------------------------------------------------------------------------
constexpr int QT_NT = 256;

struct QtNode { int16_t ulx, uly, brx, bry; };
struct QtSort { int size; int ulx; int node; };

struct QtShared {
    QtNode* bnd[2];
    int* cnt[2];
    int* proc;        // per list node: processing index or -1
    int* procNode;    // per processing index: list node
    int* childCnt;    // 4 per processing index
    int* childPos;    // 4 per processing index: new list position (or -1)
    int* keepPos;     // per list node: new position when unsplit
    int* scanA;       // scratch, 4*maxNodes
    QtSort* v;        // expandable children in creation order
    unsigned long long* best;
};

__device__ __forceinline__ int qt_child(const QtNode& n, int px, int py, int& midx, int& midy) {
    const int halfX = (int)ceilf((float)(n.brx - n.ulx) / 2);
    const int halfY = (int)ceilf((float)(n.bry - n.uly) / 2);
    midx = n.ulx + halfX; midy = n.uly + halfY;
    return (px < midx ? 0 : 1) + (py < midy ? 0 : 2);
}


// std::sort emulation by the whole CTA (see exact_math.h: rounds of independent tasks).  leaf: capacity n entries; qA / qB: n / 17 + 1 each.
__device__ void cta_libstdcxx_sort(QtSort* a, int n, SxTask* leaf, SxTask* qA, SxTask* qB, int* cnt /* shared: nCur, nNext, nLeaf */) {
    const auto less = [](const QtSort& x, const QtSort& y) { return x.size < y.size || (x.size == y.size && x.ulx < y.ulx); };
    const int tid = threadIdx.x;
    if (n < 2) return;
    if (n <= 16) { if (tid == 0) sx_insertion_sort(a, a + n, less); __syncthreads(); return; }
    if (tid == 0) {
        int depth = 0;
        for (int t = n; t > 1; t >>= 1) depth++;
        qA[0] = SxTask{0, n, 2 * depth};
        cnt[0] = 1; cnt[1] = 0; cnt[2] = 0;
    }
    __syncthreads();
    SxTask *cur = qA, *nxt = qB;
    for (;;) {
        const int nc = cnt[0];
        if (nc == 0) break;
        for (int i = tid; i < nc; i += blockDim.x) {
            const SxTask t = cur[i];
            if (t.d == 0) { sx_heapsort(a + t.lo, t.hi - t.lo, less); continue; }
            const int cut = sx_partition_step(a, t.lo, t.hi, less);
            const SxTask right = {cut, t.hi, t.d - 1}, left = {t.lo, cut, t.d - 1};
            if (right.hi - right.lo > 16) nxt[atomicAdd(&cnt[1], 1)] = right; else if (right.hi - right.lo > 1) leaf[atomicAdd(&cnt[2], 1)] = right;
            if (left.hi - left.lo > 16) nxt[atomicAdd(&cnt[1], 1)] = left; else if (left.hi - left.lo > 1) leaf[atomicAdd(&cnt[2], 1)] = left;
        }
        __syncthreads();
        if (tid == 0) { cnt[0] = cnt[1]; cnt[1] = 0; }
        SxTask* sw = cur; cur = nxt; nxt = sw;
        __syncthreads();
    }
    const int nl = cnt[2];
    for (int i = tid; i < nl; i += blockDim.x) sx_insertion_sort(a + leaf[i].lo, a + leaf[i].hi, less);
    __syncthreads();
}

__global__ void __launch_bounds__(QT_NT, 4) quadtree_orient_kernel(ExtractParams P) {
    extern __shared__ __align__(128) uint8_t smem_raw[];
    const int l = blockIdx.x, f = blockIdx.y, tid = threadIdx.x;
    const LevelGeom& G = P.lv[l];
    const int MN = P.maxNodes;
    __shared__ int s_warp[33];
    __shared__ int s_P;
    __shared__ int s_sortCnt[3];

    // carve shared memory
    QtShared S;
    uint8_t* p = smem_raw;
    S.bnd[0] = (QtNode*)p; p += sizeof(QtNode) * MN;
    S.bnd[1] = (QtNode*)p; p += sizeof(QtNode) * MN;
    S.cnt[0] = (int*)p; p += 4 * MN;
    S.cnt[1] = (int*)p; p += 4 * MN;
    S.proc = (int*)p; p += 4 * MN;
    S.procNode = (int*)p; p += 4 * MN;
    S.keepPos = (int*)p; p += 4 * MN;
    S.childCnt = (int*)p; p += 16 * MN;
    S.childPos = (int*)p; p += 16 * MN;
    S.scanA = (int*)p; p += 16 * MN;
    S.v = (QtSort*)p; p += sizeof(QtSort) * MN;
    p = (uint8_t*)(((uintptr_t)p + 7) & ~(uintptr_t)7);
    S.best = (unsigned long long*)p; p += 8 * MN;
    int* cellOff = (int*)p;   // maxCellsPerLevel + 1

    uint32_t* cand = P.cand + (size_t)f * P.candStride + G.candOff;
    uint16_t* nodeOf = P.nodeOf + (size_t)f * P.candStride + G.candOff;
    SelKp* sel = P.sel + (size_t)f * P.selStride + G.selOff;

    // ---- gather the per-cell lists into candidate order (cell row-major, then y, x) ----
    const int* cellCount = P.cellCount + (size_t)f * P.cellCountStride + G.cellBase;
    for (int c = tid; c < G.nCells; c += QT_NT) cellOff[c] = cellCount[c];
    __syncthreads();
    const int n = block_excl_scan(cellOff, G.nCells, s_warp);
    {
        const uint32_t* lists = P.cellList + (size_t)f * P.cellListStride;
        const int wid = tid >> 5, lane = tid & 31;
        for (int c = wid; c < G.nCells; c += QT_NT / 32) {
            const int cnt = cellCount[c], o = cellOff[c];
            const uint32_t* src = lists + P.cells[G.cellBase + c].listOff;
            for (int k = lane; k < cnt; k += 32) cand[o + k] = src[k];
        }
    }
    __syncthreads();

    int cur = 0;
    // ---- root nodes (:559-598) ----
    for (int i = tid; i < G.nIni; i += QT_NT) {
        QtNode nd;
        nd.ulx = (int16_t)(int)fmul(G.hX, (float)i); nd.uly = 0;
        nd.brx = (int16_t)(int)fmul(G.hX, (float)(i + 1)); nd.bry = (int16_t)(G.maxBY - G.minBY);
        S.bnd[0][i] = nd; S.cnt[0][i] = 0;
    }
    __syncthreads();
    for (int i = tid; i < n; i += QT_NT) {
        const int r = (int)fdiv((float)cand_x(cand[i]), G.hX);
        nodeOf[i] = (uint16_t)r;
        atomicAdd(&S.cnt[0][r], 1);
    }
    __syncthreads();
    // erase empty roots, keep order
    for (int i = tid; i < G.nIni; i += QT_NT) S.scanA[i] = S.cnt[0][i] > 0;
    __syncthreads();
    int m = block_excl_scan(S.scanA, G.nIni, s_warp);
    for (int i = tid; i < G.nIni; i += QT_NT) {
        S.keepPos[i] = S.cnt[0][i] > 0 ? S.scanA[i] : -1;
        if (S.cnt[0][i] > 0) { S.bnd[1][S.scanA[i]] = S.bnd[0][i]; S.cnt[1][S.scanA[i]] = S.cnt[0][i]; }
    }
    __syncthreads();
    for (int i = tid; i < n; i += QT_NT) nodeOf[i] = (uint16_t)S.keepPos[nodeOf[i]];
    cur = 1;
    __syncthreads();

    const int N = G.nDesired;
    bool finish = (n == 0);
    bool careful = false;
    int q = 0;   // size of S.v
    while (!finish) {
        const int prevSize = m;
        QtNode* bndC = S.bnd[cur]; int* cntC = S.cnt[cur];
        QtNode* bndN = S.bnd[cur ^ 1]; int* cntN = S.cnt[cur ^ 1];
        // --- A. processing order of the nodes to split ---
        int P0;
        if (!careful) {
            for (int s = tid; s < m; s += QT_NT) S.scanA[s] = cntC[s] > 1;
            __syncthreads();
            P0 = block_excl_scan(S.scanA, m, s_warp);
            for (int s = tid; s < m; s += QT_NT) {
                if (cntC[s] > 1) { S.proc[s] = S.scanA[s]; S.procNode[S.scanA[s]] = s; }
                else S.proc[s] = -1;
            }
        } else {
            // std::sort(vPrev.begin(), vPrev.end(), compareNodes) (:705): the libstdc++ introsort, whose order among equal keys the result
            // depends on, run as rounds of independent partition steps (one thread per pending range) + independent leaf insertion
            // sorts -- exact_math.h: libstdcxx_sort_rounds_host is the same sequence, checked against std::sort on the host
            cta_libstdcxx_sort(S.v, q, reinterpret_cast<SxTask*>(S.scanA), reinterpret_cast<SxTask*>(S.scanA + 3 * MN),
                               reinterpret_cast<SxTask*>(S.scanA + 3 * MN) + MN / 6, s_sortCnt);
            for (int s = tid; s < m; s += QT_NT) S.proc[s] = -1;
            __syncthreads();
            P0 = q;
            for (int j = tid; j < q; j += QT_NT) {   // processed from the back (:706)
                const int pi = q - 1 - j;
                S.proc[S.v[j].node] = pi; S.procNode[pi] = S.v[j].node;
            }
        }
        for (int i = tid; i < 4 * P0; i += QT_NT) S.childCnt[i] = 0;
        __syncthreads();
        // --- C. distribute the keys of every split node to its four children (:512-526) ---
        for (int i = tid; i < n; i += QT_NT) {
            const int s = nodeOf[i], pi = S.proc[s];
            if (pi >= 0) {
                int mx, my;
                const uint32_t c = cand[i];
                const int ch = qt_child(bndC[s], cand_x(c), cand_y(c), mx, my);
                atomicAdd(&S.childCnt[4 * pi + ch], 1);
            }
        }
        __syncthreads();
        // --- D. in the careful phase stop as soon as the list reaches N nodes (:755-756) ---
        int Pn = P0;
        for (int pi = tid; pi < P0; pi += QT_NT) {
            const int* cc = S.childCnt + 4 * pi;
            S.scanA[pi] = (cc[0] > 0) + (cc[1] > 0) + (cc[2] > 0) + (cc[3] > 0);
        }
        __syncthreads();
        // exclusive scan of k -> scanA; keep k in childPos scratch? recompute instead
        const int totalK = block_excl_scan(S.scanA, P0, s_warp);
        if (careful) {
            if (tid == 0) s_P = P0;
            __syncthreads();
            for (int pi = tid; pi < P0; pi += QT_NT) {
                const int* cc = S.childCnt + 4 * pi;
                const int k = (cc[0] > 0) + (cc[1] > 0) + (cc[2] > 0) + (cc[3] > 0);
                const int sizeAfter = prevSize + (S.scanA[pi] + k) - (pi + 1);
                if (sizeAfter >= N) atomicMin(&s_P, pi + 1);
            }
            __syncthreads();
            Pn = s_P;
            for (int pi = Pn + tid; pi < P0; pi += QT_NT) S.proc[S.procNode[pi]] = -1;
            __syncthreads();
        }
        // children of the first Pn processed nodes: inclusive prefix at Pn-1
        int totalChildren;
        if (Pn == P0) totalChildren = totalK;
        else totalChildren = S.scanA[Pn];   // exclusive prefix at Pn == sum over pi < Pn
        // --- E/G. positions of the children: blocks in reverse processing order, n4..n1 inside ---
        for (int pi = tid; pi < Pn; pi += QT_NT) {
            const int* cc = S.childCnt + 4 * pi;
            const int k = (cc[0] > 0) + (cc[1] > 0) + (cc[2] > 0) + (cc[3] > 0);
            const int blockStart = totalChildren - (S.scanA[pi] + k);
            int after = 0;
            const QtNode pn = bndC[S.procNode[pi]];
            int mx, my;
            qt_child(pn, 0, 0, mx, my);
            for (int ch = 3; ch >= 0; --ch) {
                if (cc[ch] > 0) {
                    const int pos = blockStart + after;
                    ++after;
                    S.childPos[4 * pi + ch] = pos;
                    QtNode c;
                    c.ulx = (ch & 1) ? (int16_t)mx : pn.ulx; c.brx = (ch & 1) ? pn.brx : (int16_t)mx;
                    c.uly = (ch & 2) ? (int16_t)my : pn.uly; c.bry = (ch & 2) ? pn.bry : (int16_t)my;
                    bndN[pos] = c; cntN[pos] = cc[ch];
                } else S.childPos[4 * pi + ch] = -1;
            }
        }
        // --- F. unsplit nodes keep their order behind the new blocks ---
        for (int s = tid; s < m; s += QT_NT) S.keepPos[s] = S.proc[s] < 0;
        __syncthreads();
        const int nKeep = block_excl_scan(S.keepPos, m, s_warp);
        for (int s = tid; s < m; s += QT_NT) {
            if (S.proc[s] < 0) {
                const int pos = totalChildren + S.keepPos[s];
                S.keepPos[s] = pos;
                bndN[pos] = bndC[s]; cntN[pos] = cntC[s];
            }
        }
        // --- new expandable list in creation order (processing order, n1..n4) (:634-668) ---
        for (int i = tid; i < 4 * Pn; i += QT_NT) S.scanA[i] = S.childCnt[i] > 1;
        __syncthreads();
        const int qNew = block_excl_scan(S.scanA, 4 * Pn, s_warp);
        for (int i = tid; i < 4 * Pn; i += QT_NT) {
            if (S.childCnt[i] > 1) {
                QtSort e; e.size = S.childCnt[i]; e.node = S.childPos[i]; e.ulx = bndN[e.node].ulx;
                S.v[S.scanA[i]] = e;
            }
        }
        // --- H. re-home the keys ---
        for (int i = tid; i < n; i += QT_NT) {
            const int s = nodeOf[i], pi = S.proc[s];
            if (pi >= 0) {
                int mx, my;
                const uint32_t c = cand[i];
                const int ch = qt_child(bndC[s], cand_x(c), cand_y(c), mx, my);
                nodeOf[i] = (uint16_t)S.childPos[4 * pi + ch];
            } else nodeOf[i] = (uint16_t)S.keepPos[s];
        }
        __syncthreads();
        m = totalChildren + nKeep;
        q = qNew;
        cur ^= 1;
        // --- termination (:684-764) ---
        if (m >= N || m == prevSize) finish = true;
        else if (!careful && m + 3 * q > N) careful = true;
    }

    // ---- best response per node, first maximum in candidate order (:767-783) ----
    for (int s = tid; s < m; s += QT_NT) S.best[s] = 0ull;
    __syncthreads();
    for (int i = tid; i < n; i += QT_NT) {
        const unsigned long long key = ((unsigned long long)cand_s(cand[i]) << 32) | (unsigned)(0xFFFFFFFFu - (unsigned)i);
        atomicMax(&S.best[nodeOf[i]], key);
    }
    __syncthreads();
    if (tid == 0) {
        P.selCount[f * kMaxLevels + l] = m;
        if (m > G.selCap) atomicOr(&P.status[f], 1);
    }
    const int mOut = min(m, G.selCap);
    // ---- orientation: one warp per keypoint, lane = row of the radius-15 disc (:76-103) ----
    int pitch;
    const uint8_t* img = plane_ptr(P, f, l, pitch);
    const int wid = tid >> 5, lane = tid & 31;
    for (int s = wid; s < mOut; s += QT_NT / 32) {
        const unsigned idx = 0xFFFFFFFFu - (unsigned)(S.best[s] & 0xFFFFFFFFull);
        const uint32_t c = cand[idx];
        const int x = cand_x(c) + G.minBX, y = cand_y(c) + G.minBY;
        // lane = column u of the radius-15 disc, rows walked together: every load instruction of the warp reads one contiguous 31-byte
        // row segment (one or two sectors) instead of 31 different rows (ncu r2a: 28 % of the kernel's stall samples sat on the
        // row-per-lane form of this loop, all long-scoreboard)
        int m10 = 0, m01 = 0;
        {
            const int u = lane - HALF_PATCH;
            const int au = u < 0 ? -u : u;
            const uint8_t* centre = img + (size_t)y * pitch + x + (lane < 31 ? u : 0);
            int colsum = 0;
#pragma unroll
            for (int v = -HALF_PATCH; v <= HALF_PATCH; ++v) {
                const int px = (lane < 31 && au <= c_umax[v < 0 ? -v : v]) ? centre[v * pitch] : 0;
                colsum += px; m01 += v * px;
            }
            m10 = u * colsum;
        }
#pragma unroll
        for (int o = 16; o; o >>= 1) {
            m10 += __shfl_xor_sync(0xffffffffu, m10, o);
            m01 += __shfl_xor_sync(0xffffffffu, m01, o);
        }
        if (lane == 0) {
            SelKp k; k.x = (int16_t)x; k.y = (int16_t)y; k.response = cand_s(c);
            k.angle = fast_atan2_deg((float)m01, (float)m10);
            sel[s] = k;
        }
    }
}

// ------------------------------------------------------------------------------------------
// K4: 7x7 sigma=2 Gaussian blur of every level (cv::GaussianBlur, fixed-point kernel
// [18,34,48,56,48,34,18]/256, BORDER_REFLECT_101; reference call site src/ORBextractor.cc:1133,
// SURVEY.md 9E).  Tile 64x32 per CTA, all levels in one launch.
// ------------------------------------------------------------------------------------------
constexpr int BL_TW = 128, BL_TH = 64, BL_NT = 256;   // tile per CTA; thread = 4 columns x 8 rows
constexpr int BL_SP = BL_TW + 32;                      // smem row pitch = TMA box width: columns x0-16 .. x0+TW+15 (the box must start 16-byte aligned)
constexpr int BL_C0 = 12;                              // tile byte of column x0-4, the first one the filter reads

__device__ __forceinline__ int reflect101(int i, int n) {
    if (n == 1) return 0;
    while (i < 0 || i >= n) i = i < 0 ? -i : 2 * n - 2 - i;
    return i;
}

// One thread produces a 4 (columns) x 8 (rows) strip.  Horizontal pass on pairs of pixels packed as 2 x 16 bit
// (k*p <= 56*255 and the 7-tap sum <= 65280 fit 16 bits, so one IMAD serves two pixels), vertical pass in 32 bit,
// result (v + 32768) >> 16 -- bit-identical to cv::GaussianBlur's fixed-point path.
__global__ void __launch_bounds__(BL_NT) blur_kernel(ExtractParams P, const __grid_constant__ CUtensorMap map0, const CUtensorMap* __restrict__ maps, int useTma) {
    __shared__ __align__(128) uint8_t tile[(BL_TH + 6) * BL_SP];
    __shared__ __align__(8) uint64_t s_bar;
    const int f = blockIdx.y;
    int l = 0;
    while (l + 1 < P.nlevels && (int)blockIdx.x >= P.lv[l + 1].blurTileBase) ++l;
    const LevelGeom& G = P.lv[l];
    const int t = blockIdx.x - G.blurTileBase;
    const int tyb = t / G.blurTilesX, txb = t - tyb * G.blurTilesX;
    const int x0 = txb * BL_TW, y0 = tyb * BL_TH;
    const int tid = threadIdx.x;
    if (useTma && G.w >= 8 && G.h >= 8) {
        // ---- tile: ONE TMA bulk-tensor load of rows y0-3 .. y0+TH+2, columns x0-16 .. x0+TW+15 of the pyramid plane (out-of-range
        //      elements arrive as 0); tiles on the image border then mirror the 3-px halo in shared memory (BORDER_REFLECT_101) ----
        if (tid == 0) {
            mbar_init(&s_bar, 1);
            mbar_expect_tx(&s_bar, (uint32_t)((BL_TH + 6) * BL_SP));
            tma_load_3d(tile, l == 0 ? &map0 : maps + l, x0 - 16, y0 - 3, f, &s_bar);
        }
        __syncthreads();
        mbar_wait(&s_bar, 0);
        const bool left = x0 == 0, right = x0 + BL_TW + 3 >= G.w, top = y0 == 0, bottom = y0 + BL_TH + 3 >= G.h;
        if (left || right) {
            for (int i = tid; i < (BL_TH + 6) * 8; i += BL_NT) {           // per row: 4 halo columns on either side
                const int r = i >> 3, k = i & 7;
                if (k < 4) { if (left) { const int x = -4 + k; tile[r * BL_SP + 16 + x] = tile[r * BL_SP + 16 - x]; } }
                else if (right) {
                    const int x = G.w + (k - 4);                           // columns w .. w+3
                    if (x - x0 < BL_TW + 4) tile[r * BL_SP + 16 + x - x0] = tile[r * BL_SP + 16 + (2 * G.w - 2 - x) - x0];
                }
            }
            __syncthreads();
        }
        if (top || bottom) {
            for (int i = tid; i < 6 * (BL_SP / 4); i += BL_NT) {           // 3 halo rows above / below, whole rows as words
                const int k = i / (BL_SP / 4), wq = i - k * (BL_SP / 4);
                int y;
                if (k < 3) { if (!top) continue; y = -3 + k; }
                else { if (!bottom) continue; y = G.h + (k - 3); if (y - y0 + 3 >= BL_TH + 6) continue; }
                const int ys = y < 0 ? -y : 2 * G.h - 2 - y;
                reinterpret_cast<uint32_t*>(tile + (y - y0 + 3) * BL_SP)[wq] = reinterpret_cast<const uint32_t*>(tile + (ys - y0 + 3) * BL_SP)[wq];
            }
            __syncthreads();
        }
    } else {
        int pitch;
        const uint8_t* img = plane_ptr(P, f, l, pitch);
        // ---- fallback tile load (level 0 not TMA-addressable, or a plane smaller than the halo logic assumes) ----
        for (int i = tid; i < (BL_TH + 6) * 34; i += BL_NT) {
            const int r = i / 34, wq = i - r * 34;
            const int sy = reflect101(min(y0 + r - 3, G.h + 2), G.h);
            const uint8_t* row = img + (size_t)sy * pitch;
            uint32_t v = 0;
#pragma unroll
            for (int k = 0; k < 4; ++k) v |= (uint32_t)row[reflect101(min(x0 - 4 + 4 * wq + k, G.w + 3), G.w)] << (8 * k);
            *reinterpret_cast<uint32_t*>(tile + r * BL_SP + BL_C0 + 4 * wq) = v;
        }
        __syncthreads();
    }
    const int tx = tid & 31, ty = tid >> 5;          // 32 x 8 threads
    const int xo = x0 + 4 * tx, yo = y0 + 8 * ty;
    if (xo >= G.w || yo >= G.h) return;
    // Fused H + V: the 14 input rows of the strip are walked once.  Horizontal pass on pairs of pixels packed as 2 x 16 bit; every
    // horizontal sum is unpacked ONCE and scattered with its tap weight into the (up to 7) output rows it contributes to -- the
    // accumulators of the 8 x 4 outputs stay in registers (the earlier form unpacked every sum once per output row: 7x the work).
    uint32_t acc[8][4];
#pragma unroll
    for (int r = 0; r < 8; ++r)
#pragma unroll
        for (int k = 0; k < 4; ++k) acc[r][k] = 0;
#pragma unroll
    for (int r = 0; r < 14; ++r) {
        const uint32_t* w = reinterpret_cast<const uint32_t*>(tile + (8 * ty + r) * BL_SP + BL_C0 + 4 * tx);
        const uint32_t w0 = w[0], w1 = w[1], w2 = w[2];   // bytes b0..b11 = columns xo-4 .. xo+7
        // pair words Ps = b_s | b_{s+1} << 16 (two pixels as 2 x 16 bit), s = 1..9: 32-bit window at byte s, then spread bytes 0,1
#define ORBX_PAIR(lo, hi, sh) __byte_perm(__funnelshift_r(lo, hi, sh), 0u, 0x4140)
        const uint32_t P1 = ORBX_PAIR(w0, w1, 8), P2 = ORBX_PAIR(w0, w1, 16), P3 = ORBX_PAIR(w0, w1, 24);
        const uint32_t P4 = ORBX_PAIR(w1, w2, 0), P5 = ORBX_PAIR(w1, w2, 8), P6 = ORBX_PAIR(w1, w2, 16), P7 = ORBX_PAIR(w1, w2, 24);
        const uint32_t P8 = ORBX_PAIR(w2, 0u, 0), P9 = ORBX_PAIR(w2, 0u, 8);
#undef ORBX_PAIR
        const uint32_t h01 = 18u * (P1 + P7) + 34u * (P2 + P6) + 48u * (P3 + P5) + 56u * P4;
        const uint32_t h23 = 18u * (P3 + P9) + 34u * (P4 + P8) + 48u * (P5 + P7) + 56u * P6;
        const uint32_t h[4] = {h01 & 0xFFFFu, h01 >> 16, h23 & 0xFFFFu, h23 >> 16};
#pragma unroll
        for (int j = 0; j < 7; ++j) {
            const int o = r - j;                         // input row r is tap j of output row r - j
            if (o >= 0 && o < 8) {
                const uint32_t kj = j == 0 || j == 6 ? 18u : (j == 1 || j == 5 ? 34u : (j == 2 || j == 4 ? 48u : 56u));
#pragma unroll
                for (int k = 0; k < 4; ++k) acc[o][k] += kj * h[k];
            }
        }
    }
    int bpitch;
    uint8_t* out = const_cast<uint8_t*>(blur_ptr(P, f, l, bpitch));
#pragma unroll
    for (int r = 0; r < 8; ++r) {
        if (yo + r >= G.h) break;
        uint32_t pk = 0;
#pragma unroll
        for (int k = 0; k < 4; ++k) pk |= ((acc[r][k] + 32768u) >> 16) << (8 * k);
        uint8_t* o = out + (size_t)(yo + r) * bpitch + xo;
        if (xo + 3 < G.w) *reinterpret_cast<uint32_t*>(o) = pk;     // bpitch and xo are multiples of 4
        else {
            for (int k = 0; k < 4 && xo + k < G.w; ++k) o[k] = (uint8_t)(pk >> (8 * k));
        }
    }
}

// ------------------------------------------------------------------------------------------
// K5: output assembly.  One CTA per frame: scale coordinates, split lapping / non-lapping
// keypoints (front / back fill), write the cv::KeyPoint-layout slab and the row map.
// Reference: ORBextractor::operator() src/ORBextractor.cc:1120-1164.
// ------------------------------------------------------------------------------------------
constexpr int AS_NT = 256;

__global__ void __launch_bounds__(AS_NT) assemble_kernel(ExtractParams P) {
    extern __shared__ int s_flags[];   // selStride ints
    __shared__ int s_warp[33];
    __shared__ int s_lvOff[kMaxLevels + 1];
    const int f = blockIdx.x, tid = threadIdx.x;
    if (tid == 0) {
        int o = 0;
        for (int l = 0; l < P.nlevels; ++l) { s_lvOff[l] = o; o += min(P.selCount[f * kMaxLevels + l], P.lv[l].selCap); }
        s_lvOff[P.nlevels] = o;
    }
    __syncthreads();
    const int K = s_lvOff[P.nlevels];
    const SelKp* sel = P.sel + (size_t)f * P.selStride;
    // flag = keypoint lies in the lapping area (scaled x within [lap0, lap1])
    for (int e = tid; e < K; e += AS_NT) {
        int l = 0;
        while (e >= s_lvOff[l + 1]) ++l;
        const SelKp k = sel[P.lv[l].selOff + (e - s_lvOff[l])];
        float x = (float)k.x;
        if (l != 0) x = fmul(x, P.lv[l].scale);
        s_flags[e] = (x >= (float)P.lap0 && x <= (float)P.lap1) ? 1 : 0;
    }
    __syncthreads();
    // exclusive scan -> number of lapping keypoints before e
    int* s_scan = s_flags + P.selStride;
    for (int e = tid; e < K; e += AS_NT) s_scan[e] = s_flags[e];
    __syncthreads();
    const int nLap = block_excl_scan(s_scan, K, s_warp);
    int* dstIndex = P.dstIndex + (size_t)f * P.selStride;
    OrbKeyPoint* out = reinterpret_cast<OrbKeyPoint*>(P.outKp) + (size_t)f * P.outCap;
    const bool fits = K <= P.outCap;
    for (int e = tid; e < K; e += AS_NT) {
        int l = 0;
        while (e >= s_lvOff[l + 1]) ++l;
        const int src = P.lv[l].selOff + (e - s_lvOff[l]);
        const SelKp k = sel[src];
        const int at = s_flags[e] ? (K - 1 - s_scan[e]) : (e - s_scan[e]);
        dstIndex[src] = fits ? at : -1;
        if (fits) {
            OrbKeyPoint o;
            o.x = (float)k.x; o.y = (float)k.y;
            if (l != 0) { o.x = fmul(o.x, P.lv[l].scale); o.y = fmul(o.y, P.lv[l].scale); }
            o.size = P.lv[l].sizeScaled; o.angle = k.angle; o.response = (float)k.response;
            o.octave = l; o.class_id = -1;
            out[at] = o;
        }
    }
    if (tid == 0) {
        P.outN[f] = K;
        P.outMono[f] = K - nLap;
        if (!fits) atomicOr(&P.status[f], 2);
    }
}

// ------------------------------------------------------------------------------------------
// K6: rotated BRIEF.  One warp per keypoint, lane = descriptor byte.
// Reference: computeOrbDescriptor src/ORBextractor.cc:107-146 on the blurred plane.
// ------------------------------------------------------------------------------------------
constexpr int DS_NT = 256;

__global__ void __launch_bounds__(DS_NT) describe_kernel(ExtractParams P) {
    __shared__ char4 s_pat[8][32];   // [pair-in-byte][byte] -> conflict-free across lanes
    const int f = blockIdx.y, tid = threadIdx.x, lane = tid & 31;
    for (int i = tid; i < 256; i += DS_NT) {
        const int byte = i >> 3, k = i & 7;
        s_pat[k][byte] = make_char4(c_pattern[4 * i], c_pattern[4 * i + 1], c_pattern[4 * i + 2], c_pattern[4 * i + 3]);
    }
    __syncthreads();
    const int slot = blockIdx.x * (DS_NT / 32) + (tid >> 5);   // index into the frame's sel array
    if (slot >= (int)P.selStride) return;
    int l = 0;
    while (l + 1 < P.nlevels && slot >= P.lv[l + 1].selOff) ++l;
    const int idx = slot - P.lv[l].selOff;
    if (idx >= min(P.selCount[f * kMaxLevels + l], P.lv[l].selCap)) return;
    const int at = P.dstIndex[(size_t)f * P.selStride + slot];
    if (at < 0) return;
    const SelKp k = P.sel[(size_t)f * P.selStride + slot];
    const float factorPI = 0.01745329238474369049072265625f;   // (float)(CV_PI/180.f), :106
    float a, b;
    sincosf_glibc(fmul(k.angle, factorPI), &b, &a);
    int pitch;
    const uint8_t* img = blur_ptr(P, f, l, pitch);
    const uint8_t* c = img + (size_t)k.y * pitch + k.x;
    int val = 0;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        const char4 q = s_pat[j][lane];
        const float x0 = (float)q.x, y0 = (float)q.y, x1 = (float)q.z, y1 = (float)q.w;
        const int r0 = round_half_even(ffma(x0, b, fmul(y0, a)));
        const int c0 = round_half_even(ffma(x0, a, -fmul(y0, b)));
        const int r1 = round_half_even(ffma(x1, b, fmul(y1, a)));
        const int c1 = round_half_even(ffma(x1, a, -fmul(y1, b)));
        const int t0 = c[r0 * pitch + c0], t1 = c[r1 * pitch + c1];
        val |= (t0 < t1) << j;
    }
    P.outDesc[((size_t)f * P.outCap + at) * 32 + lane] = (uint8_t)val;
}


// ------------------------------------------------------------------------------------------
// K7: Frame::ComputeStereoMatches (reference src/Frame.cc:811-982), the stereo consumer of mvImagePyramid.  One CTA per rectified
// pair; a warp per left keypoint: (1) best right keypoint on the same row band (|dy| <= 2 * scale of the right keypoint, level
// within +-1, disparity window) by Hamming distance, ties to the lower right index (= the reference's row-table order);
// (2) 11 x 11 SAD refinement over 11 shifts on the pyramid planes of the left keypoint's level, parabola fit, disparity / depth;
// then (3) the CTA removes matches whose SAD is >= 1.5 * 1.4 * median.  Float arithmetic is individually rounded like the oracle's.
// The 19-px reflected frame the reference keeps around every plane is not needed: for keypoints the extractor can produce the
// windows stay >= 5 px inside the planes; a window that would leave the plane flags status bit 4 instead of reading.
// ------------------------------------------------------------------------------------------
// copyMakeBorder(level, BORDER_REFLECT_101) around one pyramid plane (src/ORBextractor.cc:1185-1191), written on request only
__global__ void border_copy_kernel(ExtractParams P, int f, int l, int border, uint8_t* __restrict__ dst) {
    const LevelGeom& G = P.lv[l];
    const int W = G.w + 2 * border;
    const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
    if (x >= W) return;
    int pitch;
    const uint8_t* img = plane_ptr(P, f, l, pitch);
    dst[(size_t)y * W + x] = img[(size_t)reflect101(y - border, G.h) * pitch + reflect101(x - border, G.w)];
}

constexpr int ST_NT = 256;
struct StereoParams {
    ExtractParams L, R;                    // geometry + pyramid pointers of the two extractors (same image size)
    const OrbKeyPoint *kpsL, *kpsR; const uint8_t *descL, *descR; const int *nL, *nR; int capL, capR;
    float mb, mbf;
    float* uRight; float* depth;           // [batch][capL]
    int* sad;                              // scratch [batch][capL]
    int* status;                           // [batch]
};

__global__ void __launch_bounds__(ST_NT) stereo_matches_kernel(StereoParams Q) {
    __shared__ int s_cnt, s_median;
    const int f = blockIdx.x, tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const int N = min(Q.nL[f], Q.capL), Nr = min(Q.nR[f], Q.capR);
    const OrbKeyPoint* kl = Q.kpsL + (size_t)f * Q.capL; const OrbKeyPoint* kr = Q.kpsR + (size_t)f * Q.capR;
    const uint8_t* dl = Q.descL + (size_t)f * Q.capL * 32; const uint8_t* dr = Q.descR + (size_t)f * Q.capR * 32;
    float* uRight = Q.uRight + (size_t)f * Q.capL; float* depth = Q.depth + (size_t)f * Q.capL;
    int* sad = Q.sad + (size_t)f * Q.capL;
    const float minD = 0.f, maxD = fdiv(Q.mbf, Q.mb);
    if (tid == 0) s_cnt = 0;
    __syncthreads();
    for (int iL = wid; iL < N; iL += ST_NT / 32) {
        const OrbKeyPoint kpL = kl[iL];
        float outU = -1.0f, outD = -1.0f; int outSad = -1;
        const float uL = kpL.x, vL = kpL.y;
        const int rowL = (int)vL;                                    // vRowIndices[vL]: float -> index by truncation
        const float minU = fsub(uL, maxD), maxU = fsub(uL, minD);
        unsigned best = 0xFFFFFFFFu;
        if (!(maxU < 0)) {
            const uint4* pl = reinterpret_cast<const uint4*>(dl + (size_t)iL * 32);
            const uint4 a0 = __ldg(pl), a1 = __ldg(pl + 1);
            const uint32_t da[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
            for (int iR = lane; iR < Nr; iR += 32) {
                const OrbKeyPoint kpR = kr[iR];
                const float r = fmul(2.0f, Q.L.lv[min(max(kpR.octave, 0), Q.L.nlevels - 1)].scale);
                const int maxr = (int)ceilf(fadd(kpR.y, r)), minr = (int)floorf(fsub(kpR.y, r));
                if (rowL < minr || rowL > maxr) continue;
                if (kpR.octave < kpL.octave - 1 || kpR.octave > kpL.octave + 1) continue;
                if (!(kpR.x >= minU && kpR.x <= maxU)) continue;
                const uint4* pr = reinterpret_cast<const uint4*>(dr + (size_t)iR * 32);
                const uint4 b0 = __ldg(pr), b1 = __ldg(pr + 1);
                const uint32_t db[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
                const unsigned dist = (unsigned)hamming256(da, db);
                if (dist < 100u) best = min(best, (dist << 16) | (unsigned)iR);      // bestDist starts at TH_HIGH, strict <
            }
        }
        best = __reduce_min_sync(0xffffffffu, best);
        if (best != 0xFFFFFFFFu && (int)(best >> 16) < (100 + 50) / 2) {             // thOrbDist
            const int bestIdxR = (int)(best & 0xFFFFu);
            const int lvl = min(max(kpL.octave, 0), Q.L.nlevels - 1);
            const float uR0 = kr[bestIdxR].x;
            const float scaleFactor = fdiv(1.0f, Q.L.lv[lvl].scale);                  // mvInvScaleFactors[octave] = 1.0f / mvScaleFactor[octave]
            const float scaleduL = roundf(fmul(kpL.x, scaleFactor)), scaledvL = roundf(fmul(kpL.y, scaleFactor)), scaleduR0 = roundf(fmul(uR0, scaleFactor));
            const int w = 5, Lr = 5;
            const LevelGeom& G = Q.L.lv[lvl];
            const float iniu = fsub(fadd(scaleduR0, (float)Lr), (float)w), endu = fadd(fadd(fadd(scaleduR0, (float)Lr), (float)w), 1.0f);
            if (!(iniu < 0 || endu >= (float)G.w)) {
                const int y0 = (int)fsub(scaledvL, (float)w), xL0 = (int)fsub(scaleduL, (float)w), xR00 = (int)scaleduR0 - Lr - w;
                if (y0 < 0 || y0 + 2 * w + 1 > G.h || xL0 < 0 || xL0 + 2 * w + 1 > G.w || xR00 < 0) { if (lane == 0) atomicOr(&Q.status[f], 4); }
                else {
                    int pitchL, pitchR;
                    const uint8_t* IL = plane_ptr(Q.L, f, lvl, pitchL);
                    const uint8_t* IR = plane_ptr(Q.R, f, lvl, pitchR);
                    // this lane's pixels of the 11 x 11 left window
                    int pa[4], rr[4], cc[4];
#pragma unroll
                    for (int k = 0; k < 4; ++k) {
                        const int p = lane + 32 * k;
                        rr[k] = p / 11; cc[k] = p - rr[k] * 11;
                        pa[k] = p < 121 ? IL[(size_t)(y0 + rr[k]) * pitchL + xL0 + cc[k]] : 0;
                    }
                    int bestSad = 0x7fffffff, bestincR = 0;
                    float vDists[11];
#pragma unroll
                    for (int incR = -5; incR <= 5; ++incR) {
                        const int xR0 = (int)fsub(fadd(scaleduR0, (float)incR), (float)w);
                        int s = 0;
#pragma unroll
                        for (int k = 0; k < 4; ++k) {
                            const int p = lane + 32 * k;
                            if (p < 121) { const int b = IR[(size_t)(y0 + rr[k]) * pitchR + xR0 + cc[k]]; s += abs(pa[k] - b); }
                        }
                        s = __reduce_add_sync(0xffffffffu, s);
                        const float dist = (float)s;
                        if (dist < (float)bestSad) { bestSad = (int)dist; bestincR = incR; }
                        vDists[incR + 5] = dist;
                    }
                    if (!(bestincR == -Lr || bestincR == Lr)) {
                        float dist1 = 0, dist2 = 0, dist3 = 0;
#pragma unroll
                        for (int k = 1; k < 10; ++k) if (k == bestincR + 5) { dist1 = vDists[k - 1]; dist2 = vDists[k]; dist3 = vDists[k + 1]; }
                        const float deltaR = fdiv(fsub(dist1, dist3), fmul(2.0f, fsub(fadd(dist1, dist3), fmul(2.0f, dist2))));
                        if (!(deltaR < -1 || deltaR > 1)) {
                            float bestuR = fmul(G.scale, fadd(fadd(scaleduR0, (float)bestincR), deltaR));
                            float disparity = fsub(uL, bestuR);
                            if (disparity >= minD && disparity < maxD) {
                                if (disparity <= 0) { disparity = 0.01f; bestuR = (float)((double)uL - 0.01); }
                                outD = fdiv(Q.mbf, disparity); outU = bestuR; outSad = bestSad;
                            }
                        }
                    }
                }
            }
        }
        if (lane == 0) { uRight[iL] = outU; depth[iL] = outD; sad[iL] = outSad; if (outSad >= 0) atomicAdd(&s_cnt, 1); }
    }
    __syncthreads();
    // median of the SAD distances of the accepted matches: element size/2 of the (distance, iL)-sorted list (:964-966)
    const int cnt = s_cnt;
    if (cnt == 0) return;
    for (int i = tid; i < N; i += ST_NT) {
        const int di = sad[i];
        if (di < 0) continue;
        int rank = 0;
        for (int j = 0; j < N; ++j) { const int dj = sad[j]; rank += dj >= 0 && (dj < di || (dj == di && j < i)); }
        if (rank == cnt / 2) s_median = di;
    }
    __syncthreads();
    const float thDist = fmul(1.5f * 1.4f, (float)s_median);
    for (int i = tid; i < N; i += ST_NT) {
        const int di = sad[i];
        if (di >= 0 && !((float)di < thDist)) { uRight[i] = -1.0f; depth[i] = -1.0f; }
    }
}

}  // namespace orbx
