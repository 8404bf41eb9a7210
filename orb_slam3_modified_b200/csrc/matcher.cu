// B200 kernels + C-ABI for the per-frame matchers of ORBmatcher (reference src/ORBmatcher.cc):
//   SearchByProjection(Frame&, const vector<MapPoint*>&, ...)  :43-213   -> orbm_search_local_map
//   SearchByProjection(Frame&, const Frame&, th, bMono)        :1676-1887 -> orbm_search_last_frame[_batch_device]
//   DescriptorDistance                                          :2058-2074 -> hamming256 / orbm_descriptor_distance
//   cv::BFMatcher(NORM_HAMMING).knnMatch(k=2) (src/Frame.cc:1144)          -> orbm_bf_knn2
// plus Frame::AssignFeaturesToGrid / GetFeaturesInArea (src/Frame.cc:385-416,657-723) rebuilt on the device.
//
// Parallelisation of the reference's sequential claim rule (a keypoint that already holds a map point with
// observations is skipped by later map points, :84-86 / :1747-1749):
//   pass 1 (warp per map point, all streams at once): top-2 candidates by (Hamming distance, enumeration order)
//           against the claim state at entry;
//   pass 2 (one warp per stream, map points in index order): a pass-1 result is still exact unless its best or
//           second-best keypoint has been claimed meanwhile -- only then the warp rescans that map point against
//           the current claim state.  Claims live in a shared-memory bitset.
#include <cuda_runtime.h>
#include <math.h>
#include <string.h>
#include <string>
#include <vector>

#include "../../include/orb_b200.h"
#include "host_affinity.h"
#include "device_utils.cuh"
#include "exact_math.h"

namespace orbx {
void set_error(const std::string& s);
}
using orbx::set_error;

#define CK(call)                                                                   \
    do {                                                                           \
        cudaError_t e_ = (call);                                                   \
        if (e_ != cudaSuccess) {                                                   \
            set_error(std::string(#call) + ": " + cudaGetErrorString(e_));         \
            return ORB_ERR_CUDA;                                                   \
        }                                                                          \
    } while (0)

namespace orbm {

using namespace orbx;

constexpr int GRID_COLS = 64, GRID_ROWS = 48, GRID_CELLS = GRID_COLS * GRID_ROWS;   // include/Frame.h:44-45
constexpr int TH_HIGH = 100;       // src/ORBmatcher.cc:35
constexpr int HISTO_LENGTH = 30;   // :37

struct MatchParams {
    int batch, kcap, mcap, nlevels, mode;   // mode 0 = local map, 1 = last frame
    const OrbKeyPoint* kps; const uint8_t* desc; const int* nK;
    float minX, minY, maxX, maxY, gridWInv, gridHInv;
    const float* scaleFactors;
    const int* nM;
    const uint8_t *inView, *bad, *hasObs, *mpDesc, *valid;
    const float *depth, *projX, *projY, *viewCos, *xyz, *angle, *Tcw7;
    const int *level, *octave;
    float cam[4];
    float th, nnratio, thFar;
    int bFar, checkOri, resetState;
    // scratch
    int* cellStart;      // [batch][GRID_CELLS + 1]
    uint16_t* cellIdx;   // [batch][kcap]
    float4* query;       // [batch][mcap]  u, v, r, bits(minLevel+1 | (maxLevel+1) << 8 | valid << 16)
    int4 *resultIdx, *resultDist;   // [batch][mcap]  the four best (keypoint index, distance) of pass 1, ascending
    uint8_t* evBin; uint16_t* evIdx;   // [batch][mcap] rotation-histogram events
    // in/out
    int* match; uint8_t* claimed; int* nmatches;
    int* status;
};

// ---------------------------------------------------------------------------------------------
// Frame::AssignFeaturesToGrid + PosInGrid: counting sort of the keypoints into the 64x48 grid, cell lists in
// increasing keypoint index (= the reference's push_back order).  One CTA per stream.
// ---------------------------------------------------------------------------------------------
constexpr int GB_NT = 256;
__global__ void __launch_bounds__(GB_NT) grid_build_kernel(MatchParams P) {
    __shared__ int s_cnt[GRID_CELLS];
    __shared__ int s_warp[33];
    const int f = blockIdx.x, tid = threadIdx.x;
    const int K = min(P.nK[f], P.kcap);
    const OrbKeyPoint* kps = P.kps + (size_t)f * P.kcap;
    int* cellStart = P.cellStart + (size_t)f * (GRID_CELLS + 1);
    uint16_t* cellIdx = P.cellIdx + (size_t)f * P.kcap;
    for (int c = tid; c < GRID_CELLS; c += GB_NT) s_cnt[c] = 0;
    if (P.resetState) {
        int* match = P.match + (size_t)f * P.kcap;
        uint8_t* claimed = P.claimed + (size_t)f * P.kcap;
        for (int i = tid; i < P.kcap; i += GB_NT) { match[i] = -1; claimed[i] = 0; }
    }
    __syncthreads();
    for (int i = tid; i < K; i += GB_NT) {
        const int px = (int)roundf(fmul(fsub(kps[i].x, P.minX), P.gridWInv));
        const int py = (int)roundf(fmul(fsub(kps[i].y, P.minY), P.gridHInv));
        if (px >= 0 && px < GRID_COLS && py >= 0 && py < GRID_ROWS) atomicAdd(&s_cnt[px * GRID_ROWS + py], 1);
    }
    __syncthreads();
    const int total = block_excl_scan(s_cnt, GRID_CELLS, s_warp);
    for (int c = tid; c < GRID_CELLS; c += GB_NT) cellStart[c] = s_cnt[c];
    if (tid == 0) cellStart[GRID_CELLS] = total;
    __syncthreads();
    for (int i = tid; i < K; i += GB_NT) {
        const int px = (int)roundf(fmul(fsub(kps[i].x, P.minX), P.gridWInv));
        const int py = (int)roundf(fmul(fsub(kps[i].y, P.minY), P.gridHInv));
        if (px >= 0 && px < GRID_COLS && py >= 0 && py < GRID_ROWS) {
            const int pos = atomicAdd(&s_cnt[px * GRID_ROWS + py], 1);
            cellIdx[pos] = (uint16_t)i;
        }
    }
    __syncthreads();
    // restore insertion order inside each cell (lists are short)
    for (int c = tid; c < GRID_CELLS; c += GB_NT) {
        const int a = cellStart[c], b = s_cnt[c];   // s_cnt now holds the end offset
        for (int i = a + 1; i < b; ++i) {
            const uint16_t v = cellIdx[i];
            int j = i - 1;
            while (j >= a && cellIdx[j] > v) { cellIdx[j + 1] = cellIdx[j]; --j; }
            cellIdx[j + 1] = v;
        }
    }
}

// ---------------------------------------------------------------------------------------------
// Warp-wide scan of the candidates of one map point: Frame::GetFeaturesInArea order = (ix, iy, position in cell).
// Keeps the two smallest (distance, order) keys among keypoints that are not claimed.
// key = dist << 40 | cellRank << 20 | j ; returns idx/dist of best and second (idx -1 when absent).
// ---------------------------------------------------------------------------------------------
template <int K>
struct TopK {   // the K smallest keys seen, ascending
    unsigned long long k[K];
    int i[K];
    __device__ __forceinline__ void init() {
#pragma unroll
        for (int j = 0; j < K; ++j) { k[j] = ~0ull; i[j] = -1; }
    }
    __device__ __forceinline__ void insert(unsigned long long key, int idx) {
        if (key >= k[K - 1]) return;
        k[K - 1] = key; i[K - 1] = idx;
#pragma unroll
        for (int j = K - 1; j > 0; --j) {
            if (k[j] < k[j - 1]) {
                const unsigned long long tk = k[j]; k[j] = k[j - 1]; k[j - 1] = tk;
                const int ti = i[j]; i[j] = i[j - 1]; i[j - 1] = ti;
            }
        }
    }
    __device__ __forceinline__ void warp_merge() {   // afterwards every lane holds the K smallest of the whole warp
#pragma unroll
        for (int o = 16; o; o >>= 1) {
            unsigned long long ok[K]; int oi[K];
#pragma unroll
            for (int j = 0; j < K; ++j) { ok[j] = __shfl_xor_sync(0xffffffffu, k[j], o); oi[j] = __shfl_xor_sync(0xffffffffu, i[j], o); }
#pragma unroll
            for (int j = 0; j < K; ++j) insert(ok[j], oi[j]);
        }
    }
};
typedef TopK<2> Top2;

template <int K, class ClaimFn>
__device__ __forceinline__ TopK<K> scan_candidates(const MatchParams& P, int f, float u, float v, float r, int minLevel, int maxLevel,
                                                   const uint32_t* mpd, ClaimFn isClaimed) {
    const int lane = threadIdx.x & 31;
    TopK<K> t; t.init();
    const int nMinCellX = max(0, (int)floorf(fmul(fsub(fsub(u, P.minX), r), P.gridWInv)));
    const int nMaxCellX = min(GRID_COLS - 1, (int)ceilf(fmul(fadd(fsub(u, P.minX), r), P.gridWInv)));
    const int nMinCellY = max(0, (int)floorf(fmul(fsub(fsub(v, P.minY), r), P.gridHInv)));
    const int nMaxCellY = min(GRID_ROWS - 1, (int)ceilf(fmul(fadd(fsub(v, P.minY), r), P.gridHInv)));
    if (nMinCellX < GRID_COLS && nMaxCellX >= 0 && nMinCellY < GRID_ROWS && nMaxCellY >= 0) {
        const int ny = nMaxCellY - nMinCellY + 1, nx = nMaxCellX - nMinCellX + 1;
        const int* cellStart = P.cellStart + (size_t)f * (GRID_CELLS + 1);
        const uint16_t* cellIdx = P.cellIdx + (size_t)f * P.kcap;
        const OrbKeyPoint* kps = P.kps + (size_t)f * P.kcap;
        const uint8_t* desc = P.desc + (size_t)f * P.kcap * 32;
        const bool bCheckLevels = (minLevel > 0) || (maxLevel >= 0);
        for (int c = lane; c < nx * ny; c += 32) {
            const int ix = nMinCellX + c / ny, iy = nMinCellY + c % ny;
            const int cell = ix * GRID_ROWS + iy;
            const int a = cellStart[cell], b = cellStart[cell + 1];
            for (int e = a; e < b; ++e) {
                const int idx = cellIdx[e];
                const OrbKeyPoint kp = kps[idx];
                if (bCheckLevels) {
                    if (kp.octave < minLevel) continue;
                    if (maxLevel >= 0 && kp.octave > maxLevel) continue;
                }
                const float dx = fsub(kp.x, u), dy = fsub(kp.y, v);
                if (!(fabsf(dx) < r && fabsf(dy) < r)) continue;
                if (isClaimed(idx)) continue;
                const uint4* dp = reinterpret_cast<const uint4*>(desc + (size_t)idx * 32);
                const uint4 d0 = __ldg(dp), d1 = __ldg(dp + 1);
                const uint32_t dd[8] = {d0.x, d0.y, d0.z, d0.w, d1.x, d1.y, d1.z, d1.w};
                const int dist = hamming256(mpd, dd);
                t.insert(((unsigned long long)dist << 40) | ((unsigned long long)c << 20) | (unsigned)(e - a), idx);
            }
        }
    }
    t.warp_merge();
    return t;
}

__device__ __forceinline__ void load_mp_desc(const MatchParams& P, int f, int i, uint32_t* mpd) {
    const uint4* p = reinterpret_cast<const uint4*>(P.mpDesc + ((size_t)f * P.mcap + i) * 32);
    const uint4 a = __ldg(p), b = __ldg(p + 1);
    mpd[0] = a.x; mpd[1] = a.y; mpd[2] = a.z; mpd[3] = a.w; mpd[4] = b.x; mpd[5] = b.y; mpd[6] = b.z; mpd[7] = b.w;
}

// Query of map point i: search centre, radius and level window; returns false when the reference `continue`s.
__device__ __forceinline__ bool make_query(const MatchParams& P, int f, int i, float& u, float& v, float& r, int& minL, int& maxL) {
    const size_t o = (size_t)f * P.mcap + i;
    if (P.mode == 0) {   // src/ORBmatcher.cc:50-72
        if (!P.inView[o]) return false;
        if (P.bFar && P.depth[o] > P.thFar) return false;
        if (P.bad[o]) return false;
        const int lvl = P.level[o];
        float rr = ((double)P.viewCos[o] > 0.998) ? 2.5f : 4.0f;   // RadiusByViewingCos :215-221
        if (P.th != 1.0f) rr = fmul(rr, P.th);
        u = P.projX[o]; v = P.projY[o];
        r = fmul(rr, P.scaleFactors[lvl]);
        minL = lvl - 1; maxL = lvl;
        return true;
    }
    // last frame, mono (:1702-1741): x3Dc = Tcw * x3Dw with Sophus' quaternion form (so3.hpp:358-367)
    if (!P.valid[o]) return false;
    const float* T = P.Tcw7 + (size_t)f * 7;
    const float qw = T[0], qx = T[1], qy = T[2], qz = T[3];
    const float px = P.xyz[3 * o], py = P.xyz[3 * o + 1], pz = P.xyz[3 * o + 2];
    float ux = fsub(fmul(qy, pz), fmul(qz, py)), uy = fsub(fmul(qz, px), fmul(qx, pz)), uz = fsub(fmul(qx, py), fmul(qy, px));
    ux = fadd(ux, ux); uy = fadd(uy, uy); uz = fadd(uz, uz);
    const float cx_ = fsub(fmul(qy, uz), fmul(qz, uy)), cy_ = fsub(fmul(qz, ux), fmul(qx, uz)), cz_ = fsub(fmul(qx, uy), fmul(qy, ux));
    const float xc = fadd(fadd(fadd(px, fmul(qw, ux)), cx_), T[4]);
    const float yc = fadd(fadd(fadd(py, fmul(qw, uy)), cy_), T[5]);
    const float zc = fadd(fadd(fadd(pz, fmul(qw, uz)), cz_), T[6]);
    const float invzc = (float)(1.0 / (double)zc);
    if (invzc < 0) return false;
    u = fadd(fdiv(fmul(P.cam[0], xc), zc), P.cam[2]);   // Pinhole::project, Pinhole.cpp:43-49
    v = fadd(fdiv(fmul(P.cam[1], yc), zc), P.cam[3]);
    if (u < P.minX || u > P.maxX) return false;
    if (v < P.minY || v > P.maxY) return false;
    if (!(u == u) || !(v == v)) return false;   // NaN projections (zc == 0) fail every comparison above in the reference too
    const int oct = P.octave[o];
    r = fmul(P.th, P.scaleFactors[oct]);
    minL = oct - 1; maxL = oct + 1;
    return true;
}

// pass 1: warp per map point
constexpr int MC_NT = 256;
__global__ void __launch_bounds__(MC_NT) match_candidates_kernel(MatchParams P) {
    const int f = blockIdx.y;
    const int i = blockIdx.x * (MC_NT / 32) + (threadIdx.x >> 5);
    const int M = min(P.nM[f], P.mcap);
    if (i >= M) return;
    const int lane = threadIdx.x & 31;
    const size_t o = (size_t)f * P.mcap + i;
    float u = 0, v = 0, r = 0; int minL = 0, maxL = 0;
    const bool ok = make_query(P, f, i, u, v, r, minL, maxL);
    int4 ri = make_int4(-1, -1, -1, -1), rd = make_int4(256, 256, 256, 256);
    if (ok) {
        uint32_t mpd[8];
        load_mp_desc(P, f, i, mpd);
        const int* match = P.match + (size_t)f * P.kcap;
        const uint8_t* claimed = P.claimed + (size_t)f * P.kcap;
        const bool reset = P.resetState != 0;
        const TopK<4> t = scan_candidates<4>(P, f, u, v, r, minL, maxL, mpd, [&](int idx) { return !reset && match[idx] >= 0 && claimed[idx]; });
        ri = make_int4(t.i[0], t.i[1], t.i[2], t.i[3]);
        rd = make_int4(t.i[0] >= 0 ? (int)(t.k[0] >> 40) : 256, t.i[1] >= 0 ? (int)(t.k[1] >> 40) : 256,
                       t.i[2] >= 0 ? (int)(t.k[2] >> 40) : 256, t.i[3] >= 0 ? (int)(t.k[3] >> 40) : 256);
    }
    if (lane == 0) {
        P.resultIdx[o] = ri; P.resultDist[o] = rd;
        P.query[o] = make_float4(u, v, r, __int_as_float((minL + 1) | ((maxL + 1) << 8) | ((ok ? 1 : 0) << 16)));
    }
}

// pass 2: one warp per stream, map points in index order
__global__ void __launch_bounds__(32) match_commit_kernel(MatchParams P) {
    extern __shared__ uint32_t s_bits[];   // claim bitset, ceil(kcap/32) words
    __shared__ int s_hist[HISTO_LENGTH];
    const int f = blockIdx.x, lane = threadIdx.x;
    const int K = min(P.nK[f], P.kcap), M = min(P.nM[f], P.mcap);
    int* match = P.match + (size_t)f * P.kcap;
    uint8_t* claimed = P.claimed + (size_t)f * P.kcap;
    const OrbKeyPoint* kps = P.kps + (size_t)f * P.kcap;
    uint8_t* evBin = P.evBin + (size_t)f * P.mcap;
    uint16_t* evIdx = P.evIdx + (size_t)f * P.mcap;
    const int nw = (P.kcap + 31) / 32;
    for (int w = lane; w < nw; w += 32) {
        uint32_t bits = 0;
        for (int b = 0; b < 32; ++b) {
            const int idx = w * 32 + b;
            if (idx < K && match[idx] >= 0 && claimed[idx]) bits |= 1u << b;
        }
        s_bits[w] = bits;
    }
    if (lane < HISTO_LENGTH) s_hist[lane] = 0;
    __syncwarp();
    int nmatches = 0, nEvents = 0;
    const float factor = 1.0f / HISTO_LENGTH;
    const bool hist = P.mode == 1 && P.checkOri;
    auto rot_bin = [&](float lastAngle, float curAngle) {   // rotation histogram bin (:1779-1789)
        float rot = fsub(lastAngle, curAngle);
        if (rot < 0.0f) rot = fadd(rot, 360.0f);
        int bin = (int)roundf(fmul(rot, factor));
        if (bin == HISTO_LENGTH) bin = 0;
        return min(max(bin, 0), HISTO_LENGTH - 1);
    };
    auto is_claimed = [&](int idx) { return (bool)((s_bits[idx >> 5] >> (idx & 31)) & 1u); };
    for (int base = 0; base < M; base += 32) {
        // every lane prefetches everything the serial part needs for "its" map point (gathers run in parallel)
        const int i = base + lane;
        const size_t o = (size_t)f * P.mcap + i;
        int4 ri = make_int4(-1, -1, -1, -1), rd = make_int4(256, 256, 256, 256);
        int hasObs = 0;
        unsigned lv = 0, bn = 0;   // 4 x 8 bit: octave / histogram bin of the four candidates
        if (i < M) {
            ri = P.resultIdx[o]; rd = P.resultDist[o]; hasObs = P.hasObs[o];
            const float lastAngle = hist ? P.angle[o] : 0.f;
            const int ids[4] = {ri.x, ri.y, ri.z, ri.w};
#pragma unroll
            for (int c = 0; c < 4; ++c) {
                if (ids[c] >= 0) {
                    const OrbKeyPoint kb = kps[ids[c]];
                    lv |= (unsigned)(kb.octave & 0xff) << (8 * c);
                    if (hist) bn |= (unsigned)rot_bin(lastAngle, kb.angle) << (8 * c);
                }
            }
        }
        const int cnt = min(32, M - base);
        for (int j = 0; j < cnt; ++j) {
            const int c0 = __shfl_sync(0xffffffffu, ri.x, j);
            if (c0 < 0) continue;                                     // no candidate at all
            const int ids[4] = {c0, __shfl_sync(0xffffffffu, ri.y, j), __shfl_sync(0xffffffffu, ri.z, j), __shfl_sync(0xffffffffu, ri.w, j)};
            const int ds[4] = {__shfl_sync(0xffffffffu, rd.x, j), __shfl_sync(0xffffffffu, rd.y, j), __shfl_sync(0xffffffffu, rd.z, j),
                               __shfl_sync(0xffffffffu, rd.w, j)};
            const int obs = __shfl_sync(0xffffffffu, hasObs, j);
            const unsigned lvs = __shfl_sync(0xffffffffu, lv, j), bns = __shfl_sync(0xffffffffu, bn, j);
            // first (and, for the local-map ratio test, second) candidate that is still unclaimed
            int bIdx = -1, bDist = 256, sIdx = -1, sDist = 256, lv1 = -1, lv2 = -1, bnb = 0;
            bool exhausted = false;                                   // true: the list ended before 4 entries
#pragma unroll
            for (int c = 0; c < 4; ++c) {
                if (ids[c] < 0) { exhausted = true; continue; }
                if (is_claimed(ids[c])) continue;
                if (bIdx < 0) { bIdx = ids[c]; bDist = ds[c]; lv1 = (int)((lvs >> (8 * c)) & 0xff); bnb = (int)((bns >> (8 * c)) & 0xff); }
                else if (sIdx < 0) { sIdx = ids[c]; sDist = ds[c]; lv2 = (int)((lvs >> (8 * c)) & 0xff); }
            }
            const bool need2 = P.mode == 0;
            if (!exhausted && (bIdx < 0 || (need2 && sIdx < 0))) {   // the four were not enough: rescan against the current claims
                const size_t oj = (size_t)f * P.mcap + base + j;
                const float4 q = P.query[oj];
                const int bits = __float_as_int(q.w);
                uint32_t mpd[8];
                load_mp_desc(P, f, base + j, mpd);
                const Top2 t = scan_candidates<2>(P, f, q.x, q.y, q.z, (bits & 0xff) - 1, ((bits >> 8) & 0xff) - 1, mpd, is_claimed);
                bIdx = t.i[0]; bDist = t.i[0] >= 0 ? (int)(t.k[0] >> 40) : 256;
                sIdx = t.i[1]; sDist = t.i[1] >= 0 ? (int)(t.k[1] >> 40) : 256;
                if (bIdx >= 0) {
                    lv1 = kps[bIdx].octave; lv2 = sIdx >= 0 ? kps[sIdx].octave : -1;
                    if (hist) bnb = rot_bin(P.angle[oj], kps[bIdx].angle);
                }
            }
            if (bIdx < 0) continue;
            if (bDist > TH_HIGH) continue;
            if (P.mode == 0) {   // ratio test only when best and second come from the same level (:123-128)
                if (lv1 == lv2 && (float)bDist > fmul(P.nnratio, (float)sDist)) continue;
            }
            if (lane == 0) {
                match[bIdx] = base + j;
                claimed[bIdx] = (uint8_t)obs;
                if (obs) s_bits[bIdx >> 5] |= 1u << (bIdx & 31);
                else s_bits[bIdx >> 5] &= ~(1u << (bIdx & 31));
                if (hist) { evBin[nEvents] = (uint8_t)bnb; evIdx[nEvents] = (uint16_t)bIdx; s_hist[bnb]++; }
            }
            ++nmatches; ++nEvents;
            __syncwarp();
        }
    }
    if (P.mode == 1 && P.checkOri) {   // ComputeThreeMaxima (:2012-2053) + removal (:1868-1884)
        __syncwarp();
        int ind1 = -1, ind2 = -1, ind3 = -1;
        if (lane == 0) {
            int max1 = 0, max2 = 0, max3 = 0;
            for (int b = 0; b < HISTO_LENGTH; ++b) {
                const int s = s_hist[b];
                if (s > max1) { max3 = max2; max2 = max1; max1 = s; ind3 = ind2; ind2 = ind1; ind1 = b; }
                else if (s > max2) { max3 = max2; max2 = s; ind3 = ind2; ind2 = b; }
                else if (s > max3) { max3 = s; ind3 = b; }
            }
            if ((float)max2 < fmul(0.1f, (float)max1)) { ind2 = -1; ind3 = -1; }
            else if ((float)max3 < fmul(0.1f, (float)max1)) { ind3 = -1; }
        }
        ind1 = __shfl_sync(0xffffffffu, ind1, 0); ind2 = __shfl_sync(0xffffffffu, ind2, 0); ind3 = __shfl_sync(0xffffffffu, ind3, 0);
        int removed = 0;
        for (int e = lane; e < nEvents; e += 32) {
            const int b = evBin[e];
            if (b != ind1 && b != ind2 && b != ind3) { match[evIdx[e]] = -1; claimed[evIdx[e]] = 0; ++removed; }
        }
#pragma unroll
        for (int o = 16; o; o >>= 1) removed += __shfl_xor_sync(0xffffffffu, removed, o);
        nmatches -= removed;
    }
    if (lane == 0) P.nmatches[f] = nmatches;
}

// cv::BFMatcher(NORM_HAMMING).knnMatch(k=2): warp per query, lanes over train rows
constexpr int BF_NT = 256;
__global__ void __launch_bounds__(BF_NT) bf_knn2_kernel(const uint8_t* __restrict__ q, int Q, const uint8_t* __restrict__ t, int T,
                                                        int* __restrict__ idx, int* __restrict__ dist) {
    const int qi = blockIdx.x * (BF_NT / 32) + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (qi >= Q) return;
    const uint4* qp = reinterpret_cast<const uint4*>(q + (size_t)qi * 32);
    const uint4 a = __ldg(qp), b = __ldg(qp + 1);
    const uint32_t qd[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
    Top2 tt; tt.init();
    for (int j = lane; j < T; j += 32) {
        const uint4* tp = reinterpret_cast<const uint4*>(t + (size_t)j * 32);
        const uint4 c = __ldg(tp), d = __ldg(tp + 1);
        const uint32_t td[8] = {c.x, c.y, c.z, c.w, d.x, d.y, d.z, d.w};
        tt.insert(((unsigned long long)hamming256(qd, td) << 32) | (unsigned)j, j);
    }
    tt.warp_merge();
    if (lane == 0) {
        idx[2 * qi] = tt.i[0]; idx[2 * qi + 1] = tt.i[1];
        dist[2 * qi] = tt.i[0] >= 0 ? (int)(tt.k[0] >> 32) : -1;
        dist[2 * qi + 1] = tt.i[1] >= 0 ? (int)(tt.k[1] >> 32) : -1;
    }
}

__global__ void hamming_pairs_kernel(const uint8_t* __restrict__ a, const uint8_t* __restrict__ b, int n, int* __restrict__ out) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const uint4* pa = reinterpret_cast<const uint4*>(a + (size_t)i * 32);
    const uint4* pb = reinterpret_cast<const uint4*>(b + (size_t)i * 32);
    const uint4 a0 = pa[0], a1 = pa[1], b0 = pb[0], b1 = pb[1];
    const uint32_t x[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
    const uint32_t y[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
    out[i] = hamming256(x, y);
}

struct Matcher {
    int device, maxBatch, kcap, mcap;
    cudaStream_t stream = nullptr;
    // scratch
    int* d_cellStart = nullptr; uint16_t* d_cellIdx = nullptr; float4* d_query = nullptr; int4 *d_resultIdx = nullptr, *d_resultDist = nullptr;
    uint8_t* d_evBin = nullptr; uint16_t* d_evIdx = nullptr; int* d_status = nullptr;
    // staging for the host entry points (batch = 1) -- one arena
    uint8_t* d_arena = nullptr; size_t arenaBytes = 0;
    uint8_t* h_arena = nullptr;
    uint8_t* d_batch = nullptr; size_t batchBytes = 0;   // device staging for the host-pointer batch entry point
    int launches = 0;

    ~Matcher() {
        cudaSetDevice(device);
        void* ptrs[] = {d_cellStart, d_cellIdx, d_query, d_resultIdx, d_resultDist, d_evBin, d_evIdx, d_status, d_arena, d_batch};
        for (void* p : ptrs) if (p) cudaFree(p);
        if (h_arena) cudaFreeHost(h_arena);
        if (stream) cudaStreamDestroy(stream);
    }
    int init() {
        CK(cudaSetDevice(device));
        cudaDeviceProp prop;
        CK(cudaGetDeviceProperties(&prop, device));
        if (prop.major < 10) { set_error("device is not sm_100+ (Blackwell); this library has no other code path"); return ORB_ERR_CUDA; }
        const size_t B = maxBatch;
        CK(cudaMalloc(&d_cellStart, sizeof(int) * (GRID_CELLS + 1) * B));
        CK(cudaMalloc(&d_cellIdx, sizeof(uint16_t) * kcap * B));
        CK(cudaMalloc(&d_query, sizeof(float4) * mcap * B));
        CK(cudaMalloc(&d_resultIdx, sizeof(int4) * mcap * B));
        CK(cudaMalloc(&d_resultDist, sizeof(int4) * mcap * B));
        CK(cudaMalloc(&d_evBin, mcap * B));
        CK(cudaMalloc(&d_evIdx, sizeof(uint16_t) * mcap * B));
        CK(cudaMalloc(&d_status, sizeof(int) * B));
        arenaBytes = (size_t)kcap * (28 + 32 + 4 + 1) + (size_t)mcap * (32 + 12 + 4 * 6 + 4) + 4096 + 64 * 64;
        arenaBytes = (arenaBytes + 255) & ~(size_t)255;
        const size_t bf = (size_t)(kcap + mcap) * 32 + (size_t)std::max(kcap, mcap) * 16 + 4096;
        arenaBytes = std::max(arenaBytes, bf);
        CK(cudaMalloc(&d_arena, arenaBytes));
        {
            orbx::ScopedGpuAffinity numaLocal(device);     // pinned pages on the GPU's NUMA node (host_affinity.h)
            CK(cudaMallocHost(&h_arena, arenaBytes));
        }
        batchBytes = B * ((size_t)kcap * (28 + 32 + 4 + 1) + (size_t)mcap * (1 + 12 + 4 + 4 + 1 + 32) + 28 + 16 + 13 * 256) + 4096;
        CK(cudaMalloc(&d_batch, batchBytes));
        CK(cudaStreamCreateWithFlags(&stream, cudaStreamNonBlocking));
        CK(cudaFuncSetAttribute(match_commit_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 48 * 1024));
        return ORB_OK;
    }
    int run(MatchParams& P, cudaStream_t st) {
        P.gridWInv = (float)GRID_COLS / (P.maxX - P.minX);   // src/Frame.cc:342-343
        P.gridHInv = (float)GRID_ROWS / (P.maxY - P.minY);
        P.cellStart = d_cellStart; P.cellIdx = d_cellIdx; P.query = d_query; P.resultIdx = d_resultIdx; P.resultDist = d_resultDist;
        P.evBin = d_evBin; P.evIdx = d_evIdx; P.status = d_status;
        launches = 0;
        grid_build_kernel<<<P.batch, GB_NT, 0, st>>>(P);
        match_candidates_kernel<<<dim3((P.mcap + MC_NT / 32 - 1) / (MC_NT / 32), P.batch), MC_NT, 0, st>>>(P);
        const size_t sm = sizeof(uint32_t) * ((P.kcap + 31) / 32);
        if (sm > 48 * 1024) { set_error("kcap too large for the claim bitset"); return ORB_ERR_ARG; }
        match_commit_kernel<<<P.batch, 32, sm, st>>>(P);
        launches = 3;
        CK(cudaGetLastError());
        return ORB_OK;
    }
};

// bump allocator over the paired host/device arenas
// ---- Frame::isInFrustum + MapPoint::PredictScale, thread per map point (src/Frame.cc:512-574, src/MapPoint.cc:531-546).
//      Every float operation is individually rounded in the order fixed by the oracle; std::log(float) is glibc's logf
//      (exact_math.h: logf_glibc). ----
struct FrustumParams {
    int M;
    const float *P, *N, *minD, *maxD, *maxRaw;
    float R[9], t[3], Ow[3], cam[4], minX, minY, maxX, maxY, mbf, logSF, cosLimit;
    int nLevels;
    uint8_t* inView; float *projX, *projY, *projXR, *depth, *viewCos; int* level;
};
__global__ void frustum_project_kernel(FrustumParams Q) {
    using namespace orbx;
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= Q.M) return;
    uint8_t in = 0; float px = -1.f, py = -1.f, pxr = 0.f, dep = 0.f, vc = 0.f; int lvl = -1;
    const float X = Q.P[3 * i], Y = Q.P[3 * i + 1], Z = Q.P[3 * i + 2];
    const float xc = fadd(fadd(fadd(fmul(Q.R[0], X), fmul(Q.R[1], Y)), fmul(Q.R[2], Z)), Q.t[0]);
    const float yc = fadd(fadd(fadd(fmul(Q.R[3], X), fmul(Q.R[4], Y)), fmul(Q.R[5], Z)), Q.t[1]);
    const float zc = fadd(fadd(fadd(fmul(Q.R[6], X), fmul(Q.R[7], Y)), fmul(Q.R[8], Z)), Q.t[2]);
    const float pcDist = __fsqrt_rn(fadd(fadd(fmul(xc, xc), fmul(yc, yc)), fmul(zc, zc)));
    const float invz = fdiv(1.0f, zc);
    if (!(zc < 0.0f)) {
        const float u = fadd(fdiv(fmul(Q.cam[0], xc), zc), Q.cam[2]);
        const float v = fadd(fdiv(fmul(Q.cam[1], yc), zc), Q.cam[3]);
        if (!(u < Q.minX || u > Q.maxX) && !(v < Q.minY || v > Q.maxY)) {
            px = u; py = v;
            const float ox = fsub(X, Q.Ow[0]), oy = fsub(Y, Q.Ow[1]), oz = fsub(Z, Q.Ow[2]);
            const float dist = __fsqrt_rn(fadd(fadd(fmul(ox, ox), fmul(oy, oy)), fmul(oz, oz)));
            if (!(dist < Q.minD[i] || dist > Q.maxD[i])) {
                const float c = fdiv(fadd(fadd(fmul(ox, Q.N[3 * i]), fmul(oy, Q.N[3 * i + 1])), fmul(oz, Q.N[3 * i + 2])), dist);
                if (!(c < Q.cosLimit)) {
                    const float ratio = fdiv(Q.maxRaw[i], dist);
                    int n = (int)ceilf(fdiv(logf_glibc(ratio), Q.logSF));
                    if (n < 0) n = 0; else if (n >= Q.nLevels) n = Q.nLevels - 1;
                    in = 1; pxr = __fmaf_rn(-Q.mbf, invz, u) /* one FMA in the reference build, see oracle */; dep = pcDist; lvl = n; vc = c;
                }
            }
        }
    }
    Q.inView[i] = in; Q.projX[i] = px; Q.projY[i] = py; Q.projXR[i] = pxr; Q.depth[i] = dep; Q.level[i] = lvl; Q.viewCos[i] = vc;
}

struct Arena {
    uint8_t *h, *d; size_t off = 0, cap;
    Arena(uint8_t* h_, uint8_t* d_, size_t c) : h(h_), d(d_), cap(c) {}
    template <class T> bool put(const T* src, size_t n, const T** dptr) {
        off = (off + 63) & ~(size_t)63;
        const size_t bytes = n * sizeof(T);
        if (off + bytes > cap) return false;
        if (src && bytes) memcpy(h + off, src, bytes);
        *dptr = reinterpret_cast<const T*>(d + off);
        off += bytes;
        return true;
    }
};

}  // namespace orbm

using namespace orbm;

struct orbm_handle { Matcher m; };

extern "C" {

int orbm_create(orbm_handle** out, int max_batch, int max_keypoints, int max_mappoints, int device) {
    if (!out || max_batch < 1 || max_keypoints < 1 || max_mappoints < 1 || max_keypoints > 65535) { set_error("orbm_create: bad argument"); return ORB_ERR_ARG; }
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) { set_error("no CUDA device (this library has no CPU path)"); return ORB_ERR_CUDA; }
    if (device < 0 || device >= ndev) { set_error("orbm_create: bad device index"); return ORB_ERR_ARG; }
    orbm_handle* h = new orbm_handle();
    h->m.device = device; h->m.maxBatch = max_batch; h->m.kcap = max_keypoints; h->m.mcap = max_mappoints;
    int rc = h->m.init();
    if (rc) { delete h; return rc; }
    *out = h;
    return ORB_OK;
}

void orbm_destroy(orbm_handle* h) { delete h; }
int orbm_last_launch_count(const orbm_handle* h) { return h ? h->m.launches : ORB_ERR_ARG; }

static int stage_frame(Matcher& m, Arena& A, const OrbmFrame* fr, MatchParams& P, const int32_t* match, const uint8_t* claimed,
                       int** dMatch, uint8_t** dClaimed, int** dN) {
    if (!fr || fr->K < 0 || fr->K > m.kcap || !fr->scaleFactors || fr->nlevels < 1 || !(fr->maxX > fr->minX) || !(fr->maxY > fr->minY)) {
        set_error("bad OrbmFrame (K > max_keypoints?)"); return ORB_ERR_ARG;
    }
    const int32_t* dm; const uint8_t* dc; const int* dn;
    int counts[2] = {fr->K, 0};
    bool ok = A.put(fr->keypoints, (size_t)fr->K, &P.kps) && A.put(fr->descriptors, (size_t)fr->K * 32, &P.desc) &&
              A.put(fr->scaleFactors, (size_t)fr->nlevels, &P.scaleFactors) && A.put(match, (size_t)fr->K, &dm) &&
              A.put(claimed, (size_t)fr->K, &dc) && A.put(counts, (size_t)2, &dn);
    if (!ok) { set_error("matcher staging arena too small"); return ORB_ERR_CAPACITY; }
    P.kcap = m.kcap; P.nlevels = fr->nlevels; P.batch = 1;
    P.minX = fr->minX; P.minY = fr->minY; P.maxX = fr->maxX; P.maxY = fr->maxY;
    *dMatch = const_cast<int*>(dm); *dClaimed = const_cast<uint8_t*>(dc); *dN = const_cast<int*>(dn);
    P.nK = dn;
    return ORB_OK;
}

static int finish_host(Matcher& m, Arena& A, MatchParams& P, int K, int* dMatch, uint8_t* dClaimed, int* dN, int32_t* match,
                       uint8_t* claimed, int* nmatches) {
    cudaStream_t st = m.stream;
    CK(cudaMemcpyAsync(m.d_arena, m.h_arena, A.off, cudaMemcpyHostToDevice, st));
    P.match = dMatch; P.claimed = dClaimed; P.nmatches = dN + 1;
    int rc = m.run(P, st);
    if (rc) return rc;
    // results come back through the same arena offsets
    const size_t oM = (uint8_t*)dMatch - m.d_arena, oC = (uint8_t*)dClaimed - m.d_arena, oN = (uint8_t*)dN - m.d_arena;
    CK(cudaMemcpyAsync(m.h_arena + oM, dMatch, sizeof(int) * K, cudaMemcpyDeviceToHost, st));
    CK(cudaMemcpyAsync(m.h_arena + oC, dClaimed, K, cudaMemcpyDeviceToHost, st));
    CK(cudaMemcpyAsync(m.h_arena + oN, dN, 2 * sizeof(int), cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    memcpy(match, m.h_arena + oM, sizeof(int) * K);
    memcpy(claimed, m.h_arena + oC, K);
    *nmatches = ((int*)(m.h_arena + oN))[1];
    return ORB_OK;
}

int orbm_search_local_map(orbm_handle* h, const OrbmFrame* fr, const OrbmLocalPoints* pts, float th, float nnratio, int bFar,
                          float thFar, int32_t* match, uint8_t* claimed, int* nmatches) {
    if (!h || !pts || !match || !claimed || !nmatches || pts->M < 0 || pts->M > h->m.mcap) { set_error("orbm_search_local_map: bad argument"); return ORB_ERR_ARG; }
    Matcher& m = h->m;
    CK(cudaSetDevice(m.device));
    MatchParams P; memset(&P, 0, sizeof(P));
    Arena A(m.h_arena, m.d_arena, m.arenaBytes);
    int *dMatch, *dN; uint8_t* dClaimed;
    int rc = stage_frame(m, A, fr, P, match, claimed, &dMatch, &dClaimed, &dN);
    if (rc) return rc;
    ((int*)(m.h_arena + ((uint8_t*)dN - m.d_arena)))[1] = 0;
    const size_t M = pts->M;
    int mcount[1] = {pts->M};
    bool ok = A.put(pts->inView, M, &P.inView) && A.put(pts->bad, M, &P.bad) && A.put(pts->depth, M, &P.depth) &&
              A.put(pts->projX, M, &P.projX) && A.put(pts->projY, M, &P.projY) && A.put(pts->level, M, &P.level) &&
              A.put(pts->viewCos, M, &P.viewCos) && A.put(pts->hasObs, M, &P.hasObs) && A.put(pts->descriptors, M * 32, &P.mpDesc) &&
              A.put(mcount, (size_t)1, &P.nM);
    if (!ok) { set_error("matcher staging arena too small"); return ORB_ERR_CAPACITY; }
    for (size_t i = 0; i < M; ++i)
        if (pts->inView[i] && !pts->bad[i] && (pts->level[i] < 0 || pts->level[i] >= fr->nlevels)) { set_error("map point level out of range"); return ORB_ERR_ARG; }
    P.mcap = m.mcap; P.mode = 0; P.th = th; P.nnratio = nnratio; P.bFar = bFar; P.thFar = thFar;
    return finish_host(m, A, P, fr->K, dMatch, dClaimed, dN, match, claimed, nmatches);
}

int orbm_search_last_frame(orbm_handle* h, const OrbmFrame* fr, const OrbmLastFrame* last, const float* Tcw7, const float* cam4,
                           float th, int checkOri, int32_t* match, uint8_t* claimed, int* nmatches) {
    if (!h || !last || !Tcw7 || !cam4 || !match || !claimed || !nmatches || last->M < 0 || last->M > h->m.mcap) { set_error("orbm_search_last_frame: bad argument"); return ORB_ERR_ARG; }
    Matcher& m = h->m;
    CK(cudaSetDevice(m.device));
    MatchParams P; memset(&P, 0, sizeof(P));
    Arena A(m.h_arena, m.d_arena, m.arenaBytes);
    int *dMatch, *dN; uint8_t* dClaimed;
    int rc = stage_frame(m, A, fr, P, match, claimed, &dMatch, &dClaimed, &dN);
    if (rc) return rc;
    const size_t M = last->M;
    int mcount[1] = {last->M};
    bool ok = A.put(last->valid, M, &P.valid) && A.put(last->xyz, M * 3, &P.xyz) && A.put(last->octave, M, &P.octave) &&
              A.put(last->angle, M, &P.angle) && A.put(last->hasObs, M, &P.hasObs) && A.put(last->descriptors, M * 32, &P.mpDesc) &&
              A.put(Tcw7, (size_t)7, &P.Tcw7) && A.put(mcount, (size_t)1, &P.nM);
    if (!ok) { set_error("matcher staging arena too small"); return ORB_ERR_CAPACITY; }
    for (size_t i = 0; i < M; ++i)
        if (last->valid[i] && (last->octave[i] < 0 || last->octave[i] >= fr->nlevels)) { set_error("last-frame octave out of range"); return ORB_ERR_ARG; }
    memcpy(P.cam, cam4, sizeof(float) * 4);
    P.mcap = m.mcap; P.mode = 1; P.th = th; P.checkOri = checkOri;
    return finish_host(m, A, P, fr->K, dMatch, dClaimed, dN, match, claimed, nmatches);
}

int orbm_search_last_frame_batch_device(orbm_handle* h, const OrbmBatchDevice* in, float th, int checkOri, int32_t* d_match,
                                        uint8_t* d_claimed, int32_t* d_nmatches, void* stream) {
    if (!h || !in || !d_match || !d_claimed || !d_nmatches || in->batch < 1 || in->batch > h->m.maxBatch || in->kcap > 65535 ||
        in->kcap > h->m.kcap || in->mcap > h->m.mcap) { set_error("orbm_search_last_frame_batch_device: bad argument"); return ORB_ERR_ARG; }
    Matcher& m = h->m;
    CK(cudaSetDevice(m.device));
    MatchParams P; memset(&P, 0, sizeof(P));
    P.batch = in->batch; P.kcap = in->kcap; P.mcap = in->mcap; P.nlevels = in->nlevels; P.mode = 1;
    P.kps = in->kps; P.desc = in->desc; P.nK = in->nK;
    P.minX = in->minX; P.minY = in->minY; P.maxX = in->maxX; P.maxY = in->maxY;
    P.scaleFactors = in->scaleFactors; P.nM = in->nM;
    P.valid = in->valid; P.xyz = in->xyz; P.octave = in->octave; P.angle = in->angle; P.hasObs = in->hasObs; P.mpDesc = in->mpDesc;
    P.Tcw7 = in->Tcw7; memcpy(P.cam, in->cam, sizeof(float) * 4);
    P.th = th; P.checkOri = checkOri; P.resetState = in->resetState;
    P.match = d_match; P.claimed = d_claimed; P.nmatches = d_nmatches;
    return m.run(P, (cudaStream_t)stream);
}

int orbm_search_last_frame_batch(orbm_handle* h, const OrbmBatchDevice* in, float th, int checkOri, int32_t* match, uint8_t* claimed,
                                 int32_t* nmatches) {
    if (!h || !in || !match || !claimed || !nmatches || in->batch < 1 || in->batch > h->m.maxBatch || in->kcap > h->m.kcap ||
        in->mcap > h->m.mcap || in->nlevels < 1) { set_error("orbm_search_last_frame_batch: bad argument"); return ORB_ERR_ARG; }
    Matcher& m = h->m;
    CK(cudaSetDevice(m.device));
    cudaStream_t st = m.stream;
    const size_t B = in->batch, K = in->kcap, M = in->mcap;
    size_t off = 0;
    OrbmBatchDevice d = *in;
    auto up = [&](const void* src, size_t bytes, const void** dst) -> int {
        off = (off + 255) & ~(size_t)255;
        if (off + bytes > m.batchBytes) { set_error("orbm_search_last_frame_batch: staging too small"); return ORB_ERR_CAPACITY; }
        if (src) { cudaError_t e = cudaMemcpyAsync(m.d_batch + off, src, bytes, cudaMemcpyHostToDevice, st); if (e != cudaSuccess) { set_error(cudaGetErrorString(e)); return ORB_ERR_CUDA; } }
        *dst = m.d_batch + off; off += bytes;
        return ORB_OK;
    };
    int rc;
    const void *dm, *dc, *dn;
#define UP(field, bytes) if ((rc = up(in->field, bytes, (const void**)&d.field))) return rc
    UP(kps, B * K * 28); UP(desc, B * K * 32); UP(nK, B * 4); UP(scaleFactors, (size_t)in->nlevels * 4); UP(nM, B * 4);
    UP(valid, B * M); UP(xyz, B * M * 12); UP(octave, B * M * 4); UP(angle, B * M * 4); UP(hasObs, B * M); UP(mpDesc, B * M * 32); UP(Tcw7, B * 28);
#undef UP
    if ((rc = up(match, B * K * 4, &dm)) || (rc = up(claimed, B * K, &dc)) || (rc = up(nullptr, B * 4, &dn))) return rc;
    rc = orbm_search_last_frame_batch_device(h, &d, th, checkOri, (int32_t*)dm, (uint8_t*)dc, (int32_t*)dn, st);
    if (rc) return rc;
    CK(cudaMemcpyAsync(match, dm, B * K * 4, cudaMemcpyDeviceToHost, st));
    CK(cudaMemcpyAsync(claimed, dc, B * K, cudaMemcpyDeviceToHost, st));
    CK(cudaMemcpyAsync(nmatches, dn, B * 4, cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    return ORB_OK;
}

int orbm_frustum_project(orbm_handle* h, const OrbmFrustumIn* in, uint8_t* inView, float* projX, float* projY, float* projXR,
                         float* depth, int32_t* level, float* viewCos) {
    if (!h || !in || in->M < 0 || !inView || !projX || !projY || !projXR || !depth || !level || !viewCos ||
        (in->M && (!in->worldPos || !in->normal || !in->minDistInv || !in->maxDistInv || !in->maxDistance)) || in->nScaleLevels < 1) {
        set_error("orbm_frustum_project: bad argument"); return ORB_ERR_ARG;
    }
    Matcher& m = h->m;
    CK(cudaSetDevice(m.device));
    const size_t M = in->M;
    if (M == 0) return ORB_OK;
    Arena A(m.h_arena, m.d_arena, m.arenaBytes);
    FrustumParams Q; memset(&Q, 0, sizeof(Q));
    Q.M = in->M;
    bool ok = A.put(in->worldPos, M * 3, &Q.P) && A.put(in->normal, M * 3, &Q.N) && A.put(in->minDistInv, M, &Q.minD) &&
              A.put(in->maxDistInv, M, &Q.maxD) && A.put(in->maxDistance, M, &Q.maxRaw);
    const size_t inBytes = A.off;
    const uint8_t* dIn; const float *dX, *dY, *dXR, *dD, *dC; const int* dL;
    ok = ok && A.put((const float*)nullptr, M, &dX) && A.put((const float*)nullptr, M, &dY) && A.put((const float*)nullptr, M, &dXR) &&
         A.put((const float*)nullptr, M, &dD) && A.put((const float*)nullptr, M, &dC) && A.put((const int*)nullptr, M, &dL) &&
         A.put((const uint8_t*)nullptr, M, &dIn);
    if (!ok) { set_error("orbm_frustum_project: more map points than the staging arena holds (max_mappoints)"); return ORB_ERR_CAPACITY; }
    memcpy(Q.R, in->Rcw, sizeof(Q.R)); memcpy(Q.t, in->tcw, sizeof(Q.t)); memcpy(Q.Ow, in->Ow, sizeof(Q.Ow)); memcpy(Q.cam, in->cam, sizeof(Q.cam));
    Q.minX = in->minX; Q.minY = in->minY; Q.maxX = in->maxX; Q.maxY = in->maxY; Q.mbf = in->mbf; Q.logSF = in->logScaleFactor;
    Q.cosLimit = in->viewingCosLimit; Q.nLevels = in->nScaleLevels;
    Q.inView = const_cast<uint8_t*>(dIn); Q.projX = const_cast<float*>(dX); Q.projY = const_cast<float*>(dY); Q.projXR = const_cast<float*>(dXR);
    Q.depth = const_cast<float*>(dD); Q.viewCos = const_cast<float*>(dC); Q.level = const_cast<int*>(dL);
    cudaStream_t st = m.stream;
    CK(cudaMemcpyAsync(m.d_arena, m.h_arena, inBytes, cudaMemcpyHostToDevice, st));
    frustum_project_kernel<<<(unsigned)((M + 255) / 256), 256, 0, st>>>(Q);
    m.launches = 1;
    CK(cudaGetLastError());
    const size_t outOff = (const uint8_t*)dX - m.d_arena, outBytes = A.off - outOff;
    CK(cudaMemcpyAsync(m.h_arena + outOff, m.d_arena + outOff, outBytes, cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    auto hp = [&](const void* d) { return m.h_arena + ((const uint8_t*)d - m.d_arena); };
    memcpy(projX, hp(dX), 4 * M); memcpy(projY, hp(dY), 4 * M); memcpy(projXR, hp(dXR), 4 * M); memcpy(depth, hp(dD), 4 * M);
    memcpy(viewCos, hp(dC), 4 * M); memcpy(level, hp(dL), 4 * M); memcpy(inView, hp(dIn), M);
    return ORB_OK;
}

int orbm_bf_knn2(orbm_handle* h, const uint8_t* query, int Q, const uint8_t* train, int T, int32_t* idx, int32_t* dist) {
    if (!h || !query || !train || !idx || !dist || Q < 0 || T < 0) { set_error("orbm_bf_knn2: bad argument"); return ORB_ERR_ARG; }
    Matcher& m = h->m;
    CK(cudaSetDevice(m.device));
    if (Q == 0) return ORB_OK;
    Arena A(m.h_arena, m.d_arena, m.arenaBytes);
    const uint8_t *dq, *dt; const int *di, *dd;
    if (!(A.put(query, (size_t)Q * 32, &dq) && A.put(train, (size_t)T * 32, &dt))) { set_error("orbm_bf_knn2: more descriptors than max_keypoints + max_mappoints"); return ORB_ERR_CAPACITY; }
    const size_t inBytes = A.off;
    if (!(A.put((const int*)nullptr, (size_t)Q * 2, &di) && A.put((const int*)nullptr, (size_t)Q * 2, &dd))) { set_error("orbm_bf_knn2: staging arena too small"); return ORB_ERR_CAPACITY; }
    cudaStream_t st = m.stream;
    CK(cudaMemcpyAsync(m.d_arena, m.h_arena, inBytes, cudaMemcpyHostToDevice, st));
    bf_knn2_kernel<<<(Q + BF_NT / 32 - 1) / (BF_NT / 32), BF_NT, 0, st>>>(dq, Q, dt, T, const_cast<int*>(di), const_cast<int*>(dd));
    m.launches = 1;
    CK(cudaGetLastError());
    CK(cudaMemcpyAsync(idx, di, sizeof(int) * 2 * Q, cudaMemcpyDeviceToHost, st));
    CK(cudaMemcpyAsync(dist, dd, sizeof(int) * 2 * Q, cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    return ORB_OK;
}

int orbm_descriptor_distance(orbm_handle* h, const uint8_t* a, const uint8_t* b, int n, int32_t* out) {
    if (!h || !a || !b || !out || n < 0) { set_error("orbm_descriptor_distance: bad argument"); return ORB_ERR_ARG; }
    Matcher& m = h->m;
    CK(cudaSetDevice(m.device));
    if (n == 0) return ORB_OK;
    Arena A(m.h_arena, m.d_arena, m.arenaBytes);
    const uint8_t *da, *db; const int* dout;
    if (!(A.put(a, (size_t)n * 32, &da) && A.put(b, (size_t)n * 32, &db))) { set_error("orbm_descriptor_distance: too many pairs for the staging arena"); return ORB_ERR_CAPACITY; }
    const size_t inBytes = A.off;
    if (!A.put((const int*)nullptr, (size_t)n, &dout)) { set_error("orbm_descriptor_distance: staging arena too small"); return ORB_ERR_CAPACITY; }
    cudaStream_t st = m.stream;
    CK(cudaMemcpyAsync(m.d_arena, m.h_arena, inBytes, cudaMemcpyHostToDevice, st));
    hamming_pairs_kernel<<<(n + 255) / 256, 256, 0, st>>>(da, db, n, const_cast<int*>(dout));
    m.launches = 1;
    CK(cudaGetLastError());
    CK(cudaMemcpyAsync(out, dout, sizeof(int) * n, cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    return ORB_OK;
}

}  // extern "C"
