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
#include <cooperative_groups.h>
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
namespace cg = cooperative_groups;

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
    int ks, ms, descInSmem;   // shared-memory capacities of match_frame_kernel (keypoints, map points) and where the descriptors live
    // scratch
    float4* query;       // [batch][mcap]  u, v, r, bits(minLevel+1 | (maxLevel+1) << 8 | valid << 16)
    int4 *resultIdx, *resultDist;   // [batch][mcap]  the four best (keypoint index, distance) of pass 1, ascending
    // in/out
    int* match; uint8_t* claimed; int* nmatches;
    int* status;
};

// ---------------------------------------------------------------------------------------------
// Fused per-frame matcher.  One thread-block CLUSTER per frame (stream); every CTA of the cluster holds the frame in
// shared memory -- the 64x48 grid of Frame::AssignFeaturesToGrid (src/Frame.cc:385-416, cell lists in keypoint order
// = the reference's push_back order), keypoint positions / octaves and (when they fit) the descriptors -- and scans
// its share of the map points (pass 1, warp per map point).  CTA 0 of the cluster then resolves the reference's
// sequential claim rule in parallel (pass 2) and applies the rotation-histogram filter.
//
// Pass 1: the four smallest (Hamming distance, Frame::GetFeaturesInArea enumeration order) keys of every map point
//         among the keypoints that are not claimed at entry.
// Pass 2: the reference visits map points in index order; a keypoint assigned to a map point with observations is
//         skipped by all later map points (src/ORBmatcher.cc:84-86, :1747-1749).  With s_i the choice of map point i,
//         s_i = first candidate of i that no j < i with observations chose.  That recurrence is solved by fixed-point
//         iteration: every map point re-decides in parallel against minClaimer[c] = min{ j : s_j = c, j has
//         observations }; after iteration t the first t map points are final, and an iteration without a change is the
//         sequential result (induction over i).  Contention is rare, so it converges in a handful of sweeps instead of
//         the M dependent steps of a serial walk.  A map point whose four candidates are all taken (while more exist)
//         is rescanned by a warp against the current claims -- removing keys ranked after the second never changes
//         (best, second), so this is exact.
// ---------------------------------------------------------------------------------------------
constexpr int MF_NT = 512;
constexpr unsigned NONE16 = 0xFFFFu;

struct FrameSmem {          // carved from dynamic shared memory (layout computed by frame_smem_bytes)
    int* cnt;               // [GRID_CELLS] build counters; reused as histogram / scratch afterwards
    uint16_t* cellStart;    // [GRID_CELLS + 1]
    uint16_t* cellIdx;      // [Ks]
    float *kx, *ky;         // [Ks]
    uint8_t* koct;          // [Ks]
    uint32_t* initBits;     // [(Ks + 31) / 32] claimed at entry
    int* minClaimer;        // [Ks]
    uint16_t* choice;       // [Ms] chosen keypoint of every map point (NONE16: none / not accepted)
    uint8_t* mflags;        // [Ms] bit0 = has observations, bits 2.. = rotation bin + 1 (0: none)
    uint16_t* queue;        // [Ms] map points that need a rescan in the current sweep
    uint32_t* desc;         // [Ks][8] or nullptr (descriptors stay in global memory)
};

__host__ __device__ inline size_t al16(size_t v) { return (v + 15) & ~(size_t)15; }
__host__ __device__ inline size_t frame_smem_bytes(int Ks, int Ms, bool descInSmem) {
    size_t b = 0;
    b += al16(sizeof(int) * GRID_CELLS);
    b += al16(sizeof(uint16_t) * (GRID_CELLS + 1));
    b += al16(sizeof(uint16_t) * Ks);
    b += 2 * al16(sizeof(float) * Ks);
    b += al16(Ks);
    b += al16(sizeof(uint32_t) * ((Ks + 31) / 32));
    b += al16(sizeof(int) * Ks);
    b += 2 * al16(sizeof(uint16_t) * Ms);
    b += al16(Ms);
    if (descInSmem) b += al16((size_t)32 * Ks);
    return b;
}
__device__ __forceinline__ FrameSmem carve_frame_smem(uint8_t* p, int Ks, int Ms, bool descInSmem) {
    FrameSmem S;
    S.cnt = (int*)p; p += al16(sizeof(int) * GRID_CELLS);
    S.cellStart = (uint16_t*)p; p += al16(sizeof(uint16_t) * (GRID_CELLS + 1));
    S.cellIdx = (uint16_t*)p; p += al16(sizeof(uint16_t) * Ks);
    S.kx = (float*)p; p += al16(sizeof(float) * Ks);
    S.ky = (float*)p; p += al16(sizeof(float) * Ks);
    S.koct = p; p += al16(Ks);
    S.initBits = (uint32_t*)p; p += al16(sizeof(uint32_t) * ((Ks + 31) / 32));
    S.minClaimer = (int*)p; p += al16(sizeof(int) * Ks);
    S.choice = (uint16_t*)p; p += al16(sizeof(uint16_t) * Ms);
    S.mflags = p; p += al16(Ms);
    S.queue = (uint16_t*)p; p += al16(sizeof(uint16_t) * Ms);
    S.desc = descInSmem ? (uint32_t*)p : nullptr;
    return S;
}

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
    // Afterwards every lane holds the K smallest keys of the whole warp.  Keys are unique (they carry the candidate's enumeration
    // rank), every lane's list is sorted, so the warp minimum is the minimum of the heads: K rounds of a 64-bit warp minimum (two
    // REDUX.MIN on the halves), the winning lane pops its head.
    __device__ __forceinline__ void warp_merge() {
        unsigned long long rk[K]; int ri[K];
#pragma unroll
        for (int j = 0; j < K; ++j) {
            const unsigned hi = (unsigned)(k[0] >> 32), lo = (unsigned)k[0];
            const unsigned mhi = __reduce_min_sync(0xffffffffu, hi);
            const unsigned mlo = __reduce_min_sync(0xffffffffu, hi == mhi ? lo : 0xffffffffu);
            const unsigned long long mk = ((unsigned long long)mhi << 32) | mlo;
            const bool won = k[0] == mk;
            const int src = __ffs(__ballot_sync(0xffffffffu, won)) - 1;
            rk[j] = mk; ri[j] = __shfl_sync(0xffffffffu, i[0], src);
            if (won && mk != ~0ull) {
#pragma unroll
                for (int q = 0; q + 1 < K; ++q) { k[q] = k[q + 1]; i[q] = i[q + 1]; }
                k[K - 1] = ~0ull; i[K - 1] = -1;
            }
        }
#pragma unroll
        for (int j = 0; j < K; ++j) { k[j] = rk[j]; i[j] = ri[j]; }
    }
};
typedef TopK<2> Top2;

// Warp-wide scan of the candidates of one query in Frame::GetFeaturesInArea order (src/Frame.cc:657-723): (ix, iy, position in
// the cell).  key = dist << 40 | cellRank << 20 | position; keeps the K smallest keys among keypoints that pass `usable`.
struct AnyDist { __device__ __forceinline__ bool operator()(int, int) const { return true; } };
template <int K, class UsableFn, class UsableDistFn = AnyDist>
__device__ __forceinline__ TopK<K> scan_candidates(const MatchParams& P, const FrameSmem& S, int f, float u, float v, float r, int minLevel,
                                                   int maxLevel, const uint32_t* mpd, UsableFn usable, UsableDistFn usableDist = AnyDist()) {
    const int lane = threadIdx.x & 31;
    TopK<K> t; t.init();
    const int nMinCellX = max(0, (int)floorf(fmul(fsub(fsub(u, P.minX), r), P.gridWInv)));
    const int nMaxCellX = min(GRID_COLS - 1, (int)ceilf(fmul(fadd(fsub(u, P.minX), r), P.gridWInv)));
    const int nMinCellY = max(0, (int)floorf(fmul(fsub(fsub(v, P.minY), r), P.gridHInv)));
    const int nMaxCellY = min(GRID_ROWS - 1, (int)ceilf(fmul(fadd(fsub(v, P.minY), r), P.gridHInv)));
    if (nMinCellX < GRID_COLS && nMaxCellX >= 0 && nMinCellY < GRID_ROWS && nMaxCellY >= 0) {
        const int ny = nMaxCellY - nMinCellY + 1, nx = nMaxCellX - nMinCellX + 1;
        const uint8_t* gdesc = P.desc + (size_t)f * P.kcap * 32;
        const bool bCheckLevels = (minLevel > 0) || (maxLevel >= 0);
        const unsigned nyMagic = 0xFFFFFFFFu / (unsigned)ny + 1u;          // c / ny by multiply-high (exact: c * ny < 2^32)
        for (int c = lane; c < nx * ny; c += 32) {
            const int cq = (int)__umulhi((unsigned)c, nyMagic);
            const int ix = nMinCellX + cq, iy = nMinCellY + (c - cq * ny);
            const int cell = ix * GRID_ROWS + iy;
            const int a = S.cellStart[cell], b = S.cellStart[cell + 1];
            for (int e = a; e < b; ++e) {
                const int idx = S.cellIdx[e];
                const int oct = S.koct[idx];
                if (bCheckLevels) {
                    if (oct < minLevel) continue;
                    if (maxLevel >= 0 && oct > maxLevel) continue;
                }
                const float dx = fsub(S.kx[idx], u), dy = fsub(S.ky[idx], v);
                if (!(fabsf(dx) < r && fabsf(dy) < r)) continue;
                if (!usable(idx)) continue;
                uint4 d0, d1;
                if (S.desc) {
                    const uint4* dp = reinterpret_cast<const uint4*>(S.desc + (size_t)idx * 8);
                    d0 = dp[0]; d1 = dp[1];
                } else {
                    const uint4* dp = reinterpret_cast<const uint4*>(gdesc + (size_t)idx * 32);
                    d0 = __ldg(dp); d1 = __ldg(dp + 1);
                }
                const uint32_t dd[8] = {d0.x, d0.y, d0.z, d0.w, d1.x, d1.y, d1.z, d1.w};
                const int dist = hamming256(mpd, dd);
                if (!usableDist(idx, dist)) continue;
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
        const int lvl = min(max(P.level[o], 0), P.nlevels - 1);   // the host entry points reject out-of-range levels; device callers are clamped
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
    const int oct = min(max(P.octave[o], 0), P.nlevels - 1);
    r = fmul(P.th, P.scaleFactors[oct]);
    minL = oct - 1; maxL = oct + 1;
    return true;
}

// rotation histogram bin (:1779-1789)
__device__ __forceinline__ int rot_bin(float lastAngle, float curAngle) {
    float rot = fsub(lastAngle, curAngle);
    if (rot < 0.0f) rot = fadd(rot, 360.0f);
    int bin = (int)roundf(fmul(rot, 1.0f / HISTO_LENGTH));
    if (bin == HISTO_LENGTH) bin = 0;
    return min(max(bin, 0), HISTO_LENGTH - 1);
}

// Frame::AssignFeaturesToGrid + PosInGrid (src/Frame.cc:385-416, :725-735): counting sort of the keypoints into the 64x48 grid held in
// shared memory, cell lists in increasing keypoint index (= the reference's push_back order); also stages positions, octaves and
// (when S.desc is set) the descriptors.  All threads of the CTA call; ends with a barrier.
__device__ void build_frame_smem(const FrameSmem& S, const OrbKeyPoint* kps, const uint8_t* gdesc, int K, float minX, float minY, float gridWInv,
                                 float gridHInv, int* s_warp) {
    const int tid = threadIdx.x;
    for (int c = tid; c < GRID_CELLS; c += MF_NT) S.cnt[c] = 0;
    __syncthreads();
    for (int i = tid; i < K; i += MF_NT) {
        const OrbKeyPoint kp = kps[i];
        S.kx[i] = kp.x; S.ky[i] = kp.y; S.koct[i] = (uint8_t)min(max(kp.octave, 0), 255);
        const int px = (int)roundf(fmul(fsub(kp.x, minX), gridWInv));
        const int py = (int)roundf(fmul(fsub(kp.y, minY), gridHInv));
        if (px >= 0 && px < GRID_COLS && py >= 0 && py < GRID_ROWS) atomicAdd(&S.cnt[px * GRID_ROWS + py], 1);
    }
    if (S.desc) {
        const uint4* src = reinterpret_cast<const uint4*>(gdesc);
        uint4* dst = reinterpret_cast<uint4*>(S.desc);
        for (int i = tid; i < 2 * K; i += MF_NT) dst[i] = __ldg(src + i);
    }
    __syncthreads();
    const int total = block_excl_scan(S.cnt, GRID_CELLS, s_warp);
    for (int c = tid; c < GRID_CELLS; c += MF_NT) S.cellStart[c] = (uint16_t)S.cnt[c];
    if (tid == 0) S.cellStart[GRID_CELLS] = (uint16_t)total;
    __syncthreads();
    for (int i = tid; i < K; i += MF_NT) {
        const int px = (int)roundf(fmul(fsub(S.kx[i], minX), gridWInv));
        const int py = (int)roundf(fmul(fsub(S.ky[i], minY), gridHInv));
        if (px >= 0 && px < GRID_COLS && py >= 0 && py < GRID_ROWS) S.cellIdx[atomicAdd(&S.cnt[px * GRID_ROWS + py], 1)] = (uint16_t)i;
    }
    __syncthreads();
    for (int c = tid; c < GRID_CELLS; c += MF_NT) {    // restore insertion order inside each cell (lists are short)
        const int a = S.cellStart[c], b = S.cnt[c];     // cnt now holds the end offset
        for (int i = a + 1; i < b; ++i) {
            const uint16_t v = S.cellIdx[i];
            int j = i - 1;
            while (j >= a && S.cellIdx[j] > v) { S.cellIdx[j + 1] = S.cellIdx[j]; --j; }
            S.cellIdx[j + 1] = v;
        }
    }
    __syncthreads();
}

__global__ void __launch_bounds__(MF_NT) match_frame_kernel(MatchParams P) {
    extern __shared__ __align__(16) uint8_t smem_raw[];
    __shared__ int s_warp[33];
    __shared__ int s_changed, s_nq, s_nAcc, s_nRemoved, s_ind[3];
    cg::cluster_group cluster = cg::this_cluster();
    const int C = (int)cluster.num_blocks(), rank = (int)cluster.block_rank();
    const int f = blockIdx.x / C, tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const int K = min(P.nK[f], P.kcap), M = min(P.nM[f], P.mcap);
    const FrameSmem S = carve_frame_smem(smem_raw, P.ks, P.ms, P.descInSmem != 0);
    const OrbKeyPoint* kps = P.kps + (size_t)f * P.kcap;
    int* match = P.match + (size_t)f * P.kcap;
    uint8_t* claimed = P.claimed + (size_t)f * P.kcap;
    const bool reset = P.resetState != 0;

    // ---- the frame in shared memory: grid, positions, octaves, descriptors; claims at entry ----
    for (int w = tid; w < (P.ks + 31) / 32; w += MF_NT) S.initBits[w] = 0;
    build_frame_smem(S, kps, P.desc + (size_t)f * P.kcap * 32, K, P.minX, P.minY, P.gridWInv, P.gridHInv, s_warp);
    if (!reset) {
        for (int i = tid; i < K; i += MF_NT)
            if (match[i] >= 0 && claimed[i]) atomicOr(&S.initBits[i >> 5], 1u << (i & 31));
        __syncthreads();
    }

    // ---- pass 1: this CTA's share of the map points, warp per map point ----
    auto notInit = [&](int idx) { return !((S.initBits[idx >> 5] >> (idx & 31)) & 1u); };
    for (int i = rank * (MF_NT / 32) + wid; i < M; i += C * (MF_NT / 32)) {
        const size_t o = (size_t)f * P.mcap + i;
        float u = 0, v = 0, r = 0; int minL = 0, maxL = 0;
        const bool ok = make_query(P, f, i, u, v, r, minL, maxL);
        int4 ri = make_int4(-1, -1, -1, -1), rd = make_int4(256, 256, 256, 256);
        if (ok) {
            uint32_t mpd[8];
            load_mp_desc(P, f, i, mpd);
            const TopK<4> t = scan_candidates<4>(P, S, f, u, v, r, minL, maxL, mpd, notInit);
            ri = make_int4(t.i[0], t.i[1], t.i[2], t.i[3]);
            rd = make_int4(t.i[0] >= 0 ? (int)(t.k[0] >> 40) : 256, t.i[1] >= 0 ? (int)(t.k[1] >> 40) : 256,
                           t.i[2] >= 0 ? (int)(t.k[2] >> 40) : 256, t.i[3] >= 0 ? (int)(t.k[3] >> 40) : 256);
        }
        if (lane == 0) {
            P.resultIdx[o] = ri; P.resultDist[o] = rd;
            P.query[o] = make_float4(u, v, r, __int_as_float((minL + 1) | ((maxL + 1) << 8) | ((ok ? 1 : 0) << 16)));
        }
    }
    if (C > 1) { __threadfence(); cluster.sync(); }     // the lists of all CTAs are visible to CTA 0
    else __syncthreads();
    if (rank != 0) return;

    // ---- pass 2 (CTA 0): fixed-point resolution of the sequential claim rule ----
    const int4* RI = P.resultIdx + (size_t)f * P.mcap;
    const int4* RD = P.resultDist + (size_t)f * P.mcap;
    for (int i = tid; i < M; i += MF_NT) { S.choice[i] = (uint16_t)NONE16; S.mflags[i] = P.hasObs[(size_t)f * P.mcap + i] ? 1 : 0; }
    for (int c = tid; c < K; c += MF_NT) S.minClaimer[c] = 0x7fffffff;
    if (tid == 0) { s_changed = 0; s_nq = 0; }
    __syncthreads();
    // decision of map point i from (best, second) among the usable candidates
    auto decide = [&](int bIdx, int bDist, int sIdx, int sDist) -> unsigned {
        if (bIdx < 0 || bDist > TH_HIGH) return NONE16;
        if (P.mode == 0) {   // ratio test only when best and second come from the same level (:123-128)
            const int lv1 = S.koct[bIdx], lv2 = sIdx >= 0 ? (int)S.koct[sIdx] : -1;
            if (lv1 == lv2 && (float)bDist > fmul(P.nnratio, (float)sDist)) return NONE16;
        }
        return (unsigned)bIdx;
    };
    const bool need2 = P.mode == 0;
    for (int sweep = 0; sweep <= M; ++sweep) {
        // A. thread per map point: walk the four candidates against the claims of lower-index map points
        for (int i = tid; i < M; i += MF_NT) {
            const int4 ri = RI[i];
            unsigned ch = NONE16;
            if (ri.x >= 0) {
                const int4 rd = RD[i];
                const int ids[4] = {ri.x, ri.y, ri.z, ri.w};
                const int ds[4] = {rd.x, rd.y, rd.z, rd.w};
                int bIdx = -1, bDist = 256, sIdx = -1, sDist = 256;
                bool exhausted = false;
#pragma unroll
                for (int c = 0; c < 4; ++c) {
                    if (ids[c] < 0) { exhausted = true; continue; }
                    if (S.minClaimer[ids[c]] < i) continue;
                    if (bIdx < 0) { bIdx = ids[c]; bDist = ds[c]; }
                    else if (sIdx < 0) { sIdx = ids[c]; sDist = ds[c]; }
                }
                if (!exhausted && (bIdx < 0 || (need2 && sIdx < 0))) {
                    S.queue[atomicAdd(&s_nq, 1)] = (uint16_t)i;            // decided in phase B
                    continue;
                }
                ch = decide(bIdx, bDist, sIdx, sDist);
            }
            if (ch != S.choice[i]) { S.choice[i] = (uint16_t)ch; s_changed = 1; }
        }
        __syncthreads();
        // B. warp per queued map point: exact rescan against the current claims
        const int nq = s_nq;
        for (int q = wid; q < nq; q += MF_NT / 32) {
            const int i = S.queue[q];
            const float4 qq = P.query[(size_t)f * P.mcap + i];
            const int bits = __float_as_int(qq.w);
            uint32_t mpd[8];
            load_mp_desc(P, f, i, mpd);
            const Top2 t = scan_candidates<2>(P, S, f, qq.x, qq.y, qq.z, (bits & 0xff) - 1, ((bits >> 8) & 0xff) - 1, mpd,
                                              [&](int idx) { return notInit(idx) && !(S.minClaimer[idx] < i); });
            const unsigned ch = decide(t.i[0], t.i[0] >= 0 ? (int)(t.k[0] >> 40) : 256, t.i[1], t.i[1] >= 0 ? (int)(t.k[1] >> 40) : 256);
            if (lane == 0 && ch != S.choice[i]) { S.choice[i] = (uint16_t)ch; s_changed = 1; }
        }
        __syncthreads();
        const int changed = s_changed;
        __syncthreads();
        if (!changed) break;
        // C. claims of this sweep's choices
        for (int c = tid; c < K; c += MF_NT) S.minClaimer[c] = 0x7fffffff;
        if (tid == 0) { s_changed = 0; s_nq = 0; }
        __syncthreads();
        for (int i = tid; i < M; i += MF_NT) {
            const unsigned ch = S.choice[i];
            if (ch != NONE16 && (S.mflags[i] & 1)) atomicMin(&S.minClaimer[ch], i);
        }
        __syncthreads();
    }

    // ---- outcome: nmatches++ per accepted map point (:1762-1776), rotation histogram, ComputeThreeMaxima (:2012-2053) ----
    const bool hist = P.mode == 1 && P.checkOri;
    int* s_hist = S.cnt;                      // [HISTO_LENGTH]
    int* lastWriter = S.minClaimer;           // [K] highest-index accepted map point that chose the keypoint
    uint32_t* removedBits = S.initBits;       // [K bits] keypoints cleared by the rotation filter
    if (tid < HISTO_LENGTH) s_hist[tid] = 0;
    if (tid == 0) { s_nAcc = 0; s_nRemoved = 0; }
    for (int c = tid; c < K; c += MF_NT) lastWriter[c] = -1;
    for (int w = tid; w < (K + 31) / 32; w += MF_NT) removedBits[w] = 0;
    __syncthreads();
    int nAcc = 0;
    for (int i = tid; i < M; i += MF_NT) {
        const unsigned ch = S.choice[i];
        if (ch == NONE16) continue;
        ++nAcc;
        atomicMax(&lastWriter[ch], i);
        if (hist) {
            const int b = rot_bin(P.angle[(size_t)f * P.mcap + i], kps[ch].angle);
            S.mflags[i] = (uint8_t)((S.mflags[i] & 1) | ((b + 1) << 2));
            atomicAdd(&s_hist[b], 1);
        }
    }
    if (nAcc) atomicAdd(&s_nAcc, nAcc);
    __syncthreads();
    if (hist) {
        if (tid == 0) {
            int ind1 = -1, ind2 = -1, ind3 = -1, max1 = 0, max2 = 0, max3 = 0;
            for (int b = 0; b < HISTO_LENGTH; ++b) {
                const int s = s_hist[b];
                if (s > max1) { max3 = max2; max2 = max1; max1 = s; ind3 = ind2; ind2 = ind1; ind1 = b; }
                else if (s > max2) { max3 = max2; max2 = s; ind3 = ind2; ind2 = b; }
                else if (s > max3) { max3 = s; ind3 = b; }
            }
            if ((float)max2 < fmul(0.1f, (float)max1)) { ind2 = -1; ind3 = -1; }
            else if ((float)max3 < fmul(0.1f, (float)max1)) { ind3 = -1; }
            s_ind[0] = ind1; s_ind[1] = ind2; s_ind[2] = ind3;
        }
        __syncthreads();
        int removed = 0;
        for (int i = tid; i < M; i += MF_NT) {      // every event of a rejected bin clears its keypoint (:1868-1884)
            const unsigned ch = S.choice[i];
            if (ch == NONE16) continue;
            const int b = (S.mflags[i] >> 2) - 1;
            if (b != s_ind[0] && b != s_ind[1] && b != s_ind[2]) { atomicOr(&removedBits[ch >> 5], 1u << (ch & 31)); ++removed; }
        }
        if (removed) atomicAdd(&s_nRemoved, removed);
        __syncthreads();
    }
    for (int c = tid; c < K; c += MF_NT) {
        const int w = lastWriter[c];
        if ((removedBits[c >> 5] >> (c & 31)) & 1u) { match[c] = -1; claimed[c] = 0; }
        else if (w >= 0) { match[c] = w; claimed[c] = S.mflags[w] & 1; }
        else if (reset) { match[c] = -1; claimed[c] = 0; }
    }
    if (reset) for (int c = K + tid; c < P.kcap; c += MF_NT) { match[c] = -1; claimed[c] = 0; }
    if (tid == 0) P.nmatches[f] = s_nAcc - s_nRemoved;
}

// ---------------------------------------------------------------------------------------------
// ORBmatcher::SearchForInitialization (src/ORBmatcher.cc:648-763): windowed brute force between the two frames of the monocular
// initialiser.  Same cluster layout as match_frame_kernel with F2 as the frame in shared memory:
//   pass 1 (all CTAs, warp per level-0 keypoint of F1): the four smallest (distance, enumeration order) keys among the level-0
//           keypoints of F2 inside the window around vbPrevMatched[i1];
//   pass 2 (one warp of CTA 0, F1 keypoints in index order -- the vMatchedDistance / vnMatches21 overwrite rule is order dependent):
//           best and second-best among the candidates i2 whose current match is worse (vMatchedDistance[i2] > dist, :686), TH_LOW,
//           ratio test, re-assignment of an already matched F2 keypoint (:706-714); a keypoint whose four candidates do not
//           settle it is rescanned against the live vMatchedDistance;
//   then the rotation histogram on F1 indices (:716-726, :733-754) and the vbPrevMatched update (:757-759).
// MatchParams reuse: kps/desc/nK = F2, mpDesc/nM = descriptors / count of F1, projX/projY = vbPrevMatched (in), level = F1 octaves,
// angle = F1 angles, th = windowSize, nnratio; outputs: match = vnMatches12 [K1], query = updated vbPrevMatched (xy), nmatches.
// ---------------------------------------------------------------------------------------------
constexpr int TH_LOW = 50;         // src/ORBmatcher.cc:36
__global__ void __launch_bounds__(MF_NT) init_match_kernel(MatchParams P) {
    extern __shared__ __align__(16) uint8_t smem_raw[];
    __shared__ int s_warp[33];
    __shared__ int s_hist[HISTO_LENGTH], s_ind[3], s_nm;
    cg::cluster_group cluster = cg::this_cluster();
    const int C = (int)cluster.num_blocks(), rank = (int)cluster.block_rank();
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const int K2 = min(P.nK[0], P.kcap), K1 = min(P.nM[0], P.mcap);
    const FrameSmem S = carve_frame_smem(smem_raw, P.ks, P.ms, P.descInSmem != 0);
    uint16_t* m21 = reinterpret_cast<uint16_t*>(smem_raw + frame_smem_bytes(P.ks, P.ms, P.descInSmem != 0));   // vnMatches21 [K2]
    int* matchedDist = S.minClaimer;                                                                             // vMatchedDistance [K2]
    build_frame_smem(S, P.kps, P.desc, K2, P.minX, P.minY, P.gridWInv, P.gridHInv, s_warp);
    auto anyIdx = [](int) { return true; };
    // ---- pass 1 ----
    for (int i = rank * (MF_NT / 32) + wid; i < K1; i += C * (MF_NT / 32)) {
        int4 ri = make_int4(-1, -1, -1, -1), rd = make_int4(256, 256, 256, 256);
        const int level1 = P.level[i];
        if (!(level1 > 0)) {                                        // :663-665
            uint32_t mpd[8];
            load_mp_desc(P, 0, i, mpd);
            const TopK<4> t = scan_candidates<4>(P, S, 0, P.projX[i], P.projY[i], P.th, level1, level1, mpd, anyIdx);
            ri = make_int4(t.i[0], t.i[1], t.i[2], t.i[3]);
            rd = make_int4(t.i[0] >= 0 ? (int)(t.k[0] >> 40) : 256, t.i[1] >= 0 ? (int)(t.k[1] >> 40) : 256,
                           t.i[2] >= 0 ? (int)(t.k[2] >> 40) : 256, t.i[3] >= 0 ? (int)(t.k[3] >> 40) : 256);
        }
        if (lane == 0) { P.resultIdx[i] = ri; P.resultDist[i] = rd; }
    }
    if (C > 1) { __threadfence(); cluster.sync(); }
    else __syncthreads();
    if (rank != 0) return;
    // ---- pass 2 ----
    for (int c = tid; c < K2; c += MF_NT) { matchedDist[c] = 0x7fffffff; m21[c] = (uint16_t)NONE16; }
    for (int i = tid; i < K1; i += MF_NT) { S.choice[i] = (uint16_t)NONE16; S.mflags[i] = 0; }
    if (tid < HISTO_LENGTH) s_hist[tid] = 0;
    if (tid == 0) s_nm = 0;
    __syncthreads();
    if (wid == 0) {
        for (int base = 0; base < K1; base += 32) {
            const int i = base + lane;
            int4 ri = make_int4(-1, -1, -1, -1), rd = make_int4(256, 256, 256, 256);
            if (i < K1) { ri = P.resultIdx[i]; rd = P.resultDist[i]; }
            const int cnt = min(32, K1 - base);
            for (int j = 0; j < cnt; ++j) {
                const int c0 = __shfl_sync(0xffffffffu, ri.x, j);
                if (c0 < 0) continue;                                 // level > 0 or no candidate (:671-672)
                const int ids[4] = {c0, __shfl_sync(0xffffffffu, ri.y, j), __shfl_sync(0xffffffffu, ri.z, j), __shfl_sync(0xffffffffu, ri.w, j)};
                const int ds[4] = {__shfl_sync(0xffffffffu, rd.x, j), __shfl_sync(0xffffffffu, rd.y, j), __shfl_sync(0xffffffffu, rd.z, j),
                                   __shfl_sync(0xffffffffu, rd.w, j)};
                const int i1 = base + j;
                int bIdx = -1, bDist = 0x7fffffff, sDist = 0x7fffffff, found = 0;
                bool exhausted = false;
#pragma unroll
                for (int c = 0; c < 4; ++c) {
                    if (ids[c] < 0) { exhausted = true; continue; }
                    if (matchedDist[ids[c]] <= ds[c]) continue;       // :686-687
                    if (found == 0) { bIdx = ids[c]; bDist = ds[c]; }
                    else if (found == 1) sDist = ds[c];
                    ++found;
                }
                if (!exhausted && found < 2) {                        // the four do not settle best and second: exact rescan
                    uint32_t mpd[8];
                    load_mp_desc(P, 0, i1, mpd);
                    const int level1 = P.level[i1];
                    const Top2 t = scan_candidates<2>(P, S, 0, P.projX[i1], P.projY[i1], P.th, level1, level1, mpd, anyIdx,
                                                      [&](int idx, int dist) { return !(matchedDist[idx] <= dist); });
                    bIdx = t.i[0]; bDist = t.i[0] >= 0 ? (int)(t.k[0] >> 40) : 0x7fffffff;
                    sDist = t.i[1] >= 0 ? (int)(t.k[1] >> 40) : 0x7fffffff;
                }
                if (bIdx < 0 || bDist > TH_LOW) continue;             // :702
                if (!((float)bDist < fmul((float)sDist, P.nnratio))) continue;    // :704
                __syncwarp();                                         // every lane has read matchedDist[] for this keypoint before lane 0 updates it
                if (lane == 0) {
                    const unsigned prev = m21[bIdx];
                    if (prev != NONE16) S.choice[prev] = (uint16_t)NONE16;        // :706-710
                    S.choice[i1] = (uint16_t)bIdx; m21[bIdx] = (uint16_t)i1; matchedDist[bIdx] = bDist;
                    if (P.checkOri) {                                             // :716-726
                        const int b = rot_bin(P.angle[i1], P.kps[bIdx].angle);
                        S.mflags[i1] = (uint8_t)(b + 1);
                        s_hist[b]++;
                    }
                }
                __syncwarp();
            }
        }
    }
    __syncthreads();
    if (P.checkOri && tid == 0) {   // ComputeThreeMaxima (:2012-2053)
        int ind1 = -1, ind2 = -1, ind3 = -1, max1 = 0, max2 = 0, max3 = 0;
        for (int b = 0; b < HISTO_LENGTH; ++b) {
            const int sz = s_hist[b];
            if (sz > max1) { max3 = max2; max2 = max1; max1 = sz; ind3 = ind2; ind2 = ind1; ind1 = b; }
            else if (sz > max2) { max3 = max2; max2 = sz; ind3 = ind2; ind2 = b; }
            else if (sz > max3) { max3 = sz; ind3 = b; }
        }
        if ((float)max2 < fmul(0.1f, (float)max1)) { ind2 = -1; ind3 = -1; }
        else if ((float)max3 < fmul(0.1f, (float)max1)) { ind3 = -1; }
        s_ind[0] = ind1; s_ind[1] = ind2; s_ind[2] = ind3;
    }
    __syncthreads();
    int n = 0;
    for (int i = tid; i < K1; i += MF_NT) {
        unsigned ch = S.choice[i];
        if (ch != NONE16 && P.checkOri) {                             // :733-754: events of rejected bins lose their match if they still hold one
            const int b = (int)S.mflags[i] - 1;
            if (b != s_ind[0] && b != s_ind[1] && b != s_ind[2]) ch = NONE16;
        }
        P.match[i] = ch == NONE16 ? -1 : (int)ch;
        float px = P.projX[i], py = P.projY[i];
        if (ch != NONE16) { ++n; px = S.kx[ch]; py = S.ky[ch]; }      // :757-759
        P.query[i] = make_float4(px, py, 0.f, 0.f);
    }
    if (n) atomicAdd(&s_nm, n);
    __syncthreads();
    if (tid == 0) P.nmatches[0] = s_nm;
}

// ---------------------------------------------------------------------------------------------
// ORBmatcher::SearchByBoW(KeyFrame*, Frame&, vector<MapPoint*>&) (src/ORBmatcher.cc:223-425, monocular branch; Tracking::TrackReferenceKeyFrame
// src/Tracking.cc:2730 and relocalisation).  Features are matched only inside the same vocabulary node, so nodes are independent: one warp per
// keyframe node walks its features in order (the claim rule "a frame feature that already holds a match is skipped" is sequential inside a
// node only), lanes over the frame features of that node: top-2 by (distance, position).  The rotation histogram is resolved by the CTA.
// One CTA per (keyframe, frame) pair.
// ---------------------------------------------------------------------------------------------
struct BowMatchParams {
    int nKF, nF, eKF, eF;
    const OrbKeyPoint *kpsKF, *kpsF; const uint8_t *descKF, *descF, *kfPoint;
    const int *fvNodeKF, *fvFeatKF, *fvNodeF, *fvFeatF;
    float nnratio; int checkOri;
    int* match; int* nmatches; int* groupStart; uint8_t* evBin;   // match [nF]; scratch: groupStart [eKF + 1], evBin [nF]
    // SearchByBoW(KeyFrame*, KeyFrame*) (src/ORBmatcher.cc:765-905): the second side is a keyframe too -- its features need a good map point
    // (fPoint, null for a Frame), the distance test is strict (bestDist1 < TH_LOW, :859) and the result is indexed by the first side (match12)
    const uint8_t* fPoint; int strictLow; int* match12;
};
constexpr int BM_NT = 256;
__global__ void __launch_bounds__(BM_NT) bow_match_kernel(BowMatchParams Q) {
    __shared__ int s_hist[HISTO_LENGTH], s_ind[3], s_acc, s_rem, s_ng;
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    for (int i = tid; i < Q.nF; i += BM_NT) { Q.match[i] = -1; Q.evBin[i] = 0; }
    if (Q.match12) for (int i = tid; i < Q.nKF; i += BM_NT) Q.match12[i] = -1;
    if (tid < HISTO_LENGTH) s_hist[tid] = 0;
    if (tid == 0) { s_acc = 0; s_rem = 0; s_ng = 0; }
    __syncthreads();
    // group heads of the keyframe's feature vector (entries are sorted by node id)
    for (int e = tid; e < Q.eKF; e += BM_NT)
        if (e == 0 || Q.fvNodeKF[e] != Q.fvNodeKF[e - 1]) Q.groupStart[atomicAdd(&s_ng, 1)] = e;     // order of groups is irrelevant: nodes are independent
    __syncthreads();
    const int ng = s_ng;
    int acc = 0;
    for (int g = wid; g < ng; g += BM_NT / 32) {
        const int a = Q.groupStart[g];
        const int node = Q.fvNodeKF[a];
        int a1 = a;
        while (a1 < Q.eKF && Q.fvNodeKF[a1] == node) ++a1;
        int lo = 0, hi = Q.eF;                                   // lower_bound / upper_bound of `node` in the frame's vector
        while (lo < hi) { const int mid = (lo + hi) >> 1; if (Q.fvNodeF[mid] < node) lo = mid + 1; else hi = mid; }
        const int b = lo;
        hi = Q.eF;
        while (lo < hi) { const int mid = (lo + hi) >> 1; if (Q.fvNodeF[mid] <= node) lo = mid + 1; else hi = mid; }
        const int b1 = lo;
        if (b1 == b) continue;
        for (int iKF = a; iKF < a1; ++iKF) {
            const int realIdxKF = Q.fvFeatKF[iKF];
            if (Q.kfPoint[realIdxKF] != 1) continue;             // no map point, or a bad one (:255-259)
            const uint4* pk = reinterpret_cast<const uint4*>(Q.descKF + (size_t)realIdxKF * 32);
            const uint4 k0 = __ldg(pk), k1 = __ldg(pk + 1);
            const uint32_t dk[8] = {k0.x, k0.y, k0.z, k0.w, k1.x, k1.y, k1.z, k1.w};
            Top2 t; t.init();
            for (int iF = b + lane; iF < b1; iF += 32) {
                const int realIdxF = Q.fvFeatF[iF];
                if (Q.match[realIdxF] >= 0) continue;            // already matched by an earlier keyframe feature of this node (:275-276)
                if (Q.fPoint && Q.fPoint[realIdxF] != 1) continue;
                const uint4* pf = reinterpret_cast<const uint4*>(Q.descF + (size_t)realIdxF * 32);
                const uint4 f0 = __ldg(pf), f1 = __ldg(pf + 1);
                const uint32_t dfw[8] = {f0.x, f0.y, f0.z, f0.w, f1.x, f1.y, f1.z, f1.w};
                t.insert(((unsigned long long)hamming256(dk, dfw) << 32) | (unsigned)(iF - b), realIdxF);
            }
            t.warp_merge();
            const int bestDist1 = t.i[0] >= 0 ? (int)(t.k[0] >> 32) : 256, bestDist2 = t.i[1] >= 0 ? (int)(t.k[1] >> 32) : 256;
            if ((Q.strictLow ? bestDist1 < TH_LOW : bestDist1 <= TH_LOW) && (float)bestDist1 < fmul(Q.nnratio, (float)bestDist2)) {
                if (lane == 0) {
                    Q.match[t.i[0]] = realIdxKF;
                    if (Q.match12) Q.match12[realIdxKF] = t.i[0];
                    if (Q.checkOri) {
                        const int bin = rot_bin(Q.kpsKF[realIdxKF].angle, Q.kpsF[t.i[0]].angle);
                        Q.evBin[t.i[0]] = (uint8_t)(bin + 1);
                        atomicAdd(&s_hist[bin], 1);
                    }
                }
                ++acc;
            }
            __syncwarp();
        }
    }
    if (lane == 0 && acc) atomicAdd(&s_acc, acc);
    __syncthreads();
    if (Q.checkOri) {
        if (tid == 0) {
            int ind1 = -1, ind2 = -1, ind3 = -1, max1 = 0, max2 = 0, max3 = 0;
            for (int bb = 0; bb < HISTO_LENGTH; ++bb) {
                const int sz = s_hist[bb];
                if (sz > max1) { max3 = max2; max2 = max1; max1 = sz; ind3 = ind2; ind2 = ind1; ind1 = bb; }
                else if (sz > max2) { max3 = max2; max2 = sz; ind3 = ind2; ind2 = bb; }
                else if (sz > max3) { max3 = sz; ind3 = bb; }
            }
            if ((float)max2 < fmul(0.1f, (float)max1)) { ind2 = -1; ind3 = -1; }
            else if ((float)max3 < fmul(0.1f, (float)max1)) { ind3 = -1; }
            s_ind[0] = ind1; s_ind[1] = ind2; s_ind[2] = ind3;
        }
        __syncthreads();
        int rem = 0;
        for (int i = tid; i < Q.nF; i += BM_NT) {
            const int bb = (int)Q.evBin[i] - 1;
            if (bb >= 0 && bb != s_ind[0] && bb != s_ind[1] && bb != s_ind[2]) { if (Q.match12) Q.match12[Q.match[i]] = -1; Q.match[i] = -1; ++rem; }
        }
        if (rem) atomicAdd(&s_rem, rem);
        __syncthreads();
    }
    if (tid == 0) *Q.nmatches = s_acc - s_rem;
}

// ---------------------------------------------------------------------------------------------
// ORBmatcher::Fuse(KeyFrame*, vector<MapPoint*>, th) (src/ORBmatcher.cc:1148-1338), the search: the keyframe's grid in shared memory (every CTA
// builds its own copy), one warp per map point: projection with the keyframe pose, KeyFrame::IsInImage, distance and viewing-angle tests,
// MapPoint::PredictScale, then the radius search of KeyFrame::GetFeaturesInArea (src/KeyFrame.cc:704-753) with the level window
// [predicted - 1, predicted] and the chi-square gate e2 * invSigma2 > 5.99; first smallest distance wins.  No search depends on what Fuse does
// with another map point's hit, so the AddObservation / Replace bookkeeping stays with the caller (in map-point order).
// ---------------------------------------------------------------------------------------------
struct FuseParams {
    MatchParams P;          // keyframe: kps, desc, bounds, scaleFactors, nlevels, ks, descInSmem
    int K, M;
    const uint8_t *state, *mpDesc;
    const float *xyz, *normal, *minD, *maxD, *invSigma2;
    float Tcw[7], Ow[3], logSF;
    int mode, gate;         // mode 0: world point + pose + normal test (Fuse); 2: point given in the camera frame (SearchBySim3).  gate: chi-square test
    int *bestIdx, *bestDist;
};
__global__ void __launch_bounds__(MF_NT) fuse_search_kernel(FuseParams Q) {
    extern __shared__ __align__(16) uint8_t smem_raw[];
    __shared__ int s_warp[33];
    const MatchParams& P = Q.P;
    const int tid = threadIdx.x, lane = tid & 31;
    const FrameSmem S = carve_frame_smem(smem_raw, P.ks, 1, P.descInSmem != 0);
    build_frame_smem(S, P.kps, P.desc, Q.K, P.minX, P.minY, P.gridWInv, P.gridHInv, s_warp);
    const int nw = gridDim.x * (MF_NT / 32);
    for (int i = blockIdx.x * (MF_NT / 32) + (tid >> 5); i < Q.M; i += nw) {
        int bestIdx = -1, bestDist = 256;
        bool go = Q.state[i] == 1;
        float u = 0.f, v = 0.f, r = 0.f; int lvl = 0;
        if (go) {
            const float px = Q.xyz[3 * i], py = Q.xyz[3 * i + 1], pz = Q.xyz[3 * i + 2];
            const float maxRaw = Q.maxD[i];
            const float maxDist = fmul(1.2f, maxRaw), minDist = fmul(0.8f, Q.minD[i]);
            float dist3D;
            if (Q.mode == 0) {
                const float qw = Q.Tcw[0], qx = Q.Tcw[1], qy = Q.Tcw[2], qz = Q.Tcw[3];
                float ux = fsub(fmul(qy, pz), fmul(qz, py)), uy = fsub(fmul(qz, px), fmul(qx, pz)), uz = fsub(fmul(qx, py), fmul(qy, px));
                ux = fadd(ux, ux); uy = fadd(uy, uy); uz = fadd(uz, uz);
                const float cx_ = fsub(fmul(qy, uz), fmul(qz, uy)), cy_ = fsub(fmul(qz, ux), fmul(qx, uz)), cz_ = fsub(fmul(qx, uy), fmul(qy, ux));
                const float xc = fadd(fadd(fadd(px, fmul(qw, ux)), cx_), Q.Tcw[4]);
                const float yc = fadd(fadd(fadd(py, fmul(qw, uy)), cy_), Q.Tcw[5]);
                const float zc = fadd(fadd(fadd(pz, fmul(qw, uz)), cz_), Q.Tcw[6]);
                go = !(zc < 0.0f);
                u = fadd(fdiv(fmul(P.cam[0], xc), zc), P.cam[2]);
                v = fadd(fdiv(fmul(P.cam[1], yc), zc), P.cam[3]);
                go = go && (u >= P.minX && u < P.maxX && v >= P.minY && v < P.maxY);                                 // KeyFrame::IsInImage
                const float ox = fsub(px, Q.Ow[0]), oy = fsub(py, Q.Ow[1]), oz = fsub(pz, Q.Ow[2]);
                dist3D = __fsqrt_rn(fadd(fadd(fmul(ox, ox), fmul(oy, oy)), fmul(oz, oz)));
                go = go && !(dist3D < minDist || dist3D > maxDist);
                const float dot = fadd(fadd(fmul(ox, Q.normal[3 * i]), fmul(oy, Q.normal[3 * i + 1])), fmul(oz, Q.normal[3 * i + 2]));
                go = go && !((double)dot < 0.5 * (double)dist3D);
            } else {                                                   // src/ORBmatcher.cc:1510-1532 (and :1589-1611)
                go = !(pz < 0.0f);
                const float invz = (float)(1.0 / (double)pz);
                u = __fmaf_rn(P.cam[0], fmul(px, invz), P.cam[2]);       // fx*x+cx: one FMA in the reference build, see oracle
                v = __fmaf_rn(P.cam[1], fmul(py, invz), P.cam[3]);
                go = go && (u >= P.minX && u < P.maxX && v >= P.minY && v < P.maxY);
                dist3D = __fsqrt_rn(fadd(fadd(fmul(px, px), fmul(py, py)), fmul(pz, pz)));
                go = go && !(dist3D < minDist || dist3D > maxDist);
            }
            if (go) {
                lvl = (int)ceilf(fdiv(orbx::logf_glibc(fdiv(maxRaw, dist3D)), Q.logSF));
                if (lvl < 0) lvl = 0; else if (lvl >= P.nlevels) lvl = P.nlevels - 1;
                r = fmul(P.th, P.scaleFactors[lvl]);
            }
        }
        if (go) {      // warp-uniform: every lane computed the same query
            uint32_t mpd[8];
            const uint4* dp = reinterpret_cast<const uint4*>(Q.mpDesc + (size_t)i * 32);
            const uint4 a = __ldg(dp), b = __ldg(dp + 1);
            mpd[0] = a.x; mpd[1] = a.y; mpd[2] = a.z; mpd[3] = a.w; mpd[4] = b.x; mpd[5] = b.y; mpd[6] = b.z; mpd[7] = b.w;
            const float* isg = Q.invSigma2;
            const bool gate = Q.gate != 0;
            const TopK<1> t = scan_candidates<1>(P, S, 0, u, v, r, lvl - 1, lvl, mpd, [&](int idx) {
                if (!gate) return true;
                const float ex = fsub(u, S.kx[idx]), ey = fsub(v, S.ky[idx]);
                const float e2 = __fmaf_rn(ex, ex, fmul(ey, ey));                     // one FMA in the reference build, see oracle
                return !((double)fmul(e2, isg[S.koct[idx]]) > 5.99);
            });
            if (t.i[0] >= 0) { bestIdx = t.i[0]; bestDist = (int)(t.k[0] >> 40); }
        }
        if (lane == 0) { Q.bestIdx[i] = bestIdx; Q.bestDist[i] = bestDist; }
    }
}

// ---------------------------------------------------------------------------------------------
// ORBmatcher::SearchForTriangulation(pKF1, pKF2, vMatchedPairs, bOnlyStereo = false, bCoarse) (src/ORBmatcher.cc:907-1146, monocular keyframes;
// LocalMapping::CreateNewMapPoints, src/LocalMapping.cc:340-344) with Pinhole::epipolarConstrain (src/CameraModels/Pinhole.cpp:107-129).
// One CTA per (KF1, KF2) pair, one warp per KF1 feature without a map point: lanes over the KF2 features of the same vocabulary node (binary
// search in KF2's feature vector); a candidate passes with distance <= TH_LOW, no map point, the epipole-distance test and the epipolar test;
// the smallest distance wins, the last candidate on ties (:1008 is strict).  The reference never sets vbMatched2, so KF1 features are
// independent.  The epipole and the fundamental matrix are inputs (Eigen / Sophus expressions in the caller); the scalar float expressions
// are contracted as the reference's build contracts them (oracle/matcher_oracle.cpp: orbo_search_for_triangulation).
// ---------------------------------------------------------------------------------------------
struct TriPair {
    int N2, E2;
    const OrbKeyPoint* kps2; const uint8_t *desc2, *hasMP2; const int *fvNode2, *fvFeat2;
    float ep[2], F12[9];
    int* matches12; int* nmatches; uint8_t* evBin;
};
struct TriParams {
    int N1, E1, coarse, checkOri;
    const OrbKeyPoint* kps1; const uint8_t *desc1, *hasMP1; const int *fvNode1, *fvFeat1;
    const float *scaleFactors, *levelSigma2;
    const TriPair* pairs;
};
constexpr int TM_NT = 256;
__global__ void __launch_bounds__(TM_NT) tri_match_kernel(TriParams Q) {
    __shared__ int s_hist[HISTO_LENGTH], s_ind[3], s_acc, s_rem;
    const TriPair R = Q.pairs[blockIdx.x];
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    for (int i = tid; i < Q.N1; i += TM_NT) { R.matches12[i] = -1; R.evBin[i] = 0; }
    if (tid < HISTO_LENGTH) s_hist[tid] = 0;
    if (tid == 0) { s_acc = 0; s_rem = 0; }
    __syncthreads();
    int acc = 0;
    for (int e1 = wid; e1 < Q.E1; e1 += TM_NT / 32) {
        const int idx1 = Q.fvFeat1[e1];
        if (Q.hasMP1[idx1]) continue;
        const int node = Q.fvNode1[e1];
        int lo = 0, hi = R.E2;
        while (lo < hi) { const int mid = (lo + hi) >> 1; if (R.fvNode2[mid] < node) lo = mid + 1; else hi = mid; }
        const int b = lo;
        hi = R.E2;
        while (lo < hi) { const int mid = (lo + hi) >> 1; if (R.fvNode2[mid] <= node) lo = mid + 1; else hi = mid; }
        const int b1 = lo;
        if (b1 == b) continue;
        const OrbKeyPoint kp1 = Q.kps1[idx1];
        const uint4* p1 = reinterpret_cast<const uint4*>(Q.desc1 + (size_t)idx1 * 32);
        const uint4 k0 = __ldg(p1), k1 = __ldg(p1 + 1);
        const uint32_t d1[8] = {k0.x, k0.y, k0.z, k0.w, k1.x, k1.y, k1.z, k1.w};
        // epipolar line of kp1 in the second image, l = x1' F12 = [a b c]
        const float ea = fadd(__fmaf_rn(kp1.x, R.F12[0], fmul(kp1.y, R.F12[3])), R.F12[6]);
        const float eb = fadd(__fmaf_rn(kp1.x, R.F12[1], fmul(kp1.y, R.F12[4])), R.F12[7]);
        const float ec = fadd(__fmaf_rn(kp1.x, R.F12[2], fmul(kp1.y, R.F12[5])), R.F12[8]);
        const float den = __fmaf_rn(ea, ea, fmul(eb, eb));
        unsigned long long best = ~0ull;
        for (int i2 = b + lane; i2 < b1; i2 += 32) {
            const int idx2 = R.fvFeat2[i2];
            if (R.hasMP2[idx2]) continue;
            const uint4* p2 = reinterpret_cast<const uint4*>(R.desc2 + (size_t)idx2 * 32);
            const uint4 f0 = __ldg(p2), f1 = __ldg(p2 + 1);
            const uint32_t d2[8] = {f0.x, f0.y, f0.z, f0.w, f1.x, f1.y, f1.z, f1.w};
            const int dist = hamming256(d1, d2);
            if (dist > TH_LOW) continue;
            const OrbKeyPoint kp2 = R.kps2[idx2];
            const int oct = min(max(kp2.octave, 0), 255);
            const float distex = fsub(R.ep[0], kp2.x), distey = fsub(R.ep[1], kp2.y);
            if (__fmaf_rn(distex, distex, fmul(distey, distey)) < fmul(100.0f, Q.scaleFactors[oct])) continue;
            if (!Q.coarse) {
                if (den == 0.0f) continue;
                const float num = fadd(__fmaf_rn(ea, kp2.x, fmul(eb, kp2.y)), ec);
                const float dsqr = fdiv(fmul(num, num), den);
                if (!((double)dsqr < 3.84 * (double)Q.levelSigma2[oct])) continue;
            }
            const unsigned long long key = ((unsigned long long)dist << 32) | (0xFFFFFFFFu - (unsigned)(i2 - b));
            best = key < best ? key : best;
        }
#pragma unroll
        for (int o = 16; o; o >>= 1) { const unsigned long long t = __shfl_xor_sync(0xffffffffu, best, o); best = t < best ? t : best; }
        if (best != ~0ull) {
            if (lane == 0) {
                const int idx2 = R.fvFeat2[b + (int)(0xFFFFFFFFu - (unsigned)best)];
                R.matches12[idx1] = idx2;
                if (Q.checkOri) {
                    const int bin = rot_bin(kp1.angle, R.kps2[idx2].angle);
                    R.evBin[idx1] = (uint8_t)(bin + 1);
                    atomicAdd(&s_hist[bin], 1);
                }
            }
            ++acc;
        }
    }
    if (lane == 0 && acc) atomicAdd(&s_acc, acc);
    __syncthreads();
    if (Q.checkOri) {
        if (tid == 0) {
            int ind1 = -1, ind2 = -1, ind3 = -1, max1 = 0, max2 = 0, max3 = 0;
            for (int bb = 0; bb < HISTO_LENGTH; ++bb) {
                const int sz = s_hist[bb];
                if (sz > max1) { max3 = max2; max2 = max1; max1 = sz; ind3 = ind2; ind2 = ind1; ind1 = bb; }
                else if (sz > max2) { max3 = max2; max2 = sz; ind3 = ind2; ind2 = bb; }
                else if (sz > max3) { max3 = sz; ind3 = bb; }
            }
            if ((float)max2 < fmul(0.1f, (float)max1)) { ind2 = -1; ind3 = -1; }
            else if ((float)max3 < fmul(0.1f, (float)max1)) { ind3 = -1; }
            s_ind[0] = ind1; s_ind[1] = ind2; s_ind[2] = ind3;
        }
        __syncthreads();
        int rem = 0;
        for (int i = tid; i < Q.N1; i += TM_NT) {
            const int bb = (int)R.evBin[i] - 1;
            if (bb >= 0 && bb != s_ind[0] && bb != s_ind[1] && bb != s_ind[2]) { R.matches12[i] = -1; ++rem; }
        }
        if (rem) atomicAdd(&s_rem, rem);
        __syncthreads();
    }
    if (tid == 0) *R.nmatches = s_acc - s_rem;
}

// ---------------------------------------------------------------------------------------------
// MapPoint::ComputeDistinctiveDescriptors (src/MapPoint.cc:329-403) for a batch of map points: warp per map point, lane per observation row;
// the row's median distance (element (int)(0.5 (n - 1)) of the sorted row) by bisection on the distance value; first minimum wins.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) distinctive_kernel(int nPoints, const int* __restrict__ ptStart, const uint8_t* __restrict__ desc, int* __restrict__ best) {
    const int p = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (p >= nPoints) return;
    const int a = ptStart[p], n = ptStart[p + 1] - a;
    if (n <= 0) { if (lane == 0) best[p] = -1; return; }
    const int k = (int)(0.5 * (double)(n - 1));
    unsigned bestKey = 0xFFFFFFFFu;
    for (int i = lane; i < n; i += 32) {
        const uint4* pi = reinterpret_cast<const uint4*>(desc + (size_t)(a + i) * 32);
        const uint4 x0 = __ldg(pi), x1 = __ldg(pi + 1);
        const uint32_t di[8] = {x0.x, x0.y, x0.z, x0.w, x1.x, x1.y, x1.z, x1.w};
        int lo = 0, hi = 256;                                    // smallest v with #{j : d(i, j) <= v} >= k + 1
        while (lo < hi) {
            const int mid = (lo + hi) >> 1;
            int cnt = 0;
            for (int j = 0; j < n; ++j) {
                const uint4* pj = reinterpret_cast<const uint4*>(desc + (size_t)(a + j) * 32);
                const uint4 y0 = __ldg(pj), y1 = __ldg(pj + 1);
                const uint32_t dj[8] = {y0.x, y0.y, y0.z, y0.w, y1.x, y1.y, y1.z, y1.w};
                cnt += (i == j ? 0 : hamming256(di, dj)) <= mid;
            }
            if (cnt >= k + 1) hi = mid; else lo = mid + 1;
        }
        bestKey = min(bestKey, ((unsigned)lo << 16) | (unsigned)i);
    }
    bestKey = __reduce_min_sync(0xffffffffu, bestKey);
    if (lane == 0) best[p] = (int)(bestKey & 0xFFFFu);
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
    cudaEvent_t evWait = nullptr;      // blocking-sync event of the batch host entry points (device_utils.cuh)
    // scratch
    float4* d_query = nullptr; int4 *d_resultIdx = nullptr, *d_resultDist = nullptr;
    int* d_status = nullptr;
    int numSMs = 148;
    // staging for the host entry points (batch = 1) -- one arena
    uint8_t* d_arena = nullptr; size_t arenaBytes = 0;
    uint8_t* h_arena = nullptr;
    uint8_t* d_batch = nullptr; size_t batchBytes = 0;   // device staging for the host-pointer batch entry point
    int launches = 0;

    ~Matcher() {
        cudaSetDevice(device);
        void* ptrs[] = {d_query, d_resultIdx, d_resultDist, d_status, d_arena, d_batch};
        for (void* p : ptrs) if (p) cudaFree(p);
        if (h_arena) cudaFreeHost(h_arena);
        if (evWait) cudaEventDestroy(evWait);
        if (stream) cudaStreamDestroy(stream);
    }
    int init() {
        CK(cudaSetDevice(device));
        cudaDeviceProp prop;
        CK(cudaGetDeviceProperties(&prop, device));
        if (prop.major < 10) { set_error("device is not sm_100+ (Blackwell); this library has no other code path"); return ORB_ERR_CUDA; }
        const size_t B = maxBatch;
        numSMs = prop.multiProcessorCount;
        CK(cudaMalloc(&d_query, sizeof(float4) * mcap * B));
        CK(cudaMalloc(&d_resultIdx, sizeof(int4) * mcap * B));
        CK(cudaMalloc(&d_resultDist, sizeof(int4) * mcap * B));
        CK(cudaMalloc(&d_status, sizeof(int) * B));
        arenaBytes = (size_t)kcap * (28 + 32 + 4 + 1) + (size_t)mcap * (32 + 12 + 4 * 6 + 4) + 4096 + 64 * 64;
        arenaBytes = (arenaBytes + 255) & ~(size_t)255;
        const size_t bf = (size_t)(kcap + mcap) * 32 + (size_t)std::max(kcap, mcap) * 16 + 4096;
        arenaBytes = std::max(arenaBytes, bf);
        arenaBytes = std::max(arenaBytes, (size_t)(kcap + mcap) * (28 + 32 + 1 + 8 + 4 + 4 + 1) + 8192);   // SearchByBoW: keyframe + frame staged together
        CK(cudaMalloc(&d_arena, arenaBytes));
        {
            orbx::ScopedGpuAffinity numaLocal(device);     // pinned pages on the GPU's NUMA node (host_affinity.h)
            CK(cudaMallocHost(&h_arena, arenaBytes));
        }
        batchBytes = B * ((size_t)kcap * (28 + 32 + 4 + 1) + (size_t)mcap * (1 + 12 + 4 + 4 + 1 + 32) + 28 + 16 + 13 * 256) + 4096;
        CK(cudaMalloc(&d_batch, batchBytes));
        CK(cudaStreamCreateWithFlags(&stream, cudaStreamNonBlocking));
        CK(orbx::make_blocking_event(&evWait));
        return ORB_OK;
    }
    // kUse / mUse: the largest keypoint / map-point count any frame of this call can have (the slab capacities when the
    // counts live on the device): they size the kernel's shared memory.
    int run(MatchParams& P, int kUse, int mUse, cudaStream_t st) {
        P.gridWInv = (float)GRID_COLS / (P.maxX - P.minX);   // src/Frame.cc:342-343
        P.gridHInv = (float)GRID_ROWS / (P.maxY - P.minY);
        P.query = d_query; P.resultIdx = d_resultIdx; P.resultDist = d_resultDist; P.status = d_status;
        P.ks = std::max(kUse, 1); P.ms = std::max(mUse, 1);
        if (P.ks > 65535 || P.ms > 65535) { set_error("more than 65535 keypoints or map points per frame"); return ORB_ERR_CAPACITY; }
        const size_t limit = 227 * 1024 - 2048;              // opt-in maximum minus the kernel's static shared memory
        P.descInSmem = frame_smem_bytes(P.ks, P.ms, true) <= 100 * 1024 ? 1 : 0;   // two CTAs per SM when the descriptors are staged
        const size_t sm = frame_smem_bytes(P.ks, P.ms, P.descInSmem != 0);
        if (sm > limit) { set_error("frame too large for the matcher's shared-memory grid (keypoints + map points)"); return ORB_ERR_CAPACITY; }
        int rc = ensure_dynamic_smem(match_frame_kernel, sm, device);
        if (rc) return rc;
        // cluster size: as many CTAs per frame as keeps the whole call resident at once (pass 1 is spread over the cluster)
        int C = 1;
        while (C < 8 && (long)P.batch * (C * 2) <= (long)numSMs * (sm > 100 * 1024 ? 1 : 2)) C *= 2;
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3((unsigned)(P.batch * C)); cfg.blockDim = dim3(MF_NT); cfg.dynamicSmemBytes = sm; cfg.stream = st;
        cudaLaunchAttribute at[1];
        at[0].id = cudaLaunchAttributeClusterDimension; at[0].val.clusterDim.x = C; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
        cfg.attrs = at; cfg.numAttrs = 1;
        CK(cudaLaunchKernelEx(&cfg, match_frame_kernel, P));
        launches = 1;
        lastCluster = C;
        return ORB_OK;
    }
    int lastCluster = 1;

    // SearchForInitialization: F2 in shared memory (K2 keypoints), K1 keypoints of F1 as the queries
    int run_init(MatchParams& P, int K2, int K1, cudaStream_t st) {
        P.gridWInv = (float)GRID_COLS / (P.maxX - P.minX);
        P.gridHInv = (float)GRID_ROWS / (P.maxY - P.minY);
        P.query = d_query; P.resultIdx = d_resultIdx; P.resultDist = d_resultDist; P.status = d_status;
        P.ks = std::max(K2, 1); P.ms = std::max(K1, 1);
        if (P.ks > 65535 || P.ms > 65535) { set_error("more than 65535 keypoints per frame"); return ORB_ERR_CAPACITY; }
        const size_t limit = 227 * 1024 - 2048;
        auto bytes = [&](bool d) { return frame_smem_bytes(P.ks, P.ms, d) + al16(sizeof(uint16_t) * (size_t)P.ks); };
        P.descInSmem = bytes(true) <= limit ? 1 : 0;
        const size_t sm = bytes(P.descInSmem != 0);
        if (sm > limit) { set_error("frames too large for the matcher's shared-memory grid"); return ORB_ERR_CAPACITY; }
        int rc = ensure_dynamic_smem(init_match_kernel, sm, device);
        if (rc) return rc;
        cudaLaunchConfig_t cfg = {};
        const int C = 8;                                   // one pair of frames: pass 1 spread over a full cluster
        cfg.gridDim = dim3(C); cfg.blockDim = dim3(MF_NT); cfg.dynamicSmemBytes = sm; cfg.stream = st;
        cudaLaunchAttribute at[1];
        at[0].id = cudaLaunchAttributeClusterDimension; at[0].val.clusterDim.x = C; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
        cfg.attrs = at; cfg.numAttrs = 1;
        CK(cudaLaunchKernelEx(&cfg, init_match_kernel, P));
        launches = 1;
        lastCluster = C;
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

static int finish_host(Matcher& m, Arena& A, MatchParams& P, int K, int mUse, int* dMatch, uint8_t* dClaimed, int* dN, int32_t* match,
                       uint8_t* claimed, int* nmatches) {
    cudaStream_t st = m.stream;
    CK(cudaMemcpyAsync(m.d_arena, m.h_arena, A.off, cudaMemcpyHostToDevice, st));
    P.match = dMatch; P.claimed = dClaimed; P.nmatches = dN + 1;
    int rc = m.run(P, K, mUse, st);
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
    return finish_host(m, A, P, fr->K, pts->M, dMatch, dClaimed, dN, match, claimed, nmatches);
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
    return finish_host(m, A, P, fr->K, last->M, dMatch, dClaimed, dN, match, claimed, nmatches);
}

int orbm_search_last_frame_batch_device(orbm_handle* h, const OrbmBatchDevice* in, float th, int checkOri, int32_t* d_match,
                                        uint8_t* d_claimed, int32_t* d_nmatches, void* stream) {
    if (!h || !in || !d_match || !d_claimed || !d_nmatches || in->batch < 1 || in->batch > h->m.maxBatch || in->kcap > 65535 ||
        in->kcap > h->m.kcap || in->mcap > h->m.mcap || in->kcap < 1 || in->mcap < 1 || in->nlevels < 1 || !(in->maxX > in->minX) ||
        !(in->maxY > in->minY) || !in->kps || !in->desc || !in->nK || !in->nM || !in->scaleFactors) {
        set_error("orbm_search_last_frame_batch_device: bad argument (sizes, image bounds or null slabs)"); return ORB_ERR_ARG;
    }
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
    return m.run(P, in->kcap, in->mcap, (cudaStream_t)stream);
}

static int search_last_frame_batch_host(orbm_handle* h, const OrbmBatchDevice* in, float th, int checkOri, int32_t* match, uint8_t* claimed,
                                       int32_t* nmatches, bool frameOnDevice) {
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
    if (!frameOnDevice) { UP(kps, B * K * 28); UP(desc, B * K * 32); UP(nK, B * 4); }
    UP(scaleFactors, (size_t)in->nlevels * 4); UP(nM, B * 4);
    UP(valid, B * M); UP(xyz, B * M * 12); UP(octave, B * M * 4); UP(angle, B * M * 4); UP(hasObs, B * M); UP(mpDesc, B * M * 32); UP(Tcw7, B * 28);
#undef UP
    // with resetState the incoming match / claimed arrays are not read: do not ship them
    if ((rc = up(in->resetState ? nullptr : match, B * K * 4, &dm)) || (rc = up(in->resetState ? nullptr : claimed, B * K, &dc)) ||
        (rc = up(nullptr, B * 4, &dn))) return rc;
    rc = orbm_search_last_frame_batch_device(h, &d, th, checkOri, (int32_t*)dm, (uint8_t*)dc, (int32_t*)dn, st);
    if (rc) return rc;
    CK(cudaMemcpyAsync(match, dm, B * K * 4, cudaMemcpyDeviceToHost, st));
    CK(cudaMemcpyAsync(claimed, dc, B * K, cudaMemcpyDeviceToHost, st));
    CK(cudaMemcpyAsync(nmatches, dn, B * 4, cudaMemcpyDeviceToHost, st));
    CK(orbx::wait_stream_blocking(st, m.evWait));
    return ORB_OK;
}

int orbm_search_for_initialization(orbm_handle* h, const OrbmFrame* F1, const OrbmFrame* F2, float* prevMatched, int windowSize, float nnratio,
                                   int checkOrientation, int32_t* matches12, int* nmatches) {
    if (!h || !F1 || !F2 || !prevMatched || !matches12 || !nmatches || F1->K < 0 || F2->K < 0 || F1->K > h->m.mcap || F2->K > h->m.kcap ||
        F2->nlevels < 1 || !(F2->maxX > F2->minX) || !(F2->maxY > F2->minY) || windowSize < 0) {
        set_error("orbm_search_for_initialization: bad argument (F1.K <= max_mappoints, F2.K <= max_keypoints)"); return ORB_ERR_ARG;
    }
    Matcher& m = h->m;
    CK(cudaSetDevice(m.device));
    const size_t K1 = F1->K, K2 = F2->K;
    *nmatches = 0;
    for (size_t i = 0; i < K1; ++i) matches12[i] = -1;              // vnMatches12 = vector<int>(F1.mvKeysUn.size(),-1), :651
    if (K1 == 0) return ORB_OK;
    MatchParams P; memset(&P, 0, sizeof(P));
    Arena A(m.h_arena, m.d_arena, m.arenaBytes);
    std::vector<float> px(K1), py(K1), ang(K1); std::vector<int> lvl(K1);
    for (size_t i = 0; i < K1; ++i) { px[i] = prevMatched[2 * i]; py[i] = prevMatched[2 * i + 1]; ang[i] = F1->keypoints[i].angle; lvl[i] = F1->keypoints[i].octave; }
    int counts[2] = {F2->K, F1->K};
    const int* dCounts; const int32_t* dMatch; const int* dN;
    bool ok = A.put(F2->keypoints, K2, &P.kps) && A.put(F2->descriptors, K2 * 32, &P.desc) && A.put(F1->descriptors, K1 * 32, &P.mpDesc) &&
              A.put(px.data(), K1, &P.projX) && A.put(py.data(), K1, &P.projY) && A.put(ang.data(), K1, &P.angle) && A.put(lvl.data(), K1, &P.level) &&
              A.put(counts, (size_t)2, &dCounts);
    const size_t inBytes = A.off;
    ok = ok && A.put((const int32_t*)nullptr, K1, &dMatch) && A.put((const int*)nullptr, (size_t)1, &dN);
    if (!ok) { set_error("matcher staging arena too small"); return ORB_ERR_CAPACITY; }
    P.nK = dCounts; P.nM = dCounts + 1;
    P.batch = 1; P.kcap = m.kcap; P.mcap = m.mcap; P.nlevels = F2->nlevels; P.mode = 2;
    P.minX = F2->minX; P.minY = F2->minY; P.maxX = F2->maxX; P.maxY = F2->maxY;
    P.th = (float)windowSize; P.nnratio = nnratio; P.checkOri = checkOrientation;
    P.match = const_cast<int32_t*>(dMatch); P.nmatches = const_cast<int*>(dN);
    cudaStream_t st = m.stream;
    CK(cudaMemcpyAsync(m.d_arena, m.h_arena, inBytes, cudaMemcpyHostToDevice, st));
    int rc = m.run_init(P, F2->K, F1->K, st);
    if (rc) return rc;
    std::vector<float4> q(K1);
    CK(cudaMemcpyAsync(matches12, dMatch, 4 * K1, cudaMemcpyDeviceToHost, st));
    CK(cudaMemcpyAsync(nmatches, dN, 4, cudaMemcpyDeviceToHost, st));
    CK(cudaMemcpyAsync(q.data(), m.d_query, sizeof(float4) * K1, cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    for (size_t i = 0; i < K1; ++i) { prevMatched[2 * i] = q[i].x; prevMatched[2 * i + 1] = q[i].y; }
    return ORB_OK;
}

int orbm_search_last_frame_batch(orbm_handle* h, const OrbmBatchDevice* in, float th, int checkOri, int32_t* match, uint8_t* claimed,
                                 int32_t* nmatches) {
    return search_last_frame_batch_host(h, in, th, checkOri, match, claimed, nmatches, false);
}
int orbm_search_last_frame_batch_resident(orbm_handle* h, const OrbmBatchDevice* in, float th, int checkOri, int32_t* match, uint8_t* claimed,
                                          int32_t* nmatches) {
    return search_last_frame_batch_host(h, in, th, checkOri, match, claimed, nmatches, true);
}

static int search_by_bow_impl(orbm_handle* h, const OrbmBowFrame* KF, const uint8_t* kfPoint, const OrbmBowFrame* F, const uint8_t* fPoint, float nnratio,
                              int checkOrientation, int32_t* match, int32_t* match12, int* nmatches) {
    if (!h || !KF || !F || !kfPoint || (!match && !match12) || !nmatches || KF->N < 0 || F->N < 0 || KF->nEntries < 0 || F->nEntries < 0 || KF->nEntries > KF->N ||
        F->nEntries > F->N || KF->N > h->m.mcap || F->N > h->m.kcap) {
        set_error("orbm_search_by_bow: bad argument (KF.N <= max_mappoints, F.N <= max_keypoints of the handle)"); return ORB_ERR_ARG;
    }
    for (int e = 0; e < KF->nEntries; ++e) if (KF->fvFeature[e] < 0 || KF->fvFeature[e] >= KF->N || (e && KF->fvNode[e] < KF->fvNode[e - 1])) { set_error("orbm_search_by_bow: keyframe feature vector not sorted / out of range"); return ORB_ERR_ARG; }
    for (int e = 0; e < F->nEntries; ++e) if (F->fvFeature[e] < 0 || F->fvFeature[e] >= F->N || (e && F->fvNode[e] < F->fvNode[e - 1])) { set_error("orbm_search_by_bow: frame feature vector not sorted / out of range"); return ORB_ERR_ARG; }
    Matcher& m = h->m;
    CK(cudaSetDevice(m.device));
    *nmatches = 0;
    if (match) for (int i = 0; i < F->N; ++i) match[i] = -1;
    if (match12) for (int i = 0; i < KF->N; ++i) match12[i] = -1;
    if (F->N == 0 || KF->N == 0) return ORB_OK;
    Arena A(m.h_arena, m.d_arena, m.arenaBytes);
    BowMatchParams Q; memset(&Q, 0, sizeof(Q));
    Q.nKF = KF->N; Q.nF = F->N; Q.eKF = KF->nEntries; Q.eF = F->nEntries; Q.nnratio = nnratio; Q.checkOri = checkOrientation;
    bool ok = A.put(KF->keypoints, (size_t)KF->N, &Q.kpsKF) && A.put(KF->descriptors, (size_t)KF->N * 32, &Q.descKF) && A.put(kfPoint, (size_t)KF->N, &Q.kfPoint) &&
              A.put(KF->fvNode, (size_t)KF->nEntries, &Q.fvNodeKF) && A.put(KF->fvFeature, (size_t)KF->nEntries, &Q.fvFeatKF) &&
              A.put(F->keypoints, (size_t)F->N, &Q.kpsF) && A.put(F->descriptors, (size_t)F->N * 32, &Q.descF) &&
              A.put(F->fvNode, (size_t)F->nEntries, &Q.fvNodeF) && A.put(F->fvFeature, (size_t)F->nEntries, &Q.fvFeatF);
    if (fPoint) ok = ok && A.put(fPoint, (size_t)F->N, &Q.fPoint);
    const size_t inBytes = A.off;
    const int *dMatch, *dN, *dGs, *dM12 = nullptr; const uint8_t* dEv;
    ok = ok && A.put((const int*)nullptr, (size_t)F->N, &dMatch) && A.put((const int*)nullptr, (size_t)1, &dN) && A.put((const int*)nullptr, (size_t)KF->nEntries + 1, &dGs) &&
         A.put((const uint8_t*)nullptr, (size_t)F->N, &dEv);
    if (match12) ok = ok && A.put((const int*)nullptr, (size_t)KF->N, &dM12);
    if (!ok) { set_error("matcher staging arena too small"); return ORB_ERR_CAPACITY; }
    Q.match12 = const_cast<int*>(dM12); Q.strictLow = fPoint ? 1 : 0;
    Q.match = const_cast<int*>(dMatch); Q.nmatches = const_cast<int*>(dN); Q.groupStart = const_cast<int*>(dGs); Q.evBin = const_cast<uint8_t*>(dEv);
    cudaStream_t st = m.stream;
    CK(cudaMemcpyAsync(m.d_arena, m.h_arena, inBytes, cudaMemcpyHostToDevice, st));
    bow_match_kernel<<<1, BM_NT, 0, st>>>(Q);
    m.launches = 1;
    CK(cudaGetLastError());
    if (match) CK(cudaMemcpyAsync(match, dMatch, 4 * (size_t)F->N, cudaMemcpyDeviceToHost, st));
    if (match12) CK(cudaMemcpyAsync(match12, dM12, 4 * (size_t)KF->N, cudaMemcpyDeviceToHost, st));
    CK(cudaMemcpyAsync(nmatches, dN, 4, cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    return ORB_OK;
}

int orbm_search_by_bow(orbm_handle* h, const OrbmBowFrame* KF, const uint8_t* kfPoint, const OrbmBowFrame* F, float nnratio, int checkOrientation,
                       int32_t* match, int* nmatches) {
    return search_by_bow_impl(h, KF, kfPoint, F, nullptr, nnratio, checkOrientation, match, nullptr, nmatches);
}
int orbm_search_by_bow_kf(orbm_handle* h, const OrbmBowFrame* KF1, const uint8_t* point1, const OrbmBowFrame* KF2, const uint8_t* point2, float nnratio,
                          int checkOrientation, int32_t* match12, int* nmatches) {
    if (!point2 || !match12) { set_error("orbm_search_by_bow_kf: bad argument"); return ORB_ERR_ARG; }
    return search_by_bow_impl(h, KF1, point1, KF2, point2, nnratio, checkOrientation, nullptr, match12, nmatches);
}

static int project_search_impl(orbm_handle* h, const OrbmFrame* KF, const float* invLevelSigma2, float logScaleFactor, const float* Tcw7, const float* Ow3,
                               const float* cam4, const OrbmFusePoints* pts, float th, int mode, int32_t* bestIdx, int32_t* bestDist) {
    if (!h || !KF || (mode == 0 && (!Tcw7 || !Ow3)) || !cam4 || !pts || !bestIdx || !bestDist || pts->M < 0 || KF->K < 0 || KF->K > h->m.kcap ||
        !KF->scaleFactors || KF->nlevels < 1 || KF->nlevels > 256 || !(KF->maxX > KF->minX) || !(KF->maxY > KF->minY)) {
        set_error("orbm_fuse_search: bad argument (KF.K <= max_keypoints of the handle)"); return ORB_ERR_ARG;
    }
    Matcher& m = h->m;
    CK(cudaSetDevice(m.device));
    const size_t M = (size_t)pts->M, K = (size_t)KF->K;
    for (size_t i = 0; i < M; ++i) { bestIdx[i] = -1; bestDist[i] = 256; }
    if (M == 0 || K == 0) return ORB_OK;
    for (size_t k = 0; k < K; ++k) if (KF->keypoints[k].octave < 0 || KF->keypoints[k].octave >= KF->nlevels) { set_error("orbm_fuse_search: keypoint octave out of range"); return ORB_ERR_ARG; }
    size_t off = 0;
    auto take = [&](size_t bytes) { const size_t o = off; off = (off + bytes + 255) & ~(size_t)255; return o; };
    const size_t oKp = take(K * sizeof(OrbKeyPoint)), oDe = take(K * 32), oSf = take(4 * (size_t)KF->nlevels), oIs = take(4 * (size_t)KF->nlevels), oSt = take(M),
                 oXy = take(12 * M), oNo = take(12 * M), oMi = take(4 * M), oMa = take(4 * M), oMd = take(32 * M), oBi = take(4 * M), oBd = take(4 * M);
    if (off > m.batchBytes) {                                   // grow the device staging (mapping-side call: pageable copies are fine)
        if (m.d_batch) cudaFree(m.d_batch);
        m.d_batch = nullptr; m.batchBytes = 0;
        CK(cudaMalloc(&m.d_batch, off));
        m.batchBytes = off;
    }
    uint8_t* d = m.d_batch;
    cudaStream_t st = m.stream;
    CK(cudaMemcpyAsync(d + oKp, KF->keypoints, K * sizeof(OrbKeyPoint), cudaMemcpyHostToDevice, st));
    CK(cudaMemcpyAsync(d + oDe, KF->descriptors, K * 32, cudaMemcpyHostToDevice, st));
    CK(cudaMemcpyAsync(d + oSf, KF->scaleFactors, 4 * (size_t)KF->nlevels, cudaMemcpyHostToDevice, st));
    if (invLevelSigma2) CK(cudaMemcpyAsync(d + oIs, invLevelSigma2, 4 * (size_t)KF->nlevels, cudaMemcpyHostToDevice, st));
    CK(cudaMemcpyAsync(d + oSt, pts->state, M, cudaMemcpyHostToDevice, st));
    CK(cudaMemcpyAsync(d + oXy, pts->worldPos, 12 * M, cudaMemcpyHostToDevice, st));
    if (mode == 0) CK(cudaMemcpyAsync(d + oNo, pts->normal, 12 * M, cudaMemcpyHostToDevice, st));
    CK(cudaMemcpyAsync(d + oMi, pts->minDistance, 4 * M, cudaMemcpyHostToDevice, st));
    CK(cudaMemcpyAsync(d + oMa, pts->maxDistance, 4 * M, cudaMemcpyHostToDevice, st));
    CK(cudaMemcpyAsync(d + oMd, pts->descriptors, 32 * M, cudaMemcpyHostToDevice, st));
    FuseParams Q; memset(&Q, 0, sizeof(Q));
    MatchParams& P = Q.P;
    P.kps = (const OrbKeyPoint*)(d + oKp); P.desc = d + oDe; P.scaleFactors = (const float*)(d + oSf); P.kcap = m.kcap; P.nlevels = KF->nlevels; P.batch = 1;
    P.minX = KF->minX; P.minY = KF->minY; P.maxX = KF->maxX; P.maxY = KF->maxY;
    P.gridWInv = (float)GRID_COLS / (P.maxX - P.minX); P.gridHInv = (float)GRID_ROWS / (P.maxY - P.minY);
    P.th = th; memcpy(P.cam, cam4, 16);
    P.ks = (int)K; P.ms = 1;
    const size_t limit = 227 * 1024 - 2048;
    P.descInSmem = frame_smem_bytes(P.ks, 1, true) <= limit ? 1 : 0;
    const size_t sm = frame_smem_bytes(P.ks, 1, P.descInSmem != 0);
    if (sm > limit) { set_error("keyframe too large for the matcher's shared-memory grid"); return ORB_ERR_CAPACITY; }
    int rc = ensure_dynamic_smem(fuse_search_kernel, sm, m.device);
    if (rc) return rc;
    Q.K = (int)K; Q.M = (int)M; Q.state = d + oSt; Q.mpDesc = d + oMd; Q.xyz = (const float*)(d + oXy); Q.normal = (const float*)(d + oNo);
    Q.minD = (const float*)(d + oMi); Q.maxD = (const float*)(d + oMa); Q.invSigma2 = (const float*)(d + oIs);
    if (mode == 0) { memcpy(Q.Tcw, Tcw7, 28); memcpy(Q.Ow, Ow3, 12); }
    Q.logSF = logScaleFactor; Q.mode = mode; Q.gate = invLevelSigma2 ? 1 : 0;
    Q.bestIdx = (int*)(d + oBi); Q.bestDist = (int*)(d + oBd);
    const int grid = (int)std::min<size_t>(16, (M + 127) / 128);          // every CTA rebuilds the grid: a few CTAs, ~8 map points per warp
    fuse_search_kernel<<<grid, MF_NT, sm, st>>>(Q);
    m.launches = 1;
    CK(cudaGetLastError());
    CK(cudaMemcpyAsync(bestIdx, d + oBi, 4 * M, cudaMemcpyDeviceToHost, st));
    CK(cudaMemcpyAsync(bestDist, d + oBd, 4 * M, cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    return ORB_OK;
}

int orbm_fuse_search(orbm_handle* h, const OrbmFrame* KF, const float* invLevelSigma2, float logScaleFactor, const float* Tcw7, const float* Ow3,
                     const float* cam4, const OrbmFusePoints* pts, float th, int32_t* bestIdx, int32_t* bestDist) {
    if (!invLevelSigma2 || !pts || !pts->normal) { set_error("orbm_fuse_search: bad argument"); return ORB_ERR_ARG; }
    return project_search_impl(h, KF, invLevelSigma2, logScaleFactor, Tcw7, Ow3, cam4, pts, th, 0, bestIdx, bestDist);
}
int orbm_fuse_search_sim3(orbm_handle* h, const OrbmFrame* KF, float logScaleFactor, const float* Tcw7, const float* Ow3, const float* cam4, const OrbmFusePoints* pts,
                          float th, int32_t* bestIdx, int32_t* bestDist) {
    if (!pts || !pts->normal) { set_error("orbm_fuse_search_sim3: bad argument"); return ORB_ERR_ARG; }
    return project_search_impl(h, KF, nullptr, logScaleFactor, Tcw7, Ow3, cam4, pts, th, 0, bestIdx, bestDist);
}
int orbm_search_by_sim3(orbm_handle* h, const OrbmFrame* KF1, const OrbmFusePoints* pts1, const OrbmFrame* KF2, const OrbmFusePoints* pts2, float logScaleFactor,
                        const float* cam4, float th, const int32_t* pre12, int32_t* match12, int* nFound) {
    if (!h || !KF1 || !KF2 || !pts1 || !pts2 || !pre12 || !match12 || !nFound || pts1->M != KF1->K || pts2->M != KF2->K) {
        set_error("orbm_search_by_sim3: bad argument (one map-point slot per keyframe feature)"); return ORB_ERR_ARG;
    }
    const int N1 = KF1->K, N2 = KF2->K;
    for (int i = 0; i < N1; ++i) if (pre12[i] >= N2) { set_error("orbm_search_by_sim3: pre12 out of range"); return ORB_ERR_ARG; }
    // vbAlreadyMatched1 / 2 (:1478-1491) folded into the search flags
    std::vector<uint8_t> s1(N1), s2(N2);
    for (int i = 0; i < N1; ++i) s1[i] = pts1->state[i] == 1 && pre12[i] < 0;
    for (int i = 0; i < N2; ++i) s2[i] = pts2->state[i] == 1;
    for (int i = 0; i < N1; ++i) if (pre12[i] >= 0) s2[pre12[i]] = 0;
    std::vector<int32_t> m1(N1), d1(N1), m2(N2), d2(N2);
    OrbmFusePoints a = *pts1, b = *pts2;
    a.state = s1.data(); b.state = s2.data();
    int rc = project_search_impl(h, KF2, nullptr, logScaleFactor, nullptr, nullptr, cam4, &a, th, 2, m1.data(), d1.data());      // KF1's points in KF2 (:1497-1573)
    if (rc) return rc;
    rc = project_search_impl(h, KF1, nullptr, logScaleFactor, nullptr, nullptr, cam4, &b, th, 2, m2.data(), d2.data());          // KF2's points in KF1 (:1576-1652)
    if (rc) return rc;
    h->m.launches = 2;
    int n = 0;
    for (int i1 = 0; i1 < N1; ++i1) {                                                                                               // agreement (:1655-1671)
        match12[i1] = pre12[i1];
        const int idx2 = d1[i1] <= TH_HIGH ? m1[i1] : -1;
        if (idx2 >= 0 && (d2[idx2] <= TH_HIGH ? m2[idx2] : -1) == i1) { match12[i1] = idx2; ++n; }
    }
    *nFound = n;
    return ORB_OK;
}

static bool tri_frame_ok(const OrbmTriFrame* f) {
    if (!f || f->N < 0 || f->nEntries < 0 || f->nEntries > f->N) return false;
    if (f->N && (!f->keypoints || !f->descriptors || !f->hasMapPoint)) return false;
    if (f->nEntries && (!f->fvNode || !f->fvFeature)) return false;
    for (int e = 0; e < f->nEntries; ++e) if (f->fvFeature[e] < 0 || f->fvFeature[e] >= f->N || (e && f->fvNode[e] < f->fvNode[e - 1])) return false;
    return true;
}
int orbm_search_for_triangulation(orbm_handle* h, const OrbmTriFrame* KF1, int nKF2, const OrbmTriFrame* KF2, const float* scaleFactors, const float* levelSigma2,
                                  int nlevels, const float* ep2, const float* F12, int bCoarse, int checkOrientation, int32_t* matches12, int32_t* nmatches) {
    if (!h || nKF2 < 0 || !tri_frame_ok(KF1) || (nKF2 && (!KF2 || !ep2 || !F12 || !matches12 || !nmatches)) || !scaleFactors || !levelSigma2 || nlevels < 1 || nlevels > 256) {
        set_error("orbm_search_for_triangulation: bad argument (feature vectors must be sorted by node id, indices in range)"); return ORB_ERR_ARG;
    }
    for (int k = 0; k < nKF2; ++k) {
        if (!tri_frame_ok(KF2 + k)) { set_error("orbm_search_for_triangulation: bad KF2 entry"); return ORB_ERR_ARG; }
        for (int i = 0; i < KF2[k].N; ++i) if (KF2[k].keypoints[i].octave < 0 || KF2[k].keypoints[i].octave >= nlevels) { set_error("orbm_search_for_triangulation: keypoint octave out of range"); return ORB_ERR_ARG; }
    }
    if (nKF2 == 0) return ORB_OK;
    Matcher& m = h->m;
    CK(cudaSetDevice(m.device));
    const size_t N1 = (size_t)KF1->N, E1 = (size_t)KF1->nEntries;
    for (size_t i = 0; i < (size_t)nKF2 * N1; ++i) matches12[i] = -1;
    for (int k = 0; k < nKF2; ++k) nmatches[k] = 0;
    if (N1 == 0) return ORB_OK;
    // one staging buffer: tables | KF1 | per pair: KF2 arrays, outputs | pair table
    size_t off = 0;
    auto take = [&](size_t bytes) { const size_t o = off; off = (off + bytes + 255) & ~(size_t)255; return o; };
    const size_t oSf = take(4 * (size_t)nlevels), oSg = take(4 * (size_t)nlevels);
    const size_t oK1 = take(N1 * sizeof(OrbKeyPoint)), oD1 = take(N1 * 32), oM1 = take(N1), oN1 = take(4 * E1), oF1 = take(4 * E1);
    struct Off { size_t kp, de, mp, fn, ff, out, ev; };
    std::vector<Off> po(nKF2);
    for (int k = 0; k < nKF2; ++k) {
        const size_t N2 = (size_t)KF2[k].N, E2 = (size_t)KF2[k].nEntries;
        po[k] = {take(N2 * sizeof(OrbKeyPoint)), take(N2 * 32), take(N2), take(4 * E2), take(4 * E2), take(4 * N1), take(N1)};
    }
    const size_t oCnt = take(4 * (size_t)nKF2), oTab = take(sizeof(TriPair) * (size_t)nKF2);
    if (off > m.batchBytes) {                                   // grow the device staging (mapping-side call: pageable copies are fine)
        if (m.d_batch) cudaFree(m.d_batch);
        m.d_batch = nullptr; m.batchBytes = 0;
        CK(cudaMalloc(&m.d_batch, off));
        m.batchBytes = off;
    }
    uint8_t* d = m.d_batch;
    cudaStream_t st = m.stream;
    auto up = [&](size_t o, const void* src, size_t bytes) { return bytes ? cudaMemcpyAsync(d + o, src, bytes, cudaMemcpyHostToDevice, st) : cudaSuccess; };
    CK(up(oSf, scaleFactors, 4 * (size_t)nlevels)); CK(up(oSg, levelSigma2, 4 * (size_t)nlevels));
    CK(up(oK1, KF1->keypoints, N1 * sizeof(OrbKeyPoint))); CK(up(oD1, KF1->descriptors, N1 * 32)); CK(up(oM1, KF1->hasMapPoint, N1));
    CK(up(oN1, KF1->fvNode, 4 * E1)); CK(up(oF1, KF1->fvFeature, 4 * E1));
    std::vector<TriPair> tab(nKF2);
    for (int k = 0; k < nKF2; ++k) {
        const size_t N2 = (size_t)KF2[k].N, E2 = (size_t)KF2[k].nEntries;
        CK(up(po[k].kp, KF2[k].keypoints, N2 * sizeof(OrbKeyPoint))); CK(up(po[k].de, KF2[k].descriptors, N2 * 32)); CK(up(po[k].mp, KF2[k].hasMapPoint, N2));
        CK(up(po[k].fn, KF2[k].fvNode, 4 * E2)); CK(up(po[k].ff, KF2[k].fvFeature, 4 * E2));
        TriPair& T = tab[k];
        T.N2 = (int)N2; T.E2 = (int)E2;
        T.kps2 = (const OrbKeyPoint*)(d + po[k].kp); T.desc2 = d + po[k].de; T.hasMP2 = d + po[k].mp; T.fvNode2 = (const int*)(d + po[k].fn); T.fvFeat2 = (const int*)(d + po[k].ff);
        memcpy(T.ep, ep2 + 2 * (size_t)k, 8); memcpy(T.F12, F12 + 9 * (size_t)k, 36);
        T.matches12 = (int*)(d + po[k].out); T.nmatches = (int*)(d + oCnt) + k; T.evBin = d + po[k].ev;
    }
    CK(cudaMemcpyAsync(d + oTab, tab.data(), sizeof(TriPair) * (size_t)nKF2, cudaMemcpyHostToDevice, st));
    CK(cudaStreamSynchronize(st));                              // `tab` is pageable host memory: the copy must be done before it goes out of scope
    TriParams Q; memset(&Q, 0, sizeof(Q));
    Q.N1 = (int)N1; Q.E1 = (int)E1; Q.coarse = bCoarse; Q.checkOri = checkOrientation;
    Q.kps1 = (const OrbKeyPoint*)(d + oK1); Q.desc1 = d + oD1; Q.hasMP1 = d + oM1; Q.fvNode1 = (const int*)(d + oN1); Q.fvFeat1 = (const int*)(d + oF1);
    Q.scaleFactors = (const float*)(d + oSf); Q.levelSigma2 = (const float*)(d + oSg); Q.pairs = (const TriPair*)(d + oTab);
    tri_match_kernel<<<nKF2, TM_NT, 0, st>>>(Q);
    m.launches = 1;
    CK(cudaGetLastError());
    for (int k = 0; k < nKF2; ++k) CK(cudaMemcpyAsync(matches12 + (size_t)k * N1, d + po[k].out, 4 * N1, cudaMemcpyDeviceToHost, st));
    CK(cudaMemcpyAsync(nmatches, d + oCnt, 4 * (size_t)nKF2, cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    return ORB_OK;
}

int orbm_distinctive_descriptors(orbm_handle* h, int nPoints, const int32_t* obsStart, const uint8_t* descriptors, int32_t* best) {
    if (!h || nPoints < 0 || !obsStart || !best || (nPoints && obsStart[nPoints] > 0 && !descriptors)) { set_error("orbm_distinctive_descriptors: bad argument"); return ORB_ERR_ARG; }
    if (nPoints == 0) return ORB_OK;
    for (int p = 0; p < nPoints; ++p) if (obsStart[p + 1] < obsStart[p] || obsStart[p + 1] - obsStart[p] > 65535) { set_error("orbm_distinctive_descriptors: obsStart must be non-decreasing (<= 65535 observations per point)"); return ORB_ERR_ARG; }
    Matcher& m = h->m;
    CK(cudaSetDevice(m.device));
    const size_t total = (size_t)obsStart[nPoints];
    const size_t need = 4 * ((size_t)nPoints + 1) + 32 * total + 4 * (size_t)nPoints + 1024;
    if (need > m.batchBytes) {                                  // grow the device staging (cudaMemcpy from pageable host memory: rare, mapping-side call)
        if (m.d_batch) cudaFree(m.d_batch);
        m.d_batch = nullptr; m.batchBytes = 0;
        CK(cudaMalloc(&m.d_batch, need));
        m.batchBytes = need;
    }
    uint8_t* d = m.d_batch;
    const size_t oS = 0, oD = (4 * ((size_t)nPoints + 1) + 255) & ~(size_t)255, oB = (oD + 32 * total + 255) & ~(size_t)255;
    cudaStream_t st = m.stream;
    CK(cudaMemcpyAsync(d + oS, obsStart, 4 * ((size_t)nPoints + 1), cudaMemcpyHostToDevice, st));
    if (total) CK(cudaMemcpyAsync(d + oD, descriptors, 32 * total, cudaMemcpyHostToDevice, st));
    distinctive_kernel<<<(nPoints + 7) / 8, 256, 0, st>>>(nPoints, (const int*)(d + oS), d + oD, (int*)(d + oB));
    m.launches = 1;
    CK(cudaGetLastError());
    CK(cudaMemcpyAsync(best, d + oB, 4 * (size_t)nPoints, cudaMemcpyDeviceToHost, st));
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
