// B200 kernels + C-ABI for the numeric core of Optimizer::LocalBundleAdjustment
// (reference src/Optimizer.cc:1116-1498; g2o BlockSolver_6_3 + Levenberg, see include/orb_b200.h and SURVEY.md 3.2).
//
// Execution model: ONE persistent kernel per batch of problems.  Each problem is owned by one thread-block cluster
// (1..8 CTAs, chosen so that the batch fills the 148 SMs); the whole Levenberg-Marquardt loop -- accept/reject,
// lambda schedule, stop rules, stop-flag polling -- runs on the device, phases are separated by cluster barriers,
// and the host only uploads the flattened graph and downloads the result.  All arithmetic is FP64 (g2o is double).
// The linearisation is kept FACTORED: an edge stores only (x/z, y/z, 1/z, robust weight) at the linearisation point (32 B);
// its Jacobians are A = -S t (2x3, t = [R0 - xn R2; R1 - yn R2], S = diag(fx, fy)/z) and B = F b (2x6, F = diag(fx, fy),
// b a polynomial in xn, yn, 1/z), so the Hpl block W = w B^T A is rank 2 and is never materialised: every product that
// g2o forms with W (Schur complement, rhs, back-substitution) is evaluated from the factors.  Per LM trial the working
// set is ~1.5 MB per problem (L2 resident for a whole batch) instead of 11.5 MB of W / W Hll^-1 blocks.
// Phases of one LM trial (workers = all threads of the cluster):
//   residual   thread/edge      EdgeSE3ProjectXYZ::computeError + Huber rho, edge factors   (HOT LOOP A)
//   build      8 lanes/point: Hll, bl;  warp/chunk of a pose's edges: Hpp, bp               (HOT LOOP B)
//   schur      thread/point: (Hll + lambda I)^-1;  warp/chunk of (e1, e2) pairs of one pose pair:
//              Hschur = Hpp - sum B1^T (w1 w2 A1 Hll^-1 A2^T) B2                            (HOT LOOP C)
//   ldlt       CTA 0, matrix in shared memory: blocked LDL^T with look-ahead, rhs carried as an extra row
//   backsub    8 lanes/point    x_l = Hll^-1 (b_l - W^T x_p)
//   update     thread/vertex    T <- exp(dx) T, p <- p + dx  (+ backup for the LM "pop")
// Every reduction is ordered (no floating-point atomics), so results are reproducible run to run.
#include <cooperative_groups.h>
#include <cuda_runtime.h>
#include <float.h>
#include <math.h>
#include <string.h>
#include <algorithm>
#include <atomic>
#include <string>
#include <thread>
#include <vector>

#include "../../include/orb_b200.h"
#include "host_affinity.h"
#include "device_utils.cuh"

using orbx::set_error;

#define CK(call)                                                                   \
    do {                                                                           \
        cudaError_t e_ = (call);                                                   \
        if (e_ != cudaSuccess) {                                                   \
            set_error(std::string(#call) + ": " + cudaGetErrorString(e_));         \
            return ORB_ERR_CUDA;                                                   \
        }                                                                          \
    } while (0)

namespace lba {

struct Dev {
    int nP, nL, nE, nF, n;            // n = 6 nF
    double *poses, *posesB, *pts, *ptsB;        // nP x 7, nL x 3: estimate / second buffer (trial state; the two swap when a step is accepted)
    const double *initPoses, *initPts;          // the uploaded initial estimates (a run always restarts from them)
    // the caller's graph as uploaded (caller's edge order); lba_build_structure_kernel derives everything below from it
    const int *rawEdgePoint, *rawEdgePose; const double* rawObs; const float* rawInvSigma2;
    int *bCnt, *bCursor, *bTmp, *bEid, *bTaskLen, *bItemStart;   // build scratch: nL + 1, nL + 1, nE, nF x nL, nTasks, nTasks
    int pairCap, chunkCap;
    const float* cam;                 // nP x 4
    const int* hidx;                  // nP: Hessian block index or -1
    const int* freePose;              // nF: pose index of Hessian block
    // edges in INTERNAL order = sorted by point (the edges of a point are contiguous)
    const int *ePt, *ePose;           // nE
    const int* eOrig;                 // nE: caller's edge index
    const double* obs;                // nE x 2
    const float* invSigma2;           // nE
    const int* ptStart;               // nL + 1
    const int* poseEdges;             // edge lists of the free poses (internal edge ids, ascending), back to back
    const int2* pairs;                // (edge of pose i1, edge of pose i2) observing the same point
    const int* pairPt;                // that point
    // Schur work list: tasks 0..nF-1 = diagonal blocks (items = edges of the pose), nF.. = off-diagonal blocks (items = pairs);
    // every task is cut into chunks of SCH items so that all warps of the cluster get the same amount of work
    int nChunks; const int4* chunkHdr; const int* taskChunkStart;   // nChunks x (first item, items, pose A, pose B or -1), nTasks + 1
    const int* poseEdgePt;            // nE: point of poseEdges[k]
    double* Spart;                    // nChunks x 42 partial sums (36 block entries + 6 rhs entries for diagonal tasks)
    double* Ppart;                    // (chunks of the diagonal tasks) x 28: partial sums of Hpp (21) and bp (6)
    double* err;                      // nE x 2
    double *E4a, *E4b;                // nE x 4 edge factors (x/z, y/z, 1/z, w): linearisation point / trial state (swapped on accept)
    double *Hpp, *bp;                 // nF x 36, nF x 6
    double *Hll, *bl;                 // nL x 6 (00 01 02 11 12 22), nL x 3
    double* PT;                       // nL x 12 point records: upper triangle of (Hll + lambda I)^-1 (6), (Hll + lambda I)^-1 bl (3), 3 pad
    double *Hs, *bs;                  // n x n (only when it does not fit in shared memory), n
    double* x;                        // n + 3 nL
    double* partial;                  // 4 rotating slots x 16 doubles: per-CTA partial sums + broadcast words
    double* stats;                    // out: [0] iterations [1] trials [2] lambda [3] chi2 [4] initial chi2, [8..17] phase ns (CTA 0)
    double* outChi2; uint8_t* outDepthPos;   // nE, caller's edge order
    double delta, dsqr, userLambdaInit;
    int iterations;
};

__device__ __forceinline__ void qrot(const double* q, const double* v, double* o) {   // q = (w,x,y,z)
    double ux = q[2] * v[2] - q[3] * v[1], uy = q[3] * v[0] - q[1] * v[2], uz = q[1] * v[1] - q[2] * v[0];
    ux += ux; uy += uy; uz += uz;
    o[0] = v[0] + q[0] * ux + (q[2] * uz - q[3] * uy);
    o[1] = v[1] + q[0] * uy + (q[3] * ux - q[1] * uz);
    o[2] = v[2] + q[0] * uz + (q[1] * uy - q[2] * ux);
}
__device__ __forceinline__ void qnormalize(double* q) {
    if (q[0] < 0) { q[0] = -q[0]; q[1] = -q[1]; q[2] = -q[2]; q[3] = -q[3]; }
    const double n = sqrt(q[0] * q[0] + q[1] * q[1] + q[2] * q[2] + q[3] * q[3]);
    q[0] /= n; q[1] /= n; q[2] /= n; q[3] /= n;
}
__device__ __forceinline__ void qtoR(const double* q, double* R) {
    const double tx = 2 * q[1], ty = 2 * q[2], tz = 2 * q[3];
    const double twx = tx * q[0], twy = ty * q[0], twz = tz * q[0], txx = tx * q[1], txy = ty * q[1], txz = tz * q[1];
    const double tyy = ty * q[2], tyz = tz * q[2], tzz = tz * q[3];
    R[0] = 1 - (tyy + tzz); R[1] = txy - twz; R[2] = txz + twy;
    R[3] = txy + twz; R[4] = 1 - (txx + tzz); R[5] = tyz - twx;
    R[6] = txz - twy; R[7] = tyz + twx; R[8] = 1 - (txx + tyy);
}
__device__ __forceinline__ void qfromR(const double* m, double* q) {
    double t = m[0] + m[4] + m[8];
    if (t > 0) {
        t = sqrt(t + 1.0);
        q[0] = 0.5 * t; t = 0.5 / t;
        q[1] = (m[7] - m[5]) * t; q[2] = (m[2] - m[6]) * t; q[3] = (m[3] - m[1]) * t;
    } else {
        int i = 0;
        if (m[4] > m[0]) i = 1;
        if (m[8] > m[i * 3 + i]) i = 2;
        const int j = (i + 1) % 3, k = (j + 1) % 3;
        t = sqrt(m[i * 3 + i] - m[j * 3 + j] - m[k * 3 + k] + 1.0);
        double v[3];
        v[i] = 0.5 * t; t = 0.5 / t;
        q[0] = (m[k * 3 + j] - m[j * 3 + k]) * t;
        v[j] = (m[j * 3 + i] + m[i * 3 + j]) * t;
        v[k] = (m[k * 3 + i] + m[i * 3 + k]) * t;
        q[1] = v[0]; q[2] = v[1]; q[3] = v[2];
    }
}
// VertexSE3Expmap::oplusImpl: T <- SE3Quat::exp(upd) * T   (g2o/types/se3quat.h:223-256, :101-110)
__device__ void pose_oplus(double* T, const double* upd) {
    const double om0 = upd[0], om1 = upd[1], om2 = upd[2];
    const double theta = sqrt(om0 * om0 + om1 * om1 + om2 * om2);
    const double O[9] = {0, -om2, om1, om2, 0, -om0, -om1, om0, 0};
    double O2[9];
    for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) O2[i * 3 + j] = O[i * 3] * O[j] + O[i * 3 + 1] * O[3 + j] + O[i * 3 + 2] * O[6 + j];
    double a, b, c, d;
    if (theta < 0.00001) { a = 1.0; b = 1 / 2.0; c = 1 / 2.0; d = 1 / 6.0; }
    else {
        a = sin(theta) / theta; b = (1 - cos(theta)) / (theta * theta);
        c = b; d = (theta - sin(theta)) / pow(theta, 3.0);
    }
    double R[9], V[9];
    for (int i = 0; i < 9; ++i) {
        const double I = (i % 4 == 0) ? 1.0 : 0.0;
        R[i] = I + a * O[i] + b * O2[i];
        V[i] = I + c * O[i] + d * O2[i];
    }
    double qe[4]; qfromR(R, qe);
    const double te[3] = {V[0] * upd[3] + V[1] * upd[4] + V[2] * upd[5], V[3] * upd[3] + V[4] * upd[4] + V[5] * upd[5],
                          V[6] * upd[3] + V[7] * upd[4] + V[8] * upd[5]};
    qnormalize(qe);
    double rt[3]; qrot(qe, T + 4, rt);
    const double* q2 = T;
    double qn[4] = {qe[0] * q2[0] - qe[1] * q2[1] - qe[2] * q2[2] - qe[3] * q2[3], qe[0] * q2[1] + qe[1] * q2[0] + qe[2] * q2[3] - qe[3] * q2[2],
                    qe[0] * q2[2] + qe[2] * q2[0] + qe[3] * q2[1] - qe[1] * q2[3], qe[0] * q2[3] + qe[3] * q2[0] + qe[1] * q2[2] - qe[2] * q2[1]};
    qnormalize(qn);
    T[0] = qn[0]; T[1] = qn[1]; T[2] = qn[2]; T[3] = qn[3];
    T[4] = te[0] + rt[0]; T[5] = te[1] + rt[1]; T[6] = te[2] + rt[2];
}

__device__ __forceinline__ void robustify(const Dev& D, double e2, double& rho0, double& rho1) {   // RobustKernelHuber
    if (e2 <= D.dsqr) { rho0 = e2; rho1 = 1.; }
    else { const double s = sqrt(e2); rho0 = 2 * s * D.delta - D.dsqr; rho1 = D.delta / s; }
}


namespace cg = cooperative_groups;
__device__ __forceinline__ unsigned long long globaltimer_ns() { unsigned long long t; asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t)); return t; }
#ifndef LBA_NT
#define LBA_NT 512
#endif
#ifndef LBA_MINB
#define LBA_MINB 1     /* __launch_bounds__ min blocks: with LBA_NT=256, 2 caps the kernel at 128 registers = half an SM's file */
#endif
#ifndef LBA_SCH
#define LBA_SCH 128
#endif
#ifndef LBA_UNR
#define LBA_UNR 4
#endif
constexpr int NT = LBA_NT;
constexpr int NWARP = NT / 32;
constexpr int MAXC = 8;            // largest cluster
constexpr int PSLOT = 24;          // doubles per partial slot: A[0..7], flag[8], B[9..16]
constexpr int PC = 20;             // doubles per cached pose: q(4) t(3) R(9) fx fy cx cy

struct Ctx {
    int crank, csize, tid;
    int wid, nw;           // worker id / count over the cluster
    int slot;              // rotating partial slot
    double* sm;            // NT doubles of scratch
    const double* pc;      // pose cache in shared memory (state at the linearisation point), PC doubles per pose
    const double* pts;     // point estimates of that state
};

// ordered block reduction of one double per thread; every thread returns the sum
__device__ __forceinline__ double block_sum(double v, double* sm) {
    const int tid = threadIdx.x;
#pragma unroll
    for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if ((tid & 31) == 0) sm[tid >> 5] = v;
    __syncthreads();
    double r = 0;
#pragma unroll
    for (int w = 0; w < NWARP; ++w) r += sm[w];
    __syncthreads();
    return r;
}
// cluster-wide ordered sum (per-CTA partials -> global slot -> cluster barrier -> added in rank order by every thread);
// `flagIn` of CTA 0 / thread 0 is broadcast alongside.
__device__ __forceinline__ double cluster_sum(const Dev& D, Ctx& c, double local, int flagIn, int& flagOut) {
    const double s = block_sum(local, c.sm);
    double* slot = D.partial + (size_t)(c.slot & 3) * PSLOT;
    if (c.tid == 0) {
        slot[c.crank] = s;
        if (c.crank == 0) slot[MAXC] = (double)flagIn;
    }
    cg::this_cluster().sync();
    double tot = 0;
    for (int r = 0; r < c.csize; ++r) tot += slot[r];
    flagOut = (int)slot[MAXC];
    ++c.slot;
    return tot;
}
// the same for two sums at once
__device__ __forceinline__ void cluster_sum2(const Dev& D, Ctx& c, double a, double b, int flagIn, int& flagOut, double& totA, double& totB) {
    const int tid = threadIdx.x;
#pragma unroll
    for (int o = 16; o; o >>= 1) { a += __shfl_xor_sync(0xffffffffu, a, o); b += __shfl_xor_sync(0xffffffffu, b, o); }
    if ((tid & 31) == 0) { c.sm[tid >> 5] = a; c.sm[NWARP + (tid >> 5)] = b; }
    __syncthreads();
    double* slot = D.partial + (size_t)(c.slot & 3) * PSLOT;
    if (tid == 0) {
        double ra = 0, rb = 0;
        for (int w = 0; w < NWARP; ++w) { ra += c.sm[w]; rb += c.sm[NWARP + w]; }
        slot[c.crank] = ra; slot[MAXC + 1 + c.crank] = rb;
        if (c.crank == 0) slot[MAXC] = (double)flagIn;
    }
    cg::this_cluster().sync();
    totA = 0; totB = 0;
    for (int r = 0; r < c.csize; ++r) { totA += slot[r]; totB += slot[MAXC + 1 + r]; }
    flagOut = (int)slot[MAXC];
    ++c.slot;
}
__device__ __forceinline__ void csync() { cg::this_cluster().sync(); }

// (re)load every pose into the CTA's shared-memory cache: quaternion, translation, rotation matrix, camera
__device__ void load_pose_cache(const Dev& D, const double* poses, double* pc) {
    for (int i = threadIdx.x; i < D.nP; i += NT) {
        double* o = pc + PC * i;
        const double* T = poses + 7 * (size_t)i;
#pragma unroll
        for (int k = 0; k < 7; ++k) o[k] = T[k];
        qtoR(T, o + 7);
        const float* cm = D.cam + 4 * (size_t)i;
        o[16] = (double)cm[0]; o[17] = (double)cm[1]; o[18] = (double)cm[2]; o[19] = (double)cm[3];   // float params promoted (Pinhole.cpp:35-41)
    }
    __syncthreads();
}

// one edge at pose P (cached: q t R cam) and point X: residual, robust chi2 and the edge factors of this state
__device__ __forceinline__ double edge_error(const Dev& D, const double* P, const double* X, int e, double* __restrict__ Eout) {
    double r[3]; qrot(P, X, r);                                     // SE3Quat::map: _r * xyz + _t
    const double xc = r[0] + P[4], yc = r[1] + P[5], zc = r[2] + P[6];
    const double u = P[16] * xc / zc + P[18], v = P[17] * yc / zc + P[19];      // Pinhole::project(Vector3d)
    const double e0 = D.obs[2 * (size_t)e] - u, e1 = D.obs[2 * (size_t)e + 1] - v;
    *reinterpret_cast<double2*>(D.err + 2 * (size_t)e) = make_double2(e0, e1);
    const double is2 = (double)D.invSigma2[e];
    double r0, r1;
    robustify(D, is2 * (e0 * e0 + e1 * e1), r0, r1);
    const double iz = 1.0 / zc;
    double2* o = reinterpret_cast<double2*>(Eout + 4 * (size_t)e);
    o[0] = make_double2(xc * iz, yc * iz);
    o[1] = make_double2(iz, r1 * is2);
    return r0;
}
// ---- residuals + robust chi2 (SparseOptimizer::computeActiveErrors + activeRobustChi2), thread per edge.
//      Also leaves the edge factors (x/z, y/z, 1/z, rho' / sigma^2) of this state in Eout. ----
__device__ double phase_errors(const Dev& D, const Ctx& c, double* __restrict__ Eout) {
    double acc = 0;
    for (int e = c.wid; e < D.nE; e += c.nw) acc += edge_error(D, c.pc + PC * D.ePose[e], c.pts + 3 * (size_t)D.ePt[e], e, Eout);
    return acc;
}

// Edge factors.  With xn = x/z, yn = y/z, iz = 1/z (camera frame) and R the keyframe rotation, the Jacobians of
// EdgeSE3ProjectXYZ::linearizeOplus (-projectJac * R and -projectJac * [-[Xc]x | I], Pinhole.cpp:71-81) are
//   A = -diag(fx iz, fy iz) t,   t = [R0 - xn R2 ; R1 - yn R2]                           (2x3)
//   B =  diag(fx, fy) b,         b = [xn yn, -(1 + xn^2), yn, -iz, 0, xn iz ; 1 + yn^2, -xn yn, -xn, 0, -iz, yn iz]   (2x6)
// 256-bit global load (LDG.E.256, sm_100+): one L1 tag look-up per lane for a 32-byte record -- the gathers of the Schur
// phase are bound by look-ups per instruction, not by bytes
__device__ __forceinline__ void ld256(const double* p, double& a, double& b, double& c, double& d) {
    asm volatile("ld.global.v4.f64 {%0, %1, %2, %3}, [%4];" : "=d"(a), "=d"(b), "=d"(c), "=d"(d) : "l"(p));
}
__device__ __forceinline__ void load_e4(const double* __restrict__ E, int e, double& xn, double& yn, double& iz, double& w) {
    ld256(E + 4 * (size_t)e, xn, yn, iz, w);
}
__device__ __forceinline__ void edge_t(const double* R, double xn, double yn, double* t) {
#pragma unroll
    for (int k = 0; k < 3; ++k) { t[k] = R[k] - xn * R[6 + k]; t[3 + k] = R[3 + k] - yn * R[6 + k]; }
}
__device__ __forceinline__ void edge_b(double xn, double yn, double iz, double* b0, double* b1) {   // b0[4] and b1[3] are zero
    const double xy = xn * yn;
    b0[0] = xy;                 b0[1] = -(1.0 + xn * xn); b0[2] = yn;  b0[3] = -iz; b0[4] = 0.0; b0[5] = xn * iz;
    b1[0] = 1.0 + yn * yn;      b1[1] = -xy;              b1[2] = -xn; b1[3] = 0.0; b1[4] = -iz; b1[5] = yn * iz;
}

// Reduce-scatter of V register values over the 32 lanes of a warp: afterwards lane `lane` holds the totals of entries
// [lo, lo + n) in v[0..n), n <= 2.  Fixed butterfly (halve the vector at every level), ~V shuffles instead of 5 V.
template <int N, int H>
__device__ __forceinline__ void rs_level(double* v, bool up, int o) {
#pragma unroll
    for (int i = 0; i < H; ++i) {
        const double a = v[i];
        const double b = (i + H < N) ? v[i + H] : 0.0;
        const double keep = up ? b : a, send = up ? a : b;
        v[i] = keep + __shfl_xor_sync(0xffffffffu, send, o);
    }
}
template <int V>
__device__ __forceinline__ void warp_reduce_scatter(double* v, int lane, int& lo, int& n) {
    constexpr int H1 = (V + 1) / 2, H2 = (H1 + 1) / 2, H3 = (H2 + 1) / 2, H4 = (H3 + 1) / 2, H5 = (H4 + 1) / 2;
    static_assert(H5 <= 2, "vector too long");
    lo = 0; n = V;
    bool up;
    up = lane & 16; rs_level<V, H1>(v, up, 16); if (up) { lo += H1; n -= H1; } else n = min(n, H1);
    up = lane & 8;  rs_level<H1, H2>(v, up, 8); if (up) { lo += H2; n -= H2; } else n = min(n, H2);
    up = lane & 4;  rs_level<H2, H3>(v, up, 4); if (up) { lo += H3; n -= H3; } else n = min(n, H3);
    up = lane & 2;  rs_level<H3, H4>(v, up, 2); if (up) { lo += H4; n -= H4; } else n = min(n, H4);
    up = lane & 1;  rs_level<H4, H5>(v, up, 1); if (up) { lo += H5; n -= H5; } else n = min(n, H5);
}

__device__ __forceinline__ void point_record(const Dev& D, int p, const double* h, const double* bl3, double lambda);
// ---- per point Hll, bl (and, when lambda is already known, the point record of phase_point_prep); 8 lanes per point
//      (the edges of a point are contiguous) ----
__device__ void phase_build_points(const Dev& D, const Ctx& c, const double* __restrict__ E, double lambda) {
    const int sl = c.tid & 7;
    const unsigned gmask = 0xFFu << (c.tid & 24);
    for (int p = c.wid >> 3; p < D.nL; p += c.nw >> 3) {
        const int a = D.ptStart[p], b = D.ptStart[p + 1];
        double h[6] = {0, 0, 0, 0, 0, 0}, g[3] = {0, 0, 0};
        for (int e = a + sl; e < b; e += 8) {
            const double* P = c.pc + PC * D.ePose[e];
            double xn, yn, iz, w, t[6];
            load_e4(E, e, xn, yn, iz, w);
            edge_t(P + 7, xn, yn, t);
            const double2 er = *reinterpret_cast<const double2*>(D.err + 2 * (size_t)e);
            const double s0 = P[16] * iz, s1 = P[17] * iz;
            const double c0 = w * s0 * s0, c1 = w * s1 * s1;          // w A^T A = c0 t0^T t0 + c1 t1^T t1
            const double q0 = s0 * w * er.x, q1 = s1 * w * er.y;      // A^T r, r = -w err
            g[0] += t[0] * q0 + t[3] * q1; g[1] += t[1] * q0 + t[4] * q1; g[2] += t[2] * q0 + t[5] * q1;
            h[0] += c0 * t[0] * t[0] + c1 * t[3] * t[3]; h[1] += c0 * t[0] * t[1] + c1 * t[3] * t[4]; h[2] += c0 * t[0] * t[2] + c1 * t[3] * t[5];
            h[3] += c0 * t[1] * t[1] + c1 * t[4] * t[4]; h[4] += c0 * t[1] * t[2] + c1 * t[4] * t[5]; h[5] += c0 * t[2] * t[2] + c1 * t[5] * t[5];
        }
#pragma unroll
        for (int o = 4; o; o >>= 1) {
#pragma unroll
            for (int i = 0; i < 6; ++i) h[i] += __shfl_xor_sync(gmask, h[i], o);
#pragma unroll
            for (int i = 0; i < 3; ++i) g[i] += __shfl_xor_sync(gmask, g[i], o);
        }
        if (sl == 0) {
            double2* H = reinterpret_cast<double2*>(D.Hll + 6 * (size_t)p);
            H[0] = make_double2(h[0], h[1]); H[1] = make_double2(h[2], h[3]); H[2] = make_double2(h[4], h[5]);
            D.bl[3 * (size_t)p] = g[0]; D.bl[3 * (size_t)p + 1] = g[1]; D.bl[3 * (size_t)p + 2] = g[2];
            if (lambda >= 0) point_record(D, p, h, g, lambda);
        }
    }
}

constexpr int SCH = LBA_SCH;       // items per chunk of the pose / pose-pair work lists
constexpr int IPL = SCH / 32;      // items per lane
constexpr int UNR = LBA_UNR;       // items in flight per lane in the Schur chunk loops
constexpr int PTR = 12;            // doubles per point record: d00 d01 d02 d11 | d12 d22 db0 db1 | db2 - - -   (Dinv = (Hll + lambda I)^-1, db = Dinv bl)

// ---- per free pose Hpp, bp.  partial: warp per chunk of the pose's edges (the chunks of the diagonal Schur tasks), lane per
//      edge, 21 + 6 sums reduce-scattered to Spart; combine: thread per (pose, entry), chunks added in order ----
__device__ void phase_build_poses_partial(const Dev& D, const Ctx& c, const double* __restrict__ E) {
    const int lane = c.tid & 31;
    const int nDiag = D.taskChunkStart[D.nF];
    for (int ch = c.crank * NWARP + (c.tid >> 5); ch < nDiag; ch += c.csize * NWARP) {
        const int4 h = D.chunkHdr[ch];
        const double* P = c.pc + PC * h.z;
        const double fx = P[16], fy = P[17];
        double acc[27];
#pragma unroll
        for (int i = 0; i < 27; ++i) acc[i] = 0;
        for (int i = lane; i < h.y; i += 32) {
            const int e = D.poseEdges[h.x + i];
            double xn, yn, iz, w, b0[6], b1[6];
            load_e4(E, e, xn, yn, iz, w);
            edge_b(xn, yn, iz, b0, b1);
            const double2 er = *reinterpret_cast<const double2*>(D.err + 2 * (size_t)e);
            const double c0 = w * fx * fx, c1 = w * fy * fy;          // w B^T B = c0 b0^T b0 + c1 b1^T b1
            const double q0 = fx * w * er.x, q1 = fy * w * er.y;      // B^T r = -(b0 q0 + b1 q1)
            int t = 0;
#pragma unroll
            for (int a = 0; a < 6; ++a) {
                const double u0 = c0 * b0[a], u1 = c1 * b1[a];
#pragma unroll
                for (int j = a; j < 6; ++j) acc[t++] += u0 * b0[j] + u1 * b1[j];
            }
#pragma unroll
            for (int a = 0; a < 6; ++a) acc[21 + a] -= b0[a] * q0 + b1[a] * q1;
        }
        int lo, n;
        warp_reduce_scatter<27>(acc, lane, lo, n);
        if (n > 0) D.Ppart[28 * (size_t)ch + lo] = acc[0];
    }
}
__device__ void phase_build_poses_combine(const Dev& D, const Ctx& c) {
    for (int idx = c.wid; idx < D.nF * 27; idx += c.nw) {
        const int hI = idx / 27, ent = idx - hI * 27;
        double s = 0;
        for (int ch = D.taskChunkStart[hI]; ch < D.taskChunkStart[hI + 1]; ++ch) s += D.Ppart[28 * (size_t)ch + ent];
        if (ent < 21) {
            int i = 0, t = ent;             // unrank (i, j), i <= j
            while (t >= 6 - i) { t -= 6 - i; ++i; }
            const int j = i + t;
            D.Hpp[36 * (size_t)hI + i * 6 + j] = s; D.Hpp[36 * (size_t)hI + j * 6 + i] = s;
        } else D.bp[6 * (size_t)hI + (ent - 21)] = s;
    }
}

// ---- max |diag| over all Hessian blocks (computeLambdaInit); every CTA computes it redundantly ----
__device__ double phase_maxdiag(const Dev& D, const Ctx& c) {
    double m = 0;
    for (int i = c.tid; i < D.n; i += NT) m = fmax(m, fabs(D.Hpp[36 * (size_t)(i / 6) + 7 * (i % 6)]));
    for (int i = c.tid; i < 3 * D.nL; i += NT) { const int k = i % 3; m = fmax(m, fabs(D.Hll[6 * (size_t)(i / 3) + (k == 0 ? 0 : k == 1 ? 3 : 5)])); }
#pragma unroll
    for (int o = 16; o; o >>= 1) m = fmax(m, __shfl_xor_sync(0xffffffffu, m, o));
    __syncthreads();
    if ((c.tid & 31) == 0) c.sm[c.tid >> 5] = m;
    __syncthreads();
    double r = 0;
    for (int w = 0; w < NWARP; ++w) r = fmax(r, c.sm[w]);
    __syncthreads();
    return r;
}

// ---- (Hll + lambda I)^-1 and (Hll + lambda I)^-1 bl, thread per point (block_solver.hpp:381-394) ----
__device__ __forceinline__ void point_record(const Dev& D, int p, const double* h, const double* bl3, double lambda) {
    const double m0 = h[0] + lambda, m1 = h[1], m2 = h[2], m4 = h[3] + lambda, m5 = h[4], m8 = h[5] + lambda;
    const double c00 = m4 * m8 - m5 * m5, c10 = m5 * m2 - m1 * m8, c20 = m1 * m5 - m4 * m2;
    const double id = 1.0 / (m0 * c00 + m1 * c10 + m2 * c20);     // Eigen 3x3 inverse: cofactors / determinant
    const double o0 = c00 * id, o1 = c10 * id, o2 = c20 * id;
    const double o4 = (m0 * m8 - m2 * m2) * id, o5 = (m2 * m1 - m0 * m5) * id, o8 = (m0 * m4 - m1 * m1) * id;
    const double b0 = bl3[0], b1 = bl3[1], b2 = bl3[2];
    double2* O = reinterpret_cast<double2*>(D.PT + PTR * (size_t)p);
    O[0] = make_double2(o0, o1); O[1] = make_double2(o2, o4); O[2] = make_double2(o5, o8);
    O[3] = make_double2(o0 * b0 + o1 * b1 + o2 * b2, o1 * b0 + o4 * b1 + o5 * b2);
    O[4] = make_double2(o2 * b0 + o5 * b1 + o8 * b2, 0.0);
}
__device__ void phase_point_prep(const Dev& D, const Ctx& c, double lambda) {
    for (int p = c.wid; p < D.nL; p += c.nw) {
        const double2* H = reinterpret_cast<const double2*>(D.Hll + 6 * (size_t)p);
        const double2 h01 = H[0], h23 = H[1], h45 = H[2];
        const double h[6] = {h01.x, h01.y, h23.x, h23.y, h45.x, h45.y};
        const double b3[3] = {D.bl[3 * (size_t)p], D.bl[3 * (size_t)p + 1], D.bl[3 * (size_t)p + 2]};
        point_record(D, p, h, b3, lambda);
    }
}

// ---- Schur complement (block_solver.hpp:396-431), two steps.
//   partial: the work list (diagonal block i: edges of pose i; off-diagonal block (i1 < i2): the precomputed (e1, e2) pairs)
//            is cut into chunks of SCH items; one warp per chunk, one lane per item.  An item adds
//                b1^T M b2,   M[k][l] = g1_k g2_l (t1_k Hll^-1 t2_l^T),   g_k = w f_k^2 / z
//            (= W1 Hll^-1 W2^T with W = w B^T A) into the lane's 36 accumulators; diagonal items also add
//            W Hll^-1 bl = -sum_k b_k g_k (t_k . db).  The item's three records (two edges, one point) are per-lane
//            gathers -- the phase is bound by L1 look-ups (~39 clk per 32-lane LDG.256 of 32 distinct lines), so the
//            chunk's indices sit in a warp-private shared-memory buffer and the gathers of item j + 1 are issued before
//            the arithmetic of item j.  The lanes' sums are reduce-scattered into Spart.
//   combine: thread per (task, entry): ordered sum of the task's chunks -> lower triangle of the reduced camera system
//            (diagonal: Hpp_i + lambda I - sum, bs_i = bp_i - sum; off-diagonal: block (i2, i1) = -(sum)^T). ----
__device__ __forceinline__ void schur_item(const double* t1, const double* t2, const double* dv, double g10, double g11, double g20, double g21,
                                           const double* b10, const double* b11, const double* b20, const double* b21, double* acc) {
    // u = t1 Dinv (2x3), dv = 00 01 02 11 12 22
    const double u00 = t1[0] * dv[0] + t1[1] * dv[1] + t1[2] * dv[2], u01 = t1[0] * dv[1] + t1[1] * dv[3] + t1[2] * dv[4], u02 = t1[0] * dv[2] + t1[1] * dv[4] + t1[2] * dv[5];
    const double u10 = t1[3] * dv[0] + t1[4] * dv[1] + t1[5] * dv[2], u11 = t1[3] * dv[1] + t1[4] * dv[3] + t1[5] * dv[4], u12 = t1[3] * dv[2] + t1[4] * dv[4] + t1[5] * dv[5];
    const double m00 = (u00 * t2[0] + u01 * t2[1] + u02 * t2[2]) * (g10 * g20), m01 = (u00 * t2[3] + u01 * t2[4] + u02 * t2[5]) * (g10 * g21);
    const double m10 = (u10 * t2[0] + u11 * t2[1] + u12 * t2[2]) * (g11 * g20), m11 = (u10 * t2[3] + u11 * t2[4] + u12 * t2[5]) * (g11 * g21);
    // T = M b2 (2x6); b20[4] = b21[3] = 0
    double T0[6], T1[6];
#pragma unroll
    for (int q = 0; q < 6; ++q) {
        if (q == 3) { T0[q] = m00 * b20[q]; T1[q] = m10 * b20[q]; }
        else if (q == 4) { T0[q] = m01 * b21[q]; T1[q] = m11 * b21[q]; }
        else { T0[q] = m00 * b20[q] + m01 * b21[q]; T1[q] = m10 * b20[q] + m11 * b21[q]; }
    }
    // acc += b1^T T; b10[4] = b11[3] = 0
#pragma unroll
    for (int a = 0; a < 6; ++a)
#pragma unroll
        for (int q = 0; q < 6; ++q) {
            if (a == 3) acc[a * 6 + q] += b10[a] * T0[q];
            else if (a == 4) acc[a * 6 + q] += b11[a] * T1[q];
            else acc[a * 6 + q] += b10[a] * T0[q] + b11[a] * T1[q];
        }
}
struct PairOps { double a[4], b[4], dv[6]; };          // edge 1, edge 2 (x/z, y/z, 1/z, w), Dinv
struct DiagOps { double a[4], dv[6], db[3]; };
__device__ __forceinline__ void schur_chunk_offdiag(const int2* __restrict__ pairs, const int* __restrict__ pairPt, const double* __restrict__ PT, const double* __restrict__ E,
                                                  const double* pcache, double* out42, int4 h) {
    const int lane = threadIdx.x & 31;
    const int cnt = h.y;
    int2 rpr[IPL]; int rpt[IPL];
#pragma unroll
    for (int j = 0; j < IPL; ++j) { const int i = lane + 32 * j; if (i < cnt) { rpr[j] = pairs[h.x + i]; rpt[j] = pairPt[h.x + i]; } else { rpr[j] = make_int2(0, 0); rpt[j] = 0; } }
    const double* P1 = pcache + PC * h.z;
    const double* P2 = pcache + PC * h.w;
    const double f10 = P1[16] * P1[16], f11 = P1[17] * P1[17], f20 = P2[16] * P2[16], f21 = P2[17] * P2[17];
    double acc[36];
#pragma unroll
    for (int i = 0; i < 36; ++i) acc[i] = 0;
#pragma unroll UNR
    for (int j = 0; j < IPL; ++j) {
        const int i = lane + 32 * j;
        if (i < cnt) {
            PairOps o;
            const int e1 = rpr[j].x, e2 = rpr[j].y, ip = rpt[j];
            ld256(E + 4 * (size_t)e1, o.a[0], o.a[1], o.a[2], o.a[3]);
            ld256(E + 4 * (size_t)e2, o.b[0], o.b[1], o.b[2], o.b[3]);
            const double* pt = PT + PTR * (size_t)ip;
            ld256(pt, o.dv[0], o.dv[1], o.dv[2], o.dv[3]);
            const double2 t = *reinterpret_cast<const double2*>(pt + 4);
            o.dv[4] = t.x; o.dv[5] = t.y;
            double t1[6], t2[6], b10[6], b11[6], b20[6], b21[6];
            edge_t(P1 + 7, o.a[0], o.a[1], t1); edge_t(P2 + 7, o.b[0], o.b[1], t2);
            edge_b(o.a[0], o.a[1], o.a[2], b10, b11); edge_b(o.b[0], o.b[1], o.b[2], b20, b21);
            const double q1 = o.a[2] * o.a[3], q2 = o.b[2] * o.b[3];
            schur_item(t1, t2, o.dv, f10 * q1, f11 * q1, f20 * q2, f21 * q2, b10, b11, b20, b21, acc);
        }
    }
    int lo, n;
    warp_reduce_scatter<36>(acc, lane, lo, n);
    double* out = out42 + lo;
    if (n > 0) out[0] = acc[0];
    if (n > 1) out[1] = acc[1];
}
__device__ __forceinline__ void schur_chunk_diag(const int* __restrict__ poseEdges, const int* __restrict__ poseEdgePt, const double* __restrict__ PT, const double* __restrict__ E,
                                               const double* pcache, double* out42, int4 h) {
    const int lane = threadIdx.x & 31;
    const int cnt = h.y;
    int red[IPL], rpt[IPL];
#pragma unroll
    for (int j = 0; j < IPL; ++j) { const int i = lane + 32 * j; if (i < cnt) { red[j] = poseEdges[h.x + i]; rpt[j] = poseEdgePt[h.x + i]; } else { red[j] = 0; rpt[j] = 0; } }
    const double* P1 = pcache + PC * h.z;
    const double f10 = P1[16] * P1[16], f11 = P1[17] * P1[17];
    double acc[42];
#pragma unroll
    for (int i = 0; i < 42; ++i) acc[i] = 0;
#pragma unroll UNR
    for (int j = 0; j < IPL; ++j) {
        const int i = lane + 32 * j;
        if (i < cnt) {
            DiagOps o;
            const int e1 = red[j], ip = rpt[j];
            ld256(E + 4 * (size_t)e1, o.a[0], o.a[1], o.a[2], o.a[3]);
            const double* pt = PT + PTR * (size_t)ip;
            ld256(pt, o.dv[0], o.dv[1], o.dv[2], o.dv[3]);
            ld256(pt + 4, o.dv[4], o.dv[5], o.db[0], o.db[1]);
            o.db[2] = pt[8];
            double t1[6], b10[6], b11[6];
            edge_t(P1 + 7, o.a[0], o.a[1], t1);
            edge_b(o.a[0], o.a[1], o.a[2], b10, b11);
            const double q1 = o.a[2] * o.a[3], g0 = f10 * q1, g1 = f11 * q1;
            schur_item(t1, t1, o.dv, g0, g1, g0, g1, b10, b11, b10, b11, acc);
            const double y0 = g0 * (t1[0] * o.db[0] + t1[1] * o.db[1] + t1[2] * o.db[2]), y1 = g1 * (t1[3] * o.db[0] + t1[4] * o.db[1] + t1[5] * o.db[2]);
#pragma unroll
            for (int a = 0; a < 6; ++a) acc[36 + a] -= b10[a] * y0 + b11[a] * y1;
        }
    }
    int lo, n;
    warp_reduce_scatter<42>(acc, lane, lo, n);
    double* out = out42 + lo;
    if (n > 0) out[0] = acc[0];
    if (n > 1) out[1] = acc[1];
}
__device__ void phase_schur_partial(const Dev& D, const Ctx& c, const double* __restrict__ E) {
    // CTA barrier at entry: the previous phase (combine of the pose sums) leaves the warps of a CTA out of step, and warps that
    // enter this gather-bound phase out of step ran it 3x slower (measured); a barrier per round of chunks on top was slower again.
    __syncthreads();
    for (int ch = c.crank * NWARP + (c.tid >> 5); ch < D.nChunks; ch += c.csize * NWARP) {
        const int4 h = D.chunkHdr[ch];      // first item, items, pose A, pose B (-1: diagonal task)
        if (h.w >= 0) schur_chunk_offdiag(D.pairs, D.pairPt, D.PT, E, c.pc, D.Spart + 42 * (size_t)ch, h);
        else schur_chunk_diag(D.poseEdges, D.poseEdgePt, D.PT, E, c.pc, D.Spart + 42 * (size_t)ch, h);
    }
}
__device__ void phase_schur_combine(const Dev& D, const Ctx& c, double lambda, double* Hs, int ld) {
    const int nOff = D.nF * (D.nF - 1) / 2, nTasks = D.nF + nOff;
    for (int idx = c.wid; idx < nTasks * 42; idx += c.nw) {
        const int task = idx / 42, ent = idx - task * 42;
        if (task >= D.nF && ent >= 36) continue;
        double s = 0;
        for (int ch = D.taskChunkStart[task]; ch < D.taskChunkStart[task + 1]; ++ch) s += D.Spart[42 * (size_t)ch + ent];
        if (task < D.nF) {
            const int i1 = task;
            if (ent < 36) {
                const int r = ent / 6, q = ent % 6;
                if (q <= r) Hs[(size_t)(6 * i1 + r) * ld + 6 * i1 + q] = D.Hpp[36 * (size_t)i1 + ent] + (r == q ? lambda : 0.0) - s;
            } else D.bs[6 * i1 + (ent - 36)] = D.bp[6 * (size_t)i1 + (ent - 36)] - s;
        } else {
            int i1 = 0, rem = task - D.nF;
            while (rem >= D.nF - 1 - i1) { rem -= D.nF - 1 - i1; ++i1; }
            const int i2 = i1 + 1 + rem;
            const int r = ent / 6, q = ent % 6;
            Hs[(size_t)(6 * i2 + q) * ld + 6 * i1 + r] = -s;       // block (i2, i1) = -(sum)^T
        }
    }
}
// ---- dense LDL^T (no pivoting) + solve, one CTA; A = lower triangle, leading dimension ld (shared memory when it fits).
// Right-looking, blocked by the natural 6x6 pose blocks, with look-ahead: while warps 1.. apply the rank-6 trailing update
// of block step k, warp 0 updates the next diagonal block first and factors it (the serial chain of 6 reciprocals),
// so the only barriers are panel | update.  The rhs rides along as row n of the matrix (`brow`): after the last step it
// holds z = D^-1 L^-1 bs, and only the backward substitution L^T x = z remains.
// (LinearSolverEigen::solve: SimplicialLDLT fails only on an exactly zero pivot.)  Returns 1 on success (uniform).
// scratch: 2 x 6 x (n + 1) doubles (unscaled and scaled panel), brow: n doubles, both in shared memory. ----
__device__ __forceinline__ void ldlt_factor6(double* A, int ld, int k0, double* sL, double* sDinv, int* sFail) {   // one thread
    double a[21];            // lower triangle, row-major packed: (i, j) -> i (i + 1) / 2 + j
#pragma unroll
    for (int i = 0; i < 6; ++i)
#pragma unroll
        for (int j = 0; j <= i; ++j) a[i * (i + 1) / 2 + j] = A[(size_t)(k0 + i) * ld + k0 + j];
    bool bad = false;
#pragma unroll
    for (int k = 0; k < 6; ++k) {
        const double dk = a[k * (k + 1) / 2 + k];
        if (dk == 0.0) bad = true;
        const double inv = 1.0 / dk;
        sDinv[k] = inv;
        double l[6];
#pragma unroll
        for (int i = k + 1; i < 6; ++i) l[i] = a[i * (i + 1) / 2 + k] * inv;
#pragma unroll
        for (int i = k + 1; i < 6; ++i)
#pragma unroll
            for (int j = k + 1; j <= i; ++j) a[i * (i + 1) / 2 + j] -= l[i] * a[j * (j + 1) / 2 + k];
#pragma unroll
        for (int i = k + 1; i < 6; ++i) a[i * (i + 1) / 2 + k] = l[i];
    }
#pragma unroll
    for (int i = 0; i < 6; ++i)
#pragma unroll
        for (int j = 0; j <= i; ++j) {
            A[(size_t)(k0 + i) * ld + k0 + j] = a[i * (i + 1) / 2 + j];
            if (j < i) sL[i * 6 + j] = a[i * (i + 1) / 2 + j];
        }
    if (bad) *sFail = 1;
}
__device__ int phase_ldlt(const Dev& D, const Ctx& c, double* A, int ld, double* scratch, double* brow) {
    const int n = D.n, tid = c.tid, lane = tid & 31, warp = tid >> 5;
    const int n1 = n + 1;
    double* Xs = scratch;               // 6 x (n + 1): unscaled panel, column-major by panel column (conflict-free for lanes over rows)
    double* Ls = scratch + 6 * n1;      // 6 x (n + 1): scaled panel (= the final L entries)
    __shared__ double s_L[36], s_dinv[6];
    __shared__ int s_fail;
    for (int i = tid; i < n; i += NT) brow[i] = D.bs[i];
    if (tid == 0) { s_fail = 0; ldlt_factor6(A, ld, 0, s_L, s_dinv, &s_fail); }
    __syncthreads();
    if (s_fail) return 0;
    for (int k0 = 0; k0 < n; k0 += 6) {
        // (1) panel: row i below the block (and the rhs row, i == n): X = A21 L11^-T (unscaled), L21 = X D^-1
        for (int i = k0 + 6 + tid; i <= n; i += NT) {
            double* row = i < n ? A + (size_t)i * ld : brow;
            double xr[6];
#pragma unroll
            for (int cidx = 0; cidx < 6; ++cidx) {
                double v = row[k0 + cidx];
#pragma unroll
                for (int j = 0; j < cidx; ++j) v -= xr[j] * s_L[cidx * 6 + j];
                xr[cidx] = v;
            }
#pragma unroll
            for (int cidx = 0; cidx < 6; ++cidx) {
                const double l = xr[cidx] * s_dinv[cidx];
                Xs[cidx * n1 + i] = xr[cidx]; Ls[cidx * n1 + i] = l; row[k0 + cidx] = l;
            }
        }
        __syncthreads();
        // (2) trailing update A22[i][j] -= sum_c X[i][c] L21[j][c];  warp 0: next diagonal block, then its factorisation
        const int r0 = k0 + 6;
        if (warp == 0) {
            if (r0 < n) {
                if (lane < 21) {
                    int i = 0, t = lane;
                    while (t > i) { t -= i + 1; ++i; }      // lane -> (i, j = t), j <= i
                    double sacc = 0;
#pragma unroll
                    for (int cidx = 0; cidx < 6; ++cidx) sacc += Xs[cidx * n1 + r0 + i] * Ls[cidx * n1 + r0 + t];
                    A[(size_t)(r0 + i) * ld + r0 + t] -= sacc;
                }
                __syncwarp();
                if (lane == 0) ldlt_factor6(A, ld, r0, s_L, s_dinv, &s_fail);
            }
        } else {
            // two rows per warp and pass: the scaled panel entries L21[j][:] are loaded once for both
            for (int ia = r0 + 6 + (warp - 1); ia <= n; ia += 2 * (NWARP - 1)) {
                const int ib = ia + (NWARP - 1);
                const bool hasB = ib <= n;
                double* rowA = ia < n ? A + (size_t)ia * ld : brow;
                double* rowB = !hasB ? rowA : (ib < n ? A + (size_t)ib * ld : brow);
                double xa[6], xb[6];
#pragma unroll
                for (int cidx = 0; cidx < 6; ++cidx) { xa[cidx] = Xs[cidx * n1 + ia]; xb[cidx] = hasB ? Xs[cidx * n1 + ib] : 0.0; }
                const int jendA = ia < n ? ia : n - 1;
                const int jendB = hasB ? (ib < n ? ib : n - 1) : -1;
                const int jmax = jendA > jendB ? jendA : jendB;
                for (int j0 = r0 + lane; j0 <= jmax; j0 += 128) {
#pragma unroll
                    for (int u = 0; u < 4; ++u) {
                        const int j = j0 + 32 * u;
                        if (j <= jmax) {
                            double l[6];
#pragma unroll
                            for (int cidx = 0; cidx < 6; ++cidx) l[cidx] = Ls[cidx * n1 + j];
                            if (j <= jendA) {
                                double sacc = 0;
#pragma unroll
                                for (int cidx = 0; cidx < 6; ++cidx) sacc += xa[cidx] * l[cidx];
                                rowA[j] -= sacc;
                            }
                            if (j <= jendB) {
                                double sacc = 0;
#pragma unroll
                                for (int cidx = 0; cidx < 6; ++cidx) sacc += xb[cidx] * l[cidx];
                                rowB[j] -= sacc;
                            }
                        }
                    }
                }
            }
        }
        __syncthreads();
        if (s_fail) return 0;
    }
    // L^T x = z by warp 0 (column-oriented, warp-synchronous)
    if (warp == 0) {
        for (int k = n - 1; k >= 0; --k) {
            const double yk = brow[k];
            const double* Lk = A + (size_t)k * ld;
            for (int j = lane; j < k; j += 32) brow[j] -= Lk[j] * yk;
            __syncwarp();
        }
        for (int i = lane; i < n; i += 32) D.x[i] = brow[i];
    }
    __syncthreads();
    return 1;
}
// ---- the trial state of one LM step, part 1: every CTA applies the pose increments (SparseOptimizer::update,
//      VertexSE3Expmap::oplusImpl) to ALL poses into its own trial pose cache (so that no barrier is needed before the
//      residuals); CTA 0 also writes them to the trial buffer and returns the poses' share of computeScale ----
__device__ double phase_pose_trial(const Dev& D, const Ctx& c, const double* __restrict__ Pcur, double* __restrict__ Ptr, double* pcTr, double lambda) {
    double acc = 0;
    for (int v = c.tid; v < D.nP; v += NT) {
        double T[7];
#pragma unroll
        for (int i = 0; i < 7; ++i) T[i] = Pcur[7 * (size_t)v + i];
        const int h = D.hidx[v];
        if (h >= 0) {
            const double* u = D.x + 6 * (size_t)h;
            pose_oplus(T, u);
            if (c.crank == 0) {
#pragma unroll
                for (int i = 0; i < 6; ++i) acc += u[i] * (lambda * u[i] + D.bp[6 * (size_t)h + i]);
            }
        }
        if (c.crank == 0) {
#pragma unroll
            for (int i = 0; i < 7; ++i) Ptr[7 * (size_t)v + i] = T[i];
        }
        double* o = pcTr + PC * v;
#pragma unroll
        for (int k = 0; k < 7; ++k) o[k] = T[k];
        qtoR(T, o + 7);
        const float* cm = D.cam + 4 * (size_t)v;
        o[16] = (double)cm[0]; o[17] = (double)cm[1]; o[18] = (double)cm[2]; o[19] = (double)cm[3];
    }
    __syncthreads();
    return acc;
}
// ---- part 2, 8 lanes per point, one pass: landmark back-substitution x_l = Hll^-1 (bl - W^T x_p)
//      (block_solver.hpp:461-483; W^T x_p = w A^T (B x_p) = -sum_k t_k g_k (b_k . x_p)), the point update into the trial
//      buffer, its share of computeScale, and the residuals / robust chi2 / edge factors of the point's edges at the
//      trial state (computeActiveErrors + activeRobustChi2).  Returns (scale share, chi2 share). ----
__device__ double2 phase_points_trial(const Dev& D, const Ctx& c, const double* __restrict__ E, double* __restrict__ Eout, double* __restrict__ Xtr,
                                      const double* pcTr, double lambda) {
    const int sl = c.tid & 7;
    const unsigned gmask = 0xFFu << (c.tid & 24);
    double accS = 0, accC = 0;
    for (int p = c.wid >> 3; p < D.nL; p += c.nw >> 3) {
        const int a = D.ptStart[p], b = D.ptStart[p + 1];
        double cl[3] = {0, 0, 0};
        for (int e = a + sl; e < b; e += 8) {
            const int ic = D.ePose[e];
            const int h = D.hidx[ic];
            if (h < 0) continue;
            const double* P = c.pc + PC * ic;
            double xn, yn, iz, w, t[6], b0[6], b1[6];
            load_e4(E, e, xn, yn, iz, w);
            edge_t(P + 7, xn, yn, t);
            edge_b(xn, yn, iz, b0, b1);
            const double* xp = D.x + 6 * (size_t)h;
            double be0 = 0, be1 = 0;
#pragma unroll
            for (int k = 0; k < 6; ++k) { const double xa = xp[k]; be0 += b0[k] * xa; be1 += b1[k] * xa; }
            const double q = iz * w;
            be0 *= P[16] * P[16] * q; be1 *= P[17] * P[17] * q;
#pragma unroll
            for (int k = 0; k < 3; ++k) cl[k] += t[k] * be0 + t[3 + k] * be1;
        }
#pragma unroll
        for (int o = 4; o; o >>= 1) {
#pragma unroll
            for (int i = 0; i < 3; ++i) cl[i] += __shfl_xor_sync(gmask, cl[i], o);
        }
        const double* Di = D.PT + PTR * (size_t)p;      // 00 01 02 11 12 22
        const double* b3 = D.bl + 3 * (size_t)p;
        const double bl0 = b3[0], bl1 = b3[1], bl2 = b3[2];
        const double c0 = cl[0] + bl0, c1 = cl[1] + bl1, c2 = cl[2] + bl2;
        const double u0 = Di[0] * c0 + Di[1] * c1 + Di[2] * c2, u1 = Di[1] * c0 + Di[3] * c1 + Di[4] * c2, u2 = Di[2] * c0 + Di[4] * c1 + Di[5] * c2;
        const double* X0 = c.pts + 3 * (size_t)p;
        double X[3] = {X0[0] + u0, X0[1] + u1, X0[2] + u2};
        if (sl == 0) {
            Xtr[3 * (size_t)p] = X[0]; Xtr[3 * (size_t)p + 1] = X[1]; Xtr[3 * (size_t)p + 2] = X[2];
            accS += u0 * (lambda * u0 + bl0) + u1 * (lambda * u1 + bl1) + u2 * (lambda * u2 + bl2);
        }
        for (int e = a + sl; e < b; e += 8) accC += edge_error(D, pcTr + PC * D.ePose[e], X, e, Eout);
    }
    return make_double2(accS, accC);
}
__device__ void phase_finalize(const Dev& D, const Ctx& c) {
    for (int e = c.wid; e < D.nE; e += c.nw) {
        const double e0 = D.err[2 * (size_t)e], e1 = D.err[2 * (size_t)e + 1];
        const int o = D.eOrig[e];
        D.outChi2[o] = (double)D.invSigma2[e] * (e0 * e0 + e1 * e1);   // e->chi2() from the last computed _error (Optimizer.cc:1425)
        const double* P = c.pc + PC * D.ePose[e];
        double r[3]; qrot(P, c.pts + 3 * (size_t)D.ePt[e], r);
        D.outDepthPos[o] = r[2] + P[6] > 0.0;
    }
    // the estimate ends in whichever buffer was current last; the caller reads D.poses / D.pts
    if (c.pts != D.pts) for (int i = c.wid; i < 3 * D.nL; i += c.nw) D.pts[i] = c.pts[i];
}

// =============================================================================================
// The persistent kernel: SparseOptimizer::optimize (sparse_optimizer.cpp:354-418) around
// OptimizationAlgorithmLevenberg::solve (optimization_algorithm_levenberg.cpp:61-168), one cluster per problem.
// Every thread carries the (uniform) LM state; decisions use values every CTA reads identically after a cluster barrier.
// Dynamic shared memory: 2 pose caches (PC x maxP doubles each, maxP = most poses of any loaded problem, fixed keyframes included) |
// LDLT panel scratch (2 x 6 x (6 maxF + 1), maxF = most FREE poses) | rhs row (6 maxF) | reduced camera system (smemMatrixN^2).
// =============================================================================================
__global__ void __launch_bounds__(NT, LBA_MINB) lba_cluster_kernel(const Dev* __restrict__ probs, const volatile int* stop, int smemMatrixN, int maxP, int maxF) {
    extern __shared__ double s_dyn[];
    __shared__ double s_red[NT];
    cg::cluster_group cluster = cg::this_cluster();
    Ctx c;
    c.crank = (int)cluster.block_rank(); c.csize = (int)cluster.num_blocks(); c.tid = threadIdx.x;
    c.wid = c.crank * NT + c.tid; c.nw = c.csize * NT; c.slot = 0; c.sm = s_red;
    double* s_pc = s_dyn;                        // pose cache, linearisation state
    double* s_pcT = s_pc + PC * maxP;            // pose cache, trial state
    double* s_col = s_pcT + PC * maxP;           // 2 x 6 x (6 maxF + 1) panel scratch
    double* s_rhs = s_col + 72 * maxF + 12;      // 6 maxF: rhs row of the LDLT
    double* s_mat = s_rhs + 6 * maxF + 2;
    const Dev D = probs[blockIdx.x / c.csize];
    const bool matInSmem = D.n <= smemMatrixN && D.n > 0;
    // reduced camera system: CTA 0's shared memory; the other CTAs of the cluster reach it as distributed shared memory
    double* HsRemote = matInSmem ? cluster.map_shared_rank(s_mat, 0) : D.Hs;
    double* HsLocal = matInSmem ? s_mat : D.Hs;
    const int ld = D.n;
    const int stopIdx = blockIdx.x / c.csize;
    const bool poller = c.crank == 0 && c.tid == 0;
    // double-buffered state: the trial of an LM step is written next to the current estimate and the two swap when the
    // step is accepted, so a rejected step needs no restore ("pop") pass
    double *Elin = D.E4a, *Etrial = D.E4b;       // edge factors at the linearisation point / at the trial state
    double *Pcur = D.poses, *Ptr = D.posesB;     // poses
    double *Xcur = D.pts, *Xtr = D.ptsB;         // points
    c.pc = s_pc; c.pts = Xcur;

    for (int i = c.wid; i < D.nP; i += c.nw) {
        double* T = D.poses + 7 * (size_t)i;
#pragma unroll
        for (int k = 0; k < 7; ++k) T[k] = D.initPoses[7 * (size_t)i + k];
        qnormalize(T);                                              // SE3Quat(q, t) constructor
    }
    for (int i = c.wid; i < 3 * D.nL; i += c.nw) D.pts[i] = D.initPts[i];
    for (int i = c.wid; i < 2 * D.nE; i += c.nw) D.err[i] = 0.0;
    for (int i = c.wid; i < D.n; i += c.nw) D.x[i] = 0.0;
    int term = 0, dummy;
    cluster_sum(D, c, 0.0, (poller && stop && stop[stopIdx]) ? 1 : 0, term);   // initial terminate() poll + publishes the poses
    load_pose_cache(D, Pcur, s_pc);

    double lambda = -1, ni = 2, currentChi = 0, firstChi = 0;
    int nBad = 0, cj = 0, trials = 0;
    const int maxTrials = 10;
    const double goodUpper = 2. / 3., goodLower = 1. / 3., tau = 1e-5;
    bool ok = true;
    __shared__ unsigned long long s_tph[11];     // per-phase ns of CTA 0 (thread 0), [10] = tick
    if (c.tid < 11) s_tph[c.tid] = 0;
    __syncthreads();
#define TICK() do { if (c.wid == 0) s_tph[10] = globaltimer_ns(); } while (0)
#define TOCK(i) do { if (c.wid == 0) s_tph[i] += globaltimer_ns() - s_tph[10]; } while (0)
    for (int it = 0; it < D.iterations && !term && ok; ++it) {
        // computeActiveErrors at the start of solve(): after an accepted step the state, the errors and the edge factors are
        // exactly those of the accepted trial (same arithmetic), so only the first iteration evaluates them
        if (it == 0) {
            TICK();
            currentChi = cluster_sum(D, c, phase_errors(D, c, Elin), 0, dummy);
            TOCK(0);
        }
        double tempChi = currentChi;
        const double iniChi = currentChi;
        if (it == 0) firstChi = iniChi;
        const bool lambdaKnown = it > 0 || D.userLambdaInit > 0;
        if (it == 0 && lambdaKnown) lambda = D.userLambdaInit;
        TICK();
        phase_build_points(D, c, Elin, lambdaKnown ? lambda : -1.0);
        TOCK(1); TICK();
        phase_build_poses_partial(D, c, Elin);
        csync();
        phase_build_poses_combine(D, c);
        TOCK(2);
        if (it == 0) {
            if (!lambdaKnown) { csync(); lambda = tau * phase_maxdiag(D, c); }
            ni = 2; nBad = 0;
        }
        bool needPrep = !lambdaKnown;
        double rho = 0;
        int qmax = 0;
        do {
            if (needPrep) {
                TICK();
                phase_point_prep(D, c, lambda);
                csync();
                TOCK(3);
            }
            int ok2 = 1;
            if (D.nF) {
                TICK();
                phase_schur_partial(D, c, Elin);
                csync();
                TOCK(4); TICK();
                phase_schur_combine(D, c, lambda, HsRemote, ld);   // every CTA writes entries into CTA 0's shared memory through DSMEM
                csync();
                TOCK(6); TICK();
                if (c.crank == 0) {
                    ok2 = phase_ldlt(D, c, HsLocal, ld, s_col, s_rhs);
                    if (c.tid == 0) D.partial[4 * PSLOT] = (double)ok2;
                }
                csync();
                ok2 = (int)D.partial[4 * PSLOT];
                TOCK(5);
            }
            TICK();
            const double sP = phase_pose_trial(D, c, Pcur, Ptr, s_pcT, lambda);
            TOCK(7); TICK();
            const double2 sc = phase_points_trial(D, c, Elin, Etrial, Xtr, s_pcT, lambda);
            double scale0;
            cluster_sum2(D, c, sP + sc.x, sc.y, (poller && stop && stop[stopIdx]) ? 1 : 0, term, scale0, tempChi);
            TOCK(8);
            if (!ok2) tempChi = DBL_MAX;
            rho = currentChi - tempChi;
            const double scale = scale0 + 1e-3;
            rho /= scale;
            if (rho > 0 && isfinite(tempChi)) {
                double alpha = 1. - pow((2 * rho - 1), 3.0);
                alpha = fmin(alpha, goodUpper);
                const double scaleFactor = fmax(goodLower, alpha);
                lambda *= scaleFactor;
                ni = 2;
                currentChi = tempChi;
                double* sw;
                sw = Elin; Elin = Etrial; Etrial = sw;
                sw = Pcur; Pcur = Ptr; Ptr = sw;
                sw = Xcur; Xcur = Xtr; Xtr = sw;
                c.pts = Xcur;
                for (int i = c.tid; i < PC * D.nP; i += NT) s_pc[i] = s_pcT[i];
                __syncthreads();
            } else {
                lambda *= ni;                                   // the trial buffers are simply abandoned (pop)
                ni *= 2;
            }
            needPrep = true;
            ++qmax; ++trials;
        } while (rho < 0 && qmax < maxTrials && !term);
        ++cj;
        if (qmax == maxTrials || rho == 0) ok = false;
        else {
            if ((iniChi - currentChi) * 1e3 < iniChi) ++nBad; else nBad = 0;
            if (nBad >= 3) ok = false;
        }
    }
    phase_finalize(D, c);
    if (Pcur != D.poses) for (int i = c.wid; i < 7 * D.nP; i += c.nw) D.poses[i] = Pcur[i];
    if (c.wid == 0) for (int i = 0; i < 10; ++i) D.stats[8 + i] = (double)s_tph[i];
    if (c.wid == 0) { D.stats[0] = cj; D.stats[1] = trials; D.stats[2] = lambda; D.stats[3] = currentChi; D.stats[4] = firstChi; }
}

// =============================================================================================
// BlockSolver::buildStructure (block_solver.hpp:143-295) on the device: one CTA per problem turns the caller's edge list
// into the index structures of the LM kernel -- edges sorted by point (stable), the (free pose, point) -> edge table, the
// per-pose edge lists, the (e1, e2) pair list of every off-diagonal Schur block, and the chunked work list.  Only the raw
// graph (28 B per edge) crosses PCIe; the derived lists (~2.3 MB per 40k-edge problem) never exist on the host.
// Every list is in a deterministic order (by point, ties by the caller's edge index).
// status: 0 ok, 1 edge index out of range, 2 duplicate (point, keyframe) observation, 3 pair list larger than sized.
// =============================================================================================
constexpr int BT = 1024;
__device__ __forceinline__ int bt_excl_scan(int v, int* sm, int& total) {   // exclusive scan of one int per thread over the CTA
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    int inc = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { const int t = __shfl_up_sync(0xffffffffu, inc, o); if (lane >= o) inc += t; }
    if (lane == 31) sm[w] = inc;
    __syncthreads();
    if (w == 0) {
        int x = sm[lane];
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { const int t = __shfl_up_sync(0xffffffffu, x, o); if (lane >= o) x += t; }
        sm[lane] = x;
    }
    __syncthreads();
    const int base = w ? sm[w - 1] : 0;
    total = sm[31];
    __syncthreads();
    return base + inc - v;
}
__global__ void __launch_bounds__(BT) lba_build_structure_kernel(Dev* probs, int* status) {
    Dev& D = probs[blockIdx.x];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int nL = D.nL, nE = D.nE, nF = D.nF, nP = D.nP;
    const int nOff = nF * (nF - 1) / 2, nTasks = nF + nOff;
    __shared__ int s_sm[32];
    __shared__ int s_err;
    int* ptStart = const_cast<int*>(D.ptStart);
    int* ePt = const_cast<int*>(D.ePt); int* ePose = const_cast<int*>(D.ePose); int* eOrig = const_cast<int*>(D.eOrig);
    double* obs = const_cast<double*>(D.obs); float* is2 = const_cast<float*>(D.invSigma2);
    int* poseEdges = const_cast<int*>(D.poseEdges); int* poseEdgePt = const_cast<int*>(D.poseEdgePt);
    int2* pairs = const_cast<int2*>(D.pairs); int* pairPt = const_cast<int*>(D.pairPt);
    int4* chunkHdr = const_cast<int4*>(D.chunkHdr); int* taskChunkStart = const_cast<int*>(D.taskChunkStart);
    if (tid == 0) s_err = 0;
    for (int i = tid; i <= nL; i += BT) D.bCnt[i] = 0;
    for (size_t i = tid; i < (size_t)nF * nL; i += BT) D.bEid[i] = -1;
    __syncthreads();
    // 1. edges per point
    for (int e = tid; e < nE; e += BT) {
        const int p = D.rawEdgePoint[e], c = D.rawEdgePose[e];
        if (p < 0 || p >= nL || c < 0 || c >= nP) s_err = 1;
        else atomicAdd(&D.bCnt[p], 1);
    }
    __syncthreads();
    if (s_err) { if (tid == 0) status[blockIdx.x] = s_err; return; }
    // 2. ptStart = exclusive scan
    {
        int carry = 0;
        for (int b0 = 0; b0 < nL; b0 += BT) {
            const int i = b0 + tid;
            const int v = i < nL ? D.bCnt[i] : 0;
            int total;
            const int ex = bt_excl_scan(v, s_sm, total);
            if (i < nL) { ptStart[i] = carry + ex; D.bCursor[i] = carry + ex; }
            carry += total;
        }
        if (tid == 0) ptStart[nL] = carry;
    }
    __syncthreads();
    // 3. group by point (arbitrary order), then order every point's slice by the caller's edge index (stable sort by point)
    for (int e = tid; e < nE; e += BT) D.bTmp[atomicAdd(&D.bCursor[D.rawEdgePoint[e]], 1)] = e;
    __syncthreads();
    for (int p = tid; p < nL; p += BT) {
        const int a = ptStart[p], b = ptStart[p + 1];
        for (int i = a + 1; i < b; ++i) {
            const int v = D.bTmp[i];
            int j = i - 1;
            while (j >= a && D.bTmp[j] > v) { D.bTmp[j + 1] = D.bTmp[j]; --j; }
            D.bTmp[j + 1] = v;
        }
    }
    __syncthreads();
    for (int p = tid; p < nL; p += BT)
        for (int k = ptStart[p]; k < ptStart[p + 1]; ++k) {
            const int e = D.bTmp[k], c = D.rawEdgePose[e];
            ePt[k] = p; ePose[k] = c; eOrig[k] = e;
            obs[2 * (size_t)k] = D.rawObs[2 * (size_t)e]; obs[2 * (size_t)k + 1] = D.rawObs[2 * (size_t)e + 1];
            is2[k] = D.rawInvSigma2[e];
            const int h = D.hidx[c];
            if (h >= 0 && atomicCAS(&D.bEid[(size_t)h * nL + p], -1, k) != -1) s_err = 2;
        }
    __syncthreads();
    if (s_err) { if (tid == 0) status[blockIdx.x] = s_err; return; }
    // 4. items per task: edges of a free pose / points shared by a pair of free poses; warp per task, ballot over points
    auto taskRows = [&](int t, const int*& r1, const int*& r2) {
        if (t < nF) { r1 = D.bEid + (size_t)t * nL; r2 = r1; return; }
        int i1 = 0, rem = t - nF;
        while (rem >= nF - 1 - i1) { rem -= nF - 1 - i1; ++i1; }
        r1 = D.bEid + (size_t)i1 * nL; r2 = D.bEid + (size_t)(i1 + 1 + rem) * nL;
    };
    for (int t = warp; t < nTasks; t += BT / 32) {
        const int *r1, *r2;
        taskRows(t, r1, r2);
        int cnt = 0;
#pragma unroll 4
        for (int p0 = 0; p0 < nL; p0 += 32) {
            const int p = p0 + lane;
            const bool f = p < nL && r1[p] >= 0 && r2[p] >= 0;
            cnt += __popc(__ballot_sync(0xffffffffu, f));
        }
        if (lane == 0) D.bTaskLen[t] = cnt;
    }
    __syncthreads();
    // 5. where every task's items and chunks start (pose lists and pair lists are separate arrays)
    int nChunks = 0;
    {
        int carry = 0, tot;
        for (int b0 = 0; b0 < nF; b0 += BT) {                  // pose lists
            const int t = b0 + tid;
            const int ex = bt_excl_scan(t < nF ? D.bTaskLen[t] : 0, s_sm, tot);
            if (t < nF) D.bItemStart[t] = carry + ex;
            carry += tot;
        }
        carry = 0;
        for (int b0 = nF; b0 < nTasks; b0 += BT) {             // pair lists
            const int t = b0 + tid;
            const int ex = bt_excl_scan(t < nTasks ? D.bTaskLen[t] : 0, s_sm, tot);
            if (t < nTasks) D.bItemStart[t] = carry + ex;
            carry += tot;
        }
        if (carry > D.pairCap) s_err = 3;
        carry = 0;
        for (int b0 = 0; b0 < nTasks; b0 += BT) {              // chunks
            const int t = b0 + tid;
            const int ex = bt_excl_scan(t < nTasks ? (D.bTaskLen[t] + SCH - 1) / SCH : 0, s_sm, tot);
            if (t < nTasks) taskChunkStart[t] = carry + ex;
            carry += tot;
        }
        nChunks = carry;
        if (tid == 0) taskChunkStart[nTasks] = nChunks;
        if (nChunks > D.chunkCap) s_err = 3;
    }
    __syncthreads();
    if (s_err) { if (tid == 0) status[blockIdx.x] = s_err; return; }
    // 6. the lists themselves (ascending point = ascending internal edge id) and the chunk headers
    for (int t = warp; t < nTasks; t += BT / 32) {
        const int *r1, *r2;
        taskRows(t, r1, r2);
        int pos = D.bItemStart[t];
        for (int p0 = 0; p0 < nL; p0 += 32) {
            const int p = p0 + lane;
            const int e1 = p < nL ? r1[p] : -1, e2 = p < nL ? r2[p] : -1;
            const bool f = e1 >= 0 && e2 >= 0;
            const unsigned m = __ballot_sync(0xffffffffu, f);
            if (f) {
                const int k = pos + __popc(m & ((1u << lane) - 1));
                if (t < nF) { poseEdges[k] = e1; poseEdgePt[k] = p; }
                else { pairs[k] = make_int2(e1, e2); pairPt[k] = p; }
            }
            pos += __popc(m);
        }
    }
    for (int t = tid; t < nTasks; t += BT) {
        int poseA, poseB = -1;
        if (t < nF) poseA = D.freePose[t];
        else {
            int i1 = 0, rem = t - nF;
            while (rem >= nF - 1 - i1) { rem -= nF - 1 - i1; ++i1; }
            poseA = D.freePose[i1]; poseB = D.freePose[i1 + 1 + rem];
        }
        const int len = D.bTaskLen[t], first = D.bItemStart[t];
        int ch = taskChunkStart[t];
        for (int f = 0; f < len; f += SCH) chunkHdr[ch++] = make_int4(first + f, min(SCH, len - f), poseA, poseB);
    }
    if (tid == 0) { D.nChunks = nChunks; status[blockIdx.x] = 0; }
}

// =============================================================================================
// host side
// =============================================================================================
struct Packed {             // one uploaded problem: where its pieces live inside the arena
    int nP, nL, nE, nF;
    size_t posesOff, ptsOff, chi2Off, dposOff, statsOff, initPosesOff, initPtsOff;
};

struct Solver {
    int device, maxP, maxL, maxE, maxBatch;
    cudaStream_t st = nullptr;
    uint8_t* d_arena = nullptr; uint8_t* h_arena = nullptr; size_t arenaBytes = 0, perProblem = 0;
    Dev* d_probs = nullptr; std::vector<Dev> h_probs;
    std::vector<Packed> packed;
    int* h_stop = nullptr; int* d_stop = nullptr;   // mapped pinned stop flags, one per problem
    int* h_status = nullptr; int* d_status = nullptr;   // mapped pinned status words of the structure kernel
    cudaEvent_t evDone = nullptr, evRun = nullptr;   // evRun: end of the last run, on whatever stream the caller launched it
    cudaEvent_t evWait = nullptr;                    // blocking-sync event: upload / download sleep on it instead of spinning (device_utils.cuh)
    bool ranOnce = false;
    int nLoaded = 0, launches = 0, numSMs = 148;
    ~Solver() {
        cudaSetDevice(device);
        if (d_arena) cudaFree(d_arena);
        if (h_arena) cudaFreeHost(h_arena);
        if (d_probs) cudaFree(d_probs);
        if (h_stop) cudaFreeHost(h_stop);
        if (h_status) cudaFreeHost(h_status);
        if (evDone) cudaEventDestroy(evDone);
        if (evRun) cudaEventDestroy(evRun);
        if (evWait) cudaEventDestroy(evWait);
        if (st) cudaStreamDestroy(st);
    }
    static size_t al(size_t b) { return (b + 255) & ~(size_t)255; }
    static size_t outBlock(size_t nP, size_t nL, size_t nE) { return al(56 * nP) + al(24 * nL) + al(8 * nE) + al(nE) + al(512); }
    static size_t pairCapOf(size_t nP, size_t nE) { return nE * (nP > 1 ? nP - 1 : 1) / 2 + 1; }
    static size_t chunkCapOf(size_t nP, size_t nE) { return nP + nP * nP / 2 + 2 + (nE + pairCapOf(nP, nE)) / SCH + 1; }
    static size_t need(size_t nP, size_t nL, size_t nE) {
        const size_t n = 6 * nP, pc = pairCapOf(nP, nE), cc = chunkCapOf(nP, nE), tasks = nP + nP * nP / 2 + 2;
        size_t b = outBlock(nP, nL, nE);                                          // poses, pts, chi2, depth flags, stats (downloaded in one copy)
        b += al(56 * nP) + al(24 * nL) + al(16 * nP) + 2 * al(4 * nP);            // uploaded: initial estimates, cam, hidx, freePose
        b += 2 * al(4 * nE) + al(16 * nE) + al(4 * nE);                           // uploaded: the caller's edges (point, pose, obs, invSigma2)
        b += 3 * al(4 * nE) + al(16 * nE) + al(4 * nE) + al(4 * (nL + 1));        // edges sorted by point: ePt, ePose, eOrig, obs, invSigma2; ptStart
        b += 2 * al(4 * nE) + al(8 * pc) + al(4 * pc);                            // poseEdges, poseEdgePt, pairs, pairPt
        b += al(16 * cc) + al(4 * (tasks + 1));                                   // chunk headers, task chunk starts
        b += 2 * al(4 * (nL + 1)) + al(4 * nE) + al(4 * nP * nL) + 2 * al(4 * tasks);   // structure-build scratch
        b += al(56 * nP) + al(24 * nL);                                           // second state buffers
        b += al(16 * nE) + 2 * al(32 * nE);                                       // err, edge factors x 2
        b += al(288 * nP) + al(48 * nP) + al(48 * nL) + al(96 * nL) + al(24 * nL);   // Hpp, bp, Hll, PT, bl
        b += al(8 * 42 * cc) + al(8 * 28 * (nP + nE / SCH + 1));                  // Spart, Ppart
        b += al(8 * n * n) + al(8 * n) + al(8 * (n + 3 * nL)) + al(8 * (4 * PSLOT + 8));   // Hs, bs, x, partial
        return b;
    }
    int init() {
        CK(cudaSetDevice(device));
        cudaDeviceProp prop;
        CK(cudaGetDeviceProperties(&prop, device));
        if (prop.major < 10) { set_error("device is not sm_100+ (Blackwell); this library has no other code path"); return ORB_ERR_CUDA; }
        numSMs = prop.multiProcessorCount;
        perProblem = need(maxP, maxL, maxE);
        arenaBytes = perProblem * (size_t)maxBatch;
        CK(cudaMalloc(&d_arena, arenaBytes));
        {
            orbx::ScopedGpuAffinity numaLocal(device);     // pinned pages on the GPU's NUMA node (host_affinity.h)
            CK(cudaMallocHost(&h_arena, arenaBytes));
        }
        CK(cudaMalloc(&d_probs, sizeof(Dev) * maxBatch));
        CK(cudaHostAlloc(&h_stop, sizeof(int) * maxBatch, cudaHostAllocMapped));
        CK(cudaHostGetDevicePointer(&d_stop, h_stop, 0));
        CK(cudaHostAlloc(&h_status, sizeof(int) * maxBatch, cudaHostAllocMapped));
        CK(cudaHostGetDevicePointer(&d_status, h_status, 0));
        CK(cudaEventCreateWithFlags(&evDone, cudaEventDisableTiming));
        CK(cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking));
        CK(cudaEventCreateWithFlags(&evRun, cudaEventDisableTiming));
        CK(orbx::make_blocking_event(&evWait));
        h_probs.resize(maxBatch); packed.resize(maxBatch);
        return ORB_OK;
    }

    // Lays one problem out in slot `slot` of the arena and copies the caller's graph into the pinned mirror; the index
    // structures are derived on the device (lba_build_structure_kernel).
    int pack(int slot, const LbaProblem* P) {
        const int nP = P->nPoses, nL = P->nPoints, nE = P->nEdges;
        if (nP < 1 || nL < 0 || nE < 0 || nP > maxP || nL > maxL || nE > maxE || !P->poses || !P->poseFixed || !P->cam ||
            (nL && !P->points) || (nE && (!P->edgePoint || !P->edgePose || !P->obs || !P->invSigma2))) {
            set_error("lba: problem larger than the handle or null arrays"); return ORB_ERR_ARG;
        }
        std::vector<int> hidx(nP, -1), freePose;
        for (int i = 0; i < nP; ++i) if (!P->poseFixed[i]) { hidx[i] = (int)freePose.size(); freePose.push_back(i); }
        const int nF = (int)freePose.size(), n = 6 * nF;
        if (nF + nL == 0) { set_error("lba: 0 vertices to optimize"); return ORB_ERR_ARG; }
        const int nOff = nF * (nF - 1) / 2, nTasks = nF + nOff;
        const size_t pc = pairCapOf(nP, nE), cc = chunkCapOf(nP, nE);
        const size_t base = perProblem * (size_t)slot;
        size_t off = base;
        auto carve = [&](size_t bytes) { size_t o = off; off += al(bytes); return o; };
        auto put = [&](size_t o, const void* src, size_t bytes) { if (bytes) memcpy(h_arena + o, src, bytes); };
        Dev& D = h_probs[slot]; memset(&D, 0, sizeof(D));
        Packed& K = packed[slot];
        K.nP = nP; K.nL = nL; K.nE = nE; K.nF = nF;
        D.nP = nP; D.nL = nL; D.nE = nE; D.nF = nF; D.n = n; D.pairCap = (int)pc; D.chunkCap = (int)cc;
        // --- output block: fixed layout (sized by the handle's maxima) so that one strided copy downloads every problem ---
        K.posesOff = carve(56 * (size_t)maxP); D.poses = (double*)(d_arena + K.posesOff);
        K.ptsOff = carve(24 * (size_t)maxL); D.pts = (double*)(d_arena + K.ptsOff);
        K.chi2Off = carve(8 * (size_t)maxE); D.outChi2 = (double*)(d_arena + K.chi2Off);
        K.dposOff = carve((size_t)maxE); D.outDepthPos = (uint8_t*)(d_arena + K.dposOff);
        K.statsOff = carve(512); D.stats = (double*)(d_arena + K.statsOff);
        // --- uploaded block: the caller's graph ---
        const size_t upBase = off;
        size_t o;
        o = carve(56 * (size_t)nP); put(o, P->poses, 56 * (size_t)nP); D.initPoses = (const double*)(d_arena + o);
        o = carve(24 * (size_t)nL); put(o, P->points, 24 * (size_t)nL); D.initPts = (const double*)(d_arena + o);
        o = carve(16 * (size_t)nP); put(o, P->cam, 16 * (size_t)nP); D.cam = (const float*)(d_arena + o);
        o = carve(4 * (size_t)nP); put(o, hidx.data(), 4 * (size_t)nP); D.hidx = (const int*)(d_arena + o);
        o = carve(4 * (size_t)std::max(nF, 1)); put(o, freePose.data(), 4 * (size_t)nF); D.freePose = (const int*)(d_arena + o);
        o = carve(4 * (size_t)nE); put(o, P->edgePoint, 4 * (size_t)nE); D.rawEdgePoint = (const int*)(d_arena + o);
        o = carve(4 * (size_t)nE); put(o, P->edgePose, 4 * (size_t)nE); D.rawEdgePose = (const int*)(d_arena + o);
        o = carve(16 * (size_t)nE); put(o, P->obs, 16 * (size_t)nE); D.rawObs = (const double*)(d_arena + o);
        o = carve(4 * (size_t)nE); put(o, P->invSigma2, 4 * (size_t)nE); D.rawInvSigma2 = (const float*)(d_arena + o);
        uploadOff[slot] = upBase; uploadBytes[slot] = off - upBase;
        // --- derived on the device ---
        D.ePt = (const int*)(d_arena + carve(4 * (size_t)nE)); D.ePose = (const int*)(d_arena + carve(4 * (size_t)nE)); D.eOrig = (const int*)(d_arena + carve(4 * (size_t)nE));
        D.obs = (const double*)(d_arena + carve(16 * (size_t)nE)); D.invSigma2 = (const float*)(d_arena + carve(4 * (size_t)nE));
        D.ptStart = (const int*)(d_arena + carve(4 * (size_t)(nL + 1)));
        D.poseEdges = (const int*)(d_arena + carve(4 * (size_t)nE)); D.poseEdgePt = (const int*)(d_arena + carve(4 * (size_t)nE));
        D.pairs = (const int2*)(d_arena + carve(8 * pc)); D.pairPt = (const int*)(d_arena + carve(4 * pc));
        D.chunkHdr = (const int4*)(d_arena + carve(16 * cc)); D.taskChunkStart = (const int*)(d_arena + carve(4 * (size_t)(nTasks + 1)));
        D.bCnt = (int*)(d_arena + carve(4 * (size_t)(nL + 1))); D.bCursor = (int*)(d_arena + carve(4 * (size_t)(nL + 1)));
        D.bTmp = (int*)(d_arena + carve(4 * (size_t)nE)); D.bEid = (int*)(d_arena + carve(4 * (size_t)std::max(nF, 1) * (size_t)std::max(nL, 1)));
        D.bTaskLen = (int*)(d_arena + carve(4 * (size_t)std::max(nTasks, 1))); D.bItemStart = (int*)(d_arena + carve(4 * (size_t)std::max(nTasks, 1)));
        // --- state / scratch of the LM kernel ---
        D.posesB = (double*)(d_arena + carve(56 * (size_t)nP));
        D.ptsB = (double*)(d_arena + carve(24 * (size_t)nL));
        D.err = (double*)(d_arena + carve(16 * (size_t)nE));
        D.E4a = (double*)(d_arena + carve(32 * (size_t)nE)); D.E4b = (double*)(d_arena + carve(32 * (size_t)nE));
        D.Hpp = (double*)(d_arena + carve(288 * (size_t)std::max(nF, 1))); D.bp = (double*)(d_arena + carve(48 * (size_t)std::max(nF, 1)));
        D.Hll = (double*)(d_arena + carve(48 * (size_t)nL)); D.bl = (double*)(d_arena + carve(24 * (size_t)nL));
        D.PT = (double*)(d_arena + carve(96 * (size_t)nL));
        D.Spart = (double*)(d_arena + carve(8 * 42 * cc));
        D.Ppart = (double*)(d_arena + carve(8 * 28 * ((size_t)nP + (size_t)nE / SCH + 1)));
        D.Hs = (double*)(d_arena + carve(8 * (size_t)n * n)); D.bs = (double*)(d_arena + carve(8 * (size_t)std::max(n, 1)));
        D.x = (double*)(d_arena + carve(8 * ((size_t)n + 3 * (size_t)nL)));
        D.partial = (double*)(d_arena + carve(8 * (4 * PSLOT + 8)));
        if (off - base > perProblem) { set_error("lba: arena too small (internal sizing error)"); return ORB_ERR_CAPACITY; }
        D.delta = P->huberDelta; D.dsqr = P->huberDelta * P->huberDelta; D.userLambdaInit = P->userLambdaInit; D.iterations = P->iterations;
        return ORB_OK;
    }
    std::vector<size_t> uploadBytes = std::vector<size_t>(1), uploadOff = std::vector<size_t>(1);

    int upload(int count, const LbaProblem* probs) {
        if (count < 1 || count > maxBatch) { set_error("lba: batch larger than max_batch"); return ORB_ERR_ARG; }
        uploadBytes.assign(maxBatch, 0); uploadOff.assign(maxBatch, 0);
        CK(cudaSetDevice(device));
        if (ranOnce) CK(cudaStreamWaitEvent(st, evRun, 0));   // a run may still be reading the arena on the caller's stream
        {   // stage the callers' graphs in the pinned mirror on a few host threads; every worker queues the host-to-device
            // copy of a problem as soon as it is staged, so the copies overlap the staging of the next problems
            const int nth = std::max(1, std::min<int>({count, (int)std::thread::hardware_concurrency(), 8}));
            std::atomic<int> next(0), firstErr(ORB_OK);
            auto worker = [&]() {
                cudaSetDevice(device);
                for (int i = next.fetch_add(1); i < count; i = next.fetch_add(1)) {
                    int rc = pack(i, probs + i);
                    if (!rc && cudaMemcpyAsync(d_arena + uploadOff[i], h_arena + uploadOff[i], uploadBytes[i], cudaMemcpyHostToDevice, st) != cudaSuccess) rc = ORB_ERR_CUDA;
                    if (rc) { int exp = ORB_OK; firstErr.compare_exchange_strong(exp, rc); }
                }
            };
            std::vector<std::thread> pool;
            for (int t = 1; t < nth; ++t) pool.emplace_back(worker);
            worker();
            for (auto& t : pool) t.join();
            if (firstErr.load()) {
                cudaStreamSynchronize(st);
                // the workers' own messages live in their thread-local error strings: restate it for the calling thread
                set_error(firstErr.load() == ORB_ERR_CUDA ? "lba: host-to-device copy failed"
                                                          : "lba: a problem is larger than the handle, has null arrays or no vertex to optimize");
                return firstErr.load();
            }
        }
        CK(cudaSetDevice(device));
        CK(cudaMemcpyAsync(d_probs, h_probs.data(), sizeof(Dev) * count, cudaMemcpyHostToDevice, st));
        for (int i = 0; i < count; ++i) h_status[i] = -1;
        lba_build_structure_kernel<<<count, BT, 0, st>>>(d_probs, d_status);     // BlockSolver::buildStructure, one CTA per problem
        CK(cudaGetLastError());
        CK(orbx::wait_stream_blocking(st, evWait));   // the resident copy must be complete before a run on any other stream
        for (int i = 0; i < count; ++i) {
            if (h_status[i] == 0) continue;
            nLoaded = 0;
            if (h_status[i] == 1) { set_error("lba: edge index out of range"); return ORB_ERR_ARG; }
            if (h_status[i] == 2) { set_error("lba: duplicate (point, keyframe) observation"); return ORB_ERR_ARG; }
            set_error("lba: pair / chunk list larger than sized"); return ORB_ERR_CAPACITY;
        }
        nLoaded = count;
        return ORB_OK;
    }
    // (re)start from the uploaded initial estimates and run the whole LM loop for every loaded problem
    int run(cudaStream_t s) {
        if (nLoaded < 1) { set_error("lba: nothing uploaded"); return ORB_ERR_ARG; }
        CK(cudaSetDevice(device));
        int maxN = 0, mp = 1, mf = 1;
        for (int i = 0; i < nLoaded; ++i) {                             // the kernel itself restarts from the uploaded estimates
            maxN = std::max(maxN, 6 * packed[i].nF); mp = std::max(mp, packed[i].nP); mf = std::max(mf, packed[i].nF);
        }
        // dynamic shared memory: the pose caches hold every pose of the window (fixed keyframes included, the reference puts no
        // bound on lFixedCameras), the LDLT panels and rhs only the free ones; the reduced camera system joins them when it fits
        const size_t fixedSmem = 8 * ((size_t)2 * PC * mp + (size_t)78 * mf + 16);
        const size_t maxDyn = 200 * 1024;                               // of the 227 KB opt-in limit; the kernel has ~5 KB static
        if (fixedSmem > maxDyn) {
            set_error("lba: window too large for the shared-memory pose caches (about 600 keyframes incl. fixed ones)");
            return ORB_ERR_CAPACITY;
        }
        const int smemN = (int)floor(sqrt((double)(maxDyn - fixedSmem) / 8.0));
        const int matN = maxN <= smemN ? maxN : 0;   // matrix in shared memory only when every problem of the batch fits
        const size_t dyn = fixedSmem + 8 * (size_t)matN * matN;
        if (orbx::ensure_dynamic_smem(lba_cluster_kernel, dyn, device)) return ORB_ERR_CUDA;
        cudaLaunchConfig_t cfg; memset(&cfg, 0, sizeof(cfg));
        cfg.blockDim = dim3(NT);
        cfg.dynamicSmemBytes = dyn;
        cfg.stream = s;
        cudaLaunchAttribute attr[1];
        attr[0].id = cudaLaunchAttributeClusterDimension;
        attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
        cfg.attrs = attr; cfg.numAttrs = 1;
        // cluster size: the largest (any size up to 8, not only powers of two) for which every problem's cluster is resident at
        // once -- asked from the occupancy calculator, because clusters cannot straddle GPCs (24 clusters of 6 fit 148 SMs, 25 do not)
        int csize = 1;
        for (int cand = MAXC; cand >= 1; --cand) {
            if ((long)nLoaded * cand > numSMs) continue;
            attr[0].val.clusterDim.x = cand;
            cfg.gridDim = dim3(nLoaded * cand);
            int nClusters = 0;
            if (cudaOccupancyMaxActiveClusters(&nClusters, lba_cluster_kernel, &cfg) == cudaSuccess && nClusters >= nLoaded) { csize = cand; break; }
        }
        (void)cudaGetLastError();
        if (forcedCluster > 0) csize = forcedCluster;
        attr[0].val.clusterDim.x = csize;
        cfg.gridDim = dim3(nLoaded * csize);
        const Dev* dp = d_probs; const volatile int* ds = d_stop; int mn = matN;
        CK(cudaLaunchKernelEx(&cfg, lba_cluster_kernel, dp, ds, mn, mp, mf));
        CK(cudaEventRecord(evRun, s));
        ranOnce = true;
        launches = 1;
        lastCluster = csize;
        return ORB_OK;
    }
    int lastCluster = 1, forcedCluster = 0;
    int download(int count, LbaResult* res, cudaStream_t s) {
        for (int i = 0; i < count; ++i)
            if (!res[i].poses || !res[i].points || !res[i].edgeChi2 || !res[i].edgeDepthPositive) { set_error("lba: null result arrays"); return ORB_ERR_ARG; }
        // the output blocks sit at the start of every slot with one layout: a single strided copy into the pinned mirror
        const size_t ob = outBlock(maxP, maxL, maxE);
        if (ranOnce) CK(cudaStreamWaitEvent(s, evRun, 0));    // the run may be on another stream than this copy
        CK(cudaMemcpy2DAsync(h_arena, perProblem, d_arena, perProblem, ob, (size_t)count, cudaMemcpyDeviceToHost, s));
        CK(orbx::wait_stream_blocking(s, evWait));
        for (int i = 0; i < count; ++i) {
            const Packed& K = packed[i];
            LbaResult& R = res[i];
            memcpy(R.poses, h_arena + K.posesOff, 56 * (size_t)K.nP);
            if (K.nL) memcpy(R.points, h_arena + K.ptsOff, 24 * (size_t)K.nL);
            if (K.nE) { memcpy(R.edgeChi2, h_arena + K.chi2Off, 8 * (size_t)K.nE); memcpy(R.edgeDepthPositive, h_arena + K.dposOff, (size_t)K.nE); }
            const double* stt = (const double*)(h_arena + K.statsOff);
            R.iterations = (int)stt[0]; R.trials = (int)stt[1]; R.lambda = stt[2]; R.chi2 = stt[3]; R.initialChi2 = stt[4];
            R.gpuLaunches = 1;
        }
        return ORB_OK;
    }
};

}  // namespace lba

using namespace lba;

struct lba_handle { Solver s; };

extern "C" {

int lba_create_batch(lba_handle** out, int max_poses, int max_points, int max_edges, int max_batch, int device) {
    if (!out || max_poses < 1 || max_points < 1 || max_edges < 1 || max_batch < 1) { set_error("lba_create: bad argument"); return ORB_ERR_ARG; }
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) { set_error("no CUDA device (this library has no CPU path)"); return ORB_ERR_CUDA; }
    if (device < 0 || device >= ndev) { set_error("lba_create: bad device index"); return ORB_ERR_ARG; }
    lba_handle* h = new lba_handle();
    h->s.device = device; h->s.maxP = max_poses; h->s.maxL = max_points; h->s.maxE = max_edges; h->s.maxBatch = max_batch;
    int rc = h->s.init();
    if (rc) { delete h; return rc; }
    *out = h;
    return ORB_OK;
}
int lba_create(lba_handle** out, int max_poses, int max_points, int max_edges, int device) {
    return lba_create_batch(out, max_poses, max_points, max_edges, 1, device);
}
void lba_destroy(lba_handle* h) { delete h; }

int lba_upload_batch(lba_handle* h, int count, const LbaProblem* problems) {
    if (!h || !problems) { set_error("lba_upload_batch: bad argument"); return ORB_ERR_ARG; }
    return h->s.upload(count, problems);
}
int lba_run_batch_device(lba_handle* h, void* stream) {
    if (!h) { set_error("lba_run_batch_device: bad argument"); return ORB_ERR_ARG; }
    for (int i = 0; i < h->s.nLoaded; ++i) h->s.h_stop[i] = 0;
    return h->s.run((cudaStream_t)stream);
}
int lba_download_batch(lba_handle* h, int count, LbaResult* results) {
    if (!h || !results || count < 1 || count > h->s.nLoaded) { set_error("lba_download_batch: bad argument"); return ORB_ERR_ARG; }
    return h->s.download(count, results, h->s.st);
}
int lba_last_cluster_size(const lba_handle* h) { return h ? h->s.lastCluster : ORB_ERR_ARG; }
int lba_set_cluster_size(lba_handle* h, int ctas) {
    if (!h || ctas < 0 || ctas > MAXC) { set_error("lba_set_cluster_size: 0 (auto) or 1..8"); return ORB_ERR_ARG; }
    h->s.forcedCluster = ctas;
    return ORB_OK;
}
/* ns spent by CTA 0 of problem `i` in each phase of the last downloaded run: errors, build_points, build_poses, point_prep, schur, ldlt, backsub, update, errors(trial) */
int lba_get_phase_ns(const lba_handle* h, int i, double* ns10) {
    if (!h || !ns10 || i < 0 || i >= h->s.nLoaded) return ORB_ERR_ARG;
    const double* stt = (const double*)(h->s.h_arena + h->s.packed[i].statsOff);
    for (int k = 0; k < 10; ++k) ns10[k] = stt[8 + k];
    return ORB_OK;
}

int lba_solve_batch(lba_handle* h, int count, const LbaProblem* problems, LbaResult* results) {
    if (!h || !problems || !results) { set_error("lba_solve_batch: bad argument"); return ORB_ERR_ARG; }
    Solver& S = h->s;
    int rc = S.upload(count, problems);
    if (rc) return rc;
    for (int i = 0; i < count; ++i) S.h_stop[i] = (problems[i].stopFlag && *problems[i].stopFlag) ? 1 : 0;
    rc = S.run(S.st);
    if (rc) return rc;
    bool anyFlag = false;
    for (int i = 0; i < count; ++i) anyFlag = anyFlag || problems[i].stopFlag;
    if (anyFlag) {
        CK(cudaEventRecord(S.evDone, S.st));
        // forward the callers' stop flags (SparseOptimizer::terminate() polls *pbStopFlag) while the kernel runs
        while (cudaEventQuery(S.evDone) == cudaErrorNotReady) {
            for (int i = 0; i < count; ++i)
                if (problems[i].stopFlag && *problems[i].stopFlag) S.h_stop[i] = 1;
            std::this_thread::yield();
        }
    }
    CK(cudaGetLastError());
    return S.download(count, results, S.st);
}
int lba_solve(lba_handle* h, const LbaProblem* problem, LbaResult* result) { return lba_solve_batch(h, 1, problem, result); }

}  // extern "C"
