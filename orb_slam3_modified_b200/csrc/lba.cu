// B200 kernels + C-ABI for the numeric core of Optimizer::LocalBundleAdjustment
// (reference src/Optimizer.cc:1116-1498; g2o BlockSolver_6_3 + Levenberg, see include/orb_b200.h and SURVEY.md 3.2).
//
// Execution model: ONE persistent kernel per batch of problems.  Each problem is owned by one thread-block cluster
// (1..8 CTAs, chosen so that the batch fills the 148 SMs); the whole Levenberg-Marquardt loop -- accept/reject,
// lambda schedule, stop rules, stop-flag polling -- runs on the device, phases are separated by cluster barriers,
// and the host only uploads the flattened graph and downloads the result.  All arithmetic is FP64 (g2o is double).
// Phases of one LM trial (workers = all threads of the cluster):
//   residual   thread/edge      EdgeSE3ProjectXYZ::computeError + Huber rho            (HOT LOOP A)
//   build      warp/point, CTA/pose   linearizeOplus + constructQuadraticForm: Hll, bl, W=Hpl blocks, Hpp, bp   (HOT LOOP B)
//   schur      thread/point, thread/edge, warp/pose-pair   Hll^-1, Y = W Hll^-1, Hschur = Hpp - sum Y W^T       (HOT LOOP C)
//   ldlt       CTA 0, matrix in shared memory   dense LDL^T of the reduced camera system + solve
//   backsub    thread/point     x_l = Hll^-1 (b_l - W^T x_p)
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

namespace lba {

struct Dev {
    int nP, nL, nE, nF, n;            // n = 6 nF
    double *poses, *posesBk, *pts, *ptsBk;      // nP x 7, nL x 3
    const float* cam;                 // nP x 4
    const int* hidx;                  // nP: Hessian block index or -1
    const int* freePose;              // nF: pose index of Hessian block
    // edges in INTERNAL order = sorted by point (the edges of a point are contiguous)
    const int *ePt, *ePose;           // nE
    const int* eOrig;                 // nE: caller's edge index
    const double* obs;                // nE x 2
    const float* invSigma2;           // nE
    const int* ptStart;               // nL + 1
    const int *poseStart, *poseEdges; // CSR by pose (internal edge ids, ascending)
    const int* blockStart;            // nF(nF-1)/2 + 1: off-diagonal Schur blocks (i1 < i2), row-major over the strict upper triangle
    const int2* pairs;                // (edge of pose i1, edge of pose i2) observing the same point
    // Schur work list: tasks 0..nF-1 = diagonal blocks (items = edges of the pose), nF.. = off-diagonal blocks (items = pairs);
    // every task is cut into chunks of SCH items so that all warps of the cluster get the same amount of work
    int nChunks; const int* chunkTask; const int* chunkFirst; const int* taskChunkStart;   // nChunks, nChunks, nTasks + 1
    double* Spart;                    // nChunks x 42 partial sums (36 block entries + 6 rhs entries for diagonal tasks)
    double* err;                      // nE x 2
    double* W;                        // nE x 18 (6x3 Hpl block, zero for fixed poses)
    double* Y;                        // nE x 18 (W Hll^-1)
    double *Hpp, *bp;                 // nF x 36, nF x 6
    double *Hll, *bl;                 // nL x 9, nL x 3
    double *Dinv, *db;                // nL x 9, nL x 3
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
constexpr int NT = 512;
constexpr int NWARP = NT / 32;
constexpr int MAXC = 8;            // largest cluster
constexpr int PSLOT = 16;          // doubles per partial slot
constexpr int PC = 20;             // doubles per cached pose: q(4) t(3) R(9) fx fy cx cy

struct Ctx {
    int crank, csize, tid;
    int wid, nw;           // worker id / count over the cluster
    int slot;              // rotating partial slot
    double* sm;            // NT doubles of scratch
    const double* pc;      // pose cache in shared memory, PC doubles per pose
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
__device__ __forceinline__ void csync() { cg::this_cluster().sync(); }

// (re)load every pose into the CTA's shared-memory cache: quaternion, translation, rotation matrix, camera
__device__ void load_pose_cache(const Dev& D, double* pc) {
    for (int i = threadIdx.x; i < D.nP; i += NT) {
        double* o = pc + PC * i;
        const double* T = D.poses + 7 * (size_t)i;
#pragma unroll
        for (int k = 0; k < 7; ++k) o[k] = T[k];
        qtoR(T, o + 7);
        const float* cm = D.cam + 4 * (size_t)i;
        o[16] = (double)cm[0]; o[17] = (double)cm[1]; o[18] = (double)cm[2]; o[19] = (double)cm[3];   // float params promoted (Pinhole.cpp:35-41)
    }
    __syncthreads();
}

__device__ __forceinline__ void project_edge(const Dev& D, const Ctx& c, int e, double* Xc, double* uv) {
    const double* P = c.pc + PC * D.ePose[e];
    double r[3]; qrot(P, D.pts + 3 * (size_t)D.ePt[e], r);      // SE3Quat::map: _r * xyz + _t
    Xc[0] = r[0] + P[4]; Xc[1] = r[1] + P[5]; Xc[2] = r[2] + P[6];
    uv[0] = P[16] * Xc[0] / Xc[2] + P[18];                      // Pinhole::project(Vector3d)
    uv[1] = P[17] * Xc[1] / Xc[2] + P[19];
}

// ---- residuals + robust chi2 (SparseOptimizer::computeActiveErrors + activeRobustChi2), thread per edge ----
__device__ double phase_errors(const Dev& D, const Ctx& c) {
    double acc = 0;
    for (int e = c.wid; e < D.nE; e += c.nw) {
        double Xc[3], uv[2];
        project_edge(D, c, e, Xc, uv);
        const double e0 = D.obs[2 * (size_t)e] - uv[0], e1 = D.obs[2 * (size_t)e + 1] - uv[1];
        *reinterpret_cast<double2*>(D.err + 2 * (size_t)e) = make_double2(e0, e1);
        double r0, r1;
        robustify(D, (double)D.invSigma2[e] * (e0 * e0 + e1 * e1), r0, r1);
        acc += r0;
    }
    return acc;
}

// Jacobians of one edge (EdgeSE3ProjectXYZ::linearizeOplus): A = dE/dpoint (2x3), B = dE/dpose (2x6)
__device__ __forceinline__ void edge_jacobians(const Dev& D, const Ctx& c, int e, double* A, double* B, double& w, double& r0, double& r1) {
    const double* P = c.pc + PC * D.ePose[e];
    double Xc[3], uv[2];
    project_edge(D, c, e, Xc, uv);
    const double x = Xc[0], y = Xc[1], z = Xc[2];
    const double fx = P[16], fy = P[17];
    const double J00 = -(fx / z), J02 = fx * x / (z * z), J11 = -(fy / z), J12 = fy * y / (z * z);   // -projectJac (Pinhole.cpp:71-81)
    const double* R = P + 7;
#pragma unroll
    for (int k = 0; k < 3; ++k) { A[k] = J00 * R[k] + J02 * R[6 + k]; A[3 + k] = J11 * R[3 + k] + J12 * R[6 + k]; }
    // SE3deriv = [0 z -y 1 0 0; -z 0 x 0 1 0; y -x 0 0 0 1]
    B[0] = J02 * y;            B[1] = J00 * z - J02 * x;  B[2] = -J00 * y;  B[3] = J00; B[4] = 0;   B[5] = J02;
    B[6] = -J11 * z + J12 * y; B[7] = -J12 * x;           B[8] = J11 * x;   B[9] = 0;   B[10] = J11; B[11] = J12;
    const double is2 = (double)D.invSigma2[e];
    const double2 er = *reinterpret_cast<const double2*>(D.err + 2 * (size_t)e);
    double rho0, rho1;
    robustify(D, is2 * (er.x * er.x + er.y * er.y), rho0, rho1);
    w = rho1 * is2;
    r0 = -is2 * er.x * rho1; r1 = -is2 * er.y * rho1;
}

// ---- per point Hll, bl and the Hpl blocks W of its edges; 8 lanes per point (edges of a point are contiguous) ----
__device__ void phase_build_points(const Dev& D, const Ctx& c) {
    const int sl = c.tid & 7;
    const unsigned gmask = 0xFFu << (c.tid & 24);
    for (int p = c.wid >> 3; p < D.nL; p += c.nw >> 3) {
        const int a = D.ptStart[p], b = D.ptStart[p + 1];
        double h[6] = {0, 0, 0, 0, 0, 0}, g[3] = {0, 0, 0};
        for (int e = a + sl; e < b; e += 8) {
            double A[6], B[12], w, r0, r1;
            edge_jacobians(D, c, e, A, B, w, r0, r1);
            g[0] += A[0] * r0 + A[3] * r1; g[1] += A[1] * r0 + A[4] * r1; g[2] += A[2] * r0 + A[5] * r1;
            h[0] += w * (A[0] * A[0] + A[3] * A[3]); h[1] += w * (A[0] * A[1] + A[3] * A[4]); h[2] += w * (A[0] * A[2] + A[3] * A[5]);
            h[3] += w * (A[1] * A[1] + A[4] * A[4]); h[4] += w * (A[1] * A[2] + A[4] * A[5]); h[5] += w * (A[2] * A[2] + A[5] * A[5]);
            double2* We = reinterpret_cast<double2*>(D.W + 18 * (size_t)e);
            if (D.hidx[D.ePose[e]] >= 0) {
                double v[18];
#pragma unroll
                for (int i = 0; i < 6; ++i)
#pragma unroll
                    for (int j = 0; j < 3; ++j) v[i * 3 + j] = w * (B[i] * A[j] + B[6 + i] * A[3 + j]);
#pragma unroll
                for (int i = 0; i < 9; ++i) We[i] = make_double2(v[2 * i], v[2 * i + 1]);
            } else {
#pragma unroll
                for (int i = 0; i < 9; ++i) We[i] = make_double2(0.0, 0.0);
            }
        }
#pragma unroll
        for (int o = 4; o; o >>= 1) {
#pragma unroll
            for (int i = 0; i < 6; ++i) h[i] += __shfl_xor_sync(gmask, h[i], o);
#pragma unroll
            for (int i = 0; i < 3; ++i) g[i] += __shfl_xor_sync(gmask, g[i], o);
        }
        if (sl == 0) {
            double* H = D.Hll + 9 * (size_t)p;
            H[0] = h[0]; H[1] = h[1]; H[2] = h[2]; H[3] = h[1]; H[4] = h[3]; H[5] = h[4]; H[6] = h[2]; H[7] = h[4]; H[8] = h[5];
            D.bl[3 * (size_t)p] = g[0]; D.bl[3 * (size_t)p + 1] = g[1]; D.bl[3 * (size_t)p + 2] = g[2];
        }
    }
}

// ---- per free pose Hpp, bp; one CTA per pose (round-robin over the cluster), threads over its edges ----
__device__ void phase_build_poses(const Dev& D, const Ctx& c) {
    const int lane = c.tid & 31, warp = c.tid >> 5;
    double* red = c.sm;                      // NWARP x 27 <= NT doubles
    for (int hI = c.crank; hI < D.nF; hI += c.csize) {
        const int ic = D.freePose[hI];
        const int a = D.poseStart[ic], b = D.poseStart[ic + 1];
        double acc[27];
#pragma unroll
        for (int i = 0; i < 27; ++i) acc[i] = 0;
        for (int k = a + c.tid; k < b; k += NT) {
            const int e = D.poseEdges[k];
            double A[6], B[12], w, r0, r1;
            edge_jacobians(D, c, e, A, B, w, r0, r1);
            int t = 0;
#pragma unroll
            for (int i = 0; i < 6; ++i) {
#pragma unroll
                for (int j = i; j < 6; ++j) acc[t++] += w * (B[i] * B[j] + B[6 + i] * B[6 + j]);
            }
#pragma unroll
            for (int i = 0; i < 6; ++i) acc[21 + i] += B[i] * r0 + B[6 + i] * r1;
        }
#pragma unroll
        for (int o = 16; o; o >>= 1) {
#pragma unroll
            for (int i = 0; i < 27; ++i) acc[i] += __shfl_xor_sync(0xffffffffu, acc[i], o);
        }
        __syncthreads();
        if (lane == 0) {
#pragma unroll
            for (int i = 0; i < 27; ++i) red[warp * 27 + i] = acc[i];
        }
        __syncthreads();
        if (c.tid < 27) {
            double s = 0;
            for (int w2 = 0; w2 < NWARP; ++w2) s += red[w2 * 27 + c.tid];
            if (c.tid < 21) {
                int i = 0, t = c.tid;           // unrank (i, j), i <= j
                while (t >= 6 - i) { t -= 6 - i; ++i; }
                const int j = i + t;
                D.Hpp[36 * (size_t)hI + i * 6 + j] = s; D.Hpp[36 * (size_t)hI + j * 6 + i] = s;
            } else D.bp[6 * (size_t)hI + (c.tid - 21)] = s;
        }
    }
    __syncthreads();
}

// ---- max |diag| over all Hessian blocks (computeLambdaInit); every CTA computes it redundantly ----
__device__ double phase_maxdiag(const Dev& D, const Ctx& c) {
    double m = 0;
    for (int i = c.tid; i < D.n; i += NT) m = fmax(m, fabs(D.Hpp[36 * (size_t)(i / 6) + 7 * (i % 6)]));
    for (int i = c.tid; i < 3 * D.nL; i += NT) m = fmax(m, fabs(D.Hll[9 * (size_t)(i / 3) + 4 * (i % 3)]));
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

// ---- Hll^-1 (lambda on the diagonal), Hll^-1 bl, and Y = W Hll^-1; thread per edge, the inverse is recomputed per edge
//      (cheaper than a barrier), the first edge of a point stores it (block_solver.hpp:381-394) ----
__device__ void phase_point_prep(const Dev& D, const Ctx& c, double lambda) {
    for (int e = c.wid; e < D.nE; e += c.nw) {
        const int p = D.ePt[e];
        double m[9];
#pragma unroll
        for (int i = 0; i < 9; ++i) m[i] = D.Hll[9 * (size_t)p + i];
        m[0] += lambda; m[4] += lambda; m[8] += lambda;
        const double c00 = m[4] * m[8] - m[5] * m[7], c10 = m[5] * m[6] - m[3] * m[8], c20 = m[3] * m[7] - m[4] * m[6];
        const double id = 1.0 / (m[0] * c00 + m[1] * c10 + m[2] * c20);   // Eigen 3x3 inverse: cofactors / determinant
        double o[9];
        o[0] = c00 * id; o[1] = (m[2] * m[7] - m[1] * m[8]) * id; o[2] = (m[1] * m[5] - m[2] * m[4]) * id;
        o[3] = c10 * id; o[4] = (m[0] * m[8] - m[2] * m[6]) * id; o[5] = (m[2] * m[3] - m[0] * m[5]) * id;
        o[6] = c20 * id; o[7] = (m[1] * m[6] - m[0] * m[7]) * id; o[8] = (m[0] * m[4] - m[1] * m[3]) * id;
        if (e == D.ptStart[p]) {
#pragma unroll
            for (int i = 0; i < 9; ++i) D.Dinv[9 * (size_t)p + i] = o[i];
            const double* b3 = D.bl + 3 * (size_t)p;
#pragma unroll
            for (int i = 0; i < 3; ++i) D.db[3 * (size_t)p + i] = o[i * 3] * b3[0] + o[i * 3 + 1] * b3[1] + o[i * 3 + 2] * b3[2];
        }
        const double2* Wd = reinterpret_cast<const double2*>(D.W + 18 * (size_t)e);
        double wv[18];
#pragma unroll
        for (int i = 0; i < 9; ++i) { const double2 t = Wd[i]; wv[2 * i] = t.x; wv[2 * i + 1] = t.y; }
        double yv[18];
#pragma unroll
        for (int a = 0; a < 6; ++a)
#pragma unroll
            for (int b = 0; b < 3; ++b) yv[a * 3 + b] = wv[a * 3] * o[b] + wv[a * 3 + 1] * o[3 + b] + wv[a * 3 + 2] * o[6 + b];
        double2* Yd = reinterpret_cast<double2*>(D.Y + 18 * (size_t)e);
#pragma unroll
        for (int i = 0; i < 9; ++i) Yd[i] = make_double2(yv[2 * i], yv[2 * i + 1]);
    }
}

__device__ __forceinline__ void load18(const double* p, double* v) {
    const double2* q = reinterpret_cast<const double2*>(p);
#pragma unroll
    for (int i = 0; i < 9; ++i) { const double2 t = q[i]; v[2 * i] = t.x; v[2 * i + 1] = t.y; }
}
// ---- Schur complement (block_solver.hpp:396-431), two steps.
//   partial: the work list (diagonal block i: edges of pose i, contributes Y_e W_e^T and W_e Hll^-1 bl;
//            off-diagonal block (i1 < i2): the precomputed (e1, e2) pairs, contributes Y_e1 W_e2^T) is cut into chunks of
//            SCH items; one warp per chunk, two lanes share one item (each owns three rows of the 6x6 product), 16 items
//            in flight per warp; the chunk's sums go to Spart.
//   combine: thread per (task, entry): ordered sum of the task's chunks -> lower triangle of the reduced camera system
//            (diagonal: Hpp_i + lambda I - sum, bs_i = bp_i - sum; off-diagonal: block (i2, i1) = -(sum)^T). ----
constexpr int SCH = 128;
__device__ void phase_schur_partial(const Dev& D, const Ctx& c) {
    const int lane = c.tid & 31, half = lane & 1, slot = lane >> 1;
    for (int ch = c.crank * NWARP + (c.tid >> 5); ch < D.nChunks; ch += c.csize * NWARP) {
        const int task = D.chunkTask[ch], first = D.chunkFirst[ch];
        double acc[18], bacc[3] = {0, 0, 0};
#pragma unroll
        for (int i = 0; i < 18; ++i) acc[i] = 0;
        if (task >= D.nF) {
            const int blk = task - D.nF;
            const int k0 = D.blockStart[blk] + first, k1 = min(k0 + SCH, D.blockStart[blk + 1]);
            int2 pr[SCH / 16];
#pragma unroll
            for (int j = 0; j < SCH / 16; ++j) { const int k = k0 + slot + 16 * j; pr[j] = k < k1 ? D.pairs[k] : make_int2(-1, -1); }
#pragma unroll 2
            for (int j = 0; j < SCH / 16; ++j) {
                if (pr[j].x < 0) continue;
                const double* yp = D.Y + 18 * (size_t)pr[j].x + 9 * half;
                double y[9], w[18];
#pragma unroll
                for (int i = 0; i < 9; ++i) y[i] = yp[i];
                load18(D.W + 18 * (size_t)pr[j].y, w);
#pragma unroll
                for (int r = 0; r < 3; ++r)
#pragma unroll
                    for (int q = 0; q < 6; ++q) acc[r * 6 + q] += y[r * 3] * w[q * 3] + y[r * 3 + 1] * w[q * 3 + 1] + y[r * 3 + 2] * w[q * 3 + 2];
            }
        } else {
            const int ic = D.freePose[task];
            const int k0 = D.poseStart[ic] + first, k1 = min(k0 + SCH, D.poseStart[ic + 1]);
            int ed[SCH / 16];
#pragma unroll
            for (int j = 0; j < SCH / 16; ++j) { const int k = k0 + slot + 16 * j; ed[j] = k < k1 ? D.poseEdges[k] : -1; }
#pragma unroll 2
            for (int j = 0; j < SCH / 16; ++j) {
                const int e = ed[j];
                if (e < 0) continue;
                const double* yp = D.Y + 18 * (size_t)e + 9 * half;
                double y[9], w[18];
#pragma unroll
                for (int i = 0; i < 9; ++i) y[i] = yp[i];
                load18(D.W + 18 * (size_t)e, w);
                const double* dbp = D.db + 3 * (size_t)D.ePt[e];
                const double d0 = dbp[0], d1 = dbp[1], d2 = dbp[2];
#pragma unroll
                for (int r = 0; r < 3; ++r)
#pragma unroll
                    for (int q = 0; q < 6; ++q) acc[r * 6 + q] += y[r * 3] * w[q * 3] + y[r * 3 + 1] * w[q * 3 + 1] + y[r * 3 + 2] * w[q * 3 + 2];
#pragma unroll
                for (int r = 0; r < 3; ++r) {
                    const double w0 = half ? w[9 + 3 * r] : w[3 * r], w1 = half ? w[10 + 3 * r] : w[1 + 3 * r], w2 = half ? w[11 + 3 * r] : w[2 + 3 * r];
                    bacc[r] += w0 * d0 + w1 * d1 + w2 * d2;
                }
            }
        }
#pragma unroll
        for (int o = 16; o >= 2; o >>= 1) {
#pragma unroll
            for (int i = 0; i < 18; ++i) acc[i] += __shfl_xor_sync(0xffffffffu, acc[i], o);
#pragma unroll
            for (int i = 0; i < 3; ++i) bacc[i] += __shfl_xor_sync(0xffffffffu, bacc[i], o);
        }
        if (slot == 0) {
            double* out = D.Spart + 42 * (size_t)ch;
#pragma unroll
            for (int i = 0; i < 18; ++i) out[18 * half + i] = acc[i];
#pragma unroll
            for (int i = 0; i < 3; ++i) out[36 + 3 * half + i] = bacc[i];
        }
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
// Blocked by the natural 6x6 pose blocks: every thread that owns a row below the diagonal block re-factors that block in
// registers (no barrier between "factor" and "panel"), then all warps apply the rank-6 trailing update.
// (LinearSolverEigen::solve: SimplicialLDLT fails only on an exactly zero pivot.)  Returns 1 on success (uniform).
// xrow: n x 6 scratch (unscaled panel), in shared memory. ----
__device__ int phase_ldlt(const Dev& D, const Ctx& c, double* A, int ld, double* xrow) {
    const int n = D.n, tid = c.tid, lane = tid & 31, warp = tid >> 5;
    __shared__ double s_L[36], s_dinv[6];
    __shared__ int s_fail;
    if (tid == 0) s_fail = 0;
    __syncthreads();
    for (int k0 = 0; k0 < n; k0 += 6) {
        // (1) one thread factors the 6x6 diagonal block: L11 unit lower, d[6]; one reciprocal per pivot
        if (tid == 0) {
            double L[36], d[6];
#pragma unroll
            for (int i = 0; i < 6; ++i)
#pragma unroll
                for (int j = 0; j <= i; ++j) L[i * 6 + j] = A[(size_t)(k0 + i) * ld + k0 + j];
            bool bad = false;
#pragma unroll
            for (int k = 0; k < 6; ++k) {
                double dk = L[k * 6 + k];
#pragma unroll
                for (int j = 0; j < k; ++j) dk -= L[k * 6 + j] * L[k * 6 + j] * d[j];
                d[k] = dk;
                if (dk == 0.0) bad = true;
                const double inv = 1.0 / dk;
                s_dinv[k] = inv;
#pragma unroll
                for (int i = k + 1; i < 6; ++i) {
                    double v = L[i * 6 + k];
#pragma unroll
                    for (int j = 0; j < k; ++j) v -= L[i * 6 + j] * L[k * 6 + j] * d[j];
                    L[i * 6 + k] = v * inv;
                }
            }
#pragma unroll
            for (int i = 0; i < 6; ++i) {
                A[(size_t)(k0 + i) * ld + k0 + i] = d[i];
#pragma unroll
                for (int j = 0; j < i; ++j) { A[(size_t)(k0 + i) * ld + k0 + j] = L[i * 6 + j]; s_L[i * 6 + j] = L[i * 6 + j]; }
            }
            if (bad) s_fail = 1;
        }
        __syncthreads();
        if (s_fail) return 0;
        // (2) panel: row i below the block: X = A21 L11^-T (unscaled), L21 = X D^-1
        for (int i = k0 + 6 + tid; i < n; i += NT) {
            double xr[6];
#pragma unroll
            for (int cidx = 0; cidx < 6; ++cidx) {
                double v = A[(size_t)i * ld + k0 + cidx];
#pragma unroll
                for (int j = 0; j < cidx; ++j) v -= xr[j] * s_L[cidx * 6 + j];
                xr[cidx] = v;
            }
#pragma unroll
            for (int cidx = 0; cidx < 6; ++cidx) { xrow[i * 6 + cidx] = xr[cidx]; A[(size_t)i * ld + k0 + cidx] = xr[cidx] * s_dinv[cidx]; }
        }
        __syncthreads();
        // (3) trailing update: A22[i][j] -= sum_c X[i][c] * L21[j][c], warp per row
        for (int i = k0 + 6 + warp; i < n; i += NWARP) {
            double xi[6];
#pragma unroll
            for (int cidx = 0; cidx < 6; ++cidx) xi[cidx] = xrow[i * 6 + cidx];
            for (int j = k0 + 6 + lane; j <= i; j += 32) {
                const double* lj = A + (size_t)j * ld + k0;
                double sacc = 0;
#pragma unroll
                for (int cidx = 0; cidx < 6; ++cidx) sacc += xi[cidx] * lj[cidx];
                A[(size_t)i * ld + j] -= sacc;
            }
        }
        __syncthreads();
    }
    // L D L^T x = bs by warp 0 alone (warp-synchronous column-oriented substitutions), y kept in shared memory
    double* y = xrow;   // the panel scratch is free now (n <= 6 n doubles)
    if (warp == 0) {
        for (int i = lane; i < n; i += 32) y[i] = D.bs[i];
        __syncwarp();
        for (int k = 0; k < n; ++k) {
            const double yk = y[k];
            for (int i = k + 1 + lane; i < n; i += 32) y[i] -= A[(size_t)i * ld + k] * yk;
            __syncwarp();
        }
        for (int i = lane; i < n; i += 32) y[i] /= A[(size_t)i * ld + i];
        __syncwarp();
        for (int k = n - 1; k >= 0; --k) {
            const double yk = y[k];
            for (int j = lane; j < k; j += 32) y[j] -= A[(size_t)k * ld + j] * yk;
            __syncwarp();
        }
        for (int i = lane; i < n; i += 32) D.x[i] = y[i];
    }
    __syncthreads();
    return 1;
}
// ---- landmark back-substitution x_l = Hll^-1 (bl - W^T x_p)  (block_solver.hpp:461-483); 8 lanes per point ----
__device__ void phase_backsub(const Dev& D, const Ctx& c) {
    const int sl = c.tid & 7;
    const unsigned gmask = 0xFFu << (c.tid & 24);
    for (int p = c.wid >> 3; p < D.nL; p += c.nw >> 3) {
        double cl[3] = {0, 0, 0};
        for (int e = D.ptStart[p] + sl; e < D.ptStart[p + 1]; e += 8) {
            const int h = D.hidx[D.ePose[e]];
            if (h < 0) continue;
            double w[18];
            load18(D.W + 18 * (size_t)e, w);
            const double* xp = D.x + 6 * (size_t)h;
#pragma unroll
            for (int b = 0; b < 3; ++b)
#pragma unroll
                for (int a = 0; a < 6; ++a) cl[b] -= w[a * 3 + b] * xp[a];
        }
#pragma unroll
        for (int o = 4; o; o >>= 1) {
#pragma unroll
            for (int i = 0; i < 3; ++i) cl[i] += __shfl_xor_sync(gmask, cl[i], o);
        }
        if (sl < 3) {
            const double* Di = D.Dinv + 9 * (size_t)p;
            const double* b3 = D.bl + 3 * (size_t)p;
            const double c0 = cl[0] + b3[0], c1 = cl[1] + b3[1], c2 = cl[2] + b3[2];
            D.x[D.n + 3 * (size_t)p + sl] = Di[sl * 3] * c0 + Di[sl * 3 + 1] * c1 + Di[sl * 3 + 2] * c2;
        }
    }
}
// ---- push + update (SparseOptimizer::push / update) and the partial sum of computeScale ----
__device__ double phase_update(const Dev& D, const Ctx& c, double lambda) {
    double acc = 0;
    const int total = D.nP + D.nL;
    for (int v = c.wid; v < total; v += c.nw) {
        if (v < D.nP) {
            double T[7];
#pragma unroll
            for (int i = 0; i < 7; ++i) { T[i] = D.poses[7 * (size_t)v + i]; D.posesBk[7 * (size_t)v + i] = T[i]; }
            const int h = D.hidx[v];
            if (h >= 0) {
                const double* u = D.x + 6 * (size_t)h;
                pose_oplus(T, u);
#pragma unroll
                for (int i = 0; i < 7; ++i) D.poses[7 * (size_t)v + i] = T[i];
#pragma unroll
                for (int i = 0; i < 6; ++i) acc += u[i] * (lambda * u[i] + D.bp[6 * (size_t)h + i]);
            }
        } else {
            const int p = v - D.nP;
#pragma unroll
            for (int i = 0; i < 3; ++i) {
                const double u = D.x[D.n + 3 * (size_t)p + i];
                const double old = D.pts[3 * (size_t)p + i];
                D.ptsBk[3 * (size_t)p + i] = old;
                D.pts[3 * (size_t)p + i] = old + u;
                acc += u * (lambda * u + D.bl[3 * (size_t)p + i]);
            }
        }
    }
    return acc;
}
__device__ void phase_restore(const Dev& D, const Ctx& c) {   // pop
    for (int i = c.wid; i < 7 * D.nP; i += c.nw) D.poses[i] = D.posesBk[i];
    for (int i = c.wid; i < 3 * D.nL; i += c.nw) D.pts[i] = D.ptsBk[i];
}
__device__ void phase_finalize(const Dev& D, const Ctx& c) {
    for (int e = c.wid; e < D.nE; e += c.nw) {
        const double e0 = D.err[2 * (size_t)e], e1 = D.err[2 * (size_t)e + 1];
        const int o = D.eOrig[e];
        D.outChi2[o] = (double)D.invSigma2[e] * (e0 * e0 + e1 * e1);   // e->chi2() from the last computed _error (Optimizer.cc:1425)
        double Xc[3], uv[2];
        project_edge(D, c, e, Xc, uv);
        D.outDepthPos[o] = Xc[2] > 0.0;
    }
}

// =============================================================================================
// The persistent kernel: SparseOptimizer::optimize (sparse_optimizer.cpp:354-418) around
// OptimizationAlgorithmLevenberg::solve (optimization_algorithm_levenberg.cpp:61-168), one cluster per problem.
// Every thread carries the (uniform) LM state; decisions use values every CTA reads identically after a cluster barrier.
// Dynamic shared memory: pose cache (PC x maxP doubles) | LDLT panel scratch (6 maxP x 6) | reduced camera system (smemMatrixN^2).
// =============================================================================================
__global__ void __launch_bounds__(NT, 1) lba_cluster_kernel(const Dev* __restrict__ probs, const volatile int* stop, int smemMatrixN, int maxP) {
    extern __shared__ double s_dyn[];
    __shared__ double s_red[NT];
    cg::cluster_group cluster = cg::this_cluster();
    Ctx c;
    c.crank = (int)cluster.block_rank(); c.csize = (int)cluster.num_blocks(); c.tid = threadIdx.x;
    c.wid = c.crank * NT + c.tid; c.nw = c.csize * NT; c.slot = 0; c.sm = s_red;
    double* s_pc = s_dyn;
    double* s_col = s_pc + PC * maxP;
    double* s_mat = s_col + 36 * maxP;
    c.pc = s_pc;
    const Dev D = probs[blockIdx.x / c.csize];
    const bool matInSmem = D.n <= smemMatrixN && D.n > 0;
    // reduced camera system: CTA 0's shared memory; the other CTAs of the cluster reach it as distributed shared memory
    double* HsRemote = matInSmem ? cluster.map_shared_rank(s_mat, 0) : D.Hs;
    double* HsLocal = matInSmem ? s_mat : D.Hs;
    const int ld = D.n;
    const int stopIdx = blockIdx.x / c.csize;
    const bool poller = c.crank == 0 && c.tid == 0;

    for (int i = c.wid; i < D.nP; i += c.nw) qnormalize(D.poses + 7 * (size_t)i);   // SE3Quat(q, t) constructor
    for (int i = c.wid; i < 2 * D.nE; i += c.nw) D.err[i] = 0.0;
    for (int i = c.wid; i < D.n + 3 * D.nL; i += c.nw) D.x[i] = 0.0;
    int term = 0, dummy;
    cluster_sum(D, c, 0.0, (poller && stop && stop[stopIdx]) ? 1 : 0, term);   // initial terminate() poll + publishes the poses
    load_pose_cache(D, s_pc);

    double lambda = -1, ni = 2, currentChi = 0, firstChi = 0;
    int nBad = 0, cj = 0, trials = 0;
    const int maxTrials = 10;
    const double goodUpper = 2. / 3., goodLower = 1. / 3., tau = 1e-5;
    bool ok = true;
    unsigned long long tph[10] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0}, t0;
#define TICK() t0 = globaltimer_ns()
#define TOCK(i) tph[i] += globaltimer_ns() - t0
    for (int it = 0; it < D.iterations && !term && ok; ++it) {
        TICK();
        currentChi = cluster_sum(D, c, phase_errors(D, c), 0, dummy);
        TOCK(0);
        double tempChi = currentChi;
        const double iniChi = currentChi;
        if (it == 0) firstChi = iniChi;
        TICK();
        phase_build_points(D, c);
        TOCK(1); TICK();
        phase_build_poses(D, c);
        csync();
        TOCK(2);
        if (it == 0) {
            if (D.userLambdaInit > 0) lambda = D.userLambdaInit;
            else lambda = tau * phase_maxdiag(D, c);
            ni = 2; nBad = 0;
        }
        double rho = 0;
        int qmax = 0;
        do {
            TICK();
            phase_point_prep(D, c, lambda);
            csync();
            TOCK(3);
            int ok2 = 1;
            if (D.nF) {
                TICK();
                phase_schur_partial(D, c);
                csync();
                phase_schur_combine(D, c, lambda, HsRemote, ld);   // every CTA writes entries into CTA 0's shared memory through DSMEM
                csync();
                TOCK(4); TICK();
                if (c.crank == 0) {
                    ok2 = phase_ldlt(D, c, HsLocal, ld, s_col);
                    if (c.tid == 0) D.partial[4 * PSLOT] = (double)ok2;
                }
                csync();
                ok2 = (int)D.partial[4 * PSLOT];
                TOCK(5);
            }
            TICK();
            phase_backsub(D, c);
            csync();
            TOCK(6); TICK();
            const double scale0 = cluster_sum(D, c, phase_update(D, c, lambda), 0, dummy);
            load_pose_cache(D, s_pc);
            TOCK(7); TICK();
            tempChi = cluster_sum(D, c, phase_errors(D, c), (poller && stop && stop[stopIdx]) ? 1 : 0, term);
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
            } else {
                lambda *= ni;
                ni *= 2;
                phase_restore(D, c);
                csync();
                load_pose_cache(D, s_pc);
            }
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
    if (c.wid == 0) for (int i = 0; i < 10; ++i) D.stats[8 + i] = (double)tph[i];
    if (c.wid == 0) { D.stats[0] = cj; D.stats[1] = trials; D.stats[2] = lambda; D.stats[3] = currentChi; D.stats[4] = firstChi; }
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
    cudaEvent_t evDone = nullptr;
    int nLoaded = 0, launches = 0, smemN = 0, numSMs = 148;
    size_t fixedSmem = 0;
    ~Solver() {
        cudaSetDevice(device);
        if (d_arena) cudaFree(d_arena);
        if (h_arena) cudaFreeHost(h_arena);
        if (d_probs) cudaFree(d_probs);
        if (h_stop) cudaFreeHost(h_stop);
        if (evDone) cudaEventDestroy(evDone);
        if (st) cudaStreamDestroy(st);
    }
    static size_t al(size_t b) { return (b + 255) & ~(size_t)255; }
    static size_t need(size_t nP, size_t nL, size_t nE) {
        const size_t n = 6 * nP;
        size_t b = 0;
        b += 2 * al(56 * nP) + 2 * al(24 * nL) + al(56 * nP) + al(24 * nL);      // poses, bk, pts, bk, initial copies
        b += al(16 * nP) + 2 * al(4 * nP);                                        // cam, hidx, freePose
        b += 2 * al(4 * nE) + al(16 * nE) + al(4 * nE);                           // ePt, ePose, obs, invSigma2
        b += al(4 * (nL + 1)) + al(4 * nE) + al(4 * (nP + 1)) + al(4 * nE);                     // CSR, eOrig
        b += al(4 * (nP * nP / 2 + 2)) + al(8 * (nE * (nP > 1 ? nP - 1 : 1) / 2 + 1));              // Schur block starts + (e1, e2) pair lists
        b += al(16 * nE) + 2 * al(144 * nE);                                      // err, W, Y
        b += al(288 * nP) + al(48 * nP) + 2 * al(72 * nL) + 2 * al(24 * nL);      // Hpp, bp, Hll, Dinv, bl, db
        b += al(8 * n * n) + al(8 * n) + al(8 * (n + 3 * nL)) + al(8 * (4 * PSLOT + 8)) + al(256);   // Hs, bs, x, partial, stats
        {   // Schur chunk tables + partial sums: at most (#tasks + #items / SCH) chunks
            const size_t items = nE + nE * (nP > 1 ? nP - 1 : 1) / 2 + 1, tasks = nP + nP * nP / 2 + 2;
            const size_t ch = tasks + items / SCH + 1;
            b += 2 * al(4 * ch) + al(4 * (tasks + 1)) + al(8 * 42 * ch);
        }
        b += al(8 * nE) + al(nE);                                                 // outputs
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
        CK(cudaMallocHost(&h_arena, arenaBytes));
        CK(cudaMalloc(&d_probs, sizeof(Dev) * maxBatch));
        CK(cudaHostAlloc(&h_stop, sizeof(int) * maxBatch, cudaHostAllocMapped));
        CK(cudaHostGetDevicePointer(&d_stop, h_stop, 0));
        CK(cudaEventCreateWithFlags(&evDone, cudaEventDisableTiming));
        CK(cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking));
        // dynamic shared memory: pose cache + two LDLT column buffers + (when it fits) the reduced camera system,
        // within 200 KB of the 227 KB opt-in limit
        fixedSmem = 8 * (size_t)(PC + 36) * maxP;
        const size_t maxDyn = 200 * 1024;
        smemN = fixedSmem < maxDyn ? (int)floor(sqrt((double)(maxDyn - fixedSmem) / 8.0)) : 0;
        smemN = std::min(smemN, 6 * maxP);
        CK(cudaFuncSetAttribute(lba_cluster_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(fixedSmem + 8 * (size_t)smemN * smemN)));
        h_probs.resize(maxBatch); packed.resize(maxBatch);
        return ORB_OK;
    }

    // BlockSolver::buildStructure (block_solver.hpp:143-295) on the host: index maps, CSR lists, the (point, pose) -> edge table;
    // packs one problem into slot `slot` of the pinned arena.
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
        // internal edge order: stable sort by point, so that the edges of a point are contiguous
        std::vector<int> ptStart(nL + 1, 0), poseStart(nP + 1, 0), eOrig(nE), inv(nE), poseEdges(nE);
        for (int e = 0; e < nE; ++e) {
            const int p = P->edgePoint[e], c = P->edgePose[e];
            if (p < 0 || p >= nL || c < 0 || c >= nP) { set_error("lba: edge index out of range"); return ORB_ERR_ARG; }
            ++ptStart[p + 1]; ++poseStart[c + 1];
        }
        for (int i = 0; i < nL; ++i) ptStart[i + 1] += ptStart[i];
        for (int i = 0; i < nP; ++i) poseStart[i + 1] += poseStart[i];
        {
            std::vector<int> a(ptStart.begin(), ptStart.end() - 1);
            for (int e = 0; e < nE; ++e) { const int k = a[P->edgePoint[e]]++; eOrig[k] = e; inv[e] = k; }
            std::vector<int> b(poseStart.begin(), poseStart.end() - 1);
            for (int k = 0; k < nE; ++k) poseEdges[b[P->edgePose[eOrig[k]]]++] = k;     // ascending internal ids per pose
        }
        std::vector<int> ePt(nE), ePose(nE); std::vector<double> obs(2 * (size_t)nE); std::vector<float> is2(nE);
        for (int k = 0; k < nE; ++k) {
            const int e = eOrig[k];
            ePt[k] = P->edgePoint[e]; ePose[k] = P->edgePose[e]; obs[2 * (size_t)k] = P->obs[2 * (size_t)e]; obs[2 * (size_t)k + 1] = P->obs[2 * (size_t)e + 1];
            is2[k] = P->invSigma2[e];
        }
        // Schur structure: for every point, every pair of its free-pose edges (i1 < i2) goes to block (i1, i2)
        const int nOff = nF * (nF - 1) / 2;
        auto blockOf = [&](int i1, int i2) { return i1 * (nF - 1) - i1 * (i1 - 1) / 2 + (i2 - i1 - 1); };
        std::vector<int> blockStart(nOff + 1, 0);
        std::vector<int> fe;   // free edges of the current point: (hidx, internal id)
        size_t nPairs = 0;
        for (int p = 0; p < nL; ++p) {
            fe.clear();
            for (int k = ptStart[p]; k < ptStart[p + 1]; ++k) if (hidx[ePose[k]] >= 0) fe.push_back(k);
            for (size_t a2 = 0; a2 < fe.size(); ++a2)
                for (size_t b2 = a2 + 1; b2 < fe.size(); ++b2) {
                    const int h1 = hidx[ePose[fe[a2]]], h2 = hidx[ePose[fe[b2]]];
                    if (h1 == h2) { set_error("lba: duplicate (point, keyframe) observation"); return ORB_ERR_ARG; }
                    ++blockStart[blockOf(std::min(h1, h2), std::max(h1, h2)) + 1];
                    ++nPairs;
                }
        }
        if (nPairs > (size_t)nE * (nP > 1 ? nP - 1 : 1) / 2 + 1) { set_error("lba: pair list larger than sized"); return ORB_ERR_CAPACITY; }
        for (int i = 0; i < nOff; ++i) blockStart[i + 1] += blockStart[i];
        std::vector<int2> pairs(nPairs);
        {
            std::vector<int> cur(blockStart.begin(), blockStart.end() - 1);
            for (int p = 0; p < nL; ++p) {
                fe.clear();
                for (int k = ptStart[p]; k < ptStart[p + 1]; ++k) if (hidx[ePose[k]] >= 0) fe.push_back(k);
                for (size_t a2 = 0; a2 < fe.size(); ++a2)
                    for (size_t b2 = a2 + 1; b2 < fe.size(); ++b2) {
                        int k1 = fe[a2], k2 = fe[b2];
                        if (hidx[ePose[k1]] > hidx[ePose[k2]]) std::swap(k1, k2);
                        const int blk = blockOf(hidx[ePose[k1]], hidx[ePose[k2]]);
                        pairs[cur[blk]++] = make_int2(k1, k2);
                    }
            }
        }
        // Schur work list cut into chunks of SCH items
        std::vector<int> chunkTask, chunkFirst, taskChunkStart(nF + nOff + 1, 0);
        for (int t = 0; t < nF + nOff; ++t) {
            const int len = t < nF ? poseStart[freePose[t] + 1] - poseStart[freePose[t]] : blockStart[t - nF + 1] - blockStart[t - nF];
            taskChunkStart[t] = (int)chunkTask.size();
            for (int f = 0; f < len; f += SCH) { chunkTask.push_back(t); chunkFirst.push_back(f); }
        }
        taskChunkStart[nF + nOff] = (int)chunkTask.size();
        const int nChunks = (int)chunkTask.size();
        const size_t base = perProblem * (size_t)slot;
        size_t off = base;
        auto carve = [&](size_t bytes) { size_t o = off; off += al(bytes); return o; };
        auto put = [&](size_t o, const void* src, size_t bytes) { if (bytes) memcpy(h_arena + o, src, bytes); };
        Dev& D = h_probs[slot]; memset(&D, 0, sizeof(D));
        Packed& K = packed[slot];
        K.nP = nP; K.nL = nL; K.nE = nE; K.nF = nF;
        D.nP = nP; D.nL = nL; D.nE = nE; D.nF = nF; D.n = n;
        // --- uploaded block (host content matters) ---
        K.initPosesOff = carve(56 * (size_t)nP); put(K.initPosesOff, P->poses, 56 * (size_t)nP);
        K.initPtsOff = carve(24 * (size_t)nL); put(K.initPtsOff, P->points, 24 * (size_t)nL);
        size_t o;
        o = carve(16 * (size_t)nP); put(o, P->cam, 16 * (size_t)nP); D.cam = (const float*)(d_arena + o);
        o = carve(4 * (size_t)nP); put(o, hidx.data(), 4 * (size_t)nP); D.hidx = (const int*)(d_arena + o);
        o = carve(4 * (size_t)std::max(nF, 1)); put(o, freePose.data(), 4 * (size_t)nF); D.freePose = (const int*)(d_arena + o);
        o = carve(4 * (size_t)nE); put(o, ePt.data(), 4 * (size_t)nE); D.ePt = (const int*)(d_arena + o);
        o = carve(4 * (size_t)nE); put(o, ePose.data(), 4 * (size_t)nE); D.ePose = (const int*)(d_arena + o);
        o = carve(4 * (size_t)nE); put(o, eOrig.data(), 4 * (size_t)nE); D.eOrig = (const int*)(d_arena + o);
        o = carve(16 * (size_t)nE); put(o, obs.data(), 16 * (size_t)nE); D.obs = (const double*)(d_arena + o);
        o = carve(4 * (size_t)nE); put(o, is2.data(), 4 * (size_t)nE); D.invSigma2 = (const float*)(d_arena + o);
        o = carve(4 * (size_t)(nL + 1)); put(o, ptStart.data(), 4 * (size_t)(nL + 1)); D.ptStart = (const int*)(d_arena + o);
        o = carve(4 * (size_t)(nP + 1)); put(o, poseStart.data(), 4 * (size_t)(nP + 1)); D.poseStart = (const int*)(d_arena + o);
        o = carve(4 * (size_t)nE); put(o, poseEdges.data(), 4 * (size_t)nE); D.poseEdges = (const int*)(d_arena + o);
        o = carve(4 * (size_t)(nOff + 1)); put(o, blockStart.data(), 4 * (size_t)(nOff + 1)); D.blockStart = (const int*)(d_arena + o);
        o = carve(8 * std::max<size_t>(nPairs, 1)); put(o, pairs.data(), 8 * nPairs); D.pairs = (const int2*)(d_arena + o);
        D.nChunks = nChunks;
        o = carve(4 * (size_t)std::max(nChunks, 1)); put(o, chunkTask.data(), 4 * (size_t)nChunks); D.chunkTask = (const int*)(d_arena + o);
        o = carve(4 * (size_t)std::max(nChunks, 1)); put(o, chunkFirst.data(), 4 * (size_t)nChunks); D.chunkFirst = (const int*)(d_arena + o);
        o = carve(4 * (size_t)(nF + nOff + 1)); put(o, taskChunkStart.data(), 4 * (size_t)(nF + nOff + 1)); D.taskChunkStart = (const int*)(d_arena + o);
        uploadBytes[slot] = off - base;
        // --- device-only scratch / state / outputs ---
        K.posesOff = carve(56 * (size_t)nP); D.poses = (double*)(d_arena + K.posesOff);
        D.posesBk = (double*)(d_arena + carve(56 * (size_t)nP));
        K.ptsOff = carve(24 * (size_t)nL); D.pts = (double*)(d_arena + K.ptsOff);
        D.ptsBk = (double*)(d_arena + carve(24 * (size_t)nL));
        D.err = (double*)(d_arena + carve(16 * (size_t)nE));
        D.W = (double*)(d_arena + carve(144 * (size_t)nE)); D.Y = (double*)(d_arena + carve(144 * (size_t)nE));
        D.Hpp = (double*)(d_arena + carve(288 * (size_t)std::max(nF, 1))); D.bp = (double*)(d_arena + carve(48 * (size_t)std::max(nF, 1)));
        D.Hll = (double*)(d_arena + carve(72 * (size_t)nL)); D.bl = (double*)(d_arena + carve(24 * (size_t)nL));
        D.Dinv = (double*)(d_arena + carve(72 * (size_t)nL)); D.db = (double*)(d_arena + carve(24 * (size_t)nL));
        D.Spart = (double*)(d_arena + carve(8 * 42 * (size_t)std::max(nChunks, 1)));
        D.Hs = (double*)(d_arena + carve(8 * (size_t)n * n)); D.bs = (double*)(d_arena + carve(8 * (size_t)std::max(n, 1)));
        D.x = (double*)(d_arena + carve(8 * ((size_t)n + 3 * (size_t)nL)));
        D.partial = (double*)(d_arena + carve(8 * (4 * PSLOT + 8)));
        K.statsOff = carve(256); D.stats = (double*)(d_arena + K.statsOff);
        K.chi2Off = carve(8 * (size_t)nE); D.outChi2 = (double*)(d_arena + K.chi2Off);
        K.dposOff = carve((size_t)nE); D.outDepthPos = (uint8_t*)(d_arena + K.dposOff);
        if (off - base > perProblem) { set_error("lba: arena too small (internal sizing error)"); return ORB_ERR_CAPACITY; }
        D.delta = P->huberDelta; D.dsqr = P->huberDelta * P->huberDelta; D.userLambdaInit = P->userLambdaInit; D.iterations = P->iterations;
        return ORB_OK;
    }
    std::vector<size_t> uploadBytes = std::vector<size_t>(1);

    int upload(int count, const LbaProblem* probs) {
        if (count < 1 || count > maxBatch) { set_error("lba: batch larger than max_batch"); return ORB_ERR_ARG; }
        uploadBytes.assign(maxBatch, 0);
        {   // BlockSolver::buildStructure for every problem, on the host cores (problems are independent)
            const int nth = std::max(1, std::min<int>({count, (int)std::thread::hardware_concurrency(), 16}));
            std::atomic<int> next(0), firstErr(ORB_OK);
            auto worker = [&]() {
                for (int i = next.fetch_add(1); i < count; i = next.fetch_add(1)) {
                    const int rc = pack(i, probs + i);
                    if (rc) { int exp = ORB_OK; firstErr.compare_exchange_strong(exp, rc); }
                }
            };
            std::vector<std::thread> pool;
            for (int t = 1; t < nth; ++t) pool.emplace_back(worker);
            worker();
            for (auto& t : pool) t.join();
            if (firstErr.load()) { set_error("lba: malformed problem in the batch (index out of range, duplicate observation, or larger than the handle)"); return firstErr.load(); }
        }
        CK(cudaSetDevice(device));
        for (int i = 0; i < count; ++i)
            CK(cudaMemcpyAsync(d_arena + perProblem * (size_t)i, h_arena + perProblem * (size_t)i, uploadBytes[i], cudaMemcpyHostToDevice, st));
        CK(cudaMemcpyAsync(d_probs, h_probs.data(), sizeof(Dev) * count, cudaMemcpyHostToDevice, st));
        CK(cudaStreamSynchronize(st));   // the resident copy must be complete before a run on any other stream
        nLoaded = count;
        return ORB_OK;
    }
    // (re)start from the uploaded initial estimates and run the whole LM loop for every loaded problem
    int run(cudaStream_t s) {
        if (nLoaded < 1) { set_error("lba: nothing uploaded"); return ORB_ERR_ARG; }
        CK(cudaSetDevice(device));
        int maxN = 0;
        for (int i = 0; i < nLoaded; ++i) {
            const Packed& K = packed[i];
            CK(cudaMemcpyAsync(d_arena + K.posesOff, d_arena + K.initPosesOff, 56 * (size_t)K.nP, cudaMemcpyDeviceToDevice, s));
            if (K.nL) CK(cudaMemcpyAsync(d_arena + K.ptsOff, d_arena + K.initPtsOff, 24 * (size_t)K.nL, cudaMemcpyDeviceToDevice, s));
            maxN = std::max(maxN, 6 * K.nF);
        }
        int csize = 1;
        for (int cand = MAXC; cand >= 1; cand >>= 1) if (nLoaded * cand <= numSMs) { csize = cand; break; }
        if (forcedCluster > 0) csize = forcedCluster;
        const int matN = maxN <= smemN ? maxN : 0;   // matrix in shared memory only when every problem of the batch fits
        cudaLaunchConfig_t cfg; memset(&cfg, 0, sizeof(cfg));
        cfg.gridDim = dim3(nLoaded * csize); cfg.blockDim = dim3(NT);
        cfg.dynamicSmemBytes = fixedSmem + 8 * (size_t)matN * matN;
        cfg.stream = s;
        cudaLaunchAttribute attr[1];
        attr[0].id = cudaLaunchAttributeClusterDimension;
        attr[0].val.clusterDim.x = csize; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
        cfg.attrs = attr; cfg.numAttrs = 1;
        const Dev* dp = d_probs; const volatile int* ds = d_stop; int mn = matN, mp = maxP;
        CK(cudaLaunchKernelEx(&cfg, lba_cluster_kernel, dp, ds, mn, mp));
        launches = 1;
        lastCluster = csize;
        return ORB_OK;
    }
    int lastCluster = 1, forcedCluster = 0;
    int download(int count, LbaResult* res, cudaStream_t s) {
        for (int i = 0; i < count; ++i) {
            const Packed& K = packed[i];
            LbaResult& R = res[i];
            if (!R.poses || !R.points || !R.edgeChi2 || !R.edgeDepthPositive) { set_error("lba: null result arrays"); return ORB_ERR_ARG; }
            CK(cudaMemcpyAsync(R.poses, d_arena + K.posesOff, 56 * (size_t)K.nP, cudaMemcpyDeviceToHost, s));
            if (K.nL) CK(cudaMemcpyAsync(R.points, d_arena + K.ptsOff, 24 * (size_t)K.nL, cudaMemcpyDeviceToHost, s));
            if (K.nE) {
                CK(cudaMemcpyAsync(R.edgeChi2, d_arena + K.chi2Off, 8 * (size_t)K.nE, cudaMemcpyDeviceToHost, s));
                CK(cudaMemcpyAsync(R.edgeDepthPositive, d_arena + K.dposOff, (size_t)K.nE, cudaMemcpyDeviceToHost, s));
            }
            CK(cudaMemcpyAsync(h_arena + K.statsOff, d_arena + K.statsOff, 256, cudaMemcpyDeviceToHost, s));
        }
        CK(cudaStreamSynchronize(s));
        for (int i = 0; i < count; ++i) {
            const double* stt = (const double*)(h_arena + packed[i].statsOff);
            res[i].iterations = (int)stt[0]; res[i].trials = (int)stt[1]; res[i].lambda = stt[2]; res[i].chi2 = stt[3]; res[i].initialChi2 = stt[4];
            res[i].gpuLaunches = 1;
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
    if (!h || (ctas != 0 && ctas != 1 && ctas != 2 && ctas != 4 && ctas != 8)) { set_error("lba_set_cluster_size: 0 (auto), 1, 2, 4 or 8"); return ORB_ERR_ARG; }
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
    CK(cudaEventRecord(S.evDone, S.st));
    // forward the callers' stop flags (SparseOptimizer::terminate() polls *pbStopFlag) while the kernel runs
    while (cudaEventQuery(S.evDone) == cudaErrorNotReady)
        for (int i = 0; i < count; ++i)
            if (problems[i].stopFlag && *problems[i].stopFlag) S.h_stop[i] = 1;
    CK(cudaGetLastError());
    return S.download(count, results, S.st);
}
int lba_solve(lba_handle* h, const LbaProblem* problem, LbaResult* result) { return lba_solve_batch(h, 1, problem, result); }

}  // extern "C"
