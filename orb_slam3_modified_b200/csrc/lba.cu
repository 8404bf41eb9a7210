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
#include <string>
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
    const int *ePt, *ePose;           // nE
    const double* obs;                // nE x 2
    const float* invSigma2;           // nE
    const int *ptStart, *ptEdges;     // CSR by point
    const int *poseStart, *poseEdges; // CSR by pose (all poses)
    const int* edgeAt;                // nL x nF: edge of (point, free pose) or -1
    double* err;                      // nE x 2
    double* W;                        // nE x 18 (6x3 Hpl block, zero for fixed poses)
    double* Y;                        // nE x 18 (W Hll^-1)
    double *Hpp, *bp;                 // nF x 36, nF x 6
    double *Hll, *bl;                 // nL x 9, nL x 3
    double *Dinv, *db;                // nL x 9, nL x 3
    double *Hs, *bs;                  // n x n, n
    double* x;                        // n + 3 nL
    double* partial;                  // 4 rotating slots x 16 doubles: per-CTA partial sums + broadcast words
    double* stats;                    // out: [0] iterations [1] trials [2] lambda [3] chi2 [4] initial chi2
    double* outChi2; uint8_t* outDepthPos;   // nE
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

__device__ __forceinline__ void project_edge(const Dev& D, int e, double* Xc, double* uv) {
    const int ic = D.ePose[e];
    const double* T = D.poses + 7 * (size_t)ic;
    double r[3]; qrot(T, D.pts + 3 * (size_t)D.ePt[e], r);
    Xc[0] = r[0] + T[4]; Xc[1] = r[1] + T[5]; Xc[2] = r[2] + T[6];
    const float* c = D.cam + 4 * (size_t)ic;
    uv[0] = (double)c[0] * Xc[0] / Xc[2] + (double)c[2];   // Pinhole::project(Vector3d), float params promoted
    uv[1] = (double)c[1] * Xc[1] / Xc[2] + (double)c[3];
}
__device__ __forceinline__ void robustify(const Dev& D, double e2, double& rho0, double& rho1) {   // RobustKernelHuber
    if (e2 <= D.dsqr) { rho0 = e2; rho1 = 1.; }
    else { const double s = sqrt(e2); rho0 = 2 * s * D.delta - D.dsqr; rho1 = D.delta / s; }
}


namespace cg = cooperative_groups;
constexpr int NT = 256;
constexpr int MAXC = 8;            // largest cluster
constexpr int PSLOT = 16;          // doubles per partial slot

// ordered block reduction of one double per thread; every thread returns the sum
__device__ __forceinline__ double block_sum(double v, double* sm) {
    const int tid = threadIdx.x;
    sm[tid] = v;
    __syncthreads();
#pragma unroll
    for (int s = NT / 2; s > 0; s >>= 1) {
        if (tid < s) sm[tid] += sm[tid + s];
        __syncthreads();
    }
    const double r = sm[0];
    __syncthreads();
    return r;
}

struct Ctx {
    int crank, csize, tid;
    int wid, nw;           // worker id / count over the cluster
    int slot;              // rotating partial slot
    double* sm;            // NT doubles
};

// cluster-wide ordered sum: per-CTA partials -> global slot -> cluster barrier -> every thread adds them in rank order.
// `flagIn` (only meaningful on CTA 0 / thread 0) is broadcast alongside and returned in flagOut.
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

// ---- residuals + robust chi2 (SparseOptimizer::computeActiveErrors + activeRobustChi2) ----
__device__ double phase_errors(const Dev& D, const Ctx& c) {
    double acc = 0;
    for (int e = c.wid; e < D.nE; e += c.nw) {
        double Xc[3], uv[2];
        project_edge(D, e, Xc, uv);
        const double e0 = D.obs[2 * (size_t)e] - uv[0], e1 = D.obs[2 * (size_t)e + 1] - uv[1];
        D.err[2 * (size_t)e] = e0; D.err[2 * (size_t)e + 1] = e1;
        double r0, r1;
        robustify(D, (double)D.invSigma2[e] * (e0 * e0 + e1 * e1), r0, r1);
        acc += r0;
    }
    return acc;
}

// Jacobians of one edge (EdgeSE3ProjectXYZ::linearizeOplus): A = dE/dpoint (2x3), B = dE/dpose (2x6)
__device__ __forceinline__ void edge_jacobians(const Dev& D, int e, double* A, double* B, double& w, double& r0, double& r1) {
    const int ic = D.ePose[e];
    double Xc[3], uv[2];
    project_edge(D, e, Xc, uv);
    const float* c = D.cam + 4 * (size_t)ic;
    const double x = Xc[0], y = Xc[1], z = Xc[2];
    const double fx = (double)c[0], fy = (double)c[1];
    const double J00 = -(fx / z), J02 = fx * x / (z * z), J11 = -(fy / z), J12 = fy * y / (z * z);   // -projectJac
    double R[9]; qtoR(D.poses + 7 * (size_t)ic, R);
#pragma unroll
    for (int k = 0; k < 3; ++k) { A[k] = J00 * R[k] + J02 * R[6 + k]; A[3 + k] = J11 * R[3 + k] + J12 * R[6 + k]; }
    // SE3deriv = [0 z -y 1 0 0; -z 0 x 0 1 0; y -x 0 0 0 1]
    B[0] = J02 * y;            B[1] = J00 * z - J02 * x;  B[2] = -J00 * y;  B[3] = J00; B[4] = 0;   B[5] = J02;
    B[6] = -J11 * z + J12 * y; B[7] = -J12 * x;           B[8] = J11 * x;   B[9] = 0;   B[10] = J11; B[11] = J12;
    const double is2 = (double)D.invSigma2[e];
    const double e0 = D.err[2 * (size_t)e], e1 = D.err[2 * (size_t)e + 1];
    double rho0, rho1;
    robustify(D, is2 * (e0 * e0 + e1 * e1), rho0, rho1);
    w = rho1 * is2;
    r0 = -is2 * e0 * rho1; r1 = -is2 * e1 * rho1;
}

// ---- per point Hll, bl and the Hpl blocks W of its edges; one warp per point, lanes over edges ----
__device__ void phase_build_points(const Dev& D, const Ctx& c) {
    const int lane = c.tid & 31, wpb = NT / 32;
    for (int p = c.crank * wpb + (c.tid >> 5); p < D.nL; p += c.csize * wpb) {
        const int a = D.ptStart[p], b = D.ptStart[p + 1];
        double h[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0}, g[3] = {0, 0, 0};
        for (int k = a + lane; k < b; k += 32) {
            const int e = D.ptEdges[k];
            double A[6], B[12], w, r0, r1;
            edge_jacobians(D, e, A, B, w, r0, r1);
#pragma unroll
            for (int i = 0; i < 3; ++i) {
                g[i] += A[i] * r0 + A[3 + i] * r1;
#pragma unroll
                for (int j = 0; j < 3; ++j) h[i * 3 + j] += w * (A[i] * A[j] + A[3 + i] * A[3 + j]);
            }
            double* We = D.W + 18 * (size_t)e;
            if (D.hidx[D.ePose[e]] >= 0) {
#pragma unroll
                for (int i = 0; i < 6; ++i)
#pragma unroll
                    for (int j = 0; j < 3; ++j) We[i * 3 + j] = w * (B[i] * A[j] + B[6 + i] * A[3 + j]);
            } else {
#pragma unroll
                for (int i = 0; i < 18; ++i) We[i] = 0.0;
            }
        }
#pragma unroll
        for (int o = 16; o; o >>= 1) {
#pragma unroll
            for (int i = 0; i < 9; ++i) h[i] += __shfl_xor_sync(0xffffffffu, h[i], o);
#pragma unroll
            for (int i = 0; i < 3; ++i) g[i] += __shfl_xor_sync(0xffffffffu, g[i], o);
        }
        if (lane == 0) {
#pragma unroll
            for (int i = 0; i < 9; ++i) D.Hll[9 * (size_t)p + i] = h[i];
#pragma unroll
            for (int i = 0; i < 3; ++i) D.bl[3 * (size_t)p + i] = g[i];
        }
    }
}

// ---- per free pose Hpp, bp; one warp per pose, lanes over its edges ----
__device__ void phase_build_poses(const Dev& D, const Ctx& c) {
    const int lane = c.tid & 31, wpb = NT / 32;
    for (int hI = c.crank * wpb + (c.tid >> 5); hI < D.nF; hI += c.csize * wpb) {
        const int ic = D.freePose[hI];
        const int a = D.poseStart[ic], b = D.poseStart[ic + 1];
        double acc[27];
#pragma unroll
        for (int i = 0; i < 27; ++i) acc[i] = 0;
        for (int k = a + lane; k < b; k += 32) {
            const int e = D.poseEdges[k];
            double A[6], B[12], w, r0, r1;
            edge_jacobians(D, e, A, B, w, r0, r1);
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
        if (lane == 0) {
            int t = 0;
            double* H = D.Hpp + 36 * (size_t)hI;
            for (int i = 0; i < 6; ++i)
                for (int j = i; j < 6; ++j) { H[i * 6 + j] = acc[t]; H[j * 6 + i] = acc[t]; ++t; }
            for (int i = 0; i < 6; ++i) D.bp[6 * (size_t)hI + i] = acc[21 + i];
        }
    }
}

// ---- max |diag| over all Hessian blocks (computeLambdaInit); every CTA computes it redundantly ----
__device__ double phase_maxdiag(const Dev& D, const Ctx& c) {
    double m = 0;
    for (int i = c.tid; i < D.n; i += NT) m = fmax(m, fabs(D.Hpp[36 * (size_t)(i / 6) + 7 * (i % 6)]));
    for (int i = c.tid; i < 3 * D.nL; i += NT) m = fmax(m, fabs(D.Hll[9 * (size_t)(i / 3) + 4 * (i % 3)]));
    c.sm[c.tid] = m;
    __syncthreads();
    for (int s = NT / 2; s > 0; s >>= 1) { if (c.tid < s) c.sm[c.tid] = fmax(c.sm[c.tid], c.sm[c.tid + s]); __syncthreads(); }
    const double r = c.sm[0];
    __syncthreads();
    return r;
}

// ---- Hll^-1 (lambda on the diagonal) and Hll^-1 bl per point (block_solver.hpp:381-394) ----
__device__ void phase_point_prep(const Dev& D, const Ctx& c, double lambda) {
    for (int p = c.wid; p < D.nL; p += c.nw) {
        double m[9];
#pragma unroll
        for (int i = 0; i < 9; ++i) m[i] = D.Hll[9 * (size_t)p + i];
        m[0] += lambda; m[4] += lambda; m[8] += lambda;
        const double c00 = m[4] * m[8] - m[5] * m[7], c10 = m[5] * m[6] - m[3] * m[8], c20 = m[3] * m[7] - m[4] * m[6];
        const double id = 1.0 / (m[0] * c00 + m[1] * c10 + m[2] * c20);
        double o[9];
        o[0] = c00 * id; o[1] = (m[2] * m[7] - m[1] * m[8]) * id; o[2] = (m[1] * m[5] - m[2] * m[4]) * id;
        o[3] = c10 * id; o[4] = (m[0] * m[8] - m[2] * m[6]) * id; o[5] = (m[2] * m[3] - m[0] * m[5]) * id;
        o[6] = c20 * id; o[7] = (m[1] * m[6] - m[0] * m[7]) * id; o[8] = (m[0] * m[4] - m[1] * m[3]) * id;
#pragma unroll
        for (int i = 0; i < 9; ++i) D.Dinv[9 * (size_t)p + i] = o[i];
        const double* b3 = D.bl + 3 * (size_t)p;
#pragma unroll
        for (int i = 0; i < 3; ++i) D.db[3 * (size_t)p + i] = o[i * 3] * b3[0] + o[i * 3 + 1] * b3[1] + o[i * 3 + 2] * b3[2];
        // Y = W Hll^-1 for the edges of this point
        for (int k = D.ptStart[p]; k < D.ptStart[p + 1]; ++k) {
            const int e = D.ptEdges[k];
            const double* Wd = D.W + 18 * (size_t)e;
            double* Yd = D.Y + 18 * (size_t)e;
#pragma unroll
            for (int a = 0; a < 6; ++a)
#pragma unroll
                for (int b = 0; b < 3; ++b) Yd[a * 3 + b] = Wd[a * 3] * o[b] + Wd[a * 3 + 1] * o[3 + b] + Wd[a * 3 + 2] * o[6 + b];
        }
    }
}
// ---- Schur complement, one warp per block pair (i1 <= i2) (block_solver.hpp:396-431) ----
__device__ void phase_schur(const Dev& D, const Ctx& c, double lambda, double* Hs, int ld) {
    const int lane = c.tid & 31, wpb = NT / 32;
    const int nPairs = D.nF * (D.nF + 1) / 2;
    for (int pr = c.crank * wpb + (c.tid >> 5); pr < nPairs; pr += c.csize * wpb) {
        int i1 = 0, rem = pr;   // unrank (i1, i2), i1 <= i2, row-major over the upper triangle
        while (rem >= D.nF - i1) { rem -= D.nF - i1; ++i1; }
        const int i2 = i1 + rem;
        const int ic = D.freePose[i1];
        const int a = D.poseStart[ic], b = D.poseStart[ic + 1];
        double acc[36];
#pragma unroll
        for (int i = 0; i < 36; ++i) acc[i] = 0;
        double bacc[6] = {0, 0, 0, 0, 0, 0};
        for (int k = a + lane; k < b; k += 32) {
            const int e1 = D.poseEdges[k];
            const int p = D.ePt[e1];
            const int e2 = (i1 == i2) ? e1 : D.edgeAt[(size_t)p * D.nF + i2];
            if (e2 < 0) continue;
            const double* Y1 = D.Y + 18 * (size_t)e1;
            const double* W2 = D.W + 18 * (size_t)e2;
#pragma unroll
            for (int r = 0; r < 6; ++r)
#pragma unroll
                for (int q = 0; q < 6; ++q) acc[r * 6 + q] += Y1[r * 3] * W2[q * 3] + Y1[r * 3 + 1] * W2[q * 3 + 1] + Y1[r * 3 + 2] * W2[q * 3 + 2];
            if (i1 == i2) {
                const double* dbp = D.db + 3 * (size_t)p;
#pragma unroll
                for (int r = 0; r < 6; ++r) bacc[r] += W2[r * 3] * dbp[0] + W2[r * 3 + 1] * dbp[1] + W2[r * 3 + 2] * dbp[2];
            }
        }
#pragma unroll
        for (int o = 16; o; o >>= 1) {
#pragma unroll
            for (int i = 0; i < 36; ++i) acc[i] += __shfl_xor_sync(0xffffffffu, acc[i], o);
#pragma unroll
            for (int i = 0; i < 6; ++i) bacc[i] += __shfl_xor_sync(0xffffffffu, bacc[i], o);
        }
        // lanes 0..35 each write one entry (+ mirror); all lanes hold the reduced values
#pragma unroll
        for (int i = 0; i < 36; ++i) {
            if (lane == (i & 31)) {
                const int r = i / 6, q = i % 6;
                double v = -acc[i];
                if (i1 == i2) {
                    v += D.Hpp[36 * (size_t)i1 + i] + (r == q ? lambda : 0.0);
                    if (r >= q) Hs[(size_t)(6 * i1 + r) * ld + 6 * i1 + q] = v;       // lower triangle is what ldlt reads
                } else {
                    Hs[(size_t)(6 * i2 + q) * ld + 6 * i1 + r] = v;                    // block (i2, i1), lower triangle
                }
            }
        }
        if (i1 == i2 && lane < 6) D.bs[6 * i1 + lane] = D.bp[6 * (size_t)i1 + lane] - bacc[lane];
    }
}
// ---- dense LDL^T (no pivoting) + solve, one CTA; A is the lower triangle with leading dimension ld ----
// (LinearSolverEigen::solve: SimplicialLDLT fails only on an exactly zero pivot.)  Returns 1 on success (uniform).
__device__ int phase_ldlt(const Dev& D, const Ctx& c, double* A, int ld) {
    const int n = D.n, tid = c.tid;
    __shared__ int s_ok;
    if (tid == 0) s_ok = 1;
    __syncthreads();
    for (int k = 0; k < n; ++k) {
        const double d = A[(size_t)k * ld + k];
        if (d == 0.0) { if (tid == 0) s_ok = 0; break; }      // uniform: every thread reads the same d
        const int m = n - k - 1;
        // trailing update of the lower triangle with the unscaled column k
        for (int idx = tid; idx < m * m; idx += NT) {
            const int i = k + 1 + idx / m, j = k + 1 + idx % m;
            if (j <= i) A[(size_t)i * ld + j] -= A[(size_t)i * ld + k] * A[(size_t)j * ld + k] / d;
        }
        __syncthreads();
        for (int i = k + 1 + tid; i < n; i += NT) A[(size_t)i * ld + k] /= d;
        __syncthreads();
    }
    __syncthreads();
    const int ok = s_ok;
    __syncthreads();
    if (!ok) return 0;
    // solve L D L^T x = bs; y kept in D.x (pose part), column-oriented substitutions
    double* y = D.x;
    for (int i = tid; i < n; i += NT) y[i] = D.bs[i];
    __syncthreads();
    for (int k = 0; k < n; ++k) {          // forward: y_i -= L_ik y_k
        const double yk = y[k];
        for (int i = k + 1 + tid; i < n; i += NT) y[i] -= A[(size_t)i * ld + k] * yk;
        __syncthreads();
    }
    for (int i = tid; i < n; i += NT) y[i] /= A[(size_t)i * ld + i];
    __syncthreads();
    for (int k = n - 1; k >= 0; --k) {     // backward: y_j -= L_kj y_k for j < k
        const double yk = y[k];
        for (int j = tid; j < k; j += NT) y[j] -= A[(size_t)k * ld + j] * yk;
        __syncthreads();
    }
    return 1;
}
// ---- landmark back-substitution x_l = Hll^-1 (bl - W^T x_p)  (block_solver.hpp:461-483) ----
__device__ void phase_backsub(const Dev& D, const Ctx& c) {
    for (int p = c.wid; p < D.nL; p += c.nw) {
        double cl[3] = {D.bl[3 * (size_t)p], D.bl[3 * (size_t)p + 1], D.bl[3 * (size_t)p + 2]};
        for (int k = D.ptStart[p]; k < D.ptStart[p + 1]; ++k) {
            const int e = D.ptEdges[k];
            const int h = D.hidx[D.ePose[e]];
            if (h < 0) continue;
            const double* Wd = D.W + 18 * (size_t)e;
            const double* xp = D.x + 6 * (size_t)h;
#pragma unroll
            for (int b = 0; b < 3; ++b)
#pragma unroll
                for (int a = 0; a < 6; ++a) cl[b] -= Wd[a * 3 + b] * xp[a];
        }
        const double* Di = D.Dinv + 9 * (size_t)p;
#pragma unroll
        for (int a = 0; a < 3; ++a) D.x[D.n + 3 * (size_t)p + a] = Di[a * 3] * cl[0] + Di[a * 3 + 1] * cl[1] + Di[a * 3 + 2] * cl[2];
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
        D.outChi2[e] = (double)D.invSigma2[e] * (e0 * e0 + e1 * e1);   // e->chi2() from the last computed _error (Optimizer.cc:1425)
        double Xc[3], uv[2];
        project_edge(D, e, Xc, uv);
        D.outDepthPos[e] = Xc[2] > 0.0;
    }
}

// =============================================================================================
// The persistent kernel: SparseOptimizer::optimize (sparse_optimizer.cpp:354-418) around
// OptimizationAlgorithmLevenberg::solve (optimization_algorithm_levenberg.cpp:61-168), one cluster per problem.
// Every thread carries the (uniform) LM state; decisions use values every CTA reads identically after a cluster barrier.
// =============================================================================================
__global__ void __launch_bounds__(NT, 1) lba_cluster_kernel(const Dev* __restrict__ probs, const volatile int* stop, int smemMatrixN) {
    extern __shared__ double s_mat[];          // reduced camera system (CTA 0) when it fits: smemMatrixN^2 doubles
    __shared__ double s_red[NT];
    cg::cluster_group cluster = cg::this_cluster();
    Ctx c;
    c.crank = (int)cluster.block_rank(); c.csize = (int)cluster.num_blocks(); c.tid = threadIdx.x;
    c.wid = c.crank * NT + c.tid; c.nw = c.csize * NT; c.slot = 0; c.sm = s_red;
    const Dev D = probs[blockIdx.x / c.csize];
    const bool matInSmem = D.n <= smemMatrixN;
    // reduced camera system: CTA 0's shared memory, reached from the other CTAs of the cluster as distributed shared memory
    double* Hs = matInSmem ? cluster.map_shared_rank(s_mat, 0) : D.Hs;
    const int ld = D.n;
    const int stopIdx = blockIdx.x / c.csize;

    for (int i = c.wid; i < D.nP; i += c.nw) qnormalize(D.poses + 7 * (size_t)i);   // SE3Quat(q, t) constructor
    for (int i = c.wid; i < 2 * D.nE; i += c.nw) D.err[i] = 0.0;
    for (int i = c.wid; i < D.n + 3 * D.nL; i += c.nw) D.x[i] = 0.0;
    int term = 0, dummy;
    // initial terminate() poll (+ makes the normalised poses visible cluster-wide)
    cluster_sum(D, c, 0.0, (stop && stop[stopIdx]) ? 1 : 0, term);

    double lambda = -1, ni = 2, currentChi = 0, firstChi = 0;
    int nBad = 0, cj = 0, trials = 0;
    const int maxTrials = 10;
    const double goodUpper = 2. / 3., goodLower = 1. / 3., tau = 1e-5;
    bool ok = true;
    for (int it = 0; it < D.iterations && !term && ok; ++it) {
        currentChi = cluster_sum(D, c, phase_errors(D, c), 0, dummy);
        double tempChi = currentChi;
        const double iniChi = currentChi;
        if (it == 0) firstChi = iniChi;
        phase_build_points(D, c);
        phase_build_poses(D, c);
        csync();
        if (it == 0) {
            if (D.userLambdaInit > 0) lambda = D.userLambdaInit;
            else lambda = tau * phase_maxdiag(D, c);
            ni = 2; nBad = 0;
        }
        double rho = 0;
        int qmax = 0;
        do {
            phase_point_prep(D, c, lambda);
            csync();
            int ok2 = 1;
            if (D.nF) {
                phase_schur(D, c, lambda, Hs, ld);   // every CTA writes its block pairs (into CTA 0's shared memory through DSMEM)
                csync();
                if (c.crank == 0) {
                    ok2 = phase_ldlt(D, c, Hs, ld);
                    if (c.tid == 0) D.partial[4 * PSLOT] = (double)ok2;
                }
                csync();
                ok2 = (int)D.partial[4 * PSLOT];
            }
            phase_backsub(D, c);
            csync();
            const double scale0 = cluster_sum(D, c, phase_update(D, c, lambda), 0, dummy);
            tempChi = cluster_sum(D, c, phase_errors(D, c), (stop && stop[stopIdx]) ? 1 : 0, term);
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
        b += al(4 * (nL + 1)) + al(4 * nE) + al(4 * (nP + 1)) + al(4 * nE) + al(4 * nL * nP);   // CSR + edgeAt
        b += al(16 * nE) + 2 * al(144 * nE);                                      // err, W, Y
        b += al(288 * nP) + al(48 * nP) + 2 * al(72 * nL) + 2 * al(24 * nL);      // Hpp, bp, Hll, Dinv, bl, db
        b += al(8 * n * n) + al(8 * n) + al(8 * (n + 3 * nL)) + al(8 * (4 * PSLOT + 8)) + al(64);   // Hs, bs, x, partial, stats
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
        // shared-memory budget for the reduced camera system: up to 200 KB of the 227 KB opt-in limit
        const int maxDyn = 200 * 1024;
        smemN = (int)floor(sqrt((double)maxDyn / 8.0));
        smemN = std::min(smemN, 6 * maxP);
        CK(cudaFuncSetAttribute(lba_cluster_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)std::max<size_t>(8 * (size_t)smemN * smemN, 1024)));
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
        std::vector<int> ptStart(nL + 1, 0), poseStart(nP + 1, 0), ptEdges(nE), poseEdges(nE);
        for (int e = 0; e < nE; ++e) {
            const int p = P->edgePoint[e], c = P->edgePose[e];
            if (p < 0 || p >= nL || c < 0 || c >= nP) { set_error("lba: edge index out of range"); return ORB_ERR_ARG; }
            ++ptStart[p + 1]; ++poseStart[c + 1];
        }
        for (int i = 0; i < nL; ++i) ptStart[i + 1] += ptStart[i];
        for (int i = 0; i < nP; ++i) poseStart[i + 1] += poseStart[i];
        {
            std::vector<int> a(ptStart.begin(), ptStart.end() - 1), b(poseStart.begin(), poseStart.end() - 1);
            for (int e = 0; e < nE; ++e) { ptEdges[a[P->edgePoint[e]]++] = e; poseEdges[b[P->edgePose[e]]++] = e; }
        }
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
        o = carve(4 * (size_t)nE); put(o, P->edgePoint, 4 * (size_t)nE); D.ePt = (const int*)(d_arena + o);
        o = carve(4 * (size_t)nE); put(o, P->edgePose, 4 * (size_t)nE); D.ePose = (const int*)(d_arena + o);
        o = carve(16 * (size_t)nE); put(o, P->obs, 16 * (size_t)nE); D.obs = (const double*)(d_arena + o);
        o = carve(4 * (size_t)nE); put(o, P->invSigma2, 4 * (size_t)nE); D.invSigma2 = (const float*)(d_arena + o);
        o = carve(4 * (size_t)(nL + 1)); put(o, ptStart.data(), 4 * (size_t)(nL + 1)); D.ptStart = (const int*)(d_arena + o);
        o = carve(4 * (size_t)nE); put(o, ptEdges.data(), 4 * (size_t)nE); D.ptEdges = (const int*)(d_arena + o);
        o = carve(4 * (size_t)(nP + 1)); put(o, poseStart.data(), 4 * (size_t)(nP + 1)); D.poseStart = (const int*)(d_arena + o);
        o = carve(4 * (size_t)nE); put(o, poseEdges.data(), 4 * (size_t)nE); D.poseEdges = (const int*)(d_arena + o);
        o = carve(4 * (size_t)nL * std::max(nF, 1));
        {
            int* ea = (int*)(h_arena + o);
            const size_t cnt = (size_t)nL * std::max(nF, 1);
            for (size_t i = 0; i < cnt; ++i) ea[i] = -1;
            for (int e = 0; e < nE; ++e) {
                const int hI = hidx[P->edgePose[e]];
                if (hI >= 0) {
                    int& s = ea[(size_t)P->edgePoint[e] * nF + hI];
                    if (s >= 0) { set_error("lba: duplicate (point, keyframe) observation"); return ORB_ERR_ARG; }
                    s = e;
                }
            }
            D.edgeAt = (const int*)(d_arena + o);
        }
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
        D.Hs = (double*)(d_arena + carve(8 * (size_t)n * n)); D.bs = (double*)(d_arena + carve(8 * (size_t)std::max(n, 1)));
        D.x = (double*)(d_arena + carve(8 * ((size_t)n + 3 * (size_t)nL)));
        D.partial = (double*)(d_arena + carve(8 * (4 * PSLOT + 8)));
        K.statsOff = carve(64); D.stats = (double*)(d_arena + K.statsOff);
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
        for (int i = 0; i < count; ++i) { int rc = pack(i, probs + i); if (rc) return rc; }
        CK(cudaSetDevice(device));
        for (int i = 0; i < count; ++i)
            CK(cudaMemcpyAsync(d_arena + perProblem * (size_t)i, h_arena + perProblem * (size_t)i, uploadBytes[i], cudaMemcpyHostToDevice, st));
        CK(cudaMemcpyAsync(d_probs, h_probs.data(), sizeof(Dev) * count, cudaMemcpyHostToDevice, st));
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
        const int matN = maxN <= smemN ? maxN : 0;   // matrix in shared memory only when every problem of the batch fits
        cudaLaunchConfig_t cfg; memset(&cfg, 0, sizeof(cfg));
        cfg.gridDim = dim3(nLoaded * csize); cfg.blockDim = dim3(NT);
        cfg.dynamicSmemBytes = std::max<size_t>(8 * (size_t)matN * matN, 16);
        cfg.stream = s;
        cudaLaunchAttribute attr[1];
        attr[0].id = cudaLaunchAttributeClusterDimension;
        attr[0].val.clusterDim.x = csize; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
        cfg.attrs = attr; cfg.numAttrs = 1;
        const Dev* dp = d_probs; const volatile int* ds = d_stop; int mn = matN;
        CK(cudaLaunchKernelEx(&cfg, lba_cluster_kernel, dp, ds, mn));
        launches = 1;
        lastCluster = csize;
        return ORB_OK;
    }
    int lastCluster = 1;
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
            CK(cudaMemcpyAsync(h_arena + K.statsOff, d_arena + K.statsOff, 64, cudaMemcpyDeviceToHost, s));
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
