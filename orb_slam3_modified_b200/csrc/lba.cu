// B200 kernels + C-ABI for the numeric core of Optimizer::LocalBundleAdjustment
// (reference src/Optimizer.cc:1116-1498; g2o BlockSolver_6_3 + Levenberg, see include/orb_b200.h and SURVEY.md 3.2).
//
// All arithmetic is FP64 (the reference's g2o types are double).  Phases of one LM trial:
//   residual   thread/edge      EdgeSE3ProjectXYZ::computeError + Huber rho            (HOT LOOP A)
//   build      warp/point, CTA/pose   linearizeOplus + constructQuadraticForm: Hll, bl, W=Hpl blocks, Hpp, bp   (HOT LOOP B)
//   schur      thread/point, thread/edge, CTA/pose-pair   Hll^-1, Y = W Hll^-1, Hschur = Hpp - sum Y W^T        (HOT LOOP C)
//   ldlt       one CTA          dense LDL^T of the reduced camera system + solve
//   backsub    warp/point       x_l = Hll^-1 (b_l - W^T x_p)
//   update     thread/vertex    T <- exp(dx) T, p <- p + dx  (+ backup for the LM "pop")
// Every reduction is ordered (no floating-point atomics), so results are reproducible run to run.
// Each phase is a __device__ function over a grid-stride / block-stride range so that the same code serves the
// kernel-per-phase driver below and a persistent cooperative kernel.
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
    double* partial;                  // reduction scratch (>= max grid size * 2)
    double* scal;                     // [0] chi, [1] scale, [2] maxDiag, [3] ldlt ok (1/0)
    double delta, dsqr;
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

// ordered block reduction of one double per thread; result valid in thread 0
template <int NT>
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

constexpr int NT = 256;

// ---- phase: residuals + robust chi2 partial sums (SparseOptimizer::computeActiveErrors + activeRobustChi2) ----
__global__ void __launch_bounds__(NT) k_errors(Dev D) {
    __shared__ double sm[NT];
    double acc = 0;
    for (int e = blockIdx.x * NT + threadIdx.x; e < D.nE; e += gridDim.x * NT) {
        double Xc[3], uv[2];
        project_edge(D, e, Xc, uv);
        const double e0 = D.obs[2 * (size_t)e] - uv[0], e1 = D.obs[2 * (size_t)e + 1] - uv[1];
        D.err[2 * (size_t)e] = e0; D.err[2 * (size_t)e + 1] = e1;
        double r0, r1;
        robustify(D, (double)D.invSigma2[e] * (e0 * e0 + e1 * e1), r0, r1);
        acc += r0;
    }
    const double s = block_sum<NT>(acc, sm);
    if (threadIdx.x == 0) D.partial[blockIdx.x] = s;
}
// final ordered sum of `count` partials into scal[slot]
__global__ void __launch_bounds__(NT) k_reduce(Dev D, int count, int slot, int offset) {
    __shared__ double sm[NT];
    double acc = 0;
    for (int i = threadIdx.x; i < count; i += NT) acc += D.partial[offset + i];
    const double s = block_sum<NT>(acc, sm);
    if (threadIdx.x == 0) D.scal[slot] = s;
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

// ---- phase: per point Hll, bl and the Hpl blocks W of its edges; one warp per point, lanes over edges ----
__global__ void __launch_bounds__(NT) k_build_points(Dev D) {
    const int lane = threadIdx.x & 31;
    const int wpb = NT / 32;
    for (int p = blockIdx.x * wpb + (threadIdx.x >> 5); p < D.nL; p += gridDim.x * wpb) {
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

// ---- phase: per free pose Hpp, bp; one CTA per pose, threads over its edges ----
__global__ void __launch_bounds__(NT) k_build_poses(Dev D) {
    __shared__ double sm[NT];
    for (int hI = blockIdx.x; hI < D.nF; hI += gridDim.x) {
        const int ic = D.freePose[hI];
        const int a = D.poseStart[ic], b = D.poseStart[ic + 1];
        double acc[27];
#pragma unroll
        for (int i = 0; i < 27; ++i) acc[i] = 0;
        for (int k = a + threadIdx.x; k < b; k += NT) {
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
        double red[27];
#pragma unroll
        for (int i = 0; i < 27; ++i) red[i] = block_sum<NT>(acc[i], sm);
        if (threadIdx.x == 0) {
            int t = 0;
            double* H = D.Hpp + 36 * (size_t)hI;
            for (int i = 0; i < 6; ++i)
                for (int j = i; j < 6; ++j) { H[i * 6 + j] = red[t]; H[j * 6 + i] = red[t]; ++t; }
            for (int i = 0; i < 6; ++i) D.bp[6 * (size_t)hI + i] = red[21 + i];
        }
    }
}

// ---- phase: max |diag| over all Hessian blocks (computeLambdaInit) ----
__global__ void __launch_bounds__(NT) k_maxdiag(Dev D) {
    __shared__ double sm[NT];
    double m = 0;
    for (int i = threadIdx.x; i < D.n; i += NT) m = fmax(m, fabs(D.Hpp[36 * (size_t)(i / 6) + 7 * (i % 6)]));
    for (int i = threadIdx.x; i < 3 * D.nL; i += NT) m = fmax(m, fabs(D.Hll[9 * (size_t)(i / 3) + 4 * (i % 3)]));
    sm[threadIdx.x] = m;
    __syncthreads();
    for (int s = NT / 2; s > 0; s >>= 1) { if (threadIdx.x < s) sm[threadIdx.x] = fmax(sm[threadIdx.x], sm[threadIdx.x + s]); __syncthreads(); }
    if (threadIdx.x == 0) D.scal[2] = sm[0];
}

// ---- phase: Hll^-1 (with lambda on the diagonal) and Hll^-1 bl per point (block_solver.hpp:381-394) ----
__global__ void __launch_bounds__(NT) k_point_prep(Dev D, double lambda) {
    for (int p = blockIdx.x * NT + threadIdx.x; p < D.nL; p += gridDim.x * NT) {
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
    }
}
// ---- phase: Y = W Hll^-1 per edge ----
__global__ void __launch_bounds__(NT) k_edge_y(Dev D) {
    for (int e = blockIdx.x * NT + threadIdx.x; e < D.nE; e += gridDim.x * NT) {
        const double* Wd = D.W + 18 * (size_t)e;
        const double* Di = D.Dinv + 9 * (size_t)D.ePt[e];
        double* Yd = D.Y + 18 * (size_t)e;
#pragma unroll
        for (int a = 0; a < 6; ++a)
#pragma unroll
            for (int b = 0; b < 3; ++b) Yd[a * 3 + b] = Wd[a * 3] * Di[b] + Wd[a * 3 + 1] * Di[3 + b] + Wd[a * 3 + 2] * Di[6 + b];
    }
}
// ---- phase: Schur complement, one CTA per block pair (i1 <= i2) (block_solver.hpp:396-431) ----
__global__ void __launch_bounds__(NT) k_schur(Dev D, double lambda) {
    __shared__ double sm[NT];
    const int nPairs = D.nF * (D.nF + 1) / 2;
    for (int pr = blockIdx.x; pr < nPairs; pr += gridDim.x) {
        // unrank (i1, i2), i1 <= i2, row-major over the upper triangle
        int i1 = 0, rem = pr;
        while (rem >= D.nF - i1) { rem -= D.nF - i1; ++i1; }
        const int i2 = i1 + rem;
        const int ic = D.freePose[i1];
        const int a = D.poseStart[ic], b = D.poseStart[ic + 1];
        double acc[36];
#pragma unroll
        for (int i = 0; i < 36; ++i) acc[i] = 0;
        double bacc[6] = {0, 0, 0, 0, 0, 0};
        for (int k = a + threadIdx.x; k < b; k += NT) {
            const int e1 = D.poseEdges[k];
            const int p = D.ePt[e1];
            const int e2 = (i1 == i2) ? e1 : D.edgeAt[(size_t)p * D.nF + i2];
            if (e2 < 0) continue;
            const double* Y1 = D.Y + 18 * (size_t)e1;
            const double* W2 = D.W + 18 * (size_t)e2;
#pragma unroll
            for (int r = 0; r < 6; ++r)
#pragma unroll
                for (int c = 0; c < 6; ++c) acc[r * 6 + c] += Y1[r * 3] * W2[c * 3] + Y1[r * 3 + 1] * W2[c * 3 + 1] + Y1[r * 3 + 2] * W2[c * 3 + 2];
            if (i1 == i2) {
                const double* W1 = D.W + 18 * (size_t)e1;
                const double* dbp = D.db + 3 * (size_t)p;
#pragma unroll
                for (int r = 0; r < 6; ++r) bacc[r] += W1[r * 3] * dbp[0] + W1[r * 3 + 1] * dbp[1] + W1[r * 3 + 2] * dbp[2];
            }
        }
        for (int i = 0; i < 36; ++i) {
            const double s = block_sum<NT>(acc[i], sm);
            if (threadIdx.x == 0) {
                const int r = i / 6, c = i % 6;
                double v = -s;
                if (i1 == i2) v += D.Hpp[36 * (size_t)i1 + i] + (r == c ? lambda : 0.0);
                D.Hs[(size_t)(6 * i1 + r) * D.n + 6 * i2 + c] = v;
                D.Hs[(size_t)(6 * i2 + c) * D.n + 6 * i1 + r] = v;
            }
        }
        if (i1 == i2) {
            for (int r = 0; r < 6; ++r) {
                const double s = block_sum<NT>(bacc[r], sm);
                if (threadIdx.x == 0) D.bs[6 * i1 + r] = D.bp[6 * (size_t)i1 + r] - s;
            }
        }
    }
}
// ---- phase: dense LDL^T (no pivoting) of the reduced camera system + solve; one CTA, matrix in global/L2 ----
// (LinearSolverEigen::solve: SimplicialLDLT fails only on an exactly zero pivot.)
__global__ void __launch_bounds__(NT) k_ldlt(Dev D) {
    const int n = D.n, tid = threadIdx.x;
    double* A = D.Hs;
    __shared__ int s_ok;
    __shared__ double s_d;
    if (tid == 0) s_ok = 1;
    __syncthreads();
    // right-looking: after step k, column k holds L(:,k), diagonal holds d_k, trailing block updated
    for (int k = 0; k < n; ++k) {
        if (tid == 0) { s_d = A[(size_t)k * n + k]; if (s_d == 0.0) s_ok = 0; }
        __syncthreads();
        if (!s_ok) break;
        const double d = s_d;
        // l_i = A[i][k] / d ; trailing A[i][j] -= l_i * A[j][k]   (j <= i, lower triangle), A[j][k] still unscaled
        const int m = n - k - 1;
        for (int idx = tid; idx < m * m; idx += NT) {
            const int i = k + 1 + idx / m, j = k + 1 + idx % m;
            if (j <= i) A[(size_t)i * n + j] -= A[(size_t)i * n + k] * A[(size_t)j * n + k] / d;
        }
        __syncthreads();
        for (int i = k + 1 + tid; i < n; i += NT) A[(size_t)i * n + k] /= d;
        __syncthreads();
    }
    if (tid == 0) D.scal[3] = s_ok ? 1.0 : 0.0;
    if (!s_ok) return;
    // forward / diagonal / backward substitution by one warp-free loop (n is small)
    double* y = D.x;   // pose part of x
    if (tid == 0) {
        for (int i = 0; i < n; ++i) { double s = D.bs[i]; for (int j = 0; j < i; ++j) s -= A[(size_t)i * n + j] * y[j]; y[i] = s; }
        for (int i = 0; i < n; ++i) y[i] /= A[(size_t)i * n + i];
        for (int i = n - 1; i >= 0; --i) { double s = y[i]; for (int j = i + 1; j < n; ++j) s -= A[(size_t)j * n + i] * y[j]; y[i] = s; }
    }
}
// ---- phase: landmark back-substitution x_l = Hll^-1 (bl - W^T x_p)  (block_solver.hpp:461-483) ----
__global__ void __launch_bounds__(NT) k_backsub(Dev D) {
    for (int p = blockIdx.x * NT + threadIdx.x; p < D.nL; p += gridDim.x * NT) {
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
// ---- phase: push + update (SparseOptimizer::push / update) and the partial sums of computeScale ----
__global__ void __launch_bounds__(NT) k_update(Dev D, double lambda) {
    __shared__ double sm[NT];
    double acc = 0;
    const int total = D.nP + D.nL;
    for (int v = blockIdx.x * NT + threadIdx.x; v < total; v += gridDim.x * NT) {
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
    const double s = block_sum<NT>(acc, sm);
    if (threadIdx.x == 0) D.partial[4096 + blockIdx.x] = s;
}
// ---- phase: pop (restore the backup) ----
__global__ void __launch_bounds__(NT) k_restore(Dev D) {
    for (int i = blockIdx.x * NT + threadIdx.x; i < 7 * D.nP; i += gridDim.x * NT) D.poses[i] = D.posesBk[i];
    for (int i = blockIdx.x * NT + threadIdx.x; i < 3 * D.nL; i += gridDim.x * NT) D.pts[i] = D.ptsBk[i];
}
// ---- final per-edge outputs: chi2 from the last computed errors, depth sign at the final state ----
__global__ void __launch_bounds__(NT) k_finalize(Dev D, double* chi2, uint8_t* depthPos) {
    for (int e = blockIdx.x * NT + threadIdx.x; e < D.nE; e += gridDim.x * NT) {
        const double e0 = D.err[2 * (size_t)e], e1 = D.err[2 * (size_t)e + 1];
        chi2[e] = (double)D.invSigma2[e] * (e0 * e0 + e1 * e1);
        double Xc[3], uv[2];
        project_edge(D, e, Xc, uv);
        depthPos[e] = Xc[2] > 0.0;
    }
}
__global__ void k_normalize_poses(Dev D) {   // SE3Quat(q, t) constructor: normalizeRotation
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < D.nP) qnormalize(D.poses + 7 * (size_t)i);
}

struct Solver {
    int device, maxP, maxL, maxE;
    cudaStream_t st = nullptr;
    uint8_t* d_arena = nullptr; size_t arenaBytes = 0;
    double* h_scal = nullptr;   // pinned
    int launches = 0;
    ~Solver() {
        cudaSetDevice(device);
        if (d_arena) cudaFree(d_arena);
        if (h_scal) cudaFreeHost(h_scal);
        if (st) cudaStreamDestroy(st);
    }
    static size_t need(size_t nP, size_t nL, size_t nE) {
        const size_t n = 6 * nP;
        size_t b = 0;
        auto add = [&](size_t bytes) { b += (bytes + 255) & ~(size_t)255; };
        add(56 * nP); add(56 * nP); add(24 * nL); add(24 * nL); add(16 * nP); add(4 * nP); add(4 * nP);
        add(4 * nE); add(4 * nE); add(16 * nE); add(4 * nE);
        add(4 * (nL + 1)); add(4 * nE); add(4 * (nP + 1)); add(4 * nE); add(4 * nL * nP);
        add(16 * nE); add(144 * nE); add(144 * nE);
        add(288 * nP); add(48 * nP); add(72 * nL); add(24 * nL); add(72 * nL); add(24 * nL);
        add(8 * n * n); add(8 * n); add(8 * (n + 3 * nL)); add(8 * 8192); add(64);
        add(8 * nE); add(nE);
        return b + 4096;
    }
    int init() {
        CK(cudaSetDevice(device));
        cudaDeviceProp prop;
        CK(cudaGetDeviceProperties(&prop, device));
        if (prop.major < 10) { set_error("device is not sm_100+ (Blackwell); this library has no other code path"); return ORB_ERR_CUDA; }
        arenaBytes = need(maxP, maxL, maxE);
        CK(cudaMalloc(&d_arena, arenaBytes));
        CK(cudaMallocHost(&h_scal, 64));
        CK(cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking));
        return ORB_OK;
    }
};

}  // namespace lba

using namespace lba;

struct lba_handle { Solver s; };

extern "C" {

int lba_create(lba_handle** out, int max_poses, int max_points, int max_edges, int device) {
    if (!out || max_poses < 1 || max_points < 1 || max_edges < 1) { set_error("lba_create: bad argument"); return ORB_ERR_ARG; }
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) { set_error("no CUDA device (this library has no CPU path)"); return ORB_ERR_CUDA; }
    if (device < 0 || device >= ndev) { set_error("lba_create: bad device index"); return ORB_ERR_ARG; }
    lba_handle* h = new lba_handle();
    h->s.device = device; h->s.maxP = max_poses; h->s.maxL = max_points; h->s.maxE = max_edges;
    int rc = h->s.init();
    if (rc) { delete h; return rc; }
    *out = h;
    return ORB_OK;
}
void lba_destroy(lba_handle* h) { delete h; }

int lba_solve(lba_handle* h, const LbaProblem* P, LbaResult* R) {
    if (!h || !P || !R || !R->poses || !R->points || !R->edgeChi2 || !R->edgeDepthPositive) { set_error("lba_solve: bad argument"); return ORB_ERR_ARG; }
    Solver& S = h->s;
    const int nP = P->nPoses, nL = P->nPoints, nE = P->nEdges;
    if (nP < 1 || nL < 0 || nE < 0 || nP > S.maxP || nL > S.maxL || nE > S.maxE || !P->poses || !P->poseFixed || !P->cam ||
        (nL && !P->points) || (nE && (!P->edgePoint || !P->edgePose || !P->obs || !P->invSigma2))) {
        set_error("lba_solve: problem larger than the handle or null arrays"); return ORB_ERR_ARG;
    }
    CK(cudaSetDevice(S.device));
    // ---- host-side structure (BlockSolver::buildStructure, block_solver.hpp:143-295): index maps and CSR lists ----
    std::vector<int> hidx(nP, -1), freePose;
    for (int i = 0; i < nP; ++i) if (!P->poseFixed[i]) { hidx[i] = (int)freePose.size(); freePose.push_back(i); }
    const int nF = (int)freePose.size(), n = 6 * nF;
    R->iterations = 0; R->trials = 0; R->lambda = -1; R->chi2 = 0; R->initialChi2 = 0; R->gpuLaunches = 0;
    if (nF + nL == 0) { set_error("lba_solve: 0 vertices to optimize"); return ORB_ERR_ARG; }
    std::vector<int> ptStart(nL + 1, 0), poseStart(nP + 1, 0), ptEdges(nE), poseEdges(nE);
    for (int e = 0; e < nE; ++e) {
        const int p = P->edgePoint[e], c = P->edgePose[e];
        if (p < 0 || p >= nL || c < 0 || c >= nP) { set_error("lba_solve: edge index out of range"); return ORB_ERR_ARG; }
        ++ptStart[p + 1]; ++poseStart[c + 1];
    }
    for (int i = 0; i < nL; ++i) ptStart[i + 1] += ptStart[i];
    for (int i = 0; i < nP; ++i) poseStart[i + 1] += poseStart[i];
    {
        std::vector<int> a(ptStart.begin(), ptStart.end() - 1), b(poseStart.begin(), poseStart.end() - 1);
        for (int e = 0; e < nE; ++e) { ptEdges[a[P->edgePoint[e]]++] = e; poseEdges[b[P->edgePose[e]]++] = e; }
    }
    std::vector<int> edgeAt((size_t)nL * std::max(nF, 1), -1);
    for (int e = 0; e < nE; ++e) {
        const int hI = hidx[P->edgePose[e]];
        if (hI >= 0) {
            int& slot = edgeAt[(size_t)P->edgePoint[e] * nF + hI];
            if (slot >= 0) { set_error("lba_solve: duplicate (point, keyframe) observation"); return ORB_ERR_ARG; }
            slot = e;
        }
    }
    // ---- carve the arena ----
    uint8_t* base = S.d_arena; size_t off = 0;
    auto carve = [&](size_t bytes) { uint8_t* p = base + off; off += (bytes + 255) & ~(size_t)255; return p; };
    Dev D; memset(&D, 0, sizeof(D));
    D.nP = nP; D.nL = nL; D.nE = nE; D.nF = nF; D.n = n;
    D.poses = (double*)carve(56 * (size_t)nP); D.posesBk = (double*)carve(56 * (size_t)nP);
    D.pts = (double*)carve(24 * (size_t)nL); D.ptsBk = (double*)carve(24 * (size_t)nL);
    float* d_cam = (float*)carve(16 * (size_t)nP); int* d_hidx = (int*)carve(4 * (size_t)nP); int* d_free = (int*)carve(4 * (size_t)std::max(nF, 1));
    int* d_ePt = (int*)carve(4 * (size_t)nE); int* d_ePose = (int*)carve(4 * (size_t)nE);
    double* d_obs = (double*)carve(16 * (size_t)nE); float* d_is2 = (float*)carve(4 * (size_t)nE);
    int* d_ptStart = (int*)carve(4 * (size_t)(nL + 1)); int* d_ptEdges = (int*)carve(4 * (size_t)nE);
    int* d_poseStart = (int*)carve(4 * (size_t)(nP + 1)); int* d_poseEdges = (int*)carve(4 * (size_t)nE);
    int* d_edgeAt = (int*)carve(4 * edgeAt.size());
    D.err = (double*)carve(16 * (size_t)nE); D.W = (double*)carve(144 * (size_t)nE); D.Y = (double*)carve(144 * (size_t)nE);
    D.Hpp = (double*)carve(288 * (size_t)std::max(nF, 1)); D.bp = (double*)carve(48 * (size_t)std::max(nF, 1));
    D.Hll = (double*)carve(72 * (size_t)nL); D.bl = (double*)carve(24 * (size_t)nL);
    D.Dinv = (double*)carve(72 * (size_t)nL); D.db = (double*)carve(24 * (size_t)nL);
    D.Hs = (double*)carve(8 * (size_t)n * n); D.bs = (double*)carve(8 * (size_t)std::max(n, 1));
    D.x = (double*)carve(8 * ((size_t)n + 3 * (size_t)nL)); D.partial = (double*)carve(8 * 8192); D.scal = (double*)carve(64);
    double* d_chi2 = (double*)carve(8 * (size_t)nE); uint8_t* d_dpos = (uint8_t*)carve((size_t)nE);
    if (off > S.arenaBytes) { set_error("lba_solve: arena too small (internal sizing error)"); return ORB_ERR_CAPACITY; }
    D.cam = d_cam; D.hidx = d_hidx; D.freePose = d_free; D.ePt = d_ePt; D.ePose = d_ePose; D.obs = d_obs; D.invSigma2 = d_is2;
    D.ptStart = d_ptStart; D.ptEdges = d_ptEdges; D.poseStart = d_poseStart; D.poseEdges = d_poseEdges; D.edgeAt = d_edgeAt;
    D.delta = P->huberDelta; D.dsqr = P->huberDelta * P->huberDelta;
    cudaStream_t st = S.st;
#define H2D(dst, src, bytes) CK(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, st))
    H2D(D.poses, P->poses, 56 * (size_t)nP); H2D(D.pts, P->points, 24 * (size_t)nL); H2D(d_cam, P->cam, 16 * (size_t)nP);
    H2D(d_hidx, hidx.data(), 4 * (size_t)nP); if (nF) H2D(d_free, freePose.data(), 4 * (size_t)nF);
    H2D(d_ePt, P->edgePoint, 4 * (size_t)nE); H2D(d_ePose, P->edgePose, 4 * (size_t)nE); H2D(d_obs, P->obs, 16 * (size_t)nE);
    H2D(d_is2, P->invSigma2, 4 * (size_t)nE); H2D(d_ptStart, ptStart.data(), 4 * (size_t)(nL + 1)); H2D(d_ptEdges, ptEdges.data(), 4 * (size_t)nE);
    H2D(d_poseStart, poseStart.data(), 4 * (size_t)(nP + 1)); H2D(d_poseEdges, poseEdges.data(), 4 * (size_t)nE);
    H2D(d_edgeAt, edgeAt.data(), 4 * edgeAt.size());
    CK(cudaMemsetAsync(D.x, 0, 8 * ((size_t)n + 3 * (size_t)nL), st));
    if (nE) CK(cudaMemsetAsync(D.err, 0, 16 * (size_t)nE, st));   // EdgeSE3ProjectXYZ::_error before the first computeError
    int launches = 0;
    k_normalize_poses<<<(nP + 127) / 128, 128, 0, st>>>(D); ++launches;
    CK(cudaStreamSynchronize(st));   // the host vectors above must outlive the copies

    const int gE = std::min(1024, std::max(1, (nE + NT - 1) / NT));
    const int gV = std::min(1024, std::max(1, (nP + nL + NT - 1) / NT));
    const int gPt = std::min(4096, std::max(1, (nL + 7) / 8));
    const int gL = std::min(1024, std::max(1, (nL + NT - 1) / NT));
    const int nPairs = nF * (nF + 1) / 2;
    auto terminate = [&]() { return P->stopFlag && *P->stopFlag; };
    auto robust_chi2 = [&](double* out) -> int {
        k_errors<<<gE, NT, 0, st>>>(D);
        k_reduce<<<1, NT, 0, st>>>(D, gE, 0, 0);
        launches += 2;
        CK(cudaMemcpyAsync(S.h_scal, D.scal, 32, cudaMemcpyDeviceToHost, st));
        CK(cudaStreamSynchronize(st));
        *out = S.h_scal[0];
        return ORB_OK;
    };

    // ---- SparseOptimizer::optimize + OptimizationAlgorithmLevenberg::solve (control flow on the host, one small D2H per trial) ----
    double lambda = -1, ni = 2, currentChi = 0, firstChi = 0;
    int nBad = 0, cj = 0, trials = 0;
    const int maxTrials = 10;
    const double goodUpper = 2. / 3., goodLower = 1. / 3., tau = 1e-5;
    bool ok = true;
    for (int it = 0; it < P->iterations && !terminate() && ok; ++it) {
        int rc = robust_chi2(&currentChi);
        if (rc) return rc;
        double tempChi = currentChi;
        const double iniChi = currentChi;
        if (it == 0) firstChi = iniChi;
        if (nL) { k_build_points<<<gPt, NT, 0, st>>>(D); ++launches; }
        if (nF) { k_build_poses<<<nF, NT, 0, st>>>(D); ++launches; }
        if (it == 0) {
            if (P->userLambdaInit > 0) lambda = P->userLambdaInit;
            else {
                k_maxdiag<<<1, NT, 0, st>>>(D); ++launches;
                CK(cudaMemcpyAsync(S.h_scal, D.scal, 32, cudaMemcpyDeviceToHost, st));
                CK(cudaStreamSynchronize(st));
                lambda = tau * S.h_scal[2];
            }
            ni = 2; nBad = 0;
        }
        double rho = 0;
        int qmax = 0;
        do {
            if (nL) { k_point_prep<<<gL, NT, 0, st>>>(D, lambda); ++launches; }
            if (nF) {
                k_edge_y<<<gE, NT, 0, st>>>(D);
                k_schur<<<nPairs, NT, 0, st>>>(D, lambda);
                k_ldlt<<<1, NT, 0, st>>>(D);
                launches += 3;
            }
            if (nL) { k_backsub<<<gL, NT, 0, st>>>(D); ++launches; }
            k_update<<<gV, NT, 0, st>>>(D, lambda);
            k_reduce<<<1, NT, 0, st>>>(D, gV, 1, 4096);
            launches += 2;
            rc = robust_chi2(&tempChi);   // also brings scal[1] (scale) and scal[3] (ldlt ok)
            if (rc) return rc;
            const bool ok2 = nF ? (S.h_scal[3] != 0.0) : true;
            if (!ok2) tempChi = DBL_MAX;
            rho = currentChi - tempChi;
            double scale = S.h_scal[1];
            scale += 1e-3;
            rho /= scale;
            if (rho > 0 && std::isfinite(tempChi)) {
                double alpha = 1. - pow((2 * rho - 1), 3);
                alpha = std::min(alpha, goodUpper);
                const double scaleFactor = std::max(goodLower, alpha);
                lambda *= scaleFactor;
                ni = 2;
                currentChi = tempChi;
            } else {
                lambda *= ni;
                ni *= 2;
                k_restore<<<gV, NT, 0, st>>>(D); ++launches;
            }
            ++qmax; ++trials;
        } while (rho < 0 && qmax < maxTrials && !terminate());
        ++cj;
        if (qmax == maxTrials || rho == 0) ok = false;
        else {
            if ((iniChi - currentChi) * 1e3 < iniChi) ++nBad; else nBad = 0;
            if (nBad >= 3) ok = false;
        }
    }
    if (nE) { k_finalize<<<gE, NT, 0, st>>>(D, d_chi2, d_dpos); ++launches; }
    CK(cudaGetLastError());
    CK(cudaMemcpyAsync(R->poses, D.poses, 56 * (size_t)nP, cudaMemcpyDeviceToHost, st));
    if (nL) CK(cudaMemcpyAsync(R->points, D.pts, 24 * (size_t)nL, cudaMemcpyDeviceToHost, st));
    if (nE) {
        CK(cudaMemcpyAsync(R->edgeChi2, d_chi2, 8 * (size_t)nE, cudaMemcpyDeviceToHost, st));
        CK(cudaMemcpyAsync(R->edgeDepthPositive, d_dpos, (size_t)nE, cudaMemcpyDeviceToHost, st));
    }
    CK(cudaStreamSynchronize(st));
    R->iterations = cj; R->trials = trials; R->lambda = lambda; R->chi2 = currentChi; R->initialChi2 = firstChi; R->gpuLaunches = launches;
    S.launches = launches;
    return ORB_OK;
}

}  // extern "C"
