// TEST INFRASTRUCTURE ONLY -- CPU oracle for the ORB-SLAM3 hot path.
//
// This directory restates, on the CPU, the algorithm of the reference
// (lturing/ORB_SLAM3_modified) for the path BASELINE.json names.  Only tests/,
// __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs
// may load it; the product (orb_slam3_modified_b200/) never does.
//
// PARITY PINNING: the reference ships no tests or golden vectors for this path and its own
// build cannot run here (OpenCV C++, Eigen, Boost, Pangolin are absent).  What pins this restatement:
//   * the OpenCV primitives (resize, FAST, GaussianBlur, fastAtan2, BFMatcher) against the independent cv2 4.13 wheel
//     (fixtures under tests/golden/ made by tools/make_golden.py);
//   * the reference's OWN SOURCE TEXT, compiled where it lies under /root/reference by oracle/Makefile into oracle/_ref/
//     against minimal stand-in types (oracle/ref_shim/): src/ORBextractor.cc whole, the ORBmatcher / Frame / KeyFrame /
//     MapPoint / DBoW2 function bodies (bit for bit), the g2o edge types and IMU preintegration (to rounding), g2o's
//     Levenberg / Gauss-Newton / optimize() text and the optimiser loops of src/Optimizer.cc driving this oracle's
//     linear algebra (bit for bit) -- tests/test_ref_pins_*_cpu.py, DESIGN.md section 2;
//   * independent derivations (numerical Jacobians, dense solves, scipy) for what needs Eigen's solvers.
#pragma once
#include <cstdint>
#include <cmath>
#include <vector>

namespace orbo {

// Layout-compatible with cv::KeyPoint (28 bytes).
struct KeyPoint {
    float x, y;
    float size;
    float angle;
    float response;
    int octave;
    int class_id;
};
static_assert(sizeof(KeyPoint) == 28, "cv::KeyPoint layout");

// cvRound(double/float): round-half-to-even (SSE cvtsd2si), OpenCV core/fast_math.hpp.
static inline int cvRound(double v) { return (int)std::nearbyint(v); }
static inline int cvRoundf(float v) { return (int)std::nearbyintf(v); }
static inline int cvFloor(double v) { int i = (int)v; return i - (i > v); }
static inline int cvCeil(double v) { int i = (int)v; return i + (i < v); }

}  // namespace orbo

// A solver state of the oracle (LocalBundleAdjustment or LocalInertialBA) opened step by step: the operations g2o's OptimizationAlgorithmLevenberg::solve and
// SparseOptimizer::optimize call on their Solver / SparseOptimizer.  oracle/ref_shim/ref_wrap_g2o_lm.cpp drives these with the REFERENCE's own text of those two
// functions, which pins the control flow of the oracle's Levenberg-Marquardt loops (tests/test_ref_pins_lm_cpu.py).
extern "C" {
typedef struct OrboLmBackend {
    void* self;
    void (*compute_errors)(void*);
    double (*robust_chi2)(void*);
    void (*build_system)(void*);
    int (*solve)(void*, double lambda);           /* BlockSolver::setLambda + solve + restoreDiagonal; 0 = the factorisation failed */
    void (*update)(void*);                        /* SparseOptimizer::update(solver->x()) */
    void (*push)(void*);
    void (*pop)(void*);
    int (*vector_size)(void*);
    const double* (*x)(void*);
    const double* (*b)(void*);
    int (*n_diag)(void*);
    const double* (*diag)(void*);                 /* hessian(j, j) of every vertex in index-mapping order (computeLambdaInit) */
} OrboLmBackend;
}
