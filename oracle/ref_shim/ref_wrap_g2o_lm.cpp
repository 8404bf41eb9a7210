// TEST INFRASTRUCTURE ONLY -- g2o's own Levenberg-Marquardt control flow driving the oracle's linear algebra:
//   OptimizationAlgorithmLevenberg::solve / computeLambdaInit / computeScale      Thirdparty/g2o/g2o/core/optimization_algorithm_levenberg.cpp:61-194
//   SparseOptimizer::optimize                                                     Thirdparty/g2o/g2o/core/sparse_optimizer.cpp:354-419
//   the four rounds of Optimizer::PoseOptimization                                src/Optimizer.cc:996-1104
//   Optimizer::LocalInertialBA from initializeOptimization() to the FAIL test     src/Optimizer.cc:2840-2895
//   OptimizationAlgorithmGaussNewton::solve                                       Thirdparty/g2o/g2o/core/optimization_algorithm_gauss_newton.cpp:50-93
//   the four rounds + recovery of Optimizer::PoseInertialOptimizationLastKeyFrame  src/Optimizer.cc:4698-4823
//   the four rounds + recovery of Optimizer::PoseInertialOptimizationLastFrame     src/Optimizer.cc:5098-5221
// are cut out of the reference at build time (extract_ranges.py -> oracle/_ref/gen/*.inc) and compiled as they are against the class shells below, which
// carry the members those bodies touch under g2o's names (optimization_algorithm_levenberg.h, optimization_algorithm_with_hessian.h, sparse_optimizer.h,
// solver.h, property.h, batch_stats.h).  The shells' Solver and SparseOptimizer operations (buildSystem, solve, update, push / pop, computeActiveErrors,
// activeRobustChi2 ...) forward to an OrboLmBackend = a solver state of the oracle opened step by step (oracle_common.h): so the iteration / trial / lambda /
// stop logic that runs is the REFERENCE's text, and tests/test_ref_pins_lm_cpu.py requires the oracle's own loops to give the same result bit for bit.
// Compiled with -ffp-contract=off like the oracle (the reference build would fuse computeScale's multiply-adds: a last-bit matter of rho, not of control flow).
#include <cassert>
#include <cmath>
#include <iostream>
#include <limits>
#include <mutex>
#include <utility>
#include <vector>

#include "../oracle_common.h"

#define FIXED(s) (s)

namespace g2o {
using namespace std;

inline double get_monotonic_time() { return 0.0; }
inline bool g2o_isfinite(double x) { return std::isfinite(x); }

struct G2OBatchStatistics {
    int iteration = 0, numVertices = 0, numEdges = 0, levenbergIterations = 0;
    double chi2 = 0, timeResiduals = 0, timeQuadraticForm = 0, timeLinearSolution = 0, timeUpdate = 0, timeIteration = 0;
    static G2OBatchStatistics* globalStats() { return nullptr; }
    static void setGlobalStats(G2OBatchStatistics*) {}
};
template <typename T> struct Property {
    T _value;
    explicit Property(T v) : _value(v) {}
    const T& value() const { return _value; }
    void setValue(const T& v) { _value = v; }
};

struct OptimizableGraph {
    struct Vertex {
        int _dim; const double* _diag;
        int dimension() const { return _dim; }
        double hessian(int i, int j) const { return i == j ? _diag[i] : 0.0; }
    };
};
class SparseOptimizer;
class Solver {                                           // solver.h: what OptimizationAlgorithmLevenberg calls
public:
    const OrboLmBackend* be; SparseOptimizer* _optimizer;
    SparseOptimizer* optimizer() const { return _optimizer; }
    bool buildStructure(bool = false) { return true; }
    bool buildSystem() { be->build_system(be->self); return true; }
    bool setLambda(double lambda, bool = false) { _lambda = lambda; return true; }
    bool solve() { return be->solve(be->self, _lambda) != 0; }
    void restoreDiagonal() {}
    const double* x() const { return be->x(be->self); }
    const double* b() const { return be->b(be->self); }
    size_t vectorSize() const { return (size_t)be->vector_size(be->self); }
    bool schur() { return true; }
    double _lambda = 0;
};
class OptimizationAlgorithm {                            // optimization_algorithm.h
public:
    enum SolverResult { Terminate = 2, OK = 1, Fail = -1 };
    virtual ~OptimizationAlgorithm() {}
    virtual bool init(bool online = false) = 0;
    virtual SolverResult solve(int iteration, bool online = false) = 0;
    virtual void printVerbose(std::ostream&) const {}
    SparseOptimizer* _optimizer = nullptr;
};
class OptimizationAlgorithmLevenberg : public OptimizationAlgorithm {   // optimization_algorithm_levenberg.h + the constructor of the .cpp (:43-55)
public:
    explicit OptimizationAlgorithmLevenberg(Solver* solver) : _solver(solver) {
        _currentLambda = -1.;
        _tau = 1e-5;
        _goodStepUpperScale = 2. / 3.;
        _goodStepLowerScale = 1. / 3.;
        _userLambdaInit = new Property<double>(0.);
        _maxTrialsAfterFailure = new Property<int>(10);
        _ni = 2.;
        _levenbergIterations = 0;
        _nBad = 0;
    }
    ~OptimizationAlgorithmLevenberg() { delete _userLambdaInit; delete _maxTrialsAfterFailure; }
    bool init(bool = false) { return true; }             // OptimizationAlgorithmWithHessian::init: solver->init (nothing to do for the shell)
    SolverResult solve(int iteration, bool online = false);
    double computeLambdaInit() const;
    double computeScale() const;
    Solver* _solver;
    Property<int>* _maxTrialsAfterFailure;
    Property<double>* _userLambdaInit;
    double _currentLambda, _tau, _goodStepLowerScale, _goodStepUpperScale, _ni;
    int _levenbergIterations, _nBad;
    long totalTrials = 0;                                // (shell bookkeeping, not g2o's)
};
class SparseOptimizer {                                  // sparse_optimizer.h
public:
    typedef std::vector<OptimizableGraph::Vertex*> VertexContainer;
    const OrboLmBackend* be = nullptr;
    VertexContainer _ivMap, _activeVertices;
    std::vector<int> _activeEdges;
    std::vector<G2OBatchStatistics> _batchStatistics;
    bool _computeBatchStatistics = false;
    OptimizationAlgorithm* _algorithm = nullptr;
    bool* _forceStopFlag = nullptr;
    std::vector<OptimizableGraph::Vertex> _vstore;
    bool terminate() { return _forceStopFlag ? (*_forceStopFlag) : false; }
    const VertexContainer& indexMapping() const {        // the diagonal the vertices expose is the backend's, refreshed here (computeLambdaInit runs after buildSystem)
        const double* d = be->diag(be->self);
        SparseOptimizer* self = const_cast<SparseOptimizer*>(this);
        for (size_t k = 0; k < self->_vstore.size(); ++k) self->_vstore[k]._diag = d + k;
        return _ivMap;
    }
    bool verbose() const { return false; }
    void preIteration(int) {}
    void postIteration(int);
    void computeActiveErrors() { be->compute_errors(be->self); }
    double activeRobustChi2() const { return be->robust_chi2(be->self); }
    void push() { be->push(be->self); }
    void pop() { be->pop(be->self); }
    void discardTop() {}
    void update(const double*) { be->update(be->self); }
    int optimize(int iterations, bool online = false);
};

class OptimizationAlgorithmGaussNewton : public OptimizationAlgorithm {   // optimization_algorithm_gauss_newton.h
public:
    explicit OptimizationAlgorithmGaussNewton(Solver* solver) : _solver(solver) {}
    bool init(bool = false) { return true; }
    SolverResult solve(int iteration, bool online = false);
    Solver* _solver;
};

inline void SparseOptimizer::postIteration(int) {
    if (OptimizationAlgorithmLevenberg* lm = dynamic_cast<OptimizationAlgorithmLevenberg*>(_algorithm)) lm->totalTrials += lm->_levenbergIterations;
}

#include "g2o_levenberg_solve.inc"
#include "g2o_gauss_newton_solve.inc"
#include "g2o_sparse_optimizer_optimize.inc"

}  // namespace g2o

// ---- Optimizer::PoseOptimization's four rounds (src/Optimizer.cc:996-1104: the early return, the pose reset, optimize(10), the chi2 re-classification with its
//      stale-error rule, the level / robust-kernel switches, the fewer-than-10-edges exit), compiled verbatim; every object it touches is a shell over the oracle's
//      pose-only state ----
extern "C" {
void orbo_poseopt_open(const float* cam4, int N, const double* Xw3, const double* obs2, const float* invSigma2, double huberDelta, OrboLmBackend* out);
void orbo_poseopt_set_estimate(void* h, const double* pose7);
void orbo_poseopt_get_estimate(void* h, double* pose7);
void orbo_poseopt_edge_compute_error(void* h, int e);
double orbo_poseopt_edge_chi2(void* h, int e);
void orbo_poseopt_edge_set_level(void* h, int e, int level);
void orbo_poseopt_edge_set_robust(void* h, int e, int on);
int orbo_poseopt_active(void* h);
void orbo_poseopt_close(void* h);
}
namespace Sophus {
struct Quatd { double w, x, y, z; Quatd cast_double() const { return *this; } };
struct Vec3d { double v[3]; };
template <class T> struct Caster { T val; template <class U> T cast() const { return val; } };
template <typename T> struct SE3 {                     // the frame's float pose: what the loop reads is unit_quaternion().cast<double>() and translation().cast<double>()
    double p7[7];
    Caster<Quatd> unit_quaternion() const { return Caster<Quatd>{Quatd{p7[0], p7[1], p7[2], p7[3]}}; }
    Caster<Vec3d> translation() const { return Caster<Vec3d>{Vec3d{{p7[4], p7[5], p7[6]}}}; }
};
}  // namespace Sophus
namespace g2o {
struct SE3Quat { double p7[7]; SE3Quat(const Sophus::Quatd& q, const Sophus::Vec3d& t) : p7{q.w, q.x, q.y, q.z, t.v[0], t.v[1], t.v[2]} {} };
struct VertexSE3Expmap { void* h; void setEstimate(const SE3Quat& e) { orbo_poseopt_set_estimate(h, e.p7); } };
struct PoseEdgeShell {                                 // computeError / chi2 / setLevel / setRobustKernel of an only-pose edge
    void* h; int idx;
    void computeError() { orbo_poseopt_edge_compute_error(h, idx); }
    double chi2() const { return orbo_poseopt_edge_chi2(h, idx); }
    void setLevel(int l) { orbo_poseopt_edge_set_level(h, idx, l); }
    void setRobustKernel(void* k) { orbo_poseopt_edge_set_robust(h, idx, k != nullptr); }
};
struct EdgeStereoSE3ProjectXYZOnlyPose : PoseEdgeShell {};
struct PoseOptimizerShell : SparseOptimizer {          // optimizer.initializeOptimization(0) / optimize(n) / edges()
    std::vector<int> _edges; void* h = nullptr; std::vector<OptimizableGraph::Vertex> _one;
    const std::vector<int>& edges() const { return _edges; }
    void initializeOptimization(int) {                 // the active set = the level-0 edges (kept by the oracle state); no active edge -> no active vertex
        _ivMap.clear();
        if (orbo_poseopt_active(h) > 0) for (size_t k = 0; k < _vstore.size(); ++k) _ivMap.push_back(&_vstore[k]);
    }
};
}  // namespace g2o
namespace ORB_SLAM3 {
struct EdgeSE3ProjectXYZOnlyPose : g2o::PoseEdgeShell {};
struct EdgeSE3ProjectXYZOnlyPoseToBody : g2o::PoseEdgeShell {};
struct Frame {
    Sophus::SE3<float> mTcw; std::vector<bool> mvbOutlier;
    Sophus::SE3<float> GetPose() const { return mTcw; }
};
static int pose_optimization_rounds(Frame* pFrame, g2o::PoseOptimizerShell& optimizer, g2o::VertexSE3Expmap* vSE3, std::vector<EdgeSE3ProjectXYZOnlyPose*>& vpEdgesMono,
                                    std::vector<size_t>& vnIndexEdgeMono, int nInitialCorrespondences) {
    Sophus::SE3<float> Tcw = pFrame->GetPose();
    std::vector<EdgeSE3ProjectXYZOnlyPoseToBody*> vpEdgesMono_FHR; std::vector<size_t> vnIndexEdgeRight;          // no right-camera / stereo observations (monocular frame)
    std::vector<g2o::EdgeStereoSE3ProjectXYZOnlyPose*> vpEdgesStereo; std::vector<size_t> vnIndexEdgeStereo;
#include "optimizer_pose_optimization_rounds.inc"
    return nInitialCorrespondences - nBad;             // :1112
}
}  // namespace ORB_SLAM3

extern "C" {
// int Optimizer::PoseOptimization(Frame*) with the reference's own four-round loop + g2o's optimize() / Levenberg text over the oracle's pose-only state.
// pose7 in/out (double, as the oracle's), outlier [N] out; returns nInitialCorrespondences - nBad.
int ref_pose_optimization(double* pose7, const float* cam4, int N, const double* Xw3, const double* obs2, const float* invSigma2, double huberDelta, unsigned char* outlier) {
    using namespace g2o;
    OrboLmBackend be;
    orbo_poseopt_open(cam4, N, Xw3, obs2, invSigma2, huberDelta, &be);
    PoseOptimizerShell opt; opt.be = &be; opt.h = be.self;
    Solver solver; solver.be = &be; solver._optimizer = &opt;
    OptimizationAlgorithmLevenberg alg(&solver);
    alg._optimizer = &opt;
    opt._algorithm = &alg;
    opt._vstore.assign(6, OptimizableGraph::Vertex{1, nullptr});
    opt._edges.assign((size_t)N, 0);
    ORB_SLAM3::Frame frame;
    for (int i = 0; i < 7; ++i) frame.mTcw.p7[i] = pose7[i];
    frame.mvbOutlier.assign((size_t)N, false);
    VertexSE3Expmap vSE3{be.self};
    std::vector<ORB_SLAM3::EdgeSE3ProjectXYZOnlyPose> store((size_t)N);
    std::vector<ORB_SLAM3::EdgeSE3ProjectXYZOnlyPose*> edges; std::vector<size_t> index;
    for (int i = 0; i < N; ++i) { store[i].h = be.self; store[i].idx = i; edges.push_back(&store[i]); index.push_back((size_t)i); }
    int ret;
    if (N < 3) ret = 0;                                 // (the early return is inside the extracted text too; the state set-up above is harmless)
    else ret = ORB_SLAM3::pose_optimization_rounds(&frame, opt, &vSE3, edges, index, N);
    if (N >= 3) orbo_poseopt_get_estimate(be.self, pose7);
    for (int i = 0; i < N; ++i) outlier[i] = frame.mvbOutlier[i] ? 1 : 0;
    orbo_poseopt_close(be.self);
    return ret;
}
}

// ---- Optimizer::LocalInertialBA between the graph set-up and the write-back (src/Optimizer.cc:2840-2895): initializeOptimization, err, optimize(opt_it), err_end, the
//      chi2 / depth test that fills vToErase, the FAIL test -- compiled verbatim over shells of the optimizer, the edges and the map points ----
extern "C" {
double orbo_liba_edge_chi2(void* h, int e);
int orbo_liba_edge_depth_positive(void* h, int e);
}
namespace ORB_SLAM3 {
struct KeyFrame { int idx; };
struct MapPoint { float mTrackDepth; bool isBad() { return false; } };
struct Map { std::mutex mMutexMapUpdate; };
struct EdgeMono { void* h; int idx; double chi2() const { return orbo_liba_edge_chi2(h, idx); } bool isDepthPositive() { return orbo_liba_edge_depth_positive(h, idx) != 0; } };
struct EdgeStereo { double chi2() const { return 0; } };
struct LibaOptimizerShell : g2o::SparseOptimizer {
    void initializeOptimization() {}
    void setForceStopFlag(bool* f) { _forceStopFlag = f; }
};
// `failed` is true on entry and cleared after the text: its `return;` (the FAIL branch) leaves it set
static void local_inertial_ba_tail(LibaOptimizerShell& optimizer, int opt_it, bool* pbStopFlag, Map* pMap, bool bLarge, std::vector<EdgeMono*>& vpEdgesMono,
                                   std::vector<MapPoint*>& vpMapPointEdgeMono, std::vector<KeyFrame*>& vpEdgeKFMono, std::vector<std::pair<KeyFrame*, MapPoint*> >& erased,
                                   float* errOut, bool* failed) {
    using namespace std;
    const float chi2Mono2 = 5.991;                     // :2688
    const float chi2Stereo2 = 7.815;                   // :2690
    std::vector<EdgeStereo*> vpEdgesStereo; std::vector<MapPoint*> vpMapPointEdgeStereo; std::vector<KeyFrame*> vpEdgeKFStereo;
    // the text's err / err_end / vToErase are handed out where it takes the map mutex (:2887), i.e. after the inlier check and before the FAIL test
#define unique_lock errOut[0] = err; errOut[1] = err_end; erased = vToErase; unique_lock
#include "optimizer_local_inertial_ba_tail.inc"
#undef unique_lock
    *failed = false;
}
}  // namespace ORB_SLAM3

extern "C" {
// stats [4]: err, err_end, failed, iterations-not-reported (the reference does not keep optimize()'s return value here); erase [nE]
void ref_local_inertial_ba_tail(const OrboLmBackend* be, int nE, const int* edgePoint, const float* trackDepth, int nL, int opt_it, double userLambdaInit, int bLarge,
                                unsigned char* erase, double* stats) {
    using namespace g2o;
    ORB_SLAM3::LibaOptimizerShell opt; opt.be = be;
    Solver solver; solver.be = be; solver._optimizer = &opt;
    OptimizationAlgorithmLevenberg alg(&solver);
    alg._optimizer = &opt;
    alg._userLambdaInit->setValue(userLambdaInit);
    opt._algorithm = &alg;
    be->build_system(be->self);
    const int nd = be->n_diag(be->self);
    opt._vstore.assign((size_t)nd, OptimizableGraph::Vertex{1, nullptr});
    for (int k = 0; k < nd; ++k) opt._ivMap.push_back(&opt._vstore[k]);
    std::vector<ORB_SLAM3::MapPoint> pts((size_t)nL);
    for (int i = 0; i < nL; ++i) pts[i].mTrackDepth = trackDepth[i];
    std::vector<ORB_SLAM3::EdgeMono> es((size_t)nE); std::vector<ORB_SLAM3::KeyFrame> kfs((size_t)nE);
    std::vector<ORB_SLAM3::EdgeMono*> vpEdgesMono; std::vector<ORB_SLAM3::MapPoint*> vpMP; std::vector<ORB_SLAM3::KeyFrame*> vpKF;
    for (int e = 0; e < nE; ++e) { es[e].h = be->self; es[e].idx = e; kfs[e].idx = e; vpEdgesMono.push_back(&es[e]); vpMP.push_back(&pts[edgePoint[e]]); vpKF.push_back(&kfs[e]); }
    ORB_SLAM3::Map map;
    std::vector<std::pair<ORB_SLAM3::KeyFrame*, ORB_SLAM3::MapPoint*> > erased;
    float err2[2] = {0, 0};
    bool failed = true;
    ORB_SLAM3::local_inertial_ba_tail(opt, opt_it, nullptr, &map, bLarge != 0, vpEdgesMono, vpMP, vpKF, erased, err2, &failed);
    for (int e = 0; e < nE; ++e) erase[e] = 0;
    for (const auto& pr : erased) erase[pr.first->idx] = 1;      // the shell keyframe of edge e carries e
    stats[0] = err2[0]; stats[1] = err2[1]; stats[2] = failed ? 1 : 0; stats[3] = alg._currentLambda;
}
}

// ---- Optimizer::PoseInertialOptimizationLastKeyFrame's four rounds + the recovery of not-too-bad points (src/Optimizer.cc:4698-4823), compiled verbatim; the optimizer
//      runs g2o's Gauss-Newton text (optimization_algorithm_gauss_newton.cpp:50-93) ----
extern "C" {
void orbo_pikf_open(int N, const float* Xw, const float* obs, const float* invSigma2, const float* trackDepth, const float* cam4, const double* extr24, const float* P,
                    const double* kfState15, const double* state15, OrboLmBackend* out);
void orbo_pikf_edge_compute_error(void* h, int e);
double orbo_pikf_edge_chi2(void* h, int e);
int orbo_pikf_edge_depth_positive(void* h, int e);
void orbo_pikf_edge_set_level(void* h, int e, int level);
void orbo_pikf_edge_set_robust(void* h, int e, int on);
void orbo_pikf_close(void* h, double* stateOut21);
}
namespace ORB_SLAM3 {
struct EdgeOps { void (*ce)(void*, int); double (*chi2)(void*, int); int (*dp)(void*, int); void (*lvl)(void*, int, int); void (*rob)(void*, int, int); };
struct EdgeMonoOnlyPose {                               // the operations of the loop on an only-pose edge, forwarded to the oracle state (last-keyframe or last-frame variant)
    void* h; int idx; const EdgeOps* ops;
    void computeError() { ops->ce(h, idx); }
    double chi2() const { return ops->chi2(h, idx); }
    bool isDepthPositive() { return ops->dp(h, idx) != 0; }
    void setLevel(int l) { ops->lvl(h, idx, l); }
    void setRobustKernel(void* k) { ops->rob(h, idx, k != nullptr); }
};
struct EdgeStereoOnlyPose { void computeError() {} double chi2() const { return 0; } void setLevel(int) {} void setRobustKernel(void*) {} };
struct InertialFrame { std::vector<bool> mvbOutlier; std::vector<MapPoint*> mvpMapPoints; };
struct InertialOptimizerShell : g2o::SparseOptimizer {
    std::vector<int> _edges;
    const std::vector<int>& edges() const { return _edges; }
    void initializeOptimization(int) {}
};
static int pose_inertial_lf_rounds(InertialFrame* pFrame, InertialOptimizerShell& optimizer, std::vector<EdgeMonoOnlyPose*>& vpEdgesMono, std::vector<size_t>& vnIndexEdgeMono, bool bRecInit) {
    std::vector<EdgeStereoOnlyPose*> vpEdgesStereo; std::vector<size_t> vnIndexEdgeStereo;
#include "optimizer_pose_inertial_lf_rounds.inc"
    return nBad;
}
static int pose_inertial_kf_rounds(InertialFrame* pFrame, InertialOptimizerShell& optimizer, std::vector<EdgeMonoOnlyPose*>& vpEdgesMono, std::vector<size_t>& vnIndexEdgeMono, bool bRecInit) {
    std::vector<EdgeStereoOnlyPose*> vpEdgesStereo; std::vector<size_t> vnIndexEdgeStereo;
#include "optimizer_pose_inertial_kf_rounds.inc"
    return nBad;
}
}  // namespace ORB_SLAM3

extern "C" {
// int Optimizer::PoseInertialOptimizationLastKeyFrame(Frame*, bool bRecInit) up to the recovery of the state: returns nInitialCorrespondences - nBad; state21 in/out, outlier [N] out
int ref_pose_inertial_opt_last_kf(int N, const float* Xw, const float* obs, const float* invSigma2, const float* trackDepth, const float* cam4, const double* extr24, const float* P,
                                  const double* kfState21, double* state21, int bRecInit, unsigned char* outlier) {
    using namespace g2o;
    OrboLmBackend be;
    orbo_pikf_open(N, Xw, obs, invSigma2, trackDepth, cam4, extr24, P, kfState21, state21, &be);
    ORB_SLAM3::InertialOptimizerShell opt; opt.be = &be;
    Solver solver; solver.be = &be; solver._optimizer = &opt;
    OptimizationAlgorithmGaussNewton alg(&solver);
    alg._optimizer = &opt;
    opt._algorithm = &alg;
    opt._vstore.assign(15, OptimizableGraph::Vertex{1, nullptr});
    for (int k = 0; k < 15; ++k) opt._ivMap.push_back(&opt._vstore[k]);
    opt._edges.assign((size_t)N + 3, 0);                 // the mono edges + EdgeInertial + EdgeGyroRW + EdgeAccRW
    ORB_SLAM3::InertialFrame frame; frame.mvbOutlier.assign((size_t)N, false);
    std::vector<ORB_SLAM3::MapPoint> mps((size_t)N);
    std::vector<ORB_SLAM3::EdgeMonoOnlyPose> store((size_t)N);
    std::vector<ORB_SLAM3::EdgeMonoOnlyPose*> edges; std::vector<size_t> index;
    static const ORB_SLAM3::EdgeOps ops = {orbo_pikf_edge_compute_error, orbo_pikf_edge_chi2, orbo_pikf_edge_depth_positive, orbo_pikf_edge_set_level, orbo_pikf_edge_set_robust};
    for (int i = 0; i < N; ++i) { mps[i].mTrackDepth = trackDepth[i]; frame.mvpMapPoints.push_back(&mps[i]); store[i].h = be.self; store[i].idx = i; store[i].ops = &ops; edges.push_back(&store[i]); index.push_back((size_t)i); }
    const int nBad = ORB_SLAM3::pose_inertial_kf_rounds(&frame, opt, edges, index, bRecInit != 0);
    for (int i = 0; i < N; ++i) outlier[i] = frame.mvbOutlier[i] ? 1 : 0;
    orbo_pikf_close(be.self, state21);
    return N - nBad;
}
}

extern "C" {
void orbo_pilf_open(int N, const float* Xw, const float* obs, const float* invSigma2, const float* trackDepth, const float* cam4, const double* extr24, const float* Pframe,
                    const float* Pkf, const double* prior21, const double* priorH, const double* prevState21, const double* state21, OrboLmBackend* out);
void orbo_pilf_edge_compute_error(void* h, int e);
double orbo_pilf_edge_chi2(void* h, int e);
int orbo_pilf_edge_depth_positive(void* h, int e);
void orbo_pilf_edge_set_level(void* h, int e, int level);
void orbo_pilf_edge_set_robust(void* h, int e, int on);
void orbo_pilf_close(void* h, double* prevOut21, double* stateOut21);
// int Optimizer::PoseInertialOptimizationLastFrame(Frame*, bool bRecInit): the four rounds + recovery (src/Optimizer.cc:5098-5221) up to the recovery of the states
int ref_pose_inertial_opt_last_frame(int N, const float* Xw, const float* obs, const float* invSigma2, const float* trackDepth, const float* cam4, const double* extr24, const float* Pframe,
                                     const float* Pkf, const double* prior21, const double* priorH, double* prevState21, double* state21, int bRecInit, unsigned char* outlier) {
    using namespace g2o;
    OrboLmBackend be;
    orbo_pilf_open(N, Xw, obs, invSigma2, trackDepth, cam4, extr24, Pframe, Pkf, prior21, priorH, prevState21, state21, &be);
    ORB_SLAM3::InertialOptimizerShell opt; opt.be = &be;
    Solver solver; solver.be = &be; solver._optimizer = &opt;
    OptimizationAlgorithmGaussNewton alg(&solver);
    alg._optimizer = &opt;
    opt._algorithm = &alg;
    opt._vstore.assign(30, OptimizableGraph::Vertex{1, nullptr});
    for (int k = 0; k < 30; ++k) opt._ivMap.push_back(&opt._vstore[k]);
    opt._edges.assign((size_t)N + 4, 0);                 // the mono edges + EdgeInertial + EdgeGyroRW + EdgeAccRW + EdgePriorPoseImu
    ORB_SLAM3::InertialFrame frame; frame.mvbOutlier.assign((size_t)N, false);
    std::vector<ORB_SLAM3::MapPoint> mps((size_t)N);
    std::vector<ORB_SLAM3::EdgeMonoOnlyPose> store((size_t)N);
    std::vector<ORB_SLAM3::EdgeMonoOnlyPose*> edges; std::vector<size_t> index;
    static const ORB_SLAM3::EdgeOps ops = {orbo_pilf_edge_compute_error, orbo_pilf_edge_chi2, orbo_pilf_edge_depth_positive, orbo_pilf_edge_set_level, orbo_pilf_edge_set_robust};
    for (int i = 0; i < N; ++i) { mps[i].mTrackDepth = trackDepth[i]; frame.mvpMapPoints.push_back(&mps[i]); store[i].h = be.self; store[i].idx = i; store[i].ops = &ops; edges.push_back(&store[i]); index.push_back((size_t)i); }
    const int nBad = ORB_SLAM3::pose_inertial_lf_rounds(&frame, opt, edges, index, bRecInit != 0);
    for (int i = 0; i < N; ++i) outlier[i] = frame.mvbOutlier[i] ? 1 : 0;
    orbo_pilf_close(be.self, prevState21, state21);
    return N - nBad;
}
}

extern "C" {
// optimizer.optimize(iterations) with g2o's own text on the oracle state `be`.  stats [4]: final lambda, total Levenberg trials, last result of solve(), -.
int ref_g2o_optimize(const OrboLmBackend* be, int iterations, double userLambdaInit, const unsigned char* stopFlag, double* stats) {
    using namespace g2o;
    SparseOptimizer opt; opt.be = be;
    Solver solver; solver.be = be; solver._optimizer = &opt;
    OptimizationAlgorithmLevenberg alg(&solver);
    alg._optimizer = &opt;
    alg._userLambdaInit->setValue(userLambdaInit);
    opt._algorithm = &alg;
    bool stop = false;
    if (stopFlag) { stop = *stopFlag != 0; opt._forceStopFlag = &stop; }
    be->build_system(be->self);                          // sizes the diagonal view; solve() rebuilds the system itself
    const int nd = be->n_diag(be->self);
    opt._vstore.assign((size_t)nd, OptimizableGraph::Vertex{1, nullptr});     // one scalar "vertex" per diagonal entry: computeLambdaInit only takes the maximum
    for (int k = 0; k < nd; ++k) opt._ivMap.push_back(&opt._vstore[k]);
    const int it = opt.optimize(iterations);
    if (stats) { stats[0] = alg._currentLambda; stats[1] = (double)alg.totalTrials; stats[2] = 0; stats[3] = 0; }
    return it;
}
}
