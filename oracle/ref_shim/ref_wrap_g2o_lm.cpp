// TEST INFRASTRUCTURE ONLY -- g2o's own Levenberg-Marquardt control flow driving the oracle's linear algebra:
//   OptimizationAlgorithmLevenberg::solve / computeLambdaInit / computeScale      Thirdparty/g2o/g2o/core/optimization_algorithm_levenberg.cpp:61-194
//   SparseOptimizer::optimize                                                     Thirdparty/g2o/g2o/core/sparse_optimizer.cpp:354-419
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
    void postIteration(int) { static_cast<OptimizationAlgorithmLevenberg*>(_algorithm)->totalTrials += static_cast<OptimizationAlgorithmLevenberg*>(_algorithm)->_levenbergIterations; }
    void computeActiveErrors() { be->compute_errors(be->self); }
    double activeRobustChi2() const { return be->robust_chi2(be->self); }
    void push() { be->push(be->self); }
    void pop() { be->pop(be->self); }
    void discardTop() {}
    void update(const double*) { be->update(be->self); }
    int optimize(int iterations, bool online = false);
};

#include "g2o_levenberg_solve.inc"
#include "g2o_sparse_optimizer_optimize.inc"

}  // namespace g2o

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
