"""TEST INFRASTRUCTURE.  Build-time helper of oracle/Makefile: writes the listed LINE RANGES of the reference's sources
(read where they lie under <ref>) into <out>/*.inc so that ref_wrap_matcher.cpp can #include them verbatim.  <out> is
oracle/_ref/gen/ -- git-ignored; no reference text is ever committed.  Each range is guarded by an anchor that must
appear on its first line, so a drifted checkout fails the build instead of compiling something else."""
import os
import sys

RANGES = [
    # out name, file, first, last (1-based, inclusive), anchor on first line
    ('matcher_consts.inc', 'src/ORBmatcher.cc', 35, 41, 'const int ORBmatcher::TH_HIGH'),
    ('matcher_local_map.inc', 'src/ORBmatcher.cc', 43, 221, 'int ORBmatcher::SearchByProjection(Frame &F, const vector<MapPoint*>'),
    ('matcher_bow_kf_frame.inc', 'src/ORBmatcher.cc', 223, 425, 'int ORBmatcher::SearchByBoW(KeyFrame* pKF,Frame &F'),
    ('matcher_init.inc', 'src/ORBmatcher.cc', 648, 763, 'int ORBmatcher::SearchForInitialization'),
    ('matcher_bow_kf_kf.inc', 'src/ORBmatcher.cc', 765, 905, 'int ORBmatcher::SearchByBoW(KeyFrame *pKF1, KeyFrame *pKF2, vector<MapPoint *> &vpMatches12)'),
    ('matcher_triangulation.inc', 'src/ORBmatcher.cc', 907, 1146, 'int ORBmatcher::SearchForTriangulation(KeyFrame *pKF1, KeyFrame *pKF2'),
    ('pinhole_epipolar.inc', 'src/CameraModels/Pinhole.cpp', 107, 129, 'bool Pinhole::epipolarConstrain(GeometricCamera* pCamera2'),
    ('matcher_fuse.inc', 'src/ORBmatcher.cc', 1148, 1338, 'int ORBmatcher::Fuse(KeyFrame *pKF, const vector<MapPoint *> &vpMapPoints'),
    ('matcher_fuse_sim3.inc', 'src/ORBmatcher.cc', 1340, 1455, 'int ORBmatcher::Fuse(KeyFrame *pKF, Sophus::Sim3f &Scw, const vector<MapPoint *> &vpPoints'),
    ('matcher_sim3.inc', 'src/ORBmatcher.cc', 1457, 1674, 'int ORBmatcher::SearchBySim3(KeyFrame* pKF1, KeyFrame* pKF2'),
    ('mappoint_index_in_kf.inc', 'src/MapPoint.cc', 411, 418, 'tuple<int,int> MapPoint::GetIndexInKeyFrame(KeyFrame *pKF)'),
    ('matcher_last_frame.inc', 'src/ORBmatcher.cc', 1676, 1887, 'int ORBmatcher::SearchByProjection(Frame &CurrentFrame, const Frame &LastFrame'),
    ('matcher_maxima_distance.inc', 'src/ORBmatcher.cc', 2012, 2074, 'void ORBmatcher::ComputeThreeMaxima'),
    ('frame_assign_grid.inc', 'src/Frame.cc', 385, 416, 'void Frame::AssignFeaturesToGrid()'),
    ('frame_in_frustum_mono.inc', 'src/Frame.cc', 512, 574, 'bool Frame::isInFrustum(MapPoint *pMP, float viewingCosLimit)'),
    ('frame_stereo_matches.inc', 'src/Frame.cc', 811, 982, 'void Frame::ComputeStereoMatches()'),
    ('frame_features_in_area.inc', 'src/Frame.cc', 657, 735, 'vector<size_t> Frame::GetFeaturesInArea'),
    ('mappoint_distinctive.inc', 'src/MapPoint.cc', 329, 403, 'void MapPoint::ComputeDistinctiveDescriptors()'),
    ('keyframe_features_in_area.inc', 'src/KeyFrame.cc', 704, 753, 'vector<size_t> KeyFrame::GetFeaturesInArea'),
    ('mappoint_predict_scale_kf.inc', 'src/MapPoint.cc', 514, 529, 'int MapPoint::PredictScale(const float &currentDist, KeyFrame* pKF)'),
    ('mappoint_invariance.inc', 'src/MapPoint.cc', 502, 512, 'float MapPoint::GetMinDistanceInvariance()'),
    ('mappoint_predict_scale.inc', 'src/MapPoint.cc', 531, 546, 'int MapPoint::PredictScale(const float &currentDist, Frame* pF)'),
    ('g2o_imucampose_project.inc', 'src/G2oTypes.cc', 170, 175, 'Eigen::Vector2d ImuCamPose::Project(const Eigen::Vector3d &Xw, int cam_idx) const'),
    ('g2o_imucampose_depth.inc', 'src/G2oTypes.cc', 187, 190, 'bool ImuCamPose::isDepthPositive(const Eigen::Vector3d &Xw, int cam_idx) const'),
    ('g2o_imucampose_update.inc', 'src/G2oTypes.cc', 192, 220, 'void ImuCamPose::Update(const double *pu)'),
    ('g2o_edge_mono.inc', 'src/G2oTypes.cc', 349, 395, 'void EdgeMono::linearizeOplus()'),
    ('g2o_edge_inertial.inc', 'src/G2oTypes.cc', 514, 594, 'void EdgeInertial::computeError()'),
    ('g2o_edge_prior.inc', 'src/G2oTypes.cc', 731, 760, 'void EdgePriorPoseImu::computeError()'),
    ('g2o_so3.inc', 'src/G2oTypes.cc', 777, 861, 'Eigen::Matrix3d ExpSO3(const Eigen::Vector3d &w)'),
    ('pinhole_project_d.inc', 'src/CameraModels/Pinhole.cpp', 35, 41, 'Eigen::Vector2d Pinhole::project(const Eigen::Vector3d &v3D)'),
    ('pinhole_project_jac.inc', 'src/CameraModels/Pinhole.cpp', 71, 81, 'Eigen::Matrix<double, 2, 3> Pinhole::projectJac(const Eigen::Vector3d &v3D)'),
    ('g2o_se3_skew.inc', 'Thirdparty/g2o/g2o/types/se3_ops.hpp', 27, 38, 'Matrix3d skew(const Vector3d&v)'),
    ('g2o_se3quat_mul.inc', 'Thirdparty/g2o/g2o/types/se3quat.h', 104, 110, 'inline SE3Quat operator* (const SE3Quat& tr2) const{'),
    ('g2o_se3quat_map.inc', 'Thirdparty/g2o/g2o/types/se3quat.h', 217, 220, 'Vector3d map(const Vector3d & xyz) const'),
    ('g2o_se3quat_exp.inc', 'Thirdparty/g2o/g2o/types/se3quat.h', 223, 261, 'static SE3Quat exp(const Vector6d & update)'),
    ('g2o_se3quat_normalize.inc', 'Thirdparty/g2o/g2o/types/se3quat.h', 284, 289, 'void normalizeRotation(){'),
    ('edge_se3_only_pose.inc', 'src/OptimizableTypes.cpp', 49, 63, 'void EdgeSE3ProjectXYZOnlyPose::linearizeOplus() {'),
    ('edge_se3_xyz.inc', 'src/OptimizableTypes.cpp', 139, 160, 'void EdgeSE3ProjectXYZ::linearizeOplus() {'),
    ('imu_integrated_rotation.inc', 'src/ImuTypes.cc', 84, 105, 'IntegratedRotation::IntegratedRotation(const Eigen::Vector3f &angVel, const Bias &imuBias, const float &time) {'),
    ('imu_initialize.inc', 'src/ImuTypes.cc', 147, 166, 'void Preintegrated::Initialize(const Bias &b_)'),
    ('imu_integrate.inc', 'src/ImuTypes.cc', 177, 236, 'void Preintegrated::IntegrateNewMeasurement(const Eigen::Vector3f &acceleration, const Eigen::Vector3f &angVel, const float &dt)'),
    ('optimizer_pose_optimization_rounds.inc', 'src/Optimizer.cc', 996, 1104, 'if(nInitialCorrespondences<3)'),
    ('optimizer_local_inertial_ba_tail.inc', 'src/Optimizer.cc', 2840, 2895, 'optimizer.initializeOptimization();'),
    ('g2o_gauss_newton_solve.inc', 'Thirdparty/g2o/g2o/core/optimization_algorithm_gauss_newton.cpp', 50, 93, 'OptimizationAlgorithm::SolverResult OptimizationAlgorithmGaussNewton::solve(int iteration, bool online)'),
    ('optimizer_pose_inertial_kf_rounds.inc', 'src/Optimizer.cc', 4698, 4823, 'float chi2Mono[4]={12,7.5,5.991,5.991};'),
    ('optimizer_pose_inertial_lf_rounds.inc', 'src/Optimizer.cc', 5098, 5221, 'const float chi2Mono[4]={5.991,5.991,5.991,5.991};'),
    ('tracking_preintegrate_select.inc', 'src/Tracking.cc', 1646, 1678, 'while(true)'),
    ('tracking_preintegrate_steps.inc', 'src/Tracking.cc', 1680, 1729, 'const int n = mvImuFromLastFrame.size()-1;'),
    ('g2o_huber.inc', 'Thirdparty/g2o/g2o/core/robust_kernel_impl.cpp', 78, 91, 'void RobustKernelHuber::robustify(double e, Eigen::Vector3d& rho) const'),
    ('g2o_levenberg_solve.inc', 'Thirdparty/g2o/g2o/core/optimization_algorithm_levenberg.cpp', 61, 194, 'OptimizationAlgorithm::SolverResult OptimizationAlgorithmLevenberg::solve(int iteration, bool online)'),
    ('g2o_sparse_optimizer_optimize.inc', 'Thirdparty/g2o/g2o/core/sparse_optimizer.cpp', 354, 419, 'int SparseOptimizer::optimize(int iterations, bool online)'),
    ('euroc_loaders.inc', 'Examples/Monocular-Inertial/mono_inertial_euroc.cc', 252, 310, 'void LoadImages(const string &strImagePath, const string &strPathTimes,'),
    ('pinhole_project.inc', 'src/CameraModels/Pinhole.cpp', 43, 49, 'Eigen::Vector2f Pinhole::project(const Eigen::Vector3f &v3D)'),
]


def main(ref, out):
    os.makedirs(out, exist_ok=True)
    for name, rel, a, b, anchor in RANGES:
        lines = open(os.path.join(ref, rel), encoding='utf-8', errors='replace').read().split('\n')
        if anchor not in lines[a - 1]:
            sys.exit('extract_ranges: %s:%d does not start with %r (reference drifted?)' % (rel, a, anchor))
        with open(os.path.join(out, name), 'w') as f:
            f.write('// generated from %s:%d-%d of the reference at build time -- do not commit\n' % (rel, a, b))
            f.write('#line %d "%s"\n' % (a, os.path.join(ref, rel)))
            f.write('\n'.join(lines[a - 1:b]) + '\n')


if __name__ == '__main__':
    main(sys.argv[1], sys.argv[2])
