/* orb_b200.h -- C-ABI of the B200-native ORB-SLAM3 hot path (liborb_b200.so).
 *
 * Drop-in boundary for three class surfaces of lturing/ORB_SLAM3_modified:
 *   ORBextractor::operator()            include/ORBextractor.h:56-58,  src/ORBextractor.cc:1086
 *   ORBmatcher::SearchByProjection x2,  include/ORBmatcher.h:47-52,    src/ORBmatcher.cc:43,1676
 *   ORBmatcher::DescriptorDistance,     include/ORBmatcher.h:43,       src/ORBmatcher.cc:2058
 *   cv::BFMatcher(NORM_HAMMING).knnMatch(k=2) as used at src/Frame.cc:1144
 *   Optimizer::LocalBundleAdjustment    include/Optimizer.h:71,        src/Optimizer.cc:1116
 *   Optimizer::PoseOptimization         include/Optimizer.h:62,        src/Optimizer.cc:814
 *
 * Conventions: extern "C", plain pointers and sizes, caller-owned output buffers with a capacity
 * and a returned count, int status (0 = ok, <0 = error enum below), no exceptions cross the
 * boundary, one call at a time per handle (the reference's extractor is stateful in the same way,
 * SURVEY.md 8b "Ownership").  Functions suffixed _device take device pointers and a cudaStream_t
 * (passed as void*), enqueue work and return without synchronising; all others take host pointers
 * and return when the results are in the caller's buffers.
 * There is no CPU fallback: every entry point fails with ORB_ERR_CUDA when no sm_100 device exists.
 */
#ifndef ORB_B200_H
#define ORB_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

enum {
    ORB_OK = 0,
    ORB_ERR_EMPTY = -1,     /* empty image: ORBextractor::operator() returns -1 (src/ORBextractor.cc:1090) */
    ORB_ERR_ARG = -2,       /* bad argument (null pointer, size out of the configured range, ...) */
    ORB_ERR_CAPACITY = -3,  /* caller buffer or an internal slab too small; nothing was truncated silently */
    ORB_ERR_CUDA = -4,      /* CUDA runtime failure or no usable device; see orb_last_error() */
    ORB_ERR_ASPECT = -5     /* bordered aspect ratio < 0.5: the reference divides by zero (src/ORBextractor.cc:559-561) */
};

/* Byte-compatible with cv::KeyPoint (28 bytes): pt.x, pt.y, size, angle, response, octave, class_id. */
typedef struct OrbKeyPoint {
    float x, y;
    float size;
    float angle;
    float response;
    int32_t octave;
    int32_t class_id;
} OrbKeyPoint;

const char* orb_last_error(void);
/* Library/ABI version and the SM architecture the kernels were compiled for (100). */
int orb_abi_version(void);
int orb_compiled_sm(void);

/* ------------------------------------------------------------------------------------------
 * ORBextractor (reference include/ORBextractor.h:43-109)
 * ------------------------------------------------------------------------------------------ */
typedef struct orbx_handle orbx_handle;

/* ORBextractor::ORBextractor(nfeatures, scaleFactor, nlevels, iniThFAST, minThFAST)
 * (src/ORBextractor.cc:409-469) plus the device-side sizing: the largest image and the number
 * of frames one batched call may carry. */
int orbx_create(orbx_handle** out, int nfeatures, float scaleFactor, int nlevels, int iniThFAST, int minThFAST,
                int max_width, int max_height, int max_batch, int device);
void orbx_destroy(orbx_handle* h);

/* GetLevels / GetScaleFactor(s) / GetInverseScaleFactors / GetScaleSigmaSquares /
 * GetInverseScaleSigmaSquares (include/ORBextractor.h:60-82) + mnFeaturesPerLevel. Arrays hold nlevels entries. */
int orbx_get_levels(const orbx_handle* h);
int orbx_get_tables(const orbx_handle* h, float* scaleFactors, float* invScaleFactors, float* levelSigma2,
                    float* invLevelSigma2, int* featuresPerLevel);
/* Upper bound of keypoints one frame can return (sum over levels of N_l + 3). */
int orbx_max_keypoints(const orbx_handle* h);

/* int ORBextractor::operator()(image, mask [ignored], keypoints, descriptors, vLappingArea)
 * image: CV_8UC1, rows x cols, `step` bytes between rows (host memory).
 * keypoints[cap], descriptors[cap*32]; *nkeypoints receives K; returns in *monoIndex what the
 * reference returns (the count of keypoints outside [lap0, lap1], filled from the front; the others
 * are filled from the back, src/ORBextractor.cc:1120-1164).  Status ORB_ERR_EMPTY mirrors `return -1`. */
int orbx_extract(orbx_handle* h, const uint8_t* image, int rows, int cols, size_t step, int lap0, int lap1,
                 OrbKeyPoint* keypoints, uint8_t* descriptors, int cap, int* nkeypoints, int* monoIndex);

/* The same call for `batch` independent frames (one per stream), host buffers.
 * images: batch frames, `frame_stride` bytes apart.  keypoints: batch x cap, descriptors: batch x cap x 32,
 * nkeypoints/monoIndex: batch entries. */
int orbx_extract_batch(orbx_handle* h, const uint8_t* images, int batch, int rows, int cols, size_t step,
                       size_t frame_stride, int lap0, int lap1, OrbKeyPoint* keypoints, uint8_t* descriptors,
                       int cap, int* nkeypoints, int* monoIndex);

/* Device-resident variant: every pointer is a device pointer, work is enqueued on `stream`
 * (cudaStream_t) and the call returns without synchronising.  Slabs have fixed capacity `cap`
 * per frame so that they can be all-gathered as-is. */
int orbx_extract_batch_device(orbx_handle* h, const uint8_t* d_images, int batch, int rows, int cols, size_t step,
                              size_t frame_stride, int lap0, int lap1, OrbKeyPoint* d_keypoints,
                              uint8_t* d_descriptors, int cap, int* d_nkeypoints, int* d_monoIndex, void* stream);

/* The host-buffer calls above also leave their results in the handle's device slabs ([batch][cap] keypoints, [batch][cap][32]
 * descriptors, [batch] counts, cap = orbx_max_keypoints()).  These pointers stay valid until the next call on the handle, so a
 * matcher call that follows (orbm_search_last_frame_batch_resident) does not have to ship the current frame back to the device. */
int orbx_resident_slabs(const orbx_handle* h, const OrbKeyPoint** d_keypoints, const uint8_t** d_descriptors, const int** d_nkeypoints,
                        int* cap);

/* void Frame::ComputeStereoMatches() (reference src/Frame.cc:811-982), the stereo consumer of ORBextractor::mvImagePyramid
 * (src/Frame.cc:122-125 runs two extractors on the left / right image, then :126 calls this): for every left keypoint the best right
 * keypoint on its row band by Hamming distance, an 11 x 11 SAD refinement over 11 shifts on the two pyramids, parabola fit, and the
 * 1.5 * 1.4 * median SAD outlier cut.  `left` and `right` must have extracted the pair's images in their last call (`batch` pairs, pair f =
 * frame f of both); mb / mbf as in Frame.  uRight / depth [batch][cap] receive mvuRight / mvDepth (-1 = no match).
 * Host form: keypoints come from both handles' resident slabs (last call = orbx_extract[_batch]); cap = orbx_max_keypoints(left). */
int orbx_stereo_matches(orbx_handle* left, orbx_handle* right, int batch, float mb, float mbf, float* uRight, float* depth, int cap);
/* Device form: keypoint / descriptor / count slabs as written by orbx_extract_batch_device of the two handles; enqueues on `stream`
 * (the stream both extractions were enqueued on, or one that waits for them). */
int orbx_stereo_matches_device(orbx_handle* left, orbx_handle* right, int batch, const OrbKeyPoint* d_kpsL, const uint8_t* d_descL, const int* d_nL,
                               int capL, const OrbKeyPoint* d_kpsR, const uint8_t* d_descR, const int* d_nR, int capR, float mb, float mbf,
                               float* d_uRight, float* d_depth, void* stream);
/* mvImagePyramid[level] of frame `frame` of the last call including the reflected frame of `border` pixels (19 in the reference,
 * src/ORBextractor.cc:1185-1191) that ComputePyramid keeps around every level: (w + 2 border) x (h + 2 border) bytes, tightly packed. */
int orbx_copy_level_bordered(orbx_handle* h, int frame, int level, int border, uint8_t* dst);

/* Debug/inspection taps used by the parity tests (mvImagePyramid is a public member of the reference class,
 * include/ORBextractor.h:84).  Copies level `level` of frame `frame` of the last call, unbordered, tightly packed. */
int orbx_get_level_size(const orbx_handle* h, int level, int* width, int* height);
int orbx_copy_level(orbx_handle* h, int frame, int level, int blurred, uint8_t* dst);
/* FAST candidates of one level of the last call as (x, y, score) int triples in the order the reference
 * feeds them to DistributeOctTree (src/ORBextractor.cc:863-871). Returns the count or <0. */
int orbx_copy_candidates(orbx_handle* h, int frame, int level, int* xys, int cap);
/* Number of kernels launched by the last extract call (for bench.py's gpu_launches). */
int orbx_last_launch_count(const orbx_handle* h);
/* Per-kernel device times of the pipeline (the reference's REGISTER_TIMES analogue, include/Settings.h:24): with profiling on,
 * CUDA events bracket each stage on the launching stream (the blur then runs in line instead of on its forked stream).
 * ms6 = pyramid, blur, FAST cells, quadtree+orientation, assemble, BRIEF of the last profiled call. */
int orbx_set_profiling(orbx_handle* h, int on);
int orbx_get_stage_ms(orbx_handle* h, float* ms6);

/* ------------------------------------------------------------------------------------------
 * ORBmatcher (reference include/ORBmatcher.h:36-103): the per-frame projection matchers, the Hamming
 * primitive and the brute-force kNN used by the reference.  All inputs are flat arrays: the host shim
 * flattens Frame / MapPoint fields exactly as listed in SURVEY.md 8a' ("Matcher in").
 * ------------------------------------------------------------------------------------------ */
typedef struct orbm_handle orbm_handle;

/* Scratch for up to max_batch frames of max_keypoints keypoints matched against max_mappoints map points (both <= 65535).
 * One call in flight per handle: the per-frame scratch (candidate lists, queries) belongs to the handle, so two concurrent calls --
 * also on different streams -- need two handles (bench.py gives every stream group its own).  Device-side octave / level values
 * outside [0, nlevels) are clamped; the host entry points reject them. */
int orbm_create(orbm_handle** out, int max_batch, int max_keypoints, int max_mappoints, int device);
void orbm_destroy(orbm_handle* h);

/* Current frame as seen by the matchers (Frame::mvKeysUn, mDescriptors, image bounds mnMinX.., mvScaleFactors;
 * the 64x48 grid of Frame::AssignFeaturesToGrid, src/Frame.cc:385-416, is rebuilt on the device). */
typedef struct OrbmFrame {
    int K;                        /* number of keypoints */
    const OrbKeyPoint* keypoints; /* K */
    const uint8_t* descriptors;   /* K x 32 */
    float minX, minY, maxX, maxY; /* mnMinX, mnMinY, mnMaxX, mnMaxY */
    const float* scaleFactors;    /* mvScaleFactors, nlevels entries */
    int nlevels;
} OrbmFrame;

/* Map points for ORBmatcher::SearchByProjection(Frame&, const vector<MapPoint*>&, th, bFarPoints, thFarPoints)
 * (src/ORBmatcher.cc:43-213, mono branch): the tracking fields Frame::isInFrustum wrote (src/Frame.cc:563-571). */
typedef struct OrbmLocalPoints {
    int M;
    const uint8_t* inView;     /* mbTrackInView */
    const uint8_t* bad;        /* isBad() */
    const float* depth;        /* mTrackDepth */
    const float* projX;        /* mTrackProjX */
    const float* projY;        /* mTrackProjY */
    const int32_t* level;      /* mnTrackScaleLevel */
    const float* viewCos;      /* mTrackViewCos */
    const uint8_t* hasObs;     /* Observations() > 0 */
    const uint8_t* descriptors;/* GetDescriptor(), M x 32 */
} OrbmLocalPoints;

/* Last-frame map points for ORBmatcher::SearchByProjection(Frame&, const Frame&, th, bMono=true)
 * (src/ORBmatcher.cc:1676-1887): index i runs over LastFrame.N. */
typedef struct OrbmLastFrame {
    int M;
    const uint8_t* valid;      /* LastFrame.mvpMapPoints[i] != NULL && !LastFrame.mvbOutlier[i] */
    const float* xyz;          /* GetWorldPos(), M x 3 */
    const int32_t* octave;     /* LastFrame.mvKeys[i].octave */
    const float* angle;        /* LastFrame.mvKeysUn[i].angle */
    const uint8_t* hasObs;     /* Observations() > 0 */
    const uint8_t* descriptors;/* GetDescriptor(), M x 32 */
} OrbmLastFrame;

/* Both searches return the reference's `int nmatches` in *nmatches and update, per keypoint of the current frame,
 * match[K] (index of the assigned map point in the input arrays, -1 = none; the host re-attaches MapPoint*) and
 * claimed[K] (assigned point has Observations()>0: such keypoints are skipped by later candidates,
 * src/ORBmatcher.cc:84-86,1747-1749).  Both arrays are in/out (the reference reads Frame::mvpMapPoints). Host pointers. */
int orbm_search_local_map(orbm_handle* h, const OrbmFrame* frame, const OrbmLocalPoints* pts, float th, float nnratio,
                          int bFarPoints, float thFarPoints, int32_t* match, uint8_t* claimed, int* nmatches);
/* Tcw: unit quaternion (w,x,y,z) + translation (CurrentFrame.GetPose()); cam: fx, fy, cx, cy (Pinhole). */
int orbm_search_last_frame(orbm_handle* h, const OrbmFrame* frame, const OrbmLastFrame* last, const float* Tcw7,
                           const float* cam4, float th, int checkOrientation, int32_t* match, uint8_t* claimed,
                           int* nmatches);

/* Frame::isInFrustum (src/Frame.cc:512-574, monocular branch Nleft == -1) with MapPoint::PredictScale
 * (src/MapPoint.cc:531-546) for the M local map points of one frame: what Tracking::SearchLocalPoints
 * (src/Tracking.cc:3346) runs before the local-map SearchByProjection.  Host pointers. */
typedef struct OrbmFrustumIn {
    int M;
    const float* worldPos;        /* GetWorldPos(), M x 3 */
    const float* normal;          /* GetNormal(), M x 3 */
    const float* minDistInv;      /* GetMinDistanceInvariance() = 0.8f * mfMinDistance */
    const float* maxDistInv;      /* GetMaxDistanceInvariance() = 1.2f * mfMaxDistance */
    const float* maxDistance;     /* mfMaxDistance (PredictScale divides it by the current distance) */
    float Rcw[9], tcw[3], Ow[3];  /* mRcw (row-major), mtcw, mOw */
    float cam[4];                 /* fx fy cx cy (Pinhole::project) */
    float minX, minY, maxX, maxY; /* mnMinX ... mnMaxY */
    float mbf;                    /* for mTrackProjXR */
    float logScaleFactor;         /* Frame::mfLogScaleFactor */
    int nScaleLevels;             /* Frame::mnScaleLevels */
    float viewingCosLimit;        /* 0.5 in SearchLocalPoints */
} OrbmFrustumIn;
/* Outputs, one per map point: inView = mbTrackInView (the return value); projX / projY = mTrackProjX / Y (-1 when the
 * point is behind the camera or outside the image; set as soon as the bounds test passes, like the reference);
 * projXR, depth (= |Pc|), level (mnTrackScaleLevel), viewCos are written only for points in view (0 / -1 otherwise: the
 * reference leaves those fields untouched). */
int orbm_frustum_project(orbm_handle* h, const OrbmFrustumIn* in, uint8_t* inView, float* projX, float* projY, float* projXR,
                         float* depth, int32_t* level, float* viewCos);

/* Device-resident, batched last-frame search: `batch` independent streams.  All pointers are device pointers into
 * fixed-capacity slabs (stride = capacity per stream): current frame slabs as written by orbx_extract_batch_device
 * (kps[batch][kcap], desc[batch][kcap][32], nK[batch]); last-frame slabs [batch][mcap]; Tcw [batch][7].
 * match/claimed [batch][kcap] must be initialised by the caller (e.g. -1 / 0); nmatches [batch].  Enqueues on `stream`. */
typedef struct OrbmBatchDevice {
    int batch, kcap, mcap, nlevels;
    const OrbKeyPoint* kps; const uint8_t* desc; const int32_t* nK;
    float minX, minY, maxX, maxY;
    const float* scaleFactors;          /* device, nlevels */
    const int32_t* nM;                  /* [batch] */
    const uint8_t* valid; const float* xyz; const int32_t* octave; const float* angle; const uint8_t* hasObs;
    const uint8_t* mpDesc;
    const float* Tcw7;                  /* [batch][7] */
    float cam[4];
    int resetState;                     /* 1: start from match = -1, claimed = 0 (fill(mvpMapPoints, NULL), Tracking.cc:2876) instead of reading them */
} OrbmBatchDevice;
int orbm_search_last_frame_batch_device(orbm_handle* h, const OrbmBatchDevice* in, float th, int checkOrientation,
                                        int32_t* d_match, uint8_t* d_claimed, int32_t* d_nmatches, void* stream);

/* Same batched search with HOST pointers in `in` and in match / claimed / nmatches (copies in, runs, copies back). */
int orbm_search_last_frame_batch(orbm_handle* h, const OrbmBatchDevice* in, float th, int checkOrientation,
                                 int32_t* match, uint8_t* claimed, int32_t* nmatches);

/* int ORBmatcher::SearchForInitialization(Frame &F1, Frame &F2, vector<cv::Point2f> &vbPrevMatched, vector<int> &vnMatches12,
 * int windowSize) (include/ORBmatcher.h:74, src/ORBmatcher.cc:648-763): the windowed brute-force search of the monocular
 * initialiser (Tracking::MonocularInitialization, src/Tracking.cc:2494-2495, nnratio 0.9, window 100).  F1.K <= max_mappoints and
 * F2.K <= max_keypoints of the handle.  prevMatched [F1.K][2] is vbPrevMatched (in/out, updated like :757-759); matches12 [F1.K] is
 * vnMatches12 (-1 = none); *nmatches is the return value.  Host pointers. */
int orbm_search_for_initialization(orbm_handle* h, const OrbmFrame* F1, const OrbmFrame* F2, float* prevMatched, int windowSize,
                                   float nnratio, int checkOrientation, int32_t* matches12, int* nmatches);

/* Same with the CURRENT FRAME already on the device (in->kps, in->desc, in->nK are device pointers with stride in->kcap, e.g. from
 * orbx_resident_slabs()); everything else as in orbm_search_last_frame_batch (host pointers). */
int orbm_search_last_frame_batch_resident(orbm_handle* h, const OrbmBatchDevice* in, float th, int checkOrientation, int32_t* match,
                                          uint8_t* claimed, int32_t* nmatches);

/* int ORBmatcher::SearchByBoW(KeyFrame* pKF, Frame& F, vector<MapPoint*>& vpMapPointMatches) (include/ORBmatcher.h:68, src/ORBmatcher.cc:223-425,
 * monocular branch): Tracking::TrackReferenceKeyFrame (src/Tracking.cc:2730) and relocalisation.  A frame / keyframe is its keypoints,
 * descriptors and DBoW2::FeatureVector as parallel (node id, feature index) arrays in map order (what orbv_transform_batch returns).
 * kfPoint [KF.N]: 0 = no map point at that keyframe feature, 1 = a map point, 2 = a bad one.  match [F.N] receives the index of the keyframe
 * feature whose map point is assigned to each frame feature (-1 = none), *nmatches the return value.  KF.N <= max_mappoints and
 * F.N <= max_keypoints of the handle.  Host pointers. */
typedef struct OrbmBowFrame {
    int N; const OrbKeyPoint* keypoints; const uint8_t* descriptors;
    int nEntries; const int32_t* fvNode; const int32_t* fvFeature;
} OrbmBowFrame;
int orbm_search_by_bow(orbm_handle* h, const OrbmBowFrame* KF, const uint8_t* kfPoint, const OrbmBowFrame* F, float nnratio, int checkOrientation,
                       int32_t* match, int* nmatches);
/* int ORBmatcher::SearchByBoW(KeyFrame* pKF1, KeyFrame* pKF2, vector<MapPoint*>& vpMatches12) (include/ORBmatcher.h:69, src/ORBmatcher.cc:765-905;
 * LoopClosing): both sides are keyframes (point1 / point2: 0 none, 1 map point, 2 bad), the distance test is strict (bestDist1 < TH_LOW).
 * match12 [KF1.N] = feature of KF2 whose map point goes to vpMatches12[idx1] (-1 = NULL).  KF1.N <= max_mappoints, KF2.N <= max_keypoints. */
int orbm_search_by_bow_kf(orbm_handle* h, const OrbmBowFrame* KF1, const uint8_t* point1, const OrbmBowFrame* KF2, const uint8_t* point2, float nnratio,
                          int checkOrientation, int32_t* match12, int* nmatches);
/* int ORBmatcher::Fuse(KeyFrame* pKF, const vector<MapPoint*>& vpMapPoints, const float th, const bool bRight) (include/ORBmatcher.h:99,
 * src/ORBmatcher.cc:1148-1338, monocular keyframe; LocalMapping::SearchInNeighbors, src/LocalMapping.cc:878-1021): the SEARCH of every map
 * point -- projection with the keyframe pose, KeyFrame::IsInImage, distance / viewing-angle tests, MapPoint::PredictScale, the radius search
 * of KeyFrame::GetFeaturesInArea with the level window and the chi-square gate, best Hamming distance.  bestIdx[i] = keyframe keypoint
 * (-1: none), bestDist[i] its distance (256: none).  Fuse's return value is the number of i with bestDist[i] <= TH_LOW (50); for those the
 * caller applies :1310-1330 in map-point order (Replace the point with fewer observations, or AddObservation + AddMapPoint): that walk
 * mutates the pointer graph and does not feed back into any search.  state[i]: 0 = NULL, 1 = searched, 2 = isBad(), 3 = IsInKeyFrame(pKF).
 * minDistance / maxDistance are the RAW mfMinDistance / mfMaxDistance (the 0.8f / 1.2f of the getters are applied inside).
 * KF.K <= max_keypoints of the handle; M is unbounded.  Host pointers. */
typedef struct OrbmFusePoints {
    int M; const uint8_t* state; const float* worldPos; const float* normal; const float* minDistance; const float* maxDistance;
    const uint8_t* descriptors;
} OrbmFusePoints;
int orbm_fuse_search(orbm_handle* h, const OrbmFrame* KF, const float* invLevelSigma2, float logScaleFactor, const float* Tcw7, const float* Ow3,
                     const float* cam4, const OrbmFusePoints* pts, float th, int32_t* bestIdx, int32_t* bestDist);
/* int ORBmatcher::Fuse(KeyFrame* pKF, Sophus::Sim3f& Scw, const vector<MapPoint*>& vpPoints, float th, vector<MapPoint*>& vpReplacePoint)
 * (include/ORBmatcher.h:102, src/ORBmatcher.cc:1340-1455; LoopClosing::SearchAndFuse): the search of orbm_fuse_search without the chi-square
 * gate.  Tcw7 / Ow3: the caller's decomposition of Scw (:1349-1350: SE3f(Scw.rotationMatrix(), Scw.translation() / Scw.scale()) and its
 * inverse's translation).  state[i]: 1 = searched, 2 = isBad(), 3 = already among the keyframe's map points.  The caller applies :1436-1450
 * (vpReplacePoint / AddObservation) to every bestDist[i] <= TH_LOW. */
int orbm_fuse_search_sim3(orbm_handle* h, const OrbmFrame* KF, float logScaleFactor, const float* Tcw7, const float* Ow3, const float* cam4, const OrbmFusePoints* pts,
                          float th, int32_t* bestIdx, int32_t* bestDist);
/* int ORBmatcher::SearchBySim3(KeyFrame* pKF1, KeyFrame* pKF2, vector<MapPoint*>& vpMatches12, const Sophus::Sim3f& S12, const float th)
 * (include/ORBmatcher.h:95, src/ORBmatcher.cc:1457-1674): both projection searches and the agreement test.  One map-point slot per keyframe
 * feature (pts.M == KF.K; state 0 = none, 1 = good, 2 = bad); pts1.worldPos holds the points of KF1 IN KF2's CAMERA FRAME
 * (S12.inverse() * (T1w * p3Dw), :1507-1508), pts2.worldPos those of KF2 in KF1's (S12 * (T2w * p3Dw), :1586-1587) -- Sophus expressions the
 * caller evaluates; normal is not read.  cam4 = pKF1's intrinsics (the reference uses them for both directions).  pre12[i1] = KF2 feature of
 * an existing vpMatches12[i1] or -1; match12 = pre12 plus the new mutual matches, *nFound = the return value.  Host pointers. */
int orbm_search_by_sim3(orbm_handle* h, const OrbmFrame* KF1, const OrbmFusePoints* pts1, const OrbmFrame* KF2, const OrbmFusePoints* pts2, float logScaleFactor,
                        const float* cam4, float th, const int32_t* pre12, int32_t* match12, int* nFound);
/* int ORBmatcher::SearchForTriangulation(KeyFrame* pKF1, KeyFrame* pKF2, vector<pair<size_t,size_t>>& vMatchedPairs, const bool bOnlyStereo,
 * const bool bCoarse) (include/ORBmatcher.h:84-85, src/ORBmatcher.cc:907-1146; monocular keyframes, bOnlyStereo = false) for ONE new keyframe
 * against nKF2 neighbours in one launch (LocalMapping::CreateNewMapPoints loops over vpNeighKFs, src/LocalMapping.cc:296-344).
 * A keyframe is its undistorted keypoints, descriptors, hasMapPoint[i] (GetMapPoint(i) != NULL) and DBoW2::FeatureVector as (node id, feature
 * index) arrays in map order.  ep2 [nKF2][2]: the epipole pKF2->mpCamera->project(T2w * Cw) (:913-919); F12 [nKF2][9] row-major: the
 * fundamental matrix of Pinhole::epipolarConstrain (src/CameraModels/Pinhole.cpp:109-112) -- both are Eigen / Sophus expressions the caller
 * evaluates.  scaleFactors / levelSigma2: mvScaleFactors / mvLevelSigma2 (the same extractor settings for all keyframes).
 * matches12 [nKF2][KF1.N]: vMatchedPairs scattered (index in KF2 or -1); nmatches [nKF2]: the return values.  Host pointers. */
typedef struct OrbmTriFrame {
    int N; const OrbKeyPoint* keypoints; const uint8_t* descriptors; const uint8_t* hasMapPoint;
    int nEntries; const int32_t* fvNode; const int32_t* fvFeature;
} OrbmTriFrame;
int orbm_search_for_triangulation(orbm_handle* h, const OrbmTriFrame* KF1, int nKF2, const OrbmTriFrame* KF2, const float* scaleFactors, const float* levelSigma2,
                                  int nlevels, const float* ep2, const float* F12, int bCoarse, int checkOrientation, int32_t* matches12, int32_t* nmatches);
/* void MapPoint::ComputeDistinctiveDescriptors() (src/MapPoint.cc:329-403) for nPoints map points at once (LocalMapping::ProcessNewKeyFrame /
 * CreateNewMapPoints call it per point): the observed descriptors of point p are rows obsStart[p] .. obsStart[p+1] of `descriptors`;
 * best[p] = row (relative to obsStart[p]) with the least median Hamming distance to the others, -1 for a point without observations. */
int orbm_distinctive_descriptors(orbm_handle* h, int nPoints, const int32_t* obsStart, const uint8_t* descriptors, int32_t* best);

/* cv::BFMatcher(NORM_HAMMING).knnMatch(query, train, k=2) as used at src/Frame.cc:1144: idx/dist are Q x 2,
 * ordered by (distance, lower train index); missing neighbours are -1.  Host pointers. */
int orbm_bf_knn2(orbm_handle* h, const uint8_t* query, int Q, const uint8_t* train, int T, int32_t* idx, int32_t* dist);
/* ORBmatcher::DescriptorDistance (src/ORBmatcher.cc:2058-2074) for n descriptor pairs a[i], b[i].  Host pointers. */
int orbm_descriptor_distance(orbm_handle* h, const uint8_t* a, const uint8_t* b, int n, int32_t* out);
int orbm_last_launch_count(const orbm_handle* h);

/* ------------------------------------------------------------------------------------------
 * Optimizer::LocalBundleAdjustment (reference include/Optimizer.h:71, src/Optimizer.cc:1116-1498): the numeric
 * core between graph construction and write-back, i.e. optimizer.initializeOptimization(); optimizer.optimize(10)
 * (:1410-1411) with g2o's BlockSolver_6_3 + Levenberg-Marquardt semantics, and the per-edge values the outlier
 * test reads (:1417-1430).  The host shim flattens the pointer graph exactly as :1213-1403 builds it
 * (SURVEY.md 8a' "LBA in"); arrays are in g2o's Hessian order (poses by id, then points by id).
 * ------------------------------------------------------------------------------------------ */
typedef struct lba_handle lba_handle;

typedef struct LbaProblem {
    int nPoses;                  /* local + fixed keyframes (VertexSE3Expmap) */
    const double* poses;         /* nPoses x 7: unit quaternion (w,x,y,z) + translation of Tcw (float->double cast, :1217-1218) */
    const uint8_t* poseFixed;    /* vSE3->setFixed(...): InitKF or member of lFixedCameras */
    const float* cam;            /* nPoses x 4: fx, fy, cx, cy of pKFi->mpCamera (Pinhole, float parameters) */
    int nPoints;                 /* VertexSBAPointXYZ, all marginalized (:1289) */
    const double* points;        /* nPoints x 3 */
    int nEdges;                  /* EdgeSE3ProjectXYZ (mono observations, :1305-1331) */
    const int32_t* edgePoint;    /* vertex 0 */
    const int32_t* edgePose;     /* vertex 1 */
    const double* obs;           /* nEdges x 2: kpUn.pt (:1307-1309) */
    const float* invSigma2;      /* nEdges: pKFi->mvInvLevelSigma2[kpUn.octave] (:1316) */
    double huberDelta;           /* thHuberMono = (float)sqrt(5.991) (:1275,:1321) */
    int iterations;              /* optimizer.optimize(10) */
    double userLambdaInit;       /* solver->setUserLambdaInit(100.0) for inertial maps (:1197-1198), else 0 */
    const volatile uint8_t* stopFlag; /* bool* pbStopFlag (&mbAbortBA, src/LocalMapping.cc:154): one byte, so that the caller's own bool can be
                                       * handed over as it is; polled like SparseOptimizer::terminate(); may be NULL */
} LbaProblem;

typedef struct LbaResult {
    double* poses;               /* nPoses x 7, optimised (fixed ones unchanged) */
    double* points;              /* nPoints x 3 */
    double* edgeChi2;            /* nEdges: e->chi2() as the outlier test sees it (errors of the last evaluated state) */
    uint8_t* edgeDepthPositive;  /* nEdges: e->isDepthPositive() at the final state */
    int iterations;              /* return value of SparseOptimizer::optimize */
    int trials;                  /* total LM trials */
    double lambda, chi2, initialChi2;
    int gpuLaunches;
} LbaResult;

int lba_create(lba_handle** out, int max_poses, int max_points, int max_edges, int device);
/* Same, sized for up to max_batch independent problems solved by one kernel launch (one per stream / local map). */
int lba_create_batch(lba_handle** out, int max_poses, int max_points, int max_edges, int max_batch, int device);
void lba_destroy(lba_handle* h);
/* Host pointers in and out; edges may come in any order (per-edge results come back in the caller's order); the index structures of
 * BlockSolver::buildStructure are derived on the device.  Returns ORB_OK, or ORB_ERR_ARG for malformed graphs (index out of range,
 * the same point observed twice by one free keyframe, no free vertex), ORB_ERR_CAPACITY for a problem larger than the handle. */
int lba_solve(lba_handle* h, const LbaProblem* problem, LbaResult* result);
int lba_solve_batch(lba_handle* h, int count, const LbaProblem* problems, LbaResult* results);
/* Split form of lba_solve_batch: upload the flattened graphs once (host pointers), run the whole LM loop for all of
 * them from the uploaded initial estimates on `stream` (device-resident, asynchronous, repeatable), download results.
 * Ordering: lba_download_batch and the next lba_upload_batch wait (on the device, through an event) for the last
 * lba_run_batch_device whatever stream it was given, so no synchronisation is required from the caller in between. */
int lba_upload_batch(lba_handle* h, int count, const LbaProblem* problems);
int lba_run_batch_device(lba_handle* h, void* stream);
int lba_download_batch(lba_handle* h, int count, LbaResult* results);
/* Thread-block-cluster size (CTAs per problem) the last run used. */
int lba_last_cluster_size(const lba_handle* h);
/* Force the cluster size (1..8 CTAs per problem; 0 = automatic: the largest for which every problem's cluster is resident at once). */
int lba_set_cluster_size(lba_handle* h, int ctas);
/* Device-side phase timers (ns, CTA 0) of problem i of the last downloaded run: errors (first iteration), build_points,
 * build_poses, point_prep (after a rejected step), schur_partial, ldlt, schur_combine, pose_trial, points_trial (back-substitution +
 * update + residuals), spare. */
int lba_get_phase_ns(const lba_handle* h, int i, double* ns10);

/* ------------------------------------------------------------------------------------------
 * int Optimizer::PoseOptimization(Frame* pFrame) (reference include/Optimizer.h:62, src/Optimizer.cc:814-1114), monocular
 * branch, for `count` frames at once (one CTA per frame).  Per frame f (slots of `cap` entries, N[f] used):
 *   pose7 [count][7]: pFrame->GetPose() as unit quaternion (w,x,y,z) + translation, cast to double (:829-830)
 *   cam4  [count][4]: fx, fy, cx, cy;  Xw [count][cap][3]: pMP->GetWorldPos() (:884);  obs [count][cap][2]: mvKeysUn[i].pt
 *   invSigma2 [count][cap]: mvInvLevelSigma2[octave] (:877);  huberDelta = (float)sqrt(5.991)
 * Out: poseOut [count][7] (pFrame->SetPose), outlier [count][cap] (pFrame->mvbOutlier), nInliers[count] = the return value
 * nInitialCorrespondences - nBad.  Host pointers.
 * ------------------------------------------------------------------------------------------ */
int pose_optimization_batch(int count, int cap, const int32_t* N, const double* pose7, const float* cam4, const double* Xw,
                            const double* obs, const float* invSigma2, double huberDelta, double* poseOut, uint8_t* outlier,
                            int32_t* nInliers, int device);


/* ------------------------------------------------------------------------------------------
 * Inertial edges: what Optimizer::LocalInertialBA (reference include/Optimizer.h:96, src/Optimizer.cc:2383-2958)
 * and PoseInertialOptimizationLast{KeyFrame,Frame} (:4491, :4875) evaluate in every iteration, batched over all streams of a GPU.
 * The 15-DoF block solver around them is not part of this library yet; these entry points return residuals, Jacobians, chi2 and
 * robust weights in g2o's conventions.  Host pointers.
 * ------------------------------------------------------------------------------------------ */
/* One IMU::Preintegrated (reference include/ImuTypes.h:159-240) as 292 floats:
 *   [0] dT | [1..9] dR | [10..12] dV | [13..15] dP | [16..24] JRg | [25..33] JVg | [34..42] JVa | [43..51] JPg | [52..60] JPa |
 *   [61..66] b = (bax, bay, baz, bwx, bwy, bwz) | [67..291] C (15 x 15 covariance, row-major).  3 x 3 blocks are row-major. */
#define IMU_PREINT_FLOATS 292
/* Preintegrated::Initialize(b) + IntegrateNewMeasurement (src/ImuTypes.cc:177-240) for `count` intervals (one per stream / keyframe pair):
 * acc, gyr [count][maxMeas][3], dt [count][maxMeas], nMeas [count], bias6 [count][6] = (bax, bay, baz, bwx, bwy, bwz);
 * noise4 = (ng, na, ngw, naw) as given to IMU::Calib::Set (:395-407).  preint [count][IMU_PREINT_FLOATS]. */
int imu_preintegrate_batch(int count, const int32_t* nMeas, int maxMeas, const float* acc, const float* gyr, const float* dt, const float* bias6,
                           const float* noise4, float* preint, int device);
/* EdgeInertial's information matrix (src/G2oTypes.cc:499-507: inverse of C.block<9,9>, symmetrised, eigenvalues < 1e-12 clamped) and the
 * EdgeGyroRW / EdgeAccRW informations (src/Optimizer.cc:551,559: inverses of C.block<3,3>(9,9) / (12,12)).  info9 [count][81], infoG / infoA [count][9]. */
int imu_information_batch(int count, const float* preint, double* info9, double* infoG, double* infoA, int device);
/* EdgeInertial::computeError + linearizeOplus (src/G2oTypes.cc:514-594) for `count` edges; edge e uses preint[e].
 * states36 [count][36]: Rwb1 9 | twb1 3 | v1 3 | gyro bias 3 | acc bias 3 | Rwb2 9 | twb2 3 | v2 3 (estimates of the six vertices).
 * err9 [count][9] = (er, ev, ep); J9x24 [count][9][24] (may be NULL), columns: pose 1 (rotation 3, translation 3; tangent of
 * ImuCamPose::Update) | velocity 1 | gyro bias | acc bias | pose 2 | velocity 2.  With info9: chi2 [count] = e^T Omega e and, if rho is
 * given, the Huber weight rho'(chi2) for huberDelta (sqrt(16.92) in LocalInertialBA, src/Optimizer.cc:540-542; <= 0: no kernel). */
int imu_inertial_edges(int count, const float* preint, const double* states36, const double* info9, double huberDelta, double* err9, double* J9x24,
                       double* chi2, double* rho, int device);
/* int Optimizer::PoseInertialOptimizationLastKeyFrame(Frame* pFrame, bool bRecInit) (include/Optimizer.h:66, src/Optimizer.cc:4491-4873),
 * monocular frame: the inertial pose optimiser Tracking::TrackLocalMap runs per frame once the IMU is initialised and the map was updated
 * (src/Tracking.cc:2985-2994), for `count` frames at once (one CTA per frame).  Per frame f (slots of `cap` entries, N[f] used):
 *   Xw [count][cap][3] pMP->GetWorldPos(); obs [count][cap][2] mvKeysUn[i].pt; invSigma2 [count][cap] mvInvLevelSigma2[octave] / unc2;
 *   trackDepth [count][cap] pMP->mTrackDepth (the 10 m "close" rule, :4730); cam4 [count][4]; extrinsics24: Rcb 9 | tcb 3 | Rbc 9 | tbc 3;
 *   preint [count][IMU_PREINT_FLOATS]: pFrame->mpImuPreintegrated (from the last keyframe);
 *   kfState21 / state21 [count][21]: Rwb 9 | twb 3 | velocity 3 | gyro bias 3 | acc bias 3 of the last keyframe (fixed) and of the frame
 *   (in: VertexPose / Velocity / GyroBias / AccBias(pFrame); out: what SetImuPoseVelocity / mImuBias receive, :4813-4817).
 * Out: outlier [count][cap] = mvbOutlier, H15 [count][225] = the Hessian handed to ConstraintPoseImu (:4819-4870, row-major, order pose 6 |
 * velocity | gyro bias | acc bias), ret [count] = nInitialCorrespondences - nBad.  Host pointers. */
int pose_inertial_optimization_last_kf_batch(int count, int cap, const int32_t* N, const float* Xw, const float* obs, const float* invSigma2, const float* trackDepth,
                                             const float* cam4, const double* extrinsics24, const float* preint, const double* kfState21, double* state21, int bRecInit,
                                             uint8_t* outlier, double* H15, int32_t* ret, int device);
/* int Optimizer::PoseInertialOptimizationLastFrame(Frame* pFrame, bool bRecInit) (include/Optimizer.h:67, src/Optimizer.cc:4875-5289), monocular
 * frame: the variant Tracking::TrackLocalMap runs when the map was NOT updated since the last frame (src/Tracking.cc:2985-2994).  The previous
 * frame's pose / velocity / biases are free too (30 unknowns) and held by EdgePriorPoseImu: prior21 [count][21] = Rwb | twb | vwb | bg | ba and
 * priorH [count][225] = pFp->mpcpi->H (what ConstraintPoseImu's constructor made of the Hessian the previous call returned: symmetrised,
 * eigenvalues < 1e-12 clamped).  preintFrame = pFrame->mpImuPreintegratedFrame (EdgeInertial), preintKF = pFrame->mpImuPreintegrated (only the
 * covariance blocks of the two random-walk edges, :5068-5078).  prevState21 in: VertexPose / Velocity / GyroBias / AccBias(pFp); out: their
 * estimates afterwards (the reference does not write them back).  H15 = Marginalize(H, 0, 14).block<15,15>(15,15) (:5268-5270), the Hessian
 * handed to the new ConstraintPoseImu.  Everything else as in pose_inertial_optimization_last_kf_batch. */
int pose_inertial_optimization_last_frame_batch(int count, int cap, const int32_t* N, const float* Xw, const float* obs, const float* invSigma2, const float* trackDepth,
                                                const float* cam4, const double* extrinsics24, const float* preintFrame, const float* preintKF, const double* prior21,
                                                const double* priorH, double* prevState21, double* state21, int bRecInit, uint8_t* outlier, double* H15, int32_t* ret,
                                                int device);
/* EdgeMono (include/G2oTypes.h:342-385, src/G2oTypes.cc:349-373) over VertexPose = ImuCamPose (body pose + camera extrinsics, src/G2oTypes.cc:148-220). */
typedef struct ImuMonoEdges {
    int nPoses;  const double* poses;       /* [nPoses][12]: Rwb 9 | twb 3 */
    const double* extrinsics;               /* [24]: Rcb 9 | tcb 3 | Rbc 9 | tbc 3 (camera 0, IMU::Calib mTcb / mTbc) */
    const float* cam;                       /* [nPoses][4]: fx fy cx cy */
    int nPoints; const double* points;      /* [nPoints][3] */
    int nEdges;  const int32_t* edgePoint; const int32_t* edgePose; const double* obs /* [nEdges][2] */; const float* invSigma2 /* [nEdges] */;
    double huberDelta;                      /* thHuberMono = sqrt(5.991); <= 0: no kernel */
} ImuMonoEdges;
/* err2 [nEdges][2]; Jpoint2x3 [nEdges][6] / Jpose2x6 [nEdges][12] (both or neither may be NULL); chi2, rho (may be NULL) [nEdges];
 * depthPositive [nEdges] = EdgeMono::isDepthPositive(). */
int imu_mono_edges(const ImuMonoEdges* in, double* err2, double* Jpoint2x3, double* Jpose2x6, double* chi2, double* rho, uint8_t* depthPositive, int device);


/* void Optimizer::LocalInertialBA(KeyFrame* pKF, bool* pbStopFlag, Map* pMap, int& num_fixedKF, int& num_OptKF, int& num_MPs, int& num_edges,
 * bool bLarge, bool bRecInit) (include/Optimizer.h:62, src/Optimizer.cc:2383-2958; call site LocalMapping::Run, src/LocalMapping.cc:129-151),
 * monocular-inertial: the numeric core between the graph set-up (:2385-2833, the caller's walk over keyframes / map points / observations) and the
 * write-back (:2890-2957), for `count` local maps at once (one persistent CTA per map; LBA-style "replicas" across streams).
 *   keyframes [0, nOpt): the temporal window vpOptimizableKFs (VertexPose + VertexVelocity + VertexGyroBias + VertexAccBias, all free);
 *   keyframes [nOpt, nKF): lFixedKeyFrames (the window's predecessor first when it is linked by an EdgeInertial).
 *   kfState21 [nKF][21]: Rwb 9 | twb 3 | velocity 3 | gyro bias 3 | acc bias 3 (GetImuRotation / GetImuPosition / GetVelocity / GetImuBias, as doubles);
 *   kfTcw12 [nKF][12]: Rcw 9 | tcw 3 = the keyframe's own camera pose (GetRotation / GetTranslation): ImuCamPose(KeyFrame*) loads both and uses the
 *   camera pose as it is until the vertex is first updated (src/G2oTypes.cc:25-47, :213-221);  cam4 [nKF][4];  extrinsics24: Rcb 9 | tcb 3 | Rbc 9 | tbc 3.
 *   inertial edge i (:2593-2656): keyframes ieKf1[i] (mPrevKF) -> ieKf2[i], preint [nInertial][IMU_PREINT_FLOATS] = pKFi->mpImuPreintegrated,
 *   ieRobust[i] = 1 for the Huber kernel sqrt(16.92) (i == N-1 || bRecInit), ieInfoScale[i] = 1e-2 for i == N-1, else 1.  Every inertial edge
 *   brings its EdgeGyroRW and EdgeAccRW (informations from the preintegration covariance).
 *   points3 [nPoints][3] = GetWorldPos(), trackDepth [nPoints] = mTrackDepth;  mono edge e (:2737-2763): edgePoint[e], edgeKf[e], obs2 = mvKeysUn[].pt,
 *   invSigma2 = mvInvLevelSigma2[octave] / unc2; at most one edge per (point, keyframe) (monocular: no right-camera observations).
 *   iterations = opt_it (10, or 4 with bLarge), lambdaInit = setUserLambdaInit's value (1e0, or 1e-2 with bLarge).
 * Results (host pointers, per problem): the optimised kfState21 / kfTcw12 / points3 (what SetPose / SetVelocity / SetNewBias / SetWorldPos receive
 * after the casts of :2905-2950), erase [nEdges] = 1 for the (keyframe, point) pairs of vToErase (:2848-2862), edgeChi2 [nEdges] = e->chi2(),
 * stats8 = err, err_end (:2837-2839, as floats), failed (the "FAIL LOCAL-INERTIAL BA" test :2884: states and points are then returned unchanged),
 * final lambda, LM trials, optimize()'s iteration count, nanoseconds the map's CTA spent in the solver.  iterationsOut [count] (may be NULL). */
typedef struct LocalInertialBAProblem {
    int32_t nKF, nOpt;
    const double* kfState21; const double* kfTcw12; const float* cam4; const double* extrinsics24;
    int32_t nInertial; const int32_t* ieKf1; const int32_t* ieKf2; const float* preint; const uint8_t* ieRobust; const double* ieInfoScale;
    int32_t nPoints; const double* points3; const float* trackDepth;
    int32_t nEdges; const int32_t* edgePoint; const int32_t* edgeKf; const double* obs2; const float* invSigma2;
    int32_t iterations; int32_t bLarge; double lambdaInit;
} LocalInertialBAProblem;
typedef struct LocalInertialBAResult {
    double* kfState21; double* kfTcw12; double* points3; uint8_t* erase; double* edgeChi2; double* stats8;
    double* profile8;   /* may be NULL; diagnostic: nanoseconds in errors | buildSystem | Dinv, Y | Schur | LDL^T | point back-substitution | update / pop | rest */
} LocalInertialBAResult;
int local_inertial_ba_batch(int count, const LocalInertialBAProblem* problems, const LocalInertialBAResult* results, int32_t* iterationsOut, int device);

/* ------------------------------------------------------------------------------------------
 * DBoW2 transform: Frame::ComputeBoW / KeyFrame::ComputeBoW (reference src/Frame.cc:738-745) call
 * ORBVocabulary::transform(vCurrentDesc, mBowVec, mFeatVec, 4) = DBoW2::TemplatedVocabulary<FORB::TDescriptor, FORB>::transform
 * (Thirdparty/DBoW2/DBoW2/TemplatedVocabulary.h:1127-1259).  The vocabulary tree is handed over once as flat arrays (the host shim
 * flattens m_nodes after loadFromTextFile) and stays in HBM; a call transforms the descriptors of `batch` frames.
 * ------------------------------------------------------------------------------------------ */
typedef struct orbv_handle orbv_handle;
typedef struct OrbVocabulary {
    int nNodes;                   /* m_nodes.size(); node 0 is the root */
    int L;                        /* m_L (6 for ORBvoc) */
    int weighting;                /* m_weighting: 0 TF_IDF (ORB-SLAM3), 1 TF, 2 IDF, 3 BINARY */
    int norm;                     /* what m_scoring_object->mustNormalize gives: 0 none, 1 L1 (L1_NORM, ORB-SLAM3), 2 L2 */
    const int32_t* childStart;    /* [nNodes + 1]: children of node n are children[childStart[n] .. childStart[n+1]) */
    const int32_t* children;      /* node ids in the order of m_nodes[n].children */
    const uint8_t* descriptors;   /* [nNodes][32]: m_nodes[n].descriptor */
    const double* weight;         /* [nNodes]: m_nodes[n].weight (idf of a word; 0 = stopped) */
    const int32_t* wordId;        /* [nNodes]: m_nodes[n].word_id for leaves, -1 otherwise */
} OrbVocabulary;
int orbv_create(orbv_handle** out, const OrbVocabulary* voc, int device);
void orbv_destroy(orbv_handle* h);
/* transform() for `batch` frames; frame f has n[f] descriptors in desc [batch][cap][32].  Outputs in the iteration order of the reference's
 * std::map containers: BowVector as wordId / wordValue [batch][cap] (nWords[f] entries, ascending word id; values are DBoW2's doubles bit
 * for bit), FeatureVector as parallel arrays fvNode / fvFeature [batch][cap] (nEntries[f] entries: ascending node id, feature indices
 * ascending within a node).  levelsup = 4 in the reference.  Host pointers. */
int orbv_transform_batch(orbv_handle* h, int batch, const uint8_t* desc, const int32_t* n, int cap, int levelsup, int32_t* wordId, double* wordValue,
                         int32_t* nWords, int32_t* fvNode, int32_t* fvFeature, int32_t* nEntries);
/* Device form: all pointers device pointers; d_scratch* are [batch][cap] work arrays owned by the caller; enqueues on `stream`. */
int orbv_transform_batch_device(orbv_handle* h, int batch, const uint8_t* d_desc, const int* d_n, int cap, int levelsup, int* d_scratchWord,
                                double* d_scratchWeight, int* d_scratchNode, int* d_wordId, double* d_wordValue, int* d_nWords, int* d_fvNode,
                                int* d_fvFeature, int* d_nEntries, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* ORB_B200_H */
