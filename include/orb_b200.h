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

/* Debug/inspection taps used by the parity tests (mvImagePyramid is a public member of the reference class,
 * include/ORBextractor.h:84).  Copies level `level` of frame `frame` of the last call, unbordered, tightly packed. */
int orbx_get_level_size(const orbx_handle* h, int level, int* width, int* height);
int orbx_copy_level(orbx_handle* h, int frame, int level, int blurred, uint8_t* dst);
/* FAST candidates of one level of the last call as (x, y, score) int triples in the order the reference
 * feeds them to DistributeOctTree (src/ORBextractor.cc:863-871). Returns the count or <0. */
int orbx_copy_candidates(orbx_handle* h, int frame, int level, int* xys, int cap);
/* Number of kernels launched by the last extract call (for bench.py's gpu_launches). */
int orbx_last_launch_count(const orbx_handle* h);

#ifdef __cplusplus
}
#endif
#endif /* ORB_B200_H */
