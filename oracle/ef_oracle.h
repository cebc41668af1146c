/*
 * ef_oracle.h -- CPU ORACLE for the detectAndCompute hot path.  TEST INFRASTRUCTURE ONLY.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
 * load this library; nothing under cuda-efficient-features_b200/ links, imports or executes it.
 *
 * What it restates (reference = fixstars/cuda-efficient-features @ 761db2b):
 *   descriptors : modules/efficient_features/src/bad.cpp, hash_sift.cpp (the reference's CPU
 *                 ground truth).  PINNED: tests/test_oracle_vs_ref.py checks this restatement
 *                 bit-for-bit against oracle/_ref (the reference's own bad.cpp/hash_sift.cpp
 *                 compiled unmodified against oracle/shim/), and tests/golden/ holds vectors
 *                 generated from oracle/_ref.
 *   detector    : the reference has NO CPU detector.  This is a CPU restatement of its CUDA
 *                 detector (cuda_fast.cu, cuda_efficient_features.cu, cuda_efficient_features.cpp)
 *                 with the deterministic rules of DESIGN.md (no candidate cap, raster output
 *                 order, total order for top-K ties) and the FMA contraction pattern nvcc emits
 *                 for the reference source.  PINNED on the GPU box for FAST-9, the Harris response
 *                 (bit for bit), the radius NMS, limitPoints and scalePoints, and to <= 2 ulp for the IC
 *                 angle (the reference calls CUDA atan2f): tests/test_gpu_reference_kernels.py runs the
 *                 reference's own cuda_fast.cu / cuda_efficient_features.cu, compiled unmodified into
 *                 oracle/_ref/libef_ref_cuda.so against oracle/shim_cuda/, next to this restatement and
 *                 the product kernels.  PARITY UNPINNED remains for the two OpenCV-CUDA library calls the
 *                 detector depends on (cv::cuda::resize, Gaussian filter): third-party, not vendored, and
 *                 no reference test covers them.
 */
#ifndef EF_ORACLE_H
#define EF_ORACLE_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define EFO_MAX_LEVELS 16

enum { EFO_BAD_256 = 0, EFO_BAD_512 = 1, EFO_HASH_SIFT_256 = 2, EFO_HASH_SIFT_512 = 3 };

typedef struct efo_params {
    int nfeatures;      /* 5000  */
    float scale_factor; /* 1.2f  */
    int nlevels;        /* 8     */
    int first_level;    /* 0     */
    int fast_threshold; /* 20    */
    int nonmax_radius;  /* 15    */
    int desc_type;      /* EFO_* */
} efo_params;

/* descriptor input keypoint == cv::Vec4f(pt.x, pt.y, size, angle), cuda_efficient_features.cpp:125 */
typedef struct efo_kpt { float x, y, size, angle; } efo_kpt;

/* detector output keypoint, one column of the 5xN matrix (cuda_efficient_features.h:32-37) */
typedef struct efo_keypoint {
    short x, y;      /* LOCATION_ROW (image coordinates after scalePoints) */
    float response;  /* RESPONSE_ROW */
    float angle;     /* ANGLE_ROW, degrees [0,360) */
    int octave;      /* OCTAVE_ROW */
    float size;      /* SIZE_ROW = scale*31 */
    short lx, ly;    /* level coordinates (not part of the reference output; for stage tests) */
} efo_keypoint;

void efo_set_threads(int n); /* OpenMP threads for the all-cores baseline; 1 = faithful scalar */
int efo_get_max_threads(void);

/* cuda_efficient_features.cpp:136-157 -- level sizes and scales (float accumulation) */
void efo_level_geometry(int w, int h, float scale_factor, int nlevels, int* ws, int* hs, float* scales);
/* cuda_efficient_features.cpp:159-174 */
void efo_level_quotas(int nfeatures, float scale_factor, int nlevels, int* quotas);

/* cv::cuda::resize INTER_LINEAR as recalled in SURVEY Appendix A.1 (parity unpinned) */
void efo_resize_linear(const uint8_t* src, int sw, int sh, size_t spitch, uint8_t* dst, int dw, int dh, size_t dpitch);
/* cv::cuda Gaussian 7x7 sigma 2 REFLECT_101, SURVEY Appendix A.2 (parity unpinned) */
void efo_gaussian_blur7(const uint8_t* src, int w, int h, size_t spitch, uint8_t* dst, size_t dpitch);

/* cuda_fast.cu:36-222 -- 1 if (x,y) passes FAST-9/16 at threshold th (no border test) */
int efo_fast_is_corner(const uint8_t* img, size_t pitch, int x, int y, int th);
/* cuda_efficient_features.cu:99-139 with nvcc's contraction */
float efo_harris_response(const uint8_t* img, size_t pitch, int x, int y);
/* cuda_efficient_features.cu:141-172, atan2 evaluated in double and rounded once */
float efo_ic_angle(const uint8_t* img, size_t pitch, int x, int y);

/* Dense per-level stage outputs for stage-by-stage parity tests: response map (w*h floats, -INF where
 * (x,y) is not a FAST corner inside the 15-px border). Returns the number of corners. */
long efo_score_map(const uint8_t* img, int w, int h, size_t pitch, int th, float* resp);
/* radius NMS on a dense response map (cuda_efficient_features.cu:62-97,202-216): writes survivors in
 * raster order; returns count (may exceed cap; only cap are written). */
long efo_radius_nms(const float* resp, int w, int h, int radius, short* xs, short* ys, float* rs, long cap);

/* Whole detector on one frame. out has room for cap keypoints; returns N (levels concatenated in
 * ascending octave, raster order inside a level).  If level_counts != NULL it receives per-level
 * {corners, survivors, selected} triples (3*nlevels longs). */
int efo_detect(const uint8_t* img, int w, int h, size_t pitch, const efo_params* p,
               efo_keypoint* out, int cap, long* level_counts);

/* bad.cpp:254-405. nbits in {256,512}. desc: n x nbits/8 */
void efo_bad_compute(const uint8_t* img, int w, int h, size_t pitch, const efo_kpt* kpts, int n,
                     float scale_factor, int nbits, uint8_t* desc);
/* hash_sift.cpp:333-351: n x 129 floats (bias 1 first, then 128 uchar-valued floats) */
void efo_hashsift_features(const uint8_t* img, int w, int h, size_t pitch, const efo_kpt* kpts, int n,
                           float cropping_scale, float* resp129);
/* hash_sift.cpp:111-138: the 32x32 rectified patch of one keypoint */
void efo_hashsift_patch(const uint8_t* img, int w, int h, size_t pitch, const efo_kpt* kpt,
                        float cropping_scale, uint8_t* patch1024);
/* hash_sift.cpp:353-378 with cv::gemm defined as float32(sum in double, ascending k).
 * proj may be NULL; otherwise n x nbits floats (pre-binarisation). */
void efo_hashsift_project(const float* resp129, int n, int nbits, uint8_t* desc, float* proj);
void efo_hashsift_compute(const uint8_t* img, int w, int h, size_t pitch, const efo_kpt* kpts, int n,
                          float cropping_scale, int nbits, uint8_t* desc, float* proj);

/* cuda_efficient_features.cpp:225-321: detect, blur each level, describe in level coordinates with
 * size 31 and the IC angle, scale points.  desc: cap x desc_bytes. Returns N. */
int efo_detect_and_compute(const uint8_t* img, int w, int h, size_t pitch, const efo_params* p,
                           efo_keypoint* out, uint8_t* desc, int cap, long* level_counts);

/* Debug: builds the pyramid (and optionally the blurred pyramid); level l is written densely
 * (pitch = ws[l]) at levels + offsets[l]. Returns total bytes needed if levels == NULL. */
size_t efo_build_pyramid(const uint8_t* img, int w, int h, size_t pitch, float scale_factor, int nlevels,
                         int blurred, uint8_t* levels, size_t* offsets);

/* The synthetic-frame generator shared by tests and bench (counter-based hash). */
void efo_synth_frame(uint32_t seed, uint32_t frame, int w, int h, size_t pitch, uint8_t* dst);

#ifdef __cplusplus
}
#endif
#endif
