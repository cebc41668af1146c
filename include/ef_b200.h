/*
 * ef_b200.h -- C ABI of the B200-native detectAndCompute hot path (libef_b200.so).
 *
 * Plain C, POD only: no OpenCV, torch or CUDA types in the signatures (streams are passed as
 * void* holding a cudaStream_t).  Every entry point names the reference interface it replaces
 * (fixstars/cuda-efficient-features @ 761db2b, paths relative to modules/cuda_efficient_features/).
 *
 * Contract
 *   - All work is enqueued on the caller's stream; no entry point ending in _async synchronises
 *     with the host, allocates device memory or launches on another stream.  Counts stay on the
 *     device (the reference blocks twice per pyramid level, src/cuda_fast.cu:241-243,
 *     src/cuda_efficient_features.cu:337-339).
 *   - Outputs have fixed capacity: keypoints are a 5 x capacity float matrix in the reference's
 *     row layout (include/cuda_efficient_features.h:32-37), capacity = nfeatures columns;
 *     *d_count receives N <= nfeatures, columns [0,N) are valid.
 *   - Returns EF_OK or an ef_status error; ef_last_error_string() describes the last failure.
 *   - One handle per concurrent stream (the reference object is equally stateful,
 *     src/cuda_efficient_features.cpp:391-403).  Parameters are per handle and per device: no
 *     process-global __constant__ state (the reference's BAD tables are global, src/cuda_bad.cu:49-50).
 *   - There is no CPU fallback: if the CUDA device or kernels are unavailable every call fails.
 */
#ifndef EF_B200_H
#define EF_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define EF_MAX_LEVELS 16

/* EfficientFeatures::DescriptorType, include/cuda_efficient_features.h:39-45 */
typedef enum ef_desc_type { EF_BAD_256 = 0, EF_BAD_512 = 1, EF_HASH_SIFT_256 = 2, EF_HASH_SIFT_512 = 3 } ef_desc_type;

/* rows of the keypoint matrix, include/cuda_efficient_features.h:32-37 */
enum { EF_LOCATION_ROW = 0, EF_RESPONSE_ROW = 1, EF_ANGLE_ROW = 2, EF_OCTAVE_ROW = 3, EF_SIZE_ROW = 4, EF_ROWS_COUNT = 5 };

typedef enum ef_status {
    EF_OK = 0,
    EF_ERR_BAD_ARG = 1,     /* CV_Assert / CV_Error(StsBadArg) in the reference */
    EF_ERR_CUDA = 2,        /* CUDA runtime failure (the reference only prints these, src/cuda_macro.h:23-28) */
    EF_ERR_CAPACITY = 3,    /* image / batch larger than the handle was created for */
    EF_ERR_UNSUPPORTED = 4
} ef_status;

/* EfficientFeatures::create(nfeatures, scaleFactor, nlevels, firstLevel, fastThreshold, nonmaxRadius, dtype)
 * (include/cuda_efficient_features.h:47-48) plus the capacities the workspace is planned for. */
typedef struct ef_params {
    int nfeatures;        /* 5000 */
    float scale_factor;   /* 1.2f */
    int nlevels;          /* 8    */
    int first_level;      /* 0    */
    int fast_threshold;   /* 20   */
    int nonmax_radius;    /* 15   */
    int desc_type;        /* ef_desc_type; reference default HASH_SIFT_256 */
    float desc_scale;     /* BAD scaleFactor / HashSIFT croppingScale of the compute-only API; the
                             detectAndCompute path always uses 1 (src/cuda_efficient_features.cpp:48-69) */
    int max_width;        /* largest frame the workspace is planned for */
    int max_height;
    int max_batch;        /* frames per batched call */
    int max_keypoints;    /* capacity of the compute-only API (>= nfeatures) */
    int device;           /* CUDA device ordinal */
    int flags;            /* EF_FLAG_*; 0 = full detectAndCompute handle */
} ef_params;

/* ef_params.flags.  EF_FLAG_COMPUTE_ONLY: the handle serves ef_compute_async / ef_compute_rows_async only (cv::cuda::BAD,
 * cv::cuda::HashSIFT: the reference describers hold nothing but their tables, src/cuda_bad.cpp:36-44): no detection
 * workspace and no host-API staging are allocated; the detect entry points return EF_ERR_UNSUPPORTED. */
enum { EF_FLAG_COMPUTE_ONLY = 1 };

typedef struct ef_handle ef_handle;

void ef_default_params(ef_params* p);

/* replaces EfficientFeatures::create / EfficientFeaturesImpl ctor (src/cuda_efficient_features.cpp:188-195,406-411);
 * allocates the whole workspace once (the reference grows DeviceBuffers lazily, src/device_buffer.cpp:42-52). */
int ef_create(const ef_params* params, ef_handle** out);
void ef_destroy(ef_handle* h);

/* the 7 setter/getter pairs, include/cuda_efficient_features.h:78-97.  Capacities cannot change. */
typedef enum ef_param_id {
    EF_PARAM_MAX_FEATURES = 0, EF_PARAM_SCALE_FACTOR = 1, EF_PARAM_NLEVELS = 2, EF_PARAM_FIRST_LEVEL = 3,
    EF_PARAM_FAST_THRESHOLD = 4, EF_PARAM_NONMAX_RADIUS = 5, EF_PARAM_DESCRIPTOR_TYPE = 6, EF_PARAM_DESC_SCALE = 7
} ef_param_id;
int ef_set_param(ef_handle* h, int id, double value);
int ef_get_param(const ef_handle* h, int id, double* value);

size_t ef_workspace_bytes(const ef_handle* h);
int ef_descriptor_size(const ef_handle* h);            /* descriptorSize(): nbits/8, src/cuda_efficient_features.cpp:351 */
const char* ef_last_error_string(const ef_handle* h);
const char* ef_version(void);

/* replaces EfficientFeaturesImpl::detectAndComputeAsync with GpuMat arguments
 * (src/cuda_efficient_features.cpp:225-321); d_desc == NULL gives detectAsync (:215-218).
 *   d_img       CV_8UC1 device image, `pitch` bytes per row
 *   d_kpts      5 x nfeatures floats, `kpts_pitch` bytes per row (row 0 = short2, row 3 = int)
 *   d_desc      nfeatures x descriptorSize bytes, `desc_pitch` bytes per row, or NULL
 *   d_count     one int on the device: number of valid columns/rows */
int ef_detect_and_compute_async(ef_handle* h, const uint8_t* d_img, size_t pitch, int width, int height,
                                float* d_kpts, size_t kpts_pitch, uint8_t* d_desc, size_t desc_pitch,
                                int* d_count, void* stream);

/* Batched form (new; the reference processes one frame per call): nframes <= max_batch frames of equal
 * size.  Frame f reads d_imgs + f*img_stride and writes d_kpts + f*kpts_stride (bytes),
 * d_desc + f*desc_stride, d_counts[f]. */
int ef_detect_and_compute_batch_async(ef_handle* h, int nframes, const uint8_t* d_imgs, size_t img_stride, size_t pitch,
                                      int width, int height, float* d_kpts, size_t kpts_stride, size_t kpts_pitch,
                                      uint8_t* d_desc, size_t desc_stride, size_t desc_pitch, int* d_counts, void* stream);

/* replaces EfficientDescriptorsAsync::compute with std::vector<KeyPoint> (src/cuda_bad.cpp:72-75,
 * src/cuda_hash_sift.cpp:139-142; keypoints packed as (pt.x, pt.y, size, angle),
 * src/cuda_efficient_features.cpp:116-128).  d_kpts_xysa: n x 4 floats on the device. */
int ef_compute_async(ef_handle* h, const uint8_t* d_img, size_t pitch, int width, int height,
                     const float* d_kpts_xysa, int n, uint8_t* d_desc, size_t desc_pitch, void* stream);

/* replaces EfficientFeaturesImpl::computeAsync with a 5 x N GpuMat of keypoints
 * (src/cuda_efficient_features.cpp:220-223 -> getKeypointsMat :104-114 -> convertKeypointsKernel,
 * src/cuda_efficient_features.cu:250-263: only LOCATION and ANGLE rows are read, size is forced to 31). */
int ef_compute_rows_async(ef_handle* h, const uint8_t* d_img, size_t pitch, int width, int height,
                          const float* d_kpts5, size_t kpts_pitch, int n, uint8_t* d_desc, size_t desc_pitch, void* stream);

/* replaces detectAndCompute / detect with cv::Mat arguments (src/cuda_efficient_features.cpp:197-213,
 * upload :76, download :316-320): HOST buffers; does the H2D copy, the device pipeline and the D2H
 * copies on `stream`, then synchronises once.  h_desc may be NULL.  *h_count = N. */
int ef_detect_and_compute_host(ef_handle* h, const uint8_t* h_img, size_t pitch, int width, int height,
                               float* h_kpts5, uint8_t* h_desc, int* h_count, void* stream);
/* batched host form: frames contiguous with img_stride bytes between them; outputs nfeatures-capacity each */
int ef_detect_and_compute_host_batch(ef_handle* h, int nframes, const uint8_t* h_imgs, size_t img_stride, size_t pitch,
                                     int width, int height, float* h_kpts5, uint8_t* h_desc, int* h_counts, void* stream);

/* ---- single-process multi-GPU driver (new): frames are independent, a batch is cut into contiguous per-device
 * blocks and every device runs the single-GPU pipeline on its block from its own host thread and stream; no cross-GPU
 * exchange on the data path.  params->device is ignored; devices == NULL means ordinals 0 .. ndev-1; params->max_batch is the
 * sub-batch size per device.  Host buffers as in ef_detect_and_compute_host_batch (nframes may exceed max_batch). */
typedef struct ef_mg_handle ef_mg_handle;
int ef_mg_create(const ef_params* params, const int* devices, int ndev, ef_mg_handle** out);
void ef_mg_destroy(ef_mg_handle* m);
int ef_mg_device_count(const ef_mg_handle* m);
void ef_mg_shard_range(int nframes, int i, int ndev, int* begin, int* end);
int ef_mg_detect_and_compute_host_batch(ef_mg_handle* m, int nframes, const uint8_t* h_imgs, size_t img_stride, size_t pitch,
                                        int width, int height, float* h_kpts5, uint8_t* h_desc, int* h_counts);
const char* ef_mg_last_error_string(const ef_mg_handle* m);

/* ---- one oversized frame over several GPUs (new; SURVEY 8e "one oversized frame", north_star "image tiles with halo").
 * Every GPU holds the whole image (the caller broadcasts it: ncclBroadcast / torch.distributed.broadcast) and is band
 * `shard` of `nshards`: it builds the whole pyramid but runs FAST/Harris, the radius NMS and the compaction only on its
 * horizontal band of every level (score stage: + the NMS halo), keeps its local top-quota per level and packs it into
 * d_cand (ef_band_candidate_bytes() per frame).  The caller all-gathers the candidate buffers of all GPUs in shard order
 * ([nshards][nframes][bytes]; <= 8 B x nfeatures per GPU) and calls ef_band_finish_async, which re-selects the global
 * top-quota (bit-identical to the single-GPU result), writes the FULL keypoint matrix on every GPU and the descriptors of
 * this GPU's block of output rows [row0, row0 + nrows) (ef_band_desc_rows: equal blocks of ceil(nfeatures / nshards) rows; the other
 * rows are left undefined): an all-gather of the fixed-size row blocks -- in place over d_desc when its capacity is
 * nshards * nrows rows -- assembles the frame's descriptors on every GPU.  No reduction crosses the GPUs.  nshards == 1
 * degenerates to ef_detect_and_compute_batch_async.  Replaces the same reference entry (src/cuda_efficient_features.cpp:225-321). */
size_t ef_band_candidate_bytes(const ef_handle* h);
int ef_band_detect_async(ef_handle* h, int shard, int nshards, int nframes, const uint8_t* d_imgs, size_t img_stride, size_t pitch,
                         int width, int height, uint8_t* d_cand, void* stream);
int ef_band_finish_async(ef_handle* h, int shard, int nshards, int nframes, const uint8_t* d_all_cand,
                         float* d_kpts, size_t kpts_stride, size_t kpts_pitch, uint8_t* d_desc, size_t desc_stride, size_t desc_pitch,
                         int* d_counts, void* stream);
/* host arithmetic of the band partition: tile rows (32 pixel rows each) of a level with `tiles_y` tile rows owned by `shard`
 * and the rows its score stage covers with `halo_tiles` extra tile rows on either side */
void ef_band_tile_rows(int tiles_y, int shard, int nshards, int halo_tiles, int* own0, int* own_n, int* score0, int* score_n);
void ef_band_desc_rows(int nfeatures, int shard, int nshards, int* row0, int* nrows);

/* ---- callers either side of the path ------------------------------------------------------------
 * Brute-force Hamming matcher on the descriptors the path produced (32- or 64-byte rows), stateless.
 * ef_match_knn_async replaces cv::BFMatcher(NORM_HAMMING)::knnMatch(query, train, matches, k) for k = 1 or 2
 * (samples/sample_image_sequence.cpp:81,115-116): d_idx / d_dist are nq x k ints, row q = the k lexicographically smallest
 * (distance, trainIdx) pairs (OpenCV's order); missing entries (nt < k) are idx -1.
 * ef_match_cross_check_async replaces cv::BFMatcher::create(NORM_HAMMING, true)->match (samples/sample_feature_matching.cpp:99-101):
 * d_train_idx[q] = the train row matched to query q, or -1 (OpenCV omits those queries from the match list).
 * ef_match_ratio_cross_async is the filter loop of samples/sample_image_sequence.cpp:121-137 over two k = 2 results:
 * d_out_train[q] = matched train row or -1.
 * d_scratch: ef_match_scratch_bytes(nq, nt) bytes of device memory owned by the caller (no hidden allocation). */
size_t ef_match_scratch_bytes(int nq, int nt);
int ef_match_knn_async(const uint8_t* d_query, size_t qpitch, int nq, const uint8_t* d_train, size_t tpitch, int nt, int desc_bytes, int k,
                       int* d_idx, int* d_dist, void* d_scratch, void* stream);
int ef_match_cross_check_async(const uint8_t* d_query, size_t qpitch, int nq, const uint8_t* d_train, size_t tpitch, int nt, int desc_bytes,
                               int* d_train_idx, int* d_dist, void* d_scratch, void* stream);
int ef_match_ratio_cross_async(const int* d_idx12, const int* d_dist12, int nq, const int* d_idx21, const int* d_dist21, int nt,
                               double uniqueness, int* d_out_train, void* stream);
const char* ef_match_last_error_string(void);
/* replaces convertToGray (samples/sample_common.cpp:35-45: cv::cvtColor COLOR_BGR2GRAY / COLOR_BGRA2GRAY): interleaved 8-bit
 * BGR (channels 3) or BGRA (4) device image -> CV_8UC1, OpenCV's 15-bit fixed-point weights */
int ef_bgr_to_gray_async(const uint8_t* d_src, size_t src_pitch, int width, int height, int channels,
                         uint8_t* d_gray, size_t gray_pitch, void* stream);

/* ---- introspection for stage-by-stage parity tests (not part of the reference API) ---------- */
typedef struct ef_level_view {
    int width, height;
    float scale;                 /* scales_[s] */
    int quota;                   /* nfeaturesPerLevel_[s] */
    const uint8_t* d_image;      /* imagePyr_[s] (level 0 aliases the caller's image) */
    size_t image_pitch;
    const uint8_t* d_blurred;    /* blurPyr_[s] (valid after a call that computed descriptors) */
    size_t blurred_pitch;
    const float* d_response;     /* dense Harris map, -inf where not a FAST corner */
    size_t response_pitch;       /* in floats */
} ef_level_view;
int ef_debug_level_view(const ef_handle* h, int frame, int level, ef_level_view* out);
/* per level {corners, survivors, selected}; synchronises the stream passed */
int ef_debug_level_counts(ef_handle* h, int frame, int* h_counts3, void* stream);
/* HashSIFT stage outputs of the last compute call: n x 128 uint8 SIFT vector, n x nbits fp32 projection
 * (projection only kept when ef_debug_keep_projection(h,1) was set before the call) */
int ef_debug_keep_projection(ef_handle* h, int keep);
int ef_debug_hashsift_views(const ef_handle* h, const uint8_t** d_sift128, const float** d_projection);
/* the projection stage alone on caller-provided SIFT vectors (n x 128 uint8, 16-byte aligned): path 0 = default, 1 = tcgen05,
 * 2 = mma.sync, 3 = fp64 CUDA cores (1 and 2: the exact six-digit integer GEMM).  Tests only, not thread safe. */
int ef_debug_project_async(ef_handle* h, const uint8_t* d_sift128, int n, int path, uint8_t* d_desc, size_t desc_pitch, void* stream);
/* blocking device->host 2-D copy of one of the views above (tests only) */
int ef_debug_copy_to_host(ef_handle* h, void* dst, size_t dst_pitch, const void* d_src, size_t src_pitch,
                          size_t width_bytes, size_t rows);

/* ---- measurement support (bench.py) ---------------------------------------------------------- */
/* per-stage device timing with CUDA events recorded on the caller's stream between the stages of
 * ef_detect_and_compute*_async.  ef_stage_times() synchronises, returns the SUM of milliseconds per stage
 * over all calls since the last enable/reset and the number of calls, and resets the accumulators. */
enum { EF_STAGE_PYRAMID = 0, EF_STAGE_SCORE = 1, EF_STAGE_NMS = 2, EF_STAGE_COMPACT = 3, EF_STAGE_SELECT = 4,
       EF_STAGE_ANGLE_PACK = 5, EF_STAGE_BLUR = 6, EF_STAGE_DESCRIBE = 7, EF_STAGE_PROJECT = 8, EF_NUM_STAGES = 9 };
int ef_stage_timing_enable(ef_handle* h, int enable);
int ef_stage_times(ef_handle* h, float* ms_sum /* [EF_NUM_STAGES] */, int* ncalls);
/* number of kernels this library has launched in the process so far */
unsigned long long ef_kernel_launch_count(void);
/* synthetic uniform-noise frames of the benchmark: pix(f, y, x) = lowbias32(seed ^ ((f * H + y) * W + x)) >> 24 with
 * f = first_frame + frame index -- the generator the CPU arm uses (oracle efo_synth_frame), on the device, so that both arms of
 * bench.py see the same pixels.  Frame i is written at d_frames + i * frame_stride, rows `pitch` bytes apart. */
int ef_synth_frames_async(uint8_t* d_frames, size_t pitch, size_t frame_stride, int width, int height, int nframes,
                          unsigned seed, unsigned first_frame, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* EF_B200_H */
