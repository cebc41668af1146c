// ef_common.cuh -- shared device/host definitions of the B200 detectAndCompute pipeline.
//
// Everything in csrc/ is compiled with -fmad=false: no implicit contraction anywhere.  Where the
// canonical arithmetic has a fused multiply-add (the pattern nvcc emits for the reference's CUDA
// detector) it is written as an explicit fmaf(); where the canonical
// arithmetic is the reference's CPU build (generic x86-64, no FMA: BAD and HashSIFT) plain
// operators are used and stay unfused.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>
#include <math.h>

#include "../../include/ef_b200.h"
#include "ef_tma.cuh"

#define EF_HALF_PATCH 15          // HALF_PATCH_SIZE, cuda_efficient_features.cpp:34 (border mask, IC radius)
#define EF_PATCH_SIZE 31.f        // PATCH_SIZE, cuda_efficient_features.cu:36
#define EF_TILE 32                // square pixel tile of the score / NMS kernels
#define EF_NEG_INF (__int_as_float(0xff800000))

struct __align__(8) EfSurvivor { short x, y; float resp; };
// NMS block-maximum map entry: largest response of a b x b pixel block (-inf: no corner) and where it is:
// pos = tie << 31 | y << 16 | x (level coordinates; tie: the maximum is attained by more than one pixel)
struct __align__(8) EfBlockMax { float val; unsigned pos; };                 // compacted NMS survivor
struct __align__(16) EfSelected { short x, y; float resp; float angle; int pad; }; // per-level selected keypoint (level coords)

// per-frame, per-level counters (zeroed at the start of every call)
struct EfLevelCounters { int corners; int survivors; int selected; int overflow; };

struct EfLevel {
    int w, h;
    int img_pitch;      // bytes, internal pyramid level (level 0: caller's pitch)
    int blur_pitch;     // bytes
    int resp_pitch;     // floats
    int tiles_x, tiles_y;
    unsigned tiles_x_inv, strips_x_inv, blur_tiles_x_inv; // floor((2^32 - 1) / d) of tiles_x, strips_x, blur_tiles_x: ef_div_fast()
    // tile rows this call works on (whole level unless the frame is cut into bands over several GPUs, ef_band_*):
    // the score stage covers rows [score_ty0, score_ty0 + score_rows) = the owned rows plus the NMS halo,
    // NMS / compaction the owned rows [own_ty0, own_ty0 + own_rows)
    int score_ty0, score_rows, own_ty0, own_rows;
    int blk_w, blk_h;   // dimensions of the NMS block-maximum map (ceil(w / nms_block), ceil(h / nms_block))
    int tile_start;     // first tile index of this level in the all-level tile table
    int strips_x, strip_start; // NMS strips of 4 tiles per tile row; first strip index of this level
    int blur_tile_start, blur_tiles_x; // 64x64 blur tiles
    int blur_ty0, blur_rows;           // blur tile rows of this call (whole level; band-sharded frames: the rows the owned keypoints' windows touch)
    int band_start;     // first 32-row band index of this level
    int quota;          // nfeaturesPerLevel_[s]
    int surv_cap;       // capacity of the survivor list
    int kpt_block_start;// first descriptor-CTA index of this level (quota / 8 keypoints per CTA, rounded up)
    int sift_block_start;// same for the HashSIFT feature kernel (4 keypoints per CTA)
    float scale;        // scales_[s]
    float rx, ry;       // resize ratios src/dst for producing THIS level from the previous one
    // byte offsets inside one frame slot of the workspace
    unsigned long long img_off, blur_off, resp_off, blk_off, mask_off, rowcnt_off, surv_off, sel_off;
};

struct EfPipe {
    int nlevels, first_level, nframes;
    int fast_threshold;
    int nms_r2, nms_R;          // ceil(r^2); largest |d| with d^2 < nms_r2
    int nms_block, nms_K;       // block edge of the block-maximum map (0: r2 <= 1, nothing is suppressed); block reach of the disc
    int total_tiles, total_blur_tiles, total_bands, total_kpt_blocks, total_sift_blocks, total_strips;
    int nfeatures;              // output capacity (columns)
    int desc_type, desc_bytes;
    int shard_i, shard_n;       // descriptor CTAs dealt round-robin to shard_n GPUs (unused by ef_band_*: kept for callers that shard by CTA); 0, 1 otherwise
    // ef_band_finish_async: this GPU describes the keypoints of output rows [desc_row0, desc_row0 + desc_rows) only (whole matrix otherwise);
    // blur_by_slice: the blur skips the tiles none of their windows can touch (slice_y[frame][level] = {first row - 24, last row + 24})
    int desc_row0, desc_rows, blur_by_slice;
    const int* slice_y;
    int select_from_counters;   // select stage: candidate count comes from counters[].overflow (merged band candidates), not from rowcnt
    // caller buffers (frame f at base + f*stride)
    const uint8_t* img0; unsigned long long img0_stride; int img0_pitch;
    float* kpts; unsigned long long kpts_stride; int kpts_pitch;      // bytes
    uint8_t* desc; unsigned long long desc_stride; int desc_pitch;    // bytes
    int* counts;
    // workspace
    uint8_t* ws; unsigned long long ws_stride;           // per-frame slot
    EfLevelCounters* counters; /* [frame][EF_MAX_LEVELS] */
    EfLevel lv[EF_MAX_LEVELS];
};

__host__ __device__ inline int ef_div_up(int a, int b) { return (a + b - 1) / b; }
__host__ __device__ inline unsigned long long ef_align_up(unsigned long long v, unsigned long long a) { return (v + a - 1) / a * a; }

#ifdef __CUDACC__
__device__ __forceinline__ const uint8_t* ef_level_image(const EfPipe& p, int frame, int level, int& pitch)
{
    if (level == 0) { pitch = p.img0_pitch; return p.img0 + (unsigned long long)frame * p.img0_stride; }
    pitch = p.lv[level].img_pitch;
    return p.ws + (unsigned long long)frame * p.ws_stride + p.lv[level].img_off;
}
__device__ __forceinline__ uint8_t* ef_ws(const EfPipe& p, int frame, unsigned long long off)
{
    return p.ws + (unsigned long long)frame * p.ws_stride + off;
}
// saturate_cast<uchar>(float) of OpenCV CUDA: cvt.rni.sat.u8.f32
__device__ __forceinline__ unsigned ef_sat_u8_rne(float v)
{
    unsigned r;
    asm("cvt.rni.sat.u8.f32 %0, %1;" : "=r"(r) : "f"(v));
    return r;
}
__device__ __forceinline__ int ef_reflect101(int i, int n)
{
    if (n == 1) return 0;
    while (i < 0 || i >= n) i = i < 0 ? -i : 2 * (n - 1) - i;
    return i;
}
// ---- packed fp32 (Blackwell add/sub/mul/fma.rn.f32x2: two independent IEEE operations per instruction, bit-identical to the scalar forms)
__device__ __forceinline__ unsigned long long ef_pack2(float lo, float hi)
{
    unsigned long long r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
    return r;
}
__device__ __forceinline__ void ef_unpack2(unsigned long long v, float& lo, float& hi)
{
    asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v));
}
__device__ __forceinline__ unsigned long long ef_mul2(unsigned long long a, unsigned long long b)
{
    unsigned long long r;
    asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
    return r;
}
__device__ __forceinline__ unsigned long long ef_fma2(unsigned long long a, unsigned long long b, unsigned long long c)
{
    unsigned long long r;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c));
    return r;
}
__device__ __forceinline__ unsigned long long ef_add2(unsigned long long a, unsigned long long b)
{
    unsigned long long r;
    asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
    return r;
}
// round toward -infinity: x + 12582912.f (1.5 * 2^23, ulp 1) = 12582912 + floor(x) for |x| < 2^22 -- floor() on the FMA pipe
__device__ __forceinline__ unsigned long long ef_add2_rm(unsigned long long a, unsigned long long b)
{
    unsigned long long r;
    asm("add.rm.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
    return r;
}
__device__ __forceinline__ unsigned long long ef_sub2(unsigned long long a, unsigned long long b)
{
    unsigned long long r;
    asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
    return r;
}
// t / d for 0 <= t < 2^31, d >= 1, inv = floor((2^32 - 1) / d): the estimate umulhi(t, inv) is q or q - 1
__device__ __forceinline__ int ef_div_fast(int t, int d, unsigned inv)
{
    unsigned q = __umulhi((unsigned)t, inv);
    if ((unsigned)t - q * (unsigned)d >= (unsigned)d) q++;
    return (int)q;
}
// order-preserving map float -> uint32 (larger float <=> larger key)
__device__ __forceinline__ unsigned ef_float_key(float f)
{
    const unsigned b = __float_as_uint(f);
    return (b & 0x80000000u) ? ~b : (b | 0x80000000u);
}
#endif

// every kernel launch of the library is counted (bench.py reports it as gpu_launches)
extern unsigned long long g_ef_launches;
#define EF_COUNT_LAUNCH(n) ((void)__atomic_fetch_add(&g_ef_launches, (unsigned long long)(n), __ATOMIC_RELAXED)) // host threads of ef_mg_*

// ---- launchers (host) ----------------------------------------------------------------------------
void ef_launch_pyramid(const EfPipe& p, const EfTmaMaps* maps /* nullptr: no TMA */, cudaStream_t s);
void ef_launch_score(const EfPipe& p, cudaStream_t s);
void ef_launch_nms(const EfPipe& p, cudaStream_t s);
void ef_launch_compact(const EfPipe& p, cudaStream_t s);
void ef_launch_select(const EfPipe& p, cudaStream_t s);
void ef_launch_angle_pack(const EfPipe& p, cudaStream_t s);
void ef_launch_blur(const EfPipe& p, const EfTmaMaps* maps /* nullptr: no TMA */, cudaStream_t s);
// band-sharded single frame (ef_band_*): packed per-band candidates, merge of all bands, ownership mask of descriptor rows
#define EF_BAND_HDR 256
void ef_launch_band_pack(const EfPipe& p, uint8_t* cand, unsigned long long cand_stride, cudaStream_t s);
void ef_launch_band_merge(const EfPipe& p, const uint8_t* all, unsigned long long cand_stride, int nshards, cudaStream_t s);
void ef_launch_band_slice_ranges(const EfPipe& p, int* slice_y, cudaStream_t s);

// descriptor stage: keypoints either from the per-level selected lists (detectAndCompute) or from a
// flat caller array (compute-only API)
struct EfDescJob {
    // image
    const uint8_t* img; int w, h, pitch;
    // keypoints: n x float4 (x, y, size, angle)
    const float4* kpts; int n;
    float scale;             // BAD scaleFactor / HashSIFT croppingScale
    uint8_t* desc; int desc_pitch;
    int nbits;
    int staged31;            // every keypoint has integer coordinates and size 31, scale == 1, image base and pitch 16-byte aligned:
                             // the window-staging kernels of the detectAndCompute path apply (5 x N GpuMat compute path)
};
struct EfBadTables { const uchar4* boxes_xyxy; const unsigned char* radius; const float* thresholds; };

void ef_launch_integral(const uint8_t* img, int w, int h, int pitch, unsigned* integral, unsigned* segsum, cudaStream_t s);
void ef_launch_bad_flat(const EfDescJob& job, const unsigned* integral, const EfBadTables& t, cudaStream_t s);
void ef_launch_bad_pipe(const EfPipe& p, const EfBadTables& t, cudaStream_t s);

// exp_table: expf weight of the 30x30 gradient positions; grad_table[(dy+255)*511 + dx+255] = { sqrtf(dx^2+dy^2) with the sign bit = bit 2 of
// the orientation bin, orientation-bin fraction with bits 31:30 = bits 1:0 of the bin } (see ef_hashsift.cu, ef_api.cu)
struct EfHashSiftTables { const float* exp_table; const float2* grad_table; };
void ef_launch_hashsift_features_flat(const EfDescJob& job, const EfHashSiftTables& t, uint8_t* sift128, cudaStream_t s);
void ef_launch_hashsift_features_pipe(const EfPipe& p, const EfHashSiftTables& t, uint8_t* sift128, cudaStream_t s);
// projection + sign + pack.  rows = keypoint rows in sift128 (n x 128 u8).  bfrag != nullptr: exact integer-tensor-core
// path (fixed-point digits of the weights in mma fragment order, int64 bias, scale 2^-shift); else fp64 fallback on weights_t.
struct EfProjTables { const uint4* bfrag; const long long* bias; int shift; const float* weights_t; /*129 x nbits*/
                      const uint8_t* btc; /* digits in UMMA core-matrix order for the tcgen05 path (ef_project_tc.cu), or nullptr */
                      int ndigits;        /* balanced base-256 digits per weight in btc: 6 (512-bit table) or 7 (256-bit table) */ };
bool ef_launch_hashsift_project_tc(const uint8_t* sift128, int n_cap, const int* d_counts, int nframes, const EfProjTables& t, int nbits,
                                   uint8_t* desc, size_t desc_stride, int desc_pitch, float* proj_out, cudaStream_t s);
extern int g_ef_project_path;
void ef_launch_hashsift_project(const uint8_t* sift128, int n_cap, const int* d_n, const EfProjTables& t, int nbits,
                                uint8_t* desc, int desc_pitch, float* proj_out, cudaStream_t s);
void ef_launch_hashsift_project_batch(const uint8_t* sift128, int n_cap, const int* d_counts, int nframes, const EfProjTables& t, int nbits,
                                      uint8_t* desc, size_t desc_stride, int desc_pitch, float* proj_out, cudaStream_t s);
// 5xN keypoint rows -> n x float4 (x, y, 31, angle): convertKeypointsKernel, cuda_efficient_features.cu:250-263
void ef_launch_convert_rows(const float* kpts5, size_t kpts_pitch, int n, float4* out, cudaStream_t s);
