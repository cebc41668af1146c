// ef_desc.cu -- BAD and HashSIFT descriptor kernels (sm_100a), bit-exact restatement targets:
//   BAD      modules/efficient_features/src/bad.cpp:86-157,166-251,320-405   (CPU ground truth)
//   HashSIFT modules/efficient_features/src/hash_sift.cpp:68-138,150-198,200-331,353-378
// They replace the reference's GPU kernels src/cuda_bad.cu:246-316 (+ cudev integral :350-363) and
// src/cuda_hash_sift.cu:380-435 (+ cublasSgemm, cuda_hash_sift.cpp:44-60), which are NOT bit-exact
// against the CPU code (float trig, smem float atomics, fp32 SGEMM).
//
// Compiled with -fmad=false: the CPU reference is a generic x86-64 build without FMA, so every
// a*b+c below must stay two roundings.
#include "ef_common.cuh"

#include <cfloat>

#define EF_DESC_WARPS 8

// =================================================================================================
// integral image (generic compute-only path; cv::integral / cudev integral, wrapping uint32)
//   I is (h+1) x (w+1), dense pitch iw = w+1.
// =================================================================================================
#define EF_INT_SEG 64

__global__ void __launch_bounds__(256) ef_integral_rows_kernel(const uint8_t* __restrict__ img, int w, int h, int pitch,
                                                               unsigned* __restrict__ I)
{
    __shared__ unsigned s_warp[8];
    __shared__ unsigned s_carry;
    const int y = blockIdx.x; // 0..h  (row h of the grid zeroes I row 0)
    const int iw = w + 1;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (y == h) {
        for (int x = tid; x < iw; x += 256) I[x] = 0;
        return;
    }
    unsigned* out = I + (size_t)(y + 1) * iw;
    const uint8_t* row = img + (size_t)y * pitch;
    if (tid == 0) { s_carry = 0; out[0] = 0; }
    __syncthreads();
    for (int x0 = 0; x0 < w; x0 += 1024) {
        const int x = x0 + tid * 4;
        unsigned v[4];
#pragma unroll
        for (int j = 0; j < 4; j++) v[j] = (x + j < w) ? row[x + j] : 0u;
        v[1] += v[0]; v[2] += v[1]; v[3] += v[2];
        unsigned inc = v[3];
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const unsigned u = __shfl_up_sync(0xffffffffu, inc, o);
            if (lane >= o) inc += u;
        }
        if (lane == 31) s_warp[warp] = inc;
        __syncthreads();
        unsigned woff = 0;
        for (int k = 0; k < warp; k++) woff += s_warp[k];
        const unsigned base = s_carry + woff + inc - v[3];
#pragma unroll
        for (int j = 0; j < 4; j++)
            if (x + j < w) out[x + j + 1] = base + v[j];
        __syncthreads();
        if (tid == 255) s_carry = base + v[3];
        __syncthreads();
    }
}

__global__ void __launch_bounds__(128) ef_integral_segsum_kernel(const unsigned* __restrict__ I, int w, int h, unsigned* __restrict__ segsum)
{
    const int iw = w + 1;
    const int x = blockIdx.x * 128 + threadIdx.x;
    const int seg = blockIdx.y;
    if (x >= iw) return;
    const int ya = 1 + seg * EF_INT_SEG, yb = min(ya + EF_INT_SEG, h + 1);
    unsigned s = 0;
    for (int y = ya; y < yb; y++) s += I[(size_t)y * iw + x];
    segsum[(size_t)seg * iw + x] = s;
}

__global__ void __launch_bounds__(128) ef_integral_segscan_kernel(int w, int nseg, unsigned* __restrict__ segsum)
{
    const int iw = w + 1;
    const int x = blockIdx.x * 128 + threadIdx.x;
    if (x >= iw) return;
    unsigned run = 0;
    for (int s = 0; s < nseg; s++) {
        const unsigned v = segsum[(size_t)s * iw + x];
        segsum[(size_t)s * iw + x] = run;
        run += v;
    }
}

__global__ void __launch_bounds__(128) ef_integral_cols_kernel(unsigned* __restrict__ I, int w, int h, const unsigned* __restrict__ segoff)
{
    const int iw = w + 1;
    const int x = blockIdx.x * 128 + threadIdx.x;
    const int seg = blockIdx.y;
    if (x >= iw) return;
    const int ya = 1 + seg * EF_INT_SEG, yb = min(ya + EF_INT_SEG, h + 1);
    unsigned run = segoff[(size_t)seg * iw + x];
    for (int y = ya; y < yb; y++) {
        run += I[(size_t)y * iw + x];
        I[(size_t)y * iw + x] = run;
    }
}

void ef_launch_integral(const uint8_t* img, int w, int h, int pitch, unsigned* integral, unsigned* segsum, cudaStream_t s)
{
    const int iw = w + 1;
    const int nseg = ef_div_up(h, EF_INT_SEG);
    ef_integral_rows_kernel<<<h + 1, 256, 0, s>>>(img, w, h, pitch, integral);
    EF_COUNT_LAUNCH(1);
    ef_integral_segsum_kernel<<<dim3(ef_div_up(iw, 128), nseg), 128, 0, s>>>(integral, w, h, segsum);
    EF_COUNT_LAUNCH(1);
    ef_integral_segscan_kernel<<<ef_div_up(iw, 128), 128, 0, s>>>(w, nseg, segsum);
    EF_COUNT_LAUNCH(1);
    ef_integral_cols_kernel<<<dim3(ef_div_up(iw, 128), nseg), 128, 0, s>>>(integral, w, h, segsum);
    EF_COUNT_LAUNCH(1);
}

// =================================================================================================
// BAD
// =================================================================================================
struct EfBadAffine { float m00, m01, m02, m10, m11, m12, s; bool border; };

// rectifyBoxes (bad.cpp:121-147) + isKeypointInTheBorder (bad.cpp:92-102)
__device__ __forceinline__ EfBadAffine ef_bad_affine(float kx, float ky, float size, float angle, float scaleFactor, int w, int h)
{
    EfBadAffine a;
    const float s = scaleFactor * size / (0.5f * (float)(32 + 32));
    if (angle == -1) {
        a.m00 = s; a.m01 = 0.0f; a.m02 = -0.5f * s * 32.f + kx;
        a.m10 = 0.0f; a.m11 = s; a.m12 = -s * 0.5f * 32.f + ky;
    } else {
        // double cos/sin exactly like the CPU code (bad.cpp:29,138-139)
        const float cosine = (angle >= 0) ? (float)cos((double)angle * 0.017453292519943295) : 1.f;
        const float sine = (angle >= 0) ? (float)sin((double)angle * 0.017453292519943295) : 0.f;
        a.m00 = s * cosine; a.m01 = -s * sine;
        a.m02 = (-s * cosine + s * sine) * 32.f * 0.5f + kx;
        a.m10 = s * sine; a.m11 = s * cosine;
        a.m12 = (-s * sine - s * cosine) * 32.f * 0.5f + ky;
    }
    a.s = s;
    const float sb = scaleFactor * size / (float)(32 + 32);
    const float bw = 32.f * sb * 1.75f, bh = 32.f * sb * 1.75f;
    bool border = false;
    if (kx < bw || kx + bw >= (float)w) border = true;
    if (ky < bh || ky + bh >= (float)h) border = true;
    a.border = border;
    return a;
}

// integral accessors.  box(y0, x0, y1, x1) = sum of the pixels in rows [y0, y1) x columns [x0, x1)
struct EfGlobalIntegral { // (h+1) x (w+1) uint32 integral image of the whole frame (wrapping arithmetic, exact differences)
    const unsigned* I; int iw, ih;
    // A keypoint closer to the image edge than its largest box, yet not "in the border" by bad.cpp:92-102 (the test scales with the
    // keypoint size: size 2 on the last row passes it), makes the reference read past its integral image -- undefined there.  Here,
    // and in the oracle, such reads are clamped to the last row / column of the integral image: defined, and never out of the buffer.
    __device__ __forceinline__ unsigned at(int y, int x) const { return __ldg(I + (size_t)min(max(y, 0), ih - 1) * iw + min(max(x, 0), iw - 1)); }
    __device__ __forceinline__ unsigned box(int y0, int x0, int y1, int x1) const { return at(y0, x0) + at(y1, x1) - at(y0, x1) - at(y1, x0); }
};
// 16-bit modular integral of the 48-row window around a keypoint (ef_bad_pipe_kernel): P[r][a] (halfword r*PITCH + a) =
// sum over window rows < r of the bytes [0, a) of the 64-byte aligned row segment, a in [0, 63].
// Exact for boxes whose true sum is < 2^16, i.e. radius <= 7 (15 x 15 x 255 = 57375): the detectAndCompute path (size 31, scale 1).
#define EF_BW_PITCH 66 // halfwords per row: 64 entries (32 words) + one pad word: 33 words (odd) spreads the banks
struct EfWindowIntegral16 {
    const unsigned short* P; int gx0, wy0;
    __device__ __forceinline__ unsigned at(int y, int x) const { return P[(y - wy0) * EF_BW_PITCH + (x - gx0)]; }
    __device__ __forceinline__ unsigned box(int y0, int x0, int y1, int x1) const { return (at(y0, x0) + at(y1, x1) - at(y0, x1) - at(y1, x0)) & 0xffffu; }
};

template <class Integral>
__device__ __forceinline__ bool ef_bad_bit(const Integral& I, const EfBadAffine& a, uchar4 b, int r0, float thr, int w, int h)
{
    // bad.cpp:151-155 (CV_ROUNDNUM truncates)
    const float bx1 = (float)b.x, by1 = (float)b.y, bx2 = (float)b.z, by2 = (float)b.w;
    const int x1 = __float2int_rz(a.m00 * bx1 + a.m01 * by1 + a.m02 + 0.5f);
    const int y1 = __float2int_rz(a.m10 * bx1 + a.m11 * by1 + a.m12 + 0.5f);
    const int x2 = __float2int_rz(a.m00 * bx2 + a.m01 * by2 + a.m02 + 0.5f);
    const int y2 = __float2int_rz(a.m10 * bx2 + a.m11 * by2 + a.m12 + 0.5f);
    const int r = __float2int_rz(a.s * (float)r0 + 0.5f);
    if (!a.border) {
        // bad.cpp:371-393: integer box sums, threshold scaled by the box area
        const int side = 1 + (r << 1);
        const unsigned acc = I.box(y1 - r, x1 - r, y1 + r + 1, x1 + r + 1) - I.box(y2 - r, x2 - r, y2 + r + 1, x2 + r + 1);
        return (float)(int)acc <= (thr * (float)(side * side));
    }
    // bad.cpp:166-251: clamped boxes, float means.  frameWidth/Height = integral dims (w+1, h+1)
    const int fw = w + 1, fh = h + 1;
    int ax1 = x1 - r; if (ax1 < 0) ax1 = 0; else if (ax1 >= fw - 1) ax1 = fw - 2;
    int ay1 = y1 - r; if (ay1 < 0) ay1 = 0; else if (ay1 >= fh - 1) ay1 = fh - 2;
    int ax2 = x1 + r + 1; if (ax2 <= 0) ax2 = 1; else if (ax2 >= fw) ax2 = fw - 1;
    int ay2 = y1 + r + 1; if (ay2 <= 0) ay2 = 1; else if (ay2 >= fh) ay2 = fh - 1;
    int bx1i = x2 - r; if (bx1i < 0) bx1i = 0; else if (bx1i >= fw - 1) bx1i = fw - 2;
    int by1i = y2 - r; if (by1i < 0) by1i = 0; else if (by1i >= fh - 1) by1i = fh - 2;
    int bx2i = x2 + r + 1; if (bx2i <= 0) bx2i = 1; else if (bx2i >= fw) bx2i = fw - 1;
    int by2i = y2 + r + 1; if (by2i <= 0) by2i = 1; else if (by2i >= fh) by2i = fh - 1;
    const float sum1 = (float)(int)I.box(ay1, ax1, ay2, ax2);
    const float avg1 = sum1 / (float)((ay2 - ay1) * (ax2 - ax1));
    const float sum2 = (float)(int)I.box(by1i, bx1i, by2i, bx2i);
    const float avg2 = sum2 / (float)((by2i - by1i) * (bx2i - bx1i));
    return (avg1 - avg2) <= thr;
}

// one warp = one keypoint; lane l evaluates pairs l, l+32, ...; 32 bits -> 4 bytes, MSB first (bad.cpp:349,368)
template <class Integral>
__device__ __forceinline__ void ef_bad_describe(const Integral& I, const EfBadAffine& a, const EfBadTables& t, int nbits,
                                                int w, int h, uint8_t* out, int lane)
{
    for (int g = 0; g < nbits; g += 32) {
        const int i = g + lane;
        const bool bit = ef_bad_bit(I, a, t.boxes_xyxy[i], (int)t.radius[i], t.thresholds[i], w, h);
        const unsigned bal = __brev(__ballot_sync(0xffffffffu, bit)); // pair g+0 -> bit 31
        if (lane < 4) out[(g >> 3) + lane] = (uint8_t)(bal >> (24 - 8 * lane));
    }
}

// ---- generic path: flat keypoint array, global integral image ----------------------------------
__global__ void __launch_bounds__(EF_DESC_WARPS * 32) ef_bad_flat_kernel(const EfDescJob job, const unsigned* __restrict__ integral, const EfBadTables t)
{
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int i = blockIdx.x * EF_DESC_WARPS + warp;
    if (i >= job.n) return;
    const float4 k = job.kpts[i];
    const EfBadAffine a = ef_bad_affine(k.x, k.y, k.z, k.w, job.scale, job.w, job.h);
    EfGlobalIntegral I; I.I = integral; I.iw = job.w + 1; I.ih = job.h + 1;
    ef_bad_describe(I, a, t, job.nbits, job.w, job.h, job.desc + (size_t)i * job.desc_pitch, lane);
}

void ef_launch_bad_flat_window(const EfDescJob& job, const EfBadTables& t, cudaStream_t s);

void ef_launch_bad_flat(const EfDescJob& job, const unsigned* integral, const EfBadTables& t, cudaStream_t s)
{
    if (job.n <= 0) return;
    if (job.staged31) { ef_launch_bad_flat_window(job, t, s); return; }
    ef_bad_flat_kernel<<<ef_div_up(job.n, EF_DESC_WARPS), EF_DESC_WARPS * 32, 0, s>>>(job, integral, t);
    EF_COUNT_LAUNCH(1);
}

// ---- detectAndCompute path: keypoints from the per-level selected lists on the blurred level,
//      size 31, scale 1 (cuda_efficient_features.cpp:48-69,306).  Every box corner then lies in
//      [k-22, k+23] (exhaustive over both tables and all angles), so the 48 window rows k-24 .. k+23, read as 64-byte
//      aligned row segments (four 16-byte loads per row), and a 16-bit modular integral of them in shared memory
//      (6.5 KB per keypoint) replace the reference's full-frame integral image (-4P writes, -4P reads per level).
//        1. load: lane = (row mod 8, 16-byte chunk); row prefix in registers: 16 byte adds + a 4-lane shuffle scan
//        2. column prefix in shared memory, two columns (one 32-bit word) per lane, halves added independently
//        3. lane = box pair: 8 halfword lookups per pair, ballot + brev packs 32 descriptor bits
#define EF_BW_HALF 24
#define EF_BW_ROWS (2 * EF_BW_HALF)     // 48 window rows
#define EF_BW_WORDS (EF_BW_PITCH / 2)   // 33

// window integral of one keypoint (kx, ky integers) into W ((EF_BW_ROWS + 1) x EF_BW_WORDS words of shared memory), one warp
__device__ __forceinline__ void ef_bad_window_integral(unsigned* __restrict__ W, const uint8_t* __restrict__ img, int pitch, int h, int gx0, int wy0, int lane)
{
    // ---- 1. rows -> exclusive row prefixes: P[r+1][a] = sum of the bytes [0, a) of window row r (before the column pass)
    W[lane] = 0;                                                           // P[0][*] = 0
    {
        const int c = lane & 3, g = lane >> 2;
        const int gxc = gx0 + 16 * c;
        const bool colok = gxc >= 0 && gxc + 15 < pitch;
#pragma unroll
        for (int it = 0; it < EF_BW_ROWS / 8; it++) {
            const int r = 8 * it + g, gy = wy0 + r;
            uint4 v = make_uint4(0, 0, 0, 0);
            if (colok && gy >= 0 && gy < h) v = __ldg(reinterpret_cast<const uint4*>(img + (size_t)gy * pitch + gxc));
            // exclusive prefix of the 16 bytes, two 16-bit lanes per register: q[m] = (prefix[2m], prefix[2m+1])
            unsigned q[8];
            unsigned run = 0;
            const unsigned wv[4] = { v.x, v.y, v.z, v.w };
#pragma unroll
            for (int m = 0; m < 8; m++) {
                const unsigned wd = wv[m >> 1];
                const unsigned b0 = (wd >> (16 * (m & 1))) & 0xffu, b1 = (wd >> (16 * (m & 1) + 8)) & 0xffu;
                const unsigned p1 = run + b0;
                q[m] = run | (p1 << 16);
                run = p1 + b1;
            }
            // exclusive scan of the chunk totals over the 4 lanes of the row
            unsigned tot = run;
            unsigned u = __shfl_up_sync(0xffffffffu, tot, 1); if (c >= 1) tot += u;
            u = __shfl_up_sync(0xffffffffu, tot, 2); if (c >= 2) tot += u;
            const unsigned base = (tot - run) * 0x10001u;
            uint4* dst = reinterpret_cast<uint4*>(W + (r + 1) * EF_BW_WORDS + 8 * c);
            // no carry between the halves: every prefix is < 64 * 255.  (rows are 132 bytes apart: 4-byte aligned stores only)
            unsigned* d32 = reinterpret_cast<unsigned*>(dst);
#pragma unroll
            for (int m = 0; m < 8; m++) d32[m] = q[m] + base;
        }
    }
    __syncwarp();
    // ---- 2. column prefix, modulo 2^16 per halfword: one 32-bit word (two columns) per lane
    {
        unsigned lo = 0, hi = 0;
#pragma unroll 8
        for (int r = 1; r <= EF_BW_ROWS; r++) {
            const unsigned wd = W[r * EF_BW_WORDS + lane];
            lo += wd & 0xffffu; hi += wd >> 16;
            W[r * EF_BW_WORDS + lane] = (lo & 0xffffu) | (hi << 16);
        }
    }
    __syncwarp();
}

__global__ void __launch_bounds__(EF_DESC_WARPS * 32) ef_bad_pipe_kernel(const __grid_constant__ EfPipe p, const EfBadTables t)
{
    extern __shared__ __align__(16) unsigned s_win_all[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int frame = blockIdx.y;
    const EfLevelCounters* ctr = &p.counters[frame * EF_MAX_LEVELS];
    const int bx = (int)blockIdx.x * p.shard_n + p.shard_i;   // descriptor CTAs dealt round-robin over the GPUs of a band-sharded frame
    if (bx >= p.total_kpt_blocks) return;
    int level = p.first_level;
    while (level + 1 < p.nlevels && bx >= p.lv[level + 1].kpt_block_start) level++;
    const EfLevel& L = p.lv[level];
    const int i = (bx - L.kpt_block_start) * EF_DESC_WARPS + warp;
    if (i >= ctr[level].selected) return;
    int offset = 0;
    for (int l = p.first_level; l < level; l++) offset += ctr[l].selected;
    const int row = offset + i;
    if (row >= p.nfeatures) return;

    if ((unsigned)(row - p.desc_row0) >= (unsigned)p.desc_rows) return;   // band-sharded frame: another GPU's output row (warp-uniform)
    const EfSelected k = reinterpret_cast<const EfSelected*>(ef_ws(p, frame, L.sel_off))[i];
    const uint8_t* __restrict__ img = ef_ws(p, frame, L.blur_off);
    const int pitch = L.blur_pitch;
    unsigned* __restrict__ W = s_win_all + warp * ((EF_BW_ROWS + 1) * EF_BW_WORDS);
    const int gx0 = (k.x - EF_BW_HALF) & ~15, wy0 = k.y - EF_BW_HALF;
    ef_bad_window_integral(W, img, pitch, L.h, gx0, wy0, lane);

    const EfBadAffine a = ef_bad_affine((float)k.x, (float)k.y, EF_PATCH_SIZE, k.angle, 1.f, L.w, L.h);
    EfWindowIntegral16 I; I.P = reinterpret_cast<const unsigned short*>(W); I.gx0 = gx0; I.wy0 = wy0;
    uint8_t* out = p.desc + (size_t)frame * p.desc_stride + (size_t)row * p.desc_pitch;
    ef_bad_describe(I, a, t, p.desc_bytes * 8, L.w, L.h, out, lane);
}

// the same window form for a flat keypoint array whose keypoints all have integer coordinates and size 31 at scale 1
// (the 5 x N GpuMat compute path, cuda_efficient_features.cu:250-263): no full-frame integral image
__global__ void __launch_bounds__(EF_DESC_WARPS * 32) ef_bad_flat_window_kernel(const EfDescJob job, const EfBadTables t)
{
    extern __shared__ __align__(16) unsigned s_win_all[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int i = blockIdx.x * EF_DESC_WARPS + warp;
    if (i >= job.n) return;
    const float4 k = job.kpts[i];
    const int kx = (int)k.x, ky = (int)k.y;
    unsigned* __restrict__ W = s_win_all + warp * ((EF_BW_ROWS + 1) * EF_BW_WORDS);
    const int gx0 = (kx - EF_BW_HALF) & ~15, wy0 = ky - EF_BW_HALF;
    ef_bad_window_integral(W, job.img, job.pitch, job.h, gx0, wy0, lane);
    const EfBadAffine a = ef_bad_affine(k.x, k.y, k.z, k.w, job.scale, job.w, job.h);
    EfWindowIntegral16 I; I.P = reinterpret_cast<const unsigned short*>(W); I.gx0 = gx0; I.wy0 = wy0;
    ef_bad_describe(I, a, t, job.nbits, job.w, job.h, job.desc + (size_t)i * job.desc_pitch, lane);
}

void ef_launch_bad_pipe(const EfPipe& p, const EfBadTables& t, cudaStream_t s)
{
    if (p.total_kpt_blocks <= 0) return;
    const size_t smem = (size_t)EF_DESC_WARPS * (EF_BW_ROWS + 1) * EF_BW_WORDS * sizeof(unsigned);
    // function attributes are per device (ef_mg_* drives several devices from one process)
    static unsigned long long configured = 0;
    int dev = 0;
    cudaGetDevice(&dev);
    if (!((__atomic_load_n(&configured, __ATOMIC_RELAXED) >> (dev & 63)) & 1ull)) {
        cudaFuncSetAttribute(ef_bad_pipe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        __atomic_fetch_or(&configured, 1ull << (dev & 63), __ATOMIC_RELAXED);
    }
    ef_bad_pipe_kernel<<<dim3(ef_div_up(p.total_kpt_blocks, p.shard_n), p.nframes), EF_DESC_WARPS * 32, smem, s>>>(p, t);
    EF_COUNT_LAUNCH(1);
}

// =================================================================================================
// convertKeypointsKernel (cuda_efficient_features.cu:250-263): 5 x N rows -> (x, y, PATCH_SIZE, angle)
// =================================================================================================
__global__ void ef_convert_rows_kernel(const uint8_t* __restrict__ kpts5, size_t pitch, int n, float4* __restrict__ out)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const short2 pt = reinterpret_cast<const short2*>(kpts5 + (size_t)EF_LOCATION_ROW * pitch)[i];
    float4 k;
    k.x = pt.x; k.y = pt.y; k.z = EF_PATCH_SIZE;
    k.w = reinterpret_cast<const float*>(kpts5 + (size_t)EF_ANGLE_ROW * pitch)[i];
    out[i] = k;
}

void ef_launch_convert_rows(const float* kpts5, size_t kpts_pitch, int n, float4* out, cudaStream_t s)
{
    if (n <= 0) return;
    ef_convert_rows_kernel<<<ef_div_up(n, 256), 256, 0, s>>>(reinterpret_cast<const uint8_t*>(kpts5), kpts_pitch, n, out);
    EF_COUNT_LAUNCH(1);
}

void ef_launch_bad_flat_window(const EfDescJob& job, const EfBadTables& t, cudaStream_t s)
{
    const size_t smem = (size_t)EF_DESC_WARPS * (EF_BW_ROWS + 1) * EF_BW_WORDS * sizeof(unsigned);
    static unsigned long long configured = 0;   // function attributes are per device
    int dev = 0;
    cudaGetDevice(&dev);
    if (!((__atomic_load_n(&configured, __ATOMIC_RELAXED) >> (dev & 63)) & 1ull)) {
        cudaFuncSetAttribute(ef_bad_flat_window_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        __atomic_fetch_or(&configured, 1ull << (dev & 63), __ATOMIC_RELAXED);
    }
    ef_bad_flat_window_kernel<<<ef_div_up(job.n, EF_DESC_WARPS), EF_DESC_WARPS * 32, smem, s>>>(job, t);
    EF_COUNT_LAUNCH(1);
}
