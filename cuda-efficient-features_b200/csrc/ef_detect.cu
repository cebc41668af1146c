// ef_detect.cu -- detector stages of the B200 detectAndCompute path (sm_100a).
//
// Stage             replaces (reference, modules/cuda_efficient_features/src/)
//   pyramid         calcImagePyramid + cv::cuda::resize            cuda_efficient_features.cpp:136-157
//   score           createMask + calcKeypointsKernel + calcResponsesKernel
//                                                                   cuda_fast.cu:168-222, cuda_efficient_features.cu:99-139,218-225
//   nms             nptPerBlock/assignIndex/radiusSuppression       cuda_efficient_features.cu:174-216,281-342
//   compact         (atomic append in the reference, :212)          device-wide raster-order prefix-sum compaction
//   select          limitPoints (thrust::sort_by_key)               cuda_efficient_features.cu:344-358
//   angle_pack      calcAnglesKernel + scalePointsKernel + copyTo   cuda_efficient_features.cu:141-172,227-248; .cpp:310-311
//   blur            cv::cuda::createGaussianFilter(7x7, sigma 2)    cuda_efficient_features.cpp:193,305
//
// No host synchronisation anywhere: counts live in EfLevelCounters / rowcnt on the device, every
// launch is sized from capacities.  All levels of all frames of a batch go through ONE launch per
// stage (tile tables in EfPipe), except the pyramid chain whose level s needs level s-1.
#include "ef_common.cuh"

#include <cuda_fp16.h>
#include <cstdlib>
#include <cstring>

// =================================================================================================
// pyramid: bilinear x(1/scaleFactor) chain, one launch per level over the whole batch
// =================================================================================================
__global__ void __launch_bounds__(256) ef_resize_kernel(const __grid_constant__ EfPipe p, const int level)
{
    const EfLevel& L = p.lv[level];
    const EfLevel& S = p.lv[level - 1];
    const int frame = blockIdx.z;
    const int x0 = (blockIdx.x * 64 + threadIdx.x) * 4;
    const int y = blockIdx.y * 4 + threadIdx.y;
    if (y >= L.h || x0 >= L.w) return;

    int spitch;
    const uint8_t* __restrict__ src = ef_level_image(p, frame, level - 1, spitch);
    uint8_t* dst = ef_ws(p, frame, L.img_off) + (size_t)y * L.img_pitch;

    const float sy = (float)y * L.ry;
    const int y1 = __float2int_rd(sy);
    const int y2 = y1 + 1;
    const int y2r = min(y2, S.h - 1);
    const float wy1 = (float)y2 - sy, wy2 = sy - (float)y1;
    const uint8_t* __restrict__ r1 = src + (size_t)y1 * spitch;
    const uint8_t* __restrict__ r2 = src + (size_t)y2r * spitch;

    unsigned packed = 0;
#pragma unroll
    for (int i = 0; i < 4; i++) {
        const int x = x0 + i;
        if (x < L.w) {
            const float sx = (float)x * L.rx;
            const int x1 = __float2int_rd(sx);
            const int x2 = x1 + 1;
            const int x2r = min(x2, S.w - 1);
            const float wx1 = (float)x2 - sx, wx2 = sx - (float)x1;
            float out = 0.f;
            out = fmaf((float)r1[x1], wx1 * wy1, out);
            out = fmaf((float)r1[x2r], wx2 * wy1, out);
            out = fmaf((float)r2[x1], wx1 * wy2, out);
            out = fmaf((float)r2[x2r], wx2 * wy2, out);
            packed |= ef_sat_u8_rne(out) << (8 * i);
        }
    }
    *reinterpret_cast<unsigned*>(dst + x0) = packed; // img_pitch is a multiple of 128
}

// Tiled form (resize ratios <= 1.21 x 1.25, i.e. the usual 1.2 pyramid): one CTA = 128 x 32 output pixels.  The source window
// (<= 160 x 43 pixels) is staged in shared memory ONCE, read with aligned 32-bit words and converted to fp32 there (one
// conversion per source pixel instead of four per output pixel).  The column one past the right edge and the row one past the
// bottom edge are staged as copies of the last column / row, so the x2r / y2r clamps of the direct kernel vanish: the four
// taps of a pixel are s[o], s[o+1], s[o+RWP], s[o+RWP+1] -- one address per pixel, immediate offsets for the rest.
// A warp owns 4 output rows, a lane 4 adjacent output columns whose source offsets and horizontal weights are computed once;
// two pixels advance together in packed fp32 (FMUL2 / FFMA2: mul.rn.f32x2, fma.rn.f32x2 -- per-lane IEEE, so the arithmetic
// is that of ef_resize_kernel bit for bit).
#define RS_TW 128
#define RS_TH 32
#define RS_RWP 160
#define RS_RH 43


__global__ void __launch_bounds__(256) ef_resize_tiled_kernel(const __grid_constant__ EfPipe p, const int level)
{
    __shared__ __align__(16) float s_src[RS_RH * RS_RWP];

    const EfLevel& L = p.lv[level];
    const EfLevel& S = p.lv[level - 1];
    const int frame = blockIdx.z;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int X0 = blockIdx.x * RS_TW, Y0 = blockIdx.y * RS_TH;

    int spitch;
    const uint8_t* __restrict__ src = ef_level_image(p, frame, level - 1, spitch);
    const bool aligned = ((reinterpret_cast<uintptr_t>(src) | (unsigned)spitch) & 3u) == 0;

    // source window of the tile: columns sx0 .. x1(last)+1, rows sy0 .. y1(last)+1 (the +1 may be one past the edge: clamped copy)
    const int xlast = min(X0 + RS_TW, L.w) - 1, ylast = min(Y0 + RS_TH, L.h) - 1;
    const int sx0 = __float2int_rd((float)X0 * L.rx) & ~3;
    const int sy0 = __float2int_rd((float)Y0 * L.ry);
    const int nwords = min((__float2int_rd((float)xlast * L.rx) + 1 - sx0) / 4 + 1, RS_RWP / 4);
    const int nrows = min(__float2int_rd((float)ylast * L.ry) + 1 - sy0 + 1, RS_RH);
    const unsigned inv = 0xffffffffu / (unsigned)nwords + 1u; // i / nwords == umulhi(i, inv) for the small i used here (nwords >= 2)
#pragma unroll 2
    for (int i = tid; i < nrows * nwords; i += 256) {
        const int row = nwords == 1 ? i : (int)__umulhi((unsigned)i, inv), wx = i - row * nwords;
        const uint8_t* rp = src + (size_t)min(sy0 + row, S.h - 1) * spitch;
        const int gx = sx0 + 4 * wx;
        unsigned word;
        if (aligned && gx + 3 < S.w) word = *reinterpret_cast<const unsigned*>(rp + gx);
        else {
            word = rp[min(gx, S.w - 1)];
            word |= (unsigned)rp[min(gx + 1, S.w - 1)] << 8;
            word |= (unsigned)rp[min(gx + 2, S.w - 1)] << 16;
            word |= (unsigned)rp[min(gx + 3, S.w - 1)] << 24;
        }
        float4 f;
        f.x = __uint_as_float(__byte_perm(word, 0x4B000000u, 0x7540)) - 8388608.f; // 0x4B0000bb = 2^23 + b, exact
        f.y = __uint_as_float(__byte_perm(word, 0x4B000000u, 0x7541)) - 8388608.f;
        f.z = __uint_as_float(__byte_perm(word, 0x4B000000u, 0x7542)) - 8388608.f;
        f.w = __uint_as_float(__byte_perm(word, 0x4B000000u, 0x7543)) - 8388608.f;
        *reinterpret_cast<float4*>(&s_src[row * RS_RWP + 4 * wx]) = f;
    }
    __syncthreads();

    // per-lane column geometry (4 adjacent output pixels)
    const int x0 = X0 + 4 * lane;
    if (x0 >= L.w) return;
    int o[4];
    float wx1[4], wx2[4];
#pragma unroll
    for (int j = 0; j < 4; j++) {
        const int x = min(x0 + j, L.w - 1);
        const float sx = (float)x * L.rx;
        const int x1 = __float2int_rd(sx);
        o[j] = min(x1 - sx0, RS_RWP - 2);
        wx1[j] = (float)(x1 + 1) - sx; wx2[j] = sx - (float)x1;
    }
    const unsigned long long wx1a = ef_pack2(wx1[0], wx1[1]), wx1b = ef_pack2(wx1[2], wx1[3]);
    const unsigned long long wx2a = ef_pack2(wx2[0], wx2[1]), wx2b = ef_pack2(wx2[2], wx2[3]);
    uint8_t* dst = ef_ws(p, frame, L.img_off);
#pragma unroll
    for (int r = 0; r < 4; r++) {
        const int y = Y0 + 4 * warp + r;
        if (y >= L.h) break;
        const float sy = (float)y * L.ry;
        const int y1 = __float2int_rd(sy);
        const float wy1 = (float)(y1 + 1) - sy, wy2 = sy - (float)y1;
        const float* ra = s_src + min(y1 - sy0, RS_RH - 2) * RS_RWP;
        const unsigned long long wy1p = ef_pack2(wy1, wy1), wy2p = ef_pack2(wy2, wy2);
        const float* q0 = ra + o[0]; const float* q1 = ra + o[1]; const float* q2 = ra + o[2]; const float* q3 = ra + o[3];
        // pixels (0,1) and (2,3) advance together; tap order = ef_resize_kernel: (y1,x1) (y1,x2) (y2,x1) (y2,x2)
        unsigned long long a = ef_mul2(ef_pack2(q0[0], q1[0]), ef_mul2(wx1a, wy1p));
        unsigned long long b = ef_mul2(ef_pack2(q2[0], q3[0]), ef_mul2(wx1b, wy1p));
        a = ef_fma2(ef_pack2(q0[1], q1[1]), ef_mul2(wx2a, wy1p), a);
        b = ef_fma2(ef_pack2(q2[1], q3[1]), ef_mul2(wx2b, wy1p), b);
        a = ef_fma2(ef_pack2(q0[RS_RWP], q1[RS_RWP]), ef_mul2(wx1a, wy2p), a);
        b = ef_fma2(ef_pack2(q2[RS_RWP], q3[RS_RWP]), ef_mul2(wx1b, wy2p), b);
        a = ef_fma2(ef_pack2(q0[RS_RWP + 1], q1[RS_RWP + 1]), ef_mul2(wx2a, wy2p), a);
        b = ef_fma2(ef_pack2(q2[RS_RWP + 1], q3[RS_RWP + 1]), ef_mul2(wx2b, wy2p), b);
        float o0, o1, o2, o3;
        asm("mov.b64 {%0, %1}, %2;" : "=f"(o0), "=f"(o1) : "l"(a));
        asm("mov.b64 {%0, %1}, %2;" : "=f"(o2), "=f"(o3) : "l"(b));
        const unsigned packed = ef_sat_u8_rne(o0) | (ef_sat_u8_rne(o1) << 8) | (ef_sat_u8_rne(o2) << 16) | (ef_sat_u8_rne(o3) << 24);
        *reinterpret_cast<unsigned*>(dst + (size_t)y * L.img_pitch + x0) = packed; // img_pitch is a multiple of 128
    }
}

// 16-byte form of the tiled kernel (source image base, frame stride and pitch 16-byte aligned; ratios <= 1.21 x 1.22): one CTA =
// 128 x 64 output pixels, i.e. 32 pixels per thread, so that the per-thread column geometry and the CTA-wide staging are paid
// once per 32 pixels instead of once per 16.
//  * staging: the source window (<= 176 x 78 pixels, first column rounded down to a multiple of 16) is read with one 16-byte load
//    per 16 pixels, all loads of a thread issued back to back, and converted with packed adds (add.rn.f32x2).  A warp stages a
//    block of 8 rows x 4 chunks (lane = row + 8 * chunk): 64 contiguous bytes per row on the global side, and on the shared side
//    the eight lanes of a quarter-warp write eight rows whose 720-byte pitch puts their 16-byte stores on distinct bank groups.
//  * row geometry (staged row offset, wy1, wy2, "top row must be reloaded") of the 64 output rows is computed once per CTA into
//    a table read back as one broadcast 16-byte load per row.
//  * a warp owns 8 CONSECUTIVE output rows: the bottom source row of output row y is the top source row of y + 1 whenever the
//    source row advances by one (5 rows out of 6 at ratio 1.2), so its taps stay in registers (two register sets swapping roles
//    in a 2x unrolled loop; the reload is a warp-uniform branch): 2.4 instead of 4 shared-memory loads per pixel -- the loads,
//    two-way bank-conflicted by the 4.8-word lane stride, are what bounds this kernel.
// Tap arithmetic = ef_resize_kernel bit for bit (same products, same FMA chain, per-lane IEEE packed fp32).
#define RS3_TW 128
#define RS3_TH 64
#define RS3_RWP 180
#define RS3_RH 78
#define RS3_CH 11     // 16-pixel chunks per staged row
#define RS3_RG ((RS3_RH + 7) / 8)
struct __align__(16) EfRowGeo { int off; float wy1, wy2; int reload; };
#define RS3_SMEM (RS3_RH * RS3_RWP * 4 + RS3_TH * 16)
static_assert(RS3_RG * 3 <= 4 * 8, "staging: at most four 8-row x 4-chunk blocks per warp");

__device__ __forceinline__ float4 ef_u8x4_to_f32x4(unsigned word)
{
    // 0x4B0000bb = 2^23 + b: subtracting 2^23 is exact; two bytes per packed add
    const unsigned long long m = ef_pack2(-8388608.f, -8388608.f);
    unsigned long long lo = ef_pack2(__uint_as_float(__byte_perm(word, 0x4B000000u, 0x7540)), __uint_as_float(__byte_perm(word, 0x4B000000u, 0x7541)));
    unsigned long long hi = ef_pack2(__uint_as_float(__byte_perm(word, 0x4B000000u, 0x7542)), __uint_as_float(__byte_perm(word, 0x4B000000u, 0x7543)));
    asm("add.rn.f32x2 %0, %0, %1;" : "+l"(lo) : "l"(m));
    asm("add.rn.f32x2 %0, %0, %1;" : "+l"(hi) : "l"(m));
    float4 f;
    asm("mov.b64 {%0, %1}, %2;" : "=f"(f.x), "=f"(f.y) : "l"(lo));
    asm("mov.b64 {%0, %1}, %2;" : "=f"(f.z), "=f"(f.w) : "l"(hi));
    return f;
}

struct EfTaps { unsigned long long a0, b0, a1, b1; };   // (px0, px1) / (px2, px3) left taps, then right taps, of one source row

__global__ void __launch_bounds__(256, 4) ef_resize_tiled16_kernel(const __grid_constant__ EfPipe p, const int level)
{
    extern __shared__ __align__(16) float s_dyn[];
    float* __restrict__ s_src = s_dyn;
    EfRowGeo* __restrict__ s_row = reinterpret_cast<EfRowGeo*>(s_dyn + RS3_RH * RS3_RWP);

    const EfLevel& L = p.lv[level];
    const EfLevel& S = p.lv[level - 1];
    const int frame = blockIdx.z;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int X0 = blockIdx.x * RS3_TW, Y0 = blockIdx.y * RS3_TH;

    int spitch;
    const uint8_t* __restrict__ src = ef_level_image(p, frame, level - 1, spitch);

    const int xlast = min(X0 + RS3_TW, L.w) - 1, ylast = min(Y0 + RS3_TH, L.h) - 1;
    const int sx0 = __float2int_rd((float)X0 * L.rx) & ~15;
    const int sy0 = __float2int_rd((float)Y0 * L.ry);
    const int nch = min((__float2int_rd((float)xlast * L.rx) + 1 - sx0) / 16 + 1, RS3_CH);
    const int nrows = min(__float2int_rd((float)ylast * L.ry) + 1 - sy0 + 1, RS3_RH);

    // row geometry of the tile's output rows (rows past the level's last one repeat it; they are never written)
    if (tid < RS3_TH) {
        const int y = min(Y0 + tid, L.h - 1);
        const float sy = (float)y * L.ry;
        const int y1 = __float2int_rd(sy);
        const int yp = min(Y0 + tid - 1, L.h - 1);
        const int offp = min(__float2int_rd((float)yp * L.ry) - sy0, RS3_RH - 2) * (RS3_RWP * 4);
        EfRowGeo g;
        g.off = min(y1 - sy0, RS3_RH - 2) * (RS3_RWP * 4); // bytes
        g.wy1 = (float)(y1 + 1) - sy; g.wy2 = sy - (float)y1;
        g.reload = ((tid & 7) == 0 || g.off != offp + RS3_RWP * 4) ? 1 : 0;  // first row of a warp, or the source row advanced by two
        s_row[tid] = g;
    }
    // source window: staged row r = source row min(sy0 + r, S.h - 1), staged column c = source column min(sx0 + c, S.w - 1).
    // Chunks that reach past the last source column (right image edge only) are staged byte-wise with the clamp afterwards.
    {
        uint4 v[4];
        int dsto[4];
#pragma unroll
        for (int k = 0; k < 4; k++) {
            const int b = warp + 8 * k;                      // block of 8 rows x 4 chunks
            const int rg = b / 3, cg = b - 3 * rg;
            const int row = 8 * rg + (lane & 7), c = 4 * cg + (lane >> 3);
            const int gx = sx0 + 16 * c;
            const bool fast = row < nrows && c < nch && gx + 15 < S.w;
            dsto[k] = fast ? row * RS3_RWP + 16 * c : -1;
            if (fast) v[k] = __ldg(reinterpret_cast<const uint4*>(src + (size_t)min(sy0 + row, S.h - 1) * spitch + gx));
        }
#pragma unroll
        for (int k = 0; k < 4; k++) {
            if (dsto[k] >= 0) {
                float4* d = reinterpret_cast<float4*>(s_src + dsto[k]);
                d[0] = ef_u8x4_to_f32x4(v[k].x); d[1] = ef_u8x4_to_f32x4(v[k].y); d[2] = ef_u8x4_to_f32x4(v[k].z); d[3] = ef_u8x4_to_f32x4(v[k].w);
            }
        }
        const int cedge = (S.w - sx0) >> 4;     // first chunk with gx + 15 >= S.w
        if (cedge < nch) {
            const int nedge = nch - cedge;
            for (int i = tid; i < nrows * nedge * 4; i += 256) {
                const int row = i / (nedge * 4), wq = i - row * (nedge * 4);
                const uint8_t* rp = src + (size_t)min(sy0 + row, S.h - 1) * spitch;
                const int g0 = sx0 + 16 * cedge + 4 * wq;
                const unsigned wd = (unsigned)rp[min(g0, S.w - 1)] | ((unsigned)rp[min(g0 + 1, S.w - 1)] << 8) |
                                    ((unsigned)rp[min(g0 + 2, S.w - 1)] << 16) | ((unsigned)rp[min(g0 + 3, S.w - 1)] << 24);
                *reinterpret_cast<float4*>(s_src + row * RS3_RWP + 16 * cedge + 4 * wq) = ef_u8x4_to_f32x4(wd);
            }
        }
    }
    __syncthreads();

    const int x0 = X0 + 4 * lane;
    if (x0 >= L.w) return;
    int o[4];
    float wx1[4], wx2[4];
#pragma unroll
    for (int j = 0; j < 4; j++) {
        const int x = min(x0 + j, L.w - 1);
        const float sx = (float)x * L.rx;
        const int x1 = __float2int_rd(sx);
        o[j] = min(x1 - sx0, RS3_RWP - 2);
        wx1[j] = (float)(x1 + 1) - sx; wx2[j] = sx - (float)x1;
    }
    const unsigned long long wx1a = ef_pack2(wx1[0], wx1[1]), wx1b = ef_pack2(wx1[2], wx1[3]);
    const unsigned long long wx2a = ef_pack2(wx2[0], wx2[1]), wx2b = ef_pack2(wx2[2], wx2[3]);
    const int r0 = (RS3_TH / 8) * warp;
    const int nr = min(RS3_TH / 8, L.h - (Y0 + r0));       // rows of this warp inside the level
    uint8_t* dst = ef_ws(p, frame, L.img_off) + (size_t)(Y0 + r0) * L.img_pitch + x0;
    const size_t dpitch = (size_t)L.img_pitch;
    // tap pointers of row offset 0; the per-row byte offset is added to each (one integer add per pointer and row)
    const char* qb0 = reinterpret_cast<const char*>(s_src + o[0]); const char* qb1 = reinterpret_cast<const char*>(s_src + o[1]);
    const char* qb2 = reinterpret_cast<const char*>(s_src + o[2]); const char* qb3 = reinterpret_cast<const char*>(s_src + o[3]);

    auto load_taps = [&](EfTaps& t, int offb) {
        const float* q0 = reinterpret_cast<const float*>(qb0 + offb); const float* q1 = reinterpret_cast<const float*>(qb1 + offb);
        const float* q2 = reinterpret_cast<const float*>(qb2 + offb); const float* q3 = reinterpret_cast<const float*>(qb3 + offb);
        t.a0 = ef_pack2(q0[0], q1[0]); t.b0 = ef_pack2(q2[0], q3[0]);
        t.a1 = ef_pack2(q0[1], q1[1]); t.b1 = ef_pack2(q2[1], q3[1]);
    };
    auto emit_row = [&](const EfTaps& top, const EfTaps& bot, const EfRowGeo& g) {
        const unsigned long long wy1p = ef_pack2(g.wy1, g.wy1), wy2p = ef_pack2(g.wy2, g.wy2);
        // tap order = ef_resize_kernel: (y1,x1) (y1,x2) (y2,x1) (y2,x2)
        unsigned long long a = ef_mul2(top.a0, ef_mul2(wx1a, wy1p));
        unsigned long long b = ef_mul2(top.b0, ef_mul2(wx1b, wy1p));
        a = ef_fma2(top.a1, ef_mul2(wx2a, wy1p), a);
        b = ef_fma2(top.b1, ef_mul2(wx2b, wy1p), b);
        a = ef_fma2(bot.a0, ef_mul2(wx1a, wy2p), a);
        b = ef_fma2(bot.b0, ef_mul2(wx1b, wy2p), b);
        a = ef_fma2(bot.a1, ef_mul2(wx2a, wy2p), a);
        b = ef_fma2(bot.b1, ef_mul2(wx2b, wy2p), b);
        float o0, o1, o2, o3;
        asm("mov.b64 {%0, %1}, %2;" : "=f"(o0), "=f"(o1) : "l"(a));
        asm("mov.b64 {%0, %1}, %2;" : "=f"(o2), "=f"(o3) : "l"(b));
        const unsigned packed = __byte_perm(__byte_perm(ef_sat_u8_rne(o0), ef_sat_u8_rne(o1), 0x0040), __byte_perm(ef_sat_u8_rne(o2), ef_sat_u8_rne(o3), 0x0040), 0x5410);
        *reinterpret_cast<unsigned*>(dst) = packed; // img_pitch is a multiple of 128
        dst += dpitch;
    };
    EfTaps X, Y;
#pragma unroll
    for (int r = 0; r < RS3_TH / 8; r += 2) {
        if (r >= nr) break;
        {
            const EfRowGeo g = s_row[r0 + r];
            if (g.reload) load_taps(X, g.off);              // warp-uniform
            load_taps(Y, g.off + RS3_RWP * 4);
            emit_row(X, Y, g);
        }
        if (r + 1 >= nr) break;
        {
            const EfRowGeo g = s_row[r0 + r + 1];
            if (g.reload) load_taps(Y, g.off);
            load_taps(X, g.off + RS3_RWP * 4);
            emit_row(Y, X, g);
        }
    }
}

// TMA form of the tiled kernel (same tile, same warp / lane / row mapping, same tap arithmetic): the u8 source window is staged by
// ONE cp.async.bulk.tensor.3d (UTMALDG) -- no staging loop, no u8 -> fp32 conversion pass, 13.7 KB instead of 56 KB of shared memory --
// and every lane reads the <= 6 source bytes its four output pixels need from a source row as three aligned words, funnel-shifted
// to the first needed byte; two byte permutes with per-lane selectors (computed once) pick the four left and the four right taps,
// which are then expanded to fp32 (0x4B0000bb = 2^23 + b, packed subtract).  Against the fp32 window: 3 word loads per lane and
// source row at a 4.8-byte lane stride (mostly distinct banks) instead of 8 tap loads at a 4.8-WORD stride (2-way conflicts).
// Needs a 16-byte aligned source (base, pitch, frame stride) and a level whose taps never clamp at the right / bottom edge (any
// down-scaling ratio: checked exactly on the host); otherwise ef_resize_tiled16_kernel runs.
#define RS4_PITCH EF_RS_BOX_W
#define RS4_SMEM (EF_RS_BOX_H * EF_RS_BOX_W + 16 + RS3_TH * 16 + 16)

template <int NB>   // resident CTAs per SM the register allocation aims at (4 / 5 / 6 measured: 0.253 / 0.249 / 0.245 ms per 16 frames)
__global__ void __launch_bounds__(256, NB) ef_resize_tma_kernel(const __grid_constant__ EfTmaMaps maps, const __grid_constant__ EfPipe p, const int level)
{
    extern __shared__ __align__(128) unsigned char s_raw4[];
    unsigned char* __restrict__ s_u8 = s_raw4;                                                        // [78][176] + 16 bytes of slack
    EfRowGeo* __restrict__ s_row = reinterpret_cast<EfRowGeo*>(s_raw4 + EF_RS_BOX_H * EF_RS_BOX_W + 16);
    unsigned long long* s_mbar = reinterpret_cast<unsigned long long*>(s_raw4 + EF_RS_BOX_H * EF_RS_BOX_W + 16 + RS3_TH * 16);

    const EfLevel& L = p.lv[level];
    const int frame = blockIdx.z;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int X0 = blockIdx.x * RS3_TW, Y0 = blockIdx.y * RS3_TH;
    const int sx0 = __float2int_rd((float)X0 * L.rx) & ~15;
    const int sy0 = __float2int_rd((float)Y0 * L.ry);

    const unsigned mbar = ef_smem_addr(s_mbar);
    if (tid == 0) {
        ef_mbar_init(mbar, 1);
        ef_mbar_expect_tx(mbar, EF_RS_BOX_W * EF_RS_BOX_H);
        // programmatic dependent launch: this grid may have been scheduled while the previous level's grid was still draining;
        // its output (our source) is complete once this returns (no-op for a normally serialised launch)
        asm volatile("griddepcontrol.wait;" ::: "memory");
        ef_tma_load_3d(ef_smem_addr(s_u8), &maps.resize_src[level], sx0, sy0, frame, mbar);
    }
    // row geometry of the tile's output rows while the window is in flight
    if (tid < RS3_TH) {
        const int y = min(Y0 + tid, L.h - 1);
        const float sy = (float)y * L.ry;
        const int y1 = __float2int_rd(sy);
        const int yp = min(Y0 + tid - 1, L.h - 1);
        const int offp = min(__float2int_rd((float)yp * L.ry) - sy0, EF_RS_BOX_H - 2) * RS4_PITCH;
        EfRowGeo g;
        g.off = min(y1 - sy0, EF_RS_BOX_H - 2) * RS4_PITCH; // bytes
        g.wy1 = (float)(y1 + 1) - sy; g.wy2 = sy - (float)y1;
        g.reload = ((tid & 7) == 0 || g.off != offp + RS4_PITCH) ? 1 : 0;
        s_row[tid] = g;
    }
    __syncthreads();                    // mbarrier initialised, row table written
    ef_mbar_wait(mbar, 0);

    const int x0 = X0 + 4 * lane;
    if (x0 >= L.w) return;
    int o[4];
    float wx1[4], wx2[4];
#pragma unroll
    for (int j = 0; j < 4; j++) {
        const int x = min(x0 + j, L.w - 1);
        const float sx = (float)x * L.rx;
        const int x1 = __float2int_rd(sx);
        o[j] = min(x1 - sx0, EF_RS_BOX_W - 2);
        wx1[j] = (float)(x1 + 1) - sx; wx2[j] = sx - (float)x1;
    }
    const unsigned long long wx1a = ef_pack2(wx1[0], wx1[1]), wx1b = ef_pack2(wx1[2], wx1[3]);
    const unsigned long long wx2a = ef_pack2(wx2[0], wx2[1]), wx2b = ef_pack2(wx2[2], wx2[3]);
    // byte window of the lane: starts at o[0]; left taps at o[j] - o[0] in 0..4, right taps one further (<= 5)
    const int wb = o[0] & ~3;
    const unsigned sh = 8u * (unsigned)(o[0] & 3);
    unsigned selL = 0, selR = 0;
#pragma unroll
    for (int j = 0; j < 4; j++) {
        const unsigned d = (unsigned)min(max(o[j] - o[0], 0), 6);   // 0..4 for real pixels; clamped copies of the last column stay in range
        selL |= d << (4 * j); selR |= (d + 1) << (4 * j);
    }
    const unsigned char* wbase = s_u8 + wb;
    const unsigned long long m23 = ef_pack2(-8388608.f, -8388608.f);

    const int r0 = (RS3_TH / 8) * warp;
    const int nr = min(RS3_TH / 8, L.h - (Y0 + r0));
    uint8_t* dst = ef_ws(p, frame, L.img_off) + (size_t)(Y0 + r0) * L.img_pitch + x0;
    const size_t dpitch = (size_t)L.img_pitch;

    auto load_taps = [&](EfTaps& t, int offb) {
        const unsigned* wp = reinterpret_cast<const unsigned*>(wbase + offb);
        const unsigned W0 = wp[0], W1 = wp[1], W2 = wp[2];
        const unsigned A = __funnelshift_r(W0, W1, sh), B = __funnelshift_r(W1, W2, sh);   // bytes o[0] .. o[0]+7 of the row
        const unsigned tl = __byte_perm(A, B, selL), tr = __byte_perm(A, B, selR);
#define RS4_F(wd, i) __uint_as_float(__byte_perm(wd, 0x4B000000u, 0x7540 + (i)))
        t.a0 = ef_add2(ef_pack2(RS4_F(tl, 0), RS4_F(tl, 1)), m23); t.b0 = ef_add2(ef_pack2(RS4_F(tl, 2), RS4_F(tl, 3)), m23);
        t.a1 = ef_add2(ef_pack2(RS4_F(tr, 0), RS4_F(tr, 1)), m23); t.b1 = ef_add2(ef_pack2(RS4_F(tr, 2), RS4_F(tr, 3)), m23);
#undef RS4_F
    };
    auto emit_row = [&](const EfTaps& top, const EfTaps& bot, const EfRowGeo& g) {
        const unsigned long long wy1p = ef_pack2(g.wy1, g.wy1), wy2p = ef_pack2(g.wy2, g.wy2);
        // tap order = ef_resize_kernel: (y1,x1) (y1,x2) (y2,x1) (y2,x2)
        unsigned long long a = ef_mul2(top.a0, ef_mul2(wx1a, wy1p));
        unsigned long long b = ef_mul2(top.b0, ef_mul2(wx1b, wy1p));
        a = ef_fma2(top.a1, ef_mul2(wx2a, wy1p), a);
        b = ef_fma2(top.b1, ef_mul2(wx2b, wy1p), b);
        a = ef_fma2(bot.a0, ef_mul2(wx1a, wy2p), a);
        b = ef_fma2(bot.b0, ef_mul2(wx1b, wy2p), b);
        a = ef_fma2(bot.a1, ef_mul2(wx2a, wy2p), a);
        b = ef_fma2(bot.b1, ef_mul2(wx2b, wy2p), b);
        float o0, o1, o2, o3;
        asm("mov.b64 {%0, %1}, %2;" : "=f"(o0), "=f"(o1) : "l"(a));
        asm("mov.b64 {%0, %1}, %2;" : "=f"(o2), "=f"(o3) : "l"(b));
        const unsigned packed = __byte_perm(__byte_perm(ef_sat_u8_rne(o0), ef_sat_u8_rne(o1), 0x0040), __byte_perm(ef_sat_u8_rne(o2), ef_sat_u8_rne(o3), 0x0040), 0x5410);
        *reinterpret_cast<unsigned*>(dst) = packed; // img_pitch is a multiple of 128
        dst += dpitch;
    };
    EfTaps X, Y;
#pragma unroll
    for (int r = 0; r < RS3_TH / 8; r += 2) {
        if (r >= nr) break;
        {
            const EfRowGeo g = s_row[r0 + r];
            if (g.reload) load_taps(X, g.off);              // warp-uniform
            load_taps(Y, g.off + RS4_PITCH);
            emit_row(X, Y, g);
        }
        if (r + 1 >= nr) break;
        {
            const EfRowGeo g = s_row[r0 + r + 1];
            if (g.reload) load_taps(Y, g.off);
            load_taps(X, g.off + RS4_PITCH);
            emit_row(Y, X, g);
        }
    }
}

static const bool g_ef_resize_legacy = getenv("EF_RESIZE") && !strcmp(getenv("EF_RESIZE"), "legacy");
void ef_launch_pyramid(const EfPipe& p, const EfTmaMaps* maps, cudaStream_t s)
{
    static const bool use_tma = !(getenv("EF_RESIZE_TMA") && atoi(getenv("EF_RESIZE_TMA")) == 0);   // A/B switch; default: TMA staging
    bool tma_prev = false;   // the previous launch of this chain was the TMA kernel (which contains the griddepcontrol.wait)
    for (int l = 1; l < p.nlevels; l++) {
        const bool src16 = l > 1 || ((reinterpret_cast<uintptr_t>(p.img0) | p.img0_stride | (unsigned long long)p.img0_pitch) & 15ull) == 0;
        // the attribute is per device (ef_mg_* drives several devices from one process): one bit per device
        static unsigned long long configured = 0;
        int dev = 0;
        cudaGetDevice(&dev);
        if (!((__atomic_load_n(&configured, __ATOMIC_RELAXED) >> (dev & 63)) & 1ull)) {
            cudaFuncSetAttribute(ef_resize_tiled16_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, RS3_SMEM);
            __atomic_fetch_or(&configured, 1ull << (dev & 63), __ATOMIC_RELAXED);
        }
        if (maps && use_tma && ((maps->resize_src_ok >> l) & 1u) && p.lv[l].rx <= 1.21f && p.lv[l].ry <= 1.22f && !g_ef_resize_legacy) {
            const dim3 grid(ef_div_up(p.lv[l].w, RS3_TW), ef_div_up(p.lv[l].h, RS3_TH), p.nframes);
            // levels >= 2 follow another resize launch: programmatic dependent launch hides the launch gap of the 7-deep chain
            static const bool pdl = !(getenv("EF_RESIZE_PDL") && atoi(getenv("EF_RESIZE_PDL")) == 0);
            cudaLaunchConfig_t cfg = {};
            cfg.gridDim = grid; cfg.blockDim = dim3(256); cfg.dynamicSmemBytes = RS4_SMEM; cfg.stream = s;
            cudaLaunchAttribute attr[1];
            attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
            attr[0].val.programmaticStreamSerializationAllowed = 1;
            cfg.attrs = attr; cfg.numAttrs = (pdl && l >= 2 && tma_prev) ? 1 : 0;
            cudaLaunchKernelEx(&cfg, ef_resize_tma_kernel<6>, *maps, p, l);
            tma_prev = true;
            EF_COUNT_LAUNCH(1);
            continue;
        } else if (src16 && p.lv[l].rx <= 1.21f && p.lv[l].ry <= 1.22f && !g_ef_resize_legacy) {
            const dim3 grid(ef_div_up(p.lv[l].w, RS3_TW), ef_div_up(p.lv[l].h, RS3_TH), p.nframes);
            ef_resize_tiled16_kernel<<<grid, 256, RS3_SMEM, s>>>(p, l);
            tma_prev = false;
        } else if (p.lv[l].rx <= 1.21f && p.lv[l].ry <= 1.25f) {
            const dim3 grid(ef_div_up(p.lv[l].w, RS_TW), ef_div_up(p.lv[l].h, RS_TH), p.nframes);
            ef_resize_tiled_kernel<<<grid, 256, 0, s>>>(p, l);
        } else {
            const dim3 block(64, 4);
            const dim3 grid(ef_div_up(p.lv[l].w, 256), ef_div_up(p.lv[l].h, 4), p.nframes);
            ef_resize_kernel<<<grid, block, 0, s>>>(p, l);
        }
        EF_COUNT_LAUNCH(1);
    }
}

// =================================================================================================
// score: FAST-9/16 inside the 15-px border + Harris at every corner -> dense response map
//
// 32x32-pixel tile + 4-pixel halo per CTA.  Pixels are converted once to fp16 (exact for 0..255) and kept twice in shared
// memory -- s_h0[r][c] = pixel c, s_h1[r][c] = pixel c+1 -- so that the pair (c, c+1) is an aligned 32-bit word for even AND
// odd c: the FAST ring test and the Sobel gradients then run two pixels per instruction on half2 (HSET2 compare masks,
// HADD2/HFMA2; every intermediate is an integer below 2048, hence exact).
// =================================================================================================
#define SC_HALO 4
#define SC_ROWS (EF_TILE + 2 * SC_HALO)  // 40
#define SC_PAD 4                         // halves left of tile-local column 0
#define SC_HP 48                         // halves per row = 24 words: 4 consecutive rows x 8 words hit 32 distinct banks
#define SC_HW (SC_HP / 2)
#define SC_GROWS (EF_TILE + 6)           // 38 gradient rows (tile-local rows 1..38)

__device__ __forceinline__ int ef_find_level(const EfPipe& p, int idx, int EfLevel::*start)
{
    int level = p.first_level;
    while (level + 1 < p.nlevels && idx >= p.lv[level + 1].*start) level++;
    return level;
}

// rotate both 16-bit halves of m left by S
template <int S> __device__ __forceinline__ unsigned ef_rot16x2(unsigned m)
{
    constexpr unsigned A = ((0xffffu << S) & 0xffffu) * 0x10001u;
    return ((m << S) & A) | ((m >> (16 - S)) & ~A);
}
// per 16-bit half: bit i set iff bits i, i-1, ..., i-8 (circular) of that half are all set -> non-zero half <=> 9-arc
__device__ __forceinline__ unsigned ef_arc9x2(unsigned m)
{
    unsigned r = m & ef_rot16x2<1>(m);
    r &= ef_rot16x2<2>(r);
    r &= ef_rot16x2<4>(r);
    return r & __byte_perm(m, 0, 0x2301);
}

// __launch_bounds__(256, 6): ptxas still allocates 32 registers (8 resident CTAs, shared memory allows 11) but schedules a slightly shorter
// instruction stream than with (256, 8): 7.68 vs 7.77 ms per 32 frames; (256, 5) = 38 registers 7.82, (256, 4) = 48 registers 8.19 (measured)
__global__ void __launch_bounds__(256, 6) ef_score_kernel(const __grid_constant__ EfPipe p)
{
    __shared__ __align__(16) unsigned s_h0[SC_ROWS][SC_HW];
    __shared__ __align__(16) unsigned s_h1[SC_ROWS][SC_HW];
    __shared__ __align__(16) unsigned s_grad[SC_GROWS][SC_ROWS]; // Sobel sums (dx, dy) of tile-local pixel (col, row+1) as an fp16 pair: exact integers
    __shared__ __align__(16) float s_resp[EF_TILE][EF_TILE];
    __shared__ unsigned short s_list[2][EF_TILE * EF_TILE / 2]; // corners with even / odd tile column
    __shared__ int s_n[2];

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int frame = blockIdx.y;
    const int level = ef_find_level(p, blockIdx.x, &EfLevel::tile_start);
    const EfLevel& L = p.lv[level];
    const int t = blockIdx.x - L.tile_start;
    const int tyl = ef_div_fast(t, L.tiles_x, L.tiles_x_inv), ty = tyl + L.score_ty0;
    const int x0 = (t - tyl * L.tiles_x) * EF_TILE, y0 = ty * EF_TILE;

    int pitch;
    const uint8_t* __restrict__ img = ef_level_image(p, frame, level, pitch);
    const bool aligned = ((reinterpret_cast<uintptr_t>(img) | (unsigned)pitch) & 3u) == 0;

    // ---- tile + halo -> fp16 pairs.  Pixels outside the image read as 0: no corner lies within 15 px of the border and
    //      neither the ring (3) nor the Harris window (4) reaches them.  A warp takes 3 rows of 10 words per step; the
    //      shifted copy s_h1 (pair j = pixels j+1, j+2) needs the first pixel of the next word: one shuffle.
    if (tid < 2) s_n[tid] = 0;
    {
        // CTA-uniform: tile + halo inside the image and word-aligned rows (all but the tiles along the image border) -> no per-word tests
        const bool whole = aligned && x0 >= SC_HALO && y0 >= SC_HALO && x0 + EF_TILE + SC_HALO <= L.w && y0 + EF_TILE + SC_HALO <= L.h;
        const int rsub = lane / 10, wx = lane - rsub * 10;
#pragma unroll
        for (int it = 0; it < 2; it++) {
            const int row = 3 * (warp + 8 * it) + rsub;
            const bool act = lane < 30 && row < SC_ROWS;
            const int gy = y0 - SC_HALO + row, gx = x0 - SC_HALO + 4 * wx;
            unsigned word = 0;
            if (whole) { if (act) word = *reinterpret_cast<const unsigned*>(img + (size_t)gy * pitch + gx); }
            else if (act && gy >= 0 && gy < L.h && gx >= 0 && gx < L.w) {
                const uint8_t* rp = img + (size_t)gy * pitch + gx;
                if (aligned && gx + 3 < L.w) word = *reinterpret_cast<const unsigned*>(rp);
                else {
                    word = rp[0];
                    if (gx + 1 < L.w) word |= (unsigned)rp[1] << 8;
                    if (gx + 2 < L.w) word |= (unsigned)rp[2] << 16;
                    if (gx + 3 < L.w) word |= (unsigned)rp[3] << 24;
                }
            }
            // bytes -> halves: 0x6400 | b is the half 1024 + b; subtracting 1024 is exact
            const unsigned k1024 = 0x64006400u;
            unsigned h01 = __byte_perm(word, 0x64646464u, 0x4140), h23 = __byte_perm(word, 0x64646464u, 0x4342);
            const __half2 a = __hsub2(*reinterpret_cast<__half2*>(&h01), *reinterpret_cast<const __half2*>(&k1024));
            const __half2 b = __hsub2(*reinterpret_cast<__half2*>(&h23), *reinterpret_cast<const __half2*>(&k1024));
            const unsigned ua = *reinterpret_cast<const unsigned*>(&a), ub = *reinterpret_cast<const unsigned*>(&b);
            const unsigned un = __shfl_down_sync(0xffffffffu, ua, 1); // pixels 4wx+4, 4wx+5 (garbage for wx = 9: pixel 40 is never used)
            if (act) {
                *reinterpret_cast<uint2*>(&s_h0[row][SC_PAD / 2 + 2 * wx]) = make_uint2(ua, ub);
                *reinterpret_cast<uint2*>(&s_h1[row][SC_PAD / 2 + 2 * wx]) = make_uint2(__byte_perm(ua, ub, 0x5432), __byte_perm(ub, un, 0x5432));
                if (wx == 0) s_h1[row][SC_PAD / 2 - 1] = ua << 16; // pair j = -2: (pixel -1: unused, pixel 0)
            }
        }
    }
    __syncthreads();

    // ---- FAST-9/16 (cuda_fast.cu:36-40,162-166: mask1 = darker, mask2 = brighter, >= 9 contiguous), two pixels per thread
    {
        const bool interior = x0 >= EF_HALF_PATCH && y0 >= EF_HALF_PATCH && x0 + EF_TILE <= L.w - EF_HALF_PATCH && y0 + EF_TILE <= L.h - EF_HALF_PATCH;
        const unsigned thu = (unsigned)__half_as_ushort(__int2half_rn(p.fast_threshold)) * 0x10001u;
        const __half2 th2 = *reinterpret_cast<const __half2*>(&thu);
#pragma unroll
        for (int pass = 0; pass < 2; pass++) {
            const int task = warp + 8 * pass;
            const int py = 4 * (task >> 1) + (lane >> 3), px = 16 * (task & 1) + 2 * (lane & 7);
            const unsigned* r0 = &s_h0[py + SC_HALO][(SC_PAD + SC_HALO + px) / 2];
            const unsigned* r1 = &s_h1[py + SC_HALO][(SC_PAD + SC_HALO + px) / 2];
            const unsigned cu = r0[0];
            const __half2 c2 = *reinterpret_cast<const __half2*>(&cu);
            const __half2 lo = __hsub2(c2, th2), hi = __hadd2(c2, th2);
            unsigned dark = 0, bright = 0;
            // ring pixel (dy, dx) of both pixels of the pair: even dx -> s_h0 word dx/2, odd dx -> s_h1 word (dx-1)/2
#define EF_RING(k, dy, dx) { const unsigned qu = ((dx) & 1) ? r1[(dy) * SC_HW + ((dx) - 1) / 2] : r0[(dy) * SC_HW + (dx) / 2]; \
                             const __half2 q2 = *reinterpret_cast<const __half2*>(&qu); \
                             dark |= __hlt2_mask(q2, lo) & (0x10001u << (k)); bright |= __hgt2_mask(q2, hi) & (0x10001u << (k)); }
            EF_RING(0, 3, 0) EF_RING(8, -3, 0) EF_RING(4, 0, 3) EF_RING(12, 0, -3)
            unsigned arc = 0;
            if (__any_sync(0xffffffffu, (dark | bright) != 0)) { // no 9-arc without one of the 4 compass pixels
                EF_RING(1, 3, 1) EF_RING(2, 2, 2) EF_RING(3, 1, 3) EF_RING(5, -1, 3) EF_RING(6, -2, 2) EF_RING(7, -3, 1)
                EF_RING(9, -3, -1) EF_RING(10, -2, -2) EF_RING(11, -1, -3) EF_RING(13, 1, -3) EF_RING(14, 2, -2) EF_RING(15, 3, -1)
                arc = ef_arc9x2(dark) | ef_arc9x2(bright);
            }
#undef EF_RING
            bool c0 = (arc & 0xffffu) != 0, c1 = (arc >> 16) != 0;
            if (!interior) {   // CTA-uniform: only the tiles that touch the 15-pixel border band test positions
                const int gx = x0 + px, gy = y0 + py;
                const bool rowok = gy >= EF_HALF_PATCH && gy < L.h - EF_HALF_PATCH;
                c0 = c0 && rowok && gx >= EF_HALF_PATCH && gx < L.w - EF_HALF_PATCH;
                c1 = c1 && rowok && gx + 1 >= EF_HALF_PATCH && gx + 1 < L.w - EF_HALF_PATCH;
            }
            *reinterpret_cast<float2*>(&s_resp[py][px]) = make_float2(EF_NEG_INF, EF_NEG_INF);
            const unsigned bal0 = __ballot_sync(0xffffffffu, c0), bal1 = __ballot_sync(0xffffffffu, c1);
            int base = 0;
            if (lane < 2) { const int cnt = __popc(lane ? bal1 : bal0); if (cnt) base = atomicAdd(&s_n[lane], cnt); }
            const int base0 = __shfl_sync(0xffffffffu, base, 0), base1 = __shfl_sync(0xffffffffu, base, 1);
            const unsigned lt = (1u << lane) - 1u;
            if (c0) s_list[0][base0 + __popc(bal0 & lt)] = (unsigned short)((py << 5) | px);
            if (c1) s_list[1][base1 + __popc(bal1 & lt)] = (unsigned short)((py << 5) | (px + 1));
        }
    }
    __syncthreads();
    const int n_even = s_n[0], n_odd = s_n[1];
    const int n = n_even + n_odd;

    if (n > 0) {
        // ---- Sobel sums of tile-local rows/columns 1..38 (cuda_efficient_features.cu:116-128), two pixels per thread:
        //      dx = colsum(x+1) - colsum(x-1), colsum = v(y-1) + 2 v(y) + v(y+1); dy likewise with rows.  Kept as fp16 integer
        //      pairs (|d| <= 1020, exact); the scale 1/(4*7*255) is applied where the gradient is used, exactly as
        //      SCALE * (float)d in the reference.
        const unsigned twou = 0x40004000u;
        const __half2 two = *reinterpret_cast<const __half2*>(&twou);
        // Lane mapping: a half-warp takes the 16 column pairs 0..15 of one gradient row, the two halves of a warp rows r and r + 2
        // (staged rows are 24 words apart: 48 words = 16 banks, so the 32 lanes of every load hit 32 distinct banks, and the 8-byte
        // stores of a half-warp cover one 128-byte line of s_grad); the four pairs 16..19 of every row follow in a short second pass.
        // (With the plain i / 20 mapping 43 % of this stage's shared-memory wavefronts were bank conflicts.)
        auto sobel_pair = [&](int row, int pr) {
            const unsigned* hm = &s_h0[row][SC_PAD / 2 + pr];
            const unsigned* sm1 = &s_h1[row][SC_PAD / 2 + pr];
#define H2(u) (*reinterpret_cast<const __half2*>(&(u)))
            const unsigned am = sm1[-1], bm = hm[0], cm = sm1[0];                                     // row-1: columns x-1, x, x+1
            const unsigned a0 = sm1[SC_HW - 1], c0 = sm1[SC_HW];                                     // row
            const unsigned ap = sm1[2 * SC_HW - 1], bp = hm[2 * SC_HW], cp = sm1[2 * SC_HW];          // row+1
            const __half2 right = __hfma2(two, H2(c0), __hadd2(H2(cm), H2(cp)));
            const __half2 left = __hfma2(two, H2(a0), __hadd2(H2(am), H2(ap)));
            const __half2 down = __hfma2(two, H2(bp), __hadd2(H2(ap), H2(cp)));
            const __half2 up = __hfma2(two, H2(bm), __hadd2(H2(am), H2(cm)));
#undef H2
            const __half2 dx2 = __hsub2(right, left), dy2 = __hsub2(down, up);
            const unsigned ux = *reinterpret_cast<const unsigned*>(&dx2), uy = *reinterpret_cast<const unsigned*>(&dy2);
            // (dx, dy) of the first pixel, (dx, dy) of the second
            *reinterpret_cast<uint2*>(&s_grad[row][2 * pr]) = make_uint2(__byte_perm(ux, uy, 0x5410), __byte_perm(ux, uy, 0x7632));
        };
#pragma unroll
        for (int it = 0; it < 3; it++) {
            const int j = warp + 8 * it;                        // 20 warp-steps of two rows each cover the 38 gradient rows (+ 2 unused)
            const int row = 4 * (j >> 1) + (j & 1) + 2 * (lane >> 4);
            if (j < 20 && row < SC_GROWS) sobel_pair(row, lane & 15);
        }
        if (tid < SC_GROWS * 4) sobel_pair(tid >> 2, 16 + (tid & 3));
        __syncthreads();
        // ---- Harris, raster order over the 7x7 block with the reference's contraction:
        //      g = SCALE * (float)d; sxx = fmaf(gx,gx,sxx), sxy = fmaf(gx,gy,sxy), syy = fmaf(gy,gy,syy) per tap.  (gx, gy) is one
        //      packed multiply and (sxx, syy) advance together in one packed FFMA2 (mul/fma.rn.f32x2, per-lane IEEE).  The window
        //      columns are cx+1 .. cx+7 of s_grad: for an odd cx they start on an 8-byte boundary (3 x LDS.64 + LDS.32 per row),
        //      for an even cx one tap later -- the two parity lists keep every warp on one of the two layouts.
        const float SCALE = 1.f / (4 * 7 * 255);
        const unsigned long long scale2 = ef_pack2(SCALE, SCALE);
        const int npad_even = (n_even + 31) & ~31;       // odd-list work starts on a fresh warp
        for (int i = tid; i < npad_even + n_odd; i += 256) {
            const bool odd = i >= npad_even;
            const int li = odd ? i - npad_even : i;
            if (!odd && li >= n_even) continue;
            const int pos = s_list[odd ? 1 : 0][li];
            const int cx = pos & 31, cy = pos >> 5;
            unsigned long long acc = 0ull; // (sxx, syy)
            float sxy = 0.f;
#define EF_TAP(D) { const unsigned d_ = (D); \
                    const __half2 h_ = *reinterpret_cast<const __half2*>(&d_); \
                    const unsigned long long g_ = ef_mul2(ef_pack2(__low2float(h_), __high2float(h_)), scale2); float gx_, gy_; \
                    asm("mov.b64 {%0, %1}, %2;" : "=f"(gx_), "=f"(gy_) : "l"(g_)); \
                    asm("fma.rn.f32x2 %0, %1, %1, %0;" : "+l"(acc) : "l"(g_)); \
                    sxy = fmaf(gx_, gy_, sxy); }
            if (odd) {
#pragma unroll
                for (int iy = 0; iy < 7; iy++) {
                    const unsigned* row = &s_grad[cy + iy][cx + 1];
                    const uint2 q0 = *reinterpret_cast<const uint2*>(row), q1 = *reinterpret_cast<const uint2*>(row + 2),
                                q2 = *reinterpret_cast<const uint2*>(row + 4);
                    const unsigned q3 = row[6];
                    EF_TAP(q0.x) EF_TAP(q0.y) EF_TAP(q1.x) EF_TAP(q1.y) EF_TAP(q2.x) EF_TAP(q2.y) EF_TAP(q3)
                }
            } else {
#pragma unroll
                for (int iy = 0; iy < 7; iy++) {
                    const unsigned* row = &s_grad[cy + iy][cx + 1];
                    const unsigned q0 = row[0];
                    const uint2 q1 = *reinterpret_cast<const uint2*>(row + 1), q2 = *reinterpret_cast<const uint2*>(row + 3),
                                q3 = *reinterpret_cast<const uint2*>(row + 5);
                    EF_TAP(q0) EF_TAP(q1.x) EF_TAP(q1.y) EF_TAP(q2.x) EF_TAP(q2.y) EF_TAP(q3.x) EF_TAP(q3.y)
                }
            }
#undef EF_TAP
            float sxx, syy;
            asm("mov.b64 {%0, %1}, %2;" : "=f"(sxx), "=f"(syy) : "l"(acc));
            const float p2 = sxy * sxy;
            const float det = fmaf(sxx, syy, -p2);
            const float tr = sxx + syy;
            const float tt = tr * (-0.04f);
            s_resp[cy][cx] = fmaf(tr, tt, det);
        }
        if (tid == 0 && (unsigned)(ty - L.own_ty0) < (unsigned)L.own_rows) atomicAdd(&p.counters[frame * EF_MAX_LEVELS + level].corners, n);
    }
    __syncthreads();

    // ---- block maxima for the NMS stage: largest response per b x b block, its position, tie flag
    if (p.nms_block > 0) {
        EfBlockMax* bmap = reinterpret_cast<EfBlockMax*>(ef_ws(p, frame, L.blk_off));
        if (p.nms_block == 8) {
            // 16 blocks per tile, 16 threads per block (half-warp), one float4 per thread
            const int blk = tid >> 4, sub = tid & 15;
            const int by = blk >> 2, bx = blk & 3;
            const int row = by * 8 + (sub >> 1), col = bx * 8 + (sub & 1) * 4;
            const float4 v = *reinterpret_cast<const float4*>(&s_resp[row][col]);
            const float m = fmaxf(fmaxf(v.x, v.y), fmaxf(v.z, v.w));
            float g = m;
#pragma unroll
            for (int o = 8; o > 0; o >>= 1) g = fmaxf(g, __shfl_xor_sync(0xffffffffu, g, o));
            const bool has = (m == g);
            const unsigned bal = (__ballot_sync(0xffffffffu, has) >> (lane & 16)) & 0xffffu;
            const int gbx = (x0 >> 3) + bx, gby = (y0 >> 3) + by;
            if (gbx < L.blk_w && gby < L.blk_h && has && (__ffs(bal) - 1) == sub) {
                EfBlockMax e;
                e.val = g; e.pos = 0;
                if (g > EF_NEG_INF) {
                    const int cnt = (v.x == g) + (v.y == g) + (v.z == g) + (v.w == g);
                    const int j = (v.x == g) ? 0 : (v.y == g) ? 1 : (v.z == g) ? 2 : 3;
                    const unsigned tie = (__popc(bal) > 1 || cnt > 1) ? 0x80000000u : 0u;
                    e.pos = tie | ((unsigned)(y0 + row) << 16) | (unsigned)(x0 + col + j);
                }
                bmap[(size_t)gby * L.blk_w + gbx] = e;
            }
        } else {
            const int b = p.nms_block, nb = EF_TILE / b;
            for (int blk = tid; blk < nb * nb; blk += 256) {
                const int by = blk / nb, bx = blk - by * nb;
                const int gbx = x0 / b + bx, gby = y0 / b + by;
                if (gbx >= L.blk_w || gby >= L.blk_h) continue;
                float g = EF_NEG_INF; int cnt = 0; unsigned pos = 0;
                for (int yy = 0; yy < b; yy++)
                    for (int xx = 0; xx < b; xx++) {
                        const float v = s_resp[by * b + yy][bx * b + xx];
                        if (v > g) { g = v; cnt = 1; pos = ((unsigned)(y0 + by * b + yy) << 16) | (unsigned)(x0 + bx * b + xx); }
                        else if (v == g) cnt++;
                    }
                EfBlockMax e;
                e.val = g; e.pos = (g > EF_NEG_INF) ? (pos | (cnt > 1 ? 0x80000000u : 0u)) : 0u;
                bmap[(size_t)gby * L.blk_w + gbx] = e;
            }
        }
    }

    // ---- dense response map, one float4 per thread (resp_pitch is a multiple of 32 floats)
    {
        const int row = tid >> 3, c4 = (tid & 7) * 4;
        const int gy = y0 + row;
        if (gy < L.h) {
            float* resp = reinterpret_cast<float*>(ef_ws(p, frame, L.resp_off));
            const float4 v = *reinterpret_cast<const float4*>(&s_resp[row][c4]);
            *reinterpret_cast<float4*>(resp + (size_t)gy * L.resp_pitch + x0 + c4) = v;
        }
    }
}

void ef_launch_score(const EfPipe& p, cudaStream_t s)
{
    if (p.total_tiles <= 0) return;
    ef_score_kernel<<<dim3(p.total_tiles, p.nframes), 256, 0, s>>>(p);
    EF_COUNT_LAUNCH(1);
}

// =================================================================================================
// radius NMS.   i dies iff exists j != i with resp_i <= resp_j and dx^2+dy^2 < ceil(r^2)   (cuda_efficient_features.cu:90)
//
// The score stage leaves, next to the dense response map, the maximum of every b x b pixel block
// (b = 8 for the default radius 15) with its position.  b is chosen so that a whole block lies inside
// the disc of each of its pixels (2(b-1)^2 < r^2), hence
//   * only the unique maximum of a block can survive (everything else in the block is killed by it);
//   * a candidate c is killed by block B iff some pixel of B inside c's disc has a response >= resp_c:
//     if max(B) < resp_c nothing in B can; if max(B) >= resp_c and argmax(B) is inside the disc it does;
//     only when max(B) >= resp_c sits OUTSIDE the disc are the pixels of B compared one by one (dense map).
// One CTA per strip of 4 tiles (128x32 pixels = 64 candidates for b = 8), block maxima of the strip and of the K
// blocks around it staged in shared memory with one global round trip.  Four lanes per candidate share its
// (2K+1)^2 - 1 neighbours (24 for r = 15): the 29x29-pixel disc scan of the reference becomes 6 eight-byte
// shared-memory loads per lane.  The rare per-pixel comparisons are queued and done by whole warps afterwards.
// Pure comparisons: results are exact.
// Output: one 32-bit survivor word per (tile,row) in tile-major order + per-row survivor counts.
// =================================================================================================
#define EF_NMS_RT 4          // tiles per strip
#define EF_NMS_MAX_BLK 1408  // staged block maxima: (4*32/b + 2K) * (32/b + 2K) <= 1360 for every radius in [2, 64]
#define EF_NMS_LIST 512

// pixels of block (bx0, by0) inside the disc of (cx, cy) with a response >= r?  q0/qstep: this thread's share of the b*b pixels
__device__ __forceinline__ bool ef_nms_scan_block(const float* __restrict__ resp, int resp_pitch, int w, int h, int b, int bx0, int by0,
                                                  int cx, int cy, float r, int r2, int q0, int qstep)
{
    bool hit = false;
    for (int q = q0; q < b * b; q += qstep) {
        const int gx = bx0 + (q & (b - 1)), gy = by0 + q / b;
        const int ddx = gx - cx, ddy = gy - cy;
        if (gx < w && gy < h && ddx * ddx + ddy * ddy < r2 && resp[(size_t)gy * resp_pitch + gx] >= r) hit = true;
    }
    return hit;
}

// TB, TK: block edge and block reach of the disc as compile-time constants (8, 2 for the default radius 15: no runtime divisions,
// unrolled neighbour walk); TB = 0: taken from the parameter block (any radius).
template <int TB, int TK>
__global__ void __launch_bounds__(256, 8) ef_nms_kernel(const __grid_constant__ EfPipe p)
{
    __shared__ unsigned s_mask[EF_NMS_RT * EF_TILE];
    __shared__ EfBlockMax s_blk[EF_NMS_MAX_BLK];
    __shared__ unsigned s_list[EF_NMS_LIST];
    __shared__ unsigned char s_dead[64];
    __shared__ int s_nscan;

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int frame = blockIdx.y;
    const int level = ef_find_level(p, blockIdx.x, &EfLevel::strip_start);
    const EfLevel& L = p.lv[level];
    const int s = blockIdx.x - L.strip_start;
    const int tyl = ef_div_fast(s, L.strips_x, L.strips_x_inv), ty = tyl + L.own_ty0, tx0 = (s - tyl * L.strips_x) * EF_NMS_RT;
    const int ntx = min(EF_NMS_RT, L.tiles_x - tx0);
    const int x0 = tx0 * EF_TILE, y0 = ty * EF_TILE;
    const float* __restrict__ resp = reinterpret_cast<const float*>(ef_ws(p, frame, L.resp_off));

    const int b = TB ? TB : p.nms_block;
    if (tid < EF_NMS_RT * EF_TILE) s_mask[tid] = 0;
    if (b == 0) {
        // r^2 <= 1: the disc holds only the pixel itself, every corner survives
        __syncthreads();
        for (int i = warp; i < ntx * EF_TILE; i += 8) {
            const int gy = y0 + (i & 31), gx = x0 + (i >> 5) * EF_TILE + lane;
            const bool c = gy < L.h && gx < L.w && resp[(size_t)gy * L.resp_pitch + gx] > EF_NEG_INF;
            const unsigned bal = __ballot_sync(0xffffffffu, c);
            if (lane == 0) s_mask[i] = bal;
        }
    } else {
        const EfBlockMax* __restrict__ bmap = reinterpret_cast<const EfBlockMax*>(ef_ws(p, frame, L.blk_off));
        const int lb = 31 - __clz(b), nb = EF_TILE >> lb, lrx = 2 + 5 - lb;   // strip = (4 nb) x nb blocks, 4 nb = 1 << lrx
        const int K = TB ? TK : p.nms_K, r2 = p.nms_r2;
        const int side_x = (EF_NMS_RT * nb) + 2 * K, side_y = nb + 2 * K;
        const int sbx0 = (x0 >> lb) - K, sby0 = (y0 >> lb) - K;                // block coordinates of s_blk[0]
        for (int i = tid; i < side_x * side_y; i += 256) {
            const int iy = i / side_x, gby = sby0 + iy, gbx = sbx0 + i - iy * side_x;
            EfBlockMax e; e.val = EF_NEG_INF; e.pos = 0;
            if (gbx >= 0 && gby >= 0 && gbx < L.blk_w && gby < L.blk_h) e = bmap[(size_t)gby * L.blk_w + gbx];
            s_blk[i] = e;
        }
        const int wside = 2 * K + 1, nnb = wside * wside, ncand = (EF_NMS_RT * nb) * nb;
        const int cslot = warp * 8 + (lane >> 2), part = lane & 3;
        for (int c0 = 0; c0 < ncand; c0 += 64) {
            if (tid < 64) s_dead[tid] = 0;
            if (tid == 0) s_nscan = 0;
            __syncthreads();
            const int c = c0 + cslot;
            const int lx = (c & ((1 << lrx) - 1)) + K, ly = (c >> lrx) + K;     // candidate block in s_blk coordinates
            const int cidx = ly * side_x + lx;
            EfBlockMax own; own.val = EF_NEG_INF; own.pos = 0;
            if (c < ncand) own = s_blk[cidx];
            // only the unique maximum of a block can survive (a tie inside the block kills both)
            const bool cand = own.val > EF_NEG_INF && !(own.pos & 0x80000000u);
            const float r = own.val;
            const int cx = own.pos & 0xffff, cy = (own.pos >> 16) & 0x7fff;
            const int nbase = (ly - K) * side_x + lx - K;
            bool kill = false;
            if (cand) {
                int wy = part / wside, wx = part - wy * wside;
                // branch-free body (on noise about half of the neighbouring maxima are >= r: a branch diverges anyway); with the block
                // reach known at compile time the walk is unrolled
                constexpr int NNB = TB ? (2 * TK + 1) * (2 * TK + 1) : 0;
#pragma unroll
                for (int j = 0; j < (TB ? (NNB + 3) / 4 : 0x7fffffff); j++) {
                    const int n = part + 4 * j;
                    if (n >= nnb) break;
                    const EfBlockMax e = s_blk[nbase + wy * side_x + wx];
                    const int ex = (int)(e.pos & 0xffff) - cx, ey = (int)((e.pos >> 16) & 0x7fff) - cy;
                    kill |= (n != (nnb >> 1)) & (e.val >= r) & (ex * ex + ey * ey < r2);   // the centre of the window is the block itself
                    wx += 4;
                    if (wx >= wside) { wx -= wside; wy++; }          // wside >= 5 (K >= 2): one wrap at most; wside == 3: possibly two
                    if (wside < 5 && wx >= wside) { wx -= wside; wy++; }
                }
            }
            const unsigned bal = __ballot_sync(0xffffffffu, kill);
            const bool dead = !cand || ((bal >> (lane & ~3)) & 0xfu) != 0;
            if (!dead) {
                // rare: blocks with max(B) >= r whose argmax lies OUTSIDE the disc but which intersect the disc:
                // their pixels are compared one by one (queued; done inline if the queue is full)
                int wy = part / wside, wx = part - wy * wside;
                for (int n = part; n < nnb; n += 4) {
                    if (n != (nnb >> 1)) {
                        const int nidx = nbase + wy * side_x + wx;
                        if (s_blk[nidx].val >= r) {
                            const int bx0 = (sbx0 + lx - K + wx) << lb, by0 = (sby0 + ly - K + wy) << lb;
                            const int qx = min(max(cx, bx0), bx0 + b - 1) - cx, qy = min(max(cy, by0), by0 + b - 1) - cy;
                            if (qx * qx + qy * qy < r2) {                     // nearest pixel of the block is inside the disc
                                const int slot = atomicAdd(&s_nscan, 1);
                                if (slot < EF_NMS_LIST) s_list[slot] = (unsigned)cslot | ((unsigned)nidx << 6) | ((unsigned)cidx << 17);
                                else if (ef_nms_scan_block(resp, L.resp_pitch, L.w, L.h, b, bx0, by0, cx, cy, r, r2, 0, 1)) s_dead[cslot] = 1;
                            }
                        }
                    }
                    wx += 4;
                    if (wx >= wside) { wx -= wside; wy++; }          // wside >= 5 (K >= 2): one wrap at most; wside == 3: possibly two
                    if (wside < 5 && wx >= wside) { wx -= wside; wy++; }
                }
            }
            __syncthreads();
            const int nscan = min(s_nscan, EF_NMS_LIST);
            for (int e = warp; e < nscan; e += 8) {
                const unsigned ent = s_list[e];
                const int slot = ent & 63, nidx = (ent >> 6) & 2047, oidx = ent >> 17;
                const EfBlockMax o = s_blk[oidx];
                const int ny = nidx / side_x, nx = nidx - ny * side_x;
                const bool hit = ef_nms_scan_block(resp, L.resp_pitch, L.w, L.h, b, (sbx0 + nx) << lb, (sby0 + ny) << lb,
                                                   o.pos & 0xffff, (o.pos >> 16) & 0x7fff, o.val, r2, lane, 32);
                if (__any_sync(0xffffffffu, hit) && lane == 0) s_dead[slot] = 1;
            }
            __syncthreads();
            if (!dead && part == 0 && !s_dead[cslot]) atomicOr(&s_mask[((cx >> 5) - tx0) * EF_TILE + (cy & 31)], 1u << (cx & 31));
            __syncthreads();
        }
    }
    __syncthreads();

    unsigned* mask = reinterpret_cast<unsigned*>(ef_ws(p, frame, L.mask_off));
    if (tid < ntx * EF_TILE)
        mask[((size_t)ty * L.tiles_x + tx0 + (tid >> 5)) * EF_TILE + (tid & 31)] = s_mask[tid];
    if (tid < EF_TILE) {
        int cnt = 0;
        for (int i = 0; i < ntx; i++) cnt += __popc(s_mask[i * EF_TILE + tid]);
        const int gy = y0 + tid;
        if (cnt && gy < L.h) {
            int* rowcnt = reinterpret_cast<int*>(ef_ws(p, frame, L.rowcnt_off));
            atomicAdd(&rowcnt[gy], cnt);
        }
    }
}

void ef_launch_nms(const EfPipe& p, cudaStream_t s)
{
    if (p.total_strips <= 0) return;
    if (p.nms_block == 8 && p.nms_K == 2) ef_nms_kernel<8, 2><<<dim3(p.total_strips, p.nframes), 256, 0, s>>>(p);
    else ef_nms_kernel<0, 0><<<dim3(p.total_strips, p.nframes), 256, 0, s>>>(p);
    EF_COUNT_LAUNCH(1);
}

// =================================================================================================
// compact: raster-order compaction of the survivors of every level.  One CTA per 32-row band;
// base offset of the band = prefix sum of the per-row counts above it (device-wide scan folded into
// the consumer: each band CTA reduces rowcnt[0, y0) itself, <= 4320 ints, L2 resident).
// =================================================================================================
__global__ void __launch_bounds__(1024) ef_compact_kernel(const __grid_constant__ EfPipe p)
{
    __shared__ int s_warp[32];
    __shared__ int s_rowoff[32];
    __shared__ int s_base;

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int frame = blockIdx.y;
    const int level = ef_find_level(p, blockIdx.x, &EfLevel::band_start);
    const EfLevel& L = p.lv[level];
    const int band = blockIdx.x - L.band_start + L.own_ty0;
    const int y0 = band * EF_TILE;
    const int* __restrict__ rowcnt = reinterpret_cast<const int*>(ef_ws(p, frame, L.rowcnt_off));

    int partial = 0;
    for (int i = tid; i < y0; i += 1024) partial += rowcnt[i];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) partial += __shfl_xor_sync(0xffffffffu, partial, o);
    if (lane == 0) s_warp[warp] = partial;
    __syncthreads();
    if (warp == 0) {
        int v = s_warp[lane];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
        // exclusive scan of the 32 row counts of this band
        const int y = y0 + lane;
        const int c = y < L.h ? rowcnt[y] : 0;
        int inc = c;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int u = __shfl_up_sync(0xffffffffu, inc, o);
            if (lane >= o) inc += u;
        }
        s_rowoff[lane] = inc - c;
        if (lane == 0) s_base = v;
    }
    __syncthreads();

    const int y = y0 + warp;
    if (y >= L.h) return;
    const int c = rowcnt[y];
    if (c == 0) return;

    const unsigned* __restrict__ mask = reinterpret_cast<const unsigned*>(ef_ws(p, frame, L.mask_off));
    const float* __restrict__ resp = reinterpret_cast<const float*>(ef_ws(p, frame, L.resp_off));
    EfSurvivor* surv = reinterpret_cast<EfSurvivor*>(ef_ws(p, frame, L.surv_off));

    int running = s_base + s_rowoff[warp];
    for (int tx0 = 0; tx0 < L.tiles_x; tx0 += 32) {
        const int tx = tx0 + lane;
        unsigned word = tx < L.tiles_x ? mask[((size_t)band * L.tiles_x + tx) * EF_TILE + warp] : 0u;
        const int cnt = __popc(word);
        int inc = cnt;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int u = __shfl_up_sync(0xffffffffu, inc, o);
            if (lane >= o) inc += u;
        }
        int o = running + inc - cnt;
        while (word) {
            const int bit = __ffs(word) - 1;
            word &= word - 1;
            const int x = tx * EF_TILE + bit;
            if (o < L.surv_cap) {
                EfSurvivor sv;
                sv.x = (short)x; sv.y = (short)y; sv.resp = resp[(size_t)y * L.resp_pitch + x];
                surv[o] = sv;
            }
            o++;
        }
        running += __shfl_sync(0xffffffffu, inc, 31);
    }
}

void ef_launch_compact(const EfPipe& p, cudaStream_t s)
{
    if (p.total_bands <= 0) return;
    ef_compact_kernel<<<dim3(p.total_bands, p.nframes), 1024, 0, s>>>(p);
    EF_COUNT_LAUNCH(1);
}

// =================================================================================================
// select: per-level top-quota by response (limitPoints).  Total order (response desc, y asc, x asc):
// radix-select the quota-th largest key, keep everything above it and the first (in raster order)
// of the ties; output stays in raster order.  One CTA per (level, frame).
// =================================================================================================
__device__ __forceinline__ int ef_block_excl_scan_flag(bool f, int* s_warp /*[33]*/, int& total)
{
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const unsigned bal = __ballot_sync(0xffffffffu, f);
    if (lane == 0) s_warp[warp] = __popc(bal);
    __syncthreads();
    if (warp == 0) {
        const int c = s_warp[lane];
        int inc = c;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int u = __shfl_up_sync(0xffffffffu, inc, o);
            if (lane >= o) inc += u;
        }
        s_warp[lane] = inc - c;
        if (lane == 31) s_warp[32] = inc;
    }
    __syncthreads();
    const int r = s_warp[warp] + __popc(bal & ((1u << lane) - 1u));
    total = s_warp[32];
    __syncthreads();
    return r;
}

#define EF_SEL_E 16          // survivors per thread the register-resident select holds (16 K per level: every level of a 4K frame)
__global__ void __launch_bounds__(1024, 1) ef_select_kernel(const __grid_constant__ EfPipe p)
{
    __shared__ int s_warp[33];
    __shared__ unsigned s_hist[256];
    __shared__ unsigned s_prefix;
    __shared__ int s_k;

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int level = blockIdx.x, frame = blockIdx.y;
    if (level < p.first_level) return;
    const EfLevel& L = p.lv[level];
    const int* __restrict__ rowcnt = reinterpret_cast<const int*>(ef_ws(p, frame, L.rowcnt_off));
    const EfSurvivor* __restrict__ surv = reinterpret_cast<const EfSurvivor*>(ef_ws(p, frame, L.surv_off));
    EfSelected* sel = reinterpret_cast<EfSelected*>(ef_ws(p, frame, L.sel_off));
    EfLevelCounters* ctr = &p.counters[frame * EF_MAX_LEVELS + level];

    int ntotal = 0, ncand = 0;
    if (p.select_from_counters) {
        // second pass of the band-sharded path: the list holds the merged per-band candidates (ef_band_merge_kernel left their
        // number in .overflow and the true survivor count of the level in .survivors)
        if (tid == 0) { s_warp[0] = ctr->survivors; s_warp[1] = ctr->overflow; }
        __syncthreads();
        ntotal = s_warp[0]; ncand = s_warp[1];
        __syncthreads();
    } else {
        int partial = 0;
        for (int i = tid; i < L.h; i += 1024) partial += rowcnt[i];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) partial += __shfl_xor_sync(0xffffffffu, partial, o);
        if (lane == 0) s_warp[warp] = partial;
        __syncthreads();
        for (int i = 0; i < 32; i++) ntotal += s_warp[i];
        __syncthreads();
        ncand = ntotal;
    }
    const int n = min(ncand, L.surv_cap);
    const int quota = L.quota;

    if (n <= quota) {
        for (int i = tid; i < n; i += 1024) {
            const EfSurvivor sv = surv[i];
            EfSelected o; o.x = sv.x; o.y = sv.y; o.resp = sv.resp; o.angle = 0.f; o.pad = 0;
            sel[i] = o;
        }
        if (tid == 0) { ctr->survivors = ntotal; ctr->selected = n; ctr->overflow = ncand > L.surv_cap; }
        return;
    }

    if (quota <= 0) {
        // a level whose share of nfeatures rounds to zero (nfeatures = 1): nothing is selected, and the digit search below needs k >= 1
        if (tid == 0) { ctr->survivors = ntotal; ctr->selected = 0; ctr->overflow = ncand > L.surv_cap; }
        return;
    }
    if (n <= EF_SEL_E * 1024) {
        // ---- register-resident form (every level of a 4K frame): thread t owns the contiguous run [t E, (t + 1) E) of the raster-ordered
        //      list, loaded ONCE; the four radix passes and the final ordered compaction then touch no global memory, and the two block-wide
        //      prefix sums (selected / tied elements before this thread's run) replace two scans per 1024 elements.
        const int E = (n + 1023) >> 10, base = tid * E;
        unsigned key[EF_SEL_E];
#pragma unroll
        for (int e = 0; e < EF_SEL_E; e++) {
            const int i = base + e;
            const bool in = e < E && i < n;
            key[e] = in ? ef_float_key(surv[i].resp) : 0u;       // key 0 never occurs for a real response (finite floats map above 0)
        }
        unsigned prefix = 0, pmask = 0;
        int k = quota;
        for (int shift = 24; shift >= 0; shift -= 8) {
            if (tid < 256) s_hist[tid] = 0;
            __syncthreads();
#pragma unroll
            for (int e = 0; e < EF_SEL_E; e++) {
                const bool in = key[e] != 0u && (key[e] & pmask) == prefix;
                const unsigned part = __ballot_sync(0xffffffffu, in);
                if (in) {
                    const unsigned digit = (key[e] >> shift) & 255u;
                    const unsigned peers = __match_any_sync(part, digit);
                    if (lane == __ffs(peers) - 1) atomicAdd(&s_hist[digit], (unsigned)__popc(peers));
                }
            }
            __syncthreads();
            if (tid < 256) {
                const unsigned hcnt = s_hist[tid];
                unsigned suf = hcnt;
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) {
                    const unsigned u = __shfl_down_sync(0xffffffffu, suf, o);
                    if (lane + o < 32) suf += u;
                }
                if (lane == 0) s_warp[warp] = (int)suf;
                __syncwarp();
                asm volatile("bar.sync 1, 256;");
                unsigned above = 0;
                for (int wv = warp + 1; wv < 8; wv++) above += (unsigned)s_warp[wv];
                const unsigned incl = suf + above, excl = incl - hcnt;
                if ((int)excl < k && k <= (int)incl) { s_prefix = prefix | ((unsigned)tid << shift); s_k = k - (int)excl; }
            }
            __syncthreads();
            prefix = s_prefix;
            k = s_k;
            pmask |= 0xffu << shift;
            __syncthreads();
        }
        const unsigned T = prefix;   // key of the quota-th largest; k = how many keys == T are still needed (the first k in raster order)
        int my_eq = 0, my_gt = 0;
#pragma unroll
        for (int e = 0; e < EF_SEL_E; e++) { my_eq += key[e] == T; my_gt += key[e] > T; }
        // exclusive block prefix of (tied, larger) counts: warp scan + warp totals
        int inc_eq = my_eq, inc_gt = my_gt;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int a = __shfl_up_sync(0xffffffffu, inc_eq, o), b = __shfl_up_sync(0xffffffffu, inc_gt, o);
            if (lane >= o) { inc_eq += a; inc_gt += b; }
        }
        __shared__ int s_weq[32], s_wgt[32];
        if (lane == 31) { s_weq[warp] = inc_eq; s_wgt[warp] = inc_gt; }
        __syncthreads();
        int eq_before = inc_eq - my_eq, gt_before = inc_gt - my_gt;
        for (int wv = 0; wv < warp; wv++) { eq_before += s_weq[wv]; gt_before += s_wgt[wv]; }
        // output position of an element = (larger keys before it) + (tied keys before it that are taken)
        int pos = gt_before + min(eq_before, k), eqr = eq_before;
#pragma unroll
        for (int e = 0; e < EF_SEL_E; e++) {
            const bool eq = key[e] == T, take = key[e] > T || (eq && eqr < k);
            if (take && pos < quota) {
                const EfSurvivor sv = surv[base + e];
                EfSelected o; o.x = sv.x; o.y = sv.y; o.resp = sv.resp; o.angle = 0.f; o.pad = 0;
                sel[pos] = o;
            }
            pos += take; eqr += eq;
        }
        if (tid == 1023) { ctr->survivors = ntotal; ctr->selected = min(pos, quota); ctr->overflow = ncand > L.surv_cap; }
        return;
    }

    // radix select: key of the quota-th largest response
    unsigned prefix = 0, pmask = 0;
    int k = quota;
    for (int shift = 24; shift >= 0; shift -= 8) {
        if (tid < 256) s_hist[tid] = 0;
        __syncthreads();
        // responses of one level share their leading key bytes: aggregate the histogram updates per warp (one atomic per distinct
        // digit and warp) instead of thousands of same-address shared-memory atomics
        for (int start = 0; start < n; start += 1024) {
            const int i = start + tid;
            unsigned key = 0;
            bool in = false;
            if (i < n) { key = ef_float_key(surv[i].resp); in = (key & pmask) == prefix; }
            const unsigned part = __ballot_sync(0xffffffffu, in);
            if (in) {
                const unsigned digit = (key >> shift) & 255u;
                const unsigned peers = __match_any_sync(part, digit);
                if (lane == __ffs(peers) - 1) atomicAdd(&s_hist[digit], (unsigned)__popc(peers));
            }
        }
        __syncthreads();
        // the digit b with  sum(hist[b+1..255]) < k <= sum(hist[b..255]):  suffix sums by 256 threads (8 warps) instead of one thread's loop
        if (tid < 256) {
            const unsigned hcnt = s_hist[tid];
            unsigned suf = hcnt;                                   // inclusive suffix sum inside the warp
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const unsigned u = __shfl_down_sync(0xffffffffu, suf, o);
                if (lane + o < 32) suf += u;
            }
            if (lane == 0) s_warp[warp] = (int)suf;                // warp totals (warps 0..7)
            __syncwarp();
            asm volatile("bar.sync 1, 256;");                     // the 8 participating warps only
            unsigned above = 0;                                    // counts of the warps holding larger digits
            for (int wv = warp + 1; wv < 8; wv++) above += (unsigned)s_warp[wv];
            const unsigned incl = suf + above, excl = incl - hcnt; // sum(hist[tid..255]), sum(hist[tid+1..255])
            if ((int)excl < k && k <= (int)incl) { s_prefix = prefix | ((unsigned)tid << shift); s_k = k - (int)excl; }
        }
        __syncthreads();
        prefix = s_prefix;
        k = s_k;
        pmask |= 0xffu << shift;
        __syncthreads();
    }
    const unsigned T = prefix; // key of the quota-th largest; k = how many keys == T are still needed
    int run_sel = 0, run_eq = 0;
    for (int start = 0; start < n; start += 1024) {
        const int i = start + tid;
        EfSurvivor sv; sv.x = 0; sv.y = 0; sv.resp = 0.f;
        unsigned key = 0;
        const bool in = i < n;
        if (in) { sv = surv[i]; key = ef_float_key(sv.resp); }
        const bool eq = in && key == T;
        int tot_eq, tot_sel;
        const int eq_rank = run_eq + ef_block_excl_scan_flag(eq, s_warp, tot_eq);
        const bool take = in && (key > T || (eq && eq_rank < k));
        const int pos = run_sel + ef_block_excl_scan_flag(take, s_warp, tot_sel);
        if (take && pos < quota) {
            EfSelected o; o.x = sv.x; o.y = sv.y; o.resp = sv.resp; o.angle = 0.f; o.pad = 0;
            sel[pos] = o;
        }
        run_eq += tot_eq;
        run_sel += tot_sel;
    }
    if (tid == 0) { ctr->survivors = ntotal; ctr->selected = min(run_sel, quota); ctr->overflow = ncand > L.surv_cap; }
}

void ef_launch_select(const EfPipe& p, cudaStream_t s)
{
    ef_select_kernel<<<dim3(p.nlevels, p.nframes), 1024, 0, s>>>(p);
    EF_COUNT_LAUNCH(1);
}

// =================================================================================================
// angle_pack: IC angle (warp per keypoint) + scalePoints + write the 5xN output columns at the
// level's offset (prefix of the per-level selected counts).
// =================================================================================================
#define EF_KPTS_PER_CTA 8
__constant__ int c_umax[16] = { 15, 15, 15, 15, 14, 14, 14, 13, 13, 12, 11, 10, 9, 8, 6, 3 };

__global__ void __launch_bounds__(256, 8) ef_angle_pack_kernel(const __grid_constant__ EfPipe p)
{
    __shared__ int s_m[EF_KPTS_PER_CTA][2];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int frame = blockIdx.y;
    const EfLevelCounters* ctr = &p.counters[frame * EF_MAX_LEVELS];

    if (blockIdx.x == 0 && tid == 0) {
        int total = 0;
        for (int l = p.first_level; l < p.nlevels; l++) total += ctr[l].selected;
        p.counts[frame] = min(total, p.nfeatures);
    }

    const int level = ef_find_level(p, blockIdx.x, &EfLevel::kpt_block_start);
    const EfLevel& L = p.lv[level];
    const int nsel = ctr[level].selected;
    const int i0 = (blockIdx.x - L.kpt_block_start) * EF_KPTS_PER_CTA;
    if (i0 >= nsel) return;                                   // CTA-uniform
    EfSelected* sel = reinterpret_cast<EfSelected*>(ef_ws(p, frame, L.sel_off));
    int pitch;
    const uint8_t* __restrict__ img = ef_level_image(p, frame, level, pitch);

    // IC_Angle, cuda_efficient_features.cu:141-172: warp = keypoint, lane <-> dx = lane-15; integer moments (exact)
    if (i0 + warp < nsel) {
        const EfSelected k = sel[i0 + warp];
        const uint8_t* c = img + (size_t)k.y * pitch + k.x;
        int m01 = 0, m10 = 0;
        const int dx = lane - EF_HALF_PATCH;
        if (lane < 31) {
            m10 = dx * (int)c[dx];
            const int adx = abs(dx);
            const uint8_t* pT = c + dx;
            const uint8_t* pB = c + dx;
#pragma unroll
            for (int dy = 1; dy <= EF_HALF_PATCH; dy++) {
                pT -= pitch; pB += pitch;
                if (adx <= c_umax[dy]) {
                    const int vT = *pT, vB = *pB;
                    m01 += dy * (vB - vT);
                    m10 += dx * (vB + vT);
                }
            }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            m01 += __shfl_xor_sync(0xffffffffu, m01, o);
            m10 += __shfl_xor_sync(0xffffffffu, m10, o);
        }
        if (lane == 0) { s_m[warp][0] = m01; s_m[warp][1] = m10; }
    }
    __syncthreads();
    // one thread per keypoint of the CTA: the double-precision atan2 runs once per CTA instead of once per warp
    if (tid < EF_KPTS_PER_CTA && i0 + tid < nsel) {
        const int i = i0 + tid;
        const EfSelected k = sel[i];
        int offset = 0;
        for (int l = p.first_level; l < level; l++) offset += ctr[l].selected;
        // canonical: atan2 in double, rounded once (the reference's CUDA atan2f is <= 2 ulp)
        float angle = (float)atan2((double)(float)s_m[tid][0], (double)(float)s_m[tid][1]);
        const float PI = 3.14159274f; // (float)CV_PI
        if (angle < 0) angle += 2.f * PI;
        angle = (180.f / PI) * angle;
        sel[i].angle = angle;

        const int col = offset + i;
        if (col < p.nfeatures) {
            uint8_t* kp = reinterpret_cast<uint8_t*>(p.kpts) + (size_t)frame * p.kpts_stride;
            // scalePointsKernel, cuda_efficient_features.cu:236-248 (nvcc fuses scale*x+0.5f)
            short2 pt;
            pt.x = (short)__float2int_rz(fmaf(L.scale, (float)k.x, 0.5f));
            pt.y = (short)__float2int_rz(fmaf(L.scale, (float)k.y, 0.5f));
            reinterpret_cast<short2*>(kp + (size_t)EF_LOCATION_ROW * p.kpts_pitch)[col] = pt;
            reinterpret_cast<float*>(kp + (size_t)EF_RESPONSE_ROW * p.kpts_pitch)[col] = k.resp;
            reinterpret_cast<float*>(kp + (size_t)EF_ANGLE_ROW * p.kpts_pitch)[col] = angle;
            reinterpret_cast<int*>(kp + (size_t)EF_OCTAVE_ROW * p.kpts_pitch)[col] = level;
            reinterpret_cast<float*>(kp + (size_t)EF_SIZE_ROW * p.kpts_pitch)[col] = L.scale * EF_PATCH_SIZE;
        }
    }
}

void ef_launch_angle_pack(const EfPipe& p, cudaStream_t s)
{
    if (p.total_kpt_blocks <= 0) return;
    ef_angle_pack_kernel<<<dim3(p.total_kpt_blocks, p.nframes), 256, 0, s>>>(p);
    EF_COUNT_LAUNCH(1);
}

// =================================================================================================
// blur: separable 7-tap Gaussian (sigma 2), float row pass then column pass, BORDER_REFLECT_101,
// u8 -> u8 with round-half-even.  64x64 tile, all levels in one launch.
//   load : 70 rows x 72 bytes as aligned 32-bit words (byte-wise with reflection only in tiles that touch the left/right edge)
//   rows : one thread = 4 adjacent outputs of a row: 3 word loads, 10 byte->float conversions on the FMA pipe
//          (0x4B0000bb is the float 2^23 + b; subtracting 2^23 is exact), 28 FFMA, one 16-byte store
//   cols : one thread = 4x4 outputs: 10 16-byte loads, 112 FFMA, 4 packed 32-bit stores
// =================================================================================================
#define BL_TW 64
#define BL_TH 64
#define BL_IW (EF_BLUR_BOX_W / 4) // words per staged input row: columns x0-16 .. x0+79 (words 3 .. 20 are used; the rest is there for the 16-byte rules of the TMA box)
static_assert(EF_BLUR_BOX_H == BL_TH + 6 && EF_BLUR_BOX_W >= BL_TW + 20, "blur box = tile + 3 rows / 4 + 4 columns of halo");
__device__ __forceinline__ float ef_byte_to_float(unsigned word, int i)
{
    const unsigned m = __byte_perm(word, 0x4B000000u, 0x7540 + i);
    return __uint_as_float(m) - 8388608.f;
}

// TMA: interior tiles (every source row and column of the 70 x 72 window inside the image) are staged by ONE elected thread with
// cp.async.bulk.tensor (UTMALDG) onto an mbarrier; tiles that touch an image edge keep the explicit REFLECT_101 loop.
template <bool TMA>
__global__ void __launch_bounds__(256, 6) ef_blur_kernel(const __grid_constant__ EfTmaMaps maps /* first: tensor maps must lie in the first 4 KB of the parameter space */,
                                                         const __grid_constant__ EfPipe p)
{
    __shared__ __align__(128) unsigned s_in[BL_TH + 6][BL_IW];
    __shared__ __align__(16) float s_row[BL_TH + 6][BL_TW];
    __shared__ __align__(8) unsigned long long s_mbar;

    const float t0 = __uint_as_float(0x3d8fafb1u), t1 = __uint_as_float(0x3e06387eu), t2 = __uint_as_float(0x3e434a39u),
                t3 = __uint_as_float(0x3e5d4ae0u);
    const float taps[7] = { t0, t1, t2, t3, t2, t1, t0 };
    const int tid = threadIdx.x;
    const int frame = blockIdx.y;
    const int level = ef_find_level(p, blockIdx.x, &EfLevel::blur_tile_start);
    const EfLevel& L = p.lv[level];
    const int t = blockIdx.x - L.blur_tile_start;
    const int tyl = ef_div_fast(t, L.blur_tiles_x, L.blur_tiles_x_inv), tyi = tyl + L.blur_ty0;
    const int x0 = (t - tyl * L.blur_tiles_x) * BL_TW, y0 = tyi * BL_TH;
    if (p.blur_by_slice) {
        // band-sharded frame: only the rows a descriptor window of an owned keypoint can touch (CTA-uniform exit)
        const int lo = p.slice_y[(frame * EF_MAX_LEVELS + level) * 2], hi = p.slice_y[(frame * EF_MAX_LEVELS + level) * 2 + 1];
        if (y0 + BL_TH <= lo || y0 > hi) return;
    }
    int pitch;
    const uint8_t* __restrict__ img = ef_level_image(p, frame, level, pitch);
    const bool fast = ((reinterpret_cast<uintptr_t>(img) | (unsigned)pitch) & 3u) == 0 && x0 >= 4 && x0 + BL_TW + 4 <= L.w;
    const bool tma = TMA && fast && y0 >= 3 && y0 + BL_TH + 3 <= L.h && ((maps.blur_src_ok >> level) & 1u);   // CTA-uniform

    if (tma) {
        const unsigned mbar = ef_smem_addr(&s_mbar);
        if (tid == 0) {
            ef_mbar_init(mbar, 1);
            ef_mbar_expect_tx(mbar, EF_BLUR_BOX_W * EF_BLUR_BOX_H);
            ef_tma_load_3d(ef_smem_addr(&s_in[0][0]), &maps.blur_src[level], x0 - 16, y0 - 3, frame, mbar);
        }
        __syncthreads();                    // the barrier's initialisation is visible to every waiter
        ef_mbar_wait(mbar, 0);
    } else {
#pragma unroll
        for (int it = 0; it < ((BL_TH + 6) * 18 + 255) / 256; it++) {
            const int i = tid + 256 * it;
            if (i >= (BL_TH + 6) * 18) break;
            const int ly = i / 18, wx = i - ly * 18;
            const int gy = ef_reflect101(min(y0 - 3 + ly, L.h + 2), L.h);
            const uint8_t* rp = img + (size_t)gy * pitch;
            const int gx = x0 - 4 + 4 * wx;
            unsigned word;
            if (fast) word = *reinterpret_cast<const unsigned*>(rp + gx);
            else {
                word = 0;
#pragma unroll
                for (int j = 0; j < 4; j++) word |= (unsigned)rp[ef_reflect101(min(gx + j, L.w + 2), L.w)] << (8 * j);
            }
            s_in[ly][3 + wx] = word;
        }
        __syncthreads();
    }
    // row pass: output columns 4c .. 4c+3 need input columns 4c-3 .. 4c+6 = staged bytes 4c+13 .. 4c+22 (staged byte 0 = column x0-16)
    for (int i = tid; i < (BL_TH + 6) * (BL_TW / 4); i += 256) {
        const int ly = i >> 4, c = i & 15;
        const unsigned w0 = s_in[ly][c + 3], w1 = s_in[ly][c + 4], w2 = s_in[ly][c + 5];
        // packed fp32 (fma.rn.f32x2, per-lane IEEE = the scalar chain): outputs (0,1) and (2,3) advance together; tap k multiplies the
        // input pair (k, k+1) resp. (k+2, k+3).  Pairs starting at an even / odd input are converted separately (E[m] = inputs 2m,
        // 2m+1; O[m] = inputs 2m+1, 2m+2) so that every pair is born in an aligned register pair.
        const unsigned long long m23 = ef_pack2(-8388608.f, -8388608.f);
#define BL_PAIR(wa, ia, wb, ib) ef_add2(ef_pack2(__uint_as_float(__byte_perm(wa, 0x4B000000u, 0x7540 + (ia))), \
                                                 __uint_as_float(__byte_perm(wb, 0x4B000000u, 0x7540 + (ib)))), m23)
        unsigned long long E[5], O[4];
        E[0] = BL_PAIR(w0, 1, w0, 2); E[1] = BL_PAIR(w0, 3, w1, 0); E[2] = BL_PAIR(w1, 1, w1, 2); E[3] = BL_PAIR(w1, 3, w2, 0); E[4] = BL_PAIR(w2, 1, w2, 2);
        O[0] = BL_PAIR(w0, 2, w0, 3); O[1] = BL_PAIR(w1, 0, w1, 1); O[2] = BL_PAIR(w1, 2, w1, 3); O[3] = BL_PAIR(w2, 0, w2, 1);
#undef BL_PAIR
        unsigned long long s01 = 0ull, s23 = 0ull;
#pragma unroll
        for (int k = 0; k < 7; k++) {
            const unsigned long long tk = ef_pack2(taps[k], taps[k]);
            s01 = ef_fma2((k & 1) ? O[k >> 1] : E[k >> 1], tk, s01);
            s23 = ef_fma2((k & 1) ? O[(k >> 1) + 1] : E[(k >> 1) + 1], tk, s23);
        }
        float o[4];
        asm("mov.b64 {%0, %1}, %2;" : "=f"(o[0]), "=f"(o[1]) : "l"(s01));
        asm("mov.b64 {%0, %1}, %2;" : "=f"(o[2]), "=f"(o[3]) : "l"(s23));
        *reinterpret_cast<float4*>(&s_row[ly][4 * c]) = make_float4(o[0], o[1], o[2], o[3]);
    }
    __syncthreads();
    // column pass: 4 columns x 4 rows per thread
    {
        const int c = tid & 15, rg = tid >> 4;
        const int gx = x0 + 4 * c;
        float4 r[10];
#pragma unroll
        for (int k = 0; k < 10; k++) r[k] = *reinterpret_cast<const float4*>(&s_row[4 * rg + k][4 * c]);
        uint8_t* out = ef_ws(p, frame, L.blur_off);
#pragma unroll
        for (int j = 0; j < 4; j++) {
            unsigned long long sxy = 0ull, szw = 0ull;     // (x, y) and (z, w) columns advance together
#pragma unroll
            for (int k = 0; k < 7; k++) {
                const unsigned long long tk = ef_pack2(taps[k], taps[k]);
                sxy = ef_fma2(ef_pack2(r[j + k].x, r[j + k].y), tk, sxy);
                szw = ef_fma2(ef_pack2(r[j + k].z, r[j + k].w), tk, szw);
            }
            float4 sum;
            asm("mov.b64 {%0, %1}, %2;" : "=f"(sum.x), "=f"(sum.y) : "l"(sxy));
            asm("mov.b64 {%0, %1}, %2;" : "=f"(sum.z), "=f"(sum.w) : "l"(szw));
            const int gy = y0 + 4 * rg + j;
            if (gy < L.h && gx < L.w) {
                const unsigned packed = ef_sat_u8_rne(sum.x) | (ef_sat_u8_rne(sum.y) << 8) | (ef_sat_u8_rne(sum.z) << 16) | (ef_sat_u8_rne(sum.w) << 24);
                *reinterpret_cast<unsigned*>(out + (size_t)gy * L.blur_pitch + gx) = packed; // blur_pitch is a multiple of 128
            }
        }
    }
}

void ef_launch_blur(const EfPipe& p, const EfTmaMaps* maps, cudaStream_t s)
{
    if (p.total_blur_tiles <= 0) return;
    static const bool use_tma = !(getenv("EF_BLUR_TMA") && atoi(getenv("EF_BLUR_TMA")) == 0);   // A/B switch; default: TMA staging
    static const EfTmaMaps none = {};
    if (maps && use_tma && maps->blur_src_ok) ef_blur_kernel<true><<<dim3(p.total_blur_tiles, p.nframes), 256, 0, s>>>(*maps, p);
    else ef_blur_kernel<false><<<dim3(p.total_blur_tiles, p.nframes), 256, 0, s>>>(none, p);
    EF_COUNT_LAUNCH(1);
}

// =================================================================================================
// One oversized frame over several GPUs (north_star "image tiles with halo"): every GPU holds the whole image
// and builds the whole pyramid (bandwidth-trivial), but scores / suppresses / compacts only its band of tile rows (plus the
// NMS halo for the score stage).  The per-level top-quota needs all bands: each GPU keeps its LOCAL top-quota (a superset of
// its share of the global one), the packed candidate lists are all-gathered by the caller (NCCL over NVLink, <= 8 B x
// nfeatures per GPU), and every GPU re-selects over the concatenation -- bands are ordered by rows, so the concatenation is
// in raster order and the result is the single-GPU selection bit for bit.
//
// candidate buffer of one frame: int header[3][EF_MAX_LEVELS] = {selected, survivors, corners} of the band, padded to
// EF_BAND_HDR bytes; then for level l the EfSurvivor entries at EF_BAND_HDR + 8 * sum(quota[l'] : l' < l).
// =================================================================================================
__global__ void __launch_bounds__(256) ef_band_pack_kernel(const __grid_constant__ EfPipe p, uint8_t* __restrict__ cand, unsigned long long cand_stride)
{
    const int level = blockIdx.x, frame = blockIdx.y;
    const EfLevel& L = p.lv[level];
    const EfLevelCounters c = p.counters[frame * EF_MAX_LEVELS + level];
    uint8_t* out = cand + (unsigned long long)frame * cand_stride;
    const int n = level >= p.first_level ? min(c.selected, L.quota) : 0;
    if (threadIdx.x == 0) {
        int* hdr = reinterpret_cast<int*>(out);
        hdr[level] = n;
        hdr[EF_MAX_LEVELS + level] = level >= p.first_level ? c.survivors : 0;
        hdr[2 * EF_MAX_LEVELS + level] = level >= p.first_level ? c.corners : 0;
    }
    int qoff = 0;
    for (int l = 0; l < level; l++) qoff += p.lv[l].quota;
    const EfSelected* __restrict__ sel = reinterpret_cast<const EfSelected*>(ef_ws(p, frame, L.sel_off));
    EfSurvivor* dst = reinterpret_cast<EfSurvivor*>(out + EF_BAND_HDR) + qoff;
    for (int i = threadIdx.x; i < n; i += 256) {
        const EfSelected k = sel[i];
        EfSurvivor sv; sv.x = k.x; sv.y = k.y; sv.resp = k.resp;
        dst[i] = sv;
    }
}

// all: [shard][frame][cand_stride bytes].  Concatenates the bands' candidates of every level into the level's survivor list.
__global__ void __launch_bounds__(256) ef_band_merge_kernel(const __grid_constant__ EfPipe p, const uint8_t* __restrict__ all, unsigned long long cand_stride, int nshards)
{
    const int level = blockIdx.x, frame = blockIdx.y;
    if (level < p.first_level) return;
    const EfLevel& L = p.lv[level];
    int qoff = 0;
    for (int l = 0; l < level; l++) qoff += p.lv[l].quota;
    EfSurvivor* surv = reinterpret_cast<EfSurvivor*>(ef_ws(p, frame, L.surv_off));
    int base = 0, nsurv = 0, ncorn = 0;
    for (int g = 0; g < nshards; g++) {
        const uint8_t* in = all + ((unsigned long long)g * p.nframes + frame) * cand_stride;
        const int* hdr = reinterpret_cast<const int*>(in);
        const int n = min(max(hdr[level], 0), L.quota);
        nsurv += hdr[EF_MAX_LEVELS + level];
        ncorn += hdr[2 * EF_MAX_LEVELS + level];
        const EfSurvivor* __restrict__ src = reinterpret_cast<const EfSurvivor*>(in + EF_BAND_HDR) + qoff;
        for (int i = threadIdx.x; i < n; i += 256)
            if (base + i < L.surv_cap) surv[base + i] = src[i];
        base += n;
    }
    if (threadIdx.x == 0) {
        EfLevelCounters* ctr = &p.counters[frame * EF_MAX_LEVELS + level];
        ctr->corners = ncorn; ctr->survivors = nsurv; ctr->overflow = base; ctr->selected = 0;
    }
}

// Descriptor ownership of a band-sharded frame is by OUTPUT ROW: GPU g describes rows [g C, (g + 1) C), C = ceil(nfeatures / N)
// (ef_band_desc_rows), so that an in-place all-gather of equal, contiguous row blocks assembles the matrix -- no reduction, no
// zero fill, 1/N of the bytes per GPU.  Rows are ordered by level, then raster: the slice of a GPU is, per level, a contiguous run
// of keypoints whose rows y span [first.y, last.y].  This kernel leaves that span (+- the 24-pixel reach of a descriptor window)
// per (frame, level) so that the blur only computes the 64-row tiles a window of an owned keypoint can touch.
__global__ void __launch_bounds__(32) ef_band_slice_ranges_kernel(const __grid_constant__ EfPipe p, int* __restrict__ slice_y)
{
    const int frame = blockIdx.x, level = threadIdx.x;
    if (level >= p.nlevels) return;
    const EfLevelCounters* ctr = &p.counters[frame * EF_MAX_LEVELS];
    int lo = 1 << 30, hi = -(1 << 30);                       // empty: no tile intersects
    if (level >= p.first_level) {
        int offset = 0;
        for (int l = p.first_level; l < level; l++) offset += ctr[l].selected;
        const int nsel = min(ctr[level].selected, max(p.nfeatures - offset, 0));
        const int a = max(p.desc_row0 - offset, 0), b = min(p.desc_row0 + p.desc_rows - offset, nsel);
        if (a < b) {
            const EfSelected* sel = reinterpret_cast<const EfSelected*>(ef_ws(p, frame, p.lv[level].sel_off));
            lo = sel[a].y - 24; hi = sel[b - 1].y + 24;      // the selected list is in raster order
        }
    }
    slice_y[(frame * EF_MAX_LEVELS + level) * 2] = lo;
    slice_y[(frame * EF_MAX_LEVELS + level) * 2 + 1] = hi;
}

void ef_launch_band_pack(const EfPipe& p, uint8_t* cand, unsigned long long cand_stride, cudaStream_t s)
{
    ef_band_pack_kernel<<<dim3(p.nlevels, p.nframes), 256, 0, s>>>(p, cand, cand_stride);
    EF_COUNT_LAUNCH(1);
}
void ef_launch_band_merge(const EfPipe& p, const uint8_t* all, unsigned long long cand_stride, int nshards, cudaStream_t s)
{
    ef_band_merge_kernel<<<dim3(p.nlevels, p.nframes), 256, 0, s>>>(p, all, cand_stride, nshards);
    EF_COUNT_LAUNCH(1);
}
void ef_launch_band_slice_ranges(const EfPipe& p, int* slice_y, cudaStream_t s)
{
    ef_band_slice_ranges_kernel<<<p.nframes, 32, 0, s>>>(p, slice_y);
    EF_COUNT_LAUNCH(1);
}
