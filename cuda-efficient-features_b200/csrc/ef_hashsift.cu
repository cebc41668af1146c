// ef_hashsift.cu -- HashSIFT feature kernel (sm_100a): rectified 32x32 patch -> deterministic 4x4x8 gradient
// histogram -> 128 u8, bit-exact against modules/efficient_features/src/hash_sift.cpp:68-138,150-198,200-331
// (replaces computePatchSIFTKernel, src/cuda_hash_sift.cu:380-412, whose shared-memory float atomics are not).
//
// Compiled with -fmad=false: the CPU reference is a generic x86-64 build without FMA, so every a*b+c below
// stays two roundings.
//
// One HALF-WARP per keypoint (2 keypoints per warp, 4 per 64-thread CTA, 7.9 KB of shared memory each -- seven CTAs per SM):
//   1. (detectAndCompute path) the 48x48-pixel window that bounds the rotated patch is staged in shared memory
//      with aligned 32-bit loads; the 32x32 bilinear samples (hash_sift.cpp:88-106) then read bytes from it.
//   2. per gradient pixel (30x30): dx,dy in [-255,255] index ONE 8-byte table entry holding sqrtf(dx^2+dy^2),
//      the orientation-bin fraction and the bin number (finite-domain tables filled by the host libm, the libm
//      the CPU reference links); magnitude = expf-table * sqrt.  The 3-bit bin rides in the sign bit of the
//      magnitude and the two unused top bits of the fraction (< 1), so a pixel record is two floats.
//   3. trilinear histogram (hash_sift.cpp:233-290): lane c of the half-warp owns histogram cell c (4x4 cells) and
//      walks the pixels that feed it in raster order, so every accumulator receives exactly the CPU's sequence
//      of additions.  The cell scale is exactly 1/8, hence the row weight of patch row y is ((y-3) mod 8)/8 and
//      the rows feeding histogram row rb are 8(rb-2)+3 .. 8(rb-1)+10: two runs of 8 rows (weight w, then 1-w) -- the
//      same for columns -- which makes the loop nest static.  Accumulators live in shared memory as
//      hist[bin][lane]: bank = lane, conflict-free for any bin; pixel records are skewed (y*30 + x + 2*(y>>3)) so
//      that the 16 cells of one keypoint hit 16 distinct even banks and the second keypoint of the warp (array
//      base an odd number of words further) the odd ones.
//   4. fold bins 8 -> 0, L2-normalise (sequential sum), clip 0.2, renormalise, x512 -> u8 (hash_sift.cpp:293-330).
#include "ef_common.cuh"
#include "ef_libm_f32.cuh"

#include <cfloat>
#include <cstdlib>

#define EF_SIFT_WARPS 2                 // per CTA: 4 keypoints
#define EF_SIFT_KP_PER_CTA (2 * EF_SIFT_WARPS)
#define EF_SIFT_REC 916                 // floats per record array (30x30 skewed needs 906)
#define EF_SIFT_ZERO 908                // spare record slot (skewed 30x30 indices end at 905) holding an all-zero record
#define EF_SIFT_BLK 1872                // floats per keypoint block (16-byte multiple): magnitudes, fractions, slack
#define EF_SIFT_WIN_ROWS 46             // staged window rows k-22 .. k+23: every sample of a size-31 patch lies in [k-22, k+22]
#define EF_SIFT_WIN 48                  // staged window columns start at (k - 24) & ~15
#define EF_SIFT_WIN_PITCH 136           // bytes per staged row: 64 pixels (48 + up to 15 alignment bytes), TWO bytes each -- entry x holds
                                        // (pixel x, pixel x+1), so a bilinear sample is two 16-bit loads; 34-word pitch spreads the banks
#define EF_SIFT_PATCH_OFF 6464          // byte offset of the 32x32 u8 patch inside the keypoint block
#define EF_SIFT_GROWS 5                 // gradient rows per lane and step (10 table gathers in flight; 10 rows: the same time, measured)

// Shared memory of one warp = 2 keypoints: 16 128 bytes, so that SEVEN 2-warp CTAs fit one SM (2 x 16128 + 1024 reserved = 33280 = 130 x 256-byte allocation units, 7 x 33280 <= 233472;
// 18.9 KB per warp gave six).  One block per keypoint is used three times over:
//   staging   bytes [0, 46 x 136) the 46 x 64 window, while the sampler writes the patch at [PATCH_OFF, PATCH_OFF + 1024)
//   gradients magnitude[i] at float k + i (sign bit = bin bit 2), fraction[i] at float REC + k + i (bits 31:30 = bin bits 1:0), written
//             in steps of EF_SIFT_GROWS pixel rows while the patch is still being read: the fraction records of rows >= 23 land ON the patch, at
//             patch byte 4 (i + k + 916) - PATCH_OFF, always in rows that no later step reads (tests/test_layout_hashsift.py replays the
//             schedule); inside a step every lane reads before any lane writes (__syncwarp)
//   histogram records read only; afterwards the first 128 floats hold the descriptor being normalised
// The "+ k" puts the records of the second keypoint on the other bank parity (block stride = 1872 words, even).
struct EfSiftWarpSmem {
    float hist[9 * 32];                 // [bin 0..8][lane]
    __align__(16) float blk[2][EF_SIFT_BLK];
};
static_assert(sizeof(EfSiftWarpSmem) == 16128 && 7 * (EF_SIFT_WARPS * sizeof(EfSiftWarpSmem) + 1024) <= 233472 && (EF_SIFT_WARPS * sizeof(EfSiftWarpSmem) + 1024) % 256 == 0, "7 CTAs of 2 warps per SM");
static_assert(EF_SIFT_WIN_ROWS * EF_SIFT_WIN_PITCH <= EF_SIFT_PATCH_OFF, "the staged window must end below the patch");
static_assert(EF_SIFT_PATCH_OFF % 16 == 0 && EF_SIFT_PATCH_OFF + 1024 <= EF_SIFT_BLK * 4, "patch inside the block");
static_assert(4 * (2 * EF_SIFT_REC + 1) <= EF_SIFT_BLK * 4 && EF_SIFT_BLK % 4 == 0, "both record arrays inside the block");

// sum of the four unsigned bytes of a times the four signed bytes of coeff
__device__ __forceinline__ int ef_dot4_u8s8(unsigned a, unsigned coeff)
{
    int r;
    asm("dp4a.u32.s32 %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(coeff), "r"(0));
    return r;
}

// normalize(), hash_sift.cpp:150-160.  Lane hl keeps the eight bins of its cell (descriptor elements 8 hl .. 8 hl + 7) in registers; the sum of
// squares is sequential (one chain of 128 additions, computed redundantly by every lane of the half-warp) over the squares, which are
// formed once, eight per lane, into the scratch sq[128].
__device__ __forceinline__ void ef_sift_normalize(float (&v)[8], float* sq, int hl)
{
    float4* sq4 = reinterpret_cast<float4*>(sq);
    sq4[2 * hl] = make_float4(v[0] * v[0], v[1] * v[1], v[2] * v[2], v[3] * v[3]);
    sq4[2 * hl + 1] = make_float4(v[4] * v[4], v[5] * v[5], v[6] * v[6], v[7] * v[7]);
    __syncwarp();
    float sum = 0.f;
#pragma unroll 8
    for (int i = 0; i < 32; i++) {
        const float4 q = sq4[i];
        sum += q.x; sum += q.y; sum += q.z; sum += q.w;
    }
    __syncwarp();                       // every lane has read the squares before the next call overwrites them
    const float nrm = fmaxf(sqrtf(sum), FLT_EPSILON);
    const float scale = 1.f / nrm;
#pragma unroll
    for (int j = 0; j < 8; j++) v[j] *= scale;
}

// All 32 lanes call this; lanes 0-15 work on keypoint slot 0 of the warp, lanes 16-31 on slot 1.
// STAGED: integer keypoint, size 31, scale 1 (detectAndCompute path): window staging; image base and pitch 16-byte aligned.
template <bool STAGED, int V = 0>
__device__ __forceinline__ void ef_hashsift_one(const uint8_t* __restrict__ img, int w, int h, int pitch,
                                                float kx, float ky, float size, float angle, float croppingScale,
                                                const EfHashSiftTables& t, EfSiftWarpSmem& sm, uint8_t* out128, bool store)
{
    const int lane = threadIdx.x & 31, hl = lane & 15, k = lane >> 4;
    float* __restrict__ blk = sm.blk[k];
    uint8_t* __restrict__ patch = reinterpret_cast<uint8_t*>(blk) + EF_SIFT_PATCH_OFF;
    // ---- rectifyPatch + warpAffineLinear (hash_sift.cpp:68-138)
    {
        const float PI_1_0F = 3.14159274f;
        const float s = croppingScale * size / (0.5f * (float)(32 + 32));
        const float theta = PI_1_0F * angle / 180;
        // cosf/sinf: the host libm's algorithm, bit for bit (ef_libm_f32.cuh)
        const float cost = s * (angle >= 0 ? ef_libm::cosf_glibc(theta) : 1.f);
        const float sint = s * (angle >= 0 ? ef_libm::sinf_glibc(theta) : 0.f);
        const float M00 = +cost, M01 = -sint, M02 = (-cost + sint) * 32.f / 2.f + kx;
        const float M10 = +sint, M11 = +cost, M12 = (-sint - cost) * 32.f / 2.f + ky;
        const uint8_t* __restrict__ base = img;
        int bpitch = pitch, ox = 0, oy = 0;
        if (STAGED) {
            // 46 rows x 64 pixels (16-byte aligned start <= wx0), four 16-byte loads per row, all 16 lanes busy; stored as
            // overlapping pixel pairs (x, x+1): the first pixel of the next chunk comes from the neighbouring lane
            const int wx0 = (int)kx - EF_SIFT_WIN / 2, wy0 = (int)ky - (EF_SIFT_WIN_ROWS / 2 - 1);
            const int gx0 = wx0 & ~15;
            uint8_t* __restrict__ win = reinterpret_cast<uint8_t*>(blk);
            const int gxc = gx0 + 16 * (hl & 3);
            const bool colok = gxc >= 0 && gxc + 15 < pitch;
#pragma unroll
            for (int it = 0; it < (EF_SIFT_WIN_ROWS + 3) / 4; it++) {
                const int row = 4 * it + (hl >> 2), gy = wy0 + row;
                uint4 v = make_uint4(0, 0, 0, 0);
                if (colok && gy >= 0 && gy < h && row < EF_SIFT_WIN_ROWS) v = __ldg(reinterpret_cast<const uint4*>(img + (size_t)gy * pitch + gxc));
                const unsigned nx = __shfl_down_sync(0xffffffffu, v.x, 1); // chunk 3: pixel 64 is never sampled
                if (row < EF_SIFT_WIN_ROWS) {
                    uint2* dst = reinterpret_cast<uint2*>(win + row * EF_SIFT_WIN_PITCH + 32 * (hl & 3));
                    dst[0] = make_uint2(__byte_perm(v.x, 0, 0x2110), __byte_perm(v.x, v.y, 0x4332));
                    dst[1] = make_uint2(__byte_perm(v.y, 0, 0x2110), __byte_perm(v.y, v.z, 0x4332));
                    dst[2] = make_uint2(__byte_perm(v.z, 0, 0x2110), __byte_perm(v.z, v.w, 0x4332));
                    dst[3] = make_uint2(__byte_perm(v.w, 0, 0x2110), __byte_perm(v.w, nx, 0x4332));
                }
            }
            __syncwarp();
            base = win; bpitch = EF_SIFT_WIN_PITCH; ox = gx0; oy = wy0;
        }
        const float cx0 = M00 * (float)hl, cx1 = M00 * (float)(hl + 16);
        const float cy0 = M10 * (float)hl, cy1 = M10 * (float)(hl + 16);
        // STAGED: when the whole 48 x 48 window of both keypoints of the warp lies inside the image, every sample does (all of them fall
        // in [k - 22, k + 22]) and the four bounds tests per sample are dropped (warp-uniform choice; same arithmetic)
        bool inside = false;
        if (STAGED) {
            const int ikx = (int)kx, iky = (int)ky;
            inside = __all_sync(0xffffffffu, ikx >= EF_SIFT_WIN / 2 && ikx + EF_SIFT_WIN / 2 < w && iky >= EF_SIFT_WIN / 2 && iky + EF_SIFT_WIN / 2 < h);
        }
#define EF_SIFT_SAMPLE_ROWS(CHECK) \
        _Pragma("unroll 4") \
        for (int y = 0; y < 32; y++) { \
            const float ru = M01 * (float)y, rv = M11 * (float)y; \
            _Pragma("unroll") \
            for (int xx = 0; xx < 2; xx++) { \
                const float u = ((xx ? cx1 : cx0) + ru) + M02; \
                const float v = ((xx ? cy1 : cy0) + rv) + M12; \
                uint8_t dstVal = 0; \
                const int ui = (int)floorf(u); \
                const int vi = (int)floorf(v); \
                if (!(CHECK) || (ui >= 0 && ui + 1 < w && vi >= 0 && vi + 1 < h)) { \
                    const float du = u - (float)ui; \
                    const float dv = v - (float)vi; \
                    float q00, q01, q10, q11; \
                    if (STAGED) { \
                        const unsigned short* __restrict__ q = reinterpret_cast<const unsigned short*>(base + (vi - oy) * EF_SIFT_WIN_PITCH) + (ui - ox); \
                        const unsigned t0 = q[0], t1 = q[EF_SIFT_WIN_PITCH / 2]; \
                        q00 = (float)(t0 & 0xffu); q01 = (float)(t0 >> 8); q10 = (float)(t1 & 0xffu); q11 = (float)(t1 >> 8); \
                    } else { \
                        const uint8_t* __restrict__ q = base + (vi - oy) * bpitch + (ui - ox); \
                        q00 = (float)q[0]; q01 = (float)q[1]; q10 = (float)q[bpitch]; q11 = (float)q[bpitch + 1]; \
                    } \
                    const float tmp0 = (1 - du) * q00 + du * q01; \
                    const float tmp1 = (1 - du) * q10 + du * q11; \
                    const float tmp2 = (1 - dv) * tmp0 + dv * tmp1; \
                    dstVal = (uint8_t)min(__float2int_rz(tmp2 + 0.5f), 255); \
                } \
                patch[y * 32 + hl + 16 * xx] = dstVal; \
            } \
        }
        if (inside && STAGED && (V & 4)) {
            // both samples of a lane and row (columns hl and hl + 16) advance together in packed fp32: the same operations in the
            // same order as the scalar form above (separate multiplies and adds: the CPU reference is built without FMA)
            const unsigned long long cx2 = ef_pack2(cx0, cx1), cy2 = ef_pack2(cy0, cy1);
            const unsigned long long m02 = ef_pack2(M02, M02), m12 = ef_pack2(M12, M12), one2 = ef_pack2(1.f, 1.f), half2 = ef_pack2(0.5f, 0.5f);
            const float nz = __uint_as_float(0x80000000u | (blockDim.z - 1u));   // -0.0f at run time (blockDim.z == 1), opaque to the compiler
            const unsigned long long nz2 = ef_pack2(nz, nz);
            // No conversion instruction in the loop (F2I / I2F run on the quarter-rate XU pipe, and ten of them per row made this phase
            // XU-bound): floor(x) = (x + 1.5 * 2^23, rounded toward -infinity) - 1.5 * 2^23 -- exact for |x| < 2^22, the integer sits in
            // the low mantissa bits -- and byte -> float by planting the byte in the mantissa of 2^23 (PRMT) and subtracting 2^23.
            const unsigned long long fl2 = ef_pack2(12582912.f, 12582912.f), m23 = ef_pack2(8388608.f, 8388608.f);
            const int cu = 0x4B400000 + ox, cv = 0x4B400000 + oy;
#pragma unroll 4
            for (int y = 0; y < 32; y++) {
                const float ru = M01 * (float)y, rv = M11 * (float)y;
                const unsigned long long u2 = ef_add2(ef_add2(cx2, ef_pack2(ru, ru)), m02);
                const unsigned long long v2 = ef_add2(ef_add2(cy2, ef_pack2(rv, rv)), m12);
                const unsigned long long tu2 = ef_add2_rm(u2, fl2), tv2 = ef_add2_rm(v2, fl2);
                const unsigned long long du2 = ef_sub2(u2, ef_sub2(tu2, fl2)), dv2 = ef_sub2(v2, ef_sub2(tv2, fl2));
                const unsigned long long omdu2 = ef_sub2(one2, du2), omdv2 = ef_sub2(one2, dv2);
                float tua, tub, tva, tvb;
                ef_unpack2(tu2, tua, tub); ef_unpack2(tv2, tva, tvb);
                const unsigned short* __restrict__ qa = reinterpret_cast<const unsigned short*>(base + (__float_as_int(tva) - cv) * EF_SIFT_WIN_PITCH) + (__float_as_int(tua) - cu);
                const unsigned short* __restrict__ qb = reinterpret_cast<const unsigned short*>(base + (__float_as_int(tvb) - cv) * EF_SIFT_WIN_PITCH) + (__float_as_int(tub) - cu);
                const unsigned a0 = qa[0], a1 = qa[EF_SIFT_WIN_PITCH / 2], b0 = qb[0], b1 = qb[EF_SIFT_WIN_PITCH / 2];
#define EF_B2F(w, sel) __uint_as_float(__byte_perm(w, 0x4B000000u, sel))
                const unsigned long long q00 = ef_sub2(ef_pack2(EF_B2F(a0, 0x7440), EF_B2F(b0, 0x7440)), m23), q01 = ef_sub2(ef_pack2(EF_B2F(a0, 0x7441), EF_B2F(b0, 0x7441)), m23);
                const unsigned long long q10 = ef_sub2(ef_pack2(EF_B2F(a1, 0x7440), EF_B2F(b1, 0x7440)), m23), q11 = ef_sub2(ef_pack2(EF_B2F(a1, 0x7441), EF_B2F(b1, 0x7441)), m23);
#undef EF_B2F
                // ptxas contracts mul.rn.f32x2 + add.rn.f32x2 into FFMA2 even under -fmad=false (seen in SASS); the reference rounds every
                // product.  So each product is an FMA with a -0.0 addend that the compiler cannot see (x*y + -0.0 == round(x*y) bit for
                // bit), and an FMA followed by an add is not fusable.
                const unsigned long long tmp0 = ef_add2(ef_fma2(omdu2, q00, nz2), ef_fma2(du2, q01, nz2));
                const unsigned long long tmp1 = ef_add2(ef_fma2(omdu2, q10, nz2), ef_fma2(du2, q11, nz2));
                const unsigned long long tmp2 = ef_add2(ef_fma2(omdv2, tmp0, nz2), ef_fma2(dv2, tmp1, nz2));
                // (uint8_t)min((int)(tmp2 + 0.5f), 255): tmp2 is a convex combination of bytes (<= 255 + rounding), so the truncation is the
                // low byte of (tmp2 + 0.5f) + 2^23 rounded toward -infinity
                float ra, rb;
                ef_unpack2(ef_add2_rm(ef_add2(tmp2, half2), m23), ra, rb);
                patch[y * 32 + hl] = (uint8_t)__float_as_uint(ra);
                patch[y * 32 + hl + 16] = (uint8_t)__float_as_uint(rb);
            }
        }
        else if (inside) { EF_SIFT_SAMPLE_ROWS(false) } else { EF_SIFT_SAMPLE_ROWS(true) }
#undef EF_SIFT_SAMPLE_ROWS
    }
    __syncwarp();
    // ---- per-pixel magnitude / orientation (hash_sift.cpp:247-260) through the finite-domain tables (ef_api.cu).
    //      Lane hl < 15 owns the gradient columns 2 hl and 2 hl + 1 and walks the 30 rows; the four patch bytes [2 hl, 2 hl + 3] of a row serve
    //      the row above (as "below"), its own row's left/right neighbours and the row below (as "above"), so every patch row is loaded once
    //      and kept in a rolling register window; the four differences are byte dot products (IDP.4A), every index is a compile-time
    //      offset from a per-lane base.  EF_SIFT_GROWS rows per step: patch loads, then 2 x GROWS table gathers in flight, then the records.
    {
        float* __restrict__ magp = blk + k;
        float* __restrict__ ofp = blk + EF_SIFT_REC + k;
        constexpr int R = EF_SIFT_GROWS;
        static_assert(30 % R == 0, "whole steps");
        const int xl = min(2 * hl, 28);                         // lane 15 shadows lane 14 (its stores are skipped)
        const bool act = hl < 15;
        const uint8_t* __restrict__ prow = patch + xl;
        const float2* __restrict__ gt = t.grad_table + (255 * 511 + 255);
        const float2* __restrict__ et = reinterpret_cast<const float2*>(t.exp_table + xl);
        auto row4 = [&](int r) {
            const unsigned lo = *reinterpret_cast<const unsigned short*>(prow + 32 * r), hi = *reinterpret_cast<const unsigned short*>(prow + 32 * r + 2);
            return __byte_perm(lo, hi, 0x5410);
        };
        unsigned rw[R + 2];
        rw[0] = row4(0); rw[1] = row4(1);
#pragma unroll
        for (int y0 = 0; y0 < 30; y0 += R) {
#pragma unroll
            for (int j = 0; j < R; j++) rw[j + 2] = row4(y0 + j + 2);
            float2 e[2 * R], ew[R];
#pragma unroll
            for (int j = 0; j < R; j++) {
                const unsigned above = rw[j], centre = rw[j + 1], below = rw[j + 2];
                const int dxa = ef_dot4_u8s8(centre, 0x000100FFu), dxb = ef_dot4_u8s8(centre, 0x0100FF00u);           // c[1] - c[-1]
                const unsigned ab = __byte_perm(above, below, 0x6521);                                         // (above[1], above[2], below[1], below[2])
                const int dya = ef_dot4_u8s8(ab, 0x00FF0001u), dyb = ef_dot4_u8s8(ab, 0xFF000100u);               // c[-32] - c[32]
                e[2 * j] = __ldg(gt + (dya * 511 + dxa));
                e[2 * j + 1] = __ldg(gt + (dyb * 511 + dxb));
                ew[j] = __ldg(et + 15 * (y0 + j));
            }
            rw[0] = rw[R]; rw[1] = rw[R + 1];
            __syncwarp();                                       // late fraction records overwrite patch rows (read long before: see the test)
            if (act) {
#pragma unroll
                for (int j = 0; j < R; j++) {
                    const int y = y0 + j, rix = y * 30 + 2 * (y >> 3) + xl;
                    magp[rix] = ew[j].x * e[2 * j].x; ofp[rix] = e[2 * j].y;
                    magp[rix + 1] = ew[j].y * e[2 * j + 1].x; ofp[rix + 1] = e[2 * j + 1].y;
                }
            }
        }
    }
    __syncwarp();
    for (int b = 0; b < 9; b++) sm.hist[b * 32 + lane] = 0.f;
    // all-zero record (magnitude +0, fraction 0, bin 0) in a spare slot: what the out-of-patch visits of the border cells read
    if (hl == 0) { blk[k + EF_SIFT_ZERO] = 0.f; blk[EF_SIFT_REC + k + EF_SIFT_ZERO] = 0.f; }
    __syncwarp();
    // ---- trilinear histogram (hash_sift.cpp:233-290); cell (rb, cb) in 1..4
    {
        const int rb = (hl >> 2) + 1, cb = (hl & 3) + 1;
        const float* __restrict__ mp = blk + k;
        const unsigned* __restrict__ op = reinterpret_cast<const unsigned*>(blk + EF_SIFT_REC + k);
        float* __restrict__ hc = sm.hist + lane;
        const int xb = 8 * (cb - 2) + 3;
        const float nzh = __uint_as_float(0x80000000u | (blockDim.z - 1u));   // -0.0f at run time, opaque to the compiler
        const unsigned long long nzh2 = ef_pack2(nzh, nzh);
        // Out-of-patch visits (rows/columns outside 0..29 for the border cells) are not branched around: they read the all-zero
        // record and add +0.0f, which leaves every (non-negative) accumulator unchanged -- the warp executes the iteration anyway.
        // Visits whose weight is exactly zero for EVERY cell are skipped (they too would only add +0.0f): the first row of the
        // first row segment (row weight 0/8) and the first column of the first column segment (column weight 0/8) -- 31 of the
        // 256 visits of a cell.
        // Per patch row: first the 15 or 16 shares (loads + arithmetic, independent -> pipelined), then the ordered read-modify-writes.
#pragma unroll
        for (int rseg = 0; rseg < 2; rseg++) {
            for (int iy = (rseg == 0 && (V & 2)) ? 1 : 0; iy < 8; iy++) {
                const int y = 8 * (rb - 2 + rseg) + 3 + iy;
                const bool rowok = (unsigned)y < 30u;
                const float rf = 0.125f * (float)iy;
                const int rowidx = y * 30 + 2 * (y >> 3) + xb;
                float vo0[16], vo1[16];
                unsigned hoff[16];
                if (V & 8) {
                    // packed fp32 shares: visits (xo, xo + 1) of one column segment advance together.  The sign bit of the magnitude record
                    // (bin bit 2) is carried through instead of cleared -- every operation below is sign-symmetric -- and dropped by
                    // the |.| of the final additions.  Products are FMAs with the run-time -0.0 addend (see the sampling loop).
                    float mgv[16];
                    unsigned ofv[16];
#pragma unroll
                    for (int xo = 1; xo < 16; xo++) {
                        // out-of-patch visit: an all-zero record without touching memory (predicated loads at a fixed offset from the row base)
                        const bool ok = rowok && (unsigned)(xb + xo) < 30u;
                        float mg0 = 0.f;
                        unsigned ob = 0u;
                        if (ok) { mg0 = mp[rowidx + xo]; ob = op[rowidx + xo]; }
                        mgv[xo] = mg0;
                        hoff[xo] = ((ob >> 30) | ((__float_as_uint(mgv[xo]) >> 31) << 2)) * 32u;
                        ofv[xo] = ob & 0x3fffffffu;
                    }
                    {   // xo = 1 (column weight 1/8 of the first segment) has no partner
                        const float v1 = rf * mgv[1];
                        const float vr = rseg == 0 ? v1 : mgv[1] - v1;
                        const float vc = 0.125f * vr;
                        vo1[1] = __uint_as_float(ofv[1]) * vc;
                        vo0[1] = vc - vo1[1];
                    }
                    const unsigned long long rf2 = ef_pack2(rf, rf);
#pragma unroll
                    for (int xp = 2; xp < 16; xp += 2) {
                        const int cseg = xp >> 3, ix = xp & 7;
                        const unsigned long long mg2 = ef_pack2(mgv[xp], mgv[xp + 1]);
                        const unsigned long long of2 = ef_pack2(__uint_as_float(ofv[xp]), __uint_as_float(ofv[xp + 1]));
                        const unsigned long long v1 = ef_fma2(rf2, mg2, nzh2);
                        const unsigned long long vr = rseg == 0 ? v1 : ef_sub2(mg2, v1);
                        const unsigned long long c1 = ef_fma2(ef_pack2(0.125f * (float)ix, 0.125f * (float)(ix + 1)), vr, nzh2);
                        const unsigned long long vc = cseg == 0 ? c1 : ef_sub2(vr, c1);
                        const unsigned long long o1 = ef_fma2(of2, vc, nzh2);
                        const unsigned long long o0 = ef_sub2(vc, o1);
                        ef_unpack2(o1, vo1[xp], vo1[xp + 1]);
                        ef_unpack2(o0, vo0[xp], vo0[xp + 1]);
                    }
#pragma unroll
                    for (int xo = 1; xo < 16; xo++) {
                        float* h0 = hc + hoff[xo];
                        const float a0 = h0[0], a1 = h0[32];
                        h0[0] = a0 + fabsf(vo0[xo]);
                        h0[32] = a1 + fabsf(vo1[xo]);
                    }
                    continue;
                }
#pragma unroll
                for (int xo = (V & 2) ? 1 : 0; xo < 16; xo++) {
                    const int cseg = xo >> 3, ix = xo & 7;
                    const bool ok = rowok && (unsigned)(xb + xo) < 30u;
                    const int idx = ok ? rowidx + xo : ((V & 1) ? EF_SIFT_ZERO : 0);
                    const float mg = mp[idx];
                    const unsigned ob = op[idx];
                    const float mag = (V & 1) ? fabsf(mg) : (ok ? fabsf(mg) : 0.f);
                    hoff[xo] = ((ob >> 30) | ((__float_as_uint(mg) >> 31) << 2)) * 32u;
                    const float of = __uint_as_float(ob & 0x3fffffffu);
                    // distribute(): v1 = w*v; v0 = v - v1   (hash_sift.cpp:193-198)
                    const float v1 = rf * mag;
                    const float vr = rseg == 0 ? v1 : mag - v1;
                    const float c1 = (0.125f * (float)ix) * vr;
                    const float vc = cseg == 0 ? c1 : vr - c1;
                    vo1[xo] = of * vc;
                    vo0[xo] = vc - vo1[xo];
                }
#pragma unroll
                for (int xo = (V & 2) ? 1 : 0; xo < 16; xo++) {
                    float* h0 = hc + hoff[xo];
                    const float a0 = h0[0], a1 = h0[32];
                    h0[0] = a0 + vo0[xo];
                    h0[32] = a1 + vo1[xo];
                }
            }
        }
        // circular fold (hash_sift.cpp:299-302): bin0 += bin8 (bin 9 is never written: the bin number is <= 7)
        float v[8];
        v[0] = hc[0] + hc[8 * 32];
#pragma unroll
        for (int b = 1; b < 8; b++) v[b] = hc[b * 32];
        __syncwarp();                   // every lane is done with the records: their first 128 floats become the squares scratch
        // ---- L2 normalise, clip 0.2, renormalise, x512 -> uchar (hash_sift.cpp:311-330)
        ef_sift_normalize(v, blk, hl);
#pragma unroll
        for (int j = 0; j < 8; j++) v[j] = fminf(v[j], 0.2f);
        ef_sift_normalize(v, blk, hl);
        unsigned packed[2] = { 0, 0 };
#pragma unroll
        for (int j = 0; j < 8; j++) {
            const int q = __float2int_rn(512.f * v[j]);
            packed[j >> 2] |= (unsigned)min(max(q, 0), 255) << (8 * (j & 3));
        }
        if (store) reinterpret_cast<uint2*>(out128)[hl] = make_uint2(packed[0], packed[1]);
    }
    __syncwarp();
}

template <bool STAGED>
__global__ void __launch_bounds__(EF_SIFT_WARPS * 32) ef_hashsift_flat_kernel(const EfDescJob job, const EfHashSiftTables t, uint8_t* __restrict__ sift128)
{
    extern __shared__ __align__(16) unsigned char s_raw[];
    EfSiftWarpSmem* sm = reinterpret_cast<EfSiftWarpSmem*>(s_raw);
    const int slot = threadIdx.x >> 4, warp = threadIdx.x >> 5;
    const int first = blockIdx.x * EF_SIFT_KP_PER_CTA + (slot & ~1); // first keypoint of this warp
    if (first >= job.n) return;
    const int i = blockIdx.x * EF_SIFT_KP_PER_CTA + slot;
    const bool valid = i < job.n;
    const int ii = valid ? i : first;
    const float4 k = job.kpts[ii];
    ef_hashsift_one<STAGED, 15>(job.img, job.w, job.h, job.pitch, k.x, k.y, k.z, k.w, job.scale, t, sm[warp], sift128 + (size_t)ii * 128, valid);
}

void ef_launch_hashsift_features_flat(const EfDescJob& job, const EfHashSiftTables& t, uint8_t* sift128, cudaStream_t s)
{
    if (job.n <= 0) return;
    const size_t smem = sizeof(EfSiftWarpSmem) * EF_SIFT_WARPS;
    // staged31: integer keypoints of size 31 at scale 1 (the 5 x N GpuMat compute path) take the window-staging form of the pipeline kernel
    if (job.staged31) ef_hashsift_flat_kernel<true><<<ef_div_up(job.n, EF_SIFT_KP_PER_CTA), EF_SIFT_WARPS * 32, smem, s>>>(job, t, sift128);
    else ef_hashsift_flat_kernel<false><<<ef_div_up(job.n, EF_SIFT_KP_PER_CTA), EF_SIFT_WARPS * 32, smem, s>>>(job, t, sift128);
    EF_COUNT_LAUNCH(1);
}

template <int V>
__global__ void __launch_bounds__(EF_SIFT_WARPS * 32) ef_hashsift_pipe_kernel(const __grid_constant__ EfPipe p, const EfHashSiftTables t, uint8_t* __restrict__ sift128)
{
    extern __shared__ __align__(16) unsigned char s_raw[];
    EfSiftWarpSmem* sm = reinterpret_cast<EfSiftWarpSmem*>(s_raw);
    const int slot = threadIdx.x >> 4, warp = threadIdx.x >> 5;
    const int frame = blockIdx.y;
    const EfLevelCounters* ctr = &p.counters[frame * EF_MAX_LEVELS];
    const int bx = (int)blockIdx.x * p.shard_n + p.shard_i;   // CTAs dealt round-robin over the GPUs of a band-sharded frame
    if (bx >= p.total_sift_blocks) return;
    int level = p.first_level;
    while (level + 1 < p.nlevels && bx >= p.lv[level + 1].sift_block_start) level++;
    const EfLevel& L = p.lv[level];
    const int nsel = ctr[level].selected;
    const int first = (bx - L.sift_block_start) * EF_SIFT_KP_PER_CTA + (slot & ~1);
    if (first >= nsel) return;
    const int i = (bx - L.sift_block_start) * EF_SIFT_KP_PER_CTA + slot;
    int offset = 0;
    for (int l = p.first_level; l < level; l++) offset += ctr[l].selected;
    // band-sharded frame: only the output rows of this GPU's slice; a warp whose two keypoints belong to other GPUs leaves
    const bool valid = i < nsel && offset + i < p.nfeatures && (unsigned)(offset + i - p.desc_row0) < (unsigned)p.desc_rows;
    if (!__any_sync(0xffffffffu, valid)) return;
    const int ii = valid ? i : first;
    const EfSelected k = reinterpret_cast<const EfSelected*>(ef_ws(p, frame, L.sel_off))[ii];
    const uint8_t* __restrict__ img = ef_ws(p, frame, L.blur_off);
    // describer created with croppingScale 1, keypoint size PATCH_SIZE (cuda_efficient_features.cpp:58-62, .cu:260)
    ef_hashsift_one<true, V>(img, L.w, L.h, L.blur_pitch, (float)k.x, (float)k.y, EF_PATCH_SIZE, k.angle, 1.f, t, sm[warp],
                          sift128 + ((size_t)frame * p.nfeatures + offset + ii) * 128, valid);
}

void ef_launch_hashsift_features_pipe(const EfPipe& p, const EfHashSiftTables& t, uint8_t* sift128, cudaStream_t s)
{
    if (p.total_sift_blocks <= 0) return;
    const size_t smem = sizeof(EfSiftWarpSmem) * EF_SIFT_WARPS;
    // A/B switch between generations of the same kernel (all bit-identical, all on the GPU): V bit 0 = all-zero spare record for
    // out-of-patch visits, bit 1 = zero-weight visits skipped, bit 2 = packed-fp32 sampling, bit 3 = packed-fp32 histogram shares.
    // Measured per 8 frames of 4K: V=0 1.96 ms, 3 1.85 ms, 7 1.80 ms, 15 (default) 1.79 ms.
    static const int variant = getenv("EF_SIFT_V") ? atoi(getenv("EF_SIFT_V")) : 15;
    const dim3 grid(ef_div_up(p.total_sift_blocks, p.shard_n), p.nframes);
    switch (variant) {
    case 0: ef_hashsift_pipe_kernel<0><<<grid, EF_SIFT_WARPS * 32, smem, s>>>(p, t, sift128); break;
    case 3: ef_hashsift_pipe_kernel<3><<<grid, EF_SIFT_WARPS * 32, smem, s>>>(p, t, sift128); break;
    case 7: ef_hashsift_pipe_kernel<7><<<grid, EF_SIFT_WARPS * 32, smem, s>>>(p, t, sift128); break;
    default: ef_hashsift_pipe_kernel<15><<<grid, EF_SIFT_WARPS * 32, smem, s>>>(p, t, sift128); break;
    }
    EF_COUNT_LAUNCH(1);
}
