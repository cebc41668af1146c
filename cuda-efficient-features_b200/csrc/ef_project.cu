// ef_project.cu -- HashSIFT learned projection + sign + pack (hash_sift.cpp:353-378; replaces cublasSgemm +
// binarizeDescriptorsKernel, cuda_hash_sift.cpp:44-60, cuda_hash_sift.cu:414-435).
//
//   out[i][j] = [1, d_i0 .. d_i127] . B_j      (B = nbits x 129 fp32),  bit = out > 0, MSB first
//
// The SIFT vector d is u8-valued, so the contraction is done EXACTLY on the integer tensor cores:
//   * every weight of a table is a multiple of 2^-S (S = 44 for the published tables) below 2^47 * 2^-S, i.e. a
//     47-bit signed fixed-point number; it is split on the host into six balanced base-256 digits
//     w = sum_s digit_s * 256^s * 2^-S with digit_s in [-128, 127]  (s8 operands);
//   * six u8 x s8 -> s32 GEMMs (mma.sync m16n8k32, IMMA.16832.U8.S8) give exact digit sums
//     (|sum| <= 128 * 255 * 128 < 2^22), recombined in int64 with the bias B_j0 (exact, |total| < 2^62);
//   * the sign of that integer is the descriptor bit; float(total) * 2^-S -- ONE rounding of the exact dot
//     product -- is the value kept for the debug/ULP check.
// ef_hashsift_project_kernel below (fp64 CUDA cores, double accumulation in ascending k) is the generic fallback
// for a table that does not fit 6 digits.
#include "ef_common.cuh"

#include <cstdlib>

// =================================================================================================
// exact integer-tensor-core path
//   CTA = 4 warps = 256 keypoint rows; warp = 64 rows = 4 m16 tiles, A fragments (64 rows x 128 k, u8) live in
//   64 registers for the whole kernel.  The B digits are pre-packed on the host in fragment order
//   bfrag[ntile][digit][half][lane] (uint4 = b0,b1 of k-step 2*half and of k-step 2*half+1), streamed through a
//   cp.async double buffer (6 KB per 8 output bits) shared by the 4 warps.
// =================================================================================================
#define EF_PJ_ROWS_PER_WARP 64
#define EF_PJ_WARPS 4
#define EF_PJ_DIGITS 6
#define EF_PJ_STAGE_U4 (EF_PJ_DIGITS * 2 * 32) // uint4 per n-tile

__device__ __forceinline__ void ef_imma_u8s8(int (&c)[4], const unsigned (&a)[4], unsigned b0, unsigned b1)
{
    asm volatile("mma.sync.aligned.m16n8k32.row.col.s32.u8.s8.s32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+r"(c[0]), "+r"(c[1]), "+r"(c[2]), "+r"(c[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ void ef_cp_async16(void* smem, const void* gmem)
{
    const unsigned sa = (unsigned)__cvta_generic_to_shared(smem);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(sa), "l"(gmem));
}
__device__ __forceinline__ void ef_cp_async_commit() { asm volatile("cp.async.commit_group;"); }
template <int N> __device__ __forceinline__ void ef_cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N)); }

__global__ void __launch_bounds__(EF_PJ_WARPS * 32, 2)
ef_hashsift_project_imma_kernel(const uint8_t* __restrict__ sift128, int n_cap, const int* __restrict__ d_n, size_t frame_rows,
                                const uint4* __restrict__ bfrag, const long long* __restrict__ bias, int S, int ntiles,
                                uint8_t* __restrict__ desc, size_t desc_stride, int desc_pitch, float* __restrict__ proj_out)
{
    __shared__ __align__(16) uint4 s_b[2][EF_PJ_STAGE_U4];

    const int frame = blockIdx.y;
    const int n = d_n ? min(d_n[frame], n_cap) : n_cap;
    const int cta_row0 = blockIdx.x * (EF_PJ_WARPS * EF_PJ_ROWS_PER_WARP);
    if (cta_row0 >= n) return;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int g8 = lane >> 2, q = lane & 3;
    const int row0 = cta_row0 + warp * EF_PJ_ROWS_PER_WARP;
    const uint8_t* src = sift128 + (size_t)frame * frame_rows * 128;
    const int nbits = ntiles * 8;

    // prologue: first B stage in flight while the A fragments are loaded
    for (int i = tid; i < EF_PJ_STAGE_U4; i += EF_PJ_WARPS * 32) ef_cp_async16(&s_b[0][i], bfrag + i);
    ef_cp_async_commit();

    // A fragments (m16n8k32 .row u8): a0 = (row g8, k 4q..4q+3), a1 = (row g8+8, same k), a2/a3 = k + 16
    unsigned a[4][4][4];
#pragma unroll
    for (int mt = 0; mt < 4; mt++) {
        const int ra = row0 + mt * 16 + g8, rb = ra + 8;
        const unsigned* pa = reinterpret_cast<const unsigned*>(src + (size_t)ra * 128) + q;
        const unsigned* pb = reinterpret_cast<const unsigned*>(src + (size_t)rb * 128) + q;
#pragma unroll
        for (int ks = 0; ks < 4; ks++) {
            a[mt][ks][0] = ra < n ? __ldg(pa + ks * 8) : 0u;
            a[mt][ks][1] = rb < n ? __ldg(pb + ks * 8) : 0u;
            a[mt][ks][2] = ra < n ? __ldg(pa + ks * 8 + 4) : 0u;
            a[mt][ks][3] = rb < n ? __ldg(pb + ks * 8 + 4) : 0u;
        }
    }
    const float unscale = __int_as_float((127 - S) << 23); // 2^-S
    uint8_t* out = desc + (size_t)frame * desc_stride;

    for (int nt = 0; nt < ntiles; nt++) {
        if (nt + 1 < ntiles) {
            const uint4* g = bfrag + (size_t)(nt + 1) * EF_PJ_STAGE_U4;
            for (int i = tid; i < EF_PJ_STAGE_U4; i += EF_PJ_WARPS * 32) ef_cp_async16(&s_b[(nt + 1) & 1][i], g + i);
        }
        ef_cp_async_commit();
        ef_cp_async_wait<1>();
        __syncthreads();

        const uint4* sb = s_b[nt & 1];
        int acc[EF_PJ_DIGITS][4][4];
#pragma unroll
        for (int s = 0; s < EF_PJ_DIGITS; s++) {
            const uint4 f0 = sb[(s * 2 + 0) * 32 + lane], f1 = sb[(s * 2 + 1) * 32 + lane];
#pragma unroll
            for (int mt = 0; mt < 4; mt++) {
#pragma unroll
                for (int e = 0; e < 4; e++) acc[s][mt][e] = 0;
                ef_imma_u8s8(acc[s][mt], a[mt][0], f0.x, f0.y);
                ef_imma_u8s8(acc[s][mt], a[mt][1], f0.z, f0.w);
                ef_imma_u8s8(acc[s][mt], a[mt][2], f1.x, f1.y);
                ef_imma_u8s8(acc[s][mt], a[mt][3], f1.z, f1.w);
            }
        }
        // epilogue: recombine the digits (exact), add the bias, take the sign.  Accumulator element e of tile mt:
        // row = mt*16 + (e>>1)*8 + g8, column = 2q + (e&1)
        const long long bias0 = __ldg(bias + nt * 8 + 2 * q), bias1 = __ldg(bias + nt * 8 + 2 * q + 1);
        unsigned P0 = 0, P1 = 0; // byte i of P0|P1<<32 = partial descriptor byte of thread-row i = mt*2 + (e>>1)
#pragma unroll
        for (int mt = 0; mt < 4; mt++) {
#pragma unroll
            for (int e = 0; e < 4; e++) {
                const int t01 = acc[0][mt][e] + acc[1][mt][e] * 256;
                const int t23 = acc[2][mt][e] + acc[3][mt][e] * 256;
                const int t45 = acc[4][mt][e] + acc[5][mt][e] * 256;
                const long long tot = (long long)t01 + ((long long)t23 << 16) + ((long long)t45 << 32) + ((e & 1) ? bias1 : bias0);
                const unsigned bit = tot > 0 ? 1u : 0u;
                const int i = mt * 2 + (e >> 1);
                const unsigned sh = (unsigned)(8 * (i & 3) + 7 - 2 * q - (e & 1));
                if (i < 4) P0 |= bit << sh; else P1 |= bit << sh;
                if (proj_out) {
                    const int row = row0 + mt * 16 + (e >> 1) * 8 + g8;
                    if (row < n) proj_out[((size_t)frame * frame_rows + row) * nbits + nt * 8 + 2 * q + (e & 1)] = __ll2float_rn(tot) * unscale;
                }
            }
        }
        P0 |= __shfl_xor_sync(0xffffffffu, P0, 1); P0 |= __shfl_xor_sync(0xffffffffu, P0, 2);
        P1 |= __shfl_xor_sync(0xffffffffu, P1, 1); P1 |= __shfl_xor_sync(0xffffffffu, P1, 2);
        // lane q of the quad stores thread-rows q and 4+q
        {
            const int ia = q, ib = 4 + q;
            const int rowa = row0 + (ia >> 1) * 16 + (ia & 1) * 8 + g8, rowb = row0 + (ib >> 1) * 16 + (ib & 1) * 8 + g8;
            if (rowa < n) out[(size_t)rowa * desc_pitch + nt] = (uint8_t)(P0 >> (8 * q));
            if (rowb < n) out[(size_t)rowb * desc_pitch + nt] = (uint8_t)(P1 >> (8 * q));
        }
        __syncthreads();
    }
}

// =================================================================================================
// fallback: out = float32(sum_k a_k * w_k) with the sum carried in double in ascending k (every product
// u8 x fp32 is exact in double).
// weights_t: 129 x nbits (transposed at create).  16 keypoints per CTA, one output bit column per thread.
// =================================================================================================
#define EF_PROJ_KP 16
template <int NCOL>
__global__ void __launch_bounds__(256) ef_hashsift_project_kernel(const uint8_t* __restrict__ sift128, int n_cap, const int* __restrict__ d_n,
                                                                  size_t frame_rows, const float* __restrict__ weights_t,
                                                                  uint8_t* __restrict__ desc, size_t desc_stride, int desc_pitch,
                                                                  float* __restrict__ proj_out)
{
    constexpr int nbits = 256 * NCOL;
    __shared__ __align__(16) double s_a[129][EF_PROJ_KP]; // [k][keypoint]: one 16-byte broadcast load feeds 2 keypoints
    const int frame = blockIdx.y;
    const int n = d_n ? min(d_n[frame], n_cap) : n_cap;
    const int k0 = blockIdx.x * EF_PROJ_KP;
    if (k0 >= n) return;
    const int tid = threadIdx.x, lane = tid & 31;
    const uint8_t* src = sift128 + (size_t)frame * frame_rows * 128;
    for (int i = tid; i < EF_PROJ_KP * 129; i += 256) {
        const int kk = i / 129, k = i - kk * 129;
        double v = 0.0;
        if (k0 + kk < n) v = (k == 0) ? 1.0 : (double)src[(size_t)(k0 + kk) * 128 + (k - 1)];
        s_a[k][kk] = v;
    }
    __syncthreads();
    double acc[NCOL][EF_PROJ_KP];
#pragma unroll
    for (int c = 0; c < NCOL; c++)
#pragma unroll
        for (int kk = 0; kk < EF_PROJ_KP; kk++) acc[c][kk] = 0.0;
    for (int k = 0; k < 129; k++) {
        double wv[NCOL];
#pragma unroll
        for (int c = 0; c < NCOL; c++) wv[c] = (double)__ldg(weights_t + (size_t)k * nbits + tid + 256 * c);
#pragma unroll
        for (int kp = 0; kp < EF_PROJ_KP; kp += 2) {
            const double2 a = *reinterpret_cast<const double2*>(&s_a[k][kp]);
#pragma unroll
            for (int c = 0; c < NCOL; c++) {
                acc[c][kp] = fma(a.x, wv[c], acc[c][kp]);
                acc[c][kp + 1] = fma(a.y, wv[c], acc[c][kp + 1]);
            }
        }
    }
#pragma unroll
    for (int c = 0; c < NCOL; c++) {
        const int j = tid + 256 * c;
#pragma unroll
        for (int kk = 0; kk < EF_PROJ_KP; kk++) {
            const float tv = (float)acc[c][kk];
            const bool valid = k0 + kk < n;
            const unsigned bal = __brev(__ballot_sync(0xffffffffu, tv > 0));
            if (valid) {
                if (proj_out) proj_out[((size_t)frame * frame_rows + k0 + kk) * nbits + j] = tv;
                if (lane < 4)
                    desc[(size_t)frame * desc_stride + (size_t)(k0 + kk) * desc_pitch + ((j & ~31) >> 3) + lane] = (uint8_t)(bal >> (24 - 8 * lane));
            }
        }
    }
}

int g_ef_project_path = 0;   // 0: default (tcgen05, EF_PROJECT=imma -> mma.sync); ef_debug_project_async forces 1 = tcgen05, 2 = mma.sync, 3 = fp64

static void ef_project_launch(const uint8_t* sift128, int n_cap, const int* d_counts, int nframes, const EfProjTables& t, int nbits,
                              uint8_t* desc, size_t desc_stride, int desc_pitch, float* proj_out, cudaStream_t s)
{
    // EF_PROJECT=imma keeps the mma.sync kernel (A/B comparison); default: tcgen05 (ef_project_tc.cu)
    static const bool env_tc = [] { const char* e = getenv("EF_PROJECT"); return !(e && e[0] == 'i'); }();
    const bool use_tc = g_ef_project_path == 0 ? env_tc : g_ef_project_path == 1;
    if (use_tc && t.btc && ef_launch_hashsift_project_tc(sift128, n_cap, d_counts, nframes, t, nbits, desc, desc_stride, desc_pitch, proj_out, s)) {
        EF_COUNT_LAUNCH(1);
        return;
    }
    if (t.bfrag && g_ef_project_path != 3) {
        const dim3 grid(ef_div_up(n_cap, EF_PJ_WARPS * EF_PJ_ROWS_PER_WARP), nframes);
        ef_hashsift_project_imma_kernel<<<grid, EF_PJ_WARPS * 32, 0, s>>>(sift128, n_cap, d_counts, (size_t)n_cap, t.bfrag, t.bias, t.shift, nbits / 8,
                                                                         desc, desc_stride, desc_pitch, proj_out);
    } else {
        const dim3 grid(ef_div_up(n_cap, EF_PROJ_KP), nframes);
        if (nbits == 256)
            ef_hashsift_project_kernel<1><<<grid, 256, 0, s>>>(sift128, n_cap, d_counts, (size_t)n_cap, t.weights_t, desc, desc_stride, desc_pitch, proj_out);
        else
            ef_hashsift_project_kernel<2><<<grid, 256, 0, s>>>(sift128, n_cap, d_counts, (size_t)n_cap, t.weights_t, desc, desc_stride, desc_pitch, proj_out);
    }
    EF_COUNT_LAUNCH(1);
}

void ef_launch_hashsift_project(const uint8_t* sift128, int n_cap, const int* d_n, const EfProjTables& t, int nbits,
                                uint8_t* desc, int desc_pitch, float* proj_out, cudaStream_t s)
{
    if (n_cap <= 0) return;
    ef_project_launch(sift128, n_cap, d_n, 1, t, nbits, desc, 0, desc_pitch, proj_out, s);
}

void ef_launch_hashsift_project_batch(const uint8_t* sift128, int n_cap, const int* d_counts, int nframes, const EfProjTables& t, int nbits,
                                      uint8_t* desc, size_t desc_stride, int desc_pitch, float* proj_out, cudaStream_t s)
{
    if (n_cap <= 0 || nframes <= 0) return;
    ef_project_launch(sift128, n_cap, d_counts, nframes, t, nbits, desc, desc_stride, desc_pitch, proj_out, s);
}
