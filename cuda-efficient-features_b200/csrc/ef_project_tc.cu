// ef_project_tc.cu -- HashSIFT learned projection on the 5th-generation tensor cores (tcgen05.mma kind::i8, sm_100a).
//
// Same exact arithmetic as ef_hashsift_project_imma_kernel (ef_project.cu): weights = six balanced base-256 s8 digits of a
// 47-bit fixed-point number, six u8 x s8 -> s32 GEMMs, int64 recombination + bias, sign -> bit.  What changes is the machine:
//   * A (128 keypoint rows x 128 k, u8) and B (32 output bits x 128 k x 6 digits, s8) sit in shared memory in the UMMA
//     no-swizzle K-major "core matrix" layout (8 rows x 16 bytes contiguous; LBO = 128 B between the k-halves of one MMA,
//     SBO = 1024 B between 8-row groups); B is pre-packed in that byte order on the host and streamed with cp.async.
//   * one elected thread issues 4 tcgen05.mma (M128 x N192 x K32: the six digit blocks of a chunk form ONE B operand) per 32
//     output bits; the six digit accumulators are 6 x 32 TMEM columns (256 columns allocated per CTA: two CTAs per SM share the 512 columns).
//   * completion comes back through tcgen05.commit -> mbarrier; the epilogue reads the accumulators with tcgen05.ld (thread =
//     keypoint row x half of the chunk's columns, TMEM lane = row), recombines in int64 and writes 16 descriptor bits per store.
//   * the next 24 KB of B digits are in flight (cp.async double buffer) while the tensor core and the epilogue run.
// Every spin on the mbarrier is bounded (trap instead of a hung GPU if a descriptor were ever wrong).
//
// ND = 6 or 7 digits.  The published 512-bit table is a multiple of 2^-44 below 2^3: 47-bit integers, six digits.  The 256-bit
// table (the reference's DEFAULT descriptor type) is a multiple of 2^-48 with weights up to 0.987: 49-bit integers, one bit too
// wide for six balanced digits -- it takes a seventh digit block (N = 224, still one MMA per k-step, 224 of the 256 TMEM
// columns) and a two-limb recombination (units of 2^32) whose sign is exact for EVERY u8 input even where the total leaves int64.
#include "ef_common.cuh"

#define EF_TC_ROWS 128               // keypoint rows per CTA (= UMMA M = TMEM lanes)
#ifndef EF_TC_THREADS
#define EF_TC_THREADS 128            // 128: one thread per row (0.163 ms per 8 frames); 256: two threads per row, 16 output bits each (0.175 ms)
#endif
#define EF_TC_TPR (EF_TC_THREADS / EF_TC_ROWS)             // threads per row
#define EF_TC_NCH 32                 // output bits per chunk (= UMMA N)
#define EF_TC_A_BYTES (EF_TC_ROWS * 128)
#define EF_TC_BD_BYTES (EF_TC_NCH * 128)                  // one digit of one chunk: 4 KB
#define EF_TC_B_BYTES(ND) ((ND) * EF_TC_BD_BYTES)         // 24 / 28 KB per chunk
#define EF_TC_TMEM_COLS 256                               // power of two >= 7 * 32
#define EF_TC_SMEM(ND) (EF_TC_A_BYTES + 2 * EF_TC_B_BYTES(ND))   // 64 / 72 KB dynamic

__device__ __forceinline__ unsigned ef_smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }

// UMMA shared-memory descriptor, K-major, no swizzle (cute::UMMA::SmemDescriptor: start >> 4 at [0,14), LBO >> 4 at [16,30),
// SBO >> 4 at [32,46), version 1 at [46,48), layout type 0 at [61,64))
__device__ __forceinline__ unsigned long long ef_umma_desc(unsigned smem_addr, unsigned lbo_bytes, unsigned sbo_bytes)
{
    return (unsigned long long)((smem_addr & 0x3ffffu) >> 4) | ((unsigned long long)(lbo_bytes >> 4) << 16) |
           ((unsigned long long)(sbo_bytes >> 4) << 32) | (1ull << 46);
}
// instruction descriptor (cute::UMMA::InstrDescriptor): D = S32 (2 at [4,6)), A = U8 (0 at [7,10)), B = S8 (1 at [10,13)),
// both K-major, N >> 3 at [17,23), M >> 4 at [24,29)
// N = ND * 32 (192 / 224): the digit blocks are adjacent 8-row groups of ONE B operand
#define EF_TC_IDESC(ND) ((2u << 4) | (0u << 7) | (1u << 10) | ((unsigned)(((ND) * EF_TC_NCH) >> 3) << 17) | ((unsigned)(EF_TC_ROWS >> 4) << 24))

__device__ __forceinline__ void ef_tc_mma_i8(unsigned d_tmem, unsigned long long adesc, unsigned long long bdesc, unsigned idesc, unsigned accumulate)
{
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                 "tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;\n\t}"
                 ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void ef_tc_ld16(unsigned taddr, int (&v)[16])
{
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
                   "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
                 : "r"(taddr) : "memory");
}
__device__ __forceinline__ void ef_tc_cp_async16(unsigned smem, const void* gmem)
{
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem), "l"(gmem) : "memory");
}

template <int ND>
__global__ void __launch_bounds__(EF_TC_THREADS, 2)
ef_hashsift_project_tc_kernel(const uint8_t* __restrict__ sift128, int n_cap, const int* __restrict__ d_n, size_t frame_rows,
                              const uint8_t* __restrict__ btc, const long long* __restrict__ bias, int S, int nchunks,
                              uint8_t* __restrict__ desc, size_t desc_stride, int desc_pitch, float* __restrict__ proj_out)
{
    extern __shared__ __align__(1024) uint8_t s_dyn[];
    __shared__ __align__(8) unsigned long long s_mbar;
    __shared__ unsigned s_tmem;
    __shared__ long long s_bias[512];

    const int frame = blockIdx.y;
    const int n = d_n ? min(d_n[frame], n_cap) : n_cap;
    const int row0 = blockIdx.x * EF_TC_ROWS;
    if (row0 >= n) return;                                   // CTA-uniform, before any allocation
    const int tid = threadIdx.x, warp = tid >> 5;
    const int rl = tid & (EF_TC_ROWS - 1), hf = tid >> 7;     // row of the tile, half of the chunk's columns (warps w and w + 4 share TMEM lanes)
    const unsigned sA = ef_smem_u32(s_dyn), sB = sA + EF_TC_A_BYTES;
    const unsigned mbar = ef_smem_u32(&s_mbar);

    // ---- setup: TMEM allocation (warp 0), mbarrier, bias, first B chunk in flight, A tile into the core-matrix layout
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(ef_smem_u32(&s_tmem)), "n"(EF_TC_TMEM_COLS) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(mbar) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    for (int i = tid; i < nchunks * EF_TC_NCH; i += EF_TC_THREADS) s_bias[i] = bias[i];
    for (int i = tid; i < EF_TC_B_BYTES(ND) / 16; i += EF_TC_THREADS) ef_tc_cp_async16(sB + 16 * i, btc + 16 * (size_t)i);
    asm volatile("cp.async.commit_group;" ::: "memory");
    {
        // EF_TC_TPR threads per row, 8 / EF_TC_TPR chunks of 16 k-bytes each -> (row / 8) * 1024 + chunk * 128 + (row % 8) * 16
        const int r = row0 + rl;
        const uint4* src = reinterpret_cast<const uint4*>(sift128 + ((size_t)frame * frame_rows + min(r, n - 1)) * 128) + (8 / EF_TC_TPR) * hf;
        uint8_t* dstA = s_dyn + (rl >> 3) * 1024 + (rl & 7) * 16 + (8 / EF_TC_TPR) * hf * 128;
#pragma unroll
        for (int c = 0; c < 8 / EF_TC_TPR; c++) {
            uint4 v = __ldg(src + c);
            if (r >= n) v = make_uint4(0, 0, 0, 0);
            *reinterpret_cast<uint4*>(dstA + c * 128) = v;
        }
    }
    asm volatile("cp.async.wait_all;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic-proxy smem writes -> visible to the tensor core (async proxy)
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const unsigned tmem = s_tmem;

    const float unscale = __int_as_float((127 - S) << 23);    // 2^-S
    const int row = row0 + rl;
    uint8_t* out = desc + (size_t)frame * desc_stride + (size_t)row * desc_pitch;
    const bool half_ok = ((reinterpret_cast<uintptr_t>(desc) | desc_stride | (size_t)desc_pitch) & 1) == 0;
    const int nbits = nchunks * EF_TC_NCH;

    for (int c = 0; c < nchunks; c++) {
        const unsigned sBc = sB + (c & 1) * EF_TC_B_BYTES(ND);
        if (tid == 0) {
            // 4 k-steps of M128 x N192 x K32: the six 32-row digit blocks of the chunk are contiguous in shared memory (4 KB each =
            // four 8-row groups of SBO bytes), so one MMA covers all digits; digit d lands in TMEM columns [32 d, 32 d + 32)
#pragma unroll
            for (int ks = 0; ks < 4; ks++)
                ef_tc_mma_i8(tmem, ef_umma_desc(sA + ks * 256, 128, 1024), ef_umma_desc(sBc + ks * 256, 128, 1024), EF_TC_IDESC(ND), ks > 0);
            // arrives on the mbarrier when all MMAs above have completed (implies tcgen05.fence::before_thread_sync)
            asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(mbar) : "memory");
        }
        if (c + 1 < nchunks) {
            // next chunk's digits into the other buffer (its last reader, the MMAs of chunk c-1, completed before the previous wait returned)
            const uint8_t* g = btc + (size_t)(c + 1) * EF_TC_B_BYTES(ND);
            const unsigned sBn = sB + ((c + 1) & 1) * EF_TC_B_BYTES(ND);
            for (int i = tid; i < EF_TC_B_BYTES(ND) / 16; i += EF_TC_THREADS) ef_tc_cp_async16(sBn + 16 * i, g + 16 * (size_t)i);
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
        {
            // bounded wait for the MMAs of this chunk (phase parity = c & 1)
            unsigned done = 0;
            for (int spin = 0; spin < (1 << 22) && !done; spin++)
                asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                             : "=r"(done) : "r"(mbar), "r"((unsigned)(c & 1)) : "memory");
            if (!done) __trap();
        }
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");

        // ---- epilogue: TMEM lane = row (warps w and w + 4 own lanes 32 (w % 4) .. + 31), thread = 16 of the chunk's 32 columns
#pragma unroll
        for (int hh = 0; hh < 2 / EF_TC_TPR; hh++) {
            const int h2 = EF_TC_TPR == 2 ? hf : hh;         // which 16 columns of the chunk
            unsigned bits = 0;
            int a[ND][16];
#pragma unroll
            for (int d = 0; d < ND; d++) ef_tc_ld16(tmem + ((unsigned)(32 * (warp & 3)) << 16) + 32 * d + 16 * h2, a[d]);
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
            for (int j = 0; j < 16; j++) {
                const int t01 = a[0][j] + a[1][j] * 256;
                const int t23 = a[2][j] + a[3][j] * 256;
                const int t45 = a[4][j] + a[5][j] * 256;
                const int jj = 16 * h2 + j;
                bool positive;
                float value;
                if (ND == 6) {
                    const long long tot = (long long)t01 + ((long long)t23 << 16) + ((long long)t45 << 32) + s_bias[c * EF_TC_NCH + jj];
                    positive = tot > 0;
                    value = __ll2float_rn(tot) * unscale;
                } else {
                    // two limbs in units of 2^32: total = H * 2^32 + L with 0 <= L < 2^32 -- the sign is exact even where the total
                    // (up to 128 * 255 * 2^48) leaves int64; the float value is one rounding of the exact total whenever |total| < 2^62,
                    // which holds for every vector the feature kernel can produce (sum of the components <= 512 * sqrt(128))
                    const long long bias_j = s_bias[c * EF_TC_NCH + jj];
                    const long long lo = (long long)t01 + ((long long)t23 << 16) + (bias_j & 0xffffffffll);
                    const long long H = (long long)t45 + ((long long)a[ND - 1][j] << 16) + (bias_j >> 32) + (lo >> 32);
                    const unsigned long long L = (unsigned long long)lo & 0xffffffffull;
                    positive = H > 0 || (H == 0 && L != 0);
                    const bool fits = H < (1ll << 30) && H > -(1ll << 30);
                    value = fits ? __ll2float_rn((H << 32) + (long long)L) * unscale
                                 : (float)((double)H * 4294967296.0 + (double)L) * unscale;
                }
                bits |= (positive ? 1u : 0u) << (8 * (j >> 3) + 7 - (j & 7));        // MSB first inside every byte
                if (proj_out && row < n) proj_out[((size_t)frame * frame_rows + row) * nbits + c * EF_TC_NCH + jj] = value;
            }
            if (row < n) {
                uint8_t* o = out + 4 * c + 2 * h2;
                if (half_ok) *reinterpret_cast<unsigned short*>(o) = (unsigned short)bits;
                else { o[0] = (uint8_t)bits; o[1] = (uint8_t)(bits >> 8); }
            }
        }
        // the next chunk's MMAs overwrite the accumulators and read the freshly copied digits
        asm volatile("cp.async.wait_all;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        __syncthreads();
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    }
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "n"(EF_TC_TMEM_COLS) : "memory");
}

bool ef_launch_hashsift_project_tc(const uint8_t* sift128, int n_cap, const int* d_counts, int nframes, const EfProjTables& t, int nbits,
                                   uint8_t* desc, size_t desc_stride, int desc_pitch, float* proj_out, cudaStream_t s)
{
    if (!t.btc || nbits % EF_TC_NCH != 0 || nbits > 512 || (t.ndigits != 6 && t.ndigits != 7)) return false;
    static unsigned long long configured = 0;   // function attributes are per device
    int dev = 0;
    cudaGetDevice(&dev);
    if (!((__atomic_load_n(&configured, __ATOMIC_RELAXED) >> (dev & 63)) & 1ull)) {
        cudaFuncSetAttribute(ef_hashsift_project_tc_kernel<6>, cudaFuncAttributeMaxDynamicSharedMemorySize, EF_TC_SMEM(6));
        cudaFuncSetAttribute(ef_hashsift_project_tc_kernel<7>, cudaFuncAttributeMaxDynamicSharedMemorySize, EF_TC_SMEM(7));
        __atomic_fetch_or(&configured, 1ull << (dev & 63), __ATOMIC_RELAXED);
    }
    const dim3 grid(ef_div_up(n_cap, EF_TC_ROWS), nframes);
    if (t.ndigits == 6)
        ef_hashsift_project_tc_kernel<6><<<grid, EF_TC_THREADS, EF_TC_SMEM(6), s>>>(sift128, n_cap, d_counts, (size_t)n_cap, t.btc, t.bias, t.shift, nbits / EF_TC_NCH,
                                                                                 desc, desc_stride, desc_pitch, proj_out);
    else
        ef_hashsift_project_tc_kernel<7><<<grid, EF_TC_THREADS, EF_TC_SMEM(7), s>>>(sift128, n_cap, d_counts, (size_t)n_cap, t.btc, t.bias, t.shift, nbits / EF_TC_NCH,
                                                                                 desc, desc_stride, desc_pitch, proj_out);
    return true;
}
