// ef_match_tc.cu -- brute-force Hamming k-nearest (k <= 2) on the 5th-generation tensor cores (tcgen05.mma kind::i8, sm_100a).
//
// Bits become +-1 int8: dot(a, b) = K - 2 * hamming(a, b) with K = 256 or 512, so the nq x nt distance matrix is one s8 x s8 -> s32
// GEMM (exact) and the nearest neighbours are the LARGEST dot products.  Semantics = ef_match_knn2_kernel / OpenCV (ef_match.cu):
// the k lexicographically smallest (distance, trainIdx).
//
//   expand   descriptors -> +-1 bytes, written directly in the UMMA no-swizzle K-major core-matrix order of a 128-row tile
//            ((r / 8) * (K / 16) * 128 + (k / 16) * 128 + (r % 8) * 16 + (k % 16)): a tile is one contiguous 128 * K byte block that a
//            single thread moves with cp.async.bulk (TMA) onto an mbarrier -- no per-thread copy loop.
//   gemm     default: ef_match_tc2_kernel (K-sliced, warp-specialised, 256 query rows per CTA; see its header below).
//            EF_MATCH=tc1: ef_match_tc_kernel, CTA = 128 query rows (A tile resident) x the train tiles of its split, B tiles double
//            buffered (192 KB at K = 512); one elected thread issues K / 32 M128 x N128 x K32 MMAs per tile into one of two
//            128-column TMEM accumulators: the tensor core works on tile t+1 while all 8 warps read tile t back (tcgen05.ld, thread =
//            query row x half of the columns) and the TMA engine fetches tile t+2.  Bound by the L2 -> shared-memory stream of train tiles.
//   select   per thread a running top-2 of (dot, index); the maximum of every 8 columns (VIMNMX3 tree) filters out the groups that
//            cannot change it, so the common case costs half an instruction per pair.  Partial lists (one per split and column half) are
//            merged lexicographically by ef_match_merge_lex_kernel.
// Every mbarrier spin is bounded (trap instead of a hung GPU).
#include "ef_common.cuh"

#include <climits>
#include <cstdlib>

#define EF_MTC_ROWS 128
#define EF_MTC_MAX_LISTS 24            // partial top-2 lists per query the scratch buffer is sized for
#define EF_MTC_THREADS 256
#define EF_MTC_TMEM_COLS 256

__device__ __forceinline__ unsigned ef_mtc_smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ unsigned long long ef_mtc_desc(unsigned smem_addr, unsigned lbo_bytes, unsigned sbo_bytes)
{
    return (unsigned long long)((smem_addr & 0x3ffffu) >> 4) | ((unsigned long long)(lbo_bytes >> 4) << 16) |
           ((unsigned long long)(sbo_bytes >> 4) << 32) | (1ull << 46);
}
// D = S32, A = S8, B = S8, K-major both, N = 128, M = 128 (cute::UMMA::InstrDescriptor)
#define EF_MTC_IDESC ((2u << 4) | (1u << 7) | (1u << 10) | ((128u >> 3) << 17) | ((128u >> 4) << 24))

__device__ __forceinline__ void ef_mtc_mma(unsigned d_tmem, unsigned long long adesc, unsigned long long bdesc, unsigned accumulate)
{
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                 "tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;\n\t}"
                 ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(EF_MTC_IDESC), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void ef_mtc_ld32(unsigned taddr, int (&v)[32])
{
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,"
                 "%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
                   "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]),
                   "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]),
                   "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
                 : "r"(taddr) : "memory");
}
__device__ __forceinline__ void ef_mtc_wait(unsigned mbar, unsigned parity)
{
    unsigned done = 0;
    for (int spin = 0; spin < (1 << 24) && !done; spin++)
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(done) : "r"(mbar), "r"(parity) : "memory");
    if (!done) __trap();
}
// one thread: announce `bytes` on the barrier and start the bulk copies (16 KB pieces)
__device__ __forceinline__ void ef_mtc_bulk_load(unsigned smem_dst, const uint8_t* gsrc, unsigned bytes, unsigned mbar, bool first_of_phase, unsigned total_bytes)
{
    if (first_of_phase) asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(mbar), "r"(total_bytes) : "memory");
    for (unsigned off = 0; off < bytes; off += 16384)
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                     ::"r"(smem_dst + off), "l"(gsrc + off), "r"(min(16384u, bytes - off)), "r"(mbar) : "memory");
}

// bits -> +-1 bytes in operand order.  One thread = one 16-byte chunk (16 bits = 2 descriptor bytes) of one row.
//   sliced == 0 (A role, and B role of the one-tile kernel): row-group major, (r / 8) * (K / 16) * 128 + chunk * 128 + (r % 8) * 16
//   sliced == 1 (B role of the K-sliced kernel): per 128-row tile, K is cut into slices of 8 chunks (128 bytes per row): every (tile, slice)
//                is one contiguous 16 KB block  tile * 128 K + slice * 16384 + (rl / 8) * 1024 + (chunk % 8) * 128 + (rl % 8) * 16
__global__ void __launch_bounds__(256) ef_match_expand_kernel(const uint8_t* __restrict__ desc, size_t pitch, int n, int desc_bytes, uint8_t* __restrict__ out, int rows_padded, int sliced)
{
    const int kc = desc_bytes / 2;                           // 16-byte chunks per row (K / 16)
    const long long gid = (long long)blockIdx.x * 256 + threadIdx.x;
    if (gid >= (long long)rows_padded * kc) return;
    const int r = (int)(gid / kc), c = (int)(gid - (long long)r * kc);
    uint4 v = make_uint4(0, 0, 0, 0);                        // rows beyond n: zeros (dot 0, never accepted: index check in the GEMM)
    if (r < n) {
        const uint8_t* p = desc + (size_t)r * pitch + 2 * c;
        const unsigned bits = p[0] | ((unsigned)p[1] << 8);
        unsigned w[4];
#pragma unroll
        for (int i = 0; i < 4; i++) {
            const unsigned nib = (bits >> (4 * i)) & 0xfu;
            const unsigned b01 = (nib * 0x00204081u) & 0x01010101u;    // bit j of the nibble -> byte j (0 / 1)
            w[i] = (b01 * 0xfeu) ^ 0xffffffffu;                         // 1 -> 0x01 (+1), 0 -> 0xff (-1)
        }
        v = make_uint4(w[0], w[1], w[2], w[3]);
    }
    uint8_t* o;
    if (sliced) {
        const int tile = r >> 7, rl = r & 127;
        o = out + (size_t)tile * (128 * desc_bytes * 8) + (size_t)(c >> 3) * 16384 + (rl >> 3) * 1024 + (c & 7) * 128 + (rl & 7) * 16;
    } else {
        o = out + (size_t)(r >> 3) * (kc * 128) + (size_t)c * 128 + (r & 7) * 16;
    }
    *reinterpret_cast<uint4*>(o) = v;
}

// partial[(2 * split + half) * nq + q] = (dist0, idx0, dist1, idx1)
template <int K>
__global__ void __launch_bounds__(EF_MTC_THREADS, 1)
ef_match_tc_kernel(const uint8_t* __restrict__ qexp, int nq, const uint8_t* __restrict__ texp, int nt, int tiles_per_split, int4* __restrict__ partial)
{
    constexpr unsigned TILE_BYTES = 128u * K;                // 64 KB (K = 512) / 32 KB
    constexpr unsigned SBO = (K / 16) * 128;
    extern __shared__ __align__(1024) uint8_t s_dyn[];       // A | B0 | B1
    __shared__ __align__(8) unsigned long long s_full[2], s_done[2];
    __shared__ unsigned s_tmem;

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int ttiles = (nt + 127) >> 7;
    const int t_begin = blockIdx.y * tiles_per_split, T = min(tiles_per_split, ttiles - t_begin);
    if (T <= 0) return;                                       // CTA-uniform
    const unsigned sA = ef_mtc_smem_u32(s_dyn), sB = sA + TILE_BYTES;
    const unsigned full0 = ef_mtc_smem_u32(&s_full[0]), done0 = ef_mtc_smem_u32(&s_done[0]);

    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(ef_mtc_smem_u32(&s_tmem)), "n"(EF_MTC_TMEM_COLS) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    if (tid == 0) {
        for (int i = 0; i < 2; i++) {
            asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(full0 + 8 * i) : "memory");
            asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(done0 + 8 * i) : "memory");
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const unsigned tmem = s_tmem;

    const uint8_t* gA = qexp + (size_t)blockIdx.x * TILE_BYTES;
    const uint8_t* gB = texp + (size_t)t_begin * TILE_BYTES;
    auto issue_mma = [&](int t) {
        const unsigned sBt = sB + (t & 1) * TILE_BYTES;
#pragma unroll
        for (int ks = 0; ks < K / 32; ks++)
            ef_mtc_mma(tmem + 128 * (t & 1), ef_mtc_desc(sA + ks * 256, 128, SBO), ef_mtc_desc(sBt + ks * 256, 128, SBO), ks > 0);
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(done0 + 8 * (t & 1)) : "memory");
    };
    if (tid == 0) {
        ef_mtc_bulk_load(sA, gA, TILE_BYTES, full0, true, 2 * TILE_BYTES);           // A and B(0) land on full[0]
        ef_mtc_bulk_load(sB, gB, TILE_BYTES, full0, false, 0);
        if (T > 1) ef_mtc_bulk_load(sB + TILE_BYTES, gB + TILE_BYTES, TILE_BYTES, full0 + 8, true, TILE_BYTES);
        ef_mtc_wait(full0, 0);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        issue_mma(0);
    }

    // running top-2 by (dot desc, index asc) of this thread's row over its half of the columns
    const int q = blockIdx.x * EF_MTC_ROWS + 32 * (warp & 3) + lane, half = warp >> 2;
    int dot0 = INT_MIN, i0 = -1, dot1 = INT_MIN, i1 = -1;

    for (int t = 0; t < T; t++) {
        if (tid == 0 && t + 1 < T) {
            // tensor core: tile t+1 into the other accumulator (its last readers finished before the barrier that ended iteration t-1)
            ef_mtc_wait(full0 + 8 * ((t + 1) & 1), ((t + 1) >> 1) & 1);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            issue_mma(t + 1);
        }
        ef_mtc_wait(done0 + 8 * (t & 1), (t >> 1) & 1);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        if (tid == 0 && t + 2 < T)                             // MMA(t) is complete: its B buffer is free for tile t+2
            ef_mtc_bulk_load(sB + (t & 1) * TILE_BYTES, gB + (size_t)(t + 2) * TILE_BYTES, TILE_BYTES, full0 + 8 * (t & 1), true, TILE_BYTES);

        const unsigned tl = tmem + ((unsigned)(32 * (warp & 3)) << 16) + 128 * (t & 1) + 64 * half;
        const int idx_base = (t_begin + t) * 128 + 64 * half;
#pragma unroll
        for (int part = 0; part < 2; part++) {
            int a[32];
            ef_mtc_ld32(tl + 32 * part, a);
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
            // groups of 8 columns: a maximum (VIMNMX3) decides whether any of them can enter the top-2
#pragma unroll
            for (int g = 0; g < 4; g++) {
                const int m = max(max(max(a[8 * g], a[8 * g + 1]), max(a[8 * g + 2], a[8 * g + 3])), max(max(a[8 * g + 4], a[8 * g + 5]), max(a[8 * g + 6], a[8 * g + 7])));
                if (m > dot1) {
                    // columns arrive in increasing index, strict > keeps the earlier one on ties
#pragma unroll
                    for (int j = 8 * g; j < 8 * g + 8; j++) {
                        const int v = a[j], idx = idx_base + 32 * part + j;
                        if (v > dot1 && idx < nt) {
                            if (v > dot0) { dot1 = dot0; i1 = i0; dot0 = v; i0 = idx; }
                            else { dot1 = v; i1 = idx; }
                        }
                    }
                }
            }
        }
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        __syncthreads();
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    }
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "n"(EF_MTC_TMEM_COLS) : "memory");
    if (q < nq)
        partial[(size_t)(2 * blockIdx.y + half) * nq + q] = make_int4(i0 >= 0 ? (K - dot0) >> 1 : INT_MAX, i0, i1 >= 0 ? (K - dot1) >> 1 : INT_MAX, i1);
}

// ---- K-sliced, warp-specialised form: CTA = 256 query rows (two M = 128 A tiles resident: every train byte fetched from L2 feeds 256
// rows -- the one-tile kernel above is bound by that stream), train tiles of 128 rows streamed as K-slices of 16 KB through a 4-stage ring.
//   warp 8 (one lane): TMA producer   -- waits empty[stage], announces 16 KB on full[stage], cp.async.bulk
//   warp 9 (one lane): MMA issuer     -- waits acc_empty[buf]; per slice waits full[stage], issues 2 x 4 M128 x N128 x K32 MMAs (both A tiles)
//                                        and commits to empty[stage]; after the last slice commits to acc_full[buf]
//   warps 0-7        : epilogue        -- wait acc_full[buf], tcgen05.ld their 128 columns (warps 0-3: A tile 0, 4-7: A tile 1), running
//                                        top-2, arrive on acc_empty[buf] (256 arrivals)
// TMEM: 2 buffers x 2 A tiles x 128 columns = all 512 columns (one CTA per SM).
#define EF_MTC2_THREADS 320
#define EF_MTC2_STAGES 4
__device__ __forceinline__ void ef_mtc_arrive(unsigned mbar) { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(mbar) : "memory"); }

template <int K>
__global__ void __launch_bounds__(EF_MTC2_THREADS, 1)
ef_match_tc2_kernel(const uint8_t* __restrict__ qexp, int nq, const uint8_t* __restrict__ texp_sliced, int nt, int tiles_per_split, int4* __restrict__ partial)
{
    constexpr unsigned A_BYTES = 256u * K;                   // 128 KB (K = 512)
    constexpr unsigned SLICE_BYTES = 16384u;                 // 128 rows x 128 k-bytes
    constexpr int NS = K / 128;                              // slices per train tile
    constexpr unsigned SBO_A = (K / 16) * 128, SBO_B = 1024;
    extern __shared__ __align__(1024) uint8_t s_dyn[];       // A (256 rows) | ring of 4 slices
    __shared__ __align__(8) unsigned long long s_bar[2 * EF_MTC2_STAGES + 4];   // full[4] empty[4] acc_full[2] acc_empty[2]
    __shared__ unsigned s_tmem;

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int ttiles = (nt + 127) >> 7;
    const int t_begin = blockIdx.y * tiles_per_split, T = min(tiles_per_split, ttiles - t_begin);
    if (T <= 0) return;                                       // CTA-uniform
    const unsigned sA = ef_mtc_smem_u32(s_dyn), sR = sA + A_BYTES;
    const unsigned bar0 = ef_mtc_smem_u32(&s_bar[0]);
    const unsigned full0 = bar0, empty0 = bar0 + 8 * EF_MTC2_STAGES, accf0 = bar0 + 16 * EF_MTC2_STAGES, acce0 = accf0 + 16;

    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(ef_mtc_smem_u32(&s_tmem)), "n"(512) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    if (tid == 0) {
        for (int i = 0; i < EF_MTC2_STAGES; i++) {
            asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(full0 + 8 * i) : "memory");
            asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(empty0 + 8 * i) : "memory");
        }
        for (int i = 0; i < 2; i++) {
            asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(accf0 + 8 * i) : "memory");
            asm volatile("mbarrier.init.shared::cta.b64 [%0], 256;" ::"r"(acce0 + 8 * i) : "memory");
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const unsigned tmem = s_tmem;

    if (warp == 8) {
        if (lane == 0) {
            // ---- producer: the A block rides on the first slice's barrier
            const uint8_t* gB = texp_sliced + (size_t)t_begin * (128u * K);
            for (int idx = 0; idx < T * NS; idx++) {
                const int st = idx & (EF_MTC2_STAGES - 1), use = idx / EF_MTC2_STAGES;
                ef_mtc_wait(empty0 + 8 * st, (use & 1) ^ 1);
                if (idx == 0) {
                    ef_mtc_bulk_load(sA, qexp + (size_t)blockIdx.x * A_BYTES, A_BYTES, full0, true, A_BYTES + SLICE_BYTES);
                    ef_mtc_bulk_load(sR, gB, SLICE_BYTES, full0, false, 0);
                } else {
                    ef_mtc_bulk_load(sR + st * SLICE_BYTES, gB + (size_t)idx * SLICE_BYTES, SLICE_BYTES, full0 + 8 * st, true, SLICE_BYTES);
                }
            }
        }
    } else if (warp == 9) {
        if (lane == 0) {
            // ---- MMA issuer
            for (int t = 0; t < T; t++) {
                const int b = t & 1;
                ef_mtc_wait(acce0 + 8 * b, ((t >> 1) & 1) ^ 1);
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                for (int sl = 0; sl < NS; sl++) {
                    const int idx = t * NS + sl, st = idx & (EF_MTC2_STAGES - 1);
                    ef_mtc_wait(full0 + 8 * st, (idx / EF_MTC2_STAGES) & 1);
                    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
#pragma unroll
                    for (int a = 0; a < 2; a++)
#pragma unroll
                        for (int ks = 0; ks < 4; ks++)
                            ef_mtc_mma(tmem + 256 * b + 128 * a, ef_mtc_desc(sA + a * (A_BYTES / 2) + (sl * 4 + ks) * 256, 128, SBO_A),
                                       ef_mtc_desc(sR + st * SLICE_BYTES + ks * 256, 128, SBO_B), (sl | ks) != 0);
                    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(empty0 + 8 * st) : "memory");
                }
                asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(accf0 + 8 * b) : "memory");
            }
        }
    } else {
        // ---- epilogue: thread = query row
        const int atile = warp >> 2, q = blockIdx.x * 256 + 128 * atile + 32 * (warp & 3) + lane;
        int dot0 = INT_MIN, i0 = -1, dot1 = INT_MIN, i1 = -1;
        for (int t = 0; t < T; t++) {
            const int b = t & 1;
            ef_mtc_wait(accf0 + 8 * b, (t >> 1) & 1);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            const unsigned tl = tmem + ((unsigned)(32 * (warp & 3)) << 16) + 256 * b + 128 * atile;
            const int idx_base = (t_begin + t) * 128;
#pragma unroll
            for (int part = 0; part < 4; part++) {
                int a[32];
                ef_mtc_ld32(tl + 32 * part, a);
                asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
                for (int g = 0; g < 4; g++) {
                    const int m = max(max(max(a[8 * g], a[8 * g + 1]), max(a[8 * g + 2], a[8 * g + 3])), max(max(a[8 * g + 4], a[8 * g + 5]), max(a[8 * g + 6], a[8 * g + 7])));
                    if (m > dot1) {
#pragma unroll
                        for (int j = 8 * g; j < 8 * g + 8; j++) {
                            const int v = a[j], idx = idx_base + 32 * part + j;
                            if (v > dot1 && idx < nt) {
                                if (v > dot0) { dot1 = dot0; i1 = i0; dot0 = v; i0 = idx; }
                                else { dot1 = v; i1 = idx; }
                            }
                        }
                    }
                }
            }
            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
            ef_mtc_arrive(acce0 + 8 * b);
        }
        if (q < nq)
            partial[(size_t)blockIdx.y * nq + q] = make_int4(i0 >= 0 ? (K - dot0) >> 1 : INT_MAX, i0, i1 >= 0 ? (K - dot1) >> 1 : INT_MAX, i1);
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "n"(512) : "memory");
    }
}

// merge partial top-2 lists in lexicographic (distance, index) order (the lists cover arbitrary disjoint subsets of the train rows)
__global__ void __launch_bounds__(256) ef_match_merge_lex_kernel(const int4* __restrict__ partial, int nq, int nlists, int k, int* __restrict__ idx, int* __restrict__ dist)
{
    const int q = blockIdx.x * 256 + threadIdx.x;
    if (q >= nq) return;
    int d0 = INT_MAX, i0 = -1, d1 = INT_MAX, i1 = -1;
    auto less = [](int da, int ia, int db, int ib) { return ib < 0 || da < db || (da == db && ia < ib); };
    for (int s = 0; s < nlists; s++) {
        const int4 p = partial[(size_t)s * nq + q];
        const int pd[2] = { p.x, p.z }, pi[2] = { p.y, p.w };
#pragma unroll
        for (int e = 0; e < 2; e++) {
            if (pi[e] < 0) continue;
            if (less(pd[e], pi[e], d0, i0)) { d1 = d0; i1 = i0; d0 = pd[e]; i0 = pi[e]; }
            else if (less(pd[e], pi[e], d1, i1)) { d1 = pd[e]; i1 = pi[e]; }
        }
    }
    if (k == 2) { idx[2 * q] = i0; idx[2 * q + 1] = i1; dist[2 * q] = d0; dist[2 * q + 1] = d1; }
    else { idx[q] = i0; dist[q] = d0; }
}

// ---- host side ----------------------------------------------------------------------------------------------------------
// rows are padded to 256 so that either role (256-row A block, 128-row B tile) stays inside the buffer
size_t ef_match_tc_expanded_bytes(int n, int desc_bytes) { return (size_t)((n + 255) / 256) * 256 * desc_bytes * 8; }

// EF_MATCH=tc1 keeps the one-tile kernel (A/B comparison); default: the K-sliced warp-specialised kernel
static bool ef_match_tc_sliced()
{
    static const bool one_tile = [] { const char* e = getenv("EF_MATCH"); return e && e[0] == 't' && e[1] == 'c' && e[2] == '1'; }();
    return !one_tile;
}

int ef_match_tc_splits(int nq, int nt)
{
    const int qtiles = (nq + 127) / 128, ttiles = (nt + 127) / 128;
    int s = ef_div_up(148 * 4, qtiles);                      // one-tile kernel: about four waves of one CTA per SM
    s = std::max(1, std::min(std::min(s, ttiles), EF_MTC_MAX_LISTS / 2));
    return s;
}
int ef_match_tc_max_lists(void) { return EF_MTC_MAX_LISTS; }

// role_b: this set will be streamed as the train side
void ef_match_tc_expand(const uint8_t* d_desc, size_t pitch, int n, int desc_bytes, bool role_b, uint8_t* d_out, cudaStream_t s)
{
    const int rows = (n + 255) / 256 * 256;
    const long long threads = (long long)rows * (desc_bytes / 2);
    ef_match_expand_kernel<<<(unsigned)((threads + 255) / 256), 256, 0, s>>>(d_desc, pitch, n, desc_bytes, d_out, rows, (role_b && ef_match_tc_sliced()) ? 1 : 0);
    EF_COUNT_LAUNCH(1);
}

// d_partial: 2 * splits * nq int4.  texp must have been expanded with role_b = true.
void ef_match_tc_knn(const uint8_t* qexp, int nq, const uint8_t* texp, int nt, int desc_bytes, int k, int4* d_partial, int* d_idx, int* d_dist, cudaStream_t s)
{
    const int ttiles = (nt + 127) / 128;
    static unsigned long long configured = 0;                // function attributes are per device
    int dev = 0;
    cudaGetDevice(&dev);
    if (!((__atomic_load_n(&configured, __ATOMIC_RELAXED) >> (dev & 63)) & 1ull)) {
        cudaFuncSetAttribute(ef_match_tc_kernel<512>, cudaFuncAttributeMaxDynamicSharedMemorySize, 3 * 128 * 512);
        cudaFuncSetAttribute(ef_match_tc_kernel<256>, cudaFuncAttributeMaxDynamicSharedMemorySize, 3 * 128 * 256);
        cudaFuncSetAttribute(ef_match_tc2_kernel<512>, cudaFuncAttributeMaxDynamicSharedMemorySize, 256 * 512 + 4 * 16384);
        cudaFuncSetAttribute(ef_match_tc2_kernel<256>, cudaFuncAttributeMaxDynamicSharedMemorySize, 256 * 256 + 4 * 16384);
        __atomic_fetch_or(&configured, 1ull << (dev & 63), __ATOMIC_RELAXED);
    }
    if (ef_match_tc_sliced()) {
        const int qblocks = (nq + 255) / 256;
        // about four waves of one CTA per SM (finer splits fill the last wave better but reload the 128 KB A block more often: measured
        // 0.77 vs 0.79 ms at 512 bit, 0.56 vs 0.47 ms at 256 bit for 8 instead of 4 splits)
        const int splits = std::max(1, std::min(std::min(ef_div_up(148 * 4, qblocks), ttiles), EF_MTC_MAX_LISTS));
        const int tps = ef_div_up(ttiles, splits), nsplit = ef_div_up(ttiles, tps);
        const dim3 grid(qblocks, nsplit);
        if (desc_bytes == 64) ef_match_tc2_kernel<512><<<grid, EF_MTC2_THREADS, 256 * 512 + 4 * 16384, s>>>(qexp, nq, texp, nt, tps, d_partial);
        else ef_match_tc2_kernel<256><<<grid, EF_MTC2_THREADS, 256 * 256 + 4 * 16384, s>>>(qexp, nq, texp, nt, tps, d_partial);
        ef_match_merge_lex_kernel<<<ef_div_up(nq, 256), 256, 0, s>>>(d_partial, nq, nsplit, k, d_idx, d_dist);
    } else {
        const int splits = ef_match_tc_splits(nq, nt);
        const int tps = ef_div_up(ttiles, splits), nsplit = ef_div_up(ttiles, tps);
        const dim3 grid((nq + 127) / 128, nsplit);
        if (desc_bytes == 64) ef_match_tc_kernel<512><<<grid, EF_MTC_THREADS, 3 * 128 * 512, s>>>(qexp, nq, texp, nt, tps, d_partial);
        else ef_match_tc_kernel<256><<<grid, EF_MTC_THREADS, 3 * 128 * 256, s>>>(qexp, nq, texp, nt, tps, d_partial);
        ef_match_merge_lex_kernel<<<ef_div_up(nq, 256), 256, 0, s>>>(d_partial, nq, 2 * nsplit, k, d_idx, d_dist);
    }
    EF_COUNT_LAUNCH(2);
}
