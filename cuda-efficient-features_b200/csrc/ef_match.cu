// ef_match.cu -- brute-force Hamming matcher for the binary descriptors of the path (sm_100a): the step right after
// detectAndCompute in every caller of the reference:
//   cv::BFMatcher(NORM_HAMMING)->knnMatch(d1, d2, m, 2)          samples/sample_image_sequence.cpp:81,115-116
//   cv::BFMatcher::create(NORM_HAMMING, true)->match(d1, d2, m)  samples/sample_feature_matching.cpp:99-101
//   ratio (0.9) + cross-check filter over two knnMatch results   samples/sample_image_sequence.cpp:121-137
// so that 40k x 64 B descriptors never leave HBM.  Semantics pinned against OpenCV 4.13 (cv2, tests/test_matcher_cpu.py):
//   * k nearest = the k lexicographically smallest (distance, trainIdx) pairs (batchDistance scans train rows in order and
//     replaces on strict <);
//   * crossCheck: (i, j) is kept iff j is i's nearest train row AND i is j's nearest query row (both with that tie rule).
//
// Layout: one thread owns one query descriptor in registers (8 or 16 words); the train descriptors of the CTA's chunk stream
// through shared memory in tiles (every lane reads the same word: broadcast).  Per pair: W XORs, then the W population counts are
// folded through carry-save adders (LOP3 majority / 3-input XOR) so that only 5 POPCs per 512-bit pair hit the narrower POPC
// pipe.  The train set is cut into `splits` chunks (grid.y) for parallelism; partial top-2 lists are merged by a second kernel
// in (distance, index) order, which is exactly the sequential result.
#include "ef_common.cuh"

#include <algorithm>
#include <climits>
#include <cstdio>
#include <cstdlib>

#define EF_MT_THREADS 128
#define EF_MT_TILE 128          // train descriptors per shared-memory tile

__device__ __forceinline__ unsigned ef_xor3(unsigned a, unsigned b, unsigned c) { unsigned r; asm("lop3.b32 %0, %1, %2, %3, 0x96;" : "=r"(r) : "r"(a), "r"(b), "r"(c)); return r; }
__device__ __forceinline__ unsigned ef_maj3(unsigned a, unsigned b, unsigned c) { unsigned r; asm("lop3.b32 %0, %1, %2, %3, 0xE8;" : "=r"(r) : "r"(a), "r"(b), "r"(c)); return r; }

// popcount of the concatenation of W words through a carry-save adder tree
template <int W> __device__ __forceinline__ int ef_popc_words(const unsigned (&x)[W])
{
    if (W == 16) {
        // 16 -> ones/twos/fours/eights/sixteens (Harley-Seal): 11 CSAs, 5 POPCs
        const unsigned s0 = ef_xor3(x[0], x[1], x[2]), c0 = ef_maj3(x[0], x[1], x[2]);
        const unsigned s1 = ef_xor3(x[3], x[4], x[5]), c1 = ef_maj3(x[3], x[4], x[5]);
        const unsigned s2 = ef_xor3(x[6], x[7], x[8]), c2 = ef_maj3(x[6], x[7], x[8]);
        const unsigned s3 = ef_xor3(x[9], x[10], x[11]), c3 = ef_maj3(x[9], x[10], x[11]);
        const unsigned s4 = ef_xor3(x[12], x[13], x[14]), c4 = ef_maj3(x[12], x[13], x[14]);
        const unsigned s5 = ef_xor3(s0, s1, s2), c5 = ef_maj3(s0, s1, s2);
        const unsigned s6 = ef_xor3(s3, s4, x[15]), c6 = ef_maj3(s3, s4, x[15]);
        const unsigned ones = s5 ^ s6, c7 = s5 & s6;                 // weight 1 / carry into weight 2
        // weight 2: c0 c1 c2 c3 c4 c5 c6 c7 (8 words)
        const unsigned t0 = ef_xor3(c0, c1, c2), d0 = ef_maj3(c0, c1, c2);
        const unsigned t1 = ef_xor3(c3, c4, c5), d1 = ef_maj3(c3, c4, c5);
        const unsigned t2 = ef_xor3(c6, c7, t0), d2 = ef_maj3(c6, c7, t0);
        const unsigned twos = t1 ^ t2, d3 = t1 & t2;
        // weight 4: d0 d1 d2 d3
        const unsigned u0 = ef_xor3(d0, d1, d2), e0 = ef_maj3(d0, d1, d2);
        const unsigned fours = u0 ^ d3, e1 = u0 & d3;
        // weight 8: e0 e1
        const unsigned eights = e0 ^ e1, sixteens = e0 & e1;
        return __popc(ones) + 2 * __popc(twos) + 4 * __popc(fours) + 8 * __popc(eights) + 16 * __popc(sixteens);
    } else {
        // 8 words: 4 CSAs + 3 half adders, 4 POPCs
        const unsigned s0 = ef_xor3(x[0], x[1], x[2]), c0 = ef_maj3(x[0], x[1], x[2]);
        const unsigned s1 = ef_xor3(x[3], x[4], x[5]), c1 = ef_maj3(x[3], x[4], x[5]);
        const unsigned s2 = ef_xor3(x[6], x[7], s0), c2 = ef_maj3(x[6], x[7], s0);
        const unsigned ones = s1 ^ s2, c3 = s1 & s2;
        const unsigned t0 = ef_xor3(c0, c1, c2), d0 = ef_maj3(c0, c1, c2);
        const unsigned twos = t0 ^ c3, d1 = t0 & c3;
        const unsigned fours = d0 ^ d1, eights = d0 & d1;
        return __popc(ones) + 2 * __popc(twos) + 4 * __popc(fours) + 8 * __popc(eights);
    }
}

// partial[(split * nq + q) * 4 + {0,1,2,3}] = d0, i0, d1, i1 of query q over the train rows of `split`
template <int W>
__global__ void __launch_bounds__(EF_MT_THREADS) ef_match_knn2_kernel(const uint8_t* __restrict__ query, size_t qpitch, int nq,
                                                                     const uint8_t* __restrict__ train, size_t tpitch, int nt,
                                                                     int rows_per_split, int4* __restrict__ partial)
{
    __shared__ __align__(16) unsigned s_t[EF_MT_TILE][W];
    const int tid = threadIdx.x;
    const int q = blockIdx.x * EF_MT_THREADS + tid;
    const int t_begin = blockIdx.y * rows_per_split, t_end = min(nt, t_begin + rows_per_split);
    const bool aligned = ((reinterpret_cast<uintptr_t>(query) | qpitch | reinterpret_cast<uintptr_t>(train) | tpitch) & 15) == 0;

    unsigned qd[W];
    {
        const uint8_t* qp = query + (size_t)min(q, nq - 1) * qpitch;
        if (aligned) {
#pragma unroll
            for (int i = 0; i < W / 4; i++) {
                const uint4 v = reinterpret_cast<const uint4*>(qp)[i];
                qd[4 * i] = v.x; qd[4 * i + 1] = v.y; qd[4 * i + 2] = v.z; qd[4 * i + 3] = v.w;
            }
        } else {
#pragma unroll
            for (int i = 0; i < W; i++)
                qd[i] = qp[4 * i] | ((unsigned)qp[4 * i + 1] << 8) | ((unsigned)qp[4 * i + 2] << 16) | ((unsigned)qp[4 * i + 3] << 24);
        }
    }
    int d0 = INT_MAX, i0 = -1, d1 = INT_MAX, i1 = -1;
    for (int t0 = t_begin; t0 < t_end; t0 += EF_MT_TILE) {
        const int nrows = min(EF_MT_TILE, t_end - t0);
        __syncthreads();
        if (aligned) {
            for (int i = tid; i < nrows * (W / 4); i += EF_MT_THREADS) {
                const int r = i / (W / 4), c = i - r * (W / 4);
                reinterpret_cast<uint4*>(&s_t[r][0])[c] = __ldg(reinterpret_cast<const uint4*>(train + (size_t)(t0 + r) * tpitch) + c);
            }
        } else {
            for (int i = tid; i < nrows * W; i += EF_MT_THREADS) {
                const int r = i / W, c = i - r * W;
                const uint8_t* tp = train + (size_t)(t0 + r) * tpitch + 4 * c;
                s_t[r][c] = tp[0] | ((unsigned)tp[1] << 8) | ((unsigned)tp[2] << 16) | ((unsigned)tp[3] << 24);
            }
        }
        __syncthreads();
#pragma unroll 2
        for (int r = 0; r < nrows; r++) {
            unsigned x[W];
#pragma unroll
            for (int i = 0; i < W / 4; i++) {
                const uint4 v = reinterpret_cast<const uint4*>(&s_t[r][0])[i];
                x[4 * i] = v.x ^ qd[4 * i]; x[4 * i + 1] = v.y ^ qd[4 * i + 1]; x[4 * i + 2] = v.z ^ qd[4 * i + 2]; x[4 * i + 3] = v.w ^ qd[4 * i + 3];
            }
            const int d = ef_popc_words<W>(x);
            // rows arrive in increasing index: strict < keeps the earlier row on ties (OpenCV batchDistance)
            if (d < d1) {
                if (d < d0) { d1 = d0; i1 = i0; d0 = d; i0 = t0 + r; }
                else { d1 = d; i1 = t0 + r; }
            }
        }
    }
    if (q < nq) partial[(size_t)blockIdx.y * nq + q] = make_int4(d0, i0, d1, i1);
}

// merge the per-split lists in (distance, index) order; idx/dist: nq x 2 (k = 2) or nq x 1 (k = 1)
__global__ void __launch_bounds__(256) ef_match_merge_kernel(const int4* __restrict__ partial, int nq, int splits, int k, int* __restrict__ idx, int* __restrict__ dist)
{
    const int q = blockIdx.x * 256 + threadIdx.x;
    if (q >= nq) return;
    int d0 = INT_MAX, i0 = -1, d1 = INT_MAX, i1 = -1;
    for (int s = 0; s < splits; s++) {   // splits are in increasing train order: strict < keeps the earlier index
        const int4 p = partial[(size_t)s * nq + q];
        if (p.y >= 0 && p.x < d1) {
            if (p.x < d0) { d1 = d0; i1 = i0; d0 = p.x; i0 = p.y; }
            else { d1 = p.x; i1 = p.y; }
        }
        if (p.w >= 0 && p.z < d1) {
            if (p.z < d0) { d1 = d0; i1 = i0; d0 = p.z; i0 = p.w; }
            else { d1 = p.z; i1 = p.w; }
        }
    }
    if (k == 2) { idx[2 * q] = i0; idx[2 * q + 1] = i1; dist[2 * q] = d0; dist[2 * q + 1] = d1; }
    else { idx[q] = i0; dist[q] = d0; }
}

// crossCheck: keep (q, fwd[q]) iff bwd[fwd[q]] == q
__global__ void __launch_bounds__(256) ef_match_cross_kernel(const int* __restrict__ fwd_idx, const int* __restrict__ fwd_dist, const int* __restrict__ bwd_idx,
                                                             int nq, int* __restrict__ out_idx, int* __restrict__ out_dist)
{
    const int q = blockIdx.x * 256 + threadIdx.x;
    if (q >= nq) return;
    const int j = fwd_idx[q];
    const bool keep = j >= 0 && bwd_idx[j] == q;
    out_idx[q] = keep ? j : -1;
    out_dist[q] = keep ? fwd_dist[q] : INT_MAX;
}

// the filter of samples/sample_image_sequence.cpp:121-137 over knn12 (nq x 2) and knn21 (nt x 2):
//   keep q iff !(d12[0] > u * d12[1]) && !(d21[t][0] > u * d21[t][1]) && idx21[t][0] == q,  t = idx12[q][0]
// (DMatch::distance is a float, `uniqueness` a double: the comparison runs in double like the C++ expression)
__global__ void __launch_bounds__(256) ef_match_ratio_cross_kernel(const int* __restrict__ idx12, const int* __restrict__ dist12, int nq,
                                                                   const int* __restrict__ idx21, const int* __restrict__ dist21,
                                                                   double uniqueness, int* __restrict__ out_train)
{
    const int q = blockIdx.x * 256 + threadIdx.x;
    if (q >= nq) return;
    int r = -1;
    const int t = idx12[2 * q];
    if (t >= 0 && idx12[2 * q + 1] >= 0 && idx21[2 * t + 1] >= 0) {
        const bool u12 = (double)(float)dist12[2 * q] > uniqueness * (double)(float)dist12[2 * q + 1];
        const bool u21 = (double)(float)dist21[2 * t] > uniqueness * (double)(float)dist21[2 * t + 1];
        if (!u12 && !u21 && idx21[2 * t] == q) r = t;
    }
    out_train[q] = r;
}

// tcgen05 path (ef_match_tc.cu)
size_t ef_match_tc_expanded_bytes(int n, int desc_bytes);
int ef_match_tc_splits(int nq, int nt);
int ef_match_tc_max_lists(void);
void ef_match_tc_expand(const uint8_t* d_desc, size_t pitch, int n, int desc_bytes, bool role_b, uint8_t* d_out, cudaStream_t s);
void ef_match_tc_knn(const uint8_t* qexp, int nq, const uint8_t* texp, int nt, int desc_bytes, int k, int4* d_partial, int* d_idx, int* d_dist, cudaStream_t s);

namespace {
int match_splits(int nq, int nt)
{
    // enough CTAs for ~4 waves of 148 SMs x 4 resident CTAs, but at least one tile per split
    const int qblocks = ef_div_up(nq, EF_MT_THREADS);
    int s = ef_div_up(148 * 16, qblocks);
    s = std::min(s, ef_div_up(nt, EF_MT_TILE));
    return std::max(1, std::min(s, 256));
}
thread_local char g_match_err[256] = "";
int match_fail(int code, const char* msg) { std::snprintf(g_match_err, sizeof(g_match_err), "%s", msg); return code; }

// tensor-core path for problems that fill the machine; EF_MATCH=alu keeps the CUDA-core kernel (A/B comparison)
bool match_use_tc(int nq, int nt)
{
    static const bool env_alu = [] { const char* e = getenv("EF_MATCH"); return e && e[0] == 'a'; }();
    return !env_alu && nq >= 256 && nt >= 256;
}

// scratch layout: [partial lists][index/distance rows of the cross check][+-1 expansion of the query set][of the train set]
struct MatchScratch { size_t partial, rows, qexp, texp, qexp_b, texp_a, total; };   // *_b / *_a: the other role's operand order (cross check)
MatchScratch match_scratch(int nq, int nt)
{
    MatchScratch m{};
    const size_t mx = (size_t)std::max(std::max(nq, nt), 1);
    const size_t alu = (size_t)std::max(match_splits(nq, nt) * (size_t)nq, match_splits(nt, nq) * (size_t)nt) * sizeof(int4);
    const size_t tc = (size_t)ef_match_tc_max_lists() * mx * sizeof(int4);
    m.partial = 0;
    m.rows = ef_align_up(std::max(alu, tc), 256);
    m.qexp = m.rows + ef_align_up(4 * mx * sizeof(int), 256);
    m.texp = m.qexp + ef_match_tc_expanded_bytes(nq, 64);
    m.qexp_b = m.texp + ef_match_tc_expanded_bytes(nt, 64);
    m.texp_a = m.qexp_b + ef_match_tc_expanded_bytes(nq, 64);
    m.total = m.texp_a + ef_match_tc_expanded_bytes(nt, 64) + 256;
    return m;
}

int knn_launch_alu(const uint8_t* d_query, size_t qpitch, int nq, const uint8_t* d_train, size_t tpitch, int nt, int desc_bytes, int k,
                   int* d_idx, int* d_dist, void* d_partial, cudaStream_t s)
{
    const int splits = match_splits(nq, nt);
    const int rows = ef_div_up(ef_div_up(nt, splits), EF_MT_TILE) * EF_MT_TILE;
    const int nsplit = ef_div_up(nt, rows);
    int4* partial = reinterpret_cast<int4*>(d_partial);
    const dim3 grid(ef_div_up(nq, EF_MT_THREADS), nsplit);
    if (desc_bytes == 64) ef_match_knn2_kernel<16><<<grid, EF_MT_THREADS, 0, s>>>(d_query, qpitch, nq, d_train, tpitch, nt, rows, partial);
    else ef_match_knn2_kernel<8><<<grid, EF_MT_THREADS, 0, s>>>(d_query, qpitch, nq, d_train, tpitch, nt, rows, partial);
    ef_match_merge_kernel<<<ef_div_up(nq, 256), 256, 0, s>>>(partial, nq, nsplit, k, d_idx, d_dist);
    EF_COUNT_LAUNCH(2);
    return cudaGetLastError() == cudaSuccess ? EF_OK : match_fail(EF_ERR_CUDA, "kernel launch failed");
}
int match_check(const void* q, int nq, const void* t, int nt, int desc_bytes, const void* a, const void* b, const void* scratch)
{
    if (nq < 0 || nt < 0) return match_fail(EF_ERR_BAD_ARG, "negative row count");
    if (desc_bytes != 32 && desc_bytes != 64) return match_fail(EF_ERR_BAD_ARG, "descriptor size must be 32 or 64 bytes (descriptorSize() of the path)");
    if (nq > 0 && nt > 0 && (!q || !t || !a || !b || !scratch)) return match_fail(EF_ERR_BAD_ARG, "null pointer");
    if (scratch && (reinterpret_cast<uintptr_t>(scratch) & 15)) return match_fail(EF_ERR_BAD_ARG, "scratch must be 16-byte aligned");
    return EF_OK;
}
} // namespace

extern "C" {

const char* ef_match_last_error_string(void) { return g_match_err; }

size_t ef_match_scratch_bytes(int nq, int nt)
{
    if (nq <= 0 || nt <= 0) return 256;
    return match_scratch(nq, nt).total;
}

int ef_match_knn_async(const uint8_t* d_query, size_t qpitch, int nq, const uint8_t* d_train, size_t tpitch, int nt, int desc_bytes, int k,
                       int* d_idx, int* d_dist, void* d_scratch, void* stream)
{
    int rc = match_check(d_query, nq, d_train, nt, desc_bytes, d_idx, d_dist, d_scratch);
    if (rc != EF_OK) return rc;
    if (k != 1 && k != 2) return match_fail(EF_ERR_UNSUPPORTED, "k must be 1 or 2");
    cudaStream_t s = (cudaStream_t)stream;
    if (nq == 0) return EF_OK;
    if (nt == 0) {
        cudaMemsetAsync(d_idx, 0xff, sizeof(int) * (size_t)nq * k, s);
        cudaMemsetAsync(d_dist, 0x7f, sizeof(int) * (size_t)nq * k, s); // 0x7f7f7f7f: "no match" distances are never read
        return EF_OK;
    }
    if (qpitch < (size_t)desc_bytes || tpitch < (size_t)desc_bytes) return match_fail(EF_ERR_BAD_ARG, "pitch smaller than the descriptor size");
    uint8_t* sc = reinterpret_cast<uint8_t*>(d_scratch);
    if (!match_use_tc(nq, nt)) return knn_launch_alu(d_query, qpitch, nq, d_train, tpitch, nt, desc_bytes, k, d_idx, d_dist, sc, s);
    const MatchScratch m = match_scratch(nq, nt);
    ef_match_tc_expand(d_query, qpitch, nq, desc_bytes, false, sc + m.qexp, s);
    ef_match_tc_expand(d_train, tpitch, nt, desc_bytes, true, sc + m.texp, s);
    ef_match_tc_knn(sc + m.qexp, nq, sc + m.texp, nt, desc_bytes, k, reinterpret_cast<int4*>(sc + m.partial), d_idx, d_dist, s);
    return cudaGetLastError() == cudaSuccess ? EF_OK : match_fail(EF_ERR_CUDA, "kernel launch failed");
}

int ef_match_cross_check_async(const uint8_t* d_query, size_t qpitch, int nq, const uint8_t* d_train, size_t tpitch, int nt, int desc_bytes,
                               int* d_train_idx, int* d_dist, void* d_scratch, void* stream)
{
    int rc = match_check(d_query, nq, d_train, nt, desc_bytes, d_train_idx, d_dist, d_scratch);
    if (rc != EF_OK) return rc;
    cudaStream_t s = (cudaStream_t)stream;
    if (nq == 0) return EF_OK;
    if (nt == 0) { cudaMemsetAsync(d_train_idx, 0xff, sizeof(int) * (size_t)nq, s); cudaMemsetAsync(d_dist, 0x7f, sizeof(int) * (size_t)nq, s); return EF_OK; }
    if (qpitch < (size_t)desc_bytes || tpitch < (size_t)desc_bytes) return match_fail(EF_ERR_BAD_ARG, "pitch smaller than the descriptor size");
    uint8_t* sc = reinterpret_cast<uint8_t*>(d_scratch);
    const MatchScratch m = match_scratch(nq, nt);
    const size_t mx = (size_t)std::max(nq, nt);
    int* rows = reinterpret_cast<int*>(sc + m.rows);
    int *fwd_idx = rows, *fwd_dist = rows + mx, *bwd_idx = rows + 2 * mx, *bwd_dist = rows + 3 * mx;
    if (match_use_tc(nq, nt)) {
        ef_match_tc_expand(d_query, qpitch, nq, desc_bytes, false, sc + m.qexp, s);
        ef_match_tc_expand(d_train, tpitch, nt, desc_bytes, true, sc + m.texp, s);
        ef_match_tc_expand(d_query, qpitch, nq, desc_bytes, true, sc + m.qexp_b, s);
        ef_match_tc_expand(d_train, tpitch, nt, desc_bytes, false, sc + m.texp_a, s);
        ef_match_tc_knn(sc + m.qexp, nq, sc + m.texp, nt, desc_bytes, 1, reinterpret_cast<int4*>(sc + m.partial), fwd_idx, fwd_dist, s);
        ef_match_tc_knn(sc + m.texp_a, nt, sc + m.qexp_b, nq, desc_bytes, 1, reinterpret_cast<int4*>(sc + m.partial), bwd_idx, bwd_dist, s);
    } else {
        rc = knn_launch_alu(d_query, qpitch, nq, d_train, tpitch, nt, desc_bytes, 1, fwd_idx, fwd_dist, sc + m.partial, s);
        if (rc != EF_OK) return rc;
        rc = knn_launch_alu(d_train, tpitch, nt, d_query, qpitch, nq, desc_bytes, 1, bwd_idx, bwd_dist, sc + m.partial, s);
        if (rc != EF_OK) return rc;
    }
    ef_match_cross_kernel<<<ef_div_up(nq, 256), 256, 0, s>>>(fwd_idx, fwd_dist, bwd_idx, nq, d_train_idx, d_dist);
    EF_COUNT_LAUNCH(1);
    return cudaGetLastError() == cudaSuccess ? EF_OK : match_fail(EF_ERR_CUDA, "kernel launch failed");
}

int ef_match_ratio_cross_async(const int* d_idx12, const int* d_dist12, int nq, const int* d_idx21, const int* d_dist21, int nt,
                               double uniqueness, int* d_out_train, void* stream)
{
    if (nq < 0 || nt < 0) return match_fail(EF_ERR_BAD_ARG, "negative row count");
    if (nq == 0) return EF_OK;
    if (!d_idx12 || !d_dist12 || !d_out_train || (nt > 0 && (!d_idx21 || !d_dist21))) return match_fail(EF_ERR_BAD_ARG, "null pointer");
    ef_match_ratio_cross_kernel<<<ef_div_up(nq, 256), 256, 0, (cudaStream_t)stream>>>(d_idx12, d_dist12, nq, d_idx21, d_dist21, uniqueness, d_out_train);
    EF_COUNT_LAUNCH(1);
    return cudaGetLastError() == cudaSuccess ? EF_OK : match_fail(EF_ERR_CUDA, "kernel launch failed");
}

} // extern "C"
