// ef_api.cu -- C ABI (include/ef_b200.h): handle, workspace planner, level geometry, parameter
// tables, stage sequencing.  Host side of the hot path; mirrors the orchestration of
// EfficientFeaturesImpl::detectAndComputeAsync (src/cuda_efficient_features.cpp:225-321) without its
// two host synchronisations per level.
#include "ef_common.cuh"

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <new>
#include <string>
#include <vector>

#include "params/ef_bad_tables.inc"
#include "params/ef_hashsift_w256.inc"
#include "params/ef_hashsift_w512.inc"

unsigned long long g_ef_launches = 0;

// Tile rows (32-pixel rows of tiles) of one pyramid level owned by band `shard` of `nshards`, and the rows its score stage must
// cover (owned rows + `halo_tiles` on either side, clipped).  Pure host arithmetic: exported for the host mirrors and tests.
extern "C" void ef_band_tile_rows(int tiles_y, int shard, int nshards, int halo_tiles, int* own0, int* own_n, int* score0, int* score_n)
{
    if (nshards < 1) nshards = 1;
    if (shard < 0) shard = 0;
    if (shard >= nshards) shard = nshards - 1;
    const int a = (int)((long long)tiles_y * shard / nshards), b = (int)((long long)tiles_y * (shard + 1) / nshards);
    *own0 = a; *own_n = b - a;
    if (b - a <= 0) { *score0 = a; *score_n = 0; return; }
    const int sa = std::max(0, a - halo_tiles), sb = std::min(tiles_y, b + halo_tiles);
    *score0 = sa; *score_n = sb - sa;
}

// Output rows of the descriptor matrix that band `shard` of `nshards` fills (ef_band_finish_async): equal blocks of
// ceil(nfeatures / nshards) rows, so that an all-gather of fixed-size blocks assembles the matrix.  Pure host arithmetic.
extern "C" void ef_band_desc_rows(int nfeatures, int shard, int nshards, int* row0, int* nrows)
{
    if (nshards < 1) nshards = 1;
    if (shard < 0) shard = 0;
    if (shard >= nshards) shard = nshards - 1;
    const int c = (std::max(nfeatures, 0) + nshards - 1) / nshards;
    *row0 = shard * c; *nrows = c;
}

namespace {

struct Geometry {
    int nlevels = 0;
    int w[EF_MAX_LEVELS], h[EF_MAX_LEVELS];
    float scale[EF_MAX_LEVELS];
    int quota[EF_MAX_LEVELS];
};

// calcImagePyramid sizes (cuda_efficient_features.cpp:144-155) and calcNumFeaturesPerLevel (:159-174)
void compute_geometry(int w, int h, float scaleFactor, int nlevels, int nfeatures, Geometry& g)
{
    g.nlevels = nlevels;
    float scale = 1.f;
    g.w[0] = w; g.h[0] = h; g.scale[0] = scale;
    for (int s = 1; s < nlevels; s++) {
        scale *= scaleFactor;
        const float inv = 1.f / scale;
        g.h[s] = (int)lrintf(inv * (float)h); // cvRound
        g.w[s] = (int)lrintf(inv * (float)w);
        g.scale[s] = scale;
    }
    const double factor = (double)(1 / scaleFactor);
    double nf = nfeatures * (1 - factor) / (1 - std::pow(factor, nlevels));
    int sum = 0;
    for (int s = 0; s < nlevels - 1; s++) {
        g.quota[s] = (int)lrint(nf);
        sum += g.quota[s];
        nf *= factor;
    }
    g.quota[nlevels - 1] = std::max(nfeatures - sum, 0);
}

// Capacity of a level's survivor list.  The reference sizes it 0.1 * area (CORNER_DENSITY, cuda_efficient_features.cpp:35,252).
// NMS survivors are pairwise at least sqrt(r2) apart, i.e. at most ~1.155 / r2 of the pixels (hexagonal packing): below r = 4 that
// bound exceeds 10 %, so the list is sized from it (+ a boundary term) and can never overflow: results always equal the oracle's.
long surv_capacity(int w, int h, int r2)
{
    const double area = (double)w * h;
    const double dens = r2 <= 1 ? 1.0 : std::min(1.0, 1.3 / (double)r2);
    const double cap = std::max(0.1 * area, dens * area + 2.0 * (w + h));
    return std::max(1l, lrint(std::min(area, cap)));
}

int desc_bytes_of(int t) { return (t == EF_BAD_256 || t == EF_HASH_SIFT_256) ? 32 : 64; }
bool is_bad(int t) { return t == EF_BAD_256 || t == EF_BAD_512; }

struct LevelPlan { unsigned long long img_off, blur_off, resp_off, blk_off, mask_off, rowcnt_off, surv_off, sel_off; int img_pitch, resp_pitch; };

// radius NMS geometry (radiusSuppression, cuda_efficient_features.cu:291-292: imageRadius = ceil(r^2)) and the block edge b of
// the block-maximum map: the largest of 8, 4, 2 with 2(b-1)^2 < r^2, so that a block lies inside the disc of each of its pixels.
struct NmsGeom { int r2, R, block, K; };
NmsGeom nms_geometry(int radius)
{
    NmsGeom g;
    const float rf = (float)radius;
    g.r2 = (int)std::ceil(rf * rf);
    g.R = 0;
    while ((g.R + 1) * (g.R + 1) < g.r2) g.R++;
    g.block = 0;
    for (int b = 8; b >= 2; b >>= 1)
        if (2 * (b - 1) * (b - 1) < g.r2) { g.block = b; break; }
    if (g.r2 <= 1) g.block = 0; // the disc holds only the pixel itself: nothing is suppressed
    g.K = g.block ? (g.block - 1 + g.R) / g.block : 0;
    return g;
}

} // namespace

struct ef_handle {
    ef_params prm;
    std::string err;
    int device = 0;

    // workspace plan (for max_width x max_height)
    Geometry gmax;
    LevelPlan plan[EF_MAX_LEVELS];
    unsigned long long rowcnt_bytes = 0; // leading region of every frame slot, zeroed per call
    unsigned long long slot_bytes = 0;
    size_t total_bytes = 0;

    uint8_t* d_ws = nullptr;
    EfLevelCounters* d_counters = nullptr;
    int* d_slice_y = nullptr;       // [max_batch][EF_MAX_LEVELS][2]: row span of this GPU's descriptor slice per level (ef_band_finish_async)

    // descriptor tables (per handle, per device)
    uchar4* d_bad_boxes[2] = { nullptr, nullptr };
    unsigned char* d_bad_radius[2] = { nullptr, nullptr };
    float* d_bad_thr[2] = { nullptr, nullptr };
    float* d_hs_weights_t[2] = { nullptr, nullptr }; // 129 x nbits (transposed; fp64 fallback projection)
    uint8_t* d_hs_btc[2] = { nullptr, nullptr };     // the same digits in UMMA core-matrix order for tcgen05 (ef_project_tc.cu)
    uint4* d_hs_bfrag[2] = { nullptr, nullptr };     // fixed-point digits of the projection in mma fragment order (ef_project.cu)
    long long* d_hs_bias[2] = { nullptr, nullptr };  // column 0 of the projection as fixed-point integers
    int hs_shift[2] = { 0, 0 };                      // fixed-point scale 2^-shift; bfrag == nullptr: table does not fit 6 digits
    int hs_ndig[2] = { 0, 0 };                       // digits per weight in d_hs_btc (6 or 7); 0: no integer form (fp64 kernel)
    float* d_exp_table = nullptr;
    float2* d_grad_table = nullptr;

    uint8_t* d_sift128 = nullptr;   // max(max_batch*nfeatures, max_keypoints) x 128
    float* d_proj = nullptr;        // optional debug: rows x 512
    bool keep_proj = false;
    size_t sift_rows = 0;

    // compute-only scratch
    unsigned* d_integral = nullptr;
    unsigned* d_segsum = nullptr;
    float4* d_kpts4 = nullptr;

    // host-API staging
    uint8_t* d_in = nullptr; size_t in_pitch = 0, in_stride = 0;
    float* d_out_kpts = nullptr; size_t out_kpts_pitch = 0, out_kpts_stride = 0;
    uint8_t* d_out_desc = nullptr; size_t out_desc_stride = 0;
    int* d_out_counts = nullptr;
    int* h_counts_pinned = nullptr;
    // host API pipeline: upload / download streams and per-chunk events (upload of chunk c+1 and download of chunk c-1
    // overlap the kernels of chunk c)
    cudaStream_t s_in = nullptr, s_out = nullptr;
    std::vector<cudaEvent_t> ev_in, ev_cnt;
    // side stream of the device pipeline: the blur (needs only the pyramid) runs next to NMS / compaction / selection / angles
    cudaStream_t s_side = nullptr;
    cudaEvent_t ev_fork = nullptr, ev_join = nullptr;
    bool overlap_blur = false;  // measured: +0.3 % (5.181 vs 5.195 ms per 8 frames) -- the blur saturates the GPU by itself, the short kernels only queue behind it
    int slot0 = 0;              // first workspace slot of the call being enqueued (host pipeline: chunks on alternating streams use disjoint slots)
    int host_streams = 2;       // host-buffer pipeline: chunks alternate between the caller's stream and s_side (EF_B200_HOST_STREAMS=1: one stream)
    int host_chunk = 0;         // frames per chunk of the host-buffer pipeline; 0 = geometric plan 1, 2, 4, ... (measured: 10.43 vs 10.52 ms per 16 frames for chunks of 4)

    // optional per-stage timing (bench.py): events recorded between the stages
    bool timing = false;
    std::vector<cudaEvent_t> ev_pool;
    size_t ev_used = 0;      // events consumed since the last ef_stage_times()
    std::vector<int> ev_stage; // stage id that ENDS at event i (-1 = start marker)

    // tensor maps of the current geometry (re-encoded only when the caller's image, its layout or the frame size changes)
    EfTmaMaps tma;
    const void* tma_img0 = nullptr; const void* tma_ws = nullptr; size_t tma_stride = 0; int tma_pitch = 0, tma_w = 0, tma_h = 0, tma_nframes = 0;

    // last call
    int last_w = 0, last_h = 0, last_nframes = 0;
    const uint8_t* last_img0 = nullptr; size_t last_img0_stride = 0; int last_img0_pitch = 0;
    Geometry glast;
};

namespace {

int fail(ef_handle* h, int code, const std::string& msg)
{
    if (h) h->err = msg;
    return code;
}

#define EF_CUDA(h, expr)                                                                                         \
    do {                                                                                                         \
        cudaError_t e__ = (expr);                                                                                \
        if (e__ != cudaSuccess)                                                                                  \
            return fail(h, EF_ERR_CUDA, std::string(#expr) + ": " + cudaGetErrorString(e__));                    \
    } while (0)

// Entry points run on the handle's device and leave the caller's current device as they found it (a single-process multi-GPU
// caller -- PyTorch, ef_mg_* -- must not see its device change under its feet).
struct DeviceGuard {
    int prev = -1; bool switched = false, ok = true;
    explicit DeviceGuard(int dev)
    {
        if (cudaGetDevice(&prev) != cudaSuccess) { ok = false; return; }
        if (prev != dev) { ok = cudaSetDevice(dev) == cudaSuccess; switched = ok; }
    }
    ~DeviceGuard() { if (switched) cudaSetDevice(prev); }
    DeviceGuard(const DeviceGuard&) = delete;
    DeviceGuard& operator=(const DeviceGuard&) = delete;
};
#define EF_ON_DEVICE(h) DeviceGuard guard__((h)->device); if (!guard__.ok) return fail(h, EF_ERR_CUDA, "cudaSetDevice failed")

bool compute_only(const ef_handle* h) { return (h->prm.flags & EF_FLAG_COMPUTE_ONLY) != 0; }

// sum of the per-level quotas: what the path can deliver at most (the last level takes max(nfeatures - sum, 0), so the sum can
// exceed nfeatures by the rounding of the earlier levels, e.g. 8 for nfeatures = 7)
int quota_sum(const ef_params& p)
{
    Geometry g;
    compute_geometry(64, 64, p.scale_factor, p.nlevels, p.nfeatures, g);
    int s = 0;
    for (int l = 0; l < p.nlevels; l++) s += g.quota[l];
    return s;
}

void free_all(ef_handle* h)
{
    cudaFree(h->d_ws); cudaFree(h->d_counters); cudaFree(h->d_slice_y);
    for (int i = 0; i < 2; i++) { cudaFree(h->d_bad_boxes[i]); cudaFree(h->d_bad_radius[i]); cudaFree(h->d_bad_thr[i]); cudaFree(h->d_hs_weights_t[i]); cudaFree(h->d_hs_bfrag[i]); cudaFree(h->d_hs_btc[i]); cudaFree(h->d_hs_bias[i]); }
    cudaFree(h->d_exp_table); cudaFree(h->d_grad_table); cudaFree(h->d_sift128); cudaFree(h->d_proj);
    cudaFree(h->d_integral); cudaFree(h->d_segsum); cudaFree(h->d_kpts4);
    cudaFree(h->d_in); cudaFree(h->d_out_kpts); cudaFree(h->d_out_desc); cudaFree(h->d_out_counts);
    if (h->h_counts_pinned) cudaFreeHost(h->h_counts_pinned);
    if (h->s_in) cudaStreamDestroy(h->s_in);
    if (h->s_out) cudaStreamDestroy(h->s_out);
    if (h->s_side) cudaStreamDestroy(h->s_side);
    if (h->ev_fork) cudaEventDestroy(h->ev_fork);
    if (h->ev_join) cudaEventDestroy(h->ev_join);
    h->s_side = nullptr; h->ev_fork = h->ev_join = nullptr;
    for (cudaEvent_t e : h->ev_in) cudaEventDestroy(e);
    for (cudaEvent_t e : h->ev_cnt) cudaEventDestroy(e);
    h->ev_in.clear(); h->ev_cnt.clear(); h->s_in = h->s_out = nullptr; h->h_counts_pinned = nullptr;
    for (cudaEvent_t e : h->ev_pool) cudaEventDestroy(e);
    h->ev_pool.clear(); h->ev_stage.clear(); h->ev_used = 0;
}

int validate(const ef_params& p, std::string& why)
{
    if (p.nfeatures < 1) { why = "nfeatures must be >= 1"; return EF_ERR_BAD_ARG; }
    if (!(p.scale_factor > 1.f)) { why = "scale_factor must be > 1"; return EF_ERR_BAD_ARG; }
    if (p.nlevels < 1 || p.nlevels > EF_MAX_LEVELS) { why = "nlevels must be in [1,16]"; return EF_ERR_BAD_ARG; }
    if (p.first_level < 0 || p.first_level >= p.nlevels) { why = "first_level must be in [0,nlevels)"; return EF_ERR_BAD_ARG; }
    if (p.fast_threshold < 0 || p.fast_threshold > 255) { why = "fast_threshold must be in [0,255]"; return EF_ERR_BAD_ARG; }
    if (p.nonmax_radius < 0 || p.nonmax_radius > 64) { why = "nonmax_radius must be in [0,64]"; return EF_ERR_BAD_ARG; }
    if (p.desc_type < EF_BAD_256 || p.desc_type > EF_HASH_SIFT_512) { why = "unknown descriptor type"; return EF_ERR_BAD_ARG; }
    if (p.max_width < 32 || p.max_height < 32 || p.max_width > 32767 || p.max_height > 32767) { why = "max_width/max_height must be in [32,32767] (short2 locations)"; return EF_ERR_BAD_ARG; }
    if (p.max_batch < 1) { why = "max_batch must be >= 1"; return EF_ERR_BAD_ARG; }
    if (p.flags & ~EF_FLAG_COMPUTE_ONLY) { why = "unknown bits in flags"; return EF_ERR_BAD_ARG; }
    return EF_OK;
}

// plan the per-frame workspace slot for the largest frame
void plan_workspace(ef_handle* h)
{
    const ef_params& p = h->prm;
    compute_geometry(p.max_width, p.max_height, p.scale_factor, p.nlevels, p.nfeatures, h->gmax);
    unsigned long long off = 0;
    // rowcnt region first (zeroed with one 2-D memset per call)
    for (int l = 0; l < p.nlevels; l++) { h->plan[l].rowcnt_off = off; off += ef_align_up((unsigned long long)h->gmax.h[l] * 4, 128); }
    h->rowcnt_bytes = off;
    for (int l = 0; l < p.nlevels; l++) {
        LevelPlan& q = h->plan[l];
        const int w = h->gmax.w[l], hh = h->gmax.h[l];
        q.img_pitch = (int)ef_align_up(w, 128);
        q.resp_pitch = (int)ef_align_up(w, 32);
        const unsigned long long img_bytes = (unsigned long long)q.img_pitch * hh;
        q.img_off = off; if (l > 0) off += ef_align_up(img_bytes, 256);
        q.blur_off = off; off += ef_align_up(img_bytes, 256);
        q.resp_off = off; off += ef_align_up((unsigned long long)q.resp_pitch * hh * 4, 256);
        const int nb = nms_geometry(p.nonmax_radius).block;
        q.blk_off = off; if (nb) off += ef_align_up((unsigned long long)ef_div_up(w, nb) * ef_div_up(hh, nb) * sizeof(EfBlockMax), 256);
        const unsigned long long tiles = (unsigned long long)ef_div_up(w, EF_TILE) * ef_div_up(hh, EF_TILE);
        q.mask_off = off; off += ef_align_up(tiles * EF_TILE * 4, 256);
        const unsigned long long surv_cap = (unsigned long long)surv_capacity(w, hh, nms_geometry(p.nonmax_radius).r2);
        q.surv_off = off; off += ef_align_up(surv_cap * sizeof(EfSurvivor), 256);
        q.sel_off = off; off += ef_align_up((unsigned long long)p.nfeatures * sizeof(EfSelected), 256);
    }
    h->slot_bytes = ef_align_up(off, 4096);
}

int upload_tables(ef_handle* h)
{
    // BAD tables (bad.p256.h / bad.p512.h via tools/gen_params.py)
    for (int v = 0; v < 2; v++) {
        const int nbits = v == 0 ? 256 : 512;
        const unsigned char(*src)[5] = v == 0 ? ef_bad_boxes_256 : ef_bad_boxes_512;
        const unsigned int* thr = v == 0 ? ef_bad_thresholds_256_bits : ef_bad_thresholds_512_bits;
        std::vector<uchar4> boxes(nbits); std::vector<unsigned char> rad(nbits);
        for (int i = 0; i < nbits; i++) { boxes[i] = make_uchar4(src[i][0], src[i][1], src[i][2], src[i][3]); rad[i] = src[i][4]; }
        EF_CUDA(h, cudaMalloc(&h->d_bad_boxes[v], nbits * sizeof(uchar4)));
        EF_CUDA(h, cudaMalloc(&h->d_bad_radius[v], nbits));
        EF_CUDA(h, cudaMalloc(&h->d_bad_thr[v], nbits * sizeof(float)));
        EF_CUDA(h, cudaMemcpy(h->d_bad_boxes[v], boxes.data(), nbits * sizeof(uchar4), cudaMemcpyHostToDevice));
        EF_CUDA(h, cudaMemcpy(h->d_bad_radius[v], rad.data(), nbits, cudaMemcpyHostToDevice));
        EF_CUDA(h, cudaMemcpy(h->d_bad_thr[v], thr, nbits * sizeof(float), cudaMemcpyHostToDevice));
    }
    // HashSIFT projection, transposed to 129 x nbits so that one thread per output bit reads coalesced
    for (int v = 0; v < 2; v++) {
        const int nbits = v == 0 ? 256 : 512;
        const unsigned int* wb = v == 0 ? ef_hashsift_w256_bits : ef_hashsift_w512_bits;
        std::vector<unsigned int> wt((size_t)129 * nbits);
        for (int j = 0; j < nbits; j++)
            for (int k = 0; k < 129; k++) wt[(size_t)k * nbits + j] = wb[(size_t)j * 129 + k];
        EF_CUDA(h, cudaMalloc(&h->d_hs_weights_t[v], wt.size() * sizeof(float)));
        EF_CUDA(h, cudaMemcpy(h->d_hs_weights_t[v], wt.data(), wt.size() * sizeof(float), cudaMemcpyHostToDevice));
        // exact fixed-point form for the integer tensor cores (ef_project.cu): every fp32 weight is m * 2^(e-23);
        // S = -(smallest exponent of a lowest set bit) makes all of them integers W = w * 2^S, split into 6 balanced
        // base-256 digits.  Falls back to the fp64 kernel if some |W| needs more than 6 digits.
        {
            int min_lsb = 1 << 30;
            for (size_t i = 0; i < (size_t)nbits * 129; i++) {
                float f; std::memcpy(&f, &wb[i], 4);
                if (f == 0.f) continue;
                int e; const double m = std::frexp((double)f, &e);           // f = m * 2^e, 0.5 <= |m| < 1
                long long mant = (long long)std::ldexp(std::fabs(m), 24);   // 24-bit integer mantissa
                int tz = 0; while (!(mant & 1)) { mant >>= 1; tz++; }
                min_lsb = std::min(min_lsb, e - 24 + tz);
            }
            const int S = -min_lsb;
            bool ok = S > 0 && S < 100;
            std::vector<long long> W((size_t)nbits * 129);
            for (size_t i = 0; ok && i < W.size(); i++) {
                float f; std::memcpy(&f, &wb[i], 4);
                const double scaled = std::ldexp((double)f, S);             // exact: power-of-two scaling of a 24-bit mantissa
                if (std::fabs(scaled) >= 9.0e15) { ok = false; break; }
                W[i] = (long long)scaled;
                if ((double)W[i] != scaled) ok = false;
            }
            // six balanced base-256 digits when they suffice (512-bit table: 47-bit integers), seven otherwise (256-bit table: 49 bits)
            std::vector<signed char> dig; std::vector<long long> bias(nbits);
            int ndig = 0;
            for (int nd = 6; ok && nd <= 7 && !ndig; nd++) {
                bool fits = true;
                dig.assign((size_t)nd * nbits * 128, 0);
                for (int j = 0; fits && j < nbits; j++) {
                    bias[j] = W[(size_t)j * 129];
                    for (int k = 0; k < 128; k++) {
                        long long x = W[(size_t)j * 129 + 1 + k];
                        for (int d = 0; d < nd; d++) {
                            const long long r = ((x + 128) & 255) - 128;     // balanced digit in [-128, 127]
                            dig[((size_t)d * nbits + j) * 128 + k] = (signed char)r;
                            x = (x - r) / 256;
                        }
                        if (x != 0) { fits = false; break; }
                    }
                    if (std::llabs(bias[j]) >= (1ll << 61)) fits = false;
                }
                if (fits) ndig = nd;
            }
            ok = ok && ndig != 0;
            h->hs_ndig[v] = ok ? ndig : 0;
            if (ok && ndig == 6) {
                // fragment order: bfrag[ntile][digit][half][lane] = { b0(ks=2*half), b1(ks=2*half), b0(ks=2*half+1), b1(ks=2*half+1) }
                // with b0 = digits of output bit 8*ntile + lane/4 at k = 32*ks + 4*(lane%4) .. +3 and b1 the same at k + 16
                const int ntiles = nbits / 8;
                std::vector<unsigned int> frag((size_t)ntiles * 6 * 2 * 32 * 4);
                auto word = [&](int d, int j, int k) { unsigned int u; std::memcpy(&u, &dig[((size_t)d * nbits + j) * 128 + k], 4); return u; };
                for (int nt = 0; nt < ntiles; nt++)
                    for (int d = 0; d < 6; d++)
                        for (int half = 0; half < 2; half++)
                            for (int lane = 0; lane < 32; lane++) {
                                const int j = nt * 8 + lane / 4, kq = 4 * (lane % 4);
                                unsigned int* o = &frag[((((size_t)nt * 6 + d) * 2 + half) * 32 + lane) * 4];
                                o[0] = word(d, j, 32 * (2 * half) + kq);     o[1] = word(d, j, 32 * (2 * half) + kq + 16);
                                o[2] = word(d, j, 32 * (2 * half + 1) + kq); o[3] = word(d, j, 32 * (2 * half + 1) + kq + 16);
                            }
                EF_CUDA(h, cudaMalloc(&h->d_hs_bfrag[v], frag.size() * sizeof(unsigned int)));
                EF_CUDA(h, cudaMemcpy(h->d_hs_bfrag[v], frag.data(), frag.size() * sizeof(unsigned int), cudaMemcpyHostToDevice));
            }
            if (ok) {
                EF_CUDA(h, cudaMalloc(&h->d_hs_bias[v], bias.size() * sizeof(long long)));
                EF_CUDA(h, cudaMemcpy(h->d_hs_bias[v], bias.data(), bias.size() * sizeof(long long), cudaMemcpyHostToDevice));
                h->hs_shift[v] = S;
                // tcgen05 operand order: btc[chunk of 32 output bits][digit][4 KB], inside a 4 KB block the UMMA no-swizzle K-major
                // core-matrix layout: (n / 8) * 1024 + (k / 16) * 128 + (n % 8) * 16 + (k % 16)
                std::vector<signed char> btc((size_t)ndig * nbits * 128);
                for (int j = 0; j < nbits; j++)
                    for (int d = 0; d < ndig; d++)
                        for (int k = 0; k < 128; k++) {
                            const int c = j / 32, nl = j % 32;
                            btc[((size_t)c * ndig + d) * 4096 + (nl / 8) * 1024 + (k / 16) * 128 + (nl % 8) * 16 + (k % 16)] = dig[((size_t)d * nbits + j) * 128 + k];
                        }
                EF_CUDA(h, cudaMalloc(&h->d_hs_btc[v], btc.size()));
                EF_CUDA(h, cudaMemcpy(h->d_hs_btc[v], btc.data(), btc.size(), cudaMemcpyHostToDevice));
            }
        }
    }
    // Finite-domain libm tables of the CPU reference (constants, filled once per handle with the host's
    // libm -- the same libm the reference's CPU build links, which is what "bit-exact vs CPU" means):
    //   expf(distScale * ((x-15)^2 + (y-15)^2)) for the 30x30 gradient positions (hash_sift.cpp:220-224,247)
    //   for dy, dx in [-255, 255] (hash_sift.cpp:247-260): sqrtf(dx^2 + dy^2), and from ori = atan2f(dy, dx):
    //   ob = (8 / 2pi) * ori, bin = floor(ob) wrapped into [0,8), fraction = ob - floor(ob).  The 3-bit bin is packed
    //   into the sign bit of the sqrt (bit 2) and bits 31:30 of the fraction (< 1, so both are free).
    {
        std::vector<float> et(900);
        std::vector<float2> gt((size_t)511 * 511);
        const float kpScale = 1.f / 6;
        const float kpRadius = kpScale * 32.f * 0.5f;
        const float kernelSigma = 0.5f * 4 * 3.f * kpRadius;
        const float distScale = -1.f / (2 * kernelSigma * kernelSigma);
        const float cx = 0.5f * 30.f, cy = 0.5f * 30.f;
        for (int y = 0; y < 30; y++)
            for (int x = 0; x < 30; x++) {
                const float fx = (float)x - cx, fy = (float)y - cy;
                et[y * 30 + x] = expf(distScale * (fx * fx + fy * fy));
            }
        const float PI_2_0F = 6.28318548f;
        const float scaleO = 8 / PI_2_0F;
        for (int dyi = -255; dyi <= 255; dyi++)
            for (int dxi = -255; dxi <= 255; dxi++) {
                const float dx = (float)dxi, dy = (float)dyi;
                const float mag = sqrtf(dx * dx + dy * dy);
                const float ori = atan2f(dy, dx);
                const float ob = scaleO * ori;
                int oi = (int)floorf(ob);
                const float of = ob - (float)oi;
                if (oi < 0) oi += 8;
                if (oi >= 8) oi -= 8;
                unsigned mb, fb;
                std::memcpy(&mb, &mag, 4); std::memcpy(&fb, &of, 4);
                if ((fb >> 30) != 0 || oi < 0 || oi > 7) return fail(h, EF_ERR_UNSUPPORTED, "orientation fraction outside [0,1): host libm atan2f out of range");
                mb |= (unsigned)(oi >> 2) << 31;
                fb |= (unsigned)(oi & 3) << 30;
                float2 e;
                std::memcpy(&e.x, &mb, 4); std::memcpy(&e.y, &fb, 4);
                gt[(size_t)(dyi + 255) * 511 + (dxi + 255)] = e;
            }
        EF_CUDA(h, cudaMalloc(&h->d_exp_table, et.size() * sizeof(float)));
        EF_CUDA(h, cudaMalloc(&h->d_grad_table, gt.size() * sizeof(float2)));
        EF_CUDA(h, cudaMemcpy(h->d_exp_table, et.data(), et.size() * sizeof(float), cudaMemcpyHostToDevice));
        EF_CUDA(h, cudaMemcpy(h->d_grad_table, gt.data(), gt.size() * sizeof(float2), cudaMemcpyHostToDevice));
    }
    return EF_OK;
}

int allocate(ef_handle* h)
{
    const ef_params& p = h->prm;
    plan_workspace(h);
    size_t total = 0;
    auto alloc = [&](void** ptr, size_t bytes) -> cudaError_t { total += bytes; return cudaMalloc(ptr, bytes ? bytes : 1); };
    const bool conly = compute_only(h);
    if (!conly) {
        EF_CUDA(h, alloc((void**)&h->d_ws, (size_t)h->slot_bytes * p.max_batch));
        // row padding (columns w .. pitch) is read by the 16-byte window loads but never written: define it once
        EF_CUDA(h, cudaMemset(h->d_ws, 0, (size_t)h->slot_bytes * p.max_batch));
        EF_CUDA(h, alloc((void**)&h->d_counters, sizeof(EfLevelCounters) * EF_MAX_LEVELS * p.max_batch));
        EF_CUDA(h, alloc((void**)&h->d_slice_y, sizeof(int) * 2 * EF_MAX_LEVELS * p.max_batch));
    }
    h->sift_rows = conly ? (size_t)p.max_keypoints : std::max((size_t)p.max_batch * p.nfeatures, (size_t)p.max_keypoints);
    EF_CUDA(h, alloc((void**)&h->d_sift128, h->sift_rows * 128));
    EF_CUDA(h, alloc((void**)&h->d_integral, (size_t)(p.max_width + 1) * (p.max_height + 1) * 4));
    EF_CUDA(h, alloc((void**)&h->d_segsum, (size_t)(p.max_width + 1) * (ef_div_up(p.max_height, 64) + 1) * 4));
    EF_CUDA(h, alloc((void**)&h->d_kpts4, (size_t)p.max_keypoints * sizeof(float4)));
    EF_CUDA(h, cudaStreamCreateWithFlags(&h->s_side, cudaStreamNonBlocking));
    EF_CUDA(h, cudaEventCreateWithFlags(&h->ev_fork, cudaEventDisableTiming));
    EF_CUDA(h, cudaEventCreateWithFlags(&h->ev_join, cudaEventDisableTiming));
    if (conly) { h->total_bytes = total; return EF_OK; }
    // host-API staging
    h->in_pitch = ef_align_up(p.max_width, 128);
    h->in_stride = h->in_pitch * p.max_height;
    EF_CUDA(h, alloc((void**)&h->d_in, h->in_stride * p.max_batch));
    EF_CUDA(h, cudaMemset(h->d_in, 0, h->in_stride * p.max_batch));
    h->out_kpts_pitch = ef_align_up((size_t)p.nfeatures * 4, 128);
    h->out_kpts_stride = h->out_kpts_pitch * EF_ROWS_COUNT;
    EF_CUDA(h, alloc((void**)&h->d_out_kpts, h->out_kpts_stride * p.max_batch));
    h->out_desc_stride = (size_t)p.nfeatures * 64;
    EF_CUDA(h, alloc((void**)&h->d_out_desc, h->out_desc_stride * p.max_batch));
    EF_CUDA(h, alloc((void**)&h->d_out_counts, sizeof(int) * p.max_batch));
    EF_CUDA(h, cudaMallocHost((void**)&h->h_counts_pinned, sizeof(int) * (p.max_batch + EF_MAX_LEVELS * 4)));
    EF_CUDA(h, cudaStreamCreateWithFlags(&h->s_in, cudaStreamNonBlocking));
    EF_CUDA(h, cudaStreamCreateWithFlags(&h->s_out, cudaStreamNonBlocking));
    if (const char* e = std::getenv("EF_B200_OVERLAP_BLUR")) h->overlap_blur = std::atoi(e) != 0;   // 1: blur on the side stream (A/B switch)
    if (const char* e = std::getenv("EF_B200_HOST_STREAMS")) h->host_streams = std::atoi(e) == 1 ? 1 : 2;
    if (const char* e = std::getenv("EF_B200_HOST_CHUNK")) h->host_chunk = std::max(0, std::atoi(e)); // frames per pipeline chunk of the host API (0: geometric plan)
    h->ev_in.resize(p.max_batch); h->ev_cnt.resize(p.max_batch);
    for (int i = 0; i < p.max_batch; i++) {
        EF_CUDA(h, cudaEventCreateWithFlags(&h->ev_in[i], cudaEventDisableTiming));
        EF_CUDA(h, cudaEventCreateWithFlags(&h->ev_cnt[i], cudaEventDisableTiming));
    }
    h->total_bytes = total;
    return EF_OK;
}

// fill the kernel parameter block for an actual call
// shard_i / shard_n: the band of tile rows this GPU owns when one frame is cut over several GPUs (ef_band_*); 0 / 1 = whole frame
int build_pipe(ef_handle* h, int nframes, int w, int hh, EfPipe& P, int shard_i = 0, int shard_n = 1)
{
    const ef_params& p = h->prm;
    if (compute_only(h)) return fail(h, EF_ERR_UNSUPPORTED, "the handle was created with EF_FLAG_COMPUTE_ONLY: no detection workspace");
    if (w > p.max_width || hh > p.max_height) return fail(h, EF_ERR_CAPACITY, "image larger than max_width x max_height of the handle");
    if (nframes < 1 || nframes > p.max_batch) return fail(h, EF_ERR_CAPACITY, "nframes exceeds max_batch of the handle");
    if (w < 32 || hh < 32) return fail(h, EF_ERR_BAD_ARG, "image smaller than 32x32");
    Geometry& g = h->glast;
    compute_geometry(w, hh, p.scale_factor, p.nlevels, p.nfeatures, g);
    std::memset(&P, 0, sizeof(P));
    P.nlevels = p.nlevels; P.first_level = p.first_level; P.nframes = nframes;
    const NmsGeom ng = nms_geometry(p.nonmax_radius);
    P.fast_threshold = p.fast_threshold; P.nms_r2 = ng.r2; P.nms_R = ng.R; P.nms_block = ng.block; P.nms_K = ng.K;
    P.nfeatures = p.nfeatures;
    P.shard_i = 0; P.shard_n = 1; P.select_from_counters = 0;
    P.desc_row0 = 0; P.desc_rows = 0x7fffffff; P.blur_by_slice = 0; P.slice_y = h->d_slice_y;
    P.desc_type = p.desc_type; P.desc_bytes = desc_bytes_of(p.desc_type);
    if (h->slot0 < 0 || h->slot0 + nframes > p.max_batch) return fail(h, EF_ERR_CAPACITY, "internal: workspace slots");
    P.ws = h->d_ws + (size_t)h->slot0 * h->slot_bytes; P.ws_stride = h->slot_bytes;
    P.counters = h->d_counters + (size_t)h->slot0 * EF_MAX_LEVELS;
    int tiles = 0, btiles = 0, bands = 0, kblocks = 0, sblocks = 0, strips = 0;
    for (int l = 0; l < p.nlevels; l++) {
        EfLevel& L = P.lv[l];
        const LevelPlan& q = h->plan[l];
        L.w = g.w[l]; L.h = g.h[l];
        if (L.w < 1 || L.h < 1) return fail(h, EF_ERR_BAD_ARG, "pyramid level degenerates to zero size; reduce nlevels");
        L.img_pitch = q.img_pitch; L.blur_pitch = q.img_pitch; L.resp_pitch = q.resp_pitch;
        L.tiles_x = ef_div_up(L.w, EF_TILE); L.tiles_y = ef_div_up(L.h, EF_TILE);
        ef_band_tile_rows(L.tiles_y, shard_i, shard_n, ef_div_up(ng.K * ng.block, EF_TILE), &L.own_ty0, &L.own_rows, &L.score_ty0, &L.score_rows);
        L.blk_w = ng.block ? ef_div_up(L.w, ng.block) : 0; L.blk_h = ng.block ? ef_div_up(L.h, ng.block) : 0;
        L.blur_tiles_x = ef_div_up(L.w, 64);
        L.blur_ty0 = 0; L.blur_rows = ef_div_up(L.h, 64);
        if (shard_n > 1) {
            // band-sharded frame: only the rows the descriptor windows of the owned keypoints can touch (k.y +- 24)
            const int y_lo = std::max(0, L.own_ty0 * EF_TILE - 24), y_hi = std::min(L.h, (L.own_ty0 + L.own_rows) * EF_TILE + 24);
            L.blur_ty0 = y_lo / 64;
            L.blur_rows = L.own_rows > 0 ? ef_div_up(y_hi, 64) - L.blur_ty0 : 0;
        }
        L.quota = g.quota[l];
        L.surv_cap = (int)surv_capacity(L.w, L.h, ng.r2);
        L.scale = g.scale[l];
        if (l > 0) {
            L.rx = (float)(1.0 / ((double)L.w / g.w[l - 1]));
            L.ry = (float)(1.0 / ((double)L.h / g.h[l - 1]));
        }
        L.img_off = q.img_off; L.blur_off = q.blur_off; L.resp_off = q.resp_off; L.blk_off = q.blk_off; L.mask_off = q.mask_off;
        L.rowcnt_off = q.rowcnt_off; L.surv_off = q.surv_off; L.sel_off = q.sel_off;
        L.tile_start = tiles; L.blur_tile_start = btiles; L.band_start = bands; L.kpt_block_start = kblocks; L.sift_block_start = sblocks;
        L.strips_x = ef_div_up(L.tiles_x, 4); L.strip_start = strips;
        L.tiles_x_inv = 0xffffffffu / (unsigned)L.tiles_x; L.strips_x_inv = 0xffffffffu / (unsigned)L.strips_x; L.blur_tiles_x_inv = 0xffffffffu / (unsigned)L.blur_tiles_x;
        if (l >= p.first_level) {
            tiles += L.tiles_x * L.score_rows;
            strips += L.strips_x * L.own_rows;
            bands += L.own_rows;
            kblocks += ef_div_up(std::min(L.quota, p.nfeatures), 8);
            sblocks += ef_div_up(std::min(L.quota, p.nfeatures), 4);
            btiles += L.blur_tiles_x * L.blur_rows;
        }
    }
    P.total_tiles = tiles; P.total_blur_tiles = btiles; P.total_bands = bands; P.total_kpt_blocks = kblocks; P.total_sift_blocks = sblocks; P.total_strips = strips;
    h->last_w = w; h->last_h = hh; h->last_nframes = nframes;
    return EF_OK;
}

// tensor maps over the level images of this call (ef_tma.cuh); cheap, and cached while the geometry stays the same
const EfTmaMaps* prepare_tma(ef_handle* h, const EfPipe& P)
{
    if (h->tma_img0 == P.img0 && h->tma_stride == P.img0_stride && h->tma_pitch == P.img0_pitch && h->tma_w == P.lv[0].w && h->tma_h == P.lv[0].h &&
        h->tma_nframes == P.nframes && h->tma_ws == (const void*)P.ws) return &h->tma;
    h->tma.blur_src_ok = 0; h->tma.resize_src_ok = 0;
    for (int l = 1; l < P.nlevels; l++) {
        // source of pyramid level l = level l-1.  The TMA kernel zero-fills what lies outside the source instead of clamping the
        // x2 / y2 taps (resize arithmetic: SURVEY Appendix A.1); that is exact iff no tap of the level ever clamps, i.e. iff
        // floor((w-1) * rx) + 1 <= src_w - 1 and the same for rows -- true for every down-scaling ratio, checked here in the
        // kernel's own float arithmetic.
        const EfLevel& L = P.lv[l];
        const EfLevel& S = P.lv[l - 1];
        const bool noclamp = (int)std::floor((float)(L.w - 1) * L.rx) + 1 <= S.w - 1 && (int)std::floor((float)(L.h - 1) * L.ry) + 1 <= S.h - 1;
        const void* base = l == 1 ? (const void*)P.img0 : (const void*)(P.ws + S.img_off);
        const size_t pitch = l == 1 ? (size_t)P.img0_pitch : (size_t)S.img_pitch, stride = l == 1 ? (size_t)P.img0_stride : (size_t)P.ws_stride;
        if (noclamp && ef_tma_encode_u8(&h->tma.resize_src[l], base, S.w, S.h, P.nframes, pitch, stride, EF_RS_BOX_W, EF_RS_BOX_H)) h->tma.resize_src_ok |= 1u << l;
    }
    for (int l = P.first_level; l < P.nlevels; l++) {
        const EfLevel& L = P.lv[l];
        const void* base = l == 0 ? (const void*)P.img0 : (const void*)(P.ws + L.img_off);
        const size_t pitch = l == 0 ? (size_t)P.img0_pitch : (size_t)L.img_pitch, stride = l == 0 ? (size_t)P.img0_stride : (size_t)P.ws_stride;
        if (ef_tma_encode_u8(&h->tma.blur_src[l], base, L.w, L.h, P.nframes, pitch, stride, EF_BLUR_BOX_W, EF_BLUR_BOX_H)) h->tma.blur_src_ok |= 1u << l;
    }
    h->tma_img0 = P.img0; h->tma_ws = P.ws; h->tma_stride = P.img0_stride; h->tma_pitch = P.img0_pitch; h->tma_w = P.lv[0].w; h->tma_h = P.lv[0].h; h->tma_nframes = P.nframes;
    return &h->tma;
}

void mark(ef_handle* h, int stage, cudaStream_t s)
{
    if (!h->timing) return;
    if (h->ev_used == h->ev_pool.size()) {
        cudaEvent_t e;
        if (cudaEventCreate(&e) != cudaSuccess) return;
        h->ev_pool.push_back(e);
        h->ev_stage.push_back(-1);
    }
    h->ev_stage[h->ev_used] = stage;
    cudaEventRecord(h->ev_pool[h->ev_used++], s);
}

int run_pipeline(ef_handle* h, EfPipe& P, bool want_desc, cudaStream_t s)
{
    EF_CUDA(h, cudaMemset2DAsync(P.ws, h->slot_bytes, 0, h->rowcnt_bytes, P.nframes, s));
    EF_CUDA(h, cudaMemsetAsync(P.counters, 0, sizeof(EfLevelCounters) * EF_MAX_LEVELS * P.nframes, s));
    uint8_t* const sift128 = h->d_sift128 + (size_t)h->slot0 * P.nfeatures * 128;
    float* const proj = h->keep_proj ? h->d_proj + (size_t)h->slot0 * P.nfeatures * 512 : nullptr;
    mark(h, -1, s);
    ef_launch_pyramid(P, prepare_tma(h, P), s);      mark(h, EF_STAGE_PYRAMID, s);
    ef_launch_score(P, s);        mark(h, EF_STAGE_SCORE, s);
    // The blur depends on the pyramid only.  It is forked onto the side stream after the score stage (which saturates the GPU on its
    // own) and runs next to NMS, compaction, selection and angles -- latency-bound launches that leave issue slots free; the
    // descriptor stage joins it.  Markers with a negative stage id restart the elapsed-time chain of ef_stage_times().
    const bool side = want_desc && h->overlap_blur && h->s_side;
    if (side) {
        EF_CUDA(h, cudaEventRecord(h->ev_fork, s));
        EF_CUDA(h, cudaStreamWaitEvent(h->s_side, h->ev_fork, 0));
        mark(h, -2, h->s_side);
        ef_launch_blur(P, prepare_tma(h, P), h->s_side); mark(h, EF_STAGE_BLUR, h->s_side);
        EF_CUDA(h, cudaEventRecord(h->ev_join, h->s_side));
        mark(h, -2, s);
    }
    ef_launch_nms(P, s);          mark(h, EF_STAGE_NMS, s);
    ef_launch_compact(P, s);      mark(h, EF_STAGE_COMPACT, s);
    ef_launch_select(P, s);       mark(h, EF_STAGE_SELECT, s);
    ef_launch_angle_pack(P, s);   mark(h, EF_STAGE_ANGLE_PACK, s);
    if (want_desc) {
        if (side) { EF_CUDA(h, cudaStreamWaitEvent(s, h->ev_join, 0)); mark(h, -2, s); }
        else { ef_launch_blur(P, prepare_tma(h, P), s);     mark(h, EF_STAGE_BLUR, s); }
        const int v = (P.desc_bytes == 32) ? 0 : 1;
        if (is_bad(P.desc_type)) {
            EfBadTables t{ h->d_bad_boxes[v], h->d_bad_radius[v], h->d_bad_thr[v] };
            ef_launch_bad_pipe(P, t, s);
            mark(h, EF_STAGE_DESCRIBE, s);
        } else {
            EfHashSiftTables t{ h->d_exp_table, h->d_grad_table };
            ef_launch_hashsift_features_pipe(P, t, sift128, s);
            mark(h, EF_STAGE_DESCRIBE, s);
            const EfProjTables pt{ h->d_hs_bfrag[v], h->d_hs_bias[v], h->hs_shift[v], h->d_hs_weights_t[v], h->d_hs_btc[v], h->hs_ndig[v] };
            ef_launch_hashsift_project_batch(sift128, P.nfeatures, P.counts, P.nframes, pt, P.desc_bytes * 8,
                                             P.desc, (size_t)P.desc_stride, P.desc_pitch, proj, s);
            mark(h, EF_STAGE_PROJECT, s);
        }
    }
    EF_CUDA(h, cudaGetLastError());
    return EF_OK;
}

} // namespace

extern "C" {

const char* ef_version(void) { return "ef_b200 0.1 (sm_100a)"; }

void ef_default_params(ef_params* p)
{
    // defaults of EfficientFeatures::create, include/cuda_efficient_features.h:47-48
    p->nfeatures = 5000; p->scale_factor = 1.2f; p->nlevels = 8; p->first_level = 0; p->fast_threshold = 20;
    p->nonmax_radius = 15; p->desc_type = EF_HASH_SIFT_256; p->desc_scale = 1.f;
    p->max_width = 3840; p->max_height = 2160; p->max_batch = 1; p->max_keypoints = 0; p->device = 0; p->flags = 0;
}

int ef_create(const ef_params* params, ef_handle** out)
{
    if (!params || !out) return EF_ERR_BAD_ARG;
    *out = nullptr;
    ef_handle* h = new (std::nothrow) ef_handle();
    if (!h) return EF_ERR_CUDA;
    h->prm = *params;
    if (h->prm.max_keypoints < h->prm.nfeatures) h->prm.max_keypoints = h->prm.nfeatures;
    if (!(h->prm.desc_scale > 0.f)) h->prm.desc_scale = 1.f;
    std::string why;
    int rc = validate(h->prm, why);
    if (rc != EF_OK) { std::fprintf(stderr, "ef_create: %s\n", why.c_str()); delete h; return rc; }
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev <= 0 || params->device >= ndev) {
        // no CPU fallback: the path only exists as sm_100a kernels
        std::fprintf(stderr, "ef_create: no usable CUDA device (%s)\n", cudaGetErrorString(cudaGetLastError()));
        delete h; return EF_ERR_CUDA;
    }
    h->device = params->device;
    DeviceGuard guard(h->device);
    if (!guard.ok) { delete h; return EF_ERR_CUDA; }
    rc = allocate(h);
    if (rc == EF_OK) rc = upload_tables(h);
    if (rc != EF_OK) { std::fprintf(stderr, "ef_create: %s\n", h->err.c_str()); free_all(h); delete h; return rc; }
    *out = h;
    return EF_OK;
}

void ef_destroy(ef_handle* h)
{
    if (!h) return;
    DeviceGuard guard(h->device);
    free_all(h);
    delete h;
}

int ef_set_param(ef_handle* h, int id, double value)
{
    if (!h) return EF_ERR_BAD_ARG;
    ef_params q = h->prm;
    switch (id) {
    case EF_PARAM_MAX_FEATURES: q.nfeatures = (int)value; break;
    case EF_PARAM_SCALE_FACTOR: q.scale_factor = (float)value; break;
    case EF_PARAM_NLEVELS: q.nlevels = (int)value; break;
    case EF_PARAM_FIRST_LEVEL: q.first_level = (int)value; break;
    case EF_PARAM_FAST_THRESHOLD: q.fast_threshold = (int)value; break;
    case EF_PARAM_NONMAX_RADIUS: q.nonmax_radius = (int)value; break;
    case EF_PARAM_DESCRIPTOR_TYPE: q.desc_type = (int)value; break;
    case EF_PARAM_DESC_SCALE: q.desc_scale = (float)value; break;
    default: return fail(h, EF_ERR_BAD_ARG, "unknown parameter id");
    }
    std::string why;
    int rc = validate(q, why);
    if (rc != EF_OK) return fail(h, rc, why);
    const bool replan = q.nfeatures > h->prm.nfeatures || q.nlevels != h->prm.nlevels || q.scale_factor != h->prm.scale_factor ||
                        q.nonmax_radius != h->prm.nonmax_radius; // block-map and survivor-list capacities depend on the radius
    if (q.max_keypoints < q.nfeatures) q.max_keypoints = q.nfeatures;
    if (!replan) { h->prm = q; return EF_OK; }
    // capacities changed (setters are not on the hot path): build the new resources FIRST and swap them in only when everything
    // succeeded, so that a failure (e.g. out of memory after a large setMaxFeatures) leaves the handle exactly as it was
    EF_ON_DEVICE(h);
    cudaDeviceSynchronize();
    ef_handle* n = new (std::nothrow) ef_handle();
    if (!n) return fail(h, EF_ERR_CUDA, "out of host memory");
    n->prm = q; n->device = h->device;
    rc = allocate(n);
    if (rc == EF_OK) rc = upload_tables(n);
    if (rc == EF_OK && h->keep_proj) rc = ef_debug_keep_projection(n, 1);
    if (rc != EF_OK) {
        const std::string why2 = "re-planning the workspace failed, parameters unchanged: " + n->err;
        free_all(n); delete n;
        return fail(h, rc, why2);
    }
    n->timing = h->timing; n->keep_proj = h->keep_proj; n->overlap_blur = h->overlap_blur; n->host_chunk = h->host_chunk;
    free_all(h);
    *h = *n;          // plain members and resource pointers; `n` owns nothing afterwards
    delete n;
    return EF_OK;
}

int ef_get_param(const ef_handle* h, int id, double* value)
{
    if (!h || !value) return EF_ERR_BAD_ARG;
    switch (id) {
    case EF_PARAM_MAX_FEATURES: *value = h->prm.nfeatures; break;
    case EF_PARAM_SCALE_FACTOR: *value = h->prm.scale_factor; break;
    case EF_PARAM_NLEVELS: *value = h->prm.nlevels; break;
    case EF_PARAM_FIRST_LEVEL: *value = h->prm.first_level; break;
    case EF_PARAM_FAST_THRESHOLD: *value = h->prm.fast_threshold; break;
    case EF_PARAM_NONMAX_RADIUS: *value = h->prm.nonmax_radius; break;
    case EF_PARAM_DESCRIPTOR_TYPE: *value = h->prm.desc_type; break;
    case EF_PARAM_DESC_SCALE: *value = h->prm.desc_scale; break;
    default: return EF_ERR_BAD_ARG;
    }
    return EF_OK;
}

size_t ef_workspace_bytes(const ef_handle* h) { return h ? h->total_bytes : 0; }
int ef_descriptor_size(const ef_handle* h) { return h ? desc_bytes_of(h->prm.desc_type) : 0; }
const char* ef_last_error_string(const ef_handle* h) { return h ? h->err.c_str() : "null handle"; }

int ef_detect_and_compute_batch_async(ef_handle* h, int nframes, const uint8_t* d_imgs, size_t img_stride, size_t pitch,
                                      int width, int height, float* d_kpts, size_t kpts_stride, size_t kpts_pitch,
                                      uint8_t* d_desc, size_t desc_stride, size_t desc_pitch, int* d_counts, void* stream)
{
    if (!h) return EF_ERR_BAD_ARG;
    if (!d_imgs || !d_kpts || !d_counts) return fail(h, EF_ERR_BAD_ARG, "null image / keypoint / count pointer");
    if (pitch < (size_t)width) return fail(h, EF_ERR_BAD_ARG, "pitch smaller than width");
    if (kpts_pitch < (size_t)h->prm.nfeatures * 4 || (kpts_pitch & 3)) return fail(h, EF_ERR_BAD_ARG, "kpts_pitch must be >= 4*nfeatures and a multiple of 4");
    if (d_desc && desc_pitch < (size_t)desc_bytes_of(h->prm.desc_type)) return fail(h, EF_ERR_BAD_ARG, "desc_pitch smaller than the descriptor size");
    EF_ON_DEVICE(h);
    EfPipe P;
    int rc = build_pipe(h, nframes, width, height, P);
    if (rc != EF_OK) return rc;
    P.img0 = d_imgs; P.img0_stride = img_stride; P.img0_pitch = (int)pitch;
    P.kpts = d_kpts; P.kpts_stride = kpts_stride; P.kpts_pitch = (int)kpts_pitch;
    P.desc = d_desc; P.desc_stride = desc_stride; P.desc_pitch = (int)desc_pitch;
    P.counts = d_counts;
    h->last_img0 = d_imgs; h->last_img0_stride = img_stride; h->last_img0_pitch = (int)pitch;
    return run_pipeline(h, P, d_desc != nullptr, (cudaStream_t)stream);
}

int ef_detect_and_compute_async(ef_handle* h, const uint8_t* d_img, size_t pitch, int width, int height,
                                float* d_kpts, size_t kpts_pitch, uint8_t* d_desc, size_t desc_pitch, int* d_count, void* stream)
{
    return ef_detect_and_compute_batch_async(h, 1, d_img, 0, pitch, width, height, d_kpts, 0, kpts_pitch, d_desc, 0, desc_pitch, d_count, stream);
}

// rows_path: keypoints come from a 5 x N matrix (integer coordinates, size forced to 31)
static int compute_common(ef_handle* h, const uint8_t* d_img, size_t pitch, int width, int height, const float4* d_k4, int n,
                          uint8_t* d_desc, size_t desc_pitch, cudaStream_t s, bool rows_path = false)
{
    const ef_params& p = h->prm;
    EfDescJob job;
    job.img = d_img; job.w = width; job.h = height; job.pitch = (int)pitch;
    job.kpts = d_k4; job.n = n; job.scale = p.desc_scale;
    job.desc = d_desc; job.desc_pitch = (int)desc_pitch; job.nbits = desc_bytes_of(p.desc_type) * 8;
    // integer keypoints of size 31 at scale 1 on a 16-byte aligned image: the window-staging kernels of the pipeline apply
    job.staged31 = (rows_path && p.desc_scale == 1.f && ((reinterpret_cast<uintptr_t>(d_img) | pitch) & 15) == 0) ? 1 : 0;
    const int v = job.nbits == 256 ? 0 : 1;
    if (is_bad(p.desc_type)) {
        if (!job.staged31) ef_launch_integral(d_img, width, height, (int)pitch, h->d_integral, h->d_segsum, s);
        EfBadTables t{ h->d_bad_boxes[v], h->d_bad_radius[v], h->d_bad_thr[v] };
        ef_launch_bad_flat(job, h->d_integral, t, s);
    } else {
        EfHashSiftTables t{ h->d_exp_table, h->d_grad_table };
        ef_launch_hashsift_features_flat(job, t, h->d_sift128, s);
        const EfProjTables pt{ h->d_hs_bfrag[v], h->d_hs_bias[v], h->hs_shift[v], h->d_hs_weights_t[v], h->d_hs_btc[v], h->hs_ndig[v] };
        ef_launch_hashsift_project_batch(h->d_sift128, n, nullptr, 1, pt, job.nbits, d_desc, 0, (int)desc_pitch,
                                         h->keep_proj ? h->d_proj : nullptr, s);
    }
    EF_CUDA(h, cudaGetLastError());
    return EF_OK;
}

static int compute_check(ef_handle* h, const uint8_t* d_img, size_t pitch, int width, int height, const void* kp, int n, uint8_t* d_desc, size_t desc_pitch)
{
    if (!h) return EF_ERR_BAD_ARG;
    if (n == 0) return EF_OK;
    if (!d_img || !kp || !d_desc || n < 0) return fail(h, EF_ERR_BAD_ARG, "null image / keypoint / descriptor pointer");
    if (width < 2 || height < 2 || pitch < (size_t)width) return fail(h, EF_ERR_BAD_ARG, "bad image geometry");
    if (width > h->prm.max_width || height > h->prm.max_height) return fail(h, EF_ERR_CAPACITY, "image larger than the handle was created for");
    if (n > h->prm.max_keypoints) return fail(h, EF_ERR_CAPACITY, "more keypoints than max_keypoints of the handle");
    if (desc_pitch < (size_t)desc_bytes_of(h->prm.desc_type)) return fail(h, EF_ERR_BAD_ARG, "desc_pitch smaller than the descriptor size");
    EF_ON_DEVICE(h);
    return -1;
}

int ef_compute_async(ef_handle* h, const uint8_t* d_img, size_t pitch, int width, int height,
                     const float* d_kpts_xysa, int n, uint8_t* d_desc, size_t desc_pitch, void* stream)
{
    const int rc = compute_check(h, d_img, pitch, width, height, d_kpts_xysa, n, d_desc, desc_pitch);
    if (rc != -1) return rc;
    if (((uintptr_t)d_kpts_xysa & 15) != 0) return fail(h, EF_ERR_BAD_ARG, "keypoint array must be 16-byte aligned");
    return compute_common(h, d_img, pitch, width, height, reinterpret_cast<const float4*>(d_kpts_xysa), n, d_desc, desc_pitch, (cudaStream_t)stream);
}

int ef_compute_rows_async(ef_handle* h, const uint8_t* d_img, size_t pitch, int width, int height,
                          const float* d_kpts5, size_t kpts_pitch, int n, uint8_t* d_desc, size_t desc_pitch, void* stream)
{
    const int rc = compute_check(h, d_img, pitch, width, height, d_kpts5, n, d_desc, desc_pitch);
    if (rc != -1) return rc;
    ef_launch_convert_rows(d_kpts5, kpts_pitch, n, h->d_kpts4, (cudaStream_t)stream);
    return compute_common(h, d_img, pitch, width, height, h->d_kpts4, n, d_desc, desc_pitch, (cudaStream_t)stream, true);
}

int ef_detect_and_compute_host_batch(ef_handle* h, int nframes, const uint8_t* h_imgs, size_t img_stride, size_t pitch,
                                     int width, int height, float* h_kpts5, uint8_t* h_desc, int* h_counts, void* stream)
{
    if (!h) return EF_ERR_BAD_ARG;
    if (!h_imgs || !h_kpts5 || !h_counts) return fail(h, EF_ERR_BAD_ARG, "null host pointer");
    if (nframes < 1 || nframes > h->prm.max_batch) return fail(h, EF_ERR_CAPACITY, "nframes exceeds max_batch of the handle");
    if (width > h->prm.max_width || height > h->prm.max_height) return fail(h, EF_ERR_CAPACITY, "image larger than the handle was created for");
    EF_ON_DEVICE(h);
    cudaStream_t s = (cudaStream_t)stream;
    const int nf = h->prm.nfeatures, db = desc_bytes_of(h->prm.desc_type);
    // Chunked pipeline over three streams: the upload of chunk c+1 (getInputMat, cuda_efficient_features.cpp:71-77) and the
    // download of chunk c-1 (:316-320) overlap the kernels of chunk c.  Chunks reuse workspace slots [0, chunk): their
    // kernels are serialised on the caller's stream; outputs land in per-frame staging buffers.
    // chunk plan: a ONE-frame head chunk (the kernels start after 8 MB instead of a whole chunk of uploads), host_chunk frames each,
    std::vector<std::pair<int, int>> plan;   // (first frame, frames)
    if (const char* e = std::getenv("EF_B200_HOST_PLAN")) {
        // explicit chunk sizes for experiments, e.g. "2,4,8,2" (the last size repeats until the batch is covered)
        int f = 0, c = 1;
        const char* q = e;
        while (f < nframes) {
            if (*q) { c = std::max(1, std::atoi(q)); while (*q && *q != ',') q++; if (*q == ',') q++; }
            c = std::min(c, nframes - f);
            plan.emplace_back(f, c); f += c;
        }
    } else if (h->host_chunk <= 0) {
        // growing plan (EF_B200_HOST_CHUNK=0): 1, 3, 6, 12, 12, ... frames -- the kernels start after ONE frame is up, every later chunk is
        // uploaded while the previous (smaller) one computes, launches per frame fall -- and a one-frame tail, so that only 2.5 MB of
        // results are still to be downloaded when the last kernel ends.  Measured at 32 frames (tools/gpu_e2e_plan.sh, ms per step):
        // 1,3,6,12,9,1 17.94; 2,6,12,11,1 17.95; 1,2,4,8,16,1 18.21; 1,3,9,18,1 18.12; 1,4,8,18,1 18.19; no tail (1,3,6,12,10) 18.28.
        // On two alternating compute streams (below) the chunk boundaries overlap and small chunks cost less: 1, 2, 4, 4, ... + tail
        // (17.51 ms; 1,3,6,6,6,6,3,1 17.54; 1,3,6,12,9,1 17.76; 1,2,4,8,16,1 17.94).
        const bool tail = nframes >= 4;
        const int body_end = tail ? nframes - 1 : nframes;
        const bool two_streams = h->host_streams == 2 && !h->overlap_blur && h->s_side;
        const int cap = two_streams ? 4 : 12;  // larger chunks only lengthen the download that is still pending when the last kernel ends
        for (int f = 0, c = 1; f < body_end; f += c, c = two_streams ? std::min(2 * c, cap) : (c == 1 ? 3 : std::min(2 * c, cap))) {
            c = std::min(c, body_end - f); plan.emplace_back(f, c);
        }
        if (tail) plan.emplace_back(nframes - 1, 1);
    } else {
        const int chunk = std::max(1, std::min(h->host_chunk, nframes));
        // and a ONE-frame tail chunk (only 2.5 MB of results are still to be downloaded when the last kernel ends)
        const bool ends = nframes > chunk + 1 && chunk > 1;
        if (ends) plan.emplace_back(0, 1);
        const int body_end = ends ? nframes - 1 : nframes;
        for (int f = ends ? 1 : 0; f < body_end; f += chunk) plan.emplace_back(f, std::min(chunk, body_end - f));
        if (ends) plan.emplace_back(nframes - 1, 1);
    }
    const int nchunks = (int)plan.size();
    if ((int)h->ev_in.size() < nchunks || (int)h->ev_cnt.size() < nchunks) return fail(h, EF_ERR_CAPACITY, "internal: chunk events");
    EF_CUDA(h, cudaEventRecord(h->ev_in[0], s));          // order the uploads after earlier work on the caller's stream
    EF_CUDA(h, cudaStreamWaitEvent(h->s_in, h->ev_in[0], 0));
    for (int c = 0; c < nchunks; c++) {
        for (int f = plan[c].first; f < plan[c].first + plan[c].second; f++)
            EF_CUDA(h, cudaMemcpy2DAsync(h->d_in + f * h->in_stride, h->in_pitch, h_imgs + f * img_stride, pitch, width, height, cudaMemcpyHostToDevice, h->s_in));
        EF_CUDA(h, cudaEventRecord(h->ev_in[c], h->s_in));
    }
    // Chunks alternate between the caller's stream and a second one, each in its own workspace slots [f0, f0 + n): the short, latency-bound
    // launches at the end of one chunk (compaction, selection, angles, the tail of the descriptor kernels) and at the start of the next (the
    // seven dependent pyramid launches) then run next to the other chunk's throughput-bound kernels instead of in front of an idle GPU.
    const bool two = h->host_streams == 2 && nchunks >= 3 && !h->overlap_blur && h->s_side;
    if (two) { EF_CUDA(h, cudaEventRecord(h->ev_fork, s)); EF_CUDA(h, cudaStreamWaitEvent(h->s_side, h->ev_fork, 0)); }
    for (int c = 0; c < nchunks; c++) {
        const int f0 = plan[c].first, n = plan[c].second;
        cudaStream_t cs = (two && (c & 1)) ? h->s_side : s;
        EF_CUDA(h, cudaStreamWaitEvent(cs, h->ev_in[c], 0));
        h->slot0 = two ? f0 : 0;
        int rc = ef_detect_and_compute_batch_async(h, n, h->d_in + f0 * h->in_stride, h->in_stride, h->in_pitch, width, height,
                                                   (float*)((uint8_t*)h->d_out_kpts + f0 * h->out_kpts_stride), h->out_kpts_stride, h->out_kpts_pitch,
                                                   h_desc ? h->d_out_desc + f0 * h->out_desc_stride : nullptr, h->out_desc_stride, (size_t)db,
                                                   h->d_out_counts + f0, cs);
        h->slot0 = 0;
        if (rc != EF_OK) { cudaStreamSynchronize(h->s_in); cudaStreamSynchronize(s); if (two) cudaStreamSynchronize(h->s_side); return rc; }
        EF_CUDA(h, cudaMemcpyAsync(h->h_counts_pinned + f0, h->d_out_counts + f0, sizeof(int) * n, cudaMemcpyDeviceToHost, cs));
        EF_CUDA(h, cudaEventRecord(h->ev_cnt[c], cs));
    }
    if (two) { EF_CUDA(h, cudaEventRecord(h->ev_join, h->s_side)); EF_CUDA(h, cudaStreamWaitEvent(s, h->ev_join, 0)); }
    for (int c = 0; c < nchunks; c++) {
        // one host wait per chunk to size the outputs (the reference blocks 16 times per frame); later chunks keep running
        EF_CUDA(h, cudaEventSynchronize(h->ev_cnt[c]));
        for (int f = plan[c].first; f < plan[c].first + plan[c].second; f++) {
            const int n = h->h_counts_pinned[f];
            h_counts[f] = n;
            if (n <= 0) continue;
            // host layout: 5 x nfeatures floats, nfeatures x db bytes per frame
            EF_CUDA(h, cudaMemcpy2DAsync(h_kpts5 + (size_t)f * EF_ROWS_COUNT * nf, (size_t)nf * 4, (uint8_t*)h->d_out_kpts + f * h->out_kpts_stride,
                                         h->out_kpts_pitch, (size_t)n * 4, EF_ROWS_COUNT, cudaMemcpyDeviceToHost, h->s_out));
            if (h_desc)
                EF_CUDA(h, cudaMemcpyAsync(h_desc + (size_t)f * nf * db, h->d_out_desc + f * h->out_desc_stride, (size_t)n * db, cudaMemcpyDeviceToHost, h->s_out));
        }
    }
    EF_CUDA(h, cudaStreamSynchronize(h->s_out));
    EF_CUDA(h, cudaStreamSynchronize(s));
    return EF_OK;
}

int ef_detect_and_compute_host(ef_handle* h, const uint8_t* h_img, size_t pitch, int width, int height,
                               float* h_kpts5, uint8_t* h_desc, int* h_count, void* stream)
{
    return ef_detect_and_compute_host_batch(h, 1, h_img, 0, pitch, width, height, h_kpts5, h_desc, h_count, stream);
}

// ---- introspection ------------------------------------------------------------------------------
size_t ef_band_candidate_bytes(const ef_handle* h)
{
    // the per-level lists are laid out by the prefix of the quotas, whose sum can exceed nfeatures (see quota_sum)
    return h ? (size_t)ef_align_up((unsigned long long)EF_BAND_HDR + 8ull * (unsigned long long)std::max(quota_sum(h->prm), h->prm.nfeatures), 16) : 0;
}

int ef_band_detect_async(ef_handle* h, int shard, int nshards, int nframes, const uint8_t* d_imgs, size_t img_stride, size_t pitch,
                         int width, int height, uint8_t* d_cand, void* stream)
{
    if (!h) return EF_ERR_BAD_ARG;
    if (!d_imgs || !d_cand) return fail(h, EF_ERR_BAD_ARG, "null image / candidate pointer");
    if (nshards < 1 || shard < 0 || shard >= nshards) return fail(h, EF_ERR_BAD_ARG, "shard must be in [0, nshards)");
    if (pitch < (size_t)width) return fail(h, EF_ERR_BAD_ARG, "pitch smaller than width");
    EF_ON_DEVICE(h);
    cudaStream_t s = (cudaStream_t)stream;
    EfPipe P;
    int rc = build_pipe(h, nframes, width, height, P, shard, nshards);
    if (rc != EF_OK) return rc;
    P.img0 = d_imgs; P.img0_stride = img_stride; P.img0_pitch = (int)pitch;
    h->last_img0 = d_imgs; h->last_img0_stride = img_stride; h->last_img0_pitch = (int)pitch;
    EF_CUDA(h, cudaMemset2DAsync(h->d_ws, h->slot_bytes, 0, h->rowcnt_bytes, P.nframes, s));
    EF_CUDA(h, cudaMemsetAsync(h->d_counters, 0, sizeof(EfLevelCounters) * EF_MAX_LEVELS * P.nframes, s));
    mark(h, -1, s);
    ef_launch_pyramid(P, prepare_tma(h, P), s);      mark(h, EF_STAGE_PYRAMID, s);   // whole pyramid on every GPU: the halo of level s would need level s-1's anyway
    ef_launch_score(P, s);        mark(h, EF_STAGE_SCORE, s);     // owned tile rows + NMS halo
    ef_launch_nms(P, s);          mark(h, EF_STAGE_NMS, s);       // owned tile rows
    ef_launch_compact(P, s);      mark(h, EF_STAGE_COMPACT, s);
    ef_launch_select(P, s);                                       // local top-quota: superset of this band's share of the global one
    ef_launch_band_pack(P, d_cand, ef_band_candidate_bytes(h), s);
    mark(h, EF_STAGE_SELECT, s);
    EF_CUDA(h, cudaGetLastError());
    return EF_OK;
}

int ef_band_finish_async(ef_handle* h, int shard, int nshards, int nframes, const uint8_t* d_all_cand,
                         float* d_kpts, size_t kpts_stride, size_t kpts_pitch, uint8_t* d_desc, size_t desc_stride, size_t desc_pitch,
                         int* d_counts, void* stream)
{
    if (!h) return EF_ERR_BAD_ARG;
    if (!d_all_cand || !d_kpts || !d_counts) return fail(h, EF_ERR_BAD_ARG, "null candidate / keypoint / count pointer");
    if (nshards < 1 || shard < 0 || shard >= nshards) return fail(h, EF_ERR_BAD_ARG, "shard must be in [0, nshards)");
    if (h->last_w == 0 || nframes != h->last_nframes || !h->last_img0) return fail(h, EF_ERR_BAD_ARG, "ef_band_finish_async must follow ef_band_detect_async on the same handle");
    if (kpts_pitch < (size_t)h->prm.nfeatures * 4 || (kpts_pitch & 3)) return fail(h, EF_ERR_BAD_ARG, "kpts_pitch must be >= 4*nfeatures and a multiple of 4");
    const int db = desc_bytes_of(h->prm.desc_type);
    if (d_desc && desc_pitch < (size_t)db) return fail(h, EF_ERR_BAD_ARG, "desc_pitch smaller than the descriptor size");
    EF_ON_DEVICE(h);
    cudaStream_t s = (cudaStream_t)stream;
    EfPipe P;
    int rc = build_pipe(h, nframes, h->last_w, h->last_h, P, shard, nshards);
    if (rc != EF_OK) return rc;
    P.img0 = h->last_img0; P.img0_stride = h->last_img0_stride; P.img0_pitch = h->last_img0_pitch;
    P.kpts = d_kpts; P.kpts_stride = kpts_stride; P.kpts_pitch = (int)kpts_pitch;
    P.desc = d_desc; P.desc_stride = desc_stride; P.desc_pitch = (int)desc_pitch;
    P.counts = d_counts;
    P.select_from_counters = 1;
    mark(h, -1, s);
    ef_launch_band_merge(P, d_all_cand, ef_band_candidate_bytes(h), nshards, s);
    ef_launch_select(P, s);       mark(h, EF_STAGE_SELECT, s);    // global top-quota over the concatenated bands (raster order)
    ef_launch_angle_pack(P, s);   mark(h, EF_STAGE_ANGLE_PACK, s);// every GPU writes the full keypoint matrix (identical everywhere)
    if (d_desc) {
        // Descriptors: this GPU fills output rows [row0, row0 + nrows) (ef_band_desc_rows), the caller all-gathers the equal row
        // blocks in place.  The blur covers every level again (the slice cuts across levels), but each tile first checks whether
        // a window of an owned keypoint can touch it (ef_band_slice_ranges_kernel).
        int row0 = 0, nrows = 0;
        ef_band_desc_rows(h->prm.nfeatures, shard, nshards, &row0, &nrows);
        P.desc_row0 = row0; P.desc_rows = nrows;
        if (nshards > 1) {
            int btiles = 0;
            for (int l = 0; l < P.nlevels; l++) {
                EfLevel& L = P.lv[l];
                L.blur_ty0 = 0; L.blur_rows = ef_div_up(L.h, 64); L.blur_tile_start = btiles;
                if (l >= P.first_level) btiles += L.blur_tiles_x * L.blur_rows;
            }
            P.total_blur_tiles = btiles;
            P.blur_by_slice = 1;
            ef_launch_band_slice_ranges(P, h->d_slice_y, s);
        }
        ef_launch_blur(P, prepare_tma(h, P), s);     mark(h, EF_STAGE_BLUR, s);
        const int v = (db == 32) ? 0 : 1;
        if (is_bad(P.desc_type)) {
            EfBadTables t{ h->d_bad_boxes[v], h->d_bad_radius[v], h->d_bad_thr[v] };
            ef_launch_bad_pipe(P, t, s);
            mark(h, EF_STAGE_DESCRIBE, s);
        } else {
            EfHashSiftTables t{ h->d_exp_table, h->d_grad_table };
            ef_launch_hashsift_features_pipe(P, t, h->d_sift128, s);
            mark(h, EF_STAGE_DESCRIBE, s);
            // the projection runs over all rows: rows of other GPUs come out of stale SIFT vectors and are overwritten by the all-gather
            const EfProjTables pt{ h->d_hs_bfrag[v], h->d_hs_bias[v], h->hs_shift[v], h->d_hs_weights_t[v], h->d_hs_btc[v], h->hs_ndig[v] };
            ef_launch_hashsift_project_batch(h->d_sift128, P.nfeatures, P.counts, P.nframes, pt, db * 8,
                                             P.desc, (size_t)P.desc_stride, P.desc_pitch, nullptr, s);
            mark(h, EF_STAGE_PROJECT, s);
        }
    }
    EF_CUDA(h, cudaGetLastError());
    return EF_OK;
}

int ef_debug_level_view(const ef_handle* h, int frame, int level, ef_level_view* out)
{
    if (!h || !out || level < 0 || level >= h->prm.nlevels || frame < 0 || frame >= h->prm.max_batch || h->last_w == 0) return EF_ERR_BAD_ARG;
    const LevelPlan& q = h->plan[level];
    const uint8_t* slot = h->d_ws + (size_t)frame * h->slot_bytes;
    out->width = h->glast.w[level]; out->height = h->glast.h[level];
    out->scale = h->glast.scale[level]; out->quota = h->glast.quota[level];
    if (level == 0) { out->d_image = h->last_img0 + (size_t)frame * h->last_img0_stride; out->image_pitch = (size_t)h->last_img0_pitch; }
    else { out->d_image = slot + q.img_off; out->image_pitch = (size_t)q.img_pitch; }
    out->d_blurred = slot + q.blur_off; out->blurred_pitch = (size_t)q.img_pitch;
    out->d_response = reinterpret_cast<const float*>(slot + q.resp_off); out->response_pitch = (size_t)q.resp_pitch;
    return EF_OK;
}

int ef_debug_level_counts(ef_handle* h, int frame, int* h_counts3, void* stream)
{
    if (!h || !h_counts3 || frame < 0 || frame >= h->prm.max_batch) return EF_ERR_BAD_ARG;
    EfLevelCounters c[EF_MAX_LEVELS];
    EF_CUDA(h, cudaMemcpyAsync(c, h->d_counters + (size_t)frame * EF_MAX_LEVELS, sizeof(c), cudaMemcpyDeviceToHost, (cudaStream_t)stream));
    EF_CUDA(h, cudaStreamSynchronize((cudaStream_t)stream));
    for (int l = 0; l < h->prm.nlevels; l++) { h_counts3[3 * l] = c[l].corners; h_counts3[3 * l + 1] = c[l].survivors; h_counts3[3 * l + 2] = c[l].selected; }
    return EF_OK;
}

int ef_debug_keep_projection(ef_handle* h, int keep)
{
    if (!h) return EF_ERR_BAD_ARG;
    if (keep && !h->d_proj) EF_CUDA(h, cudaMalloc((void**)&h->d_proj, h->sift_rows * 512 * sizeof(float)));
    h->keep_proj = keep != 0;
    return EF_OK;
}

int ef_debug_hashsift_views(const ef_handle* h, const uint8_t** d_sift128, const float** d_projection)
{
    if (!h) return EF_ERR_BAD_ARG;
    if (d_sift128) *d_sift128 = h->d_sift128;
    if (d_projection) *d_projection = h->d_proj;
    return EF_OK;
}

int ef_debug_project_async(ef_handle* h, const uint8_t* d_sift128, int n, int path, uint8_t* d_desc, size_t desc_pitch, void* stream)
{
    if (!h || !d_sift128 || !d_desc || n < 0 || path < 0 || path > 3) return EF_ERR_BAD_ARG;
    if (is_bad(h->prm.desc_type)) return fail(h, EF_ERR_BAD_ARG, "the handle's descriptor type is not HashSIFT");
    if (n == 0) return EF_OK;
    EF_ON_DEVICE(h);
    const int db = desc_bytes_of(h->prm.desc_type), v = db == 32 ? 0 : 1;
    const EfProjTables pt{ h->d_hs_bfrag[v], h->d_hs_bias[v], h->hs_shift[v], h->d_hs_weights_t[v], h->d_hs_btc[v], h->hs_ndig[v] };
    g_ef_project_path = path;   // tests only: not thread safe
    ef_launch_hashsift_project_batch(d_sift128, n, nullptr, 1, pt, db * 8, d_desc, 0, (int)desc_pitch, nullptr, (cudaStream_t)stream);
    g_ef_project_path = 0;
    EF_CUDA(h, cudaGetLastError());
    return EF_OK;
}

int ef_stage_timing_enable(ef_handle* h, int enable)
{
    if (!h) return EF_ERR_BAD_ARG;
    h->timing = enable != 0;
    h->ev_used = 0;
    return EF_OK;
}

int ef_stage_times(ef_handle* h, float* ms_sum, int* ncalls)
{
    if (!h || !ms_sum) return EF_ERR_BAD_ARG;
    for (int i = 0; i < EF_NUM_STAGES; i++) ms_sum[i] = 0.f;
    int calls = 0;
    if (h->ev_used) EF_CUDA(h, cudaEventSynchronize(h->ev_pool[h->ev_used - 1]));
    for (size_t i = 0; i < h->ev_used; i++) {
        if (h->ev_stage[i] < 0) { if (h->ev_stage[i] == -1) calls++; continue; }
        float ms = 0.f;
        EF_CUDA(h, cudaEventElapsedTime(&ms, h->ev_pool[i - 1], h->ev_pool[i]));
        ms_sum[h->ev_stage[i]] += ms;
    }
    if (ncalls) *ncalls = calls;
    h->ev_used = 0;
    return EF_OK;
}

unsigned long long ef_kernel_launch_count(void) { return g_ef_launches; }

int ef_debug_copy_to_host(ef_handle* h, void* dst, size_t dst_pitch, const void* d_src, size_t src_pitch, size_t width_bytes, size_t rows)
{
    if (!h || !dst || !d_src) return EF_ERR_BAD_ARG;
    EF_CUDA(h, cudaDeviceSynchronize());
    EF_CUDA(h, cudaMemcpy2D(dst, dst_pitch, d_src, src_pitch, width_bytes, rows, cudaMemcpyDeviceToHost));
    return EF_OK;
}

} // extern "C"
