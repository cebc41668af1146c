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

void ef_launch_pyramid(const EfPipe& p, cudaStream_t s)
{
    for (int l = 1; l < p.nlevels; l++) {
        const dim3 block(64, 4);
        const dim3 grid(ef_div_up(p.lv[l].w, 256), ef_div_up(p.lv[l].h, 4), p.nframes);
        ef_resize_kernel<<<grid, block, 0, s>>>(p, l);
        EF_COUNT_LAUNCH(1);
    }
}

// =================================================================================================
// score: FAST-9/16 inside the 15-px border + Harris at every corner -> dense response map
// =================================================================================================
#define SC_HALO 4
#define SC_W (EF_TILE + 2 * SC_HALO) // 40
#define SC_G (EF_TILE + 6)           // 38: gradient region [-3, 34]

__device__ __forceinline__ bool ef_has_arc9(unsigned m)
{
    const unsigned mm = m | (m << 16);
    unsigned a = mm & (mm >> 1);
    a &= a >> 2;
    a &= a >> 4;
    a &= mm >> 8;
    return (a & 0xffffu) != 0;
}

__device__ __forceinline__ int ef_find_level(const EfPipe& p, int idx, int EfLevel::*start)
{
    int level = p.first_level;
    while (level + 1 < p.nlevels && idx >= p.lv[level + 1].*start) level++;
    return level;
}

__global__ void __launch_bounds__(256) ef_score_kernel(const __grid_constant__ EfPipe p)
{
    __shared__ uint8_t s_img[SC_W][SC_W + 8];
    __shared__ float2 s_grad[SC_G][SC_G + 1];
    __shared__ __align__(16) float s_resp[EF_TILE][EF_TILE];
    __shared__ unsigned short s_list[EF_TILE * EF_TILE];
    __shared__ int s_n;

    const int tid = threadIdx.x;
    const int frame = blockIdx.y;
    const int level = ef_find_level(p, blockIdx.x, &EfLevel::tile_start);
    const EfLevel& L = p.lv[level];
    const int t = blockIdx.x - L.tile_start;
    const int x0 = (t % L.tiles_x) * EF_TILE, y0 = (t / L.tiles_x) * EF_TILE;

    int pitch;
    const uint8_t* __restrict__ img = ef_level_image(p, frame, level, pitch);

    if (tid == 0) s_n = 0;
    for (int i = tid; i < SC_W * SC_W; i += 256) {
        const int ly = i / SC_W, lx = i - ly * SC_W;
        const int gy = min(max(y0 - SC_HALO + ly, 0), L.h - 1);
        const int gx = min(max(x0 - SC_HALO + lx, 0), L.w - 1);
        s_img[ly][lx] = img[(size_t)gy * pitch + gx];
    }
    __syncthreads();

    // ---- FAST-9/16 (cuda_fast.cu:36-40,162-166: mask1 = darker, mask2 = brighter, >= 9 contiguous)
    const int tx = tid & 31, ty = tid >> 5, lane = tx;
    const int th = p.fast_threshold;
#pragma unroll
    for (int i = 0; i < 4; i++) {
        const int py = ty + 8 * i;
        const int gx = x0 + tx, gy = y0 + py;
        bool corner = false;
        if (gx >= EF_HALF_PATCH && gx < L.w - EF_HALF_PATCH && gy >= EF_HALF_PATCH && gy < L.h - EF_HALF_PATCH) {
            const int cy = py + SC_HALO, cx = tx + SC_HALO;
            const int v = s_img[cy][cx];
            const int lo = v - th, hi = v + th;
            unsigned dark = 0, bright = 0;
#define EF_RING(k, dy, dx) { const int q = s_img[cy + (dy)][cx + (dx)]; dark |= (unsigned)(q < lo) << (k); bright |= (unsigned)(q > hi) << (k); }
            EF_RING(0, 3, 0) EF_RING(8, -3, 0) EF_RING(4, 0, 3) EF_RING(12, 0, -3)
            if ((dark | bright) != 0) { // at least one of the 4 compass pixels differs, else no 9-arc is possible
                EF_RING(1, 3, 1) EF_RING(2, 2, 2) EF_RING(3, 1, 3) EF_RING(5, -1, 3) EF_RING(6, -2, 2) EF_RING(7, -3, 1)
                EF_RING(9, -3, -1) EF_RING(10, -2, -2) EF_RING(11, -1, -3) EF_RING(13, 1, -3) EF_RING(14, 2, -2) EF_RING(15, 3, -1)
                corner = ef_has_arc9(dark) || ef_has_arc9(bright);
            }
#undef EF_RING
        }
        s_resp[py][tx] = EF_NEG_INF;
        const unsigned bal = __ballot_sync(0xffffffffu, corner);
        const int cnt = __popc(bal);
        int base = 0;
        if (lane == 0 && cnt) base = atomicAdd(&s_n, cnt);
        base = __shfl_sync(0xffffffffu, base, 0);
        if (corner) s_list[base + __popc(bal & ((1u << lane) - 1u))] = (unsigned short)((py << 5) | tx);
    }
    __syncthreads();
    const int n = s_n;

    if (n > 0) {
        // ---- Sobel gradients of the tile (+3 halo), cuda_efficient_features.cu:116-128
        const float SCALE = 1.f / (4 * 7 * 255);
        for (int i = tid; i < SC_G * SC_G; i += 256) {
            const int gyl = i / SC_G, gxl = i - gyl * SC_G;
            const int ly = gyl + 1, lx = gxl + 1;
            const int v00 = s_img[ly - 1][lx - 1], v01 = s_img[ly - 1][lx], v02 = s_img[ly - 1][lx + 1];
            const int v10 = s_img[ly][lx - 1], v12 = s_img[ly][lx + 1];
            const int v20 = s_img[ly + 1][lx - 1], v21 = s_img[ly + 1][lx], v22 = s_img[ly + 1][lx + 1];
            float2 g;
            g.x = SCALE * (float)((v02 + 2 * v12 + v22) - (v00 + 2 * v10 + v20));
            g.y = SCALE * (float)((v20 + 2 * v21 + v22) - (v00 + 2 * v01 + v02));
            s_grad[gyl][gxl] = g;
        }
        __syncthreads();
        // ---- Harris, raster order over the 7x7 block with the reference's contraction
        for (int i = tid; i < n; i += 256) {
            const int pos = s_list[i];
            const int cx = pos & 31, cy = pos >> 5;
            float sxx = 0.f, sxy = 0.f, syy = 0.f;
#pragma unroll
            for (int iy = 0; iy < 7; iy++) {
#pragma unroll
                for (int ix = 0; ix < 7; ix++) {
                    const float2 g = s_grad[cy + iy][cx + ix];
                    sxx = fmaf(g.x, g.x, sxx);
                    sxy = fmaf(g.x, g.y, sxy);
                    syy = fmaf(g.y, g.y, syy);
                }
            }
            const float p2 = sxy * sxy;
            const float det = fmaf(sxx, syy, -p2);
            const float tr = sxx + syy;
            const float tt = tr * (-0.04f);
            s_resp[cy][cx] = fmaf(tr, tt, det);
        }
        if (tid == 0) atomicAdd(&p.counters[frame * EF_MAX_LEVELS + level].corners, n);
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

__global__ void __launch_bounds__(256) ef_nms_kernel(const __grid_constant__ EfPipe p)
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
    const int ty = s / L.strips_x, tx0 = (s - ty * L.strips_x) * EF_NMS_RT;
    const int ntx = min(EF_NMS_RT, L.tiles_x - tx0);
    const int x0 = tx0 * EF_TILE, y0 = ty * EF_TILE;
    const float* __restrict__ resp = reinterpret_cast<const float*>(ef_ws(p, frame, L.resp_off));

    const int b = p.nms_block;
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
        const int K = p.nms_K, r2 = p.nms_r2;
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
                for (int n = part; n < nnb; n += 4) {
                    if (n != (nnb >> 1)) {                                    // the centre of the window is the block itself
                        const EfBlockMax e = s_blk[nbase + wy * side_x + wx];
                        if (e.val >= r) {
                            const int ex = (int)(e.pos & 0xffff) - cx, ey = (int)((e.pos >> 16) & 0x7fff) - cy;
                            if (ex * ex + ey * ey < r2) kill = true;
                        }
                    }
                    wx += 4;
                    while (wx >= wside) { wx -= wside; wy++; }
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
                    while (wx >= wside) { wx -= wside; wy++; }
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
    ef_nms_kernel<<<dim3(p.total_strips, p.nframes), 256, 0, s>>>(p);
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
    const int band = blockIdx.x - L.band_start;
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

__global__ void __launch_bounds__(1024) ef_select_kernel(const __grid_constant__ EfPipe p)
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

    int partial = 0;
    for (int i = tid; i < L.h; i += 1024) partial += rowcnt[i];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) partial += __shfl_xor_sync(0xffffffffu, partial, o);
    if (lane == 0) s_warp[warp] = partial;
    __syncthreads();
    int ntotal = 0;
    for (int i = 0; i < 32; i++) ntotal += s_warp[i];
    __syncthreads();
    const int n = min(ntotal, L.surv_cap);
    const int quota = L.quota;

    if (n <= quota) {
        for (int i = tid; i < n; i += 1024) {
            const EfSurvivor sv = surv[i];
            EfSelected o; o.x = sv.x; o.y = sv.y; o.resp = sv.resp; o.angle = 0.f; o.pad = 0;
            sel[i] = o;
        }
        if (tid == 0) { ctr->survivors = ntotal; ctr->selected = n; ctr->overflow = ntotal > L.surv_cap; }
        return;
    }

    // radix select: key of the quota-th largest response
    unsigned prefix = 0, pmask = 0;
    int k = quota;
    for (int shift = 24; shift >= 0; shift -= 8) {
        if (tid < 256) s_hist[tid] = 0;
        __syncthreads();
        for (int i = tid; i < n; i += 1024) {
            const unsigned key = ef_float_key(surv[i].resp);
            if ((key & pmask) == prefix) atomicAdd(&s_hist[(key >> shift) & 255u], 1u);
        }
        __syncthreads();
        if (tid == 0) {
            int acc = 0, b = 255;
            for (; b > 0; b--) {
                const int hcnt = (int)s_hist[b];
                if (acc + hcnt >= k) break;
                acc += hcnt;
            }
            s_prefix = prefix | ((unsigned)b << shift);
            s_k = k - acc;
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
    if (tid == 0) { ctr->survivors = ntotal; ctr->selected = min(run_sel, quota); ctr->overflow = ntotal > L.surv_cap; }
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

__global__ void __launch_bounds__(256) ef_angle_pack_kernel(const __grid_constant__ EfPipe p)
{
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
    const int i = (blockIdx.x - L.kpt_block_start) * EF_KPTS_PER_CTA + warp;
    if (i >= ctr[level].selected) return;
    int offset = 0;
    for (int l = p.first_level; l < level; l++) offset += ctr[l].selected;

    EfSelected* sel = reinterpret_cast<EfSelected*>(ef_ws(p, frame, L.sel_off));
    const EfSelected k = sel[i];
    int pitch;
    const uint8_t* __restrict__ img = ef_level_image(p, frame, level, pitch);
    const uint8_t* c = img + (size_t)k.y * pitch + k.x;

    // IC_Angle, cuda_efficient_features.cu:141-172: lane <-> dx = lane-15
    int m01 = 0, m10 = 0;
    const int dx = lane - EF_HALF_PATCH;
    if (lane < 31) {
        m10 = dx * (int)c[dx];
        const int adx = abs(dx);
        for (int dy = 1; dy <= EF_HALF_PATCH; dy++) {
            if (adx <= c_umax[dy]) {
                const int vT = c[-(ptrdiff_t)dy * pitch + dx];
                const int vB = c[(ptrdiff_t)dy * pitch + dx];
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
    if (lane == 0) {
        // canonical: atan2 in double, rounded once (the reference's CUDA atan2f is <= 2 ulp)
        float angle = (float)atan2((double)(float)m01, (double)(float)m10);
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
// u8 -> u8 with round-half-even.  64x32 tile, all levels in one launch.
// =================================================================================================
#define BL_TW 64
#define BL_TH 32
__global__ void __launch_bounds__(256) ef_blur_kernel(const __grid_constant__ EfPipe p)
{
    __shared__ uint8_t s_in[BL_TH + 6][BL_TW + 8];
    __shared__ float s_row[BL_TH + 6][BL_TW];

    const float taps[7] = { __uint_as_float(0x3d8fafb1u), __uint_as_float(0x3e06387eu), __uint_as_float(0x3e434a39u),
                            __uint_as_float(0x3e5d4ae0u), __uint_as_float(0x3e434a39u), __uint_as_float(0x3e06387eu),
                            __uint_as_float(0x3d8fafb1u) };
    const int tid = threadIdx.x;
    const int frame = blockIdx.y;
    const int level = ef_find_level(p, blockIdx.x, &EfLevel::blur_tile_start);
    const EfLevel& L = p.lv[level];
    const int t = blockIdx.x - L.blur_tile_start;
    const int x0 = (t % L.blur_tiles_x) * BL_TW, y0 = (t / L.blur_tiles_x) * BL_TH;
    int pitch;
    const uint8_t* __restrict__ img = ef_level_image(p, frame, level, pitch);

    for (int i = tid; i < (BL_TH + 6) * (BL_TW + 6); i += 256) {
        const int ly = i / (BL_TW + 6), lx = i - ly * (BL_TW + 6);
        const int gy = ef_reflect101(min(y0 - 3 + ly, L.h + 2), L.h);
        const int gx = ef_reflect101(min(x0 - 3 + lx, L.w + 2), L.w);
        s_in[ly][lx] = img[(size_t)gy * pitch + gx];
    }
    __syncthreads();
    for (int i = tid; i < (BL_TH + 6) * BL_TW; i += 256) {
        const int ly = i / BL_TW, lx = i - ly * BL_TW;
        float sum = 0.f;
#pragma unroll
        for (int k = 0; k < 7; k++) sum = fmaf((float)s_in[ly][lx + k], taps[k], sum);
        s_row[ly][lx] = sum;
    }
    __syncthreads();
    uint8_t* out = ef_ws(p, frame, L.blur_off);
    for (int i = tid; i < BL_TH * BL_TW / 4; i += 256) {
        const int ly = i / (BL_TW / 4), lx = (i - ly * (BL_TW / 4)) * 4;
        const int gy = y0 + ly, gx = x0 + lx;
        if (gy >= L.h || gx >= L.w) continue;
        unsigned packed = 0;
#pragma unroll
        for (int j = 0; j < 4; j++) {
            float sum = 0.f;
#pragma unroll
            for (int k = 0; k < 7; k++) sum = fmaf(s_row[ly + k][lx + j], taps[k], sum);
            packed |= ef_sat_u8_rne(sum) << (8 * j);
        }
        *reinterpret_cast<unsigned*>(out + (size_t)gy * L.blur_pitch + gx) = packed;
    }
}

void ef_launch_blur(const EfPipe& p, cudaStream_t s)
{
    if (p.total_blur_tiles <= 0) return;
    ef_blur_kernel<<<dim3(p.total_blur_tiles, p.nframes), 256, 0, s>>>(p);
    EF_COUNT_LAUNCH(1);
}
