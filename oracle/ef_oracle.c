/*
 * ef_oracle.c -- CPU ORACLE (test infrastructure, see ef_oracle.h for the rules and the pinning status).
 *
 * Build: gcc -O2 -std=gnu11 -ffp-contract=off -mfma -fopenmp -fPIC -shared   (oracle/Makefile)
 *   -ffp-contract=off : the reference CPU module is built for generic x86-64 (no FMA,
 *                       modules/efficient_features/CMakeLists.txt:22-24), so no contraction anywhere;
 *   -mfma             : only so that the EXPLICIT fmaf() calls below (the contraction pattern nvcc
 *                       emits for the reference's CUDA detector) inline to one instruction.
 * Every function cites the reference lines it follows.
 */
#include "ef_oracle.h"

#include <float.h>
#include <math.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#include "../cuda-efficient-features_b200/csrc/params/ef_bad_tables.inc"
#include "../cuda-efficient-features_b200/csrc/params/ef_hashsift_w256.inc"
#include "../cuda-efficient-features_b200/csrc/params/ef_hashsift_w512.inc"

static int g_threads = 1;

void efo_set_threads(int n) { g_threads = n < 1 ? 1 : n; }
int efo_get_max_threads(void)
{
#ifdef _OPENMP
    return omp_get_num_procs();
#else
    return 1;
#endif
}

static inline float u32_as_f32(uint32_t u) { float f; memcpy(&f, &u, 4); return f; }

/* OpenCV helpers (x86-64 SSE2 build): cvRound = round-half-even, cvFloor = floor, cvCeil = ceil,
 * saturate_cast<uchar>(float) = clamp(cvRound(v)). */
static inline int cv_round_f(float v) { return (int)lrintf(v); }
static inline int cv_round_d(double v) { return (int)lrint(v); }
static inline int cv_floor_f(float v) { int i = (int)v; return i - (i > v); }
static inline uint8_t sat_u8_rne(float v) { int i = cv_round_f(v); return (uint8_t)(i < 0 ? 0 : (i > 255 ? 255 : i)); }

/* ------------------------------------------------------------------------------------------- */
/* geometry: cuda_efficient_features.cpp:136-157 (sizes/scales), :159-174 (quotas)              */
/* ------------------------------------------------------------------------------------------- */
void efo_level_geometry(int w, int h, float scale_factor, int nlevels, int* ws, int* hs, float* scales)
{
    float scale = 1.f;
    ws[0] = w; hs[0] = h; scales[0] = scale;
    for (int s = 1; s < nlevels; s++) {
        scale *= scale_factor;                      /* :150 */
        const float inv = 1.f / scale;              /* :151 */
        hs[s] = cv_round_f(inv * (float)h);         /* :152 */
        ws[s] = cv_round_f(inv * (float)w);         /* :153 */
        scales[s] = scale;
    }
}

void efo_level_quotas(int nfeatures, float scale_factor, int nlevels, int* quotas)
{
    const double factor = (double)(1 / scale_factor);                       /* :164, float division */
    double nf = nfeatures * (1 - factor) / (1 - pow(factor, nlevels));      /* :165 */
    int sum = 0;
    for (int s = 0; s < nlevels - 1; s++) {
        quotas[s] = cv_round_d(nf);
        sum += quotas[s];
        nf *= factor;
    }
    quotas[nlevels - 1] = nfeatures - sum > 0 ? nfeatures - sum : 0;
}

/* ------------------------------------------------------------------------------------------- */
/* cv::cuda::resize(INTER_LINEAR) -- call site cuda_efficient_features.cpp:154                  */
/* Third-party (opencv_contrib cudawarping, OpenCV >= 4.6, not vendored): restated from SURVEY   */
/* Appendix A.1.  PARITY UNPINNED.                                                               */
/* ------------------------------------------------------------------------------------------- */
void efo_resize_linear(const uint8_t* src, int sw, int sh, size_t spitch, uint8_t* dst, int dw, int dh, size_t dpitch)
{
    const float rx = (float)(1.0 / ((double)dw / sw));
    const float ry = (float)(1.0 / ((double)dh / sh));
#pragma omp parallel for num_threads(g_threads) schedule(static)
    for (int y = 0; y < dh; y++) {
        const float sy = (float)y * ry;
        const int y1 = (int)floorf(sy);
        const int y2 = y1 + 1;
        const int y2r = y2 < sh - 1 ? y2 : sh - 1;
        const float wy1 = (float)y2 - sy, wy2 = sy - (float)y1;
        for (int x = 0; x < dw; x++) {
            const float sx = (float)x * rx;
            const int x1 = (int)floorf(sx);
            const int x2 = x1 + 1;
            const int x2r = x2 < sw - 1 ? x2 : sw - 1;
            const float wx1 = (float)x2 - sx, wx2 = sx - (float)x1;
            float out = 0.f;
            out = fmaf((float)src[(size_t)y1 * spitch + x1], wx1 * wy1, out);
            out = fmaf((float)src[(size_t)y1 * spitch + x2r], wx2 * wy1, out);
            out = fmaf((float)src[(size_t)y2r * spitch + x1], wx1 * wy2, out);
            out = fmaf((float)src[(size_t)y2r * spitch + x2r], wx2 * wy2, out);
            dst[(size_t)y * dpitch + x] = sat_u8_rne(out);
        }
    }
}

/* ------------------------------------------------------------------------------------------- */
/* cv::cuda::createGaussianFilter(CV_8UC1,-1,Size(7,7),2,2,BORDER_REFLECT_101)                   */
/* call sites cuda_efficient_features.cpp:193,305.  Third-party; SURVEY Appendix A.2.            */
/* PARITY UNPINNED.  Taps = cv2.getGaussianKernel(7, 2, CV_32F) bit patterns.                    */
/* ------------------------------------------------------------------------------------------- */
static const uint32_t k_gauss7_bits[7] = { 0x3d8fafb1u, 0x3e06387eu, 0x3e434a39u, 0x3e5d4ae0u,
                                           0x3e434a39u, 0x3e06387eu, 0x3d8fafb1u };

static inline int reflect101(int i, int n)
{
    if (n == 1) return 0;
    while (i < 0 || i >= n) {
        if (i < 0) i = -i;
        else i = 2 * (n - 1) - i;
    }
    return i;
}

void efo_gaussian_blur7(const uint8_t* src, int w, int h, size_t spitch, uint8_t* dst, size_t dpitch)
{
    float taps[7];
    for (int k = 0; k < 7; k++) taps[k] = u32_as_f32(k_gauss7_bits[k]);
    float* rowbuf = (float*)malloc(sizeof(float) * (size_t)w * h);
#pragma omp parallel for num_threads(g_threads) schedule(static)
    for (int y = 0; y < h; y++) {
        const uint8_t* s = src + (size_t)y * spitch;
        float* r = rowbuf + (size_t)y * w;
        for (int x = 0; x < w; x++) {
            float sum = 0.f;
            for (int k = 0; k < 7; k++) sum = fmaf((float)s[reflect101(x + k - 3, w)], taps[k], sum);
            r[x] = sum;
        }
    }
#pragma omp parallel for num_threads(g_threads) schedule(static)
    for (int y = 0; y < h; y++) {
        uint8_t* d = dst + (size_t)y * dpitch;
        const float* rows[7];
        for (int k = 0; k < 7; k++) rows[k] = rowbuf + (size_t)reflect101(y + k - 3, h) * w;
        for (int x = 0; x < w; x++) {
            float sum = 0.f;
            for (int k = 0; k < 7; k++) sum = fmaf(rows[k][x], taps[k], sum);
            d[x] = sat_u8_rne(sum);
        }
    }
    free(rowbuf);
}

/* ------------------------------------------------------------------------------------------- */
/* FAST-9/16: cuda_fast.cu:36-40 (diffType), :42-157 (calcMask bit layout), :162-166             */
/* (c_table == ">= 9 circularly contiguous bits", verified exhaustively in the survey), :168-222 */
/* ------------------------------------------------------------------------------------------- */
static const int k_ring_dy[16] = { 3, 3, 2, 1, 0, -1, -2, -3, -3, -3, -2, -1, 0, 1, 2, 3 };
static const int k_ring_dx[16] = { 0, 1, 2, 3, 3, 3, 2, 1, 0, -1, -2, -3, -3, -3, -2, -1 };

static inline int has_arc9(unsigned m)
{
    /* 16-bit circular: AND of 9 successive rotations is non-zero iff there are 9 contiguous set bits */
    unsigned mm = m | (m << 16);
    unsigned a = mm & (mm >> 1);
    a &= a >> 2;  /* 4 contiguous */
    a &= a >> 4;  /* 8 contiguous */
    a &= mm >> 8; /* 9 contiguous */
    return (a & 0xffffu) != 0;
}

int efo_fast_is_corner(const uint8_t* img, size_t pitch, int x, int y, int th)
{
    const int v = img[(size_t)y * pitch + x];
    unsigned darker = 0, brighter = 0;
    for (int k = 0; k < 16; k++) {
        const int p = img[(size_t)(y + k_ring_dy[k]) * pitch + (x + k_ring_dx[k])];
        const int diff = p - v;                 /* :38 */
        darker |= (unsigned)(diff < -th) << k;  /* mask1 */
        brighter |= (unsigned)(diff > th) << k; /* mask2 */
    }
    return has_arc9(darker) || has_arc9(brighter);
}

/* ------------------------------------------------------------------------------------------- */
/* Harris response: cuda_efficient_features.cu:99-139; contraction as emitted by nvcc 12.9 for    */
/* that source: three fmaf chains, det = fmaf(sxx,syy,-(sxy*sxy)),             */
/* resp = fmaf(tr, tr*(-0.04f), det).                                                             */
/* ------------------------------------------------------------------------------------------- */
float efo_harris_response(const uint8_t* img, size_t pitch, int x0, int y0)
{
    const float SCALE = 1.f / (4 * 7 * 255);
    float sxx = 0, sxy = 0, syy = 0;
    for (int iy = -3; iy <= 3; ++iy) {
        for (int ix = -3; ix <= 3; ++ix) {
            const uint8_t* p = img + (size_t)(y0 + iy) * pitch + (x0 + ix);
            const int v00 = p[-(ptrdiff_t)pitch - 1], v01 = p[-(ptrdiff_t)pitch], v02 = p[-(ptrdiff_t)pitch + 1];
            const int v10 = p[-1], v12 = p[1];
            const int v20 = p[pitch - 1], v21 = p[pitch], v22 = p[pitch + 1];
            const float dx = SCALE * (float)((v02 + 2 * v12 + v22) - (v00 + 2 * v10 + v20));
            const float dy = SCALE * (float)((v20 + 2 * v21 + v22) - (v00 + 2 * v01 + v02));
            sxx = fmaf(dx, dx, sxx);
            sxy = fmaf(dx, dy, sxy);
            syy = fmaf(dy, dy, syy);
        }
    }
    const float p2 = sxy * sxy;
    const float det = fmaf(sxx, syy, -p2);
    const float tr = sxx + syy;
    const float t = tr * (-0.04f);
    return fmaf(tr, t, det);
}

/* ------------------------------------------------------------------------------------------- */
/* IC angle: cuda_efficient_features.cu:141-172, :54-60.  Integer moments are exact; the          */
/* reference calls CUDA's atan2f (<= 2 ulp, not reproducible on a CPU).  Canonical definition     */
/* here and in the kernel: atan2 evaluated in double, rounded once to float.          */
/* ------------------------------------------------------------------------------------------- */
static const int k_umax[17] = { 15, 15, 15, 15, 14, 14, 14, 13, 13, 12, 11, 10, 9, 8, 6, 3, 0 };

float efo_ic_angle(const uint8_t* img, size_t pitch, int x, int y)
{
    int m_01 = 0, m_10 = 0;
    const uint8_t* c = img + (size_t)y * pitch + x;
    for (int dx = -15; dx <= 15; ++dx) m_10 += dx * c[dx];
    for (int dy = 1; dy <= 15; ++dy) {
        int y_sum = 0;
        const int d = k_umax[dy];
        for (int dx = -d; dx <= d; ++dx) {
            const int valT = c[-(ptrdiff_t)dy * (ptrdiff_t)pitch + dx];
            const int valB = c[(ptrdiff_t)dy * (ptrdiff_t)pitch + dx];
            y_sum += (valB - valT);
            m_10 += dx * (valB + valT);
        }
        m_01 += dy * y_sum;
    }
    float angle = (float)atan2((double)(float)m_01, (double)(float)m_10);
    const float PI = (float)3.1415926535897932384626433832795;
    if (angle < 0) angle += 2.f * PI;
    return (180.f / PI) * angle;
}

/* ------------------------------------------------------------------------------------------- */
/* score map = FAST (inside the 15-px border mask, cuda_efficient_features.cpp:176-182,250) +     */
/* Harris at every corner.  No candidate cap (DESIGN.md rule H1).                                 */
/* ------------------------------------------------------------------------------------------- */
long efo_score_map(const uint8_t* img, int w, int h, size_t pitch, int th, float* resp)
{
    const int B = 15;
    long total = 0;
#pragma omp parallel for num_threads(g_threads) schedule(dynamic, 16) reduction(+ : total)
    for (int y = 0; y < h; y++) {
        float* r = resp + (size_t)y * w;
        for (int x = 0; x < w; x++) r[x] = -INFINITY;
        if (y < B || y >= h - B) continue;
        for (int x = B; x < w - B; x++) {
            if (efo_fast_is_corner(img, pitch, x, y, th)) {
                r[x] = efo_harris_response(img, pitch, x, y);
                total++;
            }
        }
    }
    return total;
}

/* ------------------------------------------------------------------------------------------- */
/* radius NMS: cuda_efficient_features.cu:62-97 (IsMaxPoint), :202-216, :291-292.                 */
/* i dies iff there is j != i with resp_i <= resp_j and d^2 < ceil(r^2).  The cell grid of the    */
/* reference is only an acceleration structure (3x3 cells of 16 px cover every d^2 < r^2).        */
/* Survivors are emitted in raster order (DESIGN.md rule H2; the reference's order is atomic).    */
/* ------------------------------------------------------------------------------------------- */
long efo_radius_nms(const float* resp, int w, int h, int radius, short* xs, short* ys, float* rs, long cap)
{
    const float rf = (float)radius;
    const int r2 = (int)ceilf(rf * rf);
    int R = 0;
    while ((R + 1) * (R + 1) < r2) R++;
    uint8_t* keep = (uint8_t*)calloc((size_t)w * h, 1);
#pragma omp parallel for num_threads(g_threads) schedule(dynamic, 16)
    for (int y = 0; y < h; y++) {
        for (int x = 0; x < w; x++) {
            const float ri = resp[(size_t)y * w + x];
            if (!(ri > -INFINITY)) continue;
            int alive = 1;
            const int y0 = y - R < 0 ? 0 : y - R, y1 = y + R >= h ? h - 1 : y + R;
            const int x0 = x - R < 0 ? 0 : x - R, x1 = x + R >= w ? w - 1 : x + R;
            for (int yy = y0; yy <= y1 && alive; yy++) {
                const int dy = yy - y;
                const float* row = resp + (size_t)yy * w;
                for (int xx = x0; xx <= x1; xx++) {
                    const int dx = xx - x;
                    if ((dx | dy) == 0) continue;
                    if (ri <= row[xx] && dx * dx + dy * dy < r2) { alive = 0; break; }
                }
            }
            keep[(size_t)y * w + x] = (uint8_t)alive;
        }
    }
    long n = 0;
    for (int y = 0; y < h; y++)
        for (int x = 0; x < w; x++)
            if (keep[(size_t)y * w + x]) {
                if (n < cap) { xs[n] = (short)x; ys[n] = (short)y; rs[n] = resp[(size_t)y * w + x]; }
                n++;
            }
    free(keep);
    return n;
}

/* top-K: cuda_efficient_features.cu:344-358 sorts by response (descending) and truncates; ties are
 * arrival-order dependent there.  Total order here: (response desc, y asc, x asc). */
typedef struct { float r; short x, y; } efo_cand;
static int cmp_resp_desc(const void* a, const void* b)
{
    const efo_cand* p = (const efo_cand*)a; const efo_cand* q = (const efo_cand*)b;
    if (p->r != q->r) return p->r > q->r ? -1 : 1;
    if (p->y != q->y) return p->y < q->y ? -1 : 1;
    return p->x < q->x ? -1 : (p->x > q->x ? 1 : 0);
}
static int cmp_raster(const void* a, const void* b)
{
    const efo_cand* p = (const efo_cand*)a; const efo_cand* q = (const efo_cand*)b;
    if (p->y != q->y) return p->y < q->y ? -1 : 1;
    return p->x < q->x ? -1 : (p->x > q->x ? 1 : 0);
}

size_t efo_build_pyramid(const uint8_t* img, int w, int h, size_t pitch, float scale_factor, int nlevels,
                         int blurred, uint8_t* levels, size_t* offsets)
{
    int ws[EFO_MAX_LEVELS], hs[EFO_MAX_LEVELS]; float sc[EFO_MAX_LEVELS];
    efo_level_geometry(w, h, scale_factor, nlevels, ws, hs, sc);
    size_t total = 0;
    for (int s = 0; s < nlevels; s++) { if (offsets) offsets[s] = total; total += (size_t)ws[s] * hs[s]; }
    if (!levels) return total;
    size_t off = 0;
    uint8_t* prev = NULL; uint8_t* tmp = NULL; uint8_t* tmp2 = NULL;
    if (blurred) { tmp = (uint8_t*)malloc((size_t)w * h); tmp2 = (uint8_t*)malloc((size_t)w * h); }
    for (int s = 0; s < nlevels; s++) {
        uint8_t* plain = blurred ? (s & 1 ? tmp2 : tmp) : levels + off;
        if (s == 0) for (int y = 0; y < h; y++) memcpy(plain + (size_t)y * w, img + (size_t)y * pitch, (size_t)w);
        else efo_resize_linear(prev, ws[s - 1], hs[s - 1], (size_t)ws[s - 1], plain, ws[s], hs[s], (size_t)ws[s]);
        if (blurred) efo_gaussian_blur7(plain, ws[s], hs[s], (size_t)ws[s], levels + off, (size_t)ws[s]);
        prev = plain;
        off += (size_t)ws[s] * hs[s];
    }
    free(tmp); free(tmp2);
    return total;
}

/* ------------------------------------------------------------------------------------------- */
/* BAD: bad.cpp:86-103 (border test), :115-157 (rectifyBoxes), :166-251 (border response),        */
/* :320-405 (computeBAD).  cv::integral(CV_32S) restated as wrapping int32 prefix sums (H9).      */
/* ------------------------------------------------------------------------------------------- */
typedef struct { int x1, y1, x2, y2, r; } efo_box;

static void build_integral(const uint8_t* img, int w, int h, size_t pitch, uint32_t* I)
{
    const size_t iw = (size_t)w + 1;
    memset(I, 0, iw * sizeof(uint32_t));
    for (int y = 0; y < h; y++) {
        uint32_t rs = 0;
        uint32_t* cur = I + (size_t)(y + 1) * iw;
        const uint32_t* up = I + (size_t)y * iw;
        cur[0] = 0;
        for (int x = 0; x < w; x++) { rs += img[(size_t)y * pitch + x]; cur[x + 1] = up[x + 1] + rs; }
    }
}

static float bad_border_response(const efo_box* b, const uint32_t* I, int frameWidth, int frameHeight)
{
    /* bad.cpp:166-251; frameWidth/Height are the integral image dims (w+1, h+1) */
    int box1x1 = b->x1 - b->r; if (box1x1 < 0) box1x1 = 0; else if (box1x1 >= frameWidth - 1) box1x1 = frameWidth - 2;
    int box1y1 = b->y1 - b->r; if (box1y1 < 0) box1y1 = 0; else if (box1y1 >= frameHeight - 1) box1y1 = frameHeight - 2;
    int box1x2 = b->x1 + b->r + 1; if (box1x2 <= 0) box1x2 = 1; else if (box1x2 >= frameWidth) box1x2 = frameWidth - 1;
    int box1y2 = b->y1 + b->r + 1; if (box1y2 <= 0) box1y2 = 1; else if (box1y2 >= frameHeight) box1y2 = frameHeight - 1;
    int box2x1 = b->x2 - b->r; if (box2x1 < 0) box2x1 = 0; else if (box2x1 >= frameWidth - 1) box2x1 = frameWidth - 2;
    int box2y1 = b->y2 - b->r; if (box2y1 < 0) box2y1 = 0; else if (box2y1 >= frameHeight - 1) box2y1 = frameHeight - 2;
    int box2x2 = b->x2 + b->r + 1; if (box2x2 <= 0) box2x2 = 1; else if (box2x2 >= frameWidth) box2x2 = frameWidth - 1;
    int box2y2 = b->y2 + b->r + 1; if (box2y2 <= 0) box2y2 = 1; else if (box2y2 >= frameHeight) box2y2 = frameHeight - 1;
    const size_t iw = (size_t)frameWidth;
    uint32_t A = I[(size_t)box1y1 * iw + box1x1], B = I[(size_t)box1y1 * iw + box1x2];
    uint32_t C = I[(size_t)box1y2 * iw + box1x1], D = I[(size_t)box1y2 * iw + box1x2];
    const float sum1 = (float)(int32_t)(A + D - B - C);
    const int area1 = (box1y2 - box1y1) * (box1x2 - box1x1);
    const float avg1 = sum1 / (float)area1;
    A = I[(size_t)box2y1 * iw + box2x1]; B = I[(size_t)box2y1 * iw + box2x2];
    C = I[(size_t)box2y2 * iw + box2x1]; D = I[(size_t)box2y2 * iw + box2x2];
    const float sum2 = (float)(int32_t)(A + D - B - C);
    const int area2 = (box2y2 - box2y1) * (box2x2 - box2x1);
    const float avg2 = sum2 / (float)area2;
    return avg1 - avg2;
}

void efo_bad_compute(const uint8_t* img, int w, int h, size_t pitch, const efo_kpt* kpts, int n,
                     float scaleFactor, int nbits, uint8_t* desc)
{
    const unsigned char (*boxes)[5] = nbits == 512 ? ef_bad_boxes_512 : ef_bad_boxes_256;
    const unsigned int* thr_bits = nbits == 512 ? ef_bad_thresholds_512_bits : ef_bad_thresholds_256_bits;
    const size_t iw = (size_t)w + 1;
    uint32_t* I = (uint32_t*)malloc(iw * ((size_t)h + 1) * sizeof(uint32_t));
    build_integral(img, w, h, pitch, I);
    const int nbytes = nbits / 8;
    const int pw = 32, ph = 32; /* patch_size_ (bad.cpp:301) */

#pragma omp parallel for num_threads(g_threads) schedule(static)
    for (int ki = 0; ki < n; ki++) {
        const efo_kpt kp = kpts[ki];
        uint8_t* d = desc + (size_t)ki * nbytes;
        /* rectifyBoxes, bad.cpp:121-147 */
        float m00, m01, m02, m10, m11, m12;
        const float s = scaleFactor * kp.size / (0.5f * (float)(pw + ph));
        if (kp.angle == -1) {
            m00 = s; m01 = 0.0f; m02 = -0.5f * s * (float)pw + kp.x;
            m10 = 0.0f; m11 = s; m12 = -s * 0.5f * (float)ph + kp.y;
        } else {
            const float cosine = (kp.angle >= 0) ? (float)cos(kp.angle * 0.017453292519943295) : 1.f;
            const float sine = (kp.angle >= 0) ? (float)sin(kp.angle * 0.017453292519943295) : 0.f;
            m00 = s * cosine; m01 = -s * sine;
            m02 = (-s * cosine + s * sine) * (float)pw * 0.5f + kp.x;
            m10 = s * sine; m11 = s * cosine;
            m12 = (-s * sine - s * cosine) * (float)ph * 0.5f + kp.y;
        }
        /* isKeypointInTheBorder, bad.cpp:92-102 (frame = image size) */
        const float sb = scaleFactor * kp.size / (float)(pw + ph);
        const float bw = (float)pw * sb * 1.75f, bh = (float)ph * sb * 1.75f;
        int border = 0;
        if (kp.x < bw || kp.x + bw >= (float)w) border = 1;
        if (kp.y < bh || kp.y + bh >= (float)h) border = 1;

        uint8_t byte = 0;
        for (int i = 0; i < nbits; i++) {
            const float bx1 = (float)boxes[i][0], by1 = (float)boxes[i][1];
            const float bx2 = (float)boxes[i][2], by2 = (float)boxes[i][3];
            efo_box o;
            o.x1 = (int)(m00 * bx1 + m01 * by1 + m02 + 0.5f); /* bad.cpp:151-155, CV_ROUNDNUM truncates */
            o.y1 = (int)(m10 * bx1 + m11 * by1 + m12 + 0.5f);
            o.x2 = (int)(m00 * bx2 + m01 * by2 + m02 + 0.5f);
            o.y2 = (int)(m10 * bx2 + m11 * by2 + m12 + 0.5f);
            o.r = (int)(s * (float)boxes[i][4] + 0.5f);
            const float thr = u32_as_f32(thr_bits[i]);
            const int bit_idx = 7 - (i % 8);
            int bit;
            if (border) {
                const float resp = bad_border_response(&o, I, w + 1, h + 1);
                bit = resp <= thr; /* bad.cpp:352 */
            } else {
                /* bad.cpp:371-393 (index arithmetic in int on the reference; same elements here) */
                const int x1a = o.x1 - o.r, y1a = o.y1 - o.r, x1b = o.x1 + o.r + 1, y1b = o.y1 + o.r + 1;
                const int x2a = o.x2 - o.r, y2a = o.y2 - o.r, x2b = o.x2 + o.r + 1, y2b = o.y2 + o.r + 1;
                const int side = 1 + (o.r << 1);
                /* A keypoint closer to the edge than its largest box that still passes the (size-scaled) border test makes the reference
                 * read past its integral image: undefined there.  DEFINED here (and in the CUDA path): indices clamp to the last row /
                 * column of the integral image. */
#define EFO_I(y_, x_) I[(size_t)((y_) < 0 ? 0 : (y_) > h ? h : (y_)) * iw + ((x_) < 0 ? 0 : (x_) > w ? w : (x_))]
                const uint32_t acc = EFO_I(y1a, x1a) + EFO_I(y1b, x1b) - EFO_I(y1a, x1b) - EFO_I(y1b, x1a)
                                   - EFO_I(y2a, x2a) - EFO_I(y2b, x2b) + EFO_I(y2a, x2b) + EFO_I(y2b, x2a);
#undef EFO_I
                bit = (float)(int32_t)acc <= (thr * (float)(side * side));
            }
            byte |= (uint8_t)(bit << bit_idx);
            if (bit_idx == 0) { d[i / 8] = byte; byte = 0; }
        }
    }
    free(I);
}

/* ------------------------------------------------------------------------------------------- */
/* HashSIFT: hash_sift.cpp:68-109 (warpAffineLinear), :111-138 (rectifyPatch), :150-160           */
/* (normalize), :162-198 (HistBin, separateIF, distribute), :200-331 (computePatchSIFT),          */
/* :333-351, :353-378 (matmulAndSign).                                                            */
/* ------------------------------------------------------------------------------------------- */
void efo_hashsift_patch(const uint8_t* img, int w, int h, size_t pitch, const efo_kpt* kp,
                        float scaleFactor, uint8_t* patch)
{
    const float PI_1_0F = (float)3.1415926535897932384626433832795;
    const int pw = 32, ph = 32;
    const float s = scaleFactor * kp->size / (0.5f * (float)(pw + ph));
    const float theta = PI_1_0F * kp->angle / 180;
    const float cost = s * (kp->angle >= 0 ? cosf(theta) : 1.f);
    const float sint = s * (kp->angle >= 0 ? sinf(theta) : 0.f);
    const float M00 = +cost, M01 = -sint, M02 = (-cost + sint) * (float)pw / 2.f + kp->x;
    const float M10 = +sint, M11 = +cost, M12 = (-sint - cost) * (float)ph / 2.f + kp->y;
    for (int y = 0; y < ph; y++) {
        for (int x = 0; x < pw; x++) {
            const float u = M00 * (float)x + M01 * (float)y + M02;
            const float v = M10 * (float)x + M11 * (float)y + M12;
            uint8_t dstVal = 0;
            const int ui = cv_floor_f(u);
            const int vi = cv_floor_f(v);
            if (ui >= 0 && ui + 1 < w && vi >= 0 && vi + 1 < h) {
                const uint8_t* p = img + (size_t)vi * pitch + ui;
                const float du = u - (float)ui;
                const float dv = v - (float)vi;
                const float tmp0 = (1 - du) * (float)p[0] + du * (float)p[1];
                const float tmp1 = (1 - du) * (float)p[pitch] + du * (float)p[pitch + 1];
                const float tmp2 = (1 - dv) * tmp0 + dv * tmp1;
                int q = (int)(tmp2 + 0.5f);
                dstVal = (uint8_t)(q < 255 ? q : 255);
            }
            patch[y * pw + x] = dstVal;
        }
    }
}

static inline float squared(float x) { return x * x; }
static inline float normsq(float x, float y) { return squared(x) + squared(y); }

static void sift_normalize(float* desc, int size)
{
    float sum = 0;
    for (int i = 0; i < size; i++) sum += squared(desc[i]);
    const float nrm = fmaxf(sqrtf(sum), FLT_EPSILON);
    const float scale = 1.f / nrm;
    for (int i = 0; i < size; i++) desc[i] *= scale;
}

static void patch_sift(const uint8_t* img, float* descriptors)
{
    /* computePatchSIFT(patch 32x32, kpScale = 1.f/6), hash_sift.cpp:200-331 */
    const int h = 32, w = 32, dh = h - 2, dw = w - 2;
    const float kpScale = 1.f / 6;
    const float kpRadius = kpScale * (float)h * 0.5f;
    const float kernelSigma = 0.5f * 4 * 3.f * kpRadius;
    const float distScale = -1.f / (2 * kernelSigma * kernelSigma);
    const float cx = 0.5f * (float)dw, cy = 0.5f * (float)dh;
    const float PI_2_0F = (float)6.283185307179586476925286766559;

    float hist[6][6][10];
    memset(hist, 0, sizeof(hist));

    /* HistBin, :164-177 */
    const float cellh = 3.f * (kpScale * (float)h * 0.5f);
    const float cellw = 3.f * (kpScale * (float)w * 0.5f);
    const float scaleR = 1.f / cellh, scaleC = 1.f / cellw, scaleO = 8 / PI_2_0F;
    const float halfh = 0.5f * (float)h, halfw = 0.5f * (float)w;
    const float rbin0 = 4 / 2 - 0.5f, cbin0 = 4 / 2 - 0.5f;

    for (int y = 0; y < dh; y++) {
        const uint8_t* pT = img + (y + 0) * w + 1;
        const uint8_t* pC = img + (y + 1) * w + 1;
        const uint8_t* pB = img + (y + 2) * w + 1;
        const float rb = scaleR * ((float)(y + 1) - halfh) + rbin0;
        const int ri = cv_floor_f(rb);
        const float rf = rb - (float)ri;
        for (int x = 0; x < dw; x++) {
            const float magScale = expf(distScale * normsq((float)x - cx, (float)y - cy));
            const float dx = (float)(pC[x + 1] - pC[x - 1]);
            const float dy = (float)(pT[x] - pB[x]);
            const float mag = magScale * sqrtf(normsq(dx, dy));
            const float ori = atan2f(dy, dx);
            const float cb = scaleC * ((float)(x + 1) - halfw) + cbin0;
            const int ci = cv_floor_f(cb);
            const float cf = cb - (float)ci;
            const float ob = scaleO * ori;
            int oi = cv_floor_f(ob);
            const float of = ob - (float)oi;
            if (oi < 0) oi += 8;
            if (oi >= 8) oi -= 8;
            /* distribute(value, weight): v1 = weight*value; v0 = value - v1  (:193-198) */
            const float v1 = rf * mag, v0 = mag - v1;
            const float v01 = cf * v0, v00 = v0 - v01;
            const float v11 = cf * v1, v10 = v1 - v11;
            const float v001 = of * v00, v000 = v00 - v001;
            const float v011 = of * v01, v010 = v01 - v011;
            const float v101 = of * v10, v100 = v10 - v101;
            const float v111 = of * v11, v110 = v11 - v111;
            hist[ri + 1][ci + 1][oi + 0] += v000;
            hist[ri + 1][ci + 1][oi + 1] += v001;
            hist[ri + 1][ci + 2][oi + 0] += v010;
            hist[ri + 1][ci + 2][oi + 1] += v011;
            hist[ri + 2][ci + 1][oi + 0] += v100;
            hist[ri + 2][ci + 1][oi + 1] += v101;
            hist[ri + 2][ci + 2][oi + 0] += v110;
            hist[ri + 2][ci + 2][oi + 1] += v111;
        }
    }
    for (int r = 0; r < 4; r++)
        for (int c = 0; c < 4; c++) {
            float* ph_ = hist[r + 1][c + 1];
            ph_[0] += ph_[8];
            ph_[1] += ph_[9];
            for (int k = 0; k < 8; k++) descriptors[(r * 4 + c) * 8 + k] = ph_[k];
        }
    sift_normalize(descriptors, 128);
    for (int i = 0; i < 128; i++) descriptors[i] = fminf(descriptors[i], 0.2f);
    sift_normalize(descriptors, 128);
    for (int k = 0; k < 128; k++) descriptors[k] = (float)sat_u8_rne(512.f * descriptors[k]);
}

void efo_hashsift_features(const uint8_t* img, int w, int h, size_t pitch, const efo_kpt* kpts, int n,
                           float croppingScale, float* resp129)
{
#pragma omp parallel for num_threads(g_threads) schedule(static)
    for (int i = 0; i < n; i++) {
        uint8_t patch[1024];
        float* r = resp129 + (size_t)i * 129;
        r[0] = 1;
        efo_hashsift_patch(img, w, h, pitch, &kpts[i], croppingScale, patch);
        patch_sift(patch, r + 1);
    }
}

void efo_hashsift_project(const float* resp129, int n, int nbits, uint8_t* desc, float* proj)
{
    const unsigned int* wb = nbits == 512 ? ef_hashsift_w512_bits : ef_hashsift_w256_bits;
    const int nbytes = nbits / 8;
#pragma omp parallel for num_threads(g_threads) schedule(static)
    for (int i = 0; i < n; i++) {
        const float* a = resp129 + (size_t)i * 129;
        for (int b = 0; b < nbytes; b++) {
            uint8_t byte = 0;
            for (int j = 0; j < 8; j++) {
                const unsigned int* wr = wb + (size_t)(b * 8 + j) * 129;
                double acc = 0.0;
                for (int k = 0; k < 129; k++) acc += (double)a[k] * (double)u32_as_f32(wr[k]);
                const float t = (float)acc; /* cv::gemm fp32 output, double accumulation */
                if (proj) proj[(size_t)i * nbits + b * 8 + j] = t;
                byte |= (uint8_t)((t > 0) << (7 - j));
            }
            desc[(size_t)i * nbytes + b] = byte;
        }
    }
}

void efo_hashsift_compute(const uint8_t* img, int w, int h, size_t pitch, const efo_kpt* kpts, int n,
                          float croppingScale, int nbits, uint8_t* desc, float* proj)
{
    float* resp = (float*)malloc(sizeof(float) * 129 * (size_t)(n > 0 ? n : 1));
    efo_hashsift_features(img, w, h, pitch, kpts, n, croppingScale, resp);
    efo_hashsift_project(resp, n, nbits, desc, proj);
    free(resp);
}

/* ------------------------------------------------------------------------------------------- */
/* detect / detectAndCompute: cuda_efficient_features.cpp:225-321                                 */
/* ------------------------------------------------------------------------------------------- */
static int desc_bytes(int t) { return (t == EFO_BAD_256 || t == EFO_HASH_SIFT_256) ? 32 : 64; }

static int detect_impl(const uint8_t* img, int w, int h, size_t pitch, const efo_params* p,
                       efo_keypoint* out, uint8_t* desc, int cap, long* level_counts, int want_desc)
{
    const int L = p->nlevels;
    int ws[EFO_MAX_LEVELS], hs[EFO_MAX_LEVELS], quotas[EFO_MAX_LEVELS]; float sc[EFO_MAX_LEVELS];
    efo_level_geometry(w, h, p->scale_factor, L, ws, hs, sc);
    efo_level_quotas(p->nfeatures, p->scale_factor, L, quotas);

    uint8_t* cur = (uint8_t*)malloc((size_t)w * h);
    uint8_t* nxt = (uint8_t*)malloc((size_t)w * h);
    uint8_t* blur = want_desc ? (uint8_t*)malloc((size_t)w * h) : NULL;
    float* resp = (float*)malloc(sizeof(float) * (size_t)w * h);
    for (int y = 0; y < h; y++) memcpy(cur + (size_t)y * w, img + (size_t)y * pitch, (size_t)w);

    const int dbytes = desc_bytes(p->desc_type);
    int total = 0;
    for (int s = 0; s < L; s++) {
        const int lw = ws[s], lh = hs[s];
        if (s > 0) {
            efo_resize_linear(cur, ws[s - 1], hs[s - 1], (size_t)ws[s - 1], nxt, lw, lh, (size_t)lw);
            uint8_t* t = cur; cur = nxt; nxt = t;
        }
        if (level_counts) level_counts[3 * s] = level_counts[3 * s + 1] = level_counts[3 * s + 2] = 0;
        if (s < p->first_level) continue; /* :244 */

        const long ncorner = efo_score_map(cur, lw, lh, (size_t)lw, p->fast_threshold, resp);
        const long scap = ncorner > 0 ? ncorner : 1;
        short* xs = (short*)malloc(sizeof(short) * scap);
        short* ys = (short*)malloc(sizeof(short) * scap);
        float* rs = (float*)malloc(sizeof(float) * scap);
        const long nsurv = efo_radius_nms(resp, lw, lh, p->nonmax_radius, xs, ys, rs, scap);
        efo_cand* c = (efo_cand*)malloc(sizeof(efo_cand) * (nsurv > 0 ? nsurv : 1));
        for (long i = 0; i < nsurv; i++) { c[i].r = rs[i]; c[i].x = xs[i]; c[i].y = ys[i]; }
        long nsel = nsurv;
        if (nsurv > quotas[s]) { /* limitPoints, cuda_efficient_features.cu:344-358 */
            qsort(c, (size_t)nsurv, sizeof(efo_cand), cmp_resp_desc);
            nsel = quotas[s];
            qsort(c, (size_t)nsel, sizeof(efo_cand), cmp_raster);
        }
        if (level_counts) { level_counts[3 * s] = ncorner; level_counts[3 * s + 1] = nsurv; level_counts[3 * s + 2] = nsel; }

        if (nsel > cap - total) nsel = cap - total;
        efo_kpt* dk = (efo_kpt*)malloc(sizeof(efo_kpt) * (nsel > 0 ? nsel : 1));
#pragma omp parallel for num_threads(g_threads) schedule(static)
        for (long i = 0; i < nsel; i++) {
            efo_keypoint* k = &out[total + i];
            const float ang = efo_ic_angle(cur, (size_t)lw, c[i].x, c[i].y);   /* calcAngles :269 */
            k->lx = c[i].x; k->ly = c[i].y;
            k->response = c[i].r;
            k->angle = ang;
            /* scalePoints, cuda_efficient_features.cu:236-248 (nvcc fuses scale*x+0.5f) */
            k->x = (short)(int)fmaf(sc[s], (float)c[i].x, 0.5f);
            k->y = (short)(int)fmaf(sc[s], (float)c[i].y, 0.5f);
            k->octave = s;
            k->size = sc[s] * 31.f;
            /* convertKeypointsKernel :250-263: (x, y, PATCH_SIZE, angle) in level coordinates */
            dk[i].x = (float)c[i].x; dk[i].y = (float)c[i].y; dk[i].size = 31.f; dk[i].angle = ang;
        }
        if (want_desc && nsel > 0) {
            efo_gaussian_blur7(cur, lw, lh, (size_t)lw, blur, (size_t)lw); /* :305 */
            uint8_t* d = desc + (size_t)total * dbytes;
            switch (p->desc_type) { /* createDescriber :48-69: scale 1 */
            case EFO_BAD_256: efo_bad_compute(blur, lw, lh, (size_t)lw, dk, (int)nsel, 1.f, 256, d); break;
            case EFO_BAD_512: efo_bad_compute(blur, lw, lh, (size_t)lw, dk, (int)nsel, 1.f, 512, d); break;
            case EFO_HASH_SIFT_256: efo_hashsift_compute(blur, lw, lh, (size_t)lw, dk, (int)nsel, 1.f, 256, d, NULL); break;
            default: efo_hashsift_compute(blur, lw, lh, (size_t)lw, dk, (int)nsel, 1.f, 512, d, NULL); break;
            }
        }
        total += (int)nsel;
        free(dk); free(c); free(xs); free(ys); free(rs);
    }
    free(cur); free(nxt); free(blur); free(resp);
    return total;
}

int efo_detect(const uint8_t* img, int w, int h, size_t pitch, const efo_params* p,
               efo_keypoint* out, int cap, long* level_counts)
{
    return detect_impl(img, w, h, pitch, p, out, NULL, cap, level_counts, 0);
}

int efo_detect_and_compute(const uint8_t* img, int w, int h, size_t pitch, const efo_params* p,
                           efo_keypoint* out, uint8_t* desc, int cap, long* level_counts)
{
    return detect_impl(img, w, h, pitch, p, out, desc, cap, level_counts, 1);
}

/* ------------------------------------------------------------------------------------------- */
/* synthetic frames: i.i.d. uniform bytes from a counter-based hash                   */
/* ------------------------------------------------------------------------------------------- */
static inline uint32_t lowbias32(uint32_t x)
{
    x ^= x >> 16; x *= 0x7feb352dU; x ^= x >> 15; x *= 0x846ca68bU; x ^= x >> 16;
    return x;
}

void efo_synth_frame(uint32_t seed, uint32_t frame, int w, int h, size_t pitch, uint8_t* dst)
{
#pragma omp parallel for num_threads(g_threads) schedule(static)
    for (int y = 0; y < h; y++)
        for (int x = 0; x < w; x++) {
            const uint32_t idx = (frame * (uint32_t)h + (uint32_t)y) * (uint32_t)w + (uint32_t)x;
            dst[(size_t)y * pitch + x] = (uint8_t)(lowbias32(seed ^ idx) >> 24);
        }
}
