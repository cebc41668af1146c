// ref_cuda_api.cu -- TEST INFRASTRUCTURE: C entry points over the reference's own CUDA detector kernels, compiled UNMODIFIED from
// /root/reference/modules/cuda_efficient_features/src/{cuda_fast, cuda_efficient_features, cuda_hash_sift, cuda_bad}.cu against oracle/shim_cuda (default nvcc
// flags = the reference's build, modules/cuda_efficient_features/CMakeLists.txt:22-29: contraction on).  Used by tests/ only, to pin the
// CPU restatement of the detector (oracle/ef_oracle.c) and the product kernels against what the reference really computes:
// FAST-9 corner set, Harris responses, radius NMS, limitPoints, IC angles, scalePoints.
#include <cstdint>
#include <cstring>

#include <cuda_fast.cu>                    // -I<reference>/modules/cuda_efficient_features/src
#include <cuda_efficient_features.cu>
#include <cuda_hash_sift.cu>             // the reference's GPU HashSIFT kernels (approximately equal to its CPU descriptors: tests/descriptor_test.cpp:48-75)

#include <cuda_bad.cu>                   // the reference's GPU BAD kernel (float cos/sin: approximately equal to its CPU descriptors, tests/descriptor_test.cpp:19-46)

#include <cublas_v2.h>

using namespace cv;
using namespace cv::cuda;

namespace {
struct DevBuf {
    void* p = nullptr;
    explicit DevBuf(size_t bytes) { cudaMalloc(&p, bytes ? bytes : 16); cudaMemset(p, 0, bytes ? bytes : 16); }
    ~DevBuf() { cudaFree(p); }
};
}

extern "C" {

// createMask (cuda_efficient_features.cpp:176-182) + calcKeypoints (cuda_fast.cu:224-246).  Returns the number of corners found
// (capped at maxpoints like the reference); xy = maxpoints x 2 shorts.
int efrefcu_fast(const uint8_t* h_img, int w, int h, int threshold, int border, int maxpoints, short* h_xy)
{
    DevBuf img((size_t)w * h), mask((size_t)w * h), kp(sizeof(float) * 4 * (size_t)maxpoints), cnt(16);
    unsigned int* h_cnt = nullptr;
    cudaMallocHost((void**)&h_cnt, 16);
    cudaMemcpy(img.p, h_img, (size_t)w * h, cudaMemcpyHostToDevice);
    {
        std::vector<uint8_t> m((size_t)w * h, 0);
        for (int y = border; y < h - border; y++) std::memset(&m[(size_t)y * w + border], 255, (size_t)std::max(0, w - 2 * border));
        cudaMemcpy(mask.p, m.data(), m.size(), cudaMemcpyHostToDevice);
    }
    GpuMat gimg(h, w, CV_8UC1, img.p, (size_t)w), gmask(h, w, CV_8UC1, mask.p, (size_t)w);
    GpuMat gkp(4, maxpoints, CV_32F, kp.p, sizeof(float) * (size_t)maxpoints), gcnt(1, 4, CV_32S, cnt.p, 16);
    HostMem hcnt(1, 4, h_cnt, 16);
    calcKeypoints(gimg, gmask, gkp, maxpoints, threshold, gcnt, hcnt, 0);
    const int n = gkp.cols;
    cudaMemcpy(h_xy, kp.p, sizeof(short) * 2 * (size_t)n, cudaMemcpyDeviceToHost);
    cudaFreeHost(h_cnt);
    return n;
}

// calcResponses (:360-374) and calcAngles (:376-390) on given points (n x 2 shorts)
void efrefcu_responses_angles(const uint8_t* h_img, int w, int h, const short* h_xy, int n, float* h_resp, float* h_angle)
{
    if (n <= 0) return;
    DevBuf img((size_t)w * h), pts(sizeof(float) * 5 * (size_t)n);
    cudaMemcpy(img.p, h_img, (size_t)w * h, cudaMemcpyHostToDevice);
    cudaMemcpy(pts.p, h_xy, sizeof(short) * 2 * (size_t)n, cudaMemcpyHostToDevice);
    GpuMat gimg(h, w, CV_8UC1, img.p, (size_t)w), gp(5, n, CV_32F, pts.p, sizeof(float) * (size_t)n);
    calcResponses(gimg, gp, 0);
    calcAngles(gimg, gp, 0);
    cudaDeviceSynchronize();
    if (h_resp) cudaMemcpy(h_resp, gp.ptr<float>(1), sizeof(float) * (size_t)n, cudaMemcpyDeviceToHost);
    if (h_angle) cudaMemcpy(h_angle, gp.ptr<float>(2), sizeof(float) * (size_t)n, cudaMemcpyDeviceToHost);
}

// radiusSuppression (:281-342) then, if maxpoints >= 0, limitPoints (:344-358).  In: n points + responses; out: survivors.
int efrefcu_nms_limit(const short* h_xy, const float* h_resp, int n, int w, int h, float radius, int maxpoints, short* o_xy, float* o_resp)
{
    if (n <= 0) return 0;
    DevBuf src(sizeof(float) * 5 * (size_t)n), dst(sizeof(float) * 5 * (size_t)n);
    const int bufInts = radiusSuppressionBufferSize(Size(w, h), n);
    DevBuf buf(sizeof(int) * (size_t)bufInts);
    int* h_cnt = nullptr;
    cudaMallocHost((void**)&h_cnt, 16);
    cudaMemcpy(src.p, h_xy, sizeof(short) * 2 * (size_t)n, cudaMemcpyHostToDevice);
    cudaMemcpy((char*)src.p + sizeof(float) * (size_t)n, h_resp, sizeof(float) * (size_t)n, cudaMemcpyHostToDevice);
    GpuMat gsrc(5, n, CV_32F, src.p, sizeof(float) * (size_t)n), gdst(5, n, CV_32F, dst.p, sizeof(float) * (size_t)n);
    GpuMat gbuf(1, bufInts, CV_32S, buf.p, sizeof(int) * (size_t)bufInts);
    HostMem hcnt(1, 4, h_cnt, 16);
    radiusSuppression(gsrc, gdst, Size(w, h), radius, gbuf, hcnt, 0);
    if (maxpoints >= 0) limitPoints(gdst, maxpoints, 0);
    cudaDeviceSynchronize();
    const int m = gdst.cols;
    cudaMemcpy(o_xy, gdst.ptr<short2>(0), sizeof(short) * 2 * (size_t)m, cudaMemcpyDeviceToHost);
    cudaMemcpy(o_resp, gdst.ptr<float>(1), sizeof(float) * (size_t)m, cudaMemcpyDeviceToHost);
    cudaFreeHost(h_cnt);
    return m;
}

// scalePoints (:392-407): in-place on n points; returns scaled xy, octave and size rows
void efrefcu_scale(const short* h_xy, int n, float scale, int octave, short* o_xy, int* o_octave, float* o_size)
{
    if (n <= 0) return;
    DevBuf pts(sizeof(float) * 5 * (size_t)n);
    cudaMemcpy(pts.p, h_xy, sizeof(short) * 2 * (size_t)n, cudaMemcpyHostToDevice);
    GpuMat gp(5, n, CV_32F, pts.p, sizeof(float) * (size_t)n);
    scalePoints(gp, scale, octave, 0);
    cudaDeviceSynchronize();
    cudaMemcpy(o_xy, gp.ptr<short2>(0), sizeof(short) * 2 * (size_t)n, cudaMemcpyDeviceToHost);
    cudaMemcpy(o_octave, gp.ptr<int>(3), sizeof(int) * (size_t)n, cudaMemcpyDeviceToHost);
    cudaMemcpy(o_size, gp.ptr<float>(4), sizeof(float) * (size_t)n, cudaMemcpyDeviceToHost);
}

// The per-level detector sequence of EfficientFeaturesImpl::detectAndComputeAsync (cuda_efficient_features.cpp:244-272,310) on the
// reference's own kernels, for the level images of ONE frame (the pyramid itself is cv::cuda::resize: third-party, not timed):
// calcKeypoints (capacity 0.1 * area, :35,252) -> calcResponses -> radiusSuppression -> limitPoints -> calcAngles -> scalePoints.
// Buffers are allocated once like the reference's DeviceBuffers; returns the mean milliseconds per frame over `iters` (wall clock around
// the stream, like samples/sample_benchmark.cpp) and the keypoints of the last iteration per level in n_out.
float efrefcu_time_detect_levels(const uint8_t* const* h_levels, const int* ws, const int* hs, const float* scales, const int* quotas, int nlevels,
                                 int threshold, float radius, int iters, int* n_out)
{
    struct Lv { uint8_t *img, *mask; float *tmp, *kpts; int* buf; int maxpoints, bufInts; };
    std::vector<Lv> lv(nlevels);
    unsigned int* h_cnt = nullptr;
    cudaMallocHost((void**)&h_cnt, 16);
    for (int l = 0; l < nlevels; l++) {
        const int w = ws[l], h = hs[l];
        Lv& L = lv[l];
        L.maxpoints = cvRound(0.1 * (double)w * h);
        L.bufInts = radiusSuppressionBufferSize(Size(w, h), L.maxpoints);
        cudaMalloc((void**)&L.img, (size_t)w * h); cudaMalloc((void**)&L.mask, (size_t)w * h);
        cudaMalloc((void**)&L.tmp, sizeof(float) * 4 * (size_t)L.maxpoints); cudaMalloc((void**)&L.kpts, sizeof(float) * 5 * (size_t)L.maxpoints);
        cudaMalloc((void**)&L.buf, sizeof(int) * (size_t)L.bufInts);
        cudaMemcpy(L.img, h_levels[l], (size_t)w * h, cudaMemcpyHostToDevice);
        std::vector<uint8_t> m((size_t)w * h, 0);
        for (int y = 15; y < h - 15; y++) std::memset(&m[(size_t)y * w + 15], 255, (size_t)std::max(0, w - 30));
        cudaMemcpy(L.mask, m.data(), m.size(), cudaMemcpyHostToDevice);
    }
    cudaStream_t st;
    cudaStreamCreate(&st);
    double total_ms = 0;
    for (int it = 0; it <= iters; it++) {
        cudaStreamSynchronize(st);
        cudaEvent_t e0, e1;
        cudaEventCreate(&e0); cudaEventCreate(&e1);
        cudaEventRecord(e0, st);
        for (int l = 0; l < nlevels; l++) {
            Lv& L = lv[l];
            const int w = ws[l], h = hs[l];
            GpuMat image(h, w, CV_8UC1, L.img, (size_t)w), mask(h, w, CV_8UC1, L.mask, (size_t)w);
            GpuMat tmppoints(4, L.maxpoints, CV_32F, L.tmp, sizeof(float) * (size_t)L.maxpoints);
            GpuMat keypoints(5, L.maxpoints, CV_32F, L.kpts, sizeof(float) * (size_t)L.maxpoints);
            GpuMat d_buffer(L.bufInts, 1, CV_32S, L.buf, sizeof(int));
            HostMem h_buffer(1, 4, h_cnt, 16);
            calcKeypoints(image, mask, tmppoints, L.maxpoints, threshold, d_buffer, h_buffer, st);
            calcResponses(image, tmppoints, st);
            radiusSuppression(tmppoints, keypoints, image.size(), radius, d_buffer, h_buffer, st);
            limitPoints(keypoints, quotas[l], st);
            calcAngles(image, keypoints, st);
            scalePoints(keypoints, scales[l], l, st);
            n_out[l] = keypoints.cols;
        }
        cudaEventRecord(e1, st);
        cudaEventSynchronize(e1);
        float ms = 0;
        cudaEventElapsedTime(&ms, e0, e1);
        if (it > 0) total_ms += ms;          // one discarded warm-up iteration (sample_benchmark.cpp:39-52)
        cudaEventDestroy(e0); cudaEventDestroy(e1);
    }
    for (Lv& L : lv) { cudaFree(L.img); cudaFree(L.mask); cudaFree(L.tmp); cudaFree(L.kpts); cudaFree(L.buf); }
    cudaStreamDestroy(st);
    cudaFreeHost(h_cnt);
    return (float)(total_ms / iters);
}

// The reference's GPU HashSIFT (cuda_hash_sift.cpp:113-137): computePatchSIFTs -> cublasSgemm (hashSIFTGemm, :44-60) -> binarizeDescriptors,
// on n keypoints (x, y, size, angle) of one image.  Returns the mean milliseconds over `iters` (one discarded warm-up) and the descriptors.
float efrefcu_time_hashsift(const uint8_t* h_img, int w, int h, const float* h_kpts4, int n, int nbits, float croppingScale, int iters, uint8_t* h_desc)
{
#include "hash_sift.p512.h"
#include "hash_sift.p256.h"
    if (n <= 0 || (nbits != 256 && nbits != 512)) return -1.f;
    const double* vals = nbits == 512 ? HASH_SIFT_512_VALS : HASH_SIFT_256_VALS;
    std::vector<float> B((size_t)nbits * 129);
    for (size_t i = 0; i < B.size(); i++) B[i] = (float)vals[i];                 // Mat(...CV_64F...).convertTo(bMatrix_, CV_32F), :104-106
    DevBuf img((size_t)w * h), kp(sizeof(float) * 4 * (size_t)n), resp(sizeof(float) * 129 * (size_t)n), tmp(sizeof(float) * (size_t)nbits * n),
           bm(sizeof(float) * B.size()), desc((size_t)n * (nbits / 8));
    cudaMemcpy(img.p, h_img, (size_t)w * h, cudaMemcpyHostToDevice);
    cudaMemcpy(kp.p, h_kpts4, sizeof(float) * 4 * (size_t)n, cudaMemcpyHostToDevice);
    cudaMemcpy(bm.p, B.data(), sizeof(float) * B.size(), cudaMemcpyHostToDevice);
    GpuMat gimg(h, w, CV_8UC1, img.p, (size_t)w), gk(n, 1, CV_32FC4, kp.p, sizeof(float) * 4);
    GpuMat gresp(n, 129, CV_32F, resp.p, sizeof(float) * 129), gtmp(n, nbits, CV_32F, tmp.p, sizeof(float) * (size_t)nbits);
    GpuMat gB(nbits, 129, CV_32F, bm.p, sizeof(float) * 129), gdesc(n, nbits / 8, CV_8UC1, desc.p, (size_t)(nbits / 8));
    cublasHandle_t handle;
    cublasCreate_v2(&handle);
    cublasSetPointerMode_v2(handle, CUBLAS_POINTER_MODE_HOST);
    cudaStream_t st;
    cudaStreamCreate(&st);
    cublasSetStream_v2(handle, st);
    double total = 0;
    for (int it = 0; it <= iters; it++) {
        cudaEvent_t e0, e1;
        cudaEventCreate(&e0); cudaEventCreate(&e1);
        cudaEventRecord(e0, st);
        gpu::computePatchSIFTs(gimg, gk, gresp, croppingScale, 1.f / 6, 1.6, st);
        const float alphaf = 1.0f, betaf = 0.0f;
        cublasSgemm_v2(handle, CUBLAS_OP_T, CUBLAS_OP_N, gB.rows, gresp.rows, gB.cols, &alphaf, gB.ptr<float>(), (int)(gB.step / sizeof(float)),
                       gresp.ptr<float>(), (int)(gresp.step / sizeof(float)), &betaf, gtmp.ptr<float>(), (int)(gtmp.step / sizeof(float)));
        gpu::binarizeDescriptors(gtmp, gdesc, st);
        cudaEventRecord(e1, st);
        cudaEventSynchronize(e1);
        float ms = 0;
        cudaEventElapsedTime(&ms, e0, e1);
        if (it > 0) total += ms;
        cudaEventDestroy(e0); cudaEventDestroy(e1);
    }
    if (h_desc) cudaMemcpy(h_desc, desc.p, (size_t)n * (nbits / 8), cudaMemcpyDeviceToHost);
    cublasDestroy_v2(handle);
    cudaStreamDestroy(st);
    return (float)(total / iters);
}

// The reference's GPU BAD (cuda_bad.cpp:46-70): loadBoxPairParams + computeBAD on n keypoints (x, y, size, angle), given the integral image
// ((h + 1) x (w + 1) int32, exact prefix sums computed by the caller: cudev's integral is third-party and not part of the timing).
float efrefcu_time_bad(const int* h_integral, int w, int h, const float* h_kpts4, int n, int nbits, float scaleFactor, int iters, uint8_t* h_desc)
{
    if (n <= 0 || (nbits != 256 && nbits != 512)) return -1.f;
    DevBuf integ(sizeof(int) * (size_t)(w + 1) * (h + 1)), kp(sizeof(float) * 4 * (size_t)n), desc((size_t)n * (nbits / 8));
    cudaMemcpy(integ.p, h_integral, sizeof(int) * (size_t)(w + 1) * (h + 1), cudaMemcpyHostToDevice);
    cudaMemcpy(kp.p, h_kpts4, sizeof(float) * 4 * (size_t)n, cudaMemcpyHostToDevice);
    GpuMat gint(h + 1, w + 1, CV_32S, integ.p, sizeof(int) * (size_t)(w + 1)), gk(n, 1, CV_32FC4, kp.p, sizeof(float) * 4);
    GpuMat gdesc(n, nbits / 8, CV_8UC1, desc.p, (size_t)(nbits / 8));
    gpu::loadBoxPairParams(nbits);
    cudaStream_t st;
    cudaStreamCreate(&st);
    double total = 0;
    for (int it = 0; it <= iters; it++) {
        cudaEvent_t e0, e1;
        cudaEventCreate(&e0); cudaEventCreate(&e1);
        cudaEventRecord(e0, st);
        gpu::computeBAD(gint, gk, gdesc, scaleFactor, nbits, Size(32, 32), st);
        cudaEventRecord(e1, st);
        cudaEventSynchronize(e1);
        float ms = 0;
        cudaEventElapsedTime(&ms, e0, e1);
        if (it > 0) total += ms;
        cudaEventDestroy(e0); cudaEventDestroy(e1);
    }
    if (h_desc) cudaMemcpy(h_desc, desc.p, (size_t)n * (nbits / 8), cudaMemcpyDeviceToHost);
    cudaStreamDestroy(st);
    return (float)(total / iters);
}

} // extern "C"
