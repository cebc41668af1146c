// Minimal stand-in for <opencv2/features2d.hpp> (see core.hpp in this directory). Not OpenCV code.
#pragma once
#include "core.hpp"
namespace cv
{
// cv::Feature2D with OpenCV's signatures and defaults: detect / compute forward to detectAndCompute unless overridden
class Feature2D
{
public:
    virtual ~Feature2D() {}
    virtual void detect(InputArray image, std::vector<KeyPoint>& keypoints, InputArray mask = noArray())
    { detectAndCompute(image, mask, keypoints, noArray(), false); }
    virtual void compute(InputArray image, std::vector<KeyPoint>& keypoints, OutputArray descriptors)
    { detectAndCompute(image, noArray(), keypoints, descriptors, true); }
    virtual void detectAndCompute(InputArray, InputArray, std::vector<KeyPoint>&, OutputArray, bool = false)
    { CV_Error(Error::StsBadArg, "Feature2D::detectAndCompute is not implemented"); }
    virtual int descriptorSize() const { return 0; }
    virtual int descriptorType() const { return 0; }
    virtual int defaultNorm() const { return 0; }
};
// drawing helpers of samples/sample_common.cpp: no rasteriser here -- the output is the (first) input image, enough for the samples to reach
// their cv::imshow call (the stand-in for that prints the image size)
struct DrawMatchesFlags { enum { DEFAULT = 0, DRAW_RICH_KEYPOINTS = 4 }; };
// (InputArray / OutputArray parameters as in OpenCV: the samples' own ::drawKeypoints(const Mat&, ...) must stay the better overload)
static inline void drawKeypoints(InputArray img, const std::vector<KeyPoint>&, OutputArray out, const Scalar& = Scalar::all(-1), int = 0)
{ img.getMat().copyTo(out.getMatRef()); }
static inline void drawMatches(InputArray img1, const std::vector<KeyPoint>&, InputArray, const std::vector<KeyPoint>&, const std::vector<DMatch>&, OutputArray out)
{ img1.getMat().copyTo(out.getMatRef()); }
} // namespace cv

// cv::BFMatcher(NORM_HAMMING) as the reference's samples use it (sample_feature_matching.cpp:99-101: create(NORM_HAMMING, true)->match;
// sample_image_sequence.cpp:81,115-116: create(defaultNorm())->knnMatch k = 2), routed to the library's matcher through the C ABI
// (ef_match_cross_check_async / ef_match_knn_async): the substitution INTEGRATION.md proposes for the step after the path.
// Compiled only into the adapter binaries (-DEF_SHIM_WITH_MATCHER, oracle/Makefile): the CPU-only builds of the reference's descriptor sources
// include this header too and have neither the CUDA runtime nor include/ on their path.
#ifdef EF_SHIM_WITH_MATCHER
#include <cuda_runtime.h>
#include "ef_b200.h"
namespace cv
{
class BFMatcher
{
public:
    BFMatcher(int normType = NORM_HAMMING, bool crossCheck = false) : cross_(crossCheck) { CV_Assert(normType == NORM_HAMMING); }
    static Ptr<BFMatcher> create(int normType = NORM_HAMMING, bool crossCheck = false) { return makePtr<BFMatcher>(normType, crossCheck); }
    void match(const Mat& query, const Mat& train, std::vector<DMatch>& matches) const
    {
        std::vector<int> idx, dist;
        run(query, train, 1, cross_, idx, dist);
        matches.clear();
        for (int q = 0; q < query.rows; q++)
            if (idx[q] >= 0) { DMatch m; m.queryIdx = q; m.trainIdx = idx[q]; m.imgIdx = 0; m.distance = (float)dist[q]; matches.push_back(m); }
    }
    void knnMatch(const Mat& query, const Mat& train, std::vector<std::vector<DMatch>>& matches, int k) const
    {
        CV_Assert(k == 1 || k == 2);
        std::vector<int> idx, dist;
        run(query, train, k, false, idx, dist);
        matches.assign((size_t)query.rows, std::vector<DMatch>());
        for (int q = 0; q < query.rows; q++)
            for (int j = 0; j < k; j++)
                if (idx[(size_t)q * k + j] >= 0) {
                    DMatch m; m.queryIdx = q; m.trainIdx = idx[(size_t)q * k + j]; m.imgIdx = 0; m.distance = (float)dist[(size_t)q * k + j];
                    matches[q].push_back(m);
                }
    }
private:
    static void chk(cudaError_t e) { if (e != cudaSuccess) CV_Error(Error::GpuApiCallError, cudaGetErrorString(e)); }
    static void run(const Mat& query, const Mat& train, int k, bool cross, std::vector<int>& idx, std::vector<int>& dist)
    {
        CV_Assert(query.type() == CV_8UC1 && train.type() == CV_8UC1 && query.cols == train.cols);
        const int nq = query.rows, nt = train.rows, nb = query.cols;
        idx.assign((size_t)nq * k, -1); dist.assign((size_t)nq * k, 0);
        if (nq == 0 || nt == 0) return;
        uint8_t *dq = nullptr, *dt = nullptr; int *di = nullptr, *dd = nullptr; void* scratch = nullptr;
        chk(cudaMalloc(&dq, (size_t)nq * nb)); chk(cudaMalloc(&dt, (size_t)nt * nb));
        chk(cudaMalloc(&di, sizeof(int) * nq * k)); chk(cudaMalloc(&dd, sizeof(int) * nq * k));
        chk(cudaMalloc(&scratch, ef_match_scratch_bytes(nq, nt)));
        chk(cudaMemcpy2D(dq, nb, query.data, (size_t)query.step, nb, nq, cudaMemcpyHostToDevice));
        chk(cudaMemcpy2D(dt, nb, train.data, (size_t)train.step, nb, nt, cudaMemcpyHostToDevice));
        const int rc = cross ? ef_match_cross_check_async(dq, nb, nq, dt, nb, nt, nb, di, dd, scratch, nullptr)
                             : ef_match_knn_async(dq, nb, nq, dt, nb, nt, nb, k, di, dd, scratch, nullptr);
        if (rc != 0) CV_Error(Error::StsBadArg, ef_match_last_error_string());
        chk(cudaMemcpy(idx.data(), di, sizeof(int) * nq * k, cudaMemcpyDeviceToHost));
        chk(cudaMemcpy(dist.data(), dd, sizeof(int) * nq * k, cudaMemcpyDeviceToHost));
        cudaFree(dq); cudaFree(dt); cudaFree(di); cudaFree(dd); cudaFree(scratch);
    }
    bool cross_;
};
} // namespace cv
#endif // EF_SHIM_WITH_MATCHER
