// opencv_adapter.cpp -- drop-in cv::cuda::EfficientFeatures on top of the C ABI.
//
// NOT compiled in this repository's build (OpenCV's C++ headers are not installed here); it is the file a
// maintainer adds to modules/cuda_efficient_features/src/ IN PLACE OF cuda_efficient_features.cpp /
// cuda_fast.cu / cuda_efficient_features.cu / cuda_bad.* / cuda_hash_sift.* while keeping the reference's
// public headers unchanged, then links libef_b200.so.  See INTEGRATION.md.
#if __has_include(<opencv2/core/cuda.hpp>)
#include <opencv2/core/cuda.hpp>
#include <opencv2/core/cuda_stream_accessor.hpp>
#include <opencv2/features2d.hpp>

#include "cuda_efficient_features.h" // the reference's own header
#include "ef_b200.h"

namespace cv { namespace cuda {

class EfficientFeaturesB200 : public EfficientFeatures
{
public:
    EfficientFeaturesB200(int nfeatures, float scaleFactor, int nlevels, int firstLevel, int fastThreshold, int nonmaxRadius, DescriptorType dtype)
    {
        ef_default_params(&prm_);
        prm_.nfeatures = nfeatures; prm_.scale_factor = scaleFactor; prm_.nlevels = nlevels; prm_.first_level = firstLevel;
        prm_.fast_threshold = fastThreshold; prm_.nonmax_radius = nonmaxRadius; prm_.desc_type = (int)dtype;
        prm_.max_width = 0; prm_.max_height = 0; // sized lazily from the first image
        prm_.device = getDevice();
        count_.create(1, 1, CV_32S);
    }
    ~EfficientFeaturesB200() override { ef_destroy(h_); }

    void detect(InputArray image, std::vector<KeyPoint>& keypoints, InputArray mask) override
    { detectAsync(image, keypoints_, mask, Stream::Null()); convert(keypoints_, keypoints); }
    void compute(InputArray image, std::vector<KeyPoint>& keypoints, OutputArray descriptors) override
    {
        if (keypoints.empty()) { descriptors.release(); return; }
        GpuMat img = upload(image, Stream::Null());
        ensure(img.cols, img.rows, (int)keypoints.size());
        Mat k((int)keypoints.size(), 1, CV_32FC4);
        for (int i = 0; i < k.rows; i++) k.at<Vec4f>(i) = Vec4f(keypoints[i].pt.x, keypoints[i].pt.y, keypoints[i].size, keypoints[i].angle);
        GpuMat dk(k), dd(k.rows, descriptorSize(), CV_8U);
        check(ef_compute_async(h_, img.data, img.step, img.cols, img.rows, dk.ptr<float>(), k.rows, dd.data, dd.step, nullptr));
        dd.download(descriptors);
    }
    void detectAndCompute(InputArray image, InputArray mask, std::vector<KeyPoint>& keypoints, OutputArray descriptors, bool useProvided) override
    { detectAndComputeAsync(image, mask, keypoints_, descriptors, useProvided, Stream::Null()); convert(keypoints_, keypoints); }
    void detectAsync(InputArray image, OutputArray keypoints, InputArray mask, Stream& stream) override
    { detectAndComputeAsync(image, mask, keypoints, noArray(), false, stream); }
    void computeAsync(InputArray image, InputArray keypoints, OutputArray descriptors, Stream& stream) override
    {
        GpuMat img = upload(image, stream), k = upload(keypoints, stream);
        CV_Assert(k.rows == 5 && k.type() == CV_32F);
        if (k.cols == 0) { descriptors.release(); return; }
        ensure(img.cols, img.rows, k.cols);
        GpuMat dd = output(descriptors, k.cols, descriptorSize(), CV_8U);
        check(ef_compute_rows_async(h_, img.data, img.step, img.cols, img.rows, k.ptr<float>(), k.step, k.cols, dd.data, dd.step,
                                    StreamAccessor::getStream(stream)));
        if (descriptors.kind() == _InputArray::MAT) dd.download(descriptors, stream);
    }
    void detectAndComputeAsync(InputArray image, InputArray, OutputArray keypoints, OutputArray descriptors, bool useProvided, Stream& stream) override
    {
        CV_Assert(image.type() == CV_8U);
        CV_Assert(!useProvided);
        GpuMat img = upload(image, stream);
        ensure(img.cols, img.rows, prm_.nfeatures);
        const bool need = descriptors.needed();
        kfull_.create(ROWS_COUNT, prm_.nfeatures, CV_32F);
        if (need) dfull_.create(prm_.nfeatures, descriptorSize(), CV_8U);
        cudaStream_t s = StreamAccessor::getStream(stream);
        check(ef_detect_and_compute_async(h_, img.data, img.step, img.cols, img.rows, kfull_.ptr<float>(), kfull_.step,
                                          need ? dfull_.data : nullptr, need ? dfull_.step : 0, count_.ptr<int>(), s));
        // the ONE synchronisation needed to give the outputs their exact size (the reference does 16 per frame)
        int n = 0;
        cudaMemcpyAsync(&n, count_.ptr<int>(), sizeof(int), cudaMemcpyDeviceToHost, s);
        cudaStreamSynchronize(s);
        if (n == 0) { keypoints.release(); if (need) descriptors.release(); return; }
        deliver(kfull_.colRange(0, n), keypoints, stream);
        if (need) deliver(dfull_.rowRange(0, n), descriptors, stream);
    }
    void convert(InputArray src, std::vector<KeyPoint>& dst) override
    {
        Mat tmp; if (src.kind() == _InputArray::MAT) tmp = src.getMat(); else src.getGpuMat().download(tmp);
        dst.resize(tmp.cols);
        for (int i = 0; i < tmp.cols; i++) {
            const Vec2s p = tmp.ptr<Vec2s>(LOCATION_ROW)[i];
            dst[i] = KeyPoint(Point2f(p[0], p[1]), tmp.ptr<float>(SIZE_ROW)[i], tmp.ptr<float>(ANGLE_ROW)[i],
                              tmp.ptr<float>(RESPONSE_ROW)[i], tmp.ptr<int>(OCTAVE_ROW)[i]);
        }
    }
    int descriptorSize() const override { return (prm_.desc_type == BAD_256 || prm_.desc_type == HASH_SIFT_256) ? 32 : 64; }
    int descriptorType() const override { return CV_8U; }
    int defaultNorm() const override { return NORM_HAMMING; }
#define EF_ACCESSOR(Name, T, field, id) void set##Name(T v) override { prm_.field = v; if (h_) check(ef_set_param(h_, id, (double)v)); } T get##Name() const override { return (T)prm_.field; }
    EF_ACCESSOR(MaxFeatures, int, nfeatures, EF_PARAM_MAX_FEATURES) EF_ACCESSOR(ScaleFactor, float, scale_factor, EF_PARAM_SCALE_FACTOR)
    EF_ACCESSOR(NLevels, int, nlevels, EF_PARAM_NLEVELS) EF_ACCESSOR(FirstLevel, int, first_level, EF_PARAM_FIRST_LEVEL)
    EF_ACCESSOR(FastThreshold, int, fast_threshold, EF_PARAM_FAST_THRESHOLD) EF_ACCESSOR(NonmaxRadius, int, nonmax_radius, EF_PARAM_NONMAX_RADIUS)
#undef EF_ACCESSOR
    void setDescriptorType(DescriptorType v) override { prm_.desc_type = (int)v; if (h_) check(ef_set_param(h_, EF_PARAM_DESCRIPTOR_TYPE, (double)v)); }
    DescriptorType getDescriptorType() const override { return (DescriptorType)prm_.desc_type; }

private:
    void check(int rc) { if (rc != EF_OK) CV_Error(rc == EF_ERR_BAD_ARG ? Error::StsBadArg : Error::GpuApiCallError, ef_last_error_string(h_)); }
    void ensure(int w, int h, int nkp)
    {   // (re)create the handle when the image outgrows the planned workspace -- mirrors DeviceBuffer's grow-only policy
        if (h_ && w <= prm_.max_width && h <= prm_.max_height && nkp <= prm_.max_keypoints) return;
        if (h_) ef_destroy(h_);
        prm_.max_width = std::max(prm_.max_width, w); prm_.max_height = std::max(prm_.max_height, h);
        prm_.max_keypoints = std::max(prm_.max_keypoints, nkp);
        if (ef_create(&prm_, &h_) != EF_OK) CV_Error(Error::GpuApiCallError, "ef_create failed");
    }
    static GpuMat upload(InputArray a, Stream& s)
    {
        if (a.kind() == _InputArray::CUDA_GPU_MAT) return a.getGpuMat();
        if (a.kind() != _InputArray::MAT) CV_Error(Error::StsBadArg, "Unsupported");
        GpuMat d; d.upload(a, s); return d;
    }
    static GpuMat output(OutputArray a, int rows, int cols, int type)
    {
        if (a.kind() == _InputArray::CUDA_GPU_MAT) { a.create(rows, cols, type); return a.getGpuMat(); }
        return GpuMat(rows, cols, type);
    }
    static void deliver(const GpuMat& src, OutputArray dst, Stream& s)
    {
        if (dst.kind() == _InputArray::CUDA_GPU_MAT) { dst.create(src.size(), src.type()); src.copyTo(dst.getGpuMatRef(), s); }
        else src.download(dst, s);
    }
    ef_params prm_; ef_handle* h_ = nullptr;
    GpuMat keypoints_, kfull_, dfull_, count_;
};

Ptr<EfficientFeatures> EfficientFeatures::create(int nfeatures, float scaleFactor, int nlevels, int firstLevel, int fastThreshold,
                                                 int nonmaxRadius, DescriptorType dtype)
{
    return makePtr<EfficientFeaturesB200>(nfeatures, scaleFactor, nlevels, firstLevel, fastThreshold, nonmaxRadius, dtype);
}
EfficientFeatures::~EfficientFeatures() {}

}} // namespace cv::cuda
#endif
