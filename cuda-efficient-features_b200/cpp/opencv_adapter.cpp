// opencv_adapter.cpp -- the reference's OpenCV-facing classes on top of the C ABI (include/ef_b200.h).
//
// This is the file a maintainer adds to modules/cuda_efficient_features/src/ IN PLACE OF cuda_efficient_features.cpp, cuda_fast.cu,
// cuda_efficient_features.cu, cuda_bad.*, cuda_hash_sift.*, cuda_efficient_descriptors.cpp and device_buffer.*, keeping the
// reference's public headers (include/cuda_efficient_features.h, include/cuda_efficient_descriptors.h) UNCHANGED, and links
// libef_b200.so.  It defines every symbol those two headers declare:
//   cv::cuda::EfficientFeatures::create / ~EfficientFeatures          (replaces src/cuda_efficient_features.cpp:188-411)
//   cv::cuda::BAD::create                                              (replaces src/cuda_bad.cpp:36-101)
//   cv::cuda::HashSIFT::create                                         (replaces src/cuda_hash_sift.cpp:93-168)
//   cv::cuda::EfficientDescriptorsAsync::~EfficientDescriptorsAsync    (replaces src/cuda_efficient_descriptors.cpp:24-26)
// It needs <opencv2/core/cuda.hpp>: the real OpenCV (>= 4.6, with the cuda core module) on a maintainer's box.  In this repository,
// where OpenCV's C++ side is not installed, the test tree compiles it against a small OpenCV stand-in together with the reference's
// own UNMODIFIED tests/descriptor_test.cpp and samples/sample_benchmark.cpp, and tests/test_gpu_adapter.py runs those binaries on
// the GPU box.
#include <algorithm>
#include <vector>

#include <opencv2/core/cuda.hpp>
#include <opencv2/core/cuda_stream_accessor.hpp>
#include <opencv2/features2d.hpp>

#include "cuda_efficient_descriptors.h" // the reference's own headers
#include "cuda_efficient_features.h"
#include "ef_b200.h"

namespace cv
{
namespace cuda
{
namespace
{

// getInputMat, src/cuda_efficient_features.cpp:71-84: a cv::Mat is uploaded, a GpuMat is aliased, anything else is an error
GpuMat inputMat(InputArray src, GpuMat& staging, Stream& stream)
{
    switch (src.kind()) {
    case _InputArray::MAT: staging.upload(src, stream); return staging;
    case _InputArray::CUDA_GPU_MAT: return src.getGpuMat();
    default: CV_Error(Error::StsBadArg, "Unsupported");
    }
    return GpuMat();
}

// getOutputMat + the trailing download, :86-100,316-320: results go to a caller-owned GpuMat of the exact size, or to a cv::Mat
void deliver(const GpuMat& src, OutputArray dst, Stream& stream)
{
    switch (dst.kind()) {
    case _InputArray::CUDA_GPU_MAT: dst.create(src.rows, src.cols, src.type()); src.copyTo(dst.getGpuMatRef(), stream); break;
    case _InputArray::MAT: src.download(dst, stream); break;
    default: CV_Error(Error::StsBadArg, "Unsupported");
    }
}

int descBytes(int dtype) { return (dtype == EF_BAD_256 || dtype == EF_HASH_SIFT_256) ? 32 : 64; }

// One ef_handle that is (re)created when an image or keypoint set outgrows the planned workspace -- the grow-only policy of the
// reference's DeviceBuffer (src/device_buffer.cpp:42-52); parameter changes go through ef_set_param.
class Handle
{
public:
    explicit Handle(bool computeOnly)
    {
        ef_default_params(&prm);
        prm.max_width = 0; prm.max_height = 0; prm.max_keypoints = 0;
        prm.flags = computeOnly ? EF_FLAG_COMPUTE_ONLY : 0;
    }
    ~Handle() { ef_destroy(h); }
    Handle(const Handle&) = delete;
    Handle& operator=(const Handle&) = delete;

    void check(int rc) const
    {
        if (rc != EF_OK) CV_Error(rc == EF_ERR_BAD_ARG ? Error::StsBadArg : Error::GpuApiCallError, h ? ef_last_error_string(h) : "ef_b200 call failed");
    }
    void ensure(int w, int hh, int nkp)
    {
        if (h && w <= prm.max_width && hh <= prm.max_height && nkp <= prm.max_keypoints) return;
        if (h) { ef_destroy(h); h = nullptr; }
        prm.max_width = std::max({ prm.max_width, w, 32 }); prm.max_height = std::max({ prm.max_height, hh, 32 });
        prm.max_keypoints = std::max({ prm.max_keypoints, nkp, prm.nfeatures });
        prm.device = getDevice();
        if (ef_create(&prm, &h) != EF_OK) { h = nullptr; CV_Error(Error::GpuApiCallError, "ef_create failed (no CUDA device, out of memory or bad parameters)"); }
    }
    void set(int id, double v) { if (h) check(ef_set_param(h, id, v)); }

    ef_params prm;
    ef_handle* h = nullptr;
};

// computeBAD / computeHashSIFT (src/cuda_bad.cpp:46-70, src/cuda_hash_sift.cpp:113-137) for both keypoint kinds
class DescriberCore
{
public:
    DescriberCore(int dtype, float scale) : hd_(true) { hd_.prm.desc_type = dtype; hd_.prm.desc_scale = scale; hd_.prm.nfeatures = 1; }
    int descriptorSize() const { return descBytes(hd_.prm.desc_type); }

    // std::vector<KeyPoint>: packed as (pt.x, pt.y, size, angle), src/cuda_efficient_features.cpp:116-128
    void compute(InputArray image, const std::vector<KeyPoint>& keypoints, OutputArray descriptors, Stream& stream)
    {
        if (image.empty()) return;
        if (keypoints.empty()) { descriptors.release(); return; }
        CV_Assert(image.type() == CV_8U);
        const GpuMat img = inputMat(image, image_, stream);
        const int n = (int)keypoints.size();
        hd_.ensure(img.cols, img.rows, n);
        hkpts_.resize((size_t)n * 4);
        for (int i = 0; i < n; i++) {
            const KeyPoint& k = keypoints[i];
            hkpts_[4 * i] = k.pt.x; hkpts_[4 * i + 1] = k.pt.y; hkpts_[4 * i + 2] = k.size; hkpts_[4 * i + 3] = k.angle;
        }
        const Mat hk(1, n, CV_32FC4, hkpts_.data());   // ONE row: contiguous on the device whatever pitch the allocator gives multi-row matrices
        dkpts_.upload(hk, stream);
        run(img, n, descriptors, stream, [&](GpuMat& out) {
            return ef_compute_async(hd_.h, img.data, img.step, img.cols, img.rows, dkpts_.ptr<float>(), n, out.data, out.step, StreamAccessor::getStream(stream));
        });
    }
    // 5 x N matrix (Mat or GpuMat): only LOCATION and ANGLE rows are read, size is 31 (convertKeypointsKernel, src/cuda_efficient_features.cu:250-263)
    void computeRows(InputArray image, InputArray keypoints, OutputArray descriptors, Stream& stream)
    {
        if (image.empty()) return;
        if (keypoints.empty()) { descriptors.release(); return; }
        CV_Assert(image.type() == CV_8U);
        const GpuMat img = inputMat(image, image_, stream);
        const GpuMat k = inputMat(keypoints, dkpts_, stream);
        CV_Assert(k.rows == 5 && k.type() == CV_32F);
        const int n = k.cols;
        hd_.ensure(img.cols, img.rows, n);
        run(img, n, descriptors, stream, [&](GpuMat& out) {
            return ef_compute_rows_async(hd_.h, img.data, img.step, img.cols, img.rows, k.ptr<float>(), k.step, n, out.data, out.step, StreamAccessor::getStream(stream));
        });
    }

private:
    template <class F> void run(const GpuMat&, int n, OutputArray descriptors, Stream& stream, F call)
    {
        if (descriptors.kind() == _InputArray::CUDA_GPU_MAT) {
            descriptors.create(n, descriptorSize(), CV_8U);
            hd_.check(call(descriptors.getGpuMatRef()));
        } else if (descriptors.kind() == _InputArray::MAT) {
            desc_.create(n, descriptorSize(), CV_8U);
            hd_.check(call(desc_));
            desc_.download(descriptors, stream);
        } else CV_Error(Error::StsBadArg, "Unsupported");
    }
    Handle hd_;
    GpuMat image_, dkpts_, desc_;
    std::vector<float> hkpts_;
};

class BADB200 final : public BAD
{
public:
    BADB200(float scaleFactor, int nbits) : core_(nbits == SIZE_256_BITS ? EF_BAD_256 : EF_BAD_512, scaleFactor) {}
    void compute(InputArray image, std::vector<KeyPoint>& keypoints, OutputArray descriptors) override { core_.compute(image, keypoints, descriptors, Stream::Null()); }
    void computeAsync(InputArray image, InputArray keypoints, OutputArray descriptors, Stream& stream) override { core_.computeRows(image, keypoints, descriptors, stream); }
    int descriptorSize() const override { return core_.descriptorSize(); }
    int descriptorType() const override { return CV_8U; }
    int defaultNorm() const override { return NORM_HAMMING; }
private:
    DescriberCore core_;
};

class HashSIFTB200 final : public HashSIFT
{
public:
    HashSIFTB200(float croppingScale, int nbits) : core_(nbits == SIZE_256_BITS ? EF_HASH_SIFT_256 : EF_HASH_SIFT_512, croppingScale) {}
    void compute(InputArray image, std::vector<KeyPoint>& keypoints, OutputArray descriptors) override { core_.compute(image, keypoints, descriptors, Stream::Null()); }
    void computeAsync(InputArray image, InputArray keypoints, OutputArray descriptors, Stream& stream) override { core_.computeRows(image, keypoints, descriptors, stream); }
    int descriptorSize() const override { return core_.descriptorSize(); }
    int descriptorType() const override { return CV_8U; }
    int defaultNorm() const override { return NORM_HAMMING; }
private:
    DescriberCore core_;
};

class EfficientFeaturesB200 final : public EfficientFeatures
{
public:
    EfficientFeaturesB200(int nfeatures, float scaleFactor, int nlevels, int firstLevel, int fastThreshold, int nonmaxRadius, DescriptorType dtype) : hd_(false)
    {
        ef_params& p = hd_.prm;
        p.nfeatures = nfeatures; p.scale_factor = scaleFactor; p.nlevels = nlevels; p.first_level = firstLevel;
        p.fast_threshold = fastThreshold; p.nonmax_radius = nonmaxRadius; p.desc_type = (int)dtype;
    }

    // ---- Feature2D (src/cuda_efficient_features.cpp:197-213)
    void detect(InputArray image, std::vector<KeyPoint>& keypoints, InputArray mask) override
    {
        detectAsync(image, keypoints_, mask, Stream::Null());
        convert(keypoints_, keypoints);
    }
    void compute(InputArray image, std::vector<KeyPoint>& keypoints, OutputArray descriptors) override
    {
        describer().compute(image, keypoints, descriptors, Stream::Null());
    }
    void detectAndCompute(InputArray image, InputArray mask, std::vector<KeyPoint>& keypoints, OutputArray descriptors, bool useProvidedKeypoints) override
    {
        detectAndComputeAsync(image, mask, keypoints_, descriptors, useProvidedKeypoints, Stream::Null());
        convert(keypoints_, keypoints);
    }

    // ---- *Async (:215-321)
    void detectAsync(InputArray image, OutputArray keypoints, InputArray mask, Stream& stream) override
    {
        detectAndComputeAsync(image, mask, keypoints, noArray(), false, stream);
    }
    void computeAsync(InputArray image, InputArray keypoints, OutputArray descriptors, Stream& stream) override
    {
        describer().computeRows(image, keypoints, descriptors, stream);
    }
    void detectAndComputeAsync(InputArray image, InputArray /* mask: never read by the reference either, :225-250 */, OutputArray keypoints,
                               OutputArray descriptors, bool useProvidedKeypoints, Stream& stream) override
    {
        CV_Assert(image.type() == CV_8U);      // :228
        CV_Assert(!useProvidedKeypoints);      // :229
        const GpuMat img = inputMat(image, image_, stream);
        const ef_params& p = hd_.prm;
        hd_.ensure(img.cols, img.rows, p.nfeatures);
        const bool need = descriptors.needed();
        kfull_.create(ROWS_COUNT, p.nfeatures, CV_32F);
        if (need) dfull_.create(p.nfeatures, descriptorSize(), CV_8U);
        if (count_.empty()) count_.create(1, 1, CV_32S);
        cudaStream_t s = StreamAccessor::getStream(stream);
        hd_.check(ef_detect_and_compute_async(hd_.h, img.data, img.step, img.cols, img.rows, kfull_.ptr<float>(), kfull_.step,
                                              need ? dfull_.data : nullptr, need ? dfull_.step : 0, count_.ptr<int>(), s));
        // the ONE host synchronisation of the call: the outputs are created with their exact size, like the reference's (which
        // blocks twice per pyramid level to get there, src/cuda_fast.cu:241-243, src/cuda_efficient_features.cu:337-339)
        int n = 0;
        if (cudaMemcpyAsync(&n, count_.ptr<int>(), sizeof(int), cudaMemcpyDeviceToHost, s) != cudaSuccess || cudaStreamSynchronize(s) != cudaSuccess)
            CV_Error(Error::GpuApiCallError, "reading the keypoint count failed");
        if (n == 0) {                          // :275-281
            keypoints.release();
            if (need) descriptors.release();
            return;
        }
        deliver(kfull_.colRange(0, n), keypoints, stream);
        if (need) deliver(dfull_.rowRange(0, n), descriptors, stream);
    }

    // ---- convert (:323-349)
    void convert(InputArray src, std::vector<KeyPoint>& dst) override
    {
        if (src.empty()) { dst.clear(); return; }
        Mat tmp;
        if (src.kind() == _InputArray::CUDA_GPU_MAT) src.getGpuMat().download(tmp);
        else if (src.kind() == _InputArray::MAT) tmp = src.getMat();
        else CV_Error(Error::StsBadArg, "Unsupported");
        CV_Assert(tmp.rows == ROWS_COUNT && tmp.type() == CV_32F);
        const short* loc = tmp.ptr<short>(LOCATION_ROW);
        const float* resp = tmp.ptr<float>(RESPONSE_ROW);
        const float* ang = tmp.ptr<float>(ANGLE_ROW);
        const int* oct = tmp.ptr<int>(OCTAVE_ROW);
        const float* size = tmp.ptr<float>(SIZE_ROW);
        dst.resize((size_t)tmp.cols);
        for (int i = 0; i < tmp.cols; i++)
            dst[i] = KeyPoint(Point2f(loc[2 * i], loc[2 * i + 1]), size[i], ang[i], resp[i], oct[i]);
    }

    int descriptorSize() const override { return descBytes(hd_.prm.desc_type); }   // :351-353
    int descriptorType() const override { return CV_8U; }
    int defaultNorm() const override { return NORM_HAMMING; }

    // ---- the 7 setter / getter pairs (:355-377)
    void setMaxFeatures(int v) override { hd_.prm.nfeatures = v; hd_.set(EF_PARAM_MAX_FEATURES, v); kfull_.release(); dfull_.release(); }
    int getMaxFeatures() const override { return hd_.prm.nfeatures; }
    void setScaleFactor(float v) override { hd_.prm.scale_factor = v; hd_.set(EF_PARAM_SCALE_FACTOR, v); }
    float getScaleFactor() const override { return hd_.prm.scale_factor; }
    void setNLevels(int v) override { hd_.prm.nlevels = v; hd_.set(EF_PARAM_NLEVELS, v); }
    int getNLevels() const override { return hd_.prm.nlevels; }
    void setFirstLevel(int v) override { hd_.prm.first_level = v; hd_.set(EF_PARAM_FIRST_LEVEL, v); }
    int getFirstLevel() const override { return hd_.prm.first_level; }
    void setFastThreshold(int v) override { hd_.prm.fast_threshold = v; hd_.set(EF_PARAM_FAST_THRESHOLD, v); }
    int getFastThreshold() const override { return hd_.prm.fast_threshold; }
    void setNonmaxRadius(int v) override { hd_.prm.nonmax_radius = v; hd_.set(EF_PARAM_NONMAX_RADIUS, v); }
    int getNonmaxRadius() const override { return hd_.prm.nonmax_radius; }
    void setDescriptorType(DescriptorType v) override
    {   // the reference rebuilds its describer here (:373-377)
        hd_.prm.desc_type = (int)v; hd_.set(EF_PARAM_DESCRIPTOR_TYPE, (int)v);
        describer_.reset(); dfull_.release();
    }
    DescriptorType getDescriptorType() const override { return (DescriptorType)hd_.prm.desc_type; }

private:
    // compute()/computeAsync() of the reference forward to its BAD / HashSIFT object created with scale 1 (:48-69,220-223)
    DescriberCore& describer()
    {
        if (!describer_) describer_.reset(new DescriberCore(hd_.prm.desc_type, 1.f));
        return *describer_;
    }
    Handle hd_;
    std::unique_ptr<DescriberCore> describer_;
    GpuMat image_, keypoints_, kfull_, dfull_, count_;
};

} // namespace

Ptr<EfficientFeatures> EfficientFeatures::create(int nfeatures, float scaleFactor, int nlevels, int firstLevel, int fastThreshold,
                                                 int nonmaxRadius, DescriptorType dtype)
{
    return makePtr<EfficientFeaturesB200>(nfeatures, scaleFactor, nlevels, firstLevel, fastThreshold, nonmaxRadius, dtype);
}
EfficientFeatures::~EfficientFeatures() {}

EfficientDescriptorsAsync::~EfficientDescriptorsAsync() {}

Ptr<BAD> BAD::create(float scaleFactor, int nbits)
{
    return makePtr<BADB200>(scaleFactor, nbits);   // like the reference (src/cuda_bad.cpp:38-40): anything but SIZE_256_BITS means 512 bits
}

Ptr<HashSIFT> HashSIFT::create(float croppingScale, int nbits)
{
    if (nbits != SIZE_256_BITS && nbits != SIZE_512_BITS)                                       // src/cuda_hash_sift.cpp:100-105
        CV_Error(Error::StsBadArg, "n_bits should be either SIZE_512_BITS or SIZE_256_BITS");
    return makePtr<HashSIFTB200>(croppingScale, nbits);
}

} // namespace cuda
} // namespace cv
