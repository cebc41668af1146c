// ef_features.hpp -- C++ host mirror of the reference's operator interface over the C ABI (include/ef_b200.h).
//
// Same names, argument meaning and error behaviour as cv::cuda::EfficientFeatures
// (modules/cuda_efficient_features/include/cuda_efficient_features.h:28-98) and cv::cuda::BAD / HashSIFT
// (cuda_efficient_descriptors.h:27-121), with POD views instead of OpenCV types because OpenCV is not part
// of this build.  cpp/opencv_adapter.cpp shows the thin layer that turns this into a drop-in
// cv::cuda::EfficientFeatures where OpenCV (core, cuda) is installed.
#pragma once

#include <cstdint>
#include <cstring>
#include <memory>
#include <stdexcept>
#include <string>
#include <vector>

#include "../../include/ef_b200.h"

namespace efb200 {

// stands in for cv::Exception (CV_Assert / CV_Error)
struct Error : std::runtime_error { int status; Error(int s, const std::string& m) : std::runtime_error(m), status(s) {} };

// the four fields of cv::cuda::GpuMat / cv::Mat the path touches
struct MatView { void* data = nullptr; size_t step = 0; int rows = 0, cols = 0; };

// cv::KeyPoint fields filled by convert() (cuda_efficient_features.cpp:323-349)
struct KeyPoint { float x, y, size, angle, response; int octave; };

struct Capacity { int max_width = 3840, max_height = 2160, max_batch = 1, max_keypoints = 0, device = 0, flags = 0; };

class EfficientFeatures
{
public:
    static constexpr int LOCATION_ROW = EF_LOCATION_ROW, RESPONSE_ROW = EF_RESPONSE_ROW, ANGLE_ROW = EF_ANGLE_ROW,
                         OCTAVE_ROW = EF_OCTAVE_ROW, SIZE_ROW = EF_SIZE_ROW, ROWS_COUNT = EF_ROWS_COUNT;
    enum DescriptorType { BAD_256 = EF_BAD_256, BAD_512 = EF_BAD_512, HASH_SIFT_256 = EF_HASH_SIFT_256, HASH_SIFT_512 = EF_HASH_SIFT_512 };

    static std::unique_ptr<EfficientFeatures> create(int nfeatures = 5000, float scaleFactor = 1.2f, int nlevels = 8, int firstLevel = 0,
                                                     int fastThreshold = 20, int nonmaxRadius = 15, DescriptorType dtype = HASH_SIFT_256,
                                                     const Capacity& cap = Capacity())
    {
        ef_params p;
        ef_default_params(&p);
        p.nfeatures = nfeatures; p.scale_factor = scaleFactor; p.nlevels = nlevels; p.first_level = firstLevel;
        p.fast_threshold = fastThreshold; p.nonmax_radius = nonmaxRadius; p.desc_type = dtype;
        p.max_width = cap.max_width; p.max_height = cap.max_height; p.max_batch = cap.max_batch;
        p.max_keypoints = cap.max_keypoints; p.device = cap.device; p.flags = cap.flags;
        ef_handle* h = nullptr;
        const int rc = ef_create(&p, &h);
        if (rc != EF_OK) throw Error(rc, "ef_create failed (no CUDA device, bad argument or out of memory); there is no CPU fallback");
        return std::unique_ptr<EfficientFeatures>(new EfficientFeatures(h));
    }

    ~EfficientFeatures() { ef_destroy(h_); }
    EfficientFeatures(const EfficientFeatures&) = delete;
    EfficientFeatures& operator=(const EfficientFeatures&) = delete;

    // detectAndComputeAsync over device buffers (cuda_efficient_features.h:72-73).  keypoints: 5 x nfeatures CV_32F,
    // descriptors: nfeatures x descriptorSize() CV_8U or data == nullptr (detectAsync, :60); d_count: device int.
    // `mask` is accepted and ignored, like the reference (cuda_efficient_features.cpp:225-250).
    void detectAndComputeAsync(const MatView& image, const MatView& /*mask*/, const MatView& keypoints, const MatView& descriptors,
                               int* d_count, bool useProvidedKeypoints = false, void* stream = nullptr)
    {
        if (useProvidedKeypoints) throw Error(EF_ERR_BAD_ARG, "useProvidedKeypoints must be false");
        check(ef_detect_and_compute_async(h_, static_cast<const uint8_t*>(image.data), image.step, image.cols, image.rows,
                                          static_cast<float*>(keypoints.data), keypoints.step, static_cast<uint8_t*>(descriptors.data),
                                          descriptors.step, d_count, stream));
    }
    void detectAsync(const MatView& image, const MatView& keypoints, int* d_count, void* stream = nullptr)
    {
        detectAndComputeAsync(image, MatView(), keypoints, MatView(), d_count, false, stream);
    }
    // computeAsync with a 5 x N keypoint matrix on the device (size forced to 31, cuda_efficient_features.cu:250-263)
    void computeAsync(const MatView& image, const MatView& keypoints5xN, const MatView& descriptors, void* stream = nullptr)
    {
        check(ef_compute_rows_async(h_, static_cast<const uint8_t*>(image.data), image.step, image.cols, image.rows,
                                    static_cast<const float*>(keypoints5xN.data), keypoints5xN.step, keypoints5xN.cols,
                                    static_cast<uint8_t*>(descriptors.data), descriptors.step, stream));
    }

    // Feature2D-shaped synchronous calls on HOST images (cv::Mat path, cuda_efficient_features.cpp:197-213)
    void detectAndCompute(const MatView& hostImage, std::vector<KeyPoint>& keypoints, std::vector<uint8_t>* descriptors, void* stream = nullptr)
    {
        const int nf = getMaxFeatures(), db = descriptorSize();
        kbuf_.resize((size_t)ROWS_COUNT * nf);
        if (descriptors) descriptors->resize((size_t)nf * db);
        int n = 0;
        check(ef_detect_and_compute_host(h_, static_cast<const uint8_t*>(hostImage.data), hostImage.step, hostImage.cols, hostImage.rows,
                                         kbuf_.data(), descriptors ? descriptors->data() : nullptr, &n, stream));
        if (descriptors) descriptors->resize((size_t)n * db);
        convert(kbuf_.data(), (size_t)nf * sizeof(float), n, keypoints);
    }
    void detect(const MatView& hostImage, std::vector<KeyPoint>& keypoints, void* stream = nullptr) { detectAndCompute(hostImage, keypoints, nullptr, stream); }

    // convert(): 5 x N host matrix -> vector<KeyPoint> (cuda_efficient_features.cpp:323-349)
    static void convert(const float* kpts5, size_t pitchBytes, int n, std::vector<KeyPoint>& dst)
    {
        const uint8_t* base = reinterpret_cast<const uint8_t*>(kpts5);
        dst.resize((size_t)n);
        for (int i = 0; i < n; i++) {
            int16_t xy[2];
            std::memcpy(xy, base + LOCATION_ROW * pitchBytes + 4 * (size_t)i, 4);
            KeyPoint k;
            k.x = xy[0]; k.y = xy[1];
            std::memcpy(&k.response, base + RESPONSE_ROW * pitchBytes + 4 * (size_t)i, 4);
            std::memcpy(&k.angle, base + ANGLE_ROW * pitchBytes + 4 * (size_t)i, 4);
            std::memcpy(&k.octave, base + OCTAVE_ROW * pitchBytes + 4 * (size_t)i, 4);
            std::memcpy(&k.size, base + SIZE_ROW * pitchBytes + 4 * (size_t)i, 4);
            dst[(size_t)i] = k;
        }
    }

    int descriptorSize() const { return ef_descriptor_size(h_); }
    int descriptorType() const { return 0; /* CV_8U */ }
    int defaultNorm() const { return 6; /* NORM_HAMMING */ }

    void setMaxFeatures(int v) { set(EF_PARAM_MAX_FEATURES, v); }
    int getMaxFeatures() const { return (int)get(EF_PARAM_MAX_FEATURES); }
    void setScaleFactor(float v) { set(EF_PARAM_SCALE_FACTOR, v); }
    float getScaleFactor() const { return (float)get(EF_PARAM_SCALE_FACTOR); }
    void setNLevels(int v) { set(EF_PARAM_NLEVELS, v); }
    int getNLevels() const { return (int)get(EF_PARAM_NLEVELS); }
    void setFirstLevel(int v) { set(EF_PARAM_FIRST_LEVEL, v); }
    int getFirstLevel() const { return (int)get(EF_PARAM_FIRST_LEVEL); }
    void setFastThreshold(int v) { set(EF_PARAM_FAST_THRESHOLD, v); }
    int getFastThreshold() const { return (int)get(EF_PARAM_FAST_THRESHOLD); }
    void setNonmaxRadius(int v) { set(EF_PARAM_NONMAX_RADIUS, v); }
    int getNonmaxRadius() const { return (int)get(EF_PARAM_NONMAX_RADIUS); }
    void setDescriptorType(DescriptorType v) { set(EF_PARAM_DESCRIPTOR_TYPE, v); }
    DescriptorType getDescriptorType() const { return (DescriptorType)(int)get(EF_PARAM_DESCRIPTOR_TYPE); }

    ef_handle* handle() const { return h_; }

private:
    explicit EfficientFeatures(ef_handle* h) : h_(h) {}
    void check(int rc) const { if (rc != EF_OK) throw Error(rc, ef_last_error_string(h_)); }
    void set(int id, double v) { check(ef_set_param(h_, id, v)); }
    double get(int id) const { double v = 0; check(ef_get_param(h_, id, &v)); return v; }
    ef_handle* h_;
    std::vector<float> kbuf_;
};

// cv::cuda::BAD / cv::cuda::HashSIFT (cuda_efficient_descriptors.h:67-121): compute-only describers with a scale
class Describer
{
public:
    enum Size { SIZE_512_BITS = 100, SIZE_256_BITS = 101 };
    // keypoints: n x (x, y, size, angle) floats on the device (the std::vector<KeyPoint> path,
    // cuda_efficient_features.cpp:116-128); descriptors: n x descriptorSize() bytes on the device
    void computeAsync(const MatView& image, const float* d_kpts_xysa, int n, const MatView& descriptors, void* stream = nullptr)
    {
        const int rc = ef_compute_async(ef_->handle(), static_cast<const uint8_t*>(image.data), image.step, image.cols, image.rows,
                                        d_kpts_xysa, n, static_cast<uint8_t*>(descriptors.data), descriptors.step, stream);
        if (rc != EF_OK) throw Error(rc, ef_last_error_string(ef_->handle()));
    }
    int descriptorSize() const { return ef_->descriptorSize(); }
    int descriptorType() const { return 0; }
    int defaultNorm() const { return 6; }

protected:
    Describer(EfficientFeatures::DescriptorType t, float scale, const Capacity& cap)
        : ef_(EfficientFeatures::create(1, 1.2f, 8, 0, 20, 15, t, compute_only(cap)))   // tables + per-keypoint scratch only, like the reference describers
    {
        if (ef_set_param(ef_->handle(), EF_PARAM_DESC_SCALE, scale) != EF_OK) throw Error(EF_ERR_BAD_ARG, "bad scale");
    }
    static Capacity compute_only(Capacity c) { c.flags |= EF_FLAG_COMPUTE_ONLY; return c; }
    std::unique_ptr<EfficientFeatures> ef_;
};

class BAD : public Describer
{
public:
    static std::unique_ptr<BAD> create(float scaleFactor, int nbits = SIZE_256_BITS, const Capacity& cap = Capacity())
    {
        if (nbits != SIZE_512_BITS && nbits != SIZE_256_BITS) throw Error(EF_ERR_BAD_ARG, "n_bits should be either SIZE_512_BITS or SIZE_256_BITS");
        return std::unique_ptr<BAD>(new BAD(nbits == SIZE_512_BITS ? EfficientFeatures::BAD_512 : EfficientFeatures::BAD_256, scaleFactor, cap));
    }
private:
    using Describer::Describer;
};

class HashSIFT : public Describer
{
public:
    static std::unique_ptr<HashSIFT> create(float croppingScale, int nbits = SIZE_256_BITS, const Capacity& cap = Capacity())
    {
        if (nbits != SIZE_512_BITS && nbits != SIZE_256_BITS) throw Error(EF_ERR_BAD_ARG, "n_bits should be either SIZE_512_BITS or SIZE_256_BITS");
        return std::unique_ptr<HashSIFT>(new HashSIFT(nbits == SIZE_512_BITS ? EfficientFeatures::HASH_SIFT_512 : EfficientFeatures::HASH_SIFT_256, croppingScale, cap));
    }
private:
    using Describer::Describer;
};

// cv::BFMatcher for NORM_HAMMING over descriptors that stay on the device (callers: samples/sample_image_sequence.cpp:81,115-137,
// samples/sample_feature_matching.cpp:99-101).  All buffers are the caller's (cv::cuda::GpuMat in an OpenCV build):
//   knnMatchAsync: idx / dist = nq x k CV_32S (k = 1 or 2), row q = the k nearest train rows in OpenCV's order, -1 where missing
//   matchAsync   : trainIdx / dist = nq CV_32S; with crossCheck, trainIdx is -1 for the queries OpenCV would omit
//   scratch      : scratchBytes(nq, nt) bytes, 16-byte aligned
class BFMatcher
{
public:
    explicit BFMatcher(int normType = 6 /* NORM_HAMMING */, bool crossCheck = false) : crossCheck_(crossCheck)
    {
        if (normType != 6) throw Error(EF_ERR_UNSUPPORTED, "only NORM_HAMMING (defaultNorm() of the path's descriptors) is implemented");
    }
    static std::unique_ptr<BFMatcher> create(int normType = 6, bool crossCheck = false) { return std::unique_ptr<BFMatcher>(new BFMatcher(normType, crossCheck)); }
    static size_t scratchBytes(int nq, int nt) { return ef_match_scratch_bytes(nq, nt); }

    void knnMatchAsync(const MatView& query, const MatView& train, int k, int* d_idx, int* d_dist, void* d_scratch, void* stream = nullptr) const
    {
        check(ef_match_knn_async(static_cast<const uint8_t*>(query.data), query.step, query.rows, static_cast<const uint8_t*>(train.data), train.step,
                                 train.rows, query.cols, k, d_idx, d_dist, d_scratch, stream));
    }
    void matchAsync(const MatView& query, const MatView& train, int* d_trainIdx, int* d_dist, void* d_scratch, void* stream = nullptr) const
    {
        if (crossCheck_)
            check(ef_match_cross_check_async(static_cast<const uint8_t*>(query.data), query.step, query.rows, static_cast<const uint8_t*>(train.data),
                                             train.step, train.rows, query.cols, d_trainIdx, d_dist, d_scratch, stream));
        else knnMatchAsync(query, train, 1, d_trainIdx, d_dist, d_scratch, stream);
    }
    // the ratio + cross-check loop of samples/sample_image_sequence.cpp:121-137 over two k = 2 results: d_out[q] = train row or -1
    static void ratioCrossFilterAsync(const int* d_idx12, const int* d_dist12, int nq, const int* d_idx21, const int* d_dist21, int nt,
                                      double uniqueness, int* d_out, void* stream = nullptr)
    {
        check(ef_match_ratio_cross_async(d_idx12, d_dist12, nq, d_idx21, d_dist21, nt, uniqueness, d_out, stream));
    }

private:
    static void check(int rc) { if (rc != EF_OK) throw Error(rc, ef_match_last_error_string()); }
    bool crossCheck_;
};

// convertToGray (samples/sample_common.cpp:35-45) on the device: 8UC3 (BGR) / 8UC4 (BGRA) -> 8UC1; 8UC1 passes through
inline void convertToGrayAsync(const MatView& src, int channels, const MatView& dstGray, void* stream = nullptr)
{
    if (channels == 1) throw Error(EF_ERR_BAD_ARG, "8UC1 input needs no conversion: use it as is");
    const int rc = ef_bgr_to_gray_async(static_cast<const uint8_t*>(src.data), src.step, src.cols, src.rows, channels,
                                        static_cast<uint8_t*>(dstGray.data), dstGray.step, stream);
    if (rc != EF_OK) throw Error(rc, "Image should be 8UC1, 8UC3 or 8UC4");   // CV_Error(StsBadArg, ...), sample_common.cpp:44
}

} // namespace efb200
