// C entry points over the reference's own CPU descriptors (TEST INFRASTRUCTURE).
// The two reference sources are compiled UNMODIFIED from where they lie under /root/reference:
// hash_sift.cpp is #included so that its file-static stages (computePatchSIFTs, matmulAndSign)
// can be probed; bad.cpp is a separate translation unit (see oracle/Makefile).
#include "hash_sift.cpp" // resolved with -I/root/reference/modules/efficient_features/src

#include <cstdint>

struct efo_kpt { float x, y, size, angle; };

static cv::Mat wrap(const uint8_t* img, int w, int h, size_t pitch) { return cv::Mat(h, w, CV_8UC1, (void*)img, pitch); }
static std::vector<cv::KeyPoint> to_kps(const efo_kpt* k, int n)
{
    std::vector<cv::KeyPoint> v((size_t)n);
    for (int i = 0; i < n; i++) v[i] = cv::KeyPoint(k[i].x, k[i].y, k[i].size, k[i].angle);
    return v;
}

extern "C" {

int efref_bad_compute(const uint8_t* img, int w, int h, size_t pitch, const efo_kpt* kpts, int n,
                      float scale_factor, int nbits, uint8_t* desc)
{
    try {
        auto bad = cv::BAD::create(scale_factor, nbits == 512 ? cv::BAD::SIZE_512_BITS : cv::BAD::SIZE_256_BITS);
        cv::Mat image = wrap(img, w, h, pitch), out;
        auto kps = to_kps(kpts, n);
        bad->compute(image, kps, out);
        for (int i = 0; i < n; i++) std::memcpy(desc + (size_t)i * (nbits / 8), out.ptr<uchar>(i), (size_t)nbits / 8);
        return 0;
    } catch (const std::exception&) { return -1; }
}

int efref_hashsift_compute(const uint8_t* img, int w, int h, size_t pitch, const efo_kpt* kpts, int n,
                           float cropping_scale, int nbits, uint8_t* desc)
{
    try {
        auto hs = cv::HashSIFT::create(cropping_scale, nbits == 512 ? cv::HashSIFT::SIZE_512_BITS : cv::HashSIFT::SIZE_256_BITS);
        cv::Mat image = wrap(img, w, h, pitch), out;
        auto kps = to_kps(kpts, n);
        hs->compute(image, kps, out);
        for (int i = 0; i < n; i++) std::memcpy(desc + (size_t)i * (nbits / 8), out.ptr<uchar>(i), (size_t)nbits / 8);
        return 0;
    } catch (const std::exception&) { return -1; }
}

// hash_sift.cpp:333-351 (file-static there): the n x 129 response matrix before the projection
int efref_hashsift_features(const uint8_t* img, int w, int h, size_t pitch, const efo_kpt* kpts, int n,
                            float cropping_scale, float* resp129)
{
    try {
        cv::Mat image = wrap(img, w, h, pitch), responses;
        auto kps = to_kps(kpts, n);
        cv::computePatchSIFTs(image, kps, responses, cv::Size(32, 32), cropping_scale);
        for (int i = 0; i < n; i++) std::memcpy(resp129 + (size_t)i * 129, responses.ptr<float>(i), sizeof(float) * 129);
        return 0;
    } catch (const std::exception&) { return -1; }
}

} // extern "C"
