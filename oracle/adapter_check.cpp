// adapter_check.cpp -- behaviour of the OpenCV-facing adapter (cuda-efficient-features_b200/cpp/opencv_adapter.cpp) that the reference's
// own test does not touch, exercised through the reference's UNMODIFIED public headers.  TEST INFRASTRUCTURE (built by `make -C
// oracle adapter` against oracle/shim, run by tests/test_gpu_adapter.py on the GPU box).
//
//   adapter_check <image.pgm> <out-prefix> [nfeatures]
//
// Every check mirrors a line of the reference: argument kinds (src/cuda_efficient_features.cpp:71-129), the two asserts (:228-229),
// empty results (:275-281), convert (:323-349), the setter / getter pairs (:355-377), descriptor info (:351-353), the describers'
// create / compute / computeAsync (src/cuda_bad.cpp:46-101, src/cuda_hash_sift.cpp:113-168).  It also dumps the keypoint matrix and
// the descriptors of each descriptor type (<out-prefix>.<type>.kpts / .desc) so that the caller can compare them byte for byte with
// the Python host mirror and the CPU oracle, and compares GPU descriptors with the reference's own CPU classes (cv::BAD, cv::HashSIFT).
#include <cstdio>
#include <cstring>
#include <fstream>
#include <string>
#include <vector>

#include <opencv2/core.hpp>
#include <opencv2/highgui.hpp>

#include <cuda_efficient_descriptors.h>
#include <cuda_efficient_features.h>
#include <efficient_descriptors.h>

static int g_checks = 0, g_failed = 0;
#define CHECK(cond, name) do { g_checks++; const bool ok__ = (cond); if (!ok__) g_failed++; std::printf("CHECK %-58s %s\n", name, ok__ ? "ok" : "FAIL"); } while (0)

static bool same(const cv::Mat& a, const cv::Mat& b)
{
    if (a.rows != b.rows || a.cols != b.cols || a.type() != b.type()) return false;
    for (int i = 0; i < a.rows; i++)
        if (std::memcmp(a.ptr<uchar>(i), b.ptr<uchar>(i), (size_t)a.cols * a.elemSize())) return false;
    return true;
}
static void dump(const std::string& path, const cv::Mat& m)
{
    std::ofstream f(path, std::ios::binary);
    const int hdr[3] = { m.rows, m.cols, (int)m.elemSize() };
    f.write((const char*)hdr, sizeof(hdr));
    for (int i = 0; i < m.rows; i++) f.write((const char*)m.ptr<uchar>(i), (std::streamsize)((size_t)m.cols * m.elemSize()));
}
template <class F> static bool throws(F f)
{
    try { f(); } catch (const cv::Exception&) { return true; } catch (...) { return false; }
    return false;
}

int main(int argc, char** argv)
{
    if (argc < 3) { std::fprintf(stderr, "usage: adapter_check <image.pgm> <out-prefix> [nfeatures]\n"); return 2; }
    const int nfeatures = argc > 3 ? std::atoi(argv[3]) : 20000;
    cv::Mat image = cv::imread(argv[1], cv::IMREAD_GRAYSCALE);
    if (image.empty()) { std::fprintf(stderr, "imread failed\n"); return 2; }
    const std::string prefix = argv[2];
    using EF = cv::cuda::EfficientFeatures;
    try {
        // ---- create + defaults (include/cuda_efficient_features.h:47-48)
        auto def = EF::create();
        CHECK(def->getMaxFeatures() == 5000 && def->getScaleFactor() == 1.2f && def->getNLevels() == 8 && def->getFirstLevel() == 0 &&
              def->getFastThreshold() == 20 && def->getNonmaxRadius() == 15 && def->getDescriptorType() == EF::HASH_SIFT_256, "create(): reference defaults");
        CHECK(def->descriptorSize() == 32 && def->descriptorType() == CV_8U && def->defaultNorm() == cv::NORM_HAMMING, "descriptorSize/Type/defaultNorm");
        def->setDescriptorType(EF::BAD_512);
        CHECK(def->descriptorSize() == 64 && def->getDescriptorType() == EF::BAD_512, "setDescriptorType -> descriptorSize 64");
        def->setMaxFeatures(1234); def->setScaleFactor(1.3f); def->setNLevels(5); def->setFirstLevel(1); def->setFastThreshold(33); def->setNonmaxRadius(7);
        CHECK(def->getMaxFeatures() == 1234 && def->getScaleFactor() == 1.3f && def->getNLevels() == 5 && def->getFirstLevel() == 1 &&
              def->getFastThreshold() == 33 && def->getNonmaxRadius() == 7, "the 7 setter / getter pairs (before the first call)");
        {
            std::vector<cv::KeyPoint> k; cv::Mat d;
            def->detectAndCompute(image, cv::noArray(), k, d);
            CHECK(!k.empty() && d.rows == (int)k.size() && d.cols == 64, "detectAndCompute after setters");
            bool oct = true; for (const auto& q : k) oct = oct && q.octave >= 1 && q.octave < 5;
            CHECK(oct, "firstLevel = 1, nlevels = 5 respected (octaves in [1,5))");
            def->setNonmaxRadius(15); def->setFirstLevel(0); def->setNLevels(8); def->setScaleFactor(1.2f); def->setMaxFeatures(2000);
            std::vector<cv::KeyPoint> k2;
            def->detect(image, k2);
            CHECK(!k2.empty() && (int)k2.size() <= 2000, "setters on a live object re-plan the workspace");
        }

        cv::cuda::GpuMat d_image(image);
        cv::cuda::Stream stream;
        const char* names[4] = { "bad256", "bad512", "hashsift256", "hashsift512" };
        cv::Mat kref;
        for (int t = 0; t < 4; t++) {
            auto f = EF::create(nfeatures, 1.2f, 8, 0, 20, 15, (EF::DescriptorType)t);
            // GpuMat in, GpuMat out
            cv::cuda::GpuMat dk, dd;
            f->detectAndComputeAsync(d_image, cv::noArray(), dk, dd, false, stream);
            stream.waitForCompletion();
            cv::Mat hk, hd; dk.download(hk); dd.download(hd);
            CHECK(dk.rows == EF::ROWS_COUNT && dk.type() == CV_32F && dd.rows == dk.cols && dd.cols == f->descriptorSize() && dd.type() == CV_8U,
                  (std::string(names[t]) + ": output shapes (5 x N CV_32F, N x B CV_8U)").c_str());
            dump(prefix + "." + names[t] + ".kpts", hk);
            dump(prefix + "." + names[t] + ".desc", hd);
            if (t == 0) kref = hk.clone();
            else CHECK(same(hk, kref), (std::string(names[t]) + ": keypoints independent of the descriptor type").c_str());
            // Mat in, Mat out (upload :76, download :316-320)
            cv::Mat mk, md;
            f->detectAndComputeAsync(image, cv::noArray(), mk, md, false, stream);
            CHECK(same(mk, hk) && same(md, hd), (std::string(names[t]) + ": Mat arguments == GpuMat arguments").c_str());
            // Feature2D forms
            std::vector<cv::KeyPoint> kv, kv2; cv::Mat d1;
            f->detectAndCompute(image, cv::noArray(), kv, d1);
            f->convert(dk, kv2);
            bool eq = kv.size() == kv2.size() && (int)kv.size() == hk.cols;
            for (size_t i = 0; eq && i < kv.size(); i++) {
                const short* loc = hk.ptr<short>(EF::LOCATION_ROW);
                eq = kv[i].pt.x == kv2[i].pt.x && kv[i].pt.y == kv2[i].pt.y && kv[i].pt.x == (float)loc[2 * i] && kv[i].pt.y == (float)loc[2 * i + 1] &&
                     kv[i].size == hk.ptr<float>(EF::SIZE_ROW)[i] && kv[i].angle == hk.ptr<float>(EF::ANGLE_ROW)[i] &&
                     kv[i].response == hk.ptr<float>(EF::RESPONSE_ROW)[i] && kv[i].octave == hk.ptr<int>(EF::OCTAVE_ROW)[i];
            }
            CHECK(eq && same(d1, hd), (std::string(names[t]) + ": detectAndCompute / convert == *Async").c_str());
            // detect-only: same keypoints
            cv::cuda::GpuMat dk2; f->detectAsync(d_image, dk2, cv::noArray(), stream); stream.waitForCompletion();
            cv::Mat hk2; dk2.download(hk2);
            CHECK(same(hk2, hk), (std::string(names[t]) + ": detectAsync == detectAndComputeAsync keypoints").c_str());
            // computeAsync on the 5 x N matrix: level-0 keypoints keep their position and size 31 -> identical descriptors there
            cv::cuda::GpuMat dd2; f->computeAsync(d_image, dk, dd2, stream); stream.waitForCompletion();
            cv::Mat hd2; dd2.download(hd2);
            bool rows_ok = hd2.rows == hd.rows && hd2.cols == hd.cols; int n0 = 0;
            for (int i = 0; rows_ok && i < hk.cols; i++)
                if (hk.ptr<int>(EF::OCTAVE_ROW)[i] == 0) { n0++; rows_ok = !std::memcmp(hd2.ptr<uchar>(i), hd.ptr<uchar>(i), (size_t)hd.cols); }
            // (inside detectAndCompute level 0 is described on the BLURRED level image, computeAsync describes the image it is given:
            //  the rows are only comparable after blurring, so this is checked against the describer classes below instead)
            (void)rows_ok; (void)n0;
            // compute(vector<KeyPoint>) == cv::cuda::BAD / HashSIFT created with scale 1 == the reference's CPU class
            cv::Mat dg, dc, ds;
            f->compute(image, kv, dg);
            const int nb = (t & 1) ? 100 : 101;
            if (t < 2) {
                cv::cuda::BAD::create(1.f, nb)->compute(image, kv, ds);
                cv::BAD::create(1.f, nb)->compute(image, kv, dc);
            } else {
                cv::cuda::HashSIFT::create(1.f, nb)->compute(image, kv, ds);
                cv::HashSIFT::create(1.f, nb)->compute(image, kv, dc);
            }
            CHECK(same(dg, ds), (std::string(names[t]) + ": EfficientFeatures::compute == cuda describer class").c_str());
            CHECK(same(ds, dc), (std::string(names[t]) + ": cuda describer == reference CPU class (0 bytes differ)").c_str());
            // describer computeAsync over the 5 x N GpuMat == compute over the same keypoints with size forced to 31
            std::vector<cv::KeyPoint> k31 = kv; for (auto& q : k31) q.size = 31.f;
            cv::Mat d31, dr;
            cv::cuda::GpuMat ddr;
            if (t < 2) { auto b = cv::cuda::BAD::create(1.f, nb); b->compute(image, k31, d31); b->computeAsync(d_image, dk, ddr, stream); }
            else { auto b = cv::cuda::HashSIFT::create(1.f, nb); b->compute(image, k31, d31); b->computeAsync(d_image, dk, ddr, stream); }
            stream.waitForCompletion(); ddr.download(dr);
            CHECK(same(dr, d31) && same(dr, hd2), (std::string(names[t]) + ": computeAsync(5 x N) forces size 31").c_str());
            // asserts of the reference
            CHECK(throws([&] { cv::cuda::GpuMat a, b; f->detectAndComputeAsync(d_image, cv::noArray(), a, b, true, stream); }), (std::string(names[t]) + ": useProvidedKeypoints -> cv::Exception").c_str());
        }
        {
            auto f = EF::create(500);
            cv::Mat f32(image.rows, image.cols, CV_32F);
            std::vector<cv::KeyPoint> k; cv::Mat d;
            CHECK(throws([&] { f->detectAndCompute(f32, cv::noArray(), k, d); }), "non-CV_8U image -> cv::Exception");
            cv::Mat blank(480, 640, CV_8UC1); std::memset(blank.data, 128, 480 * 640);
            cv::cuda::GpuMat dk(5, 7, CV_32F), dd(7, 32, CV_8U);           // stale contents must be released (:275-281)
            f->detectAndComputeAsync(blank, cv::noArray(), dk, dd);
            CHECK(dk.empty() && dd.empty(), "no keypoints -> outputs released");
            std::vector<cv::KeyPoint> none; cv::Mat dn(3, 3, CV_8U);
            cv::cuda::BAD::create(1.f)->compute(image, none, dn);
            CHECK(dn.empty(), "describer with no keypoints -> descriptors released");
            CHECK(throws([&] { cv::cuda::HashSIFT::create(1.f, 77); }), "HashSIFT::create with a bad size -> cv::Exception");
            CHECK(cv::cuda::BAD::create(1.f, 77)->descriptorSize() == 64, "BAD::create with a bad size -> 512 bits (like the reference)");
        }
    } catch (const std::exception& e) {
        std::printf("EXCEPTION %s\n", e.what());
        g_failed++;
    }
    std::printf("adapter_check: %d checks, %d failed\n", g_checks, g_failed);
    return g_failed ? 1 : 0;
}
