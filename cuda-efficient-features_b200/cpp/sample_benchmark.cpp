// sample_benchmark.cpp -- the reference's benchmark sample (samples/sample_benchmark.cpp:27-52,96-142) over the C++ host mirror
// (ef_features.hpp) of the C ABI: same options, same three benchmark types, same timing protocol (one discarded iteration, mean
// wall time of the next N with a stream synchronisation after every call), same two output lines.  No OpenCV in this build:
// the input is a binary PGM (P5) file or `synthetic:WxH` (uniform noise, lowbias32 counter hash of SURVEY 8d), and
// --dump=FILE writes "n descriptorSize" followed by the keypoints (x y octave response angle size) and the descriptor bytes in
// hex, one keypoint per line (tests/test_cpp_sample.py compares it with the Python mirror).
//
//   g++ -O2 -std=c++17 sample_benchmark.cpp -I../../include -L.. -lef_b200 -L/usr/local/cuda/lib64 -lcudart -o sample_benchmark
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <string>
#include <vector>

#include <cuda_runtime_api.h>

#include "ef_features.hpp"

namespace {

struct Options {
    std::string input;
    int max_keypoints = 10000, fast_threshold = 20, num_levels = 8, nonmax_radius = 15;
    int descriptor_type = 0, descriptor_bits = 256, benchmark_type = 0, num_iterations = 100;
    std::string dump;
};

bool parse(int argc, char** argv, Options& o)
{
    for (int i = 1; i < argc; i++) {
        const std::string a = argv[i];
        auto val = [&](const char* key, int& dst) {
            const std::string k = std::string("--") + key + "=";
            if (a.compare(0, k.size(), k) == 0) { dst = std::atoi(a.c_str() + k.size()); return true; }
            return false;
        };
        if (val("max-keypoints", o.max_keypoints) || val("fast-threshold", o.fast_threshold) || val("num-levels", o.num_levels) ||
            val("nonmax-radius", o.nonmax_radius) || val("descriptor-type", o.descriptor_type) || val("descriptor-bits", o.descriptor_bits) ||
            val("benchmark-type", o.benchmark_type) || val("num-iterations", o.num_iterations)) continue;
        if (a.compare(0, 7, "--dump=") == 0) { o.dump = a.substr(7); continue; }
        if (a == "--help" || a == "-h") return false;
        if (a.compare(0, 2, "--") == 0) { std::fprintf(stderr, "unknown option %s\n", a.c_str()); return false; }
        o.input = a;
    }
    return !o.input.empty();
}

uint32_t lowbias32(uint32_t x) { x ^= x >> 16; x *= 0x7feb352dU; x ^= x >> 15; x *= 0x846ca68bU; x ^= x >> 16; return x; }

bool load_image(const std::string& path, std::vector<uint8_t>& px, int& w, int& h)
{
    if (path.compare(0, 10, "synthetic:") == 0) {
        if (std::sscanf(path.c_str() + 10, "%dx%d", &w, &h) != 2 || w < 32 || h < 32) return false;
        px.resize((size_t)w * h);
        const uint32_t seed = 0xEFB20000u;
        for (int y = 0; y < h; y++)
            for (int x = 0; x < w; x++) px[(size_t)y * w + x] = (uint8_t)(lowbias32(seed ^ ((uint32_t)y * (uint32_t)w + (uint32_t)x)) >> 24);
        return true;
    }
    std::ifstream f(path, std::ios::binary);
    std::string magic;
    int maxv = 0;
    if (!(f >> magic >> w >> h >> maxv) || magic != "P5" || maxv != 255) return false;   // the samples convert to CV_8UC1 first (sample_common.cpp:35-45)
    f.get();
    px.resize((size_t)w * h);
    f.read(reinterpret_cast<char*>(px.data()), (std::streamsize)px.size());
    return (bool)f;
}

template <class F> double perf(int niterations, F fn)   // sample_benchmark.cpp:39-52
{
    uint64_t sum = 0;
    for (int iter = 0; iter <= niterations; iter++) {
        const auto t0 = std::chrono::steady_clock::now();
        fn();
        const auto t1 = std::chrono::steady_clock::now();
        if (iter > 0) sum += (uint64_t)std::chrono::duration_cast<std::chrono::microseconds>(t1 - t0).count();
    }
    return 1e-3 * (double)sum / niterations;
}

#define CUDA_OK(x) do { if ((x) != cudaSuccess) { std::fprintf(stderr, "CUDA error at %s:%d\n", __FILE__, __LINE__); return 2; } } while (0)

} // namespace

int main(int argc, char** argv)
{
    Options o;
    if (!parse(argc, argv, o)) {
        std::printf("usage: sample_benchmark <image.pgm | synthetic:WxH> [--max-keypoints=10000] [--fast-threshold=20] [--num-levels=8]\n"
                    "       [--nonmax-radius=15] [--descriptor-type=0 (0:BAD 1:HashSIFT)] [--descriptor-bits=256] [--benchmark-type=0\n"
                    "       (0:detect-and-compute 1:detect-only 2:compute-only)] [--num-iterations=100] [--dump=FILE]\n");
        return 1;
    }
    std::vector<uint8_t> gray;
    int w = 0, h = 0;
    if (!load_image(o.input, gray, w, h)) { std::printf("imread failed.\n"); return 1; }

    using efb200::EfficientFeatures;
    // getDescriptorType, sample_common.cpp:25-33
    const EfficientFeatures::DescriptorType dtype =
        o.descriptor_type == 0 ? (o.descriptor_bits == 256 ? EfficientFeatures::BAD_256 : EfficientFeatures::BAD_512)
      : o.descriptor_type == 1 ? (o.descriptor_bits == 256 ? EfficientFeatures::HASH_SIFT_256 : EfficientFeatures::HASH_SIFT_512)
                               : EfficientFeatures::HASH_SIFT_256;
    try {
        efb200::Capacity cap;
        cap.max_width = w; cap.max_height = h; cap.max_keypoints = o.max_keypoints;
        auto feature = EfficientFeatures::create(o.max_keypoints, 1.2f, 8, 0, 20, 15, EfficientFeatures::HASH_SIFT_256, cap);
        feature->setNLevels(o.num_levels);
        feature->setFastThreshold(o.fast_threshold);
        feature->setNonmaxRadius(o.nonmax_radius);
        feature->setDescriptorType(dtype);
        const int nf = feature->getMaxFeatures(), db = feature->descriptorSize();

        // cv::cuda::GpuMat d_gray(h_gray), d_keypoints, d_descriptors; cv::cuda::Stream stream;  (sample_benchmark.cpp:110-111)
        uint8_t *d_gray = nullptr, *d_desc = nullptr;
        float* d_kpts = nullptr;
        int* d_count = nullptr;
        cudaStream_t stream;
        CUDA_OK(cudaStreamCreate(&stream));
        CUDA_OK(cudaMalloc((void**)&d_gray, gray.size()));
        CUDA_OK(cudaMalloc((void**)&d_kpts, sizeof(float) * 5 * (size_t)nf));
        CUDA_OK(cudaMalloc((void**)&d_desc, (size_t)nf * db));
        CUDA_OK(cudaMalloc((void**)&d_count, sizeof(int)));
        CUDA_OK(cudaMemcpy(d_gray, gray.data(), gray.size(), cudaMemcpyHostToDevice));
        const efb200::MatView image{ d_gray, (size_t)w, h, w }, kpts{ d_kpts, sizeof(float) * (size_t)nf, 5, nf }, desc{ d_desc, (size_t)db, nf, db };

        int n = 0;
        auto count = [&]() { cudaMemcpyAsync(&n, d_count, sizeof(int), cudaMemcpyDeviceToHost, stream); cudaStreamSynchronize(stream); };
        double time = 0;
        if (o.benchmark_type == 0) {
            time = perf(o.num_iterations, [&]() { feature->detectAndComputeAsync(image, efb200::MatView(), kpts, desc, d_count, false, stream); count(); });
        } else if (o.benchmark_type == 1) {
            time = perf(o.num_iterations, [&]() { feature->detectAsync(image, kpts, d_count, stream); count(); });
        } else {
            feature->detectAsync(image, kpts, d_count, stream);
            count();
            const efb200::MatView found{ d_kpts, sizeof(float) * (size_t)nf, 5, n };
            time = perf(o.num_iterations, [&]() { feature->computeAsync(image, found, desc, stream); cudaStreamSynchronize(stream); });
        }
        std::printf("%5d keypoints found.\n", n);
        std::printf("processing time: %.1f[milli sec]\n", time);

        if (!o.dump.empty()) {
            std::vector<float> hk((size_t)5 * nf);
            std::vector<uint8_t> hd((size_t)nf * db);
            CUDA_OK(cudaMemcpy(hk.data(), d_kpts, hk.size() * sizeof(float), cudaMemcpyDeviceToHost));
            CUDA_OK(cudaMemcpy(hd.data(), d_desc, hd.size(), cudaMemcpyDeviceToHost));
            std::vector<efb200::KeyPoint> kp;
            EfficientFeatures::convert(hk.data(), sizeof(float) * (size_t)nf, n, kp);
            FILE* f = std::fopen(o.dump.c_str(), "w");
            if (!f) return 2;
            std::fprintf(f, "%d %d\n", n, db);
            for (int i = 0; i < n; i++) {
                uint32_t rb, ab, sb;
                std::memcpy(&rb, &kp[i].response, 4); std::memcpy(&ab, &kp[i].angle, 4); std::memcpy(&sb, &kp[i].size, 4);
                std::fprintf(f, "%d %d %d %08x %08x %08x ", (int)kp[i].x, (int)kp[i].y, kp[i].octave, rb, ab, sb);
                if (o.benchmark_type != 1) for (int b = 0; b < db; b++) std::fprintf(f, "%02x", hd[(size_t)i * db + b]);
                std::fprintf(f, "\n");
            }
            std::fclose(f);
        }
        cudaFree(d_gray); cudaFree(d_kpts); cudaFree(d_desc); cudaFree(d_count);
        cudaStreamDestroy(stream);
    } catch (const efb200::Error& e) {
        std::fprintf(stderr, "error %d: %s\n", e.status, e.what());
        return 2;
    }
    return 0;
}
