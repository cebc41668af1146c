// Minimal stand-in for <opencv2/core/cuda.hpp> (+ the few core types it pulls in): just enough for the reference's detector sources
// modules/cuda_efficient_features/src/cuda_fast.cu and cuda_efficient_features.cu to compile UNMODIFIED with nvcc (oracle/Makefile,
// target ref_cuda).  TEST INFRASTRUCTURE ONLY.  OpenCV (>= 4.6, third-party, not vendored) is not installed in this image; the
// types below keep OpenCV's member names and semantics for the members those two files touch, nothing else.
#ifndef EF_SHIM_OPENCV_CORE_CUDA_HPP
#define EF_SHIM_OPENCV_CORE_CUDA_HPP

#include <cfloat>
#include <cmath>
#include <cstddef>
#include <cstdio>
#include <cstdlib>
#include <memory>
#include <string>
#include <vector>

#include <cuda_runtime.h>

#define CV_PI 3.1415926535897932384626433832795
#define CV_2PI 6.283185307179586476925286766559
#define CV_CN_SHIFT 3
#define CV_8U 0
#define CV_32S 4
#define CV_32F 5
#define CV_MAKETYPE(depth, cn) ((depth) + (((cn) - 1) << CV_CN_SHIFT))
#define CV_8UC1 CV_MAKETYPE(CV_8U, 1)
#define CV_32FC1 CV_MAKETYPE(CV_32F, 1)
#define CV_32FC4 CV_MAKETYPE(CV_32F, 4)
#define CV_Error(code, msg) do { std::fprintf(stderr, "CV_Error: %s (%s:%d)\n", msg, __FILE__, __LINE__); std::abort(); } while (0)
#define CV_Assert(expr) do { if (!(expr)) { std::fprintf(stderr, "CV_Assert failed: %s (%s:%d)\n", #expr, __FILE__, __LINE__); std::abort(); } } while (0)

namespace cv
{
typedef unsigned char uchar;
template <class T> using Ptr = std::shared_ptr<T>;
typedef std::string String;

static inline int cvCeil(double v) { return (int)std::ceil(v); }
static inline int cvRound(double v) { return (int)std::lrint(v); }

struct Size { int width = 0, height = 0; Size() {} Size(int w, int h) : width(w), height(h) {} int area() const { return width * height; } };
struct Point2f { float x = 0, y = 0; };
struct Rect { int x = 0, y = 0, width = 0, height = 0; Rect() {} Rect(int x_, int y_, int w_, int h_) : x(x_), y(y_), width(w_), height(h_) {} };
namespace Error { enum Code { StsBadArg = -5 }; }
struct KeyPoint { Point2f pt; float size = 0, angle = -1, response = 0; int octave = 0, class_id = -1; };

class _InputArray { };
class _OutputArray : public _InputArray { };
typedef const _InputArray& InputArray;
typedef const _OutputArray& OutputArray;
static inline InputArray noArray() { static _OutputArray none; return none; }

// what HostMem::createMatHeader() returns: a header over host memory
struct Mat {
    uchar* data = nullptr; size_t step = 0; int rows = 0, cols = 0;
    template <class T> T* ptr(int y = 0) { return reinterpret_cast<T*>(data + (size_t)y * step); }
};

namespace cuda
{
class Stream { public: static Stream& Null() { static Stream s; return s; } };
struct StreamAccessor { static cudaStream_t getStream(const Stream&) { return 0; } };

static inline __host__ __device__ int divUp(int total, int grain) { return (total + grain - 1) / grain; }

template <class T> struct PtrStep {
    T* data; size_t step;
    __host__ __device__ PtrStep() : data(nullptr), step(0) {}
    __host__ __device__ PtrStep(T* d, size_t s) : data(d), step(s) {}
    __host__ __device__ operator T*() { return data; }                    // DevPtr<T>::operator T*() of OpenCV
    __host__ __device__ operator const T*() const { return data; }
    __host__ __device__ T* ptr(int y = 0) { return (T*)((char*)data + (size_t)y * step); }
    __host__ __device__ const T* ptr(int y = 0) const { return (const T*)((const char*)data + (size_t)y * step); }
    __host__ __device__ T& operator()(int y, int x) { return ptr(y)[x]; }
    __host__ __device__ const T& operator()(int y, int x) const { return ptr(y)[x]; }
};
template <class T> struct PtrStepSz : public PtrStep<T> {
    int cols, rows;
    __host__ __device__ PtrStepSz() : cols(0), rows(0) {}
    __host__ __device__ PtrStepSz(int r, int c, T* d, size_t s) : PtrStep<T>(d, s), cols(c), rows(r) {}
};
typedef PtrStep<uchar> PtrStepb;
typedef PtrStepSz<uchar> PtrStepSzb;
typedef PtrStep<float> PtrStepf;
typedef PtrStepSz<float> PtrStepSzf;
typedef PtrStep<int> PtrStepi;
typedef PtrStepSz<int> PtrStepSzi;

// device matrix header (no ownership: the shim's users allocate with cudaMalloc and wrap)
class GpuMat {
public:
    int flags = 0, rows = 0, cols = 0;
    size_t step = 0;
    uchar* data = nullptr;
    GpuMat() {}
    GpuMat(int r, int c, int type, void* d, size_t s) : flags(type), rows(r), cols(c), step(s), data((uchar*)d) {}
    int type() const { return flags; }
    Size size() const { return Size(cols, rows); }
    void release() { rows = cols = 0; }   // header only: the memory belongs to the caller
    // allocation / fill / ROI are only reached through calcIntegralImage, which the test wrappers never call
    void create(int, int, int) { std::fprintf(stderr, "GpuMat::create is not available in the shim\n"); std::abort(); }
    template <class S> void setTo(int, S&) { std::abort(); }
    GpuMat operator()(const Rect& r) const { GpuMat m = *this; m.data = data + (size_t)r.y * step + (size_t)r.x * 4; m.rows = r.height; m.cols = r.width; return m; }
    template <class T> T* ptr(int y = 0) { return reinterpret_cast<T*>(data + (size_t)y * step); }
    template <class T> const T* ptr(int y = 0) const { return reinterpret_cast<const T*>(data + (size_t)y * step); }
    template <class T> operator PtrStepSz<T>() const { return PtrStepSz<T>(rows, cols, (T*)data, step); }
    template <class T> operator PtrStep<T>() const { return PtrStep<T>((T*)data, step); }
};

// page-locked host buffer header
class HostMem {
public:
    int rows = 0, cols = 0; size_t step = 0; uchar* data = nullptr;
    HostMem() {}
    HostMem(int r, int c, void* d, size_t s) : rows(r), cols(c), step(s), data((uchar*)d) {}
    Size size() const { return Size(cols, rows); }
    Mat createMatHeader() const { Mat m; m.data = data; m.step = step; m.rows = rows; m.cols = cols; return m; }
};
} // namespace cuda
} // namespace cv

#endif
