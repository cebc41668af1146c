// Minimal stand-in for <opencv2/core/cuda.hpp>, written for this repo (NOT OpenCV code; TEST INFRASTRUCTURE, see ../core.hpp):
// an owning, reference-counted cv::cuda::GpuMat over cudaMallocPitch with upload / download / copyTo / ROI, cv::cuda::Stream over a
// cudaStream_t, and the GpuMat halves of the InputArray / OutputArray proxies -- the members the reference's public headers, its
// test and its samples use, so that cuda-efficient-features_b200/cpp/opencv_adapter.cpp can be compiled and RUN where OpenCV is absent.
#pragma once
#include <cuda_runtime.h>

#include "../core.hpp"

namespace cv
{
namespace cuda
{
#define EF_SHIM_CUDA(expr) do { cudaError_t e__ = (expr); if (e__ != cudaSuccess) CV_Error(cv::Error::GpuApiCallError, cudaGetErrorString(e__)); } while (0)

class Stream
{
public:
    Stream() : impl_(std::make_shared<Impl>(true)) {}
    static Stream& Null() { static Stream s(0); return s; }
    void waitForCompletion() { EF_SHIM_CUDA(cudaStreamSynchronize(impl_->s)); }
    bool queryIfComplete() const { return cudaStreamQuery(impl_->s) == cudaSuccess; }
    cudaStream_t raw() const { return impl_->s; }
private:
    struct Impl {
        cudaStream_t s = nullptr; bool own;
        explicit Impl(bool create) : own(create) { if (create) EF_SHIM_CUDA(cudaStreamCreate(&s)); }
        ~Impl() { if (own && s) cudaStreamDestroy(s); }
    };
    explicit Stream(int) : impl_(std::make_shared<Impl>(false)) {}
    std::shared_ptr<Impl> impl_;
};

static inline int getDevice() { int d = 0; EF_SHIM_CUDA(cudaGetDevice(&d)); return d; }
static inline int getCudaEnabledDeviceCount() { int n = 0; return cudaGetDeviceCount(&n) == cudaSuccess ? n : 0; }

class GpuMat
{
public:
    int flags = 0, rows = 0, cols = 0;
    size_t step = 0;
    uchar* data = nullptr;

    GpuMat() {}
    GpuMat(int r, int c, int t) { create(r, c, t); }
    GpuMat(Size s, int t) { create(s.height, s.width, t); }
    GpuMat(int r, int c, int t, void* d, size_t s) : flags(t), rows(r), cols(c), step(s), data((uchar*)d) {}   // user memory, not owned
    explicit GpuMat(InputArray arr) { upload(arr); }

    int type() const { return flags; }
    int depth() const { return flags & 7; }
    int channels() const { return (flags >> 3) + 1; }
    size_t elemSize() const { static const int d[8] = { 1, 1, 2, 2, 4, 4, 8, 2 }; return (size_t)d[flags & 7] * ((flags >> 3) + 1); }
    Size size() const { return Size(cols, rows); }
    bool empty() const { return data == nullptr || rows == 0 || cols == 0; }

    void create(int r, int c, int t)
    {
        if (data && store_ && rows == r && cols == c && flags == t && own_rows_ == r && own_cols_ == c) return;
        release();
        if (r <= 0 || c <= 0) { flags = t; return; }
        flags = t; rows = r; cols = c; own_rows_ = r; own_cols_ = c;
        void* p = nullptr; size_t pitch = 0;
        if (r > 1 && c > 1) EF_SHIM_CUDA(cudaMallocPitch(&p, &pitch, (size_t)c * elemSize(), (size_t)r));
        else { pitch = (size_t)c * elemSize(); EF_SHIM_CUDA(cudaMalloc(&p, pitch * (size_t)r)); }   // OpenCV: a single row or column is continuous
        store_.reset((uchar*)p, [](uchar* q) { cudaFree(q); });
        data = (uchar*)p; step = pitch;
    }
    void create(Size s, int t) { create(s.height, s.width, t); }
    void release() { store_.reset(); data = nullptr; rows = cols = 0; step = 0; own_rows_ = own_cols_ = 0; }

    void upload(InputArray arr) { upload(arr, Stream::Null()); cudaStreamSynchronize(0); }
    void upload(InputArray arr, Stream& s)
    {
        const Mat m = arr.getMat();
        CV_Assert(!m.empty() && m.dims == 2);
        create(m.rows, m.cols, m.type());
        EF_SHIM_CUDA(cudaMemcpy2DAsync(data, step, m.data, (size_t)m.step, (size_t)cols * elemSize(), (size_t)rows, cudaMemcpyHostToDevice, s.raw()));
    }
    // blocking form (pageable host memory makes the async form synchronous too, as in OpenCV)
    void download(const _OutputArray& dst) const { download(dst, Stream::Null()); cudaStreamSynchronize(0); }
    void download(const _OutputArray& dst, Stream& s) const
    {
        CV_Assert(!empty());
        dst.create(rows, cols, flags);
        Mat& m = dst.getMatRef();
        EF_SHIM_CUDA(cudaMemcpy2DAsync(m.data, (size_t)m.step, data, step, (size_t)cols * elemSize(), (size_t)rows, cudaMemcpyDeviceToHost, s.raw()));
        EF_SHIM_CUDA(cudaStreamSynchronize(s.raw()));
    }
    void copyTo(GpuMat& dst, Stream& s) const
    {
        dst.create(rows, cols, flags);
        if (!empty()) EF_SHIM_CUDA(cudaMemcpy2DAsync(dst.data, dst.step, data, step, (size_t)cols * elemSize(), (size_t)rows, cudaMemcpyDeviceToDevice, s.raw()));
    }
    void copyTo(GpuMat& dst) const { copyTo(dst, Stream::Null()); }
    GpuMat colRange(int a, int b) const { GpuMat m = *this; m.data = data + (size_t)a * elemSize(); m.cols = b - a; return m; }
    GpuMat rowRange(int a, int b) const { GpuMat m = *this; m.data = data + (size_t)a * step; m.rows = b - a; return m; }
    template <class T> T* ptr(int y = 0) { return reinterpret_cast<T*>(data + (size_t)y * step); }
    template <class T> const T* ptr(int y = 0) const { return reinterpret_cast<const T*>(data + (size_t)y * step); }

private:
    std::shared_ptr<uchar> store_;
    int own_rows_ = 0, own_cols_ = 0;
};
} // namespace cuda

inline cuda::GpuMat _InputArray::getGpuMat() const { return g_ ? *g_ : cuda::GpuMat(); }
inline int _InputArray::type() const { return m_ ? m_->type() : g_ ? g_->type() : 0; }
inline bool _InputArray::empty() const { return m_ ? m_->empty() : g_ ? g_->empty() : true; }
inline Size _InputArray::size() const { return m_ ? m_->size() : g_ ? g_->size() : Size(); }
inline void _OutputArray::create(int r, int c, int t) const { CV_Assert(m_ || g_); if (m_) m_->create(r, c, t); else g_->create(r, c, t); }
inline void _OutputArray::release() const { if (m_) m_->release(); if (g_) g_->release(); }
} // namespace cv
