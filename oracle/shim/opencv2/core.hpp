// Minimal stand-in for the OpenCV core API, written for this repo (NOT OpenCV code).
// TEST INFRASTRUCTURE: it exists only so that the reference's CPU descriptor sources
// (/root/reference/modules/efficient_features/src/bad.cpp, hash_sift.cpp) compile UNMODIFIED
// into oracle/_ref/libef_ref.so, which pins oracle/ef_oracle.c.  Only the members those two files
// touch are provided.  Third-party arithmetic defined here: cv::integral (exact,
// wrapping int32), cv::gemm (fp32 out, double accumulation, ascending k), cvRound (half-even).
#pragma once
#include <cfloat>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <memory>
#include <stdexcept>
#include <string>
#include <vector>

#define CV_FINAL final
#define CV_OVERRIDE override
#define CV_WRAP
#define CV_OUT
#define CV_PI 3.1415926535897932384626433832795
#define CV_2PI 6.283185307179586476925286766559

#define CV_8U 0
#define CV_32S 4
#define CV_32F 5
#define CV_64F 6
#define CV_MAKETYPE(depth, cn) ((depth) + (((cn)-1) << 3))
#define CV_8UC1 CV_MAKETYPE(CV_8U, 1)
#define CV_8UC3 CV_MAKETYPE(CV_8U, 3)
#define CV_8UC4 CV_MAKETYPE(CV_8U, 4)
#define CV_32SC1 CV_MAKETYPE(CV_32S, 1)

#define CV_Error(code, msg) throw std::runtime_error(std::string(msg))
#define CV_Assert(expr) do { if (!(expr)) throw std::runtime_error("CV_Assert failed: " #expr); } while (0)
#define CV_DbgAssert(expr) ((void)0)

typedef unsigned char uchar;

static inline int cvRound(float v) { return (int)lrintf(v); }
static inline int cvRound(double v) { return (int)lrint(v); }
static inline int cvFloor(float v) { int i = (int)v; return i - (i > v); }
static inline int cvFloor(double v) { int i = (int)v; return i - (i > v); }

namespace cv
{
using ::uchar;
namespace Error { enum { StsBadArg = -5 }; }
enum { NORM_HAMMING = 6 };
enum { GEMM_1_T = 1, GEMM_2_T = 2, GEMM_3_T = 4 };

template <typename T> using Ptr = std::shared_ptr<T>;
template <typename T, typename... A> Ptr<T> makePtr(A&&... a) { return std::make_shared<T>(std::forward<A>(a)...); }

template <typename T> struct Point_ { T x, y; Point_() : x(0), y(0) {} Point_(T a, T b) : x(a), y(b) {} };
typedef Point_<float> Point2f;
template <typename T> struct Size_
{
    T width, height;
    Size_() : width(0), height(0) {}
    Size_(T w, T h) : width(w), height(h) {}
    T area() const { return width * height; }
};
typedef Size_<int> Size;
typedef Size_<float> Size2f;

struct KeyPoint
{
    Point2f pt; float size; float angle; float response; int octave; int class_id;
    KeyPoint() : size(0), angle(-1), response(0), octave(0), class_id(-1) {}
    KeyPoint(float x, float y, float s, float a = -1, float r = 0, int o = 0, int c = -1)
        : pt(x, y), size(s), angle(a), response(r), octave(o), class_id(c) {}
};

struct Matx23f
{
    float val[6];
    Matx23f() { for (float& v : val) v = 0; }
    float& operator()(int i, int j) { return val[i * 3 + j]; }
    const float& operator()(int i, int j) const { return val[i * 3 + j]; }
};

template <typename T> static inline T saturate_cast(float v);
template <> inline uchar saturate_cast<uchar>(float v) { int i = cvRound(v); return (uchar)(i < 0 ? 0 : (i > 255 ? 255 : i)); }

struct MatStep
{
    size_t p[3];
    MatStep() { p[0] = p[1] = p[2] = 0; }
    operator size_t() const { return p[0]; }
    size_t operator[](int i) const { return p[i]; }
};

class Mat
{
public:
    int dims, rows, cols;
    uchar* data;
    MatStep step;

    Mat() : dims(0), rows(0), cols(0), data(nullptr), type_(0) { size_[0] = size_[1] = size_[2] = 0; }
    Mat(int r, int c, int t) : Mat() { create(r, c, t); }
    Mat(int r, int c, int t, void* user, size_t stepBytes = 0) : Mat()
    {
        dims = 2; rows = r; cols = c; type_ = t; size_[0] = r; size_[1] = c;
        data = (uchar*)user; step.p[0] = stepBytes ? stepBytes : (size_t)c * elemSize(); step.p[1] = elemSize();
    }
    Mat(int nd, const int* sz, int t) : Mat() { create(nd, sz, t); }

    void create(int r, int c, int t)
    {
        if (dims == 2 && rows == r && cols == c && type_ == t && data) return;
        const int sz[2] = { r, c }; create(2, sz, t);
    }
    void create(Size s, int t) { create(s.height, s.width, t); }
    void create(int nd, const int* sz, int t)
    {
        dims = nd; type_ = t; size_t total = elemSize();
        for (int i = 0; i < 3; i++) size_[i] = i < nd ? sz[i] : 1;
        for (int i = nd - 1; i >= 0; i--) { step.p[i] = total; total *= (size_t)sz[i]; }
        rows = nd == 2 ? sz[0] : -1; cols = nd == 2 ? sz[1] : -1;
        store_.reset(new uchar[total ? total : 1], std::default_delete<uchar[]>());
        data = store_.get(); total_ = total;
    }
    void release() { store_.reset(); data = nullptr; dims = rows = cols = 0; }
    bool empty() const { return data == nullptr || total_elems() == 0; }
    int type() const { return type_; }
    int depth() const { return type_ & 7; }
    int channels() const { return (type_ >> 3) + 1; }
    size_t elemSize() const { static const int d[8] = { 1, 1, 2, 2, 4, 4, 8, 2 }; return (size_t)d[type_ & 7] * ((type_ >> 3) + 1); }
    Size size() const { return Size(cols, rows); }
    size_t total_elems() const { if (dims == 2) return (size_t)rows * cols; size_t t = 1; for (int i = 0; i < dims; i++) t *= (size_t)size_[i]; return dims ? t : 0; }

    template <typename T> T* ptr(int i0 = 0) { return (T*)(data + (size_t)i0 * step.p[0]); }
    template <typename T> const T* ptr(int i0 = 0) const { return (const T*)(data + (size_t)i0 * step.p[0]); }
    template <typename T> T& at(int i, int j) { return ((T*)(data + (size_t)i * step.p[0]))[j]; }
    template <typename T> const T& at(int i, int j) const { return ((const T*)(data + (size_t)i * step.p[0]))[j]; }

    Mat& operator=(double v)
    {   // `hist = 0` in hash_sift.cpp:229 (only zero is needed, any depth)
        CV_Assert(v == 0 && store_); std::memset(data, 0, total_); return *this;
    }
    void convertTo(Mat& dst, int t) const
    {   // only CV_64F -> CV_32F is used (hash_sift.cpp:390-392)
        CV_Assert(type_ == CV_64F && t == CV_32F && dims == 2);
        dst.create(rows, cols, CV_32F);
        for (int i = 0; i < rows; i++) { const double* s = ptr<double>(i); float* d = dst.ptr<float>(i); for (int j = 0; j < cols; j++) d[j] = (float)s[j]; }
    }

private:
    int type_; int size_[3]; size_t total_ = 0;
    std::shared_ptr<uchar> store_;
};

class _InputArray
{
public:
    _InputArray() : m_(nullptr) {}
    _InputArray(const Mat& m) : m_(const_cast<Mat*>(&m)) {}
    Mat getMat() const { return m_ ? *m_ : Mat(); }
protected:
    Mat* m_;
};
class _OutputArray : public _InputArray
{
public:
    _OutputArray() {}
    _OutputArray(Mat& m) : _InputArray(m) {}
    void create(int r, int c, int t) const { m_->create(r, c, t); }
    void release() const { if (m_) m_->release(); }
};
typedef const _InputArray& InputArray;
typedef const _OutputArray& OutputArray;

// cv::integral for CV_8UC1 -> CV_32SC1, (h+1)x(w+1), exact with int32 wrap-around (bad.cpp:286)
static inline void integral(const Mat& src, Mat& sum)
{
    CV_Assert(src.type() == CV_8UC1);
    sum.create(src.rows + 1, src.cols + 1, CV_32SC1);
    std::memset(sum.ptr<int>(0), 0, sizeof(int) * (size_t)(src.cols + 1));
    for (int y = 0; y < src.rows; y++)
    {
        const uchar* s = src.ptr<uchar>(y);
        const uint32_t* up = (const uint32_t*)sum.ptr<int>(y);
        uint32_t* cur = (uint32_t*)sum.ptr<int>(y + 1);
        uint32_t rs = 0; cur[0] = 0;
        for (int x = 0; x < src.cols; x++) { rs += s[x]; cur[x + 1] = up[x + 1] + rs; }
    }
}

// cv::gemm(A, B, 1, Mat(), 0, C, GEMM_2_T) for CV_32F: C = A * B^T, double accumulation in ascending k,
// one rounding to fp32 (what OpenCV's built-in kernel produces; LAPACK builds differ).
static inline void gemm(const Mat& A, const Mat& B, double alpha, const Mat&, double beta, Mat& C, int flags)
{
    CV_Assert(A.type() == CV_32F && B.type() == CV_32F && flags == GEMM_2_T && alpha == 1 && beta == 0 && A.cols == B.cols);
    C.create(A.rows, B.rows, CV_32F);
    for (int i = 0; i < A.rows; i++)
    {
        const float* a = A.ptr<float>(i); float* c = C.ptr<float>(i);
        for (int j = 0; j < B.rows; j++)
        {
            const float* b = B.ptr<float>(j); double acc = 0;
            for (int k = 0; k < A.cols; k++) acc += (double)a[k] * (double)b[k];
            c[j] = (float)acc;
        }
    }
}
} // namespace cv
