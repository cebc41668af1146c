// Minimal stand-in for the OpenCV core API, written for this repo (NOT OpenCV code).
// TEST INFRASTRUCTURE: it exists only so that the reference's CPU descriptor sources
// (/root/reference/modules/efficient_features/src/bad.cpp, hash_sift.cpp) compile UNMODIFIED
// into oracle/_ref/libef_ref.so, which pins oracle/ef_oracle.c.  Only the members those two files
// touch are provided.  Third-party arithmetic defined here: cv::integral (exact,
// wrapping int32), cv::gemm (fp32 out, double accumulation, ascending k), cvRound (half-even).
//
// Round 2: the same stand-in also carries what the reference's PUBLIC headers, its own test
// (tests/descriptor_test.cpp) and its benchmark sample (samples/sample_benchmark.cpp, sample_common.cpp) need, so that they
// compile UNMODIFIED against cuda-efficient-features_b200/cpp/opencv_adapter.cpp + libef_b200.so (oracle/Makefile, target
// `adapter`): cv::Exception, Vec / Scalar / DMatch, Mat ROI / clone / copyTo, _InputArray kinds (MAT, CUDA_GPU_MAT),
// absdiff / countNonZero, format, CommandLineParser; cv::cuda::GpuMat / Stream live in core/cuda.hpp (pulled in only when
// the CUDA runtime headers are on the include path).  None of this is product code.
#pragma once
#include <algorithm>
#include <cfloat>
#include <cmath>
#include <cstdint>
#include <chrono>
#include <type_traits>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <iostream>
#include <map>
#include <memory>
#include <sstream>
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
#define CV_32FC1 CV_MAKETYPE(CV_32F, 1)
#define CV_32FC4 CV_MAKETYPE(CV_32F, 4)

namespace cv
{
// cv::Exception: what CV_Error / CV_Assert throw
class Exception : public std::runtime_error
{
public:
    Exception(int c, const std::string& m) : std::runtime_error(m), code(c), msg(m) {}
    int code; std::string msg;
};
}
#define CV_Error(code_, msg_) throw cv::Exception((int)(code_), std::string(msg_))
#define CV_Assert(expr) do { if (!(expr)) throw cv::Exception(-215, "CV_Assert failed: " #expr); } while (0)
#define CV_DbgAssert(expr) ((void)0)

typedef unsigned char uchar;

static inline int cvRound(float v) { return (int)lrintf(v); }
static inline int cvRound(double v) { return (int)lrint(v); }
static inline int cvFloor(float v) { int i = (int)v; return i - (i > v); }
static inline int cvFloor(double v) { int i = (int)v; return i - (i > v); }
static inline int cvCeil(float v) { int i = (int)v; return i + (i < v); }
static inline int cvCeil(double v) { int i = (int)v; return i + (i < v); }

namespace cv
{
using ::uchar;
namespace Error { enum Code { StsBadArg = -5, StsAssert = -215, GpuApiCallError = -217 }; }
enum { NORM_HAMMING = 6 };
typedef std::string String;
enum { GEMM_1_T = 1, GEMM_2_T = 2, GEMM_3_T = 4 };

template <typename T> using Ptr = std::shared_ptr<T>;
template <typename T, typename... A> Ptr<T> makePtr(A&&... a) { return std::make_shared<T>(std::forward<A>(a)...); }

template <typename T> struct Point_
{
    T x, y; Point_() : x(0), y(0) {} Point_(T a, T b) : x(a), y(b) {}
    template <typename U> Point_(const Point_<U>& o) : x(saturate_point<T>(o.x)), y(saturate_point<T>(o.y)) {}   // Point2f -> Point rounds, like OpenCV
    Point_& operator*=(T s) { x *= s; y *= s; return *this; }
private:
    template <typename V, typename U> static V saturate_point(U v) { return std::is_integral<V>::value && !std::is_integral<U>::value ? (V)lrint((double)v) : (V)v; }
};
typedef Point_<int> Point;
template <typename T, int N> struct Vec
{
    T val[N];
    Vec() { for (T& v : val) v = T(); }
    Vec(T a, T b) { static_assert(N == 2, ""); val[0] = a; val[1] = b; }
    Vec(T a, T b, T c, T d) { static_assert(N == 4, ""); val[0] = a; val[1] = b; val[2] = c; val[3] = d; }
    T& operator[](int i) { return val[i]; }
    const T& operator[](int i) const { return val[i]; }
};
typedef Vec<short, 2> Vec2s;
typedef Vec<float, 4> Vec4f;
struct Scalar
{
    double val[4];
    Scalar() : val{ 0, 0, 0, 0 } {}
    Scalar(double a, double b = 0, double c = 0, double d = 0) : val{ a, b, c, d } {}
    static Scalar all(double v) { return Scalar(v, v, v, v); }
};
struct DMatch { int queryIdx = -1, trainIdx = -1, imgIdx = -1; float distance = FLT_MAX; };
struct Range { int start, end; Range(int s, int e) : start(s), end(e) {} };
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
static inline std::ostream& operator<<(std::ostream& os, const Size& s) { return os << "[" << s.width << " x " << s.height << "]"; }

struct KeyPoint
{
    Point2f pt; float size; float angle; float response; int octave; int class_id;
    KeyPoint() : size(0), angle(-1), response(0), octave(0), class_id(-1) {}
    KeyPoint(float x, float y, float s, float a = -1, float r = 0, int o = 0, int c = -1)
        : pt(x, y), size(s), angle(a), response(r), octave(o), class_id(c) {}
    KeyPoint(Point2f p, float s, float a = -1, float r = 0, int o = 0, int c = -1)
        : pt(p), size(s), angle(a), response(r), octave(o), class_id(c) {}
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

    // single-index element access of an n x 1 matrix (tmp.at<Vec4f>(i)); ROI headers share the storage
    template <typename T> T& at(int i) { return *(T*)(data + (size_t)i * step.p[0]); }
    template <typename T> const T& at(int i) const { return *(const T*)(data + (size_t)i * step.p[0]); }
    bool isContinuous() const { return dims != 2 || step.p[0] == (size_t)cols * elemSize(); }
    size_t step1(int i = 0) const { static const int d[8] = { 1, 1, 2, 2, 4, 4, 8, 2 }; return step.p[i] / (size_t)d[type_ & 7]; }
    Mat colRange(int a, int b) const { Mat m = *this; m.data = data + (size_t)a * elemSize(); m.cols = b - a; m.size_[1] = b - a; return m; }
    Mat rowRange(int a, int b) const { Mat m = *this; m.data = data + (size_t)a * step.p[0]; m.rows = b - a; m.size_[0] = b - a; return m; }
    Mat rowRange(const Range& r) const { return rowRange(r.start, r.end); }
    void copyTo(Mat& dst) const
    {
        CV_Assert(dims == 2);
        dst.create(rows, cols, type_);
        for (int i = 0; i < rows; i++) std::memcpy(dst.ptr<uchar>(i), ptr<uchar>(i), (size_t)cols * elemSize());
    }
    Mat clone() const { Mat m; copyTo(m); return m; }

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

namespace cuda { class GpuMat; }

// InputArray / OutputArray proxies: a Mat, a cuda::GpuMat or nothing (noArray()); kind() like OpenCV's _InputArray::KindFlag
class _InputArray
{
public:
    enum KindFlag { NONE = 0, MAT = 1 << 16, CUDA_GPU_MAT = 9 << 16 };
    _InputArray() : kind_(NONE), m_(nullptr), g_(nullptr) {}
    _InputArray(const Mat& m) : kind_(MAT), m_(const_cast<Mat*>(&m)), g_(nullptr) {}
    _InputArray(const cuda::GpuMat& g) : kind_(CUDA_GPU_MAT), m_(nullptr), g_(const_cast<cuda::GpuMat*>(&g)) {}
    KindFlag kind() const { return kind_; }
    Mat getMat() const { return m_ ? *m_ : Mat(); }
    // the members below also serve the GpuMat kind; they are defined in core/cuda.hpp when the CUDA headers are present
    inline cuda::GpuMat getGpuMat() const;
    inline int type() const;
    inline bool empty() const;
    inline Size size() const;
protected:
    KindFlag kind_; Mat* m_; cuda::GpuMat* g_;
};
class _OutputArray : public _InputArray
{
public:
    _OutputArray() {}
    _OutputArray(Mat& m) : _InputArray(m) {}
    _OutputArray(cuda::GpuMat& g) : _InputArray(g) {}
    bool needed() const { return kind_ != NONE; }
    inline void create(int r, int c, int t) const;
    inline void create(Size s, int t) const { create(s.height, s.width, t); }
    inline void release() const;
    Mat& getMatRef() const { return *m_; }
    cuda::GpuMat& getGpuMatRef() const { return *g_; }
};
typedef const _InputArray& InputArray;
typedef const _OutputArray& OutputArray;
static inline const _OutputArray& noArray() { static const _OutputArray none; return none; }

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

#if __has_include(<cuda_runtime.h>)
#include "core/cuda.hpp"
#else
// host-only build (oracle/_ref/libef_ref.so): the proxies only ever hold a Mat
namespace cv
{
inline int _InputArray::type() const { return m_ ? m_->type() : 0; }
inline bool _InputArray::empty() const { return !m_ || m_->empty(); }
inline Size _InputArray::size() const { return m_ ? m_->size() : Size(); }
inline void _OutputArray::create(int r, int c, int t) const { CV_Assert(m_); m_->create(r, c, t); }
inline void _OutputArray::release() const { if (m_) m_->release(); }
}
#endif

namespace cv
{
// cv::absdiff / cv::countNonZero for single-channel 8-bit matrices (tests/descriptor_test.cpp:40-42)
static inline void absdiff(const Mat& a, const Mat& b, Mat& dst)
{
    CV_Assert(a.type() == CV_8UC1 && b.type() == CV_8UC1 && a.rows == b.rows && a.cols == b.cols);
    dst.create(a.rows, a.cols, CV_8UC1);
    for (int i = 0; i < a.rows; i++) {
        const uchar* pa = a.ptr<uchar>(i); const uchar* pb = b.ptr<uchar>(i); uchar* pd = dst.ptr<uchar>(i);
        for (int j = 0; j < a.cols; j++) pd[j] = (uchar)(pa[j] > pb[j] ? pa[j] - pb[j] : pb[j] - pa[j]);
    }
}
static inline int countNonZero(const Mat& a)
{
    CV_Assert(a.type() == CV_8UC1);
    int n = 0;
    for (int i = 0; i < a.rows; i++) { const uchar* p = a.ptr<uchar>(i); for (int j = 0; j < a.cols; j++) n += p[j] != 0; }
    return n;
}
// cv::hconcat of equally tall 2-D matrices of one type (samples/hpatches_description.cpp:222-223)
static inline void hconcat(const std::vector<Mat>& src, Mat& dst)
{
    CV_Assert(!src.empty());
    int cols = 0;
    for (const Mat& m : src) { CV_Assert(m.dims == 2 && m.rows == src[0].rows && m.type() == src[0].type()); cols += m.cols; }
    Mat out(src[0].rows, cols, src[0].type());
    const size_t es = out.elemSize();
    for (int y = 0; y < out.rows; y++) {
        uchar* d = out.ptr<uchar>(y);
        for (const Mat& m : src) { std::memcpy(d, m.ptr<uchar>(y), (size_t)m.cols * es); d += (size_t)m.cols * es; }
    }
    dst = out;
}
// cv::fastAtan2, scalar form: the 7th-order odd polynomial of OpenCV's mathfuncs_core (degrees, [0, 360)); restated from the published
// algorithm and pinned against cv2.fastAtan2 by tests/test_hpatches_tool.py (through oracle/_ref/fastatan2_check)
static inline float fastAtan2(float y, float x)
{
    const float scale = (float)(180 / CV_PI);
    const float p1 = 0.9997878412794807f * scale, p3 = -0.3258083974640975f * scale, p5 = 0.1555786518463281f * scale, p7 = -0.04432655554792128f * scale;
    const float ax = std::abs(x), ay = std::abs(y);
    float a, c, c2;
    if (ax >= ay) { c = ay / (ax + (float)DBL_EPSILON); c2 = c * c; a = (((p7 * c2 + p5) * c2 + p3) * c2 + p1) * c; }
    else { c = ax / (ay + (float)DBL_EPSILON); c2 = c * c; a = 90.f - (((p7 * c2 + p5) * c2 + p3) * c2 + p1) * c; }
    if (x < 0) a = 180.f - a;
    if (y < 0) a = 360.f - a;
    return a;
}
static inline String format(const char* fmt, ...)
{
    char buf[1024];
    va_list ap; va_start(ap, fmt); std::vsnprintf(buf, sizeof(buf), fmt, ap); va_end(ap);
    return String(buf);
}

// cv::TickMeter (samples/sample_image_sequence.cpp:96-103)
class TickMeter
{
public:
    void start() { t0_ = std::chrono::steady_clock::now(); }
    void stop() { total_ += std::chrono::duration<double>(std::chrono::steady_clock::now() - t0_).count(); n_++; }
    void reset() { total_ = 0; n_ = 0; }
    double getTimeSec() const { return total_; }
    double getTimeMilli() const { return total_ * 1e3; }
    long getCounter() const { return n_; }
private:
    std::chrono::steady_clock::time_point t0_; double total_ = 0; long n_ = 0;
};

// cv::CommandLineParser for key strings of the form "{ name alias | default | help }" (samples/sample_benchmark.cpp:27-37)
class CommandLineParser
{
public:
    CommandLineParser(int argc, const char* const argv[], const String& keys)
    {
        size_t pos = 0;
        while ((pos = keys.find('{', pos)) != String::npos) {
            const size_t end = keys.find('}', pos);
            std::vector<String> f = split(keys.substr(pos + 1, end - pos - 1), '|');
            f.resize(3);
            Key k; k.def = trim(f[1]); k.help = trim(f[2]); k.value = k.def == "<none>" ? "" : k.def; k.set = false;
            std::istringstream names(f[0]); String nm;
            while (names >> nm) k.names.push_back(nm);
            if (!k.names.empty()) { if (k.names[0][0] == '@') positional_.push_back((int)keys_.size()); keys_.push_back(k); }
            pos = end;
        }
        size_t npos = 0;
        for (int i = 1; i < argc; i++) {
            String a = argv[i];
            if (a.rfind("--", 0) == 0 || (a.size() > 1 && a[0] == '-' && !isdigit((unsigned char)a[1]))) {
                a = a.substr(a.find_first_not_of('-'));
                String v = "true"; const size_t eq = a.find('=');
                if (eq != String::npos) { v = a.substr(eq + 1); a = a.substr(0, eq); }
                Key* k = find(a);
                if (k) { k->value = v; k->set = true; } else errors_.push_back("unknown option: " + a);
            } else if (npos < positional_.size()) { Key& k = keys_[positional_[npos++]]; k.value = a; k.set = true; }
        }
    }
    bool has(const String& name) const { const Key* k = const_cast<CommandLineParser*>(this)->find(name); return k && (k->set || (!k->def.empty() && k->def != "<none>")); }
    template <typename T> T get(const String& name) const
    {
        const Key* k = const_cast<CommandLineParser*>(this)->find(name);
        T out = T();
        if (!k) { errors_.push_back("undeclared key: " + name); return out; }
        if (k->value.empty()) { errors_.push_back("missing parameter: " + name); return out; }
        std::istringstream is(k->value);
        if (!(is >> out)) errors_.push_back("cannot parse parameter: " + name);
        return out;
    }
    bool check() const { return errors_.empty(); }
    void printErrors() const { for (const String& e : errors_) std::cout << "ERROR: " << e << std::endl; }
    void printMessage() const
    {
        std::cout << "Usage: [params]";
        for (int i : positional_) std::cout << " " << keys_[i].names[0].substr(1);
        std::cout << std::endl;
        for (const Key& k : keys_) {
            std::cout << "\t";
            for (size_t i = 0; i < k.names.size(); i++) std::cout << (i ? ", " : "") << (k.names[i][0] == '@' ? "" : "--") << k.names[i];
            if (!k.def.empty()) std::cout << " (value:" << k.def << ")";
            std::cout << "\n\t\t" << k.help << std::endl;
        }
    }
private:
    struct Key { std::vector<String> names; String def, help, value; bool set; };
    static String trim(const String& s) { const size_t a = s.find_first_not_of(" \t"), b = s.find_last_not_of(" \t"); return a == String::npos ? "" : s.substr(a, b - a + 1); }
    static std::vector<String> split(const String& s, char c) { std::vector<String> v; std::istringstream is(s); String t; while (std::getline(is, t, c)) v.push_back(t); return v; }
    Key* find(const String& name) { for (Key& k : keys_) for (const String& n : k.names) if (n == name) return &k; return nullptr; }
    std::vector<Key> keys_; std::vector<int> positional_; mutable std::vector<String> errors_;
};
template <> inline String CommandLineParser::get<String>(const String& name) const
{
    const Key* k = const_cast<CommandLineParser*>(this)->find(name);
    if (!k) { errors_.push_back("undeclared key: " + name); return ""; }
    if (k->value.empty()) errors_.push_back("missing parameter: " + name);
    return k->value;
}
} // namespace cv
