// Stand-in for the OpenCV C++ headers (TEST INFRASTRUCTURE, see oracle/refshim/README.md).
// cv::Mat here is a plain row-major buffer behind a shared header; element-wise float expressions are evaluated in
// place as the single IEEE operations OpenCV performs for CV_32F; every OpenCV ALGORITHM the pinned reference functions
// call is forwarded to the real OpenCV (the cv2 module) through the callback table below, registered by
// oracle/refbridge.py.  Calls the pinned functions never reach (drawing, DFT, ...) abort.
#ifndef REFSHIM_OPENCV_CORE
#define REFSHIM_OPENCV_CORE
#include <algorithm>          // the real OpenCV headers pull these in; the reference files rely on that
#include <cassert>
#include <cmath>
#include <cstdint>
#include <functional>
#include <iostream>
#include <sstream>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <memory>
#include <string>
#include <vector>

typedef int64_t int64;            // cvdef.h
typedef uint64_t uint64;
typedef unsigned char uchar;
typedef unsigned short ushort;

#define CV_8U 0
#define CV_32F 5
#define CV_64F 6
#define CV_MAKETYPE(depth, cn) ((depth) + (((cn)-1) << 3))
#define CV_8UC1 CV_MAKETYPE(CV_8U, 1)
#define CV_8UC3 CV_MAKETYPE(CV_8U, 3)
#define CV_32FC1 CV_MAKETYPE(CV_32F, 1)
#define CV_32FC2 CV_MAKETYPE(CV_32F, 2)
#define CV_32FC3 CV_MAKETYPE(CV_32F, 3)

extern "C" {
// the real OpenCV, reached through the Python process that loaded this library (oracle/refbridge.py)
struct RefCvCallbacks {
    // cv::filter2D(src, dst, -1, kernel, Point(-1,-1), 0, BORDER_REFLECT101) on CV_32F
    void (*filter2D)(const float *src, int rows, int cols, const float *kernel, int krows, int kcols, float *dst);
    double (*mean)(const float *src, int rows, int cols);                                   // cv::mean(src)[0]
    void (*minmax)(const float *src, int rows, int cols, double *mn, double *mx);           // minMaxIdx / minMaxLoc
    // cv::remap(src, dst, mapx, mapy, INTER_LINEAR, BORDER_CONSTANT, Scalar(0))
    void (*remap)(const float *src, int rows, int cols, const float *mapx, const float *mapy, int drows, int dcols, float *dst);
    void (*to_u8)(const float *src, int rows, int cols, double alpha, uint8_t *dst);        // src.convertTo(dst, CV_8UC1, alpha)
    // cv::imread(path, IMREAD_GRAYSCALE): returns 0 on success; buf == NULL asks for the size only
    int (*imread_gray)(const char *path, int *rows, int *cols, uint8_t *buf, long cap);
    // cv::ORB::create() + setPatchSize + setEdgeThreshold, then compute(img, keypoints, descriptors): kp_xy holds n (x, y)
    // pairs of size kp_size on entry and the m kept ones on return; desc has room for n x 32 bytes; returns m
    int (*orb_compute)(const uint8_t *img, int rows, int cols, int patch_size, int edge_threshold, float *kp_xy, int n, float kp_size,
                       uint8_t *desc);
    // cv::DescriptorMatcher::create(BRUTEFORCE_HAMMING)->knnMatch(query, train, matches, 2): per query the number of
    // matches returned (0..2) and their (trainIdx, distance)
    void (*knn_match)(const uint8_t *q, int nq, const uint8_t *t, int nt, int32_t *counts, int32_t *train_idx_nqx2, float *dist_nqx2);
};
extern RefCvCallbacks g_ref_cv;
}

namespace cv {

[[noreturn]] inline void refshim_unreached(const char *what)
{
    std::fprintf(stderr, "refshim: %s is not forwarded (the pinned reference functions never call it)\n", what);
    std::abort();
}

template <class T> struct Point_ {
    T x, y;
    Point_() : x(0), y(0) {}
    Point_(T x_, T y_) : x(x_), y(y_) {}
    template <class U> Point_(const Point_<U> &o) : x((T)o.x), y((T)o.y) {}
};
typedef Point_<int> Point;
typedef Point_<int> Point2i;
typedef Point_<float> Point2f;
typedef Point_<double> Point2d;

template <class T> struct Size_ {
    T width, height;
    Size_() : width(0), height(0) {}
    Size_(T w, T h) : width(w), height(h) {}
};
typedef Size_<int> Size;

struct Scalar {
    double val[4];
    Scalar(double a = 0, double b = 0, double c = 0, double d = 0) { val[0] = a; val[1] = b; val[2] = c; val[3] = d; }
    double operator[](int i) const { return val[i]; }
};

template <class T, int N> struct Vec {
    T val[N];
    Vec() { for (int i = 0; i < N; ++i) val[i] = T(); }
    Vec(T a, T b, T c) { static_assert(N == 3, "3 channels"); val[0] = a; val[1] = b; val[2] = c; }
    T &operator[](int i) { return val[i]; }
};
typedef Vec<uchar, 3> Vec3b;
typedef Vec<float, 3> Vec3f;

struct KeyPoint {
    Point2f pt;
    float size = 0, angle = -1, response = 0;
    int octave = 0, class_id = -1;
    KeyPoint() {}
    KeyPoint(Point2f p, float s, float a = -1, float r = 0, int o = 0, int c = -1) : pt(p), size(s), angle(a), response(r), octave(o), class_id(c) {}
    KeyPoint(float x, float y, float s, float a = -1, float r = 0, int o = 0, int c = -1) : pt(x, y), size(s), angle(a), response(r), octave(o), class_id(c) {}
};
struct DMatch { int queryIdx = -1, trainIdx = -1, imgIdx = -1; float distance = 0; };

enum { BORDER_CONSTANT = 0, BORDER_REFLECT101 = 4, BORDER_REFLECT_101 = 4, BORDER_DEFAULT = 4 };
enum { INTER_NEAREST = 0, INTER_LINEAR = 1 };
enum { IMREAD_GRAYSCALE = 0 };
enum { NORM_MINMAX = 32 };
enum { COLOR_GRAY2BGR = 8 };

class Mat {
public:
    int rows = 0, cols = 0;
    uchar *data = nullptr;

    Mat() {}
    Mat(int r, int c, int type) { create(r, c, type); }
    Mat(Size s, int type) { create(s.height, s.width, type); }
    void create(int r, int c, int type)
    {
        rows = r; cols = c; type_ = type;
        buf_ = std::make_shared<std::vector<uchar>>((size_t)r * c * elemSize(), (uchar)0);
        data = buf_->data();
    }
    static Mat zeros(int r, int c, int type) { return Mat(r, c, type); }
    static Mat zeros(Size s, int type) { return Mat(s, type); }
    static Mat ones(int r, int c, int type)
    {
        Mat m(r, c, type);
        if (m.depth() != CV_32F || m.channels() != 1) refshim_unreached("Mat::ones of a non-float type");
        for (size_t i = 0; i < (size_t)r * c; ++i) ((float *)m.data)[i] = 1.0f;
        return m;
    }
    int type() const { return type_; }
    int depth() const { return type_ & 7; }
    int channels() const { return (type_ >> 3) + 1; }
    size_t elemSize() const { return (size_t)channels() * (depth() == CV_8U ? 1 : depth() == CV_32F ? 4 : 8); }
    Size size() const { return Size(cols, rows); }
    bool empty() const { return rows == 0 || cols == 0; }
    size_t total() const { return (size_t)rows * cols; }
    Mat clone() const
    {
        Mat m;
        m.rows = rows; m.cols = cols; m.type_ = type_;
        m.buf_ = std::make_shared<std::vector<uchar>>(buf_ ? *buf_ : std::vector<uchar>());
        m.data = m.buf_->data();
        return m;
    }
    void copyTo(Mat &dst) const { dst = clone(); }
    template <class T> T &at(int i, int j)
    {
        assert(sizeof(T) == elemSize() && i >= 0 && i < rows && j >= 0 && j < cols);
        return ((T *)data)[(size_t)i * cols + j];
    }
    template <class T> const T &at(int i, int j) const
    {
        assert(sizeof(T) == elemSize() && i >= 0 && i < rows && j >= 0 && j < cols);
        return ((const T *)data)[(size_t)i * cols + j];
    }
    template <class T> T *ptr(int i = 0) { return (T *)(data + (size_t)i * cols * elemSize()); }
    template <class T> const T *ptr(int i = 0) const { return (const T *)(data + (size_t)i * cols * elemSize()); }
    Mat row(int) const { refshim_unreached("Mat::row"); }
    float *f32() { need_f32("float expression"); return (float *)data; }
    const float *f32() const { need_f32("float expression"); return (const float *)data; }
    void need_f32(const char *what) const { if (type_ != CV_32F) refshim_unreached(what); }

    // a.mul(b): cv::multiply on CV_32F, one float product per element
    Mat mul(const Mat &b) const
    {
        assert(rows == b.rows && cols == b.cols);
        Mat r(rows, cols, CV_32F);
        const float *x = f32(), *y = b.f32();
        float *z = r.f32();
        for (size_t i = 0; i < total(); ++i) z[i] = x[i] * y[i];
        return r;
    }
    // src.convertTo(dst, rtype, alpha): only the CV_32F -> CV_8U scaling polar_to_cartesian_radar_img uses (real OpenCV)
    void convertTo(Mat &dst, int rtype, double alpha = 1, double beta = 0) const
    {
        if (!(type_ == CV_32F && rtype == CV_8UC1 && beta == 0)) refshim_unreached("Mat::convertTo other than CV_32F -> CV_8UC1");
        Mat out(rows, cols, CV_8UC1);
        g_ref_cv.to_u8(f32(), rows, cols, alpha, out.data);
        dst = out;
    }

private:
    int type_ = 0;
    std::shared_ptr<std::vector<uchar>> buf_;
};

template <class T> class Mat_ : public Mat {
public:
    Mat_(const Mat &m) : Mat(m) {}
};

// ---- element-wise CV_32F expressions: what OpenCV evaluates is one float operation per element, the scalar narrowed
// to float first (arithm_op converts the scalar to the array's depth; MatExpr scaling is convertTo(-1, alpha), whose
// CV_32F kernel multiplies by (float)alpha) ----
inline Mat refshim_map(const Mat &a, float (*f)(float, float), float s)
{
    Mat r(a.rows, a.cols, CV_32F);
    const float *x = a.f32();
    float *z = r.f32();
    for (size_t i = 0; i < a.total(); ++i) z[i] = f(x[i], s);
    return r;
}
inline Mat operator-(const Mat &a, double s) { return refshim_map(a, [](float x, float t) { return x - t; }, (float)s); }
inline Mat operator+(const Mat &a, double s) { return refshim_map(a, [](float x, float t) { return x + t; }, (float)s); }
inline Mat operator-(double s, const Mat &a) { return refshim_map(a, [](float x, float t) { return t - x; }, (float)s); }
inline Mat operator*(const Mat &a, double s) { return refshim_map(a, [](float x, float t) { return x * t; }, (float)s); }
inline Mat operator/(const Mat &a, double s) { return refshim_map(a, [](float x, float t) { return x * t; }, (float)(1. / s)); }
inline Mat &operator/=(Mat &a, double s)              // mat.inl.hpp: a.convertTo(a, -1, 1./s)
{
    float *x = a.f32();
    const float t = (float)(1. / s);
    for (size_t i = 0; i < a.total(); ++i) x[i] = x[i] * t;
    return a;
}
inline Mat abs(const Mat &a) { return refshim_map(a, [](float x, float) { return std::fabs(x); }, 0.f); }

// ---- algorithms: the real OpenCV ----
inline void filter2D(const Mat &src, Mat &dst, int ddepth, const Mat &kernel, Point anchor = Point(-1, -1), double delta = 0,
                     int borderType = BORDER_DEFAULT)
{
    if (ddepth != -1 || anchor.x != -1 || anchor.y != -1 || delta != 0 || borderType != BORDER_REFLECT101)
        refshim_unreached("filter2D with other than (-1, Point(-1,-1), 0, BORDER_REFLECT101)");
    Mat out(src.rows, src.cols, CV_32F);
    g_ref_cv.filter2D(src.f32(), src.rows, src.cols, kernel.f32(), kernel.rows, kernel.cols, out.f32());
    dst = out;
}
inline Scalar mean(const Mat &src) { return Scalar(g_ref_cv.mean(src.f32(), src.rows, src.cols)); }
inline void minMaxIdx(const Mat &src, double *mn, double *mx = nullptr, int * = nullptr, int * = nullptr)
{
    double a = 0, b = 0;
    g_ref_cv.minmax(src.f32(), src.rows, src.cols, &a, &b);
    if (mn) *mn = a;
    if (mx) *mx = b;
}
inline void minMaxLoc(const Mat &src, double *mn, double *mx = nullptr, Point * = nullptr, Point * = nullptr) { minMaxIdx(src, mn, mx); }
inline void remap(const Mat &src, Mat &dst, const Mat &map1, const Mat &map2, int interpolation, int borderMode = BORDER_CONSTANT,
                  const Scalar &borderValue = Scalar())
{
    if (interpolation != INTER_LINEAR || borderMode != BORDER_CONSTANT || borderValue[0] != 0) refshim_unreached("remap with other flags");
    Mat out(map1.rows, map1.cols, CV_32F);
    g_ref_cv.remap(src.f32(), src.rows, src.cols, map1.f32(), map2.f32(), map1.rows, map1.cols, out.f32());
    dst = out;
}
inline Mat imread(const std::string &path, int flags = 1)
{
    if (flags != IMREAD_GRAYSCALE) refshim_unreached("imread other than IMREAD_GRAYSCALE");
    int r = 0, c = 0;
    if (g_ref_cv.imread_gray(path.c_str(), &r, &c, nullptr, 0) != 0) return Mat();
    Mat m(r, c, CV_8UC1);
    if (g_ref_cv.imread_gray(path.c_str(), &r, &c, m.data, (long)r * c) != 0) return Mat();
    return m;
}
// vconcat(a, b, dst): plain row stacking (dst may alias a or b: built aside first)
inline void vconcat(const Mat &a, const Mat &b, Mat &dst)
{
    assert(a.cols == b.cols && a.type() == b.type());
    Mat out(a.rows + b.rows, a.cols, a.type());
    const size_t na = (size_t)a.rows * a.cols * a.elemSize(), nb = (size_t)b.rows * b.cols * b.elemSize();
    std::memcpy(out.data, a.data, na);
    std::memcpy(out.data + na, b.data, nb);
    dst = out;
}

// ---- features2d: the real OpenCV ----
template <class T> class Ptr {
public:
    Ptr() {}
    Ptr(std::shared_ptr<T> p) : p_(p) {}
    T *operator->() const { return p_.get(); }
    T &operator*() const { return *p_; }
    explicit operator bool() const { return (bool)p_; }
private:
    std::shared_ptr<T> p_;
};

class ORB {
public:
    static Ptr<ORB> create() { return Ptr<ORB>(std::make_shared<ORB>()); }
    void setPatchSize(int p) { patch_ = p; }
    void setEdgeThreshold(int e) { edge_ = e; }
    // Feature2D::compute: keypoints the extractor cannot describe are REMOVED from the vector (the reference relies on it)
    void compute(const Mat &image, std::vector<KeyPoint> &keypoints, Mat &descriptors)
    {
        if (image.type() != CV_8UC1) refshim_unreached("ORB::compute on a non-8-bit image");
        const int n = (int)keypoints.size();
        std::vector<float> xy(2 * (size_t)std::max(n, 1));
        float size = 31;
        for (int i = 0; i < n; ++i) {
            xy[2 * i] = keypoints[i].pt.x; xy[2 * i + 1] = keypoints[i].pt.y;
            if (i == 0) size = keypoints[i].size;
            else if (keypoints[i].size != size) refshim_unreached("ORB::compute with keypoints of different sizes");
        }
        Mat d(std::max(n, 1), 32, CV_8UC1);
        const int m = g_ref_cv.orb_compute(image.data, image.rows, image.cols, patch_, edge_, xy.data(), n, size, d.data);
        keypoints.resize((size_t)m);
        for (int i = 0; i < m; ++i) keypoints[i] = KeyPoint(xy[2 * i], xy[2 * i + 1], size);
        Mat out(m, 32, CV_8UC1);
        if (m) std::memcpy(out.data, d.data, 32 * (size_t)m);
        descriptors = out;
    }
private:
    int patch_ = 31, edge_ = 31;
};

class DescriptorMatcher {
public:
    enum { FLANNBASED = 1, BRUTEFORCE = 2, BRUTEFORCE_L1 = 3, BRUTEFORCE_HAMMING = 4 };
    static Ptr<DescriptorMatcher> create(int type)
    {
        if (type != BRUTEFORCE_HAMMING) refshim_unreached("DescriptorMatcher other than BRUTEFORCE_HAMMING");
        return Ptr<DescriptorMatcher>(std::make_shared<DescriptorMatcher>());
    }
    void knnMatch(const Mat &query, const Mat &train, std::vector<std::vector<DMatch>> &matches, int k) const
    {
        if (k != 2) refshim_unreached("knnMatch with k != 2");
        const int nq = query.rows, nt = train.rows;
        std::vector<int32_t> counts((size_t)std::max(nq, 1)), idx(2 * (size_t)std::max(nq, 1));
        std::vector<float> dist(2 * (size_t)std::max(nq, 1));
        g_ref_cv.knn_match(query.data, nq, train.data, nt, counts.data(), idx.data(), dist.data());
        matches.assign((size_t)nq, std::vector<DMatch>());
        for (int i = 0; i < nq; ++i)
            for (int c = 0; c < counts[i]; ++c) {
                DMatch m;
                m.queryIdx = i; m.trainIdx = idx[2 * i + c]; m.distance = dist[2 * i + c];
                matches[i].push_back(m);
            }
    }
};

inline void imshow(const std::string &, const Mat &) { refshim_unreached("imshow"); }
inline int waitKey(int = 0) { refshim_unreached("waitKey"); }

// ---- never reached by the pinned functions ----
inline void hconcat(const Mat &, const Mat &, Mat &) { refshim_unreached("hconcat"); }
inline void transpose(const Mat &, Mat &) { refshim_unreached("transpose"); }
inline void dft(const Mat &, Mat &, int = 0, int = 0) { refshim_unreached("dft"); }
inline void merge(const Mat *, size_t, Mat &) { refshim_unreached("merge"); }
inline void split(const Mat &, Mat *) { refshim_unreached("split"); }
inline void magnitude(const Mat &, const Mat &, Mat &) { refshim_unreached("magnitude"); }
inline void normalize(const Mat &, const Mat &, double = 1, double = 0, int = 0) { refshim_unreached("normalize"); }
inline void cvtColor(const Mat &, Mat &, int) { refshim_unreached("cvtColor"); }
template <class P> inline void circle(Mat &, P, int, const Scalar &, int = 1) { refshim_unreached("circle"); }
template <class P> inline void line(Mat &, P, P, const Scalar &, int = 1) { refshim_unreached("line"); }

}  // namespace cv
#endif
