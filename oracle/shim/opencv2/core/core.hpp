// ORACLE — TEST INFRASTRUCTURE ONLY.  Inert stand-ins for the OpenCV names that appear in the reference's headers and
// in its status-image drawing code (never exercised by the tests): enough to compile, nothing is drawn.
#pragma once
#include <string>
#include <vector>
typedef unsigned char uchar;
#define CV_8UC1 0
#define CV_8UC3 16
namespace cv {
struct Vec3b { uchar v[3] = {0, 0, 0}; uchar& operator[](int i) { return v[i]; } const uchar& operator[](int i) const { return v[i]; } };
struct Scalar { double val[4]; Scalar(double a = 0, double b = 0, double c = 0, double d = 0) { val[0] = a; val[1] = b; val[2] = c; val[3] = d; } double& operator[](int i) { return val[i]; } };
struct Point { int x, y; Point(int x_ = 0, int y_ = 0) : x(x_), y(y_) {} };
struct Size { int width, height; Size(int w = 0, int h = 0) : width(w), height(h) {} };
struct Mat {
  int rows = 0, cols = 0; std::vector<uchar> buf;
  Mat() {}
  Mat(int r, int c, int) : rows(r), cols(c), buf((size_t)r * c * 4 + 4) {}
  Mat(int r, int c, int t, const Scalar&) : Mat(r, c, t) {}
  int channels() const { return 3; }
  static Mat zeros(int r, int c, int t) { return Mat(r, c, t); }
  static Mat zeros(Size s, int t) { return Mat(s.height, s.width, t); }
  template <class T> T& at(int r, int c) { if (buf.size() < ((size_t)r * cols + c + 1) * sizeof(T)) buf.resize(((size_t)r * cols + c + 1) * sizeof(T)); return *reinterpret_cast<T*>(&buf[((size_t)r * cols + c) * sizeof(T)]); }
  bool empty() const { return buf.empty(); }
  Mat clone() const { return *this; }
};
enum { COLORMAP_HOT = 11, COLOR_HSV2BGR = 54, FONT_HERSHEY_SIMPLEX = 0 };
inline void applyColorMap(const Mat& a, Mat& b, int) { b = a; }
inline void cvtColor(const Mat& a, Mat& b, int) { b = a; }
inline void putText(Mat&, const std::string&, Point, int, double, Scalar, double = 1, int = 8) {}
inline void circle(Mat&, Point, int, Scalar, int = 1) {}
inline void line(Mat&, Point, Point, Scalar, int = 1) {}
inline void rectangle(Mat&, Point, Point, Scalar, int = 1) {}
inline bool imwrite(const std::string&, const Mat&) { return true; }
inline Mat imread(const std::string&, int = 1) { return Mat(); }
inline void imshow(const std::string&, const Mat&) {}
inline int waitKey(int = 0) { return 0; }
inline void vconcat(const Mat& a, const Mat&, Mat& c) { c = a; }
inline void hconcat(const Mat& a, const Mat&, Mat& c) { c = a; }
}  // namespace cv
