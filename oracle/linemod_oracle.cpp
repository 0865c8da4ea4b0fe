// oracle/linemod_oracle.cpp
//
// TEST INFRASTRUCTURE ONLY.  CPU restatement of the `cv::linemod` algorithm that
// aelmiger/LINE-MOD-Pipeline drives through `detector->match(...)`
// (reference: src/HighLevelLinemod.cpp:152) and `detector->addTemplate(...)`
// (reference: src/HighLevelLinemod.cpp:93).  The arithmetic itself lives in the
// un-vendored third-party module opencv_contrib `modules/rgbd/src/linemod.cpp`
// (reference pins no version: README.md:68 "OPENCV4", CMakeLists.txt:53); this file
// restates that module's published algorithm function by function ([UP] tags below
// name the upstream function each block follows; SURVEY.md Appendix A is the spec).
//
// Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
// legs may load this library.  The product (line_mod_pipeline_b200/csrc) never does.
//
// PARITY PIN STATUS: the reference ships no tests for this path ("parity unpinned" by
// the reference).  The pins used instead are (a) every primitive cross-checked against
// real OpenCV 4.13 (cv2) in tests/test_oracle_primitives.py and (b) SURVEY.md §8c golden
// hashes G1..G6 computed with real OpenCV primitives on benchmark/img0.png+depth0.png
// (tests/test_oracle_golden.py).  NORMAL_LUT contents are a documented stand-in
// (upstream normal_lut.i is not available offline); DepthNormal label parity is
// "vs oracle with the same table".
//
// Build: see oracle/Makefile (g++ -O3 -ffp-contract=off -pthread -shared -fPIC).
// -ffp-contract=off matters: upstream is built for baseline x86-64 (no FMA contraction).

#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <limits>
#include <map>
#include <string>
#include <vector>
#include <atomic>
#include <thread>

namespace lmo {

typedef uint8_t u8;
typedef uint16_t u16;

// ----------------------------------------------------------------------------------
// Data model — [UP] linemod.hpp: Feature, Template, Match
// ----------------------------------------------------------------------------------
struct Feature { int x, y, label; };
struct Template {
  int width = 0, height = 0, pyramid_level = 0;
  std::vector<Feature> features;
};
typedef std::vector<Template> TemplatePyramid;  // index = level*num_modalities + modality

struct Match {
  int x, y;
  float similarity;
  int class_index;  // index into the detector's sorted class list (stands for class_id)
  int template_id;
  // [UP] Match::operator< : similarity descending, then template_id ascending
  bool operator<(const Match& r) const {
    if (similarity != r.similarity) return similarity > r.similarity;
    return template_id < r.template_id;
  }
  // [UP] Match::operator== : x, y, similarity, class_id (NOT template_id)
  bool operator==(const Match& r) const {
    return x == r.x && y == r.y && similarity == r.similarity && class_index == r.class_index;
  }
};

static inline int clampi(int v, int lo, int hi) { return v < lo ? lo : (v > hi ? hi : v); }
static inline int reflect101(int p, int n) {
  if (n == 1) return 0;
  while (p < 0 || p >= n) {
    if (p < 0) p = -p;
    else p = 2 * (n - 1) - p;
  }
  return p;
}

// ----------------------------------------------------------------------------------
// OpenCV imgproc primitives restated (integer-exact for 8-bit input)
// ----------------------------------------------------------------------------------

// Border-extended copy of an interleaved 8-bit image: pad pixels on every side, replicate (mode 0) or
// reflect-101 (mode 1).  Lets the filters below run as plain contiguous loops the compiler vectorises.
static void pad_image(const u8* src, int rows, int cols, int ch, int pad, int mode, std::vector<u8>& out) {
  int pr = rows + 2 * pad, pc = cols + 2 * pad;
  out.resize((size_t)pr * pc * ch);
  for (int y = 0; y < pr; ++y) {
    int sy = mode ? reflect101(y - pad, rows) : clampi(y - pad, 0, rows - 1);
    const u8* s = src + (size_t)sy * cols * ch;
    u8* d = out.data() + (size_t)y * pc * ch;
    std::memcpy(d + (size_t)pad * ch, s, (size_t)cols * ch);
    for (int x = 0; x < pad; ++x) {
      int sl = mode ? reflect101(x - pad, cols) : 0, sr = mode ? reflect101(cols + x, cols) : cols - 1;
      for (int c = 0; c < ch; ++c) {
        d[(size_t)x * ch + c] = s[(size_t)sl * ch + c];
        d[(size_t)(pad + cols + x) * ch + c] = s[(size_t)sr * ch + c];
      }
    }
  }
}

// cv::GaussianBlur(src, dst, Size(7,7), 0, 0, BORDER_REPLICATE) on CV_8UC{ch}.
// Fixed-point kernel 256*getGaussianKernel(7,0) = [8,28,56,72,56,28,8]; one rounding:
// (sum + 2^15) >> 16.   
static void gauss7(const u8* src, int rows, int cols, int ch, u8* dst) {
  std::vector<u8> pad;
  pad_image(src, rows, cols, ch, 3, 0, pad);
  const int pc = cols + 6, n = cols * ch;
  std::vector<u16> h((size_t)(rows + 6) * n);  // horizontal sums <= 255*256 fit 16 bits
  for (int y = 0; y < rows + 6; ++y) {
    const u8* p = pad.data() + (size_t)y * pc * ch;
    u16* o = h.data() + (size_t)y * n;
    for (int i = 0; i < n; ++i)
      o[i] = (u16)(8 * (p[i] + p[i + 6 * ch]) + 28 * (p[i + ch] + p[i + 5 * ch]) + 56 * (p[i + 2 * ch] + p[i + 4 * ch]) + 72 * p[i + 3 * ch]);
  }
  for (int y = 0; y < rows; ++y) {
    const u16 *r0 = h.data() + (size_t)y * n, *r1 = r0 + n, *r2 = r1 + n, *r3 = r2 + n, *r4 = r3 + n, *r5 = r4 + n, *r6 = r5 + n;
    u8* o = dst + (size_t)y * n;
    for (int i = 0; i < n; ++i) {
      unsigned s = 8u * (r0[i] + r6[i]) + 28u * (r1[i] + r5[i]) + 56u * (r2[i] + r4[i]) + 72u * r3[i];
      o[i] = (u8)((s + 32768u) >> 16);
    }
  }
}

// cv::Sobel(src, d, CV_16S, {1,0}|{0,1}, 3, 1, 0, BORDER_REPLICATE) on CV_8UC{ch}.
static void sobel3(const u8* src, int rows, int cols, int ch, int16_t* dx, int16_t* dy) {
  std::vector<u8> pad;
  pad_image(src, rows, cols, ch, 1, 0, pad);
  const int pc = cols + 2, n = cols * ch;
  for (int y = 0; y < rows; ++y) {
    const u8 *a = pad.data() + (size_t)y * pc * ch, *b = a + (size_t)pc * ch, *c = b + (size_t)pc * ch;
    int16_t *ox = dx + (size_t)y * n, *oy = dy + (size_t)y * n;
    for (int i = 0; i < n; ++i) {
      int l = a[i] + 2 * b[i] + c[i], r = a[i + 2 * ch] + 2 * b[i + 2 * ch] + c[i + 2 * ch];
      int t = a[i] + 2 * a[i + ch] + a[i + 2 * ch], u = c[i] + 2 * c[i + ch] + c[i + 2 * ch];
      ox[i] = (int16_t)(r - l);
      oy[i] = (int16_t)(u - t);
    }
  }
}

// cv::hal::fastAtan32f vector body, degrees.
// fused=1 evaluates the polynomial with fmaf (what the cv2 4.13 wheel does bit-exactly);
// fused=0 with separate mul/add.  The quantised label is identical either way (golden G1).
static inline float fast_atan2_deg(float y, float x, int fused) {
  const float scale = (float)(180.0 / 3.14159265358979323846);
  const float p1 = 0.9997878412794807f * scale, p3 = -0.3258083974640975f * scale;
  const float p5 = 0.1555786518463281f * scale, p7 = -0.04432655554792128f * scale;
  float ax = std::fabs(x), ay = std::fabs(y);
  float mx = ax > ay ? ax : ay, mn = ax < ay ? ax : ay;
  float c = mn / (mx + (float)2.2204460492503131e-16);
  float c2 = c * c;
  float a;
  if (fused) {
    a = fmaf(p7, c2, p5);
    a = fmaf(a, c2, p3);
    a = fmaf(a, c2, p1);
    a = a * c;
  } else {
    a = (((p7 * c2 + p5) * c2 + p3) * c2 + p1) * c;
  }
  if (!(ax >= ay)) a = 90.f - a;
  if (x < 0) a = 180.f - a;
  if (y < 0) a = 360.f - a;
  return a;
}

// 16-bucket orientation code of an integer gradient (before `& 7`): the convertTo(CV_8U, 16/360)
// step of [UP] hysteresisGradient: float multiply, round-half-to-even, saturate.
static inline int orientation16(int dx, int dy, int fused) {
  float a = fast_atan2_deg((float)dy, (float)dx, fused);
  float v = a * (float)(16.0 / 360.0);
  long r = lrintf(v);  // default FE_TONEAREST = half-to-even, like cvRound/_mm_cvtss_si32
  if (r < 0) r = 0;
  if (r > 255) r = 255;
  return (int)r;
}

// cv::pyrDown(src, dst, Size(cols/2, rows/2)) on CV_8UC{ch}: 5x5 [1 4 6 4 1]^2,
// BORDER_REFLECT_101, (sum+128)>>8.  
static void pyrdown(const u8* src, int rows, int cols, int ch, u8* dst) {
  int drows = rows / 2, dcols = cols / 2;
  std::vector<u8> pad;
  pad_image(src, rows, cols, ch, 2, 1, pad);
  const int pc = cols + 4, n = dcols * ch;
  std::vector<u16> h((size_t)(rows + 4) * n);  // horizontal sums at even columns, <= 255*16
  for (int y = 0; y < rows + 4; ++y) {
    const u8* p = pad.data() + (size_t)y * pc * ch;
    u16* o = h.data() + (size_t)y * n;
    for (int x = 0; x < dcols; ++x)
      for (int c = 0; c < ch; ++c) {
        const u8* q = p + (size_t)(2 * x) * ch + c;
        o[x * ch + c] = (u16)(q[0] + 4 * q[ch] + 6 * q[2 * ch] + 4 * q[3 * ch] + q[4 * ch]);
      }
  }
  for (int y = 0; y < drows; ++y) {
    const u16 *r0 = h.data() + (size_t)(2 * y) * n, *r1 = r0 + n, *r2 = r1 + n, *r3 = r2 + n, *r4 = r3 + n;
    u8* o = dst + (size_t)y * n;
    for (int i = 0; i < n; ++i) o[i] = (u8)((r0[i] + 4 * r1[i] + 6 * r2[i] + 4 * r3[i] + r4[i] + 128) >> 8);
  }
}

// cv::resize(src, dst, Size(dcols,drows), 0, 0, INTER_NEAREST) on CV_8UC1.
static void resize_nn(const u8* src, int rows, int cols, u8* dst, int drows, int dcols) {
  double ifx = 1.0 / ((double)dcols / cols), ify = 1.0 / ((double)drows / rows);
  for (int y = 0; y < drows; ++y) {
    int sy = std::min((int)std::floor(y * ify), rows - 1);
    for (int x = 0; x < dcols; ++x) {
      int sx = std::min((int)std::floor(x * ifx), cols - 1);
      dst[(size_t)y * dcols + x] = src[(size_t)sy * cols + sx];
    }
  }
}

// cv::medianBlur(src, dst, 5) on CV_8UC1 (replicate border, 13th of 25).
// Maps made of 0 / one-hot bytes (every linemod normal map) take a counting path: per column the 5-row
// window is folded into eight byte-wide label counters, each output adds five column counters.
static void median5(const u8* src, int rows, int cols, u8* dst) {
  std::vector<u8> out((size_t)rows * cols);
  bool onehot = true;
  for (size_t i = 0; i < (size_t)rows * cols && onehot; ++i) onehot = (src[i] & (src[i] - 1)) == 0;
  if (onehot) {
    std::vector<u8> pad;
    pad_image(src, rows, cols, 1, 2, 0, pad);
    const int pc = cols + 4;
    std::vector<uint32_t> lo(pc), hi(pc);
    for (int y = 0; y < rows; ++y) {
      for (int x = 0; x < pc; ++x) {
        uint32_t l = 0, h = 0;
        for (int i = 0; i < 5; ++i) {
          uint32_t v = pad[(size_t)(y + i) * pc + x];
          l += ((v & 15u) * 0x00204081u) & 0x01010101u;
          h += ((v >> 4) * 0x00204081u) & 0x01010101u;
        }
        lo[x] = l; hi[x] = h;
      }
      for (int x = 0; x < cols; ++x) {
        uint32_t l = lo[x] + lo[x + 1] + lo[x + 2] + lo[x + 3] + lo[x + 4];
        uint32_t h = hi[x] + hi[x + 1] + hi[x + 2] + hi[x + 3] + hi[x + 4];
        int c[8] = {(int)(l & 255), (int)((l >> 8) & 255), (int)((l >> 16) & 255), (int)(l >> 24),
                    (int)(h & 255), (int)((h >> 8) & 255), (int)((h >> 16) & 255), (int)(h >> 24)};
        int acc = 25 - (c[0] + c[1] + c[2] + c[3] + c[4] + c[5] + c[6] + c[7]);
        u8 res = 0;
        for (int b = 0; b < 8 && acc < 13; ++b) {
          acc += c[b];
          if (acc >= 13) res = (u8)(1u << b);
        }
        out[(size_t)y * cols + x] = res;
      }
    }
  } else {
    u8 v[25];
    for (int y = 0; y < rows; ++y)
      for (int x = 0; x < cols; ++x) {
        int n = 0;
        for (int i = -2; i <= 2; ++i)
          for (int j = -2; j <= 2; ++j) v[n++] = src[(size_t)clampi(y + i, 0, rows - 1) * cols + clampi(x + j, 0, cols - 1)];
        std::nth_element(v, v + 12, v + 25);
        out[(size_t)y * cols + x] = v[12];
      }
  }
  std::memcpy(dst, out.data(), out.size());
}

// cv::erode(src, dst, Mat(), Point(-1,-1), 1, BORDER_REPLICATE) — 3x3 rect minimum.
static void erode3(const u8* src, int rows, int cols, u8* dst) {
  std::vector<u8> out((size_t)rows * cols);
  for (int y = 0; y < rows; ++y)
    for (int x = 0; x < cols; ++x) {
      u8 m = 255;
      for (int i = -1; i <= 1; ++i)
        for (int j = -1; j <= 1; ++j) m = std::min(m, src[(size_t)clampi(y + i, 0, rows - 1) * cols + clampi(x + j, 0, cols - 1)]);
      out[(size_t)y * cols + x] = m;
    }
  std::memcpy(dst, out.data(), out.size());
}

// cv::distanceTransform(src, dst, DIST_C, 3): chessboard distance to the nearest zero pixel.
// Restates OpenCV's two-pass 3x3 chamfer in 16.16 fixed point with a=b=1 (exact for L-inf);
// an image without any zero pixel yields FLT_MAX like cv2 4.13.
static void dist_c(const u8* src, int rows, int cols, float* dst) {
  const int SHIFT = 16, ONE = 1 << SHIFT;
  const int INIT = std::numeric_limits<int>::max() >> 2;
  const float scale = 1.f / (1 << SHIFT);
  int W = cols + 2;
  bool any_zero = false;
  for (size_t i = 0; i < (size_t)rows * cols && !any_zero; ++i) any_zero = src[i] == 0;
  if (!any_zero) {  // cv2 4.13 returns FLT_MAX everywhere for an image without zero pixels
    for (size_t i = 0; i < (size_t)rows * cols; ++i) dst[i] = std::numeric_limits<float>::max();
    return;
  }
  std::vector<int> d((size_t)(rows + 2) * W, INIT);
  for (int y = 0; y < rows; ++y)
    for (int x = 0; x < cols; ++x) {
      int* p = &d[(size_t)(y + 1) * W + x + 1];
      if (!src[(size_t)y * cols + x]) { *p = 0; continue; }
      *p = std::min(std::min(p[-W - 1], p[-W]), std::min(p[-W + 1], p[-1])) + ONE;
    }
  for (int y = rows - 1; y >= 0; --y)
    for (int x = cols - 1; x >= 0; --x) {
      int* p = &d[(size_t)(y + 1) * W + x + 1];
      if (*p > ONE) {
        int m = std::min(std::min(p[W + 1], p[W]), std::min(p[W - 1], p[1])) + ONE;
        if (m < *p) *p = m;
      }
      dst[(size_t)y * cols + x] = (float)(*p * scale);
    }
}

// ----------------------------------------------------------------------------------
// [UP] quantizedOrientations + hysteresisGradient  (Appendix A.2)
// ----------------------------------------------------------------------------------
static void cg_quantize(const u8* bgr, int rows, int cols, float weak_threshold, int fused,
                        u8* quantized, float* magnitude) {
  size_t n = (size_t)rows * cols;
  std::vector<u8> smoothed(n * 3);
  gauss7(bgr, rows, cols, 3, smoothed.data());
  std::vector<int16_t> dx(n * 3), dy(n * 3);
  sobel3(smoothed.data(), rows, cols, 3, dx.data(), dy.data());
  std::vector<u8> q(n);
  std::vector<float> mag(n);
  for (size_t i = 0; i < n; ++i) {
    int m1 = dx[3 * i] * dx[3 * i] + dy[3 * i] * dy[3 * i];
    int m2 = dx[3 * i + 1] * dx[3 * i + 1] + dy[3 * i + 1] * dy[3 * i + 1];
    int m3 = dx[3 * i + 2] * dx[3 * i + 2] + dy[3 * i + 2] * dy[3 * i + 2];
    int c;
    if (m1 >= m2 && m1 >= m3) c = 0;
    else if (m2 >= m1 && m2 >= m3) c = 1;
    else c = 2;
    int gx = dx[3 * i + c], gy = dy[3 * i + c];
    mag[i] = (float)(gx * gx + gy * gy);
    q[i] = (u8)orientation16(gx, gy, fused);
  }
  // zero first/last row and column, & 7 on the interior
  for (int x = 0; x < cols; ++x) { q[x] = 0; q[(size_t)(rows - 1) * cols + x] = 0; }
  for (int y = 0; y < rows; ++y) { q[(size_t)y * cols] = 0; q[(size_t)y * cols + cols - 1] = 0; }
  for (int y = 1; y < rows - 1; ++y)
    for (int x = 1; x < cols - 1; ++x) q[(size_t)y * cols + x] &= 7;
  std::memset(quantized, 0, n);
  float thr = weak_threshold * weak_threshold;
  for (int y = 1; y < rows - 1; ++y)
    for (int x = 1; x < cols - 1; ++x) {
      if (mag[(size_t)y * cols + x] > thr) {
        int hist[8] = {0, 0, 0, 0, 0, 0, 0, 0};
        for (int i = -1; i <= 1; ++i)
          for (int j = -1; j <= 1; ++j) hist[q[(size_t)(y + i) * cols + x + j]]++;
        int max_votes = 0, index = -1;
        for (int i = 0; i < 8; ++i)
          if (max_votes < hist[i]) { index = i; max_votes = hist[i]; }
        if (max_votes >= 5) quantized[(size_t)y * cols + x] = (u8)(1 << index);
      }
    }
  if (magnitude) std::memcpy(magnitude, mag.data(), n * sizeof(float));
}

// ----------------------------------------------------------------------------------
// [UP] quantizedNormals (+ accumBilateral)  (Appendix A.4)
// idx_out (optional): int8 [3][rows][cols] = (v1,v2,v3) before the LUT, -1 where output is 0.
// Deviation from upstream (N4): LUT indices are clamped to 19 instead of reading out of bounds.
// ----------------------------------------------------------------------------------
static void dn_quantize(const u16* depth, int rows, int cols, int distance_threshold,
                        int difference_threshold, const u8* lut, int do_median, u8* dst, int8_t* idx_out) {
  const int r = 5, G = 20;
  size_t n = (size_t)rows * cols;
  std::memset(dst, 0, n);
  if (idx_out) std::memset(idx_out, 0xff, 3 * n);
  static const int off[8][2] = {{-5, -5}, {0, -5}, {5, -5}, {-5, 0}, {5, 0}, {-5, 5}, {0, 5}, {5, 5}};  // (i=dx, j=dy)
  for (int y = r; y < rows - r - 1; ++y)
    for (int x = r; x < cols - r - 1; ++x) {
      long d = depth[(size_t)y * cols + x];
      if (d >= distance_threshold) continue;
      long A0 = 0, A1 = 0, A3 = 0, b0 = 0, b1 = 0;
      for (int k = 0; k < 8; ++k) {
        long i = off[k][0], j = off[k][1];
        long delta = (long)depth[(size_t)(y + j) * cols + (x + i)] - d;
        long f = std::labs(delta) < difference_threshold ? 1 : 0;
        long fi = f * i, fj = f * j;
        A0 += fi * i; A1 += fi * j; A3 += fj * j;
        b0 += fi * delta; b1 += fj * delta;
      }
      long det = A0 * A3 - A1 * A1;
      long ddx = A3 * b0 - A1 * b1;
      long ddy = -A1 * b0 + A0 * b1;
      float nx = (float)(1150 * ddx), ny = (float)(1150 * ddy), nz = (float)(-det * d);
      float s = sqrtf(nx * nx + ny * ny + nz * nz);
      if (s > 0) {
        float inv = 1.0f / s;
        nx *= inv; ny *= inv; nz *= inv;
        int v1 = (int)(nx * 10 + 10), v2 = (int)(ny * 10 + 10), v3 = (int)(nz * G + G);
        v1 = clampi(v1, 0, G - 1); v2 = clampi(v2, 0, G - 1); v3 = clampi(v3, 0, G - 1);
        dst[(size_t)y * cols + x] = lut ? lut[(v3 * G + v2) * G + v1] : 0;
        if (idx_out) {
          idx_out[(size_t)y * cols + x] = (int8_t)v1;
          idx_out[n + (size_t)y * cols + x] = (int8_t)v2;
          idx_out[2 * n + (size_t)y * cols + x] = (int8_t)v3;
        }
      }
    }
  if (do_median) median5(dst, rows, cols, dst);
}

// ----------------------------------------------------------------------------------
// [UP] spread / SIMILARITY_LUT / computeResponseMaps / linearize  (A.5)
// ----------------------------------------------------------------------------------
static void spread(const u8* src, int rows, int cols, int T, u8* dst) {
  std::memset(dst, 0, (size_t)rows * cols);
  for (int r = 0; r < T; ++r)
    for (int c = 0; c < T; ++c)
      for (int y = 0; y + r < rows; ++y) {
        const u8* s = src + (size_t)(y + r) * cols + c;
        u8* d = dst + (size_t)y * cols;
        for (int x = 0; x + c < cols; ++x) d[x] |= s[x];
      }
}

// SIMILARITY_LUT[32*ori + nibble] (low nibble = orientations 0..3) and [32*ori+16+nibble]
// (high nibble = orientations 4..7): max over set bits j of max(0, 4 - dist(ori, j)).
// circular=1: dist = min(|ori-j|, 8-|ori-j|) — DEFAULT.  The formula SURVEY.md §8c G6 hashed (sum 628, sha1
//             de1dd710...); the 256-entry literal of upstream's table as recalled independently in round 2
//             (tests/golden/similarity_lut_recalled.json) is byte-identical to it: orientations are angles
//             modulo 180 degrees, so bins 0 and 7 are neighbours.
// circular=0: dist = |ori-j| (sum 528) — round 1's default; kept selectable (lmb200_config.similarity_lut = 1).
static void similarity_lut(int circular, u8* out) {
  for (int ori = 0; ori < 8; ++ori)
    for (int half = 0; half < 2; ++half)
      for (int nib = 0; nib < 16; ++nib) {
        int best = 0;
        for (int b = 0; b < 4; ++b)
          if (nib & (1 << b)) {
            int j = b + 4 * half, dd = std::abs(ori - j);
            if (circular) dd = std::min(dd, 8 - dd);
            best = std::max(best, std::max(0, 4 - dd));
          }
        out[32 * ori + 16 * half + nib] = (u8)best;
      }
}

static void response_maps(const u8* spread_img, size_t n, const u8* lut, u8* out /*[8][n]*/) {
  u8 tab[256][8];  // all 8 orientation responses of one spread byte: max(lut_low[s&15], lut_hi[s>>4])
  for (int s = 0; s < 256; ++s)
    for (int ori = 0; ori < 8; ++ori) tab[s][ori] = std::max(lut[32 * ori + (s & 15)], lut[32 * ori + 16 + (s >> 4)]);
  for (int ori = 0; ori < 8; ++ori) {
    u8* o = out + (size_t)ori * n;
    for (size_t i = 0; i < n; ++i) o[i] = tab[spread_img[i]][ori];
  }
}

static void linearize(const u8* resp, int rows, int cols, int T, u8* out /*[T*T][(rows/T)*(cols/T)]*/) {
  int mw = cols / T, mh = rows / T;
  u8* m = out;
  for (int r0 = 0; r0 < T; ++r0)
    for (int c0 = 0; c0 < T; ++c0)
      for (int r = r0; r < rows; r += T)
        for (int c = c0; c < cols; c += T) *m++ = resp[(size_t)r * cols + c];
  (void)mw; (void)mh;
}

// ----------------------------------------------------------------------------------
// Modalities and quantized pyramids — [UP] ColorGradientPyramid / DepthNormalPyramid
// ----------------------------------------------------------------------------------
struct Candidate {
  Feature f;
  float score;
  bool operator<(const Candidate& r) const { return score > r.score; }
};

// [UP] QuantizedPyramid::selectScatteredFeatures
static void select_scattered(const std::vector<Candidate>& cands, std::vector<Feature>& features,
                             size_t num_features, float distance) {
  features.clear();
  float distance_sq = distance * distance;
  int i = 0;
  while (features.size() < num_features) {
    const Candidate& c = cands[i];
    bool keep = true;
    for (int j = 0; j < (int)features.size() && keep; ++j) {
      const Feature& f = features[j];
      keep = (c.f.x - f.x) * (c.f.x - f.x) + (c.f.y - f.y) * (c.f.y - f.y) >= distance_sq;
    }
    if (keep) features.push_back(c.f);
    if (++i == (int)cands.size()) {
      i = 0;
      distance -= 1.0f;
      distance_sq = distance * distance;
    }
  }
}

static inline int get_label(int q) {
  switch (q) {
    case 1: return 0; case 2: return 1; case 4: return 2; case 8: return 3;
    case 16: return 4; case 32: return 5; case 64: return 6; case 128: return 7;
    default: return -1;
  }
}

struct ModalityCfg {
  int type;  // 0 ColorGradient, 1 DepthNormal
  float weak_threshold = 10.0f, strong_threshold = 55.0f;
  int num_features = 63;
  int distance_threshold = 2000, difference_threshold = 50, extract_threshold = 2;
};

struct Pyramid {
  const ModalityCfg* cfg;
  int rows, cols, level = 0;
  int num_features, extract_threshold;
  std::vector<u8> src;        // CG: current BGR level
  std::vector<u8> mask;       // empty = no mask
  std::vector<u8> quant;      // CG: 'angle'; DN: 'normal'
  std::vector<float> magnitude;
  int fused;

  void update_cg() {
    quant.assign((size_t)rows * cols, 0);
    magnitude.assign((size_t)rows * cols, 0.f);
    cg_quantize(src.data(), rows, cols, cfg->weak_threshold, fused, quant.data(), magnitude.data());
  }
  void pyr_down() {  // [UP] ColorGradientPyramid::pyrDown / DepthNormalPyramid::pyrDown
    num_features /= 2;
    ++level;
    int nr = rows / 2, nc = cols / 2;
    if (cfg->type == 0) {
      std::vector<u8> next((size_t)nr * nc * 3);
      pyrdown(src.data(), rows, cols, 3, next.data());
      src.swap(next);
    } else {
      extract_threshold /= 2;
      std::vector<u8> next((size_t)nr * nc);
      resize_nn(quant.data(), rows, cols, next.data(), nr, nc);
      quant.swap(next);
    }
    if (!mask.empty()) {
      std::vector<u8> nm((size_t)nr * nc);
      resize_nn(mask.data(), rows, cols, nm.data(), nr, nc);
      mask.swap(nm);
    }
    rows = nr; cols = nc;
    if (cfg->type == 0) update_cg();
  }
  void quantize(std::vector<u8>& dst) const {  // [UP] ::quantize : zeros, copyTo through mask
    dst.assign((size_t)rows * cols, 0);
    for (size_t i = 0; i < dst.size(); ++i)
      if (mask.empty() || mask[i]) dst[i] = quant[i];
  }
  bool extract(Template& t) const { return cfg->type == 0 ? extract_cg(t) : extract_dn(t); }

  bool extract_cg(Template& templ) const {  // [UP] ColorGradientPyramid::extractTemplate
    size_t n = (size_t)rows * cols;
    std::vector<u8> local;
    if (!mask.empty()) {
      local.resize(n);
      erode3(mask.data(), rows, cols, local.data());
      for (size_t i = 0; i < n; ++i) local[i] = (u8)std::max(0, (int)mask[i] - (int)local[i]);  // cv::subtract saturates
    }
    std::vector<Candidate> cands;
    float thr = cfg->strong_threshold * cfg->strong_threshold;
    for (int r = 0; r < rows; ++r)
      for (int c = 0; c < cols; ++c) {
        size_t i = (size_t)r * cols + c;
        if (local.empty() || local[i]) {
          u8 q = quant[i];
          if (q > 0) {
            float score = magnitude[i];
            if (score > thr) cands.push_back(Candidate{Feature{c, r, get_label(q)}, score});
          }
        }
      }
    if (cands.size() < (size_t)num_features) return false;
    std::stable_sort(cands.begin(), cands.end());
    float distance = (float)(cands.size() / num_features + 1);
    select_scattered(cands, templ.features, num_features, distance);
    templ.width = -1; templ.height = -1; templ.pyramid_level = level;
    return true;
  }

  bool extract_dn(Template& templ) const {  // [UP] DepthNormalPyramid::extractTemplate
    size_t n = (size_t)rows * cols;
    std::vector<u8> local;
    if (!mask.empty()) {
      local.resize(n);
      erode3(mask.data(), rows, cols, local.data());
      erode3(local.data(), rows, cols, local.data());
    }
    std::vector<u8> temp(n, 0);
    std::vector<float> dist[8];
    for (int i = 0; i < 8; ++i) {
      for (size_t p = 0; p < n; ++p)
        if (local.empty() || local[p]) temp[p] = (u8)(1 << i);
      for (size_t p = 0; p < n; ++p) temp[p] &= quant[p];
      dist[i].resize(n);
      dist_c(temp.data(), rows, cols, dist[i].data());
    }
    int label_counts[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    std::vector<Candidate> cands;
    for (int r = 0; r < rows; ++r)
      for (int c = 0; c < cols; ++c) {
        size_t i = (size_t)r * cols + c;
        if (local.empty() || local[i]) {
          u8 q = quant[i];
          if (q != 0 && q != 255) {
            int label = get_label(q);
            if (label < 0) continue;  // upstream would CV_Error; LUTs here are one-hot by contract
            float score = dist[label][i];
            if (score >= extract_threshold) {
              cands.push_back(Candidate{Feature{c, r, label}, score});
              ++label_counts[label];
            }
          }
        }
      }
    if (cands.size() < (size_t)num_features) return false;
    for (auto& c : cands) c.score /= (float)label_counts[c.f.label];
    std::stable_sort(cands.begin(), cands.end());
    float area;
    if (local.empty()) area = (float)n;
    else { size_t nz = 0; for (u8 v : local) nz += v != 0; area = (float)nz; }
    float distance = sqrtf(area) / sqrtf((float)num_features) + 1.5f;
    select_scattered(cands, templ.features, num_features, distance);
    templ.width = -1; templ.height = -1; templ.pyramid_level = level;
    return true;
  }
};

// [UP] cropTemplates
static void crop_templates(TemplatePyramid& tp, int bb[4]) {
  int min_x = std::numeric_limits<int>::max(), min_y = min_x;
  int max_x = std::numeric_limits<int>::min(), max_y = max_x;
  for (auto& t : tp)
    for (auto& f : t.features) {
      int x = f.x << t.pyramid_level, y = f.y << t.pyramid_level;
      min_x = std::min(min_x, x); min_y = std::min(min_y, y);
      max_x = std::max(max_x, x); max_y = std::max(max_y, y);
    }
  if (min_x % 2 == 1) --min_x;
  if (min_y % 2 == 1) --min_y;
  for (auto& t : tp) {
    t.width = (max_x - min_x) >> t.pyramid_level;
    t.height = (max_y - min_y) >> t.pyramid_level;
    int ox = min_x >> t.pyramid_level, oy = min_y >> t.pyramid_level;
    for (auto& f : t.features) { f.x -= ox; f.y -= oy; }
  }
  bb[0] = min_x; bb[1] = min_y; bb[2] = max_x - min_x; bb[3] = max_y - min_y;
}

// ----------------------------------------------------------------------------------
// Linear memories of one (level, modality): contiguous [8][T*T][W*H] (+64 B zero slack so that
// guarded reads in degenerate templates stay inside the vector).
// ----------------------------------------------------------------------------------
struct LevelMem {
  int T, rows, cols, W, H;   // rows/cols = quantized image size; W,H = decimated
  std::vector<u8> lm;        // [8][T*T*W*H]
  size_t per_label() const { return (size_t)T * T * W * H; }
};

// [UP] similarity.  dst: H*W u8 zeros on entry.  Reads outside the label's allocation (only
// possible for hand-made templates whose features lie outside their own bbox; upstream would
// read out of bounds) contribute 0.
static void similarity(const LevelMem& L, const Template& t, u8* dst) {
  int T = L.T, W = L.W, H = L.H;
  int wf = (t.width - 1) / T + 1, hf = (t.height - 1) / T + 1;
  int span_x = W - wf, span_y = H - hf;
  long P = (long)span_y * W + span_x + 1;
  if (P > (long)W * H) P = (long)W * H;  // dst is H*W; upstream would overrun dst for negative sizes only
  size_t per = L.per_label();
  for (const Feature& f : t.features) {
    if (f.x < 0 || f.x >= L.cols || f.y < 0 || f.y >= L.rows) continue;
    size_t base = (size_t)((f.y % T) * T + (f.x % T)) * W * H + (size_t)(f.y / T) * W + f.x / T;
    const u8* lm = L.lm.data() + (size_t)f.label * per;
    long n = P;
    if ((long)base + n > (long)per) n = (long)per - (long)base;  // guard (see above)
    const u8* p = lm + base;
    for (long j = 0; j < n; ++j) dst[j] = (u8)(dst[j] + p[j]);
  }
}

// [UP] similarityLocal.  dst: 256 u8 zeros on entry.
static void similarity_local(const LevelMem& L, const Template& t, u8* dst, int cx, int cy) {
  int T = L.T, W = L.W;
  int ox = (cx / T - 8) * T, oy = (cy / T - 8) * T;
  size_t per = L.per_label();
  for (const Feature& f0 : t.features) {
    int fx = f0.x + ox, fy = f0.y + oy;
    if (fx < 0 || fy < 0 || fx >= L.cols || fy >= L.rows) continue;
    long base = (long)((fy % T) * T + (fx % T)) * W * L.H + (long)(fy / T) * W + fx / T;
    const u8* lm = L.lm.data() + (size_t)f0.label * per;
    for (int r = 0; r < 16; ++r) {
      long rb = base + (long)r * W;
      u8* d = dst + r * 16;
      if (rb + 16 <= (long)per) {
        const u8* p = lm + rb;
        for (int c = 0; c < 16; ++c) d[c] = (u8)(d[c] + p[c]);
      } else {
        for (int c = 0; c < 16; ++c)
          if (rb + c < (long)per) d[c] = (u8)(d[c] + lm[rb + c]);  // guard (N4: upstream UB)
      }
    }
  }
}

struct MatchResult {
  std::vector<Match> final_matches, unsorted, coarse;
  std::vector<std::vector<u8>> quantized;    // [level*M + m]
  std::vector<LevelMem> mems;                // [level*M + m]
  std::vector<int> qrows, qcols;
  double t_frame = 0, t_match = 0, t_sort = 0;
  long long bytes_coarse = 0, bytes_local = 0;
  int error = 0;
};

struct Detector {
  std::vector<ModalityCfg> modalities;
  std::vector<int> T;
  int levels;
  u8 sim_lut[256];
  std::vector<u8> normal_lut;
  int fused_atan = 1;
  std::map<std::string, std::vector<TemplatePyramid>> classes;

  std::vector<std::string> class_ids() const {
    std::vector<std::string> v;
    for (auto& kv : classes) v.push_back(kv.first);
    return v;
  }

  Pyramid process(int m, const void* src, int rows, int cols, const u8* mask) const {
    Pyramid p;
    p.cfg = &modalities[m];
    p.rows = rows; p.cols = cols;
    p.num_features = modalities[m].num_features;
    p.extract_threshold = modalities[m].extract_threshold;
    p.fused = fused_atan;
    if (mask) p.mask.assign(mask, mask + (size_t)rows * cols);
    if (modalities[m].type == 0) {
      p.src.assign((const u8*)src, (const u8*)src + (size_t)rows * cols * 3);
      p.update_cg();
    } else {
      p.quant.assign((size_t)rows * cols, 0);
      dn_quantize((const u16*)src, rows, cols, modalities[m].distance_threshold,
                  modalities[m].difference_threshold, normal_lut.empty() ? nullptr : normal_lut.data(), 1,
                  p.quant.data(), nullptr);
    }
    return p;
  }

  // [UP] Detector::addTemplate
  int add_template(const std::string& class_id, const void* const* srcs, int rows, int cols,
                   const u8* mask, int bb[4]) {
    int M = (int)modalities.size();
    std::vector<TemplatePyramid>& tps = classes[class_id];  // default-inserts the class like upstream
    int template_id = (int)tps.size();
    TemplatePyramid tp(M * levels);
    for (int i = 0; i < M; ++i) {
      Pyramid qp = process(i, srcs[i], rows, cols, mask);
      for (int l = 0; l < levels; ++l) {
        if (l > 0) qp.pyr_down();
        if (!qp.extract(tp[l * M + i])) return -1;
      }
    }
    int b[4];
    crop_templates(tp, b);
    if (bb) std::memcpy(bb, b, sizeof(b));
    tps.push_back(tp);
    return template_id;
  }

  // [UP] Detector::matchClass
  void match_class(const std::vector<LevelMem>& mems, float threshold, int class_index,
                   const std::vector<TemplatePyramid>& tps, int threads, MatchResult& R, bool debug) const {
    int M = (int)modalities.size();
    size_t ntemp = tps.size();
    std::vector<std::vector<Match>> per_t(ntemp), per_t_coarse(debug ? ntemp : 0);
    std::atomic<long long> bytes_coarse_a(0), bytes_local_a(0);
    std::atomic<long> next(0);
    auto worker = [&]() {
    long long bytes_coarse = 0, bytes_local = 0;
    for (;;) {
      long ti0 = next.fetch_add(8);
      if (ti0 >= (long)ntemp) break;
    for (long ti = ti0; ti < std::min((long)ntemp, ti0 + 8); ++ti) {
      const TemplatePyramid& tp = tps[ti];
      int Lc = levels - 1;
      const LevelMem& L0 = mems[Lc * M];
      int W = L0.W, H = L0.H, Tc = L0.T;
      std::vector<u8> sim((size_t)W * H);
      std::vector<u16> total((size_t)W * H, 0);
      int nf = 0;
      for (int m = 0; m < M; ++m) {
        const Template& t = tp[Lc * M + m];
        nf += (int)t.features.size();
        std::fill(sim.begin(), sim.end(), 0);
        similarity(mems[Lc * M + m], t, sim.data());
        for (size_t j = 0; j < sim.size(); ++j) total[j] = (u16)(total[j] + sim[j]);
        int wf = (t.width - 1) / Tc + 1, hf = (t.height - 1) / Tc + 1;
        long P = (long)(H - hf) * W + (W - wf) + 1;
        if (P > 0) bytes_coarse += (long long)P * (long long)t.features.size();
      }
      int raw_threshold = (int)(2 * nf + (threshold / 100.f) * (2 * nf) + 0.5f);
      std::vector<Match> cands;
      for (int r = 0; r < H; ++r)
        for (int c = 0; c < W; ++c) {
          int raw = total[(size_t)r * W + c];
          if (raw > raw_threshold) {
            int offset = Tc / 2 + (Tc % 2 - 1);
            float score = (raw * 100.f) / (4 * nf) + 0.5f;
            cands.push_back(Match{c * Tc + offset, r * Tc + offset, score, class_index, (int)ti});
          }
        }
      if (debug) per_t_coarse[ti] = cands;
      for (int l = levels - 2; l >= 0; --l) {
        const LevelMem& Ll = mems[l * M];
        int T = Ll.T;
        int border = 8 * T, offset = T / 2 + (T % 2 - 1);
        int max_x = Ll.cols - tp[l * M].width - border;
        int max_y = Ll.rows - tp[l * M].height - border;
        u8 loc[256];
        u16 tot[256];
        for (Match& mt : cands) {
          int x = mt.x * 2 + 1, y = mt.y * 2 + 1;
          x = std::max(x, border); y = std::max(y, border);
          x = std::min(x, max_x); y = std::min(y, max_y);
          int nfl = 0;
          std::memset(tot, 0, sizeof(tot));
          for (int m = 0; m < M; ++m) {
            const Template& t = tp[l * M + m];
            nfl += (int)t.features.size();
            std::memset(loc, 0, sizeof(loc));
            similarity_local(mems[l * M + m], t, loc, x, y);
            for (int j = 0; j < 256; ++j) tot[j] = (u16)(tot[j] + loc[j]);
            bytes_local += 256LL * (long long)t.features.size();
          }
          int best = 0, br = -1, bc = -1;
          for (int r = 0; r < 16; ++r)
            for (int c = 0; c < 16; ++c) {
              int s = tot[r * 16 + c];
              if (s > best) { best = s; br = r; bc = c; }
            }
          mt.x = (x / T - 8 + bc) * T + offset;
          mt.y = (y / T - 8 + br) * T + offset;
          mt.similarity = (best * 100.f) / (4 * nfl);
        }
        cands.erase(std::remove_if(cands.begin(), cands.end(),
                                   [threshold](const Match& m) { return m.similarity < threshold; }),
                    cands.end());
      }
      per_t[ti].swap(cands);
    }
    }
    bytes_coarse_a += bytes_coarse;
    bytes_local_a += bytes_local;
    };
    if (threads <= 1) worker();
    else {
      std::vector<std::thread> pool;
      for (int i = 0; i < threads; ++i) pool.emplace_back(worker);
      for (auto& th : pool) th.join();
    }
    long long bytes_coarse = bytes_coarse_a.load(), bytes_local = bytes_local_a.load();
    for (size_t ti = 0; ti < ntemp; ++ti) {
      R.unsorted.insert(R.unsorted.end(), per_t[ti].begin(), per_t[ti].end());
      if (debug) R.coarse.insert(R.coarse.end(), per_t_coarse[ti].begin(), per_t_coarse[ti].end());
    }
    R.bytes_coarse += bytes_coarse;
    R.bytes_local += bytes_local;
  }

  // [UP] Detector::match
  MatchResult* match(const void* const* srcs, int rows, int cols, float threshold,
                     const std::vector<std::string>& class_list, const u8* const* masks, int threads,
                     bool debug) const {
    typedef std::chrono::steady_clock clk;
    MatchResult* R = new MatchResult();
    int M = (int)modalities.size();
    auto t0 = clk::now();
    R->mems.resize((size_t)levels * M);
    R->quantized.resize((size_t)levels * M);
    R->qrows.resize((size_t)levels * M);
    R->qcols.resize((size_t)levels * M);
    // Upstream walks levels outside, modalities inside; the per-modality chains are independent, so with
    // threads > 1 each modality runs its whole pyramid (process, pyrDown, quantize, spread, response,
    // linearize) on its own thread.  Results are identical.
    std::vector<int> errs(M, 0);
    auto chain = [&](int i) {
      Pyramid qp = process(i, srcs[i], rows, cols, masks ? masks[i] : nullptr);
      for (int l = 0; l < levels; ++l) {
        int Tl = T[l];
        if (l > 0) qp.pyr_down();
        std::vector<u8> q;
        qp.quantize(q);
        int r = qp.rows, c = qp.cols;
        if ((r * c) % 16 != 0 || r % Tl != 0 || c % Tl != 0) {  // [UP] CV_Assert in computeResponseMaps/linearize
          errs[i] = -3;
          return;
        }
        size_t n = (size_t)r * c;
        std::vector<u8> sp(n), resp(8 * n);
        spread(q.data(), r, c, Tl, sp.data());
        response_maps(sp.data(), n, sim_lut, resp.data());
        LevelMem& L = R->mems[(size_t)l * M + i];
        L.T = Tl; L.rows = r; L.cols = c; L.W = c / Tl; L.H = r / Tl;
        L.lm.assign(8 * n + 64, 0);
        for (int j = 0; j < 8; ++j) linearize(resp.data() + (size_t)j * n, r, c, Tl, L.lm.data() + (size_t)j * n);
        R->qrows[(size_t)l * M + i] = r;
        R->qcols[(size_t)l * M + i] = c;
        if (debug) R->quantized[(size_t)l * M + i] = q;
      }
    };
    if (threads > 1 && M > 1) {
      std::vector<std::thread> pool;
      for (int i = 0; i < M; ++i) pool.emplace_back(chain, i);
      for (auto& th : pool) th.join();
    } else {
      for (int i = 0; i < M; ++i) chain(i);
    }
    for (int i = 0; i < M; ++i)
      if (errs[i]) { R->error = errs[i]; return R; }
    auto t1 = clk::now();
    std::vector<std::string> ids = class_ids();
    if (class_list.empty()) {
      int ci = 0;
      for (auto& kv : classes) { match_class(R->mems, threshold, ci, kv.second, threads, *R, debug); ++ci; }
    } else {
      for (auto& id : class_list) {
        auto it = classes.find(id);
        if (it != classes.end()) {
          int ci = (int)(std::lower_bound(ids.begin(), ids.end(), id) - ids.begin());
          match_class(R->mems, threshold, ci, it->second, threads, *R, debug);
        }
      }
    }
    auto t2 = clk::now();
    R->final_matches = R->unsorted;
    std::sort(R->final_matches.begin(), R->final_matches.end());
    R->final_matches.erase(std::unique(R->final_matches.begin(), R->final_matches.end()), R->final_matches.end());
    auto t3 = clk::now();
    R->t_frame = std::chrono::duration<double>(t1 - t0).count();
    R->t_match = std::chrono::duration<double>(t2 - t1).count();
    R->t_sort = std::chrono::duration<double>(t3 - t2).count();
    if (!debug) {
      for (auto& m : R->mems) { std::vector<u8>().swap(m.lm); }
      std::vector<Match>().swap(R->unsorted);
    }
    return R;
  }
};

}  // namespace lmo

// ======================================================================================
// C API (ctypes).  Template pyramids travel as flat int32:
//   per template: width, height, pyramid_level, nf, then nf * (x, y, label)
// ======================================================================================
using namespace lmo;

static int decode_pyramid(const int32_t* flat, int n_ints, TemplatePyramid& tp) {
  int p = 0;
  while (p < n_ints) {
    if (p + 4 > n_ints) return -1;
    Template t;
    t.width = flat[p]; t.height = flat[p + 1]; t.pyramid_level = flat[p + 2];
    int nf = flat[p + 3];
    p += 4;
    if (nf < 0 || p + 3 * nf > n_ints) return -1;
    t.features.resize(nf);
    for (int k = 0; k < nf; ++k) { t.features[k] = Feature{flat[p], flat[p + 1], flat[p + 2]}; p += 3; }
    tp.push_back(t);
  }
  return 0;
}

extern "C" {

void lmo_gauss7(const u8* src, int rows, int cols, int ch, u8* dst) { gauss7(src, rows, cols, ch, dst); }
void lmo_sobel3(const u8* src, int rows, int cols, int ch, int16_t* dx, int16_t* dy) { sobel3(src, rows, cols, ch, dx, dy); }
void lmo_fast_atan2(const float* y, const float* x, float* out, long n, int fused) {
  for (long i = 0; i < n; ++i) out[i] = fast_atan2_deg(y[i], x[i], fused);
}
// G1: out[(dy+1020)*2041 + (dx+1020)] = orientation label (0..7) = orientation16 & 7
void lmo_label_table(u8* out, int fused) {
  for (int dy = -1020; dy <= 1020; ++dy)
    for (int dx = -1020; dx <= 1020; ++dx) out[(size_t)(dy + 1020) * 2041 + (dx + 1020)] = (u8)(orientation16(dx, dy, fused) & 7);
}
void lmo_pyrdown(const u8* src, int rows, int cols, int ch, u8* dst) { pyrdown(src, rows, cols, ch, dst); }
void lmo_resize_nn(const u8* src, int rows, int cols, u8* dst, int drows, int dcols) { resize_nn(src, rows, cols, dst, drows, dcols); }
void lmo_median5(const u8* src, int rows, int cols, u8* dst) { median5(src, rows, cols, dst); }
void lmo_erode3(const u8* src, int rows, int cols, u8* dst) { erode3(src, rows, cols, dst); }
void lmo_dist_c(const u8* src, int rows, int cols, float* dst) { dist_c(src, rows, cols, dst); }
void lmo_cg_quantize(const u8* bgr, int rows, int cols, float weak, int fused, u8* q, float* mag) {
  cg_quantize(bgr, rows, cols, weak, fused, q, mag);
}
void lmo_dn_quantize(const u16* depth, int rows, int cols, int dist_thr, int diff_thr, const u8* lut,
                     int do_median, u8* dst, int8_t* idx_out) {
  dn_quantize(depth, rows, cols, dist_thr, diff_thr, lut, do_median, dst, idx_out);
}
void lmo_spread(const u8* src, int rows, int cols, int T, u8* dst) { spread(src, rows, cols, T, dst); }
void lmo_similarity_lut(int circular, u8* out) { similarity_lut(circular, out); }
void lmo_response(const u8* sp, long n, const u8* lut, u8* out) { response_maps(sp, (size_t)n, lut, out); }
void lmo_linearize(const u8* resp, int rows, int cols, int T, u8* out) { linearize(resp, rows, cols, T, out); }

// types[i]: 0 ColorGradient / 1 DepthNormal.  fparams[i*2..]: weak, strong.
// iparams[i*4..]: num_features, distance_threshold, difference_threshold, extract_threshold.
void* lmo_create(int n_mod, const int* types, const float* fparams, const int* iparams, int levels,
                 const int* T, const u8* sim_lut, const u8* normal_lut, int fused_atan) {
  Detector* d = new Detector();
  for (int i = 0; i < n_mod; ++i) {
    ModalityCfg c;
    c.type = types[i];
    c.weak_threshold = fparams[2 * i]; c.strong_threshold = fparams[2 * i + 1];
    c.num_features = iparams[4 * i]; c.distance_threshold = iparams[4 * i + 1];
    c.difference_threshold = iparams[4 * i + 2]; c.extract_threshold = iparams[4 * i + 3];
    d->modalities.push_back(c);
  }
  d->levels = levels;
  d->T.assign(T, T + levels);
  if (sim_lut) std::memcpy(d->sim_lut, sim_lut, 256);
  else similarity_lut(1, d->sim_lut);
  if (normal_lut) d->normal_lut.assign(normal_lut, normal_lut + 8000);
  d->fused_atan = fused_atan;
  return d;
}
void lmo_destroy(void* h) { delete (Detector*)h; }

int lmo_add_template(void* h, const char* class_id, const void* const* srcs, int rows, int cols,
                     const u8* mask, int* bb4) {
  return ((Detector*)h)->add_template(class_id, srcs, rows, cols, mask, bb4);
}
int lmo_add_synthetic(void* h, const char* class_id, const int32_t* flat, int n_ints) {
  Detector* d = (Detector*)h;
  TemplatePyramid tp;
  if (decode_pyramid(flat, n_ints, tp) != 0) return -2;
  if ((int)tp.size() != d->levels * (int)d->modalities.size()) return -2;
  auto& v = d->classes[class_id];
  v.push_back(tp);
  return (int)v.size() - 1;
}
int lmo_num_classes(void* h) { return (int)((Detector*)h)->classes.size(); }
int lmo_class_id(void* h, int idx, char* out, int cap) {
  auto ids = ((Detector*)h)->class_ids();
  if (idx < 0 || idx >= (int)ids.size()) return -1;
  std::snprintf(out, cap, "%s", ids[idx].c_str());
  return (int)ids[idx].size();
}
int lmo_num_templates(void* h, const char* class_id) {
  Detector* d = (Detector*)h;
  if (!class_id) { int n = 0; for (auto& kv : d->classes) n += (int)kv.second.size(); return n; }
  auto it = d->classes.find(class_id);
  return it == d->classes.end() ? 0 : (int)it->second.size();
}
// returns number of int32 needed/written
int lmo_get_template_flat(void* h, const char* class_id, int template_id, int32_t* out, int cap) {
  Detector* d = (Detector*)h;
  auto it = d->classes.find(class_id);
  if (it == d->classes.end() || template_id < 0 || template_id >= (int)it->second.size()) return -1;
  const TemplatePyramid& tp = it->second[template_id];
  int need = 0;
  for (auto& t : tp) need += 4 + 3 * (int)t.features.size();
  if (!out || cap < need) return need;
  int p = 0;
  for (auto& t : tp) {
    out[p++] = t.width; out[p++] = t.height; out[p++] = t.pyramid_level; out[p++] = (int)t.features.size();
    for (auto& f : t.features) { out[p++] = f.x; out[p++] = f.y; out[p++] = f.label; }
  }
  return need;
}

void* lmo_match(void* h, const void* const* srcs, int rows, int cols, float threshold,
                const char* const* class_ids, int n_cls, const u8* const* masks, int threads, int debug) {
  std::vector<std::string> cl;
  for (int i = 0; i < n_cls; ++i) cl.push_back(class_ids[i]);
  if (threads < 1) threads = 1;
  return ((Detector*)h)->match(srcs, rows, cols, threshold, cl, masks, threads, debug != 0);
}
void lmo_result_free(void* r) { delete (MatchResult*)r; }
int lmo_result_error(void* r) { return ((MatchResult*)r)->error; }
static const std::vector<Match>& pick(void* r, int which) {
  MatchResult* R = (MatchResult*)r;
  return which == 0 ? R->final_matches : (which == 1 ? R->unsorted : R->coarse);
}
int lmo_result_count(void* r, int which) { return (int)pick(r, which).size(); }
void lmo_result_get(void* r, int which, int* x, int* y, float* sim, int* cls, int* tid) {
  const std::vector<Match>& v = pick(r, which);
  for (size_t i = 0; i < v.size(); ++i) {
    x[i] = v[i].x; y[i] = v[i].y; sim[i] = v[i].similarity; cls[i] = v[i].class_index; tid[i] = v[i].template_id;
  }
}
int lmo_result_quantized(void* r, int idx, u8* out, int* rows, int* cols) {
  MatchResult* R = (MatchResult*)r;
  if (idx < 0 || idx >= (int)R->quantized.size()) return -1;
  *rows = R->qrows[idx]; *cols = R->qcols[idx];
  if (out && !R->quantized[idx].empty()) std::memcpy(out, R->quantized[idx].data(), R->quantized[idx].size());
  return (int)R->quantized[idx].size();
}
long lmo_result_linmem(void* r, int idx, u8* out) {
  MatchResult* R = (MatchResult*)r;
  if (idx < 0 || idx >= (int)R->mems.size()) return -1;
  const LevelMem& L = R->mems[idx];
  long n = (long)(8 * L.per_label());
  if (out && !L.lm.empty()) std::memcpy(out, L.lm.data(), (size_t)n);
  return n;
}
void lmo_result_stats(void* r, double* out5) {
  MatchResult* R = (MatchResult*)r;
  out5[0] = R->t_frame; out5[1] = R->t_match; out5[2] = R->t_sort;
  out5[3] = (double)R->bytes_coarse; out5[4] = (double)R->bytes_local;
}
// [UP] similarity() per modality + addSimilarities() at the coarsest level: the full u16 map matchClass thresholds
// (SURVEY.md 8d parity gate "similarity maps").  Needs a result of lmo_match(..., debug=1) (linear memories kept).
// out: [H*W] u16, returns H*W; -1 unknown class/template, -2 result holds no linear memories.
long lmo_result_similarity_map(void* h, void* r, const char* class_id, int template_id, u16* out) {
  Detector* d = (Detector*)h;
  MatchResult* R = (MatchResult*)r;
  auto it = d->classes.find(class_id);
  if (it == d->classes.end() || template_id < 0 || template_id >= (int)it->second.size()) return -1;
  const int M = (int)d->modalities.size(), Lc = d->levels - 1;
  if ((int)R->mems.size() != d->levels * M || R->mems[(size_t)Lc * M].lm.empty()) return -2;
  const TemplatePyramid& tp = it->second[template_id];
  const LevelMem& L0 = R->mems[(size_t)Lc * M];
  const size_t n = (size_t)L0.W * L0.H;
  if (!out) return (long)n;
  std::vector<u8> sim(n);
  std::fill(out, out + n, (u16)0);
  for (int m = 0; m < M; ++m) {
    std::fill(sim.begin(), sim.end(), 0);
    similarity(R->mems[(size_t)Lc * M + m], tp[(size_t)Lc * M + m], sim.data());
    for (size_t j = 0; j < n; ++j) out[j] = (u16)(out[j] + sim[j]);
  }
  return (long)n;
}
void lmo_get_similarity_lut(void* h, u8* out256) { std::memcpy(out256, ((Detector*)h)->sim_lut, 256); }

int lmo_max_threads() {
  unsigned n = std::thread::hardware_concurrency();
  return n ? (int)n : 1;
}

}  // extern "C"
