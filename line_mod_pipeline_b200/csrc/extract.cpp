// extract.cpp — host side of Detector::addTemplate: feature candidates, scattered selection, crop.
// Restates ColorGradientPyramid::extractTemplate, DepthNormalPyramid::extractTemplate,
// QuantizedPyramid::selectScatteredFeatures and cropTemplates of opencv_contrib rgbd/linemod.cpp
// (SURVEY.md §8a a18, Appendix A.8), fed by the GPU quantisation kernels.  The reference reaches
// this through detector->addTemplate at src/HighLevelLinemod.cpp:93.
// Work is restricted to the mask's bounding box (+1 px), which leaves every result unchanged.
#include <algorithm>
#include <cmath>
#include <cstring>
#include <limits>

#include "detector.h"

namespace lmh {

namespace {
struct Candidate {
  Feature f;
  float score;
  bool operator<(const Candidate& r) const { return score > r.score; }
};

inline int clampi(int v, int lo, int hi) { return v < lo ? lo : (v > hi ? hi : v); }

struct Box { int x0, y0, x1, y1; bool empty() const { return x1 < x0 || y1 < y0; } };  // inclusive

Box nonzero_box(const uint8_t* m, int rows, int cols) {
  Box b{cols, rows, -1, -1};
  for (int y = 0; y < rows; ++y) {
    const uint8_t* r = m + (size_t)y * cols;
    int first = -1, last = -1;
    for (int x = 0; x < cols; ++x)
      if (r[x]) { if (first < 0) first = x; last = x; }
    if (first >= 0) {
      b.x0 = std::min(b.x0, first); b.x1 = std::max(b.x1, last);
      b.y0 = std::min(b.y0, y); b.y1 = y;
    }
  }
  return b;
}

// 3x3 rect erosion, replicate border; only evaluated inside `box`, zero elsewhere (the mask is zero there)
void erode3_box(const uint8_t* src, int rows, int cols, const Box& box, std::vector<uint8_t>& dst) {
  dst.assign((size_t)rows * cols, 0);
  if (box.empty()) return;
  for (int y = box.y0; y <= box.y1; ++y)
    for (int x = box.x0; x <= box.x1; ++x) {
      uint8_t m = 255;
      for (int i = -1; i <= 1; ++i) {
        const uint8_t* r = src + (size_t)clampi(y + i, 0, rows - 1) * cols;
        for (int j = -1; j <= 1; ++j) m = std::min(m, r[clampi(x + j, 0, cols - 1)]);
      }
      dst[(size_t)y * cols + x] = m;
    }
}

void select_scattered(const std::vector<Candidate>& cands, std::vector<Feature>& features, size_t num_features, float distance) {
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

inline int label_of(uint8_t q) {
  return (q && !(q & (q - 1))) ? __builtin_ctz(q) : -1;
}

// Chessboard distance to the nearest zero of a w x h crop (16.16 fixed point, 3x3 chamfer a=b=1):
// cv::distanceTransform(DIST_C, 3).  `whole_image` tells whether the crop is the entire image, the only
// case where "no zero pixel" can happen (cv2 then returns FLT_MAX).
void dist_c_crop(const uint8_t* src, int w, int h, bool whole_image, std::vector<float>& out) {
  const int ONE = 1 << 16, INIT = std::numeric_limits<int>::max() >> 2;
  const float scale = 1.f / (1 << 16);
  out.resize((size_t)w * h);
  if (whole_image) {
    bool any_zero = false;
    for (size_t i = 0; i < (size_t)w * h && !any_zero; ++i) any_zero = src[i] == 0;
    if (!any_zero) { std::fill(out.begin(), out.end(), std::numeric_limits<float>::max()); return; }
  }
  int W = w + 2;
  std::vector<int> d((size_t)(h + 2) * W, INIT);
  for (int y = 0; y < h; ++y)
    for (int x = 0; x < w; ++x) {
      int* p = &d[(size_t)(y + 1) * W + x + 1];
      if (!src[(size_t)y * w + x]) { *p = 0; continue; }
      *p = std::min(std::min(p[-W - 1], p[-W]), std::min(p[-W + 1], p[-1])) + ONE;
    }
  for (int y = h - 1; y >= 0; --y)
    for (int x = w - 1; x >= 0; --x) {
      int* p = &d[(size_t)(y + 1) * W + x + 1];
      if (*p > ONE) {
        int m = std::min(std::min(p[W + 1], p[W]), std::min(p[W - 1], p[1])) + ONE;
        if (m < *p) *p = m;
      }
      out[(size_t)y * w + x] = (float)(*p * scale);
    }
}
}  // namespace

void resize_nn_host(const uint8_t* src, int rows, int cols, uint8_t* dst, int drows, int dcols) {
  double ifx = 1.0 / ((double)dcols / cols), ify = 1.0 / ((double)drows / rows);
  for (int y = 0; y < drows; ++y) {
    int sy = std::min((int)std::floor(y * ify), rows - 1);
    for (int x = 0; x < dcols; ++x) dst[(size_t)y * dcols + x] = src[(size_t)sy * cols + std::min((int)std::floor(x * ifx), cols - 1)];
  }
}

bool extract_color_gradient(const uint8_t* quant, const float* magnitude, const uint8_t* mask, int rows, int cols,
                            int num_features, float strong_threshold, int level, Template& out) {
  Box box{0, 0, cols - 1, rows - 1};
  std::vector<uint8_t> eroded;
  if (mask) {
    box = nonzero_box(mask, rows, cols);
    erode3_box(mask, rows, cols, box, eroded);
  }
  std::vector<Candidate> cands;
  const float thr = strong_threshold * strong_threshold;
  if (!box.empty())
    for (int r = box.y0; r <= box.y1; ++r)
      for (int c = box.x0; c <= box.x1; ++c) {
        size_t i = (size_t)r * cols + c;
        if (mask && !((int)mask[i] - (int)eroded[i] > 0)) continue;  // outline = mask - erode(mask), saturating
        uint8_t q = quant[i];
        if (q > 0) {
          float score = magnitude[i];
          if (score > thr) cands.push_back(Candidate{Feature{c, r, label_of(q)}, score});
        }
      }
  if (num_features <= 0 || cands.size() < (size_t)num_features) return false;
  std::stable_sort(cands.begin(), cands.end());
  float distance = (float)(cands.size() / num_features + 1);
  select_scattered(cands, out.features, num_features, distance);
  out.width = -1; out.height = -1; out.pyramid_level = level;
  return true;
}

bool extract_depth_normal(const uint8_t* quant, const uint8_t* mask, int rows, int cols, int num_features,
                          int extract_threshold, int level, Template& out) {
  Box box{0, 0, cols - 1, rows - 1};
  std::vector<uint8_t> local;
  if (mask) {
    std::vector<uint8_t> e1;
    Box mb = nonzero_box(mask, rows, cols);
    erode3_box(mask, rows, cols, mb, e1);
    erode3_box(e1.data(), rows, cols, mb, local);
    box = nonzero_box(local.data(), rows, cols);
    if (box.empty()) return false;  // no candidate can exist; upstream fails the size check too
  }
  // crop = box grown by one pixel (everything outside `local` is zero in every per-label image)
  const int cx0 = std::max(0, box.x0 - 1), cy0 = std::max(0, box.y0 - 1);
  const int cx1 = std::min(cols - 1, box.x1 + 1), cy1 = std::min(rows - 1, box.y1 + 1);
  const int cw = cx1 - cx0 + 1, chh = cy1 - cy0 + 1;
  const bool whole = (cw == cols && chh == rows);
  std::vector<uint8_t> temp((size_t)cw * chh);
  std::vector<float> dist[8];
  for (int l = 0; l < 8; ++l) {
    for (int y = 0; y < chh; ++y)
      for (int x = 0; x < cw; ++x) {
        size_t i = (size_t)(y + cy0) * cols + (x + cx0);
        bool in = !mask || local[i];
        temp[(size_t)y * cw + x] = in ? (uint8_t)((1 << l) & quant[i]) : (uint8_t)0;
      }
    dist_c_crop(temp.data(), cw, chh, whole, dist[l]);
  }
  int label_counts[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  std::vector<Candidate> cands;
  long area_masked = 0;
  for (int r = box.y0; r <= box.y1; ++r)
    for (int c = box.x0; c <= box.x1; ++c) {
      size_t i = (size_t)r * cols + c;
      if (mask && !local[i]) continue;
      ++area_masked;
      uint8_t q = quant[i];
      if (q != 0 && q != 255) {
        int label = label_of(q);
        if (label < 0) continue;
        float score = dist[label][(size_t)(r - cy0) * cw + (c - cx0)];
        if (score >= (float)extract_threshold) {
          cands.push_back(Candidate{Feature{c, r, label}, score});
          ++label_counts[label];
        }
      }
    }
  if (num_features <= 0 || cands.size() < (size_t)num_features) return false;
  for (auto& c : cands) c.score /= (float)label_counts[c.f.label];
  std::stable_sort(cands.begin(), cands.end());
  float area = mask ? (float)area_masked : (float)((size_t)rows * cols);
  float distance = sqrtf(area) / sqrtf((float)num_features) + 1.5f;
  select_scattered(cands, out.features, num_features, distance);
  out.width = -1; out.height = -1; out.pyramid_level = level;
  return true;
}

void crop_templates(TemplatePyramid& tp, int bb[4]) {
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

}  // namespace lmh
