// kernels_postmatch.cu — the colour check that follows Detector::match in the reference, on the GPU:
//   HighLevelLineMOD::detectTemplate   cvtColor(BGR2HSV) + inRange(lower, upper)            src/HighLevelLinemod.cpp:159-161
//   HighLevelLineMOD::templateMask     convexHull(level-0 features + match offset), fillPoly src/HighLevelLinemod.cpp:113-135
//   HighLevelLineMOD::colorCheck       countNonZero(hue & mask) * 100 / countNonZero(mask)   src/HighLevelLinemod.cpp:424-434
// The reference allocates a full-frame mask per match on the CPU (:131).  Here one kernel turns the resident BGR frame
// into a BIT mask of the in-range pixels (one pass, 3 B in, 1 bit out), and one CTA per match builds the hull, rasterises
// it into a shared-memory bit map with cv::fillPoly's exact rules (8-connected boundary lines + 16.16 fixed-point
// scanline spans) and returns the two pixel counts; the percentage is the caller's integer division.
// The OpenCV primitives are restated from oracle/postmatch.py, which tests/test_oracle_postmatch.py pins bit for bit
// against cv2 (8-bit BGR2HSV with the hsv_shift-12 tables, hull vertex set, fillPoly on polygons inside the image).
#include "kernels.cuh"

namespace lmk {

// ---------------------------------------------------------------------------------------------- BGR2HSV + inRange -> bits
// sdiv[v] = round((255 << 12) / v), hdiv180[d] = round((180 << 12) / (6 d)) (OpenCV's RGB2HSV_b tables, hrange 180)
__constant__ int c_sdiv[256];
__constant__ int c_hdiv[256];

__global__ void __launch_bounds__(256) hsv_inrange_bits_kernel(const u8* __restrict__ bgr, int rows, int cols, int words_per_row,
                                                               u32 lower, u32 upper, u32* __restrict__ bits) {
  const int wx = blockIdx.x * 8 + (threadIdx.x >> 5), y = blockIdx.y, lane = threadIdx.x & 31;
  if (wx >= words_per_row) return;
  const int x = wx * 32 + lane;
  bool in = false;
  if (x < cols) {
    const u8* p = bgr + ((size_t)y * cols + x) * 3;
    const int b = p[0], g = p[1], r = p[2];
    const int v = max(max(b, g), r), vmin = min(min(b, g), r), diff = v - vmin;
    const int vr = v == r ? -1 : 0, vg = v == g ? -1 : 0;
    const int s = (diff * c_sdiv[v] + (1 << 11)) >> 12;
    int hh = (vr & (g - b)) + (~vr & ((vg & (b - r + 2 * diff)) + ((~vg) & (r - g + 4 * diff))));
    hh = (hh * c_hdiv[diff] + (1 << 11)) >> 12;
    hh += hh < 0 ? 180 : 0;
    const int lo_h = lower & 255, lo_s = (lower >> 8) & 255, lo_v = (lower >> 16) & 255;
    const int hi_h = upper & 255, hi_s = (upper >> 8) & 255, hi_v = (upper >> 16) & 255;
    in = hh >= lo_h && hh <= hi_h && s >= lo_s && s <= hi_s && v >= lo_v && v <= hi_v;
  }
  const u32 w = __ballot_sync(0xffffffffu, in);
  if (lane == 0) bits[(size_t)y * words_per_row + wx] = w;
}

// ---------------------------------------------------------------------------------------------- template mask counts
constexpr int PM_MAXPTS = 64 * MAX_MOD;         // level-0 features of every modality
constexpr int PM_SMEM_WORDS = 10240;            // 40 KB bit map: 327 680 pixels of bounding box (640 x 480 fits)

struct PmPoint { int x, y; };

__device__ __forceinline__ long long pm_cross(PmPoint o, PmPoint a, PmPoint b) {
  return (long long)(a.x - o.x) * (b.y - o.y) - (long long)(a.y - o.y) * (b.x - o.x);
}

// result[match] = {inside, total}; (-1, -1): the hull leaves the image or its bounding box exceeds the bit map
__global__ void __launch_bounds__(128) template_mask_count_kernel(const int2* __restrict__ xy, int n_matches,
                                                                  const int* __restrict__ g_of_match, const TplHdr* __restrict__ hdr0,
                                                                  const u32* __restrict__ feat0, int M, int rows, int cols,
                                                                  int words_per_row, const u32* __restrict__ hue_bits,
                                                                  int2* __restrict__ result) {
  __shared__ PmPoint pts[PM_MAXPTS];
  __shared__ PmPoint hull[PM_MAXPTS + 1];
  __shared__ int s_n, s_nh, s_x0, s_y0, s_x1, s_y1;
  __shared__ u32 bm[PM_SMEM_WORDS];
  __shared__ int s_cnt[2];
  const int mi = blockIdx.x, tid = threadIdx.x;
  if (mi >= n_matches) return;
  const int g = g_of_match[mi];
  const int ox = xy[mi].x, oy = xy[mi].y;
  // 1. points = features of the level-0 templates of every modality + offset (invalid/out-of-range features never occur in stored templates)
  if (tid == 0) { s_n = 0; s_cnt[0] = 0; s_cnt[1] = 0; }
  __syncthreads();
  for (int k = tid; k < M * FEAT_SLOTS; k += 128) {
    const int m = k / FEAT_SLOTS, j = k - m * FEAT_SLOTS;
    if (g >= 0 && j < (int)hdr0[g].nf[m]) {
      const u32 f = feat0[((size_t)g * M + m) * FEAT_SLOTS + j];
      if (f >> 31) { const int slot = atomicAdd(&s_n, 1); pts[slot].x = (int)(f & 0x3FFF) + ox; pts[slot].y = (int)((f >> 14) & 0x3FFF) + oy; }
    }
  }
  __syncthreads();
  const int n = s_n;
  // 2. sort lexicographically (x, then y): odd-even transposition sort, n <= 256
  for (int pass = 0; pass < n; ++pass) {
    for (int i = 2 * tid + (pass & 1); i + 1 < n; i += 256) {
      PmPoint a = pts[i], b = pts[i + 1];
      if (a.x > b.x || (a.x == b.x && a.y > b.y)) { pts[i] = b; pts[i + 1] = a; }
    }
    __syncthreads();
  }
  // 3. convex hull — Andrew's monotone chain on the unique points, collinear points dropped (the vertex set of
  //    cv::convexHull); one thread: n <= 256
  if (tid == 0) {
    int m = 0;
    for (int i = 0; i < n; ++i)
      if (m == 0 || pts[i].x != pts[m - 1].x || pts[i].y != pts[m - 1].y) pts[m++] = pts[i];
    int k = 0;
    if (m <= 2) {
      for (int i = 0; i < m; ++i) hull[k++] = pts[i];
    } else {
      for (int i = 0; i < m; ++i) {
        while (k >= 2 && pm_cross(hull[k - 2], hull[k - 1], pts[i]) <= 0) --k;
        hull[k++] = pts[i];
      }
      const int t = k + 1;
      for (int i = m - 2; i >= 0; --i) {
        while (k >= t && pm_cross(hull[k - 2], hull[k - 1], pts[i]) <= 0) --k;
        hull[k++] = pts[i];
      }
      --k;                                             // the chain closes on its first point
    }
    int x0 = 0x7fffffff, y0 = 0x7fffffff, x1 = -0x7fffffff, y1 = -0x7fffffff;
    for (int i = 0; i < k; ++i) { x0 = min(x0, hull[i].x); x1 = max(x1, hull[i].x); y0 = min(y0, hull[i].y); y1 = max(y1, hull[i].y); }
    s_nh = k; s_x0 = k ? (x0 & ~31) : 0; s_y0 = y0; s_x1 = x1; s_y1 = y1;   // bit-map columns start on a 32-pixel boundary of the image
    if (k && x0 < 0) s_x0 = -32;                                          // (marks "outside" below)
  }
  __syncthreads();
  const int nh = s_nh, bx0 = s_x0, by0 = s_y0;
  const int bw = nh > 0 ? (s_x1 - bx0) / 32 + 1 : 0, bh = nh > 0 ? s_y1 - by0 + 1 : 0;   // bit-map words per row, rows
  const bool inside_img = nh > 0 && bx0 >= 0 && by0 >= 0 && s_x1 < cols && s_y1 < rows;
  if (g < 0 || nh == 0 || !inside_img || (long long)bw * bh > PM_SMEM_WORDS) {
    if (tid == 0) result[mi] = make_int2(nh == 0 && g >= 0 ? 0 : -1, nh == 0 && g >= 0 ? 0 : -1);
    return;
  }
  for (int i = tid; i < bw * bh; i += 128) bm[i] = 0u;
  __syncthreads();
  auto setpx = [&](int x, int y) { atomicOr(&bm[(y - by0) * bw + ((x - bx0) >> 5)], 1u << ((x - bx0) & 31)); };
  // 4. cv::fillPoly, part 1 (CollectPolyEdges): every edge as an 8-connected line, left to right (cv::LineIterator)
  for (int e = tid; e < nh; e += 128) {
    const PmPoint p0 = hull[(e + nh - 1) % nh], p1 = hull[e];
    int x1 = p0.x, y1 = p0.y, dx = p1.x - p0.x, dy = p1.y - p0.y, sy = 1;
    if (dx < 0) { dx = -dx; dy = -dy; x1 = p1.x; y1 = p1.y; }
    if (dy < 0) { dy = -dy; sy = -1; }
    const bool vert = dy > dx;
    if (vert) { const int t = dx; dx = dy; dy = t; }
    int err = dx - 2 * dy;
    const int plus = 2 * dx, minus = -2 * dy;
    int x = x1, y = y1;
    for (int i = 0; i <= dx; ++i) {
      setpx(x, y);
      const bool neg = err < 0;
      err += minus + (neg ? plus : 0);
      if (vert) { y += sy; x += neg ? 1 : 0; }
      else { x += 1; y += neg ? sy : 0; }
    }
  }
  // 5. part 2 (FillEdgeCollection): scanline spans between the two active edges of a convex polygon, x in 16.16 fixed
  //    point advancing by dx = trunc((x1 - x0) << 16 / (y1 - y0)) per scanline from the edge's upper end.  Thread = scanline.
  for (int y = by0 + tid; y < s_y1; y += 128) {        // fill runs over y in [y_min, y_max) like upstream
    long long xa = 0, xb = 0;
    int found = 0;
    for (int e = 0; e < nh; ++e) {
      const PmPoint p0 = hull[(e + nh - 1) % nh], p1 = hull[e];
      if (p0.y == p1.y) continue;
      const PmPoint top = p0.y < p1.y ? p0 : p1, bot = p0.y < p1.y ? p1 : p0;
      if (!(top.y <= y && y < bot.y)) continue;
      const long long num = ((long long)p1.x << 16) - ((long long)p0.x << 16), den = p1.y - p0.y;
      const long long q = (num < 0 ? -num : num) / (den < 0 ? -den : den);
      const long long dxf = ((num >= 0) == (den >= 0)) ? q : -q;
      const long long xe = ((long long)top.x << 16) + (long long)(y - top.y) * dxf;
      if (found == 0) xa = xe; else xb = xe;
      ++found;
    }
    if (found == 2) {
      const long long lo = xa < xb ? xa : xb, hi = xa < xb ? xb : xa;
      const int xs = (int)((lo + 65535) >> 16), xe = (int)(hi >> 16);
      for (int x = max(xs, 0); x <= min(xe, cols - 1); ++x) setpx(x, y);
    }
  }
  __syncthreads();
  // 6. counts: |mask| and |mask & hue|
  int tot = 0, ins = 0;
  for (int i = tid; i < bw * bh; i += 128) {
    const u32 w = bm[i];
    if (w) {
      const int ry = i / bw, rx = i - ry * bw;
      tot += __popc(w);
      ins += __popc(w & __ldg(hue_bits + (size_t)(by0 + ry) * words_per_row + (bx0 >> 5) + rx));
    }
  }
  atomicAdd(&s_cnt[0], ins); atomicAdd(&s_cnt[1], tot);
  __syncthreads();
  if (tid == 0) result[mi] = make_int2(s_cnt[0], s_cnt[1]);
}

void postmatch_upload_tables(cudaStream_t st) {
  static bool done[64] = {false};
  int dev = 0;
  cudaGetDevice(&dev);
  if (dev < 0 || dev >= 64 || done[dev]) return;
  int sdiv[256], hdiv[256];
  sdiv[0] = hdiv[0] = 0;
  for (int i = 1; i < 256; ++i) {
    sdiv[i] = (int)((255 << 12) / (1.0 * i) + 0.5);     // cvRound of a positive value
    hdiv[i] = (int)((180 << 12) / (6.0 * i) + 0.5);
  }
  cudaMemcpyToSymbolAsync(c_sdiv, sdiv, sizeof sdiv, 0, cudaMemcpyHostToDevice, st);
  cudaMemcpyToSymbolAsync(c_hdiv, hdiv, sizeof hdiv, 0, cudaMemcpyHostToDevice, st);
  cudaStreamSynchronize(st);
  done[dev] = true;
}

void launch_hsv_inrange_bits(const u8* bgr, int rows, int cols, const u8 lower[3], const u8 upper[3], u32* bits, cudaStream_t st) {
  postmatch_upload_tables(st);
  const int wpr = (cols + 31) / 32;
  dim3 grid((wpr + 7) / 8, rows);
  const u32 lo = lower[0] | (lower[1] << 8) | (lower[2] << 16), hi = upper[0] | (upper[1] << 8) | (upper[2] << 16);
  hsv_inrange_bits_kernel<<<grid, 256, 0, st>>>(bgr, rows, cols, wpr, lo, hi, bits);
}

void launch_template_mask_count(const int2* xy, int n, const int* g_of_match, const TplHdr* hdr0, const u32* feat0,
                                int M, int rows, int cols, const u32* hue_bits, int2* result, cudaStream_t st) {
  if (n <= 0) return;
  template_mask_count_kernel<<<n, 128, 0, st>>>(xy, n, g_of_match, hdr0, feat0, M, rows, cols, (cols + 31) / 32, hue_bits, result);
}

}  // namespace lmk
