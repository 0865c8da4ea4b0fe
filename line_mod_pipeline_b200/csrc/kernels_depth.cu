// kernels_depth.cu — DepthNormal quantisation, round-2 kernel (sm_100a): quantizedNormals + medianBlur(5) fused.
// Upstream: DepthNormalPyramid ctor -> quantizedNormals(), then medianBlur(dst, dst, 5)
// (opencv_contrib rgbd/linemod.cpp).
//
// Round 1 ran two kernels (one thread per pixel with nine 2-byte global loads, then a 32x8-tile median) and
// round-tripped the raw label map through HBM.  Here one CTA stages the u16 depth tile of a 64x32 output tile
// (halo 5 for the normals + 2 for the median, as aligned 32-bit words), computes the raw labels of the 68x36
// region into shared memory as packed one-hot counters, and takes the 5x5 median with a horizontal 5-sum per row and a
// sliding vertical window per thread (8 output rows per thread).  The float normalisation keeps upstream's exact
// operation order with explicit round-to-nearest intrinsics (no FMA); integer accumulators are 32-bit when every
// intermediate provably fits (same bound as round 1's kernel), else 64-bit like upstream's `long`.
#include "kernels.cuh"

namespace lmk {

constexpr int DM_TW = 64, DM_TH = 32;
constexpr int DM_LW = DM_TW + 4, DM_LH = DM_TH + 4;      // raw-label region (median halo 2)
constexpr int DM_DW = 80, DM_DH = DM_TH + 14;            // staged depth tile: columns x0-8 .. x0+71 (40 aligned words), rows y0-7 .. y0+38
constexpr int DM_LP = DM_LW + 1;                         // label row pitch (uint2 entries)

__device__ __forceinline__ int clampi2(int v, int lo, int hi) { return v < lo ? lo : (v > hi ? hi : v); }

// raw label of one pixel from the staged depth tile (sd: u16, pitch DM_DW); (tx, ty) = position in the tile
template <typename ACC>
__device__ __forceinline__ u32 dn_label(const u16* __restrict__ sd, int tx, int ty, int dist_thr, int diff_thr,
                                        const u8* __restrict__ lut) {
  const u16* c = sd + ty * DM_DW + tx;
  const int d = c[0];
  if (d >= dist_thr) return 0u;
  // neighbours in upstream's order: (i,j) = (-5,-5) (0,-5) (5,-5) (-5,0) (5,0) (-5,5) (0,5) (5,5)
  const int n0 = c[-5 * DM_DW - 5], n1 = c[-5 * DM_DW], n2 = c[-5 * DM_DW + 5], n3 = c[-5], n4 = c[5],
            n5 = c[5 * DM_DW - 5], n6 = c[5 * DM_DW], n7 = c[5 * DM_DW + 5];
  const int e0 = n0 - d, e1 = n1 - d, e2 = n2 - d, e3 = n3 - d, e4 = n4 - d, e5 = n5 - d, e6 = n6 - d, e7 = n7 - d;
  // |e| < thr  <=>  (unsigned)(e + thr - 1) < 2 thr - 1   (the launcher clamps thr to [0, 65536]: |e| <= 65535)
  const u32 lim = diff_thr > 0 ? 2u * (u32)diff_thr - 1u : 0u, off = (u32)diff_thr - 1u;
  const int f0 = (u32)(e0 + (int)off) < lim, f1 = (u32)(e1 + (int)off) < lim, f2 = (u32)(e2 + (int)off) < lim,
            f3 = (u32)(e3 + (int)off) < lim, f4 = (u32)(e4 + (int)off) < lim, f5 = (u32)(e5 + (int)off) < lim,
            f6 = (u32)(e6 + (int)off) < lim, f7 = (u32)(e7 + (int)off) < lim;
  const int m0 = f0 ? e0 : 0, m1 = f1 ? e1 : 0, m2 = f2 ? e2 : 0, m3 = f3 ? e3 : 0, m4 = f4 ? e4 : 0, m5 = f5 ? e5 : 0,
            m6 = f6 ? e6 : 0, m7 = f7 ? e7 : 0;
  // A0 = sum f i^2, A1 = sum f i j, A3 = sum f j^2, b0 = sum f i delta, b1 = sum f j delta with i, j in {-5, 0, 5}
  const ACC A0 = 25 * (f0 + f2 + f3 + f4 + f5 + f7);
  const ACC A3 = 25 * (f0 + f1 + f2 + f5 + f6 + f7);
  const ACC A1 = 25 * (f0 - f2 - f5 + f7);
  const ACC b0 = 5 * (ACC)((m2 + m4 + m7) - (m0 + m3 + m5));
  const ACC b1 = 5 * (ACC)((m5 + m6 + m7) - (m0 + m1 + m2));
  const ACC det = A0 * A3 - A1 * A1;
  const ACC ddx = A3 * b0 - A1 * b1;
  const ACC ddy = -A1 * b0 + A0 * b1;
  float nx = (float)(1150 * ddx), ny = (float)(1150 * ddy), nz = (float)(-det * (ACC)d);
  const float s = __fsqrt_rn(__fadd_rn(__fadd_rn(__fmul_rn(nx, nx), __fmul_rn(ny, ny)), __fmul_rn(nz, nz)));
  if (!(s > 0.f)) return 0u;
  const float inv = __fdiv_rn(1.0f, s);
  nx = __fmul_rn(nx, inv); ny = __fmul_rn(ny, inv); nz = __fmul_rn(nz, inv);
  int v1 = (int)__fadd_rn(__fmul_rn(nx, 10.f), 10.f);
  int v2 = (int)__fadd_rn(__fmul_rn(ny, 10.f), 10.f);
  int v3 = (int)__fadd_rn(__fmul_rn(nz, 20.f), 20.f);
  v1 = clampi2(v1, 0, 19); v2 = clampi2(v2, 0, 19); v3 = clampi2(v3, 0, 19);  // upstream indexes [20] out of bounds there (N4)
  return (u32)__ldg(lut + (v3 * 20 + v2) * 20 + v1);
}

// one-hot (or zero) byte -> eight byte-wide counters (two words): nibble n -> (n * 0x204081) & 0x01010101
__device__ __forceinline__ uint2 onehot_counters(u32 v) {
  return make_uint2(((v & 15u) * 0x00204081u) & 0x01010101u, ((v >> 4) * 0x00204081u) & 0x01010101u);
}

// 13th of 25 from the eight label counters: sorted order is 0 < 1 < 2 < 4 < ... < 128
__device__ __forceinline__ u8 median_from_counters(u32 lo, u32 hi) {
  const u32 plo = lo * 0x01010101u;                  // running sums of labels 0..3
  const u32 tlo = plo >> 24;                         // count of labels 0..3
  const u32 phi = hi * 0x01010101u;                  // running sums of labels 4..7 (without the low half)
  const u32 zeros = 25u - tlo - (phi >> 24);
  const u32 bias = (zeros + 115u) * 0x01010101u;     // + zeros, + (128 - 13)
  const u32 ge_lo = (plo + bias) & 0x80808080u;
  const u32 ge_hi = (phi + tlo * 0x01010101u + bias) & 0x80808080u;
  const int nge = __popc(ge_lo) + __popc(ge_hi);     // labels whose running count has reached 13
  return zeros >= 13u ? (u8)0 : (u8)(1u << (8 - nge));
}

template <typename ACC>
__global__ void __launch_bounds__(256) dn_median_kernel(const u16* __restrict__ depth, size_t depth_stride, u8* __restrict__ out,
                                                        size_t out_stride, int rows, int cols, int dist_thr, int diff_thr,
                                                        const u8* __restrict__ lut) {
  __shared__ __align__(16) u16 sd[DM_DH * DM_DW];        // 46 x 80 u16 = 7 360 B
  __shared__ uint2 lab[DM_LH * DM_LP];                   // 36 x 69 x 8 B = 19 872 B
  const int tid = threadIdx.x;
  const int x0 = blockIdx.x * DM_TW, y0 = blockIdx.y * DM_TH;
  const u16* src = depth + (size_t)blockIdx.z * depth_stride;

  // 1. depth tile as aligned words (2 pixels each); outside the image: 0 (never used: the normals' own 5 px border is skipped)
  const bool al = ((cols & 1) == 0) && ((reinterpret_cast<size_t>(src) & 3) == 0);
  u32* sdw = reinterpret_cast<u32*>(sd);
  for (u32 idx = tid; idx < DM_DH * (DM_DW / 2); idx += 256) {
    const int r = idx / (u32)(DM_DW / 2), w = idx - r * (u32)(DM_DW / 2);
    const int gy = y0 - 7 + r, gx = x0 - 8 + 2 * w;
    u32 v = 0;
    if (gy >= 0 && gy < rows) {
      const u16* rp = src + (size_t)gy * cols;
      if (al && gx >= 0 && gx + 1 < cols) v = __ldg(reinterpret_cast<const u32*>(rp + gx));
      else {
        if (gx >= 0 && gx < cols) v = rp[gx];
        if (gx + 1 >= 0 && gx + 1 < cols) v |= (u32)rp[gx + 1] << 16;
      }
    }
    sdw[idx] = v;
  }
  __syncthreads();

  // 2. raw labels of the 68 x 36 region as counters; pixels outside the image replicate the (always zero) border label
  for (u32 idx = tid; idx < DM_LH * DM_LW; idx += 256) {
    const int ly = idx / (u32)DM_LW, lx = idx - ly * (u32)DM_LW;
    const int gy = y0 - 2 + ly, gx = x0 - 2 + lx;
    u32 v = 0;
    if ((unsigned)(gy - 5) < (unsigned)(rows - 11) && (unsigned)(gx - 5) < (unsigned)(cols - 11))  // [5, n-6); n <= 11: empty
      v = dn_label<ACC>(sd, lx + 6, ly + 5, dist_thr, diff_thr, lut);  // tile coordinates: column gx - (x0-8), row gy - (y0-7)
    lab[ly * DM_LP + lx] = onehot_counters(v);
  }
  __syncthreads();

  // 3. 5x5 median: thread = (column, group of 8 rows); horizontal 5-sums per label row, sliding vertical window
  const int cx = tid & 63, rgp = tid >> 6;               // 64 columns x 4 row groups
  const int gx = x0 + cx;
  if (gx >= cols) return;
  u8* o = out + (size_t)blockIdx.z * out_stride;
  const uint2* lp = lab + (rgp * 8) * DM_LP + cx;        // label row (rgp*8 + k) = image row y0 - 2 + rgp*8 + k; columns cx .. cx+4
  u32 hl[5], hh[5];                                      // horizontal sums of the last five label rows
  u32 wl = 0, wh = 0;
#pragma unroll
  for (int k = 0; k < 12; ++k) {
    const uint2 a = lp[k * DM_LP], b = lp[k * DM_LP + 1], c = lp[k * DM_LP + 2], d = lp[k * DM_LP + 3], e = lp[k * DM_LP + 4];
    const u32 sl = a.x + b.x + c.x + d.x + e.x, sh = a.y + b.y + c.y + d.y + e.y;
    if (k >= 5) { wl -= hl[k % 5]; wh -= hh[k % 5]; }
    hl[k % 5] = sl; hh[k % 5] = sh;
    wl += sl; wh += sh;
    if (k >= 4) {
      const int gy = y0 + rgp * 8 + (k - 4);
      if (gy < rows) o[(size_t)gy * cols + gx] = median_from_counters(wl, wh);
    }
  }
}

void launch_dn_median(const u16* depth, size_t depth_stride, u8* out, size_t out_stride, int rows, int cols, int dist_thr,
                      int diff_thr, const u8* lut_dev, int frames, cudaStream_t st) {
  dim3 grid((cols + DM_TW - 1) / DM_TW, (rows + DM_TH - 1) / DM_TH, frames);
  const bool narrow = diff_thr <= 200 && dist_thr <= 65535;  // 32-bit accumulators are exact (bound in kernels_frame.cu)
  diff_thr = diff_thr < 0 ? 0 : (diff_thr > 65536 ? 65536 : diff_thr);
  if (narrow)
    dn_median_kernel<int><<<grid, 256, 0, st>>>(depth, depth_stride, out, out_stride, rows, cols, dist_thr, diff_thr, lut_dev);
  else
    dn_median_kernel<long long><<<grid, 256, 0, st>>>(depth, depth_stride, out, out_stride, rows, cols, dist_thr, diff_thr, lut_dev);
}

}  // namespace lmk
