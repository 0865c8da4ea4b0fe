// kernels_color.cu — ColorGradient quantisation, round-2 kernels (sm_100a).
// Upstream: ColorGradientPyramid::update() -> quantizedOrientations() -> hysteresisGradient(), and
// ColorGradientPyramid::pyrDown() -> cv::pyrDown (opencv_contrib rgbd/linemod.cpp).
//
// Round 1's kernel was instruction-bound at 4 % of HBM: a scalar fastAtan2 with an IEEE divide per
// pixel, a horizontal blur that read 16-bit sums one at a time, five shared-memory arrays.  This version
//   * de-interleaves BGR into byte planes while staging, so every filter tap row is a contiguous byte window;
//   * runs the 7x7 Gaussian as dp4a (horizontal, 4 taps per instruction on bytes) + dp2a (vertical, on row PAIRS of the
//     16-bit horizontal sums): the integer sums are exact in any order, so the single (sum + 2^15) >> 16 rounding of
//     OpenCV's fixed-point blur is reproduced bit for bit;
//   * computes the 3x3 Sobel as dp4a with signed weights on the blurred byte planes;
//   * decides the orientation label with INTEGER compares — no float at all: SURVEY golden G1 shows the label is a
//     pure function of the integer (dx, dy); exhaustively over all 2041^2 pairs
//         k = (min*2^20 > max*N1) + (min*2^20 > max*N2),  t = |dy| > |dx| ? 4 - k : k,
//         label = sign(dx) != sign(dy) ? (8 - t) & 7 : t                       (N1 = 208571, N2 = 700819)
//     equals quantize(fastAtan2(dy, dx)) & 7 (tests/test_oracle_primitives.py::test_integer_label_rule pins it to G1);
//   * votes with packed 4-bit counters, column sums shared between the two output rows of a thread.
// pyrDown writes (and the coarser levels read) byte PLANES, so levels >= 1 skip the de-interleave.
#include "kernels.cuh"

namespace lmk {

namespace {

__device__ __forceinline__ int clampc(int v, int lo, int hi) { return v < lo ? lo : (v > hi ? hi : v); }
__device__ __forceinline__ int reflect101c(int p, int n) {
  if ((unsigned)p < (unsigned)n) return p;          // the common case costs one compare
  if (n == 1) return 0;
  while (p < 0 || p >= n) p = p < 0 ? -p : 2 * (n - 1) - p;
  return p;
}
// unsigned bytes of a  x  signed bytes of b, accumulated into c
__device__ __forceinline__ int dp4a_us(u32 a, u32 b, int c) {
  int d;
  asm("dp4a.u32.s32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c));
  return d;
}

// Stages 4 pixels (one word per colour plane) of source row `rp`.  INTERLEAVED: rp points at the BGR row, x4 = first pixel;
// PLANAR: rp points at the row of plane 0, plane_sz = bytes between planes.  border(x) maps an out-of-image column.
template <bool PLANAR, bool REFLECT>
__device__ __forceinline__ void stage4(const u8* __restrict__ rp, size_t plane_sz, int x4, int cols, bool aligned, u32& b, u32& g, u32& r) {
  if (x4 >= 0 && x4 + 3 < cols && aligned) {
    if (PLANAR) {
      b = __ldg(reinterpret_cast<const u32*>(rp + x4));
      g = __ldg(reinterpret_cast<const u32*>(rp + plane_sz + x4));
      r = __ldg(reinterpret_cast<const u32*>(rp + 2 * plane_sz + x4));
    } else {
      const u32* p = reinterpret_cast<const u32*>(rp + 3 * x4);
      const u32 w0 = __ldg(p), w1 = __ldg(p + 1), w2 = __ldg(p + 2);  // B0 G0 R0 B1 | G1 R1 B2 G2 | R2 B3 G3 R3
      b = __byte_perm(__byte_perm(w0, w1, 0x0630), w2, 0x5210);        // B0 B1 B2 | B3
      g = __byte_perm(__byte_perm(w0, w1, 0x0741), w2, 0x6210);        // G0 G1 G2 | G3
      r = __byte_perm(__byte_perm(w0, w1, 0x0052), w2, 0x7410);        // R0 R1 | R2 R3
    }
    return;
  }
  b = g = r = 0;
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const int x = x4 + k;
    const int c = REFLECT ? reflect101c(x, cols) : clampc(x, 0, cols - 1);
    u32 vb, vg, vr;
    if (PLANAR) { vb = rp[c]; vg = rp[plane_sz + c]; vr = rp[2 * plane_sz + c]; }
    else { vb = rp[3 * c]; vg = rp[3 * c + 1]; vr = rp[3 * c + 2]; }
    b |= vb << (8 * k); g |= vg << (8 * k); r |= vr << (8 * k);
  }
}

}  // namespace

// =================================================================================================================
// cv::pyrDown: 5x5 [1 4 6 4 1]^2, BORDER_REFLECT_101, (sum + 128) >> 8, dst = (rows/2, cols/2); output as byte planes.
// Tile = 64 x 16 destination pixels.  Staged source: columns 2*x0-4 .. 2*x0+131 (34 words per plane), rows 2*y0-2 .. 2*y0+33.
// =================================================================================================================
constexpr int PY_TW = 64, PY_TH = 16;
constexpr int PY_SW = 35;                      // staged words per row (34 used) — odd pitch
constexpr int PY_SR = 2 * PY_TH + 4;           // 36 staged rows = 18 row pairs (the last row only pads the last pair)
constexpr int PY_PP = PY_TW + 1;               // pair-row pitch in words

template <bool PLANAR>
__global__ void __launch_bounds__(256) pyrdown_planar_kernel(const u8* __restrict__ src, size_t src_stride, u8* __restrict__ dst,
                                                             size_t dst_stride, int rows, int cols, int drows, int dcols) {
  __shared__ u32 S[3][PY_SR][PY_SW];
  __shared__ u32 P[3][PY_SR / 2][PY_PP];       // horizontal sums of rows (2p, 2p+1) packed lo|hi<<16, one word per destination column
  const int tid = threadIdx.x;
  const int x0 = blockIdx.x * PY_TW, y0 = blockIdx.y * PY_TH;
  const u8* s = src + (size_t)blockIdx.z * src_stride;
  const size_t plane_sz = (size_t)rows * cols;
  const size_t pitch = PLANAR ? (size_t)cols : (size_t)cols * 3;
  const bool aligned = ((reinterpret_cast<size_t>(s) & 3) == 0) && ((pitch & 3) == 0) && (!PLANAR || (plane_sz & 3) == 0);
  for (u32 idx = tid; idx < PY_SR * 34; idx += 256) {
    const int r = idx / 34u, w = idx - r * 34u;
    const int gy = reflect101c(2 * y0 - 2 + r, rows);
    u32 b, g, rr;
    stage4<PLANAR, true>(s + (size_t)gy * pitch, plane_sz, 2 * x0 - 4 + 4 * w, cols, aligned, b, g, rr);
    S[0][r][w] = b; S[1][r][w] = g; S[2][r][w] = rr;
  }
  __syncthreads();
  // horizontal: destination columns (2m, 2m+1) of source rows (2p, 2p+1)
  for (u32 idx = tid; idx < 3 * (PY_SR / 2) * (PY_TW / 2); idx += 256) {
    const u32 m = idx & 31u, t = idx >> 5;
    const u32 ch = t / (u32)(PY_SR / 2), p = t - ch * (u32)(PY_SR / 2);
    u32 h[2][2];
#pragma unroll
    for (int k = 0; k < 2; ++k) {
      const u32* sr = &S[ch][2 * p + k][m];
      const u32 a = sr[0], b = sr[1], c = sr[2];
      // even column 2m: taps at staged bytes 4m+2 .. 4m+6; odd column 2m+1: bytes 4m+4 .. 4m+8
      h[k][0] = __dp4a(__funnelshift_r(a, b, 16), 0x04060401u, __dp4a(b, 0x00010000u, 0u));
      h[k][1] = __dp4a(b, 0x04060401u, __dp4a(c, 0x00000001u, 0u));
    }
    P[ch][p][2 * m] = h[0][0] | (h[1][0] << 16);
    P[ch][p][2 * m + 1] = h[0][1] | (h[1][1] << 16);
  }
  __syncthreads();
  // vertical: destination row yd takes staged rows 2yd .. 2yd+4 = pairs yd, yd+1 and the first half of yd+2
  u8* d = dst + (size_t)blockIdx.z * dst_stride;
  const size_t dplane = (size_t)drows * dcols;
  const bool al_out = ((reinterpret_cast<size_t>(d) & 3) == 0) && ((dcols & 3) == 0) && ((dplane & 3) == 0);
  for (u32 idx = tid; idx < 3 * PY_TH * (PY_TW / 4); idx += 256) {
    const u32 xq = idx & 15u, t = idx >> 4;
    const u32 yd = t & (PY_TH - 1), ch = t >> 4;
    const int gy = y0 + yd, gx = x0 + 4 * xq;
    if (gy >= drows || gx >= dcols) continue;
    u32 pack = 0;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const u32 p0 = P[ch][yd][4 * xq + k], p1 = P[ch][yd + 1][4 * xq + k], p2 = P[ch][yd + 2][4 * xq + k];
      const u32 acc = __dp2a_lo(p0, 0x00000401u, __dp2a_lo(p1, 0x00000406u, __dp2a_lo(p2, 0x00000001u, 128u)));
      pack |= (acc >> 8) << (8 * k);
    }
    u8* o = d + (size_t)ch * dplane + (size_t)gy * dcols + gx;
    if (al_out && gx + 3 < dcols) *reinterpret_cast<u32*>(o) = pack;
    else
      for (int k = 0; k < 4 && gx + k < dcols; ++k) o[k] = (u8)(pack >> (8 * k));
  }
}

void launch_pyrdown_planar(const u8* src, size_t src_stride, bool src_planar, u8* dst, size_t dst_stride, int rows, int cols,
                           int frames, cudaStream_t st) {
  const int drows = rows / 2, dcols = cols / 2;
  dim3 grid((dcols + PY_TW - 1) / PY_TW, (drows + PY_TH - 1) / PY_TH, frames);
  if (src_planar) pyrdown_planar_kernel<true><<<grid, 256, 0, st>>>(src, src_stride, dst, dst_stride, rows, cols, drows, dcols);
  else pyrdown_planar_kernel<false><<<grid, 256, 0, st>>>(src, src_stride, dst, dst_stride, rows, cols, drows, dcols);
}

// =================================================================================================================
// ColorGradient quantisation.  Tile = 64 x 32 output pixels, 256 threads.
//   S   staged source planes   [3][42][20 words]  columns x0-8 .. x0+71, rows y0-5 .. y0+36 (replicate-clamped)
//   Pk  horizontal blur sums   [3][21][18] uint4  row PAIRS (2j, 2j+1) packed lo|hi<<16, columns x0-4 .. x0+67
//   Bs  blurred planes         [3][36][19 words]  columns x0-4 .. x0+67, rows y0-2 .. y0+33
//   Vw  vote words 1 << 4q     [34][68]           columns x0-1 .. x0+64, rows y0-1 .. y0+32   (aliases S)
// =================================================================================================================
constexpr int CQ_TW = 64, CQ_TH = 32;
constexpr int CQ_SR = CQ_TH + 10, CQ_SWD = 20;                 // staged rows, words per row
constexpr int CQ_PJ = CQ_SR / 2, CQ_PW = 18;                   // row pairs, uint4 per pair row
constexpr int CQ_BR = CQ_TH + 4, CQ_BW = 18, CQ_BP = 19;       // blurred rows, words per row, pitch
constexpr int CQ_VR = CQ_TH + 2, CQ_VP = 68;                   // vote rows, pitch (words)
constexpr u32 LBL_N1 = 208571u, LBL_N2 = 700819u;              // tangent thresholds * 2^20 (see the header comment)

struct SobelPx { int dx, dy, mag; };

// orientation label 0..7 from the selected channel's integer gradient (|dx|, |dy| <= 1020)
__device__ __forceinline__ int label_from_gradient(int dx, int dy) {
  const u32 ax = (u32)abs(dx), ay = (u32)abs(dy);
  const u32 mn = min(ax, ay), mx = max(ax, ay);
  const u32 a = mn << 20;
  const int k = (a > mx * LBL_N1) + (a > mx * LBL_N2);
  const int t = ay > ax ? 4 - k : k;
  return ((dx ^ dy) < 0) ? ((8 - t) & 7) : t;
}

// Sobel of the four pixels of blurred word wq in rows (br-1, br, br+1) of one plane; keeps the best channel so far
__device__ __forceinline__ void sobel4(const u32* __restrict__ bp /* &Bs[ch][br-1][wq-1] */, SobelPx* best, bool first) {
  int dx[4] = {0, 0, 0, 0}, dy[4] = {0, 0, 0, 0};
#pragma unroll
  for (int r = 0; r < 3; ++r) {
    const u32 a = bp[r * CQ_BP], b = bp[r * CQ_BP + 1], c = bp[r * CQ_BP + 2];
    const u32 sl = __funnelshift_r(a, b, 24);  // columns -1 0 +1 +2 of the word
    const u32 sr = __funnelshift_r(b, c, 8);   // columns +1 +2 +3 +4
    const u32 wa = r == 1 ? 0x000200FEu : 0x000100FFu, wb = r == 1 ? 0x0200FE00u : 0x0100FF00u;  // (-1 0 1 0) / (0 -1 0 1), doubled in the centre row
    dx[0] = dp4a_us(sl, wa, dx[0]); dx[1] = dp4a_us(sl, wb, dx[1]);
    dx[2] = dp4a_us(sr, wa, dx[2]); dx[3] = dp4a_us(sr, wb, dx[3]);
    if (r != 1) {
      const u32 ya = r == 0 ? 0x00FFFEFFu : 0x00010201u, yb = r == 0 ? 0xFFFEFF00u : 0x01020100u;  // -/+ (1 2 1 0) / (0 1 2 1)
      dy[0] = dp4a_us(sl, ya, dy[0]); dy[1] = dp4a_us(sl, yb, dy[1]);
      dy[2] = dp4a_us(sr, ya, dy[2]); dy[3] = dp4a_us(sr, yb, dy[3]);
    }
  }
#pragma unroll
  for (int p = 0; p < 4; ++p) {
    const int mm = dx[p] * dx[p] + dy[p] * dy[p];
    if (first || mm > best[p].mag) { best[p].mag = mm; best[p].dx = dx[p]; best[p].dy = dy[p]; }  // strict >: B, then G win ties
  }
}

template <bool PLANAR, bool WRITE_MAG>
__global__ void __launch_bounds__(256) cg_quantize2_kernel(const u8* __restrict__ src, size_t src_stride, u8* __restrict__ qout,
                                                           size_t q_stride, float* __restrict__ mag_out, size_t mag_stride,
                                                           int rows, int cols, float weak_sq) {
  __shared__ __align__(16) u32 S[3 * CQ_SR * CQ_SWD];          // 10 080 B; reused as Vw[34][68] (9 248 B)
  __shared__ __align__(16) uint4 Pk[3 * CQ_PJ * CQ_PW];        // 18 144 B
  __shared__ u32 Bs[3 * CQ_BR * CQ_BP];                        //  8 208 B
  u32* Vw = S;
  const int tid = threadIdx.x;
  const int x0 = blockIdx.x * CQ_TW, y0 = blockIdx.y * CQ_TH;
  const u8* s = src + (size_t)blockIdx.z * src_stride;
  const size_t plane_sz = (size_t)rows * cols;
  const size_t pitch = PLANAR ? (size_t)cols : (size_t)cols * 3;
  const bool aligned = ((reinterpret_cast<size_t>(s) & 3) == 0) && ((pitch & 3) == 0) && (!PLANAR || (plane_sz & 3) == 0);

  // 1. source planes, replicate-clamped (GaussianBlur's BORDER_REPLICATE acts on the source)
  for (u32 idx = tid; idx < CQ_SR * CQ_SWD; idx += 256) {
    const int r = idx / (u32)CQ_SWD, w = idx - r * (u32)CQ_SWD;
    const int gy = clampc(y0 - 5 + r, 0, rows - 1);
    u32 b, g, rr;
    stage4<PLANAR, false>(s + (size_t)gy * pitch, plane_sz, x0 - 8 + 4 * w, cols, aligned, b, g, rr);
    S[idx] = b; S[CQ_SR * CQ_SWD + idx] = g; S[2 * CQ_SR * CQ_SWD + idx] = rr;
  }
  __syncthreads();

  // 2. horizontal 7-tap sums ([8 28 56 72 56 28 8], <= 256*255) of staged rows (2j, 2j+1), four columns per item
  for (u32 idx = tid; idx < 3 * CQ_PJ * CQ_PW; idx += 256) {
    const u32 t = idx / (u32)CQ_PW, wq = idx - t * (u32)CQ_PW;
    const u32 ch = t / (u32)CQ_PJ, j = t - ch * (u32)CQ_PJ;
    u32 h[2][4];
#pragma unroll
    for (int k = 0; k < 2; ++k) {
      const u32* sr = S + (ch * CQ_SR + 2 * j + k) * CQ_SWD + wq;   // words wq, wq+1, wq+2 = staged columns 4wq .. 4wq+11
      const u32 a = sr[0], b = sr[1], c = sr[2];
      // output column 4(wq+1)+i takes staged bytes 4wq+1+i .. 4wq+7+i
      h[k][0] = __dp4a(__funnelshift_r(a, b, 8), 0x48381C08u, __dp4a(__funnelshift_r(b, c, 8), 0x00081C38u, 0u));
      h[k][1] = __dp4a(__funnelshift_r(a, b, 16), 0x48381C08u, __dp4a(__funnelshift_r(b, c, 16), 0x00081C38u, 0u));
      h[k][2] = __dp4a(__funnelshift_r(a, b, 24), 0x48381C08u, __dp4a(__funnelshift_r(b, c, 24), 0x00081C38u, 0u));
      h[k][3] = __dp4a(b, 0x48381C08u, __dp4a(c, 0x00081C38u, 0u));
    }
    Pk[idx] = make_uint4(h[0][0] | (h[1][0] << 16), h[0][1] | (h[1][1] << 16), h[0][2] | (h[1][2] << 16), h[0][3] | (h[1][3] << 16));
  }
  __syncthreads();

  // 3. vertical 7-tap + the single rounding -> blurred bytes; item = blurred rows (2jb, 2jb+1) x four columns
  for (u32 idx = tid; idx < 3 * (CQ_BR / 2) * CQ_BW; idx += 256) {
    const u32 t = idx / (u32)CQ_BW, wq = idx - t * (u32)CQ_BW;
    const u32 ch = t / (u32)(CQ_BR / 2), jb = t - ch * (u32)(CQ_BR / 2);
    const uint4* pp = Pk + (ch * CQ_PJ + jb) * CQ_PW + wq;
    const uint4 p0 = pp[0], p1 = pp[CQ_PW], p2 = pp[2 * CQ_PW], p3 = pp[3 * CQ_PW];
    u32 e = 0, o = 0;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const u32 a0 = (&p0.x)[k], a1 = (&p1.x)[k], a2 = (&p2.x)[k], a3 = (&p3.x)[k];
      // blurred row 2jb   = staged rows 2jb   .. 2jb+6 : (8 28)(56 72)(56 28)(8 0)
      // blurred row 2jb+1 = staged rows 2jb+1 .. 2jb+7 : (0 8)(28 56)(72 56)(28 8)
      const u32 se = __dp2a_lo(a0, 0x48381C08u, __dp2a_hi(a1, 0x48381C08u, __dp2a_lo(a2, 0x00081C38u, __dp2a_hi(a3, 0x00081C38u, 32768u))));
      const u32 so = __dp2a_lo(a0, 0x381C0800u, __dp2a_hi(a1, 0x381C0800u, __dp2a_lo(a2, 0x081C3848u, __dp2a_hi(a3, 0x081C3848u, 32768u))));
      e |= (se >> 16) << (8 * k);
      o |= (so >> 16) << (8 * k);
    }
    Bs[(ch * CQ_BR + 2 * jb) * CQ_BP + wq] = e;
    Bs[(ch * CQ_BR + 2 * jb + 1) * CQ_BP + wq] = o;
  }
  __syncthreads();

  // 3b. Sobel's BORDER_REPLICATE acts on the BLURRED image: blurred pixels outside the image copy the clamped one.
  //     Only tiles on the image border do anything: first the (<= 4 + 4) out-of-image columns of every row, then whole
  //     out-of-image rows copy the (already fixed) nearest image row.
  if (x0 < 4 || x0 + CQ_TW + 4 > cols) {
    u8* bb = reinterpret_cast<u8*>(Bs);
    const int first_right = cols - (x0 - 4);               // first blurred column right of the image (>= 72: none)
    for (u32 idx = tid; idx < 3 * CQ_BR * 8; idx += 256) {
      const u32 j = idx & 7u, row = idx >> 3;              // row = ch * CQ_BR + br
      int bc, cx;
      if (j < 4) { bc = (int)j; cx = 4; if (x0 >= 4) continue; }             // columns x0-4 .. x0-1 exist only for x0 == 0
      else { bc = first_right + (int)j - 4; cx = first_right - 1; if (bc >= 4 * CQ_BW) continue; }
      bb[row * (CQ_BP * 4) + bc] = bb[row * (CQ_BP * 4) + cx];
    }
    __syncthreads();
  }
  if (y0 < 2 || y0 + CQ_TH + 2 > rows) {
    const int first_below = rows - (y0 - 2);               // first blurred row below the image (>= 36: none)
    const int n_above = y0 < 2 ? 2 : 0, n_below = first_below < CQ_BR ? min(2, CQ_BR - first_below) : 0;  // only row `rows` is ever read
    for (u32 idx = tid; idx < (u32)(3 * (n_above + n_below) * CQ_BW); idx += 256) {
      const u32 t = idx / (u32)CQ_BW, wq = idx - t * (u32)CQ_BW;
      const u32 ch = t / (u32)(n_above + n_below), k = t - ch * (u32)(n_above + n_below);
      const int br = (int)k < n_above ? (int)k : first_below + ((int)k - n_above);
      const int cy = (int)k < n_above ? 2 : first_below - 1;
      Bs[(ch * CQ_BR + br) * CQ_BP + wq] = Bs[(ch * CQ_BR + cy) * CQ_BP + wq];
    }
    __syncthreads();
  }

  // 4. Sobel + channel select + label -> vote words.  Main items: thread = (word column wc, row pair rp) = pixels
  //    (x0+4wc .. +3) x (y0+2rp, y0+2rp+1); the `strong` bits stay in registers for the vote of the same pixels.
  const int wc = tid & 15, rp = tid >> 4;
  u32 strong = 0;
  {
    SobelPx best[2][4];
#pragma unroll
    for (int ch = 0; ch < 3; ++ch) {
      const u32* bp = Bs + (ch * CQ_BR + 2 * rp + 1) * CQ_BP + wc;  // blurred row (2rp+2)-1 = image row y0+2rp-1; words wc .. wc+2
      sobel4(bp, best[0], ch == 0);
      sobel4(bp + CQ_BP, best[1], ch == 0);
    }
#pragma unroll
    for (int rr = 0; rr < 2; ++rr)
#pragma unroll
      for (int p = 0; p < 4; ++p) {
        const int gy = y0 + 2 * rp + rr, gx = x0 + 4 * wc + p;
        const bool inimg = gy < rows && gx < cols;
        int q = label_from_gradient(best[rr][p].dx, best[rr][p].dy);
        if (!((unsigned)(gy - 1) < (unsigned)(rows - 2) && (unsigned)(gx - 1) < (unsigned)(cols - 2))) q = 0;  // border pixels vote for bin 0
        if ((float)best[rr][p].mag > weak_sq) strong |= 1u << (4 * rr + p);
        Vw[(2 * rp + rr + 1) * CQ_VP + 4 * wc + p + 1] = 1u << (4 * q);
        if (WRITE_MAG && inimg) mag_out[(size_t)blockIdx.z * mag_stride + (size_t)gy * cols + gx] = (float)best[rr][p].mag;
      }
  }
  // halo of the vote region: rows y0-1 and y0+32 (64 + 2 columns each) and columns x0-1, x0+64 (32 rows each): 196 pixels
  if (tid < 2 * (CQ_TW + 2) + 2 * CQ_TH) {
    int vr, vc;
    if (tid < 2 * (CQ_TW + 2)) { vr = tid < CQ_TW + 2 ? 0 : CQ_VR - 1; vc = tid < CQ_TW + 2 ? tid : tid - (CQ_TW + 2); }
    else { const int k = tid - 2 * (CQ_TW + 2); vr = 1 + (k >> 1); vc = (k & 1) ? CQ_TW + 1 : 0; }
    const int gy = y0 - 1 + vr, gx = x0 - 1 + vc;
    int q = 0;
    if (gy > 0 && gy < rows - 1 && gx > 0 && gx < cols - 1) {
      const u8* bb = reinterpret_cast<const u8*>(Bs);
      int bdx = 0, bdy = 0, bm = -1;
#pragma unroll
      for (int ch = 0; ch < 3; ++ch) {
        const u8* c = bb + ((ch * CQ_BR + vr + 1) * CQ_BP) * 4 + vc + 3;  // blurred (row gy, column gx)
        const int r4 = CQ_BP * 4;
        const int dx = (c[-r4 + 1] + 2 * c[1] + c[r4 + 1]) - (c[-r4 - 1] + 2 * c[-1] + c[r4 - 1]);
        const int dy = (c[r4 - 1] + 2 * c[r4] + c[r4 + 1]) - (c[-r4 - 1] + 2 * c[-r4] + c[-r4 + 1]);
        const int mm = dx * dx + dy * dy;
        if (mm > bm) { bm = mm; bdx = dx; bdy = dy; }
      }
      q = label_from_gradient(bdx, bdy);
    }
    Vw[vr * CQ_VP + vc] = 1u << (4 * q);
  }
  __syncthreads();

  // 5. 3x3 vote: >= 5 of 9 (at most one bin can reach 5, so upstream's "first max" never decides), magnitude > weak^2
  {
    u32 cs[2][6];                               // column sums of vote rows (2rp .. 2rp+2) and (2rp+1 .. 2rp+3), columns 4wc .. 4wc+5
    const u32* vp = Vw + (2 * rp) * CQ_VP + 4 * wc;
#pragma unroll
    for (int c = 0; c < 6; ++c) {
      const u32 v0 = vp[c], v1 = vp[CQ_VP + c], v2 = vp[2 * CQ_VP + c], v3 = vp[3 * CQ_VP + c];
      const u32 mid = v1 + v2;
      cs[0][c] = v0 + mid; cs[1][c] = mid + v3;
    }
    u8* qo = qout + (size_t)blockIdx.z * q_stride;
    const bool al_out = ((reinterpret_cast<size_t>(qo) & 3) == 0) && ((cols & 3) == 0);
#pragma unroll
    for (int rr = 0; rr < 2; ++rr) {
      const int gy = y0 + 2 * rp + rr;
      if (gy >= rows) continue;
      u32 pack = 0;
#pragma unroll
      for (int p = 0; p < 4; ++p) {
        const int gx = x0 + 4 * wc + p;
        const u32 hist = cs[rr][p] + cs[rr][p + 1] + cs[rr][p + 2];
        const u32 five = (hist + 0x33333333u) & 0x88888888u;
        u32 out = 0;
        if (five && ((strong >> (4 * rr + p)) & 1u) && (unsigned)(gy - 1) < (unsigned)(rows - 2) && (unsigned)(gx - 1) < (unsigned)(cols - 2))
          out = 1u << ((__ffs((int)five) - 1) >> 2);
        pack |= out << (8 * p);
      }
      const int gx0 = x0 + 4 * wc;
      u8* o = qo + (size_t)gy * cols + gx0;
      if (al_out && gx0 + 3 < cols) *reinterpret_cast<u32*>(o) = pack;
      else
        for (int p = 0; p < 4 && gx0 + p < cols; ++p) o[p] = (u8)(pack >> (8 * p));
    }
  }
}

void launch_cg_quantize2(const u8* src, size_t src_stride, bool src_planar, u8* q, size_t q_stride, float* mag, size_t mag_stride,
                         int rows, int cols, float weak_sq, int frames, cudaStream_t st) {
  dim3 grid((cols + CQ_TW - 1) / CQ_TW, (rows + CQ_TH - 1) / CQ_TH, frames);
  if (src_planar) {
    if (mag) cg_quantize2_kernel<true, true><<<grid, 256, 0, st>>>(src, src_stride, q, q_stride, mag, mag_stride, rows, cols, weak_sq);
    else cg_quantize2_kernel<true, false><<<grid, 256, 0, st>>>(src, src_stride, q, q_stride, mag, mag_stride, rows, cols, weak_sq);
  } else {
    if (mag) cg_quantize2_kernel<false, true><<<grid, 256, 0, st>>>(src, src_stride, q, q_stride, mag, mag_stride, rows, cols, weak_sq);
    else cg_quantize2_kernel<false, false><<<grid, 256, 0, st>>>(src, src_stride, q, q_stride, mag, mag_stride, rows, cols, weak_sq);
  }
}

}  // namespace lmk
