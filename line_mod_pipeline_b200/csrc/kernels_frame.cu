// kernels_frame.cu — frame-side kernels (sm_100a): everything `Detector::match` does per frame
// before template scoring.  Each kernel restates one upstream function group bit-exactly
// (opencv_contrib rgbd/linemod.cpp):
//   pyrdown_bgr_kernel        ColorGradientPyramid::pyrDown -> cv::pyrDown            (a4)
//   cg_quantize_kernel        quantizedOrientations + hysteresisGradient, fused       (a2,a3)
//   dn_quantize_kernel        quantizedNormals body                                   (a6)
//   median5_kernel            medianBlur(5) on one-hot bytes                          (a6)
//   resize_nn_kernel          DepthNormalPyramid::pyrDown / mask pyramids             (a7)
//   spread_linearize_kernel   spread + computeResponseMaps + linearize, fused         (a8-a10)
// All are HBM-bound byte kernels: no tensor cores.  Compiled with -fmad=false; float ops that must
// match the CPU bit for bit use explicit round-to-nearest intrinsics.
#include "kernels.cuh"

namespace lmk {

__device__ __forceinline__ int clampi(int v, int lo, int hi) { return v < lo ? lo : (v > hi ? hi : v); }
__device__ __forceinline__ int reflect101(int p, int n) {
  if (n == 1) return 0;
  while (p < 0 || p >= n) p = p < 0 ? -p : 2 * (n - 1) - p;
  return p;
}

// ---------------------------------------------------------------------------------------------
// cv::pyrDown 5x5 [1 4 6 4 1]^2, BORDER_REFLECT_101, (sum+128)>>8, dst = (rows/2, cols/2).
// Separable on the interleaved byte stream (the horizontal taps of a channel are 3 bytes apart, so no
// de-interleave is needed): a CTA stages the 19 source rows of an 8 x 64 output tile as 32-bit words,
// sums them vertically two bytes per 32-bit lane (sums <= 16*255 fit 16 bits), then takes the five
// horizontal taps from the 16-bit row.  The integer sum is the same in either order: bit-exact.
// ---------------------------------------------------------------------------------------------
constexpr int PD_TW = 64, PD_TH = 8;
constexpr int PD_IR = 2 * PD_TH + 3;             // source rows per tile
constexpr int PD_WORDS = (6 * PD_TW + 8 + 4 + 3) / 4;  // source words per row: bytes [6*x0 - 8, 6*x0 + 6*TW + 3]
constexpr int PD_WP = PD_WORDS + 1;

__global__ void __launch_bounds__(256) pyrdown_bgr_kernel(const u8* __restrict__ src, size_t src_stride,
                                                          u8* __restrict__ dst, size_t dst_stride,
                                                          int rows, int cols, int drows, int dcols) {
  __shared__ u32 s_in[PD_IR][PD_WP];
  __shared__ u32 s_v[PD_TH][2 * PD_WP];  // four 16-bit vertical sums per source word
  const int tid = threadIdx.x;
  const int x0 = blockIdx.x * PD_TW, y0 = blockIdx.y * PD_TH;
  const u8* s = src + (size_t)blockIdx.z * src_stride;
  const int rowb = cols * 3;
  const int gb0 = 6 * x0 - 8;  // source byte (within a row) of staged byte 0; negative in the first tile
  const bool al_in = ((reinterpret_cast<size_t>(s) & 3) == 0) && ((rowb & 3) == 0);
  for (int idx = tid; idx < PD_IR * PD_WORDS; idx += 256) {  // (a warp per row measured slower: 19 rows on 8 warps)
    const int r = idx / PD_WORDS, w = idx - r * PD_WORDS;
    const u8* rp = s + (size_t)reflect101(2 * y0 - 2 + r, rows) * rowb;
    const int gb = gb0 + 4 * w;
    u32 v = 0;
    if (al_in && gb >= 0 && gb + 3 < rowb) {
      v = *reinterpret_cast<const u32*>(rp + gb);
    } else {  // words that touch the left/right border: reflect the pixel column byte by byte
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const int g = gb + k;
        const int c = (g + 9) / 3 - 3, ch = g - 3 * c;  // floor division (g >= -8)
        v |= (u32)rp[reflect101(c, cols) * 3 + ch] << (8 * k);
      }
    }
    s_in[r][w] = v;
  }
  __syncthreads();
  for (int idx = tid; idx < PD_TH * PD_WORDS; idx += 256) {
    const int oy = idx / PD_WORDS, w = idx - oy * PD_WORDS;
    const u32 v0 = s_in[2 * oy][w], v1 = s_in[2 * oy + 1][w], v2 = s_in[2 * oy + 2][w], v3 = s_in[2 * oy + 3][w],
              v4 = s_in[2 * oy + 4][w];
    const u32 m = 0x00FF00FFu;
    const u32 e = (v0 & m) + (v4 & m) + 4u * ((v1 & m) + (v3 & m)) + 6u * (v2 & m);                              // bytes 0, 2
    const u32 o = ((v0 >> 8) & m) + ((v4 >> 8) & m) + 4u * (((v1 >> 8) & m) + ((v3 >> 8) & m)) + 6u * ((v2 >> 8) & m);  // bytes 1, 3
    s_v[oy][2 * w] = (e & 0xFFFFu) | (o << 16);
    s_v[oy][2 * w + 1] = (e >> 16) | (o & 0xFFFF0000u);
  }
  __syncthreads();
  const u16* sv = reinterpret_cast<const u16*>(&s_v[0][0]);
  u8* d = dst + (size_t)blockIdx.z * dst_stride;
  const int drowb = dcols * 3;
  const bool al_out = ((reinterpret_cast<size_t>(d) & 3) == 0) && ((drowb & 3) == 0);
  constexpr int QPR = PD_TW * 3 / 4;  // 4-byte output groups per tile row
  for (int idx = tid; idx < PD_TH * QPR; idx += 256) {
    const int oy = idx / QPR, q = idx - oy * QPR;
    const int y = y0 + oy;
    if (y >= drows) continue;
    const u16* vr = sv + oy * (4 * PD_WP);
    u32 pack = 0;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const int ob = 4 * q + k, ox = ob / 3, ch = ob - 3 * ox;
      const u16* vp = vr + 6 * ox + ch + 2;  // staged byte of source column 2*(x0+ox) - 2, channel ch
      const int sum = (int)vp[0] + (int)vp[12] + 4 * ((int)vp[3] + (int)vp[9]) + 6 * (int)vp[6];
      pack |= (u32)((sum + 128) >> 8) << (8 * k);
    }
    const int ob0 = x0 * 3 + 4 * q;  // first output byte of the group within the row
    u8* o = d + (size_t)y * drowb + ob0;
    if (al_out && ob0 + 3 < drowb) {
      *reinterpret_cast<u32*>(o) = pack;
    } else {
#pragma unroll
      for (int k = 0; k < 4; ++k)
        if (ob0 + k < drowb) o[k] = (u8)(pack >> (8 * k));
    }
  }
}

void launch_pyrdown_bgr(const u8* src, size_t src_stride, u8* dst, size_t dst_stride, int rows, int cols,
                        int frames, cudaStream_t st) {
  int drows = rows / 2, dcols = cols / 2;
  dim3 grid((dcols + PD_TW - 1) / PD_TW, (drows + PD_TH - 1) / PD_TH, frames);
  pyrdown_bgr_kernel<<<grid, 256, 0, st>>>(src, src_stride, dst, dst_stride, rows, cols, drows, dcols);
}

// ---------------------------------------------------------------------------------------------
// ColorGradient quantisation, one fused pass per pyramid level.
//   Gaussian 7x7 fixed point ([8,28,56,72,56,28,8], single (sum+2^15)>>16 rounding, replicate)
//   -> Sobel 3x3 on the blurred image (replicate border ON THE BLURRED IMAGE)
//   -> channel with max dx^2+dy^2 (ties: B, then G) -> fastAtan2 degrees -> *16/360 round-half-even
//   -> border zero, &7 -> 3x3 vote (>=5 of 9, first max), magnitude > weak^2 (strict).
// Tile 64x16 output pixels, 256 threads, halo 5 (3 blur + 1 Sobel + 1 vote).  The blur runs on the interleaved byte
// stream: vertical pass first on whole words (two bytes per 32-bit lane), then the horizontal taps 3 sums apart.
// ---------------------------------------------------------------------------------------------
constexpr int CG_TW = 64, CG_TH = 16;
constexpr int CG_SR = CG_TH + 10, CG_SC = CG_TW + 10;                      // source tile (halo 5)
constexpr int CG_SWU = (3 * CG_SC + 1 + 3) / 4, CG_SW = CG_SWU + 1;          // ... as words per row (used, padded)
constexpr int CG_HC = CG_TW + 4, CG_BR = CG_TH + 4;                        // blurred region R2 (halo 2)
constexpr int CG_QR = CG_TH + 2, CG_QC = CG_TW + 2, CG_QCP = CG_QC + 2;    // orientation region R1 (halo 1)
constexpr int CG_BW = CG_HC / 4 + 1, CG_QG = (CG_QC + 3) / 4;              // blurred row in words (+1 spare), R1 pixel quads per row
static_assert(CG_HC % 4 == 0, "blurred rows are read as words");

__device__ __forceinline__ float fast_atan2_deg(float y, float x) {
  // OpenCV hal::fastAtan32f, fused polynomial (bit-exact with the cv2 wheel)
  const float scale = (float)(180.0 / 3.14159265358979323846);
  const float p1 = 0.9997878412794807f * scale, p3 = -0.3258083974640975f * scale;
  const float p5 = 0.1555786518463281f * scale, p7 = -0.04432655554792128f * scale;
  float ax = fabsf(x), ay = fabsf(y);
  float mx = fmaxf(ax, ay), mn = fminf(ax, ay);
  float c = __fdiv_rn(mn, __fadd_rn(mx, (float)2.2204460492503131e-16));
  float c2 = __fmul_rn(c, c);
  float a = __fmaf_rn(p7, c2, p5);
  a = __fmaf_rn(a, c2, p3);
  a = __fmaf_rn(a, c2, p1);
  a = __fmul_rn(a, c);
  if (!(ax >= ay)) a = __fsub_rn(90.f, a);
  if (x < 0) a = __fsub_rn(180.f, a);
  if (y < 0) a = __fsub_rn(360.f, a);
  return a;
}

__global__ void __launch_bounds__(256) cg_quantize_kernel(const u8* __restrict__ bgr, size_t bgr_stride,
                                                          u8* __restrict__ qout, size_t q_stride,
                                                          float* __restrict__ mag_out, size_t mag_stride,
                                                          int rows, int cols, float weak_sq) {
  __shared__ u32 s_srcw[CG_SR][CG_SW];  // interleaved BGR bytes of the source tile, staged as aligned words
  __shared__ u32 s_vw[CG_BR][2 * CG_SW];  // vertical blur sums, one u16 per staged source byte
  __shared__ u32 s_bw[3][CG_BR][CG_BW];  // blurred planes, 4 pixels per word (CG_HC = 68 columns + one spare word)
  __shared__ u32 s_q[CG_QR][CG_QCP];  // orientation code as a vote: 1 << (4 * code)
  __shared__ int s_m[CG_QR][CG_QCP];

  const int tid = threadIdx.x;
  const int x0 = blockIdx.x * CG_TW, y0 = blockIdx.y * CG_TH;
  const u8* src = bgr + (size_t)blockIdx.z * bgr_stride;

  // 1. source tile, replicate-clamped (the blur's border mode acts on the source).  The interleaved bytes are
  //    staged as they lie in memory: staged byte 0 = row byte 3*x0 - 16 (word aligned), so pixel tx (image column
  //    x0 - 5 + tx) channel ch is staged byte 3*tx + ch + 1.  Words inside the row are one 32-bit load; words that
  //    touch the left/right border are assembled byte by byte from the clamped pixel column.
  const int rowb = cols * 3;
  const bool al_in = ((reinterpret_cast<size_t>(src) & 3) == 0) && ((rowb & 3) == 0);
  for (int ty = tid >> 5; ty < CG_SR; ty += 8) {  // a warp per source row
    const u8* rp = src + (size_t)clampi(y0 - 5 + ty, 0, rows - 1) * rowb;
    for (int w = tid & 31; w < CG_SWU; w += 32) {
      const int gb = 3 * x0 - 16 + 4 * w;
      u32 v = 0;
      if (al_in && gb >= 0 && gb + 3 < rowb) {
        v = *reinterpret_cast<const u32*>(rp + gb);
      } else {
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const int g = gb + k;
          const int c = (g + 18) / 3 - 6, ch = g - 3 * c;  // floor division (g >= -16)
          v |= (u32)rp[clampi(c, 0, cols - 1) * 3 + ch] << (8 * k);
        }
      }
      s_srcw[ty][w] = v;
    }
  }
  __syncthreads();

  // 2. vertical blur sums at the CLAMPED centre row (so out-of-image R2 rows replicate the blurred border row, which
  //    is what Sobel's BORDER_REPLICATE sees).  Two bytes per 32-bit lane: a 16-bit lane holds at most 256*255.
  for (int ty2 = tid >> 5; ty2 < CG_BR; ty2 += 8) {
    const int tr = clampi(y0 - 2 + ty2, 0, rows - 1) - (y0 - 5);
    for (int w = tid & 31; w < CG_SWU; w += 32) {
      const u32 v0 = s_srcw[tr - 3][w], v1 = s_srcw[tr - 2][w], v2 = s_srcw[tr - 1][w], v3 = s_srcw[tr][w],
                v4 = s_srcw[tr + 1][w], v5 = s_srcw[tr + 2][w], v6 = s_srcw[tr + 3][w];
      const u32 m = 0x00FF00FFu;
      const u32 e = 8u * ((v0 & m) + (v6 & m)) + 28u * ((v1 & m) + (v5 & m)) + 56u * ((v2 & m) + (v4 & m)) + 72u * (v3 & m);
      const u32 o = 8u * (((v0 >> 8) & m) + ((v6 >> 8) & m)) + 28u * (((v1 >> 8) & m) + ((v5 >> 8) & m)) +
                    56u * (((v2 >> 8) & m) + ((v4 >> 8) & m)) + 72u * ((v3 >> 8) & m);
      s_vw[ty2][2 * w] = (e & 0xFFFFu) | (o << 16);          // staged bytes 4w, 4w+1
      s_vw[ty2][2 * w + 1] = (e >> 16) | (o & 0xFFFF0000u);  // staged bytes 4w+2, 4w+3
    }
  }
  __syncthreads();

  // 3. horizontal pass at the clamped centre column -> blurred u8 on R2 (taps of one channel are 3 sums apart)
  const u16* s_v = reinterpret_cast<const u16*>(&s_vw[0][0]);
  u8* s_bb = reinterpret_cast<u8*>(&s_bw[0][0][0]);
  for (int idx = tid; idx < CG_BR * CG_HC; idx += 256) {
    int ty2 = idx / CG_HC, tx2 = idx - ty2 * CG_HC;
    int cx = clampi(x0 - 2 + tx2, 0, cols - 1);
    int tc = cx - (x0 - 5);
#pragma unroll
    for (int ch = 0; ch < 3; ++ch) {
      const u16* r = s_v + ty2 * (4 * CG_SW) + 3 * (tc - 3) + ch + 1;
      int s = 8 * ((int)r[0] + r[18]) + 28 * ((int)r[3] + r[15]) + 56 * ((int)r[6] + r[12]) + 72 * (int)r[9];
      s_bb[(ch * CG_BR + ty2) * (4 * CG_BW) + tx2] = (u8)((s + 32768) >> 16);
    }
  }
  __syncthreads();

  // 4. Sobel + channel select + orientation code on R1, four pixels per thread.  The 3x3 Sobel sums of four
  //    neighbouring pixels are formed two per 32-bit word (16-bit lanes, biased by 1024 so a lane never borrows):
  //    E/O = even/odd columns of the word, X = the two columns after it.
  for (int idx = tid; idx < CG_QR * CG_QG; idx += 256) {
    const int ty1 = idx / CG_QG, k = idx - ty1 * CG_QG;
    const int gy = y0 - 1 + ty1;
    int bdx[4], bdy[4], bm[4] = {-1, -1, -1, -1};
#pragma unroll
    for (int ch = 0; ch < 3; ++ch) {
      const u32* r0 = &s_bw[ch][ty1][k];  // R2 rows ty1, ty1+1, ty1+2 = image rows gy-1, gy, gy+1; word k = R2 columns 4k..4k+3
      const u32 a0 = r0[0], b0 = r0[1], a1 = r0[CG_BW], b1 = r0[CG_BW + 1], a2 = r0[2 * CG_BW], b2 = r0[2 * CG_BW + 1];
      const u32 m = 0x00FF00FFu;
      const u32 E0 = a0 & m, O0 = (a0 >> 8) & m, X0 = __byte_perm(b0, 0u, 0x4140);
      const u32 E1 = a1 & m, O1 = (a1 >> 8) & m, X1 = __byte_perm(b1, 0u, 0x4140);
      const u32 E2 = a2 & m, O2 = (a2 >> 8) & m, X2 = __byte_perm(b2, 0u, 0x4140);
      // dx: vertical [1 2 1] per column, then right column - left column
      const u32 VE = E0 + 2u * E1 + E2, VO = O0 + 2u * O1 + O2, VX = X0 + 2u * X1 + X2;  // columns (4k,4k+2) (4k+1,4k+3) (4k+4,4k+5)
      const u32 VE2 = (VE >> 16) | (VX << 16);          // columns (4k+2, 4k+4)
      const u32 VO2 = (VO >> 16) | (VX & 0xFFFF0000u);  // columns (4k+3, 4k+5)
      const u32 dxE = VE2 + 0x04000400u - VE, dxO = VO2 + 0x04000400u - VO;  // pixels (0,2) and (1,3), + 1024
      // dy: bottom row - top row per column (+256), then horizontal [1 2 1] (+1024)
      const u32 TE = E2 + 0x01000100u - E0, TO = O2 + 0x01000100u - O0, TX = X2 + 0x01000100u - X0;
      const u32 TE2 = (TE >> 16) | (TX << 16), TO2 = (TO >> 16) | (TX & 0xFFFF0000u);
      const u32 dyE = TE + 2u * TO + TE2, dyO = TO + 2u * TE2 + TO2;
      const int dx[4] = {(int)(dxE & 0xFFFFu) - 1024, (int)(dxO & 0xFFFFu) - 1024, (int)(dxE >> 16) - 1024, (int)(dxO >> 16) - 1024};
      const int dy[4] = {(int)(dyE & 0xFFFFu) - 1024, (int)(dyO & 0xFFFFu) - 1024, (int)(dyE >> 16) - 1024, (int)(dyO >> 16) - 1024};
#pragma unroll
      for (int p = 0; p < 4; ++p) {
        const int mm = dx[p] * dx[p] + dy[p] * dy[p];
        if (mm > bm[p]) { bm[p] = mm; bdx[p] = dx[p]; bdy[p] = dy[p]; }  // strict >: earlier channel wins ties (B, G, R)
      }
    }
#pragma unroll
    for (int p = 0; p < 4; ++p) {
      const int tx1 = 4 * k + p;
      if (tx1 >= CG_QC) break;
      const int gx = x0 - 1 + tx1;
      int q = 0, mag = 0;
      if (gy >= 0 && gy < rows && gx >= 0 && gx < cols) {
        mag = bm[p];
        if (gy > 0 && gy < rows - 1 && gx > 0 && gx < cols - 1) {
          float ang = fast_atan2_deg((float)bdy[p], (float)bdx[p]);
          int r = __float2int_rn(__fmul_rn(ang, (float)(16.0 / 360.0)));
          r = r < 0 ? 0 : (r > 255 ? 255 : r);
          q = r & 7;
        }
      }
      s_q[ty1][tx1] = 1u << (4 * q);
      s_m[ty1][tx1] = mag;
    }
  }
  __syncthreads();

  // 5. 3x3 vote
  u8* qo = qout + (size_t)blockIdx.z * q_stride;
  for (int idx = tid; idx < CG_TH * CG_TW; idx += 256) {
    int ty = idx / CG_TW, tx = idx - ty * CG_TW;
    int gy = y0 + ty, gx = x0 + tx;
    if (gy >= rows || gx >= cols) continue;
    int mag = s_m[ty + 1][tx + 1];
    u8 out = 0;
    if (gy > 0 && gy < rows - 1 && gx > 0 && gx < cols - 1 && (float)mag > weak_sq) {
      u32 hist = 0;  // 8 x 4-bit counters (9 votes)
#pragma unroll
      for (int i = 0; i < 3; ++i)
#pragma unroll
        for (int j = 0; j < 3; ++j) hist += s_q[ty + i][tx + j];
      // upstream takes the first bin with the most votes and accepts it with >= 5 of 9: at most one bin can
      // have 5, so "first max" never matters — +3 carries every counter >= 5 into its bit 3.
      const u32 five = (hist + 0x33333333u) & 0x88888888u;
      if (five) out = (u8)(1u << ((__ffs((int)five) - 1) >> 2));
    }
    qo[(size_t)gy * cols + gx] = out;
    if (mag_out) mag_out[(size_t)blockIdx.z * mag_stride + (size_t)gy * cols + gx] = (float)mag;
  }
}

void launch_cg_quantize(const u8* bgr, size_t bgr_stride, u8* q, size_t q_stride, float* mag, size_t mag_stride,
                        int rows, int cols, float weak_sq, int frames, cudaStream_t st) {
  dim3 grid((cols + CG_TW - 1) / CG_TW, (rows + CG_TH - 1) / CG_TH, frames);
  cg_quantize_kernel<<<grid, 256, 0, st>>>(bgr, bgr_stride, q, q_stride, mag, mag_stride, rows, cols, weak_sq);
}

// ---------------------------------------------------------------------------------------------
// DepthNormal: quantizedNormals body.  64-bit integer accumulators like upstream's `long`;
// float normalisation with explicit rn intrinsics (no FMA); C truncation; LUT indices clamped
// to 19 (upstream reads out of bounds there — N4).  Writes every pixel (0 on border/failure).
// ---------------------------------------------------------------------------------------------
// ACC = int when every intermediate provably fits 32 bits (difference_threshold <= 200, distance_threshold
// <= 65535: |1150*ddx| <= 1150*(150+100)*30*199 < 2^31, det*d <= 22500*65535 < 2^31), else long long.
template <typename ACC>
__global__ void __launch_bounds__(256) dn_quantize_kernel(const u16* __restrict__ depth, size_t depth_stride,
                                                          u8* __restrict__ out, size_t out_stride,
                                                          int8_t* __restrict__ idx_out, int rows, int cols,
                                                          int dist_thr, int diff_thr, const u8* __restrict__ lut) {
  const int x = blockIdx.x * 32 + threadIdx.x, y = blockIdx.y * 8 + threadIdx.y;
  if (x >= cols || y >= rows) return;
  const u16* d = depth + (size_t)blockIdx.z * depth_stride;
  const size_t n = (size_t)rows * cols, pix = (size_t)y * cols + x;
  u8 res = 0;
  int v1 = -1, v2 = -1, v3 = -1;
  const int r = 5;
  if (y >= r && y < rows - r - 1 && x >= r && x < cols - r - 1) {
    const ACC dc = d[pix];
    if (dc < dist_thr) {
      ACC A0 = 0, A1 = 0, A3 = 0, b0 = 0, b1 = 0;
      const int oi[8] = {-5, 0, 5, -5, 5, -5, 0, 5};
      const int oj[8] = {-5, -5, -5, 0, 0, 5, 5, 5};
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        const ACC delta = (ACC)d[(size_t)(y + oj[k]) * cols + (x + oi[k])] - dc;
        if ((delta < 0 ? -delta : delta) < diff_thr) {  // f = 1
          A0 += oi[k] * oi[k]; A1 += oi[k] * oj[k]; A3 += oj[k] * oj[k];
          b0 += oi[k] * delta; b1 += oj[k] * delta;
        }
      }
      const ACC det = A0 * A3 - A1 * A1;
      const ACC ddx = A3 * b0 - A1 * b1;
      const ACC ddy = -A1 * b0 + A0 * b1;
      float nx = (float)(1150 * ddx), ny = (float)(1150 * ddy), nz = (float)(-det * dc);
      float s = __fsqrt_rn(__fadd_rn(__fadd_rn(__fmul_rn(nx, nx), __fmul_rn(ny, ny)), __fmul_rn(nz, nz)));
      if (s > 0) {
        float inv = __fdiv_rn(1.0f, s);
        nx = __fmul_rn(nx, inv); ny = __fmul_rn(ny, inv); nz = __fmul_rn(nz, inv);
        v1 = (int)__fadd_rn(__fmul_rn(nx, 10.f), 10.f);
        v2 = (int)__fadd_rn(__fmul_rn(ny, 10.f), 10.f);
        v3 = (int)__fadd_rn(__fmul_rn(nz, 20.f), 20.f);
        v1 = clampi(v1, 0, 19); v2 = clampi(v2, 0, 19); v3 = clampi(v3, 0, 19);
        res = lut[(v3 * 20 + v2) * 20 + v1];
      }
    }
  }
  out[(size_t)blockIdx.z * out_stride + pix] = res;
  if (idx_out) {
    int8_t* io = idx_out + (size_t)blockIdx.z * 3 * n;
    io[pix] = (int8_t)v1; io[n + pix] = (int8_t)v2; io[2 * n + pix] = (int8_t)v3;
  }
}

void launch_dn_quantize(const u16* depth, size_t depth_stride, u8* out, size_t out_stride, int8_t* idx_out,
                        int rows, int cols, int dist_thr, int diff_thr, const u8* lut_dev, int frames,
                        cudaStream_t st) {
  dim3 grid((cols + 31) / 32, (rows + 7) / 8, frames), block(32, 8);
  if (diff_thr <= 200 && dist_thr <= 65535)
    dn_quantize_kernel<int><<<grid, block, 0, st>>>(depth, depth_stride, out, out_stride, idx_out, rows, cols, dist_thr, diff_thr, lut_dev);
  else
    dn_quantize_kernel<long long><<<grid, block, 0, st>>>(depth, depth_stride, out, out_stride, idx_out, rows, cols, dist_thr, diff_thr, lut_dev);
}

// ---------------------------------------------------------------------------------------------
// medianBlur(5), replicate border, on bytes that are 0 or one-hot: the sorted order is
// 0 < 1 < 2 < 4 < ... < 128, so the 13th of 25 falls out of nine 5-bit counters.
// ---------------------------------------------------------------------------------------------
// Counting median.  Every pixel is expanded ONCE into eight byte-wide one-hot counters (two u32: a one-hot byte's
// nibble n becomes (n*0x204081)&0x01010101, one counter byte per set bit); a column entry adds five of them
// vertically, an output adds five column entries.  The 13th of 25 then falls out of packed prefix sums:
// c*0x01010101 turns four byte counts into their running sums, adding (128-13) to every byte sets bit 7 exactly
// where the running count (zeros first, then labels 0..7: the sorted order 0 < 1 < 2 < 4 < ... < 128) reaches 13.
__global__ void __launch_bounds__(256) median5_kernel(const u8* __restrict__ src, size_t src_stride,
                                                      u8* __restrict__ dst, size_t dst_stride, int rows, int cols) {
  __shared__ uint2 enc[8 + 4][32 + 4];
  __shared__ uint2 colcnt[8][32 + 4];
  const int x0 = blockIdx.x * 32, y0 = blockIdx.y * 8;
  const u8* s = src + (size_t)blockIdx.z * src_stride;
  const int tid = threadIdx.y * 32 + threadIdx.x;
  for (int idx = tid; idx < 12 * 36; idx += 256) {
    int ty = idx / 36, tx = idx - ty * 36;
    u32 v = s[(size_t)clampi(y0 - 2 + ty, 0, rows - 1) * cols + clampi(x0 - 2 + tx, 0, cols - 1)];
    enc[ty][tx] = make_uint2(((v & 15u) * 0x00204081u) & 0x01010101u, ((v >> 4) * 0x00204081u) & 0x01010101u);
  }
  __syncthreads();
  for (int idx = tid; idx < 8 * 36; idx += 256) {
    int ty = idx / 36, tx = idx - ty * 36;
    uint2 a = enc[ty][tx], b = enc[ty + 1][tx], c = enc[ty + 2][tx], d = enc[ty + 3][tx], e = enc[ty + 4][tx];
    colcnt[ty][tx] = make_uint2(a.x + b.x + c.x + d.x + e.x, a.y + b.y + c.y + d.y + e.y);
  }
  __syncthreads();
  const int x = x0 + threadIdx.x, y = y0 + threadIdx.y;
  if (x >= cols || y >= rows) return;
  u32 lo = 0, hi = 0;
#pragma unroll
  for (int j = 0; j < 5; ++j) {
    uint2 c = colcnt[threadIdx.y][threadIdx.x + j];
    lo += c.x; hi += c.y;
  }
  const u32 plo = lo * 0x01010101u;                  // running sums of labels 0..3
  const u32 tlo = plo >> 24;                         // count of labels 0..3
  const u32 phi = hi * 0x01010101u;                  // running sums of labels 4..7 (without the low half)
  const u32 zeros = 25u - tlo - (phi >> 24);
  const u32 bias = (zeros + 115u) * 0x01010101u;     // + zeros, + (128 - 13)
  const u32 ge_lo = (plo + bias) & 0x80808080u;
  const u32 ge_hi = (phi + tlo * 0x01010101u + bias) & 0x80808080u;
  const int nge = __popc(ge_lo) + __popc(ge_hi);     // labels whose running count has reached 13 (>= 1 when zeros < 13)
  dst[(size_t)blockIdx.z * dst_stride + (size_t)y * cols + x] = zeros >= 13u ? (u8)0 : (u8)(1u << (8 - nge));
}

void launch_median5(const u8* src, size_t src_stride, u8* dst, size_t dst_stride, int rows, int cols, int frames,
                    cudaStream_t st) {
  dim3 grid((cols + 31) / 32, (rows + 7) / 8, frames), block(32, 8);
  median5_kernel<<<grid, block, 0, st>>>(src, src_stride, dst, dst_stride, rows, cols);
}

// ---------------------------------------------------------------------------------------------
// cv::resize(INTER_NEAREST): sx = min(floor(x * (1/(dcols/cols))), cols-1)  (double, like OpenCV)
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) resize_nn_kernel(const u8* __restrict__ src, size_t src_stride, int rows, int cols,
                                                        u8* __restrict__ dst, size_t dst_stride, int drows, int dcols,
                                                        double ifx, double ify) {
  const int x = blockIdx.x * 32 + threadIdx.x, y = blockIdx.y * 8 + threadIdx.y;
  if (x >= dcols || y >= drows) return;
  int sx = min((int)floor(x * ifx), cols - 1), sy = min((int)floor(y * ify), rows - 1);
  dst[(size_t)blockIdx.z * dst_stride + (size_t)y * dcols + x] = src[(size_t)blockIdx.z * src_stride + (size_t)sy * cols + sx];
}

void launch_resize_nn(const u8* src, size_t src_stride, int rows, int cols, u8* dst, size_t dst_stride, int drows,
                      int dcols, int frames, cudaStream_t st) {
  dim3 grid((dcols + 31) / 32, (drows + 7) / 8, frames), block(32, 8);
  double ifx = 1.0 / ((double)dcols / cols), ify = 1.0 / ((double)drows / rows);
  resize_nn_kernel<<<grid, block, 0, st>>>(src, src_stride, rows, cols, dst, dst_stride, drows, dcols, ifx, ify);
}

// ---------------------------------------------------------------------------------------------
// spread + response maps + linearize, fused.  One CTA = one decimated row i (image rows
// [iT, iT+2T-1)) x SL_CW decimated columns.  The band is staged in shared memory with its
// (T-1) right/bottom halo (zero outside the image: OR identity), OR-reduced vertically then
// horizontally, looked up in a 256 x 8-byte table (all 8 orientations of one spread byte at once)
// and written straight into the T-strided linear memories (LevelGeom: flat rows on the coarsest level,
//     LM[ori][ (y%T*T + x%T) * W*H + (y/T)*W + x/T ],
// 16-column strips LM[ori][cell][x/T/16][y/T][16] on the finer ones).
// Each thread produces 4 consecutive decimated positions of one grid cell for all 8 orientations,
// i.e. eight 32-bit stores; a warp writes 64 B runs (flat) or 16 B runs (strips) per (ori, grid cell).
// ---------------------------------------------------------------------------------------------
constexpr int SL_CW = 64;

__device__ __forceinline__ u32 bytes_nonzero_mask(u32 m) {  // 0xFF in every byte of m that is non-zero
  u32 nz = (((m & 0x7F7F7F7Fu) + 0x7F7F7F7Fu) | m) & 0x80808080u;
  return (nz >> 7) * 0xFFu;
}

// Shared memory holds the band as 32-bit words (4 pixels): the T x T OR is T word-ORs vertically and
// T funnel-shift+ORs horizontally per 4 pixels.
__global__ void __launch_bounds__(256) spread_linearize_kernel(const u8* __restrict__ q, size_t q_stride,
                                                               const u8* __restrict__ mask, size_t mask_stride,
                                                               u8* __restrict__ lm, size_t lm_stride, LevelGeom g,
                                                               const uint2* __restrict__ table) {
  extern __shared__ __align__(16) u8 sl_smem[];
  const int T = g.T, W = g.W;
  const int i = blockIdx.x;            // decimated row
  const int c0 = blockIdx.y * SL_CW;   // first decimated column of this chunk
  const int cw = min(SL_CW, W - c0);
  const int pw = cw * T;               // pixel columns produced
  const int pww = (pw + 3) >> 2;       // ... in words
  const int wpr = pww + 5;             // words per staged row: T-1 halo bytes + slack for the 5-word window
  const int nr = 2 * T - 1;
  u32* qb = reinterpret_cast<u32*>(sl_smem);                 // [nr][wpr] raw band; later the spread band [T][wpr]
  u32* vb = qb + nr * wpr;                                   // [T][wpr]  vertical OR
  uint2* tab = reinterpret_cast<uint2*>(sl_smem + (((size_t)(nr + T) * wpr * 4 + 15) & ~(size_t)15));
  const int tid = threadIdx.x;
  const u8* qf = q + (size_t)blockIdx.z * q_stride;
  const u8* mf = mask ? mask + (size_t)blockIdx.z * mask_stride : nullptr;

  tab[tid] = table[tid];  // 256 threads, 256 entries
  const int y0 = i * T, x0 = c0 * T;
  const bool vec_in = ((g.cols & 3) == 0) && ((x0 & 3) == 0);
  for (int idx = tid; idx < nr * wpr; idx += 256) {
    const int r = idx / wpr, xw = idx - r * wpr;
    const int gy = y0 + r, gx = x0 + 4 * xw;
    u32 v = 0;
    if (gy < g.rows && gx < g.cols) {
      const size_t o = (size_t)gy * g.cols + gx;
      if (vec_in && gx + 3 < g.cols) {
        v = *reinterpret_cast<const u32*>(qf + o);
        if (mf) v &= bytes_nonzero_mask(*reinterpret_cast<const u32*>(mf + o));
      } else {
        for (int b = 0; b < 4 && gx + b < g.cols; ++b) {
          u32 px = qf[o + b];
          if (mf && !mf[o + b]) px = 0;
          v |= px << (8 * b);
        }
      }
    }
    qb[idx] = v;
  }
  __syncthreads();
  for (int idx = tid; idx < T * wpr; idx += 256) {
    const int r = idx / wpr, xw = idx - r * wpr;
    u32 v = 0;
    for (int k = 0; k < T; ++k) v |= qb[(r + k) * wpr + xw];
    vb[idx] = v;
  }
  __syncthreads();
  for (int idx = tid; idx < T * pww; idx += 256) {
    const int r = idx / pww, xw = idx - r * pww;
    const u32* wp = vb + r * wpr + xw;
    const u32 w0 = wp[0], w1 = wp[1], w2 = wp[2], w3 = wp[3], w4 = wp[4];
    const u32 ww[6] = {w0, w1, w2, w3, w4, 0u};
    u32 v = 0;
#pragma unroll
    for (int k = 0; k < 16; ++k)
      if (k < T) v |= __funnelshift_r(ww[k >> 2], ww[(k >> 2) + 1], (k & 3) * 8);
    qb[r * wpr + xw] = v;  // qb rows [0,T) now hold the spread band
  }
  __syncthreads();

  const u8* sb = reinterpret_cast<const u8*>(qb);
  const int rowb = wpr * 4;
  u8* lmf = lm + (size_t)blockIdx.z * lm_stride;
  const u32 per = g.per_label;
  const u32 plane = g.plane;
  const u32 H16 = (u32)g.H * 16u;
  const int quads = (cw + 3) >> 2;
  const bool vec = ((W & 3) == 0);
  for (int idx = tid; idx < T * T * quads; idx += 256) {
    int cell = idx / quads, k = idx - cell * quads;
    int gy = cell / T, gx = cell - gy * T;
    u32 lo[4], hi[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      int p = 4 * k + j;
      uint2 e = make_uint2(0u, 0u);
      if (p < cw) e = tab[sb[gy * rowb + p * T + gx]];
      lo[j] = e.x; hi[j] = e.y;
    }
    const u32 col = (u32)(c0 + 4 * k);  // 4 consecutive columns never straddle a 16-column strip
    const u32 dstoff = (u32)cell * plane + (g.strips ? (col >> 4) * H16 + (u32)i * 16u + (col & 15u) : (u32)i * W + col);
    if (vec && 4 * k + 3 < cw) {
      // transpose 4 positions x 8 orientations -> one u32 (4 positions) per orientation
      u32 t01 = __byte_perm(lo[0], lo[1], 0x5140), t23 = __byte_perm(lo[2], lo[3], 0x5140);
      u32 u01 = __byte_perm(lo[0], lo[1], 0x7362), u23 = __byte_perm(lo[2], lo[3], 0x7362);
      *(u32*)(lmf + 0 * (size_t)per + dstoff) = __byte_perm(t01, t23, 0x5410);
      *(u32*)(lmf + 1 * (size_t)per + dstoff) = __byte_perm(t01, t23, 0x7632);
      *(u32*)(lmf + 2 * (size_t)per + dstoff) = __byte_perm(u01, u23, 0x5410);
      *(u32*)(lmf + 3 * (size_t)per + dstoff) = __byte_perm(u01, u23, 0x7632);
      t01 = __byte_perm(hi[0], hi[1], 0x5140); t23 = __byte_perm(hi[2], hi[3], 0x5140);
      u01 = __byte_perm(hi[0], hi[1], 0x7362); u23 = __byte_perm(hi[2], hi[3], 0x7362);
      *(u32*)(lmf + 4 * (size_t)per + dstoff) = __byte_perm(t01, t23, 0x5410);
      *(u32*)(lmf + 5 * (size_t)per + dstoff) = __byte_perm(t01, t23, 0x7632);
      *(u32*)(lmf + 6 * (size_t)per + dstoff) = __byte_perm(u01, u23, 0x5410);
      *(u32*)(lmf + 7 * (size_t)per + dstoff) = __byte_perm(u01, u23, 0x7632);
    } else {
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        if (4 * k + j < cw) {
#pragma unroll
          for (int o = 0; o < 4; ++o) {
            lmf[(size_t)o * per + dstoff + j] = (u8)(lo[j] >> (8 * o));
            lmf[(size_t)(o + 4) * per + dstoff + j] = (u8)(hi[j] >> (8 * o));
          }
        }
      }
    }
  }
}

static size_t sl_smem_bytes(int T, int cw) {
  int pww = (cw * T + 3) >> 2, wpr = pww + 5;
  size_t b = ((size_t)(3 * T - 1) * wpr * 4 + 15) & ~(size_t)15;
  return b + 256 * sizeof(uint2);
}

void launch_spread_linearize(const u8* q, size_t q_stride, const u8* mask, size_t mask_stride, u8* lm,
                             size_t lm_stride, LevelGeom g, const uint2* table, int frames, cudaStream_t st) {
  size_t smem = sl_smem_bytes(g.T, g.W < SL_CW ? g.W : SL_CW);
  if (smem > 48 * 1024)  // per device, so set it on every such launch (cheap) rather than caching it per process
    cudaFuncSetAttribute(spread_linearize_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  dim3 grid(g.H, (g.W + SL_CW - 1) / SL_CW, frames);
  spread_linearize_kernel<<<grid, 256, smem, st>>>(q, q_stride, mask, mask_stride, lm, lm_stride, g, table);
}

}  // namespace lmk
