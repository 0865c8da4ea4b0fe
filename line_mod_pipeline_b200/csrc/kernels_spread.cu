// kernels_spread.cu — spread + computeResponseMaps + linearize, fused, round-2 kernels (sm_100a).
// Upstream: spread() / computeResponseMaps() / linearize() of opencv_contrib rgbd/linemod.cpp (SURVEY.md §8a a8-a10,
// Appendix A.5).  HBM-bound byte kernels (1 B in, 8 B out per pixel): no tensor cores.
//
// Round 1's kernel (kept in kernels_frame.cu as the generic fallback) divided by the runtime T in every index
// computation and wrote the fine levels in 16 B runs.  Here T is a template parameter, the band is OR-reduced on
// 32-bit words, and the output tile is chosen so that one warp instruction stores 512 contiguous bytes:
//   spread_strip_kernel<T,R>  finer levels (16-column strip layout read by similarityLocal): tile = one strip x R
//                             decimated rows; a thread owns one (grid cell, decimated row) = 16 positions and stores
//                             8 x 16 B (one per orientation); the lanes of a warp are consecutive rows of one cell.
//   spread_flat_kernel<T>     coarsest level: writes the NIBBLE-PACKED flat linear memory the coarse kernel reads
//                             directly (no byte copy, no pack_nibbles pass) and the per-frame response sum that orders
//                             the modalities; a thread owns 8 consecutive positions of one cell = 4 B per orientation.
// DepthNormal coarser levels are nearest-neighbour decimations of the level-0 map (DepthNormalPyramid::pyrDown):
// q_step > 1 reads q[y*step][x*step] of the level-0 map in place, so no decimated copy is ever written.
#include "kernels.cuh"

namespace lmk {

__device__ __forceinline__ u32 nz_mask(u32 m) {  // 0xFF in every byte of m that is non-zero
  u32 nz = (((m & 0x7F7F7F7Fu) + 0x7F7F7F7Fu) | m) & 0x80808080u;
  return (nz >> 7) * 0xFFu;
}

// One word (4 pixels) of the quantized band at image position (gy, gx..gx+3); zero outside the image (OR identity).
__device__ __forceinline__ u32 load_q_word(const SpreadArgs& a, const u8* __restrict__ qf, const u8* __restrict__ mf, int gy, int gx,
                                           bool vec) {
  const int rows = a.g.rows, cols = a.g.cols;
  if (gy >= rows || gx >= cols) return 0u;
  u32 v = 0;
  if (a.q_step == 1) {
    const size_t o = (size_t)gy * a.q_pitch + gx;
    if (vec && gx + 3 < cols) {
      v = __ldg(reinterpret_cast<const u32*>(qf + o));
    } else {
      for (int b = 0; b < 4 && gx + b < cols; ++b) v |= (u32)__ldg(qf + o + b) << (8 * b);
    }
  } else {  // nearest-neighbour decimation of the finer map, read in place
    const size_t o = (size_t)gy * a.q_step * a.q_pitch + (size_t)gx * a.q_step;
    for (int b = 0; b < 4 && gx + b < cols; ++b) v |= (u32)__ldg(qf + o + (size_t)b * a.q_step) << (8 * b);
  }
  if (mf) {  // quantize() copies through the mask
    const size_t mo = (size_t)gy * cols + gx;
    if (vec && gx + 3 < cols) v &= nz_mask(__ldg(reinterpret_cast<const u32*>(mf + mo)));
    else
      for (int b = 0; b < 4 && gx + b < cols; ++b)
        if (!__ldg(mf + mo + b)) v &= ~(0xFFu << (8 * b));
  }
  return v;
}

// 4x4 byte transpose: rows r0..r3 (one word each) -> columns c0..c3
__device__ __forceinline__ void transpose4(u32 r0, u32 r1, u32 r2, u32 r3, u32& c0, u32& c1, u32& c2, u32& c3) {
  const u32 t01 = __byte_perm(r0, r1, 0x5140), t23 = __byte_perm(r2, r3, 0x5140);  // bytes 0,1 of each pair interleaved
  const u32 u01 = __byte_perm(r0, r1, 0x7362), u23 = __byte_perm(r2, r3, 0x7362);  // bytes 2,3
  c0 = __byte_perm(t01, t23, 0x5410); c1 = __byte_perm(t01, t23, 0x7632);
  c2 = __byte_perm(u01, u23, 0x5410); c3 = __byte_perm(u01, u23, 0x7632);
}

// Staged band -> spread band.  band[NV][RW] (RW words per row, right halo of T-1 pixels included) receives
//   spread(y, x) = OR of q over [y, y+T) x [x, x+T)      (upstream spread(): dst(y, x) |= src(y + r, x + c), clipped)
// for the NV rows from image row y0 and the PW*4 pixel columns from x0.
// Phase A — a thread owns (word column, segment of SEG rows): it reads its SEG + T-1 words straight from global memory
//   (independent loads, all in flight together; round 1's row-at-a-time staging spent 41 % of the kernel's instructions
//   and most of its stalls here), keeps the last T-1 rows in registers and stores the vertical OR.
// Phase B — horizontal OR on words (funnel shifts), results held in registers across the barrier and written in place.
// FAST = 1: own-level map, 4-byte aligned rows, no mask.  FAST = 2: the 2x nearest-neighbour decimation of a finer map read
// in place (DepthNormal level 1: pixels gx..gx+3 are the even bytes of ONE aligned 8-byte load of the finer row 2*gy).
// FAST = 0: load_q_word (other decimations, masks, ragged edges).  All threads of the CTA call it; ends with a barrier.
template <int T, int SEG, int MAXV, int FAST>
__device__ __forceinline__ void stage_spread(const SpreadArgs& a, const u8* __restrict__ qf, const u8* __restrict__ mf, bool vec,
                                             int y0, int x0, int NV, int RW, int PW, u32* band, int tid, int nthr) {
  const int rows = a.g.rows, cols = a.g.cols;
  const int nseg = (NV + SEG - 1) / SEG;
  for (int it = tid; it < RW * nseg; it += nthr) {
    const int sg = it / RW, w = it - sg * RW;
    const int r0 = sg * SEG, gx = x0 + 4 * w;
    const bool colok = gx < cols;                       // FAST: cols % 4 == 0 and gx % 4 == 0, so the whole word is inside
    const u8* p = qf + (size_t)(y0 + r0) * a.q_pitch * (FAST == 2 ? 2 : 1) + (size_t)gx * (FAST == 2 ? 2 : 1);
    u32 in[SEG + T - 1];
#pragma unroll
    for (int k = 0; k < SEG + T - 1; ++k) {
      const int gy = y0 + r0 + k;
      if (FAST == 1) in[k] = (colok && gy < rows) ? __ldg(reinterpret_cast<const u32*>(p + (size_t)k * a.q_pitch)) : 0u;
      else if (FAST == 2) {
        uint2 v = make_uint2(0u, 0u);
        if (colok && gy < rows) v = __ldg(reinterpret_cast<const uint2*>(p + (size_t)k * 2 * a.q_pitch));
        in[k] = __byte_perm(v.x, v.y, 0x6420);
      } else in[k] = load_q_word(a, qf, mf, gy, gx, vec);
    }
#pragma unroll
    for (int j = 0; j < SEG; ++j) {
      u32 v = in[j];
#pragma unroll
      for (int k = 1; k < T; ++k) v |= in[j + k];
      if (r0 + j < NV) band[(r0 + j) * RW + w] = v;
    }
  }
  __syncthreads();
  constexpr int NW = (T + 2) / 4 + 1;  // words a T-wide window starting in word w can touch: bytes 4w .. 4w+3+T-1
  const int total = NV * PW;
  u32 hv[MAXV];
#pragma unroll
  for (int q = 0; q < MAXV; ++q) {
    const int it = tid + q * nthr;
    hv[q] = 0u;
    if (it < total) {
      const int r = it / PW, w = it - r * PW;
      const u32* p = band + r * RW + w;
      u32 ww[NW + 1];
#pragma unroll
      for (int j = 0; j < NW; ++j) ww[j] = (w + j < RW) ? p[j] : 0u;
      ww[NW] = 0u;
      u32 v = 0;
#pragma unroll
      for (int k = 0; k < T; ++k) v |= __funnelshift_r(ww[k >> 2], ww[(k >> 2) + 1], (k & 3) * 8);
      hv[q] = v;
    }
  }
  __syncthreads();
#pragma unroll
  for (int q = 0; q < MAXV; ++q) {
    const int it = tid + q * nthr;
    if (it < total) {
      const int r = it / PW, w = it - r * PW;
      band[r * RW + w] = hv[q];
    }
  }
  __syncthreads();
}

// ------------------------------------------------------------------------------------------------ strip layout
template <int T, int R>
struct StripCfg {
  static constexpr int PXW = 16 * T;                    // pixel columns of one strip
  static constexpr int PW = PXW / 4;                    // ... in words
  static constexpr int RW = ((PXW + T - 1 + 3) / 4) | 1;  // staged words per row (right halo T-1), odd: lanes = rows hit distinct banks for odd T
  static constexpr int NV = R * T;                      // spread rows
  static constexpr int SEG = T >= 16 ? 8 : 16;          // rows per thread in phase A (registers: SEG + T - 1 words)
  static constexpr int MAXV = (NV * PW + 255) / 256;    // phase-B words per thread
  static constexpr size_t SMEM = (size_t)NV * RW * 4 + 256 * sizeof(uint2);
};

template <int T, int R, int FAST>
__global__ void __launch_bounds__(256, T >= 16 ? 2 : 5) spread_strip_kernel(SpreadArgs a) {
  typedef StripCfg<T, R> C;
  extern __shared__ __align__(16) u8 sp_smem[];
  uint2* tab = reinterpret_cast<uint2*>(sp_smem);
  u32* band = reinterpret_cast<u32*>(sp_smem + 256 * sizeof(uint2));
  const int tid = threadIdx.x;
  const int strip = blockIdx.x, rg = blockIdx.y, frame = blockIdx.z;
  const u8* qf = a.q + (size_t)frame * a.q_stride;
  const u8* mf = a.mask ? a.mask + (size_t)frame * a.mask_stride : nullptr;
  tab[tid] = __ldg(a.table + tid);
  const int y0 = rg * C::NV, x0 = strip * C::PXW;
  const bool vec = a.q_step == 1 && ((a.q_pitch & 3) == 0) && ((reinterpret_cast<size_t>(qf) & 3) == 0) &&
                   (!mf || (((a.g.cols & 3) == 0) && ((reinterpret_cast<size_t>(mf) & 3) == 0)));
  stage_spread<T, C::SEG, C::MAXV, FAST>(a, qf, mf, vec, y0, x0, C::NV, C::RW, C::PW, band, tid, 256);

  const u8* sb = reinterpret_cast<const u8*>(band);
  u8* lmf = a.lm + (size_t)frame * a.lm_stride;
  const int H = a.g.H, W = a.g.W;
  const u32 per = a.g.per_label, plane = a.g.plane;
  const int ncols = min(16, W - strip * 16);  // valid decimated columns of this strip
  for (int it = tid; it < T * T * R; it += 256) {
    const int cell = it / R, i = it - cell * R;  // R is a power of two
    const int ig = rg * R + i;
    if (ig >= H) continue;
    const int gy = cell / T, gx = cell - gy * T;
    const u8* row = sb + (size_t)(i * T + gy) * (C::RW * 4) + gx;
    uint4 o[8];
    u32 sp[16];                                    // all 16 spread bytes first: the table lookups then overlap instead of chaining
#pragma unroll
    for (int k = 0; k < 16; ++k) sp[k] = row[k * T];
#pragma unroll
    for (int g4 = 0; g4 < 4; ++g4) {
      u32 lo[4], hi[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int k = 4 * g4 + j;
        uint2 e = tab[sp[k]];
        if (k >= ncols) e = make_uint2(0u, 0u);
        lo[j] = e.x; hi[j] = e.y;
      }
      u32 c0, c1, c2, c3;
      transpose4(lo[0], lo[1], lo[2], lo[3], c0, c1, c2, c3);
      (&o[0].x)[g4] = c0; (&o[1].x)[g4] = c1; (&o[2].x)[g4] = c2; (&o[3].x)[g4] = c3;
      transpose4(hi[0], hi[1], hi[2], hi[3], c0, c1, c2, c3);
      (&o[4].x)[g4] = c0; (&o[5].x)[g4] = c1; (&o[6].x)[g4] = c2; (&o[7].x)[g4] = c3;
    }
    u8* dst = lmf + (size_t)cell * plane + (size_t)strip * ((size_t)H * 16) + (size_t)ig * 16;
#pragma unroll
    for (int ori = 0; ori < 8; ++ori) *reinterpret_cast<uint4*>(dst + (size_t)ori * per) = o[ori];
  }
}

// ------------------------------------------------------------------------------------------------ flat nibble layout
// Tile = R decimated rows x CW decimated columns (CW a multiple of 8; the launcher makes W % 8 == 0 a precondition).
// Small tiles on purpose: the coarsest level is 1/4 of the pixels and 1/16 of the output bytes of the level below, so
// only many short CTAs fill 148 SMs (5 tiles per frame left the kernel at 0.36 waves and 25 % of the warp slots).
constexpr int FLAT_MAXV = 8;       // phase-B words per thread the launcher's tiles need at most
template <int T, int FAST>
__global__ void __launch_bounds__(256) spread_flat_kernel(SpreadArgs a, int R, int CW) {
  extern __shared__ __align__(16) u8 sp_smem[];
  uint2* tab = reinterpret_cast<uint2*>(sp_smem);
  const int tid = threadIdx.x;
  const int H = a.g.H, W = a.g.W;
  const int c0 = blockIdx.x * CW, cw = min(CW, W - c0);
  const int rg = blockIdx.y, frame = blockIdx.z;
  const int nrows = min(R, H - rg * R);            // decimated rows of this tile
  const int PW = cw * T / 4;                       // cw % 8 == 0 -> whole words
  const int RW = ((cw * T + T - 1 + 3) / 4) | 1;
  const int NV = nrows * T;
  u32* band = reinterpret_cast<u32*>(sp_smem + 256 * sizeof(uint2));
  const u8* qf = a.q + (size_t)frame * a.q_stride;
  const u8* mf = a.mask ? a.mask + (size_t)frame * a.mask_stride : nullptr;
  tab[tid] = __ldg(a.table + tid);
  const int y0 = rg * R * T, x0 = c0 * T;
  const bool vec = a.q_step == 1 && ((a.q_pitch & 3) == 0) && ((reinterpret_cast<size_t>(qf) & 3) == 0) &&
                   (!mf || (((a.g.cols & 3) == 0) && ((reinterpret_cast<size_t>(mf) & 3) == 0)));
  stage_spread<T, 4, FLAT_MAXV, FAST>(a, qf, mf, vec, y0, x0, NV, RW, PW, band, tid, 256);

  const u8* sb = reinterpret_cast<const u8*>(band);
  u8* out = a.lmn + (size_t)frame * a.lmn_stride;
  const u32 HW = (u32)H * W;
  const u32 per_n = (u32)(T * T) * HW / 2;         // nibble-packed bytes per label
  const int cq = cw >> 3;                          // 8-position chunks per tile row
  const int per_cell = nrows * cq;
  u32 rsum = 0;
  for (int it = tid; it < T * T * per_cell; it += 256) {
    const int cell = it / per_cell, q = it - cell * per_cell;
    const int i = q / cq, cc = q - i * cq;
    const int gy = cell / T, gx = cell - gy * T;
    const u8* row = sb + (size_t)(i * T + gy) * (RW * 4) + (size_t)(cc * 8) * T + gx;
    u32 lo2[4], hi2[4];                             // byte o of lo2[j]: responses of positions 2j, 2j+1 for orientation o, nibble-packed
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const uint2 e0 = tab[row[(2 * j) * T]], e1 = tab[row[(2 * j + 1) * T]];
      lo2[j] = e0.x + (e1.x << 4);                  // responses <= 4: the nibbles never carry
      hi2[j] = e0.y + (e1.y << 4);
      if (tid < 32) rsum += __dp4a(e0.x, 0x01010101u, __dp4a(e0.y, 0x01010101u, __dp4a(e1.x, 0x01010101u, __dp4a(e1.y, 0x01010101u, 0u))));
    }
    u32 w0, w1, w2, w3, w4, w5, w6, w7;
    transpose4(lo2[0], lo2[1], lo2[2], lo2[3], w0, w1, w2, w3);
    transpose4(hi2[0], hi2[1], hi2[2], hi2[3], w4, w5, w6, w7);
    const u32 pos = (u32)(rg * R + i) * W + (u32)(c0 + cc * 8);   // multiple of 8
    u8* dst = out + (size_t)cell * (HW / 2) + (pos >> 1);
    *reinterpret_cast<u32*>(dst + 0 * (size_t)per_n) = w0; *reinterpret_cast<u32*>(dst + 1 * (size_t)per_n) = w1;
    *reinterpret_cast<u32*>(dst + 2 * (size_t)per_n) = w2; *reinterpret_cast<u32*>(dst + 3 * (size_t)per_n) = w3;
    *reinterpret_cast<u32*>(dst + 4 * (size_t)per_n) = w4; *reinterpret_cast<u32*>(dst + 5 * (size_t)per_n) = w5;
    *reinterpret_cast<u32*>(dst + 6 * (size_t)per_n) = w6; *reinterpret_cast<u32*>(dst + 7 * (size_t)per_n) = w7;
  }
  if (tid < 32 && a.resp_sum) {  // 1/8 sample of the tile's responses: the statistic that orders the modalities
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) rsum += __shfl_xor_sync(0xffffffffu, rsum, d);
    if (tid == 0 && rsum) atomicAdd(a.resp_sum + (size_t)frame * a.resp_stride, rsum);
  }
}

// FAST paths of stage_spread (0: none).  1: the map is this level's own, unmasked, every row a whole number of aligned
// words.  2: the map is the finer level's, decimated by two in place: rows 2*gy, aligned 8-byte loads.
static int fast_loads(const SpreadArgs& a) {
  if (a.mask || (a.g.cols & 3) != 0) return 0;
  if (a.q_step == 1 && (a.q_pitch & 3) == 0 && (a.q_stride & 3) == 0 && (reinterpret_cast<size_t>(a.q) & 3) == 0) return 1;
  if (a.q_step == 2 && (a.q_pitch & 7) == 0 && (a.q_stride & 7) == 0 && (reinterpret_cast<size_t>(a.q) & 7) == 0 &&
      a.q_pitch >= 2 * a.g.cols)
    return 2;
  return 0;
}

template <int T, int R>
void launch_strip(const SpreadArgs& a, int frames, cudaStream_t st) {
  typedef StripCfg<T, R> C;
  static_assert(C::SMEM <= 48 * 1024, "strip tile exceeds the default dynamic shared memory");
  dim3 grid(a.g.strips, (a.g.H + R - 1) / R, frames);
  switch (fast_loads(a)) {
    case 1: spread_strip_kernel<T, R, 1><<<grid, 256, C::SMEM, st>>>(a); break;
    case 2: spread_strip_kernel<T, R, 2><<<grid, 256, C::SMEM, st>>>(a); break;
    default: spread_strip_kernel<T, R, 0><<<grid, 256, C::SMEM, st>>>(a);
  }
}

template <int T>
void launch_flat(const SpreadArgs& a, int frames, cudaStream_t st) {
  const int W = a.g.W, H = a.g.H;
  int CW = W <= 64 ? W : 64;                        // W % 8 == 0, so every chunk is a multiple of 8 too
  int R = 16 / T; if (R < 1) R = 1; if (R > H) R = H;   // ~16 image rows per tile
  while (R > 1 && (R * T) * (CW * T / 4) > 256 * FLAT_MAXV) --R;
  while (CW > 8 && (R * T) * (CW * T / 4) > 256 * FLAT_MAXV) CW -= 8;
  const int RWmax = ((CW * T + T - 1 + 3) / 4) | 1;
  const size_t smem = (size_t)(R * T) * RWmax * 4 + 256 * sizeof(uint2);
  dim3 grid((W + CW - 1) / CW, (H + R - 1) / R, frames);
  const int fast = fast_loads(a);
  if (fast == 1) {
    if (smem > 48 * 1024) cudaFuncSetAttribute(spread_flat_kernel<T, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    spread_flat_kernel<T, 1><<<grid, 256, smem, st>>>(a, R, CW);
  } else if (fast == 2) {
    if (smem > 48 * 1024) cudaFuncSetAttribute(spread_flat_kernel<T, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    spread_flat_kernel<T, 2><<<grid, 256, smem, st>>>(a, R, CW);
  } else {
    if (smem > 48 * 1024) cudaFuncSetAttribute(spread_flat_kernel<T, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    spread_flat_kernel<T, 0><<<grid, 256, smem, st>>>(a, R, CW);
  }
}

// Returns false when the geometry is not covered (the caller then runs round 1's generic kernels).
bool launch_spread_fast(const SpreadArgs& a, int frames, cudaStream_t st) {
  const int T = a.g.T;
  if (a.g.strips) {  // finer level, strip layout
    switch (T) {
      case 2: launch_strip<2, 64>(a, frames, st); return true;
      case 4: launch_strip<4, 32>(a, frames, st); return true;
      case 5: launch_strip<5, 32>(a, frames, st); return true;
      case 8: launch_strip<8, 16>(a, frames, st); return true;
      case 16: launch_strip<16, 8>(a, frames, st); return true;
      default: return false;
    }
  }
  // coarsest level: nibble-packed flat layout; chunks of 8 positions must not straddle rows
  if (!a.lmn || (a.g.W & 7) != 0) return false;
  switch (T) {
    case 2: launch_flat<2>(a, frames, st); return true;
    case 4: launch_flat<4>(a, frames, st); return true;
    case 5: launch_flat<5>(a, frames, st); return true;
    case 8: launch_flat<8>(a, frames, st); return true;
    case 16: launch_flat<16>(a, frames, st); return true;
    default: return false;
  }
}

}  // namespace lmk
