// kernels_match.cu — template-side kernels (sm_100a): the hot loops of Detector::matchClass
// (opencv_contrib rgbd/linemod.cpp):
//   build_offsets_kernel        accessLinearMemory() for every feature, once per frame geometry  (a11)
//   similarity_coarse_kernel    similarity + addSimilarities + threshold scan, fused            (a12-a14)
//   similarity_local_kernel     similarityLocal + addSimilarities + argmax + update + filter    (a15,a16)
//   pack_kernel                 ordered compaction into reference generation order              (a17)
// These are sparse u8 gather-accumulates over L2-resident linear memories: no tensor cores.
// Packed byte sums never carry between bytes (<= 63 features x 4 per modality = 252), so one 32-bit
// IADD adds four responses; unaligned feature rows are read as aligned 128-bit (coarse) / 64-bit
// (local) loads and realigned with funnel shifts.
#include "kernels.cuh"

namespace lmk {

// ---------------------------------------------------------------------------------------------
// Plan kernel: one warp per template.  offs[g][m*64+k] = label*per_label + grid_index*W*H + lm_index
// (OFF_INVALID when upstream's similarity() would skip the feature).  hdr.flags:
//   bit0  local-safe : similarityLocal can never skip a feature or leave its plane (see below)
//   bit1  coarse-safe: every feature row + template_positions stays inside its label's block
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128) build_offsets_kernel(const u32* __restrict__ feat, u32* __restrict__ offs,
                                                            TplHdr* __restrict__ hdr, int ntpl, int M, LevelGeom g) {
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (warp >= ntpl) return;
  TplHdr h = hdr[warp];
  const int T = g.T, W = g.W, H = g.H;
  const u32 plane = (u32)W * H;
  bool local_safe = h.width[0] >= 0 && h.height[0] >= 0 && h.width[0] <= g.cols - 16 * T && h.height[0] <= g.rows - 16 * T;
  bool coarse_safe = true;
  for (int s = lane; s < M * FEAT_SLOTS; s += 32) {
    int m = s / FEAT_SLOTS, k = s - m * FEAT_SLOTS;
    u32 off = OFF_INVALID;
    if (k < h.nf[m]) {
      u32 f = feat[(size_t)warp * M * FEAT_SLOTS + s];
      int x = f & 0x3FFF, y = (f >> 14) & 0x3FFF, label = (f >> 28) & 7;
      bool valid = (f >> 31) && x < g.cols && y < g.rows;
      if (valid) {
        u32 base = (u32)((y % T) * T + (x % T)) * plane + (u32)(y / T) * W + (u32)(x / T);
        off = (u32)label * g.per_label + base;
        int wf = (h.width[m] - 1) / T + 1, hf = (h.height[m] - 1) / T + 1;
        long long P = (long long)(H - hf) * W + (W - wf) + 1;
        if (P > (long long)plane) P = plane;
        if (P > 0 && (long long)base + P > (long long)g.per_label) coarse_safe = false;
      }
      if (!(f >> 31) || x > h.width[0] || y > h.height[0]) local_safe = false;
    }
    offs[(size_t)warp * M * FEAT_SLOTS + s] = off;
  }
  local_safe = __all_sync(0xffffffffu, local_safe);
  coarse_safe = __all_sync(0xffffffffu, coarse_safe);
  if (lane == 0) hdr[warp].flags = (local_safe ? 1u : 0u) | (coarse_safe ? 2u : 0u);
}

void launch_build_offsets(const u32* feat, u32* offs, TplHdr* hdr, int ntpl, int M, LevelGeom g, cudaStream_t st) {
  if (ntpl <= 0) return;
  int blocks = (ntpl * 32 + 127) / 128;
  build_offsets_kernel<<<blocks, 128, 0, st>>>(feat, offs, hdr, ntpl, M, g);
}

// ---------------------------------------------------------------------------------------------
// helpers
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ u32 ld_nc_u32(const u32* p) { return __ldg(p); }

// zero the bytes at index >= nv (nv in [0,16]) of a 16-byte value held in 4 little-endian words
__device__ __forceinline__ void mask16(u32& w0, u32& w1, u32& w2, u32& w3, int nv) {
  u32 w[4] = {w0, w1, w2, w3};
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    int r = nv - 4 * i;
    u32 m = r >= 4 ? 0xFFFFFFFFu : (r <= 0 ? 0u : (0xFFFFFFFFu >> (8 * (4 - r))));
    w[i] &= m;
  }
  w0 = w[0]; w1 = w[1]; w2 = w[2]; w3 = w[3];
}

__device__ __forceinline__ int raw_threshold(int nf, float threshold) {
  // (int)(2*nf + (threshold/100.f)*(2*nf) + 0.5f), float arithmetic without contraction
  float two_nf = (float)(2 * nf);
  return (int)__fadd_rn(__fadd_rn(two_nf, __fmul_rn(__fdiv_rn(threshold, 100.f), two_nf)), 0.5f);
}

// Accumulate one (modality, word-shift) bucket of feature rows into 16 packed byte sums.
// lmb: this modality's linear memories + 16*chunk (16 B aligned).  WS = (off>>2)&3 is uniform per bucket,
// so the realignment is straight-line code: two aligned LDG.128 + four funnel shifts per feature.
template <int WS, bool SAFE>
__device__ __forceinline__ void accum_bucket(const u8* __restrict__ lmb, const u32* __restrict__ lst,
                                             const u32* __restrict__ lim, int n, int pos0, int P,
                                             u32& a0, u32& a1, u32& a2, u32& a3) {
#pragma unroll 4
  for (int k = 0; k < n; ++k) {
    const u32 off = lst[k];
    int nv = 16;
    if (!SAFE) {
      nv = min(P, (int)lim[k]) - pos0;  // positions past the label block contribute 0 (oracle semantics)
      if (nv <= 0) continue;            // ... and are never loaded
    }
    const uint4* p = reinterpret_cast<const uint4*>(lmb + (off & ~15u));
    const uint4 A = __ldg(p), B = __ldg(p + 1);
    const u32 sh = (off & 3u) * 8u;
    u32 w0, w1, w2, w3;
    if (WS == 0) {
      w0 = __funnelshift_r(A.x, A.y, sh); w1 = __funnelshift_r(A.y, A.z, sh);
      w2 = __funnelshift_r(A.z, A.w, sh); w3 = __funnelshift_r(A.w, B.x, sh);
    } else if (WS == 1) {
      w0 = __funnelshift_r(A.y, A.z, sh); w1 = __funnelshift_r(A.z, A.w, sh);
      w2 = __funnelshift_r(A.w, B.x, sh); w3 = __funnelshift_r(B.x, B.y, sh);
    } else if (WS == 2) {
      w0 = __funnelshift_r(A.z, A.w, sh); w1 = __funnelshift_r(A.w, B.x, sh);
      w2 = __funnelshift_r(B.x, B.y, sh); w3 = __funnelshift_r(B.y, B.z, sh);
    } else {
      w0 = __funnelshift_r(A.w, B.x, sh); w1 = __funnelshift_r(B.x, B.y, sh);
      w2 = __funnelshift_r(B.y, B.z, sh); w3 = __funnelshift_r(B.z, B.w, sh);
    }
    if (!SAFE && nv < 16) mask16(w0, w1, w2, w3, nv);
    a0 += w0; a1 += w1; a2 += w2; a3 += w3;
  }
}

// ---------------------------------------------------------------------------------------------
// Coarse similarity: grid (template, frame); each thread owns chunks of 16 contiguous positions.
// ---------------------------------------------------------------------------------------------
template <bool SAFE>
__device__ __forceinline__ void coarse_body(const MatchParams& mp, const LevelParams& lp, const HdrR& hdr,
                                            const u32* lst, const u32* lim, const int* cnt, u16* bm, int NC,
                                            const u8* const* s_lm, int raw_thr) {
  const int M = mp.M, T = lp.g.T, W = lp.g.W, H = lp.g.H, HW = W * H;
  for (int c = threadIdx.x; c < NC; c += blockDim.x) {
    const int pos0 = 16 * c;
    u32 tl0 = 0, tl1 = 0, tl2 = 0, tl3 = 0, th0 = 0, th1 = 0, th2 = 0, th3 = 0;  // u16 pairs: (4w,4w+2) / (4w+1,4w+3)
    for (int m = 0; m < M; ++m) {
      int wf = (hdr.width(m) - 1) / T + 1, hf = (hdr.height(m) - 1) / T + 1;
      int P = (H - hf) * W + (W - wf) + 1;
      if (P > HW) P = HW;
      const int rem = P - pos0;
      if (rem <= 0) continue;
      u32 a0 = 0, a1 = 0, a2 = 0, a3 = 0;
      const u8* lmb = s_lm[m] + pos0;
      const u32* l = lst + m * 4 * FEAT_SLOTS;
      const u32* li = lim + m * 4 * FEAT_SLOTS;
      accum_bucket<0, SAFE>(lmb, l, li, cnt[m * 4 + 0], pos0, P, a0, a1, a2, a3);
      accum_bucket<1, SAFE>(lmb, l + FEAT_SLOTS, li + FEAT_SLOTS, cnt[m * 4 + 1], pos0, P, a0, a1, a2, a3);
      accum_bucket<2, SAFE>(lmb, l + 2 * FEAT_SLOTS, li + 2 * FEAT_SLOTS, cnt[m * 4 + 2], pos0, P, a0, a1, a2, a3);
      accum_bucket<3, SAFE>(lmb, l + 3 * FEAT_SLOTS, li + 3 * FEAT_SLOTS, cnt[m * 4 + 3], pos0, P, a0, a1, a2, a3);
      if (SAFE && rem < 16) mask16(a0, a1, a2, a3, rem);
      tl0 += a0 & 0x00FF00FFu; th0 += (a0 >> 8) & 0x00FF00FFu;
      tl1 += a1 & 0x00FF00FFu; th1 += (a1 >> 8) & 0x00FF00FFu;
      tl2 += a2 & 0x00FF00FFu; th2 += (a2 >> 8) & 0x00FF00FFu;
      tl3 += a3 & 0x00FF00FFu; th3 += (a3 >> 8) & 0x00FF00FFu;
    }
    u32 mask = 0;
    const u32 tl[4] = {tl0, tl1, tl2, tl3}, th[4] = {th0, th1, th2, th3};
#pragma unroll
    for (int w = 0; w < 4; ++w) {
      mask |= ((int)(tl[w] & 0xFFFFu) > raw_thr ? 1u : 0u) << (4 * w);
      mask |= ((int)(th[w] & 0xFFFFu) > raw_thr ? 1u : 0u) << (4 * w + 1);
      mask |= ((int)(tl[w] >> 16) > raw_thr ? 1u : 0u) << (4 * w + 2);
      mask |= ((int)(th[w] >> 16) > raw_thr ? 1u : 0u) << (4 * w + 3);
    }
    // raw scores of the hits are needed again at emission: stash them next to the bitmap
    bm[c] = (u16)mask;
    if (mask) {
      u16* raw = bm + ((NC + 1) & ~1) + 16 * c;  // raw[NC][16], only touched for chunks with hits
#pragma unroll
      for (int w = 0; w < 4; ++w) {
        raw[4 * w + 0] = (u16)(tl[w] & 0xFFFFu); raw[4 * w + 1] = (u16)(th[w] & 0xFFFFu);
        raw[4 * w + 2] = (u16)(tl[w] >> 16);     raw[4 * w + 3] = (u16)(th[w] >> 16);
      }
    }
  }
}

__global__ void __launch_bounds__(256) similarity_coarse_kernel(MatchParams mp, LevelParams lp) {
  extern __shared__ __align__(16) u32 cs_smem[];
  const int M = mp.M;
  u32* lst = cs_smem;                               // [M][4][64] feature offsets bucketed by word shift
  u32* lim = lst + M * 4 * FEAT_SLOTS;              // [M][4][64] per_label - base (guard for malformed templates)
  int* cnt = (int*)(lim + M * 4 * FEAT_SLOTS);      // [M*4]
  int* scan = cnt + 16;                             // [10]: warp totals, base
  u16* bm = (u16*)(scan + 12);                      // [NC] hit bitmap + [NC][16] raw scores

  const int isel = blockIdx.x, frame = blockIdx.y;
  const int g = mp.sel[isel];
  const HdrR hdr = load_hdr(lp.hdr + g);
  const int tid = threadIdx.x, NT = blockDim.x;
  const int T = lp.g.T, W = lp.g.W, H = lp.g.H, HW = W * H;
  const int NC = (HW + 15) >> 4;

  __shared__ const u8* s_lm[MAX_MOD];
  if (tid == 0) {
    s_lm[0] = lp.lm[0] + (size_t)frame * lp.lm_stride[0];
    s_lm[1] = lp.lm[1] + (size_t)frame * lp.lm_stride[1];
    s_lm[2] = lp.lm[2] + (size_t)frame * lp.lm_stride[2];
    s_lm[3] = lp.lm[3] + (size_t)frame * lp.lm_stride[3];
  }
  if (tid < M * 4) cnt[tid] = 0;
  __syncthreads();
  const int nf_total = hdr.nf_total();
  for (int s = tid; s < M * FEAT_SLOTS; s += NT) {
    int m = s / FEAT_SLOTS, k = s - m * FEAT_SLOTS;
    if (k < hdr.nf(m)) {
      u32 off = lp.offs[(size_t)g * M * FEAT_SLOTS + s];
      if (off != OFF_INVALID) {
        int b = m * 4 + ((off >> 2) & 3);
        int pos = atomicAdd(&cnt[b], 1);
        lst[b * FEAT_SLOTS + pos] = off;
        lim[b * FEAT_SLOTS + pos] = lp.g.per_label - off % lp.g.per_label;
      }
    }
  }
  __syncthreads();
  const int raw_thr = raw_threshold(nf_total, mp.threshold);
  if (hdr.flags & 2u) coarse_body<true>(mp, lp, hdr, lst, lim, cnt, bm, NC, s_lm, raw_thr);
  else coarse_body<false>(mp, lp, hdr, lst, lim, cnt, bm, NC, s_lm, raw_thr);
  __syncthreads();

  // ---- emission in raster order: thread t owns chunks [t*CH, (t+1)*CH)
  const int CH = (NC + NT - 1) / NT;
  const int cb = tid * CH, ce = min(NC, cb + CH);
  int mine = 0;
  for (int c = cb; c < ce; ++c) mine += __popc((u32)bm[c]);
  // block exclusive scan
  const int lane = tid & 31, wid = tid >> 5;
  int incl = mine;
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
    int v = __shfl_up_sync(0xffffffffu, incl, d);
    if (lane >= d) incl += v;
  }
  if (lane == 31) scan[wid] = incl;
  __syncthreads();
  if (tid == 0) {
    int run = 0;
    for (int w = 0; w < (NT + 31) / 32; ++w) { int v = scan[w]; scan[w] = run; run += v; }
    int base = 0;
    if (run > 0) {
      base = atomicAdd(&mp.cand_count[frame], run);
      if (base + run > mp.cand_cap) { mp.overflow[frame] = 1; run = 0; }
    }
    scan[8] = base; scan[9] = run;
    mp.tpl_start[(size_t)frame * mp.nsel_stride + isel] = base;
    mp.tpl_cnt[(size_t)frame * mp.nsel_stride + isel] = run;
  }
  __syncthreads();
  const int total = scan[9];
  if (total == 0) return;
  int pos = scan[8] + scan[wid] + incl - mine;
  Cand* out = mp.cand + (size_t)frame * mp.cand_cap;
  const int offset = T / 2 + (T % 2 - 1);
  const float denom = (float)(4 * nf_total);
  const u16* rawbase = bm + ((NC + 1) & ~1);
  for (int c = cb; c < ce; ++c) {
    u32 m = bm[c];
    while (m) {
      int b = __ffs(m) - 1;
      m &= m - 1;
      int j = 16 * c + b;
      int r = j / W, col = j - r * W;
      int raw = rawbase[16 * c + b];
      Cand cd;
      cd.tsel = isel;
      cd.x = col * T + offset;
      cd.y = r * T + offset;
      cd.sim = __fadd_rn(__fdiv_rn(__fmul_rn((float)raw, 100.f), denom), 0.5f);
      out[pos++] = cd;
    }
  }
}

static size_t coarse_smem_bytes(int M, int HW) {
  int NC = (HW + 15) >> 4;
  size_t b = (size_t)M * 4 * FEAT_SLOTS * 4 * 2 + 16 * 4 + 12 * 4;
  b += (size_t)((NC + 1) & ~1) * 2 + (size_t)NC * 16 * 2;
  return b;
}

void launch_similarity_coarse(const MatchParams& mp, const LevelParams& lp, cudaStream_t st) {
  if (mp.nsel <= 0 || mp.frames <= 0) return;
  int HW = lp.g.W * lp.g.H;
  int NC = (HW + 15) >> 4;
  int NT = ((NC + 31) / 32) * 32;
  if (NT > 256) NT = 256;
  if (NT < 32) NT = 32;
  size_t smem = coarse_smem_bytes(mp.M, HW);
  static size_t configured = 0;
  if (smem > 48 * 1024 && smem > configured) {
    cudaFuncSetAttribute(similarity_coarse_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    configured = smem;
  }
  dim3 grid(mp.nsel, mp.frames);
  similarity_coarse_kernel<<<grid, NT, smem, st>>>(mp, lp);
}

// ---------------------------------------------------------------------------------------------
// Local refinement: one warp per candidate; lane = (row 0..15, half 0..1) owns 8 of the 16x16 sums.
// Fast path (hdr.flags bit0): offsets are the plan offsets shifted by the candidate's patch origin.
// Slow path: upstream's per-feature bounds checks with guarded byte loads (malformed / oversized
// templates — N4 in SURVEY.md — where upstream itself is undefined; semantics = oracle's).
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128) similarity_local_kernel(MatchParams mp, LevelParams lp) {
  const int frame = blockIdx.y;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (mp.overflow[frame]) return;  // candidate store overflowed: the host grows it and redoes this frame
  const int n = min(mp.cand_count[frame], mp.cand_cap);
  const int T = lp.g.T, W = lp.g.W, M = mp.M;
  const int border = 8 * T, offset = T / 2 + (T % 2 - 1);
  const u32 plane = (u32)W * lp.g.H;
  Cand* cands = mp.cand + (size_t)frame * mp.cand_cap;
  const int rr = lane >> 1, hh = lane & 1;
  __shared__ const u8* s_lm[MAX_MOD];
  if (threadIdx.x == 0) {
    s_lm[0] = lp.lm[0] + (size_t)frame * lp.lm_stride[0];
    s_lm[1] = lp.lm[1] + (size_t)frame * lp.lm_stride[1];
    s_lm[2] = lp.lm[2] + (size_t)frame * lp.lm_stride[2];
    s_lm[3] = lp.lm[3] + (size_t)frame * lp.lm_stride[3];
  }
  __syncthreads();

  for (int c = blockIdx.x * 4 + warp; c < n; c += gridDim.x * 4) {
    Cand rec = cands[c];
    if (rec.sim < 0.f || rec.tsel < 0 || rec.tsel >= mp.nsel) continue;
    const int g = mp.sel[rec.tsel];
    const HdrR hdr = load_hdr(lp.hdr + g);
    int x = rec.x * 2 + 1, y = rec.y * 2 + 1;
    x = max(x, border); y = max(y, border);
    x = min(x, lp.g.cols - hdr.width(0) - border);
    y = min(y, lp.g.rows - hdr.height(0) - border);
    const int cxT = x / T - 8, cyT = y / T - 8;
    u32 t0 = 0, t1 = 0, t2 = 0, t3 = 0;  // u16 pairs: positions (0,2) (1,3) (4,6) (5,7) of this lane's 8
    int nfl = 0;
    for (int m = 0; m < M; ++m) {
      const int nf = hdr.nf(m);
      nfl += nf;
      u32 a0 = 0, a1 = 0;
      const u8* lmb = s_lm[m];
      if (hdr.flags & 1u) {
        const int shift = (cyT + rr) * W + cxT + hh * 8;
        const u32* offp = lp.offs + (size_t)g * M * FEAT_SLOTS + m * FEAT_SLOTS;
        for (int k0 = 0; k0 < nf; k0 += 32) {
          u32 myoff = (k0 + lane < nf) ? __ldg(offp + k0 + lane) : 0u;
          const int kn = min(32, nf - k0);
#pragma unroll 4
          for (int kk = 0; kk < kn; ++kk) {
            u32 a = __shfl_sync(0xffffffffu, myoff, kk) + (u32)shift;
            const uint2* p = reinterpret_cast<const uint2*>(lmb + (a & ~7u));
            uint2 A = __ldg(p), B = __ldg(p + 1);
            u32 sh = (a & 3u) * 8u;
            bool up = (a & 4u) != 0;
            u32 lo = up ? A.y : A.x, mid = up ? B.x : A.y, hi = up ? B.y : B.x;
            a0 += __funnelshift_r(lo, mid, sh);
            a1 += __funnelshift_r(mid, hi, sh);
          }
        }
      } else {
        const u32* fp = lp.feat + (size_t)g * M * FEAT_SLOTS + m * FEAT_SLOTS;
        for (int k = 0; k < nf; ++k) {
          u32 f = __ldg(fp + k);
          if (!(f >> 31)) continue;
          int fx = (int)(f & 0x3FFF) + cxT * T, fy = (int)((f >> 14) & 0x3FFF) + cyT * T;
          int label = (f >> 28) & 7;
          if (fx < 0 || fy < 0 || fx >= lp.g.cols || fy >= lp.g.rows) continue;
          long long base = (long long)((fy % T) * T + (fx % T)) * plane + (long long)(fy / T) * W + fx / T;
          const u8* lml = lmb + (size_t)label * lp.g.per_label;
          long long rb = base + (long long)rr * W + hh * 8;
#pragma unroll
          for (int b = 0; b < 8; ++b) {
            u32 v = (rb + b < (long long)lp.g.per_label) ? (u32)lml[rb + b] : 0u;
            if (b < 4) a0 += v << (8 * b); else a1 += v << (8 * (b - 4));
          }
        }
      }
      t0 += a0 & 0x00FF00FFu; t1 += (a0 >> 8) & 0x00FF00FFu;
      t2 += a1 & 0x00FF00FFu; t3 += (a1 >> 8) & 0x00FF00FFu;
    }
    // argmax, strict '>' in raster order from best = 0
    int v[8] = {(int)(t0 & 0xFFFF), (int)(t1 & 0xFFFF), (int)(t0 >> 16), (int)(t1 >> 16),
                (int)(t2 & 0xFFFF), (int)(t3 & 0xFFFF), (int)(t2 >> 16), (int)(t3 >> 16)};
    int best = 0, bi = 0;
#pragma unroll
    for (int b = 0; b < 8; ++b)
      if (v[b] > best) { best = v[b]; bi = b; }
    u32 key = best > 0 ? ((u32)best << 8) | (u32)(255 - (rr * 16 + hh * 8 + bi)) : 0u;
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) key = max(key, __shfl_xor_sync(0xffffffffu, key, d));
    if (lane == 0) {
      int br = -1, bc = -1, bs = 0;
      if (key) { int idx = 255 - (int)(key & 255u); br = idx >> 4; bc = idx & 15; bs = (int)(key >> 8); }
      rec.x = (cxT + bc) * T + offset;
      rec.y = (cyT + br) * T + offset;
      float sim = __fdiv_rn(__fmul_rn((float)bs, 100.f), (float)(4 * nfl));
      rec.sim = sim < mp.threshold ? -1.0f : sim;
      cands[c] = rec;
      if (mp.stats) atomicAdd(&mp.stats[frame], (unsigned long long)nfl * 256ull);
    }
  }
}

void launch_similarity_local(const MatchParams& mp, const LevelParams& lp, cudaStream_t st) {
  if (mp.nsel <= 0 || mp.frames <= 0) return;
  // persistent grid: the candidate count lives on the device, warps stride over it
  int bx = 148 * 4;
  if (mp.frames >= 8) bx = 148;
  dim3 grid(bx, mp.frames);
  similarity_local_kernel<<<grid, 128, 0, st>>>(mp, lp);
}

// ---------------------------------------------------------------------------------------------
// Ordered compaction: surviving candidates of frame f in (selection order, raster order).
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(1024) pack_kernel(MatchParams mp, Cand* __restrict__ out, int out_cap,
                                                    int* __restrict__ out_count) {
  __shared__ int wsum[32];
  __shared__ int s_run, s_total;
  const int frame = blockIdx.x, tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  const Cand* cands = mp.cand + (size_t)frame * mp.cand_cap;
  Cand* o = out + (size_t)frame * out_cap;
  if (tid == 0) s_run = 0;
  __syncthreads();
  if (mp.overflow[frame]) {  // host will grow the store and redo this frame
    if (tid == 0) out_count[frame] = 0;
    return;
  }
  for (int i0 = 0; i0 < mp.nsel; i0 += 1024) {
    const int i = i0 + tid;
    int st = 0, cn = 0, alive = 0;
    if (i < mp.nsel) {
      st = mp.tpl_start[(size_t)frame * mp.nsel_stride + i];
      cn = mp.tpl_cnt[(size_t)frame * mp.nsel_stride + i];
      for (int j = 0; j < cn; ++j) alive += cands[st + j].sim >= 0.f;
    }
    int incl = alive;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      int v = __shfl_up_sync(0xffffffffu, incl, d);
      if (lane >= d) incl += v;
    }
    if (lane == 31) wsum[wid] = incl;
    __syncthreads();
    if (wid == 0) {
      int v = wsum[lane], inc2 = v;
#pragma unroll
      for (int d = 1; d < 32; d <<= 1) {
        int u = __shfl_up_sync(0xffffffffu, inc2, d);
        if (lane >= d) inc2 += u;
      }
      wsum[lane] = inc2 - v;          // exclusive warp offsets
      if (lane == 31) s_total = inc2; // candidates surviving in this round of 1024 templates
    }
    __syncthreads();
    int pos = s_run + wsum[wid] + incl - alive;
    for (int j = 0; j < cn; ++j) {
      Cand cd = cands[st + j];
      if (cd.sim >= 0.f) {
        cd.tsel = mp.sel[cd.tsel];  // selection index -> global template index (rank-independent)
        if (pos < out_cap) o[pos] = cd;
        ++pos;
      }
    }
    __syncthreads();
    if (tid == 0) s_run += s_total;
    __syncthreads();
  }
  if (tid == 0) out_count[frame] = s_run;
}

void launch_pack(const MatchParams& mp, Cand* out, int out_cap, int* out_count, cudaStream_t st) {
  if (mp.frames <= 0) return;
  pack_kernel<<<mp.frames, 1024, 0, st>>>(mp, out, out_cap, out_count);
}

}  // namespace lmk
