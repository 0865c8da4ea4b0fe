// kernels.cuh — device data layout and kernel launch interface (sm_100a).
// One launcher per upstream function group: every launcher takes a frame
// count and per-frame strides so a whole batch is one launch (grid.z / grid.y = frame).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace lmk {

typedef uint8_t u8;
typedef uint16_t u16;
typedef uint32_t u32;

constexpr int MAX_MOD = 4;
constexpr int FEAT_SLOTS = 64;        // per modality per template (upstream asserts <= 63 features)
constexpr int COARSE_SLOTS = 72;      // coarsest-level plan: 63 features + padding of every word-shift bucket to a multiple of 3
constexpr u32 OFF_INVALID = 0xFFFFFFFFu;
constexpr int LM_PAD = 256;           // zero slack after each linear-memory block (realigned 16 B loads overrun)

// Packed feature: x[0:14) y[14:28) label[28:31) valid[31].  Invalid = coordinate negative / > 16383
// (upstream skips features outside the image; negative ones can never be inside).
__host__ __device__ inline u32 pack_feature(int x, int y, int label) {
  if (x < 0 || y < 0 || x > 16383 || y > 16383) return 0u;
  return (u32)x | ((u32)y << 14) | ((u32)(label & 7) << 28) | 0x80000000u;
}

// Per (level, template) header.  width/height are per modality (cropTemplates makes them equal,
// synthetic templates may differ; upstream's refinement clamp uses modality 0).
struct alignas(8) TplHdr {
  short width[MAX_MOD];
  short height[MAX_MOD];
  u8 nf[MAX_MOD];
  u32 flags;                          // bit0: local-safe, bit1: coarse-safe (set by the plan kernel)
  u32 bkt[MAX_MOD];                   // plan: per modality, 4 x u8 counts of (padded) offsets per word-shift bucket
  int P[MAX_MOD];                     // plan: upstream's template_positions per modality (clamped to W*H)
};

// Register-resident view of a TplHdr (dynamic modality index without local memory).
struct HdrR {
  unsigned long long w, h, b01, b23;
  u32 nfs, flags;
  __device__ __forceinline__ u32 bkt(int m) const { unsigned long long v = (m & 2) ? b23 : b01; return (u32)(v >> (32 * (m & 1))); }
  __device__ __forceinline__ int width(int m) const { return (int)(short)(w >> (16 * m)); }
  __device__ __forceinline__ int height(int m) const { return (int)(short)(h >> (16 * m)); }
  __device__ __forceinline__ int nf(int m) const { return (int)((nfs >> (8 * m)) & 0xFFu); }
  __device__ __forceinline__ int nf_total() const {
    return (int)((nfs & 0xFF) + ((nfs >> 8) & 0xFF) + ((nfs >> 16) & 0xFF) + (nfs >> 24));
  }
};
#ifdef __CUDACC__
__device__ __forceinline__ HdrR load_hdr(const TplHdr* p) {
  const unsigned long long* q = reinterpret_cast<const unsigned long long*>(p);
  HdrR r;
  r.w = __ldg(q); r.h = __ldg(q + 1);
  unsigned long long t = __ldg(q + 2);
  r.nfs = (u32)t; r.flags = (u32)(t >> 32);
  r.b01 = __ldg(q + 3); r.b23 = __ldg(q + 4);
  return r;
}
#endif

// Candidate / match record kept on the device (16 B).  tsel = index into the selection list.
struct Cand {
  int tsel;
  int x, y;
  float sim;                          // < 0 : filtered out
};

// Per-slot device counters, one contiguous record so a chunk needs one memset and one D2H copy.
struct SlotCtr {
  int cand_count;                     // coarse candidates appended (may exceed the store: overflow)
  int overflow;                       // candidate store overflowed -> host grows it and redoes the frame
  int out_count;                      // surviving matches written by pack_kernel
  int pad;
  unsigned long long local_bytes;     // algorithmic bytes gathered by similarityLocal (sum nf*256)
  unsigned long long coarse_chunks;   // 16-byte chunk loads the coarse kernel issued (plan rows executed x active lanes x 2):
                                      // the bytes it actually REQUESTED, after the early exit (bench roofline)
};

// Linear-memory layouts.  Coarsest level (strips == 0): upstream's flat rows, LM[label][phase][y/T * W + x/T], which the
// coarse kernel needs (its reads run across grid rows exactly like upstream's).  Finer levels (strips > 0) are only
// read as 16x16 patches by similarityLocal and are stored in 16-column strips, LM[label][phase][strip][row][16]:
// the 16 rows of a patch are 256 contiguous bytes per strip, so a patch touches 4-6 cache lines instead of 16.
struct LevelGeom {
  int T, rows, cols, W, H;            // quantized image size; W=cols/T, H=rows/T
  u32 per_label;                      // bytes per label in the stored layout: T*T*plane
  int strips;                         // 0: flat layout; else ceil(W/16)
  u32 plane;                          // bytes per (label, phase) plane: W*H (flat) or strips*H*16
};

// ------------------------------------------------------------------ frame side
void launch_pyrdown_bgr(const u8* src, size_t src_stride, u8* dst, size_t dst_stride,
                        int rows, int cols, int frames, cudaStream_t st);
void launch_cg_quantize(const u8* bgr, size_t bgr_stride, u8* q, size_t q_stride,
                        float* mag, size_t mag_stride /*elements; mag nullable*/,
                        int rows, int cols, float weak_sq, int frames, cudaStream_t st);
void launch_dn_quantize(const u16* depth, size_t depth_stride /*elements*/, u8* out, size_t out_stride,
                        int8_t* idx_out /*nullable [frames][3][rows*cols]*/,
                        int rows, int cols, int dist_thr, int diff_thr, const u8* lut_dev,
                        int frames, cudaStream_t st);
void launch_median5(const u8* src, size_t src_stride, u8* dst, size_t dst_stride, int rows, int cols,
                    int frames, cudaStream_t st);
void launch_resize_nn(const u8* src, size_t src_stride, int rows, int cols,
                      u8* dst, size_t dst_stride, int drows, int dcols, int frames, cudaStream_t st);
// spread (T x T OR) + response LUT + linearize, fused.  table: 256 x uint2 (8 orientation bytes per
// spread byte).  lm: per frame [8][T*T*W*H] (+LM_PAD).  mask nullable (quantize() through mask).
void launch_spread_linearize(const u8* q, size_t q_stride, const u8* mask, size_t mask_stride,
                             u8* lm, size_t lm_stride, LevelGeom g, const uint2* table,
                             int frames, cudaStream_t st);

// Round-2 frame-side kernels (kernels_color.cu, kernels_depth.cu).  pyrDown writes byte PLANES ([3][rows][cols]); the
// quantiser reads either the interleaved BGR frame (level 0) or planes (levels >= 1).
void launch_pyrdown_planar(const u8* src, size_t src_stride, bool src_planar, u8* dst, size_t dst_stride, int rows, int cols,
                           int frames, cudaStream_t st);
void launch_cg_quantize2(const u8* src, size_t src_stride, bool src_planar, u8* q, size_t q_stride, float* mag /*nullable*/,
                         size_t mag_stride, int rows, int cols, float weak_sq, int frames, cudaStream_t st);
// quantizedNormals + medianBlur(5) in one kernel
void launch_dn_median(const u16* depth, size_t depth_stride /*elements*/, u8* out, size_t out_stride, int rows, int cols,
                      int dist_thr, int diff_thr, const u8* lut_dev, int frames, cudaStream_t st);

// Round-2 fused spread + response + linearize (kernels_spread.cu).  Finer levels (g.strips > 0) write the strip layout
// into `lm`; the coarsest level writes the nibble-packed flat layout straight into `lmn` (+ resp_sum).  q_step > 1 reads
// the quantized map of a finer level in place: pixel (y, x) of this level = q[y*q_step][x*q_step] (DepthNormal pyrDown).
struct SpreadArgs {
  const u8* q; size_t q_stride;       // quantized map and its per-frame stride
  int q_pitch, q_step;                // bytes per row of the map `q` points to; sampling step (1 = this level's own map)
  const u8* mask; size_t mask_stride; // optional mask at THIS level's resolution (pitch = g.cols), nullable
  u8* lm; size_t lm_stride;           // strip-layout output (finer levels)
  u8* lmn; size_t lmn_stride;         // nibble-packed flat output (coarsest level)
  u32* resp_sum; int resp_stride;     // += sampled response sum per frame (coarsest level), nullable
  LevelGeom g;
  const uint2* table;                 // 256 x uint2: the 8 orientation responses of one spread byte
};
// false: geometry not covered by the fast kernels (T not in {2,4,5,8,16}, or a coarsest level with W % 8 != 0) -> the
// caller runs launch_spread_linearize (+ launch_pack_nibbles) instead.
bool launch_spread_fast(const SpreadArgs& a, int frames, cudaStream_t st);
inline bool spread_fast_covers(const LevelGeom& g) {
  const bool t_ok = g.T == 2 || g.T == 4 || g.T == 5 || g.T == 8 || g.T == 16;
  return t_ok && (g.strips ? true : (g.W & 7) == 0);
}

// ------------------------------------------------------------------ template side
// Plan: flat linear-memory offsets of every feature for one level's geometry + safe flags.
// coarsest = true: nibble-space buckets ((off>>3)&3), COARSE_SLOTS per modality, every bucket padded to a multiple of 3
// with offsets into the zero tail behind the nibble-packed linear memory (so the coarse kernel runs groups of 3 only).
void launch_build_offsets(const u32* feat, u32* offs, TplHdr* hdr, int ntpl, int M, LevelGeom g,
                          bool coarsest, cudaStream_t st);
// bytes of the zero tail the coarse kernel may read behind a nibble-packed block of 4*per_label bytes
inline size_t coarse_zero_tail(const LevelGeom& g) { return (((size_t)g.W * g.H / 2 + 64) + 255) & ~(size_t)255; }
// byte linear memory -> nibble-packed copy (coarsest level only; input of similarity_coarse_kernel)
void launch_pack_nibbles(const u8* lm, size_t lm_stride, u8* lmn, size_t lmn_stride, LevelGeom g, int frames,
                         u32* resp_sum /*+= responses, per frame*/, int resp_stride, cudaStream_t st);

struct MatchParams {
  int M, nsel, frames;
  int early_exit;                     // 1 (default): the coarse kernel's exact early exit is on; 0: measurement runs without it
  const int* sel;                     // selection list: global template indices, generation order
  float threshold;
  // candidate store, per frame
  Cand* cand; int cand_cap;
  SlotCtr* ctr;                       // [frames]
  int* tpl_start; int* tpl_cnt;       // [frames][nsel_stride] candidate block of every template
  int* tpl_alive;                     // [frames][nsel_stride] candidates of the block still alive (local kernel decrements)
  int nsel_stride;
};

struct LevelParams {
  LevelGeom g;
  const u8* lm[MAX_MOD]; size_t lm_stride[MAX_MOD];   // this level's linear memories per modality
  const TplHdr* hdr; const u32* offs; const u32* feat; // this level's template tables
  const u32* resp_sum;                                 // coarsest level: [frames][MAX_MOD] response sums (modality order)
};

void launch_similarity_coarse(const MatchParams& mp, const LevelParams& lp, bool wide, cudaStream_t st);
void launch_similarity_local(const MatchParams& mp, const LevelParams& lp, cudaStream_t st);
// debug: full u16 coarse similarity map of the single template mp.sel[0] on one frame, early exit disabled
void launch_similarity_map(const MatchParams& mp, const LevelParams& lp, bool wide, u16* map, cudaStream_t st);
// Ordered compaction: out[frame][0..n) in generation order, count[frame] = n (may exceed cap -> overflow).
void launch_pack(const MatchParams& mp, Cand* out, int out_cap, cudaStream_t st);
// multi-GPU: compact send buffer of the match all-gather: [2*frames header records][rec_cap match records, frames back to back]
void launch_gather_pack(const SlotCtr* ctr, const Cand* out, int out_cap, Cand* send, int rec_cap, int frames, cudaStream_t st);

// ------------------------------------------------------------------ device epilogue of the template-sharded step (kernels_epilogue.cu)
constexpr int EPI_SORT_CAP = 4096;     // records per frame (all ranks together) the device sort takes; longer lists go to the host
struct EpiMatch { int x, y; float sim; int class_index, template_id; };   // = lmb200_match_rec
struct EpilogueArgs {
  const Cand* gathered;                // [world][2*frames header records + gcap match records] (gather_pack_kernel, all-gathered)
  int world, rank, frames, gcap;
  const int* pos_of_g;                 // interleaved shards: global template index -> selection position; null: rank-ordered concatenation
  const int* g_class; const int* g_tid;  // global template index -> class index / template id
  EpiMatch* out_dev; EpiMatch* out_host; int out_cap;   // device scratch and its pinned host mirror (device-mapped), records
  int4* hdr;                           // pinned [2*frames]: {n_final, offset, flags, n_in}, {this rank's local_bytes lo, hi, coarse_chunks lo, hi}
};                                     // flags: 1/2 from the gathered headers (store overflow / record area too small), 4 list too long, 8 output too small
void launch_shard_epilogue(const EpilogueArgs& a, cudaStream_t st);

// ------------------------------------------------------------------ post-match colour check (kernels_postmatch.cu)
// bits: [rows][(cols+31)/32] u32, bit x%32 of word x/32 = pixel (y, x) lies in the HSV range (cvtColor BGR2HSV + inRange)
void launch_hsv_inrange_bits(const u8* bgr, int rows, int cols, const u8 lower[3], const u8 upper[3], u32* bits, cudaStream_t st);
// per match: result = {countNonZero(hue & templateMask), countNonZero(templateMask)}; (-1,-1) when the hull leaves the image
void launch_template_mask_count(const int2* xy, int n, const int* g_of_match, const TplHdr* hdr0, const u32* feat0, int M, int rows,
                                int cols, const u32* hue_bits, int2* result, cudaStream_t st);

}  // namespace lmk
