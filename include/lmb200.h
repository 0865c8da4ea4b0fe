/* lmb200.h — C ABI of the B200-native LINE-MOD matcher.
 *
 * Drop-in boundary for the `cv::linemod::Detector` path that aelmiger/LINE-MOD-Pipeline
 * drives from HighLevelLineMOD (reference file:line each entry point replaces is cited).
 * Plain pointers and sizes only; no C++/torch types; no exceptions cross this ABI.
 * All functions return LMB200_OK (0) or a negative lmb200_status, unless stated otherwise.
 *
 * Image conventions (reference: detector.cpp:12,:24-26, HighLevelLinemod.cpp:86-90):
 *   sources[] are in modality order; ColorGradient takes BGR8 (LMB200_8UC3),
 *   DepthNormal takes 16-bit depth in millimetres (LMB200_16UC1); masks are LMB200_8UC1.
 *   `step` is the row pitch in bytes (0 = tightly packed).  Buffers are caller-owned host
 *   memory, read-only for the duration of the call.
 */
#ifndef LMB200_H
#define LMB200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct lmb200_detector* lmb200_handle;

typedef enum {
  LMB200_OK = 0,
  LMB200_E_INVALID = -1,    /* bad argument / handle */
  LMB200_E_SOURCES = -2,    /* upstream CV_Assert(sources.size()==modalities.size()), wrong type or size */
  LMB200_E_SIZE = -3,       /* upstream CV_Assert in linearize/computeResponseMaps: rows%T, cols%T, (rows*cols)%16 */
  LMB200_E_FEATURES = -4,   /* upstream CV_Assert(templ.features.size() <= 63) */
  LMB200_E_CLASS = -5,      /* unknown class / class already present (readClass) / bad template id */
  LMB200_E_IO = -6,         /* file open / parse error */
  LMB200_E_CUDA = -7,       /* CUDA runtime error (message in lmb200_last_error) */
  LMB200_E_TRUNCATED = -8,  /* output buffer too small; *n_out holds the required count */
  LMB200_E_COMM = -9,       /* NCCL not available / communicator error */
  LMB200_E_NODEVICE = -10   /* no CUDA device: the library has NO CPU fallback */
} lmb200_status;

/* OpenCV type codes, so a cv::Mat's type() can be passed through unchanged. */
enum { LMB200_8UC1 = 0, LMB200_16UC1 = 2, LMB200_8UC3 = 16 };
enum { LMB200_COLOR_GRADIENT = 0, LMB200_DEPTH_NORMAL = 1 };

/* Modality parameters; defaults are upstream's (linemod.hpp ColorGradient()/DepthNormal()),
 * which the reference never overrides (HighLevelLinemod.cpp:28-30,:38). */
typedef struct {
  int type;                 /* LMB200_COLOR_GRADIENT | LMB200_DEPTH_NORMAL */
  float weak_threshold;     /* CG: 10.0 */
  int num_features;         /* 63 */
  float strong_threshold;   /* CG: 55.0 */
  int distance_threshold;   /* DN: 2000 */
  int difference_threshold; /* DN: 50 */
  int extract_threshold;    /* DN: 2 */
} lmb200_modality;

#define LMB200_MAX_MODALITIES 4
#define LMB200_MAX_LEVELS 8

/* Replaces: cv::linemod::Detector(modalities, T_pyramid)  (HighLevelLinemod.cpp:26-43). */
typedef struct {
  int num_modalities;
  lmb200_modality modalities[LMB200_MAX_MODALITIES];
  int pyramid_levels;
  int T[LMB200_MAX_LEVELS];
  int device;               /* CUDA ordinal; -1 = current device */
  int max_batch;            /* frames resident per batch call; 0 = default (64) */
  int candidate_capacity;   /* coarse candidates per frame; 0 = default (16384); grows on overflow */
  int similarity_lut;       /* LMB200_SIMLUT_CIRCULAR (0, default: upstream's table) | LMB200_SIMLUT_LINEAR */
} lmb200_config;

/* The two SIMILARITY_LUT variants (the table is data: lmb200_set_similarity_lut takes any other).
 * CIRCULAR: response = 4 - min(|ori-j|, 8-|ori-j|), sum 628 — SURVEY.md 8c G6; upstream's literal as recalled
 *           (tests/golden/similarity_lut_recalled.json) is byte-identical.  Default.
 * LINEAR  : response = 4 - |ori-j|, sum 528 — round 1's default, kept selectable. */
enum { LMB200_SIMLUT_CIRCULAR = 0, LMB200_SIMLUT_LINEAR = 1 };

typedef struct {
  const void* data;
  int rows, cols;
  int type;                 /* LMB200_8UC3 / LMB200_16UC1 / LMB200_8UC1 */
  size_t step;              /* bytes per row, 0 = packed */
} lmb200_image;

typedef struct { int x, y, label; } lmb200_feature;            /* cv::linemod::Feature */
typedef struct {                                               /* cv::linemod::Template */
  int width, height, pyramid_level, num_features;
  const lmb200_feature* features;
} lmb200_template;
typedef struct {                                               /* cv::linemod::Match */
  int x, y;
  float similarity;
  int class_index;          /* index into lmb200_class_id(); the C++ wrapper rehydrates the string */
  int template_id;
} lmb200_match_rec;

/* ---- lifetime ---------------------------------------------------------------------- */
void lmb200_default_modality(int type, lmb200_modality* out);
void lmb200_default_config(lmb200_config* out, int with_depth); /* getDefaultLINEMOD / getDefaultLINE: T={5,8} */
int lmb200_create(const lmb200_config* cfg, lmb200_handle* out);
void lmb200_destroy(lmb200_handle h);
const char* lmb200_last_error(lmb200_handle h);                /* h may be NULL: last create error */
const char* lmb200_version(void);

/* ---- introspection (HighLevelLinemod.cpp:55,:60,:65,:115,:119,:184) ----------------- */
int lmb200_num_modalities(lmb200_handle h);
const char* lmb200_modality_name(lmb200_handle h, int i);      /* "ColorGradient" / "DepthNormal" */
int lmb200_pyramid_levels(lmb200_handle h);
int lmb200_get_T(lmb200_handle h, int level);
int lmb200_num_classes(lmb200_handle h);
const char* lmb200_class_id(lmb200_handle h, int class_index); /* classIds(), std::map key order */
int lmb200_num_templates(lmb200_handle h, const char* class_id /* NULL = all classes */);
/* getTemplates(class_id, template_id)[pyramid_index]; pyramid_index = level*num_modalities+modality.
 * out->features points into handle-owned memory, valid until the template set changes. */
int lmb200_get_template(lmb200_handle h, const char* class_id, int template_id, int pyramid_index,
                        lmb200_template* out);

/* ---- template set ------------------------------------------------------------------ */
/* Replaces detector->addTemplate(sources, class_id, object_mask, &bb)  (HighLevelLinemod.cpp:93).
 * Quantisation runs on the GPU, feature selection on the host.  *template_id = -1 when extraction
 * fails (too few features), exactly what the reference tests for (HighLevelLinemod.cpp:97).
 * bb4 (nullable) = x, y, width, height of the cropped bounding box. */
int lmb200_add_template(lmb200_handle h, const char* class_id, const lmb200_image* sources, int n_sources,
                        const lmb200_image* object_mask /* nullable */, int* bb4, int* template_id);
/* Bulk form of lmb200_add_template for the reference's generateTemplates loop (HighLevelLinemod.cpp:68-110, one
 * addTemplate per rendered view): n_views views of one size, sources[view * n_sources + modality], masks[view]
 * (array nullable, entries with data == NULL mean "no mask").  The views are quantised in batches on the GPU while
 * host threads select features of the previous batch.  Result = n_views successive lmb200_add_template calls:
 * template_ids[view] = id or -1, bb4 (nullable) = 4 ints per view (untouched for failed views). */
int lmb200_add_templates(lmb200_handle h, const char* class_id, int n_views, const lmb200_image* sources, int n_sources,
                         const lmb200_image* masks, int* bb4, int* template_ids);
/* Detector::addSyntheticTemplate: n = pyramid_levels*num_modalities templates, index level*M+modality. */
int lmb200_add_synthetic_template(lmb200_handle h, const char* class_id, const lmb200_template* templates,
                                  int n, int* template_id);
int lmb200_clear_templates(lmb200_handle h);
/* Names of SURVEY.md 8b for the same entry points: lmb200_add_template_images == lmb200_add_template,
 * lmb200_add_template_pyramid == lmb200_add_synthetic_template.  lmb200_upload_templates packs the template set and
 * copies it to the device now (tables + plan for the last frame size) instead of at the next match; optional, the
 * match calls do it whenever the set has changed. */
int lmb200_add_template_images(lmb200_handle h, const char* class_id, const lmb200_image* sources, int n_sources,
                               const lmb200_image* object_mask /* nullable */, int* bb4, int* template_id);
int lmb200_add_template_pyramid(lmb200_handle h, const char* class_id, const lmb200_template* templates, int n,
                                int* template_id);
int lmb200_upload_templates(lmb200_handle h);

/* ---- persistence (OpenCV FileStorage YAML 1.0, optional .gz) ------------------------ */
/* lmb200_write: what HighLevelLineMOD::writeLinemod puts in linemod_templates.yml.gz
 *   (HighLevelLinemod.cpp:256-270): Detector::write at the root + "classes": [ {writeClass}, ... ].
 * lmb200_read : the mirror (HighLevelLinemod.cpp:288-300): creates a detector from the file's
 *   pyramid_levels/T/modalities and readClass()es every entry of "classes". */
int lmb200_write(lmb200_handle h, const char* path);
int lmb200_read(const char* path, int device, lmb200_handle* out);
/* Detector::writeClasses / readClasses with format("templates_%s.yml.gz", class_id). */
int lmb200_write_classes(lmb200_handle h, const char* format);
int lmb200_read_classes(lmb200_handle h, const char* const* class_ids, int n, const char* format);
/* Detector::writeClass(class_id, fs) / readClass(fn, class_id_override) on a file of their own: one
 * {class_id, modalities, pyramid_levels, template_pyramids} map at the root.  readClass checks modality names and
 * pyramid_levels against the detector and refuses a class that is already present (upstream CV_Asserts); with a
 * non-empty override the entry is inserted under that name (an existing entry wins, like std::map::insert). */
int lmb200_write_class(lmb200_handle h, const char* class_id, const char* path);
int lmb200_read_class(lmb200_handle h, const char* path, const char* class_id_override /* nullable */);

/* Fast binary cache of the whole detector (config + every template pyramid, 5 bytes per feature): parsing the
 * YAML of a 20 000-template set takes seconds, the cache loads in milliseconds.  Not an interchange format. */
int lmb200_write_cache(lmb200_handle h, const char* path);
int lmb200_read_cache(const char* path, int device, lmb200_handle* out);

/* The reference's pose sidecar linemod_tempPosFile.bin (HighLevelLinemod.cpp:272-284, :302-317): u32 class count,
 * then per class u64 n + n raw `struct HighLevelLineMOD::Template` records (HighLevelLinemod.h:130-148:
 * glm::vec3, glm::qua<float>, cv::Rect, uint16_t; 48 bytes with padding).  It maps template_id -> pose. */
typedef struct {
  float translation[3];
  float quaternion[4];      /* glm::qua<float> storage order as written by the reference build */
  int bb[4];                /* cv::Rect x, y, width, height */
  uint16_t median_depth;
  uint16_t pad;
} lmb200_template_pose;     /* sizeof == 48 == sizeof(HighLevelLineMOD::Template) */
int lmb200_read_pose_sidecar(const char* path, int class_index, lmb200_template_pose* out, size_t cap, size_t* n_out);
int lmb200_write_pose_sidecar(const char* path, const lmb200_template_pose* const* per_class, const size_t* counts, int n_classes);

/* ---- matching ------------------------------------------------------------------------ */
/* Replaces detector->match(sources, threshold, matches, class_ids, quantized_images, masks)
 * (HighLevelLinemod.cpp:152).  Result order is upstream's: generation order (class, template_id,
 * coarse raster) -> std::sort -> std::unique.  class_ids NULL/0 = all classes in map order; unknown
 * ids are skipped silently like upstream.  quantized_out (nullable): pyramid_levels*num_modalities
 * caller-allocated LMB200_8UC1 images, index level*M+modality, sized rows>>level x cols>>level.
 * masks (nullable): num_modalities LMB200_8UC1 images (data may be NULL per entry = no mask).
 * If cap < matches: fills cap records, sets *n_out to the full count, returns LMB200_E_TRUNCATED. */
int lmb200_match(lmb200_handle h, const lmb200_image* sources, int n_sources, float threshold,
                 const char* const* class_ids, int n_class_ids,
                 lmb200_match_rec* out, size_t cap, size_t* n_out,
                 lmb200_image* quantized_out, const lmb200_image* masks);

/* Streams n_frames frames (frames[f*n_sources + m]) through the same path: chunked H2D copies,
 * kernels and D2H of the match lists overlap on one copy and three compute streams.  Per-frame results are written
 * at out[offsets[f] .. offsets[f+1]) (offsets has n_frames+1 entries).  Use pinned host memory
 * (lmb200_host_alloc) for the frames to get asynchronous copies. */
int lmb200_match_batch(lmb200_handle h, const lmb200_image* frames, int n_frames, int n_sources, float threshold,
                       const char* const* class_ids, int n_class_ids,
                       lmb200_match_rec* out, size_t cap, size_t* offsets);

/* The same, split in two so consecutive batches pipeline into each other: submit enqueues every copy and kernel of
 * the batch without blocking and returns a ticket (at most two in flight); collect waits for it and delivers the
 * match lists.  `frames` (the descriptor array and the pixels) must stay valid until the ticket is collected. */
int lmb200_match_batch_submit(lmb200_handle h, const lmb200_image* frames, int n_frames, int n_sources, float threshold,
                              const char* const* class_ids, int n_class_ids, int* ticket);
int lmb200_match_batch_collect(lmb200_handle h, int ticket, lmb200_match_rec* out, size_t cap, size_t* offsets);

/* Device-resident variant used to time the path without PCIe: upload once, match many times.
 * lmb200_match_resident enqueues the whole device pipeline for frames [first, first+count) and
 * returns without synchronising; lmb200_fetch_resident synchronises, copies the packed match lists
 * back and applies the host sort/unique. */
int lmb200_upload_frames(lmb200_handle h, const lmb200_image* frames, int n_frames, int n_sources, int first_slot);
int lmb200_match_resident(lmb200_handle h, int first_slot, int count, float threshold,
                          const char* const* class_ids, int n_class_ids);
int lmb200_fetch_resident(lmb200_handle h, int first_slot, int count,
                          lmb200_match_rec* out, size_t cap, size_t* offsets);
int lmb200_synchronize(lmb200_handle h);
void* lmb200_stream(lmb200_handle h);                           /* cudaStream_t of the compute lane */
/* CUDA-event stopwatch on the compute stream: which = 0 records "start", 1 records "stop". */
int lmb200_timer_record(lmb200_handle h, int which);
int lmb200_timer_elapsed_ms(lmb200_handle h, float* ms);        /* synchronises on "stop" */

int lmb200_host_alloc(size_t bytes, void** out);                /* cudaHostAlloc (pinned) */
int lmb200_host_free(void* p);

/* ---- post-match checks (the per-match image operations right behind Detector::match) ------------------
 * Colour check of HighLevelLineMOD::detectTemplate / templateMask / colorCheck (src/HighLevelLinemod.cpp:159-161, :113-135,
 * :424-434) on the GPU, for the frame resident in `slot` (lmb200_match leaves its frame in slot 0): cvtColor(BGR2HSV) +
 * inRange(lower, upper) once per call, then per match the convex hull of the level-0 features of every modality shifted
 * to the match, filled like cv::fillPoly:  total[i] = countNonZero(mask), inside[i] = countNonZero(hue & mask).
 * colorCheck is then `(float)(inside * 100 / total) > percentToPassCheck` (integer division).  A match whose hull leaves
 * the image (cv::fillPoly would clip it) or whose template is unknown gets inside = total = -1. */
int lmb200_postmatch_color(lmb200_handle h, int slot, const uint8_t* lower_hsv, const uint8_t* upper_hsv,
                           const lmb200_match_rec* matches, size_t n, int* inside, int* total);
/* medianMat of HighLevelLineMOD::depthCheck (src/HighLevelLinemod.cpp:336-349, :436-457), host code: depth values 0 and 1
 * count as 65535 (threshold at 1, inverted, saturating add), the crop bb = {x, y, width, height} (must lie inside the image, as cv::Mat::operator()(Rect) demands) is
 * flattened row by row, std::nth_element puts the (n/4)-th value in place and element n / median_position is returned — with
 * the reference's median_position = 5 that is an element of the unordered lower part, whose identity depends on libstdc++'s
 * introselect: the function makes the same call on the same sequence, so it returns the reference's value.
 * depthCheck itself is then  abs((int)median - (int)template_median_depth - depth_offset) < step_size. */
int lmb200_postmatch_median_depth(const uint16_t* depth, int rows, int cols, size_t step_bytes, const int* bb4, int median_position,
                                  uint16_t* median);
/* groupSimilarMatches + discardSmallMatchGroups (src/HighLevelLinemod.cpp:206-253), host code: group_of_match[i] = index
 * of the surviving group the match belongs to (groups in order of their first member) or -1 when its group was discarded
 * (size * 100 / biggest <= discard_group_ratio). */
int lmb200_group_matches(const lmb200_match_rec* matches, size_t n, float radius_threshold, float discard_group_ratio,
                         int* group_of_match, int* n_groups);

/* ---- tables --------------------------------------------------------------------------- */
/* 256-byte SIMILARITY_LUT (layout of upstream's table: [32*ori + 16*half + nibble]); entries <= 4.
 * Default = lmb200_config.similarity_lut (circular distance, see LMB200_SIMLUT_*). */
int lmb200_set_similarity_lut(lmb200_handle h, const uint8_t* lut256);
int lmb200_get_similarity_lut(lmb200_handle h, uint8_t* lut256);
/* 8000-byte NORMAL_LUT[20][20][20] (index [v3][v2][v1]); entries must be 0 or one-hot.
 * Default = documented stand-in generator (upstream normal_lut.i is not redistributable here). */
int lmb200_set_normal_lut(lmb200_handle h, const uint8_t* lut8000);
int lmb200_get_normal_lut(lmb200_handle h, uint8_t* lut8000);
/* Loads NORMAL_LUT from a file: either upstream's `normal_lut.i` (C initialiser text: the first 8000 integer
 * literals after the first '{' are taken in order [v3][v2][v1]) or a raw 8000-byte binary. */
int lmb200_load_normal_lut(lmb200_handle h, const char* path);
/* 1 while the detector still uses the built-in stand-in NORMAL_LUT (DepthNormal labels then differ from
 * cv::linemod's by construction), 0 once lmb200_set_normal_lut / lmb200_load_normal_lut supplied a table. */
int lmb200_normal_lut_is_standin(lmb200_handle h);
/* Non-fatal diagnostics accumulated by the handle (e.g. "DepthNormal modality is running on the stand-in
 * NORMAL_LUT"); empty string when there are none.  The stand-in warning is also printed once to stderr unless
 * the environment variable LMB200_QUIET is set. */
const char* lmb200_warnings(lmb200_handle h);

/* ---- multi-GPU (one process per GPU) -------------------------------------------------- */
/* Template sharding: this handle scores only shard `rank` of `world`.  Shards are interleaved over the
 * generation-ordered selection list (positions rank, rank+world, ...), which spreads the templates of one object over
 * all ranks; lmb200_fetch_resident_allgather restores generation order by selection position.  world=1 = full set. */
int lmb200_set_template_shard(lmb200_handle h, int rank, int world);
/* NCCL plumbing (libnccl.so.2 is dlopen'ed on first use). unique_id is 128 bytes. */
int lmb200_comm_unique_id(uint8_t* unique_id128);
int lmb200_comm_init(lmb200_handle h, const uint8_t* unique_id128, int rank, int world);
int lmb200_comm_destroy(lmb200_handle h);
/* Collective fetch of a template-sharded step: every rank returns the identical, complete per-frame lists (all ranks
 * must call it for the same slot ranges in the same order).  Per step a kernel packs the per-frame match buffers of the
 * rank's shard into one compact buffer, ONE ncclAllGather over NVLink shares them, and the rank-/position-ordered
 * concatenation (= the reference's generation order) goes through the same std::sort / std::unique epilogue as the 1-GPU
 * path.  After lmb200_match_resident_sharded the gather has already been enqueued behind the step's kernels and the
 * epilogue has run on the handle's epilogue thread: the call then only waits for that thread and copies the lists out.
 * After plain lmb200_match_resident (template shard set, every rank holds the same frames) it does the same work
 * synchronously.  Overflowing stores are grown and the step's template side is redone, as in the 1-GPU fetch. */
int lmb200_fetch_resident_allgather(lmb200_handle h, int first_slot, int count,
                                    lmb200_match_rec* out, size_t cap, size_t* offsets);
/* The same template-sharded step with the expensive half of the frame side sharded as well: rank r quantises only the
 * frame block [first + r*n, first + (r+1)*n) (n = count / world) — so only those frames have to be uploaded on rank r —
 * the quantized maps are all-gathered in place over NVLink (one NCCL group), and every rank spreads all frames and
 * scores its template shard.  Needs lmb200_comm_init + lmb200_set_template_shard with the same rank/world; count not a
 * multiple of world falls back to lmb200_match_resident (replicated frame side).  The call only enqueues: quantisers
 * and map all-gather on a lane of their own (overlapping the previous step's template side), then spread, matching, the
 * match gather and its copy to pinned memory on the compute lane; a per-handle epilogue thread merges the gathered
 * lists when they land.  Follow with lmb200_fetch_resident_allgather; several steps on different slot ranges may be
 * in flight (up to 4), fetched in submission order.  Do not add templates while steps are in flight. */
int lmb200_match_resident_sharded(lmb200_handle h, int first_slot, int count, float threshold,
                                  const char* const* class_ids, int n_class_ids);
/* Pure host helper (no GPU): merge per-rank, generation-ordered partial lists exactly as above.
 * parts[r] has counts[r] records.  Used by the gloo CPU tests and by callers with their own transport. */
int lmb200_merge_matches(const lmb200_match_rec* const* parts, const size_t* counts, int world,
                         lmb200_match_rec* out, size_t cap, size_t* n_out);
/* Cost-balanced contiguous split of n templates with costs[] into `world` shards: begin[world+1]. */
int lmb200_shard_plan(const double* costs, int n, int world, int* begin);

/* ---- measurement / debug --------------------------------------------------------------- */
enum {
  LMB200_K_UPLOAD = 0, LMB200_K_PYRDOWN, LMB200_K_CG_QUANTIZE, LMB200_K_DN_QUANTIZE, LMB200_K_MEDIAN,
  LMB200_K_DECIMATE, LMB200_K_LINEARIZE, LMB200_K_SIM_COARSE, LMB200_K_SIM_LOCAL, LMB200_K_PACK,
  LMB200_K_COMM,           /* NCCL all-gather of the quantized maps (template-sharded step) */
  LMB200_K_EPILOGUE,       /* device std::sort + std::unique of the template-sharded step */
  LMB200_K_COUNT
};
typedef struct {
  double ms[LMB200_K_COUNT];          /* accumulated CUDA-event time per kernel family */
  long long launches[LMB200_K_COUNT]; /* launches per family */
  long long bytes_coarse;             /* algorithmic bytes gathered by similarity (sum nf*P) */
  long long bytes_local;              /* algorithmic bytes gathered by similarityLocal (sum nf*256) */
  long long frames;
  long long candidates, matches;
  long long chunks_coarse;            /* 16-byte chunk loads the coarse kernel actually issued (after its early exit) */
} lmb200_profile;
/* Test hook for the device epilogue's sort (csrc/sort_emul.h restates libstdc++'s std::sort so that the device returns the
 * reference's sequence, ties included): runs the restatement (host build) and the real thing on the same records.
 * mode 0: std::sort, 1: std::partial_sort over the whole range (the heap-sort fallback), 2: std::sort on an input built by
 * McIlroy's quicksort adversary (template ids are overwritten; forces the depth limit).  Both outputs hold n records. */
int lmb200_debug_sort_check(const lmb200_match_rec* in, size_t n, int mode, lmb200_match_rec* out_emulated, lmb200_match_rec* out_std);

/* Test hook for the device epilogue of the template-sharded step (csrc/kernels_epilogue.cu) on ONE GPU: `gathered` is what
 * the match all-gather leaves in device memory — per rank [2*frames header records][gcap match records], a record being
 * four 32-bit words {global template index, x, y, similarity as float bits}; header 2f = {count, flags, offset into the
 * rank's record area, 0}, header 2f+1 = counters.  pos_of_g (nullable: rank-ordered concatenation) maps a global template
 * index to its selection position.  Returns the finished records (frame f at hdr[8f+1], hdr[8f+0] of them) and the
 * per-frame header words {n_final, offset, flags, n_in}, {counters of `rank`}. */
int lmb200_debug_shard_epilogue(const int32_t* gathered, int world, int rank, int frames, int gcap, const int32_t* pos_of_g,
                                const int32_t* g_class, const int32_t* g_tid, int ntpl, lmb200_match_rec* out, size_t out_cap,
                                int32_t* hdr);

/* The HOST merge of a template-sharded step on caller-made gathered buffers (same layout as lmb200_debug_shard_epilogue;
 * no device needed): generation order restored, record -> match, std::sort, std::unique.  offsets has frames + 1 entries. */
int lmb200_debug_merge_gathered(const int32_t* gathered, int world, int frames, int gcap, const int32_t* pos_of_g,
                                const int32_t* g_class, const int32_t* g_tid, int ntpl, lmb200_match_rec* out, size_t cap,
                                size_t* offsets);

/* Measurement knobs.  "early_exit" (default 1): 0 switches off the coarse kernel's exact early exit (results are
 * identical either way; the bench reports both so the workload dependence of the exit is visible).
 * "upload_async" (default 0): 1 makes lmb200_upload_frames return without synchronising (pinned host frames that stay
 * valid until the results of the next match on those slots have been fetched).
 * "cuda_graph" (default 1; environment LMB200_NO_GRAPH=1 starts with 0): lmb200_match replays the kernel sequence of one
 * frame as a CUDA graph (re-captured when plan, templates, selection, stores or threshold change).
 * "shard_overlap" (default 1): lmb200_match_resident_sharded runs its quantisers and the all-gather of the quantized maps
 * (1), or those and spread + linearize (2), on a high-priority lane of their own, overlapping the template side of the
 * previous step (0: everything on the compute lane).
 * "resident_overlap" (default 1): lmb200_match_resident runs the frame side of a step (>= 8 frames) on the high-priority
 * frame lane and the template side on the compute lane; consecutive calls on different slot ranges overlap.
 * "host_threads" (default 8): threads the host epilogue (std::sort / std::unique per frame) of one fetch / collect may use;
 * with one process per GPU set it to (host cores / processes).
 * "shard_device_epilogue" (default 1): the std::sort + std::unique of a template-sharded step run on the device
 * (csrc/kernels_epilogue.cu, sequence-identical to libstdc++'s); 0: on the handle's host epilogue thread. */
int lmb200_set_option(lmb200_handle h, const char* name, int value);
int lmb200_set_profiling(lmb200_handle h, int enabled);        /* CUDA events around every launch */
int lmb200_get_profile(lmb200_handle h, lmb200_profile* out, int reset);

/* Measured roofs of the device the library runs on (csrc/microbench.cu): achieved GB/s of
 *   L2_READ : 128-bit loads streaming an L2-resident buffer with L1 bypassed (the similarity kernels' L2 roof),
 *   L1_READ : 128-bit loads re-reading an L1-resident window per CTA (their L1 data-pipe roof),
 *   HBM_READ: the same stream over 2 GiB,
 *   H2D     : cudaMemcpyAsync from pinned host memory (roof of the streamed end-to-end path; call it on every rank
 *             at once to get the shared-host figure).
 * bytes = 0 picks the default size of each kind; the result is the best of `iters` timed launches (H2D: their mean). */
enum { LMB200_MB_L2_READ = 0, LMB200_MB_L1_READ = 1, LMB200_MB_HBM_READ = 2, LMB200_MB_H2D = 3 };
int lmb200_microbench(int kind, size_t bytes, int iters, double* gbps);

enum { LMB200_DBG_QUANTIZED = 0, LMB200_DBG_LINMEM = 1, LMB200_DBG_COARSE = 2, LMB200_DBG_UNSORTED = 3,
       LMB200_DBG_MAGNITUDE = 4, LMB200_DBG_DN_INDICES = 5, LMB200_DBG_SIMILARITY = 6 };
/* Copies an intermediate of the LAST match on `slot` to host: QUANTIZED/LINMEM take index=level*M+modality;
 * COARSE/UNSORTED return lmb200_match_rec arrays (generation order).  SIMILARITY: index = global template index
 * (classes in classIds() order, template_id ascending); returns the u16 [H*W] coarse-level map upstream's
 * similarity() + addSimilarities() produce for that template, computed by the production kernel with its early
 * exit disabled.  *n_bytes in: capacity, out: size. */
int lmb200_debug_fetch(lmb200_handle h, int kind, int slot, int index, void* dst, size_t* n_bytes);

/* ---- headless view synthesis (host code) --------------------------------------------
 * Replaces the SDL + OpenGL offscreen passes the reference renders template views and benchmark images with
 * (OpenGLRender::renderDepthToFrontBuff / renderColorToFrontBuff + getDepthImgFromBuff / getColorImgFromBuff,
 * src/OpenglRender.cpp:33-141, shader/depth.fs): a z-buffer rasteriser, one view per host thread.
 * Camera space: looking down -z, +y up; u = cx + fx*x/z, v = cy - fy*y/z (the image after the reference's vertical
 * flip).  Depth: nearest surface in millimetres (u16, 0 = background).  Colour: 255 on the model, 0 elsewhere
 * (the reference draws models without vertex colours white and thresholds the image to binary at once). */
typedef struct {
  const double* vertices;  /* [n_vertices][3] model coordinates (mm) */
  int n_vertices;
  const int* triangles;    /* [n_triangles][3] vertex indices */
  int n_triangles;
} lmb200_mesh;
typedef struct {
  int width, height;
  double fx, fy, cx, cy;   /* the reference uses fy for both axes and the image centre (OpenglRender.cpp:3-12) */
  double near_mm, far_mm;  /* 100 / 10000 in the reference (OpenglRender.cpp:10-11); far only bounds the u16 range */
} lmb200_camera;
/* n_views cameras at eyes[i] looking at the origin with +Y up (glm::lookAt(eye, 0, up), OpenglRender.cpp:334-345).
 * depth_out [n_views][height][width] u16 and/or colour_out [n_views][height][width][3] u8 (either may be NULL).
 * threads <= 0: all host threads. */
int lmb200_render_lookat(const lmb200_mesh* mesh, const lmb200_camera* cam, const double* eyes, int n_views,
                         uint16_t* depth_out, uint8_t* colour_out, int threads);
/* The same with explicit model-view transforms x_cam = R*x + t (rotations [n][9] row-major, translations [n][3]):
 * the overloads the benchmark uses (OpenglRender.cpp:69-94,:116-141), which build `view` from in_rotMat and put
 * (in_traVec.x, -in_traVec.y, -in_traVec.z) into its translation column: pass R = the upper-left 3x3 of in_rotMat
 * (glm is column-major: R[3*i+j] = in_rotMat[j][i]) and t = (tx, -ty, -tz). */
int lmb200_render_pose(const lmb200_mesh* mesh, const lmb200_camera* cam, const double* rotations,
                       const double* translations, int n_views, uint16_t* depth_out, uint8_t* colour_out, int threads);
/* Benchmark::calculateErrorHodan (src/Benchmark.cpp:18-38 with calculateVisibilityMasks :133-154): the visibility-masked
 * depth-difference error of an estimated pose, from the input depth image and the model rendered at the ground-truth and
 * at the estimated pose (u16 millimetres, 0 = background).  The reference's defaults: visibility_threshold 15,
 * error_threshold 20 (include/Benchmark.h:92,:98); a pose counts as correct when error < 0.3 (Benchmark.cpp:33).
 * n_ok / n_comb (nullable) return the two pixel counts of the ratio. */
int lmb200_hodan_error(const uint16_t* input_depth, const uint16_t* gt_render, const uint16_t* est_render, int rows, int cols,
                       int visibility_threshold, int error_threshold, float* error, long long* n_ok, long long* n_comb);
/* The same including the two renders (lmb200_render_pose conventions): rotations[2][9], translations[2][3] =
 * ground truth, estimate. */
int lmb200_hodan_error_poses(const lmb200_mesh* mesh, const lmb200_camera* cam, const double* rotations, const double* translations,
                             const uint16_t* input_depth, int visibility_threshold, int error_threshold, float* error);
/* ASCII PLY loader (the .ply models of the reference; polygons become triangle fans).  Free both arrays with lmb200_free. */
int lmb200_load_ply(const char* path, double** vertices, int* n_vertices, int** triangles, int* n_triangles);
void lmb200_free(void* p);

#ifdef __cplusplus
}
#endif
#endif /* LMB200_H */
