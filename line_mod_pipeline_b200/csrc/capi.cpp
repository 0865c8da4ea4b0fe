// capi.cpp — the extern "C" surface of include/lmb200.h that is pure host logic: lifetime,
// introspection, template store, tables, persistence wrappers.  (Compute entry points live in
// detector.cu, multi-GPU plumbing in comm.cpp.)
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "detector.h"

using namespace lmh;

namespace lmh {
int add_template_gpu(lmb200_detector* h, const char* class_id, const lmb200_image* sources, int n_sources,
                     const lmb200_image* object_mask, int* bb4, int* template_id);

// SIMILARITY_LUT: max over set bits j of max(0, 4 - dist(ori, j)).
// circular (default): dist = min(|ori-j|, 8-|ori-j|) — SURVEY.md 8c G6 (sum 628); upstream's literal table as recalled
// in round 2 is byte-identical (tests/golden/similarity_lut_recalled.json).  linear: dist = |ori-j| (sum 528).
void similarity_lut_variant(int linear, uint8_t* out) {
  for (int ori = 0; ori < 8; ++ori)
    for (int half = 0; half < 2; ++half)
      for (int nib = 0; nib < 16; ++nib) {
        int best = 0;
        for (int b = 0; b < 4; ++b)
          if (nib & (1 << b)) {
            int d = ori - (b + 4 * half);
            if (d < 0) d = -d;
            if (!linear && 8 - d < d) d = 8 - d;
            if (4 - d > best) best = 4 - d;
          }
        out[32 * ori + 16 * half + nib] = (uint8_t)best;
      }
}
void default_similarity_lut(uint8_t* out) { similarity_lut_variant(0, out); }

// Stand-in for upstream's NORMAL_LUT[20][20][20] (normal_lut.i is not available offline).
// Cell (v3,v2,v1) -> direction (x,y) = (2*v1-19, 2*v2-19); label = the 45-degree sector of atan2(y,x)
// centred on multiples of 45 degrees, decided in exact integer arithmetic
// (|y| < |x| tan 22.5deg  <=>  (|x|+|y|)^2 < 2 x^2).  One-hot byte 1<<sector.
void default_normal_lut(uint8_t* out) {
  for (int v3 = 0; v3 < 20; ++v3)
    for (int v2 = 0; v2 < 20; ++v2)
      for (int v1 = 0; v1 < 20; ++v1) {
        int x = 2 * v1 - 19, y = 2 * v2 - 19;
        int ax = x < 0 ? -x : x, ay = y < 0 ? -y : y;
        int s = (ax + ay) * (ax + ay), bin;
        if (s < 2 * ax * ax) bin = x > 0 ? 0 : 4;
        else if (s < 2 * ay * ay) bin = y > 0 ? 2 : 6;
        else bin = x > 0 ? (y > 0 ? 1 : 7) : (y > 0 ? 3 : 5);
        out[(v3 * 20 + v2) * 20 + v1] = (uint8_t)(1 << bin);
      }
}
}  // namespace lmh

static std::string g_create_error;
namespace lmh { void set_create_error(const std::string& msg) { g_create_error = msg; } }

extern "C" {

const char* lmb200_version(void) { return "lmb200 0.1 (sm_100a)"; }

void lmb200_default_modality(int type, lmb200_modality* out) {
  std::memset(out, 0, sizeof(*out));
  out->type = type;
  out->weak_threshold = 10.0f; out->num_features = 63; out->strong_threshold = 55.0f;
  out->distance_threshold = 2000; out->difference_threshold = 50; out->extract_threshold = 2;
}

void lmb200_default_config(lmb200_config* out, int with_depth) {
  std::memset(out, 0, sizeof(*out));
  out->num_modalities = with_depth ? 2 : 1;
  lmb200_default_modality(LMB200_COLOR_GRADIENT, &out->modalities[0]);
  if (with_depth) lmb200_default_modality(LMB200_DEPTH_NORMAL, &out->modalities[1]);
  out->pyramid_levels = 2;
  out->T[0] = 5; out->T[1] = 8;
  out->device = -1;
}

int lmb200_create(const lmb200_config* cfg, lmb200_handle* out) {
  if (!cfg || !out) return LMB200_E_INVALID;
  if (cfg->num_modalities < 1 || cfg->num_modalities > LMB200_MAX_MODALITIES || cfg->pyramid_levels < 1 ||
      cfg->pyramid_levels > LMB200_MAX_LEVELS) {
    g_create_error = "num_modalities must be 1..4 and pyramid_levels 1..8";
    return LMB200_E_INVALID;
  }
  for (int m = 0; m < cfg->num_modalities; ++m)
    if (cfg->modalities[m].type != LMB200_COLOR_GRADIENT && cfg->modalities[m].type != LMB200_DEPTH_NORMAL) {
      g_create_error = "unknown modality type";
      return LMB200_E_INVALID;
    }
  for (int l = 0; l < cfg->pyramid_levels; ++l)
    if (cfg->T[l] < 1 || cfg->T[l] > 32) { g_create_error = "T must be in 1..32"; return LMB200_E_INVALID; }
  if (cfg->similarity_lut != LMB200_SIMLUT_CIRCULAR && cfg->similarity_lut != LMB200_SIMLUT_LINEAR) {
    g_create_error = "similarity_lut must be LMB200_SIMLUT_CIRCULAR or LMB200_SIMLUT_LINEAR";
    return LMB200_E_INVALID;
  }
  lmb200_detector* h = new lmb200_detector();
  h->cfg = *cfg;
  std::memset(&h->prof, 0, sizeof(h->prof));
  similarity_lut_variant(cfg->similarity_lut == LMB200_SIMLUT_LINEAR, h->sim_lut);
  default_normal_lut(h->normal_lut);
  h->normal_lut_standin = true;
  { const char* e = std::getenv("LMB200_NO_GRAPH"); h->use_graph = !(e && *e && *e != '0'); }
  *out = h;
  return LMB200_OK;
}

void lmb200_destroy(lmb200_handle h) {
  if (!h) return;
  if (h->device_ready) {
    cudaSetDevice(h->device);
    shard_jobs_stop(h);
    cudaDeviceSynchronize();
    comm_destroy(h);
    for (auto& lb : h->levels)
      for (int m = 0; m < LMB200_MAX_MODALITIES; ++m) { lb.bgr[m].release(); lb.q[m].release(); lb.mask[m].release(); lb.lm[m].release(); lb.lmn[m].release(); }
    for (int m = 0; m < LMB200_MAX_MODALITIES; ++m) { h->d_depth[m].release(); h->d_dnraw[m].release(); }
    for (int l = 0; l < LMB200_MAX_LEVELS; ++l) { h->d_hdr[l].release(); h->d_feat[l].release(); h->d_offs[l].release(); }
    lmh::DevBuf* bufs[] = {&h->d_frames, &h->d_table, &h->d_normal_lut, &h->d_sel, &h->d_mag, &h->d_dnidx, &h->d_cand, &h->d_ctr,
                           &h->d_tpl_start, &h->d_tpl_cnt, &h->d_tpl_alive, &h->d_out, &h->d_gather_send, &h->d_gather_recv, &h->d_resp_sum, &h->d_hue_bits, &h->d_gclass, &h->d_gtid, &h->d_posg};
    for (auto* b : bufs) b->release();
    if (h->h_ctr) cudaFreeHost(h->h_ctr);
    if (h->h_out) cudaFreeHost(h->h_out);
    if (h->h_gather) cudaFreeHost(h->h_gather);
    for (auto& g : h->jobs) { g.send.release(); g.recv.release(); if (g.host) cudaFreeHost(g.host); if (g.ev) cudaEventDestroy(g.ev); if (g.ev_q) cudaEventDestroy(g.ev_q); if (g.ev_g) cudaEventDestroy(g.ev_g);
      g.fin_dev.release(); if (g.fin_host) cudaFreeHost(g.fin_host); if (g.hdr_host) cudaFreeHost(g.hdr_host); }
    for (auto& tk : h->tickets) {
      if (tk.b_ctr) cudaFreeHost(tk.b_ctr);
      if (tk.b_out) cudaFreeHost(tk.b_out);
      for (auto e : tk.events) cudaEventDestroy(e);
    }
    for (auto e : h->group_done) cudaEventDestroy(e);
    for (auto& mk : h->resident_marks) if (mk.ev) cudaEventDestroy(mk.ev);
    if (h->upload_ev) cudaEventDestroy(h->upload_ev);
    if (h->match_graph) cudaGraphExecDestroy(h->match_graph);
    for (auto e : h->fork_ev) if (e) cudaEventDestroy(e);
    if (h->fs_done) cudaEventDestroy(h->fs_done);
    for (auto& r : h->prof_pending) { cudaEventDestroy(r.a); cudaEventDestroy(r.b); }
    for (auto e : h->event_pool) cudaEventDestroy(e);
    for (int i = 0; i < LMB200_LANES; ++i) {
      if (h->lanes[i].stream) cudaStreamDestroy(h->lanes[i].stream);
    }
  }
  delete h;
}

const char* lmb200_last_error(lmb200_handle h) { return h ? h->err.c_str() : g_create_error.c_str(); }

int lmb200_num_modalities(lmb200_handle h) { return h ? h->cfg.num_modalities : LMB200_E_INVALID; }
const char* lmb200_modality_name(lmb200_handle h, int i) {
  if (!h || i < 0 || i >= h->cfg.num_modalities) return nullptr;
  return h->cfg.modalities[i].type == LMB200_COLOR_GRADIENT ? "ColorGradient" : "DepthNormal";
}
int lmb200_pyramid_levels(lmb200_handle h) { return h ? h->cfg.pyramid_levels : LMB200_E_INVALID; }
int lmb200_get_T(lmb200_handle h, int level) {
  if (!h || level < 0 || level >= h->cfg.pyramid_levels) return LMB200_E_INVALID;
  return h->cfg.T[level];
}
int lmb200_num_classes(lmb200_handle h) { return h ? (int)h->classes.size() : LMB200_E_INVALID; }
const char* lmb200_class_id(lmb200_handle h, int class_index) {
  if (!h || class_index < 0 || class_index >= (int)h->classes.size()) return nullptr;
  auto it = h->classes.begin();
  std::advance(it, class_index);
  return it->first.c_str();
}
int lmb200_num_templates(lmb200_handle h, const char* class_id) {
  if (!h) return LMB200_E_INVALID;
  if (!class_id) { int n = 0; for (auto& kv : h->classes) n += (int)kv.second.size(); return n; }
  auto it = h->classes.find(class_id);
  return it == h->classes.end() ? 0 : (int)it->second.size();
}
int lmb200_get_template(lmb200_handle h, const char* class_id, int template_id, int pyramid_index, lmb200_template* out) {
  if (!h || !class_id || !out) return LMB200_E_INVALID;
  auto it = h->classes.find(class_id);
  if (it == h->classes.end() || template_id < 0 || template_id >= (int)it->second.size())
    return set_error(h, LMB200_E_CLASS, "unknown class or template id");
  const TemplatePyramid& tp = it->second[template_id];
  if (pyramid_index < 0 || pyramid_index >= (int)tp.size()) return set_error(h, LMB200_E_INVALID, "bad pyramid index");
  const Template& t = tp[pyramid_index];
  h->tmp_features.resize(t.features.size());
  for (size_t i = 0; i < t.features.size(); ++i) { h->tmp_features[i].x = t.features[i].x; h->tmp_features[i].y = t.features[i].y; h->tmp_features[i].label = t.features[i].label; }
  out->width = t.width; out->height = t.height; out->pyramid_level = t.pyramid_level;
  out->num_features = (int)t.features.size();
  out->features = h->tmp_features.data();
  return LMB200_OK;
}

int lmb200_add_template(lmb200_handle h, const char* class_id, const lmb200_image* sources, int n_sources,
                        const lmb200_image* object_mask, int* bb4, int* template_id) {
  if (!h || !class_id || !sources || !template_id) return LMB200_E_INVALID;
  return add_template_gpu(h, class_id, sources, n_sources, object_mask, bb4, template_id);
}

int lmb200_add_template_images(lmb200_handle h, const char* class_id, const lmb200_image* sources, int n_sources,
                               const lmb200_image* object_mask, int* bb4, int* template_id) {
  return lmb200_add_template(h, class_id, sources, n_sources, object_mask, bb4, template_id);
}

int lmb200_add_template_pyramid(lmb200_handle h, const char* class_id, const lmb200_template* templates, int n, int* template_id) {
  return lmb200_add_synthetic_template(h, class_id, templates, n, template_id);
}

int lmb200_upload_templates(lmb200_handle h) {
  if (!h) return LMB200_E_INVALID;
  return upload_templates_now(h);
}

int lmb200_add_templates(lmb200_handle h, const char* class_id, int n_views, const lmb200_image* sources, int n_sources,
                         const lmb200_image* masks, int* bb4, int* template_ids) {
  if (!h || !class_id || !sources || !template_ids || n_views < 0) return LMB200_E_INVALID;
  if (n_views == 0) return LMB200_OK;
  return add_templates_bulk(h, class_id, n_views, sources, n_sources, masks, bb4, template_ids);
}

int lmb200_add_synthetic_template(lmb200_handle h, const char* class_id, const lmb200_template* templates, int n, int* template_id) {
  if (!h || !class_id || !templates) return LMB200_E_INVALID;
  if (n != h->cfg.num_modalities * h->cfg.pyramid_levels)
    return set_error(h, LMB200_E_INVALID, "expected pyramid_levels*num_modalities templates");
  TemplatePyramid tp((size_t)n);
  for (int i = 0; i < n; ++i) {
    if (templates[i].num_features < 0 || templates[i].num_features > 63)
      return set_error(h, LMB200_E_FEATURES, "template has more than 63 features (upstream CV_Assert)");
    tp[i].width = templates[i].width; tp[i].height = templates[i].height; tp[i].pyramid_level = templates[i].pyramid_level;
    tp[i].features.resize(templates[i].num_features);
    for (int k = 0; k < templates[i].num_features; ++k) {
      const lmb200_feature& f = templates[i].features[k];
      if (f.label < 0 || f.label > 7) return set_error(h, LMB200_E_INVALID, "feature label must be 0..7");
      tp[i].features[k] = Feature{f.x, f.y, f.label};
    }
  }
  auto& v = h->classes[class_id];
  v.push_back(std::move(tp));
  h->templates_dirty = true;
  if (template_id) *template_id = (int)v.size() - 1;
  return LMB200_OK;
}

int lmb200_clear_templates(lmb200_handle h) {
  if (!h) return LMB200_E_INVALID;
  h->classes.clear();
  h->templates_dirty = true;
  return LMB200_OK;
}

int lmb200_write(lmb200_handle h, const char* path) { return (h && path) ? write_detector_file(h, path) : LMB200_E_INVALID; }
int lmb200_read(const char* path, int device, lmb200_handle* out) {
  if (!path || !out) return LMB200_E_INVALID;
  std::string err;
  int rc;
  try {  // no exception may cross the C ABI (a hostile file can still exhaust memory)
    rc = read_detector_file(path, device, out, err);
  } catch (const std::exception& e) {
    rc = LMB200_E_IO; err = std::string(path) + ": " + e.what();
  }
  if (rc) g_create_error = err;
  return rc;
}
int lmb200_write_class(lmb200_handle h, const char* class_id, const char* path) {
  if (!h || !class_id || !path) return LMB200_E_INVALID;
  return write_class_file(h, class_id, path);
}
int lmb200_read_class(lmb200_handle h, const char* path, const char* class_id_override) {
  if (!h || !path) return LMB200_E_INVALID;
  std::string err;
  int rc;
  try {
    rc = read_class_file(h, path, err, class_id_override);
  } catch (const std::exception& e) {
    rc = LMB200_E_IO; err = std::string(path) + ": " + e.what();
  }
  return rc ? set_error(h, rc, err) : LMB200_OK;
}
int lmb200_write_classes(lmb200_handle h, const char* format) {
  if (!h) return LMB200_E_INVALID;
  const char* fmt = format ? format : "templates_%s.yml.gz";
  for (auto& kv : h->classes) {
    char path[4096];
    std::snprintf(path, sizeof path, fmt, kv.first.c_str());
    int rc = write_class_file(h, kv.first, path);
    if (rc) return rc;
  }
  return LMB200_OK;
}
int lmb200_read_classes(lmb200_handle h, const char* const* class_ids, int n, const char* format) {
  if (!h || (n > 0 && !class_ids)) return LMB200_E_INVALID;
  const char* fmt = format ? format : "templates_%s.yml.gz";
  for (int i = 0; i < n; ++i) {
    char path[4096];
    std::snprintf(path, sizeof path, fmt, class_ids[i]);
    std::string err;
    int rc;
    try {
      rc = read_class_file(h, path, err);
    } catch (const std::exception& e) {
      rc = LMB200_E_IO; err = std::string(path) + ": " + e.what();
    }
    if (rc) return set_error(h, rc, err);
  }
  return LMB200_OK;
}

int lmb200_set_similarity_lut(lmb200_handle h, const uint8_t* lut256) {
  if (!h || !lut256) return LMB200_E_INVALID;
  for (int i = 0; i < 256; ++i)
    if (lut256[i] > 4) return set_error(h, LMB200_E_INVALID, "similarity LUT entries must be <= 4 (63 features x 4 must fit a byte)");
  std::memcpy(h->sim_lut, lut256, 256);
  h->luts_dirty = true;
  return LMB200_OK;
}
int lmb200_get_similarity_lut(lmb200_handle h, uint8_t* lut256) {
  if (!h || !lut256) return LMB200_E_INVALID;
  std::memcpy(lut256, h->sim_lut, 256);
  return LMB200_OK;
}
int lmb200_set_normal_lut(lmb200_handle h, const uint8_t* lut8000) {
  if (!h || !lut8000) return LMB200_E_INVALID;
  for (int i = 0; i < 8000; ++i) {
    uint8_t v = lut8000[i];
    if (v & (v - 1)) return set_error(h, LMB200_E_INVALID, "NORMAL_LUT entries must be 0 or one-hot");
  }
  std::memcpy(h->normal_lut, lut8000, 8000);
  h->luts_dirty = true;
  h->normal_lut_standin = false;
  return LMB200_OK;
}
int lmb200_load_normal_lut(lmb200_handle h, const char* path) {
  if (!h || !path) return LMB200_E_INVALID;
  FILE* f = std::fopen(path, "rb");
  if (!f) return set_error(h, LMB200_E_IO, std::string("cannot open ") + path);
  std::string buf;
  char tmp[65536];
  size_t n;
  while ((n = std::fread(tmp, 1, sizeof tmp, f)) > 0) {
    buf.append(tmp, n);
    if (buf.size() > (64u << 20)) break;
  }
  std::fclose(f);
  if (buf.size() == 8000) return lmb200_set_normal_lut(h, (const uint8_t*)buf.data());
  // normal_lut.i: C initialiser text.  Comments are skipped; integer literals after the first '{' are the entries.
  std::vector<uint8_t> lut;
  size_t i = buf.find('{');
  if (i == std::string::npos) return set_error(h, LMB200_E_IO, std::string(path) + ": neither 8000 raw bytes nor a C initialiser");
  for (; i < buf.size() && lut.size() < 8000; ++i) {
    const char c = buf[i];
    if (c == '/' && i + 1 < buf.size() && buf[i + 1] == '/') { while (i < buf.size() && buf[i] != '\n') ++i; continue; }
    if (c == '/' && i + 1 < buf.size() && buf[i + 1] == '*') { size_t e = buf.find("*/", i + 2); if (e == std::string::npos) break; i = e + 1; continue; }
    if (c >= '0' && c <= '9') {
      char* end = nullptr;
      unsigned long v = std::strtoul(buf.c_str() + i, &end, 0);  // decimal, 0x.., or octal literals
      if (v > 255) return set_error(h, LMB200_E_IO, std::string(path) + ": entry does not fit a byte");
      lut.push_back((uint8_t)v);
      i = (size_t)(end - buf.c_str()) - 1;
    }
  }
  if (lut.size() != 8000) return set_error(h, LMB200_E_IO, std::string(path) + ": expected 8000 NORMAL_LUT entries, found " + std::to_string(lut.size()));
  return lmb200_set_normal_lut(h, lut.data());
}
int lmb200_normal_lut_is_standin(lmb200_handle h) { return h ? (h->normal_lut_standin ? 1 : 0) : LMB200_E_INVALID; }
const char* lmb200_warnings(lmb200_handle h) { return h ? h->warnings.c_str() : ""; }
int lmb200_get_normal_lut(lmb200_handle h, uint8_t* lut8000) {
  if (!h || !lut8000) return LMB200_E_INVALID;
  std::memcpy(lut8000, h->normal_lut, 8000);
  return LMB200_OK;
}

int lmb200_set_template_shard(lmb200_handle h, int rank, int world) {
  if (!h || world < 1 || rank < 0 || rank >= world) return LMB200_E_INVALID;
  h->shard_rank = rank; h->shard_world = world;
  h->sel_key.clear();
  return LMB200_OK;
}

}  // extern "C"
