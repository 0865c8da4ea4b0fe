// persistence.cpp — OpenCV FileStorage (YAML 1.0, optional gzip) read/write of the detector and its
// template classes.  File layout = what the reference writes/reads in
// HighLevelLineMOD::writeLinemod / readLinemod (src/HighLevelLinemod.cpp:256-270, :288-300):
//   Detector::write at the root (pyramid_levels, T, modalities) + "classes": [ {writeClass}, ... ]
// and what Detector::writeClasses/readClasses produce per class (SURVEY.md §8c "Template file layout",
// verified loadable by real cv2.FileStorage in tests/test_persistence.py).
// The reader is a streaming, schema-directed line parser: a 20 000-template file is ~4 M lines and
// a DOM would cost gigabytes.
#include <zlib.h>

#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>

#include "detector.h"

namespace lmh {

namespace {

struct Writer {
  gzFile gz = nullptr;
  FILE* fp = nullptr;
  std::string buf;
  bool failed = false;  // a short write or a failing close: the caller reports LMB200_E_IO
  bool open(const char* path) {
    size_t n = std::strlen(path);
    if (n > 3 && std::strcmp(path + n - 3, ".gz") == 0) gz = gzopen(path, "wb6");
    else fp = std::fopen(path, "wb");
    return gz || fp;
  }
  void flush() {
    if (buf.empty()) return;
    if (gz) { if (gzwrite(gz, buf.data(), (unsigned)buf.size()) != (int)buf.size()) failed = true; }
    else if (std::fwrite(buf.data(), 1, buf.size(), fp) != buf.size()) failed = true;
    buf.clear();
  }
  void put(const std::string& s) {
    buf += s;
    if (buf.size() > (1u << 20)) flush();
  }
  bool close() {  // true when every byte reached the file
    flush();
    if (gz && gzclose(gz) != Z_OK) failed = true;
    if (fp && std::fclose(fp) != 0) failed = true;
    gz = nullptr; fp = nullptr;
    return !failed;
  }
};

std::string fmt_float(float v) {  // OpenCV style: "10." for integral values, shortest round-trip otherwise
  char b[64];
  if (v == (float)(long long)v && std::fabs(v) < 1e15f) { std::snprintf(b, sizeof b, "%lld.", (long long)v); return b; }
  std::snprintf(b, sizeof b, "%.8e", (double)v);
  return b;
}
std::string quote(const std::string& s) {
  std::string o = "\"";
  for (char c : s) { if (c == '"' || c == '\\') o.push_back('\\'); o.push_back(c); }
  return o + "\"";
}
std::string ind(int n) { return std::string((size_t)n, ' '); }

void write_header(Writer& w, const lmb200_detector* h) {
  w.put("%YAML:1.0\n---\n");
  w.put("pyramid_levels: " + std::to_string(h->cfg.pyramid_levels) + "\n");
  std::string t = "T: [ ";
  for (int l = 0; l < h->cfg.pyramid_levels; ++l) t += std::to_string(h->cfg.T[l]) + (l + 1 < h->cfg.pyramid_levels ? ", " : "");
  w.put(t + " ]\n");
  w.put("modalities:\n");
  for (int m = 0; m < h->cfg.num_modalities; ++m) {
    const lmb200_modality& mo = h->cfg.modalities[m];
    w.put("   -\n");
    if (mo.type == LMB200_COLOR_GRADIENT) {
      w.put("      type: ColorGradient\n");
      w.put("      weak_threshold: " + fmt_float(mo.weak_threshold) + "\n");
      w.put("      num_features: " + std::to_string(mo.num_features) + "\n");
      w.put("      strong_threshold: " + fmt_float(mo.strong_threshold) + "\n");
    } else {
      w.put("      type: DepthNormal\n");
      w.put("      distance_threshold: " + std::to_string(mo.distance_threshold) + "\n");
      w.put("      difference_threshold: " + std::to_string(mo.difference_threshold) + "\n");
      w.put("      num_features: " + std::to_string(mo.num_features) + "\n");
      w.put("      extract_threshold: " + std::to_string(mo.extract_threshold) + "\n");
    }
  }
}

const char* mod_name(int type) { return type == LMB200_COLOR_GRADIENT ? "ColorGradient" : "DepthNormal"; }

// writeClass body at indentation `in`
void write_class(Writer& w, const lmb200_detector* h, const std::string& id, const std::vector<TemplatePyramid>& tps, int in) {
  w.put(ind(in) + "class_id: " + quote(id) + "\n");
  std::string m = ind(in) + "modalities: [ ";
  for (int i = 0; i < h->cfg.num_modalities; ++i) m += std::string(mod_name(h->cfg.modalities[i].type)) + (i + 1 < h->cfg.num_modalities ? ", " : "");
  w.put(m + " ]\n");
  w.put(ind(in) + "pyramid_levels: " + std::to_string(h->cfg.pyramid_levels) + "\n");
  w.put(ind(in) + "template_pyramids:\n");
  char line[96];
  for (size_t t = 0; t < tps.size(); ++t) {
    w.put(ind(in + 3) + "-\n");
    w.put(ind(in + 6) + "template_id: " + std::to_string(t) + "\n");
    w.put(ind(in + 6) + "templates:\n");
    for (const Template& tp : tps[t]) {
      w.put(ind(in + 9) + "-\n");
      w.put(ind(in + 12) + "width: " + std::to_string(tp.width) + "\n");
      w.put(ind(in + 12) + "height: " + std::to_string(tp.height) + "\n");
      w.put(ind(in + 12) + "pyramid_level: " + std::to_string(tp.pyramid_level) + "\n");
      w.put(ind(in + 12) + "features:\n");
      std::string pre = ind(in + 15);
      for (const Feature& f : tp.features) {
        std::snprintf(line, sizeof line, "- [ %d, %d, %d ]\n", f.x, f.y, f.label);
        w.put(pre + line);
      }
    }
  }
}

// ------------------------------------------------------------------ reader
struct Reader {
  gzFile gz = nullptr;
  std::string line;
  bool open(const char* path) { gz = gzopen(path, "rb"); if (gz) gzbuffer(gz, 1 << 18); return gz != nullptr; }
  bool next(std::string& out) {
    char buf[4096];
    out.clear();
    for (;;) {
      if (!gzgets(gz, buf, sizeof buf)) return !out.empty();
      out += buf;
      if (!out.empty() && out.back() == '\n') { out.pop_back(); if (!out.empty() && out.back() == '\r') out.pop_back(); return true; }
    }
  }
  void close() { if (gz) gzclose(gz); gz = nullptr; }
};

std::string trim(const std::string& s) {
  size_t a = s.find_first_not_of(" \t"), b = s.find_last_not_of(" \t");
  return a == std::string::npos ? std::string() : s.substr(a, b - a + 1);
}
std::string unquote(const std::string& s0) {
  std::string s = trim(s0);
  if (s.size() >= 2 && (s.front() == '"' || s.front() == '\'') && s.back() == s.front()) {
    std::string o;
    for (size_t i = 1; i + 1 < s.size(); ++i) {
      if (s[i] == '\\' && s.front() == '"' && i + 2 < s.size()) ++i;
      o.push_back(s[i]);
    }
    return o;
  }
  return s;
}
// "[ a, b, c ]" -> items (handles "[:" too)
std::vector<std::string> flow_items(const std::string& s) {
  std::vector<std::string> out;
  size_t a = s.find('['), b = s.rfind(']');
  if (a == std::string::npos || b == std::string::npos || b < a) return out;
  std::string body = s.substr(a + 1, b - a - 1);
  if (!body.empty() && body[0] == ':') body.erase(0, 1);
  size_t p = 0;
  while (p <= body.size()) {
    size_t q = body.find(',', p);
    if (q == std::string::npos) q = body.size();
    std::string it = trim(body.substr(p, q - p));
    if (!it.empty()) out.push_back(unquote(it));
    p = q + 1;
  }
  return out;
}

struct ParseState {
  // detector-level (root) config
  bool have_levels = false, have_T = false;
  int pyramid_levels = 0;
  std::vector<int> T;
  std::vector<lmb200_modality> mods;
  // classes
  struct Cls { std::string id; bool have_id = false; std::vector<std::string> mod_names; int pyramid_levels = -1; std::vector<TemplatePyramid> tps; };
  std::vector<Cls> classes;
};

int parse_file(const char* path, ParseState& S, std::string& err) {
  Reader rd;
  if (!rd.open(path)) { err = std::string("cannot open ") + path; return LMB200_E_IO; }
  std::string raw;
  bool in_classes = false;         // past "classes:" (or the file is a bare writeClass file)
  bool in_root_modalities = false;
  ParseState::Cls* cls = nullptr;
  Template* tpl = nullptr;
  long lineno = 0;
  auto fail = [&](const std::string& m) { err = std::string(path) + ":" + std::to_string(lineno) + ": " + m; rd.close(); return LMB200_E_IO; };
  while (rd.next(raw)) {
    ++lineno;
    std::string s = trim(raw);
    if (s.empty() || s[0] == '%' || s == "---" || s == "..." || s[0] == '#') continue;
    // feature line (hot path): "- [ x, y, label ]"
    if (s[0] == '-' && s.find('[') != std::string::npos && tpl) {
      const char* p = s.c_str() + s.find('[') + 1;
      if (*p == ':') ++p;
      char* e;
      long x = std::strtol(p, &e, 10); while (*e == ' ' || *e == ',') ++e;
      long y = std::strtol(e, &e, 10); while (*e == ' ' || *e == ',') ++e;
      long l = std::strtol(e, &e, 10);
      tpl->features.push_back(Feature{(int)x, (int)y, (int)l});
      continue;
    }
    if (s[0] == '-') {  // "-" alone or "- key: value"
      s = trim(s.substr(1));
      if (s.empty()) continue;
    }
    // flow sequences may wrap over several lines
    size_t colon = s.find(':');
    if (colon == std::string::npos) continue;
    std::string key = trim(s.substr(0, colon)), val = trim(s.substr(colon + 1));
    // flow sequences may wrap over several lines: only an unquoted value that opens with '[' continues
    if (!val.empty() && val[0] == '[' && val.find(']') == std::string::npos) {
      std::string more;
      while (val.find(']') == std::string::npos && rd.next(more)) { ++lineno; val += " " + trim(more); }
    }
    if (key == "classes") { in_classes = true; in_root_modalities = false; continue; }
    if (key == "class_id") {
      in_classes = true; in_root_modalities = false;
      // a bare class file may have listed `modalities` first: that pending entry (no id, no templates yet) is this class
      if (!(cls && cls->id.empty() && !cls->have_id && cls->tps.empty())) { S.classes.emplace_back(); cls = &S.classes.back(); }
      tpl = nullptr;
      cls->id = unquote(val); cls->have_id = true;
      continue;
    }
    if (key == "modalities") {
      if (!val.empty()) {  // class-level flow list of names
        auto names = flow_items(val);
        if (cls) cls->mod_names = names;
        else {  // bare class file may list modalities before class_id: remember for the next class
          S.classes.emplace_back(); cls = &S.classes.back(); cls->mod_names = names; in_classes = true;
        }
      } else if (!in_classes) in_root_modalities = true;
      continue;
    }
    if (key == "pyramid_levels") {
      int v = std::atoi(val.c_str());
      if (cls) cls->pyramid_levels = v; else { S.pyramid_levels = v; S.have_levels = true; }
      continue;
    }
    if (key == "T" && !in_classes) {
      for (auto& it : flow_items(val)) S.T.push_back(std::atoi(it.c_str()));
      S.have_T = true;
      continue;
    }
    if (in_root_modalities) {
      if (key == "type") {
        lmb200_modality mo;
        std::string t = unquote(val);
        if (t == "ColorGradient") lmb200_default_modality(LMB200_COLOR_GRADIENT, &mo);
        else if (t == "DepthNormal") lmb200_default_modality(LMB200_DEPTH_NORMAL, &mo);
        else return fail("unknown modality type '" + t + "'");
        S.mods.push_back(mo);
      } else if (!S.mods.empty()) {
        lmb200_modality& mo = S.mods.back();
        if (key == "weak_threshold") mo.weak_threshold = (float)std::atof(val.c_str());
        else if (key == "strong_threshold") mo.strong_threshold = (float)std::atof(val.c_str());
        else if (key == "num_features") mo.num_features = std::atoi(val.c_str());
        else if (key == "distance_threshold") mo.distance_threshold = std::atoi(val.c_str());
        else if (key == "difference_threshold") mo.difference_threshold = std::atoi(val.c_str());
        else if (key == "extract_threshold") mo.extract_threshold = std::atoi(val.c_str());
      }
      continue;
    }
    if (!cls) continue;
    if (key == "template_pyramids") continue;
    if (key == "template_id") {
      int id = std::atoi(val.c_str());
      if (id != (int)cls->tps.size()) return fail("template_id " + std::to_string(id) + " is not consecutive (upstream CV_Assert)");
      cls->tps.emplace_back();
      tpl = nullptr;
      continue;
    }
    if (key == "templates") continue;
    if (key == "width") {
      if (cls->tps.empty()) return fail("template outside a template pyramid");
      cls->tps.back().emplace_back();
      tpl = &cls->tps.back().back();
      tpl->width = std::atoi(val.c_str());
      continue;
    }
    if (key == "height" && tpl) { tpl->height = std::atoi(val.c_str()); continue; }
    if (key == "pyramid_level" && tpl) { tpl->pyramid_level = std::atoi(val.c_str()); continue; }
    if (key == "features") {
      if (!val.empty() && tpl) {  // everything in one (wrapped) flow sequence: [ [x,y,l], ... ] or empty "[]"
        const char* p = val.c_str();
        std::vector<long> nums;
        while (*p) {
          char* e = const_cast<char*>(p);
          if ((*p >= '0' && *p <= '9') || *p == '-') { long v = std::strtol(p, &e, 10); if (e != p) nums.push_back(v); }
          p = e != p ? e : p + 1;  // a '-' that starts no number must not stall the scan
        }
        for (size_t i = 0; i + 2 < nums.size(); i += 3) tpl->features.push_back(Feature{(int)nums[i], (int)nums[i + 1], (int)nums[i + 2]});
      }
      continue;
    }
  }
  rd.close();
  return LMB200_OK;
}

int validate_pyramid(const TemplatePyramid& tp, const std::string& what, std::string& err) {
  for (const Template& t : tp) {
    if (t.features.size() > 63) { err = what + ": template has more than 63 features (upstream CV_Assert)"; return LMB200_E_FEATURES; }
    if (t.width < 0 || t.height < 0 || t.width > 32767 || t.height > 32767) { err = what + ": template width/height out of range"; return LMB200_E_IO; }
    for (const Feature& f : t.features)
      if (f.label < 0 || f.label > 7) { err = what + ": feature label must be 0..7"; return LMB200_E_IO; }
  }
  return LMB200_OK;
}

int add_parsed_class(lmb200_detector* h, ParseState::Cls& c, const char* override_id, std::string& err) {
  const int M = h->cfg.num_modalities, L = h->cfg.pyramid_levels;
  if ((int)c.mod_names.size() != M) { err = "class '" + c.id + "': modality count differs from the detector's (upstream CV_Assert)"; return LMB200_E_CLASS; }
  for (int m = 0; m < M; ++m)
    if (c.mod_names[m] != mod_name(h->cfg.modalities[m].type)) { err = "class '" + c.id + "': modality names differ from the detector's"; return LMB200_E_CLASS; }
  if (c.pyramid_levels != L) { err = "class '" + c.id + "': pyramid_levels differs from the detector's"; return LMB200_E_CLASS; }
  std::string id = (override_id && *override_id) ? override_id : c.id;
  if (!(override_id && *override_id) && h->classes.count(id)) { err = "class '" + id + "' already present (upstream CV_Assert)"; return LMB200_E_CLASS; }
  for (auto& tp : c.tps) {
    if ((int)tp.size() != M * L) { err = "class '" + id + "': template pyramid has " + std::to_string(tp.size()) + " templates, expected " + std::to_string(M * L); return LMB200_E_IO; }
    int rc = validate_pyramid(tp, "class '" + id + "'", err);   // the same limits lmb200_add_synthetic_template enforces
    if (rc) return rc;
  }
  if (!h->classes.count(id)) h->classes[id] = std::move(c.tps);  // std::map::insert semantics: existing entry wins
  h->templates_dirty = true;
  return LMB200_OK;
}

}  // namespace

int write_detector_file(lmb200_detector* h, const char* path) {
  Writer w;
  if (!w.open(path)) return set_error(h, LMB200_E_IO, std::string("cannot open ") + path + " for writing");
  write_header(w, h);
  w.put("classes:\n");
  for (auto& kv : h->classes) {
    w.put("   -\n");
    write_class(w, h, kv.first, kv.second, 6);
  }
  if (!w.close()) return set_error(h, LMB200_E_IO, std::string("write error on ") + path + " (disk full?)");
  return LMB200_OK;
}

int write_class_file(lmb200_detector* h, const std::string& class_id, const char* path) {
  auto it = h->classes.find(class_id);
  if (it == h->classes.end()) return set_error(h, LMB200_E_CLASS, "unknown class " + class_id);
  Writer w;
  if (!w.open(path)) return set_error(h, LMB200_E_IO, std::string("cannot open ") + path + " for writing");
  w.put("%YAML:1.0\n---\n");
  write_class(w, h, it->first, it->second, 0);
  if (!w.close()) return set_error(h, LMB200_E_IO, std::string("write error on ") + path + " (disk full?)");
  return LMB200_OK;
}

int read_detector_file(const char* path, int device, lmb200_handle* out, std::string& err) {
  ParseState S;
  int rc = parse_file(path, S, err);
  if (rc) return rc;
  if (!S.have_levels || !S.have_T || S.mods.empty() || (int)S.T.size() != S.pyramid_levels) {
    err = std::string(path) + ": missing pyramid_levels / T / modalities at the root";
    return LMB200_E_IO;
  }
  if (S.pyramid_levels > LMB200_MAX_LEVELS || (int)S.mods.size() > LMB200_MAX_MODALITIES) { err = "too many levels/modalities"; return LMB200_E_INVALID; }
  lmb200_config cfg;
  std::memset(&cfg, 0, sizeof cfg);
  cfg.num_modalities = (int)S.mods.size();
  for (int m = 0; m < cfg.num_modalities; ++m) cfg.modalities[m] = S.mods[m];
  cfg.pyramid_levels = S.pyramid_levels;
  for (int l = 0; l < cfg.pyramid_levels; ++l) cfg.T[l] = S.T[l];
  cfg.device = device;
  lmb200_handle h = nullptr;
  rc = lmb200_create(&cfg, &h);
  if (rc) { err = lmb200_last_error(nullptr); return rc; }
  for (auto& c : S.classes) {
    rc = add_parsed_class(h, c, nullptr, err);
    if (rc) { lmb200_destroy(h); return rc; }
  }
  *out = h;
  return LMB200_OK;
}

int read_class_file(lmb200_detector* h, const char* path, std::string& err, const char* override_id) {
  ParseState S;
  int rc = parse_file(path, S, err);
  if (rc) return rc;
  if (S.classes.empty()) { err = std::string(path) + ": no class found"; return LMB200_E_IO; }
  for (auto& c : S.classes) {
    rc = add_parsed_class(h, c, override_id, err);
    if (rc) return rc;
  }
  return LMB200_OK;
}

// ------------------------------------------------------------------ binary cache
namespace {
const char CACHE_MAGIC[8] = {'L', 'M', 'B', '2', 'T', 'P', 'L', '1'};
template <typename T> void put(std::string& b, const T& v) { b.append(reinterpret_cast<const char*>(&v), sizeof(T)); }
template <typename T> bool get(const std::string& b, size_t& p, T& v) {
  if (p + sizeof(T) > b.size()) return false;
  std::memcpy(&v, b.data() + p, sizeof(T));
  p += sizeof(T);
  return true;
}
}  // namespace

int write_cache_file(lmb200_detector* h, const char* path) {
  std::string b(CACHE_MAGIC, 8);
  put(b, h->cfg);
  put(b, (uint32_t)h->classes.size());
  for (auto& kv : h->classes) {
    put(b, (uint32_t)kv.first.size());
    b += kv.first;
    put(b, (uint32_t)kv.second.size());
    for (auto& tp : kv.second)
      for (auto& t : tp) {
        put(b, (int32_t)t.width); put(b, (int32_t)t.height); put(b, (int32_t)t.pyramid_level); put(b, (uint32_t)t.features.size());
        for (auto& f : t.features) {
          if (f.x < -32768 || f.x > 32767 || f.y < -32768 || f.y > 32767 || f.label < 0 || f.label > 255)
            return set_error(h, LMB200_E_INVALID, "feature outside the cache's 16-bit coordinate range (use lmb200_write)");
          put(b, (int16_t)f.x); put(b, (int16_t)f.y); put(b, (uint8_t)f.label);
        }
      }
  }
  FILE* fp = std::fopen(path, "wb");
  if (!fp) return set_error(h, LMB200_E_IO, std::string("cannot open ") + path + " for writing");
  size_t w = std::fwrite(b.data(), 1, b.size(), fp);
  const bool closed = std::fclose(fp) == 0;
  return (w == b.size() && closed) ? LMB200_OK : set_error(h, LMB200_E_IO, "short write");
}

int read_cache_file(const char* path, int device, lmb200_handle* out, std::string& err) {
  FILE* fp = std::fopen(path, "rb");
  if (!fp) { err = std::string("cannot open ") + path; return LMB200_E_IO; }
  std::string b;
  char buf[1 << 16];
  size_t n;
  while ((n = std::fread(buf, 1, sizeof buf, fp)) > 0) b.append(buf, n);
  std::fclose(fp);
  size_t p = 8;
  lmb200_config cfg;
  uint32_t ncls = 0;
  if (b.size() < 8 || std::memcmp(b.data(), CACHE_MAGIC, 8) != 0 || !get(b, p, cfg) || !get(b, p, ncls)) { err = "not an lmb200 template cache"; return LMB200_E_IO; }
  cfg.device = device;
  lmb200_handle h = nullptr;
  int rc = lmb200_create(&cfg, &h);
  if (rc) { err = lmb200_last_error(nullptr); return rc; }
  const int per = cfg.num_modalities * cfg.pyramid_levels;
  for (uint32_t c = 0; c < ncls; ++c) {
    uint32_t len = 0, nt = 0;
    if (!get(b, p, len) || p + len > b.size()) { lmb200_destroy(h); err = "truncated cache"; return LMB200_E_IO; }
    std::string id = b.substr(p, len);
    p += len;
    if (!get(b, p, nt)) { lmb200_destroy(h); err = "truncated cache"; return LMB200_E_IO; }
    if (per <= 0 || (unsigned long long)nt * (unsigned long long)per * 16ull > b.size() - p) {  // 16 B = an empty template record
      lmb200_destroy(h); err = "corrupt cache (template count)"; return LMB200_E_IO;
    }
    std::vector<TemplatePyramid>& tps = h->classes[id];
    tps.resize(nt);
    for (uint32_t t = 0; t < nt; ++t) {
      tps[t].resize(per);
      for (int k = 0; k < per; ++k) {
        int32_t w, hh, lv; uint32_t nf;
        if (!get(b, p, w) || !get(b, p, hh) || !get(b, p, lv) || !get(b, p, nf) || nf > 63 || p + 5ull * nf > b.size()) { lmb200_destroy(h); err = "corrupt cache"; return LMB200_E_IO; }
        Template& tm = tps[t][k];
        if (w < 0 || hh < 0 || w > 32767 || hh > 32767) { lmb200_destroy(h); err = "corrupt cache (template size)"; return LMB200_E_IO; }
        tm.width = w; tm.height = hh; tm.pyramid_level = lv;
        tm.features.resize(nf);
        for (uint32_t i = 0; i < nf; ++i) {
          int16_t x = 0, y = 0; uint8_t l = 0;
          get(b, p, x); get(b, p, y); get(b, p, l);
          if (l > 7) { lmb200_destroy(h); err = "corrupt cache (feature label > 7)"; return LMB200_E_IO; }
          tm.features[i] = Feature{x, y, l};
        }
      }
    }
  }
  h->templates_dirty = true;
  *out = h;
  return LMB200_OK;
}

}  // namespace lmh

extern "C" {

int lmb200_write_cache(lmb200_handle h, const char* path) { return (h && path) ? lmh::write_cache_file(h, path) : LMB200_E_INVALID; }
int lmb200_read_cache(const char* path, int device, lmb200_handle* out) {
  if (!path || !out) return LMB200_E_INVALID;
  std::string err;
  int rc;
  try {  // no exception may cross the C ABI
    rc = lmh::read_cache_file(path, device, out, err);
  } catch (const std::exception& e) {
    rc = LMB200_E_IO; err = std::string(path) + ": " + e.what();
  }
  if (rc) lmh::set_create_error(err);
  return rc;
}

int lmb200_read_pose_sidecar(const char* path, int class_index, lmb200_template_pose* out, size_t cap, size_t* n_out) {
  static_assert(sizeof(lmb200_template_pose) == 48, "must mirror HighLevelLineMOD::Template");
  if (!path || !n_out || class_index < 0) return LMB200_E_INVALID;
  *n_out = 0;
  FILE* fp = std::fopen(path, "rb");
  if (!fp) return LMB200_E_IO;
  uint32_t ncls = 0;
  int rc = LMB200_E_IO;
  std::fseek(fp, 0, SEEK_END);
  const long long file_size = std::ftell(fp);
  std::fseek(fp, 0, SEEK_SET);
  if (std::fread(&ncls, sizeof ncls, 1, fp) == 1) {
    for (uint32_t c = 0; c < ncls; ++c) {
      uint64_t n = 0;
      if (std::fread(&n, sizeof n, 1, fp) != 1) break;
      const long long here = std::ftell(fp);
      if (here < 0 || n > (uint64_t)(file_size - here) / sizeof(lmb200_template_pose)) break;  // count beyond the file: corrupt
      if ((int)c == class_index) {
        *n_out = (size_t)n;
        size_t take = n < cap ? (size_t)n : cap;
        if (out && take && std::fread(out, sizeof(lmb200_template_pose), take, fp) != take) break;
        rc = (take < n) ? LMB200_E_TRUNCATED : LMB200_OK;
        break;
      }
      if (std::fseek(fp, (long)(n * sizeof(lmb200_template_pose)), SEEK_CUR) != 0) break;
    }
    if (rc == LMB200_E_IO && class_index >= (int)ncls) rc = LMB200_E_CLASS;
  }
  std::fclose(fp);
  return rc;
}

int lmb200_write_pose_sidecar(const char* path, const lmb200_template_pose* const* per_class, const size_t* counts, int n_classes) {
  if (!path || n_classes < 0 || (n_classes > 0 && (!per_class || !counts))) return LMB200_E_INVALID;
  FILE* fp = std::fopen(path, "wb");
  if (!fp) return LMB200_E_IO;
  uint32_t ncls = (uint32_t)n_classes;
  bool ok = std::fwrite(&ncls, sizeof ncls, 1, fp) == 1;
  for (int c = 0; c < n_classes && ok; ++c) {
    uint64_t n = counts[c];
    ok = std::fwrite(&n, sizeof n, 1, fp) == 1;
    if (ok && n) ok = std::fwrite(per_class[c], sizeof(lmb200_template_pose), (size_t)n, fp) == (size_t)n;
  }
  if (std::fclose(fp) != 0) ok = false;
  return ok ? LMB200_OK : LMB200_E_IO;
}

}  // extern "C"
