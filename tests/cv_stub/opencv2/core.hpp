// tests/cv_stub/opencv2/core.hpp — MINIMAL stand-in for the parts of <opencv2/core.hpp> that include/lmb200_opencv.hpp and
// the reference's HighLevelLinemod.cpp touch on the cv::linemod path (OpenCV's C++ headers are not installed in this
// image).  Test infrastructure only: it exists so tests/cpp_dropin_check.cpp can compile the reference's call
// expressions verbatim against the drop-in header and run them through the C ABI.  Semantics follow OpenCV 4:
//   cv::Ptr / makePtr (shared ownership, release()), cv::Mat (non-owning view or owning buffer), cv::Rect / Point / Size,
//   cv::FileStorage streaming writer ("{" "[" "[:" "}" "]" state machine) and FileNode tree reader.
// FileStorage(WRITE) emits OpenCV-style YAML 1.0 on release (plain text; ".gz" names are written uncompressed — the
// product's reader accepts both) and keeps the tree in a process-wide map, which FileStorage(READ) consults first.
#pragma once
#include <cstdio>
#include <cstring>
#include <map>
#include <memory>
#include <string>
#include <vector>

#define CV_8U 0
#define CV_16U 2
#define CV_8UC1 0
#define CV_16UC1 2
#define CV_8UC3 16

namespace cv {

typedef std::string String;

template <typename T>
struct Ptr : std::shared_ptr<T> {
  Ptr() {}
  Ptr(T* p) : std::shared_ptr<T>(p) {}
  Ptr(const std::shared_ptr<T>& p) : std::shared_ptr<T>(p) {}
  template <typename U> Ptr(const Ptr<U>& o) : std::shared_ptr<T>(std::static_pointer_cast<T>(static_cast<const std::shared_ptr<U>&>(o))) {}
  void release() { this->reset(); }
  bool empty() const { return this->get() == nullptr; }
};
template <typename T, typename... A> Ptr<T> makePtr(A&&... a) { return Ptr<T>(std::make_shared<T>(std::forward<A>(a)...)); }

template <typename T> struct Point_ { T x, y; Point_() : x(0), y(0) {} Point_(T x_, T y_) : x(x_), y(y_) {} Point_ operator+(const Point_& o) const { return Point_(x + o.x, y + o.y); } };
typedef Point_<int> Point;
template <typename T> struct Size_ { T width, height; Size_() : width(0), height(0) {} Size_(T w, T h) : width(w), height(h) {} };
typedef Size_<int> Size;
template <typename T> struct Rect_ { T x, y, width, height; Rect_() : x(0), y(0), width(0), height(0) {} Rect_(T x_, T y_, T w, T h) : x(x_), y(y_), width(w), height(h) {} };
typedef Rect_<int> Rect;

struct MatStep { size_t v = 0; operator size_t() const { return v; } };
class Mat {
 public:
  int rows = 0, cols = 0;
  unsigned char* data = nullptr;
  MatStep step;
  Mat() {}
  Mat(int r, int c, int type) { create(r, c, type); }
  Mat(int r, int c, int type, void* external, size_t step_ = 0) : rows(r), cols(c), data((unsigned char*)external), type_(type) { step.v = step_ ? step_ : (size_t)c * elemSize(); }
  void create(int r, int c, int type) {
    rows = r; cols = c; type_ = type; step.v = (size_t)c * elemSize();
    buf_ = std::make_shared<std::vector<unsigned char>>((size_t)r * step.v, (unsigned char)0);
    data = buf_->data();
  }
  static Mat zeros(int r, int c, int type) { return Mat(r, c, type); }
  int type() const { return type_; }
  bool empty() const { return data == nullptr || rows == 0 || cols == 0; }
  size_t elemSize() const { return type_ == CV_8UC3 ? 3 : (type_ == CV_16UC1 ? 2 : 1); }
  template <typename T> T& at(int r, int c) { return *(T*)(data + (size_t)r * step.v + (size_t)c * sizeof(T)); }
  template <typename T> T* ptr(int r) { return (T*)(data + (size_t)r * step.v); }
 private:
  int type_ = 0;
  std::shared_ptr<std::vector<unsigned char>> buf_;
};

// ------------------------------------------------------------------------------------------------ persistence
struct FsNode {
  enum Kind { NONE, SCALAR, SEQ, MAP } kind = NONE;
  bool flow = false;                       // "[:" / "{:" containers
  bool quoted = false;                     // scalar written from a string
  std::string scalar;
  std::vector<std::shared_ptr<FsNode>> items;                 // SEQ
  std::vector<std::pair<std::string, std::shared_ptr<FsNode>>> fields;  // MAP (ordered)
};

class FileNode;
class FileNodeIterator {
 public:
  FileNodeIterator(const FsNode* n, size_t i) : n_(n), i_(i) {}
  FileNode operator*() const;
  FileNodeIterator& operator++() { ++i_; return *this; }
  bool operator!=(const FileNodeIterator& o) const { return i_ != o.i_ || n_ != o.n_; }
  bool operator==(const FileNodeIterator& o) const { return !(*this != o); }
 private:
  const FsNode* n_; size_t i_;
  friend FileNodeIterator& operator>>(FileNodeIterator& it, int& v);
};

class FileNode {
 public:
  FileNode() {}
  explicit FileNode(std::shared_ptr<const FsNode> n) : n_(n) {}
  FileNode(const FsNode* raw) : raw_(raw) {}
  const FsNode* node() const { return n_ ? n_.get() : raw_; }
  bool empty() const { return node() == nullptr || node()->kind == FsNode::NONE; }
  bool isSeq() const { return node() && node()->kind == FsNode::SEQ; }
  bool isMap() const { return node() && node()->kind == FsNode::MAP; }
  size_t size() const { return !node() ? 0 : node()->kind == FsNode::SEQ ? node()->items.size() : node()->kind == FsNode::MAP ? node()->fields.size() : 1; }
  FileNode operator[](const char* key) const {
    if (isMap()) for (auto& f : node()->fields) if (f.first == key) return FileNode(f.second.get());
    return FileNode();
  }
  FileNode operator[](const std::string& key) const { return (*this)[key.c_str()]; }
  FileNode operator[](int i) const { return isSeq() && (size_t)i < node()->items.size() ? FileNode(node()->items[(size_t)i].get()) : FileNode(); }
  FileNodeIterator begin() const { return FileNodeIterator(node(), 0); }
  FileNodeIterator end() const { return FileNodeIterator(node(), isSeq() ? node()->items.size() : isMap() ? node()->fields.size() : 0); }
  operator int() const { return empty() ? 0 : std::atoi(node()->scalar.c_str()); }
  operator float() const { return empty() ? 0.f : (float)std::atof(node()->scalar.c_str()); }
  operator double() const { return empty() ? 0.0 : std::atof(node()->scalar.c_str()); }
  operator std::string() const { return empty() ? std::string() : node()->scalar; }
 private:
  std::shared_ptr<const FsNode> n_;
  const FsNode* raw_ = nullptr;
};
inline FileNode FileNodeIterator::operator*() const {
  if (!n_) return FileNode();
  if (n_->kind == FsNode::SEQ) return FileNode(n_->items[i_].get());
  if (n_->kind == FsNode::MAP) return FileNode(n_->fields[i_].second.get());
  return FileNode();
}
inline FileNodeIterator& operator>>(FileNodeIterator& it, int& v) { v = (int)*it; ++it; return it; }
inline void operator>>(const FileNode& n, int& v) { v = (int)n; }
inline void operator>>(const FileNode& n, float& v) { v = (float)n; }
inline void operator>>(const FileNode& n, std::string& v) { v = (std::string)n; }
inline void operator>>(const FileNode& n, std::vector<int>& v) { v.clear(); for (auto it = n.begin(); it != n.end(); ++it) v.push_back((int)*it); }

class FileStorage {
 public:
  enum Mode { READ = 0, WRITE = 1 };
  FileStorage() {}
  FileStorage(const std::string& filename, int flags) { open(filename, flags); }
  ~FileStorage() { release(); }
  bool open(const std::string& filename, int flags) {
    name_ = filename; mode_ = flags; opened_ = true;
    if (flags == WRITE) { root_ = std::make_shared<FsNode>(); root_->kind = FsNode::MAP; stack_.assign(1, root_.get()); expect_key_ = true; }
    else { auto it = files().find(filename); if (it == files().end()) { opened_ = false; return false; } root_ = it->second; }
    return true;
  }
  bool isOpened() const { return opened_; }
  FileNode root() const { return FileNode(std::shared_ptr<const FsNode>(root_)); }
  FileNode operator[](const char* key) const { return root()[key]; }
  FileNode operator[](const std::string& key) const { return root()[key.c_str()]; }
  void release() {
    if (opened_ && mode_ == WRITE && root_) { files()[name_] = root_; dump(); }
    opened_ = false;
  }
  // ---- streaming writer
  void put_string(const std::string& s) {
    FsNode* top = stack_.back();
    if (s == "}" || s == "]") { stack_.pop_back(); expect_key_ = stack_.back()->kind == FsNode::MAP; return; }
    const bool opens = s == "{" || s == "[" || s == "{:" || s == "[:";
    if (top->kind == FsNode::MAP && expect_key_ && !opens) { key_ = s; expect_key_ = false; return; }
    auto n = std::make_shared<FsNode>();
    if (opens) { n->kind = s[0] == '{' ? FsNode::MAP : FsNode::SEQ; n->flow = s.size() > 1; }
    else { n->kind = FsNode::SCALAR; n->scalar = s; n->quoted = true; }
    attach(top, n);
    if (opens) { stack_.push_back(n.get()); expect_key_ = n->kind == FsNode::MAP; }
  }
  void put_scalar(const std::string& text) {
    auto n = std::make_shared<FsNode>();
    n->kind = FsNode::SCALAR; n->scalar = text;
    attach(stack_.back(), n);
  }
  static std::map<std::string, std::shared_ptr<FsNode>>& files() { static std::map<std::string, std::shared_ptr<FsNode>> f; return f; }
 private:
  void attach(FsNode* top, const std::shared_ptr<FsNode>& n) {
    if (top->kind == FsNode::MAP) { top->fields.emplace_back(key_, n); expect_key_ = true; }
    else top->items.push_back(n);
  }
  static std::string scalar_text(const FsNode& n) {
    if (!n.quoted) return n.scalar;
    bool plain = !n.scalar.empty();
    for (char c : n.scalar) if (!(std::isalnum((unsigned char)c) || c == '_')) plain = false;
    if (plain && !std::isdigit((unsigned char)n.scalar[0])) return n.scalar;
    std::string o = "\"";
    for (char c : n.scalar) { if (c == '"' || c == '\\') o.push_back('\\'); o.push_back(c); }
    return o + "\"";
  }
  static void emit_flow(const FsNode& n, std::string& out) {
    if (n.kind == FsNode::SCALAR) { out += scalar_text(n); return; }
    out += n.kind == FsNode::SEQ ? "[ " : "{ ";
    bool first = true;
    if (n.kind == FsNode::SEQ) for (auto& i : n.items) { if (!first) out += ", "; first = false; emit_flow(*i, out); }
    else for (auto& f : n.fields) { if (!first) out += ", "; first = false; out += f.first + ":"; emit_flow(*f.second, out); }
    out += n.kind == FsNode::SEQ ? " ]" : " }";
  }
  static void emit(const FsNode& n, int indent, std::string& out) {  // n is a block MAP or block SEQ
    const std::string pad((size_t)indent, ' ');
    if (n.kind == FsNode::MAP) {
      for (auto& f : n.fields) {
        const FsNode& v = *f.second;
        if (v.kind == FsNode::SCALAR || v.flow) { out += pad + f.first + ": "; emit_flow(v, out); out += "\n"; }
        else { out += pad + f.first + ":\n"; emit(v, indent + 3, out); }
      }
    } else {
      for (auto& i : n.items) {
        if (i->kind == FsNode::SCALAR || i->flow) { out += pad + "- "; emit_flow(*i, out); out += "\n"; }
        else { out += pad + "-\n"; emit(*i, indent + 3, out); }
      }
    }
  }
  void dump() const {
    std::string out = "%YAML:1.0\n---\n";
    emit(*root_, 0, out);
    if (FILE* f = std::fopen(name_.c_str(), "wb")) { std::fwrite(out.data(), 1, out.size(), f); std::fclose(f); }
  }
  std::string name_, key_;
  int mode_ = READ;
  bool opened_ = false, expect_key_ = true;
  std::shared_ptr<FsNode> root_;
  std::vector<FsNode*> stack_;
};
inline FileStorage& operator<<(FileStorage& fs, const char* s) { fs.put_string(s); return fs; }
inline FileStorage& operator<<(FileStorage& fs, const std::string& s) { fs.put_string(s); return fs; }
inline FileStorage& operator<<(FileStorage& fs, int v) { fs.put_scalar(std::to_string(v)); return fs; }
inline FileStorage& operator<<(FileStorage& fs, float v) {
  char b[64];
  if (v == (float)(long long)v) std::snprintf(b, sizeof b, "%lld.", (long long)v); else std::snprintf(b, sizeof b, "%.8e", (double)v);
  fs.put_scalar(b); return fs;
}
inline FileStorage& operator<<(FileStorage& fs, const std::vector<int>& v) { fs.put_string("[:"); for (int x : v) fs.put_scalar(std::to_string(x)); fs.put_string("]"); return fs; }

inline std::string format(const char* fmt, const char* arg) { char b[4096]; std::snprintf(b, sizeof b, fmt, arg); return b; }

}  // namespace cv
