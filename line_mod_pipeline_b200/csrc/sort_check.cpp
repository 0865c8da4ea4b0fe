// sort_check.cpp — host build of csrc/sort_emul.h next to the real std::sort / std::partial_sort, for the test that pins
// the device epilogue's sort to libstdc++'s (lmb200_debug_sort_check in include/lmb200.h).
#include <algorithm>
#include <cstring>
#include <vector>

#include "detector.h"
#include "sort_emul.h"

using namespace lmh;

namespace {
// M. D. McIlroy, "A Killer Adversary for Quicksort" (1999): builds, against the sort under test, the input on which its
// pivot choices are worst — std::sort then exhausts its depth limit and takes the heap-sort fallback.
struct Adversary {
  std::vector<int> val; int nsolid = 0, candidate = 0, gas;
  explicit Adversary(int n) : val(n), gas(n - 1) { for (auto& v : val) v = gas; }
  bool less(int x, int y) {
    if (val[x] == gas && val[y] == gas) { if (x == candidate) val[x] = nsolid++; else val[y] = nsolid++; }
    if (val[x] == gas) candidate = x; else if (val[y] == gas) candidate = y;
    return val[x] < val[y];
  }
};
}  // namespace

extern "C" int lmb200_debug_sort_check(const lmb200_match_rec* in, size_t n, int mode, lmb200_match_rec* out_emulated,
                                       lmb200_match_rec* out_std) {
  if ((!in && n) || !out_emulated || !out_std || mode < 0 || mode > 2 || n > (1u << 24)) return LMB200_E_INVALID;
  std::vector<Match> m(n);
  for (size_t i = 0; i < n; ++i) m[i] = Match{in[i].x, in[i].y, in[i].similarity, in[i].class_index, in[i].template_id};
  if (mode == 2 && n > 1) {   // overwrite the template ids with the adversary's permutation, one similarity for all
    Adversary adv((int)n);
    std::vector<int> idx(n);
    for (size_t i = 0; i < n; ++i) idx[i] = (int)i;
    std::sort(idx.begin(), idx.end(), [&adv](int a, int b) { return adv.less(a, b); });
    for (size_t i = 0; i < n; ++i) { m[i].template_id = adv.val[i]; m[i].similarity = 90.f; }
  }
  std::vector<uint64_t> key(n);
  std::vector<uint32_t> idx(n);
  for (size_t i = 0; i < n; ++i) { key[i] = lmsort::match_key(m[i].similarity, m[i].template_id); idx[i] = (uint32_t)i; }
  lmsort::Arr a{key.data(), idx.data()};
  std::vector<Match> s = m;
  if (mode == 1) {
    lmsort::heap_sort_(a, 0, (int)n);
    std::partial_sort(s.begin(), s.end(), s.end());
  } else {
    lmsort::std_sort(a, (int)n);
    std::sort(s.begin(), s.end());
  }
  for (size_t i = 0; i < n; ++i) {
    const Match& e = m[idx[i]];
    out_emulated[i].x = e.x; out_emulated[i].y = e.y; out_emulated[i].similarity = e.similarity; out_emulated[i].class_index = e.class_index; out_emulated[i].template_id = e.template_id;
    out_std[i].x = s[i].x; out_std[i].y = s[i].y; out_std[i].similarity = s[i].similarity; out_std[i].class_index = s[i].class_index; out_std[i].template_id = s[i].template_id;
  }
  return LMB200_OK;
}
