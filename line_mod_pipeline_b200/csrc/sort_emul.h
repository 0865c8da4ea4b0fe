// sort_emul.h — the epilogue of Detector::match, std::sort(matches) + std::unique(matches), reproduced ELEMENT FOR ELEMENT
// on the device (upstream: opencv_contrib rgbd/linemod.cpp, Detector::match, last three lines).
//
// Match::operator< orders by (similarity desc, template_id asc) only, so matches that tie on both keep whatever order the
// sort ALGORITHM leaves them in — and std::unique then removes only ADJACENT equal (x, y, similarity, class_id) records.
// The reference's result therefore depends on libstdc++'s std::sort, an introsort that is not stable.  To return the
// reference's sequence bit for bit without a host round trip, this header restates that algorithm (GNU libstdc++
// bits/stl_algo.h / bits/stl_heap.h as shipped with GCC 4.9 .. 14: introsort loop with a 2*floor(log2 n) depth limit,
// median of {first+1, mid, last-1} moved to first, unguarded Hoare partition, heap sort when the depth limit is hit,
// final insertion sort with the 16-element threshold) on an array of 64-bit keys with a 32-bit payload.  The sequence of
// comparisons and moves is the same as std::sort's on the Match records, so the resulting permutation is identical
// (tests/test_capi_host.py checks it against std::sort itself on tie-heavy inputs, including the heap-sort fallback).
//
// key(m) = (0xFFFFFFFF - bits(similarity)) << 32 | template_id :  a < b  <=>  key(a) < key(b)   (similarity >= 0, id >= 0).
#pragma once
#include <stdint.h>

#if defined(__CUDACC__)
#define LM_HD __host__ __device__ __forceinline__
#else
#define LM_HD inline
#endif

namespace lmsort {

struct Arr {              // keys and payloads move together: "one element" of the emulated std::sort
  uint64_t* k;
  uint32_t* v;
};
struct Elem { uint64_t k; uint32_t v; };

LM_HD Elem get(const Arr& a, int i) { Elem e; e.k = a.k[i]; e.v = a.v[i]; return e; }
LM_HD void put(const Arr& a, int i, const Elem& e) { a.k[i] = e.k; a.v[i] = e.v; }
LM_HD void move_(const Arr& a, int dst, int src) { a.k[dst] = a.k[src]; a.v[dst] = a.v[src]; }
LM_HD void swap_(const Arr& a, int i, int j) { Elem t = get(a, i); move_(a, i, j); put(a, j, t); }

LM_HD uint64_t match_key(float similarity, int template_id) {
  union { float f; uint32_t u; } c; c.f = similarity;
  return ((uint64_t)(0xFFFFFFFFu - c.u) << 32) | (uint32_t)template_id;
}

// std::__push_heap with a value compared through operator<
LM_HD void push_heap_(const Arr& a, int first, int hole, int top, const Elem& value) {
  int parent = (hole - 1) / 2;
  while (hole > top && a.k[first + parent] < value.k) {
    move_(a, first + hole, first + parent);
    hole = parent;
    parent = (hole - 1) / 2;
  }
  put(a, first + hole, value);
}

// std::__adjust_heap
LM_HD void adjust_heap_(const Arr& a, int first, int hole, int len, const Elem& value) {
  const int top = hole;
  int child = hole;
  while (child < (len - 1) / 2) {
    child = 2 * (child + 1);
    if (a.k[first + child] < a.k[first + child - 1]) child--;
    move_(a, first + hole, first + child);
    hole = child;
  }
  if ((len & 1) == 0 && child == (len - 2) / 2) {
    child = 2 * (child + 1);
    move_(a, first + hole, first + child - 1);
    hole = child - 1;
  }
  push_heap_(a, first, hole, top, value);
}

// std::__partial_sort(first, last, last): __heap_select over the whole range (= make_heap) + __sort_heap
LM_HD void heap_sort_(const Arr& a, int first, int last) {
  const int len = last - first;
  if (len >= 2) {
    int parent = (len - 2) / 2;
    for (;;) {
      const Elem value = get(a, first + parent);
      adjust_heap_(a, first, parent, len, value);
      if (parent == 0) break;
      parent--;
    }
  }
  int l = last;
  while (l - first > 1) {
    --l;
    const Elem value = get(a, l);           // std::__pop_heap(first, l, l)
    move_(a, l, first);
    adjust_heap_(a, first, 0, l - first, value);
  }
}

// std::__move_median_to_first(result, a, b, c)
LM_HD void move_median_to_first_(const Arr& s, int result, int a, int b, int c) {
  const uint64_t ka = s.k[a], kb = s.k[b], kc = s.k[c];
  if (ka < kb) {
    if (kb < kc) swap_(s, result, b);
    else if (ka < kc) swap_(s, result, c);
    else swap_(s, result, a);
  } else if (ka < kc) swap_(s, result, a);
  else if (kb < kc) swap_(s, result, c);
  else swap_(s, result, b);
}

// std::__unguarded_partition(first, last, pivot)
LM_HD int unguarded_partition_(const Arr& s, int first, int last, int pivot) {
  for (;;) {
    const uint64_t kp = s.k[pivot];
    while (s.k[first] < kp) ++first;
    --last;
    while (kp < s.k[last]) --last;
    if (!(first < last)) return first;
    swap_(s, first, last);
    ++first;
  }
}

// std::__unguarded_linear_insert(last)
LM_HD void unguarded_linear_insert_(const Arr& s, int last) {
  const Elem val = get(s, last);
  int next = last - 1;
  while (val.k < s.k[next]) {
    move_(s, last, next);
    last = next;
    --next;
  }
  put(s, last, val);
}

// std::__insertion_sort(first, last)
LM_HD void insertion_sort_(const Arr& s, int first, int last) {
  if (first == last) return;
  for (int i = first + 1; i != last; ++i) {
    if (s.k[i] < s.k[first]) {
      const Elem val = get(s, i);
      for (int j = i; j > first; --j) move_(s, j, j - 1);   // std::move_backward(first, i, i + 1)
      put(s, first, val);
    } else {
      unguarded_linear_insert_(s, i);
    }
  }
}

// std::sort(first, last) on n elements; the recursion of __introsort_loop (it recurses into the RIGHT part and loops on
// the left one) is unrolled on an explicit stack of (cut, last, depth_limit) records: at most one per depth level.
LM_HD void std_sort(const Arr& s, int n) {
  if (n <= 0) return;
  int lg = 0;
  for (unsigned m = (unsigned)n; m > 1; m >>= 1) ++lg;    // std::__lg(n)
  int st_first[64], st_last[64], st_depth[64];
  int sp = 0;
  st_first[0] = 0; st_last[0] = n; st_depth[0] = 2 * lg; sp = 1;
  while (sp > 0) {
    --sp;
    const int first = st_first[sp];
    int last = st_last[sp], depth = st_depth[sp];
    // one activation of __introsort_loop(first, last, depth).  Its recursive calls (right parts) must run BEFORE the
    // loop continues on the left part only in terms of data dependencies: the two ranges are disjoint, so deferring
    // the right parts to the stack leaves every comparison and move unchanged.
    while (last - first > 16) {
      if (depth == 0) {
        heap_sort_(s, first, last);
        break;
      }
      --depth;
      const int mid = first + (last - first) / 2;
      move_median_to_first_(s, first, first + 1, mid, last - 1);
      const int cut = unguarded_partition_(s, first + 1, last, first);
      st_first[sp] = cut; st_last[sp] = last; st_depth[sp] = depth; ++sp;
      last = cut;
    }
  }
  // std::__final_insertion_sort
  if (n > 16) {
    insertion_sort_(s, 0, 16);
    for (int i = 16; i != n; ++i) unguarded_linear_insert_(s, i);
  } else {
    insertion_sort_(s, 0, n);
  }
}

}  // namespace lmsort
