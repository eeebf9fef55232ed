// dev_stdsort.h -- libstdc++'s std::sort, operation for operation, on a permutation.
//
// Instance::computeVertexProperties (sloam/src/segmentation/trellis.cpp:71-82) sorts a
// vertex's points three times with std::sort (by x, by y, by z).  std::sort is not stable:
// for more than 16 elements libstdc++ runs an introsort (median-of-three quicksort, heapsort
// when the depth limit is hit, a final insertion sort), so when z values tie exactly the
// order of the tied points -- and with it the vertex radius and the feature order -- is
// whatever that particular algorithm leaves behind (SURVEY B-3).  To reproduce it the device
// runs the same algorithm: bits/stl_algo.h (__introsort_loop, __unguarded_partition_pivot,
// __move_median_to_first, __final_insertion_sort) and bits/stl_heap.h (__adjust_heap,
// __push_heap, __pop_heap, __make_heap, __sort_heap), unchanged between GCC 5 and 14.
// It sorts indices p[0..n) by key[p[i]] with the comparator key[a] < key[b]; moving an index
// is moving the point.  Host + device (tests/stdsort_test.cpp checks it against std::sort).
#ifndef SLOAM_B200_DEV_STDSORT_H
#define SLOAM_B200_DEV_STDSORT_H

#include <cstdint>

#ifndef SLOAM_HD_FN
#if defined(__CUDACC__)
#define SLOAM_HD_FN __host__ __device__ __forceinline__
#else
#define SLOAM_HD_FN inline
#endif
#endif

namespace sb {

// Elements of type Idx compared by a functor (a, b) -> bool.
template <class Idx, class Less>
struct StdSortT {
  Idx *p;
  Less less;
  SLOAM_HD_FN void swap(int i, int j) { const Idx t = p[i]; p[i] = p[j]; p[j] = t; }

  // ---- stl_heap.h ----
  SLOAM_HD_FN void push_heap(int first, int hole, int top, Idx value) {
    int parent = (hole - 1) / 2;
    while (hole > top && less(p[first + parent], value)) {
      p[first + hole] = p[first + parent];
      hole = parent;
      parent = (hole - 1) / 2;
    }
    p[first + hole] = value;
  }
  SLOAM_HD_FN void adjust_heap(int first, int hole, int len, Idx value) {
    const int top = hole;
    int second = hole;
    while (second < (len - 1) / 2) {
      second = 2 * (second + 1);
      if (less(p[first + second], p[first + (second - 1)])) --second;
      p[first + hole] = p[first + second];
      hole = second;
    }
    if ((len & 1) == 0 && second == (len - 2) / 2) {
      second = 2 * (second + 1);
      p[first + hole] = p[first + (second - 1)];
      hole = second - 1;
    }
    push_heap(first, hole, top, value);
  }
  SLOAM_HD_FN void heap_sort(int first, int last) {  // __partial_sort(first, last, last)
    const int len = last - first;
    if (len >= 2) {  // __make_heap
      int parent = (len - 2) / 2;
      for (;;) {
        const Idx value = p[first + parent];
        adjust_heap(first, parent, len, value);
        if (parent == 0) break;
        --parent;
      }
    }
    while (last - first > 1) {  // __sort_heap: __pop_heap(first, last - 1, last - 1)
      --last;
      const Idx value = p[last];
      p[last] = p[first];
      adjust_heap(first, 0, last - first, value);
    }
  }

  // ---- stl_algo.h ----
  SLOAM_HD_FN void move_median_to_first(int result, int a, int b, int c) {
    if (less(p[a], p[b])) {
      if (less(p[b], p[c])) swap(result, b);
      else if (less(p[a], p[c])) swap(result, c);
      else swap(result, a);
    } else if (less(p[a], p[c])) swap(result, a);
    else if (less(p[b], p[c])) swap(result, c);
    else swap(result, b);
  }
  SLOAM_HD_FN int unguarded_partition(int first, int last, int pivot) {
    for (;;) {
      while (less(p[first], p[pivot])) ++first;
      --last;
      while (less(p[pivot], p[last])) --last;
      if (!(first < last)) return first;
      swap(first, last);
      ++first;
    }
  }
  SLOAM_HD_FN void unguarded_linear_insert(int last) {
    const Idx val = p[last];
    int next = last - 1;
    while (less(val, p[next])) {
      p[last] = p[next];
      last = next;
      --next;
    }
    p[last] = val;
  }
  SLOAM_HD_FN void insertion_sort(int first, int last) {
    if (first == last) return;
    for (int i = first + 1; i != last; ++i) {
      if (less(p[i], p[first])) {
        const Idx val = p[i];
        for (int j = i; j > first; --j) p[j] = p[j - 1];  // move_backward(first, i, i + 1)
        p[first] = val;
      } else {
        unguarded_linear_insert(i);
      }
    }
  }
  // std::sort(p, p + n, less); n <= 2^15.  The recursion of __introsort_loop (right part
  // recursive, left part iterative) runs on an explicit stack: at most depth_limit frames.
  SLOAM_HD_FN void sort(int n) {
    if (n <= 0) return;
    int lg = 0;
    for (int v = n; v > 1; v >>= 1) ++lg;  // std::__lg
    int stack_first[32], stack_last[32], stack_depth[32], sp = 0;
    int first = 0, last = n, depth = lg * 2;
    for (;;) {
      while (last - first > 16) {
        if (depth == 0) { heap_sort(first, last); break; }
        --depth;
        const int mid = first + (last - first) / 2;
        move_median_to_first(first, first + 1, mid, last - 1);
        const int cut = unguarded_partition(first + 1, last, first);
        // __introsort_loop(cut, last, depth) happens BEFORE the loop continues on [first, cut)
        stack_first[sp] = first; stack_last[sp] = cut; stack_depth[sp] = depth; ++sp;
        first = cut;
      }
      if (sp == 0) break;
      --sp;
      first = stack_first[sp]; last = stack_last[sp]; depth = stack_depth[sp];
    }
    // __final_insertion_sort
    if (n > 16) {
      insertion_sort(0, 16);
      for (int i = 16; i != n; ++i) unguarded_linear_insert(i);
    } else {
      insertion_sort(0, n);
    }
  }
  // The first r positions of std::sort's result, exactly, without finishing the rest: a
  // partition step only permutes its own range, so a right-hand range that starts at or
  // beyond r can be left alone (its elements are >= everything before it and the final
  // insertion pass never moves them in front of it).  Positions >= the returned bound are
  // NOT in their final order.
  SLOAM_HD_FN int sort_prefix(int n, int r) {
    if (n <= 0) return 0;
    if (r >= n) { sort(n); return n; }
    int lg = 0;
    for (int v = n; v > 1; v >>= 1) ++lg;
    int stack_first[32], stack_last[32], stack_depth[32], sp = 0;
    int first = 0, last = n, depth = lg * 2, bound = n;
    for (;;) {
      while (last - first > 16) {
        if (depth == 0) { heap_sort(first, last); break; }
        --depth;
        const int mid = first + (last - first) / 2;
        move_median_to_first(first, first + 1, mid, last - 1);
        const int cut = unguarded_partition(first + 1, last, first);
        if (cut >= r) {  // [cut, last) cannot influence positions < r
          if (cut < bound) bound = cut;
          last = cut;
          continue;
        }
        stack_first[sp] = first; stack_last[sp] = cut; stack_depth[sp] = depth; ++sp;
        first = cut;
      }
      if (sp == 0) break;
      --sp;
      first = stack_first[sp]; last = stack_last[sp]; depth = stack_depth[sp];
    }
    if (n > 16) {
      insertion_sort(0, bound < 16 ? bound : 16);
      for (int i = 16; i < bound; ++i) unguarded_linear_insert(i);
    } else {
      insertion_sort(0, n);
    }
    return bound;
  }
};

// indices compared through a key array: key[a] < key[b]
struct KeyLess {
  const float *key;
  template <class Idx> SLOAM_HD_FN bool operator()(Idx a, Idx b) const { return key[a] < key[b]; }
};
template <class Idx>
struct StdSort : StdSortT<Idx, KeyLess> {
  SLOAM_HD_FN StdSort(Idx *p_, const float *key_) : StdSortT<Idx, KeyLess>{p_, KeyLess{key_}} {}
};

// (key, index) pairs packed in 64 bits, key = float bits in the high word: one load per
// element instead of two dependent ones
struct PackedLess {
  SLOAM_HD_FN static float key_of(unsigned long long e) {
    const unsigned u = (unsigned)(e >> 32);
    float f;
#if defined(__CUDA_ARCH__)
    f = __uint_as_float(u);
#else
    __builtin_memcpy(&f, &u, 4);
#endif
    return f;
  }
  SLOAM_HD_FN bool operator()(unsigned long long a, unsigned long long b) const { return key_of(a) < key_of(b); }
};
using StdSortPacked = StdSortT<unsigned long long, PackedLess>;

}  // namespace sb
#endif
