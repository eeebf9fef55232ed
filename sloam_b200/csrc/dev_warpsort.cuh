// dev_warpsort.cuh -- libstdc++'s std::sort replayed by a WARP (device only).
//
// The reference sorts with std::sort in two places whose results depend on the order it leaves
// exactly tied keys in (SURVEY B-3): the ground points of a polar cell by z (sloam.cpp:377-380)
// and the points of a tree vertex by x, by y and by z (trellis.cpp:71-82).  dev_stdsort.h
// replays libstdc++'s introsort operation for operation with one thread; this header does the
// same with the 32 lanes of a warp working together, for arrays of (float key, payload) records.
// Used by ground_cells_kernel<true> (k2_ground.cu) and vertex_replay_kernel (k3_trellis.cu);
// checked against std::sort through those kernels' parity tests and tests/test_stdsort.py
// (which covers the single-thread replay this one must agree with).
#pragma once

#include "common.cuh"
#include "dev_stdsort.h"

namespace sb {

struct SelKey { uint32_t z; uint32_t j; };

__device__ __forceinline__ bool sel_less(const SelKey &a, const SelKey &b) {
  return a.z < b.z || (a.z == b.z && a.j < b.j);
}

// z of a member from its order-preserving key (inverse of float_key)
__device__ __forceinline__ float key_to_float(uint32_t k) {
  return __uint_as_float((k & 0x80000000u) ? (k ^ 0x80000000u) : ~k);
}
// the reference's comparator p1.z < p2.z (sloam.cpp:378-380) on member records
// (the replay stores the float's own bits in SelKey::z)
struct MemberZLess {
  __device__ __forceinline__ bool operator()(const SelKey &a, const SelKey &b) const {
    return __uint_as_float(a.z) < __uint_as_float(b.z);
  }
};

// ---- std::sort replayed by a warp -----------------------------------------------------------
// The first r positions of libstdc++'s std::sort(cell, p1.z < p2.z), bit for bit what
// StdSortT::sort_prefix (dev_stdsort.h, one thread) produces, but with the warp working together:
//  * __unguarded_partition pairs the t-th element >= pivot from the left with the t-th element
//    <= pivot from the right and swaps them while the left one lies before the right one.  The
//    pairs do not depend on one another, so the lanes collect both stopper sequences 32
//    elements at a time (ballots), swap 32 pairs per step, and the cut is
//    min(l_t, r_{t-1}) at the first pair that has crossed (r_{t-1}: the scan from the left
//    stops on the element the previous swap put there at the latest).
//  * the final insertion sort is a STABLE sort of the prefix: rank by (z, current position).
//  * median-of-three and the depth-limit heapsort stay on one lane (a few operations / rare).
// `out` receives the first r elements in order; p is permuted in place (shared or global).
static __device__ int warp_partition(SelKey *p, int first, int last, int pivot, int *lq, int *rq) {
  const int lane = threadIdx.x & 31;
  const unsigned lt = (1u << lane) - 1u;
  const float piv = __uint_as_float(p[pivot].z);
  int Lp = first, Rp = last, nl = 0, nr = 0, prev_r = 0x7fffffff;
  for (;;) {
    while (nl < 32 && Lp < last) {
      const int i = Lp + lane;
      const bool stop = i < last && !(__uint_as_float(p[i].z) < piv);
      const unsigned b = __ballot_sync(kFull, stop);
      if (stop) lq[nl + __popc(b & lt)] = i;
      nl += __popc(b);
      Lp += 32;
    }
    while (nr < 32 && Rp > first) {
      const int i = Rp - 1 - lane;
      const bool stop = i >= first && !(piv < __uint_as_float(p[i].z));
      const unsigned b = __ballot_sync(kFull, stop);
      if (stop) rq[nr + __popc(b & lt)] = i;
      nr += __popc(b);
      Rp -= 32;
    }
    __syncwarp();
    const int cnt = min(min(nl, nr), 32);
    if (cnt == 0) {  // a side ran out (cannot happen behind the median-of-three sentinels)
      const int l0 = nl > 0 ? lq[0] : last;
      return min(l0, prev_r == 0x7fffffff ? last : prev_r);
    }
    int l = 0, rr = 0;
    bool ok = false;
    if (lane < cnt) { l = lq[lane]; rr = rq[lane]; ok = l < rr; }
    const unsigned fail = __ballot_sync(kFull, lane < cnt && !ok);
    const int nok = fail ? __ffs(fail) - 1 : cnt;
    if (lane < nok) { const SelKey a = p[l], b2 = p[rr]; p[l] = b2; p[rr] = a; }
    __syncwarp();
    if (fail) {
      const int lt_pos = lq[nok];
      const int rprev = nok > 0 ? rq[nok - 1] : prev_r;
      return min(lt_pos, rprev);
    }
    prev_r = rq[cnt - 1];
    // drop the consumed pairs from the queues
    const int l_keep = lane + cnt < nl ? lq[lane + cnt] : 0, l_keep2 = lane + 32 + cnt < nl ? lq[lane + 32 + cnt] : 0;
    const int r_keep = lane + cnt < nr ? rq[lane + cnt] : 0, r_keep2 = lane + 32 + cnt < nr ? rq[lane + 32 + cnt] : 0;
    __syncwarp();
    lq[lane] = l_keep; lq[lane + 32] = l_keep2;
    rq[lane] = r_keep; rq[lane + 32] = r_keep2;
    nl -= cnt; nr -= cnt;
    __syncwarp();
  }
}

static __device__ void warp_sort_prefix(SelKey *p, int n, int r, SelKey *out, int *lq, int *rq) {
  const int lane = threadIdx.x & 31;
  int bound = n;
  if (n > 16) {
    int lg = 0;
    for (int v = n; v > 1; v >>= 1) ++lg;
    int stack_first[32], stack_last[32], stack_depth[32], sp = 0;
    int first = 0, last = n, depth = lg * 2;
    const bool whole = r >= n;  // sort(): every range is needed
    StdSortT<SelKey, MemberZLess> one{p, MemberZLess{}};
    for (;;) {
      while (last - first > 16) {
        if (depth == 0) {
          if (lane == 0) one.heap_sort(first, last);
          __syncwarp();
          break;
        }
        --depth;
        const int mid = first + (last - first) / 2;
        if (lane == 0) one.move_median_to_first(first, first + 1, mid, last - 1);
        __syncwarp();
        const int cut = warp_partition(p, first + 1, last, first, lq, rq);
        if (!whole && cut >= r) {  // [cut, last) cannot influence positions < r
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
  }
  __syncwarp();
  // __final_insertion_sort == stable sort of [0, bound): rank by (z, position)
  for (int i0 = 0; i0 < bound; i0 += 32) {
    const int i = i0 + lane;
    if (i < bound) {
      const SelKey e = p[i];
      const float z = __uint_as_float(e.z);
      int rank = 0;
      for (int j = 0; j < bound; ++j) {
        const float zj = __uint_as_float(p[j].z);
        rank += (zj < z) || (zj == z && j < i);
      }
      if (rank < r) out[rank] = e;
    }
  }
  __syncwarp();
}

}  // namespace sb
