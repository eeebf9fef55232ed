// k3_trellis.cu -- stage a6 + a7: tree instance detection.
//
// Replaces Instance::computeGraph (sloam/src/segmentation/trellis.cpp:134-140):
//   findClusters -> PCL OrganizedConnectedComponentSegmentation with an
//   EuclideanClusterComparator(1.0 m) over the organized H x W tree cloud
//   (trellis.cpp:15-29), i.e. connected components of the grid graph whose
//   edges are (left, up) neighbour pairs with ||pa - pb|| < threshold; the
//   PCL label of a component is its rank by the raster index of its first
//   pixel.
//   findTrees -> per cluster with > 80 points, per scan line from the bottom
//   row up, a TreeVertex from the cluster's points of that row
//   (trellis.cpp:104-132, computeVertexProperties :63-102).
//
// GPU formulation: lock-free union-find over pixels (init / merge / flatten),
// where the smaller raster index always wins a union so that every root IS the
// first pixel of its component; component size, column extent and last row are
// accumulated at the root with warp-aggregated atomics; a per-keyframe planning
// CTA sorts the big components (= PCL label order), ranks them and emits
// (cluster, row) work items; one warp per work item builds the vertex with
// rank-based order statistics in shared memory.
#include "common.cuh"
#include "dev_stdsort.h"

namespace sb {

#ifndef SLOAM_VTX_WARPS
#define SLOAM_VTX_WARPS 8
#endif
constexpr int kVtxWarps = SLOAM_VTX_WARPS;
constexpr int kVtxCap = 128;   // members per (cluster,row) handled by the warp path
constexpr int kInvalid = -1;

__device__ __forceinline__ int uf_find(const int32_t *parent, int x) {
  int p = *((volatile const int32_t *)&parent[x]);
  while (p != x) {
    x = p;
    p = *((volatile const int32_t *)&parent[x]);
  }
  return x;
}
__device__ __forceinline__ void uf_union(int32_t *parent, int a, int b) {
  while (true) {
    a = uf_find(parent, a);
    b = uf_find(parent, b);
    if (a == b) return;
    if (a < b) { const int t = a; a = b; b = t; }
    const int old = atomicMin(&parent[a], b);  // attach the larger root under the smaller
    if (old == a) return;
    a = old;
  }
}

// tree_bits: one bit per pixel, [K][Nw] words (Nw = ceil(N / 32)).  Bit set = the pixel may
// hold a tree point (the split kernel sets it for mask == 255; for caller-supplied clouds it
// is isfinite(x)).  Pixels whose bit is clear are never read by the kernels below -- in the
// fused pipeline their tree point, parent and flag entries are not even written.
__device__ __forceinline__ bool tree_bit(const uint32_t *__restrict__ bits_k, int i) {
  return i >= 0 && ((bits_k[i >> 5] >> (i & 31)) & 1u);
}

__global__ void tree_bits_kernel(const DevParams *__restrict__ dp, const sloam_point *__restrict__ tree,
                                 uint32_t *__restrict__ bits) {
  const int N = dp->N, Nw = (N + 31) >> 5;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  const int k = blockIdx.y;
  const bool v = i < N && isfinite(tree[(size_t)k * N + i].x);
  const unsigned b = __ballot_sync(kFull, v);
  if ((threadIdx.x & 31) == 0 && i < N) bits[(size_t)k * Nw + (i >> 5)] = b;
}

// dense organized cloud from the sparse one: NaN points (intensity 0) where the bit is clear
__global__ void tree_fill_kernel(const DevParams *__restrict__ dp, const uint32_t *__restrict__ bits,
                                 sloam_point *__restrict__ tree) {
  const int N = dp->N, Nw = (N + 31) >> 5;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  const int k = blockIdx.y;
  if (i >= N) return;
  if (!((bits[(size_t)k * Nw + (i >> 5)] >> (i & 31)) & 1u)) {
    const float qnan = __int_as_float(0x7fc00000);
    st_point(tree + (size_t)k * N + i, sloam_point{qnan, qnan, qnan, 0.f});
  }
}

// The three connected-component passes below walk the NON-ZERO bit words, not the pixels
// (a forest scan is ~90 % non-tree pixels).  tree_words_kernel lists the non-zero words of
// the batch once; each pass then hands one listed word to a warp, 32 lanes = its 32 pixels,
// so the work is balanced no matter how the trees cluster in the image.
__global__ void tree_words_kernel(const DevParams *__restrict__ dp, int K, const uint32_t *__restrict__ bits,
                                  int2 *__restrict__ list, int32_t *__restrict__ n_list) {
  const int Nw = (dp->N + 31) >> 5;
  const long long total_words = (long long)K * Nw;
  const int lane = threadIdx.x & 31;
  const long long warp0 = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const long long nwarps = (long long)gridDim.x * (blockDim.x >> 5);
  for (long long base = warp0 * 32; base < total_words; base += nwarps * 32) {
    const uint32_t word = base + lane < total_words ? bits[base + lane] : 0u;
    const bool nz = word != 0u;
    const unsigned m = __ballot_sync(kFull, nz);
    if (m == 0u) continue;
    int at = 0;
    if (lane == 0) at = atomicAdd(n_list, __popc(m));
    at = __shfl_sync(kFull, at, 0);
    if (nz) {  // entry = (keyframe << 16 | word index in the keyframe, the word itself)
      const long long gw = base + lane;
      const int k = (int)(gw / Nw);
      list[at + __popc(m & ((1u << lane) - 1u))] = make_int2((k << 16) | (int)(gw - (long long)k * Nw), (int)word);
    }
  }
}

template <class F>
__device__ __forceinline__ void for_each_tree_word(const int2 *__restrict__ list,
                                                   const int32_t *__restrict__ n_list, F body) {
  const int lane = threadIdx.x & 31;
  const int n = *n_list;
  const int warp0 = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int nwarps = gridDim.x * (blockDim.x >> 5);
  // the next list entry is fetched one iteration ahead: one memory round trip less on the
  // dependent chain entry -> bits / parents -> points of every word
  int2 e = warp0 < n ? list[warp0] : make_int2(0, 0);
  for (int idx = warp0; idx < n; idx += nwarps) {
    const int2 cur = e;
    if (idx + nwarps < n) e = list[idx + nwarps];
    body((int)((unsigned)cur.x >> 16), (cur.x & 0xFFFF) * 32 + lane, (uint32_t)cur.y, idx);
  }
}

// ---- 1. init: link every valid pixel to its left (else up) neighbour ------
#ifndef SLOAM_CCINIT_MIN
#define SLOAM_CCINIT_MIN 8  // the persistent grid is 8 CTAs per SM: keep all of them resident (32 registers)
#endif
__global__ void __launch_bounds__(256, SLOAM_CCINIT_MIN)
cc_init_kernel(const DevParams *__restrict__ dp, int K, const sloam_point *__restrict__ tree,
               const uint32_t *__restrict__ bits, const int2 *__restrict__ wlist,
               const int32_t *__restrict__ n_wlist, int32_t *__restrict__ parent, uint32_t *__restrict__ mflags,
               int32_t *__restrict__ csize, int32_t *__restrict__ cmin, int32_t *__restrict__ cmax,
               int32_t *__restrict__ rmax) {
  const int N = dp->N, W = dp->p.img_w, Nw = (N + 31) >> 5;
  const float cut = dp->cluster_sq_cut;  // dist < cluster_dist_thresh  <=>  squared dist < cut
  const unsigned magic_w = dp->magic_w;
  const int lane = threadIdx.x & 31;
  for_each_tree_word(wlist, n_wlist, [&](int k, int i, uint32_t word, int idx) {
    const uint32_t *bk = bits + (size_t)k * Nw;
    const size_t g = (size_t)k * N + i;
    const int row = fast_div_w(i, magic_w), col = i - row * W;
    const bool bit = (word >> lane) & 1u;  // clear for the padding lanes of the last word
    // The four points (self, left, up, up-left) are loaded together, guarded by their bits
    // only, so that the loads are in flight at the same time instead of one after another.
    const int u = i - W;
    const uint32_t wu = (bit && row > 0) ? bk[u >> 5] : 0u;
    const bool bit_l = bit && col > 0 && (lane > 0 ? ((word >> (lane - 1)) & 1u) != 0u : tree_bit(bk, i - 1));
    const bool bit_u = (wu >> (u & 31)) & 1u;
    const bool bit_ul = bit_u && bit_l && ((u & 31) ? ((wu >> ((u & 31) - 1)) & 1u) != 0u : tree_bit(bk, u - 1));
    sloam_point p{0.f, 0.f, 0.f, 0.f}, ql = p, qu = p, qul = p;
    if (bit) p = ld_point(tree + g);
    if (bit_l) ql = ld_point(tree + g - 1);
    if (bit_u) qu = ld_point(tree + g - W);
    if (bit_ul) qul = ld_point(tree + g - W - 1);
    // PCL skips a pixel iff !isfinite(x); EuclideanClusterComparator::compare is
    // dist < threshold in float (NaN compares false)
    const bool valid = bit && isfinite(p.x);
    bool left_ok = false, up_ok = false, upleft_ok = false;  // upleft_ok: (i-W) -- (i-W-1)
    if (valid) {
      left_ok = bit_l && sqnorm3f(p.x - ql.x, p.y - ql.y, p.z - ql.z) < cut;
      up_ok = bit_u && sqnorm3f(p.x - qu.x, p.y - qu.y, p.z - qu.z) < cut;
      upleft_ok = up_ok && left_ok && bit_ul && sqnorm3f(qu.x - qul.x, qu.y - qul.y, qu.z - qul.z) < cut;
    }
    // is the left neighbour connected to ITS upper neighbour?
    int left_up = __shfl_up_sync(kFull, up_ok ? 1 : 0, 1);
    if (lane == 0) {
      left_up = 0;
      if (left_ok && row > 0 && tree_bit(bk, i - 1 - W)) {
        const sloam_point a = ld_point(tree + g - 1), b = ld_point(tree + g - 1 - W);
        left_up = sqnorm3f(a.x - b.x, a.y - b.y, a.z - b.z) < cut;
      }
    }
    // run starts inside the word: link to the start of the row run (short find chains)
    const unsigned starts = __ballot_sync(kFull, !left_ok);
    if (i < N) {
      if (valid) {
        int par;
        if (left_ok) {
          const unsigned s = starts & ((2u << lane) - 1u);   // starts at or below this lane
          par = s ? i - (lane - (31 - __clz(s))) : i - (lane + 1);
        } else {
          par = up_ok ? i - W : i;
        }
        parent[g] = par;
        if (par == i) {  // only a pixel that starts as its own parent can end up a root
          csize[g] = 0;
          cmin[g] = col; cmax[g] = col; rmax[g] = row;
        }
      } else {
        parent[g] = kInvalid;
      }
    }
    // A union with the upper neighbour is only needed when it is not implied by
    // i ~ i-1 (row link), i-1 ~ i-1-W (left neighbour's own up link) and i-W ~ i-W-1.
    const unsigned mf = __ballot_sync(kFull, left_ok && up_ok && !(left_up && upleft_ok));
    if (lane == 0) mflags[idx] = mf;  // one bit per pixel of the word, indexed like the word list
  });
}

// ---- 2. merge: pixels linked left that are also connected upwards ---------
__global__ void __launch_bounds__(256)
cc_merge_kernel(const DevParams *__restrict__ dp, int K, const uint32_t *__restrict__ bits,
                const int2 *__restrict__ wlist, const int32_t *__restrict__ n_wlist,
                const uint32_t *__restrict__ mflags, int32_t *__restrict__ parent) {
  const int N = dp->N, W = dp->p.img_w;
  const int n = *n_wlist;
  // thread per listed word: the few pixels whose flag is set are merged by that thread
  for (int idx = blockIdx.x * blockDim.x + threadIdx.x; idx < n; idx += gridDim.x * blockDim.x) {
    uint32_t mf = mflags[idx];
    if (mf == 0u) continue;
    const int2 e = wlist[idx];
    const int k = (int)((unsigned)e.x >> 16), i0 = (e.x & 0xFFFF) * 32;
    while (mf) {
      const int i = i0 + __ffs(mf) - 1;
      mf &= mf - 1;
      uf_union(parent + (size_t)k * N, i, i - W);
    }
  }
}

// ---- 3. flatten + statistics at the root ----------------------------------
__global__ void __launch_bounds__(256)
cc_flatten_kernel(const DevParams *__restrict__ dp, int K, const uint32_t *__restrict__ bits,
                  const int2 *__restrict__ wlist, const int32_t *__restrict__ n_wlist,
                  int32_t *__restrict__ parent, int32_t *__restrict__ csize, int32_t *__restrict__ cmin,
                  int32_t *__restrict__ cmax, int32_t *__restrict__ rmax, int32_t *__restrict__ row_roots,
                  int32_t *__restrict__ n_roots, int32_t *__restrict__ big_roots,
                  int32_t *__restrict__ n_big, int32_t *__restrict__ kf_flags, uint32_t *__restrict__ root_bits) {
  const int N = dp->N, W = dp->p.img_w, H = dp->p.img_h;
  const int min_pts = dp->p.min_cluster_points, T = dp->p.max_trees;
  const unsigned magic_w = dp->magic_w;
  const int lane = threadIdx.x & 31;
  for_each_tree_word(wlist, n_wlist, [&](int k, int i, uint32_t word, int) {
    const size_t g = (size_t)k * N + i;
    int root = kInvalid;
    const int par0 = ((word >> lane) & 1u) ? parent[g] : kInvalid;
    if (par0 != kInvalid) {
      root = par0 == i ? i : uf_find(parent + (size_t)k * N, par0);
      parent[g] = root;
    }
    const int row = fast_div_w(i, magic_w), col = i - row * W;
    // Consecutive lanes are consecutive pixels of a row, so the members of a component come
    // in runs: a run = maximal stretch of lanes with the same (root, row).  One lane per run
    // (its head) issues the atomics for the whole run.
    const long long key = root == kInvalid ? -1ll - lane : ((long long)root * 4096ll + row);
    const long long prev = __shfl_up_sync(kFull, key, 1);
    const bool head = root != kInvalid && (lane == 0 || prev != key);
    const unsigned heads = __ballot_sync(kFull, head || root == kInvalid);
    if (head) {
      // run length = distance to the next head / invalid lane (or the end of the word)
      const unsigned above = heads & ~((2u << lane) - 1u);
      const int len = (above ? __ffs(above) - 1 : 32) - lane;
      const size_t r = (size_t)k * N + root;
      const int old = atomicAdd(&csize[r], len);
      if (old <= min_pts && old + len > min_pts) {  // exactly one run sees the crossing
        const int slot = atomicAdd(&n_big[k], 1);
        if (slot < T) big_roots[(size_t)k * T + slot] = root;
        else atomicOr(&kf_flags[k], 1);
      }
      // extents of the component (monotone, so the racy pre-checks are safe)
      if (col < cmin[r]) atomicMin(&cmin[r], col);
      if (col + len - 1 > cmax[r]) atomicMax(&cmax[r], col + len - 1);
      if (row > rmax[r]) atomicMax(&rmax[r], row);
    }
    if (root == i) {
      atomicAdd(&row_roots[(size_t)k * H + row], 1);
      atomicAdd(&n_roots[k], 1);
      // one bit per pixel: is a component root (cc_plan counts the roots before a given one);
      // roots are few, the words were zeroed before the launch
      atomicOr(&root_bits[(size_t)k * ((N + 31) >> 5) + (i >> 5)], 1u << (i & 31));
    }
  });
}

// ---- 4. plan: sort big clusters, rank them, emit work items ----------------
__global__ void cc_plan_kernel(const DevParams *__restrict__ dp, const uint32_t *__restrict__ root_bits,
                               const int32_t *__restrict__ parent,
                               const int32_t *__restrict__ cmin, const int32_t *__restrict__ cmax,
                               const int32_t *__restrict__ rmax, const int32_t *__restrict__ row_roots,
                               int32_t *__restrict__ big_roots, int32_t *__restrict__ n_big,
                               int32_t *__restrict__ big_rank, int32_t *__restrict__ bbox,
                               int32_t *__restrict__ vwork, int32_t *__restrict__ n_vwork) {
  extern __shared__ int32_t sm[];
  const int N = dp->N, W = dp->p.img_w, H = dp->p.img_h, T = dp->p.max_trees;
  const int k = blockIdx.x;
  int32_t *s_roots = sm;            // [T]
  int32_t *s_sorted = sm + T;       // [T]
  int32_t *s_rowpre = sm + 2 * T;   // [H+1]
  const int nb = min(n_big[k], T);
  for (int s = threadIdx.x; s < nb; s += blockDim.x) s_roots[s] = big_roots[(size_t)k * T + s];
  if (threadIdx.x == 0) {
    int acc = 0;
    for (int r = 0; r < H; ++r) { s_rowpre[r] = acc; acc += row_roots[(size_t)k * H + r]; }
    s_rowpre[H] = acc;
  }
  __syncthreads();
  for (int s = threadIdx.x; s < nb; s += blockDim.x) {  // rank sort (roots are distinct)
    const int v = s_roots[s];
    int r = 0;
    for (int j = 0; j < nb; ++j) r += s_roots[j] < v;
    s_sorted[r] = v;
  }
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
  for (int s = warp; s < nb; s += nwarps) {
    const int root = s_sorted[s];
    const int row = root / W, col = root - row * W;
    // PCL label = number of component roots before this one in raster order: the roots of
    // the rows above (s_rowpre) + the set bits of root_bits in [row * W, root)
    int cnt = 0;
    const uint32_t *rbk = root_bits + (size_t)k * ((N + 31) >> 5);
    const int a = row * W, b = root;  // pixel range [a, b)
    for (int wi = (a >> 5) + lane; wi <= (b >> 5); wi += 32) {
      uint32_t v = rbk[wi];
      if (wi == (a >> 5)) v &= 0xFFFFFFFFu << (a & 31);
      if (wi == (b >> 5)) v &= (b & 31) ? (0xFFFFFFFFu >> (32 - (b & 31))) : 0u;
      cnt += __popc(v);
    }
    cnt = warp_sum(cnt);
    if (lane == 0) {
      const size_t r = (size_t)k * N + root;
      big_roots[(size_t)k * T + s] = root;
      big_rank[(size_t)k * T + s] = s_rowpre[row] + cnt;
      int32_t *bb = bbox + ((size_t)k * T + s) * 4;
      const int r1 = rmax[r];
      bb[0] = cmin[r]; bb[1] = cmax[r]; bb[2] = row; bb[3] = r1;
      const int nrows = r1 - row + 1;
      const int base = atomicAdd(n_vwork, nrows);
      for (int q = 0; q < nrows; ++q) vwork[base + q] = vw_pack(dp, k, s, row + q);
    }
  }
  if (threadIdx.x == 0) n_big[k] = nb;
}

// ---- 5. one vertex per (cluster, row) --------------------------------------
struct VtxSmem {
  float x[kVtxCap], y[kVtxCap], z[kVtxCap], w[kVtxCap];
  int16_t col[kVtxCap];
  int16_t order[kVtxCap];
};

// lexicographic (z, y, x, col): the order the reference's three std::sort calls
// produce (SURVEY B-3); col is unique so this is a total order
__device__ __forceinline__ bool key_less(float za, float ya, float xa, int ca, float zb, float yb,
                                         float xb, int cb) {
  if (za != zb) return za < zb;
  if (ya != yb) return ya < yb;
  if (xa != xb) return xa < xb;
  return ca < cb;
}

// The rare tie path: computeVertexProperties' std::sort by x, by y, by z on members that
// start in column order, replayed by ONE thread (the algorithm is sequential).
template <class Idx>
__device__ __noinline__ void replay_three_sorts(Idx *order, int n, const float *x, const float *y, const float *z) {
  for (int m = 0; m < n; ++m) order[m] = (Idx)m;
  const float *keys[3] = {x, y, z};
  for (int a = 0; a < 3; ++a) {  // one instance of the sort code
    StdSort<Idx> srt(order, keys[a]);
    srt.sort(n);
  }
}
// Same with (key, index) packed in one 64-bit word per element (one shared-memory load per
// access instead of two dependent ones); `pack` is scratch for n words.  A sort whose keys
// are all distinct has one possible result, so the replay only has to start at the first axis
// that has ties after it: `first` = 0 (x: start from column order), 1 (y: `init` = the members
// in x order) or 2 (z: `init` = the members in y order).
__device__ __noinline__ void replay_sorts_packed(int16_t *order, unsigned long long *pack, int n, int first,
                                                 const int16_t *init, const float *x, const float *y,
                                                 const float *z) {
  const float *keys[3] = {x, y, z};
  for (int m = 0; m < n; ++m) {
    const unsigned idx = first == 0 ? (unsigned)m : (unsigned)init[m];
    pack[m] = ((unsigned long long)__float_as_uint(keys[first][idx]) << 32) | idx;
  }
  for (int a = first; a < 3; ++a) {
    if (a > first)
      for (int m = 0; m < n; ++m) {
        const unsigned idx = (unsigned)(pack[m] & 0xFFFFFFFFull);
        pack[m] = ((unsigned long long)__float_as_uint(keys[a][idx]) << 32) | idx;
      }
    StdSortPacked srt{pack, PackedLess{}};
    srt.sort(n);
  }
  for (int m = 0; m < n; ++m) order[m] = (int16_t)(pack[m] & 0xFFFFFFFFull);
}

// REPLAY = false: returns true when the item has exact z ties among more than 16 points and
// must be redone by the REPLAY = true instance (nothing was written)
template <bool REPLAY>
__device__ bool build_vertex(const DevParams *dp, VtxSmem &s, unsigned long long *pack, int16_t *perm, int n,
                             int row, sloam_vertex *out, sloam_point *pool, int32_t *pool_count) {
  const int lane = threadIdx.x & 31;
  const int middle = (int)(n / 2.0);  // trellis.cpp:66
  // Order statistics by counting: member m counts the members strictly below it on each
  // axis.  Without ties that count is the rank: the member whose count equals `middle` holds
  // the median, and the z counts are the z order.  Exact ties are rare (float coordinates):
  // a z tie is detected by counting equal members and handled by the full-key pass below; a
  // tie group straddling the median leaves no member with count == middle, which is caught
  // after the loop and redone with the stable (value, index) ranks.
  // The member arrays are padded to a multiple of 4 with +inf and read as float4.
  float med[3] = {0.f, 0.f, 0.f};
  unsigned any_ztie = 0, any_xtie = 0, any_ytie = 0, found = 0;
  const int n4 = (n + 3) & ~3;
  if (lane < n4 - n) {
    const float inf = __int_as_float(0x7f800000);
    s.x[n + lane] = inf; s.y[n + lane] = inf; s.z[n + lane] = inf;
  }
  __syncwarp();
  const float4 *X4 = reinterpret_cast<const float4 *>(s.x), *Y4 = reinterpret_cast<const float4 *>(s.y),
               *Z4 = reinterpret_cast<const float4 *>(s.z);
  for (int m = lane; m < ((n + 31) & ~31); m += 32) {
    int lx = 0, ly = 0, lz = 0, ez = 0, ex = 0, ey = 0;
    float xm = 0.f, ym = 0.f, zm = 0.f;
    if (m < n) {
      xm = s.x[m]; ym = s.y[m]; zm = s.z[m];
      float fx = 0.f, fy = 0.f, fz = 0.f, fe = 0.f;
      for (int j4 = 0; j4 < (n4 >> 2); ++j4) {
        const float4 xv = X4[j4], yv = Y4[j4], zv = Z4[j4];
        if (REPLAY) {
          ex += (xv.x == xm) + (xv.y == xm) + (xv.z == xm) + (xv.w == xm);
          ey += (yv.x == ym) + (yv.y == ym) + (yv.z == ym) + (yv.w == ym);
        }
        // counts accumulate as floats (exact: n <= kVtxCap): one FSET (1.0f / 0.0f) on the ALU
        // pipe + one FADD on the FMA pipe per comparison, instead of compare + add + predicated
        // move (two ALU + one FMA) for an integer count
        fx += (flt(xv.x, xm) + flt(xv.y, xm)) + (flt(xv.z, xm) + flt(xv.w, xm));
        fy += (flt(yv.x, ym) + flt(yv.y, ym)) + (flt(yv.z, ym) + flt(yv.w, ym));
        fz += (flt(zv.x, zm) + flt(zv.y, zm)) + (flt(zv.z, zm) + flt(zv.w, zm));
        fe += (feq(zv.x, zm) + feq(zv.y, zm)) + (feq(zv.z, zm) + feq(zv.w, zm));
      }
      lx = (int)fx; ly = (int)fy; lz = (int)fz; ez = (int)fe;
    }
    any_ztie |= __ballot_sync(kFull, m < n && ez > 1);
    if (REPLAY) {
      any_xtie |= __ballot_sync(kFull, m < n && ex > 1);
      any_ytie |= __ballot_sync(kFull, m < n && ey > 1);
      if (m < n) { perm[lx] = (int16_t)m; perm[kVtxCap + ly] = (int16_t)m; }  // x / y order if tie-free
    }
    const unsigned bx = __ballot_sync(kFull, m < n && lx == middle);
    const unsigned by = __ballot_sync(kFull, m < n && ly == middle);
    const unsigned bz = __ballot_sync(kFull, m < n && lz == middle);
    if (bx) { med[0] = __shfl_sync(kFull, xm, __ffs(bx) - 1); found |= 1u; }
    if (by) { med[1] = __shfl_sync(kFull, ym, __ffs(by) - 1); found |= 2u; }
    if (bz) { med[2] = __shfl_sync(kFull, zm, __ffs(bz) - 1); found |= 4u; }
    __syncwarp();
    // final order when z has no ties (the common case); a member with a tied z skips the
    // store (two of them would hit one slot) -- the tie paths below rewrite the whole order
    if (m < n && ez == 1) s.order[lz] = (int16_t)m;
  }
  if (found != 7u) {  // ties around a median: stable ranks (value, then index)
    for (int m = lane; m < ((n + 31) & ~31); m += 32) {
      int rx = 0, ry = 0, rz = 0;
      float xm = 0.f, ym = 0.f, zm = 0.f;
      if (m < n) {
        xm = s.x[m]; ym = s.y[m]; zm = s.z[m];
        for (int j = 0; j < n; ++j) {
          const float xj = s.x[j], yj = s.y[j], zj = s.z[j];
          rx += (xj < xm) || (xj == xm && j < m);
          ry += (yj < ym) || (yj == ym && j < m);
          rz += (zj < zm) || (zj == zm && j < m);
        }
      }
      const unsigned bx = __ballot_sync(kFull, m < n && rx == middle);
      const unsigned by = __ballot_sync(kFull, m < n && ry == middle);
      const unsigned bz = __ballot_sync(kFull, m < n && rz == middle);
      if (bx) med[0] = __shfl_sync(kFull, xm, __ffs(bx) - 1);
      if (by) med[1] = __shfl_sync(kFull, ym, __ffs(by) - 1);
      if (bz) med[2] = __shfl_sync(kFull, zm, __ffs(bz) - 1);
    }
  }
  __syncwarp();
  if (any_ztie) {
    // Exact z ties: the order of the tied points is whatever the reference's three std::sort
    // calls (by x, by y, by z; trellis.cpp:71-82) leave behind.  Up to 16 points libstdc++
    // sorts by insertion (stable), so the result is the lexicographic (z, y, x, column) order.
    // Beyond that its introsort is not stable and one lane replays it (dev_stdsort.h).  That
    // code lives in a second instance of the kernel that only sees the (rare) tied items: with
    // it inside the main instance every item ran 25 % slower.
    if (n > 16) {
      if (!REPLAY) return true;
      __syncwarp();
      const int first = any_ytie ? (any_xtie ? 0 : 1) : 2;
      if (lane == 0) replay_sorts_packed(s.order, pack, n, first, first == 1 ? perm : perm + kVtxCap, s.x, s.y, s.z);
    } else
    for (int m = lane; m < n; m += 32) {
      const float xm = s.x[m], ym = s.y[m], zm = s.z[m];
      const int cm = s.col[m];
      int rk = 0;
      for (int j = 0; j < n; ++j) rk += key_less(s.z[j], s.y[j], s.x[j], s.col[j], zm, ym, xm, cm);
      s.order[rk] = (int16_t)m;
    }
    __syncwarp();
  }
  // keep points within max_dist_to_centroid of the median, in z order (trellis.cpp:89-93)
  const float cut = dp->centroid_sq_cut;  // dist < max_dist_to_centroid  <=>  squared dist < cut
  int kept = 0;
  for (int sidx = lane; sidx < ((n + 31) & ~31); sidx += 32) {
    bool keep = false;
    int m = 0;
    if (sidx < n) {
      m = s.order[sidx];
      keep = sqnorm3f(s.x[m] - med[0], s.y[m] - med[1], s.z[m] - med[2]) < cut;
    }
    const unsigned b = __ballot_sync(kFull, keep);
    __syncwarp();  // every lane has read its order[] entry before any lane overwrites one
    if (keep) s.order[kept + __popc(b & ((1u << lane) - 1u))] = (int16_t)m;  // in-place: kept+pos <= sidx
    kept += __popc(b);
    __syncwarp();
  }
  sloam_vertex v;
  v.cx = med[0]; v.cy = med[1]; v.cz = med[2];
  v.radius = 0.f; v.n_points = 0; v.point_begin = 0; v.row = row; v.is_valid = 0;
  if (kept > 1) {  // trellis.cpp:95-100
    const int a = s.order[0], b = s.order[kept - 1];
    v.radius = dist3f(s.x[a], s.y[a], s.z[a], s.x[b], s.y[b], s.z[b]);
    int base = 0;
    if (lane == 0) base = atomicAdd(pool_count, kept);
    base = __shfl_sync(kFull, base, 0);
    for (int q = lane; q < kept; q += 32) {
      const int m = s.order[q];
      sloam_point p; p.x = s.x[m]; p.y = s.y[m]; p.z = s.z[m]; p.intensity = s.w[m];
      st_point(pool + base + q, p);
    }
    v.n_points = kept; v.point_begin = base; v.is_valid = 1;
  }
  if (lane == 0) *out = v;
  return false;
}

#ifndef SLOAM_VTX_MIN
#define SLOAM_VTX_MIN 5
#endif
template <bool REPLAY>
__global__ void __launch_bounds__(kVtxWarps * 32, SLOAM_VTX_MIN)
vertex_kernel(const DevParams *__restrict__ dp, const sloam_point *__restrict__ tree,
              const uint32_t *__restrict__ bits, const int32_t *__restrict__ parent, const int32_t *__restrict__ big_roots,
              const int32_t *__restrict__ bbox, const int32_t *__restrict__ vwork,
              const int32_t *__restrict__ n_vwork, sloam_vertex *__restrict__ slot_vertices,
              sloam_point *__restrict__ pool, int32_t *__restrict__ pool_count,
              int32_t *__restrict__ overflow, int32_t *__restrict__ n_overflow,
              int32_t *__restrict__ tied, int32_t *__restrict__ n_tied) {
  __shared__ __align__(16) VtxSmem sm[kVtxWarps];
  __shared__ unsigned long long s_pack[REPLAY ? kVtxWarps * kVtxCap : 1];  // replay scratch
  __shared__ int16_t s_perm[REPLAY ? kVtxWarps * 2 * kVtxCap : 1];        // members in x and in y order
  const int N = dp->N, W = dp->p.img_w, H = dp->p.img_h, T = dp->p.max_trees;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  VtxSmem &s = sm[warp];
  unsigned long long *pack = REPLAY ? s_pack + warp * kVtxCap : s_pack;
  int16_t *perm = REPLAY ? s_perm + warp * 2 * kVtxCap : s_perm;
  const int total = *n_vwork;
  for (int item = blockIdx.x * kVtxWarps + warp; item < total; item += gridDim.x * kVtxWarps) {
    const int w = vwork[item];
    int k, slot, row;
    vw_unpack(dp, w, k, slot, row);
    const int root = big_roots[(size_t)k * T + slot];
    const int32_t *bb = bbox + ((size_t)k * T + slot) * 4;
    const int c0 = bb[0], c1 = bb[1];
    const size_t rbase = (size_t)k * N + (size_t)row * W;
    const uint32_t *bk = bits + (size_t)k * ((N + 31) >> 5);
    sloam_vertex *out = slot_vertices + ((size_t)k * T + slot) * H + row;
    // members of this cluster in this row, in column order (trellis.cpp:113-118)
    int n = 0;
    for (int c = c0 + lane; c < ((c1 - c0 + 32) & ~31) + c0; c += 32) {
      // both loads are unconditional so that they are in flight together
      const int cc = c <= c1 ? c : c1;
      const int pc = parent[rbase + cc];  // stale where the bit is clear: never trusted alone
      const uint32_t bw = bk[(row * W + cc) >> 5];
      const bool mem = c <= c1 && ((bw >> ((row * W + cc) & 31)) & 1u) && pc == root;
      const unsigned b = __ballot_sync(kFull, mem);
      if (mem) {
        const int pos = n + __popc(b & ((1u << lane) - 1u));
        if (pos < kVtxCap) {
          const sloam_point p = ld_point(tree + rbase + c);
          s.x[pos] = p.x; s.y[pos] = p.y; s.z[pos] = p.z; s.w[pos] = p.intensity;
          s.col[pos] = (int16_t)c;
        }
      }
      n += __popc(b);
    }
    __syncwarp();
    if (n > kVtxCap) {  // rare: very wide cluster, handled by the wide kernel
      if (lane == 0) overflow[atomicAdd(n_overflow, 1)] = w;
      continue;
    }
    if (n > dp->p.min_vertex_points) {  // trellis.cpp:119
      if (build_vertex<REPLAY>(dp, s, pack, perm, n, row, out, pool + (size_t)k * N, pool_count + k) && lane == 0)
        tied[atomicAdd(n_tied, 1)] = w;
    } else if (lane == 0) {
      sloam_vertex v;
      v.cx = v.cy = v.cz = 0.f; v.radius = 0.f; v.n_points = 0; v.point_begin = 0; v.row = row; v.is_valid = 0;
      *out = v;
    }
    __syncwarp();
  }
}

// wide path: one CTA per overflow item, all members in global scratch order --
// same algorithm with a block-wide rank sort; members live in dynamic smem.
__global__ void vertex_wide_kernel(const DevParams *__restrict__ dp, const sloam_point *__restrict__ tree,
                                   const uint32_t *__restrict__ bits, const int32_t *__restrict__ parent, const int32_t *__restrict__ big_roots,
                                   const int32_t *__restrict__ bbox, const int32_t *__restrict__ overflow,
                                   const int32_t *__restrict__ n_overflow,
                                   sloam_vertex *__restrict__ slot_vertices, sloam_point *__restrict__ pool,
                                   int32_t *__restrict__ pool_count) {
  extern __shared__ float dsm[];
  const int N = dp->N, W = dp->p.img_w, H = dp->p.img_h, T = dp->p.max_trees;
  float *sx = dsm, *sy = dsm + W, *sz = dsm + 2 * W, *sw = dsm + 3 * W;
  int *scol = (int *)(dsm + 4 * W);
  int *sorder = scol + W;
  int *skeep = sorder + W;
  __shared__ int s_n, s_kept, s_base;
  __shared__ float s_med[3];
  const int total = *n_overflow;
  for (int item = blockIdx.x; item < total; item += gridDim.x) {
    const int w = overflow[item];
    int k, slot, row;
    vw_unpack(dp, w, k, slot, row);
    const int root = big_roots[(size_t)k * T + slot];
    const size_t rbase = (size_t)k * N + (size_t)row * W;
    sloam_vertex *out = slot_vertices + ((size_t)k * T + slot) * H + row;
    __syncthreads();
    if (threadIdx.x == 0) {  // serial member collection keeps column order (rare path)
      int n = 0;
      const int32_t *bb = bbox + ((size_t)k * T + slot) * 4;
      for (int c = bb[0]; c <= bb[1]; ++c)
        if (tree_bit(bits + (size_t)k * ((N + 31) >> 5), row * W + c) && parent[rbase + c] == root) {
          const sloam_point p = ld_point(tree + rbase + c);
          sx[n] = p.x; sy[n] = p.y; sz[n] = p.z; sw[n] = p.intensity; scol[n] = c; ++n;
        }
      s_n = n;
    }
    __syncthreads();
    const int n = s_n;
    const int middle = (int)(n / 2.0);
    for (int m = threadIdx.x; m < n; m += blockDim.x) {
      int rx = 0, ry = 0, rz = 0, rk = 0;
      const float xm = sx[m], ym = sy[m], zm = sz[m];
      for (int j = 0; j < n; ++j) {
        rx += (sx[j] < xm) || (sx[j] == xm && j < m);
        ry += (sy[j] < ym) || (sy[j] == ym && j < m);
        rz += (sz[j] < zm) || (sz[j] == zm && j < m);
        rk += key_less(sz[j], sy[j], sx[j], scol[j], zm, ym, xm, scol[m]);
      }
      sorder[rk] = m;
      if (rx == middle) s_med[0] = xm;
      if (ry == middle) s_med[1] = ym;
      if (rz == middle) s_med[2] = zm;
    }
    __syncthreads();
    if (threadIdx.x == 0) {  // exact z ties among > 16 points: replay the three std::sort calls
      bool tie = false;
      for (int q = 0; q + 1 < n && !tie; ++q) tie = sz[sorder[q]] == sz[sorder[q + 1]];
      if (tie) {
        replay_three_sorts(sorder, n, sx, sy, sz);
      }
    }
    __syncthreads();
    const float maxd = dp->p.max_dist_to_centroid;
    for (int q = threadIdx.x; q < n; q += blockDim.x) {
      const int m = sorder[q];
      skeep[q] = dist3f(sx[m], sy[m], sz[m], s_med[0], s_med[1], s_med[2]) < maxd;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
      int kept = 0;
      for (int q = 0; q < n; ++q) if (skeep[q]) sorder[kept++] = sorder[q];
      s_kept = kept;
      sloam_vertex v;
      v.cx = s_med[0]; v.cy = s_med[1]; v.cz = s_med[2];
      v.radius = 0.f; v.n_points = 0; v.point_begin = 0; v.row = row; v.is_valid = 0;
      if (kept > 1) {
        const int a = sorder[0], b = sorder[kept - 1];
        v.radius = dist3f(sx[a], sy[a], sz[a], sx[b], sy[b], sz[b]);
        s_base = atomicAdd(pool_count + k, kept);
        v.n_points = kept; v.point_begin = s_base; v.is_valid = 1;
      }
      *out = v;
    }
    __syncthreads();
    if (s_kept > 1)
      for (int q = threadIdx.x; q < s_kept; q += blockDim.x) {
        const int m = sorder[q];
        sloam_point p; p.x = sx[m]; p.y = sy[m]; p.z = sz[m]; p.intensity = sw[m];
        pool[(size_t)k * N + s_base + q] = p;
      }
  }
}

// ---- 6. trees: keep clusters with enough vertices, bottom row first --------
__global__ void tree_compact_kernel(const DevParams *__restrict__ dp, const int32_t *__restrict__ n_big,
                                    const int32_t *__restrict__ big_rank, const int32_t *__restrict__ bbox,
                                    const sloam_vertex *__restrict__ slot_vertices,
                                    sloam_tree *__restrict__ trees, int32_t *__restrict__ n_trees,
                                    sloam_vertex *__restrict__ vertices) {
  extern __shared__ int32_t sm[];
  const int H = dp->p.img_h, T = dp->p.max_trees;
  const int minv = dp->p.min_tree_vertices, maxv = dp->p.max_tree_vertices;
  const int k = blockIdx.x;
  const int nb = n_big[k];
  int32_t *s_nv = sm;        // [T] vertices of the slot (0 when rejected)
  int32_t *s_idx = sm + T;   // [T] tree index
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
  for (int s = warp; s < nb; s += nwarps) {
    const int32_t *bb = bbox + ((size_t)k * T + s) * 4;
    const sloam_vertex *sv = slot_vertices + ((size_t)k * T + s) * H;
    int cnt = 0;
    for (int top = bb[3]; top >= bb[2]; top -= 32) {  // rows from the bottom up (trellis.cpp:111)
      const int r = top - lane;
      const bool v = r >= bb[2] && sv[r].is_valid;
      cnt += __popc(__ballot_sync(kFull, v));
    }
    if (lane == 0) s_nv[s] = cnt > minv ? min(cnt, maxv) : 0;  // trellis.cpp:124-127
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    int t = 0;
    for (int s = 0; s < nb; ++s) { s_idx[s] = t; t += s_nv[s] > 0; }
    n_trees[k] = t;
  }
  __syncthreads();
  for (int s = warp; s < nb; s += nwarps) {
    const int nv = s_nv[s];
    if (nv == 0) continue;
    const int t = s_idx[s];
    const int32_t *bb = bbox + ((size_t)k * T + s) * 4;
    const sloam_vertex *sv = slot_vertices + ((size_t)k * T + s) * H;
    sloam_vertex *dst = vertices + ((size_t)k * T + t) * maxv;
    int cnt = 0, npts = 0;
    for (int top = bb[3]; top >= bb[2]; top -= 32) {
      const int r = top - lane;
      const bool v = r >= bb[2] && sv[r].is_valid;
      const unsigned b = __ballot_sync(kFull, v);
      const int pos = cnt + __popc(b & ((1u << lane) - 1u));
      int np = 0;
      if (v && pos < nv) { const sloam_vertex vv = sv[r]; dst[pos] = vv; np = vv.n_points; }
      npts += warp_sum(np);
      cnt += __popc(b);
    }
    if (lane == 0) {
      sloam_tree tr;
      tr.tree_id = big_rank[(size_t)k * T + s];
      tr.n_vertices = nv;
      tr.vertex_begin = t * maxv;
      tr.n_points = npts;
      trees[(size_t)k * T + t] = tr;
    }
  }
}

// ---- labels for the find_clusters stage entry -------------------------------
__global__ void cc_rank_rows_kernel(const DevParams *__restrict__ dp, const uint32_t *__restrict__ bits,
                                    const int32_t *__restrict__ parent,
                                    const int32_t *__restrict__ row_roots, int32_t *__restrict__ root_rank) {
  const int N = dp->N, W = dp->p.img_w, H = dp->p.img_h;
  const int k = blockIdx.y;
  const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= H) return;
  const int lane = threadIdx.x & 31;
  int pre = 0;
  for (int r = lane; r < row; r += 32) pre += row_roots[(size_t)k * H + r];
  pre = warp_sum(pre);
  const size_t rbase = (size_t)k * N + (size_t)row * W;
  for (int c = lane; c < ((W + 31) & ~31); c += 32) {
    const bool is_root = c < W && tree_bit(bits + (size_t)k * ((N + 31) >> 5), row * W + c) &&
                         parent[rbase + c] == row * W + c;
    const unsigned b = __ballot_sync(kFull, is_root);
    if (is_root) root_rank[rbase + c] = pre + __popc(b & ((1u << lane) - 1u));
    pre += __popc(b);
  }
}
__global__ void cc_labels_kernel(const DevParams *__restrict__ dp, int K, const uint32_t *__restrict__ bits,
                                 const int32_t *__restrict__ parent,
                                 const int32_t *__restrict__ root_rank, uint32_t *__restrict__ labels) {
  const int N = dp->N;
  const long long g = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (g >= (long long)K * N) return;
  const int k = (int)(g / N);
  const int i = (int)(g - (long long)k * N);
  const int r = tree_bit(bits + (size_t)k * ((N + 31) >> 5), i) ? parent[g] : kInvalid;
  labels[g] = r == kInvalid ? 0xFFFFFFFFu : (uint32_t)root_rank[(size_t)k * N + r];
}

static int run_cc(sloam_ctx *c, int K, const sloam_point *tree, bool bits_ready) {
  Workspace &w = c->ws;
  const long long total = (long long)K * c->hp.N;
  const int H = c->hp.p.img_h;
  const bool pre_zeroed = (c->zero_valid & 2u) != 0;  // zeroed with the rest of the counters (pipeline.cu)
  c->zero_valid &= ~2u;
  if (!pre_zeroed) {
    SB_CUDA(c, cudaMemsetAsync(w.row_roots, 0, sizeof(int32_t) * (size_t)K * H, c->stream));
    SB_CUDA(c, cudaMemsetAsync(w.n_roots, 0, sizeof(int32_t) * K, c->stream));
    SB_CUDA(c, cudaMemsetAsync(w.n_big, 0, sizeof(int32_t) * K, c->stream));
    SB_CUDA(c, cudaMemsetAsync(w.kf_flags, 0, sizeof(int32_t) * K, c->stream));
  }
  const dim3 blocks((unsigned)((c->hp.N + 255) / 256), (unsigned)K);
  if (!bits_ready) {  // caller-supplied cloud: derive the bits from the points
    tree_bits_kernel<<<blocks, 256, 0, c->stream>>>(c->dp, tree, w.tree_bits);
    SB_LAUNCH_CHECK(c);
  }
  const unsigned wgrid = (unsigned)std::min<long long>((total / 32 / 32 / 8) + 1, (long long)c->sm_count * 8);
  if (!pre_zeroed) {
    SB_CUDA(c, cudaMemsetAsync(w.n_tree_words, 0, sizeof(int32_t), c->stream));
    SB_CUDA(c, cudaMemsetAsync(w.root_bits, 0, sizeof(uint32_t) * (size_t)K * ((c->hp.N + 31) / 32), c->stream));
  }
  PROF_BEGIN(c, P_CC_WORDS);
  tree_words_kernel<<<wgrid, 256, 0, c->stream>>>(c->dp, K, w.tree_bits, reinterpret_cast<int2 *>(w.tree_words), w.n_tree_words);
  PROF_END(c, P_CC_WORDS);
  SB_LAUNCH_CHECK(c);
  const int2 *wl = reinterpret_cast<const int2 *>(w.tree_words);
  PROF_BEGIN(c, P_CC_INIT);
  cc_init_kernel<<<wgrid, 256, 0, c->stream>>>(c->dp, K, tree, w.tree_bits, wl, w.n_tree_words, w.parent,
                                               reinterpret_cast<uint32_t *>(w.cc_flags), w.csize,
                                                w.ccol_min, w.ccol_max, w.crow_max);
  PROF_END(c, P_CC_INIT);
  SB_LAUNCH_CHECK(c);
  PROF_BEGIN(c, P_CC_MERGE);
  cc_merge_kernel<<<wgrid, 256, 0, c->stream>>>(c->dp, K, w.tree_bits, wl, w.n_tree_words, reinterpret_cast<const uint32_t *>(w.cc_flags), w.parent);
  PROF_END(c, P_CC_MERGE);
  SB_LAUNCH_CHECK(c);
  PROF_BEGIN(c, P_CC_FLATTEN);
  cc_flatten_kernel<<<wgrid, 256, 0, c->stream>>>(c->dp, K, w.tree_bits, wl, w.n_tree_words, w.parent, w.csize, w.ccol_min, w.ccol_max,
                                                   w.crow_max, w.row_roots, w.n_roots, w.big_roots,
                                                   w.n_big, w.kf_flags, w.root_bits);
  PROF_END(c, P_CC_FLATTEN);
  SB_LAUNCH_CHECK(c);
  return SLOAM_OK;
}

int launch_compute_graph(sloam_ctx *c, int K, const sloam_point *tree, sloam_tree *trees,
                         int32_t *n_trees, sloam_vertex *vertices, sloam_point *vertex_points,
                         bool bits_ready) {
  Workspace &w = c->ws;
  const sloam_params &p = c->hp.p;
  const int T = p.max_trees, H = p.img_h, W = p.img_w;
  int rc = run_cc(c, K, tree, bits_ready);
  if (rc != SLOAM_OK) return rc;
  if (c->zero_valid & 4u) {
    c->zero_valid &= ~4u;
  } else {
    SB_CUDA(c, cudaMemsetAsync(w.n_overflow, 0, sizeof(int32_t) * 4, c->stream));
    SB_CUDA(c, cudaMemsetAsync(w.vpool_count, 0, sizeof(int32_t) * K, c->stream));
  }
  PROF_BEGIN(c, P_CC_PLAN);
  cc_plan_kernel<<<K, 256, sizeof(int32_t) * (2 * T + H + 1), c->stream>>>(
      c->dp, w.root_bits, w.parent, w.ccol_min, w.ccol_max, w.crow_max, w.row_roots, w.big_roots, w.n_big,
      w.big_rank, w.bbox, w.vwork, w.n_overflow + 1);
  PROF_END(c, P_CC_PLAN);
  SB_LAUNCH_CHECK(c);
  // persistent grid: exactly the CTAs that are resident at once (a partial second wave would
  // double the time of its work items)
  static int vtx_occ = 0;
  if (vtx_occ == 0) {
    SB_CUDA(c, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&vtx_occ, vertex_kernel<false>, kVtxWarps * 32, 0));
    if (vtx_occ < 1) vtx_occ = 1;
  }
  const int vgrid = c->sm_count * vtx_occ;
  // n_overflow[0] rows wider than the warp path, [1] work items, [2] items with exact z ties
  PROF_BEGIN(c, P_VERTEX);
  vertex_kernel<false><<<vgrid, kVtxWarps * 32, 0, c->stream>>>(
      c->dp, tree, w.tree_bits, w.parent, w.big_roots, w.bbox, w.vwork, w.n_overflow + 1, w.slot_vertices,
      vertex_points, w.vpool_count, w.overflow_list, w.n_overflow, w.tied_list, w.n_overflow + 2);
  PROF_END(c, P_VERTEX);
  SB_LAUNCH_CHECK(c);
  PROF_BEGIN(c, P_VERTEX_REPLAY);
  // the tied items again, with the std::sort replay (usually an empty list: the CTAs exit)
  vertex_kernel<true><<<c->sm_count, kVtxWarps * 32, 0, c->stream>>>(
      c->dp, tree, w.tree_bits, w.parent, w.big_roots, w.bbox, w.tied_list, w.n_overflow + 2, w.slot_vertices,
      vertex_points, w.vpool_count, w.overflow_list, w.n_overflow, nullptr, nullptr);
  PROF_END(c, P_VERTEX_REPLAY);
  SB_LAUNCH_CHECK(c);
  const size_t wide_smem = sizeof(float) * 4 * W + sizeof(int) * 3 * W;
  // opt-in shared memory: the attribute belongs to (function, device) and must only grow -- a
  // context with a wider image than an earlier one needs more, a narrower one must not lower it
  static size_t wide_set[64] = {};
  if (wide_smem > 48 * 1024 && wide_smem > wide_set[c->device & 63]) {
    SB_CUDA(c, cudaFuncSetAttribute(vertex_wide_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)wide_smem));
    wide_set[c->device & 63] = wide_smem;
  }
  vertex_wide_kernel<<<c->sm_count, 256, wide_smem, c->stream>>>(c->dp, tree, w.tree_bits, w.parent, w.big_roots, w.bbox,
                                                                 w.overflow_list, w.n_overflow,
                                                                 w.slot_vertices, vertex_points,
                                                                 w.vpool_count);
  SB_LAUNCH_CHECK(c);
  PROF_BEGIN(c, P_TREE_COMPACT);
  tree_compact_kernel<<<K, 256, sizeof(int32_t) * 2 * T, c->stream>>>(c->dp, w.n_big, w.big_rank, w.bbox,
                                                                       w.slot_vertices, trees, n_trees,
                                                                       vertices);
  PROF_END(c, P_TREE_COMPACT);
  SB_LAUNCH_CHECK(c);
  return SLOAM_OK;
}

// materialise the dense tree cloud of the last fused run (see sloam_ctx::tree_sparse)
int launch_tree_fill(sloam_ctx *c, int K) {
  const dim3 blocks((unsigned)((c->hp.N + 255) / 256), (unsigned)K);
  tree_fill_kernel<<<blocks, 256, 0, c->stream>>>(c->dp, c->ws.tree_bits, c->ws.tree);
  SB_LAUNCH_CHECK(c);
  return SLOAM_OK;
}

}  // namespace sb

using namespace sb;

extern "C" {

int sloam_b200_find_clusters_dev(sloam_ctx *c, int K, const sloam_point *tree, uint32_t *labels,
                                 int32_t *n_clusters) {
  if (!c || K <= 0 || K > c->max_k || !tree || !labels) return set_err(c, SLOAM_E_INVALID, "find_clusters: bad arguments");
  int rc = run_cc(c, K, tree, false);
  if (rc != SLOAM_OK) return rc;
  const int H = c->hp.p.img_h;
  dim3 grid((unsigned)((H + 7) / 8), (unsigned)K);
  cc_rank_rows_kernel<<<grid, 256, 0, c->stream>>>(c->dp, c->ws.tree_bits, c->ws.parent, c->ws.row_roots, c->ws.root_rank);
  SB_LAUNCH_CHECK(c);
  const long long total = (long long)K * c->hp.N;
  cc_labels_kernel<<<(unsigned)((total + 255) / 256), 256, 0, c->stream>>>(c->dp, K, c->ws.tree_bits, c->ws.parent,
                                                                          c->ws.root_rank, labels);
  SB_LAUNCH_CHECK(c);
  if (n_clusters)
    SB_CUDA(c, cudaMemcpyAsync(n_clusters, c->ws.n_roots, sizeof(int32_t) * K, cudaMemcpyDeviceToDevice, c->stream));
  return SLOAM_OK;
}

int sloam_b200_compute_graph_dev(sloam_ctx *c, int K, const sloam_point *tree, sloam_tree *trees,
                                 int32_t *n_trees, sloam_vertex *vertices, sloam_point *vertex_points) {
  if (!c || K <= 0 || K > c->max_k || !tree || !trees || !n_trees || !vertices || !vertex_points)
    return set_err(c, SLOAM_E_INVALID, "compute_graph: bad arguments");
  return launch_compute_graph(c, K, tree, trees, n_trees, vertices, vertex_points, false);
}

}  // extern "C"
