// k6_pose.cu -- stages a12, a14..a19: feature matching, pose optimisation,
// model projection and the RunSloam orchestration, batched over keyframes.
//
// Replaces sloam::RunSloam (sloam/src/core/sloam.cpp:453-532) after
// computeModels: matchFeatures + addFeatureMatches (:257-296), the gating
// (:499-500), OptimizePose / TwoStepOptimizePose (:33-255), projectModels
// (:438-451), matchModels (:298-328).  The per-keyframe state of the reference
// (firstScan_, prevGPlanes_) and the submap are explicit inputs.
//
// The LM kernel is the fused residual / Jacobian / Huber / J^T J accumulation
// the north star asks for: every thread evaluates residual rows and adds
// r~, J~ into private 6x6 (packed 21) + 6 + 1 accumulators, a shuffle + shared
// memory block reduction produces the normal equations, and the whole <= 50
// iteration trust-region loop (dev_lm.h) runs inside the kernel, one CTA per
// (keyframe, problem).  Latency-bound: ~80 B per residual, re-read from L2 per
// iteration.
#include "common.cuh"
#include "dev_lm.h"

namespace sb {

#ifndef SLOAM_LM_WARPS
#define SLOAM_LM_WARPS 2  // measured per 1024 OS1-64 problems: 1 warp 102 us, 2 -> 91, 4 -> 118, 8 -> 178
#endif
constexpr int kLmWarps = SLOAM_LM_WARPS;  // warps that share the residual rows of ONE problem (= one CTA)
constexpr int kLmThreads = 32 * kLmWarps;

// ---------------------------------------------------------------- matching --
__global__ void __launch_bounds__(128)
build_matches_kernel(const DevParams *__restrict__ dp, const uint8_t *__restrict__ first_scan,
                     const sloam_pose *__restrict__ pose_est, const int32_t *__restrict__ n_map,
                     int map_shared, const sloam_cylinder *__restrict__ map, int map_stride,
                     const sloam_plane *__restrict__ prev_planes, const int32_t *__restrict__ n_prev,
                     int prev_stride, const int32_t *__restrict__ n_lm, const int32_t *__restrict__ lm_src,
                     const int32_t *__restrict__ assoc_idx, const double *__restrict__ assoc_dist,
                     const sloam_point *__restrict__ tree_features, const sloam_plane *__restrict__ planes_acc,
                     const int32_t *__restrict__ planes_acc_cell, const int32_t *__restrict__ n_planes_acc,
                     const sloam_point *__restrict__ cell_features, double *__restrict__ res_tree_feat,
                     sloam_cylinder *__restrict__ res_tree_obj, double *__restrict__ res_plane_feat,
                     sloam_plane *__restrict__ res_plane_obj, int32_t *__restrict__ n_tree_res,
                     int32_t *__restrict__ n_plane_res, uint8_t *__restrict__ optim_flags,
                     uint8_t *__restrict__ kf_mode, sloam_kf_result *__restrict__ results,
                     const int32_t *__restrict__ ground_count, const int32_t *__restrict__ n_trees,
                     const int32_t *__restrict__ kf_flags) {
  __shared__ int s_warp[4];
  __shared__ int s_best[kMaxCells];
  __shared__ int16_t s_pslot[kMaxCells];
  extern __shared__ int16_t s_slot[];  // [max_trees] slot of each landmark's matches, -1 = unmatched
  const sloam_params &P = dp->p;
  const int T = P.max_trees, B = dp->B, Ft = P.featuresPerTree, Fg = P.numGroundFeatures;
  const int k = blockIdx.x;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int nl = n_lm[k], npl = n_planes_acc[k];
  // counts beyond the capacities the buffers were sized for are clamped (the host entries reject them)
  const int nm = min(map_shared ? n_map[0] : n_map[k], map_stride);
  const int npv = min(n_prev[k], prev_stride);
  const sloam_cylinder *mk = map_shared ? map : map + (size_t)k * map_stride;
  int mode = 0, status = SLOAM_KF_OK;
  if (first_scan[k]) mode = 1;
  else if (nm == 0) { mode = 2; status = SLOAM_KF_EMPTY_MAP; }          // sloam.cpp:476-480
  else if (npl == 0 || nl == 0) { mode = 2; status = SLOAM_KF_NO_MODELS; }  // :482-486
  int n_tres = 0, n_pres = 0;
  if (mode == 0) {
    // ---- matchFeatures<Cylinder>: landmarks whose nearest map cylinder is within treeMatchThresh
    double *tf = res_tree_feat + (size_t)k * T * Ft * 3;
    sloam_cylinder *to = res_tree_obj + (size_t)k * T * Ft;
    int base = 0;
    for (int i0 = 0; i0 < nl; i0 += 128) {
      const int i = i0 + threadIdx.x;
      bool m = false;
      if (i < nl) {
        const int idx = assoc_idx[(size_t)k * T + i];
        m = idx >= 0 && assoc_dist[(size_t)k * T + i] < P.treeMatchThresh;
      }
      const unsigned b = __ballot_sync(kFull, m);
      if (lane == 0) s_warp[warp] = __popc(b);
      __syncthreads();
      int off = base, tot = 0;
      for (int w = 0; w < 4; ++w) { if (w < warp) off += s_warp[w]; tot += s_warp[w]; }
      if (i < nl) s_slot[i] = m ? off + __popc(b & ((1u << lane) - 1u)) : -1;
      base += tot;
      __syncthreads();
    }
    n_tres = base * Ft;
    // addFeatureMatches: one match per feature (:288-296); all threads share the copies
    for (int e = threadIdx.x; e < nl * Ft; e += 128) {
      const int i = e / Ft, q = e - i * Ft, slot = s_slot[i];
      if (slot < 0) continue;
      const sloam_point f = tree_features[((size_t)k * T + lm_src[(size_t)k * T + i]) * Ft + q];
      const size_t r = (size_t)slot * Ft + q;
      tf[3 * r] = (double)f.x; tf[3 * r + 1] = (double)f.y; tf[3 * r + 2] = (double)f.z;
      to[r] = mk[assoc_idx[(size_t)k * T + i]];
    }
    // ---- matchFeatures<Plane>: centroid distance to the previous planes, threshold 1.0 (:490)
    for (int g = threadIdx.x; g < npl; g += 128) {
      double c2[3];
      pose_apply(pose_est[k], planes_acc[(size_t)k * B + g].centroid, c2);
      double bd = P.plane_match_thresh + 100.0;
      int bi = -1;
      for (int j = 0; j < npv; ++j) {
        const double *pc = prev_planes[(size_t)k * prev_stride + j].centroid;
        const double dx = pc[0] - c2[0], dy = pc[1] - c2[1], dz = pc[2] - c2[2];
        const double d = sqrt(dx * dx + (dy * dy + dz * dz));
        if (d < bd) { bd = d; bi = j; }
      }
      s_best[g] = (bi >= 0 && bd < P.plane_match_thresh) ? bi : -1;
    }
    __syncthreads();
    if (warp == 0) {  // slots of the matched planes, in plane order (npl <= kMaxCells)
      int carry = 0;
      for (int g0 = 0; g0 < npl; g0 += 32) {
        const int g = g0 + lane;
        const bool mm = g < npl && s_best[g] >= 0;
        const unsigned b = __ballot_sync(kFull, mm);
        if (g < npl) s_pslot[g] = mm ? carry + __popc(b & ((1u << lane) - 1u)) : -1;
        carry += __popc(b);
      }
      if (lane == 0) s_warp[0] = carry * Fg;
    }
    __syncthreads();
    n_pres = s_warp[0];
    {
      double *pf = res_plane_feat + (size_t)k * B * Fg * 3;
      sloam_plane *po = res_plane_obj + (size_t)k * B * Fg;
      for (int e = threadIdx.x; e < npl * Fg; e += 128) {
        const int g = e / Fg, q = e - g * Fg, slot = s_pslot[g];
        if (slot < 0) continue;
        const sloam_point f = cell_features[((size_t)k * B + planes_acc_cell[(size_t)k * B + g]) * Fg + q];
        const size_t r = (size_t)slot * Fg + q;
        pf[3 * r] = (double)f.x; pf[3 * r + 1] = (double)f.y; pf[3 * r + 2] = (double)f.z;
        po[r] = prev_planes[(size_t)k * prev_stride + s_best[g]];
      }
    }
  }
  if (threadIdx.x == 0) {
    n_tree_res[k] = n_tres;
    n_plane_res[k] = n_pres;
    // B-1: minPlanes_ from the parameters actually set
    const double minPlanes = (double)(P.groundRadiiBins * P.groundThetaBins) * 0.1;
    const bool treeCheck = (double)nl > P.minTreeModels && (double)n_tres > 5.0 * (double)Ft;       // :499
    const bool groundCheck = (double)n_pres > P.minGroundModels && (double)npl > minPlanes;          // :500
    optim_flags[2 * k] = (mode == 0 && treeCheck) ? 1 : 0;
    optim_flags[2 * k + 1] = (mode == 0 && groundCheck) ? 1 : 0;
    kf_mode[k] = (uint8_t)mode;
    sloam_kf_result r;
    r.status = status | ((kf_flags && (kf_flags[k] & 1)) ? SLOAM_KF_FLAG_TREE_CAPACITY : 0);
    r.success = (mode == 2) ? 0 : 1;
    r.n_ground = ground_count ? ground_count[k] : 0;
    r.n_planes = npl; r.n_trees = n_trees ? n_trees[k] : 0; r.n_landmarks = nl;
    r.n_tree_matches = n_tres; r.n_plane_matches = n_pres;
    r.lm_iterations[0] = r.lm_iterations[1] = 0;
    r.lm_termination[0] = r.lm_termination[1] = -1;
    for (int i = 0; i < 3; ++i) { r.T_Map_Curr.t[i] = 0.0; r.T_Delta.t[i] = 0.0; }
    for (int i = 0; i < 3; ++i) { r.T_Map_Curr.q[i] = 0.0; r.T_Delta.q[i] = 0.0; }
    r.T_Map_Curr.q[3] = 1.0; r.T_Delta.q[3] = 1.0;
    results[k] = r;
  }
}

// ---------------------------------------------------------------- LM solve --
// One CTA of kLmWarps warps per (keyframe, problem).  The threads split the residual rows of an
// evaluation (row r on thread r % kLmThreads), reduce the J^T J / J^T r / cost partial sums
// with shuffles inside each warp and through a double-buffered shared-memory table across the
// warps (ONE block barrier per evaluation; every thread adds the warps' partial sums in the
// same order), and then all run the scalar trust-region state machine of dev_lm.h on
// identical values: a warp instruction costs the same issue slots with one lane or 32, so
// the redundancy is free, and no master/worker hand-over sits on the critical path.  The
// scalar state (LMWork) is in shared memory, one copy per WARP (warps run ahead of each other
// between barriers): every lane reads it by broadcast and writes the same value to the same
// address.
// The kernel is latency-bound on the dependent FP64 chain of one iteration.  A tree problem
// has ~260 rows, i.e. 8 sequential rows per lane with one warp; two warps halve that, but
// every further warp repeats the scalar part (the 3x3 / 6x6 solve, step and ratio tests on
// shared-memory state), which is most of an iteration: four and eight warps are slower.
struct WarpEval {
  int mode;
  const double *tree_feat; const sloam_cylinder *tree_obj; int n_tree;
  const double *plane_feat; const sloam_plane *plane_obj; int n_plane;
  double huber_a;
  double (*part)[kLmWarps][28];  // [2][warp][21 + 6 + 1] partial sums, shared memory
  int n_eval = 0;

  // N = tangent dimension (6 joint, 3 for the two-step problems): a compile-time N keeps
  // the accumulators in registers
  template <int N>
  __device__ __forceinline__ void eval(const double *xs, double *cost, double *A, double *g) {
    constexpr int NP = N * (N + 1) / 2;
    const int lane = threadIdx.x & 31;
    double x[7];
#pragma unroll
    for (int i = 0; i < 7; ++i) x[i] = xs[i];
    AaPre pre;
    if (N == 3) aa_prepare(x, pre);  // pose-only terms: once per evaluation, not per row
    double acc[NP], accg[N], accc = 0.0;
#pragma unroll
    for (int i = 0; i < NP; ++i) acc[i] = 0.0;
#pragma unroll
    for (int i = 0; i < N; ++i) accg[i] = 0.0;
    const int total = n_tree + n_plane;
    for (int r = threadIdx.x; r < total; r += kLmThreads) {
      double J[6];
      double res;
      if (r < n_tree) res = residual_row(mode, x, pre, tree_feat + 3 * (size_t)r, tree_obj + r, nullptr, J);
      else { const int q = r - n_tree; res = residual_row(mode, x, pre, plane_feat + 3 * (size_t)q, nullptr, plane_obj + q, J); }
      double sc;
      accc += huber(res, huber_a, &sc);
      const double rr = res * sc;
      int p = 0;
#pragma unroll
      for (int i = 0; i < N; ++i) {
        const double ji = J[i] * sc;
#pragma unroll
        for (int j = i; j < N; ++j) acc[p++] += ji * (J[j] * sc);
        accg[i] += ji * rr;
      }
    }
#pragma unroll
    for (int i = 0; i < NP; ++i) acc[i] = warp_sum_d(acc[i]);
#pragma unroll
    for (int i = 0; i < N; ++i) accg[i] = warp_sum_d(accg[i]);
    accc = warp_sum_d(accc);
    if (kLmWarps == 1) {
#pragma unroll
      for (int i = 0; i < NP; ++i) A[i] = acc[i];
#pragma unroll
      for (int i = 0; i < N; ++i) g[i] = accg[i];
      *cost = accc;
      return;
    }
    // across the warps: table of evaluation parity (the barrier of the next evaluation fences
    // its reuse), summed by every thread in warp order
    double (*tab)[28] = part[n_eval & 1];
    ++n_eval;
    const int warp = threadIdx.x >> 5;
    if (lane == 0) {
#pragma unroll
      for (int i = 0; i < NP; ++i) tab[warp][i] = acc[i];
#pragma unroll
      for (int i = 0; i < N; ++i) tab[warp][NP + i] = accg[i];
      tab[warp][NP + N] = accc;
    }
    __syncthreads();
#pragma unroll
    for (int i = 0; i < NP; ++i) { double v = tab[0][i]; for (int w = 1; w < kLmWarps; ++w) v += tab[w][i]; A[i] = v; }
#pragma unroll
    for (int i = 0; i < N; ++i) { double v = tab[0][NP + i]; for (int w = 1; w < kLmWarps; ++w) v += tab[w][NP + i]; g[i] = v; }
    { double v = tab[0][NP + N]; for (int w = 1; w < kLmWarps; ++w) v += tab[w][NP + N]; *cost = v; }
  }

  __device__ void operator()(const double *xs, double *cost, double *A, double *g) {
    __syncwarp();
    if (mode == LM_JOINT) eval<6>(xs, cost, A, g);
    else eval<3>(xs, cost, A, g);
    __syncwarp();
  }
};

// problems: joint -> one per keyframe; two-step -> 2 per keyframe, problem 0 = XYYaw (trees),
// 1 = ZRollPitch (planes).  CTA b of the grid solves problem b.
#ifndef SLOAM_LM_MIN_CTAS
#define SLOAM_LM_MIN_CTAS (16 / SLOAM_LM_WARPS)  // 128 registers per thread
#endif
__global__ void __launch_bounds__(kLmThreads, SLOAM_LM_MIN_CTAS)
lm_kernel(const DevParams *__restrict__ dp, int two_step, int K, const sloam_pose *__restrict__ pose_est,
          const double *__restrict__ tree_feat, const sloam_cylinder *__restrict__ tree_obj,
          const int32_t *__restrict__ n_tree_res, int tf_stride, const double *__restrict__ plane_feat,
          const sloam_plane *__restrict__ plane_obj, const int32_t *__restrict__ n_plane_res, int pf_stride,
          const uint8_t *__restrict__ optim_flags, double *__restrict__ lm_x, int32_t *__restrict__ lm_info) {
  __shared__ LMWork s_work[kLmWarps];
  __shared__ double s_part[2][kLmWarps][28];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int pid = blockIdx.x;
  const int per_kf = two_step ? 2 : 1;
  if (pid >= K * per_kf) return;  // CTA-uniform
  const int k = pid / per_kf, prob = pid % per_kf;
  const bool optimTrees = optim_flags[2 * k] != 0, optimGround = optim_flags[2 * k + 1] != 0;
  double *xo = lm_x + ((size_t)k * 2 + prob) * 8;
  int32_t *info = lm_info + ((size_t)k * 2 + prob) * 2;
  const sloam_pose T0 = pose_est[k];
  double x[7];
  int mode;
  bool run;
  if (!two_step) {
    mode = LM_JOINT;
    run = optimTrees && optimGround;  // sloam.cpp:505
    x[0] = T0.q[0]; x[1] = T0.q[1]; x[2] = T0.q[2]; x[3] = T0.q[3];
    x[4] = T0.t[0]; x[5] = T0.t[1]; x[6] = T0.t[2];
  } else {
    mode = prob == 0 ? LM_XYYAW : LM_ZROLLPITCH;
    run = prob == 0 ? optimTrees : optimGround;
    const double qw[4] = {T0.q[3], T0.q[0], T0.q[1], T0.q[2]};
    double aa[3];
    quat_to_angle_axis(qw, aa);  // sloam.cpp:58-63
    x[0] = T0.t[0]; x[1] = T0.t[1]; x[2] = T0.t[2]; x[3] = aa[0]; x[4] = aa[1]; x[5] = aa[2]; x[6] = 0.0;
  }
  LMOut o;
  o.iterations = 0; o.termination = -1; o.initial_cost = 0; o.final_cost = 0;
  if (run) {
    WarpEval ev;
    ev.mode = mode;
    const bool use_trees = mode != LM_ZROLLPITCH, use_planes = mode != LM_XYYAW;
    ev.tree_feat = tree_feat + (size_t)k * tf_stride * 3;
    ev.tree_obj = tree_obj + (size_t)k * tf_stride;
    ev.n_tree = use_trees ? n_tree_res[k] : 0;
    ev.plane_feat = plane_feat + (size_t)k * pf_stride * 3;
    ev.plane_obj = plane_obj + (size_t)k * pf_stride;
    ev.n_plane = use_planes ? n_plane_res[k] : 0;
    ev.huber_a = dp->p.huber_delta;
    ev.part = s_part;
    o = lm_minimize(ev, s_work[warp], mode, ev.n_tree + ev.n_plane, dp->p.lm_max_iterations, x);
  }
  if (threadIdx.x == 0) {
    for (int i = 0; i < 7; ++i) xo[i] = x[i];
    xo[7] = o.final_cost;
    info[0] = o.iterations;
    info[1] = run ? o.termination : -1;
  }
  if (!two_step && threadIdx.x == 1) {  // the unused second problem slot of a joint solve
    int32_t *info1 = lm_info + ((size_t)k * 2 + 1) * 2;
    info1[0] = 0; info1[1] = -1;
  }
}

// ------------------------------------------------- compose, project, output --
__global__ void __launch_bounds__(128)
finish_kernel(const DevParams *__restrict__ dp, int two_step, const sloam_pose *__restrict__ pose_est,
              const uint8_t *__restrict__ kf_mode, const uint8_t *__restrict__ optim_flags,
              const double *__restrict__ lm_x, const int32_t *__restrict__ lm_info,
              const int32_t *__restrict__ n_lm, const sloam_cylinder *__restrict__ lm_cyl,
              const int32_t *__restrict__ lm_src, const sloam_tree_model *__restrict__ tree_models,
              const sloam_plane *__restrict__ planes_acc, const int32_t *__restrict__ n_planes_acc,
              const sloam_plane *__restrict__ prev_planes, const int32_t *__restrict__ n_prev, int prev_stride,
              sloam_pose *__restrict__ curr_pose, sloam_kf_result *__restrict__ results,
              sloam_cylinder *__restrict__ tm, int32_t *__restrict__ tm_id, int32_t *__restrict__ matches,
              sloam_plane *__restrict__ planes_out, int32_t *__restrict__ n_planes_out, int planes_stride) {
  __shared__ sloam_pose s_pose;
  const sloam_params &P = dp->p;
  const int T = P.max_trees, B = dp->B;
  const int k = blockIdx.x;
  const int mode = kf_mode[k];
  if (threadIdx.x == 0) {
    sloam_kf_result r = results[k];
    sloam_pose cur = pose_est[k];  // currPose = in.poseEstimate (:498)
    if (mode == 0) {
      const double *x0 = lm_x + ((size_t)k * 2 + 0) * 8, *x1 = lm_x + ((size_t)k * 2 + 1) * 8;
      const int32_t *i0 = lm_info + ((size_t)k * 2 + 0) * 2, *i1 = lm_info + ((size_t)k * 2 + 1) * 2;
      if (two_step) {
        // TwoStepOptimizePose (:33-53): fall back to the estimate unless optimize && CONVERGENCE
        const sloam_pose T0 = pose_est[k];
        const double qw[4] = {T0.q[3], T0.q[0], T0.q[1], T0.q[2]};
        double rpy[3];
        quat_to_angle_axis(qw, rpy);
        const bool okT = optim_flags[2 * k] && i0[1] == 0, okG = optim_flags[2 * k + 1] && i1[1] == 0;
        const double treeOut[3] = {okT ? x0[0] : T0.t[0], okT ? x0[1] : T0.t[1], okT ? x0[5] : rpy[2]};
        const double groundOut[3] = {okG ? x1[2] : T0.t[2], okG ? x1[3] : rpy[0], okG ? x1[4] : rpy[1]};
        const double aa[3] = {groundOut[1], groundOut[2], treeOut[2]};
        double q[4];
        angle_axis_to_quat(aa, q);
        const double nq = sqrt(q[0] * q[0] + q[1] * q[1] + q[2] * q[2] + q[3] * q[3]);
        cur.q[0] = q[1] / nq; cur.q[1] = q[2] / nq; cur.q[2] = q[3] / nq; cur.q[3] = q[0] / nq;
        cur.t[0] = treeOut[0]; cur.t[1] = treeOut[1]; cur.t[2] = groundOut[0];
        r.lm_iterations[0] = i0[0]; r.lm_termination[0] = i0[1];
        r.lm_iterations[1] = i1[0]; r.lm_termination[1] = i1[1];
        r.success = 1;  // always returns true (:52)
      } else if (optim_flags[2 * k] && optim_flags[2 * k + 1]) {
        r.lm_iterations[0] = i0[0]; r.lm_termination[0] = i0[1];
        if (i0[1] == 0) {  // success: tf.setQuaternion (normalises), currPose = T_Delta (:241-246,:508-509)
          const double nq = sqrt(x0[0] * x0[0] + x0[1] * x0[1] + x0[2] * x0[2] + x0[3] * x0[3]);
          sloam_pose td;
          td.q[0] = x0[0] / nq; td.q[1] = x0[1] / nq; td.q[2] = x0[2] / nq; td.q[3] = x0[3] / nq;
          td.t[0] = x0[4]; td.t[1] = x0[5]; td.t[2] = x0[6];
          cur = td;
          r.T_Delta = td;
          r.success = 1;
        } else {
          r.success = 0;
          r.status = SLOAM_KF_NOT_CONVERGED | (r.status & ~0xFF);
        }
      }
    }
    if (mode != 2) r.T_Map_Curr = cur;
    results[k] = r;
    curr_pose[k] = cur;
    s_pose = cur;
  }
  __syncthreads();
  const sloam_pose cur = s_pose;
  if (mode == 2) {
    // RunSloam returned false before touching its state: prevGPlanes_ is unchanged
    const int np = min(n_prev[k], prev_stride);
    for (int g = threadIdx.x; g < np; g += 128) planes_out[(size_t)k * planes_stride + g] = prev_planes[(size_t)k * prev_stride + g];
    if (threadIdx.x == 0) n_planes_out[k] = np;
    return;
  }
  // projectModels (:438-451): Cylinder::project and Plane::project into the map frame
  const int nl = n_lm[k];
  for (int i = threadIdx.x; i < nl; i += 128) {
    sloam_cylinder c = lm_cyl[(size_t)k * T + i];
    double other[3] = {c.root[0] + c.ray[0], c.root[1] + c.ray[1], c.root[2] + c.ray[2]}, r2[3], o2[3];
    pose_apply(cur, c.root, r2);
    pose_apply(cur, other, o2);
    for (int a = 0; a < 3; ++a) { c.root[a] = r2[a]; c.ray[a] = o2[a] - r2[a]; }
    tm[(size_t)k * T + i] = c;
    tm_id[(size_t)k * T + i] = tree_models[(size_t)k * T + lm_src[(size_t)k * T + i]].id;
    matches[(size_t)k * T + i] = -1;
  }
  const int npl = n_planes_acc[k];
  for (int g = threadIdx.x; g < npl; g += 128) {
    const sloam_plane p = planes_acc[(size_t)k * B + g];
    sloam_plane o;
    plane_transform(cur, p.plane, o.plane);
    pose_apply(cur, p.centroid, o.centroid);
    planes_out[(size_t)k * planes_stride + g] = o;
  }
  if (threadIdx.x == 0) n_planes_out[k] = npl;
}

// matchModels (:298-328) on the map-frame distances of the second association
__global__ void matches_kernel(const DevParams *__restrict__ dp, const uint8_t *__restrict__ kf_mode,
                               const int32_t *__restrict__ n_lm, const int32_t *__restrict__ assoc_idx,
                               const double *__restrict__ assoc_dist, int32_t *__restrict__ matches) {
  const sloam_params &P = dp->p;
  const int T = P.max_trees, k = blockIdx.y;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (kf_mode[k] != 0 || i >= n_lm[k]) return;
  const int idx = assoc_idx[(size_t)k * T + i];
  const double d = assoc_dist[(size_t)k * T + i];
  matches[(size_t)k * T + i] = (idx >= 0 && d < P.treeMatchThresh + 100.0 && d < P.AddNewTreeThreshDist) ? idx : -1;
}

int launch_associate(sloam_ctx *c, int K, const sloam_cylinder *det, const int32_t *n_det, int det_stride,
                     int det_cap, const sloam_pose *tf, const sloam_cylinder *map, const int32_t *n_map,
                     int map_stride, int map_shared, int map_cap, int32_t *best_index, double *best_dist);

int launch_lm(sloam_ctx *c, int K, int two_step, const sloam_pose *pose_est, const double *tree_feat,
              const sloam_cylinder *tree_obj, const int32_t *n_tree_res, int tf_stride,
              const double *plane_feat, const sloam_plane *plane_obj, const int32_t *n_plane_res,
              int pf_stride, const uint8_t *optim_flags) {
  const int problems = K * (two_step ? 2 : 1);
  lm_kernel<<<problems, kLmThreads, 0, c->stream>>>(c->dp, two_step, K, pose_est, tree_feat, tree_obj, n_tree_res,
                                                tf_stride, plane_feat, plane_obj, n_plane_res, pf_stride,
                                                optim_flags, c->ws.lm_x, c->ws.lm_info);
  SB_LAUNCH_CHECK(c);
  return SLOAM_OK;
}

// everything of RunSloam after computeModels
int launch_sloam_core(sloam_ctx *c, int K, const sloam_batch_in *in, const sloam_batch_out *out) {
  Workspace &w = c->ws;
  const sloam_params &p = c->hp.p;
  const int T = p.max_trees, two_step = p.twoStepOptim ? 1 : 0;
  PROF_BEGIN(c, P_ASSOC_1);
  int rc = launch_associate(c, K, w.lm_cyl, w.n_lm, T, T, in->pose_est, in->map_models, in->n_map_models,
                            p.max_map_models, in->map_shared, p.max_map_models, w.assoc_idx, w.assoc_dist);
  PROF_END(c, P_ASSOC_1);
  if (rc != SLOAM_OK) return rc;
  PROF_BEGIN(c, P_BUILD_MATCHES);
  build_matches_kernel<<<K, 128, sizeof(int16_t) * (size_t)p.max_trees, c->stream>>>(
      c->dp, in->first_scan, in->pose_est, in->n_map_models, in->map_shared, in->map_models,
      p.max_map_models, in->prev_planes, in->n_prev_planes, p.max_prev_planes, w.n_lm, w.lm_src,
      w.assoc_idx, w.assoc_dist, w.tree_features, w.planes_acc, w.planes_acc_cell, w.n_planes_acc,
      w.cell_features, w.res_tree_feat, w.res_tree_obj, w.res_plane_feat, w.res_plane_obj, w.n_tree_res,
      w.n_plane_res, w.optim_flags, w.kf_mode, out->results, w.ground_count, w.n_trees,
      c->kf_flags_valid ? w.kf_flags : nullptr);
  PROF_END(c, P_BUILD_MATCHES);
  SB_LAUNCH_CHECK(c);
  PROF_BEGIN(c, P_LM);
  rc = launch_lm(c, K, two_step, in->pose_est, w.res_tree_feat, w.res_tree_obj, w.n_tree_res,
                 T * p.featuresPerTree, w.res_plane_feat, w.res_plane_obj, w.n_plane_res,
                 c->hp.B * p.numGroundFeatures, w.optim_flags);
  PROF_END(c, P_LM);
  if (rc != SLOAM_OK) return rc;
  PROF_BEGIN(c, P_FINISH);
  finish_kernel<<<K, 128, 0, c->stream>>>(c->dp, two_step, in->pose_est, w.kf_mode, w.optim_flags, w.lm_x,
                                          w.lm_info, w.n_lm, w.lm_cyl, w.lm_src, w.tree_models, w.planes_acc,
                                          w.n_planes_acc, in->prev_planes, in->n_prev_planes,
                                          p.max_prev_planes, w.curr_pose, out->results, out->tm, out->tm_id,
                                          out->matches, out->planes, out->n_planes, p.max_prev_planes);
  PROF_END(c, P_FINISH);
  SB_LAUNCH_CHECK(c);
  PROF_BEGIN(c, P_ASSOC_2);
  rc = launch_associate(c, K, w.lm_cyl, w.n_lm, T, T, w.curr_pose, in->map_models, in->n_map_models,
                        p.max_map_models, in->map_shared, p.max_map_models, w.assoc_idx, w.assoc_dist);
  if (rc != SLOAM_OK) return rc;
  dim3 g((unsigned)((T + 127) / 128), (unsigned)K);
  matches_kernel<<<g, 128, 0, c->stream>>>(c->dp, w.kf_mode, w.n_lm, w.assoc_idx, w.assoc_dist, out->matches);
  PROF_END(c, P_ASSOC_2);
  SB_LAUNCH_CHECK(c);
  return SLOAM_OK;
}

__global__ void pose_pack_kernel(int K, int two_step, const sloam_pose *__restrict__ pose_est,
                                 const uint8_t *__restrict__ flags, const double *__restrict__ lm_x, const int32_t *__restrict__ lm_info,
                                 sloam_pose *__restrict__ out_pose, int32_t *__restrict__ iterations,
                                 int32_t *__restrict__ termination) {
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= K) return;
  const double *x0 = lm_x + ((size_t)k * 2 + 0) * 8, *x1 = lm_x + ((size_t)k * 2 + 1) * 8;
  const int32_t *i0 = lm_info + ((size_t)k * 2 + 0) * 2, *i1 = lm_info + ((size_t)k * 2 + 1) * 2;
  sloam_pose cur = pose_est[k];
  iterations[2 * k] = i0[0]; termination[2 * k] = i0[1];
  iterations[2 * k + 1] = two_step ? i1[0] : 0; termination[2 * k + 1] = two_step ? i1[1] : -1;
  if (two_step) {
    const sloam_pose T0 = pose_est[k];
    const double qw[4] = {T0.q[3], T0.q[0], T0.q[1], T0.q[2]};
    double rpy[3];
    quat_to_angle_axis(qw, rpy);
    const bool okT = flags[2 * k] && i0[1] == 0, okG = flags[2 * k + 1] && i1[1] == 0;
    const double treeOut[3] = {okT ? x0[0] : T0.t[0], okT ? x0[1] : T0.t[1], okT ? x0[5] : rpy[2]};
    const double groundOut[3] = {okG ? x1[2] : T0.t[2], okG ? x1[3] : rpy[0], okG ? x1[4] : rpy[1]};
    const double aa[3] = {groundOut[1], groundOut[2], treeOut[2]};
    double q[4];
    angle_axis_to_quat(aa, q);
    const double nq = sqrt(q[0] * q[0] + q[1] * q[1] + q[2] * q[2] + q[3] * q[3]);
    cur.q[0] = q[1] / nq; cur.q[1] = q[2] / nq; cur.q[2] = q[3] / nq; cur.q[3] = q[0] / nq;
    cur.t[0] = treeOut[0]; cur.t[1] = treeOut[1]; cur.t[2] = groundOut[0];
  } else if (i0[1] == 0) {
    const double nq = sqrt(x0[0] * x0[0] + x0[1] * x0[1] + x0[2] * x0[2] + x0[3] * x0[3]);
    cur.q[0] = x0[0] / nq; cur.q[1] = x0[1] / nq; cur.q[2] = x0[2] / nq; cur.q[3] = x0[3] / nq;
    cur.t[0] = x0[4]; cur.t[1] = x0[5]; cur.t[2] = x0[6];
  }
  out_pose[k] = cur;
}

__global__ void flags_pack_kernel(int K, int mode, const uint8_t *ot, const uint8_t *og, uint8_t *flags) {
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= K) return;
  flags[2 * k] = mode == 0 ? 1 : (ot ? ot[k] : 1);
  flags[2 * k + 1] = mode == 0 ? 1 : (og ? og[k] : 1);
}

}  // namespace sb

using namespace sb;

extern "C" int sloam_b200_optimize_pose_dev(sloam_ctx *c, int K, int mode, const sloam_pose *pose_est,
                                            const double *tree_feat, const sloam_cylinder *tree_obj,
                                            const int32_t *n_tree_res, int tf_stride, const double *plane_feat,
                                            const sloam_plane *plane_obj, const int32_t *n_plane_res,
                                            int pf_stride, const uint8_t *optim_trees,
                                            const uint8_t *optim_ground, sloam_pose *out_pose,
                                            int32_t *iterations, int32_t *termination) {
  if (!c || K <= 0 || K > c->max_k || !pose_est || !tree_feat || !tree_obj || !n_tree_res || !plane_feat ||
      !plane_obj || !n_plane_res || !out_pose || !iterations || !termination || (mode != 0 && mode != 1))
    return set_err(c, SLOAM_E_INVALID, "optimize_pose: bad arguments");
  Workspace &w = c->ws;
  flags_pack_kernel<<<(K + 127) / 128, 128, 0, c->stream>>>(K, mode, optim_trees, optim_ground, w.optim_flags);
  SB_LAUNCH_CHECK(c);
  int rc = launch_lm(c, K, mode, pose_est, tree_feat, tree_obj, n_tree_res, tf_stride, plane_feat, plane_obj,
                     n_plane_res, pf_stride, w.optim_flags);
  if (rc != SLOAM_OK) return rc;
  // reuse kf-sized scratch for the unpacked flags
  pose_pack_kernel<<<(K + 127) / 128, 128, 0, c->stream>>>(K, mode, pose_est, w.optim_flags, w.lm_x,
                                                          w.lm_info, out_pose, iterations, termination);
  SB_LAUNCH_CHECK(c);
  return SLOAM_OK;
}
