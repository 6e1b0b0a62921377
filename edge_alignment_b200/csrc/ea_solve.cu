// Fused residual / analytic-Jacobian / normal-equation kernels and the persistent on-device
// Levenberg-Marquardt solve.  One thread-block cluster (1..8 CTAs, DSMEM reduction) owns one
// frame pair for its whole coarse-to-fine solve: no host round trip per iteration.
//
// Replaces, for n pairs at once: the residual-block loop + ceres::Solve of
// standalone_edge_align.cpp:265-291 (and src/SolveEA.cpp:163-198).
#include <cooperative_groups.h>

#include "ea_internal.h"
#include "ea_solve.cuh"

namespace cg = cooperative_groups;

#define EA_MAX_WARPS 32
#define EA_FLUSH_EVERY 4  // points per thread between fp32->fp64 flushes of the normal-equation slots

struct EaCtrl {  // double-buffered hand-off from the LM thread to every CTA of the cluster
  double cand[7];
  int cmd, pad;
};

struct EaSolveSmem {
  EaCtrl ctrl[2];
  double part[EA_MAX_WARPS][EA_NSUM];       // per-warp lane-slot sums
  double cpart[EA_MAX_WARPS];               // per-warp cost
  double cluster_sums[8][EA_SUMS + 3];      // rank 0 only: one row per cluster rank
  EaLmState lm;                             // rank 0 / thread 0 only
};

// Evaluate residual indices [j0, j1) (point index = j * stride) with the CTA's threads; leaves the CTA totals in
// out[0..EA_SUMS) (valid for all threads after the trailing __syncthreads()).
template <bool XYZ, int THREADS>
__device__ __forceinline__ void ea_eval_slice(const EaLevelDesc& rd, const EaLevelDesc& nd, const EaLevelGeom& rg,
                                              const EaLevelGeom& ng, double inv_depth_scale, const ea_solve_params& sp,
                                              const EaPose& P, int j0, int j1, EaSolveSmem& S, double* out) {
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const float loss_a = float(sp.loss_scale);
  const int loss_type = sp.loss_type, stride = sp.point_stride;
  float acc[EA_NSUM];
#pragma unroll
  for (int k = 0; k < EA_NSUM; ++k) acc[k] = 0.0f;
  double acc64 = 0.0, cost64 = 0.0;
  int since_flush = 0;
  for (int base = j0 + warp * 32; base < j1; base += THREADS) {
    const int j = base + lane;
    if (j < j1) {
      const float4 p = __ldg(rd.pts + size_t(j) * stride);
      EaPointEval e;
      ea_point_eval<XYZ>(p, rg, ng, inv_depth_scale, P, nd.dt, e);
      float rho0;
      const float w = ea_loss_eval(loss_type, loss_a, e.f, rho0);
      float J[6];
      ea_jacobian(e, ng, w, J);
      ea_accumulate(acc, J, e.f * w);
      acc[27] += e.fail ? 1.0f : 0.0f;
      cost64 += double(0.5f * rho0);
    }
    if (++since_flush == EA_FLUSH_EVERY) {
      acc64 += double(ea_warp_transpose_reduce(acc, lane));
#pragma unroll
      for (int k = 0; k < EA_NSUM; ++k) acc[k] = 0.0f;
      since_flush = 0;
    }
  }
  if (since_flush) acc64 += double(ea_warp_transpose_reduce(acc, lane));
  cost64 = ea_warp_sum(cost64);
  S.part[warp][lane] = acc64;
  if (lane == 0) S.cpart[warp] = cost64;
  __syncthreads();
  if (tid < EA_SUMS) {
    double s = 0.0;
    if (tid < 28) {
#pragma unroll 1
      for (int w2 = 0; w2 < THREADS / 32; ++w2) s += S.part[w2][tid];
    } else {
#pragma unroll 1
      for (int w2 = 0; w2 < THREADS / 32; ++w2) s += S.cpart[w2];
    }
    out[tid] = s;
  }
}

template <int THREADS, bool CLUSTER>
__global__ void __launch_bounds__(THREADS) ea_k_solve_batch(const __grid_constant__ EaSolveArgs A) {
  __shared__ EaSolveSmem S;
  __shared__ double cta_sums[EA_SUMS + 3];
  cg::cluster_group cluster = cg::this_cluster();
  const unsigned crank = CLUSTER ? cluster.block_rank() : 0u;
  const unsigned csize = CLUSTER ? cluster.num_blocks() : 1u;
  const int cluster_id = blockIdx.x / csize, n_clusters = gridDim.x / csize;
  const int tid = threadIdx.x;
  const bool boss = (crank == 0 && tid == 0);
  EaSolveSmem* S0 = CLUSTER ? cluster.map_shared_rank(&S, 0) : &S;
  auto sync_all = [&]() { if (CLUSTER) cluster.sync(); else __syncthreads(); };
  unsigned g = 0;  // control-slot parity, uniform across the cluster

  for (int pair = cluster_id; pair < A.n_pairs; pair += n_clusters) {
    const int rs = A.ref_slots[pair], ns = A.now_slots[pair];
    const int pi = A.pose_index ? A.pose_index[pair] : pair;
    if (boss) {
#pragma unroll 1
      for (int i = 0; i < 7; ++i) S.lm.x[i] = A.poses[size_t(pi) * 7 + i];
    }
    for (int level = A.coarsest; level >= A.finest; --level) {
      const EaLevelDesc rd = A.ref_desc[size_t(rs) * EA_MAX_LEVELS + level];
      const EaLevelDesc nd = A.now_desc[size_t(ns) * EA_MAX_LEVELS + level];
      const EaLevelGeom& rg = A.ref_geom[level];
      const EaLevelGeom& ng = A.now_geom[level];
      const int n_pts = min(*rd.n_pts, A.ref_cap[level]);
      const int n_res = (n_pts + A.sp.point_stride - 1) / A.sp.point_stride;
      ea_summary* sum_out = A.summaries ? A.summaries + (size_t(pair) * A.n_levels + level) : nullptr;
      if (n_res == 0) {
        if (boss && sum_out) { ea_summary z = {}; z.termination = EA_TERM_SKIPPED_NO_POINTS; *sum_out = z; }
        continue;
      }
      const int j0 = int((long long)n_res * crank / csize), j1 = int((long long)n_res * (crank + 1) / csize);
      // level start: publish x as the first candidate in a control slot nobody is reading
      g += 1;
      if (boss) {
        EaLmState& L = S.lm;
        L.phase = 0; L.iter = 0; L.accepted = 0; L.rejected = 0; L.invalid_run = 0; L.evals = 0; L.term = EA_TERM_NONE;
#pragma unroll 1
        for (int i = 0; i < 7; ++i) L.cand[i] = L.x[i];
        for (unsigned r = 0; r < csize; ++r) {
          EaSolveSmem* Sr = CLUSTER ? cluster.map_shared_rank(&S, r) : &S;
#pragma unroll 1
          for (int i = 0; i < 7; ++i) Sr->ctrl[g & 1].cand[i] = L.x[i];
          Sr->ctrl[g & 1].cmd = EA_CMD_EVAL;
        }
      }
      sync_all();
      for (;;) {
        const EaCtrl& C = S.ctrl[g & 1];
        if (C.cmd == EA_CMD_DONE) break;
        EaPose P;
        ea_pose_from_q(C.cand, P);
        if (rd.pts_mode == EA_POINTS_XYZ)
          ea_eval_slice<true, THREADS>(rd, nd, rg, ng, A.inv_depth_scale, A.sp, P, j0, j1, S, cta_sums);
        else
          ea_eval_slice<false, THREADS>(rd, nd, rg, ng, A.inv_depth_scale, A.sp, P, j0, j1, S, cta_sums);
        if (CLUSTER) {
          __syncthreads();
          if (tid < EA_SUMS) S0->cluster_sums[crank][tid] = cta_sums[tid];
        }
        sync_all();
        if (boss) {
          double sums[EA_SUMS];
          if (CLUSTER) {
#pragma unroll 1
            for (int k = 0; k < EA_SUMS; ++k) {
              double s = 0.0;
              for (unsigned r = 0; r < csize; ++r) s += S.cluster_sums[r][k];
              sums[k] = s;
            }
          } else {
#pragma unroll 1
            for (int k = 0; k < EA_SUMS; ++k) sums[k] = cta_sums[k];
          }
          const int cmd = ea_lm_advance(S.lm, sums, A.sp);
          const unsigned nx = (g + 1) & 1;
          for (unsigned r = 0; r < csize; ++r) {
            EaSolveSmem* Sr = CLUSTER ? cluster.map_shared_rank(&S, r) : &S;
            if (cmd == EA_CMD_EVAL) {
#pragma unroll 1
              for (int i = 0; i < 7; ++i) Sr->ctrl[nx].cand[i] = S.lm.cand[i];
            }
            Sr->ctrl[nx].cmd = cmd;
          }
        }
        sync_all();
        g += 1;
      }
      if (boss && sum_out) {
        const EaLmState& L = S.lm;
        ea_summary z;
        z.termination = L.term; z.iterations = L.iter; z.accepted = L.accepted; z.rejected = L.rejected;
        z.n_residuals = n_res; z.evaluations = L.evals; z.initial_cost = L.initial_cost; z.final_cost = L.cost;
        *sum_out = z;
      }
    }
    if (boss) {
#pragma unroll 1
      for (int i = 0; i < 7; ++i) A.poses[size_t(pi) * 7 + i] = S.lm.x[i];
    }
  }
  if (CLUSTER) cluster.sync();  // no CTA may exit while a peer can still touch its shared memory
}

// Per-point outputs for ea_eval (parity tests / EAResidue facade): same device functions as the solve.
template <bool XYZ>
__global__ void __launch_bounds__(256) ea_k_eval_points(EaLevelDesc rd, EaLevelDesc nd, EaLevelGeom rg, EaLevelGeom ng,
                                                        double inv_depth_scale, ea_solve_params sp, const double* pose7,
                                                        int n_res, double* raw, double* res, double* jac, int* failed) {
  EaPose P;
  ea_pose_from_q(pose7, P);
  for (int j = blockIdx.x * blockDim.x + threadIdx.x; j < n_res; j += gridDim.x * blockDim.x) {
    const float4 p = __ldg(rd.pts + size_t(j) * sp.point_stride);
    EaPointEval e;
    ea_point_eval<XYZ>(p, rg, ng, inv_depth_scale, P, nd.dt, e);
    float rho0;
    const float w = ea_loss_eval(sp.loss_type, float(sp.loss_scale), e.f, rho0);
    float J[6];
    ea_jacobian(e, ng, w, J);
    if (raw) raw[j] = double(e.f);
    if (res) res[j] = double(e.f * w);
    if (jac) {
#pragma unroll
      for (int k = 0; k < 6; ++k) jac[size_t(j) * 6 + k] = double(J[k]);
    }
    if (e.fail && failed) atomicAdd(failed, 1);
  }
}

// Normal-equation sums through the production reduction path (one cluster-less CTA per call slice).
template <int THREADS>
__global__ void __launch_bounds__(THREADS) ea_k_eval_sums(EaLevelDesc rd, EaLevelDesc nd, EaLevelGeom rg, EaLevelGeom ng,
                                                          double inv_depth_scale, ea_solve_params sp,
                                                          const double* pose7, int n_res, double* sums /*[grid][EA_SUMS]*/) {
  __shared__ EaSolveSmem S;
  __shared__ double cta_sums[EA_SUMS + 3];
  EaPose P;
  ea_pose_from_q(pose7, P);
  const int j0 = int((long long)n_res * blockIdx.x / gridDim.x), j1 = int((long long)n_res * (blockIdx.x + 1) / gridDim.x);
  if (rd.pts_mode == EA_POINTS_XYZ)
    ea_eval_slice<true, THREADS>(rd, nd, rg, ng, inv_depth_scale, sp, P, j0, j1, S, cta_sums);
  else
    ea_eval_slice<false, THREADS>(rd, nd, rg, ng, inv_depth_scale, sp, P, j0, j1, S, cta_sums);
  __syncthreads();
  if (threadIdx.x < EA_SUMS) sums[size_t(blockIdx.x) * EA_SUMS + threadIdx.x] = cta_sums[threadIdx.x];
}

// ---- host launchers -----------------------------------------------------------------------------------
#define EA_SOLVE_THREADS 512

cudaError_t ea_launch_solve_batch(const EaSolveArgs& A, int cluster_size, int sm_count, cudaStream_t stream) {
  if (A.n_pairs <= 0) return cudaSuccess;
  cudaLaunchConfig_t cfg = {};
  cfg.blockDim = dim3(EA_SOLVE_THREADS);
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  if (cluster_size <= 1) {
    // persistent: at most one CTA per SM worth of pairs in flight; CTAs loop over pairs
    cfg.gridDim = dim3(unsigned(A.n_pairs < sm_count * 2 ? A.n_pairs : sm_count * 2));
    cfg.numAttrs = 0;
    return cudaLaunchKernelEx(&cfg, ea_k_solve_batch<EA_SOLVE_THREADS, false>, A);
  }
  int n_clusters = A.n_pairs;
  const int max_clusters = (sm_count / cluster_size) * 2;
  if (n_clusters > max_clusters) n_clusters = max_clusters;
  if (n_clusters < 1) n_clusters = 1;
  cfg.gridDim = dim3(unsigned(n_clusters * cluster_size));
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = unsigned(cluster_size);
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  return cudaLaunchKernelEx(&cfg, ea_k_solve_batch<EA_SOLVE_THREADS, true>, A);
}

cudaError_t ea_launch_eval_points(const EaLevelDesc& rd, const EaLevelDesc& nd, const EaLevelGeom& rg,
                                  const EaLevelGeom& ng, double inv_depth_scale, const ea_solve_params& sp,
                                  const double* d_pose7, int n_res, double* d_raw, double* d_res, double* d_jac,
                                  int* d_failed, cudaStream_t stream) {
  if (n_res <= 0) return cudaSuccess;
  const int blocks = (n_res + 255) / 256;
  if (rd.pts_mode == EA_POINTS_XYZ)
    ea_k_eval_points<true><<<blocks, 256, 0, stream>>>(rd, nd, rg, ng, inv_depth_scale, sp, d_pose7, n_res, d_raw, d_res, d_jac, d_failed);
  else
    ea_k_eval_points<false><<<blocks, 256, 0, stream>>>(rd, nd, rg, ng, inv_depth_scale, sp, d_pose7, n_res, d_raw, d_res, d_jac, d_failed);
  return cudaGetLastError();
}

cudaError_t ea_launch_eval_sums(const EaLevelDesc& rd, const EaLevelDesc& nd, const EaLevelGeom& rg,
                                const EaLevelGeom& ng, double inv_depth_scale, const ea_solve_params& sp,
                                const double* d_pose7, int n_res, int n_blocks, double* d_sums, cudaStream_t stream) {
  if (n_res <= 0 || n_blocks <= 0) return cudaSuccess;
  ea_k_eval_sums<EA_SOLVE_THREADS><<<n_blocks, EA_SOLVE_THREADS, 0, stream>>>(rd, nd, rg, ng, inv_depth_scale, sp, d_pose7, n_res, d_sums);
  return cudaGetLastError();
}
