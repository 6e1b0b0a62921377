// Fused residual / analytic-Jacobian / normal-equation kernels and the persistent on-device
// Levenberg-Marquardt solve.  One thread-block cluster (1..8 CTAs, DSMEM reduction) owns one
// frame pair for its whole coarse-to-fine solve: no host round trip per iteration.
//
// Replaces, for n pairs at once: the residual-block loop + ceres::Solve of
// standalone_edge_align.cpp:265-291 (and src/SolveEA.cpp:163-198).
//
// Control flow: the "boss" (thread 0 of cluster rank 0) owns the LM state machine and the work queue.  Every loop
// iteration of every CTA is exactly one evaluation: read the boss's message {point list, DT, candidate pose}, reduce
// the CTA's slice of the points to 29 sums, hand them to the boss, who advances LM / the pyramid level / the pair
// and publishes the next message.  Messages are parity double-buffered, so two barriers per evaluation suffice.
// Several small CTAs share an SM, so one pair's serial LM step overlaps the other pairs' evaluations.
#include <cooperative_groups.h>

#include "ea_internal.h"
#include <cstdlib>

#include "ea_solve.cuh"

namespace cg = cooperative_groups;

#ifndef EA_SOLVE_THREADS
#define EA_SOLVE_THREADS 512
#endif
#define EA_SOLVE_WARPS (EA_SOLVE_THREADS / 32)
#ifndef EA_SOLVE_MIN_CTAS
#define EA_SOLVE_MIN_CTAS 1
#endif
// points per thread between fp32->fp64 flushes of the normal-equation slots
#define EA_CMD_EXIT 2

struct EaMsg {  // boss -> all threads of the cluster
  double cand[7];
  const void* pts;
  const float* dt;
  float2 affine;
  int n_res, level, pts_mode, cmd;
  int rev, pad;   // sweep direction of this evaluation: alternates per evaluation of the (pair, level), see ea_eval_slice
};

struct EaSolveSmem {
  EaMsg msg[2];
  double part[EA_SOLVE_WARPS][EA_NSUM];     // per-warp lane-slot sums
  double cpart[EA_SOLVE_WARPS];             // per-warp cost
  double sums[EA_SUMS + 3];                 // CTA totals
  double cluster_sums[8][EA_SUMS + 3];      // rank 0 only: one row per cluster rank
  EaLmState lm;                             // boss only
  int pair, level;                          // boss only
};

// Boss: move to the next (pair, level) that has points and publish its first evaluation, or EXIT.
template <bool INL>
__device__ __forceinline__ void ea_boss_next_impl(const EaSolveArgs& A, EaSolveSmem& S, EaMsg& out, bool new_pair_needed) {
  for (;;) {
    if (new_pair_needed) {
      const int ticket = atomicAdd(A.work_counter, 1);
      if (ticket >= A.n_pairs) { out.cmd = EA_CMD_EXIT; S.pair = -1; return; }
      const int pair = A.order ? A.order[ticket] : ticket;   // longest-expected-first when the caller knows better
      S.pair = pair; S.level = A.coarsest;
      const int pi = A.pose_index ? A.pose_index[pair] : pair;
#pragma unroll 1
      for (int i = 0; i < 7; ++i) S.lm.x[i] = A.poses[size_t(pi) * 7 + i];
      new_pair_needed = false;
    }
    const int pair = S.pair, level = S.level;
    if (level < A.finest) {  // pair finished: write the pose back, fetch another
      const int pi = A.pose_index ? A.pose_index[pair] : pair;
#pragma unroll 1
      for (int i = 0; i < 7; ++i) A.poses[size_t(pi) * 7 + i] = S.lm.x[i];
      new_pair_needed = true;
      continue;
    }
    const EaLevelDesc rd = A.ref_desc[size_t(A.ref_slots[pair]) * EA_MAX_LEVELS + level];
    const EaLevelDesc nd = A.now_desc[size_t(A.now_slots[pair]) * EA_MAX_LEVELS + level];
    const int n_pts = min(*rd.n_pts, A.ref_cap[level]);
    const int n_res = (n_pts + A.sp.point_stride - 1) / A.sp.point_stride;
    if (n_res == 0) {
      if (A.summaries) { ea_summary z = {}; z.termination = EA_TERM_SKIPPED_NO_POINTS; A.summaries[size_t(pair) * A.n_levels + level] = z; }
      S.level = level - 1;
      continue;
    }
    EaLmState& L = S.lm;
    L.phase = 0; L.iter = 0; L.accepted = 0; L.rejected = 0; L.invalid_run = 0; L.evals = 0; L.term = EA_TERM_NONE;
#pragma unroll 1
    for (int i = 0; i < 7; ++i) { L.cand[i] = L.x[i]; out.cand[i] = L.x[i]; }
    out.pts = rd.pts; out.dt = nd.dt; out.affine = *nd.dt_affine; out.n_res = n_res; out.level = level; out.pts_mode = rd.pts_mode; out.cmd = EA_CMD_EVAL; out.rev = 0; out.pad = 0;
    return;
  }
}

// Boss: consume the sums of one evaluation; fills `out` with the next message.
__device__ __noinline__ void ea_boss_next(const EaSolveArgs& A, EaSolveSmem& S, EaMsg& out, bool new_pair_needed) { ea_boss_next_impl<false>(A, S, out, new_pair_needed); }

template <bool INL>
__device__ __forceinline__ void ea_boss_step_impl(const EaSolveArgs& A, EaSolveSmem& S, const double* sums, const EaMsg& cur, EaMsg& out) {
  const int acc0 = S.lm.accepted, rej0 = S.lm.rejected;
  const int cmd = INL ? ea_lm_advance_impl(S.lm, sums, A.sp) : ea_lm_advance(S.lm, sums, A.sp);
  if (A.trace) {   // iteration log of a single-pair solve (ea_solve_traced): what Ceres prints with minimizer_progress_to_stdout
    const int k = (*A.trace_count)++;
    if (k < A.trace_cap) {
      double* r = A.trace + size_t(k) * EA_TRACE_DOUBLES;
      r[0] = cur.level; r[1] = S.lm.accepted + S.lm.rejected; r[2] = S.lm.cost; r[3] = sums[28]; r[4] = S.lm.radius;
      r[5] = S.lm.accepted > acc0 ? 1.0 : (S.lm.rejected > rej0 ? 0.0 : -1.0);   // accepted / rejected / no decision (first evaluation, termination)
    }
  }
  if (cmd == EA_CMD_EVAL) {
#pragma unroll
    for (int i = 0; i < 7; ++i) out.cand[i] = S.lm.cand[i];
    out.pts = cur.pts; out.dt = cur.dt; out.affine = cur.affine; out.n_res = cur.n_res; out.level = cur.level; out.pts_mode = cur.pts_mode; out.cmd = EA_CMD_EVAL; out.rev = S.lm.evals & 1; out.pad = 0;
    return;
  }
  if (A.summaries) {
    const EaLmState& L = S.lm;
    ea_summary z;
    z.termination = L.term; z.iterations = L.iter; z.accepted = L.accepted; z.rejected = L.rejected;
    z.n_residuals = cur.n_res; z.evaluations = L.evals; z.initial_cost = L.initial_cost; z.final_cost = L.cost;
    A.summaries[size_t(S.pair) * A.n_levels + cur.level] = z;
  }
  S.level = cur.level - 1;
  if (INL) ea_boss_next_impl<true>(A, S, out, false); else ea_boss_next(A, S, out, false);
}
__device__ __noinline__ void ea_boss_step(const EaSolveArgs& A, EaSolveSmem& S, const double* sums, const EaMsg& cur, EaMsg& out) { ea_boss_step_impl<false>(A, S, sums, cur, out); }


#ifndef EA_ALTERNATE_SWEEP
#define EA_ALTERNATE_SWEEP 1
#endif

template <int THREADS, bool CLUSTER>
__global__ void __launch_bounds__(THREADS, EA_SOLVE_MIN_CTAS) ea_k_solve_batch(const __grid_constant__ EaSolveArgs A) {
  __shared__ EaSolveSmem S;
  cg::cluster_group cluster = cg::this_cluster();
  const unsigned crank = CLUSTER ? cluster.block_rank() : 0u;
  const unsigned csize = CLUSTER ? cluster.num_blocks() : 1u;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const bool boss = (crank == 0 && tid == 0);
  auto sync_all = [&]() { if (CLUSTER) cluster.sync(); else __syncthreads(); };
  auto publish = [&](const EaMsg& m, unsigned slot) {
    for (unsigned r = 0; r < csize; ++r) {
      EaSolveSmem* Sr = CLUSTER ? cluster.map_shared_rank(&S, r) : &S;
      Sr->msg[slot] = m;
    }
  };
  unsigned g = 0;
  if (boss) {
    EaMsg m;
    ea_boss_next(A, S, m, true);
    publish(m, 0);
  }
  sync_all();
  for (;;) {
    const EaMsg& M = S.msg[g & 1];
    if (M.cmd == EA_CMD_EXIT) break;
    const EaLevelGeom& rg = A.ref_geom[M.level];
    const EaLevelGeom& ng = A.now_geom[M.level];
    const int j0 = int((long long)M.n_res * crank / csize), j1 = int((long long)M.n_res * (crank + 1) / csize);
    EaPose P;
    if (M.pts_mode == EA_POINTS_XYZ) {
      ea_pose_setup<true>(M.cand, rg, ng, P);
      ea_eval_slice<true, THREADS>(M.pts, M.dt, M.affine, ng, A.inv_depth_scale, A.sp, P, j0, j1, S.part, S.cpart, EA_ALTERNATE_SWEEP && M.rev);
    } else {
      ea_pose_setup<false>(M.cand, rg, ng, P);
      ea_eval_slice<false, THREADS>(M.pts, M.dt, M.affine, ng, A.inv_depth_scale, A.sp, P, j0, j1, S.part, S.cpart, EA_ALTERNATE_SWEEP && M.rev);
    }
    __syncthreads();
    if (warp == 0) {
      const double tot = ea_cta_total<THREADS / 32>(S.part, S.cpart, lane);
      if (CLUSTER) {
        EaSolveSmem* S0 = cluster.map_shared_rank(&S, 0);
        if (lane < EA_SUMS) S0->cluster_sums[crank][lane] = tot;
      } else {
        if (lane < EA_SUMS) S.sums[lane] = tot;
        __syncwarp();
        if (lane == 0) {
          EaMsg m;
          ea_boss_step(A, S, S.sums, M, m);
          S.msg[(g + 1) & 1] = m;
        }
      }
    }
    if (CLUSTER) {
      cluster.sync();
      if (crank == 0 && warp == 0) {
        double s = 0.0;
        if (lane < EA_SUMS) for (unsigned r = 0; r < csize; ++r) s += S.cluster_sums[r][lane];
        if (lane < EA_SUMS) S.sums[lane] = s;
        __syncwarp();
        if (lane == 0) {
          EaMsg m;
          ea_boss_step(A, S, S.sums, M, m);
          publish(m, (g + 1) & 1);
        }
      }
    }
    sync_all();
    g += 1;
  }
  if (CLUSTER) cluster.sync();  // no CTA may exit while a peer can still touch its shared memory
}

// Per-point outputs for ea_eval (parity tests / EAResidue facade): same device functions as the solve.
template <bool XYZ>
__global__ void __launch_bounds__(256) ea_k_eval_points(EaLevelDesc rd, EaLevelDesc nd, EaLevelGeom rg, EaLevelGeom ng,
                                                        double inv_depth_scale, ea_solve_params sp, const double* pose7,
                                                        int n_res, double* raw, double* res, double* jac, int* failed) {
  EaPose P;
  ea_pose_setup<XYZ>(pose7, rg, ng, P);
  const float2 affine = *nd.dt_affine;
  const int n_round = (n_res + 31) & ~31;   // keep warps converged: the gather's fast path votes with __all_sync
  for (int j = blockIdx.x * blockDim.x + threadIdx.x; j < n_round; j += gridDim.x * blockDim.x) {
    const bool valid = j < n_res;
    typedef EaPtStream<XYZ> PS;
    const typename PS::T p = valid ? PS::load(rd.pts, size_t(j) * sp.point_stride) : PS::pad();
    EaPointEval e;
    double a0, a1, a2;
    PS::unpack(p, a0, a1, a2);
    ea_point_eval<XYZ>(a0, a1, a2, ng, inv_depth_scale, P, nd.dt, affine, e);
    float rho0;
    const float w = ea_loss_eval(sp.loss_type, float(sp.loss_scale), e.f, rho0);
    float J[6];
    ea_jacobian(e, P, w, J);
    if (!valid) continue;
    if (raw) raw[j] = double(e.f);
    if (res) res[j] = double(e.f * w);
    if (jac) {
#pragma unroll
      for (int k = 0; k < 6; ++k) jac[size_t(j) * 6 + k] = double(J[k]);
    }
    if (e.fail && failed) atomicAdd(failed, 1);
  }
}

// Normal-equation sums through the production reduction path (one CTA per slice of the points).
template <int THREADS>
__global__ void __launch_bounds__(THREADS) ea_k_eval_sums(EaLevelDesc rd, EaLevelDesc nd, EaLevelGeom rg, EaLevelGeom ng,
                                                          double inv_depth_scale, ea_solve_params sp, const double* pose7,
                                                          const int* done, int j_begin, int j_end, double* sums /*[grid][EA_SUMS]*/) {
  if (done && *done) return;
  __shared__ double part[THREADS / 32][EA_NSUM];
  __shared__ double cpart[THREADS / 32];
  const int n = j_end - j_begin;
  const int j0 = j_begin + int((long long)n * blockIdx.x / gridDim.x), j1 = j_begin + int((long long)n * (blockIdx.x + 1) / gridDim.x);
  EaPose P;
  if (rd.pts_mode == EA_POINTS_XYZ) {
    ea_pose_setup<true>(pose7, rg, ng, P);
    ea_eval_slice<true, THREADS>(rd.pts, nd.dt, *nd.dt_affine, ng, inv_depth_scale, sp, P, j0, j1, part, cpart);
  } else {
    ea_pose_setup<false>(pose7, rg, ng, P);
    ea_eval_slice<false, THREADS>(rd.pts, nd.dt, *nd.dt_affine, ng, inv_depth_scale, sp, P, j0, j1, part, cpart);
  }
  __syncthreads();
  if (threadIdx.x < 32) {
    const double tot = ea_cta_total<THREADS / 32>(part, cpart, threadIdx.x);
    if (threadIdx.x < EA_SUMS) sums[size_t(blockIdx.x) * EA_SUMS + threadIdx.x] = tot;
  }
}

// ---- warp-specialised solve kernel -------------------------------------------------------------------------------------
// tools/probe_occupancy.sh: the memory side of an evaluation (point stream, projection, 16-texel gather) is latency-bound
// and scales with resident warps (2.2 / 3.6 / 5.2 ms per launch at 32 / 16 / 8 warps per SM); the fused loop needs 128
// registers and therefore runs 16.  Here the CTA has 24 warps: warpgroup 0 (4 "math" warps, 160 registers after
// setmaxnreg) and warpgroups 1..5 (20 "gather" warps, 64 registers).  Gather warp (p, m) projects the points of
// warp-iterations k = 4 q + m, gathers their texels and hands {16 texels, du, dv, ub, vb, z', 1/z', flags} per point to math
// warp m through a 5-slot shared-memory ring (one slot per producer, mbarrier full/empty pairs); the math warp
// interpolates, builds the Jacobian and accumulates in a fixed order (k ascending), so the sums stay reproducible.
#ifndef EA_WS_MATH
#define EA_WS_MATH 4          // math warps (a multiple of 4: setmaxnreg works on warpgroups)
#endif
#ifndef EA_WS_PROD
#define EA_WS_PROD 5          // gather warps per math warp
#endif
#ifndef EA_WS_MATH_REGS
#define EA_WS_MATH_REGS 160
#endif
#ifndef EA_WS_GATHER_REGS
#define EA_WS_GATHER_REGS 64
#endif
#ifndef EA_WS_BACKOFF_NS
#define EA_WS_BACKOFF_NS 0
#endif
#define EA_WS_THREADS (32 * EA_WS_MATH * (1 + EA_WS_PROD))
#define EA_WS_WORDS 24
#define EA_STR2(x) #x
#define EA_STR(x) EA_STR2(x)

struct EaWsRing {
  unsigned long long full[EA_WS_MATH][EA_WS_PROD];
  unsigned long long empty[EA_WS_MATH][EA_WS_PROD];
  float slot[EA_WS_MATH][EA_WS_PROD][EA_WS_WORDS][32];
};

__device__ __forceinline__ unsigned ea_smem_u32(const void* p) { return unsigned(__cvta_generic_to_shared(p)); }
__device__ __forceinline__ void ea_mbar_init(unsigned long long* b, unsigned count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(ea_smem_u32(b)), "r"(count) : "memory");
}
__device__ __forceinline__ void ea_mbar_arrive(unsigned long long* b) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(ea_smem_u32(b)) : "memory");
}
__device__ __forceinline__ void ea_mbar_wait(unsigned long long* b, unsigned parity) {
  const unsigned a = ea_smem_u32(b);
  unsigned done;
  do {
    asm volatile("{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n selp.u32 %0, 1, 0, p;\n}"
                 : "=r"(done) : "r"(a), "r"(parity) : "memory");
#if EA_WS_BACKOFF_NS > 0
    if (!done) __nanosleep(EA_WS_BACKOFF_NS);     // polling warps would otherwise take issue slots from the warps they wait for
#endif
  } while (!done);
}

// warp-iterations of an evaluation that belong to math warp m: k = 4 q + m, q in [0, Qm)
__device__ __forceinline__ int ea_ws_count(int n_res, int m) {
  const int K = (n_res + 31) >> 5;
  return K > m ? (K - m + EA_WS_MATH - 1) / EA_WS_MATH : 0;
}

template <bool XYZ>
__device__ __forceinline__ void ea_ws_produce(const void* __restrict__ pts, const float* __restrict__ dt, const int n_res, const int stride,
                                              const EaLevelGeom& ng, const double inv_depth_scale, const EaPose& P, EaWsRing& R,
                                              const int m, const int p, const int lane, unsigned& seq) {
  typedef EaPtStream<XYZ> PS;
  const int Qm = ea_ws_count(n_res, m);
  const unsigned n0 = seq;
  seq = n0 + unsigned(Qm);
  int q = int((unsigned(p) + EA_WS_PROD - (n0 % EA_WS_PROD)) % EA_WS_PROD);   // first q with (n0 + q) % 5 == p
  if (q >= Qm) return;
  int j = ((q * EA_WS_MATH + m) << 5) + lane;
  typename PS::T p_next = j < n_res ? PS::load(pts, size_t(j) * stride) : PS::pad();
  float* const s = &R.slot[m][p][0][lane];
  for (; q < Qm; q += EA_WS_PROD) {
    const unsigned n = n0 + unsigned(q);
    const typename PS::T pt = p_next;
    const bool valid = j < n_res;
    j += EA_WS_PROD * EA_WS_MATH * 32;
    if (j < n_res) p_next = PS::load(pts, size_t(j) * stride);
    double a0, a1, a2;
    PS::unpack(pt, a0, a1, a2);
    EaProj r;
    ea_project<XYZ>(a0, a1, a2, ng, inv_depth_scale, P, r);
    float t[16];
    ea_gather(r, ng.w, ng.h, dt, t);
    ea_mbar_wait(&R.empty[m][p], ((n / EA_WS_PROD) & 1u) ^ 1u);     // the math warp is done with this slot's previous content
#pragma unroll
    for (int k = 0; k < 16; ++k) s[k * 32] = t[k];
    s[16 * 32] = r.du; s[17 * 32] = r.dv; s[18 * 32] = r.ub; s[19 * 32] = r.vb; s[20 * 32] = r.pz; s[21 * 32] = r.iz;
    s[22 * 32] = __int_as_float((valid ? 1 : 0) | (r.fail ? 2 : 0));
    __syncwarp();
    if (lane == 0) ea_mbar_arrive(&R.full[m][p]);
  }
}

__device__ __forceinline__ void ea_ws_consume(const int n_res, const float2 affine, const ea_solve_params& sp, const EaPose& P, EaWsRing& R,
                                              const int m, const int lane, unsigned& seq, double (*part)[EA_NSUM], double* cpart) {
  const int Qm = ea_ws_count(n_res, m);
  const unsigned n0 = seq;
  seq = n0 + unsigned(Qm);
  const float loss_a = float(sp.loss_scale);
  const int loss_type = sp.loss_type;
  float acc[EA_NSUM];
#pragma unroll
  for (int k = 0; k < EA_NSUM; ++k) acc[k] = 0.0f;
  double acc64 = 0.0, cost64 = 0.0;
  int since_flush = 0;
  for (int q = 0; q < Qm; ++q) {
    const unsigned n = n0 + unsigned(q);
    const unsigned slot = n % EA_WS_PROD;
    ea_mbar_wait(&R.full[m][slot], (n / EA_WS_PROD) & 1u);
    const float* const s = &R.slot[m][slot][0][lane];
    float t[16];
#pragma unroll
    for (int k = 0; k < 16; ++k) t[k] = s[k * 32];
    EaPointEval e;
    const float du = s[16 * 32], dv = s[17 * 32];
    e.ub = s[18 * 32]; e.vb = s[19 * 32]; e.pz = s[20 * 32]; e.iz = s[21 * 32];
    const int flags = __float_as_int(s[22 * 32]);
    ea_interp(t, du, dv, affine, e.f, e.dfdu, e.dfdv);
    const bool valid = (flags & 1) != 0;
    e.fail = (flags & 2) != 0;
    float rho0;
    float w = ea_loss_eval(loss_type, loss_a, e.f, rho0);
    if (!valid) { w = 0.0f; rho0 = 0.0f; e.f = 0.0f; e.fail = false; }
    float J[6];
    ea_jacobian(e, P, w, J);
    if (!valid) {
#pragma unroll
      for (int k = 0; k < 6; ++k) J[k] = 0.0f;
    }
    __syncwarp();                                           // every lane has consumed what it read from the slot
    if (lane == 0) ea_mbar_arrive(&R.empty[m][slot]);
    ea_accumulate(acc, J, e.f * w);
    acc[27] += e.fail ? 1.0f : 0.0f;
    cost64 += double(0.5f * rho0);
    if (++since_flush == EA_FLUSH_EVERY) {
      acc64 += double(ea_warp_transpose_reduce(acc, lane));
#pragma unroll
      for (int k = 0; k < EA_NSUM; ++k) acc[k] = 0.0f;
      since_flush = 0;
    }
  }
  if (since_flush) acc64 += double(ea_warp_transpose_reduce(acc, lane));
  cost64 = ea_warp_sum(cost64);
  part[m][lane] = acc64;
  if (lane == 0) cpart[m] = cost64;
}

__global__ void __launch_bounds__(EA_WS_THREADS, 1) ea_k_solve_ws(const __grid_constant__ EaSolveArgs A) {
  __shared__ EaSolveSmem S;
  extern __shared__ __align__(16) unsigned char ea_ws_raw[];
  EaWsRing& R = *reinterpret_cast<EaWsRing*>(ea_ws_raw);
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  if (tid < EA_WS_MATH * EA_WS_PROD) {
    ea_mbar_init(&R.full[tid / EA_WS_PROD][tid % EA_WS_PROD], 1);
    ea_mbar_init(&R.empty[tid / EA_WS_PROD][tid % EA_WS_PROD], 1);
  }
  if (tid == 0) {
    EaMsg m0;
    ea_boss_next_impl<true>(A, S, m0, true);
    S.msg[0] = m0;
  }
  __syncthreads();
  // Two separate code paths from here on (each with its own loop and the same sequence of CTA barriers): ptxas gives every
  // region the register budget of the setmaxnreg that dominates it, code shared by both would have to live in 64.
  if (warp < EA_WS_MATH) {
    asm volatile("setmaxnreg.inc.sync.aligned.u32 " EA_STR(EA_WS_MATH_REGS) ";");
    unsigned seq = 0;                                  // warp-iterations this math warp has consumed so far
    for (unsigned g = 0;; ++g) {
      const EaMsg& M = S.msg[g & 1];
      if (M.cmd == EA_CMD_EXIT) break;
      const EaLevelGeom& rg = A.ref_geom[M.level];
      const EaLevelGeom& ng = A.now_geom[M.level];
      EaPose P;
      if (M.pts_mode == EA_POINTS_XYZ) ea_pose_setup<true>(M.cand, rg, ng, P);
      else ea_pose_setup<false>(M.cand, rg, ng, P);
      ea_ws_consume(M.n_res, M.affine, A.sp, P, R, warp, lane, seq, S.part, S.cpart);
      asm volatile("bar.sync 0;" ::: "memory");
      if (warp == 0) {
        const double tot = ea_cta_total<EA_WS_MATH>(S.part, S.cpart, lane);
        if (lane < EA_SUMS) S.sums[lane] = tot;
        __syncwarp();
        if (lane == 0) {
          EaMsg mn;
          ea_boss_step_impl<true>(A, S, S.sums, M, mn);
          S.msg[(g + 1) & 1] = mn;
        }
      }
      asm volatile("bar.sync 0;" ::: "memory");
    }
  } else {
    asm volatile("setmaxnreg.dec.sync.aligned.u32 " EA_STR(EA_WS_GATHER_REGS) ";");
    const int mw = (warp - EA_WS_MATH) % EA_WS_MATH;   // the math warp this gather warp feeds
    const int pw = (warp - EA_WS_MATH) / EA_WS_MATH;   // its index among that math warp's five producers
    unsigned seq = 0;
    for (unsigned g = 0;; ++g) {
      const EaMsg& M = S.msg[g & 1];
      if (M.cmd == EA_CMD_EXIT) break;
      const EaLevelGeom& rg = A.ref_geom[M.level];
      const EaLevelGeom& ng = A.now_geom[M.level];
      EaPose P;
      if (M.pts_mode == EA_POINTS_XYZ) {
        ea_pose_setup<true>(M.cand, rg, ng, P);
        ea_ws_produce<true>(M.pts, M.dt, M.n_res, A.sp.point_stride, ng, A.inv_depth_scale, P, R, mw, pw, lane, seq);
      } else {
        ea_pose_setup<false>(M.cand, rg, ng, P);
        ea_ws_produce<false>(M.pts, M.dt, M.n_res, A.sp.point_stride, ng, A.inv_depth_scale, P, R, mw, pw, lane, seq);
      }
      asm volatile("bar.sync 0;" ::: "memory");
      asm volatile("bar.sync 0;" ::: "memory");         // the boss' LM step happens between the two
    }
  }
}

// ---- memory-side roof of the fused evaluation (measurement only) -----------------------------------------------------
// The same point stream, the same fp64 projection (it generates the addresses) and the same 16-texel clamp-to-edge gather
// as ea_eval_slice, and nothing else: no interpolation, no Jacobian, no reduction, no LM, few registers, full occupancy.
// Its rate is the L1/L2-gather roofline of the access pattern for the frames and poses at hand.
__global__ void __launch_bounds__(256) ea_k_gather_probe(const __grid_constant__ EaSolveArgs A, int level, int slices, int repeats,
                                                         float* __restrict__ sink) {
  const int pair = blockIdx.x / slices, slice = blockIdx.x % slices;
  if (pair >= A.n_pairs) return;
  const EaLevelDesc rd = A.ref_desc[size_t(A.ref_slots[pair]) * EA_MAX_LEVELS + level];
  const EaLevelDesc nd = A.now_desc[size_t(A.now_slots[pair]) * EA_MAX_LEVELS + level];
  const EaLevelGeom& rg = A.ref_geom[level];
  const EaLevelGeom& ng = A.now_geom[level];
  const int n = min(*rd.n_pts, A.ref_cap[level]);
  const int j0 = int((long long)n * slice / slices), j1 = int((long long)n * (slice + 1) / slices);
  if (rd.pts_mode != EA_POINTS_PIXEL) return;
  EaPose P;
  ea_pose_setup<false>(A.poses + size_t(pair) * 7, rg, ng, P);
  const float2 affine = make_float2(1.0f, 0.0f);
  float s = 0.0f;
  const int n_round = j0 + (((j1 - j0) + 31) & ~31);   // whole warps: the gather votes
  for (int r = 0; r < repeats; ++r)
    for (int j = j0 + int(threadIdx.x); j < n_round; j += 256) {
      typedef EaPtStream<false> PS;
      const PS::T p = j < j1 ? PS::load(rd.pts, size_t(j)) : PS::pad();
      double a0, a1, a2;
      PS::unpack(p, a0, a1, a2);
      s += ea_point_gather_sum<false>(a0, a1, a2, ng, A.inv_depth_scale, P, nd.dt);
    }
  (void)affine;
  sink[size_t(blockIdx.x) * 256 + threadIdx.x] = s;
}

cudaError_t ea_launch_gather_probe(const EaSolveArgs& A, int level, int slices, int repeats, float* d_sink, cudaStream_t stream) {
  // EA_PROBE_SMEM_KB: occupancy experiment (dynamic shared memory only limits the number of resident CTAs)
  static const int smem_kb = getenv("EA_PROBE_SMEM_KB") ? atoi(getenv("EA_PROBE_SMEM_KB")) : 0;
  if (smem_kb > 48) cudaFuncSetAttribute(ea_k_gather_probe, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_kb * 1024);
  ea_k_gather_probe<<<unsigned(A.n_pairs * slices), 256, size_t(smem_kb) * 1024, stream>>>(A, level, slices, repeats, d_sink);
  return cudaGetLastError();
}

// ---- host launchers -----------------------------------------------------------------------------------
cudaError_t ea_launch_solve_batch(const EaSolveArgs& A, int cluster_size, int sm_count, cudaStream_t stream) {
  if (A.n_pairs <= 0) return cudaSuccess;
  cudaError_t e = cudaMemsetAsync(A.work_counter, 0, sizeof(int), stream);
  if (e != cudaSuccess) return e;
  cudaLaunchConfig_t cfg = {};
  cfg.blockDim = dim3(EA_SOLVE_THREADS);
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  if (cluster_size == EA_KERNEL_WS) {
    static bool ws_attr = false;
    if (!ws_attr) {
      e = cudaFuncSetAttribute(ea_k_solve_ws, cudaFuncAttributeMaxDynamicSharedMemorySize, int(sizeof(EaWsRing)));
      if (e != cudaSuccess) return e;
      ws_attr = true;
    }
    cfg.blockDim = dim3(EA_WS_THREADS);
    cfg.dynamicSmemBytes = sizeof(EaWsRing);
    cfg.gridDim = dim3(unsigned(A.n_pairs < sm_count ? A.n_pairs : sm_count));
    cfg.numAttrs = 0;
    return cudaLaunchKernelEx(&cfg, ea_k_solve_ws, A);
  }
  if (cluster_size <= 1) {
    // persistent CTAs (EA_SOLVE_MIN_CTAS per SM) pulling pairs from the work queue
    int max_ctas = sm_count * EA_SOLVE_MIN_CTAS;
    static const int env_ctas = getenv("EA_SOLVE_MAX_CTAS") ? atoi(getenv("EA_SOLVE_MAX_CTAS")) : 0;   // experiment knob
    if (env_ctas > 0 && env_ctas < max_ctas) max_ctas = env_ctas;
    cfg.gridDim = dim3(unsigned(A.n_pairs < max_ctas ? A.n_pairs : max_ctas));
    cfg.numAttrs = 0;
    return cudaLaunchKernelEx(&cfg, ea_k_solve_batch<EA_SOLVE_THREADS, false>, A);
  }
  int n_clusters = A.n_pairs;
  const int max_clusters = (sm_count * EA_SOLVE_MIN_CTAS) / cluster_size;
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
                                const double* d_pose7, const int* d_done, int j_begin, int j_end, int n_blocks, double* d_sums,
                                cudaStream_t stream) {
  if (n_blocks <= 0) return cudaSuccess;
  ea_k_eval_sums<EA_SOLVE_THREADS><<<n_blocks, EA_SOLVE_THREADS, 0, stream>>>(rd, nd, rg, ng, inv_depth_scale, sp, d_pose7, d_done, j_begin, j_end, d_sums);
  return cudaGetLastError();
}

// ---- longest-first ordering for the next launch: sort pairs by the work their last solve needed -----------------
// One CTA, bitonic sort of (work, index) in shared memory; n <= 4096.
__global__ void __launch_bounds__(1024) ea_k_order_by_work(const ea_summary* __restrict__ summaries, int n, int n_levels, int32_t* __restrict__ order) {
  __shared__ unsigned long long key[4096];
  int m = 1;
  while (m < n) m <<= 1;
  for (int i = threadIdx.x; i < m; i += blockDim.x) {
    unsigned long long k = 0;
    if (i < n) {
      unsigned long long work = 0;
      for (int l = 0; l < n_levels; ++l) {
        const ea_summary s = summaries[size_t(i) * n_levels + l];
        work += (unsigned long long)(s.n_residuals > 0 ? s.n_residuals : 0) * (unsigned long long)(s.evaluations > 0 ? s.evaluations : 0);
      }
      if (work > 0xFFFFFFFFFFFull) work = 0xFFFFFFFFFFFull;
      k = (work << 20) | (unsigned long long)(0xFFFFF - i);       // descending work, ascending index on ties
    }
    key[i] = k;
  }
  __syncthreads();
  for (int size = 2; size <= m; size <<= 1)
    for (int stride = size >> 1; stride > 0; stride >>= 1) {
      for (int i = threadIdx.x; i < m; i += blockDim.x) {
        const int j = i ^ stride;
        if (j > i) {
          const bool desc = (i & size) == 0;
          const unsigned long long a = key[i], b = key[j];
          if (desc ? (a < b) : (a > b)) { key[i] = b; key[j] = a; }
        }
      }
      __syncthreads();
    }
  for (int i = threadIdx.x; i < n; i += blockDim.x) order[i] = int32_t(0xFFFFF - int(key[i] & 0xFFFFF));
}

cudaError_t ea_launch_order_by_work(const ea_summary* d_summaries, int n, int n_levels, int32_t* d_order, cudaStream_t stream) {
  if (n <= 0 || n > 4096) return cudaErrorInvalidValue;
  ea_k_order_by_work<<<1, 1024, 0, stream>>>(d_summaries, n, n_levels, d_order);
  return cudaGetLastError();
}
