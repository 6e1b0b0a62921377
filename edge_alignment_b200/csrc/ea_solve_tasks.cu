// Task-graph variant of the batched on-device solve (the default for batches).
//
// Unit of work = one CHUNK of one evaluation of one pair: (pair, chunk) tasks flow through a device-side MPMC ring
// queue.  Any CTA evaluates any chunk and writes its 29 partial sums into the pair's fixed slot; the CTA that
// finishes an evaluation's last chunk adds the slots in fixed order (deterministic, independent of scheduling and
// of the SM count), runs the 6x6 LM step for that pair and enqueues the next evaluation's chunks -- or moves the pair
// to the next pyramid level, or retires it and admits the next pair of the batch.
//
// Why (measured with the one-CTA-per-pair kernel in ea_solve.cu, profiles/): 13 % of SM time sat behind the serial
// LM step, and the launch tail was set by the slowest pair (50 LM iterations at level 0 ~ 5 ms on one SM).  Here an
// LM step only stalls one small CTA while the SM's other CTAs keep evaluating, and a straggler's evaluations are
// spread over the whole GPU.  A bounded admission window keeps the distance transforms of the active pairs L2-sized.
//
// Replaces the same reference code as ea_solve.cu: standalone_edge_align.cpp:265-291.
#include "ea_internal.h"
#include "ea_solve.cuh"

#ifndef EA_TASK_THREADS
#define EA_TASK_THREADS 256
#endif
#ifndef EA_TASK_MIN_CTAS
#define EA_TASK_MIN_CTAS 2
#endif
#define EA_TASK_EXIT 0xFFFFFFFFu

namespace {

__device__ __forceinline__ unsigned long long ld_acquire_u64(const unsigned long long* p) {
  unsigned long long v;
  asm volatile("ld.acquire.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_release_u64(unsigned long long* p, unsigned long long v) {
  asm volatile("st.release.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}

// ---- producer side (one thread) --------------------------------------------------------------------------------
__device__ void push_tasks(const EaSolveArgs& A, int pair, int n_chunks) {
  __threadfence();   // the pair state written above must be visible before its tasks are
  const unsigned base = atomicAdd(&A.queue->tail, unsigned(n_chunks));
  for (int c = 0; c < n_chunks; ++c) {
    const unsigned t = base + unsigned(c);
    st_release_u64(A.slots + (t & (EA_QUEUE_CAP - 1)), (static_cast<unsigned long long>(t + 1u) << 32) | unsigned(pair * EA_MAX_CHUNKS + c));
  }
}
__device__ void push_exit(const EaSolveArgs& A, int n) {
  const unsigned base = atomicAdd(&A.queue->tail, unsigned(n));
  for (int c = 0; c < n; ++c) {
    const unsigned t = base + unsigned(c);
    st_release_u64(A.slots + (t & (EA_QUEUE_CAP - 1)), (static_cast<unsigned long long>(t + 1u) << 32) | EA_TASK_EXIT);
  }
}

// Start the first evaluation of the highest remaining level of `pair` that has points.  Returns false when no level is
// left (the pair is finished).  L is the caller's working copy of the LM state (x holds the current pose).
__device__ bool start_level(const EaSolveArgs& A, int pair, EaPairState* st, EaLmState& L, int level) {
  for (; level >= A.finest; --level) {
    const EaLevelDesc rd = A.ref_desc[size_t(A.ref_slots[pair]) * EA_MAX_LEVELS + level];
    const EaLevelDesc nd = A.now_desc[size_t(A.now_slots[pair]) * EA_MAX_LEVELS + level];
    const int n_pts = min(__ldcg(rd.n_pts), A.ref_cap[level]);
    const int n_res = (n_pts + A.sp.point_stride - 1) / A.sp.point_stride;
    if (n_res == 0) {
      if (A.summaries) { ea_summary z = {}; z.termination = EA_TERM_SKIPPED_NO_POINTS; A.summaries[size_t(pair) * A.n_levels + level] = z; }
      continue;
    }
    L.phase = 0; L.iter = 0; L.accepted = 0; L.rejected = 0; L.invalid_run = 0; L.evals = 0; L.term = EA_TERM_NONE;
    int n_chunks = (n_res + A.chunk_points - 1) / A.chunk_points;
    n_chunks = max(1, min(n_chunks, EA_MAX_CHUNKS));
#pragma unroll 1
    for (int i = 0; i < 7; ++i) { L.cand[i] = L.x[i]; st->cand[i] = L.x[i]; }
    st->pts = rd.pts; st->dt = nd.dt; st->affine = __ldcg(nd.dt_affine); st->n_res = n_res; st->level = level;
    st->pts_mode = rd.pts_mode; st->n_chunks = n_chunks; st->remaining = n_chunks;
    st->lm = L;
    push_tasks(A, pair, n_chunks);
    return true;
  }
  return false;
}

// Admit pairs from the batch until one actually has work (or the batch is exhausted); retire finished ones.
__device__ void admit_next(const EaSolveArgs& A, EaLmState& L) {
  for (;;) {
    const int pair = atomicAdd(&A.queue->next_pair, 1);
    if (pair >= A.n_pairs) return;
    const int pi = A.pose_index ? A.pose_index[pair] : pair;
#pragma unroll 1
    for (int i = 0; i < 7; ++i) L.x[i] = __ldcg(A.poses + size_t(pi) * 7 + i);
    L.cost = 0.0; L.initial_cost = 0.0;
    if (start_level(A, pair, A.states + pair, L, A.coarsest)) return;
    // nothing to solve at any level: pose stays as it is
    const int done = atomicAdd(&A.queue->done_pairs, 1) + 1;
    if (done == A.n_pairs) { push_exit(A, int(gridDim.x)); return; }
  }
}

// The last chunk of an evaluation has landed: advance this pair.  `sums` = the evaluation's 29 totals.
__device__ __noinline__ void advance_pair(const EaSolveArgs& A, int pair, const double* sums, EaLmState& L) {
  EaPairState* st = A.states + pair;
  // working copy of the LM state (it was last written by another SM: bypass L1)
  {
    const double* src = reinterpret_cast<const double*>(&st->lm);
    double* dst = reinterpret_cast<double*>(&L);
#pragma unroll 1
    for (int i = 0; i < int(sizeof(EaLmState) / 8); ++i) dst[i] = __ldcg(src + i);
  }
  const int level = __ldcg(&st->level);
  const int cmd = ea_lm_advance(L, sums, A.sp);
  if (cmd == EA_CMD_EVAL) {
    const int n_chunks = __ldcg(&st->n_chunks);
#pragma unroll 1
    for (int i = 0; i < 7; ++i) st->cand[i] = L.cand[i];
    st->lm = L;
    st->remaining = n_chunks;
    push_tasks(A, pair, n_chunks);
    return;
  }
  if (A.summaries) {
    ea_summary z;
    z.termination = L.term; z.iterations = L.iter; z.accepted = L.accepted; z.rejected = L.rejected;
    z.n_residuals = __ldcg(&st->n_res); z.evaluations = L.evals; z.initial_cost = L.initial_cost; z.final_cost = L.cost;
    A.summaries[size_t(pair) * A.n_levels + level] = z;
  }
  if (start_level(A, pair, st, L, level - 1)) return;
  // pair finished
  const int pi = A.pose_index ? A.pose_index[pair] : pair;
#pragma unroll 1
  for (int i = 0; i < 7; ++i) A.poses[size_t(pi) * 7 + i] = L.x[i];
  const int done = atomicAdd(&A.queue->done_pairs, 1) + 1;
  if (done == A.n_pairs) { push_exit(A, int(gridDim.x)); return; }
  admit_next(A, L);
}

struct TaskSmem {
  double part[EA_TASK_THREADS / 32][EA_NSUM];
  double cpart[EA_TASK_THREADS / 32];
  double sums[EA_SUMS + 3];
  double cand[7];
  EaLmState lm;          // scratch for the thread that advances a pair
  unsigned payload;
  int last;
};

__global__ void __launch_bounds__(EA_TASK_THREADS, EA_TASK_MIN_CTAS) ea_k_solve_tasks(const __grid_constant__ EaSolveArgs A) {
  __shared__ TaskSmem S;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  if (tid == 0) {   // initial admission: `window` pairs spread over the CTAs
    for (int k = int(blockIdx.x); k < A.window; k += int(gridDim.x)) admit_next(A, S.lm);
  }
  for (;;) {
    if (tid == 0) {
      const unsigned ticket = atomicAdd(&A.queue->head, 1u);
      const unsigned long long* slot = A.slots + (ticket & (EA_QUEUE_CAP - 1));
      unsigned long long v;
      unsigned ns = 32;
      while (((v = ld_acquire_u64(slot)) >> 32) != static_cast<unsigned long long>(ticket + 1u)) {
        __nanosleep(ns);
        if (ns < 1024) ns <<= 1;
      }
      S.payload = unsigned(v);
    }
    __syncthreads();
    const unsigned payload = S.payload;
    if (payload == EA_TASK_EXIT) break;
    const int pair = int(payload / EA_MAX_CHUNKS), chunk = int(payload % EA_MAX_CHUNKS);
    EaPairState* st = A.states + pair;
    // pair state was written by another SM and changes every evaluation: read it around L1
    if (tid < 7) S.cand[tid] = __ldcg(&st->cand[tid]);
    const void* pts = reinterpret_cast<const void*>(__ldcg(reinterpret_cast<const unsigned long long*>(&st->pts)));
    const float* dt = reinterpret_cast<const float*>(__ldcg(reinterpret_cast<const unsigned long long*>(&st->dt)));
    const float2 affine = __ldcg(&st->affine);
    const int n_res = __ldcg(&st->n_res), level = __ldcg(&st->level), pts_mode = __ldcg(&st->pts_mode), n_chunks = __ldcg(&st->n_chunks);
    __syncthreads();
    const EaLevelGeom& rg = A.ref_geom[level];
    const EaLevelGeom& ng = A.now_geom[level];
    const int j0 = int((long long)n_res * chunk / n_chunks), j1 = int((long long)n_res * (chunk + 1) / n_chunks);
    EaPose P;
    if (pts_mode == EA_POINTS_XYZ) {
      ea_pose_setup<true>(S.cand, rg, ng, P);
      ea_eval_slice<true, EA_TASK_THREADS>(pts, dt, affine, ng, A.inv_depth_scale, A.sp, P, j0, j1, S.part, S.cpart);
    } else {
      ea_pose_setup<false>(S.cand, rg, ng, P);
      ea_eval_slice<false, EA_TASK_THREADS>(pts, dt, affine, ng, A.inv_depth_scale, A.sp, P, j0, j1, S.part, S.cpart);
    }
    __syncthreads();
    if (warp == 0) {
      const double tot = ea_cta_total<EA_TASK_THREADS / 32>(S.part, S.cpart, lane);
      if (lane < EA_SUMS) { st->partial[chunk][lane] = tot; __threadfence(); }   // every writer fences its own store
      __syncwarp();
      int last = 0;
      if (lane == 0) {
        __threadfence();                                   // partials before the counter
        last = (atomicSub(&st->remaining, 1) == 1);
        if (last) __threadfence();                         // counter before the other chunks' partials
      }
      last = __shfl_sync(0xffffffffu, last, 0);
      if (last) {
        if (lane < EA_SUMS) {
          double s = 0.0;
          for (int c = 0; c < n_chunks; ++c) s += __ldcg(&st->partial[c][lane]);   // fixed order
          S.sums[lane] = s;
        }
        __syncwarp();
        if (lane == 0) advance_pair(A, pair, S.sums, S.lm);
      }
    }
    // (the next iteration's first __syncthreads orders this iteration's shared-memory reads before new writes)
  }
}

}  // namespace

cudaError_t ea_launch_solve_tasks(const EaSolveArgs& A, int sm_count, cudaStream_t stream) {
  if (A.n_pairs <= 0) return cudaSuccess;
  cudaError_t e = cudaMemsetAsync(A.queue, 0, sizeof(EaQueue), stream);
  if (e != cudaSuccess) return e;
  e = cudaMemsetAsync(A.slots, 0, size_t(EA_QUEUE_CAP) * sizeof(unsigned long long), stream);
  if (e != cudaSuccess) return e;
  int per_sm = 0;
  e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, ea_k_solve_tasks, EA_TASK_THREADS, 0);
  if (e != cudaSuccess) return e;
  if (per_sm < 1) return cudaErrorLaunchOutOfResources;
  // persistent grid: never more CTAs than can be co-resident (consumers wait on producers)
  int grid = sm_count * per_sm;
  EaSolveArgs B = A;
  if (B.window <= 0) B.window = grid / 2;
  if (B.window > A.n_pairs) B.window = A.n_pairs;
  if (B.window < 1) B.window = 1;
  ea_k_solve_tasks<<<grid, EA_TASK_THREADS, 0, stream>>>(B);
  return cudaGetLastError();
}
