// Point-sharded solve of ONE large frame pair across ranks (BASELINE config "4K, ~2M edge points"):
// every rank holds the full distance transform and a contiguous slice of the ordered point list, reduces its
// slice to the 29 normal-equation sums, and one ncclAllReduce(sum, 29 x fp64 = 232 B) per evaluation makes
// them global.  All ranks then run the identical 6x6 LM step on bit-identical sums, so they take identical
// accept/reject decisions with no broadcast.  Nothing in the reference corresponds to this (it is
// single-threaded); the math per point and the LM semantics are those of ea_solve.cu / ea_solve.cuh.
//
// NCCL is bound lazily with dlopen("libnccl.so.2") so the library has no hard NCCL dependency and, inside a
// torch process, resolves to the copy torch already loaded.
#include <dlfcn.h>
#include <nccl.h>

#include <algorithm>
#include <cstring>
#include <new>
#include <vector>

#include "ea_internal.h"
#include "ea_solve.cuh"

namespace {

struct NcclApi {
  void* handle = nullptr;
  ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
  ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
  ncclResult_t (*AllReduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
  const char* (*GetErrorString)(ncclResult_t) = nullptr;
};
NcclApi g_nccl;

int load_nccl() {
  if (g_nccl.handle) return EA_OK;
  void* h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
  if (!h) h = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
  if (!h) return ea_fail(EA_ERR_NCCL, "cannot dlopen libnccl.so.2: %s", dlerror());
  g_nccl.GetUniqueId = (decltype(g_nccl.GetUniqueId))dlsym(h, "ncclGetUniqueId");
  g_nccl.CommInitRank = (decltype(g_nccl.CommInitRank))dlsym(h, "ncclCommInitRank");
  g_nccl.AllReduce = (decltype(g_nccl.AllReduce))dlsym(h, "ncclAllReduce");
  g_nccl.CommDestroy = (decltype(g_nccl.CommDestroy))dlsym(h, "ncclCommDestroy");
  g_nccl.GetErrorString = (decltype(g_nccl.GetErrorString))dlsym(h, "ncclGetErrorString");
  if (!g_nccl.GetUniqueId || !g_nccl.CommInitRank || !g_nccl.AllReduce || !g_nccl.CommDestroy || !g_nccl.GetErrorString)
    return ea_fail(EA_ERR_NCCL, "libnccl.so.2 lacks a required symbol");
  g_nccl.handle = h;
  return EA_OK;
}
#define NC(call)                                                                                               \
  do {                                                                                                         \
    ncclResult_t r_ = (call);                                                                                  \
    if (r_ != ncclSuccess) return ea_fail(EA_ERR_NCCL, "%s -> %s", #call, g_nccl.GetErrorString(r_));          \
  } while (0)

struct ShardState {        // device-resident control block
  EaLmState lm;
  double cand[7];
  int done, started;
};

// ---- general (rig / distortion) evaluation kernels: residual variants of standalone/utils.h:101-421 ----------------
#define GV_THREADS 256
template <bool XYZ>
__global__ void __launch_bounds__(GV_THREADS) k_view_eval_sums(EaLevelDesc rd, EaLevelDesc nd, EaLevelGeom rg, EaLevelGeom ng, EaViewXf V,
                                                               double inv_depth_scale, ea_solve_params sp, const double* pose7,
                                                               const int* done, int j_begin, int j_end, double* sums /*[grid][EA_SUMS]*/) {
  if (done && *done) return;
  __shared__ double part[GV_THREADS / 32][EA_NSUM];
  __shared__ double cpart[GV_THREADS / 32];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int n = j_end - j_begin;
  const int j0 = j_begin + int((long long)n * blockIdx.x / gridDim.x), j1 = j_begin + int((long long)n * (blockIdx.x + 1) / gridDim.x);
  EaPoseG P;
  ea_pose_setup_general<XYZ>(pose7, rg, ng, V, P);
  const float2 affine = *nd.dt_affine;
  float acc[EA_NSUM];
#pragma unroll
  for (int k = 0; k < EA_NSUM; ++k) acc[k] = 0.0f;
  double acc64 = 0.0, cost64 = 0.0;
  for (int base = j0 + warp * 32; base < j1; base += GV_THREADS) {
    const int j = base + lane;
    if (j < j1) {
      const float4 p = EaPtStream<XYZ>::as_float4(EaPtStream<XYZ>::load(rd.pts, size_t(j) * sp.point_stride));
      float f, w, rho0, J[6];
      const bool fail = ea_point_eval_general<XYZ>(p, ng, inv_depth_scale, P, nd.dt, affine, sp.loss_type, float(sp.loss_scale), f, w, rho0, J);
      ea_accumulate(acc, J, f * w);
      acc[27] += fail ? 1.0f : 0.0f;
      cost64 += double(0.5f * rho0);
    }
    acc64 += double(ea_warp_transpose_reduce(acc, lane));
#pragma unroll
    for (int k = 0; k < EA_NSUM; ++k) acc[k] = 0.0f;
  }
  cost64 = ea_warp_sum(cost64);
  part[warp][lane] = acc64;
  if (lane == 0) cpart[warp] = cost64;
  __syncthreads();
  if (tid < 32) {
    const double tot = ea_cta_total<GV_THREADS / 32>(part, cpart, tid);
    if (tid < EA_SUMS) sums[size_t(blockIdx.x) * EA_SUMS + tid] = tot;
  }
}

template <bool XYZ>
__global__ void __launch_bounds__(256) k_view_eval_points(EaLevelDesc rd, EaLevelDesc nd, EaLevelGeom rg, EaLevelGeom ng, EaViewXf V,
                                                          double inv_depth_scale, ea_solve_params sp, const double* pose7, int n_res,
                                                          double* raw, double* res, double* jac, int* failed) {
  EaPoseG P;
  ea_pose_setup_general<XYZ>(pose7, rg, ng, V, P);
  const float2 affine = *nd.dt_affine;
  for (int j = blockIdx.x * blockDim.x + threadIdx.x; j < n_res; j += gridDim.x * blockDim.x) {
    const float4 p = EaPtStream<XYZ>::as_float4(EaPtStream<XYZ>::load(rd.pts, size_t(j) * sp.point_stride));
    float f, w, rho0, J[6];
    const bool fail = ea_point_eval_general<XYZ>(p, ng, inv_depth_scale, P, nd.dt, affine, sp.loss_type, float(sp.loss_scale), f, w, rho0, J);
    if (raw) raw[j] = double(f);
    if (res) res[j] = double(f * w);
    if (jac) for (int k = 0; k < 6; ++k) jac[size_t(j) * 6 + k] = double(J[k]);
    if (fail && failed) atomicAdd(failed, 1);
  }
}

__global__ void k_shard_reduce(const ShardState* st, const double* partials, int n_blocks, double* sums) {
  if (st->done) return;
  const int k = threadIdx.x;
  if (k < EA_SUMS) {
    double s = 0.0;
    for (int b = 0; b < n_blocks; ++b) s += partials[size_t(b) * EA_SUMS + k];   // fixed order: deterministic
    sums[k] = s;
  }
}

__global__ void k_shard_lm(ShardState* st, const double* sums, ea_solve_params sp) {
  if (st->done) return;
  double local[EA_SUMS];
  for (int k = 0; k < EA_SUMS; ++k) local[k] = sums[k];
  const int cmd = ea_lm_advance(st->lm, local, sp);
  if (cmd == EA_CMD_EVAL) { for (int i = 0; i < 7; ++i) st->cand[i] = st->lm.cand[i]; }
  else st->done = 1;
}

__global__ void k_shard_init(ShardState* st, const double* pose7) {
  EaLmState& L = st->lm;
  L.phase = 0; L.iter = 0; L.accepted = 0; L.rejected = 0; L.invalid_run = 0; L.evals = 0; L.term = EA_TERM_NONE;
  L.cost = 0.0; L.initial_cost = 0.0;
  for (int i = 0; i < 7; ++i) { L.x[i] = pose7[i]; L.cand[i] = pose7[i]; st->cand[i] = pose7[i]; }
  st->done = 0; st->started = 1;
}

}  // namespace

struct ea_shard {
  ea_context* ctx = nullptr;
  ncclComm_t comm = nullptr;
  int rank = 0, world = 1;
  ShardState* d_state = nullptr;
  double* d_partials = nullptr;
  double* d_sums = nullptr;
  double* d_pose = nullptr;
  int n_blocks = 0;
  int* h_done = nullptr;  // pinned
};

extern "C" {

int ea_shard_unique_id(uint8_t id128[128]) {
  if (!id128) return ea_fail(EA_ERR_INVALID_ARG, "null id");
  int rc = load_nccl();
  if (rc) return rc;
  ncclUniqueId id;
  NC(g_nccl.GetUniqueId(&id));
  static_assert(sizeof(id) == 128, "ncclUniqueId size");
  std::memcpy(id128, &id, 128);
  return EA_OK;
}

int ea_shard_create(ea_context* ctx, const uint8_t id128[128], int rank, int world, ea_shard** out) {
  if (!ctx || !out) return ea_fail(EA_ERR_INVALID_ARG, "null argument");
  *out = nullptr;
  if (world < 1 || rank < 0 || rank >= world) return ea_fail(EA_ERR_INVALID_ARG, "bad rank/world %d/%d", rank, world);
  CU(cudaSetDevice(ctx->device));
  ea_shard* s = new (std::nothrow) ea_shard();
  if (!s) return ea_fail(EA_ERR_INVALID_ARG, "out of host memory");
  s->ctx = ctx; s->rank = rank; s->world = world;
  if (world > 1) {
    if (!id128) { delete s; return ea_fail(EA_ERR_INVALID_ARG, "world > 1 needs the NCCL unique id"); }
    int rc = load_nccl();
    if (rc) { delete s; return rc; }
    ncclUniqueId id;
    std::memcpy(&id, id128, 128);
    ncclResult_t r = g_nccl.CommInitRank(&s->comm, world, id, rank);
    if (r != ncclSuccess) { delete s; return ea_fail(EA_ERR_NCCL, "ncclCommInitRank -> %s", g_nccl.GetErrorString(r)); }
  }
  s->n_blocks = ctx->sm_count * 2;
  CU(cudaMalloc((void**)&s->d_state, sizeof(ShardState)));
  CU(cudaMalloc((void**)&s->d_partials, size_t(s->n_blocks) * EA_SUMS * 8));
  CU(cudaMalloc((void**)&s->d_sums, EA_SUMS * 8));
  CU(cudaMalloc((void**)&s->d_pose, 7 * 8));
  CU(cudaHostAlloc((void**)&s->h_done, sizeof(int), cudaHostAllocDefault));
  *out = s;
  return EA_OK;
}

int ea_shard_destroy(ea_shard* s) {
  if (!s) return EA_OK;
  cudaSetDevice(s->ctx->device);
  cudaStreamSynchronize(s->ctx->stream);
  if (s->comm) g_nccl.CommDestroy(s->comm);
  cudaFree(s->d_state); cudaFree(s->d_partials); cudaFree(s->d_sums); cudaFree(s->d_pose);
  if (s->h_done) cudaFreeHost(s->h_done);
  delete s;
  return EA_OK;
}

int ea_shard_solve(ea_shard* s, ea_frameset* ref, int ref_slot, ea_frameset* now, int now_slot, int level, double* pose7,
                   const ea_solve_params* sp, ea_summary* summary) {
  if (!s || !ref || !now || !pose7 || !sp) return ea_fail(EA_ERR_INVALID_ARG, "null argument");
  if (level < 0 || level >= ref->p.n_levels || level >= now->p.n_levels) return ea_fail(EA_ERR_INVALID_ARG, "bad level");
  if (ref_slot < 0 || ref_slot >= ref->n_slots || now_slot < 0 || now_slot >= now->n_slots) return ea_fail(EA_ERR_INVALID_ARG, "bad slot");
  if (sp->point_stride < 1) return ea_fail(EA_ERR_INVALID_ARG, "point_stride must be >= 1");
  ea_context* c = s->ctx;
  CU(cudaSetDevice(c->device));
  cudaStream_t st = c->stream;
  int n_pts = 0;
  int rc = ea_frameset_get_num_points(ref, ref_slot, level, &n_pts);
  if (rc) return rc;
  const int n_res = (n_pts + sp->point_stride - 1) / sp->point_stride;
  // this rank's contiguous slice of the ordered residual list
  const int j0 = int((long long)n_res * s->rank / s->world), j1 = int((long long)n_res * (s->rank + 1) / s->world);
  const EaLevelDesc& rd = ref->h_desc[size_t(ref_slot) * EA_MAX_LEVELS + level];
  const EaLevelDesc& nd = now->h_desc[size_t(now_slot) * EA_MAX_LEVELS + level];
  const double ids = ref->inv_depth_unit;
  int nb = s->n_blocks;
  const int per_block = 2048;
  if ((j1 - j0 + per_block - 1) / per_block < nb) nb = std::max(1, (j1 - j0 + per_block - 1) / per_block);
  CU(cudaMemcpyAsync(s->d_pose, pose7, 56, cudaMemcpyHostToDevice, st));
  k_shard_init<<<1, 1, 0, st>>>(s->d_state, s->d_pose);
  c->launches++;
  const int max_evals = sp->max_num_iterations + 2;
  int evals = 0;
  *s->h_done = 0;
  while (evals < max_evals && !*s->h_done) {
    const int chunk = std::min(8, max_evals - evals);
    for (int i = 0; i < chunk; ++i) {
      cudaError_t le = ea_launch_eval_sums(rd, nd, ref->geom[level], now->geom[level], ids, *sp, s->d_state->cand, &s->d_state->done, j0, j1, nb, s->d_partials, st);
      if (le != cudaSuccess) return ea_fail(EA_ERR_CUDA, "shard eval launch: %s", cudaGetErrorString(le));
      k_shard_reduce<<<1, 32, 0, st>>>(s->d_state, s->d_partials, nb, s->d_sums);
      if (s->world > 1) NC(g_nccl.AllReduce(s->d_sums, s->d_sums, EA_SUMS, ncclDouble, ncclSum, s->comm, st));
      k_shard_lm<<<1, 1, 0, st>>>(s->d_state, s->d_sums, *sp);
      c->launches += 3;
    }
    evals += chunk;
    CU(cudaMemcpyAsync(s->h_done, &s->d_state->done, sizeof(int), cudaMemcpyDeviceToHost, st));
    CU(cudaStreamSynchronize(st));
  }
  ShardState h;
  CU(cudaMemcpy(&h, s->d_state, sizeof h, cudaMemcpyDeviceToHost));
  for (int i = 0; i < 7; ++i) pose7[i] = h.lm.x[i];
  if (summary) {
    summary->termination = h.done ? h.lm.term : EA_TERM_NO_CONVERGENCE;
    summary->iterations = h.lm.iter; summary->accepted = h.lm.accepted; summary->rejected = h.lm.rejected;
    summary->n_residuals = n_res; summary->evaluations = h.lm.evals;
    summary->initial_cost = h.lm.initial_cost; summary->final_cost = h.lm.cost;
  }
  return EA_OK;
}


// ---- multi-camera problems: every view adds its residual blocks to the same 6-DoF problem ----------------------------
// (standalone_edge_align.cpp:791-803 / 3204-3219: one AddResidualBlock loop per camera with EAResidue /
//  EAResidueSecondCam / EAResidueEx / EAResidueSecondCamEx, all on the same b_quat_a, b_t_a)
static int view_xf(const ea_view& v, EaViewXf& X) {
  std::memset(&X, 0, sizeof X);
  X.use_rig = v.use_rig != 0; X.use_dist = v.use_distortion != 0;
  for (int i = 0; i < 12; ++i) { X.T21[i] = v.cam_T_first[i]; X.T12[i] = v.first_T_cam[i]; }
  if (!X.use_rig) { for (int i = 0; i < 12; ++i) X.T21[i] = X.T12[i] = (i % 5 == 0) ? 1.0 : 0.0; }
  for (int i = 0; i < 5; ++i) X.dist[i] = X.use_dist ? v.dist[i] : 0.0;
  return EA_OK;
}
static int check_view(const ea_view& v, int level) {
  if (!v.ref || !v.now) return ea_fail(EA_ERR_INVALID_ARG, "view with a null frameset");
  if (v.ref->ctx->device != v.now->ctx->device) return ea_fail(EA_ERR_INVALID_ARG, "view framesets on different devices");
  if (level < 0 || level >= v.ref->p.n_levels || level >= v.now->p.n_levels) return ea_fail(EA_ERR_INVALID_ARG, "bad level");
  if (v.ref_slot < 0 || v.ref_slot >= v.ref->n_slots || v.now_slot < 0 || v.now_slot >= v.now->n_slots) return ea_fail(EA_ERR_INVALID_ARG, "bad slot");
  return EA_OK;
}

int ea_eval_views(ea_context* c, int n_views, const ea_view* views, int level, const double* pose7, const ea_solve_params* sp,
                  int* n_residuals, double* raw, double* residuals, double* jac, double* sums28, int* failed) {
  if (!c || !views || n_views < 1 || !pose7 || !sp) return ea_fail(EA_ERR_INVALID_ARG, "null argument");
  if (sp->point_stride < 1) return ea_fail(EA_ERR_INVALID_ARG, "point_stride must be >= 1");
  CU(cudaSetDevice(c->device));
  cudaStream_t st = c->stream;
  std::vector<int> nres(n_views);
  int total = 0;
  for (int i = 0; i < n_views; ++i) {
    int rc = check_view(views[i], level);
    if (rc) return rc;
    int n_pts = 0;
    rc = ea_frameset_get_num_points(views[i].ref, views[i].ref_slot, level, &n_pts);
    if (rc) return rc;
    nres[i] = (n_pts + sp->point_stride - 1) / sp->point_stride;
    total += nres[i];
  }
  if (n_residuals) *n_residuals = total;
  CU(cudaMemcpyAsync(c->d_pose, pose7, 56, cudaMemcpyHostToDevice, st));
  CU(cudaMemsetAsync(c->d_failed, 0, sizeof(int), st));
  int rc = ea_ensure_tmp(c, size_t(std::max(total, 1)) * 8 * sizeof(double) + size_t(n_views) * 64 * EA_SUMS * 8);
  if (rc) return rc;
  double* d_raw = (double*)c->d_tmp; double* d_res = d_raw + total; double* d_jac = d_res + total; double* d_part = d_jac + size_t(total) * 6;
  int off = 0, nb_total = 0;
  for (int i = 0; i < n_views; ++i) {
    const ea_view& v = views[i];
    const EaLevelDesc& rd = v.ref->h_desc[size_t(v.ref_slot) * EA_MAX_LEVELS + level];
    const EaLevelDesc& nd = v.now->h_desc[size_t(v.now_slot) * EA_MAX_LEVELS + level];
    EaViewXf X; view_xf(v, X);
    const double ids = v.ref->inv_depth_unit;
    if (nres[i] > 0) {
      const int blocks = (nres[i] + 255) / 256;
      if (rd.pts_mode == EA_POINTS_XYZ)
        k_view_eval_points<true><<<blocks, 256, 0, st>>>(rd, nd, v.ref->geom[level], v.now->geom[level], X, ids, *sp, c->d_pose, nres[i], d_raw + off, d_res + off, d_jac + size_t(off) * 6, c->d_failed);
      else
        k_view_eval_points<false><<<blocks, 256, 0, st>>>(rd, nd, v.ref->geom[level], v.now->geom[level], X, ids, *sp, c->d_pose, nres[i], d_raw + off, d_res + off, d_jac + size_t(off) * 6, c->d_failed);
      const int nb = std::max(1, std::min(64, (nres[i] + 2047) / 2048));
      if (rd.pts_mode == EA_POINTS_XYZ)
        k_view_eval_sums<true><<<nb, GV_THREADS, 0, st>>>(rd, nd, v.ref->geom[level], v.now->geom[level], X, ids, *sp, c->d_pose, nullptr, 0, nres[i], d_part + size_t(nb_total) * EA_SUMS);
      else
        k_view_eval_sums<false><<<nb, GV_THREADS, 0, st>>>(rd, nd, v.ref->geom[level], v.now->geom[level], X, ids, *sp, c->d_pose, nullptr, 0, nres[i], d_part + size_t(nb_total) * EA_SUMS);
      c->launches += 2;
      nb_total += nb;
    }
    off += nres[i];
  }
  CU(cudaGetLastError());
  if (raw && total) CU(cudaMemcpyAsync(raw, d_raw, size_t(total) * 8, cudaMemcpyDeviceToHost, st));
  if (residuals && total) CU(cudaMemcpyAsync(residuals, d_res, size_t(total) * 8, cudaMemcpyDeviceToHost, st));
  if (jac && total) CU(cudaMemcpyAsync(jac, d_jac, size_t(total) * 48, cudaMemcpyDeviceToHost, st));
  if (failed) CU(cudaMemcpyAsync(failed, c->d_failed, sizeof(int), cudaMemcpyDeviceToHost, st));
  std::vector<double> h(size_t(std::max(nb_total, 1)) * EA_SUMS, 0.0);
  if (nb_total) CU(cudaMemcpyAsync(h.data(), d_part, size_t(nb_total) * EA_SUMS * 8, cudaMemcpyDeviceToHost, st));
  CU(cudaStreamSynchronize(st));
  if (sums28) {
    double t[EA_SUMS] = {0};
    for (int b2 = 0; b2 < nb_total; ++b2) for (int k = 0; k < EA_SUMS; ++k) t[k] += h[size_t(b2) * EA_SUMS + k];
    sums28[0] = t[28];
    for (int k = 0; k < 6; ++k) sums28[1 + k] = t[21 + k];
    for (int k = 0; k < 21; ++k) sums28[7 + k] = t[k];
  }
  return EA_OK;
}

int ea_solve_views(ea_context* c, int n_views, const ea_view* views, int level, double* pose7, const ea_solve_params* sp,
                   ea_summary* summary) {
  if (!c || !views || n_views < 1 || !pose7 || !sp) return ea_fail(EA_ERR_INVALID_ARG, "null argument");
  if (sp->point_stride < 1) return ea_fail(EA_ERR_INVALID_ARG, "point_stride must be >= 1");
  CU(cudaSetDevice(c->device));
  cudaStream_t st = c->stream;
  std::vector<int> nres(n_views), nb(n_views);
  int total = 0, nb_total = 0;
  for (int i = 0; i < n_views; ++i) {
    int rc = check_view(views[i], level);
    if (rc) return rc;
    int n_pts = 0;
    rc = ea_frameset_get_num_points(views[i].ref, views[i].ref_slot, level, &n_pts);
    if (rc) return rc;
    nres[i] = (n_pts + sp->point_stride - 1) / sp->point_stride;
    nb[i] = nres[i] > 0 ? std::max(1, std::min(c->sm_count, (nres[i] + 2047) / 2048)) : 0;
    total += nres[i]; nb_total += nb[i];
  }
  ShardState* d_state = nullptr;
  double *d_partials = nullptr, *d_sums = nullptr;
  int* h_done = nullptr;
  CU(cudaMalloc((void**)&d_state, sizeof(ShardState)));
  CU(cudaMalloc((void**)&d_partials, size_t(std::max(nb_total, 1)) * EA_SUMS * 8));
  CU(cudaMalloc((void**)&d_sums, EA_SUMS * 8));
  CU(cudaHostAlloc((void**)&h_done, sizeof(int), cudaHostAllocDefault));
  CU(cudaMemcpyAsync(c->d_pose, pose7, 56, cudaMemcpyHostToDevice, st));
  k_shard_init<<<1, 1, 0, st>>>(d_state, c->d_pose);
  c->launches++;
  const int max_evals = sp->max_num_iterations + 2;
  int evals = 0;
  *h_done = 0;
  while (evals < max_evals && !*h_done) {
    const int chunk = std::min(8, max_evals - evals);
    for (int it = 0; it < chunk; ++it) {
      int boff = 0;
      for (int i = 0; i < n_views; ++i) {
        if (!nb[i]) continue;
        const ea_view& v = views[i];
        const EaLevelDesc& rd = v.ref->h_desc[size_t(v.ref_slot) * EA_MAX_LEVELS + level];
        const EaLevelDesc& nd = v.now->h_desc[size_t(v.now_slot) * EA_MAX_LEVELS + level];
        EaViewXf X; view_xf(v, X);
        const double ids = v.ref->inv_depth_unit;
        if (rd.pts_mode == EA_POINTS_XYZ)
          k_view_eval_sums<true><<<nb[i], GV_THREADS, 0, st>>>(rd, nd, v.ref->geom[level], v.now->geom[level], X, ids, *sp, d_state->cand, &d_state->done, 0, nres[i], d_partials + size_t(boff) * EA_SUMS);
        else
          k_view_eval_sums<false><<<nb[i], GV_THREADS, 0, st>>>(rd, nd, v.ref->geom[level], v.now->geom[level], X, ids, *sp, d_state->cand, &d_state->done, 0, nres[i], d_partials + size_t(boff) * EA_SUMS);
        boff += nb[i];
        c->launches++;
      }
      k_shard_reduce<<<1, 32, 0, st>>>(d_state, d_partials, nb_total, d_sums);
      k_shard_lm<<<1, 1, 0, st>>>(d_state, d_sums, *sp);
      c->launches += 2;
    }
    evals += chunk;
    CU(cudaMemcpyAsync(h_done, &d_state->done, sizeof(int), cudaMemcpyDeviceToHost, st));
    CU(cudaStreamSynchronize(st));
  }
  ShardState h;
  CU(cudaMemcpy(&h, d_state, sizeof h, cudaMemcpyDeviceToHost));
  for (int i = 0; i < 7; ++i) pose7[i] = h.lm.x[i];
  if (summary) {
    summary->termination = h.done ? h.lm.term : EA_TERM_NO_CONVERGENCE;
    summary->iterations = h.lm.iter; summary->accepted = h.lm.accepted; summary->rejected = h.lm.rejected;
    summary->n_residuals = total; summary->evaluations = h.lm.evals;
    summary->initial_cost = h.lm.initial_cost; summary->final_cost = h.lm.cost;
  }
  cudaFree(d_state); cudaFree(d_partials); cudaFree(d_sums); cudaFreeHost(h_done);
  return EA_OK;
}

}  // extern "C"
