// Internal host-side declarations shared by the .cu translation units (not part of the ABI).
#pragma once
#include <cuda_runtime.h>

#include <string>
#include <vector>

#include "ea_device.cuh"

#include "ea_solve.cuh"

#define EA_TRACE_DOUBLES 6    // {level, iteration, cost of the accepted iterate, cost at the candidate, radius, decision}

struct EaSolveArgs {
  const EaLevelDesc* ref_desc;  // [ref slots][EA_MAX_LEVELS]
  const EaLevelDesc* now_desc;  // [now slots][EA_MAX_LEVELS]
  const int32_t* ref_slots;     // [n_pairs] device
  const int32_t* now_slots;     // [n_pairs] device
  const int32_t* pose_index;    // [n_pairs] device or null (identity)
  const int32_t* order;         // [n_pairs] device or null: processing order of the work queue (longest first)
  double* poses;                // [*][7] device, in/out
  int* work_counter;            // device: next pair index of the dynamic work queue (zeroed per launch)
  struct EaHelpBoard* boards;   // [grid] tail-helper boards of the persistent kernel (zeroed per launch), or null: no helpers
  unsigned long long* debug;    // null, or [grid][6] cycle counters of each CTA's thread 0 (EA_SOLVE_DEBUG=1, development aid)
  ea_summary* summaries;        // [n_pairs][n_levels] device or null
  double* trace;                // [trace_cap][EA_TRACE_DOUBLES] device or null: one record per evaluation (single-pair solves only)
  int* trace_count;             // device counter of records written
  int trace_cap;
  int n_pairs, n_levels, coarsest, finest;
  double inv_depth_scale;
  EaLevelGeom ref_geom[EA_MAX_LEVELS];
  EaLevelGeom now_geom[EA_MAX_LEVELS];
  int ref_cap[EA_MAX_LEVELS];   // point-list capacity per level (guards a truncated list)
  ea_solve_params sp;
  EaEvalConsts loss;                 // sp's loss in fp32 (ea_eval_consts), for the evaluation loop
};

cudaError_t ea_launch_solve_batch(const EaSolveArgs& A, int cluster_size, int sm_count, cudaStream_t stream);
size_t ea_help_boards_bytes(int sm_count);
cudaError_t ea_launch_gather_probe(const EaSolveArgs& A, int level, int slices, int repeats, float* d_sink, cudaStream_t stream);
int ea_probe_gather_device(ea_context* c, int n, ea_frameset* ref, const int32_t* d_ref_slots, ea_frameset* now, const int32_t* d_now_slots,
                           const double* d_poses7, int level, int repeats, float* ms, double* point_gathers);
cudaError_t ea_launch_eval_points(const EaLevelDesc& rd, const EaLevelDesc& nd, const EaLevelGeom& rg,
                                  const EaLevelGeom& ng, double inv_depth_scale, const ea_solve_params& sp,
                                  const double* d_pose7, int n_res, double* d_raw, double* d_res, double* d_jac,
                                  int* d_failed, cudaStream_t stream);
cudaError_t ea_launch_eval_sums(const EaLevelDesc& rd, const EaLevelDesc& nd, const EaLevelGeom& rg,
                                const EaLevelGeom& ng, double inv_depth_scale, const ea_solve_params& sp,
                                const double* d_pose7, const int* d_done, int j_begin, int j_end, int n_blocks, double* d_sums,
                                cudaStream_t stream);

// ---- preprocessing (ea_preprocess.cu) ---------------------------------------------------------------
struct EaPrepLevel {            // device pointers of one pyramid level, slot-major pools
  uint8_t* bgr;                 // [slots][h][w][3]   (level 0: null, the caller's input is read directly)
  void* depth;                  // [slots][h][w] u16 or f32 (level 0: null)
  uint32_t* edge_bits;          // [slots][h][words]  raw Laplacian>threshold mask, 1 bit / pixel
  uint32_t* ref_bits;           // [slots][h][words]  edge & depth>0
  uint32_t* med_bits;           // [slots][h][words]  3x3 median of edge_bits
  float* dt;                    // [slots][h + 2 PAD][dt_pitch] replicate-padded (EA_DT_PAD, ea_device.cuh)
  int dt_pitch; size_t dt_slot; // floats per padded row / per slot
  float4* pts;                  // [slots][cap]
  int w, h, words, cap;
};
struct EaCannyCfg { double low, high; int l2, on_color; };
struct EaScratch {              // Canny / exact-EDT work buffers, [n_slots][stride] each (stride = level-0 pixels)
  uint8_t* gray; int* mag; short2* dxy; uint8_t* map; size_t stride;
};
struct EaPrepArgs {
  EaPrepLevel lv[EA_MAX_LEVELS];
  int n_levels;
  const int32_t* slots;         // [n] device: destination slot of frame i
  const uint8_t* in_bgr;        // [n][h0][w0][3]
  const void* in_depth;         // [n][h0][w0] u16 raw units or f32 metres (depth_type), or null
  const uint8_t* in_mask;       // [n][h0][w0] or null: reference points only where mask > 0 (get_aX_mask, utils.cpp:283-369)
  const uint8_t* in_now_mask;   // [n][h0][w0] or null: now-frame edges only where mask > 1 (get_distance_transform2_masked, utils.cpp:124-128)
  int* n_pts;                   // [slots][EA_MAX_LEVELS]
  unsigned* dt_minmax;          // [slots][EA_MAX_LEVELS][2]  (min,max of the fixed-point DT)
  float2* dt_affine;            // [slots][EA_MAX_LEVELS]     {scale, shift} of the min-max normalisation
  int* overflow;                // [slots][EA_MAX_LEVELS]: that point list was truncated at cap (rewritten by every compaction)
  int n, roles, grad_threshold, use_median, dt_normalize;
  int edge_detector, dt_kind;
  int depth_type;               // 0 = u16 raw units, 1 = f32 metres
  int zero_to_one;              // src/SolveEA.cpp:69: an edge pixel without depth is kept at Z = 1
  float depth_one;              // the stored value meaning Z = 1 (depth_scale for u16, 1.0 for f32)
  EaCannyCfg canny;
  EaScratch scratch;
};
// enqueue the whole preprocessing pipeline for n frames; returns number of kernel launches via *launches
// Optional second lane for big batches: the distance transforms (DRAM / latency-bound, few warps) of one part of the frames run on
// an auxiliary stream beside the edge kernels (issue-bound) of the next part.  Streams and events belong to the context.
#define EA_PREP_MAX_PARTS 4
struct EaPrepPipe {
  cudaStream_t aux[EA_PREP_MAX_PARTS] = {};
  cudaEvent_t ev_front[EA_PREP_MAX_PARTS] = {}, ev_dt[EA_PREP_MAX_PARTS] = {};
  bool ready = false;
};
cudaError_t ea_launch_preprocess(const EaPrepArgs& A, int sm_count, cudaStream_t stream, int* launches, const EaPrepPipe* pipe = nullptr);
__host__ __device__ inline float* ea_dt_origin(const EaPrepLevel& L, int slot) {   // pixel (0,0) of a slot's padded DT
  return L.dt + size_t(slot) * L.dt_slot + size_t(EA_DT_PAD) * L.dt_pitch + EA_DT_PAD;
}
cudaError_t ea_launch_dt_normalized_copy(const float* origin, int pitch, const float2* affine, int w, int h, float* out, cudaStream_t stream);
cudaError_t ea_launch_dt_import(const float* src, float* origin, int w, int h, int pitch, cudaStream_t stream);
cudaError_t ea_launch_dt_fill_pad(const EaPrepLevel& L, const int32_t* d_slots, int n, cudaStream_t stream);
cudaError_t ea_launch_canny_level(const EaPrepArgs& A, int l, const uint8_t* bgr, const void* depth, size_t frame_stride_px,
                                  bool slot_indexed, const EaCannyCfg& cfg, const EaScratch& S, cudaStream_t stream, int* launches);
cudaError_t ea_launch_exact_edt_level(const EaPrepArgs& A, int l, const EaScratch& S, cudaStream_t stream, int* launches);
cudaError_t ea_launch_unpack_mask(const uint32_t* bits, int w, int h, int words, int median, uint8_t* out,
                                  cudaStream_t stream);

// ---- host-side objects behind the opaque ABI handles ------------------------------------------------
struct ea_context {
  int device = 0, sm_count = 0, cc_major = 0, cc_minor = 0;
  cudaStream_t stream = nullptr;
  bool own_stream = false;
  int64_t launches = 0;
  // small scratch
  double* d_pose = nullptr;      // [7]
  int* d_failed = nullptr;
  int* d_work = nullptr;         // work-queue counter of the batched solve
  void* d_boards = nullptr; size_t boards_bytes = 0;        // tail-helper boards of the persistent solve kernel
  unsigned long long* d_debug = nullptr; double debug_sum[6] = {0, 0, 0, 0, 0, 0}; long debug_launches = 0;   // EA_SOLVE_DEBUG=1
  double* d_sums = nullptr;      // [max blocks][EA_SUMS]
  int32_t* d_idx = nullptr;      // scratch slot indices
  size_t idx_cap = 0;
  void* d_tmp = nullptr;
  size_t tmp_cap = 0;
  void* d_views = nullptr; size_t views_cap = 0;   // ea_solve_views: view descriptors, control block and CTA totals of the persistent kernel
  EaPrepPipe prep_pipe;           // auxiliary streams of the preprocessing pipeline (created on first use)
  // optional profiling: event pairs around preprocessing pipelines [0] and solve launches [1]
  bool profile = false;
  double* d_trace = nullptr; int* d_trace_count = nullptr; int trace_cap = 0;   // ea_solve_traced
  std::vector<cudaEvent_t> ev[2];
};
struct EaProfileScope {   // records a start/stop event pair on the launching stream when profiling is on
  ea_context* c; int kind; cudaStream_t s;
  EaProfileScope(ea_context* c_, int kind_, cudaStream_t s_ = nullptr) : c(c_), kind(kind_), s(s_ ? s_ : c_->stream) {
    if (c->profile) { cudaEvent_t e; cudaEventCreate(&e); cudaEventRecord(e, s); c->ev[kind].push_back(e); }
  }
  ~EaProfileScope() {
    if (c->profile) { cudaEvent_t e; cudaEventCreate(&e); cudaEventRecord(e, s); c->ev[kind].push_back(e); }
  }
};

struct ea_frameset {
  ea_context* ctx = nullptr;
  ea_frame_params p;
  int n_slots = 0;
  EaPrepLevel lv[EA_MAX_LEVELS];
  EaLevelGeom geom[EA_MAX_LEVELS];
  EaLevelDesc* d_desc = nullptr;          // [n_slots][EA_MAX_LEVELS]
  std::vector<EaLevelDesc> h_desc;
  int* d_npts = nullptr;                  // [n_slots][EA_MAX_LEVELS]
  unsigned* d_minmax = nullptr;
  float2* d_affine = nullptr;             // [n_slots][EA_MAX_LEVELS]
  int* d_overflow = nullptr;
  uint8_t* stage_bgr = nullptr;           // [n_slots][h][w][3]  (host-upload staging)
  void* stage_depth = nullptr;
  size_t depth_elem = 2;                  // bytes per depth sample (2: u16, 4: f32)
  double inv_depth_unit = 1.0 / 5000.0;   // metres per stored depth unit
  EaScratch scratch{};                    // only allocated for Canny / exact EDT
  std::vector<void*> allocs;
};


int ea_fail(int code, const char* fmt, ...);
#define CU(call)                                                                                       \
  do {                                                                                                 \
    cudaError_t e_ = (call);                                                                           \
    if (e_ != cudaSuccess) return ea_fail(EA_ERR_CUDA, "%s:%d %s -> %s", __FILE__, __LINE__, #call, cudaGetErrorString(e_)); \
  } while (0)
// Scope guards for the entry points: CU() returns early on an error, these release what the call had acquired by then.
struct EaAsyncBuf {   // cudaMallocAsync / cudaFreeAsync on one stream
  void* p = nullptr; cudaStream_t s = nullptr;
  explicit EaAsyncBuf(cudaStream_t s_) : s(s_) {}
  cudaError_t alloc(size_t bytes) { return cudaMallocAsync(&p, bytes, s); }
  template <class T> T* as() const { return static_cast<T*>(p); }
  ~EaAsyncBuf() { if (p) cudaFreeAsync(p, s); }
  EaAsyncBuf(const EaAsyncBuf&) = delete; EaAsyncBuf& operator=(const EaAsyncBuf&) = delete;
};
struct EaDevBuf {     // cudaMalloc / cudaFree
  void* p = nullptr;
  cudaError_t alloc(size_t bytes) { return cudaMalloc(&p, bytes); }
  template <class T> T* as() const { return static_cast<T*>(p); }
  ~EaDevBuf() { if (p) cudaFree(p); }
  EaDevBuf() = default; EaDevBuf(const EaDevBuf&) = delete; EaDevBuf& operator=(const EaDevBuf&) = delete;
};
struct EaEventPair {
  cudaEvent_t a = nullptr, b = nullptr;
  cudaError_t create() { cudaError_t e = cudaEventCreate(&a); return e != cudaSuccess ? e : cudaEventCreate(&b); }
  ~EaEventPair() { if (a) cudaEventDestroy(a); if (b) cudaEventDestroy(b); }
  EaEventPair() = default; EaEventPair(const EaEventPair&) = delete; EaEventPair& operator=(const EaEventPair&) = delete;
};
int ea_ensure_tmp(ea_context* c, size_t bytes);
extern "C" int ea_check_solve_params(const ea_solve_params* sp);   // shared argument validation of every solve entry point
int ea_solve_batch_device_ordered(ea_context* c, int n, ea_frameset* ref, const int32_t* d_ref_slots, ea_frameset* now,
                                   const int32_t* d_now_slots, double* d_poses7, const int32_t* d_pose_index,
                                   const int32_t* d_order, const ea_solve_params* sp, ea_summary* d_summaries, bool tail_helpers = true);
cudaError_t ea_launch_order_by_work(const ea_summary* d_summaries, int n, int n_levels, int32_t* d_order, cudaStream_t stream);
int ea_preprocess_impl(ea_frameset* fs, int n, const int32_t* d_slots, const uint8_t* d_bgr, const void* d_depth, int roles,
                       const uint8_t* d_mask = nullptr, const uint8_t* d_now_mask = nullptr, cudaStream_t stream = nullptr);
