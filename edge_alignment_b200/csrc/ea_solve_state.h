// Device-resident state of the task-graph solve (ea_solve_tasks.cu).
#pragma once
#include "ea_solve.cuh"

#define EA_MAX_CHUNKS 32          // chunks (tasks) per evaluation of one pair
#define EA_QUEUE_CAP 65536        // ring capacity, power of two, >> window * EA_MAX_CHUNKS + grid

struct EaPairState {              // one per pair of the batch
  EaLmState lm;
  double cand[7];                 // candidate pose of the evaluation in flight
  const void* pts;
  const float* dt;
  float2 affine;
  int n_res, level, pts_mode, n_chunks;
  int remaining;                  // chunks of the evaluation in flight that have not landed yet
  int pad;
  double partial[EA_MAX_CHUNKS][EA_SUMS + 3];   // per-chunk sums, added in chunk order by the last finisher
};

struct EaQueue {
  unsigned head, tail;            // consumer / producer tickets
  int next_pair, done_pairs;
};
