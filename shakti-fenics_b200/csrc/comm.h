// Multi-GPU plumbing: one process per GPU, NCCL over NVLink/NVSwitch, loaded at run time so
// that the single-GPU path has no NCCL dependency.  Replaces MPI.COMM_WORLD (reference
// source/main.py:11) and the DOLFINx/PETSc ghost scatters and VecNorm all-reduces that
// solvers.py:179,197,229 trigger.
#pragma once
#include "common.h"
#include "prep.h"

namespace shakti {

struct Comm {
  int rank = 0, nranks = 1;
  bool active() const { return nranks > 1; }
};
Comm& comm();

void comm_unique_id(uint8_t id[128]);
void comm_init(const uint8_t id[128], int rank, int nranks, int device);
void comm_finalize();
// in-place sum of `count` doubles (device memory) over all ranks, on stream s
void comm_allreduce_sum(double* dev, int count, cudaStream_t s);

// Halo exchange plan of one level (fine mesh or an AMG level): owner -> ghost copies.
struct HaloPlan {
  struct Peer { int rank; int32_t send_off, send_cnt, recv_begin, recv_cnt; };
  std::vector<Peer> peers;
  DevBuf<int32_t> send_idx;   // concatenated owned local ids to pack
  DevBuf<double> send_buf;    // up to 4 fields packed
  int32_t n_send = 0;
  void build(const std::vector<Neighbor>& nbrs);
  // v: n_local vector; ghosts [recv_begin, ...) are overwritten with the owners' values
  void exchange(double* v, cudaStream_t s);
};

}  // namespace shakti
