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
// in-place sum / max of `count` doubles (device memory) over all ranks, on stream s
void comm_allreduce_sum(double* dev, int count, cudaStream_t s);
void comm_allreduce_max(double* dev, int count, cudaStream_t s);
// recv[r*count .. (r+1)*count) = rank r's send[0..count)   (device memory; copy when one rank)
void comm_allgather(const double* send, double* recv, int count, cudaStream_t s);
// host-level helpers for set-up phases (synchronise the stream)
double comm_host_sum(double v, cudaStream_t s);
double comm_host_max(double v, cudaStream_t s);
std::vector<double> comm_host_allgather(double v, cudaStream_t s);
// personalised all-to-all of small lists: out[r] goes to rank r; returns what every rank sent here
std::vector<std::vector<double>> comm_exchange_lists(const std::vector<std::vector<double>>& out, cudaStream_t s);

// Halo exchange plan of one level (fine mesh or an AMG level): owner -> ghost copies.
struct HaloPlan {
  struct Peer { int rank; int32_t send_off, send_cnt, recv_begin, recv_cnt; };
  std::vector<Peer> peers;
  DevBuf<int32_t> send_idx;   // concatenated owned local ids to pack
  DevBuf<double> send_buf;
  DevBuf<float> send_buf_f;
  int32_t n_send = 0;
  void build(const std::vector<Neighbor>& nbrs);
  // v: n_local vector; ghosts [recv_begin, ...) are overwritten with the owners' values
  void exchange(double* v, cudaStream_t s);
  void exchange(float* v, cudaStream_t s);
  // same for `width` doubles per vertex, v laid out [vertex][width]
  void exchange_block(double* v, int width, cudaStream_t s);
  // pre-packed variant: `sendbuf` holds width doubles per entry of the send list (in send-list
  // order); the values for ghost g (local id ghost_base + k) arrive in recvbuf[k*width ...]
  void exchange_packed(const double* sendbuf, double* recvbuf, int32_t ghost_base, int width, cudaStream_t s);
  int32_t n_send_total() const { return n_send; }
  std::vector<int32_t> send_idx_host;   // copy of the packed send list (set-up phases)
};

}  // namespace shakti
