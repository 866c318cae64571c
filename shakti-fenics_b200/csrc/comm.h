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
  bool p2p = false;   // peer-mapped symmetric heap available: halo / small collectives run as our own kernels over NVLink
  bool active() const { return nranks > 1; }
};
// The communicator of the calling code.  Inside a CommSerialScope it is the trivial one-rank communicator:
// that is how the replicated coarse part of the AMG hierarchy runs the ordinary serial code on every rank.
Comm& comm();
struct CommSerialScope {
  CommSerialScope();
  ~CommSerialScope();
};

// ---- symmetric heap: one cudaMalloc region per rank, mapped into every other rank with CUDA IPC, so a
// kernel can store straight into a neighbour's memory over NVLink.  Offsets are per rank (allocation
// sequences differ); whoever needs a peer's offset exchanges it once at set-up.
size_t p2p_alloc(size_t bytes);            // 256-byte aligned offset; throws when the heap is exhausted
void p2p_free(size_t off);
char* p2p_local(size_t off);
char* p2p_peer(int rank, size_t off);
int heap_selftest(size_t heap_bytes, int rounds);   // host-only check of the allocator (tests)
int p2p_error();                           // != 0 after a kernel-side wait timed out (checked at host syncs)
// host-level allgather of k doubles per rank (set-up phases): out[r*k + i]
std::vector<double> comm_host_allgather_k(const double* v, int k, cudaStream_t s);

// One-shot all-to-all of a small block per rank through the symmetric heap: every rank stores its block
// into every peer's buffer, raises a flag there, waits for the peers' flags and then either leaves the
// gathered blocks in `out` or reduces them in rank order (bitwise identical on all ranks).  One kernel,
// no NCCL call: usable inside CUDA graphs.
struct P2pGather {
  int stride = 0;                          // doubles per rank and slot
  size_t buf_off = 0, flag_off = 0;
  bool built = false;
  DevBuf<char*> peer_buf;
  DevBuf<unsigned long long*> peer_flag;
  DevBuf<unsigned long long> ctr;          // [0] epoch, [1] arrival counter
  DevBuf<int32_t> off, cnt;                // gatherv layout (allgatherv only)
  void build(int stride_doubles, cudaStream_t s);
  void release();
  ~P2pGather() { release(); }
  // dev[0..count) <- sum / max over ranks
  void allreduce(double* dev, int count, bool is_max, cudaStream_t s);
  // out[off[r] + t] = rank r's local[t], t < cnt[r]   (counts set by set_counts; TS source, TO output type)
  void set_counts(const std::vector<int32_t>& counts);
  template <class TS, class TO> void allgatherv(const TS* local, TO* out, cudaStream_t s);
  int my_count = 0, max_count = 0;
};

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
struct HaloPeerDev {     // per neighbour, as the P2P exchange kernel sees it
  char* r_stage;                   // neighbour's staging buffer (slot 0) -- where MY values go
  unsigned long long* r_flag;      // neighbour's flag for me
  unsigned long long r_slot_stride;  // bytes between the neighbour's two slots
  int32_t r_recv_off;              // entry offset of my range inside the neighbour's staging slot
  int32_t send_off, send_cnt;      // my send list range
  int32_t recv_off, recv_cnt, recv_begin;   // my staging range (entries) and ghost range (local ids)
};
struct HaloPlan {
  struct Peer { int rank; int32_t send_off, send_cnt, recv_begin, recv_cnt; };
  std::vector<Peer> peers;
  DevBuf<int32_t> send_idx;   // concatenated owned local ids to pack
  DevBuf<double> send_buf;
  DevBuf<float> send_buf_f;
  int32_t n_send = 0, n_recv = 0;
  // P2P path (comm().p2p): two staging slots + one flag per neighbour in the symmetric heap
  bool p2p = false;
  int cap_bytes = 0;               // bytes per entry the staging slots are sized for
  size_t stage_off = 0, flag_off = 0;
  unsigned long long slot_stride = 0;
  DevBuf<HaloPeerDev> dpeers;
  DevBuf<unsigned long long> ctr;  // [0] epoch, [1] arrival counter
  HaloPlan() = default;
  HaloPlan(const HaloPlan&) = delete;
  HaloPlan& operator=(const HaloPlan&) = delete;
  ~HaloPlan() { release_p2p(); }
  void release_p2p();
  void ensure_p2p(int bytes_per_entry, cudaStream_t s);   // collective: (re)allocates staging and exchanges addresses
  template <class T> void exchange_p2p(const T* src, const int32_t* idx, T* dst, int width, cudaStream_t s);
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
