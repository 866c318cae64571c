// NCCL plumbing (see comm.h).  libnccl is dlopen'ed so the library loads on boxes without it.
#include "comm.h"
#include "device.h"

#include <dlfcn.h>

#include <algorithm>
#include <cstdlib>

#include <cstring>

namespace shakti {

// Minimal NCCL ABI (stable across 2.x): opaque comm, 128-byte unique id, enums below.
typedef struct ncclComm* ncclComm_t;
typedef struct { char internal[128]; } ncclUniqueId;
enum { ncclSuccess = 0 };
enum { ncclFloat32 = 7, ncclFloat64 = 8 };  // ncclDataType_t
enum { ncclSum = 0, ncclMax = 2 };

struct NcclApi {
  void* lib = nullptr;
  int (*GetUniqueId)(ncclUniqueId*) = nullptr;
  int (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
  int (*CommDestroy)(ncclComm_t) = nullptr;
  int (*AllReduce)(const void*, void*, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
  int (*AllGather)(const void*, void*, size_t, int, ncclComm_t, cudaStream_t) = nullptr;
  int (*Send)(const void*, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
  int (*Recv)(void*, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
  int (*GroupStart)() = nullptr;
  int (*GroupEnd)() = nullptr;
  const char* (*GetErrorString)(int) = nullptr;
};

static NcclApi g_api;
static ncclComm_t g_nccl = nullptr;
static Comm g_comm;
static Comm g_serial;            // rank 0 of 1
static int g_serial_depth = 0;   // CommSerialScope nesting (one host thread drives a model)
Comm& comm() { return g_serial_depth > 0 ? g_serial : g_comm; }
CommSerialScope::CommSerialScope() { ++g_serial_depth; }
CommSerialScope::~CommSerialScope() { --g_serial_depth; }

static void load_api() {
  if (g_api.lib) return;
  const char* names[] = {"libnccl.so.2", "libnccl.so"};
  for (const char* n : names) {
    g_api.lib = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
    if (g_api.lib) break;
  }
  if (!g_api.lib) throw Error(SHAKTI_ERR_COMM, std::string("cannot load libnccl.so.2: ") + dlerror());
  auto sym = [&](const char* s) {
    void* p = dlsym(g_api.lib, s);
    if (!p) throw Error(SHAKTI_ERR_COMM, std::string("NCCL symbol missing: ") + s);
    return p;
  };
  g_api.GetUniqueId = (decltype(g_api.GetUniqueId))sym("ncclGetUniqueId");
  g_api.CommInitRank = (decltype(g_api.CommInitRank))sym("ncclCommInitRank");
  g_api.CommDestroy = (decltype(g_api.CommDestroy))sym("ncclCommDestroy");
  g_api.AllReduce = (decltype(g_api.AllReduce))sym("ncclAllReduce");
  g_api.AllGather = (decltype(g_api.AllGather))sym("ncclAllGather");
  g_api.Send = (decltype(g_api.Send))sym("ncclSend");
  g_api.Recv = (decltype(g_api.Recv))sym("ncclRecv");
  g_api.GroupStart = (decltype(g_api.GroupStart))sym("ncclGroupStart");
  g_api.GroupEnd = (decltype(g_api.GroupEnd))sym("ncclGroupEnd");
  g_api.GetErrorString = (decltype(g_api.GetErrorString))sym("ncclGetErrorString");
}

#define SHAKTI_NCCL(call)                                                                    \
  do {                                                                                       \
    int r_ = (call);                                                                         \
    if (r_ != ncclSuccess)                                                                   \
      throw Error(SHAKTI_ERR_COMM, std::string(#call) + ": " + g_api.GetErrorString(r_));    \
  } while (0)

void comm_unique_id(uint8_t id[128]) {
  load_api();
  ncclUniqueId u;
  SHAKTI_NCCL(g_api.GetUniqueId(&u));
  std::memcpy(id, u.internal, 128);
}

// ------------------------------------------------------------------ symmetric heap (CUDA IPC)
struct HeapBlock { size_t off, size; bool free; };
struct P2pState {
  char* heap = nullptr;
  size_t bytes = 0;
  std::vector<char*> peer;           // peer[r]: rank r's heap mapped here (peer[me] = heap)
  std::vector<HeapBlock> blocks;
  int* err_host = nullptr;           // mapped pinned: kernels report a timed-out wait here
  int* err_dev = nullptr;
  unsigned generation = 0;
};
static P2pState g_p2p;
static P2pGather* g_ar = nullptr;    // the small all-reduce (Krylov dots, norms, spectrum bounds)

// first-fit with splitting; SIZE_MAX when nothing fits
static size_t heap_take(std::vector<HeapBlock>& bl, size_t bytes) {
  bytes = (std::max<size_t>(bytes, 1) + 255) & ~(size_t)255;
  for (size_t i = 0; i < bl.size(); ++i) {
    HeapBlock& b = bl[i];
    if (!b.free || b.size < bytes) continue;
    const size_t off = b.off;
    if (b.size > bytes) {
      const HeapBlock rest{b.off + bytes, b.size - bytes, true};
      b.size = bytes;
      b.free = false;
      bl.insert(bl.begin() + i + 1, rest);
    } else {
      b.free = false;
    }
    return off;
  }
  return (size_t)-1;
}
// release with coalescing of free neighbours; false when `off` is not an allocated block
static bool heap_give(std::vector<HeapBlock>& bl, size_t off) {
  for (size_t i = 0; i < bl.size(); ++i) {
    if (bl[i].off != off || bl[i].free) continue;
    bl[i].free = true;
    if (i + 1 < bl.size() && bl[i + 1].free) { bl[i].size += bl[i + 1].size; bl.erase(bl.begin() + i + 1); }
    if (i > 0 && bl[i - 1].free) { bl[i - 1].size += bl[i].size; bl.erase(bl.begin() + i); }
    return true;
  }
  return false;
}
size_t p2p_alloc(size_t bytes) {
  SHAKTI_REQUIRE(g_comm.p2p, "p2p_alloc without a symmetric heap");
  const size_t off = heap_take(g_p2p.blocks, bytes);
  if (off == (size_t)-1)
    throw Error(SHAKTI_ERR_COMM, "symmetric heap exhausted (" + std::to_string(g_p2p.bytes >> 20) +
                                     " MiB): raise SHAKTI_P2P_HEAP_MB");
  return off;
}
void p2p_free(size_t off) { heap_give(g_p2p.blocks, off); }

// host-only self test of the heap allocator (no device needed): a deterministic pseudo-random sequence of
// allocations and releases on a heap of `heap_bytes`; checks alignment, that live blocks never overlap, that
// releasing everything leaves one free block again.  Returns the number of violations.
int heap_selftest(size_t heap_bytes, int rounds) {
  std::vector<HeapBlock> bl(1, HeapBlock{0, heap_bytes, true});
  std::vector<std::pair<size_t, size_t>> live;   // (offset, requested bytes)
  unsigned long long rng = 88172645463325252ULL;
  auto next = [&]() { rng ^= rng << 13; rng ^= rng >> 7; rng ^= rng << 17; return rng; };
  int bad = 0;
  for (int r = 0; r < rounds; ++r) {
    if (live.empty() || next() % 3 != 0) {
      const size_t want = 1 + next() % (heap_bytes / 16);
      const size_t off = heap_take(bl, want);
      if (off == (size_t)-1) continue;             // full: fine
      if (off % 256 != 0 || off + want > heap_bytes) ++bad;
      for (const auto& l : live) {
        const size_t a0 = l.first, a1 = l.first + ((l.second + 255) & ~(size_t)255);
        const size_t b0 = off, b1 = off + ((want + 255) & ~(size_t)255);
        if (a0 < b1 && b0 < a1) ++bad;
      }
      live.push_back({off, want});
    } else {
      const size_t k = next() % live.size();
      if (!heap_give(bl, live[k].first)) ++bad;
      if (heap_give(bl, live[k].first)) ++bad;     // double free must be refused
      live.erase(live.begin() + k);
    }
  }
  for (const auto& l : live)
    if (!heap_give(bl, l.first)) ++bad;
  if (bl.size() != 1 || !bl[0].free || bl[0].off != 0 || bl[0].size != heap_bytes) ++bad;
  return bad;
}
char* p2p_local(size_t off) { return g_p2p.heap + off; }
char* p2p_peer(int rank, size_t off) { return g_p2p.peer[rank] + off; }
int p2p_error() { return g_p2p.err_host ? *(volatile int*)g_p2p.err_host : 0; }

static void p2p_setup(int rank, int nranks) {
  const char* off_env = getenv("SHAKTI_P2P");
  if (off_env && atoi(off_env) == 0) return;
  const char* mb = getenv("SHAKTI_P2P_HEAP_MB");
  const size_t bytes = (size_t)(mb ? std::max(8, atoi(mb)) : 256) << 20;
  // every rank must be able to map every other one; the decision is made collectively so that all
  // ranks take the same path
  int dev = 0, ndev = 0;
  SHAKTI_CUDA(cudaGetDevice(&dev));
  SHAKTI_CUDA(cudaGetDeviceCount(&ndev));
  double ok = 1.0;
  char* heap = nullptr;
  if (cudaMalloc(&heap, bytes) != cudaSuccess) { cudaGetLastError(); ok = 0.0; heap = nullptr; }
  cudaIpcMemHandle_t h;
  std::memset(&h, 0, sizeof(h));
  if (heap && cudaIpcGetMemHandle(&h, heap) != cudaSuccess) { cudaGetLastError(); ok = 0.0; }
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
  double rec[9];
  std::memcpy(rec, &h, 64);
  rec[8] = ok;
  // NaN payloads survive: all-gather only moves bytes
  g_comm.rank = rank; g_comm.nranks = nranks;
  const std::vector<double> all = comm_host_allgather_k(rec, 9, 0);
  bool all_ok = true;
  for (int r = 0; r < nranks; ++r) all_ok &= all[(size_t)r * 9 + 8] == 1.0;
  std::vector<char*> peer(nranks, nullptr);
  double mapped = 1.0;
  if (all_ok) {
    SHAKTI_CUDA(cudaMemset(heap, 0, bytes));
    for (int r = 0; r < nranks && mapped == 1.0; ++r) {
      if (r == rank) { peer[r] = heap; continue; }
      cudaIpcMemHandle_t hr;
      std::memcpy(&hr, &all[(size_t)r * 9], 64);
      void* p = nullptr;
      if (cudaIpcOpenMemHandle(&p, hr, cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) { cudaGetLastError(); mapped = 0.0; }
      peer[r] = (char*)p;
    }
  } else {
    mapped = 0.0;
  }
  const std::vector<double> m_all = comm_host_allgather_k(&mapped, 1, 0);
  bool every = true;
  for (double v : m_all) every &= v == 1.0;
  if (!every) {   // some rank cannot map a peer: everybody stays on NCCL
    for (int r = 0; r < nranks; ++r)
      if (r != rank && peer[r]) cudaIpcCloseMemHandle(peer[r]);
    if (heap) cudaFree(heap);
    return;
  }
  g_p2p.heap = heap;
  g_p2p.bytes = bytes;
  g_p2p.peer = peer;
  g_p2p.blocks.assign(1, HeapBlock{0, bytes, true});
  g_p2p.generation++;
  SHAKTI_CUDA(cudaHostAlloc((void**)&g_p2p.err_host, sizeof(int), cudaHostAllocMapped));
  *g_p2p.err_host = 0;
  SHAKTI_CUDA(cudaHostGetDevicePointer((void**)&g_p2p.err_dev, g_p2p.err_host, 0));
  g_comm.p2p = true;
  g_ar = new P2pGather();
  g_ar->build(64, 0);
}

void comm_init(const uint8_t id[128], int rank, int nranks, int device) {
  SHAKTI_REQUIRE(nranks >= 1 && rank >= 0 && rank < nranks, "bad rank / nranks");
  if (nranks == 1) { g_comm.rank = 0; g_comm.nranks = 1; return; }
  load_api();
  if (device >= 0) SHAKTI_CUDA(cudaSetDevice(device));
  ncclUniqueId u;
  std::memcpy(u.internal, id, 128);
  SHAKTI_NCCL(g_api.CommInitRank(&g_nccl, nranks, u, rank));
  g_comm.rank = rank;
  g_comm.nranks = nranks;
  g_comm.p2p = false;
  p2p_setup(rank, nranks);
}

void comm_finalize() {
  if (g_ar) { delete g_ar; g_ar = nullptr; }
  if (g_p2p.heap) {
    cudaDeviceSynchronize();
    for (size_t r = 0; r < g_p2p.peer.size(); ++r)
      if ((int)r != g_comm.rank && g_p2p.peer[r]) cudaIpcCloseMemHandle(g_p2p.peer[r]);
    cudaFree(g_p2p.heap);
    if (g_p2p.err_host) cudaFreeHost(g_p2p.err_host);
    g_p2p = P2pState();
  }
  if (g_nccl) g_api.CommDestroy(g_nccl);
  g_nccl = nullptr;
  g_comm = Comm();
}

// ------------------------------------------------------------------ kernel-side signalling
__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long* p) {
  unsigned long long v;
  asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_release_sys(unsigned long long* p, unsigned long long v) {
  asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long global_timer_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
  return t;
}
// spin until *flag >= epoch; a wait longer than 20 s means a peer died: report instead of hanging the GPU
__device__ __forceinline__ void wait_flag(const unsigned long long* flag, unsigned long long epoch, int* err) {
  if (ld_acquire_sys(flag) >= epoch) return;
  const unsigned long long t0 = global_timer_ns();
  while (ld_acquire_sys(flag) < epoch) {
    __nanosleep(40);
    if (global_timer_ns() - t0 > 20000000000ull) { *err = 1; __threadfence_system(); return; }
  }
}

// ---- small all-to-all through the heap.  MODE 0: gatherv (out[off[r] + t]); 1: sum; 2: max
template <int MODE, class TS, class TO>
__global__ void __launch_bounds__(256)
p2p_gather_kernel(int nranks, int me, int stride, char* const* __restrict__ peer_buf, unsigned long long* const* __restrict__ peer_flag,
                  const unsigned long long* __restrict__ my_flag, const double* __restrict__ my_buf, unsigned long long* ctr,
                  const TS* local, int count, const int32_t* __restrict__ off, const int32_t* __restrict__ cnt,
                  TO* out, int* err) {   // local may alias out (in-place all-reduce): no __restrict__
  pdl_sync();
  const unsigned long long e = *(volatile unsigned long long*)ctr + 1;
  const size_t slot = (size_t)(e & 1) * nranks * stride;
  const int tid = blockIdx.x * blockDim.x + threadIdx.x, nth = gridDim.x * blockDim.x;
  for (int r = 0; r < nranks; ++r) {
    double* dst = reinterpret_cast<double*>(peer_buf[r]) + slot + (size_t)me * stride;
    for (int t = tid; t < count; t += nth) dst[t] = (double)local[t];
  }
  __threadfence_system();
  __syncthreads();
  __shared__ bool last;
  if (threadIdx.x == 0) last = atomicAdd(ctr + 1, 1ull) == gridDim.x - 1;
  __syncthreads();
  if (last) {
    __threadfence_system();
    if ((int)threadIdx.x < nranks) st_release_sys(peer_flag[threadIdx.x] + me, e);
    if (threadIdx.x == 0) { ctr[1] = 0; ctr[0] = e; }
  }
  if ((int)threadIdx.x < nranks) wait_flag(my_flag + threadIdx.x, e, err);
  __syncthreads();
  const double* buf = my_buf + slot;
  if (MODE == 0) {
    for (int r = 0; r < nranks; ++r) {
      const int c = cnt[r], o = off[r];
      for (int t = tid; t < c; t += nth) out[o + t] = (TO)__ldcg(buf + (size_t)r * stride + t);
    }
  } else {
    for (int t = tid; t < count; t += nth) {
      double acc = __ldcg(buf + t);
      for (int r = 1; r < nranks; ++r) {
        const double v = __ldcg(buf + (size_t)r * stride + t);
        acc = MODE == 1 ? acc + v : fmax(acc, v);
      }
      out[t] = (TO)acc;
    }
  }
}

void P2pGather::build(int stride_doubles, cudaStream_t s) {
  release();
  const int nr = g_comm.nranks, me = g_comm.rank;
  stride = stride_doubles;
  buf_off = p2p_alloc((size_t)2 * nr * stride * sizeof(double));
  flag_off = p2p_alloc((size_t)nr * sizeof(unsigned long long));
  SHAKTI_CUDA(cudaMemsetAsync(p2p_local(flag_off), 0, (size_t)nr * sizeof(unsigned long long), s));
  SHAKTI_CUDA(cudaStreamSynchronize(s));
  ctr.alloc_zero(2, s);
  // offsets differ from rank to rank: publish mine, learn everybody's (also the barrier after zeroing the flags)
  const double mine[2] = {(double)buf_off, (double)flag_off};
  const std::vector<double> all = comm_host_allgather_k(mine, 2, s);
  std::vector<char*> pb(nr);
  std::vector<unsigned long long*> pf(nr);
  for (int r = 0; r < nr; ++r) {
    pb[r] = p2p_peer(r, (size_t)all[2 * r]);
    pf[r] = reinterpret_cast<unsigned long long*>(p2p_peer(r, (size_t)all[2 * r + 1]));
  }
  (void)me;
  peer_buf.upload(pb);
  peer_flag.upload(pf);
  built = true;
}
void P2pGather::release() {
  if (!built) return;
  if (g_comm.p2p) { p2p_free(buf_off); p2p_free(flag_off); }
  built = false;
}
void P2pGather::allreduce(double* dev, int count, bool is_max, cudaStream_t s) {
  SHAKTI_REQUIRE(built && count <= stride, "P2pGather::allreduce: block too large");
  const int nr = g_comm.nranks, me = g_comm.rank;
  const unsigned long long* mf = reinterpret_cast<const unsigned long long*>(p2p_local(flag_off));
  const double* mb = reinterpret_cast<const double*>(p2p_local(buf_off));
  if (is_max)
    SHAKTI_LAUNCH((p2p_gather_kernel<2, double, double>), 1, 64, 0, s, nr, me, stride, peer_buf.p, peer_flag.p, mf, mb, ctr.p, dev, count,
                  nullptr, nullptr, dev, g_p2p.err_dev);
  else
    SHAKTI_LAUNCH((p2p_gather_kernel<1, double, double>), 1, 64, 0, s, nr, me, stride, peer_buf.p, peer_flag.p, mf, mb, ctr.p, dev, count,
                  nullptr, nullptr, dev, g_p2p.err_dev);
}
void P2pGather::set_counts(const std::vector<int32_t>& counts) {
  std::vector<int32_t> o(counts.size(), 0);
  max_count = 0;
  for (size_t r = 0; r < counts.size(); ++r) {
    if (r) o[r] = o[r - 1] + counts[r - 1];
    max_count = std::max(max_count, counts[r]);
  }
  my_count = counts[g_comm.rank];
  SHAKTI_REQUIRE(max_count <= stride, "P2pGather::set_counts: block too large");
  off.upload(o);
  cnt.upload(counts);
}
template <class TS, class TO>
void P2pGather::allgatherv(const TS* local, TO* out, cudaStream_t s) {
  SHAKTI_REQUIRE(built && cnt.p, "P2pGather::allgatherv before build / set_counts");
  const int nr = g_comm.nranks, me = g_comm.rank;
  const unsigned long long* mf = reinterpret_cast<const unsigned long long*>(p2p_local(flag_off));
  const double* mb = reinterpret_cast<const double*>(p2p_local(buf_off));
  const int blocks = std::max(1, std::min(8, (max_count + 1023) / 1024));
  SHAKTI_LAUNCH_PDL((p2p_gather_kernel<0, TS, TO>), blocks, 256, 0, s, nr, me, stride, peer_buf.p, peer_flag.p, mf, mb, ctr.p, local, my_count,
                off.p, cnt.p, out, g_p2p.err_dev);
}
template void P2pGather::allgatherv<float, float>(const float*, float*, cudaStream_t);
template void P2pGather::allgatherv<double, double>(const double*, double*, cudaStream_t);

void comm_allreduce_sum(double* dev, int count, cudaStream_t s) {
  if (!comm().active() || count == 0) return;
  if (comm().p2p && g_ar && count <= g_ar->stride) { g_ar->allreduce(dev, count, false, s); return; }
  SHAKTI_NCCL(g_api.AllReduce(dev, dev, (size_t)count, ncclFloat64, ncclSum, g_nccl, s));
}

void comm_allreduce_max(double* dev, int count, cudaStream_t s) {
  if (!comm().active() || count == 0) return;
  if (comm().p2p && g_ar && count <= g_ar->stride) { g_ar->allreduce(dev, count, true, s); return; }
  SHAKTI_NCCL(g_api.AllReduce(dev, dev, (size_t)count, ncclFloat64, ncclMax, g_nccl, s));
}

void comm_allgather(const double* send, double* recv, int count, cudaStream_t s) {
  if (count == 0) return;
  if (!comm().active()) {
    if (send != recv) SHAKTI_CUDA(cudaMemcpyAsync(recv, send, sizeof(double) * count, cudaMemcpyDeviceToDevice, s));
    return;
  }
  SHAKTI_NCCL(g_api.AllGather(send, recv, (size_t)count, ncclFloat64, g_nccl, s));
}

static DevBuf<double>& scratch(size_t n) {
  static DevBuf<double> buf;
  if (buf.n < n) buf.alloc(std::max<size_t>(n, 256));
  return buf;
}

double comm_host_sum(double v, cudaStream_t s) {
  if (!comm().active()) return v;
  DevBuf<double>& b = scratch(1);
  SHAKTI_CUDA(cudaMemcpyAsync(b.p, &v, sizeof(double), cudaMemcpyHostToDevice, s));
  comm_allreduce_sum(b.p, 1, s);
  SHAKTI_CUDA(cudaMemcpyAsync(&v, b.p, sizeof(double), cudaMemcpyDeviceToHost, s));
  SHAKTI_CUDA(cudaStreamSynchronize(s));
  return v;
}
double comm_host_max(double v, cudaStream_t s) {
  if (!comm().active()) return v;
  DevBuf<double>& b = scratch(1);
  SHAKTI_CUDA(cudaMemcpyAsync(b.p, &v, sizeof(double), cudaMemcpyHostToDevice, s));
  comm_allreduce_max(b.p, 1, s);
  SHAKTI_CUDA(cudaMemcpyAsync(&v, b.p, sizeof(double), cudaMemcpyDeviceToHost, s));
  SHAKTI_CUDA(cudaStreamSynchronize(s));
  return v;
}
std::vector<double> comm_host_allgather_k(const double* v, int k, cudaStream_t s) {
  const int nr = comm().nranks;
  std::vector<double> out((size_t)nr * k);
  if (!comm().active()) { std::copy(v, v + k, out.begin()); return out; }
  DevBuf<double>& b = scratch((size_t)k * (1 + nr));
  SHAKTI_CUDA(cudaMemcpyAsync(b.p, v, sizeof(double) * k, cudaMemcpyHostToDevice, s));
  SHAKTI_NCCL(g_api.AllGather(b.p, b.p + k, (size_t)k, ncclFloat64, g_nccl, s));
  SHAKTI_CUDA(cudaMemcpyAsync(out.data(), b.p + k, sizeof(double) * nr * k, cudaMemcpyDeviceToHost, s));
  SHAKTI_CUDA(cudaStreamSynchronize(s));
  return out;
}
std::vector<double> comm_host_allgather(double v, cudaStream_t s) {
  const int nr = comm().nranks;
  std::vector<double> out(nr, v);
  if (!comm().active()) return out;
  DevBuf<double>& b = scratch(1 + nr);
  SHAKTI_CUDA(cudaMemcpyAsync(b.p, &v, sizeof(double), cudaMemcpyHostToDevice, s));
  comm_allgather(b.p, b.p + 1, 1, s);
  SHAKTI_CUDA(cudaMemcpyAsync(out.data(), b.p + 1, sizeof(double) * nr, cudaMemcpyDeviceToHost, s));
  SHAKTI_CUDA(cudaStreamSynchronize(s));
  return out;
}

std::vector<std::vector<double>> comm_exchange_lists(const std::vector<std::vector<double>>& out, cudaStream_t s) {
  const int nr = comm().nranks, me = comm().rank;
  std::vector<std::vector<double>> in(nr);
  if (!comm().active()) { in[0] = out[0]; return in; }
  // counts[r][q] = how many doubles rank r sends to rank q
  std::vector<double> mine(nr);
  for (int r = 0; r < nr; ++r) mine[r] = (double)out[r].size();
  DevBuf<double>& b = scratch((size_t)nr + (size_t)nr * nr);
  SHAKTI_CUDA(cudaMemcpyAsync(b.p, mine.data(), sizeof(double) * nr, cudaMemcpyHostToDevice, s));
  comm_allgather(b.p, b.p + nr, nr, s);
  std::vector<double> counts((size_t)nr * nr);
  SHAKTI_CUDA(cudaMemcpyAsync(counts.data(), b.p + nr, sizeof(double) * nr * nr, cudaMemcpyDeviceToHost, s));
  SHAKTI_CUDA(cudaStreamSynchronize(s));
  size_t n_out = 0, n_in = 0;
  for (int r = 0; r < nr; ++r) { n_out += out[r].size(); n_in += (size_t)counts[(size_t)r * nr + me]; }
  DevBuf<double> sb, rb;
  sb.alloc(std::max<size_t>(n_out, 1));
  rb.alloc(std::max<size_t>(n_in, 1));
  std::vector<double> flat;
  flat.reserve(n_out);
  for (int r = 0; r < nr; ++r) flat.insert(flat.end(), out[r].begin(), out[r].end());
  if (n_out) SHAKTI_CUDA(cudaMemcpyAsync(sb.p, flat.data(), sizeof(double) * n_out, cudaMemcpyHostToDevice, s));
  SHAKTI_NCCL(g_api.GroupStart());
  size_t so = 0, ro = 0;
  for (int r = 0; r < nr; ++r) {
    const size_t cs = out[r].size(), cr = (size_t)counts[(size_t)r * nr + me];
    if (r != me) {
      if (cs) SHAKTI_NCCL(g_api.Send(sb.p + so, cs, ncclFloat64, r, g_nccl, s));
      if (cr) SHAKTI_NCCL(g_api.Recv(rb.p + ro, cr, ncclFloat64, r, g_nccl, s));
    }
    so += cs;
    ro += cr;
  }
  SHAKTI_NCCL(g_api.GroupEnd());
  std::vector<double> rflat(std::max<size_t>(n_in, 1));
  if (n_in) SHAKTI_CUDA(cudaMemcpyAsync(rflat.data(), rb.p, sizeof(double) * n_in, cudaMemcpyDeviceToHost, s));
  SHAKTI_CUDA(cudaStreamSynchronize(s));
  ro = 0;
  for (int r = 0; r < nr; ++r) {
    const size_t cr = (size_t)counts[(size_t)r * nr + me];
    if (r == me) in[r] = out[r];
    else in[r].assign(rflat.begin() + ro, rflat.begin() + ro + cr);
    ro += cr;
  }
  return in;
}

void HaloPlan::build(const std::vector<Neighbor>& nbrs) {
  peers.clear();
  std::vector<int32_t> idx;
  for (const auto& nb : nbrs) {
    Peer p;
    p.rank = nb.rank;
    p.send_off = (int32_t)idx.size();
    p.send_cnt = (int32_t)nb.send_local.size();
    p.recv_begin = nb.recv_begin;
    p.recv_cnt = nb.recv_count;
    idx.insert(idx.end(), nb.send_local.begin(), nb.send_local.end());
    peers.push_back(p);
  }
  n_send = (int32_t)idx.size();
  n_recv = 0;
  for (const auto& p : peers) n_recv += p.recv_cnt;
  send_idx_host = idx;
  if (n_send) {
    send_idx.upload(idx);
    send_buf.alloc((size_t)n_send);
  }
  release_p2p();
  p2p = comm().active() && comm().p2p;
  cap_bytes = 0;
}

// ---- P2P halo exchange: ONE kernel packs my interface values straight into the neighbours' staging
// slots over NVLink, raises their flags, waits for the neighbours' flags and unpacks my ghosts.
// Two slots (epoch parity) suffice because neighbour relations are symmetric: a rank can start
// exchange e+1 only after it consumed exchange e, and it needs my e+1 signal before it can reach e+2.
template <class T>
__global__ void __launch_bounds__(256)
halo_p2p_kernel(int npeers, const HaloPeerDev* __restrict__ peers, char* my_stage, unsigned long long slot_stride,
                const unsigned long long* __restrict__ my_flags, unsigned long long* ctr, const T* __restrict__ src,
                const int32_t* __restrict__ idx, T* dst, int width, int* err) {
  pdl_sync();
  const unsigned long long e = *(volatile unsigned long long*)ctr + 1;
  const unsigned long long par = e & 1ull;
  const int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x, nth = (int64_t)gridDim.x * blockDim.x;
  for (int p = 0; p < npeers; ++p) {
    const HaloPeerDev pe = peers[p];
    T* rs = reinterpret_cast<T*>(pe.r_stage + par * pe.r_slot_stride) + (size_t)pe.r_recv_off * width;
    const int64_t n = (int64_t)pe.send_cnt * width;
    if (width == 1) {
      if (idx) for (int64_t t = tid; t < n; t += nth) rs[t] = src[idx[pe.send_off + t]];
      else for (int64_t t = tid; t < n; t += nth) rs[t] = src[pe.send_off + t];
    } else {
      for (int64_t t = tid; t < n; t += nth) {
        const int64_t k = t / width, w = t - k * width;
        rs[t] = src[(idx ? (int64_t)idx[pe.send_off + k] : (int64_t)(pe.send_off + k)) * width + w];
      }
    }
  }
  __threadfence_system();
  __syncthreads();
  __shared__ bool last;
  if (threadIdx.x == 0) last = atomicAdd(ctr + 1, 1ull) == gridDim.x - 1;
  __syncthreads();
  if (last) {
    __threadfence_system();
    if ((int)threadIdx.x < npeers) st_release_sys(peers[threadIdx.x].r_flag, e);
    if (threadIdx.x == 0) { ctr[1] = 0; ctr[0] = e; }
  }
  for (int p = threadIdx.x; p < npeers; p += blockDim.x) wait_flag(my_flags + p, e, err);
  __syncthreads();
  for (int p = 0; p < npeers; ++p) {
    const HaloPeerDev pe = peers[p];
    const T* ls = reinterpret_cast<const T*>(my_stage + par * slot_stride) + (size_t)pe.recv_off * width;
    T* d = dst + (int64_t)pe.recv_begin * width;
    const int64_t n = (int64_t)pe.recv_cnt * width;
    for (int64_t t = tid; t < n; t += nth) d[t] = __ldcg(ls + t);
  }
}

void HaloPlan::release_p2p() {
  if (cap_bytes > 0 && g_comm.p2p && g_p2p.heap) { p2p_free(stage_off); p2p_free(flag_off); }
  cap_bytes = 0;
}

void HaloPlan::ensure_p2p(int bytes_per_entry, cudaStream_t s) {
  if (bytes_per_entry <= cap_bytes) return;
  // collective (every rank of the plan calls with the same width): new staging, fresh flags and epoch
  SHAKTI_CUDA(cudaStreamSynchronize(s));
  release_p2p();
  cap_bytes = std::max(bytes_per_entry, 8);
  const int np = (int)peers.size();
  slot_stride = (((unsigned long long)std::max(n_recv, 1) * cap_bytes) + 255ull) & ~255ull;
  stage_off = p2p_alloc((size_t)2 * slot_stride);
  flag_off = p2p_alloc((size_t)std::max(np, 1) * sizeof(unsigned long long));
  SHAKTI_CUDA(cudaMemsetAsync(p2p_local(flag_off), 0, (size_t)std::max(np, 1) * sizeof(unsigned long long), s));
  ctr.alloc_zero(2, s);
  SHAKTI_CUDA(cudaStreamSynchronize(s));
  // tell every neighbour where its values go here: stage offset, slot stride, entry offset of its range, flag address
  const int nr = comm().nranks;
  std::vector<std::vector<double>> out(nr);
  int32_t roff = 0;
  std::vector<int32_t> my_roff(np);
  for (int i = 0; i < np; ++i) {
    my_roff[i] = roff;
    out[peers[i].rank] = {(double)stage_off, (double)slot_stride, (double)roff, (double)(flag_off + (size_t)i * sizeof(unsigned long long))};
    roff += peers[i].recv_cnt;
  }
  const std::vector<std::vector<double>> in = comm_exchange_lists(out, s);
  std::vector<HaloPeerDev> dp(np);
  for (int i = 0; i < np; ++i) {
    const Peer& p = peers[i];
    const std::vector<double>& r = in[p.rank];
    if (r.size() != 4) throw Error(SHAKTI_ERR_COMM, "halo plan: neighbour relation is not symmetric");
    HaloPeerDev& d = dp[i];
    d.r_stage = p2p_peer(p.rank, (size_t)r[0]);
    d.r_slot_stride = (unsigned long long)r[1];
    d.r_recv_off = (int32_t)r[2];
    d.r_flag = reinterpret_cast<unsigned long long*>(p2p_peer(p.rank, (size_t)r[3]));
    d.send_off = p.send_off; d.send_cnt = p.send_cnt;
    d.recv_off = my_roff[i]; d.recv_cnt = p.recv_cnt; d.recv_begin = p.recv_begin;
  }
  if (np) dpeers.upload(dp);
}

template <class T>
void HaloPlan::exchange_p2p(const T* src, const int32_t* idx, T* dst, int width, cudaStream_t s) {
  ensure_p2p((int)sizeof(T) * width, s);   // collective when the staging must grow: every rank gets here, peers or not
  if (peers.empty()) return;
  const int64_t work = (int64_t)std::max(n_send, n_recv) * width;
  const int blocks = (int)std::max<int64_t>(1, std::min<int64_t>(16, (work + 2047) / 2048));
  SHAKTI_LAUNCH_PDL((halo_p2p_kernel<T>), blocks, 256, 0, s, (int)peers.size(), dpeers.p, p2p_local(stage_off), slot_stride,
                reinterpret_cast<const unsigned long long*>(p2p_local(flag_off)), ctr.p, src, idx, dst, width, g_p2p.err_dev);
}
template void HaloPlan::exchange_p2p<double>(const double*, const int32_t*, double*, int, cudaStream_t);
template void HaloPlan::exchange_p2p<float>(const float*, const int32_t*, float*, int, cudaStream_t);

__global__ void halo_pack_block_kernel(int32_t n, int width, const int32_t* __restrict__ idx, const double* __restrict__ v,
                                       double* __restrict__ out) {
  const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= (int64_t)n * width) return;
  const int32_t k = (int32_t)(t / width), w = (int32_t)(t - (int64_t)k * width);
  out[t] = v[(int64_t)idx[k] * width + w];
}

void HaloPlan::exchange_block(double* v, int width, cudaStream_t s) {
  if (!comm().active() || width <= 0) return;
  if (p2p) { exchange_p2p<double>(v, send_idx.p, v, width, s); return; }
  if (peers.empty()) return;
  if (send_buf.n < (size_t)n_send * width) send_buf.alloc((size_t)std::max(n_send, 1) * width);
  if (n_send)
    SHAKTI_LAUNCH(halo_pack_block_kernel, div_up((int64_t)n_send * width, 256), 256, 0, s, n_send, width, send_idx.p, v, send_buf.p);
  SHAKTI_NCCL(g_api.GroupStart());
  for (const auto& p : peers) {
    if (p.send_cnt)
      SHAKTI_NCCL(g_api.Send(send_buf.p + (size_t)p.send_off * width, (size_t)p.send_cnt * width, ncclFloat64, p.rank, g_nccl, s));
    if (p.recv_cnt)
      SHAKTI_NCCL(g_api.Recv(v + (size_t)p.recv_begin * width, (size_t)p.recv_cnt * width, ncclFloat64, p.rank, g_nccl, s));
  }
  SHAKTI_NCCL(g_api.GroupEnd());
}

void HaloPlan::exchange_packed(const double* sendbuf, double* recvbuf, int32_t ghost_base, int width, cudaStream_t s) {
  if (!comm().active() || width <= 0) return;
  if (p2p) { exchange_p2p<double>(sendbuf, nullptr, recvbuf - (int64_t)ghost_base * width, width, s); return; }
  if (peers.empty()) return;
  SHAKTI_NCCL(g_api.GroupStart());
  for (const auto& p : peers) {
    if (p.send_cnt)
      SHAKTI_NCCL(g_api.Send(sendbuf + (size_t)p.send_off * width, (size_t)p.send_cnt * width, ncclFloat64, p.rank, g_nccl, s));
    if (p.recv_cnt)
      SHAKTI_NCCL(g_api.Recv(recvbuf + (size_t)(p.recv_begin - ghost_base) * width, (size_t)p.recv_cnt * width, ncclFloat64,
                             p.rank, g_nccl, s));
  }
  SHAKTI_NCCL(g_api.GroupEnd());
}

__global__ void halo_gather_f32_kernel(int32_t n, const int32_t* __restrict__ idx, const float* __restrict__ v,
                                       float* __restrict__ out) {
  const int32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = v[idx[i]];
}

void HaloPlan::exchange(float* v, cudaStream_t s) {
  if (!comm().active()) return;
  if (p2p) { exchange_p2p<float>(v, send_idx.p, v, 1, s); return; }
  if (peers.empty()) return;
  if (send_buf_f.n < (size_t)n_send) send_buf_f.alloc((size_t)std::max(n_send, 1));
  if (n_send) SHAKTI_LAUNCH(halo_gather_f32_kernel, div_up(n_send, 256), 256, 0, s, n_send, send_idx.p, v, send_buf_f.p);
  SHAKTI_NCCL(g_api.GroupStart());
  for (const auto& p : peers) {
    if (p.send_cnt) SHAKTI_NCCL(g_api.Send(send_buf_f.p + p.send_off, (size_t)p.send_cnt, ncclFloat32, p.rank, g_nccl, s));
    if (p.recv_cnt) SHAKTI_NCCL(g_api.Recv(v + p.recv_begin, (size_t)p.recv_cnt, ncclFloat32, p.rank, g_nccl, s));
  }
  SHAKTI_NCCL(g_api.GroupEnd());
}

void HaloPlan::exchange(double* v, cudaStream_t s) {
  if (!comm().active()) return;
  if (p2p) { exchange_p2p<double>(v, send_idx.p, v, 1, s); return; }
  if (peers.empty()) return;
  if (n_send) launch_gather(n_send, send_idx.p, v, send_buf.p, s);
  SHAKTI_NCCL(g_api.GroupStart());
  for (const auto& p : peers) {
    if (p.send_cnt) SHAKTI_NCCL(g_api.Send(send_buf.p + p.send_off, (size_t)p.send_cnt, ncclFloat64, p.rank, g_nccl, s));
    if (p.recv_cnt) SHAKTI_NCCL(g_api.Recv(v + p.recv_begin, (size_t)p.recv_cnt, ncclFloat64, p.rank, g_nccl, s));
  }
  SHAKTI_NCCL(g_api.GroupEnd());
}

}  // namespace shakti
