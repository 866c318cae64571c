// NCCL plumbing (see comm.h).  libnccl is dlopen'ed so the library loads on boxes without it.
#include "comm.h"
#include "device.h"

#include <dlfcn.h>

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
Comm& comm() { return g_comm; }

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
}

void comm_finalize() {
  if (g_nccl) g_api.CommDestroy(g_nccl);
  g_nccl = nullptr;
  g_comm = Comm();
}

void comm_allreduce_sum(double* dev, int count, cudaStream_t s) {
  if (!g_comm.active() || count == 0) return;
  SHAKTI_NCCL(g_api.AllReduce(dev, dev, (size_t)count, ncclFloat64, ncclSum, g_nccl, s));
}

void comm_allreduce_max(double* dev, int count, cudaStream_t s) {
  if (!g_comm.active() || count == 0) return;
  SHAKTI_NCCL(g_api.AllReduce(dev, dev, (size_t)count, ncclFloat64, ncclMax, g_nccl, s));
}

void comm_allgather(const double* send, double* recv, int count, cudaStream_t s) {
  if (count == 0) return;
  if (!g_comm.active()) {
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
  if (!g_comm.active()) return v;
  DevBuf<double>& b = scratch(1);
  SHAKTI_CUDA(cudaMemcpyAsync(b.p, &v, sizeof(double), cudaMemcpyHostToDevice, s));
  comm_allreduce_sum(b.p, 1, s);
  SHAKTI_CUDA(cudaMemcpyAsync(&v, b.p, sizeof(double), cudaMemcpyDeviceToHost, s));
  SHAKTI_CUDA(cudaStreamSynchronize(s));
  return v;
}
double comm_host_max(double v, cudaStream_t s) {
  if (!g_comm.active()) return v;
  DevBuf<double>& b = scratch(1);
  SHAKTI_CUDA(cudaMemcpyAsync(b.p, &v, sizeof(double), cudaMemcpyHostToDevice, s));
  comm_allreduce_max(b.p, 1, s);
  SHAKTI_CUDA(cudaMemcpyAsync(&v, b.p, sizeof(double), cudaMemcpyDeviceToHost, s));
  SHAKTI_CUDA(cudaStreamSynchronize(s));
  return v;
}
std::vector<double> comm_host_allgather(double v, cudaStream_t s) {
  const int nr = g_comm.nranks;
  std::vector<double> out(nr, v);
  if (!g_comm.active()) return out;
  DevBuf<double>& b = scratch(1 + nr);
  SHAKTI_CUDA(cudaMemcpyAsync(b.p, &v, sizeof(double), cudaMemcpyHostToDevice, s));
  comm_allgather(b.p, b.p + 1, 1, s);
  SHAKTI_CUDA(cudaMemcpyAsync(out.data(), b.p + 1, sizeof(double) * nr, cudaMemcpyDeviceToHost, s));
  SHAKTI_CUDA(cudaStreamSynchronize(s));
  return out;
}

std::vector<std::vector<double>> comm_exchange_lists(const std::vector<std::vector<double>>& out, cudaStream_t s) {
  const int nr = g_comm.nranks, me = g_comm.rank;
  std::vector<std::vector<double>> in(nr);
  if (!g_comm.active()) { in[0] = out[0]; return in; }
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
  send_idx_host = idx;
  if (n_send) {
    send_idx.upload(idx);
    send_buf.alloc((size_t)n_send);
  }
}

__global__ void halo_pack_block_kernel(int32_t n, int width, const int32_t* __restrict__ idx, const double* __restrict__ v,
                                       double* __restrict__ out) {
  const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= (int64_t)n * width) return;
  const int32_t k = (int32_t)(t / width), w = (int32_t)(t - (int64_t)k * width);
  out[t] = v[(int64_t)idx[k] * width + w];
}

void HaloPlan::exchange_block(double* v, int width, cudaStream_t s) {
  if (!g_comm.active() || peers.empty() || width <= 0) return;
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
  if (!g_comm.active() || peers.empty() || width <= 0) return;
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
  if (!g_comm.active() || peers.empty()) return;
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
  if (!g_comm.active() || peers.empty()) return;
  if (n_send) launch_gather(n_send, send_idx.p, v, send_buf.p, s);
  SHAKTI_NCCL(g_api.GroupStart());
  for (const auto& p : peers) {
    if (p.send_cnt) SHAKTI_NCCL(g_api.Send(send_buf.p + p.send_off, (size_t)p.send_cnt, ncclFloat64, p.rank, g_nccl, s));
    if (p.recv_cnt) SHAKTI_NCCL(g_api.Recv(v + p.recv_begin, (size_t)p.recv_cnt, ncclFloat64, p.rank, g_nccl, s));
  }
  SHAKTI_NCCL(g_api.GroupEnd());
}

}  // namespace shakti
