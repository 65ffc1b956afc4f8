// comm.cu -- multi-GPU behind the C ABI: NCCL communicators, replicated nets, the fan-out + all-reduce of the
// sharded sampling paths.
//
// The reference's only entry is infer(im, pgm, query; evidence) (src/ProbabilisticGraphicalModels/inference.jl:48-49):
// one call answers the query with whatever the machine has.  Particles (src/Inference/likelihood.jl:34-54) and Gibbs
// chains are i.i.d., so GPU g of G takes the contiguous GLOBAL index range [first + n*g/G, first + n*(g+1)/G) of the
// job (Philox counters are global indices: every world size consumes the same random streams) and the only exchange
// step is ONE ncclAllReduce(sum, fp64) of the (cells + 2) doubles -- issued here, on the stream the kernel ran on, not
// by a Python / torch.distributed layer.  NCCL is dlopen'ed (libnccl.so.2) so that libbn_cuda.so loads, and every
// single-GPU entry point works, on a machine without it.
//
// Two process models share the code: one process driving all GPUs (ncclCommInitAll, one stream per device; what a
// Julia session uses) and one process per GPU (ncclCommInitRank with a unique id shipped out of band; torchrun / MPI).
#include "common.cuh"

#include <dlfcn.h>
#include <nccl.h>

#include <functional>

namespace bn {

// ------------------------------------------------------------------------------------------ NCCL (dlopen)
struct Nccl {
  void* h = nullptr;
  ncclResult_t (*GetVersion)(int*) = nullptr;
  ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
  ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
  ncclResult_t (*CommInitAll)(ncclComm_t*, int, const int*) = nullptr;
  ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
  ncclResult_t (*AllReduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*GroupStart)() = nullptr;
  ncclResult_t (*GroupEnd)() = nullptr;
  const char* (*GetErrorString)(ncclResult_t) = nullptr;
  bool ok = false;
  std::string path;
};

static Nccl& nccl() {
  static Nccl n;
  static std::once_flag once;
  std::call_once(once, [] {
    // an already-loaded libnccl.so.2 (e.g. the one a host framework brought along) is reused by the loader
    const char* cands[] = {getenv("BN_NCCL_PATH"), "libnccl.so.2", "/usr/lib/x86_64-linux-gnu/libnccl.so.2", "libnccl.so"};
    for (const char* c : cands) {
      if (!c) continue;
      n.h = dlopen(c, RTLD_NOW | RTLD_GLOBAL);
      if (n.h) { n.path = c; break; }
    }
    if (!n.h) return;
#define BN_SYM(field, name) *(void**)(&n.field) = dlsym(n.h, name); if (!n.field) return;
    BN_SYM(GetVersion, "ncclGetVersion")
    BN_SYM(GetUniqueId, "ncclGetUniqueId")
    BN_SYM(CommInitRank, "ncclCommInitRank")
    BN_SYM(CommInitAll, "ncclCommInitAll")
    BN_SYM(CommDestroy, "ncclCommDestroy")
    BN_SYM(AllReduce, "ncclAllReduce")
    BN_SYM(GroupStart, "ncclGroupStart")
    BN_SYM(GroupEnd, "ncclGroupEnd")
    BN_SYM(GetErrorString, "ncclGetErrorString")
#undef BN_SYM
    n.ok = true;
  });
  return n;
}

#define BN_NCCL_TRY(expr)                                                                          \
  do {                                                                                             \
    ncclResult_t _r = (expr);                                                                      \
    if (_r != ncclSuccess) {                                                                       \
      bn::set_error("%s failed: %s (%s:%d)", #expr, nccl().GetErrorString(_r), __FILE__, __LINE__); \
      return BN_NCCL_ERR;                                                                          \
    }                                                                                              \
  } while (0)

// ------------------------------------------------------------------------------------------ communicator
struct Comm {
  int world = 1;       // GPUs in the communicator
  int first_rank = 0;  // global rank of this process's first local GPU
  std::vector<int> devices;           // local devices
  std::vector<ncclComm_t> comms;      // one per local device
  std::vector<cudaStream_t> streams;  // one per local device; the FIRST local device runs on the caller's stream instead
  std::vector<double*> res;           // per local device: result staging for devices 1.. (grown on demand)
  std::vector<size_t> res_n;
  std::mutex mu;                      // one host thread at a time (NCCL's rule for a communicator)
  int n_local() const { return (int)devices.size(); }
};

static std::mutex g_comm_mu;
static std::unordered_map<uint64_t, Comm*> g_comms;
static uint64_t g_comm_next = 1;

static Comm* get_comm(bn_comm_t h) {
  std::lock_guard<std::mutex> lk(g_comm_mu);
  auto it = g_comms.find(h);
  return it == g_comms.end() ? nullptr : it->second;
}

static int32_t require_nccl() {
  Nccl& nc = nccl();
  BN_REQUIRE(nc.ok, BN_NCCL_ERR, "libnccl.so.2 could not be loaded (set BN_NCCL_PATH); multi-GPU entry points need NCCL");
  return BN_OK;
}

static void comm_free(Comm* c) {
  Nccl& nc = nccl();
  for (int d = 0; d < c->n_local(); ++d) {
    cudaSetDevice(c->devices[d]);
    if (d < (int)c->streams.size() && c->streams[d]) cudaStreamSynchronize(c->streams[d]);
    if (d < (int)c->comms.size() && c->comms[d] && nc.ok) nc.CommDestroy(c->comms[d]);
    if (d < (int)c->res.size() && c->res[d]) cudaFree(c->res[d]);
    if (d < (int)c->streams.size() && c->streams[d]) cudaStreamDestroy(c->streams[d]);
  }
  delete c;
}

static int32_t comm_finish_init(Comm* c) {
  const int nl = c->n_local();
  c->streams.assign(nl, nullptr);
  c->res.assign(nl, nullptr);
  c->res_n.assign(nl, 0);
  for (int d = 0; d < nl; ++d) {
    BN_CUDA_TRY(cudaSetDevice(c->devices[d]));
    ensure_pool(c->devices[d]);
    BN_CUDA_TRY(cudaStreamCreateWithFlags(&c->streams[d], cudaStreamNonBlocking));
  }
  BN_CUDA_TRY(cudaSetDevice(c->devices[0]));
  return BN_OK;
}

static void shard(uint64_t first, uint64_t n, int rank, int world, uint64_t* lo, uint64_t* cnt) {
  const unsigned __int128 a = (unsigned __int128)n * (unsigned)rank / (unsigned)world;
  const unsigned __int128 b = (unsigned __int128)n * (unsigned)(rank + 1) / (unsigned)world;
  *lo = first + (uint64_t)a;
  *cnt = (uint64_t)(b - a);
}

// Fan one sharded job out over the local GPUs of net->comm and all-reduce its fp64 result.
//   launch(replica, lo, cnt, out) enqueues the shard [lo, lo + cnt) of the job on `replica` (current device and
//   tl_stream are already set to that replica's) and leaves n_out doubles in `out` (zeroed by launch).
//   out_dev0: device buffer on the first local device, n_out doubles, on the caller's stream.
int32_t comm_fanout(Net* net, uint64_t first, uint64_t n, size_t n_out, double* out_dev0,
                    const std::function<int32_t(Net*, int, uint64_t, uint64_t, double*)>& launch) {
  Comm* c = net->comm;
  BN_TRY(require_nccl());
  Nccl& nc = nccl();
  std::lock_guard<std::mutex> lk(c->mu);
  const int nl = c->n_local();
  BN_REQUIRE((int)net->peers.size() == nl - 1, BN_BAD_ARG, "net / communicator mismatch");
  cudaStream_t s0 = tl_stream;
  std::vector<double*> outs(nl, nullptr);
  std::vector<cudaStream_t> ss(nl, nullptr);
  int32_t rc = BN_OK;
  for (int d = 0; d < nl && rc == BN_OK; ++d) {
    Net* r = d == 0 ? net : net->peers[d - 1];
    BN_CUDA_TRY(cudaSetDevice(r->device));
    ss[d] = d == 0 ? s0 : c->streams[d];
    if (d == 0) {
      outs[d] = out_dev0;
    } else {
      if (c->res_n[d] < n_out) {
        if (c->res[d]) { cudaStreamSynchronize(c->streams[d]); cudaFree(c->res[d]); c->res[d] = nullptr; }
        BN_CUDA_TRY(cudaMalloc((void**)&c->res[d], std::max<size_t>(n_out, 4096) * 8));
        c->res_n[d] = std::max<size_t>(n_out, 4096);
      }
      outs[d] = c->res[d];
    }
    uint64_t lo, cnt;
    shard(first, n, c->first_rank + d, c->world, &lo, &cnt);
    StreamScope sc(ss[d]);
    rc = launch(r, d, lo, cnt, outs[d]);
  }
  if (rc == BN_OK && c->world > 1) {
    // one collective for the whole job: (cells + 2) doubles, latency-only over NVLink / NVSwitch
    ncclResult_t r = nc.GroupStart();
    for (int d = 0; d < nl && r == ncclSuccess; ++d)
      r = nc.AllReduce(outs[d], outs[d], n_out, ncclDouble, ncclSum, c->comms[d], ss[d]);
    ncclResult_t r2 = nc.GroupEnd();
    if (r == ncclSuccess) r = r2;
    if (r != ncclSuccess) {
      set_error("ncclAllReduce failed: %s", nc.GetErrorString(r));
      rc = BN_NCCL_ERR;
    }
  }
  cudaSetDevice(net->device);
  return rc;
}

// Wait for the peers' streams (the caller's own stream is the caller's business).
int32_t comm_sync_peers(Net* net) {
  Comm* c = net->comm;
  for (int d = 1; d < c->n_local(); ++d) {
    BN_CUDA_TRY(cudaSetDevice(c->devices[d]));
    BN_CUDA_TRY(cudaStreamSynchronize(c->streams[d]));
  }
  BN_CUDA_TRY(cudaSetDevice(net->device));
  return BN_OK;
}

void comm_describe(const Net* net, int* world, int* first_rank, int* n_local) {
  *world = net->comm ? net->comm->world : 1;
  *first_rank = net->comm ? net->comm->first_rank : 0;
  *n_local = net->comm ? net->comm->n_local() : 1;
}

cudaStream_t comm_stream(const Net* net, int local_index) { return net->comm->streams[local_index]; }

void comm_shard_of(const Net* net, int local_index, uint64_t first, uint64_t n, uint64_t* lo, uint64_t* cnt) {
  shard(first, n, net->comm->first_rank + local_index, net->comm->world, lo, cnt);
}

}  // namespace bn

using namespace bn;

extern "C" {

int32_t bn_comm_shard(uint64_t first, uint64_t n, int32_t rank, int32_t world, uint64_t* out_first, uint64_t* out_n) {
  BN_REQUIRE(out_first && out_n, BN_BAD_ARG, "bn_comm_shard: NULL output");
  BN_REQUIRE(world >= 1 && rank >= 0 && rank < world, BN_BAD_ARG, "rank %d outside 0..%d", rank, world - 1);
  bn::shard(first, n, rank, world, out_first, out_n);
  return BN_OK;
}

int32_t bn_comm_init_all(const int32_t* devices, int32_t n_devices, bn_comm_t* out_comm) {
  BN_REQUIRE(out_comm && devices && n_devices >= 1, BN_BAD_ARG, "bn_comm_init_all: bad arguments");
  int ndev = 0;
  BN_CUDA_TRY(cudaGetDeviceCount(&ndev));
  for (int i = 0; i < n_devices; ++i) {
    BN_REQUIRE(devices[i] >= 0 && devices[i] < ndev, BN_BAD_ARG, "device %d out of range (%d visible)", devices[i], ndev);
    for (int j = 0; j < i; ++j) BN_REQUIRE(devices[j] != devices[i], BN_BAD_ARG, "device %d listed twice", devices[i]);
  }
  Comm* c = new Comm();
  c->world = n_devices;
  c->first_rank = 0;
  c->devices.assign(devices, devices + n_devices);
  c->comms.assign(n_devices, nullptr);
  if (n_devices > 1) {  // a one-GPU "communicator" needs no NCCL at all
    int32_t rc = require_nccl();
    if (rc != BN_OK) { delete c; return rc; }
    ncclResult_t r = nccl().CommInitAll(c->comms.data(), n_devices, c->devices.data());
    if (r != ncclSuccess) {
      set_error("ncclCommInitAll failed: %s", nccl().GetErrorString(r));
      delete c;
      return BN_NCCL_ERR;
    }
  }
  int32_t rc = comm_finish_init(c);
  if (rc != BN_OK) { comm_free(c); return rc; }
  std::lock_guard<std::mutex> lk(g_comm_mu);
  *out_comm = g_comm_next++;
  g_comms[*out_comm] = c;
  return BN_OK;
}

int32_t bn_comm_unique_id(uint8_t* out_id) {
  BN_REQUIRE(out_id, BN_BAD_ARG, "bn_comm_unique_id: NULL output");
  BN_TRY(require_nccl());
  static_assert(sizeof(ncclUniqueId) == BN_COMM_ID_BYTES, "ncclUniqueId size");
  ncclUniqueId id;
  BN_NCCL_TRY(nccl().GetUniqueId(&id));
  memcpy(out_id, &id, sizeof(id));
  return BN_OK;
}

int32_t bn_comm_init_rank(const uint8_t* id, int32_t rank, int32_t world, int32_t device, bn_comm_t* out_comm) {
  BN_REQUIRE(out_comm && id, BN_BAD_ARG, "bn_comm_init_rank: NULL argument");
  BN_REQUIRE(world >= 1 && rank >= 0 && rank < world, BN_BAD_ARG, "rank %d outside 0..%d", rank, world - 1);
  int ndev = 0;
  BN_CUDA_TRY(cudaGetDeviceCount(&ndev));
  BN_REQUIRE(device >= 0 && device < ndev, BN_BAD_ARG, "device %d out of range (%d visible)", device, ndev);
  BN_TRY(require_nccl());
  BN_CUDA_TRY(cudaSetDevice(device));
  Comm* c = new Comm();
  c->world = world;
  c->first_rank = rank;
  c->devices.assign(1, device);
  c->comms.assign(1, nullptr);
  ncclUniqueId nid;
  memcpy(&nid, id, sizeof(nid));
  ncclResult_t r = nccl().CommInitRank(&c->comms[0], world, nid, rank);
  if (r != ncclSuccess) {
    set_error("ncclCommInitRank failed: %s", nccl().GetErrorString(r));
    delete c;
    return BN_NCCL_ERR;
  }
  int32_t rc = comm_finish_init(c);
  if (rc != BN_OK) { comm_free(c); return rc; }
  std::lock_guard<std::mutex> lk(g_comm_mu);
  *out_comm = g_comm_next++;
  g_comms[*out_comm] = c;
  return BN_OK;
}

int32_t bn_comm_info(bn_comm_t h, int32_t* out_world, int32_t* out_first_rank, int32_t* out_n_local) {
  Comm* c = get_comm(h);
  BN_REQUIRE(c, BN_BAD_ARG, "bn_comm_info: unknown communicator");
  if (out_world) *out_world = c->world;
  if (out_first_rank) *out_first_rank = c->first_rank;
  if (out_n_local) *out_n_local = c->n_local();
  return BN_OK;
}

int32_t bn_comm_destroy(bn_comm_t h) {
  Comm* c = nullptr;
  {
    std::lock_guard<std::mutex> lk(g_comm_mu);
    auto it = g_comms.find(h);
    BN_REQUIRE(it != g_comms.end(), BN_BAD_ARG, "bn_comm_destroy: unknown communicator");
    c = it->second;
    g_comms.erase(it);
  }
  comm_free(c);
  return BN_OK;
}

static int32_t nccl_types(int32_t dtype, int32_t op, ncclDataType_t* t, ncclRedOp_t* o) {
  BN_REQUIRE(dtype == BN_DTYPE_F64 || dtype == BN_DTYPE_I64, BN_BAD_ARG, "bad dtype %d", dtype);
  BN_REQUIRE(op == BN_REDUCE_SUM || op == BN_REDUCE_MAX, BN_BAD_ARG, "bad reduction %d", op);
  *t = dtype == BN_DTYPE_F64 ? ncclDouble : ncclInt64;
  *o = op == BN_REDUCE_SUM ? ncclSum : ncclMax;
  return BN_OK;
}

int32_t bn_comm_allreduce_dev(bn_comm_t h, void* dev_inout, int64_t count, int32_t dtype, int32_t op) {
  Comm* c = get_comm(h);
  BN_REQUIRE(c, BN_BAD_ARG, "bn_comm_allreduce_dev: unknown communicator");
  BN_REQUIRE(dev_inout && count >= 0, BN_BAD_ARG, "bn_comm_allreduce_dev: bad buffer");
  BN_REQUIRE(c->n_local() == 1, BN_UNSUPPORTED, "bn_comm_allreduce_dev takes one buffer: one local GPU per process only");
  if (c->world == 1 || count == 0) return BN_OK;
  ncclDataType_t t;
  ncclRedOp_t o;
  BN_TRY(nccl_types(dtype, op, &t, &o));
  std::lock_guard<std::mutex> lk(c->mu);
  BN_CUDA_TRY(cudaSetDevice(c->devices[0]));
  BN_NCCL_TRY(nccl().AllReduce(dev_inout, dev_inout, (size_t)count, t, o, c->comms[0], tl_stream));
  return BN_OK;
}

int32_t bn_comm_allreduce(bn_comm_t h, void* host_inout, int64_t count, int32_t dtype, int32_t op) {
  Comm* c = get_comm(h);
  BN_REQUIRE(c, BN_BAD_ARG, "bn_comm_allreduce: unknown communicator");
  BN_REQUIRE(host_inout && count >= 0, BN_BAD_ARG, "bn_comm_allreduce: bad buffer");
  if (c->world == c->n_local() || count == 0) return BN_OK;  // one process: nothing to combine
  BN_REQUIRE(c->n_local() == 1, BN_UNSUPPORTED, "mixed process / device layouts are not supported");
  BN_CUDA_TRY(cudaSetDevice(c->devices[0]));
  TmpBuf<unsigned char> d;
  BN_CUDA_TRY(d.alloc((size_t)count * 8));
  BN_CUDA_TRY(cudaMemcpyAsync(d.p, host_inout, (size_t)count * 8, cudaMemcpyHostToDevice, tl_stream));
  BN_TRY(bn_comm_allreduce_dev(h, d.p, count, dtype, op));
  BN_CUDA_TRY(cudaMemcpyAsync(host_inout, d.p, (size_t)count * 8, cudaMemcpyDeviceToHost, tl_stream));
  BN_CUDA_TRY(cudaStreamSynchronize(tl_stream));
  return BN_OK;
}

int32_t bn_comm_barrier(bn_comm_t h) {
  Comm* c = get_comm(h);
  BN_REQUIRE(c, BN_BAD_ARG, "bn_comm_barrier: unknown communicator");
  if (c->world > 1) {
    BN_TRY(require_nccl());
    std::lock_guard<std::mutex> lk(c->mu);
    const int nl = c->n_local();
    std::vector<TmpBuf<double>> bufs(nl);
    for (int d = 0; d < nl; ++d) {
      BN_CUDA_TRY(cudaSetDevice(c->devices[d]));
      StreamScope sc(d == 0 ? tl_stream : c->streams[d]);
      BN_CUDA_TRY(bufs[d].alloc(1));
      BN_CUDA_TRY(cudaMemsetAsync(bufs[d].p, 0, 8, tl_stream));
    }
    ncclResult_t r = nccl().GroupStart();
    for (int d = 0; d < nl && r == ncclSuccess; ++d)
      r = nccl().AllReduce(bufs[d].p, bufs[d].p, 1, ncclDouble, ncclSum, c->comms[d], d == 0 ? tl_stream : c->streams[d]);
    ncclResult_t r2 = nccl().GroupEnd();
    if (r == ncclSuccess) r = r2;
    BN_REQUIRE(r == ncclSuccess, BN_NCCL_ERR, "barrier all-reduce failed: %s", nccl().GetErrorString(r));
    for (int d = nl - 1; d >= 0; --d) {
      BN_CUDA_TRY(cudaSetDevice(c->devices[d]));
      BN_CUDA_TRY(cudaStreamSynchronize(d == 0 ? tl_stream : c->streams[d]));
      bufs[d].release();
    }
  } else {
    BN_CUDA_TRY(cudaSetDevice(c->devices[0]));
    BN_CUDA_TRY(cudaStreamSynchronize(tl_stream));
  }
  return BN_OK;
}

int32_t bn_net_create_comm(int32_t n_nodes, const int32_t* card, const int32_t* par_ptr, const int32_t* par_idx,
                           const int64_t* cpt_off, const double* cpt, bn_comm_t comm, bn_net_t* out_net) {
  BN_REQUIRE(out_net, BN_BAD_ARG, "bn_net_create_comm: NULL argument");
  Comm* c = get_comm(comm);
  BN_REQUIRE(c, BN_BAD_ARG, "bn_net_create_comm: unknown communicator");
  Net* net = nullptr;
  BN_TRY(make_net(n_nodes, card, par_ptr, par_idx, cpt_off, cpt, c->devices[0], &net));
  net->comm = c;
  for (int d = 1; d < c->n_local(); ++d) {
    Net* peer = nullptr;
    StreamScope sc(c->streams[d]);
    int32_t rc = make_net(n_nodes, card, par_ptr, par_idx, cpt_off, cpt, c->devices[d], &peer);
    if (rc != BN_OK) {
      delete net;
      return rc;
    }
    peer->comm = c;
    net->peers.push_back(peer);
  }
  BN_CUDA_TRY(cudaSetDevice(net->device));
  *out_net = register_net(net);
  return BN_OK;
}

}  // extern "C"
