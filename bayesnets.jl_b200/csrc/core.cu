// core.cu -- handles, error state, network upload (the device side of device_layout.jl).
#include "common.cuh"

namespace bn {

static thread_local char tl_error[512] = "";
thread_local cudaStream_t tl_stream = nullptr;
std::atomic<uint64_t> g_launches{0};

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(tl_error, sizeof(tl_error), fmt, ap);
  va_end(ap);
}

static std::mutex g_mu;
static std::unordered_map<uint64_t, Net*> g_nets;
static std::unordered_map<uint64_t, Factor*> g_factors;
static uint64_t g_next = 1;
static std::unordered_map<int, DevProps> g_props;

const DevProps& dev_props(int device) {
  std::lock_guard<std::mutex> lk(g_mu);
  auto it = g_props.find(device);
  if (it != g_props.end()) return it->second;
  DevProps p;
  cudaDeviceGetAttribute(&p.sm_count, cudaDevAttrMultiProcessorCount, device);
  cudaDeviceGetAttribute(&p.max_smem_optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, device);
  cudaDeviceGetAttribute(&p.smem_per_sm, cudaDevAttrMaxSharedMemoryPerMultiprocessor, device);
  return g_props[device] = p;
}

void ensure_pool(int device) {
  static std::mutex mu;
  static std::unordered_map<int, bool> done;
  std::lock_guard<std::mutex> lk(mu);
  if (done[device]) return;
  cudaMemPool_t pool;
  if (cudaDeviceGetDefaultMemPool(&pool, device) == cudaSuccess) {
    uint64_t thr = ~0ull;
    cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &thr);
  }
  done[device] = true;
}

namespace {
struct BlockKey {
  int device;
  cudaStream_t s;
  size_t bytes;
  bool operator==(const BlockKey& o) const { return device == o.device && s == o.s && bytes == o.bytes; }
};
struct BlockKeyHash {
  size_t operator()(const BlockKey& k) const {
    return std::hash<size_t>()(k.bytes) ^ (std::hash<const void*>()((const void*)k.s) * 31u) ^ ((size_t)k.device << 56);
  }
};
std::mutex g_cache_mu;
std::unordered_map<BlockKey, std::vector<void*>, BlockKeyHash> g_cache;
size_t g_cache_bytes = 0;
size_t cache_cap() {
  static const size_t cap = (getenv("BN_POOL_CACHE_MB") ? (size_t)atoll(getenv("BN_POOL_CACHE_MB")) : (size_t)16384) << 20;
  return cap;
}
}  // namespace

size_t pool_trim() {
  std::unordered_map<BlockKey, std::vector<void*>, BlockKeyHash> take;
  size_t parked = 0;
  {
    std::lock_guard<std::mutex> lk(g_cache_mu);
    take.swap(g_cache);
    parked = g_cache_bytes;
    g_cache_bytes = 0;
  }
  for (auto& kv : take)  // cudaFree (synchronising) rather than cudaFreeAsync: the stream a block was parked under may be gone
    for (void* q : kv.second) cudaFree(q);
  return parked;
}

cudaError_t pool_alloc(void** p, size_t bytes, cudaStream_t s, int* device_out) {
  int device = 0;
  cudaError_t e = cudaGetDevice(&device);
  if (e != cudaSuccess) return e;
  *device_out = device;
  {
    std::lock_guard<std::mutex> lk(g_cache_mu);
    auto it = g_cache.find(BlockKey{device, s, bytes});
    if (it != g_cache.end() && !it->second.empty()) {
      *p = it->second.back();
      it->second.pop_back();
      g_cache_bytes -= bytes;
      return cudaSuccess;
    }
  }
  e = cudaMallocAsync(p, bytes, s);
  if (e == cudaErrorMemoryAllocation) {  // give the parked blocks back to the driver pool and retry once
    cudaGetLastError();
    pool_trim();
    e = cudaMallocAsync(p, bytes, s);
  }
  return e;
}

void pool_free(void* p, size_t bytes, cudaStream_t s, int device) {
  {
    std::lock_guard<std::mutex> lk(g_cache_mu);
    if (g_cache_bytes + bytes <= cache_cap()) {
      g_cache[BlockKey{device, s, bytes}].push_back(p);
      g_cache_bytes += bytes;
      return;
    }
  }
  cudaFreeAsync(p, s);
}

Net* get_net(bn_net_t h) {
  std::lock_guard<std::mutex> lk(g_mu);
  auto it = g_nets.find(h);
  return it == g_nets.end() ? nullptr : it->second;
}
Factor* get_factor(bn_factor_t h) {
  std::lock_guard<std::mutex> lk(g_mu);
  auto it = g_factors.find(h);
  return it == g_factors.end() ? nullptr : it->second;
}
Factor* take_factor(bn_factor_t h) {
  std::lock_guard<std::mutex> lk(g_mu);
  auto it = g_factors.find(h);
  if (it == g_factors.end()) return nullptr;
  Factor* f = it->second;
  g_factors.erase(it);
  return f;
}
bn_factor_t register_factor(Factor* f) {
  std::lock_guard<std::mutex> lk(g_mu);
  uint64_t h = g_next++;
  g_factors[h] = f;
  return h;
}

// children CSR, ascending child index (Graphs.outneighbors order, src/bayes_nets.jl:96-99)
void build_children(Net* net) {
  const int n = net->n;
  const int32_t n_par = net->par_ptr[n];
  net->child_ptr.assign(n + 1, 0);
  for (int i = 0; i < n; ++i)
    for (int k = net->par_ptr[i]; k < net->par_ptr[i + 1]; ++k) net->child_ptr[net->par_idx[k] + 1]++;
  for (int i = 0; i < n; ++i) net->child_ptr[i + 1] += net->child_ptr[i];
  net->child_idx.assign(n_par, 0);
  std::vector<int32_t> fill(n, 0);
  for (int i = 0; i < n; ++i)
    for (int k = net->par_ptr[i]; k < net->par_ptr[i + 1]; ++k) {
      int p = net->par_idx[k];
      net->child_idx[net->child_ptr[p] + fill[p]++] = i;
    }
}

int32_t new_factor(const std::vector<int32_t>& dims, const std::vector<int32_t>& card, int device,
                   Factor** out) {
  Factor* f = new Factor();
  f->device = device;
  f->dims = dims;
  f->card = card;
  f->len = 1;
  for (int32_t c : card) f->len *= c;
  cudaError_t e = cudaSetDevice(device);
  if (e == cudaSuccess) ensure_pool(device);
  if (e == cudaSuccess) e = f->data.alloc((size_t)f->len);
  if (e != cudaSuccess) {
    set_error("cudaMalloc of %lld doubles failed: %s", (long long)f->len, cudaGetErrorString(e));
    delete f;
    cudaGetLastError();
    return e == cudaErrorMemoryAllocation ? BN_OOM : BN_CUDA_ERR;
  }
  *out = f;
  return BN_OK;
}


Net::~Net() {
  spec_cache_clear(this);
  loopy_plan_free(loopy);
  if (gtab) gibbs_tables_free(gtab);
  for (StatsPlan* sp : stats_plan) if (sp) stats_plan_free(sp);
  for (Net* p : peers) {
    cudaSetDevice(p->device);
    delete p;
  }
  cudaSetDevice(device);
}

bn_net_t register_net(Net* net) {
  std::lock_guard<std::mutex> lk(g_mu);
  uint64_t h = g_next++;
  g_nets[h] = net;
  return h;
}

// Validate a flat network and upload it to `device` (the body of bn_net_create; bn_net_create_comm calls it once per
// local device of the communicator).
int32_t make_net(int32_t n_nodes, const int32_t* card, const int32_t* par_ptr, const int32_t* par_idx,
                 const int64_t* cpt_off, const double* cpt, int32_t device, Net** out) {
  BN_REQUIRE(out && card && par_ptr && cpt_off && cpt, BN_BAD_ARG, "bn_net_create: NULL argument");
  BN_REQUIRE(n_nodes > 0, BN_BAD_ARG, "bn_net_create: n_nodes must be > 0");
  BN_REQUIRE(par_ptr[0] == 0 && cpt_off[0] == 0, BN_BAD_ARG, "bn_net_create: CSR offsets must start at 0");
  int32_t n_par = par_ptr[n_nodes];
  BN_REQUIRE(n_par == 0 || par_idx, BN_BAD_ARG, "bn_net_create: par_idx is NULL");
  for (int i = 0; i < n_nodes; ++i) {
    // sampled states travel as uint8 with 0xFF reserved
    BN_REQUIRE(card[i] >= 1 && card[i] <= BN_MAX_CARD, BN_BAD_ARG, "node %d: cardinality %d outside 1..%d", i, card[i],
               BN_MAX_CARD);
    BN_REQUIRE(par_ptr[i + 1] >= par_ptr[i], BN_BAD_ARG, "par_ptr not monotone at node %d", i);
    int64_t q = 1;
    for (int k = par_ptr[i]; k < par_ptr[i + 1]; ++k) {
      int p = par_idx[k];
      // topological order is the contract (src/bayes_nets.jl:26-48)
      BN_REQUIRE(p >= 0 && p < i, BN_BAD_ARG, "node %d: parent %d breaks the topological order", i, p);
      for (int k2 = par_ptr[i]; k2 < k; ++k2)
        BN_REQUIRE(par_idx[k2] != p, BN_BAD_ARG, "node %d lists parent %d twice", i, p);
      q *= card[p];
    }
    BN_REQUIRE(cpt_off[i + 1] - cpt_off[i] == q * card[i], BN_BAD_ARG,
               "node %d: CPT has %lld entries, expected %lld", i, (long long)(cpt_off[i + 1] - cpt_off[i]),
               (long long)(q * card[i]));
  }
  int ndev = 0;
  BN_CUDA_TRY(cudaGetDeviceCount(&ndev));
  BN_REQUIRE(device >= 0 && device < ndev, BN_BAD_ARG, "device %d out of range (%d visible)", device, ndev);
  BN_CUDA_TRY(cudaSetDevice(device));
  ensure_pool(device);  // per-call temporaries of the host-buffer entry points come from the stream-ordered pool

  Net* net = new Net();
  net->device = device;
  net->n = n_nodes;
  net->card.assign(card, card + n_nodes);
  net->par_ptr.assign(par_ptr, par_ptr + n_nodes + 1);
  net->par_idx.assign(par_idx, par_idx + n_par);
  net->cpt_off.assign(cpt_off, cpt_off + n_nodes + 1);
  net->cpt.assign(cpt, cpt + cpt_off[n_nodes]);
  build_children(net);
  // only what kernels read lives in HBM: the fp64 CPTs (Gibbs, Factor(bn, node)) and the cardinalities; graph
  // structure is consumed by the host-side planners (sampling programs, code generators, elimination order).
  // The upload is complete on return (the stream must belong to `device`: comm.cu switches tl_stream per replica).
  cudaStream_t s = tl_stream;
  cudaError_t e = cudaSuccess;
  auto chk = [&](cudaError_t x) { if (e == cudaSuccess) e = x; };
  chk(net->d_card.upload(net->card.data(), net->card.size(), s));
  chk(net->d_cpt.upload(net->cpt.data(), net->cpt.size(), s));
  chk(cudaStreamSynchronize(s));
  if (e != cudaSuccess) {
    set_error("bn_net_create: upload failed: %s", cudaGetErrorString(e));
    delete net;
    cudaGetLastError();
    return e == cudaErrorMemoryAllocation ? BN_OOM : BN_CUDA_ERR;
  }
  *out = net;
  return BN_OK;
}

}  // namespace bn

using namespace bn;

extern "C" {

const char* bn_last_error(void) { return bn::tl_error; }
int32_t bn_version(void) { return 100; }

int32_t bn_device_count(int32_t* out_n) {
  BN_REQUIRE(out_n, BN_BAD_ARG, "out_n is NULL");
  int n = 0;
  BN_CUDA_TRY(cudaGetDeviceCount(&n));
  *out_n = n;
  return BN_OK;
}

int32_t bn_set_stream(void* cuda_stream) {
  bn::tl_stream = (cudaStream_t)cuda_stream;
  return BN_OK;
}

uint64_t bn_launch_count(void) { return bn::g_launches.load(); }

int32_t bn_synchronize(void) {
  BN_CUDA_TRY(cudaStreamSynchronize(bn::tl_stream));
  return BN_OK;
}

int32_t bn_pool_trim(uint64_t* out_bytes) {
  const size_t parked = bn::pool_trim();
  if (out_bytes) *out_bytes = (uint64_t)parked;
  return BN_OK;
}

int32_t bn_dev_alloc(int32_t device, uint64_t bytes, void** out_ptr) {
  BN_REQUIRE(out_ptr, BN_BAD_ARG, "bn_dev_alloc: out_ptr is NULL");
  BN_CUDA_TRY(cudaSetDevice(device));
  bn::ensure_pool(device);
  BN_CUDA_TRY(cudaMallocAsync(out_ptr, bytes ? bytes : 1, bn::tl_stream));
  return BN_OK;
}
int32_t bn_dev_free(int32_t device, void* ptr) {
  if (!ptr) return BN_OK;
  BN_CUDA_TRY(cudaSetDevice(device));
  BN_CUDA_TRY(cudaFreeAsync(ptr, bn::tl_stream));
  return BN_OK;
}
int32_t bn_dev_memset(void* ptr, int32_t value, uint64_t bytes) {
  BN_REQUIRE(ptr || !bytes, BN_BAD_ARG, "bn_dev_memset: NULL pointer");
  BN_CUDA_TRY(cudaMemsetAsync(ptr, value, bytes, bn::tl_stream));
  return BN_OK;
}
int32_t bn_dev_upload(void* dst_dev, const void* src, uint64_t bytes) {
  BN_REQUIRE((dst_dev && src) || !bytes, BN_BAD_ARG, "bn_dev_upload: NULL pointer");
  BN_CUDA_TRY(cudaMemcpyAsync(dst_dev, src, bytes, cudaMemcpyHostToDevice, bn::tl_stream));
  return BN_OK;
}
int32_t bn_dev_download(void* dst, const void* src_dev, uint64_t bytes) {
  BN_REQUIRE((dst && src_dev) || !bytes, BN_BAD_ARG, "bn_dev_download: NULL pointer");
  BN_CUDA_TRY(cudaMemcpyAsync(dst, src_dev, bytes, cudaMemcpyDeviceToHost, bn::tl_stream));
  BN_CUDA_TRY(cudaStreamSynchronize(bn::tl_stream));
  return BN_OK;
}

int32_t bn_net_create(int32_t n_nodes, const int32_t* card, const int32_t* par_ptr, const int32_t* par_idx,
                      const int64_t* cpt_off, const double* cpt, int32_t device, bn_net_t* out_net) {
  BN_REQUIRE(out_net, BN_BAD_ARG, "bn_net_create: NULL argument");
  Net* net = nullptr;
  BN_TRY(bn::make_net(n_nodes, card, par_ptr, par_idx, cpt_off, cpt, device, &net));
  *out_net = bn::register_net(net);
  return BN_OK;
}

int32_t bn_net_destroy(bn_net_t h) {
  Net* net = nullptr;
  {
    std::lock_guard<std::mutex> lk(bn::g_mu);
    auto it = bn::g_nets.find(h);
    BN_REQUIRE(it != bn::g_nets.end(), BN_BAD_ARG, "bn_net_destroy: unknown handle");
    net = it->second;
    bn::g_nets.erase(it);
  }
  cudaSetDevice(net->device);
  delete net;
  return BN_OK;
}

}  // extern "C"
