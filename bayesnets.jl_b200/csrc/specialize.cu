// specialize.cu -- network-specialised likelihood-weighting kernel (NVRTC, sm_100a).
//
// Replaces the particle loop of src/Inference/likelihood.jl:34-54 for one (network, evidence mask, query):
// the host emits straight-line CUDA for THAT network -- one block of code per node in topological order,
// particle states in registers, table offsets / parent strides as immediates, the row index as <= 3 IMAD,
// one LDS per node -- compiles it for sm_100a with NVRTC and caches the cubin in the net handle.
// Evidence VALUES stay run-time data (they only change table contents, which are uploaded per call).
//
// Why: the generic interpreter (sampling.cu) spends ~78 issue slots per node-sample, 3/4 of it decoding
// the program and moving u8 states through shared memory.  LW on a B200 is bound by issue slots
// (ALU + FMA pipes), not by HBM, so removing the interpreter is the roofline move.
//
// Same random-stream spec as the generic kernel and oracle/bn_oracle.c (DEVICE_SPEC): draw d of a particle
// is word d&3 of Philox4x32-10(counter = (pid_lo, pid_hi, d>>2, 0), key = seed), d = index of the node
// among the FREE nodes in topological order; state = #{k : (r >> 1) >= T_k}.  Because draw indices are
// static, nodes that can influence neither the query nor the weight (not an ancestor of query or evidence:
// "barren" nodes) may be skipped without changing any count bit.  prune = 0 keeps every node alive (each
// state feeds a checksum the compiler cannot fold), which is what bench.py's headline number uses.
#include "common.cuh"

#include <cuda.h>
#include <dlfcn.h>
#include <nvrtc.h>

#include <sys/stat.h>
#include <unistd.h>

#include <condition_variable>
#include <cstring>
#include <functional>
#include <memory>
#include <sstream>
#include <thread>

namespace bn {

// ------------------------------------------------------------------------------------------ NVRTC (dlopen)
struct Nvrtc {
  void* h = nullptr;
  nvrtcResult (*CreateProgram)(nvrtcProgram*, const char*, const char*, int, const char* const*, const char* const*);
  nvrtcResult (*CompileProgram)(nvrtcProgram, int, const char* const*);
  nvrtcResult (*GetCUBINSize)(nvrtcProgram, size_t*);
  nvrtcResult (*GetCUBIN)(nvrtcProgram, char*);
  nvrtcResult (*GetProgramLogSize)(nvrtcProgram, size_t*);
  nvrtcResult (*GetProgramLog)(nvrtcProgram, char*);
  nvrtcResult (*DestroyProgram)(nvrtcProgram*);
  const char* (*GetErrorString)(nvrtcResult);
  nvrtcResult (*Version)(int*, int*) = nullptr;
  bool ok = false;
};

static Nvrtc& nvrtc() {
  static Nvrtc n;
  static std::once_flag once;
  std::call_once(once, [] {
    const char* cands[] = {getenv("BN_NVRTC_PATH"), "libnvrtc.so.12", "/usr/local/cuda/lib64/libnvrtc.so.12",
                           "/usr/local/cuda/lib64/libnvrtc.so", "libnvrtc.so"};
    for (const char* c : cands) {
      if (!c) continue;
      n.h = dlopen(c, RTLD_NOW | RTLD_GLOBAL);
      if (n.h) break;
    }
    if (!n.h) return;
#define BN_SYM(field, name) *(void**)(&n.field) = dlsym(n.h, name); if (!n.field) return;
    BN_SYM(CreateProgram, "nvrtcCreateProgram")
    BN_SYM(CompileProgram, "nvrtcCompileProgram")
    BN_SYM(GetCUBINSize, "nvrtcGetCUBINSize")
    BN_SYM(GetCUBIN, "nvrtcGetCUBIN")
    BN_SYM(GetProgramLogSize, "nvrtcGetProgramLogSize")
    BN_SYM(GetProgramLog, "nvrtcGetProgramLog")
    BN_SYM(DestroyProgram, "nvrtcDestroyProgram")
    BN_SYM(GetErrorString, "nvrtcGetErrorString")
    BN_SYM(Version, "nvrtcVersion")
#undef BN_SYM
    n.ok = true;
  });
  return n;
}

// ------------------------------------------------------------------------------------------ driver API
struct Driver {
  CUresult (*ModuleLoadData)(CUmodule*, const void*) = nullptr;
  CUresult (*ModuleUnload)(CUmodule) = nullptr;
  CUresult (*ModuleGetFunction)(CUfunction*, CUmodule, const char*) = nullptr;
  CUresult (*FuncSetAttribute)(CUfunction, CUfunction_attribute, int) = nullptr;
  CUresult (*FuncGetAttribute)(int*, CUfunction_attribute, CUfunction) = nullptr;
  CUresult (*OccupancyMaxActiveBlocksPerMultiprocessor)(int*, CUfunction, int, size_t) = nullptr;
  CUresult (*LaunchKernel)(CUfunction, unsigned, unsigned, unsigned, unsigned, unsigned, unsigned, unsigned, CUstream,
                           void**, void**) = nullptr;
  CUresult (*GetErrorString)(CUresult, const char**) = nullptr;
  bool ok = false;
};

static Driver& driver() {
  static Driver d;
  static std::once_flag once;
  std::call_once(once, [] {
    auto get = [](const char* name, void** fn) {
      cudaDriverEntryPointQueryResult q;
      return cudaGetDriverEntryPoint(name, fn, cudaEnableDefault, &q) == cudaSuccess && *fn;
    };
    d.ok = get("cuModuleLoadData", (void**)&d.ModuleLoadData) && get("cuModuleUnload", (void**)&d.ModuleUnload) &&
           get("cuModuleGetFunction", (void**)&d.ModuleGetFunction) &&
           get("cuFuncSetAttribute", (void**)&d.FuncSetAttribute) &&
           get("cuFuncGetAttribute", (void**)&d.FuncGetAttribute) &&
           get("cuOccupancyMaxActiveBlocksPerMultiprocessor", (void**)&d.OccupancyMaxActiveBlocksPerMultiprocessor) &&
           get("cuLaunchKernel", (void**)&d.LaunchKernel) && get("cuGetErrorString", (void**)&d.GetErrorString);
    cudaGetLastError();
  });
  return d;
}

// ------------------------------------------------------------------------------------------ code generation
constexpr int SPEC_MAX_REG_CELLS = 16;  // query tables up to this size accumulate in registers

// 512 x 1 (<= 128 registers) measured best on B200 for the 100-node case (profiles/r1b_lw_spec_sweep.txt);
// the kernel is bound by the fmaheavy (IMAD.WIDE of Philox) and ALU pipes, not by occupancy.
struct SpecShape {
  int threads = 512;
  int min_blocks = 1;
};

static void live_nodes(const Net& net, const SampleProgram& prg, const int32_t* query, uint32_t n_query, bool prune,
                       std::vector<char>* live) {
  const int n = net.n;
  live->assign(n, prune ? 0 : 1);
  if (!prune) return;
  // ancestors (through FREE parents only: evidence parents are folded into the tables) of query and evidence nodes
  std::vector<int> stack;
  for (uint32_t j = 0; j < n_query; ++j) stack.push_back(query[j]);
  for (int i = 0; i < n; ++i) if (prg.plan[i].is_evidence) stack.push_back(i);
  while (!stack.empty()) {
    int i = stack.back();
    stack.pop_back();
    if ((*live)[i]) continue;
    (*live)[i] = 1;
    for (auto& pr : prg.plan[i].parents) stack.push_back(pr.first);
  }
}

static const char* kPrelude = R"BNSRC(
typedef unsigned int u32;
typedef unsigned long long u64;

__device__ __forceinline__ u32 smem_u32(const void* p) { return (u32)__cvta_generic_to_shared(p); }

// Philox4x32-10 (Salmon et al. SC'11); identical to common.cuh / oracle/bn_oracle.c
__device__ __forceinline__ void px(u32 c0, u32 c1, u32 c2, u32 c3, u32 k0, u32 k1, u32& o0, u32& o1, u32& o2, u32& o3) {
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    const u32 hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
    const u32 hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
    const u32 n0 = hi1 ^ c1 ^ k0, n2 = hi0 ^ c3 ^ k1;
    c0 = n0; c1 = lo1; c2 = n2; c3 = lo0;
    k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
  }
  o0 = c0; o1 = c1; o2 = c2; o3 = c3;
}

// one-shot TMA bulk copy global -> shared (SASS: UBLKCP), completion on an mbarrier
__device__ __forceinline__ void stage_tables(void* dst, const void* src, u32 bytes, u64* bar) {
  const u32 b = smem_u32(bar);
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(b));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(b), "r"(bytes) : "memory");
    u32 done = 0;
    while (done < bytes) {
      const u32 chunk = bytes - done < 65536u ? bytes - done : 65536u;
      asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                       smem_u32((const char*)dst + done)),
                   "l"((const char*)src + done), "r"(chunk), "r"(b)
                   : "memory");
      done += chunk;
    }
  }
  u32 ok = 0;
  while (!ok) {
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(ok) : "r"(b) : "memory");
  }
}
)BNSRC";

enum SpecMode { SPEC_INFER = 0, SPEC_TABLE = 1, SPEC_REJECT = 2 };

// Emit the kernel for this (net structure, evidence mask, query | table mode).
//   SPEC_INFER : weighted counts of the query cell (bn_lw_infer)
//   SPEC_TABLE : one row of states (+ weight) per particle; evidence columns are clamped (bn_sample ancestral/weighted)
//   SPEC_REJECT: ancestral rows retried (Philox counter word 3 = attempt << 8) until consistent with the evidence
static std::string gen_source(SpecMode mode, const Net& net, const SampleProgram& prg, const std::vector<char>& ev_mask,
                              const int32_t* query, uint32_t n_query, const std::vector<int64_t>& qstride,
                              uint32_t ncell, bool prune, const SpecShape& shp, uint32_t* n_live_out) {
  const int n = net.n;
  const bool infer = mode == SPEC_INFER;
  std::vector<char> live;
  live_nodes(net, prg, query, n_query, prune && infer, &live);
  const bool reg_bins = ncell <= (uint32_t)SPEC_MAX_REG_CELLS;
  std::ostringstream o;
  o << "// generated by libbn_cuda (specialize.cu): " << n << " nodes, " << prg.n_free << " free, mode " << (int)mode;
  if (infer) o << ", query cells " << ncell << (prune ? ", barren nodes pruned" : ", all nodes kept alive");
  o << "\n" << kPrelude;
  o << "#define NT " << shp.threads << "\n#define TAB_WORDS " << prg.table_words << "u\n#define NCELL " << (infer ? ncell : 0u) << "u\n";
  o << "extern \"C\" __global__ void __launch_bounds__(NT, " << shp.min_blocks << ") lw_spec(const u32* __restrict__ tab_g, "
       "u64 first, u64 n, u32 k0, u32 k1, u32 never, double* __restrict__ out, const signed char* __restrict__ ev, "
       "unsigned char* __restrict__ out_states, double* __restrict__ out_w, unsigned char* __restrict__ out_ok, u32 max_att) {\n";
  o << "  extern __shared__ __align__(16) unsigned char smem_raw[];\n";
  o << "  __shared__ __align__(8) u64 bar;\n";
  o << "  const u32* __restrict__ tab = reinterpret_cast<const u32*>(smem_raw);\n";
  o << "  const u32 tid = threadIdx.x;\n";
  o << "  stage_tables(smem_raw, tab_g, TAB_WORDS * 4u, &bar);\n";
  if (infer) {
    o << "  double* bins = reinterpret_cast<double*>(smem_raw + TAB_WORDS * 4u);\n";
    o << "  for (u32 c = tid; c < NCELL + 2u; c += NT) bins[c] = 0.0;\n";
  }
  o << "  __syncthreads();\n";
  if (infer) {
    o << "  double sw = 0.0, sw2 = 0.0;\n";
    if (reg_bins)
      for (uint32_t c = 0; c < ncell; ++c) o << "  double a" << c << " = 0.0;\n";
  }
  // states that no child, query cell or weight reads (only matters when every node is kept alive)
  std::vector<char> has_reader(n, 0);
  for (int i = 0; i < n; ++i)
    for (auto& pr : prg.plan[i].parents) has_reader[pr.first] = 1;
  for (uint32_t j = 0; j < n_query; ++j) has_reader[query[j]] = 1;

  o << "  for (u64 row = (u64)blockIdx.x * NT + tid; row < n; row += (u64)gridDim.x * NT) {\n";
  o << "    const u64 pid = first + row;\n";
  o << "    const u32 c0 = (u32)pid, c1 = (u32)(pid >> 32);\n";
  o << "    double w = 1.0;\n";
  if (infer && !prune) o << "    u32 chk = 0u;\n";
  if (mode == SPEC_REJECT) {
    for (int i = 0; i < n; ++i) o << "    u32 s" << i << " = 0u;\n";
    o << "    u32 ok = 0u;\n";
    o << "    for (u32 att = 0u; att < max_att && !ok; ++att) {\n";
    o << "    const u32 c3 = att << 8;\n";
  }
  const char* c3 = mode == SPEC_REJECT ? "c3" : "0u";
  const char* decl = mode == SPEC_REJECT ? "" : "u32 ";
  uint32_t d = 0, n_live = 0;
  int cur_blk = -1;
  for (int i = 0; i < n; ++i) {
    const NodePlan& np = prg.plan[i];
    const uint32_t my_d = d;
    if (!np.is_evidence) ++d;
    if (!live[i]) continue;
    ++n_live;
    std::ostringstream addr;  // row address (words)
    addr << np.toff << "u";
    for (auto& pr : np.parents) addr << " + s" << pr.first << " * " << (pr.second * np.row_words) << "u";
    if (np.is_evidence) {
      o << "    w *= *reinterpret_cast<const double*>(tab + (" << addr.str() << "));  // node " << i << " (evidence)\n";
      continue;
    }
    const int blk = (int)(my_d >> 2);
    if (blk != cur_blk) {
      cur_blk = blk;
      o << "    u32 r" << blk << "_0, r" << blk << "_1, r" << blk << "_2, r" << blk << "_3;\n";
      o << "    px(c0, c1, " << blk << "u, " << c3 << ", k0, k1, r" << blk << "_0, r" << blk << "_1, r" << blk << "_2, r"
        << blk << "_3);\n";
    }
    std::ostringstream rw;
    rw << "r" << blk << "_" << (my_d & 3u);
    const int nth = np.card - 1;  // real thresholds (row padding is never compared)
    o << "    // node " << i << ": card " << np.card << ", " << np.parents.size() << " free parents\n";
    if (np.card <= 1) {
      o << "    " << decl << "s" << i << " = 0u;\n";
    } else {
      // sign-bit form: rb = (r >> 1) | 2^31 (one funnel shift); rb - T has bit 31 set iff (r >> 1) >= T for every
      // T in [0, 2^31]; two ALU ops per threshold (IADD + LEA.HI / SHF), no predicates, no moves
      o << "    " << decl << "s" << i << ";\n";
      o << "    { const u32 rb = __funnelshift_r(" << rw.str() << ", 1u, 1); const u32* tp = tab + (" << addr.str() << ");\n";
      if (nth == 1) o << "      s" << i << " = (rb - tp[0]) >> 31;\n";
      else if (nth == 2) o << "      { const uint2 t = *reinterpret_cast<const uint2*>(tp); s" << i << " = ((rb - t.x) >> 31) + ((rb - t.y) >> 31); }\n";
      else {
        o << "      s" << i << " = 0u;\n";
        for (int k = 0; k < nth; k += 4) {
          o << "      { const uint4 t = *reinterpret_cast<const uint4*>(tp + " << k << "); s" << i << " += ((rb - t.x) >> 31)";
          if (k + 1 < nth) o << " + ((rb - t.y) >> 31)";
          if (k + 2 < nth) o << " + ((rb - t.z) >> 31)";
          if (k + 3 < nth) o << " + ((rb - t.w) >> 31)";
          o << "; }\n";
        }
      }
      o << "    }\n";
    }
    if (infer && !prune && !has_reader[i]) o << "    chk += s" << i << ";  // nothing else reads this state\n";
  }
  if (mode == SPEC_REJECT) {
    // consistent(a, evidence): src/ProbabilisticGraphicalModels/assignments.jl:13-22
    o << "    ok = 1u;\n";
    for (int i = 0; i < n; ++i)
      if (ev_mask[i]) o << "    ok &= (u32)(s" << i << " == (u32)ev[" << i << "]);\n";
    o << "    }\n";
  }
  if (infer) {
    o << "    const u32 cell = 0u";
    for (uint32_t j = 0; j < n_query; ++j) o << " + s" << query[j] << " * " << (uint32_t)qstride[j] << "u";
    o << ";\n";
    if (!prune) o << "    if (chk == never) w = 0.0;  // keeps every state live; never true (host passes 0xFFFFFFFF)\n";
    if (reg_bins) {
      for (uint32_t c = 0; c < ncell; ++c) o << "    a" << c << " += (cell == " << c << "u) ? w : 0.0;\n";
    } else {
      o << "    atomicAdd(&bins[cell], w);\n";
    }
    o << "    sw += w; sw2 += w * w;\n";
  } else {
    // one DataFrame column per node: node-major uint8, 32 consecutive bytes per warp store
    for (int i = 0; i < n; ++i) {
      o << "    out_states[" << i << "ull * n + row] = (unsigned char)(";
      if (mode == SPEC_REJECT) o << "ok ? s" << i << " : 0xFFu";
      else if (prg.plan[i].is_evidence) o << "ev[" << i << "]";
      else o << "s" << i;
      o << ");\n";
    }
    o << "    if (out_w) out_w[row] = w;\n";
    o << "    if (out_ok) out_ok[row] = (unsigned char)" << (mode == SPEC_REJECT ? "ok" : "1u") << ";\n";
  }
  o << "  }\n";
  if (infer) {
    // reductions: warp shuffle, one smem atomic per warp, then one global atomic per (CTA, cell)
    o << "#pragma unroll\n  for (int of = 16; of > 0; of >>= 1) {\n";
    o << "    sw += __shfl_xor_sync(0xffffffffu, sw, of); sw2 += __shfl_xor_sync(0xffffffffu, sw2, of);\n";
    if (reg_bins)
      for (uint32_t c = 0; c < ncell; ++c) o << "    a" << c << " += __shfl_xor_sync(0xffffffffu, a" << c << ", of);\n";
    o << "  }\n";
    o << "  if ((tid & 31u) == 0u) {\n    atomicAdd(&bins[NCELL], sw); atomicAdd(&bins[NCELL + 1u], sw2);\n";
    if (reg_bins)
      for (uint32_t c = 0; c < ncell; ++c) o << "    if (a" << c << " != 0.0) atomicAdd(&bins[" << c << "], a" << c << ");\n";
    o << "  }\n  __syncthreads();\n";
    o << "  for (u32 c = tid; c < NCELL + 2u; c += NT) if (bins[c] != 0.0) atomicAdd(&out[c], bins[c]);\n";
  }
  o << "}\n";
  if (n_live_out) *n_live_out = n_live;
  return o.str();
}

static int32_t nvrtc_compile(const std::string& src, std::vector<char>* cubin, std::string* log) {
  Nvrtc& nv = nvrtc();
  BN_REQUIRE(nv.ok, BN_UNSUPPORTED, "libnvrtc.so.12 could not be loaded (set BN_NVRTC_PATH)");
  nvrtcProgram prog;
  nvrtcResult r = nv.CreateProgram(&prog, src.c_str(), "lw_spec.cu", 0, nullptr, nullptr);
  BN_REQUIRE(r == NVRTC_SUCCESS, BN_CUDA_ERR, "nvrtcCreateProgram: %s", nv.GetErrorString(r));
  const char* opts[] = {"--gpu-architecture=sm_100a", "-lineinfo", "--std=c++17", "-fmad=false"};
  r = nv.CompileProgram(prog, 4, opts);
  size_t ls = 0;
  nv.GetProgramLogSize(prog, &ls);
  if (ls > 1 && log) { log->resize(ls); nv.GetProgramLog(prog, &(*log)[0]); }
  if (r != NVRTC_SUCCESS) {
    set_error("nvrtcCompileProgram: %s: %.300s", nv.GetErrorString(r), log ? log->c_str() : "");
    nv.DestroyProgram(&prog);
    return BN_CUDA_ERR;
  }
  size_t cs = 0;
  nv.GetCUBINSize(prog, &cs);
  cubin->resize(cs);
  nv.GetCUBIN(prog, cubin->data());
  nv.DestroyProgram(&prog);
  return BN_OK;
}
static const char* kToolchainTag = "nvrtc sm_100a -lineinfo c++17 fmad=false v2";

// ------------------------------------------------------------------------------------------ cubin cache
// Generated source -> cubin, three levels:
//   1. process-wide table keyed by the generated source (any net handle / device with the same structure, evidence
//      mask and query reuses it);
//   2. an on-disk cache (BN_CUBIN_CACHE, default $XDG_CACHE_HOME/bn_cuda or ~/.cache/bn_cuda; "off" disables): one file
//      per source, named by a 128-bit hash of (toolchain tag, NVRTC version, source) and carrying the source itself, which
//      is compared byte for byte on load -- a second process finds its kernels without paying NVRTC (~10 ms per node);
//   3. NVRTC, either blocking (explicit "specialised" mode) or on a background host thread while the interpreter kernel
//      keeps serving calls ("auto" mode): the first-ever call of a process never waits for a compile.
struct CubinEntry {
  enum State { COMPILING, READY, FAILED } state = COMPILING;
  std::vector<char> cubin;
  std::string log;
  bool from_disk = false;
};
static std::mutex g_cubin_mu;
static std::condition_variable g_cubin_cv;
static std::unordered_map<std::string, std::shared_ptr<CubinEntry>> g_cubins;
static std::atomic<uint64_t> g_spec_compiles{0}, g_disk_hits{0}, g_bg_compiles{0};

struct BgThreads {  // background compiles are joined before the library is torn down
  std::mutex mu;
  std::vector<std::thread> th;
  ~BgThreads() {
    std::lock_guard<std::mutex> lk(mu);
    for (auto& t : th) if (t.joinable()) t.join();
  }
};
static BgThreads& bg_threads() { static BgThreads b; return b; }

static uint64_t fnv1a(const std::string& s, uint64_t h) {
  for (unsigned char c : s) { h ^= c; h *= 0x100000001b3ull; }
  return h;
}

static std::string cache_dir() {
  const char* e = getenv("BN_CUBIN_CACHE");
  if (e && (!strcmp(e, "off") || !strcmp(e, "0") || !*e)) return "";
  std::string d;
  if (e) d = e;
  else if (const char* x = getenv("XDG_CACHE_HOME")) d = std::string(x) + "/bn_cuda";
  else if (const char* h = getenv("HOME")) d = std::string(h) + "/.cache/bn_cuda";
  else return "";
  // mkdir -p (two levels are enough for the defaults)
  for (size_t i = 1; i <= d.size(); ++i)
    if (i == d.size() || d[i] == '/') ::mkdir(d.substr(0, i).c_str(), 0755);
  return d;
}

static std::string cache_file(const std::string& src) {
  const std::string dir = cache_dir();
  if (dir.empty()) return "";
  int major = 0, minor = 0;
  if (nvrtc().Version) nvrtc().Version(&major, &minor);
  const std::string tag = std::string(kToolchainTag) + " " + std::to_string(major) + "." + std::to_string(minor);
  const uint64_t h1 = fnv1a(src, fnv1a(tag, 0xcbf29ce484222325ull)), h2 = fnv1a(src, fnv1a(tag, 0x84222325cbf29ce4ull));
  char name[64];
  snprintf(name, sizeof(name), "/%016llx%016llx.cubin", (unsigned long long)h1, (unsigned long long)h2);
  return dir + name;
}

static bool disk_load(const std::string& src, std::vector<char>* cubin) {
  const std::string f = cache_file(src);
  if (f.empty()) return false;
  FILE* fp = fopen(f.c_str(), "rb");
  if (!fp) return false;
  bool ok = false;
  uint64_t hdr[3];  // magic, source bytes, cubin bytes
  if (fread(hdr, 8, 3, fp) == 3 && hdr[0] == 0x314e4942554342ull && hdr[1] == src.size() && hdr[2] < (1ull << 31)) {
    std::string stored(hdr[1], '\0');
    cubin->resize(hdr[2]);
    ok = fread(&stored[0], 1, hdr[1], fp) == hdr[1] && fread(cubin->data(), 1, hdr[2], fp) == hdr[2] && stored == src;
  }
  fclose(fp);
  return ok;
}

static void disk_store(const std::string& src, const std::vector<char>& cubin) {
  const std::string f = cache_file(src);
  if (f.empty()) return;
  const std::string tmp = f + ".tmp" + std::to_string((long long)getpid()) + "_" + std::to_string((long long)(uintptr_t)&cubin);
  FILE* fp = fopen(tmp.c_str(), "wb");
  if (!fp) return;
  const uint64_t hdr[3] = {0x314e4942554342ull, src.size(), cubin.size()};
  const bool ok = fwrite(hdr, 8, 3, fp) == 3 && fwrite(src.data(), 1, src.size(), fp) == src.size() &&
                  fwrite(cubin.data(), 1, cubin.size(), fp) == cubin.size();
  fclose(fp);
  if (ok) ::rename(tmp.c_str(), f.c_str());  // atomic: readers see the old file or the complete new one
  else ::remove(tmp.c_str());
}

static void compile_into(const std::string& src, const std::shared_ptr<CubinEntry>& e) {
  std::vector<char> cubin;
  std::string log;
  g_spec_compiles.fetch_add(1);
  const int32_t rc = nvrtc_compile(src, &cubin, &log);
  if (rc == BN_OK) disk_store(src, cubin);
  std::lock_guard<std::mutex> lk(g_cubin_mu);
  e->log = log;
  if (rc == BN_OK) { e->cubin.swap(cubin); e->state = CubinEntry::READY; }
  else e->state = CubinEntry::FAILED;
  g_cubin_cv.notify_all();
}

// wait = true : returns a READY or FAILED entry (compiling in this thread if nobody else is).
// wait = false: returns the entry in whatever state it is; starts a background compile when there is none.
static std::shared_ptr<CubinEntry> get_cubin(const std::string& src, bool wait) {
  std::shared_ptr<CubinEntry> e;
  bool mine = false;
  {
    std::unique_lock<std::mutex> lk(g_cubin_mu);
    auto it = g_cubins.find(src);
    if (it != g_cubins.end()) {
      e = it->second;
      if (wait) g_cubin_cv.wait(lk, [&] { return e->state != CubinEntry::COMPILING; });
      return e;
    }
    e = std::make_shared<CubinEntry>();
    g_cubins[src] = e;
    mine = true;
  }
  if (mine) {
    std::vector<char> cubin;
    if (disk_load(src, &cubin)) {
      g_disk_hits.fetch_add(1);
      std::lock_guard<std::mutex> lk(g_cubin_mu);
      e->cubin.swap(cubin);
      e->from_disk = true;
      e->state = CubinEntry::READY;
      g_cubin_cv.notify_all();
      return e;
    }
    if (wait) {
      compile_into(src, e);
    } else {
      g_bg_compiles.fetch_add(1);
      BgThreads& b = bg_threads();
      std::lock_guard<std::mutex> lk(b.mu);
      b.th.emplace_back([src, e] { compile_into(src, e); });
    }
  }
  return e;
}

// ------------------------------------------------------------------------------------------ kernel cache
struct SpecKernel {
  CUmodule mod = nullptr;
  CUfunction fn = nullptr;
  int threads = 0, blocks_per_sm = 0, regs = 0;
  size_t smem = 0;
  uint32_t table_words = 0, n_live = 0;
};

// Loaded kernels are owned by a process-wide table keyed by (device, generated source): re-uploading the same
// network (a new bn_net_t), another network with the same structure, or the replicas of a multi-GPU net reuse the
// cubin instead of paying NVRTC again.  Net::spec_cache is a non-owning per-handle shortcut (guarded by Net::mu).
static std::mutex g_spec_mu;
static std::unordered_map<std::string, std::unique_ptr<SpecKernel>> g_spec;

void spec_cache_clear(Net* net) { net->spec_cache.clear(); }

static int32_t cu_fail(const char* what, CUresult r) {
  const char* s = nullptr;
  driver().GetErrorString(r, &s);
  set_error("%s failed: %s", what, s ? s : "?");
  return BN_CUDA_ERR;
}

thread_local int tl_lw_mode = 0;   // 0 auto, 1 generic only, 2 specialised required
thread_local int tl_lw_prune = 1;  // barren-node elimination in the specialised kernel

// The specialised kernel for `key` on net's device, or *out = nullptr when the interpreter kernel should serve this
// call: the network does not qualify, the compile failed earlier, or ("auto" modes) the cubin is neither in the
// process nor on disk yet -- then NVRTC is started on a background thread once the cumulative work requested for the key
// reaches `threshold` (tiny calls never trigger a compile) and later calls switch over when it is ready.
// must = true (explicit "specialised" mode) blocks until the kernel exists and reports failures.
struct KernelRequest {
  const char* entry;       // __global__ function name
  int threads;
  size_t smem;
  uint32_t table_words, n_live;
  bool want_occupancy;
};
static int32_t acquire_kernel(Net* net, const std::string& key, bool must, double work, double threshold,
                              const std::function<std::string(KernelRequest*)>& gen, SpecKernel** out) {
  *out = nullptr;
  Driver& drv = driver();
  std::lock_guard<std::mutex> nlk(net->mu);
  auto it = net->spec_cache.find(key);
  if (it != net->spec_cache.end()) {
    *out = it->second;
    BN_REQUIRE(*out || !must, BN_CUDA_ERR, "specialised kernel failed to compile earlier for this (net, evidence, query)");
    return BN_OK;
  }
  if (!must) {
    double& d = net->spec_demand[key];
    d += work;
    if (d < threshold) return BN_OK;
  }
  KernelRequest rq{};
  const std::string src = gen(&rq);
  std::shared_ptr<CubinEntry> ce = get_cubin(src, must);
  if (ce->state == CubinEntry::COMPILING) return BN_OK;  // background compile in flight: interpreter for now
  if (ce->state == CubinEntry::FAILED) {
    net->spec_cache[key] = nullptr;  // do not retry a failing compile on every call
    BN_REQUIRE(!must, BN_CUDA_ERR, "nvrtcCompileProgram failed: %.300s", ce->log.c_str());
    return BN_OK;
  }
  const std::string gkey = std::to_string(net->device) + "|" + src;
  std::lock_guard<std::mutex> lk(g_spec_mu);
  auto git = g_spec.find(gkey);
  SpecKernel* k = nullptr;
  if (git != g_spec.end()) {
    k = git->second.get();
  } else {
    std::unique_ptr<SpecKernel> nk(new SpecKernel());
    CUresult r = drv.ModuleLoadData(&nk->mod, ce->cubin.data());
    if (r != CUDA_SUCCESS) return cu_fail("cuModuleLoadData", r);
    r = drv.ModuleGetFunction(&nk->fn, nk->mod, rq.entry);
    if (r != CUDA_SUCCESS) return cu_fail("cuModuleGetFunction", r);
    nk->threads = rq.threads;
    nk->smem = rq.smem;
    nk->table_words = rq.table_words;
    nk->n_live = rq.n_live;
    r = drv.FuncSetAttribute(nk->fn, CU_FUNC_ATTRIBUTE_MAX_DYNAMIC_SHARED_SIZE_BYTES, (int)rq.smem);
    if (r != CUDA_SUCCESS) return cu_fail("cuFuncSetAttribute", r);
    drv.FuncGetAttribute(&nk->regs, CU_FUNC_ATTRIBUTE_NUM_REGS, nk->fn);
    nk->blocks_per_sm = 1;
    if (rq.want_occupancy) {
      int occ = 0;
      r = drv.OccupancyMaxActiveBlocksPerMultiprocessor(&occ, nk->fn, nk->threads, rq.smem);
      if (r != CUDA_SUCCESS) return cu_fail("cuOccupancyMaxActiveBlocksPerMultiprocessor", r);
      nk->blocks_per_sm = occ > 0 ? occ : 1;
    }
    k = nk.get();
    g_spec[gkey] = std::move(nk);
  }
  net->spec_cache[key] = k;
  *out = k;
  return BN_OK;
}

// "auto" mode thresholds: cumulative particles / rows (node updates for Gibbs) requested for one key before NVRTC is
// started in the background.  The interpreter needs ~15 ms for 5e7 particles of a 100-node net; below that a compile
// (~1 s of one host core) is not worth starting.
static double auto_threshold(const char* env, double dflt) {
  const char* e = getenv(env);
  return e ? atof(e) : dflt;
}

struct SpecLaunch {
  SpecMode mode = SPEC_INFER;
  const int8_t* evidence = nullptr;  // host, n entries or NULL
  bool all_free = false;             // ancestral / rejection: evidence is ignored while sampling
  const int32_t* query = nullptr;
  uint32_t n_query = 0, ncell = 1;
  const std::vector<int64_t>* qstride = nullptr;
  uint64_t first = 0, n = 0, seed = 0;
  double* out = nullptr;             // SPEC_INFER
  uint8_t* out_states = nullptr;     // table modes
  double* out_w = nullptr;
  uint8_t* out_ok = nullptr;
  uint32_t max_attempts = 1;
};

static int32_t launch_specialized(Net* net, const SpecLaunch& a, bool* used) {
  *used = false;
  if (tl_lw_mode == 1) return BN_OK;
  const bool must = tl_lw_mode == 2;
  Driver& drv = driver();
  if (!nvrtc().ok || !drv.ok) {
    BN_REQUIRE(!must, BN_UNSUPPORTED, "specialised sampling kernel unavailable: NVRTC / driver entry points missing");
    return BN_OK;
  }
  const DevProps& dp = dev_props(net->device);
  SampleProgram prg;
  BN_TRY(compile_program(*net, a.evidence, a.all_free, 1, &prg));
  const bool infer = a.mode == SPEC_INFER;
  const size_t smem = (size_t)prg.table_words * 4 + (infer ? (size_t)(a.ncell + 2) * 8 : 0);
  if (smem + 1024 > (size_t)dp.max_smem_optin || net->n > 2048 || prg.table_words > (1u << 22)) {
    BN_REQUIRE(!must, BN_UNSUPPORTED, "network too large for the specialised sampling kernel");
    return BN_OK;
  }
  const bool prune = tl_lw_prune != 0 && infer;
  std::vector<char> ev_mask((size_t)net->n, 0);
  for (int i = 0; i < net->n; ++i) ev_mask[i] = (a.evidence && a.evidence[i] >= 0) ? 1 : 0;
  std::string key((size_t)net->n, '0');
  for (int i = 0; i < net->n; ++i) key[i] = ev_mask[i] ? '1' : '0';
  key += "|m" + std::to_string((int)a.mode) + (a.all_free ? "f" : "e") + (prune ? "p" : "a");
  for (uint32_t j = 0; j < a.n_query; ++j) key += "," + std::to_string(a.query[j]);
  SpecKernel* k = nullptr;
  static const double thr = auto_threshold("BN_SPEC_AUTO_THRESHOLD", 5e7);
  BN_TRY(acquire_kernel(net, key, must, (double)a.n, thr, [&](KernelRequest* rq) {
    SpecShape shp;
    if (const char* e = getenv("BN_SPEC_THREADS")) shp.threads = atoi(e);
    if (const char* e = getenv("BN_SPEC_MINB")) shp.min_blocks = atoi(e);
    uint32_t n_live = 0;
    static const std::vector<int64_t> no_stride;
    std::string src = gen_source(a.mode, *net, prg, ev_mask, a.query, a.n_query, a.qstride ? *a.qstride : no_stride,
                                 a.ncell, prune, shp, &n_live);
    rq->entry = "lw_spec";
    rq->threads = shp.threads;
    rq->smem = smem;
    rq->table_words = prg.table_words;
    rq->n_live = n_live;
    rq->want_occupancy = true;
    return src;
  }, &k));
  if (!k) return BN_OK;  // the interpreter kernel serves this call
  BN_REQUIRE(k->table_words == prg.table_words, BN_CUDA_ERR, "specialised kernel/table mismatch");
  // tables (+ evidence values) -> per-call device staging from the stream-ordered pool (freed in stream order after
  // the kernel when `stage` goes out of scope): evidence VALUES live only here, and concurrent calls never share it
  cudaStream_t s = tl_stream;
  const size_t tbytes = (size_t)prg.table_words * 4;
  const size_t ev_off = (tbytes + 15) / 16 * 16;
  const size_t total = ev_off + (size_t)net->n + 16;
  TmpBuf<unsigned char> stage;
  BN_CUDA_TRY(stage.alloc(total));
  std::vector<unsigned char> host(total, 0);
  if (tbytes) memcpy(host.data(), prg.words.data() + prg.prog_words, tbytes);
  for (int i = 0; i < net->n; ++i) host[ev_off + i] = (unsigned char)(a.evidence ? a.evidence[i] : -1);
  BN_CUDA_TRY(cudaMemcpyAsync(stage.p, host.data(), total, cudaMemcpyHostToDevice, s));
  long want = (long)((a.n + k->threads - 1) / k->threads);
  long full = (long)dp.sm_count * k->blocks_per_sm;
  unsigned grid = (unsigned)(want < full ? (want > 0 ? want : 1) : full);
  const uint32_t* tab_dev = reinterpret_cast<const uint32_t*>(stage.p);
  const signed char* ev_dev = reinterpret_cast<const signed char*>(stage.p + ev_off);
  uint64_t first = a.first, n = a.n;
  uint32_t k0 = (uint32_t)a.seed, k1 = (uint32_t)(a.seed >> 32), never = 0xFFFFFFFFu, max_att = a.max_attempts;
  double* out = a.out;
  uint8_t* out_states = a.out_states;
  double* out_w = a.out_w;
  uint8_t* out_ok = a.out_ok;
  void* args[] = {(void*)&tab_dev, (void*)&first, (void*)&n, (void*)&k0, (void*)&k1, (void*)&never, (void*)&out,
                  (void*)&ev_dev, (void*)&out_states, (void*)&out_w, (void*)&out_ok, (void*)&max_att};
  CUresult r = drv.LaunchKernel(k->fn, grid, 1, 1, (unsigned)k->threads, 1, 1, (unsigned)k->smem, (CUstream)s, args, nullptr);
  if (r != CUDA_SUCCESS) return cu_fail("cuLaunchKernel(lw_spec)", r);
  count_launch();
  *used = true;
  return BN_OK;
}

int32_t launch_lw_specialized(Net* net, const int8_t* evidence, const int32_t* query, uint32_t n_query,
                              const std::vector<int64_t>& qstride, uint32_t ncell, uint64_t first, uint64_t n,
                              uint64_t seed, double* out_dev, bool* used) {
  SpecLaunch a;
  a.mode = SPEC_INFER;
  a.evidence = evidence;
  a.query = query; a.n_query = n_query; a.ncell = ncell; a.qstride = &qstride;
  a.first = first; a.n = n; a.seed = seed; a.out = out_dev;
  return launch_specialized(net, a, used);
}

// kind: BN_SAMPLE_ANCESTRAL / _REJECTION / _WEIGHTED (include/bn_cuda.h)
int32_t launch_table_specialized(Net* net, const int8_t* evidence, int32_t kind, uint32_t max_attempts, uint64_t first,
                                 uint64_t n, uint64_t seed, uint8_t* out_states, double* out_w, uint8_t* out_ok,
                                 bool* used) {
  SpecLaunch a;
  a.mode = kind == BN_SAMPLE_REJECTION ? SPEC_REJECT : SPEC_TABLE;
  a.evidence = kind == BN_SAMPLE_ANCESTRAL ? nullptr : evidence;
  a.all_free = kind != BN_SAMPLE_WEIGHTED;
  a.first = first; a.n = n; a.seed = seed;
  a.out_states = out_states; a.out_w = out_w; a.out_ok = out_ok;
  a.max_attempts = max_attempts;
  return launch_specialized(net, a, used);
}

// ------------------------------------------------------------------------------------------ Gibbs
// Network-specialised batched Gibbs inference (src/Inference/gibbs.jl:66-145 Nodewise, :161-234 Full): one chain per
// thread, states u8 [node][thread] in shared memory (offsets are immediates), one straight-line block per free
// node with its Markov blanket unrolled: p[v] = prod_children P(x_c | pa_c, x_n = v) * P(v | pa_n) in fp64, children
// first and own CPD last exactly like the reference's foldl, then StatsBase's cumulative scan.  Same Philox stream
// and arithmetic as gibbs_infer_kernel / oracle -> identical integer counts.  burn-in / thinning bookkeeping is
// chain-independent (a countdown in uniform registers).
constexpr int GIBBS_SPEC_MAX_CARD = 8;

// The sweep is one long straight-line block (~55 instructions per node, > 100 KB for 200 nodes).  Everything that
// is not per-node arithmetic stays out of line or out of the loop: the Philox refill every 4th update is a call (inlined
// where its site is static), the thinned record step is a call that counts in per-thread shared-memory counters (<= 16
// query cells) or through one MATCH.ANY-grouped atomic per cell.
static const char* kGibbsPrelude = R"BNSRC(
#ifndef REFILL_ATTR
#define REFILL_ATTR __noinline__
#endif
__device__ REFILL_ATTR uint4 refill(u32 c0, u32 c1, u32 blk, u32 k0, u32 k1) {  // by value: the block stays in registers
  uint4 b;
  px(c0, c1, blk, 2u, k0, k1, b.x, b.y, b.z, b.w);
  return b;
}
#define NEXT_R(r) { if ((d & 3u) == 0u) { const uint4 nb = refill(c0, c1, d >> 2, k0, k1); b0 = nb.x; b1 = nb.y; b2 = nb.z; b3 = nb.w; } \
                    r = (d & 2u) ? ((d & 1u) ? b3 : b2) : ((d & 1u) ? b1 : b0); ++d; }
)BNSRC";

// Rows a node's blanket may have to get a table.  The tables share L1 with the fp64 CPTs (what is left of the SM's 256 KB
// after two CTAs' state columns, ~124 KB): measured on the C4 net (CPT 64 KB), 256 rows (61 KB of tables) is 8% faster
// than no tables, 1024 rows (126 KB) 7% slower.  So: the largest power of two whose tables + CPTs stay within 128 KB.
// BN_GIBBS_TABLE_ROWS overrides (0 = no tables).
uint32_t gibbs_table_row_limit(const Net& net) {
  if (const char* e = getenv("BN_GIBBS_TABLE_ROWS")) return (uint32_t)atoi(e);
  if (net.gibbs_row_limit >= 0) return (uint32_t)net.gibbs_row_limit;  // decided once per net (structure only)
  const uint64_t budget = 128u << 10, cpt_bytes = (uint64_t)net.cpt_off[net.n] * 8;
  uint32_t best = 0;
  for (uint32_t lim = 16; lim <= 4096; lim *= 2) {
    GibbsTables g;
    gibbs_tables_layout(net, lim, &g);
    if (cpt_bytes + g.words * 4 <= budget) best = lim; else break;
  }
  net.gibbs_row_limit = (int)best;
  return best;
}

// Update of a node whose Markov-blanket conditional is tabulated (gibbs_tables.cu): row index from the blanket's states,
// one vector load of the K-1 integer thresholds, pick = #{v : r >= R_v} -- the same decision as the fp64 scan for every r.
static void emit_table_update(std::ostringstream& o, const Net& net, const GibbsTables& gt, int i, const char* ind) {
  const int K = net.card[i];
  const std::vector<int>& mb = gt.members[i];
  o << ind << "const u32* tp = tab + (" << (uint32_t)gt.tab_off[i] << "u";
  const uint32_t W = K <= 2 ? 1u : (K == 3 ? 2u : (K == 4 ? 4u : 8u));
  for (size_t j = 0; j < mb.size(); ++j) o << " + (u32)st[" << mb[j] << " * NT] * " << gt.strides[i][j] * W << "u";
  o << ");\n";
  if (K == 2) {
    o << ind << "const u32 pick = (u32)(r >= __ldg(tp));\n";
  } else if (K == 3) {
    o << ind << "const uint2 t = __ldg(reinterpret_cast<const uint2*>(tp));\n";
    o << ind << "const u32 pick = (u32)(r >= t.x) + (u32)(r >= t.y);\n";
  } else {
    o << ind << "const uint4 t = __ldg(reinterpret_cast<const uint4*>(tp));\n";
    o << ind << "u32 pick = (u32)(r >= t.x) + (u32)(r >= t.y) + (u32)(r >= t.z);\n";
    if (K >= 5) {
      o << ind << "pick += (u32)(r >= t.w);\n";
      if (K >= 6) {
        o << ind << "{ const uint4 t2 = __ldg(reinterpret_cast<const uint4*>(tp) + 1); pick += (u32)(r >= t2.x)";
        if (K >= 7) o << " + (u32)(r >= t2.y)";
        if (K >= 8) o << " + (u32)(r >= t2.z)";
        o << "; }\n";
      }
    }
  }
}

static std::string gen_gibbs_source(const Net& net, const std::vector<char>& ev_mask, const int32_t* query,
                                    uint32_t n_query, const std::vector<int64_t>& qstride, uint32_t ncell, bool full,
                                    int threads, int min_blocks, uint32_t* n_free_out, const GibbsTables* gt = nullptr) {
  const int n = net.n;
  std::ostringstream o;
  uint32_t n_free = 0;
  for (int i = 0; i < n; ++i) n_free += ev_mask[i] ? 0 : 1;

  o << "// generated by libbn_cuda (specialize.cu): Gibbs " << (full ? "Full" : "Nodewise") << ", " << n << " nodes, "
    << n_free << " free, query cells " << ncell << "\n";
  // Draw d of a chain is word d & 3 of Philox block d >> 2, and d advances by one per update: when the number of free
  // nodes is a multiple of 4 the compiler resolves d & 3 at every node (the refill sits at every 4th node, the word is
  // a fixed register), and the refill can be INLINED there: the round keys become uniform-register operands (20 VIADD
  // per block gone), no call / argument moves.  Otherwise the refill site depends on the sweep and stays a call.
  {
    const char* e = getenv("BN_GIBBS_INLINE_REFILL");
    const bool inline_refill = e ? atoi(e) != 0 : (n_free % 4u == 0u);
    if (inline_refill) o << "#define REFILL_ATTR __forceinline__\n";
  }
  o << kPrelude << kGibbsPrelude;
  o << "#define NT " << threads << "\n#define NN " << n << "u\n#define NCELL " << ncell << "u\n";
  // record step, out of line (it runs once per `thin` updates, but inline it was ~35 instructions in EVERY node's
  // block): <= 16 query cells count in per-thread shared-memory counters cnt[cell][thread] (no atomics until the end);
  // larger tables group the warp's lanes by cell (MATCH.ANY) and the group leader adds the group size to the CTA's bins
  const bool reg_bins = ncell <= (uint32_t)SPEC_MAX_REG_CELLS;
  o << "#define CELL (0u";
  for (uint32_t j = 0; j < n_query; ++j) o << " + (u32)st[" << query[j] << " * NT] * " << (uint32_t)qstride[j] << "u";
  o << ")\n";
  if (reg_bins) {
    o << "__device__ __noinline__ void record(const unsigned char* st, u32* cnt, double* cc, u64 ci, bool live) {\n";
    o << "  const u32 cell = CELL;\n";
    o << "  cnt[cell * NT] += 1u;\n";
    o << "  if (cc && live) cc[ci * NCELL + cell] += 1.0;\n}\n";
    o << "#define RECORD record(st, cnt, cc, ci, live);\n";
  } else {
    o << "__device__ __noinline__ void record(const unsigned char* st, unsigned long long* bins, double* cc, u64 ci, bool live) {\n";
    o << "  const u32 cell = live ? CELL : 0xFFFFFFFFu;\n";
    o << "  const u32 grp = __match_any_sync(0xffffffffu, cell);\n";
    o << "  if (live) { if ((u32)(__ffs(grp) - 1) == (threadIdx.x & 31u)) atomicAdd(&bins[cell], (unsigned long long)__popc(grp));\n";
    o << "              if (cc) cc[ci * NCELL + cell] += 1.0; }\n}\n";
    o << "#define RECORD record(st, bins, cc, ci, live);\n";
  }
  // 32-bit countdowns (the launcher sends burn-in + thin >= 2^31 or more than 2^26 kept samples to the interpreter)
  o << "#define TICK { if (--until == 0u) { until = (u32)thin; RECORD; if (++smp >= (u32)total) goto finished; } }\n";
  o << "extern \"C\" __global__ void __launch_bounds__(NT, " << min_blocks << ") gibbs_spec(const double* __restrict__ cpt, "
       "const signed char* __restrict__ ev, u64 first_chain, u64 n_chains, u32 k0, u32 k1, long long burn_in, long long thin, "
       "long long total, double* __restrict__ out, double* __restrict__ cc, unsigned char* __restrict__ state_io, "
       "int init_from_state, const u32* __restrict__ tab) {\n";
  o << "  extern __shared__ __align__(16) unsigned char smem_raw[];\n";
  o << "  unsigned long long* bins = reinterpret_cast<unsigned long long*>(smem_raw);\n";
  o << "  const u32 tid = threadIdx.x;\n";
  if (reg_bins) {
    o << "  u32* cnt = reinterpret_cast<u32*>(smem_raw + NCELL * 8u) + tid;\n";
    o << "  for (u32 c = 0; c < NCELL; ++c) cnt[c * NT] = 0u;\n";
    o << "  unsigned char* st = smem_raw + NCELL * 8u + NCELL * NT * 4u + tid;\n";
  } else {
    o << "  unsigned char* st = smem_raw + NCELL * 8u + tid;\n";
  }
  o << "  for (u32 c = tid; c < NCELL; c += NT) bins[c] = 0ull;\n";
  o << "  __syncthreads();\n";
  o << "  const u64 ci = (u64)blockIdx.x * NT + tid;\n";
  o << "  const bool live = ci < n_chains;  // surplus threads of the last CTA run a dummy chain (barriers need them)\n";
  o << "  {\n";
  o << "    const u64 chain = first_chain + ci;\n";
  o << "    const u32 c0 = (u32)chain, c1 = (u32)(chain >> 32);\n";
  // ---- initial state: resume (im.state) or uniform (src/Inference/gibbs.jl:6-20), stream 1, static draw indices
  o << "    if (state_io && init_from_state) {\n";
  o << "      for (u32 i = 0; i < NN; ++i) st[i * NT] = live ? state_io[ci * NN + i] : (unsigned char)0;\n";
  o << "    } else {\n";
  {
    uint32_t d = 0;
    int cur = -1;
    for (int i = 0; i < n; ++i) {
      if (ev_mask[i]) { o << "      st[" << i << " * NT] = (unsigned char)ev[" << i << "];\n"; continue; }
      const int blk = (int)(d >> 2);
      if (blk != cur) {
        cur = blk;
        o << "      u32 i" << blk << "_0, i" << blk << "_1, i" << blk << "_2, i" << blk << "_3; px(c0, c1, " << blk
          << "u, 1u, k0, k1, i" << blk << "_0, i" << blk << "_1, i" << blk << "_2, i" << blk << "_3);\n";
      }
      o << "      st[" << i << " * NT] = (unsigned char)__umulhi(i" << blk << "_" << (d & 3u) << ", " << net.card[i] << "u);\n";
      ++d;
    }
  }
  o << "    }\n";
  o << "    u32 d = 0u, b0 = 0u, b1 = 0u, b2 = 0u, b3 = 0u;\n";
  o << "    u32 until = (u32)(burn_in + thin), smp = 0u;\n";
  if (n_free == 0 && !full) {
    o << "    goto finished;\n";
  }
  o << "    for (;;) {\n";
  int emitted = 0;
  // Re-convergence barrier every N nodes: the generated code (~1.2 KB per node) is larger than the instruction caches,
  // and 20% of the warp-state samples were `no_inst`; 256-thread CTAs that meet every 32 nodes fetch each code
  // window once for 8 warps (C4: 0.84 -> 0.76 ms; profiles/r1f_gibbs_sync_sweep.txt).  64-thread CTAs do not sync.
  // With the trimmed blocks of round 2 any interval between 48 and a whole sweep measures the same (0.57-0.58 ms),
  // 16 is 4% slower, none at all 12% slower.
  const int sync_every = getenv("BN_GIBBS_SYNC") ? std::max(1, atoi(getenv("BN_GIBBS_SYNC"))) : (threads >= 128 ? 64 : (1 << 30));
  for (int i = 0; i < n; ++i) {
    if (ev_mask[i]) continue;
    const int K = net.card[i];
    if (gt && gt->tab_off[i] >= 0) {
      o << "      {  // node " << i << ": card " << K << ", tabulated conditional over " << gt->members[i].size() << " blanket nodes\n";
      o << "        u32 r; NEXT_R(r);\n";
      emit_table_update(o, net, *gt, i, "        ");
      o << "        st[" << i << " * NT] = (unsigned char)pick;\n";
      if (!full) o << "        TICK\n";
      o << "      }\n";
      if (++emitted % sync_every == 0) o << "      __syncthreads();  // keep the CTA's warps on the same code window\n";
      continue;
    }
    o << "      {  // node " << i << ": card " << K << ", " << (net.child_ptr[i + 1] - net.child_ptr[i]) << " children\n";
    o << "        u32 r; NEXT_R(r);\n";
    o << "        const double uu = ((double)r + 0.5) * (1.0 / 4294967296.0);\n";
    o << "        double";
    for (int v = 0; v < K; ++v) o << (v ? ", p" : " p") << v;
    o << ";\n";
    bool first = true;
    auto emit_cpd = [&](int c, bool own) {
      // element index of the row P(. | pa_c) with x_i left out; `step` = stride of x_i in elements
      const int Kc = net.card[c];
      std::ostringstream idx;
      idx << (uint32_t)net.cpt_off[c] << "u";
      uint32_t step = own ? 1u : 0u;
      int64_t stride = 1;
      for (int k = net.par_ptr[c]; k < net.par_ptr[c + 1]; ++k) {
        const int pa = net.par_idx[k];
        if (!own && pa == i) step = (uint32_t)(stride * Kc);
        else idx << " + (u32)st[" << pa << " * NT] * " << (uint32_t)(stride * Kc) << "u";
        stride *= net.card[pa];
      }
      if (!own) idx << " + (u32)st[" << c << " * NT]";
      o << "        { const double* pr = cpt + (" << idx.str() << ");";
      for (int v = 0; v < K; ++v)
        o << " p" << v << (first ? " = " : " *= ") << "__ldg(pr + " << (uint32_t)v * step << "u);";
      o << " }  // " << (own ? "own CPD" : "child ") << c << "\n";
      first = false;
    };
    for (int k = net.child_ptr[i]; k < net.child_ptr[i + 1]; ++k) emit_cpd(net.child_idx[k], false);
    emit_cpd(i, true);
    o << "        double sum = p0;";
    for (int v = 1; v < K; ++v) o << " sum += p" << v << ";";
    o << "\n        const double t = uu * sum;\n";
    o << "        u32 pick = 0u; double cw = p0; bool go = true;\n";
    for (int v = 1; v < K; ++v)
      o << "        go = go && (cw < t); if (go) { cw += p" << v << "; pick = " << v << "u; }\n";
    o << "        st[" << i << " * NT] = (unsigned char)pick;\n";
    if (!full) o << "        TICK\n";
    o << "      }\n";
    if (++emitted % sync_every == 0) o << "      __syncthreads();  // keep the CTA's warps on the same code window\n";
  }
  if (full) o << "      TICK\n";
  o << "    }\n";
  o << "  finished:\n";
  o << "    if (state_io && live) for (u32 i = 0; i < NN; ++i) state_io[ci * NN + i] = st[i * NT];\n";
  o << "  }\n";
  if (reg_bins) {
    for (uint32_t c = 0; c < ncell; ++c) {
      o << "  { u32 v = live ? cnt[" << c << "u * NT] : 0u;\n";
      o << "#pragma unroll\n    for (int of = 16; of > 0; of >>= 1) v += __shfl_xor_sync(0xffffffffu, v, of);\n";
      o << "    if ((tid & 31u) == 0u && v) atomicAdd(&bins[" << c << "], (unsigned long long)v); }\n";
    }
  }
  o << "  __syncthreads();\n";
  o << "  for (u32 c = tid; c < NCELL; c += NT) if (bins[c]) atomicAdd(&out[c], (double)bins[c]);\n";
  o << "}\n";
  if (n_free_out) *n_free_out = n_free;
  return o.str();
}

// gibbs_sample(bn, nsamples, burn_in; thinning, variable_order) specialised the same way (replaces the interpreter
// gibbs_sample_kernel of gibbs.cu for src/gibbs.jl:183-245): the per-node blocks are the ones gibbs_spec uses, but
// the sweep order changes every sweep (a fresh permutation, shared by all chains, :222-234), so the blocks sit behind
// one `switch` on the sweep's order entry -- a warp-uniform indexed branch per update.  At small chain counts the
// interpreter is bound by its ~2000-cycle dependent latency per update (4096 chains: one warp per scheduler); the
// straight-line block cuts that by an order of magnitude.
static std::string gen_gibbs_sample_source(const Net& net, const std::vector<char>& ev_mask, int threads, int min_blocks,
                                           int sync_every, uint32_t* n_free_out, const GibbsTables* gt = nullptr) {
  const int n = net.n;
  std::ostringstream o;
  uint32_t n_free = 0;
  for (int i = 0; i < n; ++i) n_free += ev_mask[i] ? 0 : 1;
  o << "// generated by libbn_cuda (specialize.cu): gibbs_sample, " << n << " nodes, " << n_free << " free\n";
  o << kPrelude << kGibbsPrelude;
  o << "#define NT " << threads << "\n#define NN " << n << "u\n#define NFREE " << n_free << "u\n";
  o << "extern \"C\" __global__ void __launch_bounds__(NT, " << min_blocks << ") gibbs_sample_spec(const double* __restrict__ cpt, "
       "const int* __restrict__ orders, u32 order_stride, const unsigned char* __restrict__ init, u64 first_chain, "
       "u64 n_chains, u32 k0, u32 k1, long long burn_in, long long nsamples, long long thinning, "
       "unsigned char* __restrict__ out_states, u64 n_rows, const u32* __restrict__ tab) {\n";
  o << "  extern __shared__ __align__(16) unsigned char smem_raw[];\n";
  o << "  const u32 tid = threadIdx.x;\n";
  o << "  unsigned char* st = smem_raw + tid;\n";
  o << "  const u64 ci = (u64)blockIdx.x * NT + tid;\n";
  o << "  const bool live = ci < n_chains;  // surplus threads of the last CTA run a dummy chain (barriers need them)\n";
  o << "  const u64 chain = first_chain + ci;\n";
  o << "  const u32 c0 = (u32)chain, c1 = (u32)(chain >> 32);\n";
  o << "  for (u32 i = 0; i < NN; ++i) st[i * NT] = live ? init[ci * NN + i] : (unsigned char)0;\n";
  o << "  u32 d = 0u, b0 = 0u, b1 = 0u, b2 = 0u, b3 = 0u;\n";
  o << "  const long long total_sweeps = burn_in + nsamples * (thinning + 1);\n";
  o << "  long long until = burn_in + thinning + 1, row = 0;  // sweeps until the next kept sample (src/gibbs.jl:226-240)\n";
  o << "  for (long long sweep = 0; sweep < total_sweeps; ++sweep) {\n";
  o << "    const int* ord = orders + (u64)sweep * order_stride;\n";
  o << "    for (u32 f = 0; f < NFREE; ++f) {\n";
  o << "      u32 r; NEXT_R(r);\n";
  o << "      switch (__ldg(ord + f)) {\n";
  uint32_t rank = 0;
  for (int i = 0; i < n; ++i) {
    if (ev_mask[i]) continue;
    const int K = net.card[i];
    if (gt && gt->tab_off[i] >= 0) {
      o << "      case " << rank++ << ": {  // node " << i << ": card " << K << ", tabulated conditional over " << gt->members[i].size() << " blanket nodes\n";
      emit_table_update(o, net, *gt, i, "        ");
      o << "        st[" << i << " * NT] = (unsigned char)pick;\n";
      o << "      } break;\n";
      continue;
    }
    o << "      case " << rank++ << ": {  // node " << i << ": card " << K << ", " << (net.child_ptr[i + 1] - net.child_ptr[i]) << " children\n";
    o << "        const double uu = ((double)r + 0.5) * (1.0 / 4294967296.0);\n";
    o << "        double";
    for (int v = 0; v < K; ++v) o << (v ? ", p" : " p") << v;
    o << ";\n";
    bool first = true;
    auto emit_cpd = [&](int c, bool own) {
      const int Kc = net.card[c];
      std::ostringstream idx;
      idx << (uint32_t)net.cpt_off[c] << "u";
      uint32_t step = own ? 1u : 0u;
      int64_t stride = 1;
      for (int k = net.par_ptr[c]; k < net.par_ptr[c + 1]; ++k) {
        const int pa = net.par_idx[k];
        if (!own && pa == i) step = (uint32_t)(stride * Kc);
        else idx << " + (u32)st[" << pa << " * NT] * " << (uint32_t)(stride * Kc) << "u";
        stride *= net.card[pa];
      }
      if (!own) idx << " + (u32)st[" << c << " * NT]";
      o << "        { const double* pr = cpt + (" << idx.str() << ");";
      for (int v = 0; v < K; ++v)
        o << " p" << v << (first ? " = " : " *= ") << "__ldg(pr + " << (uint32_t)v * step << "u);";
      o << " }  // " << (own ? "own CPD" : "child ") << c << "\n";
      first = false;
    };
    for (int k = net.child_ptr[i]; k < net.child_ptr[i + 1]; ++k) emit_cpd(net.child_idx[k], false);
    emit_cpd(i, true);
    o << "        double sum = p0;";
    for (int v = 1; v < K; ++v) o << " sum += p" << v << ";";
    o << "\n        const double t = uu * sum;\n";
    o << "        u32 pick = 0u; double cw = p0; bool go = true;\n";
    for (int v = 1; v < K; ++v)
      o << "        go = go && (cw < t); if (go) { cw += p" << v << "; pick = " << v << "u; }\n";
    o << "        st[" << i << " * NT] = (unsigned char)pick;\n";
    o << "      } break;\n";
  }
  o << "      default: break;\n      }\n";
  if (sync_every > 0) o << "      if ((f % " << sync_every << "u) == " << (sync_every - 1) << "u) __syncthreads();  // one instruction fetch per block for the whole CTA\n";
  o << "    }\n";
  o << "    if (--until == 0) {\n      until = thinning + 1;\n";
  o << "      if (live) for (u32 i = 0; i < NN; ++i) out_states[(u64)i * n_rows + ci * (u64)nsamples + (u64)row] = st[i * NT];\n";
  o << "      ++row;\n    }\n  }\n}\n";
  if (n_free_out) *n_free_out = n_free;
  return o.str();
}

thread_local int tl_gibbs_mode = 0;  // 0 auto, 1 generic only, 2 specialised required

int32_t launch_gibbs_specialized(Net* net, const int8_t* evidence, const int32_t* query, uint32_t n_query,
                                 const std::vector<int64_t>& qstride, uint32_t ncell, bool full, uint64_t first_chain,
                                 uint64_t n_chains, int64_t burn_in, int64_t thin, int64_t total, uint64_t seed,
                                 double* out_dev, double* chain_counts_dev, uint8_t* state_dev, int32_t init_from_state,
                                 bool* used) {
  *used = false;
  if (tl_gibbs_mode == 1) return BN_OK;
  const bool must = tl_gibbs_mode == 2;
  Driver& drv = driver();
  if (!nvrtc().ok || !drv.ok) {
    BN_REQUIRE(!must, BN_UNSUPPORTED, "specialised Gibbs kernel unavailable: NVRTC / driver entry points missing");
    return BN_OK;
  }
  bool ok = net->n <= 1024 && net->cpt_off[net->n] < (int64_t)0x7fffffff && total < (int64_t)(1ll << 26) &&
            burn_in + thin < (int64_t)0x7fffffff;  // 32-bit countdowns; 32 lanes x
  // total kept samples must fit the u32 warp sums
  for (int i = 0; i < net->n && ok; ++i) ok = net->card[i] <= GIBBS_SPEC_MAX_CARD;
  const DevProps& dp = dev_props(net->device);
  // small CTAs: the grid balances dynamically and 64-thread CTAs measured fastest (profiles/r1d_gibbs_spec_sweep.txt)
  int T = n_chains >= (uint64_t)dp.sm_count * 256 ? 256 : 64;  // one 256-thread CTA per SM or more: see sync_every
  while (T > 64 && (size_t)ncell * 8 + (size_t)(net->n + 4 * SPEC_MAX_REG_CELLS) * T + 1024 > (size_t)dp.max_smem_optin) T -= 32;
  if (const char* e = getenv("BN_GIBBS_THREADS")) T = atoi(e);
  const size_t cnt_bytes = ncell <= (uint32_t)SPEC_MAX_REG_CELLS ? (size_t)ncell * T * 4 : 0;  // per-thread record counters
  const size_t smem = (size_t)ncell * 8 + cnt_bytes + (size_t)net->n * T;
  if (!ok || smem + 1024 > (size_t)dp.max_smem_optin) {
    BN_REQUIRE(!must, BN_UNSUPPORTED, "network not supported by the specialised Gibbs kernel (card > %d or too large)",
               GIBBS_SPEC_MAX_CARD);
    return BN_OK;
  }
  std::vector<char> ev_mask((size_t)net->n, 0);
  for (int i = 0; i < net->n; ++i) ev_mask[i] = (evidence && evidence[i] >= 0) ? 1 : 0;
  std::string key((size_t)net->n, '0');
  for (int i = 0; i < net->n; ++i) key[i] = ev_mask[i] ? '1' : '0';
  key += std::string("|g") + (full ? "F" : "N") + std::to_string(T);
  for (uint32_t j = 0; j < n_query; ++j) key += "," + std::to_string(query[j]);
  // tabulated conditionals for nodes with small Markov blankets (gibbs_tables.cu); BN_GIBBS_TABLE_ROWS=0 turns them off
  const GibbsTables* gt = nullptr;
  {
    const uint32_t row_limit = gibbs_table_row_limit(*net);
    if (row_limit > 0) {
      BN_TRY(gibbs_tables_get(net, row_limit, &gt));
      if (gt && (gt->words == 0 || gt->has_never)) gt = nullptr;
      if (gt) key += "|t" + std::to_string(row_limit);
    }
  }
  SpecKernel* k = nullptr;
  static const double thr = auto_threshold("BN_GIBBS_AUTO_THRESHOLD", 2e8);
  BN_TRY(acquire_kernel(net, key, must, (double)n_chains * (double)(burn_in + total * thin) * (full ? net->n : 1), thr,
                        [&](KernelRequest* rq) {
    uint32_t n_free = 0;
    int minb = 512 / T;  // <= 128 registers per thread
    if (const char* e = getenv("BN_GIBBS_MINB")) minb = atoi(e);
    std::string src = gen_gibbs_source(*net, ev_mask, query, n_query, qstride, ncell, full, T, minb, &n_free, gt);
    rq->entry = "gibbs_spec";
    rq->threads = T;
    rq->smem = smem;
    rq->n_live = n_free;
    return src;
  }, &k));
  if (!k) return BN_OK;  // the interpreter kernel serves this call
  // evidence values -> per-call device staging (stream-ordered pool; freed in stream order after the kernel)
  cudaStream_t s = tl_stream;
  const size_t total_b = (size_t)net->n + 16;
  TmpBuf<unsigned char> stage;
  BN_CUDA_TRY(stage.alloc(total_b));
  std::vector<unsigned char> host(total_b, 0);
  for (int i = 0; i < net->n; ++i) host[i] = (unsigned char)(evidence ? evidence[i] : -1);
  BN_CUDA_TRY(cudaMemcpyAsync(stage.p, host.data(), total_b, cudaMemcpyHostToDevice, s));
  const double* cpt = net->d_cpt.p;
  const signed char* ev_dev = reinterpret_cast<const signed char*>(stage.p);
  uint32_t k0 = (uint32_t)seed, k1 = (uint32_t)(seed >> 32);
  long long bi = burn_in, th = thin, tot = total;
  const uint32_t* tab_dev = gt ? gt->tab.p : nullptr;
  void* args[] = {(void*)&cpt, (void*)&ev_dev, (void*)&first_chain, (void*)&n_chains, (void*)&k0, (void*)&k1, (void*)&bi,
                  (void*)&th, (void*)&tot, (void*)&out_dev, (void*)&chain_counts_dev, (void*)&state_dev,
                  (void*)&init_from_state, (void*)&tab_dev};
  const unsigned grid = (unsigned)((n_chains + k->threads - 1) / k->threads);
  CUresult r = drv.LaunchKernel(k->fn, grid, 1, 1, (unsigned)k->threads, 1, 1, (unsigned)k->smem, (CUstream)s, args, nullptr);
  if (r != CUDA_SUCCESS) return cu_fail("cuLaunchKernel(gibbs_spec)", r);
  count_launch();
  *used = true;
  return BN_OK;
}

// Specialised gibbs_sample: orders / init / out are device pointers prepared by bn_gibbs_sample (gibbs.cu).
int32_t launch_gibbs_sample_specialized(Net* net, const int8_t* evidence, const int32_t* orders_dev, uint32_t order_stride,
                                        const uint8_t* init_dev, uint64_t first_chain, uint64_t n_chains, int64_t burn_in,
                                        int64_t nsamples, int64_t thinning, uint64_t seed, uint8_t* out_states_dev,
                                        uint64_t n_rows, bool* used) {
  *used = false;
  if (tl_gibbs_mode == 1) return BN_OK;
  const bool must = tl_gibbs_mode == 2;
  Driver& drv = driver();
  if (!nvrtc().ok || !drv.ok) return BN_OK;  // the interpreter kernel serves the call
  bool ok = net->n <= 1024 && net->cpt_off[net->n] < (int64_t)0x7fffffff;
  for (int i = 0; i < net->n && ok; ++i) ok = net->card[i] <= GIBBS_SPEC_MAX_CARD;
  const DevProps& dp = dev_props(net->device);
  int T = 64, sync_every = 0;
  if (const char* e = getenv("BN_GSAMPLE_THREADS")) T = atoi(e);
  if (const char* e = getenv("BN_GSAMPLE_SYNC")) sync_every = atoi(e);
  const size_t smem = (size_t)net->n * T;
  if (!ok || smem + 1024 > (size_t)dp.max_smem_optin) return BN_OK;
  std::vector<char> ev_mask((size_t)net->n, 0);
  uint32_t n_free = 0;
  for (int i = 0; i < net->n; ++i) { ev_mask[i] = (evidence && evidence[i] >= 0) ? 1 : 0; n_free += ev_mask[i] ? 0 : 1; }
  if (n_free == 0) return BN_OK;
  std::string key((size_t)net->n, '0');
  for (int i = 0; i < net->n; ++i) key[i] = ev_mask[i] ? '1' : '0';
  key += "|gsample" + std::to_string(T) + "s" + std::to_string(sync_every);
  const GibbsTables* gt = nullptr;  // tabulated conditionals, exactly as in launch_gibbs_specialized
  {
    const uint32_t row_limit = gibbs_table_row_limit(*net);
    if (row_limit > 0) {
      BN_TRY(gibbs_tables_get(net, row_limit, &gt));
      if (gt && (gt->words == 0 || gt->has_never)) gt = nullptr;
      if (gt) key += "|t" + std::to_string(row_limit);
    }
  }
  SpecKernel* k = nullptr;
  static const double thr = auto_threshold("BN_GIBBS_AUTO_THRESHOLD", 2e8);
  int32_t rc = acquire_kernel(net, key, must, (double)n_chains * (double)(burn_in + nsamples * (thinning + 1)) * n_free, thr,
                              [&](KernelRequest* rq) {
    std::string src = gen_gibbs_sample_source(*net, ev_mask, T, 512 / T, sync_every, nullptr, gt);
    rq->entry = "gibbs_sample_spec";
    rq->threads = T;
    rq->smem = smem;
    rq->n_live = n_free;
    return src;
  }, &k);
  if (rc != BN_OK || !k) return BN_OK;  // the interpreter kernel serves this call (it covers every network)
  const double* cpt = net->d_cpt.p;
  uint32_t k0 = (uint32_t)seed, k1 = (uint32_t)(seed >> 32);
  long long bi = burn_in, ns = nsamples, th = thinning;
  const uint32_t* tab_dev = gt ? gt->tab.p : nullptr;
  void* args[] = {(void*)&cpt, (void*)&orders_dev, (void*)&order_stride, (void*)&init_dev, (void*)&first_chain, (void*)&n_chains,
                  (void*)&k0, (void*)&k1, (void*)&bi, (void*)&ns, (void*)&th, (void*)&out_states_dev, (void*)&n_rows,
                  (void*)&tab_dev};
  const unsigned grid = (unsigned)((n_chains + k->threads - 1) / k->threads);
  CUresult r = drv.LaunchKernel(k->fn, grid, 1, 1, (unsigned)k->threads, 1, 1, (unsigned)k->smem, (CUstream)tl_stream, args, nullptr);
  if (r != CUDA_SUCCESS) return cu_fail("cuLaunchKernel(gibbs_sample_spec)", r);
  count_launch();
  *used = true;
  return BN_OK;
}

}  // namespace bn

using namespace bn;

extern "C" {

uint64_t bn_lw_spec_compile_count(void) { return bn::g_spec_compiles.load(); }

int32_t bn_spec_cache_stats(uint64_t* out4) {
  BN_REQUIRE(out4, BN_BAD_ARG, "bn_spec_cache_stats: NULL output");
  out4[0] = bn::g_spec_compiles.load();
  out4[1] = bn::g_bg_compiles.load();
  out4[2] = bn::g_disk_hits.load();
  std::lock_guard<std::mutex> lk(bn::g_spec_mu);
  out4[3] = (uint64_t)bn::g_spec.size();
  return BN_OK;
}

int32_t bn_set_gibbs_mode(int32_t mode) {
  BN_REQUIRE(mode >= 0 && mode <= 2, BN_BAD_ARG, "bn_set_gibbs_mode: mode %d outside 0..2", mode);
  bn::tl_gibbs_mode = mode;
  return BN_OK;
}

int32_t bn_set_lw_mode(int32_t mode, int32_t prune) {
  BN_REQUIRE(mode >= 0 && mode <= 2, BN_BAD_ARG, "bn_set_lw_mode: mode %d outside 0..2", mode);
  bn::tl_lw_mode = mode;
  bn::tl_lw_prune = prune ? 1 : 0;
  return BN_OK;
}

// Host-only: the specialised Gibbs inference source for a flat network (and its NVRTC cubin size when asked).
int32_t bn_gibbs_spec_source(int32_t n_nodes, const int32_t* card, const int32_t* par_ptr, const int32_t* par_idx,
                             const int64_t* cpt_off, const double* cpt, const int8_t* evidence, const int32_t* query,
                             int32_t n_query, int32_t full, int32_t threads, char* out_src, int64_t cap, int64_t* out_len,
                             int64_t* cubin_size) {
  BN_REQUIRE(card && par_ptr && cpt_off && cpt && out_len, BN_BAD_ARG, "bn_gibbs_spec_source: NULL argument");
  Net net;
  net.n = n_nodes;
  net.card.assign(card, card + n_nodes);
  net.par_ptr.assign(par_ptr, par_ptr + n_nodes + 1);
  net.par_idx.assign(par_idx, par_idx + par_ptr[n_nodes]);
  net.cpt_off.assign(cpt_off, cpt_off + n_nodes + 1);
  net.cpt.assign(cpt, cpt + cpt_off[n_nodes]);
  build_children(&net);
  std::vector<int64_t> qstride;
  int64_t cells = 1;
  for (int j = 0; j < n_query; ++j) { qstride.push_back(cells); cells *= card[query[j]]; }
  std::vector<char> ev_mask((size_t)n_nodes, 0);
  for (int i = 0; i < n_nodes; ++i) ev_mask[i] = (evidence && evidence[i] >= 0) ? 1 : 0;
  const int T = threads > 0 ? threads : 64;  // same launch bounds as launch_gibbs_specialized: <= 128 registers
  GibbsTables gt;  // layout only (host): the source launch_gibbs_specialized generates when no tabled row needs 2^32
  const uint32_t row_limit = gibbs_table_row_limit(net);
  if (row_limit > 0) gibbs_tables_layout(net, row_limit, &gt);
  const std::string src = gen_gibbs_source(net, ev_mask, query, (uint32_t)n_query, qstride, (uint32_t)cells, full != 0,
                                           T, std::max(1, 512 / T), nullptr, (row_limit > 0 && gt.words > 0) ? &gt : nullptr);
  *out_len = (int64_t)src.size();
  if (out_src && cap > 0) {
    const size_t m = src.size() < (size_t)(cap - 1) ? src.size() : (size_t)(cap - 1);
    memcpy(out_src, src.data(), m);
    out_src[m] = 0;
  }
  if (cubin_size) {  // through the process + disk cache, so the cubin can be inspected (cuobjdump -sass) afterwards
    BN_REQUIRE(nvrtc().ok, BN_UNSUPPORTED, "libnvrtc.so.12 could not be loaded (set BN_NVRTC_PATH)");
    std::shared_ptr<CubinEntry> ce = get_cubin(src, true);
    BN_REQUIRE(ce->state == CubinEntry::READY, BN_CUDA_ERR, "nvrtcCompileProgram failed: %.300s", ce->log.c_str());
    *cubin_size = (int64_t)ce->cubin.size();
  }
  return BN_OK;
}

// Host-only (no GPU needed): emit the specialised source for a flat network and, when cubin_size is non-NULL,
// compile it with NVRTC for sm_100a.  Used by the CPU tests and to inspect / commit the generated code.
int32_t bn_lw_spec_source(int32_t n_nodes, const int32_t* card, const int32_t* par_ptr, const int32_t* par_idx,
                          const int64_t* cpt_off, const double* cpt, const int8_t* evidence, const int32_t* query,
                          int32_t n_query, int32_t prune, char* out_src, int64_t cap, int64_t* out_len,
                          int64_t* cubin_size, int32_t* n_live) {
  BN_REQUIRE(card && par_ptr && cpt_off && cpt && out_len, BN_BAD_ARG, "bn_lw_spec_source: NULL argument");
  Net net;
  net.n = n_nodes;
  net.card.assign(card, card + n_nodes);
  net.par_ptr.assign(par_ptr, par_ptr + n_nodes + 1);
  net.par_idx.assign(par_idx, par_idx + par_ptr[n_nodes]);
  net.cpt_off.assign(cpt_off, cpt_off + n_nodes + 1);
  net.cpt.assign(cpt, cpt + cpt_off[n_nodes]);
  SampleProgram prg;
  BN_TRY(compile_program(net, evidence, false, 1, &prg));
  std::vector<int64_t> qstride;
  int64_t cells = 1;
  for (int j = 0; j < n_query; ++j) {
    BN_REQUIRE(query[j] >= 0 && query[j] < n_nodes && !(evidence && evidence[query[j]] >= 0), BN_BAD_ARG, "bad query");
    qstride.push_back(cells);
    cells *= card[query[j]];
  }
  SpecShape shp;
  uint32_t nl = 0;
  std::vector<char> ev_mask((size_t)n_nodes, 0);
  for (int i = 0; i < n_nodes; ++i) ev_mask[i] = (evidence && evidence[i] >= 0) ? 1 : 0;
  // n_query == 0 dumps the table sampler (rand(bn, n) / weighted rows) instead of the inference kernel
  const std::string src = gen_source(n_query == 0 ? SPEC_TABLE : SPEC_INFER, net, prg, ev_mask, query, (uint32_t)n_query, qstride,
                                     (uint32_t)cells, prune != 0 && n_query != 0, shp, &nl);
  if (n_live) *n_live = (int32_t)nl;
  *out_len = (int64_t)src.size();
  if (out_src && cap > 0) {
    const size_t m = src.size() < (size_t)(cap - 1) ? src.size() : (size_t)(cap - 1);
    memcpy(out_src, src.data(), m);
    out_src[m] = 0;
  }
  if (cubin_size) {  // through the process + disk cache, so the cubin can be inspected (cuobjdump -sass) afterwards
    BN_REQUIRE(nvrtc().ok, BN_UNSUPPORTED, "libnvrtc.so.12 could not be loaded (set BN_NVRTC_PATH)");
    std::shared_ptr<CubinEntry> ce = get_cubin(src, true);
    BN_REQUIRE(ce->state == CubinEntry::READY, BN_CUDA_ERR, "nvrtcCompileProgram failed: %.300s", ce->log.c_str());
    *cubin_size = (int64_t)ce->cubin.size();
  }
  return BN_OK;
}

}  // extern "C"
