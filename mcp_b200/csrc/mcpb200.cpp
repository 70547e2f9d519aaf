// mcpb200.cpp — C-ABI entry points of libmcpb200.so (include/mcpb200.h).
//
// Host runtime around the generated sm_100a kernels: NVRTC compilation with an on-disk cubin cache,
// lazy binding of the CUDA driver (the library must load on a GPU-less build box), per-device module /
// scratch / staging-buffer management, launches timed with CUDA events on the launching stream, and the
// host-pointer entry points that shard the θ batch over the selected devices (one host thread per
// device, no collective: instances are independent — SURVEY.md §8e).
#include <cuda.h>
#include <cuda_runtime.h>
#include <dlfcn.h>
#include <nvrtc.h>
#include <sys/stat.h>
#include <unistd.h>

#include <algorithm>
#include <atomic>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <map>
#include <memory>
#include <mutex>
#include <sstream>
#include <string>
#include <thread>
#include <vector>

#include "plan.h"

extern const char* mcpb200_kernel_template_source;  // kernel_template_embed.cpp (generated at build time)
extern "C" int mcpb200_static_fp64_peak(double* tflops_out, char* err, int errlen);
extern "C" int mcpb200_static_flush_l2(void* stream, char* err, int errlen);

namespace {

using mcpb200::Plan;

// Must match `struct SolveParams` / `struct SensParams` in kernel_template.cuh.
struct SolveParams {
  long long B;
  const double *theta, *x0, *y0, *s0;
  double *x_out, *y_out, *s_out, *kkt_out, *eps_out;
  int *outer_out, *status_out, *steps_out;
  double* scratch;
  double* state;
  unsigned long long* counters;
  int* deferred;
  double tol, tightening_rate, loosening_rate, min_stepsize;
  int max_inner, max_outer;
  int pass, step_budget;
};
struct SensParams {
  long long B;
  const double *theta, *x, *y, *s;
  double* dzdtheta;
  const double* zbar;
  double* thetabar;
  const double* theta_p;
  double* z_p;
  int* status_out;
  double* scratch;
  double* state;
  unsigned long long* counters;
  int P;
};

thread_local std::string g_global_error;

// ---- lazily bound driver API ---------------------------------------------------------------------------
struct Driver {
  CUresult (*ModuleLoadData)(CUmodule*, const void*) = nullptr;
  CUresult (*ModuleUnload)(CUmodule) = nullptr;
  CUresult (*ModuleGetFunction)(CUfunction*, CUmodule, const char*) = nullptr;
  CUresult (*FuncSetAttribute)(CUfunction, CUfunction_attribute, int) = nullptr;
  CUresult (*FuncGetAttribute)(int*, CUfunction_attribute, CUfunction) = nullptr;
  CUresult (*LaunchKernel)(CUfunction, unsigned, unsigned, unsigned, unsigned, unsigned, unsigned, unsigned, CUstream,
                           void**, void**) = nullptr;
  CUresult (*GetErrorString)(CUresult, const char**) = nullptr;
  bool ok = false;
  std::string err;
};

Driver& driver() {
  static Driver d;
  static std::once_flag once;
  std::call_once(once, [] {
    auto get = [&](const char* name, void** fn) {
      cudaDriverEntryPointQueryResult q;
      cudaError_t e = cudaGetDriverEntryPoint(name, fn, cudaEnableDefault, &q);
      if (e != cudaSuccess || q != cudaDriverEntryPointSuccess || !*fn) {
        d.err = std::string("CUDA driver entry point ") + name + " unavailable: " + cudaGetErrorString(e);
        cudaGetLastError();
        return false;
      }
      return true;
    };
    d.ok = get("cuModuleLoadData", (void**)&d.ModuleLoadData) && get("cuModuleUnload", (void**)&d.ModuleUnload) &&
           get("cuModuleGetFunction", (void**)&d.ModuleGetFunction) &&
           get("cuFuncSetAttribute", (void**)&d.FuncSetAttribute) && get("cuFuncGetAttribute", (void**)&d.FuncGetAttribute) &&
           get("cuLaunchKernel", (void**)&d.LaunchKernel) && get("cuGetErrorString", (void**)&d.GetErrorString);
  });
  return d;
}

std::string cu_err(CUresult r) {
  const char* s = nullptr;
  if (driver().GetErrorString) driver().GetErrorString(r, &s);
  return s ? s : "unknown CUDA driver error";
}

// ---- per-device state ---------------------------------------------------------------------------------------
struct DevBuf {
  void* p = nullptr;
  size_t cap = 0;
  int ensure(size_t bytes) {
    if (bytes <= cap) return 0;
    if (p) cudaFree(p);
    p = nullptr;
    cap = 0;
    if (cudaMalloc(&p, bytes) != cudaSuccess) return 1;
    cap = bytes;
    return 0;
  }
  void release() {
    if (p) cudaFree(p);
    p = nullptr;
    cap = 0;
  }
};

struct DeviceState {
  int dev = -1;
  CUmodule mod = nullptr;
  CUfunction f_solve = nullptr, f_sens = nullptr, f_adj = nullptr;
  int num_sms = 0, regs_solve = 0, regs_sens = 0;
  DevBuf scratch, state, counters, deferred, steps_tmp;
  cudaEvent_t ev0 = nullptr, ev1 = nullptr, ev_mid = nullptr, ev_h2d0 = nullptr, ev_h2d1 = nullptr, ev_d2h1 = nullptr;
  cudaStream_t stream = nullptr;
  // staging buffers of the host entry points
  DevBuf theta, x, y, s, kkt, eps, outer, status, steps, big0, big1, big2, big3;
  bool timed = false;
  bool two_pass = false;  // ev_mid was recorded between pass 0 and pass 1 of the last solve
  bool pending = false;   // ev1 marks the end of the last enqueued launch sequence on this device
  long long launches = 0;
};

}  // namespace

struct mcpb200_problem {
  Plan plan;
  std::vector<char> cubin;
  std::mutex mu;
  std::string err;
  bool compile_only = false;
  bool cache_hit = false;
  std::vector<int> devices{0};
  std::map<int, std::unique_ptr<DeviceState>> dev;
  mcpb200_timing timing{};
  std::vector<int> last_devs;  // devices used by the last call (for timing)
};

namespace {

// Restores the caller's current device on EVERY exit path of an entry point that hops over the shards' devices.
struct DeviceGuard {
  int cur = -1;
  DeviceGuard() {
    if (cudaGetDevice(&cur) != cudaSuccess) {
      cur = -1;
      cudaGetLastError();
    }
  }
  ~DeviceGuard() {
    }
};

// Page-locks caller-owned PAGEABLE host arrays for the duration of one call (cudaHostRegister) so that the
// cudaMemcpyAsync of a Julia / numpy caller's arrays runs at full PCIe speed and truly asynchronously, as it does for
// the pinned buffers of bench.py.  Already pinned / registered / managed / device memory is left alone.  A failed
// registration is not an error: the copy then takes CUDA's pageable path.  MCPB200_HOST_REGISTER=0 disables it.
struct HostPin {
  std::vector<void*> regs;
  static bool enabled() {   // MCPB200_HOST_REGISTER=0: never page-lock caller memory
    static const bool on = [] {
      const char* e = getenv("MCPB200_HOST_REGISTER");
      return !(e && atoi(e) == 0);
    }();
    return on;
  }
  void pin(const void* p, size_t bytes) {
    if (!p || bytes < (1u << 20) || !enabled()) return;   // small arrays: the registration costs more than it saves
    cudaPointerAttributes attr;
    if (cudaPointerGetAttributes(&attr, p) != cudaSuccess) {
      cudaGetLastError();
      return;
    }
    if (attr.type != cudaMemoryTypeUnregistered) return;
    if (cudaHostRegister(const_cast<void*>(p), bytes, cudaHostRegisterPortable) == cudaSuccess) regs.push_back(const_cast<void*>(p));
    else cudaGetLastError();
  }
  ~HostPin() {
    for (void* p : regs)
      if (cudaHostUnregister(p) != cudaSuccess) cudaGetLastError();
  }
};

bool is_pageable_host(const void* p) {
  cudaPointerAttributes attr;
  if (cudaPointerGetAttributes(&attr, p) != cudaSuccess) {
    cudaGetLastError();
    return true;
  }
  return attr.type == cudaMemoryTypeUnregistered;
}

// Device → host download of one shard's slice of a caller array, in two stages so that every page-locking call of a
// shard is made BEFORE its first copy is queued (a cudaHostRegister issued behind a queued copy waits for the stream).
// A PAGEABLE destination (a Julia or numpy caller's plain array) is page-locked — only the page-aligned INTERIOR of the
// slice, so that the registrations of neighbouring shards (one host thread per device, running concurrently, in the
// shadow of their kernels) never share a page — and copied in three parts: the interior asynchronously at full PCIe
// speed, the two ragged edges (< one page each) through CUDA's pageable path.  The registrations last for the call
// (unregistering costs ≈ 0.15 ms per MB: ≈ 200 ms for the bench batch's 1.4 GB — keeping them between calls would save
// that but is unsafe without owning the arrays' lifetime, so it is not done).
struct Download {
  char* dst = nullptr;
  const char* src = nullptr;
  size_t bytes = 0;
  uintptr_t i0 = 0, i1 = 0;   // registered interior [i0, i1), or empty
  void prepare(HostPin& pins, void* dst_, const void* src_, size_t bytes_) {
    constexpr size_t PAGE = 4096;
    dst = (char*)dst_;
    src = (const char*)src_;
    bytes = bytes_;
    i0 = i1 = 0;
    if (bytes < (4u << 20) || !HostPin::enabled() || !is_pageable_host(dst)) return;
    const uintptr_t a0 = (uintptr_t)dst, a1 = a0 + bytes;
    const uintptr_t b0 = (a0 + PAGE - 1) / PAGE * PAGE, b1 = a1 / PAGE * PAGE;
    if (b1 <= b0) return;
    if (cudaHostRegister((void*)b0, b1 - b0, cudaHostRegisterPortable) != cudaSuccess) {
      cudaGetLastError();
      return;
    }
    pins.regs.push_back((void*)b0);   // released when the shard's worker returns (after its stream sync)
    i0 = b0;
    i1 = b1;
  }
  cudaError_t issue(cudaStream_t sm) const {
    if (!bytes) return cudaSuccess;
    if (i1 <= i0) return cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToHost, sm);
    const uintptr_t a0 = (uintptr_t)dst, a1 = a0 + bytes;
    cudaError_t e;
    if ((e = cudaMemcpyAsync((void*)i0, src + (i0 - a0), i1 - i0, cudaMemcpyDeviceToHost, sm)) != cudaSuccess) return e;
    if (i0 > a0 && (e = cudaMemcpyAsync(dst, src, i0 - a0, cudaMemcpyDeviceToHost, sm)) != cudaSuccess) return e;
    if (a1 > i1 && (e = cudaMemcpyAsync((void*)i1, src + (i1 - a0), a1 - i1, cudaMemcpyDeviceToHost, sm)) != cudaSuccess) return e;
    return cudaSuccess;
  }
};

int set_err(mcpb200_problem* h, int code, const std::string& msg) {
  if (h) h->err = msg;
  g_global_error = msg;
  return code;
}

#define CUDA_TRY(h, call)                                                                             \
  do {                                                                                                \
    cudaError_t e_ = (call);                                                                          \
    if (e_ != cudaSuccess) {                                                                          \
      cudaGetLastError();                                                                             \
      return set_err(h, MCPB200_ERR_CUDA, std::string(#call) + ": " + cudaGetErrorString(e_));       \
    }                                                                                                 \
  } while (0)

#define CU_TRY(h, call)                                                                   \
  do {                                                                                    \
    CUresult r_ = (call);                                                                 \
    if (r_ != CUDA_SUCCESS) return set_err(h, MCPB200_ERR_CUDA, std::string(#call) + ": " + cu_err(r_)); \
  } while (0)

uint64_t fnv1a(const std::string& s) {
  uint64_t hsh = 1469598103934665603ULL;
  for (unsigned char c : s) {
    hsh ^= c;
    hsh *= 1099511628211ULL;
  }
  return hsh;
}

std::string cache_dir() {
  if (const char* e = getenv("MCPB200_CACHE_DIR")) return e;
  Dl_info info;
  if (dladdr((void*)&fnv1a, &info) && info.dli_fname) {
    std::string p(info.dli_fname);
    size_t slash = p.rfind('/');
    std::string dir = slash == std::string::npos ? "." : p.substr(0, slash);
    return dir + "/_kcache";
  }
  return "/tmp/mcpb200_kcache";
}

bool read_file(const std::string& path, std::vector<char>& out) {
  std::ifstream f(path, std::ios::binary);
  if (!f) return false;
  out.assign(std::istreambuf_iterator<char>(f), std::istreambuf_iterator<char>());
  return !out.empty();
}

void write_file(const std::string& path, const char* data, size_t n) {
  std::string tmp = path + ".tmp" + std::to_string((long long)getpid());
  {
    std::ofstream f(tmp, std::ios::binary);
    if (!f) return;
    f.write(data, (std::streamsize)n);
  }
  rename(tmp.c_str(), path.c_str());
}

// ---- nvJitLink, bound at run time ------------------------------------------------------------------------------
// The library is NOT linked: a host process that imported torch has already mapped torch's own libnvJitLink.so.12
// (an older minor version without the symbol versions this toolkit's header would bind).  The toolkit's copy is
// opened by absolute path into its own namespace and its entry points are looked up by name.
struct JitLink {
  using Handle = void*;
  int (*Create)(Handle*, uint32_t, const char**) = nullptr;
  int (*Destroy)(Handle*) = nullptr;
  int (*AddData)(Handle, int, const void*, size_t, const char*) = nullptr;
  int (*Complete)(Handle) = nullptr;
  int (*GetLinkedCubinSize)(Handle, size_t*) = nullptr;
  int (*GetLinkedCubin)(Handle, void*) = nullptr;
  int (*GetErrorLogSize)(Handle, size_t*) = nullptr;
  int (*GetErrorLog)(Handle, char*) = nullptr;
  bool ok = false;
  std::string err;
};
constexpr int kJitLinkInputCubin = 1;   // NVJITLINK_INPUT_CUBIN

JitLink& jitlink() {
  static JitLink j;
  static std::once_flag once;
  std::call_once(once, [] {
    void* lib = nullptr;
    std::vector<std::string> names;
    if (const char* e = getenv("CUDA_HOME")) names.push_back(std::string(e) + "/lib64/libnvJitLink.so.12");
    names.push_back("/usr/local/cuda/lib64/libnvJitLink.so.12");
    names.push_back("libnvJitLink.so.12");
    for (const std::string& n : names)
      if ((lib = dlopen(n.c_str(), RTLD_NOW | RTLD_LOCAL))) break;
    if (!lib) {
      j.err = "libnvJitLink.so.12 not found (needed to link the separately compiled evaluation units)";
      return;
    }
    auto sym = [&](const char* base) -> void* {
      for (int minor = 9; minor >= 0; --minor) {
        const std::string n = std::string("__") + base + "_12_" + std::to_string(minor);
        if (void* f = dlsym(lib, n.c_str())) return f;
      }
      return dlsym(lib, base);
    };
    j.Create = (decltype(j.Create))sym("nvJitLinkCreate");
    j.Destroy = (decltype(j.Destroy))sym("nvJitLinkDestroy");
    j.AddData = (decltype(j.AddData))sym("nvJitLinkAddData");
    j.Complete = (decltype(j.Complete))sym("nvJitLinkComplete");
    j.GetLinkedCubinSize = (decltype(j.GetLinkedCubinSize))sym("nvJitLinkGetLinkedCubinSize");
    j.GetLinkedCubin = (decltype(j.GetLinkedCubin))sym("nvJitLinkGetLinkedCubin");
    j.GetErrorLogSize = (decltype(j.GetErrorLogSize))sym("nvJitLinkGetErrorLogSize");
    j.GetErrorLog = (decltype(j.GetErrorLog))sym("nvJitLinkGetErrorLog");
    j.ok = j.Create && j.Destroy && j.AddData && j.Complete && j.GetLinkedCubinSize && j.GetLinkedCubin && j.GetErrorLogSize && j.GetErrorLog;
    if (!j.ok) j.err = "libnvJitLink.so.12 lacks an expected entry point";
  });
  return j;
}

// One NVRTC compilation to a cubin (relocatable when `rdc`).  Returns an empty string on success, else the log.
std::string nvrtc_to_cubin(const std::string& src, const std::string& name, bool rdc, bool lineinfo, std::vector<char>& cubin,
                           int maxreg = 0) {
  nvrtcProgram prog;
  if (nvrtcCreateProgram(&prog, src.c_str(), name.c_str(), 0, nullptr, nullptr) != NVRTC_SUCCESS) return "nvrtcCreateProgram failed";
  std::vector<const char*> opts = {"--gpu-architecture=sm_100a", "--std=c++17", "-default-device"};
  if (lineinfo) opts.push_back("-lineinfo");
  if (rdc) opts.push_back("--relocatable-device-code=true");
  const std::string mr = "--maxrregcount=" + std::to_string(maxreg);
  if (maxreg > 0) opts.push_back(mr.c_str());
  nvrtcResult r = nvrtcCompileProgram(prog, (int)opts.size(), opts.data());
  if (r != NVRTC_SUCCESS) {
    size_t n = 0;
    nvrtcGetProgramLogSize(prog, &n);
    std::string log(n, '\0');
    if (n) nvrtcGetProgramLog(prog, &log[0]);
    nvrtcDestroyProgram(&prog);
    if (log.size() > 4000) log.resize(4000);
    return std::string("NVRTC: ") + nvrtcGetErrorString(r) + "\n" + log;
  }
  size_t n = 0;
  nvrtcGetCUBINSize(prog, &n);
  cubin.resize(n);
  nvrtcGetCUBIN(prog, cubin.data());
  nvrtcDestroyProgram(&prog);
  return n ? std::string() : std::string("NVRTC produced an empty cubin");
}

constexpr uint32_t MCPB200_NO_CACHE_READ_INTERNAL = 0x40000000u;   // (internal) recompile and overwrite the cache entry

int compile_source(mcpb200_problem* h, uint32_t flags) {
  const std::string& src = h->plan.source;
  const std::vector<std::string>& units = h->plan.units;
  int nvrtc_major = 0, nvrtc_minor = 0;
  nvrtcVersion(&nvrtc_major, &nvrtc_minor);
  unsigned long long hsh = fnv1a(src + "|sm_100a|v" + std::to_string(MCPB200_VERSION) + "|nvrtc" + std::to_string(nvrtc_major) + "." +
                                 std::to_string(nvrtc_minor) + "|std=c++17,-default-device,-lineinfo");
  for (const std::string& u : units) hsh = hsh * 1099511628211ULL ^ fnv1a(u);
  char key[32];
  snprintf(key, sizeof key, "%016llx", hsh);
  const std::string dir = cache_dir();
  const std::string cu_path = dir + "/mcp_" + key + ".cu";
  const std::string bin_path = dir + "/mcp_" + key + ".cubin";
  const bool use_cache = !(flags & MCPB200_NO_CACHE);
  if (use_cache && !(flags & MCPB200_NO_CACHE_READ_INTERNAL) && read_file(bin_path, h->cubin)) {
    h->cache_hit = true;
    return MCPB200_OK;
  }
  if (use_cache) {
    mkdir(dir.c_str(), 0755);
    write_file(cu_path, src.data(), src.size());  // lets ncu --import-source map SASS to the generated source
  }
  if (units.empty()) {
    const std::string err = nvrtc_to_cubin(src, cu_path, false, true, h->cubin);
    if (!err.empty()) return set_err(h, MCPB200_ERR_COMPILE, err);
  } else {
    // Big problems: the evaluation parts are separate relocatable units, compiled concurrently (NVRTC is
    // thread-safe across programs) and linked with the kernels by nvJitLink.
    const size_t nu = units.size();
    std::vector<std::vector<char>> objs(nu + 1);
    std::vector<std::string> errs(nu + 1);
    std::atomic<size_t> next{0};
    unsigned nthreads = std::max(1u, std::min<unsigned>(std::thread::hardware_concurrency(), (unsigned)(nu + 1)));
    if (const char* e = getenv("MCPB200_COMPILE_THREADS")) nthreads = (unsigned)std::max(1, atoi(e));
    // separately compiled functions do not see the kernels' launch bounds: cap their registers at what the widest
    // kernel of this plan may use
    const Plan& P = h->plan;
    // (dense kernels v1/v2 are compiled for two CTAs per SM: __launch_bounds__(256, 2))
    const int dense_resident = P.dense_threads * ((P.dense_kernel == 1 || P.dense_kernel == 2) ? 2 : 1);
    const int threads = std::max({P.dense_kernel ? dense_resident : P.sub * P.ipc_solve * P.nwide, P.has_jt ? P.sub * P.ipc_sens : 0,
                                  P.has_adjoint ? P.sub * P.ipc_adj : 0, 32});
    const int maxreg = std::min(255, (65536 / threads) / 8 * 8);
    auto worker = [&] {
      for (;;) {
        const size_t i = next.fetch_add(1);
        if (i > nu) break;
        errs[i] = (i == 0) ? nvrtc_to_cubin(src, cu_path, true, true, objs[0])
                           : nvrtc_to_cubin(units[i - 1], cu_path + ".unit" + std::to_string(i - 1), true, false, objs[i], maxreg);  // (sources of the units are not kept: no line info)
      }
    };
    std::vector<std::thread> pool;
    for (unsigned t = 1; t < nthreads; ++t) pool.emplace_back(worker);
    worker();
    for (auto& t : pool) t.join();
    for (size_t i = 0; i <= nu; ++i)
      if (!errs[i].empty()) return set_err(h, MCPB200_ERR_COMPILE, "unit " + std::to_string(i) + ": " + errs[i]);
    JitLink& J = jitlink();
    if (!J.ok) return set_err(h, MCPB200_ERR_COMPILE, J.err);
    JitLink::Handle L = nullptr;
    const char* lopts[] = {"-arch=sm_100a", "-lineinfo"};
    if (J.Create(&L, 2, lopts) != 0) return set_err(h, MCPB200_ERR_COMPILE, "nvJitLinkCreate failed");
    auto link_fail = [&](const char* what) {
      size_t n = 0;
      std::string log;
      if (J.GetErrorLogSize(L, &n) == 0 && n) {
        log.resize(n);
        J.GetErrorLog(L, &log[0]);
      }
      J.Destroy(&L);
      if (log.size() > 4000) log.resize(4000);
      return set_err(h, MCPB200_ERR_COMPILE, std::string("nvJitLink: ") + what + "\n" + log);
    };
    for (size_t i = 0; i <= nu; ++i)
      if (J.AddData(L, kJitLinkInputCubin, objs[i].data(), objs[i].size(), ("unit" + std::to_string(i)).c_str()) != 0)
        return link_fail("add failed");
    if (J.Complete(L) != 0) return link_fail("link failed");
    size_t n = 0;
    if (J.GetLinkedCubinSize(L, &n) != 0 || n == 0) return link_fail("empty result");
    h->cubin.resize(n);
    J.GetLinkedCubin(L, h->cubin.data());
    J.Destroy(&L);
  }
  if (use_cache) write_file(bin_path, h->cubin.data(), h->cubin.size());
  return MCPB200_OK;
}

// Makes `dev` current and returns its state (module loaded, scratch allocated).
int device_state(mcpb200_problem* h, int dev, DeviceState** out) {
  if (h->compile_only)
    return set_err(h, MCPB200_ERR_CUDA, "handle was created with MCPB200_COMPILE_ONLY: no device work possible");
  CUDA_TRY(h, cudaSetDevice(dev));
  CUDA_TRY(h, cudaFree(0));  // force the primary context
  Driver& D = driver();
  if (!D.ok) return set_err(h, MCPB200_ERR_CUDA, D.err.empty() ? "CUDA driver unavailable" : D.err);
  auto it = h->dev.find(dev);
  if (it == h->dev.end() || !it->second) {
    // built in a local and inserted only after every step succeeded: a failure (not an sm_100 device, module load,
    // allocation) must not leave a null entry behind for mcpb200_destroy / mcpb200_get_timing to trip over
    auto st = std::make_unique<DeviceState>();
    st->dev = dev;
    cudaDeviceProp prop;
    CUDA_TRY(h, cudaGetDeviceProperties(&prop, dev));
    if (prop.major != 10)
      return set_err(h, MCPB200_ERR_CUDA, std::string("device '") + prop.name + "' is not sm_100: this library targets B200 only");
    st->num_sms = prop.multiProcessorCount;
    {
      CUresult r = D.ModuleLoadData(&st->mod, h->cubin.data());
      if (r != CUDA_SUCCESS && h->cache_hit) {
        // a corrupt or stale cubin in the on-disk cache: compile afresh once (which overwrites the cache entry)
        h->cache_hit = false;
        h->cubin.clear();
        const int rc = compile_source(h, MCPB200_NO_CACHE_READ_INTERNAL);
        if (rc) return rc;
        r = D.ModuleLoadData(&st->mod, h->cubin.data());
      }
      if (r != CUDA_SUCCESS) return set_err(h, MCPB200_ERR_CUDA, std::string("cuModuleLoadData: ") + cu_err(r));
    }
    CU_TRY(h, D.ModuleGetFunction(&st->f_solve, st->mod, "mcp_solve_kernel"));
    if (!h->plan.tiny_kernel)
      CU_TRY(h, D.FuncSetAttribute(st->f_solve, CU_FUNC_ATTRIBUTE_MAX_DYNAMIC_SHARED_SIZE_BYTES, (int)h->plan.smem_solve));
    if (const char* e = getenv("MCPB200_CARVEOUT"))   // tuning: shared-memory carve-out in percent (the rest of the 256 KB is L1)
      D.FuncSetAttribute(st->f_solve, CU_FUNC_ATTRIBUTE_PREFERRED_SHARED_MEMORY_CARVEOUT, atoi(e));
    D.FuncGetAttribute(&st->regs_solve, CU_FUNC_ATTRIBUTE_NUM_REGS, st->f_solve);
    if (h->plan.has_jt) {
      CU_TRY(h, D.ModuleGetFunction(&st->f_sens, st->mod, "mcp_sens_kernel"));
      CU_TRY(h, D.FuncSetAttribute(st->f_sens, CU_FUNC_ATTRIBUTE_MAX_DYNAMIC_SHARED_SIZE_BYTES, (int)h->plan.smem_sens));
      D.FuncGetAttribute(&st->regs_sens, CU_FUNC_ATTRIBUTE_NUM_REGS, st->f_sens);
      if (h->plan.has_adjoint) {
        CU_TRY(h, D.ModuleGetFunction(&st->f_adj, st->mod, "mcp_adj_kernel"));
        CU_TRY(h, D.FuncSetAttribute(st->f_adj, CU_FUNC_ATTRIBUTE_MAX_DYNAMIC_SHARED_SIZE_BYTES, (int)h->plan.smem_adj));
      }
    }
    CUDA_TRY(h, cudaEventCreate(&st->ev0));
    CUDA_TRY(h, cudaEventCreate(&st->ev1));
    CUDA_TRY(h, cudaEventCreate(&st->ev_mid));
    CUDA_TRY(h, cudaEventCreate(&st->ev_h2d0));
    CUDA_TRY(h, cudaEventCreate(&st->ev_h2d1));
    CUDA_TRY(h, cudaEventCreate(&st->ev_d2h1));
    CUDA_TRY(h, cudaStreamCreateWithFlags(&st->stream, cudaStreamNonBlocking));
    if (st->counters.ensure(64)) return set_err(h, MCPB200_ERR_CUDA, "cudaMalloc(counters) failed");
    it = h->dev.insert_or_assign(dev, std::move(st)).first;
  }
  *out = it->second.get();
  return MCPB200_OK;
}

int launch_solve(mcpb200_problem* h, DeviceState* st, SolveParams& p, cudaStream_t stream) {
  const Plan& P = h->plan;
  const long long ctas_needed = (p.B + P.ipc_solve - 1) / P.ipc_solve;
  const long long max_ctas = (long long)st->num_sms * (P.dense_kernel ? P.dense_ctas_per_sm : 1);
  unsigned grid = (unsigned)std::max<long long>(1, std::min<long long>(max_ctas, ctas_needed));
  unsigned block = P.dense_kernel ? (unsigned)P.dense_threads : (unsigned)(P.sub * P.ipc_solve * P.nwide);
  unsigned smem_bytes = (unsigned)P.smem_solve;
  if (P.tiny_kernel) {   // one thread per instance, no shared memory, not persistent
    block = 128;
    grid = (unsigned)std::max<long long>(1, (p.B + block - 1) / block);
    smem_bytes = 0;
  }
  const size_t scratch_bytes = (size_t)st->num_sms * P.ipc_solve * P.scratch_doubles_solve * 8;
  if (st->scratch.ensure(std::max(scratch_bytes, (size_t)st->num_sms * std::max(P.ipc_sens, P.ipc_adj) * P.scratch_doubles_sens * 8)))
    return set_err(h, MCPB200_ERR_CUDA, "cudaMalloc(scratch) failed");
  p.scratch = (double*)st->scratch.p;
  p.counters = (unsigned long long*)st->counters.p;
  p.state = nullptr;
  if (P.large_state || P.state_doubles_solve > 0) {   // (also the S_GLOBAL plans: `s` alone lives in the global block)
    if (st->state.ensure((size_t)st->num_sms * std::max({(size_t)P.ipc_solve * P.state_doubles_solve,
                                                         (size_t)P.ipc_sens * P.state_doubles_sens,
                                                         (size_t)P.ipc_adj * P.state_doubles_adj}) * 8 + 64))
      return set_err(h, MCPB200_ERR_CUDA, "cudaMalloc(state) failed");
    p.state = (double*)st->state.p;
  }
  // two-pass scheduling (kernel_template.cuh): pass 0 parks instances that exceed the step budget, pass 1
  // resumes them together.  MCPB200_PASS1_STEPS overrides the budget (0 disables the second pass).
  int budget = 2 * std::max(p.max_inner, 1) + 24;
  if (const char* e = getenv("MCPB200_PASS1_STEPS")) budget = atoi(e);
  // deferred list: the window kernels re-queue unfinished instances during pass 1 (slices of `budget` steps; at most
  // (max_outer·max_inner)/budget + 1 entries per instance), entries not yet produced read −1
  const bool requeue = !P.tiny_kernel && !P.dense_kernel && budget > 0;
  const size_t slices = requeue ? (size_t)std::max(1, p.max_outer) * (size_t)std::max(1, p.max_inner) / (size_t)budget + 2 : 1;
  const size_t deferred_bytes = (size_t)p.B * 4 * slices + 16;
  if (st->deferred.ensure(deferred_bytes)) return set_err(h, MCPB200_ERR_CUDA, "cudaMalloc(deferred) failed");
  p.deferred = (int*)st->deferred.p;
  if (!p.steps_out) {
    if (st->steps_tmp.ensure((size_t)p.B * 4 + 16)) return set_err(h, MCPB200_ERR_CUDA, "cudaMalloc(steps) failed");
    p.steps_out = (int*)st->steps_tmp.p;
  }
  p.step_budget = budget;
  // Scratch, queue counters, deferred list and state block are per (handle, device): a sequence enqueued on
  // another stream (e.g. torch's, by the _device entry points) must finish before this one may reuse them.
  if (st->pending) CUDA_TRY(h, cudaStreamWaitEvent(stream, st->ev1, 0));
  CUDA_TRY(h, cudaMemsetAsync(st->counters.p, 0, 64, stream));
  if (requeue) CUDA_TRY(h, cudaMemsetAsync(st->deferred.p, 0xff, deferred_bytes, stream));
  CUDA_TRY(h, cudaEventRecord(st->ev0, stream));
  void* args[] = {&p};
  p.pass = 0;
  CU_TRY(h, driver().LaunchKernel(st->f_solve, grid, 1, 1, block, 1, 1, smem_bytes, (CUstream)stream, args, nullptr));
  st->launches = 1;
  st->two_pass = budget > 0;
  if (budget > 0) {
    CUDA_TRY(h, cudaEventRecord(st->ev_mid, stream));
    p.pass = 1;
    if (const char* e = getenv("MCPB200_REQUEUE")) {   // A/B: 0 = pass 1 runs every deferred instance to its end in one go
      if (atoi(e) == 0) p.step_budget = 0;
      else p.step_budget = atoi(e);                    // or the slice length in Newton steps
    }
    CU_TRY(h, driver().LaunchKernel(st->f_solve, grid, 1, 1, block, 1, 1, smem_bytes, (CUstream)stream, args, nullptr));
    st->launches = 2;
  }
  CUDA_TRY(h, cudaEventRecord(st->ev1, stream));
  st->timed = true;
  st->pending = true;
  return MCPB200_OK;
}

int launch_sens(mcpb200_problem* h, DeviceState* st, SensParams& p, cudaStream_t stream) {
  const Plan& P = h->plan;
  // pullback only → the adjoint kernel (one transposed solve per instance, its own denser layout)
  const bool adj = P.has_adjoint && st->f_adj && p.zbar && p.thetabar && !p.dzdtheta && !p.z_p;
  const int ipc = adj ? P.ipc_adj : P.ipc_sens;
  const long long ctas_needed = (p.B + ipc - 1) / ipc;
  const unsigned grid = (unsigned)std::max<long long>(1, std::min<long long>(st->num_sms, ctas_needed));
  const size_t scratch_bytes = (size_t)st->num_sms * std::max((size_t)std::max(P.ipc_sens, P.ipc_adj) * P.scratch_doubles_sens,
                                                              (size_t)P.ipc_solve * P.scratch_doubles_solve) * 8;
  if (st->scratch.ensure(scratch_bytes)) return set_err(h, MCPB200_ERR_CUDA, "cudaMalloc(scratch) failed");
  p.scratch = (double*)st->scratch.p;
  p.counters = (unsigned long long*)st->counters.p;
  p.state = nullptr;
  if (P.large_state) {
    if (st->state.ensure((size_t)st->num_sms * std::max({(size_t)P.ipc_solve * P.state_doubles_solve,
                                                         (size_t)P.ipc_sens * P.state_doubles_sens,
                                                         (size_t)P.ipc_adj * P.state_doubles_adj}) * 8 + 64))
      return set_err(h, MCPB200_ERR_CUDA, "cudaMalloc(state) failed");
    p.state = (double*)st->state.p;
  }
  if (st->pending) CUDA_TRY(h, cudaStreamWaitEvent(stream, st->ev1, 0));
  CUDA_TRY(h, cudaMemsetAsync(st->counters.p, 0, 64, stream));
  CUDA_TRY(h, cudaEventRecord(st->ev0, stream));
  void* args[] = {&p};
  CU_TRY(h, driver().LaunchKernel(adj ? st->f_adj : st->f_sens, grid, 1, 1, (unsigned)(P.sub * ipc), 1, 1,
                                  (unsigned)(adj ? P.smem_adj : P.smem_sens), (CUstream)stream, args, nullptr));
  CUDA_TRY(h, cudaEventRecord(st->ev1, stream));
  st->timed = true;
  st->pending = true;
  st->launches = 1;
  return MCPB200_OK;
}

void fill_opts(SolveParams& p, const mcpb200_solver_opts* o) {
  mcpb200_solver_opts d;
  mcpb200_default_opts(&d);
  if (!o) o = &d;
  p.tol = o->tol;
  p.max_inner = o->max_inner_iters;
  p.max_outer = o->max_outer_iters;
  p.tightening_rate = o->tightening_rate;
  p.loosening_rate = o->loosening_rate;
  p.min_stepsize = o->min_stepsize;
}

struct Shard {
  int dev;
  int64_t begin, count;
};

std::vector<Shard> make_shards(const std::vector<int>& devs, int64_t B) {
  std::vector<Shard> out;
  const int64_t G = (int64_t)devs.size();
  for (int64_t g = 0; g < G; ++g) {
    int64_t b0 = B * g / G, b1 = B * (g + 1) / G;  // contiguous column blocks
    if (b1 > b0) out.push_back({devs[g], b0, b1 - b0});
  }
  return out;
}

}  // namespace

// =============================================================================================================
extern "C" {

const char* mcpb200_global_error(void) { return g_global_error.c_str(); }

const char* mcpb200_last_error(mcpb200_handle h) { return h ? h->err.c_str() : g_global_error.c_str(); }

void mcpb200_default_opts(mcpb200_solver_opts* o) {
  if (!o) return;
  o->tol = 1e-4;
  o->max_inner_iters = 20;
  o->max_outer_iters = 50;
  o->tightening_rate = 0.1;
  o->loosening_rate = 0.5;
  o->min_stepsize = 1e-4;
}

int mcpb200_create(const mcpb200_problem_desc* desc, uint32_t flags, mcpb200_handle* out) {
  if (!desc || !out) return set_err(nullptr, MCPB200_ERR_INVALID_ARGUMENT, "null argument");
  *out = nullptr;
  try {
    auto h = std::make_unique<mcpb200_problem>();
    h->compile_only = (flags & MCPB200_COMPILE_ONLY) != 0;
    int rc = mcpb200::build_plan(*desc, mcpb200_kernel_template_source, h->plan);
    if (rc != MCPB200_OK) return set_err(nullptr, rc, h->plan.error);
    rc = compile_source(h.get(), flags);
    if (rc != MCPB200_OK) return set_err(nullptr, rc, h->err);
    *out = h.release();
    return MCPB200_OK;
  } catch (const std::exception& e) {
    return set_err(nullptr, MCPB200_ERR_INTERNAL, std::string("exception: ") + e.what());
  } catch (...) {
    return set_err(nullptr, MCPB200_ERR_INTERNAL, "unknown exception");
  }
}

int mcpb200_destroy(mcpb200_handle h) {
  if (!h) return MCPB200_OK;
  DeviceGuard device_guard;
  for (auto& kv : h->dev) {
    DeviceState* st = kv.second.get();
    if (!st || cudaSetDevice(st->dev) != cudaSuccess) continue;
    for (DevBuf* b : {&st->scratch, &st->state, &st->counters, &st->deferred, &st->steps_tmp, &st->theta, &st->x, &st->y, &st->s, &st->kkt, &st->eps, &st->outer,
                      &st->status, &st->steps, &st->big0, &st->big1, &st->big2, &st->big3})
      b->release();
    for (cudaEvent_t e : {st->ev0, st->ev1, st->ev_mid, st->ev_h2d0, st->ev_h2d1, st->ev_d2h1})
      if (e) cudaEventDestroy(e);
    if (st->stream) cudaStreamDestroy(st->stream);
    if (st->mod && driver().ModuleUnload) driver().ModuleUnload(st->mod);
  }
  delete h;
  return MCPB200_OK;
}

int mcpb200_get_source(mcpb200_handle h, const char** src, int64_t* len) {
  if (!h || !src) return MCPB200_ERR_INVALID_ARGUMENT;
  *src = h->plan.source.c_str();
  if (len) *len = (int64_t)h->plan.source.size();
  return MCPB200_OK;
}

int mcpb200_get_info(mcpb200_handle h, mcpb200_info* info) {
  if (!h || !info) return MCPB200_ERR_INVALID_ARGUMENT;
  const Plan& P = h->plan;
  memset(info, 0, sizeof *info);
  info->nx = P.nx;
  info->ny = P.ny;
  info->ntheta = P.nt;
  info->n_reduced = P.N;
  info->kl = P.kl;
  info->ku = P.ku;
  info->window_rows = P.R;
  info->window_cols = P.WC;
  info->row_stride = P.WS1;
  info->n_jac_computed = (int)P.jv_nodes.size();
  info->n_jac_constant = P.n_const_entries;
  info->n_assembly_dests = (int)P.d_row.size();
  info->n_assembly_terms = (int)P.t_coef.size();
  info->threads_per_instance = P.tiny_kernel ? 1 : (P.dense_kernel ? P.dense_threads : P.sub * P.nwide);
  info->instances_per_cta = P.ipc_solve;
  info->ctas_per_sm = P.dense_kernel ? P.dense_ctas_per_sm : 1;
  info->smem_bytes_per_cta = (int)P.smem_solve;
  info->has_sensitivities = P.has_jt;
  info->cache_hit = h->cache_hit;
  info->flops_per_newton_step_band = P.flops_band;
  for (auto& kv : h->dev) {
    if (!kv.second) continue;
    info->regs_solve = kv.second->regs_solve;
    info->regs_sens = kv.second->regs_sens;
  }
  return MCPB200_OK;
}

int mcpb200_set_devices(mcpb200_handle h, const int32_t* ids, int32_t count) {
  if (!h || !ids || count <= 0) return set_err(h, MCPB200_ERR_INVALID_ARGUMENT, "set_devices: need ≥ 1 device id");
  std::lock_guard<std::mutex> lock(h->mu);
  h->devices.assign(ids, ids + count);
  return MCPB200_OK;
}

int mcpb200_get_timing(mcpb200_handle h, mcpb200_timing* t) {
  if (!h || !t) return MCPB200_ERR_INVALID_ARGUMENT;
  std::lock_guard<std::mutex> lock(h->mu);
  mcpb200_timing out = h->timing;
  out.kernel_ms = 0;
  out.launches = 0;
  out.newton_steps = 0;
  out.solved = 0;
  out.pass0_ms = 0;
  out.deferred = 0;
  DeviceGuard device_guard;
  for (int dev : h->last_devs) {
    auto it = h->dev.find(dev);
    if (it == h->dev.end() || !it->second || !it->second->timed) continue;
    DeviceState* st = it->second.get();
    CUDA_TRY(h, cudaSetDevice(dev));
    CUDA_TRY(h, cudaEventSynchronize(st->ev1));
    float ms = 0;
    CUDA_TRY(h, cudaEventElapsedTime(&ms, st->ev0, st->ev1));
    if ((double)ms >= out.kernel_ms) {  // max over devices; the pass split of that device
      out.kernel_ms = (double)ms;
      out.pass0_ms = (double)ms;
      if (st->two_pass && st->launches == 2) {
        float m0 = 0;
        CUDA_TRY(h, cudaEventElapsedTime(&m0, st->ev0, st->ev_mid));
        out.pass0_ms = (double)m0;
      }
    }
    unsigned long long c[4] = {0, 0, 0, 0};
    CUDA_TRY(h, cudaMemcpy(c, st->counters.p, sizeof c, cudaMemcpyDeviceToHost));
    out.newton_steps += (int64_t)c[1];
    out.solved += (int64_t)c[2];
    out.deferred += (int64_t)c[3];
    out.launches += st->launches;
  }
  *t = out;
  return MCPB200_OK;
}

// ---- device-pointer entry points -------------------------------------------------------------------------------
int mcpb200_solve_batched_device(mcpb200_handle h, int64_t B, const double* theta, const double* x0, const double* y0,
                                 const double* s0, const mcpb200_solver_opts* opts, double* x_out, double* y_out,
                                 double* s_out, double* kkt_out, double* eps_out, int32_t* outer_out, int32_t* status_out,
                                 int32_t* steps_out, void* stream) {
  if (!h) return set_err(nullptr, MCPB200_ERR_INVALID_ARGUMENT, "null handle");
  std::lock_guard<std::mutex> lock(h->mu);
  if (B < 0 || !x_out || !y_out || !s_out || !kkt_out || !eps_out || !outer_out || !status_out || (h->plan.nt > 0 && !theta && B > 0))
    return set_err(h, MCPB200_ERR_INVALID_ARGUMENT, "solve_batched: null output/θ pointer or negative B");
  int dev = 0;
  CUDA_TRY(h, cudaGetDevice(&dev));
  DeviceState* st = nullptr;
  int rc = device_state(h, dev, &st);
  if (rc) return rc;
  h->last_devs = {dev};
  h->timing = mcpb200_timing{};
  if (B == 0) {
    st->timed = false;
    return MCPB200_OK;
  }
  SolveParams p{};
  p.B = B;
  p.theta = theta;
  p.x0 = x0;
  p.y0 = y0;
  p.s0 = s0;
  p.x_out = x_out;
  p.y_out = y_out;
  p.s_out = s_out;
  p.kkt_out = kkt_out;
  p.eps_out = eps_out;
  p.outer_out = outer_out;
  p.status_out = status_out;
  p.steps_out = steps_out;
  fill_opts(p, opts);
  return launch_solve(h, st, p, (cudaStream_t)stream);
}

int mcpb200_sensitivities_device(mcpb200_handle h, int64_t B, const double* theta, const double* x, const double* y,
                                 const double* s, const double* eps, double* dzdtheta_out, const double* zbar,
                                 double* thetabar_out, int32_t P, const double* theta_p, double* z_p_out,
                                 int32_t* sens_status_out, void* stream) {
  (void)eps;  // ∇F_z and ∇F_θ do not depend on ϵ; kept for signature parity with `_solve_jacobian_θ`
  if (!h) return set_err(nullptr, MCPB200_ERR_INVALID_ARGUMENT, "null handle");
  std::lock_guard<std::mutex> lock(h->mu);
  if (!h->plan.has_jt)
    return set_err(h, MCPB200_ERR_NO_SENSITIVITIES,
                   "Missing sensitivities. Set `compute_sensitivities = true` when constructing the PrimalDualMCP.");
  if (B < 0 || !x || !y || !s || (zbar && !thetabar_out) || (theta_p && (!z_p_out || P <= 0)))
    return set_err(h, MCPB200_ERR_INVALID_ARGUMENT, "sensitivities: inconsistent arguments");
  int dev = 0;
  CUDA_TRY(h, cudaGetDevice(&dev));
  DeviceState* st = nullptr;
  int rc = device_state(h, dev, &st);
  if (rc) return rc;
  h->last_devs = {dev};
  h->timing = mcpb200_timing{};
  if (B == 0) {
    st->timed = false;
    return MCPB200_OK;
  }
  SensParams p{};
  p.B = B;
  p.theta = theta;
  p.x = x;
  p.y = y;
  p.s = s;
  p.dzdtheta = dzdtheta_out;
  p.zbar = zbar;
  p.thetabar = zbar ? thetabar_out : nullptr;
  p.theta_p = theta_p;
  p.z_p = theta_p ? z_p_out : nullptr;
  p.status_out = sens_status_out;
  p.P = theta_p ? P : 0;
  return launch_sens(h, st, p, (cudaStream_t)stream);
}

// ---- host-pointer entry points ------------------------------------------------------------------------------------
int mcpb200_solve_batched(mcpb200_handle h, int64_t B, const double* theta, const double* x0, const double* y0,
                          const double* s0, const mcpb200_solver_opts* opts, double* x_out, double* y_out, double* s_out,
                          double* kkt_out, double* eps_out, int32_t* outer_out, int32_t* status_out, int32_t* steps_out) {
  if (!h) return set_err(nullptr, MCPB200_ERR_INVALID_ARGUMENT, "null handle");
  std::lock_guard<std::mutex> lock(h->mu);
  if (B < 0 || !x_out || !y_out || !s_out || !kkt_out || !eps_out || !outer_out || !status_out || (h->plan.nt > 0 && !theta && B > 0))
    return set_err(h, MCPB200_ERR_INVALID_ARGUMENT, "solve_batched: null output/θ pointer or negative B");
  h->timing = mcpb200_timing{};
  h->last_devs.clear();
  if (B == 0) return MCPB200_OK;
  const int nx = h->plan.nx, ny = h->plan.ny, nt = h->plan.nt;
  std::vector<Shard> shards = make_shards(h->devices, B);
  std::vector<int> rcs(shards.size(), 0);
  std::vector<std::string> errs(shards.size());
  std::vector<float> h2d(shards.size(), 0), d2h(shards.size(), 0);
  std::mutex state_mu;
  // Pageable caller arrays: the (small) inputs are page-locked for the duration of the call; the (much larger) outputs
  // are page-locked per shard by the shard's own host thread while its kernels run (download()).
  HostPin pins;
  {
    DeviceGuard g0;
    if (!shards.empty()) cudaSetDevice(shards[0].dev);
    pins.pin(theta, (size_t)B * nt * 8);
    pins.pin(x0, (size_t)B * nx * 8);
    pins.pin(y0, (size_t)B * ny * 8);
    pins.pin(s0, (size_t)B * ny * 8);
  }
  auto work = [&](size_t i) {
    const Shard sh = shards[i];
    DeviceState* st = nullptr;
    {
      std::lock_guard<std::mutex> l2(state_mu);
      rcs[i] = device_state(h, sh.dev, &st);
      if (rcs[i]) errs[i] = h->err;
    }
    if (rcs[i]) return;
    auto fail = [&](const char* what, cudaError_t e) {
      if (st && st->stream) cudaStreamSynchronize(st->stream);   // nothing may still be copying into memory about to be unregistered
      rcs[i] = MCPB200_ERR_CUDA;
      errs[i] = std::string(what) + ": " + cudaGetErrorString(e);
      cudaGetLastError();
    };
    cudaSetDevice(sh.dev);
    const size_t n = (size_t)sh.count;
    if (st->theta.ensure(std::max<size_t>(8, n * nt * 8)) || st->x.ensure(n * nx * 8) || st->y.ensure(std::max<size_t>(8, n * ny * 8)) ||
        st->s.ensure(std::max<size_t>(8, n * ny * 8)) || st->kkt.ensure(n * 8) || st->eps.ensure(n * 8) || st->outer.ensure(n * 4) ||
        st->status.ensure(n * 4) || st->steps.ensure(n * 4)) {
      rcs[i] = MCPB200_ERR_CUDA;
      errs[i] = "cudaMalloc of staging buffers failed (batch too large for device memory?)";
      return;
    }
    cudaStream_t sm = st->stream;
    cudaError_t e;
    cudaEventRecord(st->ev_h2d0, sm);
    if (nt && (e = cudaMemcpyAsync(st->theta.p, theta + sh.begin * nt, n * nt * 8, cudaMemcpyHostToDevice, sm))) return fail("H2D θ", e);
    if (x0 && (e = cudaMemcpyAsync(st->x.p, x0 + sh.begin * nx, n * nx * 8, cudaMemcpyHostToDevice, sm))) return fail("H2D x0", e);
    if (y0 && ny && (e = cudaMemcpyAsync(st->y.p, y0 + sh.begin * ny, n * ny * 8, cudaMemcpyHostToDevice, sm))) return fail("H2D y0", e);
    if (s0 && ny && (e = cudaMemcpyAsync(st->s.p, s0 + sh.begin * ny, n * ny * 8, cudaMemcpyHostToDevice, sm))) return fail("H2D s0", e);
    cudaEventRecord(st->ev_h2d1, sm);
    SolveParams p{};
    p.B = sh.count;
    p.theta = (const double*)st->theta.p;
    p.x0 = x0 ? (const double*)st->x.p : nullptr;  // in-place: the kernel reads an instance's x₀ before writing its x
    p.y0 = y0 ? (const double*)st->y.p : nullptr;
    p.s0 = s0 ? (const double*)st->s.p : nullptr;
    p.x_out = (double*)st->x.p;
    p.y_out = (double*)st->y.p;
    p.s_out = (double*)st->s.p;
    p.kkt_out = (double*)st->kkt.p;
    p.eps_out = (double*)st->eps.p;
    p.outer_out = (int*)st->outer.p;
    p.status_out = (int*)st->status.p;
    p.steps_out = (int*)st->steps.p;
    fill_opts(p, opts);
    {
      std::lock_guard<std::mutex> l2(state_mu);
      rcs[i] = launch_solve(h, st, p, sm);
      if (rcs[i]) errs[i] = h->err;
    }
    if (rcs[i]) return;
    HostPin out_pins;
    const bool dbg = getenv("MCPB200_DEBUG_TIMING") != nullptr;
    auto now = [] { return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count(); };
    const double t0 = now();
    Download dx, dy, ds;
    dx.prepare(out_pins, x_out + sh.begin * nx, st->x.p, n * nx * 8);
    if (ny) dy.prepare(out_pins, y_out + sh.begin * ny, st->y.p, n * ny * 8);
    if (ny) ds.prepare(out_pins, s_out + sh.begin * ny, st->s.p, n * ny * 8);
    const double t1 = now();
    if ((e = dx.issue(sm))) return fail("D2H x", e);
    if ((e = dy.issue(sm))) return fail("D2H y", e);
    if ((e = ds.issue(sm))) return fail("D2H s", e);
    if ((e = cudaMemcpyAsync(kkt_out + sh.begin, st->kkt.p, n * 8, cudaMemcpyDeviceToHost, sm))) return fail("D2H kkt", e);
    if ((e = cudaMemcpyAsync(eps_out + sh.begin, st->eps.p, n * 8, cudaMemcpyDeviceToHost, sm))) return fail("D2H eps", e);
    if ((e = cudaMemcpyAsync(outer_out + sh.begin, st->outer.p, n * 4, cudaMemcpyDeviceToHost, sm))) return fail("D2H outer", e);
    if ((e = cudaMemcpyAsync(status_out + sh.begin, st->status.p, n * 4, cudaMemcpyDeviceToHost, sm))) return fail("D2H status", e);
    if (steps_out && (e = cudaMemcpyAsync(steps_out + sh.begin, st->steps.p, n * 4, cudaMemcpyDeviceToHost, sm))) return fail("D2H steps", e);
    cudaEventRecord(st->ev_d2h1, sm);
    const double t2 = now();
    if ((e = cudaStreamSynchronize(sm))) return fail("solve kernel / stream sync", e);
    if (dbg) fprintf(stderr, "[mcpb200] dev %d: register %.1f ms, issue %.1f ms, sync %.1f ms\n", sh.dev, t1 - t0, t2 - t1, now() - t2);
    cudaEventElapsedTime(&h2d[i], st->ev_h2d0, st->ev_h2d1);
    cudaEventElapsedTime(&d2h[i], st->ev1, st->ev_d2h1);
  };
  DeviceGuard device_guard;
  if (shards.size() == 1) {
    work(0);
  } else {
    std::vector<std::thread> ts;
    for (size_t i = 0; i < shards.size(); ++i) ts.emplace_back(work, i);
    for (auto& t : ts) t.join();
  }
  for (size_t i = 0; i < shards.size(); ++i) {
    if (rcs[i]) return set_err(h, rcs[i], "device " + std::to_string(shards[i].dev) + ": " + errs[i]);
    h->last_devs.push_back(shards[i].dev);
    h->timing.h2d_ms = std::max(h->timing.h2d_ms, (double)h2d[i]);
    h->timing.d2h_ms = std::max(h->timing.d2h_ms, (double)d2h[i]);
  }
  return MCPB200_OK;
}

int mcpb200_sensitivities(mcpb200_handle h, int64_t B, const double* theta, const double* x, const double* y,
                          const double* s, const double* eps, double* dzdtheta_out, const double* zbar,
                          double* thetabar_out, int32_t P, const double* theta_p, double* z_p_out,
                          int32_t* sens_status_out) {
  (void)eps;
  if (!h) return set_err(nullptr, MCPB200_ERR_INVALID_ARGUMENT, "null handle");
  std::lock_guard<std::mutex> lock(h->mu);
  if (!h->plan.has_jt)
    return set_err(h, MCPB200_ERR_NO_SENSITIVITIES,
                   "Missing sensitivities. Set `compute_sensitivities = true` when constructing the PrimalDualMCP.");
  if (B < 0 || !x || !y || !s || (zbar && !thetabar_out) || (theta_p && (!z_p_out || P <= 0)) || (h->plan.nt > 0 && !theta && B > 0))
    return set_err(h, MCPB200_ERR_INVALID_ARGUMENT, "sensitivities: inconsistent arguments");
  h->timing = mcpb200_timing{};
  h->last_devs.clear();
  if (B == 0) return MCPB200_OK;
  const int nx = h->plan.nx, ny = h->plan.ny, nt = h->plan.nt, nz = nx + 2 * ny;
  if (!theta_p) P = 0;
  // single device per call is enough for the sensitivities host path; shard sequentially over the devices
  std::vector<Shard> shards = make_shards(h->devices, B);
  DeviceGuard device_guard;
  for (const Shard& sh : shards) {
    DeviceState* st = nullptr;
    int rc = device_state(h, sh.dev, &st);
    if (rc) return rc;
    const size_t n = (size_t)sh.count;
    if (st->theta.ensure(std::max<size_t>(8, n * nt * 8)) || st->x.ensure(n * nx * 8) || st->y.ensure(std::max<size_t>(8, n * ny * 8)) ||
        st->s.ensure(std::max<size_t>(8, n * ny * 8)) || st->status.ensure(n * 4) ||
        (dzdtheta_out && st->big0.ensure(n * nz * nt * 8)) || (zbar && (st->big1.ensure(n * nz * 8) || st->kkt.ensure(std::max<size_t>(8, n * nt * 8)))) ||
        (P && (st->big2.ensure(n * nt * P * 8) || st->big3.ensure(n * nz * P * 8))))
      return set_err(h, MCPB200_ERR_CUDA, "cudaMalloc of staging buffers failed");
    cudaStream_t sm = st->stream;
    CUDA_TRY(h, cudaMemcpyAsync(st->theta.p, theta + sh.begin * nt, n * nt * 8, cudaMemcpyHostToDevice, sm));
    CUDA_TRY(h, cudaMemcpyAsync(st->x.p, x + sh.begin * nx, n * nx * 8, cudaMemcpyHostToDevice, sm));
    CUDA_TRY(h, cudaMemcpyAsync(st->y.p, y + sh.begin * ny, n * ny * 8, cudaMemcpyHostToDevice, sm));
    CUDA_TRY(h, cudaMemcpyAsync(st->s.p, s + sh.begin * ny, n * ny * 8, cudaMemcpyHostToDevice, sm));
    if (zbar) CUDA_TRY(h, cudaMemcpyAsync(st->big1.p, zbar + sh.begin * nz, n * nz * 8, cudaMemcpyHostToDevice, sm));
    if (P) CUDA_TRY(h, cudaMemcpyAsync(st->big2.p, theta_p + sh.begin * nt * P, n * nt * P * 8, cudaMemcpyHostToDevice, sm));
    SensParams p{};
    p.B = sh.count;
    p.theta = (const double*)st->theta.p;
    p.x = (const double*)st->x.p;
    p.y = (const double*)st->y.p;
    p.s = (const double*)st->s.p;
    p.dzdtheta = dzdtheta_out ? (double*)st->big0.p : nullptr;
    p.zbar = zbar ? (const double*)st->big1.p : nullptr;
    p.thetabar = zbar ? (double*)st->kkt.p : nullptr;
    p.theta_p = P ? (const double*)st->big2.p : nullptr;
    p.z_p = P ? (double*)st->big3.p : nullptr;
    p.status_out = (int*)st->status.p;
    p.P = P;
    rc = launch_sens(h, st, p, sm);
    if (rc) return rc;
    if (dzdtheta_out) CUDA_TRY(h, cudaMemcpyAsync(dzdtheta_out + sh.begin * nz * nt, st->big0.p, n * nz * nt * 8, cudaMemcpyDeviceToHost, sm));
    if (zbar) CUDA_TRY(h, cudaMemcpyAsync(thetabar_out + sh.begin * nt, st->kkt.p, n * nt * 8, cudaMemcpyDeviceToHost, sm));
    if (P) CUDA_TRY(h, cudaMemcpyAsync(z_p_out + sh.begin * nz * P, st->big3.p, n * nz * P * 8, cudaMemcpyDeviceToHost, sm));
    if (sens_status_out) CUDA_TRY(h, cudaMemcpyAsync(sens_status_out + sh.begin, st->status.p, n * 4, cudaMemcpyDeviceToHost, sm));
    CUDA_TRY(h, cudaStreamSynchronize(sm));
    h->last_devs.push_back(sh.dev);
  }
  return MCPB200_OK;
}

int mcpb200_measure_fp64_peak(double* tflops_out) {
  char err[256] = {0};
  int rc = mcpb200_static_fp64_peak(tflops_out, err, sizeof err);
  if (rc) return set_err(nullptr, MCPB200_ERR_CUDA, err);
  return MCPB200_OK;
}

int mcpb200_flush_l2(void* stream) {
  char err[256] = {0};
  int rc = mcpb200_static_flush_l2(stream, err, sizeof err);
  if (rc) return set_err(nullptr, MCPB200_ERR_CUDA, err);
  return MCPB200_OK;
}

}  // extern "C"
