// Run-time compiler and cubin cache of the fused tier.
//
// A generated kernel (codegen.cpp) is compiled for sm_100a by NVRTC -- loaded with dlopen, so
// the library itself has no link-time dependency on it -- and loaded with the CUDA runtime's
// library API (cudaLibraryLoadData / cudaLibraryGetKernel).  Compiled cubins are cached in
// memory and on disk, keyed by a hash of the generated source and of the headers it includes;
// the cache directory lives next to this library (csrc/build/fused, override with
// PYSTL_B200_CACHE), so cubins compiled at build time travel with the tree and a formula is
// compiled at most once per machine.  Compiling needs no GPU.
#include <cuda_runtime.h>
#include <dlfcn.h>
#include <sys/stat.h>
#include <unistd.h>

#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <map>
#include <mutex>
#include <sstream>
#include <string>
#include <vector>

#include "../../include/pystl_b200.h"
#include "fused_host.h"

namespace pystl {
namespace {

// ---- the slice of the NVRTC API we use (nvrtc.h is not needed at build time)
typedef struct _nvrtcProgram* nvrtcProgram;
typedef int nvrtcResult;
struct Nvrtc {
  void* h = nullptr;
  nvrtcResult (*CreateProgram)(nvrtcProgram*, const char*, const char*, int, const char* const*, const char* const*) = nullptr;
  nvrtcResult (*DestroyProgram)(nvrtcProgram*) = nullptr;
  nvrtcResult (*CompileProgram)(nvrtcProgram, int, const char* const*) = nullptr;
  nvrtcResult (*GetCUBINSize)(nvrtcProgram, size_t*) = nullptr;
  nvrtcResult (*GetCUBIN)(nvrtcProgram, char*) = nullptr;
  nvrtcResult (*GetProgramLogSize)(nvrtcProgram, size_t*) = nullptr;
  nvrtcResult (*GetProgramLog)(nvrtcProgram, char*) = nullptr;
  const char* (*GetErrorString)(nvrtcResult) = nullptr;
  std::string why;
  bool load() {
    if (h) return true;
    const char* names[] = {getenv("PYSTL_B200_NVRTC"), "libnvrtc.so.12", "/usr/local/cuda/lib64/libnvrtc.so.12",
                           "/usr/local/cuda/lib64/libnvrtc.so", "libnvrtc.so"};
    for (const char* n : names) {
      if (!n || !*n) continue;
      h = dlopen(n, RTLD_NOW | RTLD_LOCAL);
      if (h) break;
    }
    if (!h) {
      why = std::string("cannot load libnvrtc: ") + (dlerror() ? dlerror() : "not found");
      return false;
    }
#define PYSTL_SYM(field, name)                                          \
  field = reinterpret_cast<decltype(field)>(dlsym(h, name));            \
  if (!field) {                                                         \
    why = std::string("libnvrtc lacks ") + name;                        \
    h = nullptr;                                                        \
    return false;                                                       \
  }
    PYSTL_SYM(CreateProgram, "nvrtcCreateProgram")
    PYSTL_SYM(DestroyProgram, "nvrtcDestroyProgram")
    PYSTL_SYM(CompileProgram, "nvrtcCompileProgram")
    PYSTL_SYM(GetCUBINSize, "nvrtcGetCUBINSize")
    PYSTL_SYM(GetCUBIN, "nvrtcGetCUBIN")
    PYSTL_SYM(GetProgramLogSize, "nvrtcGetProgramLogSize")
    PYSTL_SYM(GetProgramLog, "nvrtcGetProgramLog")
    PYSTL_SYM(GetErrorString, "nvrtcGetErrorString")
#undef PYSTL_SYM
    return true;
  }
};

std::mutex g_mu;
Nvrtc g_nvrtc;
std::map<std::string, FusedKernel> g_loaded;  // key -> kernel
std::string g_src_dir, g_cache_dir, g_hdr_hash;

uint64_t fnv1a(const std::string& s, uint64_t h = 1469598103934665603ull) {
  for (unsigned char ch : s) {
    h ^= ch;
    h *= 1099511628211ull;
  }
  return h;
}

std::string read_file(const std::string& path) {
  std::ifstream f(path, std::ios::binary);
  if (!f) return std::string();
  std::stringstream ss;
  ss << f.rdbuf();
  return ss.str();
}

void mkdirs(const std::string& path) {
  std::string cur;
  for (size_t i = 0; i <= path.size(); ++i) {
    if (i == path.size() || path[i] == '/') {
      if (!cur.empty()) mkdir(cur.c_str(), 0777);
    }
    if (i < path.size()) cur += path[i];
  }
}

// directory of this shared object = the csrc directory holding the headers
void locate() {
  if (!g_src_dir.empty()) return;
  const char* e = getenv("PYSTL_B200_SRC");
  if (e && *e) {
    g_src_dir = e;
  } else {
    Dl_info info;
    if (dladdr(reinterpret_cast<void*>(&locate), &info) && info.dli_fname) {
      std::string p = info.dli_fname;
      const size_t k = p.rfind('/');
      g_src_dir = k == std::string::npos ? "." : p.substr(0, k);
    } else {
      g_src_dir = ".";
    }
  }
  const char* c = getenv("PYSTL_B200_CACHE");
  g_cache_dir = (c && *c) ? std::string(c) : g_src_dir + "/build/fused";
  uint64_t h = 1469598103934665603ull;
  for (const char* name : {"fused_rt.cuh", "tile_eval.h", "tile_bwd.h", "program.h", "fused_host.h", "../../include/pystl_b200.h"})
    h = fnv1a(read_file(g_src_dir + "/" + name), h);
  char buf[32];
  snprintf(buf, sizeof(buf), "%016llx", (unsigned long long)h);
  g_hdr_hash = buf;
}

std::string key_of(const FusedSource& src) {
  locate();
  char buf[40];
  const char* extra = getenv("PYSTL_B200_NVRTC_FLAGS");
  snprintf(buf, sizeof(buf), "%016llx",
           (unsigned long long)fnv1a(src.source + "|" + g_hdr_hash + "|sm_100a|v1|" + (extra ? extra : "")));
  return buf;
}

int fail(char* msg, int msg_len, int code, const std::string& what) {
  if (msg && msg_len > 0) snprintf(msg, msg_len, "%s", what.c_str());
  return code;
}

// Next to every <key>.cubin (a plain cubin: cuobjdump / nvdisasm read it as it is) lies <key>.meta = [magic "PYSTLCB2"]
// [u64 source length][u64 second hash of the source (FNV-1a, other seed)][u64 cubin length]: a truncated or foreign file,
// or another formula's kernel under a colliding 64-bit key, is recognised and recompiled instead of being loaded.
constexpr char kCacheMagic[8] = {'P', 'Y', 'S', 'T', 'L', 'C', 'B', '2'};
struct CacheHeader {
  char magic[8];
  uint64_t src_len, src_hash2, cubin_len;
};
uint64_t second_hash(const std::string& s) { return fnv1a(s, 0x9e3779b97f4a7c15ull); }
bool cache_entry_ok(const std::string& meta, const std::string& cubin, const FusedSource& src) {
  if (meta.size() != sizeof(CacheHeader) || cubin.empty()) return false;
  CacheHeader h;
  memcpy(&h, meta.data(), sizeof(h));
  return memcmp(h.magic, kCacheMagic, 8) == 0 && h.src_len == src.source.size() && h.src_hash2 == second_hash(src.source) &&
         h.cubin_len == cubin.size();
}

// source -> cubin bytes (disk cache, else NVRTC)
int get_cubin(const FusedSource& src, const std::string& key, std::string* cubin, char* msg, int msg_len, bool ignore_cache = false) {
  const std::string path = g_cache_dir + "/" + key + ".cubin";
  const std::string mpath = g_cache_dir + "/" + key + ".meta";
  if (!ignore_cache) {
    *cubin = read_file(path);
    if (!cubin->empty()) {
      if (cache_entry_ok(read_file(mpath), *cubin, src)) return PYSTL_OK;
      unlink(path.c_str());  // stale, truncated or colliding entry: recompile (a read-only directory just recompiles every time)
      unlink(mpath.c_str());
    }
  }
  if (!g_nvrtc.load()) return fail(msg, msg_len, PYSTL_E_UNSUPPORTED, g_nvrtc.why);
  nvrtcProgram prog = nullptr;
  nvrtcResult r = g_nvrtc.CreateProgram(&prog, src.source.c_str(), "pystl_fused.cu", 0, nullptr, nullptr);
  if (r) return fail(msg, msg_len, PYSTL_E_CUDA, std::string("nvrtcCreateProgram: ") + g_nvrtc.GetErrorString(r));
  const std::string inc = "-I" + g_src_dir;
  // PYSTL_B200_NVRTC_FLAGS=-lineinfo for profiling builds (part of the cache key)
  std::vector<const char*> opts = {"--gpu-architecture=sm_100a", "-std=c++17", "-default-device", inc.c_str()};
  const char* extra = getenv("PYSTL_B200_NVRTC_FLAGS");
  std::vector<std::string> extras;
  if (extra && *extra) {
    std::stringstream ss(extra);
    std::string tok;
    while (ss >> tok) extras.push_back(tok);
    for (auto& t : extras) opts.push_back(t.c_str());
  }
  r = g_nvrtc.CompileProgram(prog, (int)opts.size(), opts.data());
  if (r) {
    size_t n = 0;
    g_nvrtc.GetProgramLogSize(prog, &n);
    std::string log(n, '\0');
    if (n) g_nvrtc.GetProgramLog(prog, &log[0]);
    g_nvrtc.DestroyProgram(&prog);
    const char* dump = getenv("PYSTL_B200_DUMP_FAILED");
    if (dump && *dump) {
      std::ofstream f(dump);
      f << src.source << "\n/*\n" << log << "\n*/\n";
    }
    return fail(msg, msg_len, PYSTL_E_CUDA, "nvrtc: " + log.substr(0, 400));
  }
  size_t n = 0;
  g_nvrtc.GetCUBINSize(prog, &n);
  cubin->assign(n, '\0');
  g_nvrtc.GetCUBIN(prog, &(*cubin)[0]);
  g_nvrtc.DestroyProgram(&prog);
  // publish atomically; failure to write the cache is not an error
  mkdirs(g_cache_dir);
  char tmp[64];
  snprintf(tmp, sizeof(tmp), ".%d.tmp", (int)getpid());
  const std::string tpath = path + tmp;
  {
    std::ofstream f(tpath, std::ios::binary);
    if (f) {
      f.write(cubin->data(), (std::streamsize)cubin->size());
      f.close();
      CacheHeader h;
      memcpy(h.magic, kCacheMagic, 8);
      h.src_len = src.source.size();
      h.src_hash2 = second_hash(src.source);
      h.cubin_len = cubin->size();
      const std::string tmeta = mpath + tmp;
      std::ofstream fm(tmeta, std::ios::binary);
      fm.write(reinterpret_cast<const char*>(&h), sizeof(h));
      fm.close();
      // the cubin first, then its descriptor: a reader that finds no (or a stale) descriptor recompiles
      if (rename(tpath.c_str(), path.c_str()) != 0) unlink(tpath.c_str());
      if (!fm || rename(tmeta.c_str(), mpath.c_str()) != 0) unlink(tmeta.c_str());
    }
  }
  if (getenv("PYSTL_B200_KEEP_SOURCE")) {
    std::ofstream f(g_cache_dir + "/" + key + ".cu");
    f << src.source;
  }
  return PYSTL_OK;
}

}  // namespace

const char* fused_cache_dir() {
  std::lock_guard<std::mutex> lk(g_mu);
  locate();
  return g_cache_dir.c_str();
}

int fused_jit_precompile(const FusedSource& src, char* msg, int msg_len) {
  std::lock_guard<std::mutex> lk(g_mu);
  const std::string key = key_of(src);
  std::string cubin;
  return get_cubin(src, key, &cubin, msg, msg_len);
}

int fused_jit_get(const FusedSource& src, FusedKernel* out, char* msg, int msg_len) {
  std::lock_guard<std::mutex> lk(g_mu);
  const std::string key = key_of(src);
  int dev = 0;
  cudaGetDevice(&dev);
  const std::string lkey = key + "@" + std::to_string(dev);
  auto it = g_loaded.find(lkey);
  if (it != g_loaded.end()) {
    *out = it->second;
    return PYSTL_OK;
  }
  std::string cubin;
  int rc = get_cubin(src, key, &cubin, msg, msg_len);
  if (rc) return rc;
  cudaLibrary_t lib = nullptr;
  cudaError_t e = cudaLibraryLoadData(&lib, cubin.data(), nullptr, nullptr, 0, nullptr, nullptr, 0);
  if (e != cudaSuccess) {  // a cached image the driver rejects (another toolkit, a damaged file): compile once more, past the cache
    cudaGetLastError();
    rc = get_cubin(src, key, &cubin, msg, msg_len, true);
    if (rc) return rc;
    e = cudaLibraryLoadData(&lib, cubin.data(), nullptr, nullptr, 0, nullptr, nullptr, 0);
  }
  if (e != cudaSuccess) return fail(msg, msg_len, PYSTL_E_CUDA, std::string("cudaLibraryLoadData: ") + cudaGetErrorString(e));
  cudaKernel_t k = nullptr;
  e = cudaLibraryGetKernel(&k, lib, src.name.c_str());
  if (e != cudaSuccess) return fail(msg, msg_len, PYSTL_E_CUDA, std::string("cudaLibraryGetKernel: ") + cudaGetErrorString(e));
  e = cudaFuncSetAttribute(reinterpret_cast<const void*>(k), cudaFuncAttributeMaxDynamicSharedMemorySize, src.smem_floats * 4);
  if (e != cudaSuccess) return fail(msg, msg_len, PYSTL_E_CUDA, std::string("cudaFuncSetAttribute: ") + cudaGetErrorString(e));
  FusedKernel fk;
  fk.kernel = reinterpret_cast<void*>(k);
  cudaFuncAttributes attr;
  if (cudaFuncGetAttributes(&attr, reinterpret_cast<const void*>(k)) == cudaSuccess) fk.regs = attr.numRegs;
  g_loaded[lkey] = fk;
  *out = fk;
  return PYSTL_OK;
}

}  // namespace pystl
