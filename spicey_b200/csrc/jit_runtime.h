// NVRTC, loaded lazily with dlopen so that the library itself has no link-time dependency on it: compiles the
// kernels the generators (sparse_codegen.h, tran_codegen.h) write for one netlist to an sm_100a cubin.
//
// Compiled cubins are kept on disk, keyed by a hash of the source text (which contains the netlist's structure,
// constants and launch shape) and of the NVRTC version: a topology is compiled once per machine, not once per
// process (cfg 2: ~4 s of compile against 0.66 ms per sweep).  Directory: $SPICEY_CACHE_DIR, else
// $XDG_CACHE_HOME/spicey_b200, else ~/.cache/spicey_b200; SPICEY_CACHE_DIR=off disables it.
#pragma once
#include <dlfcn.h>
#include <sys/stat.h>
#include <unistd.h>

#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

namespace spicey {
namespace host {

struct Nvrtc {
  typedef int (*create_t)(void**, const char*, const char*, int, const char* const*, const char* const*);
  typedef int (*compile_t)(void*, int, const char* const*);
  typedef int (*size_t_fn)(void*, size_t*);
  typedef int (*get_t)(void*, char*);
  typedef int (*destroy_t)(void**);
  void* lib = nullptr;
  create_t create = nullptr; compile_t compile = nullptr; size_t_fn cubin_size = nullptr; get_t cubin = nullptr;
  size_t_fn log_size = nullptr; get_t log = nullptr; destroy_t destroy = nullptr;
  int (*version)(int*, int*) = nullptr;
  bool ok = false;
  Nvrtc() {
    const char* names[] = {"libnvrtc.so.12", "libnvrtc.so", "/usr/local/cuda/lib64/libnvrtc.so.12", "/usr/local/cuda/lib64/libnvrtc.so"};
    for (const char* n : names) if ((lib = dlopen(n, RTLD_NOW | RTLD_LOCAL))) break;
    if (!lib) return;
    create = (create_t)dlsym(lib, "nvrtcCreateProgram");
    compile = (compile_t)dlsym(lib, "nvrtcCompileProgram");
    cubin_size = (size_t_fn)dlsym(lib, "nvrtcGetCUBINSize");
    cubin = (get_t)dlsym(lib, "nvrtcGetCUBIN");
    log_size = (size_t_fn)dlsym(lib, "nvrtcGetProgramLogSize");
    log = (get_t)dlsym(lib, "nvrtcGetProgramLog");
    destroy = (destroy_t)dlsym(lib, "nvrtcDestroyProgram");
    version = (int (*)(int*, int*))dlsym(lib, "nvrtcVersion");
    ok = create && compile && cubin_size && cubin && log_size && log && destroy;
  }
};
inline Nvrtc& nvrtc() { static Nvrtc n; return n; }

// source -> cubin (sm_100a).  Returns false with `why` filled when NVRTC is missing or the compile fails.
inline bool jit_compile(const std::string& src, std::vector<char>& cubin, std::string& why) {
  Nvrtc& N = nvrtc();
  if (!N.ok) { why = "libnvrtc not available"; return false; }
  void* prog = nullptr;
  if (N.create(&prog, src.c_str(), "spicey_sparse_jit.cu", 0, nullptr, nullptr) != 0) { why = "nvrtcCreateProgram failed"; return false; }
  const char* opts[] = {"--gpu-architecture=sm_100a", "-std=c++17", "-lineinfo", "-default-device"};
  int rc = N.compile(prog, 4, opts);
  if (rc != 0) {
    size_t n = 0;
    N.log_size(prog, &n);
    std::string lg(n, '\0');
    if (n) N.log(prog, &lg[0]);
    why = "nvrtcCompileProgram failed: " + lg.substr(0, 2000);
    N.destroy(&prog);
    return false;
  }
  size_t n = 0;
  N.cubin_size(prog, &n);
  cubin.resize(n);
  N.cubin(prog, cubin.data());
  N.destroy(&prog);
  return n > 0;
}

inline std::string jit_cache_dir() {
  if (const char* e = getenv("SPICEY_CACHE_DIR")) return std::string(e) == "off" ? std::string() : std::string(e);
  if (const char* e = getenv("XDG_CACHE_HOME")) return std::string(e) + "/spicey_b200";
  if (const char* e = getenv("HOME")) return std::string(e) + "/.cache/spicey_b200";
  return std::string();
}

// Path of the cache entry of one source text (empty when the cache is off).
inline std::string jit_cache_path(const std::string& src) {
  const std::string dir = jit_cache_dir();
  if (dir.empty()) return std::string();
  unsigned long long h = 1469598103934665603ull;
  auto mix = [&](const void* p, size_t n) { const unsigned char* q = (const unsigned char*)p; for (size_t i = 0; i < n; ++i) { h ^= q[i]; h *= 1099511628211ull; } };
  mix(src.data(), src.size());
  int ver[2] = {0, 0};
  if (nvrtc().ok && nvrtc().version) nvrtc().version(&ver[0], &ver[1]);
  mix(ver, sizeof ver);
  char name[64];
  snprintf(name, sizeof name, "/%016llx_%zu.sm_100a.jit", h, src.size());
  return dir + name;
}

// Drops the cache entry of `src` (a cubin that fails to load is evicted and compiled again once).
inline void jit_cache_evict(const std::string& src) {
  const std::string path = jit_cache_path(src);
  if (!path.empty()) remove(path.c_str());
}

// jit_compile through the on-disk cache.  `from_cache` (optional) tells which way it went.
// Entry layout: "SPCYJIT1" | u64 source bytes | u64 cubin bytes | source text | cubin.  The file name is only a
// 64-bit hash: a hit is accepted when the stored source text equals `src` byte for byte and the sizes add up, so a
// hash collision, a truncated file or a foreign file is a miss (and is overwritten), never a wrong kernel.
inline bool jit_compile_cached(const std::string& src, std::vector<char>& cubin, std::string& why, bool* from_cache = nullptr) {
  if (from_cache) *from_cache = false;
  const std::string path = jit_cache_path(src);
  static const char kMagic[8] = {'S', 'P', 'C', 'Y', 'J', 'I', 'T', '1'};
  if (!path.empty()) {
    if (FILE* f = fopen(path.c_str(), "rb")) {
      char magic[8];
      unsigned long long ns = 0, nc = 0;
      bool got = fread(magic, 1, 8, f) == 8 && memcmp(magic, kMagic, 8) == 0 && fread(&ns, 8, 1, f) == 1 && fread(&nc, 8, 1, f) == 1 &&
                 ns == src.size() && nc > 0 && nc < (1ull << 31);
      if (got) {
        std::string stored(ns, '\0');
        got = fread(&stored[0], 1, ns, f) == ns && stored == src;
      }
      if (got) {
        cubin.resize(nc);
        got = fread(cubin.data(), 1, nc, f) == nc && fgetc(f) == EOF;
      }
      fclose(f);
      if (got) { if (from_cache) *from_cache = true; return true; }
    }
  }
  if (!jit_compile(src, cubin, why)) return false;
  if (!path.empty()) {   // best effort: write to a temporary name, then rename (concurrent ranks compile the same source)
    const std::string dir = jit_cache_dir();
    mkdir(dir.substr(0, dir.find_last_of('/')).c_str(), 0755);
    mkdir(dir.c_str(), 0755);
    const std::string tmp = path + "." + std::to_string((long long)getpid()) + ".tmp";
    if (FILE* f = fopen(tmp.c_str(), "wb")) {
      const unsigned long long ns = src.size(), nc = cubin.size();
      const bool okw = fwrite(kMagic, 1, 8, f) == 8 && fwrite(&ns, 8, 1, f) == 1 && fwrite(&nc, 8, 1, f) == 1 &&
                       fwrite(src.data(), 1, src.size(), f) == src.size() && fwrite(cubin.data(), 1, cubin.size(), f) == cubin.size();
      fclose(f);
      if (!okw || rename(tmp.c_str(), path.c_str()) != 0) remove(tmp.c_str());
    }
  }
  return true;
}

}  // namespace host
}  // namespace spicey
