// NVRTC, loaded lazily with dlopen so that the library itself has no link-time dependency on it: compiles the
// kernels the generators (sparse_codegen.h, tran_codegen.h) write for one netlist to an sm_100a cubin.
#pragma once
#include <dlfcn.h>

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


}  // namespace host
}  // namespace spicey
