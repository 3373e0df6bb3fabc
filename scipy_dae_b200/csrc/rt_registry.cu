// rt_registry.cu — residuals registered at RUN time (b200dae_register_problem, include/b200dae.h).
//
// The reference takes any Python callable fun(t, y, yp, *args) (base.py:44-45, dae.py:365-383).
// On the device the residual is a functor the kernels are instantiated with; the built-in ones are
// compiled by nvcc (problem_*.cu).  A functor registered at run time arrives as CUDA source text
// and goes through NVRTC: the kernel templates are header-only (launch_shapes.cuh and below), so
// the run-time translation unit is
//
//     #include "launch_shapes.cuh"
//     namespace b200dae { namespace rt_user_K { <the caller's source: struct Problem {...}> } }
//
// plus ONE name expression per kernel variant, compiled for sm_100a on first use (a few seconds
// per variant), cached as a cubin on disk, loaded with the driver API and launched with
// cuLaunchKernel.  The variant is chosen by the same rules as launchers.cuh, so a registered copy
// of a built-in functor runs the same template instantiation as the built-in.
//
// libnvrtc and the driver entry points are resolved at run time (dlopen /
// cudaGetDriverEntryPoint): the library loads and exports every symbol on a machine without them,
// and the compile step itself needs no GPU (the -m "not gpu" tests exercise it).
#include <cuda.h>
#include <cuda_runtime.h>
#include <dlfcn.h>
#include <nvrtc.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <sys/stat.h>

#include <atomic>
#include <map>
#include <mutex>
#include <string>
#include <vector>

#include "../../include/b200dae.h"
#include "rt_registry.h"

namespace b200dae {

extern std::atomic<int64_t> g_launches;

namespace {

struct Nvrtc {
    void* h = nullptr;
    decltype(&nvrtcCreateProgram) create = nullptr;
    decltype(&nvrtcDestroyProgram) destroy = nullptr;
    decltype(&nvrtcCompileProgram) compile = nullptr;
    decltype(&nvrtcGetProgramLogSize) log_size = nullptr;
    decltype(&nvrtcGetProgramLog) log = nullptr;
    decltype(&nvrtcAddNameExpression) add_name = nullptr;
    decltype(&nvrtcGetLoweredName) lowered = nullptr;
    decltype(&nvrtcGetCUBINSize) cubin_size = nullptr;
    decltype(&nvrtcGetCUBIN) cubin = nullptr;
    decltype(&nvrtcVersion) version = nullptr;
    bool ok = false;
};

Nvrtc& nvrtc() {
    static Nvrtc n;
    static bool tried = false;
    if (tried) return n;
    tried = true;
    const char* names[] = {"libnvrtc.so.12", "libnvrtc.so", "/usr/local/cuda/lib64/libnvrtc.so.12",
                           "/usr/local/cuda/lib64/libnvrtc.so"};
    for (const char* nm : names) {
        n.h = dlopen(nm, RTLD_NOW | RTLD_LOCAL);
        if (n.h) break;
    }
    if (!n.h) return n;
#define B200DAE_SYM(field, sym) n.field = (decltype(n.field))dlsym(n.h, sym)
    B200DAE_SYM(create, "nvrtcCreateProgram");
    B200DAE_SYM(destroy, "nvrtcDestroyProgram");
    B200DAE_SYM(compile, "nvrtcCompileProgram");
    B200DAE_SYM(log_size, "nvrtcGetProgramLogSize");
    B200DAE_SYM(log, "nvrtcGetProgramLog");
    B200DAE_SYM(add_name, "nvrtcAddNameExpression");
    B200DAE_SYM(lowered, "nvrtcGetLoweredName");
    B200DAE_SYM(cubin_size, "nvrtcGetCUBINSize");
    B200DAE_SYM(cubin, "nvrtcGetCUBIN");
    B200DAE_SYM(version, "nvrtcVersion");
#undef B200DAE_SYM
    n.ok = n.create && n.destroy && n.compile && n.log_size && n.log && n.add_name && n.lowered &&
           n.cubin_size && n.cubin && n.version;
    return n;
}

struct Driver {
    decltype(&cuModuleLoadData) module_load = nullptr;
    decltype(&cuModuleGetFunction) get_function = nullptr;
    decltype(&cuFuncSetAttribute) set_attr = nullptr;
    decltype(&cuLaunchKernel) launch = nullptr;
    decltype(&cuGetErrorString) err_string = nullptr;
    bool ok = false;
};

Driver& driver() {
    static Driver d;
    static bool tried = false;
    if (tried) return d;
    tried = true;
    cudaFree(nullptr);                                   // the runtime's primary context is current
    auto get = [](const char* sym, void** fn) {
        cudaDriverEntryPointQueryResult st;
        return cudaGetDriverEntryPoint(sym, fn, cudaEnableDefault, &st) == cudaSuccess &&
               st == cudaDriverEntryPointSuccess && *fn != nullptr;
    };
    d.ok = get("cuModuleLoadData", (void**)&d.module_load) && get("cuModuleGetFunction", (void**)&d.get_function) &&
           get("cuFuncSetAttribute", (void**)&d.set_attr) && get("cuLaunchKernel", (void**)&d.launch) &&
           get("cuGetErrorString", (void**)&d.err_string);
    return d;
}

std::string self_dir() {
    Dl_info info;
    if (dladdr((void*)&self_dir, &info) && info.dli_fname) {
        std::string p = info.dli_fname;
        const size_t k = p.rfind('/');
        return k == std::string::npos ? "." : p.substr(0, k);
    }
    return ".";
}

bool read_file(const std::string& path, std::string& out) {
    FILE* f = fopen(path.c_str(), "rb");
    if (!f) return false;
    char buf[1 << 16];
    size_t n;
    out.clear();
    while ((n = fread(buf, 1, sizeof(buf), f)) > 0) out.append(buf, n);
    fclose(f);
    return true;
}

uint64_t fnv1a(const std::string& s, uint64_t h = 1469598103934665603ull) {
    for (unsigned char c : s) { h ^= c; h *= 1099511628211ull; }
    return h;
}

// the kernel headers, wherever this build of the library lives: <lib dir>/csrc or, for the A/B
// variant libraries under <pkg>/variants, <lib dir>/../csrc; B200DAE_RT_INCLUDE overrides
const std::string& csrc_dir() {
    static std::string dir;
    if (!dir.empty()) return dir;
    const char* env = getenv("B200DAE_RT_INCLUDE");
    std::vector<std::string> cands;
    if (env && *env) cands.push_back(env);
    const std::string here = self_dir();
    cands.push_back(here + "/csrc");
    cands.push_back(here + "/../csrc");
    for (const std::string& c : cands) {
        struct stat st;
        if (stat((c + "/launch_shapes.cuh").c_str(), &st) == 0) { dir = c; break; }
    }
    return dir;
}

// every header the run-time translation unit can see takes part in the cache key
const std::string& headers_digest() {
    static std::string digest;
    if (!digest.empty()) return digest;
    static const char* files[] = {"launch_shapes.cuh", "radau.cuh", "radau_group.cuh", "bdf.cuh", "bdf_group.cuh",
                                  "cic.cuh", "group.cuh", "dense_lu.cuh", "fd_jacobian.cuh", "fastmath.cuh",
                                  "events.cuh", "problems.cuh", "rt_compat.cuh", "../../include/b200dae.h"};
    uint64_t h = 1469598103934665603ull;
    std::string text;
    for (const char* f : files) {
        if (read_file(csrc_dir() + "/" + f, text)) h = fnv1a(text, h);
    }
    char buf[32];
    snprintf(buf, sizeof(buf), "%016llx", (unsigned long long)h);
    digest = buf;
    return digest;
}

std::string cache_dir() {
    const char* env = getenv("B200DAE_RT_CACHE");
    std::string d;
    if (env && *env) d = env;
    else {
        const char* home = getenv("HOME");
        d = std::string(home && *home ? home : "/tmp") + "/.cache/scipy_dae_b200/rt";
    }
    std::string acc;
    for (size_t i = 0; i <= d.size(); ++i) {             // mkdir -p
        if (i == d.size() || d[i] == '/') {
            if (!acc.empty()) mkdir(acc.c_str(), 0755);
        }
        if (i < d.size()) acc.push_back(d[i]);
    }
    return d;
}

// the integers of a lowered name "..._kernelILi3ELi64ELin1ELb1E...EEvv" in order
std::vector<long long> mangled_ints(const std::string& lowered, const char* after) {
    std::vector<long long> v;
    size_t p = lowered.find(after);
    if (p == std::string::npos) return v;
    p += strlen(after);
    while (p + 1 < lowered.size() && lowered[p] == 'L') {
        size_t q = p + 2;                                 // skip 'L' and the type letter
        bool neg = false;
        if (q < lowered.size() && lowered[q] == 'n') { neg = true; ++q; }
        long long x = 0;
        while (q < lowered.size() && lowered[q] >= '0' && lowered[q] <= '9') x = 10 * x + (lowered[q++] - '0');
        if (q >= lowered.size() || lowered[q] != 'E') break;
        v.push_back(neg ? -x : x);
        p = q + 1;
    }
    return v;
}

struct RtKernel {
    std::string lowered;
    std::vector<char> cubin;
    int G = 1, T = 0, MB = 0;
    long long smem = 0;
    int carveout = 0;                                     // preferred shared-memory carve-out in percent (0: driver's choice)
    CUfunction fn = nullptr;                              // null until loaded on a device
};

struct RtProblem {
    std::string name, src;
    int ns = 0;                                           // namespace number of the wrapped source
    int n = 0, n_params = 0, n_events = 0;
    bool has_jac = false;
    std::map<std::string, RtKernel> kernels;
};

std::mutex g_mu;
std::vector<RtProblem*> g_problems;

std::string translation_unit(const RtProblem& p, const std::string& shape_expr) {
    char head[256];
    snprintf(head, sizeof(head), "namespace b200dae { namespace rt_user_%d {\n#line 1 \"%s.cu\"\n", p.ns, p.name.c_str());
    std::string tu = "#include \"launch_shapes.cuh\"\n";
    tu += head;
    tu += p.src;
    tu += "\n} }\n";
    char tail[512];
    snprintf(tail, sizeof(tail),
             "typedef b200dae::rt_user_%d::Problem P;\n"
             "static_assert(P::N == %d && P::NP == %d, \"struct Problem: N / NP differ from the registered n / n_params\");\n"
             "namespace b200dae { template <int A, int B, int C, long long D> __global__ void rt_shape_kernel() {} }\n"
             "using namespace b200dae;\n",
             p.ns, p.n, p.n_params);
    tu += tail;
    (void)shape_expr;
    return tu;
}

// Compiles (or fetches from the disk cache) one kernel variant.  kernel_expr / shape_expr are C++
// constant expressions over P (see rt_launch_*); shape = (G, T, MB, dynamic shared memory bytes).
int get_kernel(RtProblem& p, const std::string& key, const std::string& kernel_expr, const std::string& shape_expr,
               RtKernel** out, std::string& err) {
    auto it = p.kernels.find(key);
    if (it != p.kernels.end()) { *out = &it->second; return 0; }
    Nvrtc& rt = nvrtc();
    if (!rt.ok) { err = "libnvrtc.so.12 not found: run-time registered problems need NVRTC"; return B200DAE_ERTC; }
    if (csrc_dir().empty()) { err = "kernel headers (csrc/launch_shapes.cuh) not found next to the library; set B200DAE_RT_INCLUDE"; return B200DAE_ERTC; }
    const std::string tu = translation_unit(p, shape_expr);
    const std::string shape_name = "b200dae::rt_shape_kernel<" + shape_expr + ">";
    int major = 0, minor = 0;
    rt.version(&major, &minor);
    char vbuf[64];
    snprintf(vbuf, sizeof(vbuf), "nvrtc %d.%d sm_100a|", major, minor);
    const uint64_t h = fnv1a(kernel_expr + "|" + shape_expr + "|" + tu, fnv1a(std::string(vbuf) + headers_digest()));
    char hbuf[32];
    snprintf(hbuf, sizeof(hbuf), "%016llx", (unsigned long long)h);
    const std::string base = cache_dir() + "/" + p.name + "_" + hbuf;
    RtKernel k;
    std::string meta, bin;
    bool cached = false;
    if (!getenv("B200DAE_RT_NOCACHE") && read_file(base + ".meta", meta) && read_file(base + ".cubin", bin) && !bin.empty()) {
        char low[1024];
        long long smem;
        if (sscanf(meta.c_str(), "%1023s %d %d %d %lld", low, &k.G, &k.T, &k.MB, &smem) == 5) {
            k.lowered = low;
            k.smem = smem;
            k.cubin.assign(bin.begin(), bin.end());
            cached = true;
        }
    }
    if (!cached) {
        nvrtcProgram prog;
        if (rt.create(&prog, tu.c_str(), (p.name + "_rt.cu").c_str(), 0, nullptr, nullptr) != NVRTC_SUCCESS) {
            err = "nvrtcCreateProgram failed";
            return B200DAE_ERTC;
        }
        rt.add_name(prog, kernel_expr.c_str());
        rt.add_name(prog, shape_name.c_str());
        const std::string inc1 = "-I" + csrc_dir(), inc2 = "-I" + csrc_dir() + "/../../include";
        const char* opts[] = {"--gpu-architecture=sm_100a", "-std=c++17", "-lineinfo", "-default-device",
                              inc1.c_str(), inc2.c_str()};
        const nvrtcResult rc = rt.compile(prog, (int)(sizeof(opts) / sizeof(opts[0])), opts);
        if (rc != NVRTC_SUCCESS) {
            size_t n = 0;
            rt.log_size(prog, &n);
            std::string log(n, '\0');
            if (n) rt.log(prog, &log[0]);
            rt.destroy(&prog);
            // the first errors are the useful ones; the thread-local message buffer is small
            size_t e = log.find("error");
            err = "NVRTC: " + (e == std::string::npos ? log : log.substr(log.rfind('\n', e) == std::string::npos ? 0 : log.rfind('\n', e) + 1));
            return B200DAE_ERTC;
        }
        const char* low = nullptr;
        const char* low_shape = nullptr;
        if (rt.lowered(prog, kernel_expr.c_str(), &low) != NVRTC_SUCCESS || rt.lowered(prog, shape_name.c_str(), &low_shape) != NVRTC_SUCCESS) {
            rt.destroy(&prog);
            err = "nvrtcGetLoweredName failed for " + kernel_expr;
            return B200DAE_ERTC;
        }
        k.lowered = low;
        const std::vector<long long> sh = mangled_ints(low_shape, "rt_shape_kernelI");
        if (sh.size() != 4) { rt.destroy(&prog); err = std::string("cannot read the launch shape from ") + low_shape; return B200DAE_ERTC; }
        k.G = (int)sh[0]; k.T = (int)sh[1]; k.MB = (int)sh[2]; k.smem = sh[3];
        size_t sz = 0;
        rt.cubin_size(prog, &sz);
        k.cubin.resize(sz);
        rt.cubin(prog, k.cubin.data());
        rt.destroy(&prog);
        if (!getenv("B200DAE_RT_NOCACHE")) {
            FILE* f = fopen((base + ".cubin.tmp").c_str(), "wb");
            if (f) {
                fwrite(k.cubin.data(), 1, k.cubin.size(), f);
                fclose(f);
                rename((base + ".cubin.tmp").c_str(), (base + ".cubin").c_str());
                f = fopen((base + ".meta").c_str(), "w");
                if (f) { fprintf(f, "%s %d %d %d %lld\n", k.lowered.c_str(), k.G, k.T, k.MB, k.smem); fclose(f); }
            }
        }
    }
    auto ins = p.kernels.emplace(key, std::move(k));
    *out = &ins.first->second;
    return 0;
}

int load_kernel(RtKernel& k, std::string& err) {
    if (k.fn) return 0;
    Driver& d = driver();
    if (!d.ok) { err = "CUDA driver entry points not available (no device?)"; return B200DAE_ECUDA; }
    auto msg = [&](CUresult rc, const char* what) {
        const char* s = nullptr;
        d.err_string(rc, &s);
        err = std::string(what) + ": " + (s ? s : "unknown driver error");
        return B200DAE_ECUDA;
    };
    CUmodule mod;
    CUresult rc = d.module_load(&mod, k.cubin.data());
    if (rc != CUDA_SUCCESS) return msg(rc, "cuModuleLoadData");
    rc = d.get_function(&k.fn, mod, k.lowered.c_str());
    if (rc != CUDA_SUCCESS) { k.fn = nullptr; return msg(rc, "cuModuleGetFunction"); }
    if (k.smem > 0) {
        rc = d.set_attr(k.fn, CU_FUNC_ATTRIBUTE_MAX_DYNAMIC_SHARED_SIZE_BYTES, (int)k.smem);
        if (rc != CUDA_SUCCESS) { k.fn = nullptr; return msg(rc, "cuFuncSetAttribute(max dynamic shared memory)"); }
    }
    if (k.carveout > 0) (void)d.set_attr(k.fn, CU_FUNC_ATTRIBUTE_PREFERRED_SHARED_MEMORY_CARVEOUT, k.carveout);   // a hint
    return 0;
}

int launch(RtKernel& k, int64_t units, void* arg, cudaStream_t st, std::string& err) {
    int rc = load_kernel(k, err);
    if (rc) return rc;
    const int64_t blocks = (units * k.G + k.T - 1) / k.T;
    void* params[] = {arg};
    const CUresult cr = driver().launch(k.fn, (unsigned)blocks, 1, 1, (unsigned)k.T, 1, 1, (unsigned)k.smem,
                                        (CUstream)st, params, nullptr);
    if (cr != CUDA_SUCCESS) {
        const char* s = nullptr;
        driver().err_string(cr, &s);
        err = std::string("cuLaunchKernel: ") + (s ? s : "unknown driver error");
        return B200DAE_ECUDA;
    }
    g_launches.fetch_add(1);
    return 0;
}

const char* tf(bool b) { return b ? "true" : "false"; }

}  // namespace

// ---- registration ---------------------------------------------------------------------------
int rt_register(const char* name, const char* src, int n, int n_params, std::string& err) {
    std::lock_guard<std::mutex> lock(g_mu);
    for (size_t i = 0; i < g_problems.size(); ++i)
        if (g_problems[i]->name == name) {
            if (g_problems[i]->src == src && g_problems[i]->n == n && g_problems[i]->n_params == n_params) return (int)i;
            err = std::string("problem '") + name + "' is already registered with a different source";
            return B200DAE_EINVAL;
        }
    RtProblem* p = new RtProblem;
    p->name = name;
    p->src = src;
    p->ns = (int)g_problems.size();
    p->n = n;
    p->n_params = n_params;
    // the residual-check kernel doubles as the syntax / interface check of the source, and its
    // "shape" carries what the registry needs to know about the functor
    RtKernel* k = nullptr;
    const int rc = get_kernel(*p, "residual", "b200dae::residual_kernel<P>",
                              "(int)P::kHasJac, 128, b200dae::EventOf<P>::kNum, 0", &k, err);
    if (rc) { delete p; return rc; }
    p->has_jac = k->G != 0;
    p->n_events = k->MB;
    k->G = 1;
    k->MB = 1;
    g_problems.push_back(p);
    return (int)g_problems.size() - 1;
}

int rt_count() {
    std::lock_guard<std::mutex> lock(g_mu);
    return (int)g_problems.size();
}

bool rt_info(int idx, const char** name, int* n, int* n_params, int* n_events, int* has_jac) {
    std::lock_guard<std::mutex> lock(g_mu);
    if (idx < 0 || idx >= (int)g_problems.size()) return false;
    const RtProblem* p = g_problems[idx];
    if (name) *name = p->name.c_str();
    if (n) *n = p->n;
    if (n_params) *n_params = p->n_params;
    if (n_events) *n_events = p->n_events;
    if (has_jac) *has_jac = p->has_jac ? 1 : 0;
    return true;
}

// ---- variant selection: the rules of launchers.cuh, as strings ---------------------------------
int rt_launch_radau(int idx, const RadauLaunch& L, cudaStream_t st, bool do_launch, std::string& err) {
    std::lock_guard<std::mutex> lock(g_mu);
    RtProblem& p = *g_problems[idx];
    const int S = L.tb.s;
    const bool jac_ok = p.has_jac && S == 3;
    if (L.analytic_jac && !jac_ok) return kRtErrNoJac;
    char key[128], kern[512], shape[256];
    if (p.n > B200DAE_REG_MAXN) {                          // group regime (launch_radau_group)
        const bool glean = L.newton_iter_embedded == 1 && L.cew == 0.0 && L.out.t_eval == nullptr &&
                           L.out.t_steps == nullptr && L.ev.spec.n_events == 0;
        const bool J = L.analytic_jac != 0, lean = !J && glean;
        snprintf(key, sizeof(key), "radau_group s%d j%d l%d", S, J, lean);
        snprintf(kern, sizeof(kern), "b200dae::radau_group_kernel<P, %d, RadauGroupShape<P, %d>::G, RadauGroupShape<P, %d>::T, "
                 "RadauGroupShape<P, %d>::MB, %s, %s>", S, S, S, S, tf(J), tf(lean));
        snprintf(shape, sizeof(shape), "RadauGroupShape<P, %d>::G, RadauGroupShape<P, %d>::T, RadauGroupShape<P, %d>::MB, "
                 "(long long)RadauGroupShape<P, %d>::smem", S, S, S, S);
    } else {                                               // thread regime (launch_radau)
        const bool lean_opts = L.newton_iter_embedded == 1 && L.cew == 0.0 && L.out.t_steps == nullptr &&
                               L.ev.spec.n_events == 0;
        const bool lean = lean_opts && L.out.t_eval == nullptr;
        bool kl = lean, kj = false, ke = false, kio = false;
        int bits = 5;
        if (L.analytic_jac) kj = true;
        else if (S == 3 && lean_opts && !lean) { kl = true; ke = true; }
        else if (S == 3 && lean && L.newton_max_iter <= 8) { bits = 3; kio = L.io_aos != 0; }
        snprintf(key, sizeof(key), "radau s%d l%d j%d e%d b%d io%d", S, kl, kj, ke, bits, kio);
        snprintf(kern, sizeof(kern), "b200dae::radau_kernel<P, %d, RadauTune<P, %d>::kThreads, RadauTune<P, %d>::kMinBlocks, "
                 "RadauTune<P, %d>::kSharedCold, %s, %s, %s, %d, %s>", S, S, S, S, tf(kl), tf(kj), tf(ke), bits, tf(kio));
        snprintf(shape, sizeof(shape), "1, RadauTune<P, %d>::kThreads, RadauTune<P, %d>::kMinBlocks, (long long)RadauTune<P, %d>::kDynSmem", S, S, S);
    }
    RtKernel* k = nullptr;
    int rc = get_kernel(p, key, kern, shape, &k, err);
    if (rc) return rc;
    // the unrolled thread regime beyond n = 4 wants ONE resident block per SM (launchers.cuh: B200DAE_BIG_CARVEOUT_DEFAULT)
    if (p.n > 4 && p.n <= B200DAE_REG_MAXN && strstr(key, "b3") != nullptr && k->fn == nullptr) k->carveout = 40;
    if (!do_launch) return 0;
    RadauLaunch Lc = L;
    return launch(*k, L.N, &Lc, st, err);
}

int rt_launch_bdf(int idx, const BdfLaunch& L, cudaStream_t st, bool do_launch, std::string& err) {
    std::lock_guard<std::mutex> lock(g_mu);
    RtProblem& p = *g_problems[idx];
    if (L.analytic_jac && !p.has_jac) return kRtErrNoJac;
    char key[128], kern[512], shape[256];
    const bool ev = L.ev.spec.n_events > 0;
    const bool plain = !ev && L.out.t_eval == nullptr && L.out.t_steps == nullptr;
    if (p.n > B200DAE_REG_MAXN) {
        const bool J = L.analytic_jac != 0;
        const bool E = ev, O = J ? true : !plain;
        snprintf(key, sizeof(key), "bdf_group j%d e%d o%d", J, E, O);
        snprintf(kern, sizeof(kern), "b200dae::bdf_group_kernel<P, BdfGroupShape<P>::G, BdfGroupShape<P>::T, BdfGroupShape<P>::MB, %s, %s, %s>",
                 tf(J), tf(E), tf(O));
        snprintf(shape, sizeof(shape), "BdfGroupShape<P>::G, BdfGroupShape<P>::T, BdfGroupShape<P>::MB, (long long)BdfGroupShape<P>::smem");
    } else {
        const bool J = L.analytic_jac != 0, lean = !J && plain;
        snprintf(key, sizeof(key), "bdf j%d l%d", J, lean);
        snprintf(kern, sizeof(kern), "b200dae::bdf_kernel<P, BdfTune<P>::kThreads, %s, %s, BdfTune<P>::kSharedCold>", tf(J), tf(lean));
        snprintf(shape, sizeof(shape), "1, BdfTune<P>::kThreads, 1, (long long)BdfTune<P>::kDynSmem");
    }
    RtKernel* k = nullptr;
    int rc = get_kernel(p, key, kern, shape, &k, err);
    if (rc) return rc;
    if (!do_launch) return 0;
    BdfLaunch Lc = L;
    return launch(*k, L.N, &Lc, st, err);
}

int rt_launch_cic(int idx, const CicLaunch& L, cudaStream_t st, bool do_launch, std::string& err) {
    std::lock_guard<std::mutex> lock(g_mu);
    RtProblem& p = *g_problems[idx];
    if (L.analytic_jac && !p.has_jac) return kRtErrNoJac;
    const bool J = L.analytic_jac != 0;
    char key[64], kern[128];
    snprintf(key, sizeof(key), "cic j%d", J);
    snprintf(kern, sizeof(kern), "b200dae::cic_kernel<P, %s>", tf(J));
    RtKernel* k = nullptr;
    int rc = get_kernel(p, key, kern, "1, 32, 1, 0", &k, err);
    if (rc) return rc;
    if (!do_launch) return 0;
    CicLaunch Lc = L;
    return launch(*k, L.N, &Lc, st, err);
}

int rt_launch_residual(int idx, int64_t N, int64_t ld, double t, const double* y, const double* yp,
                       const double* params, double* out, cudaStream_t st, std::string& err) {
    std::lock_guard<std::mutex> lock(g_mu);
    RtProblem& p = *g_problems[idx];
    RtKernel& k = p.kernels.at("residual");
    int rc = load_kernel(k, err);
    if (rc) return rc;
    void* params_v[] = {&N, &ld, &t, &y, &yp, &params, &out};
    const CUresult cr = driver().launch(k.fn, (unsigned)((N + 127) / 128), 1, 1, 128, 1, 1, 0, (CUstream)st, params_v, nullptr);
    if (cr != CUDA_SUCCESS) { err = "cuLaunchKernel(residual_kernel) failed"; return B200DAE_ECUDA; }
    g_launches.fetch_add(1);
    return 0;
}

}  // namespace b200dae
