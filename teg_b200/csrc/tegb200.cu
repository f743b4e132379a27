// tegb200.cu -- libtegb200.so: the C-ABI runtime of the Teg CUDA (sm_100a) evaluation backend.
//
// See include/tegb200.h for the contract and for which reference code each entry point replaces.
// Contents:
//   * NVRTC JIT of generated translation units (teg_b200/codegen.py) to sm_100a cubins;
//   * module loading / kernel launch through the CUDA driver API (resolved lazily with dlopen so the
//     library loads -- and compiles programs -- on machines without a GPU);
//   * static kernels: deterministic column sums (gradient reduction), FMA-peak micro-benchmark;
//   * NCCL all-reduce of the parameter-gradient vector (libnccl resolved lazily).
//
// Build (see __graft_entry__.build):
//   nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 --fmad=false -shared
//        -Xcompiler -fPIC tegb200.cu -o libtegb200.so -lnvrtc -ldl

#include "tegb200.h"

#include <cuda.h>
#include <cuda_runtime.h>
#include <dlfcn.h>
#include <nvrtc.h>

#include <atomic>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <string>
#include <vector>

#include "tegb200_runtime_header.inc"   // const char TEGB_RUNTIME_HEADER[] = contents of teg_b200_runtime.cuh

// ---------------------------------------------------------------------------------------------- errors
static thread_local char g_err[8192];
static std::atomic<long long> g_launches{0};

static int fail(int code, const char *fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof g_err, fmt, ap);
    va_end(ap);
    return code;
}

extern "C" const char *tegb_last_error(void) { return g_err; }
extern "C" int tegb_version(void) { return 100; }
extern "C" int64_t tegb_launch_count(void) { return g_launches.load(); }
// FNV-1a of the embedded runtime header: part of the cubin cache key (generated units #include it by name only)
extern "C" uint64_t tegb_runtime_header_hash(void) {
    uint64_t h = 1469598103934665603ull;
    for (const char *c = TEGB_RUNTIME_HEADER; *c; ++c) { h ^= (unsigned char)*c; h *= 1099511628211ull; }
    return h;
}

// ------------------------------------------------------------------------------------- driver API shim
namespace drv {
typedef CUresult (*cuInit_t)(unsigned int);
typedef CUresult (*cuGetErrorString_t)(CUresult, const char **);
typedef CUresult (*cuModuleLoadData_t)(CUmodule *, const void *);
typedef CUresult (*cuModuleUnload_t)(CUmodule);
typedef CUresult (*cuModuleGetFunction_t)(CUfunction *, CUmodule, const char *);
typedef CUresult (*cuModuleGetGlobal_t)(CUdeviceptr *, size_t *, CUmodule, const char *);
typedef CUresult (*cuLaunchKernel_t)(CUfunction, unsigned, unsigned, unsigned, unsigned, unsigned, unsigned,
                                     unsigned, CUstream, void **, void **);
typedef CUresult (*cuMemcpyDtoDAsync_t)(CUdeviceptr, CUdeviceptr, size_t, CUstream);

static void *lib = nullptr;
static cuInit_t Init;
static cuGetErrorString_t GetErrorString;
static cuModuleLoadData_t ModuleLoadData;
static cuModuleUnload_t ModuleUnload;
static cuModuleGetFunction_t ModuleGetFunction;
static cuModuleGetGlobal_t ModuleGetGlobal;
static cuLaunchKernel_t LaunchKernel;
static cuMemcpyDtoDAsync_t MemcpyDtoDAsync;
static std::once_flag once;
static bool ok = false;
static std::string why;

static void load() {
    lib = dlopen("libcuda.so.1", RTLD_NOW | RTLD_GLOBAL);
    if (!lib) { why = std::string("cannot load libcuda.so.1: ") + dlerror(); return; }
#define SYM(var, name)                                                   \
    var = (decltype(var))dlsym(lib, name);                               \
    if (!var) { why = std::string("libcuda.so.1 lacks ") + name; return; }
    SYM(Init, "cuInit");
    SYM(GetErrorString, "cuGetErrorString");
    SYM(ModuleLoadData, "cuModuleLoadData");
    SYM(ModuleUnload, "cuModuleUnload");
    SYM(ModuleGetFunction, "cuModuleGetFunction");
    SYM(ModuleGetGlobal, "cuModuleGetGlobal_v2");
    SYM(LaunchKernel, "cuLaunchKernel");
    SYM(MemcpyDtoDAsync, "cuMemcpyDtoDAsync_v2");
#undef SYM
    CUresult r = Init(0);
    if (r != CUDA_SUCCESS) { why = "cuInit failed (no usable CUDA driver/device)"; return; }
    ok = true;
}

static bool ready() {
    std::call_once(once, load);
    return ok;
}

static const char *errstr(CUresult r) {
    const char *s = nullptr;
    if (GetErrorString && GetErrorString(r, &s) == CUDA_SUCCESS && s) return s;
    return "unknown CUDA driver error";
}
}  // namespace drv

#define CU_CHECK(call)                                                                      \
    do {                                                                                    \
        CUresult _r = (call);                                                               \
        if (_r != CUDA_SUCCESS) return fail(TEGB_ERR_CUDA, "%s failed: %s", #call, drv::errstr(_r)); \
    } while (0)

#define RT_CHECK(call)                                                                           \
    do {                                                                                         \
        cudaError_t _e = (call);                                                                 \
        if (_e != cudaSuccess) return fail(TEGB_ERR_CUDA, "%s failed: %s", #call, cudaGetErrorString(_e)); \
    } while (0)

extern "C" int tegb_device_count(void) {
    if (!drv::ready()) return 0;
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
    return n;
}

// --------------------------------------------------------------------------------------- NVRTC shim
// NVRTC is loaded by PATH, not by soname: a process that imported torch already has torch's bundled
// libnvrtc.so.12 (an older 12.x) mapped, and a soname lookup would silently reuse it.  The toolkit's own
// NVRTC (>= 12.9: PTX 8.8, 3-input min/max for sm_100) is preferred; TEGB200_NVRTC overrides the path.
namespace rtc {
typedef nvrtcResult (*Version_t)(int *, int *);
typedef nvrtcResult (*CreateProgram_t)(nvrtcProgram *, const char *, const char *, int, const char *const *, const char *const *);
typedef nvrtcResult (*DestroyProgram_t)(nvrtcProgram *);
typedef nvrtcResult (*CompileProgram_t)(nvrtcProgram, int, const char *const *);
typedef nvrtcResult (*GetSize_t)(nvrtcProgram, size_t *);
typedef nvrtcResult (*GetData_t)(nvrtcProgram, char *);
typedef const char *(*GetErrorString_t)(nvrtcResult);
static void *lib = nullptr;
static Version_t Version;
static CreateProgram_t CreateProgram;
static DestroyProgram_t DestroyProgram;
static CompileProgram_t CompileProgram;
static GetSize_t GetProgramLogSize, GetCUBINSize;
static GetData_t GetProgramLog, GetCUBIN;
static GetErrorString_t GetErrorString;
static int major = 0, minor = 0;
static std::once_flag once;
static bool ok = false;
static std::string why;
static void load() {
    std::vector<std::string> cands;
    if (const char *e = getenv("TEGB200_NVRTC")) cands.push_back(e);
    cands.push_back("/usr/local/cuda/lib64/libnvrtc.so.12");
    cands.push_back("/usr/local/cuda/lib64/libnvrtc.so");
    cands.push_back("libnvrtc.so.12");
    cands.push_back("libnvrtc.so");
    for (auto &c : cands) {
        lib = dlopen(c.c_str(), RTLD_NOW | RTLD_LOCAL);
        if (lib) break;
    }
    if (!lib) { why = std::string("cannot load NVRTC: ") + dlerror(); return; }
#define SYM(var, name)                                                   \
    var = (decltype(var))dlsym(lib, name);                               \
    if (!var) { why = std::string("NVRTC lacks ") + name; return; }
    SYM(Version, "nvrtcVersion");
    SYM(CreateProgram, "nvrtcCreateProgram");
    SYM(DestroyProgram, "nvrtcDestroyProgram");
    SYM(CompileProgram, "nvrtcCompileProgram");
    SYM(GetProgramLogSize, "nvrtcGetProgramLogSize");
    SYM(GetProgramLog, "nvrtcGetProgramLog");
    SYM(GetCUBINSize, "nvrtcGetCUBINSize");
    SYM(GetCUBIN, "nvrtcGetCUBIN");
    SYM(GetErrorString, "nvrtcGetErrorString");
#undef SYM
    Version(&major, &minor);
    ok = true;
}
static bool ready() { std::call_once(once, load); return ok; }
}  // namespace rtc

extern "C" int tegb_nvrtc_version(void) { return rtc::ready() ? rtc::major * 1000 + rtc::minor * 10 : 0; }

// ------------------------------------------------------------------------------------------------ JIT
struct tegb_module {
    std::vector<char> cubin;
    std::string log;
};

extern "C" int tegb_compile(const char *cuda_src, const char *arch, const char *opts, tegb_module **out) {
    if (!cuda_src || !out) return fail(TEGB_ERR_INVALID, "tegb_compile: null argument");
    *out = nullptr;
    if (!rtc::ready()) return fail(TEGB_ERR_COMPILE, "%s", rtc::why.c_str());
    const char *header_names[] = {"teg_b200_runtime.cuh"};
    const char *header_srcs[] = {TEGB_RUNTIME_HEADER};
    nvrtcProgram prog;
    nvrtcResult r = rtc::CreateProgram(&prog, cuda_src, "teg_program.cu", 1, header_srcs, header_names);
    if (r != NVRTC_SUCCESS) return fail(TEGB_ERR_COMPILE, "nvrtcCreateProgram: %s", rtc::GetErrorString(r));

    std::vector<std::string> o;
    o.push_back(std::string("--gpu-architecture=") + (arch && *arch ? arch : "sm_100a"));
    o.push_back("--fmad=false");       // one IEEE operation per reference operation (parity contract)
    if (rtc::major > 12 || (rtc::major == 12 && rtc::minor >= 9))
        o.push_back("-DTEGB_HAVE_MIN3=1");     // PTX 8.8: min.NaN.f32 with three inputs (FMNMX3 on sm_100)
    o.push_back("--std=c++17");
    o.push_back("-lineinfo");
    if (opts) {
        std::string s(opts), tok;
        for (size_t i = 0; i <= s.size(); i++) {
            if (i == s.size() || s[i] == ' ') { if (!tok.empty()) o.push_back(tok); tok.clear(); }
            else tok.push_back(s[i]);
        }
    }
    std::vector<const char *> oc;
    for (auto &s : o) oc.push_back(s.c_str());
    r = rtc::CompileProgram(prog, (int)oc.size(), oc.data());
    size_t logsz = 0;
    rtc::GetProgramLogSize(prog, &logsz);
    std::string log(logsz > 0 ? logsz : 1, '\0');
    if (logsz > 1) rtc::GetProgramLog(prog, &log[0]);
    if (r != NVRTC_SUCCESS) {
        int rc = fail(TEGB_ERR_COMPILE, "NVRTC: %s\n%s", rtc::GetErrorString(r), log.c_str());
        rtc::DestroyProgram(&prog);
        return rc;
    }
    size_t sz = 0;
    r = rtc::GetCUBINSize(prog, &sz);
    if (r != NVRTC_SUCCESS || sz == 0) {
        rtc::DestroyProgram(&prog);
        return fail(TEGB_ERR_COMPILE, "nvrtcGetCUBINSize: %s (arch must be a real sm_XX)", rtc::GetErrorString(r));
    }
    tegb_module *m = new tegb_module();
    m->cubin.resize(sz);
    rtc::GetCUBIN(prog, m->cubin.data());
    m->log = log.c_str();
    rtc::DestroyProgram(&prog);
    *out = m;
    return TEGB_OK;
}

extern "C" int tegb_module_from_cubin(const void *cubin, size_t nbytes, tegb_module **out) {
    if (!cubin || !nbytes || !out) return fail(TEGB_ERR_INVALID, "tegb_module_from_cubin: null argument");
    tegb_module *m = new tegb_module();
    m->cubin.assign((const char *)cubin, (const char *)cubin + nbytes);
    *out = m;
    return TEGB_OK;
}

extern "C" int tegb_module_cubin(const tegb_module *m, const void **cubin, size_t *nbytes) {
    if (!m || !cubin || !nbytes) return fail(TEGB_ERR_INVALID, "tegb_module_cubin: null argument");
    *cubin = m->cubin.data();
    *nbytes = m->cubin.size();
    return TEGB_OK;
}

extern "C" const char *tegb_module_log(const tegb_module *m) { return m ? m->log.c_str() : ""; }
extern "C" void tegb_module_destroy(tegb_module *m) { delete m; }

// ------------------------------------------------------------------------------------------- programs
struct tegb_peer;
struct tegb_program {
    tegb_program_desc d;
    int device = 0;
    // A handle owns device scratch (the published uniform table, tile tables, partial sums, staging): its launches are
    // serialised on the host by `mu`, and a launch on another stream than the previous one first waits (on the device)
    // for that previous launch -- so one handle may be shared by threads and streams; independent handles overlap.
    std::recursive_mutex mu;
    cudaEvent_t ev_last = nullptr;    // recorded after the last launch sequence of this handle
    bool ev_valid = false;
    CUstream ev_stream = nullptr;
    // measurement: CUDA events around the MAIN kernel of the last launch (tegb_program_enable_timing)
    bool timing = false;
    cudaEvent_t ev_t0 = nullptr, ev_t1 = nullptr;
    bool timing_valid = false;
    int n_groups_total = 0;           // grouped programs with a peer: rows of the all-reduced [groups][outputs] result
    // grouped programs: what the last per-group prologue (uniform rows + tile classification) covered
    bool prep_valid = false;
    int prep_begin = 0, prep_count = 0, prep_max_rows = 0, prep_max_cols = 0;
    tegb_grid prep_grid{};
    const int32_t *prep_windows = nullptr;
    const int *group_map = nullptr;   // device [n_groups_total]: this rank's index of every global group, or -1 (NULL: identity)
    CUmodule mod = nullptr;
    CUfunction f_uniform = nullptr, f_main = nullptr, f_tiles = nullptr, f_gather = nullptr;
    CUdeviceptr u_const = 0;      // __constant__ table inside the module
    size_t u_const_bytes = 0;
    void *u_stage = nullptr;      // device buffer the uniform kernel writes
    // staging for tegb_program_eval_host
    std::vector<void *> d_in, d_out;
    int64_t cap = 0;
    // grouped programs: one uniform-table row per group
    void *d_utab = nullptr;
    size_t utab_bytes = 0;
    // scalars of the last uniform prologue: identical scalars => the published table is still valid
    std::vector<double> last_scalars;
    bool uniform_valid = false;
    CUstream last_stream = nullptr;
    // per-block partial sums of reduce_outputs programs
    void *d_partials = nullptr;
    unsigned int *d_red_done = nullptr;     // per-column counters of the one-launch column sum (zero between launches)
    size_t partials_bytes = 0;
    // set by tegb_program_set_peer: the sums of reduce_outputs launches are all-reduced over the peer mailboxes
    struct tegb_peer *peer = nullptr;
    // tri-states of the tile conditions (render programs with n_tile_conds > 0)
    void *d_tiles = nullptr;
    size_t tiles_bytes = 0;
    void *tiles_ord_ptr = nullptr;    // the launch-ordered copy inside d_tiles (plain tiled renders)
    int *q_ulist = nullptr, *q_counters = nullptr;     // unresolved-row queue of the last plain tiled launch
    bool counters_dirty = true;                        // the counters are not known to be zero (first launch / after a failure)
    int q_n_tx = 0, q_n_ty = 0, q_workers = 0;
    // per tile row / tile column: [min, max] of the box edges and sample abscissae (valid for ranges_grid)
    int *d_order = nullptr;       // [n_tiles] tile indices in launch order, then 2 counters
    size_t order_count = 0;
    void *d_ranges = nullptr;
    size_t ranges_bytes = 0;
    tegb_grid ranges_grid{};
    int ranges_stride = 1, ranges_phase = 0;
    bool ranges_valid = false;
    // pixel-edge tables for tegb_render_launch
    void *d_xe = nullptr, *d_ye = nullptr;
    std::vector<char> h_xe, h_ye;
    tegb_grid grid_cached{};
    bool grid_valid = false;
};

extern "C" int tegb_program_create(tegb_module *m, const tegb_program_desc *desc, int device, tegb_program **out) {
    if (!m || !desc || !out) return fail(TEGB_ERR_INVALID, "tegb_program_create: null argument");
    *out = nullptr;
    if (desc->elem_size != 4 && desc->elem_size != 8) return fail(TEGB_ERR_INVALID, "elem_size must be 4 or 8");
    if (desc->n_outputs < 1 || desc->block_size < 32 || desc->block_size > 1024)
        return fail(TEGB_ERR_INVALID, "bad program description");
    if (desc->reduce_outputs < 0 || desc->reduce_outputs > desc->n_outputs)
        return fail(TEGB_ERR_INVALID, "reduce_outputs must be in [0, n_outputs]");
    if (desc->grouped && desc->reduce_outputs != 0 && desc->reduce_outputs != desc->n_outputs)
        return fail(TEGB_ERR_INVALID, "grouped programs reduce all outputs or none");
    if (desc->team > 1 && ((desc->team & (desc->team - 1)) != 0 || desc->team > 1024 ||
                           (desc->team > 32 && desc->team != desc->block_size) ||
                           (desc->team <= 32 && desc->block_size % desc->team != 0)))
        return fail(TEGB_ERR_INVALID, "bad team size");
    if (!drv::ready()) return fail(TEGB_ERR_NO_DEVICE, "CUDA driver unavailable: %s", drv::why.c_str());
    RT_CHECK(cudaSetDevice(device));
    RT_CHECK(cudaFree(0));      // make the primary context current for the driver calls below
    tegb_program *p = new tegb_program();
    p->d = *desc;
    p->device = device;
    CUresult r = drv::ModuleLoadData(&p->mod, m->cubin.data());
    if (r != CUDA_SUCCESS) { delete p; return fail(TEGB_ERR_CUDA, "cuModuleLoadData: %s", drv::errstr(r)); }
    auto bail = [&](int rc) { drv::ModuleUnload(p->mod); if (p->u_stage) cudaFree(p->u_stage); delete p; return rc; };
    r = drv::ModuleGetFunction(&p->f_uniform, p->mod, desc->uniform_kernel);
    if (r != CUDA_SUCCESS) return bail(fail(TEGB_ERR_CUDA, "kernel %s: %s", desc->uniform_kernel, drv::errstr(r)));
    r = drv::ModuleGetFunction(&p->f_main, p->mod, desc->main_kernel);
    if (r != CUDA_SUCCESS) return bail(fail(TEGB_ERR_CUDA, "kernel %s: %s", desc->main_kernel, drv::errstr(r)));
    if (desc->n_tile_conds > 0) {
        if (!desc->tile_kernel || desc->tile_rows < 1 || desc->tile_samples < 1 || desc->n_grid_args != 4 ||
            desc->tile_cols < 32 || desc->tile_cols % 32 != 0 || desc->block_size % desc->tile_cols != 0)
            return bail(fail(TEGB_ERR_INVALID, "tile conditions need a render program, tile_rows >= 1 and a tile kernel"));
        r = drv::ModuleGetFunction(&p->f_tiles, p->mod, desc->tile_kernel);
        if (r != CUDA_SUCCESS) return bail(fail(TEGB_ERR_CUDA, "kernel %s: %s", desc->tile_kernel, drv::errstr(r)));
    }
    if (desc->grouped && desc->accumulate && !desc->reduce_outputs) {
        // accumulating grouped image programs also carry <name>_gather (same module, name derived from <name>_main);
        // modules generated without one simply cannot take tegb_render_groups_gather
        std::string nm(desc->main_kernel);
        const std::string suffix("_main");
        if (nm.size() > suffix.size() && nm.compare(nm.size() - suffix.size(), suffix.size(), suffix) == 0) {
            nm.replace(nm.size() - suffix.size(), suffix.size(), "_gather");
            if (drv::ModuleGetFunction(&p->f_gather, p->mod, nm.c_str()) != CUDA_SUCCESS) p->f_gather = nullptr;
        }
    }
    size_t want = (size_t)(desc->n_uniform > 0 ? desc->n_uniform : 1) * desc->elem_size;
    if (!desc->grouped) {
        r = drv::ModuleGetGlobal(&p->u_const, &p->u_const_bytes, p->mod, desc->uniform_symbol);
        if (r != CUDA_SUCCESS) return bail(fail(TEGB_ERR_CUDA, "symbol %s: %s", desc->uniform_symbol, drv::errstr(r)));
        if (p->u_const_bytes < want) return bail(fail(TEGB_ERR_INVALID, "uniform table smaller than described"));
    }
    if (cudaMalloc(&p->u_stage, want) != cudaSuccess) return bail(fail(TEGB_ERR_CUDA, "cudaMalloc(uniform stage) failed"));
    *out = p;
    return TEGB_OK;
}

// Entry of every launch path (p->mu held): order this launch after the handle's previous one when the stream changed.
static int begin_launch(tegb_program *p, CUstream st) {
    if (p->ev_valid && p->ev_stream != st) {
        RT_CHECK(cudaStreamWaitEvent((cudaStream_t)st, p->ev_last, 0));
        // the published uniform table belongs to the previous stream's launch order: re-publish on this one
        p->uniform_valid = false;
    }
    return TEGB_OK;
}

// Exit of every launch path: remember where the handle's scratch was last used.  (Not while the stream is being
// captured into a CUDA graph: events recorded there cannot be waited on from outside the capture.)
static int end_launch(tegb_program *p, CUstream st) {
    cudaStreamCaptureStatus cs = cudaStreamCaptureStatusNone;
    if (cudaStreamIsCapturing((cudaStream_t)st, &cs) != cudaSuccess) { cudaGetLastError(); cs = cudaStreamCaptureStatusNone; }
    if (cs != cudaStreamCaptureStatusNone) { p->ev_valid = false; return TEGB_OK; }
    if (!p->ev_last) RT_CHECK(cudaEventCreateWithFlags(&p->ev_last, cudaEventDisableTiming));
    RT_CHECK(cudaEventRecord(p->ev_last, (cudaStream_t)st));
    p->ev_valid = true;
    p->ev_stream = st;
    return TEGB_OK;
}

// measurement hooks around the main kernel of a launch (no-ops unless enabled; never inside a stream capture)
static void time_begin(tegb_program *p, CUstream st) {
    if (p->timing) { cudaEventRecord(p->ev_t0, (cudaStream_t)st); p->timing_valid = false; }
}
static void time_end(tegb_program *p, CUstream st) {
    if (p->timing) { cudaEventRecord(p->ev_t1, (cudaStream_t)st); p->timing_valid = true; }
}

extern "C" int tegb_program_enable_timing(tegb_program *p, int on) {
    if (!p) return fail(TEGB_ERR_INVALID, "tegb_program_enable_timing: null program");
    std::lock_guard<std::recursive_mutex> lock(p->mu);
    RT_CHECK(cudaSetDevice(p->device));
    if (on && !p->ev_t0) {
        RT_CHECK(cudaEventCreate(&p->ev_t0));
        RT_CHECK(cudaEventCreate(&p->ev_t1));
    }
    p->timing = on != 0;
    p->timing_valid = false;
    return TEGB_OK;
}

extern "C" int tegb_program_main_kernel_ms(tegb_program *p, double *ms) {
    if (!p || !ms) return fail(TEGB_ERR_INVALID, "tegb_program_main_kernel_ms: null argument");
    std::lock_guard<std::recursive_mutex> lock(p->mu);
    if (!p->timing || !p->timing_valid) return fail(TEGB_ERR_INVALID, "no timed launch (tegb_program_enable_timing, then launch)");
    RT_CHECK(cudaSetDevice(p->device));
    RT_CHECK(cudaEventSynchronize(p->ev_t1));
    float f = 0;
    RT_CHECK(cudaEventElapsedTime(&f, p->ev_t0, p->ev_t1));
    *ms = f;
    return TEGB_OK;
}

extern "C" void tegb_program_destroy(tegb_program *p) {
    if (!p) return;
    cudaSetDevice(p->device);
    if (p->ev_last) cudaEventDestroy(p->ev_last);
    if (p->ev_t0) cudaEventDestroy(p->ev_t0);
    if (p->ev_t1) cudaEventDestroy(p->ev_t1);
    for (void *q : p->d_in) if (q) cudaFree(q);
    for (void *q : p->d_out) if (q) cudaFree(q);
    if (p->u_stage) cudaFree(p->u_stage);
    if (p->d_partials) cudaFree(p->d_partials);
    if (p->d_red_done) cudaFree(p->d_red_done);
    if (p->d_utab) cudaFree(p->d_utab);
    if (p->d_tiles) cudaFree(p->d_tiles);
    if (p->d_ranges) cudaFree(p->d_ranges);
    if (p->d_order) cudaFree(p->d_order);
    if (p->d_xe) cudaFree(p->d_xe);
    if (p->d_ye) cudaFree(p->d_ye);
    if (p->mod) drv::ModuleUnload(p->mod);
    delete p;
}

// uniform prologue: 1 thread evaluates the launch-invariant values, then they are published to __constant__
static int run_uniform(tegb_program *p, const double *scalars, CUstream st) {
    const int ns = p->d.n_scalar_args;
    if (ns > 0 && !scalars) return fail(TEGB_ERR_INVALID, "scalars is NULL but the program has scalar arguments");
    if (p->uniform_valid && p->last_stream == st && (int)p->last_scalars.size() == ns &&
        (ns == 0 || memcmp(p->last_scalars.data(), scalars, sizeof(double) * ns) == 0))
        return TEGB_OK;       // same launch-wide arguments as last time: tegU already holds their values
    std::vector<float> sf(ns > 0 ? ns : 1);
    std::vector<double> sd(ns > 0 ? ns : 1);
    std::vector<void *> args;
    for (int i = 0; i < ns; i++) {
        if (!scalars) return fail(TEGB_ERR_INVALID, "scalars is NULL but the program has scalar arguments");
        if (p->d.elem_size == 4) { sf[i] = (float)scalars[i]; args.push_back(&sf[i]); }
        else { sd[i] = scalars[i]; args.push_back(&sd[i]); }
    }
    args.push_back(&p->u_stage);
    CU_CHECK(drv::LaunchKernel(p->f_uniform, 1, 1, 1, 1, 1, 1, 0, st, args.data(), nullptr));
    g_launches++;
    CU_CHECK(drv::MemcpyDtoDAsync(p->u_const, (CUdeviceptr)p->u_stage,
                                  (size_t)(p->d.n_uniform > 0 ? p->d.n_uniform : 1) * p->d.elem_size, st));
    p->last_scalars.assign(scalars, scalars + ns);
    p->uniform_valid = true;
    p->last_stream = st;
    return TEGB_OK;
}


// --------------------------------------------------------------------------------- deterministic sums
static const int RED_THREADS = 256;
static const int RED_MAX_BLOCKS = 1184;     // 8 blocks per SM x 148 SMs
static const int RED_SLICES = 64;

template <typename T>
__device__ __forceinline__ T block_sum_fixed(T v, T *smem) {
    for (int off = 16; off > 0; off >>= 1) v = v + __shfl_down_sync(0xffffffffu, v, off);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (lane == 0) smem[warp] = v;
    __syncthreads();
    T r = 0;
    if (warp == 0) {
        r = lane < (blockDim.x >> 5) ? smem[lane] : (T)0;
        for (int off = 16; off > 0; off >>= 1) r = r + __shfl_down_sync(0xffffffffu, r, off);
    }
    __syncthreads();
    return r;   // valid in thread 0
}

// ---------------------------------------------------------------------- all-reduce over NVLink peer memory
// The reverse-mode gradient of a sharded render is a handful of numbers per rank.  Instead of handing them to a
// collective library after the column sums, the last column-sum kernel does the exchange itself: every rank owns a
// mailbox (cudaMalloc + CUDA IPC, mapped by every peer of the node), block c writes its column sum into slot
// [bank][my rank][c] of EVERY mailbox (value, fence, sequence number), then polls its own mailbox until all ranks'
// slots for column c carry this call's sequence number and adds them in rank order.  One NVLink store latency plus one
// poll; every rank forms the same sum in the same order (bit-identical across ranks and runs).  Two banks alternate
// with the sequence number: a rank can run at most one call ahead of the slowest peer, because finishing call s needs
// every peer's call-s slot, which a peer writes only after its call s-1 kernel completed (stream order).
#define TEGB_MAX_PEERS 16
struct PeerSlot { unsigned long long bits; unsigned long long seq; };
struct PeerView {
    PeerSlot *box[TEGB_MAX_PEERS];      // mailbox of every rank, mapped here
    int rank, world, maxv;
    volatile int *err;                  // (pinned host memory, mapped) set when a peer did not answer within the time-out
    unsigned long long timeout_ns;      // 0: wait for ever
    // The sequence number of the exchange lives on the DEVICE (not in a launch argument): every block of an exchanging
    // kernel reads seq = *d_seq + 1, the last block to finish publishes it.  Stream order makes the next launch see it,
    // so a captured CUDA graph that contains an exchange can be replayed (each replay is a new exchange).
    unsigned long long *d_seq;
    unsigned int *d_done;
};

__device__ __forceinline__ unsigned long long peer_seq_begin(const PeerView &v) {
    return *(volatile unsigned long long *)v.d_seq + 1ull;
}

// called by every block of the exchanging kernel when it is done (all threads)
__device__ __forceinline__ void peer_seq_end(const PeerView &v, unsigned long long seq, unsigned int n_blocks = 0u) {
    if (n_blocks == 0u) n_blocks = gridDim.x;
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence();
        if (atomicAdd(v.d_done, 1u) == n_blocks - 1u) {
            *(volatile unsigned int *)v.d_done = 0u;
            __threadfence();
            *(volatile unsigned long long *)v.d_seq = seq;
        }
    }
}

__device__ __forceinline__ unsigned long long tegb_globaltimer_ns() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}

// called by all 32 lanes of one warp; `value` is taken from lane 0; every lane returns the sum.  Lane p serves peer p
// (stores / polls run in parallel over the ranks), the contributions are then added in rank order.
template <typename T>
__device__ __forceinline__ T peer_exchange_sum(const PeerView &v, int col, T value, unsigned long long seq) {
    const int lane = threadIdx.x & 31;
    const int bank = (int)(seq & 1ull);
    value = __shfl_sync(0xffffffffu, value, 0);
    unsigned long long bits = 0;
    if (sizeof(T) == 4) bits = (unsigned long long)__float_as_uint((float)value);
    else bits = (unsigned long long)__double_as_longlong((double)value);
    const size_t mine = ((size_t)bank * v.world + v.rank) * v.maxv + col;
    if (lane < v.world) *(volatile unsigned long long *)&v.box[lane][mine].bits = bits;
    __threadfence_system();
    if (lane < v.world) *(volatile unsigned long long *)&v.box[lane][mine].seq = seq;
    T x = 0;
    if (lane < v.world) {
        PeerSlot *s = &v.box[v.rank][((size_t)bank * v.world + lane) * v.maxv + col];
        const unsigned long long t0 = tegb_globaltimer_ns();
        bool arrived = true;
        unsigned int spins = 0;
        while (*(volatile unsigned long long *)&s->seq != seq) {
            // a peer that never arrives (crashed rank) must not hang the GPU for ever: after the time-out the sticky
            // error flag is raised (every later launch on this peer group fails) and the result is poisoned below
            if (v.timeout_ns && (++spins & 1023u) == 0 && tegb_globaltimer_ns() - t0 > v.timeout_ns) {
                *v.err = 1;
                arrived = false;
                break;
            }
        }
        __threadfence_system();
        const unsigned long long b = *(volatile unsigned long long *)&s->bits;
        if (sizeof(T) == 4) x = (T)__uint_as_float((unsigned int)b);
        else x = (T)__longlong_as_double((long long)b);
        if (!arrived) x = (T)__longlong_as_double(0x7ff8000000000000ll);       // NaN: the failure shows in the data
    }
    T acc = __shfl_sync(0xffffffffu, x, 0);
    for (int r = 1; r < v.world; ++r) acc = acc + __shfl_sync(0xffffffffu, x, r);
    return acc;
}

// column sums fused with the peer all-reduce: out[c] = sum over ranks (in rank order) of this rank's column sum
template <typename T>
__global__ void __launch_bounds__(RED_THREADS) colsum_stage2_peer(const T *partials, long long n_blocks, T *out,
                                                                  PeerView view) {
    __shared__ T smem[32];
    const unsigned long long seq = peer_seq_begin(view);
    const int c = blockIdx.x;
    T acc = 0;
    for (long long i = threadIdx.x; i < n_blocks; i += RED_THREADS) acc = acc + partials[(long long)c * n_blocks + i];
    T s = block_sum_fixed(acc, smem);
    if (threadIdx.x < 32) {
        const T total = peer_exchange_sum<T>(view, c, s, seq);
        if (threadIdx.x == 0) out[c] = total;
    }
    peer_seq_end(view, seq);
}

// stand-alone form: in-place all-reduce of a device vector (count <= maxv), one warp per element (grid = count blocks)
template <typename T>
__global__ void peer_allreduce_kernel(T *buf, int count, PeerView view) {
    const unsigned long long seq = peer_seq_begin(view);
    const int c = blockIdx.x;
    if (c < count) {
        const T total = peer_exchange_sum<T>(view, c, buf[c], seq);
        if (threadIdx.x == 0) buf[c] = total;
    }
    peer_seq_end(view, seq);
}

struct tegb_peer {
    PeerView view{};
    int device = 0;
    void *d_box = nullptr;
    size_t box_bytes = 0;
    std::vector<void *> opened;
    int *h_err = nullptr;               // pinned + mapped: kernels raise it, the host reads it without synchronising
};

// Checked (one host load) before every exchange is enqueued: once a peer timed out, results were poisoned and the
// group is unusable -- fail loudly instead of handing out more sums.
static int peer_failed(tegb_peer *peer) {
    if (peer && peer->h_err && *(volatile int *)peer->h_err) {
        fail(TEGB_ERR_PEER, "peer all-reduce: a rank did not arrive within the time-out (TEGB200_PEER_TIMEOUT_S); "
                            "the affected sums were set to NaN and this peer group is unusable");
        return 1;
    }
    return 0;
}

template <typename T>
__global__ void __launch_bounds__(RED_THREADS) colsum_stage2(const T *partials, long long n_blocks, T *out) {
    __shared__ T smem[32];
    const int c = blockIdx.x;
    T acc = 0;
    for (long long i = threadIdx.x; i < n_blocks; i += RED_THREADS) acc = acc + partials[(long long)c * n_blocks + i];
    T s = block_sum_fixed(acc, smem);
    if (threadIdx.x == 0) out[c] = s;
}

// partials buffer of a reduce_outputs program: n_outputs x n_blocks elements
static int ensure_partials(tegb_program *p, unsigned long long n_blocks) {
    const size_t need = (size_t)p->d.reduce_outputs * (n_blocks + RED_SLICES) * p->d.elem_size;
    if (need <= p->partials_bytes) return TEGB_OK;
    if (p->d_partials) { cudaFree(p->d_partials); p->d_partials = nullptr; p->partials_bytes = 0; }
    RT_CHECK(cudaMalloc(&p->d_partials, need));
    p->partials_bytes = need;
    return TEGB_OK;
}

// Sum n_cols columns of n_blocks partials each.  Few long columns would leave the GPU idle (6 blocks for
// a 6-parameter gradient), so long columns are first cut into RED_SLICES contiguous slices (one block each),
// then the slice sums are added: still a fixed order for a given n_blocks, i.e. deterministic.

template <typename T>
__global__ void __launch_bounds__(RED_THREADS) colsum_slices(const T *partials, long long n_blocks, long long chunk,
                                                             T *slices) {
    __shared__ T smem[32];
    const int c = blockIdx.y;
    const long long lo = (long long)blockIdx.x * chunk;
    long long hi = lo + chunk;
    if (hi > n_blocks) hi = n_blocks;
    T acc = 0;
    for (long long i = lo + threadIdx.x; i < hi; i += RED_THREADS) acc = acc + partials[(long long)c * n_blocks + i];
    T s = block_sum_fixed(acc, smem);
    if (threadIdx.x == 0) slices[(long long)c * gridDim.x + blockIdx.x] = s;
}

// Both stages in ONE launch: the block that finishes a column's last slice (a counter per column, reset by that block:
// replay-safe) adds the slice sums -- the same additions in the same order as colsum_stage2 over the slices, whoever
// that block is -- and, with peers, runs the exchange.  One launch and one launch gap less on the critical path of
// every gradient.
template <typename T, bool PEER>
__global__ void __launch_bounds__(RED_THREADS) colsum_slices_finish(const T *partials, long long n_blocks, long long chunk,
                                                                    T *slices, unsigned int *done, T *out, PeerView view) {
    __shared__ T smem[32];
    __shared__ int last_;
    unsigned long long seq = 0;
    if (PEER) seq = peer_seq_begin(view);
    const int c = blockIdx.y;
    const long long lo = (long long)blockIdx.x * chunk;
    long long hi = lo + chunk;
    if (hi > n_blocks) hi = n_blocks;
    T acc = 0;
    for (long long i = lo + threadIdx.x; i < hi; i += RED_THREADS) acc = acc + partials[(long long)c * n_blocks + i];
    T s = block_sum_fixed(acc, smem);
    if (threadIdx.x == 0) {
        slices[(long long)c * gridDim.x + blockIdx.x] = s;
        __threadfence();
        last_ = atomicAdd(&done[c], 1u) == gridDim.x - 1u;
    }
    __syncthreads();
    if (last_) {                                   // (block-uniform)
        __threadfence();
        T a2 = 0;
        for (long long i = threadIdx.x; i < (long long)gridDim.x; i += RED_THREADS)
            a2 = a2 + __ldcg(&slices[(long long)c * gridDim.x + i]);
        T t = block_sum_fixed(a2, smem);
        if (PEER) {
            if (threadIdx.x < 32) {
                const T total = peer_exchange_sum<T>(view, c, t, seq);
                if (threadIdx.x == 0) out[c] = total;
            }
        } else if (threadIdx.x == 0) {
            out[c] = t;
        }
        if (threadIdx.x == 0) done[c] = 0u;
    }
    if (PEER) peer_seq_end(view, seq, gridDim.x * gridDim.y);
}

template <typename T>
static void sum_columns(const T *partials, int n_cols, unsigned long long n_blocks, T *out, T *slices, cudaStream_t st,
                        unsigned int *done) {
    if (n_blocks >= 16384 && slices && done) {
        const long long chunk = (long long)((n_blocks + RED_SLICES - 1) / RED_SLICES);
        colsum_slices_finish<T, false><<<dim3(RED_SLICES, n_cols), RED_THREADS, 0, st>>>(partials, (long long)n_blocks, chunk,
                                                                                         slices, done, out, PeerView{});
        g_launches++;
    } else if (n_blocks >= 16384 && slices) {
        const long long chunk = (long long)((n_blocks + RED_SLICES - 1) / RED_SLICES);
        colsum_slices<T><<<dim3(RED_SLICES, n_cols), RED_THREADS, 0, st>>>(partials, (long long)n_blocks, chunk, slices);
        colsum_stage2<T><<<n_cols, RED_THREADS, 0, st>>>(slices, (long long)RED_SLICES, out);
        g_launches += 2;
    } else {
        colsum_stage2<T><<<n_cols, RED_THREADS, 0, st>>>(partials, (long long)n_blocks, out);
        g_launches++;
    }
}

template <typename T>
static void sum_columns_peer(const T *partials, int n_cols, unsigned long long n_blocks, T *out, T *slices, cudaStream_t st,
                             tegb_peer *peer, unsigned int *done) {
    if (n_blocks >= 16384 && slices && done) {
        const long long chunk = (long long)((n_blocks + RED_SLICES - 1) / RED_SLICES);
        colsum_slices_finish<T, true><<<dim3(RED_SLICES, n_cols), RED_THREADS, 0, st>>>(partials, (long long)n_blocks, chunk,
                                                                                        slices, done, out, peer->view);
        g_launches++;
    } else if (n_blocks >= 16384 && slices) {
        const long long chunk = (long long)((n_blocks + RED_SLICES - 1) / RED_SLICES);
        colsum_slices<T><<<dim3(RED_SLICES, n_cols), RED_THREADS, 0, st>>>(partials, (long long)n_blocks, chunk, slices);
        colsum_stage2_peer<T><<<n_cols, RED_THREADS, 0, st>>>(slices, (long long)RED_SLICES, out, peer->view);
        g_launches += 2;
    } else {
        colsum_stage2_peer<T><<<n_cols, RED_THREADS, 0, st>>>(partials, (long long)n_blocks, out, peer->view);
        g_launches++;
    }
}

static int finish_partials(tegb_program *p, unsigned long long n_blocks, void *out, cudaStream_t st) {
    // slice sums live behind the partials (ensure_partials reserves the room)
    char *slices = (char *)p->d_partials + (size_t)p->d.reduce_outputs * n_blocks * p->d.elem_size;
    if (!p->d_red_done) {                        // per-column "slices done" counters of the one-launch column sum
        RT_CHECK(cudaMalloc(&p->d_red_done, (size_t)p->d.reduce_outputs * sizeof(unsigned int)));
        RT_CHECK(cudaMemsetAsync(p->d_red_done, 0, (size_t)p->d.reduce_outputs * sizeof(unsigned int), st));
    }
    unsigned int *done = p->d_red_done;
    if (p->peer) {
        if (p->d.reduce_outputs > p->peer->view.maxv) return fail(TEGB_ERR_INVALID, "peer mailbox smaller than the gradient");
        if (peer_failed(p->peer)) return TEGB_ERR_PEER;
        if (p->d.elem_size == 4)
            sum_columns_peer<float>((const float *)p->d_partials, p->d.reduce_outputs, n_blocks, (float *)out, (float *)slices, st, p->peer, done);
        else
            sum_columns_peer<double>((const double *)p->d_partials, p->d.reduce_outputs, n_blocks, (double *)out, (double *)slices, st, p->peer, done);
        RT_CHECK(cudaGetLastError());
        return TEGB_OK;
    }
    if (p->d.elem_size == 4)
        sum_columns<float>((const float *)p->d_partials, p->d.reduce_outputs, n_blocks, (float *)out, (float *)slices, st, done);
    else
        sum_columns<double>((const double *)p->d_partials, p->d.reduce_outputs, n_blocks, (double *)out, (double *)slices, st, done);
    RT_CHECK(cudaGetLastError());
    return TEGB_OK;
}

extern "C" int tegb_program_launch(tegb_program *p, const double *scalars, const void *const *in_ptrs,
                                   void *const *out_ptrs, int64_t n, void *stream) {
    if (!p || !out_ptrs || n < 0) return fail(TEGB_ERR_INVALID, "tegb_program_launch: bad argument");
    if (p->d.n_grid_args != 0 || p->d.grouped) return fail(TEGB_ERR_INVALID, "render program: use tegb_render_launch");
    if (p->d.n_array_args > 0 && !in_ptrs) return fail(TEGB_ERR_INVALID, "in_ptrs is NULL");
    if (n == 0) return TEGB_OK;
    std::lock_guard<std::recursive_mutex> lock(p->mu);
    RT_CHECK(cudaSetDevice(p->device));
    CUstream st = (CUstream)stream;
    int rc = begin_launch(p, st);
    if (rc) return rc;
    rc = run_uniform(p, scalars, st);
    if (rc) return rc;
    std::vector<const void *> ins(p->d.n_array_args > 0 ? p->d.n_array_args : 1);
    std::vector<void *> outs(p->d.n_outputs);
    std::vector<void *> args;
    for (int i = 0; i < p->d.n_array_args; i++) { ins[i] = in_ptrs[i]; args.push_back(&ins[i]); }
    const unsigned bs = (unsigned)p->d.block_size;
    const long long team = p->d.team > 1 ? p->d.team : 1;
    const long long vec = p->d.vec > 1 ? p->d.vec : 1;
    if (vec > 1) {
        for (int i = 0; i < p->d.n_array_args; i++)
            if (((uintptr_t)in_ptrs[i]) & 15) return fail(TEGB_ERR_INVALID, "vectorised program: input %d is not 16-byte aligned", i);
        for (int k = 0; k < p->d.n_outputs; k++)
            if (((uintptr_t)out_ptrs[k]) & 15) return fail(TEGB_ERR_INVALID, "vectorised program: output %d is not 16-byte aligned", k);
    }
    const long long blocks = (((n + vec - 1) / vec) * team + bs - 1) / bs;
    if (blocks > 2147483647LL) return fail(TEGB_ERR_INVALID, "too many instances for one launch");
    const int n_store = p->d.n_outputs - p->d.reduce_outputs;
    for (int k = 0; k < n_store; k++) { outs[k] = out_ptrs[k]; args.push_back(&outs[k]); }
    if (p->d.reduce_outputs) {
        rc = ensure_partials(p, (unsigned long long)blocks);
        if (rc) return rc;
        args.push_back(&p->d_partials);
    }
    long long nn = n;
    args.push_back(&nn);
    time_begin(p, st);
    CU_CHECK(drv::LaunchKernel(p->f_main, (unsigned)blocks, 1, 1, bs, 1, 1, 0, st, args.data(), nullptr));
    time_end(p, st);
    g_launches++;
    if (p->d.reduce_outputs) {
        rc = finish_partials(p, (unsigned long long)blocks, out_ptrs[n_store], (cudaStream_t)st);
        if (rc) return rc;
    }
    return end_launch(p, st);
}

static int ensure_staging(tegb_program *p, int64_t n) {
    if (n <= p->cap && (int)p->d_in.size() == p->d.n_array_args && (int)p->d_out.size() == p->d.n_outputs)
        return TEGB_OK;
    for (void *q : p->d_in) if (q) cudaFree(q);
    for (void *q : p->d_out) if (q) cudaFree(q);
    p->d_in.assign(p->d.n_array_args, nullptr);
    p->d_out.assign(p->d.n_outputs, nullptr);
    p->cap = 0;
    const size_t bytes = (size_t)n * p->d.elem_size;
    // a reduce_outputs program delivers its n_outputs sums into d_out[0]
    const size_t out_bytes = (size_t)(n > p->d.n_outputs ? n : p->d.n_outputs) * p->d.elem_size;
    for (auto &q : p->d_in) RT_CHECK(cudaMalloc(&q, bytes));
    for (auto &q : p->d_out) RT_CHECK(cudaMalloc(&q, out_bytes));
    p->cap = n;
    return TEGB_OK;
}

extern "C" int tegb_program_eval_host(tegb_program *p, const double *scalars, const void *const *host_in,
                                      void *const *host_out, int64_t n) {
    if (!p || !host_out || n < 0) return fail(TEGB_ERR_INVALID, "tegb_program_eval_host: bad argument");
    if (p->d.n_array_args > 0 && !host_in) return fail(TEGB_ERR_INVALID, "host_in is NULL");
    if (n == 0) return TEGB_OK;
    std::lock_guard<std::recursive_mutex> lock(p->mu);
    RT_CHECK(cudaSetDevice(p->device));
    int rc = ensure_staging(p, n);
    if (rc) return rc;
    const size_t bytes = (size_t)n * p->d.elem_size;
    cudaStream_t st = 0;
    for (int i = 0; i < p->d.n_array_args; i++)
        RT_CHECK(cudaMemcpyAsync(p->d_in[i], host_in[i], bytes, cudaMemcpyHostToDevice, st));
    rc = tegb_program_launch(p, scalars, (const void *const *)p->d_in.data(), p->d_out.data(), n, st);
    if (rc) return rc;
    const int n_store = p->d.n_outputs - p->d.reduce_outputs;
    for (int k = 0; k < n_store; k++)
        RT_CHECK(cudaMemcpyAsync(host_out[k], p->d_out[k], bytes, cudaMemcpyDeviceToHost, st));
    if (p->d.reduce_outputs)    // the sums were written to the start of d_out[n_store]; host_out[n_store] receives them
        RT_CHECK(cudaMemcpyAsync(host_out[n_store], p->d_out[n_store], (size_t)p->d.reduce_outputs * p->d.elem_size,
                                 cudaMemcpyDeviceToHost, st));
    RT_CHECK(cudaStreamSynchronize(st));
    return TEGB_OK;
}

// np.linspace(lo, hi, res + 1) as numpy computes it (float64): step = (hi-lo)/res; y[j] = j*step + lo;
// y[res] = hi  (numpy/_core/function_base.py; tests/plot.py:11-12 is the caller being reproduced)
static void linspace_edges(double lo, double hi, int res, int elem_size, std::vector<char> &out) {
    out.resize((size_t)(res + 1) * elem_size);
    const double delta = hi - lo;
    const double step = delta / (double)res;
    for (int j = 0; j <= res; j++) {
        double v;
        if (step == 0.0) { v = ((double)j / (double)res) * delta + lo; }
        else { volatile double t = (double)j * step; v = t + lo; }
        if (j == res) v = hi;
        if (elem_size == 4) ((float *)out.data())[j] = (float)v;
        else ((double *)out.data())[j] = v;
    }
}

static int ensure_edges(tegb_program *p, const tegb_grid *g, CUstream st) {
    const bool same = p->grid_valid && p->grid_cached.x_lo == g->x_lo && p->grid_cached.x_hi == g->x_hi &&
                      p->grid_cached.y_lo == g->y_lo && p->grid_cached.y_hi == g->y_hi &&
                      p->grid_cached.res_x == g->res_x && p->grid_cached.res_y == g->res_y;
    if (same) return TEGB_OK;
    // the previous tables may still be in use by an earlier launch on this stream
    RT_CHECK(cudaStreamSynchronize((cudaStream_t)st));
    if (p->d_xe) { cudaFree(p->d_xe); p->d_xe = nullptr; }
    if (p->d_ye) { cudaFree(p->d_ye); p->d_ye = nullptr; }
    linspace_edges(g->x_lo, g->x_hi, g->res_x, p->d.elem_size, p->h_xe);
    linspace_edges(g->y_lo, g->y_hi, g->res_y, p->d.elem_size, p->h_ye);
    RT_CHECK(cudaMalloc(&p->d_xe, p->h_xe.size()));
    RT_CHECK(cudaMalloc(&p->d_ye, p->h_ye.size()));
    RT_CHECK(cudaMemcpyAsync(p->d_xe, p->h_xe.data(), p->h_xe.size(), cudaMemcpyHostToDevice, (cudaStream_t)st));
    RT_CHECK(cudaMemcpyAsync(p->d_ye, p->h_ye.data(), p->h_ye.size(), cudaMemcpyHostToDevice, (cudaStream_t)st));
    p->grid_cached = *g;
    p->grid_valid = true;
    p->ranges_valid = false;
    return TEGB_OK;
}

// Blocks of a tiled render launch that serve the queue of unresolved rows (they come first in the grid, so they start in the
// first wave): TEGB200_QUEUE_WORKERS per SM (default 2 of the 8 resident 128-thread blocks; the rest of the wave renders;
// measured on the 4096^2 triangle: 2 -> 28.5 us, 4 -> 30.7 us, 6 -> 32.7 us for classification + value kernel).
static int queue_workers(int device) {
    static int per_sm = -1;
    if (per_sm < 0) {
        per_sm = 2;
        if (const char *e = getenv("TEGB200_QUEUE_WORKERS")) per_sm = atoi(e) > 0 ? atoi(e) : 1;
    }
    int sms = 148;
    if (cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device) != cudaSuccess) { cudaGetLastError(); sms = 148; }
    return sms * per_sm;
}

// Tile classification support (generated <name>_tiles kernels read this table): for every tile row (x axis, entries
// [0, n_ty)) and every tile column (y axis, entries [n_ty, n_ty + n_tx)) the exact [min, max] over its pixels of the
// lower box edge, the upper box edge and the first / last midpoint abscissa -- computed by walking the pixels with
// the very operations the generated code uses (lb + step * (i + 0.5), step = (ub - lb) / N; this file is compiled
// with --fmad=false like the generated units) -- plus a poison sum that is finite iff every edge, width and abscissa
// of the walk is.  7 values per entry: lb_lo lb_hi ub_lo ub_hi s_lo s_hi poison.  One warp per entry.
template <typename T>
__global__ void tile_ranges(const T *__restrict__ xe, const T *__restrict__ ye, int res_y, int row0, int n_rows,
                            int tile_rows, int tile_cols, int num_samples, int n_tx, int n_ty, int row_stride,
                            int row_phase, T *__restrict__ out) {
    const int e = (int)((blockIdx.x * blockDim.x + threadIdx.x) >> 5);
    const int lane = threadIdx.x & 31;
    if (e >= n_tx + n_ty) return;
    const bool is_x = e < n_ty;
    const T *tab = is_x ? xe : ye;
    int first, last;
    if (is_x) { first = row0 + (e * row_stride + row_phase) * tile_rows; last = min(first + tile_rows, row0 + n_rows); }
    else { first = (e - n_ty) * tile_cols; last = min(first + tile_cols, res_y); }
    if (first >= last) return;          // a padding tile column beyond the image: never read
    T lb_lo = tab[first], lb_hi = tab[first], ub_lo = tab[first + 1], ub_hi = tab[first + 1];
    T s_lo = tab[first], s_hi = tab[first], pz = (T)0;
    const T half = (T)0.5;
    for (int q = first + lane; q < last; q += 32) {
        const T lb = tab[q], ub = tab[q + 1];
        lb_lo = fmin(lb_lo, lb); lb_hi = fmax(lb_hi, lb);
        ub_lo = fmin(ub_lo, ub); ub_hi = fmax(ub_hi, ub);
        const T st = ((T)(ub - lb)) / (T)num_samples;
        const T sa = lb + st * ((T)0u + half), sb = lb + st * ((T)(unsigned)(num_samples - 1) + half);
        s_lo = fmin(s_lo, fmin(sa, sb)); s_hi = fmax(s_hi, fmax(sa, sb));
        pz = pz + lb; pz = pz + ub; pz = pz + (ub - lb); pz = pz + sa; pz = pz + sb;
    }
    for (int off = 16; off > 0; off >>= 1) {
        lb_lo = fmin(lb_lo, __shfl_xor_sync(0xffffffffu, lb_lo, off)); lb_hi = fmax(lb_hi, __shfl_xor_sync(0xffffffffu, lb_hi, off));
        ub_lo = fmin(ub_lo, __shfl_xor_sync(0xffffffffu, ub_lo, off)); ub_hi = fmax(ub_hi, __shfl_xor_sync(0xffffffffu, ub_hi, off));
        s_lo = fmin(s_lo, __shfl_xor_sync(0xffffffffu, s_lo, off)); s_hi = fmax(s_hi, __shfl_xor_sync(0xffffffffu, s_hi, off));
        pz = pz + __shfl_xor_sync(0xffffffffu, pz, off);
    }
    if (lane == 0) {
        T *o = out + (size_t)e * 7;
        o[0] = lb_lo; o[1] = lb_hi; o[2] = ub_lo; o[3] = ub_hi; o[4] = s_lo; o[5] = s_hi; o[6] = pz;
    }
}

static int ensure_ranges(tegb_program *p, const tegb_grid *g, int n_tx, int n_ty, int row_stride, int row_phase,
                         cudaStream_t st) {
    const tegb_grid &c = p->ranges_grid;
    if (p->ranges_valid && p->ranges_stride == row_stride && p->ranges_phase == row_phase && c.x_lo == g->x_lo && c.x_hi == g->x_hi && c.y_lo == g->y_lo && c.y_hi == g->y_hi &&
        c.res_x == g->res_x && c.res_y == g->res_y && c.row_begin == g->row_begin && c.row_end == g->row_end)
        return TEGB_OK;
    const size_t need = (size_t)(n_tx + n_ty) * 7 * p->d.elem_size;
    if (need > p->ranges_bytes) {
        RT_CHECK(cudaStreamSynchronize(st));
        if (p->d_ranges) { cudaFree(p->d_ranges); p->d_ranges = nullptr; p->ranges_bytes = 0; }
        RT_CHECK(cudaMalloc(&p->d_ranges, need));
        p->ranges_bytes = need;
    }
    const int n_rows = g->row_end - g->row_begin;
    const unsigned tb = 128, tg = (unsigned)(((size_t)(n_tx + n_ty) * 32 + tb - 1) / tb);
    if (p->d.elem_size == 4)
        tile_ranges<float><<<tg, tb, 0, st>>>((const float *)p->d_xe, (const float *)p->d_ye, g->res_y, g->row_begin, n_rows,
                                             p->d.tile_rows, p->d.tile_cols, p->d.tile_samples, n_tx, n_ty, row_stride, row_phase, (float *)p->d_ranges);
    else
        tile_ranges<double><<<tg, tb, 0, st>>>((const double *)p->d_xe, (const double *)p->d_ye, g->res_y, g->row_begin, n_rows,
                                              p->d.tile_rows, p->d.tile_cols, p->d.tile_samples, n_tx, n_ty, row_stride, row_phase, (double *)p->d_ranges);
    RT_CHECK(cudaGetLastError());
    g_launches++;
    p->ranges_grid = *g;
    p->ranges_stride = row_stride;
    p->ranges_phase = row_phase;
    p->ranges_valid = true;
    return TEGB_OK;
}

// per-group sums of a grouped reduce launch: partials [group][component][tile] -> out[group][component].  With peer
// mailboxes attached (tegb_program_set_peer_groups) `out` is the whole [n_groups_total][n_outputs] result: every
// column is exchanged, the columns of groups this launch did not cover contribute zero.
template <typename T>
__global__ void __launch_bounds__(RED_THREADS) colsum_groups_peer(const T *partials, long long n_tiles, int group_begin,
                                                                  int group_count, int K, int total_cols,
                                                                  const int *__restrict__ local_of, T *out, PeerView view) {
    // One WARP per column of the all-reduced result (global group, component): a column has only a few hundred tile
    // partials, and a scene has thousands of columns (7 per triangle) -- a block per column made the exchange of 14336
    // columns at 8 GPUs cost 0.4 ms.  Lanes sum the column's partials strided, a fixed shuffle tree combines them, the
    // same warp exchanges the sum with the peers and stores the total.
    const unsigned long long seq = peer_seq_begin(view);
    const int lane = threadIdx.x & 31;
    const int c = blockIdx.x * (RED_THREADS / 32) + (threadIdx.x >> 5);
    if (c < total_cols) {                   // warp-uniform
        const int g = c / K, k = c - g * K;
        const int lg = local_of ? local_of[g] : g;       // this rank's index of the group (-1: not launched here)
        T acc = 0;
        if (lg >= group_begin && lg < group_begin + group_count) {
            const T *col = partials + ((long long)(lg - group_begin) * K + k) * n_tiles;
            for (long long i = lane; i < n_tiles; i += 32) acc = acc + col[i];
        }
        for (int off = 16; off > 0; off >>= 1) acc = acc + __shfl_down_sync(0xffffffffu, acc, off);
        const T total = peer_exchange_sum<T>(view, c, acc, seq);
        if (lane == 0) out[c] = total;
    }
    peer_seq_end(view, seq);
}

static int finish_group_partials(tegb_program *p, unsigned long long tiles, int group_begin, int group_count, void *out,
                                 cudaStream_t st) {
    const int K = p->d.n_outputs;
    if (p->peer && p->n_groups_total > 0) {
        const int total = p->n_groups_total * K;
        if (total > p->peer->view.maxv) return fail(TEGB_ERR_INVALID, "peer mailbox smaller than the [groups][outputs] gradient");
        if (peer_failed(p->peer)) return TEGB_ERR_PEER;
        const int cblocks = (total + RED_THREADS / 32 - 1) / (RED_THREADS / 32);
        if (p->d.elem_size == 4)
            colsum_groups_peer<float><<<cblocks, RED_THREADS, 0, st>>>((const float *)p->d_partials, (long long)tiles, group_begin,
                                                                       group_count, K, total, p->group_map, (float *)out, p->peer->view);
        else
            colsum_groups_peer<double><<<cblocks, RED_THREADS, 0, st>>>((const double *)p->d_partials, (long long)tiles, group_begin,
                                                                        group_count, K, total, p->group_map, (double *)out, p->peer->view);
        g_launches++;
        RT_CHECK(cudaGetLastError());
        return TEGB_OK;
    }
    if (group_count == 0) return TEGB_OK;
    const int cols = group_count * K;
    if (p->d.elem_size == 4)
        colsum_stage2<float><<<cols, RED_THREADS, 0, st>>>((const float *)p->d_partials, (long long)tiles, (float *)out);
    else
        colsum_stage2<double><<<cols, RED_THREADS, 0, st>>>((const double *)p->d_partials, (long long)tiles, (double *)out);
    g_launches++;
    RT_CHECK(cudaGetLastError());
    return TEGB_OK;
}

// launch geometry of a grouped render: blocks of block_size columns x tile_rows rows over the largest window
struct GroupGeom { unsigned gx, gy; int tile_rows, n_tx, n_ty; };

static GroupGeom group_geom(const tegb_program *p, const tegb_grid *g, int max_rows, int max_cols, bool aligned = false) {
    GroupGeom q;
    const unsigned bs = (unsigned)p->d.block_size;
    q.gx = (unsigned)((max_cols + bs - 1) / bs);
    const int band_rows = g->row_end - g->row_begin;
    q.tile_rows = p->d.tile_rows > 0 ? p->d.tile_rows : 1;
    const int win_rows = max_rows < band_rows ? max_rows : band_rows;
    q.gy = (unsigned)((win_rows + q.tile_rows - 1) / q.tile_rows);
    q.n_tx = (int)q.gx * (p->d.tile_cols > 0 ? p->d.block_size / p->d.tile_cols : 1);
    q.n_ty = (int)q.gy;
    if (aligned) {
        // gather launches: the tiles of a window are those of the ABSOLUTE grid it reaches into (a window that starts
        // inside a tile spans one more); the launch itself covers the image (blocks of block_size columns x tile_rows rows)
        const int tc = p->d.tile_cols > 0 ? p->d.tile_cols : p->d.block_size;
        q.n_tx = (max_cols + tc - 1) / tc + 1;
        q.n_ty = (win_rows + q.tile_rows - 1) / q.tile_rows + 1;
        q.gx = (unsigned)((g->res_y + bs - 1) / bs);
        q.gy = (unsigned)((band_rows + q.tile_rows - 1) / q.tile_rows);
    }
    return q;
}

// Per-group prologue of a grouped render: the uniform rows and (tile-culled programs) the tri-states of every (group,
// window tile), for the groups' CURRENT parameters.  Tables are indexed by absolute group number, so one prologue
// over all groups can serve several launches over sub-ranges (tegb_render_groups_prepare).
static int groups_prologue(tegb_program *p, const void *const *group_ptrs, int group_begin, int group_count,
                           const tegb_grid *g, const int32_t *windows, int max_rows, int max_cols, CUstream st,
                           bool aligned_tiles = false) {
    int rc = ensure_edges(p, g, st);
    if (rc) return rc;
    const GroupGeom q = group_geom(p, g, max_rows, max_cols, aligned_tiles);
    const int tab_groups = group_begin + group_count;
    const int nu = p->d.n_uniform > 0 ? p->d.n_uniform : 1;
    const size_t need = (size_t)tab_groups * nu * p->d.elem_size;
    if (need > p->utab_bytes) {
        RT_CHECK(cudaStreamSynchronize((cudaStream_t)st));
        if (p->d_utab) { cudaFree(p->d_utab); p->d_utab = nullptr; p->utab_bytes = 0; }
        RT_CHECK(cudaMalloc(&p->d_utab, need));
        p->utab_bytes = need;
    }
    {
        std::vector<const void *> gp(p->d.n_scalar_args > 0 ? p->d.n_scalar_args : 1);
        std::vector<void *> args;
        for (int i = 0; i < p->d.n_scalar_args; i++) { gp[i] = group_ptrs[i]; args.push_back(&gp[i]); }
        args.push_back(&group_begin);
        args.push_back(&group_count);
        args.push_back(&p->d_utab);
        CU_CHECK(drv::LaunchKernel(p->f_uniform, (unsigned)((group_count + 127) / 128), 1, 1, 128, 1, 1, 0, st,
                                   args.data(), nullptr));
        g_launches++;
    }
    if (p->d.n_tile_conds > 0) {
        // classify every (group, window tile): one thread per tile
        int n_tx = q.n_tx, n_ty = q.n_ty, tile_rows = q.tile_rows, tabg = tab_groups;
        int res_y = g->res_y, row0 = g->row_begin, row1 = g->row_end;
        const size_t need_t = (size_t)p->d.n_tile_conds * tab_groups * n_tx * n_ty * p->d.elem_size;
        if (need_t > p->tiles_bytes) {
            RT_CHECK(cudaStreamSynchronize((cudaStream_t)st));
            if (p->d_tiles) { cudaFree(p->d_tiles); p->d_tiles = nullptr; p->tiles_bytes = 0; }
            RT_CHECK(cudaMalloc(&p->d_tiles, need_t));
            p->tiles_bytes = need_t;
        }
        std::vector<const void *> gp2(p->d.n_scalar_args > 0 ? p->d.n_scalar_args : 1);
        std::vector<void *> targs;
        for (int i = 0; i < p->d.n_scalar_args; i++) { gp2[i] = group_ptrs[i]; targs.push_back(&gp2[i]); }
        targs.push_back(&p->d_xe);
        targs.push_back(&p->d_ye);
        targs.push_back(&res_y);
        targs.push_back(&row0);
        targs.push_back(&row1);
        targs.push_back((void *)&windows);
        targs.push_back(&group_begin);
        targs.push_back(&group_count);
        targs.push_back(&tile_rows);
        targs.push_back(&n_tx);
        targs.push_back(&n_ty);
        targs.push_back(&tabg);
        int aligned = aligned_tiles ? 1 : 0;
        targs.push_back(&aligned);
        targs.push_back(&p->d_tiles);
        const unsigned long long threads = (unsigned long long)group_count * n_tx * n_ty;
        const unsigned tb = 128, tg = (unsigned)((threads + tb - 1) / tb);
        CU_CHECK(drv::LaunchKernel(p->f_tiles, tg, 1, 1, tb, 1, 1, 0, st, targs.data(), nullptr));
        g_launches++;
    }
    p->prep_valid = !aligned_tiles;              // (a gather launch classifies for itself, every time)
    p->prep_begin = group_begin;
    p->prep_count = group_count;
    p->prep_grid = *g;
    p->prep_windows = windows;
    p->prep_max_rows = max_rows;
    p->prep_max_cols = max_cols;
    return TEGB_OK;
}

static int groups_launch_impl(tegb_program *p, const void *const *group_ptrs, int group_begin, int group_count,
                              const tegb_grid *g, const int32_t *windows, int max_rows, int max_cols,
                              const void *const *pixel_ptrs, void *const *out_ptrs, void *stream, bool prepared) {
    if (!p || !g || !windows || !out_ptrs) return fail(TEGB_ERR_INVALID, "tegb_render_groups_launch: bad argument");
    if (!p->d.grouped || p->d.n_grid_args != 4) return fail(TEGB_ERR_INVALID, "not a grouped render program");
    if (p->d.n_scalar_args > 0 && !group_ptrs) return fail(TEGB_ERR_INVALID, "group_ptrs is NULL");
    if (p->d.n_pixel_args > 0 && !pixel_ptrs) return fail(TEGB_ERR_INVALID, "pixel_ptrs is NULL");
    if (group_begin < 0 || group_count < 0 || max_rows < 1 || max_cols < 1)
        return fail(TEGB_ERR_INVALID, "bad group range or window bounds");
    if (g->res_x < 1 || g->res_y < 1 || g->row_begin < 0 || g->row_end > g->res_x || g->row_begin > g->row_end)
        return fail(TEGB_ERR_INVALID, "bad grid");
    const bool peer_groups = p->d.reduce_outputs && p->peer && p->n_groups_total > 0;
    if (peer_groups && !p->group_map && (group_begin + group_count > p->n_groups_total))
        return fail(TEGB_ERR_INVALID, "group range exceeds the total declared with tegb_program_set_peer_groups");
    if ((group_count == 0 || g->row_end == g->row_begin) && !peer_groups) return TEGB_OK;
    if (group_count > 65535 || max_rows > 65535 * 16) return fail(TEGB_ERR_INVALID, "too many groups or rows in one launch");
    std::lock_guard<std::recursive_mutex> lock(p->mu);
    RT_CHECK(cudaSetDevice(p->device));
    CUstream st = (CUstream)stream;
    int rc = begin_launch(p, st);
    if (rc) return rc;
    if (group_count == 0 || g->row_end == g->row_begin) {
        // nothing to render on this rank, but the other ranks wait for its (zero) contribution
        rc = finish_group_partials(p, 0, 0, 0, out_ptrs[0], (cudaStream_t)st);
        if (rc) return rc;
        return end_launch(p, st);
    }
    if (prepared) {
        const tegb_grid &c = p->prep_grid;
        const bool same = p->prep_valid && group_begin >= p->prep_begin &&
                          group_begin + group_count <= p->prep_begin + p->prep_count && p->prep_windows == windows &&
                          p->prep_max_rows == max_rows && p->prep_max_cols == max_cols && c.x_lo == g->x_lo &&
                          c.x_hi == g->x_hi && c.y_lo == g->y_lo && c.y_hi == g->y_hi && c.res_x == g->res_x &&
                          c.res_y == g->res_y && c.row_begin == g->row_begin && c.row_end == g->row_end;
        if (!same) return fail(TEGB_ERR_INVALID, "tegb_render_groups_launch_prepared: no matching tegb_render_groups_prepare "
                                                 "(same grid, windows and window bounds, covering this group range)");
    } else {
        rc = groups_prologue(p, group_ptrs, group_begin, group_count, g, windows, max_rows, max_cols, st);
        if (rc) return rc;
    }
    const GroupGeom q = group_geom(p, g, max_rows, max_cols);
    const unsigned bs = (unsigned)p->d.block_size;
    const unsigned gx = q.gx, gy = q.gy, gz = (unsigned)group_count;
    std::vector<const void *> pix(p->d.n_pixel_args > 0 ? p->d.n_pixel_args : 1);
    std::vector<void *> outs(p->d.n_outputs);
    std::vector<void *> args;
    int res_y = g->res_y, row0 = g->row_begin, row1 = g->row_end, tile_rows = q.tile_rows;
    int tab_groups = p->prep_begin + p->prep_count;
    args.push_back(&p->d_xe);
    args.push_back(&p->d_ye);
    args.push_back(&res_y);
    args.push_back(&row0);
    args.push_back(&row1);
    args.push_back((void *)&windows);
    args.push_back(&group_begin);
    args.push_back(&p->d_utab);
    args.push_back(&tile_rows);
    if (p->d.n_tile_conds > 0) { args.push_back(&p->d_tiles); args.push_back(&tab_groups); }
    for (int i = 0; i < p->d.n_pixel_args; i++) { pix[i] = pixel_ptrs[i]; args.push_back(&pix[i]); }
    const unsigned long long tiles = (unsigned long long)gx * gy;
    if (p->d.reduce_outputs) {
        rc = ensure_partials(p, tiles * gz);
        if (rc) return rc;
        args.push_back(&p->d_partials);
    } else {
        for (int k = 0; k < p->d.n_outputs; k++) { outs[k] = out_ptrs[k]; args.push_back(&outs[k]); }
    }
    time_begin(p, st);
    CU_CHECK(drv::LaunchKernel(p->f_main, gx, gy, gz, bs, 1, 1, 0, st, args.data(), nullptr));
    time_end(p, st);
    g_launches++;
    if (p->d.reduce_outputs) {
        rc = finish_group_partials(p, tiles, group_begin, (int)gz, out_ptrs[0], (cudaStream_t)st);
        if (rc) return rc;
    }
    return end_launch(p, st);
}

extern "C" int tegb_render_groups_launch(tegb_program *p, const void *const *group_ptrs, int group_begin,
                                         int group_count, const tegb_grid *g, const int32_t *windows,
                                         int max_rows, int max_cols, const void *const *pixel_ptrs,
                                         void *const *out_ptrs, void *stream) {
    return groups_launch_impl(p, group_ptrs, group_begin, group_count, g, windows, max_rows, max_cols, pixel_ptrs,
                              out_ptrs, stream, false);
}

// One launch over the IMAGE for groups with overlapping windows (accumulating image programs): block (bx, by) of the
// image walks the groups of its bin in order, so every pixel adds its contributions in ascending group order.
extern "C" int tegb_render_groups_gather(tegb_program *p, const void *const *group_ptrs, int group_count,
                                         const tegb_grid *g, const int32_t *windows, int max_rows, int max_cols,
                                         const int32_t *bin_begin, const int32_t *bin_groups, int clear,
                                         const void *const *pixel_ptrs, void *const *out_ptrs, void *stream) {
    if (!p || !g || !windows || !out_ptrs || !bin_begin || !bin_groups)
        return fail(TEGB_ERR_INVALID, "tegb_render_groups_gather: bad argument");
    if (!p->d.grouped || p->d.n_grid_args != 4) return fail(TEGB_ERR_INVALID, "not a grouped render program");
    if (!p->f_gather) return fail(TEGB_ERR_INVALID, "not an accumulating grouped image program (no <name>_gather kernel)");
    if (p->d.n_scalar_args > 0 && !group_ptrs) return fail(TEGB_ERR_INVALID, "group_ptrs is NULL");
    if (p->d.n_pixel_args > 0 && !pixel_ptrs) return fail(TEGB_ERR_INVALID, "pixel_ptrs is NULL");
    if (group_count < 0 || max_rows < 1 || max_cols < 1) return fail(TEGB_ERR_INVALID, "bad group count or window bounds");
    if (g->res_x < 1 || g->res_y < 1 || g->row_begin < 0 || g->row_end > g->res_x || g->row_begin > g->row_end)
        return fail(TEGB_ERR_INVALID, "bad grid");
    if (g->row_end == g->row_begin || (group_count == 0 && !clear)) return TEGB_OK;
    std::lock_guard<std::recursive_mutex> lock(p->mu);
    RT_CHECK(cudaSetDevice(p->device));
    CUstream st = (CUstream)stream;
    int rc = begin_launch(p, st);
    if (rc) return rc;
    if (group_count == 0) {                          // nothing to add: the cleared image
        const size_t bytes = (size_t)(g->row_end - g->row_begin) * g->res_y * p->d.elem_size;
        for (int k = 0; k < p->d.n_outputs; k++) RT_CHECK(cudaMemsetAsync(out_ptrs[k], 0, bytes, (cudaStream_t)st));
        return end_launch(p, st);
    }
    rc = groups_prologue(p, group_ptrs, 0, group_count, g, windows, max_rows, max_cols, st, true);
    if (rc) return rc;
    const GroupGeom q = group_geom(p, g, max_rows, max_cols, true);
    if (q.gy > 65535u) return fail(TEGB_ERR_INVALID, "too many tile rows in one launch");
    std::vector<const void *> pix(p->d.n_pixel_args > 0 ? p->d.n_pixel_args : 1);
    std::vector<void *> outs(p->d.n_outputs);
    std::vector<void *> args;
    int res_y = g->res_y, row0 = g->row_begin, row1 = g->row_end, tile_rows = q.tile_rows;
    int tab_groups = group_count, n_tx = q.n_tx, n_ty = q.n_ty;
    args.push_back(&p->d_xe);
    args.push_back(&p->d_ye);
    args.push_back(&res_y);
    args.push_back(&row0);
    args.push_back(&row1);
    args.push_back((void *)&windows);
    args.push_back((void *)&bin_begin);
    args.push_back((void *)&bin_groups);
    args.push_back(&clear);
    args.push_back(&p->d_utab);
    args.push_back(&tile_rows);
    if (p->d.n_tile_conds > 0) {
        args.push_back(&p->d_tiles);
        args.push_back(&tab_groups);
        args.push_back(&n_tx);
        args.push_back(&n_ty);
    }
    for (int i = 0; i < p->d.n_pixel_args; i++) { pix[i] = pixel_ptrs[i]; args.push_back(&pix[i]); }
    for (int k = 0; k < p->d.n_outputs; k++) { outs[k] = out_ptrs[k]; args.push_back(&outs[k]); }
    time_begin(p, st);
    CU_CHECK(drv::LaunchKernel(p->f_gather, q.gx, q.gy, 1, (unsigned)p->d.block_size, 1, 1, 0, st, args.data(), nullptr));
    time_end(p, st);
    g_launches++;
    return end_launch(p, st);
}

extern "C" int tegb_render_groups_prepare(tegb_program *p, const void *const *group_ptrs, int group_begin,
                                          int group_count, const tegb_grid *g, const int32_t *windows, int max_rows,
                                          int max_cols, void *stream) {
    if (!p || !g || !windows) return fail(TEGB_ERR_INVALID, "tegb_render_groups_prepare: bad argument");
    if (!p->d.grouped || p->d.n_grid_args != 4) return fail(TEGB_ERR_INVALID, "not a grouped render program");
    if (p->d.n_scalar_args > 0 && !group_ptrs) return fail(TEGB_ERR_INVALID, "group_ptrs is NULL");
    if (group_begin < 0 || group_count < 1 || group_count > 65535 || max_rows < 1 || max_cols < 1)
        return fail(TEGB_ERR_INVALID, "bad group range or window bounds");
    if (g->res_x < 1 || g->res_y < 1 || g->row_begin < 0 || g->row_end > g->res_x || g->row_begin >= g->row_end)
        return fail(TEGB_ERR_INVALID, "bad grid");
    std::lock_guard<std::recursive_mutex> lock(p->mu);
    RT_CHECK(cudaSetDevice(p->device));
    CUstream st = (CUstream)stream;
    int rc = begin_launch(p, st);
    if (rc) return rc;
    rc = groups_prologue(p, group_ptrs, group_begin, group_count, g, windows, max_rows, max_cols, st);
    if (rc) return rc;
    return end_launch(p, st);
}

extern "C" int tegb_render_groups_launch_prepared(tegb_program *p, const void *const *group_ptrs, int group_begin,
                                                  int group_count, const tegb_grid *g, const int32_t *windows,
                                                  int max_rows, int max_cols, const void *const *pixel_ptrs,
                                                  void *const *out_ptrs, void *stream) {
    return groups_launch_impl(p, group_ptrs, group_begin, group_count, g, windows, max_rows, max_cols, pixel_ptrs,
                              out_ptrs, stream, true);
}

extern "C" int tegb_render_launch_strided(tegb_program *p, const double *scalars, const tegb_grid *g, int tile_stride,
                                          int tile_phase, void *const *out_ptrs, void *stream);

extern "C" int tegb_render_launch(tegb_program *p, const double *scalars, const tegb_grid *g,
                                  void *const *out_ptrs, void *stream) {
    return tegb_render_launch_strided(p, scalars, g, 1, 0, out_ptrs, stream);
}

extern "C" int tegb_render_launch_strided(tegb_program *p, const double *scalars, const tegb_grid *g, int tile_stride,
                                          int tile_phase, void *const *out_ptrs, void *stream) {
    if (!p || !g || !out_ptrs) return fail(TEGB_ERR_INVALID, "tegb_render_launch: bad argument");
    if (tile_stride < 1 || tile_phase < 0 || tile_phase >= tile_stride)
        return fail(TEGB_ERR_INVALID, "tegb_render_launch_strided: need 0 <= tile_phase < tile_stride");
    if (p->d.n_grid_args != 4 || p->d.n_array_args != 0 || p->d.grouped)
        return fail(TEGB_ERR_INVALID, "not a plain render program (needs exactly the 4 pixel-box arguments on the grid)");
    if (g->res_x < 1 || g->res_y < 1 || g->row_begin < 0 || g->row_end > g->res_x || g->row_begin > g->row_end)
        return fail(TEGB_ERR_INVALID, "bad grid");
    std::lock_guard<std::recursive_mutex> lock(p->mu);
    RT_CHECK(cudaSetDevice(p->device));
    CUstream st = (CUstream)stream;
    int rc = begin_launch(p, st);
    if (rc) return rc;
    {
        // nothing to render for this rank (empty row range, or no tile row of its phase): image outputs stay untouched,
        // summed outputs are zero -- and still take part in the peer exchange, the other ranks wait for this one
        const int tr = p->d.tile_rows > 0 ? p->d.tile_rows : 1;
        const int n_tr0 = (g->row_end - g->row_begin + tr - 1) / tr;
        if (tile_phase >= n_tr0) {
            if (!p->d.reduce_outputs) return TEGB_OK;
            int rc0 = ensure_partials(p, 0);
            if (rc0) return rc0;
            rc0 = finish_partials(p, 0, out_ptrs[p->d.n_outputs - p->d.reduce_outputs], (cudaStream_t)st);
            if (rc0) return rc0;
            return end_launch(p, st);
        }
    }
    rc = ensure_edges(p, g, st);
    if (rc) return rc;
    rc = run_uniform(p, scalars, st);
    if (rc) return rc;
    std::vector<void *> outs(p->d.n_outputs);
    std::vector<void *> args;
    int res_y = g->res_y, row0 = g->row_begin;
    args.push_back(&p->d_xe);
    args.push_back(&p->d_ye);
    args.push_back(&res_y);
    args.push_back(&row0);
    const unsigned bs = (unsigned)p->d.block_size;
    const unsigned gx = (unsigned)((g->res_y + bs - 1) / bs);
    // one block per tile of block_size columns x tile_rows rows
    int tile_rows = p->d.tile_rows > 0 ? p->d.tile_rows : 1;
    int n_rows = g->row_end - g->row_begin;
    // tile rows tile_phase, tile_phase + tile_stride, ... of the row range are rendered by this launch
    const int n_tr = (n_rows + tile_rows - 1) / tile_rows;
    const unsigned gy = (unsigned)((n_tr - tile_phase + tile_stride - 1) / tile_stride);
    if (gy > 65535u) return fail(TEGB_ERR_INVALID, "more than 65535 tile rows in one render launch");
    if (gx > 65535u) return fail(TEGB_ERR_INVALID, "more than 65535 tile columns in one render launch");
    args.push_back(&n_rows);
    args.push_back(&tile_rows);
    args.push_back(&tile_stride);
    args.push_back(&tile_phase);
    // Tiled programs whose outputs are ALL summed, one partial per 32-column tile (the generated code makes the same
    // test): tiles without work get their zero partials from the classification kernel, and the main kernel is a fixed
    // number of blocks walking the (few) listed tile blocks next to the workers that serve the unresolved tiles.
    const bool sparse = p->d.n_tile_conds > 0 && p->d.reduce_outputs == p->d.n_outputs && p->d.tile_cols == 32;
    if (p->d.n_tile_conds > 0) {
        // classify the tiles for these launch-wide arguments and this band (tile_cols columns x tile_rows rows each;
        // a block of the main kernel spans block_size / tile_cols of them)
        int n_tx = (int)gx, n_ty = (int)gy;
        int n_txs = n_tx * (p->d.block_size / p->d.tile_cols);
        // two copies of the tri-states: by tile (diagnostics, tegb_program_tile_table) and in launch order (what the main
        // kernel reads: its order entry and its states are then independent loads)
        const size_t one_table = (size_t)p->d.n_tile_conds * n_txs * n_ty * p->d.elem_size;
        const size_t need = 2 * one_table;
        if (need > p->tiles_bytes) {
            RT_CHECK(cudaStreamSynchronize((cudaStream_t)st));      // an earlier launch may still read the old table
            if (p->d_tiles) { cudaFree(p->d_tiles); p->d_tiles = nullptr; p->tiles_bytes = 0; }
            RT_CHECK(cudaMalloc(&p->d_tiles, need));
            p->tiles_bytes = need;
        }
        void *d_tiles_ord = (char *)p->d_tiles + one_table;
        const int ns = p->d.n_scalar_args;
        std::vector<float> sf(ns > 0 ? ns : 1);
        std::vector<double> sd(ns > 0 ? ns : 1);
        std::vector<void *> targs;
        for (int i = 0; i < ns; i++) {
            if (p->d.elem_size == 4) { sf[i] = (float)scalars[i]; targs.push_back(&sf[i]); }
            else { sd[i] = scalars[i]; targs.push_back(&sd[i]); }
        }
        rc = ensure_ranges(p, g, n_txs, n_ty, tile_stride, tile_phase, (cudaStream_t)st);
        if (rc) return rc;
        const size_t n_tiles = (size_t)n_tx * n_ty;
        const size_t n_sub = (size_t)n_txs * n_ty;                   // tiles of tile_cols columns (what is classified)
        if (n_txs > 65535) return fail(TEGB_ERR_INVALID, "more than 65535 tiles per image row in one render launch");
        if (n_sub > p->order_count) {
            RT_CHECK(cudaStreamSynchronize((cudaStream_t)st));
            if (p->d_order) { cudaFree(p->d_order); p->d_order = nullptr; p->order_count = 0; }
            // launch order [blocks] | 4 counters | unresolved-tile list [tiles]
            RT_CHECK(cudaMalloc((void **)&p->d_order, (2 * n_sub + 4) * sizeof(int)));
            p->order_count = n_sub;
            p->counters_dirty = true;
        }
        int *d_counters = p->d_order + p->order_count;
        int *d_ulist = d_counters + 4;
        // counters: blocks filed from the front / from the back of the launch order, unresolved tiles, queue cursor.
        // They are zeroed BEHIND the main kernel for the next launch (off the critical path of this one: a 16-byte
        // memset node in front of the classification kernel cost ~2 us per call); only a first launch, or one after a
        // call that failed half-way, zeroes them here.
        if (p->counters_dirty) RT_CHECK(cudaMemsetAsync(d_counters, 0, 4 * sizeof(int), (cudaStream_t)st));
        p->counters_dirty = true;
        targs.push_back(&p->d_ranges);
        targs.push_back(&res_y);
        targs.push_back(&n_tx);
        targs.push_back(&n_ty);
        targs.push_back(&n_txs);
        targs.push_back(&p->d_tiles);
        targs.push_back(&d_tiles_ord);
        targs.push_back(&p->d_order);
        targs.push_back(&d_counters);
        targs.push_back(&d_ulist);
        if (sparse) {
            // (the classification kernel writes the zero partials of the tiles without work)
            rc = ensure_partials(p, (unsigned long long)n_sub);
            if (rc) return rc;
            targs.push_back(&p->d_partials);
        }
        const unsigned tb = 128, tg = (unsigned)(((size_t)n_txs * n_ty + tb - 1) / tb);     // a thread per tile
        CU_CHECK(drv::LaunchKernel(p->f_tiles, tg, 1, 1, tb, 1, 1, 0, st, targs.data(), nullptr));
        g_launches++;
        p->tiles_ord_ptr = d_tiles_ord;
        p->q_ulist = d_ulist;
        p->q_counters = d_counters;
        args.push_back(&p->tiles_ord_ptr);
        args.push_back(&p->d_order);
        args.push_back(&p->q_ulist);
        args.push_back(&p->q_counters);
        p->q_n_tx = n_tx;
        p->q_n_ty = n_ty;
        // image programs: dedicated blocks serve the queue of unresolved rows; programs with summed outputs walk their
        // unresolved tiles inside the owning block (their partial sums stay tied to the block)
        // (programs with summed outputs and 32-column tiles: a queue of whole tiles, one partial sum per tile; wider
        // tiles keep the walk inside the owning block)
        p->q_workers = (p->d.reduce_outputs && p->d.tile_cols != 32) ? 0
                       : queue_workers(p->device) * (p->d.reduce_outputs ? 2 : 1);
        args.push_back(&p->q_n_tx);
        args.push_back(&p->q_n_ty);
        args.push_back(&p->q_workers);
    }
    const int n_store = p->d.n_outputs - p->d.reduce_outputs;
    for (int k = 0; k < n_store; k++) { outs[k] = out_ptrs[k]; args.push_back(&outs[k]); }
    // partial sums: one per block, or one per tile for tiled programs with 32-column tiles
    const unsigned long long n_partials = (p->d.n_tile_conds > 0 && p->d.tile_cols == 32)
                                              ? (unsigned long long)gx * (bs / 32) * gy : (unsigned long long)gx * gy;
    if (p->d.reduce_outputs) {
        rc = ensure_partials(p, n_partials);
        if (rc) return rc;
        args.push_back(&p->d_partials);
    }
    time_begin(p, st);
    if (p->d.n_tile_conds > 0) {    // 1-D grid: the queue workers first, then one block per tile block (sparse: per listed one)
        unsigned ordinary = gx * gy;
        if (sparse) {
            const unsigned cap = (unsigned)queue_workers(p->device) * 2u;
            if (ordinary > cap) ordinary = cap;
        }
        CU_CHECK(drv::LaunchKernel(p->f_main, (unsigned)p->q_workers + ordinary, 1, 1, bs, 1, 1, 0, st, args.data(), nullptr));
    }
    else
        CU_CHECK(drv::LaunchKernel(p->f_main, gx, gy, 1, bs, 1, 1, 0, st, args.data(), nullptr));
    time_end(p, st);
    g_launches++;
    if (p->d.n_tile_conds > 0) {
        RT_CHECK(cudaMemsetAsync(p->q_counters, 0, 4 * sizeof(int), (cudaStream_t)st));
        p->counters_dirty = false;
    }
    if (p->d.reduce_outputs) {
        rc = finish_partials(p, n_partials, out_ptrs[n_store], (cudaStream_t)st);
        if (rc) return rc;
    }
    return end_launch(p, st);
}

extern "C" int tegb_program_tile_table(tegb_program *p, void *host_out, size_t nbytes, void *stream) {
    if (!p || !host_out) return fail(TEGB_ERR_INVALID, "tegb_program_tile_table: null argument");
    if (!p->d_tiles || nbytes > p->tiles_bytes / (p->d.grouped ? 1 : 2))
        return fail(TEGB_ERR_INVALID, "no tile table of that size (render first; n_tile_conds x tiles x elem_size bytes)");
    std::lock_guard<std::recursive_mutex> lock(p->mu);
    RT_CHECK(cudaSetDevice(p->device));
    RT_CHECK(cudaStreamSynchronize((cudaStream_t)stream));
    RT_CHECK(cudaMemcpy(host_out, p->d_tiles, nbytes, cudaMemcpyDeviceToHost));
    return TEGB_OK;
}

// Stage 1: block b sums elements [b*chunk, (b+1)*chunk) of every column: each thread walks its strided
// subsequence in order, then a fixed shuffle/shared-memory tree combines the 256 partials.  Stage 2: one
// block combines the per-block partials in the same fixed way.  The result depends only on (n, values).
template <typename T>
__global__ void __launch_bounds__(RED_THREADS) colsum_stage1(const T *const *cols, int n_cols, long long n,
                                                             long long chunk, T *partials) {
    __shared__ T smem[32];
    const long long lo = (long long)blockIdx.x * chunk;
    long long hi = lo + chunk;
    if (hi > n) hi = n;
    for (int c = 0; c < n_cols; c++) {
        const T *col = cols[c];
        T acc = 0;
        for (long long i = lo + threadIdx.x; i < hi; i += RED_THREADS) acc = acc + col[i];
        T s = block_sum_fixed(acc, smem);
        if (threadIdx.x == 0) partials[(long long)c * gridDim.x + blockIdx.x] = s;
    }
}

static int reduce_blocks(int64_t n) {
    long long b = (n + (long long)RED_THREADS * 8 - 1) / ((long long)RED_THREADS * 8);
    if (b < 1) b = 1;
    if (b > RED_MAX_BLOCKS) b = RED_MAX_BLOCKS;
    return (int)b;
}

extern "C" size_t tegb_reduce_workspace_bytes(int n_cols, int64_t n, int elem_size) {
    return (size_t)n_cols * sizeof(void *) + 256 + (size_t)n_cols * reduce_blocks(n) * (size_t)elem_size;
}

extern "C" int tegb_reduce_sum(const void *const *in_ptrs, int n_cols, int64_t n, int elem_size, void *out,
                               void *workspace, size_t workspace_bytes, void *stream) {
    if (!in_ptrs || n_cols < 1 || n < 0 || !out || !workspace) return fail(TEGB_ERR_INVALID, "tegb_reduce_sum: bad argument");
    if (elem_size != 4 && elem_size != 8) return fail(TEGB_ERR_INVALID, "elem_size must be 4 or 8");
    if (workspace_bytes < tegb_reduce_workspace_bytes(n_cols, n, elem_size))
        return fail(TEGB_ERR_INVALID, "workspace too small");
    cudaStream_t st = (cudaStream_t)stream;
    const int blocks = reduce_blocks(n);
    const long long chunk = (n + blocks - 1) / blocks;
    // column pointer table lives at the start of the workspace
    RT_CHECK(cudaMemcpyAsync(workspace, in_ptrs, (size_t)n_cols * sizeof(void *), cudaMemcpyHostToDevice, st));
    size_t off = ((size_t)n_cols * sizeof(void *) + 255) & ~(size_t)255;
    char *partials = (char *)workspace + off;
    if (elem_size == 4) {
        colsum_stage1<float><<<blocks, RED_THREADS, 0, st>>>((const float *const *)workspace, n_cols, n, chunk, (float *)partials);
        colsum_stage2<float><<<n_cols, RED_THREADS, 0, st>>>((const float *)partials, (long long)blocks, (float *)out);
    } else {
        colsum_stage1<double><<<blocks, RED_THREADS, 0, st>>>((const double *const *)workspace, n_cols, n, chunk, (double *)partials);
        colsum_stage2<double><<<n_cols, RED_THREADS, 0, st>>>((const double *)partials, (long long)blocks, (double *)out);
    }
    g_launches += 2;
    RT_CHECK(cudaGetLastError());
    return TEGB_OK;
}

// --------------------------------------------------------------------------------------- peer mailboxes
extern "C" int tegb_peer_create(int rank, int world, int max_values, int device, tegb_peer **out, void *handle64) {
    if (!out || !handle64 || world < 1 || world > TEGB_MAX_PEERS || rank < 0 || rank >= world || max_values < 1)
        return fail(TEGB_ERR_INVALID, "tegb_peer_create: bad argument (world <= %d)", TEGB_MAX_PEERS);
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
    static_assert(TEGB_MAX_PEERS <= 32, "one lane per peer");
    RT_CHECK(cudaSetDevice(device));
    tegb_peer *p = new tegb_peer();
    p->device = device;
    p->view.rank = rank; p->view.world = world; p->view.maxv = max_values;
    // mailbox slots, then this rank's sequence number and its block counter
    const size_t slots_bytes = sizeof(PeerSlot) * 2 * (size_t)world * max_values;
    p->box_bytes = slots_bytes + 64;
    if (cudaMalloc(&p->d_box, p->box_bytes) != cudaSuccess) { delete p; return fail(TEGB_ERR_CUDA, "cudaMalloc(mailbox) failed"); }
    cudaMemset(p->d_box, 0, p->box_bytes);
    cudaDeviceSynchronize();
    p->view.d_seq = (unsigned long long *)((char *)p->d_box + slots_bytes);
    p->view.d_done = (unsigned int *)((char *)p->d_box + slots_bytes + 16);
    if (cudaHostAlloc((void **)&p->h_err, sizeof(int), cudaHostAllocMapped) != cudaSuccess) {
        cudaFree(p->d_box); delete p;
        return fail(TEGB_ERR_CUDA, "cudaHostAlloc(peer error flag) failed");
    }
    *p->h_err = 0;
    void *d_err = nullptr;
    if (cudaHostGetDevicePointer(&d_err, p->h_err, 0) != cudaSuccess) {
        cudaFreeHost(p->h_err); cudaFree(p->d_box); delete p;
        return fail(TEGB_ERR_CUDA, "cudaHostGetDevicePointer(peer error flag) failed");
    }
    p->view.err = (volatile int *)d_err;
    // how long an exchange waits for a peer: rank skew of seconds is ordinary (a JIT compile on one rank, a host stall),
    // so the default is generous; 0 waits for ever
    double secs = 120.0;
    if (const char *e = getenv("TEGB200_PEER_TIMEOUT_S")) secs = atof(e);
    p->view.timeout_ns = secs > 0 ? (unsigned long long)(secs * 1e9) : 0ull;
    cudaIpcMemHandle_t h;
    if (cudaIpcGetMemHandle(&h, p->d_box) != cudaSuccess) {
        cudaFree(p->d_box); delete p;
        return fail(TEGB_ERR_CUDA, "cudaIpcGetMemHandle failed: %s", cudaGetErrorString(cudaGetLastError()));
    }
    memcpy(handle64, &h, 64);
    *out = p;
    return TEGB_OK;
}

/* handles: world x 64 bytes, the handle of rank r at offset 64 r (gathered by the caller, e.g. torch.distributed) */
extern "C" int tegb_peer_connect(tegb_peer *p, const void *handles) {
    if (!p || !handles) return fail(TEGB_ERR_INVALID, "tegb_peer_connect: null argument");
    RT_CHECK(cudaSetDevice(p->device));
    for (int r = 0; r < p->view.world; r++) {
        if (r == p->view.rank) { p->view.box[r] = (PeerSlot *)p->d_box; continue; }
        cudaIpcMemHandle_t h;
        memcpy(&h, (const char *)handles + 64 * r, 64);
        void *ptr = nullptr;
        cudaError_t e = cudaIpcOpenMemHandle(&ptr, h, cudaIpcMemLazyEnablePeerAccess);
        if (e != cudaSuccess) return fail(TEGB_ERR_CUDA, "cudaIpcOpenMemHandle(rank %d): %s", r, cudaGetErrorString(e));
        p->opened.push_back(ptr);
        p->view.box[r] = (PeerSlot *)ptr;
    }
    return TEGB_OK;
}

extern "C" int tegb_peer_allreduce_sum(tegb_peer *p, void *buf, int count, int elem_size, void *stream) {
    if (!p || !buf || count < 0 || count > p->view.maxv || (elem_size != 4 && elem_size != 8))
        return fail(TEGB_ERR_INVALID, "tegb_peer_allreduce_sum: bad argument");
    if (count == 0) return TEGB_OK;
    if (peer_failed(p)) return TEGB_ERR_PEER;
    RT_CHECK(cudaSetDevice(p->device));
    const int tb = 32, tg = count;
    if (elem_size == 4) peer_allreduce_kernel<float><<<tg, tb, 0, (cudaStream_t)stream>>>((float *)buf, count, p->view);
    else peer_allreduce_kernel<double><<<tg, tb, 0, (cudaStream_t)stream>>>((double *)buf, count, p->view);
    RT_CHECK(cudaGetLastError());
    g_launches++;
    return TEGB_OK;
}

/* 1 when some exchange gave up waiting for a peer (its sums are NaN, later launches fail); synchronises the device */
extern "C" int tegb_peer_error(tegb_peer *p) {
    if (!p) return 1;
    cudaSetDevice(p->device);
    cudaDeviceSynchronize();
    return *(volatile int *)p->h_err;
}

/* all ranks must have finished using the mailboxes (barrier) before any rank destroys its own */
extern "C" void tegb_peer_destroy(tegb_peer *p) {
    if (!p) return;
    cudaSetDevice(p->device);
    cudaDeviceSynchronize();
    for (void *q : p->opened) cudaIpcCloseMemHandle(q);
    if (p->d_box) cudaFree(p->d_box);
    if (p->h_err) cudaFreeHost(p->h_err);
    delete p;
}

/* reduce_outputs launches of `prog` deliver the sum over all ranks of `peer` (NULL restores local sums) */
extern "C" int tegb_program_set_peer(tegb_program *prog, tegb_peer *peer) {
    if (!prog) return fail(TEGB_ERR_INVALID, "tegb_program_set_peer: null program");
    std::lock_guard<std::recursive_mutex> lock(prog->mu);
    prog->peer = peer;
    prog->n_groups_total = 0;
    prog->group_map = nullptr;
    return TEGB_OK;
}

/* grouped reduce_outputs programs: tegb_render_groups_launch then delivers the all-reduced [n_groups_total][n_outputs]
 * sums to out_ptrs[0] (ALL rows, whatever group range the launch covered; groups it did not cover contribute zero).
 * local_of (device int32 [n_groups_total], or NULL for the identity): this rank's group index of every GLOBAL group, -1
 * for groups it does not hold -- a rank may keep only the groups whose windows reach into its rows.
 * Every rank must make the same sequence of such launches, also ranks with nothing to render. */
extern "C" int tegb_program_set_peer_groups(tegb_program *prog, tegb_peer *peer, int n_groups_total, const int32_t *local_of) {
    if (!prog || (peer && n_groups_total < 1)) return fail(TEGB_ERR_INVALID, "tegb_program_set_peer_groups: bad argument");
    if (!prog->d.grouped || !prog->d.reduce_outputs) return fail(TEGB_ERR_INVALID, "not a grouped reduce program");
    if (peer && (long long)n_groups_total * prog->d.n_outputs > peer->view.maxv)
        return fail(TEGB_ERR_INVALID, "peer mailbox smaller than n_groups_total x n_outputs");
    std::lock_guard<std::recursive_mutex> lock(prog->mu);
    prog->peer = peer;
    prog->n_groups_total = peer ? n_groups_total : 0;
    prog->group_map = peer ? (const int *)local_of : nullptr;
    return TEGB_OK;
}

// ------------------------------------------------------------------------------------------------ NCCL
namespace nccl {
typedef struct { char internal[128]; } UniqueId;
typedef int (*GetUniqueId_t)(UniqueId *);
typedef int (*CommInitRank_t)(void **, int, UniqueId, int);
typedef int (*CommDestroy_t)(void *);
typedef int (*AllReduce_t)(const void *, void *, size_t, int, int, void *, cudaStream_t);
typedef const char *(*GetErrorString_t)(int);
static void *lib = nullptr;
static GetUniqueId_t GetUniqueId;
static CommInitRank_t CommInitRank;
static CommDestroy_t CommDestroy;
static AllReduce_t AllReduce;
static GetErrorString_t GetErrorString;
static std::once_flag once;
static bool ok = false;
static std::string why;
static void load() {
    // RTLD_NOLOAD first: reuse the libnccl the process (e.g. torch) already mapped, if any
    lib = dlopen("libnccl.so.2", RTLD_NOW | RTLD_NOLOAD);
    if (!lib) lib = dlopen("libnccl.so.2", RTLD_NOW);
    if (!lib) { why = std::string("cannot load libnccl.so.2: ") + dlerror(); return; }
#define SYM(var, name)                                                   \
    var = (decltype(var))dlsym(lib, name);                               \
    if (!var) { why = std::string("libnccl lacks ") + name; return; }
    SYM(GetUniqueId, "ncclGetUniqueId");
    SYM(CommInitRank, "ncclCommInitRank");
    SYM(CommDestroy, "ncclCommDestroy");
    SYM(AllReduce, "ncclAllReduce");
    SYM(GetErrorString, "ncclGetErrorString");
#undef SYM
    ok = true;
}
static bool ready() { std::call_once(once, load); return ok; }
}  // namespace nccl

#define NCCL_CHECK(call)                                                                         \
    do {                                                                                         \
        int _r = (call);                                                                         \
        if (_r != 0) return fail(TEGB_ERR_NCCL, "%s failed: %s", #call, nccl::GetErrorString(_r)); \
    } while (0)

extern "C" int tegb_nccl_unique_id(void *id128) {
    if (!id128) return fail(TEGB_ERR_INVALID, "null id");
    if (!nccl::ready()) return fail(TEGB_ERR_NCCL, "%s", nccl::why.c_str());
    NCCL_CHECK(nccl::GetUniqueId((nccl::UniqueId *)id128));
    return TEGB_OK;
}

extern "C" int tegb_nccl_comm_create(const void *id128, int world_size, int rank, int device, void **comm) {
    if (!id128 || !comm) return fail(TEGB_ERR_INVALID, "null argument");
    if (!nccl::ready()) return fail(TEGB_ERR_NCCL, "%s", nccl::why.c_str());
    RT_CHECK(cudaSetDevice(device));
    nccl::UniqueId id;
    memcpy(&id, id128, sizeof id);
    NCCL_CHECK(nccl::CommInitRank(comm, world_size, id, rank));
    return TEGB_OK;
}

extern "C" int tegb_nccl_comm_destroy(void *comm) {
    if (!comm) return TEGB_OK;
    if (!nccl::ready()) return fail(TEGB_ERR_NCCL, "%s", nccl::why.c_str());
    NCCL_CHECK(nccl::CommDestroy(comm));
    return TEGB_OK;
}

extern "C" int tegb_allreduce_sum(void *buf, int64_t count, int elem_size, void *comm, void *stream) {
    if (!buf || count < 0 || !comm) return fail(TEGB_ERR_INVALID, "tegb_allreduce_sum: bad argument");
    if (elem_size != 4 && elem_size != 8) return fail(TEGB_ERR_INVALID, "elem_size must be 4 or 8");
    if (!nccl::ready()) return fail(TEGB_ERR_NCCL, "%s", nccl::why.c_str());
    const int dtype = elem_size == 4 ? 7 /* ncclFloat32 */ : 8 /* ncclFloat64 */;
    NCCL_CHECK(nccl::AllReduce(buf, buf, (size_t)count, dtype, 0 /* ncclSum */, comm, (cudaStream_t)stream));
    return TEGB_OK;
}

// ------------------------------------------------------------------------------------ FMA peak (roofline)
// 8 independent dependent-FMA chains per thread; enough warps to saturate the 4 SMSP issue ports.
template <typename T>
__global__ void __launch_bounds__(256) fma_peak_kernel(T *sink, int iters, T a, T b) {
    T x0 = a + (T)threadIdx.x, x1 = x0 + (T)1, x2 = x0 + (T)2, x3 = x0 + (T)3;
    T x4 = x0 + (T)4, x5 = x0 + (T)5, x6 = x0 + (T)6, x7 = x0 + (T)7;
    for (int i = 0; i < iters; i++) {
#pragma unroll
        for (int u = 0; u < 16; u++) {
            x0 = fma(x0, a, b); x1 = fma(x1, a, b); x2 = fma(x2, a, b); x3 = fma(x3, a, b);
            x4 = fma(x4, a, b); x5 = fma(x5, a, b); x6 = fma(x6, a, b); x7 = fma(x7, a, b);
        }
    }
    T s = ((x0 + x1) + (x2 + x3)) + ((x4 + x5) + (x6 + x7));
    if (s == (T)123456789) sink[0] = s;     // never true; keeps the chains alive
}

extern "C" int tegb_fma_peak(int elem_size, int device, double *tflops, double *sm_clock_mhz) {
    if (!tflops) return fail(TEGB_ERR_INVALID, "null argument");
    if (elem_size != 4 && elem_size != 8) return fail(TEGB_ERR_INVALID, "elem_size must be 4 or 8");
    if (!drv::ready()) return fail(TEGB_ERR_NO_DEVICE, "CUDA driver unavailable: %s", drv::why.c_str());
    RT_CHECK(cudaSetDevice(device));
    cudaDeviceProp prop;
    RT_CHECK(cudaGetDeviceProperties(&prop, device));
    void *sink = nullptr;
    RT_CHECK(cudaMalloc(&sink, 64));
    const int blocks = prop.multiProcessorCount * 8, threads = 256;
    const int iters = elem_size == 4 ? 4096 : 1024;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    float best = 1e30f;
    for (int rep = 0; rep < 5; rep++) {
        cudaEventRecord(e0);
        if (elem_size == 4) fma_peak_kernel<float><<<blocks, threads>>>((float *)sink, iters, 0.999f, 0.001f);
        else fma_peak_kernel<double><<<blocks, threads>>>((double *)sink, iters, 0.999, 0.001);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        g_launches++;
        float ms = 0;
        cudaEventElapsedTime(&ms, e0, e1);
        if (rep > 0 && ms < best) best = ms;
    }
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    cudaFree(sink);
    RT_CHECK(cudaGetLastError());
    const double fmas = (double)blocks * threads * (double)iters * 16.0 * 8.0;
    *tflops = 2.0 * fmas / (best * 1e-3) / 1e12;
    if (sm_clock_mhz) {
        int khz = 0;
        cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, device);
        *sm_clock_mhz = khz / 1000.0;
    }
    return TEGB_OK;
}
