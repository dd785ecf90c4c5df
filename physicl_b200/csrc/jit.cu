// Run-time compiled kernels (NVRTC) behind the same C ABI.
//
// The reference builds its OpenCL kernels at run time from Python strings (CLProgram.build_kernel,
// physicl/__init__.py:583-597: cl.Program(ctx, src).build()), and two of its features only exist as
// run-time text: the user's number-density expression of ScatterIsotropicStep(variable_n=True)
// (light.py:295-299) and user-written CLProgram kernels (README.md:8).  This file is the equivalent of
// cl.Program(...).build() + prog.<kernel>(queue, global, local, *args) for sm_100a: source in, cubin
// out (nvrtcCompileProgram, --gpu-architecture=sm_100a, -fmad=false), launched through the driver API.
// libnvrtc and libcuda are opened with dlopen on first use, so the library itself has no link-time
// dependency on either.
#include <cuda.h>
#include <dlfcn.h>
#include <nvrtc.h>
#include <stdlib.h>

#include "pcl_common.cuh"

#include "pcl_photon_body.cuh"

struct pcl_jit_kernel {
    CUfunction fn;
    pcl_jit_kernel *next;
};
struct pcl_jit_module {
    CUmodule mod;
    pcl_jit_kernel *kernels;  // handed out by pcl_jit_get_kernel, owned by the module
};

namespace {
struct Api {
    void *nvrtc = nullptr, *cuda = nullptr;
    nvrtcResult (*CreateProgram)(nvrtcProgram *, const char *, const char *, int, const char *const *, const char *const *);
    nvrtcResult (*CompileProgram)(nvrtcProgram, int, const char *const *);
    nvrtcResult (*GetProgramLogSize)(nvrtcProgram, size_t *);
    nvrtcResult (*GetProgramLog)(nvrtcProgram, char *);
    nvrtcResult (*GetCUBINSize)(nvrtcProgram, size_t *);
    nvrtcResult (*GetCUBIN)(nvrtcProgram, char *);
    nvrtcResult (*DestroyProgram)(nvrtcProgram *);
    const char *(*GetErrorString)(nvrtcResult);
    CUresult (*ModuleLoadData)(CUmodule *, const void *);
    CUresult (*ModuleGetFunction)(CUfunction *, CUmodule, const char *);
    CUresult (*ModuleUnload)(CUmodule);
    CUresult (*LaunchKernel)(CUfunction, unsigned, unsigned, unsigned, unsigned, unsigned, unsigned, unsigned, CUstream, void **, void **);
    CUresult (*GetErrorStringCu)(CUresult, const char **);
    bool ok = false;
} g_api;

template <typename F>
bool sym(void *lib, const char *name, F &out) {
    out = (F)dlsym(lib, name);
    return out != nullptr;
}

int load_api(pcl_ctx *ctx) {
    if (g_api.ok) return 0;
    const char *nv[] = {"/usr/local/cuda/lib64/libnvrtc.so.12", "libnvrtc.so.12", "libnvrtc.so"};
    for (const char *n : nv)
        if (!g_api.nvrtc) g_api.nvrtc = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
    PCL_REQUIRE(ctx, g_api.nvrtc != nullptr, "libnvrtc.so.12 not found (run-time kernels need NVRTC)");
    g_api.cuda = dlopen("libcuda.so.1", RTLD_NOW | RTLD_GLOBAL);
    PCL_REQUIRE(ctx, g_api.cuda != nullptr, "libcuda.so.1 not found");
    bool ok = sym(g_api.nvrtc, "nvrtcCreateProgram", g_api.CreateProgram) && sym(g_api.nvrtc, "nvrtcCompileProgram", g_api.CompileProgram) &&
              sym(g_api.nvrtc, "nvrtcGetProgramLogSize", g_api.GetProgramLogSize) && sym(g_api.nvrtc, "nvrtcGetProgramLog", g_api.GetProgramLog) &&
              sym(g_api.nvrtc, "nvrtcGetCUBINSize", g_api.GetCUBINSize) && sym(g_api.nvrtc, "nvrtcGetCUBIN", g_api.GetCUBIN) &&
              sym(g_api.nvrtc, "nvrtcDestroyProgram", g_api.DestroyProgram) && sym(g_api.nvrtc, "nvrtcGetErrorString", g_api.GetErrorString) &&
              sym(g_api.cuda, "cuModuleLoadData", g_api.ModuleLoadData) && sym(g_api.cuda, "cuModuleGetFunction", g_api.ModuleGetFunction) &&
              sym(g_api.cuda, "cuModuleUnload", g_api.ModuleUnload) && sym(g_api.cuda, "cuLaunchKernel", g_api.LaunchKernel) &&
              sym(g_api.cuda, "cuGetErrorString", g_api.GetErrorStringCu);
    PCL_REQUIRE(ctx, ok, "NVRTC / driver entry points missing");
    g_api.ok = true;
    return 0;
}
}  // namespace

static const char *const kJitOpts[] = {"--gpu-architecture=sm_100a", "-fmad=false", "-default-device", "--std=c++17",
                                       "-lineinfo"};

// Compiles `source` to an sm_100a cubin.  On failure the NVRTC log goes to `log`.  No device needed.
static int jit_compile_cubin(const char *source, int n_headers, const char *const *header_names,
                             const char *const *header_texts, char **cubin, size_t *cubin_bytes, char *log, size_t log_cap) {
    nvrtcProgram prog;
    nvrtcResult r = g_api.CreateProgram(&prog, source, "physicl_b200_jit.cu", n_headers, header_texts, header_names);
    if (r != NVRTC_SUCCESS) {
        if (log && log_cap) snprintf(log, log_cap, "nvrtcCreateProgram: %s", g_api.GetErrorString(r));
        return -20;
    }
    r = g_api.CompileProgram(prog, (int)(sizeof(kJitOpts) / sizeof(kJitOpts[0])), kJitOpts);
    if (r != NVRTC_SUCCESS) {
        size_t n = 0;
        g_api.GetProgramLogSize(prog, &n);
        char *full = (char *)malloc(n + 1);
        if (full) {
            g_api.GetProgramLog(prog, full);
            full[n] = 0;
            if (log && log_cap) snprintf(log, log_cap, "kernel build failed:\n%s", full);
            free(full);
        }
        g_api.DestroyProgram(&prog);
        return -21;
    }
    size_t nb = 0;
    g_api.GetCUBINSize(prog, &nb);
    char *bin = (char *)malloc(nb ? nb : 1);
    if (bin) g_api.GetCUBIN(prog, bin);
    g_api.DestroyProgram(&prog);
    if (!bin) return -24;
    *cubin = bin;
    *cubin_bytes = nb;
    return 0;
}

// Compile only (no device, no context): lets a step report a bad user expression when it is built
// and lets the CPU test-suite cover the generated kernels.  Returns 0 and the cubin size on success.
extern "C" int pcl_jit_check(const char *source, int n_headers, const char *const *header_names,
                             const char *const *header_texts, char *log, uint64_t log_cap, uint64_t *cubin_bytes) {
    if (log && log_cap) log[0] = 0;
    if (!source) return -1;
    if (!g_api.nvrtc) {
        const char *nv[] = {"/usr/local/cuda/lib64/libnvrtc.so.12", "libnvrtc.so.12", "libnvrtc.so"};
        for (const char *n : nv)
            if (!g_api.nvrtc) g_api.nvrtc = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
    }
    bool ok = g_api.nvrtc && sym(g_api.nvrtc, "nvrtcCreateProgram", g_api.CreateProgram) &&
              sym(g_api.nvrtc, "nvrtcCompileProgram", g_api.CompileProgram) &&
              sym(g_api.nvrtc, "nvrtcGetProgramLogSize", g_api.GetProgramLogSize) &&
              sym(g_api.nvrtc, "nvrtcGetProgramLog", g_api.GetProgramLog) && sym(g_api.nvrtc, "nvrtcGetCUBINSize", g_api.GetCUBINSize) &&
              sym(g_api.nvrtc, "nvrtcGetCUBIN", g_api.GetCUBIN) && sym(g_api.nvrtc, "nvrtcDestroyProgram", g_api.DestroyProgram) &&
              sym(g_api.nvrtc, "nvrtcGetErrorString", g_api.GetErrorString);
    if (!ok) {
        if (log && log_cap) snprintf(log, log_cap, "libnvrtc.so.12 not found (run-time kernels need NVRTC)");
        return -1;
    }
    char *bin = nullptr;
    size_t nb = 0;
    int rc = jit_compile_cubin(source, n_headers, header_names, header_texts, &bin, &nb, log, (size_t)log_cap);
    if (rc) return rc;
    free(bin);
    if (cubin_bytes) *cubin_bytes = nb;
    return 0;
}

// cl.Program(ctx, src).build() (physicl/__init__.py:597): CUDA C++ text defining one or more
// `extern "C" __global__` kernels; headers: n_headers in-memory includes (name, text).  On a compile
// error the NVRTC log is the error text (pcl_last_error), as with pyopencl's build().
extern "C" int pcl_jit_build(pcl_ctx *ctx, const char *source, int n_headers, const char *const *header_names,
                             const char *const *header_texts, pcl_jit_module **out) {
    PCL_ENTER(ctx);
    PCL_REQUIRE(ctx, source && out, "null argument");
    *out = nullptr;
    int rc = load_api(ctx);
    if (rc) return rc;
    PCL_CUDA(ctx, cudaFree(0));  // make sure the primary context exists and is current
    char *bin = nullptr;
    size_t nb = 0;
    char log[480];
    rc = jit_compile_cubin(source, n_headers, header_names, header_texts, &bin, &nb, log, sizeof(log));
    if (rc) {
        pcl_set_error(ctx, "%s", log);
        return rc;
    }
    pcl_jit_module *m = (pcl_jit_module *)calloc(1, sizeof(pcl_jit_module));
    CUresult cr = m ? g_api.ModuleLoadData(&m->mod, bin) : CUDA_ERROR_OUT_OF_MEMORY;
    free(bin);
    if (cr != CUDA_SUCCESS) {
        const char *msg = "?";
        g_api.GetErrorStringCu(cr, &msg);
        pcl_set_error(ctx, "loading run-time compiled module: %s", msg);
        free(m);
        return -22;
    }
    *out = m;
    return 0;
}

// prog.<kernel> (physicl/__init__.py:656): look a kernel up by name.  The handle belongs to the module.
extern "C" int pcl_jit_get_kernel(pcl_ctx *ctx, pcl_jit_module *m, const char *kernel_name, pcl_jit_kernel **out) {
    PCL_ENTER(ctx);
    PCL_REQUIRE(ctx, m && kernel_name && out, "null argument");
    *out = nullptr;
    CUfunction fn;
    CUresult cr = g_api.ModuleGetFunction(&fn, m->mod, kernel_name);
    if (cr != CUDA_SUCCESS) {
        const char *msg = "?";
        g_api.GetErrorStringCu(cr, &msg);
        pcl_set_error(ctx, "kernel '%s': %s", kernel_name, msg);
        return -22;
    }
    pcl_jit_kernel *k = (pcl_jit_kernel *)calloc(1, sizeof(pcl_jit_kernel));
    PCL_REQUIRE(ctx, k != nullptr, "out of memory");
    k->fn = fn;
    k->next = m->kernels;
    m->kernels = k;
    *out = k;
    return 0;
}

static int jit_launch(pcl_ctx *ctx, cudaStream_t st, pcl_jit_kernel *k, unsigned grid, unsigned block, void **args) {
    CUresult cr = g_api.LaunchKernel(k->fn, grid, 1, 1, block, 1, 1, 0, (CUstream)st, args, nullptr);
    if (cr != CUDA_SUCCESS) {
        const char *msg = "?";
        g_api.GetErrorStringCu(cr, &msg);
        pcl_set_error(ctx, "cuLaunchKernel: %s", msg);
        return -23;
    }
    ctx->launches++;
    return 0;
}

// One-dimensional launch over n work items (the reference's global=(N,), local=None,
// physicl/__init__.py:640-656): 256 threads per CTA, a grid that is a multiple of the SM count;
// generated kernels walk their items in grid-stride order.  args: pointers to the argument values.
extern "C" int pcl_jit_launch(pcl_ctx *ctx, uintptr_t stream, pcl_jit_kernel *k, uint64_t n, void **args) {
    PCL_ENTER(ctx);
    PCL_REQUIRE(ctx, k && k->fn && args, "null argument");
    if (n == 0) return 0;
    return jit_launch(ctx, (cudaStream_t)stream, k, pcl_stream_grid(ctx, n, 256, 8), 256, args);
}

extern "C" int pcl_jit_free(pcl_ctx *ctx, pcl_jit_module *m) {
    if (!m) return 0;
    if (ctx) cudaSetDevice(ctx->device);
    if (g_api.ok && m->mod) g_api.ModuleUnload(m->mod);
    for (pcl_jit_kernel *k = m->kernels; k;) {
        pcl_jit_kernel *nx = k->next;
        free(k);
        k = nx;
    }
    free(m);
    return 0;
}

// ---------------------------------------------------------------------------------------------
// Photon steps whose kernel was compiled at run time (pcl_jit_photon.cuh): same arguments as
// pcl_photon_steps / pcl_scatter plus the float64 constants of the variable-density law.
// ---------------------------------------------------------------------------------------------
static int fill_varn(pcl_ctx *ctx, StepK &K, const pcl_varn_params *vn) {
    PCL_REQUIRE(ctx, vn != nullptr, "variable-density parameters are required");
    K.kd = vn->kd;
    K.e0 = vn->e0;
    K.a_slot = vn->a_slot;
    K.n_slot = vn->n_slot;
    return 0;
}

extern "C" int pcl_photon_steps_jit(pcl_ctx *ctx, uintptr_t stream, pcl_jit_kernel *k, const pcl_soa *p, float dt,
                                    const pcl_scatter_params *sp, const pcl_varn_params *vn, const pcl_rng *rng,
                                    float escape_r2, const pcl_planes *planes, int64_t *tally_table, uint32_t nsteps) {
    PCL_ENTER(ctx);
    PCL_REQUIRE(ctx, k && k->fn && p && sp && rng && tally_table, "null argument");
    PCL_REQUIRE(ctx, p->x && p->y && p->z && p->vx && p->vy && p->vz, "r and v planes are required");
    PCL_REQUIRE(ctx, p->n < (1ull << 32), "a shard holds fewer than 2^32 slots");
    PCL_REQUIRE(ctx, (p->id_base & 0xffffffffull) + p->n <= (1ull << 32), "the global ids of a shard must not cross a multiple of 2^32");
    if (sp->mode & PCL_SCATTER_WAVELENGTH) PCL_REQUIRE(ctx, p->e != nullptr, "wavelength law needs the e plane");
    PCL_REQUIRE(ctx, !(sp->mode & PCL_SCATTER_SFU), "PCL_SCATTER_SFU applies to the pre-compiled fused photon steps only");
    PCL_REQUIRE(ctx, nsteps == 1 || rng->u_rand == nullptr, "multi-step runs draw from Philox; injected uniforms are per step");
    cudaStream_t st = (cudaStream_t)stream;
    PCL_CUDA(ctx, cudaMemsetAsync(tally_table, 0, (size_t)nsteps * PCL_TALLY_COLS * sizeof(int64_t), st));
    if (p->n == 0) return 0;
    StepK K;
    int rc = pcl_fill_stepk(ctx, K, dt, sp, rng, escape_r2, planes, p->id_base);
    if (rc == 0) rc = fill_varn(ctx, K, vn);
    if (rc) return rc;
    int aligned = pcl_aligned16(p->x) && pcl_aligned16(p->y) && pcl_aligned16(p->z) && pcl_aligned16(p->vx) &&
                  pcl_aligned16(p->vy) && pcl_aligned16(p->vz) && pcl_aligned16(p->e) && pcl_aligned16(p->id) &&
                  pcl_aligned16(p->nscat) && pcl_aligned16(K.u_theta) && pcl_aligned16(K.u_phi) && pcl_aligned16(K.u_rand);
    pcl_soa view = *p;
    const unsigned grid = pcl_stream_grid(ctx, aligned ? (p->n + 3) / 4 : p->n, 256, 8);
    for (uint32_t s = 0; s < nsteps; ++s) {
        K.step = rng->step + s;
        int64_t *row = tally_table + (size_t)s * PCL_TALLY_COLS;
        void *args[] = {&view, &K, &row, &aligned};
        rc = jit_launch(ctx, st, k, grid, 256, args);
        if (rc) return rc;
    }
    return 0;
}

extern "C" int pcl_scatter_jit(pcl_ctx *ctx, uintptr_t stream, pcl_jit_kernel *k, const pcl_soa *p,
                               const pcl_scatter_params *sp, const pcl_varn_params *vn, const pcl_rng *rng, int32_t *flags,
                               int64_t *tally_row) {
    PCL_ENTER(ctx);
    PCL_REQUIRE(ctx, k && k->fn && p && sp, "null argument");
    PCL_REQUIRE(ctx, p->x && p->y && p->z && p->vx && p->vy && p->vz, "r and v planes are required");
    PCL_REQUIRE(ctx, p->dx && p->dy && p->dz, "stand-alone scatter reads the dr planes");
    PCL_REQUIRE(ctx, p->n_dev == nullptr, "this step needs the exact slot count on the host (n_dev must be null)");
    if (sp->mode & PCL_SCATTER_WAVELENGTH) PCL_REQUIRE(ctx, p->e != nullptr, "wavelength law needs the e plane");
    PCL_REQUIRE(ctx, !(sp->mode & PCL_SCATTER_SFU), "PCL_SCATTER_SFU applies to the pre-compiled fused photon steps only");
    if (p->n == 0) return 0;
    StepK K;
    int rc = pcl_fill_stepk(ctx, K, 0.f, sp, rng, 0.f, nullptr, p->id_base);
    if (rc == 0) rc = fill_varn(ctx, K, vn);
    if (rc) return rc;
    pcl_soa view = *p;
    uint64_t n = p->n;
    void *args[] = {&view, &K, &flags, &tally_row, &n};
    return jit_launch(ctx, (cudaStream_t)stream, k, pcl_stream_grid(ctx, n, 256, 8), 256, args);
}
