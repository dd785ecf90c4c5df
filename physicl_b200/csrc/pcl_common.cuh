// Shared device/host helpers for the physicl_b200 kernels (sm_100a only).
// Arithmetic contract: this library is compiled with -fmad=false, so `a*b+c` is never contracted;
// every fused multiply-add is an explicit fmaf().  The CPU oracle twin (oracle/c/oracle.c, orc_*_f32) uses
// the same sequence of IEEE-754 binary32 operations (mul, add, fma, floor, int<->float conversions) and the
// same direction table, which is what makes the integer tallies bit-exact between the two.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include "../../include/physicl_b200.h"

#ifndef PCL_BLOCK
#define PCL_BLOCK 256
#endif
// minimum resident CTAs per SM requested from ptxas for the fused photon kernels (register cap)
#ifndef PCL_PHOTON_MINB
#define PCL_PHOTON_MINB 3
#endif
#define PCL_WARPS (PCL_BLOCK / 32)
// most timesteps one photon launch advances in registers (pcl_k_photon_multi)
#define PCL_FUSE_MAX 8

struct pcl_hostpipe;  // hostpipe.cu

struct pcl_ctx {
    int device;
    int sm_count;
    size_t l2_bytes;
    size_t hbm_bytes;
    char name[128];
    char err[512];
    uint64_t launches;
    // compaction scratch: per-block live counts / offsets
    uint32_t *scan_buf;
    size_t scan_cap;
    // gravity: partial accelerations of the j splits
    float *grav_part;
    size_t grav_cap;
    // gravity, TMA-fed form: the j-bodies re-packed pair by pair, padded to whole tiles
    void *grav_pairs;
    size_t grav_pairs_cap;
    pcl_hostpipe *pipe;
    // (sin, cos)(2 pi k / 512): the direction table of the photon kernels (pcl_device.cuh), built at pcl_init
    float *trig;
};

void pcl_set_error(pcl_ctx *ctx, const char *fmt, ...);

#define PCL_CUDA(ctx, call)                                                                  \
    do {                                                                                     \
        cudaError_t e__ = (call);                                                            \
        if (e__ != cudaSuccess) {                                                            \
            pcl_set_error(ctx, "%s:%d: %s -> %s", __FILE__, __LINE__, #call,                 \
                          cudaGetErrorString(e__));                                          \
            return -(int)e__ - 1000;                                                         \
        }                                                                                    \
    } while (0)

#define PCL_REQUIRE(ctx, cond, msg)                                                          \
    do {                                                                                     \
        if (!(cond)) {                                                                       \
            pcl_set_error(ctx, "%s:%d: %s (%s)", __FILE__, __LINE__, msg, #cond);            \
            return -1;                                                                       \
        }                                                                                    \
    } while (0)

#define PCL_ENTER(ctx)                                                                       \
    do {                                                                                     \
        if (!(ctx)) {                                                                        \
            pcl_set_error(nullptr, "%s: null context", __func__);                            \
            return -2;                                                                       \
        }                                                                                    \
        PCL_CUDA(ctx, cudaSetDevice((ctx)->device));                                         \
    } while (0)

#define PCL_LAUNCHED(ctx)                                                                    \
    do {                                                                                     \
        (ctx)->launches++;                                                                   \
        PCL_CUDA(ctx, cudaGetLastError());                                                   \
    } while (0)

static inline bool pcl_aligned16(const void *p) { return (((uintptr_t)p) & 15u) == 0; }

// persistent-style grid for streaming kernels: a multiple of the SM count, capped by the work
static inline unsigned pcl_stream_grid(const pcl_ctx *ctx, uint64_t work_items, int per_block,
                                       int blocks_per_sm) {
    uint64_t need = (work_items + (uint64_t)per_block - 1) / (uint64_t)per_block;
    uint64_t cap = (uint64_t)ctx->sm_count * (uint64_t)blocks_per_sm;
    if (need < 1) need = 1;
    return (unsigned)(need < cap ? need : cap);
}

#ifdef __CUDACC__
#include "pcl_device.cuh"

struct StepK;  // pcl_photon_body.cuh; filled by photon.cu for the pre-compiled and the run-time kernels
int pcl_fill_stepk(pcl_ctx *ctx, StepK &K, float dt, const pcl_scatter_params *sp, const pcl_rng *rng, float escape_r2,
                   const pcl_planes *planes, uint64_t id_base);

#endif  // __CUDACC__
