// Host-buffer entry point: one fused photon step over particle planes that live in HOST memory.
//
// This is the shape of the reference's own per-step marshalling (CLProgram.run,
// physicl/__init__.py:602-664: cl_array.to_device per input :614, launch :656, .get() per output
// :659-662), restated for PCIe Gen5 + B200: the shard is cut into chunks, and chunk c's H2D copies,
// its kernel and its D2H copies are queued on stream c % PIPE_SLOTS, so the two copy engines and
// the SMs all stay busy.  Bytes over PCIe per photon-step: in-place form 24 B up (r, v) + 24 B down
// (+4 B e up); compacting form 28 B up per photon entering + 28 B down per survivor (r, v, id; +4 B e).
#include <stdlib.h>

#include "pcl_common.cuh"

int pcl_photon_step_impl(pcl_ctx *ctx, cudaStream_t st, const pcl_soa *p, const pcl_soa *dst, float dt,
                         const pcl_scatter_params *sp, const pcl_rng *rng, float escape_r2, const pcl_planes *planes,
                         int64_t *tally_row, uint64_t *n_out, uint32_t nsteps, bool keep_count);

int pcl_kinematics_impl(pcl_ctx *ctx, cudaStream_t st, const pcl_soa *p, float dt, int accel, const float *a_uniform,
                        uint32_t nsteps);

#define PIPE_SLOTS 4
#define PIPE_PLANES 11  // x y z vx vy vz e id nscat + u_theta u_phi (u_rand shares a slot below)

struct pcl_hostpipe {
    uint64_t chunk;
    cudaStream_t stream[PIPE_SLOTS];
    float *buf[PIPE_SLOTS][12];
    float *out[PIPE_SLOTS][9];  // compacting form: survivors of the chunk (x y z vx vy vz e id nscat)
    uint64_t *cnt_dev;          // [PIPE_SLOTS] survivors per in-flight chunk
    uint64_t *cnt_pinned;       // [PIPE_SLOTS]
    int64_t *tally_dev;
    int64_t *tally_pinned;
    cudaEvent_t up_done[PIPE_SLOTS];  // zero-copy form: uploads complete in chunk order
    cudaEvent_t ready;
    uint64_t *total_dev;              // survivors of the whole step (all chunks append to one output)
};

void pcl_hostpipe_destroy(pcl_ctx *ctx) {
    pcl_hostpipe *hp = ctx->pipe;
    if (!hp) return;
    for (int s = 0; s < PIPE_SLOTS; ++s) {
        if (hp->stream[s]) cudaStreamDestroy(hp->stream[s]);
        for (int q = 0; q < 12; ++q)
            if (hp->buf[s][q]) cudaFree(hp->buf[s][q]);
        for (int q = 0; q < 9; ++q)
            if (hp->out[s][q]) cudaFree(hp->out[s][q]);
    }
    for (int s = 0; s < PIPE_SLOTS; ++s)
        if (hp->up_done[s]) cudaEventDestroy(hp->up_done[s]);
    if (hp->ready) cudaEventDestroy(hp->ready);
    if (hp->total_dev) cudaFree(hp->total_dev);
    if (hp->cnt_dev) cudaFree(hp->cnt_dev);
    if (hp->cnt_pinned) cudaFreeHost(hp->cnt_pinned);
    if (hp->tally_dev) cudaFree(hp->tally_dev);
    if (hp->tally_pinned) cudaFreeHost(hp->tally_pinned);
    free(hp);
    ctx->pipe = nullptr;
}

static int pipe_alloc(pcl_ctx *ctx, pcl_hostpipe *hp, uint64_t chunk) {
    for (int s = 0; s < PIPE_SLOTS; ++s) {
        PCL_CUDA(ctx, cudaStreamCreateWithFlags(&hp->stream[s], cudaStreamNonBlocking));
        for (int q = 0; q < 12; ++q) PCL_CUDA(ctx, cudaMalloc(&hp->buf[s][q], chunk * sizeof(float)));
    }
    PCL_CUDA(ctx, cudaMalloc(&hp->tally_dev, PCL_FUSE_MAX * PCL_TALLY_COLS * sizeof(int64_t)));
    PCL_CUDA(ctx, cudaMallocHost(&hp->tally_pinned, PCL_FUSE_MAX * PCL_TALLY_COLS * sizeof(int64_t)));
    return 0;
}

static int pipe_prepare(pcl_ctx *ctx, uint64_t chunk) {
    if (ctx->pipe && ctx->pipe->chunk == chunk) return 0;
    pcl_hostpipe_destroy(ctx);
    pcl_hostpipe *hp = (pcl_hostpipe *)calloc(1, sizeof(pcl_hostpipe));
    PCL_REQUIRE(ctx, hp != nullptr, "out of host memory");
    ctx->pipe = hp;
    int rc = pipe_alloc(ctx, hp, chunk);
    if (rc) {  // a half-built pipe must not be found (and reused with null buffers) by the next call
        pcl_hostpipe_destroy(ctx);
        return rc;
    }
    hp->chunk = chunk;
    return 0;
}

extern "C" int pcl_photon_step_host(pcl_ctx *ctx, const pcl_soa *host, float dt, const pcl_scatter_params *sp,
                                    const pcl_rng *rng, float escape_r2, const pcl_planes *planes,
                                    int64_t *tally_row_host, uint64_t chunk) {
    PCL_ENTER(ctx);
    PCL_REQUIRE(ctx, host && sp && rng && tally_row_host, "null argument");
    PCL_REQUIRE(ctx, host->x && host->y && host->z && host->vx && host->vy && host->vz, "r and v planes are required");
    if (chunk == 0) chunk = 1u << 20;
    chunk = (chunk + 3) & ~(uint64_t)3;
    int rc = pipe_prepare(ctx, chunk);
    if (rc) return rc;
    pcl_hostpipe *hp = ctx->pipe;
    const bool wave = sp->mode & PCL_SCATTER_WAVELENGTH;
    if (wave) PCL_REQUIRE(ctx, host->e != nullptr, "wavelength law needs the e plane");
    const bool inj = rng->u_rand != nullptr;
    PCL_CUDA(ctx, cudaMemsetAsync(hp->tally_dev, 0, PCL_TALLY_COLS * sizeof(int64_t), hp->stream[0]));
    PCL_CUDA(ctx, cudaStreamSynchronize(hp->stream[0]));
    uint64_t nchunks = (host->n + chunk - 1) / chunk;
    for (uint64_t c = 0; c < nchunks; ++c) {
        const int s = (int)(c % PIPE_SLOTS);
        cudaStream_t st = hp->stream[s];
        const uint64_t off = c * chunk;
        const uint64_t m = (host->n - off < chunk) ? host->n - off : chunk;
        const size_t bytes = m * sizeof(float);
        float **b = hp->buf[s];
        const float *src[6] = {host->x, host->y, host->z, host->vx, host->vy, host->vz};
        for (int q = 0; q < 6; ++q)
            PCL_CUDA(ctx, cudaMemcpyAsync(b[q], src[q] + off, bytes, cudaMemcpyHostToDevice, st));
        pcl_soa d;
        memset(&d, 0, sizeof(d));
        d.n = m;
        d.x = b[0]; d.y = b[1]; d.z = b[2]; d.vx = b[3]; d.vy = b[4]; d.vz = b[5];
        d.id_base = host->id_base + (host->id ? 0 : off);
        if (wave) {
            PCL_CUDA(ctx, cudaMemcpyAsync(b[6], host->e + off, bytes, cudaMemcpyHostToDevice, st));
            d.e = b[6];
        }
        if (host->id) {
            PCL_CUDA(ctx, cudaMemcpyAsync(b[7], host->id + off, bytes, cudaMemcpyHostToDevice, st));
            d.id = (uint32_t *)b[7];
        }
        if (host->nscat) {
            PCL_CUDA(ctx, cudaMemcpyAsync(b[8], host->nscat + off, bytes, cudaMemcpyHostToDevice, st));
            d.nscat = (uint32_t *)b[8];
        }
        pcl_rng r = *rng;
        if (inj) {
            const bool del = sp->mode & PCL_SCATTER_DELETE;
            PCL_CUDA(ctx, cudaMemcpyAsync(b[11], rng->u_rand + off, bytes, cudaMemcpyHostToDevice, st));
            r.u_rand = b[11];
            r.u_theta = r.u_phi = nullptr;
            if (!del) {
                PCL_REQUIRE(ctx, rng->u_theta && rng->u_phi, "injected uniforms need u_theta and u_phi");
                PCL_CUDA(ctx, cudaMemcpyAsync(b[9], rng->u_theta + off, bytes, cudaMemcpyHostToDevice, st));
                PCL_CUDA(ctx, cudaMemcpyAsync(b[10], rng->u_phi + off, bytes, cudaMemcpyHostToDevice, st));
                r.u_theta = b[9];
                r.u_phi = b[10];
            }
        }
        rc = pcl_photon_step_impl(ctx, st, &d, nullptr, dt, sp, &r, escape_r2, planes, hp->tally_dev, nullptr, 1, false);
        if (rc) return rc;
        float *dst[6] = {host->x, host->y, host->z, host->vx, host->vy, host->vz};
        for (int q = 0; q < 6; ++q)
            PCL_CUDA(ctx, cudaMemcpyAsync(dst[q] + off, b[q], bytes, cudaMemcpyDeviceToHost, st));
        if (host->nscat)
            PCL_CUDA(ctx, cudaMemcpyAsync(host->nscat + off, b[8], bytes, cudaMemcpyDeviceToHost, st));
    }
    for (int s = 0; s < PIPE_SLOTS; ++s) PCL_CUDA(ctx, cudaStreamSynchronize(hp->stream[s]));
    PCL_CUDA(ctx, cudaMemcpy(hp->tally_pinned, hp->tally_dev, PCL_TALLY_COLS * sizeof(int64_t), cudaMemcpyDeviceToHost));
    memcpy(tally_row_host, hp->tally_pinned, PCL_TALLY_COLS * sizeof(int64_t));
    return 0;
}

// ---------------------------------------------------------------------------------------------
// Compacting host-buffer step: same timestep, but each chunk runs the retire-and-compact kernel and
// only the SURVIVORS travel back, written densely from the front of the host planes (the host-side
// meaning of the reference's sim.remove_obj, physicl/__init__.py:455-459).  PCIe bytes per step are
// proportional to live photons: 28 B up + 28 B down each (r, v, id).  Chunks are completed in order
// so the output offset is a running sum of survivor counts; an output region never overtakes a chunk
// that has not been uploaded yet because a chunk yields at most as many photons as it received.
// ---------------------------------------------------------------------------------------------
static int host_compact_staged(pcl_ctx *ctx, const pcl_soa *host, float dt, const pcl_scatter_params *sp,
                               const pcl_rng *rng, float escape_r2, const pcl_planes *planes,
                               int64_t *tally_row_host, uint64_t chunk, uint64_t *n_out_host, uint32_t nsteps) {
    PCL_REQUIRE(ctx, nsteps >= 1 && nsteps <= PCL_FUSE_MAX, "1 to 8 timesteps per host round trip");
    PCL_REQUIRE(ctx, host && sp && rng && tally_row_host && n_out_host, "null argument");
    PCL_REQUIRE(ctx, host->x && host->y && host->z && host->vx && host->vy && host->vz && host->id,
                "r, v and id planes are required (ids travel with the photons)");
    PCL_REQUIRE(ctx, rng->u_rand == nullptr, "the compacting host step draws from Philox");
    if (chunk == 0) chunk = 1u << 20;
    chunk = (chunk + 3) & ~(uint64_t)3;
    int rc = pipe_prepare(ctx, chunk);
    if (rc) return rc;
    pcl_hostpipe *hp = ctx->pipe;
    const bool wave = sp->mode & PCL_SCATTER_WAVELENGTH;
    if (wave) PCL_REQUIRE(ctx, host->e != nullptr, "wavelength law needs the e plane");
    if (!hp->out[PIPE_SLOTS - 1][8]) {  // the LAST buffer: an allocation that failed part-way is resumed, not skipped
        if (!hp->cnt_dev) PCL_CUDA(ctx, cudaMalloc(&hp->cnt_dev, PIPE_SLOTS * sizeof(uint64_t)));
        if (!hp->cnt_pinned) PCL_CUDA(ctx, cudaMallocHost(&hp->cnt_pinned, PIPE_SLOTS * sizeof(uint64_t)));
        for (int s = 0; s < PIPE_SLOTS; ++s)
            for (int q = 0; q < 9; ++q)
                if (!hp->out[s][q]) PCL_CUDA(ctx, cudaMalloc(&hp->out[s][q], chunk * sizeof(float)));
    }
    PCL_CUDA(ctx, cudaMemsetAsync(hp->tally_dev, 0, (size_t)nsteps * PCL_TALLY_COLS * sizeof(int64_t), hp->stream[0]));
    PCL_CUDA(ctx, cudaStreamSynchronize(hp->stream[0]));
    const uint64_t nchunks = (host->n + chunk - 1) / chunk;
    uint64_t out_off = 0;
    // planes that travel: index into buf/out
    float *hplane[9] = {host->x, host->y, host->z, host->vx, host->vy, host->vz, host->e,
                        (float *)host->id, (float *)host->nscat};
    auto complete = [&](uint64_t c) -> int {  // read the survivor count of chunk c, queue its D2H copies
        const int s = (int)(c % PIPE_SLOTS);
        cudaStream_t st = hp->stream[s];
        PCL_CUDA(ctx, cudaStreamSynchronize(st));
        const uint64_t cnt = hp->cnt_pinned[s];
        for (int q = 0; q < 9; ++q)
            if (hplane[q] && cnt)
                PCL_CUDA(ctx, cudaMemcpyAsync(hplane[q] + out_off, hp->out[s][q], cnt * sizeof(float), cudaMemcpyDeviceToHost, st));
        out_off += cnt;
        return 0;
    };
    for (uint64_t c = 0; c < nchunks; ++c) {
        const int s = (int)(c % PIPE_SLOTS);
        cudaStream_t st = hp->stream[s];
        const uint64_t off = c * chunk;
        const uint64_t m = (host->n - off < chunk) ? host->n - off : chunk;
        pcl_soa src, dst;
        memset(&src, 0, sizeof(src));
        memset(&dst, 0, sizeof(dst));
        src.n = dst.n = m;
        src.id_base = dst.id_base = host->id_base;
        float **in = hp->buf[s], **ou = hp->out[s];
        for (int q = 0; q < 9; ++q)
            if (hplane[q]) PCL_CUDA(ctx, cudaMemcpyAsync(in[q], hplane[q] + off, m * sizeof(float), cudaMemcpyHostToDevice, st));
        src.x = in[0]; src.y = in[1]; src.z = in[2]; src.vx = in[3]; src.vy = in[4]; src.vz = in[5];
        dst.x = ou[0]; dst.y = ou[1]; dst.z = ou[2]; dst.vx = ou[3]; dst.vy = ou[4]; dst.vz = ou[5];
        if (host->e) { src.e = in[6]; dst.e = ou[6]; }
        src.id = (uint32_t *)in[7]; dst.id = (uint32_t *)ou[7];
        if (host->nscat) { src.nscat = (uint32_t *)in[8]; dst.nscat = (uint32_t *)ou[8]; }
        rc = pcl_photon_step_impl(ctx, st, &src, &dst, dt, sp, rng, escape_r2, planes, hp->tally_dev, hp->cnt_dev + s, nsteps, false);
        if (rc) return rc;
        PCL_CUDA(ctx, cudaMemcpyAsync(hp->cnt_pinned + s, hp->cnt_dev + s, sizeof(uint64_t), cudaMemcpyDeviceToHost, st));
        if (c + 1 >= PIPE_SLOTS) {  // oldest chunk in flight: its slot is needed next
            rc = complete(c + 1 - PIPE_SLOTS);
            if (rc) return rc;
        }
    }
    for (uint64_t c = (nchunks >= PIPE_SLOTS ? nchunks - PIPE_SLOTS + 1 : 0); c < nchunks; ++c) {
        rc = complete(c);
        if (rc) return rc;
    }
    for (int s = 0; s < PIPE_SLOTS; ++s) PCL_CUDA(ctx, cudaStreamSynchronize(hp->stream[s]));
    PCL_CUDA(ctx, cudaMemcpy(hp->tally_pinned, hp->tally_dev, (size_t)nsteps * PCL_TALLY_COLS * sizeof(int64_t), cudaMemcpyDeviceToHost));
    memcpy(tally_row_host, hp->tally_pinned, (size_t)nsteps * PCL_TALLY_COLS * sizeof(int64_t));
    *n_out_host = out_off;
    return 0;
}

// nsteps (1..8) timesteps per host round trip: each chunk is uploaded once, advanced nsteps timesteps in registers
// by ONE launch of the retire-and-compact kernel, and its survivors come back.  For callers that do not need the
// particles on the host after every single timestep (the reference's Simulation.run with no host step in between,
// physicl/__init__.py:512-516), PCIe bytes per photon-step drop by nsteps.  tally_rows_host: int64[nsteps][COLS].
extern "C" int pcl_photon_steps_host_compact(pcl_ctx *ctx, const pcl_soa *host, float dt, const pcl_scatter_params *sp,
                                             const pcl_rng *rng, float escape_r2, const pcl_planes *planes,
                                             int64_t *tally_rows_host, uint64_t chunk, uint32_t nsteps,
                                             uint64_t *n_out_host) {
    PCL_ENTER(ctx);
    return host_compact_staged(ctx, host, dt, sp, rng, escape_r2, planes, tally_rows_host, chunk, n_out_host, nsteps);
}

// Is this host plane page-locked and mapped into the device's address space (torch pin_memory(),
// cudaHostAlloc, pcl_host_register)?  Then kernels can store to it directly; *dev gets the alias.
static bool host_plane_mapped(const void *h, void **dev) {
    cudaPointerAttributes a;
    if (cudaPointerGetAttributes(&a, h) != cudaSuccess) {
        cudaGetLastError();
        return false;
    }
    if (a.type != cudaMemoryTypeHost || a.devicePointer == nullptr) return false;
    *dev = a.devicePointer;
    return true;
}

// Zero-copy form of the compacting host-buffer step (PCL_HOST_ZEROCOPY=1 and every plane page-locked and mapped):
// uploads go through the copy engine chunk by chunk, and each chunk's kernel stores its survivors
// STRAIGHT into the host planes over PCIe (coalesced runs out of the shared-memory stage), appending
// at one device-side counter shared by all chunks.  No staging buffers on the way back, no survivor
// counts read by the host, one synchronisation per timestep; the two PCIe directions run concurrently.
// In-place safety: uploads complete in chunk order (event chain), a kernel starts after its own
// upload, and at any moment the survivors written so far number at most the slots of the chunks whose
// kernels have started, so they only overwrite host regions that are already on the device.
extern "C" int pcl_photon_step_host_compact(pcl_ctx *ctx, const pcl_soa *host, float dt, const pcl_scatter_params *sp,
                                            const pcl_rng *rng, float escape_r2, const pcl_planes *planes,
                                            int64_t *tally_row_host, uint64_t chunk, uint64_t *n_out_host) {
    PCL_ENTER(ctx);
    PCL_REQUIRE(ctx, host && sp && rng && tally_row_host && n_out_host, "null argument");
    PCL_REQUIRE(ctx, host->x && host->y && host->z && host->vx && host->vy && host->vz && host->id,
                "r, v and id planes are required (ids identify photons afterwards)");
    PCL_REQUIRE(ctx, rng->u_rand == nullptr, "the compacting host step draws from Philox");
    const bool wave = sp->mode & PCL_SCATTER_WAVELENGTH;
    if (wave) PCL_REQUIRE(ctx, host->e != nullptr, "wavelength law needs the e plane");
    float *hplane[9] = {host->x, host->y, host->z, host->vx, host->vy, host->vz, host->e,
                        (float *)host->id, (float *)host->nscat};
    float *dplane[9] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
    bool mapped = true;
    // measured on B200 / PCIe Gen5 (20 steps of the default bench): staged copies 1.40 G photon-steps/s,
    // zero-copy stores 1.17-1.22 G, so the staged form is the default and this one is opt-in
    const char *zc = getenv("PCL_HOST_ZEROCOPY");
    const bool zero_copy = zc && atoi(zc) != 0;
    for (int q = 0; q < 9 && mapped; ++q)
        if (hplane[q]) mapped = host_plane_mapped(hplane[q], (void **)&dplane[q]);
    if (!mapped || !zero_copy)
        return host_compact_staged(ctx, host, dt, sp, rng, escape_r2, planes, tally_row_host, chunk, n_out_host, 1);
    if (chunk == 0) chunk = 1u << 20;
    chunk = (chunk + 3) & ~(uint64_t)3;
    int rc = pipe_prepare(ctx, chunk);
    if (rc) return rc;
    pcl_hostpipe *hp = ctx->pipe;
    if (!hp->total_dev) {
        PCL_CUDA(ctx, cudaMalloc(&hp->total_dev, sizeof(uint64_t)));
        PCL_CUDA(ctx, cudaEventCreateWithFlags(&hp->ready, cudaEventDisableTiming));
        for (int s = 0; s < PIPE_SLOTS; ++s) PCL_CUDA(ctx, cudaEventCreateWithFlags(&hp->up_done[s], cudaEventDisableTiming));
    }
    if (!hp->cnt_pinned) PCL_CUDA(ctx, cudaMallocHost(&hp->cnt_pinned, PIPE_SLOTS * sizeof(uint64_t)));
    PCL_CUDA(ctx, cudaMemsetAsync(hp->tally_dev, 0, PCL_TALLY_COLS * sizeof(int64_t), hp->stream[0]));
    PCL_CUDA(ctx, cudaMemsetAsync(hp->total_dev, 0, sizeof(uint64_t), hp->stream[0]));
    PCL_CUDA(ctx, cudaEventRecord(hp->ready, hp->stream[0]));
    const uint64_t nchunks = (host->n + chunk - 1) / chunk;
    pcl_soa dst;
    memset(&dst, 0, sizeof(dst));
    dst.n = host->n;
    dst.id_base = host->id_base;
    dst.x = dplane[0]; dst.y = dplane[1]; dst.z = dplane[2]; dst.vx = dplane[3]; dst.vy = dplane[4]; dst.vz = dplane[5];
    dst.e = dplane[6];
    dst.id = (uint32_t *)dplane[7];
    dst.nscat = (uint32_t *)dplane[8];
    for (uint64_t c = 0; c < nchunks; ++c) {
        const int s = (int)(c % PIPE_SLOTS);
        cudaStream_t st = hp->stream[s];
        const uint64_t off = c * chunk;
        const uint64_t m = (host->n - off < chunk) ? host->n - off : chunk;
        if (c < PIPE_SLOTS) PCL_CUDA(ctx, cudaStreamWaitEvent(st, hp->ready, 0));
        if (c > 0) PCL_CUDA(ctx, cudaStreamWaitEvent(st, hp->up_done[(c - 1) % PIPE_SLOTS], 0));
        float **in = hp->buf[s];
        for (int q = 0; q < 9; ++q)
            if (hplane[q]) PCL_CUDA(ctx, cudaMemcpyAsync(in[q], hplane[q] + off, m * sizeof(float), cudaMemcpyHostToDevice, st));
        PCL_CUDA(ctx, cudaEventRecord(hp->up_done[s], st));
        pcl_soa src;
        memset(&src, 0, sizeof(src));
        src.n = m;
        src.id_base = host->id_base;
        src.x = in[0]; src.y = in[1]; src.z = in[2]; src.vx = in[3]; src.vy = in[4]; src.vz = in[5];
        if (host->e) src.e = in[6];
        src.id = (uint32_t *)in[7];
        if (host->nscat) src.nscat = (uint32_t *)in[8];
        pcl_soa d = dst;
        d.n = m;
        rc = pcl_photon_step_impl(ctx, st, &src, &d, dt, sp, rng, escape_r2, planes, hp->tally_dev, hp->total_dev, 1, true);
        if (rc) return rc;
    }
    for (int s = 0; s < PIPE_SLOTS; ++s) PCL_CUDA(ctx, cudaStreamSynchronize(hp->stream[s]));
    PCL_CUDA(ctx, cudaMemcpy(hp->tally_pinned, hp->tally_dev, PCL_TALLY_COLS * sizeof(int64_t), cudaMemcpyDeviceToHost));
    PCL_CUDA(ctx, cudaMemcpy(hp->cnt_pinned, hp->total_dev, sizeof(uint64_t), cudaMemcpyDeviceToHost));
    memcpy(tally_row_host, hp->tally_pinned, PCL_TALLY_COLS * sizeof(int64_t));
    *n_out_host = hp->cnt_pinned[0];
    return 0;
}

// ---------------------------------------------------------------------------------------------
// Kinematics over HOST planes: the same chunked pipeline for NewtonianKinematicsStep (physicl/newton.py:14-16; with
// accel, the constant-acceleration law of BASELINE configs[0]).  A chunk goes up once (r, v [, a]), is advanced nsteps
// timesteps in registers by one launch (pcl_kinematics_steps) and comes back (r [, v when accel] [, dr when the host
// has dr planes]); a planes are inputs only.  Bytes over PCIe per particle and round trip with a and dr planes:
// 36 B up + 36 B down.
// ---------------------------------------------------------------------------------------------
extern "C" int pcl_kinematics_steps_host(pcl_ctx *ctx, const pcl_soa *host, float dt, int accel, const float *a_uniform,
                                         uint32_t nsteps, uint64_t chunk) {
    PCL_ENTER(ctx);
    PCL_REQUIRE(ctx, host != nullptr, "null particle view");
    PCL_REQUIRE(ctx, host->x && host->y && host->z && host->vx && host->vy && host->vz, "r and v planes are required");
    PCL_REQUIRE(ctx, host->n_dev == nullptr, "host planes carry their slot count in n");
    const bool has_a = accel && host->ax != nullptr, has_d = host->dx != nullptr;
    if (has_a) PCL_REQUIRE(ctx, host->ay && host->az, "a planes must come as a triple");
    if (has_d) PCL_REQUIRE(ctx, host->dy && host->dz, "dr planes must come as a triple");
    if (accel && !has_a) PCL_REQUIRE(ctx, a_uniform != nullptr, "accel requested without a planes or a_uniform");
    if (nsteps == 0 || host->n == 0) return 0;
    if (chunk == 0) chunk = 1u << 20;
    chunk = (chunk + 3) & ~(uint64_t)3;
    int rc = pipe_prepare(ctx, chunk);
    if (rc) return rc;
    pcl_hostpipe *hp = ctx->pipe;
    float *hplane[12] = {host->x, host->y, host->z, host->vx, host->vy, host->vz, host->ax, host->ay, host->az,
                         host->dx, host->dy, host->dz};
    const uint64_t nchunks = (host->n + chunk - 1) / chunk;
    for (uint64_t c = 0; c < nchunks; ++c) {
        const int s = (int)(c % PIPE_SLOTS);
        cudaStream_t st = hp->stream[s];  // chunk c + PIPE_SLOTS reuses these buffers: stream order keeps it behind c's D2H
        const uint64_t off = c * chunk;
        const uint64_t m = (host->n - off < chunk) ? host->n - off : chunk;
        const size_t bytes = m * sizeof(float);
        float **b = hp->buf[s];
        for (int q = 0; q < (has_a ? 9 : 6); ++q)
            PCL_CUDA(ctx, cudaMemcpyAsync(b[q], hplane[q] + off, bytes, cudaMemcpyHostToDevice, st));
        pcl_soa d;
        memset(&d, 0, sizeof(d));
        d.n = m;
        d.x = b[0]; d.y = b[1]; d.z = b[2]; d.vx = b[3]; d.vy = b[4]; d.vz = b[5];
        if (has_a) { d.ax = b[6]; d.ay = b[7]; d.az = b[8]; }
        if (has_d) { d.dx = b[9]; d.dy = b[10]; d.dz = b[11]; }
        rc = pcl_kinematics_impl(ctx, st, &d, dt, accel, a_uniform, nsteps);
        if (rc) return rc;
        for (int q = 0; q < 12; ++q) {
            const bool back = q < 3 || (q < 6 && accel) || (q >= 9 && has_d);
            if (back) PCL_CUDA(ctx, cudaMemcpyAsync(hplane[q] + off, b[q], bytes, cudaMemcpyDeviceToHost, st));
        }
    }
    for (int s = 0; s < PIPE_SLOTS; ++s) PCL_CUDA(ctx, cudaStreamSynchronize(hp->stream[s]));
    return 0;
}
