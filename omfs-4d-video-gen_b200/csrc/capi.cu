// capi.cu — error plumbing, device check and the level-2 render session of include/omfs_b200.h.
//
// The session is the in-process replacement for the body of the reference's
// render_with_gaussians (02_Visual_Engine/render_surgery.py:245-362): model and avatar are
// uploaded once and stay resident in HBM; frames are processed in batches of `max_batch`
// segments (frame x view), every stage of a batch being ONE launch over all its segments.
// FLAME evaluation (pose prep, tensor-core blendshape GEMM, skinning) runs over larger frame
// chunks so that the GEMM sees M >= 128 rows whenever the clip is long enough.
// Finished frames leave through a second stream (double-buffered images) so the device->host
// copy of batch i overlaps the kernels of batch i+1.
#include <stdarg.h>
#include <string.h>

#include <algorithm>
#include <chrono>
#include <cstdlib>
#include <vector>

#include "common.cuh"

namespace omfs {

static thread_local char g_err[1024] = "";
unsigned long long g_launches = 0;

void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

int cuda_fail(cudaError_t e, const char* what, const char* file, int line) {
    set_error("CUDA error %d (%s) at %s:%d: %s", (int)e, cudaGetErrorString(e), file, line, what);
    return OMFS_ERR_CUDA;
}

}  // namespace omfs

using namespace omfs;

extern "C" int omfs_flame_fold_subject(int V, int n_shape, int npad, const float* d_template,
                                       const float* d_shapedirs, const float* d_shape, const float* d_static,
                                       const float* d_plan, const float* d_jreg, float* d_base, void* stream);
extern "C" int omfs_to_uint8(int S, int width, int height, const float* d_image, uint8_t* d_out, void* stream);
namespace omfs {
int png_encode_launch(int S, int width, int height, const uint8_t* d_frames, uint8_t* d_png, size_t png_capacity,
                      unsigned long long* d_offsets, void* d_workspace, size_t workspace_bytes, cudaStream_t stream);
}

extern "C" const char* omfs_last_error(void) { return g_err; }
extern "C" int omfs_abi_version(void) { return OMFS_ABI_VERSION; }
extern "C" unsigned long long omfs_launch_count(void) { return g_launches; }

extern "C" int omfs_device_check(int dev) {
    int count = 0;
    cudaError_t e = cudaGetDeviceCount(&count);
    if (e != cudaSuccess || dev < 0 || dev >= count) {
        set_error("no CUDA device %d (count=%d, %s)", dev, count, cudaGetErrorString(e));
        return OMFS_ERR_UNSUPPORTED;
    }
    cudaDeviceProp p;
    OMFS_CUDA(cudaGetDeviceProperties(&p, dev));
    if (p.major != 10) {
        set_error("device %d is sm_%d%d; this library is built for sm_100a (B200) only", dev, p.major, p.minor);
        return OMFS_ERR_UNSUPPORTED;
    }
    return OMFS_OK;
}

extern "C" int omfs_device_count(void) {
    int count = 0;
    return cudaGetDeviceCount(&count) == cudaSuccess ? count : 0;
}
extern "C" int omfs_set_device(int dev) {
    int rc = omfs_device_check(dev);
    if (rc) return rc;
    OMFS_CUDA(cudaSetDevice(dev));
    return OMFS_OK;
}

extern "C" int omfs_host_alloc(void** p, size_t bytes) {
    OMFS_REQUIRE(p != nullptr, "null pointer");
    OMFS_CUDA(cudaMallocHost(p, bytes));
    return OMFS_OK;
}
extern "C" int omfs_host_free(void* p) {
    OMFS_CUDA(cudaFreeHost(p));
    return OMFS_OK;
}
extern "C" int omfs_device_alloc(void** p, size_t bytes) {
    OMFS_REQUIRE(p != nullptr, "null pointer");
    OMFS_CUDA(cudaMalloc(p, bytes));
    return OMFS_OK;
}
extern "C" int omfs_device_free(void* p) {
    OMFS_CUDA(cudaFree(p));
    return OMFS_OK;
}
extern "C" int omfs_memcpy_h2d(void* d_dst, const void* h_src, size_t bytes) {
    OMFS_CUDA(cudaMemcpy(d_dst, h_src, bytes, cudaMemcpyHostToDevice));
    return OMFS_OK;
}
extern "C" int omfs_memcpy_d2h(void* h_dst, const void* d_src, size_t bytes) {
    OMFS_CUDA(cudaDeviceSynchronize());
    OMFS_CUDA(cudaMemcpy(h_dst, d_src, bytes, cudaMemcpyDeviceToHost));
    return OMFS_OK;
}
extern "C" int omfs_device_memset(void* d_dst, int value, size_t bytes) {
    OMFS_REQUIRE(d_dst != nullptr || bytes == 0, "null pointer");
    OMFS_CUDA(cudaMemset(d_dst, value, bytes));
    return OMFS_OK;
}
extern "C" int omfs_device_sync(void) {
    OMFS_CUDA(cudaDeviceSynchronize());
    return OMFS_OK;
}

// ---- frame exchange (include/omfs_b200.h): CUDA IPC + copy-engine peer-to-peer pushes
static_assert(sizeof(cudaIpcMemHandle_t) == OMFS_IPC_HANDLE_BYTES, "IPC handle size");
extern "C" int omfs_ipc_export(void* d_ptr, void* h_handle) {
    OMFS_REQUIRE(d_ptr && h_handle, "null pointer");
    cudaIpcMemHandle_t h;
    OMFS_CUDA(cudaIpcGetMemHandle(&h, d_ptr));
    memcpy(h_handle, &h, sizeof(h));
    return OMFS_OK;
}
extern "C" int omfs_ipc_open(const void* h_handle, void** d_ptr) {
    OMFS_REQUIRE(d_ptr && h_handle, "null pointer");
    cudaIpcMemHandle_t h;
    memcpy(&h, h_handle, sizeof(h));
    OMFS_CUDA(cudaIpcOpenMemHandle(d_ptr, h, cudaIpcMemLazyEnablePeerAccess));
    return OMFS_OK;
}
extern "C" int omfs_ipc_close(void* d_ptr) {
    OMFS_REQUIRE(d_ptr, "null pointer");
    OMFS_CUDA(cudaIpcCloseMemHandle(d_ptr));
    return OMFS_OK;
}
extern "C" int omfs_push_frames(void* d_dst, const void* d_src, size_t bytes, void* stream) {
    OMFS_REQUIRE(d_dst && d_src, "null pointer");
    if (bytes == 0) return OMFS_OK;
    // cudaMemcpyDefault: UVA resolves local vs peer; a device-to-device copy is run by the copy engine of
    // the device that owns `stream`, i.e. the sender pushes
    OMFS_CUDA(cudaMemcpyAsync(d_dst, d_src, bytes, cudaMemcpyDefault, (cudaStream_t)stream));
    return OMFS_OK;
}

// --------------------------------------------------------------------------------------------
struct DevBuf {
    void* p = nullptr;
    size_t bytes = 0;
    int ensure(size_t need) {
        if (need <= bytes) return OMFS_OK;
        if (p) cudaFree(p);
        p = nullptr;
        bytes = 0;
        const size_t want = need + need / 4;
        cudaError_t e = cudaMalloc(&p, want);
        if (e != cudaSuccess) {
            p = nullptr;
            set_error("cudaMalloc(%zu) failed: %s", want, cudaGetErrorString(e));
            return OMFS_ERR_NOMEM;
        }
        bytes = want;
        return OMFS_OK;
    }
    void release() {
        if (p) cudaFree(p);
        p = nullptr;
        bytes = 0;
    }
    template <typename T>
    T* as() const {
        return reinterpret_cast<T*>(p);
    }
};

#ifndef OMFS_COMP_PIPELINED_WARPS
#define OMFS_COMP_PIPELINED_WARPS 20
#endif
constexpr int kCompPipelinedWarps = OMFS_COMP_PIPELINED_WARPS;

// Host destination of the device frame sink for one call: packed PNG streams and their offsets.
struct PngSink {
    uint8_t* h_png = nullptr;
    size_t capacity = 0;
    uint64_t* h_offsets = nullptr;   // [segments + 1]
    size_t written = 0;              // bytes handed to the copy engine so far
    bool streaming = false;          // a submit / collect call: failures are kept for the collect
    int error = OMFS_OK;             // first failure while draining
};

struct omfs_session {
    omfs_session_config cfg{};
    int V = 0, F = 0, n_expr = 0, N = 0, kpad = 0, npad = 0, tiles = 0;
    size_t capacity = 0;
    int geo_chunk = 512;  // frames per FLAME launch group
    cudaStream_t stream = nullptr, copy_stream = nullptr;
    // compositing of batch b runs on comp_stream (lowest priority) while the caller's stream already
    // prepares batch b+1 (geometry, binning): ev_front = "batch's lists are ready", ev_comp = "batch's
    // buffers may be overwritten"
    cudaStream_t comp_stream = nullptr;
    // side stream of the front end (tile-range scan beside the depth sort) and its two events
    cudaStream_t aux_stream = nullptr;
    cudaEvent_t ev_pre = nullptr, ev_scan = nullptr;
    bool fuse_front = true;   // OMFS_FUSE_FRONT=0: the separate histogram / tile-count kernels (A/B, debugging)
    bool last_batch_full = true;   // OMFS_COMP_LAST_FULL=0: the last batch keeps the pipelined warp count (A/B)
    bool defer_join = false;       // omfs_session_set_deferred_join: device-output calls leave their compositing un-joined
    int set_parity = 0;            // deferred join: buffer set the next call starts with (the other one may still composite)
    int comp_pipelined_warps = kCompPipelinedWarps;   // OMFS_COMP_PIPE_WARPS=n: compositing warps per SM beside a front end (A/B)
    cudaEvent_t ev_done[2]{}, ev_copied[2]{}, ev_front[2]{}, ev_comp[2]{};
    bool ev_comp_pending[2]{};
    bool subject_set = false;
    // model (resident)
    DevBuf template_, shapedirs, bt, jreg, weights, faces, xyzb, scale_lo, rot, sh, base;
    // subject staging
    DevBuf shape, static_off, plan_off;
    // per call
    DevBuf expr, rotation, neck, jaw, eyes, transl, dyn, jdyn, cams, seg_frame;
    // per geometry chunk
    DevBuf acoef, rmats, vp, verts;
    // per batch
    // (P0, P1, P2, vals, ranges are what compositing reads: double-buffered for the two-stream pipeline)
    DevBuf ff, P0[2], P1[2], P2[2], tt, depth_keys, vals[2], keys64, ranges[2], counters, tickets, ws, image[2], image_u8[2], cams_in;
    int last_set = 0;
    size_t ws_bytes = 0;
    int last_sorted_buffer = 0, last_image_buffer = 0, last_batch_segments = 0;
    uint64_t stats[4]{};
    // optional per-stage timing (bench.py's roofline pass): events around every stage of every batch
    bool profiling = false;
    std::vector<cudaEvent_t> prof_pool;   // events recorded during a call, resolved after the final sync
    std::vector<int> prof_stage;          // stage id each event OPENS (-1: closes the previous one only)
    double stage_ms[8]{};
    uint64_t stage_calls[8]{};
    uint32_t last_pairs = 0, max_batch_pairs = 0;
    bool capacity_auto = true;   // pair_capacity == 0 at create: the host entry point grows it on overflow
    cudaStream_t user_stream = nullptr;   // stream of the last render_device call
    // device frame sink (png.cu): a ring of packed-PNG buffers, so that the encode of batch b, the device->host copy
    // of batch b-1 and the rendering of batch b+1 overlap.  A slot is reused only after the host has issued (and the
    // copy engine finished) the copy of the batch that used it.
    static constexpr int kPngRing = 4;
    cudaStream_t png_stream = nullptr;
    DevBuf png_ws, png_buf[kPngRing], png_off[kPngRing];
    unsigned long long* png_off_host[kPngRing]{};   // pinned: frame offsets of the slot's batch, [S+1]
    cudaEvent_t ev_png_off[kPngRing]{}, ev_png_copied[kPngRing]{};
    struct PngSlot {
        bool active = false;
        size_t seg0 = 0;
        int S = 0;
        PngSink* sink = nullptr;   // the call the slot's batch belongs to
    } png_slot[kPngRing];
    // streaming host calls (omfs_session_submit_host_png / omfs_session_collect_host_png): up to kMaxPending calls
    // whose frames are still on their way to host memory
    static constexpr int kMaxPending = 3;
    PngSink* pending[kMaxPending]{};
    int n_pending = 0;
    long long stream_base = 0;        // batches of all streaming calls so far: ring slot and buffer set continue across calls
    bool copied_pending[2]{};         // ev_copied[ib] has been recorded and not yet waited for by a compositing launch
    cudaEvent_t ev_collect = nullptr;
    size_t png_frame_cap = 0, png_ws_bytes = 0;
    std::vector<int32_t> seg_frame_host;
    int seg_table_fpb = -1, seg_table_views = -1;
};

enum Stage { kStFlame = 0, kStFaceFrames, kStBindPre, kStDepthSort, kStTileRanges, kStEmitScatter, kStComposite, kStCount };

__global__ void tile_cams_kernel(const float* __restrict__ in, int n_in, int n_out, float* __restrict__ out) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n_out) out[i] = in[i % n_in];
}

static int upload(DevBuf& b, const void* h, size_t bytes, cudaStream_t st) {
    int rc = b.ensure(bytes);
    if (rc) return rc;
    OMFS_CUDA(cudaMemcpyAsync(b.p, h, bytes, cudaMemcpyHostToDevice, st));
    return OMFS_OK;
}

static inline float tf32_hi_host(float x) {
    uint32_t u;
    memcpy(&u, &x, 4);
    u &= 0xffffe000u;
    float r;
    memcpy(&r, &u, 4);
    return r;
}

extern "C" void omfs_session_destroy(omfs_session* s) {
    if (!s) return;
    cudaSetDevice(s->cfg.device);
    if (s->stream) cudaStreamSynchronize(s->stream);
    if (s->copy_stream) cudaStreamSynchronize(s->copy_stream);
    if (s->comp_stream) cudaStreamSynchronize(s->comp_stream);
    if (s->aux_stream) {
        cudaStreamSynchronize(s->aux_stream);
        cudaStreamDestroy(s->aux_stream);
    }
    if (s->ev_pre) cudaEventDestroy(s->ev_pre);
    if (s->ev_scan) cudaEventDestroy(s->ev_scan);
    if (s->png_stream) cudaStreamSynchronize(s->png_stream);
    if (s->ev_collect) cudaEventDestroy(s->ev_collect);
    while (s->n_pending > 0) delete s->pending[--s->n_pending];   // clips submitted and never collected
    DevBuf* all[] = {&s->template_, &s->shapedirs, &s->bt, &s->jreg, &s->weights, &s->faces, &s->xyzb,
                     &s->scale_lo, &s->rot, &s->sh, &s->base, &s->shape, &s->static_off, &s->plan_off,
                     &s->expr, &s->rotation, &s->neck, &s->jaw, &s->eyes, &s->transl, &s->dyn, &s->jdyn,
                     &s->cams, &s->seg_frame, &s->acoef, &s->rmats, &s->vp, &s->verts, &s->ff, &s->P0[0], &s->P1[0],
                     &s->P2[0], &s->P0[1], &s->P1[1], &s->P2[1], &s->tt, &s->depth_keys, &s->vals[0], &s->vals[1],
                     &s->keys64, &s->ranges[0], &s->ranges[1],
                     &s->counters, &s->tickets, &s->cams_in, &s->ws, &s->image[0], &s->image[1], &s->image_u8[0], &s->image_u8[1]};
    for (DevBuf* b : all) b->release();
    for (int i = 0; i < 2; i++) {
        if (s->ev_done[i]) cudaEventDestroy(s->ev_done[i]);
        if (s->ev_copied[i]) cudaEventDestroy(s->ev_copied[i]);
        if (s->ev_front[i]) cudaEventDestroy(s->ev_front[i]);
        if (s->ev_comp[i]) cudaEventDestroy(s->ev_comp[i]);
    }
    if (s->png_stream) cudaStreamSynchronize(s->png_stream);
    s->png_ws.release();
    for (int i = 0; i < omfs_session::kPngRing; i++) {
        s->png_buf[i].release();
        s->png_off[i].release();
        if (s->png_off_host[i]) cudaFreeHost(s->png_off_host[i]);
        if (s->ev_png_off[i]) cudaEventDestroy(s->ev_png_off[i]);
        if (s->ev_png_copied[i]) cudaEventDestroy(s->ev_png_copied[i]);
    }
    if (s->png_stream) cudaStreamDestroy(s->png_stream);
    for (cudaEvent_t e : s->prof_pool) cudaEventDestroy(e);
    if (s->comp_stream) cudaStreamDestroy(s->comp_stream);
    if (s->stream) cudaStreamDestroy(s->stream);
    if (s->copy_stream) cudaStreamDestroy(s->copy_stream);
    delete s;
}

extern "C" int omfs_session_create(const omfs_model_desc* m, const omfs_session_config* cfg, omfs_session** out) {
    OMFS_REQUIRE(m && cfg && out, "null argument");
    OMFS_REQUIRE(m->n_verts > 0 && m->n_faces > 0 && m->n_gauss > 0 && m->n_expr >= 0 && m->n_expr <= 400,
                 "bad model sizes");
    OMFS_REQUIRE(m->v_template && m->shapedirs && m->posedirs && m->j_regressor && m->lbs_weights && m->faces &&
                     m->xyzb && m->scale_lo && m->rot && m->sh,
                 "null model array");
    OMFS_REQUIRE(cfg->width > 0 && cfg->height > 0 && cfg->max_batch > 0 && cfg->max_batch <= 65535, "bad config");
    // The kernels index device memory with these arrays (face_frames: verts[faces[..]], bind_preprocess:
    // ff[binding]): an avatar or mesh from an untrusted file must be rejected here, not discovered as a fault.
    for (long long i = 0; i < 3ll * m->n_faces; i++)
        if (m->faces[i] < 0 || m->faces[i] >= m->n_verts) {
            set_error("omfs_session_create: faces[%lld] = %d is not a vertex index (n_verts = %d)", i / 3,
                      (int)m->faces[i], (int)m->n_verts);
            return OMFS_ERR_INVALID;
        }
    for (long long n = 0; n < m->n_gauss; n++) {
        int32_t b;
        memcpy(&b, &m->xyzb[4 * n + 3], 4);  // binding index rides as raw bits
        if (b < 0 || b >= m->n_faces) {
            set_error("omfs_session_create: Gaussian %lld is bound to face %d, the mesh has %d faces (a GaussianAvatars "
                      "PLY binds to the teeth-augmented FLAME mesh: load the matching model)",
                      n, (int)b, (int)m->n_faces);
            return OMFS_ERR_INVALID;
        }
    }
    int rc = omfs_device_check(cfg->device);
    if (rc) return rc;
    OMFS_CUDA(cudaSetDevice(cfg->device));
    omfs_session* s = new omfs_session();
    s->cfg = *cfg;
    s->V = m->n_verts;
    s->F = m->n_faces;
    s->n_expr = m->n_expr;
    s->N = m->n_gauss;
    s->kpad = ((m->n_expr + OMFS_N_POSE_FEAT) + 7) / 8 * 8;
    s->npad = ((3 * s->V + 15) + 127) / 128 * 128;
    s->tiles = ((cfg->width + kTile - 1) / kTile) * ((cfg->height + kTile - 1) / kTile);
    s->capacity = cfg->pair_capacity ? (size_t)cfg->pair_capacity : (size_t)6 * cfg->max_batch * (size_t)s->N;
    s->capacity_auto = cfg->pair_capacity == 0;
    if (s->capacity >= (1ull << 30)) s->capacity = (1ull << 30) - 1;
    const int V = s->V, V3 = 3 * V, N = s->N;
#define TRY(x)                       \
    do {                             \
        rc = (x);                    \
        if (rc) {                    \
            omfs_session_destroy(s); \
            return rc;               \
        }                            \
    } while (0)
#define TRY_CUDA(x)                                                \
    do {                                                           \
        cudaError_t _e = (x);                                      \
        if (_e != cudaSuccess) {                                   \
            rc = omfs::cuda_fail(_e, #x, __FILE__, __LINE__);      \
            omfs_session_destroy(s);                               \
            return rc;                                             \
        }                                                          \
    } while (0)
    TRY_CUDA(cudaStreamCreateWithFlags(&s->stream, cudaStreamNonBlocking));
    TRY_CUDA(cudaStreamCreateWithFlags(&s->copy_stream, cudaStreamNonBlocking));
    {
        // lowest priority: when both streams have CTAs to place, the next batch's front end goes first and
        // the (issue-bound) compositing grid fills whatever is left
        int prio_least = 0, prio_greatest = 0;
        TRY_CUDA(cudaDeviceGetStreamPriorityRange(&prio_least, &prio_greatest));
        TRY_CUDA(cudaStreamCreateWithPriority(&s->comp_stream, cudaStreamNonBlocking, prio_least));
    }
    TRY_CUDA(cudaStreamCreateWithFlags(&s->aux_stream, cudaStreamNonBlocking));
    TRY_CUDA(cudaEventCreateWithFlags(&s->ev_pre, cudaEventDisableTiming));
    TRY_CUDA(cudaEventCreateWithFlags(&s->ev_scan, cudaEventDisableTiming));
    {
        const char* e = getenv("OMFS_FUSE_FRONT");
        s->fuse_front = !(e && e[0] == '0');
        e = getenv("OMFS_COMP_LAST_FULL");
        s->last_batch_full = !(e && e[0] == '0');
        e = getenv("OMFS_COMP_PIPE_WARPS");
        if (e && atoi(e) > 0 && atoi(e) <= 32) s->comp_pipelined_warps = atoi(e);
    }
    for (int i = 0; i < 2; i++) {
        TRY_CUDA(cudaEventCreateWithFlags(&s->ev_done[i], cudaEventDisableTiming));
        TRY_CUDA(cudaEventCreateWithFlags(&s->ev_copied[i], cudaEventDisableTiming));
        TRY_CUDA(cudaEventCreateWithFlags(&s->ev_front[i], cudaEventDisableTiming));
        TRY_CUDA(cudaEventCreateWithFlags(&s->ev_comp[i], cudaEventDisableTiming));
    }
    cudaStream_t st = s->stream;
    TRY(upload(s->template_, m->v_template, sizeof(float) * V3, st));
    TRY(upload(s->shapedirs, m->shapedirs, sizeof(float) * (size_t)300 * V3, st));
    TRY(upload(s->jreg, m->j_regressor, sizeof(float) * 5 * V, st));
    TRY(upload(s->weights, m->lbs_weights, sizeof(float) * 5 * V, st));
    TRY(upload(s->faces, m->faces, sizeof(int32_t) * 3 * s->F, st));
    TRY(upload(s->xyzb, m->xyzb, sizeof(float) * 4 * (size_t)N, st));
    TRY(upload(s->scale_lo, m->scale_lo, sizeof(float) * 4 * (size_t)N, st));
    TRY(upload(s->rot, m->rot, sizeof(float) * 4 * (size_t)N, st));
    TRY(upload(s->sh, m->sh, sizeof(float) * 48 * (size_t)N, st));
    TRY(s->base.ensure(sizeof(float) * s->npad));

    // B' = [Bh | Bl | Bh] per output column, K-major.  Rows of B: n_expr expression directions, 36
    // pose correctives; columns: 3V vertex coordinates then the 15 regressed joint coordinates.
    {
        const int kpad = s->kpad, K3 = 3 * kpad, npad = s->npad, ne = s->n_expr;
        std::vector<float> bt((size_t)npad * K3, 0.f);
        const float* exprdirs = m->shapedirs + (size_t)300 * V3;
        for (int n = 0; n < V3; n++) {
            float* row = bt.data() + (size_t)n * K3;
            for (int k = 0; k < ne + OMFS_N_POSE_FEAT; k++) {
                const float b = (k < ne) ? exprdirs[(size_t)k * V3 + n] : m->posedirs[(size_t)(k - ne) * V3 + n];
                const float hi = tf32_hi_host(b), lo = tf32_hi_host(b - hi);
                row[k] = hi;
                row[kpad + k] = lo;
                row[2 * kpad + k] = hi;
            }
        }
        for (int j = 0; j < 5; j++)
            for (int c = 0; c < 3; c++) {
                float* row = bt.data() + (size_t)(V3 + j * 3 + c) * K3;
                for (int k = 0; k < ne; k++) {
                    double a = 0.0;
                    const float* d = exprdirs + (size_t)k * V3;
                    const float* jr = m->j_regressor + (size_t)j * V;
                    for (int v = 0; v < V; v++) a += (double)jr[v] * (double)d[v * 3 + c];
                    const float b = (float)a;
                    const float hi = tf32_hi_host(b), lo = tf32_hi_host(b - hi);
                    row[k] = hi;
                    row[kpad + k] = lo;
                    row[2 * kpad + k] = hi;
                }
            }
        TRY(s->bt.ensure(sizeof(float) * bt.size()));
        TRY_CUDA(cudaMemcpy(s->bt.p, bt.data(), sizeof(float) * bt.size(), cudaMemcpyHostToDevice));
    }

    // per-batch buffers
    const size_t Sb = (size_t)cfg->max_batch;
    const size_t hw = (size_t)cfg->width * cfg->height;
    TRY(s->ff.ensure(sizeof(float) * kFF * Sb * s->F));
    for (int i = 0; i < 2; i++) {
        TRY(s->P0[i].ensure(sizeof(float) * 4 * Sb * N));
        TRY(s->P1[i].ensure(sizeof(float) * 4 * Sb * N));
        TRY(s->P2[i].ensure(sizeof(float) * 4 * Sb * N));
        TRY(s->vals[i].ensure(sizeof(uint32_t) * s->capacity));
        TRY(s->ranges[i].ensure(sizeof(uint32_t) * 2 * Sb * s->tiles));
    }
    TRY(s->tt.ensure(sizeof(uint32_t) * Sb * N));
    TRY(s->depth_keys.ensure(sizeof(uint32_t) * Sb * N));
    if (cfg->debug_keys) TRY(s->keys64.ensure(sizeof(uint64_t) * s->capacity));
    TRY(s->counters.ensure(256));
    TRY_CUDA(cudaMemsetAsync(s->counters.p, 0, 256, st));
    // ticket counter of the persistent compositing kernel: zero once, the kernel leaves it zeroed
    TRY(s->tickets.ensure(OMFS_COMPOSITE_TICKET_BYTES));
    TRY_CUDA(cudaMemsetAsync(s->tickets.p, 0, OMFS_COMPOSITE_TICKET_BYTES, st));
    s->ws_bytes = omfs_binning_workspace_bytes((int)Sb, N, cfg->width, cfg->height, s->capacity);
    TRY(s->ws.ensure(s->ws_bytes));
    for (int i = 0; i < 2; i++) {
        TRY(s->image[i].ensure(sizeof(float) * 3 * Sb * hw));
        TRY(s->image_u8[i].ensure(3 * Sb * hw));
    }
    TRY_CUDA(cudaStreamSynchronize(st));
#undef TRY
#undef TRY_CUDA
    *out = s;
    return OMFS_OK;
}

extern "C" int omfs_session_set_subject(omfs_session* s, const float* h_shape300, const float* h_static_offset,
                                        const float* h_plan_offset) {
    OMFS_REQUIRE(s && h_shape300, "null argument");
    OMFS_CUDA(cudaSetDevice(s->cfg.device));
    const int V3 = 3 * s->V;
    int rc = upload(s->shape, h_shape300, sizeof(float) * 300, s->stream);
    if (rc) return rc;
    if (h_static_offset) {
        rc = upload(s->static_off, h_static_offset, sizeof(float) * V3, s->stream);
        if (rc) return rc;
    }
    if (h_plan_offset) {
        rc = upload(s->plan_off, h_plan_offset, sizeof(float) * V3, s->stream);
        if (rc) return rc;
    }
    rc = omfs_flame_fold_subject(s->V, 300, s->npad, s->template_.as<float>(), s->shapedirs.as<float>(),
                                 s->shape.as<float>(), h_static_offset ? s->static_off.as<float>() : nullptr,
                                 h_plan_offset ? s->plan_off.as<float>() : nullptr, s->jreg.as<float>(),
                                 s->base.as<float>(), s->stream);
    if (rc) return rc;
    OMFS_CUDA(cudaStreamSynchronize(s->stream));
    s->subject_set = true;
    return OMFS_OK;
}


// Lazily created state of the device frame sink (streams, ring buffers) for batches of up to max_batch frames.
static int png_prepare(omfs_session* s) {
    if (s->png_stream) return OMFS_OK;
    const int Sb = s->cfg.max_batch, W = s->cfg.width, H = s->cfg.height;
    s->png_frame_cap = omfs_png_max_bytes(W, H);
    s->png_ws_bytes = omfs_png_workspace_bytes(Sb, W, H);
    if (!s->png_frame_cap || !s->png_ws_bytes) {
        set_error("png sink: image size %dx%d is not supported", W, H);
        return OMFS_ERR_UNSUPPORTED;
    }
    int rc;
    if ((rc = s->png_ws.ensure(s->png_ws_bytes))) return rc;
    for (int i = 0; i < omfs_session::kPngRing; i++) {
        if ((rc = s->png_buf[i].ensure(s->png_frame_cap * (size_t)Sb))) return rc;
        if ((rc = s->png_off[i].ensure(sizeof(unsigned long long) * ((size_t)Sb + 1)))) return rc;
        OMFS_CUDA(cudaMallocHost((void**)&s->png_off_host[i], sizeof(unsigned long long) * ((size_t)Sb + 1)));
        OMFS_CUDA(cudaEventCreateWithFlags(&s->ev_png_off[i], cudaEventDisableTiming));
        OMFS_CUDA(cudaEventCreateWithFlags(&s->ev_png_copied[i], cudaEventDisableTiming));
    }
    OMFS_CUDA(cudaStreamCreateWithFlags(&s->png_stream, cudaStreamNonBlocking));
    return OMFS_OK;
}

// The batch in ring slot r has been encoded (or will be shortly): wait for its frame offsets, hand exactly the bytes
// it produced to the copy engine, publish the offsets to the caller.
static int png_drain(omfs_session* s, int r) {
    omfs_session::PngSlot& slot = s->png_slot[r];
    if (!slot.active) return OMFS_OK;
    slot.active = false;
    PngSink* sink = slot.sink;
    OMFS_CUDA(cudaEventSynchronize(s->ev_png_off[r]));
    if (sink->error != OMFS_OK) return OMFS_OK;   // the call already failed: nothing more is copied for it
    const unsigned long long* off = s->png_off_host[r];
    const size_t total = (size_t)off[slot.S];
    if (sink->written + total > sink->capacity) {
        set_error("png sink: the caller's buffer (%zu bytes) is too small: %zu bytes needed so far (size it with "
                  "omfs_png_max_bytes per frame)", sink->capacity, sink->written + total);
        sink->error = OMFS_ERR_CAPACITY;
        return sink->streaming ? OMFS_OK : OMFS_ERR_CAPACITY;   // a streaming call hears of it at its collect
    }
    OMFS_CUDA(cudaMemcpyAsync(sink->h_png + sink->written, s->png_buf[r].p, total, cudaMemcpyDeviceToHost, s->copy_stream));
    OMFS_CUDA(cudaEventRecord(s->ev_png_copied[r], s->copy_stream));
    for (int i = 0; i <= slot.S; i++) sink->h_offsets[slot.seg0 + i] = (uint64_t)(sink->written + off[i]);
    sink->written += total;
    return OMFS_OK;
}

// Core loop.  All `p_*` pointers are DEVICE pointers valid on s->stream.
static int render_core(omfs_session* s, int T, int n_views, const float* p_expr, const float* p_rot,
                       const float* p_neck, const float* p_jaw, const float* p_eyes, const float* p_transl,
                       const float* p_dyn, const float* p_cams, uint8_t* out_u8, float* out_f32, bool out_on_host,
                       cudaStream_t st, PngSink* png = nullptr, bool streaming = false) {
    const int V = s->V, F = s->F, N = s->N, W = s->cfg.width, H = s->cfg.height;
    const size_t hw = (size_t)W * H;
    if (png) {
        int prc = png_prepare(s);
        if (prc) return prc;
        if (!streaming)
            for (auto& slot : s->png_slot) slot.active = false;
    }
    if (!streaming) s->copied_pending[0] = s->copied_pending[1] = false;   // a blocking call left nothing in flight
    const unsigned long long launches0 = g_launches;
    s->stats[0] = s->stats[2] = s->stats[3] = 0;
    int rc;
    const int fpb = std::max(1, s->cfg.max_batch / n_views);  // frames per render batch
    const int geo = std::max(fpb, s->geo_chunk / fpb * fpb);   // frames per FLAME chunk (multiple of fpb)
    if ((rc = s->acoef.ensure(sizeof(float) * 3 * s->kpad * (size_t)geo))) return rc;
    if ((rc = s->rmats.ensure(sizeof(float) * 45 * (size_t)geo))) return rc;
    if ((rc = s->vp.ensure(sizeof(float) * (size_t)s->npad * geo))) return rc;
    if ((rc = s->verts.ensure(sizeof(float) * 3 * (size_t)V * geo))) return rc;
    if (p_dyn && (rc = s->jdyn.ensure(sizeof(float) * 15 * (size_t)geo))) return rc;
    // segment tables for one full batch: seg -> local frame, seg -> camera.  The seg->frame table only
    // depends on (frames per batch, views): it is uploaded once and reused by later calls.
    {
        const int Sb = fpb * n_views;
        if (s->seg_table_fpb != fpb || s->seg_table_views != n_views) {
            s->seg_frame_host.resize(Sb);
            for (int i = 0; i < Sb; i++) s->seg_frame_host[i] = i / n_views;
            if ((rc = s->seg_frame.ensure(sizeof(int32_t) * Sb))) return rc;
            OMFS_CUDA(cudaMemcpyAsync(s->seg_frame.p, s->seg_frame_host.data(), sizeof(int32_t) * Sb,
                                      cudaMemcpyHostToDevice, st));
            OMFS_CUDA(cudaStreamSynchronize(st));
            s->seg_table_fpb = fpb;
            s->seg_table_views = n_views;
        }
        // cameras tiled per frame of the batch: one launch (it was one device-to-device copy per frame of the
        // batch — 60 copy-engine round trips in front of every call)
        if ((rc = s->cams.ensure(sizeof(float) * kCam * Sb))) return rc;
        {
            const int n_in = kCam * n_views, n_out = n_in * fpb;
            tile_cams_kernel<<<ceil_div(n_out, 256), 256, 0, st>>>(p_cams, n_in, n_out, s->cams.as<float>());
            count_launch();
            OMFS_LAUNCH_CHECK();
        }
    }
    uint32_t* d_num_pairs = s->counters.as<uint32_t>();
    int* d_flag = s->counters.as<int>() + 1;
    // (streaming calls keep the counters of the calls before them: the overflow flag of clip i must survive until its
    // collect, which may come after clip i+1 was submitted)
    if (!(streaming && s->stream_base > 0)) OMFS_CUDA(cudaMemsetAsync(s->counters.p, 0, 256, st));

    unsigned long long* d_pair_accum = reinterpret_cast<unsigned long long*>(s->counters.as<unsigned char>() + 16);
    uint32_t* d_pair_max = s->counters.as<uint32_t>() + 2;
    const bool prof = s->profiling;
    const bool pipelined = !prof && s->cfg.max_batch > 0;
    size_t prof_used = 0;
    s->prof_stage.clear();
    // stage timing: mark(stage) records an event that closes the previous stage and opens `stage`
    // (-1 = only close).  Nothing is synchronised here — the CPU keeps running ahead of the GPU, so the
    // intervals contain kernel time, not launch gaps; they are resolved after the call's final sync.
    auto mark = [&](int stage) -> int {
        if (!prof) return OMFS_OK;
        if (prof_used == s->prof_pool.size()) {
            cudaEvent_t e;
            OMFS_CUDA(cudaEventCreate(&e));
            s->prof_pool.push_back(e);
        }
        OMFS_CUDA(cudaEventRecord(s->prof_pool[prof_used++], st));
        s->prof_stage.push_back(stage);
        return OMFS_OK;
    };

    int batch_index = 0;
    // Deferred join (device output only): this call's last compositing launch may still run when the next call's
    // front end starts, so the next call begins with the OTHER buffer set, and the last batch keeps the pipelined
    // occupancy (a front end will run beside it).
    const bool deferred = s->defer_join && !out_on_host;
    // Streaming host calls go one step further: ring slot, image buffer and record set all continue from the
    // previous call, whose last batches may still be encoding or copying.
    const long long idx0 = streaming ? s->stream_base : (deferred ? s->set_parity : 0);
    for (int g0 = 0; g0 < T; g0 += geo) {
        const int gT = std::min(geo, T - g0);
        // ---- FLAME: operand prep, blendshape GEMM (tensor cores), skinning
        if ((rc = mark(kStFlame))) return rc;
        if ((rc = omfs_flame_pose_prep(gT, s->n_expr, s->kpad, p_expr + (size_t)g0 * s->n_expr, p_rot + (size_t)g0 * 3,
                                       p_neck + (size_t)g0 * 3, p_jaw + (size_t)g0 * 3, p_eyes + (size_t)g0 * 6,
                                       s->acoef.as<float>(), s->rmats.as<float>(), st)))
            return rc;
        const int impl = (s->cfg.gemm_impl == 0) ? 0 : 1;
        if ((rc = omfs_flame_blend_gemm(gT, s->kpad, s->npad, s->acoef.as<float>(), s->bt.as<float>(),
                                        s->base.as<float>(), s->vp.as<float>(), impl, st)))
            return rc;
        const float* dyn_g = p_dyn ? p_dyn + (size_t)g0 * V * 3 : nullptr;
        if (dyn_g && (rc = omfs_flame_joint_dyn(gT, V, s->jreg.as<float>(), dyn_g, s->jdyn.as<float>(), st)))
            return rc;
        if ((rc = omfs_flame_lbs(gT, V, s->npad, s->vp.as<float>(), s->rmats.as<float>(), s->weights.as<float>(),
                                 p_transl + (size_t)g0 * 3, dyn_g, dyn_g ? s->jdyn.as<float>() : nullptr,
                                 s->verts.as<float>(), st)))
            return rc;
        if ((rc = mark(-1))) return rc;
        // ---- render batches inside the chunk
        // batch schedule of this chunk: full batches, except that the LAST batch of a synchronous
        // host-output call is tapered (halved down to ~fpb/8 frames).  Inside a call the compositing of
        // batch b overlaps the front end of batch b+1 and the device->host copy of batch b-1; the last
        // compositing pass and the last copy run alone — make those small.  (Tapering the first batch as
        // well was measured slower: 11.25 vs 10.72 ms per 300-frame call.)  A call of one or two batches
        // keeps its batches whole (the debug taps then describe the whole call).
        std::vector<int> sizes;
        for (int b0 = 0; b0 < gT; b0 += fpb) sizes.push_back(std::min(fpb, gT - b0));
        if (out_on_host && pipelined && !streaming && T >= 3 * fpb) {
            const int floor_frames = std::max(2, fpb / 8);
            auto taper = [&](int frames) {  // descending pieces
                std::vector<int> out;
                while (frames > 2 * floor_frames) {
                    out.push_back((frames + 1) / 2);
                    frames -= (frames + 1) / 2;
                }
                if (frames > 0) out.push_back(frames);
                return out;
            };
            if (g0 + gT >= T) {
                std::vector<int> t = taper(sizes.back());
                sizes.pop_back();
                sizes.insert(sizes.end(), t.begin(), t.end());
            }
        }
        int b0 = 0;
        for (size_t bi = 0; bi < sizes.size(); b0 += sizes[bi], bi++) {
            const int bT = sizes[bi];
            const int S = bT * n_views;
            const long long gi = idx0 + batch_index;   // position in the stream of batches this numbering continues
            const int ib = (int)(gi & 1);
            // the buffer set `ib` is free once the compositing of two batches ago has read it
            if (s->ev_comp_pending[ib]) {
                OMFS_CUDA(cudaStreamWaitEvent(st, s->ev_comp[ib], 0));
                s->ev_comp_pending[ib] = false;
            }
            float* P0 = s->P0[ib].as<float>();
            float* P1 = s->P1[ib].as<float>();
            float* P2 = s->P2[ib].as<float>();
            uint32_t* vals = s->vals[ib].as<uint32_t>();
            uint32_t* ranges = s->ranges[ib].as<uint32_t>();
            if ((rc = mark(kStFaceFrames))) return rc;
            if ((rc = omfs_face_frames(bT, V, F, s->verts.as<float>() + (size_t)b0 * V * 3, s->faces.as<int32_t>(),
                                       s->ff.as<float>(), st)))
                return rc;
            if ((rc = mark(kStBindPre))) return rc;
            // the binning's counters are zeroed first: the fused bind + preprocess kernel fills the depth-digit
            // histograms and the tile counts from the values it holds in registers
            uint32_t *d_hist = nullptr, *d_tcnt = nullptr;
            const bool fused = s->fuse_front;
            if (fused && (rc = binning_prepare(S, N, W, H, s->capacity, s->ws.p, &d_hist, &d_tcnt, st))) return rc;
            if ((rc = bind_preprocess_launch(S, N, F, W, H, s->ff.as<float>(), s->seg_frame.as<int32_t>(),
                                             s->cams.as<float>(), s->xyzb.as<float>(), s->scale_lo.as<float>(),
                                             s->rot.as<float>(), s->sh.as<float>(), P0, P1, P2, s->tt.as<uint32_t>(),
                                             s->depth_keys.as<uint32_t>(), d_hist, d_tcnt, st)))
                return rc;
            // The tile-range scan (ONE CTA, ~35 us of latency) needs only the tile counts: with the fused front end it
            // runs on a side stream beside the depth sort instead of between it and the emission.  Stage timing keeps
            // everything on one stream so that its intervals mean one stage each.
            const bool side_scan = fused && pipelined;
            if (side_scan) {
                OMFS_CUDA(cudaEventRecord(s->ev_pre, st));
                OMFS_CUDA(cudaStreamWaitEvent(s->aux_stream, s->ev_pre, 0));
                if ((rc = binning_tile_ranges(S, N, W, H, s->capacity, P0, s->tt.as<uint32_t>(), ranges, d_num_pairs,
                                              d_flag, d_pair_accum, d_pair_max, s->ws.p, s->aux_stream, true)))
                    return rc;
                OMFS_CUDA(cudaEventRecord(s->ev_scan, s->aux_stream));
            }
            if ((rc = mark(kStDepthSort))) return rc;
            if ((rc = binning_depth_sort(S, N, W, H, s->capacity, s->depth_keys.as<uint32_t>(), s->ws.p, st, fused)))
                return rc;
            if ((rc = mark(kStTileRanges))) return rc;
            if (side_scan) {
                OMFS_CUDA(cudaStreamWaitEvent(st, s->ev_scan, 0));
            } else if ((rc = binning_tile_ranges(S, N, W, H, s->capacity, P0, s->tt.as<uint32_t>(), ranges, d_num_pairs,
                                                 d_flag, d_pair_accum, d_pair_max, s->ws.p, st, fused))) {
                return rc;
            }
            if ((rc = mark(kStEmitScatter))) return rc;
            if ((rc = binning_emit_scatter(S, N, W, H, s->capacity, P0, s->tt.as<uint32_t>(), vals, s->ws.p, st)))
                return rc;
            if (s->cfg.debug_keys &&
                (rc = binning_rebuild_keys(S, N, W, H, s->capacity, ranges, vals, s->depth_keys.as<uint32_t>(),
                                           s->keys64.as<uint64_t>(), s->ws.p, st)))
                return rc;
            // ---- compositing: on its own stream unless stage timing is on (then everything stays in order
            // on the caller's stream so that the event intervals mean one stage each)
            cudaStream_t cst = pipelined ? s->comp_stream : st;
            if (pipelined) {
                OMFS_CUDA(cudaEventRecord(s->ev_front[ib], st));
                OMFS_CUDA(cudaStreamWaitEvent(cst, s->ev_front[ib], 0));
            }
            // the image buffer may still be draining to the host from two batches ago
            if (out_on_host && s->copied_pending[ib]) {
                OMFS_CUDA(cudaStreamWaitEvent(cst, s->ev_copied[ib], 0));
                s->copied_pending[ib] = false;
            }
            float* img = s->image[ib].as<float>();
            uint8_t* img8 = s->image_u8[ib].as<uint8_t>();
            const size_t seg0 = (size_t)(g0 + b0) * n_views;
            const bool direct = !out_on_host;  // device output: composite straight into the caller's buffers
            float* dst_f = direct ? (out_f32 ? out_f32 + seg0 * 3 * hw : nullptr) : (out_f32 ? img : nullptr);
            uint8_t* dst_8 = direct ? (out_u8 ? out_u8 + seg0 * 3 * hw : nullptr) : ((out_u8 || png) ? img8 : nullptr);
            // The sink keeps two batches in flight: before batch b is enqueued the host waits for the frame offsets
            // of batch b-2 and hands its streams to the copy engine (the GPU still has batch b-1 queued meanwhile),
            // so the copies overlap the rendering of the following batches and a ring slot is always drained long
            // before it comes round again.
            const int pr = (int)(gi % omfs_session::kPngRing);
            if (png && gi >= 2 && (rc = png_drain(s, (int)((gi - 2) % omfs_session::kPngRing)))) return rc;
            if (!dst_f && !dst_8) dst_f = img;  // nothing requested: still render (debug taps)
            if ((rc = mark(kStComposite))) return rc;
            // Pipelined: the persistent compositing warps leave 12 of the 32 warp slots per SM to the front end of
            // the next batch (measured on the 512^2 / 100k clip: 20 -> 32.0k frames/s, 32 -> 31.4k, 16 -> 30.5k).
            // The LAST batch of a call has no front end beside it (the call joins its streams at the end): it gets the
            // whole SM.
            const bool last_batch = (g0 + gT >= T) && (bi + 1 == sizes.size());
            if ((rc = composite_launch(S, N, W, H, P0, P1, P2, vals, ranges, s->cfg.bg, dst_f, dst_8, s->tickets.p,
                                       (pipelined && !(last_batch && s->last_batch_full && !deferred && !streaming)) ? s->comp_pipelined_warps : 0, cst)))
                return rc;
            if ((rc = mark(-1))) return rc;
            if (pipelined) {
                OMFS_CUDA(cudaEventRecord(s->ev_comp[ib], cst));
                s->ev_comp_pending[ib] = true;
            }
            s->last_image_buffer = ib;
            s->last_set = ib;
            s->last_batch_segments = S;
            if (out_on_host) {
                OMFS_CUDA(cudaEventRecord(s->ev_done[ib], cst));
                OMFS_CUDA(cudaStreamWaitEvent(s->copy_stream, s->ev_done[ib], 0));
                if (png) {
                    // encode on the sink's own stream: beside the next batch's compositing, after this batch's
                    // frames are complete and after the copy engine has drained the slot's previous contents
                    OMFS_CUDA(cudaStreamWaitEvent(s->png_stream, s->ev_done[ib], 0));
                    OMFS_CUDA(cudaStreamWaitEvent(s->png_stream, s->ev_png_copied[pr], 0));
                    if ((rc = png_encode_launch(S, W, H, img8, s->png_buf[pr].as<uint8_t>(), s->png_buf[pr].bytes,
                                                s->png_off[pr].as<unsigned long long>(), s->png_ws.p, s->png_ws_bytes,
                                                s->png_stream)))
                        return rc;
                    OMFS_CUDA(cudaMemcpyAsync(s->png_off_host[pr], s->png_off[pr].p,
                                              sizeof(unsigned long long) * ((size_t)S + 1), cudaMemcpyDeviceToHost,
                                              s->png_stream));
                    OMFS_CUDA(cudaEventRecord(s->ev_png_off[pr], s->png_stream));
                    s->png_slot[pr].active = true;
                    s->png_slot[pr].seg0 = seg0;
                    s->png_slot[pr].S = S;
                    s->png_slot[pr].sink = png;
                }
                if (out_u8)
                    OMFS_CUDA(cudaMemcpyAsync(out_u8 + seg0 * 3 * hw, img8, (size_t)S * 3 * hw,
                                              cudaMemcpyDeviceToHost, s->copy_stream));
                if (out_f32)
                    OMFS_CUDA(cudaMemcpyAsync(out_f32 + seg0 * 3 * hw, img, sizeof(float) * S * 3 * hw,
                                              cudaMemcpyDeviceToHost, s->copy_stream));
                // ev_copied = "this buffer set's images may be overwritten": after the raw copies and, with the
                // sink, after the encoder has read the uint8 frames.  Sink alone: recorded on the sink's stream, so
                // that the copy stream never queues a later batch's stream copy behind this batch's encode.
                if (png && !out_u8 && !out_f32) {
                    OMFS_CUDA(cudaEventRecord(s->ev_copied[ib], s->png_stream));
                    s->copied_pending[ib] = true;
                } else {
                    if (png) OMFS_CUDA(cudaStreamWaitEvent(s->copy_stream, s->ev_png_off[pr], 0));
                    OMFS_CUDA(cudaEventRecord(s->ev_copied[ib], s->copy_stream));
                    s->copied_pending[ib] = true;
                }
            }
            batch_index++;
        }
    }
    // the sink's remaining batches, oldest first
    if (png && !streaming)
        for (int k = 0; k < omfs_session::kPngRing; k++)
            if ((rc = png_drain(s, (batch_index + k) % omfs_session::kPngRing))) return rc;
    // everything this call launched is complete when the caller's stream is: join the compositing stream — unless
    // the caller asked to join by itself (omfs_session_join), so that consecutive calls overlap like the batches
    // of one call do
    if (streaming) {
        s->stream_base = idx0 + batch_index;
    } else if (deferred) {
        s->set_parity = (int)((idx0 + batch_index) & 1);
    } else {
        for (int ib = 0; ib < 2; ib++)
            if (s->ev_comp_pending[ib]) {
                OMFS_CUDA(cudaStreamWaitEvent(st, s->ev_comp[ib], 0));
                s->ev_comp_pending[ib] = false;
            }
    }
    s->stats[1] = g_launches - launches0;
    s->stats[2] = (uint64_t)batch_index;
    return OMFS_OK;
}

static int resolve_profile(omfs_session* s) {
    const size_t n = s->prof_stage.size();
    for (size_t i = 0; i + 1 < n; i++) {
        const int stage = s->prof_stage[i];
        if (stage < 0) continue;
        float ms = 0.f;
        OMFS_CUDA(cudaEventElapsedTime(&ms, s->prof_pool[i], s->prof_pool[i + 1]));
        s->stage_ms[stage] += ms;
        s->stage_calls[stage]++;
    }
    s->prof_stage.clear();
    return OMFS_OK;
}

static int finish_stats(omfs_session* s) {
    if (s->profiling) {
        OMFS_CUDA(cudaDeviceSynchronize());  // the caller may have rendered on its own stream
        int prc = resolve_profile(s);
        if (prc) return prc;
    }
    uint32_t h[8] = {0};
    OMFS_CUDA(cudaMemcpy(h, s->counters.p, sizeof(h), cudaMemcpyDeviceToHost));
    unsigned long long total = 0;
    memcpy(&total, &h[4], 8);
    s->stats[0] = total;  // tile pairs over the whole call
    s->stats[3] = h[1];
    s->last_pairs = h[0];
    s->max_batch_pairs = h[2];
    if (h[1]) {
        set_error("tile-pair list overflowed the session capacity (%zu; the largest batch needs %u): raise "
                  "pair_capacity or call omfs_session_reserve_pairs",
                  s->capacity, h[2]);
        return OMFS_ERR_CAPACITY;
    }
    return OMFS_OK;
}

// Grow (never shrink) the per-batch tile-pair capacity: the sorted index lists and the binning workspace.
extern "C" int omfs_session_reserve_pairs(omfs_session* s, uint64_t capacity) {
    OMFS_REQUIRE(s, "null argument");
    OMFS_REQUIRE(capacity > 0 && capacity < (1ull << 30), "capacity must be in (0, 2^30) pairs per batch");
    OMFS_CUDA(cudaSetDevice(s->cfg.device));
    if ((size_t)capacity <= s->capacity) return OMFS_OK;
    OMFS_CUDA(cudaDeviceSynchronize());
    int rc;
    for (int i = 0; i < 2; i++)
        if ((rc = s->vals[i].ensure(sizeof(uint32_t) * (size_t)capacity))) return rc;
    if (s->cfg.debug_keys && (rc = s->keys64.ensure(sizeof(uint64_t) * (size_t)capacity))) return rc;
    const size_t ws = omfs_binning_workspace_bytes(s->cfg.max_batch, s->N, s->cfg.width, s->cfg.height, (size_t)capacity);
    if ((rc = s->ws.ensure(ws))) return rc;
    s->ws_bytes = ws;
    s->capacity = (size_t)capacity;
    return OMFS_OK;
}

extern "C" int omfs_session_render_host(omfs_session* s, const omfs_frames_desc* fr, uint8_t* h_out_u8,
                                        float* h_out_f32) {
    OMFS_REQUIRE(s && fr, "null argument");
    OMFS_REQUIRE(s->subject_set, "omfs_session_set_subject must be called first");
    OMFS_REQUIRE(fr->n_frames >= 0 && fr->n_views > 0 && fr->n_views <= s->cfg.max_batch, "bad frame/view counts");
    OMFS_REQUIRE(fr->expr && fr->rotation && fr->neck_pose && fr->jaw_pose && fr->eyes_pose && fr->translation &&
                     fr->cams,
                 "null frame array");
    OMFS_REQUIRE(s->n_pending == 0, "streaming calls are outstanding: collect them first (omfs_session_collect_host_png)");
    OMFS_CUDA(cudaSetDevice(s->cfg.device));
    const int T = fr->n_frames;
    if (T == 0) return OMFS_OK;
    cudaStream_t st = s->stream;
    int rc;
    if ((rc = upload(s->expr, fr->expr, sizeof(float) * (size_t)T * s->n_expr, st))) return rc;
    if ((rc = upload(s->rotation, fr->rotation, sizeof(float) * 3 * T, st))) return rc;
    if ((rc = upload(s->neck, fr->neck_pose, sizeof(float) * 3 * T, st))) return rc;
    if ((rc = upload(s->jaw, fr->jaw_pose, sizeof(float) * 3 * T, st))) return rc;
    if ((rc = upload(s->eyes, fr->eyes_pose, sizeof(float) * 6 * T, st))) return rc;
    if ((rc = upload(s->transl, fr->translation, sizeof(float) * 3 * T, st))) return rc;
    if (fr->dynamic_offset &&
        (rc = upload(s->dyn, fr->dynamic_offset, sizeof(float) * 3 * (size_t)s->V * T, st)))
        return rc;
    if ((rc = upload(s->cams_in, fr->cams, sizeof(float) * kCam * fr->n_views, st))) return rc;
    // OMFS_TRACE=1: host-side timeline of the call (microseconds since entry) on stderr
    static const bool trace = getenv("OMFS_TRACE") != nullptr;
    const auto t0 = std::chrono::steady_clock::now();
    auto us = [&]() { return std::chrono::duration<double, std::micro>(std::chrono::steady_clock::now() - t0).count(); };
    for (int attempt = 0;; attempt++) {
        rc = render_core(s, T, fr->n_views, s->expr.as<float>(), s->rotation.as<float>(), s->neck.as<float>(),
                         s->jaw.as<float>(), s->eyes.as<float>(), s->transl.as<float>(),
                         fr->dynamic_offset ? s->dyn.as<float>() : nullptr, s->cams_in.as<float>(), h_out_u8,
                         h_out_f32, true, st);
        const double t_enq = us();
        cudaError_t e1 = cudaStreamSynchronize(st);
        const double t_st = us();
        cudaError_t e2 = cudaStreamSynchronize(s->copy_stream);
        const double t_cp = us();
        if (rc == OMFS_OK && e1 != cudaSuccess) rc = cuda_fail(e1, "stream sync", __FILE__, __LINE__);
        if (rc == OMFS_OK && e2 != cudaSuccess) rc = cuda_fail(e2, "copy stream sync", __FILE__, __LINE__);
        if (rc == OMFS_OK) rc = finish_stats(s);
        if (trace)
            fprintf(stderr, "[omfs trace] render_host T=%d: enqueued %.0f us, kernels done %.0f us, copies done %.0f us, "
                            "stats %.0f us\n", T, t_enq, t_st, t_cp, us());
        // A session created with pair_capacity = 0 sizes itself: an overflowing batch emitted nothing, the
        // scan still counted what it needs, so grow once to that (plus headroom) and render the call again.
        if (rc == OMFS_ERR_CAPACITY && s->capacity_auto && attempt == 0 && s->max_batch_pairs > 0) {
            uint64_t want = (uint64_t)s->max_batch_pairs + (uint64_t)s->max_batch_pairs / 8 + 4096;
            if (want >= (1ull << 30)) want = (1ull << 30) - 1;
            if (want > s->capacity && omfs_session_reserve_pairs(s, want) == OMFS_OK) continue;
        }
        break;
    }
    return rc;
}

// Host in, PNG streams out: the frames are encoded on the device (png.cu) and only the compressed streams (plus,
// optionally, the raw uint8 frames) cross PCIe.
extern "C" int omfs_session_render_host_png(omfs_session* s, const omfs_frames_desc* fr, uint8_t* h_png,
                                            size_t h_png_capacity, uint64_t* h_offsets, uint8_t* h_out_u8) {
    OMFS_REQUIRE(s && fr && h_png && h_offsets, "null argument");
    OMFS_REQUIRE(s->subject_set, "omfs_session_set_subject must be called first");
    OMFS_REQUIRE(fr->n_frames >= 0 && fr->n_views > 0 && fr->n_views <= s->cfg.max_batch, "bad frame/view counts");
    OMFS_REQUIRE(fr->expr && fr->rotation && fr->neck_pose && fr->jaw_pose && fr->eyes_pose && fr->translation &&
                     fr->cams,
                 "null frame array");
    OMFS_REQUIRE(s->n_pending == 0, "streaming calls are outstanding: collect them first (omfs_session_collect_host_png)");
    OMFS_CUDA(cudaSetDevice(s->cfg.device));
    const int T = fr->n_frames;
    h_offsets[0] = 0;
    if (T == 0) return OMFS_OK;
    cudaStream_t st = s->stream;
    int rc;
    if ((rc = upload(s->expr, fr->expr, sizeof(float) * (size_t)T * s->n_expr, st))) return rc;
    if ((rc = upload(s->rotation, fr->rotation, sizeof(float) * 3 * T, st))) return rc;
    if ((rc = upload(s->neck, fr->neck_pose, sizeof(float) * 3 * T, st))) return rc;
    if ((rc = upload(s->jaw, fr->jaw_pose, sizeof(float) * 3 * T, st))) return rc;
    if ((rc = upload(s->eyes, fr->eyes_pose, sizeof(float) * 6 * T, st))) return rc;
    if ((rc = upload(s->transl, fr->translation, sizeof(float) * 3 * T, st))) return rc;
    if (fr->dynamic_offset &&
        (rc = upload(s->dyn, fr->dynamic_offset, sizeof(float) * 3 * (size_t)s->V * T, st)))
        return rc;
    if ((rc = upload(s->cams_in, fr->cams, sizeof(float) * kCam * fr->n_views, st))) return rc;
    static const bool trace = getenv("OMFS_TRACE") != nullptr;
    const auto t0 = std::chrono::steady_clock::now();
    auto us = [&]() { return std::chrono::duration<double, std::micro>(std::chrono::steady_clock::now() - t0).count(); };
    for (int attempt = 0;; attempt++) {
        PngSink sink;
        sink.h_png = h_png;
        sink.capacity = h_png_capacity;
        sink.h_offsets = h_offsets;
        rc = render_core(s, T, fr->n_views, s->expr.as<float>(), s->rotation.as<float>(), s->neck.as<float>(),
                         s->jaw.as<float>(), s->eyes.as<float>(), s->transl.as<float>(),
                         fr->dynamic_offset ? s->dyn.as<float>() : nullptr, s->cams_in.as<float>(), h_out_u8, nullptr,
                         true, st, &sink);
        const double t_enq = us();
        cudaError_t e1 = cudaStreamSynchronize(st);
        const double t_st = us();
        cudaError_t e3 = s->png_stream ? cudaStreamSynchronize(s->png_stream) : cudaSuccess;
        cudaError_t e2 = cudaStreamSynchronize(s->copy_stream);
        if (trace)
            fprintf(stderr, "[omfs trace] render_host_png T=%d: enqueued+drained %.0f us, kernels done %.0f us, sink and "
                            "copies done %.0f us, %zu bytes\n", T, t_enq, t_st, us(), sink.written);
        if (rc == OMFS_OK && e1 != cudaSuccess) rc = cuda_fail(e1, "stream sync", __FILE__, __LINE__);
        if (rc == OMFS_OK && e3 != cudaSuccess) rc = cuda_fail(e3, "sink stream sync", __FILE__, __LINE__);
        if (rc == OMFS_OK && e2 != cudaSuccess) rc = cuda_fail(e2, "copy stream sync", __FILE__, __LINE__);
        if (rc != OMFS_OK) {   // a failed call must not leave events of a half-run pipeline pending
            cudaDeviceSynchronize();
            for (auto& slot : s->png_slot) slot.active = false;
        }
        if (rc == OMFS_OK) rc = finish_stats(s);
        if (rc == OMFS_ERR_CAPACITY && s->stats[3] && s->capacity_auto && attempt == 0 && s->max_batch_pairs > 0) {
            uint64_t want = (uint64_t)s->max_batch_pairs + (uint64_t)s->max_batch_pairs / 8 + 4096;
            if (want >= (1ull << 30)) want = (1ull << 30) - 1;
            if (want > s->capacity && omfs_session_reserve_pairs(s, want) == OMFS_OK) continue;
        }
        break;
    }
    return rc;
}

// ---- streaming form of the call above.  submit enqueues a clip and returns; collect completes the OLDEST submitted
// clip: its PNG streams and offsets are then in the caller's buffers.  Up to kMaxPending clips may be outstanding, so
// clip i+1 renders while clip i's last batches are still being encoded and copied — the overlap a blocking call
// cannot have.  Everything the call reads and writes on the host (parameter arrays, h_png, h_offsets) must stay
// valid until its collect.
extern "C" int omfs_session_submit_host_png(omfs_session* s, const omfs_frames_desc* fr, uint8_t* h_png,
                                            size_t h_png_capacity, uint64_t* h_offsets) {
    OMFS_REQUIRE(s && fr && h_png && h_offsets, "null argument");
    OMFS_REQUIRE(s->subject_set, "omfs_session_set_subject must be called first");
    OMFS_REQUIRE(fr->n_frames > 0 && fr->n_views > 0 && fr->n_views <= s->cfg.max_batch, "bad frame/view counts");
    OMFS_REQUIRE(fr->expr && fr->rotation && fr->neck_pose && fr->jaw_pose && fr->eyes_pose && fr->translation &&
                     fr->cams,
                 "null frame array");
    OMFS_REQUIRE(s->n_pending < omfs_session::kMaxPending, "too many clips outstanding: collect one first");
    OMFS_REQUIRE(!s->profiling, "stage profiling needs the blocking calls");
    OMFS_CUDA(cudaSetDevice(s->cfg.device));
    const int T = fr->n_frames;
    cudaStream_t st = s->stream;
    int rc;
    if ((rc = upload(s->expr, fr->expr, sizeof(float) * (size_t)T * s->n_expr, st))) return rc;
    if ((rc = upload(s->rotation, fr->rotation, sizeof(float) * 3 * T, st))) return rc;
    if ((rc = upload(s->neck, fr->neck_pose, sizeof(float) * 3 * T, st))) return rc;
    if ((rc = upload(s->jaw, fr->jaw_pose, sizeof(float) * 3 * T, st))) return rc;
    if ((rc = upload(s->eyes, fr->eyes_pose, sizeof(float) * 6 * T, st))) return rc;
    if ((rc = upload(s->transl, fr->translation, sizeof(float) * 3 * T, st))) return rc;
    if (fr->dynamic_offset &&
        (rc = upload(s->dyn, fr->dynamic_offset, sizeof(float) * 3 * (size_t)s->V * T, st)))
        return rc;
    if ((rc = upload(s->cams_in, fr->cams, sizeof(float) * kCam * fr->n_views, st))) return rc;
    PngSink* sink = new PngSink();
    sink->h_png = h_png;
    sink->capacity = h_png_capacity;
    sink->h_offsets = h_offsets;
    sink->streaming = true;
    h_offsets[0] = 0;
    s->pending[s->n_pending++] = sink;
    rc = render_core(s, T, fr->n_views, s->expr.as<float>(), s->rotation.as<float>(), s->neck.as<float>(),
                     s->jaw.as<float>(), s->eyes.as<float>(), s->transl.as<float>(),
                     fr->dynamic_offset ? s->dyn.as<float>() : nullptr, s->cams_in.as<float>(), nullptr, nullptr, true, st,
                     sink, true);
    if (rc != OMFS_OK) sink->error = rc;   // reported again by the collect, which also cleans up
    return rc;
}

extern "C" int omfs_session_collect_host_png(omfs_session* s) {
    OMFS_REQUIRE(s, "null argument");
    OMFS_REQUIRE(s->n_pending > 0, "no clip outstanding");
    OMFS_CUDA(cudaSetDevice(s->cfg.device));
    PngSink* sink = s->pending[0];
    // drain what is left of this clip, oldest batch first (batches of a later clip are not touched: their slots
    // carry their own sink), then wait for its last copy
    int rc = OMFS_OK;
    for (int k = 0; k < omfs_session::kPngRing && rc == OMFS_OK; k++) {
        const int r = (int)((s->stream_base + k) % omfs_session::kPngRing);   // oldest slot first
        if (s->png_slot[r].active && s->png_slot[r].sink == sink) rc = png_drain(s, r);
    }
    if (!s->ev_collect) OMFS_CUDA(cudaEventCreateWithFlags(&s->ev_collect, cudaEventDisableTiming));
    cudaError_t e = cudaEventRecord(s->ev_collect, s->copy_stream);
    if (e == cudaSuccess) e = cudaEventSynchronize(s->ev_collect);
    if (rc == OMFS_OK && e != cudaSuccess) rc = cuda_fail(e, "collect", __FILE__, __LINE__);
    if (rc == OMFS_OK) rc = sink->error;
    if (rc == OMFS_OK) {   // tile-pair overflow of any batch so far (sticky until the stream of calls is reset)
        uint32_t h[2] = {0, 0};
        e = cudaMemcpyAsync(h, s->counters.p, sizeof(h), cudaMemcpyDeviceToHost, s->copy_stream);
        if (e == cudaSuccess) e = cudaStreamSynchronize(s->copy_stream);
        if (e != cudaSuccess) rc = cuda_fail(e, "collect (flags)", __FILE__, __LINE__);
        else if (h[1]) {
            set_error("tile-pair list overflowed the session capacity (%zu): raise pair_capacity or call "
                      "omfs_session_reserve_pairs, then submit the clip again", s->capacity);
            rc = OMFS_ERR_CAPACITY;
        }
    }
    delete sink;
    for (int i = 1; i < s->n_pending; i++) s->pending[i - 1] = s->pending[i];
    s->pending[--s->n_pending] = nullptr;
    if (rc != OMFS_OK) {   // a failed stream of calls is wound up: nothing half-run stays pending
        cudaDeviceSynchronize();
        for (auto& slot : s->png_slot) slot.active = false;
        while (s->n_pending > 0) delete s->pending[--s->n_pending];
        for (int i = 0; i < omfs_session::kMaxPending; i++) s->pending[i] = nullptr;
        cudaMemset(s->counters.p, 0, 256);
        s->stream_base = 0;
        s->copied_pending[0] = s->copied_pending[1] = false;
    }
    return rc;
}

extern "C" int omfs_session_render_device(omfs_session* s, const omfs_frames_desc* fr, uint8_t* d_out_u8,
                                          float* d_out_f32, void* stream) {
    OMFS_REQUIRE(s && fr, "null argument");
    OMFS_REQUIRE(s->subject_set, "omfs_session_set_subject must be called first");
    OMFS_REQUIRE(fr->n_frames >= 0 && fr->n_views > 0 && fr->n_views <= s->cfg.max_batch, "bad frame/view counts");
    OMFS_REQUIRE(fr->expr && fr->rotation && fr->neck_pose && fr->jaw_pose && fr->eyes_pose && fr->translation &&
                     fr->cams,
                 "null frame array");
    OMFS_REQUIRE(s->n_pending == 0, "streaming calls are outstanding: collect them first (omfs_session_collect_host_png)");
    OMFS_CUDA(cudaSetDevice(s->cfg.device));
    if (fr->n_frames == 0) return OMFS_OK;
    cudaStream_t st = (cudaStream_t)stream;  // NULL = the legacy default stream, as everywhere in this ABI
    s->user_stream = st;
    return render_core(s, fr->n_frames, fr->n_views, fr->expr, fr->rotation, fr->neck_pose, fr->jaw_pose,
                       fr->eyes_pose, fr->translation, fr->dynamic_offset, fr->cams, d_out_u8, d_out_f32, false, st);
}

// Waits for the session's streams and reports overflow; call after render_device before reading results.
extern "C" int omfs_session_sync(omfs_session* s) {
    OMFS_REQUIRE(s, "null argument");
    OMFS_CUDA(cudaSetDevice(s->cfg.device));
    OMFS_CUDA(cudaStreamSynchronize(s->user_stream));
    OMFS_CUDA(cudaStreamSynchronize(s->stream));
    OMFS_CUDA(cudaStreamSynchronize(s->copy_stream));
    // with the deferred join the compositing stream is not joined into the caller's: wait for it here
    if (s->comp_stream) OMFS_CUDA(cudaStreamSynchronize(s->comp_stream));
    if (s->png_stream) OMFS_CUDA(cudaStreamSynchronize(s->png_stream));
    if (s->defer_join)
        for (int ib = 0; ib < 2; ib++) s->ev_comp_pending[ib] = false;
    return finish_stats(s);
}

extern "C" void* omfs_session_stream(omfs_session* s) { return s ? (void*)s->stream : nullptr; }

extern "C" int omfs_session_set_deferred_join(omfs_session* s, int on) {
    OMFS_REQUIRE(s, "null argument");
    OMFS_CUDA(cudaSetDevice(s->cfg.device));
    if (!on && s->defer_join) {   // leaving the mode: whatever is still compositing is joined on the last caller stream
        for (int ib = 0; ib < 2; ib++)
            if (s->ev_comp_pending[ib]) {
                OMFS_CUDA(cudaStreamWaitEvent(s->user_stream, s->ev_comp[ib], 0));
                s->ev_comp_pending[ib] = false;
            }
        s->set_parity = 0;
    }
    s->defer_join = on != 0;
    return OMFS_OK;
}

extern "C" int omfs_session_join(omfs_session* s, void* stream) {
    OMFS_REQUIRE(s, "null argument");
    OMFS_CUDA(cudaSetDevice(s->cfg.device));
    // the flags stay set: the rendering stream itself has not waited, and the next call must before it reuses a set
    for (int ib = 0; ib < 2; ib++)
        if (s->ev_comp_pending[ib]) OMFS_CUDA(cudaStreamWaitEvent((cudaStream_t)stream, s->ev_comp[ib], 0));
    return OMFS_OK;
}

extern "C" int omfs_session_stats(omfs_session* s, uint64_t* out4) {
    OMFS_REQUIRE(s && out4, "null argument");
    for (int i = 0; i < 4; i++) out4[i] = s->stats[i];
    return OMFS_OK;
}

extern "C" int omfs_session_tap(omfs_session* s, const char* name, void** d_ptr, size_t* bytes) {
    OMFS_REQUIRE(s && name && d_ptr && bytes, "null argument");
    struct Tap {
        const char* n;
        DevBuf* b;
    };
    DevBuf* sorted_k = &s->keys64;
    DevBuf* sorted_v = &s->vals[s->last_set];
    Tap taps[] = {{"verts", &s->verts}, {"ff", &s->ff}, {"P0", &s->P0[s->last_set]}, {"P1", &s->P1[s->last_set]}, {"P2", &s->P2[s->last_set]},
                  {"tiles_touched", &s->tt}, {"depth_keys", &s->depth_keys}, {"keys", sorted_k}, {"vals", sorted_v},
                  {"ranges", &s->ranges[s->last_set]}, {"image", &s->image[s->last_image_buffer]},
                  {"image_u8", &s->image_u8[s->last_image_buffer]}, {"vp", &s->vp}, {"acoef", &s->acoef},
                  {"base", &s->base}, {"counters", &s->counters}, {"rmats", &s->rmats}};
    for (const Tap& t : taps)
        if (strcmp(t.n, name) == 0) {
            if (!t.b->p) {
                set_error("tap '%s' is not materialised (create the session with debug_keys=1)", name);
                return OMFS_ERR_INVALID;
            }
            *d_ptr = t.b->p;
            *bytes = t.b->bytes;
            return OMFS_OK;
        }
    set_error("unknown tap '%s'", name);
    return OMFS_ERR_INVALID;
}

// Per-stage device time, accumulated while profiling is on (it costs a host sync per batch, so the
// headline throughput is measured with profiling off).  out_ms/out_calls: flame, face_frames,
// bind_preprocess, depth_sort, tile_ranges, emit_scatter, composite.
extern "C" int omfs_session_set_profiling(omfs_session* s, int on) {
    OMFS_REQUIRE(s, "null argument");
    s->profiling = on != 0;
    for (int i = 0; i < 8; i++) {
        s->stage_ms[i] = 0.0;
        s->stage_calls[i] = 0;
    }
    return OMFS_OK;
}
extern "C" int omfs_session_stage_ms(omfs_session* s, double* out_ms8, uint64_t* out_calls8) {
    OMFS_REQUIRE(s && out_ms8 && out_calls8, "null argument");
    for (int i = 0; i < 8; i++) {
        out_ms8[i] = s->stage_ms[i];
        out_calls8[i] = s->stage_calls[i];
    }
    return OMFS_OK;
}

extern "C" int omfs_session_dims(omfs_session* s, int32_t* out9) {
    int32_t* out8 = out9;
    OMFS_REQUIRE(s && out8, "null argument");
    out8[0] = s->V; out8[1] = s->F; out8[2] = s->n_expr; out8[3] = s->N;
    out8[4] = s->kpad; out8[5] = s->npad; out8[6] = s->tiles; out8[7] = s->last_batch_segments;
    out8[8] = (int32_t)s->last_pairs;
    return OMFS_OK;
}
