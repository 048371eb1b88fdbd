/*
 * omfs_b200.h — C-ABI of the B200-native surgery-render hot path.
 *
 * The reference (cwlachap/OMFS-4D-Video-Gen) has no FFI / operator table for this path: it
 * reaches its renderer by spawning a process
 * (02_Visual_Engine/render_surgery.py:289-315, `render_with_gaussians`).  The entry points
 * below are what an in-process binding at that replacement point needs; each comment names
 * the reference interface (file:line) or the un-vendored upstream unit
 * it stands in for.  INTEGRATION.md shows the ctypes stub a maintainer would add.
 *
 * Conventions
 *   - plain pointers and sizes only; no torch / C++ types cross the boundary;
 *   - every function returns 0 on success, a negative omfs_status otherwise, and never
 *     throws; omfs_last_error() gives the message of the last failure on this thread;
 *   - "d_" arguments are device pointers, "h_" host pointers; stream is a cudaStream_t
 *     passed as void* (NULL = default stream);
 *   - level-1 functions (kernels) never allocate: scratch comes from the caller
 *     (omfs_binning_workspace_bytes); level-2 (omfs_session_*) owns its device buffers
 *     between create and destroy;
 *   - all float data is IEEE binary32, images are [S,3,H,W] planar float or [S,H,W,3] uint8.
 */
#ifndef OMFS_B200_H
#define OMFS_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define OMFS_ABI_VERSION 2
#define OMFS_TILE 16          /* compositing tile edge, pixels */
#define OMFS_FF_STRIDE 20     /* floats per face-frame record */
#define OMFS_CAM_FLOATS 40    /* floats per camera record (cameras.py: Camera.pack) */
#define OMFS_N_JOINTS 5
#define OMFS_N_POSE_FEAT 36

typedef enum omfs_status {
    OMFS_OK = 0,
    OMFS_ERR_INVALID = -1,     /* bad argument */
    OMFS_ERR_CUDA = -2,        /* a CUDA call failed; see omfs_last_error() */
    OMFS_ERR_CAPACITY = -3,    /* tile-pair list overflowed the caller's capacity */
    OMFS_ERR_UNSUPPORTED = -4, /* no sm_100 device / feature missing */
    OMFS_ERR_NOMEM = -5
} omfs_status;

const char* omfs_last_error(void);
int omfs_abi_version(void);
/* 0 when device `dev` is an sm_100 part this library can run on. */
int omfs_device_check(int dev);

/* ------------------------------------------------------------------------------------------
 * Level 1 — one call per pipeline stage, device pointers in and out.
 * ---------------------------------------------------------------------------------------- */

/* U1+U2 operand prep (the in-tree anchor is SimpleFLAME.forward,
 * 02_Visual_Engine/flame_fitter.py:154-175).  Per frame: Rodrigues of the 5 joint rotations
 * -> d_rmats[T,5,9]; coefficient row [expr | pose features | 0] split into tf32 hi/lo and laid
 * out as the GEMM's A operand d_acoef[T,3*kpad] = [hi | hi | lo].  kpad = round_up(n_expr+36, 8). */
int omfs_flame_pose_prep(int T, int n_expr, int kpad,
                         const float* d_expr, const float* d_rotation, const float* d_neck,
                         const float* d_jaw, const float* d_eyes,
                         float* d_acoef, float* d_rmats, void* stream);

/* U1+U2 contraction: d_vp[T,npad] = d_base[npad] + A[T,3*kpad] . Bt[npad,3*kpad]^T.
 * Columns 0..3V-1 are posed-template vertices before skinning, 3V..3V+14 the 5 joints.
 * impl 0 = tcgen05/TMA tensor-core kernels (tf32x3: A = [Ah | Ah | Al], Bt = [Bh | Bl | Bh]; below 1024 rows the
 * kernel that streams the concatenated operands, from there on the one that stages the four panels once per K
 * block and issues the three products from the same tiles), 2 / 3 = force the first / the second,
 * 1 = fp32 CUDA-core kernel (same operands; kept as the cross-check). */
int omfs_flame_blend_gemm(int T, int kpad, int npad,
                          const float* d_acoef, const float* d_bt, const float* d_base,
                          float* d_vp, int impl, void* stream);

/* U3: linear-blend skinning.  d_verts[T,V,3] = sum_j W[v,j] A_j(t) [v_posed;1] + transl[t];
 * the kinematic chain A_j(t) is rebuilt per block from d_rmats and the joint columns of d_vp.
 * d_dyn (optional, [T,V,3]) is the per-frame dynamic offset; d_jdyn (optional, [T,15]) its
 * joint contribution. */
int omfs_flame_lbs(int T, int V, int npad,
                   const float* d_vp, const float* d_rmats, const float* d_weights /*[V,5]*/,
                   const float* d_transl /*[T,3]*/, const float* d_dyn, const float* d_jdyn,
                   float* d_verts, void* stream);

/* J_reg . dyn[t] -> d_jdyn[T,15] (only needed when dynamic offsets are non-zero). */
int omfs_flame_joint_dyn(int T, int V, const float* d_jreg /*[5,V]*/, const float* d_dyn, float* d_jdyn,
                         void* stream);

/* U4: per-face centre, orthonormal frame, isotropic scale, unit quaternion.
 * d_ff[T,F,20] = [cx cy cz s | qw qx qy qz | R row0,0 | R row1,0 | R row2,0]. */
int omfs_face_frames(int T, int V, int F, const float* d_verts, const int32_t* d_faces, float* d_ff,
                     void* stream);

/* U5+U6 fused: parent-triangle transform of every Gaussian, then cull / project / EWA / SH.
 * One segment = one (frame, camera) pair; d_seg_frame[S] gives the frame of each segment and
 * d_cams[S,40] its camera.  Outputs are [S,N,4] float4 streams plus d_tiles_touched[S,N]:
 *   P0 = (px, py, cull extents, radius as int32 bits)  P1 = (ca, cb, cc, lo)  P2 = (r, g, b, depth)
 * (the published fields; the cull extents — two half-precision half-widths of the alpha >= 1/255 footprint box —
 * are an acceleration hint outside the parity surface, placed beside the centre so that binning and compositing
 * read one record for both)
 * and (optional) d_depth_keys[S,N] = the depth's float bits, 0 for culled Gaussians: the key of the
 * depth sort in omfs_binning. */
int omfs_bind_preprocess(int S, int N, int F, int width, int height,
                         const float* d_ff, const int32_t* d_seg_frame, const float* d_cams,
                         const float* d_xyzb, const float* d_scale_lo, const float* d_rot, const float* d_sh,
                         float* d_P0, float* d_P1, float* d_P2, uint32_t* d_tiles_touched,
                         uint32_t* d_depth_keys, void* stream);

/* U7+U8+U9.  Result: the pair list of the batch ordered by the published key
 * ((seg*tiles+tile)<<32 | depth bits, ties by Gaussian index): d_sorted_vals[capacity] (Gaussian index
 * inside its segment in the low OMFS_VAL_INDEX_BITS bits; the top four bits are block hints for omfs_composite:
 * bit 2*yhalf + xhalf is set when the Gaussian's footprint box reaches that 8x8 pixel block of the pair's tile)
 * and d_ranges[S*tiles,2] (untouched tiles stay (0,0)).  N <= 2^28 - 1.  Internally (binning.cu):
 * segmented onesweep depth sort of the Gaussians, per-tile counts and their scan, then one fused
 * emit + counting-sort-by-tile kernel that writes every index straight to its final position.
 * d_sorted_keys[capacity] (optional) receives the 64-bit keys in final order, for parity / debugging.
 * d_num_pairs (uint32 on device) receives the pair count; if it exceeds capacity nothing is emitted
 * and d_status_flag (int on device) is set to 1.  All scratch comes from d_workspace. */
#define OMFS_VAL_INDEX_BITS 28
size_t omfs_binning_workspace_bytes(int S, int N, int width, int height, size_t capacity);
int omfs_binning(int S, int N, int width, int height, size_t capacity,
                 const float* d_P0, const uint32_t* d_depth_keys, const uint32_t* d_tiles_touched,
                 uint32_t* d_sorted_vals, uint64_t* d_sorted_keys, uint32_t* d_ranges /*[S*tiles,2]*/,
                 uint32_t* d_num_pairs, int* d_status_flag, void* d_workspace, size_t workspace_bytes,
                 void* stream);

/* U10: front-to-back alpha compositing, one warp per 8x8 pixel block; a warp only fetches the list entries whose
 * block hint (see omfs_binning) names its block.  d_image[S,3,H,W];
 * d_image_u8 (optional) [S,H,W,3] gets the save_image quantisation in the same kernel.
 * d_tickets (optional): OMFS_COMPOSITE_TICKET_BYTES of device memory, zero before its first use and not
 * shared by launches that may run concurrently; the kernel leaves it zeroed.  With it the launch is one
 * resident wave of persistent warps drawing work units from the counter; without it, one CTA per unit. */
#define OMFS_COMPOSITE_TICKET_BYTES 16
int omfs_composite(int S, int N, int width, int height,
                   const float* d_P0, const float* d_P1, const float* d_P2,
                   const uint32_t* d_sorted_vals, const uint32_t* d_ranges, const float* bg3,
                   float* d_image, uint8_t* d_image_u8, void* d_tickets, void* stream);

/* Frame sink on the device: S uint8 frames [S,H,W,3] in HBM -> S complete 8-bit RGB PNG files
 * packed back to back in d_png, frame i at d_png[d_offsets[i] .. d_offsets[i+1]) (d_offsets: S+1 uint64 on the
 * device).  Lossless; every decoder reads them (Up filter on every row, one deflate block per strip of rows —
 * dynamic Huffman with one of 8 fixed code tables chosen per strip by exact cost, run-length matches, or stored —
 * one IDAT chunk per strip with its CRC-32, Adler-32 combined over the strips; csrc/png_core.cuh describes the
 * stream).  png_capacity must be >= S * omfs_png_max_bytes(width, height); scratch comes from d_workspace
 * (omfs_png_workspace_bytes).  Width up to 5450 pixels (0 is returned by the size queries beyond that). */
size_t omfs_png_max_bytes(int width, int height);
size_t omfs_png_workspace_bytes(int S, int width, int height);
int omfs_png_encode(int S, int width, int height, const uint8_t* d_frames_u8, uint8_t* d_png, size_t png_capacity,
                    uint64_t* d_offsets, void* d_workspace, size_t workspace_bytes, void* stream);

/* R9 on device (02_Visual_Engine/validation_reporting.py:16-37): the moments behind PSNR and the global SSIM
 * of T pairs of uint8 frames [T,H,W,3] resident in HBM, one pass over both sets.
 *   d_moments[T][6] (float64) = sum (a-b)^2 over all channel values (exact), then with x, y the float32
 *   BT.601 luma of a, b: sum x, sum y, sum x^2, sum y^2, sum xy.
 * The host finishes MSE -> PSNR (99.0 when identical) and moments -> SSIM.  Frame sets must be 4-byte
 * aligned; H*W*3 must be a multiple of 4 when T > 1. */
int omfs_frame_metrics(int T, int height, int width, const uint8_t* d_a_u8, const uint8_t* d_b_u8,
                       double* d_moments, void* stream);

/* R5/R6 (01_Clinical_Engine/surgical_sim.py:25-47, 180-204, 262-329): half-space masks and the
 * rigid move of the two mobile segments, on an arbitrary point set, in float64.
 *   planes[3][8]  = {nx,ny,nz, ox,oy,oz, 0,0} for Le Fort, BSSO-L, BSSO-R (normals from
 *                   _angle_to_normal, evaluated on the host in float64);
 *   moves[2][12]  = row-major 3x3 rotation then translation, for maxilla and mandible
 *                   (rotation about the segment's bounding-box centre, as PyVista's
 *                   `mesh.center` is, surgical_sim.py:300,312);
 *   d_jaw_weight  optional [P]: points with weight > 0.5 belong to the mandible side
 *                   (NULL: `mandible_first` points by index belong to it);
 *   d_mask[P]     bit0 (p-o).n<=0 for Le Fort, bit1 (p-o).n>0 for BSSO-L, bit2 (p-o).n<=0 for
 *                 BSSO-R, bit3 = mobile maxilla, bit4 = distal mandible;
 *   d_out[P,3]    moved points (float32), d_bbox[2][6] the two segment bounding boxes. */
int omfs_displace_points(int P, const float* d_points, const double* h_planes, const double* h_moves,
                         const float* d_jaw_weight, int mandible_first,
                         uint8_t* d_mask, float* d_out, float* d_bbox /*[12]*/, void* stream);

/* ------------------------------------------------------------------------------------------
 * Level 2 — a render session: model + avatar resident in HBM, frames streamed in batches.
 * This is the call that replaces the body of render_with_gaussians
 * (02_Visual_Engine/render_surgery.py:245-362).
 * ---------------------------------------------------------------------------------------- */
typedef struct omfs_session omfs_session;

typedef struct omfs_model_desc {
    int32_t n_verts, n_faces, n_expr, n_gauss;
    const float* v_template;   /* [V,3] */
    const float* shapedirs;    /* [300+n_expr, 3V] row k = direction k */
    const float* posedirs;     /* [36, 3V] */
    const float* j_regressor;  /* [5,V] */
    const float* lbs_weights;  /* [V,5] */
    const int32_t* faces;      /* [F,3] */
    /* baked avatar (avatar.py: bake_avatar) */
    const float* xyzb;         /* [N,4] */
    const float* scale_lo;     /* [N,4] */
    const float* rot;          /* [N,4] */
    const float* sh;           /* [12,N,4] */
} omfs_model_desc;

typedef struct omfs_session_config {
    int32_t width, height;
    int32_t max_batch;          /* segments (frame x view) per launch group */
    int32_t device;
    int32_t gemm_impl;          /* 0 tensor core, 1 CUDA core */
    int32_t debug_keys;         /* also materialise the 64-bit sorted keys of each batch (parity taps) */
    uint64_t pair_capacity;     /* tile pairs per batch; 0 = start at 6 * max_batch * n_gauss and let
                                   omfs_session_render_host grow it when a batch overflows */
    float bg[3];
} omfs_session_config;

int omfs_session_create(const omfs_model_desc* model, const omfs_session_config* cfg, omfs_session** out);
void omfs_session_destroy(omfs_session* s);

/* Per-subject fold: base = v_template + shapedirs[:300].shape + static_offset (+ plan_offset),
 * joint base = J_reg . base.  Host pointers; plan_offset may be NULL. */
int omfs_session_set_subject(omfs_session* s, const float* h_shape300, const float* h_static_offset,
                             const float* h_plan_offset);

typedef struct omfs_frames_desc {
    int32_t n_frames;           /* T */
    int32_t n_views;            /* cameras per frame; segments = T * n_views, view-minor */
    const float* expr;          /* [T,n_expr] */
    const float* rotation;      /* [T,3] */
    const float* neck_pose;     /* [T,3] */
    const float* jaw_pose;      /* [T,3] */
    const float* eyes_pose;     /* [T,6] */
    const float* translation;   /* [T,3] */
    const float* dynamic_offset;/* [T,V,3] or NULL */
    const float* cams;          /* [n_views,40] */
} omfs_frames_desc;

/* Host in, host out: parameters are copied host->device, frames come back as uint8 [S,H,W,3]
 * (h_out_u8) and/or float [S,3,H,W] (h_out_f32); either may be NULL.  Pinned host memory
 * makes the copies asynchronous.  Blocks until the frames are in host memory. */
int omfs_session_render_host(omfs_session* s, const omfs_frames_desc* frames,
                             uint8_t* h_out_u8, float* h_out_f32);

/* Host in, PNG files out (replaces the frame sink behind
 * 02_Visual_Engine/render_surgery.py:289-315 — upstream's save_image per frame — and the PNG round trip of
 * stitch_video, :412-449).  The frames are filtered, deflated and framed ON THE DEVICE (omfs_png_encode), so only
 * the compressed streams cross PCIe.  h_png receives the S complete PNG files back to back, frame i being
 * h_png[h_offsets[i] .. h_offsets[i+1]); h_offsets has S+1 entries.  h_png_capacity >= S * omfs_png_max_bytes()
 * always suffices; a smaller buffer is accepted and OMFS_ERR_CAPACITY is returned if the streams do not fit.
 * h_out_u8 (optional) also returns the raw uint8 frames [S,H,W,3].  Blocks until everything is in host memory. */
int omfs_session_render_host_png(omfs_session* s, const omfs_frames_desc* frames, uint8_t* h_png,
                                 size_t h_png_capacity, uint64_t* h_offsets, uint8_t* h_out_u8);

/* The same call for a STREAM of clips (one after another, or one frame block per step): submit enqueues a clip and
 * returns without waiting; collect completes the OLDEST submitted clip — its PNG files and offsets are then in the
 * buffers given to its submit.  Up to three clips may be outstanding, so clip i+1 is uploaded and rendered while the last
 * batches of clip i are still being encoded and copied: the overlap a blocking call cannot have (the session's
 * buffers and ring slots simply continue from one call to the next).  Everything a clip reads or writes on the host
 * (parameter arrays, h_png, h_offsets) must stay valid and untouched until its collect; a tile-pair overflow or a
 * too small h_png is reported by the collect (OMFS_ERR_CAPACITY: every outstanding clip is dropped, reserve and
 * submit again).  The blocking calls refuse to run while clips are outstanding. */
int omfs_session_submit_host_png(omfs_session* s, const omfs_frames_desc* frames, uint8_t* h_png,
                                 size_t h_png_capacity, uint64_t* h_offsets);
int omfs_session_collect_host_png(omfs_session* s);

/* Device in, device out (inputs already resident; same field meaning, device pointers).
 * d_out_u8 / d_out_f32 may be NULL.  Asynchronous on `stream`. */
int omfs_session_render_device(omfs_session* s, const omfs_frames_desc* d_frames,
                               uint8_t* d_out_u8, float* d_out_f32, void* stream);

/* Streaming callers (one clip after another, or one frame block per step): with the deferred join on,
 * omfs_session_render_device returns without making `stream` wait for the call's last compositing launch, so the
 * next call's front end overlaps it exactly as the batches inside one call overlap.  The caller orders its own
 * consumers with omfs_session_join(s, consumer_stream): that stream then waits for every compositing launch
 * enqueued so far (it may be the rendering stream itself, or e.g. the stream that pushes the frames to a peer).
 * Host-output calls always join.  Off by default: then every call is complete in stream order, as documented. */
int omfs_session_set_deferred_join(omfs_session* s, int on);
int omfs_session_join(omfs_session* s, void* stream);

/* Counters of the last render call: [0] tile pairs (whole call), [1] kernel launches, [2] batches,
 * [3] overflow flag.  Valid after omfs_session_render_host / omfs_session_sync. */
int omfs_session_stats(omfs_session* s, uint64_t* out4);

/* Debug taps for the parity tests: pointers into the session's buffers for the LAST batch
 * rendered (valid until the next call).  Names: "verts" "ff" "P0" "P1" "P2" "tiles_touched"
 * "depth_keys" "keys" (needs debug_keys) "vals" "ranges" "image" "image_u8" "vp" "acoef" "base" "rmats". */
int omfs_session_tap(omfs_session* s, const char* name, void** d_ptr, size_t* bytes);

/* Waits for the session's streams; returns OMFS_ERR_CAPACITY if any batch overflowed. */
int omfs_session_sync(omfs_session* s);
/* Grow the per-batch tile-pair capacity (never shrinks).  omfs_session_render_device cannot re-run a
 * call by itself: on OMFS_ERR_CAPACITY from omfs_session_sync, reserve and render again. */
int omfs_session_reserve_pairs(omfs_session* s, uint64_t capacity);
void* omfs_session_stream(omfs_session* s);
/* out9 = V, F, n_expr, N, kpad, npad, tiles, segments of the last batch, tile pairs of the last batch */
int omfs_session_dims(omfs_session* s, int32_t* out9);
/* Per-stage device time (ms) and launch-group counts while profiling is on; stages: flame,
 * face_frames, bind_preprocess, depth_sort, tile_ranges, emit_scatter, composite, unused.  Profiling adds a host
 * sync per batch: headline numbers are taken with it off. */
int omfs_session_set_profiling(omfs_session* s, int on);
int omfs_session_stage_ms(omfs_session* s, double* out_ms8, uint64_t* out_calls8);

/* ------------------------------------------------------------------------------------------
 * Helpers for callers without a CUDA runtime of their own (ctypes, cgo, JNI ...).
 * ---------------------------------------------------------------------------------------- */
int omfs_device_count(void);                    /* CUDA devices visible to this process (0 if none) */
int omfs_set_device(int dev);                  /* device the helpers below allocate on (one process per GPU) */
int omfs_host_alloc(void** p, size_t bytes);   /* page-locked host memory */
int omfs_host_free(void* p);
int omfs_device_alloc(void** p, size_t bytes);
int omfs_device_free(void* p);
int omfs_memcpy_h2d(void* d_dst, const void* h_src, size_t bytes);
int omfs_memcpy_d2h(void* h_dst, const void* d_src, size_t bytes); /* synchronises the device first */
int omfs_device_memset(void* d_dst, int value, size_t bytes);
int omfs_device_sync(void);
unsigned long long omfs_launch_count(void);    /* kernels launched by this library so far */

/* ------------------------------------------------------------------------------------------
 * Frame exchange between the GPUs of one box (the path's only exchange step; the
 * reference is single-GPU, 02_Visual_Engine/app.py:194-196).  One process per GPU.  The root
 * allocates its receive buffer with omfs_device_alloc and exports it; every other rank opens the
 * handle once and pushes its finished frames with omfs_push_frames: a peer-to-peer copy over
 * NVLink executed by the SENDER's copy engine on the given stream — no SM is taken from the
 * rendering kernels on either side.  Completion is ordered by the stream (record an event or
 * synchronise after the push); the root learns of it through the caller's barrier.
 * ---------------------------------------------------------------------------------------- */
#define OMFS_IPC_HANDLE_BYTES 64
int omfs_ipc_export(void* d_ptr /* from omfs_device_alloc */, void* h_handle /*[OMFS_IPC_HANDLE_BYTES]*/);
int omfs_ipc_open(const void* h_handle, void** d_ptr);   /* in ANOTHER process than the exporter */
int omfs_ipc_close(void* d_ptr);
int omfs_push_frames(void* d_dst /* local or opened peer memory */, const void* d_src, size_t bytes,
                     void* stream);

/* Level-1 extras used by the parity tests and by omfs_session_set_subject. */
int omfs_flame_fold_subject(int V, int n_shape, int npad, const float* d_template, const float* d_shapedirs,
                            const float* d_shape, const float* d_static, const float* d_plan, const float* d_jreg,
                            float* d_base, void* stream);
int omfs_binning_sort_bits(int S, int width, int height);
int omfs_to_uint8(int S, int width, int height, const float* d_image, uint8_t* d_out, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* OMFS_B200_H */
