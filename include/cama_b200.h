/*
 * cama_b200 — C ABI of the B200-native reprojection hot path of manymuch/CAMA.
 *
 * The reference (pure Python, no FFI of its own) implements this path in
 * cama/reproject.py + cama/pose_transformer.py, driven by cama/dataset.py
 * (file:line below are relative to the reference tree).  This header is what a
 * binding in any host language loads from libcama_b200.so; cama_b200/_native.py
 * is the ctypes binding shipped with the package, INTEGRATION.md shows the stub
 * a maintainer of the reference would add.
 *
 * Conventions
 *   - every entry point returns a cama_status (0 = ok, < 0 = error); the text of the
 *     last error of the calling thread is cama_last_error().  Nothing throws.
 *   - the library never allocates caller-visible memory: all buffers (inputs,
 *     outputs, scratch workspace) are owned by the caller.  "device" pointers are
 *     CUDA device pointers of the context's device (e.g. torch.Tensor.data_ptr());
 *     "host" pointers are ordinary host memory, read synchronously during the call.
 *   - every device call takes the cudaStream_t to enqueue on as a void* (0 = legacy
 *     default stream) and returns after enqueueing; no call synchronises unless its
 *     comment says so.  Calls on different streams are independent as long as they use different
 *     workspaces (everything a later call needs to know about an earlier one — e.g. the layout
 *     cama_clip_stats_read has to read — is in the workspace); the context itself carries a launch
 *     counter, the phase-profiling events (cama_ctx_profile_*: enable/read from one thread at a time)
 *     and the helper streams of the frame-group pipeline, which every cama_clip_render call orders
 *     behind its own stream.
 *   - images are uint8 [H,W,3] BGR, row-major, as in the reference (cv2 convention).
 *   - point arrays are row-major [n,3] (x,y,z) or [n,2] (v,u) = (row,col), like the
 *     reference's "points" arrays; ragged instance lists are flat arrays plus
 *     int64 offsets[I+1] (instance i owns rows offsets[i]..offsets[i+1]).
 *   - arithmetic contract (bit-exactness with the reference's NumPy/BLAS path): every
 *     matrix-vector product accumulates in index order with fused multiply-adds,
 *     a0*b0 first; divisions are IEEE double.  See DESIGN.md "Numerics".
 */
#ifndef CAMA_B200_H_
#define CAMA_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define CAMA_ABI_VERSION 3
#define CAMA_MAX_CAMERAS 8
#define CAMA_TILE_VERTICES 256
#define CAMA_WARP_VERTICES 32
#define CAMA_MAX_PEERS 8          /* GPUs of one box that exchange the sparse output (cama_peer_*) */
#define CAMA_PEER_HEADER_BYTES 256
#define CAMA_PEER_HANDLE_BYTES 64
#define CAMA_CAMERA_TABLE_HEADER 16
#define CAMA_CAMERA_TABLE_BYTES (CAMA_CAMERA_TABLE_HEADER + 256 * 256) /* device bytes of a camera table */

typedef enum cama_status {
    CAMA_OK = 0,
    CAMA_E_INVALID = -1,    /* bad argument (null pointer, negative size, misaligned buffer ...) */
    CAMA_E_CUDA = -2,       /* a CUDA runtime call failed; text in cama_last_error() */
    CAMA_E_WORKSPACE = -3,  /* caller's workspace is smaller than cama_*_workspace_bytes() */
    CAMA_E_CAPACITY = -4,   /* clip path: a record list overflowed; rerun with the capacity in cama_clip_stats */
    CAMA_E_NODEVICE = -5,   /* no CUDA device / not an sm_100 device */
    CAMA_E_UNSUPPORTED = -6 /* shape outside what the requested mode supports */
} cama_status;

typedef struct cama_ctx cama_ctx;

/* ---- library / context -------------------------------------------------------------------- */
int cama_abi_version(void);
const char *cama_last_error(void);
int cama_device_count(int *count);
/* Binds a context to a device (does not change the caller's current device permanently). */
int cama_ctx_create(int device, cama_ctx **out);
int cama_ctx_destroy(cama_ctx *ctx);
/* Number of kernels this context has launched so far (bench.py's gpu_launches claim). */
int cama_ctx_launch_count(const cama_ctx *ctx, uint64_t *count);
int cama_ctx_sm_count(const cama_ctx *ctx, int *count);

/* Per-phase device timing of cama_clip_render (bench.py's roofline figure).  After
 * cama_ctx_profile_enable(ctx, max_calls) every cama_clip_render records CUDA events on ITS OWN
 * stream around its phases (no synchronisation, a few hundred ns each) until max_calls calls have
 * been recorded; max_calls = 0 disables and frees the events.  cama_ctx_profile_read waits for
 * call `call` (0-based since enable) to finish and returns the milliseconds of each phase. */
#define CAMA_CLIP_PHASES 4 /* 0 prep + clears, 1 geometry, 2 work lists of the raster (BINNED), 3 raster */
int cama_ctx_profile_enable(cama_ctx *ctx, int max_calls);
int cama_ctx_profile_calls(const cama_ctx *ctx, int *calls);
int cama_ctx_profile_read(cama_ctx *ctx, int call, float *phase_ms /* [CAMA_CLIP_PHASES] */);

/* ---- per-call operators: one per reference method ------------------------------------------ */

/* MapManager.transform_3d_instance_maps (cama/reproject.py:108-116): out = (T @ [p;1])[:3].
 *   pts      device, [n,3] float32 (pts_is_f32 != 0) or float64
 *   T        host,   16 doubles row-major (a float32 matrix is widened by the caller, exactly)
 *   out      device, [n,3] float64 */
int cama_transform_points(cama_ctx *ctx, const void *pts, int pts_is_f32, int64_t n, const double *T,
                          double *out, void *stream);

/* Scratch needed by the three compacting operators below for n points. */
int cama_compact_workspace_bytes(int64_t n, size_t *bytes);

/* MapManager.crop_3d_instance_maps (cama/reproject.py:118-131), optionally fused with the
 * transform that always precedes it (cama/dataset.py:99-105).  Order-preserving compaction.
 *   pts          device [n,3] (float32 or float64, see pts_is_f32)
 *   T            host 16 doubles or NULL (no transform; pts must then be float64)
 *   box          host 6 doubles {x_min,x_max,y_min,y_max,z_min,z_max}, inclusive
 *   in_offsets   device int64 [n_inst+1]
 *   out_pts      device [n,3] float64 (first out_offsets[n_inst] rows valid)
 *   out_offsets  device int64 [n_inst+1]; an instance with no survivor has equal neighbours */
int cama_crop_points(cama_ctx *ctx, const void *pts, int pts_is_f32, int64_t n, const double *T,
                     const double *box, const int64_t *in_offsets, int64_t n_inst, double *out_pts,
                     int64_t *out_offsets, void *workspace, size_t workspace_bytes, void *stream);

/* CameraManager.project_to_image (cama/reproject.py:187-205), optionally fused with the
 * chassis->camera transform that precedes it (cama/dataset.py:110-115).
 *   pts          device [n,3] float64
 *   T            host 16 doubles or NULL
 *   K            host 9 doubles row-major (already rescaled to the output size)
 *   out_vu       device [n,2] float64, (v,u) per visible point, order preserved
 *   out_offsets  device int64 [n_inst+1] */
int cama_project_points(cama_ctx *ctx, const double *pts, int64_t n, const double *T, const double *K,
                        int width, int height, const int64_t *in_offsets, int64_t n_inst,
                        double *out_vu, int64_t *out_offsets, void *workspace, size_t workspace_bytes,
                        void *stream);

/* CameraManager.render_maps (cama/reproject.py:246-257): stamps the radius-2 filled disc
 * (13 px, |dx|+|dy|<=2, border-clipped) of every point, instance after instance (painter's
 * order), into the image IN PLACE.
 *   vu           device [n,2] float64 (v,u); centres are truncated like astype(np.int32)
 *   in_offsets   device int64 [n_inst+1]
 *   inst_bgr     device uint8 [n_inst,3]
 *   image        device uint8 [height,width,3], in/out
 *   workspace    cama_render_workspace_bytes(height,width) bytes */
int cama_render_workspace_bytes(int height, int width, size_t *bytes);
int cama_render_points(cama_ctx *ctx, const double *vu, int64_t n, const int64_t *in_offsets,
                       int64_t n_inst, const uint8_t *inst_bgr, uint8_t *image, int height, int width,
                       void *workspace, size_t workspace_bytes, void *stream);

/* ---- the batched clip path's record type (also produced by cama_render_points_overlay below) is declared further down;
 * forward declaration for the per-call operator that returns the same records: */
struct cama_overlay_record;
/* render_maps for an image that lives in HOST memory: same centres, same painter's order, but instead of a device image
 * the call returns the lit 8-pixel chunks as cama_overlay_record (CAMA_OVERLAY_BGR) — chunk index inside the
 * [height, width] image, mask of painted pixels, their 24 BGR bytes — which cama_overlay_apply_host(..., CAMA_OVERLAY_DRAW)
 * writes into the host image: exactly the bytes cama_render_points would have changed, without the image crossing PCIe
 * twice.  width % 8 == 0.  records: device, `capacity` records (height * width / 8 always suffice); count: device uint32,
 * set to the number of lit chunks (may exceed capacity: the excess was dropped). */
int cama_render_points_overlay(cama_ctx *ctx, const double *vu, int64_t n, const int64_t *in_offsets, int64_t n_inst,
                               const uint8_t *inst_bgr, int height, int width, struct cama_overlay_record *records,
                               uint32_t *count, int64_t capacity, void *workspace, size_t workspace_bytes, void *stream);

/* ---- load-time densify on the device (scope row N2) ------------------------------------------- */

/* MapManager.load_3d_instance_maps / calculate_3d_instance_maps (cama/reproject.py:42-106, pixel2world_xy
 * :36-40): every polyline segment becomes num = int(len / resolution) float32 points start + (end-start)/num*j,
 * j < num (segment end excluded, num == 0 segments dropped); CAMA (pixel) labels also look the height up in
 * the BEV map and move to world metres.  Bit-identical to the reference's float32 NumPy arithmetic.
 *   raw_xy     device float32 [n_raw,2]  label vertices of all polylines, polyline-major (polylines with
 *              fewer than 2 vertices removed by the caller, as the reference skips them)
 *   raw_poly   device int32 [n_raw]      instance ordinal of each raw vertex (non-decreasing)
 * Two steps, because the output size is data dependent:
 *   cama_densify_plan  -> seg_start device int64 [n_raw+1]: exclusive scan of the per-segment point counts
 *                         (seg_start[n_raw] = total; the caller reads it back to size the output)
 *   cama_densify_fill  -> out_vertices device float4 [total] {x,y,z, bit-cast int32 ordinal}: the
 *                         CAMA_VERTEX_F32X4 layout cama_clip_render takes, so the dense map never visits the host.
 *                         bev_height == NULL: metric labels (x,y kept, z = 0, cama/reproject.py:42-70);
 *                         else float32 [bev_rows,bev_cols] and x = p1*solution - half_width + center_x,
 *                         y = p0*solution - half_height + center_y, z = bev[clip(round(p1)), clip(round(p0))]. */
int cama_densify_plan(cama_ctx *ctx, const float *raw_xy, const int32_t *raw_poly, int64_t n_raw, float resolution,
                      int64_t *seg_start, void *stream);
int cama_densify_fill(cama_ctx *ctx, const float *raw_xy, const int32_t *raw_poly, int64_t n_raw, const int64_t *seg_start,
                      int64_t total, const float *bev_height, int bev_rows, int bev_cols, float solution, float half_width,
                      float half_height, float center_x, float center_y, float *out_vertices, void *stream);

/* ---- undistort-resize of the camera images (scope row N1) --------------------------------------- */

/* CameraManager.resize_image (cama/reproject.py:232-240): cv2.remap(image, map_x, map_y, INTER_LINEAR) on uint8
 * BGR images, bit-identical to OpenCV's fixed-point bilinear path (5 fractional bits, 15-bit weights, border 0).
 *   src    device uint8 [n_images, src_height, src_width, 3]
 *   map_x, map_y  device float32 [n_maps, dst_height, dst_width] from cv2.initUndistortRectifyMap — computed once
 *          per camera (the reference recomputes them per image); image i uses map i % n_maps, so frame-major /
 *          camera-minor images take n_maps = n_cams
 *   dst    device uint8 [n_images, dst_height, dst_width, 3] (may be the `background`/`frames` of cama_clip_render) */
int cama_remap_bilinear(cama_ctx *ctx, const uint8_t *src, int64_t n_images, int src_height, int src_width,
                        const float *map_x, const float *map_y, int64_t n_maps, uint8_t *dst, int dst_height, int dst_width,
                        void *stream);

/* ---- the batched clip path: the loop of cama/dataset.py:78-126 in one call ------------------- */

enum { CAMA_VERTEX_F32X4 = 0,   /* float4 {x,y,z, bit-cast int32 instance ordinal} */
       CAMA_VERTEX_F64X3 = 1 }; /* double[n,3] + int32 vertex_instance[n] */
enum { CAMA_CLIP_AUTO = 0,      /* BINNED when the shape allows it, else PLANE */
       CAMA_CLIP_PLANE = 1,     /* global uint32 centre-id plane + per-pixel dilation (simple, any shape) */
       CAMA_CLIP_BINNED = 2 };  /* band-binned centre records + shared-memory plane + bulk-store raster */

/* One lit 8-pixel chunk of the rendered clip (sparse "overlay" output, BINNED mode only): the chunk
 * covers pixels 8*chunk .. 8*chunk+7 of the flattened [n_frames, n_cams, H, W] pixel array (W % 8 == 0,
 * so a chunk never straddles a row), bit k of mask = pixel k is painted, bgr = the 24 output bytes
 * (painted pixels: instance colour; the others: zero).  Every lit chunk appears exactly once, in no
 * particular order.  cama_overlay_apply_host() draws such records into host frames. */
typedef struct cama_overlay_record {
    uint32_t chunk;
    uint32_t mask;
    uint8_t bgr[24];
} cama_overlay_record;

/* Compact form of the same (CAMA_OVERLAY_PALETTE), usable when the instances have at most 255 distinct
 * colours: one byte per pixel, 0 = not painted, k = palette entry k (`instance_palette` of cama_clip_desc). */
typedef struct cama_overlay_record_palette {
    uint32_t chunk;
    uint8_t index[8];
} cama_overlay_record_palette;
enum { CAMA_OVERLAY_BGR = 0, CAMA_OVERLAY_PALETTE = 1 };
enum { CAMA_PHASE_ALL = 0, CAMA_PHASE_GEOMETRY = 1, CAMA_PHASE_RASTER = 2 };

typedef struct cama_clip_desc {
    uint32_t struct_bytes;          /* sizeof(cama_clip_desc), ABI guard */
    int32_t mode;                   /* CAMA_CLIP_* */
    int32_t n_frames, n_cams, n_instances;
    int32_t height, width;
    int32_t vertex_layout;          /* CAMA_VERTEX_* */
    int64_t n_vertices;
    const void *vertices;           /* device */
    const int32_t *vertex_instance; /* device, only for CAMA_VERTEX_F64X3 */
    const float *world2chassis;     /* device float32 [n_frames,16]: np.linalg.inv(chassis2world.astype(f32)) (cama/dataset.py:92,99) */
    const double *chassis2cam;      /* host   float64 [n_cams,16]  (cama/reproject.py:170) */
    const double *intrinsics;       /* host   float64 [n_cams,9]   rescaled K (cama/reproject.py:180-182) */
    double crop_box[6];             /* x_min,x_max,y_min,y_max,z_min,z_max (cama/reproject.py:28-34) */
    const uint8_t *instance_bgr;    /* device uint8 [n_instances,3]: colour of each instance (cama/reproject.py:251-254) */
    const uint8_t *background;      /* device uint8 [n_frames,n_cams,H,W,3] or NULL (= black); may alias frames */
    uint8_t *frames;                /* device uint8 [n_frames,n_cams,H,W,3] out */
    int32_t *crop_counts;           /* device int32 [n_frames,n_instances] or NULL; caller zero-fills */
    int32_t *visible_counts;        /* device int32 [n_frames,n_cams,n_instances] or NULL; caller zero-fills */
    double *vu_dense;               /* device float64 [n_frames,n_cams,n_vertices,2] or NULL; NaN where not visible */
    int64_t record_capacity;        /* BINNED: centre records each (frame, camera, band group) list of the workspace holds;
                                     * 0 = default (min(max(n_vertices / 4, 2048), 8192 per band of the group)) */
    /* Sparse output (BINNED mode, background must be NULL): when overlay_records != NULL the lit chunks are
     * appended there and `frames` is not written (and may be NULL).  What a host consumer needs crosses PCIe
     * as ~10 % of the dense bytes; the dense frames never exist. */
    /* Optional culling aid: for every tile of CAMA_TILE_VERTICES consecutive vertices the centre and
     * half-extent {cx,cy,cz,ex,ey,ez} of an axis-aligned box containing them.  A (tile, frame) whose
     * transformed box misses the crop box is skipped as a whole; results are unchanged. */
    const double *tile_bounds;      /* device float64 [ceil(n_vertices / CAMA_TILE_VERTICES), 6] or NULL */
    const double *warp_bounds;      /* the same for every CAMA_WARP_VERTICES consecutive vertices (the unit one geometry warp
                                     * works on): device float64 [ceil(n_vertices / CAMA_WARP_VERTICES), 6] or NULL */
    void *overlay_records;          /* device [overlay_capacity] records of overlay_format, or NULL */
    uint32_t *overlay_count;        /* device [1]: records appended (may exceed the capacity: the excess was dropped) */
    int64_t overlay_capacity;
    int32_t overlay_format;         /* CAMA_OVERLAY_BGR | CAMA_OVERLAY_PALETTE (needs instance_palette) */
    int32_t pipeline_frames;        /* BINNED: frames per group of the frame-group pipeline (geometry of group g+1 under the raster of
                                     * group g, useful for clips of hundreds of frames); 0 = library default, < 0 = off */
    const uint8_t *instance_palette; /* device uint8 [n_instances]: palette entry (1..255) of every instance, or NULL */
    /* Frame-sharded clips (cama_peer_* below): every record appended to overlay_records is also stored, at the same
     * position, into overlay_mirrors[0 .. overlay_n_mirrors) — record arrays in the memory of peer GPUs, written over
     * NVLink while the raster runs; chunk indices count from image overlay_image_base (this call's first
     * (frame, camera) image inside the assembled clip: frame_lo * n_cams). */
    void *overlay_mirrors[CAMA_MAX_PEERS];
    int32_t overlay_n_mirrors;
    int32_t reserved0;
    int64_t overlay_image_base;
    /* Optional culling aid: the table cama_camera_table_build made for THESE cameras, crop box and image size (a table
     * built for other ones is recognised by its signature and ignored).  Results are unchanged. */
    const void *camera_table;       /* device, CAMA_CAMERA_TABLE_BYTES, or NULL */
    /* Resident CTAs per SM of the (persistent) geometry kernel; 0 = default (4, which takes every register of an SM).
     * 3 leaves room for a memory-bound kernel of another stream — e.g. the zero-fill of the assembled frames of a
     * frame-sharded clip — to run beside it. */
    int32_t geometry_ctas_per_sm;
    /* Resident CTAs per SM of the raster kernel; 0 = default (4: fastest for one clip alone).  3 for independent clips
     * enqueued on several streams: the geometry CTAs of one clip then fit beside the raster CTAs of another, and the
     * issue-bound geometry runs under the store-bound raster (measured: 104.5 us per clip against 108.2 on config 2). */
    int32_t raster_ctas_per_sm;
    /* Output layout.  mosaic_cols == 0: plain frames uint8 [n_frames,n_cams,H,W,3].  mosaic_cols > 0 (BINNED mode): `frames`
     * (and `background`) are the camera mosaic VideoGenerator.concate_image builds before encoding (cama/tools.py:22-25):
     * uint8 [n_frames, rows*H, mosaic_cols*W, 3] with rows = ceil(n_cams / mosaic_cols) and camera c in tile
     * mosaic_tile_of_cam[c] (row-major) — the raster writes straight into it, concate_image becomes a no-op and a GPU
     * encoder gets its frames without a host hop.  Tiles no camera maps to are not written. */
    int32_t mosaic_cols;
    int32_t mosaic_tile_of_cam[CAMA_MAX_CAMERAS];
    /* Frame-sharded clips, exchange of the CENTRE RECORDS (BINNED mode; cama_peer_* below, cama_b200/shard.py::ListExchange).
     * The record lists may live outside the workspace — in memory the peers can write — and a call may run only one half
     * of the pipeline:
     *   phases = CAMA_PHASE_GEOMETRY  prep + geometry of this call's n_frames frames, which are frames list_frame_base ..
     *            of the lists; only the cursors of this call's frames are cleared and advanced (cama_peer_publish_lists
     *            then copies the filled lists to the peers);
     *   phases = CAMA_PHASE_RASTER    work lists + raster of n_frames frames (normally all list_frames) from lists that are
     *            complete (cursors of the other ranks' frames delivered by cama_peer_publish_cursors / cama_peer_wait).
     * list_records: device uint32 [list_frames * n_cams * n_bands][record_capacity]; list_cursor: device uint32
     * [list_frames * n_cams * n_bands] (n_bands: cama_clip_stats.n_bands of this shape). */
    int32_t phases;                 /* 0 = the whole pipeline */
    int32_t list_frame_base;
    int32_t list_frames;            /* 0 = n_frames */
    int32_t reserved2;
    void *list_records;             /* NULL = inside the workspace */
    uint32_t *list_cursor;
} cama_clip_desc;

typedef struct cama_clip_stats {
    int64_t records_total;          /* centre records emitted (incl. band-halo duplicates) */
    int64_t record_capacity_needed; /* length of the fullest record list: the record_capacity a rerun needs after an overflow */
    int64_t record_capacity;        /* per-list capacity this run used */
    int32_t overflow;               /* != 0: some list exceeded the capacity, frames are incomplete */
    int32_t mode;                   /* mode actually used (CAMA_CLIP_PLANE / CAMA_CLIP_BINNED) */
    int32_t band_rows;              /* BINNED: output rows per band */
    int32_t n_bands;
    int64_t overlay_records;        /* lit chunks produced (sparse output), 0 otherwise */
    int32_t lists_per_image;        /* BINNED: record lists per (frame, camera) image (= n_bands unless bands share lists) */
    int32_t reserved;
} cama_clip_stats;

/* Which cameras can see a point of the crop box at all?  Cuts the crop box's x-y rectangle into cells of about a metre
 * and stores, per cell, the set of cameras for which the visibility conditions of CameraManager.project_to_image
 * (cama/reproject.py:191-198: q_z > 0, 0 <= u < W, 0 <= v < H after cama/dataset.py:110-115's chassis->camera
 * transform) can hold somewhere in the cell (conservatively: maximised over the cell and the box's z range, with a slack
 * far above the rounding error).  cama_clip_render then runs, for every 32 vertices, only the cameras some vertex's cell
 * lists (typically 1-2 of 6) — the frames are bit-identical with and without the table.  Depends on the camera rig, the
 * crop box and the image size only: build once, pass as cama_clip_desc.camera_table.
 *   chassis2cam, intrinsics, crop_box: host, as in cama_clip_desc;  table: device, CAMA_CAMERA_TABLE_BYTES, 16-byte aligned */
int cama_camera_table_build(cama_ctx *ctx, const double *chassis2cam, const double *intrinsics, int n_cams,
                            const double *crop_box, int height, int width, void *table, void *stream);

/* Workspace (device bytes) a cama_clip_render call with this descriptor needs. */
int cama_clip_workspace_bytes(const cama_clip_desc *desc, size_t *bytes);
/* Enqueues the whole clip.  workspace: device, 256-byte aligned. */
int cama_clip_render(cama_ctx *ctx, const cama_clip_desc *desc, void *workspace, size_t workspace_bytes,
                     void *stream);
/* Synchronises `stream` and reads back the counters of the last cama_clip_render that used this
 * workspace.  Returns CAMA_E_CAPACITY when a record list overflowed. */
int cama_clip_stats_read(cama_ctx *ctx, const cama_clip_desc *desc, const void *workspace, void *stream,
                         cama_clip_stats *stats);

/* ---- host side of the sparse output ------------------------------------------------------------ */

/* Where overlay records are drawn: host frames uint8 [n_frames,n_cams,H,W,3] (grid_cols == 0), or the camera
 * mosaic that VideoGenerator.concate_image (cama/tools.py:22-25) builds before encoding: uint8
 * [n_frames, rows*H, grid_cols*W, 3] with rows = ceil(n_cams / grid_cols) and camera c in tile tile_of_cam[c]
 * (row-major) — drawing there makes concate_image a no-op. */
typedef struct cama_overlay_target {
    uint8_t *pixels;
    int64_t n_frames;
    int32_t n_cams, height, width;
    int32_t grid_cols;              /* 0: plain frames */
    const int32_t *tile_of_cam;     /* host int32 [n_cams], only for the mosaic */
} cama_overlay_target;

/* Applies `n` overlay records (format CAMA_OVERLAY_BGR: cama_overlay_record; CAMA_OVERLAY_PALETTE:
 * cama_overlay_record_palette + palette_bgr, host uint8 [256,3], entry 0 unused) to the target; records whose
 * chunk lies outside it are ignored.  DRAW is the in-place draw of CameraManager.render_maps
 * (cama/reproject.py:246-257) for pixels whose colour the GPU has already decided: only painted pixels are
 * written.  BLANK sets them to 0,0,0.  The *_CHUNKS variants write all 24 bytes of each chunk: valid when the
 * unpainted pixels are black anyway (frames without a background), and cheaper.  Only moves bytes
 * (n_threads <= 0: all cores). */
enum { CAMA_OVERLAY_DRAW = 0, CAMA_OVERLAY_BLANK = 1, CAMA_OVERLAY_DRAW_CHUNKS = 2, CAMA_OVERLAY_BLANK_CHUNKS = 3 };
int cama_overlay_apply_host(const void *records, int64_t n, int format, const uint8_t *palette_bgr,
                            const cama_overlay_target *target, int op, int n_threads);

/* The whole return path of the sparse output in one call: `n` records (device) are copied into `staging_pinned`
 * (pinned host memory, at least n records) in slices on `stream`, and every slice is applied to the target with
 * cama_overlay_apply_host(op) while the next one is still in flight.  Synchronises with the copies it issued. */
int cama_overlay_fetch_apply(cama_ctx *ctx, const void *records_dev, int64_t n, int format, const uint8_t *palette_bgr,
                             void *staging_pinned, const cama_overlay_target *target, int op, int n_threads, void *stream);

/* STREAM-like probe of the host memory system on the library's worker pool (n_threads <= 0: all cores): GB/s of a
 * parallel fill (bytes written) and of a parallel copy (bytes read + written) over `bytes`-sized buffers (>= 1 MiB; use
 * several times the last-level cache).  The host side of the sparse output is bound by this. */
int cama_host_bandwidth_probe(int64_t bytes, int n_threads, double *fill_gbs, double *copy_gbs);

/* ---- device side of the sparse output: records -> dense frames ---------------------------------- */

/* Expands `n` overlay records (device) into dense frames uint8 [n_frames,n_cams,H,W,3] (device): the frames are
 * zero-filled first when zero_first != 0, then every record's 8 pixels are written.  With the records of
 * cama_clip_render's sparse output this reproduces its dense output on blank frames byte for byte.  It is the
 * receiving end of a sparse all-gather (cama_b200/shard.py): ranks exchange the lit chunks (~4 % of the dense
 * bytes) instead of the uint8 frames north_star's all-gather moves, and rebuild the frames at HBM speed.
 * palette_bgr: device uint8 [256,3] (CAMA_OVERLAY_PALETTE only); palette_scratch: device, 1 KiB. */
int cama_overlay_expand(cama_ctx *ctx, const void *records, int64_t n, int format, const uint8_t *palette_bgr,
                        void *palette_scratch, uint8_t *frames, int64_t n_frames, int n_cams, int height, int width,
                        int zero_first, void *stream);

/* ---- frame-sharded clips: the sparse output exchanged between the GPUs of one box (BASELINE.json configs[3]) ------
 *
 * The frames of a clip are independent units (cama/dataset.py:88-106), so a site is split into contiguous frame blocks,
 * one per GPU (cama_b200/shard.py).  To end up with EVERY frame on EVERY GPU the ranks do not all-gather the dense uint8
 * frames (NVLink-bound: 3 GB per site) but the lit 8-pixel chunks: cama_clip_render's sparse output is written into a
 * slot of the rank's own mailbox and mirrored, flush by flush, into the same slot of every peer's mailbox
 * (overlay_mirrors: peer stores over NVLink while the raster runs — there is no separate collective);
 * cama_peer_publish then releases the slot, and cama_peer_expand on every rank waits for all slots of the step and
 * rebuilds the dense frames at HBM speed.  No host round trip, no NCCL call on the data path.
 *
 * A mailbox is device memory the LIBRARY allocates (the one exception to "the caller owns all buffers": it has to come
 * from cudaMalloc to be exportable through CUDA IPC), laid out by the caller as [P parities][world] slots of
 * cama_peer_slot_bytes(); a slot = CAMA_PEER_HEADER_BYTES header {uint32 count, uint32 step, ...} | records.  Step s uses
 * parity s % P (P = 2 when a step's calls all go to one stream, 4 when the render of the next step is enqueued on a second
 * stream beside the expand of this one: csrc/peer.cu); steps are numbered from 1 and never reused. */
int cama_peer_slot_bytes(int64_t capacity_records, int record_bytes, size_t *bytes);
/* cudaMalloc + zero-fill of `bytes` on the context's device; ipc_handle: CAMA_PEER_HANDLE_BYTES bytes out (host), to be sent
 * to the peer processes.  Synchronises. */
int cama_peer_alloc(cama_ctx *ctx, size_t bytes, void **dev_ptr, void *ipc_handle);
int cama_peer_free(cama_ctx *ctx, void *dev_ptr);
/* Maps the mailbox another process allocated on device `peer_device` (same box) into this process; enables peer access
 * from the context's device.  CAMA_E_UNSUPPORTED when there is no peer-to-peer path between the two devices. */
int cama_peer_open(cama_ctx *ctx, int peer_device, const void *ipc_handle, void **dev_ptr);
int cama_peer_close(cama_ctx *ctx, void *dev_ptr);
/* After cama_clip_render on `stream`: writes *overlay_count (device) and then, with release semantics at system scope,
 * `step` into the n slot headers (device pointers, own and peers'; host array). */
int cama_peer_publish(cama_ctx *ctx, const uint32_t *overlay_count, uint32_t step, void *const *slot_headers, int n, void *stream);
/* slots: host array of `world` device pointers, the slots of this rank's own mailbox that hold the records of rank
 * 0..world-1 for `step`.  Slot after slot, starting with own_rank's: waits on the device until the slot carries `step`
 * (at most timeout_ms in all, <= 0: 10000), then writes the 8 pixels of each of its records into frames uint8
 * [n_frames,n_cams,H,W,3] (device, zero-filled by the caller, e.g. with cama_frames_clear).
 * status: device int32 [1], caller zero-fills once; set to 1 when a peer's step did not arrive in time (its records are
 * missing), 2 when a slot held more records than capacity_records (frames incomplete).  palette_*: as
 * cama_overlay_expand. */
int cama_frames_clear(cama_ctx *ctx, uint8_t *frames, size_t bytes, void *stream);   /* cudaMemsetAsync(frames, 0, bytes) on `stream` */
int cama_peer_expand(cama_ctx *ctx, void *const *slots, int world, int own_rank, uint32_t step, int64_t capacity_records, int format,
                     const uint8_t *palette_bgr, void *palette_scratch, uint8_t *frames, int64_t n_frames, int n_cams,
                     int height, int width, int timeout_ms, int32_t *status, void *stream);

/* The same hand-off for the exchange of the CENTRE RECORDS (cama_clip_desc.phases / list_*): every rank's geometry has
 * mirrored the records of its frames into the peers' list arrays; cama_peer_publish_cursors copies the rank's cursor range
 * [first, first + count) (own_cursor: this rank's array) into the same range of the n_peers peer arrays and then writes
 * `step`, with release semantics at system scope, into the n_headers headers (own and peers'); cama_peer_wait (one warp)
 * waits on `stream` until the `world` headers in this rank's own memory carry `step` (timeout_ms <= 0: 10000; status 1 on a
 * time-out), so that the raster-phase call enqueued after it sees complete lists.  Every rank then rasters EVERY frame from
 * the lists: the dense frames are written once, by the HBM-bound kernel that writes them anyway — no zero-fill, no expand. */
int cama_peer_publish_cursors(cama_ctx *ctx, const uint32_t *own_cursor, int64_t first, int64_t count,
                              void *const *peer_cursors, int n_peers, uint32_t step, void *const *headers, int n_headers,
                              void *stream);
int cama_peer_wait(cama_ctx *ctx, void *const *headers, int world, uint32_t step, int timeout_ms, int32_t *status,
                   void *stream);
/* cama_peer_publish_cursors preceded by the records themselves: the filled part of each of the rank's lists [first, first +
 * count) is copied to the same place of every peer's list array with wide coalesced peer stores (a geometry call that ran
 * (mirroring record by record from the geometry kernel was measured: 8x slower over NVLink at 8 GPUs).
 * capacity: records per list, a multiple of 4; arrays 16-byte aligned. */
int cama_peer_publish_lists(cama_ctx *ctx, const void *own_records, const uint32_t *own_cursor, int64_t capacity, int64_t first,
                            int64_t count, void *const *peer_records, void *const *peer_cursors, int n_peers, uint32_t step,
                            void *const *headers, int n_headers, void *stream);

/* ---- LiDAR aggregation (BASELINE.json configs[4]) -------------------------------- */

/* Axis-aligned voxel grid: voxel (ix,iy,iz) covers origin + [i, i+1) * voxel on every axis; counts are stored
 * uint32 [nz, ny, nx]. */
typedef struct cama_voxel_grid {
    double origin[3];               /* world coordinates of the grid's minimum corner */
    double voxel[3];                /* edge lengths, > 0 */
    int32_t dims[3];                /* nx, ny, nz */
    int32_t reserved;
} cama_voxel_grid;

/* For every sweep s and every point row i in [sweep_offsets[s], sweep_offsets[s+1]):
 *   world = (transforms[s] @ [x y z 1])[:3]   — MapManager.transform_3d_instance_maps, cama/reproject.py:108-116,
 *           with transforms[s] = chassis2world(t_s) @ lidar2chassis (PoseTransformer.seek_by_timestamp,
 *           cama/pose_transformer.py:589-652; DatasetReader.get_extrinsic, cama/dataset_reader.py:222-248)
 *   counts[floor((world - origin) / voxel)] += 1 when the voxel index is inside the grid.
 * points: device float64 rows of `row_doubles` values, x y z first (DatasetReader.yield_lidar,
 * cama/dataset_reader.py:45-51: 6 per row); sweep_offsets: device int64 [n_sweeps+1]; transforms: device float64
 * [n_sweeps,16] row-major 4x4; counts: device uint32 [nz,ny,nx], accumulated into (caller zero-fills); n_inside:
 * optional device uint64 [1], incremented by the number of points counted.  The voxel accumulation is not in the
 * reference snapshot (camav2 branch, README.md:17-20): this definition is the library's own, restated by
 * oracle/lidar_oracle.py — parity unpinned. */
int cama_lidar_accumulate(cama_ctx *ctx, const double *points, int row_doubles, const int64_t *sweep_offsets,
                          int n_sweeps, const double *transforms, const cama_voxel_grid *grid, uint32_t *counts,
                          uint64_t *n_inside, void *stream);

#ifdef __cplusplus
}
#endif
#endif /* CAMA_B200_H_ */
