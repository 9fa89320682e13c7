// libcama_b200: the batched clip path — the whole frame x camera x vertex loop of
// /root/reference/cama/dataset.py:78-126 (yield_frame -> project_all_camera -> render_maps)
// for every frame of a clip in a fixed handful of launches.
//
// Pipeline (BINNED mode, the throughput path):
//   prep      w2c float32 -> float64 (exact), instance colours -> packed LUT
//   geometry  FP64: world->chassis, crop box, cameras x (chassis->camera, K, /z, mask); every visible
//             centre becomes a 4-byte record {ordinal+1 : 16 | row in its band group : 16 - x_bits | x : x_bits}
//             appended straight to the record list of its (frame, camera, band group) — a group is a few
//             consecutive bands; lists have a fixed capacity and a cursor, no sort pass follows
//   classify  the (frame, camera, band) work items of the raster by the weight of their list
//   raster    one CTA per (frame, camera, band), reading the band's rows out of its group's list: uint16 centre plane in shared memory
//             (atomic max of ordinals), L1-radius-2 max-dilation with packed u16x2 max,
//             colour LUT, optional composite over a background, rows staged in shared memory
//             and written with bulk async copies (cp.async.bulk shared -> global).
// PLANE mode keeps the centre ids in a global uint32 plane and dilates per pixel: slower, no
// shape restrictions, and an independent implementation the tests cross-check BINNED against.
#include <algorithm>
#include <climits>
#include <cstdlib>

#include <nvtx3/nvToolsExt.h>      // header-only; the ranges cost nothing unless a profiler is attached

#include "common.cuh"
#include "geom.cuh"

namespace cama {

// NVTX range over a scope: `ncu --nvtx --nvtx-include "cama_clip_render/"` (or any NVTX-aware profiler) can pick a
// call or one of its phases.
struct NvtxRange {
    explicit NvtxRange(const char *name) { nvtxRangePushA(name); }
    ~NvtxRange() { nvtxRangePop(); }
};

constexpr unsigned kFull = 0xffffffffu;

// Programmatic dependent launch: the kernels of a clip are launched with programmatic stream
// serialisation, so a kernel's CTAs may become resident (and run the prologue that does not depend on
// earlier kernels) while the previous kernel drains.  pdl_wait() blocks until the previous grid has
// completed and its writes are visible (a no-op when the launch did not carry the attribute);
// pdl_trigger() lets the next kernel of the stream start launching once every CTA of this one has called it.
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

struct CamBlock {                      // by-value kernel parameter (constant bank)
    double E[CAMA_MAX_CAMERAS][12];    // chassis -> camera, rows 0..2 of the 4x4
    double K[CAMA_MAX_CAMERAS][9];
    double box[6];
    int k_row2_is_001[CAMA_MAX_CAMERAS];
    double wlim, hlim;                 // (double)(width + 1), (double)(height + 1): bounds of the frustum pre-test
    int all_pinhole;                   // every K is exactly [[fx,0,cx],[0,fy,cy],[0,0,1]]
};

struct ClipArgs {
    int n_frames, n_cams, n_instances, height, width;
    long long n_vertices;
    const void *vertices;
    const int *vertex_instance;
    const double *w2c64;               // [F,12]
    const double *tile_bounds;         // [n_tiles,6] centre + half-extent, or nullptr
    const unsigned long long *worklist; // live units {unit << 8 | frame mask} from geometry_cull_kernel, or nullptr (all units)
    const unsigned *n_live;
    const double *warp_bounds;         // [ceil(n_vertices / 32),6] centre + half-extent of every warp's 32 vertices, or nullptr
    unsigned *geo_counter;             // [1] dynamic claims of the geometry warps (cleared by prep)
    const unsigned char *cam_table;    // 16-byte header {signature} | [tab_ny][tab_nx] cameras that can see a chassis-frame cell (cama_camera_table_build), or nullptr
    unsigned long long tab_signature;  // of the cameras / crop box / image size of THIS call: a table built for others is ignored
    int unit_frames;                   // frames per warp unit of the geometry kernel: 1, 2, 4 or 8
    unsigned static_units;             // warp units dealt round-robin before the warps start claiming dynamically
    double tab_x0, tab_y0;
    float tab_inv_sx, tab_inv_sy;
    int tab_nx, tab_ny;
    int *crop_counts;
    int *visible_counts;
    double *vu_dense;
    // PLANE
    unsigned *plane;                   // [F,C,H,W]
    // BINNED
    int band_rows, n_bands, x_bits;
    int group_rows, n_groups;          // rows of a band group (band_rows * bands per group), groups per image
    unsigned group_magic;              // ceil(2^32 / group_rows): row / group_rows == umulhi(row, group_magic) for rows < 65536
    unsigned list_cap;                 // records a list holds
    unsigned *cursor;                  // [F*C*n_groups] records appended (attempted) to each list so far (cleared by prep)
    unsigned *records;                 // [F*C*n_groups][list_cap] payloads
    int list_frame_base;               // frame f of this call is frame list_frame_base + f of the lists (frame-sharded clips)
};

// ------------------------------------------------------------------------------------------------ camera table
// Which cameras can see a point of the crop box at all?  The box's x-y rectangle is cut into cells (about a metre);
// for every cell and camera the five linear forms that decide visibility — q_z > 0, q_x >= 0, W q_z - q_x > 0,
// q_y >= 0, H q_z - q_y > 0 with q = K (E [x y z 1]) (cama/dataset.py:110-115, cama/reproject.py:191-198) — are
// maximised over the cell (grown by 2 % for the rounding of the cell lookup) x the box's z range; a camera stays in
// the cell's mask unless one of them is negative by more than a slack ~1e7 times the rounding error of the chain.
// The geometry warps then only run the cameras in the OR of their lanes' masks (typically 1-2 of 6).  Conservative:
// a masked-out camera could not have produced a visible point, so results are unchanged; NaN/inf entries never cull.
struct CamTable {
    double x0, y0, sx, sy;                 // cell (ix, iy) covers x0 + [ix, ix+1) sx, y0 + [iy, iy+1) sy
    int nx, ny;
};
constexpr int kCamTableMaxDim = 256;

__device__ __forceinline__ double form_max(const double *f, const double *lo, const double *hi, double &scale) {
    double m = f[3];
    scale = fabs(f[3]);
#pragma unroll
    for (int j = 0; j < 3; ++j) {
        if (f[j] != 0.0) {
            m += fmax(f[j] * lo[j], f[j] * hi[j]);
            scale += fabs(f[j]) * fmax(fabs(lo[j]), fabs(hi[j]));
        }
    }
    return m;
}

// one (cell, camera): may the camera see anything inside the cell?
__device__ bool camera_sees_cell(const CamBlock &cams, int c, int width, int height, const CamTable &t, int cell) {
    const int ix = cell % t.nx, iy = cell / t.nx;
    const double lo[3] = {t.x0 + (ix - 0.02) * t.sx, t.y0 + (iy - 0.02) * t.sy, cams.box[4]};
    const double hi[3] = {t.x0 + (ix + 1.02) * t.sx, t.y0 + (iy + 1.02) * t.sy, cams.box[5]};
    const double *E = cams.E[c], *K = cams.K[c];
    double q[3][4];                                                // q_r = sum_j K[r][j] * (row j of E), as a form over [x y z 1]
#pragma unroll
    for (int r = 0; r < 3; ++r)
#pragma unroll
        for (int j = 0; j < 4; ++j) q[r][j] = K[3 * r] * E[j] + K[3 * r + 1] * E[4 + j] + K[3 * r + 2] * E[8 + j];
    bool out = false;
#pragma unroll
    for (int k = 0; k < 5; ++k) {
        double f[4];
#pragma unroll
        for (int j = 0; j < 4; ++j)
            f[j] = k == 0 ? q[2][j] : k == 1 ? q[0][j] : k == 2 ? (double)width * q[2][j] - q[0][j] : k == 3 ? q[1][j] : (double)height * q[2][j] - q[1][j];
        double scale;
        const double m = form_max(f, lo, hi, scale);
        out = out || (m < -(1e-6 * scale + 1e-9));               // (false for NaN: never culls)
    }
    return !out;
}

// 8 lanes per cell, one camera each; the lane of camera 0 writes the cell's mask.  The table depends on the camera rig,
// the crop box and the image size only: built once per rig (cama_camera_table_build), not per clip.
__global__ void __launch_bounds__(256) camera_table_kernel(const __grid_constant__ CamBlock cams, int n_cams, int width, int height, const CamTable tab,
                                                          unsigned long long signature, unsigned char *__restrict__ table) {
    static_assert(CAMA_MAX_CAMERAS == 8, "camera table: one byte per cell, 8 lanes per cell");
    const int n_cells = tab.nx * tab.ny;
    const int k = blockIdx.x * 256 + threadIdx.x, cell = k >> 3, c = k & 7;
    const bool sees = cell < n_cells && c < n_cams && camera_sees_cell(cams, c, width, height, tab, cell);
    const unsigned votes = __ballot_sync(kFull, sees);
    if (c == 0 && cell < n_cells) table[CAMA_CAMERA_TABLE_HEADER + cell] = (unsigned char)((votes >> (threadIdx.x & 24)) & 0xffu);
    if (k == 0) *reinterpret_cast<unsigned long long *>(table) = signature;
}

// ------------------------------------------------------------------------------------------------ prep
// (also clears the per-call counters, so that a clip costs kernel launches only: zero_words 32-bit words at `zero`,
// the statistics block, the sparse-output counter)
__global__ void prep_kernel(const float *__restrict__ w2c, int n_frames, double *__restrict__ w2c64,
                            const uint8_t *__restrict__ inst_bgr, int n_inst, unsigned *__restrict__ lut,
                            unsigned *__restrict__ zero, long long zero_words, unsigned *__restrict__ stats, int stats_words,
                            unsigned *__restrict__ overlay_count, const uint8_t *__restrict__ inst_palette,
                            unsigned *__restrict__ zero2, long long zero2_words, unsigned groups_used) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    pdl_wait();                                    // (the previous clip's raster still reads the counters cleared here)
    pdl_trigger();
    for (long long k = i; k < zero_words; k += (long long)gridDim.x * blockDim.x) zero[k] = 0u;
    for (long long k = i; k < zero2_words; k += (long long)gridDim.x * blockDim.x) zero2[k] = 0u;      // (external list cursors of this call's frames)
    if (i < stats_words) stats[i] = i == stats_words - 1 ? groups_used : 0u;       // (ClipStatsDev: counters cleared, groups_used set)
    if (i == 0 && overlay_count) *overlay_count = 0u;
    if (w2c && i < n_frames * 12) w2c64[i] = (double)w2c[(i / 12) * 16 + (i % 12)];
    // lut[ordinal+1] = B | G << 8 | R << 16 | palette entry << 24 (the raster never looks at the top byte of a colour)
    if (i <= n_inst) lut[i] = i == 0 ? 0u
                                      : (unsigned)inst_bgr[3 * (i - 1)] | ((unsigned)inst_bgr[3 * (i - 1) + 1] << 8) |
                                            ((unsigned)inst_bgr[3 * (i - 1) + 2] << 16) | (inst_palette ? (unsigned)inst_palette[i - 1] << 24 : 0u);
}

// ------------------------------------------------------------------------------------------------ geometry
template <int LAYOUT>
__device__ __forceinline__ bool load_vertex(const ClipArgs &a, long long n, double &x, double &y, double &z, int &ord) {
    if (n >= a.n_vertices) return false;
    if (LAYOUT == CAMA_VERTEX_F32X4) {
        const float4 v = __ldg(static_cast<const float4 *>(a.vertices) + n);
        x = (double)v.x; y = (double)v.y; z = (double)v.z;
        ord = __float_as_int(v.w);
    } else {
        const double *p = static_cast<const double *>(a.vertices) + 3 * n;
        x = p[0]; y = p[1]; z = p[2];
        ord = a.vertex_instance[n];
    }
    return true;
}

// One chassis-frame point against one camera (reference cama/dataset.py:110-115 + reproject.py:187-205)
// up to, but not including, the perspective divide.  Returns false for points the visibility mask
// is certain to discard; the early exits only skip work whose result the mask would discard anyway:
//  * K row 2 == (0,0,1) makes q_z == p_z bit-for-bit, so p_z <= 0 rejects before x,y are formed;
//  * q_x < -q_z or q_x > (W+1) q_z (same for y) puts u (v) outside [0,W) by a whole pixel, far
//    beyond what the rounding of the division could undo (NaN coordinates may or may not pass: the
//    pixel test that follows masks them).
// PINHOLE: K is exactly [[fx,0,cx],[0,fy,cy],[0,0,1]] (what CameraManager builds, cama/reproject.py:180-182).  The
// products with the zero entries add an exact zero in NumPy's accumulation (a0*b0, then one fma per further term)
// and the one with the 1 returns p_z, so q = (fma(cx,pz,fx*px), fma(cy,pz,fy*py), pz) bit for bit for finite
// camera coordinates — 4 FP64 operations instead of 9.  (A non-finite p_x or p_y makes q_x or q_y non-finite here
// and NaN there: the point is masked either way.)
template <bool PINHOLE>
__device__ __forceinline__ bool camera_candidate(const CamBlock &cams, int c, double cx, double cy, double cz, int width, int height,
                                                 double &qx, double &qy, double &qz) {
    const double *E = cams.E[c];
    const double *K = cams.K[c];
    const double pz = affine_row(E + 8, cx, cy, cz);
    if (PINHOLE) {                              // q_z == p_z: 0 < p_z <= DBL_MAX decided before x and y are formed
        if (!((pz > 0.0) & ((unsigned)__double2hiint(pz) < 0x7ff00000u))) return false;
    } else if (cams.k_row2_is_001[c] && !(pz > 0.0)) {
        return false;
    }
    const double px = affine_row(E, cx, cy, cz);
    const double py = affine_row(E + 4, cx, cy, cz);
    if (PINHOLE) {
        qz = pz;
        qx = __fma_rn(K[2], pz, __dmul_rn(K[0], px));
        qy = __fma_rn(K[5], pz, __dmul_rn(K[4], py));
    } else {
        qz = linear_row(K + 6, px, py, pz);
        if (!((qz > 0.0) & (qz <= DBL_MAX))) return false;
        qx = linear_row(K, px, py, pz);
        qy = linear_row(K + 3, px, py, pz);
    }
    // q_x + q_z < 0, q_y + q_z < 0, (W+1) q_z - q_x < 0 or (H+1) q_z - q_y < 0: one operation each, the four sign bits
    // ORed (the comparison form compiled to a NaN-aware min/max sequence of 16 instructions).  The sum of two doubles
    // has the sign of the exact sum, the fused product-difference that of the exact value; a visible point has
    // q_x >= 0 and q_x < W q_z (1 + 2^-52) < (W+1) q_z, so none of the four is negative for it.
    const double lo_x = __dadd_rn(qx, qz), lo_y = __dadd_rn(qy, qz);
    const double hi_x = __fma_rn(cams.wlim, qz, -qx), hi_y = __fma_rn(cams.hlim, qz, -qy);
    return (__double2hiint(lo_x) | __double2hiint(lo_y) | __double2hiint(hi_x) | __double2hiint(hi_y)) >= 0;
}

// Pixel of a candidate: trunc(fl(q_x / q_z)), trunc(fl(q_y / q_z)) and the in-image test of the reference,
// bit-exact, without a double-precision division for almost every point: a float32 estimate of the
// quotient (error < 1.1e-3 for |u| <= 2049: two conversions, an approximate reciprocal, a multiply)
// decides the pixel whenever it is further than kPixelGuard from an integer; only the rest take
// the IEEE divisions.  `exact_v/u` receive the quotients when `want_exact` (dense (v,u) output).
constexpr float kPixelGuard = 4e-3f;
__device__ __forceinline__ bool candidate_pixel(bool cand, double qx, double qy, double qz, int width, int height, bool want_exact,
                                                int &vi, int &ui, double &v, double &u) {
    const float rz = __frcp_rn(__double2float_rn(qz));
    const float ue = __double2float_rn(qx) * rz, ve = __double2float_rn(qy) * rz;
    const float uf = floorf(ue), vf = floorf(ve);
    const bool safe = (ue - uf > kPixelGuard) & (uf + 1.0f - ue > kPixelGuard) & (ve - vf > kPixelGuard) & (vf + 1.0f - ve > kPixelGuard) &
                      (fabsf(ue) < 4096.0f) & (fabsf(ve) < 4096.0f);
    bool vis = cand & (uf >= 0.0f) & (uf < (float)width) & (vf >= 0.0f) & (vf < (float)height);
    ui = (int)uf; vi = (int)vf;
    if (cand && (want_exact || !safe)) {
        u = __ddiv_rn(qx, qz);
        v = __ddiv_rn(qy, qz);
        vis = (u >= 0.0) & (u < (double)width) & (v >= 0.0) & (v < (double)height);
        ui = __double2int_rz(u);       // reference cama/reproject.py:249 (values are >= 0: trunc == floor)
        vi = __double2int_rz(v);
    }
    return vis;
}

// Records {list, payload} are staged in shared memory, per warp, and flushed to their band lists >= 128 at a time
// (stage_flush): the returning atomics that reserve the places are issued per flush round, not per append, and no
// block-level barrier is involved.  (A version that reserved with returning atomics per warp and append spent a
// quarter of its time waiting for them; one that staged per CTA spent a third of it in the per-frame barrier.)
#ifndef CAMA_PIPE_FRAMES_DEFAULT
#define CAMA_PIPE_FRAMES_DEFAULT 160
#endif
#ifndef CAMA_GEO_MINB
#define CAMA_GEO_MINB 4                    // resident CTAs per SM the geometry kernel is compiled for
#endif
#ifndef CAMA_GEO_STAGE_FLUSH
#define CAMA_GEO_STAGE_FLUSH 128
#endif
constexpr int kGeoThreads = 256;
constexpr int kStageFlush = CAMA_GEO_STAGE_FLUSH;                                       // a warp flushes once this many records are staged ...
constexpr int kStageCap = kStageFlush + 32 * 2 * CAMA_MAX_CAMERAS;     // ... so one more frame (<= 2 records per vertex and camera) always fits

struct GeoStage {                          // one per warp
    uint2 rec[kStageCap];
    unsigned count;
    unsigned pad[3];
};

__device__ __forceinline__ void warp_append(const ClipArgs &a, GeoStage &st, bool pred, unsigned list, unsigned payload) {
    const unsigned mask = __ballot_sync(kFull, pred);
    if (mask == 0) return;
    const int lane = threadIdx.x & 31;
    const unsigned slot = st.count + __popc(mask & ((1u << lane) - 1u));                  // the warp owns its stage: no atomic
    if (pred) st.rec[slot] = make_uint2(list, payload);
    __syncwarp();
    if (lane == 0) st.count += (unsigned)__popc(mask);
    __syncwarp();
}

// Warp-wide: move the warp's staged records {list, payload} to their lists.  32 records per round: match.any groups
// the lanes of a list (staged records come in runs: one list per (frame, camera) of the warp's 32 vertices), the first
// lane of every group reserves the group's places with ONE returning atomic on the list's cursor, and the atomics of
// up to four rounds are in flight before the first result is used — the latency of a reservation is paid once per
// flush (>= 128 records), not once per append.  Records past a list's capacity are dropped (the cursor keeps counting:
// band_classify_kernel reports the overflow).
__device__ __forceinline__ void stage_flush(const ClipArgs &a, GeoStage &st) {
    const int lane = threadIdx.x & 31;
    const unsigned n = st.count;
    const unsigned lt = (1u << lane) - 1u;
    for (unsigned i0 = 0; i0 < n; i0 += 128u) {
        uint2 r[4];
        unsigned peers[4], first[4];
        bool live[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const unsigned i = i0 + 32u * k + lane;
            live[k] = i < n;
            r[k] = live[k] ? st.rec[i] : make_uint2(0xffffffffu, 0u);
            peers[k] = __match_any_sync(kFull, r[k].x);
            first[k] = 0u;
            if (live[k] && lane == __ffs(peers[k]) - 1) first[k] = atomicAdd(&a.cursor[r[k].x], (unsigned)__popc(peers[k]));
        }
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const unsigned pos = __shfl_sync(kFull, first[k], __ffs(peers[k]) - 1) + (unsigned)__popc(peers[k] & lt);
            if (live[k] && pos < a.list_cap) a.records[(size_t)r[k].x * a.list_cap + pos] = r[k].y;
        }
    }
    __syncwarp();
    if (lane == 0) st.count = 0;
    __syncwarp();
}

template <bool BINNED, bool DEBUG>
__device__ __forceinline__ void emit_centre(const ClipArgs &a, GeoStage &st, int f, int c, bool vis, int vi, int ui, double v, double u, int ord, long long n) {
    if (DEBUG && vis) {
        if (a.visible_counts) atomicAdd(&a.visible_counts[((size_t)f * a.n_cams + c) * a.n_instances + ord], 1);
        if (a.vu_dense) {
            double *o = a.vu_dense + (((size_t)f * a.n_cams + c) * a.n_vertices + n) * 2;
            o[0] = v; o[1] = u;
        }
    }
    if (!BINNED) {
        if (vis) atomicMax(&a.plane[(((size_t)f * a.n_cams + c) * a.height + vi) * a.width + ui], (unsigned)(ord + 1));
    } else {
        // a centre equal to the previous lane's (same pixel, same instance: dense far-away vertices) adds nothing to a max
        const unsigned code = ((unsigned)vi << 16) | (unsigned)ui;
        const unsigned prev_code = __shfl_up_sync(kFull, vis ? code : 0xffffffffu, 1);
        const int prev_ord = __shfl_up_sync(kFull, ord, 1);
        if ((threadIdx.x & 31) != 0 && prev_code == code && prev_ord == ord) vis = false;
        if (!__any_sync(kFull, vis)) return;
        const int rg = a.group_rows;
        const int g0 = (int)__umulhi((unsigned)vi, a.group_magic);  // vi / rg
        const int r = vi - g0 * rg;                                  // row inside the band group (stored + 2: rows -2, -1 belong to the halo)
        const unsigned list = (unsigned)(((a.list_frame_base + f) * a.n_cams + c) * a.n_groups + g0);
        const unsigned key = (unsigned)(ord + 1) << 16;
        warp_append(a, st, vis, list, key | (unsigned)(((r + 2) << a.x_bits) | ui));
        // the two rows next to a group edge also matter to the neighbouring group (dilation radius 2); band edges
        // inside a group need nothing: the bands of a group read the same list
        const bool up = vis && r < 2 && g0 > 0;
        const bool down = vis && r >= rg - 2 && g0 + 1 < a.n_groups;
        if (__any_sync(kFull, up || down)) {
            const unsigned list2 = up ? list - 1 : list + 1;
            const int r2 = up ? r + rg : r - rg;
            warp_append(a, st, up || down, list2, key | (unsigned)(((r2 + 2) << a.x_bits) | ui));
        }
    }
}

constexpr int kGeoFrames = 8;              // frames per work unit (vertex loads amortised over them)
static_assert(kGeoThreads == CAMA_TILE_VERTICES, "tile_bounds are per geometry tile");

// Tile culling: the tile's bounding box, moved to the chassis frame of a frame (centre by the affine
// map, half-extent by |R|), against the crop box.  Conservative: the slack is ~1e7 times the rounding
// error of these few operations, and a surviving tile is decided vertex by vertex as before.
// Non-finite bounds or poses never cull (comparisons with NaN are false).
__device__ __forceinline__ bool tile_may_survive(const double *__restrict__ b, const double *__restrict__ T, const double *box) {
    bool out = false;
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        const double c = T[4 * k] * b[0] + T[4 * k + 1] * b[1] + T[4 * k + 2] * b[2] + T[4 * k + 3];
        const double e = fabs(T[4 * k]) * b[3] + fabs(T[4 * k + 1]) * b[4] + fabs(T[4 * k + 2]) * b[5];
        const double slack = 1e-6 + 1e-9 * (fabs(c) + e);
        out = out || (c - e - slack > box[2 * k + 1]) || (c + e + slack < box[2 * k]);
    }
    return !out;
}

// Large clips (sites): most (tile, 8-frame chunk) units contain nothing near the vehicle.  One thread
// per (unit, frame) evaluates one culling test (8 lanes per unit, 4 units per warp); the live units
// {unit index, frame mask} are appended to a work list, one atomic per warp, so that the geometry
// kernel only ever starts units with something to do.  (One thread per unit, walking its 8 frames,
// took 56 us on the 120 k units of a site: a chain of dependent loads per thread.)
static_assert(kGeoFrames == 8, "geometry_cull_kernel packs 8 frames of a unit into 8 lanes");
__global__ void __launch_bounds__(256) geometry_cull_kernel(const double *__restrict__ tile_bounds, const double *__restrict__ w2c64,
                                                           long long units, int n_chunks, int n_frames, const __grid_constant__ CamBlock cams,
                                                           unsigned long long *__restrict__ worklist, unsigned *__restrict__ n_live) {
    const long long t = (long long)blockIdx.x * 256 + threadIdx.x;
    const long long unit = t >> 3;
    const int fi = (int)(t & 7);
    const int lane = threadIdx.x & 31;
    pdl_wait();
    pdl_trigger();
    bool frame_live = false;
    if (unit < units) {
        const long long tile = unit / n_chunks;
        const int f = (int)(unit % n_chunks) * kGeoFrames + fi;
        frame_live = f < n_frames && tile_may_survive(tile_bounds + tile * 6, w2c64 + (size_t)f * 12, cams.box);
    }
    const unsigned votes = __ballot_sync(kFull, frame_live);
    const unsigned mask = (votes >> (lane & ~7)) & 0xffu;            // frame mask of this lane's unit
    const bool leader = fi == 0 && mask != 0u;
    const unsigned live = __ballot_sync(kFull, leader);
    if (live == 0u) return;
    const int first = __ffs(live) - 1;
    unsigned base = 0;
    if (lane == first) base = atomicAdd(n_live, (unsigned)__popc(live));
    base = __shfl_sync(kFull, base, first);
    if (leader) worklist[base + __popc(live & ((1u << lane) - 1u))] = ((unsigned long long)unit << 8) | mask;
}

// Work unit of a WARP = (32 consecutive vertices, kUnitFrames consecutive frames).  The kernel is a persistent grid of
// independent warps: every warp claims its next unit from a global counter (one returning atomic per unit,
// issued a whole unit ahead of its use), keeps the poses of the unit in its own slice of shared memory, and
// never meets a block-level barrier — so a warp whose vertices all fall outside the crop box moves on at once,
// and no CTA slot idles behind its slowest warp.  Per unit: the warp's own bounding box (`warp_bounds`) decides
// which of the unit's frames it has to look at (lane = frame); a thread keeps its vertex in registers; lanes hold
// consecutive vertices, so crop survival — and with it the camera tail — is almost warp-uniform (polylines are
// spatially coherent).  Cameras: the table of cama_camera_table_build gives every surviving lane the cameras that
// can see its cell; the warp runs the OR of them (typically 1-2 of 6), in a ROLLED loop over the set bits: with the
// loop unrolled six times the kernel was 52 KB of code and a quarter of the stall samples were instruction fetches
// (independent warps are all over the code).
// With a work list (sites: geometry_cull_kernel has culled (256-vertex tile, 8 frames) units) entry e covers the
// 8 * kGeoFrames / kUnitFrames warp units of its tile and frames.
// Frames per unit (a.unit_frames, a power of two <= 8): 4 for ordinary clips (shorter units balance better: 53.8 us
// against 57.9 with 8 on config 2), 8 behind a work list (sites: most warp units die in the bounds test, and the
// per-unit overhead — claim, poses, bounds — is what counts: 504 us against 651 with 4 on config 3).
// Claims: the first a.static_units units are dealt round-robin (no atomic), the rest dynamically — the claims only
// matter for the tail, and 30 k returning atomics on one address were 18 % of the stall samples.
#ifndef CAMA_GEO_UNIT_FRAMES
#define CAMA_GEO_UNIT_FRAMES 4
#endif
#ifndef CAMA_GEO_UNIT_FRAMES_SITE
#define CAMA_GEO_UNIT_FRAMES_SITE 8
#endif
#ifndef CAMA_GEO_STATIC_PCT
#define CAMA_GEO_STATIC_PCT 60
#endif

template <int LAYOUT, bool BINNED, bool DEBUG, bool PINHOLE>
__global__ void __launch_bounds__(kGeoThreads, CAMA_GEO_MINB) clip_geometry_kernel(const ClipArgs a, const __grid_constant__ CamBlock cams) {
    __shared__ double sT_all[kGeoThreads / 32][kGeoFrames][12];
    __shared__ GeoStage stages[kGeoThreads / 32];
    constexpr int kWarps = kGeoThreads / 32;
    const int tid = threadIdx.x, lane = tid & 31;
    GeoStage &stage = stages[tid >> 5];
    double (*sT)[12] = sT_all[tid >> 5];
    const unsigned n_tiles = (unsigned)((a.n_vertices + 31) / 32);                          // warp tiles
    const unsigned n_chunks8 = (unsigned)((a.n_frames + kGeoFrames - 1) / kGeoFrames);
    const int kUnitFrames = a.unit_frames;                                                    // (runtime: 1, 2, 4 or 8)
    const unsigned kUnitSplit = (unsigned)(kGeoFrames / kUnitFrames);
    const unsigned n_chunks = (unsigned)((a.n_frames + kUnitFrames - 1) / kUnitFrames);
    const bool want_exact = DEBUG && a.vu_dense != nullptr;
    if (BINNED && lane == 0) stage.count = 0;
    __syncwarp();
    pdl_wait();
    pdl_trigger();
    const unsigned n_work = a.worklist ? *a.n_live * (8u * kUnitSplit) : n_tiles * n_chunks;   // (< 2^32: checked by the host)
    const unsigned n_warps = gridDim.x * kWarps;
    const unsigned first_dynamic = min(a.static_units, n_work);                              // units [0, first_dynamic) are dealt round-robin
    const bool use_table = a.cam_table != nullptr && *reinterpret_cast<const unsigned long long *>(a.cam_table) == a.tab_signature;
    const unsigned char *table = a.cam_table + CAMA_CAMERA_TABLE_HEADER;
    const unsigned all_cams = (1u << a.n_cams) - 1u;
    unsigned w = blockIdx.x * kWarps + (tid >> 5);
    if (w >= first_dynamic) {                                                                // (more warps than static units)
        unsigned c0 = 0;
        if (lane == 0) c0 = atomicAdd(a.geo_counter, 1u);
        w = first_dynamic + __shfl_sync(kFull, c0, 0);
    }
    while (w < n_work) {
        const bool next_static = w + n_warps < first_dynamic;               // (uniform)
        unsigned claim = 0;
        if (!next_static && lane == 0) claim = atomicAdd(a.geo_counter, 1u);   // the unit after this one (used at the bottom of the loop)
        unsigned tile;
        int f0;
        unsigned frame_mask = (1u << kUnitFrames) - 1u;
        if (a.worklist) {
            const unsigned long long e = a.worklist[w / (8u * kUnitSplit)];
            const unsigned rem = w % (8u * kUnitSplit), unit8 = (unsigned)(e >> 8);
            tile = unit8 / n_chunks8 * 8u + rem / kUnitSplit;              // (a 256-vertex tile = 8 warp tiles)
            f0 = (int)(unit8 % n_chunks8) * kGeoFrames + (int)(rem % kUnitSplit) * kUnitFrames;
            frame_mask &= (unsigned)(e & 0xffu) >> ((rem % kUnitSplit) * kUnitFrames);
        } else {
            tile = w / n_chunks;
            f0 = (int)(w % n_chunks) * kUnitFrames;
        }
        const int nf = min(kUnitFrames, a.n_frames - f0);
        if (tile < n_tiles && nf > 0 && frame_mask) {
            // which of the unit's frames can the warp's 32 vertices matter to?  lane = frame, poses straight from global
            // memory (L1): seven units in ten end here, and only the others stage their poses in shared memory
            const double *bounds = a.warp_bounds ? a.warp_bounds + (size_t)tile * 6 : (!a.worklist && a.tile_bounds) ? a.tile_bounds + (size_t)(tile >> 3) * 6 : nullptr;
            if (bounds) frame_mask &= __ballot_sync(kFull, lane < nf && tile_may_survive(bounds, a.w2c64 + (size_t)(f0 + (lane & (kGeoFrames - 1))) * 12, cams.box));
            frame_mask &= (1u << nf) - 1u;
        } else {
            frame_mask = 0u;
        }
        if (frame_mask) {
            __syncwarp();
            for (int i = lane; i < nf * 12; i += 32) (&sT[0][0])[i] = a.w2c64[(size_t)f0 * 12 + i];
            __syncwarp();
            const long long n = (long long)tile * 32 + lane;
            double vx, vy, vz;
            int ord;
            const bool valid = load_vertex<LAYOUT>(a, n, vx, vy, vz, ord);
            for (unsigned left = frame_mask; left; left &= left - 1u) {       // (uniform over the warp)
                const int fi = __ffs(left) - 1;
                const int f = f0 + fi;
                const double *T = sT[fi];
                // reference cama/dataset.py:99-105: world -> chassis, then the crop box
                const double cx = affine_row(T, vx, vy, vz);
                const double cy = affine_row(T + 4, vx, vy, vz);
                const double cz = affine_row(T + 8, vx, vy, vz);
                const bool alive = valid && in_box(cams.box, cx, cy, cz);
                if (__any_sync(kFull, alive)) {
                    if (DEBUG && alive && a.crop_counts) atomicAdd(&a.crop_counts[(size_t)f * a.n_instances + ord], 1);
                    unsigned lane_cams = alive ? all_cams : 0u;               // cameras that can see this lane's point at all
                    if (use_table && alive) {
                        const int ix = min(max((int)((float)(cx - a.tab_x0) * a.tab_inv_sx), 0), a.tab_nx - 1);
                        const int iy = min(max((int)((float)(cy - a.tab_y0) * a.tab_inv_sy), 0), a.tab_ny - 1);
                        lane_cams = __ldg(table + iy * a.tab_nx + ix);
                    }
                    double qx = 0.0, qy = 0.0, qz = 1.0;            // (lanes that are no candidate keep whatever the last camera left: masked)
#pragma unroll 1
                    for (unsigned cams_left = __reduce_or_sync(kFull, lane_cams) & all_cams; cams_left; cams_left &= cams_left - 1u) {
                        const int c = __ffs(cams_left) - 1;         // (uniform)
                        const bool cand = ((lane_cams >> c) & 1u) && camera_candidate<PINHOLE>(cams, c, cx, cy, cz, a.width, a.height, qx, qy, qz);
                        if (!__any_sync(kFull, cand)) continue;
                        int vi = 0, ui = 0;
                        double v = 0.0, u = 0.0;
                        const bool vis = candidate_pixel(cand, qx, qy, qz, a.width, a.height, want_exact, vi, ui, v, u);
                        if (!BINNED || DEBUG || __any_sync(kFull, vis)) emit_centre<BINNED, DEBUG>(a, stage, f, c, vis, vi, ui, v, u, ord, n);
                    }
                }
                if (BINNED && stage.count >= (unsigned)kStageFlush) stage_flush(a, stage);     // (warp-uniform)
            }
        }
        w = next_static ? w + n_warps : first_dynamic + __shfl_sync(kFull, claim, 0);
    }
    if (BINNED && stage.count > 0u) stage_flush(a, stage);
}

// ------------------------------------------------------------------------------------------------ PLANE raster
__global__ void __launch_bounds__(256) plane_raster_kernel(const unsigned *__restrict__ plane, const unsigned *__restrict__ lut,
                                                          const uint8_t *bg, uint8_t *frames, int height, int width, long long n_images) {
    const long long p = (long long)blockIdx.x * 256 + threadIdx.x;
    const long long per = (long long)height * width;
    if (p >= n_images * per) return;
    const long long img = p / per;
    const int y = (int)((p % per) / width), x = (int)(p % width);
    const unsigned *pl = plane + img * per;
    unsigned m = 0;
#pragma unroll
    for (int dy = -2; dy <= 2; ++dy) {
        const int yy = y + dy;
        if (yy < 0 || yy >= height) continue;
        const int span = 2 - (dy < 0 ? -dy : dy);
        for (int dx = -span; dx <= span; ++dx) {
            const int xx = x + dx;
            if (xx < 0 || xx >= width) continue;
            m = max(m, pl[(size_t)yy * width + xx]);
        }
    }
    uint8_t *o = frames + 3 * p;
    if (m) {
        const unsigned c = lut[m];
        o[0] = (uint8_t)c; o[1] = (uint8_t)(c >> 8); o[2] = (uint8_t)(c >> 16);
    } else if (bg) {
        const uint8_t *b = bg + 3 * p;
        const uint8_t b0 = b[0], b1 = b[1], b2 = b[2];
        o[0] = b0; o[1] = b1; o[2] = b2;
    } else {
        o[0] = 0; o[1] = 0; o[2] = 0;
    }
}

// ------------------------------------------------------------------------------------------------ BINNED: work lists of the raster
struct ClipStatsDev {
    unsigned long long records_total;      // appended to all lists (attempted)
    unsigned long long list_max;           // the fullest list (attempted)
    unsigned overflow;
    unsigned groups_used;                  // written by prep: statistics blocks (frame groups) the render that owns this block used
};
static_assert(sizeof(ClipStatsDev) == 24, "prep_kernel writes groups_used as the last 32-bit word");

// One thread per raster work item = (frame, camera, band).  The items are sorted into four work lists by the weight
// of their band group's record list: >= kHeavyBand / >= kMediumBand (per band of the group) / >= 1 records (claimed dynamically by the raster CTAs, in
// this order) and empty ones (a band of an empty group has nothing to draw: its zeros are bulk stores).  Within a list
// the order is whatever the warps' atomics make it.  The first group-count threads also fold the statistics: total
// and largest list length, overflow.
constexpr unsigned kHeavyBand = 2048, kMediumBand = 256;       // records per band of the group
__global__ void __launch_bounds__(256) band_classify_kernel(const unsigned *__restrict__ cursor, int n_lists, int n_items, int n_bands, int group_bands,
                                                            int n_groups, unsigned list_cap, ClipStatsDev *__restrict__ stats,
                                                            unsigned *__restrict__ lists /* [4][n_items] */, unsigned *__restrict__ list_counts /* [4] */) {
    const int i = blockIdx.x * 256 + threadIdx.x, lane = threadIdx.x & 31;
    pdl_wait();
    pdl_trigger();
    int cls = -1;
    if (i < n_items) {
        const int image = i / n_bands, band = i - image * n_bands;
        const unsigned n = cursor[image * n_groups + band / group_bands];
        cls = n == 0u ? 3 : n >= kHeavyBand * (unsigned)group_bands ? 0 : n >= kMediumBand * (unsigned)group_bands ? 1 : 2;
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        const unsigned m = __ballot_sync(kFull, cls == j);
        if (m == 0u) continue;
        unsigned base = 0;
        if (lane == __ffs(m) - 1) base = atomicAdd(&list_counts[j], (unsigned)__popc(m));
        base = __shfl_sync(kFull, base, __ffs(m) - 1);
        if (cls == j) lists[(size_t)j * n_items + base + __popc(m & ((1u << lane) - 1u))] = (unsigned)i;
    }
    // statistics over the lists (threads 0 .. n_lists-1), one atomic per warp and quantity
    unsigned long long mine = i < n_lists ? cursor[i] : 0ull;
    unsigned long long total = mine, most = mine;
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) {
        total += __shfl_xor_sync(kFull, total, d);
        const unsigned long long o = __shfl_xor_sync(kFull, most, d);
        most = o > most ? o : most;
    }
    if (lane == 0 && total) {
        atomicAdd(&stats->records_total, total);
        atomicMax(&stats->list_max, most);
        if (most > list_cap) atomicOr(&stats->overflow, 1u);
    }
}

// ------------------------------------------------------------------------------------------------ BINNED raster
__device__ __forceinline__ void fence_proxy_async_shared() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void bulk_store_shared_to_global(void *gptr, const void *sptr, unsigned bytes) {
    const unsigned s = (unsigned)__cvta_generic_to_shared(sptr);
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(gptr), "r"(s), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_commit_group() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void bulk_wait_group_read() { asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory"); }

__device__ __forceinline__ unsigned max3_u16x2(unsigned a, unsigned b, unsigned c) { return __vimax3_u16x2(a, b, c); }

// max of a 16-bit lane inside shared memory (ordinals of different instances race for a pixel)
__device__ __forceinline__ void smem_max_u16(unsigned short *plane, unsigned pix, unsigned val) {
    unsigned *word = reinterpret_cast<unsigned *>(plane) + (pix >> 1);
    const unsigned sh = (pix & 1u) * 16u;
    unsigned old = *reinterpret_cast<volatile unsigned *>(word);
    while (((old >> sh) & 0xffffu) < val) {
        const unsigned want = (old & ~(0xffffu << sh)) | (val << sh);
        const unsigned prev = atomicCAS(word, old, want);
        if (prev == old) break;
        old = prev;
    }
}

// CAMA_RASTER_DEBUG & 32: every raster CTA logs {start, end (globaltimer ns), active bands processed} here
// (read with cama_debug_raster_timeline; a tuning aid, not part of the ABI header)
__device__ unsigned long long g_raster_timeline[3 * 4096];
__device__ __forceinline__ unsigned long long global_ns() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}

constexpr int kRasterThreads = 128;            // compute threads of a raster CTA
constexpr int kRasterWarps = kRasterThreads / 32;
constexpr int kRasterBlock = kRasterThreads;
constexpr int kSegPx = 64;                     // pixels of one cell: 8 lanes x 8 px, so a warp computes four cells at a time
constexpr int kMaxSegs = 32;                   // width <= 2048
constexpr int kHitBytes = 4 * kMaxSegs;        // per-segment hit masks
constexpr int kZeroRows = 2;                   // image rows of zeros kept in shared memory (source of dark output)
constexpr int kMaxPlaneRows = 32;              // band_rows + 4: one bit per plane row in a hit mask

struct RasterArgs {
    int n_items;                           // F*C*NB
    int n_bands, band_rows, height, width, n_instances;
    int x_bits;                            // record = ord1 : 16 | plane row : 16 - x_bits | x : x_bits
    int n_strips;                          // 64-px segments per row
    int debug;                             // CAMA_RASTER_DEBUG experiments: 2 = no colour lookup
    int image_base;                        // sparse output: index of the first (frame, camera) image of this launch inside the clip
    int group_bands, n_groups;             // bands per band group, groups per image
    unsigned list_cap;
    const unsigned *cursor;                // [n_images * n_groups] records appended to each list
    const unsigned *records;               // [n_images * n_groups][list_cap]
    const unsigned *lut;
    const uint8_t *bg;
    uint8_t *frames;
    const unsigned *lists;                 // [4][n_items]: bands whose list holds >= 2048 / >= 256 / >= 1 records (claimed dynamically, in this order) | bands with an empty list (dealt round-robin)
    const unsigned *list_counts;           // [4]
    unsigned *work_counter;                // claims of active buckets beyond the first three of every CTA
    unsigned *empty_counter;               // claims of the dynamically dealt empty buckets
    int dyn_empty_pct, dyn_empty_per_cta;  // empty buckets held back for the CTAs that finish their active ones early: share of all, cap per CTA
    void *ov_records;                      // MODE 2 / 3: sparse output records
    unsigned *ov_count;
    long long ov_cap;
    void *ov_mirrors[CAMA_MAX_PEERS];      // record arrays on peer GPUs that receive every flush too (frame-sharded clips)
    int ov_n_mirrors;
    // where image (frame f, camera c) lives in `frames` / `bg`: f * frame_stride + cam_offset[c], rows `pitch` bytes apart.
    // Plain [F,C,H,W,3] frames: frame_stride = C*H*W*3, cam_offset[c] = c*H*W*3, pitch = W*3; the camera mosaic of
    // VideoGenerator.concate_image (cama/tools.py:22-25): frame_stride = rows*H*cols*W*3, pitch = cols*W*3.
    int n_cams;
    unsigned pitch;
    unsigned long long image_bytes;        // H*W*3
    unsigned long long frame_stride;
    unsigned long long cam_offset[CAMA_MAX_CAMERAS];
};

// Sparse output (MODE 2: 32-byte BGR records, MODE 3: 12-byte palette records): lit 8-pixel chunks are
// staged per warp and appended to the global record list >= kOvFlush at a time (one returning atomic
// per flush).  RW = 32-bit words per record.
constexpr int kOvFlush = 64;
constexpr int kOvCap = kOvFlush + 32;
struct OvStage {
    unsigned words[kOvCap * 8];
    unsigned count;
    unsigned pad[3];
};
struct OvSink {
    OvStage *st;
    unsigned *records;
    unsigned *count;
    long long cap;
    const RasterArgs *args;                // (mirrors)
};

template <int RW>
__device__ __forceinline__ void ov_flush(const OvSink &o) {
    const int lane = threadIdx.x & 31;
    const unsigned n = o.st->count;
    unsigned base = 0;
    if (lane == 0) base = atomicAdd(o.count, n);
    const long long b = __shfl_sync(kFull, base, 0);
    const long long room = o.cap > b ? o.cap - b : 0;                  // records that still fit
    const unsigned n_words = (unsigned)min((long long)n, room) * RW;
    for (unsigned i = lane; i < n_words; i += 32) o.records[b * RW + i] = o.st->words[i];
    // frame-sharded clips: the same words to the same place in every peer's mailbox (stores over NVLink; the slot's
    // count and step number follow in cama_peer_publish, after this kernel)
    for (int m = 0; m < o.args->ov_n_mirrors; ++m) {
        unsigned *dst = static_cast<unsigned *>(o.args->ov_mirrors[m]) + b * RW;
        for (unsigned i = lane; i < n_words; i += 32) dst[i] = o.st->words[i];
    }
    __syncwarp();
    if (lane == 0) o.st->count = 0;
    __syncwarp();
}

// warp-collective: every lane calls it; lanes with `lit` contribute one record of RW words
template <int RW>
__device__ __forceinline__ void ov_append(const OvSink &o, bool lit, const unsigned (&rec)[RW]) {
    const unsigned bal = __ballot_sync(kFull, lit);
    if (bal == 0) return;
    const int lane = threadIdx.x & 31;
    const unsigned slot = o.st->count + __popc(bal & ((1u << lane) - 1u));
    if (lit) {
#pragma unroll
        for (int k = 0; k < RW; ++k) o.st->words[slot * RW + k] = rec[k];
    }
    __syncwarp();
    if (lane == 0) o.st->count += (unsigned)__popc(bal);
    __syncwarp();
    if (o.st->count >= (unsigned)kOvFlush) ov_flush<RW>(o);
}

// 8 packed 24-bit values -> the 24 bytes of 8 BGR pixels (6 words), one byte-permute per word
__device__ __forceinline__ void pack8x24(const unsigned (&c)[8], unsigned (&w)[6]) {
    w[0] = __byte_perm(c[0], c[1], 0x4210); w[1] = __byte_perm(c[1], c[2], 0x5421); w[2] = __byte_perm(c[2], c[3], 0x6542);
    w[3] = __byte_perm(c[4], c[5], 0x4210); w[4] = __byte_perm(c[5], c[6], 0x5421); w[5] = __byte_perm(c[6], c[7], 0x6542);
}

// Colours of the 8 pixels of a lane (ids packed as u16x2 in m[4]; lut[0] == 0) into w[6].
// MODE 0: unlit pixels are black.  MODE 1: unlit pixels keep the background bytes already in w[].
template <int MODE>
__device__ __forceinline__ void colour8(const unsigned *__restrict__ lut, const unsigned (&m)[4], unsigned (&w)[6]) {
    unsigned c[8];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        c[2 * k] = __ldg(lut + (m[k] & 0xffffu));
        c[2 * k + 1] = __ldg(lut + (m[k] >> 16));
    }
    unsigned v[6];
    pack8x24(c, v);
    if (MODE == 0) {
#pragma unroll
        for (int k = 0; k < 6; ++k) w[k] = v[k];
    } else {
        unsigned lit[8], keep[6];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            lit[2 * k] = (m[k] & 0xffffu) ? 0xffffffu : 0u;
            lit[2 * k + 1] = (m[k] >> 16) ? 0xffffffu : 0u;
        }
        pack8x24(lit, keep);
#pragma unroll
        for (int k = 0; k < 6; ++k) w[k] = (w[k] & ~keep[k]) | v[k];
    }
}

template <int MODE>
__device__ __forceinline__ void store_row(const unsigned *__restrict__ lut, const unsigned (&m)[4], bool lane_on, uint8_t *out_px, const uint8_t *bg_px,
                                          const OvSink &ov, unsigned chunk) {
    if (MODE == 2) {                           // sparse output: a record per lit chunk, nothing dense
        const bool lit = lane_on && (m[0] | m[1] | m[2] | m[3]) != 0u;
        unsigned w[6] = {0u, 0u, 0u, 0u, 0u, 0u};
        unsigned mask = 0u;
        if (lit) {
            colour8<0>(lut, m, w);
#pragma unroll
            for (int k = 0; k < 4; ++k) mask |= ((m[k] & 0xffffu) ? 1u << (2 * k) : 0u) | ((m[k] >> 16) ? 2u << (2 * k) : 0u);
        }
        const unsigned rec[8] = {chunk, mask, w[0], w[1], w[2], w[3], w[4], w[5]};
        ov_append<8>(ov, lit, rec);
        return;
    }
    if (MODE == 3) {                           // sparse output, palette form: one byte per pixel (lut[id] >> 24, 0 = not painted)
        const bool lit = lane_on && (m[0] | m[1] | m[2] | m[3]) != 0u;
        unsigned rec[3] = {chunk, 0u, 0u};
        if (lit) {
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const unsigned p_lo = __ldg(lut + (m[k] & 0xffffu)) >> 24, p_hi = __ldg(lut + (m[k] >> 16)) >> 24;
                rec[1 + (k >> 1)] |= (p_lo | (p_hi << 8)) << (16 * (k & 1));
            }
        }
        ov_append<3>(ov, lit, rec);
        return;
    }
    if (!lane_on) return;
    unsigned w[6] = {0u, 0u, 0u, 0u, 0u, 0u};
    if (MODE == 1) {
        const uint2 *b = reinterpret_cast<const uint2 *>(bg_px);
        const uint2 b0 = b[0], b1 = b[1], b2 = b[2];
        w[0] = b0.x; w[1] = b0.y; w[2] = b1.x; w[3] = b1.y; w[4] = b2.x; w[5] = b2.y;
    }
    if (m[0] | m[1] | m[2] | m[3]) {
        if (lut) colour8<MODE>(lut, m, w);
        else { w[0] = m[0]; w[1] = m[1]; w[2] = m[2]; w[3] = m[3]; }
    }
    uint2 *d = reinterpret_cast<uint2 *>(out_px);
    d[0] = make_uint2(w[0], w[1]); d[1] = make_uint2(w[2], w[3]); d[2] = make_uint2(w[4], w[5]);
}

// 3-wide (t) and, optionally, 5-wide (f) horizontal max of one plane row held as 4 u16x2 words + halo words
__device__ __forceinline__ void row_max(const uint4 &A, unsigned L, unsigned R, uint4 &t, uint4 *f) {
    const unsigned S0 = __byte_perm(L, A.x, 0x5432);     // (p-1, p0)
    const unsigned S1 = __byte_perm(A.x, A.y, 0x5432);   // (p1, p2)
    const unsigned S2 = __byte_perm(A.y, A.z, 0x5432);   // (p3, p4)
    const unsigned S3 = __byte_perm(A.z, A.w, 0x5432);   // (p5, p6)
    const unsigned S4 = __byte_perm(A.w, R, 0x5432);     // (p7, p8)
    t.x = max3_u16x2(S0, A.x, S1); t.y = max3_u16x2(S1, A.y, S2);
    t.z = max3_u16x2(S2, A.z, S3); t.w = max3_u16x2(S3, A.w, S4);
    if (f) {
        f->x = max3_u16x2(t.x, L, A.y);   f->y = max3_u16x2(t.y, A.x, A.z);
        f->z = max3_u16x2(t.z, A.y, A.w); f->w = max3_u16x2(t.w, A.z, R);
    }
}

// One warp, one cell = output rows y, y+1 (band rows) of one 256-pixel strip; lane = 8 pixels.
// Stateless: the six plane rows y .. y+5 are read from shared memory (14 independent loads).
//   out(y) = max(raw[y-2], h3[y-1], h5[y], h3[y+1], raw[y+2])     (the 13-px L1 ball; h3/h5 = 3/5-wide row max)
template <int MODE>
__device__ __forceinline__ void raster_cell(const unsigned short *plane, const unsigned *__restrict__ lut, int W, int xc, bool lane_on,
                                            int y, bool two, uint8_t *out_px, const uint8_t *bg_px, unsigned pitch,
                                            const OvSink &ov, unsigned chunk) {
    const unsigned short *row = plane + (size_t)y * W + xc;
    const bool has_l = xc > 0, has_r = xc + 8 < W;
    uint4 A[6];
    unsigned L[4], R[4];
#pragma unroll
    for (int r = 0; r < 6; ++r) A[r] = *reinterpret_cast<const uint4 *>(row + (size_t)r * W);
#pragma unroll
    for (int r = 0; r < 4; ++r) {
        L[r] = has_l ? *reinterpret_cast<const unsigned *>(row + (size_t)(r + 1) * W - 2) : 0u;
        R[r] = has_r ? *reinterpret_cast<const unsigned *>(row + (size_t)(r + 1) * W + 8) : 0u;
    }
    uint4 t1, t2, t3, t4, f2, f3;
    row_max(A[1], L[0], R[0], t1, nullptr);
    row_max(A[2], L[1], R[1], t2, &f2);
    row_max(A[3], L[2], R[2], t3, &f3);
    row_max(A[4], L[3], R[3], t4, nullptr);
    unsigned m[4];
    m[0] = max3_u16x2(max3_u16x2(A[0].x, t1.x, f2.x), t3.x, A[4].x);
    m[1] = max3_u16x2(max3_u16x2(A[0].y, t1.y, f2.y), t3.y, A[4].y);
    m[2] = max3_u16x2(max3_u16x2(A[0].z, t1.z, f2.z), t3.z, A[4].z);
    m[3] = max3_u16x2(max3_u16x2(A[0].w, t1.w, f2.w), t3.w, A[4].w);
    store_row<MODE>(lut, m, lane_on, out_px, bg_px, ov, chunk);
    // (`two` differs between the lane groups of a warp only in a band with an odd row count; the sparse modes append
    // their records warp-collectively, so every lane goes through the second row with its own predicate)
    m[0] = max3_u16x2(max3_u16x2(A[1].x, t2.x, f3.x), t4.x, A[5].x);
    m[1] = max3_u16x2(max3_u16x2(A[1].y, t2.y, f3.y), t4.y, A[5].y);
    m[2] = max3_u16x2(max3_u16x2(A[1].z, t2.z, f3.z), t4.z, A[5].z);
    m[3] = max3_u16x2(max3_u16x2(A[1].w, t2.w, f3.w), t4.w, A[5].w);
    store_row<MODE>(lut, m, lane_on && two, out_px + pitch, MODE == 1 ? bg_px + pitch : nullptr, ov, chunk + (unsigned)(W >> 3));
}

// One work item = one (frame, camera, band).  Shared memory: uint16 centre plane [(band_rows+5)][W] | hit
// masks (bit r of hits[g]: plane row r has a centre within 2 px of the 64-px segment g) | kZeroRows image
// rows of zeros.  Per item: scatter the bucket's records into the plane; then every warp lists the lit
// cells of the band (cell = 2 output rows x 64 px whose 6-row window has a hit: one ballot per row pair),
// keeps the ones that fall to it, and computes four of them per pass (8 lanes x 8 px each) straight from the
// plane, storing from registers; the dark stretches of a row pair go out as bulk stores from the shared
// zeros, one per run of dark segments (MODE 0).  Only ~6 % of the pixels of a clip are painted: with
// 256-px cells (one warp strip) 18 % of the cells were lit and computed, with 64-px cells far fewer.
// Bands without a record cost one thread a few bulk stores.
#ifndef CAMA_RASTER_MINB
#define CAMA_RASTER_MINB 4
#endif
template <int MODE>
__global__ void __launch_bounds__(kRasterBlock, CAMA_RASTER_MINB) binned_raster_kernel(const RasterArgs a) {
    extern __shared__ __align__(128) unsigned char smem[];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int W = a.width;
    const int plane_rows = a.band_rows + 4;
    const unsigned plane_bytes = (unsigned)((plane_rows + 1) * W) * 2u;      // one spare (always zero) row: the odd last row reads it
    const unsigned row_bytes = (unsigned)W * 3u, pitch = a.pitch;               // bytes of a row / distance between rows (equal for plain frames)
    const bool contiguous = pitch == row_bytes;
    const unsigned zero_bytes = (unsigned)kZeroRows * row_bytes;
    // (plain frames: image i starts at i * H * W * 3; the division and the per-camera table are only needed for the mosaic)
    auto image_offset = [&](int image) -> size_t {
        return contiguous ? (size_t)image * a.image_bytes : (size_t)(image / a.n_cams) * a.frame_stride + a.cam_offset[image % a.n_cams];
    };
    unsigned short *plane = reinterpret_cast<unsigned short *>(smem);
    unsigned *hits = reinterpret_cast<unsigned *>(smem + plane_bytes);
    unsigned char *zeros = smem + plane_bytes + kHitBytes;
    OvSink ov{};
    if (MODE >= 2) {                                                     // per-warp staging of sparse records, after the zeros
        ov.st = reinterpret_cast<OvStage *>(zeros + zero_bytes) + warp;
        ov.records = reinterpret_cast<unsigned *>(a.ov_records); ov.count = a.ov_count; ov.cap = a.ov_cap; ov.args = &a;
        if (lane == 0) ov.st->count = 0;
    }
    const bool inplace = MODE == 1 && a.bg == a.frames;
    const unsigned x_mask = (1u << a.x_bits) - 1u;
    const int n_strips = a.n_strips;
    {   // plane and hit masks start out (and are kept) all-zero between items; the zero rows stay zero
        uint4 *p4 = reinterpret_cast<uint4 *>(smem);
        const int n16 = (int)((plane_bytes + kHitBytes + zero_bytes) >> 4);
        for (int i = tid; i < n16; i += kRasterBlock) p4[i] = make_uint4(0u, 0u, 0u, 0u);
    }
    if (MODE == 0) fence_proxy_async_shared();   // the zeros are read by bulk stores
    __syncthreads();
    pdl_wait();                                  // everything above overlapped the tail of the previous kernel
    pdl_trigger();
    if ((a.debug & 32) && tid == 0 && blockIdx.x < 4096) g_raster_timeline[3 * blockIdx.x] = global_ns();
    auto sync_compute = [] { asm volatile("bar.sync 1, %0;" ::"n"(kRasterThreads) : "memory"); };

    // a record of the band group's list -> this band's plane, if its row is one of the band's (row_off = the band's first
    // row inside the group; rows are stored + 2, so the band's window [row_off - 2, row_off + rows + 2) is plane rows 0 ..)
    auto scatter = [&](unsigned rec, unsigned row_off) {
        // (x and ordinal checks: the fetch pads with 0xffffffff, and a list past its capacity holds stale records)
        const unsigned x = rec & x_mask, row = ((rec & 0xffffu) >> a.x_bits) - row_off, ord1 = rec >> 16;
        if (row < (unsigned)plane_rows && x < (unsigned)W && ord1 <= (unsigned)a.n_instances) {
            smem_max_u16(plane, row * (unsigned)W + x, ord1);
            const unsigned bit = 1u << row, g = x / kSegPx, off = x % kSegPx;
            atomicOr(&hits[g], bit);
            if (off < 2u && g > 0u) atomicOr(&hits[g - 1], bit);
            if (off >= kSegPx - 2u && g + 1u < (unsigned)n_strips) atomicOr(&hits[g + 1], bit);
        }
    };
    auto fetch = [&](long long begin, long long end, unsigned (&rec)[4]) {
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const long long i = begin + tid + q * kRasterThreads;
            rec[q] = i < end ? __ldg(a.records + i) : 0xffffffffu;
        }
    };

    // Work distribution.  Buckets with records ("active") come from a list ordered heaviest class first
    // and are claimed dynamically, two claims ahead of use, so the claim, the bucket bounds and the first
    // records of the next bucket are in flight while the current one is processed.  Buckets without
    // records are dealt round-robin; their zeros are issued by thread 0 a few at a time between active
    // buckets, so that the store queue stays fed while the warps compute.
    __shared__ int s_claim[2];
    __shared__ unsigned short s_cells[kRasterWarps][kMaxPlaneRows / 2 * kMaxSegs / kRasterWarps + 8];    // lit cells of the band that fall to each warp: pair << 8 | segment
    const unsigned G = gridDim.x;
    const unsigned n0 = a.list_counts[0], n1 = a.list_counts[1], n2 = a.list_counts[2], n_emp = a.list_counts[3];
    const unsigned n_act = n0 + n1 + n2;
    const unsigned *empty_list = a.lists + (size_t)3 * a.n_items;
    const bool copy_empties = MODE == 1 && !inplace;          // out-of-place composite: empty bands are copied by the active path
    const unsigned n_work = n_act + (copy_empties ? n_emp : 0u);
    auto item_at = [&](unsigned pos) -> int {
        if (pos < n0) return (int)a.lists[pos];
        if (pos < n0 + n1) return (int)a.lists[(size_t)a.n_items + (pos - n0)];
        if (pos < n_act) return (int)a.lists[(size_t)2 * a.n_items + (pos - n0 - n1)];
        if (pos < n_work) return (int)empty_list[pos - n_act];
        return -1;
    };
    auto bounds = [&](int item, long long &blo, long long &bhi) {
        blo = bhi = 0;
        if (item >= 0) {                       // the list of the item's band group
            const int image = item / a.n_bands, list = image * a.n_groups + (item - image * a.n_bands) / a.group_bands;
            blo = (long long)list * a.list_cap;
            bhi = blo + min(a.cursor[list], a.list_cap);
        }
    };
    // The last dyn_empty_pct % of the empty list are not dealt: a CTA that has finished its active buckets claims
    // them two at a time until they are gone, so the CTAs that drew light buckets take store work off the ones
    // still computing (CTA end times spread 47..73 us around a mean of 60 with every empty dealt statically).
    // (the held-back pool only has to cover the spread of the CTAs' finishing times: at most dyn_empty_per_cta buckets per CTA)
    const unsigned n_dyn = MODE == 0 ? min((unsigned)((unsigned long long)n_emp * (unsigned)a.dyn_empty_pct / 100u), G * (unsigned)a.dyn_empty_per_cta) : 0u;
    const unsigned n_static = n_emp - n_dyn;
    unsigned e_pos = blockIdx.x;                               // (thread 0) next position of this CTA in the empty list
    const unsigned my_emp = e_pos < n_static ? (n_static - e_pos + G - 1u) / G : 0u;
    const unsigned my_act = max(1u, (n_act + G - 1u) / G);
    const unsigned empties_per_active = (my_emp + my_act - 1u) / my_act;
    auto store_empty = [&](unsigned pos) {
            const int item = (int)empty_list[pos];
            const int y_first = (item % a.n_bands) * a.band_rows;
            const int rows_out = min(a.band_rows, a.height - y_first);
            uint8_t *out_base = a.frames + image_offset(item / a.n_bands) + (size_t)y_first * pitch;
            if (contiguous) {
                unsigned left = (unsigned)rows_out * row_bytes, off = 0;
                while (left) {
                    const unsigned n = min(left, zero_bytes);
                    bulk_store_shared_to_global(out_base + off, zeros, n);
                    off += n; left -= n;
                }
            } else {                                   // (mosaic: the rows of an image are not adjacent)
                for (int y = 0; y < rows_out; ++y) bulk_store_shared_to_global(out_base + (size_t)y * pitch, zeros, row_bytes);
            }
            bulk_commit_group();
    };
    auto issue_empties = [&](unsigned k) {
        if (MODE != 0 || tid != 0) return;
        for (; k > 0u && e_pos < n_static; --k, e_pos += G) store_empty(e_pos);
    };

    unsigned a2 = blockIdx.x + 2u * G;
    int item0 = item_at(blockIdx.x), item1 = item_at(blockIdx.x + G);
    long long lo, hi, nlo, nhi;
    unsigned pre[4];
    bounds(item0, lo, hi);
    bounds(item1, nlo, nhi);
    fetch(lo, hi, pre);
    int n_done = 0;
    for (int it = 0; item0 >= 0; ++it, ++n_done) {
        int claim = 0;
        if (tid == 0) claim = (int)(3u * G + atomicAdd(a.work_counter, 1u));     // list position for three iterations from now
        const int item2 = item_at(a2);
        long long n2lo, n2hi;
        bounds(item2, n2lo, n2hi);
        unsigned pre_next[4];
        fetch(nlo, nhi, pre_next);         // (nlo == nhi == 0 past the last bucket: nothing is loaded)
        issue_empties(empties_per_active);
        {
            const int y_first = (item0 % a.n_bands) * a.band_rows;
            const int rows_out = min(a.band_rows, a.height - y_first);
            const size_t band_offset = image_offset(item0 / a.n_bands) + (size_t)y_first * pitch;
            uint8_t *out_base = a.frames + band_offset;
            const uint8_t *bg_base = MODE == 1 ? a.bg + band_offset : nullptr;
            const unsigned chunk_base = (unsigned)((((size_t)(a.image_base + item0 / a.n_bands) * a.height + y_first) * W) >> 3);   // MODE 2
            // 1. centres of this band (out of its group's list) -> plane (max ordinal per pixel) + hit masks
            const unsigned row_off = (unsigned)(((item0 % a.n_bands) % a.group_bands) * a.band_rows);
#pragma unroll
            for (int q = 0; q < 4; ++q) scatter(pre[q], row_off);
            // (16-byte fetches of four consecutive records per thread and a software-pipelined loop were both measured slower)
            for (long long base = lo + 4 * kRasterThreads; base < hi; base += 4 * kRasterThreads) {
                unsigned rec[4];
                fetch(base, hi, rec);
#pragma unroll
                for (int q = 0; q < 4; ++q) scatter(rec[q], row_off);
            }
            sync_compute();
            // 2. cells: dilation + colour + store.  Every warp walks the row pairs (lane = segment): a ballot gives the
            // pair's lit segments, lit cells are numbered in band order, batches of four go to the warps round-robin
            // and each warp notes its own; the dark runs of pair p are stored by warp p % 4.
            const int n_pairs = (rows_out + 1) >> 1;
            const bool all_cells = MODE == 1 && !inplace;                       // out-of-place composite: every cell is written
            const unsigned my_hits = lane < n_strips ? hits[lane] : 0u;
            const unsigned seg_all = n_strips >= 32 ? kFull : (1u << n_strips) - 1u;
            unsigned short *my_cells = s_cells[warp];
            unsigned n_lit = 0;                                                  // (warp-uniform)
            for (int pp = 0; pp < n_pairs; ++pp) {
                const int y = 2 * pp;
                const bool two = y + 1 < rows_out;
                const bool lit = lane < n_strips && (all_cells || ((my_hits >> y) & (two ? 0x3fu : 0x1fu)) != 0u);
                const unsigned m = __ballot_sync(kFull, lit);
                if (lit) {
                    const unsigned idx = n_lit + (unsigned)__popc(m & ((1u << lane) - 1u)), batch = idx >> 2;
                    if (batch % kRasterWarps == (unsigned)warp) my_cells[((batch / kRasterWarps) << 2) | (idx & 3u)] = (unsigned short)((pp << 8) | lane);
                }
                n_lit += (unsigned)__popc(m);
                if (MODE == 0 && pp % kRasterWarps == warp) {
                    const unsigned dark = ~m & seg_all;
                    if (((dark >> lane) & 1u) && (lane == 0 || !((dark >> (lane - 1)) & 1u))) {          // first segment of a dark run
                        const unsigned after = m & ~((2u << lane) - 1u);                                    // lit segments beyond it
                        const int x_a = lane * kSegPx, x_b = min((after ? __ffs(after) - 1 : n_strips) * kSegPx, W);
                        uint8_t *dst = out_base + (size_t)y * pitch + (size_t)x_a * 3;
                        const unsigned bytes = (unsigned)(x_b - x_a) * 3u;
                        if (two && bytes == row_bytes && contiguous) {
                            bulk_store_shared_to_global(dst, zeros, 2u * row_bytes);                        // dark across the width: both rows at once
                        } else {
                            bulk_store_shared_to_global(dst, zeros, bytes);
                            if (two) bulk_store_shared_to_global(dst + pitch, zeros, bytes);
                        }
                        bulk_commit_group();
                    }
                }
            }
            __syncwarp();
            const unsigned n_batches = (n_lit + 3u) >> 2;
            for (unsigned b = (unsigned)warp; b < n_batches; b += kRasterWarps) {
                const unsigned k = (b << 2) | (unsigned)(lane >> 3);           // the cell of this lane group, in band order
                const bool active = k < n_lit;
                const unsigned cell = active ? my_cells[((b / kRasterWarps) << 2) | (unsigned)(lane >> 3)] : 0u;
                const int y = 2 * (int)(cell >> 8);
                const bool two = y + 1 < rows_out;
                const int x0 = (int)(cell & 0xffu) * kSegPx + (lane & 7) * 8;
                const bool lane_on = active && x0 < W;
                const int xc = min(x0, W - 8);                                 // lanes past the edge recompute the last 8 px, store nothing
                raster_cell<MODE>(plane, (a.debug & 2) ? nullptr : a.lut, W, xc, lane_on, y, two, out_base + (size_t)y * pitch + (size_t)x0 * 3,
                                  MODE == 1 ? bg_base + (size_t)y * pitch + (size_t)x0 * 3 : nullptr, pitch,
                                  ov, chunk_base + (unsigned)((y * W + x0) >> 3));
            }
            sync_compute();
            // 3. restore the all-zero plane (clearing all of it costs fewer instructions than picking the rows hit)
            {
                uint4 *p4 = reinterpret_cast<uint4 *>(plane);
                const int n16 = (plane_rows * W) >> 3;
#pragma unroll 4
                for (int i = tid; i < n16; i += kRasterThreads) p4[i] = make_uint4(0u, 0u, 0u, 0u);
            }
            if (tid < kMaxSegs) hits[tid] = 0u;
        }
        // publish the claim (this barrier also orders the clean-up before the next scatter) and advance the pipeline
        if (tid == 0) s_claim[it & 1] = claim;
        __syncthreads();
        item0 = item1; lo = nlo; hi = nhi;
#pragma unroll
        for (int q = 0; q < 4; ++q) pre[q] = pre_next[q];
        item1 = item2; nlo = n2lo; nhi = n2hi;
        a2 = (unsigned)s_claim[it & 1];
    }
    issue_empties(0xffffffffu);
    if (MODE == 0 && tid == 0) {
        unsigned pos = n_static + atomicAdd(a.empty_counter, 2u);
        while (pos < n_emp) {
            const unsigned next = n_static + atomicAdd(a.empty_counter, 2u);      // (in flight while this pair is stored)
            store_empty(pos);
            if (pos + 1u < n_emp) store_empty(pos + 1u);
            pos = next;
        }
    }
    if ((a.debug & 32) && tid == 0 && blockIdx.x < 4096) {
        g_raster_timeline[3 * blockIdx.x + 1] = global_ns();
        g_raster_timeline[3 * blockIdx.x + 2] = (unsigned long long)n_done;
    }
    if (MODE == 2 && ov.st->count > 0u) ov_flush<8>(ov);
    if (MODE == 3 && ov.st->count > 0u) ov_flush<3>(ov);
    // the shared zeros must stay valid until the last bulk stores have read them
    if (MODE == 0 && lane == 0) bulk_wait_group_read<0>();
}

}  // namespace cama

using namespace cama;

namespace {

// Kernel launch with (optionally) programmatic stream serialisation, see pdl_wait().  CAMA_NO_PDL=1
// in the environment turns the attribute off (plain stream order), for A/B measurements.
bool pdl_enabled() {
    static const bool on = getenv("CAMA_NO_PDL") == nullptr;
    return on;
}

// Frames per group of the frame-group pipeline; CAMA_PIPE_FRAMES=<n> overrides (0 = off).  A multiple of the
// geometry kernel's 8-frame work unit.
int pipe_group_frames() {
    static const int frames = [] {
        const char *env = getenv("CAMA_PIPE_FRAMES");
        int f = env ? atoi(env) : CAMA_PIPE_FRAMES_DEFAULT;
        return f <= 0 ? 0 : (f + 7) / 8 * 8;
    }();
    return frames;
}

template <typename... KArgs, typename... Args>
cudaError_t launch_k(bool pdl, void (*kernel)(KArgs...), unsigned grid, unsigned block, size_t smem, cudaStream_t s, Args &&...args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(grid, 1, 1);
    cfg.blockDim = dim3(block, 1, 1);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = s;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = pdl ? 1 : 0;
    return cudaLaunchKernelEx(&cfg, kernel, KArgs(std::forward<Args>(args))...);
}

// Cell grid of the camera table over the crop box's x-y rectangle: cells of about a metre, at most 256 x 256; none
// for a degenerate or unbounded box.
bool table_geometry(const double *box, CamTable &tab) {
    const double rx = box[1] - box[0], ry = box[3] - box[2];
    if (!(rx > 0.0 && ry > 0.0 && rx <= 1e6 && ry <= 1e6)) return false;                  // (false for NaN / infinite boxes)
    tab.nx = (int)std::min<double>(kCamTableMaxDim, std::max(1.0, std::ceil(rx)));
    tab.ny = (int)std::min<double>(kCamTableMaxDim, std::max(1.0, std::ceil(ry)));
    tab.x0 = box[0]; tab.y0 = box[2];
    tab.sx = rx / tab.nx; tab.sy = ry / tab.ny;
    return true;
}

// FNV-1a over everything the table depends on: a table built for another rig / box / image size is ignored by the kernel
unsigned long long table_signature(const CamBlock &cams, int n_cams, int width, int height) {
    unsigned long long h = 1469598103934665603ull;
    auto mix = [&](const void *p, size_t n) {
        const unsigned char *b = static_cast<const unsigned char *>(p);
        for (size_t i = 0; i < n; ++i) { h ^= b[i]; h *= 1099511628211ull; }
    };
    mix(cams.E, sizeof(cams.E)); mix(cams.K, sizeof(cams.K)); mix(cams.box, sizeof(cams.box));
    const int dims[3] = {n_cams, width, height};
    mix(dims, sizeof(dims));
    return h | 1ull;                                           // (never 0: a zero-filled buffer is not a table)
}

void fill_cam_block(CamBlock &cams, int n_cams, const double *chassis2cam, const double *intrinsics, const double *crop_box, int width, int height) {
    for (int c = 0; c < CAMA_MAX_CAMERAS; ++c) {
        const bool live = c < n_cams;
        for (int i = 0; i < 12; ++i) cams.E[c][i] = live ? chassis2cam[16 * c + i] : 0.0;
        for (int i = 0; i < 9; ++i) cams.K[c][i] = live ? intrinsics[9 * c + i] : 0.0;
        cams.k_row2_is_001[c] = live && cams.K[c][6] == 0.0 && cams.K[c][7] == 0.0 && cams.K[c][8] == 1.0;
    }
    for (int i = 0; i < 6; ++i) cams.box[i] = crop_box[i];
    cams.wlim = (double)(width + 1);
    cams.hlim = (double)(height + 1);
    cams.all_pinhole = 1;
    for (int c = 0; c < n_cams; ++c) {
        const double *K = cams.K[c];
        if (!(cams.k_row2_is_001[c] && K[1] == 0.0 && K[3] == 0.0)) cams.all_pinhole = 0;
    }
}

// bytes from one frame of `frames` / `background` to the next: n_cams images, or one mosaic of rows x mosaic_cols tiles
size_t frame_stride_bytes(const cama_clip_desc *d) {
    const size_t image = (size_t)d->height * d->width * 3;
    if (d->mosaic_cols <= 0) return (size_t)d->n_cams * image;
    const int rows = (d->n_cams + d->mosaic_cols - 1) / d->mosaic_cols;
    return (size_t)rows * d->mosaic_cols * image;
}

struct ClipPlan {
    int mode;
    int band_rows, n_bands, x_bits;
    int group_bands, n_groups;      // bands per band group (one record list per (frame, camera, group)), groups per image
    int n_strips;
    long long cap;          // records per list
    int n_buckets;          // raster work items: frames x cameras x bands
    long long n_lists;
    size_t raster_smem;
    // workspace offsets
    size_t off_zero, zero_bytes;     // region cleared by prep each call: counters | list cursors (BINNED, lists inside the workspace)
    size_t off_counter, off_cursor, off_stats, off_w2c64, off_lut, off_records, off_plane, off_worklist, off_lists;
    long long geo_units;
    size_t total;
    // frame-group pipeline (BINNED): the clip is rendered as `groups` sub-clips of `group_frames` frames, each with
    // its own workspace slice of `group_stride` bytes (laid out by the plan of a `group_frames`-frame clip)
    int groups, group_frames;
    size_t group_stride;
};

constexpr int kRasterCtasPerSm = 4;
constexpr int kDynEmptyPerCta = 5;
constexpr int kDynEmptyPct = 60;                 // measured on config 2: 0 % 78.5 us, 40 % 74.4, 60 % 73.4, 80 % 75.5, 100 % 78.6 (CAMA_RASTER_DYN_EMPTY)
constexpr int kDefaultBandRows = 16;
// Bands that share one record list.  1: every band has its own list and reads nothing it does not draw.  Larger groups
// mean fewer, longer lists (less workspace, fewer halo duplicates) but every band scans its whole group: measured on
// config 2, raster 81 us with 1, 90 with 2, 110 with 3 (CAMA_GROUP_BANDS).
constexpr int kDefaultGroupBands = 1;
bool BinnedSite(const cama_clip_desc *d, const cama_ctx *ctx, long long units) { return d->tile_bounds && units >= (long long)ctx->sm_count * 64; }

template <bool BINNED>
cudaError_t launch_geometry(bool pdl, bool f32, bool debug, unsigned grid, cudaStream_t s, const ClipArgs &a, const CamBlock &cams) {
    if (f32 && !debug) {                                            // the production shape: float32 vertices, no per-instance outputs
        if (cams.all_pinhole) return launch_k(pdl, clip_geometry_kernel<CAMA_VERTEX_F32X4, BINNED, false, true>, grid, kGeoThreads, 0, s, a, cams);
        return launch_k(pdl, clip_geometry_kernel<CAMA_VERTEX_F32X4, BINNED, false, false>, grid, kGeoThreads, 0, s, a, cams);
    }
    if (f32) return launch_k(pdl, clip_geometry_kernel<CAMA_VERTEX_F32X4, BINNED, true, false>, grid, kGeoThreads, 0, s, a, cams);
    return debug ? launch_k(pdl, clip_geometry_kernel<CAMA_VERTEX_F64X3, BINNED, true, false>, grid, kGeoThreads, 0, s, a, cams)
                 : launch_k(pdl, clip_geometry_kernel<CAMA_VERTEX_F64X3, BINNED, false, false>, grid, kGeoThreads, 0, s, a, cams);
}

constexpr size_t kRasterSmemBudget = 56 * 1024;      // four CTAs per SM
constexpr size_t kRasterStageSmem = kHitBytes;   // + the plane + kZeroRows image rows

int make_plan(const cama_clip_desc *d, ClipPlan &p, bool allow_groups = true) {
    CAMA_REQUIRE(d, "desc is NULL");
    CAMA_REQUIRE(d->struct_bytes == sizeof(cama_clip_desc), "cama_clip_desc size mismatch: caller %u, library %zu", d->struct_bytes, sizeof(cama_clip_desc));
    CAMA_REQUIRE(d->n_frames >= 0 && d->n_cams > 0 && d->n_instances >= 0 && d->n_vertices >= 0, "negative size");
    CAMA_REQUIRE(d->height > 0 && d->width > 0, "bad image size");
    if (d->n_cams > CAMA_MAX_CAMERAS) return fail(CAMA_E_UNSUPPORTED, "at most %d cameras per call", CAMA_MAX_CAMERAS);
    CAMA_REQUIRE(d->vertex_layout == CAMA_VERTEX_F32X4 || d->vertex_layout == CAMA_VERTEX_F64X3, "bad vertex_layout");
    CAMA_REQUIRE((long long)d->n_frames * d->n_cams * d->height * d->width < (1ll << 40), "clip too large");
    CAMA_REQUIRE(((d->n_vertices + 31) / 32 + 8) * (((long long)d->n_frames + 7) / 8 * 8) < (1ll << 32), "too many (vertex tile, frame) units for 32-bit indices");
    const int W = d->width, H = d->height;
    // BINNED needs: 16-byte rows for the bulk stores, a 16-bit {plane row | x} pixel code with at least
    // 8 plane rows, 32 occupancy segments of 64 px, a 16-bit ordinal
    int band_rows = 0, x_bits = 0;
    bool binned_ok = (W % 16 == 0) && W <= 2048 && d->n_instances <= 65534;
    if (binned_ok) {
        while ((1 << x_bits) < W) ++x_bits;
        long long rows = ((long long)kRasterSmemBudget - (long long)kRasterStageSmem - (long long)kZeroRows * W * 3) / (2ll * W) - 5;
        rows = std::min<long long>(rows, (1ll << (16 - x_bits)) - 4);
        rows = std::min<long long>(rows, kMaxPlaneRows - 4);
        rows = std::min<long long>(rows, kDefaultBandRows);       // measured optimum on config 2 (profiles/)
        if (const char *env = getenv("CAMA_BAND_ROWS")) {        // tuning knob for experiments
            const long long want = atoll(env);
            if (want >= 1) rows = std::min(rows, want);
        }
        rows = std::min<long long>(rows, H);
        if (rows < 4 && rows < H) binned_ok = false;
        if (binned_ok) {                                          // even out the bands
            const long long nb = (H + rows - 1) / rows;
            rows = (H + nb - 1) / nb;
        }
        band_rows = (int)rows;
    }
    int mode = d->mode;
    if (mode == CAMA_CLIP_AUTO) mode = binned_ok ? CAMA_CLIP_BINNED : CAMA_CLIP_PLANE;
    CAMA_REQUIRE(mode == CAMA_CLIP_PLANE || mode == CAMA_CLIP_BINNED, "bad mode");
    if (mode == CAMA_CLIP_BINNED && !binned_ok)
        return fail(CAMA_E_UNSUPPORTED, "BINNED mode needs width %% 16 == 0, width <= 2048, <= 65534 instances");
    p = ClipPlan();
    p.mode = mode;
    size_t off = 0;
    auto take = [&](size_t bytes) { size_t o = off; off = align_up(off + bytes, 256); return o; };
    p.off_stats = take(sizeof(ClipStatsDev));                             // first, in every layout: cama_clip_stats_read finds block 0 at the workspace's start
    p.off_w2c64 = take(sizeof(double) * 12 * (size_t)std::max(d->n_frames, 1));
    p.off_lut = take(sizeof(unsigned) * ((size_t)d->n_instances + 1));
    p.geo_units = ((d->n_vertices + kGeoThreads - 1) / kGeoThreads) * ((d->n_frames + kGeoFrames - 1) / kGeoFrames);
    p.off_worklist = take(sizeof(unsigned long long) * (size_t)std::max<long long>(p.geo_units, 1));
    p.off_zero = off;                                                      // cleared by prep every call: counters | list cursors (BINNED)
    p.off_counter = take(256);
    p.zero_bytes = off - p.off_zero;
    if (mode == CAMA_CLIP_PLANE) {
        p.off_plane = take(sizeof(unsigned) * (size_t)d->n_frames * d->n_cams * H * W);
    } else {
        p.band_rows = band_rows;
        p.x_bits = x_bits;
        p.n_bands = (H + band_rows - 1) / band_rows;
        p.n_strips = (W + kSegPx - 1) / kSegPx;

        const long long nb = (long long)d->n_frames * d->n_cams * p.n_bands;
        CAMA_REQUIRE(nb < INT_MAX, "too many (frame, camera, band) work items");
        p.n_buckets = (int)nb;
        // band groups: a record stores its row inside the group (+ 2 halo rows either side) in 16 - x_bits bits
        static const int max_group_bands = getenv("CAMA_GROUP_BANDS") ? std::max(1, atoi(getenv("CAMA_GROUP_BANDS"))) : kDefaultGroupBands;   // tuning knob
        p.group_bands = (int)std::max<long long>(1, std::min<long long>(max_group_bands, ((1ll << (16 - x_bits)) - 4) / band_rows));
        p.n_groups = (p.n_bands + p.group_bands - 1) / p.group_bands;
        p.n_lists = (long long)d->n_frames * d->n_cams * p.n_groups;
        long long cap = d->record_capacity;
        if (cap <= 0) cap = std::min<long long>(std::max<long long>(d->n_vertices / 4, 2048), 8192 * (long long)p.group_bands);   // per list; an overflow is reported and the caller reruns
        CAMA_REQUIRE(cap < (1ll << 31), "record_capacity too large");
        p.cap = cap;
        p.raster_smem = (size_t)(band_rows + 5) * W * 2 + kRasterStageSmem + (size_t)kZeroRows * W * 3;
        if (d->overlay_records) p.raster_smem += sizeof(OvStage) * kRasterWarps;
        if (!d->list_records) {                                              // (external lists: frame-sharded clips)
            p.off_cursor = take(sizeof(unsigned) * ((size_t)p.n_lists + 1));    // (directly after the counter block: cleared with it)
            p.zero_bytes = off - p.off_zero;
        }
        p.off_lists = take(sizeof(unsigned) * 4 * (size_t)nb);             // work lists of the raster, one per weight class
        if (!d->list_records) p.off_records = take(sizeof(unsigned) * (size_t)p.n_lists * (size_t)cap);
    }
    p.total = std::max<size_t>(off, 256);
    p.groups = 1;
    p.group_frames = d->n_frames;
    p.group_stride = 0;
    // Frame groups: the geometry of group g+1 runs while group g is sorted and rastered (three stream lanes, see
    // cama_clip_render).  Not with the per-instance debug outputs (their consumers want one pass), not for small clips.
    if (allow_groups && mode == CAMA_CLIP_BINNED && !d->crop_counts && !d->visible_counts && !d->vu_dense && d->phases == CAMA_PHASE_ALL && !d->list_records) {
        int want = d->pipeline_frames > 0 ? (d->pipeline_frames + 7) / 8 * 8 : d->pipeline_frames < 0 ? 0 : pipe_group_frames();
        if (want > 0 && d->n_frames > want) {
            cama_clip_desc sub = *d;
            sub.n_frames = want;
            ClipPlan q;
            const int rc = make_plan(&sub, q, false);
            if (rc != CAMA_OK) return rc;
            p.group_frames = want;
            p.groups = (d->n_frames + want - 1) / want;
            p.group_stride = align_up(q.total, 256);
            p.total = std::max(p.total, p.group_stride * (size_t)p.groups);     // (the un-grouped layout is used while phase profiling is on)
        }
    }
    return CAMA_OK;
}

// Streams one pass (a whole clip, or one frame group) is issued on, and the events that order them
// (nullptr: the three lanes are one stream).
struct Lanes {
    cudaStream_t geo, sort, raster;
    cudaEvent_t geo_done, sort_done;
};

// Enqueues one pass over the frames of `d` with the workspace slice `ws` laid out by `p`.  image_base: index of the
// pass's first (frame, camera) image inside the clip (sparse output); first: this pass clears the clip-wide counters.
int render_pass(cama_ctx *ctx, const cama_clip_desc *d, const ClipPlan &p, unsigned char *ws, const CamBlock &cams, const Lanes &lanes,
                cudaEvent_t *prof, int image_base, bool first, int groups_used) {
    cudaStream_t s = lanes.geo;
    auto mark = [&](int i) { return prof ? cudaEventRecord(prof[i], s) : cudaSuccess; };
    CAMA_CUDA_TRY(mark(0));
    ClipArgs a{};
    a.n_frames = d->n_frames; a.n_cams = d->n_cams; a.n_instances = d->n_instances;
    a.height = d->height; a.width = d->width; a.n_vertices = d->n_vertices;
    a.vertices = d->vertices; a.vertex_instance = d->vertex_instance;
    a.w2c64 = reinterpret_cast<const double *>(ws + p.off_w2c64);
    a.crop_counts = d->crop_counts; a.visible_counts = d->visible_counts; a.vu_dense = d->vu_dense;
    a.tile_bounds = d->tile_bounds;
    unsigned *lut = reinterpret_cast<unsigned *>(ws + p.off_lut);
    ClipStatsDev *stats = reinterpret_cast<ClipStatsDev *>(ws + p.off_stats);
    const bool binned = p.mode == CAMA_CLIP_BINNED;
    const bool pdl = binned && pdl_enabled() && !prof;

    // camera table (cama_camera_table_build, once per rig): which cameras can see a cell of the crop box at all
    {
        static const bool no_table = getenv("CAMA_GEO_NO_CAMTABLE") != nullptr;          // experiment knob
        CamTable tab{};
        if (!no_table && d->camera_table && table_geometry(d->crop_box, tab)) {
            a.cam_table = static_cast<const unsigned char *>(d->camera_table);
            a.tab_signature = table_signature(cams, d->n_cams, d->width, d->height);
            a.tab_x0 = tab.x0; a.tab_y0 = tab.y0;
            a.tab_inv_sx = (float)(1.0 / tab.sx); a.tab_inv_sy = (float)(1.0 / tab.sy);
            a.tab_nx = tab.nx; a.tab_ny = tab.ny;
        }
    }
    static const bool no_warp_bounds = getenv("CAMA_GEO_NO_WARPBOUNDS") != nullptr;     // experiment knob
    a.warp_bounds = no_warp_bounds ? nullptr : d->warp_bounds;
    a.geo_counter = reinterpret_cast<unsigned *>(ws + p.off_counter) + 8;
    if (d->vu_dense)
        CAMA_CUDA_TRY(cudaMemsetAsync(d->vu_dense, 0xff, sizeof(double) * 2 * (size_t)d->n_frames * d->n_cams * d->n_vertices, s));
    {
        NvtxRange nvtx_phase("prep");
        const long long zero_words = (long long)(p.zero_bytes / 4);
        const long long n = std::max<long long>(std::max(d->n_frames * 12, d->n_instances + 1), std::min<long long>(zero_words, 1 << 20));
        // (PDL only in BINNED mode: PLANE mode has memsets between its kernels)
        // external record lists (frame-sharded clips): the cursors of THIS call's frames are cleared, unless the call only rasters
        unsigned *zero2 = nullptr;
        long long zero2_words = 0;
        if (binned && d->list_records && d->phases != CAMA_PHASE_RASTER) {
            zero2 = d->list_cursor + (size_t)d->list_frame_base * d->n_cams * p.n_groups;
            zero2_words = (long long)d->n_frames * d->n_cams * p.n_groups;
        }
        CAMA_CUDA_TRY(launch_k(pdl, prep_kernel, (unsigned)((std::max(n, std::min<long long>(zero2_words, 1 << 20)) + 255) / 256), 256, 0, s,
                               d->world2chassis, d->n_frames, reinterpret_cast<double *>(ws + p.off_w2c64),
                               d->instance_bgr, d->n_instances, lut,
                               reinterpret_cast<unsigned *>(ws + p.off_zero), zero_words,
                               reinterpret_cast<unsigned *>(stats), (int)(sizeof(ClipStatsDev) / 4),
                               d->overlay_records && first ? d->overlay_count : (unsigned *)nullptr, d->instance_palette, zero2, zero2_words,
                               (unsigned)groups_used));
        CAMA_LAUNCHED(ctx);
    }
    const long long n_tiles = (d->n_vertices + kGeoThreads - 1) / kGeoThreads;
    const long long units = n_tiles * ((d->n_frames + kGeoFrames - 1) / kGeoFrames);          // (256-vertex tile, 8 frames): what the cull kernel works on
    const bool site = BinnedSite(d, ctx, units);                     // big clips: the cull kernel makes a work list first
    static const int uf_clip = getenv("CAMA_GEO_UNIT_FRAMES") ? atoi(getenv("CAMA_GEO_UNIT_FRAMES")) : CAMA_GEO_UNIT_FRAMES;
    static const int uf_site = getenv("CAMA_GEO_UNIT_FRAMES_SITE") ? atoi(getenv("CAMA_GEO_UNIT_FRAMES_SITE")) : CAMA_GEO_UNIT_FRAMES_SITE;
    static const int static_pct = getenv("CAMA_GEO_STATIC_PCT") ? atoi(getenv("CAMA_GEO_STATIC_PCT")) : CAMA_GEO_STATIC_PCT;
    int unit_frames = site && p.mode == CAMA_CLIP_BINNED ? uf_site : uf_clip;
    if (unit_frames != 1 && unit_frames != 2 && unit_frames != 4 && unit_frames != 8) unit_frames = 4;
    a.unit_frames = unit_frames;
    const long long warp_units = ((d->n_vertices + 31) / 32) * ((d->n_frames + unit_frames - 1) / unit_frames);
    // (with a work list the number of live units is only known on the device: everything is claimed dynamically there)
    a.static_units = site && p.mode == CAMA_CLIP_BINNED ? 0u : (unsigned)(warp_units * std::min(100, std::max(0, static_pct)) / 100);
    // persistent grid: every resident warp claims (32-vertex, 8-frame) units until none are left
    static const int geo_ctas_env = getenv("CAMA_GEO_CTAS") ? std::max(1, atoi(getenv("CAMA_GEO_CTAS"))) : 0;      // experiment knob
    const int geo_ctas = geo_ctas_env ? geo_ctas_env : d->geometry_ctas_per_sm > 0 ? std::min(d->geometry_ctas_per_sm, CAMA_GEO_MINB) : CAMA_GEO_MINB;
    const unsigned geo_grid = (unsigned)std::max<long long>(1, std::min<long long>((warp_units + 7) / 8, (long long)ctx->sm_count * geo_ctas));
    const bool f32 = d->vertex_layout == CAMA_VERTEX_F32X4;
    const bool debug = d->crop_counts || d->visible_counts || d->vu_dense;

    if (p.mode == CAMA_CLIP_PLANE) {
        a.plane = reinterpret_cast<unsigned *>(ws + p.off_plane);
        const size_t px = (size_t)d->n_frames * d->n_cams * d->height * d->width;
        CAMA_CUDA_TRY(cudaMemsetAsync(a.plane, 0, sizeof(unsigned) * px, s));
        CAMA_CUDA_TRY(mark(1));
        if (units > 0) {
            CAMA_CUDA_TRY(launch_geometry<false>(false, f32, debug, geo_grid, s, a, cams));
            CAMA_LAUNCHED(ctx);
        }
        CAMA_CUDA_TRY(mark(2));
        CAMA_CUDA_TRY(mark(3));
        plane_raster_kernel<<<(unsigned)((px + 255) / 256), 256, 0, s>>>(a.plane, lut, d->background, d->frames, d->height, d->width,
                                                                        (long long)d->n_frames * d->n_cams);
        CAMA_LAUNCHED(ctx);
        CAMA_CUDA_TRY(mark(4));
        return CAMA_OK;
    }

    // BINNED (the workspace slice may be laid out for more frames than this pass has: the last frame group)
    const int n_buckets = d->n_frames * d->n_cams * p.n_bands;
    const int n_lists = d->n_frames * d->n_cams * p.n_groups;
    a.band_rows = p.band_rows; a.n_bands = p.n_bands; a.x_bits = p.x_bits;
    a.group_rows = p.band_rows * p.group_bands; a.n_groups = p.n_groups;
    a.group_magic = (unsigned)(((1ull << 32) + a.group_rows - 1) / a.group_rows);
    a.list_cap = (unsigned)p.cap;
    if (d->list_records) {                       // lists outside the workspace, shared with the peers; frame f of this call = list frame base + f
        a.records = static_cast<unsigned *>(d->list_records);
        a.cursor = d->list_cursor;
        a.list_frame_base = d->list_frame_base;
    } else {
        a.cursor = reinterpret_cast<unsigned *>(ws + p.off_cursor);
        a.records = reinterpret_cast<unsigned *>(ws + p.off_records);
    }
    CAMA_CUDA_TRY(mark(1));
    if (units > 0 && d->phases != CAMA_PHASE_RASTER) {
        NvtxRange nvtx_phase("geometry");
        // big clips: cull (tile, frame chunk) units first and run the geometry over the live ones only
        if (site) {
            unsigned *n_live = reinterpret_cast<unsigned *>(ws + p.off_counter) + 2;
            unsigned long long *worklist = reinterpret_cast<unsigned long long *>(ws + p.off_worklist);
            CAMA_CUDA_TRY(launch_k(pdl, geometry_cull_kernel, (unsigned)((units * kGeoFrames + 255) / 256), 256, 0, s, d->tile_bounds, a.w2c64, units,
                                   (d->n_frames + kGeoFrames - 1) / kGeoFrames, d->n_frames, cams, worklist, n_live));
            CAMA_LAUNCHED(ctx);
            a.worklist = worklist;
            a.n_live = n_live;
        }
        CAMA_CUDA_TRY(launch_geometry<true>(pdl, f32, debug, geo_grid, s, a, cams));
        CAMA_LAUNCHED(ctx);
    }
    CAMA_CUDA_TRY(mark(2));
    if (d->phases == CAMA_PHASE_GEOMETRY) {      // the lists are handed to the peers; classify + raster run in a later call
        CAMA_CUDA_TRY(mark(3));
        CAMA_CUDA_TRY(mark(4));
        return CAMA_OK;
    }
    // (a raster-only call reads lists whose frame 0 is this call's frame 0 + list_frame_base)
    const unsigned *cursor = a.cursor + (size_t)a.list_frame_base * d->n_cams * p.n_groups;
    const unsigned *records = a.records + (size_t)a.list_frame_base * d->n_cams * p.n_groups * a.list_cap;
    const bool lanes_split = lanes.geo_done != nullptr;
    if (lanes_split) {
        CAMA_CUDA_TRY(cudaEventRecord(lanes.geo_done, s));
        CAMA_CUDA_TRY(cudaStreamWaitEvent(lanes.sort, lanes.geo_done, 0));
    }
    s = lanes.sort;
    unsigned *lists = reinterpret_cast<unsigned *>(ws + p.off_lists);
    unsigned *list_counts = reinterpret_cast<unsigned *>(ws + p.off_counter) + 4;
    {
        NvtxRange nvtx_phase("classify");
        CAMA_CUDA_TRY(launch_k(pdl && !lanes_split, band_classify_kernel, (unsigned)((std::max(n_buckets, n_lists) + 255) / 256), 256, 0, s, cursor, n_lists, n_buckets,
                               p.n_bands, p.group_bands, p.n_groups, a.list_cap, stats, lists, list_counts));
        CAMA_LAUNCHED(ctx);
    }
    if (lanes_split) {
        CAMA_CUDA_TRY(cudaEventRecord(lanes.sort_done, s));
        CAMA_CUDA_TRY(cudaStreamWaitEvent(lanes.raster, lanes.sort_done, 0));
    }
    s = lanes.raster;
    CAMA_CUDA_TRY(mark(3));
    NvtxRange nvtx_raster("raster");
    RasterArgs r{};
    r.n_items = n_buckets; r.n_bands = p.n_bands; r.band_rows = p.band_rows; r.height = d->height; r.width = d->width;
    r.n_instances = d->n_instances; r.group_bands = p.group_bands; r.n_groups = p.n_groups; r.list_cap = (unsigned)p.cap;
    r.x_bits = p.x_bits; r.n_strips = p.n_strips; r.image_base = image_base;
    if (const char *env = getenv("CAMA_RASTER_DEBUG")) r.debug = atoi(env);
    r.cursor = cursor; r.records = records; r.lut = lut; r.bg = d->background; r.frames = d->frames;
    r.n_cams = d->n_cams;
    r.image_bytes = (unsigned long long)d->height * d->width * 3;
    r.frame_stride = frame_stride_bytes(d);
    {
        const size_t image = (size_t)d->height * d->width * 3, row = (size_t)d->width * 3;
        r.pitch = (unsigned)(d->mosaic_cols > 0 ? row * d->mosaic_cols : row);
        for (int c = 0; c < d->n_cams; ++c) {
            if (d->mosaic_cols > 0) {
                const int tile = d->mosaic_tile_of_cam[c];
                r.cam_offset[c] = (unsigned long long)(tile / d->mosaic_cols) * d->height * r.pitch + (unsigned long long)(tile % d->mosaic_cols) * row;
            } else {
                r.cam_offset[c] = (unsigned long long)c * image;
            }
        }
    }
    r.lists = lists; r.list_counts = list_counts;
    r.work_counter = reinterpret_cast<unsigned *>(ws + p.off_counter);
    r.empty_counter = reinterpret_cast<unsigned *>(ws + p.off_counter) + 3;
    static const int dyn_empty_pct = getenv("CAMA_RASTER_DYN_EMPTY") ? std::min(100, std::max(0, atoi(getenv("CAMA_RASTER_DYN_EMPTY")))) : kDynEmptyPct;
    static const int dyn_empty_per_cta = getenv("CAMA_RASTER_DYN_PER_CTA") ? std::max(0, atoi(getenv("CAMA_RASTER_DYN_PER_CTA"))) : kDynEmptyPerCta;
    r.dyn_empty_pct = dyn_empty_pct;
    r.dyn_empty_per_cta = dyn_empty_per_cta;
    static const int raster_ctas_env = getenv("CAMA_RASTER_CTAS") ? std::max(1, atoi(getenv("CAMA_RASTER_CTAS"))) : 0;      // experiment knob
    const int raster_ctas = raster_ctas_env ? raster_ctas_env : d->raster_ctas_per_sm > 0 ? std::min(d->raster_ctas_per_sm, kRasterCtasPerSm) : kRasterCtasPerSm;
    const unsigned raster_grid = (unsigned)std::min<long long>(n_buckets, (long long)ctx->sm_count * raster_ctas);
    const bool rpdl = pdl && !lanes_split;
    if (d->overlay_records) {
        r.ov_records = d->overlay_records; r.ov_count = d->overlay_count; r.ov_cap = d->overlay_capacity;
        r.ov_n_mirrors = d->overlay_n_mirrors;
        for (int m = 0; m < d->overlay_n_mirrors; ++m) r.ov_mirrors[m] = d->overlay_mirrors[m];
        r.image_base = image_base + (int)d->overlay_image_base;
        if (d->overlay_format == CAMA_OVERLAY_PALETTE) {
            CAMA_CUDA_TRY(cudaFuncSetAttribute(binned_raster_kernel<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)p.raster_smem));
            CAMA_CUDA_TRY(launch_k(rpdl, binned_raster_kernel<3>, raster_grid, kRasterBlock, p.raster_smem, s, r));
        } else {
            CAMA_CUDA_TRY(cudaFuncSetAttribute(binned_raster_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)p.raster_smem));
            CAMA_CUDA_TRY(launch_k(rpdl, binned_raster_kernel<2>, raster_grid, kRasterBlock, p.raster_smem, s, r));
        }
    } else if (d->background) {
        CAMA_CUDA_TRY(cudaFuncSetAttribute(binned_raster_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)p.raster_smem));
        CAMA_CUDA_TRY(launch_k(rpdl, binned_raster_kernel<1>, raster_grid, kRasterBlock, p.raster_smem, s, r));
    } else {
        CAMA_CUDA_TRY(cudaFuncSetAttribute(binned_raster_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)p.raster_smem));
        CAMA_CUDA_TRY(launch_k(rpdl, binned_raster_kernel<0>, raster_grid, kRasterBlock, p.raster_smem, s, r));
    }
    CAMA_LAUNCHED(ctx);
    CAMA_CUDA_TRY(mark(4));
    return CAMA_OK;
}

}  // namespace

extern "C" {

int cama_camera_table_build(cama_ctx *ctx, const double *chassis2cam, const double *intrinsics, int n_cams, const double *crop_box, int height,
                            int width, void *table, void *stream) {
    CAMA_REQUIRE(ctx && chassis2cam && intrinsics && crop_box && table, "NULL argument");
    CAMA_REQUIRE(n_cams > 0 && n_cams <= CAMA_MAX_CAMERAS && height > 0 && width > 0, "bad shape");
    CAMA_REQUIRE(((uintptr_t)table & 15) == 0, "table must be 16-byte aligned");
    DeviceGuard guard(ctx->device);
    cudaStream_t s = (cudaStream_t)stream;
    CamBlock cams;
    fill_cam_block(cams, n_cams, chassis2cam, intrinsics, crop_box, width, height);
    CamTable tab{};
    if (!table_geometry(crop_box, tab)) {                      // no grid for this box: an all-zero header never matches a signature
        CAMA_CUDA_TRY(cudaMemsetAsync(table, 0, CAMA_CAMERA_TABLE_HEADER, s));
        return CAMA_OK;
    }
    const int n = tab.nx * tab.ny * 8;
    camera_table_kernel<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(cams, n_cams, width, height, tab, table_signature(cams, n_cams, width, height),
                                                                static_cast<unsigned char *>(table));
    CAMA_LAUNCHED(ctx);
    return CAMA_OK;
}

int cama_clip_workspace_bytes(const cama_clip_desc *desc, size_t *bytes) {
    CAMA_REQUIRE(bytes, "bytes is NULL");
    ClipPlan p;
    const int rc = make_plan(desc, p);
    if (rc != CAMA_OK) return rc;
    *bytes = p.total;
    return CAMA_OK;
}

int cama_clip_render(cama_ctx *ctx, const cama_clip_desc *d, void *workspace, size_t workspace_bytes, void *stream) {
    CAMA_REQUIRE(ctx, "ctx is NULL");
    ClipPlan p;
    int rc = make_plan(d, p);
    if (rc != CAMA_OK) return rc;
    if (!workspace || workspace_bytes < p.total) return fail(CAMA_E_WORKSPACE, "clip workspace: need %zu bytes, got %zu", p.total, workspace_bytes);
    CAMA_REQUIRE(((uintptr_t)workspace & 255) == 0, "workspace must be 256-byte aligned");
    if (d->n_frames == 0) return CAMA_OK;
    CAMA_REQUIRE(d->phases == CAMA_PHASE_ALL || d->phases == CAMA_PHASE_GEOMETRY || d->phases == CAMA_PHASE_RASTER, "bad phases");
    if (d->phases != CAMA_PHASE_ALL || d->list_records) {
        if (p.mode != CAMA_CLIP_BINNED) return fail(CAMA_E_UNSUPPORTED, "split phases / external record lists need BINNED mode");
        CAMA_REQUIRE(d->list_records && d->list_cursor, "split phases need external record lists (list_records, list_cursor)");
        CAMA_REQUIRE(d->list_frame_base >= 0 && (d->list_frames == 0 ? d->list_frame_base == 0 : d->list_frame_base + d->n_frames <= d->list_frames),
                     "this call's frames do not fit the external lists");
        CAMA_REQUIRE(!d->crop_counts && !d->visible_counts && !d->vu_dense, "the per-instance debug outputs need the whole pipeline in one call");
    }
    CAMA_REQUIRE(d->chassis2cam && d->intrinsics, "NULL buffer in desc");
    CAMA_REQUIRE(d->world2chassis || d->phases == CAMA_PHASE_RASTER, "world2chassis is NULL");
    CAMA_REQUIRE(d->frames || d->overlay_records || d->phases == CAMA_PHASE_GEOMETRY, "neither frames nor overlay_records given");
    if (d->overlay_records) {
        if (p.mode != CAMA_CLIP_BINNED) return fail(CAMA_E_UNSUPPORTED, "the sparse overlay output needs BINNED mode");
        CAMA_REQUIRE(!d->background, "the sparse overlay output takes no background (the host composites)");
        CAMA_REQUIRE(d->overlay_count && d->overlay_capacity > 0, "overlay_count / overlay_capacity missing");
        CAMA_REQUIRE(((uintptr_t)d->overlay_records & 15) == 0, "overlay_records must be 16-byte aligned");
        CAMA_REQUIRE(d->overlay_format == CAMA_OVERLAY_BGR || d->overlay_format == CAMA_OVERLAY_PALETTE, "bad overlay_format");
        CAMA_REQUIRE(d->overlay_format != CAMA_OVERLAY_PALETTE || d->n_instances == 0 || d->instance_palette, "the palette overlay format needs instance_palette");
        CAMA_REQUIRE(d->overlay_image_base >= 0 && ((long long)d->n_frames * d->n_cams + d->overlay_image_base) * d->height * d->width / 8 < (1ll << 32),
                     "clip too large for 32-bit chunk indices");
        CAMA_REQUIRE(d->overlay_n_mirrors >= 0 && d->overlay_n_mirrors <= CAMA_MAX_PEERS, "overlay_n_mirrors out of range");
        for (int m = 0; m < d->overlay_n_mirrors; ++m) CAMA_REQUIRE(d->overlay_mirrors[m] && ((uintptr_t)d->overlay_mirrors[m] & 15) == 0, "overlay_mirrors[%d] is NULL or misaligned", m);
    }
    if (d->mosaic_cols != 0) {
        CAMA_REQUIRE(d->mosaic_cols > 0, "mosaic_cols is negative");
        if (p.mode != CAMA_CLIP_BINNED) return fail(CAMA_E_UNSUPPORTED, "the mosaic layout needs BINNED mode");
        const int tiles = (d->n_cams + d->mosaic_cols - 1) / d->mosaic_cols * d->mosaic_cols;
        for (int c = 0; c < d->n_cams; ++c) {
            CAMA_REQUIRE(d->mosaic_tile_of_cam[c] >= 0 && d->mosaic_tile_of_cam[c] < tiles, "mosaic_tile_of_cam[%d] out of range", c);
            for (int e = 0; e < c; ++e) CAMA_REQUIRE(d->mosaic_tile_of_cam[e] != d->mosaic_tile_of_cam[c], "cameras %d and %d share a mosaic tile", e, c);
        }
    }
    CAMA_REQUIRE(d->n_vertices == 0 || d->vertices, "vertices is NULL");
    CAMA_REQUIRE(d->n_instances == 0 || d->instance_bgr, "instance_bgr is NULL");
    CAMA_REQUIRE(d->vertex_layout != CAMA_VERTEX_F64X3 || d->n_vertices == 0 || d->vertex_instance, "vertex_instance is NULL");
    CAMA_REQUIRE(((uintptr_t)d->frames & 15) == 0 && ((uintptr_t)d->background & 15) == 0 && ((uintptr_t)d->vertices & 15) == 0,
                 "frames/background/vertices must be 16-byte aligned");
    DeviceGuard guard(ctx->device);
    NvtxRange nvtx_call("cama_clip_render");
    cudaStream_t s = (cudaStream_t)stream;
    unsigned char *ws = static_cast<unsigned char *>(workspace);
    // phase events of this call, when profiling is on (cama_ctx_profile_enable)
    cudaEvent_t *prof = nullptr;
    if (ctx->prof_calls < ctx->prof_capacity) prof = ctx->prof_events.data() + (size_t)(ctx->prof_calls++) * (CAMA_CLIP_PHASES + 1);

    CamBlock cams;
    fill_cam_block(cams, d->n_cams, d->chassis2cam, d->intrinsics, d->crop_box, d->width, d->height);

    if (p.groups <= 1 || prof) {                       // one pass, one stream (phase events would serialise the lanes anyway)
        ClipPlan whole = p;
        if (p.groups > 1) {
            rc = make_plan(d, whole, false);
            if (rc != CAMA_OK) return rc;
        }
        Lanes lanes{s, s, s, nullptr, nullptr};
        return render_pass(ctx, d, whole, ws, cams, lanes, prof, 0, true, 1);
    }

    // Frame-group pipeline.  Lane 0 = the caller's stream: prep + geometry of every group, back to back; lane 1:
    // scan + scatter of a group once its geometry is done; lane 2: its raster once it is sorted.  The geometry
    // (FP64 issue-bound, no memory traffic) of group g+1 thus runs under the raster (store-bound) of group g.
    if (!ctx->pipe_streams[0]) {
        int lo = 0, hi = 0;
        CAMA_CUDA_TRY(cudaDeviceGetStreamPriorityRange(&lo, &hi));
        static const int raster_prio = getenv("CAMA_PIPE_RASTER_PRIO") ? atoi(getenv("CAMA_PIPE_RASTER_PRIO")) : 0;
        CAMA_CUDA_TRY(cudaStreamCreateWithPriority(&ctx->pipe_streams[0], cudaStreamNonBlocking, raster_prio ? hi : lo));
        CAMA_CUDA_TRY(cudaStreamCreateWithPriority(&ctx->pipe_streams[1], cudaStreamNonBlocking, raster_prio ? hi : lo));
    }
    while (ctx->pipe_events.size() < (size_t)(2 * p.groups + 1)) {
        cudaEvent_t e;
        CAMA_CUDA_TRY(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
        ctx->pipe_events.push_back(e);
    }
    cudaStream_t s_sort = ctx->pipe_streams[0], s_raster = ctx->pipe_streams[1];
    // (the side lanes only ever start after an event recorded on the caller's stream, i.e. after everything enqueued there before)
    cudaEvent_t e_join = ctx->pipe_events[2 * p.groups];
    ClipPlan q;                                                // every slice has the layout of a full group
    {
        cama_clip_desc full = *d;
        full.n_frames = p.group_frames;
        rc = make_plan(&full, q, false);
        if (rc != CAMA_OK) return rc;
        if (q.total > p.group_stride) return fail(CAMA_E_WORKSPACE, "internal: group plan larger than its slice");
    }
    for (int g = 0; g < p.groups; ++g) {
        const int f0 = g * p.group_frames;
        cama_clip_desc sub = *d;
        sub.n_frames = std::min(p.group_frames, d->n_frames - f0);
        sub.world2chassis = d->world2chassis + (size_t)f0 * 16;
        if (d->frames) sub.frames = d->frames + (size_t)f0 * frame_stride_bytes(d);
        if (d->background) sub.background = d->background + (size_t)f0 * frame_stride_bytes(d);
        Lanes lanes{s, s_sort, s_raster, ctx->pipe_events[2 * g], ctx->pipe_events[2 * g + 1]};
        rc = render_pass(ctx, &sub, q, ws + (size_t)g * p.group_stride, cams, lanes, nullptr, f0 * d->n_cams, g == 0, p.groups);
        if (rc != CAMA_OK) return rc;
    }
    CAMA_CUDA_TRY(cudaEventRecord(e_join, s_raster));          // (the raster lane's last kernel waited for everything on the sort lane)
    CAMA_CUDA_TRY(cudaStreamWaitEvent(s, e_join, 0));
    return CAMA_OK;
}

int cama_clip_stats_read(cama_ctx *ctx, const cama_clip_desc *d, const void *workspace, void *stream, cama_clip_stats *out) {
    CAMA_REQUIRE(ctx && out && workspace, "NULL argument");
    ClipPlan p;
    int rc = make_plan(d, p);
    if (rc != CAMA_OK) return rc;
    if (d->n_frames == 0) {                        // cama_clip_render launched nothing: the workspace holds no statistics of this call
        *out = cama_clip_stats{};
        out->record_capacity = p.cap;
        out->mode = p.mode;
        out->band_rows = p.band_rows;
        out->n_bands = p.n_bands;
        out->lists_per_image = p.n_groups;
        return CAMA_OK;
    }
    DeviceGuard guard(ctx->device);
    // The statistics block of a pass sits at the start of its workspace slice and names the number of slices (frame
    // groups) the render used: 1 for a one-pass render (also a phase-profiled one), p.groups for the pipeline.  So the
    // layout of the render that last used this workspace is read from the workspace itself, not remembered in the context.
    const int n_blocks = std::max(p.groups, 1);
    std::vector<ClipStatsDev> h((size_t)n_blocks);
    unsigned n_overlay = 0;
    const bool sparse = d->overlay_records && d->overlay_count;
    for (int g = 0; g < n_blocks; ++g)
        CAMA_CUDA_TRY(cudaMemcpyAsync(&h[g], static_cast<const unsigned char *>(workspace) + (size_t)g * p.group_stride + p.off_stats, sizeof(ClipStatsDev),
                                      cudaMemcpyDeviceToHost, (cudaStream_t)stream));
    if (sparse) CAMA_CUDA_TRY(cudaMemcpyAsync(&n_overlay, d->overlay_count, sizeof(n_overlay), cudaMemcpyDeviceToHost, (cudaStream_t)stream));
    CAMA_CUDA_TRY(cudaStreamSynchronize((cudaStream_t)stream));      // one round trip for all counters
    h.resize((size_t)std::min<unsigned>(std::max(h[0].groups_used, 1u), (unsigned)n_blocks));
    unsigned long long total = 0, max_per_frame = 0;
    unsigned overflow = 0;
    for (const ClipStatsDev &b : h) {
        total += b.records_total;
        max_per_frame = std::max(max_per_frame, b.list_max);
        overflow |= b.overflow;
    }
    out->records_total = (int64_t)total;
    out->record_capacity_needed = (int64_t)max_per_frame;
    out->record_capacity = p.cap;
    out->overflow = (int32_t)overflow;
    out->mode = p.mode;
    out->band_rows = p.band_rows;
    out->n_bands = p.n_bands;
    out->lists_per_image = p.n_groups;
    out->overlay_records = sparse ? n_overlay : 0;
    if (overflow)
        return fail(CAMA_E_CAPACITY, "record list overflow: the fullest list needs %llu records, capacity %lld", max_per_frame, p.cap);
    return CAMA_OK;
}

// tuning aid (CAMA_RASTER_DEBUG & 32): copies the per-CTA {start ns, end ns, active bands} log of the last raster launch
int cama_debug_raster_timeline(unsigned long long *out, int n_ctas) {
    CAMA_REQUIRE(out && n_ctas > 0 && n_ctas <= 4096, "bad argument");
    CAMA_CUDA_TRY(cudaDeviceSynchronize());
    CAMA_CUDA_TRY(cudaMemcpyFromSymbol(out, g_raster_timeline, sizeof(unsigned long long) * 3 * (size_t)n_ctas));
    return CAMA_OK;
}

}  // extern "C"
