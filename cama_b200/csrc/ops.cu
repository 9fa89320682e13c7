// libcama_b200: context management and the per-call operators, one per reference method
// (transform / crop / project / render of /root/reference/cama/reproject.py).  These serve the
// call-compatible Python managers; the throughput path is clip.cu.
#include <climits>

#include "common.cuh"
#include "geom.cuh"

namespace cama {
thread_local char g_last_error[512] = "";

constexpr int kBlock = 256;
constexpr unsigned kFullMask = 0xffffffffu;

// ------------------------------------------------------------------------------------------------
// One point through [optional 4x4] -> {crop box | pinhole projection}.
struct PointOp {
    const void *pts;
    long long n;
    int is_f32, has_T, project, width, height;
    double T[12];
    double box[6];
    double K[9];
};

__device__ __forceinline__ bool eval_point(const PointOp &op, long long i, double &o0, double &o1, double &o2) {
    double x, y, z;
    if (op.is_f32) {
        const float *p = static_cast<const float *>(op.pts) + 3 * i;
        x = (double)p[0]; y = (double)p[1]; z = (double)p[2];
    } else {
        const double *p = static_cast<const double *>(op.pts) + 3 * i;
        x = p[0]; y = p[1]; z = p[2];
    }
    if (op.has_T) {
        const double tx = affine_row(op.T, x, y, z);
        const double ty = affine_row(op.T + 4, x, y, z);
        const double tz = affine_row(op.T + 8, x, y, z);
        x = tx; y = ty; z = tz;
    }
    if (!op.project) {
        o0 = x; o1 = y; o2 = z;
        return in_box(op.box, x, y, z);
    }
    o2 = 0.0;
    return project_point(op.K, x, y, z, op.width, op.height, o0, o1);
}

// R9 alone: no compaction.
__global__ void __launch_bounds__(kBlock) transform_kernel(const __grid_constant__ PointOp op, double *__restrict__ out) {
    const long long i = (long long)blockIdx.x * kBlock + threadIdx.x;
    if (i >= op.n) return;
    double a, b, c;
    eval_point(op, i, a, b, c);
    out[3 * i] = a; out[3 * i + 1] = b; out[3 * i + 2] = c;
}

// pass 1 of the order-preserving compaction: survivors per block
__global__ void __launch_bounds__(kBlock) compact_count_kernel(const __grid_constant__ PointOp op, long long *__restrict__ block_counts) {
    const long long i = (long long)blockIdx.x * kBlock + threadIdx.x;
    double a, b, c;
    const bool keep = (i < op.n) && eval_point(op, i, a, b, c);
    const int cnt = __syncthreads_count(keep);
    if (threadIdx.x == 0) block_counts[blockIdx.x] = cnt;
}

// pass 2: exclusive scan of the per-block counts, in place; total goes to counts[nb]
__global__ void __launch_bounds__(1024) scan_i64_kernel(long long *__restrict__ counts, long long nb) {
    __shared__ long long warp_sum[32];
    __shared__ long long carry;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid == 0) carry = 0;
    __syncthreads();
    for (long long base = 0; base < nb; base += 1024) {
        const long long idx = base + tid;
        const long long v = idx < nb ? counts[idx] : 0;
        long long inc = v;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const long long t = __shfl_up_sync(kFullMask, inc, d);
            if (lane >= d) inc += t;
        }
        if (lane == 31) warp_sum[warp] = inc;
        __syncthreads();
        if (warp == 0) {
            long long w = warp_sum[lane];
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                const long long t = __shfl_up_sync(kFullMask, w, d);
                if (lane >= d) w += t;
            }
            warp_sum[lane] = w;      // inclusive over warps
        }
        __syncthreads();
        const long long before = carry + (warp > 0 ? warp_sum[warp - 1] : 0) + inc - v;
        if (idx < nb) counts[idx] = before;
        __syncthreads();
        if (tid == 1023) carry = before + v;
        __syncthreads();
    }
    if (tid == 0) counts[nb] = carry;
}

// pass 3: recompute, place survivors, remember every point's exclusive rank for the offsets
template <int COLS>
__global__ void __launch_bounds__(kBlock) compact_emit_kernel(const __grid_constant__ PointOp op, const long long *__restrict__ block_offsets,
                                                             double *__restrict__ out, long long *__restrict__ excl) {
    __shared__ int warp_tot[kBlock / 32];
    const long long i = (long long)blockIdx.x * kBlock + threadIdx.x;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    double a = 0, b = 0, c = 0;
    const bool keep = (i < op.n) && eval_point(op, i, a, b, c);
    const unsigned bal = __ballot_sync(kFullMask, keep);
    if (lane == 0) warp_tot[warp] = __popc(bal);
    __syncthreads();
    int before = __popc(bal & ((1u << lane) - 1u));
    for (int w = 0; w < warp; ++w) before += warp_tot[w];
    const long long pos = block_offsets[blockIdx.x] + before;
    if (i < op.n) excl[i] = pos;
    if (keep) {
        out[COLS * pos] = a;
        out[COLS * pos + 1] = b;
        if (COLS == 3) out[COLS * pos + 2] = c;
    }
}

__global__ void compact_offsets_kernel(const long long *__restrict__ in_offsets, long long n_inst, long long n,
                                       const long long *__restrict__ excl, const long long *__restrict__ total,
                                       long long *__restrict__ out_offsets) {
    const long long j = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (j > n_inst) return;
    const long long idx = in_offsets[j];
    out_offsets[j] = idx >= n ? *total : excl[idx];
}

// ------------------------------------------------------------------------------------------------
// R14.  Painter's loop of 13-px discs == per pixel, the colour of the highest-ordinal instance
// that has a centre within L1 distance 2.  Centres go to a plane padded by
// 2 px so discs centred just outside the image clip exactly like cv2.circle does.
__global__ void __launch_bounds__(kBlock) render_scatter_kernel(const double *__restrict__ vu, long long n, const long long *__restrict__ offsets,
                                                               long long n_inst, unsigned *__restrict__ plane, int height, int width) {
    const long long i = (long long)blockIdx.x * kBlock + threadIdx.x;
    if (i >= n) return;
    bool okv, oku;
    const int cv = trunc_i32(vu[2 * i], okv);
    const int cu = trunc_i32(vu[2 * i + 1], oku);
    if (!okv || !oku) return;
    if (cv < -2 || cv >= height + 2 || cu < -2 || cu >= width + 2) return;
    // instance owning row i: last j with offsets[j] <= i
    long long lo = 0, hi = n_inst;     // invariant: offsets[lo] <= i < offsets[hi]
    while (hi - lo > 1) {
        const long long mid = (lo + hi) >> 1;
        if (offsets[mid] <= i) lo = mid; else hi = mid;
    }
    atomicMax(&plane[(size_t)(cv + 2) * (width + 4) + (cu + 2)], (unsigned)(lo + 1));
}

__global__ void __launch_bounds__(kBlock) render_dilate_kernel(const unsigned *__restrict__ plane, const uint8_t *__restrict__ inst_bgr,
                                                              uint8_t *__restrict__ image, int height, int width) {
    const long long p = (long long)blockIdx.x * kBlock + threadIdx.x;
    if (p >= (long long)height * width) return;
    const int y = (int)(p / width), x = (int)(p % width);
    const int pw = width + 4;
    const unsigned *c = plane + (size_t)(y + 2) * pw + (x + 2);
    unsigned m = c[0];
    m = max(m, max(max(c[-1], c[1]), max(c[-2], c[2])));
    m = max(m, max(max(c[-pw - 1], c[-pw]), c[-pw + 1]));
    m = max(m, max(max(c[pw - 1], c[pw]), c[pw + 1]));
    m = max(m, max(c[-2 * pw], c[2 * pw]));
    if (m) {
        const uint8_t *col = inst_bgr + 3 * (size_t)(m - 1);
        uint8_t *px = image + 3 * (size_t)p;
        px[0] = col[0]; px[1] = col[1]; px[2] = col[2];
    }
}

// The same dilation, but instead of writing into a device image: one thread per 8-pixel chunk, and every chunk with a
// painted pixel becomes a cama_overlay_record {chunk, mask, 24 BGR bytes} (one returning atomic per warp).  A host
// image then gets exactly the bytes render_dilate_kernel would have changed (cama_overlay_apply_host, DRAW) without
// crossing PCIe twice.
__global__ void __launch_bounds__(kBlock) render_chunks_kernel(const unsigned *__restrict__ plane, const uint8_t *__restrict__ inst_bgr, int height, int width,
                                                              cama_overlay_record *__restrict__ records, unsigned *__restrict__ count, long long capacity) {
    const long long q = (long long)blockIdx.x * kBlock + threadIdx.x;
    const int chunks_per_row = width >> 3;
    const bool in_range = q < (long long)height * chunks_per_row;
    cama_overlay_record rec;
    rec.chunk = (uint32_t)q;
    rec.mask = 0u;
    if (in_range) {
        const int y = (int)(q / chunks_per_row), x0 = (int)(q % chunks_per_row) * 8;
        const int pw = width + 4;
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            const unsigned *c = plane + (size_t)(y + 2) * pw + (x0 + k + 2);
            unsigned m = c[0];
            m = max(m, max(max(c[-1], c[1]), max(c[-2], c[2])));
            m = max(m, max(max(c[-pw - 1], c[-pw]), c[-pw + 1]));
            m = max(m, max(max(c[pw - 1], c[pw]), c[pw + 1]));
            m = max(m, max(c[-2 * pw], c[2 * pw]));
            rec.bgr[3 * k] = rec.bgr[3 * k + 1] = rec.bgr[3 * k + 2] = 0;
            if (m) {
                const uint8_t *col = inst_bgr + 3 * (size_t)(m - 1);
                rec.bgr[3 * k] = col[0]; rec.bgr[3 * k + 1] = col[1]; rec.bgr[3 * k + 2] = col[2];
                rec.mask |= 1u << k;
            }
        }
    }
    const bool lit = rec.mask != 0u;
    const unsigned votes = __ballot_sync(0xffffffffu, lit);
    if (votes == 0u) return;
    const int lane = threadIdx.x & 31, leader = __ffs(votes) - 1;
    unsigned base = 0;
    if (lane == leader) base = atomicAdd(count, (unsigned)__popc(votes));
    base = __shfl_sync(0xffffffffu, base, leader);
    const long long slot = (long long)base + __popc(votes & ((1u << lane) - 1u));
    if (lit && slot < capacity) records[slot] = rec;
}

static int fill_op(PointOp &op, const void *pts, int is_f32, int64_t n, const double *T) {
    op.pts = pts;
    op.n = n;
    op.is_f32 = is_f32 != 0;
    op.has_T = T != nullptr;
    op.project = 0;
    op.width = op.height = 0;
    for (int i = 0; i < 12; ++i) op.T[i] = T ? T[i] : 0.0;
    for (int i = 0; i < 6; ++i) op.box[i] = 0.0;
    for (int i = 0; i < 9; ++i) op.K[i] = 0.0;
    return CAMA_OK;
}

static int run_compaction(cama_ctx *ctx, const PointOp &op, int cols, const int64_t *in_offsets, int64_t n_inst,
                          double *out, int64_t *out_offsets, void *workspace, size_t workspace_bytes, cudaStream_t s) {
    size_t need = 0;
    cama_compact_workspace_bytes(op.n, &need);
    if (workspace_bytes < need || (need && !workspace))
        return fail(CAMA_E_WORKSPACE, "compaction workspace: need %zu bytes, got %zu", need, workspace_bytes);
    const long long n = op.n;
    const long long nb = (n + kBlock - 1) / kBlock;
    long long *block_counts = static_cast<long long *>(workspace);
    long long *excl = block_counts + (nb + 1);
    if (n > 0) {
        compact_count_kernel<<<(unsigned)nb, kBlock, 0, s>>>(op, block_counts);
        CAMA_LAUNCHED(ctx);
    }
    scan_i64_kernel<<<1, 1024, 0, s>>>(block_counts, nb);
    CAMA_LAUNCHED(ctx);
    if (n > 0) {
        if (cols == 3) compact_emit_kernel<3><<<(unsigned)nb, kBlock, 0, s>>>(op, block_counts, out, excl);
        else compact_emit_kernel<2><<<(unsigned)nb, kBlock, 0, s>>>(op, block_counts, out, excl);
        CAMA_LAUNCHED(ctx);
    }
    const unsigned ob = (unsigned)((n_inst + 1 + kBlock - 1) / kBlock);
    compact_offsets_kernel<<<ob, kBlock, 0, s>>>(reinterpret_cast<const long long *>(in_offsets), n_inst, n, excl,
                                                 block_counts + nb, reinterpret_cast<long long *>(out_offsets));
    CAMA_LAUNCHED(ctx);
    return CAMA_OK;
}

}  // namespace cama

using namespace cama;

extern "C" {

int cama_abi_version(void) { return CAMA_ABI_VERSION; }

const char *cama_last_error(void) { return g_last_error; }

int cama_device_count(int *count) {
    CAMA_REQUIRE(count, "count is NULL");
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) {
        cudaGetLastError();
        n = 0;
    }
    *count = n;
    return CAMA_OK;
}

int cama_ctx_create(int device, cama_ctx **out) {
    CAMA_REQUIRE(out, "out is NULL");
    *out = nullptr;
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess || n == 0) {
        cudaGetLastError();
        return fail(CAMA_E_NODEVICE, "no CUDA device visible");
    }
    CAMA_REQUIRE(device >= 0 && device < n, "device %d out of range (0..%d)", device, n - 1);
    cudaDeviceProp prop;
    CAMA_CUDA_TRY(cudaGetDeviceProperties(&prop, device));
    if (prop.major != 10)
        return fail(CAMA_E_NODEVICE, "device %d is sm_%d%d; libcama_b200 carries sm_100a code only", device, prop.major, prop.minor);
    cama_ctx *ctx = new cama_ctx();
    ctx->device = device;
    ctx->sm_count = prop.multiProcessorCount;
    ctx->cc_major = prop.major;
    ctx->cc_minor = prop.minor;
    ctx->smem_optin = prop.sharedMemPerBlockOptin;
    *out = ctx;
    return CAMA_OK;
}

int cama_ctx_destroy(cama_ctx *ctx) {
    if (ctx) {
        cama_ctx_profile_enable(ctx, 0);
        DeviceGuard guard(ctx->device);
        for (cudaEvent_t e : ctx->pipe_events) cudaEventDestroy(e);
        for (cudaEvent_t e : ctx->fetch_events) cudaEventDestroy(e);
        for (cudaStream_t s : ctx->pipe_streams)
            if (s) cudaStreamDestroy(s);
    }
    delete ctx;
    return CAMA_OK;
}

int cama_ctx_profile_enable(cama_ctx *ctx, int max_calls) {
    CAMA_REQUIRE(ctx && max_calls >= 0 && max_calls <= (1 << 20), "bad argument");
    DeviceGuard guard(ctx->device);
    for (cudaEvent_t e : ctx->prof_events) cudaEventDestroy(e);
    ctx->prof_events.clear();
    ctx->prof_capacity = 0;
    ctx->prof_calls = 0;
    ctx->prof_events.reserve((size_t)max_calls * (CAMA_CLIP_PHASES + 1));
    for (int i = 0; i < max_calls * (CAMA_CLIP_PHASES + 1); ++i) {
        cudaEvent_t e;
        CAMA_CUDA_TRY(cudaEventCreate(&e));
        ctx->prof_events.push_back(e);
    }
    ctx->prof_capacity = max_calls;
    return CAMA_OK;
}

int cama_ctx_profile_calls(const cama_ctx *ctx, int *calls) {
    CAMA_REQUIRE(ctx && calls, "NULL argument");
    *calls = ctx->prof_calls;
    return CAMA_OK;
}

int cama_ctx_profile_read(cama_ctx *ctx, int call, float *phase_ms) {
    CAMA_REQUIRE(ctx && phase_ms, "NULL argument");
    CAMA_REQUIRE(call >= 0 && call < ctx->prof_calls, "call %d was not recorded (%d recorded)", call, ctx->prof_calls);
    DeviceGuard guard(ctx->device);
    cudaEvent_t *e = ctx->prof_events.data() + (size_t)call * (CAMA_CLIP_PHASES + 1);
    CAMA_CUDA_TRY(cudaEventSynchronize(e[CAMA_CLIP_PHASES]));
    for (int p = 0; p < CAMA_CLIP_PHASES; ++p) CAMA_CUDA_TRY(cudaEventElapsedTime(&phase_ms[p], e[p], e[p + 1]));
    return CAMA_OK;
}

int cama_ctx_launch_count(const cama_ctx *ctx, uint64_t *count) {
    CAMA_REQUIRE(ctx && count, "NULL argument");
    *count = ctx->launches.load(std::memory_order_relaxed);
    return CAMA_OK;
}

int cama_ctx_sm_count(const cama_ctx *ctx, int *count) {
    CAMA_REQUIRE(ctx && count, "NULL argument");
    *count = ctx->sm_count;
    return CAMA_OK;
}

int cama_transform_points(cama_ctx *ctx, const void *pts, int pts_is_f32, int64_t n, const double *T, double *out, void *stream) {
    CAMA_REQUIRE(ctx && T, "NULL argument");
    CAMA_REQUIRE(n >= 0, "negative n");
    if (n == 0) return CAMA_OK;
    CAMA_REQUIRE(pts && out, "NULL buffer");
    DeviceGuard guard(ctx->device);
    PointOp op;
    fill_op(op, pts, pts_is_f32, n, T);
    transform_kernel<<<(unsigned)((n + kBlock - 1) / kBlock), kBlock, 0, (cudaStream_t)stream>>>(op, out);
    CAMA_LAUNCHED(ctx);
    return CAMA_OK;
}

int cama_compact_workspace_bytes(int64_t n, size_t *bytes) {
    CAMA_REQUIRE(bytes && n >= 0, "bad argument");
    const size_t nb = (size_t)((n + kBlock - 1) / kBlock);
    *bytes = align_up(sizeof(long long) * (nb + 1 + (size_t)n), 256);
    return CAMA_OK;
}

int cama_crop_points(cama_ctx *ctx, const void *pts, int pts_is_f32, int64_t n, const double *T, const double *box,
                     const int64_t *in_offsets, int64_t n_inst, double *out_pts, int64_t *out_offsets, void *workspace,
                     size_t workspace_bytes, void *stream) {
    CAMA_REQUIRE(ctx && box && in_offsets && out_offsets, "NULL argument");
    CAMA_REQUIRE(n >= 0 && n_inst >= 0, "negative size");
    CAMA_REQUIRE(n == 0 || (pts && out_pts), "NULL buffer");
    CAMA_REQUIRE(T || !pts_is_f32, "float32 points need a transform (the reference always widens in R9)");
    DeviceGuard guard(ctx->device);
    PointOp op;
    fill_op(op, pts, pts_is_f32, n, T);
    for (int i = 0; i < 6; ++i) op.box[i] = box[i];
    return run_compaction(ctx, op, 3, in_offsets, n_inst, out_pts, out_offsets, workspace, workspace_bytes, (cudaStream_t)stream);
}

int cama_project_points(cama_ctx *ctx, const double *pts, int64_t n, const double *T, const double *K, int width, int height,
                        const int64_t *in_offsets, int64_t n_inst, double *out_vu, int64_t *out_offsets, void *workspace,
                        size_t workspace_bytes, void *stream) {
    CAMA_REQUIRE(ctx && K && in_offsets && out_offsets, "NULL argument");
    CAMA_REQUIRE(n >= 0 && n_inst >= 0 && width > 0 && height > 0, "bad size");
    CAMA_REQUIRE(n == 0 || (pts && out_vu), "NULL buffer");
    DeviceGuard guard(ctx->device);
    PointOp op;
    fill_op(op, pts, 0, n, T);
    op.project = 1;
    op.width = width;
    op.height = height;
    for (int i = 0; i < 9; ++i) op.K[i] = K[i];
    return run_compaction(ctx, op, 2, in_offsets, n_inst, out_vu, out_offsets, workspace, workspace_bytes, (cudaStream_t)stream);
}

int cama_render_workspace_bytes(int height, int width, size_t *bytes) {
    CAMA_REQUIRE(bytes && height > 0 && width > 0, "bad argument");
    *bytes = align_up(sizeof(unsigned) * (size_t)(height + 4) * (size_t)(width + 4), 256);
    return CAMA_OK;
}

int cama_render_points(cama_ctx *ctx, const double *vu, int64_t n, const int64_t *in_offsets, int64_t n_inst,
                       const uint8_t *inst_bgr, uint8_t *image, int height, int width, void *workspace, size_t workspace_bytes,
                       void *stream) {
    CAMA_REQUIRE(ctx && image, "NULL argument");
    CAMA_REQUIRE(n >= 0 && n_inst >= 0 && width > 0 && height > 0, "bad size");
    if (n == 0) return CAMA_OK;
    CAMA_REQUIRE(vu && in_offsets && inst_bgr && n_inst > 0, "NULL buffer");
    CAMA_REQUIRE(n_inst < (int64_t)UINT_MAX, "too many instances");
    size_t need = 0;
    cama_render_workspace_bytes(height, width, &need);
    if (workspace_bytes < need || !workspace) return fail(CAMA_E_WORKSPACE, "render workspace: need %zu bytes, got %zu", need, workspace_bytes);
    DeviceGuard guard(ctx->device);
    cudaStream_t s = (cudaStream_t)stream;
    unsigned *plane = static_cast<unsigned *>(workspace);
    CAMA_CUDA_TRY(cudaMemsetAsync(plane, 0, sizeof(unsigned) * (size_t)(height + 4) * (width + 4), s));
    render_scatter_kernel<<<(unsigned)((n + kBlock - 1) / kBlock), kBlock, 0, s>>>(vu, n, reinterpret_cast<const long long *>(in_offsets), n_inst,
                                                                                  plane, height, width);
    CAMA_LAUNCHED(ctx);
    const long long px = (long long)height * width;
    render_dilate_kernel<<<(unsigned)((px + kBlock - 1) / kBlock), kBlock, 0, s>>>(plane, inst_bgr, image, height, width);
    CAMA_LAUNCHED(ctx);
    return CAMA_OK;
}

int cama_render_points_overlay(cama_ctx *ctx, const double *vu, int64_t n, const int64_t *in_offsets, int64_t n_inst, const uint8_t *inst_bgr,
                               int height, int width, cama_overlay_record *records, uint32_t *count, int64_t capacity, void *workspace,
                               size_t workspace_bytes, void *stream) {
    CAMA_REQUIRE(ctx && count, "NULL argument");
    CAMA_REQUIRE(n >= 0 && n_inst >= 0 && width > 0 && height > 0 && capacity >= 0, "bad size");
    CAMA_REQUIRE(width % 8 == 0, "the overlay form needs a width that is a multiple of 8 (chunks never straddle a row)");
    DeviceGuard guard(ctx->device);
    cudaStream_t s = (cudaStream_t)stream;
    CAMA_CUDA_TRY(cudaMemsetAsync(count, 0, sizeof(uint32_t), s));
    if (n == 0) return CAMA_OK;
    CAMA_REQUIRE(vu && in_offsets && inst_bgr && n_inst > 0 && (records || capacity == 0), "NULL buffer");
    CAMA_REQUIRE(n_inst < (int64_t)UINT_MAX, "too many instances");
    CAMA_REQUIRE(((uintptr_t)records & 15) == 0, "records must be 16-byte aligned");
    size_t need = 0;
    cama_render_workspace_bytes(height, width, &need);
    if (workspace_bytes < need || !workspace) return fail(CAMA_E_WORKSPACE, "render workspace: need %zu bytes, got %zu", need, workspace_bytes);
    unsigned *plane = static_cast<unsigned *>(workspace);
    CAMA_CUDA_TRY(cudaMemsetAsync(plane, 0, sizeof(unsigned) * (size_t)(height + 4) * (width + 4), s));
    render_scatter_kernel<<<(unsigned)((n + kBlock - 1) / kBlock), kBlock, 0, s>>>(vu, n, reinterpret_cast<const long long *>(in_offsets), n_inst,
                                                                                  plane, height, width);
    CAMA_LAUNCHED(ctx);
    const long long chunks = (long long)height * (width / 8);
    render_chunks_kernel<<<(unsigned)((chunks + kBlock - 1) / kBlock), kBlock, 0, s>>>(plane, inst_bgr, height, width, records, count, capacity);
    CAMA_LAUNCHED(ctx);
    return CAMA_OK;
}

}  // extern "C"
