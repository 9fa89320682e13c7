// libcama_b200: load-time densify of map polylines on the device — "next" row N2 of the scope
// table: MapManager.load_3d_instance_maps / calculate_3d_instance_maps of
// /root/reference/cama/reproject.py:42-106 (+ pixel2world_xy :36-40).
//
// Float32 throughout, every operation individually rounded exactly like the reference's NumPy
// expression (NEP-50 keeps python-float scalars weak, so `length / 0.1` and `delta / num * j` stay
// float32):   num = int(sqrt(dx*dx + dy*dy) / 0.1f)       per segment, segments with num == 0 dropped
//             p_j = start + (end - start) / num * j        j = 0 .. num-1 (segment end never emitted)
// pixel labels additionally: cell = clip(uint16(round(p))[::-1], 0, rows-1); z = bev[cell];
//             x = p1 * 0.1f - 300 + 0;  y = p0 * 0.1f - 300 + 0.
// The library is built with -fmad=false, and the intrinsics below pin the rounding anyway.
#include "common.cuh"

namespace cama {

// dense points emitted by the segment that starts at raw vertex i (0 for the last vertex of a polyline)
__global__ void __launch_bounds__(256) densify_count_kernel(const float2 *__restrict__ raw, const int *__restrict__ raw_poly, long long n_raw,
                                                           float resolution, long long *__restrict__ seg_count) {
    const long long i = (long long)blockIdx.x * 256 + threadIdx.x;
    if (i >= n_raw) return;
    long long num = 0;
    if (i + 1 < n_raw && raw_poly[i + 1] == raw_poly[i]) {
        const float2 a = raw[i], b = raw[i + 1];
        const float dx = __fsub_rn(b.x, a.x), dy = __fsub_rn(b.y, a.y);
        const float len = __fsqrt_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)));
        const float q = __fdiv_rn(len, resolution);
        num = (q >= 1.0f && q < 9.0e18f) ? (long long)q : 0;       // astype(int64) truncation; NaN / huge never happen for finite labels
    }
    seg_count[i] = num;
}

struct DensifyArgs {
    const float2 *raw;
    const int *raw_poly;
    const long long *seg_start;      // exclusive scan of seg_count, [n_raw + 1]
    long long n_raw, total;
    const float *bev;                // nullptr: metric labels (z = 0, x,y as they are)
    int bev_rows, bev_cols;
    float solution, half_w, half_h, center_x, center_y;
    float4 *out;
};

__global__ void __launch_bounds__(256) densify_fill_kernel(const DensifyArgs a) {
    const long long i = (long long)blockIdx.x * 256 + threadIdx.x;
    if (i >= a.total) return;
    // segment owning dense point i: last s with seg_start[s] <= i
    long long lo = 0, hi = a.n_raw;
    while (hi - lo > 1) {
        const long long mid = (lo + hi) >> 1;
        if (a.seg_start[mid] <= i) lo = mid; else hi = mid;
    }
    const long long s = lo;
    const float num = (float)(a.seg_start[s + 1] - a.seg_start[s]);
    const float j = (float)(i - a.seg_start[s]);
    const float2 p0 = a.raw[s], p1 = a.raw[s + 1];
    const float px = __fadd_rn(p0.x, __fmul_rn(__fdiv_rn(__fsub_rn(p1.x, p0.x), num), j));
    const float py = __fadd_rn(p0.y, __fmul_rn(__fdiv_rn(__fsub_rn(p1.y, p0.y), num), j));
    float4 v;
    if (a.bev == nullptr) {
        v = make_float4(px, py, 0.0f, 0.0f);
    } else {
        // .round() = half to even; astype(uint16) of the integral value wraps modulo 65536; [:, ::-1]; both indices clipped with rows-1
        const long long r0 = (long long)rintf(py), c0 = (long long)rintf(px);
        int row = (int)(unsigned short)r0, col = (int)(unsigned short)c0;
        row = min(row, a.bev_rows - 1);
        col = min(col, a.bev_rows - 1);
        const float z = a.bev[(size_t)row * a.bev_cols + col];
        const float x = __fadd_rn(__fsub_rn(__fmul_rn(py, a.solution), a.half_w), a.center_x);
        const float y = __fadd_rn(__fsub_rn(__fmul_rn(px, a.solution), a.half_h), a.center_y);
        v = make_float4(x, y, z, 0.0f);
    }
    v.w = __int_as_float(a.raw_poly[s]);
    a.out[i] = v;
}

// exclusive scan of long long counts, single CTA (label sets have thousands of raw vertices)
__global__ void __launch_bounds__(1024) densify_scan_kernel(const long long *__restrict__ counts, long long n, long long *__restrict__ start) {
    __shared__ long long warp_sum[32];
    __shared__ long long carry_s;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid == 0) carry_s = 0;
    __syncthreads();
    for (long long base = 0; base < n; base += 1024) {
        const long long idx = base + tid;
        const long long v = idx < n ? counts[idx] : 0;
        long long inc = v;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const long long t = __shfl_up_sync(0xffffffffu, inc, d);
            if (lane >= d) inc += t;
        }
        if (lane == 31) warp_sum[warp] = inc;
        __syncthreads();
        if (warp == 0) {
            long long w = warp_sum[lane];
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                const long long t = __shfl_up_sync(0xffffffffu, w, d);
                if (lane >= d) w += t;
            }
            warp_sum[lane] = w;
        }
        __syncthreads();
        const long long before = carry_s + (warp > 0 ? warp_sum[warp - 1] : 0) + inc - v;
        if (idx < n) start[idx] = before;
        __syncthreads();
        if (tid == 1023) carry_s = before + v;
        __syncthreads();
    }
    if (tid == 0) start[n] = carry_s;
}

}  // namespace cama

using namespace cama;

extern "C" {

int cama_densify_plan(cama_ctx *ctx, const float *raw_xy, const int32_t *raw_poly, int64_t n_raw, float resolution,
                      int64_t *seg_start, void *stream) {
    CAMA_REQUIRE(ctx && seg_start, "NULL argument");
    CAMA_REQUIRE(n_raw >= 0 && resolution > 0.0f, "bad argument");
    CAMA_REQUIRE(n_raw == 0 || (raw_xy && raw_poly), "NULL buffer");
    DeviceGuard guard(ctx->device);
    cudaStream_t s = (cudaStream_t)stream;
    long long *counts = reinterpret_cast<long long *>(seg_start) + 1;        // counts live in seg_start[1..n_raw], scanned in place below
    if (n_raw > 0) {
        densify_count_kernel<<<(unsigned)((n_raw + 255) / 256), 256, 0, s>>>(reinterpret_cast<const float2 *>(raw_xy), raw_poly, n_raw, resolution, counts);
        CAMA_LAUNCHED(ctx);
    }
    // exclusive scan of counts[0..n) into seg_start[0..n]: reading counts[i] = seg_start[i+1] and writing seg_start[i] never collide
    // within a 1024-wide round because every value of the round is read before any is written
    densify_scan_kernel<<<1, 1024, 0, s>>>(counts, n_raw, reinterpret_cast<long long *>(seg_start));
    CAMA_LAUNCHED(ctx);
    return CAMA_OK;
}

int cama_densify_fill(cama_ctx *ctx, const float *raw_xy, const int32_t *raw_poly, int64_t n_raw, const int64_t *seg_start,
                      int64_t total, const float *bev_height, int bev_rows, int bev_cols, float solution, float half_width,
                      float half_height, float center_x, float center_y, float *out_vertices, void *stream) {
    CAMA_REQUIRE(ctx, "ctx is NULL");
    CAMA_REQUIRE(n_raw >= 0 && total >= 0, "negative size");
    if (total == 0) return CAMA_OK;
    CAMA_REQUIRE(raw_xy && raw_poly && seg_start && out_vertices, "NULL buffer");
    CAMA_REQUIRE(!bev_height || (bev_rows > 0 && bev_cols > 0), "bad height map shape");
    // both indices are clipped with bev_rows - 1, like the reference (cama/reproject.py:98): a map with fewer columns
    // than rows would be read past the end of a row (the reference raises IndexError there)
    CAMA_REQUIRE(!bev_height || bev_cols >= bev_rows, "height map with fewer columns (%d) than rows (%d): the reference's lookup fails on it", bev_cols, bev_rows);
    CAMA_REQUIRE(((uintptr_t)out_vertices & 15) == 0, "out_vertices must be 16-byte aligned");
    DeviceGuard guard(ctx->device);
    DensifyArgs a{};
    a.raw = reinterpret_cast<const float2 *>(raw_xy); a.raw_poly = raw_poly;
    a.seg_start = reinterpret_cast<const long long *>(seg_start);
    a.n_raw = n_raw; a.total = total;
    a.bev = bev_height; a.bev_rows = bev_rows; a.bev_cols = bev_cols;
    a.solution = solution; a.half_w = half_width; a.half_h = half_height; a.center_x = center_x; a.center_y = center_y;
    a.out = reinterpret_cast<float4 *>(out_vertices);
    densify_fill_kernel<<<(unsigned)((total + 255) / 256), 256, 0, (cudaStream_t)stream>>>(a);
    CAMA_LAUNCHED(ctx);
    return CAMA_OK;
}

}  // extern "C"
