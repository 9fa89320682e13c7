// libcama_b200: LiDAR aggregation (BASELINE.json configs[4]) — every point of
// every sweep is moved lidar -> chassis -> world with the sweep's pose and counted into a voxel grid.
//
// The reference snapshot only has the inputs of this step: the sweep reader
// (/root/reference/cama/dataset_reader.py:45-51: (n,6) float64 rows "x y z intensity ring timestamp"), the
// calibration lookup (dataset_reader.py:222-248), the pose lookup (pose_transformer.py:589-652) and the rigid
// transform of points (reproject.py:108-116).  The voxel accumulation itself lives on its camav2 branch
// (README.md:17-20), so its definition here is OURS (oracle/lidar_oracle.py restates it in NumPy): parity unpinned.
//   world = (T @ [x y z 1]^T)[:3]                 float64, accumulation order of NumPy's matmul (geom.cuh)
//   index = floor((world - origin) / voxel)       float64 subtraction, IEEE division, floor; per axis
//   counts[iz, iy, ix] += 1  when 0 <= index < dims on every axis (non-finite coordinates never are)
#include "common.cuh"
#include "geom.cuh"

using namespace cama;

namespace {

struct VoxelGrid {
    double origin[3], voxel[3];
    int dims[3];          // nx, ny, nz
};

__device__ __forceinline__ bool voxel_axis(double w, double origin, double voxel, int dim, int &i) {
    const double q = floor(__ddiv_rn(__dsub_rn(w, origin), voxel));
    const bool ok = (q >= 0.0) & (q < (double)dim);          // false for NaN
    i = ok ? __double2int_rz(q) : 0;
    return ok;
}

// grid.y = sweep, grid.x strides over the sweep's points; the sweep's 3x4 transform sits in shared memory
__global__ void __launch_bounds__(256) lidar_accumulate_kernel(const double *__restrict__ points, int row_doubles, const long long *__restrict__ sweep_offsets,
                                                             const double *__restrict__ transforms, const __grid_constant__ VoxelGrid g,
                                                             unsigned *__restrict__ counts, unsigned long long *__restrict__ n_inside) {
    __shared__ double T[12];
    const int sweep = blockIdx.y;
    if (threadIdx.x < 12) T[threadIdx.x] = transforms[(size_t)sweep * 16 + threadIdx.x];
    __syncthreads();
    const long long lo = sweep_offsets[sweep], hi = sweep_offsets[sweep + 1];
    unsigned inside = 0;
    for (long long i = lo + (long long)blockIdx.x * 256 + threadIdx.x; i < hi; i += (long long)gridDim.x * 256) {
        const double *p = points + (size_t)i * row_doubles;
        const double x = p[0], y = p[1], z = p[2];
        const double wx = affine_row(T, x, y, z), wy = affine_row(T + 4, x, y, z), wz = affine_row(T + 8, x, y, z);
        int ix, iy, iz;
        const bool ok = voxel_axis(wx, g.origin[0], g.voxel[0], g.dims[0], ix) & voxel_axis(wy, g.origin[1], g.voxel[1], g.dims[1], iy) &
                        voxel_axis(wz, g.origin[2], g.voxel[2], g.dims[2], iz);
        if (ok) {
            atomicAdd(&counts[((size_t)iz * g.dims[1] + iy) * g.dims[0] + ix], 1u);
            ++inside;
        }
    }
    if (n_inside) {
        for (int d = 16; d > 0; d >>= 1) inside += __shfl_down_sync(0xffffffffu, inside, d);
        if ((threadIdx.x & 31) == 0 && inside) atomicAdd(n_inside, (unsigned long long)inside);
    }
}

}  // namespace

extern "C" int cama_lidar_accumulate(cama_ctx *ctx, const double *points, int row_doubles, const int64_t *sweep_offsets, int n_sweeps,
                                     const double *transforms, const cama_voxel_grid *grid, uint32_t *counts, uint64_t *n_inside, void *stream) {
    CAMA_REQUIRE(ctx && grid, "NULL argument");
    CAMA_REQUIRE(n_sweeps >= 0 && row_doubles >= 3, "bad shape");
    CAMA_REQUIRE(grid->dims[0] > 0 && grid->dims[1] > 0 && grid->dims[2] > 0, "voxel grid dims must be positive");
    CAMA_REQUIRE(grid->voxel[0] > 0.0 && grid->voxel[1] > 0.0 && grid->voxel[2] > 0.0, "voxel sizes must be positive");
    CAMA_REQUIRE((long long)grid->dims[0] * grid->dims[1] * grid->dims[2] < (1ll << 40), "voxel grid too large");
    if (n_sweeps == 0) return CAMA_OK;
    CAMA_REQUIRE(points && sweep_offsets && transforms && counts, "NULL buffer");
    CAMA_REQUIRE(n_sweeps <= 65535, "at most 65535 sweeps per call");
    DeviceGuard guard(ctx->device);
    VoxelGrid g;
    for (int k = 0; k < 3; ++k) { g.origin[k] = grid->origin[k]; g.voxel[k] = grid->voxel[k]; g.dims[k] = grid->dims[k]; }
    const dim3 blocks((unsigned)std::max(1, ctx->sm_count * 8 / std::max(1, n_sweeps)), (unsigned)n_sweeps, 1);
    lidar_accumulate_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(points, row_doubles, reinterpret_cast<const long long *>(sweep_offsets), transforms, g, counts,
                                                                     reinterpret_cast<unsigned long long *>(n_inside));
    CAMA_LAUNCHED(ctx);
    return CAMA_OK;
}
