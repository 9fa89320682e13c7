// Device math shared by every kernel of libcama_b200.  All arithmetic that decides a pixel is
// written with explicit round-to-nearest intrinsics so that nvcc can neither contract nor split
// it: the accumulation order is the one NumPy/OpenBLAS dgemm uses for the reference's
// `T @ P.T` (/root/reference/cama/reproject.py:114,191), i.e. a0*b0 first, then one fused
// multiply-add per further term, in index order (DESIGN.md "Numerics").
#pragma once
#include <cfloat>
#include <cstdint>

namespace cama {

// (row of a 4x4) . [x y z 1]   — reference cama/reproject.py:113-114
__device__ __forceinline__ double affine_row(const double *m, double x, double y, double z) {
    double a = __dmul_rn(m[0], x);
    a = __fma_rn(m[1], y, a);
    a = __fma_rn(m[2], z, a);
    return __dadd_rn(a, m[3]);          // == fma(m3, 1.0, a)
}
__device__ __forceinline__ double affine_row4(double m0, double m1, double m2, double m3, double x, double y, double z) {
    double a = __dmul_rn(m0, x);
    a = __fma_rn(m1, y, a);
    a = __fma_rn(m2, z, a);
    return __dadd_rn(a, m3);
}

// (row of a 3x3) . [x y z]   — reference cama/reproject.py:191
__device__ __forceinline__ double linear_row(const double *k, double x, double y, double z) {
    double a = __dmul_rn(k[0], x);
    a = __fma_rn(k[1], y, a);
    return __fma_rn(k[2], z, a);
}

// inclusive crop box {x_min,x_max,y_min,y_max,z_min,z_max} — reference cama/reproject.py:123-125
__device__ __forceinline__ bool in_box(const double *b, double x, double y, double z) {
    return (x >= b[0]) & (x <= b[1]) & (y >= b[2]) & (y <= b[3]) & (z >= b[4]) & (z <= b[5]);
}

// Pinhole projection and visibility of one camera-frame point — reference cama/reproject.py:191-198.
//   q = K p ; mask_z = q_z > 0 ; q /= q_z ; keep = (q_z/q_z > 0) & 0<=u<W & 0<=v<H & mask_z
// q_z/q_z > 0 holds exactly when q_z is finite and non-zero, so with mask_z: 0 < q_z <= DBL_MAX.
__device__ __forceinline__ bool project_point(const double *K, double x, double y, double z, int width,
                                              int height, double &v, double &u) {
    const double qx = linear_row(K, x, y, z);
    const double qy = linear_row(K + 3, x, y, z);
    const double qz = linear_row(K + 6, x, y, z);
    u = __ddiv_rn(qx, qz);
    v = __ddiv_rn(qy, qz);
    return (qz > 0.0) & (qz <= DBL_MAX) & (u >= 0.0) & (u < (double)width) & (v >= 0.0) & (v < (double)height);
}

// astype(np.int32) of the reference (cama/reproject.py:249): truncation; values that do not fit
// (and NaN) become INT32_MIN there, i.e. a centre that can never touch the image -> `ok` false.
__device__ __forceinline__ int trunc_i32(double a, bool &ok) {
    ok = (a > -2147483649.0) & (a < 2147483648.0);
    return ok ? __double2int_rz(a) : INT32_MIN;
}

}  // namespace cama
