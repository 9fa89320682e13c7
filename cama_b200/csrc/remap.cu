// libcama_b200: undistort-resize of the camera images — "next" row N1 of the scope table:
// CameraManager.resize_image of /root/reference/cama/reproject.py:232-240, i.e. cv2.remap with
// float32 maps, INTER_LINEAR, BORDER_CONSTANT(0), on uint8 BGR images.
//
// OpenCV's 8-bit bilinear remap is fixed point, and this kernel restates it exactly:
//   sx = cvRound(map_x * 32), sy = cvRound(map_y * 32)          (round half to even, float multiply)
//   x = sx >> 5, y = sy >> 5, fx = sx & 31, fy = sy & 31
//   w = {(32-fx)(32-fy), fx(32-fy), (32-fx)fy, fx fy} * 32      (the INTER_LINEAR table at 15 bits; sums to 32768)
//   dst = (sum_k tap_k * w_k + 16384) >> 15                      (taps outside the source are the border value 0)
// The maps come from cv2.initUndistortRectifyMap, computed ONCE per camera on the host (the
// reference recomputes them for every image) and kept on the device.
#include "common.cuh"

namespace cama {

constexpr int kRemapPx = 4;          // output pixels per thread: 12 bytes = 3 aligned words

__global__ void __launch_bounds__(256) remap_bilinear_kernel(const uint8_t *__restrict__ src, const float *__restrict__ map_x,
                                                            const float *__restrict__ map_y, uint8_t *__restrict__ dst,
                                                            int src_h, int src_w, int dst_h, int dst_w, int n_maps) {
    const long long groups = (long long)dst_h * dst_w / kRemapPx;
    const long long g = (long long)blockIdx.x * 256 + threadIdx.x;
    if (g >= groups) return;
    const long long image = blockIdx.y;
    const uint8_t *s = src + (size_t)image * src_h * src_w * 3;
    const size_t map_off = (size_t)(image % n_maps) * dst_h * dst_w + g * kRemapPx;      // images cycle through the maps (frame-major, camera-minor)
    const float4 mx = *reinterpret_cast<const float4 *>(map_x + map_off);
    const float4 my = *reinterpret_cast<const float4 *>(map_y + map_off);
    const float mxs[4] = {mx.x, mx.y, mx.z, mx.w}, mys[4] = {my.x, my.y, my.z, my.w};
    unsigned char out[12];
#pragma unroll
    for (int k = 0; k < kRemapPx; ++k) {
        const int sx = __float2int_rn(__fmul_rn(mxs[k], 32.0f)), sy = __float2int_rn(__fmul_rn(mys[k], 32.0f));
        const int x = sx >> 5, y = sy >> 5, fx = sx & 31, fy = sy & 31;
        const int w00 = (32 - fx) * (32 - fy) * 32, w01 = fx * (32 - fy) * 32, w10 = (32 - fx) * fy * 32, w11 = fx * fy * 32;
        const bool x0 = x >= 0 && x < src_w, x1 = x + 1 >= 0 && x + 1 < src_w;
        const bool y0 = y >= 0 && y < src_h, y1 = y + 1 >= 0 && y + 1 < src_h;
        const uint8_t *p00 = s + ((size_t)y * src_w + x) * 3;
        const uint8_t *p10 = p00 + (size_t)src_w * 3;
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            int acc = 1 << 14;
            if (y0 && x0) acc += p00[c] * w00;
            if (y0 && x1) acc += p00[3 + c] * w01;
            if (y1 && x0) acc += p10[c] * w10;
            if (y1 && x1) acc += p10[3 + c] * w11;
            out[3 * k + c] = (unsigned char)(acc >> 15);          // weights sum to 2^15: never above 255
        }
    }
    unsigned *d = reinterpret_cast<unsigned *>(dst + ((size_t)image * dst_h * dst_w + g * kRemapPx) * 3);
    d[0] = out[0] | (out[1] << 8) | (out[2] << 16) | ((unsigned)out[3] << 24);
    d[1] = out[4] | (out[5] << 8) | (out[6] << 16) | ((unsigned)out[7] << 24);
    d[2] = out[8] | (out[9] << 8) | (out[10] << 16) | ((unsigned)out[11] << 24);
}

}  // namespace cama

using namespace cama;

extern "C" int cama_remap_bilinear(cama_ctx *ctx, const uint8_t *src, int64_t n_images, int src_height, int src_width,
                                   const float *map_x, const float *map_y, int64_t n_maps, uint8_t *dst, int dst_height, int dst_width,
                                   void *stream) {
    CAMA_REQUIRE(ctx, "ctx is NULL");
    CAMA_REQUIRE(n_images >= 0 && src_height > 0 && src_width > 0 && dst_height > 0 && dst_width > 0, "bad size");
    CAMA_REQUIRE(n_maps >= 1 && n_maps <= n_images + (n_images == 0), "n_maps must be in 1..n_images");
    CAMA_REQUIRE(dst_width % 4 == 0, "dst_width must be a multiple of 4");
    CAMA_REQUIRE(n_images <= 65535, "at most 65535 images per call");
    if (n_images == 0) return CAMA_OK;
    CAMA_REQUIRE(src && map_x && map_y && dst, "NULL buffer");
    CAMA_REQUIRE(((uintptr_t)map_x & 15) == 0 && ((uintptr_t)map_y & 15) == 0 && ((uintptr_t)dst & 3) == 0, "maps must be 16-byte, dst 4-byte aligned");
    DeviceGuard guard(ctx->device);
    const long long groups = (long long)dst_height * dst_width / kRemapPx;
    remap_bilinear_kernel<<<dim3((unsigned)((groups + 255) / 256), (unsigned)n_images), 256, 0, (cudaStream_t)stream>>>(
        src, map_x, map_y, dst, src_height, src_width, dst_height, dst_width, (int)n_maps);
    CAMA_LAUNCHED(ctx);
    return CAMA_OK;
}
