// libcama_b200: host side of the sparse overlay output — draws the lit 8-pixel chunks the GPU
// produced into host frames (the in-place draw of /root/reference/cama/reproject.py:246-257 for
// pixels whose colour is already decided).  Pure byte movement, OpenMP over the records.
#include <cstring>
#include <omp.h>

#include "common.cuh"

using namespace cama;

extern "C" int cama_overlay_apply_host(const cama_overlay_record *records, int64_t n, uint8_t *frames, int64_t n_chunks,
                                       int op, int n_threads) {
    CAMA_REQUIRE(n >= 0 && n_chunks >= 0, "negative size");
    CAMA_REQUIRE(op >= CAMA_OVERLAY_DRAW && op <= CAMA_OVERLAY_BLANK_CHUNKS, "bad op");
    if (n == 0) return CAMA_OK;
    CAMA_REQUIRE(records && frames, "NULL buffer");
    int threads = n_threads > 0 ? n_threads : omp_get_max_threads();
    if (n < 4096) threads = 1;
    constexpr int64_t kAhead = 24;             // the destination lines are scattered: prefetch them for writing
#pragma omp parallel for num_threads(threads) schedule(static)
    for (int64_t i = 0; i < n; ++i) {
        if (i + kAhead < n) {
            const uint32_t c = records[i + kAhead].chunk;
            if ((int64_t)c < n_chunks) __builtin_prefetch(frames + (size_t)c * 24, 1, 0);
        }
        const cama_overlay_record &r = records[i];
        if ((int64_t)r.chunk >= n_chunks) continue;
        uint8_t *dst = frames + (size_t)r.chunk * 24;
        const unsigned mask = r.mask & 0xffu;
        if (op == CAMA_OVERLAY_DRAW_CHUNKS || (op == CAMA_OVERLAY_DRAW && mask == 0xffu)) {
            memcpy(dst, r.bgr, 24);
        } else if (op == CAMA_OVERLAY_BLANK_CHUNKS || (op == CAMA_OVERLAY_BLANK && mask == 0xffu)) {
            memset(dst, 0, 24);
        } else {
            const bool blank = op == CAMA_OVERLAY_BLANK;
            for (int k = 0; k < 8; ++k) {
                if ((mask >> k) & 1u) {
                    dst[3 * k] = blank ? 0 : r.bgr[3 * k];
                    dst[3 * k + 1] = blank ? 0 : r.bgr[3 * k + 1];
                    dst[3 * k + 2] = blank ? 0 : r.bgr[3 * k + 2];
                }
            }
        }
    }
    return CAMA_OK;
}

// Same, into the 2x3-style camera mosaic the reference builds with np.concatenate before encoding
// (/root/reference/cama/tools.py:22-25): frame f is one [rows*H, cols*W, 3] image, camera c sits at
// tile tile_of_cam[c] (row-major).  Writing there directly makes concate_image a no-op.
extern "C" int cama_overlay_apply_host_mosaic(const cama_overlay_record *records, int64_t n, uint8_t *mosaic, int64_t n_frames,
                                              int n_cams, int height, int width, int grid_cols, const int32_t *tile_of_cam,
                                              int op, int n_threads) {
    CAMA_REQUIRE(n >= 0 && n_frames >= 0 && n_cams > 0 && height > 0 && width > 0 && grid_cols > 0, "bad size");
    CAMA_REQUIRE(width % 8 == 0, "width must be a multiple of 8");
    CAMA_REQUIRE(op >= CAMA_OVERLAY_DRAW && op <= CAMA_OVERLAY_BLANK_CHUNKS, "bad op");
    if (n == 0) return CAMA_OK;
    CAMA_REQUIRE(records && mosaic && tile_of_cam, "NULL buffer");
    const int grid_rows = (n_cams + grid_cols - 1) / grid_cols;
    for (int c = 0; c < n_cams; ++c) CAMA_REQUIRE(tile_of_cam[c] >= 0 && tile_of_cam[c] < grid_rows * grid_cols, "tile_of_cam[%d] out of range", c);
    const int64_t chunks_per_row = width / 8, chunks_per_image = chunks_per_row * height;
    const int64_t n_chunks = n_frames * n_cams * chunks_per_image;
    const size_t mosaic_pitch = (size_t)grid_cols * width * 3, mosaic_frame = mosaic_pitch * grid_rows * height;
    int threads = n_threads > 0 ? n_threads : omp_get_max_threads();
    if (n < 4096) threads = 1;
#pragma omp parallel for num_threads(threads) schedule(static)
    for (int64_t i = 0; i < n; ++i) {
        const cama_overlay_record &r = records[i];
        if ((int64_t)r.chunk >= n_chunks) continue;
        const int64_t image = r.chunk / chunks_per_image, within = r.chunk % chunks_per_image;
        const int64_t f = image / n_cams;
        const int tile = tile_of_cam[image % n_cams];
        const int64_t y = within / chunks_per_row, x = (within % chunks_per_row) * 8;
        uint8_t *dst = mosaic + (size_t)f * mosaic_frame + ((size_t)(tile / grid_cols) * height + y) * mosaic_pitch +
                       ((size_t)(tile % grid_cols) * width + x) * 3;
        const unsigned mask = r.mask & 0xffu;
        if (op == CAMA_OVERLAY_DRAW_CHUNKS || (op == CAMA_OVERLAY_DRAW && mask == 0xffu)) {
            memcpy(dst, r.bgr, 24);
        } else if (op == CAMA_OVERLAY_BLANK_CHUNKS || (op == CAMA_OVERLAY_BLANK && mask == 0xffu)) {
            memset(dst, 0, 24);
        } else {
            const bool blank = op == CAMA_OVERLAY_BLANK;
            for (int k = 0; k < 8; ++k) {
                if ((mask >> k) & 1u) {
                    dst[3 * k] = blank ? 0 : r.bgr[3 * k];
                    dst[3 * k + 1] = blank ? 0 : r.bgr[3 * k + 1];
                    dst[3 * k + 2] = blank ? 0 : r.bgr[3 * k + 2];
                }
            }
        }
    }
    return CAMA_OK;
}
