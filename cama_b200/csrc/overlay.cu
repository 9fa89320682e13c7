// libcama_b200: host side of the sparse overlay output — draws the lit 8-pixel chunks the GPU
// produced into host frames (the in-place draw of /root/reference/cama/reproject.py:246-257 for
// pixels whose colour is already decided), either plain [F,C,H,W,3] frames or the 2x3 camera mosaic of
// /root/reference/cama/tools.py:22-25.  Pure byte movement, on the library's worker pool (host_pool.h).
#include <algorithm>
#include <chrono>
#include <cstdlib>
#include <cstring>
#include <thread>

#include "host_pool.h"
#if defined(__x86_64__)
#include <tmmintrin.h>
#define CAMA_HAVE_SSSE3 1
#endif

#include "common.cuh"

using namespace cama;

namespace {

constexpr int64_t kHostChunk = 2048;           // records a worker claims at a time (~20 us of drawing)

struct Target {
    uint8_t *pixels;
    int64_t n_chunks, chunks_per_row, chunks_per_image;
    int n_cams, height, width, grid_cols;
    const int32_t *tile_of_cam;
    size_t mosaic_pitch, mosaic_frame;

    inline uint8_t *chunk_ptr(uint32_t chunk) const {
        if (grid_cols == 0) return pixels + (size_t)chunk * 24;
        const int64_t image = chunk / chunks_per_image, within = chunk % chunks_per_image;
        const int tile = tile_of_cam[image % n_cams];
        const int64_t y = within / chunks_per_row, x = (within % chunks_per_row) * 8;
        return pixels + (size_t)(image / n_cams) * mosaic_frame + ((size_t)(tile / grid_cols) * height + y) * mosaic_pitch +
               ((size_t)(tile % grid_cols) * width + x) * 3;
    }
};

inline void apply_one(uint8_t *dst, const uint8_t *bgr, unsigned mask, int op) {
    if (op == CAMA_OVERLAY_DRAW_CHUNKS || (op == CAMA_OVERLAY_DRAW && mask == 0xffu)) {
        memcpy(dst, bgr, 24);
    } else if (op == CAMA_OVERLAY_BLANK_CHUNKS || (op == CAMA_OVERLAY_BLANK && mask == 0xffu)) {
        memset(dst, 0, 24);
    } else {
        const bool blank = op == CAMA_OVERLAY_BLANK;
        for (int k = 0; k < 8; ++k) {
            if ((mask >> k) & 1u) {
                dst[3 * k] = blank ? 0 : bgr[3 * k];
                dst[3 * k + 1] = blank ? 0 : bgr[3 * k + 1];
                dst[3 * k + 2] = blank ? 0 : bgr[3 * k + 2];
            }
        }
    }
}

#ifdef CAMA_HAVE_SSSE3
// Palette records with at most 15 colours (CAMA has two): the 8 index bytes of a record become its 24 BGR bytes with
// byte shuffles — each index replicated three times (one shuffle per 16 output bytes), then looked up in the three
// 16-entry channel tables (one shuffle each) and merged by position.  The scalar path (two pixels per lookup in a
// 64 K-entry table) spent ~40 cycles per record, three times what blanking the same chunks costs.
struct Pal16 {
    __m128i b, g, r;
};
template <bool MASKED>
__attribute__((target("ssse3"))) inline void expand8_ssse3(const Pal16 &pal, const uint8_t *ix, uint8_t *dst) {
    __m128i idx = _mm_loadl_epi64(reinterpret_cast<const __m128i *>(ix));
    idx = _mm_or_si128(idx, _mm_cmpgt_epi8(idx, _mm_set1_epi8(15)));       // entries past the 16-entry tables: black, like the scalar path
    const __m128i rep0 = _mm_shuffle_epi8(idx, _mm_setr_epi8(0, 0, 0, 1, 1, 1, 2, 2, 2, 3, 3, 3, 4, 4, 4, 5));
    const __m128i rep1 = _mm_shuffle_epi8(idx, _mm_setr_epi8(5, 5, 6, 6, 6, 7, 7, 7, -1, -1, -1, -1, -1, -1, -1, -1));
    // output byte p of the chunk holds channel p % 3 (B, G, R) of pixel p / 3
    const __m128i mb0 = _mm_setr_epi8(-1, 0, 0, -1, 0, 0, -1, 0, 0, -1, 0, 0, -1, 0, 0, -1);
    const __m128i mg0 = _mm_setr_epi8(0, -1, 0, 0, -1, 0, 0, -1, 0, 0, -1, 0, 0, -1, 0, 0);
    const __m128i mr0 = _mm_setr_epi8(0, 0, -1, 0, 0, -1, 0, 0, -1, 0, 0, -1, 0, 0, -1, 0);
    const __m128i mb1 = _mm_setr_epi8(0, 0, -1, 0, 0, -1, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0);      // bytes 16..23: G R | B G R | B G R
    const __m128i mg1 = _mm_setr_epi8(-1, 0, 0, -1, 0, 0, -1, 0, 0, 0, 0, 0, 0, 0, 0, 0);
    const __m128i mr1 = _mm_setr_epi8(0, -1, 0, 0, -1, 0, 0, -1, 0, 0, 0, 0, 0, 0, 0, 0);
    __m128i out0 = _mm_or_si128(_mm_or_si128(_mm_and_si128(_mm_shuffle_epi8(pal.b, rep0), mb0), _mm_and_si128(_mm_shuffle_epi8(pal.g, rep0), mg0)),
                                _mm_and_si128(_mm_shuffle_epi8(pal.r, rep0), mr0));
    __m128i out1 = _mm_or_si128(_mm_or_si128(_mm_and_si128(_mm_shuffle_epi8(pal.b, rep1), mb1), _mm_and_si128(_mm_shuffle_epi8(pal.g, rep1), mg1)),
                                _mm_and_si128(_mm_shuffle_epi8(pal.r, rep1), mr1));
    if (MASKED) {                                  // only painted pixels (index != 0) are written: the others keep the image
        const __m128i zero = _mm_setzero_si128();
        const __m128i keep0 = _mm_cmpeq_epi8(rep0, zero), keep1 = _mm_cmpeq_epi8(rep1, zero);
        const __m128i old0 = _mm_loadu_si128(reinterpret_cast<const __m128i *>(dst));
        const __m128i old1 = _mm_loadl_epi64(reinterpret_cast<const __m128i *>(dst + 16));
        out0 = _mm_or_si128(_mm_and_si128(old0, keep0), _mm_andnot_si128(keep0, out0));
        out1 = _mm_or_si128(_mm_and_si128(old1, keep1), _mm_andnot_si128(keep1, out1));
    }
    _mm_storeu_si128(reinterpret_cast<__m128i *>(dst), out0);
    _mm_storel_epi64(reinterpret_cast<__m128i *>(dst + 16), out1);
}

struct DrawPal16Job {
    const cama_overlay_record_palette *rec;
    int64_t n;
    const Target *t;
    Pal16 pal;
};
template <bool MASKED>
__attribute__((target("ssse3"))) void draw_pal16_range(void *ctx, int64_t lo, int64_t hi) {
    const DrawPal16Job &j = *static_cast<const DrawPal16Job *>(ctx);
    const cama_overlay_record_palette *rec = j.rec;
    const Target &t = *j.t;
    constexpr int64_t kAhead = 24;                 // the destination lines are scattered: prefetch them for writing
    for (int64_t i = lo; i < hi; ++i) {
        if (i + kAhead < j.n && (int64_t)rec[i + kAhead].chunk < t.n_chunks) __builtin_prefetch(t.chunk_ptr(rec[i + kAhead].chunk), 1, 0);
        if ((int64_t)rec[i].chunk >= t.n_chunks) continue;
        expand8_ssse3<MASKED>(j.pal, rec[i].index, t.chunk_ptr(rec[i].chunk));
    }
}

template <bool MASKED>
__attribute__((target("ssse3"))) void draw_pal16(const cama_overlay_record_palette *rec, int64_t n, const uint8_t *palette_bgr, const Target &t, int threads) {
    alignas(16) uint8_t tb[16] = {0}, tg[16] = {0}, tr[16] = {0};
    for (int e = 1; e < 16; ++e) { tb[e] = palette_bgr[3 * e]; tg[e] = palette_bgr[3 * e + 1]; tr[e] = palette_bgr[3 * e + 2]; }
    DrawPal16Job job{rec, n, &t, {}};
    job.pal.b = _mm_load_si128(reinterpret_cast<const __m128i *>(tb));
    job.pal.g = _mm_load_si128(reinterpret_cast<const __m128i *>(tg));
    job.pal.r = _mm_load_si128(reinterpret_cast<const __m128i *>(tr));
    HostPool::instance().run(threads, n, kHostChunk, draw_pal16_range<MASKED>, &job);
}
#endif

struct BgrJob {
    const cama_overlay_record *rec;
    int64_t n;
    const Target *t;
    int op;
};
void apply_bgr_range(void *ctx, int64_t lo, int64_t hi) {
    const BgrJob &j = *static_cast<const BgrJob *>(ctx);
    const cama_overlay_record *rec = j.rec;
    const Target &t = *j.t;
    constexpr int64_t kAhead = 24;                 // the destination lines are scattered: prefetch them for writing
    for (int64_t i = lo; i < hi; ++i) {
        if (i + kAhead < j.n && (int64_t)rec[i + kAhead].chunk < t.n_chunks) __builtin_prefetch(t.chunk_ptr(rec[i + kAhead].chunk), 1, 0);
        if ((int64_t)rec[i].chunk >= t.n_chunks) continue;
        apply_one(t.chunk_ptr(rec[i].chunk), rec[i].bgr, rec[i].mask & 0xffu, j.op);
    }
}

struct PalJob {
    const cama_overlay_record_palette *rec;
    int64_t n;
    const Target *t;
    int op;
    const uint64_t *pair_tab;                      // two pixels per lookup: pair[a | b << 8] = the 6 bytes of pixel a followed by pixel b
};
void apply_pal_range(void *ctx, int64_t lo, int64_t hi) {
    const PalJob &j = *static_cast<const PalJob *>(ctx);
    const cama_overlay_record_palette *rec = j.rec;
    const Target &t = *j.t;
    const uint64_t *pair_tab = j.pair_tab;
    const int op = j.op;
    constexpr int64_t kAhead = 24;
    for (int64_t i = lo; i < hi; ++i) {
        if (i + kAhead < j.n && (int64_t)rec[i + kAhead].chunk < t.n_chunks) __builtin_prefetch(t.chunk_ptr(rec[i + kAhead].chunk), 1, 0);
        if ((int64_t)rec[i].chunk >= t.n_chunks) continue;
        uint8_t *dst = t.chunk_ptr(rec[i].chunk);
        if (op == CAMA_OVERLAY_BLANK_CHUNKS) {
            memset(dst, 0, 24);
            continue;
        }
        const uint8_t *ix = rec[i].index;
        uint16_t pr[4];
        memcpy(pr, ix, 8);
        const uint64_t p01 = pair_tab[pr[0]], p23 = pair_tab[pr[1]], p45 = pair_tab[pr[2]], p67 = pair_tab[pr[3]];
        const uint64_t w[3] = {p01 | p23 << 48, p23 >> 16 | p45 << 32, p45 >> 32 | p67 << 16};
        if (op == CAMA_OVERLAY_DRAW_CHUNKS) {              // the 24 bytes, assembled in registers (little endian)
            memcpy(dst, w, 24);
            continue;
        }
        uint8_t bgr[24];
        memcpy(bgr, w, 24);
        unsigned mask = 0;
        for (int k = 0; k < 8; ++k) mask |= (ix[k] ? 1u : 0u) << k;
        apply_one(dst, bgr, mask, op);
    }
}

}  // namespace

// ---- device side: records -> dense frames (the receiving end of a sparse all-gather) --------------------
// One thread per record: 24 bytes (8 BGR pixels) written as three 8-byte stores at chunk * 24.  The frames are
// zero-filled first (cudaMemsetAsync), so the result equals the dense render of the same clip on blank frames.
template <int FORMAT>
__global__ void __launch_bounds__(256) overlay_expand_kernel(const uint32_t *__restrict__ records, long long n, const uint32_t *__restrict__ palette,
                                                            uint8_t *__restrict__ frames, long long n_chunks) {
    const long long i = (long long)blockIdx.x * 256 + threadIdx.x;
    if (i >= n) return;
    uint32_t w[6];
    uint32_t chunk;
    if (FORMAT == CAMA_OVERLAY_BGR) {
        const uint4 a = reinterpret_cast<const uint4 *>(records)[2 * i], b = reinterpret_cast<const uint4 *>(records)[2 * i + 1];
        chunk = a.x;                                   // a.y = mask: unpainted pixels of a record are 0,0,0 already
        w[0] = a.z; w[1] = a.w; w[2] = b.x; w[3] = b.y; w[4] = b.z; w[5] = b.w;
    } else {
        const uint32_t *r = records + 3 * i;
        chunk = r[0];
        const uint32_t lo = r[1], hi = r[2];
        uint32_t c[8];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            c[k] = palette[(lo >> (8 * k)) & 0xffu];
            c[4 + k] = palette[(hi >> (8 * k)) & 0xffu];
        }
        w[0] = __byte_perm(c[0], c[1], 0x4210); w[1] = __byte_perm(c[1], c[2], 0x5421); w[2] = __byte_perm(c[2], c[3], 0x6542);
        w[3] = __byte_perm(c[4], c[5], 0x4210); w[4] = __byte_perm(c[5], c[6], 0x5421); w[5] = __byte_perm(c[6], c[7], 0x6542);
    }
    if ((long long)chunk >= n_chunks) return;
    uint2 *d = reinterpret_cast<uint2 *>(frames + (size_t)chunk * 24);
    d[0] = make_uint2(w[0], w[1]); d[1] = make_uint2(w[2], w[3]); d[2] = make_uint2(w[4], w[5]);
}

__global__ void palette_pack_kernel(const uint8_t *__restrict__ palette_bgr, uint32_t *__restrict__ packed) {
    const int e = threadIdx.x;
    packed[e] = e == 0 ? 0u : (uint32_t)palette_bgr[3 * e] | ((uint32_t)palette_bgr[3 * e + 1] << 8) | ((uint32_t)palette_bgr[3 * e + 2] << 16);
}

extern "C" int cama_overlay_expand(cama_ctx *ctx, const void *records, int64_t n, int format, const uint8_t *palette_bgr, void *palette_scratch,
                                   uint8_t *frames, int64_t n_frames, int n_cams, int height, int width, int zero_first, void *stream) {
    CAMA_REQUIRE(ctx, "ctx is NULL");
    CAMA_REQUIRE(n >= 0 && n_frames >= 0 && n_cams > 0 && height > 0 && width > 0 && width % 8 == 0, "bad shape");
    CAMA_REQUIRE(format == CAMA_OVERLAY_BGR || format == CAMA_OVERLAY_PALETTE, "bad format");
    CAMA_REQUIRE(format != CAMA_OVERLAY_PALETTE || (palette_bgr && palette_scratch), "the palette format needs palette_bgr and palette_scratch (device)");
    const int64_t n_chunks = n_frames * n_cams * height * (width / 8);
    if (n_chunks == 0) return CAMA_OK;
    CAMA_REQUIRE(frames && ((uintptr_t)frames & 7) == 0, "frames must be 8-byte aligned");
    CAMA_REQUIRE(n == 0 || (records && ((uintptr_t)records & (format == CAMA_OVERLAY_BGR ? 15 : 3)) == 0), "records must be 16-byte (BGR) / 4-byte (palette) aligned");
    DeviceGuard guard(ctx->device);
    cudaStream_t s = (cudaStream_t)stream;
    if (zero_first) CAMA_CUDA_TRY(cudaMemsetAsync(frames, 0, (size_t)n_chunks * 24, s));
    if (n == 0) return CAMA_OK;
    const unsigned grid = (unsigned)((n + 255) / 256);
    if (format == CAMA_OVERLAY_PALETTE) {
        palette_pack_kernel<<<1, 256, 0, s>>>(palette_bgr, static_cast<uint32_t *>(palette_scratch));
        CAMA_LAUNCHED(ctx);
        overlay_expand_kernel<CAMA_OVERLAY_PALETTE><<<grid, 256, 0, s>>>(static_cast<const uint32_t *>(records), n, static_cast<const uint32_t *>(palette_scratch), frames, n_chunks);
    } else {
        overlay_expand_kernel<CAMA_OVERLAY_BGR><<<grid, 256, 0, s>>>(static_cast<const uint32_t *>(records), n, nullptr, frames, n_chunks);
    }
    CAMA_LAUNCHED(ctx);
    return CAMA_OK;
}

extern "C" int cama_overlay_apply_host(const void *records, int64_t n, int format, const uint8_t *palette_bgr,
                                       const cama_overlay_target *target, int op, int n_threads) {
    CAMA_REQUIRE(n >= 0, "negative size");
    CAMA_REQUIRE(op >= CAMA_OVERLAY_DRAW && op <= CAMA_OVERLAY_BLANK_CHUNKS, "bad op");
    CAMA_REQUIRE(format == CAMA_OVERLAY_BGR || format == CAMA_OVERLAY_PALETTE, "bad format");
    if (n == 0) return CAMA_OK;
    CAMA_REQUIRE(records && target && target->pixels, "NULL buffer");
    CAMA_REQUIRE(format != CAMA_OVERLAY_PALETTE || palette_bgr, "palette_bgr is NULL");
    CAMA_REQUIRE(target->n_frames >= 0 && target->n_cams > 0 && target->height > 0 && target->width > 0 && target->width % 8 == 0 &&
                     target->grid_cols >= 0, "bad target shape");
    Target t{};
    t.pixels = target->pixels;
    t.n_cams = target->n_cams; t.height = target->height; t.width = target->width; t.grid_cols = target->grid_cols;
    t.chunks_per_row = target->width / 8;
    t.chunks_per_image = t.chunks_per_row * target->height;
    t.n_chunks = target->n_frames * target->n_cams * t.chunks_per_image;
    t.tile_of_cam = target->tile_of_cam;
    if (t.grid_cols > 0) {
        CAMA_REQUIRE(t.tile_of_cam, "tile_of_cam is NULL");
        const int grid_rows = (t.n_cams + t.grid_cols - 1) / t.grid_cols;
        for (int c = 0; c < t.n_cams; ++c)
            CAMA_REQUIRE(t.tile_of_cam[c] >= 0 && t.tile_of_cam[c] < grid_rows * t.grid_cols, "tile_of_cam[%d] out of range", c);
        t.mosaic_pitch = (size_t)t.grid_cols * t.width * 3;
        t.mosaic_frame = t.mosaic_pitch * grid_rows * t.height;
    }
    int threads = n_threads > 0 ? n_threads : (int)std::max(1u, std::thread::hardware_concurrency());
    if (n < 4096) threads = 1;
    if (format == CAMA_OVERLAY_BGR) {
        BgrJob job{static_cast<const cama_overlay_record *>(records), n, &t, op};
        HostPool::instance().run(threads, n, kHostChunk, apply_bgr_range, &job);
    } else {
        const cama_overlay_record_palette *rec = static_cast<const cama_overlay_record_palette *>(records);
        uint32_t pal32[256];                                  // packed B | G << 8 | R << 16, entry 0 (not painted) = black
        pal32[0] = 0;
        int used = 1;                                          // entries above the last non-black one are never referenced in practice
        for (int e = 1; e < 256; ++e) {
            pal32[e] = (uint32_t)palette_bgr[3 * e] | ((uint32_t)palette_bgr[3 * e + 1] << 8) | ((uint32_t)palette_bgr[3 * e + 2] << 16);
            if (pal32[e]) used = e + 1;
        }
#ifdef CAMA_HAVE_SSSE3
        if ((op == CAMA_OVERLAY_DRAW_CHUNKS || op == CAMA_OVERLAY_DRAW) && used <= 16 && __builtin_cpu_supports("ssse3")) {
            if (op == CAMA_OVERLAY_DRAW) draw_pal16<true>(rec, n, palette_bgr, t, threads);
            else draw_pal16<false>(rec, n, palette_bgr, t, threads);
            return CAMA_OK;
        }
#endif
        // two pixels per lookup: pair[a | b << 8] = the 6 bytes of pixel a followed by pixel b
        static thread_local uint64_t *pair = nullptr;
        static thread_local int pair_used = 0;                 // entries >= pair_used (either half) are zero = black
        if (!pair) pair = new uint64_t[65536]();
        const int fill = used > pair_used ? used : pair_used;  // also overwrites what an earlier, larger palette left behind
        if (op != CAMA_OVERLAY_BLANK_CHUNKS) {
            for (int b2 = 0; b2 < 256; ++b2) {                 // every pair with a non-black half (entries >= used are black)
                const int a_end = b2 < fill ? 256 : fill;
                for (int a2 = 0; a2 < a_end; ++a2) pair[a2 | b2 << 8] = (uint64_t)pal32[a2] | (uint64_t)pal32[b2] << 24;
            }
            pair_used = used;
        }
        PalJob job{rec, n, &t, op, pair};
        HostPool::instance().run(threads, n, kHostChunk, apply_pal_range, &job);
    }
    return CAMA_OK;
}

// STREAM-like probe of what the host's memory system gives the worker pool: a parallel fill and a parallel copy over
// buffers far larger than the caches.  bench.py quotes it next to the end-to-end figure, which is bound by exactly this
// (every lit chunk of the host frames is a read-for-ownership and a write-back of a cache line, twice per call).
namespace {
struct ProbeJob {
    unsigned char *a, *b;
    int op;
};
void probe_range(void *ctx, int64_t lo, int64_t hi) {
    const ProbeJob &j = *static_cast<const ProbeJob *>(ctx);
    if (j.op == 0) memset(j.a + lo, 1, (size_t)(hi - lo));
    else memcpy(j.b + lo, j.a + lo, (size_t)(hi - lo));
}
}  // namespace

extern "C" int cama_host_bandwidth_probe(int64_t bytes, int n_threads, double *fill_gbs, double *copy_gbs) {
    CAMA_REQUIRE(bytes >= (1 << 20) && fill_gbs && copy_gbs, "bad argument");
    int threads = n_threads > 0 ? n_threads : (int)std::max(1u, std::thread::hardware_concurrency());
    unsigned char *a = static_cast<unsigned char *>(malloc((size_t)bytes)), *b = static_cast<unsigned char *>(malloc((size_t)bytes));
    if (!a || !b) {
        free(a); free(b);
        return fail(CAMA_E_INVALID, "cama_host_bandwidth_probe: out of host memory");
    }
    ProbeJob job{a, b, 0};
    const int64_t chunk = 1 << 20;
    auto seconds = [&](int op) {
        job.op = op;
        double best = 1e30;
        for (int rep = 0; rep < 4; ++rep) {                    // (the first pass also faults the pages in)
            const auto t0 = std::chrono::steady_clock::now();
            HostPool::instance().run(threads, bytes, chunk, probe_range, &job);
            best = std::min(best, std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count());
        }
        return best;
    };
    *fill_gbs = (double)bytes / seconds(0) / 1e9;               // bytes written
    seconds(1);
    *copy_gbs = 2.0 * (double)bytes / seconds(1) / 1e9;         // bytes read + bytes written
    free(a); free(b);
    return CAMA_OK;
}

// Device records -> host frames in one call: the records are copied to the caller's pinned staging buffer in a few
// slices (six by default) on `stream`, and each slice is applied to the target while the next one is still crossing
// PCIe (records are independent: unique chunks, any order).  The same pipeline Reproject.__call__ ran from Python,
// without the per-slice interpreter and tensor-dispatch overhead.
extern "C" int cama_overlay_fetch_apply(cama_ctx *ctx, const void *records_dev, int64_t n, int format, const uint8_t *palette_bgr,
                                        void *staging_pinned, const cama_overlay_target *target, int op, int n_threads, void *stream) {
    CAMA_REQUIRE(ctx, "ctx is NULL");
    CAMA_REQUIRE(n >= 0, "negative size");
    CAMA_REQUIRE(format == CAMA_OVERLAY_BGR || format == CAMA_OVERLAY_PALETTE, "bad format");
    if (n == 0) return CAMA_OK;
    CAMA_REQUIRE(records_dev && staging_pinned, "NULL buffer");
    const size_t rb = format == CAMA_OVERLAY_BGR ? sizeof(cama_overlay_record) : sizeof(cama_overlay_record_palette);
    // slices: a small first one (its draw starts early), then growing ones; the call ends one slice's draw after the
    // last byte has crossed PCIe, so the last slice should be short too (three slices 1/8, 1/4, 5/8 left the draw of
    // 5/8 of the records exposed)
    constexpr int kMaxSlices = 8;
    static const int want_slices = getenv("CAMA_FETCH_SLICES") ? std::min(kMaxSlices, std::max(1, atoi(getenv("CAMA_FETCH_SLICES")))) : 6;
    int64_t cuts[kMaxSlices + 1];
    int n_slices = 1;
    cuts[0] = 0;
    cuts[1] = n;
    if (n >= (1 << 16) && want_slices > 1) {
        n_slices = want_slices;
        // weights 1, 2, 3, 3, 3, ... normalised
        double w[kMaxSlices], total = 0.0;
        for (int k = 0; k < n_slices; ++k) { w[k] = k < 2 ? k + 1.0 : 3.0; total += w[k]; }
        double acc = 0.0;
        for (int k = 0; k < n_slices; ++k) { acc += w[k]; cuts[k + 1] = k + 1 == n_slices ? n : (int64_t)((double)n * acc / total); }
    }
    DeviceGuard guard(ctx->device);
    cudaStream_t s = (cudaStream_t)stream;
    while (ctx->fetch_events.size() < (size_t)kMaxSlices) {
        cudaEvent_t e;
        CAMA_CUDA_TRY(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
        ctx->fetch_events.push_back(e);
    }
    const unsigned char *src = static_cast<const unsigned char *>(records_dev);
    unsigned char *dst = static_cast<unsigned char *>(staging_pinned);
    for (int k = 0; k < n_slices; ++k) {
        const int64_t count = cuts[k + 1] - cuts[k];
        if (count > 0) CAMA_CUDA_TRY(cudaMemcpyAsync(dst + (size_t)cuts[k] * rb, src + (size_t)cuts[k] * rb, (size_t)count * rb, cudaMemcpyDeviceToHost, s));
        CAMA_CUDA_TRY(cudaEventRecord(ctx->fetch_events[k], s));
    }
    for (int k = 0; k < n_slices; ++k) {
        CAMA_CUDA_TRY(cudaEventSynchronize(ctx->fetch_events[k]));
        const int64_t count = cuts[k + 1] - cuts[k];
        if (count > 0) {
            const int rc = cama_overlay_apply_host(dst + (size_t)cuts[k] * rb, count, format, palette_bgr, target, op, n_threads);
            if (rc != CAMA_OK) return rc;
        }
    }
    return CAMA_OK;
}
