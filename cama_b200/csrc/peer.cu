// libcama_b200: a frame-sharded clip assembled on every GPU of one box without a host in the loop.
//
// BASELINE.json configs[3] shards the frames of a site over the GPUs (frames are independent units of
// /root/reference/cama/dataset.py:88-106) and wants every rank to end up with every frame.  Moving the dense uint8
// frames is NVLink-bound (every rank receives (world-1)/world of 3 GB); only ~6 % of the pixels are painted, so the
// ranks exchange the LIT 8-PIXEL CHUNKS instead and rebuild the frames locally at HBM speed:
//
//   raster (clip.cu, sparse output)   every flush of lit-chunk records goes to the rank's slot of its own mailbox AND,
//                                      by peer stores over NVLink, to the same slot of every peer's mailbox — the
//                                      exchange happens tile by tile while the raster computes, there is no separate
//                                      collective;
//   cama_peer_publish                  one warp: record count + step number into the slot headers on every GPU
//                                      (release at system scope, after the records);
//   cama_peer_expand                   waits (acquire at system scope, with a timeout) until the slots of all ranks
//                                      carry this step's number, then writes every record's 8 pixels into the
//                                      zero-filled frames.
//
// Mailbox of a rank (device memory allocated here with cudaMalloc, so that cudaIpcGetMemHandle can export it):
//   [P parities][world sources] slots of  cama_peer_slot_bytes(capacity, record_bytes)  =  256-byte header | records,
// laid out by the caller; step s uses parity s % P, so a slot is rewritten P steps later.  With everything of a step on
// one stream (render(s) -> publish(s) -> expand(s) -> render(s+1)) P = 2 is enough: a rank can only publish step s+2
// after its expand of step s+1 has seen every peer's step s+1, which those peers published after finishing their expand
// of step s.  cama_b200/shard.py runs the render of step s+1 beside the fill + expand of step s (it only waits for the
// expand of step s-1) and uses P = 4 by the same argument two steps further.
#include <algorithm>
#include <cstdlib>
#include <cstring>

#include "common.cuh"

using namespace cama;

namespace {

struct SlotHeader {                    // first 256 bytes of a slot
    unsigned count;                    // records appended (may exceed the capacity: the excess was dropped)
    unsigned step;                     // written last, with release semantics
    unsigned pad[62];
};
static_assert(sizeof(SlotHeader) == CAMA_PEER_HEADER_BYTES, "slot header size is part of the ABI");

__device__ __forceinline__ void st_release_sys(unsigned *p, unsigned v) { asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory"); }
__device__ __forceinline__ unsigned ld_acquire_sys(const unsigned *p) {
    unsigned v;
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ unsigned long long global_ns() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}

struct PeerPtrs {
    void *p[CAMA_MAX_PEERS];
};

__global__ void peer_publish_kernel(const unsigned *__restrict__ count, unsigned step, PeerPtrs headers, int n) {
    const int i = threadIdx.x;
    if (i >= n) return;
    SlotHeader *h = static_cast<SlotHeader *>(headers.p[i]);
    h->count = *count;
    __threadfence_system();            // the records (written by the kernels before this one) and the count, then the step
    st_release_sys(&h->step, step);
}

// status[0]: 0 = ok, 1 = timed out waiting for a peer (that peer's records are not expanded), 2 = a slot overflowed its capacity
// Slot after slot, starting with the rank's own (which is there first): wait for the slot's step number, then write its
// records grid-stride — the records of the ranks that are already done are expanded while the others still render.
template <int FORMAT>
__global__ void __launch_bounds__(256) peer_expand_kernel(PeerPtrs slots, int world, int first, unsigned step, long long capacity, const uint32_t *__restrict__ palette,
                                                         uint8_t *__restrict__ frames, long long n_chunks, unsigned long long timeout_ns, int *status) {
    __shared__ long long s_count;
    __shared__ uint32_t s_pal[256];
    constexpr int RW = FORMAT == CAMA_OVERLAY_BGR ? 8 : 3;
    constexpr int kPer = 4;                                // independent records per thread and pass: four load -> lookup -> store chains in flight
    if (FORMAT == CAMA_OVERLAY_PALETTE) s_pal[threadIdx.x] = palette[threadIdx.x];
    const unsigned long long t0 = global_ns();
    for (int k = 0; k < world; ++k) {
        const int r = (first + k) % world;
        const SlotHeader *h = static_cast<const SlotHeader *>(slots.p[r]);
        __syncthreads();                                   // (s_count of the previous slot has been read by everyone; the palette is in place)
        if (threadIdx.x == 0) {
            long long c = -1;
            while (ld_acquire_sys(&h->step) != step) {
                if (global_ns() - t0 > timeout_ns) break;
                __nanosleep(100);
            }
            if (ld_acquire_sys(&h->step) == step) {
                c = h->count;
                if (c > capacity) {
                    if (blockIdx.x == 0) atomicMax(status, 2);
                    c = capacity;
                }
            } else if (blockIdx.x == 0) {
                atomicMax(status, 1);
            }
            s_count = c;
        }
        __syncthreads();
        const long long total = s_count;
        const uint32_t *records = reinterpret_cast<const uint32_t *>(static_cast<const unsigned char *>(slots.p[r]) + CAMA_PEER_HEADER_BYTES);
        for (long long i0 = (long long)blockIdx.x * (256 * kPer) + threadIdx.x; i0 < total; i0 += (long long)gridDim.x * (256 * kPer)) {
            uint32_t chunk[kPer], w[kPer][6];
#pragma unroll
            for (int q = 0; q < kPer; ++q) {
                const long long i = i0 + (long long)q * 256;
                chunk[q] = 0xffffffffu;
                if (i >= total) continue;
                const uint32_t *rec = records + i * RW;
                if (FORMAT == CAMA_OVERLAY_BGR) {
                    const uint4 a = reinterpret_cast<const uint4 *>(rec)[0], b = reinterpret_cast<const uint4 *>(rec)[1];
                    chunk[q] = a.x;
                    w[q][0] = a.z; w[q][1] = a.w; w[q][2] = b.x; w[q][3] = b.y; w[q][4] = b.z; w[q][5] = b.w;
                } else {
                    chunk[q] = rec[0];
                    w[q][0] = rec[1]; w[q][1] = rec[2];
                }
            }
#pragma unroll
            for (int q = 0; q < kPer; ++q) {
                if ((long long)chunk[q] >= n_chunks) continue;                 // (also the lanes past the end)
                if (FORMAT == CAMA_OVERLAY_PALETTE) {
                    const uint32_t lo = w[q][0], hi = w[q][1];
                    uint32_t c[8];
#pragma unroll
                    for (int e = 0; e < 4; ++e) {
                        c[e] = s_pal[(lo >> (8 * e)) & 0xffu];
                        c[4 + e] = s_pal[(hi >> (8 * e)) & 0xffu];
                    }
                    w[q][0] = __byte_perm(c[0], c[1], 0x4210); w[q][1] = __byte_perm(c[1], c[2], 0x5421); w[q][2] = __byte_perm(c[2], c[3], 0x6542);
                    w[q][3] = __byte_perm(c[4], c[5], 0x4210); w[q][4] = __byte_perm(c[5], c[6], 0x5421); w[q][5] = __byte_perm(c[6], c[7], 0x6542);
                }
                uint2 *d = reinterpret_cast<uint2 *>(frames + (size_t)chunk[q] * 24);
                d[0] = make_uint2(w[q][0], w[q][1]); d[1] = make_uint2(w[q][2], w[q][3]); d[2] = make_uint2(w[q][4], w[q][5]);
            }
        }
    }
}

// ---- exchange of the centre-record lists (cama_b200/shard.py::ListExchange) ----------------------------------------
// The geometry kernel of a rank has mirrored the records of its frames into every peer's list array; what the peers still
// need are the list lengths.  One kernel copies the rank's cursor range into the same range of every peer's cursor
// array; the step number follows in peer_publish_step_kernel (release, system scope), after that kernel in stream order.
__global__ void __launch_bounds__(256) peer_cursor_copy_kernel(const unsigned *__restrict__ own, long long first, long long count, PeerPtrs peer_cursors, int n_peers) {
    for (long long i = (long long)blockIdx.x * 256 + threadIdx.x; i < count; i += (long long)gridDim.x * 256) {
        const unsigned v = own[first + i];
        for (int p = 0; p < n_peers; ++p) static_cast<unsigned *>(peer_cursors.p[p])[first + i] = v;
    }
}

// The records themselves: one CTA per list (grid-stride), the filled part of the rank's list read once and stored to the
// same place of every peer's array with 16-byte stores, 4 KB contiguous per CTA pass and peer.  (Mirroring every record
// from the geometry kernel as it is appended — 4-byte stores in runs of a few dozen — reached 76 GB/s over NVLink at
// 8 GPUs and took 0.8 ms; this copy moves the same 55 MB in ~0.1 ms.)  capacity % 4 == 0, arrays 16-byte aligned.
__global__ void __launch_bounds__(256) peer_list_push_kernel(const unsigned *__restrict__ own_records, const unsigned *__restrict__ own_cursor, long long first,
                                                            long long count, unsigned capacity, PeerPtrs peer_records, int n_peers) {
    // one WARP per list (a list holds a few hundred records: a CTA per list left three quarters of its threads idle and
    // the lists of a CTA in a queue); 512 contiguous bytes per warp store and peer
    const int lane = threadIdx.x & 31;
    const long long warps = (long long)gridDim.x * 8;
    for (long long l = first + (long long)blockIdx.x * 8 + (threadIdx.x >> 5); l < first + count; l += warps) {
        const unsigned n = min(own_cursor[l], capacity);
        const unsigned n4 = (n + 3u) >> 2;                                   // (capacity % 4 == 0: the padding stays inside the list)
        const uint4 *src = reinterpret_cast<const uint4 *>(own_records + (size_t)l * capacity);
        for (unsigned i0 = 0; i0 < n4; i0 += 128u) {                          // four 16-byte loads in flight per lane
            uint4 v[4];
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const unsigned i = i0 + 32u * k + lane;
                if (i < n4) v[k] = src[i];
            }
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const unsigned i = i0 + 32u * k + lane;
                if (i < n4)
                    for (int p = 0; p < n_peers; ++p) reinterpret_cast<uint4 *>(static_cast<unsigned *>(peer_records.p[p]) + (size_t)l * capacity)[i] = v[k];
            }
        }
    }
}

__global__ void peer_publish_step_kernel(unsigned step, PeerPtrs headers, int n) {
    const int i = threadIdx.x;
    if (i >= n) return;
    SlotHeader *h = static_cast<SlotHeader *>(headers.p[i]);
    __threadfence_system();
    st_release_sys(&h->step, step);
}

// One CTA: waits until the `world` headers (in this rank's own memory) carry `step`; the kernels after it in the stream
// then see everything the peers wrote before they published.
__global__ void peer_wait_kernel(PeerPtrs headers, int world, unsigned step, unsigned long long timeout_ns, int *status) {
    const int r = threadIdx.x;
    if (r >= world) return;
    const SlotHeader *h = static_cast<const SlotHeader *>(headers.p[r]);
    const unsigned long long t0 = global_ns();
    while (ld_acquire_sys(&h->step) != step) {
        if (global_ns() - t0 > timeout_ns) {
            atomicMax(status, 1);
            break;
        }
        __nanosleep(100);
    }
    __threadfence_system();
}

// Zero-fill with many short-lived CTAs (64 KiB each) and streaming stores.  cudaMemsetAsync is as fast alone, but its
// CTAs stay resident until the fill is done, so the render of the same step — enqueued on a higher-priority stream
// exactly so that it can run beside the fill — only got its CTAs when the fill had finished; with short CTAs the block
// scheduler hands every freed slot to the pending high-priority CTAs first.  evict-first stores keep the fill from
// flushing the vertices and records of the render out of L2.
constexpr int kClearThreads = 256, kClearPerThread = 16;             // 256 threads x 16 x 16 B = 64 KiB per CTA
__global__ void __launch_bounds__(kClearThreads) frames_clear_kernel(uint4 *__restrict__ dst, long long n16) {
    const long long base = (long long)blockIdx.x * (kClearThreads * kClearPerThread) + threadIdx.x;
    const uint4 zero = make_uint4(0u, 0u, 0u, 0u);
#pragma unroll
    for (int k = 0; k < kClearPerThread; ++k) {
        const long long i = base + (long long)k * kClearThreads;
        if (i < n16) __stcs(dst + i, zero);
    }
}

__global__ void palette_pack256_kernel(const uint8_t *__restrict__ palette_bgr, uint32_t *__restrict__ packed) {
    const int e = threadIdx.x;
    packed[e] = e == 0 ? 0u : (uint32_t)palette_bgr[3 * e] | ((uint32_t)palette_bgr[3 * e + 1] << 8) | ((uint32_t)palette_bgr[3 * e + 2] << 16);
}

}  // namespace

extern "C" {

int cama_peer_slot_bytes(int64_t capacity_records, int record_bytes, size_t *bytes) {
    CAMA_REQUIRE(bytes && capacity_records >= 0 && record_bytes > 0, "bad argument");
    *bytes = align_up((size_t)CAMA_PEER_HEADER_BYTES + (size_t)capacity_records * (size_t)record_bytes, 256);
    return CAMA_OK;
}

int cama_peer_alloc(cama_ctx *ctx, size_t bytes, void **dev_ptr, void *ipc_handle) {
    CAMA_REQUIRE(ctx && dev_ptr && ipc_handle && bytes > 0, "bad argument");
    static_assert(sizeof(cudaIpcMemHandle_t) == CAMA_PEER_HANDLE_BYTES, "IPC handle size is part of the ABI");
    DeviceGuard guard(ctx->device);
    void *p = nullptr;
    CAMA_CUDA_TRY(cudaMalloc(&p, bytes));
    cudaError_t e = cudaMemset(p, 0, bytes);
    cudaIpcMemHandle_t h;
    if (e == cudaSuccess) e = cudaIpcGetMemHandle(&h, p);
    if (e != cudaSuccess) {
        cudaFree(p);
        return fail(CAMA_E_CUDA, "cama_peer_alloc: %s", cudaGetErrorString(e));
    }
    memcpy(ipc_handle, &h, sizeof(h));
    *dev_ptr = p;
    return CAMA_OK;
}

int cama_peer_free(cama_ctx *ctx, void *dev_ptr) {
    CAMA_REQUIRE(ctx, "ctx is NULL");
    if (!dev_ptr) return CAMA_OK;
    DeviceGuard guard(ctx->device);
    CAMA_CUDA_TRY(cudaFree(dev_ptr));
    return CAMA_OK;
}

int cama_peer_open(cama_ctx *ctx, int peer_device, const void *ipc_handle, void **dev_ptr) {
    CAMA_REQUIRE(ctx && ipc_handle && dev_ptr, "bad argument");
    DeviceGuard guard(ctx->device);
    int can = 0;
    CAMA_CUDA_TRY(cudaDeviceCanAccessPeer(&can, ctx->device, peer_device));
    if (!can) return fail(CAMA_E_UNSUPPORTED, "device %d cannot access the memory of device %d (no peer-to-peer path)", ctx->device, peer_device);
    cudaError_t e = cudaDeviceEnablePeerAccess(peer_device, 0);
    if (e == cudaErrorPeerAccessAlreadyEnabled) {
        cudaGetLastError();
    } else if (e != cudaSuccess) {
        return fail(CAMA_E_CUDA, "cudaDeviceEnablePeerAccess(%d): %s", peer_device, cudaGetErrorString(e));
    }
    cudaIpcMemHandle_t h;
    memcpy(&h, ipc_handle, sizeof(h));
    void *p = nullptr;
    CAMA_CUDA_TRY(cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess));
    *dev_ptr = p;
    return CAMA_OK;
}

int cama_peer_close(cama_ctx *ctx, void *dev_ptr) {
    CAMA_REQUIRE(ctx, "ctx is NULL");
    if (!dev_ptr) return CAMA_OK;
    DeviceGuard guard(ctx->device);
    CAMA_CUDA_TRY(cudaIpcCloseMemHandle(dev_ptr));
    return CAMA_OK;
}

int cama_peer_publish(cama_ctx *ctx, const uint32_t *overlay_count, uint32_t step, void *const *slot_headers, int n, void *stream) {
    CAMA_REQUIRE(ctx && overlay_count && slot_headers, "NULL argument");
    CAMA_REQUIRE(n > 0 && n <= CAMA_MAX_PEERS, "1..%d slots", CAMA_MAX_PEERS);
    DeviceGuard guard(ctx->device);
    PeerPtrs hp{};
    for (int i = 0; i < n; ++i) {
        CAMA_REQUIRE(slot_headers[i], "slot_headers[%d] is NULL", i);
        hp.p[i] = slot_headers[i];
    }
    peer_publish_kernel<<<1, 32, 0, (cudaStream_t)stream>>>(overlay_count, step, hp, n);
    CAMA_LAUNCHED(ctx);
    return CAMA_OK;
}

int cama_peer_publish_lists(cama_ctx *ctx, const void *own_records, const uint32_t *own_cursor, int64_t capacity, int64_t first, int64_t count,
                            void *const *peer_records, void *const *peer_cursors, int n_peers, uint32_t step, void *const *headers, int n_headers,
                            void *stream) {
    CAMA_REQUIRE(ctx && own_records && own_cursor, "NULL argument");
    CAMA_REQUIRE(capacity > 0 && capacity % 4 == 0 && ((uintptr_t)own_records & 15) == 0, "the list capacity must be a multiple of 4 records and the arrays 16-byte aligned");
    CAMA_REQUIRE(first >= 0 && count >= 0 && n_peers >= 0 && n_peers <= CAMA_MAX_PEERS, "bad argument");
    CAMA_REQUIRE(n_peers == 0 || (peer_records && peer_cursors), "peer arrays are NULL");
    if (count > 0 && n_peers > 0) {
        DeviceGuard guard(ctx->device);
        PeerPtrs pr{};
        for (int i = 0; i < n_peers; ++i) {
            CAMA_REQUIRE(peer_records[i] && ((uintptr_t)peer_records[i] & 15) == 0, "peer_records[%d] is NULL or misaligned", i);
            pr.p[i] = peer_records[i];
        }
        peer_list_push_kernel<<<(unsigned)std::min<long long>((count + 7) / 8, (long long)ctx->sm_count * 8), 256, 0, (cudaStream_t)stream>>>(
            static_cast<const unsigned *>(own_records), own_cursor, first, count, (unsigned)capacity, pr, n_peers);
        CAMA_LAUNCHED(ctx);
    }
    return cama_peer_publish_cursors(ctx, own_cursor, first, count, peer_cursors, n_peers, step, headers, n_headers, stream);
}

int cama_peer_publish_cursors(cama_ctx *ctx, const uint32_t *own_cursor, int64_t first, int64_t count, void *const *peer_cursors, int n_peers,
                              uint32_t step, void *const *headers, int n_headers, void *stream) {
    CAMA_REQUIRE(ctx && own_cursor && headers, "NULL argument");
    CAMA_REQUIRE(first >= 0 && count >= 0 && n_peers >= 0 && n_peers <= CAMA_MAX_PEERS && n_headers > 0 && n_headers <= CAMA_MAX_PEERS, "bad argument");
    CAMA_REQUIRE(n_peers == 0 || peer_cursors, "peer_cursors is NULL");
    DeviceGuard guard(ctx->device);
    cudaStream_t s = (cudaStream_t)stream;
    PeerPtrs pc{}, hp{};
    for (int i = 0; i < n_peers; ++i) {
        CAMA_REQUIRE(peer_cursors[i], "peer_cursors[%d] is NULL", i);
        pc.p[i] = peer_cursors[i];
    }
    for (int i = 0; i < n_headers; ++i) {
        CAMA_REQUIRE(headers[i], "headers[%d] is NULL", i);
        hp.p[i] = headers[i];
    }
    if (count > 0 && n_peers > 0) {
        peer_cursor_copy_kernel<<<(unsigned)std::min<long long>((count + 255) / 256, 1024), 256, 0, s>>>(own_cursor, first, count, pc, n_peers);
        CAMA_LAUNCHED(ctx);
    }
    peer_publish_step_kernel<<<1, 32, 0, s>>>(step, hp, n_headers);
    CAMA_LAUNCHED(ctx);
    return CAMA_OK;
}

int cama_peer_wait(cama_ctx *ctx, void *const *headers, int world, uint32_t step, int timeout_ms, int32_t *status, void *stream) {
    CAMA_REQUIRE(ctx && headers && status, "NULL argument");
    CAMA_REQUIRE(world > 0 && world <= CAMA_MAX_PEERS, "1..%d ranks", CAMA_MAX_PEERS);
    DeviceGuard guard(ctx->device);
    PeerPtrs hp{};
    for (int i = 0; i < world; ++i) {
        CAMA_REQUIRE(headers[i], "headers[%d] is NULL", i);
        hp.p[i] = headers[i];
    }
    const unsigned long long timeout_ns = (unsigned long long)(timeout_ms > 0 ? timeout_ms : 10000) * 1000000ull;
    peer_wait_kernel<<<1, 32, 0, (cudaStream_t)stream>>>(hp, world, step, timeout_ns, status);
    CAMA_LAUNCHED(ctx);
    return CAMA_OK;
}

int cama_frames_clear(cama_ctx *ctx, uint8_t *frames, size_t bytes, void *stream) {
    CAMA_REQUIRE(ctx, "ctx is NULL");
    if (bytes == 0) return CAMA_OK;
    CAMA_REQUIRE(frames, "frames is NULL");
    DeviceGuard guard(ctx->device);
    cudaStream_t s = (cudaStream_t)stream;
    static const bool use_memset = getenv("CAMA_CLEAR_MEMSET") != nullptr;          // experiment knob
    if (use_memset || ((uintptr_t)frames & 15) != 0) {
        CAMA_CUDA_TRY(cudaMemsetAsync(frames, 0, bytes, s));
        return CAMA_OK;
    }
    const long long n16 = (long long)(bytes / 16);
    const long long per_cta = (long long)kClearThreads * kClearPerThread;
    if (n16 > 0) {
        frames_clear_kernel<<<(unsigned)((n16 + per_cta - 1) / per_cta), kClearThreads, 0, s>>>(reinterpret_cast<uint4 *>(frames), n16);
        CAMA_LAUNCHED(ctx);
    }
    if (bytes % 16) CAMA_CUDA_TRY(cudaMemsetAsync(frames + (size_t)n16 * 16, 0, bytes % 16, s));
    return CAMA_OK;
}

int cama_peer_expand(cama_ctx *ctx, void *const *slots, int world, int own_rank, uint32_t step, int64_t capacity_records, int format, const uint8_t *palette_bgr,
                     void *palette_scratch, uint8_t *frames, int64_t n_frames, int n_cams, int height, int width, int timeout_ms, int32_t *status,
                     void *stream) {
    CAMA_REQUIRE(ctx && slots && status, "NULL argument");
    CAMA_REQUIRE(world > 0 && world <= CAMA_MAX_PEERS, "1..%d ranks", CAMA_MAX_PEERS);
    CAMA_REQUIRE(own_rank >= 0 && own_rank < world, "own_rank out of range");
    CAMA_REQUIRE(format == CAMA_OVERLAY_BGR || format == CAMA_OVERLAY_PALETTE, "bad format");
    CAMA_REQUIRE(format != CAMA_OVERLAY_PALETTE || (palette_bgr && palette_scratch), "the palette format needs palette_bgr and palette_scratch (device)");
    CAMA_REQUIRE(n_frames >= 0 && n_cams > 0 && height > 0 && width > 0 && width % 8 == 0 && capacity_records >= 0, "bad shape");
    const int64_t n_chunks = n_frames * n_cams * height * (width / 8);
    CAMA_REQUIRE(n_chunks == 0 || (frames && ((uintptr_t)frames & 7) == 0), "frames must be 8-byte aligned");
    DeviceGuard guard(ctx->device);
    cudaStream_t s = (cudaStream_t)stream;
    PeerPtrs sp{};
    for (int r = 0; r < world; ++r) {
        CAMA_REQUIRE(slots[r] && ((uintptr_t)slots[r] & 255) == 0, "slots[%d] must be a 256-byte aligned device pointer", r);
        sp.p[r] = slots[r];
    }
    const unsigned long long timeout_ns = (unsigned long long)(timeout_ms > 0 ? timeout_ms : 10000) * 1000000ull;
    // a few CTAs per SM: the record loop is grid-stride, and every CTA first waits for the peers' step numbers
    const unsigned grid = (unsigned)ctx->sm_count * 8u;
    if (format == CAMA_OVERLAY_PALETTE) {
        palette_pack256_kernel<<<1, 256, 0, s>>>(palette_bgr, static_cast<uint32_t *>(palette_scratch));
        CAMA_LAUNCHED(ctx);
        peer_expand_kernel<CAMA_OVERLAY_PALETTE><<<grid, 256, 0, s>>>(sp, world, own_rank, step, capacity_records, static_cast<const uint32_t *>(palette_scratch), frames, n_chunks, timeout_ns, status);
    } else {
        peer_expand_kernel<CAMA_OVERLAY_BGR><<<grid, 256, 0, s>>>(sp, world, own_rank, step, capacity_records, nullptr, frames, n_chunks, timeout_ns, status);
    }
    CAMA_LAUNCHED(ctx);
    return CAMA_OK;
}

}  // extern "C"
