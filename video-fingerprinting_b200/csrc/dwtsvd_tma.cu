// TMA-staged, persistent variants of the Haar-DWT / block-SVD embed and extract kernels (sm_100a).
//
// Same arithmetic and outputs as dwtsvd.cu (the per-tile functions are shared through
// dwtsvd_tile.cuh); what changes is how the frame strips travel:
//
//   * A work item is one STRIP: the 8 sample rows of (a column chunk of) one tile row of one frame.
//     With tight rows (pitch == width) and at most 256 tiles per row a strip is 8*width contiguous
//     bytes (15 KB at 1080p), so it moves with ONE bulk-copy TMA instruction (cp.async.bulk, SASS
//     UBLKCP) instead of one vector load per thread and row.  Wider planes (4K: 480 tiles per row)
//     are cut into column chunks of at most 256 tiles and pitched planes keep their row gaps out
//     of shared memory: both move with one bulk copy per sample row (1920 bytes at 4K), issued by
//     eight lanes of the producer warp.  (A first version used 2-D tensor-map boxes of 256x8 bytes
//     per warp; ncu showed the copy engine saturating on those small boxes at 2.3 TB/s -
//     profiles/r01_summary.md - which is why the copies are now whole rows or whole strips.)
//   * Each CTA is persistent and owns a ring of kStages strip buffers in shared memory.  It is warp
//     specialised: four consumer warps do the arithmetic, a fifth PRODUCER warp only moves data.
//     The producer issues the bulk loads for the strips the CTA will need next and arms a "full"
//     mbarrier with the byte count; consumers wait on its phase, and each consumer warp arrives
//     on the slot's "done" mbarrier when it is finished, so consumer warps never wait for each other
//     (no CTA-wide barrier in the loop).  HBM latency is hidden by the ring depth, no consumer
//     spends issue slots on global address arithmetic and the 64 source bytes of a tile never
//     occupy registers for the length of the SVD.
//   * Narrow planes (at most 128 tiles per row, e.g. the 960-wide chroma planes of 1080p yuv420p) take
//     two tile rows per work item, so that both packed lanes of every thread stay busy.
//   * Consumer thread t owns TWO tiles of the strip, t and t + half (half = half the tiles of the strip; narrow
//     planes: tile t of the first and of the second tile row), and runs the eigen-iteration for both in packed FP32
//     (FFMA2/FMUL2/FADD2, svd4x2.cuh): the kernels were issue-bound, and a packed instruction does the
//     work of two for one issue slot.  It reads its 8x8 bytes from shared memory (conflict-free: a warp
//     reads 256 contiguous bytes per row).  Embed updates the strip in
//     shared memory and writes it back with a bulk store (cp.async.bulk ... bulk_group); a slot
//     is refilled one iteration after its store was committed
//     (cp.async.bulk.wait_group.read 1), so loads, math and stores of neighbouring strips
//     overlap.
//   * Strips are aligned to tile rows, not to the 32-bit words of the packed bit arrays, so the
//     extract kernel ORs each warp's ballot into (at most two) words of a zeroed raw-bit array
//     and the embed kernel funnel-shifts its 32 watermark bits out of two words.
//
// Requirements (otherwise the launcher uses the vectorised-load kernels): planar uint8, at least 64 tiles
// per row, base and frame stride multiples of 16 bytes, and either tight rows of at most 256 tiles (any
// parity) or 16-byte aligned rows with an even number of tiles; rows of 129..183 tiles are left to the
// vectorised-load kernels, which measured faster there.
#include <cstdlib>
#include <mutex>
#include "common.cuh"
#include "svd4.cuh"
#include "dwtsvd_tile.cuh"

namespace b200wm {

constexpr int kStripThreads = 128;                 // consumer threads: two 8x8 tiles each (t and t + half)
constexpr int kMaxStripTiles = 2 * kStripThreads;
#ifndef B200WM_EMBED_STAGES
#define B200WM_EMBED_STAGES 3
#endif
constexpr int kEmbedStages = B200WM_EMBED_STAGES;      // >= 3: load in flight + strip being updated + store in flight
#ifndef B200WM_EXTRACT_STAGES
#define B200WM_EXTRACT_STAGES 3
#endif
#ifndef B200WM_EXTRACT_MIN_CTAS
#define B200WM_EXTRACT_MIN_CTAS 4
#endif
#ifndef B200WM_EMBED_MIN_CTAS
#define B200WM_EMBED_MIN_CTAS 3
#endif
#ifndef B200WM_EMBED_WIDE_CTAS
#define B200WM_EMBED_WIDE_CTAS 2
#endif
constexpr int kExtractStages = B200WM_EXTRACT_STAGES;

// ---- PTX wrappers ---------------------------------------------------------------------------------
__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(unsigned bar, unsigned count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(unsigned bar, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(unsigned bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
// Wait for the phase with the given parity.  The suspend-time hint lets the hardware park the warp
// until the barrier completes (it is woken by the arrival) instead of spinning through issue slots.
__device__ __forceinline__ void mbar_wait(unsigned bar, unsigned parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "B200WM_WAIT:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1, %2;\n"
        "@p bra B200WM_DONE;\n"
        "bra B200WM_WAIT;\n"
        "B200WM_DONE:\n"
        "}\n" ::"r"(bar), "r"(parity), "r"(1000000u) : "memory");
}
// global -> shared bulk copy, completion signalled on an mbarrier (bytes % 16 == 0, 16-byte aligned)
__device__ __forceinline__ void bulk_load(unsigned dst, const void* src, unsigned bytes, unsigned bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}
// shared -> global bulk copy, tracked by the bulk async-group of the issuing thread
__device__ __forceinline__ void bulk_store(void* dst, unsigned src, unsigned bytes) {
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst), "r"(src), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
template <int kPending>
__device__ __forceinline__ void bulk_wait_read() {
    asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(kPending) : "memory");
}
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ uint2 lds_u2(unsigned addr) {
    uint2 r;
    asm volatile("ld.shared.v2.u32 {%0, %1}, [%2];" : "=r"(r.x), "=r"(r.y) : "r"(addr));
    return r;
}
// predicated store: no branch (and no reconvergence bookkeeping) around the 16 row stores of a tile
__device__ __forceinline__ void sts_u2_if(bool pred, unsigned addr, uint2 v) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "setp.ne.u32 p, %3, 0;\n"
        "@p st.shared.v2.u32 [%0], {%1, %2};\n"
        "}\n" ::"r"(addr), "r"(v.x), "r"(v.y), "r"((unsigned)pred) : "memory");
}

// ---- work items ---------------------------------------------------------------------------------------
// Items are numbered i = (frame * tiles_y + ty) * chunks_x + cx; CTA b processes i = b, b + grid, ...
struct StripGeom {
    TileGeom g;
    int total;                       // n_frames * frame_items (< 2^26 so that the magic divisions are exact)
    int chunks_x;                    // column chunks per tile row (1 up to 256 tiles per row)
    int chunk_tiles;                 // tiles per chunk, even; the last chunk of a row may hold fewer
    int frame_items;                 // work items per frame: tiles_y * chunks_x, or ceil(tiles_y / 2) on narrow planes
    unsigned long long frame_magic;  // i / frame_items == (i * magic) >> 40
    unsigned long long chunk_magic;  // j / chunks_x    == (j * magic) >> 40
    unsigned pitch;                  // row pitch of the plane in global memory
    unsigned slot_pitch;             // row pitch of a strip in shared memory = 8 * chunk_tiles
    unsigned slot_bytes;             // 8 (narrow planes: 16) * slot_pitch
    int whole;                       // 1: the strip is contiguous in global memory too -> one bulk copy
    int narrow;                      // 1: two tile rows per item
    long long frame_stride;
};

struct Item {
    int frame, ty, cx, tiles, rows;  // tiles = tiles of this chunk; rows = tile rows of this item (narrow planes: 1 or 2)
};

// kWhole: one chunk per tile row and tight rows (the 1080p case) - a strip is contiguous in global
// memory, moves with a single bulk copy, and the item index needs one division only.
// kNarrow: tiles_x <= 128 - an item is two consecutive tile rows (one at the bottom of an odd plane).
template <bool kWhole, bool kNarrow>
__device__ __forceinline__ Item item_of(int i, const StripGeom& sg) {
    Item it;
    it.frame = (int)(((unsigned long long)(unsigned)i * sg.frame_magic) >> 40);
    const int j = i - it.frame * sg.frame_items;
    if (kNarrow) {
        it.ty = 2 * j;
        it.cx = 0;
        it.tiles = sg.g.tiles_x;
        it.rows = min(2, sg.g.tiles_y - it.ty);
    } else if (kWhole) {
        it.ty = j;
        it.cx = 0;
        it.tiles = sg.g.tiles_x;
        it.rows = 1;
    } else {
        it.ty = (int)(((unsigned long long)(unsigned)j * sg.chunk_magic) >> 40);
        it.cx = j - it.ty * sg.chunks_x;
        it.tiles = min(sg.chunk_tiles, sg.g.tiles_x - it.cx * sg.chunk_tiles);
        it.rows = 1;
    }
    return it;
}

__device__ __forceinline__ long long item_offset(const Item& it, const StripGeom& sg) {
    return it.frame * sg.frame_stride + (long long)(it.ty * 8) * sg.pitch + (long long)it.cx * sg.slot_pitch;
}

// full[s]: armed by the producer with the byte count of a load into slot s, completed by the copy engine.
// done[s]: one arrival per consumer warp when it has finished with slot s.
template <int kStages, int kCW>
__device__ __forceinline__ void init_ring(unsigned full0, unsigned done0) {
    if (threadIdx.x == 0) {
#pragma unroll
        for (int s = 0; s < kStages; ++s) {
            mbar_init(full0 + 8 * s, 1);
            mbar_init(done0 + 8 * s, kCW);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
}

// Called by the whole (converged) producer warp: lane 0 arms the barrier, then it moves the whole
// strip (kWhole) or 8 (16) lanes move one sample row each.
template <bool kWhole, bool kNarrow>
__device__ __forceinline__ void load_item(const uint8_t* src, int i, unsigned slot, unsigned bar, const StripGeom& sg, int lane) {
    const Item it = item_of<kWhole, kNarrow>(i, sg);
    const unsigned row_bytes = (unsigned)it.tiles * 8u, n_rows = 8u * (unsigned)it.rows;
    if (kWhole) {
        if (lane == 0) {
            mbar_arrive_expect_tx(bar, n_rows * row_bytes);
            bulk_load(slot, src + item_offset(it, sg), n_rows * row_bytes, bar);
        }
        return;
    }
    if (lane == 0) mbar_arrive_expect_tx(bar, n_rows * row_bytes);
    __syncwarp();
    if ((unsigned)lane < n_rows)
        bulk_load(slot + lane * sg.slot_pitch, src + item_offset(it, sg) + (unsigned long long)lane * sg.pitch, row_bytes, bar);
}

template <bool kWhole, bool kNarrow>
__device__ __forceinline__ void store_item(uint8_t* dst, int i, unsigned slot, const StripGeom& sg, int lane) {
    const Item it = item_of<kWhole, kNarrow>(i, sg);
    const unsigned row_bytes = (unsigned)it.tiles * 8u, n_rows = 8u * (unsigned)it.rows;
    if (kWhole) {
        if (lane == 0) bulk_store(dst + item_offset(it, sg), slot, n_rows * row_bytes);
    } else if ((unsigned)lane < n_rows) {
        bulk_store(dst + item_offset(it, sg) + (unsigned long long)lane * sg.pitch, slot + lane * sg.slot_pitch, row_bytes);
    }
    bulk_commit();
}

// The two tiles of consumer thread t in an item: shared-memory byte offsets (dead lanes read tile 0 and
// are masked later), liveness, and the distance between their block indices.
template <bool kNarrow>
struct TilePair {
    bool live_lo, live_hi, warp_live_lo, warp_live_hi;
    unsigned off_lo, off_hi, hi_delta;
    __device__ __forceinline__ TilePair(int t, const Item& it, const StripGeom& sg, unsigned slot_pitch) {
        live_lo = t < it.tiles;
        warp_live_lo = (t & ~31) < it.tiles;
        if (kNarrow) {
            live_hi = live_lo && it.rows == 2;
            warp_live_hi = warp_live_lo && it.rows == 2;
            hi_delta = (unsigned)sg.g.tiles_x;
            off_hi = live_hi ? 8u * slot_pitch + (unsigned)t * 8u : 0u;
        } else {
            // the strip's tiles are split in two equal halves, so that planes with 129..255 tiles per row (720p: 160,
            // portrait 1080p: 135) keep both packed lanes busy in fewer warps instead of leaving the upper lane idle
            const int half = (it.tiles + 1) >> 1;
            live_lo = t < half;
            warp_live_lo = (t & ~31) < half;
            live_hi = live_lo && t + half < it.tiles;
            warp_live_hi = warp_live_lo && (t & ~31) + half < it.tiles;
            hi_delta = (unsigned)half;
            off_hi = live_hi ? (unsigned)(t + half) * 8u : 0u;
        }
        off_lo = live_lo ? (unsigned)t * 8u : 0u;
    }
};

// ---- extract -------------------------------------------------------------------------------------------
// raw_bits must be zero on entry (the launcher clears it): warps OR their bits in.
// kPitch: the shared-memory row pitch when it is known at compile time (1920 for 1080p strips and 4K chunks,
// 960 for the chroma planes of 1080p yuv420p), 0 otherwise: row offsets then fold into the LDS / STS immediates.
// kCW: consumer warps per CTA.  4 (two tiles per thread cover strips of up to 256 tiles, or two tile rows of up to 128)
// for every width but strips of 170..192 tiles, whose half-strips of 85..96 tiles would leave a quarter of four warps
// idle: those run with 3 consumer warps and one more CTA per SM (0.91-0.94 of peak against 0.87 for the vectorised-load
// kernels, scripts/midwidth_probe.py).
template <bool kWhole, bool kNarrow, unsigned kPitch, int kCW>
__global__ void __launch_bounds__((kCW + 1) * 32, kCW == 3 ? B200WM_EXTRACT_MIN_CTAS + 1 : (kCW == 8 ? 2 : B200WM_EXTRACT_MIN_CTAS))
dwtsvd_extract_tma_kernel(const uint8_t* __restrict__ src, ExtractArgs ex, StripGeom sg) {
    constexpr int kStages = kExtractStages;
    constexpr int kConsumerWarps = kCW;
    extern __shared__ __align__(128) uint8_t smem[];
    __shared__ __align__(8) unsigned long long bars[2 * kStages];
    __shared__ int cta_counts[kStages][32];          // per-strip vote counts, one buffer per ring slot
    const unsigned ring = smem_u32(smem);
    const unsigned full0 = smem_u32(&bars[0]), done0 = smem_u32(&bars[kStages]);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const TileGeom& g = sg.g;
    if (threadIdx.x < kStages * 32) cta_counts[threadIdx.x >> 5][threadIdx.x & 31] = 0;
    init_ring<kStages, kCW>(full0, done0);
    const int step = (int)gridDim.x;
    const int L = ex.payload_len;
    int stage = 0;
    unsigned parity = 0;

    if (warp == kConsumerWarps) {
        // ===== producer warp: loads, and the per-strip flush of the vote counters =====
#pragma unroll
        for (int s = 0; s < kStages; ++s)
            if ((int)blockIdx.x + s * step < sg.total)
                load_item<kWhole, kNarrow>(src, (int)blockIdx.x + s * step, ring + s * sg.slot_bytes, full0 + 8 * s, sg, lane);
        for (int i = (int)blockIdx.x; i < sg.total; i += step) {
            mbar_wait(done0 + 8 * stage, parity);        // every consumer warp is done with this slot
            if (ex.pos_counts && lane < L) {
                const int n = cta_counts[stage][lane];
                if (n) {
                    atomicAdd(&ex.pos_counts[(long long)item_of<kWhole, kNarrow>(i, sg).frame * L + lane], n);
                    cta_counts[stage][lane] = 0;
                }
            }
            __syncwarp();
            if (i + kStages * step < sg.total)
                load_item<kWhole, kNarrow>(src, i + kStages * step, ring + stage * sg.slot_bytes, full0 + 8 * stage, sg, lane);
            if (++stage == kStages) { stage = 0; parity ^= 1u; }
        }
        return;
    }

    // ===== consumer warps =====
    const int t = threadIdx.x;                                   // the launcher guarantees chunk_tiles <= 2 * kStripThreads
    for (int i = (int)blockIdx.x; i < sg.total; i += step) {
        const Item it = item_of<kWhole, kNarrow>(i, sg);
        const unsigned slot = ring + stage * sg.slot_bytes;
        const unsigned sp = kPitch ? kPitch : sg.slot_pitch;
        const TilePair<kNarrow> tp(t, it, sg, sp);
        mbar_wait(full0 + 8 * stage, parity);
        unsigned bits;
        {
            f2 S[16];
            {
                uint2 ra[8], rb[8];
#pragma unroll
                for (int r = 0; r < 8; ++r) {
                    ra[r] = lds_u2(slot + tp.off_lo + r * sp);
                    rb[r] = lds_u2(slot + tp.off_hi + r * sp);
                }
                sums_from_rows_x2(ra, rb, S);
            }
            bits = extract_bits_x2(S, ex.scale, ex.inv_scale,
                                   [&](int which) { return flat_probe_shared(slot + (which ? tp.off_hi : tp.off_lo), sp); });
        }
        const unsigned ballot_lo = __ballot_sync(0xFFFFFFFFu, tp.live_lo && (bits & 1u));
        const unsigned ballot_hi = __ballot_sync(0xFFFFFFFFu, tp.live_hi && (bits & 2u));
        const unsigned c0 = (unsigned)(it.ty * g.tiles_x + it.cx * sg.chunk_tiles + (t & ~31));     // first block of the low half
        const unsigned c1 = c0 + tp.hi_delta;                                                       // ... of the high half
        if (lane == 0) {
            uint32_t* frame_bits = ex.raw_bits + (long long)it.frame * g.words;
            if (tp.warp_live_lo && ballot_lo) {
                const unsigned sh = c0 & 31u;
                atomicOr(frame_bits + (c0 >> 5), ballot_lo << sh);
                if (sh && (ballot_lo >> (32u - sh))) atomicOr(frame_bits + (c0 >> 5) + 1, ballot_lo >> (32u - sh));
            }
            if (tp.warp_live_hi && ballot_hi) {
                const unsigned sh = c1 & 31u;
                atomicOr(frame_bits + (c1 >> 5), ballot_hi << sh);
                if (sh && (ballot_hi >> (32u - sh))) atomicOr(frame_bits + (c1 >> 5) + 1, ballot_hi >> (32u - sh));
            }
        }
        if (ex.pos_counts && lane < L) {
            // lane i sees the bits of blocks c+i, c+i+L, ...: payload position (c + i) mod L.  When the distance between
            // the halves is a multiple of L (1080p: 120 blocks) both land on the same position.
            if (kNarrow || (tp.hi_delta & (unsigned)(L - 1))) {
                const int n0 = __popc(ballot_lo & (ex.every << lane)), n1 = __popc(ballot_hi & (ex.every << lane));
                if (n0) atomicAdd(&cta_counts[stage][(c0 + lane) & (unsigned)(L - 1)], n0);
                if (n1) atomicAdd(&cta_counts[stage][(c1 + lane) & (unsigned)(L - 1)], n1);
            } else {
                const int n = __popc(ballot_lo & (ex.every << lane)) + __popc(ballot_hi & (ex.every << lane));
                if (n) atomicAdd(&cta_counts[stage][(c0 + lane) & (unsigned)(L - 1)], n);
            }
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(done0 + 8 * stage);
        if (++stage == kStages) { stage = 0; parity ^= 1u; }
    }
}

// ---- embed ----------------------------------------------------------------------------------------------
template <bool kWhole, bool kNarrow, unsigned kPitch, int kCW>
__global__ void __launch_bounds__((kCW + 1) * 32, kCW == 3 ? B200WM_EMBED_MIN_CTAS + 1 : (kCW == 8 ? B200WM_EMBED_WIDE_CTAS : B200WM_EMBED_MIN_CTAS))
dwtsvd_embed_tma_kernel(const uint8_t* __restrict__ src, uint8_t* __restrict__ dst, EmbedArgs em, StripGeom sg) {
    constexpr int kStages = kEmbedStages;
    constexpr int kConsumerWarps = kCW;
    static_assert(kStages >= 3, "a slot is refilled one iteration after its store was committed");
    extern __shared__ __align__(128) uint8_t smem[];
    __shared__ __align__(8) unsigned long long bars[2 * kStages];
    const unsigned ring = smem_u32(smem);
    const unsigned full0 = smem_u32(&bars[0]), done0 = smem_u32(&bars[kStages]);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const TileGeom& g = sg.g;
    init_ring<kStages, kCW>(full0, done0);
    const int step = (int)gridDim.x;
    int stage = 0;
    unsigned parity = 0;

    if (warp == kConsumerWarps) {
        // ===== producer warp: loads, stores, refills (a single lane when a strip is one copy) =====
        if (kWhole && lane != 0) return;
#pragma unroll
        for (int s = 0; s < kStages - 1; ++s)            // the last slot is filled by the first refill
            if ((int)blockIdx.x + s * step < sg.total)
                load_item<kWhole, kNarrow>(src, (int)blockIdx.x + s * step, ring + s * sg.slot_bytes, full0 + 8 * s, sg, lane);
        int refill = kStages - 1;
        for (int i = (int)blockIdx.x; i < sg.total; i += step) {
            mbar_wait(done0 + 8 * stage, parity);        // every consumer warp has rewritten its tiles of this strip
            store_item<kWhole, kNarrow>(dst, i, ring + stage * sg.slot_bytes, sg, lane);
            // the slot stored one iteration ago has been read by now (every lane: at most its newest store pending)
            bulk_wait_read<1>();
            if (!kWhole) __syncwarp();
            const int nxt = i + (kStages - 1) * step;    // the strip kStages-1 iterations ahead
            if (nxt < sg.total) load_item<kWhole, kNarrow>(src, nxt, ring + refill * sg.slot_bytes, full0 + 8 * refill, sg, lane);
            refill = stage;
            if (++stage == kStages) { stage = 0; parity ^= 1u; }
        }
        bulk_wait_read<0>();
        return;
    }

    // ===== consumer warps =====
    const int t = threadIdx.x;                                   // the launcher guarantees chunk_tiles <= 2 * kStripThreads
    // The watermark bits of a thread's two tiles come from global memory through a dependent chain (the frame's row
    // number, then the word of that row holding the bit).  Both levels are fetched ahead - the row number two items
    // ahead, the words one item ahead - so that each load has a whole item's arithmetic to land in: with the chain
    // inside the item a batch with per-frame rows ran 6 % slower than one with a single row.
    auto row_of = [&](int i) -> int {
        if (!em.frame_row || i >= sg.total) return 0;
        const int frame = (int)(((unsigned long long)(unsigned)i * sg.frame_magic) >> 40);
        return em.frame_row[frame];
    };
    // -> the two words holding the bits of this thread's low / high tile in item i, and the bit positions inside them
    auto words_of = [&](int i, int row, unsigned& w_lo, unsigned& w_hi) {
        w_lo = w_hi = 0u;
        if (i >= sg.total) return;
        const Item it = item_of<kWhole, kNarrow>(i, sg);
        const TilePair<kNarrow> tp(t, it, sg, 0u);
        const uint32_t* wrow = em.wm + (long long)clamp_row(row, em.n_rows) * em.wm_words;
        const unsigned c = (unsigned)(it.ty * g.tiles_x + it.cx * sg.chunk_tiles + t);
        const int wi_lo = (int)(c >> 5), wi_hi = (int)((c + tp.hi_delta) >> 5);
        if (tp.live_lo && wi_lo < em.wm_words) w_lo = wrow[wi_lo];
        if (tp.live_hi && wi_hi < em.wm_words) w_hi = wrow[wi_hi];
    };
    unsigned next_lo, next_hi;
    words_of((int)blockIdx.x, row_of((int)blockIdx.x), next_lo, next_hi);
    int row_ahead = row_of((int)blockIdx.x + step);
    for (int i = (int)blockIdx.x; i < sg.total; i += step) {
        const Item it = item_of<kWhole, kNarrow>(i, sg);
        const unsigned slot = ring + stage * sg.slot_bytes;
        const unsigned sp = kPitch ? kPitch : sg.slot_pitch;
        const TilePair<kNarrow> tp(t, it, sg, sp);
        const unsigned c = (unsigned)(it.ty * g.tiles_x + it.cx * sg.chunk_tiles + t);
        const unsigned mybits = ((next_lo >> (c & 31u)) & 1u) | (((next_hi >> ((c + tp.hi_delta) & 31u)) & 1u) << 1);
        words_of(i + step, row_ahead, next_lo, next_hi);          // issued here, used in the next iteration
        row_ahead = row_of(i + 2 * step);
        mbar_wait(full0 + 8 * stage, parity);
        {
            const unsigned mine_lo = slot + tp.off_lo, mine_hi = slot + tp.off_hi;
            f2 D[16];
            {
                f2 S[16];
                {
                    uint2 ra[8], rb[8];
#pragma unroll
                    for (int r = 0; r < 8; ++r) {
                        ra[r] = lds_u2(mine_lo + r * sp);
                        rb[r] = lds_u2(mine_hi + r * sp);
                    }
                    sums_from_rows_x2(ra, rb, S);
                }
                embed_deltas_x2(S, mybits, em.scale, em.inv_scale, 12582912.0f, D,
                                [&](int which) { return flat_probe_shared(which ? mine_hi : mine_lo, sp); });
            }
#pragma unroll
            for (int i2 = 0; i2 < 4; ++i2) {
                const unsigned a01 = __byte_perm(__float_as_uint(D[4 * i2 + 0].x), __float_as_uint(D[4 * i2 + 1].x), 0x5410);
                const unsigned a23 = __byte_perm(__float_as_uint(D[4 * i2 + 2].x), __float_as_uint(D[4 * i2 + 3].x), 0x5410);
                const unsigned b01 = __byte_perm(__float_as_uint(D[4 * i2 + 0].y), __float_as_uint(D[4 * i2 + 1].y), 0x5410);
                const unsigned b23 = __byte_perm(__float_as_uint(D[4 * i2 + 2].y), __float_as_uint(D[4 * i2 + 3].y), 0x5410);
#pragma unroll
                for (int rr = 0; rr < 2; ++rr) {
                    const unsigned ro = (2 * i2 + rr) * sp;
                    sts_u2_if(tp.live_lo, mine_lo + ro, add_clamp_row(lds_u2(mine_lo + ro), a01, a23));
                    sts_u2_if(tp.live_hi, mine_hi + ro, add_clamp_row(lds_u2(mine_hi + ro), b01, b23));
                }
            }
        }
        fence_async_smem();          // generic-proxy writes -> visible to the producer's bulk store
        __syncwarp();
        if (lane == 0) mbar_arrive(done0 + 8 * stage);
        if (++stage == kStages) { stage = 0; parity ^= 1u; }
    }
}

// ---- host side ----------------------------------------------------------------------------------------------
// three strips of at most 2 KB x 8 rows leave room for four (extract) / three (embed) CTAs per SM
static int strip_chunks(const TileGeom& g, int* chunk_tiles, int max_tiles = kMaxStripTiles) {
    const int chunks = (g.tiles_x + max_tiles - 1) / max_tiles;
    int per = (g.tiles_x + chunks - 1) / chunks;
    if (chunks > 1) per += per & 1;                   // even: 16-byte aligned chunk starts
    *chunk_tiles = per;
    return (g.tiles_x + per - 1) / per;
}

// Two ways in.  A strip that is contiguous in global memory (tight rows covered exactly by tiles, at most 256
// tiles per row) moves as one bulk copy of 8 * pitch bytes, which is a multiple of 16 whatever the width: this
// also takes planes whose rows are only 8-byte aligned or hold an odd number of tiles.  Everything else moves row by row and needs 16-byte aligned rows and chunk starts.
bool tma_eligible(const void* a, const void* b, const b200wm_plane* pl, const TileGeom& g) {
    if (!(pl->dtype == B200WM_U8 && pl->elem_stride == 1 && ((uintptr_t)a % 16) == 0 && ((uintptr_t)b % 16) == 0 &&
          (pl->frame_stride_bytes % 16) == 0 && g.tiles_x >= 64 && g.tiles_y > 0))
        return false;
    if (pl->n_frames > 1 && pl->frame_stride_bytes < pl->pitch_bytes * (long long)pl->height) return false;
    // Strips of 129..169 tiles (720p: 160, portrait 1080p: 135) and of 193..215 tiles carry too few tiles per consumer
    // thread whatever the CTA shape (measured, embed + extract, fraction of the HBM peak: 720p 0.86 with three consumer
    // warps and one tile row per item, 0.88 with five and two rows, against 0.90 for the vectorised-load kernels;
    // 135 tiles 0.71 / 0.74 against 0.86; 208 tiles 0.83 against 0.86 - scripts/midwidth_probe.py), so those planes
    // stay on the vectorised-load kernels.
    if ((g.tiles_x > kStripThreads && g.tiles_x < 170) || (g.tiles_x > 192 && g.tiles_x < 216)) return false;
    const bool whole = g.tiles_x <= kMaxStripTiles && pl->pitch_bytes == 8ll * g.tiles_x;
    if (!whole && ((pl->pitch_bytes % 16) != 0 || (g.tiles_x % 2) != 0)) return false;
    int chunk_tiles = 0;
    const long long frame_items = (long long)g.tiles_y * strip_chunks(g, &chunk_tiles);     // (narrow planes: about half of it)
    const long long items = pl->n_frames * frame_items;
    return items < (1ll << 26) && items * frame_items < (1ll << 40);     // exact magic divisions
}

// Grid of a persistent kernel: SMs x resident CTAs.  Asked from the runtime once per (kernel, device, shared-memory
// size) and remembered.  (All instantiations of one kernel template share a function-pointer TYPE, hence this
// function and its tables: entries are keyed by the kernel's address.)
template <typename Kernel>
static int persistent_grid(Kernel kernel, int cta_threads, size_t smem, int* blocks) {
    struct Entry { const void* fn; int dev; size_t smem; int blocks; };
    constexpr int kMax = 64;
    static std::mutex mu;
    static Entry cache[kMax];        // smem == ~0: "dynamic shared-memory limit raised for (fn, dev)"
    static int used = 0;
    const void* fn = reinterpret_cast<const void*>(kernel);
    int dev = 0;
    B200WM_CUDA_TRY(cudaGetDevice(&dev));
    std::lock_guard<std::mutex> lock(mu);
    bool opted_in = false;
    for (int i = 0; i < used; ++i) {
        if (cache[i].fn != fn || cache[i].dev != dev) continue;
        if (cache[i].smem == smem) { *blocks = cache[i].blocks; return B200WM_OK; }
        if (cache[i].smem == ~(size_t)0) opted_in = true;
    }
    if (!opted_in) {
        // once per kernel and device: allow everything the SM offers, so that launches with another strip size need no call
        int optin = 0;
        cudaFuncAttributes fa;
        B200WM_CUDA_TRY(cudaDeviceGetAttribute(&optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev));
        B200WM_CUDA_TRY(cudaFuncGetAttributes(&fa, kernel));
        B200WM_CUDA_TRY(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, optin - (int)fa.sharedSizeBytes));
        if (used < kMax) cache[used++] = Entry{fn, dev, ~(size_t)0, 0};
    }
    int sms = 0, per_sm = 0;
    B200WM_CUDA_TRY(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    B200WM_CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, cta_threads, smem));
    if (per_sm < 1) per_sm = 1;
    *blocks = sms * per_sm;
    if (used < kMax) cache[used++] = Entry{fn, dev, smem, *blocks};
    return B200WM_OK;
}

static StripGeom make_strip_geom(const TileGeom& g, const b200wm_plane* pl, int max_tiles = kMaxStripTiles) {
    StripGeom sg;
    sg.g = g;
    sg.chunks_x = strip_chunks(g, &sg.chunk_tiles, max_tiles);
    sg.narrow = g.tiles_x <= kStripThreads ? 1 : 0;
    sg.frame_items = sg.narrow ? (g.tiles_y + 1) / 2 : g.tiles_y * sg.chunks_x;
    sg.total = pl->n_frames * sg.frame_items;
    sg.frame_magic = (1ull << 40) / (unsigned long long)sg.frame_items + 1ull;
    sg.chunk_magic = (1ull << 40) / (unsigned long long)sg.chunks_x + 1ull;
    sg.pitch = (unsigned)pl->pitch_bytes;
    sg.slot_pitch = 8u * (unsigned)sg.chunk_tiles;
    sg.slot_bytes = (sg.narrow ? 16u : 8u) * sg.slot_pitch;
    sg.whole = (sg.chunks_x == 1 && sg.slot_pitch == sg.pitch) ? 1 : 0;
    sg.frame_stride = pl->frame_stride_bytes;
    return sg;
}

static bool mid_width(const StripGeom& sg) { return !sg.narrow && sg.chunks_x == 1 && sg.g.tiles_x <= 192; }

template <bool kWhole, bool kNarrow, unsigned kPitch, int kCW = 4>
static int launch_extract_t(const uint8_t* src, const ExtractArgs& xa, const StripGeom& sg, cudaStream_t stream) {
    const size_t smem = (size_t)kExtractStages * sg.slot_bytes;
    int blocks = 0;
    const int rc = persistent_grid(dwtsvd_extract_tma_kernel<kWhole, kNarrow, kPitch, kCW>, (kCW + 1) * 32, smem, &blocks);
    if (rc) return rc;
    if (sg.total < blocks) blocks = sg.total;
    dwtsvd_extract_tma_kernel<kWhole, kNarrow, kPitch, kCW><<<blocks, (kCW + 1) * 32, smem, stream>>>(src, xa, sg);
    B200WM_LAUNCH_CHECK("dwtsvd_extract_tma_kernel");
    return B200WM_OK;
}

template <bool kWhole, bool kNarrow, unsigned kPitch, int kCW = 4>
static int launch_embed_t(const uint8_t* src, uint8_t* dst, const EmbedArgs& ea, const StripGeom& sg, cudaStream_t stream) {
    const size_t smem = (size_t)kEmbedStages * sg.slot_bytes;
    int blocks = 0;
    const int rc = persistent_grid(dwtsvd_embed_tma_kernel<kWhole, kNarrow, kPitch, kCW>, (kCW + 1) * 32, smem, &blocks);
    if (rc) return rc;
    if (sg.total < blocks) blocks = sg.total;
    dwtsvd_embed_tma_kernel<kWhole, kNarrow, kPitch, kCW><<<blocks, (kCW + 1) * 32, smem, stream>>>(src, dst, ea, sg);
    B200WM_LAUNCH_CHECK("dwtsvd_embed_tma_kernel");
    return B200WM_OK;
}

int launch_dwtsvd_extract_tma(const void* src, const b200wm_plane* pl, const TileGeom& g, ExtractArgs xa, cudaStream_t stream) {
    const uint8_t* p = (const uint8_t*)src;
    // Wide planes with tight rows (4K: 480 tiles per row): the whole 8-row strip is contiguous, so the extract kernel takes
    // it as ONE 30 KB bulk copy with eight consumer warps (two CTAs per SM) instead of two column chunks moved row by row
    // (sixteen 1.9 KB copies): 0.86 -> 0.94 of the HBM peak at 4K, 0.965 with the row pitch as a compile-time constant.
    if (g.tiles_x > kMaxStripTiles && g.tiles_x <= 2 * kMaxStripTiles && pl->pitch_bytes == 8ll * g.tiles_x) {
        const StripGeom wide = make_strip_geom(g, pl, 2 * kMaxStripTiles);
        if (wide.whole && wide.total < (1 << 26))
            return wide.slot_pitch == 3840 ? launch_extract_t<true, false, 3840, 8>(p, xa, wide, stream) : launch_extract_t<true, false, 0, 8>(p, xa, wide, stream);
    }
    const StripGeom sg = make_strip_geom(g, pl);
    if (sg.narrow) {
        if (sg.whole) return sg.slot_pitch == 960 ? launch_extract_t<true, true, 960>(p, xa, sg, stream) : launch_extract_t<true, true, 0>(p, xa, sg, stream);
        return launch_extract_t<false, true, 0>(p, xa, sg, stream);
    }
    if (mid_width(sg)) return sg.whole ? launch_extract_t<true, false, 0, 3>(p, xa, sg, stream) : launch_extract_t<false, false, 0, 3>(p, xa, sg, stream);
    if (sg.whole) return sg.slot_pitch == 1920 ? launch_extract_t<true, false, 1920>(p, xa, sg, stream) : launch_extract_t<true, false, 0>(p, xa, sg, stream);
    return sg.slot_pitch == 1920 ? launch_extract_t<false, false, 1920>(p, xa, sg, stream) : launch_extract_t<false, false, 0>(p, xa, sg, stream);
}

int launch_dwtsvd_embed_tma(const void* src, void* dst, const b200wm_plane* pl, const TileGeom& g, EmbedArgs ea,
                            cudaStream_t stream) {
    const uint8_t* p = (const uint8_t*)src;
    uint8_t* d = (uint8_t*)dst;
    // The same strip shape as the 4K extract: one 30 KB bulk copy in, one out, eight consumer warps, two CTAs per SM.  Two
    // CTAs of nine warps leave 96 registers per thread (a handful of spilled words) and still beat the column chunks,
    // whose strips move as sixteen 1.9 KB row copies: 0.91 -> 0.96 of the HBM peak at 4K (scripts/probe_plane.py).  The
    // same shape for two tile rows of a 1080p plane was measured too and lost to the four-warp CTAs (0.96 against 0.99).
    // B200WM_NO_EMBED_WIDE=1 in the environment brings the chunks back (A/B evidence).
    static const bool wide_embed = getenv("B200WM_NO_EMBED_WIDE") == nullptr;
    if (wide_embed && g.tiles_x > kMaxStripTiles && g.tiles_x <= 2 * kMaxStripTiles && pl->pitch_bytes == 8ll * g.tiles_x) {
        const StripGeom wide = make_strip_geom(g, pl, 2 * kMaxStripTiles);
        if (wide.whole && wide.total < (1 << 26))
            return wide.slot_pitch == 3840 ? launch_embed_t<true, false, 3840, 8>(p, d, ea, wide, stream) : launch_embed_t<true, false, 0, 8>(p, d, ea, wide, stream);
    }
    const StripGeom sg = make_strip_geom(g, pl);
    if (sg.narrow) {
        if (sg.whole) return sg.slot_pitch == 960 ? launch_embed_t<true, true, 960>(p, d, ea, sg, stream) : launch_embed_t<true, true, 0>(p, d, ea, sg, stream);
        return launch_embed_t<false, true, 0>(p, d, ea, sg, stream);
    }
    if (mid_width(sg)) return sg.whole ? launch_embed_t<true, false, 0, 3>(p, d, ea, sg, stream) : launch_embed_t<false, false, 0, 3>(p, d, ea, sg, stream);
    if (sg.whole) return sg.slot_pitch == 1920 ? launch_embed_t<true, false, 1920>(p, d, ea, sg, stream) : launch_embed_t<true, false, 0>(p, d, ea, sg, stream);
    return sg.slot_pitch == 1920 ? launch_embed_t<false, false, 1920>(p, d, ea, sg, stream) : launch_embed_t<false, false, 0>(p, d, ea, sg, stream);
}

}  // namespace b200wm
