// Vote kernels: per-position counts, the per-frame finish of DeShuffler.degenerate, and the
// device half of the cross-frame pattern vote.
//
//   counts / finish  src/offmark/degenerator/de_shuffler.py:14-22
//   pattern vote     tests/segment_mark_detect_hls.py:144-155 (Counter.most_common)
#include "common.cuh"
#include <limits.h>

namespace b200wm {

// ---- general per-position counts from packed raw bits (any payload_len) ---------------------
// grid = (chunks, frames); each thread walks words with a stride; set bits are added to a
// shared-memory histogram (payload_len <= kSmemPositions) or straight to global memory.
constexpr int kSmemPositions = 8192;

__global__ void __launch_bounds__(256) vote_counts_kernel(const uint32_t* __restrict__ raw_bits, int words,
                                                          long long block_num, int L, int32_t* __restrict__ counts,
                                                          int frame0) {
    extern __shared__ int smem_counts[];
    const int frame = frame0 + blockIdx.y;
    const bool use_smem = L <= kSmemPositions;
    if (use_smem) {
        for (int i = threadIdx.x; i < L; i += blockDim.x) smem_counts[i] = 0;
        __syncthreads();
    }
    int32_t* out = counts + (long long)frame * L;
    const uint32_t* bits = raw_bits + (long long)frame * words;
    for (int w = blockIdx.x * blockDim.x + threadIdx.x; w < words; w += gridDim.x * blockDim.x) {
        uint32_t m = bits[w];
        const long long base = (long long)w * 32;
        while (m) {
            const int b = __ffs(m) - 1;
            m &= m - 1;
            const long long c = base + b;
            if (c < block_num) {
                const int pos = (int)(c % L);
                if (use_smem) atomicAdd(&smem_counts[pos], 1);
                else atomicAdd(&out[pos], 1);
            }
        }
    }
    if (use_smem) {
        __syncthreads();
        for (int i = threadIdx.x; i < L; i += blockDim.x)
            if (smem_counts[i]) atomicAdd(&out[i], smem_counts[i]);
    }
}

int launch_vote_counts(const uint32_t* raw_bits, int n_frames, int words_per_frame, long long block_num,
                       int payload_len, int32_t* pos_counts, cudaStream_t stream) {
    if (!raw_bits || !pos_counts || n_frames < 0 || words_per_frame < 0 || payload_len <= 0 || block_num < 0)
        return B200WM_ERR_INVALID;
    if ((long long)words_per_frame * 32 < block_num) return B200WM_ERR_INVALID;
    if (n_frames == 0) return B200WM_OK;
    B200WM_CUDA_TRY(cudaMemsetAsync(pos_counts, 0, sizeof(int32_t) * (size_t)n_frames * payload_len, stream));
    if (words_per_frame == 0) return B200WM_OK;
    const int threads = 256;
    int chunks = (words_per_frame + threads * 4 - 1) / (threads * 4);
    if (chunks < 1) chunks = 1;
    const size_t smem = payload_len <= kSmemPositions ? sizeof(int) * (size_t)payload_len : 0;
    for (int f0 = 0; f0 < n_frames; f0 += 65535) {
        const dim3 grid(chunks, (unsigned)((n_frames - f0) < 65535 ? (n_frames - f0) : 65535));
        vote_counts_kernel<<<grid, threads, smem, stream>>>(raw_bits, words_per_frame, block_num, payload_len,
                                                            pos_counts, f0);
        B200WM_LAUNCH_CHECK("vote_counts_kernel");
    }
    return B200WM_OK;
}

// ---- per-frame finish -------------------------------------------------------------------------
// One warp per frame.  m_i = count_i / n_i in float64 with n_i = number of block indices
// congruent to i (de_shuffler.py:18: wm_bits[i::L].mean()), scattered through perm
// (payload[payload_idx] = payload.copy(), :19), threshold 0.5*(max+min) (:20), strict > (:21).
// IEEE fp64 division, addition and the exact halving give the same doubles as numpy.
__global__ void __launch_bounds__(128) vote_finish_kernel(const int32_t* __restrict__ counts, int n_frames, int L,
                                                          long long block_num, const int32_t* __restrict__ perm,
                                                          uint8_t* __restrict__ patterns,
                                                          unsigned long long* __restrict__ packed) {
    const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (warp >= n_frames) return;
    const int32_t* cnt = counts + (long long)warp * L;
    // an empty slice (payload_len > block_num) makes numpy's mean NaN; np.max/np.min then return
    // NaN, the threshold is NaN and every comparison is False.
    double vmax = -1.0, vmin = 2.0;
    bool empty = false;
    for (int i = lane; i < L; i += 32) {
        const long long n = i < block_num ? (block_num - i + L - 1) / L : 0;
        if (n > 0) {
            const double m = (double)cnt[i] / (double)n;
            vmax = fmax(vmax, m);
            vmin = fmin(vmin, m);
        } else {
            empty = true;
        }
    }
#pragma unroll
    for (int s = 16; s > 0; s >>= 1) {
        vmax = fmax(vmax, __shfl_xor_sync(0xFFFFFFFFu, vmax, s));
        vmin = fmin(vmin, __shfl_xor_sync(0xFFFFFFFFu, vmin, s));
    }
    empty = __any_sync(0xFFFFFFFFu, empty);
    const double thr = 0.5 * (vmax + vmin);
    unsigned long long word = 0ull;
    for (int i = lane; i < L; i += 32) {
        const long long n = i < block_num ? (block_num - i + L - 1) / L : 0;
        const int bit = (!empty && n > 0 && (double)cnt[i] / (double)n > thr) ? 1 : 0;
        const int j = perm[i];
        if ((unsigned)j >= (unsigned)L) continue;          // not an index into the payload: nothing is written for it
        patterns[(long long)warp * L + j] = (uint8_t)bit;
        if (L <= 64 && bit) word |= 1ull << (L - 1 - j);
    }
    if (packed) {
#pragma unroll
        for (int s = 16; s > 0; s >>= 1) word |= __shfl_xor_sync(0xFFFFFFFFu, word, s);
        if (lane == 0) packed[warp] = word;
    }
}

int launch_vote_finish(const int32_t* pos_counts, int n_frames, int payload_len, long long block_num,
                       const int32_t* perm, uint8_t* patterns, uint64_t* packed, cudaStream_t stream) {
    if (!pos_counts || !perm || !patterns || n_frames < 0 || payload_len <= 0 || block_num < 0) return B200WM_ERR_INVALID;
    if (packed && payload_len > 64) return B200WM_ERR_UNSUPPORTED;
    if (n_frames == 0) return B200WM_OK;
    const int threads = 128;
    const long long blocks = ((long long)n_frames * 32 + threads - 1) / threads;
    vote_finish_kernel<<<(unsigned)blocks, threads, 0, stream>>>(pos_counts, n_frames, payload_len, block_num, perm,
                                                                patterns, (unsigned long long*)packed);
    B200WM_LAUNCH_CHECK("vote_finish_kernel");
    return B200WM_OK;
}

// ---- cross-frame pattern histogram ---------------------------------------------------------------
__global__ void __launch_bounds__(256) pattern_hist_kernel(const unsigned long long* __restrict__ packed,
                                                           const int32_t* __restrict__ frame_segment,
                                                           const int32_t* __restrict__ frame_order, int order_offset,
                                                           int n_frames, int L, int n_segments, int32_t* hist,
                                                           int32_t* first_seen, int32_t* bit_votes, int32_t* seg_frames) {
    const int f = blockIdx.x * blockDim.x + threadIdx.x;
    if (f >= n_frames) return;
    const int seg = frame_segment ? frame_segment[f] : 0;
    if (seg < 0 || seg >= n_segments) return;
    const unsigned long long p = packed[f];
    if (p >> L) return;                       // not a pattern of L bits: ignored like a frame of no segment
    const long long bin = (long long)seg * (1ll << L) + (long long)p;
    atomicAdd(&hist[bin], 1);
    atomicMin(&first_seen[bin], frame_order ? frame_order[f] : order_offset + f);
    atomicAdd(&seg_frames[seg], 1);
    for (int j = 0; j < L; ++j)
        if ((p >> (L - 1 - j)) & 1ull) atomicAdd(&bit_votes[(long long)seg * L + j], 1);
}

// ---- pattern histogram fused with its exchange over NVLink --------------------------------------------------
// Multi-GPU form of the cross-frame vote when whole segments are dealt to ranks (BASELINE config 4): every rank
// keeps the per-segment state of ALL ranks in one buffer, rank-major, [world][block_len] int32 followed by `world`
// arrival flags, allocated as symmetric memory so that every rank holds a mapped pointer to every peer's copy
// (torch.distributed._symmetric_memory: plumbing).  ONE kernel accumulates this rank's frames into its own block
// (the histogram atomics of pattern_hist_kernel) and, in helper CTAs of the same launch that wait for the histogram
// CTAs' tickets, stores the finished block straight into the same slot of every peer's buffer with 128-bit stores over
// NVLink / NVSwitch (four CTAs per peer) and raises this rank's arrival flag on every peer (release, system scope) - what the NCCL all-gather delivered before, without a
// second launch, a collective call or a rendezvous on the host.  The exchange is ONE-SIDED: nothing in the step waits
// for a peer, so a rank that runs a little slower does not hold the others back every step (with the wait inside the
// kernel the 8-GPU step measured 0.14 ms of vote time, all of it skew between the ranks).  Whoever needs the peers'
// blocks - the host reading the result - first runs vote_exchange_wait_kernel, which waits (acquire) until every
// peer's flag shows the epoch; it gives up after two seconds (a peer that never launches must not hang the GPU) and
// reports it in *status.
struct ExchangeArgs {
    int32_t* const* peers;     // device array [world]: base of every rank's state buffer (index `rank` = the local one)
    long long block_len;       // int32 entries per rank block, a multiple of 4
    int world, rank;
    unsigned epoch;            // increases by one per exchange
    unsigned* ticket;          // local scratch, two counters, zero before the first launch
    int* status;               // local: set to 1 when the wait timed out
};

__device__ __forceinline__ unsigned long long global_ns() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}

__device__ __forceinline__ void hist_and_ticket(const unsigned long long* __restrict__ packed, const int32_t* __restrict__ frame_segment,
                                                const int32_t* __restrict__ frame_order, int order_offset, int n_frames, int L,
                                                int n_segments, int32_t* hist, int32_t* first_seen, int32_t* bit_votes,
                                                int32_t* seg_frames, const ExchangeArgs& ex) {
    const int f = blockIdx.x * blockDim.x + threadIdx.x;
    if (f < n_frames) {
        const int seg = frame_segment ? frame_segment[f] : 0;
        const unsigned long long p = packed[f];
        if (seg >= 0 && seg < n_segments && (p >> L) == 0) {
            const long long bin = (long long)seg * (1ll << L) + (long long)p;
            atomicAdd(&hist[bin], 1);
            atomicMin(&first_seen[bin], frame_order ? frame_order[f] : order_offset + f);
            atomicAdd(&seg_frames[seg], 1);
            for (int j = 0; j < L; ++j)
                if ((p >> (L - 1 - j)) & 1ull) atomicAdd(&bit_votes[(long long)seg * L + j], 1);
        }
    }
    // every histogram CTA takes a ticket behind a fence: whoever reads ticket == n_hist (acquire) sees all their atomics
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) atomicAdd(&ex.ticket[0], 1u);
}

// Helper CTAs of the same launch (blockIdx >= n_hist; kPublishSplit per peer): wait until every histogram CTA has
// taken its ticket, then store one slice of the finished block into one peer's buffer; the helper that finishes last
// raises this rank's flag on every peer and re-arms the counters.  All CTAs of the launch are co-resident (a few dozen
// on 148 SMs), so the wait cannot starve the CTAs it waits for; it still gives up after two seconds.
constexpr int kPublishSplit = 4;

__device__ __forceinline__ void publish_slice(const ExchangeArgs& ex, int n_hist, int helper) {
    __shared__ bool go;
    if (threadIdx.x == 0) {
        const unsigned long long t0 = global_ns();
        unsigned seen;
        go = true;
        do {
            asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(seen) : "l"(ex.ticket) : "memory");
            if (seen >= (unsigned)n_hist) break;
            if (global_ns() - t0 > 2000000000ull) { *ex.status = 1; go = false; break; }
            __nanosleep(200);
        } while (true);
    }
    __syncthreads();
    const int peer_slot = helper / kPublishSplit, part = helper % kPublishSplit;
    const int peer = (ex.rank + 1 + peer_slot) % ex.world;              // every rank starts on another peer
    if (go) {
        const int4* mine = reinterpret_cast<const int4*>(ex.peers[ex.rank] + (long long)ex.rank * ex.block_len);
        int4* theirs = reinterpret_cast<int4*>(ex.peers[peer] + (long long)ex.rank * ex.block_len);
        const int n16 = (int)(ex.block_len / 4), per = (n16 + kPublishSplit - 1) / kPublishSplit;
        const int hi = min(n16, (part + 1) * per);
        for (int i = part * per + threadIdx.x; i < hi; i += blockDim.x) theirs[i] = __ldcg(mine + i);
    }
    __threadfence_system();
    __syncthreads();
    __shared__ bool is_last;
    const int n_helpers = (ex.world - 1) * kPublishSplit;
    if (threadIdx.x == 0) is_last = atomicAdd(&ex.ticket[1], 1u) == (unsigned)(n_helpers - 1);
    __syncthreads();
    if (!is_last) return;
    __threadfence_system();
    if (go && (int)threadIdx.x < ex.world && (int)threadIdx.x != ex.rank) {
        const long long flags = (long long)ex.world * ex.block_len;
        unsigned* there = reinterpret_cast<unsigned*>(ex.peers[threadIdx.x] + flags) + ex.rank;
        asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(there), "r"(ex.epoch) : "memory");
    }
    if (threadIdx.x == 0) { ex.ticket[0] = 0u; ex.ticket[1] = 0u; }      // armed for the next launch
}

__global__ void __launch_bounds__(256) pattern_hist_publish_kernel(const unsigned long long* __restrict__ packed,
                                                                   const int32_t* __restrict__ frame_segment,
                                                                   const int32_t* __restrict__ frame_order, int order_offset,
                                                                   int n_frames, int L, int n_segments, int32_t* hist,
                                                                   int32_t* first_seen, int32_t* bit_votes, int32_t* seg_frames,
                                                                   ExchangeArgs ex, int n_hist) {
    if ((int)blockIdx.x >= n_hist) {
        publish_slice(ex, n_hist, (int)blockIdx.x - n_hist);
        return;
    }
    hist_and_ticket(packed, frame_segment, frame_order, order_offset, n_frames, L, n_segments, hist, first_seen, bit_votes, seg_frames, ex);
    if (ex.world == 1 && threadIdx.x == 0 && blockIdx.x == 0) {
        // no helpers to re-arm the ticket: the first CTA waits for the others' tickets (all resident) and clears it
        unsigned seen;
        do { asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(seen) : "l"(ex.ticket) : "memory"); } while (seen < (unsigned)n_hist);
        ex.ticket[0] = 0u;
    }
}

__global__ void __launch_bounds__(256) vote_exchange_wait_kernel(ExchangeArgs ex) {
    const long long flags = (long long)ex.world * ex.block_len;
    if ((int)threadIdx.x < ex.world && (int)threadIdx.x != ex.rank) {
        const unsigned* here = reinterpret_cast<const unsigned*>(ex.peers[ex.rank] + flags) + threadIdx.x;
        const unsigned long long t0 = global_ns();
        unsigned seen;
        do {
            asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(seen) : "l"(here) : "memory");
            if ((int)(seen - ex.epoch) >= 0) break;
            if (global_ns() - t0 > 2000000000ull) { *ex.status = 1; break; }
        } while (true);
    }
}

int launch_vote_exchange_wait(void* const* peers, long long block_len, int world, int rank, unsigned epoch, int* status,
                              cudaStream_t stream) {
    if (!peers || !status || world < 1 || world > 256 || rank < 0 || rank >= world || block_len <= 0) return B200WM_ERR_INVALID;
    ExchangeArgs ex{reinterpret_cast<int32_t* const*>(peers), block_len, world, rank, epoch, nullptr, status};
    vote_exchange_wait_kernel<<<1, 256, 0, stream>>>(ex);
    B200WM_LAUNCH_CHECK("vote_exchange_wait_kernel");
    return B200WM_OK;
}

int launch_pattern_hist_publish(const uint64_t* packed, const int32_t* frame_segment, const int32_t* frame_order,
                                int order_offset, int n_frames, int payload_len, int n_segments, int32_t* hist,
                                int32_t* first_seen, int32_t* bit_votes, int32_t* seg_frames, void* const* peers,
                                long long block_len, int world, int rank, unsigned epoch, unsigned* ticket, int* status,
                                cudaStream_t stream) {
    if ((!packed && n_frames > 0) || !hist || !first_seen || !bit_votes || !seg_frames || n_frames < 0 || n_segments <= 0 || payload_len <= 0)
        return B200WM_ERR_INVALID;
    if (!peers || !ticket || !status || world < 1 || world > 256 || rank < 0 || rank >= world || block_len <= 0 || block_len % 4)
        return B200WM_ERR_INVALID;
    if (payload_len > 16) return B200WM_ERR_UNSUPPORTED;
    const int threads = 256;
    const int n_hist = n_frames > 0 ? (n_frames + threads - 1) / threads : 1;      // an empty shard still takes part in the exchange
    if (n_hist + (world - 1) * kPublishSplit > 128) return B200WM_ERR_UNSUPPORTED;  // helpers spin: every CTA must be resident at once
    ExchangeArgs ex{reinterpret_cast<int32_t* const*>(peers), block_len, world, rank, epoch, ticket, status};
    pattern_hist_publish_kernel<<<n_hist + (world - 1) * kPublishSplit, threads, 0, stream>>>(
        (const unsigned long long*)packed, frame_segment, frame_order, order_offset, n_frames, payload_len, n_segments, hist, first_seen,
        bit_votes, seg_frames, ex, n_hist);
    B200WM_LAUNCH_CHECK("pattern_hist_publish_kernel");
    return B200WM_OK;
}

// One launch that puts a vote state back to "nothing seen": zeros for the counters, INT32_MAX for first_seen.
__global__ void __launch_bounds__(256) vote_state_reset_kernel(int32_t* __restrict__ state, long long n_zero, long long n_total) {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n_total; i += (long long)gridDim.x * blockDim.x)
        state[i] = i < n_zero ? 0 : INT_MAX;
}

int launch_vote_state_reset(int32_t* state, long long n_zero, long long n_total, cudaStream_t stream) {
    if (!state || n_zero < 0 || n_total < n_zero) return B200WM_ERR_INVALID;
    if (n_total == 0) return B200WM_OK;
    long long blocks = (n_total + 256 * 4 - 1) / (256 * 4);
    if (blocks > 1184) blocks = 1184;
    vote_state_reset_kernel<<<(unsigned)blocks, 256, 0, stream>>>(state, n_zero, n_total);
    B200WM_LAUNCH_CHECK("vote_state_reset_kernel");
    return B200WM_OK;
}

int launch_pattern_hist(const uint64_t* packed, const int32_t* frame_segment, const int32_t* frame_order,
                        int order_offset, int n_frames, int payload_len, int n_segments, int32_t* hist,
                        int32_t* first_seen, int32_t* bit_votes, int32_t* seg_frames, cudaStream_t stream) {
    if (!packed || !hist || !first_seen || !bit_votes || !seg_frames || n_frames < 0 || n_segments <= 0 || payload_len <= 0)
        return B200WM_ERR_INVALID;
    if (payload_len > 16) return B200WM_ERR_UNSUPPORTED;
    if (n_frames == 0) return B200WM_OK;
    const int threads = 256;
    pattern_hist_kernel<<<(n_frames + threads - 1) / threads, threads, 0, stream>>>(
        (const unsigned long long*)packed, frame_segment, frame_order, order_offset, n_frames, payload_len, n_segments,
        hist, first_seen, bit_votes, seg_frames);
    B200WM_LAUNCH_CHECK("pattern_hist_kernel");
    return B200WM_OK;
}

}  // namespace b200wm
