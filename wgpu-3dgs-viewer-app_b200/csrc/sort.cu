// sort.cu — K2: onesweep-style LSD radix sort of (u32 key, u32 value) pairs.
//
// Replaces viewer.radix_sorter.sort(encoder, bind_group, radix_sort_indirect_args)
// (reference src/tab/scene.rs:865-869): stable ascending sort of the depth keys (f32 bits as
// u32) with the Gaussian indices as payload; the element count lives on the device (the
// reference sizes an indirect dispatch from it), here `*d_n`.
//
// Design: one histogram kernel (all digit histograms in one read of the keys), then one
// kernel per 8-bit digit.  A digit pass is a single sweep: each 6144-key tile ranks its keys
// with warp-level same-digit peer masks (ballots or MATCH.ANY), publishes its per-digit counts and resolves
// its global offsets by decoupled look-back over epoch-tagged status words (chained scan, no
// separate scan kernel, no second read of the keys), then scatters keys and values through
// shared memory so that global writes are runs of consecutive addresses.  Tiles are handed
// out by an atomic ticket so that every predecessor a tile waits on is owned by a running CTA.
// The same kernels sort the (tile id, splat) entries of the binning stage with 2 passes.
#include <initializer_list>
#include <mutex>

#include "common.cuh"

namespace {

constexpr int kThreads = 256;
constexpr int kWarps = kThreads / 32;
constexpr int kKpt = 24;                       // keys per thread
constexpr int kTile = kThreads * kKpt;         // 6144 keys per tile
constexpr int kRadix = 256;
static_assert(kTile <= 65535, "tile_start holds 16-bit positions");

// ------------------------------------------------------------------ histogram kernel
// hist[pass][digit] += count, for `passes` digits.  Warp-aggregated (match_any) shared-memory
// atomics: depth keys share their top bytes, so naive atomics would serialise on one bin.
__global__ void __launch_bounds__(kThreads) k_sort_hist(const uint32_t* __restrict__ keys, const uint32_t* d_n,
                                                        uint32_t n_max, uint32_t* hist, uint32_t passes) {
    __shared__ uint32_t s_hist[4 * kRadix];
    for (int i = threadIdx.x; i < 4 * kRadix; i += kThreads) s_hist[i] = 0;
    __syncthreads();
    uint32_t n = *d_n;
    if (n > n_max) n = n_max;
    const int lane = threadIdx.x & 31;
    // each warp walks 32-key groups, grid-stride
    const uint32_t warps_total = gridDim.x * kWarps;
    const uint32_t gw = blockIdx.x * kWarps + (threadIdx.x >> 5);
    const uint32_t ngroups = (n + 31) / 32;
    for (uint32_t g = gw; g < ngroups; g += warps_total) {
        uint32_t i = g * 32 + lane;
        bool ok = i < n;
        uint32_t k = ok ? keys[i] : 0u;
        uint32_t act = __ballot_sync(0xffffffffu, ok);
        if (ok) {
            for (uint32_t p = 0; p < passes; p++) {
                uint32_t d = (k >> (8 * p)) & 0xffu;
                uint32_t peers = __match_any_sync(act, d);
                if ((uint32_t)lane == (uint32_t)(__ffs((int)peers) - 1)) atomicAdd(&s_hist[p * kRadix + d], __popc(peers));
            }
        }
    }
    __syncthreads();
    for (int i = threadIdx.x; i < (int)passes * kRadix; i += kThreads) {
        uint32_t c = s_hist[i];
        if (c) atomicAdd(&hist[i], c);
    }
}

// ---------------------------------------------------------------------- digit pass
// One tile at a time per CTA (kNBuf = 1): load, count, publish, rank, scatter into the exchange buffer, resolve the
// look-back, write out.  Until late in round 2 the pass was software-pipelined by one tile with double-buffered exchange
// buffers (kNBuf = 2: tile t + 1 ranked before tile t is resolved), on the assumption that the look-back is what a tile
// waits for.  It is not — with the look-back switched off the pipelined pass took the same time — and the second
// buffer's shared memory is what capped the tile at 12 keys per thread and left nothing on the SM for another stream's
// kernels.  Measured on the bench frame (1 viewer / 2 viewers, frames/s): pipelined 12 keys per thread 1221 / 1322,
// 13: 1234 / 1339, 16: 1256 / 1295; one tile at a time 12: 1228 / 1367, 16: 1281 / 1400, 24 at 3 CTAs per SM (172 KB of
// shared memory): **1290 / 1440**, 24 at 2 CTAs: 1258 / 1413, 32 at 2: 1254 / 1360 — bigger tiles mean fewer barrier
// crossings per key, and what is left of the SM's shared memory decides how much of another viewer's frame overlaps.
#ifndef GS_SORT_NBUF
#define GS_SORT_NBUF 1
#endif
#ifndef GS_SORT_MAXCTAS
#define GS_SORT_MAXCTAS 3
#endif
constexpr int kMaxCtasPerSm = GS_SORT_MAXCTAS;   // resident CTAs per SM the launcher asks for
constexpr int kNBuf = GS_SORT_NBUF;               // exchange buffers: 1 = one tile at a time, 2 = pipelined by one tile
struct PassSmem {
    uint32_t warp_hist[kWarps][kRadix];  // per-warp digit counts -> running per-warp offsets
    uint32_t exch_k[kNBuf][kTile];       // keys / values in tile-sorted order
    uint32_t exch_v[kNBuf][kTile];
    uint16_t tile_start[kNBuf][kRadix];  // first position of each digit inside the sorted tile (< kTile <= 65535)
    int32_t global_off[kRadix];          // global index = global_off[digit] + position in sorted tile
    uint32_t scan_tmp[kWarps];
    uint32_t tile_id;
};

// kVote: how the lanes of a warp find their same-digit peers.  MATCH.ANY costs ADU cycles per DISTINCT
// value among the 32 lanes (measured: ~2 cycles each; a pass over uniformly spread digits is ADU-bound),
// one ballot per digit bit costs ~4 ALU instructions whatever the digits are.  kVote = number of low digit
// bits resolved by ballots, the remaining high bits (<= 2^(8 - kVote) distinct values) by MATCH.ANY; measured
// on 5.9 M spread keys: 0 bits 63 us, 3: 60, 4: 53, 5: 47, 6: 48, 8: 51.  The host picks per pass: kVoteBits
// for spread digits (the low bytes of depth keys, the low byte of tile ids), pure MATCH.ANY for concentrated
// ones (the top bytes of depth keys, the row-band byte of tile ids).
constexpr int kVoteBits = 5;
template <int kVote>
__global__ void __launch_bounds__(kThreads, 3) k_sort_pass(uint32_t* __restrict__ keys_a, uint32_t* __restrict__ vals_a,
                                                           uint32_t* __restrict__ keys_b, uint32_t* __restrict__ vals_b,
                                                           const uint32_t* d_n, uint32_t n_max,
                                                           const uint32_t* __restrict__ hist_all, uint32_t pass,
                                                           uint32_t passes, uint64_t* lookback, uint32_t epoch,
                                                           uint32_t* ticket, uint32_t* result_in_b, uint32_t vals_identity) {
    extern __shared__ __align__(16) uint8_t smem_raw[];
    PassSmem& sm = *reinterpret_cast<PassSmem*>(smem_raw);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    uint32_t n = *d_n;
    if (n > n_max) n = n_max;
    const uint32_t ntiles = (n + kTile - 1) / kTile;
    const uint32_t shift = 8 * pass;

    // A digit whose histogram has a single non-empty bin leaves the order unchanged: the pass is
    // skipped (depth keys share their top byte).  Every CTA derives the same plan from the global
    // histograms: which passes run, hence which buffer holds this pass's input.
    uint32_t executed_before = 0;
    bool skip_me = false;
    for (uint32_t q = 0; q < passes; q++) {
        const int degenerate = __syncthreads_or(n > 0 && hist_all[q * kRadix + tid] == n);
        if (q < pass) executed_before += degenerate ? 0u : 1u;
        if (q == pass) skip_me = degenerate != 0;
    }
    const bool src_b = (executed_before & 1u) != 0;
    if (pass == passes - 1 && blockIdx.x == 0 && tid == 0)
        *result_in_b = ((executed_before + (skip_me ? 0u : 1u)) & 1u);
    if (skip_me || n == 0) return;
    const uint32_t* __restrict__ keys_in = src_b ? keys_b : keys_a;
    const uint32_t* __restrict__ vals_in = src_b ? vals_b : vals_a;
    uint32_t* __restrict__ keys_out = src_b ? keys_a : keys_b;
    uint32_t* __restrict__ vals_out = src_b ? vals_a : vals_b;
    const bool synth_vals = vals_identity && executed_before == 0;  // first executed pass: value = input position
    const uint32_t* __restrict__ hist = hist_all + pass * kRadix;

    // exclusive prefix of the global histogram of this digit (thread d owns digit d)
    uint32_t gbase;
    bool digit_used;   // does any key of the whole input carry this thread's digit?  (an unused digit takes no part in the look-back)
    {
        uint32_t c = hist[tid];
        digit_used = c != 0u;
        uint32_t incl = c;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            uint32_t t = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= o) incl += t;
        }
        if (lane == 31) sm.scan_tmp[warp] = incl;
        __syncthreads();
        uint32_t wbase = 0;
        for (int k = 0; k < warp; k++) wbase += sm.scan_tmp[k];
        gbase = wbase + incl - c;
    }
    uint64_t* lb = lookback + (size_t)tid;

    bool have_prev = false;
    uint32_t p_tile = 0, p_count = 0, p_valid = 0;
    for (uint32_t iter = 0;; iter++) {
        const uint32_t buf = kNBuf == 2 ? (iter & 1u) : 0u;
        __syncthreads();  // previous iteration's ranking / write-out are done
        if (tid == 0) sm.tile_id = atomicAdd(ticket, 1u);
#pragma unroll
        for (int k = 0; k < kRadix / 32; k++) sm.warp_hist[warp][k * 32 + lane] = 0;
        __syncthreads();
        const uint32_t tile = sm.tile_id;
        const bool valid_tile = tile < ntiles;
        uint32_t count = 0, valid = 0;
        if (valid_tile) {
            const uint32_t tile_base = tile * kTile;
            valid = min((uint32_t)kTile, n - tile_base);
            // ---- load keys (warp-striped: slot = warp*32*KPT + k*32 + lane keeps index order inside a warp)
            // (a full tile — every tile but the last — takes straight-line code: with a bounds test per element the compiler
            // builds one branch region per element, and in the write-out below the dependent shared-memory loads of the
            // regions then run one after the other instead of overlapped)
            uint32_t key[kKpt];
            const uint32_t wbase_idx = warp * (32 * kKpt);
            const bool full = valid == (uint32_t)kTile;
            const uint32_t* __restrict__ kp = keys_in + tile_base + wbase_idx + lane;
            if (full) {
#pragma unroll
                for (int k = 0; k < kKpt; k++) key[k] = kp[k * 32];
            } else {
#pragma unroll
                for (int k = 0; k < kKpt; k++) {
                    const uint32_t s = wbase_idx + k * 32 + lane;
                    key[k] = s < valid ? kp[k * 32] : 0xffffffffu;
                }
            }
            // ---- early counts (per-warp histograms), so the aggregates can be published at once
#pragma unroll
            for (int k = 0; k < kKpt; k++) atomicAdd(&sm.warp_hist[warp][(key[k] >> shift) & 0xffu], 1u);
            __syncthreads();
            // per digit (thread d): exclusive scan over warps, tile total
#pragma unroll
            for (int w2 = 0; w2 < kWarps; w2++) {
                const uint32_t t = sm.warp_hist[w2][tid];
                sm.warp_hist[w2][tid] = count;
                count += t;
            }
            // publish this tile's aggregate for digit `tid`
            if (digit_used)
                gs_st_status(&lb[(size_t)tile * kRadix], epoch, (tile == 0 ? GS_LOOKBACK_FLAG_INCL : GS_LOOKBACK_FLAG_AGG) | count);
            // exclusive scan of the tile totals over digits
            uint32_t incl = count;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const uint32_t t = __shfl_up_sync(0xffffffffu, incl, o);
                if (lane >= o) incl += t;
            }
            if (lane == 31) sm.scan_tmp[warp] = incl;
            __syncthreads();
            uint32_t wb = 0;
            for (int k = 0; k < warp; k++) wb += sm.scan_tmp[k];
            sm.tile_start[buf][tid] = wb + incl - count;
            __syncthreads();
            // ---- values (loaded late to keep registers low), then rank + scatter into the exchange buffer
            uint32_t val[kKpt];
            if (synth_vals) {
#pragma unroll
                for (int k = 0; k < kKpt; k++) val[k] = tile_base + wbase_idx + k * 32 + lane;
            } else {
                const uint32_t* __restrict__ vp = vals_in + tile_base + wbase_idx + lane;
                if (full) {
#pragma unroll
                    for (int k = 0; k < kKpt; k++) val[k] = vp[k * 32];
                } else {
#pragma unroll
                    for (int k = 0; k < kKpt; k++) {
                        const uint32_t s = wbase_idx + k * 32 + lane;
                        val[k] = s < valid ? vp[k * 32] : 0u;
                    }
                }
            }
            // Two phases, so that the votes of all keys of a thread overlap (they touch no memory and do not depend on
            // each other) and only the short counter update runs as a dependent chain through shared memory.
            uint32_t peers[kKpt];
#pragma unroll
            for (int k = 0; k < kKpt; k++) {
                if (kVote > 0) {
                    // the low kVote bits by ballots, the remaining high bits (few distinct values) by MATCH.ANY
                    uint32_t bal[kVote > 0 ? kVote : 1];
#pragma unroll
                    for (int b = 0; b < kVote; b++) bal[b] = __ballot_sync(0xffffffffu, (key[k] >> (shift + b)) & 1u);
                    uint32_t pm = kVote < 8 ? __match_any_sync(0xffffffffu, ((key[k] >> shift) & 0xffu) >> kVote) : 0xffffffffu;
#pragma unroll
                    for (int b = 0; b < kVote; b++) pm &= ((key[k] >> (shift + b)) & 1u) ? bal[b] : ~bal[b];
                    peers[k] = pm;
                } else {
                    peers[k] = __match_any_sync(0xffffffffu, (key[k] >> shift) & 0xffu);
                }
            }
            const uint32_t lane_lt = (1u << lane) - 1u;
#pragma unroll
            for (int k = 0; k < kKpt; k++) {
                const uint32_t d = (key[k] >> shift) & 0xffu;
                const uint32_t old = sm.warp_hist[warp][d];        // every peer reads the same running offset ...
                const uint32_t mine = __popc(peers[k] & lane_lt);
                __syncwarp();
                if (mine == 0) sm.warp_hist[warp][d] = old + __popc(peers[k]);   // ... and the first peer advances it
                __syncwarp();
                const uint32_t pos = sm.tile_start[buf][d] + old + mine;
                sm.exch_k[buf][pos] = key[k];
                sm.exch_v[buf][pos] = val[k];
            }
        }
        if (kNBuf == 1) { have_prev = valid_tile; p_tile = tile; p_count = count; p_valid = valid; }   // (no pipeline: this tile, now)
        if (have_prev) {
            // ---- resolve the previous tile's look-back for digit `tid`, kLb status words per round trip
            const uint32_t pbuf = kNBuf == 2 ? (buf ^ 1u) : 0u;
            uint32_t excl = 0;
            if (p_tile > 0 && digit_used) {
                constexpr int kLb = 4;
                int64_t p = (int64_t)p_tile - 1;
                bool done = false;
                while (!done) {
                    uint64_t v[kLb];
#pragma unroll
                    for (int j = 0; j < kLb; j++)
                        v[j] = (p - j >= 0) ? gs_ld_status(&lb[(size_t)(p - j) * kRadix])
                                            : (((uint64_t)epoch << 32) | GS_LOOKBACK_FLAG_INCL);
                    int used = 0;
#pragma unroll
                    for (int j = 0; j < kLb; j++) {
                        if (!done && used == j) {
                            const uint32_t fl = gs_status_flag(v[j], epoch);
                            if (fl != 0u) {
                                excl += (uint32_t)v[j] & GS_LOOKBACK_VALUE_MASK;
                                used = j + 1;
                                done = fl == 2u;
                            }
                        }
                    }
                    p -= used;
                }
                gs_st_status(&lb[(size_t)p_tile * kRadix], epoch, GS_LOOKBACK_FLAG_INCL | (excl + p_count));
            }
            sm.global_off[tid] = (int32_t)(gbase + excl) - (int32_t)sm.tile_start[pbuf][tid];
            __syncthreads();
            // ---- write out the previous tile: consecutive positions of one digit are consecutive addresses
            if (p_valid == (uint32_t)kTile) {
                uint32_t kk[kKpt], vv[kKpt], g[kKpt];
#pragma unroll
                for (int k = 0; k < kKpt; k++) { kk[k] = sm.exch_k[pbuf][k * kThreads + tid]; vv[k] = sm.exch_v[pbuf][k * kThreads + tid]; }
#pragma unroll
                for (int k = 0; k < kKpt; k++) g[k] = (uint32_t)(sm.global_off[(kk[k] >> shift) & 0xffu] + (int32_t)(k * kThreads + tid));
#pragma unroll
                for (int k = 0; k < kKpt; k++) { keys_out[g[k]] = kk[k]; vals_out[g[k]] = vv[k]; }
            } else {
#pragma unroll
                for (int k = 0; k < kKpt; k++) {
                    const uint32_t p = k * kThreads + tid;
                    if (p < p_valid) {
                        const uint32_t kk = sm.exch_k[pbuf][p];
                        const uint32_t g = (uint32_t)(sm.global_off[(kk >> shift) & 0xffu] + (int32_t)p);
                        keys_out[g] = kk;
                        vals_out[g] = sm.exch_v[pbuf][p];
                    }
                }
            }
        }
        if (!valid_tile) break;
        if (kNBuf == 2) {
            have_prev = true;
            p_tile = tile;
            p_count = count;
            p_valid = valid;
        }
    }
}

}  // namespace

size_t gs_sort_lookback_words(uint32_t n_max, uint32_t passes) {
    size_t tiles = ((size_t)n_max + kTile - 1) / kTile;
    if (tiles < 1) tiles = 1;
    return tiles * kRadix * passes;
}

cudaError_t gs_launch_sort(const GsSortArgs& a, int num_sms, cudaStream_t st) {
    if (a.passes < 1 || a.passes > 4 || !a.result_in_b) return cudaErrorInvalidValue;
    size_t tiles = ((size_t)a.n_max + kTile - 1) / kTile;
    if (tiles < 1) tiles = 1;
    if (!a.hist_prefilled) {
        uint32_t grid = (uint32_t)(num_sms * 4);
        uint32_t need = (uint32_t)((a.n_max + kThreads - 1) / kThreads);
        if (grid > need) grid = need < 1 ? 1 : need;
        k_sort_hist<<<grid, kThreads, 0, st>>>(a.keys_a, a.d_n, a.n_max, a.hist, a.passes);
    }
    static std::mutex mu;
    static int bps_dev[64] = {0};   // per device (attributes and occupancy are per device)
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return e;
    if (dev < 0 || dev >= 64) return cudaErrorInvalidDevice;
    int blocks_per_sm;
    {
        std::lock_guard<std::mutex> lock(mu);
        if (bps_dev[dev] == 0) {
            for (auto k : {k_sort_pass<0>, k_sort_pass<kVoteBits>}) {
                e = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(PassSmem));
                if (e != cudaSuccess) return e;
            }
            e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&bps_dev[dev], k_sort_pass<0>, kThreads, sizeof(PassSmem));
            if (e != cudaSuccess) { bps_dev[dev] = 0; return e; }
            if (bps_dev[dev] < 1) bps_dev[dev] = 1;
            if (bps_dev[dev] > kMaxCtasPerSm) bps_dev[dev] = kMaxCtasPerSm;
        }
        blocks_per_sm = bps_dev[dev];
    }
    uint32_t grid = (uint32_t)(blocks_per_sm * num_sms);
    if (grid > tiles) grid = (uint32_t)tiles;
    for (uint32_t p = 0; p < a.passes; p++) {
        auto kern = ((a.vote_mask >> p) & 1u) ? k_sort_pass<kVoteBits> : k_sort_pass<0>;
        kern<<<grid, kThreads, sizeof(PassSmem), st>>>(a.keys_a, a.vals_a, a.keys_b, a.vals_b, a.d_n, a.n_max, a.hist, p, a.passes,
                                               a.lookback + (size_t)p * tiles * kRadix, a.epoch, a.tickets + p,
                                               a.result_in_b, a.vals_identity ? 1u : 0u);
    }
    return cudaGetLastError();
}
