// bin.cu — K3a: tile binning for the compositor.
//
// Part of the replacement of renderer.render_with_pass (reference src/tab/scene.rs:2302-2314):
// the reference draws one instanced quad per visible Gaussian in sorted order and lets the
// rasteriser find the covered pixels; here every depth-sorted splat is expanded into one
// (tile id, splat id) entry per 16x16 tile it can actually touch.  Entries are produced IN
// DEPTH ORDER, so a STABLE sort by tile id alone (2 onesweep passes over 16 bits, sort.cu)
// yields per-tile lists that are still front-to-back.  Models are expanded nearest first, each
// appended after the previous one, which reproduces the reference's per-model layering
// (scene.rs:533-558).
//
// Depth order puts the nearest = largest splats into the first ranks, so a rank-chunked
// expansion is badly imbalanced (the first 1024 ranks hold ~40x the mean work).  The
// expansion is therefore done over the CANDIDATE index space in two balanced kernels:
//   k_bin_count : per depth rank, the number of candidate tiles (the tile rectangle of the extent
//                 square) -> exclusive prefix cand_off[rank] (block scan + decoupled look-back),
//                 plus, for every 2048-candidate block, the rank that owns its first candidate.
//   k_bin_emit  : one CTA per block of 2048 consecutive candidates, one candidate per thread slot:
//                 exact footprint test (a tile is kept only if some pixel of tile ∩ extent square
//                 can reach alpha >= 1/255 — skipped tiles cannot change the image), order-
//                 preserving compaction, decoupled look-back over the blocks' kept counts,
//                 coalesced write-out through shared memory, and the digit histograms of the
//                 tile sort accumulated on the way out.
#include "common.cuh"

namespace {

constexpr int kThreads = 256;
constexpr int kIpt = 4;                      // depth ranks per thread in k_bin_count
constexpr int kChunk = kThreads * kIpt;      // 1024 ranks per chunk
constexpr int kCpt = 8;                      // candidates per thread in k_bin_emit
constexpr uint32_t kBlock = kThreads * kCpt; // 2048 candidates per block

using Cand = GsCand;

// can candidate tile (x, y) of the rectangle be touched?  (gs_min_q_rect with the divisions hoisted)
__device__ __forceinline__ bool tile_hit(const Cand& c, uint32_t x, uint32_t y) {
    const float tx = (float)((c.tx0 + x) * GS_TILE), ty = (float)((c.ty0 + y) * GS_TILE);
    const float dx0 = fmaxf(tx, c.fx0) - c.mx, dx1 = fminf(tx + (float)(GS_TILE - 1), c.fx1) - c.mx;
    const float dy0 = fmaxf(ty, c.fy0) - c.my, dy1 = fminf(ty + (float)(GS_TILE - 1), c.fy1) - c.my;
    const bool inx = dx0 <= 0.0f && dx1 >= 0.0f, iny = dy0 <= 0.0f && dy1 >= 0.0f;
    if (inx && iny) return true;
    float best = 3.0e38f;
    if (!inx) {
        const float dx = dx0 > 0.0f ? dx0 : dx1;
        const float dy = fminf(dy1, fmaxf(dy0, c.nbc * dx));
        best = c.a * dx * dx + 2.0f * c.b * dx * dy + c.c * dy * dy;
    }
    if (!iny) {
        const float dy = dy0 > 0.0f ? dy0 : dy1;
        const float dx = fminf(dx1, fmaxf(dx0, c.nba * dy));
        best = fminf(best, c.a * dx * dx + 2.0f * c.b * dx * dy + c.c * dy * dy);
    }
    return best <= c.tau;
}

// ------------------------------------------------------------------ kernel A: candidate counts
__global__ void __launch_bounds__(kThreads) k_bin_count(const uint32_t* __restrict__ sorted_a,
                                                        const uint32_t* __restrict__ sorted_b,
                                                        const uint32_t* sorted_in_b,
                                                        const uint32_t* __restrict__ ncand,
                                                        const uint32_t* d_v, uint32_t v_max, uint64_t* lookback,
                                                        uint32_t epoch, uint32_t* ticket, uint2* __restrict__ cand_off,
                                                        uint32_t* __restrict__ block_rank, uint32_t block_cap,
                                                        uint32_t* cand_total, uint32_t q_lo, uint32_t q_hi) {
    __shared__ uint32_t s_wsum[kThreads / 32];
    __shared__ uint32_t s_chunk, s_base;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const uint32_t* __restrict__ sorted_slot = (sorted_in_b && *sorted_in_b) ? sorted_b : sorted_a;
    uint32_t vis = *d_v;
    if (vis > v_max) vis = v_max;
    // depth slab: ranks [lo, v) of this model, as 16.16 fractions of the visible count
    const uint32_t lo = (uint32_t)(((uint64_t)vis * q_lo) >> 16), v = (uint32_t)(((uint64_t)vis * q_hi) >> 16);
    const uint32_t nchunks = (v - lo + kChunk - 1) / kChunk;
    if (v == lo) {
        if (blockIdx.x == 0 && tid == 0) *cand_total = 0;
        return;
    }
    // Software-pipelined by one chunk: chunk k+1 is counted and its aggregate published BEFORE chunk
    // k's prefix is resolved, so the look-back never waits for chunks drawn at the same moment.
    bool have_prev = false;
    uint32_t p_c = 0, p_total = 0, p_r0 = 0, p_local = 0, p_cnt[kIpt], p_slot[kIpt];
#pragma unroll
    for (int k = 0; k < kIpt; k++) p_cnt[k] = p_slot[k] = 0;
    while (true) {
        __syncthreads();
        if (tid == 0) s_chunk = atomicAdd(ticket, 1u);
        __syncthreads();
        const uint32_t c = s_chunk;
        const bool valid = c < nchunks;
        uint32_t cnt[kIpt], slot[kIpt], sum = 0, chunk_total = 0, local = 0;
        const uint32_t r0 = lo + c * kChunk + tid * kIpt;
#pragma unroll
        for (int k = 0; k < kIpt; k++) cnt[k] = slot[k] = 0;
        if (valid) {
#pragma unroll
            for (int k = 0; k < kIpt; k++) slot[k] = (r0 + k < v) ? (sorted_slot ? sorted_slot[r0 + k] : r0 + k) : 0u;
            // candidate-tile counts were stored per compaction slot by the preprocess kernel (a 4-byte
            // gather from an L2-resident array instead of a 32-byte splat gather from HBM)
#pragma unroll
            for (int k = 0; k < kIpt; k++) {
                cnt[k] = (r0 + k < v) ? __ldg(&ncand[slot[k]]) : 0u;
                sum += cnt[k];
            }
            uint32_t incl = sum;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                uint32_t t = __shfl_up_sync(0xffffffffu, incl, o);
                if (lane >= o) incl += t;
            }
            if (lane == 31) s_wsum[warp] = incl;
            __syncthreads();
            uint32_t woff = 0;
#pragma unroll
            for (int k = 0; k < kThreads / 32; k++) {
                const uint32_t t = s_wsum[k];
                if (k < warp) woff += t;
                chunk_total += t;
            }
            if (tid == 0) gs_lookback_publish(lookback, epoch, c, chunk_total);
            local = woff + incl - sum;
        }
        if (have_prev) {
            if (warp == 0) {
                const uint32_t excl = gs_lookback_resolve(lookback, epoch, p_c, p_total, lane);
                if (lane == 0) {
                    s_base = excl;
                    if (p_c == nchunks - 1) *cand_total = excl + p_total;
                }
            }
            __syncthreads();
            uint32_t o = s_base + p_local;
#pragma unroll
            for (int k = 0; k < kIpt; k++) {
                if (p_r0 + k < v) {
                    cand_off[p_r0 + k] = make_uint2(o, p_slot[k]);
                    if (p_cnt[k]) {
                        // this rank owns the first candidate of every block whose start falls in its run
                        const uint32_t first = (o + kBlock - 1) / kBlock, last = (o + p_cnt[k] - 1) / kBlock;
                        for (uint32_t b = first; b <= last && b < block_cap; b++) block_rank[b] = p_r0 + k;
                    }
                }
                o += p_cnt[k];
            }
        }
        if (!valid) break;
        have_prev = true;
        p_c = c; p_total = chunk_total; p_r0 = r0; p_local = local;
#pragma unroll
        for (int k = 0; k < kIpt; k++) { p_cnt[k] = cnt[k]; p_slot[k] = slot[k]; }
    }
}

// ------------------------------------------------------------------ kernel B: test + emit
__global__ void __launch_bounds__(kThreads) k_bin_emit(const b200gs_splat* __restrict__ splats,
                                                       const uint32_t* d_v, uint32_t v_max, uint32_t splat_base,
                                                       const uint2* __restrict__ cand_off,
                                                       const uint32_t* __restrict__ block_rank, uint32_t block_cap,
                                                       const uint32_t* cand_total_p, uint64_t* lookback, uint32_t epoch,
                                                       uint32_t* ticket, const uint32_t* entry_base_in,
                                                       uint32_t* entry_total_out, uint32_t* overflow,
                                                       uint32_t* __restrict__ tile_keys, uint32_t* __restrict__ tile_vals,
                                                       uint32_t capacity, uint32_t* tile_hist, float W, float H,
                                                       uint32_t tiles_x, uint32_t flat, uint32_t q_lo, uint32_t q_hi,
                                                       const uint8_t* __restrict__ tile_done) {
    __shared__ int32_t s_owner[kBlock];       // rank owning each candidate (after the max-scan)
    __shared__ uint32_t s_keys[2][kBlock];    // kept entries of the block, double-buffered:
    __shared__ uint32_t s_vals[2][kBlock];    // block b+1 is staged before block b is written out
    __shared__ uint32_t s_hist[512];
    __shared__ uint32_t s_cnt[kCpt][kThreads / 32];
    __shared__ int32_t s_wmax[kThreads / 32];
    __shared__ uint32_t s_blk, s_base, s_kept;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    uint32_t vis = *d_v;
    if (vis > v_max) vis = v_max;
    const uint32_t lo = (uint32_t)(((uint64_t)vis * q_lo) >> 16), v = (uint32_t)(((uint64_t)vis * q_hi) >> 16);
    const uint32_t total = v > lo ? *cand_total_p : 0u;
    uint32_t nblocks = (total + kBlock - 1) / kBlock;
    if (nblocks > block_cap - 1) {  // candidate space larger than the scratch: drop the tail, flag it
        nblocks = block_cap - 1;
        if (blockIdx.x == 0 && threadIdx.x == 0) *overflow = 1u;
    }
    const uint32_t ebase = *entry_base_in;
    if (nblocks == 0) {
        if (blockIdx.x == 0 && tid == 0) *entry_total_out = ebase;
        return;
    }
    for (int i = tid; i < 512; i += kThreads) s_hist[i] = 0;

    // Software-pipelined by one block: block b+1 is tested, counted, published and staged BEFORE
    // block b's prefix is resolved and its entries are written out.
    bool have_prev = false;
    uint32_t p_b = 0, p_kept = 0;
    for (uint32_t iter = 0;; iter++) {
        const uint32_t buf = iter & 1u;
        __syncthreads();
        if (tid == 0) s_blk = atomicAdd(ticket, 1u);
        __syncthreads();
        const uint32_t b = s_blk;
        const bool valid = b < nblocks;
        uint32_t kept_total = 0;
        if (valid) {
        const uint32_t cbase = b * kBlock;
        const uint32_t nc = min(kBlock, total - cbase);

        // ---- owners: rank r starts at cand_off[r]; mark the starts, then an inclusive max-scan
        for (uint32_t i = tid; i < kBlock; i += kThreads) s_owner[i] = -1;
        __syncthreads();
        const uint32_t r_lo = block_rank[b];
        const uint32_t r_hi = (b + 1 < nblocks) ? block_rank[b + 1] : v - 1;
        if (tid == 0) s_owner[0] = (int32_t)r_lo;
        for (uint32_t r = r_lo + 1 + tid; r <= r_hi; r += kThreads) {
            const uint32_t o = cand_off[r].x;
            const uint32_t nxt = (r + 1 < v) ? cand_off[r + 1].x : total;
            if (nxt > o && o >= cbase && o < cbase + kBlock) s_owner[o - cbase] = (int32_t)r;
        }
        __syncthreads();
        {
            // thread t scans its 8 consecutive slots, then warps / block combine
            int32_t loc[kCpt], run = -1;
#pragma unroll
            for (int k = 0; k < kCpt; k++) { run = max(run, s_owner[tid * kCpt + k]); loc[k] = run; }
            int32_t inc = run;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const int32_t t = __shfl_up_sync(0xffffffffu, inc, o);
                if (lane >= o) inc = max(inc, t);
            }
            if (lane == 31) s_wmax[warp] = inc;
            __syncthreads();
            int32_t pre = -1;
            for (int k = 0; k < warp; k++) pre = max(pre, s_wmax[k]);
            const int32_t excl = __shfl_up_sync(0xffffffffu, inc, 1);
            pre = max(pre, lane ? excl : -1);
#pragma unroll
            for (int k = 0; k < kCpt; k++) s_owner[tid * kCpt + k] = max(loc[k], pre);
        }
        __syncthreads();

        // ---- one candidate per thread slot (p = k*256 + tid: consecutive lanes share splats)
        uint32_t key[kCpt], val[kCpt];
        bool keep[kCpt];
#pragma unroll
        for (int k = 0; k < kCpt; k++) {
            const uint32_t p = k * kThreads + tid;
            keep[k] = false;
            key[k] = val[k] = 0;
            if (p < nc) {
                const uint32_t r = (uint32_t)s_owner[p];
                const uint2 os = __ldg(&cand_off[r]);  // {first candidate, splat slot} of the owning rank
                const uint32_t slot = os.y;
                const uint4* sp = reinterpret_cast<const uint4*>(splats + slot);
                const uint4 q0 = __ldg(sp), q1 = __ldg(sp + 1);
                Cand cd;
                if (gs_make_rect(q0, q1, W, H, flat != 0, cd)) {
                    cd.nbc = __fdividef(-cd.b, cd.c); cd.nba = __fdividef(-cd.b, cd.a);  // tau carries the slack
                    const uint32_t e = cbase + p - os.x;
                    // e / nx without the integer-division sequence (e < 2^24, nx <= 1024: exact after one fix-up)
                    uint32_t y = (uint32_t)(__fdividef((float)e + 0.5f, (float)cd.nx));
                    if (y * cd.nx > e) y--;
                    else if ((y + 1) * cd.nx <= e) y++;
                    const uint32_t x = e - y * cd.nx;
                    key[k] = (cd.ty0 + y) * tiles_x + cd.tx0 + x;
                    // tiles already finished by a nearer depth slab take no more entries
                    keep[k] = !(tile_done && tile_done[key[k]]) && tile_hit(cd, x, y);
                    val[k] = splat_base + slot;
                }
            }
            const uint32_t bal = __ballot_sync(0xffffffffu, keep[k]);
            if (lane == 0) s_cnt[k][warp] = __popc(bal);
        }
        __syncthreads();
        // ---- exclusive scan of the 64 (k, warp) counts in candidate order; block total
        if (warp == 0) {
            uint32_t a0 = s_cnt[lane >> 3][lane & 7], a1 = s_cnt[(lane >> 3) + 4][lane & 7];
            uint32_t i0 = a0, i1 = a1;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const uint32_t t0 = __shfl_up_sync(0xffffffffu, i0, o), t1 = __shfl_up_sync(0xffffffffu, i1, o);
                if (lane >= o) { i0 += t0; i1 += t1; }
            }
            const uint32_t half = __shfl_sync(0xffffffffu, i0, 31);
            const uint32_t kept = half + __shfl_sync(0xffffffffu, i1, 31);
            s_cnt[lane >> 3][lane & 7] = i0 - a0;
            s_cnt[(lane >> 3) + 4][lane & 7] = half + i1 - a1;
            if (lane == 0) {
                s_kept = kept;
                gs_lookback_publish(lookback, epoch, b, kept);
            }
        }
        __syncthreads();
        kept_total = s_kept;
        // ---- stage the kept entries in candidate order
#pragma unroll
        for (int k = 0; k < kCpt; k++) {
            const uint32_t bal = __ballot_sync(0xffffffffu, keep[k]);
            if (keep[k]) {
                const uint32_t o = s_cnt[k][warp] + __popc(bal & ((1u << lane) - 1u));
                s_keys[buf][o] = key[k];
                s_vals[buf][o] = val[k];
            }
        }
        }  // valid
        if (have_prev) {
            const uint32_t pbuf = buf ^ 1u;
            if (warp == 0) {
                const uint32_t excl = gs_lookback_resolve(lookback, epoch, p_b, p_kept, lane);
                if (lane == 0) {
                    s_base = excl;
                    if (p_b == nblocks - 1) {
                        uint32_t t = ebase + excl + p_kept;
                        if (t > capacity) { *overflow = 1u; t = capacity; }
                        *entry_total_out = t;
                    }
                }
            }
            __syncthreads();
            const uint32_t gbase = ebase + s_base;
            // ---- coalesced write-out + digit histograms for the tile sort
            for (uint32_t i0 = 0; i0 < p_kept; i0 += kThreads) {
                const uint32_t i = i0 + tid;
                const uint32_t g = gbase + i;
                const bool ok = i < p_kept && g < capacity;
                const uint32_t act = __ballot_sync(0xffffffffu, ok);
                if (ok) {
                    const uint32_t kk = s_keys[pbuf][i];
                    tile_keys[g] = kk;
                    tile_vals[g] = s_vals[pbuf][i];
#pragma unroll
                    for (int p = 0; p < 2; p++) {
                        const uint32_t d = (kk >> (8 * p)) & 0xffu;
                        const uint32_t peers = __match_any_sync(act, d);
                        if (lane == __ffs((int)peers) - 1) atomicAdd(&s_hist[p * 256 + d], (uint32_t)__popc(peers));
                    }
                }
            }
        }
        if (!valid) break;
        have_prev = true;
        p_b = b;
        p_kept = kept_total;
    }
    __syncthreads();
    for (int i = tid; i < 512; i += kThreads) {
        const uint32_t cnt = s_hist[i];
        if (cnt) atomicAdd(&tile_hist[i], cnt);
    }
}

// ranges[tile] = first entry, ranges[n_tiles + tile] = one past the last entry (both 0 if none)
__global__ void __launch_bounds__(256) k_tile_ranges(const uint32_t* __restrict__ keys_a, const uint32_t* __restrict__ keys_b,
                                                     const uint32_t* in_b, const uint32_t* d_entries,
                                                     uint32_t capacity, uint32_t* ranges, uint32_t n_tiles,
                                                     unsigned long long* entry_stat) {
    const uint32_t* __restrict__ tile_keys = *in_b ? keys_b : keys_a;
    uint32_t n = *d_entries;
    if (n > capacity) n = capacity;
    if (entry_stat && blockIdx.x == 0 && threadIdx.x == 0) atomicAdd(entry_stat, (unsigned long long)n);
    // 4 consecutive entries per thread (one 128-bit load) + the two neighbours
    for (uint32_t e0 = (blockIdx.x * blockDim.x + threadIdx.x) * 4; e0 < n; e0 += gridDim.x * blockDim.x * 4) {
        uint32_t k[6];
        if (e0 + 4 <= n) {
            const uint4 q = *reinterpret_cast<const uint4*>(tile_keys + e0);
            k[1] = q.x; k[2] = q.y; k[3] = q.z; k[4] = q.w;
        } else {
            for (int j = 0; j < 4; j++) k[1 + j] = e0 + j < n ? tile_keys[e0 + j] : 0xffffffffu;
        }
        k[0] = e0 > 0 ? tile_keys[e0 - 1] : 0xffffffffu;
        k[5] = e0 + 4 < n ? tile_keys[e0 + 4] : 0xffffffffu;
#pragma unroll
        for (int j = 0; j < 4; j++) {
            const uint32_t e = e0 + j, kk = k[1 + j];
            if (e >= n || kk >= n_tiles) continue;
            if (k[j] != kk) ranges[kk] = e;
            if (e == n - 1 || k[j + 2] != kk) ranges[n_tiles + kk] = e + 1;
        }
    }
}

// Launch order of the compositor's tiles: longest list first (LPT), so that the few very long tiles
// of a frame start at once instead of wherever their index falls.  Counting sort on the number of
// 256-entry rounds, one CTA.
__global__ void __launch_bounds__(1024) k_tile_order(const uint32_t* __restrict__ ranges, uint32_t n_tiles,
                                                     uint32_t* __restrict__ order) {
    __shared__ uint32_t s_cnt[256];
    const int tid = threadIdx.x;
    if (tid < 256) s_cnt[tid] = 0;
    __syncthreads();
    for (uint32_t t = tid; t < n_tiles; t += 1024) {
        const uint32_t len = ranges[n_tiles + t] - ranges[t];
        atomicAdd(&s_cnt[255u - min(255u, (len + 255u) >> 8)], 1u);   // bucket 0 = longest
    }
    __syncthreads();
    if (tid < 32) {  // exclusive scan of 256 buckets by one warp (8 per lane)
        uint32_t loc[8], sum = 0;
#pragma unroll
        for (int k = 0; k < 8; k++) { loc[k] = sum; sum += s_cnt[tid * 8 + k]; }
        uint32_t incl = sum;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const uint32_t x = __shfl_up_sync(0xffffffffu, incl, o);
            if (tid >= o) incl += x;
        }
        const uint32_t base = incl - sum;
#pragma unroll
        for (int k = 0; k < 8; k++) s_cnt[tid * 8 + k] = base + loc[k];
    }
    __syncthreads();
    for (uint32_t t = tid; t < n_tiles; t += 1024) {
        const uint32_t len = ranges[n_tiles + t] - ranges[t];
        order[atomicAdd(&s_cnt[255u - min(255u, (len + 255u) >> 8)], 1u)] = t;
    }
}

}  // namespace

size_t gs_bin_block_words(uint32_t capacity_candidates) { return (size_t)capacity_candidates / kBlock + 2; }

cudaError_t gs_launch_bin(const GsBinArgs& a, const GsFrame& f, int num_sms, cudaStream_t st) {
    static int bps_count = 0, bps_emit = 0;
    if (bps_count == 0) {
        cudaError_t e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&bps_count, k_bin_count, kThreads, 0);
        if (e != cudaSuccess) return e;
        e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&bps_emit, k_bin_emit, kThreads, 0);
        if (e != cudaSuccess) return e;
        if (bps_count < 1) bps_count = 1;
        if (bps_emit < 1) bps_emit = 1;
    }
    const uint32_t flat = f.display_mode != B200GS_DISPLAY_SPLAT ? 1u : 0u;
    uint32_t nchunks = (a.v_max + kChunk - 1) / kChunk;
    uint32_t grid = (uint32_t)(bps_count * num_sms);
    if (grid > nchunks) grid = nchunks;
    if (grid < 1) grid = 1;
    k_bin_count<<<grid, kThreads, 0, st>>>(a.sorted_slot, a.sorted_slot_b, a.sorted_in_b, a.ncand, a.d_v, a.v_max, a.lookback, a.epoch, a.ticket,
                                           a.cand_off, a.block_rank, a.block_cap, a.cand_total, a.q_lo, a.q_hi);
    k_bin_emit<<<(uint32_t)(bps_emit * num_sms), kThreads, 0, st>>>(
        a.splats, a.d_v, a.v_max, a.splat_base, a.cand_off, a.block_rank, a.block_cap, a.cand_total,
        a.lookback_emit,
        a.epoch, a.ticket + 1, a.entry_base_in, a.entry_total_out, a.overflow, a.tile_keys, a.tile_vals, a.capacity,
        a.tile_hist, f.W, f.H, f.tiles_x, flat, a.q_lo, a.q_hi, a.tile_done);
    return cudaGetLastError();
}

cudaError_t gs_launch_tile_ranges(const uint32_t* keys_a, const uint32_t* keys_b, const uint32_t* in_b,
                                  const uint32_t* d_entries, uint32_t capacity, uint32_t* ranges, uint32_t n_tiles,
                                  unsigned long long* entry_stat, int num_sms, cudaStream_t st) {
    cudaError_t e = cudaMemsetAsync(ranges, 0, (size_t)n_tiles * 2 * sizeof(uint32_t), st);
    if (e != cudaSuccess) return e;
    k_tile_ranges<<<num_sms * 8, 256, 0, st>>>(keys_a, keys_b, in_b, d_entries, capacity, ranges, n_tiles, entry_stat);
    k_tile_order<<<1, 1024, 0, st>>>(ranges, n_tiles, ranges + 2 * (size_t)n_tiles);
    return cudaGetLastError();
}
