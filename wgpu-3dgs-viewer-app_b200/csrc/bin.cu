// bin.cu — K3a: binning for the compositor.
//
// Part of the replacement of renderer.render_with_pass (reference src/tab/scene.rs:2302-2314):
// the reference draws one instanced quad per visible Gaussian in sorted order and lets the
// rasteriser find the covered pixels; here every depth-sorted splat is expanded into one
// (bin id, splat id) entry per 32x32-pixel BIN its candidate rectangle touches.  A bin is 2 x 2 compositor
// tiles (16x16 pixels, one CTA each); every entry's key also carries, above the bin id, the mask of the
// bin's quadrants the rectangle reaches, and the compositor CTA of a quadrant stages only the entries with
// its bit.  Binning at 32 pixels emits ~40 % fewer entries than binning at 16 (a garden-scale splat is
// 5-15 pixels across), which is what the bin sort and this kernel are paid per.  Entries are produced IN
// DEPTH ORDER, so a STABLE sort by bin id alone (onesweep passes over the bin id bits, sort.cu)
// yields per-bin lists that are still front-to-back.  Models are expanded nearest first, each
// appended after the previous one, which reproduces the reference's per-model layering
// (scene.rs:533-558).
//
// One kernel, k_bin, persistent CTAs over chunks of 1024 depth ranks:
//   * the preprocess kernel left a 4-byte BIN WORD per compaction slot (candidate rectangle of the splat in
//     16-pixel tiles; common.cuh).  A thread reads sorted slot -> bin word (an L2-resident 4-byte gather) and
//     knows its entry count without touching the 32-byte splat;
//   * the common small splat (<= 4 bins) is expanded by its thread; a bigger one by its whole warp, 32 bins
//     per round; only the rare huge splat (a side of more than 32 tiles) has its rectangle rebuilt from the
//     stored record;
//   * order-preserving compaction: per-rank counts -> block scan -> decoupled look-back over the chunks
//     (pipelined by one chunk: chunk c+1 is counted and published before chunk c is resolved and
//     written), entries written at their final depth-ordered position;
//   * every written entry bumps its bin's counter (RED), from which k_tile_finish derives the per-bin
//     list boundaries, the digit histograms of the bin sort and the compositor's launch order — no pass
//     over the entries is needed for any of them.
// The candidate rectangle is kept whole (see common.cuh): the compositor culls every staged splat against
// its sub-tiles exactly, so a kept quadrant the splat cannot reach costs one staged record, not pixels.
#include <mutex>

#include "common.cuh"

namespace {

constexpr int kThreads = 256;
constexpr int kWarps = kThreads / 32;
constexpr int kIpt = 4;                      // depth ranks per thread
constexpr int kChunk = kThreads * kIpt;      // 1024 ranks per chunk; rank = chunk base + k * 256 + tid
constexpr uint32_t kStage = 4096;            // entries of one chunk staged in shared memory for a coalesced write-out

// candidate rectangle of a splat in 16-pixel tiles: from the bin word, or rebuilt from the stored record for a huge
// one (every lane for its own splat, so the loads of a warp's huge splats are in flight together)
struct Rect { uint32_t tx0, ty0, tx1, ty1; };
__device__ __forceinline__ Rect word_rect(uint32_t w) {
    Rect r;
    r.tx0 = w & 1023u; r.ty0 = (w >> 10) & 1023u;
    r.tx1 = r.tx0 + ((w >> 20) & 31u); r.ty1 = r.ty0 + ((w >> 25) & 31u);
    return r;
}
__device__ __forceinline__ uint32_t rect_bins(const Rect& r) { return ((r.tx1 >> 1) - (r.tx0 >> 1) + 1u) * ((r.ty1 >> 1) - (r.ty0 >> 1) + 1u); }
__device__ __forceinline__ uint32_t word_count(uint32_t w) {
    if (w & GS_BIN_HUGE) return w & 0x3fffffffu;
    return w ? rect_bins(word_rect(w)) : 0u;
}
__device__ __forceinline__ Rect huge_rect(const b200gs_splat* __restrict__ splats, uint32_t slot, float W, float H, bool flat) {
    const uint4* sp = reinterpret_cast<const uint4*>(splats + slot);
    const uint4 q0 = __ldg(sp), q1 = __ldg(sp + 1);
    GsCand cd;
    Rect r = {1u, 1u, 0u, 0u};
    if (gs_make_rect(q0, q1, W, H, flat, cd)) { r.tx0 = cd.tx0; r.ty0 = cd.ty0; r.tx1 = cd.tx0 + cd.nx - 1u; r.ty1 = cd.ty0 + cd.ny - 1u; }
    return r;
}
// the whole warp walks the bins of ONE big splat, 32 per round
__device__ __forceinline__ void big_rounds(const Rect r, uint32_t bins_x, int lane, uint32_t o, uint32_t val,
                                           uint32_t capacity, uint32_t* tile_keys, uint32_t* tile_vals,
                                           uint32_t* __restrict__ tile_count, uint32_t count_off) {
    const uint32_t bx0 = r.tx0 >> 1, by0 = r.ty0 >> 1, nbx = (r.tx1 >> 1) - bx0 + 1u;
    const uint32_t total = nbx * ((r.ty1 >> 1) - by0 + 1u);
    const float inv_nx = __frcp_rn((float)nbx);
    for (uint32_t e0 = 0; e0 < total; e0 += 32) {
        const uint32_t e = e0 + lane;
        // e / nbx without the integer-division sequence (e < 2^24, nbx <= 512: exact after one fix-up)
        uint32_t y = (uint32_t)(((float)e + 0.5f) * inv_nx);
        if (y * nbx > e) y--;
        else if ((y + 1) * nbx <= e) y++;
        const uint32_t bx = bx0 + (e - y * nbx), by = by0 + y;
        const uint32_t g = o + e;
        if (e < total && g < capacity) {
            const uint32_t bin = by * bins_x + bx;
            tile_keys[g] = bin | (gs_quadrant_mask(bx, by, r.tx0, r.tx1, r.ty0, r.ty1) << GS_QMASK_SHIFT);
            tile_vals[g] = val;
            atomicAdd(&tile_count[count_off + bin], 1u);
        }
    }
}

// The entries of the kIpt ranks of one thread (bin word, slot, offset inside the chunk).  A small splat (<= GS_BIN_INLINE
// bins: a run along a row, a run down a column, or a 2 x 2 block) is expanded by its thread, walking (bx, by, bin) from
// entry to entry with per-splat steps, so that an entry costs three additions, the quadrant mask, two stores and the
// counter bump; a bigger one by the whole warp.  Entries at offsets >= cap are dropped together with their counts.
__device__ __forceinline__ void expand_ranks(const uint32_t (&p_w)[kIpt], const uint32_t (&p_slot)[kIpt], const uint32_t (&p_loc)[kIpt],
                                             uint32_t obase, uint32_t cap, uint32_t splat_base, uint32_t* dst_k, uint32_t* dst_v,
                                             uint32_t* __restrict__ tile_count, uint32_t count_off,
                                             const b200gs_splat* __restrict__ splats, float W, float H, bool is_flat,
                                             uint32_t bins_x, int lane) {
#pragma unroll
    for (int k = 0; k < kIpt; k++) {
        const uint32_t wk = p_w[k];
        const uint32_t o = obase + p_loc[k];
        const uint32_t val = splat_base + p_slot[k];
        Rect r = word_rect(wk);
        const bool huge = (wk & GS_BIN_HUGE) != 0u;
        uint32_t n = 0;
        if (wk && !huge) {
            const uint32_t bx0 = r.tx0 >> 1, by0 = r.ty0 >> 1, nbx = (r.tx1 >> 1) - bx0 + 1u;
            n = nbx * ((r.ty1 >> 1) - by0 + 1u);
            if (n <= GS_BIN_INLINE) {
                const uint32_t ne = o < cap ? min(n, cap - o) : 0u;   // entries that fit
                // step from entry e - 1 to entry e: along the row; down for a single column; back and down at the row end of
                // a 2 x 2 block (n <= 4, so two columns and a third entry mean 2 x 2)
                const bool column = nbx == 1u, two = nbx == 2u;
                const uint32_t sx1 = column ? 0u : 1u, sy1 = column ? 1u : 0u;
                const uint32_t sx2 = column ? 0u : (two ? 0xffffffffu : 1u), sy2 = (column || two) ? 1u : 0u;
                const uint32_t sb1 = column ? bins_x : 1u, sb2 = column ? bins_x : (two ? bins_x - 1u : 1u);
                uint32_t bx = bx0, by = by0, bin = by0 * bins_x + bx0;
#pragma unroll
                for (uint32_t e = 0; e < GS_BIN_INLINE; e++) {
                    if (e == 1u || e == 3u) { bx += sx1; by += sy1; bin += sb1; }
                    if (e == 2u) { bx += sx2; by += sy2; bin += sb2; }
                    if (e < ne) {
                        dst_k[o + e] = bin | (gs_quadrant_mask(bx, by, r.tx0, r.tx1, r.ty0, r.ty1) << GS_QMASK_SHIFT);
                        dst_v[o + e] = val;
                        atomicAdd(&tile_count[count_off + bin], 1u);
                    }
                }
            }
        }
        uint32_t big = __ballot_sync(0xffffffffu, huge || n > GS_BIN_INLINE);
        if (big) {
            if (huge) r = huge_rect(splats, p_slot[k], W, H, is_flat);
            while (big) {
                const int src = __ffs((int)big) - 1;
                big &= big - 1;
                Rect rr;
                rr.tx0 = __shfl_sync(0xffffffffu, r.tx0, src); rr.ty0 = __shfl_sync(0xffffffffu, r.ty0, src);
                rr.tx1 = __shfl_sync(0xffffffffu, r.tx1, src); rr.ty1 = __shfl_sync(0xffffffffu, r.ty1, src);
                if (rr.tx0 <= rr.tx1)
                    big_rounds(rr, bins_x, lane, __shfl_sync(0xffffffffu, o, src), __shfl_sync(0xffffffffu, val, src), cap,
                               dst_k, dst_v, tile_count, count_off);
            }
        }
    }
}

__global__ void __launch_bounds__(kThreads) k_bin(const uint32_t* __restrict__ sorted_a, const uint32_t* __restrict__ sorted_b,
                                                  const uint32_t* sorted_in_b, const uint32_t* __restrict__ binword,
                                                  const b200gs_splat* __restrict__ splats, const uint32_t* d_v,
                                                  uint32_t v_max, uint32_t splat_base, uint64_t* lookback, uint32_t epoch,
                                                  uint32_t* ticket, const uint32_t* entry_base_in,
                                                  uint32_t* entry_total_out, uint32_t* overflow,
                                                  uint32_t* __restrict__ tile_keys, uint32_t* __restrict__ tile_vals,
                                                  uint32_t capacity, uint32_t* __restrict__ tile_count,
                                                  uint32_t count_copies, uint32_t count_stride, float W, float H,
                                                  uint32_t bins_x, uint32_t flat) {
    __shared__ uint32_t s_cnt[kIpt][kWarps];   // per (k, warp) kept counts -> exclusive offsets
    __shared__ uint32_t s_keys[kStage], s_vals[kStage];   // the chunk's entries, in order, before the write-out
    __shared__ uint32_t s_total;
    __shared__ uint32_t s_chunk, s_base;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const uint32_t* __restrict__ sorted_slot = (sorted_in_b && *sorted_in_b) ? sorted_b : sorted_a;
    // per-bin entry counters are replicated (copy = SM id mod copies): same-address atomics from many SMs
    // serialise in L2 and the centre bins of a frame are hot
    uint32_t count_off;
    {
        uint32_t smid;
        asm("mov.u32 %0, %%smid;" : "=r"(smid));
        // (a 32-bit word offset, opaque to the compiler: it would otherwise re-derive the replica's 64-bit address from the
        // parameters at every counter bump — seven instructions per entry — rather than hold it in registers)
        count_off = (smid & (count_copies - 1u)) * count_stride;
        asm volatile("" : "+r"(count_off));
    }
    uint32_t v = *d_v;
    if (v > v_max) v = v_max;
    const uint32_t nchunks = (v + kChunk - 1) / kChunk;
    const uint32_t ebase = *entry_base_in;
    if (v == 0) {
        if (blockIdx.x == 0 && tid == 0) *entry_total_out = ebase;
        return;
    }
    const bool is_flat = flat != 0;

    // Software-pipelined by one chunk: chunk c+1 is counted and its aggregate published BEFORE chunk c's
    // prefix is resolved, so the look-back never waits for chunks drawn at the same moment.
    bool have_prev = false;
    uint32_t p_c = 0, p_total = 0, p_w[kIpt], p_slot[kIpt], p_loc[kIpt];
#pragma unroll
    for (int k = 0; k < kIpt; k++) p_w[k] = p_slot[k] = p_loc[k] = 0;
    while (true) {
        __syncthreads();
        if (tid == 0) s_chunk = atomicAdd(ticket, 1u);
        __syncthreads();
        const uint32_t c = s_chunk;
        const bool valid = c < nchunks;
        uint32_t w[kIpt], slot[kIpt], loc[kIpt], chunk_total = 0;
#pragma unroll
        for (int k = 0; k < kIpt; k++) w[k] = slot[k] = loc[k] = 0;
        if (valid) {
            // ---------------- phase A: entry count of every rank of the chunk (no splat is touched)
            const uint32_t r0 = c * kChunk + tid;
#pragma unroll
            for (int k = 0; k < kIpt; k++) {
                const uint32_t r = r0 + k * kThreads;
                if (r < v) slot[k] = sorted_slot[r];
            }
#pragma unroll
            for (int k = 0; k < kIpt; k++) {
                const uint32_t r = r0 + k * kThreads;
                if (r < v) w[k] = __ldg(&binword[slot[k]]);
            }
#pragma unroll
            for (int k = 0; k < kIpt; k++) {
                const uint32_t n = word_count(w[k]);
                // warp-inclusive scan of the counts of slot k
                uint32_t incl = n;
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) {
                    const uint32_t t = __shfl_up_sync(0xffffffffu, incl, o);
                    if (lane >= o) incl += t;
                }
                loc[k] = incl - n;
                if (lane == 31) s_cnt[k][warp] = incl;
            }
            __syncthreads();
            // exclusive scan of the 32 (k, warp) counts in rank order, by warp 0
            if (warp == 0) {
                const uint32_t a = s_cnt[lane >> 3][lane & 7];
                uint32_t incl = a;
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) {
                    const uint32_t t = __shfl_up_sync(0xffffffffu, incl, o);
                    if (lane >= o) incl += t;
                }
                s_cnt[lane >> 3][lane & 7] = incl - a;
                if (lane == 31) {
                    s_total = incl;
                    gs_lookback_publish(lookback, epoch, c, incl);
                }
            }
            __syncthreads();
            chunk_total = s_total;
#pragma unroll
            for (int k = 0; k < kIpt; k++) loc[k] += s_cnt[k][warp];
        }
        if (have_prev) {
            // ---------------- phase B: resolve the previous chunk's base, write its entries
            if (warp == 0) {
                const uint32_t excl = gs_lookback_resolve(lookback, epoch, p_c, p_total, lane);
                if (lane == 0) {
                    s_base = excl;
                    if (p_c == nchunks - 1) {
                        uint32_t t = ebase + excl + p_total;
                        if (t > capacity || t < ebase) { *overflow = 1u; t = capacity; }
                        *entry_total_out = t;
                    }
                }
            }
            __syncthreads();
            const uint32_t gbase = ebase + s_base;
            // Entries go through shared memory (the usual chunk holds ~1.2 K of them) so that the global writes
            // are whole lines; a chunk that does not fit (the nearest, hugest splats) writes directly.  Either
            // way an entry past the capacity is dropped together with its count.
            const bool staged = p_total <= kStage;
            const uint32_t room = capacity > gbase ? capacity - gbase : 0u;
            const uint32_t cap = staged ? min(room, kStage) : capacity;
            // (two instantiations, so that the staged one stores with shared-memory instructions at immediate offsets
            // instead of through a generic pointer chosen at run time)
            if (staged) expand_ranks(p_w, p_slot, p_loc, 0u, cap, splat_base, s_keys, s_vals, tile_count, count_off, splats, W, H, is_flat, bins_x, lane);
            else expand_ranks(p_w, p_slot, p_loc, gbase, cap, splat_base, tile_keys, tile_vals, tile_count, count_off, splats, W, H, is_flat, bins_x, lane);
            if (staged) {
                __syncthreads();
                const uint32_t nw = min(p_total, cap);
                for (uint32_t i = tid; i < nw; i += kThreads) {
                    tile_keys[gbase + i] = s_keys[i];
                    tile_vals[gbase + i] = s_vals[i];
                }
            }
        }
        if (!valid) break;
        have_prev = true;
        p_c = c;
        p_total = chunk_total;
#pragma unroll
        for (int k = 0; k < kIpt; k++) { p_w[k] = w[k]; p_slot[k] = slot[k]; p_loc[k] = loc[k]; }
    }
}

// Everything the bin sort and the compositor need from the replicated per-bin counters, in one kernel
// ("tile" in the names below is a 32-pixel bin): per-bin sums (counters cleared on the way for the next frame), exclusive scan -> per-tile list
// boundaries (ranges[tile] = first entry, ranges[n_tiles + tile] = one past the last), the digit
// histograms of the tile sort, the total, and the compositor's launch order: longest list first (LPT), so
// that the few very long tiles of a frame start at once instead of wherever their index falls — a counting
// sort on the number of 256-entry rounds.  Chunks of 64 tiles per CTA, chained by decoupled look-back; the
// last CTA to finish turns the bucket ranks into the launch order.
constexpr int kTfTiles = 64;
__global__ void __launch_bounds__(256) k_tile_finish(uint32_t* __restrict__ replicas, uint32_t copies, uint32_t stride,
                                                     uint32_t n_tiles, uint32_t n_chunks, uint32_t* __restrict__ ranges,
                                                     uint32_t* __restrict__ hist, uint32_t passes,
                                                     unsigned long long* entry_stat, uint64_t* lookback, uint32_t epoch,
                                                     uint32_t* ticket, uint32_t* done_ctr, uint32_t* buckets) {
    __shared__ uint32_t s_part[4][kTfTiles];
    __shared__ uint32_t s_bk[256];
    __shared__ uint32_t s_chunk, s_last;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    uint32_t* __restrict__ tmp = ranges + 3 * (size_t)n_tiles;     // per tile: bucket << 24 | rank inside the bucket
    uint32_t* __restrict__ order = ranges + 2 * (size_t)n_tiles;
    if (tid == 0) s_chunk = atomicAdd(ticket, 1u);
    __syncthreads();
    const uint32_t c = s_chunk;
    {
        // 4 groups of threads share the copies of the chunk's 64 tiles
        const uint32_t q = tid >> 6, t = c * kTfTiles + (tid & 63);
        const uint32_t per = (copies + 3u) / 4u, c0 = min(copies, q * per), c1 = min(copies, c0 + per);
        uint32_t sum = 0;
        if (t < n_tiles) {
            for (uint32_t cc = c0; cc < c1; cc += 8) {
                uint32_t x[8];
#pragma unroll
                for (int j = 0; j < 8; j++) x[j] = cc + j < c1 ? replicas[(size_t)(cc + j) * stride + t] : 0u;
#pragma unroll
                for (int j = 0; j < 8; j++) {
                    sum += x[j];
                    if (x[j]) replicas[(size_t)(cc + j) * stride + t] = 0u;
                }
            }
        }
        s_part[q][tid & 63] = sum;
    }
    __syncthreads();
    if (warp == 0) {
        // lane l owns tiles 2l and 2l + 1 of the chunk
        const uint32_t a0 = s_part[0][2 * lane] + s_part[1][2 * lane] + s_part[2][2 * lane] + s_part[3][2 * lane];
        const uint32_t a1 = s_part[0][2 * lane + 1] + s_part[1][2 * lane + 1] + s_part[2][2 * lane + 1] + s_part[3][2 * lane + 1];
        uint32_t incl = a0 + a1;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const uint32_t x = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= o) incl += x;
        }
        const uint32_t total = __shfl_sync(0xffffffffu, incl, 31);
        const uint32_t base = gs_lookback_warp(lookback, epoch, c, total, lane);
        if (c == n_chunks - 1 && lane == 0 && entry_stat) atomicAdd(entry_stat, (unsigned long long)(base + total));
        uint32_t run = base + incl - (a0 + a1);
#pragma unroll
        for (int h = 0; h < 2; h++) {
            const uint32_t t = c * kTfTiles + 2 * lane + h, cnt = h ? a1 : a0;
            if (t < n_tiles) {
                ranges[t] = run;
                ranges[n_tiles + t] = run + cnt;
                if (cnt)
                    for (uint32_t p = 0; p < passes; p++) atomicAdd(&hist[p * 256 + ((t >> (8 * p)) & 0xffu)], cnt);
                const uint32_t bk = 255u - min(255u, (cnt + 255u) >> 8);   // bucket 0 = longest
                tmp[t] = (bk << 24) | atomicAdd(&buckets[bk], 1u);
            }
            run += cnt;
        }
    }
    // ---- the last CTA to get here turns (bucket, rank) into the launch order
    __threadfence();
    __syncthreads();
    if (tid == 0) s_last = atomicAdd(done_ctr, 1u) == n_chunks - 1 ? 1u : 0u;
    __syncthreads();
    if (!s_last) return;
    __threadfence();
    s_bk[tid] = __ldcg(&buckets[tid]);
    __syncthreads();
    if (warp == 0) {  // exclusive scan of 256 buckets by one warp (8 per lane)
        uint32_t loc[8], sum = 0;
#pragma unroll
        for (int k = 0; k < 8; k++) { loc[k] = sum; sum += s_bk[lane * 8 + k]; }
        uint32_t incl = sum;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const uint32_t x = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= o) incl += x;
        }
        const uint32_t base = incl - sum;
#pragma unroll
        for (int k = 0; k < 8; k++) s_bk[lane * 8 + k] = base + loc[k];
    }
    __syncthreads();
    for (uint32_t t0 = tid; t0 < n_tiles; t0 += 256 * 8) {
        uint32_t x[8];
#pragma unroll
        for (int j = 0; j < 8; j++) x[j] = t0 + j * 256 < n_tiles ? __ldcg(&tmp[t0 + j * 256]) : 0u;
#pragma unroll
        for (int j = 0; j < 8; j++)
            if (t0 + j * 256 < n_tiles) order[s_bk[x[j] >> 24] + (x[j] & 0xffffffu)] = t0 + j * 256;
    }
}

}  // namespace

// replicas of the per-tile counters: a power of two <= 128, at most 2M words in total
uint32_t gs_tile_count_copies(uint32_t n_tiles) {
    uint32_t c = 128;   // measured on a 6M-splat frame: 8 copies 186 us, 32 copies 152 us, 128 copies 144 us
    while (c > 1 && (uint64_t)c * n_tiles > (2u << 20)) c >>= 1;
    return c;
}
// enough for this and every smaller viewport (the copy count grows as the tile count shrinks)
size_t gs_tile_count_words(uint32_t n_tiles) { return (size_t)(2u << 20) + (size_t)n_tiles; }

cudaError_t gs_launch_bin(const GsBinArgs& a, const GsFrame& f, int num_sms, cudaStream_t st) {
    static std::mutex mu;
    static int bps_dev[64] = {0};   // per device
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return e;
    if (dev < 0 || dev >= 64) return cudaErrorInvalidDevice;
    int bps;
    {
        std::lock_guard<std::mutex> lock(mu);
        if (bps_dev[dev] == 0) {
            e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&bps_dev[dev], k_bin, kThreads, 0);
            if (e != cudaSuccess) { bps_dev[dev] = 0; return e; }
            if (bps_dev[dev] < 1) bps_dev[dev] = 1;
        }
        bps = bps_dev[dev];
    }
    const uint32_t flat = f.display_mode != B200GS_DISPLAY_SPLAT ? 1u : 0u;
    const uint32_t n_bins = f.bins_x * f.bins_y;
    uint32_t nchunks = (a.v_max + kChunk - 1) / kChunk;
    uint32_t grid = (uint32_t)(bps * num_sms);
    if (grid > nchunks) grid = nchunks;
    if (grid < 1) grid = 1;
    k_bin<<<grid, kThreads, 0, st>>>(a.sorted_slot, a.sorted_slot_b, a.sorted_in_b, a.binword, a.splats, a.d_v, a.v_max,
                                     a.splat_base, a.lookback, a.epoch, a.ticket, a.entry_base_in, a.entry_total_out,
                                     a.overflow, a.tile_keys, a.tile_vals, a.capacity, a.tile_count,
                                     gs_tile_count_copies(n_bins), n_bins, f.W, f.H, f.bins_x, flat);
    return cudaGetLastError();
}

size_t gs_tile_lookback_words(uint32_t n_tiles) { return (size_t)(n_tiles + kTfTiles - 1) / kTfTiles + 1; }

cudaError_t gs_launch_tile_ranges(const GsTileRangesArgs& a, cudaStream_t st) {
    const uint32_t n_chunks = (a.n_tiles + kTfTiles - 1) / kTfTiles;
    k_tile_finish<<<n_chunks, 256, 0, st>>>(a.tile_count, gs_tile_count_copies(a.n_tiles), a.n_tiles, a.n_tiles, n_chunks,
                                            a.ranges, a.hist, a.passes, a.entry_stat, a.lookback, a.epoch, a.ticket,
                                            a.done_ctr, a.buckets);
    return cudaGetLastError();
}
