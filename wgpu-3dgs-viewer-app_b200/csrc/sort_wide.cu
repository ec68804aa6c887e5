// sort_wide.cu — K2w: onesweep-style LSD radix sort of (u32 key, u32 value) pairs, 11-bit digits, one thread-block
// cluster per super-tile.  Used where it saves a pass: the bin ids of the binning stage (<= 2048 bins: ONE pass).
//
// Replaces viewer.radix_sorter.sort(encoder, bind_group, radix_sort_indirect_args)
// (reference src/tab/scene.rs:865-869): stable ascending sort of the depth keys (f32 bits as
// u32) with the Gaussian indices as payload; the element count lives on the device (the
// reference sizes an indirect dispatch from it), here `*d_n`.
//
// Design.  A digit pass is a single sweep over the keys (one read, one write): tiles rank their keys
// with warp-level same-digit peer masks, publish per-digit counts, resolve global offsets by decoupled
// look-back over epoch-tagged status words (no scan kernel, no second read), and scatter through shared
// memory so that global writes are runs of consecutive addresses.  Round 1 used 8-bit digits: 3 executed
// passes for the ~23 live bits of a depth key, each bound by its per-pass overhead (and by 2-cycle-per-lane
// shared-memory atomics in the counting step), not by HBM.  This version sorts 11 bits per pass — depth keys
// in [0.5, 1) take TWO passes (bits 0..10, 11..21; the 10-bit top digit is degenerate and skipped) and the
// tile ids of a 1080p frame ONE — which needs 2048 status words per look-back step instead of 256.  To keep
// that affordable the unit of look-back is a THREAD-BLOCK CLUSTER of 8 CTAs: each CTA ranks 4096 keys in
// shared memory, the cluster's 32768 keys form one super-tile, and CTA r of the cluster owns digits
// [256 r, 256 r + 256): one digit per thread.  The owner reads the 8 CTAs' counts of its digit through
// distributed shared memory (DSMEM), publishes the super-tile's count, walks the look-back, and writes every
// CTA's global base for the digit back into that CTA's shared memory.  Status traffic per key drops 8x
// (16 KB per 32768 keys) and only ~50 super-tiles are in flight, so look-back walks stay short.
// Ranking uses no shared-memory atomics: per-warp u16 histograms are updated by one leader lane per digit.
// Super-tiles are handed out by an atomic ticket (drawn by CTA 0 of the cluster, distributed through DSMEM),
// so that every predecessor a super-tile waits on is owned by a running cluster.
#include <atomic>
#include <initializer_list>
#include <mutex>

#include "common.cuh"

namespace {

constexpr int kThreads = 256;
constexpr int kWarps = kThreads / 32;
constexpr int kKpt = 16;                         // keys per thread
constexpr int kTile = kThreads * kKpt;           // 4096 keys per CTA
constexpr int kMaxCluster = 8;                   // CTAs per cluster: 8 (default), 4, 2 or 1 (tuning knob, gs_sort_set_cluster)
constexpr int kBins = GS_SORT_BINS;              // 2048
constexpr int kWarpSpan = 32 * kKpt;             // 512 consecutive keys per warp
static_assert(kBins == kMaxCluster * kThreads, "one owned digit per thread at the largest cluster size");

// ------------------------------------------------------------------ histogram kernel (raw sort API only)
// hist[pass][digit] += count.  The frame path never runs it: the preprocess kernel and the tile-finish
// kernel accumulate the histograms of the keys they produce.
__global__ void __launch_bounds__(kThreads) k_sort_wide_hist(const uint32_t* __restrict__ keys, const uint32_t* d_n,
                                                        uint32_t n_max, uint32_t* hist, uint32_t key_bits) {
    extern __shared__ uint32_t s_hist[];   // passes x kBins
    const uint32_t passes = gs_sort_passes(key_bits);
    for (uint32_t i = threadIdx.x; i < passes * kBins; i += kThreads) s_hist[i] = 0;
    __syncthreads();
    uint32_t n = *d_n;
    if (n > n_max) n = n_max;
    for (uint32_t i = blockIdx.x * kThreads + threadIdx.x; i < n; i += gridDim.x * kThreads) {
        const uint32_t k = keys[i];
        for (uint32_t p = 0; p < passes; p++) {
            const uint32_t bits = min((uint32_t)GS_SORT_DIGIT_BITS, key_bits - GS_SORT_DIGIT_BITS * p);
            atomicAdd(&s_hist[p * kBins + ((k >> (GS_SORT_DIGIT_BITS * p)) & ((1u << bits) - 1u))], 1u);
        }
    }
    __syncthreads();
    for (uint32_t i = threadIdx.x; i < passes * kBins; i += kThreads) {
        const uint32_t c = s_hist[i];
        if (c) atomicAdd(&hist[i], c);
    }
}

// ---------------------------------------------------------------------- cluster / DSMEM helpers
__device__ __forceinline__ uint32_t cl_rank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ uint32_t cl_map(const void* p, uint32_t rank) {  // my shared address -> the same variable in CTA `rank`
    uint32_t r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(gs_smem_u32(p)), "r"(rank));
    return r;
}
__device__ __forceinline__ void cl_st_u32(uint32_t addr, uint32_t v) {
    asm volatile("st.shared::cluster.u32 [%0], %1;" ::"r"(addr), "r"(v) : "memory");
}
// Full cluster barrier with release / acquire semantics.  ptxas implements the cluster-scope release as MEMBAR.ALL.GPU
// (+ CCTL.IVALL on the acquire side), far too heavy for a per-tile handshake: it is used ONCE, at kernel start.
__device__ __forceinline__ void cl_sync() {
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// Asynchronous DSMEM store that signals the destination CTA's mbarrier with the bytes it delivered (SASS: STAS): data
// and completion travel together, so the receiver needs no fence — it waits on its own mbarrier.
__device__ __forceinline__ void cl_st_async_u32(uint32_t addr, uint32_t v, uint32_t mbar) {
    asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.b32 [%0], %1, [%2];" ::"r"(addr), "r"(v), "r"(mbar) : "memory");
}
__device__ __forceinline__ void cl_st_async_v4(uint32_t addr, uint4 v, uint32_t mbar) {
    asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.v4.b32 [%0], {%1, %2, %3, %4}, [%5];" ::"r"(addr), "r"(v.x),
                 "r"(v.y), "r"(v.z), "r"(v.w), "r"(mbar)
                 : "memory");
}

// ---------------------------------------------------------------------- digit pass
struct __align__(16) PassSmem {
    union {
        uint16_t whist[kWarps][kBins];   // per-warp digit counts -> exclusive offsets of the warp inside the CTA's digit run
        uint32_t gpos[kBins];            // (after the scatter) global index of this CTA's first key of each digit, sent by the digit's owner
    };
    uint32_t exch_k[kTile];              // keys / values of the CTA in digit order
    uint32_t exch_v[kTile];
    uint16_t rcnt[kBins];                // owner side: [r][x] = count in CTA r of my x-th owned digit, pushed by CTA r
    uint16_t tile_start[kBins];          // first position of each digit inside the sorted tile
    uint32_t scan_tmp[kWarps];
    uint32_t tile_id[2];                 // super-tile of this / the next iteration (sent by CTA 0 of the cluster)
    uint64_t bar_cnt;                    // mbarrier: one phase per tile, completes when all CTAs' counts of my digits are here
    uint64_t bar_gpos;                   // mbarrier: one phase per tile, completes when gpos[] (and the next ticket) are here
};

// peers of this lane = lanes of the warp whose digit equals mine.  One ballot per digit bit; MATCH.ANY costs ADU cycles
// per DISTINCT value and 11-bit digits are spread (measured in round 1: a MATCH.ANY pass over spread digits was
// ADU-bound).  The ballots of a row are independent of each other: they are issued back to back and only then
// combined (a serial vote -> xor -> and chain per bit left the warp waiting on the vote latency eleven times per row).
template <int BITS>
__device__ __forceinline__ uint32_t digit_peers(uint32_t key, uint32_t shift) {
    uint32_t bal[BITS];
#pragma unroll
    for (int b = 0; b < BITS; b++) bal[b] = __ballot_sync(0xffffffffu, (key >> (shift + b)) & 1u);
    uint32_t peers = 0xffffffffu;
#pragma unroll
    for (int b = 0; b < BITS; b++) peers &= ((key >> (shift + b)) & 1u) ? bal[b] : ~bal[b];
    return peers;
}

// CLAIM: wide digits are spread — most rows of 32 keys hold 32 DIFFERENT digits — so a row first tries the cheap
// test: every lane stores its lane id into a per-warp claim table at its digit and reads it back; if every lane reads
// its own id the row is collision-free, each lane is the only peer of its digit, and the eleven ballots are skipped.
template <int BITS, int CL, bool CLAIM>
__global__ void __launch_bounds__(kThreads, 3)
k_sort_wide_pass(uint32_t* __restrict__ keys_a, uint32_t* __restrict__ vals_a, uint32_t* __restrict__ keys_b,
            uint32_t* __restrict__ vals_b, const uint32_t* d_n, uint32_t n_max, const uint32_t* __restrict__ hist_all,
            uint32_t pass, uint32_t key_bits, uint64_t* lookback, uint32_t epoch, uint32_t* ticket, uint32_t* result_in_b,
            uint32_t vals_identity) {
    extern __shared__ __align__(16) uint8_t smem_raw[];
    PassSmem& sm = *reinterpret_cast<PassSmem*>(smem_raw);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    constexpr int kSuper = kTile * CL;        // keys per super-tile (one cluster)
    constexpr int kOwn = kMaxCluster / CL;    // digits owned per thread: (rank * 256 + tid) * kOwn + j
    const uint32_t rank = CL > 1 ? cl_rank() : 0u;
    uint32_t n = *d_n;
    if (n > n_max) n = n_max;
    const uint32_t nsuper = (n + kSuper - 1) / kSuper;
    const uint32_t passes = gs_sort_passes(key_bits);
    const uint32_t shift = GS_SORT_DIGIT_BITS * pass;
    const uint32_t dmask = (1u << BITS) - 1u;

    // A digit whose histogram has a single non-empty bin leaves the order unchanged: the pass is skipped
    // (depth keys in [0.5, 1) share their top 10 bits).  Every CTA derives the same plan from the global
    // histograms: which passes run, hence which buffer holds this pass's input.
    uint32_t executed_before = 0;
    bool skip_me = false;
    for (uint32_t q = 0; q < passes; q++) {
        const uint4* h4 = reinterpret_cast<const uint4*>(hist_all + q * kBins) + 2 * tid;
        const uint4 x = h4[0], y = h4[1];
        const bool hit = n > 0 && (x.x == n || x.y == n || x.z == n || x.w == n || y.x == n || y.y == n || y.z == n || y.w == n);
        const int degenerate = __syncthreads_or(hit);
        if (q < pass) executed_before += degenerate ? 0u : 1u;
        if (q == pass) skip_me = degenerate != 0;
    }
    const bool src_b = (executed_before & 1u) != 0;
    if (pass == passes - 1 && blockIdx.x == 0 && tid == 0) *result_in_b = ((executed_before + (skip_me ? 0u : 1u)) & 1u);
    if (skip_me || n == 0) return;   // (uniform over the grid)
    const uint32_t* __restrict__ keys_in = src_b ? keys_b : keys_a;
    const uint32_t* __restrict__ vals_in = src_b ? vals_b : vals_a;
    uint32_t* __restrict__ keys_out = src_b ? keys_a : keys_b;
    uint32_t* __restrict__ vals_out = src_b ? vals_a : vals_b;
    const bool synth_vals = vals_identity && executed_before == 0;  // first executed pass: value = input position

    // global bases of the digits this thread owns: exclusive prefix of the pass's histogram
    const uint32_t own = (rank * kThreads + tid) * kOwn;
    uint32_t gbase[kOwn];
    {
        const uint4* h4 = reinterpret_cast<const uint4*>(hist_all + pass * kBins) + 2 * tid;
        const uint4 x = h4[0], y = h4[1];
        const uint32_t c[8] = {x.x, x.y, x.z, x.w, y.x, y.y, y.z, y.w};
        uint32_t sum = 0;
#pragma unroll
        for (int j = 0; j < 8; j++) sum += c[j];
        uint32_t incl = sum;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const uint32_t t = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= o) incl += t;
        }
        if (lane == 31) sm.scan_tmp[warp] = incl;
        __syncthreads();
        uint32_t run = incl - sum;
        for (int k = 0; k < warp; k++) run += sm.scan_tmp[k];
#pragma unroll
        for (int j = 0; j < 8; j++) { sm.exch_k[8 * tid + j] = run; run += c[j]; }
        __syncthreads();
#pragma unroll
        for (int j = 0; j < kOwn; j++) gbase[j] = sm.exch_k[own + j];
    }
    uint64_t* lb = lookback + own;

    // Cluster protocol (CL > 1).  Per tile, every CTA pushes its digit counts to the digits' owners and every owner
    // pushes the global bases back, both with st.async + complete_tx on the receiver's mbarrier: no cluster barrier
    // and no fence inside the loop.  Each mbarrier runs one phase per tile (one arrival: thread 0's expect_tx, armed a
    // whole tile ahead; the transaction bytes do the rest).
    constexpr uint32_t kCntBytes = kBins * 2;          // counts of my owned digits from all CTAs
    constexpr uint32_t kGposBytes = kBins * 4 + 4;     // gpos[] + the next ticket
    if (CL > 1) {
        if (tid == 0) {
            gs_mbar_init(&sm.bar_cnt, 1);
            gs_mbar_init(&sm.bar_gpos, 1);
            gs_fence_mbar_init();
            gs_mbar_expect_tx(&sm.bar_cnt, kCntBytes);
            gs_mbar_expect_tx(&sm.bar_gpos, kGposBytes);
        }
        cl_sync();   // every CTA of the cluster is running, its barriers are armed: its shared memory may be written
    }
    if (rank == 0 && tid == 0) {
        const uint32_t t = atomicAdd(ticket, 1u);
        if (CL > 1) {
#pragma unroll
            for (int r = 0; r < CL; r++) cl_st_u32(cl_map(&sm.tile_id[0], r), t);
        } else sm.tile_id[0] = t;
    }
    if (CL > 1) cl_sync();
    else __syncthreads();

    for (uint32_t it = 0;; it++) {
        const uint32_t super = sm.tile_id[it & 1u];
        if (super >= nsuper) break;   // (uniform over the cluster)
        // ---- clear the per-warp histograms (the previous write-out, which read gpos = the same memory, is done:
        // barrier at the end of the loop body)
        {
            uint4* z = reinterpret_cast<uint4*>(&sm.whist[0][0]);
#pragma unroll
            for (int k = 0; k < (int)(sizeof(sm.whist) / 16 / kThreads); k++) z[k * kThreads + tid] = make_uint4(0, 0, 0, 0);
        }
        // ---- load keys (warp-striped: slot = warp*512 + k*32 + lane keeps index order inside a warp); slots past
        // the end hold 0xffffffff: they carry the largest digit and the highest positions, so they rank behind
        // every real key of that digit and fall off the end of the sorted tile
        const uint32_t tile_base = super * kSuper + rank * kTile;
        const uint32_t valid = tile_base < n ? min((uint32_t)kTile, n - tile_base) : 0u;
        const uint32_t wslot = warp * kWarpSpan + lane;
        uint32_t key[kKpt];
#pragma unroll
        for (int k = 0; k < kKpt; k++) {
            const uint32_t s = wslot + k * 32;
            key[k] = s < valid ? keys_in[tile_base + s] : 0xffffffffu;
        }
        __syncthreads();
        // ---- rank inside the warp: position among the warp's earlier keys of the same digit
        // Two phases, so that the votes of all 16 rows can overlap (they touch no memory and do not depend on each
        // other), and only the short counter update runs as a dependent chain through shared memory:
        //   1. peer masks of every row;
        //   2. row by row: every peer reads the warp's counter of its digit, the first peer advances it.
        uint32_t rk[kKpt / 2];   // two u16 ranks per register
        uint16_t* wh = sm.whist[warp];
        const uint32_t lane_lt = (1u << lane) - 1u;
        uint32_t peers[kKpt];
        if (CLAIM) {
            // rows whose 32 digits are all different (the common case for spread digits) skip the votes: every lane
            // stores its lane id into a per-warp claim table at its digit and reads it back
            uint8_t* claim = reinterpret_cast<uint8_t*>(sm.exch_k) + warp * kBins;   // (the exchange buffer is idle while ranking)
#pragma unroll
            for (int k = 0; k < kKpt; k++) {
                const uint32_t d = (key[k] >> shift) & dmask;
                claim[d] = (uint8_t)lane;
                __syncwarp();
                const bool alone = !__any_sync(0xffffffffu, claim[d] != (uint8_t)lane);
                __syncwarp();
                peers[k] = alone ? (1u << lane) : digit_peers<BITS>(key[k], shift);   // (warp-uniform branch)
            }
        } else {
#pragma unroll
            for (int k = 0; k < kKpt; k++) peers[k] = digit_peers<BITS>(key[k], shift);
        }
#pragma unroll
        for (int k = 0; k < kKpt; k++) {
            const uint32_t d = (key[k] >> shift) & dmask;
            const uint32_t before = wh[d];                    // every peer reads the same counter ...
            const uint32_t mine = __popc(peers[k] & lane_lt);
            __syncwarp();
            if (mine == 0) wh[d] = (uint16_t)(before + __popc(peers[k]));   // ... and the first peer advances it
            __syncwarp();
            const uint32_t r = before + mine;
            if (k & 1) rk[k >> 1] |= r << 16;
            else rk[k >> 1] = r;
        }
        __syncthreads();
        // ---- per digit: exclusive scan over the warps (thread t owns digits 8t .. 8t+7 = one 16-byte word per warp
        // row; two u16 counters per u32 add, no carry: a CTA holds 4096 keys), the CTA's count, and the exclusive
        // scan of the counts over all digits -> start of each digit's run in the sorted tile
        uint4 cnt8;   // this CTA's counts of digits 8 tid .. 8 tid + 7 (u16 x 8)
        {
            uint4 run = make_uint4(0, 0, 0, 0);
#pragma unroll
            for (int w2 = 0; w2 < kWarps; w2++) {
                uint4* p = reinterpret_cast<uint4*>(&sm.whist[w2][0]) + tid;
                const uint4 v = *p;
                *p = run;
                run.x += v.x; run.y += v.y; run.z += v.z; run.w += v.w;
            }
            cnt8 = run;
            const uint32_t c[8] = {run.x & 0xffffu, run.x >> 16, run.y & 0xffffu, run.y >> 16,
                                   run.z & 0xffffu, run.z >> 16, run.w & 0xffffu, run.w >> 16};
            uint32_t sum = 0;
#pragma unroll
            for (int j = 0; j < 8; j++) sum += c[j];
            uint32_t incl = sum;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const uint32_t t = __shfl_up_sync(0xffffffffu, incl, o);
                if (lane >= o) incl += t;
            }
            if (lane == 31) sm.scan_tmp[warp] = incl;
            __syncthreads();
            uint32_t s0 = incl - sum;
            for (int k = 0; k < warp; k++) s0 += sm.scan_tmp[k];
            uint32_t st[8];
#pragma unroll
            for (int j = 0; j < 8; j++) { st[j] = s0; s0 += c[j]; }
            reinterpret_cast<uint4*>(sm.tile_start)[tid] =
                make_uint4(st[0] | (st[1] << 16), st[2] | (st[3] << 16), st[4] | (st[5] << 16), st[6] | (st[7] << 16));
        }
        __syncthreads();
        // ---- values, then scatter keys and values into digit order
        {
            uint32_t val[kKpt];
#pragma unroll
            for (int k = 0; k < kKpt; k++) {
                const uint32_t s = wslot + k * 32;
                val[k] = s < valid ? (synth_vals ? tile_base + s : vals_in[tile_base + s]) : 0u;
            }
#pragma unroll
            for (int k = 0; k < kKpt; k++) {
                const uint32_t d = (key[k] >> shift) & dmask;
                const uint32_t r = (k & 1) ? (rk[k >> 1] >> 16) : (rk[k >> 1] & 0xffffu);
                const uint32_t pos = (uint32_t)sm.tile_start[d] + (uint32_t)wh[d] + r;
                sm.exch_k[pos] = key[k];
                sm.exch_v[pos] = val[k];
            }
        }
        // ---- counts of digits 8 tid .. 8 tid + 7 go to their owner: one 16-byte asynchronous DSMEM store, issued AFTER
        // the CTA's barrier — every thread of the CTA is done reading whist and writing the exchange buffer.  An owner
        // sends global bases only once the counts of every CTA have arrived, so gpos[] (which aliases whist) is never
        // written under a reader.
        constexpr uint32_t kOwnedPerCta = kBins / CL;
        const uint32_t cnt_owner = (8u * tid) / kOwnedPerCta, cnt_x = 8u * tid - cnt_owner * kOwnedPerCta;
        if (CL == 1) *reinterpret_cast<uint4*>(&sm.rcnt[cnt_x]) = cnt8;
        __syncthreads();
        if (CL > 1) cl_st_async_v4(cl_map(&sm.rcnt[rank * kOwnedPerCta + cnt_x], cnt_owner), cnt8, cl_map(&sm.bar_cnt, cnt_owner));
        const uint32_t parity = it & 1u;
        if (CL > 1) {
            gs_mbar_wait(&sm.bar_cnt, parity);
            if (tid == 0) gs_mbar_expect_tx(&sm.bar_cnt, kCntBytes);   // next tile's phase
        }
        // ---- the next super-tile's ticket travels with the global bases
        if (rank == 0 && tid == 0) {
            const uint32_t t = atomicAdd(ticket, 1u);
            if (CL > 1) {
#pragma unroll
                for (int r = 0; r < CL; r++) cl_st_async_u32(cl_map(&sm.tile_id[(it + 1) & 1u], r), t, cl_map(&sm.bar_gpos, r));
            } else sm.tile_id[(it + 1) & 1u] = t;
        }
        // ---- owner of digits own .. own+kOwn-1: counts of the CTAs, the super-tile's totals, look-back, global bases
        {
            uint32_t c[CL][kOwn], total[kOwn];
#pragma unroll
            for (int j = 0; j < kOwn; j++) total[j] = 0;
#pragma unroll
            for (int r = 0; r < CL; r++)
#pragma unroll
                for (int j = 0; j < kOwn; j++) {
                    c[r][j] = sm.rcnt[r * kOwnedPerCta + tid * kOwn + j];
                    total[j] += c[r][j];
                }
            uint64_t* my = lb + (size_t)super * kBins;
            uint32_t excl[kOwn];
#pragma unroll
            for (int j = 0; j < kOwn; j++) excl[j] = 0;
            if (super == 0) {
#pragma unroll
                for (int j = 0; j < kOwn; j++) gs_st_status(my + j, epoch, GS_LOOKBACK_FLAG_INCL | total[j]);
            } else {
#pragma unroll
                for (int j = 0; j < kOwn; j++) gs_st_status(my + j, epoch, GS_LOOKBACK_FLAG_AGG | total[j]);
                if (kOwn == 1) {
                    // one digit per thread (cluster of 8): kLb predecessors per round trip, consumed in order up to the
                    // first one that carries an inclusive prefix
                    constexpr int kLb = 8;
                    int64_t p = (int64_t)super - 1;
                    bool done = false;
                    while (!done) {
                        uint64_t v[kLb];
#pragma unroll
                        for (int q = 0; q < kLb; q++)
                            v[q] = (p - q >= 0) ? gs_ld_status(lb + (size_t)(p - q) * kBins) : (((uint64_t)epoch << 32) | GS_LOOKBACK_FLAG_INCL);
                        int used = 0;
#pragma unroll
                        for (int q = 0; q < kLb; q++) {
                            if (!done && used == q) {
                                const uint32_t fl = gs_status_flag(v[q], epoch);
                                if (fl != 0u) {
                                    excl[0] += (uint32_t)v[q] & GS_LOOKBACK_VALUE_MASK;
                                    used = q + 1;
                                    done = fl == 2u;
                                }
                            }
                        }
                        p -= used;
                    }
                } else {
                    // several digits per thread (smaller clusters): a predecessor is consumed once ALL still-pending
                    // digits find it published
                    constexpr int kLb = kOwn == 2 ? 2 : 1;   // predecessors per round trip
                    int64_t p = (int64_t)super - 1;
                    uint32_t pending = (1u << kOwn) - 1u;   // digits whose inclusive prefix has not been met yet
                    while (pending) {
                        uint64_t v[kLb][kOwn];
#pragma unroll
                        for (int q = 0; q < kLb; q++)
#pragma unroll
                            for (int j = 0; j < kOwn; j++)
                                v[q][j] = (p - q >= 0) ? gs_ld_status(lb + (size_t)(p - q) * kBins + j)
                                                       : (((uint64_t)epoch << 32) | GS_LOOKBACK_FLAG_INCL);
                        int used = 0;
#pragma unroll
                        for (int q = 0; q < kLb; q++) {
                            if (pending && used == q) {
                                bool ready = true;
#pragma unroll
                                for (int j = 0; j < kOwn; j++)
                                    if (((pending >> j) & 1u) && gs_status_flag(v[q][j], epoch) == 0u) ready = false;
                                if (ready) {
#pragma unroll
                                    for (int j = 0; j < kOwn; j++) {
                                        if ((pending >> j) & 1u) {
                                            excl[j] += (uint32_t)v[q][j] & GS_LOOKBACK_VALUE_MASK;
                                            if (gs_status_flag(v[q][j], epoch) == 2u) pending &= ~(1u << j);
                                        }
                                    }
                                    used = q + 1;
                                }
                            }
                        }
                        p -= used;
                    }
                }
#pragma unroll
                for (int j = 0; j < kOwn; j++) gs_st_status(my + j, epoch, GS_LOOKBACK_FLAG_INCL | (excl[j] + total[j]));
            }
#pragma unroll
            for (int j = 0; j < kOwn; j++) {
                uint32_t g = gbase[j] + excl[j];
#pragma unroll
                for (int r = 0; r < CL; r++) {
                    if (CL > 1) cl_st_async_u32(cl_map(&sm.gpos[own + j], r), g, cl_map(&sm.bar_gpos, r));
                    else sm.gpos[own + j] = g;
                    g += c[r][j];
                }
            }
        }
        // ---- gpos of every digit (and the next ticket) has arrived from the owners
        if (CL > 1) {
            gs_mbar_wait(&sm.bar_gpos, parity);
            if (tid == 0) gs_mbar_expect_tx(&sm.bar_gpos, kGposBytes);   // next tile's phase
        } else __syncthreads();
        // ---- write out: consecutive positions of one digit are consecutive addresses
#pragma unroll
        for (int k = 0; k < kKpt; k++) {
            const uint32_t p = k * kThreads + tid;
            if (p < valid) {
                const uint32_t kk = sm.exch_k[p];
                const uint32_t d = (kk >> shift) & dmask;
                const uint32_t g = sm.gpos[d] + p - (uint32_t)sm.tile_start[d];
                keys_out[g] = kk;
                vals_out[g] = sm.exch_v[p];
            }
        }
        __syncthreads();
    }
}

// co-resident clusters of the pass kernel, per device and per cluster size (index log2(CL))
struct DevInfo { int clusters[4] = {0, 0, 0, 0}; };
std::mutex g_mu;
DevInfo g_dev[64];
std::atomic<int> g_cluster{kMaxCluster};
std::atomic<int> g_claim{0};

using PassKernel = void (*)(uint32_t*, uint32_t*, uint32_t*, uint32_t*, const uint32_t*, uint32_t, const uint32_t*, uint32_t, uint32_t,
                            uint64_t*, uint32_t, uint32_t*, uint32_t*, uint32_t);

template <int CL>
PassKernel pick_kernel(uint32_t bits, bool claim) {
    if (bits <= 2) return k_sort_wide_pass<2, CL, false>;
    if (bits <= 5) return k_sort_wide_pass<5, CL, false>;
    if (bits <= 8) return k_sort_wide_pass<8, CL, false>;
    if (bits <= 10) return claim ? k_sort_wide_pass<10, CL, true> : k_sort_wide_pass<10, CL, false>;
    return claim ? k_sort_wide_pass<11, CL, true> : k_sort_wide_pass<11, CL, false>;
}
PassKernel pick_kernel(int cl, uint32_t bits, bool claim) {
    return cl == 8 ? pick_kernel<8>(bits, claim) : cl == 4 ? pick_kernel<4>(bits, claim)
         : cl == 2 ? pick_kernel<2>(bits, claim) : pick_kernel<1>(bits, claim);
}

void fill_config(cudaLaunchConfig_t* cfg, cudaLaunchAttribute* at, int cl, uint32_t clusters, cudaStream_t st) {
    *cfg = cudaLaunchConfig_t{};
    cfg->gridDim = dim3(clusters * cl, 1, 1);
    cfg->blockDim = dim3(kThreads, 1, 1);
    cfg->dynamicSmemBytes = sizeof(PassSmem);
    cfg->stream = st;
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = cl; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
    cfg->attrs = at;
    cfg->numAttrs = cl > 1 ? 1 : 0;
}

cudaError_t setup_device(int cl, int* clusters) {
    int nc_min = 0;
    for (uint32_t variant : {2u, 5u, 8u, 10u, 11u, 110u, 111u}) {   // (1xx: the claim instantiations)
        PassKernel kern = pick_kernel(cl, variant % 100u, variant >= 100u);
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(PassSmem));
        if (e != cudaSuccess) return e;
        int nc = 0;
        if (cl > 1) {
            cudaLaunchConfig_t cfg;
            cudaLaunchAttribute at[1];
            fill_config(&cfg, at, cl, 1, nullptr);
            e = cudaOccupancyMaxActiveClusters(&nc, kern, &cfg);
            if (e != cudaSuccess) return e;
        } else {
            int bps = 0, sms = 0, dev = 0;
            e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&bps, kern, kThreads, sizeof(PassSmem));
            if (e == cudaSuccess) e = cudaGetDevice(&dev);
            if (e == cudaSuccess) e = cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
            if (e != cudaSuccess) return e;
            nc = bps * sms;
        }
        if (nc < 1) nc = 1;
        if (nc_min == 0 || nc < nc_min) nc_min = nc;
    }
    *clusters = nc_min;
    return cudaSuccess;
}

int cluster_index(int cl) { return cl == 8 ? 3 : cl == 4 ? 2 : cl == 2 ? 1 : 0; }

}  // namespace

// tuning knob: CTAs per cluster (8, 4, 2 or 1 = no clusters).  Process-wide; set it before viewers are created
// (look-back buffers are sized from it).
cudaError_t gs_sort_set_cluster(int cl) {
    if (cl != 1 && cl != 2 && cl != 4 && cl != 8) return cudaErrorInvalidValue;
    g_cluster.store(cl);
    return cudaSuccess;
}
int gs_sort_get_cluster() { return g_cluster.load(); }
// tuning knob: try the collision-free fast path per row of wide (10/11-bit) digits before the ballots (default off)
void gs_sort_set_claim(int on) { g_claim.store(on ? 1 : 0); }
int gs_sort_get_claim() { return g_claim.load(); }
int gs_sort_resident_clusters(int device) {
    std::lock_guard<std::mutex> lock(g_mu);
    return (device >= 0 && device < 64) ? g_dev[device].clusters[cluster_index(g_cluster.load())] : 0;
}

size_t gs_sort_wide_lookback_words(uint32_t n_max, uint32_t key_bits) {
    const size_t super = (size_t)kTile * g_cluster.load();
    size_t supers = ((size_t)n_max + super - 1) / super;
    if (supers < 1) supers = 1;
    return supers * kBins * gs_sort_passes(key_bits);
}

cudaError_t gs_launch_sort_wide(const GsSortWideArgs& a, int num_sms, cudaStream_t st) {
    (void)num_sms;
    if (a.key_bits < 1 || a.key_bits > 32 || !a.result_in_b) return cudaErrorInvalidValue;
    const uint32_t passes = gs_sort_passes(a.key_bits);
    const int cl = g_cluster.load();
    const size_t super = (size_t)kTile * cl;
    size_t supers = ((size_t)a.n_max + super - 1) / super;
    if (supers < 1) supers = 1;
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return e;
    if (dev < 0 || dev >= 64) return cudaErrorInvalidDevice;
    int clusters;
    {
        std::lock_guard<std::mutex> lock(g_mu);
        int& c = g_dev[dev].clusters[cluster_index(cl)];
        if (c == 0) {
            e = setup_device(cl, &c);
            if (e != cudaSuccess) { c = 0; return e; }
        }
        clusters = c;
    }
    if (!a.hist_prefilled) {
        uint32_t grid = 296;
        const uint32_t need = (uint32_t)((a.n_max + kThreads - 1) / kThreads);
        if (grid > need) grid = need < 1 ? 1 : need;
        k_sort_wide_hist<<<grid, kThreads, passes * kBins * 4, st>>>(a.keys_a, a.d_n, a.n_max, a.hist, a.key_bits);
    }
    uint32_t grid_clusters = (uint32_t)clusters;
    if (grid_clusters > supers) grid_clusters = (uint32_t)supers;
    for (uint32_t p = 0; p < passes; p++) {
        // the kernel is instantiated for a few ballot counts; a narrower digit is ranked on the next wider
        // instantiation that still fits below bit 32: the extra bits lie above key_bits, are zero in every real
        // key and set only in the 0xffffffff padding, so real digits are unchanged and padding still sorts last
        const uint32_t shift = GS_SORT_DIGIT_BITS * p;
        const uint32_t bits = a.key_bits - shift < GS_SORT_DIGIT_BITS ? a.key_bits - shift : GS_SORT_DIGIT_BITS;
        PassKernel kern = pick_kernel(cl, bits, g_claim.load() != 0);
        cudaLaunchConfig_t cfg;
        cudaLaunchAttribute at[1];
        fill_config(&cfg, at, cl, grid_clusters, st);
        e = cudaLaunchKernelEx(&cfg, kern, a.keys_a, a.vals_a, a.keys_b, a.vals_b, a.d_n, a.n_max, (const uint32_t*)a.hist, p, a.key_bits,
                               a.lookback + (size_t)p * supers * kBins, a.epoch, a.tickets + p, a.result_in_b,
                               a.vals_identity ? 1u : 0u);
        if (e != cudaSuccess) return e;
    }
    return cudaGetLastError();
}
