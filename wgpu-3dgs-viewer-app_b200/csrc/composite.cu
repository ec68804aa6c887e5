// composite.cu — K3: tile-binned front-to-back splat compositor.
//
// Replaces the fragment + ROP-blend half of renderer.render_with_pass (reference
// src/tab/scene.rs:2302-2314; depth state scene.rs:1972-1978): the reference blends quads back
// to front in hardware; here one CTA owns a 16x16 pixel tile, stages the tile's depth-sorted
// splats in shared memory 256 at a time and blends them FRONT TO BACK
//     C += c·α·T,  T -= α·T,   α = min(0.99, o·exp(-½ dᵀQd)),
// dropping α < 1/255 and power > 0; a pixel has stopped when T < 1/1024, and a warp / the CTA
// leave as soon as all their pixels have stopped.  Each warp owns an 8x4 pixel sub-tile and first
// culls the staged splats against it 32 at a time with a ballot (exact ellipse-vs-rectangle
// footprint test, one splat per lane), so only splats that can reach alpha >= 1/255 inside the
// sub-tile are evaluated; the hitting lanes leave the splats' shared-memory addresses in a per-warp
// hit list that the blend loop walks two splats per trip.  Issue-slot and shared-memory-pipe bound,
// not HBM bound.
//
// Lists are per 32x32-pixel BIN (bin.cu): the four tiles of a bin (its quadrants) are four CTAs that walk the same
// list, and every entry's key says which quadrants its splat can reach.  A ninth PRODUCER warp per CTA streams the
// bin's (key, splat id) entries, keeps the ids that carry this quadrant's bit and feeds them, in order, through a
// shared-memory ring to the eight compositing warps — which therefore gather, convert and cull exactly the splats a
// 16-pixel binning would have given them, while the binning stage and the bin sort handle 26-40 % fewer entries.
#include "common.cuh"

namespace {

constexpr int kConsumers = 256;              // 8 compositing warps: one 8x4 sub-tile each
constexpr int kThreads = kConsumers + 32;    // + the producer warp
constexpr uint32_t kRing = 2048;             // ring of filtered splat ids (8 rounds)
constexpr uint32_t kPrologue = 1024;         // entries filtered cooperatively by the compositing warps at CTA start
constexpr uint32_t kBatch = 512;             // entries examined per producer step (16 per lane)
constexpr int kPlane = kConsumers + 1;        // entries per staged plane: 256 splats + one NULL splat (empty extent masks)
constexpr float kLog2e = 1.4426950408889634f;

__device__ __forceinline__ float ex2_approx(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
__device__ __forceinline__ float rcp_approx(float x) {   // (operands here are conic entries: never denormal, never huge)
    float y;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
// loads from a shared-memory window address (the hit list holds addresses, not indices)
__device__ __forceinline__ float4 lds128(uint32_t a) {
    float4 v;
    asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(a));
    return v;
}
__device__ __forceinline__ float2 lds64(uint32_t a) {
    float2 v;
    asm volatile("ld.shared.v2.f32 {%0, %1}, [%2];" : "=f"(v.x), "=f"(v.y) : "r"(a));
    return v;
}

// The producer warp and the compositing warps run different code between barriers, so the barriers are spelled as PTX named
// barriers with an explicit thread count (what warp-specialised kernels use) rather than as __syncthreads(), whose C++
// contract wants one call site for the whole CTA.  Every warp executes the same sequence of them.
// compute-sanitizer's synccheck reports "divergent thread(s) in block" whenever a block-wide rendezvous is reached from two
// code addresses, correct or not (tools/scratch/bar_pattern.cu is the minimal reproducer: right answer, flagged; the same
// program with the barrier in a non-inlined function: clean).  Building with -DB200GS_SYNCCHECK_BUILD (build.py
// --synccheck, used by tools/sanitize.sh) keeps the two helpers out of line — one barrier instruction at one address for
// both roles — which synccheck accepts; the production build inlines them (the call costs 16 us per frame: 334 -> 350).
#ifdef B200GS_SYNCCHECK_BUILD
#define GS_BAR_FN __noinline__
#else
#define GS_BAR_FN __forceinline__
#endif
__device__ GS_BAR_FN void cta_bar() { asm volatile("bar.sync 1, %0;" ::"n"(kThreads) : "memory"); }
__device__ GS_BAR_FN bool cta_bar_and(bool pred) {
    uint32_t r;
    asm volatile(
        "{\n"
        ".reg .pred p, q;\n"
        "setp.ne.u32 p, %1, 0;\n"
        "bar.red.and.pred q, 1, %2, p;\n"
        "selp.u32 %0, 1, 0, q;\n"
        "}\n"
        : "=r"(r)
        : "r"((uint32_t)pred), "n"(kThreads)
        : "memory");
    return r != 0;
}

template <bool FLAT, bool COUNT>
__global__ void __launch_bounds__(kThreads, 4) k_composite(const uint32_t* __restrict__ tile_keys_a,
                                                           const uint32_t* __restrict__ tile_vals_a,
                                                           const uint32_t* __restrict__ tile_keys_b,
                                                           const uint32_t* __restrict__ tile_vals_b,
                                                           const uint32_t* tile_in_b,
                                                           const uint32_t* __restrict__ ranges,
                                                           const b200gs_splat* __restrict__ splats, uint8_t* out,
                                                           size_t pitch, uint32_t W, uint32_t H, uint32_t bins_x,
                                                           uint32_t n_bins, float bg0, float bg1, float bg2, float bg3,
                                                           unsigned long long* evals) {
    // 64 bytes per staged splat as four float4 planes [j][splat] (a lane reading its own splat in the cull
    // loop touches consecutive 16-byte words: no bank conflicts): {mx, my, a', b'} {c', opacity, red, green}
    // {blue, extent masks, tau', -} {cx, hx, cy, hy}; conic pre-scaled so that power is in log2 units.  The extent
    // square clipped to the viewport is stored twice: as centre / half-size (exact: half-integers) for the cull, and,
    // clipped to this tile, as two 16-bit masks (columns x0..x1 | rows y0..y1 << 16) for the blend loop, where "is my
    // pixel inside the square" is then ONE logic instruction against a per-lane constant instead of two subtractions and
    // two compares on a fourth 16-byte load.  Double-buffered: the next round's splats are in flight while this round is
    // blended.
    __shared__ float4 sS[2][4 * kPlane];
    __shared__ __align__(8) uint32_t s_hit[kConsumers / 32][34];   // per warp: shared-memory byte offsets of the splats that hit its sub-tile
    __shared__ uint32_t s_ring[kRing];   // splat ids of this quadrant, in list order; position p lives in slot p % kRing
    __shared__ uint32_t s_wcnt[8];       // prologue: ids kept by each compositing warp
    __shared__ uint32_t s_prod[2];       // ids produced so far; bit 31: the list is exhausted (the count is final).  Written before
                                         // barrier number i into word i & 1 and read after it, so no word is ever written while read
    constexpr uint32_t kDone = 0x80000000u;

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const bool in_b = *tile_in_b != 0;
    const uint32_t* __restrict__ tile_keys = in_b ? tile_keys_b : tile_keys_a;
    const uint32_t* __restrict__ tile_vals = in_b ? tile_vals_b : tile_vals_a;
    // longest lists first; the 4 quadrants of a bin are neighbours in the launch order, so that they run at the same time and
    // share the bin's entries and the splats that straddle quadrants in L2 (measured: quadrant-major order 384 us vs 334 us)
    const uint32_t bin = ranges[2 * n_bins + (blockIdx.x >> 2)];
    const uint32_t quad = blockIdx.x & 3u;
    const uint32_t tx = 2u * (bin % bins_x) + (quad & 1u), ty = 2u * (bin / bins_x) + (quad >> 1);
    if (tx * GS_TILE >= W || ty * GS_TILE >= H) return;   // quadrant outside the viewport (CTA-uniform)
    const uint32_t start = ranges[bin], end = ranges[n_bins + bin];
    const uint32_t qbit = 1u << (GS_QMASK_SHIFT + quad);
    const uint32_t lane_lt = (1u << lane) - 1u;

    // Every barrier of this kernel is CTA-wide and executed by all nine warps in the same order (A, B, C, then one per
    // round), so shared memory needs no flags, fences or polling: what a warp wrote before a barrier is there after it.
    if (warp == kConsumers / 32) {
        // ------------------------------------------------------------ producer warp
        // Streams the bin's entries beyond the prologue kBatch at a time (8 independent coalesced loads of key and id
        // per lane), keeps the ids whose key carries this quadrant's bit and appends them to the ring in list (= depth)
        // order, a few rounds ahead of the compositing warps.
        uint32_t next = min(end, start + kPrologue);   // next entry to examine
        constexpr int kPer = kBatch / 32;
        uint32_t k[kPer], v[kPer];
        bool loaded = false;                           // k / v hold the batch at `next` (its loads may still be in flight)
        auto prefetch = [&]() {
            if (loaded || next >= end) return;
#pragma unroll
            for (int j = 0; j < kPer; j++) {
                const uint32_t e = next + 32u * j + lane;
                const bool ok = e < end;
                k[j] = ok ? __ldg(tile_keys + e) : 0u;
                v[j] = ok ? __ldg(tile_vals + e) : 0u;
            }
            loaded = true;
        };
        uint32_t w = 0;                    // ids in the ring so far
        auto append = [&]() {              // filter the loaded batch into the ring
#pragma unroll
            for (int j = 0; j < kPer; j++) {
                const bool keep = (k[j] & qbit) != 0u;
                const uint32_t bal = __ballot_sync(0xffffffffu, keep);
                if (keep) s_ring[(w + __popc(bal & lane_lt)) & (kRing - 1u)] = v[j];
                w += __popc(bal);
            }
            next += kBatch;
            loaded = false;
        };
        // One phase = the time between two barriers.  A batch whose loads were issued in the previous phase is filtered for
        // free (if the ring has room and the compositing warps are less than `ahead` ids ahead served); only when the ids
        // the NEXT fetch needs are still missing does the warp wait on memory inside a phase; then it issues the loads of
        // the following batch and goes to the barrier, so that it is never what the other eight warps wait for.
        uint32_t barrier_no = 0;           // barriers passed since A: B is number 0, C number 1, the one ending round r number r + 2
        auto phase = [&](uint32_t need, uint32_t ahead, uint32_t cap) {
            if (loaded && w < ahead && w + kBatch <= cap) append();
            while (next < end && w < need) { prefetch(); append(); }   // (need <= cap - kBatch by construction)
            if (w < ahead || ahead == 0u) prefetch();
            if (lane == 0) s_prod[barrier_no & 1u] = next < end ? w : (w | kDone);
            barrier_no++;
        };
        prefetch();                        // in flight across barrier A
        cta_bar();                         // A: the prologue counts are in s_wcnt
#pragma unroll
        for (int i = 0; i < 8; i++) w += s_wcnt[i];
        // (before B and C only what the next fetch needs: the compositing warps are waiting to start)
        phase(1u * kConsumers, 0u, kRing);   // round 0 is fetched right after B
        cta_bar();                           // B
        phase(2u * kConsumers, 0u, kRing);   // round 1 is fetched right after C
        cta_bar();                           // C
        for (uint32_t round = 0;; round++) {
            if (next >= end && w <= round * kConsumers) break;   // the compositing warps leave at the top of this round
            // the barrier that ends round r is followed by the fetch of round r + 2; everything the compositing warps
            // read before the PREVIOUS barrier (ids of rounds <= r) may be overwritten
            phase((round + 3u) * kConsumers, (round + 5u) * kConsumers, (round + 1u) * kConsumers + kRing);
            if (cta_bar_and(true)) break;   // every pixel of the tile has stopped
        }
        return;
    }

    // ---------------------------------------------------------------- compositing warps
    // prologue: the first kPrologue entries of the list are filtered by these eight warps together (one memory round
    // trip for all of them, warp w takes entries [128 w, 128 w + 128)), which for most tiles is most of what they will
    // ever consume; the producer warp continues from there
    {
        constexpr int kPer = kPrologue / kConsumers;
        uint32_t k[kPer], v[kPer], bal[kPer], cnt = 0;
        const uint32_t e0 = start + 32u * kPer * warp + lane;
#pragma unroll
        for (int j = 0; j < kPer; j++) {
            const uint32_t e = e0 + 32u * j;
            const bool ok = e < end;
            k[j] = ok ? __ldg(tile_keys + e) : 0u;
            v[j] = ok ? __ldg(tile_vals + e) : 0u;
        }
#pragma unroll
        for (int j = 0; j < kPer; j++) {
            bal[j] = __ballot_sync(0xffffffffu, (k[j] & qbit) != 0u);
            cnt += __popc(bal[j]);
        }
        if (lane == 0) s_wcnt[warp] = cnt;
        if (tid < 2) sS[tid][2 * kPlane + kConsumers] = make_float4(0.0f, 0.0f, 0.0f, 0.0f);   // the NULL splat: empty extent masks
        cta_bar();                         // A
        uint32_t pos = 0;
#pragma unroll
        for (int i = 0; i < 8; i++) pos += i < warp ? s_wcnt[i] : 0u;
#pragma unroll
        for (int j = 0; j < kPer; j++) {
            if (k[j] & qbit) s_ring[pos + __popc(bal[j] & lane_lt)] = v[j];   // (< kPrologue <= kRing: no wrap)
            pos += __popc(bal[j]);
        }
        cta_bar();                         // B: the ids of round 0 are in the ring (or the list is exhausted)
    }

    // warp -> 8x4 sub-tile, lane -> pixel
    const int wx0 = (int)(tx * GS_TILE) + (warp & 1) * 8, wy0 = (int)(ty * GS_TILE) + (warp >> 1) * 4;
    const float fwx0 = (float)wx0, fwx1 = (float)(wx0 + 7), fwy0 = (float)wy0, fwy1 = (float)(wy0 + 3);
    const int px = wx0 + (lane & 7), py = wy0 + (lane >> 3);
    const float fpx = (float)px, fpy = (float)py;
    const bool inside = px < (int)W && py < (int)H;
    const int tpx0 = (int)(tx * GS_TILE), tpy0 = (int)(ty * GS_TILE);
    const uint32_t sel = (1u << (px - tpx0)) | (0x10000u << (py - tpy0));   // this pixel's column bit | row bit << 16

    // A pixel has stopped when T < 1/1024; pixels outside the viewport start stopped (T = 0).  "Stopped" is always read off
    // T itself: no separate flag to carry through the blend loop.  Stopping is a property of the WARP and the CTA (they leave
    // when all their pixels have stopped): a stopped pixel whose warp is still running keeps blending the splats the warp
    // evaluates — contributions below 1/1024 that the back-to-front reference includes anyway — which takes the "has this
    // pixel stopped" test out of the blend loop.
    float T = inside ? 1.0f : 0.0f, Cr = 0.0f, Cg = 0.0f, Cb = 0.0f;
    unsigned long long my_evals = 0;
    const float Wf = (float)W, Hf = (float)H;

    // software pipeline of the staging: the splat records of the next round are in registers (fetched one round
    // ahead through the ring), converted into the other shared-memory buffer at the end of the current round, so a
    // round costs ONE barrier (which also carries the early-exit vote)
    uint4 q0 = make_uint4(0, 0, 0, 0), q1 = make_uint4(0, 0, 0, 0);
    auto stage = [&](float4* dst) {
        const float mx = __uint_as_float(q0.x), my = __uint_as_float(q0.y);
        const float r = (float)(q0.z & 0xffffu);
        const float op = __half2float(__ushort_as_half((unsigned short)(q0.z >> 16)));
        const float cr = __half2float(__ushort_as_half((unsigned short)(q0.w & 0xffffu)));
        const float cg = __half2float(__ushort_as_half((unsigned short)(q0.w >> 16)));
        const float cb = __half2float(__ushort_as_half((unsigned short)(q1.w & 0xffffu)));
        const float ca = __uint_as_float(q1.x), cbq = __uint_as_float(q1.y), cc = __uint_as_float(q1.z);
        // same bounds expression as bin.cu / the oracle (exact in float)
        float fx0 = ceilf(mx - r), fx1 = floorf(mx + r), fy0 = ceilf(my - r), fy1 = floorf(my + r);
        if (fx0 < 0.0f) fx0 = 0.0f;
        if (fy0 < 0.0f) fy0 = 0.0f;
        if (fx1 > Wf - 1.0f) fx1 = Wf - 1.0f;
        if (fy1 > Hf - 1.0f) fy1 = Hf - 1.0f;
        // the square clipped to this tile, as column / row bit masks (empty masks if it misses the tile)
        const int x0 = max((int)fx0 - tpx0, 0), x1 = min((int)fx1 - tpx0, GS_TILE - 1);
        const int y0 = max((int)fy0 - tpy0, 0), y1 = min((int)fy1 - tpy0, GS_TILE - 1);
        const uint32_t xm = x0 <= x1 ? (2u << x1) - (1u << x0) : 0u, ym = y0 <= y1 ? (2u << y1) - (1u << y0) : 0u;
        dst[tid] = make_float4(mx, my, -0.5f * kLog2e * ca, -kLog2e * cbq);
        dst[kPlane + tid] = make_float4(-0.5f * kLog2e * cc, op, cr, cg);
        dst[2 * kPlane + tid] = make_float4(cb, __uint_as_float(xm | (ym << 16)), 0.5f * kLog2e * gs_footprint_tau(op, FLAT), 0.0f);
        dst[3 * kPlane + tid] = make_float4(0.5f * (fx0 + fx1), 0.5f * (fx1 - fx0), 0.5f * (fy0 + fy1), 0.5f * (fy1 - fy0));
    };
    // splat record of filtered position `pos` into the registers; false if the list ended before it.  (Called right
    // after a barrier before which the producer made sure the position exists unless the list is exhausted.)
    auto fetch = [&](uint32_t pos, uint32_t barrier_no) -> bool {
        if (pos >= (s_prod[barrier_no & 1u] & ~kDone)) return false;
        const uint4* sp = reinterpret_cast<const uint4*>(splats + s_ring[pos & (kRing - 1u)]);
        q0 = __ldg(sp); q1 = __ldg(sp + 1);
        return true;
    };
    if (fetch((uint32_t)tid, 0u)) stage(sS[0]);                      // round 0 -> buffer 0
    cta_bar();                                                       // C: the ids of round 1 are in the ring, too
    bool have_next = fetch((uint32_t)(kConsumers + tid), 1u);        // round 1 -> registers

    uint32_t buf = 0;
    for (uint32_t round = 0;; round++, buf ^= 1u) {
        // (before the barrier just passed the producer had delivered this whole round and the next, or the whole list)
        const uint32_t produced = s_prod[(round + 1u) & 1u] & ~kDone, first = round * kConsumers;   // (barrier number round + 1 was the last)
        if (produced <= first) break;
        const uint32_t cnt = min((uint32_t)kConsumers, produced - first);
        if (COUNT && tid == 0) atomicAdd(evals + 1, (unsigned long long)cnt);  // entries staged before the tile finished
        const float4* sSb = sS[buf];
        const uint32_t plane0 = gs_smem_u32(sSb);   // shared-memory address of this round's first plane

        if (!__all_sync(0xffffffffu, T < GS_T_EPS)) {
            for (uint32_t g = 0; g < cnt; g += 32) {
                const uint32_t s = g + lane;
                // splat s against this warp's 8x4 sub-tile: extent-square overlap, then the exact footprint test (can any
                // pixel of the overlap reach alpha >= 1/255?: the minimum of the quadratic over the overlap rectangle is at
                // the centre if it is inside, else on the edge nearest to it).  Straight-line code: lanes past the end of the
                // list and splats that miss the square compute on whatever the slot holds and are masked at the end.
                bool ov;
                {
                    const float4 C = sSb[3 * kPlane + s];
                    const float x0 = fmaxf(C.x - C.y, fwx0), x1 = fminf(C.x + C.y, fwx1);
                    const float y0 = fmaxf(C.z - C.w, fwy0), y1 = fminf(C.z + C.w, fwy1);
                    const float4 A = sSb[s];
                    const float tau = sSb[2 * kPlane + s].z;
                    const float pa = -A.z, pb = -0.5f * A.w, pc = -sSb[kPlane + s].x;  // 0.5*log2e * (a, b, c)
                    const float dx0 = x0 - A.x, dx1 = x1 - A.x, dy0 = y0 - A.y, dy1 = y1 - A.y;
                    const bool inx = dx0 <= 0.0f && dx1 >= 0.0f, iny = dy0 <= 0.0f && dy1 >= 0.0f;
                    // the vertical edge nearest the centre, minimised over y; the horizontal edge nearest the centre, over x
                    const float dxe = dx0 > 0.0f ? dx0 : dx1;
                    const float dye = fminf(dy1, fmaxf(dy0, -pb * dxe * rcp_approx(pc)));
                    const float qe = pa * dxe * dxe + 2.0f * pb * dxe * dye + pc * dye * dye;
                    const float dyf = dy0 > 0.0f ? dy0 : dy1;
                    const float dxf = fminf(dx1, fmaxf(dx0, -pb * dyf * rcp_approx(pa)));
                    const float qf = pa * dxf * dxf + 2.0f * pb * dxf * dyf + pc * dyf * dyf;
                    const float best = fminf(inx ? (iny ? 0.0f : 3.0e38f) : qe, iny ? 3.0e38f : qf);
                    ov = s < cnt && x0 <= x1 && y0 <= y1 && best <= tau;
                }
                // The lanes whose splat hits write its shared-memory offset into the warp's hit list, in order (an odd count is
                // padded with the NULL splat, whose extent masks are empty); the blend loop then walks the list two splats per
                // trip — their alpha evaluations are independent (ILP), the blends are applied in order — and costs one 8-byte
                // load per pair for "which splats" instead of a find-first-set chain on the ballot.
                const uint32_t m = __ballot_sync(0xffffffffu, ov);
                const int n_hit = __popc(m);
                if (ov) s_hit[warp][__popc(m & lane_lt)] = plane0 + s * 16u;
                if (lane == 0 && (n_hit & 1)) s_hit[warp][n_hit] = plane0 + (uint32_t)kConsumers * 16u;
                __syncwarp();
                for (int i = 0; i < n_hit; i += 2) {
                    const uint2 h = *reinterpret_cast<const uint2*>(&s_hit[warp][i]);
                    const float4 Aa = lds128(h.x), Ba = lds128(h.x + kPlane * 16);
                    const float4 Ab = lds128(h.y), Bb = lds128(h.y + kPlane * 16);
                    const float2 Ca = lds64(h.x + 2 * kPlane * 16);   // {blue, extent masks}
                    const float2 Cbb = lds64(h.y + 2 * kPlane * 16);
                    const float dxa = fpx - Aa.x, dya = fpy - Aa.y, dxb = fpx - Ab.x, dyb = fpy - Ab.y;
                    const bool ina = (__float_as_uint(Ca.y) & sel) == sel;
                    const bool inb = (__float_as_uint(Cbb.y) & sel) == sel;
                    const float pa2 = __fmaf_rn(__fmaf_rn(Aa.w, dya, Aa.z * dxa), dxa, (Ba.x * dya) * dya);
                    const float pb2 = __fmaf_rn(__fmaf_rn(Ab.w, dyb, Ab.z * dxb), dxb, (Bb.x * dyb) * dyb);
                    float ala, alb;
                    if (FLAT) {
                        ala = (pa2 >= -0.5f * GS_FLAT_D2 * kLog2e) ? fminf(GS_ALPHA_MAX, Ba.y) : 0.0f;
                        alb = (pb2 >= -0.5f * GS_FLAT_D2 * kLog2e) ? fminf(GS_ALPHA_MAX, Bb.y) : 0.0f;
                    } else {
                        ala = fminf(GS_ALPHA_MAX, Ba.y * ex2_approx(pa2));
                        alb = fminf(GS_ALPHA_MAX, Bb.y * ex2_approx(pb2));
                    }
                    if (COUNT) my_evals += (ina && T >= GS_T_EPS) ? 1ull : 0ull;
                    if (ina && pa2 <= 0.0f && ala >= GS_ALPHA_MIN) {
                        const float w = ala * T;
                        Cr = __fmaf_rn(Ba.z, w, Cr);
                        Cg = __fmaf_rn(Ba.w, w, Cg);
                        Cb = __fmaf_rn(Ca.x, w, Cb);
                        T -= w;
                    }
                    if (COUNT) my_evals += (inb && T >= GS_T_EPS) ? 1ull : 0ull;
                    if (inb && pb2 <= 0.0f && alb >= GS_ALPHA_MIN) {
                        const float w = alb * T;
                        Cr = __fmaf_rn(Bb.z, w, Cr);
                        Cg = __fmaf_rn(Bb.w, w, Cg);
                        Cb = __fmaf_rn(Cbb.x, w, Cb);
                        T -= w;
                    }
                }
                __syncwarp();   // every lane is done reading the hit list before the next group overwrites it
                if (__all_sync(0xffffffffu, T < GS_T_EPS)) break;
            }
        }
        // next round: registers -> the other buffer, then (behind the barrier) fetch the round after it
        if (have_next) stage(sS[buf ^ 1u]);
        if (cta_bar_and(T < GS_T_EPS)) break;
        have_next = fetch((round + 2u) * kConsumers + tid, round + 2u);
    }

    if (inside) {
        const float r = Cr + bg0 * T, g = Cg + bg1 * T, b = Cb + bg2 * T, a = (1.0f - T) + bg3 * T;
        auto q = [](float v) -> uint32_t {
            v = v < 0.0f ? 0.0f : (v > 1.0f ? 1.0f : v);
            return (uint32_t)(v * 255.0f + 0.5f);
        };
        const uint32_t rgba = q(r) | (q(g) << 8) | (q(b) << 16) | (q(a) << 24);
        *reinterpret_cast<uint32_t*>(out + (size_t)py * pitch + (size_t)px * 4) = rgba;
    }
    if (COUNT) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) my_evals += __shfl_xor_sync(0xffffffffu, my_evals, o);
        if (lane == 0 && my_evals) atomicAdd(evals, my_evals);
    }
}

// Hit query (row N3): the ordered list of splats contributing to ONE pixel — a single-warp walk of
// the pixel's tile list with the compositor's own alpha rule.  Replaces the query_results buffer
// the reference renderer appends to (src/tab/scene.rs:635-657).
__global__ void __launch_bounds__(32) k_query_hits(const uint32_t* __restrict__ tile_keys_a, const uint32_t* __restrict__ tile_vals_a,
                                                   const uint32_t* __restrict__ tile_keys_b, const uint32_t* __restrict__ tile_vals_b,
                                                   const uint32_t* tile_in_b,
                                                   const uint32_t* __restrict__ ranges,
                                                   const b200gs_splat* __restrict__ splats, uint32_t W, uint32_t H,
                                                   uint32_t bins_x, uint32_t n_bins, uint32_t px, uint32_t py,
                                                   uint32_t flat, uint2* out, uint32_t cap, uint32_t* count) {
    const bool in_b = *tile_in_b != 0;
    const uint32_t* __restrict__ tile_keys = in_b ? tile_keys_b : tile_keys_a;
    const uint32_t* __restrict__ tile_vals = in_b ? tile_vals_b : tile_vals_a;
    const uint32_t bin = (py / GS_BIN) * bins_x + px / GS_BIN;
    const uint32_t qbit = 1u << (GS_QMASK_SHIFT + (((py / GS_TILE) & 1u) * 2u + ((px / GS_TILE) & 1u)));   // the pixel's quadrant of the bin
    const uint32_t start = ranges[bin], end = ranges[n_bins + bin];
    const int lane = threadIdx.x;
    const float fpx = (float)px, fpy = (float)py, Wf = (float)W, Hf = (float)H;
    uint32_t n_out = 0;
    for (uint32_t base = start; base < end; base += 32) {
        const uint32_t e = base + lane;
        bool ok = false;
        uint32_t id = 0;
        float al = 0.0f;
        if (e < end && (tile_keys[e] & qbit)) {
            id = tile_vals[e];
            const uint4* sp = reinterpret_cast<const uint4*>(splats + id);
            const uint4 q0 = __ldg(sp), q1 = __ldg(sp + 1);
            const float mx = __uint_as_float(q0.x), my = __uint_as_float(q0.y);
            const float r = (float)(q0.z & 0xffffu);
            const float op = __half2float(__ushort_as_half((unsigned short)(q0.z >> 16)));
            float fx0 = ceilf(mx - r), fx1 = floorf(mx + r), fy0 = ceilf(my - r), fy1 = floorf(my + r);
            fx0 = fmaxf(fx0, 0.0f); fy0 = fmaxf(fy0, 0.0f); fx1 = fminf(fx1, Wf - 1.0f); fy1 = fminf(fy1, Hf - 1.0f);
            const float dx = fpx - mx, dy = fpy - my;
            const float power = -0.5f * (__uint_as_float(q1.x) * dx * dx + __uint_as_float(q1.z) * dy * dy) - __uint_as_float(q1.y) * dx * dy;
            if (flat) al = power >= -0.5f * GS_FLAT_D2 ? fminf(GS_ALPHA_MAX, op) : 0.0f;
            else al = fminf(GS_ALPHA_MAX, op * __expf(power));
            ok = fpx >= fx0 && fpx <= fx1 && fpy >= fy0 && fpy <= fy1 && power <= 0.0f && al >= GS_ALPHA_MIN;
        }
        const uint32_t bal = __ballot_sync(0xffffffffu, ok);
        const uint32_t o = n_out + __popc(bal & ((1u << lane) - 1u));
        if (ok && o < cap) out[o] = make_uint2(id, __float_as_uint(al));
        n_out += __popc(bal);
    }
    if (lane == 0) *count = n_out;
}

}  // namespace

cudaError_t gs_launch_query_hits(const GsCompositeArgs& a, const GsFrame& f, uint32_t px, uint32_t py, uint2* out,
                                 uint32_t cap, uint32_t* count, cudaStream_t st) {
    k_query_hits<<<1, 32, 0, st>>>(a.tile_keys, a.tile_vals, a.tile_keys_b, a.tile_vals_b, a.tile_in_b, a.ranges, a.splats,
                                   (uint32_t)f.W, (uint32_t)f.H, f.bins_x, f.bins_x * f.bins_y, px, py,
                                   f.display_mode != B200GS_DISPLAY_SPLAT ? 1u : 0u, out, cap, count);
    return cudaGetLastError();
}

cudaError_t gs_launch_composite(const GsCompositeArgs& a, const GsFrame& f, cudaStream_t st) {
    const uint32_t n_bins = f.bins_x * f.bins_y;
    const bool flat = f.display_mode != B200GS_DISPLAY_SPLAT;
    const uint32_t W = (uint32_t)f.W, H = (uint32_t)f.H;
    // one CTA per compositor tile: 4 per bin (quadrants outside the viewport exit at once)
#define GS_LAUNCH_COMPOSITE(FLAT, COUNT)                                                                              \
    k_composite<FLAT, COUNT><<<4u * n_bins, kThreads, 0, st>>>(a.tile_keys, a.tile_vals, a.tile_keys_b, a.tile_vals_b, a.tile_in_b,  \
                                                               a.ranges, a.splats, a.out, a.pitch, W, H, f.bins_x, n_bins,  \
                                                               f.bg[0], f.bg[1], f.bg[2], f.bg[3], a.evals)
    if (flat) {
        if (a.evals) GS_LAUNCH_COMPOSITE(true, true);
        else GS_LAUNCH_COMPOSITE(true, false);
    } else {
        if (a.evals) GS_LAUNCH_COMPOSITE(false, true);
        else GS_LAUNCH_COMPOSITE(false, false);
    }
#undef GS_LAUNCH_COMPOSITE
    return cudaGetLastError();
}
