// composite.cu — K3: tile-binned front-to-back splat compositor.
//
// Replaces the fragment + ROP-blend half of renderer.render_with_pass (reference
// src/tab/scene.rs:2302-2314; depth state scene.rs:1972-1978): the reference blends quads back
// to front in hardware; here one CTA owns a 16x16 pixel tile, stages the tile's depth-sorted
// splats in shared memory 256 at a time and blends them FRONT TO BACK
//     C += c·α·T,  T -= α·T,   α = min(0.99, o·exp(-½ dᵀQd)),
// dropping α < 1/255 and power > 0, and stops a pixel when T < 1/1024 (whole warp / whole CTA
// exit as soon as all their pixels stopped).  Each warp owns an 8x4 pixel sub-tile and first
// culls the staged splats against it 32 at a time with a ballot (exact ellipse-vs-rectangle
// footprint test, one splat per lane), so only splats that can reach alpha >= 1/255 inside the
// sub-tile are evaluated.  FP32-pipe + MUFU bound, not HBM bound.
#include "common.cuh"

namespace {

constexpr int kThreads = 256;
constexpr float kLog2e = 1.4426950408889634f;

__device__ __forceinline__ float ex2_approx(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}

template <bool FLAT, bool COUNT>
__global__ void __launch_bounds__(kThreads) k_composite(float4* __restrict__ state, uint8_t* tile_done, uint32_t resume,
                                                        uint32_t last,
                                                        const uint32_t* __restrict__ tile_vals_a,
                                                        const uint32_t* __restrict__ tile_vals_b,
                                                        const uint32_t* tile_in_b,
                                                        const uint32_t* __restrict__ ranges,
                                                        const b200gs_splat* __restrict__ splats, uint8_t* out,
                                                        size_t pitch, uint32_t W, uint32_t H, uint32_t tiles_x,
                                                        uint32_t n_tiles, float bg0, float bg1, float bg2, float bg3,
                                                        unsigned long long* evals) {
    // 64 bytes per staged splat as four float4 planes [j][splat] (a lane reading its own splat in the cull
    // loop touches consecutive 16-byte words: no bank conflicts): {mx, my, a', b'} {c', opacity, red, green}
    // {cx, hx, cy, hy} {blue, tau', -, -}; conic pre-scaled so that power is in log2 units, extent
    // square clipped to the viewport stored as centre / half-size (exact: half-integers).
    // Double-buffered: the next round's splats are in flight while this round is blended.
    __shared__ float4 sS[2][4 * kThreads];

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const uint32_t* __restrict__ tile_vals = *tile_in_b ? tile_vals_b : tile_vals_a;
    const uint32_t tile = ranges[2 * n_tiles + blockIdx.x];  // longest lists first
    const uint32_t tx = tile % tiles_x, ty = tile / tiles_x;
    const uint32_t start = ranges[tile], end = ranges[n_tiles + tile];

    // warp -> 8x4 sub-tile, lane -> pixel
    const int wx0 = (int)(tx * GS_TILE) + (warp & 1) * 8, wy0 = (int)(ty * GS_TILE) + (warp >> 1) * 4;
    const float fwx0 = (float)wx0, fwx1 = (float)(wx0 + 7), fwy0 = (float)wy0, fwy1 = (float)(wy0 + 3);
    const int px = wx0 + (lane & 7), py = wy0 + (lane >> 3);
    const float fpx = (float)px, fpy = (float)py;
    const bool inside = px < (int)W && py < (int)H;

    float T = 1.0f, Cr = 0.0f, Cg = 0.0f, Cb = 0.0f;
    bool done = !inside;
    unsigned long long my_evals = 0;
    const float Wf = (float)W, Hf = (float)H;
    // depth slabs: a tile finished by a nearer slab has its final pixels already; otherwise pick up
    // the accumulated colour / transmittance where the previous slab left them
    if (resume) {
        if (tile_done[tile]) return;
        if (inside) {
            const float4 st = state[(size_t)py * W + px];
            Cr = st.x; Cg = st.y; Cb = st.z; T = st.w;
            done = T < GS_T_EPS;
        }
    }

    // software pipeline of the staging: entry ids two rounds ahead, splat records one round ahead in
    // registers, converted into the other shared-memory buffer at the end of the current round, so a
    // round costs ONE barrier (which also carries the early-exit vote)
    uint4 q0 = make_uint4(0, 0, 0, 0), q1 = make_uint4(0, 0, 0, 0);
    uint32_t id_next = 0;
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
        dst[tid] = make_float4(mx, my, -0.5f * kLog2e * ca, -kLog2e * cbq);
        dst[kThreads + tid] = make_float4(-0.5f * kLog2e * cc, op, cr, cg);
        dst[2 * kThreads + tid] = make_float4(0.5f * (fx0 + fx1), 0.5f * (fx1 - fx0), 0.5f * (fy0 + fy1), 0.5f * (fy1 - fy0));
        dst[3 * kThreads + tid] = make_float4(cb, 0.5f * kLog2e * gs_footprint_tau(op, FLAT), r, 0.0f);
    };
    auto load_splat = [&](uint32_t id) {
        const uint4* sp = reinterpret_cast<const uint4*>(splats + id);
        q0 = __ldg(sp); q1 = __ldg(sp + 1);
    };
    if (start + tid < end) { load_splat(tile_vals[start + tid]); stage(sS[0]); }              // round 0 -> buffer 0
    if (start + kThreads + tid < end) load_splat(tile_vals[start + kThreads + tid]);            // round 1 -> registers
    if (start + 2 * kThreads + tid < end) id_next = tile_vals[start + 2 * kThreads + tid];      // round 2 ids
    __syncthreads();

    uint32_t buf = 0;
    for (uint32_t base = start; base < end; base += kThreads, buf ^= 1u) {
        const uint32_t cnt = min((uint32_t)kThreads, end - base);
        if (COUNT && tid == 0) atomicAdd(evals + 1, (unsigned long long)cnt);  // entries staged before the tile finished
        const float4* sSb = sS[buf];

        if (!__all_sync(0xffffffffu, done)) {
            for (uint32_t g = 0; g < cnt; g += 32) {
                const uint32_t s = g + lane;
                bool ov = false;
                if (s < cnt) {
                    // splat s against this warp's 8x4 sub-tile: extent-square overlap first, then the exact
                    // footprint test (can any pixel of the overlap reach alpha >= 1/255?)
                    const float4 C = sSb[2 * kThreads + s];
                    const float x0 = fmaxf(C.x - C.y, fwx0), x1 = fminf(C.x + C.y, fwx1);
                    const float y0 = fmaxf(C.z - C.w, fwy0), y1 = fminf(C.z + C.w, fwy1);
                    if (x0 <= x1 && y0 <= y1) {
                        const float4 A = sSb[s];
                        const float4 D = sSb[3 * kThreads + s];
                        const float pa = -A.z, pb = -0.5f * A.w, pc = -sSb[kThreads + s].x;  // 0.5*log2e * (a, b, c)
                        const float dx0 = x0 - A.x, dx1 = x1 - A.x, dy0 = y0 - A.y, dy1 = y1 - A.y;
                        const bool inx = dx0 <= 0.0f && dx1 >= 0.0f, iny = dy0 <= 0.0f && dy1 >= 0.0f;
                        float best = (inx && iny) ? 0.0f : 3.0e38f;
                        if (!inx) {
                            const float dx = dx0 > 0.0f ? dx0 : dx1;
                            const float dy = fminf(dy1, fmaxf(dy0, __fdividef(-pb * dx, pc)));
                            best = pa * dx * dx + 2.0f * pb * dx * dy + pc * dy * dy;
                        }
                        if (!iny) {
                            const float dy = dy0 > 0.0f ? dy0 : dy1;
                            const float dx = fminf(dx1, fmaxf(dx0, __fdividef(-pb * dy, pa)));
                            best = fminf(best, pa * dx * dx + 2.0f * pb * dx * dy + pc * dy * dy);
                        }
                        ov = best <= D.y;
                    }
                }
                uint32_t m = __ballot_sync(0xffffffffu, ov);
                // two splats per trip: their alpha evaluations are independent (ILP), the blends are
                // applied in order
                while (m) {
                    const int sa = (int)g + __ffs((int)m) - 1;
                    m &= m - 1;
                    const bool has_b = m != 0;
                    const int sb = has_b ? (int)g + __ffs((int)m) - 1 : sa;
                    m &= m - 1;  // no-op when m == 0
                    const float4 Aa = sSb[sa], Ba = sSb[kThreads + sa], Ca = sSb[2 * kThreads + sa];
                    const float4 Ab = sSb[sb], Bb = sSb[kThreads + sb], Cbb = sSb[2 * kThreads + sb];
                    const float dxa = fpx - Aa.x, dya = fpy - Aa.y, dxb = fpx - Ab.x, dyb = fpy - Ab.y;
                    const bool ina = fabsf(fpx - Ca.x) <= Ca.y && fabsf(fpy - Ca.z) <= Ca.w;
                    const bool inb = has_b && fabsf(fpx - Cbb.x) <= Cbb.y && fabsf(fpy - Cbb.z) <= Cbb.w;
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
                    if (COUNT) my_evals += (ina && !done) ? 1ull : 0ull;
                    if (ina && !done && pa2 <= 0.0f && ala >= GS_ALPHA_MIN) {
                        const float w = ala * T;
                        Cr = __fmaf_rn(Ba.z, w, Cr);
                        Cg = __fmaf_rn(Ba.w, w, Cg);
                        Cb = __fmaf_rn(sSb[3 * kThreads + sa].x, w, Cb);
                        T -= w;
                        done = T < GS_T_EPS;
                    }
                    if (COUNT) my_evals += (inb && !done) ? 1ull : 0ull;
                    if (inb && !done && pb2 <= 0.0f && alb >= GS_ALPHA_MIN) {
                        const float w = alb * T;
                        Cr = __fmaf_rn(Bb.z, w, Cr);
                        Cg = __fmaf_rn(Bb.w, w, Cg);
                        Cb = __fmaf_rn(sSb[3 * kThreads + sb].x, w, Cb);
                        T -= w;
                        done = T < GS_T_EPS;
                    }
                }
                if (__all_sync(0xffffffffu, done)) break;
            }
        }
        // next round: registers -> the other buffer, then fetch the round after it
        if (base + kThreads + tid < end) stage(sS[buf ^ 1u]);
        if (base + 2 * kThreads + tid < end) load_splat(id_next);
        if (base + 3 * kThreads + tid < end) id_next = tile_vals[base + 3 * kThreads + tid];
        if (__syncthreads_and(done)) break;
    }

    bool final_write = true;
    if (!last) {
        // not the last slab: only finished tiles write pixels now, the others park their state
        const bool all_done = __syncthreads_and(done) != 0;
        if (all_done) { if (tid == 0) tile_done[tile] = 1; }
        else {
            final_write = false;
            if (inside) state[(size_t)py * W + px] = make_float4(Cr, Cg, Cb, T);
        }
    }
    if (inside && final_write) {
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
__global__ void __launch_bounds__(32) k_query_hits(const uint32_t* __restrict__ tile_vals_a,
                                                   const uint32_t* __restrict__ tile_vals_b, const uint32_t* tile_in_b,
                                                   const uint32_t* __restrict__ ranges,
                                                   const b200gs_splat* __restrict__ splats, uint32_t W, uint32_t H,
                                                   uint32_t tiles_x, uint32_t n_tiles, uint32_t px, uint32_t py,
                                                   uint32_t flat, uint2* out, uint32_t cap, uint32_t* count) {
    const uint32_t* __restrict__ tile_vals = *tile_in_b ? tile_vals_b : tile_vals_a;
    const uint32_t tile = (py / GS_TILE) * tiles_x + px / GS_TILE;
    const uint32_t start = ranges[tile], end = ranges[n_tiles + tile];
    const int lane = threadIdx.x;
    const float fpx = (float)px, fpy = (float)py, Wf = (float)W, Hf = (float)H;
    uint32_t n_out = 0;
    for (uint32_t base = start; base < end; base += 32) {
        const uint32_t e = base + lane;
        bool ok = false;
        uint32_t id = 0;
        float al = 0.0f;
        if (e < end) {
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
    k_query_hits<<<1, 32, 0, st>>>(a.tile_vals, a.tile_vals_b, a.tile_in_b, a.ranges, a.splats, (uint32_t)f.W, (uint32_t)f.H,
                                   f.tiles_x, f.tiles_x * f.tiles_y, px, py, f.display_mode != B200GS_DISPLAY_SPLAT ? 1u : 0u,
                                   out, cap, count);
    return cudaGetLastError();
}

cudaError_t gs_launch_composite(const GsCompositeArgs& a, const GsFrame& f, cudaStream_t st) {
    const uint32_t n_tiles = f.tiles_x * f.tiles_y;
    const bool flat = f.display_mode != B200GS_DISPLAY_SPLAT;
    const uint32_t W = (uint32_t)f.W, H = (uint32_t)f.H;
#define GS_LAUNCH_COMPOSITE(FLAT, COUNT)                                                                              \
    k_composite<FLAT, COUNT><<<n_tiles, kThreads, 0, st>>>(a.state, a.tile_done, a.resume ? 1u : 0u, a.last ? 1u : 0u, a.tile_vals, a.tile_vals_b, a.tile_in_b, a.ranges, a.splats, a.out, a.pitch, W, H,      \
                                                           f.tiles_x, n_tiles, f.bg[0], f.bg[1], f.bg[2], f.bg[3],     \
                                                           a.evals)
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
