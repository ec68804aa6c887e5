// preprocess.cu — K1: the preprocess kernel of the B200 3DGS render core.
//
// Replaces viewer.preprocessor.preprocess(encoder, bind_group, N) (reference
// src/tab/scene.rs:856-863, bindings :1835-1852) and, per the north star, the per-splat vertex
// work of renderer.render_with_pass (scene.rs:2306-2313): model/view/projection transform,
// frustum cull, mask + hidden-edit + selection test, 3D->2D covariance projection, conic and
// extent, SH colour up to degree 3, edits and highlight.  Emits, for the V visible Gaussians in
// ASCENDING INDEX ORDER (order-preserving compaction by decoupled look-back, so that the later
// stable sort is deterministic): depth key, Gaussian index, a 32-byte projected splat and a 4-byte
// bin word (the splat's candidate tile rectangle, consumed by the binning kernel, common.cuh).
//
// Data movement: persistent CTAs of 7 compute warps + 1 control warp (4 per SM for records up to 76 bytes); each
// 224-Gaussian chunk (224*R contiguous bytes, 16-byte aligned because R is a multiple of 4) is pulled into
// shared memory with one 1-D TMA bulk copy (cp.async.bulk -> UBLKCP) on an mbarrier, NSTAGE
// chunks in flight per CTA; each compute thread then reads its own record from shared memory
// (word stride R/4).  The control warp owns tickets, TMA issue and the decoupled look-back, so
// the look-back latency overlaps the compute warps' heavy phase (named barriers, no
// __syncthreads in the loop).
//
// THIS FILE IS COMPILED WITH -fmad=false: in the EXACT class (position chain, cull, depth key,
// pixel centre, covariance -> conic/extent) every float operation rounds separately, in source
// order, and multiply-adds are fused ONLY where __fmaf_rn is spelled out — the oracle (built with
// -ffp-contract=off) spells fmaf() in the same places — so that those results are bit-identical
// to the CPU oracle's.  Do not re-associate that arithmetic.  The TOLERANCE class (SH colour,
// edits) uses explicit __fmaf_rn / rsqrtf and is compared within the RGBA tolerance.
#include <mutex>

#include "common.cuh"

namespace {

constexpr int kChunk = 224;          // Gaussians per chunk = 7 compute warps: with 3 stages of 224 x 76 B a CTA needs 55 KB and 4 fit an SM
constexpr int kCW = kChunk / 32;    // compute warps
template <int SH> struct ShBytes { static constexpr int v = SH == 0 ? 180 : SH == 1 ? 92 : SH == 2 ? 48 : 0; };
template <int COV> struct CovBytes { static constexpr int v = COV == 0 ? 24 : 12; };


__device__ __forceinline__ float h2f_lo(uint32_t w) { return __half2float(__ushort_as_half((unsigned short)(w & 0xffffu))); }
__device__ __forceinline__ float h2f_hi(uint32_t w) { return __half2float(__ushort_as_half((unsigned short)(w >> 16))); }
__device__ __forceinline__ float clamp01(float x) { return x < 0.0f ? 0.0f : (x > 1.0f ? 1.0f : x); }

// SH coefficient k (0..44) of the record whose SH field starts at word pointer `w`
template <int SH>
__device__ __forceinline__ float sh_coef(const uint32_t* w, int k) {
    if (SH == 0) return __uint_as_float(w[k]);
    if (SH == 1) { uint32_t x = w[k >> 1]; return (k & 1) ? h2f_hi(x) : h2f_lo(x); }
    if (SH == 2) { uint32_t x = w[k >> 2]; return (float)((x >> ((k & 3) * 8)) & 0xffu) * (2.0f / 255.0f) - 1.0f; }
    return 0.0f;
}

__device__ void rgb_to_hsv(const float c[3], float hsv[3]) {
    float mx = fmaxf(c[0], fmaxf(c[1], c[2])), mn = fminf(c[0], fminf(c[1], c[2]));
    float d = mx - mn, h = 0.0f;
    if (d > 0.0f) {
        if (mx == c[0]) h = (c[1] - c[2]) / d;
        else if (mx == c[1]) h = 2.0f + (c[2] - c[0]) / d;
        else h = 4.0f + (c[0] - c[1]) / d;
        h = h / 6.0f;
        if (h < 0.0f) h = h + 1.0f;
    }
    hsv[0] = h;
    hsv[1] = mx > 0.0f ? d / mx : 0.0f;
    hsv[2] = mx;
}
__device__ void hsv_to_rgb(const float hsv[3], float c[3]) {
    float h = hsv[0] * 6.0f, s = hsv[1], v = hsv[2];
    float i = floorf(h), f = h - i;
    int k = ((int)i) % 6;
    if (k < 0) k += 6;
    float p = v * (1.0f - s), q = v * (1.0f - s * f), t = v * (1.0f - s * (1.0f - f));
    switch (k) {
        case 0: c[0] = v; c[1] = t; c[2] = p; break;
        case 1: c[0] = q; c[1] = v; c[2] = p; break;
        case 2: c[0] = p; c[1] = v; c[2] = t; break;
        case 3: c[0] = p; c[1] = q; c[2] = v; break;
        case 4: c[0] = t; c[1] = p; c[2] = v; break;
        default: c[0] = v; c[1] = p; c[2] = q; break;
    }
}
// GaussianEditPod application (inputs: src/app.rs:1533-1564): colour -> contrast -> exposure ->
// gamma -> alpha.
__device__ void apply_edit(const b200gs_edit_pod& e, float rgb[3], float& op) {
    if (!(e.flag & B200GS_EDIT_ENABLED)) return;
    if (e.flag & B200GS_EDIT_OVERRIDE_COLOR) {
        rgb[0] = e.color[0]; rgb[1] = e.color[1]; rgb[2] = e.color[2];
    } else {
        float hsv[3];
        rgb_to_hsv(rgb, hsv);
        float h = hsv[0] + e.color[0];
        hsv[0] = h - floorf(h);
        hsv[1] = clamp01(hsv[1] * e.color[1]);
        hsv[2] = hsv[2] * e.color[2];
        hsv_to_rgb(hsv, rgb);
    }
    float ex = exp2f(e.exposure);
#pragma unroll
    for (int c = 0; c < 3; c++) {
        float v = (rgb[c] - 0.5f) * (1.0f + e.contrast) + 0.5f;
        v = v * ex;
        v = powf(fmaxf(v, 0.0f), e.gamma);
        rgb[c] = v;
    }
    op = clamp01(op * e.alpha);
}

// named barriers (id 0 is __syncthreads); each kind has 4 ids, indexed by chunk sequence number & 3,
// because up to 3 generations of one kind can be outstanding (see the loop comments)
constexpr int kBarCounts = 1;   // compute warps -> control warp: per-warp visible counts of chunk j are in smem
constexpr int kBarBase = 5;     // control warp -> compute warps: chunk j's output base is in smem
constexpr int kBarFree = 9;     // compute warps -> control warp: chunk j's stage can be refilled
constexpr int kThreads = kChunk + 32;  // the compute warps + 1 control warp

__device__ __forceinline__ void bar_sync(int id) { asm volatile("bar.sync %0, %1;" ::"r"(id), "n"(kThreads) : "memory"); }
__device__ __forceinline__ void bar_arrive(int id) { asm volatile("bar.arrive %0, %1;" ::"r"(id), "n"(kThreads) : "memory"); }

// byte k of x as a float, without I2F: PRMT builds 0x4B0000bb (= 8388608 + b), one FADD removes the bias
__device__ __forceinline__ float byte_to_float(uint32_t x, int k) {
    return __uint_as_float(__byte_perm(x, 0x4B000000u, 0x7440u | (uint32_t)k)) - 8388608.0f;
}

// SH bands 1..DEG added to rgb (TOLERANCE class: FMA).  Inria computeColorFromSH sign convention.
template <int SH, int DEG>
__device__ __forceinline__ void sh_colour(const uint32_t* shw, float dx, float dy, float dz, float rgb[3]) {
    constexpr int NCOEF = DEG >= 3 ? 15 : (DEG == 2 ? 8 : 3);
    float bs[NCOEF];
    bs[0] = -0.4886025119029199f * dy;
    bs[1] = 0.4886025119029199f * dz;
    bs[2] = -0.4886025119029199f * dx;
    if (DEG >= 2) {
        const float xx = dx * dx, yy = dy * dy, zz = dz * dz, xy = dx * dy, yz = dy * dz, xz = dx * dz;
        bs[3] = 1.0925484305920792f * xy;
        bs[4] = -1.0925484305920792f * yz;
        bs[5] = 0.31539156525252005f * (2.0f * zz - xx - yy);
        bs[6] = -1.0925484305920792f * xz;
        bs[7] = 0.5462742152960396f * (xx - yy);
        if (DEG >= 3) {
            bs[8] = -0.5900435899266435f * dy * (3.0f * xx - yy);
            bs[9] = 2.890611442640554f * xy * dz;
            bs[10] = -0.4570457994644658f * dy * (4.0f * zz - xx - yy);
            bs[11] = 0.3731763325901154f * dz * (2.0f * zz - 3.0f * xx - 3.0f * yy);
            bs[12] = -0.4570457994644658f * dx * (4.0f * zz - xx - yy);
            bs[13] = 1.445305721320277f * dz * (xx - yy);
            bs[14] = -0.5900435899266435f * dx * (xx - 3.0f * yy);
        }
    }
    float acc[3] = {0.0f, 0.0f, 0.0f};
    if (SH == 2) {
        // unorm8 coefficient q -> x = 1 + q·2^-15 by ONE PRMT (the byte dropped into mantissa bits 8..15 of
        // 1.0f), so a coefficient costs PRMT + FFMA:  Σ b_k x_k = Σ b_k + 2^-15 Σ b_k q_k,  and
        // Σ b_k (q_k·2/255 − 1) = (2^16/255)·(Σ b_k x_k − Σ b_k) − Σ b_k   (error ~3e-4, tolerance class)
        float sum_b = 0.0f;
#pragma unroll
        for (int k = 0; k < NCOEF; k++) {
            sum_b += bs[k];
#pragma unroll
            for (int ch = 0; ch < 3; ch++) {
                const int e = 3 * k + ch;
                const float x = __uint_as_float(__byte_perm(shw[e >> 2], 0x3F800000u, 0x7604u | ((uint32_t)(e & 3) << 4)));
                acc[ch] = __fmaf_rn(bs[k], x, acc[ch]);
            }
        }
#pragma unroll
        for (int ch = 0; ch < 3; ch++) rgb[ch] += __fmaf_rn(acc[ch] - sum_b, 65536.0f / 255.0f, -sum_b);
    } else {
#pragma unroll
        for (int k = 0; k < NCOEF; k++) {
#pragma unroll
            for (int ch = 0; ch < 3; ch++) acc[ch] = __fmaf_rn(bs[k], sh_coef<SH>(shw, 3 * k + ch), acc[ch]);
        }
#pragma unroll
        for (int ch = 0; ch < 3; ch++) rgb[ch] += acc[ch];
    }
}

// EXACT class: model/view/projection chain, frustum cull.  Returns visibility; pw/pv/ndc out.
template <bool FAST>
__device__ __forceinline__ bool project_and_cull(const uint32_t* w, const GsFrame& f, const GsModelXf& m, float pw[3],
                                                 float pv[3], float& nx, float& ny, float& nz) {
    const float p0 = __uint_as_float(w[0]), p1 = __uint_as_float(w[1]), p2 = __uint_as_float(w[2]);
    if (FAST || m.identity) {  // bit-identical to the general form when R = I, s = 1, t = 0
        pw[0] = p0; pw[1] = p1; pw[2] = p2;
    } else {
        // world = q*(s⊙p)+t  (src/app.rs:1044-1046)
        const float ps0 = m.s[0] * p0, ps1 = m.s[1] * p1, ps2 = m.s[2] * p2;
#pragma unroll
        for (int r = 0; r < 3; r++) pw[r] = __fmaf_rn(m.R[r][0], ps0, __fmaf_rn(m.R[r][1], ps1, __fmaf_rn(m.R[r][2], ps2, m.t[r])));
    }
#pragma unroll
    for (int r = 0; r < 3; r++) pv[r] = __fmaf_rn(f.V[r][0], pw[0], __fmaf_rn(f.V[r][1], pw[1], __fmaf_rn(f.V[r][2], pw[2], f.V[r][3])));
    float pc[4];
    if (FAST || f.std_proj) {  // zero terms of glam's perspective_rh skipped: fma(0, x, y) = y, so the bits are the same
        pc[0] = f.P[0][0] * pv[0];
        pc[1] = f.P[1][1] * pv[1];
        pc[2] = __fmaf_rn(f.P[2][2], pv[2], f.P[2][3]);
        pc[3] = f.P[3][2] * pv[2];
    } else {
#pragma unroll
        for (int r = 0; r < 4; r++) pc[r] = __fmaf_rn(f.P[r][0], pv[0], __fmaf_rn(f.P[r][1], pv[1], __fmaf_rn(f.P[r][2], pv[2], f.P[r][3])));
    }
    if (!(pc[3] > 0.0f)) return false;
    const float iw = 1.0f / pc[3];
    nx = pc[0] * iw; ny = pc[1] * iw; nz = pc[2] * iw;
    return nz > 0.0f && nz < 1.0f && fabsf(nx) <= GS_CULL_XY && fabsf(ny) <= GS_CULL_XY;
}

// Selection query: the splat centre against the shape itself (immediate mode) or against the viewer's query texture
// (non-immediate mode, scene.rs:767-791).  EXACT class: same float operations as the oracle.
__device__ __forceinline__ bool query_hit(const GsFrame& f, float sx, float sy) {
    if (f.query.kind == B200GS_QUERY_TEXTURE) {
        const float fx = floorf(sx), fy = floorf(sy);
        if (!(fx >= 0.0f && fy >= 0.0f && fx < (float)f.query_tex_w && fy < (float)f.query_tex_h)) return false;
        return f.query_tex[(size_t)(uint32_t)fy * f.query_tex_w + (uint32_t)fx] != 0;
    }
    return gs_query_shape_hit(f.query, sx, sy);
}

// mask / hidden-edit / selection tests that precede the frustum cull (reference preprocess bindings,
// src/tab/scene.rs:1835-1852; flags src/app.rs:1548-1551)
__device__ __forceinline__ bool pre_tests(uint32_t i, uint32_t n, const uint32_t* __restrict__ mask,
                                          const uint32_t* selection,
                                          const b200gs_edit_pod* __restrict__ edits, const GsFrame& f, bool& selected,
                                          b200gs_edit_pod& ed) {
    bool vis = i < n;
    selected = false;
    ed.flag = 0;
    if (vis && mask) vis = (mask[i >> 5] >> (i & 31)) & 1u;
    if (vis && selection) selected = (selection[i >> 5] >> (i & 31)) & 1u;
    if (vis && edits) {
        const uint4* ep = reinterpret_cast<const uint4*>(edits + i);
        const uint4 e0 = ep[0], e1 = ep[1];
        ed.flag = e0.x; ed.color[0] = __uint_as_float(e0.y); ed.color[1] = __uint_as_float(e0.z);
        ed.color[2] = __uint_as_float(e0.w); ed.contrast = __uint_as_float(e1.x);
        ed.exposure = __uint_as_float(e1.y); ed.gamma = __uint_as_float(e1.z); ed.alpha = __uint_as_float(e1.w);
        if ((ed.flag & B200GS_EDIT_ENABLED) && (ed.flag & B200GS_EDIT_HIDDEN)) vis = false;
    }
    if (vis && selected && (f.sel_edit.flag & B200GS_EDIT_ENABLED) && (f.sel_edit.flag & B200GS_EDIT_HIDDEN)) vis = false;
    return vis;
}

// FAST: the launcher found the common frame — identity model transform, glam perspective projection, no mask /
// selection / edit buffers, no selection query, Splat display, SH degree 3 with SH0 — and picks the instantiation in
// which those uniform branches (and the loads of their operands) are compiled out.  Same arithmetic, same bits.
template <int SH, int COV, bool FAST>
__global__ void __launch_bounds__(kThreads, (16 + ShBytes<SH>::v + CovBytes<COV>::v) <= 76 ? 4 : 3) k_preprocess(const uint8_t* __restrict__ recs, uint32_t n,
                                                         const uint32_t* __restrict__ mask, uint32_t* selection,
                                                         const b200gs_edit_pod* __restrict__ edits,
                                                         const __grid_constant__ GsFrame f,
                                                         const __grid_constant__ GsModelXf m, uint32_t* ctrl,
                                                         uint64_t* lookback, uint32_t epoch,
                                                         uint32_t* __restrict__ keys, uint32_t* __restrict__ idx,
                                                         b200gs_splat* __restrict__ splats, uint32_t* __restrict__ binword,
                                                         uint32_t* sort_hist) {
    constexpr int RB = 16 + ShBytes<SH>::v + CovBytes<COV>::v;  // record bytes
    constexpr int RW = RB / 4;                                  // record words
    constexpr int NSTAGE = 3;                                   // chunk j, chunk j+1 (counted ahead), one in flight
    constexpr uint32_t STAGE_BYTES = kChunk * RB;

    extern __shared__ __align__(128) uint8_t smem[];
    uint8_t* stage_mem = smem;  // NSTAGE * STAGE_BYTES
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + NSTAGE * STAGE_BYTES);
    uint32_t* s_chunk = reinterpret_cast<uint32_t*>(bars + NSTAGE);  // NSTAGE
    uint32_t* s_wcount_all = s_chunk + NSTAGE + 1;                   // 4 x 8 warp counts (by sequence number & 3)
    uint32_t* s_base_all = s_wcount_all + 32;                        // 4 x 8 output bases of the warps (by sequence number & 3)
    uint32_t* s_hist = s_base_all + 32;                              // 4 x 256 (only if sort_hist)

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const uint32_t nchunks = (n + kChunk - 1) / kChunk;
    const bool is_control = warp == kChunk / 32;

    // fill `stage` with chunk c; a chunk past the end completes the stage's barrier with no data, so that the
    // compute warps can always wait on the barrier BEFORE they read which chunk the stage holds (they run
    // ahead of the refill: the barrier is what orders their read of s_chunk behind it)
    auto issue = [&](int stage, uint32_t c) {
        if (c >= nchunks) { gs_mbar_arrive(&bars[stage]); return; }
        uint32_t cnt = min((uint32_t)kChunk, n - c * kChunk);
        uint32_t bytes = (cnt * RB + 15u) & ~15u;  // buffer is padded, see api.cu
        gs_mbar_expect_tx(&bars[stage], bytes);
        gs_tma_load_1d(stage_mem + (size_t)stage * STAGE_BYTES, recs + (size_t)c * STAGE_BYTES, bytes, &bars[stage]);
        // Chunks are handed out in order, so chunk c + 2 * gridDim.x is what SOME CTA will draw two rounds from now: pull
        // it into L2 today, and that CTA's bulk copy then completes in an L2 round trip instead of a DRAM one (the
        // stage refill is issued only one heavy phase before the data is needed; 14 % of warp time used to wait on it).
        const uint32_t cp = c + 2u * gridDim.x;
        if (cp < nchunks) {
            const uint32_t pb = (min((uint32_t)kChunk, n - cp * kChunk) * RB) & ~15u;
            if (pb) asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(recs + (size_t)cp * STAGE_BYTES), "r"(pb) : "memory");
        }
    };

    if (tid == kChunk) {
        for (int s = 0; s < NSTAGE; s++) gs_mbar_init(&bars[s], 1);
        gs_fence_mbar_init();
        for (int s = 0; s < NSTAGE; s++) {
            uint32_t c = atomicAdd(&ctrl[GS_CTRL_TICKET], 1u);
            s_chunk[s] = c;
            issue(s, c);
        }
    }
    if (sort_hist)
        for (int i = tid; i < 1024; i += kThreads) s_hist[i] = 0;
    __syncthreads();

    // The CTA works through the chunks it drew, j = 0, 1, 2, ... (stage j % 3).  The cull/count of
    // chunk j+1 runs BEFORE the heavy phase of chunk j, so chunk j+1's visible count is published a
    // whole iteration before its output base is needed: the decoupled look-back (control warp)
    // overlaps a full heavy phase and is off the critical path.
    if (is_control) {
        // ------------------------------------------------------------ control warp
        // publish(j+1) happens one iteration BEFORE resolve(j+1): by the time a chunk's prefix is
        // resolved, its predecessors (drawn at the same moment by other CTAs) have long published.
        // (woff: this lane's warp's offset inside the chunk, lanes 0..7 — the exclusive scan of the 8 warp counts)
        auto publish = [&](uint32_t j, uint32_t c, uint32_t& woff) -> uint32_t {
            bar_sync(kBarCounts + (int)(j & 3u));  // counts of chunk j are in smem
            const uint32_t cnt = lane < kCW ? s_wcount_all[(j & 3u) * 8 + lane] : 0u;
            uint32_t incl = cnt;
#pragma unroll
            for (int o = 1; o < 8; o <<= 1) {
                const uint32_t t = __shfl_up_sync(0xffffffffu, incl, o);
                if (lane >= o) incl += t;
            }
            woff = incl - cnt;
            const uint32_t total = __shfl_sync(0xffffffffu, incl, kCW - 1);
            if (lane == 0) gs_lookback_publish(lookback, epoch, c, total);
            return total;
        };
        uint32_t c = s_chunk[0];
        uint32_t woff = 0, woff_next = 0;
        uint32_t total = c < nchunks ? publish(0, c, woff) : 0u;
        for (uint32_t j = 0; c < nchunks; j++) {
            // the compute warps count chunk j+1 at the start of their iteration j
            const uint32_t c_next = s_chunk[(j + 1) % NSTAGE];
            const uint32_t total_next = c_next < nchunks ? publish(j + 1, c_next, woff_next) : 0u;
            if (j > 0) {
                // chunk j-1's stage was read for the last time in its heavy phase: refill it
                bar_sync(kBarFree + (int)((j - 1) & 3u));
                if (lane == 0) {
                    const uint32_t c2 = atomicAdd(&ctrl[GS_CTRL_TICKET], 1u);
                    s_chunk[(j - 1) % NSTAGE] = c2;
                    issue((int)((j - 1) % NSTAGE), c2);
                }
                __syncwarp();
            }
            const uint32_t excl = gs_lookback_resolve(lookback, epoch, c, total, lane);
            if (lane < kCW) s_base_all[(j & 3u) * 8 + lane] = excl + woff;   // output base of every warp of the chunk
            if (lane == 0 && c == nchunks - 1) ctrl[GS_CTRL_VISIBLE] = excl + total;
            bar_arrive(kBarBase + (int)(j & 3u));
            c = c_next;
            total = total_next;
            woff = woff_next;
        }
    } else {
        // ------------------------------------------------------------ compute warps
        // phase 1 of chunk j (cheap, exact class): mask / edit / selection tests, projection, frustum cull;
        // publishes the warp's visible count and keeps the results in registers for the heavy phase
        struct Phase1 { bool vis; uint32_t ballot; float pv[3], nx, ny, nz; };
        auto phase1 = [&](uint32_t j) -> Phase1 {
            Phase1 r;
            r.vis = false; r.ballot = 0u; r.pv[0] = r.pv[1] = r.pv[2] = r.nx = r.ny = r.nz = 0.0f;
            const int stage = j % NSTAGE;
            gs_mbar_wait(&bars[stage], (j / NSTAGE) & 1u);   // also orders the read of s_chunk behind the refill
            const uint32_t c = s_chunk[stage];
            if (c >= nchunks) return r;
            const uint32_t i = c * kChunk + tid;
            const uint32_t* w = reinterpret_cast<const uint32_t*>(stage_mem + (size_t)stage * STAGE_BYTES) + tid * RW;
            bool selected;
            b200gs_edit_pod ed;
            r.vis = FAST ? i < n : pre_tests(i, n, mask, selection, edits, f, selected, ed);
            if (r.vis) {
                float pw[3];
                r.vis = project_and_cull<FAST>(w, f, m, pw, r.pv, r.nx, r.ny, r.nz);
            }
            r.ballot = __ballot_sync(0xffffffffu, r.vis);
            if (lane == 0) s_wcount_all[(j & 3u) * 8 + warp] = __popc(r.ballot);
            bar_arrive(kBarCounts + (int)(j & 3u));
            return r;
        };
        // The outputs of chunk it are written TWO iterations later (after phase 1 of chunk it+3, at the start of
        // iteration it+2): the chunk's output base needs the counts of all 8 warps and the control warp's
        // look-back; written one iteration later, 11 % of warp time still sat at the base barrier.
        struct Pending { bool valid, vis; uint32_t seq, ballot, key, index, bw; uint4 q0, q1; };
        Pending pend, pend2;   // pend: chunk it-1, pend2: chunk it-2 (the one written in iteration it)
        pend.valid = false; pend.vis = false; pend.seq = pend.ballot = pend.key = pend.index = pend.bw = 0;
        pend.q0 = pend.q1 = make_uint4(0, 0, 0, 0);
        pend2 = pend;
        auto flush = [&](Pending& pd) {
            if (!pd.valid) return;   // (CTA-uniform)
            bar_sync(kBarBase + (int)(pd.seq & 3u));
            if (pd.vis) {
                const uint32_t off = s_base_all[(pd.seq & 3u) * 8 + warp] + __popc(pd.ballot & ((1u << lane) - 1u));
                keys[off] = pd.key;
                idx[off] = pd.index;
                uint4* sp = reinterpret_cast<uint4*>(splats + off);
                sp[0] = pd.q0;
                sp[1] = pd.q1;
                binword[off] = pd.bw;
            }
            pd.valid = false;
        };
        Phase1 cur = phase1(0);
#pragma unroll 2
        for (uint32_t it = 0;; it++) {
            const int stage = it % NSTAGE;
            const uint32_t c = s_chunk[stage];
            if (c >= nchunks) break;
            // chunk it+1 is culled and counted BEFORE the heavy phase of chunk it (see above)
            const Phase1 nxt = phase1(it + 1);
            flush(pend2);   // chunk it-2: its base has had two heavy phases to resolve
            pend2 = pend;

            const uint32_t i = c * kChunk + tid;
            const uint32_t* w = reinterpret_cast<const uint32_t*>(stage_mem + (size_t)stage * STAGE_BYTES) + tid * RW;
            const bool vis = cur.vis;
            const uint32_t ballot = cur.ballot;
            const float nx = cur.nx, ny = cur.ny, nz = cur.nz;
            float pv[3] = {cur.pv[0], cur.pv[1], cur.pv[2]};
            // per-Gaussian inputs of the heavy phase that phase 1 did not keep
            bool selected = false;
            b200gs_edit_pod ed;
            ed.flag = 0;
            if (!FAST && (mask || selection || edits)) (void)pre_tests(i, n, mask, selection, edits, f, selected, ed);
            float pw[3];
            {
                const float p0 = __uint_as_float(w[0]), p1 = __uint_as_float(w[1]), p2 = __uint_as_float(w[2]);
                if (FAST || m.identity) { pw[0] = p0; pw[1] = p1; pw[2] = p2; }
                else {
                    const float ps0 = m.s[0] * p0, ps1 = m.s[1] * p1, ps2 = m.s[2] * p2;
#pragma unroll
                    for (int r = 0; r < 3; r++)
                        pw[r] = __fmaf_rn(m.R[r][0], ps0, __fmaf_rn(m.R[r][1], ps1, __fmaf_rn(m.R[r][2], ps2, m.t[r])));
                }
            }

            // ---------------- selection query: rewrites this warp's selection word (32 Gaussians) -------
            if (!FAST && f.query.kind >= B200GS_QUERY_RECT && selection) {
                const float sx = __fmaf_rn(nx + 1.0f, f.W, -1.0f) * 0.5f + 0.5f, sy = __fmaf_rn(1.0f - ny, f.H, -1.0f) * 0.5f + 0.5f;
                const bool hit = vis && query_hit(f, sx, sy);
                const uint32_t hits = __ballot_sync(0xffffffffu, hit);
                if (i - lane < n) {  // warp-uniform: the word exists
                    uint32_t word = lane == 0 ? selection[i >> 5] : 0u;
                    word = __shfl_sync(0xffffffffu, word, 0);
                    const uint32_t neww = f.query.op == B200GS_SELECT_SET ? hits
                                          : (f.query.op == B200GS_SELECT_ADD ? (word | hits) : (word & ~hits));
                    if (lane == 0 && neww != word) selection[i >> 5] = neww;
                    selected = (neww >> lane) & 1u;   // shown (highlight / edit) with the new selection
                }
            }

            // ---------------- phase 2: projected splat for the visible ones ----------------
            // (warp-uniform branch: the culled lanes of a warp with any visible Gaussian run the arithmetic on
            // whatever their record holds and store nothing — no divergence bookkeeping inside)
            uint4 q0 = make_uint4(0, 0, 0, 0), q1 = make_uint4(0, 0, 0, 0);
            uint32_t bw = 0;
            if (ballot != 0u) {
                const uint32_t* shw = w + 4;
                const uint32_t* cw = w + 4 + ShBytes<SH>::v / 4;
                float cv[6];
                if (COV == 0) {
#pragma unroll
                    for (int k = 0; k < 6; k++) cv[k] = __uint_as_float(cw[k]);
                } else {
#pragma unroll
                    for (int k = 0; k < 3; k++) { uint32_t x = cw[k]; cv[2 * k] = h2f_lo(x); cv[2 * k + 1] = h2f_hi(x); }
                }
                // Σ' = (R_m S_m) Σ (R_m S_m)^T · size²  (upper triangle; exact class)
                float Sw[3][3];
                if (FAST || m.identity) {
                    Sw[0][0] = cv[0] * f.sz2; Sw[0][1] = cv[1] * f.sz2; Sw[0][2] = cv[2] * f.sz2;
                    Sw[1][1] = cv[3] * f.sz2; Sw[1][2] = cv[4] * f.sz2; Sw[2][2] = cv[5] * f.sz2;
                } else {
                    const float S[3][3] = {{cv[0], cv[1], cv[2]}, {cv[1], cv[3], cv[4]}, {cv[2], cv[4], cv[5]}};
                    float B[3][3];
#pragma unroll
                    for (int r = 0; r < 3; r++)
#pragma unroll
                        for (int k = 0; k < 3; k++) B[r][k] = __fmaf_rn(m.M[r][0], S[0][k], __fmaf_rn(m.M[r][1], S[1][k], m.M[r][2] * S[2][k]));
#pragma unroll
                    for (int r = 0; r < 3; r++)
#pragma unroll
                        for (int k = r; k < 3; k++)
                            Sw[r][k] = __fmaf_rn(B[r][0], m.M[k][0], __fmaf_rn(B[r][1], m.M[k][1], B[r][2] * m.M[k][2])) * f.sz2;
                }
                Sw[1][0] = Sw[0][1]; Sw[2][0] = Sw[0][2]; Sw[2][1] = Sw[1][2];

                // Jacobian of the pixel mapping (view space RH, looking down -z; rows flipped in y)
                const float tz = -pv[2];
                const float itz = 1.0f / tz;
                float txz = pv[0] * itz, tyz = pv[1] * itz;
                txz = fminf(f.limx, fmaxf(-f.limx, txz));
                tyz = fminf(f.limy, fmaxf(-f.limy, tyz));
                const float xc = txz * tz, yc = tyz * tz;
                const float itz2 = itz * itz;
                const float J00 = f.fx * itz, J02 = (f.fx * xc) * itz2;
                const float J11 = -(f.fy * itz), J12 = -((f.fy * yc) * itz2);
                float T0[3], T1[3], U0[3], U1[3];
#pragma unroll
                for (int k = 0; k < 3; k++) {
                    T0[k] = __fmaf_rn(J00, f.V[0][k], J02 * f.V[2][k]);
                    T1[k] = __fmaf_rn(J11, f.V[1][k], J12 * f.V[2][k]);
                }
#pragma unroll
                for (int k = 0; k < 3; k++) {
                    U0[k] = __fmaf_rn(T0[0], Sw[0][k], __fmaf_rn(T0[1], Sw[1][k], T0[2] * Sw[2][k]));
                    U1[k] = __fmaf_rn(T1[0], Sw[0][k], __fmaf_rn(T1[1], Sw[1][k], T1[2] * Sw[2][k]));
                }
                float a = __fmaf_rn(U0[0], T0[0], __fmaf_rn(U0[1], T0[1], U0[2] * T0[2]));
                const float b = __fmaf_rn(U0[0], T1[0], __fmaf_rn(U0[1], T1[1], U0[2] * T1[2]));
                float d = __fmaf_rn(U1[0], T1[0], __fmaf_rn(U1[1], T1[1], U1[2] * T1[2]));
                a = a + GS_LOWPASS;
                d = d + GS_LOWPASS;
                const float det = __fmaf_rn(a, d, -(b * b));
                float ca = 0.0f, cb = 0.0f, cc = 0.0f, radf = 0.0f;
                if (det > 0.0f) {
                    const float di = 1.0f / det;
                    ca = d * di; cb = -b * di; cc = a * di;
                    const float mid = 0.5f * (a + d);
                    float disc = __fmaf_rn(mid, mid, -det);
                    if (disc < GS_MIN_DISC) disc = GS_MIN_DISC;
                    const float lam = mid + sqrtf(disc);
                    radf = ceilf(GS_EXTENT_SIGMA * sqrtf(lam));
                }
                if (!FAST && f.display_mode == B200GS_DISPLAY_POINT) {
                    ca = GS_FLAT_D2 / (GS_POINT_RADIUS * GS_POINT_RADIUS); cb = 0.0f; cc = ca;
                    radf = ceilf(GS_POINT_RADIUS);
                }
                if (!(radf <= 65535.0f)) radf = 65535.0f;

                // ---- colour (tolerance class: FMA and rsqrt allowed): baked SH0 (u8) + bands 1..deg,
                // view direction in world space
                float dx = pw[0] - f.cam[0], dy = pw[1] - f.cam[1], dz = pw[2] - f.cam[2];
                const float il = rsqrtf(__fmaf_rn(dx, dx, __fmaf_rn(dy, dy, dz * dz)));
                dx *= il; dy *= il; dz *= il;
                const uint32_t colw = w[3];
                float rgb[3];
#pragma unroll
                for (int ch = 0; ch < 3; ch++) rgb[ch] = (!FAST && f.no_sh0) ? 0.0f : byte_to_float(colw, ch) * (1.0f / 255.0f);
                if (SH != 3 && FAST) sh_colour<SH, 3>(shw, dx, dy, dz, rgb);
                else if (SH != 3) {  // the degree is uniform per frame: one fully unrolled variant per degree
                    switch (f.sh_deg) {
                        case 1: sh_colour<SH, 1>(shw, dx, dy, dz, rgb); break;
                        case 2: sh_colour<SH, 2>(shw, dx, dy, dz, rgb); break;
                        case 3: sh_colour<SH, 3>(shw, dx, dy, dz, rgb); break;
                        default: break;
                    }
                }
                if (!FAST) {   // (FAST has no edits: the clamp after them is the only one needed)
#pragma unroll
                    for (int ch = 0; ch < 3; ch++) rgb[ch] = clamp01(rgb[ch]);
                }
                float op = (float)(colw >> 24) * (1.0f / 255.0f);
                if (!FAST && edits) apply_edit(ed, rgb, op);
                if (!FAST && selected) {
                    apply_edit(f.sel_edit, rgb, op);
                    const float ha = f.hl[3];
#pragma unroll
                    for (int ch = 0; ch < 3; ch++) rgb[ch] = rgb[ch] + (f.hl[ch] - rgb[ch]) * ha;
                }
#pragma unroll
                for (int ch = 0; ch < 3; ch++) rgb[ch] = clamp01(rgb[ch]);

                const float mx = __fmaf_rn(nx + 1.0f, f.W, -1.0f) * 0.5f;
                const float my = __fmaf_rn(1.0f - ny, f.H, -1.0f) * 0.5f;
                q0.x = __float_as_uint(mx);
                q0.y = __float_as_uint(my);
                q0.z = (uint32_t)radf | ((uint32_t)__half_as_ushort(__float2half_rn(op)) << 16);
                q0.w = (uint32_t)__half_as_ushort(__float2half_rn(rgb[0])) |
                       ((uint32_t)__half_as_ushort(__float2half_rn(rgb[1])) << 16);
                q1.x = __float_as_uint(ca);
                q1.y = __float_as_uint(cb);
                q1.z = __float_as_uint(cc);
                q1.w = (uint32_t)__half_as_ushort(__float2half_rn(rgb[2])) | ((selected ? 1u : 0u) << 16);
                // bin word (from the STORED record, exactly as the binning kernel decodes big splats): the tile
                // tests of the common small splats are done here, where the splat is in registers
                if (FAST || f.display_mode == B200GS_DISPLAY_SPLAT) bw = gs_make_bin_word_cov(q0, q1, radf, a, d, f.W, f.H);
                else bw = gs_make_bin_word(q0, q1, f.W, f.H, true);
            }
            bar_arrive(kBarFree + (int)(it & 3u));  // every read of this stage's shared memory is done

            // digit histograms of the emitted keys for the depth sort (saves its histogram kernel);
            // warp-aggregated: depth keys share their top bytes
            if (sort_hist && ballot) {
                const uint32_t key = __float_as_uint(nz);
                const int first = __ffs((int)ballot) - 1;
                const uint32_t key0 = __shfl_sync(0xffffffffu, key, first);
#pragma unroll
                for (int p = 3; p >= 2; p--) {
                    const uint32_t dgt = (key >> (8 * p)) & 0xffu;
                    // the top bytes are usually the same for the whole warp: one vote instead of a match
                    const bool uniform = __all_sync(0xffffffffu, !vis || dgt == ((key0 >> (8 * p)) & 0xffu));
                    if (uniform) {
                        if (lane == first) atomicAdd(&s_hist[p * 256 + dgt], (uint32_t)__popc(ballot));
                    } else if (vis) {
                        const uint32_t peers = __match_any_sync(ballot, dgt);
                        if (lane == __ffs((int)peers) - 1) atomicAdd(&s_hist[p * 256 + dgt], (uint32_t)__popc(peers));
                    }
                }
                // the low bytes are spread: ~30 distinct values among 32 lanes, where MATCH.ANY costs more ADU
                // cycles than 32 fire-and-forget shared-memory atomics cost the (idle) LSU
                if (vis) {
                    atomicAdd(&s_hist[256 + ((key >> 8) & 0xffu)], 1u);
                    atomicAdd(&s_hist[key & 0xffu], 1u);
                }
            }

            pend.valid = true; pend.vis = vis; pend.seq = it; pend.ballot = ballot;
            pend.key = __float_as_uint(nz); pend.index = i; pend.bw = bw; pend.q0 = q0; pend.q1 = q1;
            cur = nxt;
        }
        flush(pend2);
        flush(pend);
    }
    if (sort_hist) {
        __syncthreads();
        for (int i = tid; i < 1024; i += kThreads) {
            const uint32_t cnt = s_hist[i];
            if (cnt) atomicAdd(&sort_hist[i], cnt);
        }
    }
    if (n == 0 && blockIdx.x == 0 && tid == 0) ctrl[GS_CTRL_VISIBLE] = 0;
}

struct DevInfo { int blocks_per_sm[16] = {0}; };   // per layout x {general, FAST}
std::mutex g_mu;
DevInfo g_dev[64];   // attributes and occupancy are per device: a process may hold viewers on several GPUs

template <int SH, int COV, bool FAST>
cudaError_t launch_t(const GsPreprocessArgs& a, const GsFrame& f, const GsModelXf& m, int num_sms, cudaStream_t st) {
    constexpr int RB = 16 + ShBytes<SH>::v + CovBytes<COV>::v;
    constexpr int NSTAGE = 3;
    constexpr size_t smem = (size_t)NSTAGE * kChunk * RB + NSTAGE * 8 + (NSTAGE + 1) * 4 + 32 * 4 + 32 * 4 + 1024 * 4 + 16;
    auto kern = k_preprocess<SH, COV, FAST>;
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return e;
    if (dev < 0 || dev >= 64) return cudaErrorInvalidDevice;
    int blocks_per_sm;
    {
        std::lock_guard<std::mutex> lock(g_mu);
        int& b = g_dev[dev].blocks_per_sm[(SH * 2 + COV) * 2 + (FAST ? 1 : 0)];
        if (b == 0) {
            e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
            if (e != cudaSuccess) return e;
            e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&b, kern, kThreads, smem);
            if (e != cudaSuccess) { b = 0; return e; }
            if (b < 1) b = 1;
        }
        blocks_per_sm = b;
    }
    uint32_t nchunks = (a.n + kChunk - 1) / kChunk;
    uint32_t grid = (uint32_t)(blocks_per_sm * num_sms);
    if (grid > nchunks) grid = nchunks;
    if (grid < 1) grid = 1;
    kern<<<grid, kThreads, smem, st>>>(a.recs, a.n, a.mask, a.selection, a.edits, f, m, a.ctrl, a.lookback, a.epoch,
                                       a.keys, a.idx, a.splats, a.binword, a.sort_hist);
    return cudaGetLastError();
}

}  // namespace

size_t gs_preprocess_lookback_words(uint64_t n) { return (size_t)((n + kChunk - 1) / kChunk) + 1; }

cudaError_t gs_launch_preprocess(const GsPreprocessArgs& a, const GsFrame& f, const GsModelXf& m, int num_sms,
                                 cudaStream_t st) {
    const bool fast = m.identity && f.std_proj && !a.mask && !a.selection && !a.edits && f.query.kind < B200GS_QUERY_RECT &&
                      f.display_mode == B200GS_DISPLAY_SPLAT && f.sh_deg == 3 && !f.no_sh0;
#define GS_PRE_CASE(k, SH, COV) \
    case k: return fast ? launch_t<SH, COV, true>(a, f, m, num_sms, st) : launch_t<SH, COV, false>(a, f, m, num_sms, st)
    switch (a.sh * 2 + a.cov) {
        GS_PRE_CASE(0, 0, 0);
        GS_PRE_CASE(1, 0, 1);
        GS_PRE_CASE(2, 1, 0);
        GS_PRE_CASE(3, 1, 1);
        GS_PRE_CASE(4, 2, 0);
        GS_PRE_CASE(5, 2, 1);
        GS_PRE_CASE(6, 3, 0);
        GS_PRE_CASE(7, 3, 1);
    }
#undef GS_PRE_CASE
    return cudaErrorInvalidValue;
}
