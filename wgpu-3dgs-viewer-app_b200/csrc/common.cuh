// common.cuh — shared declarations of the B200 (sm_100a) 3DGS render core.
//
// Product code.  Nothing here includes or links anything under oracle/.
#pragma once

#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/b200gs.h"

// ---- constants of the path ([CANON] = Inria diff-gaussian-rasterization) ----
#define GS_CULL_XY 1.3f                 // NDC x/y frustum margin
#define GS_CLAMP_XY 1.3f                // Jacobian clamp, x/z and y/z to 1.3 tan(fov/2)
#define GS_LOWPASS 0.3f                 // low-pass added to the cov2d diagonal
#define GS_EXTENT_SIGMA 3.0f            // extent = ceil(3 sqrt(lambda_max))
#define GS_MIN_DISC 0.1f
#define GS_ALPHA_MAX 0.99f
#define GS_ALPHA_MIN (1.0f / 255.0f)
#define GS_T_EPS (1.0f / 1024.0f)       // front-to-back termination threshold
#define GS_FLAT_D2 4.0f                 // Ellipse / Point display modes
#define GS_POINT_RADIUS 1.5f

#define GS_TILE 16                      // compositor tile edge in pixels (one CTA)
#define GS_BIN 32                       // binning tile edge in pixels: a bin is 2 x 2 compositor tiles (its quadrants)
#define GS_QMASK_SHIFT 28               // an entry's sort key is bin id | quadrant mask << 28 (bit q = ty&1 * 2 + tx&1)
#define GS_NUM_SMS_FALLBACK 148

// Per-frame uniforms (CameraPod + GaussianTransformPod + selection pods of the reference,
// scene.rs:785-835) pre-digested on the host so the kernels do no setup arithmetic.
struct GsFrame {
    float V[3][4];  // rows 0..2 of the view matrix
    float P[4][4];  // projection, row-major
    float cam[3];   // camera position in world space
    float W, H;     // viewport in pixels
    float fx, fy;   // P00*W/2, P11*H/2
    float limx, limy;
    float sz2;      // gaussian size²
    uint32_t display_mode, sh_deg, no_sh0;
    b200gs_edit_pod sel_edit;
    float hl[4];
    float bg[4];
    uint32_t tiles_x, tiles_y;   // 16-pixel compositor tiles
    uint32_t bins_x, bins_y;     // 32-pixel binning tiles
    uint32_t std_proj;  // 1: only P00,P11,P22,P23,P32 are non-zero (glam perspective_rh): kernels skip the zero terms
    b200gs_query_pod query;  // selection query tested in the preprocess kernel (rect / brush / texture)
    const uint8_t* query_tex;  // query texture (u8 per pixel, query_tex_w x query_tex_h), sampled when query.kind == TEXTURE
    uint32_t query_tex_w, query_tex_h;
};

// Per-model uniforms (ModelTransformPod, scene.rs:796-802)
struct GsModelXf {
    float R[3][3];
    float t[3];
    float s[3];
    float M[3][3];  // R * diag(s)
    uint32_t identity;  // 1: R = I, s = 1, t = 0 exactly: kernels skip the model transform
};

// Control block of one model in device memory (u32 words)
enum {
    GS_CTRL_TICKET = 0,    // preprocess chunk ticket
    GS_CTRL_VISIBLE = 1,   // V, written by the preprocess kernel
    GS_CTRL_SORT_IN_B = 2,   // 1 = sorted keys/values are in the *_b buffers
    GS_CTRL_WORDS = 16
};

#define GS_LOOKBACK_FLAG_AGG (1u << 30)
#define GS_LOOKBACK_FLAG_INCL (2u << 30)
#define GS_LOOKBACK_VALUE_MASK ((1u << 30) - 1u)

// ------------------------------------------------------------- launch API (csrc/*.cu)
struct GsPreprocessArgs {
    const uint8_t* recs; uint32_t n; uint32_t sh, cov;
    const uint32_t* mask; uint32_t* selection; const b200gs_edit_pod* edits;  // selection is rewritten by a rect/brush query
    uint32_t* ctrl;      // GS_CTRL_WORDS, zeroed before launch
    uint64_t* lookback;  // one status word per 256-Gaussian chunk (epoch-tagged, never cleared)
    uint32_t epoch;
    uint32_t* keys; uint32_t* idx; b200gs_splat* splats;
    uint32_t* binword;    // per compaction slot: bin word of the splat (gs_make_bin_word; consumed by k_bin)
    uint32_t* sort_hist;  // 4 x 256 digit histogram of the emitted keys (zeroed before launch), or null
};
cudaError_t gs_launch_preprocess(const GsPreprocessArgs& a, const GsFrame& f, const GsModelXf& m, int num_sms,
                                 cudaStream_t st);
size_t gs_preprocess_lookback_words(uint64_t n);   // status words of the preprocess kernel for n Gaussians

// K2 (sort.cu): onesweep LSD radix sort of (key,value) u32 pairs, 8-bit digits, CTA-local tiles; n is read from the
// device (*d_n).  Sorts the depth keys.
struct GsSortArgs {
    uint32_t* keys_a; uint32_t* vals_a;   // input, and final output
    uint32_t* keys_b; uint32_t* vals_b;   // scratch (ping-pong)
    const uint32_t* d_n; uint32_t n_max;  // element count on device, and its upper bound
    uint32_t* hist;      // 4 x 256 global digit histogram (zeroed before launch unless prefilled)
    uint64_t* lookback;  // passes x tiles x 256 status words (epoch-tagged, never cleared)
    uint32_t epoch;
    uint32_t* tickets;   // passes words, zeroed before launch
    uint32_t passes;     // 1..4 (bits = 8*passes, starting at bit 0)
    uint32_t* result_in_b; // device flag written by the last pass: 1 = the sorted data is in keys_b/vals_b
    bool hist_prefilled; // histogram already accumulated by the producer
    bool vals_identity;  // the first executed pass synthesises value = input position instead of reading vals_a
    uint32_t vote_mask;  // bit p: pass p finds same-digit peers with ballots (spread digits) instead of MATCH.ANY (concentrated)
};
size_t gs_sort_lookback_words(uint32_t n_max, uint32_t passes);
cudaError_t gs_launch_sort(const GsSortArgs& a, int num_sms, cudaStream_t st);

// K2w (sort_wide.cu): the same sort with 11-bit digits and one thread-block cluster per 32768-key super-tile; sorts
// the bin ids of the binning stage in ONE pass (<= 2048 bins).
#define GS_SORT_DIGIT_BITS 11u
#define GS_SORT_BINS 2048u
__host__ __device__ inline uint32_t gs_sort_passes(uint32_t key_bits) { return (key_bits + GS_SORT_DIGIT_BITS - 1u) / GS_SORT_DIGIT_BITS; }
struct GsSortWideArgs {
    uint32_t* keys_a; uint32_t* vals_a;   // input, and final output
    uint32_t* keys_b; uint32_t* vals_b;   // scratch (ping-pong)
    const uint32_t* d_n; uint32_t n_max;  // element count on device, and its upper bound
    uint32_t* hist;      // passes x 2048 global digit histogram: digit p = (key >> 11 p) & 2047 (zeroed before launch unless prefilled)
    uint64_t* lookback;  // gs_sort_wide_lookback_words(n_max, key_bits) status words (epoch-tagged, never cleared)
    uint32_t epoch;
    uint32_t* tickets;   // passes words, zeroed before launch
    uint32_t key_bits;   // 1..32: keys are < 2^key_bits; passes = ceil(key_bits / 11)
    uint32_t* result_in_b; // device flag written by the last pass: 1 = the sorted data is in keys_b/vals_b
    bool hist_prefilled; // histogram already accumulated by the producer
    bool vals_identity;  // the first executed pass synthesises value = input position instead of reading vals_a
};
size_t gs_sort_wide_lookback_words(uint32_t n_max, uint32_t key_bits);
cudaError_t gs_launch_sort_wide(const GsSortWideArgs& a, int num_sms, cudaStream_t st);
cudaError_t gs_sort_set_cluster(int ctas_per_cluster);   // tuning knob: 8 (default), 4, 2, 1
int gs_sort_get_cluster();
void gs_sort_set_claim(int on);                          // tuning knob: collision-free fast path of the ranking (default off)
int gs_sort_get_claim();
int gs_sort_resident_clusters(int device);               // co-resident clusters of the pass kernel (0 until the first sort)
// Binning: expand depth-sorted splats into (tile, splat) entries in depth order.
struct GsBinArgs {
    const uint32_t* sorted_slot;   // per depth rank: slot of the splat in `splats` (compaction order)
    const uint32_t* sorted_slot_b; // the sort's other buffer, selected when *sorted_in_b != 0
    const uint32_t* sorted_in_b;   // device flag written by the sort (GsSortArgs::result_in_b)
    const b200gs_splat* splats;    // this model's splats (compaction order)
    const uint32_t* binword;       // this model's per-slot bin words (from the preprocess kernel)
    const uint32_t* d_v;           // visible count of this model on device
    uint32_t v_max;
    uint32_t splat_base;           // global id of this model's splat 0 in the frame arena
    uint64_t* lookback;            // per-1024-rank-chunk status words (epoch-tagged, never cleared)
    uint32_t epoch;
    uint32_t* ticket;              // 1 word, zeroed before launch
    const uint32_t* entry_base_in; // entries already emitted by nearer models
    uint32_t* entry_total_out;     // entry_base_in + this model's entries (a different word)
    uint32_t* overflow;            // set to 1 when the capacity is exceeded
    uint32_t* tile_keys; uint32_t* tile_vals; uint32_t capacity;
    uint32_t* tile_count;          // gs_tile_count_words(n_bins) words: replicated per-bin entry counters (all zero between frames)
};
cudaError_t gs_launch_bin(const GsBinArgs& a, const GsFrame& f, int num_sms, cudaStream_t st);
size_t gs_tile_count_words(uint32_t n_tiles);
// per-tile list boundaries + launch order + the tile sort's digit histograms, from the per-tile counters
struct GsTileRangesArgs {
    uint32_t* tile_count;          // replicated per-tile counters (cleared on the way)
    uint32_t* ranges;              // 4 x n_tiles: start, end, launch order, scratch
    uint32_t n_tiles;
    uint32_t* hist; uint32_t passes;   // passes x 256 digit histogram of the tile ids (zeroed before launch)
    unsigned long long* entry_stat;    // += total entries (may be null)
    uint64_t* lookback; uint32_t epoch;   // gs_tile_lookback_words(n_tiles) status words (epoch-tagged, never cleared)
    uint32_t* ticket; uint32_t* done_ctr; uint32_t* buckets;   // 1 + 1 + 256 words, zeroed before launch
};
size_t gs_tile_lookback_words(uint32_t n_tiles);
cudaError_t gs_launch_tile_ranges(const GsTileRangesArgs& a, cudaStream_t st);

struct GsCompositeArgs {
    const uint32_t* tile_keys;      // entries sorted by bin, depth order inside a bin: key = bin | quadrant mask << 28 ...
    const uint32_t* tile_vals;      // ... value = splat id in the frame arena
    const uint32_t* tile_keys_b;    // the bin sort's other buffers, selected when *tile_in_b != 0
    const uint32_t* tile_vals_b;
    const uint32_t* tile_in_b;
    const uint32_t* ranges;         // [bin] = start, [n_bins + bin] = end, [2 n_bins + i] = i-th bin to launch
    const b200gs_splat* splats;     // frame arena
    uint8_t* out; size_t pitch;     // RGBA8
    unsigned long long* evals;      // optional work counters: [0] evaluations, [1] entries staged (may be null)
};
cudaError_t gs_launch_composite(const GsCompositeArgs& a, const GsFrame& f, cudaStream_t st);
cudaError_t gs_launch_query_hits(const GsCompositeArgs& a, const GsFrame& f, uint32_t px, uint32_t py, uint2* out,
                                 uint32_t cap, uint32_t* count, cudaStream_t st);

// Mask evaluation / postprocess (row N2)
cudaError_t gs_launch_eval_mask(const uint8_t* recs, uint32_t n, uint32_t record_bytes, const GsModelXf& m,
                                const b200gs_mask_op* ops_dev, uint32_t n_ops, const b200gs_mask_shape* shapes_dev,
                                const float* shape_rot_dev, uint32_t* words, cudaStream_t st);
cudaError_t gs_launch_postprocess(uint32_t n, const uint32_t* selection, b200gs_edit_pod* edits,
                                  b200gs_edit_pod sel_edit, cudaStream_t st);
cudaError_t gs_launch_fill_default_edits(uint32_t n, b200gs_edit_pod* edits, cudaStream_t st);
cudaError_t gs_launch_paint_query_texture(uint8_t* tex, uint32_t w, uint32_t h, const b200gs_query_pod& stroke, cudaStream_t st);

// ------------------------------------------------------------------ device helpers
#ifdef __CUDACC__
__device__ __forceinline__ uint32_t gs_smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void gs_mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(gs_smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void gs_fence_mbar_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void gs_mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(gs_smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void gs_mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(gs_smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void gs_mbar_wait(uint64_t* bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra DONE_%=;\n"
        "bra WAIT_%=;\n"
        "DONE_%=:\n"
        "}\n" ::"r"(gs_smem_u32(bar)),
        "r"(parity)
        : "memory");
}
// 1-D TMA bulk copy global -> shared (SASS: UBLKCP), completion on an mbarrier
__device__ __forceinline__ void gs_tma_load_1d(void* dst_smem, const void* src_gmem, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     gs_smem_u32(dst_smem)),
                 "l"(src_gmem), "r"(bytes), "r"(gs_smem_u32(bar))
                 : "memory");
}
// ---- epoch-tagged look-back status words -------------------------------------------------
// A status word is (epoch << 32) | (flag << 30) | value.  A word whose epoch differs from the
// launch's epoch reads as "not published", so status arrays never need clearing between
// launches: the host hands every launch a fresh epoch (buffers are zero-filled at allocation
// and epochs start at 1).
__device__ __forceinline__ uint64_t gs_ld_status(const uint64_t* p) {
    uint64_t v;
    asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void gs_st_status(uint64_t* p, uint32_t epoch, uint32_t flag_value) {
    uint64_t v = ((uint64_t)epoch << 32) | flag_value;
    asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ uint32_t gs_status_flag(uint64_t v, uint32_t epoch) {
    return ((uint32_t)(v >> 32) == epoch) ? (((uint32_t)v) >> 30) : 0u;
}

// Decoupled look-back executed by ONE full warp, split in two so that callers can put work
// between publishing their aggregate and needing the prefix.  All lanes return the same value.
// Values saturate at 2^30 - 1 instead of carrying into the flag bits: a saturated total is larger than any buffer
// capacity the host allows (< 2^30), so callers that bound their writes by a capacity flag overflow as usual.
__device__ __forceinline__ uint32_t gs_sat30(uint32_t a, uint32_t b) {
    const uint32_t s = a + b;
    return (s < a || s > GS_LOOKBACK_VALUE_MASK) ? GS_LOOKBACK_VALUE_MASK : s;
}
__device__ __forceinline__ void gs_lookback_publish(uint64_t* status, uint32_t epoch, uint32_t tile, uint32_t aggregate) {
    gs_st_status(&status[tile], epoch, (tile == 0 ? GS_LOOKBACK_FLAG_INCL : GS_LOOKBACK_FLAG_AGG) | gs_sat30(aggregate, 0u));
}
__device__ __forceinline__ uint32_t gs_lookback_resolve(uint64_t* status, uint32_t epoch, uint32_t tile,
                                                        uint32_t aggregate, int lane) {
    // Each round inspects the 64 predecessors p .. p-63 with 2 independent loads per lane (group j holds
    // p-32j-lane).  Measured on the preprocess kernel (444 CTAs in flight): 1 or 2 groups per round 206 us,
    // 4 groups 212 us, 8 groups 227 us — the pipelined callers publish a whole iteration before they resolve,
    // so an inclusive prefix is usually within the first few dozen predecessors and wider rounds only add
    // L2 traffic on the status words.
    if (tile == 0) return 0;
    constexpr int kGroups = 2;
    uint32_t part = 0;   // this lane's share of the sum: ONE warp reduction at the end, none per group
    int64_t p = (int64_t)tile - 1;
    while (true) {
        uint64_t v[kGroups];
#pragma unroll
        for (int j = 0; j < kGroups; j++) {
            const int64_t i = p - 32 * j - lane;
            v[j] = (i >= 0) ? gs_ld_status(&status[i]) : (((uint64_t)epoch << 32) | GS_LOOKBACK_FLAG_INCL);
        }
        bool done = false, stalled = false;
#pragma unroll
        for (int j = 0; j < kGroups; j++) {
            if (!done && !stalled) {
                const uint32_t flag = gs_status_flag(v[j], epoch);
                const uint32_t m_incl = __ballot_sync(0xffffffffu, flag == 2u);
                const uint32_t m_inv = __ballot_sync(0xffffffffu, flag == 0u);
                const uint32_t first = m_incl ? (uint32_t)(__ffs((int)m_incl) - 1) : 32u;
                const uint32_t needed = first >= 31u ? 0xffffffffu : ((2u << first) - 1u);
                if (m_inv & needed) stalled = true;  // a needed predecessor has not published yet: retry from here
                else {
                    if ((needed >> lane) & 1u) part = gs_sat30(part, (uint32_t)v[j] & GS_LOOKBACK_VALUE_MASK);
                    if (first < 32u) done = true;
                    else p -= 32;
                }
            }
        }
        if (done) break;
    }
    uint32_t excl = part;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) excl = gs_sat30(excl, __shfl_xor_sync(0xffffffffu, excl, o));
    if (lane == 0) gs_st_status(&status[tile], epoch, GS_LOOKBACK_FLAG_INCL | gs_sat30(excl, gs_sat30(aggregate, 0u)));
    return excl;
}
__device__ __forceinline__ uint32_t gs_lookback_warp(uint64_t* status, uint32_t epoch, uint32_t tile,
                                                     uint32_t aggregate, int lane) {
    if (lane == 0) gs_lookback_publish(status, epoch, tile, aggregate);
    return gs_lookback_resolve(status, epoch, tile, aggregate, lane);
}

// ---- selection query shapes ---------------------------------------------------------------
// gs::QueryToolset rect / brush (reference src/tab/scene.rs:758-791, 1224-1263): is the point (viewport pixels,
// top-left origin, pixel i covers [i, i+1)) inside the rectangle / within `radius` of the brush segment?
// EXACT class: single-rounding operations, same sequence as the oracle (callers are compiled with -fmad=false).
__device__ __forceinline__ bool gs_query_shape_hit(const b200gs_query_pod& q, float sx, float sy) {
    if (q.kind == B200GS_QUERY_RECT) return sx >= q.p0[0] && sx <= q.p1[0] && sy >= q.p0[1] && sy <= q.p1[1];
    const float vx = q.p1[0] - q.p0[0], vy = q.p1[1] - q.p0[1];
    const float wx = sx - q.p0[0], wy = sy - q.p0[1];
    const float vv = vx * vx + vy * vy;
    float t = vv > 0.0f ? (wx * vx + wy * vy) / vv : 0.0f;
    t = fminf(1.0f, fmaxf(0.0f, t));
    const float dx = wx - t * vx, dy = wy - t * vy;
    return dx * dx + dy * dy <= q.radius * q.radius;
}

// ---- footprint threshold ------------------------------------------------------------------
// A splat contributes to a pixel only if alpha = o*exp(-q/2) >= 1/255, i.e. q <= tau with
// q = a dx^2 + 2 b dx dy + c dy^2 and tau = 2 ln(255 o) (flat display modes: tau = GS_FLAT_D2).
// footprint threshold (with slack so that rounding can never cull a contributing pixel)
// single-instruction MUFU forms (flush-to-zero, no denormal fix-up code around them): both are deterministic, so the
// translation units that rebuild a rectangle (preprocess: bin word, binning: enumeration of a huge splat) agree
__device__ __forceinline__ float gs_rsqrt_approx(float x) {
    float y;
    asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
__device__ __forceinline__ float gs_lg2_approx(float x) {
    float y;
    asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
__device__ __forceinline__ float gs_footprint_tau(float opacity, bool flat) {
    if (!(opacity * 255.0f >= 1.0f)) return -1.0f;  // alpha < 1/255 everywhere
    // explicit single-rounding ops: this value feeds the candidate tile rectangle, which must come out
    // identical in translation units compiled with and without -fmad.  2 ln x = (2 ln 2) lg2 x
    const float t = flat ? GS_FLAT_D2 : __fmul_rn(1.3862943611198906f, gs_lg2_approx(__fmul_rn(opacity, 255.0f)));
    return __fadd_rn(__fmul_rn(t, 1.001f), 0.01f);
}
// ---- candidate tile rectangle of a projected splat (shared by preprocess and binning) --------
struct GsCand {
    float mx, my, a, b, c, tau;   // ellipse
    float fx0, fx1, fy0, fy1;     // pixel bounds of the extent square clipped to the viewport
    uint32_t tx0, ty0, nx, ny;    // candidate tile rectangle
};

// candidate tile rectangle of a projected splat; false if it cannot touch anything.  The pixel
// range is the extent square (same expression as the compositor / the oracle, exact in float)
// intersected with a conservative axis-aligned box of the footprint ellipse q <= tau (half-widths
// sqrt(tau * cov_xx), sqrt(tau * cov_yy), inflated): tiles outside it cannot be touched.
// Deterministic function of the STORED record, so the preprocess kernel (bin word: candidate count of a
// huge splat) and the binning kernel (its enumeration) agree exactly.
// pixel bounds + tile rectangle from the centre and the half-widths of the candidate box
__device__ __forceinline__ bool gs_rect_from_box(float rx, float ry, float W, float H, GsCand& c) {
    // (rx, ry <= r and rounding is monotonic, so this box never reaches beyond the extent square itself)
    c.fx0 = ceilf(__fsub_rn(c.mx, rx)); c.fx1 = floorf(__fadd_rn(c.mx, rx));
    c.fy0 = ceilf(__fsub_rn(c.my, ry)); c.fy1 = floorf(__fadd_rn(c.my, ry));
    if (c.fx0 < 0.0f) c.fx0 = 0.0f;
    if (c.fy0 < 0.0f) c.fy0 = 0.0f;
    if (c.fx1 > W - 1.0f) c.fx1 = W - 1.0f;
    if (c.fy1 > H - 1.0f) c.fy1 = H - 1.0f;
    if (!(c.fx0 <= c.fx1 && c.fy0 <= c.fy1)) return false;
    c.tx0 = (uint32_t)c.fx0 / GS_TILE;
    c.ty0 = (uint32_t)c.fy0 / GS_TILE;
    c.nx = (uint32_t)c.fx1 / GS_TILE - c.tx0 + 1;
    c.ny = (uint32_t)c.fy1 / GS_TILE - c.ty0 + 1;
    return true;
}
// half-width of the footprint box along one axis: min(r, sqrt(v) inflated), v = tau * covariance of that axis.
// sqrt(x) as x * rsqrt(x) (MUFU.RSQ: ~1 ulp, covered by the inflation; a NaN from x = 0 or inf leaves the extent
// radius in place).  Explicit single-rounding ops, see gs_footprint_tau.
__device__ __forceinline__ float gs_box_halfwidth(float r, float v) {
    return fminf(r, __fadd_rn(__fmul_rn(__fmul_rn(v, gs_rsqrt_approx(v)), 1.002f), 0.02f));
}
__device__ __forceinline__ bool gs_make_rect(const uint4& q0, const uint4& q1, float W, float H, bool flat, GsCand& c) {
    const uint32_t radius = q0.z & 0xffffu;
    if (radius == 0) return false;
    c.mx = __uint_as_float(q0.x);
    c.my = __uint_as_float(q0.y);
    const float op = __half2float(__ushort_as_half((unsigned short)(q0.z >> 16)));
    c.tau = gs_footprint_tau(op, flat);
    if (c.tau < 0.0f) return false;
    c.a = __uint_as_float(q1.x);
    c.b = __uint_as_float(q1.y);
    c.c = __uint_as_float(q1.z);
    const float r = (float)radius;
    float rx = r, ry = r;
    const float det = __fsub_rn(__fmul_rn(c.a, c.c), __fmul_rn(c.b, c.b));
    if (det > 0.0f) {
        const float k = __fdividef(c.tau, det);
        rx = gs_box_halfwidth(r, __fmul_rn(k, c.c));
        ry = gs_box_halfwidth(r, __fmul_rn(k, c.a));
    }
    return gs_rect_from_box(rx, ry, W, H, c);
}

// ---- bin word: what the binning kernel needs to know about a splat, one u32 per compaction slot ----
// Written by the preprocess kernel (which has the projected splat in registers), so that the binning
// kernel expands splats from a 4-byte L2-resident gather instead of a 32-byte one.  The candidate rectangle
// (gs_make_rect: extent square ∩ bounding box of the alpha >= 1/255 footprint) is kept whole and is spelled in
// 16-pixel COMPOSITOR tiles; the binning kernel derives from it the 32-pixel bins the splat falls in (one entry
// each) and, per entry, which of the bin's four quadrants (compositor tiles) the rectangle reaches — the
// compositor CTA of a quadrant stages only the entries that carry its bit.
//   0                  : touches nothing
//   bit 31 clear       : bit 30 set; bits 0..9 tx0, 10..19 ty0 (first tile column / row), bits 20..24 nx - 1,
//                        bits 25..29 ny - 1 (nx, ny <= 32 tiles)
//   bit 31 set (huge)  : bits 0..29 number of BINS of the rectangle; the binning kernel rebuilds the rectangle from the
//                        stored splat
#define GS_BIN_INLINE 4u
#define GS_BIN_VALID 0x40000000u
#define GS_BIN_HUGE 0x80000000u
// bins covered by the tile range [t0, t0 + n - 1] along one axis
__device__ __forceinline__ uint32_t gs_bins_of_tiles(uint32_t t0, uint32_t n) { return ((t0 + n - 1u) >> 1) - (t0 >> 1) + 1u; }
// quadrant mask of bin (bx, by) for the tile rectangle [tx0, tx1] x [ty0, ty1] (the bin is known to intersect it)
__device__ __forceinline__ uint32_t gs_quadrant_mask(uint32_t bx, uint32_t by, uint32_t tx0, uint32_t tx1, uint32_t ty0, uint32_t ty1) {
    const uint32_t col = (2u * bx >= tx0 ? 1u : 0u) | (2u * bx + 1u <= tx1 ? 2u : 0u);
    const uint32_t row = (2u * by >= ty0 ? 1u : 0u) | (2u * by + 1u <= ty1 ? 2u : 0u);
    return ((row & 1u) ? col : 0u) | ((row & 2u) ? col << 2 : 0u);
}
__device__ __forceinline__ uint32_t gs_encode_bin_word(const GsCand& cd) {
    if (cd.nx <= 32u && cd.ny <= 32u && cd.tx0 < 1024u && cd.ty0 < 1024u)
        return GS_BIN_VALID | cd.tx0 | (cd.ty0 << 10) | ((cd.nx - 1u) << 20) | ((cd.ny - 1u) << 25);
    return GS_BIN_HUGE | (gs_bins_of_tiles(cd.tx0, cd.nx) * gs_bins_of_tiles(cd.ty0, cd.ny));
}
__device__ __forceinline__ uint32_t gs_make_bin_word(const uint4& q0, const uint4& q1, float W, float H, bool flat) {
    GsCand cd;
    if (!gs_make_rect(q0, q1, W, H, flat, cd)) return 0u;
    return gs_encode_bin_word(cd);
}
// The same word for the Splat display mode from what the preprocess kernel holds in registers: the footprint box
// half-widths are sqrt(tau * cov_xx), sqrt(tau * cov_yy) with the 2-D covariance itself (the stored conic is its
// inverse: tau * conic_c / det(conic) is the same number up to rounding, far inside the box's 0.2 % + 0.02 px
// inflation).  A regular word spells its rectangle out, so it need not be reproducible from the record; a HUGE one
// is a count the binning kernel re-derives from the stored record, so that case defers to gs_make_bin_word.
__device__ __forceinline__ uint32_t gs_make_bin_word_cov(const uint4& q0, const uint4& q1, float radf, float cov_xx, float cov_yy,
                                                         float W, float H) {
    if (!(radf > 0.0f)) return 0u;
    GsCand cd;
    cd.mx = __uint_as_float(q0.x);
    cd.my = __uint_as_float(q0.y);
    const float op = __half2float(__ushort_as_half((unsigned short)(q0.z >> 16)));   // the stored (f16) opacity
    const float tau = gs_footprint_tau(op, false);
    if (tau < 0.0f) return 0u;
    if (!gs_rect_from_box(gs_box_halfwidth(radf, __fmul_rn(tau, cov_xx)), gs_box_halfwidth(radf, __fmul_rn(tau, cov_yy)), W, H, cd))
        return 0u;
    if (cd.nx > 32u || cd.ny > 32u) return gs_make_bin_word(q0, q1, W, H, false);
    return gs_encode_bin_word(cd);
}

#endif
