// api.cu — the C ABI of include/b200gs.h: viewer / model handles, device memory, stream,
// uploads, downloads and the per-frame launch sequence.  Mirrors the resource model of the
// reference's gs::MultiModelViewer (src/tab/scene.rs:1969-1980, 2102-2139): one viewer = one
// CUDA stream on one device; calls on a viewer are externally serialised (scene.rs:1926-1930).
//
// There is no CPU fallback: every compute entry point needs a CUDA device.
#include <algorithm>
#include <atomic>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>

#include "common.cuh"
#include "host_api.h"

// ---------------------------------------------------------------------------- errors
static thread_local std::string g_last_error;
void gs_set_error(const char* fmt, ...) {
    char buf[1024];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    g_last_error = buf;
}
extern "C" const char* b200gs_last_error(void) { return g_last_error.c_str(); }
extern "C" const char* b200gs_version(void) { return "b200gs 0.1 (sm_100a)"; }

#define CK(call)                                                                                        \
    do {                                                                                                \
        cudaError_t e__ = (call);                                                                       \
        if (e__ != cudaSuccess) {                                                                       \
            gs_set_error("%s:%d: %s -> %s", __FILE__, __LINE__, #call, cudaGetErrorString(e__));        \
            return e__ == cudaErrorMemoryAllocation ? B200GS_ERR_OOM : B200GS_ERR_CUDA;                 \
        }                                                                                               \
    } while (0)
#define REQUIRE(cond, msg)                          \
    do {                                            \
        if (!(cond)) {                              \
            gs_set_error("%s: %s", __func__, msg);  \
            return B200GS_ERR_INVALID;              \
        }                                           \
    } while (0)

// ---------------------------------------------------------------------------- handles
enum {  // viewer control block (u32 words), zeroed at the start of every render
    VC_BIN_TICKET = 0,     // 64 words: chunk ticket per model in the frame
    VC_ENTRY_TOTAL = 128,  // 65 words: running (tile, splat) entry count after each model
    VC_TSORT_TICKET = 196, // 3 words
    VC_TSORT_IN_B = 200,   // 1 = tile-sorted entries are in the *_b buffers
    VC_TILE_TICKET = 201,  // tile-finish kernel: chunk ticket, finished-CTA counter
    VC_TILE_DONE = 202,
    VC_TILE_BUCKETS = 256, // 256: tiles per list-length bucket (launch order)
    VC_TSORT_HIST = 512,   // 2 x 2048: digit histograms of the tile-sort keys (accumulated by the tile-finish kernel)
    VC_WORDS = 512 + 2 * 2048
};
enum {  // model control block layout
    MC_CTRL = 0,               // GS_CTRL_WORDS
    MC_SORT_TICKET = 16,       // 4
    MC_SORT_HIST = 32,         // 4 x 256: digit histograms of the depth keys (accumulated by the preprocess kernel)
    MC_WORDS = 32 + 1024
};
constexpr uint32_t kMaxModelsPerFrame = 64;

// Packed records of a model: reference counted, so that several viewers on one device can render the same
// resident copy (the reference's buffers are ref-counted clones, src/tab/scene.rs:641, 648).
struct GsRecStore {
    uint8_t* p = nullptr;
    std::atomic<int> refs{1};
    int device = 0;
};

struct b200gs_model {
    b200gs_viewer* v = nullptr;
    std::string key;
    uint64_t cap = 0;
    GsRecStore* store = nullptr;
    uint8_t* recs = nullptr;   // = store->p
    uint32_t* mask = nullptr;
    uint32_t* selection = nullptr;
    b200gs_edit_pod* edits = nullptr;
    float pos[3] = {0, 0, 0}, quat[4] = {0, 0, 0, 1}, scale[3] = {1, 1, 1};
    uint32_t* ctrl = nullptr;
    uint32_t *keys_a = nullptr, *vals_a = nullptr, *keys_b = nullptr, *vals_b = nullptr, *idx = nullptr, *binword = nullptr;
    uint64_t *lb_pre = nullptr, *lb_sort = nullptr;
    uint64_t arena_offset = 0;
    bool preprocessed = false, sorted = false;
};

struct b200gs_viewer {
    int device = 0;
    cudaStream_t stream = nullptr;
    int num_sms = GS_NUM_SMS_FALLBACK;
    uint32_t sh = 0, cov = 0, rb = 0;
    uint32_t W = 1, H = 1;
    float view[16], proj[16], size[2];
    float gsize = 1.0f;
    uint32_t display_mode = 0, sh_deg = 3, no_sh0 = 0;
    b200gs_edit_pod sel_edit;
    float hl[4] = {0, 0, 0, 0}, bg[4] = {0, 0, 0, 0};
    b200gs_query_pod query;
    std::vector<b200gs_model*> models;
    b200gs_splat* arena = nullptr;
    uint64_t arena_cap = 0;
    bool layout_dirty = true;
    uint32_t *tk_a = nullptr, *tv_a = nullptr, *tk_b = nullptr, *tv_b = nullptr;
    uint64_t entry_cap = 0, entry_cap_user = 0;
    uint64_t *lb_bin = nullptr, *lb_tsort = nullptr, *lb_tiles = nullptr;
    uint64_t lb_bin_words = 0;
    uint32_t* ranges = nullptr;
    uint32_t* tile_count = nullptr;        // per tile: entries written by the binning kernel (cleared by the tile scan)
    uint32_t ranges_tiles = 0;
    unsigned long long* stats = nullptr;   // [0] evals, [1] staged entries, [2] tile entries (per frame)
    cudaEvent_t ev_bin[3] = {};            // render: start, after binning + bin sort, after compositing
    uint32_t* vctrl = nullptr;
    uint8_t* image = nullptr;  // internal RGBA8 targets for render_frame_host (2 slots, image + image_bytes)
    size_t image_bytes = 0;
    cudaStream_t copy_stream = nullptr;
    cudaEvent_t ev_rendered[2] = {nullptr, nullptr}, ev_copied[2] = {nullptr, nullptr};
    uint32_t ring_head = 0, ring_pending = 0;
    uint64_t launches = 0;     // kernels launched by this viewer (b200gs_launch_count)
    uint32_t epoch = 0;
    bool timing = false, count_evals = false, rendered = false;
    cudaEvent_t ev[6] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
    b200gs_timings last = {};
    uint32_t* h_small = nullptr;  // pinned scratch (words 0..15: downloads; 32..33: overflow flag of the two host-frame slots)
    // pinned staging ring of the uploads: a host buffer handed to an upload call is copied (or packed) into a slot
    // before the call returns, so uploads enqueue and return without synchronising the stream
    uint8_t* stage_buf[2] = {nullptr, nullptr};
    cudaEvent_t stage_done[2] = {nullptr, nullptr};
    bool stage_used[2] = {false, false};
    uint32_t stage_next = 0;
    uint8_t* mask_scratch = nullptr;   // device: postfix ops | shapes | shape rotations of the last eval_mask
    uint8_t* query_tex = nullptr;      // device: selection query texture (u8 per pixel), or null
    uint32_t query_tex_w = 0, query_tex_h = 0;
};
constexpr size_t kStageBytes = 8u << 20;
constexpr size_t kMaskScratchBytes = 64 * sizeof(b200gs_mask_op) + 64 * sizeof(b200gs_mask_shape) + 64 * 36;

static void identity16(float* m) {
    memset(m, 0, 64);
    m[0] = m[5] = m[10] = m[15] = 1.0f;
}

static void quat_to_mat3(const float q[4], float R[3][3]) {  // glam Mat3::from_quat
    float x = q[0], y = q[1], z = q[2], w = q[3];
    float x2 = x + x, y2 = y + y, z2 = z + z;
    float xx = x * x2, xy = x * y2, xz = x * z2;
    float yy = y * y2, yz = y * z2, zz = z * z2;
    float wx = w * x2, wy = w * y2, wz = w * z2;
    R[0][0] = 1.0f - (yy + zz); R[0][1] = xy - wz;          R[0][2] = xz + wy;
    R[1][0] = xy + wz;          R[1][1] = 1.0f - (xx + zz); R[1][2] = yz - wx;
    R[2][0] = xz - wy;          R[2][1] = yz + wx;          R[2][2] = 1.0f - (xx + yy);
}

static GsFrame make_frame(const b200gs_viewer* v) {
    GsFrame f;
    memset(&f, 0, sizeof f);
    for (int r = 0; r < 3; r++)
        for (int c = 0; c < 4; c++) f.V[r][c] = v->view[c * 4 + r];
    for (int r = 0; r < 4; r++)
        for (int c = 0; c < 4; c++) f.P[r][c] = v->proj[c * 4 + r];
    for (int a = 0; a < 3; a++) f.cam[a] = -(f.V[0][a] * f.V[0][3] + f.V[1][a] * f.V[1][3] + f.V[2][a] * f.V[2][3]);
    f.W = v->size[0];
    f.H = v->size[1];
    f.fx = f.P[0][0] * f.W * 0.5f;
    f.fy = f.P[1][1] * f.H * 0.5f;
    f.limx = GS_CLAMP_XY / f.P[0][0];
    f.limy = GS_CLAMP_XY / f.P[1][1];
    f.sz2 = v->gsize * v->gsize;
    f.display_mode = v->display_mode;
    f.sh_deg = v->sh_deg;
    f.no_sh0 = v->no_sh0;
    f.sel_edit = v->sel_edit;
    memcpy(f.hl, v->hl, 16);
    memcpy(f.bg, v->bg, 16);
    f.tiles_x = (v->W + GS_TILE - 1) / GS_TILE;
    f.tiles_y = (v->H + GS_TILE - 1) / GS_TILE;
    f.bins_x = (v->W + GS_BIN - 1) / GS_BIN;
    f.bins_y = (v->H + GS_BIN - 1) / GS_BIN;
    f.query = v->query;
    f.query_tex = v->query_tex;
    f.query_tex_w = v->query_tex_w;
    f.query_tex_h = v->query_tex_h;
    const float (*P)[4] = f.P;
    f.std_proj = (P[0][1] == 0.0f && P[0][2] == 0.0f && P[0][3] == 0.0f && P[1][0] == 0.0f && P[1][2] == 0.0f &&
                  P[1][3] == 0.0f && P[2][0] == 0.0f && P[2][1] == 0.0f && P[3][0] == 0.0f && P[3][1] == 0.0f &&
                  P[3][3] == 0.0f) ? 1u : 0u;
    return f;
}

static GsModelXf make_xf(const b200gs_model* m) {
    GsModelXf x;
    quat_to_mat3(m->quat, x.R);
    for (int a = 0; a < 3; a++) { x.t[a] = m->pos[a]; x.s[a] = m->scale[a]; }
    for (int a = 0; a < 3; a++)
        for (int b = 0; b < 3; b++) x.M[a][b] = x.R[a][b] * x.s[b];
    x.identity = 1u;
    for (int a = 0; a < 3; a++) {
        if (x.t[a] != 0.0f || x.s[a] != 1.0f) x.identity = 0u;
        for (int b = 0; b < 3; b++)
            if (x.R[a][b] != (a == b ? 1.0f : 0.0f)) x.identity = 0u;
    }
    return x;
}

static int set_device(const b200gs_viewer* v) {
    CK(cudaSetDevice(v->device));
    return B200GS_OK;
}

template <typename T>
static int dev_alloc(T** p, size_t count, bool zero, cudaStream_t st) {
    *p = nullptr;
    size_t bytes = std::max<size_t>(count, 1) * sizeof(T);
    CK(cudaMalloc((void**)p, bytes));
    if (zero) CK(cudaMemsetAsync(*p, 0, bytes, st));
    return B200GS_OK;
}
#define TRY(x)                        \
    do {                              \
        int rc__ = (x);               \
        if (rc__ != B200GS_OK) return rc__; \
    } while (0)

// (re)build the viewer-level buffers whose size depends on the set of models / the viewport
static int ensure_frame_buffers(b200gs_viewer* v) {
    const uint32_t n_tiles = ((v->W + GS_BIN - 1) / GS_BIN) * ((v->H + GS_BIN - 1) / GS_BIN);   // 32-pixel bins
    if (v->ranges_tiles < n_tiles) {
        CK(cudaStreamSynchronize(v->stream));
        if (v->ranges) CK(cudaFree(v->ranges));
        TRY(dev_alloc(&v->ranges, (size_t)n_tiles * 4, true, v->stream));
        if (v->lb_tiles) CK(cudaFree(v->lb_tiles));
        TRY(dev_alloc(&v->lb_tiles, gs_tile_lookback_words(n_tiles), true, v->stream));
        if (v->tile_count) CK(cudaFree(v->tile_count));
        TRY(dev_alloc(&v->tile_count, gs_tile_count_words(n_tiles), true, v->stream));
        v->ranges_tiles = n_tiles;
    }
    const size_t img = (size_t)v->W * v->H * 4;
    if (v->image_bytes < img) {
        CK(cudaStreamSynchronize(v->stream));
        CK(cudaStreamSynchronize(v->copy_stream));   // a pending host frame may still be reading the old target
        if (v->image) CK(cudaFree(v->image));
        TRY(dev_alloc(&v->image, img * 3, true, v->stream));   // two host-frame slots + the hit query's scratch target
        v->image_bytes = img;
    }
    if (!v->layout_dirty) return B200GS_OK;
    CK(cudaStreamSynchronize(v->stream));
    uint64_t total = 0, maxcap = 0;
    for (auto* m : v->models) {
        m->arena_offset = total;
        total += m->cap;
        maxcap = std::max(maxcap, m->cap);
        m->preprocessed = m->sorted = false;
    }
    if (total > v->arena_cap) {
        if (v->arena) CK(cudaFree(v->arena));
        TRY(dev_alloc(&v->arena, total, false, v->stream));
        v->arena_cap = total;
    }
    uint64_t want = v->entry_cap_user ? v->entry_cap_user : std::max<uint64_t>(total * 8, 1u << 20);
    want = std::min<uint64_t>(want, 0x3fffffffull);
    if (want != v->entry_cap) {
        for (uint32_t** p : {&v->tk_a, &v->tv_a, &v->tk_b, &v->tv_b}) {
            if (*p) CK(cudaFree(*p));
            TRY(dev_alloc(p, want, false, v->stream));
        }
        if (v->lb_tsort) CK(cudaFree(v->lb_tsort));
        TRY(dev_alloc(&v->lb_tsort, gs_sort_lookback_words((uint32_t)want, 3), true, v->stream));
        v->entry_cap = want;
    }
    uint64_t lbw = (maxcap + 1023) / 1024 + 1;
    if (lbw > v->lb_bin_words) {
        if (v->lb_bin) CK(cudaFree(v->lb_bin));
        TRY(dev_alloc(&v->lb_bin, lbw * kMaxModelsPerFrame, true, v->stream));
        v->lb_bin_words = lbw;
    }
    v->layout_dirty = false;
    return B200GS_OK;
}

// ---------------------------------------------------------------------------- library
extern "C" int b200gs_device_count(int* out) {
    REQUIRE(out, "null out");
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess) { (void)cudaGetLastError(); n = 0; }
    *out = n;
    return B200GS_OK;
}

extern "C" uint32_t b200gs_record_bytes(uint32_t sh, uint32_t cov3d) { return gs_record_bytes(sh, cov3d); }

// process-wide tuning knobs (set before viewers are created) and read-only facts for the bench / profiles
extern "C" int b200gs_set_tuning(const char* name, int64_t value) {
    REQUIRE(name, "null name");
    if (!strcmp(name, "sort.cluster")) {
        REQUIRE(gs_sort_set_cluster((int)value) == cudaSuccess, "sort.cluster must be 1, 2, 4 or 8");
        return B200GS_OK;
    }
    if (!strcmp(name, "sort.claim")) { gs_sort_set_claim((int)value); return B200GS_OK; }
    REQUIRE(false, "unknown tuning knob");
}
extern "C" int b200gs_get_info(b200gs_viewer* v, const char* name, int64_t* out) {
    REQUIRE(name && out, "null argument");
    if (!strcmp(name, "sort.cluster")) { *out = gs_sort_get_cluster(); return B200GS_OK; }
    if (!strcmp(name, "sort.claim")) { *out = gs_sort_get_claim(); return B200GS_OK; }
    if (!strcmp(name, "sort.resident_clusters")) { REQUIRE(v, "null viewer"); *out = gs_sort_resident_clusters(v->device); return B200GS_OK; }
    if (!strcmp(name, "num_sms")) { REQUIRE(v, "null viewer"); *out = v->num_sms; return B200GS_OK; }
    REQUIRE(false, "unknown info name");
}

// ---------------------------------------------------------------------------- viewer
extern "C" int b200gs_viewer_create(int device, uint32_t sh, uint32_t cov3d, uint32_t width, uint32_t height,
                                    b200gs_viewer** out) {
    REQUIRE(out, "null out");
    *out = nullptr;
    REQUIRE(gs_record_bytes(sh, cov3d) != 0, "invalid layout");
    REQUIRE(width >= 1 && height >= 1 && width <= 16384 && height <= 16384, "invalid size");
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess || n == 0) {
        (void)cudaGetLastError();
        gs_set_error("no CUDA device: the B200 render core has no CPU fallback");
        return B200GS_ERR_CUDA;
    }
    REQUIRE(device >= 0 && device < n, "device out of range");
    CK(cudaSetDevice(device));
    cudaDeviceProp prop;
    CK(cudaGetDeviceProperties(&prop, device));
    if (prop.major < 10) {
        gs_set_error("device %d is sm_%d%d; this library is built for sm_100a only", device, prop.major, prop.minor);
        return B200GS_ERR_CUDA;
    }
    auto* v = new b200gs_viewer();
    v->device = device;
    v->num_sms = prop.multiProcessorCount;
    v->sh = sh; v->cov = cov3d; v->rb = gs_record_bytes(sh, cov3d);
    v->W = width; v->H = height;
    identity16(v->view);
    identity16(v->proj);
    v->size[0] = (float)width; v->size[1] = (float)height;
    v->sel_edit = b200gs_edit_pod{0, {0.0f, 1.0f, 1.0f}, 0.0f, 0.0f, 1.0f, 1.0f};
    memset(&v->query, 0, sizeof v->query);
    cudaError_t e = cudaStreamCreateWithFlags(&v->stream, cudaStreamNonBlocking);
    if (e == cudaSuccess) e = cudaMalloc((void**)&v->vctrl, VC_WORDS * 4);
    if (e == cudaSuccess) e = cudaMemsetAsync(v->vctrl, 0, VC_WORDS * 4, v->stream);
    if (e == cudaSuccess) e = cudaMallocHost((void**)&v->h_small, 4096);
    if (e == cudaSuccess) e = cudaMalloc((void**)&v->stats, 64);
    if (e == cudaSuccess) e = cudaMemsetAsync(v->stats, 0, 64, v->stream);
    for (int j = 0; j < 3 && e == cudaSuccess; j++) e = cudaEventCreate(&v->ev_bin[j]);
    if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&v->copy_stream, cudaStreamNonBlocking);
    for (int i = 0; i < 2 && e == cudaSuccess; i++) {
        e = cudaEventCreateWithFlags(&v->ev_rendered[i], cudaEventDisableTiming);
        if (e == cudaSuccess) e = cudaEventCreateWithFlags(&v->ev_copied[i], cudaEventDisableTiming);
    }
    for (int i = 0; i < 6 && e == cudaSuccess; i++) e = cudaEventCreate(&v->ev[i]);
    if (e != cudaSuccess) {
        gs_set_error("viewer_create: %s", cudaGetErrorString(e));
        b200gs_viewer_destroy(v);
        return B200GS_ERR_CUDA;
    }
    *out = v;
    return B200GS_OK;
}

static void free_model(b200gs_model* m) {
    if (m->store && m->store->refs.fetch_sub(1) == 1) {
        if (m->store->p) cudaFree(m->store->p);
        delete m->store;
    }
    void* ps[] = {m->mask, m->selection, m->edits, m->ctrl, m->keys_a, m->vals_a, m->keys_b, m->vals_b, m->idx, m->binword,
                  m->lb_pre, m->lb_sort};
    for (void* p : ps)
        if (p) cudaFree(p);
    delete m;
}

extern "C" int b200gs_viewer_destroy(b200gs_viewer* v) {
    if (!v) return B200GS_OK;
    cudaSetDevice(v->device);
    if (v->stream) cudaStreamSynchronize(v->stream);
    for (auto* m : v->models) free_model(m);
    void* ps[] = {v->arena, v->tk_a, v->tv_a, v->tk_b, v->tv_b, v->lb_bin, v->lb_tsort, v->lb_tiles, v->tile_count,
                  v->ranges, v->vctrl, v->image, v->stats};
    for (void* p : ps)
        if (p) cudaFree(p);
    if (v->h_small) cudaFreeHost(v->h_small);
    for (int i = 0; i < 2; i++) {
        if (v->stage_buf[i]) cudaFreeHost(v->stage_buf[i]);
        if (v->stage_done[i]) cudaEventDestroy(v->stage_done[i]);
    }
    if (v->mask_scratch) cudaFree(v->mask_scratch);
    if (v->query_tex) cudaFree(v->query_tex);
    for (auto& e : v->ev)
        if (e) cudaEventDestroy(e);
    for (auto& e : v->ev_bin)
        if (e) cudaEventDestroy(e);
    for (int i = 0; i < 2; i++) {
        if (v->ev_rendered[i]) cudaEventDestroy(v->ev_rendered[i]);
        if (v->ev_copied[i]) cudaEventDestroy(v->ev_copied[i]);
    }
    if (v->copy_stream) { cudaStreamSynchronize(v->copy_stream); cudaStreamDestroy(v->copy_stream); }
    if (v->stream) cudaStreamDestroy(v->stream);
    delete v;
    return B200GS_OK;
}

// the projected splats (and their bin words) are computed for one viewport / display mode:
// changing either makes the preprocessed state stale
static void invalidate_models(b200gs_viewer* v) {
    for (auto* m : v->models) m->preprocessed = m->sorted = false;
    v->rendered = false;
}

extern "C" int b200gs_resize(b200gs_viewer* v, uint32_t width, uint32_t height) {
    REQUIRE(v, "null viewer");
    REQUIRE(width >= 1 && height >= 1 && width <= 16384 && height <= 16384, "invalid size");
    if (v->W != width || v->H != height) invalidate_models(v);
    v->W = width; v->H = height;
    v->size[0] = (float)width; v->size[1] = (float)height;
    return B200GS_OK;
}

extern "C" int b200gs_set_camera(b200gs_viewer* v, const float view[16], const float proj[16], const float size[2]) {
    REQUIRE(v && view && proj, "null argument");
    memcpy(v->view, view, 64);
    memcpy(v->proj, proj, 64);
    if (size) {
        REQUIRE(size[0] >= 1.0f && size[1] >= 1.0f && size[0] <= 16384.0f && size[1] <= 16384.0f, "invalid size");
        if (v->W != (uint32_t)size[0] || v->H != (uint32_t)size[1]) invalidate_models(v);
        v->size[0] = size[0]; v->size[1] = size[1];
        v->W = (uint32_t)size[0]; v->H = (uint32_t)size[1];
    }
    return B200GS_OK;
}

extern "C" int b200gs_set_gaussian_transform(b200gs_viewer* v, float size, uint32_t display_mode, uint32_t sh_deg,
                                             uint32_t no_sh0) {
    REQUIRE(v, "null viewer");
    REQUIRE(display_mode <= 2, "display_mode out of range");
    REQUIRE(sh_deg <= 3, "sh_deg out of range (gs::GaussianShDegree::new)");
    if (v->display_mode != display_mode) invalidate_models(v);
    v->gsize = size; v->display_mode = display_mode; v->sh_deg = sh_deg; v->no_sh0 = no_sh0 ? 1 : 0;
    return B200GS_OK;
}
extern "C" int b200gs_set_selection_edit(b200gs_viewer* v, const b200gs_edit_pod* pod) {
    REQUIRE(v && pod, "null argument");
    v->sel_edit = *pod;
    return B200GS_OK;
}
extern "C" int b200gs_set_selection_highlight(b200gs_viewer* v, const float rgba[4]) {
    REQUIRE(v && rgba, "null argument");
    memcpy(v->hl, rgba, 16);
    return B200GS_OK;
}
extern "C" int b200gs_set_query(b200gs_viewer* v, const b200gs_query_pod* pod) {
    REQUIRE(v && pod, "null argument");
    REQUIRE(pod->kind <= B200GS_QUERY_TEXTURE, "query kind out of range");
    v->query = *pod;
    return B200GS_OK;
}
static int stage_upload(b200gs_viewer* v, void* dst_dev, const void* src_host, size_t bytes);

// the query texture always matches the viewport (update_query_texture_size, scene.rs:740): (re)allocated cleared
static int ensure_query_texture(b200gs_viewer* v) {
    if (v->query_tex && v->query_tex_w == v->W && v->query_tex_h == v->H) return B200GS_OK;
    TRY(set_device(v));
    if (v->query_tex) {
        CK(cudaStreamSynchronize(v->stream));
        CK(cudaFree(v->query_tex));
        v->query_tex = nullptr;
    }
    TRY(dev_alloc(&v->query_tex, (size_t)v->W * v->H, true, v->stream));
    v->query_tex_w = v->W;
    v->query_tex_h = v->H;
    return B200GS_OK;
}
extern "C" int b200gs_query_texture_clear(b200gs_viewer* v) {
    REQUIRE(v, "null viewer");
    TRY(ensure_query_texture(v));
    CK(cudaMemsetAsync(v->query_tex, 0, (size_t)v->W * v->H, v->stream));
    return B200GS_OK;
}
extern "C" int b200gs_query_texture_paint(b200gs_viewer* v, const b200gs_query_pod* stroke) {
    REQUIRE(v && stroke, "null argument");
    REQUIRE(stroke->kind == B200GS_QUERY_RECT || stroke->kind == B200GS_QUERY_BRUSH, "a stroke is a rect or a brush segment");
    TRY(ensure_query_texture(v));
    CK(gs_launch_paint_query_texture(v->query_tex, v->W, v->H, *stroke, v->stream));
    v->launches += 1;
    return B200GS_OK;
}
extern "C" int b200gs_query_texture_upload(b200gs_viewer* v, const uint8_t* texels, uint32_t width, uint32_t height) {
    REQUIRE(v && texels, "null argument");
    REQUIRE(width == v->W && height == v->H, "the query texture has the size of the viewport");
    TRY(ensure_query_texture(v));
    return stage_upload(v, v->query_tex, texels, (size_t)width * height);
}
extern "C" int b200gs_query_texture_download(b200gs_viewer* v, uint8_t* texels, size_t cap) {
    REQUIRE(v && texels, "null argument");
    REQUIRE(cap >= (size_t)v->W * v->H, "buffer too small");
    TRY(ensure_query_texture(v));
    CK(cudaMemcpyAsync(texels, v->query_tex, (size_t)v->W * v->H, cudaMemcpyDeviceToHost, v->stream));
    CK(cudaStreamSynchronize(v->stream));
    return B200GS_OK;
}

extern "C" int b200gs_set_background(b200gs_viewer* v, const float rgba[4]) {
    REQUIRE(v && rgba, "null argument");
    memcpy(v->bg, rgba, 16);
    return B200GS_OK;
}
extern "C" int b200gs_set_tile_entry_capacity(b200gs_viewer* v, uint64_t entries) {
    REQUIRE(v, "null viewer");
    REQUIRE(entries < 0x40000000ull, "capacity too large");
    v->entry_cap_user = entries;
    v->layout_dirty = true;
    return B200GS_OK;
}
extern "C" int b200gs_enable_timings(b200gs_viewer* v, int on, int count_evals) {
    REQUIRE(v, "null viewer");
    v->timing = on != 0;
    v->count_evals = count_evals != 0;
    return B200GS_OK;
}
extern "C" int b200gs_sync(b200gs_viewer* v) {
    REQUIRE(v, "null viewer");
    TRY(set_device(v));
    if (v->rendered) CK(cudaMemcpyAsync(v->h_small + 34, v->stats + 3, 4, cudaMemcpyDeviceToHost, v->stream));
    CK(cudaStreamSynchronize(v->stream));
    if (v->rendered && v->h_small[34]) {
        gs_set_error("tile-entry capacity exceeded in the last frame (b200gs_set_tile_entry_capacity): the image is truncated");
        return B200GS_ERR_OVERFLOW;
    }
    return B200GS_OK;
}
extern "C" void* b200gs_stream(b200gs_viewer* v) { return v ? (void*)v->stream : nullptr; }
extern "C" void* b200gs_image_device(b200gs_viewer* v) {
    if (!v || set_device(v) != B200GS_OK || ensure_frame_buffers(v) != B200GS_OK) return nullptr;
    return v->image;
}
extern "C" int b200gs_host_alloc(size_t bytes, void** out) {
    REQUIRE(out, "null out");
    CK(cudaMallocHost(out, bytes ? bytes : 1));
    return B200GS_OK;
}
extern "C" int b200gs_host_free(void* p) {
    if (p) CK(cudaFreeHost(p));
    return B200GS_OK;
}

// ---------------------------------------------------------------------------- upload staging
// A slot of the pinned ring, free to be written by the host: waits only for the copy that last used THIS slot.
static int stage_acquire(b200gs_viewer* v, uint8_t** buf, int* slot) {
    for (int i = 0; i < 2; i++) {
        if (!v->stage_buf[i]) {
            CK(cudaMallocHost((void**)&v->stage_buf[i], kStageBytes));
            CK(cudaEventCreateWithFlags(&v->stage_done[i], cudaEventDisableTiming));
        }
    }
    const int k = (int)(v->stage_next++ & 1u);
    if (v->stage_used[k]) CK(cudaEventSynchronize(v->stage_done[k]));
    *buf = v->stage_buf[k];
    *slot = k;
    return B200GS_OK;
}
static int stage_submit(b200gs_viewer* v, int slot, void* dst_dev, size_t bytes) {
    CK(cudaMemcpyAsync(dst_dev, v->stage_buf[slot], bytes, cudaMemcpyHostToDevice, v->stream));
    CK(cudaEventRecord(v->stage_done[slot], v->stream));
    v->stage_used[slot] = true;
    return B200GS_OK;
}
// host -> device through the ring: the host buffer is only borrowed for the call, the call does not wait for the GPU
static int stage_upload(b200gs_viewer* v, void* dst_dev, const void* src_host, size_t bytes) {
    for (size_t o = 0; o < bytes; o += kStageBytes) {
        const size_t c = std::min(kStageBytes, bytes - o);
        uint8_t* buf;
        int slot;
        TRY(stage_acquire(v, &buf, &slot));
        memcpy(buf, (const uint8_t*)src_host + o, c);
        TRY(stage_submit(v, slot, (uint8_t*)dst_dev + o, c));
    }
    return B200GS_OK;
}

// ---------------------------------------------------------------------------- models
static int model_create(b200gs_viewer* v, const char* key, uint64_t capacity, GsRecStore* shared, b200gs_model** out) {
    REQUIRE(v && key && out, "null argument");
    *out = nullptr;
    REQUIRE(capacity < 0x3fffff00ull, "capacity too large");
    REQUIRE(v->models.size() < kMaxModelsPerFrame, "too many models");
    for (auto* m : v->models) REQUIRE(m->key != key, "duplicate model key");
    TRY(set_device(v));
    auto* m = new b200gs_model();
    m->v = v; m->key = key; m->cap = capacity;
    cudaStream_t st = v->stream;
    int rc = B200GS_OK;
    if (shared) {
        shared->refs.fetch_add(1);
        m->store = shared;
    } else {
        m->store = new GsRecStore();
        m->store->device = v->device;
        // records: padded so that the last chunk's 16-byte-rounded TMA copy stays inside the buffer
        rc = dev_alloc(&m->store->p, (size_t)capacity * v->rb + 64, true, st);
    }
    m->recs = m->store->p;
    if (rc == B200GS_OK) rc = dev_alloc(&m->ctrl, MC_WORDS, true, st);
    if (rc == B200GS_OK) rc = dev_alloc(&m->keys_a, capacity, false, st);
    if (rc == B200GS_OK) rc = dev_alloc(&m->vals_a, capacity, false, st);
    if (rc == B200GS_OK) rc = dev_alloc(&m->keys_b, capacity, false, st);
    if (rc == B200GS_OK) rc = dev_alloc(&m->vals_b, capacity, false, st);
    if (rc == B200GS_OK) rc = dev_alloc(&m->idx, capacity, false, st);
    if (rc == B200GS_OK) rc = dev_alloc(&m->binword, capacity, false, st);
    if (rc == B200GS_OK) rc = dev_alloc(&m->lb_pre, gs_preprocess_lookback_words(capacity), true, st);
    if (rc == B200GS_OK) rc = dev_alloc(&m->lb_sort, gs_sort_lookback_words((uint32_t)capacity, 4), true, st);
    if (rc != B200GS_OK) { free_model(m); return rc; }
    v->models.push_back(m);
    v->layout_dirty = true;
    *out = m;
    return B200GS_OK;
}

extern "C" int b200gs_model_create(b200gs_viewer* v, const char* key, uint64_t capacity, b200gs_model** out) {
    return model_create(v, key, capacity, nullptr, out);
}

// A model of `v` that renders the packed records ALREADY resident in `source` (a model of another viewer on the same
// device, same layout): the records are reference counted, nothing is copied.  Mask / selection / edit buffers,
// transform and all per-frame state stay per model.
extern "C" int b200gs_model_create_shared(b200gs_viewer* v, const char* key, b200gs_model* source, b200gs_model** out) {
    REQUIRE(v && source && out, "null argument");
    REQUIRE(source->v->device == v->device, "shared records must live on the viewer's device");
    REQUIRE(source->v->rb == v->rb && source->v->sh == v->sh && source->v->cov == v->cov, "viewers use different record layouts");
    TRY(set_device(v));
    // uploads already enqueued on the source viewer's stream are ordered before anything this viewer does
    cudaEvent_t ev;
    CK(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
    CK(cudaEventRecord(ev, source->v->stream));
    CK(cudaStreamWaitEvent(v->stream, ev, 0));
    CK(cudaEventDestroy(ev));
    return model_create(v, key, source->cap, source->store, out);
}

extern "C" int b200gs_model_destroy(b200gs_viewer* v, b200gs_model* m) {
    REQUIRE(v && m, "null argument");
    auto it = std::find(v->models.begin(), v->models.end(), m);
    REQUIRE(it != v->models.end(), "model does not belong to this viewer");
    TRY(set_device(v));
    CK(cudaStreamSynchronize(v->stream));
    v->models.erase(it);
    free_model(m);
    v->layout_dirty = true;
    return B200GS_OK;
}

extern "C" b200gs_model* b200gs_model_find(b200gs_viewer* v, const char* key) {
    if (!v || !key) return nullptr;
    for (auto* m : v->models)
        if (m->key == key) return m;
    return nullptr;
}

extern "C" uint64_t b200gs_model_len(const b200gs_model* m) { return m ? m->cap : 0; }

extern "C" int b200gs_model_upload_packed(b200gs_model* m, uint64_t start, const void* packed, uint64_t count) {
    REQUIRE(m && (packed || count == 0), "null argument");
    REQUIRE(start <= m->cap && count <= m->cap - start, "range exceeds the model's capacity");
    if (count == 0) return B200GS_OK;
    TRY(set_device(m->v));
    // the host buffer is only borrowed for the call: it is copied into the pinned ring, the H2D copies are enqueued
    return stage_upload(m->v, m->recs + start * m->v->rb, packed, count * m->v->rb);
}

extern "C" int b200gs_model_upload_packed_device(b200gs_model* m, uint64_t start, const void* packed_dev, uint64_t count) {
    REQUIRE(m && (packed_dev || count == 0), "null argument");
    REQUIRE(start <= m->cap && count <= m->cap - start, "range exceeds the model's capacity");
    if (count == 0) return B200GS_OK;
    TRY(set_device(m->v));
    CK(cudaMemcpyAsync(m->recs + start * m->v->rb, packed_dev, count * m->v->rb, cudaMemcpyDeviceToDevice, m->v->stream));
    return B200GS_OK;
}

extern "C" int b200gs_model_update_range(b200gs_model* m, uint64_t start, const b200gs_gaussian* gaussians, uint64_t count) {
    REQUIRE(m && (gaussians || count == 0), "null argument");
    REQUIRE(start <= m->cap && count <= m->cap - start, "range exceeds the model's capacity");
    if (count == 0) return B200GS_OK;
    b200gs_viewer* v = m->v;
    TRY(set_device(v));
    // packed on the host straight into the viewer's pinned ring and streamed up chunk by chunk (scene.rs:2069-2085: the
    // app calls this every frame while a model loads): packing chunk k+1 overlaps the copy of chunk k, no allocation and
    // no stream synchronisation per call — the call returns once the last chunk is enqueued
    const uint32_t rb = v->rb;
    const uint64_t chunk = kStageBytes / rb;
    for (uint64_t o = 0; o < count; o += chunk) {
        const uint64_t c = std::min(chunk, count - o);
        uint8_t* buf;
        int slot;
        TRY(stage_acquire(v, &buf, &slot));
        TRY(b200gs_pack_gaussians(v->sh, v->cov, gaussians + o, c, buf));
        TRY(stage_submit(v, slot, m->recs + (start + o) * rb, c * rb));
    }
    return B200GS_OK;
}

extern "C" int b200gs_model_set_transform(b200gs_model* m, const float pos[3], const float quat_xyzw[4], const float scale[3]) {
    REQUIRE(m && pos && quat_xyzw && scale, "null argument");
    memcpy(m->pos, pos, 12);
    memcpy(m->quat, quat_xyzw, 16);
    memcpy(m->scale, scale, 12);
    return B200GS_OK;
}

static int upload_words(b200gs_model* m, uint32_t** dst, const uint32_t* words, uint64_t nwords) {
    uint64_t need = (m->cap + 31) / 32;
    REQUIRE(words && nwords == need, "bitset must hold ceil(N/32) words");
    TRY(set_device(m->v));
    if (!*dst) TRY(dev_alloc(dst, need, true, m->v->stream));
    return stage_upload(m->v, *dst, words, need * 4);
}
extern "C" int b200gs_model_upload_mask(b200gs_model* m, const uint32_t* words, uint64_t nwords) {
    REQUIRE(m, "null model");
    return upload_words(m, &m->mask, words, nwords);
}
extern "C" int b200gs_model_upload_selection(b200gs_model* m, const uint32_t* words, uint64_t nwords) {
    REQUIRE(m, "null model");
    return upload_words(m, &m->selection, words, nwords);
}

static int ensure_edits(b200gs_model* m) {
    if (m->edits) return B200GS_OK;
    TRY(dev_alloc(&m->edits, m->cap, false, m->v->stream));
    CK(gs_launch_fill_default_edits((uint32_t)m->cap, m->edits, m->v->stream));
    return B200GS_OK;
}
extern "C" int b200gs_model_upload_edits(b200gs_model* m, uint64_t start, const b200gs_edit_pod* pods, uint64_t count) {
    REQUIRE(m && (pods || count == 0), "null argument");
    REQUIRE(start <= m->cap && count <= m->cap - start, "range exceeds the model's capacity");
    TRY(set_device(m->v));
    TRY(ensure_edits(m));
    if (count) TRY(stage_upload(m->v, m->edits + start, pods, count * sizeof(b200gs_edit_pod)));
    return B200GS_OK;
}

extern "C" int b200gs_model_eval_mask(b200gs_model* m, const b200gs_mask_op* postfix, uint32_t n_ops,
                                      const b200gs_mask_shape* shapes, uint32_t n_shapes) {
    REQUIRE(m, "null model");
    REQUIRE(n_ops <= 64, "too many mask ops");
    REQUIRE((postfix || n_ops == 0) && (shapes || n_shapes == 0), "null argument");
    // validate the postfix program on the host (stack depth, shape indices)
    int sp = n_ops == 0 ? 1 : 0;
    for (uint32_t o = 0; o < n_ops; o++) {
        switch (postfix[o].kind) {
            case B200GS_MASKOP_SHAPE: REQUIRE(postfix[o].arg < n_shapes, "shape index out of range"); sp++; break;
            case B200GS_MASKOP_RESET: sp++; break;
            case B200GS_MASKOP_COMPLEMENT: REQUIRE(sp >= 1, "malformed op tree"); break;
            case B200GS_MASKOP_UNION: case B200GS_MASKOP_INTERSECTION: case B200GS_MASKOP_DIFFERENCE:
            case B200GS_MASKOP_SYMDIFF: REQUIRE(sp >= 2, "malformed op tree"); sp--; break;
            default: REQUIRE(false, "unknown mask op");
        }
        REQUIRE(sp <= 60, "op tree too deep");
    }
    REQUIRE(sp == 1, "malformed op tree");
    REQUIRE(n_shapes <= 64, "too many mask shapes");
    b200gs_viewer* v = m->v;
    TRY(set_device(v));
    cudaStream_t st = v->stream;
    uint64_t need = (m->cap + 31) / 32;
    if (!m->mask) TRY(dev_alloc(&m->mask, need, true, st));
    if (!v->mask_scratch) TRY(dev_alloc(&v->mask_scratch, kMaskScratchBytes, false, st));
    // postfix ops | shapes | 3x3 shape rotations, staged in one pinned block (no allocation, no synchronisation:
    // the evaluation is enqueued behind the upload on the viewer's stream)
    uint8_t* buf;
    int slot;
    TRY(stage_acquire(v, &buf, &slot));
    b200gs_mask_op* h_ops = reinterpret_cast<b200gs_mask_op*>(buf);
    b200gs_mask_shape* h_shapes = reinterpret_cast<b200gs_mask_shape*>(buf + 64 * sizeof(b200gs_mask_op));
    float* h_rot = reinterpret_cast<float*>(buf + 64 * sizeof(b200gs_mask_op) + 64 * sizeof(b200gs_mask_shape));
    if (n_ops) memcpy(h_ops, postfix, n_ops * sizeof(b200gs_mask_op));
    if (n_shapes) memcpy(h_shapes, shapes, n_shapes * sizeof(b200gs_mask_shape));
    for (uint32_t s2 = 0; s2 < n_shapes; s2++) {
        float R[3][3];
        quat_to_mat3(shapes[s2].quat, R);
        memcpy(h_rot + 9 * s2, R, 36);
    }
    TRY(stage_submit(v, slot, v->mask_scratch, kMaskScratchBytes));
    const b200gs_mask_op* d_ops = reinterpret_cast<const b200gs_mask_op*>(v->mask_scratch);
    const b200gs_mask_shape* d_shapes = reinterpret_cast<const b200gs_mask_shape*>(v->mask_scratch + 64 * sizeof(b200gs_mask_op));
    const float* d_rot = reinterpret_cast<const float*>(v->mask_scratch + 64 * sizeof(b200gs_mask_op) + 64 * sizeof(b200gs_mask_shape));
    GsModelXf xf = make_xf(m);
    CK(gs_launch_eval_mask(m->recs, (uint32_t)m->cap, v->rb, xf, d_ops, n_ops, d_shapes, d_rot, m->mask, st));
    return B200GS_OK;
}

extern "C" int b200gs_model_postprocess(b200gs_model* m) {
    REQUIRE(m, "null model");
    if (!m->selection || !(m->v->sel_edit.flag & B200GS_EDIT_ENABLED)) return B200GS_OK;
    TRY(set_device(m->v));
    TRY(ensure_edits(m));
    CK(gs_launch_postprocess((uint32_t)m->cap, m->selection, m->edits, m->v->sel_edit, m->v->stream));
    return B200GS_OK;
}

// ---------------------------------------------------------------------------- hot path
extern "C" int b200gs_model_preprocess(b200gs_model* m, int use_unedited) {
    REQUIRE(m, "null model");
    b200gs_viewer* v = m->v;
    TRY(set_device(v));
    TRY(ensure_frame_buffers(v));
    CK(cudaMemsetAsync(m->ctrl, 0, MC_WORDS * 4, v->stream));
    if (v->query.kind == B200GS_QUERY_TEXTURE) TRY(ensure_query_texture(v));
    GsFrame f = make_frame(v);
    if (f.query.kind >= B200GS_QUERY_RECT && !m->selection)  // a selection query writes the selection bitset
        TRY(dev_alloc(&m->selection, (m->cap + 31) / 32, true, v->stream));
    if (use_unedited) f.sel_edit = b200gs_edit_pod{0, {0.0f, 1.0f, 1.0f}, 0.0f, 0.0f, 1.0f, 1.0f};
    GsModelXf xf = make_xf(m);
    GsPreprocessArgs a;
    a.recs = m->recs; a.n = (uint32_t)m->cap; a.sh = v->sh; a.cov = v->cov;
    a.mask = m->mask; a.selection = m->selection; a.edits = use_unedited ? nullptr : m->edits;
    a.ctrl = m->ctrl + MC_CTRL; a.lookback = m->lb_pre; a.epoch = ++v->epoch;
    a.keys = m->keys_a; a.idx = m->idx; a.splats = v->arena + m->arena_offset;
    a.sort_hist = m->ctrl + MC_SORT_HIST;
    a.binword = m->binword;
    if (v->timing) CK(cudaEventRecord(v->ev[0], v->stream));
    CK(gs_launch_preprocess(a, f, xf, v->num_sms, v->stream));
    v->launches += 1;
    if (v->timing) CK(cudaEventRecord(v->ev[1], v->stream));
    m->preprocessed = true;
    m->sorted = false;
    return B200GS_OK;
}

extern "C" int b200gs_model_sort(b200gs_model* m) {
    REQUIRE(m, "null model");
    REQUIRE(m->preprocessed, "sort called before preprocess");
    if (m->sorted) return B200GS_OK;  // already sorted since the last preprocess
    b200gs_viewer* v = m->v;
    TRY(set_device(v));
    GsSortArgs a;
    a.keys_a = m->keys_a; a.vals_a = m->vals_a; a.keys_b = m->keys_b; a.vals_b = m->vals_b;
    a.d_n = m->ctrl + MC_CTRL + GS_CTRL_VISIBLE; a.n_max = (uint32_t)m->cap;
    a.hist = m->ctrl + MC_SORT_HIST; a.lookback = m->lb_sort; a.epoch = ++v->epoch;
    a.tickets = m->ctrl + MC_SORT_TICKET; a.passes = 4; a.hist_prefilled = true; a.vals_identity = true;
    a.vote_mask = 0x3;  // depth keys: the two low bytes are spread, the two high bytes concentrated
    a.result_in_b = m->ctrl + MC_CTRL + GS_CTRL_SORT_IN_B;
    CK(gs_launch_sort(a, v->num_sms, v->stream));
    v->launches += a.passes;
    if (v->timing) CK(cudaEventRecord(v->ev[2], v->stream));
    m->sorted = true;
    return B200GS_OK;
}

// Binning (nearest model first: its splats come first in every bin's front-to-back list), per-bin ranges + launch order,
// the bin sort, and the compositor.  (An earlier revision could split the frame into depth slabs with finished-tile
// skipping; on the benchmark scene the extra launches cost more than the skipped entries saved, and it was removed when
// binning moved to 32-pixel bins.)
static int render_binned(b200gs_viewer* v, b200gs_model* const* far_to_near, uint32_t n_models, void* rgba8_out, size_t pitch) {
    cudaStream_t st = v->stream;
    GsFrame f = make_frame(v);
    const uint32_t n_bins = f.bins_x * f.bins_y;
    CK(cudaMemsetAsync(v->stats, 0, 64, st));
    if (v->timing) CK(cudaEventRecord(v->ev_bin[0], st));
    CK(cudaMemsetAsync(v->vctrl, 0, VC_WORDS * 4, st));
    for (uint32_t k = 0; k < n_models; k++) {
        b200gs_model* m = far_to_near[n_models - 1 - k];
        GsBinArgs b;
        b.sorted_slot = m->vals_a;
        b.sorted_slot_b = m->vals_b;
        b.sorted_in_b = m->ctrl + MC_CTRL + GS_CTRL_SORT_IN_B;
        b.splats = v->arena + m->arena_offset;
        b.binword = m->binword;
        b.d_v = m->ctrl + MC_CTRL + GS_CTRL_VISIBLE;
        b.v_max = (uint32_t)m->cap;
        b.splat_base = (uint32_t)m->arena_offset;
        b.lookback = v->lb_bin + (size_t)k * v->lb_bin_words;
        b.epoch = ++v->epoch;
        b.ticket = v->vctrl + VC_BIN_TICKET + k;
        b.entry_base_in = v->vctrl + VC_ENTRY_TOTAL + k;
        b.entry_total_out = v->vctrl + VC_ENTRY_TOTAL + k + 1;
        b.overflow = (uint32_t*)(v->stats + 3);
        b.tile_keys = v->tk_a; b.tile_vals = v->tv_a; b.capacity = (uint32_t)v->entry_cap;
        b.tile_count = v->tile_count;
        CK(gs_launch_bin(b, f, v->num_sms, st));
        v->launches += 1;
    }
    // bin ids are sorted on 16 bits (2 onesweep passes); viewports with more than 65536 bins take a third.  The
    // quadrant masks in the top key bits ride along untouched.
    const uint32_t tpasses = n_bins > 65536u ? 3u : 2u;
    GsTileRangesArgs tr;
    tr.tile_count = v->tile_count; tr.ranges = v->ranges; tr.n_tiles = n_bins;
    tr.hist = v->vctrl + VC_TSORT_HIST; tr.passes = tpasses; tr.entry_stat = v->stats + 2;
    tr.lookback = v->lb_tiles; tr.epoch = ++v->epoch;
    tr.ticket = v->vctrl + VC_TILE_TICKET; tr.done_ctr = v->vctrl + VC_TILE_DONE; tr.buckets = v->vctrl + VC_TILE_BUCKETS;
    CK(gs_launch_tile_ranges(tr, st));
    GsSortArgs s;
    s.keys_a = v->tk_a; s.vals_a = v->tv_a; s.keys_b = v->tk_b; s.vals_b = v->tv_b;
    s.d_n = v->vctrl + VC_ENTRY_TOTAL + n_models; s.n_max = (uint32_t)v->entry_cap;
    s.hist = v->vctrl + VC_TSORT_HIST; s.lookback = v->lb_tsort; s.epoch = ++v->epoch;
    s.tickets = v->vctrl + VC_TSORT_TICKET; s.passes = tpasses; s.hist_prefilled = true; s.vals_identity = false;
    s.vote_mask = 0x1;  // bin ids: the low byte is spread, the row-band bytes concentrated
    s.result_in_b = v->vctrl + VC_TSORT_IN_B;
    CK(gs_launch_sort(s, v->num_sms, st));
    if (v->timing) CK(cudaEventRecord(v->ev_bin[1], st));
    GsCompositeArgs c;
    c.tile_keys = v->tk_a; c.tile_vals = v->tv_a; c.tile_keys_b = v->tk_b; c.tile_vals_b = v->tv_b;
    c.tile_in_b = s.result_in_b; c.ranges = v->ranges; c.splats = v->arena;
    c.out = (uint8_t*)rgba8_out; c.pitch = pitch;
    c.evals = v->count_evals ? v->stats : nullptr;
    CK(gs_launch_composite(c, f, st));
    if (v->timing) CK(cudaEventRecord(v->ev_bin[2], st));
    v->launches += tpasses + 2;  // bin sort passes, tile finish, compositor
    if (v->timing) CK(cudaEventRecord(v->ev[4], st));
    v->rendered = true;
    return B200GS_OK;
}

extern "C" int b200gs_render(b200gs_viewer* v, b200gs_model* const* far_to_near, uint32_t n_models, void* rgba8_out, size_t pitch) {
    REQUIRE(v && rgba8_out && (far_to_near || n_models == 0), "null argument");
    REQUIRE(n_models <= kMaxModelsPerFrame, "too many models");
    REQUIRE(pitch >= (size_t)v->W * 4, "pitch too small");
    REQUIRE(pitch % 4 == 0 && ((uintptr_t)rgba8_out & 3u) == 0, "rgba8_out and pitch must be 4-byte aligned (pixels are stored as u32)");
    for (uint32_t i = 0; i < n_models; i++) {
        REQUIRE(far_to_near[i] && far_to_near[i]->v == v, "model does not belong to this viewer");
        REQUIRE(far_to_near[i]->sorted, "render called before preprocess + sort");
    }
    TRY(set_device(v));
    TRY(ensure_frame_buffers(v));
    for (uint32_t i = 0; i < n_models; i++) REQUIRE(far_to_near[i]->sorted, "model layout changed; preprocess + sort again");
    return render_binned(v, far_to_near, n_models, rgba8_out, pitch);
}

extern "C" int b200gs_render_frame(b200gs_viewer* v, b200gs_model* const* far_to_near, uint32_t n_models, void* rgba8_out, size_t pitch) {
    REQUIRE(v && (far_to_near || n_models == 0), "null argument");
    for (uint32_t i = 0; i < n_models; i++) {
        REQUIRE(far_to_near[i], "null model");
        TRY(b200gs_model_preprocess(far_to_near[i], 0));
        TRY(b200gs_model_sort(far_to_near[i]));
    }
    return b200gs_render(v, far_to_near, n_models, rgba8_out, pitch);
}

// Pipelined host frames: _begin enqueues the frame and its D2H copy (on a second stream) into one
// of two slots and returns; _end waits for the OLDEST outstanding frame.  With begin(i+1) issued
// before end(i), the copy of frame i overlaps the rendering of frame i+1.
extern "C" int b200gs_render_frame_host_begin(b200gs_viewer* v, b200gs_model* const* far_to_near, uint32_t n_models,
                                              const float view[16], const float proj[16], void* rgba8_host) {
    REQUIRE(v && rgba8_host, "null argument");
    REQUIRE(v->ring_pending < 2, "two frames are already in flight: call b200gs_render_frame_host_end first");
    if (view && proj) TRY(b200gs_set_camera(v, view, proj, nullptr));
    TRY(set_device(v));
    TRY(ensure_frame_buffers(v));
    const uint32_t slot = (v->ring_head + v->ring_pending) & 1u;
    const size_t img = (size_t)v->W * v->H * 4;
    uint8_t* dev = v->image + (size_t)slot * v->image_bytes;
    // this slot's previous image (two frames ago) must have left the device before it is overwritten
    CK(cudaStreamWaitEvent(v->stream, v->ev_copied[slot], 0));
    TRY(b200gs_render_frame(v, far_to_near, n_models, dev, (size_t)v->W * 4));
    CK(cudaEventRecord(v->ev_rendered[slot], v->stream));
    CK(cudaStreamWaitEvent(v->copy_stream, v->ev_rendered[slot], 0));
    CK(cudaMemcpyAsync(rgba8_host, dev, img, cudaMemcpyDeviceToHost, v->copy_stream));
    // the frame's overflow flag travels with the image (checked by _end: no extra synchronisation)
    CK(cudaMemcpyAsync(v->h_small + 32 + slot, v->stats + 3, 4, cudaMemcpyDeviceToHost, v->copy_stream));
    CK(cudaEventRecord(v->ev_copied[slot], v->copy_stream));
    v->ring_pending++;
    return B200GS_OK;
}

extern "C" int b200gs_render_frame_host_end(b200gs_viewer* v) {
    REQUIRE(v, "null viewer");
    REQUIRE(v->ring_pending > 0, "no frame in flight");
    TRY(set_device(v));
    const uint32_t slot = v->ring_head & 1u;
    CK(cudaEventSynchronize(v->ev_copied[slot]));
    v->ring_head++;
    v->ring_pending--;
    if (v->h_small[32 + slot]) {   // the image was delivered, but entries beyond the capacity were dropped
        gs_set_error("tile-entry capacity exceeded in this frame (b200gs_set_tile_entry_capacity): the image is truncated");
        return B200GS_ERR_OVERFLOW;
    }
    return B200GS_OK;
}

extern "C" int b200gs_render_frame_host(b200gs_viewer* v, b200gs_model* const* far_to_near, uint32_t n_models,
                                        const float view[16], const float proj[16], void* rgba8_host) {
    REQUIRE(v && rgba8_host, "null argument");
    while (v->ring_pending) {
        const int rc = b200gs_render_frame_host_end(v);
        if (rc != B200GS_OK && rc != B200GS_ERR_OVERFLOW) return rc;
    }
    TRY(b200gs_render_frame_host_begin(v, far_to_near, n_models, view, proj, rgba8_host));
    return b200gs_render_frame_host_end(v);
}

extern "C" int b200gs_launch_count(b200gs_viewer* v, uint64_t* out) {
    REQUIRE(v && out, "null argument");
    *out = v->launches;
    return B200GS_OK;
}

extern "C" int b200gs_order_models(b200gs_viewer* v, b200gs_model* const* models, const float* centers, uint32_t n, uint32_t* order_out) {
    REQUIRE(v && (models || n == 0) && (centers || n == 0) && (order_out || n == 0), "null argument");
    GsFrame f = make_frame(v);
    std::vector<float> d(n);
    for (uint32_t i = 0; i < n; i++) {
        REQUIRE(models[i], "null model");
        GsModelXf x = make_xf(models[i]);
        float cs[3] = {centers[3 * i] * x.s[0], centers[3 * i + 1] * x.s[1], centers[3 * i + 2] * x.s[2]};
        float acc = 0.0f;
        for (int r = 0; r < 3; r++) {
            float w = x.R[r][0] * cs[0] + x.R[r][1] * cs[1] + x.R[r][2] * cs[2] + x.t[r];
            float dd = w - f.cam[r];
            acc = acc + dd * dd;
        }
        d[i] = acc;
        order_out[i] = i;
    }
    std::stable_sort(order_out, order_out + n, [&](uint32_t a, uint32_t b) { return d[a] > d[b]; });
    return B200GS_OK;
}

// ---------------------------------------------------------------------------- downloads
static int visible_count(b200gs_model* m, uint64_t* out) {
    b200gs_viewer* v = m->v;
    TRY(set_device(v));
    if (!m->preprocessed) { *out = 0; return B200GS_OK; }
    CK(cudaMemcpyAsync(v->h_small, m->ctrl + MC_CTRL + GS_CTRL_VISIBLE, 4, cudaMemcpyDeviceToHost, v->stream));
    CK(cudaStreamSynchronize(v->stream));
    *out = v->h_small[0];
    return B200GS_OK;
}
extern "C" int b200gs_model_visible_count(b200gs_model* m, uint64_t* out) {
    REQUIRE(m && out, "null argument");
    return visible_count(m, out);
}
// after a sort the data may sit in the *_b buffers (a degenerate digit pass was skipped)
static int sorted_in_b(b200gs_model* m, bool* out) {
    *out = false;
    if (!m->sorted) return B200GS_OK;
    b200gs_viewer* v = m->v;
    CK(cudaMemcpyAsync(v->h_small, m->ctrl + MC_CTRL + GS_CTRL_SORT_IN_B, 4, cudaMemcpyDeviceToHost, v->stream));
    CK(cudaStreamSynchronize(v->stream));
    *out = v->h_small[0] != 0;
    return B200GS_OK;
}
static int download_u32(b200gs_model* m, const uint32_t* src, uint32_t* dst, uint64_t n) {
    if (n) CK(cudaMemcpyAsync(dst, src, n * 4, cudaMemcpyDeviceToHost, m->v->stream));
    CK(cudaStreamSynchronize(m->v->stream));
    return B200GS_OK;
}
extern "C" int b200gs_model_download_depth_keys(b200gs_model* m, uint32_t* keys, uint64_t cap, uint64_t* n) {
    REQUIRE(m && n && (keys || cap == 0), "null argument");
    uint64_t vc;
    TRY(visible_count(m, &vc));
    *n = vc;
    REQUIRE(cap >= vc, "buffer too small");
    bool in_b;
    TRY(sorted_in_b(m, &in_b));
    return download_u32(m, in_b ? m->keys_b : m->keys_a, keys, vc);
}
extern "C" int b200gs_model_download_indices(b200gs_model* m, uint32_t* idx, uint64_t cap, uint64_t* n) {
    REQUIRE(m && n && (idx || cap == 0), "null argument");
    uint64_t vc;
    TRY(visible_count(m, &vc));
    *n = vc;
    REQUIRE(cap >= vc, "buffer too small");
    if (!m->sorted) return download_u32(m, m->idx, idx, vc);
    std::vector<uint32_t> slots(vc), orig(vc);
    bool in_b;
    TRY(sorted_in_b(m, &in_b));
    TRY(download_u32(m, in_b ? m->vals_b : m->vals_a, slots.data(), vc));
    TRY(download_u32(m, m->idx, orig.data(), vc));
    for (uint64_t i = 0; i < vc; i++) idx[i] = slots[i] < vc ? orig[slots[i]] : 0xffffffffu;
    return B200GS_OK;
}
extern "C" int b200gs_model_download_splats(b200gs_model* m, b200gs_splat* out, uint64_t cap, uint64_t* n) {
    REQUIRE(m && n && (out || cap == 0), "null argument");
    uint64_t vc;
    TRY(visible_count(m, &vc));
    *n = vc;
    REQUIRE(cap >= vc, "buffer too small");
    b200gs_viewer* v = m->v;
    if (!m->sorted) {
        if (vc) CK(cudaMemcpyAsync(out, v->arena + m->arena_offset, vc * sizeof(b200gs_splat), cudaMemcpyDeviceToHost, v->stream));
        CK(cudaStreamSynchronize(v->stream));
        return B200GS_OK;
    }
    std::vector<uint32_t> slots(vc);
    std::vector<b200gs_splat> tmp(vc);
    bool in_b;
    TRY(sorted_in_b(m, &in_b));
    TRY(download_u32(m, in_b ? m->vals_b : m->vals_a, slots.data(), vc));
    if (vc) CK(cudaMemcpyAsync(tmp.data(), v->arena + m->arena_offset, vc * sizeof(b200gs_splat), cudaMemcpyDeviceToHost, v->stream));
    CK(cudaStreamSynchronize(v->stream));
    for (uint64_t i = 0; i < vc; i++) {
        if (slots[i] < vc) out[i] = tmp[slots[i]];
        else memset(&out[i], 0xff, sizeof(b200gs_splat));
    }
    return B200GS_OK;
}
static int download_words(b200gs_model* m, const uint32_t* src, uint32_t fill, uint32_t* words, uint64_t cap_words, uint64_t* n) {
    uint64_t need = (m->cap + 31) / 32;
    *n = need;
    REQUIRE(cap_words >= need, "buffer too small");
    TRY(set_device(m->v));
    if (!src) {
        for (uint64_t i = 0; i < need; i++) words[i] = fill;
        if (fill && (m->cap & 31) && need) words[need - 1] = (1u << (m->cap & 31)) - 1u;
        return B200GS_OK;
    }
    return download_u32(m, src, words, need);
}
extern "C" int b200gs_model_download_mask(b200gs_model* m, uint32_t* words, uint64_t cap_words, uint64_t* n) {
    REQUIRE(m && n && (words || cap_words == 0), "null argument");
    return download_words(m, m->mask, 0xffffffffu, words, cap_words, n);  // MaskOpTree::Reset = all shown
}
extern "C" int b200gs_model_download_selection(b200gs_model* m, uint32_t* words, uint64_t cap_words, uint64_t* n) {
    REQUIRE(m && n && (words || cap_words == 0), "null argument");
    return download_words(m, m->selection, 0u, words, cap_words, n);
}
extern "C" int b200gs_model_download_edits(b200gs_model* m, b200gs_edit_pod* pods, uint64_t cap, uint64_t* n) {
    REQUIRE(m && n && (pods || cap == 0), "null argument");
    *n = m->cap;
    REQUIRE(cap >= m->cap, "buffer too small");
    TRY(set_device(m->v));
    TRY(ensure_edits(m));
    if (m->cap) CK(cudaMemcpyAsync(pods, m->edits, m->cap * sizeof(b200gs_edit_pod), cudaMemcpyDeviceToHost, m->v->stream));
    CK(cudaStreamSynchronize(m->v->stream));
    return B200GS_OK;
}
extern "C" int b200gs_model_download_packed(b200gs_model* m, uint64_t start, void* packed, uint64_t count) {
    REQUIRE(m && (packed || count == 0), "null argument");
    REQUIRE(start <= m->cap && count <= m->cap - start, "range exceeds the model's capacity");
    TRY(set_device(m->v));
    if (count) CK(cudaMemcpyAsync(packed, m->recs + start * m->v->rb, count * m->v->rb, cudaMemcpyDeviceToHost, m->v->stream));
    CK(cudaStreamSynchronize(m->v->stream));
    return B200GS_OK;
}

extern "C" int b200gs_last_timings(b200gs_viewer* v, b200gs_timings* out) {
    REQUIRE(v && out, "null argument");
    TRY(set_device(v));
    CK(cudaStreamSynchronize(v->stream));
    memset(out, 0, sizeof *out);
    if (v->timing) {
        float t;
        if (cudaEventElapsedTime(&t, v->ev[0], v->ev[1]) == cudaSuccess) out->preprocess_ms = t;
        if (cudaEventElapsedTime(&t, v->ev[1], v->ev[2]) == cudaSuccess) out->sort_ms = t;
        if (cudaEventElapsedTime(&t, v->ev_bin[0], v->ev_bin[1]) == cudaSuccess) out->bin_ms = t;
        if (cudaEventElapsedTime(&t, v->ev_bin[1], v->ev_bin[2]) == cudaSuccess) out->composite_ms = t;
        if (cudaEventElapsedTime(&t, v->ev[0], v->ev[4]) == cudaSuccess) out->total_ms = t;
        (void)cudaGetLastError();
    }
    uint64_t vis = 0;
    for (auto* m : v->models) {
        uint64_t c = 0;
        if (m->preprocessed) TRY(visible_count(m, &c));
        vis += c;
    }
    out->visible = vis;
    CK(cudaMemcpyAsync(v->h_small, v->stats, 64, cudaMemcpyDeviceToHost, v->stream));
    CK(cudaStreamSynchronize(v->stream));
    uint64_t st64[4];
    memcpy(st64, v->h_small, 32);
    out->evals = st64[0];
    out->staged_entries = st64[1];
    out->tile_entries = st64[2];
    out->overflow = (uint32_t)st64[3];
    return B200GS_OK;
}

// ---------------------------------------------------------------------------- hit query (row N3)
extern "C" int b200gs_query_hits(b200gs_viewer* v, b200gs_model* const* far_to_near, uint32_t n_models, uint32_t px,
                                 uint32_t py, b200gs_hit* out, uint64_t cap, uint64_t* n) {
    REQUIRE(v && n && (out || cap == 0) && (far_to_near || n_models == 0), "null argument");
    REQUIRE(px < v->W && py < v->H, "pixel outside the viewport");
    REQUIRE(v->rendered, "query_hits needs a rendered frame");
    for (uint32_t i = 0; i < n_models; i++) REQUIRE(far_to_near[i] && far_to_near[i]->v == v && far_to_near[i]->sorted, "model not rendered");
    TRY(set_device(v));
    cudaStream_t st = v->stream;
    // the hit list needs the per-bin lists of THIS frame: re-bin it (event-driven call) into the
    // query's own scratch target — the two host-frame slots may hold frames that are still being copied out
    TRY(ensure_frame_buffers(v));
    TRY(render_binned(v, far_to_near, n_models, v->image + 2 * v->image_bytes, (size_t)v->W * 4));
    const uint32_t c32 = (uint32_t)std::min<uint64_t>(cap, 1u << 20);
    uint2* d_out = nullptr;
    uint32_t* d_cnt = nullptr;
    TRY(dev_alloc(&d_out, c32, false, st));
    TRY(dev_alloc(&d_cnt, 1, true, st));
    GsFrame f = make_frame(v);
    GsCompositeArgs c;
    c.tile_keys = v->tk_a; c.tile_vals = v->tv_a; c.tile_keys_b = v->tk_b; c.tile_vals_b = v->tv_b;
    c.tile_in_b = v->vctrl + VC_TSORT_IN_B; c.ranges = v->ranges;
    c.splats = v->arena; c.out = nullptr; c.pitch = 0; c.evals = nullptr;
    CK(gs_launch_query_hits(c, f, px, py, d_out, c32, d_cnt, st));
    std::vector<uint2> raw(c32);
    uint32_t cnt = 0;
    CK(cudaMemcpyAsync(&cnt, d_cnt, 4, cudaMemcpyDeviceToHost, st));
    if (c32) CK(cudaMemcpyAsync(raw.data(), d_out, (size_t)c32 * 8, cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    cudaFree(d_out); cudaFree(d_cnt);
    *n = cnt;
    const uint64_t take = std::min<uint64_t>(cnt, c32);
    // splat id -> (model, Gaussian index, depth): the arena slot is the compaction slot of its model
    struct MInfo { uint64_t lo, vc; std::vector<uint32_t> idx, key_sorted, slot_sorted; };
    std::vector<MInfo> info(n_models);
    for (uint32_t k = 0; k < n_models; k++) {
        b200gs_model* m = far_to_near[k];
        info[k].lo = m->arena_offset;
        TRY(visible_count(m, &info[k].vc));
    }
    for (uint64_t h = 0; h < take; h++) {
        const uint32_t id = raw[h].x;
        out[h].model = 0xffffffffu; out[h].index = 0xffffffffu; out[h].depth = 0.0f;
        memcpy(&out[h].alpha, &raw[h].y, 4);
        for (uint32_t k = 0; k < n_models; k++) {
            if (id < info[k].lo || id >= info[k].lo + info[k].vc) continue;
            MInfo& mi = info[k];
            b200gs_model* m = far_to_near[k];
            if (mi.idx.empty() && mi.vc) {  // lazily fetch the model's index / key tables
                bool in_b;
                TRY(sorted_in_b(m, &in_b));
                mi.idx.resize(mi.vc); mi.key_sorted.resize(mi.vc); mi.slot_sorted.resize(mi.vc);
                TRY(download_u32(m, m->idx, mi.idx.data(), mi.vc));
                TRY(download_u32(m, in_b ? m->keys_b : m->keys_a, mi.key_sorted.data(), mi.vc));
                TRY(download_u32(m, in_b ? m->vals_b : m->vals_a, mi.slot_sorted.data(), mi.vc));
                std::vector<uint32_t> key_by_slot(mi.vc);
                for (uint64_t r = 0; r < mi.vc; r++)
                    if (mi.slot_sorted[r] < mi.vc) key_by_slot[mi.slot_sorted[r]] = mi.key_sorted[r];
                mi.key_sorted.swap(key_by_slot);  // now indexed by slot
            }
            const uint32_t slot = id - (uint32_t)mi.lo;
            out[h].model = k;
            out[h].index = mi.idx[slot];
            memcpy(&out[h].depth, &mi.key_sorted[slot], 4);
            break;
        }
    }
    return B200GS_OK;
}

// ---------------------------------------------------------------------------- raw sort
// raw sort entry points: bits = 16 / 32 run the 8-bit-digit sort (K2), `wide` != 0 the 11-bit cluster sort (K2w, any bits)
static int sort_pairs_device(b200gs_viewer* v, uint32_t* keys_dev, uint32_t* values_dev, uint64_t n, uint32_t bits, bool wide) {
    REQUIRE(v && (keys_dev || n == 0) && (values_dev || n == 0), "null argument");
    REQUIRE(wide ? (bits >= 1 && bits <= 32) : (bits == 16 || bits == 32), "bits must be 16 or 32 (1..32 for the wide sort)");
    REQUIRE(n < 0x3fffff00ull, "too many elements");
    if (n == 0) return B200GS_OK;
    TRY(set_device(v));
    cudaStream_t st = v->stream;
    uint32_t *kb = nullptr, *vb = nullptr, *ctl = nullptr;
    uint64_t* lb = nullptr;
    TRY(dev_alloc(&kb, n, false, st));
    TRY(dev_alloc(&vb, n, false, st));
    TRY(dev_alloc(&ctl, 1024 + 3 * 2048, true, st));
    TRY(dev_alloc(&lb, wide ? gs_sort_wide_lookback_words((uint32_t)n, bits) : gs_sort_lookback_words((uint32_t)n, bits / 8), true, st));
    uint32_t nn = (uint32_t)n;
    CK(cudaMemcpyAsync(ctl, &nn, 4, cudaMemcpyHostToDevice, st));
    if (wide) {
        GsSortWideArgs a;
        a.keys_a = keys_dev; a.vals_a = values_dev; a.keys_b = kb; a.vals_b = vb;
        a.d_n = ctl; a.n_max = nn; a.hist = ctl + 1024; a.lookback = lb; a.epoch = ++v->epoch;
        a.tickets = ctl + 8; a.key_bits = bits; a.hist_prefilled = false; a.vals_identity = false;
        a.result_in_b = ctl + 16;
        CK(gs_launch_sort_wide(a, v->num_sms, st));
    } else {
        GsSortArgs a;
        a.keys_a = keys_dev; a.vals_a = values_dev; a.keys_b = kb; a.vals_b = vb;
        a.d_n = ctl; a.n_max = nn; a.hist = ctl + 1024; a.lookback = lb; a.epoch = ++v->epoch;
        a.tickets = ctl + 8; a.passes = bits / 8; a.hist_prefilled = false; a.vals_identity = false;
        a.vote_mask = 0xf;  // arbitrary keys: assume spread digits
        a.result_in_b = ctl + 16;
        CK(gs_launch_sort(a, v->num_sms, st));
    }
    CK(cudaMemcpyAsync(v->h_small, ctl + 16, 4, cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    if (v->h_small[0]) {  // an odd number of passes ran: the result is in the scratch buffers
        CK(cudaMemcpyAsync(keys_dev, kb, n * 4, cudaMemcpyDeviceToDevice, st));
        CK(cudaMemcpyAsync(values_dev, vb, n * 4, cudaMemcpyDeviceToDevice, st));
        CK(cudaStreamSynchronize(st));
    }
    cudaFree(kb); cudaFree(vb); cudaFree(ctl); cudaFree(lb);
    return B200GS_OK;
}
extern "C" int b200gs_sort_pairs_device(b200gs_viewer* v, uint32_t* keys_dev, uint32_t* values_dev, uint64_t n, uint32_t bits) {
    return sort_pairs_device(v, keys_dev, values_dev, n, bits, false);
}
extern "C" int b200gs_sort_pairs_wide_device(b200gs_viewer* v, uint32_t* keys_dev, uint32_t* values_dev, uint64_t n, uint32_t bits) {
    return sort_pairs_device(v, keys_dev, values_dev, n, bits, true);
}

static int sort_pairs_host(b200gs_viewer* v, uint32_t* keys, uint32_t* values, uint64_t n, uint32_t bits, bool wide) {
    REQUIRE(v && (keys || n == 0) && (values || n == 0), "null argument");
    if (n == 0) return B200GS_OK;
    TRY(set_device(v));
    uint32_t *dk = nullptr, *dv = nullptr;
    TRY(dev_alloc(&dk, n, false, v->stream));
    TRY(dev_alloc(&dv, n, false, v->stream));
    CK(cudaMemcpyAsync(dk, keys, n * 4, cudaMemcpyHostToDevice, v->stream));
    CK(cudaMemcpyAsync(dv, values, n * 4, cudaMemcpyHostToDevice, v->stream));
    int rc = sort_pairs_device(v, dk, dv, n, bits, wide);
    if (rc == B200GS_OK) {
        CK(cudaMemcpyAsync(keys, dk, n * 4, cudaMemcpyDeviceToHost, v->stream));
        CK(cudaMemcpyAsync(values, dv, n * 4, cudaMemcpyDeviceToHost, v->stream));
        CK(cudaStreamSynchronize(v->stream));
    }
    cudaFree(dk); cudaFree(dv);
    return rc;
}
extern "C" int b200gs_sort_pairs_host(b200gs_viewer* v, uint32_t* keys, uint32_t* values, uint64_t n, uint32_t bits) {
    return sort_pairs_host(v, keys, values, n, bits, false);
}
extern "C" int b200gs_sort_pairs_wide_host(b200gs_viewer* v, uint32_t* keys, uint32_t* values, uint64_t n, uint32_t bits) {
    return sort_pairs_host(v, keys, values, n, bits, true);
}
