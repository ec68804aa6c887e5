// aux.cu — event-driven side kernels next to the hot path.
//
//  * mask evaluation: mask_evaluator.evaluate(device, queue, &MaskOpTree, &mask_buffer,
//    &model_transform_buffer, &gaussians_buffer) — reference src/tab/scene.rs:2124-2131,
//    2201-2209; tree built at src/app.rs:1816-1837.  The tree arrives flattened to postfix.
//  * postprocess: postprocessor.postprocess(...) — reference src/tab/scene.rs:604-610: commit
//    the viewer's selection edit into the per-Gaussian edit buffer of the selected Gaussians.
//  * query texture painting: query_toolset.render(queue, encoder, query_texture) — scene.rs:787-791.
// Compiled with -fmad=false like preprocess.cu so the point-in-shape tests match the oracle.
#include "common.cuh"

namespace {

constexpr int kMaxOps = 64;

__global__ void __launch_bounds__(256) k_eval_mask(const uint8_t* __restrict__ recs, uint32_t n, uint32_t rb,
                                                   const __grid_constant__ GsModelXf m,
                                                   const b200gs_mask_op* __restrict__ ops, uint32_t n_ops,
                                                   const b200gs_mask_shape* __restrict__ shapes,
                                                   const float* __restrict__ shape_rot, uint32_t* words) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    bool shown = false;
    if (i < n) {
        const float* p = reinterpret_cast<const float*>(recs + (size_t)i * rb);
        float ps0 = m.s[0] * p[0], ps1 = m.s[1] * p[1], ps2 = m.s[2] * p[2];
        float pw[3];
#pragma unroll
        for (int r = 0; r < 3; r++) pw[r] = m.R[r][0] * ps0 + m.R[r][1] * ps1 + m.R[r][2] * ps2 + m.t[r];
        uint64_t stack = 0;  // bit stack
        int sp = 0;
        if (n_ops == 0) { stack = 1; sp = 1; }
        for (uint32_t o = 0; o < n_ops; o++) {
            const uint32_t k = ops[o].kind;
            if (k == B200GS_MASKOP_RESET) { stack |= 1ull << sp; sp++; continue; }
            if (k == B200GS_MASKOP_SHAPE) {
                const b200gs_mask_shape& sh = shapes[ops[o].arg];
                const float* Q = shape_rot + 9 * ops[o].arg;
                float d0 = pw[0] - sh.pos[0], d1 = pw[1] - sh.pos[1], d2 = pw[2] - sh.pos[2];
                float l[3];
#pragma unroll
                for (int c = 0; c < 3; c++) l[c] = (Q[0 * 3 + c] * d0 + Q[1 * 3 + c] * d1 + Q[2 * 3 + c] * d2) / sh.scale[c];
                bool in;
                if (sh.kind == B200GS_MASK_BOX) in = fabsf(l[0]) <= 0.5f && fabsf(l[1]) <= 0.5f && fabsf(l[2]) <= 0.5f;
                else in = (l[0] * l[0] + l[1] * l[1] + l[2] * l[2]) <= 0.25f;
                stack = (stack & ~(1ull << sp)) | ((uint64_t)in << sp);
                sp++;
                continue;
            }
            if (k == B200GS_MASKOP_COMPLEMENT) { stack ^= 1ull << (sp - 1); continue; }
            bool b = (stack >> (sp - 1)) & 1ull, a = (stack >> (sp - 2)) & 1ull, r = false;
            sp -= 2;
            if (k == B200GS_MASKOP_UNION) r = a | b;
            else if (k == B200GS_MASKOP_INTERSECTION) r = a & b;
            else if (k == B200GS_MASKOP_DIFFERENCE) r = a & !b;
            else if (k == B200GS_MASKOP_SYMDIFF) r = a ^ b;
            stack = (stack & ~(1ull << sp)) | ((uint64_t)r << sp);
            sp++;
        }
        shown = sp > 0 && ((stack >> (sp - 1)) & 1ull);
    }
    uint32_t w = __ballot_sync(0xffffffffu, shown);
    if ((threadIdx.x & 31) == 0 && i < n) words[i >> 5] = w;
}

__global__ void __launch_bounds__(256) k_postprocess(uint32_t n, const uint32_t* __restrict__ selection,
                                                     b200gs_edit_pod* edits, const __grid_constant__ b200gs_edit_pod e) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    if ((selection[i >> 5] >> (i & 31)) & 1u) edits[i] = e;
}

__global__ void __launch_bounds__(256) k_fill_edits(uint32_t n, b200gs_edit_pod* edits) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    b200gs_edit_pod e;
    e.flag = 0; e.color[0] = 0.0f; e.color[1] = 1.0f; e.color[2] = 1.0f;
    e.contrast = 0.0f; e.exposure = 0.0f; e.gamma = 1.0f; e.alpha = 1.0f;
    edits[i] = e;
}

// query_toolset.render(queue, encoder, query_texture) (scene.rs:787-791): one stroke painted into the query texture
__global__ void __launch_bounds__(256) k_paint_query_texture(uint8_t* tex, uint32_t w, uint32_t h,
                                                             const __grid_constant__ b200gs_query_pod q) {
    const uint32_t x = blockIdx.x * 32 + (threadIdx.x & 31), y = blockIdx.y * 8 + (threadIdx.x >> 5);
    if (x >= w || y >= h) return;
    if (gs_query_shape_hit(q, (float)x + 0.5f, (float)y + 0.5f)) tex[(size_t)y * w + x] = 255;
}

}  // namespace

cudaError_t gs_launch_paint_query_texture(uint8_t* tex, uint32_t w, uint32_t h, const b200gs_query_pod& stroke, cudaStream_t st) {
    if (w == 0 || h == 0) return cudaSuccess;
    k_paint_query_texture<<<dim3((w + 31) / 32, (h + 7) / 8), 256, 0, st>>>(tex, w, h, stroke);
    return cudaGetLastError();
}

cudaError_t gs_launch_eval_mask(const uint8_t* recs, uint32_t n, uint32_t record_bytes, const GsModelXf& m,
                                const b200gs_mask_op* ops_dev, uint32_t n_ops, const b200gs_mask_shape* shapes_dev,
                                const float* shape_rot_dev, uint32_t* words, cudaStream_t st) {
    if (n_ops > kMaxOps) return cudaErrorInvalidValue;
    if (n == 0) return cudaSuccess;
    k_eval_mask<<<(n + 255) / 256, 256, 0, st>>>(recs, n, record_bytes, m, ops_dev, n_ops, shapes_dev, shape_rot_dev, words);
    return cudaGetLastError();
}

cudaError_t gs_launch_postprocess(uint32_t n, const uint32_t* selection, b200gs_edit_pod* edits,
                                  b200gs_edit_pod sel_edit, cudaStream_t st) {
    if (n == 0) return cudaSuccess;
    k_postprocess<<<(n + 255) / 256, 256, 0, st>>>(n, selection, edits, sel_edit);
    return cudaGetLastError();
}

cudaError_t gs_launch_fill_default_edits(uint32_t n, b200gs_edit_pod* edits, cudaStream_t st) {
    if (n == 0) return cudaSuccess;
    k_fill_edits<<<(n + 255) / 256, 256, 0, st>>>(n, edits);
    return cudaGetLastError();
}
