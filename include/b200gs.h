/*
 * b200gs.h — C ABI of the B200-native 3D Gaussian splatting render core.
 *
 * This is the drop-in boundary for the per-frame hot path that
 * LioQing/wgpu-3dgs-viewer-app drives through the crate `wgpu-3dgs-viewer`
 * (imported `as gs`): preprocess -> depth sort -> splat compositing.
 * Every entry point cites the reference call site (file:line under the
 * reference tree) whose `gs::` call it replaces.  All functions return a
 * status (`B200GS_OK` == 0 on success); `b200gs_last_error()` gives the
 * thread-local message for the last failure.  No torch types, plain
 * pointers and sizes only.  There is NO CPU fallback behind this ABI: every
 * compute entry point fails with B200GS_ERR_CUDA when no sm_100 device is
 * usable.
 *
 * Conventions (reference evidence in brackets):
 *  - matrices are column-major float[16] exactly as glam's Mat4::to_cols_array
 *    [src/app.rs:1236-1244]; view = look_at_rh, proj = perspective_rh
 *    (depth 0..1);
 *  - quaternions are (x,y,z,w) as glam's Quat [src/app.rs:1123-1130];
 *  - model transform: world = quat * (scale ⊙ p) + pos [src/app.rs:1044-1046];
 *  - bitsets (mask, selection) are ceil(N/32) u32 words, bit (i & 31) of word
 *    (i >> 5) belongs to Gaussian i [src/app.rs:626, 806]; mask bit 1 = shown;
 *  - images are RGBA8, premultiplied alpha, row 0 = top (NDC y = +1).
 */
#ifndef B200GS_H
#define B200GS_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define B200GS_API __attribute__((visibility("default")))

/* ------------------------------------------------------------------ status */
enum {
    B200GS_OK = 0,
    B200GS_ERR_INVALID = 1,  /* bad argument (null handle, range, layout ...)  */
    B200GS_ERR_CUDA = 2,     /* CUDA runtime/driver failure or no device        */
    B200GS_ERR_OOM = 3,      /* device or host allocation failed                */
    B200GS_ERR_IO = 4,       /* PLY / file error (gs::Error::Io, scene.rs:234)  */
    B200GS_ERR_FORMAT = 5,   /* malformed PLY                                    */
    B200GS_ERR_OVERFLOW = 6  /* tile-entry capacity exceeded in the last frame: returned by b200gs_sync,
                              * b200gs_render_frame_host and _host_end AFTER the (truncated) image was delivered */
};

/* ------------------------------------------------ record layouts (gs::GaussianPod)
 * The 8 `GaussianPodWithSh{Single,Half,Norm8,None}Cov3d{Single,Half}Configs`
 * of src/app.rs:250-257.  A packed record is
 *     pos f32x3 | color u8x4 | SH field | Cov3d field        (scene.rs:907-978)
 * SH field   : Single = 45 f32 (15 x Vec3, coefficient-major: sh[k] = (r,g,b))
 *              Half   = 46 f16 (45 used + 1 pad)
 *              Norm8  = 48 u8  (45 used + 3 pad), value = q/255*2-1, range [-1,1]
 *              None   = nothing (SH bands 1..3 read as 0)
 * Cov3d field: upper triangle (xx,xy,xz,yy,yz,zz) of R S S^T R^T as 6 f32 / 6 f16
 * color      : rgb = clamp(0.5 + 0.2820948*f_dc), a = sigmoid(opacity), each x255 RN.
 */
typedef enum { B200GS_SH_SINGLE = 0, B200GS_SH_HALF = 1, B200GS_SH_NORM8 = 2, B200GS_SH_NONE = 3 } b200gs_sh_config;
typedef enum { B200GS_COV3D_SINGLE = 0, B200GS_COV3D_HALF = 1 } b200gs_cov3d_config;

/* Unpacked Gaussian, the mirror of `gs::Gaussian` (src/app.rs:512, 1066). 224 B. */
typedef struct b200gs_gaussian {
    float rot[4];     /* unit quaternion x,y,z,w */
    float pos[3];
    uint8_t color[4]; /* rgb = SH0 colour, a = opacity */
    float sh[45];     /* bands 1..3, sh[3*k + c], k = 0..14, c = r,g,b */
    float scale[3];   /* linear (already exp'ed) */
} b200gs_gaussian;

/* One PLY vertex of the Inria 3DGS format, mirror of `gs::PlyGaussianPod`
 * (scene.rs:997, metadata.rs:54): 62 f32 = 248 B. */
typedef struct b200gs_ply_gaussian {
    float pos[3];
    float normal[3];
    float f_dc[3];
    float f_rest[45]; /* channel-major: 15 R, 15 G, 15 B */
    float opacity;    /* logit */
    float scale[3];   /* log */
    float rot[4];     /* w,x,y,z (not normalised) */
} b200gs_ply_gaussian;

/* gs::GaussianDisplayMode (src/tab/transform.rs:129-131) */
typedef enum { B200GS_DISPLAY_SPLAT = 0, B200GS_DISPLAY_ELLIPSE = 1, B200GS_DISPLAY_POINT = 2 } b200gs_display_mode;

/* gs::GaussianEditFlag (src/app.rs:1548-1553) */
enum { B200GS_EDIT_ENABLED = 1u, B200GS_EDIT_HIDDEN = 2u, B200GS_EDIT_OVERRIDE_COLOR = 4u };

/* gs::GaussianEditPod (src/app.rs:1556-1563; ranges src/tab/selection.rs:172-204). 32 B.
 * Default (flag = 0, hsv = (0,1,1), contrast 0, exposure 0, gamma 1, alpha 1) is a no-op
 * (scene.rs:821, 848). */
typedef struct b200gs_edit_pod {
    uint32_t flag;
    float color[3]; /* HSV (h add 0..1, s mul 0..2, v mul 0..2) or override RGB */
    float contrast; /* -1..1  */
    float exposure; /* -5..5  */
    float gamma;    /*  0..5  */
    float alpha;    /*  0..2  */
} b200gs_edit_pod;

/* gs::Query*Pod (scene.rs:1622, 1633; QueryToolset rect/brush scene.rs:1260-1263). */
typedef enum {
    B200GS_QUERY_NONE = 0, B200GS_QUERY_HIT = 1,
    B200GS_QUERY_RECT = 2, B200GS_QUERY_BRUSH = 3, /* immediate mode: the shape itself is tested (set_use_texture(false)) */
    B200GS_QUERY_TEXTURE = 4                       /* non-immediate: the viewer's query texture is sampled (scene.rs:767-791) */
} b200gs_query_kind;
typedef enum { B200GS_SELECT_SET = 0, B200GS_SELECT_ADD = 1, B200GS_SELECT_REMOVE = 2 } b200gs_selection_op;
typedef struct b200gs_query_pod {
    uint32_t kind;   /* b200gs_query_kind */
    uint32_t op;     /* b200gs_selection_op (rect / brush)                         */
    float p0[2];     /* hit: pixel; rect: top-left; brush: segment start (pixels)  */
    float p1[2];     /* rect: bottom-right; brush: segment end                     */
    float radius;    /* brush radius in pixels                                     */
    uint32_t _pad;
} b200gs_query_pod;

/* gs::MaskShapeKind / gs::MaskOpShapePod / gs::MaskOpTree (src/app.rs:1816-1837,
 * src/tab/mask.rs:141-230).  The op tree is passed flattened in postfix order. */
typedef enum { B200GS_MASK_BOX = 0, B200GS_MASK_ELLIPSOID = 1 } b200gs_mask_shape_kind;
typedef struct b200gs_mask_shape {
    uint32_t kind;
    float pos[3];
    float quat[4]; /* x,y,z,w */
    float scale[3]; /* full extents of the box / diameters of the ellipsoid: the unit
                       shape is the cube [-0.5,0.5]^3 or the sphere of radius 0.5 */
} b200gs_mask_shape;
typedef enum {
    B200GS_MASKOP_SHAPE = 0,      /* push shape[arg]           */
    B200GS_MASKOP_UNION = 1,      /* a | b                     */
    B200GS_MASKOP_INTERSECTION = 2, /* a & b                   */
    B200GS_MASKOP_DIFFERENCE = 3, /* a & ~b                    */
    B200GS_MASKOP_SYMDIFF = 4,    /* a ^ b                     */
    B200GS_MASKOP_COMPLEMENT = 5, /* ~a                        */
    B200GS_MASKOP_RESET = 6       /* push 1 (everything shown) */
} b200gs_mask_op_kind;
typedef struct b200gs_mask_op { uint32_t kind; uint32_t arg; } b200gs_mask_op;

/* Projected splat written by the preprocess kernel, one per VISIBLE Gaussian (parity tap;
 * also the compositor's input).  32 B, two 16-byte halves. */
typedef struct b200gs_splat {
    float mx, my;        /* centre in pixel-centre coordinates (pixel i is at i)  */
    uint16_t radius;     /* ceil(3*sqrt(lambda_max)), clamped to 65535; 0 = empty */
    uint16_t opacity_h;  /* f16                                                    */
    uint16_t r_h, g_h;   /* f16 colour                                             */
    float ca, cb, cc;    /* conic (inverse 2-D covariance: a, b, c)               */
    uint16_t b_h;        /* f16 colour                                             */
    uint16_t flags;      /* bit0 selected                                          */
} b200gs_splat;

/* One entry of the per-pixel hit list (mirror of gs::QueryHitResultPod, src/tab/scene.rs:650-657):
 * the splats that contribute to a pixel, front to back. */
typedef struct b200gs_hit {
    uint32_t model;   /* position of the model in the far_to_near array passed to the render   */
    uint32_t index;   /* Gaussian index inside that model                                       */
    float alpha;      /* the splat's alpha at the pixel (>= 1/255)                              */
    float depth;      /* ndc.z of the Gaussian (the depth key as a float)                       */
} b200gs_hit;

/* Per-stage device times of the last b200gs_render_frame / explicit stage calls (ms). */
typedef struct b200gs_timings {
    float preprocess_ms, sort_ms, bin_ms, composite_ms, total_ms;
    uint64_t visible;      /* sum over models of V                                  */
    uint64_t tile_entries; /* duplicated (tile, splat) entries binned               */
    uint64_t evals;        /* splat-pixel evaluations (only when counting is enabled) */
    uint64_t staged_entries; /* tile entries the compositor read before its tiles finished (counting only) */
    uint32_t overflow;     /* 1 = tile-entry capacity was exceeded (entries dropped)  */
    uint32_t _pad;
} b200gs_timings;

typedef struct b200gs_viewer b200gs_viewer;
typedef struct b200gs_model b200gs_model;

/* ----------------------------------------------------------------- library */
B200GS_API const char* b200gs_last_error(void);
B200GS_API const char* b200gs_version(void);
B200GS_API int b200gs_device_count(int* out);
/* bytes of one packed record of the layout; 0 if the layout is invalid */
B200GS_API uint32_t b200gs_record_bytes(uint32_t sh, uint32_t cov3d);

/* process-wide tuning knobs (set BEFORE viewers are created) and read-only facts for benches / profiles.
 * knobs: "sort.cluster" = CTAs per thread-block cluster of the radix sort (8 default, 4, 2, 1 = no clusters);
 *        "sort.claim" = 0 (default) / 1: collision-free fast path of the ranking for 10/11-bit digits.
 * info:  "sort.cluster", "sort.resident_clusters" (after the first sort on the viewer's device), "num_sms". */
B200GS_API int b200gs_set_tuning(const char* name, int64_t value);
B200GS_API int b200gs_get_info(b200gs_viewer* v, const char* name, int64_t* out);

/* ------------------------------------------------------------------ viewer
 * gs::MultiModelViewer::<G>::new_with(device, format, depth_stencil, uvec2)  scene.rs:1969-1980 */
B200GS_API int b200gs_viewer_create(int device, uint32_t sh, uint32_t cov3d, uint32_t width, uint32_t height,
                                    b200gs_viewer** out);
B200GS_API int b200gs_viewer_destroy(b200gs_viewer* v);
/* viewer.update_query_texture_size(device, size)  scene.rs:740 — also the render-target size */
B200GS_API int b200gs_resize(b200gs_viewer* v, uint32_t width, uint32_t height);
/* viewer.update_camera(queue, &impl CameraTrait, size)  scene.rs:795; CameraPod{view,proj,size}
 * src/shader/measurement.wgsl:14-19 */
B200GS_API int b200gs_set_camera(b200gs_viewer* v, const float view[16], const float proj[16], const float size[2]);
/* viewer.update_gaussian_transform(queue, size, display_mode, sh_deg, no_sh0)  scene.rs:803-809 */
B200GS_API int b200gs_set_gaussian_transform(b200gs_viewer* v, float size, uint32_t display_mode, uint32_t sh_deg,
                                             uint32_t no_sh0);
/* viewer.update_selection_edit_with_pod(queue, &pod)  scene.rs:815, 821, 848 */
B200GS_API int b200gs_set_selection_edit(b200gs_viewer* v, const b200gs_edit_pod* pod);
/* viewer.update_selection_highlight(queue, vec4) / _with_pod  scene.rs:816, 822-829, 833 */
B200GS_API int b200gs_set_selection_highlight(b200gs_viewer* v, const float rgba[4]);
/* viewer.update_query(queue, pod)  scene.rs:785 */
B200GS_API int b200gs_set_query(b200gs_viewer* v, const b200gs_query_pod* pod);
/* The query texture of gs::QueryToolset in non-immediate mode (query_toolset.set_use_texture(!immediate) +
 * query_toolset.render(queue, encoder, &viewer.world_buffers.query_texture)  scene.rs:767-791; sized by
 * update_query_texture_size  scene.rs:740): one u8 per viewport pixel, non-zero = painted.  _paint rasterises one
 * rect / brush-segment stroke (a B200GS_QUERY_RECT / _BRUSH pod) into it, _clear empties it (toolset.start), _upload
 * replaces it with host texels (width x height must equal the viewport), _download reads it back.  A preprocess
 * with query kind B200GS_QUERY_TEXTURE selects (op Set / Add / Remove) the visible Gaussians whose projected centre
 * falls on a painted texel.  Resizing the viewer clears the texture. */
B200GS_API int b200gs_query_texture_clear(b200gs_viewer* v);
B200GS_API int b200gs_query_texture_paint(b200gs_viewer* v, const b200gs_query_pod* stroke);
B200GS_API int b200gs_query_texture_upload(b200gs_viewer* v, const uint8_t* texels, uint32_t width, uint32_t height);
B200GS_API int b200gs_query_texture_download(b200gs_viewer* v, uint8_t* texels, size_t cap);
/* headless clear colour (premultiplied RGBA in 0..1); default transparent black */
B200GS_API int b200gs_set_background(b200gs_viewer* v, const float rgba[4]);
/* capacity of the (bin, splat) entry list per frame (one entry per 32x32-pixel bin a splat touches); default 8 x total
 * Gaussian capacity */
B200GS_API int b200gs_set_tile_entry_capacity(b200gs_viewer* v, uint64_t entries);
/* record per-stage CUDA-event times (and optionally count splat evaluations) for
 * b200gs_last_timings; off by default */
B200GS_API int b200gs_enable_timings(b200gs_viewer* v, int on, int count_evals);
/* queue.submit + device.poll(Maintain::Wait)  scene.rs:613-614, 872-873 */
B200GS_API int b200gs_sync(b200gs_viewer* v);
/* the viewer's cudaStream_t (as void*): all work of a viewer is ordered on it */
B200GS_API void* b200gs_stream(b200gs_viewer* v);
/* the viewer's own RGBA8 render target (device pointer, width*4 pitch), used by
 * b200gs_render_frame_host; valid until the next resize */
B200GS_API void* b200gs_image_device(b200gs_viewer* v);
/* page-locked host memory for images / uploads (cudaMallocHost) */
B200GS_API int b200gs_host_alloc(size_t bytes, void** out);
B200GS_API int b200gs_host_free(void* p);

/* ------------------------------------------------------------------- models
 * MultiModelViewerGaussianBuffers::new_empty(device,count) + BindGroups::new +
 * viewer.models.insert(key, ..) + MaskOpTree::Reset  scene.rs:2111-2139 */
B200GS_API int b200gs_model_create(b200gs_viewer* v, const char* key, uint64_t capacity, b200gs_model** out);
/* a model of `v` over the packed records already resident in `source` (a model of ANOTHER viewer on the same device
 * with the same layout): the reference's buffers are ref-counted clones (scene.rs:641, 648); nothing is copied,
 * the records live until the last model using them is destroyed.  Mask / selection / edits stay per model. */
B200GS_API int b200gs_model_create_shared(b200gs_viewer* v, const char* key, b200gs_model* source, b200gs_model** out);
/* viewer.remove_model(key)  scene.rs:2176 */
B200GS_API int b200gs_model_destroy(b200gs_viewer* v, b200gs_model* m);
B200GS_API b200gs_model* b200gs_model_find(b200gs_viewer* v, const char* key);
/* gaussians_buffer.len()  scene.rs:608, 862, 1832 */
B200GS_API uint64_t b200gs_model_len(const b200gs_model* m);
/* gaussians_buffer.update_range(queue, start, &[Gaussian])  scene.rs:2076-2084 — unpacked
 * Gaussians are packed on the host into the viewer's layout, then copied H2D.  Like queue.write_buffer, every upload
 * below ENQUEUES and returns: the host buffer is borrowed for the call only (it is packed / copied into the viewer's
 * pinned staging ring before the call returns), nothing is allocated per call and the stream is not synchronised,
 * so the app can call this every frame while a model loads (scene.rs:341-380). */
B200GS_API int b200gs_model_update_range(b200gs_model* m, uint64_t start, const b200gs_gaussian* gaussians,
                                         uint64_t count);
/* same, records already packed (host pointer) */
B200GS_API int b200gs_model_upload_packed(b200gs_model* m, uint64_t start, const void* packed, uint64_t count);
/* same, records already packed and resident on this device (e.g. after an NCCL broadcast) */
B200GS_API int b200gs_model_upload_packed_device(b200gs_model* m, uint64_t start, const void* packed_dev,
                                                 uint64_t count);
/* viewer.update_model_transform(queue, key, pos, quat, scale)  scene.rs:796-802 */
B200GS_API int b200gs_model_set_transform(b200gs_model* m, const float pos[3], const float quat_xyzw[4],
                                          const float scale[3]);
/* mask / selection / edit buffers of MultiModelViewerGaussianBuffers  scene.rs:2111-2139 */
B200GS_API int b200gs_model_upload_mask(b200gs_model* m, const uint32_t* words, uint64_t nwords);
B200GS_API int b200gs_model_upload_selection(b200gs_model* m, const uint32_t* words, uint64_t nwords);
B200GS_API int b200gs_model_upload_edits(b200gs_model* m, uint64_t start, const b200gs_edit_pod* pods, uint64_t count);
/* mask_evaluator.evaluate(device, queue, &tree, &mask, &model_transform, &gaussians)
 * scene.rs:2124-2131, 2201-2209; n_ops == 0 or a single RESET op = MaskOpTree::Reset */
B200GS_API int b200gs_model_eval_mask(b200gs_model* m, const b200gs_mask_op* postfix, uint32_t n_ops,
                                      const b200gs_mask_shape* shapes, uint32_t n_shapes);
/* postprocessor.postprocess(encoder, bg.0, bg.1, count, args)  scene.rs:604-610: commit the
 * viewer's selection edit into the per-Gaussian edit buffer for the selected Gaussians */
B200GS_API int b200gs_model_postprocess(b200gs_model* m);

/* ------------------------------------------------------------- the hot path
 * preprocessor.preprocess(encoder, bind_group, N)  scene.rs:856-863 (use_unedited selects
 * UneditedModel's blank edit buffer, scene.rs:858-861).  Enqueue only, no host sync. */
B200GS_API int b200gs_model_preprocess(b200gs_model* m, int use_unedited);
/* radix_sorter.sort(encoder, bind_group, radix_sort_indirect_args)  scene.rs:865-869 */
B200GS_API int b200gs_model_sort(b200gs_model* m);
/* renderer.render_with_pass(pass, bind_group, indirect_args) over model_render_keys
 * scene.rs:2302-2314: models are given FARTHEST FIRST (scene.rs:533-558) and layered whole.
 * rgba8_out is a DEVICE pointer to height rows of pitch bytes.  Enqueue only. */
B200GS_API int b200gs_render(b200gs_viewer* v, b200gs_model* const* far_to_near, uint32_t n_models, void* rgba8_out,
                             size_t pitch);
/* preprocess + sort for each model, then render: one whole frame, enqueue only */
B200GS_API int b200gs_render_frame(b200gs_viewer* v, b200gs_model* const* far_to_near, uint32_t n_models,
                                   void* rgba8_out, size_t pitch);
/* whole frame from HOST state to a HOST image: sets the camera, renders, copies the image to
 * `rgba8_host` (width*4 pitch) and waits.  This is the end-to-end call a headless caller makes. */
B200GS_API int b200gs_render_frame_host(b200gs_viewer* v, b200gs_model* const* far_to_near, uint32_t n_models,
                                        const float view[16], const float proj[16], void* rgba8_host);
/* pipelined variant: _begin enqueues frame + D2H copy (second stream) and returns; _end waits for
 * the OLDEST frame in flight (at most two).  rgba8_host should be page-locked (b200gs_host_alloc). */
B200GS_API int b200gs_render_frame_host_begin(b200gs_viewer* v, b200gs_model* const* far_to_near, uint32_t n_models,
                                              const float view[16], const float proj[16], void* rgba8_host);
B200GS_API int b200gs_render_frame_host_end(b200gs_viewer* v);
/* number of CUDA kernels this viewer has launched so far */
B200GS_API int b200gs_launch_count(b200gs_viewer* v, uint64_t* out);
/* sort models by squared distance of `world_center` to the camera, farthest first
 * (scene.rs:533-558).  centers: n x 3 model-space centres; order_out: n indices. */
B200GS_API int b200gs_order_models(b200gs_viewer* v, b200gs_model* const* models, const float* centers, uint32_t n,
                                   uint32_t* order_out);

/* -------------------------------------------------------- downloads / taps
 * buffer.download(&device,&queue)  src/app.rs:789, 806 (edits, mask); the rest are parity taps
 * with no reference equivalent.  All of them synchronise the viewer's stream. */
B200GS_API int b200gs_model_download_mask(b200gs_model* m, uint32_t* words, uint64_t cap_words, uint64_t* n);
B200GS_API int b200gs_model_download_selection(b200gs_model* m, uint32_t* words, uint64_t cap_words, uint64_t* n);
B200GS_API int b200gs_model_download_edits(b200gs_model* m, b200gs_edit_pod* pods, uint64_t cap, uint64_t* n);
B200GS_API int b200gs_model_download_packed(b200gs_model* m, uint64_t start, void* packed, uint64_t count);
B200GS_API int b200gs_model_visible_count(b200gs_model* m, uint64_t* out);
/* keys/indices as currently stored: after preprocess = compaction order (ascending index),
 * after sort = ascending depth key */
B200GS_API int b200gs_model_download_depth_keys(b200gs_model* m, uint32_t* keys, uint64_t cap, uint64_t* n);
B200GS_API int b200gs_model_download_indices(b200gs_model* m, uint32_t* idx, uint64_t cap, uint64_t* n);
/* projected splats in the same order as the indices */
B200GS_API int b200gs_model_download_splats(b200gs_model* m, b200gs_splat* out, uint64_t cap, uint64_t* n);
B200GS_API int b200gs_last_timings(b200gs_viewer* v, b200gs_timings* out);
/* gs::QueryHitPod + gs::query::download (src/tab/scene.rs:617-657): the ordered hit list of pixel
 * (px, py) of the LAST rendered frame (same models, same camera), nearest first, at most `cap`. */
B200GS_API int b200gs_query_hits(b200gs_viewer* v, b200gs_model* const* far_to_near, uint32_t n_models, uint32_t px,
                                 uint32_t py, b200gs_hit* out, uint64_t cap, uint64_t* n);
/* gs::query::hit_pos_by_closest / hit_pos_by_alpha_range (src/tab/scene.rs:659-676): world position of
 * the picked surface point on the ray through the pixel.  closest = depth of the first hit;
 * alpha_range = alpha-weighted mean depth of the hits with alpha >= threshold (0.05 in the app).
 * Returns B200GS_ERR_INVALID if no hit qualifies. */
B200GS_API int b200gs_hit_pos_by_closest(const b200gs_hit* hits, uint64_t n, const float view[16], const float proj[16],
                                         const float size[2], uint32_t px, uint32_t py, float pos_out[3]);
B200GS_API int b200gs_hit_pos_by_alpha_range(const b200gs_hit* hits, uint64_t n, float alpha_threshold, const float view[16],
                                             const float proj[16], const float size[2], uint32_t px, uint32_t py,
                                             float pos_out[3]);
/* raw sort entry point (keys/values DEVICE arrays of n u32, sorted ascending & stable in place;
 * bits = number of low key bits to sort, multiple of 8).  Used by the sort parity tests. */
B200GS_API int b200gs_sort_pairs_device(b200gs_viewer* v, uint32_t* keys_dev, uint32_t* values_dev, uint64_t n,
                                        uint32_t bits);
B200GS_API int b200gs_sort_pairs_host(b200gs_viewer* v, uint32_t* keys, uint32_t* values, uint64_t n, uint32_t bits);
/* the same through the 11-bit-digit cluster sort that orders the bin ids of the binning stage (bits = 1..32) */
B200GS_API int b200gs_sort_pairs_wide_device(b200gs_viewer* v, uint32_t* keys_dev, uint32_t* values_dev, uint64_t n,
                                             uint32_t bits);
B200GS_API int b200gs_sort_pairs_wide_host(b200gs_viewer* v, uint32_t* keys, uint32_t* values, uint64_t n, uint32_t bits);

/* -------------------------------------------- host side (no GPU required)
 * Packing: GaussiansBuffer::update_range's host half (scene.rs:2069-2085). */
B200GS_API int b200gs_pack_gaussians(uint32_t sh, uint32_t cov3d, const b200gs_gaussian* in, uint64_t count, void* out);
B200GS_API int b200gs_unpack_gaussians(uint32_t sh, uint32_t cov3d, const void* in, uint64_t count, b200gs_gaussian* out);
/* Gaussian::from(PlyGaussianPod)  src/app.rs:1066 */
B200GS_API int b200gs_gaussian_from_ply(const b200gs_ply_gaussian* in, uint64_t count, b200gs_gaussian* out);
B200GS_API int b200gs_gaussian_to_ply(const b200gs_gaussian* in, uint64_t count, b200gs_ply_gaussian* out);
/* Gaussians::read_ply_header + header.count()  src/app.rs:1056-1057 */
typedef struct b200gs_ply_reader b200gs_ply_reader;
B200GS_API int b200gs_ply_open(const char* path, b200gs_ply_reader** out, uint64_t* count);
B200GS_API int b200gs_ply_open_memory(const void* data, size_t size, b200gs_ply_reader** out, uint64_t* count);
/* Gaussians::read_ply_gaussians iterator  src/app.rs:1062-1070: read up to max vertices */
B200GS_API int b200gs_ply_read(b200gs_ply_reader* r, b200gs_ply_gaussian* out, uint64_t max, uint64_t* n_read);
B200GS_API int b200gs_ply_close(b200gs_ply_reader* r);
/* Gaussians::write_ply  src/app.rs:910-914 (binary little endian) */
B200GS_API int b200gs_ply_write(const char* path, const b200gs_ply_gaussian* verts, uint64_t count);
/* Export with edits and mask: Gaussians::write_ply(writer, setting.edit.then_some(&edits), setting.mask.then_some(mask))
 * src/app.rs:904-914, 935-943 (pods / mask words as downloaded at app.rs:789, 806; either may be NULL = None).
 * A Gaussian whose mask bit is clear, or whose pod is ENABLED|HIDDEN, is not written; an ENABLED pod is applied to the
 * Gaussian's base colour and opacity (colour -> contrast -> exposure -> gamma -> alpha, as the preprocess kernel
 * shows it) and re-quantised to u8 before Gaussian -> PlyGaussianPod; SH bands 1..3 are written unchanged.
 * b200gs_apply_edits_for_export is the in-memory half (out holds up to `count` Gaussians). */
B200GS_API int b200gs_ply_write_edited(const char* path, const b200gs_gaussian* gaussians, uint64_t count,
                                       const b200gs_edit_pod* edits_or_null, const uint32_t* mask_words_or_null);
B200GS_API int b200gs_apply_edits_for_export(const b200gs_gaussian* in, uint64_t count, const b200gs_edit_pod* edits_or_null,
                                             const uint32_t* mask_words_or_null, b200gs_gaussian* out, uint64_t* n_out);
/* camera helpers: glam look_at_rh / perspective_rh as the app uses them  src/app.rs:1236-1244 */
B200GS_API void b200gs_look_at_rh(const float eye[3], const float target[3], const float up[3], float out[16]);
B200GS_API void b200gs_perspective_rh(float vfov, float aspect, float z_near, float z_far, float out[16]);
/* Quat::from_euler(EulerRot::ZYX, rz, ry, rx) of degrees  src/app.rs:1123-1130 */
B200GS_API void b200gs_quat_from_euler_zyx_deg(const float rot_deg[3], float quat_xyzw[4]);
/* deterministic synthetic scene of SURVEY.md §8d (bench / test input generator) */
B200GS_API int b200gs_synth_scene(uint64_t seed, uint64_t start, uint64_t count, b200gs_ply_gaussian* out);

#ifdef __cplusplus
}
#endif
#endif /* B200GS_H */
