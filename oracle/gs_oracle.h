/*
 * gs_oracle.h — CPU ORACLE (test infrastructure, NOT product code).
 *
 * A plain-C restatement of the per-frame hot path the reference app drives through the
 * crate `wgpu-3dgs-viewer 0.2.0` (preprocess -> depth sort -> splat render), used ONLY as
 * the checker by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / reference
 * arm.  Nothing in the product path (the CUDA library) links, loads or calls this file.
 *
 * PARITY UNPINNED: the crate's source (Cargo.lock:3731-3734) is not under /root/reference,
 * the reference tree holds no tests or golden vectors for this path,
 * and it cannot be built here (no rustc/cargo, no wgpu adapter).  The oracle therefore
 * follows (i) what the app itself pins [APP] and (ii) canonical 3DGS maths (Kerbl et al.
 * 2023, Inria diff-gaussian-rasterization) [CANON] for what lives inside the crate; every
 * such constant is a named macro below so it can be flipped when the crate is readable.
 *
 * Compile with -ffp-contract=off: every float operation rounds separately, in source order.
 */
#ifndef GS_ORACLE_H
#define GS_ORACLE_H

#include "../include/b200gs.h" /* POD types of the boundary only */

#ifdef __cplusplus
extern "C" {
#endif

/* ---- named constants of the restatement ---- */
#define ORC_CULL_XY 1.3f          /* [RECALLED] NDC x/y cull margin                              */
#define ORC_CLAMP_XY 1.3f         /* [CANON] clamp of x/z, y/z to 1.3 tan(fov/2) in the Jacobian  */
#define ORC_LOWPASS 0.3f          /* [CANON] added to the cov2d diagonal                          */
#define ORC_EXTENT_SIGMA 3.0f     /* [CANON] quad extent = ceil(3 sqrt(lambda_max))               */
#define ORC_MIN_DISC 0.1f         /* [CANON] max(0.1, mid^2 - det)                                */
#define ORC_ALPHA_MAX 0.99f       /* [CANON] alpha clamp                                          */
#define ORC_ALPHA_MIN (1.0f / 255.0f) /* [CANON] alpha cut                                        */
#define ORC_T_EPS (1.0f / 1024.0f) /* front-to-back termination threshold (our design, §7.4)      */
#define ORC_FLAT_D2 4.0f          /* Ellipse/Point display: flat alpha inside d^2 <= 4            */
#define ORC_POINT_RADIUS 1.5f     /* Point display: disc radius in pixels                         */
#define ORC_SH_C0 0.28209479177387814f
#define ORC_SH_C1 0.4886025119029199f

typedef struct orc_frame {
    float view[16], proj[16]; /* column-major (glam) */
    float size[2];            /* viewport in pixels  */
    float gaussian_size;      /* gs::update_gaussian_transform size           */
    uint32_t display_mode, sh_deg, no_sh0;
    b200gs_edit_pod selection_edit;
    float highlight[4];
    float background[4];
    b200gs_query_pod query; /* selection query tested during preprocess (rect / brush / texture) */
    const uint8_t* query_tex; /* kind TEXTURE: u8 per pixel, row-major, query_tex_w x query_tex_h (may be NULL) */
    uint32_t query_tex_w, query_tex_h;
} orc_frame;

typedef struct orc_model {
    uint32_t sh, cov3d;  /* layout */
    const void* packed;  /* N records */
    uint64_t n;
    float pos[3], quat[4], scale[3];
    const uint32_t* mask;      /* may be NULL = all shown   */
    const uint32_t* selection; /* may be NULL = none        */
    const b200gs_edit_pod* edits; /* may be NULL = no edits */
} orc_model;

/* The oracle's OWN projected splat: fp32 throughout, nothing rounded to f16.  (The product stores a 32-byte
 * record with f16 colour and opacity, `b200gs_splat`; the *_f32 entry points below never touch that format, so
 * the image tolerance measured against them includes the product's quantisation.) */
typedef struct orc_splat_f32 {
    float mx, my, radius; /* pixel centre, extent-square half-size */
    float opacity, r, g, b;
    float ca, cb, cc;     /* conic */
    uint32_t flags;       /* bit 0: selected */
} orc_splat_f32;

/* host-side conversions */
uint32_t orc_record_bytes(uint32_t sh, uint32_t cov3d);
uint16_t orc_f32_to_f16(float f);
float orc_f16_to_f32(uint16_t h);
void orc_synth_scene(uint64_t seed, uint64_t start, uint64_t count, b200gs_ply_gaussian* out);
void orc_gaussian_from_ply(const b200gs_ply_gaussian* in, uint64_t count, b200gs_gaussian* out);
void orc_pack(uint32_t sh, uint32_t cov3d, const b200gs_gaussian* in, uint64_t count, void* out);
void orc_look_at_rh(const float eye[3], const float target[3], const float up[3], float out[16]);
void orc_perspective_rh(float vfov, float aspect, float z_near, float z_far, float out[16]);
void orc_quat_from_euler_zyx_deg(const float rot_deg[3], float quat[4]);

/* a1: preprocess.  Outputs in ascending Gaussian index order; returns V. */
uint64_t orc_preprocess(const orc_frame* f, const orc_model* m, uint32_t* indices, uint32_t* keys, b200gs_splat* splats);
uint64_t orc_preprocess_f32(const orc_frame* f, const orc_model* m, uint32_t* indices, uint32_t* keys, orc_splat_f32* splats);
void orc_sort_f32(uint64_t v, uint32_t* keys, uint32_t* indices, orc_splat_f32* splats);
void orc_composite_b2f_f32(const orc_frame* f, const orc_splat_f32* splats, uint64_t n_total, float* rgba_f, uint8_t* rgba8);
uint64_t orc_composite_f2b_f32(const orc_frame* f, const orc_splat_f32* splats, uint64_t n_total, float* rgba_f, uint8_t* rgba8);
uint64_t orc_render_frame_f32(const orc_frame* f, const orc_model* far_to_near, uint32_t n_models, int front_to_back,
                              uint8_t* rgba8, double stage_seconds[3]);
/* a2: stable ascending sort of (key, index) with the splats carried along */
void orc_sort(uint64_t v, uint32_t* keys, uint32_t* indices, b200gs_splat* splats);
void orc_sort_pairs(uint64_t n, uint32_t* keys, uint32_t* values, uint32_t bits);
/* a3: compositing.  `splats` = concatenation of the models' depth-sorted (near->far) lists,
 * NEAREST model first; n_total entries.  rgba_f (4 floats / pixel) and rgba8 may be NULL.
 * b2f = reference-style back-to-front "over" blending; f2b = front-to-back with early
 * termination (returns the number of splat evaluations performed before termination). */
void orc_composite_b2f(const orc_frame* f, const b200gs_splat* splats, uint64_t n_total, float* rgba_f, uint8_t* rgba8);
uint64_t orc_composite_f2b(const orc_frame* f, const b200gs_splat* splats, uint64_t n_total, float* rgba_f, uint8_t* rgba8);
/* a4: model order, farthest world_center first (scene.rs:533-558) */
void orc_order_models(const orc_frame* f, const orc_model* models, const float* centers, uint32_t n, uint32_t* order);
/* whole frame: models in far-to-near order; returns total visible */
uint64_t orc_render_frame(const orc_frame* f, const orc_model* far_to_near, uint32_t n_models, int front_to_back,
                          uint8_t* rgba8, double stage_seconds[3]);
/* N2: mask evaluation */
void orc_eval_mask(const orc_model* m, const b200gs_mask_op* postfix, uint32_t n_ops, const b200gs_mask_shape* shapes,
                   uint32_t n_shapes, uint32_t* words);
/* N2: selection query (rect / brush, Set/Add/Remove) of f->query applied to m->selection -> words_out */
void orc_query_selection(const orc_frame* f, const orc_model* m, uint32_t* words_out);
/* N2: one rect / brush stroke painted into a query texture (query_toolset.render, scene.rs:787-791) */
void orc_query_texture_paint(uint8_t* tex, uint32_t w, uint32_t h, const b200gs_query_pod* stroke);
/* N2: postprocess (scene.rs:604-610): the viewer's selection edit committed into the edit pods of the selected Gaussians */
void orc_postprocess(uint64_t n, const uint32_t* selection, b200gs_edit_pod* edits, const b200gs_edit_pod* selection_edit);
void orc_apply_edit(const b200gs_edit_pod* e, float rgb[3], float* opacity);
/* N4: export with edits + mask (either may be NULL); returns the number of vertices written to out (<= count) */
uint64_t orc_export_edited(const b200gs_gaussian* in, uint64_t count, const b200gs_edit_pod* edits, const uint32_t* mask,
                           b200gs_ply_gaussian* out);
int orc_num_threads(void);
void orc_set_num_threads(int n);

#ifdef __cplusplus
}
#endif
#endif
