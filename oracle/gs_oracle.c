/*
 * gs_oracle.c — CPU ORACLE (test infrastructure, NOT product code).  See gs_oracle.h.
 * PARITY UNPINNED (no reference tests / goldens exist for this path).
 *
 * Build: gcc -O2 -ffp-contract=off -fopenmp -shared -fPIC gs_oracle.c -o _build/libgs_oracle.so -lm
 * Every float expression below is meant to be evaluated exactly as written, one rounding
 * per operation, left to right, and fused multiply-adds ONLY where fmaf() is spelled out (libm's
 * fmaf is correctly rounded with or without hardware FMA).  Do not "simplify" the arithmetic: the
 * CUDA kernels are compared bit-for-bit against it for culling, depth keys and pixel bounds.
 */
#include "gs_oracle.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>
#ifdef _OPENMP
#include <omp.h>
#endif

int orc_num_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

/* the thread count is set explicitly by callers that time the oracle (bench.py), because launchers such as
 * torch.distributed.run export OMP_NUM_THREADS=1 to their children */
void orc_set_num_threads(int n) {
#ifdef _OPENMP
    if (n >= 1) omp_set_num_threads(n);
#else
    (void)n;
#endif
}

static double now_s(void) {
    struct timespec ts;
    clock_gettime(CLOCK_MONOTONIC, &ts);
    return (double)ts.tv_sec + 1e-9 * (double)ts.tv_nsec;
}

/* ------------------------------------------------------------------ f16 */
/* IEEE binary16 <-> binary32, round-to-nearest-even (crate dep `half 2.4.1`,
 * Cargo.lock:3731-3747: f16::from_f32 is RN-even). */
uint16_t orc_f32_to_f16(float f) {
    uint32_t x;
    memcpy(&x, &f, 4);
    uint32_t sign = (x >> 16) & 0x8000u;
    uint32_t e = (x >> 23) & 0xffu;
    uint32_t m = x & 0x7fffffu;
    if (e == 0xff) return (uint16_t)(sign | 0x7c00u | (m ? 0x200u | (m >> 13) : 0));
    int32_t ee = (int32_t)e - 127 + 15;
    if (ee >= 31) return (uint16_t)(sign | 0x7c00u); /* overflow -> inf */
    if (ee <= 0) {
        if (ee < -10) return (uint16_t)sign; /* underflow -> 0 */
        m |= 0x800000u;
        uint32_t shift = (uint32_t)(14 - ee);
        uint32_t hm = m >> shift;
        uint32_t rem = m & ((1u << shift) - 1);
        uint32_t half = 1u << (shift - 1);
        if (rem > half || (rem == half && (hm & 1))) hm++;
        return (uint16_t)(sign | hm);
    }
    uint32_t h = ((uint32_t)ee << 10) | (m >> 13);
    uint32_t rem = m & 0x1fffu;
    if (rem > 0x1000u || (rem == 0x1000u && (h & 1))) h++;
    return (uint16_t)(sign | h);
}

float orc_f16_to_f32(uint16_t h) {
    uint32_t sign = ((uint32_t)h & 0x8000u) << 16;
    uint32_t e = (h >> 10) & 0x1fu;
    uint32_t m = h & 0x3ffu;
    uint32_t x;
    if (e == 0) {
        if (m == 0) x = sign;
        else {
            int s = 0;
            while (!(m & 0x400u)) { m <<= 1; s++; }
            m &= 0x3ffu;
            x = sign | ((uint32_t)(127 - 15 - s + 1) << 23) | (m << 13);
        }
    } else if (e == 31) x = sign | 0x7f800000u | (m << 13);
    else x = sign | ((e - 15 + 127) << 23) | (m << 13);
    float f;
    memcpy(&f, &x, 4);
    return f;
}

/* ------------------------------------------------------------ layouts (a5) */
/* Record = pos f32x3 | color u8x4 | SH | Cov3d (scene.rs:907-978; app.rs:250-257). */
static uint32_t sh_bytes(uint32_t sh) { return sh == 0 ? 180u : sh == 1 ? 92u : sh == 2 ? 48u : sh == 3 ? 0u : ~0u; }
static uint32_t cov_bytes(uint32_t c) { return c == 0 ? 24u : c == 1 ? 12u : ~0u; }
uint32_t orc_record_bytes(uint32_t sh, uint32_t cov3d) {
    if (sh > 3 || cov3d > 1) return 0;
    return 16u + sh_bytes(sh) + cov_bytes(cov3d);
}

/* glam Mat3::from_quat */
static void quat_to_mat3(const float q[4], float R[3][3]) {
    float x = q[0], y = q[1], z = q[2], w = q[3];
    float x2 = x + x, y2 = y + y, z2 = z + z;
    float xx = x * x2, xy = x * y2, xz = x * z2;
    float yy = y * y2, yz = y * z2, zz = z * z2;
    float wx = w * x2, wy = w * y2, wz = w * z2;
    R[0][0] = 1.0f - (yy + zz); R[0][1] = xy - wz;          R[0][2] = xz + wy;
    R[1][0] = xy + wz;          R[1][1] = 1.0f - (xx + zz); R[1][2] = yz - wx;
    R[2][0] = xz - wy;          R[2][1] = yz + wx;          R[2][2] = 1.0f - (xx + yy);
}

/* -------------------------------------------------- synthetic scene (§8d) */
static uint64_t mix64(uint64_t z) {
    z += 0x9E3779B97F4A7C15ULL;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
    return z ^ (z >> 31);
}
static double u01(uint64_t seed, uint64_t i, uint64_t k) {
    uint64_t h = mix64(mix64(seed ^ (i * 0xD1342543DE82EF95ULL)) + k);
    return ((double)(h >> 11) + 0.5) * (1.0 / 9007199254740992.0);
}
static void npair(uint64_t seed, uint64_t i, uint64_t k, double* a, double* b) {
    double u1 = u01(seed, i, 2 * k), u2 = u01(seed, i, 2 * k + 1);
    double r = sqrt(-2.0 * log(u1));
    double t = 6.283185307179586476925286766559 * u2;
    *a = r * cos(t);
    *b = r * sin(t);
}
#define SYNTH_CLUSTERS 64
#define SYNTH_CLUSTER_SALT 0xC1A57E2500000000ULL
void orc_synth_scene(uint64_t seed, uint64_t start, uint64_t count, b200gs_ply_gaussian* out) {
    double cc[SYNTH_CLUSTERS][4];
    for (int c = 0; c < SYNTH_CLUSTERS; c++) {
        uint64_t s2 = seed ^ SYNTH_CLUSTER_SALT;
        cc[c][0] = (2.0 * u01(s2, (uint64_t)c, 0) - 1.0) * 4.0;
        cc[c][1] = (2.0 * u01(s2, (uint64_t)c, 1) - 1.0) * 1.5;
        cc[c][2] = (2.0 * u01(s2, (uint64_t)c, 2) - 1.0) * 4.0;
        cc[c][3] = 0.05 + 0.35 * u01(s2, (uint64_t)c, 3);
    }
    const double ln_scale = log(0.006);
#pragma omp parallel for schedule(static)
    for (int64_t jj = 0; jj < (int64_t)count; jj++) {
        uint64_t i = start + (uint64_t)jj;
        b200gs_ply_gaussian* g = &out[jj];
        double n[64];
        for (int k = 0; k < 32; k++) npair(seed, i, (uint64_t)k, &n[2 * k], &n[2 * k + 1]);
        if (u01(seed, i, 100) < 0.1) {
            g->pos[0] = (float)((2.0 * u01(seed, i, 102) - 1.0) * 4.0);
            g->pos[1] = (float)((2.0 * u01(seed, i, 103) - 1.0) * 1.5);
            g->pos[2] = (float)((2.0 * u01(seed, i, 104) - 1.0) * 4.0);
        } else {
            int c = (int)(u01(seed, i, 101) * (double)SYNTH_CLUSTERS);
            if (c >= SYNTH_CLUSTERS) c = SYNTH_CLUSTERS - 1;
            for (int a = 0; a < 3; a++) g->pos[a] = (float)(cc[c][a] + cc[c][3] * n[a]);
        }
        g->normal[0] = g->normal[1] = g->normal[2] = 0.0f;
        for (int a = 0; a < 3; a++) g->scale[a] = (float)(ln_scale + 0.6 * n[4 + a]);
        for (int a = 0; a < 4; a++) g->rot[a] = (float)n[8 + a];
        g->opacity = (float)(0.5 + 2.0 * n[12]);
        for (int a = 0; a < 3; a++) g->f_dc[a] = (float)(0.8 * n[14 + a]);
        for (int c = 0; c < 3; c++)
            for (int k = 0; k < 15; k++) {
                double band = k < 3 ? 1.0 : (k < 8 ? 2.0 : 3.0);
                g->f_rest[c * 15 + k] = (float)((0.15 / band) * n[18 + c * 15 + k]);
            }
    }
}

/* ---------------------------------------------- Gaussian::from(PlyGaussianPod)
 * [CANON] SURVEY.md §8c.1 (call site src/app.rs:1066). */
static uint8_t unorm8(float x) {
    if (!(x > 0.0f)) x = 0.0f;
    if (x > 1.0f) x = 1.0f;
    return (uint8_t)(x * 255.0f + 0.5f);
}
void orc_gaussian_from_ply(const b200gs_ply_gaussian* in, uint64_t count, b200gs_gaussian* out) {
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < (int64_t)count; i++) {
        const b200gs_ply_gaussian* p = &in[i];
        b200gs_gaussian* g = &out[i];
        float w = p->rot[0], x = p->rot[1], y = p->rot[2], z = p->rot[3];
        float len = sqrtf(x * x + y * y + z * z + w * w);
        if (len > 0.0f) { x = x / len; y = y / len; z = z / len; w = w / len; }
        else { x = y = z = 0.0f; w = 1.0f; }
        g->rot[0] = x; g->rot[1] = y; g->rot[2] = z; g->rot[3] = w;
        for (int a = 0; a < 3; a++) g->pos[a] = p->pos[a];
        for (int a = 0; a < 3; a++) g->scale[a] = expf(p->scale[a]);
        for (int c = 0; c < 3; c++) g->color[c] = unorm8(0.5f + ORC_SH_C0 * p->f_dc[c]);
        g->color[3] = unorm8(1.0f / (1.0f + expf(-p->opacity)));
        for (int k = 0; k < 15; k++)
            for (int c = 0; c < 3; c++) g->sh[3 * k + c] = p->f_rest[c * 15 + k];
    }
}

/* ------------------------------------------------------------------ pack */
/* Host half of GaussiansBuffer::update_range (scene.rs:2069-2085); Σ = R S S^T R^T
 * stored (xx,xy,xz,yy,yz,zz) [CANON §8c.2]; Norm8: q = RN(clamp((x+1)/2)·255). */
void orc_pack(uint32_t sh, uint32_t cov3d, const b200gs_gaussian* in, uint64_t count, void* out) {
    uint32_t rb = orc_record_bytes(sh, cov3d);
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < (int64_t)count; i++) {
        const b200gs_gaussian* g = &in[i];
        uint8_t* r = (uint8_t*)out + (size_t)i * rb;
        memcpy(r, g->pos, 12);
        memcpy(r + 12, g->color, 4);
        uint8_t* p = r + 16;
        if (sh == 0) { memcpy(p, g->sh, 180); p += 180; }
        else if (sh == 1) {
            uint16_t h[46];
            for (int k = 0; k < 45; k++) h[k] = orc_f32_to_f16(g->sh[k]);
            h[45] = 0;
            memcpy(p, h, 92); p += 92;
        } else if (sh == 2) {
            uint8_t q[48];
            for (int k = 0; k < 45; k++) q[k] = unorm8((g->sh[k] + 1.0f) * 0.5f);
            q[45] = q[46] = q[47] = 0;
            memcpy(p, q, 48); p += 48;
        }
        float R[3][3], M[3][3];
        quat_to_mat3(g->rot, R);
        for (int a = 0; a < 3; a++)
            for (int b = 0; b < 3; b++) M[a][b] = R[a][b] * g->scale[b];
        float c6[6];
        c6[0] = M[0][0] * M[0][0] + M[0][1] * M[0][1] + M[0][2] * M[0][2];
        c6[1] = M[0][0] * M[1][0] + M[0][1] * M[1][1] + M[0][2] * M[1][2];
        c6[2] = M[0][0] * M[2][0] + M[0][1] * M[2][1] + M[0][2] * M[2][2];
        c6[3] = M[1][0] * M[1][0] + M[1][1] * M[1][1] + M[1][2] * M[1][2];
        c6[4] = M[1][0] * M[2][0] + M[1][1] * M[2][1] + M[1][2] * M[2][2];
        c6[5] = M[2][0] * M[2][0] + M[2][1] * M[2][1] + M[2][2] * M[2][2];
        if (cov3d == 0) memcpy(p, c6, 24);
        else {
            uint16_t h[6];
            for (int k = 0; k < 6; k++) h[k] = orc_f32_to_f16(c6[k]);
            memcpy(p, h, 12);
        }
    }
}

/* -------------------------------------------------------------- camera (a6)
 * glam 0.29 Mat4::look_at_rh / perspective_rh (depth 0..1) as called at src/app.rs:1236-1244. */
void orc_look_at_rh(const float eye[3], const float target[3], const float up[3], float out[16]) {
    float d[3] = {target[0] - eye[0], target[1] - eye[1], target[2] - eye[2]};
    float dl = sqrtf(d[0] * d[0] + d[1] * d[1] + d[2] * d[2]);
    float f[3] = {d[0] / dl, d[1] / dl, d[2] / dl};
    float s[3] = {f[1] * up[2] - f[2] * up[1], f[2] * up[0] - f[0] * up[2], f[0] * up[1] - f[1] * up[0]};
    float sl = sqrtf(s[0] * s[0] + s[1] * s[1] + s[2] * s[2]);
    s[0] = s[0] / sl; s[1] = s[1] / sl; s[2] = s[2] / sl;
    float u[3] = {s[1] * f[2] - s[2] * f[1], s[2] * f[0] - s[0] * f[2], s[0] * f[1] - s[1] * f[0]};
    out[0] = s[0]; out[1] = u[0]; out[2] = -f[0]; out[3] = 0.0f;
    out[4] = s[1]; out[5] = u[1]; out[6] = -f[1]; out[7] = 0.0f;
    out[8] = s[2]; out[9] = u[2]; out[10] = -f[2]; out[11] = 0.0f;
    out[12] = -(eye[0] * s[0] + eye[1] * s[1] + eye[2] * s[2]);
    out[13] = -(eye[0] * u[0] + eye[1] * u[1] + eye[2] * u[2]);
    out[14] = eye[0] * f[0] + eye[1] * f[1] + eye[2] * f[2];
    out[15] = 1.0f;
}
void orc_perspective_rh(float vfov, float aspect, float z_near, float z_far, float out[16]) {
    float sf = sinf(0.5f * vfov), cf = cosf(0.5f * vfov);
    float h = cf / sf;
    float w = h / aspect;
    float r = z_far / (z_near - z_far);
    memset(out, 0, 64);
    out[0] = w;
    out[5] = h;
    out[10] = r;
    out[11] = -1.0f;
    out[14] = r * z_near;
}
/* Quat::from_euler(EulerRot::ZYX, rz, ry, rx) with degrees->radians (src/app.rs:1123-1130):
 * q = qz(rz) * qy(ry) * qx(rx). */
void orc_quat_from_euler_zyx_deg(const float rot_deg[3], float q[4]) {
    const float d2r = 0.017453292519943295f;
    float hx = rot_deg[0] * d2r * 0.5f, hy = rot_deg[1] * d2r * 0.5f, hz = rot_deg[2] * d2r * 0.5f;
    float sx = sinf(hx), cx = cosf(hx), sy = sinf(hy), cy = cosf(hy), sz = sinf(hz), cz = cosf(hz);
    q[0] = cz * cy * sx - sz * sy * cx;
    q[1] = cz * sy * cx + sz * cy * sx;
    q[2] = sz * cy * cx - cz * sy * sx;
    q[3] = cz * cy * cx + sz * sy * sx;
}

/* ---------------------------------------------------------------- edits (a3)
 * Inputs pinned by src/app.rs:1533-1564 and src/tab/selection.rs:172-204; the ORDER of the
 * operations inside the crate is unknown (§8c.9) — fixed here as
 * colour (HSV or override) -> contrast -> exposure -> gamma -> alpha. */
static float clamp01(float x) { return x < 0.0f ? 0.0f : (x > 1.0f ? 1.0f : x); }
static void rgb_to_hsv(const float c[3], float hsv[3]) {
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
static void hsv_to_rgb(const float hsv[3], float c[3]) {
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
void orc_apply_edit(const b200gs_edit_pod* e, float rgb[3], float* opacity) {
    if (!(e->flag & B200GS_EDIT_ENABLED)) return;
    if (e->flag & B200GS_EDIT_OVERRIDE_COLOR) {
        rgb[0] = e->color[0]; rgb[1] = e->color[1]; rgb[2] = e->color[2];
    } else {
        float hsv[3];
        rgb_to_hsv(rgb, hsv);
        float h = hsv[0] + e->color[0];
        hsv[0] = h - floorf(h);
        hsv[1] = clamp01(hsv[1] * e->color[1]);
        hsv[2] = hsv[2] * e->color[2];
        hsv_to_rgb(hsv, rgb);
    }
    float ex = exp2f(e->exposure);
    for (int c = 0; c < 3; c++) {
        float v = (rgb[c] - 0.5f) * (1.0f + e->contrast) + 0.5f;
        v = v * ex;
        v = powf(fmaxf(v, 0.0f), e->gamma);
        rgb[c] = v;
    }
    *opacity = clamp01(*opacity * e->alpha);
}

/* --------------------------------------------------------- preprocess (a1)
 * Replaces viewer.preprocessor.preprocess (src/tab/scene.rs:856-863, bindings :1835-1852)
 * plus the per-splat vertex work of renderer.render_with_pass (scene.rs:2306-2313) that the
 * north_star moves into this stage (cov3d->cov2d, conic, extent, SH colour, edits). */
typedef struct pre_ctx {
    float V[4][4], P[4][4]; /* row-major [r][c] */
    float R[3][3], t[3], s[3], M[3][3];
    float cam[3];
    float W, H, fx, fy, limx, limy, sz2;
} pre_ctx;

static void pre_setup(const orc_frame* f, const orc_model* m, pre_ctx* c) {
    for (int r = 0; r < 4; r++)
        for (int k = 0; k < 4; k++) { c->V[r][k] = f->view[k * 4 + r]; c->P[r][k] = f->proj[k * 4 + r]; }
    quat_to_mat3(m->quat, c->R);
    for (int a = 0; a < 3; a++) { c->t[a] = m->pos[a]; c->s[a] = m->scale[a]; }
    for (int a = 0; a < 3; a++)
        for (int b = 0; b < 3; b++) c->M[a][b] = c->R[a][b] * c->s[b];
    /* camera position = -R_v^T t_v for a rigid view matrix */
    for (int a = 0; a < 3; a++)
        c->cam[a] = -(c->V[0][a] * c->V[0][3] + c->V[1][a] * c->V[1][3] + c->V[2][a] * c->V[2][3]);
    c->W = f->size[0]; c->H = f->size[1];
    c->fx = c->P[0][0] * c->W * 0.5f;
    c->fy = c->P[1][1] * c->H * 0.5f;
    c->limx = ORC_CLAMP_XY / c->P[0][0];
    c->limy = ORC_CLAMP_XY / c->P[1][1];
    c->sz2 = f->gaussian_size * f->gaussian_size;
}

static void decode_record(const orc_model* m, const uint8_t* r, float pos[3], uint8_t col[4], float sh[45], float cov[6]) {
    memcpy(pos, r, 12);
    memcpy(col, r + 12, 4);
    const uint8_t* p = r + 16;
    if (m->sh == 0) { memcpy(sh, p, 180); p += 180; }
    else if (m->sh == 1) {
        uint16_t h[46];
        memcpy(h, p, 92);
        for (int k = 0; k < 45; k++) sh[k] = orc_f16_to_f32(h[k]);
        p += 92;
    } else if (m->sh == 2) {
        for (int k = 0; k < 45; k++) sh[k] = (float)p[k] * (2.0f / 255.0f) - 1.0f;
        p += 48;
    } else {
        for (int k = 0; k < 45; k++) sh[k] = 0.0f;
    }
    if (m->cov3d == 0) memcpy(cov, p, 24);
    else {
        uint16_t h[6];
        memcpy(h, p, 12);
        for (int k = 0; k < 6; k++) cov[k] = orc_f16_to_f32(h[k]);
    }
}

/* SH basis for bands 1..3 in the Inria sign convention (computeColorFromSH) [CANON §8c.3] */
static void sh_basis(float x, float y, float z, uint32_t deg, float b[15]) {
    for (int k = 0; k < 15; k++) b[k] = 0.0f;
    if (deg < 1) return;
    b[0] = -ORC_SH_C1 * y;
    b[1] = ORC_SH_C1 * z;
    b[2] = -ORC_SH_C1 * x;
    if (deg < 2) return;
    float xx = x * x, yy = y * y, zz = z * z, xy = x * y, yz = y * z, xz = x * z;
    b[3] = 1.0925484305920792f * xy;
    b[4] = -1.0925484305920792f * yz;
    b[5] = 0.31539156525252005f * (2.0f * zz - xx - yy);
    b[6] = -1.0925484305920792f * xz;
    b[7] = 0.5462742152960396f * (xx - yy);
    if (deg < 3) return;
    b[8] = -0.5900435899266435f * y * (3.0f * xx - yy);
    b[9] = 2.890611442640554f * xy * z;
    b[10] = -0.4570457994644658f * y * (4.0f * zz - xx - yy);
    b[11] = 0.3731763325901154f * z * (2.0f * zz - 3.0f * xx - 3.0f * yy);
    b[12] = -0.4570457994644658f * x * (4.0f * zz - xx - yy);
    b[13] = 1.445305721320277f * z * (xx - yy);
    b[14] = -0.5900435899266435f * x * (xx - 3.0f * yy);
}

/* Selection query in immediate mode (gs::QueryToolset rect / brush, src/tab/scene.rs:758-791,
 * 1224-1263; ops src/app.rs:1453): splat centre in viewport pixels (top-left origin) inside the
 * rectangle, or within `radius` of the brush segment. */
static int query_shape_hit(const b200gs_query_pod* q, float sx, float sy) {
    if (q->kind == B200GS_QUERY_RECT) return sx >= q->p0[0] && sx <= q->p1[0] && sy >= q->p0[1] && sy <= q->p1[1];
    float vx = q->p1[0] - q->p0[0], vy = q->p1[1] - q->p0[1];
    float wx = sx - q->p0[0], wy = sy - q->p0[1];
    float vv = vx * vx + vy * vy;
    float t = vv > 0.0f ? (wx * vx + wy * vy) / vv : 0.0f;
    t = fminf(1.0f, fmaxf(0.0f, t));
    float dx = wx - t * vx, dy = wy - t * vy;
    return dx * dx + dy * dy <= q->radius * q->radius;
}

/* non-immediate mode (set_use_texture(true), scene.rs:767-791): the texel under the splat centre */
static int query_hit(const orc_frame* f, float sx, float sy) {
    if (f->query.kind == B200GS_QUERY_TEXTURE) {
        float fx = floorf(sx), fy = floorf(sy);
        if (!f->query_tex || !(fx >= 0.0f && fy >= 0.0f && fx < (float)f->query_tex_w && fy < (float)f->query_tex_h)) return 0;
        return f->query_tex[(size_t)(uint32_t)fy * f->query_tex_w + (uint32_t)fx] != 0;
    }
    return query_shape_hit(&f->query, sx, sy);
}
void orc_query_texture_paint(uint8_t* tex, uint32_t w, uint32_t h, const b200gs_query_pod* stroke) {
    for (uint32_t y = 0; y < h; y++)
        for (uint32_t x = 0; x < w; x++)
            if (query_shape_hit(stroke, (float)x + 0.5f, (float)y + 0.5f)) tex[(size_t)y * w + x] = 255;
}
void orc_postprocess(uint64_t n, const uint32_t* selection, b200gs_edit_pod* edits, const b200gs_edit_pod* selection_edit) {
    if (!(selection_edit->flag & B200GS_EDIT_ENABLED)) return;
    for (uint64_t i = 0; i < n; i++)
        if ((selection[i >> 5] >> (i & 31)) & 1u) edits[i] = *selection_edit;
}

/* one Gaussian; returns 1 if visible */
static int pre_one(const orc_frame* f, const orc_model* m, const pre_ctx* c, uint64_t i, uint32_t* key, orc_splat_f32* out,
                   int* selected_out) {
    if (selected_out) *selected_out = m->selection ? (int)((m->selection[i >> 5] >> (i & 31)) & 1u) : 0;
    if (m->mask && !((m->mask[i >> 5] >> (i & 31)) & 1u)) return 0;
    const b200gs_edit_pod* ed = m->edits ? &m->edits[i] : NULL;
    int selected = m->selection ? (int)((m->selection[i >> 5] >> (i & 31)) & 1u) : 0;
    if (ed && (ed->flag & B200GS_EDIT_ENABLED) && (ed->flag & B200GS_EDIT_HIDDEN)) return 0;
    if (selected && (f->selection_edit.flag & B200GS_EDIT_ENABLED) && (f->selection_edit.flag & B200GS_EDIT_HIDDEN))
        return 0;
    uint32_t rb = orc_record_bytes(m->sh, m->cov3d);
    float p[3], sh[45], cv[6];
    uint8_t col[4];
    decode_record(m, (const uint8_t*)m->packed + (size_t)i * rb, p, col, sh, cv);

    /* world = q*(s⊙p)+t  (src/app.rs:1044-1046).  Dot products are written as explicit fused chains (fmaf): the
     * CUDA kernel issues the same chain as FFMA, and a fused multiply-add is what a GPU shader compiler emits for
     * `a*b + c` in the reference's WGSL as well. */
    float ps[3] = {c->s[0] * p[0], c->s[1] * p[1], c->s[2] * p[2]};
    float pw[3], pv[3], pc[4];
    for (int r = 0; r < 3; r++) pw[r] = fmaf(c->R[r][0], ps[0], fmaf(c->R[r][1], ps[1], fmaf(c->R[r][2], ps[2], c->t[r])));
    for (int r = 0; r < 3; r++) pv[r] = fmaf(c->V[r][0], pw[0], fmaf(c->V[r][1], pw[1], fmaf(c->V[r][2], pw[2], c->V[r][3])));
    for (int r = 0; r < 4; r++) pc[r] = fmaf(c->P[r][0], pv[0], fmaf(c->P[r][1], pv[1], fmaf(c->P[r][2], pv[2], c->P[r][3])));
    if (!(pc[3] > 0.0f)) return 0;
    float iw = 1.0f / pc[3]; /* one IEEE reciprocal, then multiplies */
    float nx = pc[0] * iw, ny = pc[1] * iw, nz = pc[2] * iw;
    /* frustum cull [RECALLED §8c.5] */
    if (!(nz > 0.0f && nz < 1.0f && fabsf(nx) <= ORC_CULL_XY && fabsf(ny) <= ORC_CULL_XY)) return 0;
    /* depth key [§8c.6]: bits(ndc.z), ascending = near -> far */
    memcpy(key, &nz, 4);
    /* selection query: the new selection state is what this frame shows */
    if (f->query.kind >= B200GS_QUERY_RECT) {
        float sx = fmaf(nx + 1.0f, c->W, -1.0f) * 0.5f + 0.5f, sy = fmaf(1.0f - ny, c->H, -1.0f) * 0.5f + 0.5f;
        int hit = query_hit(f, sx, sy);
        if (f->query.op == B200GS_SELECT_SET) selected = hit;
        else if (f->query.op == B200GS_SELECT_ADD) selected = selected | hit;
        else selected = selected & !hit;
        if (selected_out) *selected_out = selected;
    }

    /* Σ' = (R_m S_m) Σ (R_m S_m)^T · size² */
    float S[3][3] = {{cv[0], cv[1], cv[2]}, {cv[1], cv[3], cv[4]}, {cv[2], cv[4], cv[5]}};
    float B[3][3], Sw[3][3];
    for (int r = 0; r < 3; r++)
        for (int k = 0; k < 3; k++) B[r][k] = fmaf(c->M[r][0], S[0][k], fmaf(c->M[r][1], S[1][k], c->M[r][2] * S[2][k]));
    for (int r = 0; r < 3; r++)
        for (int k = 0; k < 3; k++)
            Sw[r][k] = fmaf(B[r][0], c->M[k][0], fmaf(B[r][1], c->M[k][1], B[r][2] * c->M[k][2])) * c->sz2;
    /* make it exactly symmetric the same way the kernel does: use the upper triangle */
    Sw[1][0] = Sw[0][1]; Sw[2][0] = Sw[0][2]; Sw[2][1] = Sw[1][2];

    /* Jacobian of the pixel mapping [CANON §8c.7]; view space is RH, looking down -z */
    float tz = -pv[2];
    float itz = 1.0f / tz;
    float txz = pv[0] * itz, tyz = pv[1] * itz;
    txz = fminf(c->limx, fmaxf(-c->limx, txz));
    tyz = fminf(c->limy, fmaxf(-c->limy, tyz));
    float xc = txz * tz, yc = tyz * tz;
    float itz2 = itz * itz;
    float J00 = c->fx * itz, J02 = (c->fx * xc) * itz2;
    float J11 = -(c->fy * itz), J12 = -((c->fy * yc) * itz2);
    float T0[3], T1[3];
    for (int k = 0; k < 3; k++) {
        T0[k] = fmaf(J00, c->V[0][k], J02 * c->V[2][k]);
        T1[k] = fmaf(J11, c->V[1][k], J12 * c->V[2][k]);
    }
    float U0[3], U1[3];
    for (int k = 0; k < 3; k++) {
        U0[k] = fmaf(T0[0], Sw[0][k], fmaf(T0[1], Sw[1][k], T0[2] * Sw[2][k]));
        U1[k] = fmaf(T1[0], Sw[0][k], fmaf(T1[1], Sw[1][k], T1[2] * Sw[2][k]));
    }
    float a = fmaf(U0[0], T0[0], fmaf(U0[1], T0[1], U0[2] * T0[2]));
    float b = fmaf(U0[0], T1[0], fmaf(U0[1], T1[1], U0[2] * T1[2]));
    float d = fmaf(U1[0], T1[0], fmaf(U1[1], T1[1], U1[2] * T1[2]));
    a = a + ORC_LOWPASS;
    d = d + ORC_LOWPASS;
    float det = fmaf(a, d, -(b * b));
    float ca = 0.0f, cb = 0.0f, cc = 0.0f, radf = 0.0f;
    if (det > 0.0f) {
        float di = 1.0f / det;
        ca = d * di; cb = -b * di; cc = a * di;
        float mid = 0.5f * (a + d);
        float disc = fmaf(mid, mid, -det);
        if (disc < ORC_MIN_DISC) disc = ORC_MIN_DISC;
        float lam = mid + sqrtf(disc);
        radf = ceilf(ORC_EXTENT_SIGMA * sqrtf(lam));
    }
    if (f->display_mode == B200GS_DISPLAY_POINT) {
        ca = ORC_FLAT_D2 / (ORC_POINT_RADIUS * ORC_POINT_RADIUS); cb = 0.0f; cc = ca;
        radf = ceilf(ORC_POINT_RADIUS);
    }
    if (!(radf <= 65535.0f)) radf = 65535.0f;

    /* colour: base u8 (SH0 baked, toggled by no_sh0 — src/tab/transform.rs:142-145) + bands
     * 1..sh_deg (transform.rs:135-139), view direction in WORLD space (§8c.3) */
    float dx = pw[0] - c->cam[0], dy = pw[1] - c->cam[1], dz = pw[2] - c->cam[2];
    float dl = sqrtf(dx * dx + dy * dy + dz * dz);
    dx = dx / dl; dy = dy / dl; dz = dz / dl;
    float basis[15];
    sh_basis(dx, dy, dz, f->sh_deg, basis);
    float rgb[3];
    uint32_t ncoef = f->sh_deg >= 3 ? 15u : (f->sh_deg == 2 ? 8u : (f->sh_deg == 1 ? 3u : 0u));
    for (int ch = 0; ch < 3; ch++) {
        float v = f->no_sh0 ? 0.0f : (float)col[ch] / 255.0f;
        for (uint32_t k = 0; k < ncoef; k++) v = v + basis[k] * sh[3 * k + ch];
        rgb[ch] = clamp01(v);
    }
    float op = (float)col[3] / 255.0f;
    if (ed) orc_apply_edit(ed, rgb, &op);
    if (selected) {
        orc_apply_edit(&f->selection_edit, rgb, &op);
        float ha = f->highlight[3];
        for (int ch = 0; ch < 3; ch++) rgb[ch] = rgb[ch] + (f->highlight[ch] - rgb[ch]) * ha;
    }
    for (int ch = 0; ch < 3; ch++) rgb[ch] = clamp01(rgb[ch]);

    out->mx = fmaf(nx + 1.0f, c->W, -1.0f) * 0.5f;
    out->my = fmaf(1.0f - ny, c->H, -1.0f) * 0.5f;
    out->radius = radf;
    out->opacity = op;
    out->r = rgb[0]; out->g = rgb[1]; out->b = rgb[2];
    out->ca = ca; out->cb = cb; out->cc = cc;
    out->flags = (uint32_t)(selected ? 1 : 0);
    return 1;
}

/* the product's 32-byte record of a projected splat: colour and opacity rounded to f16 (RN-even) */
static void to_record(const orc_splat_f32* s, b200gs_splat* out) {
    memset(out, 0, sizeof *out);
    out->mx = s->mx; out->my = s->my;
    out->radius = (uint16_t)s->radius;
    out->opacity_h = orc_f32_to_f16(s->opacity);
    out->r_h = orc_f32_to_f16(s->r);
    out->g_h = orc_f32_to_f16(s->g);
    out->b_h = orc_f32_to_f16(s->b);
    out->ca = s->ca; out->cb = s->cb; out->cc = s->cc;
    out->flags = (uint16_t)s->flags;
}
static void from_record(const b200gs_splat* s, orc_splat_f32* out) {
    out->mx = s->mx; out->my = s->my;
    out->radius = (float)s->radius;
    out->opacity = orc_f16_to_f32(s->opacity_h);
    out->r = orc_f16_to_f32(s->r_h); out->g = orc_f16_to_f32(s->g_h); out->b = orc_f16_to_f32(s->b_h);
    out->ca = s->ca; out->cb = s->cb; out->cc = s->cc;
    out->flags = s->flags;
}

/* Outputs in ascending Gaussian index order.  `splats` (the product's record, f16 colour / opacity) and
 * `splats_f32` (the oracle's own fp32 splat, nothing rounded) may each be NULL. */
static uint64_t preprocess_any(const orc_frame* f, const orc_model* m, uint32_t* indices, uint32_t* keys, b200gs_splat* splats,
                               orc_splat_f32* splats_f32) {
    pre_ctx c;
    pre_setup(f, m, &c);
    /* two-pass so that the output keeps ascending-index order under threads */
    int nt = orc_num_threads();
    uint64_t* cnt = (uint64_t*)calloc((size_t)nt + 1, sizeof(uint64_t));
    uint8_t* vis = (uint8_t*)malloc(m->n ? m->n : 1);
    uint64_t chunk = (m->n + (uint64_t)nt - 1) / (uint64_t)nt;
#pragma omp parallel for schedule(static, 1)
    for (int t = 0; t < nt; t++) {
        uint64_t lo = (uint64_t)t * chunk, hi = lo + chunk < m->n ? lo + chunk : m->n;
        uint64_t k = 0;
        for (uint64_t i = lo; i < hi; i++) {
            uint32_t key;
            orc_splat_f32 s;
            vis[i] = (uint8_t)pre_one(f, m, &c, i, &key, &s, NULL);
            k += vis[i];
        }
        cnt[t + 1] = k;
    }
    for (int t = 0; t < nt; t++) cnt[t + 1] += cnt[t];
#pragma omp parallel for schedule(static, 1)
    for (int t = 0; t < nt; t++) {
        uint64_t lo = (uint64_t)t * chunk, hi = lo + chunk < m->n ? lo + chunk : m->n;
        uint64_t k = cnt[t];
        for (uint64_t i = lo; i < hi; i++) {
            if (!vis[i]) continue;
            uint32_t key;
            orc_splat_f32 s;
            memset(&s, 0, sizeof s);
            pre_one(f, m, &c, i, &key, &s, NULL);
            if (indices) indices[k] = (uint32_t)i;
            if (keys) keys[k] = key;
            if (splats) to_record(&s, &splats[k]);
            if (splats_f32) splats_f32[k] = s;
            k++;
        }
    }
    uint64_t v = cnt[nt];
    free(cnt);
    free(vis);
    return v;
}
uint64_t orc_preprocess(const orc_frame* f, const orc_model* m, uint32_t* indices, uint32_t* keys, b200gs_splat* splats) {
    return preprocess_any(f, m, indices, keys, splats, NULL);
}
uint64_t orc_preprocess_f32(const orc_frame* f, const orc_model* m, uint32_t* indices, uint32_t* keys, orc_splat_f32* splats) {
    return preprocess_any(f, m, indices, keys, NULL, splats);
}

void orc_query_selection(const orc_frame* f, const orc_model* m, uint32_t* words_out) {
    pre_ctx c;
    pre_setup(f, m, &c);
    uint64_t nw = (m->n + 31) / 32;
    for (uint64_t w = 0; w < nw; w++) words_out[w] = 0;
    for (uint64_t i = 0; i < m->n; i++) {
        uint32_t key;
        orc_splat_f32 s;
        int sel = 0;
        int vis = pre_one(f, m, &c, i, &key, &s, &sel);
        /* a culled Gaussian is never hit: Set clears it, Add / Remove keep its old state */
        if (!vis) sel = (f->query.kind >= B200GS_QUERY_RECT && f->query.op == B200GS_SELECT_SET)
                            ? 0 : (m->selection ? (int)((m->selection[i >> 5] >> (i & 31)) & 1u) : 0);
        if (sel) words_out[i >> 5] |= 1u << (i & 31);
    }
}

/* --------------------------------------------------------------- sort (a2)
 * Replaces viewer.radix_sorter.sort (src/tab/scene.rs:865-869): stable ascending sort of
 * (key = f32 depth bits as u32, value = index).  Restated as a merge sort on
 * (key, original position) so that it shares nothing with the GPU radix sort. */
typedef struct kv { uint32_t key, pos; } kv;
static void merge_sort_kv(kv* a, kv* tmp, uint64_t n) {
    if (n < 2) return;
    if (n <= 16) {
        for (uint64_t i = 1; i < n; i++) {
            kv x = a[i];
            uint64_t j = i;
            while (j > 0 && (a[j - 1].key > x.key)) { a[j] = a[j - 1]; j--; }
            a[j] = x;
        }
        return;
    }
    uint64_t h = n / 2;
    merge_sort_kv(a, tmp, h);
    merge_sort_kv(a + h, tmp + h, n - h);
    uint64_t i = 0, j = h, k = 0;
    while (i < h && j < n) tmp[k++] = (a[j].key < a[i].key) ? a[j++] : a[i++];
    while (i < h) tmp[k++] = a[i++];
    while (j < n) tmp[k++] = a[j++];
    memcpy(a, tmp, n * sizeof(kv));
}
static void sort_perm(uint64_t n, const uint32_t* keys, uint32_t mask, kv* a) {
    kv* tmp = (kv*)malloc((n ? n : 1) * sizeof(kv));
    for (uint64_t i = 0; i < n; i++) { a[i].key = keys[i] & mask; a[i].pos = (uint32_t)i; }
    int nt = orc_num_threads();
    if (n < 65536 || nt < 2) merge_sort_kv(a, tmp, n);
    else {
        /* sort nt runs in parallel, then merge them pairwise (stable: left run wins ties) */
        uint64_t run = (n + (uint64_t)nt - 1) / (uint64_t)nt;
#pragma omp parallel for schedule(static, 1)
        for (int t = 0; t < nt; t++) {
            uint64_t lo = (uint64_t)t * run, hi = lo + run < n ? lo + run : n;
            if (lo < hi) merge_sort_kv(a + lo, tmp + lo, hi - lo);
        }
        for (uint64_t w = run; w < n; w *= 2) {
            int64_t npairs = (int64_t)((n + 2 * w - 1) / (2 * w));
#pragma omp parallel for schedule(static, 1)
            for (int64_t pi = 0; pi < npairs; pi++) {
                uint64_t lo = (uint64_t)pi * 2 * w, mid = lo + w < n ? lo + w : n, hi = lo + 2 * w < n ? lo + 2 * w : n;
                uint64_t i = lo, j = mid, k = lo;
                while (i < mid && j < hi) tmp[k++] = (a[j].key < a[i].key) ? a[j++] : a[i++];
                while (i < mid) tmp[k++] = a[i++];
                while (j < hi) tmp[k++] = a[j++];
                memcpy(a + lo, tmp + lo, (hi - lo) * sizeof(kv));
            }
        }
    }
    free(tmp);
}
void orc_sort_pairs(uint64_t n, uint32_t* keys, uint32_t* values, uint32_t bits) {
    kv* a = (kv*)malloc((n ? n : 1) * sizeof(kv));
    uint32_t mask = bits >= 32 ? 0xffffffffu : ((1u << bits) - 1u);
    sort_perm(n, keys, mask, a);
    uint32_t* k2 = (uint32_t*)malloc((n ? n : 1) * 4);
    uint32_t* v2 = (uint32_t*)malloc((n ? n : 1) * 4);
    for (uint64_t i = 0; i < n; i++) { k2[i] = keys[a[i].pos]; v2[i] = values ? values[a[i].pos] : 0; }
    memcpy(keys, k2, n * 4);
    if (values) memcpy(values, v2, n * 4);
    free(k2); free(v2); free(a);
}
static void sort_any(uint64_t v, uint32_t* keys, uint32_t* indices, void* splats, size_t splat_bytes) {
    kv* a = (kv*)malloc((v ? v : 1) * sizeof(kv));
    sort_perm(v, keys, 0xffffffffu, a);
    uint32_t* k2 = (uint32_t*)malloc((v ? v : 1) * 4);
    uint32_t* i2 = (uint32_t*)malloc((v ? v : 1) * 4);
    for (uint64_t i = 0; i < v; i++) { k2[i] = keys[a[i].pos]; i2[i] = indices[a[i].pos]; }
    memcpy(keys, k2, v * 4);
    memcpy(indices, i2, v * 4);
    free(k2); free(i2);
    if (splats) {
        uint8_t* s2 = (uint8_t*)malloc((v ? v : 1) * splat_bytes);
        for (uint64_t i = 0; i < v; i++) memcpy(s2 + i * splat_bytes, (const uint8_t*)splats + (size_t)a[i].pos * splat_bytes, splat_bytes);
        memcpy(splats, s2, v * splat_bytes);
        free(s2);
    }
    free(a);
}
void orc_sort(uint64_t v, uint32_t* keys, uint32_t* indices, b200gs_splat* splats) { sort_any(v, keys, indices, splats, sizeof(b200gs_splat)); }
void orc_sort_f32(uint64_t v, uint32_t* keys, uint32_t* indices, orc_splat_f32* splats) { sort_any(v, keys, indices, splats, sizeof(orc_splat_f32)); }

/* ---------------------------------------------------------- compositing (a3)
 * Fragment rule [CANON §8c.7]: a splat covers the pixels of the screen-aligned square of
 * half-size `radius` around its centre; alpha = min(0.99, o·exp(power)),
 * power = -½(a dx² + c dy²) - b dx dy; dropped if power > 0 or alpha < 1/255.
 * Ellipse/Point display (src/tab/transform.rs:129-131): flat alpha min(0.99, o) inside
 * d² <= ORC_FLAT_D2 [our definition; the crate's is unknown]. */
static inline int splat_alpha(const orc_frame* f, const orc_splat_f32* s, float op, float px, float py, float* alpha) {
    float dx = px - s->mx, dy = py - s->my;
    float power = -0.5f * (s->ca * dx * dx + s->cc * dy * dy) - s->cb * dx * dy;
    if (power > 0.0f) return 0;
    float al;
    if (f->display_mode == B200GS_DISPLAY_SPLAT) al = fminf(ORC_ALPHA_MAX, op * expf(power));
    else al = (power >= -0.5f * ORC_FLAT_D2) ? fminf(ORC_ALPHA_MAX, op) : 0.0f;
    if (al < ORC_ALPHA_MIN) return 0;
    *alpha = al;
    return 1;
}
static inline int splat_bounds(const orc_frame* f, const orc_splat_f32* s, int* x0, int* x1, int* y0, int* y1) {
    if (!(s->radius > 0.0f)) return 0;
    float r = s->radius;
    float fx0 = ceilf(s->mx - r), fx1 = floorf(s->mx + r), fy0 = ceilf(s->my - r), fy1 = floorf(s->my + r);
    float W = f->size[0], H = f->size[1];
    if (fx0 < 0.0f) fx0 = 0.0f;
    if (fy0 < 0.0f) fy0 = 0.0f;
    if (fx1 > W - 1.0f) fx1 = W - 1.0f;
    if (fy1 > H - 1.0f) fy1 = H - 1.0f;
    if (!(fx0 <= fx1 && fy0 <= fy1)) return 0;
    *x0 = (int)fx0; *x1 = (int)fx1; *y0 = (int)fy0; *y1 = (int)fy1;
    return 1;
}
static void finish_image(const orc_frame* f, const float* acc, uint64_t npx, float* rgba_f, uint8_t* rgba8) {
    (void)f;
    for (uint64_t p = 0; p < npx; p++)
        for (int c = 0; c < 4; c++) {
            float v = acc[4 * p + c];
            if (rgba_f) rgba_f[4 * p + c] = v;
            if (rgba8) rgba8[4 * p + c] = unorm8(v);
        }
}

/* reference-style: hardware "over" blending, back to front, premultiplied [§8c.8]:
 * C = c·α + C·(1-α), A = α + A·(1-α), starting from the clear colour. */
void orc_composite_b2f_f32(const orc_frame* f, const orc_splat_f32* splats, uint64_t n_total, float* rgba_f, uint8_t* rgba8) {
    int W = (int)f->size[0], H = (int)f->size[1];
    float* acc = (float*)malloc((size_t)W * H * 16);
    int nt = orc_num_threads();
    int band = (H + nt * 4 - 1) / (nt * 4);
    if (band < 1) band = 1;
    int nb = (H + band - 1) / band;
#pragma omp parallel for schedule(dynamic, 1)
    for (int bi = 0; bi < nb; bi++) {
        int by0 = bi * band, by1 = by0 + band - 1 < H - 1 ? by0 + band - 1 : H - 1;
        for (int y = by0; y <= by1; y++)
            for (int x = 0; x < W; x++)
                for (int c = 0; c < 4; c++) acc[4 * ((size_t)y * W + x) + c] = f->background[c];
        for (uint64_t k = n_total; k-- > 0;) {
            const orc_splat_f32* s = &splats[k];
            int x0, x1, y0, y1;
            if (!splat_bounds(f, s, &x0, &x1, &y0, &y1)) continue;
            if (y0 < by0) y0 = by0;
            if (y1 > by1) y1 = by1;
            if (y0 > y1) continue;
            float op = s->opacity;
            float col[3] = {s->r, s->g, s->b};
            for (int y = y0; y <= y1; y++)
                for (int x = x0; x <= x1; x++) {
                    float al;
                    if (!splat_alpha(f, s, op, (float)x, (float)y, &al)) continue;
                    float* a = &acc[4 * ((size_t)y * W + x)];
                    float om = 1.0f - al;
                    a[0] = col[0] * al + a[0] * om;
                    a[1] = col[1] * al + a[1] * om;
                    a[2] = col[2] * al + a[2] * om;
                    a[3] = al + a[3] * om;
                }
        }
    }
    finish_image(f, acc, (uint64_t)W * H, rgba_f, rgba8);
    free(acc);
}

/* the new design's order: front to back, C += c·α·T, T *= (1-α), stop when T < ORC_T_EPS */
uint64_t orc_composite_f2b_f32(const orc_frame* f, const orc_splat_f32* splats, uint64_t n_total, float* rgba_f, uint8_t* rgba8) {
    int W = (int)f->size[0], H = (int)f->size[1];
    float* acc = (float*)malloc((size_t)W * H * 16);
    float* Tr = (float*)malloc((size_t)W * H * 4);
    int nt = orc_num_threads();
    int band = (H + nt * 4 - 1) / (nt * 4);
    if (band < 1) band = 1;
    int nb = (H + band - 1) / band;
    uint64_t evals = 0;
#pragma omp parallel for schedule(dynamic, 1) reduction(+ : evals)
    for (int bi = 0; bi < nb; bi++) {
        int by0 = bi * band, by1 = by0 + band - 1 < H - 1 ? by0 + band - 1 : H - 1;
        for (int y = by0; y <= by1; y++)
            for (int x = 0; x < W; x++) {
                size_t p = (size_t)y * W + x;
                acc[4 * p] = acc[4 * p + 1] = acc[4 * p + 2] = acc[4 * p + 3] = 0.0f;
                Tr[p] = 1.0f;
            }
        for (uint64_t k = 0; k < n_total; k++) {
            const orc_splat_f32* s = &splats[k];
            int x0, x1, y0, y1;
            if (!splat_bounds(f, s, &x0, &x1, &y0, &y1)) continue;
            if (y0 < by0) y0 = by0;
            if (y1 > by1) y1 = by1;
            if (y0 > y1) continue;
            float op = s->opacity;
            float col[3] = {s->r, s->g, s->b};
            for (int y = y0; y <= y1; y++)
                for (int x = x0; x <= x1; x++) {
                    size_t p = (size_t)y * W + x;
                    float T = Tr[p];
                    if (T < ORC_T_EPS) continue;
                    evals++;
                    float al;
                    if (!splat_alpha(f, s, op, (float)x, (float)y, &al)) continue;
                    float w = al * T;
                    acc[4 * p] += col[0] * w;
                    acc[4 * p + 1] += col[1] * w;
                    acc[4 * p + 2] += col[2] * w;
                    Tr[p] = T * (1.0f - al);
                }
        }
        for (int y = by0; y <= by1; y++)
            for (int x = 0; x < W; x++) {
                size_t p = (size_t)y * W + x;
                float T = Tr[p];
                acc[4 * p] += f->background[0] * T;
                acc[4 * p + 1] += f->background[1] * T;
                acc[4 * p + 2] += f->background[2] * T;
                acc[4 * p + 3] = (1.0f - T) + f->background[3] * T;
            }
    }
    finish_image(f, acc, (uint64_t)W * H, rgba_f, rgba8);
    free(acc);
    free(Tr);
    return evals;
}

/* the same two compositors over the product's 32-byte records (f16 colour / opacity read back exactly) */
static orc_splat_f32* records_to_f32(const b200gs_splat* splats, uint64_t n) {
    orc_splat_f32* t = (orc_splat_f32*)malloc((n ? n : 1) * sizeof(orc_splat_f32));
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < (int64_t)n; i++) from_record(&splats[i], &t[i]);
    return t;
}
void orc_composite_b2f(const orc_frame* f, const b200gs_splat* splats, uint64_t n_total, float* rgba_f, uint8_t* rgba8) {
    orc_splat_f32* t = records_to_f32(splats, n_total);
    orc_composite_b2f_f32(f, t, n_total, rgba_f, rgba8);
    free(t);
}
uint64_t orc_composite_f2b(const orc_frame* f, const b200gs_splat* splats, uint64_t n_total, float* rgba_f, uint8_t* rgba8) {
    orc_splat_f32* t = records_to_f32(splats, n_total);
    uint64_t e = orc_composite_f2b_f32(f, t, n_total, rgba_f, rgba8);
    free(t);
    return e;
}

/* ----------------------------------------------------------- model order (a4)
 * src/tab/scene.rs:533-558: squared distance camera -> world_center, farthest first;
 * world_center = quat*(center*scale)+pos (src/app.rs:1044-1046). */
void orc_order_models(const orc_frame* f, const orc_model* models, const float* centers, uint32_t n, uint32_t* order) {
    float* d = (float*)malloc((n ? n : 1) * sizeof(float));
    float V[4][4];
    for (int r = 0; r < 4; r++)
        for (int k = 0; k < 4; k++) V[r][k] = f->view[k * 4 + r];
    float cam[3];
    for (int a = 0; a < 3; a++) cam[a] = -(V[0][a] * V[0][3] + V[1][a] * V[1][3] + V[2][a] * V[2][3]);
    for (uint32_t i = 0; i < n; i++) {
        float R[3][3];
        quat_to_mat3(models[i].quat, R);
        float cs[3] = {centers[3 * i] * models[i].scale[0], centers[3 * i + 1] * models[i].scale[1],
                       centers[3 * i + 2] * models[i].scale[2]};
        float acc = 0.0f;
        for (int r = 0; r < 3; r++) {
            float w = R[r][0] * cs[0] + R[r][1] * cs[1] + R[r][2] * cs[2] + models[i].pos[r];
            float dd = w - cam[r];
            acc = acc + dd * dd;
        }
        d[i] = acc;
        order[i] = i;
    }
    /* stable insertion sort, descending distance */
    for (uint32_t i = 1; i < n; i++) {
        uint32_t x = order[i];
        uint32_t j = i;
        while (j > 0 && d[order[j - 1]] < d[x]) { order[j] = order[j - 1]; j--; }
        order[j] = x;
    }
    free(d);
}

/* ------------------------------------------------------------- whole frame */
static uint64_t render_frame_any(const orc_frame* f, const orc_model* far_to_near, uint32_t n_models, int front_to_back, int fp32,
                                 uint8_t* rgba8, double stage_seconds[3]) {
    uint64_t cap = 0;
    for (uint32_t i = 0; i < n_models; i++) cap += far_to_near[i].n;
    b200gs_splat* all = fp32 ? NULL : (b200gs_splat*)malloc((cap ? cap : 1) * sizeof(b200gs_splat));
    orc_splat_f32* all32 = fp32 ? (orc_splat_f32*)malloc((cap ? cap : 1) * sizeof(orc_splat_f32)) : NULL;
    uint32_t* keys = (uint32_t*)malloc((cap ? cap : 1) * 4);
    uint32_t* idx = (uint32_t*)malloc((cap ? cap : 1) * 4);
    uint64_t total = 0;
    double t_pre = 0, t_sort = 0, t_comp = 0;
    /* nearest model first in the concatenated near->far list */
    for (uint32_t mi = n_models; mi-- > 0;) {
        const orc_model* m = &far_to_near[mi];
        double t0 = now_s();
        uint64_t v = preprocess_any(f, m, idx + total, keys + total, all ? all + total : NULL, all32 ? all32 + total : NULL);
        double t1 = now_s();
        if (fp32) orc_sort_f32(v, keys + total, idx + total, all32 + total);
        else orc_sort(v, keys + total, idx + total, all + total);
        double t2 = now_s();
        t_pre += t1 - t0;
        t_sort += t2 - t1;
        total += v;
    }
    double t3 = now_s();
    if (fp32) {
        if (front_to_back) orc_composite_f2b_f32(f, all32, total, NULL, rgba8);
        else orc_composite_b2f_f32(f, all32, total, NULL, rgba8);
    } else {
        if (front_to_back) orc_composite_f2b(f, all, total, NULL, rgba8);
        else orc_composite_b2f(f, all, total, NULL, rgba8);
    }
    t_comp = now_s() - t3;
    if (stage_seconds) { stage_seconds[0] = t_pre; stage_seconds[1] = t_sort; stage_seconds[2] = t_comp; }
    free(all); free(all32); free(keys); free(idx);
    return total;
}
uint64_t orc_render_frame(const orc_frame* f, const orc_model* far_to_near, uint32_t n_models, int front_to_back,
                          uint8_t* rgba8, double stage_seconds[3]) {
    return render_frame_any(f, far_to_near, n_models, front_to_back, 0, rgba8, stage_seconds);
}
/* the whole path in fp32 end to end: the projected splat is never rounded to the product's f16 record */
uint64_t orc_render_frame_f32(const orc_frame* f, const orc_model* far_to_near, uint32_t n_models, int front_to_back,
                              uint8_t* rgba8, double stage_seconds[3]) {
    return render_frame_any(f, far_to_near, n_models, front_to_back, 1, rgba8, stage_seconds);
}

/* ------------------------------------------------------------ export (N4)
 * Gaussians::write_ply(writer, Option<&edits>, Option<mask>) as called at src/app.rs:904-914, 935-943: masked-out
 * and hidden Gaussians are dropped, an enabled edit pod is baked into the base colour / opacity (u8), then the
 * inverse of Gaussian::from(PlyGaussianPod).  [RECALLED: the crate's own treatment is not readable here.] */
uint64_t orc_export_edited(const b200gs_gaussian* in, uint64_t count, const b200gs_edit_pod* edits, const uint32_t* mask,
                           b200gs_ply_gaussian* out) {
    uint64_t k = 0;
    for (uint64_t i = 0; i < count; i++) {
        if (mask && !((mask[i >> 5] >> (i & 31)) & 1u)) continue;
        b200gs_gaussian g = in[i];
        if (edits) {
            const b200gs_edit_pod* e = &edits[i];
            if ((e->flag & B200GS_EDIT_ENABLED) && (e->flag & B200GS_EDIT_HIDDEN)) continue;
            float rgb[3] = {(float)g.color[0] / 255.0f, (float)g.color[1] / 255.0f, (float)g.color[2] / 255.0f};
            float op = (float)g.color[3] / 255.0f;
            orc_apply_edit(e, rgb, &op);
            for (int c = 0; c < 3; c++) g.color[c] = unorm8(rgb[c]);
            g.color[3] = unorm8(op);
        }
        b200gs_ply_gaussian* p = &out[k++];
        memcpy(p->pos, g.pos, 12);
        p->normal[0] = p->normal[1] = p->normal[2] = 0.0f;
        for (int c = 0; c < 3; c++) p->f_dc[c] = ((float)g.color[c] / 255.0f - 0.5f) / ORC_SH_C0;
        for (int kk = 0; kk < 15; kk++)
            for (int c = 0; c < 3; c++) p->f_rest[c * 15 + kk] = g.sh[3 * kk + c];
        float o = (float)g.color[3] / 255.0f;
        o = o < 1e-6f ? 1e-6f : (o > 1.0f - 1e-6f ? 1.0f - 1e-6f : o);
        p->opacity = logf(o / (1.0f - o));
        for (int a = 0; a < 3; a++) p->scale[a] = logf(g.scale[a]);
        p->rot[0] = g.rot[3]; p->rot[1] = g.rot[0]; p->rot[2] = g.rot[1]; p->rot[3] = g.rot[2];
    }
    return k;
}

/* ------------------------------------------------------------ mask eval (N2)
 * Replaces mask_evaluator.evaluate (src/tab/scene.rs:2124-2131, 2201-2209) with the op tree
 * of src/app.rs:1816-1837 flattened to postfix.  A Gaussian is tested by its WORLD position
 * (model transform applied) against each shape: p_s = S^-1 R^T (p_w - pos); box: |p_s| <= ½
 * per axis; ellipsoid: |p_s|² <= ¼. */
void orc_eval_mask(const orc_model* m, const b200gs_mask_op* postfix, uint32_t n_ops, const b200gs_mask_shape* shapes,
                   uint32_t n_shapes, uint32_t* words) {
    uint64_t nw = (m->n + 31) / 32;
    for (uint64_t w = 0; w < nw; w++) words[w] = 0;
    float R[3][3];
    quat_to_mat3(m->quat, R);
    uint32_t rb = orc_record_bytes(m->sh, m->cov3d);
    float (*SR)[3][3] = (float (*)[3][3])malloc((n_shapes ? n_shapes : 1) * sizeof(float[3][3]));
    for (uint32_t s = 0; s < n_shapes; s++) quat_to_mat3(shapes[s].quat, SR[s]);
    for (uint64_t i = 0; i < m->n; i++) {
        float p[3];
        memcpy(p, (const uint8_t*)m->packed + (size_t)i * rb, 12);
        float ps[3] = {m->scale[0] * p[0], m->scale[1] * p[1], m->scale[2] * p[2]};
        float pw[3];
        for (int r = 0; r < 3; r++) pw[r] = R[r][0] * ps[0] + R[r][1] * ps[1] + R[r][2] * ps[2] + m->pos[r];
        uint8_t stack[64];
        int sp = 0;
        if (n_ops == 0) stack[sp++] = 1;
        for (uint32_t o = 0; o < n_ops; o++) {
            uint32_t k = postfix[o].kind;
            if (k == B200GS_MASKOP_RESET) { stack[sp++] = 1; continue; }
            if (k == B200GS_MASKOP_SHAPE) {
                const b200gs_mask_shape* sh = &shapes[postfix[o].arg];
                float (*Q)[3] = SR[postfix[o].arg];
                float d[3] = {pw[0] - sh->pos[0], pw[1] - sh->pos[1], pw[2] - sh->pos[2]};
                float l[3];
                for (int c = 0; c < 3; c++) l[c] = (Q[0][c] * d[0] + Q[1][c] * d[1] + Q[2][c] * d[2]) / sh->scale[c];
                uint8_t in;
                if (sh->kind == B200GS_MASK_BOX) in = fabsf(l[0]) <= 0.5f && fabsf(l[1]) <= 0.5f && fabsf(l[2]) <= 0.5f;
                else in = (l[0] * l[0] + l[1] * l[1] + l[2] * l[2]) <= 0.25f;
                stack[sp++] = in;
                continue;
            }
            if (k == B200GS_MASKOP_COMPLEMENT) { stack[sp - 1] = !stack[sp - 1]; continue; }
            uint8_t b = stack[--sp], a = stack[--sp], r = 0;
            if (k == B200GS_MASKOP_UNION) r = a | b;
            else if (k == B200GS_MASKOP_INTERSECTION) r = a & b;
            else if (k == B200GS_MASKOP_DIFFERENCE) r = a & (uint8_t)!b;
            else if (k == B200GS_MASKOP_SYMDIFF) r = a ^ b;
            stack[sp++] = r;
        }
        if (sp > 0 && stack[sp - 1]) words[i >> 5] |= 1u << (i & 31);
    }
    free(SR);
}
