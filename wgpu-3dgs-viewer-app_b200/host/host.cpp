// host.cpp — host side of the B200 3DGS render core (no CUDA in this file).
//
// What the reference keeps on the CPU inside crate `wgpu-3dgs-viewer` and the app:
//   * Gaussian::from(PlyGaussianPod)                      src/app.rs:1066
//   * GaussiansBuffer::update_range's packing into the 8 GaussianPod layouts
//                                                         src/tab/scene.rs:2069-2085, app.rs:250-257
//   * Gaussians::read_ply_header / read_ply_gaussians / write_ply (streaming Inria PLY)
//                                                         src/app.rs:1056-1070, 910-914
//   * camera matrices (glam look_at_rh / perspective_rh)  src/app.rs:1236-1244
//   * model Euler rotation                                src/app.rs:1123-1130
// plus the deterministic synthetic-scene generator of SURVEY.md §8d used by bench and tests.
// Built with -ffp-contract=off so that float results do not depend on FMA availability.
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <string>
#include <thread>
#include <vector>

#include "../csrc/host_api.h"

// ------------------------------------------------------------------ small helpers
template <typename F>
static void parallel_for(uint64_t n, F&& fn) {
    unsigned hw = std::thread::hardware_concurrency();
    uint64_t nt = std::min<uint64_t>(hw ? hw : 1, (n + 16383) / 16384);
    if (nt <= 1) { fn(0, n); return; }
    std::vector<std::thread> th;
    uint64_t per = (n + nt - 1) / nt;
    for (uint64_t t = 0; t < nt; t++) {
        uint64_t lo = t * per, hi = std::min(n, lo + per);
        if (lo >= hi) break;
        th.emplace_back([=, &fn] { fn(lo, hi); });
    }
    for (auto& t : th) t.join();
}

static uint32_t sh_field_bytes(uint32_t sh) { return sh == 0 ? 180u : sh == 1 ? 92u : sh == 2 ? 48u : 0u; }

uint32_t gs_record_bytes(uint32_t sh, uint32_t cov3d) {
    if (sh > 3 || cov3d > 1) return 0;
    return 16u + sh_field_bytes(sh) + (cov3d == 0 ? 24u : 12u);
}

// IEEE binary16 conversion, round-to-nearest-even (what `half::f16::from_f32` does)
static uint16_t f32_to_f16(float f) {
    uint32_t x;
    memcpy(&x, &f, 4);
    const uint32_t sign = (x >> 16) & 0x8000u;
    const int32_t exp = (int32_t)((x >> 23) & 0xffu) - 127;
    uint32_t man = x & 0x7fffffu;
    if (exp == 128) return (uint16_t)(sign | 0x7c00u | (man ? (0x200u | (man >> 13)) : 0u));
    if (exp > 15) return (uint16_t)(sign | 0x7c00u);
    if (exp >= -14) {
        uint32_t h = ((uint32_t)(exp + 15) << 10) | (man >> 13);
        uint32_t rest = man & 0x1fffu;
        if (rest > 0x1000u || (rest == 0x1000u && (h & 1u))) h += 1;  // carries into the exponent correctly
        return (uint16_t)(sign | h);
    }
    if (exp < -25) return (uint16_t)sign;
    man |= 0x800000u;
    const uint32_t shift = (uint32_t)(-exp - 14 + 13);  // 14..24
    uint32_t h = man >> shift;
    const uint32_t rest = man & ((1u << shift) - 1u), half = 1u << (shift - 1);
    if (rest > half || (rest == half && (h & 1u))) h += 1;
    return (uint16_t)(sign | h);
}
static float f16_to_f32(uint16_t h) {
    const uint32_t sign = ((uint32_t)h & 0x8000u) << 16;
    uint32_t exp = (h >> 10) & 0x1fu, man = h & 0x3ffu, x;
    if (exp == 31) x = sign | 0x7f800000u | (man << 13);
    else if (exp != 0) x = sign | ((exp + 112u) << 23) | (man << 13);
    else if (man == 0) x = sign;
    else {
        int e = -1;
        do { man <<= 1; e++; } while (!(man & 0x400u));
        x = sign | ((uint32_t)(112 - e) << 23) | ((man & 0x3ffu) << 13);
    }
    float f;
    memcpy(&f, &x, 4);
    return f;
}

static inline uint8_t to_unorm8(float x) {
    x = x > 0.0f ? x : 0.0f;
    x = x < 1.0f ? x : 1.0f;
    return (uint8_t)(x * 255.0f + 0.5f);
}

static void mat3_from_quat(const float q[4], float R[9]) {  // glam Mat3::from_quat, row-major out
    const float x = q[0], y = q[1], z = q[2], w = q[3];
    const float x2 = x + x, y2 = y + y, z2 = z + z;
    const float xx = x * x2, xy = x * y2, xz = x * z2, yy = y * y2, yz = y * z2, zz = z * z2;
    const float wx = w * x2, wy = w * y2, wz = w * z2;
    R[0] = 1.0f - (yy + zz); R[1] = xy - wz;          R[2] = xz + wy;
    R[3] = xy + wz;          R[4] = 1.0f - (xx + zz); R[5] = yz - wx;
    R[6] = xz - wy;          R[7] = yz + wx;          R[8] = 1.0f - (xx + yy);
}

// ------------------------------------------------------------------ PLY <-> Gaussian
extern "C" int b200gs_gaussian_from_ply(const b200gs_ply_gaussian* in, uint64_t count, b200gs_gaussian* out) {
    if ((!in || !out) && count) { gs_set_error("gaussian_from_ply: null argument"); return B200GS_ERR_INVALID; }
    parallel_for(count, [&](uint64_t lo, uint64_t hi) {
        for (uint64_t i = lo; i < hi; i++) {
            const b200gs_ply_gaussian& p = in[i];
            b200gs_gaussian& g = out[i];
            float w = p.rot[0], x = p.rot[1], y = p.rot[2], z = p.rot[3];
            const float len = sqrtf(x * x + y * y + z * z + w * w);
            if (len > 0.0f) { x = x / len; y = y / len; z = z / len; w = w / len; }
            else { x = y = z = 0.0f; w = 1.0f; }
            g.rot[0] = x; g.rot[1] = y; g.rot[2] = z; g.rot[3] = w;
            memcpy(g.pos, p.pos, 12);
            for (int a = 0; a < 3; a++) g.scale[a] = expf(p.scale[a]);
            for (int c = 0; c < 3; c++) g.color[c] = to_unorm8(0.5f + 0.28209479177387814f * p.f_dc[c]);
            g.color[3] = to_unorm8(1.0f / (1.0f + expf(-p.opacity)));
            for (int k = 0; k < 15; k++)
                for (int c = 0; c < 3; c++) g.sh[3 * k + c] = p.f_rest[c * 15 + k];
        }
    });
    return B200GS_OK;
}

// inverse mapping used by export (write_ply of edited models, src/app.rs:897-947); the colour
// and opacity round-trip through their u8 quantisation
extern "C" int b200gs_gaussian_to_ply(const b200gs_gaussian* in, uint64_t count, b200gs_ply_gaussian* out) {
    if ((!in || !out) && count) { gs_set_error("gaussian_to_ply: null argument"); return B200GS_ERR_INVALID; }
    parallel_for(count, [&](uint64_t lo, uint64_t hi) {
        for (uint64_t i = lo; i < hi; i++) {
            const b200gs_gaussian& g = in[i];
            b200gs_ply_gaussian& p = out[i];
            memcpy(p.pos, g.pos, 12);
            p.normal[0] = p.normal[1] = p.normal[2] = 0.0f;
            for (int c = 0; c < 3; c++) p.f_dc[c] = ((float)g.color[c] / 255.0f - 0.5f) / 0.28209479177387814f;
            for (int k = 0; k < 15; k++)
                for (int c = 0; c < 3; c++) p.f_rest[c * 15 + k] = g.sh[3 * k + c];
            float o = (float)g.color[3] / 255.0f;
            o = std::min(std::max(o, 1e-6f), 1.0f - 1e-6f);
            p.opacity = logf(o / (1.0f - o));
            for (int a = 0; a < 3; a++) p.scale[a] = logf(g.scale[a]);
            p.rot[0] = g.rot[3]; p.rot[1] = g.rot[0]; p.rot[2] = g.rot[1]; p.rot[3] = g.rot[2];
        }
    });
    return B200GS_OK;
}

// ------------------------------------------------------------------ packing
extern "C" int b200gs_pack_gaussians(uint32_t sh, uint32_t cov3d, const b200gs_gaussian* in, uint64_t count, void* out) {
    const uint32_t rb = gs_record_bytes(sh, cov3d);
    if (!rb) { gs_set_error("pack_gaussians: invalid layout"); return B200GS_ERR_INVALID; }
    if ((!in || !out) && count) { gs_set_error("pack_gaussians: null argument"); return B200GS_ERR_INVALID; }
    parallel_for(count, [&](uint64_t lo, uint64_t hi) {
        for (uint64_t i = lo; i < hi; i++) {
            const b200gs_gaussian& g = in[i];
            uint8_t* rec = (uint8_t*)out + i * rb;
            memcpy(rec, g.pos, 12);
            memcpy(rec + 12, g.color, 4);
            uint8_t* w = rec + 16;
            switch (sh) {
                case B200GS_SH_SINGLE: memcpy(w, g.sh, 180); break;
                case B200GS_SH_HALF: {
                    uint16_t h[46];
                    for (int k = 0; k < 45; k++) h[k] = f32_to_f16(g.sh[k]);
                    h[45] = 0;
                    memcpy(w, h, 92);
                    break;
                }
                case B200GS_SH_NORM8:
                    for (int k = 0; k < 45; k++) w[k] = to_unorm8((g.sh[k] + 1.0f) * 0.5f);
                    w[45] = w[46] = w[47] = 0;
                    break;
                default: break;
            }
            w += sh_field_bytes(sh);
            // Σ = (R S)(R S)^T, upper triangle
            float R[9], M[9];
            mat3_from_quat(g.rot, R);
            for (int r = 0; r < 3; r++)
                for (int c = 0; c < 3; c++) M[3 * r + c] = R[3 * r + c] * g.scale[c];
            float cv[6];
            int k = 0;
            for (int r = 0; r < 3; r++)
                for (int c = r; c < 3; c++, k++)
                    cv[k] = M[3 * r] * M[3 * c] + M[3 * r + 1] * M[3 * c + 1] + M[3 * r + 2] * M[3 * c + 2];
            if (cov3d == B200GS_COV3D_SINGLE) memcpy(w, cv, 24);
            else {
                uint16_t h[6];
                for (int j = 0; j < 6; j++) h[j] = f32_to_f16(cv[j]);
                memcpy(w, h, 12);
            }
        }
    });
    return B200GS_OK;
}

// Unpacking recovers pos, colour and SH exactly as the kernels decode them; rotation and scale
// are NOT recoverable from Σ without an eigen-decomposition, so they come back as identity /
// sqrt of the diagonal (enough for export previews; the app keeps the unpacked Gaussians on
// the host anyway, src/app.rs:1029-1031).
extern "C" int b200gs_unpack_gaussians(uint32_t sh, uint32_t cov3d, const void* in, uint64_t count, b200gs_gaussian* out) {
    const uint32_t rb = gs_record_bytes(sh, cov3d);
    if (!rb) { gs_set_error("unpack_gaussians: invalid layout"); return B200GS_ERR_INVALID; }
    if ((!in || !out) && count) { gs_set_error("unpack_gaussians: null argument"); return B200GS_ERR_INVALID; }
    parallel_for(count, [&](uint64_t lo, uint64_t hi) {
        for (uint64_t i = lo; i < hi; i++) {
            const uint8_t* rec = (const uint8_t*)in + i * rb;
            b200gs_gaussian& g = out[i];
            memcpy(g.pos, rec, 12);
            memcpy(g.color, rec + 12, 4);
            const uint8_t* w = rec + 16;
            if (sh == B200GS_SH_SINGLE) memcpy(g.sh, w, 180);
            else if (sh == B200GS_SH_HALF) {
                uint16_t h[46];
                memcpy(h, w, 92);
                for (int k = 0; k < 45; k++) g.sh[k] = f16_to_f32(h[k]);
            } else if (sh == B200GS_SH_NORM8) {
                for (int k = 0; k < 45; k++) g.sh[k] = (float)w[k] * (2.0f / 255.0f) - 1.0f;
            } else memset(g.sh, 0, 180);
            w += sh_field_bytes(sh);
            float cv[6];
            if (cov3d == B200GS_COV3D_SINGLE) memcpy(cv, w, 24);
            else {
                uint16_t h[6];
                memcpy(h, w, 12);
                for (int j = 0; j < 6; j++) cv[j] = f16_to_f32(h[j]);
            }
            g.rot[0] = g.rot[1] = g.rot[2] = 0.0f; g.rot[3] = 1.0f;
            g.scale[0] = sqrtf(std::max(cv[0], 0.0f));
            g.scale[1] = sqrtf(std::max(cv[3], 0.0f));
            g.scale[2] = sqrtf(std::max(cv[5], 0.0f));
        }
    });
    return B200GS_OK;
}

// ------------------------------------------------------------------ camera / transforms
static float dot3(const float* a, const float* b) { return a[0] * b[0] + a[1] * b[1] + a[2] * b[2]; }
static void cross3(const float* a, const float* b, float* o) {
    o[0] = a[1] * b[2] - a[2] * b[1];
    o[1] = a[2] * b[0] - a[0] * b[2];
    o[2] = a[0] * b[1] - a[1] * b[0];
}
static void normalize3(float* v) {
    const float l = sqrtf(dot3(v, v));
    v[0] = v[0] / l; v[1] = v[1] / l; v[2] = v[2] / l;
}

// glam Mat4::look_at_rh(eye, center, up) = look_to_rh(eye, center - eye, up)
extern "C" void b200gs_look_at_rh(const float eye[3], const float target[3], const float up[3], float out[16]) {
    float f[3] = {target[0] - eye[0], target[1] - eye[1], target[2] - eye[2]};
    normalize3(f);
    float s[3], u[3];
    cross3(f, up, s);
    normalize3(s);
    cross3(s, f, u);
    const float cols[16] = {s[0], u[0], -f[0], 0.0f, s[1], u[1], -f[1], 0.0f, s[2], u[2], -f[2], 0.0f,
                            -dot3(eye, s), -dot3(eye, u), dot3(eye, f), 1.0f};
    memcpy(out, cols, 64);
}

// glam Mat4::perspective_rh(fov_y, aspect, near, far), depth range 0..1
extern "C" void b200gs_perspective_rh(float vfov, float aspect, float z_near, float z_far, float out[16]) {
    const float sin_fov = sinf(0.5f * vfov), cos_fov = cosf(0.5f * vfov);
    const float h = cos_fov / sin_fov, w = h / aspect, r = z_far / (z_near - z_far);
    const float cols[16] = {w, 0, 0, 0, 0, h, 0, 0, 0, 0, r, -1.0f, 0, 0, r * z_near, 0};
    memcpy(out, cols, 64);
}

// Quat::from_euler(EulerRot::ZYX, rz, ry, rx) on degrees (src/app.rs:1123-1130) = qz * qy * qx
extern "C" void b200gs_quat_from_euler_zyx_deg(const float rot_deg[3], float quat_xyzw[4]) {
    const float k = 0.017453292519943295f;
    const float ax = rot_deg[0] * k * 0.5f, ay = rot_deg[1] * k * 0.5f, az = rot_deg[2] * k * 0.5f;
    const float sx = sinf(ax), cx = cosf(ax), sy = sinf(ay), cy = cosf(ay), sz = sinf(az), cz = cosf(az);
    quat_xyzw[0] = cz * cy * sx - sz * sy * cx;
    quat_xyzw[1] = cz * sy * cx + sz * cy * sx;
    quat_xyzw[2] = sz * cy * cx - cz * sy * sx;
    quat_xyzw[3] = cz * cy * cx + sz * sy * sx;
}

// ------------------------------------------------------------------ hit positions (row N3)
// gs::query::hit_pos_by_closest / hit_pos_by_alpha_range (src/tab/scene.rs:659-676): unproject the
// pixel centre at the chosen depth.  The crate's exact selection rule is not visible from the app;
// ours: closest = first hit; alpha_range = alpha-weighted mean ndc depth of the hits whose alpha is
// >= threshold.
static int unproject_pixel(const float view[16], const float proj[16], const float size[2], uint32_t px, uint32_t py,
                           float ndc_z, float out[3]) {
    // ndc of the pixel centre (row 0 = top)
    const double nx = (2.0 * ((double)px + 0.5)) / size[0] - 1.0, ny = 1.0 - (2.0 * ((double)py + 0.5)) / size[1];
    // perspective_rh-shaped projection: clip = (P00 x, P11 y, P22 z + P23, P32 z); solve for view-space point
    const double P00 = proj[0], P11 = proj[5], P22 = proj[10], P23 = proj[14], P32 = proj[11];
    const double den = (double)ndc_z * P32 - P22;
    if (den == 0.0 || P00 == 0.0 || P11 == 0.0) return B200GS_ERR_INVALID;
    const double zv = P23 / den, w = P32 * zv;
    const double xv = nx * w / P00, yv = ny * w / P11;
    // world = R^T (p_view - t) for a rigid view matrix (column-major)
    const double px_ = xv - view[12], py_ = yv - view[13], pz_ = zv - view[14];
    for (int a = 0; a < 3; a++) out[a] = (float)(view[4 * a + 0] * px_ + view[4 * a + 1] * py_ + view[4 * a + 2] * pz_);
    return B200GS_OK;
}
extern "C" int b200gs_hit_pos_by_closest(const b200gs_hit* hits, uint64_t n, const float view[16], const float proj[16],
                                         const float size[2], uint32_t px, uint32_t py, float pos_out[3]) {
    if (!hits || !n || !view || !proj || !size || !pos_out) { gs_set_error("hit_pos_by_closest: no hit"); return B200GS_ERR_INVALID; }
    return unproject_pixel(view, proj, size, px, py, hits[0].depth, pos_out);
}
extern "C" int b200gs_hit_pos_by_alpha_range(const b200gs_hit* hits, uint64_t n, float alpha_threshold, const float view[16],
                                             const float proj[16], const float size[2], uint32_t px, uint32_t py,
                                             float pos_out[3]) {
    if (!hits || !view || !proj || !size || !pos_out) { gs_set_error("hit_pos_by_alpha_range: null argument"); return B200GS_ERR_INVALID; }
    double wsum = 0.0, zsum = 0.0;
    for (uint64_t i = 0; i < n; i++)
        if (hits[i].alpha >= alpha_threshold) { wsum += hits[i].alpha; zsum += (double)hits[i].alpha * hits[i].depth; }
    if (wsum == 0.0) { gs_set_error("hit_pos_by_alpha_range: no hit above the threshold"); return B200GS_ERR_INVALID; }
    return unproject_pixel(view, proj, size, px, py, (float)(zsum / wsum), pos_out);
}

// ------------------------------------------------------------------ synthetic scene (§8d)
// Counter-based: value = f(seed, Gaussian index, stream), so any sub-range can be generated
// independently (and on any number of threads) with identical bytes.
namespace synth {
static inline uint64_t mix(uint64_t z) {
    z += 0x9E3779B97F4A7C15ULL;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
    return z ^ (z >> 31);
}
struct Rng {
    uint64_t base;
    Rng(uint64_t seed, uint64_t i) : base(mix(seed ^ (i * 0xD1342543DE82EF95ULL))) {}
    double uniform(uint64_t stream) const { return ((double)(mix(base + stream) >> 11) + 0.5) * 0x1.0p-53; }
    void normal2(uint64_t pair, double& a, double& b) const {
        const double r = sqrt(-2.0 * log(uniform(2 * pair)));
        const double t = 6.283185307179586476925286766559 * uniform(2 * pair + 1);
        a = r * cos(t);
        b = r * sin(t);
    }
};
constexpr int kClusters = 64;
constexpr uint64_t kClusterSalt = 0xC1A57E2500000000ULL;
}  // namespace synth

extern "C" int b200gs_synth_scene(uint64_t seed, uint64_t start, uint64_t count, b200gs_ply_gaussian* out) {
    if (!out && count) { gs_set_error("synth_scene: null argument"); return B200GS_ERR_INVALID; }
    using namespace synth;
    double centre[kClusters][3], sigma[kClusters];
    for (int c = 0; c < kClusters; c++) {
        Rng r(seed ^ kClusterSalt, (uint64_t)c);
        centre[c][0] = (2.0 * r.uniform(0) - 1.0) * 4.0;
        centre[c][1] = (2.0 * r.uniform(1) - 1.0) * 1.5;
        centre[c][2] = (2.0 * r.uniform(2) - 1.0) * 4.0;
        sigma[c] = 0.05 + 0.35 * r.uniform(3);
    }
    const double log_scale_mean = log(0.006);
    parallel_for(count, [&](uint64_t lo, uint64_t hi) {
        for (uint64_t j = lo; j < hi; j++) {
            Rng r(seed, start + j);
            b200gs_ply_gaussian& g = out[j];
            double nrm[64];
            for (uint64_t k = 0; k < 32; k++) r.normal2(k, nrm[2 * k], nrm[2 * k + 1]);
            if (r.uniform(100) < 0.1) {  // 10 % uniform background
                g.pos[0] = (float)((2.0 * r.uniform(102) - 1.0) * 4.0);
                g.pos[1] = (float)((2.0 * r.uniform(103) - 1.0) * 1.5);
                g.pos[2] = (float)((2.0 * r.uniform(104) - 1.0) * 4.0);
            } else {
                int c = (int)(r.uniform(101) * (double)kClusters);
                c = std::min(c, kClusters - 1);
                for (int a = 0; a < 3; a++) g.pos[a] = (float)(centre[c][a] + sigma[c] * nrm[a]);
            }
            g.normal[0] = g.normal[1] = g.normal[2] = 0.0f;
            for (int a = 0; a < 3; a++) g.scale[a] = (float)(log_scale_mean + 0.6 * nrm[4 + a]);
            for (int a = 0; a < 4; a++) g.rot[a] = (float)nrm[8 + a];
            g.opacity = (float)(0.5 + 2.0 * nrm[12]);
            for (int a = 0; a < 3; a++) g.f_dc[a] = (float)(0.8 * nrm[14 + a]);
            for (int c = 0; c < 3; c++)
                for (int k = 0; k < 15; k++) {
                    const double band = k < 3 ? 1.0 : (k < 8 ? 2.0 : 3.0);
                    g.f_rest[c * 15 + k] = (float)((0.15 / band) * nrm[18 + c * 15 + k]);
                }
        }
    });
    return B200GS_OK;
}

// ------------------------------------------------------------------ PLY io
// Inria 3DGS PLY: `element vertex N` with float properties x,y,z,nx,ny,nz,f_dc_0..2,
// f_rest_0..44,opacity,scale_0..2,rot_0..3 (any order, extra properties skipped, missing
// f_rest / normals read as 0).  binary_little_endian and ascii are both accepted.
struct b200gs_ply_reader {
    FILE* fp = nullptr;
    const uint8_t* mem = nullptr;
    size_t mem_size = 0, mem_pos = 0;
    bool ascii = false;
    uint64_t count = 0, done = 0;
    size_t stride = 0;                  // bytes per vertex (binary)
    struct Prop { int field; int type; size_t offset; };  // field = index into the 62 floats, -1 = skip
    std::vector<Prop> props;
    std::vector<uint8_t> buf;
};

static int ply_type_size(const std::string& t, int* code) {
    struct { const char* n; int sz; int c; } T[] = {
        {"char", 1, 0}, {"int8", 1, 0}, {"uchar", 1, 1}, {"uint8", 1, 1}, {"short", 2, 2}, {"int16", 2, 2},
        {"ushort", 2, 3}, {"uint16", 2, 3}, {"int", 4, 4}, {"int32", 4, 4}, {"uint", 4, 5}, {"uint32", 4, 5},
        {"float", 4, 6}, {"float32", 4, 6}, {"double", 8, 7}, {"float64", 8, 7}};
    for (auto& e : T)
        if (t == e.n) { *code = e.c; return e.sz; }
    return 0;
}
static int ply_field_index(const std::string& n) {
    static const char* base[] = {"x", "y", "z", "nx", "ny", "nz", "f_dc_0", "f_dc_1", "f_dc_2"};
    for (int i = 0; i < 9; i++)
        if (n == base[i]) return i;
    if (n.rfind("f_rest_", 0) == 0) {
        int k = atoi(n.c_str() + 7);
        return (k >= 0 && k < 45) ? 9 + k : -1;
    }
    if (n == "opacity") return 54;
    if (n.rfind("scale_", 0) == 0) { int k = atoi(n.c_str() + 6); return (k >= 0 && k < 3) ? 55 + k : -1; }
    if (n.rfind("rot_", 0) == 0) { int k = atoi(n.c_str() + 4); return (k >= 0 && k < 4) ? 58 + k : -1; }
    return -1;
}

static bool reader_getline(b200gs_ply_reader* r, std::string& line) {
    line.clear();
    if (r->fp) {
        int c;
        while ((c = fgetc(r->fp)) != EOF) {
            if (c == '\n') return true;
            if (c != '\r') line.push_back((char)c);
        }
        return !line.empty();
    }
    if (r->mem_pos >= r->mem_size) return false;
    while (r->mem_pos < r->mem_size) {
        char c = (char)r->mem[r->mem_pos++];
        if (c == '\n') return true;
        if (c != '\r') line.push_back(c);
    }
    return true;
}

static int reader_parse_header(b200gs_ply_reader* r) {
    std::string line;
    if (!reader_getline(r, line) || line != "ply") { gs_set_error("ply: missing magic"); return B200GS_ERR_FORMAT; }
    bool have_format = false, in_vertex = false, have_vertex = false;
    size_t off = 0;
    while (true) {
        if (!reader_getline(r, line)) { gs_set_error("ply: truncated header"); return B200GS_ERR_FORMAT; }
        char a[64] = {0}, b[64] = {0}, c[64] = {0};
        int n = sscanf(line.c_str(), "%63s %63s %63s", a, b, c);
        if (n < 1) continue;
        std::string kw = a;
        if (kw == "end_header") break;
        if (kw == "comment" || kw == "obj_info") continue;
        if (kw == "format") {
            if (n < 2) { gs_set_error("ply: bad format line"); return B200GS_ERR_FORMAT; }
            if (std::string(b) == "binary_little_endian") r->ascii = false;
            else if (std::string(b) == "ascii") r->ascii = true;
            else { gs_set_error("ply: unsupported format '%s'", b); return B200GS_ERR_FORMAT; }
            have_format = true;
        } else if (kw == "element") {
            if (n < 3) { gs_set_error("ply: bad element line"); return B200GS_ERR_FORMAT; }
            if (std::string(b) == "vertex") {
                if (have_vertex) { gs_set_error("ply: duplicate vertex element"); return B200GS_ERR_FORMAT; }
                r->count = strtoull(c, nullptr, 10);
                in_vertex = have_vertex = true;
            } else {
                if (!have_vertex) { gs_set_error("ply: element '%s' before vertex is not supported", b); return B200GS_ERR_FORMAT; }
                in_vertex = false;
            }
        } else if (kw == "property") {
            if (!in_vertex) continue;
            if (std::string(b) == "list") { gs_set_error("ply: list property in vertex element"); return B200GS_ERR_FORMAT; }
            int code = 0, sz = ply_type_size(b, &code);
            if (!sz || n < 3) { gs_set_error("ply: bad property line '%s'", line.c_str()); return B200GS_ERR_FORMAT; }
            r->props.push_back({ply_field_index(c), code, off});
            off += (size_t)sz;
        }
    }
    if (!have_format || !have_vertex) { gs_set_error("ply: header lacks format or vertex element"); return B200GS_ERR_FORMAT; }
    r->stride = off;
    bool has_xyz[3] = {false, false, false};
    for (auto& p : r->props)
        if (p.field >= 0 && p.field < 3) has_xyz[p.field] = true;
    if (!(has_xyz[0] && has_xyz[1] && has_xyz[2])) { gs_set_error("ply: vertex lacks x/y/z"); return B200GS_ERR_FORMAT; }
    return B200GS_OK;
}

static float ply_scalar(const uint8_t* p, int code) {
    switch (code) {
        case 0: return (float)*(const int8_t*)p;
        case 1: return (float)*p;
        case 2: { int16_t v; memcpy(&v, p, 2); return (float)v; }
        case 3: { uint16_t v; memcpy(&v, p, 2); return (float)v; }
        case 4: { int32_t v; memcpy(&v, p, 4); return (float)v; }
        case 5: { uint32_t v; memcpy(&v, p, 4); return (float)v; }
        case 6: { float v; memcpy(&v, p, 4); return v; }
        default: { double v; memcpy(&v, p, 8); return (float)v; }
    }
}

static int open_common(b200gs_ply_reader* r, b200gs_ply_reader** out, uint64_t* count) {
    int rc = reader_parse_header(r);
    if (rc != B200GS_OK) {
        if (r->fp) fclose(r->fp);
        delete r;
        return rc;
    }
    *out = r;
    if (count) *count = r->count;
    return B200GS_OK;
}

extern "C" int b200gs_ply_open(const char* path, b200gs_ply_reader** out, uint64_t* count) {
    if (!path || !out) { gs_set_error("ply_open: null argument"); return B200GS_ERR_INVALID; }
    *out = nullptr;
    FILE* fp = fopen(path, "rb");
    if (!fp) { gs_set_error("ply_open: cannot open '%s'", path); return B200GS_ERR_IO; }
    auto* r = new b200gs_ply_reader();
    r->fp = fp;
    return open_common(r, out, count);
}

extern "C" int b200gs_ply_open_memory(const void* data, size_t size, b200gs_ply_reader** out, uint64_t* count) {
    if (!data || !out) { gs_set_error("ply_open_memory: null argument"); return B200GS_ERR_INVALID; }
    *out = nullptr;
    auto* r = new b200gs_ply_reader();
    r->mem = (const uint8_t*)data;
    r->mem_size = size;
    return open_common(r, out, count);
}

extern "C" int b200gs_ply_read(b200gs_ply_reader* r, b200gs_ply_gaussian* out, uint64_t max, uint64_t* n_read) {
    if (!r || (!out && max) || !n_read) { gs_set_error("ply_read: null argument"); return B200GS_ERR_INVALID; }
    uint64_t want = std::min(max, r->count - r->done);
    *n_read = 0;
    if (want == 0) return B200GS_OK;
    if (r->ascii) {
        std::string line;
        for (uint64_t i = 0; i < want; i++) {
            if (!reader_getline(r, line)) { gs_set_error("ply: truncated vertex data"); return B200GS_ERR_IO; }
            float* f = (float*)&out[i];
            memset(f, 0, sizeof(b200gs_ply_gaussian));
            const char* s = line.c_str();
            for (auto& p : r->props) {
                char* e = nullptr;
                double v = strtod(s, &e);
                if (e == s) { gs_set_error("ply: malformed ascii vertex %llu", (unsigned long long)(r->done + i)); return B200GS_ERR_FORMAT; }
                s = e;
                if (p.field >= 0) f[p.field] = (float)v;
            }
            (*n_read)++;
        }
        r->done += want;
        return B200GS_OK;
    }
    const size_t bytes = (size_t)want * r->stride;
    const uint8_t* src;
    if (r->fp) {
        r->buf.resize(bytes);
        size_t got = fread(r->buf.data(), 1, bytes, r->fp);
        if (got != bytes) { gs_set_error("ply: truncated vertex data"); return B200GS_ERR_IO; }
        src = r->buf.data();
    } else {
        if (r->mem_size - r->mem_pos < bytes) { gs_set_error("ply: truncated vertex data"); return B200GS_ERR_IO; }
        src = r->mem + r->mem_pos;
        r->mem_pos += bytes;
    }
    // fast path: the canonical 62-float layout is a straight copy
    bool canonical = r->stride == sizeof(b200gs_ply_gaussian) && r->props.size() == 62;
    for (size_t k = 0; canonical && k < r->props.size(); k++)
        canonical = r->props[k].field == (int)k && r->props[k].type == 6;
    if (canonical) memcpy(out, src, bytes);
    else {
        parallel_for(want, [&](uint64_t lo, uint64_t hi) {
            for (uint64_t i = lo; i < hi; i++) {
                float* f = (float*)&out[i];
                memset(f, 0, sizeof(b200gs_ply_gaussian));
                const uint8_t* v = src + i * r->stride;
                for (auto& p : r->props)
                    if (p.field >= 0) f[p.field] = ply_scalar(v + p.offset, p.type);
            }
        });
    }
    r->done += want;
    *n_read = want;
    return B200GS_OK;
}

extern "C" int b200gs_ply_close(b200gs_ply_reader* r) {
    if (!r) return B200GS_OK;
    if (r->fp) fclose(r->fp);
    delete r;
    return B200GS_OK;
}

// GaussianEditPod applied to a displayed colour and opacity: colour (HSV shift/scale or override) -> contrast ->
// exposure -> gamma -> alpha, the same order as the preprocess kernel (csrc/preprocess.cu apply_edit).
static inline float clamp01f(float x) { return x < 0.0f ? 0.0f : (x > 1.0f ? 1.0f : x); }
static void edit_colour(const b200gs_edit_pod& e, float rgb[3], float& op) {
    if (!(e.flag & B200GS_EDIT_ENABLED)) return;
    if (e.flag & B200GS_EDIT_OVERRIDE_COLOR) {
        rgb[0] = e.color[0]; rgb[1] = e.color[1]; rgb[2] = e.color[2];
    } else {
        const float mx = fmaxf(rgb[0], fmaxf(rgb[1], rgb[2])), mn = fminf(rgb[0], fminf(rgb[1], rgb[2]));
        const float d = mx - mn;
        float h = 0.0f;
        if (d > 0.0f) {
            if (mx == rgb[0]) h = (rgb[1] - rgb[2]) / d;
            else if (mx == rgb[1]) h = 2.0f + (rgb[2] - rgb[0]) / d;
            else h = 4.0f + (rgb[0] - rgb[1]) / d;
            h = h / 6.0f;
            if (h < 0.0f) h = h + 1.0f;
        }
        float s = mx > 0.0f ? d / mx : 0.0f, v = mx;
        h = h + e.color[0];
        h = h - floorf(h);
        s = clamp01f(s * e.color[1]);
        v = v * e.color[2];
        const float h6 = h * 6.0f, i = floorf(h6), f = h6 - i;
        int k = ((int)i) % 6;
        if (k < 0) k += 6;
        const float p = v * (1.0f - s), q = v * (1.0f - s * f), t = v * (1.0f - s * (1.0f - f));
        switch (k) {
            case 0: rgb[0] = v; rgb[1] = t; rgb[2] = p; break;
            case 1: rgb[0] = q; rgb[1] = v; rgb[2] = p; break;
            case 2: rgb[0] = p; rgb[1] = v; rgb[2] = t; break;
            case 3: rgb[0] = p; rgb[1] = q; rgb[2] = v; break;
            case 4: rgb[0] = t; rgb[1] = p; rgb[2] = v; break;
            default: rgb[0] = v; rgb[1] = p; rgb[2] = q; break;
        }
    }
    const float ex = exp2f(e.exposure);
    for (int c = 0; c < 3; c++) {
        float v = (rgb[c] - 0.5f) * (1.0f + e.contrast) + 0.5f;
        v = v * ex;
        rgb[c] = powf(fmaxf(v, 0.0f), e.gamma);
    }
    op = clamp01f(op * e.alpha);
}

// Export of an edited, masked model: Gaussians::write_ply(writer, Option<&[GaussianEditPod]>, Option<mask words>)
// as the app calls it (src/app.rs:904-914, 935-943; pods and mask words downloaded at app.rs:789, 806).
extern "C" int b200gs_apply_edits_for_export(const b200gs_gaussian* in, uint64_t count, const b200gs_edit_pod* edits,
                                             const uint32_t* mask_words, b200gs_gaussian* out, uint64_t* n_out) {
    if ((!in || !out) && count) { gs_set_error("apply_edits_for_export: null argument"); return B200GS_ERR_INVALID; }
    if (!n_out) { gs_set_error("apply_edits_for_export: null n_out"); return B200GS_ERR_INVALID; }
    uint64_t k = 0;
    for (uint64_t i = 0; i < count; i++) {
        if (mask_words && !((mask_words[i >> 5] >> (i & 31)) & 1u)) continue;        // masked out: not exported
        b200gs_gaussian g = in[i];
        if (edits) {
            const b200gs_edit_pod& e = edits[i];
            if ((e.flag & B200GS_EDIT_ENABLED) && (e.flag & B200GS_EDIT_HIDDEN)) continue;   // hidden: not exported
            float rgb[3] = {(float)g.color[0] / 255.0f, (float)g.color[1] / 255.0f, (float)g.color[2] / 255.0f};
            float op = (float)g.color[3] / 255.0f;
            edit_colour(e, rgb, op);
            for (int c = 0; c < 3; c++) g.color[c] = to_unorm8(rgb[c]);
            g.color[3] = to_unorm8(op);
        }
        out[k++] = g;
    }
    *n_out = k;
    return B200GS_OK;
}

extern "C" int b200gs_ply_write_edited(const char* path, const b200gs_gaussian* gaussians, uint64_t count,
                                       const b200gs_edit_pod* edits_or_null, const uint32_t* mask_words_or_null) {
    if (!path || (!gaussians && count)) { gs_set_error("ply_write_edited: null argument"); return B200GS_ERR_INVALID; }
    std::vector<b200gs_gaussian> kept(count ? count : 1);
    uint64_t n = 0;
    int rc = b200gs_apply_edits_for_export(gaussians, count, edits_or_null, mask_words_or_null, kept.data(), &n);
    if (rc != B200GS_OK) return rc;
    std::vector<b200gs_ply_gaussian> verts(n ? n : 1);
    rc = b200gs_gaussian_to_ply(kept.data(), n, verts.data());
    if (rc != B200GS_OK) return rc;
    return b200gs_ply_write(path, verts.data(), n);
}

extern "C" int b200gs_ply_write(const char* path, const b200gs_ply_gaussian* verts, uint64_t count) {
    if (!path || (!verts && count)) { gs_set_error("ply_write: null argument"); return B200GS_ERR_INVALID; }
    FILE* fp = fopen(path, "wb");
    if (!fp) { gs_set_error("ply_write: cannot open '%s'", path); return B200GS_ERR_IO; }
    fprintf(fp, "ply\nformat binary_little_endian 1.0\nelement vertex %llu\n", (unsigned long long)count);
    static const char* names[] = {"x", "y", "z", "nx", "ny", "nz", "f_dc_0", "f_dc_1", "f_dc_2"};
    for (auto n : names) fprintf(fp, "property float %s\n", n);
    for (int k = 0; k < 45; k++) fprintf(fp, "property float f_rest_%d\n", k);
    fprintf(fp, "property float opacity\n");
    for (int k = 0; k < 3; k++) fprintf(fp, "property float scale_%d\n", k);
    for (int k = 0; k < 4; k++) fprintf(fp, "property float rot_%d\n", k);
    fprintf(fp, "end_header\n");
    size_t wrote = count ? fwrite(verts, sizeof(b200gs_ply_gaussian), count, fp) : 0;
    int bad = (wrote != count) | (fclose(fp) != 0);
    if (bad) { gs_set_error("ply_write: short write to '%s'", path); return B200GS_ERR_IO; }
    return B200GS_OK;
}
