// gs.hpp — header-only C++ mirror of the `gs::` (crate wgpu-3dgs-viewer) API surface that the
// reference app touches on the hot path, implemented over the C ABI of include/b200gs.h.
//
// The reference's host side is Rust; this image has no Rust toolchain, so the host layer above the
// C ABI is C++ with the same names, argument meaning and error behaviour (Result<_, gs::Error> ->
// exceptions of type gs::Error).  Call sites cited are under /root/reference.  A Rust shim with the
// same shape is sketched in INTEGRATION.md.
#pragma once

#include <array>
#include <cmath>
#include <cstdint>
#include <map>
#include <memory>
#include <optional>
#include <stdexcept>
#include <string>
#include <vector>

#include "../../include/b200gs.h"

namespace gs {

// gs::Error (src/app.rs:548, 1055; Error::Io at src/tab/scene.rs:234)
struct Error : std::runtime_error {
    int code;
    Error(int c, const std::string& what) : std::runtime_error(what), code(c) {}
    bool is_io() const { return code == B200GS_ERR_IO; }
};
inline void check(int rc) {
    if (rc != B200GS_OK) throw Error(rc, b200gs_last_error());
}

using Vec3 = std::array<float, 3>;
using Vec4 = std::array<float, 4>;
using Quat = std::array<float, 4>;  // x, y, z, w (glam)
using Mat4 = std::array<float, 16>; // column-major (glam)
using UVec2 = std::array<uint32_t, 2>;

using Gaussian = b200gs_gaussian;              // src/app.rs:512
using PlyGaussianPod = b200gs_ply_gaussian;    // src/tab/scene.rs:997
using GaussianEditPod = b200gs_edit_pod;       // src/app.rs:1556
using MaskOpShapePod = b200gs_mask_shape;      // src/app.rs:1580
using QueryPod = b200gs_query_pod;

enum class GaussianDisplayMode : uint32_t { Splat = 0, Ellipse = 1, Point = 2 };  // src/tab/transform.rs:129-131
namespace GaussianEditFlag { enum : uint32_t { ENABLED = 1, HIDDEN = 2, OVERRIDE_COLOR = 4 }; }  // src/app.rs:1548-1553

// gs::GaussianShDegree::{new, new_unchecked, degree} (src/app.rs:1161; src/tab/transform.rs:137-139)
class GaussianShDegree {
    uint8_t d_;
    explicit GaussianShDegree(uint8_t d) : d_(d) {}
public:
    static std::optional<GaussianShDegree> new_(uint8_t d) { return d <= 3 ? std::optional<GaussianShDegree>(GaussianShDegree(d)) : std::nullopt; }
    static GaussianShDegree new_unchecked(uint8_t d) { return GaussianShDegree(d); }
    uint8_t degree() const { return d_; }
};

// The 8 GaussianPod layouts (src/app.rs:250-257): GaussianPodWithSh{Single,Half,Norm8,None}Cov3d{Single,Half}Configs
struct GaussianShSingleConfig { static constexpr uint32_t id = B200GS_SH_SINGLE; static constexpr size_t field_bytes = 180; };
struct GaussianShHalfConfig { static constexpr uint32_t id = B200GS_SH_HALF; static constexpr size_t field_bytes = 92; };
struct GaussianShNorm8Config { static constexpr uint32_t id = B200GS_SH_NORM8; static constexpr size_t field_bytes = 48; };
struct GaussianShNoneConfig { static constexpr uint32_t id = B200GS_SH_NONE; static constexpr size_t field_bytes = 0; };
struct GaussianCov3dSingleConfig { static constexpr uint32_t id = B200GS_COV3D_SINGLE; static constexpr size_t field_bytes = 24; };
struct GaussianCov3dHalfConfig { static constexpr uint32_t id = B200GS_COV3D_HALF; static constexpr size_t field_bytes = 12; };
template <class Sh, class Cov>
struct GaussianPodWith {
    using ShConfig = Sh;
    using Cov3dConfig = Cov;
    static constexpr size_t bytes = 16 + Sh::field_bytes + Cov::field_bytes;
};
using GaussianPodWithShSingleCov3dSingleConfigs = GaussianPodWith<GaussianShSingleConfig, GaussianCov3dSingleConfig>;
using GaussianPodWithShSingleCov3dHalfConfigs = GaussianPodWith<GaussianShSingleConfig, GaussianCov3dHalfConfig>;
using GaussianPodWithShHalfCov3dSingleConfigs = GaussianPodWith<GaussianShHalfConfig, GaussianCov3dSingleConfig>;
using GaussianPodWithShHalfCov3dHalfConfigs = GaussianPodWith<GaussianShHalfConfig, GaussianCov3dHalfConfig>;
using GaussianPodWithShNorm8Cov3dSingleConfigs = GaussianPodWith<GaussianShNorm8Config, GaussianCov3dSingleConfig>;
using GaussianPodWithShNorm8Cov3dHalfConfigs = GaussianPodWith<GaussianShNorm8Config, GaussianCov3dHalfConfig>;
using GaussianPodWithShNoneCov3dSingleConfigs = GaussianPodWith<GaussianShNoneConfig, GaussianCov3dSingleConfig>;
using GaussianPodWithShNoneCov3dHalfConfigs = GaussianPodWith<GaussianShNoneConfig, GaussianCov3dHalfConfig>;

// gs::CameraTrait{view, projection} (src/app.rs:1236-1244, 1329-1343)
struct CameraTrait {
    virtual ~CameraTrait() = default;
    virtual Mat4 view() const = 0;
    virtual Mat4 projection(float aspect_ratio) const = 0;
};
inline Mat4 look_at_rh(const Vec3& eye, const Vec3& target, const Vec3& up) {
    Mat4 m;
    b200gs_look_at_rh(eye.data(), target.data(), up.data(), m.data());
    return m;
}
inline Mat4 perspective_rh(float vfov, float aspect, float z_near, float z_far) {
    Mat4 m;
    b200gs_perspective_rh(vfov, aspect, z_near, z_far, m.data());
    return m;
}
// gs::Camera — the first-person camera (src/app.rs:1247, 1299-1316; src/tab/scene.rs:1395-1457)
struct Camera : CameraTrait {
    Vec3 pos{0, 0, 0};
    float yaw = 0, pitch = 0;
    std::array<float, 2> z{0.1f, 1e4f};
    float vertical_fov = 1.0471975512f;
    Camera(std::array<float, 2> z_range, float vfov) : z(z_range), vertical_fov(vfov) {}
    Vec3 get_forward() const { return {std::sin(yaw) * std::cos(pitch), std::sin(pitch), std::cos(yaw) * std::cos(pitch)}; }
    Vec3 get_right() const { return {-std::cos(yaw), 0.0f, std::sin(yaw)}; }
    void yaw_by(float d) { yaw += d; }
    void pitch_by(float d) { pitch = std::fmax(-1.5607963f, std::fmin(1.5607963f, pitch + d)); }
    Mat4 view() const override {
        Vec3 f = get_forward();
        return look_at_rh(pos, {pos[0] + f[0], pos[1] + f[1], pos[2] + f[2]}, {0, 1, 0});
    }
    Mat4 projection(float aspect) const override { return perspective_rh(vertical_fov, aspect, z[0], z[1]); }
};

// gs::Gaussians{gaussians} + read_ply_header / read_ply_gaussians / write_ply (src/app.rs:1029-1031, 1056-1070, 910-914)
struct PlyHeader {
    std::shared_ptr<b200gs_ply_reader> reader;
    uint64_t n = 0;
    uint64_t count() const { return n; }
};
struct Gaussians {
    std::vector<Gaussian> gaussians;

    static PlyHeader read_ply_header(const std::string& path) {
        b200gs_ply_reader* r = nullptr;
        uint64_t n = 0;
        check(b200gs_ply_open(path.c_str(), &r, &n));
        return PlyHeader{std::shared_ptr<b200gs_ply_reader>(r, [](b200gs_ply_reader* p) { b200gs_ply_close(p); }), n};
    }
    // streaming: calls `sink(chunk)` for every chunk of vertices parsed, like the iterator the app drains
    template <class Sink>
    static void read_ply_gaussians(const PlyHeader& h, Sink&& sink, uint64_t chunk = 1u << 16) {
        std::vector<PlyGaussianPod> buf(chunk);
        for (;;) {
            uint64_t got = 0;
            check(b200gs_ply_read(h.reader.get(), buf.data(), chunk, &got));
            if (!got) break;
            sink(buf.data(), got);
        }
    }
    static std::vector<Gaussian> from_ply(const PlyGaussianPod* p, uint64_t n) {  // Gaussian::from(PlyGaussianPod)
        std::vector<Gaussian> out(n);
        check(b200gs_gaussian_from_ply(p, n, out.data()));
        return out;
    }
    // Gaussians::write_ply(writer, Option<&[GaussianEditPod]>, Option<mask words>) (src/app.rs:904-914): nullptr = None
    void write_ply(const std::string& path, const std::vector<b200gs_edit_pod>* edits = nullptr,
                   const std::vector<uint32_t>* mask = nullptr) const {
        if (edits && edits->size() != gaussians.size()) throw Error(B200GS_ERR_INVALID, "write_ply: one edit pod per Gaussian");
        if (mask && mask->size() != (gaussians.size() + 31) / 32) throw Error(B200GS_ERR_INVALID, "write_ply: mask must hold ceil(N/32) words");
        check(b200gs_ply_write_edited(path.c_str(), gaussians.data(), gaussians.size(), edits ? edits->data() : nullptr,
                                      mask ? mask->data() : nullptr));
    }
};

template <class G> class MultiModelViewer;

// gs::MultiModelViewerModel{gaussian_buffers, bind_groups} (src/tab/scene.rs:2135-2138)
template <class G>
class MultiModelViewerModel {
    friend class MultiModelViewer<G>;
    b200gs_model* h_ = nullptr;
public:
    b200gs_model* handle() const { return h_; }
    size_t len() const { return (size_t)b200gs_model_len(h_); }  // gaussians_buffer.len()
    // gaussians_buffer.update_range(queue, start, &[Gaussian]) — src/tab/scene.rs:2076-2084
    void update_range(size_t start, const Gaussian* g, size_t n) { check(b200gs_model_update_range(h_, start, g, n)); }
    void upload_mask(const std::vector<uint32_t>& w) { check(b200gs_model_upload_mask(h_, w.data(), w.size())); }
    void upload_selection(const std::vector<uint32_t>& w) { check(b200gs_model_upload_selection(h_, w.data(), w.size())); }
    std::vector<uint32_t> download_mask() {  // mask_buffer.download() — src/app.rs:806
        std::vector<uint32_t> w((len() + 31) / 32);
        uint64_t n = 0;
        check(b200gs_model_download_mask(h_, w.data(), w.size(), &n));
        return w;
    }
    std::vector<GaussianEditPod> download_edits() {  // gaussians_edit_buffer.download() — src/app.rs:789
        std::vector<GaussianEditPod> e(len());
        uint64_t n = 0;
        check(b200gs_model_download_edits(h_, e.data(), e.size(), &n));
        return e;
    }
};

// gs::MultiModelViewer<G> (src/tab/scene.rs:1969-1980) with the stage objects the app calls through
template <class G>
class MultiModelViewer {
    b200gs_viewer* h_ = nullptr;
public:
    std::map<std::string, MultiModelViewerModel<G>> models;

    struct Preprocessor {  // viewer.preprocessor.preprocess(encoder, bind_group, n) — scene.rs:856-863
        void preprocess(MultiModelViewerModel<G>& m, bool unedited = false) const { check(b200gs_model_preprocess(m.handle(), unedited)); }
    } preprocessor;
    struct RadixSorter {   // viewer.radix_sorter.sort(encoder, bind_group, args) — scene.rs:865-869
        void sort(MultiModelViewerModel<G>& m) const { check(b200gs_model_sort(m.handle())); }
    } radix_sorter;
    struct Postprocessor { // viewer.postprocessor.postprocess(...) — scene.rs:604-610
        void postprocess(MultiModelViewerModel<G>& m) const { check(b200gs_model_postprocess(m.handle())); }
    } postprocessor;

    // MultiModelViewer::new_with(device, format, depth_stencil, size)
    static MultiModelViewer new_with(int device, UVec2 size) {
        MultiModelViewer v;
        check(b200gs_viewer_create(device, G::ShConfig::id, G::Cov3dConfig::id, size[0], size[1], &v.h_));
        return v;
    }
    MultiModelViewer() = default;
    MultiModelViewer(MultiModelViewer&& o) noexcept : h_(o.h_), models(std::move(o.models)) { o.h_ = nullptr; }
    MultiModelViewer& operator=(MultiModelViewer&& o) noexcept {
        if (this != &o) { if (h_) b200gs_viewer_destroy(h_); h_ = o.h_; models = std::move(o.models); o.h_ = nullptr; }
        return *this;
    }
    MultiModelViewer(const MultiModelViewer&) = delete;
    ~MultiModelViewer() { if (h_) b200gs_viewer_destroy(h_); }

    // MultiModelViewerGaussianBuffers::new_empty + BindGroups::new + models.insert — scene.rs:2111-2139
    MultiModelViewerModel<G>& insert_model(const std::string& key, size_t count) {
        MultiModelViewerModel<G> m;
        check(b200gs_model_create(h_, key.c_str(), count, &m.h_));
        return models[key] = m;
    }
    void remove_model(const std::string& key) {  // scene.rs:2176
        auto it = models.find(key);
        if (it == models.end()) return;
        check(b200gs_model_destroy(h_, it->second.handle()));
        models.erase(it);
    }
    void update_query_texture_size(UVec2 s) { check(b200gs_resize(h_, s[0], s[1])); }                  // scene.rs:740
    void update_query(const QueryPod& q) { check(b200gs_set_query(h_, &q)); }                          // scene.rs:785
    void update_camera(const CameraTrait& c, UVec2 size) {                                             // scene.rs:795
        Mat4 v = c.view(), p = c.projection((float)size[0] / (float)size[1]);
        float s[2] = {(float)size[0], (float)size[1]};
        check(b200gs_set_camera(h_, v.data(), p.data(), s));
    }
    void update_model_transform(const std::string& key, Vec3 pos, Quat quat, Vec3 scale) {             // scene.rs:796-802
        check(b200gs_model_set_transform(models.at(key).handle(), pos.data(), quat.data(), scale.data()));
    }
    void update_gaussian_transform(float size, GaussianDisplayMode mode, GaussianShDegree deg, bool no_sh0) {  // scene.rs:803-809
        check(b200gs_set_gaussian_transform(h_, size, (uint32_t)mode, deg.degree(), no_sh0));
    }
    void update_selection_edit_with_pod(const GaussianEditPod& e) { check(b200gs_set_selection_edit(h_, &e)); }   // scene.rs:815
    void update_selection_highlight(Vec4 rgba) { check(b200gs_set_selection_highlight(h_, rgba.data())); }       // scene.rs:816
    // renderer.render_with_pass over model_render_keys (farthest first) — scene.rs:2302-2314
    void render(const std::vector<std::string>& model_render_keys, void* rgba8_device, size_t pitch) {
        std::vector<b200gs_model*> hs;
        for (auto& k : model_render_keys) hs.push_back(models.at(k).handle());
        check(b200gs_render(h_, hs.data(), (uint32_t)hs.size(), rgba8_device, pitch));
    }
    void render_frame_host(const std::vector<std::string>& model_render_keys, const CameraTrait& c, UVec2 size, void* rgba8_host) {
        std::vector<b200gs_model*> hs;
        for (auto& k : model_render_keys) hs.push_back(models.at(k).handle());
        update_camera(c, size);
        check(b200gs_render_frame_host(h_, hs.data(), (uint32_t)hs.size(), nullptr, nullptr, rgba8_host));
    }
    void poll_wait() { check(b200gs_sync(h_)); }  // queue.submit + device.poll(Maintain::Wait) — scene.rs:872-873
    b200gs_viewer* handle() const { return h_; }
};

}  // namespace gs
