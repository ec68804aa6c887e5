// Compile-and-run check of host/gs.hpp (the C++ mirror of the reference's gs:: API) over libb200gs.so.
// usage: gs_hpp_check <tmpdir> [gpu]
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>

#include "../wgpu-3dgs-viewer-app_b200/host/gs.hpp"

#define EXPECT(c) do { if (!(c)) { std::printf("FAIL %s:%d %s\n", __FILE__, __LINE__, #c); return 1; } } while (0)

struct Orbit : gs::CameraTrait {  // CameraOrbitControl of src/app.rs:1174-1244
    gs::Vec3 target{0, 0, 0}, pos{0, 0, 5};
    gs::Mat4 view() const override { return gs::look_at_rh(pos, target, {0, 1, 0}); }
    gs::Mat4 projection(float a) const override { return gs::perspective_rh(1.0471975512f, a, 0.1f, 1e4f); }
};

int main(int argc, char** argv) {
    const std::string dir = argc > 1 ? argv[1] : "/tmp";
    const bool gpu = argc > 2 && std::strcmp(argv[2], "gpu") == 0;
    using G = gs::GaussianPodWithShNorm8Cov3dHalfConfigs;
    static_assert(G::bytes == 76, "default layout is 76 bytes");
    static_assert(gs::GaussianPodWithShSingleCov3dSingleConfigs::bytes == 220, "largest layout");
    EXPECT(b200gs_record_bytes(G::ShConfig::id, G::Cov3dConfig::id) == G::bytes);
    EXPECT(!gs::GaussianShDegree::new_(4).has_value() && gs::GaussianShDegree::new_(3)->degree() == 3);

    // PLY: write, stream back, convert (GaussianSplattingModel::init_load, src/app.rs:1053-1096)
    const uint64_t n = 5000;
    std::vector<gs::PlyGaussianPod> ply(n);
    gs::check(b200gs_synth_scene(0xB2000001, 0, n, ply.data()));
    gs::Gaussians gsn;
    gsn.gaussians = gs::Gaussians::from_ply(ply.data(), n);
    const std::string path = dir + "/gs_hpp_check.ply";
    gsn.write_ply(path);
    auto header = gs::Gaussians::read_ply_header(path);
    EXPECT(header.count() == n);
    gs::Gaussians back;
    gs::Gaussians::read_ply_gaussians(header, [&](const gs::PlyGaussianPod* p, uint64_t k) {
        auto g = gs::Gaussians::from_ply(p, k);
        back.gaussians.insert(back.gaussians.end(), g.begin(), g.end());
    }, 777);
    EXPECT(back.gaussians.size() == n);
    EXPECT(std::memcmp(back.gaussians[123].pos, gsn.gaussians[123].pos, 12) == 0);
    bool threw = false;
    try { gs::Gaussians::read_ply_header(dir + "/does_not_exist.ply"); } catch (const gs::Error& e) { threw = e.is_io(); }
    EXPECT(threw);

    if (!gpu) {
        bool no_gpu = false;
        try { auto v = gs::MultiModelViewer<G>::new_with(0, {64, 64}); } catch (const gs::Error& e) { no_gpu = e.code == B200GS_ERR_CUDA; }
        std::printf("host-only ok (viewer creation %s)\n", no_gpu ? "correctly refused: no GPU" : "succeeded");
        return 0;
    }
    // one frame the way Scene::loaded drives it (src/tab/scene.rs:699-874, 2263-2327)
    auto viewer = gs::MultiModelViewer<G>::new_with(0, {320, 180});
    auto& model = viewer.insert_model("model", n);
    model.update_range(0, gsn.gaussians.data(), n);
    Orbit cam;
    viewer.update_camera(cam, {320, 180});
    viewer.update_model_transform("model", {0, 0, 0}, {0, 0, 0, 1}, {1, 1, 1});
    viewer.update_gaussian_transform(1.0f, gs::GaussianDisplayMode::Splat, gs::GaussianShDegree::new_unchecked(3), false);
    viewer.preprocessor.preprocess(model);
    viewer.radix_sorter.sort(model);
    viewer.poll_wait();
    std::vector<uint8_t> img(320 * 180 * 4);
    viewer.render_frame_host({"model"}, cam, {320, 180}, img.data());
    uint64_t vis = 0, lit = 0;
    gs::check(b200gs_model_visible_count(model.handle(), &vis));
    for (size_t i = 3; i < img.size(); i += 4) lit += img[i] != 0;
    EXPECT(vis > 0 && lit > 0);
    EXPECT(model.download_mask().size() == (n + 31) / 32);
    viewer.remove_model("model");
    EXPECT(viewer.models.empty());
    std::printf("gpu ok: visible %llu, lit pixels %llu\n", (unsigned long long)vis, (unsigned long long)lit);
    return 0;
}
