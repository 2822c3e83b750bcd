// host_mirror_test.cpp — exercises the C++ host mirror (include/ptb200.hpp) the way the reference's
// renderer/src/main.rs:39-42,118,122 uses the crate, and checks the result against the CPU oracle.
//   host_mirror_test --cpu   : no GPU expected — the product must fail loudly (no CPU fallback)
//   host_mirror_test --gpu   : render through Tracer::render / render_spp, compare with the oracle
#include <array>
#include <cstdio>
#include <cstring>

#include "ptb200.hpp"
#include "../../oracle/pt_oracle.hpp"   // test infrastructure: the checker

using namespace rust_pathtracer;

struct HostOnlyScene : Scene {   // a scene that does not export itself
    Pinhole cam; AnalyticalLight l = AnalyticalLight::spherical(F3(0, 5, 0), 1, F3(1, 1, 1));
    Camera3D& camera() override { return cam; }
    size_t number_of_lights() const override { return 1; }
    const AnalyticalLight& light_at(size_t) const override { return l; }
};

#define CHECK(cond) do { if (!(cond)) { std::printf("FAILED: %s (line %d)\n", #cond, __LINE__); return 1; } } while (0)

static int test_cpu() {
    ColorBuffer buffer = ColorBuffer::new_(5, 3);                      // buffer.rs:18-26
    CHECK(buffer.width == 5 && buffer.height == 3 && buffer.frames == 0 && buffer.pixels.size() == 60);
    buffer.pixels[(2 * 5 + 4) * 4 + 1] = 7;
    CHECK(buffer.at(4, 2)[1] == 7);
    Material m = Material::new_();                                     // material.rs:82-114
    CHECK(m.rgb.x == F(1.5) && m.roughness == F(0.5) && m.ior == F(1.45) && m.metallic == 0);
    Pinhole p = Pinhole::new_();                                       // pinhole.rs:14-25
    CHECK(p.origin.z == 3 && p.fov == 80);
    CHECK(std::fabs(AnalyticalLight::spherical(F3(3, 2, 2), 1, F3(3, 3, 3)).light.area - F(12.566371)) < 1e-5);
    // the two light kinds and the Medium the reference declares (globals.rs:69-84, material.rs:5-34): carried by the mirror
    AnalyticalLight quad = AnalyticalLight::rectangular(F3(-1, 3, 0), F3(2, 0, 0), F3(0, 0, 1.5), F3(9, 8, 7));
    CHECK(quad.light.light_type == PTB_LIGHT_RECTANGULAR && std::fabs(quad.light.area - F(3)) < 1e-6 && quad.light.u.x == 2);
    CHECK(AnalyticalLight::distant(F3(0, 1, 0), F3(1, 1, 1)).light.area == 0);
    CHECK(m.medium.medium_type == PTB_MEDIUM_NONE && m.medium.density == 0 && Medium::new_().anisotropy == 0);
    AnalyticalScene demo;
    auto e = demo.device_export();
    CHECK(e && e->spheres.size() == 2 && e->planes.size() == 1 && e->materials.size() == 3 && e->depth == 4);
    bool threw = false;
    try { Tracer t(std::make_unique<HostOnlyScene>()); } catch (const std::runtime_error& ex) { threw = std::strstr(ex.what(), "no CPU fallback") != nullptr; }
    CHECK(threw);                                                      // a scene without device_export() is rejected
    if (ptb_device_count() == 0) {
        threw = false;
        try { Tracer t(std::make_unique<AnalyticalScene>()); } catch (const std::runtime_error& ex) { threw = std::strstr(ex.what(), "no CPU fallback") != nullptr; }
        CHECK(threw);                                                  // no GPU: loud failure, never a silent CPU render
    }
    threw = false;
    try { buffer.convert_to_u8(nullptr); } catch (const std::runtime_error&) { threw = true; }
    CHECK(threw);
    std::printf("cpu checks passed\n");
    return 0;
}

static int test_gpu() {
    const size_t W = 160, H = 120, N = 4;
    // --- the reference's usage, main.rs:39-42 + 118 + 122 ---
    ColorBuffer buffer = ColorBuffer::new_(W, H);
    auto pt = Tracer::new_(std::make_unique<AnalyticalScene>());
    for (size_t i = 0; i < N; ++i) pt->render(buffer);
    std::vector<uint8_t> frame(W * H * 4);
    buffer.convert_to_u8(frame.data());
    CHECK(buffer.frames == N);
    // --- the oracle on the same samples ---
    pto::AnalyticalSceneLiteral<F> oscene;
    pto::Tracer<F> otr(&oscene);
    pto::ColorBuffer<F> obuf(W, H);
    for (size_t i = 0; i < N; ++i) otr.render(obuf);
    size_t close = 0;
    for (size_t px = 0; px < W * H; ++px) {
        double worst = 0, scale = 1e-3;
        for (int c = 0; c < 3; ++c) {
            worst = std::fmax(worst, std::fabs((double)buffer.pixels[px * 4 + c] - obuf.pixels[px * 4 + c]));
            scale = std::fmax(scale, std::fabs((double)obuf.pixels[px * 4 + c]));
        }
        if (worst / scale < 1e-4) ++close;
        CHECK(buffer.pixels[px * 4 + 3] == F(1));
    }
    std::printf("pixels within 1e-4 rel of the oracle: %.5f\n", (double)close / (W * H));
    CHECK(close >= (size_t)(0.99 * W * H));
    std::vector<uint8_t> oframe(W * H * 4);
    pto::convert_to_u8<F>(buffer.pixels.data(), W * H, oframe.data());
    size_t off = 0;
    for (size_t i = 0; i < frame.size(); ++i) {
        int d = (int)frame[i] - (int)oframe[i];
        CHECK(d >= -1 && d <= 1);
        off += d != 0;
    }
    CHECK(off < frame.size() / 500);
    CHECK(buffer.to_u8_vec() == frame);
    // reset idiom (frames = 0) and the batch extension
    buffer.frames = 0;
    pt->render_spp(buffer, (uint32_t)N);
    CHECK(buffer.frames == N);
    close = 0;
    for (size_t i = 0; i < W * H * 4; ++i) close += std::fabs((double)buffer.pixels[i] - obuf.pixels[i]) <= 1e-4 * std::fmax(1e-3, std::fabs((double)obuf.pixels[i]));
    CHECK(close >= (size_t)(0.99 * W * H * 4));
    // convert_to_u8_at: strict bounds, no gamma (buffer.rs:67-102)
    const size_t FW = 200, FH = 150;
    std::vector<uint8_t> big(FW * FH * 4, 9), obig(FW * FH * 4, 9);
    buffer.convert_to_u8_at(big.data(), 10, 7, FW, FH);
    pto::convert_to_u8_at<F>(buffer.pixels.data(), W, H, obig.data(), 10, 7, FW, FH);
    CHECK(big == obig);
    std::printf("gpu checks passed\n");
    return 0;
}

int main(int argc, char** argv) {
    if (argc > 1 && !std::strcmp(argv[1], "--gpu")) return test_gpu();
    return test_cpu();
}
