// TEST INFRASTRUCTURE.  Drives the C++ host mirror (include/equilibrium.hpp) the way the reference's own callers and
// tests drive `Fluid`, and compares with the CPU oracle (oracle/fluid_ref.h) bit for bit.  Linked against either the real
// CUDA library (pytest -m gpu) or the emulated build of the same sources (CPU tests); tests/test_zz_host_mirrors.py builds it.
#include <chrono>
#include <cstdio>
#include <cstring>
#include <random>

#include "equilibrium.hpp"
#include "fluid_ref.h"

using namespace equilibrium;

static int failures = 0;
#define EXPECT(cond, ...)                                  \
    do {                                                   \
        if (!(cond)) {                                     \
            ++failures;                                    \
            std::printf("FAIL %s:%d: ", __FILE__, __LINE__); \
            std::printf(__VA_ARGS__);                      \
            std::printf("\n");                             \
        }                                                  \
    } while (0)

static size_t diff_bits(const std::vector<float> &a, const float *b) {
    size_t n = 0;
    for (size_t i = 0; i < a.size(); ++i) {
        uint32_t x, y;
        std::memcpy(&x, &a[i], 4);
        std::memcpy(&y, &b[i], 4);
        const bool nan_both = (a[i] != a[i]) && (b[i] != b[i]);   // NaN payloads are not compared (tests/parity.py)
        n += (x != y && !nan_both);
    }
    return n;
}

static void compare(Fluid &dev, ref_fluid *ref, const char *where) {
    dev.refresh();
    const float *rd = static_cast<const float *>(ref_fluid_field(ref, REF_F_DENSITY));
    const float *rx = static_cast<const float *>(ref_fluid_field(ref, REF_F_VX));
    const float *ry = static_cast<const float *>(ref_fluid_field(ref, REF_F_VY));
    const uint8_t *rc = static_cast<const uint8_t *>(ref_fluid_field(ref, REF_F_CELLS));
    EXPECT(diff_bits(dev.density, rd) == 0, "%s: density differs in %zu cells", where, diff_bits(dev.density, rd));
    EXPECT(diff_bits(dev.velocities_x, rx) == 0, "%s: velocities_x differs in %zu cells", where, diff_bits(dev.velocities_x, rx));
    EXPECT(diff_bits(dev.velocities_y, ry) == 0, "%s: velocities_y differs in %zu cells", where, diff_bits(dev.velocities_y, ry));
    size_t bad = 0;
    for (size_t i = 0; i < dev.cells_type.size(); ++i) bad += static_cast<uint8_t>(dev.cells_type[i]) != rc[i];
    EXPECT(bad == 0, "%s: cells_type differs in %zu cells", where, bad);
    for (EqField f : {EQ_F_VX0, EQ_F_VY0, EQ_F_SCRATCH}) {   // the private arrays too
        std::vector<float> v;
        dev.download(f, v);
        EXPECT(diff_bits(v, static_cast<const float *>(ref_fluid_field(ref, f))) == 0, "%s: field %d differs", where, (int)f);
    }
}

// CurrentSimulation::simulate (renderer_helpers.rs:52-72) on the default scene, with scripted impulses for add_noise
static void default_scene_frames() {
    FluidConfigs fc;
    SimulationConfigs sc;                        // 0.02, 16, 128  (configs.rs:14-22)
    sc.frames = 2;                               // frames is also the GS iteration count (fluid.rs:445)
    Fluid fluid = Fluid::new_(fc, sc);
    ref_fluid *ref = ref_fluid_new(sc.size, sc.size, sc.delta_t, sc.frames, 0, fc.diffusion, fc.viscousity);
    Rectangle rect;                              // (80,80)-(110,110)  (obstacle.rs:47-51)
    fluid.fill_obstacle(rect);                   // mark_fluid_obstacles (renderer_helpers.rs:76-80)
    ref_fill_rect(ref, 80, 80, 110, 110);
    compare(fluid, ref, "after construction");
    std::mt19937 rng(0);
    std::uniform_real_distribution<float> u(-256.f, 256.f);
    for (int64_t i = 0; i < 2; ++i) {
        const float ax = u(rng), ay = u(rng);
        fluid.add_velocity(sc.size / 2, sc.size / 2, ax, ay);
        ref_add_velocity(ref, sc.size / 2, sc.size / 2, ax, ay);
        fluid.step();
        ref_fluid_step(ref);
        Fluid copy = fluid.clone();              // what the reference sends to the render thread every frame (:61-65)
        compare(copy, ref, "clone of a frame");
    }
    compare(fluid, ref, "default scene, 2 frames");

    // the pub config structs are live (the GUI edits them between runs)
    fluid.simulation_configs.delta_t = 0.05f;
    fluid.fluid_configs.viscousity = 0.01f;
    fluid.fluid_configs.diffusion = 1e-4f;
    ref->delta_t = 0.05f;
    ref->viscosity = 0.01f;
    ref->diffusion = 1e-4f;
    fluid.step();
    ref_fluid_step(ref);
    compare(fluid, ref, "after editing the configs");
    ref_fluid_free(ref);
}

// renderer_helpers.rs:222-252: the default scene renders 1408 obstacle pixels (508 frame cells + the 30x30 rectangle)
static void default_scene_obstacle_pixels() {
    Fluid fluid = Fluid::new_(FluidConfigs(), SimulationConfigs());
    Rectangle rect;
    fluid.fill_obstacle(rect);
    std::vector<uint8_t> px;
    fluid.render_rgba(px);
    size_t red = 0;
    for (size_t i = 0; i < px.size(); i += 4) red += px[i] == 255 && px[i + 1] == 0 && px[i + 2] == 0 && px[i + 3] == 255;
    EXPECT(red == 1408, "obstacle pixels: %zu, the reference's test expects 1408", red);
    fluid.refresh();
    size_t walls = 0;
    for (ContainerWall c : fluid.cells_type) walls += c == ContainerWall::DefaultWall;
    EXPECT(walls == 1408, "wall cells: %zu", walls);
}

// Default::default double-initialises (fluid.rs:83-89)
static void default_double_init() {
    Fluid fluid;
    ref_fluid *ref = ref_fluid_new(128, 128, 0.02f, 16, 0, 0.0f, 0.001f);
    ref_fluid_init(ref);
    compare(fluid, ref, "Default::default");
    ref_fluid_free(ref);
}

// obstacle.rs:100-115: invalid rectangles panic; fluid.rs:626-635: idx is row-major
static void reference_unit_tests() {
    bool threw = false;
    try { Rectangle r({50, 120}, {127, 110}, 128); } catch (const std::invalid_argument &) { threw = true; }
    EXPECT(threw, "Rectangle (50,120)-(127,110) must panic");
    threw = false;
    try { Rectangle r({12, 12}, {10, 10}, 128); } catch (const std::invalid_argument &) { threw = true; }
    EXPECT(threw, "Rectangle (12,12)-(10,10) must panic");
    threw = false;
    try { Rectangle r({1, 1}, {128, 5}, 128); } catch (const std::invalid_argument &) { threw = true; }
    EXPECT(threw, "a point on the container size must panic");
    Rectangle ok({0, 0}, {127, 127}, 128);
    EXPECT(ok.get_approximate_points().size() == 2, "two approximate points");
    const int64_t n = 128;
    for (int64_t y : {0, 5, 127})
        for (int64_t x : {0, 17, 127}) EXPECT(Fluid::idx(x, y, n) == static_cast<size_t>(x + y * n), "idx(%lld,%lld)", (long long)x, (long long)y);
    EXPECT(Fluid::idx(-3, 200, n) == static_cast<size_t>(0 + 127 * n), "idx clamps");
    SimulationConfigs s;
    FluidConfigs f;
    EXPECT(s.delta_t == 0.02f && s.frames == 16 && s.size == 128, "SimulationConfigs::default");
    EXPECT(f.diffusion == 0.0f && f.viscousity == 0.001f && f.has_perlin_noise, "FluidConfigs::default");
    EXPECT((f.fluid_color == Color32{208, 88, 157, 220}) && (f.world_color == Color32{94, 146, 162, 128}), "default colours");
}

// a GUI-edited rectangle may leave the grid: fill_obstacle clamps like idx! (fluid.rs:610-619, obstacle_widget.rs:176-188)
static void edited_rectangle_clamps() {
    SimulationConfigs sc(0.02f, 2, 64);
    Fluid fluid = Fluid::new_(FluidConfigs(), sc);
    ref_fluid *ref = ref_fluid_new(64, 64, 0.02f, 2, 0, 0.0f, 0.001f);
    Rectangle rect({10, 10}, {20, 20}, 64);
    rect.get_approximate_points()[1] = {90, 30};
    fluid.fill_obstacle(rect);
    ref_fill_rect(ref, 10, 10, 90, 30);
    fluid.step();
    ref_fluid_step(ref);
    compare(fluid, ref, "rectangle edited past the grid");
    ref_fluid_free(ref);
}

// add_noise on the device against the oracle's restatement of the impulse
static void device_noise() {
    SimulationConfigs sc(0.02f, 3, 96);
    DeviceOptions opt;
    opt.noise_seed = 0x1234ABCD5678ull;
    Fluid fluid = Fluid::new_(FluidConfigs(), sc, opt);
    ref_fluid *ref = ref_fluid_new(96, 96, 0.02f, 3, 0, 0.0f, 0.001f);
    for (uint64_t fr = 0; fr < 3; ++fr) {
        const EqNoise nz = fluid.device_noise(fr);
        uint32_t xy[2];
        float a[2];
        ref_noise_impulse(nz.seed, fr, 96, nz.cos_t, nz.sin_t, nz.gain, xy, a);
        ref_add_velocity(ref, xy[0], xy[1], a[0], a[1]);
        ref_fluid_step(ref);
        if (fr < 2) {
            if (fluid.fluid_configs.has_perlin_noise) fluid.add_noise();   // renderer_helpers.rs:55-57
            fluid.step();
        } else {
            fluid.step_n_noise(1);                                         // continues the same stream
        }
    }
    compare(fluid, ref, "device-side add_noise");
    ref_fluid_free(ref);
}

// errors arrive as exceptions carrying eq_last_error, and leave the object usable
static void error_behaviour() {
    bool threw = false;
    try { Fluid f = Fluid::new_(FluidConfigs(), SimulationConfigs(0.02f, 2, 10)); } catch (const Error &e) { threw = e.code == EQ_ERR_INVALID; }
    EXPECT(threw, "size 10 must be rejected (init_density underflows, fluid.rs:534)");
    Fluid fluid = Fluid::new_(FluidConfigs(), SimulationConfigs(0.02f, 2, 64));
    fluid.simulation_configs.size = 128;
    threw = false;
    try { fluid.step(); } catch (const Error &) { threw = true; }
    EXPECT(threw, "a Fluid cannot be resized (renderer.rs:145-149 builds a new one)");
    fluid.simulation_configs.size = 64;
    fluid.step();
    fluid.sync();
}

// CurrentSimulation::simulate (renderer_helpers.rs:52-72): per-frame clones, then the same run over the snapshot path
static void current_simulation() {
    SimulationConfigs sc(0.02f, 3, 64);   // small: this also runs on the host emulator
    DeviceOptions opt;
    opt.noise_seed = 77;
    auto make_ref = [&]() {
        ref_fluid *r = ref_fluid_new(64, 64, 0.02f, 3, 0, 0.0f, 0.001f);
        ref_fill_rect(r, 20, 24, 40, 44);
        return r;
    };
    auto ref_frame = [&](ref_fluid *r, const Fluid &f, uint64_t fr) {
        const EqNoise nz = f.device_noise(fr);
        uint32_t xy[2];
        float a[2];
        ref_noise_impulse(nz.seed, fr, 64, nz.cos_t, nz.sin_t, nz.gain, xy, a);
        ref_add_velocity(r, xy[0], xy[1], a[0], a[1]);
        ref_fluid_step(r);
    };
    {   // the reference's way: a FluidStep with a deep copy per frame
        CurrentSimulation sim(Fluid::new_(FluidConfigs(), sc, opt), {Rectangle({20, 24}, {40, 44}, 64)});
        ref_fluid *ref = make_ref();
        int64_t seen = 0;
        sim.simulate([&](FluidStep step) {
            EXPECT(step.frame_number == seen, "frame order");
            ref_frame(ref, sim.fluid, static_cast<uint64_t>(seen));
            compare(step.fluid, ref, "FluidStep clone");
            ++seen;
        });
        EXPECT(seen == 3, "3 frames sent, got %lld", (long long)seen);
        ref_fluid_free(ref);
    }
    for (int kind : {EQ_SNAP_DENSITY, EQ_SNAP_RGBA}) {   // the snapshot path: one array per frame, overlapped
        CurrentSimulation sim(Fluid::new_(FluidConfigs(), sc, opt), {Rectangle({20, 24}, {40, 44}, 64)});
        ref_fluid *ref = make_ref();
        int64_t seen = 0;
        const FluidConfigs fc;
        std::vector<uint8_t> want(64 * 64 * 4);
        sim.simulate_frames(kind, [&](FrameView fv) {
            EXPECT(fv.frame_number == seen && fv.bytes == want.size(), "frame order / size");
            ref_frame(ref, sim.fluid, static_cast<uint64_t>(seen));
            const float *rd = static_cast<const float *>(ref_fluid_field(ref, REF_F_DENSITY));
            if (kind == EQ_SNAP_DENSITY) {
                std::memcpy(want.data(), rd, want.size());
            } else {
                const uint8_t world[4] = {fc.world_color.r, fc.world_color.g, fc.world_color.b, fc.world_color.a};
                const uint8_t fl[4] = {fc.fluid_color.r, fc.fluid_color.g, fc.fluid_color.b, fc.fluid_color.a};
                const uint8_t obs[4] = {255, 0, 0, 255};
                ref_render_rgba(rd, static_cast<const uint8_t *>(ref_fluid_field(ref, REF_F_CELLS)), 64, 64, world, fl, obs, want.data());
            }
            EXPECT(std::memcmp(fv.data, want.data(), want.size()) == 0, "snapshot kind %d frame %lld differs", kind, (long long)seen);
            ++seen;
        });
        EXPECT(seen == 3, "3 frames delivered, got %lld", (long long)seen);
        ref_fluid_free(ref);
    }
    CurrentSimulation def;   // Default: Fluid::default() + the default rectangle
    EXPECT(def.obstacles.size() == 1 && def.fluid.simulation_configs.frames == 16, "CurrentSimulation::default");
}

int main() {
    if (eq_device_count() < 1) {
        std::printf("no CUDA device: the product path has no CPU fallback\n");
        return 2;
    }
    struct { const char *name; void (*fn)(); } sections[] = {
        {"reference_unit_tests", reference_unit_tests}, {"default_scene_obstacle_pixels", default_scene_obstacle_pixels},
        {"default_double_init", default_double_init},   {"default_scene_frames", default_scene_frames},
        {"edited_rectangle_clamps", edited_rectangle_clamps}, {"device_noise", device_noise},
        {"error_behaviour", error_behaviour},           {"current_simulation", current_simulation}};
    for (auto &sec : sections) {
        const auto t0 = std::chrono::steady_clock::now();
        sec.fn();
        std::printf("%-32s %.2f s\n", sec.name, std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count());
    }
    if (failures) std::printf("%d FAILED\n", failures);
    else std::printf("host mirror ok\n");
    return failures ? 1 : 0;
}
