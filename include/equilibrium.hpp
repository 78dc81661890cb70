// equilibrium.hpp -- C++17 host-side mirror of the reference's `simulation` module, above the C ABI of
// equilibrium_cuda.h.  The reference is Rust and no Rust toolchain exists in the build image, so this header is the
// compiled-language host side that is actually built and tested here (tests/cpp/host_mirror_test.cpp); the Rust crates
// under rust/ bind the same entry points.  Same names, argument meaning and error behaviour as the reference:
//
//   SimulationConfigs, FluidConfigs      src/simulation/configs.rs:5-60
//   ContainerWall                        src/simulation/fluid.rs:11-17
//   Rectangle, ObstaclesType             src/simulation/obstacle.rs:4-95   (a panic becomes std::invalid_argument)
//   Fluid::new / Default / step / add_noise / fill_obstacle / Clone / pub fields    src/simulation/fluid.rs:51-110,
//                                        437-524, 575-599, 610-619
//
// Header-only; link with libequilibrium_cuda.so.  There is no CPU fallback: without the library the program does not
// link, without a device the constructor throws.
#ifndef EQUILIBRIUM_HPP
#define EQUILIBRIUM_HPP

#include <cmath>
#include <cstdint>
#include <stdexcept>
#include <string>
#include <utility>
#include <vector>

#include "equilibrium_cuda.h"

namespace equilibrium {

// egui's Color32 (r, g, b, a bytes), as far as configs.rs:50-60 and renderer_helpers.rs:122-143 use it
struct Color32 {
    uint8_t r = 0, g = 0, b = 0, a = 255;
    static constexpr Color32 from_rgba_premultiplied(uint8_t r, uint8_t g, uint8_t b, uint8_t a) { return Color32{r, g, b, a}; }
    bool operator==(const Color32 &o) const { return r == o.r && g == o.g && b == o.b && a == o.a; }
};
inline constexpr Color32 RED{255, 0, 0, 255};   // RenderingListener's obstacles_color (renderer_helpers.rs:94-101)

// configs.rs:5-33
struct SimulationConfigs {
    float delta_t = 0.02f;
    int64_t frames = 16;
    uint32_t size = 128;
    SimulationConfigs() = default;
    SimulationConfigs(float delta_t_, int64_t frames_, uint32_t fluid_container_size) : delta_t(delta_t_), frames(frames_), size(fluid_container_size) {}
    static SimulationConfigs new_(float delta_t, int64_t frames, uint32_t fluid_container_size) { return {delta_t, frames, fluid_container_size}; }
};

// configs.rs:36-60 (`viscousity` is the reference's spelling)
struct FluidConfigs {
    float diffusion = 0.0f;
    float viscousity = 0.001f;
    bool has_perlin_noise = true;
    Color32 fluid_color = Color32::from_rgba_premultiplied(208, 88, 157, 220);
    Color32 world_color = Color32::from_rgba_premultiplied(94, 146, 162, 128);
};

// fluid.rs:11-17; the numbering is the ABI's (0 = NoWall, 1 = DefaultWall), see INTEGRATION.md
enum class ContainerWall : uint8_t { NoWall = 0, DefaultWall = 1 };

using Point = std::pair<int64_t, int64_t>;   // line_drawing::Point<i64>

// obstacle.rs:29-95
class Rectangle {
public:
    Point down_left_point, up_right_point;
    // Rectangle::new (obstacle.rs:55-71): panics on invalid input
    Rectangle(Point down_left, Point up_right, uint32_t fluid_container_size)
        : down_left_point(down_left), up_right_point(up_right), approximate_points_{down_left, up_right} {
        if (!are_all_points_valid(static_cast<int64_t>(fluid_container_size))) throw std::invalid_argument("Invalid input for Rectangle");
    }
    // Default (obstacle.rs:47-51)
    Rectangle() : Rectangle({80, 80}, {110, 110}, SimulationConfigs().size) {}
    // obstacle.rs:74-87
    bool are_all_points_valid(int64_t fluid_container_size) const {
        const int64_t v[4] = {down_left_point.first, down_left_point.second, up_right_point.first, up_right_point.second};
        bool all_below = true;
        for (int64_t e : v) all_below = all_below && e < fluid_container_size;
        return v[0] != v[2] && v[1] != v[3] && v[0] < v[2] && v[1] < v[3] && all_below;
    }
    // trait Obstacle (obstacle.rs:4-7, 90-94): mutable, as the GUI edits the points in place (obstacle_widget.rs:176-188)
    std::vector<Point> &get_approximate_points() { return approximate_points_; }

private:
    std::vector<Point> approximate_points_;
};
using ObstaclesType = Rectangle;   // obstacle.rs:12-15: an enum with the single variant Rectangle(Rectangle)

// what a panic of the CUDA path becomes: the status code and eq_last_error()
class Error : public std::runtime_error {
public:
    int code;
    Error(int code_, const std::string &msg) : std::runtime_error("equilibrium_cuda error " + std::to_string(code_) + ": " + msg), code(code_) {}
};
inline void check(int rc) {
    if (rc != EQ_OK) throw Error(rc, eq_last_error());
}

// what the CUDA path adds to Fluid::new's two arguments
struct DeviceOptions {
    EqMode mode = EQ_MODE_EXACT;   // lexicographic wavefront (bit-identical to fluid.rs) or red-black
    int64_t gs_iterations = 0;     // 0 => `frames`, like the reference (fluid.rs:445)
    int device = 0;
    uint64_t noise_seed = 0;       // add_noise draws from Philox4x32-10 keyed by this instead of thread_rng
};

// fluid.rs:51-81
class Fluid {
public:
    // the reference's pub fields.  The config structs are live: edits are pushed before the next step.  The three
    // arrays and the mask are host mirrors of the device state, refreshed by refresh() (one download each).
    FluidConfigs fluid_configs;
    SimulationConfigs simulation_configs;
    std::vector<float> density, velocities_x, velocities_y;
    std::vector<ContainerWall> cells_type;

    // Fluid::new (fluid.rs:93-110)
    Fluid(FluidConfigs init_fluid, SimulationConfigs init_simulation, DeviceOptions opt = {})
        : fluid_configs(init_fluid), simulation_configs(init_simulation), opt_(opt) {
        EqParams p = params();
        check(eq_create(&p, &h_));
        pushed_ = p;
    }
    static Fluid new_(FluidConfigs init_fluid, SimulationConfigs init_simulation, DeviceOptions opt = {}) { return Fluid(init_fluid, init_simulation, opt); }
    // Default::default (fluid.rs:83-89): new() and then init() a second time
    Fluid() : Fluid(FluidConfigs(), SimulationConfigs()) { check(eq_init_default(h_)); }
    // #[derive(Clone)] (fluid.rs:51): a deep copy of the device state
    Fluid(const Fluid &o)
        : fluid_configs(o.fluid_configs), simulation_configs(o.simulation_configs), density(o.density), velocities_x(o.velocities_x),
          velocities_y(o.velocities_y), cells_type(o.cells_type), opt_(o.opt_), pushed_(o.pushed_), noise_frame_(o.noise_frame_) {
        check(eq_clone(o.h_, &h_));
    }
    Fluid clone() const { return Fluid(*this); }
    Fluid(Fluid &&o) noexcept { *this = std::move(o); }
    Fluid &operator=(Fluid &&o) noexcept {
        if (this != &o) {
            if (h_) eq_destroy(h_);
            h_ = o.h_;
            o.h_ = nullptr;
            fluid_configs = o.fluid_configs;
            simulation_configs = o.simulation_configs;
            density = std::move(o.density);
            velocities_x = std::move(o.velocities_x);
            velocities_y = std::move(o.velocities_y);
            cells_type = std::move(o.cells_type);
            opt_ = o.opt_;
            pushed_ = o.pushed_;
            noise_frame_ = o.noise_frame_;
        }
        return *this;
    }
    Fluid &operator=(const Fluid &o) { return *this = Fluid(o); }
    ~Fluid() {
        if (h_) eq_destroy(h_);   // Drop
    }

    // Fluid::step (fluid.rs:437-524); enqueued, reading a field synchronises
    void step() {
        push_params();
        check(eq_step(h_));
    }
    // the loop of CurrentSimulation::simulate (renderer_helpers.rs:54-60) without returning to the host
    void step_n(int64_t n, const std::vector<EqSource> &sources = {}) {
        push_params();
        check(eq_step_n(h_, n, sources.empty() ? nullptr : sources.data(), static_cast<int64_t>(sources.size())));
    }
    // Fluid::add_noise (fluid.rs:575-599) on the device.  thread_rng -> a seeded Philox stream (one counter per call);
    // the angle is a function of delta_t only (:578-583) -- noise-0.7's Perlin is not available, noise_angle() is the
    // same stand-in the Python mirror uses.
    void add_noise() {
        EqNoise nz = device_noise(noise_frame_++);
        check(eq_add_noise(h_, &nz));
    }
    // n x { add_noise(); step() } in one call
    void step_n_noise(int64_t n) {
        push_params();
        EqNoise nz = device_noise(noise_frame_);
        check(eq_step_n_noise(h_, n, &nz));
        noise_frame_ += static_cast<uint64_t>(n);
    }
    float noise_angle() const {
        const double dt = simulation_configs.delta_t;
        return static_cast<float>(std::sin(12.9898 * dt + 78.233 * dt) * 6.28 * 2.0);
    }
    EqNoise device_noise(uint64_t frame) const {
        const double th = static_cast<double>(noise_angle()) * (3.14159265358979323846 / 180.0);   // geo rotates by degrees
        EqNoise nz{};
        nz.seed = opt_.noise_seed;
        nz.first_frame = frame;
        nz.cos_t = static_cast<float>(std::cos(th));
        nz.sin_t = static_cast<float>(std::sin(th));
        nz.gain = 2.0f;   // fluid.rs:595-596
        return nz;
    }
    // Fluid::fill_obstacle (fluid.rs:610-619): [p0.x, p1.x) x [p0.y, p1.y) becomes DefaultWall, indices clamped like idx!
    void fill_obstacle(ObstaclesType &obstacle) {
        const std::vector<Point> &p = obstacle.get_approximate_points();
        check(eq_fill_rect(h_, p[0].first, p[0].second, p[1].first, p[1].second));
    }
    // private in the reference (fluid.rs:120-131), public here because scripted sources need them
    void add_density(uint32_t x, uint32_t y, float amount) { check(eq_add_density(h_, x, y, amount)); }
    void add_velocity(uint32_t x, uint32_t y, float amount_x, float amount_y) { check(eq_add_velocity(h_, x, y, amount_x, amount_y)); }

    // idx! (fluid.rs:31-35)
    static size_t idx(int64_t x, int64_t y, int64_t size) {
        auto cl = [size](int64_t v) { return v < 0 ? 0 : (v > size - 1 ? size - 1 : v); };
        return static_cast<size_t>(cl(x) + cl(y) * size);
    }

    // ---- host mirrors -----------------------------------------------------------------------------------------
    void sync() { check(eq_sync(h_)); }
    size_t cells() const { return static_cast<size_t>(simulation_configs.size) * simulation_configs.size; }
    void download(EqField field, std::vector<float> &out) {
        out.resize(cells());
        check(eq_download(h_, field, out.data(), out.size() * sizeof(float)));
    }
    void upload(EqField field, const std::vector<float> &in) { check(eq_upload(h_, field, in.data(), in.size() * sizeof(float))); }
    // refresh density, velocities_x, velocities_y and cells_type from the device
    void refresh() {
        download(EQ_F_DENSITY, density);
        download(EQ_F_VX, velocities_x);
        download(EQ_F_VY, velocities_y);
        cells_type.resize(cells());
        static_assert(sizeof(ContainerWall) == 1, "cells_type travels as bytes");
        check(eq_download(h_, EQ_F_CELLS, cells_type.data(), cells_type.size()));
    }
    // render_image's pixel loop (renderer_helpers.rs:145-167) on the device: size*size RGBA pixels
    void render_rgba(std::vector<uint8_t> &rgba, Color32 obstacles_color = RED) {
        rgba.resize(cells() * 4);
        EqColors c = colors(obstacles_color);
        check(eq_render_rgba(h_, &c, rgba.data(), rgba.size()));
    }
    // the frame hand-off that replaces `fluid.clone()` + send (renderer_helpers.rs:61-65): asynchronous, into `dst`
    // (pinned memory from eq_host_alloc for a real overlap), complete after snapshot_wait(slot)
    void snapshot_begin(int kind, int slot, void *dst, size_t bytes, Color32 obstacles_color = RED) {
        EqColors c = colors(obstacles_color);
        check(eq_snapshot_begin(h_, kind, slot, &c, dst, bytes));
    }
    void snapshot_wait(int slot) { check(eq_snapshot_wait(h_, slot)); }
    eq_fluid *handle() { return h_; }

private:
    EqParams params() const {
        EqParams p{};
        p.size = simulation_configs.size;
        p.delta_t = simulation_configs.delta_t;
        p.frames = simulation_configs.frames;
        p.gs_iterations = opt_.gs_iterations;
        p.diffusion = fluid_configs.diffusion;
        p.viscosity = fluid_configs.viscousity;
        p.mode = opt_.mode;
        p.device = opt_.device;
        p.rank = 0;
        p.world = 1;
        return p;
    }
    void push_params() {
        const EqParams p = params();
        if (p.size != pushed_.size || p.delta_t != pushed_.delta_t || p.frames != pushed_.frames || p.diffusion != pushed_.diffusion ||
            p.viscosity != pushed_.viscosity) {
            check(eq_set_params(h_, &p));
            pushed_ = p;
        }
    }
    EqColors colors(Color32 obstacles) const {
        EqColors c{};
        const Color32 src[3] = {fluid_configs.world_color, fluid_configs.fluid_color, obstacles};
        uint8_t *dst[3] = {c.world, c.fluid, c.obstacle};
        for (int i = 0; i < 3; ++i) {
            dst[i][0] = src[i].r;
            dst[i][1] = src[i].g;
            dst[i][2] = src[i].b;
            dst[i][3] = src[i].a;
        }
        return c;
    }

    eq_fluid *h_ = nullptr;
    DeviceOptions opt_;
    EqParams pushed_{};
    uint64_t noise_frame_ = 0;
};

// A frame buffer in pinned host memory (eq_host_alloc) for the snapshot path
class PinnedFrame {
public:
    explicit PinnedFrame(size_t bytes) : bytes_(bytes) { check(eq_host_alloc(&p_, bytes)); }
    PinnedFrame(const PinnedFrame &) = delete;
    PinnedFrame &operator=(const PinnedFrame &) = delete;
    PinnedFrame(PinnedFrame &&o) noexcept : p_(o.p_), bytes_(o.bytes_) { o.p_ = nullptr; }
    ~PinnedFrame() {
        if (p_) eq_host_free(p_);
    }
    void *data() { return p_; }
    const void *data() const { return p_; }
    size_t bytes() const { return bytes_; }

private:
    void *p_ = nullptr;
    size_t bytes_ = 0;
};

// FluidStep (renderer_helpers.rs:21-24): what simulate() sends to the render thread
struct FluidStep {
    Fluid fluid;
    int64_t frame_number;
};
// what the snapshot path sends instead: the one array the render thread consumes (renderer_helpers.rs:145-167), valid
// until the callback returns
struct FrameView {
    const void *data;       // size*size f32 density (EQ_SNAP_DENSITY) or size*size RGBA pixels (EQ_SNAP_RGBA)
    size_t bytes;
    int64_t frame_number;
};

// CurrentSimulation (renderer_helpers.rs:29-81)
class CurrentSimulation {
public:
    Fluid fluid;
    std::vector<ObstaclesType> obstacles;

    // Default (renderer_helpers.rs:39-48): Fluid::default() and the default rectangle
    CurrentSimulation() : fluid(), obstacles{Rectangle()} {}
    CurrentSimulation(Fluid f, std::vector<ObstaclesType> obs) : fluid(std::move(f)), obstacles(std::move(obs)) {}

    // simulate (renderer_helpers.rs:52-72) as the reference does it: a deep copy of the Fluid per frame goes to `tx`
    template <class Tx>
    void simulate(Tx &&tx) {
        mark_fluid_obstacles();
        for (int64_t i = 0; i < fluid.simulation_configs.frames; ++i) {
            if (fluid.fluid_configs.has_perlin_noise) fluid.add_noise();
            fluid.step();
            tx(FluidStep{fluid.clone(), i});
        }
    }
    // The same loop over the snapshot path: frame i's array travels to a pinned buffer on the copy
    // stream while step i+1 runs; `tx` gets the frames in order, each as soon as it has landed.
    template <class Tx>
    void simulate_frames(int kind, Tx &&tx, Color32 obstacles_color = RED) {
        mark_fluid_obstacles();
        const size_t bytes = fluid.cells() * 4;      // f32 density and RGBA pixels are both 4 bytes per cell
        PinnedFrame buf[2] = {PinnedFrame(bytes), PinnedFrame(bytes)};
        const int64_t frames = fluid.simulation_configs.frames;
        for (int64_t i = 0; i < frames; ++i) {
            if (fluid.fluid_configs.has_perlin_noise) fluid.add_noise();
            fluid.step();
            const int slot = static_cast<int>(i & 1);
            fluid.snapshot_begin(kind, slot, buf[slot].data(), bytes, obstacles_color);
            if (i > 0) {
                fluid.snapshot_wait(1 - slot);       // frame i-1 has landed
                tx(FrameView{buf[1 - slot].data(), bytes, i - 1});
            }
        }
        if (frames > 0) {
            const int last = static_cast<int>((frames - 1) & 1);
            fluid.snapshot_wait(last);
            tx(FrameView{buf[last].data(), bytes, frames - 1});
        }
    }

private:
    // mark_fluid_obstacles (renderer_helpers.rs:76-80)
    void mark_fluid_obstacles() {
        for (ObstaclesType &o : obstacles) fluid.fill_obstacle(o);
    }
};

}  // namespace equilibrium
#endif
