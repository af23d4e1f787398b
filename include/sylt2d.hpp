// sylt2d.hpp — header-only C++ mirror of the reference crate's public API for the step path, on top of the C ABI.
// The reference is compiled (Rust) code and this image has no Rust toolchain, so this is the compiled-language
// host side a caller can use today; names / argument meaning / error behaviour follow the Rust items:
//   Vec2 (src/math_utils.rs:23), Body::new / new_polygon (src/body.rs:204,246), Joint::new (src/joint.rs:26),
//   WorldContext (src/world.rs:11), World::new / add_body / add_joint / clear / step / broad_phase (src/world.rs:36-153),
//   Sylt2DErrors (src/errors.rs:5) -> sylt2d::Error (step returns Result in Rust; here it throws).
#pragma once
#include <atomic>
#include <cfloat>
#include <map>
#include <stdexcept>
#include <string>
#include <utility>
#include <vector>
#include "sylt2d_b200.h"

namespace sylt2d {

struct Vec2 { float x = 0, y = 0; Vec2() = default; Vec2(float x_, float y_) : x(x_), y(y_) {} };

struct Error : std::runtime_error {
    int code;                  // SYLT_ERR_*
    float matrix[4] = {0, 0, 0, 0};  // NoInverse { matrix }: col1.x, col2.x, col1.y, col2.y
    int joint = -1;
    Error(int c, const std::string& what) : std::runtime_error(what), code(c) {}
};

enum class Shape { Box, ConvexPolygon };

struct Body {
    size_t id = 0;
    Vec2 position, velocity, force, width;
    float rotation = 0, angular_velocity = 0, torque = 0, friction = 0, mass = 0;
    std::vector<Vec2> vertices;
    Shape shape = Shape::Box;
    static size_t next_id() { static std::atomic<size_t> c{1}; return c.fetch_add(1); }  // BODY_ID_COUNTER (body.rs:201)
    static Body make(Vec2 width, float mass) { Body b; b.id = next_id(); b.width = width; b.mass = mass; return b; }          // Body::new
    static Body make_polygon(std::vector<Vec2> v, float mass) {                                                              // Body::new_polygon
        Body b; b.id = next_id(); b.vertices = std::move(v); b.mass = mass; b.shape = Shape::ConvexPolygon; return b;
    }
    void add_force(Vec2 f) { force.x += f.x; force.y += f.y; }
};

struct WorldContext { bool accumulate_impulse = true, warm_starting = false, position_correction = true; };  // world.rs:38-42

struct Joint { size_t body_1, body_2; Vec2 anchor; float softness = 0.0f, bias_factor = 0.2f; int index = -1; };

struct Arbiter { float friction; std::vector<SyltContact> contacts; int color; };

class World {
  public:
    WorldContext world_context;
    World(Vec2 gravity, unsigned iterations, int device = 0) {
        int e = sylt_world_create(gravity.x, gravity.y, iterations, device, &h_);
        if (e) throw Error(e, std::string("sylt_world_create: ") + sylt_last_error_string());
    }
    ~World() { sylt_world_destroy(h_); }
    World(const World&) = delete;
    World& operator=(const World&) = delete;

    int add_body(const Body& b) {
        int32_t idx = -1;
        int e;
        if (b.shape == Shape::Box)
            e = sylt_world_add_box(h_, b.id, b.width.x, b.width.y, b.mass, b.friction, b.position.x, b.position.y, b.rotation, b.velocity.x,
                                   b.velocity.y, b.angular_velocity, &idx);
        else {
            std::vector<float> flat;
            for (const Vec2& v : b.vertices) { flat.push_back(v.x); flat.push_back(v.y); }
            e = sylt_world_add_polygon(h_, b.id, flat.data(), (int32_t)b.vertices.size(), b.mass, b.friction, b.position.x, b.position.y,
                                       b.rotation, b.velocity.x, b.velocity.y, b.angular_velocity, &idx);
        }
        check(e);
        if (b.force.x != 0 || b.force.y != 0) check(sylt_world_add_force(h_, idx, b.force.x, b.force.y));
        return idx;
    }
    // Joint::new(body_1, body_2, anchor, &world) + add_joint; throws BODY_NOT_FOUND where the reference panics
    Joint add_joint(const Body& b1, const Body& b2, Vec2 anchor, float softness = 0.0f, float bias_factor = 0.2f) {
        Joint j{b1.id, b2.id, anchor, softness, bias_factor, -1};
        int32_t idx = -1;
        check(sylt_world_add_joint(h_, b1.id, b2.id, anchor.x, anchor.y, &idx));
        check(sylt_world_set_joint_params(h_, idx, softness, bias_factor));
        j.index = idx;
        return j;
    }
    void clear() { check(sylt_world_clear(h_)); }
    void broad_phase() { push(); check(sylt_world_broad_phase(h_)); }
    void step(float dt) { push(); check(sylt_world_step(h_, dt)); }
    void step_n(float dt, int n) { push(); check(sylt_world_step_n(h_, dt, n)); }

    std::vector<SyltBodyState> bodies() const {
        std::vector<SyltBodyState> out((size_t)sylt_world_num_bodies(h_));
        sylt_world_get_bodies(h_, out.data(), (int32_t)out.size());
        return out;
    }
    void set_bodies(const std::vector<SyltBodyState>& s, int first = 0) { check(sylt_world_set_bodies(h_, first, (int32_t)s.size(), s.data())); }
    std::map<std::pair<uint64_t, uint64_t>, Arbiter> arbiters() const {  // HashMap<ArbiterKey, Arbiter>, here ordered by key
        int na = sylt_world_num_arbiters(h_), nc = sylt_world_num_contacts(h_);
        std::vector<SyltArbiterInfo> a((size_t)na);
        std::vector<SyltContact> c((size_t)nc);
        int got = sylt_world_get_arbiters(h_, a.data(), na, c.data(), nc);
        std::map<std::pair<uint64_t, uint64_t>, Arbiter> m;
        for (int i = 0; i < got; i++)
            m[{a[i].id1, a[i].id2}] = Arbiter{a[i].friction, {c.begin() + a[i].contact_offset, c.begin() + a[i].contact_offset + a[i].num_contacts}, a[i].color};
        return m;
    }
    SyltWorld* handle() const { return h_; }

  private:
    SyltWorld* h_ = nullptr;
    void push() { check(sylt_world_set_context(h_, world_context.accumulate_impulse, world_context.warm_starting, world_context.position_correction)); }
    void check(int e) const {
        if (!e) return;
        static const char* names[] = {"ok", "NoInverse", "Arbiter", "BodyOrder", "BodyNotFound", "Capacity", "CUDA", "Invalid", "Unsupported"};
        Error err(e, std::string("sylt2d: ") + (e >= 0 && e <= 8 ? names[e] : "?") + (e == SYLT_ERR_CUDA ? std::string(": ") + sylt_last_error_string() : ""));
        if (e == SYLT_ERR_NO_INVERSE) err.joint = sylt_world_last_error_matrix(h_, err.matrix);
        throw err;
    }
};

}  // namespace sylt2d
