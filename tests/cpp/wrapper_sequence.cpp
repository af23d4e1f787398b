// wrapper_sequence.cpp — TEST: replays, in C++, the exact call sequence the Rust wrapper crate issues around World::step
// (rust/src/world.rs: upload_dirty -> sylt_world_step -> mirror_back, lazy arbiter fetch) and checks the result against the CPU
// oracle after every step, bit for bit.  There is no rustc in this image, so this is how the wrapper's protocol is exercised:
// host-side Body / Joint objects with pub fields that the "caller" mutates between steps, diffed against a mirror of the device
// state exactly as the crate does.  Links libsylt2d_b200.so (the product) and oracle/liborc.so (the checker).
#include <algorithm>
#include <cfloat>
#include <cstdio>
#include <cstring>
#include <memory>
#include <vector>
#include "../../include/sylt2d_b200.h"
#include "../../oracle/sylt_oracle.h"

struct HBody {   // rust/src/body.rs Body: the pub fields
    uint64_t id; int shape; float wx, wy, mass, inv_mass, moi, inv_moi, friction;
    SyltBodyState st;
    std::vector<float> verts;
};
struct HJoint { uint64_t id1, id2; float la[4], softness, bias_factor; int index; };

static bool differ(const void* a, const void* b, size_t n) { return memcmp(a, b, n) != 0; }
#define CK(x) do { int e_ = (x); if (e_ != 0) { fprintf(stderr, "%s -> %d (%s)\n", #x, e_, sylt_last_error_string()); exit(3); } } while (0)

struct Mirrored {  // rust/src/world.rs World
    SyltWorld* h = nullptr;
    std::vector<HBody> bodies;
    std::vector<HJoint> joints;
    std::vector<SyltBodyState> mirror;
    std::vector<std::vector<float>> props;
    int ctx[3] = {1, 0, 1};
    bool arbiters_valid = false;
    std::vector<SyltArbiterInfo> arb; std::vector<SyltContact> con;

    static std::vector<float> props_of(const HBody& b) { return {b.mass, b.inv_mass, b.moi, b.inv_moi, b.wx, b.wy, b.friction, (float)b.shape}; }
    void add_body(HBody b) {
        int32_t idx;
        if (b.shape == 0) CK(sylt_world_add_box(h, b.id, b.wx, b.wy, b.mass, b.friction, b.st.x, b.st.y, b.st.rotation, b.st.vx, b.st.vy, b.st.angular_velocity, &idx));
        else CK(sylt_world_add_polygon(h, b.id, b.verts.data(), (int)b.verts.size() / 2, b.mass, b.friction, b.st.x, b.st.y, b.st.rotation, b.st.vx, b.st.vy, b.st.angular_velocity, &idx));
        SyltBodyState m = b.st; m.fx = m.fy = m.torque = 0.0f;
        mirror.push_back(m);
        float p8[8];
        bodies.push_back(b);
        std::vector<float> all(8 * bodies.size());
        sylt_world_get_body_props(h, all.data(), (int)bodies.size());   // what the library derived (Body::derived_props in the crate)
        memcpy(p8, &all[8 * (bodies.size() - 1)], sizeof p8);
        HBody& nb = bodies.back();
        nb.inv_mass = p8[1]; nb.moi = p8[2]; nb.inv_moi = p8[3]; if (nb.shape) { nb.wx = p8[4]; nb.wy = p8[5]; }
        props.push_back(props_of(nb));
    }
    void add_joint(HJoint j) {
        int32_t idx;
        CK(sylt_world_add_joint_local(h, j.id1, j.id2, j.la[0], j.la[1], j.la[2], j.la[3], j.softness, j.bias_factor, &idx));
        j.index = idx;
        joints.push_back(j);
    }
    void upload_dirty() {
        size_t first = 0, pfirst = 0;
        std::vector<SyltBodyState> run;
        std::vector<float> prun;
        for (size_t i = 0; i < bodies.size(); i++) {
            if (differ(&bodies[i].st, &mirror[i], sizeof(SyltBodyState))) {
                if (!run.empty() && first + run.size() != i) { CK(sylt_world_set_bodies(h, (int)first, (int)run.size(), run.data())); run.clear(); }
                if (run.empty()) first = i;
                run.push_back(bodies[i].st);
                mirror[i] = bodies[i].st;
            }
            std::vector<float> p = props_of(bodies[i]);
            if (differ(p.data(), props[i].data(), 7 * sizeof(float))) {
                if (!prun.empty() && pfirst + prun.size() / 8 != i) { CK(sylt_world_set_body_props(h, (int)pfirst, (int)prun.size() / 8, prun.data())); prun.clear(); }
                if (prun.empty()) pfirst = i;
                prun.insert(prun.end(), p.begin(), p.end());
                props[i] = p;
            }
        }
        if (!run.empty()) CK(sylt_world_set_bodies(h, (int)first, (int)run.size(), run.data()));
        if (!prun.empty()) CK(sylt_world_set_body_props(h, (int)pfirst, (int)prun.size() / 8, prun.data()));
        for (const HJoint& j : joints) {
            CK(sylt_world_set_joint_params(h, j.index, j.softness, j.bias_factor));
            CK(sylt_world_set_joint_local_anchors(h, j.index, j.la[0], j.la[1], j.la[2], j.la[3]));
        }
        CK(sylt_world_set_context(h, ctx[0], ctx[1], ctx[2]));
    }
    void mirror_back() {
        mirror.resize(bodies.size());
        if (sylt_world_get_bodies(h, mirror.data(), (int)mirror.size()) != (int)mirror.size()) exit(4);
        for (size_t i = 0; i < bodies.size(); i++) bodies[i].st = mirror[i];
        arbiters_valid = false;
    }
    int step(float dt) { upload_dirty(); int rc = sylt_world_step(h, dt); mirror_back(); return rc; }
    void fetch_arbiters() {   // ArbiterMap::deref
        if (arbiters_valid) return;
        int na = sylt_world_num_arbiters(h), nc = sylt_world_num_contacts(h);
        arb.assign((size_t)na, SyltArbiterInfo{}); con.assign((size_t)nc, SyltContact{});
        if (sylt_world_get_arbiters(h, arb.data(), na, con.data(), nc) != na) exit(5);
        arbiters_valid = true;
    }
};

int main() {
    Mirrored w;
    if (sylt_world_create(0.0f, -10.0f, 10, 0, &w.h) != 0) { fprintf(stderr, "no device: %s\n", sylt_last_error_string()); return 2; }
    OrcWorld* o = orc_world_create(0.0f, -10.0f, 10);
    orc_world_set_mode(o, 1, 0);
    uint64_t nid = 1;
    auto box = [&](float wx, float wy, float mass, float fr, float x, float y, float rot = 0.0f, float vx = 0.0f, float vy = 0.0f, float om = 0.0f) {
        HBody b{}; b.id = nid++; b.shape = 0; b.wx = wx; b.wy = wy; b.mass = mass; b.friction = fr;
        b.st = SyltBodyState{x, y, rot, vx, vy, om, 0, 0, 0};
        w.add_body(b);
        orc_world_add_box(o, b.id, wx, wy, mass, fr, x, y, rot, vx, vy, om);
        return (int)w.bodies.size() - 1;
    };
    box(100.0f, 20.0f, FLT_MAX, 0.2f, 0.0f, -10.0f);
    for (int r = 0; r < 5; r++) for (int c = r; c < 5; c++) box(1.0f, 1.0f, 10.0f, 0.2f, -3.0f + 0.5625f * r + 1.125f * (c - r), 0.75f + 1.05f * r);
    {   // a polygon (box-polygon and polygon-polygon pairs take the polygon path)
        HBody b{}; b.id = nid++; b.shape = 1; b.mass = 12.0f; b.friction = 0.3f; b.verts = {0.6f, 0.0f, 0.3f, 0.52f, -0.3f, 0.52f, -0.6f, 0.0f, -0.3f, -0.52f, 0.3f, -0.52f};
        b.st = SyltBodyState{4.5f, 1.0f, 0.2f, 0, 0, 0, 0, 0, 0};
        w.add_body(b);
        orc_world_add_polygon(o, b.id, b.verts.data(), 6, b.mass, b.friction, 4.5f, 1.0f, 0.2f, 0, 0, 0);
    }
    const int pend = box(0.5f, 1.5f, 20.0f, 0.2f, 9.0f, 6.0f);
    {   // Joint::new(ground, pendulum, anchor): local anchors computed on the host as the crate does, passed as they are
        const float ax = 9.0f, ay = 9.0f;
        HJoint j{w.bodies[0].id, w.bodies[(size_t)pend].id, {ax - 0.0f, ay + 10.0f, ax - 9.0f, ay - 6.0f}, 0.0f, 0.2f, -1};  // rotations are 0: R^T = I
        w.add_joint(j);
        int oj = orc_world_add_joint(o, j.id1, j.id2, ax, ay);
        orc_world_set_joint_params(o, oj, 0.0f, 0.2f);
    }
    w.ctx[1] = 1;
    orc_world_set_context(o, 1, 1, 1);
    const float dt = 1.0f / 60.0f;
    for (int s = 0; s < 45; s++) {
        if (s == 10) box(1.0f, 1.0f, 50.0f, 0.2f, -8.0f, 12.0f, 0.3f, 12.0f, -18.0f, 7.0f);   // launch_bomb: add_body between steps
        if (s == 15) {   // the caller writes pub fields
            w.bodies[3].st.vx += 2.5f; w.bodies[7].st.angular_velocity = -3.0f; w.bodies[4].st.fx = 40.0f;
            w.bodies[5].friction = 0.7f;
            w.bodies[6].mass = 25.0f; w.bodies[6].inv_mass = 1.0f / 25.0f; w.bodies[6].moi = 25.0f * 2.0f / 12.0f; w.bodies[6].inv_moi = 1.0f / w.bodies[6].moi;
            w.joints[0].softness = 0.01f; w.joints[0].la[2] += 0.1f;
            std::vector<OrcBodyState> os((size_t)orc_world_num_bodies(o));
            orc_world_get_bodies(o, os.data(), (int)os.size());
            os[3].vx += 2.5f; os[7].angular_velocity = -3.0f; os[4].fx = 40.0f;
            orc_world_set_bodies(o, 0, (int)os.size(), os.data());
            std::vector<float> op(8 * os.size());
            orc_world_get_body_props(o, op.data(), (int)os.size());
            op[8 * 5 + 6] = 0.7f; op[8 * 6 + 0] = 25.0f; op[8 * 6 + 1] = 1.0f / 25.0f; op[8 * 6 + 2] = 25.0f * 2.0f / 12.0f; op[8 * 6 + 3] = 1.0f / (25.0f * 2.0f / 12.0f);
            orc_world_set_body_props(o, 0, (int)os.size(), op.data());
            orc_world_set_joint_params(o, 0, 0.01f, 0.2f);
            orc_world_set_joint_local_anchors(o, 0, w.joints[0].la[0], w.joints[0].la[1], w.joints[0].la[2], w.joints[0].la[3]);
        }
        if (s == 25) { w.ctx[1] = 0; orc_world_set_context(o, 1, 0, 1); }     // the sample's warm-starting check box
        if (w.step(dt) != 0) { fprintf(stderr, "step %d failed\n", s); return 6; }
        w.fetch_arbiters();                                                   // (a renderer reads world.arbiters every frame)
        std::vector<int> ord(w.arb.size());
        for (size_t k = 0; k < ord.size(); k++) ord[k] = (int)k;
        std::stable_sort(ord.begin(), ord.end(), [&](int a, int b) { return w.arb[(size_t)a].color < w.arb[(size_t)b].color; });  // arbiters come key-sorted
        std::vector<uint64_t> keys;
        for (int k : ord) { keys.push_back(w.arb[(size_t)k].id1); keys.push_back(w.arb[(size_t)k].id2); }
        std::vector<int32_t> jo((size_t)sylt_world_num_joints(w.h));
        sylt_world_get_joint_order(w.h, jo.data(), (int)jo.size());
        if (orc_world_step_ordered(o, dt, keys.data(), (int)ord.size(), jo.data(), (int)jo.size()) != 0) { fprintf(stderr, "oracle step %d failed\n", s); return 7; }
        std::vector<OrcBodyState> os((size_t)orc_world_num_bodies(o));
        orc_world_get_bodies(o, os.data(), (int)os.size());
        if (os.size() != w.bodies.size()) return 8;
        for (size_t i = 0; i < os.size(); i++)
            if (memcmp(&os[i], &w.bodies[i].st, 6 * sizeof(float)) != 0 || os[i].fx != w.bodies[i].st.fx) {
                fprintf(stderr, "step %d body %zu differs: gpu (%.9g %.9g %.9g) oracle (%.9g %.9g %.9g)\n", s, i, w.bodies[i].st.x, w.bodies[i].st.y,
                        w.bodies[i].st.vx, os[i].x, os[i].y, os[i].vx);
                return 1;
            }
    }
    w.fetch_arbiters();
    printf("wrapper sequence ok: %zu bodies, %zu arbiters, %zu contacts bit-identical to the oracle over 45 steps\n", w.bodies.size(), w.arb.size(), w.con.size());
    sylt_world_destroy(w.h);
    orc_world_destroy(o);
    return 0;
}
