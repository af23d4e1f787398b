/* sylt_oracle.h — TEST INFRASTRUCTURE, NOT PRODUCT.
 *
 * C ABI of the CPU oracle: a literal C++ restatement of the reference's World::step hot path
 * (/root/reference/src/{world,arbiter,collide,collide_polygon,joint,body,math_utils}.rs).
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference leg may
 * load this library; the product (libsylt2d_b200.so) never links or calls it.
 *
 * Parity status: the reference cannot be compiled here (no cargo/rustc), and its own tests pin
 * only contact COUNTS for four box pairs, Vec2/Mat2x2 spot values and the singular-matrix error
 * (SURVEY.md §4).  Those are checked in tests/test_oracle_kat.py.  Everything else on the path
 * (step, update, pre_step, apply_impulse, joints, polygons) is "parity unpinned" by the
 * reference: it is cross-validated by an independent numpy-f32 restatement (oracle/narrow_np.py).
 */
#ifndef SYLT_ORACLE_H
#define SYLT_ORACLE_H
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

typedef struct OrcWorld OrcWorld;

/* Same record layouts as include/sylt2d_b200.h (SyltBodyState / SyltContact / SyltArbiterInfo). */
typedef struct {
    float x, y, rotation, vx, vy, angular_velocity, fx, fy, torque;
} OrcBodyState;

typedef struct {
    float px, py, nx, ny, r1x, r1y, r2x, r2y;
    float separation, pn, pt, pnb, mass_normal, mass_tangent, bias;
    uint8_t in_edge_1, out_edge_1, in_edge_2, out_edge_2;
    int32_t feature_value;
} OrcContact;

typedef struct {
    uint64_t id1, id2;
    int32_t body1, body2; /* indices into world.bodies */
    float friction;
    int32_t num_contacts;
    int32_t contact_offset;
    int32_t color; /* always 0 in the oracle */
} OrcArbiterInfo;

typedef struct {
    float px, py, bias_x, bias_y, r1x, r1y, r2x, r2y;
    float m11, m12, m21, m22; /* col1.x, col2.x, col1.y, col2.y */
    float bias_factor, softness, la1x, la1y, la2x, la2y;
    int32_t body1, body2;
} OrcJointState;

OrcWorld* orc_world_create(float gx, float gy, uint32_t iterations);
void orc_world_destroy(OrcWorld* w);
void orc_world_clear(OrcWorld* w);
void orc_world_set_context(OrcWorld* w, int accumulate_impulse, int warm_starting, int position_correction);
/* sincos_mode: 0 = libm sinf/cosf (reference-faithful), 1 = shared sylt_sincos.h (device-matched).
 * broadphase : 0 = all pairs (reference, world.rs:74-77), 1 = uniform grid candidates (same pair set). */
void orc_world_set_mode(OrcWorld* w, int sincos_mode, int broadphase);
void orc_world_set_iterations(OrcWorld* w, uint32_t iterations);
/* opt-in corrected physics (NOT the reference's behaviour; mirrors sylt_world_set_fixes): contact matching by packed
 * feature edges instead of the never-written FeaturePair.value (quirk Q-2), Mat2x2 inverse scaled by 1/det (quirk Q-3) */
void orc_world_set_fixes(OrcWorld* w, int fix_features, int fix_inverse);

/* return body index >= 0, or -(error code) */
int orc_world_add_box(OrcWorld* w, uint64_t id, float wx, float wy, float mass, float friction,
                      float x, float y, float rot, float vx, float vy, float omega);
int orc_world_add_polygon(OrcWorld* w, uint64_t id, const float* verts_xy, int n, float mass, float friction,
                          float x, float y, float rot, float vx, float vy, float omega);
/* Joint::new (joint.rs:26-55): anchor in world space; returns joint index or -(error) */
int orc_world_add_joint(OrcWorld* w, uint64_t id1, uint64_t id2, float ax, float ay);
int orc_world_set_joint_params(OrcWorld* w, int j, float softness, float bias_factor);
int orc_world_add_force(OrcWorld* w, int body, float fx, float fy);

int orc_world_num_bodies(OrcWorld* w);
int orc_world_num_joints(OrcWorld* w);
int orc_world_num_arbiters(OrcWorld* w);
int orc_world_num_contacts(OrcWorld* w);

int orc_world_get_bodies(OrcWorld* w, OrcBodyState* out, int cap);
int orc_world_set_bodies(OrcWorld* w, int first, int count, const OrcBodyState* in);
/* mass, inv_mass, moi, inv_moi, width.x, width.y, friction, shape  (8 floats per body) */
int orc_world_get_body_props(OrcWorld* w, float* out8, int cap);
int orc_world_set_body_props(OrcWorld* w, int first, int count, const float* in8);
int orc_world_set_joint_local_anchors(OrcWorld* w, int j, float la1x, float la1y, float la2x, float la2y);
int orc_world_get_arbiters(OrcWorld* w, OrcArbiterInfo* arb, int arb_cap, OrcContact* contacts, int contact_cap);
int orc_world_get_joints(OrcWorld* w, OrcJointState* out, int cap);

/* 0 ok, else error code (same numbering as include/sylt2d_b200.h) */
int orc_world_broad_phase(OrcWorld* w);
int orc_world_step(OrcWorld* w, float dt);
/* Same as step, but arbiters are pre-stepped / iterated in the order of `keys` (id1,id2 pairs, n_arb of
 * them; must be a permutation of the arbiter set after this step's broad phase, else SYLT_ERR_INVALID)
 * and joints in the order `joint_order` (n_joint indices, or NULL for Vec order). */
int orc_world_step_ordered(OrcWorld* w, float dt, const uint64_t* keys, int n_arb, const int32_t* joint_order, int n_joint);
/* matrix of the last NoInverse error (col1.x, col2.x, col1.y, col2.y) and joint index */
int orc_world_last_error_matrix(OrcWorld* w, float* out4);

/* Single-pair narrow phase on explicit body descriptions (collide.rs:140-297 / collide_polygon.rs:195-203).
 * shape: 0 box (verts ignored), 1 polygon.  Returns number of contacts written (<= cap) or -(error). */
int orc_collide_pair(int sincos_mode,
                     int shape_a, float wax, float way, const float* verts_a, int na, float xa, float ya, float ra,
                     int shape_b, float wbx, float wby, const float* verts_b, int nb, float xb, float yb, float rb,
                     OrcContact* out, int cap);

/* math_utils spot checks (math_utils.rs:185-199): returns 0 and writes 4 floats, or SYLT_ERR_NO_INVERSE */
int orc_mat_invert(const float* m4, float* out4);
void orc_sincos(int mode, float a, float* s, float* c);

/* Multi-threaded batch stepping for the CPU baseline: n_worlds handles, each stepped `steps` times,
 * worlds distributed over `threads` host threads.  Returns first non-zero error code or 0. */
int orc_worlds_step_mt(OrcWorld** worlds, int n_worlds, float dt, int steps, int threads);

#ifdef __cplusplus
}
#endif
#endif
