/* sylt2d_b200.h — C ABI of libsylt2d_b200.so: the B200 (sm_100a) implementation of sylt-2d's
 * World::step hot path.  Plain pointers and sizes only; no torch / C++ types.
 *
 * The reference (/root/reference, a pure-Rust crate) has NO FFI / plugin boundary for this path: the path
 * sits behind its public Rust API.  Each entry point below therefore cites the Rust item it stands in
 * for, and INTEGRATION.md shows the `extern "C"` block + `World` wrapper a maintainer adds on the Rust side.
 *
 *   reference item (file:line)                               entry point
 *   World::new(gravity, iterations)      src/world.rs:37-51   sylt_world_create
 *   drop(World)                                               sylt_world_destroy
 *   World::clear                         src/world.rs:67-71   sylt_world_clear
 *   pub world_context                    src/world.rs:11-16   sylt_world_set_context
 *   Body::new + World::add_body          src/body.rs:204-245, src/world.rs:53-55   sylt_world_add_box / _add_bodies
 *   Body::new_polygon + World::add_body  src/body.rs:246-284  sylt_world_add_polygon / _add_bodies
 *   Joint::new + World::add_joint        src/joint.rs:26-55, src/world.rs:63-65    sylt_world_add_joint
 *   pub Joint.softness / .bias_factor    src/joint.rs:17-18   sylt_world_set_joint_params
 *   pub Joint.local_anchor_1 / _2        src/joint.rs:19-20   sylt_world_add_joint_local / sylt_world_set_joint_local_anchors
 *   pub Body.friction / mass / inv_mass / moi / inv_moi / width   src/body.rs:190-196   sylt_world_set_body_props
 *   Body::add_force                      src/body.rs:286-288  sylt_world_add_force
 *   pub Body fields (mutated between steps, e.g. examples/samples/src/main.rs:56-64)   sylt_world_set_bodies
 *   World::broad_phase                   src/world.rs:73-105  sylt_world_broad_phase
 *   World::step(dt) -> Result            src/world.rs:107-152 sylt_world_step / sylt_world_step_n
 *   World::iter_bodies / pub bodies      src/world.rs:57-61   sylt_world_get_bodies
 *   pub arbiters (HashMap<ArbiterKey,Arbiter>) src/world.rs:23, src/arbiter.rs:69-114   sylt_world_get_arbiters
 *   pub joints                           src/world.rs:22      sylt_world_get_joints
 *   collide / collide_polygons           src/collide.rs:140, src/collide_polygon.rs:195   sylt_collide_pairs
 *   Sylt2DErrors                         src/errors.rs:5-9    SYLT_ERR_* codes
 *
 * Batched independent worlds (BASELINE.json configs 4 and 5; one CTA per world, no counterpart in the
 * reference — it is N separate `World`s stepped in a loop): sylt_batch_*.
 *
 * Threading: a handle is bound to one CUDA device and one stream and is not thread-safe; distinct handles
 * may be driven from distinct threads.  All calls are synchronous on return unless named *_async.
 * There is no CPU fallback: every compute entry point returns SYLT_ERR_CUDA when no device is usable.
 */
#ifndef SYLT2D_B200_H
#define SYLT2D_B200_H
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

#define SYLT_OK 0
#define SYLT_ERR_NO_INVERSE 1     /* Sylt2DErrors::MathOperations(MathErrors::NoInverse{matrix}); step aborted mid-way (world.rs:127-129) */
#define SYLT_ERR_ARBITER 2        /* Sylt2DErrors::Arbiter(..) — unreachable in the reference, kept for the mapping */
#define SYLT_ERR_BODY_ORDER 3     /* reference panics: bodies not added in increasing-id order (arbiter.rs:120-122) */
#define SYLT_ERR_BODY_NOT_FOUND 4 /* reference panics: Joint::new body lookup (joint.rs:31,36) */
#define SYLT_ERR_CAPACITY 5       /* a fixed limit was exceeded: > 64 arbiter colours at one body, a polygon manifold over 2*SYLT_MAX_POLY_VERTS
                                     contacts, a batch world over its slot capacity.  (Pair / arbiter / contact buffers of a single world grow
                                     by themselves.)  The failed step changed nothing. */
#define SYLT_ERR_CUDA 6           /* CUDA runtime error / no device; sylt_last_error_string() has the text */
#define SYLT_ERR_INVALID 7        /* bad argument */
#define SYLT_ERR_UNSUPPORTED 8    /* e.g. polygons with > SYLT_MAX_POLY_VERTS vertices */
#define SYLT_ERR_INTERNAL 9       /* solver watchdog: an ordered wait never completed (a bug; the step result is invalid) */

#define SYLT_MAX_POLY_VERTS 32    /* per polygon; a manifold of two polygons holds up to n1 + n2 contacts (reference: unbounded) */
#define SYLT_SHAPE_BOX 0
#define SYLT_SHAPE_POLYGON 1

typedef struct SyltWorld SyltWorld;
typedef struct SyltBatch SyltBatch;

/* Body construction record (Body::new / Body::new_polygon + the pub fields the samples set). */
typedef struct {
    uint64_t id;              /* Body.id (process-global counter on the Rust side) */
    int32_t shape;            /* SYLT_SHAPE_* */
    int32_t vert_offset;      /* polygons: first vertex in the `verts` array passed alongside */
    int32_t vert_count;
    int32_t group;            /* collision group, 0 .. 65535 (default 0): bodies of different groups never collide, so ONE World can hold many
                                 independent worlds — boxes and polygons — stepped by the same launches (ids must stay unique and increasing) */
    float width_x, width_y;   /* boxes: Body.width; polygons: ignored (bounding box is recomputed, body.rs:263) */
    float mass, friction;     /* mass == FLT_MAX  =>  static (body.rs:208-216) */
    float x, y, rotation, vx, vy, angular_velocity;
} SyltBodyDesc;               /* 64 bytes */

typedef struct {
    uint64_t id1, id2;        /* Body.id of body_1 / body_2 as passed to Joint::new */
    float anchor_x, anchor_y; /* world-space anchor */
    float softness, bias_factor; /* Joint::new defaults are 0.0 / 0.2 (joint.rs:47-48) */
} SyltJointDesc;              /* 32 bytes */

typedef struct {
    float x, y, rotation, vx, vy, angular_velocity, fx, fy, torque;
} SyltBodyState;              /* 36 bytes: the mutable pub fields of Body */

typedef struct {              /* ContactInfo (arbiter.rs:69-83) */
    float px, py, nx, ny, r1x, r1y, r2x, r2y;
    float separation, pn, pt, pnb, mass_normal, mass_tangent, bias;
    uint8_t in_edge_1, out_edge_1, in_edge_2, out_edge_2; /* FeaturePair.edges */
    int32_t feature_value;                                 /* FeaturePair.value: always 0 (reference never writes it) */
} SyltContact;                /* 68 bytes */

typedef struct {              /* (ArbiterKey, Arbiter) (arbiter.rs:85-114) */
    uint64_t id1, id2;        /* ArbiterKey */
    int32_t body1, body2;     /* indices into World.bodies */
    float friction;
    int32_t num_contacts;
    int32_t contact_offset;   /* first contact of this arbiter in the contacts array */
    int32_t color;            /* Gauss-Seidel colour this step (solve order = colour ascending) */
} SyltArbiterInfo;            /* 40 bytes */

typedef struct {              /* Joint (joint.rs:10-23) */
    float px, py, bias_x, bias_y, r1x, r1y, r2x, r2y;
    float m11, m12, m21, m22; /* M.col1.x, M.col2.x, M.col1.y, M.col2.y */
    float bias_factor, softness, la1x, la1y, la2x, la2y;
    int32_t body1, body2;
} SyltJointState;             /* 80 bytes */

typedef struct {              /* per-step counters of the last completed step (SURVEY.md §8d symbols) */
    int32_t bodies;           /* B */
    int32_t candidate_pairs;  /* P: AABB-overlapping, not static-static */
    int32_t arbiters;         /* A */
    int32_t contacts;         /* C */
    int32_t joints;           /* J */
    int32_t colors;           /* arbiter colours used */
    int32_t joint_colors;
    int32_t color_rounds;     /* colouring rounds */
} SyltStepStats;

typedef struct {              /* result of sylt_batch_stats: the only quantity ever reduced across GPUs */
    uint64_t body_steps;
    uint64_t contacts;        /* sum over worlds of contacts in the last step */
    uint64_t errors;          /* worlds whose last step raised an error */
    double kinetic_energy;    /* sum over bodies of 0.5 m v^2 + 0.5 I w^2 (dynamic bodies) */
    float max_penetration;    /* max over contacts of -separation */
    float _pad;
} SyltBatchStats;

const char* sylt_last_error_string(void);
int sylt_device_count(void);

/* ------------------------------------------------------------------ single (large) world */
int sylt_world_create(float gravity_x, float gravity_y, uint32_t iterations, int device, SyltWorld** out);
int sylt_world_destroy(SyltWorld* w);
int sylt_world_clear(SyltWorld* w);
int sylt_world_set_context(SyltWorld* w, int accumulate_impulse, int warm_starting, int position_correction);
int sylt_world_set_iterations(SyltWorld* w, uint32_t iterations);
/* Opt-in corrected physics, default OFF (= bit-for-bit the reference).  fix_features: old and new contacts are matched by
 * their packed feature edges (Box2D-lite) instead of the never-written FeaturePair.value, which makes every contact
 * inherit old contact 0's impulses (arbiter.rs:145-176).  fix_inverse: Mat2x2 inverse scaled by 1/det instead of det
 * (math_utils.rs:185-199), i.e. rigid instead of very soft joints. */
int sylt_world_set_fixes(SyltWorld* w, int fix_features, int fix_inverse);
/* 0 = keep; defaults: pairs 16*B, arbiters 8*B, contacts 16*B (grown automatically when bodies are added).  These are starting
 * sizes, not limits: a step that overflows a buffer changes nothing, the buffers are grown and the step is repeated (the
 * reference's arbiter container is a HashMap, world.rs:24). */
int sylt_world_set_capacity(SyltWorld* w, int64_t max_pairs, int64_t max_arbiters, int64_t max_contacts);

int sylt_world_add_box(SyltWorld* w, uint64_t id, float width_x, float width_y, float mass, float friction,
                       float x, float y, float rotation, float vx, float vy, float angular_velocity, int32_t* out_index);
int sylt_world_add_polygon(SyltWorld* w, uint64_t id, const float* verts_xy, int32_t n_verts, float mass, float friction,
                           float x, float y, float rotation, float vx, float vy, float angular_velocity, int32_t* out_index);
int sylt_world_add_bodies(SyltWorld* w, const SyltBodyDesc* bodies, int32_t n, const float* verts_xy, int32_t n_verts);
int sylt_world_add_joint(SyltWorld* w, uint64_t id1, uint64_t id2, float anchor_x, float anchor_y, int32_t* out_index);
int sylt_world_add_joints(SyltWorld* w, const SyltJointDesc* joints, int32_t n);
int sylt_world_set_joint_params(SyltWorld* w, int32_t joint, float softness, float bias_factor);
/* Joint with the local anchors given as they are (pub Joint.local_anchor_1 / _2, joint.rs:19-20: Joint::new computes them once
 * from the bodies' poses, joint.rs:37-40, and a caller may edit them afterwards) */
int sylt_world_add_joint_local(SyltWorld* w, uint64_t id1, uint64_t id2, float la1x, float la1y, float la2x, float la2y,
                               float softness, float bias_factor, int32_t* out_index);
int sylt_world_set_joint_local_anchors(SyltWorld* w, int32_t joint, float la1x, float la1y, float la2x, float la2y);
int sylt_world_add_force(SyltWorld* w, int32_t body, float fx, float fy);

int sylt_world_num_bodies(SyltWorld* w);
int sylt_world_num_joints(SyltWorld* w);
int sylt_world_num_arbiters(SyltWorld* w);
int sylt_world_num_contacts(SyltWorld* w);

int sylt_world_get_bodies(SyltWorld* w, SyltBodyState* out, int32_t cap);
int sylt_world_set_bodies(SyltWorld* w, int32_t first, int32_t count, const SyltBodyState* in);
/* mass, inv_mass, moi, inv_moi, width.x, width.y, friction, shape — 8 floats per body (Body's read-mostly pub fields) */
int sylt_world_get_body_props(SyltWorld* w, float* out8, int32_t cap);
/* The same record written back: the reference reads these pub Body fields on every step (inv_mass / inv_moi: world.rs:115-118,
 * arbiter.rs:198-211, joint.rs:74-92; width: collide.rs:141-142; friction: arbiter.rs:128 when an arbiter is created), so a caller
 * may overwrite them between steps.  Taken as given (nothing is recomputed from mass); the shape entry is ignored. */
int sylt_world_set_body_props(SyltWorld* w, int32_t first, int32_t count, const float* props8);
/* arbiters in ascending (id1,id2) order with their contacts; returns count or -(error) */
int sylt_world_get_arbiters(SyltWorld* w, SyltArbiterInfo* arbiters, int32_t arbiter_cap, SyltContact* contacts, int32_t contact_cap);
int sylt_world_get_joints(SyltWorld* w, SyltJointState* out, int32_t cap);
/* solve order used by the last step: joint indices, colour-major (arbiters: sort get_arbiters by (color,key)) */
int sylt_world_get_joint_order(SyltWorld* w, int32_t* out, int32_t cap);
int sylt_world_get_stats(SyltWorld* w, SyltStepStats* out);
/* matrix (col1.x, col2.x, col1.y, col2.y) of the last SYLT_ERR_NO_INVERSE and its joint index (return value) */
int sylt_world_last_error_matrix(SyltWorld* w, float* out4);

int sylt_world_broad_phase(SyltWorld* w);
int sylt_world_step(SyltWorld* w, float dt);
int sylt_world_step_n(SyltWorld* w, float dt, int32_t n);
/* enqueue n steps on the world's stream without waiting; errors surface at sylt_world_sync */
int sylt_world_step_n_async(SyltWorld* w, float dt, int32_t n);
int sylt_world_sync(SyltWorld* w);
/* device time of everything enqueued between the two marks (CUDA events on the world's stream) */
int sylt_world_timer_start(SyltWorld* w);
int sylt_world_timer_stop(SyltWorld* w, float* out_ms);
/* number of kernel launches issued by this handle so far */
int64_t sylt_world_launch_count(SyltWorld* w);
/* Diagnostics: with kernel timing on, CUDA events are recorded between the launches of every step (this serialises the chain of
 * short kernels, so the step itself gets slower).  get_kernel_times returns the durations [us] of the LAST step's launches in
 * order — prepare, cells, fill cells, rows count, rows emit, narrow, (re-cut +) build, solve — and their number (or -error). */
int sylt_world_set_kernel_timing(SyltWorld* w, int on);
int sylt_world_get_kernel_times(SyltWorld* w, float* out_us, int32_t cap);

/* Narrow phase on n explicit pairs (collide / collide_polygons called directly, as examples/collision-debug does).
 * a[i], b[i] describe the two bodies of pair i; verts holds polygon vertices.  Writes up to `cap` contacts per pair
 * into out[i*cap ...] and the count into out_counts[i]. */
int sylt_collide_pairs(int device, const SyltBodyDesc* a, const SyltBodyDesc* b, int32_t n, const float* verts_xy, int32_t n_verts,
                       SyltContact* out, int32_t cap, int32_t* out_counts);
/* (cos, sin) of n angles with the device routine of sylt_sincos.h (Mat2x2::new_from_angle, math_utils.rs:164-172) */
int sylt_sincos(int device, const float* angles, int32_t n, float* out_sin, float* out_cos);

/* ------------------------------------------------------------------ batched independent worlds */
/* CSR description: world k owns bodies[body_offsets[k] .. body_offsets[k+1]) and joints[joint_offsets[k] ..).
 * Boxes only.  All worlds share gravity / iterations / context. */
int sylt_batch_create(int32_t n_worlds, const int32_t* body_offsets, const SyltBodyDesc* bodies,
                      const int32_t* joint_offsets, const SyltJointDesc* joints,
                      float gravity_x, float gravity_y, uint32_t iterations, int device, SyltBatch** out);
int sylt_batch_destroy(SyltBatch* b);
int sylt_batch_set_context(SyltBatch* b, int accumulate_impulse, int warm_starting, int position_correction);
/* Candidate-pair slots per world (AABB-overlapping pairs a step may see; default 2 * max bodies per world + 32, 0 restores it).
 * A step that finds more raises SYLT_ERR_CAPACITY for that world (sylt_batch_get_errors) and freezes it.  The slots live in
 * shared memory (97 B each): SYLT_ERR_UNSUPPORTED when a world no longer fits one CTA.  Drops the contact history (next step
 * starts cold), so call it before stepping.  The reference has no such limit (HashMap, world.rs:24). */
int sylt_batch_set_capacity(SyltBatch* b, int32_t slots_per_world);
/* Worlds whose arbiters sylt_batch_get_arbiters can read back (<= 64; default: the first 64 worlds of the batch).  The
 * choice takes effect with the next step.  Body states of every world are always readable (sylt_batch_get_bodies). */
int sylt_batch_set_export_worlds(SyltBatch* b, const int32_t* worlds, int32_t n);
/* see sylt_world_set_fixes; fix_features is not available for batches (SYLT_ERR_UNSUPPORTED) */
int sylt_batch_set_fixes(SyltBatch* b, int fix_features, int fix_inverse);
int sylt_batch_step_n(SyltBatch* b, float dt, int32_t n);
int sylt_batch_step_n_async(SyltBatch* b, float dt, int32_t n);
int sylt_batch_sync(SyltBatch* b);
/* step_n with HOST buffers (pinned memory recommended): upload every body state from `in` (NULL: keep device state), run n
 * steps, download every body state to `out` (NULL: skip).  World chunks are pipelined over several streams so that the
 * H2D copy, the kernels and the D2H copy overlap.  Equivalent to set_bodies + step_n + get_bodies. */
int sylt_batch_step_host(SyltBatch* b, float dt, int32_t n, const SyltBodyState* in, SyltBodyState* out);
/* same with compact 24-byte records per body: x, y, rotation, vx, vy, angular_velocity (the fields a caller mirrors every
 * step; a batch has no external forces) */
int sylt_batch_step_host6(SyltBatch* b, float dt, int32_t n, const float* in6, float* out6);
int sylt_batch_timer_start(SyltBatch* b);
int sylt_batch_timer_stop(SyltBatch* b, float* out_ms);
int64_t sylt_batch_launch_count(SyltBatch* b);
int sylt_batch_num_worlds(SyltBatch* b);
int64_t sylt_batch_num_bodies(SyltBatch* b);
/* body states of worlds [first_world, first_world+n_worlds), concatenated in CSR order */
int sylt_batch_get_bodies(SyltBatch* b, int32_t first_world, int32_t n_worlds, SyltBodyState* out, int64_t cap);
int sylt_batch_set_bodies(SyltBatch* b, int32_t first_world, int32_t n_worlds, const SyltBodyState* in, int64_t count);
/* arbiters (key order) + contacts of one world; returns count or -(error) */
int sylt_batch_get_arbiters(SyltBatch* b, int32_t world, SyltArbiterInfo* arbiters, int32_t arbiter_cap, SyltContact* contacts, int32_t contact_cap);
int sylt_batch_get_joints(SyltBatch* b, int32_t world, SyltJointState* out, int32_t cap);
int sylt_batch_get_joint_order(SyltBatch* b, int32_t world, int32_t* out, int32_t cap);
/* per-world error code of the last step (SYLT_OK / SYLT_ERR_NO_INVERSE / SYLT_ERR_CAPACITY / SYLT_ERR_BODY_ORDER) */
int sylt_batch_get_errors(SyltBatch* b, int32_t* out, int32_t cap);
int sylt_batch_stats(SyltBatch* b, SyltBatchStats* out);

#ifdef __cplusplus
}
#endif
#endif
