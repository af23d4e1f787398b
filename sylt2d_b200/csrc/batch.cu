// batch.cu — batches of independent small worlds (BASELINE.json configs 4 and 5): ONE CTA PER WORLD.
//
// The reference has no batched API: this is N separate `World`s each running World::step
// (/root/reference/src/world.rs:107-152).  Here a CTA loads its world into shared memory once, runs `n_steps`
// complete steps there, and writes the state back:
//
//   per step, all in shared memory
//     bodies     (cos,sin), AABB, force integration                         world.rs:113-120
//     pairs      warp-per-row all-pairs AABB filter -> key-ordered candidate slots (static-static skipped, :79-81)
//     narrow     one thread per candidate: box_box (narrow.cuh)             collide.rs:140-297
//     match      binary search of last step's key-sorted slots: Arbiter::update with quirks Q-2 / Q-17
//     colour     persisting arbiters keep their colour, new ones get the lowest free colour (sequential, warp 0)
//     solve      Arbiter::pre_step, Joint::pre_step, iterations x apply_impulse, colour by colour with
//                __syncthreads() between colours                            arbiter.rs:182-298, joint.rs:57-130
//     integrate  positions, all bodies                                      world.rs:143-150
//
// Worlds never exchange data, so a batch shards across GPUs with no collective; the only cross-GPU quantity is the
// SyltBatchStats record (sylt_batch_stats), reduced by the caller.
#include <algorithm>
#include <numeric>
#include "common.h"
#include "narrow.cuh"
#include "solver.cuh"

using namespace sylt;

namespace {

constexpr int BNC = 32;        // colours per world
constexpr int MAX_JC = 16;     // joint colours per world
constexpr unsigned char NO_COLOR = 0xff;
constexpr int MAX_EXPORT = 64;  // worlds whose arbiters are exported in full (parity tests / renderer read-back)
constexpr int ROWT = 8;         // partners remembered per row between the count and the emit pass

struct SlotExport {  // full record of one arbiter slot, written for worlds in the export range only
    unsigned key;    // body1 << 16 | body2 (indices local to the world)
    float friction;
    int count, color;
    float nx, ny;
    float geo[2][3];  // px, py, separation   (written by the narrow phase of the launch's last step)
    float sol[2][9];  // r1x,r1y,r2x,r2y,mass_n,mass_t,bias,pn,pt   (written after the solve)
    unsigned feat[2];
};

struct BatchArgs {
    int n_worlds, world_begin, n_steps, iterations, accumulate, warm, poscorr, fix_inverse;
    float dt, inv_dt, gx, gy, cell;
    int Bmax, Jmax, cap, nbk;
    const int* exp_map;              // per world: export slot (full arbiter read-back, sylt_batch_set_export_worlds) or -1
    const int *body_off, *joint_off;
    float4 *pose, *vel;              // (x,y,rot,_), (vx,vy,w,_)
    const float4* mp;                // (inv_mass, inv_moi, friction, _)
    const float2* wh;                // Body.width
    const unsigned char* bflags;     // bit0 static, bit1 large (bypasses the grid)
    const int2* jbodies;             // world-local body indices
    const float4* jla;
    const float2* jparam;
    float2* jp;
    float4 *jr, *jm;
    float2* jbias;
    const int* jorder;               // per joint slot (CSR like joints): colour-major joint index, world-local
    const int* jcolor_start;         // per world: MAX_JC + 1 entries
    const int* n_jcolors;            // per world
    // persistent per-world slot cache (cap entries per world)
    int* c_count;
    unsigned* c_key;
    float4* c_cache[2];              // (friction, pn, pt, -) of contact 0 per slot, double buffered: a step reads [parity], writes [parity ^ 1]
    int* c_par;                      // per world: which buffer holds the last finished step
    float4 *c_stg0, *c_stg1;         // narrow-phase staging per slot: (px0,py0,sep0,nx), (px1,py1,sep1,ny)
    unsigned char* c_color;          // NO_COLOR = inactive slot (candidate without contacts)
    // per-world summary
    int *w_err, *w_ncon, *w_narb, *w_ncolors, *w_errjoint;
    float* w_maxpen;
    float4* w_badk;
    // optional host-format records (sylt_batch_step_host*): the kernel reads its world's bodies from `rec_in` instead of pose / vel
    // and also writes the result to `rec_out` — rec_floats = 6 (x, y, rotation, vx, vy, w) or 9 (SyltBodyState) floats per body,
    // indexed by the global body index: no separate unpack / pack launches around the step
    const float* rec_in;
    float* rec_out;
    int rec_floats;
    SlotExport* exp_slots;           // export slots * cap
    int* exp_count;                  // per export slot
    // shared-memory layout: byte offset of every Smem array, computed once on the host (carve) and read from the constant
    // bank, so that re-deriving a pointer inside a hot loop is one load + one add instead of the whole chain of align-ups
    unsigned long long smem_off[48];
};

struct Smem {
    float2* pos; float* rot; float2* cs; float4* vel; float2* inv; float4* aabb; unsigned char* stat;
    unsigned* used;
    int* rowcnt;                       // Bmax + 1
    float4* caabb; int* bcell; int* bstart; unsigned short* cellb; unsigned short* rank; unsigned short* rowtmp; unsigned short* large;
    // slots
    unsigned* key; unsigned char* cnt; unsigned char* color;
    unsigned short* order;             // colour-major position k -> slot
    // solver records, indexed by colour-major position k so that the lanes of a warp read consecutive addresses:
    //   kb = body1 | body2 << 12 | (contacts - 1) << 24;  per contact q: cA[q] = (r1, r2), cB[q] = (mass_normal, mass_tangent, bias, pn)
    unsigned* kb; float2* nrmk; float* frk; float4* cA[2]; float4* cB[2]; float2* ptk;
    // old cache
    unsigned* okey; unsigned char* ocolor;  // last step's slots: keys + colours here, (friction, pn0, pt0) stay in global memory
    // joints
    int2* jb; float4* jla; float2* jparam; float2* jp; float4* jr; float4* jm; float2* jbias; unsigned short* jorder;
    int* misc;                         // scan scratch (40) + colour bins
};

__host__ __device__ inline size_t align16(size_t x) { return (x + 15) & ~(size_t)15; }

// identical carve-up on host (size only) and device
__host__ __device__ inline size_t carve(Smem* s, unsigned char* base, int B, int J, int cap, int nbk) {
    size_t o = 0;
#define TAKE(field, type, n)                          \
    do {                                              \
        if (s) s->field = (type*)(base + o);          \
        o = align16(o + sizeof(type) * (size_t)(n)); \
    } while (0)
    TAKE(pos, float2, B); TAKE(rot, float, B); TAKE(cs, float2, B); TAKE(vel, float4, B); TAKE(inv, float2, B);
    TAKE(stat, unsigned char, B); TAKE(used, unsigned, B); TAKE(large, unsigned short, B + 1);
    TAKE(key, unsigned, cap); TAKE(kb, unsigned, cap); TAKE(nrmk, float2, cap); TAKE(cnt, unsigned char, cap); TAKE(color, unsigned char, cap);
    // the contact-constant region doubles as broad-phase scratch (dead before the narrow phase writes contacts)
    const size_t scratch = 2 * align16(sizeof(float4) * (size_t)B) + align16(sizeof(int) * (size_t)(B + 1)) + align16(sizeof(int) * (size_t)B) +
                           align16(sizeof(int) * (size_t)(nbk + 1)) + 2 * align16(sizeof(unsigned short) * (size_t)B) +
                           align16(sizeof(unsigned short) * (size_t)B * ROWT);
    const size_t con_floats = (size_t)18 * cap > scratch / 4 ? (size_t)18 * cap : scratch / 4;
    const size_t con_at = o;
    if (s) {
        float4* f = (float4*)(base + o);
        s->cA[0] = f; s->cA[1] = f + cap; s->cB[0] = f + 2 * (size_t)cap; s->cB[1] = f + 3 * (size_t)cap; s->ptk = (float2*)(f + 4 * (size_t)cap);
    }
    o = align16(o + sizeof(float) * con_floats);
    TAKE(order, unsigned short, cap);
    TAKE(okey, unsigned, cap); TAKE(ocolor, unsigned char, cap);
    if (s) s->frk = (float*)s->okey;  // last step's keys are dead between the narrow phase and the end of the step
    TAKE(jb, int2, J + 1); TAKE(jla, float4, J + 1); TAKE(jparam, float2, J + 1); TAKE(jp, float2, J + 1); TAKE(jr, float4, J + 1);
    TAKE(jm, float4, J + 1); TAKE(jbias, float2, J + 1); TAKE(jorder, unsigned short, J + 1);
    TAKE(misc, int, 40 + 3 * BNC + 8);
    const size_t total = o;
    o = con_at;
    TAKE(aabb, float4, B); TAKE(caabb, float4, B); TAKE(rowcnt, int, B + 1); TAKE(bcell, int, B); TAKE(bstart, int, nbk + 1); TAKE(cellb, unsigned short, B);
    TAKE(rank, unsigned short, B); TAKE(rowtmp, unsigned short, B * ROWT);
    o = total;
#undef TAKE
    return o;
}

__device__ __forceinline__ int warp_incl_scan_i(int v) {
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        int u = __shfl_up_sync(0xffffffffu, v, o);
        if ((threadIdx.x & 31) >= o) v += u;
    }
    return v;
}
// in-place exclusive scan of a[0..n) in shared memory by the whole block; returns the total
__device__ int block_scan_inplace(int* a, int n, int* sh /* 34 ints */) {
    const int T = blockDim.x, tid = threadIdx.x;
    const int chunk = (n + T - 1) / T;
    const int s0 = min(tid * chunk, n), s1 = min(s0 + chunk, n);
    int s = 0;
    for (int k = s0; k < s1; k++) s += a[k];
    int inc = warp_incl_scan_i(s);
    const int w = tid >> 5, l = tid & 31;
    if (l == 31) sh[w] = inc;
    __syncthreads();
    if (w == 0) {
        int v = l < (T >> 5) ? sh[l] : 0;
        int vi = warp_incl_scan_i(v);
        sh[l] = vi - v;
        if (l == 31) sh[32] = vi;
    }
    __syncthreads();
    int run = sh[w] + inc - s;
    for (int k = s0; k < s1; k++) { int t = a[k]; a[k] = run; run += t; }
    int total = sh[32];
    __syncthreads();
    return total;
}

// slot key = owner << 16 | partner (slot order); body1 = lower index (Arbiter::new, arbiter.rs:120-122)
__device__ __forceinline__ void key_bodies(unsigned key, int& b1, int& b2) {
    int x = (int)(key >> 16), y = (int)(key & 0xffffu);
    b1 = min(x, y); b2 = max(x, y);
}
__device__ __forceinline__ int cellc(float x, float cell) { return (int)fminf(fmaxf(floorf(x / cell), -1.0e9f), 1.0e9f); }
// bucket of cell (cx, cy): rows of cells are contiguous runs of buckets (mod the table size), so a row of a search window is
// ONE range of the cell-ordered body list.  Two cells of a window (<= 3 x 3) never share a bucket: k * CELL_ROW_STRIDE mod 2^m
// stays outside [-3, 3] for k = 1, 2, 3 and every table size 2^m >= 64.
constexpr unsigned CELL_ROW_STRIDE = 37u;
__device__ __forceinline__ unsigned cellh(int cx, int cy) { return (unsigned)cx + (unsigned)cy * CELL_ROW_STRIDE; }

__device__ __forceinline__ bool overlap4(float4 a, float4 b) { return !(a.x > b.z || b.x > a.z || a.y > b.w || b.y > a.w); }

__device__ __forceinline__ Vel ld_vel(const float4* v, int b) { float4 t = v[b]; Vel o; o.v = mk(t.x, t.y); o.w = t.z; return o; }
__device__ __forceinline__ void st_vel(float4* v, int b, const Vel& x) { v[b] = make_float4(x.v.x, x.v.y, x.w, 0.0f); }

// The iterated part of the solver (world.rs:132-140, Arbiter::apply_impulse) for `iters` sweeps over the arbiter colours.
// Deliberately NOT inlined: the step kernel around it keeps ~60 values alive, so inlined the compiler re-derives all ten
// shared-memory base addresses on every colour hop; as a function of its own it keeps them in registers for the whole sweep.
// (Byte offsets into the dynamic shared memory, not pointers: across a call the compiler would otherwise fall back to generic
// 64-bit addressing.)
struct HotOffsets {
    unsigned kb, inv, nrmk, frk, ptk, vel, cA0, cA1, cB0, cB1, cstart;
};
struct HotArrays {
    const unsigned* kb; const float2* inv; const float2* nrmk; const float* frk; float2* ptk; float4* vel;
    const float4 *cA0, *cA1; float4 *cB0, *cB1;
    const int* cstart;
};
extern __shared__ __align__(16) unsigned char smem_raw[];
template <bool ACC>
__device__ __noinline__ void contact_sweeps(HotOffsets o, int ncolors, int iters) {
    const int tid = threadIdx.x, T = blockDim.x;
    HotArrays h;
    h.kb = (const unsigned*)(smem_raw + o.kb); h.inv = (const float2*)(smem_raw + o.inv); h.nrmk = (const float2*)(smem_raw + o.nrmk);
    h.frk = (const float*)(smem_raw + o.frk); h.ptk = (float2*)(smem_raw + o.ptk); h.vel = (float4*)(smem_raw + o.vel);
    h.cA0 = (const float4*)(smem_raw + o.cA0); h.cA1 = (const float4*)(smem_raw + o.cA1);
    h.cB0 = (float4*)(smem_raw + o.cB0); h.cB1 = (float4*)(smem_raw + o.cB1); h.cstart = (const int*)(smem_raw + o.cstart);
    for (int it = 0; it < iters; it++) {
        for (int c = 0; c < ncolors; c++) {
            for (int k = h.cstart[c] + tid; k < h.cstart[c + 1]; k += T) {
                const unsigned kbv = h.kb[k];
                const int b1 = kbv & 0xfffu, b2 = (kbv >> 12) & 0xfffu;
                const float4 ca0 = h.cA0[k], cb0 = h.cB0[k], ca1 = h.cA1[k], cb1 = h.cB1[k];  // contact 1: unused garbage when absent
                const float2 i1 = h.inv[b1], i2 = h.inv[b2], nn = h.nrmk[k];
                const float fr = h.frk[k];
                float2 pt = h.ptk[k];
                Vel v1 = ld_vel(h.vel, b1), v2 = ld_vel(h.vel, b2);
                ContactConst k0;
                k0.n = mk(nn.x, nn.y); k0.r1 = mk(ca0.x, ca0.y); k0.r2 = mk(ca0.z, ca0.w);
                k0.mass_n = cb0.x; k0.mass_t = cb0.y; k0.bias = cb0.z;
                float pn = cb0.w;
                contact_apply(k0, fr, i1.x, i1.y, i2.x, i2.y, ACC, pn, pt.x, v1, v2);
                if (ACC) h.cB0[k].w = pn;
                if (kbv >> 24) {
                    k0.r1 = mk(ca1.x, ca1.y); k0.r2 = mk(ca1.z, ca1.w);
                    k0.mass_n = cb1.x; k0.mass_t = cb1.y; k0.bias = cb1.z;
                    pn = cb1.w;
                    contact_apply(k0, fr, i1.x, i1.y, i2.x, i2.y, ACC, pn, pt.y, v1, v2);
                    if (ACC) h.cB1[k].w = pn;
                }
                if (ACC) h.ptk[k] = pt;
                if (i1.x != 0.0f || i1.y != 0.0f) st_vel(h.vel, b1, v1);
                if (i2.x != 0.0f || i2.y != 0.0f) st_vel(h.vel, b2, v2);
            }
            __syncthreads();
        }
    }
}

// The ordered half of Arbiter::pre_step (arbiter.rs:216-223): apply last step's accumulated impulses, colour by colour.
__device__ __noinline__ void warm_sweep(HotOffsets o, int ncolors) {
    const int tid = threadIdx.x, T = blockDim.x;
    const unsigned* kb = (const unsigned*)(smem_raw + o.kb);
    const float2 *inv = (const float2*)(smem_raw + o.inv), *nrmk = (const float2*)(smem_raw + o.nrmk), *ptk = (const float2*)(smem_raw + o.ptk);
    float4* vel = (float4*)(smem_raw + o.vel);
    const float4 *cA0 = (const float4*)(smem_raw + o.cA0), *cA1 = (const float4*)(smem_raw + o.cA1);
    const float4 *cB0 = (const float4*)(smem_raw + o.cB0), *cB1 = (const float4*)(smem_raw + o.cB1);
    const int* cstart = (const int*)(smem_raw + o.cstart);
    for (int c = 0; c < ncolors; c++) {
        for (int k = cstart[c] + tid; k < cstart[c + 1]; k += T) {
            const unsigned kbv = kb[k];
            const int b1 = kbv & 0xfffu, b2 = (kbv >> 12) & 0xfffu;
            const float2 i1 = inv[b1], i2 = inv[b2], nn = nrmk[k], pt = ptk[k];
            Vel v1 = ld_vel(vel, b1), v2 = ld_vel(vel, b2);
            ContactConst k0;
            k0.n = mk(nn.x, nn.y);
            float4 ca = cA0[k];
            k0.r1 = mk(ca.x, ca.y); k0.r2 = mk(ca.z, ca.w);
            contact_warmstart(k0, i1.x, i1.y, i2.x, i2.y, cB0[k].w, pt.x, v1, v2);
            if (kbv >> 24) {
                ca = cA1[k];
                k0.r1 = mk(ca.x, ca.y); k0.r2 = mk(ca.z, ca.w);
                contact_warmstart(k0, i1.x, i1.y, i2.x, i2.y, cB1[k].w, pt.y, v1, v2);
            }
            if (i1.x != 0.0f || i1.y != 0.0f) st_vel(vel, b1, v1);
            if (i2.x != 0.0f || i2.y != 0.0f) st_vel(vel, b2, v2);
        }
        __syncthreads();
    }
}

#ifdef SYLT_BATCH_TIMING
// profiling build only (tools/profiling): cycles thread 0 of every CTA spends in each phase of the step
__device__ unsigned long long g_phase_cycles[16];
#define PHASE(i) do { if (tid == 0) { long long t_ = clock64(); atomicAdd(&g_phase_cycles[i], (unsigned long long)(t_ - t_last)); t_last = t_; } } while (0)
// inside a hop (-DSYLT_BATCH_TIMING=2; the probes themselves cost ~40 cycles each): the clock is read after `dep` is available
#if SYLT_BATCH_TIMING < 2
#define HOP(i, dep) do { } while (0)
#else
#define HOP(i, dep) do { if (tid == 0) { long long t_; asm volatile("mov.u64 %0, %%clock64;" : "=l"(t_) : "f"(dep) : "memory"); \
                                         atomicAdd(&g_phase_cycles[i], (unsigned long long)(t_ - t_hop)); t_hop = t_; } } while (0)
#endif
#else
#define PHASE(i) do { } while (0)
#define HOP(i, dep) do { } while (0)
#endif

__global__ void __launch_bounds__(256, 4) k_batch_step(BatchArgs a) {
    const int wld = blockIdx.x + a.world_begin;
    if (wld >= a.n_worlds) return;
    const int tid = threadIdx.x, T = blockDim.x, lane = tid & 31, warp = tid >> 5;
    const int b0 = a.body_off[wld], B = a.body_off[wld + 1] - b0;
    const int j0 = a.joint_off[wld], J = a.joint_off[wld + 1] - j0;
    const int cap = a.cap;
    Smem s;
    static_assert(sizeof(Smem) % sizeof(void*) == 0 && sizeof(Smem) / sizeof(void*) <= 48, "Smem must be an array of pointers");
    {
        unsigned char* ptrs[sizeof(Smem) / sizeof(void*)];
#pragma unroll
        for (int i = 0; i < (int)(sizeof(Smem) / sizeof(void*)); i++) ptrs[i] = smem_raw + a.smem_off[i];
        memcpy(&s, ptrs, sizeof(Smem));  // (not a cast: the fields have different pointer types)
    }
    int* scan_sh = s.misc;           // 34
    int* s_P = s.misc + 36;          // current candidate count
    int* s_Pold = s.misc + 37;
    int* s_err = s.misc + 38;
    int* ccount = s.misc + 40;       // BNC
    int* cstart = ccount + BNC;      // BNC + 1
    int* cfill = cstart + BNC + 1;   // BNC

    // ---- load the world
    for (int i = tid; i < B; i += T) {
        float4 p, v;
        const float4 m = a.mp[b0 + i];
        if (a.rec_in) {
            const float* r = a.rec_in + (size_t)(b0 + i) * a.rec_floats;
            p = make_float4(r[0], r[1], r[2], 0.0f);
            v = make_float4(r[3], r[4], r[5], 0.0f);
        } else {
            p = a.pose[b0 + i]; v = a.vel[b0 + i];
        }
        s.pos[i] = make_float2(p.x, p.y); s.rot[i] = p.z; s.vel[i] = v; s.inv[i] = make_float2(m.x, m.y);
        s.stat[i] = a.bflags[b0 + i];
    }
    for (int j = tid; j < J; j += T) {
        s.jb[j] = a.jbodies[j0 + j]; s.jla[j] = a.jla[j0 + j]; s.jparam[j] = a.jparam[j0 + j]; s.jp[j] = a.jp[j0 + j];
        s.jorder[j] = (unsigned short)a.jorder[j0 + j];
        s.jr[j] = make_float4(0, 0, 0, 0); s.jm[j] = make_float4(0, 0, 0, 0); s.jbias[j] = make_float2(0, 0);
    }
    int par = a.c_par[wld];
    const size_t gslot = (size_t)wld * cap;
    const int Pold0 = min(a.c_count[wld], cap);
    for (int k = tid; k < Pold0; k += T) {
        size_t g = (size_t)wld * cap + k;
        s.okey[k] = a.c_key[g]; s.ocolor[k] = a.c_color[g];
    }
    if (tid == 0) {
        *s_Pold = Pold0; *s_err = a.w_err[wld]; *s_P = 0; s.misc[138] = 0; s.misc[139] = 0;
        int nl = 0;
        for (int i = 0; i < B; i++)
            if (a.bflags[b0 + i] & 2) s.large[nl++] = (unsigned short)i;
        s.misc[34] = nl;
    }
    const int njc = a.n_jcolors[wld];
    const int* jcs = a.jcolor_start + (size_t)wld * (MAX_JC + 1);
    __syncthreads();
    const bool acc = a.accumulate != 0, warm = a.warm != 0, poscorr = a.poscorr != 0;
    const float kbias = poscorr ? 0.2f : 0.0f;
    int ncolors = 0, P = 0;
    float maxpen = 0.0f;
    const int xw = a.exp_map[wld];
    const bool exporting = xw >= 0;

#ifdef SYLT_BATCH_TIMING
    long long t_last = clock64(), t_hop = 0;
#endif
    for (int step = 0; step < a.n_steps; step++) {
        if (*s_err) break;  // a world that raised an error stays frozen (block-uniform: read after a barrier)
        maxpen = 0.0f;
        PHASE(0);
        // ---- 1. bodies: rotation cache, AABB, force integration (world.rs:113-120; no external forces in a batch)
        float ext_max = 0.0f, abs_max = 0.0f;
        for (int i = tid; i < B; i += T) {
            float r = s.rot[i];
            float c = sylt_cosf(r), sn = sylt_sinf(r);
            s.cs[i] = make_float2(c, sn);
            float2 p = s.pos[i], w = a.wh[b0 + i];
            float hx = w.x * 0.5f, hy = w.y * 0.5f;
            float ex = fabsf(c) * hx + fabsf(sn) * hy, ey = fabsf(sn) * hx + fabsf(c) * hy;
            float mg = 1.0e-6f * (fabsf(p.x) + fabsf(p.y)) + 2.0e-5f * (ex + ey) + 1.0e-6f;
            const float4 ab = make_float4(p.x - ex - mg, p.y - ey - mg, p.x + ex + mg, p.y + ey + mg);
            s.aabb[i] = ab;
            if (!(s.stat[i] & 2)) {
                ext_max = fmaxf(ext_max, fmaxf(ab.z - ab.x, ab.w - ab.y));
                abs_max = fmaxf(abs_max, fmaxf(fabsf(p.x), fabsf(p.y)));
            }
            float2 iv = s.inv[i];
            if (iv.x != 0.0f) {
                float4 v = s.vel[i];
                V2 nv = mk(v.x, v.y) + (mk(a.gx, a.gy) + mk(0.0f, 0.0f) * iv.x) * a.dt;
                v.x = nv.x; v.y = nv.y;
                v.z += iv.y * 0.0f * a.dt;
                s.vel[i] = v;
            }
            s.used[i] = 0u;
        }
        for (int c = tid; c < 3 * BNC + 1; c += T) ccount[c] = 0;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            ext_max = fmaxf(ext_max, __shfl_xor_sync(0xffffffffu, ext_max, o));
            abs_max = fmaxf(abs_max, __shfl_xor_sync(0xffffffffu, abs_max, o));
        }
        if (lane == 0) {  // non-negative floats order like their bit patterns
            atomicMax(&s.misc[138], __float_as_int(fminf(ext_max, 3.0e38f)));
            atomicMax(&s.misc[139], __float_as_int(fminf(abs_max, 3.0e38f)));
        }
        __syncthreads();
        PHASE(1);
        // ---- 2. candidate rows through a hashed uniform grid in shared memory (large bodies bypass it):
        //         bucket counts -> scan -> fill; thread per body: count (remembering partners) -> scan -> emit sorted.
        //         A pair is owned by its lower index, a small-large pair by the small body (slot order = (owner, partner)).
        //         The cell is re-derived every step from the largest AABB the small bodies have NOW (a resting box spans
        //         its width, not its bounding circle), with slack for the rounding of centres at large coordinates.
        const int nbk = a.nbk, nl = s.misc[34];
        const unsigned bmask = (unsigned)nbk - 1u;
        const float cell = fminf(a.cell, 1.05f * __int_as_float(s.misc[138]) + 0.01f + 1.0e-6f * __int_as_float(s.misc[139]));
        for (int k = tid; k <= nbk; k += T) s.bstart[k] = 0;
        __syncthreads();
        for (int i = tid; i < B; i += T) {
            if (s.stat[i] & 2) continue;
            float4 ab = s.aabb[i];
            int h = (int)(cellh(cellc(0.5f * (ab.x + ab.z), cell), cellc(0.5f * (ab.y + ab.w), cell)) & bmask);
            s.bcell[i] = h;
            s.rank[i] = (unsigned short)atomicAdd(&s.bstart[h], 1);
        }
        __syncthreads();
        PHASE(9);
        block_scan_inplace(s.bstart, nbk + 1, scan_sh);
        for (int i = tid; i < B; i += T) {
            if (s.stat[i] & 2) continue;
            int k = s.bstart[s.bcell[i]] + s.rank[i];
            s.cellb[k] = (unsigned short)(i | ((s.stat[i] & 1) << 15));  // body index (<= 4095) + static flag
            s.caabb[k] = s.aabb[i];
        }
        __syncthreads();
        PHASE(10);
        auto row_search = [&](int i, auto&& found) {
            const float4 my = s.aabb[i];
            const int fl = s.stat[i];
            if (!(fl & 2)) {
                const int x0 = cellc(my.x - 0.5f * cell, cell), x1 = cellc(my.z + 0.5f * cell, cell);
                const int y0 = cellc(my.y - 0.5f * cell, cell), y1 = cellc(my.w + 0.5f * cell, cell);
                const unsigned skip = (fl & 1) ? 0x8000u : 0u;  // static-static pairs are never candidates
                auto scan = [&](int k0, int k1) {
                    for (int k = k0; k < k1; k++) {
                        const unsigned e = s.cellb[k];
                        const int j = (int)(e & 0x7fffu);
                        if (j <= i || (e & skip)) continue;
                        if (overlap4(my, s.caabb[k])) found(j);
                    }
                };
                if (x1 - x0 <= 3 && y1 - y0 <= 3) {  // always, unless coordinates are so large that the AABB margin outgrows the cell
                    const unsigned wx = (unsigned)(x1 - x0);
                    for (int cy = y0; cy <= y1; cy++) {
                        const unsigned hb = cellh(x0, cy);
                        const int h0 = (int)(hb & bmask), h1 = (int)((hb + wx) & bmask);
                        if (h1 >= h0) scan(s.bstart[h0], s.bstart[h1 + 1]);
                        else { scan(s.bstart[h0], s.bstart[nbk]); scan(s.bstart[0], s.bstart[h1 + 1]); }
                    }
                } else {
                    scan(0, s.bstart[nbk]);
                }
                for (int k = 0; k < nl; k++) {
                    int j = s.large[k];
                    if ((fl & 1) && (s.stat[j] & 1)) continue;
                    if (overlap4(my, s.aabb[j])) found(j);
                }
            } else {
                for (int k = 0; k < nl; k++) {
                    int j = s.large[k];
                    if (j <= i) continue;
                    if ((fl & 1) && (s.stat[j] & 1)) continue;
                    if (overlap4(my, s.aabb[j])) found(j);
                }
            }
        };
        for (int i = tid; i < B; i += T) {
            int n = 0;
            unsigned short* tmp = s.rowtmp + i * ROWT;
            row_search(i, [&](int j) { if (n < ROWT) tmp[n] = (unsigned short)j; n++; });
            s.rowcnt[i] = n;
        }
        __syncthreads();
        PHASE(11);
        P = block_scan_inplace(s.rowcnt, B, scan_sh);
        PHASE(12);
        if (P > cap) {  // block-uniform
            if (tid == 0) *s_err = SYLT_ERR_CAPACITY;
            __syncthreads();
            break;
        }
        for (int i = tid; i < B; i += T) {
            const int base = s.rowcnt[i], n = (i + 1 < B ? s.rowcnt[i + 1] : P) - base;
            if (n == 0) continue;
            if (n <= ROWT) {
                unsigned short v[ROWT];
                const unsigned short* tmp = s.rowtmp + i * ROWT;
#pragma unroll
                for (int q = 0; q < ROWT; q++) v[q] = q < n ? tmp[q] : (unsigned short)0xffff;
#pragma unroll
                for (int q = 1; q < ROWT; q++) {  // insertion sort by partner
                    unsigned short x = v[q];
                    int r = q - 1;
                    while (r >= 0 && v[r] > x) { v[r + 1] = v[r]; r--; }
                    v[r + 1] = x;
                }
                for (int q = 0; q < n; q++) s.key[base + q] = ((unsigned)i << 16) | (unsigned)v[q];
            } else {
                int m = 0;
                row_search(i, [&](int j) { s.key[base + m++] = ((unsigned)i << 16) | (unsigned)j; });
                for (int q = 1; q < n; q++) {
                    unsigned x = s.key[base + q];
                    int r = q - 1;
                    while (r >= 0 && s.key[base + r] > x) { s.key[base + r + 1] = s.key[base + r]; r--; }
                    s.key[base + r + 1] = x;
                }
            }
        }
        __syncthreads();
        PHASE(2);
        // ---- 3. narrow phase + match with last step's slots, one thread per candidate slot
        const int Pold = *s_Pold;
        for (int p = tid; p < P; p += T) {
            unsigned key = s.key[p];
            int b1, b2;
            key_bodies(key, b1, b2);
            ContactOut out[2];
            float2 p1 = s.pos[b1], p2 = s.pos[b2], c1 = s.cs[b1], c2 = s.cs[b2], w1 = a.wh[b0 + b1], w2 = a.wh[b0 + b2];
            int n = box_box(mk(p1.x, p1.y), c1.x, c1.y, mk(w1.x, w1.y), mk(p2.x, p2.y), c2.x, c2.y, mk(w2.x, w2.y), out);
            s.cnt[p] = (unsigned char)n;
            unsigned char color = NO_COLOR;
            if (n > 0) {
                int lo = 0, hi = Pold;  // first old slot with key >= this key
                while (lo < hi) { int mid = (lo + hi) >> 1; if (s.okey[mid] < key) lo = mid + 1; else hi = mid; }
                float fr, pn = 0.0f, pt = 0.0f;
                if (lo < Pold && s.okey[lo] == key && s.ocolor[lo] != NO_COLOR) {  // Arbiter::update (arbiter.rs:137-181)
                    const float4 oc = a.c_cache[par][gslot + lo];  // written by this CTA at the end of the previous step
                    fr = oc.x;                     // quirk Q-17
                    color = s.ocolor[lo];
                    if (warm) { pn = oc.y; pt = oc.z; }  // quirk Q-2: all contacts inherit old contact 0
                } else {
                    fr = sqrtf(a.mp[b0 + b1].z * a.mp[b0 + b2].z);  // arbiter.rs:128
                }
                // staged through (L2-resident) global memory: the solver records are laid out in colour order, which is only
                // known after the new arbiters are coloured
                a.c_stg0[gslot + p] = make_float4(out[0].px, out[0].py, out[0].sep, out[0].nx);
                a.c_stg1[gslot + p] = n > 1 ? make_float4(out[1].px, out[1].py, out[1].sep, out[0].ny) : make_float4(0.0f, 0.0f, 0.0f, out[0].ny);
                a.c_cache[par ^ 1][gslot + p] = make_float4(fr, pn, pt, 0.0f);
                if (color != NO_COLOR) {
                    if (!(s.stat[b1] & 1)) atomicOr(&s.used[b1], 1u << color);
                    if (!(s.stat[b2] & 1)) atomicOr(&s.used[b2], 1u << color);
                }
            }
            s.color[p] = color;
            if (exporting && step == a.n_steps - 1) {
                SlotExport* e = &a.exp_slots[(size_t)xw * cap + p];
                { int kb1, kb2; key_bodies(key, kb1, kb2); e->key = ((unsigned)kb1 << 16) | (unsigned)kb2; }
                e->count = n; e->nx = n > 0 ? out[0].nx : 0.0f; e->ny = n > 0 ? out[0].ny : 0.0f;
                for (int q = 0; q < 2; q++) {
                    e->geo[q][0] = q < n ? out[q].px : 0.0f; e->geo[q][1] = q < n ? out[q].py : 0.0f; e->geo[q][2] = q < n ? out[q].sep : 0.0f;
                    e->feat[q] = q < n ? out[q].feat : 0u;
                }
            }
        }
        __syncthreads();
        PHASE(3);
        // ---- 4. colour the new arbiters: lowest free colour, in key order.  Every warp flags its slots into a bit mask
        //         (parked in the not yet written `order` array); lane 0 of warp 0 walks the set bits — a settled world has few.
        unsigned* newmask = reinterpret_cast<unsigned*>(s.order);
        const int nwords = (P + 31) >> 5;
        for (int pb = warp << 5; pb < P; pb += T) {
            const int p = pb + lane;
            const bool need = p < P && s.cnt[p] > 0 && s.color[p] == NO_COLOR;
            const unsigned m = __ballot_sync(0xffffffffu, need);
            if (lane == 0) newmask[pb >> 5] = m;
        }
        __syncthreads();
        if (warp == 0) {
            for (int wb = 0; wb < nwords; wb += 32) {
                const unsigned mine = wb + lane < nwords ? newmask[wb + lane] : 0u;
                unsigned nz = __ballot_sync(0xffffffffu, mine != 0u);
                while (nz) {  // warp-uniform
                    const int wi = __ffs(nz) - 1;
                    nz &= nz - 1;
                    unsigned m = __shfl_sync(0xffffffffu, mine, wi);
                    if (lane == 0) {
                        while (m) {
                            const int q = ((wb + wi) << 5) + __ffs(m) - 1;
                            m &= m - 1;
                            int b1, b2;
                            key_bodies(s.key[q], b1, b2);
                            const unsigned um = ((s.stat[b1] & 1) ? 0u : s.used[b1]) | ((s.stat[b2] & 1) ? 0u : s.used[b2]);
                            const unsigned fr = ~um;
                            if (!fr) { *s_err = SYLT_ERR_CAPACITY; continue; }
                            const int c = __ffs(fr) - 1;
                            s.color[q] = (unsigned char)c;
                            if (!(s.stat[b1] & 1)) s.used[b1] |= 1u << c;
                            if (!(s.stat[b2] & 1)) s.used[b2] |= 1u << c;
                        }
                    }
                    __syncwarp();
                }
            }
        }
        __syncthreads();
        if (*s_err) break;
        //         colour-major order: per-colour counts -> starts -> fill.  Lanes of a warp that share a colour are grouped with
        //         match.any, so one shared-memory atomic per (warp, colour) replaces one per arbiter on 5-6 hot addresses.
        for (int pb = warp << 5; pb < P; pb += T) {
            const int p = pb + lane;
            const unsigned col = (p < P && s.cnt[p] > 0) ? (unsigned)s.color[p] : (unsigned)NO_COLOR;
            const unsigned grp = __match_any_sync(0xffffffffu, col);
            if (col != NO_COLOR && (__ffs(grp) - 1) == lane) atomicAdd(&ccount[col], __popc(grp));
        }
        __syncthreads();
        if (warp == 0) {
            const int cntc = ccount[lane];  // BNC == 32 == warp size
            const int incl = warp_incl_scan_i(cntc);
            cstart[lane] = incl - cntc;
            if (lane == 31) cstart[BNC] = incl;
            const unsigned present = __ballot_sync(0xffffffffu, cntc > 0);
            int mx = cntc;
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) mx = max(mx, __shfl_xor_sync(0xffffffffu, mx, o));
            if (lane == 0) { s.misc[35] = 32 - __clz(present); s.misc[137] = mx; }
        }
        __syncthreads();
        ncolors = s.misc[35];
        for (int pb = warp << 5; pb < P; pb += T) {
            const int p = pb + lane;
            const unsigned col = (p < P && s.cnt[p] > 0) ? (unsigned)s.color[p] : (unsigned)NO_COLOR;
            const unsigned grp = __match_any_sync(0xffffffffu, col);
            const int leader = __ffs(grp) - 1;
            int base = 0;
            if (col != NO_COLOR && leader == lane) base = atomicAdd(&cfill[col], __popc(grp));
            base = __shfl_sync(0xffffffffu, base, leader);
            if (col != NO_COLOR) s.order[cstart[col] + base + __popc(grp & ((1u << lane) - 1u))] = (unsigned short)p;
        }
        __syncthreads();
        PHASE(4);
        // ---- 5. Arbiter::pre_step.  (a) the velocity-independent half (arbiter.rs:192-215), one thread per arbiter, written as
        //         colour-ordered solver records;  (b) the ordered half (arbiter.rs:216-223): last step's impulses, colour by colour
        const int A = cstart[BNC];
        HotOffsets hot;
        {
            auto off = [&](const void* p) { return (unsigned)((const unsigned char*)p - smem_raw); };
            hot.kb = off(s.kb); hot.inv = off(s.inv); hot.nrmk = off(s.nrmk); hot.frk = off(s.frk); hot.ptk = off(s.ptk); hot.vel = off(s.vel);
            hot.cA0 = off(s.cA[0]); hot.cA1 = off(s.cA[1]); hot.cB0 = off(s.cB[0]); hot.cB1 = off(s.cB[1]); hot.cstart = off(cstart);
        }
        for (int k = tid; k < A; k += T) {
            const int p = s.order[k];
            int b1, b2;
            key_bodies(s.key[p], b1, b2);
            const int n = s.cnt[p];
            const float4 g0 = a.c_stg0[gslot + p], g1 = a.c_stg1[gslot + p], cw = a.c_cache[par ^ 1][gslot + p];
            const float2 p1 = s.pos[b1], p2 = s.pos[b2], i1 = s.inv[b1], i2 = s.inv[b2];
            const V2 nn = mk(g0.w, g1.w);
            ContactConst k0;
            contact_constants(mk(g0.x, g0.y), nn, g0.z, mk(p1.x, p1.y), mk(p2.x, p2.y), i1.x, i1.y, i2.x, i2.y, a.inv_dt, kbias, k0);
            maxpen = fmaxf(maxpen, -g0.z);
            s.cA[0][k] = make_float4(k0.r1.x, k0.r1.y, k0.r2.x, k0.r2.y);
            s.cB[0][k] = make_float4(k0.mass_n, k0.mass_t, k0.bias, cw.y);
            if (n > 1) {
                contact_constants(mk(g1.x, g1.y), nn, g1.z, mk(p1.x, p1.y), mk(p2.x, p2.y), i1.x, i1.y, i2.x, i2.y, a.inv_dt, kbias, k0);
                maxpen = fmaxf(maxpen, -g1.z);
                s.cA[1][k] = make_float4(k0.r1.x, k0.r1.y, k0.r2.x, k0.r2.y);
                s.cB[1][k] = make_float4(k0.mass_n, k0.mass_t, k0.bias, cw.y);
            }
            s.kb[k] = (unsigned)b1 | ((unsigned)b2 << 12) | ((unsigned)(n - 1) << 24);
            s.nrmk[k] = make_float2(nn.x, nn.y);
            s.frk[k] = cw.x;
            s.ptk[k] = make_float2(cw.z, cw.z);
        }
        __syncthreads();
        if (acc) warm_sweep(hot, ncolors);
        PHASE(5);
        // ---- 6. Joint::pre_step by joint colour.  The reference's loop (world.rs:127-129) returns Err at the FIRST joint in Vec
        //         order whose K is singular: joints behind it are never touched.  K depends on the poses only, so that joint is
        //         found before anything is pre-stepped (s.misc[140] = J - index, 0 = none).
        if (tid == 0) s.misc[140] = 0;
        __syncthreads();
        for (int j = tid; j < J; j += T) {
            const int2 b = s.jb[j];
            const float2 i1 = s.inv[b.x], i2 = s.inv[b.y], c1 = s.cs[b.x], c2 = s.cs[b.y];
            const float4 la = s.jla[j];
            if (joint_singular(mk(la.x, la.y), mk(la.z, la.w), s.jparam[j].x, c1.x, c1.y, c2.x, c2.y, i1.x, i1.y, i2.x, i2.y)) atomicMax(&s.misc[140], J - j);
        }
        __syncthreads();
        const int jfail = s.misc[140] > 0 ? J - s.misc[140] : 0x7fffffff;
        for (int c = 0; c < njc; c++) {
            for (int k = jcs[c] + tid; k < jcs[c + 1]; k += T) {
                int j = s.jorder[k];
                if (j > jfail) continue;
                int2 b = s.jb[j];
                float2 i1 = s.inv[b.x], i2 = s.inv[b.y], p1 = s.pos[b.x], p2 = s.pos[b.y], c1 = s.cs[b.x], c2 = s.cs[b.y];
                Vel v1 = ld_vel(s.vel, b.x), v2 = ld_vel(s.vel, b.y);
                float4 la = s.jla[j];
                float2 pr = s.jparam[j], pa = s.jp[j];
                V2 pacc = mk(pa.x, pa.y);
                JointConst jc;
                M2 bad;
                bool ok = joint_prestep(mk(la.x, la.y), mk(la.z, la.w), pr.x, pr.y, mk(p1.x, p1.y), c1.x, c1.y, mk(p2.x, p2.y), c2.x, c2.y,
                                        i1.x, i1.y, i2.x, i2.y, a.inv_dt, poscorr, warm, pacc, v1, v2, jc, &bad, a.fix_inverse != 0);
                s.jr[j] = make_float4(jc.r1.x, jc.r1.y, jc.r2.x, jc.r2.y);
                if (!ok) {  // quirk Q-15
                    if (atomicCAS(s_err, 0, SYLT_ERR_NO_INVERSE) == 0) {
                        a.w_errjoint[wld] = j;
                        a.w_badk[wld] = make_float4(bad.c1.x, bad.c2.x, bad.c1.y, bad.c2.y);
                    }
                    continue;
                }
                s.jm[j] = make_float4(jc.m.c1.x, jc.m.c2.x, jc.m.c1.y, jc.m.c2.y);
                s.jbias[j] = make_float2(jc.bias.x, jc.bias.y);
                s.jp[j] = make_float2(pacc.x, pacc.y);
                if (warm) {
                    if (i1.x != 0.0f || i1.y != 0.0f) st_vel(s.vel, b.x, v1);
                    if (i2.x != 0.0f || i2.y != 0.0f) st_vel(s.vel, b.y, v2);
                }
            }
            __syncthreads();
        }
        if (*s_err) break;
        PHASE(6);
        // Joint::apply_impulse by joint colour (all threads; barrier 0)
        auto joint_hops = [&]() {
            for (int c = 0; c < njc; c++) {
                for (int k = jcs[c] + tid; k < jcs[c + 1]; k += T) {
                    int j = s.jorder[k];
                    int2 b = s.jb[j];
                    float2 i1 = s.inv[b.x], i2 = s.inv[b.y];
                    Vel v1 = ld_vel(s.vel, b.x), v2 = ld_vel(s.vel, b.y);
                    float4 r = s.jr[j], mm = s.jm[j];
                    float2 bs = s.jbias[j], pa = s.jp[j];
                    JointConst jc;
                    jc.r1 = mk(r.x, r.y); jc.r2 = mk(r.z, r.w); jc.bias = mk(bs.x, bs.y);
                    jc.m = mk(mk(mm.x, mm.z), mk(mm.y, mm.w));
                    jc.softness = s.jparam[j].x;
                    V2 pacc = mk(pa.x, pa.y);
                    joint_apply(jc, i1.x, i1.y, i2.x, i2.y, pacc, v1, v2);
                    s.jp[j] = make_float2(pacc.x, pacc.y);
                    if (i1.x != 0.0f || i1.y != 0.0f) st_vel(s.vel, b.x, v1);
                    if (i2.x != 0.0f || i2.y != 0.0f) st_vel(s.vel, b.y, v2);
                }
                __syncthreads();
            }
        };
        // ---- 7. iterations (world.rs:132-140): one thread per arbiter of the colour, records read by colour-major position
        {
            const int sweeps = njc > 0 ? 1 : a.iterations;  // joints are solved after the contacts inside every iteration
            for (int it = 0; it < a.iterations; it += sweeps) {
                if (acc) contact_sweeps<true>(hot, ncolors, sweeps); else contact_sweeps<false>(hot, ncolors, sweeps);
                if (njc > 0) joint_hops();
            }
        }
        PHASE(7);
        // ---- 8. integrate positions (world.rs:143-150, all bodies)
        for (int i = tid; i < B; i += T) {
            float4 v = s.vel[i];
            float2 p = s.pos[i];
            V2 np = mk(p.x, p.y) + mk(v.x, v.y) * a.dt;
            s.pos[i] = make_float2(np.x, np.y);
            s.rot[i] += v.z * a.dt;
        }
        // ---- 9. this step's slots become the cache of the next one
        for (int k = tid; k < A; k += T)  // (friction, accumulated impulses of contact 0) by slot, for the next step's Arbiter::update
            a.c_cache[par ^ 1][gslot + s.order[k]] = make_float4(s.frk[k], s.cB[0][k].w, s.ptk[k].x, 0.0f);
        par ^= 1;
        __syncthreads();  // frk aliases okey
        for (int p = tid; p < P; p += T) {
            s.okey[p] = s.key[p];
            s.ocolor[p] = s.cnt[p] > 0 ? s.color[p] : NO_COLOR;
        }
        if (tid == 0) { *s_Pold = P; s.misc[138] = 0; s.misc[139] = 0; }
        __syncthreads();
        PHASE(8);
    }

    // ---- store the world
    __syncthreads();
    const int err = *s_err;
    for (int i = tid; i < B; i += T) {
        const float2 p = s.pos[i];
        const float4 v = s.vel[i];
        const float rt = s.rot[i];
        a.pose[b0 + i] = make_float4(p.x, p.y, rt, 0.0f);
        a.vel[b0 + i] = v;
        if (a.rec_out) {
            float* r = a.rec_out + (size_t)(b0 + i) * a.rec_floats;
            r[0] = p.x; r[1] = p.y; r[2] = rt; r[3] = v.x; r[4] = v.y; r[5] = v.z;
            for (int k = 6; k < a.rec_floats; k++) r[k] = 0.0f;  // SyltBodyState: force / torque (a batch has no external forces)
        }
    }
    for (int j = tid; j < J; j += T) {
        a.jp[j0 + j] = s.jp[j]; a.jr[j0 + j] = s.jr[j]; a.jm[j0 + j] = s.jm[j]; a.jbias[j0 + j] = s.jbias[j];
    }
    const int Pn = *s_Pold;
    int ncon = 0, narb = 0;
    for (int k = tid; k < Pn; k += T) {
        size_t g = (size_t)wld * cap + k;
        a.c_key[g] = s.okey[k]; a.c_color[g] = s.ocolor[k];
        if (!err && k < P && s.cnt[k] > 0) { ncon += s.cnt[k]; narb++; }
    }
    // block reductions of (ncon, narb, maxpen) through shared memory atomics
    if (tid == 0) { s.misc[0] = 0; s.misc[1] = 0; s.misc[2] = 0; }
    __syncthreads();
    atomicAdd(&s.misc[0], ncon);
    atomicAdd(&s.misc[1], narb);
    atomicMax(&s.misc[2], __float_as_int(fmaxf(maxpen, 0.0f)));
    __syncthreads();
    if (tid == 0) {
        a.c_count[wld] = Pn;
        a.c_par[wld] = par;
        a.w_err[wld] = err;
        a.w_ncon[wld] = s.misc[0];
        a.w_narb[wld] = s.misc[1];
        a.w_maxpen[wld] = __int_as_float(s.misc[2]);
        a.w_ncolors[wld] = ncolors;
    }
    // ---- optional full export of the last step's arbiters (parity tests / renderer read-back)
    if (exporting) {
        const bool valid = !err && a.n_steps > 0;
        for (int k = tid; k < (valid ? cstart[BNC] : 0); k += T) {
            const int p = s.order[k];
            SlotExport* e = &a.exp_slots[(size_t)xw * cap + p];
            e->friction = a.c_cache[par][gslot + p].x; e->color = s.color[p];
            const float2 pt = s.ptk[k];
            for (int q = 0; q < 2; q++) {
                const float4 ca = s.cA[q][k], cb = s.cB[q][k];
                float* o = e->sol[q];
                o[0] = ca.x; o[1] = ca.y; o[2] = ca.z; o[3] = ca.w; o[4] = cb.x; o[5] = cb.y; o[6] = cb.z; o[7] = cb.w; o[8] = q ? pt.y : pt.x;
            }
        }
        if (tid == 0) a.exp_count[xw] = valid ? P : 0;
    }
}

// AoS (SyltBodyState, the ABI record) <-> SoA (pose, vel) on the device, so host buffers move with ONE DMA each way
__global__ void k_unpack_states(int n, const SyltBodyState* in, float4* pose, float4* vel) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    SyltBodyState s = in[i];
    pose[i] = make_float4(s.x, s.y, s.rotation, 0.0f);
    vel[i] = make_float4(s.vx, s.vy, s.angular_velocity, 0.0f);
}
__global__ void k_pack_states(int n, const float4* pose, const float4* vel, SyltBodyState* out) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    float4 p = pose[i], v = vel[i];
    SyltBodyState s = {p.x, p.y, p.z, v.x, v.y, v.z, 0.0f, 0.0f, 0.0f};
    out[i] = s;
}

struct StatsOut {
    unsigned long long contacts, errors;
    double ke;
    float maxpen;
};

__global__ void k_batch_stats(int n_worlds, const int* body_off, const float4* vel, const float4* mp, const float* mass_moi /*2 per body*/,
                              const int* w_err, const int* w_ncon, const float* w_maxpen, StatsOut* out) {
    __shared__ double sh_ke[256];
    __shared__ unsigned long long sh_c[256], sh_e[256];
    __shared__ float sh_p[256];
    double ke = 0.0;
    unsigned long long c = 0, e = 0;
    float mp_ = 0.0f;
    for (int w = blockIdx.x * blockDim.x + threadIdx.x; w < n_worlds; w += gridDim.x * blockDim.x) {
        for (int i = body_off[w]; i < body_off[w + 1]; i++) {
            float4 m = mp[i];
            if (m.x != 0.0f) {
                float4 v = vel[i];
                ke += 0.5 * (double)mass_moi[2 * i] * ((double)v.x * v.x + (double)v.y * v.y) + 0.5 * (double)mass_moi[2 * i + 1] * (double)v.z * v.z;
            }
        }
        c += (unsigned long long)w_ncon[w];
        e += w_err[w] ? 1ull : 0ull;
        mp_ = fmaxf(mp_, w_maxpen[w]);
    }
    sh_ke[threadIdx.x] = ke; sh_c[threadIdx.x] = c; sh_e[threadIdx.x] = e; sh_p[threadIdx.x] = mp_;
    __syncthreads();
    for (int o = blockDim.x / 2; o > 0; o >>= 1) {
        if ((int)threadIdx.x < o) {
            sh_ke[threadIdx.x] += sh_ke[threadIdx.x + o]; sh_c[threadIdx.x] += sh_c[threadIdx.x + o]; sh_e[threadIdx.x] += sh_e[threadIdx.x + o];
            sh_p[threadIdx.x] = fmaxf(sh_p[threadIdx.x], sh_p[threadIdx.x + o]);
        }
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        atomicAdd(&out->contacts, sh_c[0]);
        atomicAdd(&out->errors, sh_e[0]);
        atomicAdd(&out->ke, sh_ke[0]);
        atomicMax((int*)&out->maxpen, __float_as_int(sh_p[0]));
    }
}

}  // namespace

// ====================================================================== host side
struct SyltBatch {
    int device = 0;
    cudaStream_t stream = nullptr;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    int n_worlds = 0, Bmax = 0, Jmax = 0, cap = 0, threads = 256, nbk = 64;
    float cell = 1.0f;
    int64_t n_bodies = 0, n_joints = 0;
    float gx = 0, gy = 0;
    uint32_t iterations = 10;
    int accumulate = 1, warm = 0, poscorr = 1;  // World::new defaults (world.rs:38-42)
    int fix_inverse = 0;                        // opt-in corrected Mat2x2 inverse, default off
    size_t smem = 0;
    int64_t launches = 0;
    uint64_t body_steps = 0;
    int export_count = 0;             // worlds with full arbiter read-back (<= MAX_EXPORT)
    std::vector<int> h_exp_map;       // per world: export slot or -1
    DBuf<int> exp_map;
    int last_steps = 0, wave = 0;     // wave: CTAs of k_batch_step the device holds at once
    bool inflight = false;
    std::vector<int> h_body_off, h_joint_off, h_jorder;
    std::vector<uint64_t> h_ids;
    struct HJ { int b1, b2; float la1x, la1y, la2x, la2y, softness, bias_factor; };
    std::vector<HJ> h_joints;
    DBuf<int> body_off, joint_off, jorder, jcolor_start, n_jcolors, c_count, w_err, w_ncon, w_narb, w_ncolors, w_errjoint, exp_count;
    DBuf<float4> pose, vel, mp, jla, jr, jm, w_badk;
    DBuf<float2> wh, jparam, jp, jbias;
    DBuf<float> mass_moi, w_maxpen;
    DBuf<float4> c_cache0, c_cache1, c_stg0, c_stg1;
    DBuf<int> c_par;
    DBuf<int2> jbodies;
    DBuf<unsigned> c_key;
    DBuf<unsigned char> c_color, bflags;
    DBuf<SlotExport> exp_slots;
    DBuf<StatsOut> stats;
    DBuf<SyltBodyState> stage;
    static constexpr int NS = 4, MAX_CHUNKS = 64;
    cudaEvent_t pipe_ev[2 * MAX_CHUNKS] = {};
    cudaStream_t xs[NS] = {nullptr, nullptr, nullptr, nullptr};  // copy/compute pipeline of sylt_batch_step_host
};

namespace {

void batch_free(SyltBatch* b) {
    b->body_off.release(); b->joint_off.release(); b->jorder.release(); b->jcolor_start.release(); b->n_jcolors.release();
    b->c_count.release(); b->w_err.release(); b->w_ncon.release(); b->w_narb.release(); b->w_ncolors.release(); b->w_errjoint.release();
    b->exp_count.release(); b->pose.release(); b->vel.release(); b->mp.release(); b->jla.release(); b->jr.release(); b->jm.release();
    b->w_badk.release(); b->wh.release(); b->jparam.release(); b->jp.release(); b->jbias.release(); b->mass_moi.release();
    b->c_cache0.release(); b->c_cache1.release(); b->c_stg0.release(); b->c_stg1.release(); b->c_par.release(); b->w_maxpen.release(); b->jbodies.release(); b->c_key.release();
    b->c_color.release(); b->bflags.release(); b->exp_slots.release(); b->exp_map.release(); b->stats.release(); b->stage.release();
}

template <typename T>
int upload(DBuf<T>& d, const std::vector<T>& h, size_t min_n = 1) {
    SYLT_CUDA(d.ensure(std::max(h.size(), min_n)));
    if (!h.empty()) SYLT_CUDA(cudaMemcpy(d.p, h.data(), h.size() * sizeof(T), cudaMemcpyHostToDevice));
    return SYLT_OK;
}
template <typename T>
int zeroed(DBuf<T>& d, size_t n) {
    SYLT_CUDA(d.ensure(std::max<size_t>(n, 1)));
    SYLT_CUDA(cudaMemset(d.p, 0, std::max<size_t>(n, 1) * sizeof(T)));
    return SYLT_OK;
}

// Everything that depends on the per-world candidate-slot capacity: shared-memory size, CTA width, the per-world slot caches
// (zeroed: last step's contact history is dropped) and the export buffers.
int batch_set_cap(SyltBatch* b, int cap) {
    cap = std::min(65535, std::max(32, (cap + 1) & ~1));
    const size_t smem = carve(nullptr, nullptr, b->Bmax, b->Jmax, cap, b->nbk);
    int max_smem = 0;
    SYLT_CUDA(cudaDeviceGetAttribute(&max_smem, cudaDevAttrMaxSharedMemoryPerBlockOptin, b->device));
    if (smem > (size_t)max_smem) {
        set_last_error("world too large for the one-CTA-per-world path (shared memory)");
        return SYLT_ERR_UNSUPPORTED;
    }
    SYLT_CUDA(cudaFuncSetAttribute(k_batch_step, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    b->cap = cap;
    b->smem = smem;
    b->wave = 0;
    int t = std::max(b->Bmax, (cap + 1) / 2);
    t = ((t + 31) / 32) * 32;
    b->threads = std::min(256, std::max(32, t));
    const size_t W = (size_t)b->n_worlds, WC = W * (size_t)cap;
    int e = 0;
    if ((e = zeroed(b->c_count, W)) || (e = zeroed(b->c_key, WC)) || (e = zeroed(b->c_cache0, WC)) || (e = zeroed(b->c_cache1, WC)) ||
        (e = zeroed(b->c_stg0, WC)) || (e = zeroed(b->c_stg1, WC)) || (e = zeroed(b->c_par, W)) || (e = zeroed(b->c_color, WC)) ||
        (e = zeroed(b->exp_slots, (size_t)b->export_count * (size_t)cap)) || (e = zeroed(b->exp_count, (size_t)std::max(1, b->export_count))))
        return e;
    return SYLT_OK;
}

int batch_enqueue(SyltBatch* b, float dt, int n, int w0 = 0, int w1 = -1, cudaStream_t st = nullptr, const float* rec_in = nullptr,
                  float* rec_out = nullptr, int rec_floats = 0) {
    if (w1 < 0) w1 = b->n_worlds;
    if (!st) st = b->stream;
    BatchArgs a;
    memset(&a, 0, sizeof a);
    a.n_worlds = w1; a.world_begin = w0; a.n_steps = n; a.iterations = (int)b->iterations;
    a.accumulate = b->accumulate; a.warm = b->warm; a.poscorr = b->poscorr; a.fix_inverse = b->fix_inverse;
    a.dt = dt; a.inv_dt = dt > 0.0f ? 1.0f / dt : 0.0f;  // world.rs:108
    {
        Smem hs;
        static unsigned char anchor[16];  // any non-null base: only the differences are used
        carve(&hs, anchor, b->Bmax, b->Jmax, b->cap, b->nbk);
        unsigned char* ptrs[sizeof(Smem) / sizeof(void*)];
        memcpy(ptrs, &hs, sizeof(Smem));
        for (size_t i = 0; i < sizeof(Smem) / sizeof(void*); i++) a.smem_off[i] = (unsigned long long)(ptrs[i] - anchor);
    }
    a.rec_in = rec_in; a.rec_out = rec_out; a.rec_floats = rec_floats;
    a.gx = b->gx; a.gy = b->gy; a.Bmax = b->Bmax; a.Jmax = b->Jmax; a.cap = b->cap; a.nbk = b->nbk; a.cell = b->cell; a.bflags = b->bflags.p;
    a.exp_map = b->exp_map.p;
    a.body_off = b->body_off.p; a.joint_off = b->joint_off.p; a.pose = b->pose.p; a.vel = b->vel.p; a.mp = b->mp.p; a.wh = b->wh.p;
    a.jbodies = b->jbodies.p; a.jla = b->jla.p; a.jparam = b->jparam.p; a.jp = b->jp.p; a.jr = b->jr.p; a.jm = b->jm.p; a.jbias = b->jbias.p;
    a.jorder = b->jorder.p; a.jcolor_start = b->jcolor_start.p; a.n_jcolors = b->n_jcolors.p;
    a.c_count = b->c_count.p; a.c_key = b->c_key.p; a.c_cache[0] = b->c_cache0.p; a.c_cache[1] = b->c_cache1.p; a.c_par = b->c_par.p; a.c_stg0 = b->c_stg0.p; a.c_stg1 = b->c_stg1.p; a.c_color = b->c_color.p;
    a.w_err = b->w_err.p; a.w_ncon = b->w_ncon.p; a.w_narb = b->w_narb.p; a.w_ncolors = b->w_ncolors.p; a.w_errjoint = b->w_errjoint.p;
    a.w_maxpen = b->w_maxpen.p; a.w_badk = b->w_badk.p; a.exp_slots = b->exp_slots.p; a.exp_count = b->exp_count.p;
    k_batch_step<<<w1 - w0, b->threads, b->smem, st>>>(a);
    SYLT_CUDA(cudaGetLastError());
    b->launches++;
    b->body_steps += (uint64_t)(b->h_body_off[(size_t)w1] - b->h_body_off[(size_t)w0]) * (uint64_t)n;
    b->last_steps = n;
    b->inflight = true;
    return SYLT_OK;
}

}  // namespace

extern "C" {

int sylt_batch_create(int32_t n_worlds, const int32_t* body_offsets, const SyltBodyDesc* bodies, const int32_t* joint_offsets,
                      const SyltJointDesc* joints, float gx, float gy, uint32_t iterations, int device, SyltBatch** out) {
    if (!out || n_worlds < 1 || !body_offsets || !bodies) return SYLT_ERR_INVALID;
    *out = nullptr;
    int ndev = 0;
    SYLT_CUDA(cudaGetDeviceCount(&ndev));
    if (device < 0 || device >= ndev) { set_last_error("no such CUDA device"); return SYLT_ERR_CUDA; }
    SYLT_CUDA(cudaSetDevice(device));
    const int64_t NB = body_offsets[n_worlds];
    const int64_t NJ = joint_offsets ? joint_offsets[n_worlds] : 0;
    if (NJ > 0 && !joints) return SYLT_ERR_INVALID;
    SyltBatch* b = new SyltBatch();
    b->device = device; b->n_worlds = n_worlds; b->gx = gx; b->gy = gy; b->iterations = iterations;
    b->n_bodies = NB; b->n_joints = NJ;
    b->h_body_off.assign(body_offsets, body_offsets + n_worlds + 1);
    if (joint_offsets) b->h_joint_off.assign(joint_offsets, joint_offsets + n_worlds + 1);
    else b->h_joint_off.assign((size_t)n_worlds + 1, 0);
    std::vector<float4> pose((size_t)NB), vel((size_t)NB), mp((size_t)NB);
    std::vector<float2> wh((size_t)NB);
    std::vector<float> mm((size_t)NB * 2), radius((size_t)NB);
    b->h_ids.resize((size_t)NB);
    int Bmax = 0, Jmax = 0;
    std::vector<V2> scratch;
    for (int w = 0; w < n_worlds; w++) {
        int o0 = body_offsets[w], o1 = body_offsets[w + 1];
        if (o1 < o0) { delete b; return SYLT_ERR_INVALID; }
        Bmax = std::max(Bmax, o1 - o0);
        Jmax = std::max(Jmax, b->h_joint_off[(size_t)w + 1] - b->h_joint_off[(size_t)w]);
        std::vector<HostBody> hbs;
        for (int i = o0; i < o1; i++) {
            const SyltBodyDesc& d = bodies[i];
            if (d.shape != SYLT_SHAPE_BOX) { delete b; return SYLT_ERR_UNSUPPORTED; }
            HostBody hb;
            scratch.clear();
            int e = make_host_body(d, nullptr, 0, &hb, &scratch);
            if (e) { delete b; return e; }
            hbs.push_back(hb);
            pose[(size_t)i] = make_float4(d.x, d.y, d.rotation, 0.0f);
            vel[(size_t)i] = make_float4(d.vx, d.vy, d.angular_velocity, 0.0f);
            mp[(size_t)i] = make_float4(hb.inv_mass, hb.inv_moi, hb.friction, 0.0f);
            wh[(size_t)i] = make_float2(hb.width_x, hb.width_y);
            mm[(size_t)i * 2] = hb.mass; mm[(size_t)i * 2 + 1] = hb.moi;
            radius[(size_t)i] = hb.radius;
            b->h_ids[(size_t)i] = d.id;
        }
        if (body_order_violated(hbs)) { delete b; return SYLT_ERR_BODY_ORDER; }  // quirk Q-6
    }
    if (Bmax > 4096) { delete b; return SYLT_ERR_UNSUPPORTED; }
    // grid classification shared by all worlds of the batch: cell just above the largest "small" bounding diameter
    std::vector<unsigned char> fl((size_t)NB, 0);
    {
        std::vector<float> d;
        for (int64_t i = 0; i < NB; i++)
            if (mp[(size_t)i].x != 0.0f) d.push_back(2.0f * radius[(size_t)i]);
        float cell = 1.0f;
        if (!d.empty()) {
            std::vector<float> t = d;
            std::nth_element(t.begin(), t.begin() + (long)(t.size() / 2), t.end());
            float med = t[t.size() / 2], mx = 0.0f;
            for (float x : d)
                if (x <= 4.0f * med && x > mx) mx = x;
            if (mx > 0.0f) cell = mx;
        }
        b->cell = 1.1f * cell + 0.1f;
        for (int64_t i = 0; i < NB; i++) {
            unsigned char f = mp[(size_t)i].x == 0.0f ? 1 : 0;
            if (2.2f * radius[(size_t)i] + 0.1f > b->cell || !(radius[(size_t)i] == radius[(size_t)i])) f |= 2;
            fl[(size_t)i] = f;
        }
        // >= 4 buckets per body: with CELL_ROW_STRIDE = 37 two rows of cells only share buckets when they are >= nbk / 37 rows apart
        int nbk = 64;
        while (nbk < 4 * Bmax && nbk < 4096) nbk <<= 1;
        b->nbk = nbk;
    }
    // joints: Joint::new (joint.rs:26-55) per world, then a greedy colouring of each world's joint graph
    std::vector<int2> jb((size_t)NJ);
    std::vector<float4> jla((size_t)NJ);
    std::vector<float2> jparam((size_t)NJ);
    std::vector<int> jorder((size_t)NJ), jcs((size_t)n_worlds * (MAX_JC + 1), 0), njc((size_t)n_worlds, 0);
    b->h_joints.resize((size_t)NJ);
    for (int w = 0; w < n_worlds; w++) {
        int o0 = body_offsets[w], nb = body_offsets[w + 1] - o0;
        int q0 = b->h_joint_off[(size_t)w], nj = b->h_joint_off[(size_t)w + 1] - q0;
        std::vector<unsigned> used((size_t)nb, 0u);
        std::vector<int> color((size_t)nj, 0);
        int nc = 0;
        for (int j = 0; j < nj; j++) {
            const SyltJointDesc& d = joints[q0 + j];
            int b1 = -1, b2 = -1;
            for (int i = 0; i < nb && (b1 < 0 || b2 < 0); i++) {
                if (b1 < 0 && bodies[o0 + i].id == d.id1) b1 = i;
                if (b2 < 0 && bodies[o0 + i].id == d.id2) b2 = i;
            }
            if (b1 < 0 || b2 < 0) { delete b; return SYLT_ERR_BODY_NOT_FOUND; }
            const SyltBodyDesc &s1 = bodies[o0 + b1], &s2 = bodies[o0 + b2];
            M2 rt1 = transpose(rot_angle(s1.rotation)), rt2 = transpose(rot_angle(s2.rotation));
            V2 anchor = mk(d.anchor_x, d.anchor_y);
            V2 la1 = rt1 * (anchor - mk(s1.x, s1.y)), la2 = rt2 * (anchor - mk(s2.x, s2.y));
            jb[(size_t)(q0 + j)] = make_int2(b1, b2);
            jla[(size_t)(q0 + j)] = make_float4(la1.x, la1.y, la2.x, la2.y);
            jparam[(size_t)(q0 + j)] = make_float2(d.softness, d.bias_factor);
            SyltBatch::HJ hj = {b1, b2, la1.x, la1.y, la2.x, la2.y, d.softness, d.bias_factor};
            b->h_joints[(size_t)(q0 + j)] = hj;
            bool d1 = mp[(size_t)(o0 + b1)].x != 0.0f || mp[(size_t)(o0 + b1)].y != 0.0f;
            bool d2 = mp[(size_t)(o0 + b2)].x != 0.0f || mp[(size_t)(o0 + b2)].y != 0.0f;
            unsigned m = (d1 ? used[(size_t)b1] : 0u) | (d2 ? used[(size_t)b2] : 0u);
            int c = 0;
            while (c < MAX_JC && ((m >> c) & 1u)) c++;
            if (c >= MAX_JC) { delete b; return SYLT_ERR_UNSUPPORTED; }
            color[(size_t)j] = c;
            if (d1) used[(size_t)b1] |= 1u << c;
            if (d2) used[(size_t)b2] |= 1u << c;
            nc = std::max(nc, c + 1);
        }
        njc[(size_t)w] = nc;
        int* st = &jcs[(size_t)w * (MAX_JC + 1)];
        for (int j = 0; j < nj; j++) st[color[(size_t)j] + 1]++;
        for (int c = 0; c < MAX_JC; c++) st[c + 1] += st[c];
        std::vector<int> fill(st, st + MAX_JC);
        for (int j = 0; j < nj; j++) jorder[(size_t)(q0 + fill[(size_t)color[(size_t)j]]++)] = j;
    }
    b->h_jorder = jorder;
    b->Bmax = Bmax; b->Jmax = Jmax;
    b->export_count = std::min(n_worlds, MAX_EXPORT);  // default: the first worlds (sylt_batch_set_export_worlds picks others)
    b->h_exp_map.assign((size_t)n_worlds, -1);
    for (int k = 0; k < b->export_count; k++) b->h_exp_map[(size_t)k] = k;
    { int e0 = upload(b->exp_map, b->h_exp_map); if (e0) { sylt_batch_destroy(b); return e0; } }
    // default slot capacity: two candidate pairs per body (a pyramid or a chain has fewer); sylt_batch_set_capacity raises it
    { int e0 = batch_set_cap(b, 2 * Bmax + 32); if (e0) { sylt_batch_destroy(b); return e0; } }
    SYLT_CUDA(cudaStreamCreateWithFlags(&b->stream, cudaStreamNonBlocking));
    SYLT_CUDA(cudaEventCreate(&b->ev0));
    SYLT_CUDA(cudaEventCreate(&b->ev1));
    for (int k = 0; k < SyltBatch::NS; k++) SYLT_CUDA(cudaStreamCreateWithFlags(&b->xs[k], cudaStreamNonBlocking));
    int e = 0;
    if ((e = upload(b->body_off, b->h_body_off)) || (e = upload(b->joint_off, b->h_joint_off)) || (e = upload(b->pose, pose)) ||
        (e = upload(b->vel, vel)) || (e = upload(b->mp, mp)) || (e = upload(b->bflags, fl)) || (e = upload(b->wh, wh)) || (e = upload(b->mass_moi, mm)) ||
        (e = upload(b->jbodies, jb)) || (e = upload(b->jla, jla)) || (e = upload(b->jparam, jparam)) || (e = upload(b->jorder, jorder)) ||
        (e = upload(b->jcolor_start, jcs)) || (e = upload(b->n_jcolors, njc))) { sylt_batch_destroy(b); return e; }
    const size_t W = (size_t)n_worlds;
    if ((e = zeroed(b->jp, (size_t)NJ)) || (e = zeroed(b->jr, (size_t)NJ)) || (e = zeroed(b->jm, (size_t)NJ)) || (e = zeroed(b->jbias, (size_t)NJ)) ||
        (e = zeroed(b->w_err, W)) || (e = zeroed(b->w_ncon, W)) ||
        (e = zeroed(b->w_narb, W)) || (e = zeroed(b->w_ncolors, W)) || (e = zeroed(b->w_errjoint, W)) || (e = zeroed(b->w_maxpen, W)) ||
        (e = zeroed(b->w_badk, W)) || (e = zeroed(b->stats, 1))) { sylt_batch_destroy(b); return e; }
    *out = b;
    return SYLT_OK;
}

int sylt_batch_destroy(SyltBatch* b) {
    if (!b) return SYLT_OK;
    cudaSetDevice(b->device);
    if (b->stream) cudaStreamSynchronize(b->stream);
    batch_free(b);
    if (b->ev0) cudaEventDestroy(b->ev0);
    if (b->ev1) cudaEventDestroy(b->ev1);
    if (b->stream) cudaStreamDestroy(b->stream);
    for (int k = 0; k < SyltBatch::NS; k++)
        if (b->xs[k]) cudaStreamDestroy(b->xs[k]);
    for (cudaEvent_t ev : b->pipe_ev)
        if (ev) cudaEventDestroy(ev);
    delete b;
    return SYLT_OK;
}

int sylt_batch_set_context(SyltBatch* b, int acc, int warm, int pc) {
    if (!b) return SYLT_ERR_INVALID;
    b->accumulate = acc != 0; b->warm = warm != 0; b->poscorr = pc != 0;
    return SYLT_OK;
}

int sylt_batch_set_capacity(SyltBatch* b, int32_t slots_per_world) {
    if (!b || slots_per_world < 0) return SYLT_ERR_INVALID;
    SYLT_CUDA(cudaSetDevice(b->device));
    SYLT_CUDA(cudaStreamSynchronize(b->stream));
    return batch_set_cap(b, slots_per_world == 0 ? 2 * b->Bmax + 32 : slots_per_world);
}

int sylt_batch_set_export_worlds(SyltBatch* b, const int32_t* worlds, int32_t n) {
    if (!b || n < 0 || n > MAX_EXPORT || (n > 0 && !worlds)) return SYLT_ERR_INVALID;
    std::vector<int> m((size_t)b->n_worlds, -1);
    for (int k = 0; k < n; k++) {
        if (worlds[k] < 0 || worlds[k] >= b->n_worlds || m[(size_t)worlds[k]] >= 0) return SYLT_ERR_INVALID;
        m[(size_t)worlds[k]] = k;
    }
    SYLT_CUDA(cudaSetDevice(b->device));
    SYLT_CUDA(cudaStreamSynchronize(b->stream));
    b->h_exp_map = m;
    b->export_count = n;
    int e = upload(b->exp_map, b->h_exp_map);
    if (e) return e;
    if ((e = zeroed(b->exp_slots, (size_t)n * (size_t)b->cap)) || (e = zeroed(b->exp_count, (size_t)std::max(1, n)))) return e;
    return SYLT_OK;
}

int sylt_batch_set_fixes(SyltBatch* b, int fix_features, int fix_inverse) {
    if (!b) return SYLT_ERR_INVALID;
    if (fix_features) return SYLT_ERR_UNSUPPORTED;  // the per-world slot cache keeps contact 0's impulses only
    b->fix_inverse = fix_inverse != 0;
    return SYLT_OK;
}

int sylt_batch_sync(SyltBatch* b) {
    if (!b) return SYLT_ERR_INVALID;
    SYLT_CUDA(cudaSetDevice(b->device));
    SYLT_CUDA(cudaStreamSynchronize(b->stream));
    b->inflight = false;
    return SYLT_OK;
}
int sylt_batch_step_n_async(SyltBatch* b, float dt, int32_t n) {
    if (!b || n < 0) return SYLT_ERR_INVALID;
    if (n == 0) return SYLT_OK;
    SYLT_CUDA(cudaSetDevice(b->device));
    return batch_enqueue(b, dt, n);
}
int sylt_batch_step_n(SyltBatch* b, float dt, int32_t n) {
    int e = sylt_batch_step_n_async(b, dt, n);
    if (e) return e;
    return sylt_batch_sync(b);
}
// One step_n with HOST buffers: upload all body states, step, download all body states — chunked over worlds and
// pipelined on several streams so the two DMA directions overlap the kernels (pinned host memory recommended).
// floats_per_body: 9 (SyltBodyState) or 6 (x, y, rotation, vx, vy, angular_velocity).
static int step_host_impl(SyltBatch* b, float dt, int32_t n, const float* in, float* out, int fpb) {
    if (!b || n < 0) return SYLT_ERR_INVALID;
    SYLT_CUDA(cudaSetDevice(b->device));
    SYLT_CUDA(cudaStreamSynchronize(b->stream));
    SYLT_CUDA(b->stage.ensure((size_t)b->n_bodies));
    float* stage = (float*)b->stage.p;  // 9 floats per body of room; the compact layout uses the first 6 * n_bodies floats
    const int W = b->n_worlds;
    // chunks: each at least two resident waves of CTAs (a shorter launch leaves SMs idle at its tail), at most 16
    if (b->wave <= 0) {
        int per_sm = 1, sms = 1;
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_batch_step, b->threads, b->smem);
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, b->device);
        b->wave = std::max(1, per_sm) * std::max(1, sms);
    }
    static const int env_chunks = getenv("SYLT_BATCH_CHUNKS") ? atoi(getenv("SYLT_BATCH_CHUNKS")) : 0;
    static const int env_mode = getenv("SYLT_BATCH_PIPE") ? atoi(getenv("SYLT_BATCH_PIPE")) : 1;  // 0: every chunk on one of four streams (upload, step, download in stream order)
    const int chunks = std::max(1, std::min(env_chunks > 0 ? std::min(env_chunks, SyltBatch::MAX_CHUNKS) : 16, W / (2 * b->wave)));
    const int per = (W + chunks - 1) / chunks;
    for (int c = 0; c < chunks; c++) {
        const int w0 = std::min(W, c * per), w1 = std::min(W, w0 + per);
        if (w1 <= w0) continue;
        const int o0 = b->h_body_off[(size_t)w0], cnt = b->h_body_off[(size_t)w1] - o0;
        float* sg = stage + (size_t)o0 * fpb;
        if (env_mode == 1) {  // three-stage pipeline: one upload stream, two compute streams, one download stream, events between them
            cudaStream_t up = b->xs[0], down = b->xs[1], cs = b->xs[2 + (c & 1)];
            if (!b->pipe_ev[2 * c]) { SYLT_CUDA(cudaEventCreateWithFlags(&b->pipe_ev[2 * c], cudaEventDisableTiming)); SYLT_CUDA(cudaEventCreateWithFlags(&b->pipe_ev[2 * c + 1], cudaEventDisableTiming)); }
            if (in && cnt) {
                SYLT_CUDA(cudaMemcpyAsync(sg, in + (size_t)o0 * fpb, (size_t)cnt * fpb * sizeof(float), cudaMemcpyHostToDevice, up));
                SYLT_CUDA(cudaEventRecord(b->pipe_ev[2 * c], up));
                SYLT_CUDA(cudaStreamWaitEvent(cs, b->pipe_ev[2 * c], 0));
            }
            int e = batch_enqueue(b, dt, n, w0, w1, cs, (in && cnt) ? stage : nullptr, (out && cnt) ? stage : nullptr, fpb);
            if (e) return e;
            if (out && cnt) {
                SYLT_CUDA(cudaEventRecord(b->pipe_ev[2 * c + 1], cs));
                SYLT_CUDA(cudaStreamWaitEvent(down, b->pipe_ev[2 * c + 1], 0));
                SYLT_CUDA(cudaMemcpyAsync(out + (size_t)o0 * fpb, sg, (size_t)cnt * fpb * sizeof(float), cudaMemcpyDeviceToHost, down));
            }
            continue;
        }
        cudaStream_t st = b->xs[c % SyltBatch::NS];
        if (in && cnt) SYLT_CUDA(cudaMemcpyAsync(sg, in + (size_t)o0 * fpb, (size_t)cnt * fpb * sizeof(float), cudaMemcpyHostToDevice, st));
        // the step kernel itself reads the host-format records and writes them back (n == 0: zero steps, i.e. a pure conversion)
        int e = batch_enqueue(b, dt, n, w0, w1, st, (in && cnt) ? stage : nullptr, (out && cnt) ? stage : nullptr, fpb);
        if (e) return e;
        if (out && cnt) SYLT_CUDA(cudaMemcpyAsync(out + (size_t)o0 * fpb, sg, (size_t)cnt * fpb * sizeof(float), cudaMemcpyDeviceToHost, st));
    }
    for (int k = 0; k < SyltBatch::NS; k++) SYLT_CUDA(cudaStreamSynchronize(b->xs[k]));
    b->inflight = false;
    return SYLT_OK;
}
int sylt_batch_step_host(SyltBatch* b, float dt, int32_t n, const SyltBodyState* in, SyltBodyState* out) {
    return step_host_impl(b, dt, n, (const float*)in, (float*)out, 9);
}
int sylt_batch_step_host6(SyltBatch* b, float dt, int32_t n, const float* in6, float* out6) { return step_host_impl(b, dt, n, in6, out6, 6); }

int sylt_batch_timer_start(SyltBatch* b) {
    if (!b) return SYLT_ERR_INVALID;
    SYLT_CUDA(cudaSetDevice(b->device));
    SYLT_CUDA(cudaEventRecord(b->ev0, b->stream));
    return SYLT_OK;
}
int sylt_batch_timer_stop(SyltBatch* b, float* ms) {
    if (!b || !ms) return SYLT_ERR_INVALID;
    SYLT_CUDA(cudaSetDevice(b->device));
    SYLT_CUDA(cudaEventRecord(b->ev1, b->stream));
    SYLT_CUDA(cudaEventSynchronize(b->ev1));
    SYLT_CUDA(cudaEventElapsedTime(ms, b->ev0, b->ev1));
    b->inflight = false;
    return SYLT_OK;
}
int64_t sylt_batch_launch_count(SyltBatch* b) { return b ? b->launches : 0; }
int sylt_batch_num_worlds(SyltBatch* b) { return b ? b->n_worlds : 0; }
int64_t sylt_batch_num_bodies(SyltBatch* b) { return b ? b->n_bodies : 0; }

int sylt_batch_get_bodies(SyltBatch* b, int32_t first, int32_t nw, SyltBodyState* out, int64_t cap) {
    if (!b || first < 0 || nw < 0 || first + nw > b->n_worlds) return -SYLT_ERR_INVALID;
    const int o0 = b->h_body_off[(size_t)first], o1 = b->h_body_off[(size_t)(first + nw)];
    const int n = o1 - o0;
    if (n > cap) return -SYLT_ERR_CAPACITY;
    if (n == 0) return 0;
    if (cudaSetDevice(b->device) != cudaSuccess) return -SYLT_ERR_CUDA;
    if (b->stage.ensure((size_t)n) != cudaSuccess) return -SYLT_ERR_CUDA;
    k_pack_states<<<(n + 255) / 256, 256, 0, b->stream>>>(n, b->pose.p + o0, b->vel.p + o0, b->stage.p);
    b->launches++;
    cudaError_t ce = cudaMemcpyAsync(out, b->stage.p, (size_t)n * sizeof(SyltBodyState), cudaMemcpyDeviceToHost, b->stream);
    if (ce == cudaSuccess) ce = cudaStreamSynchronize(b->stream);
    if (ce != cudaSuccess) { set_last_error(cudaGetErrorString(ce)); return -SYLT_ERR_CUDA; }
    return n;
}

int sylt_batch_set_bodies(SyltBatch* b, int32_t first, int32_t nw, const SyltBodyState* in, int64_t count) {
    if (!b || first < 0 || nw < 0 || first + nw > b->n_worlds) return SYLT_ERR_INVALID;
    const int o0 = b->h_body_off[(size_t)first], o1 = b->h_body_off[(size_t)(first + nw)];
    const int n = o1 - o0;
    if (count != n) return SYLT_ERR_INVALID;
    if (n == 0) return SYLT_OK;
    SYLT_CUDA(cudaSetDevice(b->device));
    SYLT_CUDA(b->stage.ensure((size_t)n));
    SYLT_CUDA(cudaMemcpyAsync(b->stage.p, in, (size_t)n * sizeof(SyltBodyState), cudaMemcpyHostToDevice, b->stream));
    k_unpack_states<<<(n + 255) / 256, 256, 0, b->stream>>>(n, b->stage.p, b->pose.p + o0, b->vel.p + o0);
    b->launches++;
    SYLT_CUDA(cudaStreamSynchronize(b->stream));
    return SYLT_OK;
}

int sylt_batch_get_arbiters(SyltBatch* b, int32_t world, SyltArbiterInfo* arbiters, int32_t arbiter_cap, SyltContact* contacts, int32_t contact_cap) {
    if (!b || world < 0 || world >= b->n_worlds) return -SYLT_ERR_INVALID;
    if (b->h_exp_map[(size_t)world] < 0) return -SYLT_ERR_UNSUPPORTED;  // not in the export set (sylt_batch_set_export_worlds)
    if (cudaSetDevice(b->device) != cudaSuccess) return -SYLT_ERR_CUDA;
    const int xw = b->h_exp_map[(size_t)world];
    int P = 0;
    cudaStreamSynchronize(b->stream);
    if (cudaMemcpy(&P, b->exp_count.p + xw, sizeof(int), cudaMemcpyDeviceToHost) != cudaSuccess) return -SYLT_ERR_CUDA;
    std::vector<SlotExport> sl((size_t)std::max(P, 1));
    if (P > 0 && cudaMemcpy(sl.data(), b->exp_slots.p + (size_t)xw * b->cap, (size_t)P * sizeof(SlotExport), cudaMemcpyDeviceToHost) != cudaSuccess)
        return -SYLT_ERR_CUDA;
    const int o0 = b->h_body_off[(size_t)world];
    int na = 0, nc = 0;
    std::vector<int> ord((size_t)P);
    for (int p = 0; p < P; p++) ord[(size_t)p] = p;
    std::sort(ord.begin(), ord.end(), [&](int x, int y) { return sl[(size_t)x].key < sl[(size_t)y].key; });
    for (int pi = 0; pi < P; pi++) {
        const SlotExport& e = sl[(size_t)ord[(size_t)pi]];
        if (e.count <= 0) continue;
        if (na >= arbiter_cap || nc + e.count > contact_cap) return -SYLT_ERR_CAPACITY;
        SyltArbiterInfo& o = arbiters[na++];
        o.body1 = (int)(e.key >> 16); o.body2 = (int)(e.key & 0xffffu);
        o.id1 = b->h_ids[(size_t)(o0 + o.body1)]; o.id2 = b->h_ids[(size_t)(o0 + o.body2)];
        o.friction = e.friction; o.num_contacts = e.count; o.contact_offset = nc; o.color = e.color;
        const bool zero_r = b->iterations == 0;
        for (int q = 0; q < e.count; q++) {
            SyltContact& c = contacts[nc++];
            memset(&c, 0, sizeof c);
            c.px = e.geo[q][0]; c.py = e.geo[q][1]; c.separation = e.geo[q][2]; c.nx = e.nx; c.ny = e.ny;
            if (!zero_r) { c.r1x = e.sol[q][0]; c.r1y = e.sol[q][1]; c.r2x = e.sol[q][2]; c.r2y = e.sol[q][3]; }
            c.mass_normal = e.sol[q][4]; c.mass_tangent = e.sol[q][5]; c.bias = e.sol[q][6]; c.pn = e.sol[q][7]; c.pt = e.sol[q][8];
            c.in_edge_1 = e.feat[q] & 0xff; c.out_edge_1 = (e.feat[q] >> 8) & 0xff; c.in_edge_2 = (e.feat[q] >> 16) & 0xff;
            c.out_edge_2 = (e.feat[q] >> 24) & 0xff;
        }
    }
    return na;
}

int sylt_batch_get_joints(SyltBatch* b, int32_t world, SyltJointState* out, int32_t cap) {
    if (!b || world < 0 || world >= b->n_worlds) return -SYLT_ERR_INVALID;
    if (cudaSetDevice(b->device) != cudaSuccess) return -SYLT_ERR_CUDA;
    const int q0 = b->h_joint_off[(size_t)world], J = std::min<int>(cap, b->h_joint_off[(size_t)world + 1] - q0);
    if (J <= 0) return 0;
    std::vector<float2> p((size_t)J), bs((size_t)J);
    std::vector<float4> r((size_t)J), m((size_t)J);
    cudaStreamSynchronize(b->stream);
    cudaMemcpy(p.data(), b->jp.p + q0, (size_t)J * sizeof(float2), cudaMemcpyDeviceToHost);
    cudaMemcpy(bs.data(), b->jbias.p + q0, (size_t)J * sizeof(float2), cudaMemcpyDeviceToHost);
    cudaMemcpy(r.data(), b->jr.p + q0, (size_t)J * sizeof(float4), cudaMemcpyDeviceToHost);
    if (cudaMemcpy(m.data(), b->jm.p + q0, (size_t)J * sizeof(float4), cudaMemcpyDeviceToHost) != cudaSuccess) return -SYLT_ERR_CUDA;
    for (int j = 0; j < J; j++) {
        const SyltBatch::HJ& h = b->h_joints[(size_t)(q0 + j)];
        SyltJointState s;
        s.px = p[(size_t)j].x; s.py = p[(size_t)j].y; s.bias_x = bs[(size_t)j].x; s.bias_y = bs[(size_t)j].y;
        s.r1x = r[(size_t)j].x; s.r1y = r[(size_t)j].y; s.r2x = r[(size_t)j].z; s.r2y = r[(size_t)j].w;
        s.m11 = m[(size_t)j].x; s.m12 = m[(size_t)j].y; s.m21 = m[(size_t)j].z; s.m22 = m[(size_t)j].w;
        s.bias_factor = h.bias_factor; s.softness = h.softness; s.la1x = h.la1x; s.la1y = h.la1y; s.la2x = h.la2x; s.la2y = h.la2y;
        s.body1 = h.b1; s.body2 = h.b2;
        out[j] = s;
    }
    return J;
}

int sylt_batch_get_joint_order(SyltBatch* b, int32_t world, int32_t* out, int32_t cap) {
    if (!b || world < 0 || world >= b->n_worlds) return -SYLT_ERR_INVALID;
    const int q0 = b->h_joint_off[(size_t)world], J = std::min<int>(cap, b->h_joint_off[(size_t)world + 1] - q0);
    for (int j = 0; j < J; j++) out[j] = b->h_jorder[(size_t)(q0 + j)];
    return std::max(J, 0);
}

int sylt_batch_get_errors(SyltBatch* b, int32_t* out, int32_t cap) {
    if (!b || !out) return -SYLT_ERR_INVALID;
    if (cudaSetDevice(b->device) != cudaSuccess) return -SYLT_ERR_CUDA;
    int n = std::min<int>(cap, b->n_worlds);
    cudaStreamSynchronize(b->stream);
    if (cudaMemcpy(out, b->w_err.p, (size_t)n * sizeof(int), cudaMemcpyDeviceToHost) != cudaSuccess) return -SYLT_ERR_CUDA;
    return n;
}

int sylt_batch_stats(SyltBatch* b, SyltBatchStats* out) {
    if (!b || !out) return SYLT_ERR_INVALID;
    SYLT_CUDA(cudaSetDevice(b->device));
    SYLT_CUDA(cudaMemsetAsync(b->stats.p, 0, sizeof(StatsOut), b->stream));
    k_batch_stats<<<std::min(1024, (b->n_worlds + 255) / 256), 256, 0, b->stream>>>(b->n_worlds, b->body_off.p, b->vel.p, b->mp.p, b->mass_moi.p,
                                                                                    b->w_err.p, b->w_ncon.p, b->w_maxpen.p, b->stats.p);
    b->launches++;
    StatsOut so;
    SYLT_CUDA(cudaMemcpyAsync(&so, b->stats.p, sizeof so, cudaMemcpyDeviceToHost, b->stream));
    SYLT_CUDA(cudaStreamSynchronize(b->stream));
    memset(out, 0, sizeof *out);
    out->body_steps = b->body_steps; out->contacts = so.contacts; out->errors = so.errors; out->kinetic_energy = so.ke;
    out->max_penetration = so.maxpen;
    return SYLT_OK;
}

}  // extern "C"

#ifdef SYLT_BATCH_TIMING
extern "C" int sylt_batch_debug_phases(unsigned long long* out16, int reset) {
    cudaDeviceSynchronize();
    if (out16) cudaMemcpyFromSymbol(out16, g_phase_cycles, sizeof(unsigned long long) * 16);
    if (reset) { unsigned long long z[16] = {0}; cudaMemcpyToSymbol(g_phase_cycles, z, sizeof(z)); }
    return 0;
}
#endif
