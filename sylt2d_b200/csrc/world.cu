// world.cu — the single (large) world: World::step on one B200.
//
// Restates the behaviour of /root/reference/src/world.rs:73-152 (broad_phase + step) with a different
// algorithmic structure (nothing here is a translation of the Rust loops):
//
//   k_prepare      body        (cos,sin) cache, conservative AABB, hashed uniform-grid cell, force integration
//                              (world.rs:113-120); zeroes the per-step counters
//   k_cells_local  bucket      counting-sort offsets of the grid: exclusive offsets inside tiles of 2048 buckets + one total
//                              per tile (no device-wide scan: the consumer adds up the tile totals)
//   k_fill_cells   body        bodies grouped by cell (a cell-ordered copy of index / cell / flags / AABB)
//   k_rows_count   body        candidate rows, one warp per body: for body i every AABB-overlapping partner it "owns";
//                              static-static pairs skipped (world.rs:79-81), other collision groups skipped.  Large bodies
//                              (ground, walls) bypass the grid through a short list.  Leaves one total per 128 bodies.
//   k_rows_emit    body        rows laid out in body order (prefix = block totals in front + in-block scan), partner-sorted
//   k_narrow       pair        one thread per candidate pair: box_box / poly_poly (narrow.cuh); one total per 256 pairs
//   k_build        pair        compaction into the arbiter / contact SoA (offsets = tile totals in front + in-tile scan);
//                              match against last step's arbiters (Arbiter::update, arbiter.rs:137-181 incl. quirk Q-2;
//                              friction persistence Q-17); persisting arbiters keep their colour
//   k_solve        cooperative persistent kernel: incremental graph colouring of new arbiters, then
//                              Arbiter::pre_step / Joint::pre_step / iterations x apply_impulse in colour order as a
//                              version-gated dataflow sweep through L2 (arbiter.rs:182-298, joint.rs:57-130), then
//                              position integration (world.rs:143-150).  Worlds with joints, small worlds.
//   k_solve_regions            the same sweep for joint-free worlds of >= 1000 bodies: bodies partitioned into regions
//                              (k_recut, every 128 steps) that live in shared memory, cross-region arbiters replayed
//                              redundantly from ghost bodies (see the comment at the kernel)
//
// Gauss-Seidel order: arbiter colours ascending, then joint colours ascending (within a colour constraints touch
// disjoint dynamic bodies, so the result equals the sequential sweep in that order bit for bit; static bodies
// are never written).  The order is reported through sylt_world_get_arbiters(...).color / get_joint_order so the
// oracle can replay it.
#include <cooperative_groups.h>
#include <algorithm>
#include <cstdlib>
#include <mutex>
#include <thread>
#include "common.h"
#include "narrow.cuh"
#include "solver.cuh"

namespace cg = cooperative_groups;
using namespace sylt;

namespace sylt {
static thread_local std::string g_last_error;
static std::mutex g_err_mu;
static std::string g_last_error_global;
void set_last_error(const std::string& s) {
    g_last_error = s;
    std::lock_guard<std::mutex> l(g_err_mu);
    g_last_error_global = s;
}
}  // namespace sylt

namespace {

constexpr int NC = 64;          // max arbiter colours (one bit each in the per-body 64-bit `used` mask)
constexpr int GEN_BIT = 32;     // region solver: colours [0, GEN_BIT) are interior colours (bits of `used_int`), colour GEN_BIT + k is bit k of `used`
constexpr int NCX = NC + GEN_BIT;
constexpr int MAX_REGIONS = 1024;  // region solver: CTAs of the cooperative launch
constexpr int MAX_ROUNDS = 256; // colouring rounds per step (only new arbiters take part)
constexpr int FLAG_POLY = 1, FLAG_LARGE = 2, FLAG_STATIC = 4;

struct Counters {
    // sticky across the steps of one step_n call (reset by do_steps).  Once err is set every kernel of the following steps
    // returns at once: a failed step changes nothing (bodies, forces, last step's arbiters), later steps of the call do not run.
    int err, err_joint;
    float badk[4];
    int steps_done;  // steps of this call that completed
    int good_n_arb, good_n_con;  // arbiters / contacts of the last completed step (the set a repeated step matches against)
    unsigned neg_zero_tag;       // serial of the last step in which k_prepare saw a dynamic body with a velocity component of exactly -0.0
    // per step (reset by enqueue_step):
    int n_pairs, n_arb, n_con;
    int n_colors, rounds, n_uncolored, warn_ext;
    int color_count[NCX];
    int color_start[NC + 1];
    int color_fill[NC];
    int remaining[MAX_ROUNDS];
    unsigned bar;  // grid barrier arrivals of this step's k_solve
    int jfail1;    // J - (index of the first joint in Vec order whose K is singular), 0 = none (world.rs:127-129)
    // region solver (k_solve_regions):
    int n_generic;                  // arbiters of the generic colours (cross-region pairs, interior overflow)
    int n_ghost;                    // replicas of generic arbiters replayed by neighbouring regions
    int max_slots, n_nonres;        // load of the busiest CTA; slots whose records / contacts did not fit into shared memory
    int why;                        // diagnostics: bit mask of what overflowed in k_solve_regions
    int n_promoted;                 // new interior arbiters that found no free interior colour (they take a generic one, grid-wide)
    int bin_overflow;               // a region has more new arbiters than its bin holds: this step is coloured by the grid-wide rounds
    int region_new[MAX_REGIONS + 1];  // new arbiters per owner region (interior ones; the cross-region ones are counted at [n_regions])
    unsigned long long color_mask;  // generic colours in use
    unsigned color_mask_int;        // interior colours in use
};

// ------------------------------------------------------------------ programmatic dependent launch
// The kernels of a step form one dependent chain of short launches.  Launched with the programmatic-serialization
// attribute, the next kernel's CTAs may be scheduled while the previous kernel is still draining; pdl_wait() at the top of
// every kernel is the actual dependency (it returns once the previous kernel has completed and its writes are visible),
// pdl_trigger() lets the successor start scheduling early.  Without the attribute both are no-ops.
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
// Once a step has failed, the kernels that would change persistent state (k_build: the arbiter sets; the solvers: bodies and
// impulses; the counters) return at once.  The streaming kernels in front of them only write per-step scratch and are NOT gated: a
// gate is a dependent load at the top of every short-lived warp (measured: +8 us on k_rows_count at 100k bodies).
__device__ __forceinline__ bool step_failed(const Counters* cn) { return *(volatile const int*)&cn->err != 0; }
__device__ __forceinline__ void step_completed(Counters* cn) {  // one thread, at the very end of a step
    cn->steps_done++;
    cn->good_n_arb = cn->n_arb;
    cn->good_n_con = cn->n_con;
}
template <typename... KArgs, typename... Args>
cudaError_t launch_pdl(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args&&... args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr; cfg.numAttrs = 1;
    return cudaLaunchKernelEx(&cfg, kern, static_cast<KArgs>(args)...);
}

// 32-byte records move as ONE request (sm_100: 256-bit global loads / stores; p must be 32-byte aligned)
struct Rec32 { int4 a, b; };
__device__ __forceinline__ Rec32 ld256(const void* p) {
    Rec32 r;
    asm volatile("ld.global.v8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                 : "=r"(r.a.x), "=r"(r.a.y), "=r"(r.a.z), "=r"(r.a.w), "=r"(r.b.x), "=r"(r.b.y), "=r"(r.b.z), "=r"(r.b.w) : "l"(p));
    return r;
}
__device__ __forceinline__ void st256(void* p, const int4 a, const int4 b) {
    asm volatile("st.global.v8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};"
                 ::"l"(p), "r"(a.x), "r"(a.y), "r"(a.z), "r"(a.w), "r"(b.x), "r"(b.y), "r"(b.z), "r"(b.w) : "memory");
}
__device__ __forceinline__ float4 as_f4(const int4 v) { return make_float4(__int_as_float(v.x), __int_as_float(v.y), __int_as_float(v.z), __int_as_float(v.w)); }
__device__ __forceinline__ int4 as_i4(const float4 v) { return make_int4(__float_as_int(v.x), __float_as_int(v.y), __float_as_int(v.z), __float_as_int(v.w)); }

// ------------------------------------------------------------------ block-level scans
template <typename T>
__device__ __forceinline__ T warp_incl_scan(T v) {
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        T u = __shfl_up_sync(0xffffffffu, v, o);
        if ((threadIdx.x & 31) >= o) v += u;
    }
    return v;
}
// exclusive scan of one value per thread across the block; returns the exclusive prefix, *total = block sum
template <typename T>
__device__ __forceinline__ T block_excl_scan(T v, T* total, T* sh /* >= 33 */) {
    T inc = warp_incl_scan(v);
    int w = threadIdx.x >> 5, l = threadIdx.x & 31;
    if (l == 31) sh[w] = inc;
    __syncthreads();
    if (w == 0) {
        T s = (l < (int)(blockDim.x >> 5)) ? sh[l] : T(0);
        T si = warp_incl_scan(s);
        sh[l] = si - s;
        if (l == 31) sh[32] = si;
    }
    __syncthreads();
    T r = sh[w] + inc - v;
    *total = sh[32];
    __syncthreads();
    return r;
}

// The two prefix sums of the candidate chain (rows -> pairs, pairs -> arbiters / contacts) have no kernel of their own: the
// producer leaves one total per block of ROW_BLK bodies / PAIR_TILE candidates, and every consumer block adds up the totals in
// front of it (a few hundred coalesced loads) before the exclusive scan of its own block.
constexpr int ROW_BLK = 128;     // bodies per k_rows_emit block
constexpr int PAIR_TILE = 256;   // candidates per tile of k_narrow / k_build

// ------------------------------------------------------------------ broad phase
struct GridParams {
    float cell;
    int mask;  // buckets - 1
    float inv_cell;
};
// (any monotone function of x works, as long as the bodies are binned and searched with the same one: a multiplication, not a division)
__device__ __forceinline__ int cell_coord(float x, float inv_cell) {
    float q = floorf(x * inv_cell);
    q = fminf(fmaxf(q, -1.0e9f), 1.0e9f);
    return (int)q;
}
__device__ __forceinline__ int cell_hash(int cx, int cy, int mask, int group = 0) {
    return (int)(((unsigned)cx * 73856093u) ^ ((unsigned)cy * 19349663u) ^ ((unsigned)group * 2654435761u)) & mask;
}
// Collision groups (SyltBodyDesc.group): bodies of different groups never pair, i.e. one SyltWorld can hold many independent worlds
// (boxes AND polygons) that share gravity / iterations / context and are stepped by the same launches.  The group sits in the upper
// half of the flag word that travels with every cell entry, so the candidate test costs one more compare.
constexpr int GROUP_SHIFT = 16;
__device__ __forceinline__ int group_of(int flags) { return (int)((unsigned)flags >> GROUP_SHIFT); }
__device__ __forceinline__ bool aabb_overlap(float4 a, float4 b) {
    return !(a.x > b.z || b.x > a.z || a.y > b.w || b.y > a.w);
}

// per-step counters back to zero — unless an earlier step of the call failed: its counts are what the host grows the buffers from.
// Done by block 0 of k_prepare (nothing reads or writes them before the kernels behind it).
__device__ __forceinline__ void reset_step_counters(Counters* cn) {
    if (step_failed(cn)) return;
    int* p = &cn->n_pairs;
    const int n = (int)((sizeof(Counters) - offsetof(Counters, n_pairs)) / sizeof(int));
    for (int k = threadIdx.x; k < n; k += blockDim.x) p[k] = 0;
}

__global__ void __launch_bounds__(256) k_prepare(int B, float4* pose, const float4* vel, const float4* force, const float4* mp, const float4* geom,
                                                 const int* bflags, const float2* lverts, float2* rotc, float4* aabb, int4* cinfo,
                                                 int* cell_count, int* rank, GridParams gp, float gx, float gy, float dt, int integrate,
                                                 Counters* cn, ulonglong2* vrec, unsigned long long* used, unsigned* used_int, int* row_btot, unsigned step_serial) {
    pdl_trigger();
    if (blockIdx.x == 0) reset_step_counters(cn);
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= B) return;
    if ((i & (ROW_BLK - 1)) == 0) row_btot[i / ROW_BLK] = 0;  // candidates per block of bodies: accumulated by k_rows_count
    used[i] = 0ull;  // colours taken at this body: rebuilt by k_build / the colouring rounds every step
    used_int[i] = 0u;
    float4 p = pose[i];
    float4 m = mp[i];
    float4 g = geom[i];
    int fl = bflags[i];
    float c = sylt_cosf(p.z), s = sylt_sinf(p.z);
    rotc[i] = make_float2(c, s);
    {
        float4 v = vel[i];
        if (integrate && m.x != 0.0f) {  // world.rs:113-120
            float4 f = force[i];
            V2 nv = mk(v.x, v.y) + (mk(gx, gy) + mk(f.x, f.y) * m.x) * dt;
            v.x = nv.x; v.y = nv.y;
            v.z += m.y * f.z * dt;
            // NOT written to vel[]: the solver reads and writes velocities through vrec and stores vel[] when the step has
            // succeeded, so a step that fails later on (capacity) leaves the bodies untouched
        }
        // (a warm start with all-zero impulses — warm starting off, the World default — leaves every velocity as it is, bit for bit,
        // unless one is exactly -0.0: then the solver has to run its ordered pre-step sweep to get the signs of the zeros right)
        if (m.x != 0.0f && (__float_as_uint(v.x) == 0x80000000u || __float_as_uint(v.y) == 0x80000000u || __float_as_uint(v.z) == 0x80000000u))
            cn->neg_zero_tag = step_serial;
        // versioned velocity record of the dataflow solver: (vx|ver, vy|ver), (w|ver, inv_mass|inv_moi), ver = 0
        vrec[2 * (size_t)i] = make_ulonglong2((unsigned long long)__float_as_uint(v.x), (unsigned long long)__float_as_uint(v.y));
        vrec[2 * (size_t)i + 1] = make_ulonglong2((unsigned long long)__float_as_uint(v.z),
                                                  (unsigned long long)__float_as_uint(m.x) | ((unsigned long long)__float_as_uint(m.y) << 32));
    }
    float minx, miny, maxx, maxy;
    if (fl & FLAG_POLY) {
        int off = __float_as_int(g.z), nv = (fl >> 8) & 0xff;
        M2 r = rot_cs(c, s);
        minx = miny = 3.402823466e+38f; maxx = maxy = -3.402823466e+38f;
        for (int k = 0; k < nv; k++) {
            float2 lv = lverts[off + k];
            V2 w = r * mk(lv.x, lv.y) + mk(p.x, p.y);
            minx = fminf(minx, w.x); maxx = fmaxf(maxx, w.x);
            miny = fminf(miny, w.y); maxy = fmaxf(maxy, w.y);
        }
    } else {
        float hx = g.x * 0.5f, hy = g.y * 0.5f;
        float ex = fabsf(c) * hx + fabsf(s) * hy, ey = fabsf(s) * hx + fabsf(c) * hy;
        minx = p.x - ex; maxx = p.x + ex; miny = p.y - ey; maxy = p.y + ey;
    }
    // conservative margin: a few ulps of the absolute coordinates plus a relative slack on the extent
    float mg = 1.0e-6f * (fabsf(p.x) + fabsf(p.y)) + 1.0e-5f * ((maxx - minx) + (maxy - miny)) + 1.0e-6f;
    minx -= mg; miny -= mg; maxx += mg; maxy += mg;
    aabb[i] = make_float4(minx, miny, maxx, maxy);
    int cx = 0, cy = 0;
    if (!(fl & FLAG_LARGE)) {
        cx = cell_coord(0.5f * (minx + maxx), gp.inv_cell);
        cy = cell_coord(0.5f * (miny + maxy), gp.inv_cell);
        int h = cell_hash(cx, cy, gp.mask, group_of(fl));
        rank[i] = atomicAdd(&cell_count[h], 1);
    }
    cinfo[i] = make_int4(cx, cy, fl, 0);
}

// bodies grouped by cell: a cell-ordered COPY of (index, cell, flags) and of the AABB, so that the row search streams
// contiguous 32-byte entries instead of chasing body indices
// Counting-sort offsets of the grid without a device-wide scan: k_cells_local leaves, per tile of CELL_TILE buckets, the exclusive
// offsets inside the tile and the tile's total; k_fill_cells adds up the (few hundred) totals in its prologue.
constexpr int CELL_TILE = 2048, MAX_CELL_TILES = 2048;  // (the bucket count stops growing at CELL_TILE * MAX_CELL_TILES)
__global__ void __launch_bounds__(256) k_cells_local(int buckets, int* cell_count, int* cell_local, int* tile_tot) {
    __shared__ int sh[34];
    pdl_wait();
    pdl_trigger();
    const int base = blockIdx.x * CELL_TILE + threadIdx.x * 8;
    int v[8], sum = 0;
    if (base + 8 <= buckets) {
        const int4 a = *(const int4*)(cell_count + base), b = *(const int4*)(cell_count + base + 4);
        v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
        *(int4*)(cell_count + base) = make_int4(0, 0, 0, 0);  // the counters are ready for the next step
        *(int4*)(cell_count + base + 4) = make_int4(0, 0, 0, 0);
    } else {
#pragma unroll
        for (int k = 0; k < 8; k++) { v[k] = base + k < buckets ? cell_count[base + k] : 0; if (base + k < buckets) cell_count[base + k] = 0; }
    }
#pragma unroll
    for (int k = 0; k < 8; k++) sum += v[k];
    int tot;
    int ex = block_excl_scan<int>(sum, &tot, sh);
#pragma unroll
    for (int k = 0; k < 8; k++) { if (base + k < buckets) cell_local[base + k] = ex; ex += v[k]; }
    if (threadIdx.x == 0) tile_tot[blockIdx.x] = tot;
}

// Bodies into their cells: a cell-ordered COPY of (index, cell, flags) and of the AABB, so that the row search streams contiguous
// 32-byte entries instead of chasing body indices.  The large bodies (which bypass the grid) get entries of the same kind behind
// the small ones, in the order of the `large` list.  Also turns the tile-local bucket offsets into absolute ones for the row kernels.
__global__ void __launch_bounds__(256) k_fill_cells(int B, const int4* cinfo, const float4* aabb, const int* rank, const int* cell_local, const int* tile_tot,
                                                    int ntiles, int buckets, int* cell_start, int4* cent_info, float4* cent_aabb, GridParams gp,
                                                    const int* large, int n_large) {
    __shared__ int s_tp[MAX_CELL_TILES + 1];  // buckets in front of tile t
    __shared__ int sh[34];
    pdl_wait();
    pdl_trigger();
    {
        const int per = (ntiles + (int)blockDim.x - 1) / (int)blockDim.x, t0 = min((int)threadIdx.x * per, ntiles), t1 = min(t0 + per, ntiles);
        int sum = 0;
        for (int t = t0; t < t1; t++) sum += tile_tot[t];
        int tot;
        int ex = block_excl_scan<int>(sum, &tot, sh);
        for (int t = t0; t < t1; t++) { s_tp[t] = ex; ex += tile_tot[t]; }
        if (threadIdx.x == 0) s_tp[ntiles] = tot;
    }
    __syncthreads();
    const int tid = blockIdx.x * blockDim.x + threadIdx.x, nth = gridDim.x * blockDim.x;
    for (int h = tid; h <= buckets; h += nth) cell_start[h] = h < buckets ? s_tp[h / CELL_TILE] + cell_local[h] : s_tp[ntiles];
    const int i = tid;
    if (i < n_large) {  // the i-th large body
        const int j = large[i];
        const int4 cj = cinfo[j];
        const int pos = B - n_large + i;
        st256(&cent_info[2 * pos], make_int4(j, cj.x, cj.y, cj.z), as_i4(aabb[j]));
    }
    if (i >= B) return;
    const int4 ci = cinfo[i];
    if (ci.z & FLAG_LARGE) return;
    const int h = cell_hash(ci.x, ci.y, gp.mask, group_of(ci.z));
    const int pos = s_tp[h / CELL_TILE] + cell_local[h] + rank[i];
    st256(&cent_info[2 * pos], make_int4(i, ci.x, ci.y, ci.z), as_i4(aabb[i]));  // info and AABB of an entry: one 32-byte sector, one request
}

// Row i = partners body i owns: small-small pairs belong to the lower index, small-large pairs to the small
// body, large-large pairs to the lower index.  Partners are emitted sorted ascending.
// The count pass remembers the first ROWTMP partners it finds so that the emit pass does not repeat the grid walk.
constexpr int ROWTMP = 16;
constexpr int WIDE_WINDOW = 64;  // a search window of more cells than this is searched by walking the (cell-ordered) entries instead

template <typename F>
__device__ __forceinline__ void row_search(int i, int B, float4 my, int4 ci, const float4* aabb, const int4* cinfo, const int* cell_start,
                                           const int4* cent_info, const float4* cent_aabb, const int* large_all, const int* large_start, int n_small,
                                           GridParams gp, F&& found) {
    const bool st_i = (ci.z & FLAG_STATIC) != 0;
    const int grp = group_of(ci.z);
    const int* large = large_all + large_start[grp];                       // the large bodies of this body's group
    const int n_large = large_start[grp + 1] - large_start[grp];
    if (!(ci.z & FLAG_LARGE)) {
        int x0 = cell_coord(my.x - 0.5f * gp.cell, gp.inv_cell), x1 = cell_coord(my.z + 0.5f * gp.cell, gp.inv_cell);
        int y0 = cell_coord(my.y - 0.5f * gp.cell, gp.inv_cell), y1 = cell_coord(my.w + 0.5f * gp.cell, gp.inv_cell);
        if (x1 - x0 <= 2 && y1 - y0 <= 2) {
            // the usual 3x3 block: fetch all bucket ranges first (independent loads), then stream the entries
            int k0[9], k1[9];
#pragma unroll
            for (int q = 0; q < 9; q++) {
                int cx = x0 + q % 3, cy = y0 + q / 3;
                bool ok = cx <= x1 && cy <= y1;
                int h = cell_hash(cx, cy, gp.mask, grp);
                k0[q] = ok ? cell_start[h] : 0;
                k1[q] = ok ? cell_start[h + 1] : 0;
            }
#pragma unroll
            for (int q = 0; q < 9; q++) {
                int cx = x0 + q % 3, cy = y0 + q / 3;
                for (int k = k0[q]; k < k1[q]; k++) {
                    int4 cj = cent_info[2 * k];
                    float4 ab = cent_aabb[2 * k];
                    if (cj.x <= i || cj.y != cx || cj.z != cy || group_of(cj.w) != grp) continue;  // lower index owns the pair; another cell (or group) may share the bucket
                    if (st_i && (cj.w & FLAG_STATIC)) continue;
                    if (aabb_overlap(my, ab)) found(cj.x);
                }
            }
        } else if ((long long)(x1 - x0 + 1) * (long long)(y1 - y0 + 1) > WIDE_WINDOW) {
            // a window of many cells (a body at huge coordinates: the AABB margin grows with |x|): walk the entries, not the cells
            for (int k = 0; k < n_small; k++) {
                const int4 cj = cent_info[2 * k];
                if (cj.x <= i || group_of(cj.w) != grp) continue;
                if (st_i && (cj.w & FLAG_STATIC)) continue;
                if (aabb_overlap(my, cent_aabb[2 * k])) found(cj.x);
            }
        } else {
            for (int cy = y0; cy <= y1; cy++)
                for (int cx = x0; cx <= x1; cx++) {
                    int h = cell_hash(cx, cy, gp.mask, grp);
                    int ke = cell_start[h + 1];
                    for (int k = cell_start[h]; k < ke; k++) {
                        int4 cj = cent_info[2 * k];
                        if (cj.x <= i || cj.y != cx || cj.z != cy || group_of(cj.w) != grp) continue;
                        if (st_i && (cj.w & FLAG_STATIC)) continue;
                        if (aabb_overlap(my, cent_aabb[2 * k])) found(cj.x);
                    }
                }
        }
        for (int k = 0; k < n_large; k++) {
            int j = large[k];
            if (st_i && (cinfo[j].z & FLAG_STATIC)) continue;
            if (!aabb_overlap(my, aabb[j])) continue;
            found(j);
        }
    } else {
        for (int k = 0; k < n_large; k++) {
            int j = large[k];
            if (j <= i) continue;
            if (st_i && (cinfo[j].z & FLAG_STATIC)) continue;
            if (!aabb_overlap(my, aabb[j])) continue;
            found(j);
        }
    }
}

// Count pass: one warp per body, one lane per candidate ENTRY.  In the usual case (the body's search window is a 3 x 3 block of
// cells) lanes 0..8 fetch the bucket ranges of the nine cells and lane 9 stands for the large list; a prefix over those ten
// counts numbers all entries of the window, every lane takes entry base + lane (one 32-byte load, one test) and a ballot
// packs the hits into row_tmp — two dependent memory round trips per body, no per-lane loops, no divergence.  Anything else
// (a large body, a wider window) is searched by lane 0 alone.  Bodies are taken in cell order, so neighbouring warps search the
// same buckets.  (Measured at 100k bodies: one thread per body 36 us, 16 lanes with a cell each 49 us — issue-bound, 32 M warp
// instructions —, 4 lanes 39 us with 40 % of the lanes active.)
template <int LPB>  // lanes per body: 16 (two bodies per warp — the kernel is bound by the number of warp lifetimes, see above) or 32
__global__ void __launch_bounds__(128) k_rows_count(int B, const float4* aabb, const int4* cinfo, const int* cell_start, const int4* cent_info,
                                                    const float4* cent_aabb, const int* large_all, int n_large_all, const int* large_start, GridParams gp,
                                                    int* row_count, int* row_tmp, const Counters* cn, int* row_btot) {
    pdl_wait();
    pdl_trigger();
    static_assert(LPB == 16 || LPB == 32, "one lane per segment (9 cells + the large list) is needed");
    const int g = (blockIdx.x * blockDim.x + threadIdx.x) / LPB, lane = threadIdx.x & (LPB - 1);   // lane: inside the body's group of lanes
    const int gshift = (threadIdx.x & 31) & ~(LPB - 1);                                              // first lane of the group inside its warp
    const unsigned gmask = LPB == 32 ? 0xffffffffu : (0xffffu << gshift);
    if (g >= B) return;
    int i;
    float4 my;
    int4 ci;
    if (g < B - n_large_all) {  // the cell-ordered entry already carries the body's index, cell, flags and AABB: one load level less
        const Rec32 own = ld256(&cent_info[2 * g]);
        const int4 ce = own.a;
        my = as_f4(own.b);
        i = ce.x;
        ci = make_int4(ce.y, ce.z, ce.w, 0);
    } else {
        i = large_all[g - (B - n_large_all)];
        my = aabb[i];
        ci = cinfo[i];
    }
    const bool st_i = (ci.z & FLAG_STATIC) != 0;
    const int grp = group_of(ci.z);
    const int large_base = B - n_large_all;  // the entries of the large bodies follow those of the small ones (k_fill_cells)
    int* tmp = row_tmp + (size_t)i * ROWTMP;
    int x0 = 0, x1 = 0, y0 = 0, y1 = 0;
    bool fast = false;
    if (!(ci.z & FLAG_LARGE)) {
        x0 = cell_coord(my.x - 0.5f * gp.cell, gp.inv_cell); x1 = cell_coord(my.z + 0.5f * gp.cell, gp.inv_cell);
        y0 = cell_coord(my.y - 0.5f * gp.cell, gp.inv_cell); y1 = cell_coord(my.w + 0.5f * gp.cell, gp.inv_cell);
        fast = x1 - x0 <= 2 && y1 - y0 <= 2;
    }
    if (!fast && !(ci.z & FLAG_LARGE) && (long long)(x1 - x0 + 1) * (long long)(y1 - y0 + 1) > WIDE_WINDOW) {
        // a body at huge coordinates (the AABB margin grows with |x|, and so does the window): all lanes walk the entries of the small
        // bodies, then the large list of the group — O(B) for this body, but no million-cell window (an exploding world stays steppable)
        int found = 0;
        const int l0 = large_start[grp], l1 = large_start[grp + 1];
        const int n_all = large_base + (l1 - l0);
        for (int base = 0; base < n_all; base += LPB) {
            const int e = base + lane;
            bool hit = false;
            int partner = 0;
            if (e < n_all) {
                const int k = e < large_base ? e : large_base + l0 + (e - large_base);
                const Rec32 ent = ld256(&cent_info[2 * k]);
                const int4 cj = ent.a;
                const float4 ab = as_f4(ent.b);
                const bool mine = e < large_base ? (cj.x > i && group_of(cj.w) == grp) : true;
                hit = mine && !(st_i && (cj.w & FLAG_STATIC)) && aabb_overlap(my, ab);
                partner = cj.x;
            }
            const unsigned m = (__ballot_sync(gmask, hit) >> gshift) & (LPB == 32 ? 0xffffffffu : 0xffffu);
            const int pos = found + __popc(m & ((1u << lane) - 1u));
            if (hit && pos < ROWTMP) tmp[pos] = partner;
            found += __popc(m);
        }
        if (lane == 0) {
            row_count[i] = found;
            if (found) atomicAdd(&row_btot[i / ROW_BLK], found);
        }
        return;
    }
    if (!fast) {  // uniform over the body's lanes
        if (lane == 0) {
            int cnt = 0;
            row_search(i, B, my, ci, aabb, cinfo, cell_start, cent_info, cent_aabb, large_all, large_start, large_base, gp, [&](int j) {
                if (cnt < ROWTMP) tmp[cnt] = j;
                cnt++;
            });
            row_count[i] = cnt;
            if (cnt) atomicAdd(&row_btot[i / ROW_BLK], cnt);
        }
        return;
    }
    // segments: lanes 0..8 = the cells of the window, lane 9 = the large list
    int seg_k0 = 0, seg_n = 0;
    const int cx = x0 + lane % 3, cy = y0 + lane / 3;
    if (lane < 9) {
        if (cx <= x1 && cy <= y1) {
            const int h = cell_hash(cx, cy, gp.mask, grp);
            seg_k0 = cell_start[h];
            seg_n = cell_start[h + 1] - seg_k0;
        }
    } else if (lane == 9) {  // the large bodies of this body's group
        const int l0 = large_start[grp];
        seg_k0 = large_base + l0;
        seg_n = large_start[grp + 1] - l0;
    }
    int seg_end = seg_n;  // inclusive prefix
#pragma unroll
    for (int o = 1; o < 16; o <<= 1) {
        const int u = __shfl_up_sync(gmask, seg_end, o, LPB);
        if (lane >= o) seg_end += u;
    }
    const int total = __shfl_sync(gmask, seg_end, 9, LPB);
    int found = 0;
    for (int base = 0; base < total; base += LPB) {
        const int e = base + lane;
        int seg = 0;  // the segment entry e belongs to: the number of segments (of the first nine) that end at or before e — the ends ascend
#pragma unroll
        for (int st = 8; st >= 1; st >>= 1) {
            const int v = __shfl_sync(gmask, seg_end, min(seg + st - 1, 8), LPB);
            if (seg + st <= 9 && e >= v) seg += st;
        }
        const int s_end = __shfl_sync(gmask, seg_end, seg, LPB), s_n = __shfl_sync(gmask, seg_n, seg, LPB), s_k0 = __shfl_sync(gmask, seg_k0, seg, LPB);
        bool hit = false;
        int partner = 0;
        if (e < total) {
            const int off = e - (s_end - s_n);
            {
                const int k = s_k0 + off;
                const Rec32 ent = ld256(&cent_info[2 * k]);
                const int4 cj = ent.a;
                const float4 ab = as_f4(ent.b);
                const int scx = x0 + seg % 3, scy = y0 + seg / 3;
                // a cell entry: the lower index owns the pair, another cell (or group) may share the bucket; an entry of the large list: the
                // small body owns the pair.  Static-static pairs are skipped (world.rs:79-81).
                const bool mine = seg < 9 ? (cj.x > i && cj.y == scx && cj.z == scy && group_of(cj.w) == grp) : true;
                hit = mine && !(st_i && (cj.w & FLAG_STATIC)) && aabb_overlap(my, ab);
                partner = cj.x;
            }
        }
        const unsigned m = (__ballot_sync(gmask, hit) >> gshift) & (LPB == 32 ? 0xffffffffu : 0xffffu);
        const int pos = found + __popc(m & ((1u << lane) - 1u));
        if (hit && pos < ROWTMP) tmp[pos] = partner;
        found += __popc(m);
    }
    if (lane == 0) {
        row_count[i] = found;
        if (found) atomicAdd(&row_btot[i / ROW_BLK], found);
    }
}

__global__ void __launch_bounds__(ROW_BLK) k_rows_emit(int B, const float4* aabb, const int4* cinfo, const int* cell_start, const int4* cent_info,
                                                       const float4* cent_aabb, const int* large, const int* large_start, GridParams gp, int* row_start, const int* row_tmp,
                                                       int2* pairs, int cap_pairs, Counters* cn, const int* row_count, const int* row_btot, int n_small) {
    __shared__ int sh[34];
    pdl_wait();
    pdl_trigger();
    const int failed = __ldcg(&cn->err);  // (checked where the counters are written: the load overlaps the ones below)
    const int i = blockIdx.x * ROW_BLK + threadIdx.x;
    // row_start[i] = candidates owned by the bodies in front of i: the blocks in front of this one (their totals), then this block
    int before = 0;
    for (int k = threadIdx.x; k < (int)blockIdx.x; k += ROW_BLK) before += row_btot[k];
    int blocks_before;
    block_excl_scan<int>(before, &blocks_before, sh);
    const int n = i < B ? row_count[i] : 0;
    int block_total;
    const int base = blocks_before + block_excl_scan<int>(n, &block_total, sh), end = base + n;
    if (i >= B) return;
    row_start[i] = base;
    if (i == B - 1) {  // the grand total (SyltStepStats.candidate_pairs)
        row_start[B] = end;
        if (!failed) {  // (a failed step's counts stay: the host grows the buffers from them)
            cn->n_pairs = end;
            if (end > cap_pairs) cn->err = SYLT_ERR_CAPACITY;
        }
    }
    if (n == 0) return;
    if (end > cap_pairs) {  // the row does not fit: flag and drop
        if (!failed) atomicExch(&cn->err, SYLT_ERR_CAPACITY);
        return;
    }
    if (n <= ROWTMP) {
        const int* tmp = row_tmp + (size_t)i * ROWTMP;
        int v[ROWTMP];
#pragma unroll
        for (int a = 0; a < ROWTMP; a++) v[a] = a < n ? tmp[a] : 0x7fffffff;
        for (int a = 1; a < n; a++) {  // insertion sort by partner (rows are short)
            int x = v[a], b = a - 1;
            while (b >= 0 && v[b] > x) { v[b + 1] = v[b]; b--; }
            v[b + 1] = x;
        }
        for (int a = 0; a < n; a++) pairs[base + a] = make_int2(i, v[a]);
        return;
    }
    int m = 0;
    row_search(i, B, aabb[i], cinfo[i], aabb, cinfo, cell_start, cent_info, cent_aabb, large, large_start, n_small, gp, [&](int j) { pairs[base + m++] = make_int2(i, j); });
    for (int a = 1; a < n; a++) {
        int2 x = pairs[base + a];
        int b = a - 1;
        while (b >= 0 && pairs[base + b].y > x.y) { pairs[base + b + 1] = pairs[base + b]; b--; }
        pairs[base + b + 1] = x;
    }
}

// ------------------------------------------------------------------ narrow phase: one thread per candidate pair
template <int MAXC>
__global__ void __launch_bounds__(128) k_narrow(const Counters* cn, int cap_pairs, const int2* pairs, const float4* pose, const float2* rotc,
                                                const float4* geom, const int* bflags, const float2* lverts,
                                                unsigned long long* info, float4* tmp_a, float2* tmp_b, Counters* cnw, unsigned long long* pair_btot) {
    __shared__ unsigned long long sh[34];
    pdl_wait();
    pdl_trigger();
    const int failed = __ldcg(&cn->err);  // same cache line as n_pairs: no extra round trip
    int P = min(__ldcg(&cn->n_pairs), cap_pairs);
    if (failed) return;
    const int ntiles = (P + PAIR_TILE - 1) / PAIR_TILE;
    for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    unsigned long long tile_sum = 0ull;  // (arbiters << 32 | contacts) of this thread's candidates
    for (int p = tile * PAIR_TILE + threadIdx.x; p < min(P, (tile + 1) * PAIR_TILE); p += blockDim.x) {
        int2 pr = pairs[p];
        int a = min(pr.x, pr.y), b = max(pr.x, pr.y);  // Arbiter::new: body1 = lower id (arbiter.rs:120-122)
        float4 pa = pose[a], pb = pose[b];
        float2 ra = rotc[a], rb = rotc[b];
        float4 ga = geom[a], gb = geom[b];
        int fa = bflags[a], fb = bflags[b];
        ContactOut out[MAXC];
        int n;
        if (MAXC == 2 || !((fa | fb) & FLAG_POLY)) {
            n = box_box(mk(pa.x, pa.y), ra.x, ra.y, mk(ga.x, ga.y), mk(pb.x, pb.y), rb.x, rb.y, mk(gb.x, gb.y), out);
        } else {
            constexpr int MAXV = MAXC / 2 > 0 ? MAXC / 2 : 1;
            V2 w0[MAXV], w1[MAXV];
            int n0 = (fa >> 8) & 0xff, n1 = (fb >> 8) & 0xff;
            int o0 = __float_as_int(ga.z), o1 = __float_as_int(gb.z);
            M2 r0 = rot_cs(ra.x, ra.y), r1 = rot_cs(rb.x, rb.y);
            for (int k = 0; k < n0; k++) { float2 l = lverts[o0 + k]; w0[k] = r0 * mk(l.x, l.y) + mk(pa.x, pa.y); }  // rotate + translate (body.rs:148-168)
            for (int k = 0; k < n1; k++) { float2 l = lverts[o1 + k]; w1[k] = r1 * mk(l.x, l.y) + mk(pb.x, pb.y); }
            bool ovf = false;
            n = poly_poly<MAXV>(w0, n0, w1, n1, out, &ovf);
            if (ovf) atomicExch(&cnw->err, SYLT_ERR_CAPACITY);
        }
        info[p] = n > 0 ? ((1ull << 32) | (unsigned)n) : 0ull;
        tile_sum += n > 0 ? ((1ull << 32) | (unsigned)n) : 0ull;
        for (int k = 0; k < n; k++) {
            tmp_a[(size_t)p * MAXC + k] = make_float4(out[k].px, out[k].py, out[k].nx, out[k].ny);
            tmp_b[(size_t)p * MAXC + k] = make_float2(out[k].sep, __uint_as_float(out[k].feat));
        }
    }
    unsigned long long tot;
    block_excl_scan<unsigned long long>(tile_sum, &tot, sh);
    if (threadIdx.x == 0) pair_btot[tile] = tot;
    }
}

// Arbiter arrays of one step (double buffered: "cur" is being built, "old" is last step's result)
struct ArbSet {
    int2* bodies;      // (body1, body2), body1 < body2
    int2* meta;        // (num_contacts, contact_offset)
    int2* partner;     // (partner index in the owner's row: the lookup key of the next step, first contact)
    float* friction;
    int* color;
    int* row_start;    // per body: first arbiter of its row (B+1 entries)
    float4 *ca, *cb, *cc, *cd;  // contacts: (nx,ny,r1x,r1y) (r2x,r2y,mass_n,mass_t) (bias,pn,pt,sep) (px,py,pnb,feat)
    float4* cin;       // region solver: 2 per contact, what its set-up reads: (position, normal) (separation, pn, pt, -); k_build
};

__device__ __forceinline__ void pair_owner(int a, int b, int fa, int fb, int* owner, int* partner) {
    bool la = (fa & FLAG_LARGE) != 0, lb = (fb & FLAG_LARGE) != 0;
    if (la == lb) { *owner = a; *partner = b; }  // a < b
    else if (la) { *owner = b; *partner = a; }
    else { *owner = a; *partner = b; }
}

constexpr int UNC_STATIC = 0x40000000;  // body reference of an uncoloured-arbiter record: the body is static (takes no colour)
constexpr int SR_STATIC = 0x40000000, SR_OWNER2 = 0x20000000, SR_BODY = 0x1fffffff;  // body references of a solver record (srec)
template <int MAXC>
__global__ void __launch_bounds__(256) k_build(Counters* cn, int cap_pairs, int cap_arb, int cap_con, int B, int B_old, int have_old,
                                               const int2* pairs, const unsigned long long* info, unsigned long long* scan, const unsigned long long* pair_btot,
                                               const int* row_start, const float4* tmp_a, const float2* tmp_b, const float4* mp,
                                               const int* bflags, const int* bflags_old, ArbSet cur, ArbSet old, int warm,
                                               unsigned long long* used, int4* unc_rec, int fix_features, int keep_colors, unsigned* used_int, int regions,
                                               const int* rloc, int2* gadj, int n_regions, int4* unc_bin, int* unc_bin_idx, int bin_cap, int4* srec) {
    pdl_wait();
    const int failed = __ldcg(&cn->err);
    int P = min(__ldcg(&cn->n_pairs), cap_pairs);
    if (failed) return;  // (the set "being built" by a step behind a failed one is the last good set: it must not be touched)
    int tid = blockIdx.x * blockDim.x + threadIdx.x, nth = gridDim.x * blockDim.x;
    __shared__ int s_hist[NCX];
    __shared__ unsigned long long s_sc[34];
    if (threadIdx.x < NCX) s_hist[threadIdx.x] = 0;
    __syncthreads();
    (void)tid; (void)nth;
    // Candidates in tiles of PAIR_TILE: (arbiter, contact) offset of a candidate = the tiles in front (k_narrow's totals) + the
    // exclusive scan inside the tile.  The offsets are also stored (scan[p]): the solver kernel turns them into the per-body arbiter rows.
    const int ntiles = (P + PAIR_TILE - 1) / PAIR_TILE;
    for (int tile = blockIdx.x; tile < max(ntiles, 1); tile += gridDim.x) {
    unsigned long long before = 0ull;
    for (int k = threadIdx.x; k < tile; k += blockDim.x) before += pair_btot[k];
    unsigned long long run;
    block_excl_scan<unsigned long long>(before, &run, s_sc);
    for (int p0 = tile * PAIR_TILE; p0 < (tile + 1) * PAIR_TILE; p0 += blockDim.x) {  // block-uniform trip count
        const int p = p0 + threadIdx.x;
        const unsigned long long in = p < P ? info[p] : 0ull;
        unsigned long long part;
        const unsigned long long s = run + block_excl_scan<unsigned long long>(in, &part, s_sc);
        run += part;
        if (p < P) scan[p] = s;
        if (p == P - 1 || (P == 0 && p == 0)) {  // the grand totals
            const unsigned long long total = P > 0 ? s + in : 0ull;
            scan[P] = total;
            cn->n_arb = (int)(total >> 32);
            cn->n_con = (int)(unsigned)total;
        }
        if (!in) continue;
        int n = (int)(unsigned)in;
        int a = (int)(s >> 32), coff = (int)(unsigned)s;
        if (a >= cap_arb || coff + n > cap_con) { atomicExch(&cn->err, SYLT_ERR_CAPACITY); continue; }
        int2 pr = pairs[p];
        int b1 = min(pr.x, pr.y), b2 = max(pr.x, pr.y);
        float4 m1 = mp[b1], m2 = mp[b2];
        // ---- Arbiter::update / insert (world.rs:85-98): look the pair up in last step's arbiters
        int found = -1, found_con = 0;  // last step's arbiter of the pair and its first contact (carried in the row entry: one dependent level less)
        if (have_old && b2 < B_old) {
            int oo, op;
            pair_owner(b1, b2, bflags_old[b1], bflags_old[b2], &oo, &op);
            const int lo = old.row_start[oo], hi = old.row_start[oo + 1];
            int2 part[8];  // rows are short: fetch the entries side by side (independent loads), then compare
#pragma unroll
            for (int k = 0; k < 8; k++) part[k] = lo + k < hi ? old.partner[lo + k] : make_int2(-1, 0);
#pragma unroll
            for (int k = 7; k >= 0; k--)
                if (part[k].x == op) { found = lo + k; found_con = part[k].y; }
            for (int k = lo + 8; k < hi && found < 0; k++) {
                const int2 pe = old.partner[k];
                if (pe.x == op) { found = k; found_con = pe.y; }
            }
        }
        float fr, pn = 0.0f, pt = 0.0f, pnb = 0.0f;
        int color = -1;
        const bool s1 = (bflags[b1] & FLAG_STATIC) != 0, s2 = (bflags[b2] & FLAG_STATIC) != 0;
        // region solver: (region << 16 | index inside the region) of the two bodies; pr.x owns the pair
        const int rl1 = regions ? rloc[b1] : 0, rl2 = regions ? rloc[b2] : 0;
        const int Rown = (pr.x == b1 ? rl1 : rl2) >> 16;
        const bool interior_pair = (s1 || (rl1 >> 16) == Rown) && (s2 || (rl2 >> 16) == Rown);
        if (found >= 0) {
            fr = old.friction[found];  // quirk Q-17: friction fixed at first insertion
            if (keep_colors) color = old.color[found];  // (all dropped when the region solver is switched on or off: the ranges mean something else)
            // Region solver: a colour survives if its class does — interior (all dynamic bodies in the owner's region, colour < GEN_BIT)
            // or generic.  The class of a pair changes when the regions are re-cut (only arbiters near the new borders) and when one
            // of its bodies is made static or dynamic between steps (set_body_props).
            if (regions && color >= 0 && (color < GEN_BIT) != interior_pair) color = -1;
            if (warm) {                // quirk Q-2: every new contact inherits old contact 0
                const int oc = found_con;
                float4 c = old.cc[oc];
                pn = c.y; pt = c.z; pnb = old.cd[oc].z;
            }
        } else {
            fr = sqrtf(m1.z * m2.z);  // arbiter.rs:128
        }
        cur.bodies[a] = make_int2(b1, b2);
        cur.meta[a] = make_int2(n, coff);
        cur.partner[a] = make_int2(pr.y, coff);
        cur.friction[a] = fr;
        cur.color[a] = color;
        if (regions) {  // everything the region solver's set-up needs to know about the arbiter, in one 32-byte record
            st256(&srec[2 * (size_t)a], make_int4(b1 | (s1 ? SR_STATIC : 0) | (pr.x == b2 ? SR_OWNER2 : 0), b2 | (s2 ? SR_STATIC : 0), rl1, rl2),
                  make_int4(n, coff, __float_as_int(fr), 0));
        }
        for (int k = 0; k < n; k++) {
            float4 ta = tmp_a[(size_t)p * MAXC + k];
            float2 tb = tmp_b[(size_t)p * MAXC + k];
            if (fix_features && found >= 0) {  // opt-in (NOT the reference): match old contacts by their packed feature edges
                pn = 0.0f; pt = 0.0f; pnb = 0.0f;
                int2 om = old.meta[found];
                for (int q = 0; q < om.x; q++) {
                    float4 od = old.cd[om.y + q];
                    if (__float_as_uint(od.w) == __float_as_uint(tb.y)) {
                        if (warm) { float4 oc = old.cc[om.y + q]; pn = oc.y; pt = oc.z; pnb = od.z; }
                        break;
                    }
                }
            }
            if (regions != 2) {  // (regions == 2: a full step of the region solver reads `cin` and writes these three itself)
                cur.ca[coff + k] = make_float4(ta.z, ta.w, 0.0f, 0.0f);
                cur.cb[coff + k] = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
                cur.cc[coff + k] = make_float4(0.0f, pn, pt, tb.x);
            }
            cur.cd[coff + k] = make_float4(ta.x, ta.y, pnb, tb.y);
            if (regions) {  // ... and about the contact: (position, normal | separation, pn, pt)
                st256(&cur.cin[2 * (size_t)(coff + k)], as_i4(ta), as_i4(make_float4(tb.x, pn, pt, 0.0f)));
            }
        }
        if (color >= 0) {
            const bool d1 = !s1, d2 = !s2;
            if (regions && color < GEN_BIT) {
                if (d1) atomicOr(&used_int[b1], 1u << color);
                if (d2) atomicOr(&used_int[b2], 1u << color);
            } else {
                const int bit = regions ? color - GEN_BIT : color;
                if (d1) atomicOr(&used[b1], 1ull << bit);
                if (d2) atomicOr(&used[b2], 1ull << bit);
                if (regions) {  // the per-body table of generic arbiters the region solver's cone walk reads (slot = colour bit)
                    const int2 entry = make_int2(a, Rown | (color << 16));
                    if (d1) gadj[(size_t)b1 * NC + bit] = entry;
                    if (d2) gadj[(size_t)b2 * NC + bit] = entry;
                }
            }
            atomicAdd(&s_hist[color], 1);
        } else {
            // everything the colouring rounds need to know about a new arbiter, in one 16-byte record: the rounds are chains of
            // dependent L2 round trips, and this takes two links (bodies -> flags -> regions) out of every one of them
            const int interior = !regions ? 0 : interior_pair ? Rown + 1 : -(Rown + 1);  // region solver: +-(owner region + 1), + = interior
            const int4 rec = make_int4(a, b1 | (s1 ? UNC_STATIC : 0), b2 | (s2 ? UNC_STATIC : 0), interior);
            const int ui = atomicAdd(&cn->n_uncolored, 1);
            unc_rec[ui] = rec;
            if (regions) {  // ... and once more in the bin of the region CTA that will colour it (the cross-region ones: one bin behind the regions')
                const int bin = interior > 0 ? interior - 1 : n_regions;
                const int q = atomicAdd(&cn->region_new[bin], 1);
                if (q < bin_cap) { unc_bin[(size_t)bin * bin_cap + q] = rec; unc_bin_idx[(size_t)bin * bin_cap + q] = ui; }
                else cn->bin_overflow = 1;
            }
        }
    }
    }
    __syncthreads();
    if (threadIdx.x < NCX && s_hist[threadIdx.x]) atomicAdd(&cn->color_count[threadIdx.x], s_hist[threadIdx.x]);
}

// The arbiter rows of the bodies (ArbSet::row_start: first arbiter owned by body i) from the candidate rows and the candidate ->
// arbiter offsets k_build stored; run by the solver kernels before their first grid barrier.
__device__ __forceinline__ void fill_arbiter_rows(int B, int P, const int* row_start, const unsigned long long* scan, int* arow, int tid, int nth) {
    for (int i = tid; i <= B; i += nth) {
        const int rs = i < B ? min(row_start[i], P) : P;
        arow[i] = (int)(__ldcg(&scan[rs]) >> 32);
    }
}

// ------------------------------------------------------------------ solver (cooperative persistent kernel)
__device__ __forceinline__ unsigned fmix32(unsigned h) {  // bijective mix: unique priority per arbiter index
    h ^= h >> 16; h *= 0x85ebca6bu; h ^= h >> 13; h *= 0xc2b2ae35u; h ^= h >> 16;
    return h;
}

__device__ __forceinline__ unsigned hash_u32(unsigned x) { return fmix32(x * 2654435761u); }

struct SolveArgs {
    Counters* cn;
    int B, cap_arb, cap_con, J, n_joint_colors, iterations;
    int accumulate, warm, poscorr, run_solver;
    float dt, inv_dt;
    unsigned epoch_base;
    int smem_slots, smem_contacts;  // shared-memory residency budget per CTA
    float4 *pose, *vel, *force, *pose_prev;
    const float4* mp;
    const float2* rotc;
    const int* bflags;
    ArbSet cur;
    unsigned long long* used;
    unsigned* used_int;         // region solver: interior colours taken at a body
    unsigned long long* claim;  // B * NCX
    int *uncolored, *uncolored2, *prop, *order;   // uncolored / uncolored2: the losers of a round (record indices), ping-pong; prop: by record
    const int4* unc_rec;                         // new arbiters of this step (k_build): (arbiter, body1 | UNC_STATIC, body2 | UNC_STATIC, interior)
    ulonglong2* vrec;           // 2 per body: versioned velocities (see k_prepare)
    int sleep_ns;               // back-off between polls of the dataflow wait (0 = spin)
    int fix_inverse;            // opt-in corrected Mat2x2 inverse (NOT the reference)
    int slack_pct;              // polling throttle slack (see wait_progress)
    unsigned long long* dbg;    // optional: globaltimer stamps of CTA 0 at the section boundaries (SYLT_SOLVER_TIMING=1)
    const int* jdeg;            // joints touching each body
    const int2* jrank;          // per joint: its rank among the joints of body1 / body2 (joint solve order)
    // joints
    const int2* jbodies;
    const float4* jla;       // (la1x, la1y, la2x, la2y)
    const float2* jparam;    // (softness, bias_factor)
    float2* jp;              // accumulated impulse
    float4 *jr, *jm;         // (r1, r2), M
    float2* jbias;
    const int *jorder, *jcolor_start;
    // regions (k_solve_regions): every body belongs to one region, region r is solved by CTA r
    const int *region_of, *lidx, *reg_start, *reg_bodies;  // body -> region, body -> index inside its region, CSR of the regions
    int n_regions, smem_bytes;                             // dynamic shared memory of k_solve_regions
    const int4* srec;                                      // 2 per arbiter: (b1 | flags, b2 | flags, region << 16 | local index of b1, of b2) (contacts, first contact, friction, -)
    int2* gadj;                                            // per body: its generic arbiters (arbiter, owner region | colour << 16), indexed by the colour bit
    ulonglong2* xrec;                                      // export planes (one per pass) of versioned boundary-body velocities
    unsigned step_tag, step_serial;
    int probe_cta;                                         // diagnostics: the CTA whose set-up phases are stamped
    const int4* unc_bin;                                   // new arbiters binned by owner region (k_build); bin n_regions: the cross-region ones
    const int* unc_bin_idx;                                //   their indices in unc_rec
    int bin_cap;
    int color_local;                                       // CTA-local colouring of new interior arbiters (SYLT_COLOR_LOCAL=0: grid-wide rounds only)
    // candidate rows + candidate -> (arbiter, contact) offsets of this step (fill_arbiter_rows)
    const int* cand_row_start;
    const unsigned long long* pscan;
    int cap_pairs;
};

// ---- light grid barrier: one monotonic counter, release-add on arrival, acquire-poll until everybody arrived.
// (cg::grid.sync() cost ~5 us per colour phase here: profiles/r01_k_solve_lines.txt.)  All CTAs of the cooperative
// launch are co-resident, all threads call it the same number of times; `target` lives in a register of every thread.
__device__ __forceinline__ void grid_barrier(unsigned* counter, unsigned& target) {
    target += gridDim.x;
    __syncthreads();
    if (threadIdx.x == 0) {
        asm volatile("red.release.gpu.global.add.u32 [%0], 1;" ::"l"(counter) : "memory");
        unsigned v;
        do {
            asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(counter) : "memory");
        } while (v < target);
        __threadfence();  // acquire side: one gpu-scope fence after the poll loop (also drops this SM's stale L1 lines)
    }
    __syncthreads();
}

constexpr int SOLVE_THREADS = 512;

constexpr int VSTATIC = 0x40000000;  // body reference flag: static body (never written, never waited for)

// One arbiter as the solver sees it: record + contact constants either resident in this CTA's shared memory for the
// whole solve or, when the CTA's share of the world exceeds the shared-memory budget, in the global arrays.
struct SlotView {
    int r1, r2, n;         // body index | VSTATIC
    unsigned ord;          // D1 | r1 << 8 | D2 << 16 | r2 << 24  (see "residency")
    float fr;
    float4 *ca, *cb, *cc;  // generic pointers (shared or global)
};

// Versioned velocity records in global memory (L2): per body two 16-byte halves
//   { vx | ver << 32,  vy | ver << 32 }   { w | ver << 32,  inv_mass | inv_moi << 32 }
// Every 64-bit word carries its own copy of the version, each word is read and written with a single-copy-atomic
// 64-bit access (relaxed, gpu scope: served by L2, never by L1), so a reader needs no fence: it simply retries until
// all three words show the version it expects (the protocol NCCL's LL transport uses).
struct BodyIO {
    ulonglong2* vrec;
    int sleep_ns;
    int* err;  // watchdog: a wait that never completes (a bug, never the data) flags SYLT_ERR_INTERNAL instead of hanging the GPU
    __device__ __forceinline__ static ulonglong2 ld(const ulonglong2* p) {
        ulonglong2 r;
        asm volatile("ld.relaxed.gpu.global.v2.b64 {%0, %1}, [%2];" : "=l"(r.x), "=l"(r.y) : "l"(p) : "memory");
        return r;
    }
    __device__ __forceinline__ bool try_get(int b, unsigned expect, bool check, Vel& v, float& im, float& ii) const {
        unsigned w0, w1, w2, w3, w4, w5, w6, w7;  // one 256-bit access (sm_100): a record is 32 bytes; every 64-bit word carries the version
        asm volatile("ld.relaxed.gpu.global.v8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                     : "=r"(w0), "=r"(w1), "=r"(w2), "=r"(w3), "=r"(w4), "=r"(w5), "=r"(w6), "=r"(w7) : "l"(vrec + 2 * (size_t)b) : "memory");
        v.v = mk(__uint_as_float(w0), __uint_as_float(w2));
        v.w = __uint_as_float(w4);
        im = __uint_as_float(w6);
        ii = __uint_as_float(w7);
        return !check || (w1 == expect && w3 == expect && w5 == expect);
    }
    __device__ __forceinline__ void get(int b, Vel& v, float& im, float& ii) const { try_get(b, 0u, false, v, im, ii); }
    // wait for both bodies of a constraint (static bodies are read once and never waited for)
    __device__ __forceinline__ void wait2(int r1, unsigned e1, Vel& v1, float& im1, float& ii1, int r2, unsigned e2, Vel& v2, float& im2,
                                          float& ii2) const {
        const int b1 = r1 & ~VSTATIC, b2 = r2 & ~VSTATIC;
        bool ok1 = false, ok2 = false;
        unsigned spins = 0;
        for (;;) {
            if (!ok1) ok1 = try_get(b1, e1, !(r1 & VSTATIC), v1, im1, ii1);
            if (!ok2) ok2 = try_get(b2, e2, !(r2 & VSTATIC), v2, im2, ii2);
            if (ok1 && ok2) break;
            if (sleep_ns) __nanosleep(sleep_ns);
            if ((++spins & 0xfffu) == 0u) {
                if (*(volatile int*)err == SYLT_ERR_INTERNAL) break;
                if (spins > (1u << 23)) { atomicExch(err, SYLT_ERR_INTERNAL); break; }
            }
        }
    }
    // The same records moved as ONE 256-bit access (sm_100; the export planes of k_solve_regions: the fourth word is not used there).  A
    // 32-byte access need not be single-copy atomic as a whole: every 64-bit word still carries its own copy of the version.
    __device__ __forceinline__ bool try_get256(int b, unsigned expect, Vel& v) const {
        unsigned w0, w1, w2, w3, w4, w5, w6, w7;
        asm volatile("ld.relaxed.gpu.global.v8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                     : "=r"(w0), "=r"(w1), "=r"(w2), "=r"(w3), "=r"(w4), "=r"(w5), "=r"(w6), "=r"(w7) : "l"(vrec + 2 * (size_t)b) : "memory");
        v.v = mk(__uint_as_float(w0), __uint_as_float(w2));
        v.w = __uint_as_float(w4);
        return w1 == expect && w3 == expect && w5 == expect;
    }
    __device__ __forceinline__ void put256(int b, const Vel& v, unsigned ver) const {
        asm volatile("st.relaxed.gpu.global.v8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};"
                     ::"l"(vrec + 2 * (size_t)b), "r"(__float_as_uint(v.v.x)), "r"(ver), "r"(__float_as_uint(v.v.y)), "r"(ver), "r"(__float_as_uint(v.w)), "r"(ver), "r"(0u), "r"(0u) : "memory");
    }
    __device__ __forceinline__ void put(int b, const Vel& v, unsigned ver, float im, float ii) const {  // (im, ii: the record's fourth word, unchanged)
        asm volatile("st.relaxed.gpu.global.v8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};"
                     ::"l"(vrec + 2 * (size_t)b), "r"(__float_as_uint(v.v.x)), "r"(ver), "r"(__float_as_uint(v.v.y)), "r"(ver), "r"(__float_as_uint(v.w)), "r"(ver),
                       "r"(__float_as_uint(im)), "r"(__float_as_uint(ii)) : "memory");
    }
};

// Colour the arbiters that are new this step (persisting ones kept their colour in k_build): lowest free colour,
// conflicts between new arbiters resolved by a hashed priority (64-bit atomicMin claims per body and colour, stamped
// with a round epoch so nothing is ever cleared).  A handful of new arbiters (the usual case in a settled scene) is
// coloured by CTA 0 alone with block barriers; a flood (scene start) by the grid.  On return every CTA sees the final
// colours and `used` masks.
// REGIONS: colours below GEN_BIT are "interior" colours, reserved for arbiters whose dynamic bodies all live in the
// region of the pair's owner body (k_solve_regions solves those in shared memory); every other arbiter, and an interior
// candidate that finds no free interior colour, takes a colour >= GEN_BIT ("generic": solved through the versioned
// records in L2).  The class of an arbiter IS its colour range, so it persists with the colour.
constexpr int PROMOTE_ROUND0 = 128;     // grid-wide rounds of the interior arbiters that overflowed into the generic colours
constexpr int CB = 4;                   // records a thread has in flight at once: a round is a chain of dependent L2 round trips
// `work`: something every CTA has to do that does not depend on the colours; it is called exactly once, where the CTA would wait.
template <bool REGIONS, typename Work>
__device__ __forceinline__ bool color_new_arbiters(const SolveArgs& a, Counters* cn, unsigned* bar, unsigned& bar_target, unsigned char* scratch, int scratch_bytes, Work work) {  // true: the grid has met at a barrier
    const int tid = blockIdx.x * blockDim.x + threadIdx.x, nth = gridDim.x * blockDim.x;
    const int T = blockDim.x, bid = blockIdx.x, lt = threadIdx.x;
    const int nunc = *(volatile int*)&cn->n_uncolored;
    if (nunc <= 0) { work(); return false; }
    const int gen_cta = REGIONS ? max(a.n_regions - 1, 0) : 0;  // the CTA that colours the cross-region arbiters (box piles: the last region is the lightest)
    int cq_ = 0;
    auto cstamp = [&]() { if (a.dbg && (bid == gen_cta || bid == a.probe_cta) && lt == 0 && cq_ < 8) { unsigned long long t; asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t)); a.dbg[(bid == gen_cta ? 48 : 56) + cq_++] = t; } };
    cstamp();
    int* lists[2] = {a.uncolored, a.uncolored2};
    const int4 no_rec = make_int4(-1, UNC_STATIC, UNC_STATIC, 0);  // (a lane without a record: loads body 0, does nothing)
    struct Masks { unsigned fi; unsigned long long fr; };
    // The loads of a round are unconditional (static bodies included, masked afterwards) and issued for CB records before any of
    // them is used: one L2 round trip per phase instead of one per record and branch.
    auto load_masks = [&](const int4 rec, bool local) -> Masks {  // free interior / generic colours at the dynamic bodies of the pair
        const int bx = rec.y & ~UNC_STATIC, by = rec.z & ~UNC_STATIC;
        const bool d1 = !(rec.y & UNC_STATIC), d2 = !(rec.z & UNC_STATIC);
        // (the CTA-local rounds promote an interior arbiter without a free interior colour instead of looking at the generic ones)
        const bool want_gen = !REGIONS || !local || rec.w <= 0, want_int = REGIONS && rec.w > 0;
        unsigned long long y1 = 0ull, y2 = 0ull;
        unsigned x1 = 0u, x2 = 0u;
        if (want_gen) { y1 = __ldcg(&a.used[bx]); y2 = __ldcg(&a.used[by]); }
        if (want_int) { x1 = __ldcg(&a.used_int[bx]); x2 = __ldcg(&a.used_int[by]); }
        Masks m;
        m.fi = ~((d1 ? x1 : 0u) | (d2 ? x2 : 0u));
        m.fr = ~((d1 ? y1 : 0ull) | (d2 ? y2 : 0ull));
        return m;
    };
    auto claim_val = [&](int e, int round) -> unsigned long long {
        return ((unsigned long long)(~(a.epoch_base + (unsigned)round)) << 32) | fmix32((unsigned)e * 2654435761u + (unsigned)round);
    };
    // one record of one round: the colour it goes for (-1: none free at all, -2: no free INTERIOR colour and `promote`)
    auto pick = [&](const int4 rec, const Masks m, bool promote) -> int {
        int c = -1;
        if (REGIONS && rec.w > 0) {  // interior: lowest free interior colour
            if (m.fi) c = __ffs((int)m.fi) - 1;
            else if (promote) return -2;
        }
        if (c < 0 && m.fr) c = (REGIONS ? GEN_BIT : 0) + __ffsll((long long)m.fr) - 1;  // lowest free (generic) colour
        if (c < 0) atomicExch(&cn->err, SYLT_ERR_CAPACITY);
        return c;
    };
    auto propose = [&](const int4 rec, const Masks m, int round) -> int {  // ... and its claims in the global cells
        const int c = pick(rec, m, false);
        if (c < 0) return -1;
        const unsigned long long val = claim_val(rec.x, round);
        if (!(rec.y & UNC_STATIC)) atomicMin(&a.claim[(size_t)(rec.y & ~UNC_STATIC) * NCX + c], val);
        if (!(rec.z & UNC_STATIC)) atomicMin(&a.claim[(size_t)(rec.z & ~UNC_STATIC) * NCX + c], val);
        return c;
    };
    // the colour of a record whose claims survived is final
    auto finalize = [&](const int4 rec, int c) {
        const int e = rec.x, bx = rec.y & ~UNC_STATIC, by = rec.z & ~UNC_STATIC;
        const bool d1 = !(rec.y & UNC_STATIC), d2 = !(rec.z & UNC_STATIC);
        a.cur.color[e] = c;
        if (REGIONS && c < GEN_BIT) {
            if (d1) atomicOr(&a.used_int[bx], 1u << c);
            if (d2) atomicOr(&a.used_int[by], 1u << c);
        } else {
            const int bit = REGIONS ? c - GEN_BIT : c;
            if (d1) atomicOr(&a.used[bx], 1ull << bit);
            if (d2) atomicOr(&a.used[by], 1ull << bit);
            if (REGIONS) {  // (as k_build does for the persisting ones; rec.w = +-(owner region + 1))
                const int2 entry = make_int2(e, ((rec.w < 0 ? -rec.w : rec.w) - 1) | (c << 16));
                if (d1) a.gadj[(size_t)bx * NC + bit] = entry;
                if (d2) a.gadj[(size_t)by * NC + bit] = entry;
            }
        }
        if (!REGIONS) atomicAdd(&cn->color_count[c], 1);  // (k_solve's colour-major order)
    };
    auto settle = [&](const int4 rec, int round, int c, unsigned long long cl1, unsigned long long cl2) -> bool {  // (cl1, cl2: its claim cells)
        const unsigned long long val = claim_val(rec.x, round);
        if ((!(rec.y & UNC_STATIC) && cl1 != val) || (!(rec.z & UNC_STATIC) && cl2 != val)) return false;
        finalize(rec, c);
        return true;
    };
    // Grid-wide rounds.  Round r reads the list L[r & 1] (n entries; `identity`: the first round takes every record); arbiters that
    // lose a claim append themselves to the other list, whose length is counted in cn->remaining[r].  The grid runs the rounds
    // while many arbiters are left, then CTA 0 finishes the stragglers alone with block barriers while the others wait at one
    // grid barrier.
    auto grid_rounds = [&](int n, int round0, bool identity) {
        bool solo = n <= T;  // block-uniform and identical in every CTA (one entry per thread of CTA 0)
        int round = round0;
        for (; round < round0 + PROMOTE_ROUND0; round++) {
            if (solo && bid != 0) break;
            const int ct = solo ? lt : tid, cs = solo ? T : nth;
            const int* cur_list = lists[round & 1];
            int* next_list = lists[(round + 1) & 1];
            const bool ident = identity && round == round0;
            for (int k0 = ct; k0 < n; k0 += CB * cs) {
                int r[CB];
                int4 rec[CB];
                Masks m[CB];
#pragma unroll
                for (int u = 0; u < CB; u++) r[u] = k0 + u * cs < n ? (ident ? k0 + u * cs : __ldcg(&cur_list[k0 + u * cs])) : -1;
#pragma unroll
                for (int u = 0; u < CB; u++) rec[u] = r[u] >= 0 ? __ldcg(&a.unc_rec[r[u]]) : no_rec;
#pragma unroll
                for (int u = 0; u < CB; u++) m[u] = load_masks(rec[u], false);
#pragma unroll
                for (int u = 0; u < CB; u++)
                    if (r[u] >= 0) a.prop[r[u]] = propose(rec[u], m[u], round);
            }
            if (solo) { __threadfence(); __syncthreads(); } else grid_barrier(bar, bar_target);
            for (int k0 = ct; k0 < n; k0 += CB * cs) {
                int r[CB], c[CB];
                int4 rec[CB];
                unsigned long long cl1[CB], cl2[CB];
#pragma unroll
                for (int u = 0; u < CB; u++) r[u] = k0 + u * cs < n ? (ident ? k0 + u * cs : __ldcg(&cur_list[k0 + u * cs])) : -1;
#pragma unroll
                for (int u = 0; u < CB; u++) { rec[u] = r[u] >= 0 ? __ldcg(&a.unc_rec[r[u]]) : no_rec; c[u] = r[u] >= 0 ? __ldcg(&a.prop[r[u]]) : -1; }
#pragma unroll
                for (int u = 0; u < CB; u++) {
                    const int cc = max(c[u], 0);
                    cl1[u] = __ldcg(&a.claim[(size_t)(rec[u].y & ~UNC_STATIC) * NCX + cc]);
                    cl2[u] = __ldcg(&a.claim[(size_t)(rec[u].z & ~UNC_STATIC) * NCX + cc]);
                }
#pragma unroll
                for (int u = 0; u < CB; u++)
                    if (c[u] >= 0 && !settle(rec[u], round, c[u], cl1[u], cl2[u])) next_list[atomicAdd(&cn->remaining[round], 1)] = r[u];
            }
            if (solo) { __threadfence(); __syncthreads(); } else grid_barrier(bar, bar_target);
            n = *(volatile int*)&cn->remaining[round];
            if (n == 0) break;
            if (n <= T) solo = true;  // the stragglers are CTA 0's job; the others go and wait
        }
        if (bid == 0 && lt == 0) atomicMax(&cn->rounds, round - round0 + 1);
        if (solo) grid_barrier(bar, bar_target);  // everybody waits for CTA 0's colours
    };
    // Region solver, the usual step (a few thousand new arbiters): an interior arbiter only ever meets interior arbiters of its
    // own region at its dynamic bodies, so every CTA colours the new interior arbiters of its region by itself — the same
    // propose / claim rounds, but with block barriers instead of grid barriers, on the records k_build dropped into the region's bin
    // — while CTA 0 also takes the (few) cross-region ones along.  One grid barrier at the end.  An interior arbiter without a free interior colour is "promoted": it takes a generic
    // colour in grid-wide rounds behind that barrier (a body touching more than 32 others inside its region; rare).
    // The claims of these rounds are cells of a hash table in shared memory, keyed by (body, colour): two keys that share a cell
    // only cost the loser another round.  A claim value (round | priority of the arbiter) is unique.
    if (REGIONS && a.color_local && *(volatile int*)&cn->bin_overflow == 0) {
        __shared__ int s_nn;
        const int n_int = bid < a.n_regions ? min(__ldcg(&cn->region_new[bid]), a.bin_cap) : 0;
        const int n_gen = bid == gen_cta ? min(__ldcg(&cn->region_new[a.n_regions]), a.bin_cap) : 0;
        const int n0 = n_int + n_gen;
        int HT = 16384;                        // claim cells: as many as fit next to the lists
        while (HT > 1024 && (size_t)HT * 8 + (size_t)n0 * 28 + 64 > (size_t)scratch_bytes) HT >>= 1;
        unsigned long long* ht = (unsigned long long*)scratch;
        int4* l_rec = (int4*)(ht + HT);        // my records
        int* l_cur = (int*)(l_rec + n0);       // the records still uncoloured (indices into l_rec), ping-pong
        int* l_next = l_cur + n0;
        int* l_prop = l_next + n0;
        int n = n0, round = 0;
        if (n0 > 0) {
            for (int k = lt; k < HT; k += T) ht[k] = ~0ull;
            for (int k = lt; k < n0; k += T) {
                l_rec[k] = __ldcg(&a.unc_bin[(size_t)(k < n_int ? bid : a.n_regions) * a.bin_cap + (k < n_int ? k : k - n_int)]);
                l_cur[k] = k;
            }
            __syncthreads();
        }
        cstamp();
        auto cell = [&](int body, int c) -> unsigned long long* { return &ht[hash_u32((unsigned)body * 128u + (unsigned)c) & (unsigned)(HT - 1)]; };
        // (later rounds claim with smaller values, so the table is never cleared; fmix32 is a bijection: the value is unique per
        // arbiter and does not depend on the order in which k_build filled the bin — the colours are reproducible run to run)
        auto local_val = [&](int e, int rnd) -> unsigned long long {
            return ((unsigned long long)(unsigned)(~rnd) << 32) | fmix32((unsigned)e * 2654435761u + (unsigned)rnd);
        };
        for (; n > 0 && round < 60; round++) {
            if (lt == 0) s_nn = 0;
            for (int k0 = lt; k0 < n; k0 += CB * T) {
                int q[CB];
                int4 rec[CB];
                Masks m[CB];
#pragma unroll
                for (int u = 0; u < CB; u++) { q[u] = k0 + u * T < n ? l_cur[k0 + u * T] : -1; rec[u] = q[u] >= 0 ? l_rec[q[u]] : no_rec; }
#pragma unroll
                for (int u = 0; u < CB; u++) m[u] = load_masks(rec[u], true);
#pragma unroll
                for (int u = 0; u < CB; u++)
                    if (q[u] >= 0) {
                        const int c = pick(rec[u], m[u], true);
                        l_prop[q[u]] = c;
                        if (c >= 0) {
                            const unsigned long long val = local_val(rec[u].x, round);
                            if (!(rec[u].y & UNC_STATIC)) atomicMin(cell(rec[u].y, c), val);
                            if (!(rec[u].z & UNC_STATIC)) atomicMin(cell(rec[u].z, c), val);
                        } else if (c == -2) {
                            lists[PROMOTE_ROUND0 & 1][atomicAdd(&cn->n_promoted, 1)] =
                                __ldcg(&a.unc_bin_idx[(size_t)(q[u] < n_int ? bid : a.n_regions) * a.bin_cap + (q[u] < n_int ? q[u] : q[u] - n_int)]);
                        }
                    }
            }
            __syncthreads();
            for (int k = lt; k < n; k += T) {
                const int q = l_cur[k], c = l_prop[q];
                if (c < 0) continue;
                const int4 rec = l_rec[q];
                const unsigned long long val = local_val(rec.x, round);
                const bool win = ((rec.y & UNC_STATIC) || *cell(rec.y, c) == val) && ((rec.z & UNC_STATIC) || *cell(rec.z, c) == val);
                if (win) finalize(rec, c);
                else l_next[atomicAdd(&s_nn, 1)] = q;
            }
            __threadfence();  // (the winners' bits in used / used_int: read by the losers in the next round)
            __syncthreads();
            n = s_nn;
            int* t = l_cur; l_cur = l_next; l_next = t;
            __syncthreads();
        }
        for (int k = lt; k < n; k += T) {  // (never in practice: 60 rounds were not enough — the rest joins the grid-wide rounds)
            const int q = l_cur[k];
            lists[PROMOTE_ROUND0 & 1][atomicAdd(&cn->n_promoted, 1)] = __ldcg(&a.unc_bin_idx[(size_t)(q < n_int ? bid : a.n_regions) * a.bin_cap + (q < n_int ? q : q - n_int)]);
        }
        cstamp();
        if (lt == 0 && round > 0) atomicMax(&cn->rounds, round);
        __syncthreads();  // (the scratch is dead: `work` may use the shared memory)
        if (bid != gen_cta) work();  // everybody else would only wait for the cross-region colours
        grid_barrier(bar, bar_target);
        if (bid == gen_cta) work();
        cstamp();
        const int np = *(volatile int*)&cn->n_promoted;
        if (np > 0) grid_rounds(np, PROMOTE_ROUND0, false);
        return true;
    }
    grid_rounds(nunc, 0, true);
    work();
    return true;
}

__global__ void __launch_bounds__(SOLVE_THREADS, 1) k_solve(SolveArgs a) {
    extern __shared__ __align__(16) unsigned char sm_raw[];
    Counters* cn = a.cn;
    const int tid = blockIdx.x * blockDim.x + threadIdx.x, nth = gridDim.x * blockDim.x;
    const int T = blockDim.x, G = gridDim.x, bid = blockIdx.x, lt = threadIdx.x;
    const int A = min(*(volatile int*)&cn->n_arb, a.cap_arb);
    unsigned bar_target = 0;
    unsigned* bar = &cn->bar;
    auto stamp = [&](int k) {
        if (a.dbg && tid == 0) { unsigned long long t; asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t)); a.dbg[k] = t; }
    };
    stamp(0);
    if (step_failed(cn)) return;  // grid-uniform: err only changes inside this kernel from here on
    fill_arbiter_rows(a.B, min(*(volatile int*)&cn->n_pairs, a.cap_pairs), a.cand_row_start, a.pscan, a.cur.row_start, tid, nth);  // (for the next step's k_build)
    // The first joint in Vec order whose K is singular is where World::step returns Err (world.rs:127-129, quirk Q-15); K depends
    // on the poses only, so it is known before anything is pre-stepped.  (Visible to everybody after the barrier below.)
    for (int j = tid; j < a.J; j += nth) {
        const int2 b = a.jbodies[j];
        const float4 la = a.jla[j], m1 = a.mp[b.x], m2 = a.mp[b.y];
        const float2 c1 = a.rotc[b.x], c2 = a.rotc[b.y];
        if (joint_singular(mk(la.x, la.y), mk(la.z, la.w), a.jparam[j].x, c1.x, c1.y, c2.x, c2.y, m1.x, m1.y, m2.x, m2.y)) atomicMax(&cn->jfail1, a.J - j);
    }
    __shared__ int s_start[NC + 1];   // global colour-major start of colour c
    __shared__ int s_lo[NC + 1];      // first slot of colour c this CTA owns (relative to s_start[c])
    __shared__ int s_lbase[NC + 1];   // local index of that slot
    __shared__ int s_ncolors;
    __shared__ int s_scan[34];
    // shared-memory carve-up: records (2 x float4 per slot), slot contact counts (scan scratch), contacts (3 x float4)
    const int SLOTCAP = a.smem_slots, CONCAP = a.smem_contacts;
    float4* s_rec = (float4*)sm_raw;
    int* s_n = (int*)(s_rec + 2 * (size_t)SLOTCAP);
    float4* s_ca = (float4*)(s_n + ((SLOTCAP + 3) & ~3));
    float4* s_cb = s_ca + CONCAP;
    float4* s_cc = s_cb + CONCAP;

    color_new_arbiters<false>(a, cn, bar, bar_target, nullptr, 0, [] {});

    stamp(1);
    // ---- 2. colour-major order (within a colour the order is irrelevant: disjoint dynamic bodies) and this CTA's share
    if (lt == 0) {
        int run = 0, nc = 0, lrun = 0;
        for (int c = 0; c < NC; c++) {
            s_start[c] = run;
            int k = *(volatile int*)&cn->color_count[c];
            if (k > 0) nc = c + 1;
            int chunk = (k + G - 1) / G;
            int lo = min(k, bid * chunk), hi = min(k, lo + chunk);
            s_lo[c] = lo;
            s_lbase[c] = lrun;
            lrun += hi - lo;
            run += k;
        }
        s_start[NC] = run;
        s_lo[NC] = 0;
        s_lbase[NC] = lrun;
        s_ncolors = nc;
        if (bid == 0) {
            for (int c = 0; c <= NC; c++) cn->color_start[c] = s_start[c];
            cn->n_colors = nc;
        }
    }
    __syncthreads();
    for (int e0 = tid - (lt & 31); e0 < A; e0 += nth) {  // warp-uniform trip count
        const int lane = lt & 31, e = e0 + lane;
        int c = e < A ? __ldcg(&a.cur.color[e]) : -1;
        unsigned act = __ballot_sync(0xffffffffu, c >= 0);
        if (c >= 0) {  // warp-aggregated slot allocation per colour
            unsigned peers = __match_any_sync(act, c);
            int leader = __ffs(peers) - 1;
            int base = 0;
            if (lane == leader) base = atomicAdd(&cn->color_fill[c], __popc(peers));
            base = __shfl_sync(peers, base, leader);
            a.order[s_start[c] + base + __popc(peers & ((1u << lane) - 1))] = e;
        }
    }
    grid_barrier(bar, bar_target);
    if (!a.run_solver) {
        if (tid == 0 && !step_failed(cn)) step_completed(cn);
        return;
    }
    if (*(volatile int*)&cn->err == SYLT_ERR_CAPACITY) return;  // e.g. an arbiter without a free colour: no solve, error reported
    const int jfail1 = *(volatile int*)&cn->jfail1;
    const int jfail = jfail1 > 0 ? a.J - jfail1 : 0x7fffffff;
    const int ncolors = s_ncolors;
    const int L = s_lbase[NC];  // slots this CTA owns
    const bool acc = a.accumulate != 0;
    const float kbias = a.poscorr ? 0.2f : 0.0f;

    stamp(2);
    // ---- 3. residency: load the records of the owned slots; contacts get shared-memory room by an exclusive scan.
    //         Each record also gets, per dynamic body, (D, r): D = constraints touching the body, r = how many of them
    //         precede this arbiter in the solve order (colours are distinct around a body, so r = popc(used & lower bits)).
    const int Ls = min(L, SLOTCAP);
    for (int c = 0; c < ncolors; c++) {
        const int cnt = s_lbase[c + 1] - s_lbase[c];
        for (int i = lt; i < cnt; i += T) {
            int ls = s_lbase[c] + i;
            if (ls >= SLOTCAP) continue;
            int e = __ldcg(&a.order[s_start[c] + s_lo[c] + i]);
            s_n[ls] = a.cur.meta[e].x;
        }
    }
    __syncthreads();
    {   // in-place exclusive scan of s_n[0..Ls)
        const int chunk = (Ls + T - 1) / T, q0 = min(lt * chunk, Ls), q1 = min(q0 + chunk, Ls);
        int sum = 0;
        for (int k = q0; k < q1; k++) sum += s_n[k];
        int tot;
        int ex = block_excl_scan<int>(sum, &tot, s_scan);
        for (int k = q0; k < q1; k++) { int t = s_n[k]; s_n[k] = ex; ex += t; }
    }
    __syncthreads();
    auto order_pack = [&](int b1, int b2, int c) -> unsigned {  // D1 | r1 << 8 | D2 << 16 | r2 << 24
        unsigned long long u1 = __ldcg(&a.used[b1]), u2 = __ldcg(&a.used[b2]);
        unsigned long long lower = (1ull << c) - 1ull;
        unsigned d1 = (unsigned)(__popcll(u1) + a.jdeg[b1]), d2 = (unsigned)(__popcll(u2) + a.jdeg[b2]);
        return (d1 & 0xffu) | ((unsigned)__popcll(u1 & lower) << 8) | ((d2 & 0xffu) << 16) | ((unsigned)__popcll(u2 & lower) << 24);
    };
    for (int c = 0; c < ncolors; c++) {
        const int cnt = s_lbase[c + 1] - s_lbase[c];
        for (int i = lt; i < cnt; i += T) {
            int ls = s_lbase[c] + i;
            if (ls >= SLOTCAP) continue;
            int e = __ldcg(&a.order[s_start[c] + s_lo[c] + i]);
            int2 b = a.cur.bodies[e];
            int2 mt = a.cur.meta[e];
            int lc = s_n[ls];
            const bool res = lc + mt.x <= CONCAP;
            int r1 = b.x | ((a.bflags[b.x] & FLAG_STATIC) ? VSTATIC : 0), r2 = b.y | ((a.bflags[b.y] & FLAG_STATIC) ? VSTATIC : 0);
            s_rec[2 * ls] = make_float4(__int_as_float(r1), __int_as_float(r2), __int_as_float(res ? (lc << 8) | mt.x : -1), a.cur.friction[e]);
            s_rec[2 * ls + 1] = make_float4(__uint_as_float(order_pack(b.x, b.y, c)), 0.0f, 0.0f, 0.0f);
            if (res)
                for (int q = 0; q < mt.x; q++) {
                    s_ca[lc + q] = a.cur.ca[mt.y + q];
                    s_cc[lc + q] = a.cur.cc[mt.y + q];
                }
        }
    }
    __syncthreads();

    auto view = [&](int c, int i, SlotView& v) {
        int ls = s_lbase[c] + i;
        if (ls < SLOTCAP) {
            float4 r0 = s_rec[2 * ls];
            int pk = __float_as_int(r0.z);
            if (pk >= 0) {
                v.r1 = __float_as_int(r0.x); v.r2 = __float_as_int(r0.y); v.n = pk & 0xff; v.fr = r0.w;
                v.ord = __float_as_uint(s_rec[2 * ls + 1].x);
                int lc = pk >> 8;
                v.ca = s_ca + lc; v.cb = s_cb + lc; v.cc = s_cc + lc;
                return;
            }
        }
        int e = __ldcg(&a.order[s_start[c] + s_lo[c] + i]);
        int2 b = a.cur.bodies[e];
        int2 mt = a.cur.meta[e];
        v.r1 = b.x | ((a.bflags[b.x] & FLAG_STATIC) ? VSTATIC : 0); v.r2 = b.y | ((a.bflags[b.y] & FLAG_STATIC) ? VSTATIC : 0);
        v.n = mt.x; v.fr = a.cur.friction[e];
        v.ord = order_pack(b.x, b.y, c);
        v.ca = a.cur.ca + mt.y; v.cb = a.cur.cb + mt.y; v.cc = a.cur.cc + mt.y;
    };

    // Dataflow Gauss-Seidel: no barrier between colours.  A constraint of pass t (0 = pre_step, 1..I = iterations) waits
    // until both of its dynamic bodies carry version t*D + r, i.e. until every constraint that precedes it in the
    // sequential (pass, colour) order has written them, then writes them back with version + 1.  Threads walk their
    // slots in (pass, colour) order and all CTAs are co-resident, so the globally smallest unfinished constraint can
    // always run: no deadlock.  The result is exactly the sequential sweep in (colour, then joint colour) order.
    BodyIO io{a.vrec, a.sleep_ns, &cn->err};
    // Polling throttle, CTA-local: s_done counts the slots this CTA has finished (all passes).  Before a warp starts its
    // slots of (pass, colour) it waits (on shared memory only) until the CTA has finished as many slots as precede that
    // phase locally, i.e. the CTA walks the phases roughly in step without bar.sync, and different CTAs drift freely.
    // Cross-CTA ordering is carried by the versions alone; only threads of a CTA's current phase ever poll L2.
    // (With all ~75k threads polling L2 at once the poll traffic itself was the bottleneck; a global progress counter
    // serialises on one L2 address.)  Exactness never depends on this counter, only on the versions.
    __shared__ unsigned s_done;
    if (lt == 0) s_done = 0u;
    __syncthreads();
    const int lane = lt & 31;
    // Slot i of a colour belongs to thread i (the write-back below relies on it: no barrier between the last pass and it).
    // Dealing the slots round-robin to the warps instead (a few lanes in each of the 16 warps rather than a few full warps)
    // was measured: iterations 252 -> 289 us at 100k bodies — more warps polling L2 costs more than shorter waits save.
    const int slot_of_thread = lt;
    auto wait_progress = [&](unsigned need) {
        if (lane == 0) {
            unsigned spins = 0;
            while (*(volatile unsigned*)&s_done < need)
                if ((++spins & 0xffffu) == 0u && (*(volatile int*)&cn->err == SYLT_ERR_INTERNAL || spins > (1u << 28))) break;
        }
        __syncwarp();
    };
    auto slack_of = [&](int c) -> int {  // slots of the previous local phase that may still be unfinished when a warp moves on
        int prev = c > 0 ? s_lbase[c] - s_lbase[c - 1] : s_lbase[ncolors] - s_lbase[ncolors - 1];
        return prev * a.slack_pct / 100;
    };
    auto add_progress = [&](bool did) {
        unsigned m = __ballot_sync(0xffffffffu, did);
        if (lane == 0 && m) atomicAdd(&s_done, (unsigned)__popc(m));
    };
    auto prestep_slot = [&](int c, int i) {
        SlotView v;
        view(c, i, v);
        int e = __ldcg(&a.order[s_start[c] + s_lo[c] + i]);
        int coff = a.cur.meta[e].y;
        const int b1 = v.r1 & ~VSTATIC, b2 = v.r2 & ~VSTATIC;
        float4 p1 = a.pose[b1], p2 = a.pose[b2];
        const unsigned e1 = (v.ord >> 8) & 0xffu, e2 = v.ord >> 24;  // pass 0: expected version = rank
        Vel v1, v2;
        float im1, ii1, im2, ii2;
        io.wait2(v.r1, e1, v1, im1, ii1, v.r2, e2, v2, im2, ii2);
        for (int q = 0; q < v.n; q++) {
            float4 ca = v.ca[q], cc = v.cc[q], cd = a.cur.cd[coff + q];
            ContactConst k0;
            contact_prestep(mk(cd.x, cd.y), mk(ca.x, ca.y), cc.w, mk(p1.x, p1.y), mk(p2.x, p2.y), im1, ii1, im2, ii2,
                            a.inv_dt, kbias, acc, cc.y, cc.z, v1, v2, k0);
            v.ca[q] = make_float4(ca.x, ca.y, k0.r1.x, k0.r1.y);
            v.cb[q] = make_float4(k0.r2.x, k0.r2.y, k0.mass_n, k0.mass_t);
            v.cc[q] = make_float4(k0.bias, cc.y, cc.z, cc.w);
        }
        if (!(v.r1 & VSTATIC)) io.put(b1, v1, e1 + 1, im1, ii1);
        if (!(v.r2 & VSTATIC)) io.put(b2, v2, e2 + 1, im2, ii2);
    };
    auto apply_slot = [&](int c, int i, unsigned pass) {
        SlotView v;
        view(c, i, v);
        const int b1 = v.r1 & ~VSTATIC, b2 = v.r2 & ~VSTATIC;
        const unsigned e1 = pass * (v.ord & 0xffu) + ((v.ord >> 8) & 0xffu), e2 = pass * ((v.ord >> 16) & 0xffu) + (v.ord >> 24);
        Vel v1, v2;
        float im1, ii1, im2, ii2;
        io.wait2(v.r1, e1, v1, im1, ii1, v.r2, e2, v2, im2, ii2);
        for (int q = 0; q < v.n; q++) {
            float4 ca = v.ca[q], cb = v.cb[q], cc = v.cc[q];
            ContactConst k0;
            k0.n = mk(ca.x, ca.y); k0.r1 = mk(ca.z, ca.w); k0.r2 = mk(cb.x, cb.y);
            k0.mass_n = cb.z; k0.mass_t = cb.w; k0.bias = cc.x;
            float pn = cc.y, pt = cc.z;
            contact_apply(k0, v.fr, im1, ii1, im2, ii2, acc, pn, pt, v1, v2);
            if (acc) v.cc[q] = make_float4(cc.x, pn, pt, cc.w);
        }
        if (!(v.r1 & VSTATIC)) io.put(b1, v1, e1 + 1, im1, ii1);
        if (!(v.r2 & VSTATIC)) io.put(b2, v2, e2 + 1, im2, ii2);
    };
    auto joint_expect = [&](int j, int2 b, unsigned pass, unsigned& e1, unsigned& e2) {
        int2 jr = a.jrank[j];
        unsigned da1 = (unsigned)__popcll(__ldcg(&a.used[b.x])), da2 = (unsigned)__popcll(__ldcg(&a.used[b.y]));
        e1 = pass * (da1 + (unsigned)a.jdeg[b.x]) + da1 + (unsigned)jr.x;
        e2 = pass * (da2 + (unsigned)a.jdeg[b.y]) + da2 + (unsigned)jr.y;
    };

    stamp(3);
    // ---- 4. pass 0: Arbiter::pre_step (arbiter.rs:182-228) then Joint::pre_step (joint.rs:57-115)
    for (int c = 0; c < ncolors; c++) {
        const int cnt = s_lbase[c + 1] - s_lbase[c];
        for (int r0 = 0; r0 < cnt; r0 += T) {  // warp-uniform trip count
            const int i = r0 + slot_of_thread;
            const bool did = i < cnt;
            wait_progress((unsigned)max(0, s_lbase[c] - slack_of(c)));
            if (did) prestep_slot(c, i);
            add_progress(did);
        }
    }
    for (int c = 0; c < a.n_joint_colors; c++) {
        for (int k0 = a.jcolor_start[c] + tid - lane; k0 < a.jcolor_start[c + 1]; k0 += nth) {
            const int k = k0 + lane;
            const bool did = k < a.jcolor_start[c + 1];
            if (did) {
            int j = a.jorder[k];
            int2 b = a.jbodies[j];
            float4 p1 = a.pose[b.x], p2 = a.pose[b.y];
            float2 r1c = a.rotc[b.x], r2c = a.rotc[b.y];
            const int r1 = b.x | ((a.bflags[b.x] & FLAG_STATIC) ? VSTATIC : 0), r2 = b.y | ((a.bflags[b.y] & FLAG_STATIC) ? VSTATIC : 0);
            unsigned e1, e2;
            joint_expect(j, b, 0u, e1, e2);
            Vel v1, v2;
            float im1, ii1, im2, ii2;
            io.wait2(r1, e1, v1, im1, ii1, r2, e2, v2, im2, ii2);
            Vel w1 = v1, w2 = v2;
            if (j <= jfail) {  // the reference's loop returns at joint `jfail`: later joints (Vec order) are never touched
                float4 la = a.jla[j];
                float2 pr = a.jparam[j];
                float2 pa = a.jp[j];
                V2 pacc = mk(pa.x, pa.y);
                JointConst jc;
                M2 bad;
                bool ok = joint_prestep(mk(la.x, la.y), mk(la.z, la.w), pr.x, pr.y, mk(p1.x, p1.y), r1c.x, r1c.y, mk(p2.x, p2.y), r2c.x, r2c.y,
                                        im1, ii1, im2, ii2, a.inv_dt, a.poscorr != 0, a.warm != 0, pacc, w1, w2, jc, &bad, a.fix_inverse != 0);
                a.jr[j] = make_float4(jc.r1.x, jc.r1.y, jc.r2.x, jc.r2.y);  // self.r1 / self.r2 are stored before the `?` (joint.rs:67-68)
                if (!ok) {  // quirk Q-15: NoInverse aborts the step before the iterations (only joint `jfail` gets here)
                    if (atomicCAS(&cn->err, 0, SYLT_ERR_NO_INVERSE) == 0) {
                        cn->err_joint = j;
                        cn->badk[0] = bad.c1.x; cn->badk[1] = bad.c2.x; cn->badk[2] = bad.c1.y; cn->badk[3] = bad.c2.y;
                    }
                    w1 = v1; w2 = v2;  // velocities untouched, versions still advance so that nobody waits forever
                } else {
                    a.jm[j] = make_float4(jc.m.c1.x, jc.m.c2.x, jc.m.c1.y, jc.m.c2.y);
                    a.jbias[j] = make_float2(jc.bias.x, jc.bias.y);
                    a.jp[j] = make_float2(pacc.x, pacc.y);
                }
            }
            if (!(r1 & VSTATIC)) io.put(b.x, w1, e1 + 1, im1, ii1);
            if (!(r2 & VSTATIC)) io.put(b.y, w2, e2 + 1, im2, ii2);
            }
        }
    }
    grid_barrier(bar, bar_target);
    const bool aborted = *(volatile int*)&cn->err == SYLT_ERR_NO_INVERSE;
    stamp(4);

    // ---- 5. passes 1..I: iterations x (arbiter colours, joint colours)   (world.rs:132-140)
    for (int it = 1; it <= (aborted ? 0 : a.iterations); it++) {
        for (int c = 0; c < ncolors; c++) {
            const int cnt = s_lbase[c + 1] - s_lbase[c];
            for (int r0 = 0; r0 < cnt; r0 += T) {
                const int i = r0 + slot_of_thread;
                const bool did = i < cnt;
                wait_progress((unsigned)max(0, it * L + s_lbase[c] - slack_of(c)));
                if (did) apply_slot(c, i, (unsigned)it);
                add_progress(did);
            }
        }
        for (int c = 0; c < a.n_joint_colors; c++) {
            for (int k0 = a.jcolor_start[c] + tid - lane; k0 < a.jcolor_start[c + 1]; k0 += nth) {
                const int k = k0 + lane;
                const bool did = k < a.jcolor_start[c + 1];
                if (did) {
                int j = a.jorder[k];
                int2 b = a.jbodies[j];
                const int r1 = b.x | ((a.bflags[b.x] & FLAG_STATIC) ? VSTATIC : 0), r2 = b.y | ((a.bflags[b.y] & FLAG_STATIC) ? VSTATIC : 0);
                unsigned e1, e2;
                joint_expect(j, b, (unsigned)it, e1, e2);
                Vel v1, v2;
                float im1, ii1, im2, ii2;
                io.wait2(r1, e1, v1, im1, ii1, r2, e2, v2, im2, ii2);
                float4 r = a.jr[j], mm = a.jm[j];
                float2 bs = a.jbias[j], pa = a.jp[j];
                JointConst jc;
                jc.r1 = mk(r.x, r.y); jc.r2 = mk(r.z, r.w); jc.bias = mk(bs.x, bs.y);
                jc.m = mk(mk(mm.x, mm.z), mk(mm.y, mm.w));
                jc.softness = a.jparam[j].x;
                V2 pacc = mk(pa.x, pa.y);
                joint_apply(jc, im1, ii1, im2, ii2, pacc, v1, v2);
                a.jp[j] = make_float2(pacc.x, pacc.y);
                if (!(r1 & VSTATIC)) io.put(b.x, v1, e1 + 1, im1, ii1);
                if (!(r2 & VSTATIC)) io.put(b.y, v2, e2 + 1, im2, ii2);
                }
            }
        }
    }

    stamp(5);
    // ---- 6. write the shared-memory resident contacts back (warm start of the next step, read-back)
    for (int c = 0; c < ncolors; c++) {
        const int cnt = s_lbase[c + 1] - s_lbase[c];
        for (int i = lt; i < cnt; i += T) {
            int ls = s_lbase[c] + i;
            if (ls >= SLOTCAP) continue;
            int pk = __float_as_int(s_rec[2 * ls].z);
            if (pk < 0) continue;
            int e = __ldcg(&a.order[s_start[c] + s_lo[c] + i]);
            int coff = a.cur.meta[e].y, lc = pk >> 8, n = pk & 0xff;
            for (int q = 0; q < n; q++) {
                a.cur.ca[coff + q] = s_ca[lc + q];
                a.cur.cb[coff + q] = s_cb[lc + q];
                a.cur.cc[coff + q] = s_cc[lc + q];
            }
        }
    }
    grid_barrier(bar, bar_target);

    stamp(6);
    // ---- 7. velocities out of the versioned records; integrate positions, clear forces (world.rs:143-150; static
    //         bodies too, quirk Q-16).  After a NoInverse the step stops before the integration (quirk Q-15).
    for (int i = tid; i < a.B; i += nth) {
        Vel v;
        float im, ii;
        io.get(i, v, im, ii);
        a.vel[i] = make_float4(v.v.x, v.v.y, v.w, 0.0f);
        if (aborted) continue;
        float4 p = a.pose[i];
        V2 np = mk(p.x, p.y) + v.v * a.dt;
        p.x = np.x; p.y = np.y;
        p.z += v.w * a.dt;
        a.pose[i] = p;
        a.force[i] = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
    }
    if (tid == 0 && !aborted && *(volatile int*)&cn->err == 0) step_completed(cn);
    stamp(7);
}

// ------------------------------------------------------------------ region solver
// The same sequential sweep as k_solve — passes x colours ascending — with a different execution plan.
// Bodies are partitioned into compact spatial regions (k_recut, every `regions_period` steps); region r belongs to CTA r for the
// whole solve and its body velocities live in that CTA's shared memory.  Arbiters come in two classes (the colour
// range IS the class, see color_new_arbiters): interior (colour < GEN_BIT: all dynamic bodies in the owner's region)
// and generic (cross-region pairs).  Interior arbiters of different regions touch disjoint dynamic bodies, so the
// regions sweep their interior colours side by side with one __syncthreads() per colour.
// The generic colours are solved by REDUNDANT computation instead of communication: what a region needs from the
// generic phase of a pass is the final state of its own boundary bodies, and that depends only on the generic arbiters
// reachable from its own through "a generic arbiter of a lower colour shares a dynamic body" — a cone that is closed
// after at most (number of generic colours) steps and stays near the region's border.  Every CTA collects that cone
// (a breadth-first walk over a per-body table of generic arbiters), imports the "ghost" bodies the cone touches once
// per pass — their state after the owner's interior sweep, exported through a per-pass plane of versioned records in
// L2 — and replays the cone in colour order in its own shared memory.  f32 arithmetic is deterministic, so every
// replica of a generic arbiter computes the same bits; a pass costs ONE L2 hand-off instead of one per generic colour.
// Only the owner writes an arbiter's contacts back.  Worlds with joints use k_solve.
constexpr int GADJ = NC;              // per body: its generic arbiters, indexed by the colour bit (GEN_BIT + k -> entry k)
constexpr int GHOST_FLAG = 0x40000000;
}  // namespace
// table sizes of a region CTA that has an SM to itself (k_solve_regions<PER_SM> divides them by PER_SM)
constexpr int CONE_CAP_1 = 2048;      // ghost arbiters a CTA may replay
constexpr int HASH_CAP_1 = 4096;      // open-addressing tables of the cone walk (set-up scratch at the top of shared memory)
constexpr int GB_MAX_1 = 2048;        // ghost bodies of a CTA
constexpr int OWN_MAX_1 = 4096;       // arbiters a region may own
constexpr int UNC_BIN_CAP_1 = 2048;   // new arbiters of one step a region CTA colours by itself (more: grid-wide rounds for that step)
namespace {
constexpr int MAX_PLANES = 16;        // export planes: pass t uses plane t % MAX_PLANES; a grid barrier every MAX_PLANES passes keeps
                                      // every CTA within one lap of the others, so a plane is never overwritten before it has been read
constexpr int MAX_PASSES = 4096;      // passes (1 + iterations) the 12-bit pass field of an export tag can tell apart


// PER_SM region CTAs share an SM (each with 1 / PER_SM of the threads, the shared memory and the table sizes): the passes are a chain
// of short latency-bound colour hops, so two half-size regions per SM overlap each other's latencies.
template <int PER_SM>
__global__ void __launch_bounds__(SOLVE_THREADS / PER_SM, PER_SM) k_solve_regions(SolveArgs a) {
    constexpr int CONE_CAP = ::CONE_CAP_1 / PER_SM, HASH_CAP = ::HASH_CAP_1 / PER_SM, GB_MAX = ::GB_MAX_1 / PER_SM, OWN_MAX = ::OWN_MAX_1 / PER_SM;
    extern __shared__ __align__(16) unsigned char sm_raw[];
    Counters* cn = a.cn;
    const int tid = blockIdx.x * blockDim.x + threadIdx.x, nth = gridDim.x * blockDim.x;
    const int T = blockDim.x, bid = blockIdx.x, lt = threadIdx.x;
    const int A = min(*(volatile int*)&cn->n_arb, a.cap_arb);
    unsigned bar_target = 0;
    unsigned* bar = &cn->bar;
    auto stamp = [&](int k) {
        if (a.dbg && tid == 0) { unsigned long long t; asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t)); a.dbg[k] = t; }
    };
    stamp(0);
    __shared__ int s_cnt[NCX], s_fill[NCX], s_cstart[NCX + 1], s_clist[NCX], s_ccnt[NCX], s_cfill[NCX];  // slots / contacts by colour
    __shared__ int s_ncl, s_nint, s_nbnd, s_L, s_ngh, s_ngb, s_nown, s_bad;
    __shared__ int2 s_hop[NCX];
    __shared__ int s_wtot[3], s_wctot[3];
    __shared__ unsigned s_wbal[3];
    // Shared memory, carved as the sizes become known: boundary lists [nb], bodies (32 bytes each: own, then ghosts), ghost
    // list, slot records (12 bytes), contacts (36 bytes, all that is left).  During set-up the top of the area holds the
    // scratch of the cone walk: hash tables, the ghost ids, the (arbiter, colour) list of the own arbiters.
    const int R = bid;
    const int rb0 = R < a.n_regions ? a.reg_start[R] : 0;
    const int nb = R < a.n_regions ? a.reg_start[R + 1] - rb0 : 0;
    int* s_bnd = (int*)sm_raw;                            // own boundary bodies: local index
    int* s_bndg = s_bnd + nb;                             //                      global index
    float4* s_body = (float4*)(sm_raw + (((size_t)nb * 8 + 15) & ~(size_t)15));  // [2i] = (vx, vy, w, -), [2i+1] = (x, y, inv_mass, inv_moi)
    int* h_arb = (int*)(sm_raw + a.smem_bytes) - 3 * HASH_CAP;  // set of ghost arbiters (key = arbiter + 1)
    int* h_body = h_arb + HASH_CAP;                       // ghost body -> index (key = body + 1, value)
    int* h_bval = h_body + HASH_CAP;
    int* t_gb = h_arb - GB_MAX;                           // ghost bodies in discovery order
    int2* t_own = (int2*)t_gb - OWN_MAX;                  // (arbiter, colour) of the arbiters this region owns
    int2* ghosts = t_own - CONE_CAP;                      // ghost arbiters of this CTA (arbiter, colour), in discovery order
    const unsigned char* scratch_lo = (const unsigned char*)ghosts;

    int tq_ = 32;
    auto tstamp = [&]() { if (a.dbg && bid == a.probe_cta && lt == 0 && tq_ < 48) { unsigned long long t; asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t)); a.dbg[tq_++] = t; } };
    // a step that already failed in the broad phase (capacity) has holes in its arbiter arrays: nothing to solve, error reported
    if (step_failed(cn)) return;
    fill_arbiter_rows(a.B, min(*(volatile int*)&cn->n_pairs, a.cap_pairs), a.cand_row_start, a.pscan, a.cur.row_start, tid, nth);  // (for the next step's k_build and k_recut; this kernel takes its rows from the candidate rows)
    // (the per-body table of generic arbiters `gadj` — slot = colour bit, entry = (arbiter, owner region | colour << 16) — was written by
    // k_build and by the colouring; the rows and the colours of other CTAs' arbiters are visible behind the grid barrier)
    // The part of the set-up that needs no colours — the tables, the region's bodies, the list of its own arbiters — runs where the CTA
    // would otherwise wait for the cross-region colours.  (The rows of its own bodies come straight from the candidate rows.)
    auto arb_insert = [&](int f) -> bool {  // true: f is new
        unsigned h = hash_u32((unsigned)f) & (HASH_CAP - 1);
        for (int probe = 0; probe < HASH_CAP; probe++) {
            const int old = atomicCAS(&h_arb[h], 0, f + 1);
            if (old == 0) return true;
            if (old == f + 1) return false;
            h = (h + 1) & (HASH_CAP - 1);
        }
        s_bad = 1;
        return false;
    };
    auto add_ghost = [&](int f, int c) {
        if (!arb_insert(f)) return;
        const int g = atomicAdd(&s_ngh, 1);
        if (g < CONE_CAP) { ghosts[g] = make_int2(f, c); atomicAdd(&s_cnt[c], 1); }
        else s_bad = 1;
    };
    // ghost bodies: bodies of other regions that this CTA's slots touch (dynamic: re-imported every pass; static: once)
    auto body_insert = [&](int z, bool is_static) {
        unsigned h = hash_u32((unsigned)z ^ 0x9e3779b9u) & (HASH_CAP - 1);
        for (int probe = 0; probe < HASH_CAP; probe++) {
            const int old = atomicCAS(&h_body[h], 0, z + 1);
            if (old == 0) {
                const int g = atomicAdd(&s_ngb, 1);
                h_bval[h] = g;
                if (g < GB_MAX) t_gb[g] = z | (is_static ? GHOST_FLAG : 0); else s_bad = 1;
                return;
            }
            if (old == z + 1) return;
            h = (h + 1) & (HASH_CAP - 1);
        }
        s_bad = 1;
    };
    auto body_lookup = [&](int z) -> int {
        unsigned h = hash_u32((unsigned)z ^ 0x9e3779b9u) & (HASH_CAP - 1);
        for (int probe = 0; probe < HASH_CAP; probe++) {
            const int key = h_body[h];
            if (key == z + 1) return h_bval[h];
            if (key == 0) break;
            h = (h + 1) & (HASH_CAP - 1);
        }
        s_bad = 1;
        return 0;
    };
    // One arbiter joins this CTA's slots (own: e is a row of one of its bodies; ghost: a replica): its contacts are counted for the
    // layout, its bodies of other regions become ghost bodies, and — a generic arbiter of colour c — everything it depends on
    // inside one generic phase, the lower-coloured generic arbiters at its dynamic bodies, joins too.
    auto adopt = [&](int c, const int4 ra, int n) {
        atomicAdd(&s_ccnt[c], n);
        const int z1 = ra.x & SR_BODY, z2 = ra.y & SR_BODY;
        const bool st1 = (ra.x & SR_STATIC) != 0, st2 = (ra.y & SR_STATIC) != 0;
        if ((ra.z >> 16) != R) body_insert(z1, st1);
        if ((ra.w >> 16) != R) body_insert(z2, st2);
        if (c < GEN_BIT) return;
        unsigned long long u1 = 0ull, u2 = 0ull;
        if (!st1) u1 = __ldcg(&a.used[z1]);
        if (!st2) u2 = __ldcg(&a.used[z2]);
        const unsigned long long lower = (1ull << (c - GEN_BIT)) - 1ull;
#pragma unroll
        for (int side = 0; side < 2; side++) {
            const int z = side ? z2 : z1;
            for (unsigned long long m = (side ? u2 : u1) & lower; m; m &= m - 1ull) {
                const int k = __ffsll((long long)m) - 1;
                const int2 f = __ldcg(&a.gadj[(size_t)z * GADJ + k]);
                if ((f.y & 0xffff) != R) add_ghost(f.x, GEN_BIT + k);  // (an own arbiter is a root of the walk anyway)
            }
        }
    };
    constexpr int SB = 4;  // arbiters a thread has in flight at once
    auto bodies_in = [&]() {
        if (lt < NCX) { s_cnt[lt] = 0; s_fill[lt] = 0; s_ccnt[lt] = 0; }
        if (lt == 0) { s_nbnd = 0; s_ngh = 0; s_ngb = 0; s_nown = 0; s_bad = (const unsigned char*)(s_body + 2 * (size_t)nb) > scratch_lo ? 1 : 0; }
        for (int k = lt; k < 3 * HASH_CAP; k += T) h_arb[k] = 0;
        __syncthreads();
        const int P = min(*(volatile int*)&cn->n_pairs, a.cap_pairs);
        for (int lb = lt; lb < nb && !s_bad; lb += T) {  // bodies: state, the arbiters of the body's row
            const int b = a.reg_bodies[rb0 + lb];
            const Rec32 vr = ld256(a.vrec + 2 * (size_t)b);  // { vx | ver, vy | ver } { w | ver, inv_mass | inv_moi } (k_prepare)
            const float4 p = a.pose[b];
            const int r0 = min(a.cand_row_start[b], P), r1 = b + 1 < a.B ? min(a.cand_row_start[b + 1], P) : P;
            const int e0 = (int)(__ldcg(&a.pscan[r0]) >> 32), e1 = min((int)(__ldcg(&a.pscan[r1]) >> 32), A);
            s_body[2 * lb] = make_float4(__int_as_float(vr.a.x), __int_as_float(vr.a.z), __int_as_float(vr.b.x), 0.0f);
            s_body[2 * lb + 1] = make_float4(p.x, p.y, __int_as_float(vr.b.z), __int_as_float(vr.b.w));
            if (e1 > e0) {
                const int k = atomicAdd(&s_nown, e1 - e0);
                for (int e = e0; e < e1; e++)
                    if (k + e - e0 < OWN_MAX) t_own[k + e - e0] = make_int2(e, -1); else s_bad = 1;
            }
        }
        __syncthreads();
        // own arbiters with an interior colour (persisting, or coloured by this CTA a moment ago): counted and adopted here; the generic
        // ones and the ones still without a colour (cross-region, promoted) are marked and wait for the grid barrier (t_own.y = -2)
        const int nown_ = min(s_nown, OWN_MAX);
        for (int k0 = lt; k0 < nown_; k0 += SB * T) {
            int e[SB], c[SB], n[SB];
            int4 ra[SB];
#pragma unroll
            for (int u = 0; u < SB; u++) e[u] = k0 + u * T < nown_ ? t_own[k0 + u * T].x : -1;
#pragma unroll
            for (int u = 0; u < SB; u++)
                if (e[u] >= 0) { c[u] = __ldcg(&a.cur.color[e[u]]); const Rec32 r = ld256(&a.srec[2 * (size_t)e[u]]); ra[u] = r.a; n[u] = r.b.x; }
#pragma unroll
            for (int u = 0; u < SB; u++) {
                if (e[u] < 0) continue;
                if (c[u] < 0 || c[u] >= GEN_BIT) { t_own[k0 + u * T].y = -2; continue; }
                t_own[k0 + u * T].y = c[u];
                atomicAdd(&s_cnt[c[u]], 1);
                adopt(c[u], ra[u], n[u]);
            }
        }
        __syncthreads();
    };
    if (!color_new_arbiters<true>(a, cn, bar, bar_target, sm_raw, a.smem_bytes, bodies_in)) grid_barrier(bar, bar_target);
    stamp(1);
    tstamp();
    if (*(volatile int*)&cn->err == SYLT_ERR_CAPACITY) return;  // an arbiter without a free colour (every CTA sees it: the colouring ends with a barrier)
    tstamp();
    unsigned long long t_setup0 = 0;
    if (a.dbg && lt == 0) asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t_setup0));

    // The set-up below is bound by the number of scattered 32-byte sectors a CTA has to fetch (~2 ns each per SM), so everything
    // it needs to know about an arbiter comes in one 32-byte record k_build wrote (`srec`: bodies with their static flags,
    // (region, local index) of both bodies | contact count, first contact, friction) and about a contact in another
    // (`cin`: position, normal | separation, pn, pt).
    // ---- 2b. the region's bodies into shared memory; its own arbiters (the rows its bodies own) listed and counted by colour
    const bool fits = (const unsigned char*)(s_body + 2 * (size_t)nb) <= scratch_lo;  // the region's bodies alone must leave room for the scratch (the host sizes the regions accordingly)
    for (int lb = lt; lb < nb && fits; lb += T) {  // boundary bodies (the colours at every body are final now)
        const int b = a.reg_bodies[rb0 + lb];
        const unsigned long long ub = __ldcg(&a.used[b]);  // (never set at a static body)
        if (ub != 0ull) {  // a boundary body: exported every pass
            const int k = atomicAdd(&s_nbnd, 1);
            s_bnd[k] = lb;
            s_bndg[k] = b;
            for (unsigned long long m = ub; m; m &= m - 1ull) {  // every generic arbiter at a boundary body is replayed here, whoever owns it
                const int kb = __ffsll((long long)m) - 1;
                const int2 f = __ldcg(&a.gadj[(size_t)b * GADJ + kb]);
                if ((f.y & 0xffff) != R) add_ghost(f.x, GEN_BIT + kb);
            }
        }
    }
    __syncthreads();
    const int nown = min(s_nown, OWN_MAX);
    for (int k0 = lt; k0 < nown; k0 += SB * T) {  // the own arbiters left over: the generic ones are the roots of the cone walk (level 0)
        int e[SB], c[SB], n[SB];
        int4 ra[SB];
#pragma unroll
        for (int u = 0; u < SB; u++) { e[u] = -1; if (k0 + u * T < nown) { const int2 t = t_own[k0 + u * T]; if (t.y == -2) e[u] = t.x; } }
#pragma unroll
        for (int u = 0; u < SB; u++)
            if (e[u] >= 0) { c[u] = __ldcg(&a.cur.color[e[u]]); const Rec32 r = ld256(&a.srec[2 * (size_t)e[u]]); ra[u] = r.a; n[u] = r.b.x; }
#pragma unroll
        for (int u = 0; u < SB; u++) {
            if (e[u] < 0) continue;
            t_own[k0 + u * T].y = c[u];
            if (c[u] < 0) continue;
            atomicAdd(&s_cnt[c[u]], 1);
            adopt(c[u], ra[u], n[u]);
        }
    }
    __syncthreads();
    tstamp();
    // ---- 2c. cone walk: ghost arbiters found in one level are adopted in the next, until nothing new turns up
    int levels = 0;
    for (int f0 = 0;; levels++) {
        const int f1 = min(s_ngh, CONE_CAP);
        __syncthreads();
        if (f1 == f0) break;
        for (int g = f0 + lt; g < f1; g += T) {
            const int2 f = ghosts[g];
            const Rec32 r = ld256(&a.srec[2 * (size_t)f.x]);
            adopt(f.y, r.a, r.b.x);
        }
        __syncthreads();
        f0 = f1;
    }
    const int ngh = min(s_ngh, CONE_CAP);
    tstamp();
    const int ngb_found = min(s_ngb, GB_MAX);
    tstamp();
    // ---- 2e. slot layout: generic colours first (their contacts must be shared-memory resident), then the interior colours; the
    //          contacts of a colour's slots are laid out colour by colour in the same order
    int* s_gb = (int*)(s_body + 2 * ((size_t)nb + ngb_found));  // ghost bodies: global index | GHOST_FLAG if static (never imported again)
    {   // (96 colours: one thread each, three warps; layout order of the warps: generic 32-63, generic 64-95, interior 0-31)
        static_assert(NCX == 96 && GEN_BIT == 32, "slot layout: three warps of colours");
        const int wq = lt >> 5, lq = lt & 31;
        int cntc = 0, ccn = 0, incl = 0, cincl = 0;
        unsigned bal = 0u;
        if (lt < NCX) {
            cntc = s_cnt[lt]; ccn = s_ccnt[lt];
            bal = __ballot_sync(0xffffffffu, cntc > 0);
            incl = warp_incl_scan(cntc); cincl = warp_incl_scan(ccn);
            if (lq == 31) { s_wtot[wq] = incl; s_wctot[wq] = cincl; }
            if (lq == 0) s_wbal[wq] = bal;
        }
        __syncthreads();
        if (lt < NCX) {
            const int before = wq == 1 ? 0 : wq == 2 ? s_wtot[1] : s_wtot[1] + s_wtot[2];
            const int cbefore = wq == 1 ? 0 : wq == 2 ? s_wctot[1] : s_wctot[1] + s_wctot[2];
            s_cstart[lt] = before + incl - cntc;
            s_cfill[lt] = cbefore + cincl - ccn;
            if (cntc > 0) {  // the colour list: interior colours ascending, then the generic ones
                const int below = __popc(bal & ((1u << lq) - 1u));
                s_clist[wq == 0 ? below : __popc(s_wbal[0]) + (wq == 2 ? __popc(s_wbal[1]) : 0) + below] = lt;
            }
        }
        if (lt == 0) {
            const int run = s_wtot[0] + s_wtot[1] + s_wtot[2], ngen = s_wtot[1] + s_wtot[2];
            const unsigned mask_int = s_wbal[0];
            const unsigned long long mask = (unsigned long long)s_wbal[1] | ((unsigned long long)s_wbal[2] << 32);
            s_cstart[NCX] = run;
            s_nint = __popc(mask_int); s_ncl = __popc(mask_int) + __popcll(mask); s_L = run;
            if (mask) atomicOr(&cn->color_mask, mask);
            if (mask_int) atomicOr(&cn->color_mask_int, mask_int);
            if (ngen) atomicAdd(&cn->n_generic, ngen - ngh);
            if (ngh) atomicAdd(&cn->n_ghost, ngh);
            atomicMax(&cn->max_slots, run);
            // everything but the contacts must fit below the scratch; bodies are referenced through 15-bit indices
            const bool room = (const unsigned char*)(s_gb + ngb_found + 4 * (size_t)run) <= scratch_lo && nb + ngb_found < 0x8000;
            if (s_bad || !room) {
                s_bad = 1;
                atomicExch(&cn->err, SYLT_ERR_CAPACITY);
                atomicOr(&cn->why, 1 | (!room ? 2 : 0) | (s_ngh > CONE_CAP ? 4 : 0) | (s_ngb > GB_MAX ? 8 : 0) | (!fits ? 16 : 0) | (s_nown > OWN_MAX ? 64 : 0));
            }
        }
    }
    __syncthreads();
    if (!a.run_solver) {
        if (tid == 0 && !step_failed(cn)) step_completed(cn);  // (a failed set-up is reported; nothing else depends on it here)
        return;
    }
    const bool bad = s_bad != 0;  // this CTA cannot solve; the error is flagged, the grid barrier below is still joined
    const int ngb = bad ? 0 : ngb_found;
    const int L = bad ? 0 : s_L, ncl = bad ? 0 : s_ncl, nint = bad ? 0 : s_nint, nbnd = bad ? 0 : s_nbnd;  // (a CTA that cannot solve idles through the passes)
    int* s_sb = s_gb + ngb;                               // slot: body refs, r1 | r2 << 16; ref = index into s_body | SREF_STATIC
    int* s_sc = s_sb + L;                                 //       contact offset << 8 | count, or -1 - count: contacts in the global arrays
    float* s_sf = (float*)(s_sc + L);                     //       friction
    int* s_so = (int*)(s_sf + L);                         //       first contact in the global arrays
    float4* s_c0 = (float4*)(((size_t)(s_so + L) + 15) & ~(size_t)15);  // contact: (position, normal)
    const int CONCAP = bad ? 0 : (int)(((sm_raw + a.smem_bytes - (unsigned char*)s_c0) / 36) & ~(size_t)3);
    float4* s_c1 = s_c0 + CONCAP;                         //          (mass_normal, mass_tangent, bias, pn)
    float* s_c2 = (float*)(s_c1 + CONCAP);                //          pt
    for (int g = lt; g < ngb && !bad; g += T) {  // ghosts: positions, inverse masses (and, for static ones, the velocity) once; dynamic velocities arrive every pass
        const int zf = t_gb[g], z = zf & ~GHOST_FLAG;
        s_gb[g] = zf;
        const Rec32 vr = ld256(a.vrec + 2 * (size_t)z);
        const float4 p = a.pose[z];
        s_body[2 * (nb + g)] = make_float4(__int_as_float(vr.a.x), __int_as_float(vr.a.z), __int_as_float(vr.b.x), 0.0f);
        s_body[2 * (nb + g) + 1] = make_float4(p.x, p.y, __int_as_float(vr.b.z), __int_as_float(vr.b.w));
    }
    const bool acc = a.accumulate != 0;
    const float kbias = a.poscorr ? 0.2f : 0.0f;

    stamp(2);
    tstamp();
    // ---- 3. slots: position by colour, record (a ghost body's shared-memory index comes from the hash table) and contacts.  Every
    //         record and contact is read into registers (PB slots per thread) before the first contact is stored: the contacts
    //         overwrite the set-up scratch (hash tables, lists).
    constexpr int SREF_STATIC = 0x8000;
    const int n_generic_slots = s_cstart[0];  // generic colours come first in the layout
    auto place = [&](int c, int flag, const int4 ra, const int4 rb) {
        const int ls = s_cstart[c] + atomicAdd(&s_fill[c], 1);
        if (ls >= s_cstart[c] + s_cnt[c]) { atomicExch(&cn->err, SYLT_ERR_INTERNAL); atomicOr(&cn->why, 128); return; }  // (never: counted and placed from the same lists)
        const int lc = atomicAdd(&s_cfill[c], rb.x);
        const bool res = lc + rb.x <= CONCAP;
        const int ref1 = ((ra.z >> 16) == R ? (ra.z & 0xffff) : nb + body_lookup(ra.x & SR_BODY)) | ((ra.x & SR_STATIC) ? SREF_STATIC : 0);
        const int ref2 = ((ra.w >> 16) == R ? (ra.w & 0xffff) : nb + body_lookup(ra.y & SR_BODY)) | ((ra.y & SR_STATIC) ? SREF_STATIC : 0);
        s_sb[ls] = ref1 | (ref2 << 16);
        s_sc[ls] = res ? (lc << 8) | rb.x : -1 - rb.x;
        s_so[ls] = rb.y | flag;  // contact offsets stay below 2^30
        s_sf[ls] = __int_as_float(rb.z);
        if (!res) {  // contacts stay in the global arrays
            atomicAdd(&cn->n_nonres, 1);
            if (ls < n_generic_slots) { atomicExch(&cn->err, SYLT_ERR_CAPACITY); atomicOr(&cn->why, 32); }  // a replica has no global working storage
        }
    };
    const int nall = bad ? 0 : nown + ngh;
    for (int k0 = lt; k0 < nall; k0 += SB * T) {
        int e[SB], c[SB];
        int4 ra[SB], rb[SB];
#pragma unroll
        for (int u = 0; u < SB; u++) {
            const int k = k0 + u * T;
            e[u] = -1; c[u] = 0;
            if (k < nown) { const int2 ec = t_own[k]; if (ec.y >= 0) { e[u] = ec.x; c[u] = ec.y; } }
            else if (k < nall) { const int2 f = ghosts[k - nown]; e[u] = f.x; c[u] = f.y; }
        }
#pragma unroll
        for (int u = 0; u < SB; u++)
            if (e[u] >= 0) { const Rec32 r = ld256(&a.srec[2 * (size_t)e[u]]); ra[u] = r.a; rb[u] = r.b; }
#pragma unroll
        for (int u = 0; u < SB; u++)
            if (e[u] >= 0) place(c[u], k0 + u * T < nown ? 0 : GHOST_FLAG, ra[u], rb[u]);
    }
    __syncthreads();  // the scratch is dead from here on
    tstamp();
    tstamp();
    for (int ls0 = lt; ls0 < L; ls0 += 2 * T) {  // contacts: two slots (up to four contacts) in flight per thread
        int sc[2], co[2];
        float4 ci[2][2][2];
#pragma unroll
        for (int v = 0; v < 2; v++) {
            const int ls = ls0 + v * T;
            sc[v] = ls < L ? s_sc[ls] : -1;
            co[v] = ls < L ? (s_so[ls] & ~GHOST_FLAG) : 0;
        }
#pragma unroll
        for (int v = 0; v < 2; v++)
#pragma unroll
            for (int q = 0; q < 2; q++)
                if (sc[v] >= 0 && q < (sc[v] & 0xff)) {
                    const Rec32 r = ld256(&a.cur.cin[2 * (size_t)(co[v] + q)]);
                    ci[v][q][0] = as_f4(r.a);
                    ci[v][q][1] = as_f4(r.b);
                }
#pragma unroll
        for (int v = 0; v < 2; v++) {
            if (sc[v] < 0) continue;
            const int lc = sc[v] >> 8, n = sc[v] & 0xff;
            // the velocity-independent half of Arbiter::pre_step (arbiter.rs:192-215) right here: this loop waits for memory, not for arithmetic
            const int sb = s_sb[ls0 + v * T];
            const float4 m1 = s_body[2 * (sb & 0x7fff) + 1], m2 = s_body[2 * ((sb >> 16) & 0x7fff) + 1];
            auto put_contact = [&](int l, const float4 x0, const float4 x1) {  // x0 = (position, normal), x1 = (separation, pn, pt, -)
                ContactConst k0;
                contact_constants(mk(x0.x, x0.y), mk(x0.z, x0.w), x1.x, mk(m1.x, m1.y), mk(m2.x, m2.y), m1.z, m1.w, m2.z, m2.w, a.inv_dt, kbias, k0);
                s_c0[l] = x0;
                s_c1[l] = make_float4(k0.mass_n, k0.mass_t, k0.bias, x1.y);
                s_c2[l] = x1.z;
            };
#pragma unroll
            for (int q = 0; q < 2; q++)
                if (q < n) put_contact(lc + q, ci[v][q][0], ci[v][q][1]);
            for (int q = 2; q < n; q++)  // (polygon manifolds)
                put_contact(lc + q, a.cur.cin[2 * (size_t)(co[v] + q)], a.cur.cin[2 * (size_t)(co[v] + q) + 1]);
        }
    }
    tstamp();
    if (lt < ncl) { const int c = s_clist[lt]; s_hop[lt] = make_int2(s_cstart[c], s_cnt[c]); }  // (visible behind the barrier in front of the passes)
    if (a.dbg && lt == 0) {  // diagnostics: the CTA's load
        int mxc = 0;
        for (int k = 0; k < ncl; k++) mxc = max(mxc, s_cnt[s_clist[k]]);
        unsigned long long t_setup1;
        asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t_setup1));
        unsigned long long* d = a.dbg + 64 + 4 * 1024 + bid * 4;
        d[0] = (unsigned long long)L | ((unsigned long long)levels << 24) | ((t_setup1 - t_setup0) << 32);
        d[1] = (unsigned long long)nint | ((unsigned long long)(ncl - nint) << 32);
        d[2] = (unsigned long long)ngh | ((unsigned long long)ngb << 32);
        d[3] = (unsigned long long)nb | ((unsigned long long)mxc << 32);
    }
    __syncthreads();
    tstamp();
    // No grid barrier here: a CTA starts its passes as soon as it is set up and from then on only ever waits for the exports
    // of its neighbours.  (A CTA that could not set itself up has flagged the error; whoever waits for it gives up, see the
    // import loop, and nothing is written back.)

    // One slot of one colour hop.  Bodies and (normally) contacts are in shared memory; r1, r2 are recomputed from the
    // contact position exactly as Arbiter::pre_step / apply_impulse do (arbiter.rs:192-193, 236-237).
    // -DSYLT_HOP_PROBE: cycle stamps inside the colour hops of pass 5 in the probe CTA (diagnostics build only: the probes cost ~10 % of
    // the loop's instructions)
#ifdef SYLT_HOP_PROBE
    bool hop_probe = false;
    int hop_k = 0;
#define HOPCLK(k, dep) do { if (hop_probe) { unsigned long long t_; float d_ = (dep); asm volatile("mov.u64 %0, %%clock64;" : "=l"(t_) : "f"(d_) : "memory"); a.dbg[64 + 8 * 1024 + hop_k * 8 + (k)] = t_; } } while (0)
#else
#define HOPCLK(k, dep) do { } while (0)
#endif
    auto solve_slot = [&](int ls, unsigned pass) {
        const int sb = s_sb[ls], sc = s_sc[ls];
        const float fr = s_sf[ls];
        const int r1 = sb & 0xffff, r2 = (sb >> 16) & 0xffff;
        const int i1 = r1 & ~SREF_STATIC, i2 = r2 & ~SREF_STATIC;
        const float4 q1 = s_body[2 * i1], q2 = s_body[2 * i2], m1 = s_body[2 * i1 + 1], m2 = s_body[2 * i2 + 1];
        Vel v1, v2;
        v1.v = mk(q1.x, q1.y); v1.w = q1.z; v2.v = mk(q2.x, q2.y); v2.w = q2.z;
        const V2 p1 = mk(m1.x, m1.y), p2 = mk(m2.x, m2.y);
        const float im1 = m1.z, ii1 = m1.w, im2 = m2.z, ii2 = m2.w;
        if (sc >= 0 && pass != 0u && (sc & 0xff) <= 2) {
            // the usual slot of an iteration (box contacts: one or two, resident): everything is loaded before the dependent chain starts
            const int lc = sc >> 8, n = sc & 0xff, l2 = lc + max(n, 1) - 1;
            const float4 c0a = s_c0[lc], c1a = s_c1[lc], c0b = s_c0[l2], c1b = s_c1[l2];
            const float pta = s_c2[lc], ptb = s_c2[l2];
            HOPCLK(1, c0b.x + c1b.x + ptb + q1.x + q2.x + m1.x + m2.x);
            if (n >= 1) {
                ContactConst k0;
                k0.n = mk(c0a.z, c0a.w); k0.r1 = mk(c0a.x, c0a.y) - p1; k0.r2 = mk(c0a.x, c0a.y) - p2;
                k0.mass_n = c1a.x; k0.mass_t = c1a.y; k0.bias = c1a.z;
                float pn = c1a.w, pt = pta;
                contact_apply(k0, fr, im1, ii1, im2, ii2, acc, pn, pt, v1, v2);
                if (acc) { s_c1[lc].w = pn; s_c2[lc] = pt; }
            }
            HOPCLK(2, v1.v.x + v2.v.x + v1.w + v2.w);
            if (n == 2) {
                ContactConst k0;
                k0.n = mk(c0b.z, c0b.w); k0.r1 = mk(c0b.x, c0b.y) - p1; k0.r2 = mk(c0b.x, c0b.y) - p2;
                k0.mass_n = c1b.x; k0.mass_t = c1b.y; k0.bias = c1b.z;
                float pn = c1b.w, pt = ptb;
                contact_apply(k0, fr, im1, ii1, im2, ii2, acc, pn, pt, v1, v2);
                if (acc) { s_c1[l2].w = pn; s_c2[l2] = pt; }
            }
            HOPCLK(3, v1.v.x + v2.v.x + v1.w + v2.w);
        } else if (sc >= 0) {
            const int lc = sc >> 8, n = sc & 0xff;
            for (int q = 0; q < n; q++) {
                const float4 c0 = s_c0[lc + q], c1 = s_c1[lc + q];
                const float pt0 = s_c2[lc + q];
                ContactConst k0;
                if (pass == 0u) {  // the ordered half of Arbiter::pre_step (arbiter.rs:216-223); the constants were computed with the set-up
                    k0.n = mk(c0.z, c0.w); k0.r1 = mk(c0.x, c0.y) - p1; k0.r2 = mk(c0.x, c0.y) - p2;
                    if (acc) contact_warmstart(k0, im1, ii1, im2, ii2, c1.w, pt0, v1, v2);
                } else {
                    k0.n = mk(c0.z, c0.w); k0.r1 = mk(c0.x, c0.y) - p1; k0.r2 = mk(c0.x, c0.y) - p2;
                    k0.mass_n = c1.x; k0.mass_t = c1.y; k0.bias = c1.z;
                    float pn = c1.w, pt = pt0;
                    contact_apply(k0, fr, im1, ii1, im2, ii2, acc, pn, pt, v1, v2);
                    if (acc) { s_c1[lc + q].w = pn; s_c2[lc + q] = pt; }
                }
            }
        } else if (!(s_so[ls] & GHOST_FLAG)) {  // an own slot beyond the shared-memory budget: the global contact arrays are its working
            // storage (a replica in that situation has failed the step, SYLT_ERR_CAPACITY, and is skipped: it must not touch them)
            const int2 mt = make_int2(-1 - sc, s_so[ls]);
            for (int q = 0; q < mt.x; q++) {
                ContactConst k0;
                if (pass == 0u) {
                    const float4 x0 = a.cur.cin[2 * (size_t)(mt.y + q)], x1 = a.cur.cin[2 * (size_t)(mt.y + q) + 1];  // (position, normal) (separation, pn, pt, -)
                    contact_prestep(mk(x0.x, x0.y), mk(x0.z, x0.w), x1.x, p1, p2, im1, ii1, im2, ii2, a.inv_dt, kbias, acc, x1.y, x1.z, v1, v2, k0);
                    a.cur.ca[mt.y + q] = make_float4(x0.z, x0.w, k0.r1.x, k0.r1.y);
                    a.cur.cb[mt.y + q] = make_float4(k0.r2.x, k0.r2.y, k0.mass_n, k0.mass_t);
                    a.cur.cc[mt.y + q] = make_float4(k0.bias, x1.y, x1.z, x1.x);
                } else {
                    const float4 ca = a.cur.ca[mt.y + q], cc = a.cur.cc[mt.y + q];
                    const float4 cb = a.cur.cb[mt.y + q];
                    k0.n = mk(ca.x, ca.y); k0.r1 = mk(ca.z, ca.w); k0.r2 = mk(cb.x, cb.y);
                    k0.mass_n = cb.z; k0.mass_t = cb.w; k0.bias = cc.x;
                    float pn = cc.y, pt = cc.z;
                    contact_apply(k0, fr, im1, ii1, im2, ii2, acc, pn, pt, v1, v2);
                    if (acc) a.cur.cc[mt.y + q] = make_float4(cc.x, pn, pt, cc.w);
                }
            }
        }
        if (!(r1 & SREF_STATIC)) s_body[2 * i1] = make_float4(v1.v.x, v1.v.y, v1.w, 0.0f);
        if (!(r2 & SREF_STATIC)) s_body[2 * i2] = make_float4(v2.v.x, v2.v.y, v2.w, 0.0f);
    };
    int pa_ = 0;
    auto pstamp = [&](unsigned pass) {  // diagnostics: the phases of pass 5 in every CTA
        if (a.dbg && lt == 0 && pass == 5u && pa_ < 4) { unsigned long long t; asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t)); a.dbg[64 + bid * 4 + pa_++] = t; }
    };
    auto colours = [&](unsigned pass, int k0, int k1) {  // colour hops k0..k1 of the list
        for (int k = k0; k < k1; k++) {
            const int2 h = s_hop[k];  // (first slot, slots) of the hop's colour
#ifdef SYLT_HOP_PROBE
            hop_probe = a.dbg && bid == a.probe_cta && lt == 0 && pass == 5u && k < 8;
            hop_k = k;
#endif
            HOPCLK(0, 0.0f);
            for (int i = lt; i < h.y; i += T) solve_slot(h.x + i, pass);
            HOPCLK(4, 0.0f);
            __syncthreads();
            HOPCLK(5, 0.0f);
#ifdef SYLT_HOP_PROBE
            if (hop_probe) a.dbg[64 + 8 * 1024 + k * 8 + 6] = (unsigned long long)h.y;
            hop_probe = false;
#endif
        }
    };
    // One pass: interior colours; boundary bodies out (plane `pass` of the export records); ghosts in; generic colours.
    auto sweep = [&](unsigned pass) {
        pstamp(pass);
        colours(pass, 0, nint);
        pstamp(pass);
        BodyIO xio{a.xrec + (size_t)(pass % MAX_PLANES) * 2 * (size_t)a.B, a.sleep_ns, &cn->err};
        const unsigned tag = (a.step_tag << 12) | pass;
        for (int k = lt; k < nbnd; k += T) {
            const float4 q = s_body[2 * s_bnd[k]];
            Vel o; o.v = mk(q.x, q.y); o.w = q.z;
            xio.put256(s_bndg[k], o, tag);
        }
        for (int g = lt; g < ngb; g += T) {
            const int z = s_gb[g];
            if (z & GHOST_FLAG) continue;  // static
            Vel o;
            unsigned spins = 0;
            while (!xio.try_get256(z, tag, o)) {
                if ((++spins & 0x3ffu) == 0u) {
                    if (*(volatile int*)&cn->err != 0) break;  // a failed step: the owner may never export
                    if (spins > (1u << 23)) { atomicExch(&cn->err, SYLT_ERR_INTERNAL); break; }
                }
            }
            s_body[2 * (nb + g)] = make_float4(o.v.x, o.v.y, o.w, 0.0f);
        }
        __syncthreads();
        pstamp(pass);
        colours(pass, nint, ncl);
        pstamp(pass);
    };

    stamp(3);
    // ---- 4. pass 0: Arbiter::pre_step (arbiter.rs:182-228).  Its ordered half is the warm start (:216-223); without accumulation there is
    //         none, and with every accumulated impulse zero (warm starting off: k_build zeroes them) it subtracts +-0 from every
    //         velocity — which changes nothing unless a velocity is exactly -0.0 (k_prepare looks).  Then the pre-step is just the
    //         constants of every contact (:192-215), and those were computed while the contacts were loaded: no colour hops, no hand-off.
    const bool flat_prestep = !acc || (!a.warm && *(volatile unsigned*)&cn->neg_zero_tag != a.step_serial);  // grid-uniform
    if (flat_prestep) {
        for (int ls = lt; ls < L; ls += T) {  // (only the own slots beyond the shared-memory budget are left: the global arrays are their storage)
            const int sc = s_sc[ls];
            if (sc >= 0 || (s_so[ls] & GHOST_FLAG)) continue;
            const int sb = s_sb[ls];
            const float4 m1 = s_body[2 * (sb & 0x7fff) + 1], m2 = s_body[2 * ((sb >> 16) & 0x7fff) + 1];
            const int n = -1 - sc, coff = s_so[ls];
            for (int q = 0; q < n; q++) {
                const float4 x0 = a.cur.cin[2 * (size_t)(coff + q)], x1 = a.cur.cin[2 * (size_t)(coff + q) + 1];
                ContactConst k0;
                contact_constants(mk(x0.x, x0.y), mk(x0.z, x0.w), x1.x, mk(m1.x, m1.y), mk(m2.x, m2.y), m1.z, m1.w, m2.z, m2.w, a.inv_dt, kbias, k0);
                a.cur.ca[coff + q] = make_float4(x0.z, x0.w, k0.r1.x, k0.r1.y);
                a.cur.cb[coff + q] = make_float4(k0.r2.x, k0.r2.y, k0.mass_n, k0.mass_t);
                a.cur.cc[coff + q] = make_float4(k0.bias, x1.y, x1.z, x1.x);
            }
        }
        __syncthreads();
    } else {
        sweep(0u);
    }
    stamp(4);
    // ---- 5. passes 1..I: apply_impulse (world.rs:132-140)
    for (int it = 1; it <= a.iterations; it++) {
        if (it % MAX_PLANES == 0) grid_barrier(bar, bar_target);  // (the reference's samples run 100 iterations: examples/samples/src/main.rs:10)
        sweep((unsigned)it);
    }

    stamp(5);
    // ---- 6. the owner writes its shared-memory resident contacts back (warm start of the next step, read-back).
    //         The one barrier of the solve: every replica has read the global contact arrays and the poses (set-up) before
    //         any owner overwrites them.
    grid_barrier(bar, bar_target);
    if (*(volatile int*)&cn->err != 0) return;
    for (int ls = lt; ls < L; ls += T) {
        const int sc = s_sc[ls], so = s_so[ls];
        if (sc < 0 || (so & GHOST_FLAG)) continue;
        // Only what the next step reads (the accumulated impulses) and the bias go back: r1, r2, mass_normal, mass_tangent and the
        // separation are only ever read by sylt_world_get_arbiters, which has k_fill_contacts recompute them — the same
        // contact_constants() on the same inputs (cin, the poses from before the integration, the masses) gives the same bits.
        const int coff = so, lc = sc >> 8, n = sc & 0xff;
        for (int q = 0; q < n; q++) {
            const float4 c1 = s_c1[lc + q];
            a.cur.cc[coff + q] = make_float4(c1.z, c1.w, s_c2[lc + q], 0.0f);
        }
    }

    stamp(6);
    // ---- 7. integrate the region's bodies (world.rs:143-150; static bodies too, quirk Q-16)
    for (int lb = lt; lb < nb; lb += T) {
        const int b = a.reg_bodies[rb0 + lb];
        const float4 q = s_body[2 * lb];
        a.vel[b] = make_float4(q.x, q.y, q.z, 0.0f);
        float4 p = a.pose[b];
        a.pose_prev[b] = p;  // (for k_fill_contacts)
        V2 np = mk(p.x, p.y) + mk(q.x, q.y) * a.dt;
        p.x = np.x; p.y = np.y;
        p.z += q.z * a.dt;
        a.pose[b] = p;
        a.force[b] = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
    }
    if (tid == 0) step_completed(cn);  // (past the barrier above every CTA has seen err == 0)
    stamp(7);
}

// ------------------------------------------------------------------ re-cut of the regions (every `regions_period` steps)
// Spatial order of the bodies for the region solver: a counting sort by the Z-order (Morton) index of a 128 x 128 grid laid
// over the bounding box of the bodies, then by body index inside a cell (so the order, and with it the solve order, is
// reproducible); consecutive runs of that order are compact blobs.  The order is then cut into G regions of equal load: weight
// of a body = 1 + 2 x (arbiters it owned last step), floored at the average weight so that no region gets more than twice the
// average number of bodies (k_solve_regions keeps a region's bodies in shared memory).  Regions are consecutive runs of `perm`,
// so perm itself is the CSR body list.  Any order is valid; compactness only decides how many arbiters cross regions.
// ONE cooperative launch over all SMs (grid barriers between the phases; every sum is an integer, so the cut does not depend on
// the number of CTAs): ~50 us at 100k bodies, where the single-CTA kernels it replaces took ~800 us.
constexpr int RECUT_THREADS = 512, SORT_BITS = 7, SORT_CELLS = 1 << (2 * SORT_BITS), TRIM_BINS = 1024, TRIM_ROUNDS = 4;
struct RecutScratch {  // zeroed by a memset in front of every launch
    unsigned bar;
    unsigned box[4];                         // order-preserving unsigned images of (-x0, x1, -y0, y1): all reduced with atomicMax (0 = nothing yet)
    int trim[TRIM_ROUNDS][2 * TRIM_BINS];    // coordinate histograms of the trimming rounds (x, then y)
    int hist[SORT_CELLS + 1];                // bodies per Morton cell
    int fill[SORT_CELLS];                    // scatter cursors
    long long part[1024], epart[1024];       // per-CTA sums of the weights / the floored weights
    int first[1025];                         // B - (first position of region g), reduced with atomicMax (0 = empty)
    unsigned long long stamp[16];            // %globaltimer of CTA 0 at the phase boundaries (diagnostics, SYLT_SOLVER_TIMING)
};
__device__ __forceinline__ unsigned morton_spread(unsigned v) {  // 0b abcdefg -> 0b 0a0b0c0d0e0f0g
    v = (v | (v << 4)) & 0x0f0fu;
    v = (v | (v << 2)) & 0x3333u;
    v = (v | (v << 1)) & 0x5555u;
    return v;
}
__device__ __forceinline__ int float_order(float f) { const int i = __float_as_int(f); return i >= 0 ? i : i ^ 0x7fffffff; }   // monotonic float -> int
__device__ __forceinline__ float order_float(int i) { return __int_as_float(i >= 0 ? i : i ^ 0x7fffffff); }

__global__ void __launch_bounds__(RECUT_THREADS, 1) k_recut(int B, int G, const float4* pose, const int* bflags, int n_groups, const int* old_row_start,
                                                           int B_old, int* perm, int* key_tmp, int* region_of, int* lidx, int* reg_start,
                                                           RecutScratch* sc, const Counters* cn) {
    extern __shared__ int s_cells[];  // SORT_CELLS + 1: cell offsets (also staging for the trimming histograms)
    __shared__ long long sh64[34];
    __shared__ int sh32[34];
    __shared__ int s_lohi[4];
    if (step_failed(cn)) return;  // grid-uniform at entry
    const int lt = threadIdx.x, T = blockDim.x, lane = lt & 31, wid = lt >> 5;
    const int tid = blockIdx.x * T + lt, nth = gridDim.x * T;
    unsigned bar_target = 0;
    int n_stamp = 0;
    auto stamp = [&]() { if (tid == 0 && n_stamp < 16) { unsigned long long t; asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t)); sc->stamp[n_stamp++] = t; } };
    stamp();
    auto finite = [](float4 p) { return fabsf(p.x) < 1.0e30f && fabsf(p.y) < 1.0e30f; };
    // ---- 1. bounding box of the (finite) body positions
    {
        float x0 = 3.0e38f, x1 = -3.0e38f, y0 = 3.0e38f, y1 = -3.0e38f;
        for (int i = tid; i < B; i += nth) {
            const float4 p = pose[i];
            if (finite(p)) { x0 = fminf(x0, p.x); x1 = fmaxf(x1, p.x); y0 = fminf(y0, p.y); y1 = fmaxf(y1, p.y); }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            x0 = fminf(x0, __shfl_xor_sync(0xffffffffu, x0, o)); x1 = fmaxf(x1, __shfl_xor_sync(0xffffffffu, x1, o));
            y0 = fminf(y0, __shfl_xor_sync(0xffffffffu, y0, o)); y1 = fmaxf(y1, __shfl_xor_sync(0xffffffffu, y1, o));
        }
        if (lane == 0 && x0 <= x1) {
            atomicMax(&sc->box[0], (unsigned)float_order(-x0) ^ 0x80000000u); atomicMax(&sc->box[1], (unsigned)float_order(x1) ^ 0x80000000u);
            atomicMax(&sc->box[2], (unsigned)float_order(-y0) ^ 0x80000000u); atomicMax(&sc->box[3], (unsigned)float_order(y1) ^ 0x80000000u);
        }
    }
    grid_barrier(&sc->bar, bar_target);
    float x0, x1, y0, y1;
    {
        auto get = [&](int k) { const unsigned v = *(volatile unsigned*)&sc->box[k]; return v == 0u ? 0.0f : order_float((int)(v ^ 0x80000000u)); };
        x0 = -get(0); x1 = get(1); y0 = -get(2); y1 = get(3);
    }
    // ---- 2. a body that has been shot far away must not blow the grid up to one cell for everybody else: trim the box to the
    //         0.1 % .. 99.9 % range of the coordinates (1024-bin histograms), again while that shrinks it a lot.  Bodies outside
    //         land in the edge cells.  (Every CTA takes the same decision from the same histogram.)
    for (int round = 0; round < TRIM_ROUNDS; round++) {
        const float sx = (float)TRIM_BINS / fmaxf(x1 - x0, 1.0e-6f), sy = (float)TRIM_BINS / fmaxf(y1 - y0, 1.0e-6f);
        for (int k = lt; k < 2 * TRIM_BINS; k += T) s_cells[k] = 0;
        __syncthreads();
        for (int i = tid; i < B; i += nth) {
            const float4 p = pose[i];
            if (!finite(p)) continue;
            atomicAdd(&s_cells[min(TRIM_BINS - 1, max(0, (int)((p.x - x0) * sx)))], 1);
            atomicAdd(&s_cells[TRIM_BINS + min(TRIM_BINS - 1, max(0, (int)((p.y - y0) * sy)))], 1);
        }
        __syncthreads();
        for (int k = lt; k < 2 * TRIM_BINS; k += T)
            if (s_cells[k]) atomicAdd(&sc->trim[round][k], s_cells[k]);
        grid_barrier(&sc->bar, bar_target);
        for (int k = lt; k < 2 * TRIM_BINS; k += T) s_cells[k] = __ldcg(&sc->trim[round][k]);
        __syncthreads();
        if (wid < 2) {  // warp 0: x, warp 1: y — lane l owns bins [32 l, 32 l + 32)
            const int* h = s_cells + TRIM_BINS * wid;
            const int q = max(1, B / 1000);
            int own = 0;
            for (int k = 0; k < 32; k++) own += h[32 * lane + k];
            int incl = own;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) { const int u = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += u; }
            const int total = __shfl_sync(0xffffffffu, incl, 31);
            // lo = the first bin at which the running count from the left exceeds q; hi = the same from the right
            int lo = TRIM_BINS - 1, hi = 0;
            {
                const unsigned m = __ballot_sync(0xffffffffu, incl > q);
                if (m) {
                    const int l = __ffs(m) - 1;
                    int acc = __shfl_sync(0xffffffffu, incl - own, l), b = 32 * l;
                    if (lane == l) { for (;; b++) { acc += h[b]; if (acc > q || b == TRIM_BINS - 1) break; } }
                    lo = __shfl_sync(0xffffffffu, b, l);
                }
                const int from_right = total - (incl - own);  // bins [32 lane, end)
                const unsigned mr = __ballot_sync(0xffffffffu, from_right > q);
                if (mr) {
                    const int l = 31 - __clz(mr);
                    int acc = __shfl_sync(0xffffffffu, total - incl, l), b = 32 * l + 31;
                    if (lane == l) { for (;; b--) { acc += h[b]; if (acc > q || b == 0) break; } }
                    hi = __shfl_sync(0xffffffffu, b, l);
                }
            }
            if (hi < lo) hi = lo;
            if (lane == 0) { s_lohi[2 * wid] = lo; s_lohi[2 * wid + 1] = hi + 1; }
        }
        __syncthreads();
        const float wx = (x1 - x0) / (float)TRIM_BINS, wy = (y1 - y0) / (float)TRIM_BINS;
        const float nx0 = x0 + (float)s_lohi[0] * wx, nx1 = x0 + (float)s_lohi[1] * wx, ny0 = y0 + (float)s_lohi[2] * wy, ny1 = y0 + (float)s_lohi[3] * wy;
        const bool shrunk = (nx1 - nx0) < 0.25f * (x1 - x0) || (ny1 - ny0) < 0.25f * (y1 - y0);
        __syncthreads();
        if (!shrunk) break;
        x0 = nx0; x1 = nx1; y0 = ny0; y1 = ny1;
    }
    stamp();
    // ---- 3. Morton keys + bodies per cell
    const float span = fmaxf(fmaxf(x1 - x0, y1 - y0), 1.0e-3f);
    const float inv = (float)(1 << SORT_BITS) / (span * 1.0001f);
    for (int i = tid; i < B; i += nth) {
        const float4 p = pose[i];
        int k = 0;
        if (n_groups > 1) k = (int)((long long)group_of(bflags[i]) * SORT_CELLS / n_groups);  // collision groups: a region is a run of whole groups
        else if (finite(p)) {
            const int cx = min((1 << SORT_BITS) - 1, max(0, (int)((p.x - x0) * inv))), cy = min((1 << SORT_BITS) - 1, max(0, (int)((p.y - y0) * inv)));
            k = (int)(morton_spread((unsigned)cx) | (morton_spread((unsigned)cy) << 1));
        }
        key_tmp[i] = k;
        atomicAdd(&sc->hist[k], 1);
    }
    grid_barrier(&sc->bar, bar_target);
    stamp();
    // ---- 4. cell offsets (every CTA scans the histogram for itself), scatter, then ascending body index inside a cell
    {
        for (int k = lt; k < SORT_CELLS; k += T) s_cells[k] = __ldcg(&sc->hist[k]);  // coalesced
        __syncthreads();
        // thread t scans cells [32 t, 32 t + 32) — read through a rotation so that the lanes of a warp hit different banks
        const int per = SORT_CELLS / RECUT_THREADS, q0 = lt * per;
        int c[SORT_CELLS / RECUT_THREADS];
        int sum = 0;
#pragma unroll
        for (int k = 0; k < per; k++) { const int kk = (k + lane) & (per - 1); c[k] = s_cells[q0 + kk]; sum += c[k]; }
        int tot;
        const int ex0 = block_excl_scan<int>(sum, &tot, sh32);
        // (the rotated order is only a read order: rebuild the running sum in cell order)
        int run = ex0;
#pragma unroll
        for (int k = 0; k < per; k++) {
            // cell q0 + k was read at rotated position (k - lane) mod per
            const int pos = (k - lane) & (per - 1);
            int v = 0;
#pragma unroll
            for (int j = 0; j < per; j++) v = j == pos ? c[j] : v;
            s_cells[q0 + k] = run;  // (own cells only: no other thread reads them before the barrier below)
            run += v;
        }
        if (lt == 0) s_cells[SORT_CELLS] = tot;
    }
    __syncthreads();
    stamp();
    for (int i = tid; i < B; i += nth) {
        const int k = key_tmp[i];
        perm[s_cells[k] + atomicAdd(&sc->fill[k], 1)] = i;
    }
    grid_barrier(&sc->bar, bar_target);
    stamp();
    // (one warp per cell, a rank sort in registers: the bodies of a cell sit in consecutive positions, ids are distinct; a pile of
    //  30 bodies in a cell would cost an insertion sort ~450 dependent global round trips)
    for (int k = tid >> 5; k < SORT_CELLS; k += nth >> 5) {
        const int lo = s_cells[k], hi = s_cells[k + 1], n = hi - lo;
        if (n <= 1 || n > 256) continue;  // a degenerate pile-up in one cell: leave it unordered rather than sort it here (warp-uniform)
        const int nr = (n + 31) >> 5;
        int v[8], rank[8];
#pragma unroll
        for (int r = 0; r < 8; r++) {
            const int idx = lo + lane + 32 * r;
            v[r] = (r < nr && idx < hi) ? __ldcg(&perm[idx]) : 0x7fffffff;
            rank[r] = 0;
        }
        __syncwarp();
#pragma unroll
        for (int r2 = 0; r2 < 8; r2++) {
            if (r2 >= nr) break;
            for (int src = 0; src < 32; src++) {
                const int y = __shfl_sync(0xffffffffu, v[r2], src);
#pragma unroll
                for (int r = 0; r < 8; r++) rank[r] += y < v[r] ? 1 : 0;
            }
        }
#pragma unroll
        for (int r = 0; r < 8; r++)
            if (v[r] != 0x7fffffff) perm[lo + rank[r]] = v[r];
    }
    grid_barrier(&sc->bar, bar_target);
    stamp();
    // ---- 5. cut the order into G regions of equal (floored) weight: CTA c owns positions [c0, c1), thread t a slice of those
    auto weight = [&](int b) -> long long { return (old_row_start && b < B_old) ? 1 + 2 * max(0, __ldg(&old_row_start[b + 1]) - __ldg(&old_row_start[b])) : 1; };
    const int cchunk = (B + (int)gridDim.x - 1) / (int)gridDim.x, c0 = min((int)blockIdx.x * cchunk, B), c1 = min(c0 + cchunk, B);
    const int tchunk = (c1 - c0 + T - 1) / T, q0 = min(c0 + lt * tchunk, c1), q1 = min(q0 + tchunk, c1);
    {
        long long sum = 0;
        for (int i = q0; i < q1; i++) sum += weight(__ldcg(&perm[i]));
        long long tot;
        block_excl_scan<long long>(sum, &tot, sh64);
        if (lt == 0) sc->part[blockIdx.x] = tot;
    }
    grid_barrier(&sc->bar, bar_target);
    long long total;  // (gridDim.x <= RECUT_THREADS: one partial per thread, summed by the block)
    block_excl_scan<long long>(lt < (int)gridDim.x ? __ldcg(&sc->part[lt]) : 0ll, &total, sh64);
    const long long wmin = (total + B - 1) / max(B, 1);
    long long ex;
    {
        long long esum = 0;
        for (int i = q0; i < q1; i++) esum += max(weight(__ldcg(&perm[i])), wmin);
        long long tot;
        ex = block_excl_scan<long long>(esum, &tot, sh64);
        if (lt == 0) sc->epart[blockIdx.x] = tot;
    }
    grid_barrier(&sc->bar, bar_target);
    long long etotal;
    {
        __shared__ long long s_before;
        const long long pre = block_excl_scan<long long>(lt < (int)gridDim.x ? __ldcg(&sc->epart[lt]) : 0ll, &etotal, sh64);
        if (lt == (int)blockIdx.x) s_before = pre;  // the floored weights of the CTAs in front of this one
        __syncthreads();
        ex += s_before;
    }
    for (int i = q0; i < q1; i++) {
        const int b = __ldcg(&perm[i]);
        const int reg = (int)min((long long)(G - 1), ex * (long long)G / max(etotal, 1ll));
        ex += max(weight(b), wmin);
        region_of[b] = reg;
        atomicMax(&sc->first[reg], B - i);  // = min over the positions
    }
    grid_barrier(&sc->bar, bar_target);
    // ---- 6. region starts (an empty region starts where the next one starts), then every body's index inside its region
    for (int g = lt; g < G; g += T) { const int v = __ldcg(&sc->first[g]); s_cells[g] = v > 0 ? B - v : B; }
    if (lt == 0) s_cells[G] = B;
    __syncthreads();
    if (lt == 0)
        for (int g = G - 1; g >= 0; g--) if (s_cells[g] > s_cells[g + 1]) s_cells[g] = s_cells[g + 1];
    __syncthreads();
    if (blockIdx.x == 0)
        for (int g = lt; g <= G; g += T) reg_start[g] = s_cells[g];
    for (int i = q0; i < q1; i++) { const int b = __ldcg(&perm[i]); const int reg = region_of[b]; lidx[b] = (reg << 16) | (i - s_cells[reg]); }  // (region, index inside it: < 2^15)
    stamp();
}

// the flags a step used become the "old" flags of the next one — unless the step failed (it will be repeated)
// The solve fields of a set the region solver wrote back that nothing on the device reads again — r1, r2, mass_normal, mass_tangent
// (arbiter.rs:192-215) and the separation — when somebody asks for the arbiters: contact_constants() on what the step itself used.
__global__ void k_fill_contacts(int A, const int2* bodies, const int2* meta, const float4* cin, const float4* pose_prev, const float4* mp,
                                float inv_dt, float kbias, float4* ca, float4* cb, float4* cc) {
    const int e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= A) return;
    const int2 b = bodies[e], mt = meta[e];
    const float4 p1 = pose_prev[b.x], p2 = pose_prev[b.y], m1 = mp[b.x], m2 = mp[b.y];
    for (int q = 0; q < mt.x; q++) {
        const float4 x0 = cin[2 * (size_t)(mt.y + q)], x1 = cin[2 * (size_t)(mt.y + q) + 1];
        ContactConst k0;
        contact_constants(mk(x0.x, x0.y), mk(x0.z, x0.w), x1.x, mk(p1.x, p1.y), mk(p2.x, p2.y), m1.x, m1.y, m2.x, m2.y, inv_dt, kbias, k0);
        ca[mt.y + q] = make_float4(x0.z, x0.w, k0.r1.x, k0.r1.y);
        cb[mt.y + q] = make_float4(k0.r2.x, k0.r2.y, k0.mass_n, k0.mass_t);
        cc[mt.y + q].w = x1.x;
    }
}

// SyltBodyState records (x, y, rotation, vx, vy, angular_velocity, fx, fy, torque) for sylt_world_get_bodies
__global__ void k_pack_states(int n, const float4* pose, const float4* vel, const float4* force, float* out) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float4 p = pose[i], v = vel[i], f = force[i];
    float* o = out + 9 * (size_t)i;
    o[0] = p.x; o[1] = p.y; o[2] = p.z; o[3] = v.x; o[4] = v.y; o[5] = v.z; o[6] = f.x; o[7] = f.y; o[8] = f.z;
}

__global__ void k_copy_flags(int B, const int* src, int* dst, const Counters* cn) {
    if (step_failed(cn)) return;
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < B) dst[i] = src[i];
}

__global__ void k_collide_pairs(int n, const SyltBodyDesc* A, const SyltBodyDesc* Bd, const float2* lverts, const int* loff,
                                SyltContact* out, int cap, int* counts) {
    int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= n) return;
    const SyltBodyDesc a = A[p], b = Bd[p];
    ContactOut o[MAX_MANIFOLD];
    float ca = sylt_cosf(a.rotation), sa = sylt_sinf(a.rotation), cb = sylt_cosf(b.rotation), sb = sylt_sinf(b.rotation);
    int m;
    if (a.shape == SYLT_SHAPE_BOX && b.shape == SYLT_SHAPE_BOX) {
        m = box_box(mk(a.x, a.y), ca, sa, mk(a.width_x, a.width_y), mk(b.x, b.y), cb, sb, mk(b.width_x, b.width_y), o);
    } else {
        V2 w0[MAX_POLY_VERTS], w1[MAX_POLY_VERTS];
        int n0 = loff[4 * p + 1], n1 = loff[4 * p + 3];
        M2 r0 = rot_cs(ca, sa), r1 = rot_cs(cb, sb);
        for (int k = 0; k < n0; k++) { float2 l = lverts[loff[4 * p] + k]; w0[k] = r0 * mk(l.x, l.y) + mk(a.x, a.y); }
        for (int k = 0; k < n1; k++) { float2 l = lverts[loff[4 * p + 2] + k]; w1[k] = r1 * mk(l.x, l.y) + mk(b.x, b.y); }
        bool ovf = false;
        m = poly_poly<MAX_POLY_VERTS>(w0, n0, w1, n1, o, &ovf);
    }
    counts[p] = m;
    for (int k = 0; k < m && k < cap; k++) {
        SyltContact c;
        memset(&c, 0, sizeof c);
        c.px = o[k].px; c.py = o[k].py; c.nx = o[k].nx; c.ny = o[k].ny; c.separation = o[k].sep;
        c.in_edge_1 = o[k].feat & 0xff; c.out_edge_1 = (o[k].feat >> 8) & 0xff;
        c.in_edge_2 = (o[k].feat >> 16) & 0xff; c.out_edge_2 = (o[k].feat >> 24) & 0xff;
        out[(size_t)p * cap + k] = c;
    }
}

__global__ void k_sincos(int n, const float* a, float* s, float* c) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) { c[i] = sylt_cosf(a[i]); s[i] = sylt_sinf(a[i]); }
}

}  // namespace

// ====================================================================== host side
struct HostJoint {
    int b1, b2;
    float la1x, la1y, la2x, la2y, softness, bias_factor;
    int color;
};

struct SyltWorld {
    int device = 0;
    cudaStream_t stream = nullptr;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    float gx = 0, gy = 0;
    uint32_t iterations = 10;
    int accumulate = 1, warm = 0, poscorr = 1;  // World::new defaults (world.rs:38-42, quirk Q-4)
    int fix_features = 0, fix_inverse = 0;      // opt-in corrected physics, default off (= the reference's behaviour)
    int64_t launches = 0;
    int num_sms = 0, coop_blocks = 0, smem_slots = 0, smem_contacts = 0;
    size_t solve_smem = 0;

    // host mirror of construction-time data
    std::vector<HostBody> hb;
    std::vector<V2> hverts;
    std::vector<SyltBodyState> pending;  // initial states of bodies not yet uploaded
    std::vector<HostJoint> hj;
    int uploaded = 0;        // bodies already on the device
    bool class_dirty = true; // grid classification / flags need recomputing
    bool joints_dirty = true;
    bool jparams_dirty = false;  // only softness / bias_factor / local anchors changed: re-upload, no re-colouring
    bool any_poly = false;
    int max_poly_verts = 0;  // largest polygon of the world: picks the tier (8 / 16 / 32 vertices) of the narrow phase
    int n_large = 0;
    GridParams gp{1.0f, 1023, 1.0f};
    int buckets = 1024;

    // device state
    DBuf<float4> pose, vel, force, mp, geom, aabb, pose_prev;  // pose_prev: the poses the last region-solver step solved with
    DBuf<float2> rotc, lverts;
    DBuf<int> bflags, bflags_old, rank, cell_count, cell_start, cell_local, tile_tot, large, large_start, row_count, row_start, row_tmp;
    int n_groups = 1;        // collision groups in use (SyltBodyDesc.group)
    DBuf<int4> cinfo, cent_info;  // cent_info: interleaved (info, AABB) entries of the cell-ordered copy
    DBuf<int2> pairs;
    DBuf<unsigned long long> info, pscan, claim, pair_btot;
    DBuf<int> row_btot;
    DBuf<float4> tmp_a;
    DBuf<float2> tmp_b;
    DBuf<unsigned long long> used;
    DBuf<unsigned> used_int;
    DBuf<ulonglong2> vrec;
    int sleep_ns = 0, ctas_per_sm = 1, solve_threads = 512, slack_pct = 0;
    bool timing = false;
    DBuf<unsigned long long> dbg;
    DBuf<int> jdeg;
    DBuf<int2> jrank;
    DBuf<int> uncolored, uncolored2, prop, order;
    DBuf<int4> unc_rec, unc_bin, srec;   // srec: 2 per arbiter, cin: 2 per contact (region solver set-up records, k_build)
    DBuf<int> unc_bin_idx;
    DBuf<Counters> counters;
    bool cells_clean = false;  // cell_count is all zero (the cell scan re-zeroes what it reads)
    // arbiters, double buffered
    struct ArbBufs {
        DBuf<int2> bodies, meta;
        DBuf<int2> partner;
        DBuf<int> color, row_start;
        DBuf<float> friction;
        DBuf<float4> ca, cb, cc, cd, cin;
        bool sep_in_cin = false;  // the set was solved by the region solver: r1, r2, masses and the separation are filled in on demand
        float lazy_inv_dt = 0.0f, lazy_kbias = 0.0f;
    } arb[2];
    int cur = 0;             // index being built next
    bool have_old = false;   // an arbiter set from a previous step/broad phase exists
    int B_old = 0;
    bool flags_stale = false, inflight = false, inflight_full = false;
    // joints
    DBuf<int2> jbodies;
    DBuf<float4> jla, jr, jm;
    DBuf<float2> jparam, jp, jbias;
    DBuf<int> jorder, jcolor_start;
    int n_joint_colors = 0;
    std::vector<int> h_jorder;

    // regions (k_solve_regions): spatial partition of the bodies, recomputed on the host when bodies were added or the
    // share of cross-region arbiters has grown
    DBuf<int> region_of, lidx, reg_start, reg_bodies, sort_keys;
    DBuf<RecutScratch> recut;
    DBuf<int2> gadj;
    DBuf<ulonglong2> xrec;
    unsigned step_tag = 0, step_serial = 0;
    int regions_mode = 1;            // SYLT_REGIONS: 0 = never (k_solve only), 1 = joint-free worlds of at least regions_min_bodies (default), 2 = always
    int regions_min_bodies = 1000, regions_count = 0, regions_period = 128;
    int regions_mode_saved = 1, regions_suspended_at = 0;  // region solver suspended (a region did not fit) at this many contacts
    bool color_local = true;         // SYLT_COLOR_LOCAL=0: new arbiters are coloured by the grid-wide rounds only (diagnostics)
    bool regions_on = false;         // the partition on the device is valid for the current bodies
    bool part_dirty = true;          // bodies were added / the mode may have changed
    bool last_regions = false;       // the last enqueued step ran k_solve_regions
    bool recolor_all = false;        // next step: ignore last step's colours (they encode the old partition)
    int n_regions = 0;
    int64_t steps_since_part = 0;
    bool resort_pending = false;     // the partition of the current bodies has not been computed yet
    cudaEvent_t kt_ev[16] = {};      // SYLT_KERNEL_TIMING / sylt_world_set_kernel_timing
    int kt_n = 0;
    bool kernel_timing = false, kernel_timing_print = false;
    void* rb_host = nullptr;         // pinned staging of sylt_world_get_arbiters
    size_t rb_cap = 0;
    void* sb_host = nullptr;         // ... and of sylt_world_get_bodies (packed on the device first)
    size_t sb_cap = 0;
    DBuf<float> pack_dev;
    size_t solve_budget = 0;         // dynamic shared memory of one k_solve_regions CTA
    int regions_per_sm = 1, regions_blocks = 0, regions_threads = SOLVE_THREADS;  // SYLT_REGION_CTAS_PER_SM

    int64_t cap_pairs = 0, cap_arb = 0, cap_con = 0;
    int64_t user_pairs = 0, user_arb = 0, user_con = 0;
    int64_t grow_pairs = 0, grow_arb = 0, grow_con = 0;  // raised automatically when a step overflowed (the step is then repeated)
    // the steps enqueued since the last finish(): what to repeat when one of them fails for lack of capacity
    struct Segment { float dt; int n; bool full; };
    std::vector<Segment> segs;
    int call_cur0 = 0, call_B_old0 = 0;
    bool call_have_old0 = false, call_flags_stale0 = false, call_pending0 = false, call_recolor0 = false;
    int64_t call_since0 = 0;
    int n_arb_valid = 0, n_con_valid = 0;  // size of the arbiter set sylt_world_get_arbiters reads (the last set that was completed)
    unsigned epoch = 1;
    Counters last{};         // counters of the last completed step
    bool have_last = false;
    int sticky_err = 0;
    uint32_t last_iterations = 0;
    bool last_was_step = false;

    ArbSet set(int k) {
        ArbSet s;
        ArbBufs& b = arb[k];
        s.bodies = b.bodies.p; s.meta = b.meta.p; s.partner = b.partner.p; s.friction = b.friction.p; s.color = b.color.p;
        s.row_start = b.row_start.p; s.ca = b.ca.p; s.cb = b.cb.p; s.cc = b.cc.p; s.cd = b.cd.p; s.cin = b.cin.p;
        return s;
    }
};

namespace {

// contacts a candidate pair may have: 2 (boxes only) or twice the tier (8 / 16 / 32 vertices) of the world's largest polygon
int manifold_tier(const SyltWorld* w) { return !w->any_poly ? 2 : w->max_poly_verts <= 8 ? 16 : w->max_poly_verts <= 16 ? 32 : 64; }
int use_device(SyltWorld* w) {
    SYLT_CUDA(cudaSetDevice(w->device));
    return SYLT_OK;
}

// Host-side greedy colouring of the (static) joint graph, in joint index order.
void color_joints(SyltWorld* w) {
    size_t J = w->hj.size();
    std::vector<uint64_t> used(w->hb.size(), 0);
    int ncol = 0;
    for (size_t j = 0; j < J; j++) {
        HostJoint& q = w->hj[j];
        bool d1 = w->hb[(size_t)q.b1].inv_mass != 0.0f || w->hb[(size_t)q.b1].inv_moi != 0.0f;
        bool d2 = w->hb[(size_t)q.b2].inv_mass != 0.0f || w->hb[(size_t)q.b2].inv_moi != 0.0f;
        uint64_t m = (d1 ? used[(size_t)q.b1] : 0) | (d2 ? used[(size_t)q.b2] : 0);
        int c = 0;
        while (c < 63 && (m >> c) & 1) c++;
        q.color = c;
        if (d1) used[(size_t)q.b1] |= 1ull << c;
        if (d2) used[(size_t)q.b2] |= 1ull << c;
        ncol = std::max(ncol, c + 1);
    }
    w->n_joint_colors = J ? ncol : 0;
    std::vector<int> start((size_t)w->n_joint_colors + 1, 0);
    for (auto& q : w->hj) start[(size_t)q.color + 1]++;
    for (int c = 0; c < w->n_joint_colors; c++) start[(size_t)c + 1] += start[(size_t)c];
    w->h_jorder.assign(J, 0);
    std::vector<int> fill(start.begin(), start.end());
    for (size_t j = 0; j < J; j++) w->h_jorder[(size_t)fill[(size_t)w->hj[j].color]++] = (int)j;
    // per joint: rank among the joints of each of its bodies in solve order; per body: number of joints (dataflow solver)
    std::vector<int> jdeg(w->hb.size() + 1, 0);
    std::vector<int2> jrank(J + 1, make_int2(0, 0));
    for (size_t k = 0; k < J; k++) {
        int j = w->h_jorder[k];
        const HostJoint& q = w->hj[(size_t)j];
        jrank[(size_t)j] = make_int2(jdeg[(size_t)q.b1], jdeg[(size_t)q.b2] + (q.b1 == q.b2 ? 1 : 0));
        jdeg[(size_t)q.b1]++;
        jdeg[(size_t)q.b2]++;
    }
    w->jdeg.ensure(w->hb.size() + 1);
    w->jrank.ensure(J + 1);
    cudaMemcpyAsync(w->jdeg.p, jdeg.data(), jdeg.size() * sizeof(int), cudaMemcpyHostToDevice, w->stream);
    cudaMemcpyAsync(w->jrank.p, jrank.data(), jrank.size() * sizeof(int2), cudaMemcpyHostToDevice, w->stream);
    // upload
    w->jorder.ensure(J + 1);
    w->jcolor_start.ensure((size_t)w->n_joint_colors + 2);
    if (J) cudaMemcpyAsync(w->jorder.p, w->h_jorder.data(), J * sizeof(int), cudaMemcpyHostToDevice, w->stream);
    cudaMemcpyAsync(w->jcolor_start.p, start.data(), start.size() * sizeof(int), cudaMemcpyHostToDevice, w->stream);
    cudaStreamSynchronize(w->stream);
}

// Upload newly added bodies / joints and (re)compute the static grid classification.
int finish(SyltWorld* w);
int flush(SyltWorld* w) {
    int e = use_device(w);
    if (e) return e;
    if (w->inflight) finish(w);  // buffers may be reallocated below: drain the queue and read the counters first
    const size_t B = w->hb.size();
    if ((size_t)w->uploaded < B) {
        const size_t keep = (size_t)w->uploaded, add = B - keep;
        SYLT_CUDA(w->pose.ensure(B, keep, w->stream));
        SYLT_CUDA(w->vel.ensure(B, keep, w->stream));
        SYLT_CUDA(w->force.ensure(B, keep, w->stream));
        SYLT_CUDA(w->pose_prev.ensure(B, keep, w->stream));
        SYLT_CUDA(w->mp.ensure(B, keep, w->stream));
        SYLT_CUDA(w->geom.ensure(B, keep, w->stream));
        SYLT_CUDA(w->bflags_old.ensure(B, keep, w->stream));
        SYLT_CUDA(w->bflags.ensure(B, keep, w->stream));
        SYLT_CUDA(w->aabb.ensure(B));
        SYLT_CUDA(w->rotc.ensure(B));
        SYLT_CUDA(w->rank.ensure(B));
        SYLT_CUDA(w->cinfo.ensure(B));
        SYLT_CUDA(w->cent_info.ensure(2 * B));
        SYLT_CUDA(w->row_count.ensure(B + 1));
        SYLT_CUDA(w->row_btot.ensure(B / ROW_BLK + 2));
        SYLT_CUDA(w->row_start.ensure(B + 2));
        SYLT_CUDA(w->row_tmp.ensure(B * ROWTMP));
        SYLT_CUDA(w->used.ensure(B));
        SYLT_CUDA(w->used_int.ensure(B));
        SYLT_CUDA(w->vrec.ensure(2 * B));
        w->joints_dirty = true;  // jdeg is per body
        SYLT_CUDA(w->large.ensure(B));
        for (int k = 0; k < 2; k++) SYLT_CUDA(w->arb[k].row_start.ensure(B + 2, (size_t)w->B_old + 1, w->stream));
        {   // claims: all ones == "no claim"
            size_t oldcap = w->claim.cap;
            SYLT_CUDA(w->claim.ensure(B * NCX));
            if (w->claim.cap != oldcap) SYLT_CUDA(cudaMemsetAsync(w->claim.p, 0xff, w->claim.cap * sizeof(unsigned long long), w->stream));
        }
        std::vector<float4> hp(add), hv(add), hf(add), hm(add), hg(add);
        for (size_t k = 0; k < add; k++) {
            const HostBody& b = w->hb[keep + k];
            const SyltBodyState& s = w->pending[k];
            hp[k] = make_float4(s.x, s.y, s.rotation, 0.0f);
            hv[k] = make_float4(s.vx, s.vy, s.angular_velocity, 0.0f);
            hf[k] = make_float4(s.fx, s.fy, s.torque, 0.0f);
            hm[k] = make_float4(b.inv_mass, b.inv_moi, b.friction, 0.0f);
            float off;
            int o = b.vert_offset;
            memcpy(&off, &o, 4);
            hg[k] = make_float4(b.width_x, b.width_y, off, b.radius);
        }
        SYLT_CUDA(cudaMemcpyAsync(w->pose.p + keep, hp.data(), add * sizeof(float4), cudaMemcpyHostToDevice, w->stream));
        SYLT_CUDA(cudaMemcpyAsync(w->vel.p + keep, hv.data(), add * sizeof(float4), cudaMemcpyHostToDevice, w->stream));
        SYLT_CUDA(cudaMemcpyAsync(w->force.p + keep, hf.data(), add * sizeof(float4), cudaMemcpyHostToDevice, w->stream));
        SYLT_CUDA(cudaMemcpyAsync(w->mp.p + keep, hm.data(), add * sizeof(float4), cudaMemcpyHostToDevice, w->stream));
        SYLT_CUDA(cudaMemcpyAsync(w->geom.p + keep, hg.data(), add * sizeof(float4), cudaMemcpyHostToDevice, w->stream));
        SYLT_CUDA(w->lverts.ensure(w->hverts.size() + 1));
        if (!w->hverts.empty())
            SYLT_CUDA(cudaMemcpyAsync(w->lverts.p, w->hverts.data(), w->hverts.size() * sizeof(float2), cudaMemcpyHostToDevice, w->stream));
        SYLT_CUDA(cudaStreamSynchronize(w->stream));
        w->pending.clear();
        w->uploaded = (int)B;
        w->class_dirty = true;
        w->part_dirty = true;
    }
    if (w->class_dirty && B > 0) {
        // static classification: cell = slightly more than the largest "small" bounding diameter; bodies above go to the large list
        std::vector<float> d;
        for (const HostBody& b : w->hb)
            if (b.inv_mass != 0.0f) d.push_back(2.0f * b.radius);
        float cell = 1.0f;
        if (!d.empty()) {
            std::vector<float> t = d;
            std::nth_element(t.begin(), t.begin() + (long)(t.size() / 2), t.end());
            float med = t[t.size() / 2], mx = 0.0f;
            for (float x : d)
                if (x <= 4.0f * med && x > mx) mx = x;
            if (mx > 0.0f) cell = mx;
        }
        cell = 1.1f * cell + 0.1f;
        std::vector<int> fl(B), large;
        int n_groups = 1;
        for (size_t i = 0; i < B; i++) {
            const HostBody& b = w->hb[i];
            int f = (b.shape == SYLT_SHAPE_POLYGON ? FLAG_POLY : 0) | (b.inv_mass == 0.0f ? FLAG_STATIC : 0) | (b.nverts << 8) | (b.group << GROUP_SHIFT);
            if (2.2f * b.radius + 0.1f > cell || !(b.radius == b.radius)) { f |= FLAG_LARGE; large.push_back((int)i); }
            fl[i] = f;
            n_groups = std::max(n_groups, b.group + 1);
        }
        w->n_large = (int)large.size();
        w->n_groups = n_groups;
        // the large bodies (ground, walls) bypass the grid: every body tests the ones of its own group — a CSR by group, index order inside
        std::stable_sort(large.begin(), large.end(), [&](int x, int y) { return w->hb[(size_t)x].group < w->hb[(size_t)y].group; });
        std::vector<int> lstart((size_t)n_groups + 1, 0);
        for (int j : large) lstart[(size_t)w->hb[(size_t)j].group + 1]++;
        for (int g = 0; g < n_groups; g++) lstart[(size_t)g + 1] += lstart[(size_t)g];
        SYLT_CUDA(w->large_start.ensure((size_t)n_groups + 2));
        SYLT_CUDA(cudaMemcpyAsync(w->large_start.p, lstart.data(), lstart.size() * sizeof(int), cudaMemcpyHostToDevice, w->stream));
        int buckets = 2048;
        while ((size_t)buckets < 2 * B && buckets < CELL_TILE * MAX_CELL_TILES) buckets <<= 1;  // (beyond 2 M bodies the buckets fill up: still correct, slower)
        w->buckets = buckets;
        w->gp.cell = cell;
        w->gp.inv_cell = 1.0f / cell;
        w->gp.mask = buckets - 1;
        if ((size_t)buckets + 1 > w->cell_count.cap) w->cells_clean = false;
        SYLT_CUDA(w->cell_count.ensure((size_t)buckets + 1));
        SYLT_CUDA(w->cell_start.ensure((size_t)buckets + 2));
        SYLT_CUDA(w->cell_local.ensure((size_t)buckets + 8));
        SYLT_CUDA(w->tile_tot.ensure((size_t)buckets / CELL_TILE + 2));
        SYLT_CUDA(w->counters.ensure(1));
        if (w->have_old && w->B_old > 0 && !w->flags_stale)  // keep the flags the last step used for the next arbiter match
            SYLT_CUDA(cudaMemcpyAsync(w->bflags_old.p, w->bflags.p, (size_t)w->B_old * sizeof(int), cudaMemcpyDeviceToDevice, w->stream));  // (flags_stale: bflags_old already holds them)
        SYLT_CUDA(cudaMemcpyAsync(w->bflags.p, fl.data(), B * sizeof(int), cudaMemcpyHostToDevice, w->stream));
        w->flags_stale = true;
        if (!large.empty()) SYLT_CUDA(cudaMemcpyAsync(w->large.p, large.data(), large.size() * sizeof(int), cudaMemcpyHostToDevice, w->stream));
        // capacities
        int64_t cp = w->user_pairs ? w->user_pairs : std::max<int64_t>(16 * (int64_t)B, 4096);
        int64_t ca = w->user_arb ? w->user_arb : std::max<int64_t>(8 * (int64_t)B, 2048);
        int64_t cc = w->user_con ? w->user_con : std::max<int64_t>(16 * (int64_t)B, 4096);
        if (w->any_poly && !w->user_con) cc *= 2;
        cp = std::max(cp, w->grow_pairs); ca = std::max(ca, w->grow_arb); cc = std::max(cc, w->grow_con);  // a step overflowed before
        const int maxc = manifold_tier(w);
        SYLT_CUDA(w->pairs.ensure((size_t)cp));
        SYLT_CUDA(w->info.ensure((size_t)cp + 1));
        SYLT_CUDA(w->pscan.ensure((size_t)cp + 2));
        SYLT_CUDA(w->pair_btot.ensure((size_t)cp / PAIR_TILE + 2));
        SYLT_CUDA(w->tmp_a.ensure((size_t)cp * maxc));
        SYLT_CUDA(w->tmp_b.ensure((size_t)cp * maxc));
        SYLT_CUDA(w->uncolored.ensure((size_t)ca));
        SYLT_CUDA(w->unc_rec.ensure((size_t)ca));
        SYLT_CUDA(w->uncolored2.ensure((size_t)ca));
        SYLT_CUDA(w->prop.ensure((size_t)ca));
        SYLT_CUDA(w->order.ensure((size_t)ca));
        SYLT_CUDA(w->srec.ensure(2 * (size_t)ca));
        for (int k = 0; k < 2; k++) {
            SyltWorld::ArbBufs& a = w->arb[k];
            const bool keep_set = (k != w->cur) && w->have_old && w->have_last;  // last step's arbiters survive a regrow
            size_t ka = keep_set ? (size_t)std::min<int64_t>(w->n_arb_valid, w->cap_arb) : 0;
            size_t kc = keep_set ? (size_t)std::min<int64_t>(w->n_con_valid, w->cap_con) : 0;
            SYLT_CUDA(a.bodies.ensure((size_t)ca, ka, w->stream));
            SYLT_CUDA(a.meta.ensure((size_t)ca, ka, w->stream));
            SYLT_CUDA(a.partner.ensure((size_t)ca, ka, w->stream));
            SYLT_CUDA(a.color.ensure((size_t)ca, ka, w->stream));
            SYLT_CUDA(a.friction.ensure((size_t)ca, ka, w->stream));
            SYLT_CUDA(a.ca.ensure((size_t)cc, kc, w->stream));
            SYLT_CUDA(a.cb.ensure((size_t)cc, kc, w->stream));
            SYLT_CUDA(a.cc.ensure((size_t)cc, kc, w->stream));
            SYLT_CUDA(a.cd.ensure((size_t)cc, kc, w->stream));
            SYLT_CUDA(a.cin.ensure(2 * (size_t)cc, 2 * kc, w->stream));
        }
        w->cap_pairs = (int64_t)std::min<size_t>(w->pairs.cap, (size_t)cp);
        w->cap_arb = ca;
        w->cap_con = cc;
        SYLT_CUDA(cudaStreamSynchronize(w->stream));
        w->class_dirty = false;
    }
    if (w->joints_dirty) {
        size_t J = w->hj.size();
        SYLT_CUDA(w->jbodies.ensure(J + 1));
        SYLT_CUDA(w->jla.ensure(J + 1));
        SYLT_CUDA(w->jparam.ensure(J + 1));
        SYLT_CUDA(w->jr.ensure(J + 1));
        SYLT_CUDA(w->jm.ensure(J + 1));
        SYLT_CUDA(w->jbias.ensure(J + 1));
        std::vector<int2> jb(J);
        std::vector<float4> la(J);
        std::vector<float2> pr(J);
        for (size_t j = 0; j < J; j++) {
            jb[j] = make_int2(w->hj[j].b1, w->hj[j].b2);
            la[j] = make_float4(w->hj[j].la1x, w->hj[j].la1y, w->hj[j].la2x, w->hj[j].la2y);
            pr[j] = make_float2(w->hj[j].softness, w->hj[j].bias_factor);
        }
        if (J) {
            SYLT_CUDA(cudaMemcpyAsync(w->jbodies.p, jb.data(), J * sizeof(int2), cudaMemcpyHostToDevice, w->stream));
            SYLT_CUDA(cudaMemcpyAsync(w->jla.p, la.data(), J * sizeof(float4), cudaMemcpyHostToDevice, w->stream));
            SYLT_CUDA(cudaMemcpyAsync(w->jparam.p, pr.data(), J * sizeof(float2), cudaMemcpyHostToDevice, w->stream));
        }
        color_joints(w);
        SYLT_CUDA(cudaStreamSynchronize(w->stream));
        w->joints_dirty = false;
        w->jparams_dirty = false;
    }
    if (w->jparams_dirty) {  // pub Joint fields edited by the caller: same joint graph, new numbers
        const size_t J = w->hj.size();
        std::vector<float4> la(J);
        std::vector<float2> pr(J);
        for (size_t j = 0; j < J; j++) {
            la[j] = make_float4(w->hj[j].la1x, w->hj[j].la1y, w->hj[j].la2x, w->hj[j].la2y);
            pr[j] = make_float2(w->hj[j].softness, w->hj[j].bias_factor);
        }
        if (J) {
            SYLT_CUDA(cudaMemcpyAsync(w->jla.p, la.data(), J * sizeof(float4), cudaMemcpyHostToDevice, w->stream));
            SYLT_CUDA(cudaMemcpyAsync(w->jparam.p, pr.data(), J * sizeof(float2), cudaMemcpyHostToDevice, w->stream));
            SYLT_CUDA(cudaStreamSynchronize(w->stream));
        }
        w->jparams_dirty = false;
    }
    return SYLT_OK;
}

int unc_bin_cap(const SyltWorld* w) { return UNC_BIN_CAP_1 / std::max(w->regions_per_sm, 1); }

// Region solver on / off for the current bodies, and its buffers.  The partition itself (k_recut) is enqueued
// with the next step; nothing here synchronises with the device beyond (re)allocation.
int repartition(SyltWorld* w) {
    const size_t B = w->hb.size();
    w->part_dirty = false;
    const bool was_on = w->regions_on;
    const int G = std::min(1024, w->regions_count > 0 ? std::min(w->regions_count, w->regions_blocks) : w->regions_blocks);
    bool on = w->regions_mode == 2 || (w->regions_mode == 1 && B >= (size_t)w->regions_min_bodies);
    if (B == 0 || G < 1 || !w->hj.empty() || w->iterations + 1 > (uint32_t)MAX_PASSES) on = false;  // joints: k_solve
    // a region may hold twice the average number of bodies (k_recut floors the weights at the average); 40 bytes each must leave room for the set-up scratch and the slots
    const size_t maxper = 2 * ((B + (size_t)G - 1) / (size_t)std::max(G, 1)) + 16;
    if (on && maxper * 40 + ((size_t)3 * HASH_CAP_1 * 4 + (size_t)GB_MAX_1 * 4 + (size_t)OWN_MAX_1 * 8 + (size_t)CONE_CAP_1 * 8 + 32768) / (size_t)w->regions_per_sm > w->solve_budget) on = false;
    if (on != was_on) w->recolor_all = true;  // the colour ranges mean something else
    w->regions_on = on;
    if (!on) return SYLT_OK;
    if (w->inflight) finish(w);  // buffers may move
    SYLT_CUDA(w->region_of.ensure(B));
    SYLT_CUDA(w->lidx.ensure(B));
    SYLT_CUDA(w->reg_bodies.ensure(B));
    SYLT_CUDA(w->sort_keys.ensure(B));
    SYLT_CUDA(w->recut.ensure(1));
    SYLT_CUDA(w->reg_start.ensure((size_t)G + 1));
    SYLT_CUDA(w->gadj.ensure(B * (size_t)GADJ));
    SYLT_CUDA(w->unc_bin.ensure((size_t)(G + 1) * (size_t)unc_bin_cap(w)));
    SYLT_CUDA(w->unc_bin_idx.ensure((size_t)(G + 1) * (size_t)unc_bin_cap(w)));
    {   // export planes: a record is valid when its tag matches (step_tag, pass); fresh memory must not hold a look-alike
        const size_t oldcap = w->xrec.cap;
        SYLT_CUDA(w->xrec.ensure((size_t)MAX_PLANES * 2 * B));
        if (w->xrec.cap != oldcap) SYLT_CUDA(cudaMemsetAsync(w->xrec.p, 0, w->xrec.cap * sizeof(ulonglong2), w->stream));
    }
    w->n_regions = G;
    w->resort_pending = true;
    return SYLT_OK;
}

// The re-cut schedule of the region solver is host-side bookkeeping that runs ahead of the device (steps are enqueued, not
// awaited): recut_due + the updates in enqueue_step.  replay_schedule() re-derives it for the steps of a call that completed
// when a later step failed, so that the repeated steps re-cut (and recolour) exactly where the original ones would have.
bool recut_due(const SyltWorld* w, bool full_step) {
    return w->regions_on && (w->resort_pending || (full_step && w->steps_since_part >= w->regions_period));
}
void replay_schedule(SyltWorld* w, int64_t since0, bool pending0, bool recolor0, const std::vector<SyltWorld::Segment>& segs, int completed) {
    w->steps_since_part = since0; w->resort_pending = pending0; w->recolor_all = recolor0;
    for (const auto& sg : segs)
        for (int k = 0; k < sg.n && completed > 0; k++, completed--) {
            if (recut_due(w, sg.full)) { w->steps_since_part = 0; w->resort_pending = false; }
            w->recolor_all = false;
            if (w->regions_on) w->steps_since_part++;
        }
}

// Enqueue one step (or one broad phase only) on the stream.  No host synchronisation.
int enqueue_step(SyltWorld* w, float dt, bool full_step) {
    const int B = (int)w->hb.size();
    cudaStream_t st = w->stream;
    // SYLT_KERNEL_TIMING=1: events between the launches of every step (diagnostics; it serialises the launch chain)
    const bool ktime = w->kernel_timing;
    int kt_n = 0;
    auto mark = [&]() { if (ktime && kt_n < 16) { if (!w->kt_ev[kt_n]) cudaEventCreate(&w->kt_ev[kt_n]); cudaEventRecord(w->kt_ev[kt_n++], st); } };
    mark();
    Counters* cn = w->counters.p;
    if (++w->step_serial == 0u) w->step_serial = 1u;
    static_assert((sizeof(Counters) - offsetof(Counters, n_pairs)) % sizeof(int) == 0, "Counters: per-step part is a whole number of ints");
    if (B == 0) {  // an empty world: nothing runs, nothing is counted
        SYLT_CUDA(cudaMemsetAsync(&cn->n_pairs, 0, sizeof(Counters) - offsetof(Counters, n_pairs), st));
        return SYLT_OK;
    }
    if (!w->cells_clean) {
        SYLT_CUDA(cudaMemsetAsync(w->cell_count.p, 0, w->cell_count.cap * sizeof(int), st));
        w->cells_clean = true;
    }
    const int nbB = (B + 255) / 256;
    k_prepare<<<nbB, 256, 0, st>>>(B, w->pose.p, w->vel.p, w->force.p, w->mp.p, w->geom.p, w->bflags.p, w->lverts.p, w->rotc.p, w->aabb.p,
                                   w->cinfo.p, w->cell_count.p, w->rank.p, w->gp, w->gx, w->gy, dt, full_step ? 1 : 0, cn, w->vrec.p, w->used.p, w->used_int.p, w->row_btot.p, w->step_serial);
    w->launches++;
    mark();
    const int ntiles = (w->buckets + CELL_TILE - 1) / CELL_TILE;
    SYLT_CUDA(launch_pdl(k_cells_local, dim3((unsigned)ntiles), dim3(256), 0, st, w->buckets, w->cell_count.p, w->cell_local.p, w->tile_tot.p));
    w->launches++;
    mark();
    SYLT_CUDA(launch_pdl(k_fill_cells, dim3((unsigned)nbB), dim3(256), 0, st, B, w->cinfo.p, w->aabb.p, w->rank.p, w->cell_local.p, w->tile_tot.p, ntiles, w->buckets,
                         w->cell_start.p, w->cent_info.p, (float4*)(w->cent_info.p + 1), w->gp, w->large.p, w->n_large));
    mark();
    const int nbR = (B + ROW_BLK - 1) / ROW_BLK;
    static const int rows_lpb = getenv("SYLT_ROWS_LPB") ? atoi(getenv("SYLT_ROWS_LPB")) : 16;
    SYLT_CUDA(launch_pdl(rows_lpb == 32 ? k_rows_count<32> : k_rows_count<16>, dim3((unsigned)(rows_lpb == 32 ? (B + 3) / 4 : (B + 7) / 8)), dim3(128), 0, st, B, w->aabb.p, w->cinfo.p, w->cell_start.p, w->cent_info.p,
                         (float4*)(w->cent_info.p + 1), w->large.p, w->n_large, w->large_start.p, w->gp, w->row_count.p, w->row_tmp.p, cn, w->row_btot.p));
    w->launches += 2;
    mark();
    SYLT_CUDA(launch_pdl(k_rows_emit, dim3((unsigned)nbR), dim3(ROW_BLK), 0, st, B, w->aabb.p, w->cinfo.p, w->cell_start.p, w->cent_info.p, (float4*)(w->cent_info.p + 1),
                         w->large.p, w->large_start.p, w->gp, w->row_start.p, w->row_tmp.p, w->pairs.p, (int)w->cap_pairs, cn, w->row_count.p, w->row_btot.p, B - w->n_large));
    w->launches += 1;
    mark();
    const int gridP = w->num_sms * 8;
    const int maxc = manifold_tier(w);  // contacts a pair may have: 2 (boxes only) or twice the largest polygon's tier (8 / 16 / 32 vertices)
    {
        auto kn = maxc == 2 ? k_narrow<2> : maxc == 16 ? k_narrow<16> : maxc == 32 ? k_narrow<32> : k_narrow<64>;
        SYLT_CUDA(launch_pdl(kn, dim3((unsigned)gridP), dim3(128), 0, st, cn, (int)w->cap_pairs, w->pairs.p, w->pose.p, w->rotc.p, w->geom.p, w->bflags.p,
                             w->lverts.p, w->info.p, w->tmp_a.p, w->tmp_b.p, cn, w->pair_btot.p));
    }
    w->launches++;
    mark();
    ArbSet cur = w->set(w->cur), old = w->set(w->cur ^ 1);
    bool recut_now = false;
    if (recut_due(w, full_step)) {
        // (re)partition by the current positions and last step's load (device side, no synchronisation)
        SYLT_CUDA(cudaMemsetAsync(w->recut.p, 0, sizeof(RecutScratch), st));
        {
            int rB = B, rG = w->n_regions, rng = w->n_groups, rBold = w->B_old;
            const float4* rpose = w->pose.p;
            const int *rfl = w->bflags.p, *rold = w->have_old ? old.row_start : nullptr;
            int *rperm = w->reg_bodies.p, *rkey = w->sort_keys.p, *rreg = w->region_of.p, *rlidx = w->lidx.p, *rstart = w->reg_start.p;
            RecutScratch* rsc = w->recut.p;
            const Counters* rcn = cn;
            void* rargs[] = {&rB, &rG, &rpose, &rfl, &rng, &rold, &rBold, &rperm, &rkey, &rreg, &rlidx, &rstart, &rsc, &rcn};
            SYLT_CUDA(cudaLaunchCooperativeKernel((void*)k_recut, dim3((unsigned)w->num_sms), dim3(RECUT_THREADS), rargs, (SORT_CELLS + 1) * sizeof(int), st));
        }
        w->launches += 1;
        w->steps_since_part = 0;
        w->resort_pending = false;
        recut_now = true;
    }
    const int keep_colors = w->recolor_all ? 0 : (recut_now ? 2 : 1);
    w->recolor_all = false;
    static const int build_mult = getenv("SYLT_BUILD_MULT") ? atoi(getenv("SYLT_BUILD_MULT")) : 4;  // CTAs per SM of k_build (8 measured slower: 32.8 vs 26.7 us at 100k bodies)
    {
        auto kb = maxc == 2 ? k_build<2> : maxc == 16 ? k_build<16> : maxc == 32 ? k_build<32> : k_build<64>;
        SYLT_CUDA(launch_pdl(kb, dim3((unsigned)(w->num_sms * build_mult)), dim3(256), 0, st, cn, (int)w->cap_pairs, (int)w->cap_arb, (int)w->cap_con, B, w->B_old,
                             w->have_old ? 1 : 0, w->pairs.p, w->info.p, w->pscan.p, w->pair_btot.p, w->row_start.p, w->tmp_a.p, w->tmp_b.p, w->mp.p, w->bflags.p, w->bflags_old.p, cur, old,
                             w->warm, w->used.p, w->unc_rec.p, w->fix_features, keep_colors, w->used_int.p, w->regions_on ? (full_step ? 2 : 1) : 0, w->lidx.p, w->gadj.p, w->n_regions,
                             w->unc_bin.p, w->unc_bin_idx.p, unc_bin_cap(w), w->srec.p));
    }
    w->launches += 1;
    SYLT_CUDA(cudaGetLastError());
    mark();

    SolveArgs a;
    memset(&a, 0, sizeof a);
    a.cn = cn; a.B = B; a.cap_arb = (int)w->cap_arb; a.cap_con = (int)w->cap_con;
    a.J = (int)w->hj.size(); a.n_joint_colors = w->n_joint_colors; a.iterations = (int)w->iterations;
    a.accumulate = w->accumulate; a.warm = w->warm; a.poscorr = w->poscorr; a.run_solver = full_step ? 1 : 0;
    a.dt = dt; a.inv_dt = dt > 0.0f ? 1.0f / dt : 0.0f;  // world.rs:108
    a.epoch_base = w->epoch;
    w->epoch += MAX_ROUNDS;
    if (w->epoch > 0xf0000000u) {  // wrap: forget all claims
        SYLT_CUDA(cudaMemsetAsync(w->claim.p, 0xff, w->claim.cap * sizeof(unsigned long long), st));
        w->epoch = 1;
    }
    a.pose = w->pose.p; a.vel = w->vel.p; a.force = w->force.p; a.pose_prev = w->pose_prev.p; a.mp = w->mp.p; a.rotc = w->rotc.p; a.bflags = w->bflags.p;
    a.cur = cur; a.used = w->used.p; a.used_int = w->used_int.p; a.claim = w->claim.p; a.uncolored = w->uncolored.p; a.uncolored2 = w->uncolored2.p; a.unc_rec = w->unc_rec.p; a.prop = w->prop.p; a.order = w->order.p;
    a.jbodies = w->jbodies.p; a.jla = w->jla.p; a.jparam = w->jparam.p; a.jp = w->jp.p; a.jr = w->jr.p; a.jm = w->jm.p; a.jbias = w->jbias.p;
    a.jorder = w->jorder.p; a.jcolor_start = w->jcolor_start.p;
    void* args[] = {&a};
    a.smem_slots = w->smem_slots; a.smem_contacts = w->smem_contacts;
    a.cand_row_start = w->row_start.p; a.pscan = w->pscan.p; a.cap_pairs = (int)w->cap_pairs;
    a.vrec = w->vrec.p; a.jdeg = w->jdeg.p; a.jrank = w->jrank.p; a.sleep_ns = w->sleep_ns; a.slack_pct = w->slack_pct; a.fix_inverse = w->fix_inverse; a.dbg = w->timing ? w->dbg.p : nullptr;
    if (w->regions_on) {
        a.region_of = w->region_of.p; a.lidx = w->lidx.p; a.reg_start = w->reg_start.p; a.reg_bodies = w->reg_bodies.p;
        a.n_regions = w->n_regions; a.smem_bytes = (int)w->solve_budget;
        a.gadj = w->gadj.p; a.xrec = w->xrec.p; a.srec = w->srec.p;
        if (++w->step_tag > 0xffff0u) {  // 20-bit step tags (+ 12 bits of pass): start over with clean planes
            SYLT_CUDA(cudaMemsetAsync(w->xrec.p, 0, w->xrec.cap * sizeof(ulonglong2), st));
            w->step_tag = 1;
        }
        a.step_tag = w->step_tag;
        a.step_serial = w->step_serial;
        a.color_local = w->color_local ? 1 : 0;
        a.unc_bin = w->unc_bin.p; a.unc_bin_idx = w->unc_bin_idx.p; a.bin_cap = unc_bin_cap(w);
        a.probe_cta = getenv("SYLT_PROBE_CTA") ? atoi(getenv("SYLT_PROBE_CTA")) : w->regions_blocks / 2;
        SYLT_CUDA(cudaLaunchCooperativeKernel(w->regions_per_sm == 2 ? (void*)k_solve_regions<2> : (void*)k_solve_regions<1>, dim3((unsigned)w->regions_blocks),
                                              dim3((unsigned)w->regions_threads), args, w->solve_budget, st));
        w->arb[w->cur].sep_in_cin = full_step;
        w->arb[w->cur].lazy_inv_dt = a.inv_dt;
        w->arb[w->cur].lazy_kbias = a.poscorr ? 0.2f : 0.0f;
        w->steps_since_part++;
        w->last_regions = true;
    } else {
        w->last_regions = false;
        w->arb[w->cur].sep_in_cin = false;
        SYLT_CUDA(cudaLaunchCooperativeKernel((void*)k_solve, dim3((unsigned)w->coop_blocks), dim3((unsigned)w->solve_threads), args, w->solve_smem, st));
    }
    w->launches++;
    mark();
    w->kt_n = kt_n;
    // the set just built becomes "last step's"
    w->cur ^= 1;
    return SYLT_OK;
}

// SYLT_SOLVER_TIMING=1: section times of the last step's solver kernel (%globaltimer stamps of CTA 0) and, for
// k_solve_regions, the phases of pass 5 over all CTAs plus the load of the slowest / median / fastest CTA.
void print_solver_timing(SyltWorld* w) {
    if (w->last.err)
        fprintf(stderr, "step error %d (why %d, max slots %d, ghosts %d, uncoloured %d; pairs %d / %lld, arbiters %d / %lld, contacts %d / %lld, colouring rounds %d)\n",
                w->last.err, w->last.why, w->last.max_slots, w->last.n_ghost, w->last.n_uncolored, w->last.n_pairs, (long long)w->cap_pairs, w->last.n_arb,
                (long long)w->cap_arb, w->last.n_con, (long long)w->cap_con, w->last.rounds);
    const int G = w->last_regions ? w->regions_blocks : w->coop_blocks;
    std::vector<unsigned long long> t(64 + 8 * 1024 + 64);
    if (cudaMemcpy(t.data(), w->dbg.p, t.size() * sizeof(unsigned long long), cudaMemcpyDeviceToHost) != cudaSuccess) return;
    if (w->last_regions && t[32]) {
        fprintf(stderr, "  set-up of CTA %d [us]: generic table + barrier | bodies, rows, cone roots | cone walk | ghost bodies | layout, ghost data | slot records | scan | contacts:", getenv("SYLT_PROBE_CTA") ? atoi(getenv("SYLT_PROBE_CTA")) : w->regions_blocks / 2);
        for (int k = 33; k < 48 && t[k]; k++) fprintf(stderr, " %.1f", (t[k] - t[k - 1]) * 1e-3);
        fprintf(stderr, "\n");
    }
    if (w->last_regions && t[64] && G <= 1024) {
        const unsigned long long *pc = &t[64], *ld = &t[64 + 4 * 1024];  // per CTA: 4 stamps of pass 5; 4 words of load figures
        double mn[3] = {1e30, 1e30, 1e30}, mx[3] = {0, 0, 0}, sm[3] = {0, 0, 0};
        for (int c = 0; c < G; c++)
            for (int k = 0; k < 3; k++) {
                const double d = (double)(pc[c * 4 + k + 1] - pc[c * 4 + k]) * 1e-3;
                mn[k] = std::min(mn[k], d); mx[k] = std::max(mx[k], d); sm[k] += d;
            }
        std::vector<int> ord((size_t)G);
        for (int c = 0; c < G; c++) ord[(size_t)c] = c;
        std::sort(ord.begin(), ord.end(), [&](int x, int y) { return (ld[x * 4] >> 32) > (ld[y * 4] >> 32); });  // slowest set-up first
        for (int q : {0, 1, 2, G / 2, G - 1}) {
            const int c = ord[(size_t)q];
            const unsigned long long* d = &ld[c * 4];
            fprintf(stderr, "  CTA %3d: interior %.2f us | set-up %.1f us, cone walk %llu levels | %llu slots, %llu + %llu colours, ghosts %llu arbiters / %llu bodies, %llu own bodies, widest colour %llu\n", c,
                    (double)(pc[c * 4 + 1] - pc[c * 4]) * 1e-3, (double)(d[0] >> 32) * 1e-3, (d[0] >> 24) & 0xffull, d[0] & 0xffffffull, d[1] & 0xffffffffull, d[1] >> 32, d[2] & 0xffffffffull, d[2] >> 32,
                    d[3] & 0xffffffffull, d[3] >> 32);
        }
        fprintf(stderr, "  pass 5, all CTAs [us] min/avg/max: interior %.2f/%.2f/%.2f  export+import %.2f/%.2f/%.2f  generic %.2f/%.2f/%.2f\n", mn[0], sm[0] / G, mx[0],
                mn[1], sm[1] / G, mx[1], mn[2], sm[2] / G, mx[2]);
    }
    fprintf(stderr, "new arbiters %d (%d cross-region, %d promoted) of %d, colouring rounds %d\n", w->last.n_uncolored, w->last_regions ? w->last.region_new[w->n_regions] : 0, w->last.n_promoted, w->last.n_arb, w->last.rounds);
    for (int base : {48, 56})
        if (w->last_regions && t[(size_t)base]) {
            fprintf(stderr, "  colouring of CTA %s [us since kernel start]:", base == 48 ? "0" : "probe");
            for (int k = base; k < base + 8 && t[(size_t)k]; k++) fprintf(stderr, " %.1f", (double)(t[(size_t)k] - t[0]) * 1e-3);
            fprintf(stderr, "\n");
        }
    for (int k = 0; k < 8; k++) {
        const unsigned long long* h = &t[(size_t)(64 + 8 * 1024 + k * 8)];
        if (w->last_regions && h[5])
            fprintf(stderr, "  hop %d of pass 5, probe CTA, thread 0 [cycles]: loads %lld, contact a %lld, contact b %lld, stores %lld, barrier %lld (%llu slots in the colour)\n", k,
                    (long long)(h[1] - h[0]), (long long)(h[2] - h[1]), (long long)(h[3] - h[2]), (long long)(h[4] - h[3]), (long long)(h[5] - h[4]), h[6]);
    }
    fprintf(stderr, "k_solve sections [us]: colour %.1f order %.1f residency %.1f pass0 %.1f iterations %.1f writeback %.1f integrate %.1f total %.1f\n",
            (t[1] - t[0]) * 1e-3, (t[2] - t[1]) * 1e-3, (t[3] - t[2]) * 1e-3, (t[4] - t[3]) * 1e-3, (t[5] - t[4]) * 1e-3, (t[6] - t[5]) * 1e-3,
            (t[7] - t[6]) * 1e-3, (t[7] - t[0]) * 1e-3);
}

// Enqueue the steps of `segs` (no host synchronisation).
int enqueue_segments(SyltWorld* w, const std::vector<SyltWorld::Segment>& segs) {
    const int B = (int)w->hb.size();
    for (const SyltWorld::Segment& sg : segs)
        for (int s = 0; s < sg.n; s++) {
            int e = enqueue_step(w, sg.dt, sg.full);
            if (e) return e;
            w->inflight = true;
            w->inflight_full = sg.full;
            if (B > 0) {
                w->have_old = true;  // the set just built is what the next step matches against
                w->B_old = B;
                if (w->flags_stale) {  // the flags this step used become the "old" flags of the next one (unless the step fails)
                    k_copy_flags<<<(B + 255) / 256, 256, 0, w->stream>>>(B, w->bflags.p, w->bflags_old.p, w->counters.p);
                    SYLT_CUDA(cudaGetLastError());
                    w->launches++;
                    w->flags_stale = false;
                }
            }
        }
    return SYLT_OK;
}

int prepare_call(SyltWorld* w);

// After the queue drains: read the counters of the last step and update host-side bookkeeping.  A step that failed for lack
// of capacity changed nothing (every kernel after the overflow returned at once, and so did the steps enqueued behind it): the
// buffers are grown from the counts the failed step reported and the remaining steps of the call are run again.  The reference
// has no such limit (a HashMap, src/world.rs:24), so to the caller the overflow never happened.
int finish(SyltWorld* w) {
    for (int attempt = 0;; attempt++) {
        SYLT_CUDA(cudaStreamSynchronize(w->stream));
        if (!w->inflight) return w->have_last ? w->last.err : SYLT_OK;
        SYLT_CUDA(cudaMemcpy(&w->last, w->counters.p, sizeof(Counters), cudaMemcpyDeviceToHost));
        w->inflight = false;
        w->have_last = true;
        if (w->last_regions) w->last.n_colors = __builtin_popcountll(w->last.color_mask) + __builtin_popcount(w->last.color_mask_int);
        if (w->timing) print_solver_timing(w);
    if (w->timing && w->recut.p && w->steps_since_part <= 1) {
        RecutScratch* hs = new RecutScratch;
        if (cudaMemcpy(hs, w->recut.p, sizeof(RecutScratch), cudaMemcpyDeviceToHost) == cudaSuccess && hs->stamp[0]) {
            fprintf(stderr, "k_recut phases [us] (box + trim | keys + histogram | cell scan | scatter | sort in cells | cut):");
            for (int k = 1; k < 16 && hs->stamp[k]; k++) fprintf(stderr, " %.1f", (double)(hs->stamp[k] - hs->stamp[k - 1]) * 1e-3);
            fprintf(stderr, "\n");
        }
        delete hs;
    }
        if (w->kt_n > 1 && w->kernel_timing_print) {  // SYLT_KERNEL_TIMING: the last step's launches
            static const char* names[] = {"prepare", "scan cells", "fill cells", "rows count", "rows emit", "narrow", "(re-cut +) build", "solve"};
            fprintf(stderr, "kernels [us]:");
            for (int k = 1; k < w->kt_n; k++) {
                float ms = 0.0f;
                cudaEventElapsedTime(&ms, w->kt_ev[k - 1], w->kt_ev[k]);
                fprintf(stderr, " %s %.1f", k - 1 < 8 ? names[k - 1] : "?", ms * 1e3f);
            }
            fprintf(stderr, "\n");
        }
        w->last_was_step = w->inflight_full;
        w->last_iterations = w->iterations;
        const int err = w->last.err;
        int total = 0;
        for (const auto& sg : w->segs) total += sg.n;
        if (err == SYLT_OK) {
            if (w->regions_suspended_at > 0 && w->last.n_con > 0 && (int64_t)w->last.n_con * 4 < (int64_t)w->regions_suspended_at * 3) {
                // a world that fell back to k_solve because a region did not fit: three quarters of the contacts it had then — try again
                // (the next failure suspends it at the lower count, so the retries die out)
                w->regions_mode = w->regions_mode_saved;
                w->regions_suspended_at = 0;
                w->part_dirty = true;
            }
            w->n_arb_valid = (int)std::min<int64_t>(w->last.n_arb, w->cap_arb);
            w->n_con_valid = (int)std::min<int64_t>(w->last.n_con, w->cap_con);
            w->segs.clear();
            return SYLT_OK;
        }
        // ---- a step of this call failed: the steps behind it did not run.  Put the host-side bookkeeping where the device is.
        const int completed = std::max(0, std::min(w->last.steps_done, total));
        const int ran = completed + (err == SYLT_ERR_NO_INVERSE ? 1 : 0);  // a NoInverse step keeps its arbiter set (world.rs:110-129 ran)
        if (!w->hb.empty()) {
            w->cur = w->call_cur0 ^ (ran & 1);
            if (ran == 0) { w->have_old = w->call_have_old0; w->B_old = w->call_B_old0; w->flags_stale = w->call_flags_stale0; }
        }
        w->cells_clean = false;
        replay_schedule(w, w->call_since0, w->call_pending0, w->call_recolor0, w->segs, ran);  // re-cuts enqueued behind the failure did not happen
        if (err == SYLT_ERR_NO_INVERSE) {
            w->n_arb_valid = (int)std::min<int64_t>(w->last.n_arb, w->cap_arb);
            w->n_con_valid = (int)std::min<int64_t>(w->last.n_con, w->cap_con);
        } else if (completed > 0) {  // the last good set is the one of the last completed step of this call
            w->n_arb_valid = (int)std::min<int64_t>(w->last.good_n_arb, w->cap_arb);
            w->n_con_valid = (int)std::min<int64_t>(w->last.good_n_con, w->cap_con);
        }
        std::vector<SyltWorld::Segment> rest;  // the steps that did not run
        {
            int skip = completed;
            for (const auto& sg : w->segs) {
                const int take = std::min(skip, sg.n);
                skip -= take;
                if (sg.n > take) rest.push_back({sg.dt, sg.n - take, sg.full});
            }
        }
        w->segs.clear();
        if (err != SYLT_ERR_CAPACITY || attempt >= 8) return err;
        bool grown = false;
        if (w->last.n_pairs > w->cap_pairs) { w->grow_pairs = std::max<int64_t>(2 * w->cap_pairs, (int64_t)w->last.n_pairs * 3 / 2); grown = true; }
        if (w->last.n_arb > w->cap_arb) { w->grow_arb = std::max<int64_t>(2 * w->cap_arb, (int64_t)w->last.n_arb * 3 / 2); grown = true; }
        if (w->last.n_con > w->cap_con) { w->grow_con = std::max<int64_t>(2 * w->cap_con, (int64_t)w->last.n_con * 3 / 2); grown = true; }
        if (w->last_regions && w->last.why != 0) {  // a region outgrew k_solve_regions' tables: this world goes on with k_solve —
            // until its contact count has come down again (a pile is densest while it lands): see the success path above
            w->regions_mode_saved = w->regions_mode;
            w->regions_suspended_at = std::max(1, w->last.n_con);
            w->regions_mode = 0;
            w->part_dirty = true;
            grown = true;
        }
        if (w->user_pairs && w->grow_pairs > w->user_pairs) w->user_pairs = 0;  // an explicit capacity is a starting point, not a limit
        if (w->user_arb && w->grow_arb > w->user_arb) w->user_arb = 0;
        if (w->user_con && w->grow_con > w->user_con) w->user_con = 0;
        if (!grown) return err;  // nothing a larger buffer would cure (> 64 colours at a body, a polygon manifold over the cap)
        w->class_dirty = true;
        int e = prepare_call(w);
        if (e) return e;
        w->segs = rest;
        e = enqueue_segments(w, rest);
        if (e) return e;
    }
}

// Everything that has to happen on the host before steps are enqueued: uploads, (re)allocation, the solver choice; and the
// snapshot of the bookkeeping a failed step restores.
int prepare_call(SyltWorld* w) {
    int e = flush(w);
    if (e) return e;
    if (body_order_violated(w->hb)) return SYLT_ERR_BODY_ORDER;
    if (w->part_dirty) {  // bodies / joints / iterations changed: region solver on or off, buffers
        e = repartition(w);
        if (e) return e;
    }
    if (!w->inflight) {
        SYLT_CUDA(cudaMemsetAsync(w->counters.p, 0, offsetof(Counters, n_pairs), w->stream));  // sticky error fields, steps_done
        w->segs.clear();
        w->call_cur0 = w->cur; w->call_have_old0 = w->have_old; w->call_B_old0 = w->B_old; w->call_flags_stale0 = w->flags_stale;
        w->call_since0 = w->steps_since_part; w->call_pending0 = w->resort_pending; w->call_recolor0 = w->recolor_all;
    }
    return SYLT_OK;
}

// r1 / r2 / masses / separation of the last completed arbiter set, if the region solver left them to be computed on demand.  Must run
// before anything they are computed from changes (the masses: sylt_world_set_body_props).
int materialise_arbiters(SyltWorld* w) {
    SyltWorld::ArbBufs& ab = w->arb[w->cur ^ 1];
    if (!ab.sep_in_cin) return SYLT_OK;
    ab.sep_in_cin = false;
    const int A = w->have_last ? (int)std::min<int64_t>(w->n_arb_valid, w->cap_arb) : 0;
    if (A <= 0) return SYLT_OK;
    SYLT_CUDA(cudaSetDevice(w->device));
    k_fill_contacts<<<(A + 127) / 128, 128, 0, w->stream>>>(A, ab.bodies.p, ab.meta.p, ab.cin.p, w->pose_prev.p, w->mp.p, ab.lazy_inv_dt, ab.lazy_kbias,
                                                             ab.ca.p, ab.cb.p, ab.cc.p);
    SYLT_CUDA(cudaGetLastError());
    return SYLT_OK;
}

int do_steps(SyltWorld* w, float dt, int n, bool full, bool wait) {
    if (!w) return SYLT_ERR_INVALID;
    int e = prepare_call(w);
    if (e) return e;
    std::vector<SyltWorld::Segment> sg = {{dt, n, full}};
    w->segs.push_back(sg[0]);
    e = enqueue_segments(w, sg);
    if (e) return e;
    if (!wait) return SYLT_OK;
    return finish(w);
}

}  // namespace

// ====================================================================== C ABI
extern "C" {

const char* sylt_last_error_string(void) { return g_last_error.c_str(); }

int sylt_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) return 0;
    return n;
}

int sylt_world_create(float gx, float gy, uint32_t iterations, int device, SyltWorld** out) {
    if (!out) return SYLT_ERR_INVALID;
    *out = nullptr;
    int n = 0;
    SYLT_CUDA(cudaGetDeviceCount(&n));
    if (device < 0 || device >= n) { set_last_error("no such CUDA device"); return SYLT_ERR_CUDA; }
    SYLT_CUDA(cudaSetDevice(device));
    SyltWorld* w = new SyltWorld();
    w->device = device; w->gx = gx; w->gy = gy; w->iterations = iterations;
    cudaDeviceProp prop;
    SYLT_CUDA(cudaGetDeviceProperties(&prop, device));
    w->num_sms = prop.multiProcessorCount;
    int coop = 0;
    SYLT_CUDA(cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, device));
    if (!coop) { delete w; set_last_error("device lacks cooperative launch"); return SYLT_ERR_CUDA; }
    // k_solve: one persistent CTA per SM; its shared memory holds the CTA's share of arbiter records and contact constants
    int max_smem = 0;
    SYLT_CUDA(cudaDeviceGetAttribute(&max_smem, cudaDevAttrMaxSharedMemoryPerBlockOptin, device));
    cudaFuncAttributes fa;
    SYLT_CUDA(cudaFuncGetAttributes(&fa, k_solve));
    w->ctas_per_sm = 1;
    if (const char* ev = getenv("SYLT_SOLVER_CTAS_PER_SM")) w->ctas_per_sm = std::min(4, std::max(1, atoi(ev)));
    w->solve_threads = SOLVE_THREADS / w->ctas_per_sm;
    size_t budget = ((size_t)max_smem + 1024) / (size_t)w->ctas_per_sm - fa.sharedSizeBytes - 2048;  // 1 KB per CTA is reserved by the system
    w->smem_slots = 1792 / w->ctas_per_sm;
    if (const char* ev = getenv("SYLT_SOLVER_SLEEP_NS")) w->sleep_ns = std::max(0, atoi(ev));
    if (const char* ev = getenv("SYLT_SOLVER_SLACK")) w->slack_pct = std::max(0, atoi(ev));
    if (const char* ev = getenv("SYLT_SOLVER_TIMING")) w->timing = atoi(ev) != 0;
    SYLT_CUDA(w->dbg.ensure(64 + 8 * 1024 + 64));
    SYLT_CUDA(cudaMemset(w->dbg.p, 0, (64 + 8 * 1024 + 64) * sizeof(unsigned long long)));
    if (const char* ev = getenv("SYLT_SOLVER_SMEM_SLOTS")) w->smem_slots = std::max(0, atoi(ev));  // tests: force the global-memory path
    size_t fixed = (size_t)w->smem_slots * 32 + (((size_t)w->smem_slots + 3) & ~(size_t)3) * 4;
    if (fixed + 48 > budget) { w->smem_slots = 0; fixed = 0; }
    w->smem_contacts = w->smem_slots ? (int)((budget - fixed) / 48) : 0;
    if (const char* ev = getenv("SYLT_SOLVER_SMEM_CONTACTS")) w->smem_contacts = std::min(w->smem_contacts, std::max(0, atoi(ev)));
    w->solve_smem = fixed + (size_t)w->smem_contacts * 48;
    SYLT_CUDA(cudaFuncSetAttribute(k_solve, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)w->solve_smem));
    {   // k_solve_regions: the SM's shared memory, carved per partition (bodies of the region, slot records, contacts)
        if (const char* ev = getenv("SYLT_REGION_CTAS_PER_SM")) w->regions_per_sm = atoi(ev) == 2 ? 2 : 1;
        const void* kern = w->regions_per_sm == 2 ? (const void*)k_solve_regions<2> : (const void*)k_solve_regions<1>;
        cudaFuncAttributes fr;
        SYLT_CUDA(cudaFuncGetAttributes(&fr, kern));
        const size_t b2 = ((size_t)max_smem + 1024) / (size_t)w->regions_per_sm - 1024 - fr.sharedSizeBytes - 1024;  // 1 KB per CTA is reserved by the system
        w->solve_budget = b2 & ~(size_t)15;
        w->regions_threads = SOLVE_THREADS / w->regions_per_sm;
        w->regions_blocks = w->num_sms * w->regions_per_sm;
        SYLT_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)w->solve_budget));
        SYLT_CUDA(cudaFuncSetAttribute(k_recut, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)((SORT_CELLS + 1) * sizeof(int))));
        int per2 = 0;
        SYLT_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per2, kern, w->regions_threads, w->solve_budget));
        if (const char* ev = getenv("SYLT_REGIONS")) w->regions_mode = std::min(2, std::max(0, atoi(ev)));
        if (const char* ev = getenv("SYLT_REGIONS_MIN_BODIES")) w->regions_min_bodies = std::max(1, atoi(ev));
        if (const char* ev = getenv("SYLT_REGIONS_COUNT")) w->regions_count = std::max(0, atoi(ev));
        if (const char* ev = getenv("SYLT_REGIONS_PERIOD")) w->regions_period = std::max(1, atoi(ev));
        if (const char* ev = getenv("SYLT_COLOR_LOCAL")) w->color_local = atoi(ev) != 0;
        if (const char* ev = getenv("SYLT_KERNEL_TIMING")) w->kernel_timing = w->kernel_timing_print = atoi(ev) != 0;
        if (per2 < w->regions_per_sm || w->ctas_per_sm != 1) w->regions_mode = 0;
    }
    int per_sm = 0;
    SYLT_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_solve, w->solve_threads, w->solve_smem));
    if (per_sm < w->ctas_per_sm) { delete w; set_last_error("k_solve does not fit on an SM"); return SYLT_ERR_CUDA; }
    w->coop_blocks = w->num_sms * w->ctas_per_sm;
    SYLT_CUDA(cudaStreamCreateWithFlags(&w->stream, cudaStreamNonBlocking));
    SYLT_CUDA(cudaEventCreate(&w->ev0));
    SYLT_CUDA(cudaEventCreate(&w->ev1));
    SYLT_CUDA(w->counters.ensure(1));
    SYLT_CUDA(cudaMemset(w->counters.p, 0, sizeof(Counters)));
    *out = w;
    return SYLT_OK;
}

static void free_all(SyltWorld* w) {
    if (w->rb_host) { cudaFreeHost(w->rb_host); w->rb_host = nullptr; w->rb_cap = 0; }
    if (w->sb_host) { cudaFreeHost(w->sb_host); w->sb_host = nullptr; w->sb_cap = 0; }
    w->pack_dev.release();
    w->pose.release(); w->vel.release(); w->force.release(); w->pose_prev.release(); w->mp.release(); w->geom.release(); w->aabb.release();
    w->rotc.release(); w->lverts.release(); w->bflags.release(); w->bflags_old.release(); w->rank.release();
    w->cell_count.release(); w->cell_start.release(); w->cell_local.release(); w->tile_tot.release(); w->cent_info.release(); w->large.release(); w->row_count.release();
    w->row_start.release(); w->row_tmp.release(); w->large_start.release(); w->row_btot.release(); w->pair_btot.release(); w->cinfo.release(); w->pairs.release(); w->info.release(); w->pscan.release(); w->claim.release();
    w->tmp_a.release(); w->tmp_b.release(); w->used.release(); w->used_int.release(); w->vrec.release(); w->jdeg.release(); w->jrank.release(); w->uncolored.release(); w->uncolored2.release(); w->unc_rec.release();
    w->prop.release(); w->order.release();
    for (int k = 0; k < 2; k++) {
        auto& a = w->arb[k];
        a.bodies.release(); a.meta.release(); a.partner.release(); a.color.release(); a.row_start.release(); a.friction.release();
        a.ca.release(); a.cb.release(); a.cc.release(); a.cd.release(); a.cin.release();
    }
    w->jbodies.release(); w->jla.release(); w->jr.release(); w->jm.release(); w->jparam.release(); w->jp.release(); w->jbias.release();
    w->jorder.release(); w->jcolor_start.release();
    w->region_of.release(); w->lidx.release(); w->reg_start.release(); w->reg_bodies.release(); w->sort_keys.release(); w->recut.release(); w->gadj.release(); w->xrec.release();
}

int sylt_world_destroy(SyltWorld* w) {
    if (!w) return SYLT_OK;
    cudaSetDevice(w->device);
    cudaStreamSynchronize(w->stream);
    free_all(w);
    w->counters.release();
    cudaEventDestroy(w->ev0); cudaEventDestroy(w->ev1);
    cudaStreamDestroy(w->stream);
    delete w;
    return SYLT_OK;
}

int sylt_world_clear(SyltWorld* w) {  // world.rs:67-71
    if (!w) return SYLT_ERR_INVALID;
    cudaSetDevice(w->device);
    cudaStreamSynchronize(w->stream);
    w->hb.clear(); w->hverts.clear(); w->pending.clear(); w->hj.clear(); w->h_jorder.clear();
    w->uploaded = 0; w->class_dirty = true; w->joints_dirty = true; w->any_poly = false; w->max_poly_verts = 0; w->n_large = 0;
    w->have_old = false; w->B_old = 0; w->have_last = false; w->n_joint_colors = 0; w->inflight = false; w->flags_stale = false;
    w->regions_on = false; w->part_dirty = true; w->recolor_all = false; w->last_regions = false;
    w->segs.clear(); w->n_arb_valid = 0; w->n_con_valid = 0;
    memset(&w->last, 0, sizeof w->last);
    return SYLT_OK;
}

int sylt_world_set_context(SyltWorld* w, int acc, int warm, int pc) {
    if (!w) return SYLT_ERR_INVALID;
    w->accumulate = acc != 0; w->warm = warm != 0; w->poscorr = pc != 0;
    return SYLT_OK;
}
int sylt_world_set_fixes(SyltWorld* w, int fix_features, int fix_inverse) {
    if (!w) return SYLT_ERR_INVALID;
    w->fix_features = fix_features != 0; w->fix_inverse = fix_inverse != 0;
    return SYLT_OK;
}
int sylt_world_set_iterations(SyltWorld* w, uint32_t it) {
    if (!w) return SYLT_ERR_INVALID;
    if (w->iterations != it) w->part_dirty = true;  // (the region solver handles up to MAX_PASSES - 1 iterations)
    w->iterations = it;
    return SYLT_OK;
}
int sylt_world_set_capacity(SyltWorld* w, int64_t p, int64_t a, int64_t c) {
    if (!w || p < 0 || a < 0 || c < 0) return SYLT_ERR_INVALID;
    if (p) w->user_pairs = p;
    if (a) w->user_arb = a;
    if (c) w->user_con = c;
    w->class_dirty = true;
    return SYLT_OK;
}

int sylt_world_add_bodies(SyltWorld* w, const SyltBodyDesc* bodies, int32_t n, const float* verts_xy, int32_t n_verts) {
    if (!w || n < 0 || (n > 0 && !bodies)) return SYLT_ERR_INVALID;
    size_t b0 = w->hb.size(), v0 = w->hverts.size();
    for (int i = 0; i < n; i++) {
        HostBody hb;
        int e = make_host_body(bodies[i], verts_xy, n_verts, &hb, &w->hverts);
        if (e) { w->hb.resize(b0); w->hverts.resize(v0); w->pending.resize(b0 - (size_t)w->uploaded); return e; }
        if (hb.shape == SYLT_SHAPE_POLYGON) { w->any_poly = true; w->max_poly_verts = std::max(w->max_poly_verts, hb.nverts); }
        w->hb.push_back(hb);
        SyltBodyState s = {bodies[i].x, bodies[i].y, bodies[i].rotation, bodies[i].vx, bodies[i].vy, bodies[i].angular_velocity, 0.0f, 0.0f, 0.0f};
        w->pending.push_back(s);
    }
    return SYLT_OK;
}
int sylt_world_add_box(SyltWorld* w, uint64_t id, float wx, float wy, float mass, float friction, float x, float y, float rot,
                       float vx, float vy, float om, int32_t* out_index) {
    SyltBodyDesc d;
    memset(&d, 0, sizeof d);
    d.id = id; d.shape = SYLT_SHAPE_BOX; d.width_x = wx; d.width_y = wy; d.mass = mass; d.friction = friction;
    d.x = x; d.y = y; d.rotation = rot; d.vx = vx; d.vy = vy; d.angular_velocity = om;
    int e = sylt_world_add_bodies(w, &d, 1, nullptr, 0);
    if (!e && out_index) *out_index = (int32_t)w->hb.size() - 1;
    return e;
}
int sylt_world_add_polygon(SyltWorld* w, uint64_t id, const float* verts_xy, int32_t nv, float mass, float friction, float x, float y,
                           float rot, float vx, float vy, float om, int32_t* out_index) {
    SyltBodyDesc d;
    memset(&d, 0, sizeof d);
    d.id = id; d.shape = SYLT_SHAPE_POLYGON; d.vert_offset = 0; d.vert_count = nv; d.mass = mass; d.friction = friction;
    d.x = x; d.y = y; d.rotation = rot; d.vx = vx; d.vy = vy; d.angular_velocity = om;
    int e = sylt_world_add_bodies(w, &d, 1, verts_xy, nv);
    if (!e && out_index) *out_index = (int32_t)w->hb.size() - 1;
    return e;
}

static int find_body(SyltWorld* w, uint64_t id) {  // joint.rs:27-36: first body with this id
    for (size_t i = 0; i < w->hb.size(); i++)
        if (w->hb[i].id == id) return (int)i;
    return -1;
}

static int add_joint_impl(SyltWorld* w, const SyltJointDesc& d, const std::vector<SyltBodyState>* states) {
    int b1 = find_body(w, d.id1), b2 = find_body(w, d.id2);
    if (b1 < 0 || b2 < 0) return SYLT_ERR_BODY_NOT_FOUND;
    const SyltBodyState& s1 = (*states)[(size_t)b1];
    const SyltBodyState& s2 = (*states)[(size_t)b2];
    // Joint::new (joint.rs:37-40): local anchors from the bodies' current pose
    M2 rt1 = transpose(rot_angle(s1.rotation)), rt2 = transpose(rot_angle(s2.rotation));
    V2 anchor = mk(d.anchor_x, d.anchor_y);
    V2 la1 = rt1 * (anchor - mk(s1.x, s1.y)), la2 = rt2 * (anchor - mk(s2.x, s2.y));
    HostJoint j = {b1, b2, la1.x, la1.y, la2.x, la2.y, d.softness, d.bias_factor, 0};
    w->hj.push_back(j);
    return SYLT_OK;
}

int sylt_world_get_bodies(SyltWorld* w, SyltBodyState* out, int32_t cap);

int sylt_world_add_joints(SyltWorld* w, const SyltJointDesc* joints, int32_t n) {
    if (!w || n < 0 || (n > 0 && !joints)) return SYLT_ERR_INVALID;
    if (n == 0) return SYLT_OK;
    // current poses: device state for uploaded bodies, pending states for the rest
    std::vector<SyltBodyState> st(w->hb.size());
    if (w->uploaded > 0) {
        int e = sylt_world_get_bodies(w, st.data(), w->uploaded);
        if (e < 0) return -e;
    }
    for (size_t k = 0; k < w->pending.size(); k++) st[(size_t)w->uploaded + k] = w->pending[k];
    size_t j0 = w->hj.size();
    for (int i = 0; i < n; i++) {
        int e = add_joint_impl(w, joints[i], &st);
        if (e) { w->hj.resize(j0); return e; }
    }
    // accumulated impulses of the new joints start at zero; existing ones are kept
    cudaSetDevice(w->device);
    size_t J = w->hj.size();
    SYLT_CUDA(w->jp.ensure(J + 1, j0, w->stream));
    SYLT_CUDA(cudaMemsetAsync(w->jp.p + j0, 0, (J - j0) * sizeof(float2), w->stream));
    SYLT_CUDA(cudaStreamSynchronize(w->stream));
    w->joints_dirty = true;
    w->part_dirty = true;  // worlds with joints use k_solve
    return SYLT_OK;
}
int sylt_world_add_joint(SyltWorld* w, uint64_t id1, uint64_t id2, float ax, float ay, int32_t* out_index) {
    SyltJointDesc d = {id1, id2, ax, ay, 0.0f, 0.2f};  // joint.rs:47-48 defaults
    int e = sylt_world_add_joints(w, &d, 1);
    if (!e && out_index) *out_index = (int32_t)w->hj.size() - 1;
    return e;
}
static bool same_bits(float a, float b) { return memcmp(&a, &b, sizeof a) == 0; }
int sylt_world_set_joint_params(SyltWorld* w, int32_t j, float softness, float bias_factor) {
    if (!w || j < 0 || (size_t)j >= w->hj.size()) return SYLT_ERR_INVALID;
    HostJoint& q = w->hj[(size_t)j];
    if (same_bits(q.softness, softness) && same_bits(q.bias_factor, bias_factor)) return SYLT_OK;  // (wrappers push the pub fields before every step)
    q.softness = softness;
    q.bias_factor = bias_factor;
    w->jparams_dirty = true;
    return SYLT_OK;
}
int sylt_world_set_joint_local_anchors(SyltWorld* w, int32_t j, float la1x, float la1y, float la2x, float la2y) {
    if (!w || j < 0 || (size_t)j >= w->hj.size()) return SYLT_ERR_INVALID;
    HostJoint& q = w->hj[(size_t)j];
    if (same_bits(q.la1x, la1x) && same_bits(q.la1y, la1y) && same_bits(q.la2x, la2x) && same_bits(q.la2y, la2y)) return SYLT_OK;
    q.la1x = la1x; q.la1y = la1y; q.la2x = la2x; q.la2y = la2y;
    w->jparams_dirty = true;
    return SYLT_OK;
}
int sylt_world_add_joint_local(SyltWorld* w, uint64_t id1, uint64_t id2, float la1x, float la1y, float la2x, float la2y, float softness,
                               float bias_factor, int32_t* out_index) {
    if (!w) return SYLT_ERR_INVALID;
    const int b1 = find_body(w, id1), b2 = find_body(w, id2);
    if (b1 < 0 || b2 < 0) return SYLT_ERR_BODY_NOT_FOUND;
    if (w->inflight) finish(w);
    cudaSetDevice(w->device);
    const size_t j0 = w->hj.size();
    HostJoint j = {b1, b2, la1x, la1y, la2x, la2y, softness, bias_factor, 0};
    w->hj.push_back(j);
    SYLT_CUDA(w->jp.ensure(j0 + 2, j0, w->stream));  // accumulated impulses: the new joint starts at zero, existing ones are kept
    SYLT_CUDA(cudaMemsetAsync(w->jp.p + j0, 0, sizeof(float2), w->stream));
    SYLT_CUDA(cudaStreamSynchronize(w->stream));
    w->joints_dirty = true;
    w->part_dirty = true;  // worlds with joints use k_solve
    if (out_index) *out_index = (int32_t)j0;
    return SYLT_OK;
}

// pub Body fields the reference reads on every step, written back as the caller has them (nothing is re-derived from mass)
int sylt_world_set_body_props(SyltWorld* w, int32_t first, int32_t count, const float* props8) {
    if (!w || first < 0 || count < 0 || (size_t)first + (size_t)count > w->hb.size() || (count > 0 && !props8)) return SYLT_ERR_INVALID;
    if (count == 0) return SYLT_OK;
    int e = flush(w);  // (drains the queue, uploads pending bodies)
    if (e) return e;
    if ((e = materialise_arbiters(w))) return e;  // (computed from the masses the last step used)
    std::vector<float4> hm((size_t)count), hg((size_t)count);
    for (int k = 0; k < count; k++) {
        HostBody& b = w->hb[(size_t)(first + k)];
        const float* p = props8 + 8 * (size_t)k;
        b.mass = p[0]; b.inv_mass = p[1]; b.moi = p[2]; b.inv_moi = p[3]; b.friction = p[6];
        if (b.shape == SYLT_SHAPE_BOX) {  // collide() reads Body.width (collide.rs:141-142); box-polygon pairs keep using the vertices stored at Body::new
            b.width_x = p[4]; b.width_y = p[5];
            float r = 0.0f;
            for (int v = 0; v < b.nverts; v++) {
                const V2 l = w->hverts[(size_t)(b.vert_offset + v)];
                r = std::max(r, sqrtf(l.x * l.x + l.y * l.y));
            }
            const float rb = 0.5f * sqrtf(b.width_x * b.width_x + b.width_y * b.width_y);
            b.radius = (rb > r || !(r == r)) ? rb : r;
            if (!(b.radius == b.radius)) b.radius = 0.0f;
        }
        hm[(size_t)k] = make_float4(b.inv_mass, b.inv_moi, b.friction, 0.0f);
        float off;
        const int o = b.vert_offset;
        memcpy(&off, &o, 4);
        hg[(size_t)k] = make_float4(b.width_x, b.width_y, off, b.radius);
    }
    SYLT_CUDA(cudaMemcpyAsync(w->mp.p + first, hm.data(), (size_t)count * sizeof(float4), cudaMemcpyHostToDevice, w->stream));
    SYLT_CUDA(cudaMemcpyAsync(w->geom.p + first, hg.data(), (size_t)count * sizeof(float4), cudaMemcpyHostToDevice, w->stream));
    SYLT_CUDA(cudaStreamSynchronize(w->stream));
    w->class_dirty = true;   // static / large classification and the grid cell depend on inv_mass and the radius
    w->joints_dirty = true;  // the joint colouring treats bodies without inverse mass as static
    return SYLT_OK;
}

int sylt_world_add_force(SyltWorld* w, int32_t body, float fx, float fy) {  // body.rs:286-288
    if (!w || body < 0 || (size_t)body >= w->hb.size()) return SYLT_ERR_INVALID;
    if (body >= w->uploaded) {
        SyltBodyState& s = w->pending[(size_t)(body - w->uploaded)];
        s.fx = s.fx + fx; s.fy = s.fy + fy;
        return SYLT_OK;
    }
    cudaSetDevice(w->device);
    float4 f;
    SYLT_CUDA(cudaMemcpyAsync(&f, w->force.p + body, sizeof f, cudaMemcpyDeviceToHost, w->stream));
    SYLT_CUDA(cudaStreamSynchronize(w->stream));
    f.x = f.x + fx; f.y = f.y + fy;
    SYLT_CUDA(cudaMemcpyAsync(w->force.p + body, &f, sizeof f, cudaMemcpyHostToDevice, w->stream));
    SYLT_CUDA(cudaStreamSynchronize(w->stream));
    return SYLT_OK;
}

int sylt_world_num_bodies(SyltWorld* w) { return w ? (int)w->hb.size() : 0; }
int sylt_world_num_joints(SyltWorld* w) { return w ? (int)w->hj.size() : 0; }
int sylt_world_num_arbiters(SyltWorld* w) { return (w && w->have_last) ? w->n_arb_valid : 0; }
int sylt_world_num_contacts(SyltWorld* w) { return (w && w->have_last) ? w->n_con_valid : 0; }

int sylt_world_get_bodies(SyltWorld* w, SyltBodyState* out, int32_t cap) {
    if (!w || (!out && cap > 0)) return -SYLT_ERR_INVALID;
    int n = std::min<int>(cap, (int)w->hb.size());
    int nd = std::min(n, w->uploaded);
    if (nd > 0) {
        if (cudaSetDevice(w->device) != cudaSuccess) return -SYLT_ERR_CUDA;
        // packed on the device (9 floats per body), one copy through the handle's pinned staging area
        const size_t bytes = (size_t)nd * sizeof(SyltBodyState);
        if (w->pack_dev.ensure((size_t)nd * 9) != cudaSuccess) return -SYLT_ERR_CUDA;
        if (bytes > w->sb_cap) {
            if (w->sb_host) cudaFreeHost(w->sb_host);
            w->sb_host = nullptr; w->sb_cap = 0;
            if (cudaMallocHost(&w->sb_host, bytes + bytes / 4 + 64) != cudaSuccess) { set_last_error("cudaMallocHost (body read-back)"); return -SYLT_ERR_CUDA; }
            w->sb_cap = bytes + bytes / 4 + 64;
        }
        k_pack_states<<<(nd + 255) / 256, 256, 0, w->stream>>>(nd, w->pose.p, w->vel.p, w->force.p, w->pack_dev.p);
        cudaError_t e = cudaMemcpyAsync(w->sb_host, w->pack_dev.p, bytes, cudaMemcpyDeviceToHost, w->stream);
        if (e == cudaSuccess) e = cudaStreamSynchronize(w->stream);
        if (e != cudaSuccess) { set_last_error(cudaGetErrorString(e)); return -SYLT_ERR_CUDA; }
        memcpy(out, w->sb_host, bytes);
    }
    for (int i = nd; i < n; i++) out[i] = w->pending[(size_t)(i - w->uploaded)];
    return n;
}

int sylt_world_set_bodies(SyltWorld* w, int32_t first, int32_t count, const SyltBodyState* in) {
    if (!w || first < 0 || count < 0 || (size_t)first + (size_t)count > w->hb.size() || (count > 0 && !in)) return SYLT_ERR_INVALID;
    int e = flush(w);
    if (e) return e;
    std::vector<float4> p((size_t)count), v((size_t)count), f((size_t)count);
    for (int i = 0; i < count; i++) {
        p[(size_t)i] = make_float4(in[i].x, in[i].y, in[i].rotation, 0.0f);
        v[(size_t)i] = make_float4(in[i].vx, in[i].vy, in[i].angular_velocity, 0.0f);
        f[(size_t)i] = make_float4(in[i].fx, in[i].fy, in[i].torque, 0.0f);
    }
    if (count) {
        SYLT_CUDA(cudaMemcpyAsync(w->pose.p + first, p.data(), (size_t)count * sizeof(float4), cudaMemcpyHostToDevice, w->stream));
        SYLT_CUDA(cudaMemcpyAsync(w->vel.p + first, v.data(), (size_t)count * sizeof(float4), cudaMemcpyHostToDevice, w->stream));
        SYLT_CUDA(cudaMemcpyAsync(w->force.p + first, f.data(), (size_t)count * sizeof(float4), cudaMemcpyHostToDevice, w->stream));
        SYLT_CUDA(cudaStreamSynchronize(w->stream));
    }
    return SYLT_OK;
}

int sylt_world_get_body_props(SyltWorld* w, float* o, int32_t cap) {
    if (!w) return -SYLT_ERR_INVALID;
    int n = std::min<int>(cap, (int)w->hb.size());
    for (int i = 0; i < n; i++) {
        const HostBody& b = w->hb[(size_t)i];
        float* p = o + 8 * i;
        p[0] = b.mass; p[1] = b.inv_mass; p[2] = b.moi; p[3] = b.inv_moi; p[4] = b.width_x; p[5] = b.width_y; p[6] = b.friction; p[7] = (float)b.shape;
    }
    return n;
}

int sylt_world_get_arbiters(SyltWorld* w, SyltArbiterInfo* arbiters, int32_t arbiter_cap, SyltContact* contacts, int32_t contact_cap) {
    if (!w) return -SYLT_ERR_INVALID;
    int A = sylt_world_num_arbiters(w), Cn = sylt_world_num_contacts(w);
    if (A > arbiter_cap || Cn > contact_cap) return -SYLT_ERR_CAPACITY;
    if (A == 0) return 0;
    if (cudaSetDevice(w->device) != cudaSuccess) return -SYLT_ERR_CUDA;
    SyltWorld::ArbBufs& ab = w->arb[w->cur ^ 1];  // last completed step
    if (materialise_arbiters(w)) return -SYLT_ERR_CUDA;
    // one pinned staging area (kept and grown by the handle: no page faults, DMA at link speed), the eight arrays behind one another
    const size_t bytes = (size_t)A * (2 * sizeof(int2) + sizeof(float) + sizeof(int)) + (size_t)Cn * 4 * sizeof(float4) + 8 * 16;
    if (bytes > w->rb_cap) {
        if (w->rb_host) cudaFreeHost(w->rb_host);
        w->rb_host = nullptr; w->rb_cap = 0;
        if (cudaMallocHost(&w->rb_host, bytes + bytes / 4) != cudaSuccess) { set_last_error("cudaMallocHost (arbiter read-back)"); return -SYLT_ERR_CUDA; }
        w->rb_cap = bytes + bytes / 4;
    }
    unsigned char* hp = (unsigned char*)w->rb_host;
    auto take = [&](const void* dev, size_t n) {
        void* h = hp;
        cudaMemcpyAsync(h, dev, n, cudaMemcpyDeviceToHost, w->stream);
        hp += (n + 15) & ~(size_t)15;
        return h;
    };
    const int2* bodies = (const int2*)take(ab.bodies.p, (size_t)A * sizeof(int2));
    const int2* meta = (const int2*)take(ab.meta.p, (size_t)A * sizeof(int2));
    const float* fr = (const float*)take(ab.friction.p, (size_t)A * sizeof(float));
    const int* col = (const int*)take(ab.color.p, (size_t)A * sizeof(int));
    const float4* ca = (const float4*)take(ab.ca.p, (size_t)Cn * sizeof(float4));
    const float4* cb = (const float4*)take(ab.cb.p, (size_t)Cn * sizeof(float4));
    const float4* cc = (const float4*)take(ab.cc.p, (size_t)Cn * sizeof(float4));
    const float4* cd = (const float4*)take(ab.cd.p, (size_t)Cn * sizeof(float4));
    cudaError_t e = cudaStreamSynchronize(w->stream);
    if (e != cudaSuccess) { set_last_error(cudaGetErrorString(e)); return -SYLT_ERR_CUDA; }
    // key order (ArbiterKey: id pair = body pair); the device order (owner body, partner) already is that order unless small-large
    // pairs are around: sort only then
    // key order (ArbiterKey: id pair = body pair).  The device order is (owner body, partner): already key order except for the pairs
    // of a small body with a large one of lower index (the small body owns those).  So: the ascending run that the device order
    // contains, the few stragglers sorted, and a merge — not a sort of everything.
    std::vector<int> idx((size_t)A), run, rest;
    std::vector<int> coff((size_t)A + 1);
    run.reserve((size_t)A);
    auto less = [&](int x, int y) {
        const int2 &a = bodies[x], &b = bodies[y];
        return a.x != b.x ? a.x < b.x : a.y < b.y;
    };
    for (int i = 0; i < A; i++) {
        if (run.empty() || !less(i, run.back())) run.push_back(i);
        else rest.push_back(i);
    }
    std::sort(rest.begin(), rest.end(), less);
    std::merge(run.begin(), run.end(), rest.begin(), rest.end(), idx.begin(), less);
    coff[0] = 0;
    for (int k = 0; k < A; k++) coff[(size_t)k + 1] = coff[(size_t)k] + meta[idx[(size_t)k]].x;
    const bool zero_r = !w->last_was_step || w->last_iterations == 0;  // r1/r2 are only stored by apply_impulse (arbiter.rs:236-237)
    auto fill = [&](int k0, int k1) {
        for (int k = k0; k < k1; k++) {
            const int a = idx[(size_t)k];
            SyltArbiterInfo& o = arbiters[k];
            o.body1 = bodies[a].x; o.body2 = bodies[a].y;
            o.id1 = w->hb[(size_t)o.body1].id; o.id2 = w->hb[(size_t)o.body2].id;
            o.friction = fr[a]; o.num_contacts = meta[a].x; o.contact_offset = coff[(size_t)k]; o.color = col[a];
            for (int q = 0; q < meta[a].x; q++) {
                const size_t ci = (size_t)(meta[a].y + q);
                SyltContact& c = contacts[coff[(size_t)k] + q];
                c.px = cd[ci].x; c.py = cd[ci].y; c.nx = ca[ci].x; c.ny = ca[ci].y;
                c.r1x = zero_r ? 0.0f : ca[ci].z; c.r1y = zero_r ? 0.0f : ca[ci].w; c.r2x = zero_r ? 0.0f : cb[ci].x; c.r2y = zero_r ? 0.0f : cb[ci].y;
                c.separation = cc[ci].w; c.pn = cc[ci].y; c.pt = cc[ci].z; c.pnb = cd[ci].z;
                c.mass_normal = cb[ci].z; c.mass_tangent = cb[ci].w; c.bias = cc[ci].x;
                uint32_t ft;
                memcpy(&ft, &cd[ci].w, 4);
                c.in_edge_1 = ft & 0xff; c.out_edge_1 = (ft >> 8) & 0xff; c.in_edge_2 = (ft >> 16) & 0xff; c.out_edge_2 = (ft >> 24) & 0xff;
                c.feature_value = 0;
            }
        }
    };
    const int nthreads = A >= 32768 ? 8 : 1;  // (a 100k-body world has ~230k arbiters: the assembly is worth a few host threads)
    if (nthreads == 1) {
        fill(0, A);
    } else {
        std::vector<std::thread> th;
        for (int t = 0; t < nthreads; t++) th.emplace_back(fill, (int)((long long)A * t / nthreads), (int)((long long)A * (t + 1) / nthreads));
        for (auto& t : th) t.join();
    }
    return A;
}

int sylt_world_get_joints(SyltWorld* w, SyltJointState* out, int32_t cap) {
    if (!w) return -SYLT_ERR_INVALID;
    int e = flush(w);
    if (e) return -e;
    int J = std::min<int>(cap, (int)w->hj.size());
    if (J == 0) return 0;
    std::vector<float2> p((size_t)J), bs((size_t)J);
    std::vector<float4> r((size_t)J), m((size_t)J);
    cudaMemcpyAsync(p.data(), w->jp.p, (size_t)J * sizeof(float2), cudaMemcpyDeviceToHost, w->stream);
    cudaMemcpyAsync(bs.data(), w->jbias.p, (size_t)J * sizeof(float2), cudaMemcpyDeviceToHost, w->stream);
    cudaMemcpyAsync(r.data(), w->jr.p, (size_t)J * sizeof(float4), cudaMemcpyDeviceToHost, w->stream);
    cudaMemcpyAsync(m.data(), w->jm.p, (size_t)J * sizeof(float4), cudaMemcpyDeviceToHost, w->stream);
    cudaError_t ce = cudaStreamSynchronize(w->stream);
    if (ce != cudaSuccess) { set_last_error(cudaGetErrorString(ce)); return -SYLT_ERR_CUDA; }
    const bool stepped = w->have_last && w->last_was_step;
    for (int j = 0; j < J; j++) {
        const HostJoint& h = w->hj[(size_t)j];
        SyltJointState s;
        memset(&s, 0, sizeof s);
        if (stepped) {
            s.px = p[(size_t)j].x; s.py = p[(size_t)j].y; s.bias_x = bs[(size_t)j].x; s.bias_y = bs[(size_t)j].y;
            s.r1x = r[(size_t)j].x; s.r1y = r[(size_t)j].y; s.r2x = r[(size_t)j].z; s.r2y = r[(size_t)j].w;
            s.m11 = m[(size_t)j].x; s.m12 = m[(size_t)j].y; s.m21 = m[(size_t)j].z; s.m22 = m[(size_t)j].w;
        }
        s.bias_factor = h.bias_factor; s.softness = h.softness; s.la1x = h.la1x; s.la1y = h.la1y; s.la2x = h.la2x; s.la2y = h.la2y;
        s.body1 = h.b1; s.body2 = h.b2;
        out[j] = s;
    }
    return J;
}

int sylt_world_get_joint_order(SyltWorld* w, int32_t* out, int32_t cap) {
    if (!w) return -SYLT_ERR_INVALID;
    int e = flush(w);
    if (e) return -e;
    int J = std::min<int>(cap, (int)w->h_jorder.size());
    for (int j = 0; j < J; j++) out[j] = w->h_jorder[(size_t)j];
    return J;
}

int sylt_world_get_stats(SyltWorld* w, SyltStepStats* o) {
    if (!w || !o) return SYLT_ERR_INVALID;
    memset(o, 0, sizeof *o);
    o->bodies = (int)w->hb.size();
    o->joints = (int)w->hj.size();
    o->joint_colors = w->n_joint_colors;
    if (w->have_last) {
        o->candidate_pairs = w->last.n_pairs; o->arbiters = w->last.n_arb; o->contacts = w->last.n_con;
        o->colors = w->last.n_colors; o->color_rounds = w->last.rounds;
    }
    return SYLT_OK;
}

int sylt_world_last_error_matrix(SyltWorld* w, float* o) {
    if (!w || !o) return -1;
    for (int k = 0; k < 4; k++) o[k] = w->last.badk[k];
    return w->last.err == SYLT_ERR_NO_INVERSE ? w->last.err_joint : -1;
}

int sylt_world_broad_phase(SyltWorld* w) { return do_steps(w, 0.0f, 1, false, true); }
int sylt_world_step(SyltWorld* w, float dt) { return do_steps(w, dt, 1, true, true); }
int sylt_world_step_n(SyltWorld* w, float dt, int32_t n) { return n < 0 ? SYLT_ERR_INVALID : (n == 0 ? SYLT_OK : do_steps(w, dt, n, true, true)); }
int sylt_world_step_n_async(SyltWorld* w, float dt, int32_t n) { return n < 0 ? SYLT_ERR_INVALID : (n == 0 ? SYLT_OK : do_steps(w, dt, n, true, false)); }
int sylt_world_sync(SyltWorld* w) {
    if (!w) return SYLT_ERR_INVALID;
    if (cudaSetDevice(w->device) != cudaSuccess) return SYLT_ERR_CUDA;
    return finish(w);
}
int sylt_world_timer_start(SyltWorld* w) {
    if (!w) return SYLT_ERR_INVALID;
    SYLT_CUDA(cudaSetDevice(w->device));
    SYLT_CUDA(cudaEventRecord(w->ev0, w->stream));
    return SYLT_OK;
}
int sylt_world_timer_stop(SyltWorld* w, float* ms) {
    if (!w || !ms) return SYLT_ERR_INVALID;
    SYLT_CUDA(cudaSetDevice(w->device));
    SYLT_CUDA(cudaEventRecord(w->ev1, w->stream));
    SYLT_CUDA(cudaEventSynchronize(w->ev1));
    SYLT_CUDA(cudaEventElapsedTime(ms, w->ev0, w->ev1));
    return SYLT_OK;
}
int64_t sylt_world_launch_count(SyltWorld* w) { return w ? w->launches : 0; }

int sylt_world_set_kernel_timing(SyltWorld* w, int on) {
    if (!w) return SYLT_ERR_INVALID;
    w->kernel_timing = on != 0;
    if (!on) w->kt_n = 0;
    return SYLT_OK;
}
int sylt_world_get_kernel_times(SyltWorld* w, float* out_us, int32_t cap) {
    if (!w || (cap > 0 && !out_us)) return -SYLT_ERR_INVALID;
    if (cudaSetDevice(w->device) != cudaSuccess) return -SYLT_ERR_CUDA;
    if (w->inflight) { const int e = finish(w); if (e) return -e; }
    int n = 0;
    for (int k = 1; k < w->kt_n && n < cap; k++, n++) {
        float ms = 0.0f;
        if (cudaEventElapsedTime(&ms, w->kt_ev[k - 1], w->kt_ev[k]) != cudaSuccess) return -SYLT_ERR_CUDA;
        out_us[n] = ms * 1e3f;
    }
    return n;
}

int sylt_collide_pairs(int device, const SyltBodyDesc* a, const SyltBodyDesc* b, int32_t n, const float* verts_xy, int32_t n_verts,
                       SyltContact* out, int32_t cap, int32_t* out_counts) {
    if (n < 0 || cap < 0 || (n > 0 && (!a || !b || !out || !out_counts))) return SYLT_ERR_INVALID;
    if (n == 0) return SYLT_OK;
    SYLT_CUDA(cudaSetDevice(device));
    std::vector<V2> lv;
    std::vector<int> loff((size_t)n * 4);
    for (int i = 0; i < n; i++) {
        HostBody ha, hb2;
        int e = make_host_body(a[i], verts_xy, n_verts, &ha, &lv);
        if (e) return e;
        e = make_host_body(b[i], verts_xy, n_verts, &hb2, &lv);
        if (e) return e;
        loff[(size_t)i * 4] = ha.vert_offset; loff[(size_t)i * 4 + 1] = ha.nverts;
        loff[(size_t)i * 4 + 2] = hb2.vert_offset; loff[(size_t)i * 4 + 3] = hb2.nverts;
    }
    SyltBodyDesc *da = nullptr, *db = nullptr;
    float2* dl = nullptr;
    int *dof = nullptr, *dcnt = nullptr;
    SyltContact* dout = nullptr;
    SYLT_CUDA(cudaMalloc(&da, (size_t)n * sizeof *da));
    SYLT_CUDA(cudaMalloc(&db, (size_t)n * sizeof *db));
    SYLT_CUDA(cudaMalloc(&dl, (lv.size() + 1) * sizeof(float2)));
    SYLT_CUDA(cudaMalloc(&dof, loff.size() * sizeof(int)));
    SYLT_CUDA(cudaMalloc(&dcnt, (size_t)n * sizeof(int)));
    SYLT_CUDA(cudaMalloc(&dout, ((size_t)n * cap + 1) * sizeof(SyltContact)));
    SYLT_CUDA(cudaMemcpy(da, a, (size_t)n * sizeof *da, cudaMemcpyHostToDevice));
    SYLT_CUDA(cudaMemcpy(db, b, (size_t)n * sizeof *db, cudaMemcpyHostToDevice));
    if (!lv.empty()) SYLT_CUDA(cudaMemcpy(dl, lv.data(), lv.size() * sizeof(float2), cudaMemcpyHostToDevice));
    SYLT_CUDA(cudaMemcpy(dof, loff.data(), loff.size() * sizeof(int), cudaMemcpyHostToDevice));
    k_collide_pairs<<<(n + 127) / 128, 128>>>(n, da, db, dl, dof, dout, cap, dcnt);
    SYLT_CUDA(cudaGetLastError());
    SYLT_CUDA(cudaMemcpy(out_counts, dcnt, (size_t)n * sizeof(int), cudaMemcpyDeviceToHost));
    SYLT_CUDA(cudaMemcpy(out, dout, (size_t)n * cap * sizeof(SyltContact), cudaMemcpyDeviceToHost));
    cudaFree(da); cudaFree(db); cudaFree(dl); cudaFree(dof); cudaFree(dcnt); cudaFree(dout);
    return SYLT_OK;
}

int sylt_sincos(int device, const float* angles, int32_t n, float* out_sin, float* out_cos) {
    if (n < 0 || (n > 0 && (!angles || !out_sin || !out_cos))) return SYLT_ERR_INVALID;
    if (n == 0) return SYLT_OK;
    SYLT_CUDA(cudaSetDevice(device));
    float *da, *ds, *dc;
    SYLT_CUDA(cudaMalloc(&da, (size_t)n * 4));
    SYLT_CUDA(cudaMalloc(&ds, (size_t)n * 4));
    SYLT_CUDA(cudaMalloc(&dc, (size_t)n * 4));
    SYLT_CUDA(cudaMemcpy(da, angles, (size_t)n * 4, cudaMemcpyHostToDevice));
    k_sincos<<<(n + 255) / 256, 256>>>(n, da, ds, dc);
    SYLT_CUDA(cudaGetLastError());
    SYLT_CUDA(cudaMemcpy(out_sin, ds, (size_t)n * 4, cudaMemcpyDeviceToHost));
    SYLT_CUDA(cudaMemcpy(out_cos, dc, (size_t)n * 4, cudaMemcpyDeviceToHost));
    cudaFree(da); cudaFree(ds); cudaFree(dc);
    return SYLT_OK;
}

}  // extern "C"
