// common.h — host-side helpers shared by world.cu and batch.cu (error plumbing, device buffers, and the
// host restatement of Body::new / Body::new_polygon mass properties, /root/reference/src/body.rs:44-146,204-284).
#pragma once
#include <cuda_runtime.h>
#include <cfloat>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>
#include "../../include/sylt2d_b200.h"
#include "sylt_math.cuh"

namespace sylt {

void set_last_error(const std::string& s);

#define SYLT_CUDA(call)                                                                        \
    do {                                                                                       \
        cudaError_t e__ = (call);                                                              \
        if (e__ != cudaSuccess) {                                                              \
            char buf__[512];                                                                   \
            snprintf(buf__, sizeof buf__, "%s:%d: %.300s -> %s", __FILE__, __LINE__, #call, cudaGetErrorString(e__)); \
            sylt::set_last_error(buf__);                                                       \
            return SYLT_ERR_CUDA;                                                              \
        }                                                                                      \
    } while (0)

template <typename T>
struct DBuf {
    T* p = nullptr;
    size_t cap = 0;
    // grow (preserving the first `keep` elements); returns cudaError_t
    cudaError_t ensure(size_t n, size_t keep = 0, cudaStream_t s = 0) {
        if (n <= cap) return cudaSuccess;
        size_t ncap = n + n / 2 + 16;
        T* q = nullptr;
        cudaError_t e = cudaMalloc(&q, ncap * sizeof(T));
        if (e != cudaSuccess) return e;
        if (keep && p) {
            e = cudaMemcpyAsync(q, p, keep * sizeof(T), cudaMemcpyDeviceToDevice, s);
            if (e != cudaSuccess) return e;
            cudaStreamSynchronize(s);
        }
        if (p) cudaFree(p);
        p = q;
        cap = ncap;
        return cudaSuccess;
    }
    void release() {
        if (p) cudaFree(p);
        p = nullptr;
        cap = 0;
    }
};

// Mass properties and geometry of one body, computed once on the host exactly as the reference constructors do.
struct HostBody {
    uint64_t id;
    int shape, nverts, vert_offset;
    int group;     // collision group (SyltBodyDesc.group): bodies of different groups never pair
    float width_x, width_y, mass, inv_mass, moi, inv_moi, friction;
    float radius;  // bounding radius about the body position (broad phase only; not in the reference)
};

// ConvexPolygon::get_vertex (+1 shift, body.rs:19-24)
inline V2 hv(const std::vector<V2>& v, long i) { long n = (long)v.size(); return v[(size_t)((i + n + 1) % n)]; }
// ConvexPolygon::centroid (body.rs:61-81)
inline V2 host_centroid(const std::vector<V2>& v) {
    float cx = 0.0f, cy = 0.0f, area = 0.0f;
    for (size_t i = 0; i < v.size(); i++) {
        V2 p1 = hv(v, (long)i), p2 = hv(v, (long)i + 1);
        float cr = p1.x * p2.y - p1.y * p2.x;
        area += cr;
        cx += (p1.x + p2.x) * cr;
        cy += (p1.y + p2.y) * cr;
    }
    area /= 2.0f;
    cx /= 6.0f * area;
    cy /= 6.0f * area;
    return mk(cx, cy);
}
// ConvexPolygon::moi (body.rs:95-121, quirk Q-9: area second moment, not divided by area)
inline float host_poly_moi(const std::vector<V2>& v) {
    V2 c = host_centroid(v);
    float moi = 0.0f;
    for (size_t i = 0; i < v.size(); i++) {
        V2 a = hv(v, (long)i), b = hv(v, (long)i + 1);
        V2 p1 = mk(a.x - c.x, a.y - c.y), p2 = mk(b.x - c.x, b.y - c.y);
        float cr = p1.x * p2.y - p1.y * p2.x;
        moi += cr * (p1.x * p1.x + p1.x * p2.x + p2.x * p2.x + p1.y * p1.y + p1.y * p2.y + p2.y * p2.y);
    }
    return fabsf(moi) / 12.0f;
}

// Fills mass props (Body::new body.rs:204-216 / Body::new_polygon :246-263) and appends the body's
// centroid-relative vertices (what ConvexPolygon::rotate subtracts every call, body.rs:148-158, quirk Q-12)
// to `local_verts`.  Boxes get their 4 stored vertices (body.rs:217-224) so Box-Polygon pairs work (quirk Q-11).
inline int make_host_body(const SyltBodyDesc& d, const float* verts_xy, int n_verts_total, HostBody* hb, std::vector<V2>* local_verts) {
    hb->id = d.id;
    hb->shape = d.shape;
    if (d.group < 0 || d.group > 0xffff) return SYLT_ERR_INVALID;
    hb->group = d.group;
    hb->mass = d.mass;
    hb->friction = d.friction;
    std::vector<V2> v;
    if (d.shape == SYLT_SHAPE_BOX) {
        hb->width_x = d.width_x;
        hb->width_y = d.width_y;
        if (d.mass < FLT_MAX) {
            hb->inv_mass = 1.0f / d.mass;
            hb->moi = d.mass * (d.width_x * d.width_x + d.width_y * d.width_y) / 12.0f;
            hb->inv_moi = 1.0f / hb->moi;
        } else {
            hb->inv_mass = 0.0f; hb->moi = FLT_MAX; hb->inv_moi = 0.0f;
        }
        float hw = d.width_x / 2.0f, hh = d.width_y / 2.0f;
        v = {mk(hw, hh), mk(-hw, hh), mk(-hw, -hh), mk(hw, -hh)};
    } else if (d.shape == SYLT_SHAPE_POLYGON) {
        if (d.vert_count < 1 || d.vert_offset < 0 || d.vert_offset + d.vert_count > n_verts_total || !verts_xy) return SYLT_ERR_INVALID;
        if (d.vert_count > SYLT_MAX_POLY_VERTS) return SYLT_ERR_UNSUPPORTED;
        for (int i = 0; i < d.vert_count; i++) v.push_back(mk(verts_xy[2 * (d.vert_offset + i)], verts_xy[2 * (d.vert_offset + i) + 1]));
        if (d.mass < FLT_MAX) {
            hb->inv_mass = 1.0f / d.mass;
            hb->moi = d.mass * host_poly_moi(v);
            hb->inv_moi = 1.0f / hb->moi;
        } else {
            hb->inv_mass = 0.0f; hb->moi = FLT_MAX; hb->inv_moi = 0.0f;
        }
        float min_x = FLT_MAX, max_x = -FLT_MAX, min_y = FLT_MAX, max_y = -FLT_MAX;  // bounding_box (body.rs:124-146)
        for (V2 p : v) {
            if (p.x < min_x) min_x = p.x;
            if (p.x > max_x) max_x = p.x;
            if (p.y < min_y) min_y = p.y;
            if (p.y > max_y) max_y = p.y;
        }
        hb->width_x = max_x - min_x;
        hb->width_y = max_y - min_y;
    } else {
        return SYLT_ERR_INVALID;
    }
    V2 c = host_centroid(v);
    hb->nverts = (int)v.size();
    hb->vert_offset = (int)local_verts->size();
    float r2 = 0.0f;
    for (V2 p : v) {
        V2 l = mk(p.x - c.x, p.y - c.y);
        local_verts->push_back(l);
        float q = l.x * l.x + l.y * l.y;
        if (q > r2) r2 = q;
    }
    float r = sqrtf(r2);
    if (d.shape == SYLT_SHAPE_BOX) {
        float rb = 0.5f * sqrtf(d.width_x * d.width_x + d.width_y * d.width_y);
        if (rb > r || !(r == r)) r = rb;
    }
    hb->radius = (r == r) ? r : 0.0f;
    return SYLT_OK;
}

// exists i<j with id_i > id_j and not both static  (the pair the reference's Arbiter::new would panic on, quirk Q-6)
inline bool body_order_violated(const std::vector<HostBody>& b) {
    uint64_t max_all = 0, max_dyn = 0;
    bool any = false, any_dyn = false;
    for (const HostBody& x : b) {
        bool dyn = x.inv_mass != 0.0f;
        if (dyn && any && max_all > x.id) return true;
        if (any_dyn && max_dyn > x.id) return true;
        if (!any || x.id > max_all) max_all = x.id;
        any = true;
        if (dyn && (!any_dyn || x.id > max_dyn)) max_dyn = x.id;
        any_dyn = any_dyn || dyn;
    }
    return false;
}

}  // namespace sylt
