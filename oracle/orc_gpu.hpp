// ORACLE — TEST INFRASTRUCTURE ONLY (see orc_math.hpp header).
//
// Restatement of the `strolle-gpu` crate (L1 shared device library): rays, BVH
// traversal, triangles, hits, G-buffer packing, camera, noise, reservoirs,
// MIS, lights, BRDFs, atmosphere sampling.  Each item cites the reference
// file:line (relative to /root/reference) it follows.
#pragma once
#include <vector>
#ifdef _OPENMP
#include <omp.h>
#endif
#include "orc_math.hpp"

namespace orc {

// ---------------------------------------------------------------------------
// Scene views (strolle-gpu/src/{bvh_view,triangles,materials,lights,world}.rs)
// ---------------------------------------------------------------------------
struct Material {  // strolle-gpu/src/material.rs:7-21 (112 B)
    V4 base_color, base_color_texture, emissive, emissive_texture;
    float roughness, metallic, reflectance, ior;
    V4 metallic_roughness_texture, normal_map_texture;
};
struct Light {  // strolle-gpu/src/light.rs:13-42 (112 B)
    V4 d0, d1, d2, d3, prev_d0, prev_d1, prev_d2;
};
struct World {  // strolle-gpu/src/world.rs:6-13
    u32 light_count; float sun_azimuth, sun_altitude; u32 _pad;
};
struct Lut {   // a sampled Rgba16Float texture with a linear, clamp-to-edge sampler
    int w, h; const V4* texels;  // texel values already rounded to f16 precision
};
struct Scene {
    const V4* triangles;  // 9 vec4 per triangle (strolle-gpu/src/triangle.rs:8-21)
    const V4* bvh;        // strolle/src/bvh/serializer.rs:53-104
    size_t bvh_len = 0;   // vec4 count; 0 = no primitive alive: the serialiser emits nothing (serializer.rs:20-110) and the
                          // reference's rasteriser draws nothing (passes/prim_raster.rs:134-246), so every ray misses
    const Material* materials;
    const Light* lights;
    World world;
    const uint8_t* blue_noise;  // 256x256 RGBA8 (strolle/assets/blue-noise.png)
    Lut transmittance_lut, sky_lut;
    const u32* tri_instance;    // per triangle: index into instance_xforms (6 vec4 each), prim_raster.rs push constants
    const V4* instance_xforms;  // per instance: curr_xform_inv d0..d2, prev_xform d0..d2 (PrimRasterPassParams::encode_affine, passes.rs:54-77)
    const uint8_t* atlas;       // ATLAS_SIZE^2 RGBA8 (Rgba8UnormSrgb), or null when no image was inserted
    const float* srgb_lut;      // 256-entry sRGB -> linear table (hardware decode of the atlas format)
};
static const u32 ATLAS_SIZE = 8192;   // strolle/src/images.rs:29-30

static const u32 BVH_STACK_SIZE = 24;  // strolle-gpu/src/lib.rs:76

// World::sun_dir / sun_pos (strolle-gpu/src/world.rs:19-29)
static inline V3 world_sun_dir(const World& w) {
    return v3(cos_(w.sun_altitude) * sin_(w.sun_azimuth), sin_(w.sun_altitude), -cos_(w.sun_altitude) * cos_(w.sun_azimuth));
}

// ---------------------------------------------------------------------------
// Normal (strolle-gpu/src/normal.rs:9-34)
// ---------------------------------------------------------------------------
static inline V2 normal_encode(V3 n) {
    n = n / (abs_(n.x) + abs_(n.y) + abs_(n.z));
    V2 r;
    if (n.z >= 0.0f) r = v2(n.x, n.y);
    else {
        V2 t = v2(1.0f - abs_(n.y), 1.0f - abs_(n.x));
        t.x = copysign_(t.x, n.x);
        t.y = copysign_(t.y, n.y);
        r = t;
    }
    return r * 0.5f + v2(0.5f, 0.5f);
}
static inline V3 normal_decode(V2 e) {
    V2 n2 = e * 2.0f - v2(1.0f, 1.0f);
    V3 n = v3(n2.x, n2.y, 1.0f - abs_(n2.x) - abs_(n2.y));
    float t = fmax_(-n.z, 0.0f);
    n.x -= copysign_(t, n.x);
    n.y -= copysign_(t, n.y);
    return normalize(n);
}

// ---------------------------------------------------------------------------
// TriangleHit (strolle-gpu/src/hit.rs:75-129)
// ---------------------------------------------------------------------------
struct TriangleHit {
    float distance; V3 point; V3 normal; V2 uv; u32 material_id;
    // bookkeeping that is not in the reference struct: the id of the accepted
    // triangle (parity tests compare it bit-exactly)
    u32 triangle_id;
};
static inline TriangleHit trihit_none() {
    TriangleHit h; h.distance = F32_MAX; h.point = v3s(0); h.normal = v3s(0); h.uv = v2(0, 0); h.material_id = 0; h.triangle_id = 0xffffffffu; return h;
}
static inline bool trihit_is_some(const TriangleHit& h) { return h.distance < F32_MAX; }
static inline void trihit_pack(const TriangleHit& h, V4* d0, V4* d1) {  // hit.rs:112-120
    *d0 = v4(h.point, u2f(h.material_id));
    V2 n = normal_encode(h.normal);
    *d1 = v4(n.x, n.y, h.uv.x, h.uv.y);
}
static inline TriangleHit trihit_unpack(V4 d0, V4 d1) {  // hit.rs:95-110
    if (xyz(d0) == v3s(0)) return trihit_none();
    TriangleHit h;
    h.distance = 0.0f; h.point = xyz(d0); h.normal = normal_decode(v2(d1.x, d1.y)); h.uv = v2(d1.z, d1.w);
    h.material_id = f2u(d0.w); h.triangle_id = 0xffffffffu;
    return h;
}

// ---------------------------------------------------------------------------
// Ray (strolle-gpu/src/ray.rs)
// ---------------------------------------------------------------------------
struct Ray { V3 origin, dir, inv_dir; float len; };
static inline Ray ray_default() { Ray r; r.origin = v3s(0); r.dir = v3s(0); r.inv_dir = v3s(0); r.len = 0.0f; return r; }
static inline Ray ray_new(V3 origin, V3 dir) {  // ray.rs:22-30
    Ray r; r.origin = origin; r.dir = dir; r.inv_dir = 1.0f / dir; r.len = F32_MAX; return r;
}
static inline Ray ray_with_len(Ray r, float len) { r.len = len; return r; }
static inline V3 ray_at(const Ray& r, float t) { return r.origin + r.dir * t; }

// ray.rs:273-302
static inline float ray_intersect_box(const Ray& r, V3 bmin, V3 bmax) {
    float tmin = 0.0f, tmax = F32_MAX;
    V3 t1 = (bmin - r.origin) * r.inv_dir;
    V3 t2 = (bmax - r.origin) * r.inv_dir;
    tmin = fmax_(tmin, fmin_(t1.x, t2.x)); tmax = fmin_(tmax, fmax_(t1.x, t2.x));
    tmin = fmax_(tmin, fmin_(t1.y, t2.y)); tmax = fmin_(tmax, fmax_(t1.y, t2.y));
    tmin = fmax_(tmin, fmin_(t1.z, t2.z)); tmax = fmin_(tmax, fmax_(t1.z, t2.z));
    return (tmin <= tmax) ? tmin : F32_MAX;
}
// ray.rs:304-322
static inline float ray_intersect_sphere(const Ray& r, float radius) {
    float b = dot(r.origin, r.dir);
    float c = dot(r.origin, r.origin) - radius * radius;
    if (c > 0.0f && b > 0.0f) return -1.0f;
    float discr = b * b - c;
    if (discr < 0.0f) return -1.0f;
    else if (discr > b * b) return -b + sqrt_(discr);
    else return -b - sqrt_(discr);
}

// Triangle::hit — Möller–Trumbore (strolle-gpu/src/triangle.rs:64-113)
static inline bool triangle_hit(const V4* t, const Ray& ray, TriangleHit* hit) {
    V3 p0 = xyz(t[0]), p1 = xyz(t[3]), p2 = xyz(t[6]);
    V3 v0v1 = p1 - p0, v0v2 = p2 - p0;
    V3 pvec = cross(ray.dir, v0v2);
    float det = dot(v0v1, pvec);
    if (abs_(det) < F32_EPSILON) return false;
    float inv_det = 1.0f / det;
    V3 tvec = ray.origin - p0;
    float u = dot(tvec, pvec) * inv_det;
    V3 qvec = cross(tvec, v0v1);
    float v = dot(ray.dir, qvec) * inv_det;
    float distance = dot(v0v2, qvec) * inv_det;
    if ((u < 0.0f) | (u > 1.0f) | (v < 0.0f) | (u + v > 1.0f) | (distance <= 0.0f) | (distance >= hit->distance)) return false;
    V3 n0 = xyz(t[1]), n1 = xyz(t[4]), n2 = xyz(t[7]);
    V3 normal = u * n1 + v * n2 + (1.0f - u - v) * n0;
    normal = normalize(normal) * copysign_(1.0f, inv_det);
    V2 uv0 = v2(t[0].w, t[1].w), uv1 = v2(t[3].w, t[4].w), uv2_ = v2(t[6].w, t[7].w);
    V2 uv = uv0 + (uv1 - uv0) * u + (uv2_ - uv0) * v;
    hit->uv = uv; hit->normal = normal; hit->distance = distance;
    return true;
}

// Material::sample_atlas (strolle-gpu/src/material.rs:76-104).  The atlas sampler is wgpu's default
// (nearest filter, clamp-to-edge, strolle/src/images.rs:38-43); the texture format Rgba8UnormSrgb decodes
// r,g,b through the sRGB transfer function (table) and alpha linearly.
static inline float wrap_uv(float t) { return (t > 0.0f) ? fmod_(t, 1.0f) : 1.0f - fmod_(-t, 1.0f); }
static inline V4 atlas_fetch(const Scene& sc, V2 uv) {
    if (!sc.atlas) return v4z();
    i32 x = f2i_sat(floor_(uv.x * (float)ATLAS_SIZE)), y = f2i_sat(floor_(uv.y * (float)ATLAS_SIZE));
    if (x < 0) x = 0; if (x > (i32)ATLAS_SIZE - 1) x = (i32)ATLAS_SIZE - 1;
    if (y < 0) y = 0; if (y > (i32)ATLAS_SIZE - 1) y = (i32)ATLAS_SIZE - 1;
    const uint8_t* p = sc.atlas + 4 * ((size_t)y * ATLAS_SIZE + (size_t)x);
    return v4(sc.srgb_lut[p[0]], sc.srgb_lut[p[1]], sc.srgb_lut[p[2]], (float)p[3] / 255.0f);
}
static inline V4 material_sample_atlas(const Scene& sc, V2 hit_uv, V4 multiplier, V4 texture) {
    if (is_zero(texture)) return multiplier;
    hit_uv.x = wrap_uv(hit_uv.x); hit_uv.y = wrap_uv(hit_uv.y);
    V2 uv = v2(texture.x, texture.y) + hit_uv * v2(texture.z, texture.w);
    return multiplier * atlas_fetch(sc, uv);
}
static inline V4 material_base_color(const Scene& sc, const Material& m, V2 uv) { return material_sample_atlas(sc, uv, m.base_color, m.base_color_texture); }
static inline V3 material_emissive(const Scene& sc, const Material& m, V2 uv) { return xyz(material_sample_atlas(sc, uv, m.emissive, m.emissive_texture)); }
// Material::metallic_roughness (material.rs:44-58): (metallic, roughness) = (1, roughness, metallic, 1) * texel -> .zy()
static inline V2 material_metallic_roughness(const Scene& sc, const Material& m, V2 uv) {
    V4 t = material_sample_atlas(sc, uv, v4(1.0f, m.roughness, m.metallic, 1.0f), m.metallic_roughness_texture);
    return v2(t.z, t.y);
}
static inline void material_regularize(Material& m) { m.roughness = fmax_(m.roughness, 0.75f * 0.75f); }  // material.rs:25-27

enum Tracing { ReturnClosest, ReturnFirst };

// executed Ray::trace / Ray::intersect calls (the Mrays/s numerator); one padded slot per OpenMP thread
struct RayCounter { unsigned long long n; char pad[56]; };
static RayCounter g_ray_counters[256];
static inline void count_ray() {
#ifdef _OPENMP
    g_ray_counters[omp_get_thread_num() & 255].n += 1;
#else
    g_ray_counters[0].n += 1;
#endif
}

// Ray::traverse (strolle-gpu/src/ray.rs:114-266).  Returns `used_memory`.
// `overflow` (not in the reference) is set when the 24-entry stack would
// overflow — the reference silently corrupts a neighbour's stack there
// (strolle-gpu/src/lib.rs:72-76); both oracle and product assert at upload.
static inline size_t ray_traverse(const Ray& self, const Scene& sc, Tracing tracing, TriangleHit* hit, u32* visited_nodes = nullptr) {
    count_ray();
    if (sc.bvh_len == 0) { if (visited_nodes) *visited_nodes = 0; return 0; }   // empty scene: a miss, nothing touched
    size_t used_memory = 0;
    u32 bvh_ptr = 0;
    u32 stack[BVH_STACK_SIZE];
    u32 stack_ptr = 0;
    u32 visits = 0;
    for (;;) {
        used_memory += 16;
        visits++;
        V4 d0 = sc.bvh[bvh_ptr];
        bool is_internal = f2u(d0.w) == 0;
        if (is_internal) {
            used_memory += 3 * 16;
            V4 d1 = sc.bvh[bvh_ptr + 1], d2 = sc.bvh[bvh_ptr + 2], d3 = sc.bvh[bvh_ptr + 3];
            u32 near_ptr = bvh_ptr + 4;
            u32 far_ptr = f2u(d1.w);
            float near_d = ray_intersect_box(self, xyz(d0), xyz(d1));
            float far_d = ray_intersect_box(self, xyz(d2), xyz(d3));
            if (far_d < near_d) { u32 t = near_ptr; near_ptr = far_ptr; far_ptr = t; float f = near_d; near_d = far_d; far_d = f; }
            if (far_d < hit->distance) { if (stack_ptr < BVH_STACK_SIZE) stack[stack_ptr] = far_ptr; stack_ptr += 1; }
            if (near_d < hit->distance) { bvh_ptr = near_ptr; continue; }
        } else {
            used_memory += 144;
            u32 flags = f2u(d0.x);
            bool got_more = (flags & 1u) == 1u;
            bool has_alpha = (flags & 2u) == 2u;
            u32 triangle_id = f2u(d0.y), material_id = f2u(d0.z);
            V2 prev_uv = hit->uv; V3 prev_normal = hit->normal; float prev_distance = hit->distance;
            bool found = triangle_hit(sc.triangles + 9 * (size_t)triangle_id, self, hit);
            if (found && has_alpha) {
                used_memory += 112; used_memory += 16;
                V4 base = material_base_color(sc, sc.materials[material_id], hit->uv);
                if (base.w < 1.0f) { found = false; hit->uv = prev_uv; hit->normal = prev_normal; hit->distance = prev_distance; }
            }
            if (found) {
                hit->material_id = material_id;
                hit->triangle_id = triangle_id;
                if (tracing == ReturnFirst) break;
            }
            if (got_more) { bvh_ptr += 1; continue; }
        }
        if (stack_ptr > 0) { stack_ptr -= 1; bvh_ptr = stack[stack_ptr]; }
        else break;
    }
    if (trihit_is_some(*hit)) hit->point = ray_at(self, hit->distance);
    if (visited_nodes) *visited_nodes = visits;
    return used_memory;
}
// Ray::trace (ray.rs:55-80)
static inline TriangleHit ray_trace(const Ray& r, const Scene& sc, size_t* used_memory = nullptr) {
    TriangleHit h = trihit_none();
    size_t um = ray_traverse(r, sc, ReturnClosest, &h);
    if (used_memory) *used_memory = um;
    return h;
}
// Ray::intersect (ray.rs:84-112)
static inline bool ray_intersect(const Ray& r, const Scene& sc) {
    TriangleHit h = trihit_none();
    h.distance = r.len;
    ray_traverse(r, sc, ReturnFirst, &h);
    return h.distance < r.len;
}

// ---------------------------------------------------------------------------
// GBufferEntry (strolle-gpu/src/gbuffer.rs:19-112)
// ---------------------------------------------------------------------------
struct GBufferEntry { V4 base_color; V3 normal; float metallic; V3 emissive; float roughness, reflectance, depth; };
static inline GBufferEntry gbuffer_default() { GBufferEntry g; g.base_color = v4z(); g.normal = v3s(0); g.metallic = 0; g.emissive = v3s(0); g.roughness = 0; g.reflectance = 0; g.depth = 0; return g; }
static inline GBufferEntry gbuffer_unpack(V4 d0, V4 d1) {
    GBufferEntry g;
    g.depth = d0.x;
    g.normal = normal_decode(v2(d0.y, d0.z));
    u32 b = f2u(d0.w);
    g.metallic = (float)(b & 0xff) / 255.0f;
    g.roughness = sqr((float)((b >> 8) & 0xff) / 255.0f);
    g.reflectance = (float)((b >> 16) & 0xff) / 255.0f;
    g.emissive = xyz(d1);
    u32 c = f2u(d1.w);
    g.base_color = v4(pow_((float)(c & 0xff) / 255.0f, 2.2f), pow_((float)((c >> 8) & 0xff) / 255.0f, 2.2f),
                      pow_((float)((c >> 16) & 0xff) / 255.0f, 2.2f), pow_((float)((c >> 24) & 0xff) / 63.0f, 2.2f));
    return g;
}
static inline void gbuffer_pack(const GBufferEntry& g, V4* d0, V4* d1) {
    V2 n = normal_encode(g.normal);
    float metallic = clampf(g.metallic, 0.0f, 1.0f) * 255.0f;
    float roughness = clampf(sqrt_(g.roughness), 0.0f, 1.0f) * 255.0f;
    float reflectance = clampf(g.reflectance, 0.0f, 1.0f) * 255.0f;
    *d0 = v4(g.depth, n.x, n.y, u2f(from_bytes(f2u_sat(metallic), f2u_sat(roughness), f2u_sat(reflectance), 1)));
    const float ig = 1.0f / 2.2f;
    V4 bc = v4(clampf(pow_(g.base_color.x, ig), 0.0f, 1.0f), clampf(pow_(g.base_color.y, ig), 0.0f, 1.0f),
               clampf(pow_(g.base_color.z, ig), 0.0f, 1.0f), clampf(pow_(g.base_color.w, ig), 0.0f, 1.0f));
    *d1 = v4(g.emissive.x, g.emissive.y, g.emissive.z,
             u2f(from_bytes(f2u_sat(bc.x * 255.0f), f2u_sat(bc.y * 255.0f), f2u_sat(bc.z * 255.0f), f2u_sat(bc.w * 63.0f))));
}
static inline bool gbuffer_is_some(const GBufferEntry& g) { return g.depth != 0.0f; }
static inline float gbuffer_clamped_roughness(const GBufferEntry& g) { return clampf(g.roughness, 0.089f * 0.089f, 1.0f); }

// ---------------------------------------------------------------------------
// Camera (strolle-gpu/src/camera.rs:8-106), 160 B
// ---------------------------------------------------------------------------
struct Camera { M4 projection_view, ndc_to_world; V4 origin, screen; };
static inline V4 camera_world_to_clip(const Camera& c, V3 p) { return mul(c.projection_view, v4(p, 1.0f)); }
static inline V2 camera_clip_to_screen(const Camera& c, V4 pos) {
    V2 ndc = v2(pos.x, pos.y) / pos.w;
    ndc = v2(ndc.x, -ndc.y);
    return (0.5f * ndc + v2(0.5f, 0.5f)) * v2(c.screen.x, c.screen.y);
}
static inline size_t camera_screen_to_idx(const Camera& c, UV2 p) { return (size_t)(p.y * f2u_sat(c.screen.x) + p.x); }
static inline bool camera_contains(const Camera& c, UV2 p) { return p.x < f2u_sat(c.screen.x) && p.y < f2u_sat(c.screen.y); }
static inline bool camera_contains(const Camera& c, IV2 p) { return p.x >= 0 && p.y >= 0 && p.x < f2i_sat(c.screen.x) && p.y < f2i_sat(c.screen.y); }
static inline bool camera_contains(const Camera& c, V2 p) { return p.x >= 0.0f && p.y >= 0.0f && p.x < c.screen.x && p.y < c.screen.y; }
static inline UV2 camera_contain(const Camera& c, IV2 pos) {  // camera.rs:57-77 (wrapping i32 arithmetic)
    i32 sx = f2i_sat(c.screen.x), sy = f2i_sat(c.screen.y);
    if (pos.x < 0) pos.x = (i32)(0u - (u32)pos.x);
    if (pos.y < 0) pos.y = (i32)(0u - (u32)pos.y);
    if (pos.x >= sx) pos.x = (i32)((u32)sx - (u32)pos.x + (u32)sx - 1u);
    if (pos.y >= sy) pos.y = (i32)((u32)sy - (u32)pos.y + (u32)sy - 1u);
    return uv2((u32)pos.x, (u32)pos.y);
}
static inline Ray camera_ray(const Camera& c, UV2 sp) {  // camera.rs:80-93
    V2 screen_size = v2(c.screen.x, c.screen.y);
    V2 p = v2((float)sp.x, (float)sp.y) + v2(0.5f, 0.5f);
    V2 ndc = p * 2.0f / screen_size - v2(1.0f, 1.0f);
    ndc = v2(ndc.x, -ndc.y);
    V3 far_plane = project_point3(c.ndc_to_world, v3(ndc.x, ndc.y, F32_EPSILON));
    V3 near_plane = project_point3(c.ndc_to_world, v3(ndc.x, ndc.y, 1.0f));
    return ray_new(near_plane, normalize(far_plane - near_plane));
}
static inline bool camera_is_eq(const Camera& a, const Camera& b) {  // camera.rs:103-106
    for (int i = 0; i < 4; i++) {
        const float* p = &a.projection_view.c[i].x; const float* q = &b.projection_view.c[i].x;
        for (int j = 0; j < 4; j++) if (!(abs_(p[j] - q[j]) <= 0.0025f)) return false;
    }
    return true;
}

// ---------------------------------------------------------------------------
// Hit (strolle-gpu/src/hit.rs:8-73), Surface (surface.rs)
// ---------------------------------------------------------------------------
struct Hit { V3 origin, dir, point; GBufferEntry gbuffer; };
static inline Hit hit_default() { Hit h; h.origin = v3s(0); h.dir = v3s(0); h.point = v3s(0); h.gbuffer = gbuffer_default(); return h; }
static inline Hit hit_new(const Ray& ray, const GBufferEntry& g) {
    Hit h; h.origin = ray.origin; h.dir = ray.dir; h.point = ray_at(ray, g.depth - 0.01f); h.gbuffer = g; return h;
}
static inline bool hit_is_some(const Hit& h) { return gbuffer_is_some(h.gbuffer); }

struct Surface { V3 normal; float depth, roughness; };
static inline Surface surface_get(const V4* tex, int w, UV2 p) {  // surface.rs:60-68
    V4 d0 = tex[(size_t)p.y * w + p.x];
    Surface s; s.normal = normal_decode(v2(d0.x, d0.y)); s.depth = d0.z; s.roughness = d0.w; return s;
}
static inline bool surface_is_sky(const Surface& s) { return s.depth == 0.0f; }
static inline float surface_similarity(const Surface& self, const Surface& other) {  // surface.rs:21-47
    if (surface_is_sky(self) || surface_is_sky(other)) return 0.0f;
    float d = fmax_(dot(self.normal, other.normal), 0.0f);
    float normal_score = (d <= 0.5f) ? 0.0f : 2.0f * d;
    float t = abs_(self.depth - other.depth);
    float depth_score = (t >= 0.1f * other.depth) ? 0.0f : 1.0f;
    return normal_score * depth_score;
}

// ---------------------------------------------------------------------------
// Reprojection (strolle-gpu/src/reprojection.rs) + BilinearFilter (utils/bilinear_filter.rs)
// ---------------------------------------------------------------------------
struct Reprojection { float prev_x, prev_y, confidence; u32 validity; };
static inline Reprojection reprojection_default() { Reprojection r = {0, 0, 0, 0}; return r; }
static inline V4 reprojection_serialize(const Reprojection& r) { return v4(r.prev_x, r.prev_y, r.confidence, u2f(r.validity)); }
static inline Reprojection reprojection_deserialize(V4 d) { Reprojection r = {d.x, d.y, d.z, f2u(d.w)}; return r; }
static inline bool reprojection_is_some(const Reprojection& r) { return r.confidence > 0.0f; }
static inline UV2 reprojection_prev_pos_round(const Reprojection& r) { return uv2(f2u_sat(round_(r.prev_x)), f2u_sat(round_(r.prev_y))); }
static inline bool reprojection_is_exact(const Reprojection& r) {
    // glam Vec2::fract = v - floor(v)
    V2 f = v2(r.prev_x - floor_(r.prev_x), r.prev_y - floor_(r.prev_y));
    return length_squared(f) == 0.0f;
}
static inline void reprojection_coords(float px, float py, IV2 out[4]) {  // bilinear_filter.rs:79-86
    out[0] = iv2(f2i_sat(floor_(px)), f2i_sat(floor_(py)));
    out[1] = iv2(f2i_sat(ceil_(px)), f2i_sat(floor_(py)));
    out[2] = iv2(f2i_sat(floor_(px)), f2i_sat(ceil_(py)));
    out[3] = iv2(f2i_sat(ceil_(px)), f2i_sat(ceil_(py)));
}
// BilinearFilter::reproject with sample = |pos| (tex[pos], 1.0) (bilinear_filter.rs:27-108)
static inline V4 bilinear_reproject(const Reprojection& r, const V4* tex, int w, int h) {
    if (reprojection_is_exact(r)) {
        UV2 p = reprojection_prev_pos_round(r);
        if ((int)p.x >= w || (int)p.y >= h) return v4z();  // guard: reference would read out of bounds
        return tex[(size_t)p.y * w + p.x];
    }
    IV2 p[4]; reprojection_coords(r.prev_x, r.prev_y, p);
    V4 s[4] = {v4z(), v4z(), v4z(), v4z()};
    float wt[4] = {0, 0, 0, 0};
    for (int i = 0; i < 4; i++) {
        if ((r.validity & (1u << i)) > 0 && p[i].x >= 0 && p[i].y >= 0) {
            if (p[i].x < w && p[i].y < h) s[i] = tex[(size_t)p[i].y * w + p[i].x];  // guard as above
            wt[i] = 1.0f;
        }
    }
    // Rust f32::fract = x - trunc(x)
    float ux = r.prev_x - trunc_(r.prev_x), uy = r.prev_y - trunc_(r.prev_y);
    V4 weights = v4(wt[0], wt[1], wt[2], wt[3]) * v4((1.0f - ux) * (1.0f - uy), ux * (1.0f - uy), (1.0f - ux) * uy, ux * uy);
    float w_sum = dot(weights, v4(1, 1, 1, 1));
    if (w_sum == 0.0f) return v4z();
    return (s[0] * weights.x + s[1] * weights.y + s[2] * weights.z + s[3] * weights.w) / w_sum;
}

// ---------------------------------------------------------------------------
// Noise (strolle-gpu/src/noise/{white,blue}.rs)
// ---------------------------------------------------------------------------
struct WhiteNoise { u32 state; };
static inline WhiteNoise wnoise_new(u32 seed, UV2 id) { WhiteNoise n; n.state = seed ^ (48619u * id.x) ^ (95461u * id.y); return n; }
static inline u32 wnoise_sample_int(WhiteNoise& n) {
    n.state = n.state * 747796405u + 2891336453u;
    u32 word = ((n.state >> ((n.state >> 28) + 4u)) ^ n.state) * 277803737u;
    return (word >> 22) ^ word;
}
static inline float wnoise_sample(WhiteNoise& n) { return (float)wnoise_sample_int(n) / 4294967296.0f; }  // u32::MAX as f32 == 2^32
static inline V2 wnoise_sample_circle(WhiteNoise& n) { float a = wnoise_sample(n) * PI * 2.0f; return v2(cos_(a), sin_(a)); }
static inline V2 wnoise_sample_disk(WhiteNoise& n) { float radius = sqrt_(wnoise_sample(n)); return wnoise_sample_circle(n) * radius; }
static inline V3 wnoise_sample_sphere(WhiteNoise& n) {
    float phi = wnoise_sample(n) * 2.0f * PI;
    float cos_theta = wnoise_sample(n) * 2.0f - 1.0f;
    float u = wnoise_sample(n);
    float theta = acos_(cos_theta);
    float r = sqrt_(u);
    return v3(r * sin_(theta) * cos_(phi), r * sin_(theta) * sin_(phi), r * cos_(theta));
}
static inline V3 wnoise_sample_hemisphere(WhiteNoise& n, V3 normal) {
    float cos_theta = wnoise_sample(n);
    float sin_theta = sqrt_(1.0f - sqr(cos_theta));
    float phi = 2.0f * PI * wnoise_sample(n);
    V3 t, b; any_orthonormal_pair(normal, &t, &b);
    return (t * cos_(phi) + b * sin_(phi)) * sin_theta + normal * cos_theta;
}
// BlueNoise::new + texel fetch (noise/blue.rs:15-27); RGBA8 unorm -> f32 = byte / 255
static inline V4 bnoise_texel(const uint8_t* tex, UV2 id, u32 frame) {
    u32 ux = (id.x + 71u * frame) % 256u, uy = (id.y + 11u * frame) % 256u;
    const uint8_t* p = tex + 4 * ((size_t)uy * 256 + ux);
    return v4((float)p[0] / 255.0f, (float)p[1] / 255.0f, (float)p[2] / 255.0f, (float)p[3] / 255.0f);
}

// ---------------------------------------------------------------------------
// BRDFs (strolle-gpu/src/brdf.rs)
// ---------------------------------------------------------------------------
struct BrdfSample { V3 dir; float pdf; V3 radiance; };
static inline V3 diffuse_brdf_eval(const GBufferEntry& g) { return xyz(g.base_color) * (1.0f - g.metallic) / PI; }
static inline float ggx_distribution(float n_dot_h, float roughness) {
    float a2 = roughness * roughness;
    float d = (n_dot_h * a2 - n_dot_h) * n_dot_h + 1.0f;
    return a2 / (PI * d * d);
}
static inline float ggx_schlick_masking_term(float n_dot_l, float n_dot_v, float roughness) {
    float k = roughness * roughness / 2.0f;
    float g_v = n_dot_v / (n_dot_v * (1.0f - k) + k);
    float g_l = n_dot_l / (n_dot_l * (1.0f - k) + k);
    return g_v * g_l;
}
static inline V3 ggx_schlick_fresnel(V3 f0, float l_dot_h) {
    float f90 = saturate(dot(f0, v3s(50.0f * 0.33f)));
    return f0 + (v3s(f90) - f0) * pow_(fmax_(1.0f - l_dot_h, 0.001f), 5.0f);
}
static inline V3 specular_brdf_eval(const GBufferEntry& g, V3 l, V3 v) {  // brdf.rs:46-79
    if (g.metallic <= 0.0f) return v3s(0);
    float a = gbuffer_clamped_roughness(g);
    V3 n = g.normal;
    V3 h = normalize(l + v);
    float n_dot_l = saturate(dot(n, l)), n_dot_h = saturate(dot(n, h)), l_dot_h = saturate(dot(l, h)), n_dot_v = saturate(dot(n, v));
    if (n_dot_l <= 0.0f || n_dot_v <= 0.0f) return v3s(0);
    float d = ggx_distribution(n_dot_h, a);
    float gg = ggx_schlick_masking_term(n_dot_l, n_dot_v, a);
    V3 f0 = v3s(0.16f * g.reflectance * g.reflectance * (1.0f - g.metallic)) + xyz(g.base_color) * g.metallic;
    V3 f = ggx_schlick_fresnel(f0, l_dot_h);
    return d * gg * f / (4.0f * n_dot_l * n_dot_v);
}
static inline BrdfSample diffuse_brdf_sample(const GBufferEntry& g, WhiteNoise& wn) {
    BrdfSample s; s.dir = wnoise_sample_hemisphere(wn, g.normal); s.pdf = 1.0f / PI; s.radiance = diffuse_brdf_eval(g); return s;
}
static inline BrdfSample specular_brdf_sample(const GBufferEntry& g, WhiteNoise& wn, V3 v) {  // brdf.rs:82-113
    float r0 = wnoise_sample(wn), r1 = wnoise_sample(wn);
    float a = gbuffer_clamped_roughness(g);
    V3 n = g.normal;
    float a2 = sqr(a);
    V3 b, t; any_orthonormal_pair(n, &b, &t);
    float cos_theta = sqrt_(fmax_(0.0f, (1.0f - r0) / ((a2 - 1.0f) * r0 + 1.0f)));
    float sin_theta = sqrt_(fmax_(0.0f, 1.0f - cos_theta * cos_theta));
    float phi = r1 * PI * 2.0f;
    V3 h = t * (sin_theta * cos_(phi)) + b * (sin_theta * sin_(phi)) + n * cos_theta;
    float n_dot_h = saturate(dot(n, h)), h_dot_v = saturate(dot(h, v));
    BrdfSample s;
    s.dir = normalize(2.0f * h_dot_v * h - v);
    s.pdf = ggx_distribution(n_dot_h, a) * n_dot_h / (4.0f * h_dot_v);
    s.radiance = specular_brdf_eval(g, s.dir, v);
    return s;
}
static inline BrdfSample layered_brdf_sample(const GBufferEntry& g, WhiteNoise& wn, V3 l) {  // brdf.rs:125-138
    BrdfSample s;
    if (wnoise_sample(wn) < g.metallic) { s = specular_brdf_sample(g, wn, l); s.pdf /= g.metallic; }
    else { s = diffuse_brdf_sample(g, wn); s.pdf /= 1.0f - g.metallic; }
    return s;
}

// ---------------------------------------------------------------------------
// Lights (strolle-gpu/src/light.rs)
// ---------------------------------------------------------------------------
struct LightRadiance { V3 radiance, diff_brdf, spec_brdf; };
static inline LightRadiance light_radiance_default() { LightRadiance r; r.radiance = v3s(0); r.diff_brdf = v3s(0); r.spec_brdf = v3s(0); return r; }
static inline V3 light_radiance_sum(const LightRadiance& r) { return r.radiance * (r.diff_brdf + r.spec_brdf); }
static inline V3 light_center(const Light& l) { return xyz(l.d0); }
static inline float light_radius(const Light& l) { return l.d0.w; }
static inline bool light_is_none(const Light& l) { return f2u(l.d2.x) == 0; }
static inline bool light_is_point(const Light& l) { return f2u(l.d2.x) == 1; }
static inline bool light_contains(const Light& l, V3 p) { return distance(light_center(l), p) <= light_radius(l); }
static inline bool light_is_slot_killed(const Light& l) { return f2u(l.d3.x) == 0xcafebabeu; }
static inline bool light_is_slot_remapped(const Light& l) { return f2u(l.d3.x) > 0 && f2u(l.d3.x) != 0xcafebabeu; }
static inline Light light_rollback(Light l) { l.d0 = l.prev_d0; l.d1 = l.prev_d1; l.d2 = l.prev_d2; return l; }
// glam 0.24.2 (crates.io, Cargo.lock; not vendored) `math::acos_approx`, restated from its published definition: DirectXMath's
// XMScalarACos, a 7th-degree minimax polynomial times sqrt(1 - |x|), mirrored for negative arguments.  `Vec3::angle_between`
// is acos_approx(dot / sqrt(|a|^2 |b|^2)); the only call site on the path is the spot-light cone (strolle-gpu/src/light.rs:149-152).
static inline float acos_approx(float v) {
    bool nonnegative = v >= 0.0f;
    float x = abs_(v);
    float omx = 1.0f - x;
    if (omx < 0.0f) omx = 0.0f;
    float root = sqrt_(omx);
    float result = ((((((-0.0012624911f * x + 0.0066700901f) * x - 0.0170881256f) * x + 0.0308918810f) * x - 0.0501743046f) * x + 0.0889789874f) * x - 0.2145988016f) * x + 1.5707963050f;
    result *= root;
    return nonnegative ? result : PI - result;
}
static inline float angle_between(V3 a, V3 b) { return acos_approx(dot(a, b) / sqrt_(length_squared(a) * length_squared(b))); }

static inline LightRadiance light_radiance(const Light& self, const Hit& hit) {  // light.rs:143-207
    V3 l = light_center(self) - hit.point;
    float f_angle;
    if (light_is_point(self)) f_angle = 1.0f;
    else {
        float angle = angle_between(normal_decode(v2(self.d2.y, self.d2.z)), hit.point - light_center(self));
        f_angle = saturate(1.0f - pow_(angle / self.d2.w, 3.0f));
    }
    float f_dist;
    float range = self.d1.w;
    if (range == F32_INF) f_dist = 1.0f;
    else {
        float l2 = length_squared(l);
        float inv_r2 = 1.0f / sqr(range);
        float factor = l2 * inv_r2;
        float smooth_factor = saturate(1.0f - factor * factor);
        float attenuation = smooth_factor * smooth_factor;
        f_dist = attenuation / fmax_(l2, 0.0001f);
    }
    float f_cosine = saturate(dot(hit.gbuffer.normal, normalize(l)));
    V3 diff_brdf = diffuse_brdf_eval(hit.gbuffer);
    V3 spec_brdf;
    {
        V3 v = -hit.dir;
        V3 n = hit.gbuffer.normal;
        V3 r = reflect(-v, n);
        V3 center_to_ray = dot(l, r) * r - l;
        V3 closest_point;
        {
            float t = light_radius(self) * (1.0f / sqrt_(dot(center_to_ray, center_to_ray)));
            closest_point = l + center_to_ray * saturate(t);
        }
        float l_spec_length_inverse = 1.0f / sqrt_(dot(closest_point, closest_point));
        float i_roughness;
        {
            float t = gbuffer_clamped_roughness(hit.gbuffer) + light_radius(self) * 0.5f * l_spec_length_inverse;
            i_roughness = gbuffer_clamped_roughness(hit.gbuffer) / saturate(t);
        }
        float intensity = sqr(i_roughness);
        V3 ll = closest_point * l_spec_length_inverse;
        spec_brdf = intensity * specular_brdf_eval(hit.gbuffer, ll, v);
    }
    LightRadiance out;
    out.radiance = xyz(self.d1) * f_angle * f_dist * f_cosine;
    out.diff_brdf = diff_brdf; out.spec_brdf = spec_brdf;
    return out;
}
static inline Ray light_ray_wnoise(const Light& self, WhiteNoise& wn, V3 hit_point) {  // light.rs:209-215
    V3 light_pos = light_center(self) + light_radius(self) * wnoise_sample_sphere(wn);
    V3 light_to_hit = hit_point - light_pos;
    return ray_with_len(ray_new(light_pos, normalize(light_to_hit)), length(light_to_hit));
}
static inline Ray light_ray_bnoise(const Light& self, V2 sample, V3 hit_point) {  // light.rs:217-239
    V3 to_light = light_center(self) - hit_point;
    V3 light_dir = normalize(to_light);
    float light_distance = length(to_light);
    float light_radius_ = light_radius(self) / light_distance;
    V3 tangent, bitangent; any_orthonormal_pair(light_dir, &tangent, &bitangent);
    float angle = 2.0f * PI * sample.x;
    float radius = sqrt_(sample.y);
    V2 disk_point = v2(sin_(angle), cos_(angle)) * radius * light_radius_;
    V3 ray_dir = light_dir + disk_point.x * tangent + disk_point.y * bitangent;
    ray_dir = normalize(ray_dir);
    return ray_with_len(ray_new(hit_point + ray_dir * light_distance, -ray_dir), light_distance);
}

// ---------------------------------------------------------------------------
// Atmosphere sampling (strolle-gpu/src/atmosphere.rs:86-205)
// ---------------------------------------------------------------------------
static const float ATM_GROUND_RADIUS_MM = 6.360f;
static const float ATM_ATMOSPHERE_RADIUS_MM = 6.460f;
static const float ATM_EXPOSURE = 20.0f;
static inline V3 atm_view_pos() { return v3(0.0f, ATM_GROUND_RADIUS_MM + 0.0002f, 0.0f); }

// sample_by_lod(linear sampler, clamp-to-edge, lod 0) restated as an explicit
// f32 bilinear fetch at texel centres (hardware filters use driver-defined
// fixed-point weights; the strict f32 form is normative here).
static inline V3 lut_sample(const Lut& lut, V2 uv) {
    float fx = uv.x * (float)lut.w - 0.5f, fy = uv.y * (float)lut.h - 0.5f;
    float x0f = floor_(fx), y0f = floor_(fy);
    float tx = fx - x0f, ty = fy - y0f;
    i32 x0 = f2i_sat(x0f), y0 = f2i_sat(y0f);
    i32 x1 = x0 + 1, y1 = y0 + 1;
    if (x0 < 0) x0 = 0; if (x0 > lut.w - 1) x0 = lut.w - 1;
    if (x1 < 0) x1 = 0; if (x1 > lut.w - 1) x1 = lut.w - 1;
    if (y0 < 0) y0 = 0; if (y0 > lut.h - 1) y0 = lut.h - 1;
    if (y1 < 0) y1 = 0; if (y1 > lut.h - 1) y1 = lut.h - 1;
    V3 a = xyz(lut.texels[(size_t)y0 * lut.w + x0]), b = xyz(lut.texels[(size_t)y0 * lut.w + x1]);
    V3 c = xyz(lut.texels[(size_t)y1 * lut.w + x0]), d = xyz(lut.texels[(size_t)y1 * lut.w + x1]);
    V3 top = a + (b - a) * tx, bot = c + (d - c) * tx;
    return top + (bot - top) * ty;
}
static inline V3 atm_sample_lut(const Lut& lut, V3 pos, V3 sun_dir) {  // atmosphere.rs:183-204
    float height = length(pos);
    V3 up = pos / height;
    float sun_cos_zenith = dot(sun_dir, up);
    float u = saturate(0.5f + 0.5f * sun_cos_zenith);
    float v = saturate((height - ATM_GROUND_RADIUS_MM) / (ATM_ATMOSPHERE_RADIUS_MM - ATM_GROUND_RADIUS_MM));
    return lut_sample(lut, v2(u, v));
}
static inline V3 atm_sample_sky_lut(const Scene& sc, V3 ray_dir, V3 sun_dir) {  // atmosphere.rs:108-146
    V3 vp = atm_view_pos();
    float height = length(vp);
    V3 up = vp / height;
    float horizon;
    { float t = sqr(height) - sqr(ATM_GROUND_RADIUS_MM); t = sqrt_(t) / height; horizon = acos_(clampf(t, -1.0f, 1.0f)); }
    float altitude = horizon - acos_(dot(ray_dir, up));
    float azimuth;
    if (abs_(altitude) > (0.5f * PI - 0.0001f)) azimuth = 0.0f;
    else {
        V3 right = cross(sun_dir, up);
        V3 forward = cross(up, right);
        V3 projected_dir = normalize(ray_dir - up * dot(ray_dir, up));
        float sin_theta = dot(projected_dir, right);
        float cos_theta = dot(projected_dir, forward);
        azimuth = atan2_(sin_theta, cos_theta) + PI;
    }
    float u = azimuth / (2.0f * PI);
    float v = 0.5f + 0.5f * copysign_(sqrt_(abs_(altitude) * 2.0f / PI), altitude);
    return lut_sample(sc.sky_lut, v2(u, v));
}
static inline V3 atm_evaluate_bloom(V3 ray_dir, V3 sun_dir) {  // atmosphere.rs:148-163
    const float SUN_SOLID_ANGLE = 0.53f * PI / 180.0f;
    float min_sun_cos_theta = cos_(SUN_SOLID_ANGLE);
    float cos_theta = dot(ray_dir, sun_dir);
    if (cos_theta >= min_sun_cos_theta) return v3s(1.0f);
    float offset = min_sun_cos_theta - cos_theta;
    float gaussian_bloom = exp_(-offset * 50000.0f) * 0.5f;
    float inv_bloom = 1.0f / (0.02f + offset * 300.0f) * 0.01f;
    return v3s(gaussian_bloom + inv_bloom);
}
static inline V3 atm_interpolate_bloom(V3 bloom) {  // atmosphere.rs:165-172
    V3 t = vclamp((bloom - v3s(0.002f)) / (v3s(1.0f) - v3s(0.002f)), v3s(0), v3s(1));
    return t * t * (v3s(3.0f) - 2.0f * t);
}
static inline V3 atmosphere_sample(const Scene& sc, V3 sun_dir, V3 ray_dir) {  // atmosphere.rs:86-106
    V3 lum = atm_sample_sky_lut(sc, ray_dir, sun_dir);
    V3 sun_lum = atm_evaluate_bloom(ray_dir, sun_dir);
    sun_lum = atm_interpolate_bloom(sun_lum);
    if (length_squared(sun_lum) > 0.0f) {
        Ray ray = ray_new(atm_view_pos(), ray_dir);
        if (ray_intersect_sphere(ray, ATM_GROUND_RADIUS_MM) >= 0.0f) sun_lum = v3s(0);
        else sun_lum *= atm_sample_lut(sc.transmittance_lut, atm_view_pos(), sun_dir);
    }
    lum += sun_lum;
    lum *= ATM_EXPOSURE;
    return lum;
}

// ---------------------------------------------------------------------------
// Reservoirs (strolle-gpu/src/reservoir.rs + reservoir/*.rs)
// ---------------------------------------------------------------------------
template <typename T> struct Reservoir {
    T sample; float m, w;
    bool update(WhiteNoise& wn, const T& s, float weight) {  // reservoir.rs:24-39
        m += 1.0f; w += weight;
        if (wnoise_sample(wn) * w < weight) { sample = s; return true; }
        return false;
    }
    bool merge(WhiteNoise& wn, const Reservoir<T>& s, float pdf) {  // reservoir.rs:41-53
        if (s.m <= 0.0f) return false;
        m += s.m - 1.0f;
        return update(wn, s.sample, s.w * s.m * pdf);
    }
    void clamp_m(float mx) { m = fmin_(m, mx); }
    void clamp_w(float mx) { w = fmin_(w, mx); }
    void norm(float pdf, float num, float den) { float d = pdf * den; w = (d == 0.0f) ? 0.0f : (w * num) / d; }
    void norm_avg(float pdf) { norm(pdf, 1.0f, m); }
    void norm_mis(float pdf) { norm(pdf, 1.0f, 1.0f); }
};

struct DiSample { float pdf, confidence; u32 light_id; V3 light_point; bool is_occluded; };
static inline DiSample di_sample_default() { DiSample s; s.pdf = 0; s.confidence = 0; s.light_id = 0; s.light_point = v3s(0); s.is_occluded = false; return s; }
typedef Reservoir<DiSample> DiReservoir;
static inline DiReservoir di_default() { DiReservoir r; r.sample = di_sample_default(); r.m = 0; r.w = 0; return r; }
static inline DiReservoir di_read(const V4* buf, size_t id) {  // reservoir/di.rs:17-35
    V4 d0 = buf[2 * id], d1 = buf[2 * id + 1];
    u32 b = f2u(d0.w);
    DiReservoir r;
    r.sample.pdf = d0.z; r.sample.confidence = (float)((b >> 8) & 0xff); r.sample.light_id = f2u(d1.w);
    r.sample.light_point = xyz(d1); r.sample.is_occluded = (b & 0xff) > 0;
    r.m = d0.x; r.w = d0.y;
    return r;
}
static inline void di_write(const DiReservoir& r, V4* buf, size_t id) {  // reservoir/di.rs:37-59
    // `confidence as u32` then `u32::from_bytes` ORs un-masked values
    buf[2 * id] = v4(r.m, r.w, r.sample.pdf, u2f(from_bytes(r.sample.is_occluded ? 1u : 0u, f2u_sat(r.sample.confidence), 0, 0)));
    buf[2 * id + 1] = v4(r.sample.light_point, u2f(r.sample.light_id));
}
static inline bool di_is_empty(const DiReservoir& r) { return r.m == 0.0f; }
static inline float di_sample_pdf_ex(const DiSample& s, const Light& light, Hit hit) {  // di.rs:108-117
    hit.gbuffer.base_color = v4(1, 1, 1, 1);
    if (!light_is_none(light) && light_contains(light, s.light_point)) return luma(light_radiance_sum(light_radiance(light, hit)));
    return 0.0f;
}
static inline float di_sample_pdf(const DiSample& s, const Scene& sc, const Hit& hit) { return di_sample_pdf_ex(s, sc.lights[s.light_id], hit); }
static inline float di_sample_pdf_prev(const DiSample& s, const Scene& sc, const Hit& hit) { return di_sample_pdf_ex(s, light_rollback(sc.lights[s.light_id]), hit); }
static inline Ray di_sample_ray(const DiSample& s, V3 hit_point) {  // di.rs:119-123
    V3 dir = hit_point - s.light_point;
    return ray_with_len(ray_new(s.light_point, normalize(dir)), length(dir));
}

struct GiSample { float pdf; u32 rng; V3 radiance, v1_point, v2_point, v2_normal; };
struct GiReservoir : Reservoir<GiSample> { float confidence; };
static inline GiReservoir gi_default() {
    GiReservoir r; r.sample.pdf = 0; r.sample.rng = 0; r.sample.radiance = v3s(0); r.sample.v1_point = v3s(0); r.sample.v2_point = v3s(0); r.sample.v2_normal = v3s(0);
    r.m = 0; r.w = 0; r.confidence = 0; return r;
}
static inline GiReservoir gi_read(const V4* buf, size_t id) {  // reservoir/gi.rs:19-40
    V4 d0 = buf[4 * id], d1 = buf[4 * id + 1], d2 = buf[4 * id + 2], d3 = buf[4 * id + 3];
    GiReservoir r;
    r.sample.pdf = d2.w; r.sample.rng = f2u(d3.w); r.sample.radiance = xyz(d0); r.sample.v1_point = xyz(d1); r.sample.v2_point = xyz(d2);
    r.sample.v2_normal = normal_decode(v2(d3.x, d3.y));
    r.m = d0.w; r.w = d1.w; r.confidence = d3.z;
    return r;
}
static inline void gi_write(const GiReservoir& r, V4* buf, size_t id) {  // reservoir/gi.rs:42-57
    V2 n = normal_encode(r.sample.v2_normal);
    buf[4 * id] = v4(r.sample.radiance, r.m);
    buf[4 * id + 1] = v4(r.sample.v1_point, r.w);
    buf[4 * id + 2] = v4(r.sample.v2_point, r.sample.pdf);
    buf[4 * id + 3] = v4(n.x, n.y, r.confidence, u2f(r.sample.rng));
}
static inline bool gi_is_empty(const GiReservoir& r) { return r.m == 0.0f; }
static inline bool gi_sample_exists(const GiSample& s) { return s.v2_point != v3s(0); }
static inline V3 gi_sample_dir(const GiSample& s, V3 p) { return normalize(s.v2_point - p); }
static inline float gi_sample_cosine(const GiSample& s, const Hit& hit) { return fmax_(dot(gi_sample_dir(s, hit.point), hit.gbuffer.normal), 0.0f); }
static inline V3 gi_sample_spec_brdf(const GiSample& s, const Hit& hit) { return specular_brdf_eval(hit.gbuffer, gi_sample_dir(s, hit.point), -hit.dir); }
static inline float gi_sample_pdf(const GiSample& s, Hit hit) {  // gi.rs:98-112
    if (!gi_sample_exists(s)) return 0.0f;
    hit.gbuffer.base_color = v4(1, 1, 1, 1);
    float diff = luma(diffuse_brdf_eval(hit.gbuffer));
    float spec = luma(gi_sample_spec_brdf(s, hit));
    return luma(s.radiance) * gi_sample_cosine(s, hit) * (diff + spec);
}
static inline Ray gi_sample_ray(const GiSample& s, V3 hit_point) {  // gi.rs:114-117
    return ray_with_len(ray_new(hit_point, gi_sample_dir(s, hit_point)), distance(s.v2_point, hit_point) - 0.01f);
}
static inline void gi_partial_jacobian(const GiSample& s, V3 hit_point, float* dist, float* cosv) {
    V3 vec = hit_point - s.v2_point;
    *dist = length(vec);
    *cosv = saturate(dot(s.v2_normal, vec / *dist));
}
static inline float gi_sample_jacobian(const GiSample& s, V3 new_hit_point) {  // gi.rs:135-151
    if (!gi_sample_exists(s)) return 1.0f;
    float new_dist, new_cos, old_dist, old_cos;
    gi_partial_jacobian(s, new_hit_point, &new_dist, &new_cos);
    gi_partial_jacobian(s, s.v1_point, &old_dist, &old_cos);
    float x = new_cos * old_dist * old_dist;
    float y = old_cos * new_dist * new_dist;
    return (y == 0.0f) ? 0.0f : x / y;
}

// EphemeralReservoir::build (reservoir/ephemeral.rs:14-55)
struct EphemeralSample { u32 light_id; LightRadiance light_rad; };
typedef Reservoir<EphemeralSample> EphemeralReservoir;
static inline EphemeralReservoir ephemeral_build(WhiteNoise& wn, const Scene& sc, const Hit& hit) {
    EphemeralReservoir res; res.sample.light_id = 0; res.sample.light_rad = light_radiance_default(); res.m = 0; res.w = 0;
    float res_pdf = 0.0f;
    u32 lc = sc.world.light_count;
    u32 max_samples = lc < 16 ? lc : 16;
    float sample_ipdf = (float)lc;
    for (u32 nth = 0; nth < max_samples; nth++) {
        EphemeralSample s;
        s.light_id = wnoise_sample_int(wn) % lc;
        s.light_rad = light_radiance(sc.lights[s.light_id], hit);
        float sample_pdf = perc_luma(s.light_rad.radiance);
        if (res.update(wn, s, sample_pdf * sample_ipdf)) res_pdf = sample_pdf;
    }
    res.norm_avg(res_pdf);
    return res;
}

// Mis (reservoir/mis.rs:12-155)
struct Mis { float lhs_m, rhs_m, rhs_jacobian, lhs_lhs_pdf, lhs_rhs_pdf, rhs_lhs_pdf, rhs_rhs_pdf; };
struct MisResult { float m, lhs_pdf, lhs_mis, rhs_pdf, rhs_mis; };
static inline float mis_mis(float x, float y) { float s = x + y; return (s == 0.0f) ? 0.0f : x / s; }
static inline float mis_m(float q0, float q1) { return (q0 <= 0.0f) ? 1.0f : saturate(pow_(fmin_(q1 / q0, 1.0f), 8.0f)); }
static inline MisResult mis_eval(const Mis& s) {
    MisResult r;
    r.m = s.rhs_m * fmin_(mis_m(s.rhs_rhs_pdf, s.rhs_lhs_pdf), mis_m(s.lhs_rhs_pdf, s.lhs_lhs_pdf));
    float t = mis_mis(s.lhs_m, s.rhs_m);
    r.lhs_mis = t + (1.0f - t) * mis_mis(s.lhs_m * s.lhs_lhs_pdf, s.rhs_m * s.lhs_rhs_pdf);
    r.rhs_mis = (1.0f - t) * mis_mis(s.rhs_m * s.rhs_rhs_pdf * s.rhs_jacobian, s.lhs_m * s.rhs_lhs_pdf);
    r.lhs_pdf = s.lhs_lhs_pdf; r.rhs_pdf = s.rhs_lhs_pdf;
    return r;
}

// checkerboard helpers (strolle-gpu/src/utils.rs:33-43)
static inline UV2 resolve_checkerboard(UV2 g, u32 frame) { return uv2(g.x * 2u + ((frame + g.y) % 2u), g.y); }
static inline UV2 resolve_checkerboard_alt(UV2 g, u32 frame) { return resolve_checkerboard(g, frame + 1u); }
static inline bool got_checkerboard_at(UV2 p, u32 frame) { UV2 r = resolve_checkerboard(uv2(p.x / 2u, p.y), frame); return r.x == p.x && r.y == p.y; }
static inline bool frame_is_gi_tracing(u32 frame) { return frame % 6u < 4u; }  // frame.rs:19-21

}  // namespace orc
