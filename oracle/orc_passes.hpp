// ORACLE — TEST INFRASTRUCTURE ONLY (see orc_math.hpp header).
//
// Restatement of the `strolle-shaders` compute entry points (L2 kernels) as
// plain loops over the dispatch grid, plus the per-camera buffers of
// strolle/src/camera_controller/buffers.rs.  One function per reference
// entry point; "return" in a kernel body == `continue` here (no write).
#pragma once
#include <string>
#include <vector>
#include "orc_gpu.hpp"

namespace orc {

typedef std::vector<V4> Buf;

// strolle/src/camera_controller/buffers.rs:53-339 (textures are row-major vec4 arrays)
struct CamState {
    int w, h;
    Camera curr_camera, prev_camera;
    Buf prim_gbuffer_d0[2], prim_gbuffer_d1[2], prim_surface_map[2];
    Buf reprojection_map, velocity_map;
    Buf di_reservoirs[3];
    Buf di_diff_samples, di_diff_prev_colors, di_diff_curr_colors, di_diff_moments[2], di_diff_stash, di_spec_samples;
    Buf gi_d0, gi_d1, gi_d2, gi_reservoirs[4];
    Buf gi_diff_samples, gi_diff_prev_colors, gi_diff_curr_colors, gi_diff_moments[2], gi_diff_stash, gi_spec_samples;
    Buf ref_hits, ref_rays, ref_colors;
    Buf prim_triangle_ids;   // not in the reference: x = bits of the primary-hit triangle id (parity hook)
    Buf output;              // composed frame (frame_composition target), linear HDR rgba
    void init(int w_, int h_) {
        w = w_; h = h_;
        size_t n = (size_t)w * h;
        Buf* one[] = {&prim_gbuffer_d0[0], &prim_gbuffer_d0[1], &prim_gbuffer_d1[0], &prim_gbuffer_d1[1], &prim_surface_map[0], &prim_surface_map[1],
                      &reprojection_map, &velocity_map, &di_diff_samples, &di_diff_prev_colors, &di_diff_curr_colors, &di_diff_moments[0],
                      &di_diff_moments[1], &di_diff_stash, &di_spec_samples, &gi_d0, &gi_d1, &gi_d2, &gi_diff_samples, &gi_diff_prev_colors,
                      &gi_diff_curr_colors, &gi_diff_moments[0], &gi_diff_moments[1], &gi_diff_stash, &gi_spec_samples, &ref_colors,
                      &prim_triangle_ids, &output};
        for (Buf* b : one) b->assign(n, v4z());
        for (int i = 0; i < 3; i++) di_reservoirs[i].assign(2 * n, v4z());
        for (int i = 0; i < 4; i++) gi_reservoirs[i].assign(4 * n, v4z());
        ref_hits.assign(2 * n, v4z());
        ref_rays.assign(3 * n, v4z());
    }
    Buf* by_name(const std::string& s) {
        struct E { const char* n; Buf* b; };
        E tab[] = {{"prim_gbuffer_d0_a", &prim_gbuffer_d0[0]}, {"prim_gbuffer_d0_b", &prim_gbuffer_d0[1]},
                   {"prim_gbuffer_d1_a", &prim_gbuffer_d1[0]}, {"prim_gbuffer_d1_b", &prim_gbuffer_d1[1]},
                   {"prim_surface_map_a", &prim_surface_map[0]}, {"prim_surface_map_b", &prim_surface_map[1]},
                   {"reprojection_map", &reprojection_map}, {"velocity_map", &velocity_map},
                   {"di_reservoirs_0", &di_reservoirs[0]}, {"di_reservoirs_1", &di_reservoirs[1]}, {"di_reservoirs_2", &di_reservoirs[2]},
                   {"di_diff_samples", &di_diff_samples}, {"di_diff_prev_colors", &di_diff_prev_colors}, {"di_diff_curr_colors", &di_diff_curr_colors},
                   {"di_diff_moments_a", &di_diff_moments[0]}, {"di_diff_moments_b", &di_diff_moments[1]}, {"di_diff_stash", &di_diff_stash},
                   {"di_spec_samples", &di_spec_samples}, {"gi_d0", &gi_d0}, {"gi_d1", &gi_d1}, {"gi_d2", &gi_d2},
                   {"gi_reservoirs_0", &gi_reservoirs[0]}, {"gi_reservoirs_1", &gi_reservoirs[1]}, {"gi_reservoirs_2", &gi_reservoirs[2]},
                   {"gi_reservoirs_3", &gi_reservoirs[3]}, {"gi_diff_samples", &gi_diff_samples}, {"gi_diff_prev_colors", &gi_diff_prev_colors},
                   {"gi_diff_curr_colors", &gi_diff_curr_colors}, {"gi_diff_moments_a", &gi_diff_moments[0]}, {"gi_diff_moments_b", &gi_diff_moments[1]},
                   {"gi_diff_stash", &gi_diff_stash}, {"gi_spec_samples", &gi_spec_samples}, {"ref_hits", &ref_hits}, {"ref_rays", &ref_rays},
                   {"ref_colors", &ref_colors}, {"prim_triangle_ids", &prim_triangle_ids}, {"output", &output}};
        for (const E& e : tab) if (s == e.n) return e.b;
        return nullptr;
    }
};

#define ORC_FOR_FULL_GRID(cs)                                   \
    _Pragma("omp parallel for schedule(dynamic, 4)")            \
    for (int gy_ = 0; gy_ < (cs).h; gy_++)                      \
        for (int gx_ = 0; gx_ < (cs).w; gx_++)

// half-width checkerboard dispatch: workgroups ((W+7)/8/2, (H+7)/8) of 8x8 threads
// (strolle/src/camera_controller/passes/di_spatial_resampling.rs:81-86)
static inline int half_grid_w(int w) { return 8 * (((w + 7) / 8) / 2); }
static inline int full_grid_h(int h) { return 8 * ((h + 7) / 8); }
#define ORC_FOR_HALF_GRID(cs)                                   \
    _Pragma("omp parallel for schedule(dynamic, 4)")            \
    for (int gy_ = 0; gy_ < full_grid_h((cs).h); gy_++)         \
        for (int gx_ = 0; gx_ < half_grid_w((cs).w); gx_++)

static inline V4& at(Buf& b, int w, UV2 p) { return b[(size_t)p.y * w + p.x]; }
static inline const V4& at(const Buf& b, int w, UV2 p) { return b[(size_t)p.y * w + p.x]; }
// storage-image reads outside the texture return 0 (tap loops on screens smaller than the
// 128 px tap radius land there after Camera::contain's single mirror, camera.rs:57-77)
static inline V4 at_or_zero(const Buf& b, int w, UV2 p) { return ((int)p.x < w && (size_t)p.y * w + p.x < b.size()) ? b[(size_t)p.y * w + p.x] : v4z(); }
static inline Hit load_hit(const Camera& cam, const Buf& d0, const Buf& d1, int w, UV2 p) {
    return hit_new(camera_ray(cam, p), gbuffer_unpack(at_or_zero(d0, w, p), at_or_zero(d1, w, p)));
}


// storage-image semantics: out-of-bounds writes are dropped, reads give 0
static inline void tex_wr(Buf& b, const CamState& cs, UV2 p, V4 v) { if ((int)p.x < cs.w && (int)p.y < cs.h) b[(size_t)p.y * cs.w + p.x] = v; }
static inline V4 tex_rd(const Buf& b, const CamState& cs, UV2 p) { return ((int)p.x < cs.w && (int)p.y < cs.h) ? b[(size_t)p.y * cs.w + p.x] : v4z(); }

// ---------------------------------------------------------------------------
// Primary-visibility G-buffer (replaces the rasteriser).
// Follows strolle-shaders/src/prim_raster.rs:41-128 with the pixel's own
// primary ray standing in for the rasterised fragment (quirk C-9 documented).
// ---------------------------------------------------------------------------
static void pass_prim_gbuffer(CamState& cs, const Scene& sc, bool alternate) {
    int cur = alternate ? 1 : 0;
    const Camera& cam = cs.curr_camera;
    ORC_FOR_FULL_GRID(cs) {
        UV2 p = uv2(gx_, gy_);
        Ray ray = camera_ray(cam, p);
        TriangleHit th = ray_trace(ray, sc);
        V4 g0 = v4z(), g1 = v4z(), surf = v4z(), vel = v4z(), tid = v4(u2f(0xffffffffu), 0, 0, 0);
        if (trihit_is_some(th)) {
            const Material& m = sc.materials[th.material_id];
            GBufferEntry g;
            V2 mr = material_metallic_roughness(sc, m, th.uv);
            g.base_color = material_base_color(sc, m, th.uv);
            g.normal = th.normal;
            g.metallic = mr.x;
            g.emissive = material_emissive(sc, m, th.uv);
            g.roughness = mr.y;
            g.reflectance = m.reflectance;
            g.depth = distance(ray.origin, th.point);
            gbuffer_pack(g, &g0, &g1);
            V2 n = normal_encode(th.normal);
            surf = v4(n.x, n.y, g.depth, m.roughness);
            // prim_raster::vs (prim_raster.rs:25-34): prev_point = prev_xform * (curr_xform_inv * point), per instance
            const V4* xf = sc.instance_xforms + 6 * (size_t)sc.tri_instance[th.triangle_id];
            V3 local = ((v3(xf[0].x, xf[0].y, xf[0].z) * th.point.x + v3(xf[1].x, xf[1].y, xf[1].z) * th.point.y) + v3(xf[2].x, xf[2].y, xf[2].z) * th.point.z) + v3(xf[0].w, xf[1].w, xf[2].w);
            V3 prev_point = ((v3(xf[3].x, xf[3].y, xf[3].z) * local.x + v3(xf[4].x, xf[4].y, xf[4].z) * local.y) + v3(xf[5].x, xf[5].y, xf[5].z) * local.z) + v3(xf[3].w, xf[4].w, xf[5].w);
            V2 velocity = camera_clip_to_screen(cam, camera_world_to_clip(cam, th.point)) -
                          camera_clip_to_screen(cs.prev_camera, camera_world_to_clip(cs.prev_camera, prev_point));
            if (length_squared(velocity) >= 0.001f) vel = v4(velocity.x, velocity.y, 0, 0);
            tid.x = u2f(th.triangle_id);
        }
        at(cs.prim_gbuffer_d0[cur], cs.w, p) = g0;
        at(cs.prim_gbuffer_d1[cur], cs.w, p) = g1;
        at(cs.prim_surface_map[cur], cs.w, p) = surf;
        at(cs.velocity_map, cs.w, p) = vel;
        at(cs.prim_triangle_ids, cs.w, p) = tid;
    }
}

// ---------------------------------------------------------------------------
// K4 frame_reprojection::main (strolle-shaders/src/frame_reprojection.rs:7-95)
// ---------------------------------------------------------------------------
static void pass_frame_reprojection(CamState& cs, bool alternate) {
    const Buf& surf_curr = cs.prim_surface_map[alternate ? 1 : 0];
    const Buf& surf_prev = cs.prim_surface_map[alternate ? 0 : 1];
    ORC_FOR_FULL_GRID(cs) {
        UV2 p = uv2(gx_, gy_);
        Reprojection rp = reprojection_default();
        Surface surface = surface_get(surf_curr.data(), cs.w, p);
        if (surface_is_sky(surface)) { at(cs.reprojection_map, cs.w, p) = reprojection_serialize(rp); continue; }
        V4 vel = at(cs.velocity_map, cs.w, p);
        V2 prev = v2((float)p.x, (float)p.y) - v2(vel.x, vel.y);
        V2 prev_r = v2(round_(prev.x), round_(prev.y));
        if (camera_contains(cs.prev_camera, prev_r)) {
            Surface ps = surface_get(surf_prev.data(), cs.w, uv2(f2u_sat(prev_r.x), f2u_sat(prev_r.y)));
            float confidence = surface_similarity(ps, surface);
            if (confidence > 0.0f) { rp.prev_x = prev.x; rp.prev_y = prev.y; rp.confidence = confidence; rp.validity = 0; }
        }
        if (reprojection_is_some(rp)) {
            IV2 c[4]; reprojection_coords(rp.prev_x, rp.prev_y, c);
            for (int i = 0; i < 4; i++) {
                if (!camera_contains(cs.curr_camera, c[i])) continue;
                if (surface_similarity(surface_get(surf_prev.data(), cs.w, uv2((u32)c[i].x, (u32)c[i].y)), surface) >= 0.25f) rp.validity |= (1u << i);
            }
        }
        at(cs.reprojection_map, cs.w, p) = reprojection_serialize(rp);
    }
}

// ---------------------------------------------------------------------------
// K5 di_sampling::main (strolle-shaders/src/di_sampling.rs:4-94)
// ---------------------------------------------------------------------------
static void pass_di_sampling(CamState& cs, const Scene& sc, bool alternate, u32 seed, u32 frame) {
    int cur = alternate ? 1 : 0;
    const Camera& cam = cs.curr_camera;
    ORC_FOR_FULL_GRID(cs) {
        UV2 p = uv2(gx_, gy_);
        size_t idx = camera_screen_to_idx(cam, p);
        WhiteNoise wn = wnoise_new(seed, p);
        Hit hit = load_hit(cam, cs.prim_gbuffer_d0[cur], cs.prim_gbuffer_d1[cur], cs.w, p);
        if (!hit_is_some(hit)) continue;
        EphemeralReservoir res = ephemeral_build(wn, sc, hit);
        DiReservoir out = di_default();
        if (res.m > 0.0f) {
            V4 bn = bnoise_texel(sc.blue_noise, p, frame);
            Ray ray = light_ray_bnoise(sc.lights[res.sample.light_id], v2(bn.x, bn.y), hit.point);
            bool is_occluded = ray_intersect(ray, sc);
            if (is_occluded) res.w = 0.0f;
            out.sample.pdf = 0.0f; out.sample.confidence = 0.0f; out.sample.light_id = res.sample.light_id;
            out.sample.light_point = ray.origin; out.sample.is_occluded = is_occluded;
            out.m = 1.0f; out.w = res.w;
        }
        di_write(out, cs.di_reservoirs[1].data(), idx);
    }
}

// ---------------------------------------------------------------------------
// K6 di_temporal_resampling::main (strolle-shaders/src/di_temporal_resampling.rs:4-112)
// ---------------------------------------------------------------------------
static void pass_di_temporal(CamState& cs, const Scene& sc, bool alternate, u32 seed) {
    int cur = alternate ? 1 : 0, prv = alternate ? 0 : 1;
    size_t npx = (size_t)cs.w * cs.h;
    ORC_FOR_FULL_GRID(cs) {
        UV2 lhs_pos = uv2(gx_, gy_);
        size_t lhs_idx = camera_screen_to_idx(cs.curr_camera, lhs_pos);
        WhiteNoise wn = wnoise_new(seed, lhs_pos);
        Hit lhs_hit = load_hit(cs.curr_camera, cs.prim_gbuffer_d0[cur], cs.prim_gbuffer_d1[cur], cs.w, lhs_pos);
        if (!hit_is_some(lhs_hit)) continue;
        DiReservoir lhs = di_read(cs.di_reservoirs[1].data(), lhs_idx);
        if (!di_is_empty(lhs)) lhs.sample.pdf = di_sample_pdf(lhs.sample, sc, lhs_hit);
        DiReservoir rhs = di_default();
        Hit rhs_hit = hit_default();
        bool rhs_killed = false;
        Reprojection rp = reprojection_deserialize(at(cs.reprojection_map, cs.w, lhs_pos));
        if (reprojection_is_some(rp)) {
            UV2 rhs_pos = reprojection_prev_pos_round(rp);
            size_t ridx = camera_screen_to_idx(cs.curr_camera, rhs_pos);
            if (ridx < npx) rhs = di_read(cs.di_reservoirs[0].data(), ridx);
            rhs.clamp_m(64.0f);
            if (!di_is_empty(rhs)) {
                const Light& rl = sc.lights[rhs.sample.light_id];
                if (light_is_slot_killed(rl)) { rhs.w = 0.0f; rhs_killed = true; }
                else if (light_is_slot_remapped(rl)) rhs.sample.light_id = f2u(rl.d3.x) - 1u;
                rhs_hit = load_hit(cs.prev_camera, cs.prim_gbuffer_d0[prv], cs.prim_gbuffer_d1[prv], cs.w, rhs_pos);
            }
        }
        DiReservoir main = di_default();
        float main_pdf = 0.0f;
        Mis mis;
        mis.lhs_m = lhs.m; mis.rhs_m = rhs.m; mis.rhs_jacobian = 1.0f;
        mis.lhs_lhs_pdf = lhs.sample.pdf; mis.rhs_rhs_pdf = rhs.sample.pdf;
        mis.lhs_rhs_pdf = ((lhs.m > 0.0f) & hit_is_some(rhs_hit)) ? di_sample_pdf_prev(lhs.sample, sc, rhs_hit) : 0.0f;
        mis.rhs_lhs_pdf = ((rhs.m > 0.0f) & !rhs_killed) ? di_sample_pdf(rhs.sample, sc, lhs_hit) : 0.0f;
        MisResult mr = mis_eval(mis);
        if (main.update(wn, lhs.sample, mr.lhs_mis * mr.lhs_pdf * lhs.w)) main_pdf = mr.lhs_pdf;
        if (main.update(wn, rhs.sample, mr.rhs_mis * mr.rhs_pdf * rhs.w)) main_pdf = mr.rhs_pdf;
        main.m = lhs.m + mr.m;
        main.sample.pdf = main_pdf;
        main.sample.confidence = rhs_killed ? 0.0f : 1.0f;
        main.norm_mis(main_pdf);
        di_write(main, cs.di_reservoirs[1].data(), lhs_idx);
    }
}

// ---------------------------------------------------------------------------
// K7 di_spatial_resampling::pick (strolle-shaders/src/di_spatial_resampling.rs:4-147)
// scratch: buf_d0 = di_diff_samples, buf_d1 = di_diff_curr_colors, buf_d2 = di_diff_stash
// (strolle/src/camera_controller/passes/di_spatial_resampling.rs:24-28)
// Deviation: the reference returns for sky pixels
// without clearing its scratch texels and later decodes stale colours as
// rays / reservoir indices (out-of-bounds).  Here sky pixels clear buf_d1
// like the GI twin does; only sky-pixel reservoirs (never shaded) differ.
// ---------------------------------------------------------------------------
static void pass_di_spatial_pick(CamState& cs, const Scene& sc, bool alternate, u32 seed, u32 frame) {
    int cur = alternate ? 1 : 0;
    const Camera& cam = cs.curr_camera;
    Buf& buf_d0 = cs.di_diff_samples; Buf& buf_d1 = cs.di_diff_curr_colors;
    ORC_FOR_HALF_GRID(cs) {
        UV2 gid = uv2(gx_, gy_);
        UV2 lhs_pos = resolve_checkerboard_alt(gid, frame / 2);
        size_t lhs_idx = camera_screen_to_idx(cam, lhs_pos);
        WhiteNoise wn = wnoise_new(seed, lhs_pos);
        UV2 pa = uv2(gid.x * 2, gid.y), pb = uv2(gid.x * 2 + 1, gid.y);
        if (!camera_contains(cam, lhs_pos)) continue;
        Hit lhs_hit = load_hit(cam, cs.prim_gbuffer_d0[cur], cs.prim_gbuffer_d1[cur], cs.w, lhs_pos);
        if (!hit_is_some(lhs_hit)) { tex_wr(buf_d1, cs, pa, v4z()); tex_wr(buf_d1, cs, pb, v4z()); continue; }
        DiReservoir lhs = di_read(cs.di_reservoirs[1].data(), lhs_idx);
        DiReservoir rhs = di_default();
        u32 rhs_nth = 0; size_t rhs_idx = 0;
        Hit rhs_hit = hit_default();
        float max_radius = 128.0f;
        while (rhs_nth < 8) {
            rhs_nth += 1;
            V2 off = wnoise_sample_disk(wn) * max_radius;
            V2 fp = v2((float)lhs_pos.x, (float)lhs_pos.y) + off;
            UV2 rhs_pos = camera_contain(cam, iv2(f2i_sat(fp.x), f2i_sat(fp.y)));
            if (rhs_pos.x == lhs_pos.x && rhs_pos.y == lhs_pos.y) continue;
            rhs_hit = load_hit(cam, cs.prim_gbuffer_d0[cur], cs.prim_gbuffer_d1[cur], cs.w, rhs_pos);
            if (!hit_is_some(rhs_hit)) { max_radius = fmax_(max_radius * 0.5f, 5.0f); continue; }
            if (abs_(rhs_hit.gbuffer.depth - lhs_hit.gbuffer.depth) > 0.33f * lhs_hit.gbuffer.depth) { max_radius = fmax_(max_radius * 0.5f, 5.0f); continue; }
            if (dot(rhs_hit.gbuffer.normal, lhs_hit.gbuffer.normal) < 0.33f) { max_radius = fmax_(max_radius * 0.5f, 5.0f); continue; }
            rhs_idx = camera_screen_to_idx(cam, rhs_pos);
            rhs = di_read(cs.di_reservoirs[1].data(), rhs_idx);
            if (!di_is_empty(rhs)) break;
        }
        if (di_is_empty(rhs)) { tex_wr(buf_d1, cs, pa, v4z()); tex_wr(buf_d1, cs, pb, v4z()); continue; }
        float lhs_rhs_pdf = di_sample_pdf(lhs.sample, sc, rhs_hit);
        float rhs_lhs_pdf = di_sample_pdf(rhs.sample, sc, lhs_hit);
        Ray ray_a = (lhs_rhs_pdf > 0.0f) ? di_sample_ray(lhs.sample, rhs_hit.point) : ray_default();
        Ray ray_b = (rhs_lhs_pdf > 0.0f) ? di_sample_ray(rhs.sample, lhs_hit.point) : ray_default();
        V2 na = normal_encode(ray_a.dir), nb = normal_encode(ray_b.dir);
        tex_wr(buf_d0, cs, pa, v4(ray_a.origin, ray_a.len));
        tex_wr(buf_d1, cs, pa, v4(na.x, na.y, u2f((u32)rhs_idx + 1u), 0.0f));
        tex_wr(buf_d0, cs, pb, v4(ray_b.origin, ray_b.len));
        tex_wr(buf_d1, cs, pb, v4(nb.x, nb.y, lhs_rhs_pdf, rhs_lhs_pdf));
    }
}

// K8 / K16 *_spatial_resampling::trace (di_spatial_resampling.rs:150-209, gi_spatial_resampling.rs:163-222)
static void pass_spatial_trace(CamState& cs, const Scene& sc, const Buf& buf_d0, const Buf& buf_d1, Buf& buf_d2) {
    ORC_FOR_FULL_GRID(cs) {
        UV2 p = uv2(gx_, gy_);
        V4 d0 = at(buf_d0, cs.w, p), d1 = at(buf_d1, cs.w, p);
        if (is_zero(d1)) { at(buf_d2, cs.w, p) = v4z(); continue; }
        Ray ray = ray_with_len(ray_new(xyz(d0), normal_decode(v2(d1.x, d1.y))), d0.w);
        bool occluded = ray_intersect(ray, sc);
        at(buf_d2, cs.w, p) = v4(occluded ? 0.0f : 1.0f, d1.z, d1.w, 0.0f);
    }
}

// K9 di_spatial_resampling::sample (di_spatial_resampling.rs:212-297)
static void pass_di_spatial_sample(CamState& cs, u32 seed, u32 frame) {
    const Camera& cam = cs.curr_camera;
    const Buf& buf_d2 = cs.di_diff_stash;
    const V4* in = cs.di_reservoirs[1].data(); V4* out = cs.di_reservoirs[2].data();
    size_t npx = (size_t)cs.w * cs.h;
    ORC_FOR_HALF_GRID(cs) {
        UV2 gid = uv2(gx_, gy_);
        UV2 lhs_pos = resolve_checkerboard_alt(gid, frame / 2);
        size_t lhs_idx = camera_screen_to_idx(cam, lhs_pos);
        WhiteNoise wn = wnoise_new(seed, lhs_pos);
        UV2 pa = uv2(gid.x * 2, gid.y), pb = uv2(gid.x * 2 + 1, gid.y);
        if (!camera_contains(cam, lhs_pos)) continue;
        V4 d0 = tex_rd(buf_d2, cs, pa), d1 = tex_rd(buf_d2, cs, pb);
        float lhs_rhs_vis = d0.x; u32 rhs_idx = f2u(d0.y);
        float rhs_lhs_vis = d1.x, lhs_rhs_pdf = d1.y, rhs_lhs_pdf = d1.z;
        DiReservoir lhs = di_read(in, lhs_idx);
        if (rhs_idx > 0 && (size_t)rhs_idx - 1 < npx) {
            DiReservoir rhs = di_read(in, (size_t)rhs_idx - 1);
            DiReservoir main = di_default();
            float main_pdf = 0.0f;
            Mis mis;
            mis.lhs_m = lhs.m; mis.rhs_m = rhs.m; mis.rhs_jacobian = 1.0f; mis.lhs_lhs_pdf = lhs.sample.pdf;
            mis.lhs_rhs_pdf = lhs_rhs_pdf * lhs_rhs_vis; mis.rhs_lhs_pdf = rhs_lhs_pdf * rhs_lhs_vis; mis.rhs_rhs_pdf = rhs.sample.pdf;
            MisResult mr = mis_eval(mis);
            if (main.update(wn, lhs.sample, mr.lhs_mis * mr.lhs_pdf * lhs.w)) main_pdf = mr.lhs_pdf;
            if (main.update(wn, rhs.sample, mr.rhs_mis * mr.rhs_pdf * rhs.w)) { main_pdf = mr.rhs_pdf; main.sample.is_occluded = lhs_rhs_vis == 0.0f; }
            main.m = lhs.m + mr.m;
            main.sample.pdf = main_pdf;
            main.norm_mis(main_pdf);
            di_write(main, out, lhs_idx);
        } else di_write(lhs, out, lhs_idx);
        UV2 other = resolve_checkerboard(gid, frame / 2);
        size_t other_idx = camera_screen_to_idx(cam, other);
        if (camera_contains(cam, other)) di_write(di_read(in, other_idx), out, other_idx);  // guard: reference copies unchecked
    }
}

// ---------------------------------------------------------------------------
// K10 di_resolving::main (strolle-shaders/src/di_resolving.rs:4-119)
// ---------------------------------------------------------------------------
static void pass_di_resolving(CamState& cs, const Scene& sc, bool alternate) {
    int cur = alternate ? 1 : 0;
    const Camera& cam = cs.curr_camera;
    V3 sun_dir = world_sun_dir(sc.world);
    ORC_FOR_FULL_GRID(cs) {
        UV2 p = uv2(gx_, gy_);
        size_t idx = camera_screen_to_idx(cam, p);
        Hit hit = load_hit(cam, cs.prim_gbuffer_d0[cur], cs.prim_gbuffer_d1[cur], cs.w, p);
        DiReservoir res = di_read(cs.di_reservoirs[2].data(), idx);
        float confidence;
        LightRadiance radiance;
        if (hit_is_some(hit)) {
            bool is_occluded = ray_intersect(di_sample_ray(res.sample, hit.point), sc);
            confidence = (res.sample.is_occluded == is_occluded) ? res.sample.confidence : 0.0f;
            res.sample.confidence = 1.0f;
            res.sample.is_occluded = is_occluded;
            if (is_occluded) radiance = light_radiance_default();
            else { radiance = light_radiance(sc.lights[res.sample.light_id], hit); radiance.radiance *= res.w; }
        } else {
            confidence = 1.0f;
            radiance.radiance = atmosphere_sample(sc, sun_dir, hit.dir);
            radiance.diff_brdf = v3s(1.0f); radiance.spec_brdf = v3s(0.0f);
        }
        float diff_brdf = (1.0f - hit.gbuffer.metallic) / PI;
        at(cs.di_diff_samples, cs.w, p) = v4(radiance.radiance * diff_brdf, confidence);
        at(cs.di_spec_samples, cs.w, p) = v4(radiance.radiance * radiance.spec_brdf, confidence);
        di_write(res, cs.di_reservoirs[0].data(), idx);
    }
}

// ---------------------------------------------------------------------------
// K11 gi_reprojection::main (strolle-shaders/src/gi_reprojection.rs:4-51)
// ---------------------------------------------------------------------------
static void pass_gi_reprojection(CamState& cs, bool alternate) {
    int cur = alternate ? 1 : 0;
    const Camera& cam = cs.curr_camera;
    size_t npx = (size_t)cs.w * cs.h;
    ORC_FOR_FULL_GRID(cs) {
        UV2 p = uv2(gx_, gy_);
        size_t idx = camera_screen_to_idx(cam, p);
        Hit hit = load_hit(cam, cs.prim_gbuffer_d0[cur], cs.prim_gbuffer_d1[cur], cs.w, p);
        if (!hit_is_some(hit)) continue;
        Reprojection rp = reprojection_deserialize(at(cs.reprojection_map, cs.w, p));
        GiReservoir res = gi_default();
        if (reprojection_is_some(rp)) {
            size_t ridx = camera_screen_to_idx(cam, reprojection_prev_pos_round(rp));
            if (ridx < npx) res = gi_read(cs.gi_reservoirs[0].data(), ridx);
        }
        res.confidence = 1.0f;
        res.sample.v1_point = hit.point;
        gi_write(res, cs.gi_reservoirs[2].data(), idx);
    }
}

// ---------------------------------------------------------------------------
// K12 gi_sampling_a::main (strolle-shaders/src/gi_sampling_a.rs:4-122)
// ---------------------------------------------------------------------------
static void pass_gi_sampling_a(CamState& cs, const Scene& sc, bool alternate, u32 seed, u32 frame) {
    int cur = alternate ? 1 : 0;
    const Camera& cam = cs.curr_camera;
    bool tracing = frame_is_gi_tracing(frame);
    ORC_FOR_HALF_GRID(cs) {
        UV2 gid = uv2(gx_, gy_);
        UV2 sp = tracing ? resolve_checkerboard(gid, frame / 2) : resolve_checkerboard(gid, frame);
        size_t idx = camera_screen_to_idx(cam, sp);
        if (!camera_contains(cam, sp)) continue;
        Ray gi_ray; float gi_ray_pdf;
        if (tracing) {
            WhiteNoise wn = wnoise_new(seed, sp);
            Hit hit = load_hit(cam, cs.prim_gbuffer_d0[cur], cs.prim_gbuffer_d1[cur], cs.w, sp);
            if (!hit_is_some(hit)) continue;
            BrdfSample s = layered_brdf_sample(hit.gbuffer, wn, -hit.dir);
            gi_ray = ray_new(hit.point, s.dir);
            gi_ray_pdf = s.pdf;
        } else {
            GiReservoir res = gi_read(cs.gi_reservoirs[2].data(), idx);
            if (gi_is_empty(res)) continue;
            gi_ray = ray_new(res.sample.v1_point, gi_sample_dir(res.sample, res.sample.v1_point));
            gi_ray_pdf = 1.0f;
        }
        TriangleHit gh = ray_trace(gi_ray, sc);
        GBufferEntry gg = gbuffer_default();
        if (trihit_is_some(gh)) {
            Material m = sc.materials[gh.material_id];
            material_regularize(m);
            gg.base_color = material_base_color(sc, m, gh.uv);
            gg.normal = gh.normal; gg.metallic = m.metallic; gg.emissive = material_emissive(sc, m, gh.uv);
            gg.roughness = m.roughness; gg.reflectance = m.reflectance;
            gg.depth = distance(gi_ray.origin, gh.point);
        }
        V4 d1, d2; gbuffer_pack(gg, &d1, &d2);
        at(cs.gi_d0, cs.w, gid) = v4(gi_ray.dir, gi_ray_pdf);
        at(cs.gi_d1, cs.w, gid) = d1;
        at(cs.gi_d2, cs.w, gid) = d2;
    }
}

// ---------------------------------------------------------------------------
// K13 gi_sampling_b::main (strolle-shaders/src/gi_sampling_b.rs:4-235)
// ---------------------------------------------------------------------------
static void pass_gi_sampling_b(CamState& cs, const Scene& sc, bool alternate, u32 seed, u32 frame) {
    int cur = alternate ? 1 : 0;
    const Camera& cam = cs.curr_camera;
    bool tracing = frame_is_gi_tracing(frame);
    V3 sun_dir = world_sun_dir(sc.world);
    ORC_FOR_HALF_GRID(cs) {
        UV2 gid = uv2(gx_, gy_);
        UV2 sp = tracing ? resolve_checkerboard(gid, frame / 2) : resolve_checkerboard(gid, frame);
        size_t idx = camera_screen_to_idx(cam, sp);
        if (!camera_contains(cam, sp)) continue;
        Hit prim_hit = load_hit(cam, cs.prim_gbuffer_d0[cur], cs.prim_gbuffer_d1[cur], cs.w, sp);
        if (!hit_is_some(prim_hit)) continue;
        V4 d0 = at(cs.gi_d0, cs.w, gid), d1 = at(cs.gi_d1, cs.w, gid), d2 = at(cs.gi_d2, cs.w, gid);
        WhiteNoise wn; Hit gi_hit; float gi_ray_pdf;
        if (tracing) {
            wn = wnoise_new(seed, sp);
            gi_hit = hit_new(ray_new(prim_hit.point, xyz(d0)), gbuffer_unpack(d1, d2));
            gi_ray_pdf = d0.w;
        } else {
            GiReservoir res = gi_read(cs.gi_reservoirs[2].data(), idx);
            if (gi_is_empty(res)) continue;
            wn.state = res.sample.rng;
            gi_hit = hit_new(ray_new(res.sample.v1_point, xyz(d0)), gbuffer_unpack(d1, d2));
            gi_ray_pdf = 1.0f;
        }
        u32 rng = wn.state;
        const u32 SKY = 0xffffffffu;
        u32 light_id; float light_pdf; V3 light_rad; V3 light_dir = v3s(0);
        if (!hit_is_some(gi_hit)) {
            light_id = SKY; light_pdf = 1.0f; light_rad = atmosphere_sample(sc, sun_dir, gi_hit.dir);
        } else {
            float atmosphere_pdf = (sc.world.sun_altitude <= -1.0f) ? 0.0f : 0.25f;
            if (sc.world.light_count == 0 || wnoise_sample(wn) < atmosphere_pdf) {
                light_id = SKY; light_pdf = atmosphere_pdf;
                light_dir = wnoise_sample_hemisphere(wn, gi_hit.gbuffer.normal);
                light_rad = atmosphere_sample(sc, sun_dir, light_dir) * dot(gi_hit.gbuffer.normal, light_dir);
            } else {
                EphemeralReservoir res = ephemeral_build(wn, sc, gi_hit);
                if (res.w > 0.0f) {
                    light_id = res.sample.light_id;
                    light_pdf = (1.0f / res.w) * (1.0f - atmosphere_pdf);
                    light_rad = res.sample.light_rad.radiance * (v3s(1.0f) + res.sample.light_rad.spec_brdf);
                } else { light_id = 0; light_pdf = 1.0f; light_rad = v3s(0); }
            }
        }
        V3 radiance;
        if (light_pdf > 0.0f) {
            float light_vis;
            if (hit_is_some(gi_hit)) {
                Ray ray = (light_id == SKY) ? ray_new(gi_hit.point, light_dir) : light_ray_wnoise(sc.lights[light_id], wn, gi_hit.point);
                light_vis = ray_intersect(ray, sc) ? 0.0f : 1.0f;
            } else light_vis = 1.0f;
            radiance = light_rad * light_vis / light_pdf;
        } else radiance = v3s(0);
        if (hit_is_some(gi_hit)) {
            radiance *= xyz(gi_hit.gbuffer.base_color) / PI;
            radiance += gi_hit.gbuffer.emissive;
        }
        GiReservoir res = gi_default();
        if (gi_ray_pdf > 0.0f) {
            V3 v1 = prim_hit.point, v2p, v2n;
            if (hit_is_some(gi_hit)) { v2p = gi_hit.point; v2n = gi_hit.gbuffer.normal; }
            else { v2p = v1 + gi_hit.dir * 1000.0f; v2n = -gi_hit.dir; }
            res.sample.pdf = 0.0f; res.sample.rng = rng; res.sample.radiance = radiance;
            res.sample.v1_point = v1; res.sample.v2_point = v2p; res.sample.v2_normal = v2n;
            res.m = 1.0f; res.w = 1.0f / gi_ray_pdf;
            res.sample.pdf = gi_sample_pdf(res.sample, prim_hit);
        }
        gi_write(res, cs.gi_reservoirs[1].data(), idx);
    }
}

// ---------------------------------------------------------------------------
// K14 gi_temporal_resampling::main (strolle-shaders/src/gi_temporal_resampling.rs:4-156)
// ---------------------------------------------------------------------------
static void pass_gi_temporal(CamState& cs, bool alternate, u32 seed, u32 frame) {
    int cur = alternate ? 1 : 0, prv = alternate ? 0 : 1;
    bool tracing = frame_is_gi_tracing(frame);
    ORC_FOR_FULL_GRID(cs) {
        UV2 lhs_pos = uv2(gx_, gy_);
        size_t lhs_idx = camera_screen_to_idx(cs.curr_camera, lhs_pos);
        WhiteNoise wn = wnoise_new(seed, lhs_pos);
        Hit lhs_hit = load_hit(cs.curr_camera, cs.prim_gbuffer_d0[cur], cs.prim_gbuffer_d1[cur], cs.w, lhs_pos);
        V4* curr = cs.gi_reservoirs[1].data();
        if (!hit_is_some(lhs_hit)) { gi_write(gi_default(), curr, lhs_idx); continue; }
        bool got_sample = tracing ? (frame % 2 == 0 && got_checkerboard_at(lhs_pos, frame / 2)) : got_checkerboard_at(lhs_pos, frame);
        GiReservoir lhs = got_sample ? gi_read(curr, lhs_idx) : gi_default();
        GiReservoir rhs = gi_default();
        Hit rhs_hit = hit_default();
        Reprojection rp = reprojection_deserialize(at(cs.reprojection_map, cs.w, lhs_pos));
        if (reprojection_is_some(rp)) {
            rhs = gi_read(cs.gi_reservoirs[2].data(), lhs_idx);
            rhs.confidence = 1.0f;
            rhs.clamp_m(128.0f);
            if (!tracing && !gi_is_empty(lhs) && !gi_is_empty(rhs) && gi_sample_exists(rhs.sample)) {
                if (distance(lhs.sample.radiance, rhs.sample.radiance) > 0.33f) rhs.confidence = 0.0f;
                rhs.sample.radiance = lhs.sample.radiance;
                rhs.sample.v2_point = lhs.sample.v2_point;
                rhs.sample.v2_normal = lhs.sample.v2_normal;
            }
            if (!gi_is_empty(rhs)) {
                UV2 rhs_pos = reprojection_prev_pos_round(rp);
                rhs_hit = load_hit(cs.prev_camera, cs.prim_gbuffer_d0[prv], cs.prim_gbuffer_d1[prv], cs.w, rhs_pos);
            }
        }
        GiReservoir main = gi_default();
        float main_pdf = 0.0f;
        if (tracing) {
            Mis mis;
            mis.lhs_m = lhs.m; mis.rhs_m = rhs.m; mis.rhs_jacobian = 1.0f; mis.lhs_lhs_pdf = lhs.sample.pdf; mis.rhs_rhs_pdf = rhs.sample.pdf;
            mis.lhs_rhs_pdf = ((lhs.m > 0.0f) & hit_is_some(rhs_hit)) ? gi_sample_pdf(lhs.sample, rhs_hit) : 0.0f;
            mis.rhs_lhs_pdf = (rhs.m > 0.0f) ? gi_sample_pdf(rhs.sample, lhs_hit) : 0.0f;
            MisResult mr = mis_eval(mis);
            if (main.update(wn, lhs.sample, mr.lhs_mis * mr.lhs_pdf * lhs.w)) main_pdf = mr.lhs_pdf;
            if (main.update(wn, rhs.sample, mr.rhs_mis * mr.rhs_pdf * rhs.w)) main_pdf = mr.rhs_pdf;
            main.m = lhs.m + mr.m;
            main.confidence = 1.0f;
            main.norm_mis(main_pdf);
        } else {
            if (main.merge(wn, rhs, rhs.sample.pdf)) main_pdf = rhs.sample.pdf;
            main.confidence = rhs.confidence;
            main.norm_avg(main_pdf);
        }
        main.sample.pdf = main_pdf;
        main.sample.v1_point = lhs_hit.point;
        main.clamp_w(5.0f);
        gi_write(main, curr, lhs_idx);
    }
}

// ---------------------------------------------------------------------------
// K15 gi_spatial_resampling::pick (strolle-shaders/src/gi_spatial_resampling.rs:4-160)
// ---------------------------------------------------------------------------
static void pass_gi_spatial_pick(CamState& cs, bool alternate, u32 seed, u32 frame) {
    int cur = alternate ? 1 : 0;
    const Camera& cam = cs.curr_camera;
    Buf& buf_d0 = cs.gi_d0; Buf& buf_d1 = cs.gi_d1;
    const V4* reservoirs = cs.gi_reservoirs[1].data();
    ORC_FOR_HALF_GRID(cs) {
        UV2 gid = uv2(gx_, gy_);
        UV2 lhs_pos = resolve_checkerboard_alt(gid, frame / 2);
        size_t lhs_idx = camera_screen_to_idx(cam, lhs_pos);
        WhiteNoise wn = wnoise_new(seed, lhs_pos);
        UV2 pa = uv2(gid.x * 2, gid.y), pb = uv2(gid.x * 2 + 1, gid.y);
        if (!camera_contains(cam, lhs_pos)) continue;
        Hit lhs_hit = load_hit(cam, cs.prim_gbuffer_d0[cur], cs.prim_gbuffer_d1[cur], cs.w, lhs_pos);
        GiReservoir lhs = gi_read(reservoirs, lhs_idx);
        if (!hit_is_some(lhs_hit) || gi_is_empty(lhs)) { tex_wr(buf_d1, cs, pa, v4z()); tex_wr(buf_d1, cs, pb, v4z()); continue; }
        GiReservoir rhs = gi_default();
        u32 rhs_nth = 0; size_t rhs_idx = 0;
        Hit rhs_hit = hit_default();
        float rhs_jacobian = 0.0f;
        float max_radius = 128.0f;
        while (rhs_nth < 8) {
            rhs_nth += 1;
            V2 off = wnoise_sample_disk(wn) * max_radius;
            V2 fp = v2((float)lhs_pos.x, (float)lhs_pos.y) + off;
            UV2 rhs_pos = camera_contain(cam, iv2(f2i_sat(fp.x), f2i_sat(fp.y)));
            if (rhs_pos.x == lhs_pos.x && rhs_pos.y == lhs_pos.y) continue;
            rhs_hit = load_hit(cam, cs.prim_gbuffer_d0[cur], cs.prim_gbuffer_d1[cur], cs.w, rhs_pos);
            if (!hit_is_some(rhs_hit)) { max_radius = fmax_(max_radius * 0.5f, 5.0f); continue; }
            if (abs_(rhs_hit.gbuffer.depth - lhs_hit.gbuffer.depth) > 0.33f * lhs_hit.gbuffer.depth) { max_radius = fmax_(max_radius * 0.5f, 5.0f); continue; }
            if (dot(rhs_hit.gbuffer.normal, lhs_hit.gbuffer.normal) < 0.33f) { max_radius = fmax_(max_radius * 0.5f, 5.0f); continue; }
            rhs_idx = camera_screen_to_idx(cam, rhs_pos);
            rhs = gi_read(reservoirs, rhs_idx);
            if (gi_is_empty(rhs)) continue;
            rhs_jacobian = gi_sample_jacobian(rhs.sample, lhs_hit.point);
            if (rhs_jacobian < 1.0f / 10.0f || rhs_jacobian > 10.0f) { rhs.m = 0.0f; continue; }
            rhs_jacobian = clampf(rhs_jacobian, 1.0f / 3.0f, 3.0f);
            break;
        }
        if (gi_is_empty(rhs) || !hit_is_some(rhs_hit)) { tex_wr(buf_d1, cs, pa, v4z()); tex_wr(buf_d1, cs, pb, v4z()); continue; }
        float lhs_rhs_pdf = gi_sample_pdf(lhs.sample, rhs_hit);
        float rhs_lhs_pdf = gi_sample_pdf(rhs.sample, lhs_hit);
        Ray ray_a = (lhs_rhs_pdf > 0.0f) ? gi_sample_ray(lhs.sample, rhs_hit.point) : ray_default();
        Ray ray_b = (rhs_lhs_pdf > 0.0f) ? gi_sample_ray(rhs.sample, lhs_hit.point) : ray_default();
        V2 na = normal_encode(ray_a.dir), nb = normal_encode(ray_b.dir);
        tex_wr(buf_d0, cs, pa, v4(ray_a.origin, ray_a.len));
        tex_wr(buf_d1, cs, pa, v4(na.x, na.y, u2f((u32)rhs_idx + 1u), rhs_jacobian));
        tex_wr(buf_d0, cs, pb, v4(ray_b.origin, ray_b.len));
        tex_wr(buf_d1, cs, pb, v4(nb.x, nb.y, lhs_rhs_pdf, rhs_lhs_pdf));
    }
}

// K17 gi_spatial_resampling::sample (gi_spatial_resampling.rs:225-314)
static void pass_gi_spatial_sample(CamState& cs, u32 seed, u32 frame) {
    const Camera& cam = cs.curr_camera;
    const Buf& buf_d2 = cs.gi_d2;
    const V4* in = cs.gi_reservoirs[1].data(); V4* out = cs.gi_reservoirs[2].data();
    size_t npx = (size_t)cs.w * cs.h;
    ORC_FOR_HALF_GRID(cs) {
        UV2 gid = uv2(gx_, gy_);
        UV2 sp = resolve_checkerboard_alt(gid, frame / 2);
        size_t idx = camera_screen_to_idx(cam, sp);
        WhiteNoise wn = wnoise_new(seed, sp);
        UV2 pa = uv2(gid.x * 2, gid.y), pb = uv2(gid.x * 2 + 1, gid.y);
        if (!camera_contains(cam, sp)) continue;
        V4 d0 = tex_rd(buf_d2, cs, pa), d1 = tex_rd(buf_d2, cs, pb);
        float lhs_rhs_vis = d0.x; u32 rhs_idx = f2u(d0.y); float rhs_jacobian = d0.z;
        float rhs_lhs_vis = d1.x, lhs_rhs_pdf = d1.y, rhs_lhs_pdf = d1.z;
        GiReservoir lhs = gi_read(in, idx);
        if (rhs_idx > 0 && (size_t)rhs_idx - 1 < npx) {
            GiReservoir rhs = gi_read(in, (size_t)rhs_idx - 1);
            GiReservoir main = gi_default();
            float main_pdf = 0.0f;
            Mis mis;
            mis.lhs_m = lhs.m; mis.rhs_m = rhs.m; mis.rhs_jacobian = rhs_jacobian; mis.lhs_lhs_pdf = lhs.sample.pdf;
            mis.lhs_rhs_pdf = lhs_rhs_pdf * lhs_rhs_vis; mis.rhs_lhs_pdf = rhs_lhs_pdf * rhs_lhs_vis; mis.rhs_rhs_pdf = rhs.sample.pdf;
            MisResult mr = mis_eval(mis);
            if (main.update(wn, lhs.sample, mr.lhs_mis * mr.lhs_pdf * lhs.w)) main_pdf = mr.lhs_pdf;
            if (main.update(wn, rhs.sample, mr.rhs_mis * mr.rhs_pdf * rhs.w * rhs_jacobian)) main_pdf = mr.rhs_pdf;
            main.m = lhs.m + mr.m;
            main.confidence = 1.0f;
            main.sample.pdf = main_pdf;
            main.sample.v1_point = lhs.sample.v1_point;
            main.norm_mis(main_pdf);
            main.clamp_w(5.0f);
            gi_write(main, out, idx);
        } else gi_write(lhs, out, idx);
        UV2 other = resolve_checkerboard(gid, frame / 2);
        size_t other_idx = camera_screen_to_idx(cam, other);
        if (camera_contains(cam, other)) gi_write(gi_read(in, other_idx), out, other_idx);
    }
}

// ---------------------------------------------------------------------------
// K18 gi_preview_resampling::main (strolle-shaders/src/gi_preview_resampling.rs:4-138)
// ---------------------------------------------------------------------------
static void pass_gi_preview(CamState& cs, bool alternate, u32 seed, u32 source, u32 nth, const Buf& in_a, const Buf& in_b, Buf& outb) {
    int cur = alternate ? 1 : 0;
    const Camera& cam = cs.curr_camera;
    const V4* in = (source == 0) ? in_a.data() : in_b.data();
    V4* out = outb.data();
    const Buf& surf = cs.prim_surface_map[cur];
    ORC_FOR_FULL_GRID(cs) {
        UV2 cp = uv2(gx_, gy_);
        size_t cidx = camera_screen_to_idx(cam, cp);
        WhiteNoise wn = wnoise_new(seed, cp);
        Hit chit = load_hit(cam, cs.prim_gbuffer_d0[cur], cs.prim_gbuffer_d1[cur], cs.w, cp);
        if (!hit_is_some(chit)) { gi_write(gi_default(), out, cidx); continue; }
        GiReservoir main = gi_default();
        float main_pdf = 0.0f;
        GiReservoir center = gi_read(in, cidx);
        if (main.merge(wn, center, center.sample.pdf)) main_pdf = center.sample.pdf;
        u32 max_samples = f2u_sat(lerp_c(8.0f, 0.0f, main.m / 8.0f));
        float max_radius = (nth == 0) ? 128.0f : 64.0f;
        u32 sample_nth = 0;
        bool bail = false;
        while (sample_nth < max_samples) {
            sample_nth += 1;
            V2 off = wnoise_sample_disk(wn) * max_radius;
            V2 fp = v2((float)cp.x, (float)cp.y) + off;
            UV2 sp = camera_contain(cam, iv2(f2i_sat(fp.x), f2i_sat(fp.y)));
            if (sp.x == cp.x && sp.y == cp.y) { bail = true; break; }   // quirk C-6: `return` without writing
            if (!camera_contains(cam, sp)) continue;   // out-of-texture read == sky
            Surface ss = surface_get(surf.data(), cs.w, sp);
            if (surface_is_sky(ss)) continue;
            if (abs_(ss.depth - chit.gbuffer.depth) > 0.25f * chit.gbuffer.depth) continue;
            if (dot(ss.normal, chit.gbuffer.normal) < 0.5f) continue;
            GiReservoir sample = gi_read(in, camera_screen_to_idx(cam, sp));
            if (gi_is_empty(sample)) continue;
            float sample_pdf = gi_sample_pdf(sample.sample, chit);
            float sample_jacobian = gi_sample_jacobian(sample.sample, chit.point);
            if (sample_jacobian < 1.0f / 10.0f || sample_jacobian > 10.0f) continue;
            sample_jacobian = clampf(sample_jacobian, 1.0f / 3.0f, 3.0f);
            if (main.merge(wn, sample, sample_pdf * sample_jacobian)) main_pdf = sample_pdf;
        }
        if (bail) continue;
        main.confidence = center.confidence;
        main.sample.pdf = main_pdf;
        main.sample.v1_point = center.sample.v1_point;
        main.norm_avg(main_pdf);
        main.clamp_w(5.0f);
        gi_write(main, out, cidx);
    }
}

// ---------------------------------------------------------------------------
// K19 gi_resolving::main (strolle-shaders/src/gi_resolving.rs:4-67)
// ---------------------------------------------------------------------------
static void pass_gi_resolving(CamState& cs, bool alternate, u32 source) {
    int cur = alternate ? 1 : 0;
    const Camera& cam = cs.curr_camera;
    const V4* in = (source == 0) ? cs.gi_reservoirs[1].data() : cs.gi_reservoirs[2].data();
    V4* out = cs.gi_reservoirs[0].data();
    ORC_FOR_FULL_GRID(cs) {
        UV2 p = uv2(gx_, gy_);
        size_t idx = camera_screen_to_idx(cam, p);
        Hit hit = load_hit(cam, cs.prim_gbuffer_d0[cur], cs.prim_gbuffer_d1[cur], cs.w, p);
        GiReservoir res = gi_read(out, idx);
        float confidence; V3 radiance;
        if (hit_is_some(hit)) { confidence = res.confidence; radiance = res.w * gi_sample_cosine(res.sample, hit) * res.sample.radiance; }
        else { confidence = 1.0f; radiance = v3s(0); }
        float diff_brdf = (1.0f - hit.gbuffer.metallic) / PI;
        V3 spec_brdf = gi_sample_spec_brdf(res.sample, hit);
        at(cs.gi_diff_samples, cs.w, p) = v4(radiance * diff_brdf, confidence);
        at(cs.gi_spec_samples, cs.w, p) = v4(radiance * spec_brdf, confidence);
        gi_write(gi_read(in, idx), out, idx);
    }
}

// ---------------------------------------------------------------------------
// K20 frame_denoising::reproject (strolle-shaders/src/frame_denoising.rs:4-78)
// ---------------------------------------------------------------------------
static void pass_denoise_reproject(CamState& cs, bool alternate, const Buf& prev_colors, const Buf& prev_moments, const Buf& samples, Buf& colors, Buf& moments) {
    const Buf& surf = cs.prim_surface_map[alternate ? 1 : 0];
    ORC_FOR_FULL_GRID(cs) {
        UV2 p = uv2(gx_, gy_);
        if (surface_is_sky(surface_get(surf.data(), cs.w, p))) { at(colors, cs.w, p) = at(samples, cs.w, p); continue; }
        V4 sample = at(samples, cs.w, p);
        float sample_luma = luma(xyz(sample));
        Reprojection rp = reprojection_deserialize(at(cs.reprojection_map, cs.w, p));
        V3 color, moment;
        if (reprojection_is_some(rp) && sample.w > 0.0f) {
            V4 pc = bilinear_reproject(rp, prev_colors.data(), cs.w, cs.h);
            V4 pm = bilinear_reproject(rp, prev_moments.data(), cs.w, cs.h);
            V3 prev_color = xyz(pc);
            float prev_history = pm.x, prev_m1 = pm.y, prev_m2 = pm.z;
            V3 curr_color = xyz(sample);
            float curr_history = fmin_(prev_history + 1.0f, 16.0f);
            float curr_m1 = sample_luma, curr_m2 = sample_luma * sample_luma;
            float alpha = 1.0f / curr_history;
            color = lerp_c(prev_color, curr_color, alpha);
            moment = v3(curr_history, lerp_c(prev_m1, curr_m1, alpha), lerp_c(prev_m2, curr_m2, alpha));
        } else {
            color = xyz(sample);
            moment = v3(1.0f, sample_luma, sample_luma * sample_luma);
        }
        at(colors, cs.w, p) = v4(color, 0.0f);
        at(moments, cs.w, p) = v4(moment, 0.0f);
    }
}

// frame_denoising::sample_weight (frame_denoising.rs:363-392)
static inline float svgf_sample_weight(float center_luma, const Surface& cs_, float sample_luma, const Surface& ss, float luma_sigma, float depth_sigma) {
    float luma_weight = abs_(sqrt_(center_luma) - sqrt_(sample_luma)) * luma_sigma;
    float leeway = cs_.depth * depth_sigma;
    float diff = abs_(ss.depth - cs_.depth);
    float depth_weight = (diff >= leeway) ? 0.0f : 1.0f - diff / leeway;
    float normal_weight = pow_(fmax_(dot(ss.normal, cs_.normal), 0.0f), 64.0f);
    return exp_(-luma_weight) * depth_weight * normal_weight;
}

// K21 frame_denoising::estimate_variance (frame_denoising.rs:81-217)
static void pass_denoise_estimate_variance(CamState& cs, bool alternate) {
    int cur = alternate ? 1 : 0;
    const Buf& surf = cs.prim_surface_map[cur];
    const Buf& di_colors = cs.di_diff_curr_colors; const Buf& di_moments = cs.di_diff_moments[cur]; Buf& di_out = cs.di_diff_stash;
    const Buf& gi_colors = cs.gi_diff_curr_colors; const Buf& gi_moments = cs.gi_diff_moments[cur]; Buf& gi_out = cs.gi_diff_stash;
    const Camera& cam = cs.curr_camera;
    ORC_FOR_FULL_GRID(cs) {
        UV2 p = uv2(gx_, gy_);
        Surface csf = surface_get(surf.data(), cs.w, p);
        V4 cdi = at(di_colors, cs.w, p); float cdi_luma = luma(xyz(cdi)); V4 cdim = at(di_moments, cs.w, p);
        V4 cgi = at(gi_colors, cs.w, p); float cgi_luma = luma(xyz(cgi)); V4 cgim = at(gi_moments, cs.w, p);
        if (surface_is_sky(csf)) { at(di_out, cs.w, p) = cdi; at(gi_out, cs.w, p) = cgi; continue; }
        float di_var, gi_var;
        if (cdim.x >= 4.0f) {
            di_var = cdim.z - sqr(cdim.y);
            gi_var = cgim.z - sqr(cgim.y);
        } else {
            V3 sum_di = v3s(0), sum_gi = v3s(0);
            IV2 off = iv2(-2, -2);
            for (;;) {
                IV2 sp = iv2((i32)p.x + off.x, (i32)p.y + off.y);
                if (camera_contains(cam, sp)) {
                    UV2 spu = uv2((u32)sp.x, (u32)sp.y);
                    Surface ssf = surface_get(surf.data(), cs.w, spu);
                    if (!surface_is_sky(ssf)) {
                        float sdl = luma(xyz(at(di_colors, cs.w, spu)));
                        float wdi = svgf_sample_weight(cdi_luma, csf, sdl, ssf, 1.0f, 0.2f);
                        sum_di += v3(sdl, sdl * sdl, 1.0f) * v3s(wdi);
                        float sgl = luma(xyz(at(gi_colors, cs.w, spu)));
                        float wgi = svgf_sample_weight(cgi_luma, csf, sgl, ssf, 1.0f, 0.2f);
                        sum_gi += v3(sgl, sgl * sgl, 1.0f) * v3s(wgi);
                    }
                }
                off.x += 1;
                if (off.x == 3) { off.x = -3; off.y += 1; if (off.y == 3) break; }   // quirk C-3
            }
            { float m1 = sum_di.x / sum_di.z, m2 = sum_di.y / sum_di.z; di_var = abs_(m2 - m1 * m1) * 4.0f; }
            { float m1 = sum_gi.x / sum_gi.z, m2 = sum_gi.y / sum_gi.z; gi_var = abs_(m2 - m1 * m1) * 4.0f; }
        }
        di_var = fmax_(di_var, 0.0f); gi_var = fmax_(gi_var, 0.0f);
        at(di_out, cs.w, p) = v4(xyz(cdi), di_var);
        at(gi_out, cs.w, p) = v4(xyz(cgi), gi_var);
    }
}

// K22 frame_denoising::wavelet (frame_denoising.rs:220-361)
static void pass_denoise_wavelet(CamState& cs, const Scene& sc, bool alternate, u32 frame, u32 stride, float strength,
                                 const Buf& di_in, Buf& di_out, const Buf& gi_in, Buf& gi_out) {
    const Buf& surf = cs.prim_surface_map[alternate ? 1 : 0];
    const Camera& cam = cs.curr_camera;
    ORC_FOR_FULL_GRID(cs) {
        UV2 p = uv2(gx_, gy_);
        V4 bn = bnoise_texel(sc.blue_noise, p, frame);
        Surface csf = surface_get(surf.data(), cs.w, p);
        V4 cdi = at(di_in, cs.w, p);
        V3 cdi_color = xyz(cdi); float cdi_var = cdi.w; float cdi_luma = luma(cdi_color);
        if (surface_is_sky(csf)) { at(di_out, cs.w, p) = v4(cdi_color, cdi_var); continue; }
        V4 cgi = at(gi_in, cs.w, p);
        V3 cgi_color = xyz(cgi); float cgi_var = cgi.w; float cgi_luma = luma(cgi_color);
        float luma_sigma_di = lerp_c(2.5f, 0.5f, sqrt_(cdi_var));
        float depth_sigma_di = 0.33f / strength;
        float luma_sigma_gi = lerp_c(1.0f, 0.0f, sqrt_(cgi_var));
        float depth_sigma_gi = 0.33f / strength;
        V2 jf = (v2(bn.z, bn.w) - v2(0.5f, 0.5f)) * ((float)stride - 1.0f) * 0.5f;
        IV2 jitter = iv2(f2i_sat(jf.x), f2i_sat(jf.y));
        float sum_di_w = 1.0f; V3 sum_di_c = cdi_color; float sum_di_v = cdi_var;
        float sum_gi_w = 1.0f; V3 sum_gi_c = cgi_color; float sum_gi_v = cgi_var;
        IV2 off = iv2(-1, -1);
        for (;;) {
            IV2 sp = iv2((i32)p.x + jitter.x + off.x * (i32)stride, (i32)p.y + jitter.y + off.y * (i32)stride);
            if (camera_contains(cam, sp) && !(off.x == 0 && off.y == 0)) {
                UV2 spu = uv2((u32)sp.x, (u32)sp.y);
                Surface ssf = surface_get(surf.data(), cs.w, spu);
                if (!surface_is_sky(ssf)) {
                    V4 sdi = at(di_in, cs.w, spu);
                    float wdi = svgf_sample_weight(cdi_luma, csf, luma(xyz(sdi)), ssf, luma_sigma_di, depth_sigma_di);
                    if (wdi > 0.0f) { sum_di_w += wdi; sum_di_c += wdi * xyz(sdi); sum_di_v += sqr(wdi) * sdi.w; }
                    V4 sgi = at(gi_in, cs.w, spu);
                    float wgi = svgf_sample_weight(cgi_luma, csf, luma(xyz(sgi)), ssf, luma_sigma_gi, depth_sigma_gi);
                    if (wgi > 0.0f) { sum_gi_w += wgi; sum_gi_c += wgi * xyz(sgi); sum_gi_v += sqr(wgi) * sgi.w; }
                }
            }
            off.x += 1;
            if (off.x == 2) { off.x = -1; off.y += 1; if (off.y == 2) break; }
        }
        at(di_out, cs.w, p) = v4(sum_di_c / sum_di_w, sum_di_v / (sum_di_w * sum_di_w));
        at(gi_out, cs.w, p) = v4(sum_gi_c / sum_gi_w, sum_gi_v / (sum_gi_w * sum_gi_w));
    }
}

// ---------------------------------------------------------------------------
// R2 frame_composition::fs (strolle-shaders/src/frame_composition.rs:19-82)
// writes linear HDR (the reference's target view applies its own format)
// ---------------------------------------------------------------------------
static void pass_frame_composition(CamState& cs, bool alternate, u32 camera_mode, bool denoise_di, bool denoise_gi) {
    int cur = alternate ? 1 : 0;
    const Buf& di_diff = denoise_di ? cs.di_diff_curr_colors : cs.di_diff_samples;
    const Buf& gi_diff = denoise_gi ? cs.gi_diff_curr_colors : cs.gi_diff_samples;
    ORC_FOR_FULL_GRID(cs) {
        UV2 p = uv2(gx_, gy_);
        GBufferEntry g = gbuffer_unpack(at(cs.prim_gbuffer_d0[cur], cs.w, p), at(cs.prim_gbuffer_d1[cur], cs.w, p));
        V3 color;
        switch (camera_mode) {
            case 0: {
                V3 dd = xyz(at(di_diff, cs.w, p)), ds = xyz(at(cs.di_spec_samples, cs.w, p));
                V3 gd = xyz(at(gi_diff, cs.w, p)), gs = xyz(at(cs.gi_spec_samples, cs.w, p));
                if (gbuffer_is_some(g)) color = g.emissive + (dd + gd) * xyz(g.base_color) + ds + gs;
                else color = dd;
                break;
            }
            case 1: color = xyz(at(di_diff, cs.w, p)); break;
            case 2: color = xyz(at(cs.di_spec_samples, cs.w, p)); break;
            case 3: color = xyz(at(gi_diff, cs.w, p)); break;
            case 4: color = xyz(at(cs.gi_spec_samples, cs.w, p)); break;
            case 5: color = xyz(at(cs.ref_colors, cs.w, p)); break;
            case 6: { V4 c = at(cs.ref_colors, cs.w, p); color = xyz(c) / c.w; break; }
            default: color = v3s(0);
        }
        at(cs.output, cs.w, p) = v4(color, 1.0f);
    }
}

// ---------------------------------------------------------------------------
// K1 ref_tracing::main (strolle-shaders/src/ref_tracing.rs:4-60)
// ---------------------------------------------------------------------------
static void pass_ref_tracing(CamState& cs, const Scene& sc, u32 depth) {
    const Camera& cam = cs.curr_camera;
    ORC_FOR_FULL_GRID(cs) {
        UV2 p = uv2(gx_, gy_);
        size_t idx = camera_screen_to_idx(cam, p);
        Ray ray;
        if (depth == 0) ray = camera_ray(cam, p);
        else {
            V4 d0 = cs.ref_rays[3 * idx], d1 = cs.ref_rays[3 * idx + 1];
            if (is_zero(d1)) continue;
            ray = ray_new(xyz(d0), xyz(d1));
        }
        TriangleHit h = ray_trace(ray, sc);
        trihit_pack(h, &cs.ref_hits[2 * idx], &cs.ref_hits[2 * idx + 1]);
    }
}

// K2 ref_shading::main (strolle-shaders/src/ref_shading.rs:4-177)
static void pass_ref_shading(CamState& cs, const Scene& sc, u32 seed, u32 depth) {
    const Camera& cam = cs.curr_camera;
    V3 sun_dir = world_sun_dir(sc.world);
    bool cam_eq = camera_is_eq(cs.curr_camera, cs.prev_camera);
    ORC_FOR_FULL_GRID(cs) {
        UV2 p = uv2(gx_, gy_);
        size_t idx = camera_screen_to_idx(cam, p);
        WhiteNoise wn = wnoise_new(seed, p);
        V4* rays = cs.ref_rays.data();
        if (depth == 255) {
            V4 prev_color = cam_eq ? at(cs.ref_colors, cs.w, p) : v4z();
            V3 curr = xyz(rays[3 * idx + 2]);
            at(cs.ref_colors, cs.w, p) = prev_color + v4(curr, 1.0f);
            continue;
        }
        Ray ray; V3 color, throughput;
        if (depth == 0) { ray = camera_ray(cam, p); color = v3s(0); throughput = v3s(1.0f); }
        else {
            V4 d0 = rays[3 * idx], d1 = rays[3 * idx + 1], d2 = rays[3 * idx + 2];
            // Deviation C-19: a dead path (d1 == 0, the marker K1 also tests, ref_tracing.rs:39-41) is skipped.
            // The reference shades it with a zero direction and relies on 0 * sample(NaN uv) == 0 on GPUs;
            // in strict IEEE arithmetic that is NaN.  The net effect on a GPU is a no-op, restated here.
            if (is_zero(d1)) continue;
            ray = ray_new(xyz(d0), xyz(d1)); color = xyz(d2); throughput = v3(d0.w, d1.w, d2.w);
        }
        TriangleHit th = trihit_unpack(cs.ref_hits[2 * idx], cs.ref_hits[2 * idx + 1]);
        if (!trihit_is_some(th)) {
            color += throughput * atmosphere_sample(sc, sun_dir, ray.dir);
            rays[3 * idx] = v4z(); rays[3 * idx + 1] = v4z(); rays[3 * idx + 2] = v4(color, 0.0f);
            continue;
        }
        Material material = sc.materials[th.material_id];
        if (depth > 0) material_regularize(material);
        Hit hit;
        hit.point = th.point + th.normal * 0.01f; hit.origin = ray.origin; hit.dir = ray.dir;
        hit.gbuffer.base_color = material_base_color(sc, material, th.uv); hit.gbuffer.normal = th.normal; hit.gbuffer.metallic = material.metallic;
        hit.gbuffer.emissive = material_emissive(sc, material, th.uv); hit.gbuffer.roughness = material.roughness;
        hit.gbuffer.reflectance = material.reflectance; hit.gbuffer.depth = 0.0f;
        color += throughput * hit.gbuffer.emissive;
        if (sc.world.light_count > 0) {
            u32 light_id = wnoise_sample_int(wn) % sc.world.light_count;
            float light_pdf = 1.0f / (float)sc.world.light_count;
            const Light& light = sc.lights[light_id];
            bool occluded = ray_intersect(light_ray_wnoise(light, wn, hit.point), sc);
            if (!occluded) color += throughput * light_radiance_sum(light_radiance(light, hit)) / light_pdf;
        }
        BrdfSample rs = layered_brdf_sample(hit.gbuffer, wn, -hit.dir);
        if (rs.pdf == 0.0f) { rays[3 * idx] = v4z(); rays[3 * idx + 1] = v4z(); continue; }
        Ray rr = ray_new(hit.point, rs.dir);
        throughput *= dot(rs.dir, hit.gbuffer.normal);
        throughput *= rs.radiance / rs.pdf;
        rays[3 * idx] = v4(rr.origin, throughput.x);
        rays[3 * idx + 1] = v4(rr.dir, throughput.y);
        rays[3 * idx + 2] = v4(color, throughput.z);
    }
}

// K3 bvh_heatmap::main (strolle-shaders/src/bvh_heatmap.rs:4-77)
static inline V3 heat_gradient(float progress) {
    const V3 colors[4] = {v3(0, 0, 1), v3(0, 1, 0), v3(1, 0, 0), v3(0, 0, 0)};
    if (progress <= 0.0f) return colors[0];
    float step = 1.0f / (4.0f - 1.0f);
    for (int i = 0; i < 3; i++) {
        float mn = step * (float)i, mx = step * ((float)i + 1.0f);
        if (progress >= mn && progress <= mx) { float rhs = (progress - mn) / step; float lhs = 1.0f - rhs; return lhs * colors[i] + rhs * colors[i + 1]; }
    }
    return colors[3];
}
static void pass_bvh_heatmap(CamState& cs, const Scene& sc) {
    ORC_FOR_FULL_GRID(cs) {
        UV2 p = uv2(gx_, gy_);
        size_t used = 0;
        ray_trace(camera_ray(cs.curr_camera, p), sc, &used);
        at(cs.ref_colors, cs.w, p) = v4(heat_gradient((float)used / 8192.0f), 1.0f);
    }
}

}  // namespace orc
