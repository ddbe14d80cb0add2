// gsr_collide.cuh -- collision detection of one environment by one lane group (collider.py:684-808).
//
//   broad phase : the lanes sweep the static valid-pair table in chunks of G pairs, test the world AABBs (shared memory)
//                 and append the overlapping pairs, in pair order, to a shared-memory work queue through a ballot /
//                 popcount prefix (warp-level compaction);
//   narrow phase: whenever G candidates are queued (and once more at the end) every lane takes one pair and runs the
//                 reference's per-pair routine (plane support queries, plane-box corners, MPR with warm-started portal,
//                 up to 4 perturbed multi-contact probes) with its simplex in registers; the contacts of a batch are
//                 compacted with a prefix sum over the lanes into the environment's staging buffer;
//   ordering    : every contact carries the key of the reference's canonical order (CT:439-538: x of the first contact of
//                 its pair, then convex-before-plane-box, pair index, detection index); each lane ranks its contact
//                 against the others and writes it to its final slot of the CONTACT record (and of the on-chip copy).
#pragma once
#include "gsr_group.cuh"
#include "gsr_gjk.cuh"
#undef IN
#undef WK
#if GSR_LDS_MASK & 2
#define IN(field) IN_SHARED(field)
#define WK(field) WK_SHARED(field)
#else
#define IN(field) IN_GENERIC(field)
#define WK(field) WK_GENERIC(field)
#endif

// SF:138-180 support point of a mesh-backed geom from its 180x180 direction table (L2 resident, read-only path)
template <int G>
__device__ __noinline__ f3 dev_support_mesh(const Ctx<G>& C, f3 d, int g) {
  const float PI = 3.14159265358979323846f;
  const float theta = atan2f(d.y, d.x), phi = acosf(d.z);
  const int res = 180;
  const float ii = (theta + PI) / PI / 2.f * res, jj = phi / PI * res;
  float dot_max = -1e20f;
  f3 best;
  const float* tab = CM(support_v) + 3 * (size_t)GM(g_support_start[g]);
#pragma unroll
  for (int i4 = 0; i4 < 4; i4++) {
    int i = (i4 & 1) ? ((int)ceilf(ii)) % res : ((int)floorf(ii)) % res;
    int j;
    if (i4 >> 1) { j = (int)fminf(fmaxf(ceilf(jj), 0.f), (float)(res - 1)); if (j == res - 1) j = res - 2; }
    else { j = (int)fminf(fmaxf(floorf(jj), 0.f), (float)(res - 1)); if (j == 0) j = 1; }
    const f3 v = g3(tab + 3 * (i * res + j));
    const float dt = dot(v, d);
    if (dt > dot_max) { best = v; dot_max = dt; }
  }
  return best;
}

// the same lookup with the identity of the winning vertex (support_field.py support_vid: index inside the geom), for GJK
template <int G>
__device__ __noinline__ f3 dev_support_mesh_id(const Ctx<G>& C, f3 d, int g, int& vid) {
  const float PI = 3.14159265358979323846f;
  const float theta = atan2f(d.y, d.x), phi = acosf(d.z);
  const int res = 180;
  const float ii = (theta + PI) / PI / 2.f * res, jj = phi / PI * res;
  float dot_max = -1e20f;
  f3 best;
  const int start = GM(g_support_start[g]);
  const float* tab = CM(support_v) + 3 * (size_t)start;
  const int* ids = CM(support_vid) + start;
  vid = 0;
  for (int i4 = 0; i4 < 4; i4++) {
    int i = (i4 & 1) ? ((int)ceilf(ii)) % res : ((int)floorf(ii)) % res;
    int j;
    if (i4 >> 1) { j = (int)fminf(fmaxf(ceilf(jj), 0.f), (float)(res - 1)); if (j == res - 1) j = res - 2; }
    else { j = (int)fminf(fmaxf(floorf(jj), 0.f), (float)(res - 1)); if (j == 0) j = 1; }
    const f3 v = g3(tab + 3 * (i * res + j));
    const float dt = dot(v, d);
    if (dt > dot_max) { best = v; dot_max = dt; vid = __ldg(ids + i * res + j); }
  }
  return best;
}

// MPR:146-176 support_driver
template <int G>
__device__ __noinline__ f3 dev_support(const Ctx<G>& C, f3 d, int g, f3 pos, f4 quat) {
  const int type = GM(g_type[g]);
  if (type == GEOM_SPHERE) return pos + d * GM(g_data[g][0]);
  if (type == GEOM_BOX) {
    const f3 db = rotate_inv(d, quat);
    const f3 v((db.x < 0.f ? -1.f : 1.f) * GM(g_data[g][0]) * 0.5f, (db.y < 0.f ? -1.f : 1.f) * GM(g_data[g][1]) * 0.5f,
               (db.z < 0.f ? -1.f : 1.f) * GM(g_data[g][2]) * 0.5f);
    return transform(v, pos, quat);
  }
  if (type == GEOM_CAPSULE) {
    const float r = GM(g_data[g][0]), hl = 0.5f * GM(g_data[g][1]);
    const f3 axis = rotate(f3(0.f, 0.f, 1.f), quat);
    const float side = dot(d, axis) < 0.f ? -1.f : 1.f;
    return pos + axis * (hl * side) + d * r;
  }
  if (type == GEOM_ELLIPSOID) {
    const float a = GM(g_data[g][0]), b = GM(g_data[g][1]), c = GM(g_data[g][2]);
    const f3 dl = rotate_inv(d, quat);
    const float n = sqrtf(a * a * dl.x * dl.x + b * b * dl.y * dl.y + c * c * dl.z * dl.z);
    return pos + rotate(f3(a * a * dl.x / n, b * b * dl.y / n, c * c * dl.z / n), quat);
  }
  return transform(dev_support_mesh(C, rotate_inv(d, quat), g), pos, quat);
}

// ---- MPR (mpr.py) ---------------------------------------------------------------------------------
struct Simplex {
  f3 v[4], v1[4], v2[4];
};
__device__ __forceinline__ void simplex_set(Simplex& s, int i, f3 v, f3 v1, f3 v2) {
  // explicit selects keep the simplex in registers (no dynamically indexed local memory)
  if (i == 1) { s.v[1] = v; s.v1[1] = v1; s.v2[1] = v2; }
  else if (i == 2) { s.v[2] = v; s.v1[2] = v1; s.v2[2] = v2; }
  else { s.v[3] = v; s.v1[3] = v1; s.v2[3] = v2; }
}
struct PairPose { int ga, gb; f3 pa, pb; f4 qa, qb; };
template <int G>
__device__ __forceinline__ void mpr_support(const Ctx<G>& C, const PairPose& P, f3 d, f3& v, f3& v1, f3& v2) {
  v1 = dev_support(C, d, P.ga, P.pa, P.qa);
  v2 = dev_support(C, -d, P.gb, P.pb, P.qb);
  v = v1 - v2;
}
template <int G>
struct GroupSupport {  // support functor of the pair for gsr_epa.cuh
  const Ctx<G>& C;
  const PairPose& P;
  __device__ __forceinline__ void operator()(f3 d, f3& v, f3& v1, f3& v2) const { mpr_support(C, P, d, v, v1, v2); }
};
// the pair as gsr_gjk.cuh sees it (gjk_support.py:66-112 support_driver, gjk.py:1648-1698 discrete vertices, :1841-1900 counts)
template <int G>
struct GroupGeo {
  const Ctx<G>& C;
  const PairPose& P;
  int ta, tb;
  __device__ __forceinline__ float eps() const { return c_m.eps; }
  __device__ __forceinline__ bool discrete() const { return (ta == GEOM_MESH || ta == GEOM_BOX) && (tb == GEOM_MESH || tb == GEOM_BOX); }
  __device__ void support(f3 d, int which, f3& v, int& vid) const {
    const int g = which ? P.gb : P.ga, type = which ? tb : ta;
    const f3 pos = which ? P.pb : P.pa;
    const f4 quat = which ? P.qb : P.qa;
    vid = -1;
    if (type == GEOM_BOX) {
      const f3 db = rotate_inv(d, quat);
      const f3 vl((db.x < 0.f ? -1.f : 1.f) * GM(g_data[g][0]) * 0.5f, (db.y < 0.f ? -1.f : 1.f) * GM(g_data[g][1]) * 0.5f,
                  (db.z < 0.f ? -1.f : 1.f) * GM(g_data[g][2]) * 0.5f);
      vid = (vl.x > 0.f ? 1 : 0) + (vl.y > 0.f ? 2 : 0) + (vl.z > 0.f ? 4 : 0) + GM(g_vert_start[g]);
      v = transform(vl, pos, quat);
    } else if (type == GEOM_SPHERE || type == GEOM_ELLIPSOID || type == GEOM_CAPSULE) {
      v = dev_support(C, d, g, pos, quat);
    } else {
      v = transform(dev_support_mesh_id(C, rotate_inv(d, quat), g, vid), pos, quat);
    }
  }
  __device__ int count_support(f3 dir) const {  // boxes: 2^(zero components); the reference's mesh count always answers 1
    int count = 1;
    if (ta == GEOM_BOX) { const f3 db = rotate_inv(dir, P.qa); count <<= (db.x == 0.f) + (db.y == 0.f) + (db.z == 0.f); }
    if (tb == GEOM_BOX) { const f3 db = rotate_inv(-dir, P.qb); count <<= (db.x == 0.f) + (db.y == 0.f) + (db.z == 0.f); }
    return count;
  }
  __device__ __forceinline__ int n_verts(int which) const { const int g = which ? P.gb : P.ga; return GM(g_vert_end[g]) - GM(g_vert_start[g]); }
  __device__ void vertex(int which, int iv, f3& v, int& id) const {
    const int g = which ? P.gb : P.ga, type = which ? tb : ta;
    f3 vl;
    if (type == GEOM_BOX) vl = f3(((iv & 1) ? 1.f : -1.f) * GM(g_data[g][0]) * 0.5f, ((iv & 2) ? 1.f : -1.f) * GM(g_data[g][1]) * 0.5f,
                                  ((iv & 4) ? 1.f : -1.f) * GM(g_data[g][2]) * 0.5f);
    else vl = g3(CM(verts) + 3 * (size_t)(GM(g_vert_start[g]) + iv));
    id = GM(g_vert_start[g]) + iv;
    v = transform(vl, which ? P.pb : P.pa, which ? P.qb : P.qa);
  }
};
__device__ __forceinline__ f3 portal_dir(const Simplex& s) { return normalized(cross(s.v[2] - s.v[1], s.v[3] - s.v[1])); }
__device__ __forceinline__ bool reach_tolerance(const Simplex& s, f3 v, f3 dir) {  // MPR:134-143
  const float dv4 = dot(v, dir);
  const float dot1 = fminf(fminf(dv4 - dot(s.v[1], dir), dv4 - dot(s.v[2], dir)), dv4 - dot(s.v[3], dir));
  return dot1 < c_m.ccd_tolerance + c_m.ccd_eps * fmaxf(1.0f, dot1);
}
__device__ __forceinline__ void expand_portal(Simplex& s, f3 v, f3 v1, f3 v2) {  // MPR:426-442
  const f3 v4v0 = cross(v, s.v[0]);
  int is;
  if (dot(s.v[1], v4v0) > 0.f) is = dot(s.v[2], v4v0) > 0.f ? 1 : 3;
  else is = dot(s.v[3], v4v0) > 0.f ? 2 : 1;
  simplex_set(s, is, v, v1, v2);
}
// MPR:445-598.  `sup(d, v, v1, v2)` is the support mapping of the pair: GroupSupport for two posed geoms, TerrainSupport for a geom
// against one prism of the height field.
template <int G, class Sup>
__device__ int discover_portal(const Ctx<G>& C, Simplex& s, const Sup& sup, f3 ca, f3 cb) {
  const float CE = c_m.ccd_eps;
  s.v1[0] = ca; s.v2[0] = cb; s.v[0] = ca - cb;
  if (fabsf(s.v[0].x) < CE && fabsf(s.v[0].y) < CE && fabsf(s.v[0].z) < CE) s.v[0].x += 10.0f * CE;
  f3 dir = -normalized(s.v[0]);
  f3 v, v1, v2;
  sup(dir, v, v1, v2);
  s.v[1] = v; s.v1[1] = v1; s.v2[1] = v2;
  if (dot(v, dir) < CE) return -1;
  dir = cross(s.v[0], s.v[1]);
  if (dot(dir, dir) < CE) {
    if (fabsf(s.v[1].x) < CE && fabsf(s.v[1].y) < CE && fabsf(s.v[1].z) < CE) return 1;
    return 2;
  }
  dir = normalized(dir);
  sup(dir, v, v1, v2);
  if (dot(v, dir) < CE) return -1;
  s.v[2] = v; s.v1[2] = v1; s.v2[2] = v2;
  dir = normalized(cross(s.v[1] - s.v[0], s.v[2] - s.v[0]));
  if (dot(dir, s.v[0]) > 0.f) {
    f3 t;
    t = s.v[1]; s.v[1] = s.v[2]; s.v[2] = t;
    t = s.v1[1]; s.v1[1] = s.v1[2]; s.v1[2] = t;
    t = s.v2[1]; s.v2[1] = s.v2[2]; s.v2[2] = t;
    dir = -dir;
  }
  int num_trials = 0;
  while (true) {
    sup(dir, v, v1, v2);
    if (dot(v, dir) < CE) return -1;
    bool cont = false;
    if (dot(cross(s.v[1], v), s.v[0]) < -CE) { s.v[2] = v; s.v1[2] = v1; s.v2[2] = v2; cont = true; }
    if (!cont && dot(cross(v, s.v[2]), s.v[0]) < -CE) { s.v[1] = v; s.v1[1] = v1; s.v2[1] = v2; cont = true; }
    if (cont) {
      dir = normalized(cross(s.v[1] - s.v[0], s.v[2] - s.v[0]));
      if (++num_trials == 15) return -1;
    } else {
      s.v[3] = v; s.v1[3] = v1; s.v2[3] = v2;
      return 0;
    }
  }
}
// MPR:686-765 func_mpr_contact_from_centers (portal discovery / refinement + penetration)
template <int G, class Sup>
__device__ __forceinline__ void dev_mpr_from_centers(const Ctx<G>& C, const Sup& sup, f3 ca, f3 cb, bool& is_col, f3& normal, float& pen, f3& pos,
                                                     Simplex& s, bool& tetra) {
  const float CE = c_m.ccd_eps;
  const long long mt0 = NARROW_CLOCK();
  int n_refine = 0;
  int res = discover_portal(C, s, sup, ca, cb);
  const long long mt1 = NARROW_CLOCK();
  is_col = false; pos = f3(); normal = f3(); pen = 0.f;
  tetra = false;  // true when s ends as the interior point + final portal, a tetrahedron around the origin (EPA seed)
  if (res == 1) { is_col = true; normal = -normalized(s.v[0]); pos = (s.v1[1] + s.v2[1]) * 0.5f; }
  else if (res == 2) { is_col = true; pen = norm(s.v[1]); normal = -normalized(s.v[1]); pos = (s.v1[1] + s.v2[1]) * 0.5f; }
  else if (res == 0) {
    while (true) {  // MPR:232-278 refine portal
      const f3 dir = portal_dir(s);
      if (dot(s.v[1], dir) > -CE) { res = 0; break; }
      f3 v, v1, v2;
      sup(dir, v, v1, v2);
      if (!(dot(v, dir) > -CE) || reach_tolerance(s, v, dir)) { res = -1; break; }
      expand_portal(s, v, v1, v2);
      n_refine++;
    }
    if (res >= 0) {
      int it = 0;
      while (true) {  // MPR:338-423 find penetration
        const f3 dir = portal_dir(s);
        f3 v, v1, v2;
        sup(dir, v, v1, v2);
        if (reach_tolerance(s, v, dir) || it > c_m.ccd_iterations) {
          pen = dot(dir, s.v[1]);
          normal = -dir;
          is_col = true;
          // MPR:281-316 find_pos: barycentric weights of the origin's projection on the portal
          const float b1 = dot(cross(s.v[2], s.v[3]), dir), b2 = dot(cross(s.v[3], s.v[1]), dir), b3 = dot(cross(s.v[1], s.v[2]), dir);
          const float sum = b1 + b2 + b3;
          const f3 p1 = s.v1[1] * b1 + s.v1[2] * b2 + s.v1[3] * b3, p2 = s.v2[1] * b1 + s.v2[2] * b2 + s.v2[3] * b3;
          pos = (p1 + p2) * (0.5f / sum);
          tetra = true;
          break;
        }
        expand_portal(s, v, v1, v2);
        it++;
      }
      NARROW_ADD(5, it); NARROW_MAX(6, it);
    }
  }
  // developer counters (GSR_NARROW_TRACE, tools/go2_tail.py): queries, cycles, portal discovery cycles, refinement / penetration iterations
  NARROW_ADD(0, 1); NARROW_ADD(1, NARROW_CLOCK() - mt0); NARROW_MAX(2, NARROW_CLOCK() - mt0); NARROW_ADD(3, n_refine); NARROW_MAX(4, n_refine);
  NARROW_ADD(7, mt1 - mt0); NARROW_MAX(8, mt1 - mt0);
  (void)mt0; (void)mt1; (void)n_refine;
}

// MPR:601-819 func_mpr_contact (centre guess + portal discovery/refinement + penetration)
template <int G>
__device__ __noinline__ void dev_mpr_contact(const Ctx<G>& C, const PairPose& P, f3 normal_ws, bool& is_col, f3& normal, float& pen, f3& pos,
                                             Simplex& s, bool& tetra) {
  const float CE = c_m.ccd_eps;
  f3 ca = transform(f3(GM(g_center[P.ga][0]), GM(g_center[P.ga][1]), GM(g_center[P.ga][2])), P.pa, P.qa);
  f3 cb = transform(f3(GM(g_center[P.gb][0]), GM(g_center[P.gb][1]), GM(g_center[P.gb][2])), P.pb, P.qb);
  if (fabsf(normal_ws.x) > CE || fabsf(normal_ws.y) > CE || fabsf(normal_ws.z) > CE) {
    const f3 loa(GM(g_aabb_lo[P.ga][0]), GM(g_aabb_lo[P.ga][1]), GM(g_aabb_lo[P.ga][2])), hia(GM(g_aabb_hi[P.ga][0]), GM(g_aabb_hi[P.ga][1]), GM(g_aabb_hi[P.ga][2]));
    const f3 lob(GM(g_aabb_lo[P.gb][0]), GM(g_aabb_lo[P.gb][1]), GM(g_aabb_lo[P.gb][2])), hib(GM(g_aabb_hi[P.gb][0]), GM(g_aabb_hi[P.gb][1]), GM(g_aabb_hi[P.gb][2]));
    ca = transform((hia + loa) * 0.5f, P.pa, P.qa);
    cb = transform((hib + lob) * 0.5f, P.pb, P.qb);
    const f3 delta = ca - cb;
    if (norm(cross(normal_ws, normalized(delta))) > 0.01f) {
      const f3 offset = normal_ws * dot(delta, normal_ws) - delta;
      const float on = norm(offset);
      if (on > c_m.eps) {
        const f3 dir = offset / on;
        const f3 dla = rotate_inv(dir, P.qa), dlb = rotate_inv(dir, P.qb);
        const f3 bsa = hia - loa, bsb = hib - lob;
        const float la = bsa.x * fabsf(dla.x) + bsa.y * fabsf(dla.y) + bsa.z * fabsf(dla.z);
        const float lb = bsb.x * fabsf(dlb.x) + bsb.y * fabsf(dlb.y) + bsb.z * fabsf(dlb.z);
        const float ratio = fminf(on / (la + lb), 0.5f);
        ca = ca + dir * (la * ratio);
        cb = cb - dir * (lb * ratio);
      }
    }
  }
  dev_mpr_from_centers(C, GroupSupport<G>{C, P}, ca, cb, is_col, normal, pen, pos, s, tetra);
}

// contacts found by one lane for one pair (registers / local memory)
struct LaneContacts {
  int n;
  float pen[5];
  f3 pos[5], nrm[5];
  __device__ __forceinline__ void add(f3 normal, f3 p, float pe) {
    if (n < 5) { nrm[n] = normal; pos[n] = p; pen[n] = pe; n++; }
  }
};

// CT:317-336
template <int G>
__device__ __forceinline__ float dev_tolerance(const Ctx<G>& C, int ga, int gb) {
  float size = norm(f3(GM(g_aabb_hi[gb][0]) - GM(g_aabb_lo[gb][0]), GM(g_aabb_hi[gb][1]) - GM(g_aabb_lo[gb][1]), GM(g_aabb_hi[gb][2]) - GM(g_aabb_lo[gb][2])));
  if (GM(g_type[ga]) != GEOM_PLANE)
    size = fminf(norm(f3(GM(g_aabb_hi[ga][0]) - GM(g_aabb_lo[ga][0]), GM(g_aabb_hi[ga][1]) - GM(g_aabb_lo[ga][1]), GM(g_aabb_hi[ga][2]) - GM(g_aabb_lo[ga][2]))), size);
  return 0.5f * c_m.mc_tolerance * size;
}
// CT:415-436
__device__ __forceinline__ void rotate_frame(f3 pos, f4 quat, f3 cpos, f4 qrot, f3& npos, f4& nquat) {
  nquat = quat_compose(quat, qrot);
  const f3 rel = cpos - pos;
  npos = pos - (rotate(rel, qrot) - rel);
}

// NP:519-970 func_convex_convex_contact for one broad-phase pair (type[ga] <= type[gb], not plane-box), split in two so that
// the four perturbed multi-contact detections of a pair -- which depend only on the unperturbed one -- run on four lanes
// instead of back to back on the pair's lane (the MPR of an unusual pose is the long pole of a launch, profiles/README.md).
struct PairDet {
  int ga, gb, ta, tb, p;
  bool multi, is_col_0;
  float tolerance;
  f3 pa0, pb0, normal_0, cpos_0, axis_0, axis_1;
  f4 qa0, qb0;
};

// one detection at the (possibly perturbed) poses in P: analytic capsule pairs, plane support query, or MPR with warm start
// MPR is a level: 0 compiles the portal refinement and the GJK / EPA fallback out, 1 = the default, 2 = the full variant (GJK-first mode)
// MPR = 0 compiles the portal refinement and the EPA fallback out: models without a convex-convex pair that needs them
// (gsr_host.h has_mpr; e.g. a robot without self-collision on a plane) run a kernel a few thousand instructions shorter
template <int G, int MPR>
__device__ __forceinline__ void dev_detect(const Ctx<G>& C, const PairDet& D, const PairPose& P, int i_det, bool& is_col, f3& normal, float& pen, f3& cpos) {
  const float EPS = c_m.eps;
  const int ga = D.ga, gb = D.gb;
  if (D.ta == GEOM_CAPSULE && D.tb == GEOM_CAPSULE) {  // NP:637-648 analytic capsule pairs
    capsule_capsule_contact(P.pa, P.qa, GM(g_data[ga][0]), 0.5f * GM(g_data[ga][1]), P.pb, P.qb, GM(g_data[gb][0]), 0.5f * GM(g_data[gb][1]), EPS,
                            is_col, normal, cpos, pen);
  } else if (D.ta == GEOM_SPHERE && D.tb == GEOM_CAPSULE) {  // NP:649-661 (types are sorted: A is the sphere)
    sphere_capsule_contact(P.pa, GM(g_data[ga][0]), P.pb, P.qb, GM(g_data[gb][0]), 0.5f * GM(g_data[gb][1]), EPS, is_col, normal, cpos, pen);
  } else if (D.ta == GEOM_PLANE) {
    normal = -normalized(rotate(f3(GM(g_data[ga][0]), GM(g_data[ga][1]), GM(g_data[ga][2])), P.qa));
    const f3 v1 = dev_support(C, normal, gb, P.pb, P.qb);
    pen = dot(normal, v1 - P.pa);
    cpos = v1 - normal * (0.5f * pen);
    is_col = pen > 0.f;
  } else if constexpr (!MPR) {
    is_col = false; normal = f3(); pen = 0.f; cpos = f3();  // unreachable: the host launches this variant only without such pairs
  } else {
    bool run_gjk = false;  // RigidOptions.use_gjk_collision: no MPR at all (NP:604-607); compiled into the full variants (level 2) only
    if constexpr (MPR == 2) run_gjk = c_m.use_gjk != 0;
    if (!run_gjk) {
      // the perturbed detections read the normal the unperturbed one has just cached (same as the sequential reference flow)
      int* cvalid = reinterpret_cast<int*>(C.cache) + c_h.cache_valid;
      f3 normal_ws;
      const bool valid = (cvalid[D.p >> 5] >> (D.p & 31)) & 1;
      if (valid) normal_ws = ld3(C.cache, D.p);
      const bool guess = fabsf(normal_ws.x) > EPS || fabsf(normal_ws.y) > EPS || fabsf(normal_ws.z) > EPS;
      Simplex sx;
      bool tetra = false, guess_used = guess;
      dev_mpr_contact(C, P, normal_ws, is_col, normal, pen, cpos, sx, tetra);
      if (i_det == 0 && !is_col && guess) { dev_mpr_contact(C, P, f3(), is_col, normal, pen, cpos, sx, tetra); guess_used = false; }
      // NP:726-733: MPR's answer is replaced by the exact one when the overlap is large or no direction was cached
      if (c_m.gjk_fallback && is_col && pen > D.tolerance && (!guess_used || c_m.mc_tolerance * pen >= 0.25f * D.tolerance)) {
#ifndef GSR_DEV_NO_GJK  // developer A/B builds: what does compiling the exact-penetration path in cost a scene that never runs it?
        run_gjk = true;  // (the round-1 approximation, EPA seeded with MPR's portal, is gone from this kernel: carrying it cost Go2 time)
#endif
      }
    }
#ifndef GSR_DEV_NO_GJK
    if (run_gjk) {  // the reference's path: a fresh safe-GJK query and safe EPA from its tetrahedron (gsr_gjk.cuh)
      const GroupGeo<G> geo{C, P, D.ta, D.tb};
      gjk_contact(geo, is_col, normal, pen, cpos);
    }
#endif
  }
}

// unperturbed detection of a pair: contact 0, warm-start cache update, perturbation axes
template <int G, int MPR>
__device__ __noinline__ void dev_pair_det0(const Ctx<G>& C, LaneContacts& out, PairDet& D) {
  const float EPS = c_m.eps;
  const int ga = D.ga, gb = D.gb;
  D.ta = GM(g_type[ga]); D.tb = GM(g_type[gb]);
  D.multi = c_m.enable_multi_contact && D.ta != GEOM_SPHERE && D.ta != GEOM_ELLIPSOID && D.tb != GEOM_SPHERE && D.tb != GEOM_ELLIPSOID;
  D.tolerance = dev_tolerance(C, ga, gb);
  geom_pose(C, ga, D.pa0, D.qa0);
  geom_pose(C, gb, D.pb0, D.qb0);
  PairPose P;
  P.ga = ga; P.gb = gb; P.pa = D.pa0; P.pb = D.pb0; P.qa = D.qa0; P.qb = D.qb0;
  bool is_col;
  f3 normal, cpos;
  float pen = 0.f;
  dev_detect<G, MPR>(C, D, P, 0, is_col, normal, pen, cpos);
  D.is_col_0 = is_col; D.normal_0 = normal; D.cpos_0 = cpos;
  if (is_col) {
    out.add(normal, cpos, pen);
    if (D.multi) {
      // CT:353-412 func_contact_orthogonals: perturbation axes from the principal inertia frame of the smaller geom
      int g = gb;
      if (D.ta != GEOM_PLANE) {
        const float va = GM(g_aabb_hi[ga][0]) * GM(g_aabb_hi[ga][1]) * GM(g_aabb_hi[ga][2]);
        const float vb = GM(g_aabb_hi[gb][0]) * GM(g_aabb_hi[gb][1]) * GM(g_aabb_hi[gb][2]);
        g = (va < vb) ? ga : gb;
      }
      const int l = GM(g_link[g]);
      const mat3 rot = quat_to_R(quat_compose(f4(GM(l_iquat[l][0]), GM(l_iquat[l][1]), GM(l_iquat[l][2]), GM(l_iquat[l][3])), ld4(IN(links_quat), l)), EPS);
      int axis_idx = 0;
      float amax = 0.f;
      for (int i = 0; i < 3; i++) { const float a = fabsf(dot(mat_col(rot, i), normal)); if (a > amax) { amax = a; axis_idx = i; } }
      axis_idx = (axis_idx + 1) % 3;
      D.axis_0 = mat_col(rot, axis_idx);
      D.axis_0 = normalized(D.axis_0 - normal * dot(normal, D.axis_0));
      D.axis_1 = cross(normal, D.axis_0);
    }
  }
  if (D.ta != GEOM_PLANE) {
    int* cvalid = reinterpret_cast<int*>(C.cache) + c_h.cache_valid;
    if (is_col) { st3(C.cache, D.p, normal); atomicOr(&cvalid[D.p >> 5], 1 << (D.p & 31)); }
    else atomicAnd(&cvalid[D.p >> 5], ~(1 << (D.p & 31)));
  }
}

// perturbed detection i_det (1..4) of a colliding pair, corrected back to the unperturbed frame (NP:889-936)
template <int G, int MPR>
__device__ __noinline__ void dev_pair_perturbed(const Ctx<G>& C, const PairDet& D, int i_det, bool& is_col, f3& normal, float& pen, f3& cpos) {
  const float EPS = c_m.eps;
  const f3 axis = D.axis_0 * (float)(2 * (i_det % 2) - 1) + D.axis_1 * (float)(1 - 2 * ((i_det / 2) % 2));
  const f4 qrot = rotvec_to_quat(axis * c_m.mc_perturbation, EPS);
  PairPose P;
  P.ga = D.ga; P.gb = D.gb;
  rotate_frame(D.pa0, D.qa0, D.cpos_0, qrot, P.pa, P.qa);
  rotate_frame(D.pb0, D.qb0, D.cpos_0, quat_inv(qrot), P.pb, P.qb);
  dev_detect<G, MPR>(C, D, P, i_det, is_col, normal, pen, cpos);
  if (is_col) {
    const f3 cpa = rotate((cpos - normal * (0.5f * pen)) - D.cpos_0, quat_inv(qrot)) + D.cpos_0;
    const f3 cpb = rotate((cpos + normal * (0.5f * pen)) - D.cpos_0, qrot) + D.cpos_0;
    cpos = (cpa + cpb) * 0.5f;
    f3 twist = cross(normal, D.normal_0);
    const float mcp = c_m.mc_perturbation;
    twist = f3(fminf(fmaxf(twist.x, -mcp), mcp), fminf(fmaxf(twist.y, -mcp), mcp), fminf(fmaxf(twist.z, -mcp), mcp));
    normal = normal + cross(twist, normal);
    pen = dot(normal, cpb - cpa);
  }
}

#include "gsr_boxbox.cuh"  // BOX:100-785 func_box_box_contact (needs LaneContacts)

#include "gsr_terrain.cuh"  // NP:346-517 func_contact_mpr_terrain (needs LaneContacts, dev_tolerance, dev_mpr_from_centers)

// BOX:24-96 func_plane_box_contact
template <int G>
__device__ __noinline__ void dev_plane_box(const Ctx<G>& C, LaneContacts& out, int ga, int gb) {
  f3 pa, pb;
  f4 qa, qb;
  geom_pose(C, ga, pa, qa);
  geom_pose(C, gb, pb, qb);
  const f3 normal = -normalized(rotate(f3(GM(g_data[ga][0]), GM(g_data[ga][1]), GM(g_data[ga][2])), qa));
  const f3 v1 = dev_support(C, normal, gb, pb, qb);
  float pen = dot(normal, v1 - pa);
  if (pen > 0.f) {
    f3 cpos = v1 - normal * (0.5f * pen);
    out.add(normal, cpos, pen);
    if (c_m.enable_multi_contact) {
      int n_con = 1;
      const f3 cpos_0 = cpos;
      const float tolerance = dev_tolerance(C, ga, gb);
      const int vs = GM(g_vert_start[gb]), ve = GM(g_vert_end[gb]);
      for (int iv = vs; iv < ve; iv++) {
        if (n_con < c_m.n_contacts_per_pair) {
          const f3 pc = transform(g3(CM(verts) + 3 * (size_t)iv), pb, qb);
          pen = dot(normal, pc - pa);
          if (pen > 0.f) {
            cpos = pc - normal * (0.5f * pen);
            if (norm(cpos - cpos_0) > tolerance) { out.add(normal, cpos, pen); n_con++; }
          }
        }
      }
    }
  }
}

// narrow phase of up to G queued pairs; appends the contacts (unordered) to the staging records and returns the new count
// BB: the "full" variant -- RigidOptions.box_box_detection with a box-box pair in the table (gsr_boxbox.cuh) or a height-field
// terrain (gsr_terrain.cuh)
template <int G, bool MPR, bool BB>
__device__ int narrow_batch(const Ctx<G>& C, const int* queue, int n_batch, int n_staged) {
  const int lane = C.lane;
  LaneContacts out;
  out.n = 0;
  const long long nb0 = NARROW_CLOCK();
  int ga = 0, gb = 0, p = 0, plane_box = 0;
  PairDet D;
  D.multi = false; D.is_col_0 = false;
  if (lane < n_batch) {
    p = queue[lane];
    const int ga0 = __ldg(CM(pairs) + 2 * p), gb0 = __ldg(CM(pairs) + 2 * p + 1);
    ga = ga0; gb = gb0;
    if (GM(g_type[ga0]) > GM(g_type[gb0])) { ga = gb0; gb = ga0; }
    // convex specialisations (NP:2135-2189): their contacts follow those of the convex-convex pass in the contact list
    const bool is_plane_box = GM(g_type[ga]) == GEOM_PLANE && GM(g_type[gb]) == GEOM_BOX;
    const bool is_box_box = BB && c_m.box_box_detection && GM(g_type[ga]) == GEOM_BOX && GM(g_type[gb]) == GEOM_BOX;
    const bool is_terrain = BB && GM(g_type[gb]) == GEOM_TERRAIN;  // NP:2193-2239: after the specialisations
    plane_box = is_terrain ? 2 : ((is_plane_box || is_box_box) ? 1 : 0);
    if (is_plane_box) dev_plane_box(C, out, ga, gb);
    else if (is_box_box) { if constexpr (BB) dev_box_box(C, out, ga, gb); }
    else if (is_terrain) { if constexpr (BB) dev_terrain(C, out, ga, gb); }
    else { D.ga = ga; D.gb = gb; D.p = p; dev_pair_det0<G, (MPR ? (BB ? 2 : 1) : 0)>(C, out, D); }
  }
  // ---- perturbed detections: task (pair k of the colliding multi-contact pairs, i_det) -> lane, results through the
  // staging area behind the contact staging records, then every pair lane applies the de-duplication in detection order
  const bool want = lane < n_batch && !plane_box && D.multi && D.is_col_0;
  const unsigned wm = C.ballot(want);
  const int n_task = 4 * __popc(wm);
  float* stage = C.spill + (size_t)CM(max_contact_pairs) * GSR_CONTACT_WORDS;  // [lane][4][8]
  __threadfence_block();
  C.sync();
  const long long nb1 = NARROW_CLOCK();
  if (lane == 0) { NARROW_MAX(12, nb1 - nb0); NARROW_ADD(14, nb1 - nb0); }
  for (int base = 0; base < n_task; base += G) {
    const int tt = base + lane;
    const bool active = tt < n_task;
    const int src = active ? (int)__fns(wm, 0, (tt >> 2) + 1) : 0;
    const int i_det = (tt & 3) + 1;
    PairDet T;
#define SHF(v) __shfl_sync(C.gmask, v, src, G)
#define SHF3(a) f3(SHF(a.x), SHF(a.y), SHF(a.z))
#define SHF4(a) f4(SHF(a.w), SHF(a.x), SHF(a.y), SHF(a.z))
    T.ga = SHF(D.ga); T.gb = SHF(D.gb); T.ta = SHF(D.ta); T.tb = SHF(D.tb); T.p = SHF(D.p);
    T.pa0 = SHF3(D.pa0); T.pb0 = SHF3(D.pb0); T.qa0 = SHF4(D.qa0); T.qb0 = SHF4(D.qb0);
    T.normal_0 = SHF3(D.normal_0); T.cpos_0 = SHF3(D.cpos_0); T.axis_0 = SHF3(D.axis_0); T.axis_1 = SHF3(D.axis_1);
    T.tolerance = SHF(D.tolerance);
#undef SHF4
#undef SHF3
#undef SHF
    if (active) {
      bool is_col;
      f3 normal, cpos;
      float pen = 0.f;
      dev_pair_perturbed<G, (MPR ? (BB ? 2 : 1) : 0)>(C, T, i_det, is_col, normal, pen, cpos);
      float* r = stage + ((size_t)src * 4 + (i_det - 1)) * 8;
      r[0] = is_col ? 1.f : 0.f; r[1] = normal.x; r[2] = normal.y; r[3] = normal.z; r[4] = cpos.x; r[5] = cpos.y; r[6] = cpos.z; r[7] = pen;
    }
  }
  __threadfence_block();
  C.sync();
  if (lane == 0) { NARROW_MAX(13, NARROW_CLOCK() - nb1); NARROW_ADD(15, NARROW_CLOCK() - nb1); }
  if (want) {
    for (int i_det = 1; i_det < 5; i_det++) {
      const float* r = stage + ((size_t)lane * 4 + (i_det - 1)) * 8;
      if (r[0] == 0.f) continue;
      const f3 normal(r[1], r[2], r[3]), cpos(r[4], r[5], r[6]);
      const float pen = r[7];
      bool repeated = false;
      for (int c = 0; c < out.n; c++)
        if (norm(cpos - out.pos[c]) < D.tolerance) repeated = true;
      if (!repeated && pen > -D.tolerance) out.add(normal, cpos, fmaxf(pen, 0.f));
    }
  }
  // exclusive prefix of the per-lane counts (lane order = pair order)
  int incl = out.n;
#pragma unroll
  for (int o = 1; o < G; o <<= 1) {
    const int v = __shfl_up_sync(C.gmask, incl, o, G);
    if (lane >= o) incl += v;
  }
  const int total = __shfl_sync(C.gmask, incl, G - 1, G);
  const int base = n_staged + incl - out.n;
  const int cap = CM(max_contact_pairs);
  if (out.n > 0) {
    const float mu = fmaxf(fmaxf(GM(g_friction[ga]) * C.rec[c_h.friction_ratio + ga], GM(g_friction[gb]) * C.rec[c_h.friction_ratio + gb]), 1e-2f);
    const int la = GM(g_link[ga]), lb = GM(g_link[gb]);
    for (int k = 0; k < 5; k++) {
      if (k < out.n && base + k < cap) {
        float* r = C.spill + (size_t)(base + k) * GSR_CONTACT_WORDS;
        int* ri = reinterpret_cast<int*>(r);
        ri[CT_GA] = ga; ri[CT_GB] = gb; ri[CT_LA] = la; ri[CT_LB] = lb;
        r[CT_PEN] = out.pen[k]; r[CT_MU] = mu;
        r[CT_POS] = out.pos[k].x; r[CT_POS + 1] = out.pos[k].y; r[CT_POS + 2] = out.pos[k].z;
        r[CT_NRM] = out.nrm[k].x; r[CT_NRM + 1] = out.nrm[k].y; r[CT_NRM + 2] = out.nrm[k].z;
        ri[CT_FRC] = (plane_box << 24) | (p << 3) | k;  // canonical detection order (convex pairs first, pair, index)
        r[CT_FRC + 1] = 0.f; r[CT_FRC + 2] = 0.f;
        r[CT_KEY] = out.pos[0].x;                        // CT:455-470: sort key = x of the pair's first contact
      }
    }
  }
  int n_new = n_staged + total;
  if (n_new > cap) {
    if (lane == 0) atomicOr(&C.SI()[c_h.errno_], ERR_CONTACT_OVERFLOW);
    n_new = cap;
  }
  return n_new;
}

template <int G, bool MPR, bool BB>
__device__ void phase_collide(const Ctx<G>& C) {
  const int lane = C.lane;
  const int nP = CM(nP), nG = CM(nG);
  phase_geoms(C);
  int* SI = C.SI();
  const int n_prev = SI[c_h.n_contacts];
  int n_staged = 0;
  if (nP > 0) {
    const float* amin = WK(aabb);
    const float* amax = WK(aabb) + 3 * nG;
    int* queue = reinterpret_cast<int*>(WK(queue));
    int* cvalid = reinterpret_cast<int*>(C.cache) + c_h.cache_valid;
    int qn = 0, n_broad = 0;
    for (int base = 0; base < nP; base += G) {
      const int p = base + lane;
      bool cand = false;
      if (p < nP) {
        const int ga = __ldg(CM(pairs) + 2 * p), gb = __ldg(CM(pairs) + 2 * p + 1);
        const f3 a0 = ld3(amin, ga), a1 = ld3(amax, ga), b0 = ld3(amin, gb), b1 = ld3(amax, gb);
        const bool sep = a1.x <= b0.x || a0.x >= b1.x || a1.y <= b0.y || a0.y >= b1.y || a1.z <= b0.z || a0.z >= b1.z;
        cand = !sep;
        if (sep && GM(g_type[ga]) != GEOM_PLANE && GM(g_type[gb]) != GEOM_PLANE) {  // BP:120-133: separated pairs lose their warm start
          if ((cvalid[p >> 5] >> (p & 31)) & 1) atomicAnd(&cvalid[p >> 5], ~(1 << (p & 31)));
        }
      }
      const unsigned bal = C.ballot(cand);
      const int before = __popc(bal & ((1u << lane) - 1u));
      const int cnt = __popc(bal);
      if (cand) {
        if (n_broad + before < CM(max_broad_pairs)) queue[qn + before] = p;
      }
      int accepted = cnt;
      if (n_broad + cnt > CM(max_broad_pairs)) {
        accepted = max(0, CM(max_broad_pairs) - n_broad);
        if (lane == 0) atomicOr(&SI[c_h.errno_], ERR_BROAD_OVERFLOW);
      }
      n_broad += accepted;
      qn += accepted;
      C.sync();
      if (qn >= G) {
        n_staged = narrow_batch<G, MPR, BB>(C, queue, G, n_staged);
        C.sync();
        const int rest = qn - G;  // move the tail (< G entries) to the front
        int t = 0;
        if (lane < rest) t = queue[G + lane];
        C.sync();
        if (lane < rest) queue[lane] = t;
        qn = rest;
        C.sync();
      }
    }
    if (qn > 0) n_staged = narrow_batch<G, MPR, BB>(C, queue, qn, n_staged);
    if (lane == 0) SI[c_h.n_broad] = n_broad;
  }
  __threadfence_block();
  C.sync();
  // ---- canonical order (CT:439-538) by ranking, then the final records
  const int n = n_staged;
  const bool sorted = CM(has_cc) && CM(sort_contacts);
  float* ct_s = WK(contacts);
  const int NCC = c_s.NCC;
  for (int c = lane; c < n; c += G) {
    const float* r = C.spill + (size_t)c * GSR_CONTACT_WORDS;
    const float key = sorted ? r[CT_KEY] : 0.f;
    const int ord = reinterpret_cast<const int*>(r)[CT_FRC];
    int rank = 0;
    for (int o = 0; o < n; o++) {
      const float* ro = C.spill + (size_t)o * GSR_CONTACT_WORDS;
      const float ko = sorted ? ro[CT_KEY] : 0.f;
      const int oo = reinterpret_cast<const int*>(ro)[CT_FRC];
      rank += (ko < key || (ko == key && oo < ord)) ? 1 : 0;
    }
    float v[GSR_CONTACT_WORDS];
#pragma unroll
    for (int k = 0; k < GSR_CONTACT_WORDS; k++) v[k] = r[k];
    v[CT_FRC] = 0.f;
    float* dst = C.contacts + (size_t)rank * GSR_CONTACT_WORDS;
#pragma unroll
    for (int k = 0; k < GSR_CONTACT_WORDS; k += 4) *reinterpret_cast<float4*>(dst + k) = make_float4(v[k], v[k + 1], v[k + 2], v[k + 3]);
    if (rank < NCC) {
#pragma unroll
      for (int k = 0; k < GSR_CONTACT_WORDS; k++) ct_s[rank * GSR_CONTACT_WORDS + k] = v[k];
    }
  }
  // BP:74-139: slots that held a contact last step and do not any more read as cleared (-1 / 0)
  for (int c = n + lane; c < n_prev; c += G) {
    float* dst = C.contacts + (size_t)c * GSR_CONTACT_WORDS;
    const float m1 = __int_as_float(-1);
    *reinterpret_cast<float4*>(dst) = make_float4(m1, m1, m1, m1);
    *reinterpret_cast<float4*>(dst + 4) = make_float4(0.f, 0.f, 0.f, 0.f);
    *reinterpret_cast<float4*>(dst + 8) = make_float4(0.f, 0.f, 0.f, 0.f);
    *reinterpret_cast<float4*>(dst + 12) = make_float4(0.f, 0.f, 0.f, 0.f);
  }
  if (lane == 0) SI[c_h.n_contacts] = n;
  C.sync();
}
