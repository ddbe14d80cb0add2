// gsr_gjk.cuh -- the reference's GJK + EPA contact query for one convex pair (scope row a17): what the narrow phase runs
// instead of keeping MPR's answer when no contact direction is cached or the overlap is large (narrowphase.py:726-733, 756-850).
//
//   gjk_contact      gjk.py:170-446   func_gjk_contact, non-MuJoCo branch: safe GJK, then safe EPA from its tetrahedron; result =
//                                     witness points w1 / w2, pos = (w1 + w2) / 2, normal = (w2 - w1) / |w2 - w1|, penetration
//   safe_gjk         gjk.py:1209-1425 func_safe_gjk (+ :1428-1540 vertex validity, :1545-1645 brute-force vertex search,
//                                     :1701-1730 triangle info, :1733-1838 perturbed support, :1841-1900 support counting)
//   safe_epa         epa.py:970-1183  func_safe_epa (+ :1251-1301 init, :1304-1386 attach face, :1389-1427 plane normal,
//                                     :274-433 horizon / delete face, :1184-1249 witness; gjk_utils.py:45-100, 183-235)
//   constants        gjk.py:53-110    50 + 50 iterations, FLOAT_MIN 1e-15, tolerance 1e-6 (EPS for two discrete geoms),
//                                     simplex degeneracy 1e-5, reprojection error 1e-4, <= 300 polytope faces
//
// One thread runs one query.  The polytope lives in the thread's local memory in a compact form (16-bit indices, 32-byte faces,
// Minkowski points recomputed as obj1 - obj2: the same subtraction the reference stores): ~14 KB against the reference's
// global gjk_state arrays (array_class.py:870-1116).  `Geo` supplies the pair: support points with vertex identities, support
// multiplicity of boxes, the discrete-vertex enumeration of boxes / meshes, and EPS.
#pragma once
#include "gsr_math.cuh"

#define GSR_GJK_MAXV 56    /* 4 simplex vertices + one per EPA iteration (50) */
#define GSR_GJK_MAXF 300   /* gjk.py:62 polytope_max_faces = 6 * epa_max_iterations */
#define GSR_GJK_STACK 192

struct GjkVert {
  f3 o1, o2;
  int id1, id2;
  __host__ __device__ __forceinline__ f3 mink() const { return o1 - o2; }
};
struct GjkFace {
  unsigned short v[3], adj[3];
  short map_idx;
  f3 n;
  float d2;
};

#pragma nv_exec_check_disable
template <class Geo>
struct GjkSolver {
  const Geo& geo;
  GjkVert sv[4];
  int nverts;
  int last_searched;
  int n_search = 0, n_iter = 0;  // developer counters (GSR_NARROW_TRACE): candidates walked by search_valid_vertex, GJK iterations
  __host__ __device__ GjkSolver(const Geo& g) : geo(g), nverts(0), last_searched(0) {}

  __host__ __device__ bool duplicate(int id1, int id2) const {  // gjk.py:1451-1470
    for (int i = 0; i < nverts; i++) {
      if (id1 == -1 || sv[i].id1 != id1) continue;
      if (id2 == -1 || sv[i].id2 != id2) continue;
      return true;
    }
    return false;
  }
  __host__ __device__ bool degenerate(f3 mk) const {  // gjk.py:1473-1540
    const float deg2 = 1e-5f * 1e-5f;
    for (int i = 0; i < nverts; i++) { const f3 d = sv[i].mink() - mk; if (dot(d, d) < deg2) return true; }
    if (nverts == 2) {
      const f3 e1 = sv[1].mink() - sv[0].mink(), e2 = mk - sv[0].mink(), n = cross(e1, e2);
      return dot(n, n) < deg2 * dot(e1, e1) * dot(e2, e2);
    }
    if (nverts == 3) {
      const f3 e1 = normalized(sv[1].mink() - sv[0].mink()), e2 = normalized(sv[2].mink() - sv[0].mink()), n = cross(e1, e2), diff = mk - sv[0].mink();
      const float nd = dot(n, diff);
      return nd * nd < deg2 * dot(n, n) * dot(diff, diff);
    }
    return false;
  }
  __host__ __device__ bool valid(const GjkVert& r) const { return !duplicate(r.id1, r.id2) && !degenerate(r.mink()); }

  __host__ __device__ GjkVert safe_support(f3 dir) const {  // gjk.py:1733-1838
    const float EPS = geo.eps();
    GjkVert r;
    r.id1 = r.id2 = -1;
    for (int i = 0; i < 9; i++) {
      f3 nd = dir;
      if (i > 0) {
        const int j = i - 1;
        nd.x += -(1.0f - 2.0f * (float)(j & 1)) * EPS;
        nd.y += -(1.0f - 2.0f * (float)(j & 2)) * EPS;
        nd.z += -(1.0f - 2.0f * (float)(j & 4)) * EPS;
      }
      nd = nd * (2.0f - dot(nd, dir));
      const int num = geo.count_support(nd);
      if (i > 0 && num > 1) continue;
      geo.support(nd, 0, r.o1, r.id1);
      geo.support(-nd, 1, r.o2, r.id2);
      if (i == 0) { if (num > 1) continue; else break; }
      if (i == 8) break;
      if (valid(r)) break;
    }
    return r;
  }
  __host__ __device__ bool search_valid_vertex(GjkVert& r) {  // gjk.py:1545-1645
    r.o1 = f3(); r.o2 = f3(); r.id1 = r.id2 = -1;
    if (geo.discrete()) {
      // Exhaustive walk over the vertex pairs (i of A, j of B) from where the last search stopped, first valid pair wins.
      // Minkowski differences of boxes / hulls put long runs of candidates on a common plane with the simplex, so the walk
      // is the hot loop of the exact-penetration path (Go2 self-contacts: two thirds of its instructions before this
      // rewrite): the indices advance incrementally (no division), A's vertex is fetched once per row, and the parts of the
      // validity test (`valid`: duplicate + degenerate) that do not depend on the candidate are evaluated once.
      const int n1 = geo.n_verts(1), num = geo.n_verts(0) * n1;
      if (num <= 0) return false;
      const float deg2 = 1e-5f * 1e-5f;
      f3 m0, m1, m2, e1, n;  // (explicit names, not an array: the simplex stays in registers)
      float e1e1 = 0.f, nn = 0.f;
      if (nverts > 0) m0 = sv[0].mink();
      if (nverts > 1) m1 = sv[1].mink();
      if (nverts > 2) m2 = sv[2].mink();
      if (nverts == 2) { e1 = m1 - m0; e1e1 = dot(e1, e1); }
      if (nverts == 3) { n = cross(normalized(m1 - m0), normalized(m2 - m0)); nn = dot(n, n); }
      int mm = last_searched % num, i = mm / n1, j = mm - i * n1;
      geo.vertex(0, i, r.o1, r.id1);
      for (int k = 0; k < num; k++) {
        n_search++;
        geo.vertex(1, j, r.o2, r.id2);
        bool ok = !duplicate(r.id1, r.id2);
        if (ok) {
          const f3 mk = r.mink();
          if (nverts > 0) { const f3 d = m0 - mk; if (dot(d, d) < deg2) ok = false; }
          if (nverts > 1) { const f3 d = m1 - mk; if (dot(d, d) < deg2) ok = false; }
          if (nverts > 2) { const f3 d = m2 - mk; if (dot(d, d) < deg2) ok = false; }
          if (ok && nverts == 2) { const f3 e2 = mk - m0, c = cross(e1, e2); ok = !(dot(c, c) < deg2 * e1e1 * dot(e2, e2)); }
          if (ok && nverts == 3) { const f3 diff = mk - m0; const float nd = dot(n, diff); ok = !(nd * nd < deg2 * nn * dot(diff, diff)); }
        }
        mm++; j++;
        if (ok) { last_searched = mm == num ? 0 : mm; return true; }
        if (j == n1) { j = 0; i++; if (mm == num) { mm = 0; i = 0; } geo.vertex(0, i, r.o1, r.id1); }
      }
      return false;
    }
    if (nverts == 3) {
      const f3 dir = normalized(cross(sv[2].mink() - sv[0].mink(), sv[1].mink() - sv[0].mink()));
      for (int i = 0; i < 2; i++) {
        r = safe_support(i == 0 ? dir : -dir);
        if (valid(r)) return true;
      }
    }
    return false;
  }
  __host__ __device__ bool safe_gjk() {  // gjk.py:1209-1425: true = the origin is inside the final tetrahedron
    nverts = 0;
    bool init_ok = true;
    for (int i = 0; i < 4; i++) {
      f3 dir;
      set_comp(dir, 2 - i / 2, 1.0f - 2.0f * (float)(i % 2));
      GjkVert r = safe_support(dir);
      if (!valid(r)) {
        if (!search_valid_vertex(r)) { init_ok = false; break; }
      }
      sv[i] = r;
      nverts++;
    }
    if (!init_ok) return false;
    int si[4] = {0, 1, 2, 3};
    for (int it = 0; it < 50; it++) {
      n_iter++;
      f3 min_normal;
      float min_sd = 0.f;
      int min_i = 0;
      for (int j = 0; j < 4; j++) {
        int s0 = si[2], s1 = si[1], s2 = si[3], ap = si[0];
        if (j == 1) { s0 = si[0]; s1 = si[2]; s2 = si[3]; ap = si[1]; }
        else if (j == 2) { s0 = si[1]; s1 = si[0]; s2 = si[3]; ap = si[2]; }
        else if (j == 3) { s0 = si[0]; s1 = si[1]; s2 = si[2]; ap = si[3]; }
        const f3 v1 = sv[s0].mink(), v2 = sv[s1].mink(), v3 = sv[s2].mink(), apex = sv[ap].mink();  // gjk.py:1701-1730
        f3 n = normalized(cross(v3 - v1, v2 - v1));
        if (dot(n, apex - v1) > 0.f) n = -n;
        const float sd = dot(n, v1);
        if (j == 0 || sd < min_sd) { min_sd = sd; min_normal = n; min_i = j; }
      }
      const int min_si = si[min_i];
      if (min_sd >= 0.f) return true;
      nverts = 3;
      if (min_si != 3) sv[min_si] = sv[3];
      const GjkVert r = safe_support(min_normal);
      if (duplicate(r.id1, r.id2)) return false;
      if (degenerate(r.mink())) return false;
      if (dot(r.mink(), min_normal) < 0.f) return false;
      sv[3] = r;
      nverts = 4;
    }
    return false;
  }
};

__host__ __device__ __forceinline__ bool gjk_plane_normal(f3 v1, f3 v2, f3 v3, f3& normal) {  // epa.py:1389-1427
  const f3 d21 = v2 - v1, d31 = v3 - v1, d32 = v3 - v2;
  for (int i = 0; i < 3; i++) {
    const f3 n = i == 0 ? cross(d32, d21) : (i == 1 ? cross(d21, d31) : cross(d31, d32));
    const float nn = norm(n);
    if (nn == 0.f) return false;
    if (nn > 1e-15f) { normal = normalized(n); return true; }
  }
  return false;
}
__host__ __device__ __forceinline__ f3 gjk_affine_coords(f3 point, f3 t1, f3 t2, f3 t3) {  // gjk_utils.py:45-100
  float ms[3];
  for (int i = 0; i < 3; i++) {
    int i1 = (i + 1) % 3, i2 = (i + 2) % 3;
    if (i == 1) { const int t = i1; i1 = i2; i2 = t; }
    ms[i] = comp(t2, i1) * comp(t3, i2) - comp(t2, i2) * comp(t3, i1) - comp(t1, i1) * comp(t3, i2) + comp(t1, i2) * comp(t3, i1) +
            comp(t1, i1) * comp(t2, i2) - comp(t1, i2) * comp(t2, i1);
  }
  float m_max = 0.f;
  int ix = 0, iy = 0;
  for (int i = 0; i < 3; i++) {
    if (fabsf(ms[i]) >= fabsf(ms[(i + 1) % 3]) && fabsf(ms[i]) >= fabsf(ms[(i + 2) % 3])) {
      m_max = ms[i]; ix = (i + 1) % 3; iy = (i + 2) % 3;
      if (i == 1) { const int t = ix; ix = iy; iy = t; }
      break;
    }
  }
  f3 cs;
  for (int i = 0; i < 3; i++) {
    const f3 a = i == 0 ? t2 : (i == 1 ? t3 : t1), b = i == 0 ? t3 : (i == 1 ? t1 : t2);
    set_comp(cs, i, comp(point, ix) * comp(a, iy) + comp(point, iy) * comp(b, ix) + comp(a, ix) * comp(b, iy) - comp(point, ix) * comp(b, iy) -
                        comp(point, iy) * comp(a, ix) - comp(b, ix) * comp(a, iy));
  }
  return cs / m_max;
}
__host__ __device__ __forceinline__ bool gjk_project_origin(f3 v1, f3 v2, f3 v3, f3& point) {  // gjk_utils.py:183-235
  const f3 d21 = v2 - v1, d31 = v3 - v1, d32 = v3 - v2;
  point = f3();
  for (int i = 0; i < 3; i++) {
    f3 n, v;
    if (i == 0) { n = cross(d32, d21); v = v2; } else if (i == 1) { n = cross(d21, d31); v = v1; } else { n = cross(d31, d32); v = v3; }
    const float nv = dot(n, v), nn = dot(n, n);
    if (nn == 0.f) return false;
    if (nn > 1e-15f) { point = n * (nv / nn); return true; }
    if (i == 2) { if (nn < 1e-15f) return false; point = n * (nv / nn); return true; }
  }
  return true;
}

#if defined(GSR_NARROW_TRACE) && defined(__CUDACC__)  // developer build (tools/go2_tail.py): where does one exact-penetration query spend its time?
__device__ unsigned long long g_narrow[16];
#endif
#if defined(GSR_NARROW_TRACE) && defined(__CUDA_ARCH__)
#define NARROW_CLOCK() clock64()
#define NARROW_ADD(k, v) atomicAdd(&g_narrow[k], (unsigned long long)(v))
#define NARROW_MAX(k, v) atomicMax(&g_narrow[k], (unsigned long long)(v))
#else
#define NARROW_CLOCK() 0ll
#define NARROW_ADD(k, v)
#define NARROW_MAX(k, v)
#endif

#pragma nv_exec_check_disable
template <class Geo>
__host__ __device__ __noinline__ void gjk_contact(const Geo& geo, bool& is_col, f3& normal, float& pen, f3& pos) {
  is_col = false; pen = 0.f;
  GjkSolver<Geo> gs(geo);
  const long long nt0 = NARROW_CLOCK();
  const bool gjk_inside = gs.safe_gjk();
  const long long nt1 = NARROW_CLOCK();
  NARROW_ADD(9, 1); NARROW_ADD(10, nt1 - nt0); NARROW_ADD(11, gs.n_search);  // exact-penetration queries, GJK cycles, candidates walked
  (void)nt0; (void)nt1;
  if (!gjk_inside) return;
  // ---- safe EPA (epa.py:970-1183) from the GJK tetrahedron (epa.py:1251-1301)
  const float EPS = geo.eps(), FMAX = 1e15f;
  GjkVert pv[GSR_GJK_MAXV];
  GjkFace pf[GSR_GJK_MAXF];
  unsigned short fmap[GSR_GJK_MAXF], hz_face[GSR_GJK_MAXF], hz_edge[GSR_GJK_MAXF], st_face[GSR_GJK_STACK];
  unsigned char st_edge[GSR_GJK_STACK];
  int npv = 0, nfaces = 0, nfmap = 0, nhz = 0;
  auto attach = [&](int v1, int v2, int v3, int a1, int a2, int a3) -> bool {  // epa.py:1304-1386
    GjkFace& f = pf[nfaces];
    f.v[0] = (unsigned short)v1; f.v[1] = (unsigned short)v2; f.v[2] = (unsigned short)v3;
    f.adj[0] = (unsigned short)a1; f.adj[1] = (unsigned short)a2; f.adj[2] = (unsigned short)a3;
    nfaces++;
    const f3 m1 = pv[v1].mink(), m2 = pv[v2].mink(), m3 = pv[v3].mink();
    f3 nrm;
    if (!gjk_plane_normal(m3, m2, m1, nrm)) return false;
    const f3 center = (m1 + m2 + m3) / 3.0f;
    float max_orient = -dot(nrm, center), max_abs = fabsf(max_orient);
    for (int iv = 0; iv < npv; iv++) {
      if (iv != v1 && iv != v2 && iv != v3) {
        const float orient = dot(nrm, pv[iv].mink() - center);
        if (fabsf(orient) > max_abs) { max_abs = fabsf(orient); max_orient = orient; }
      }
    }
    if (max_orient > 0.f) nrm = -nrm;
    f.n = nrm;
    const float d1 = dot(nrm, m1), d2 = dot(nrm, m2), d3 = dot(nrm, m3);
    float mn = FMAX;
    if (d1 * d1 < mn) mn = d1 * d1;
    if (d2 * d2 < mn) mn = d2 * d2;
    if (d3 * d3 < mn) mn = d3 * d3;
    f.d2 = mn;
    f.map_idx = -1;
    return true;
  };
  for (int i = 0; i < 4; i++) pv[npv++] = gs.sv[i];
  {
    const int fv[4][3] = {{0, 1, 2}, {0, 3, 1}, {0, 2, 3}, {3, 2, 1}}, fa[4][3] = {{1, 3, 2}, {2, 3, 0}, {0, 3, 1}, {2, 0, 1}};
    for (int i = 0; i < 4; i++) attach(fv[i][0], fv[i][1], fv[i][2], fa[i][0], fa[i][1], fa[i][2]);
    for (int i = 0; i < 4; i++) { fmap[i] = (unsigned short)i; pf[i].map_idx = (short)i; }
    nfmap = 4;
  }
  float upper = FMAX, upper2 = FMAX * FMAX, lower = 0.f, tolerance = 1e-6f;
  int nearest = -1, prev_nearest = -1;
  const bool discrete = geo.discrete();
  if (discrete) tolerance = EPS;
  for (int k = 0; k < 50; k++) {
    prev_nearest = nearest;
    float lower2 = FMAX * FMAX;
    for (int i = 0; i < nfmap; i++) {
      const int i_f = fmap[i];
      if (pf[i_f].d2 < lower2) { lower2 = pf[i_f].d2; nearest = i_f; }
    }
    if (lower2 > upper2 || nearest == -1) { nearest = prev_nearest; break; }
    lower = sqrtf(lower2);
    const f3 dir = pf[nearest].n;
    const int wi = npv;  // epa.py:810-883 func_epa_support(dir, 1.0): the plain support pair joins the polytope
    geo.support(dir, 0, pv[wi].o1, pv[wi].id1);
    geo.support(-dir, 1, pv[wi].o2, pv[wi].id2);
    npv++;
    const f3 w = pv[wi].mink();
    const float upper_k = dot(w, dir);
    if (upper_k < upper) { upper = upper_k; upper2 = upper * upper; }
    if ((upper - lower) < tolerance) break;
    if (discrete) {
      bool repeated = false;
      for (int i = 0; i < npv; i++) {
        if (i == wi) continue;
        if (pv[i].id1 == pv[wi].id1 && pv[i].id2 == pv[wi].id2) { repeated = true; break; }
      }
      if (repeated) break;
    }
    {  // epa.py:274-341 func_epa_horizon: depth-first walk over the faces visible from w; horizon edges come out as a loop
      st_face[0] = (unsigned short)nearest; st_edge[0] = 0;
      int top = 1;
      bool is_first = true, overflow = false;
      while (top > 0) {
        const int i_f = st_face[top - 1], i_e = st_edge[top - 1];
        const f3 v = pv[pf[i_f].v[0]].mink();
        top--;
        const bool deleted = pf[i_f].map_idx == -2;
        if (!is_first && deleted) continue;
        const bool visible = dot(pf[i_f].n, w - v) > 1e-15f;
        if (visible || is_first) {
          const int mi = pf[i_f].map_idx;  // epa.py:384-405 delete face
          if (mi >= 0) {
            const int last = fmap[nfmap - 1];
            fmap[mi] = (unsigned short)last;
            pf[last].map_idx = (short)mi;
            nfmap--;
          }
          pf[i_f].map_idx = -2;
          for (int kk = is_first ? 0 : 1; kk < 3; kk++) {
            const int e2 = (i_e + kk) % 3, adj = pf[i_f].adj[e2];
            if (pf[adj].map_idx != -2) {
              const int start_v = pf[i_f].v[(e2 + 1) % 3];
              const int ae = pf[adj].v[0] == start_v ? 0 : (pf[adj].v[1] == start_v ? 1 : 2);
              if (top >= GSR_GJK_STACK) { overflow = true; break; }
              st_face[top] = (unsigned short)adj; st_edge[top] = (unsigned char)ae; top++;
            }
          }
          if (overflow) break;
        } else {
          hz_edge[nhz] = (unsigned short)i_e; hz_face[nhz] = (unsigned short)i_f; nhz++;
        }
        is_first = false;
      }
      if (overflow) { nearest = -1; break; }
    }
    if (nhz < 3) { nearest = -1; break; }
    const int nf0 = nfaces, nedges = nhz;
    if (nf0 + nedges >= GSR_GJK_MAXF || npv >= GSR_GJK_MAXV) break;
    bool attach_ok = true;
    for (int i = 0; i < nedges; i++) {
      const int i_f0 = nf0 + i, i_f1 = nf0 + (i + 1) % nedges;
      const int hf = hz_face[i], he = hz_edge[i];
      const int hv1 = pf[hf].v[he], hv2 = pf[hf].v[(he + 1) % 3];
      pf[hf].adj[he] = (unsigned short)i_f0;
      const int adj0 = i > 0 ? i_f0 - 1 : nf0 + nedges - 1;
      if (!attach(wi, hv2, hv1, i_f1, hf, adj0)) { attach_ok = false; break; }
      const float d2 = pf[nfaces - 1].d2;
      if (d2 >= lower2 - EPS && d2 <= upper2 + EPS) { fmap[nfmap] = (unsigned short)i_f0; pf[i_f0].map_idx = (short)nfmap; nfmap++; }
    }
    if (!attach_ok) { nearest = -1; break; }
    nhz = 0;
    if (nfmap == 0 || nearest == -1) { nearest = -1; break; }
  }
  if (nearest == -1) return;
  // ---- witness points (epa.py:1184-1249) and the contact (gjk.py:417-446)
  const GjkFace& f = pf[nearest];
  const f3 f1 = pv[f.v[0]].mink(), f2 = pv[f.v[1]].mink(), f3_ = pv[f.v[2]].mink();
  f3 proj;
  gjk_project_origin(f1, f2, f3_, proj);
  const f3 lam = gjk_affine_coords(proj, f1, f2, f3_);
  const f3 back = f1 * lam.x + f2 * lam.y + f3_ * lam.z;
  const float err = norm(proj - back);
  const float e12 = dot(f1 - f2, f1 - f2), e23 = dot(f2 - f3_, f2 - f3_), e31 = dot(f3_ - f1, f3_ - f1);
  const float inv_len = 1.0f / sqrtf(fmaxf(fmaxf(e12, e23), fmaxf(e31, 1e-30f)));
  if (err * inv_len > 1e-4f && err > 1e-4f) return;
  const f3 w1 = pv[f.v[0]].o1 * lam.x + pv[f.v[1]].o1 * lam.y + pv[f.v[2]].o1 * lam.z;
  const f3 w2 = pv[f.v[0]].o2 * lam.x + pv[f.v[1]].o2 * lam.y + pv[f.v[2]].o2 * lam.z;
  const float distance = -sqrtf(f.d2);
  if (!(distance < 0.f)) return;
  const f3 d = w2 - w1;
  const float len = norm(d);
  if (len < 1e-15f) return;
  is_col = true;
  pen = -distance;
  pos = (w1 + w2) * 0.5f;
  normal = d / len;
}
