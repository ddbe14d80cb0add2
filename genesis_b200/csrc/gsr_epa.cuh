// gsr_epa.cuh -- exact penetration of two overlapping convex geoms by the expanding polytope algorithm (row a17 of the
// scope table: the GJK / EPA fallback the reference takes after MPR when the overlap is large or no contact direction is
// cached, narrowphase.py:726-733; result convention of gjk.py:420-446: witness points w1 on geom A and w2 on geom B of the
// minimum-translation face of the Minkowski difference A - B, pos = (w1 + w2) / 2, normal = (w2 - w1) / |w2 - w1| (pointing
// from B to A), penetration = distance of that face from the origin).
//
// The reference seeds EPA with a tetrahedron found by its "safe" GJK (gjk.py:1209-1425); here the seed is the tetrahedron MPR
// has just built around the origin -- its interior point v0 and the final portal -- so no second intersection search is
// needed: the minimum over the boundary of a convex set does not depend on the seed (van den Bergen, "Proximity queries and
// penetration depth computation on 3D game objects", 2001).  Limits follow gjk.py:53-56, 89-90: 50 iterations, tolerance
// 1e-6; the polytope is bounded by 40 vertices / 128 faces (the reference allows 300 faces for its 5-detection variant).
// One thread runs one query; vertices and faces live in its local memory.  Shared by the lane-group and the lane kernels
// (and compiled for the host by tests/lane_emu); the oracle has its own std::vector implementation of the same algorithm.
#pragma once
#include "gsr_math.cuh"

#define GSR_EPA_MAXV 40
#define GSR_EPA_MAXF 128
#define GSR_EPA_MAXE 48

struct EpaFace {
  int a, b, c;
  f3 n;
  float d;
  int alive;
};

__host__ __device__ __forceinline__ bool epa_make_face(const f3* pv, f3 inside, int a, int b, int c, EpaFace& f) {
  f3 n = cross(pv[b] - pv[a], pv[c] - pv[a]);
  const float len = norm(n);
  if (!(len > 1e-20f)) return false;
  n = n / len;
  if (dot(n, inside - pv[a]) > 0.f) { const int t = b; b = c; c = t; n = -n; }
  f.a = a; f.b = b; f.c = c; f.n = n; f.d = dot(n, pv[a]); f.alive = 1;
  return true;
}

// sup(d, v, v1, v2): support point of A - B in direction d with its witnesses.  sv / sv1 / sv2: the 4 seed vertices.
#pragma nv_exec_check_disable
template <class Sup>
__host__ __device__ __noinline__ bool epa_penetration(const f3* sv, const f3* sv1, const f3* sv2, const Sup& sup, f3& normal, float& pen, f3& pos) {
  const float TOL = 1e-6f;
  f3 pv[GSR_EPA_MAXV], p1[GSR_EPA_MAXV], p2[GSR_EPA_MAXV];
  EpaFace faces[GSR_EPA_MAXF];
  int ea[GSR_EPA_MAXE], eb[GSR_EPA_MAXE];
  int nv = 4, nf = 0;
  for (int i = 0; i < 4; i++) { pv[i] = sv[i]; p1[i] = sv1[i]; p2[i] = sv2[i]; }
  const f3 inside = (pv[0] + pv[1] + pv[2] + pv[3]) * 0.25f;
  const int tri[4][3] = {{0, 1, 2}, {0, 2, 3}, {0, 3, 1}, {1, 3, 2}};
  for (int i = 0; i < 4; i++) {
    if (!epa_make_face(pv, inside, tri[i][0], tri[i][1], tri[i][2], faces[nf])) return false;
    if (faces[nf].d < -TOL) return false;  // the origin is not inside the seed
    nf++;
  }
  int best = -1;
  for (int it = 0; it < 50; it++) {
    best = -1;
    for (int i = 0; i < nf; i++)
      if (faces[i].alive && (best < 0 || faces[i].d < faces[best].d)) best = i;
    if (best < 0) return false;
    const f3 bn = faces[best].n;
    const float bd = faces[best].d;
    f3 w, w1, w2;
    sup(bn, w, w1, w2);
    if (dot(w, bn) - bd < TOL) break;
    bool dup = false;
    for (int i = 0; i < nv; i++)
      if (norm_sqr(pv[i] - w) < 1e-16f) dup = true;
    if (dup || nv >= GSR_EPA_MAXV) break;
    const int iw = nv;
    pv[nv] = w; p1[nv] = w1; p2[nv] = w2; nv++;
    // faces seen from w go away; the edges used by exactly one removed face are the horizon
    int ne = 0, n_alive = 0;
    bool overflow = false;
    for (int i = 0; i < nf; i++) {
      EpaFace& f = faces[i];
      if (!f.alive) continue;
      if (!(dot(f.n, w) - f.d > 1e-12f)) { n_alive++; continue; }
      f.alive = 0;
      const int e0[3] = {f.a, f.b, f.c}, e1[3] = {f.b, f.c, f.a};
      for (int k = 0; k < 3; k++) {
        bool found = false;
        for (int q = 0; q < ne; q++)
          if (ea[q] == e1[k] && eb[q] == e0[k]) { ea[q] = ea[ne - 1]; eb[q] = eb[ne - 1]; ne--; found = true; break; }
        if (!found) {
          if (ne >= GSR_EPA_MAXE) { overflow = true; break; }
          ea[ne] = e0[k]; eb[ne] = e1[k]; ne++;
        }
      }
      if (overflow) break;
    }
    if (overflow || ne < 3 || nf + ne > GSR_EPA_MAXF) { faces[best].alive = 1; break; }
    bool ok = true;
    for (int q = 0; q < ne; q++) {
      if (!epa_make_face(pv, inside, ea[q], eb[q], iw, faces[nf])) { ok = false; break; }
      nf++;
    }
    if (!ok) return false;
  }
  if (best < 0 || !faces[best].alive) {  // the loop ended right after a removal: take the nearest surviving face
    best = -1;
    for (int i = 0; i < nf; i++)
      if (faces[i].alive && (best < 0 || faces[i].d < faces[best].d)) best = i;
    if (best < 0) return false;
  }
  const EpaFace& f = faces[best];
  const f3 p = f.n * f.d, e0 = pv[f.b] - pv[f.a], e1 = pv[f.c] - pv[f.a], e2 = p - pv[f.a];
  const float d00 = dot(e0, e0), d01 = dot(e0, e1), d11 = dot(e1, e1), d20 = dot(e2, e0), d21 = dot(e2, e1);
  const float den = d00 * d11 - d01 * d01;
  if (!(fabsf(den) > 1e-30f)) return false;
  const float lb = (d11 * d20 - d01 * d21) / den, lc = (d00 * d21 - d01 * d20) / den, la = 1.0f - lb - lc;
  const f3 w1 = p1[f.a] * la + p1[f.b] * lb + p1[f.c] * lc, w2 = p2[f.a] * la + p2[f.b] * lb + p2[f.c] * lc;
  const f3 dlt = w2 - w1;
  const float len = norm(dlt);
  normal = len > 1e-10f ? dlt / len : -f.n;
  pen = f.d;
  pos = (w1 + w2) * 0.5f;
  return true;
}
