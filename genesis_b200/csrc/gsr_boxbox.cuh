// gsr_boxbox.cuh -- box-box contacts for RigidOptions.box_box_detection: the separating-axis search over the 15 axes and the
// face / edge clipping of MuJoCo's engine_collision_box.c as the reference runs it
// (genesis/engine/solvers/rigid/collider/box_contact.py:100-785 func_box_box_contact; dispatch narrowphase.py:2028-2032, 2175-2189).
// One lane handles one pair, like the other analytic pairs; its scratch (<= BB_MAX candidate points, four clip lines) lives in
// the lane's local memory -- the routine is only compiled into the k_step<G, true, true> variants, which the host launches
// when the option is on and the valid-pair table holds a box-box pair.
//
// Reference behaviour kept: face contacts report HALF the clipped depth as penetration (the reference reads z after halving
// it, BOX:417-419 / 433), edge contacts the full depth; candidates are reported in the reference's order, the first
// n_contacts_per_pair that are not duplicates (all |dx| < EPS) are kept.
#pragma once
#include "gsr_math.cuh"

#define BB_MAX 16

struct BBFrame {
  int i0, i1, i2;
  float f0, f1, f2;
};
// the signed axis permutation that turns face `f` (0..5) into the +z face (BOX:247-289, 499-527) and its matrix rotmore
__host__ __device__ __forceinline__ BBFrame bb_face_frame(int f, mat3& rm) {
  for (int a = 0; a < 3; a++) for (int b = 0; b < 3; b++) rm.m[a][b] = 0.f;
  BBFrame F{0, 1, 2, 1.f, 1.f, 1.f};
  switch (f) {
    case 0: rm.m[0][2] = -1.f; rm.m[1][1] = 1.f; rm.m[2][0] = 1.f; F.i0 = 2; F.f0 = -1.f; F.i2 = 0; break;
    case 1: rm.m[0][0] = 1.f; rm.m[1][2] = -1.f; rm.m[2][1] = 1.f; F.i1 = 2; F.f1 = -1.f; F.i2 = 1; break;
    case 2: rm.m[0][0] = 1.f; rm.m[1][1] = 1.f; rm.m[2][2] = 1.f; break;
    case 3: rm.m[0][2] = 1.f; rm.m[1][1] = 1.f; rm.m[2][0] = -1.f; F.i0 = 2; F.i2 = 0; F.f2 = -1.f; break;
    case 4: rm.m[0][0] = 1.f; rm.m[1][2] = 1.f; rm.m[2][1] = -1.f; F.i1 = 2; F.i2 = 1; F.f2 = -1.f; break;
    default: rm.m[0][0] = -1.f; rm.m[1][1] = 1.f; rm.m[2][2] = -1.f; F.f0 = -1.f; F.f2 = -1.f; break;
  }
  return F;
}
__host__ __device__ __forceinline__ f3 bb_perm(f3 v, const BBFrame& F) { return f3(comp(v, F.i0) * F.f0, comp(v, F.i1) * F.f1, comp(v, F.i2) * F.f2); }
__host__ __device__ __forceinline__ mat3 bb_perm_rows(const mat3& a, const BBFrame& F) {
  mat3 r;
  for (int c = 0; c < 3; c++) { r.m[0][c] = a.m[F.i0][c] * F.f0; r.m[1][c] = a.m[F.i1][c] * F.f1; r.m[2][c] = a.m[F.i2][c] * F.f2; }
  return r;
}
__host__ __device__ __forceinline__ f3 bb_row(const mat3& a, int i) { return f3(a.m[i][0], a.m[i][1], a.m[i][2]); }
__host__ __device__ __forceinline__ f3 bb_mul(const mat3& a, f3 v) { return f3(dot(bb_row(a, 0), v), dot(bb_row(a, 1), v), dot(bb_row(a, 2), v)); }
__host__ __device__ __forceinline__ f3 bb_mul_t(const mat3& a, f3 v) { return bb_row(a, 0) * v.x + bb_row(a, 1) * v.y + bb_row(a, 2) * v.z; }  // a^T v
__host__ __device__ __forceinline__ mat3 bb_matmul(const mat3& a, const mat3& b, bool ta, bool tb) {  // (ta ? a^T : a) (tb ? b^T : b)
  mat3 r;
  for (int i = 0; i < 3; i++)
    for (int j = 0; j < 3; j++) {
      float s = 0.f;
      for (int k = 0; k < 3; k++) s += (ta ? a.m[k][i] : a.m[i][k]) * (tb ? b.m[j][k] : b.m[k][j]);
      r.m[i][j] = s;
    }
  return r;
}
struct BBLine { f3 o, d; };
__host__ __device__ __forceinline__ float bb_lc(const BBLine& L, int k) { return k < 3 ? comp(L.o, k) : comp(L.d, k - 3); }

// Geometry only (host + device: tests/test_boxbox_host.py runs it on the CPU next to the oracle).  `Out` offers n, pos[], add(normal, pos, pen).
template <class Out>
__host__ __device__ void bb_contacts(f3 pos1, const mat3& mat1, f3 size1, f3 pos2, const mat3& mat2, f3 size2, float EPS, int ncpp, Out& out) {
  const f3 pos21 = bb_mul_t(mat1, pos2 - pos1), pos12 = bb_mul_t(mat2, pos1 - pos2);
  const mat3 rot = bb_matmul(mat1, mat2, true, false);  // box 2 axes in box 1's frame; rot^T rows are rott
  // ---- 15 separating axes (BOX:160-243); margin = 0
  float penetration = 0.f;
  for (int i = 0; i < 3; i++) penetration = penetration + comp(size1, i) * 3.f + comp(size2, i) * 3.f;
  int code = -1, cle1 = 0, cle2 = 0;
  bool sep = false, in_ = false;
  for (int i = 0; i < 3; i++) {
    float plen2 = 0.f, plen1 = 0.f;
    for (int k = 0; k < 3; k++) { plen2 += fabsf(rot.m[i][k]) * comp(size2, k); plen1 += fabsf(rot.m[k][i]) * comp(size1, k); }
    const float c1 = -fabsf(comp(pos21, i)) + comp(size1, i) + plen2;
    const float c2 = -fabsf(comp(pos12, i)) + comp(size2, i) + plen1;
    if (c1 < 0.f || c2 < 0.f) sep = true;
    if (c1 < penetration) { penetration = c1; code = i + 3 * (comp(pos21, i) < 0.f ? 1 : 0); }
    if (c2 < penetration) { penetration = c2; code = i + 3 * (comp(pos12, i) < 0.f ? 1 : 0) + 6; }
  }
  f3 clnorm;
  for (int i = 0; i < 3; i++)
    for (int j = 0; j < 3; j++) {
      const float rj0 = rot.m[0][j], rj1 = rot.m[1][j], rj2 = rot.m[2][j];  // row j of rot^T
      f3 ax = i == 0 ? f3(0.f, -rj2, rj1) : (i == 1 ? f3(rj2, 0.f, -rj0) : f3(-rj1, rj0, 0.f));
      const float c1 = norm(ax);
      if (c1 >= EPS) {
        ax = ax / c1;
        const float c2 = dot(pos21, ax);
        float c3 = 0.f;
        for (int k = 0; k < 3; k++) if (k != i) c3 = c3 + comp(size1, k) * fabsf(comp(ax, k));
        for (int k = 0; k < 3; k++) if (k != j) c3 = c3 + comp(size2, k) * fabsf(rot.m[i][3 - k - j]) / c1;
        c3 = c3 - fabsf(c2);
        if (c3 < 0.f) sep = true;
        if (c3 < penetration) {  // the reference's factor (1 - 1e-12) is 1 in fp32
          penetration = c3;
          cle1 = 0; cle2 = 0;
          for (int k = 0; k < 3; k++) {
            if (k != i && ((comp(ax, k) > 0.f) != (c2 < 0.f))) cle1 += 1 << k;
            if (k != j && (((rot.m[i][3 - k - j] > 0.f) != (c2 < 0.f)) != ((k - j + 3) % 3 == 1))) cle2 += 1 << k;
          }
          code = 12 + i * 3 + j;
          clnorm = ax;
          in_ = c2 < 0.f;
        }
      }
    }
  if (sep || code == -1) return;

  f3 pt[BB_MAX];       // candidate points, frame of the clipping face
  float dep[BB_MAX];   // and their signed distances
  BBLine ln[4], lnu[4];
  mat3 rm;
  int n = 0;
  f3 normal_0, origin;
  mat3 rw;  // clipping-face frame -> world
  float hz;
  if (code < 12) {  // ---- a face of one box against the other box (BOX:244-452)
    const int q1 = code % 6, q2 = code / 6;
    const BBFrame F = bb_face_frame(q1, rm);
    mat3 r;
    f3 p, ss, s;
    if (q2) { r = bb_matmul(rm, rot, false, true); p = bb_perm(pos12, F); ss = bb_perm(size2, F); s = size1; }
    else { r = bb_perm_rows(rot, F); p = bb_perm(pos21, F); ss = bb_perm(size1, F); s = size2; }
    ss = f3(fabsf(ss.x), fabsf(ss.y), fabsf(ss.z));
    const float lx = ss.x, ly = ss.y;
    hz = ss.z;
    p.z -= hz;
    int clcorner = 0;
    for (int i = 0; i < 3; i++) if (r.m[2][i] < 0.f) clcorner += 1 << i;
    f3 lp = p;
    for (int i = 0; i < 3; i++) lp = lp + mat_col(r, i) * comp(s, i) * ((clcorner & (1 << i)) ? 1.f : -1.f);
    f3 q[6];  // lowest corner of the other box, its (up to two) nearly horizontal edges, and the corners they span
    int m = 1, k = 0;
    q[0] = lp; q[1] = f3(); q[2] = f3();
    for (int i = 0; i < 3; i++)
      if (fabsf(r.m[2][i]) < 0.5f) { q[m] = mat_col(r, i) * comp(s, i) * ((clcorner & (1 << i)) ? -2.f : 2.f); m++; }
    q[3] = q[0] + q[1]; q[4] = q[0] + q[2]; q[5] = q[3] + q[2];
    if (m > 1) { ln[k].o = q[0]; ln[k].d = q[1]; k++; }
    if (m > 2) {
      ln[k].o = q[0]; ln[k].d = q[2]; k++;
      ln[k].o = q[3]; ln[k].d = q[2]; k++;
      ln[k].o = q[4]; ln[k].d = q[1]; k++;
    }
    for (int i = 0; i < k; i++)  // edges clipped by the face rectangle
      for (int a_ = 0; a_ < 2; a_++) {
        const float a = bb_lc(ln[i], a_), b = bb_lc(ln[i], 3 + a_), c = bb_lc(ln[i], 1 - a_), d = bb_lc(ln[i], 4 - a_);
        if (fabsf(b) > EPS)
          for (int sj = -1; sj <= 1; sj += 2) {
            const float l = comp(ss, a_) * (float)sj;
            const float c1 = (l - a) / b;
            if (0.f <= c1 && c1 <= 1.f) {
              const float c2 = c + d * c1;
              if (fabsf(c2) <= comp(ss, 1 - a_) && n < BB_MAX) { pt[n] = ln[i].o + ln[i].d * c1; n++; }
            }
          }
      }
    if (m > 2) {  // corners of the face rectangle under the other box's face
      const float a = q[1].x, b = q[2].x, c = q[1].y, d = q[2].y;
      const float det = a * d - b * c;
      for (int i = 0; i < 4; i++) {
        const float llx = (i / 2) ? lx : -lx, lly = (i % 2) ? ly : -ly;
        const float x = llx - q[0].x, y = lly - q[0].y;
        const float u = (x * d - y * b) / det, v = (y * a - x * c) / det;
        if (0.f < u && u < 1.f && 0.f < v && v < 1.f && n < BB_MAX) { pt[n] = f3(llx, lly, q[0].z + u * q[1].z + v * q[2].z); n++; }
      }
    }
    for (int i = 0; i < (1 << (m - 1)); i++) {  // corners of the other box inside the rectangle
      const f3 t = q[i == 0 ? 0 : i + 2];
      if (!(i && (t.x <= -lx || t.x >= lx || t.y <= -ly || t.y >= ly)) && n < BB_MAX) { pt[n] = t; n++; }
    }
    const int m_all = n;
    n = 0;
    for (int i = 0; i < m_all; i++)
      if (pt[i].z <= 0.f) {
        f3 t = pt[i];
        t.z *= 0.5f;
        pt[n] = t;
        dep[n] = t.z;  // what the reference passes on: the halved z
        n++;
      }
    rw = bb_matmul(q2 ? mat2 : mat1, rm, false, true);
    origin = q2 ? pos2 : pos1;
    normal_0 = mat_col(rw, 2) * (q2 ? -1.f : 1.f);
  } else {  // ---- an edge of box 2 against an edge of box 1 (BOX:453-785)
    code -= 12;
    const int q1 = code / 3, q2 = code % 3;
    int ax1 = q2 == 1 ? 0 : 1, ax2 = q2 == 2 ? 0 : 2;
    int pax1 = q1 == 1 ? 0 : 1, pax2 = q1 == 2 ? 0 : 2;
    if (fabsf(rot.m[q1][ax1]) < fabsf(rot.m[q1][ax2])) { ax1 = ax2; ax2 = 3 - q2 - ax1; }
    if (fabsf(rot.m[pax1][q2]) < fabsf(rot.m[pax2][q2])) { pax1 = pax2; pax2 = 3 - q1 - pax1; }
    const int clface = (cle1 & (1 << pax2)) ? pax2 : pax2 + 3;
    const BBFrame F = bb_face_frame(clface, rm);
    f3 p = bb_perm(pos21, F);
    const f3 rnorm = bb_perm(clnorm, F);
    const mat3 r = bb_perm_rows(rot, F);
    f3 s = bb_mul_t(rm, size1);
    s = f3(fabsf(s.x), fabsf(s.y), fabsf(s.z));
    const float lx = s.x, ly = s.y;
    hz = s.z;
    p.z -= hz;
    const float s1 = (cle2 & (1 << ax1)) ? 1.f : -1.f, s2 = (cle2 & (1 << ax2)) ? 1.f : -1.f;
    const f3 e1 = mat_col(r, ax1) * comp(size2, ax1), e2 = mat_col(r, ax2) * comp(size2, ax2), eq = mat_col(r, q2) * comp(size2, q2);
    f3 pu[4];  // the two closest parallel edges of box 2 (end points)
    pu[0] = p + e1 * s1 + e2 * s2;
    pu[1] = pu[0];
    pu[0] = pu[0] + eq;
    pu[1] = pu[1] - eq;
    pu[2] = p + e1 * (-s1) + e2 * s2;
    pu[3] = pu[2];
    pu[2] = pu[2] + eq;
    pu[3] = pu[3] - eq;
    if (fabsf(rnorm.z) < EPS) return;
    const float innorm = (1.0f / rnorm.z) * (in_ ? -1.f : 1.f);
    f3 pp[4];  // ... projected along the contact normal onto the face plane z = 0
    for (int i = 0; i < 4; i++) pp[i] = pu[i] + rnorm * (-pu[i].z / rnorm.z);
    const f3 a0 = pu[0], a1 = pu[1] - pu[0], a2 = pu[2] - pu[0];
    const f3 b0 = pp[0], b1 = pp[1] - pp[0], b2 = pp[2] - pp[0];
    ln[0].o = b0; ln[0].d = b1; lnu[0].o = a0; lnu[0].d = a1;
    ln[1].o = b0; ln[1].d = b2; lnu[1].o = a0; lnu[1].d = a2;
    ln[2].o = b0 + b1; ln[2].d = b2; lnu[2].o = a0 + a1; lnu[2].d = a2;
    ln[3].o = b0 + b2; ln[3].d = b1; lnu[3].o = a0 + a2; lnu[3].d = a1;
    for (int i = 0; i < 4; i++)
      for (int a_ = 0; a_ < 2; a_++) {
        const float a = bb_lc(ln[i], a_), b = bb_lc(ln[i], a_ + 3), c = bb_lc(ln[i], 1 - a_), d = bb_lc(ln[i], 4 - a_);
        if (fabsf(b) > EPS)
          for (int sj = -1; sj <= 1; sj += 2) {
            const float l = comp(s, a_) * (float)sj;
            const float c1 = (l - a) / b;
            if (0.f <= c1 && c1 <= 1.f) {
              const float c2 = c + d * c1;
              if (fabsf(c2) <= comp(s, 1 - a_) && (lnu[i].o.z + lnu[i].d.z * c1) * innorm <= 0.f && n < BB_MAX) {
                f3 t = lnu[i].o * 0.5f + lnu[i].d * (c1 * 0.5f);
                set_comp(t, a_, comp(t, a_) + 0.5f * l);
                set_comp(t, 1 - a_, comp(t, 1 - a_) + 0.5f * c2);
                pt[n] = t;
                dep[n] = t.z * innorm * 2.f;
                n++;
              }
            }
          }
      }
    const int nl = n;
    {
      const float a = b1.x, b = b2.x, c = b1.y, d = b2.y;
      const float det = a * d - b * c;
      for (int i = 0; i < 4; i++) {
        const float llx = (i / 2) ? lx : -lx, lly = (i % 2) ? ly : -ly;
        const float x = llx - b0.x, y = lly - b0.y;
        float u = (x * d - y * b) / det, v = (y * a - x * c) / det;
        if (nl == 0) { if ((u < 0.f || u > 1.f) && (v < 0.f || v > 1.f)) continue; }
        else if (u < 0.f || u > 1.f || v < 0.f || v > 1.f) continue;
        u = u < 0.f ? 0.f : (u > 1.f ? 1.f : u);
        v = v < 0.f ? 0.f : (v > 1.f ? 1.f : v);
        const f3 w = pu[0] * (1.f - u - v) + pu[1] * u + pu[2] * v;
        const f3 corner(llx, lly, 0.f);
        const f3 dlt = corner - w;
        const float c2 = dot(dlt, dlt);
        if (!(w.z > 0.f && c2 > 0.f) && n < BB_MAX) { pt[n] = (corner + w) * 0.5f; dep[n] = sqrtf(c2) * (w.z < 0.f ? -1.f : 1.f); n++; }
      }
    }
    const int nf = n;
    for (int i = 0; i < 4; i++) {
      const float x = pp[i].x, y = pp[i].y;
      if (nl == 0) { if (nf != 0 && (x < -lx || x > lx) && (y < -ly || y > ly)) continue; }
      else if (x < -lx || x > lx || y < -ly || y > ly) continue;
      float c1 = 0.f;
      f3 t(x * 0.5f, y * 0.5f, 0.f);
      if (x < -s.x) { c1 = c1 + (x + s.x) * (x + s.x); t.x = -s.x * 0.5f; } else if (x > s.x) { c1 = c1 + (x - s.x) * (x - s.x); t.x = s.x * 0.5f; }
      if (y < -s.y) { c1 = c1 + (y + s.y) * (y + s.y); t.y = -s.y * 0.5f; } else if (y > s.y) { c1 = c1 + (y - s.y) * (y - s.y); t.y = s.y * 0.5f; }
      c1 = c1 + (pu[i].z * innorm) * (pu[i].z * innorm);
      if (pu[i].z > 0.f && c1 > 0.f) continue;
      if (n < BB_MAX) { pt[n] = t + pu[i] * 0.5f; dep[n] = sqrtf(c1) * (pu[i].z < 0.f ? -1.f : 1.f); n++; }
    }
    rw = bb_matmul(mat1, rm, false, true);
    origin = pos1;
    normal_0 = bb_mul(rw, rnorm) * (in_ ? -1.f : 1.f);
  }
  // ---- contact records: the first n_contacts_per_pair candidates that do not repeat an earlier one (BOX:425-452, 752-785)
  const f3 nrm = -normal_0;
  for (int i = 0; i < n; i++) {
    if (out.n >= ncpp) break;
    f3 t = pt[i];
    t.z += hz;
    const f3 cp = origin + bb_mul(rw, t);
    bool valid = true;
    for (int j = 0; j < out.n; j++) {
      const f3 d = cp - out.pos[j];
      if (fabsf(d.x) < EPS && fabsf(d.y) < EPS && fabsf(d.z) < EPS) valid = false;
    }
    if (valid) out.add(nrm, cp, -dep[i]);
  }
}

#ifdef __CUDACC__
template <int G>
__device__ __noinline__ void dev_box_box(const Ctx<G>& C, LaneContacts& out, int ga, int gb) {
  const float EPS = c_m.eps;
  f3 pos1, pos2;
  f4 quat1, quat2;
  geom_pose(C, ga, pos1, quat1);
  geom_pose(C, gb, pos2, quat2);
  const mat3 mat1 = quat_to_R(quat1, EPS), mat2 = quat_to_R(quat2, EPS);
  const f3 size1 = f3(GM(g_data[ga][0]), GM(g_data[ga][1]), GM(g_data[ga][2])) * 0.5f;
  const f3 size2 = f3(GM(g_data[gb][0]), GM(g_data[gb][1]), GM(g_data[gb][2])) * 0.5f;
  bb_contacts(pos1, mat1, size1, pos2, mat2, size2, EPS, min(c_m.n_contacts_per_pair, 5), out);
}
#endif
