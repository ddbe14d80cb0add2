// gsr_math.cuh -- register-resident 3-vector / quaternion / symmetric 3x3 algebra for the rigid-step kernels.
// Conventions follow the reference's device helpers (genesis/utils/geom.py:18-426): quaternions are (w,x,y,z),
// spatial quantities are (ang, vel) pairs expressed about the subtree COM in world axes.
#pragma once
// the small math helpers are real functions in the generic kernels (code size: instruction fetch is a contended resource there); the
// model-specialised builds are a quarter smaller and inline them (GSR_MATH_NOINLINE / GSR_ROUTINE_NOINLINE empty)
#ifndef GSR_MATH_NOINLINE
#define GSR_MATH_NOINLINE __noinline__
#endif
#ifndef GSR_ROUTINE_NOINLINE
#define GSR_ROUTINE_NOINLINE __noinline__
#endif
#ifndef GSR_ROWS_NOINLINE
#define GSR_ROWS_NOINLINE GSR_ROUTINE_NOINLINE
#endif
#ifndef GSR_UPD_NOINLINE
#define GSR_UPD_NOINLINE GSR_ROUTINE_NOINLINE
#endif
#ifndef GSR_HESS_NOINLINE
#define GSR_HESS_NOINLINE GSR_ROUTINE_NOINLINE
#endif
#ifndef GSR_GRAD_NOINLINE
#define GSR_GRAD_NOINLINE GSR_ROUTINE_NOINLINE
#endif
#ifndef GSR_FK_NOINLINE
#define GSR_FK_NOINLINE GSR_ROUTINE_NOINLINE
#endif
#ifndef GSR_MASS_NOINLINE
#define GSR_MASS_NOINLINE GSR_ROUTINE_NOINLINE
#endif
#include <cuda_runtime.h>
#include <math.h>

struct f3 {
  float x, y, z;
  __host__ __device__ f3() : x(0.f), y(0.f), z(0.f) {}
  __host__ __device__ f3(float a, float b, float c) : x(a), y(b), z(c) {}
};
struct f4 {
  float w, x, y, z;
  __host__ __device__ f4() : w(1.f), x(0.f), y(0.f), z(0.f) {}
  __host__ __device__ f4(float a, float b, float c, float d) : w(a), x(b), y(c), z(d) {}
};
// symmetric 3x3: xx xy xz yy yz zz
struct sym3 {
  float xx, xy, xz, yy, yz, zz;
};

__host__ __device__ __forceinline__ f3 operator+(f3 a, f3 b) { return f3(a.x + b.x, a.y + b.y, a.z + b.z); }
__host__ __device__ __forceinline__ f3 operator-(f3 a, f3 b) { return f3(a.x - b.x, a.y - b.y, a.z - b.z); }
__host__ __device__ __forceinline__ f3 operator-(f3 a) { return f3(-a.x, -a.y, -a.z); }
__host__ __device__ __forceinline__ f3 operator*(f3 a, float s) { return f3(a.x * s, a.y * s, a.z * s); }
__host__ __device__ __forceinline__ f3 operator*(float s, f3 a) { return f3(a.x * s, a.y * s, a.z * s); }
__host__ __device__ __forceinline__ f3 operator/(f3 a, float s) { return f3(a.x / s, a.y / s, a.z / s); }
__host__ __device__ __forceinline__ void operator+=(f3& a, f3 b) { a.x += b.x; a.y += b.y; a.z += b.z; }
__host__ __device__ __forceinline__ void operator-=(f3& a, f3 b) { a.x -= b.x; a.y -= b.y; a.z -= b.z; }
__host__ __device__ __forceinline__ float dot(f3 a, f3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
__host__ __device__ __forceinline__ f3 cross(f3 a, f3 b) { return f3(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x); }
__host__ __device__ __forceinline__ float norm_sqr(f3 a) { return dot(a, a); }
__host__ __device__ __forceinline__ float norm(f3 a) { return sqrtf(dot(a, a)); }
__host__ __device__ __forceinline__ f3 normalized(f3 a) { float n = norm(a); return f3(a.x / n, a.y / n, a.z / n); }
__host__ __device__ __forceinline__ float comp(f3 a, int i) { return i == 0 ? a.x : (i == 1 ? a.y : a.z); }
__host__ __device__ __forceinline__ void set_comp(f3& a, int i, float v) { if (i == 0) a.x = v; else if (i == 1) a.y = v; else a.z = v; }

__host__ __device__ __forceinline__ f4 quat_inv(f4 q) { return f4(q.w, -q.x, -q.y, -q.z); }
// geom.py:235-242
__host__ __device__ __forceinline__ f4 quat_mul(f4 u, f4 v) {
  return f4(u.w * v.w - u.x * v.x - u.y * v.y - u.z * v.z, u.w * v.x + u.x * v.w + u.y * v.z - u.z * v.y,
            u.w * v.y - u.x * v.z + u.y * v.w + u.z * v.x, u.w * v.z + u.x * v.y - u.y * v.x + u.z * v.w);
}
// geom.py:224-232 quat_mul_axis: q * (0, axis)
__host__ __device__ __forceinline__ f4 quat_mul_axis(f4 q, f3 a) {
  return f4(-q.x * a.x - q.y * a.y - q.z * a.z, q.w * a.x + q.y * a.z - q.z * a.y, q.w * a.y + q.z * a.x - q.x * a.z, q.w * a.z + q.x * a.y - q.y * a.x);
}
// geom.py:245-252: transform_quat_by_quat(v, u) = normalize(u * v)
__host__ __device__ GSR_MATH_NOINLINE f4 quat_compose(f4 v, f4 u) {
  f4 r = quat_mul(u, v);
  float n = 1.0f / sqrtf(r.w * r.w + r.x * r.x + r.y * r.y + r.z * r.z);
  return f4(r.w * n, r.x * n, r.y * n, r.z * n);
}
// geom.py:255-270
__host__ __device__ GSR_MATH_NOINLINE f3 rotate(f3 v, f4 q) {
  float xx = q.x * q.x, xy = q.x * q.y, xz = q.x * q.z, wx = q.x * q.w;
  float yy = q.y * q.y, yz = q.y * q.z, wy = q.y * q.w;
  float zz = q.z * q.z, wz = q.z * q.w, ww = q.w * q.w;
  float inv = 1.0f / (ww + xx + yy + zz);
  return f3((v.x * (xx + ww - yy - zz) + v.y * (2.f * xy - 2.f * wz) + v.z * (2.f * xz + 2.f * wy)) * inv,
            (v.x * (2.f * wz + 2.f * xy) + v.y * (ww - xx + yy - zz) + v.z * (-2.f * wx + 2.f * yz)) * inv,
            (v.x * (-2.f * wy + 2.f * xz) + v.y * (2.f * wx + 2.f * yz) + v.z * (ww - xx - yy + zz)) * inv);
}
__host__ __device__ __forceinline__ f3 rotate_inv(f3 v, f4 q) { return rotate(v, quat_inv(q)); }
__host__ __device__ __forceinline__ f3 transform(f3 p, f3 t, f4 q) { return rotate(p, q) + t; }

struct mat3 {
  float m[3][3];
};
// geom.py:136-161
__host__ __device__ GSR_MATH_NOINLINE mat3 quat_to_R(f4 q, float eps) {
  mat3 R;
  R.m[0][0] = 1.f; R.m[0][1] = 0.f; R.m[0][2] = 0.f; R.m[1][0] = 0.f; R.m[1][1] = 1.f; R.m[1][2] = 0.f;
  R.m[2][0] = 0.f; R.m[2][1] = 0.f; R.m[2][2] = 1.f;
  float d = q.w * q.w + q.x * q.x + q.y * q.y + q.z * q.z;
  if (d > eps) {
    float s = 2.0f / d;
    float xs = q.x * s, ys = q.y * s, zs = q.z * s;
    float wx = q.w * xs, wy = q.w * ys, wz = q.w * zs;
    float xx = q.x * xs, xy = q.x * ys, xz = q.x * zs;
    float yy = q.y * ys, yz = q.y * zs, zz = q.z * zs;
    R.m[0][0] = 1.f - (yy + zz); R.m[0][1] = xy - wz; R.m[0][2] = xz + wy;
    R.m[1][0] = xy + wz; R.m[1][1] = 1.f - (xx + zz); R.m[1][2] = yz - wx;
    R.m[2][0] = xz - wy; R.m[2][1] = yz + wx; R.m[2][2] = 1.f - (xx + yy);
  }
  return R;
}
__host__ __device__ __forceinline__ f3 mat_col(const mat3& R, int j) { return f3(R.m[0][j], R.m[1][j], R.m[2][j]); }

// geom.py:110-133
__host__ __device__ GSR_MATH_NOINLINE f4 rotvec_to_quat(f3 rv, float eps) {
  f4 q(0.f, 0.f, 0.f, 0.f);
  float thetasq = norm_sqr(rv);
  if (thetasq > eps * eps) {
    float theta = sqrtf(thetasq);
    float s, c;
    sincosf(0.5f * theta, &s, &c);
    float k = s / theta;
    q = f4(c, k * rv.x, k * rv.y, k * rv.z);
    float r = 0.5f * (3.0f - (q.w * q.w + q.x * q.x + q.y * q.y + q.z * q.z));
    q = f4(q.w * r, q.x * r, q.y * r, q.z * r);
  } else {
    q.w = 1.f;
  }
  return q;
}
// geom.py:164-193
__host__ __device__ __noinline__ f3 quat_to_xyz(f4 q, float eps) {
  float roll = 0.f, pitch = 0.f, yaw = 0.f;
  float n2 = q.w * q.w + q.x * q.x + q.y * q.y + q.z * q.z;
  if (n2 > eps) {
    float s = 2.0f / n2;
    float xs = q.x * s, ys = q.y * s, zs = q.z * s;
    float wx = q.w * xs, wy = q.w * ys, wz = q.w * zs;
    float xx = q.x * xs, xy = q.x * ys, xz = q.x * zs;
    float yy = q.y * ys, yz = q.y * zs, zz = q.z * zs;
    float sinycosp = wz - xy, cosycosp = 1.f - (yy + zz);
    float cosp = sqrtf(cosycosp * cosycosp + sinycosp * sinycosp);
    pitch = atan2f(xz + wy, cosp);
    if (cosp > eps) {
      roll = atan2f(wx - yz, 1.f - (xx + yy));
      yaw = atan2f(sinycosp, cosycosp);
    } else {
      yaw = atan2f(wz + xy, 1.f - (xx + zz));
    }
  }
  return f3(roll, pitch, yaw);
}
// geom.py:45-65
__host__ __device__ __noinline__ f4 xyz_to_quat(f3 xyz) {
  float si, ci, sj, cj, sk, ck;
  sincosf(0.5f * xyz.z, &si, &ci);
  sincosf(-0.5f * xyz.y, &sj, &cj);
  sincosf(0.5f * xyz.x, &sk, &ck);
  float cc = ci * ck, cs = ci * sk, sc = si * ck, ss = si * sk;
  return f4(cj * cc + sj * ss, cj * cs - sj * sc, -cj * ss - sj * cc, cj * sc - sj * cs);
}
// geom.py:196-207
__host__ __device__ __noinline__ f3 quat_to_rotvec(f4 q, float eps) {
  f3 rv(q.x, q.y, q.z);
  float s2 = norm(rv);
  if (s2 > eps) {
    float angle = 2.0f * atan2f(s2, fabsf(q.w));
    rv = rv * ((q.w < 0.f ? -1.f : 1.f) * (angle / s2));
  }
  return rv;
}

__host__ __device__ __forceinline__ f3 sym_mul(const sym3& I, f3 v) {
  return f3(I.xx * v.x + I.xy * v.y + I.xz * v.z, I.xy * v.x + I.yy * v.y + I.yz * v.z, I.xz * v.x + I.yz * v.y + I.zz * v.z);
}
// geom.py:365-369 inertial_mul: (ang, vel) -> (I ang + pos x vel, m vel - pos x ang)
__host__ __device__ __forceinline__ void inertial_mul(f3 pos, const sym3& I, float mass, f3 vel, f3 ang, f3& o_ang, f3& o_vel) {
  o_ang = sym_mul(I, ang) + cross(pos, vel);
  o_vel = vel * mass - cross(pos, ang);
}
// geom.py:386-403
__host__ __device__ __noinline__ void orthogonals(f3 a, f3& d1, f3& d2) {
  f3 b;
  if (fabsf(a.y) < 0.5f) b = f3(-a.x * a.y, 1.f - a.y * a.y, -a.z * a.y);
  else b = f3(-a.x * a.z, -a.y * a.z, 1.f - a.z * a.z);
  b = normalized(b);
  d1 = cross(b, a);
  d2 = b;
}
// geom.py:406-424
__host__ __device__ __noinline__ void imp_aref(const float* p, float neg_penetration, float vel, float pos, float& imp, float& aref) {
  float timeconst = p[0], dampratio = p[1], dmin = p[2], dmax = p[3], width = p[4], mid = p[5], power = p[6];
  float imp_x = fabsf(neg_penetration) / width;
  float imp_a = (1.0f / powf(mid, power - 1.f)) * powf(imp_x, power);
  float imp_b = 1.0f - (1.0f / powf(1.0f - mid, power - 1.f)) * powf(1.0f - imp_x, power);
  float imp_y = imp_x < mid ? imp_a : imp_b;
  imp = dmin + imp_y * (dmax - dmin);
  imp = fminf(fmaxf(imp, dmin), dmax);
  if (imp_x > 1.0f) imp = dmax;
  float b = 2.0f / (dmax * timeconst);
  float k = 1.0f / (dmax * dmax * timeconst * timeconst * dampratio * dampratio);
  aref = -b * vel - k * imp * pos;
}

// ---- analytic capsule contacts (collider/capsule_contact.py, collider/utils.py:9-65) ----------------------------------------
// closest points of two segments (Ericson, Real-Time Collision Detection 5.1.9)
__host__ __device__ __noinline__ void closest_points_on_segments(f3 p1, f3 p2, f3 q1, f3 q2, float eps, f3& ca, f3& cb) {
  const f3 da = p2 - p1, db = q2 - q1, r = p1 - q1;
  const float a = dot(da, da), b = dot(da, db), c = dot(db, db), d = dot(da, r), f = dot(db, r);
  const float denom = a * c - b * b;
  float sp = 0.f, tp = 0.f;
  if (denom < eps) {
    tp = c > eps ? fminf(fmaxf(f / c, 0.f), 1.f) : 0.f;
  } else {
    sp = fminf(fmaxf((b * f - c * d) / denom, 0.f), 1.f);
    tp = fminf(fmaxf(c > eps ? (b * sp + f) / c : 0.f, 0.f), 1.f);
    const float s_new = fminf(fmaxf(a > eps ? (b * tp - d) / a : 0.f, 0.f), 1.f);
    if (a > eps) sp = s_new;
  }
  ca = p1 + da * sp;
  cb = q1 + db * tp;
}
// capsule_contact.py:10-98: capsules along their local z, data = (radius, half length); normal points from B to A
__host__ __device__ __noinline__ void capsule_capsule_contact(f3 pa, f4 qa, float ra, float hla, f3 pb, f4 qb, float rb, float hlb, float eps,
                                                              bool& is_col, f3& normal, f3& cpos, float& pen) {
  const f3 axa = rotate(f3(0.f, 0.f, 1.f), qa), axb = rotate(f3(0.f, 0.f, 1.f), qb);
  f3 Pa, Pb;
  closest_points_on_segments(pa - axa * hla, pa + axa * hla, pb - axb * hlb, pb + axb * hlb, eps, Pa, Pb);
  const f3 diff = Pa - Pb;
  const float dist_sq = dot(diff, diff), cr = ra + rb;
  is_col = false; normal = f3(1.f, 0.f, 0.f); cpos = f3(); pen = 0.f;
  if (dist_sq < cr * cr) {
    is_col = true;
    const float dist = sqrtf(dist_sq);
    if (dist > eps) normal = diff / dist;
    else {
      const f3 nd = cross(axa, axb);
      const float ndl = dot(nd, nd);
      if (ndl > eps) normal = nd / ndl;  // (sic) the reference divides by the squared length
      else normal = fabsf(axa.x) < 0.9f ? f3(0.f, -axa.z, axa.y) : f3(axa.z, 0.f, -axa.x);
    }
    pen = cr - dist;
    cpos = Pa - normal * ra + normal * (0.5f * pen);
  }
}
// capsule_contact.py:102-196 with A the sphere and B the capsule (the geoms of a pair are sorted by type)
__host__ __device__ __noinline__ void sphere_capsule_contact(f3 sc, float rs, f3 cc, f4 cq, float rc, float hl, float eps,
                                                             bool& is_col, f3& normal, f3& cpos, float& pen) {
  const f3 ax = rotate(f3(0.f, 0.f, 1.f), cq);
  const f3 P1 = cc - ax * hl, seg = (cc + ax * hl) - P1;
  const float l2 = dot(seg, seg);
  float t = 0.5f;
  if (l2 > eps) t = fminf(fmaxf(dot(sc - P1, seg) / l2, 0.f), 1.f);
  const f3 diff = sc - (P1 + seg * t);
  const float dist_sq = dot(diff, diff), cr = rs + rc;
  is_col = false; normal = f3(1.f, 0.f, 0.f); cpos = f3(); pen = 0.f;
  if (dist_sq < cr * cr) {
    is_col = true;
    const float dist = sqrtf(dist_sq);
    if (dist > eps) normal = diff / dist;
    else normal = fabsf(ax.x) < 0.9f ? f3(0.f, -ax.z, ax.y) : f3(ax.z, 0.f, -ax.x);
    pen = cr - dist;
    cpos = sc - normal * (rs - 0.5f * pen);
  }
}
