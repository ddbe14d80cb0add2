// TEST INFRASTRUCTURE ONLY (oracle/): small vector/quaternion algebra for the CPU restatement.
// Every helper restates one device helper of the reference's genesis/utils/geom.py (file:line cited per function).
#pragma once
#include <cmath>
#include <algorithm>

namespace orc {

template <typename T>
struct V3 {
  T x, y, z;
  V3() : x(0), y(0), z(0) {}
  V3(T a, T b, T c) : x(a), y(b), z(c) {}
  T& operator[](int i) { return i == 0 ? x : (i == 1 ? y : z); }
  T operator[](int i) const { return i == 0 ? x : (i == 1 ? y : z); }
  V3 operator+(const V3& o) const { return V3(x + o.x, y + o.y, z + o.z); }
  V3 operator-(const V3& o) const { return V3(x - o.x, y - o.y, z - o.z); }
  V3 operator-() const { return V3(-x, -y, -z); }
  V3 operator*(T s) const { return V3(x * s, y * s, z * s); }
  V3 operator/(T s) const { return V3(x / s, y / s, z / s); }
  V3& operator+=(const V3& o) { x += o.x; y += o.y; z += o.z; return *this; }
  V3& operator-=(const V3& o) { x -= o.x; y -= o.y; z -= o.z; return *this; }
  T dot(const V3& o) const { return x * o.x + y * o.y + z * o.z; }
  V3 cross(const V3& o) const { return V3(y * o.z - z * o.y, z * o.x - x * o.z, x * o.y - y * o.x); }
  T norm_sqr() const { return x * x + y * y + z * z; }
  T norm() const { return std::sqrt(norm_sqr()); }
  V3 normalized() const { T n = norm(); return V3(x / n, y / n, z / n); }
};
template <typename T> V3<T> operator*(T s, const V3<T>& v) { return v * s; }

template <typename T>
struct Q4 {
  T w, x, y, z;
  Q4() : w(1), x(0), y(0), z(0) {}
  Q4(T a, T b, T c, T d) : w(a), x(b), y(c), z(d) {}
};

template <typename T>
struct M3 {
  T m[3][3];
  M3() { for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) m[i][j] = 0; }
  V3<T> mul(const V3<T>& v) const {
    return V3<T>(m[0][0] * v.x + m[0][1] * v.y + m[0][2] * v.z, m[1][0] * v.x + m[1][1] * v.y + m[1][2] * v.z,
                 m[2][0] * v.x + m[2][1] * v.y + m[2][2] * v.z);
  }
  V3<T> col(int j) const { return V3<T>(m[0][j], m[1][j], m[2][j]); }
  V3<T> row(int i) const { return V3<T>(m[i][0], m[i][1], m[i][2]); }
  M3 operator+(const M3& o) const { M3 r; for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) r.m[i][j] = m[i][j] + o.m[i][j]; return r; }
  M3& operator+=(const M3& o) { for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) m[i][j] += o.m[i][j]; return *this; }
};

// geom.py:218-220
template <typename T> Q4<T> inv_quat(const Q4<T>& q) { return Q4<T>(q.w, -q.x, -q.y, -q.z); }

// geom.py:235-242 (quat_mul(u, v) = u (x) v)
template <typename T> Q4<T> quat_mul(const Q4<T>& u, const Q4<T>& v) {
  return Q4<T>(u.w * v.w - u.x * v.x - u.y * v.y - u.z * v.z, u.w * v.x + u.x * v.w + u.y * v.z - u.z * v.y,
               u.w * v.y - u.x * v.z + u.y * v.w + u.z * v.x, u.w * v.z + u.x * v.y - u.y * v.x + u.z * v.w);
}

// geom.py:245-252: transform_quat_by_quat(v, u) = normalize(u (x) v)
template <typename T> Q4<T> transform_quat_by_quat(const Q4<T>& v, const Q4<T>& u) {
  Q4<T> r = quat_mul(u, v);
  T n = std::sqrt(r.w * r.w + r.x * r.x + r.y * r.y + r.z * r.z);
  return Q4<T>(r.w / n, r.x / n, r.y / n, r.z / n);
}

// geom.py:255-270
template <typename T> V3<T> transform_by_quat(const V3<T>& v, const Q4<T>& q) {
  T xx = q.x * q.x, xy = q.x * q.y, xz = q.x * q.z, wx = q.x * q.w;
  T yy = q.y * q.y, yz = q.y * q.z, wy = q.y * q.w;
  T zz = q.z * q.z, wz = q.z * q.w, ww = q.w * q.w;
  T d = ww + xx + yy + zz;
  return V3<T>((v.x * (xx + ww - yy - zz) + v.y * (2 * xy - 2 * wz) + v.z * (2 * xz + 2 * wy)) / d,
               (v.x * (2 * wz + 2 * xy) + v.y * (ww - xx + yy - zz) + v.z * (-2 * wx + 2 * yz)) / d,
               (v.x * (-2 * wy + 2 * xz) + v.y * (2 * wx + 2 * yz) + v.z * (ww - xx - yy + zz)) / d);
}
template <typename T> V3<T> inv_transform_by_quat(const V3<T>& v, const Q4<T>& q) { return transform_by_quat(v, inv_quat(q)); }
// geom.py:288-290
template <typename T> V3<T> transform_by_trans_quat(const V3<T>& p, const V3<T>& t, const Q4<T>& q) { return transform_by_quat(p, q) + t; }

// geom.py:136-161
template <typename T> M3<T> quat_to_R(const Q4<T>& q, T eps) {
  M3<T> R;
  R.m[0][0] = R.m[1][1] = R.m[2][2] = 1;
  T d = q.w * q.w + q.x * q.x + q.y * q.y + q.z * q.z;
  if (d > eps) {
    T s = T(2) / d;
    T xs = q.x * s, ys = q.y * s, zs = q.z * s;
    T wx = q.w * xs, wy = q.w * ys, wz = q.w * zs;
    T xx = q.x * xs, xy = q.x * ys, xz = q.x * zs;
    T yy = q.y * ys, yz = q.y * zs, zz = q.z * zs;
    R.m[0][0] = 1 - (yy + zz); R.m[0][1] = xy - wz; R.m[0][2] = xz + wy;
    R.m[1][0] = xy + wz; R.m[1][1] = 1 - (xx + zz); R.m[1][2] = yz - wx;
    R.m[2][0] = xz - wy; R.m[2][1] = yz + wx; R.m[2][2] = 1 - (xx + yy);
  }
  return R;
}

// geom.py:110-133
template <typename T> Q4<T> rotvec_to_quat(const V3<T>& rv, T eps) {
  Q4<T> q(0, 0, 0, 0);
  T thetasq = rv.norm_sqr();
  if (thetasq > eps * eps) {
    T theta = std::sqrt(thetasq);
    T c = std::cos(T(0.5) * theta), s = std::sin(T(0.5) * theta);
    q.w = c;
    q.x = s / theta * rv.x; q.y = s / theta * rv.y; q.z = s / theta * rv.z;
    T k = T(0.5) * (T(3) - (q.w * q.w + q.x * q.x + q.y * q.y + q.z * q.z));
    q.w *= k; q.x *= k; q.y *= k; q.z *= k;
  } else {
    q.w = 1;
  }
  return q;
}

// geom.py:164-193
template <typename T> V3<T> quat_to_xyz(const Q4<T>& q, T eps) {
  T roll = 0, pitch = 0, yaw = 0;
  T n2 = q.w * q.w + q.x * q.x + q.y * q.y + q.z * q.z;
  if (n2 > eps) {
    T s = T(2) / n2;
    T xs = q.x * s, ys = q.y * s, zs = q.z * s;
    T wx = q.w * xs, wy = q.w * ys, wz = q.w * zs;
    T xx = q.x * xs, xy = q.x * ys, xz = q.x * zs;
    T yy = q.y * ys, yz = q.y * zs, zz = q.z * zs;
    T sinycosp = wz - xy, cosycosp = 1 - (yy + zz);
    T cosp = std::sqrt(cosycosp * cosycosp + sinycosp * sinycosp);
    pitch = std::atan2(xz + wy, cosp);
    if (cosp > eps) {
      roll = std::atan2(wx - yz, 1 - (xx + yy));
      yaw = std::atan2(sinycosp, cosycosp);
    } else {
      yaw = std::atan2(wz + xy, 1 - (xx + zz));
    }
  }
  return V3<T>(roll, pitch, yaw);
}

// geom.py:45-65
template <typename T> Q4<T> xyz_to_quat(const V3<T>& xyz) {
  T ai = T(0.5) * xyz.z, aj = -T(0.5) * xyz.y, ak = T(0.5) * xyz.x;
  T si = std::sin(ai), sj = std::sin(aj), sk = std::sin(ak);
  T ci = std::cos(ai), cj = std::cos(aj), ck = std::cos(ak);
  T cc = ci * ck, cs = ci * sk, sc = si * ck, ss = si * sk;
  return Q4<T>(cj * cc + sj * ss, cj * cs - sj * sc, -cj * ss - sj * cc, cj * sc - sj * cs);
}

// geom.py:196-207
template <typename T> V3<T> quat_to_rotvec(const Q4<T>& q, T eps) {
  V3<T> rv(q.x, q.y, q.z);
  T s2 = rv.norm();
  if (s2 > eps) {
    T angle = T(2) * std::atan2(s2, std::abs(q.w));
    T inv_sinc = angle / s2;
    rv = rv * ((q.w < 0 ? T(-1) : T(1)) * inv_sinc);
  }
  return rv;
}

// geom.py:365-369: (ang, vel) = I_spatial * (vel, ang) with first-moment `pos`
template <typename T> void inertial_mul(const V3<T>& pos, const M3<T>& I, T mass, const V3<T>& vel, const V3<T>& ang, V3<T>& o_ang, V3<T>& o_vel) {
  o_ang = I.mul(ang) + pos.cross(vel);
  o_vel = vel * mass - pos.cross(ang);
}
// geom.py:372-376
template <typename T> void motion_cross_force(const V3<T>& m_ang, const V3<T>& m_vel, const V3<T>& f_ang, const V3<T>& f_vel, V3<T>& o_ang, V3<T>& o_vel) {
  o_vel = m_ang.cross(f_vel);
  o_ang = m_ang.cross(f_ang) + m_vel.cross(f_vel);
}
// geom.py:379-383
template <typename T> void motion_cross_motion(const V3<T>& s_ang, const V3<T>& s_vel, const V3<T>& m_ang, const V3<T>& m_vel, V3<T>& o_ang, V3<T>& o_vel) {
  o_vel = s_ang.cross(m_vel) + s_vel.cross(m_ang);
  o_ang = s_ang.cross(m_ang);
}

// geom.py:386-403
template <typename T> void orthogonals(const V3<T>& a, V3<T>& d1, V3<T>& d2) {
  V3<T> b;
  if (std::abs(a.y) < T(0.5)) {
    b = V3<T>(-a.x * a.y, T(1) - a.y * a.y, -a.z * a.y);
  } else {
    b = V3<T>(-a.x * a.z, -a.y * a.z, T(1) - a.z * a.z);
  }
  b = b.normalized();
  d1 = b.cross(a);
  d2 = b;
}

// geom.py:406-424
template <typename T> void imp_aref(const T* p, T neg_penetration, T vel, T pos, T& imp, T& aref) {
  T timeconst = p[0], dampratio = p[1], dmin = p[2], dmax = p[3], width = p[4], mid = p[5], power = p[6];
  T imp_x = std::abs(neg_penetration) / width;
  T imp_a = (T(1) / std::pow(mid, power - 1)) * std::pow(imp_x, power);
  T imp_b = T(1) - (T(1) / std::pow(T(1) - mid, power - 1)) * std::pow(T(1) - imp_x, power);
  T imp_y = imp_x < mid ? imp_a : imp_b;
  imp = dmin + imp_y * (dmax - dmin);
  imp = std::min(std::max(imp, dmin), dmax);
  if (imp_x > T(1)) imp = dmax;
  T b = T(2) / (dmax * timeconst);
  T k = T(1) / (dmax * dmax * timeconst * timeconst * dampratio * dampratio);
  aref = -b * vel - k * imp * pos;
}

}  // namespace orc
