// gsr_group.cuh -- lane-group primitives and shared-memory accessors of the group-per-environment kernels.
#pragma once
#include "gsr_types.h"
#include "gsr_math.cuh"

__constant__ DevModel c_m;     // warp-uniform reads (scalars, loop bounds, tables indexed by a uniform loop variable)
#ifdef GSR_SPEC_HEADER
// Model-specialised build (tools/specialize.py): the record / workspace layouts and the model's counts and options are literals of
// the generated header -- offsets become immediates, loop bounds and option branches fold.  The library only accepts the model
// it was built for (gsr_api.cu checks the layouts at commit).  The run-time copies stay for the batch-dependent segment bases.
#include GSR_SPEC_HEADER
__constant__ HotLayout c_h_rt;
__constant__ SmemLayout c_s_rt;
#define c_h SpecH()
#define c_s SpecS()
#define c_hb c_h_rt
#else
__constant__ HotLayout c_h;
__constant__ SmemLayout c_s;
#define c_hb c_h
#endif

// Lane-varying table reads go to a global-memory copy of the model (read-only path, L1 resident): constant memory
// serialises divergent addresses.
#define CM(x) CMX_##x
#define CMI(x) CMX_##x  // the same, for a use nested inside CM(...) (a macro does not expand inside itself)
#include "gsr_cm_fields.h"
#define GM(x) (__ldg(&(C.gm->x)))

template <int G>
struct Ctx {
  float* S;            // this env's shared-memory workspace
  float* rec;          // this env's HOT record (global)
  float* contacts;     // this env's CONTACT records (global)
  float* cache;        // this env's CACHE record (global)
  float* spill;        // this env's SPILL record (global)
  const DevModel* gm;  // global-memory copy of the model
  int lane;            // lane index inside the group
  int env;             // environment index
  unsigned gmask;      // warp lanes that belong to this group

  __device__ __forceinline__ void sync() const { __syncwarp(gmask); }
  __device__ __forceinline__ float sum(float v) const {
#pragma unroll
    for (int o = G / 2; o > 0; o >>= 1) v += __shfl_xor_sync(gmask, v, o, G);
    return v;
  }
  __device__ __forceinline__ int isum(int v) const {
#pragma unroll
    for (int o = G / 2; o > 0; o >>= 1) v += __shfl_xor_sync(gmask, v, o, G);
    return v;
  }
  // ballot restricted to the group, bit i = lane i of the group
  __device__ __forceinline__ unsigned ballot(bool p) const {
    const unsigned b = __ballot_sync(gmask, p);
    if (G == 32) return b;
    const int shift = (threadIdx.x & 31) - lane;
    return (b >> shift) & ((1u << (G & 31)) - 1u);
  }
  __device__ __forceinline__ bool any(bool p) const { return ballot(p) != 0u; }
  __device__ __forceinline__ int* SI() const { return reinterpret_cast<int*>(S); }
};

__device__ __forceinline__ f3 ld3(const float* p, int i) { return f3(p[3 * i], p[3 * i + 1], p[3 * i + 2]); }
__device__ __forceinline__ void st3(float* p, int i, f3 v) { p[3 * i] = v.x; p[3 * i + 1] = v.y; p[3 * i + 2] = v.z; }
__device__ __forceinline__ f4 ld4(const float* p, int i) { return f4(p[4 * i], p[4 * i + 1], p[4 * i + 2], p[4 * i + 3]); }
__device__ __forceinline__ void st4(float* p, int i, f4 v) { p[4 * i] = v.w; p[4 * i + 1] = v.x; p[4 * i + 2] = v.y; p[4 * i + 3] = v.z; }
__device__ __forceinline__ sym3 ld6(const float* p, int i) {
  sym3 r; r.xx = p[6 * i]; r.xy = p[6 * i + 1]; r.xz = p[6 * i + 2]; r.yy = p[6 * i + 3]; r.yz = p[6 * i + 4]; r.zz = p[6 * i + 5];
  return r;
}
__device__ __forceinline__ void st6(float* p, int i, const sym3& v) {
  p[6 * i] = v.xx; p[6 * i + 1] = v.xy; p[6 * i + 2] = v.xz; p[6 * i + 3] = v.yy; p[6 * i + 4] = v.yz; p[6 * i + 5] = v.zz;
}
__device__ __forceinline__ f3 g3(const float* p) { return f3(__ldg(p), __ldg(p + 1), __ldg(p + 2)); }
__device__ __forceinline__ f4 g4(const float* p) { return f4(__ldg(p), __ldg(p + 1), __ldg(p + 2), __ldg(p + 3)); }
__device__ __forceinline__ bool bit64(unsigned long long m, int i) { return (m >> i) & 1ull; }
// iterate the set bits of a 64-bit mask in ascending / descending order
#define FOR_BITS_UP(m_, i_) for (unsigned long long mm_ = (m_); mm_ != 0ull; mm_ &= mm_ - 1ull) { const int i_ = __ffsll((long long)mm_) - 1;
#define FOR_BITS_DOWN(m_, i_) for (unsigned long long mm_ = (m_); mm_ != 0ull; ) { const int i_ = 63 - __clzll((long long)mm_); mm_ &= ~(1ull << i_);
#define END_BITS }

// lower-triangle task decode: t = i (i + 1) / 2 + j, j <= i (table built on the host, read through the L1 read-only path)
#define tri_decode(t, i, j) { const unsigned tr_ = GM(tri[t]); i = (int)(tr_ >> 8); j = (int)(tr_ & 255u); }

// The mass matrix lives in the workspace as a PACKED lower triangle: entry (i, j), j <= i, at i (i + 1) / 2 + j -- the order of the
// `tri` task table, so the assembly loops address it by their task index.  Row i times a vector, terms in ascending j like a dense row:
__device__ __forceinline__ float sym_row_dot(const float* Mp, int i, const float* x, int n) {
  float s = 0.f;
  const float* r = Mp + i * (i + 1) / 2;
  for (int j = 0; j <= i; j++) s += r[j] * x[j];
  int idx = (i + 1) * (i + 2) / 2 + i;
  for (int j = i + 1; j < n; j++) { s += Mp[idx] * x[j]; idx += j + 1; }
  return s;
}
__device__ __forceinline__ float sym_get(const float* Mp, int i, int j) { return i >= j ? Mp[i * (i + 1) / 2 + j] : Mp[j * (j + 1) / 2 + i]; }

// per-environment inertial parameters: model value plus the domain-randomisation shift of the record (FK:377-382, 437)
template <int G>
__device__ __forceinline__ float link_mass(const Ctx<G>& C, int l, bool shifted) {
  float m = GM(l_mass[l]);
  if (shifted) m += C.rec[c_h.mass_shift + l];
  return m;
}
template <int G>
__device__ __forceinline__ f3 link_ipos(const Ctx<G>& C, int l, bool shifted) {
  f3 p(GM(l_ipos[l][0]), GM(l_ipos[l][1]), GM(l_ipos[l][2]));
  if (shifted) p += ld3(C.rec + c_h.com_shift, l);
  return p;
}

// dof parameters: model constants, or the environment's own values when RigidOptions.batch_dofs_info is on (uniform branch)
#ifdef GSR_NO_BATCH_DOFS
#define GSR_BD(x) false
#else
#define GSR_BD(x) (x)
#endif
#define GSR_DOF_PARAM(fn, field, expr)                                                                     \
  template <int G>                                                                                         \
  __device__ __forceinline__ float fn(const Ctx<G>& C, int d) { return GSR_BD(CM(batch_dofs)) ? C.rec[c_h.field + d] : (expr); }
GSR_DOF_PARAM(dof_kp, bd_kp, GM(d_kp[d]))
GSR_DOF_PARAM(dof_armature, bd_armature, GM(d_armature[d]))
GSR_DOF_PARAM(dof_damping, bd_damping, GM(d_damping[d]))
GSR_DOF_PARAM(dof_stiffness, bd_stiffness, GM(d_stiffness[d]))
GSR_DOF_PARAM(dof_frictionloss, bd_frictionloss, GM(d_frictionloss[d]))
GSR_DOF_PARAM(dof_invweight, bd_invweight, GM(d_invweight[d]))
#undef GSR_DOF_PARAM
template <int G>
__device__ __forceinline__ float dof_bias(const Ctx<G>& C, int d, int k) { return GSR_BD(CM(batch_dofs)) ? C.rec[c_h.bd_bias + 3 * d + k] : GM(d_act_bias[d][k]); }
template <int G>
__device__ __forceinline__ float dof_limit(const Ctx<G>& C, int d, int k) { return GSR_BD(CM(batch_dofs)) ? C.rec[c_h.bd_limit + 2 * d + k] : GM(d_limit[d][k]); }
template <int G>
__device__ __forceinline__ float dof_force_range(const Ctx<G>& C, int d, int k) { return GSR_BD(CM(batch_dofs)) ? C.rec[c_h.bd_force_range + 2 * d + k] : GM(d_force_range[d][k]); }

// IN / WK reach shared memory through a generic pointer (generic LD / ST in SASS): telling the compiler the address space of
// EVERY access (`__builtin_assume(__isShared(C.S))` in both macros) gave wrong results in 16-20 parity tests, for either half
// of the accessors (profiles/README.md item 13); the routines that are known to be safe use WKS below.
__device__ __forceinline__ float* smem_ptr(float* p) { __builtin_assume(__isShared(p)); return p; }
#define IN_GENERIC(field) (C.S + c_h.field)
#define WK_GENERIC(field) (C.S + c_s.field)
#define IN_SHARED(field) (smem_ptr(C.S) + c_h.field)
#define WK_SHARED(field) (smem_ptr(C.S) + c_s.field)
// GSR_LDS_MASK (developer builds): bit 0 dynamics, 1 collide, 2 solver, 3 kernels -- which files address the workspace as shared
#ifndef GSR_LDS_MASK
#define GSR_LDS_MASK 0
#endif
#define IN(field) IN_GENERIC(field)   // IN part of the record, mirrored in shared memory
#define WK(field) WK_GENERIC(field)   // workspace regions
// Shared-space variant for the self-contained dense-algebra routines (factorisations, triangular solves, rank-1 updates:
// explicit warp barriers between dependent steps): LDS / STS instead of generic LD / ST, Franka 73.8 -> 75.9 M env-steps/s.
// Applied to ALL accessors it produced wrong results in 16-20 parity tests (profiles/README.md item 13), hence the opt-in.
#ifndef GSR_NO_LDS_ROUTINES
#define WKS(field) (smem_ptr(C.S) + c_s.field)
#else
#define WKS(field) WK(field)
#endif
