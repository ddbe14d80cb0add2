// gsr_kernels.cuh -- kernels of the batched rigid-body step.
//
//   k_step<G>    : the whole substep of one environment per lane group, on chip (see gsr_types.h): bulk-load the record's
//                  IN part -> forward dynamics -> collision detection -> constraint rows + Newton/CG solve -> integration +
//                  forward kinematics -> write back.  Replaces the ~65 Quadrants launches of one reference substep
//                  (kernel_step_1, _func_constraint_force, kernel_step_2) with ONE launch.
//   k_forward<G> : forward kinematics (+ geom poses / AABBs) for the masked environments (set_qpos / set_dofs_velocity /
//                  set_state paths, RS:1545-1572) and for gsr_init_state.
//   k_reset, k_init, k_scatter, k_gather : thread-per-env housekeeping (collider.clear / constraint_solver.reset; initial
//                  state; host-buffer staging of gsr_step_host).
#pragma once
#include "gsr_group.cuh"
#include "gsr_dynamics.cuh"
#include "gsr_collide.cuh"
#include "gsr_solver.cuh"
#undef IN
#undef WK
#if GSR_LDS_MASK & 8
#define IN(field) IN_SHARED(field)
#define WK(field) WK_SHARED(field)
#else
#define IN(field) IN_GENERIC(field)
#define WK(field) WK_GENERIC(field)
#endif

#define GSR_MAX_WARPS_PER_CTA 16
// All warps of a CTA walk the phases of the step together (CTA barriers between phases): the step is ~20k instructions of
// mostly straight-line code per environment, far more than the instruction cache holds, so warps that drift apart each
// stream the whole kernel from L2 on their own (ncu: stall_no_instruction dominated before this was added).
#define PHASE_BARRIER(k) if ((c_m.phase_barriers >> (k)) & 1) __syncthreads()  // CTA-uniform condition

template <int G>
__device__ __forceinline__ bool make_ctx(Ctx<G>& C, float* state, int B, const DevModel* gm, float* smem, const int* order = nullptr) {
  constexpr int EPW = 32 / G;
  const int warp = threadIdx.x >> 5, lane32 = threadIdx.x & 31;
  const int sub = lane32 / G;
  const int slot = warp * EPW + sub;
  const int gslot = blockIdx.x * ((blockDim.x >> 5) * EPW) + slot;  // position in the launch order
  if (gslot >= B) return false;
  const int env = order ? __ldg(order + gslot) : gslot;  // cost-ordered scheduling (gsr_api.cu: reorder_envs)
  C.env = env;
  C.lane = lane32 % G;
  C.gmask = (G == 32) ? 0xffffffffu : (((1u << (G & 31)) - 1u) << (sub * G));
  C.S = smem + (size_t)slot * c_s.env_words - c_h.in_start;  // record-relative workspace offsets (SmemLayout)
  C.rec = state + (size_t)env * c_h.hot_stride;
  C.contacts = state + c_hb.contact_base + (size_t)env * c_h.contact_stride;
  C.cache = state + c_hb.cache_base + (size_t)env * c_h.cache_stride;
  C.spill = state + c_hb.spill_base + (size_t)env * c_h.spill_stride;
  C.gm = gm;
  return true;
}

// ---- record IN part <-> shared memory by the TMA (bulk async copies, one per environment, issued by one lane per warp).
// Both sides are 16-byte aligned and n_in is a multiple of 4 words by construction (gsr_model_commit).
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// Loads the IN parts of the (up to 32 / G) live environments of this warp; `bar` is this warp's mbarrier (one phase per launch).
template <int G>
__device__ __forceinline__ void bulk_load_in(float* state, int B, float* smem, uint64_t* bar, const int* order, int n_end) {
  constexpr int EPW = 32 / G;
  const int warp = threadIdx.x >> 5, lane32 = threadIdx.x & 31;
  const int env0 = (blockIdx.x * (blockDim.x >> 5) + warp) * EPW;
  const int n_live = min(EPW, B - env0);  // > 0 for a live warp
  const uint32_t bytes = (uint32_t)(n_end - c_h.in_start) * 4u;
  const uint32_t b = smem_u32(bar);
  if (lane32 == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(b), "r"(1) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(b), "r"(bytes * (uint32_t)n_live) : "memory");
    for (int e = 0; e < n_live; e++) {
      const float* src = state + (size_t)(order ? __ldg(order + env0 + e) : env0 + e) * c_h.hot_stride + c_h.in_start;
      const uint32_t dst = smem_u32(smem + (size_t)(warp * EPW + e) * c_s.env_words);
      asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                   ::"r"(dst), "l"(src), "r"(bytes), "r"(b) : "memory");
    }
  }
  __syncwarp();
  asm volatile(
      "{\n\t.reg .pred p;\n\tWAIT_%=:\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t@!p bra WAIT_%=;\n\t}"
      ::"r"(b), "r"(0) : "memory");
}
// Stores the IN parts back (smem -> global bulk copies); returns after the TMA has finished reading shared memory.
template <int G>
__device__ __forceinline__ void bulk_store_in(float* state, int B, float* smem, const int* order, int n_end) {
  constexpr int EPW = 32 / G;
  const int warp = threadIdx.x >> 5, lane32 = threadIdx.x & 31;
  const int env0 = (blockIdx.x * (blockDim.x >> 5) + warp) * EPW;
  const int n_live = min(EPW, B - env0);
  const uint32_t bytes = (uint32_t)(n_end - c_h.in_start) * 4u;
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // generic-proxy writes -> visible to the async proxy
  __syncwarp();
  if (lane32 == 0) {
    for (int e = 0; e < n_live; e++) {
      float* dst = state + (size_t)(order ? __ldg(order + env0 + e) : env0 + e) * c_h.hot_stride + c_h.in_start;
      const uint32_t src = smem_u32(smem + (size_t)(warp * EPW + e) * c_s.env_words);
      asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst), "r"(src), "r"(bytes) : "memory");
    }
    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
    asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
  }
  __syncwarp();
}

template <int G>
__device__ __forceinline__ void step_epilogue(const Ctx<G>& C, int clear_ext, bool end_fk);

// gsr_step_host: the step kernel itself reads this step's control targets from a dense (B, nD) device staging buffer and
// writes the new qpos / vel to dense (B, n) buffers -- device staging, or the device-side aliases of the caller's PAGE-LOCKED
// host buffers (posted PCIe writes that overlap the compute of the same launch) -- so no scatter / gather launch is left.
struct HostIO {
  const float* ctrl_pos;  // (B, nD) or null
  float* qpos;            // (B, nQ) or null
  float* vel;             // (B, nD) or null
};

// launch bounds of the step kernel (developer builds may override: ptxas fits the kernel into 96 registers without spilling, so
// shared memory, not registers, bounds the number of resident CTAs)
// Full-warp groups run multi-warp CTAs (one big CTA or a few per SM); small groups run one-warp CTAs, as many as the workspace
// lets an SM hold -- 17 for Franka-sized models, which needs <= 120 registers per thread.
#ifndef GSR_LB_T
#ifndef GSR_DEV_SMALL_WARPS
#define GSR_DEV_SMALL_WARPS 1
#endif
#define GSR_LB_T(G) ((G) == 32 ? GSR_MAX_WARPS_PER_CTA * 32 : 32 * GSR_DEV_SMALL_WARPS)
#ifndef GSR_DEV_SMALL_BLOCKS
#define GSR_DEV_SMALL_BLOCKS 17
#endif
#define GSR_LB_B(G) ((G) == 32 ? 1 : GSR_DEV_SMALL_BLOCKS)
#endif
template <int G, bool MPR, bool BB = false>
__global__ void __launch_bounds__(GSR_LB_T(G), GSR_LB_B(G)) k_step(float* state, int B, const DevModel* gm, int clear_ext, const int* order,
                                                                        unsigned long long* trace, HostIO io, int fk_flags) {
  extern __shared__ __align__(16) float smem[];
  Ctx<G> C;
  // order[B] tells whether the measured costs are spread enough for the permutation to pay (uniform batches keep the
  // identity order and with it the adjacency of the records a CTA touches)
  if (order && !__ldg(order + B)) order = nullptr;
  const bool live = make_ctx(C, state, B, gm, smem, order);
  // gsr_debug_trace: per environment, the time (globaltimer, ns) at the start and after every phase, and the SM
  unsigned long long tr[6] = {0ull, 0ull, 0ull, 0ull, 0ull, 0ull};
#define TRACE_MARK(k) if (trace) asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(tr[k]))
  // own busy time of the environment (SM cycles between leaving a phase barrier and finishing the phase): the cost the next
  // launches are ordered by, so that the expensive environments start first and their tail overlaps the bulk of the batch
  long long busy = 0, t_phase = clock64();
#define PHASE_END() busy += clock64() - t_phase
#define PHASE_BEGIN() t_phase = clock64()
  TRACE_MARK(0);
  // a warp is live when its first environment exists; its mbarrier sits after the workspaces
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + (size_t)(blockDim.x >> 5) * (32 / G) * c_s.env_words);
  const bool warp_live = (blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5)) * (32 / G) < B;
  // fk_flags: bit 0 = the kinematic cache of the records is stale: load the state block only and run the kinematics first; bit 1 =
  // leave a valid cache behind (kinematics of the new configuration in the epilogue, whole IN part stored).  Eager stepping is
  // (0, 1): what the reference does (FK at the end of kernel_step_2).  Lazy stepping is (1, 0): the same arithmetic per step, but the
  // cache never travels through HBM -- the host asks for it (k_forward) only when an accessor needs link poses / velocities.
  const bool start_fk = fk_flags & 1, end_fk = (fk_flags & 2) != 0;
  if (warp_live) bulk_load_in<G>(state, B, smem, bars + (threadIdx.x >> 5), order, start_fk ? c_h.n_state : c_h.n_in);
  if (start_fk && live) {  // group-uniform: fixed roots keep a user-set pose (state, not cache), everything else follows from qpos / vel
    for (int l = C.lane; l < CM(nL); l += G)
      if (GM(l_parent[l]) == -1 && GM(l_is_fixed[l])) {
        for (int k = 0; k < 3; k++) IN(links_pos)[3 * l + k] = C.rec[c_h.links_pos + 3 * l + k];
        for (int k = 0; k < 4; k++) IN(links_quat)[4 * l + k] = C.rec[c_h.links_quat + 4 * l + k];
      }
    C.sync();
    phase_kinematics(C);
  }
  if (io.ctrl_pos && live) {  // this step's control targets from the dense staging buffer
    const int nD = CM(nD);
    const float* src = io.ctrl_pos + (size_t)C.env * nD;
    for (int d = C.lane; d < nD; d += G) IN(ctrl_pos)[d] = __ldcs(src + d);
  }
  PHASE_BARRIER(0);
  PHASE_BEGIN();
  if (live) phase_dynamics(C);
  PHASE_END();
  TRACE_MARK(1);
  PHASE_BARRIER(1);
  PHASE_BEGIN();
  if (live && CM(enable_collision) && CM(nG) > 0) phase_collide<G, MPR, BB>(C);
  PHASE_END();
  TRACE_MARK(2);
  PHASE_BARRIER(2);
  PHASE_BEGIN();
  if (live) phase_constraints(C);
  PHASE_END();
  TRACE_MARK(3);
  PHASE_BARRIER(3);
  PHASE_BEGIN();
  if (live) phase_integrate(C);
  PHASE_END();
  TRACE_MARK(4);
  PHASE_BARRIER(4);
  PHASE_BEGIN();
  if (live) step_epilogue(C, clear_ext, end_fk);
  PHASE_END();
  if (live && C.lane == 0) reinterpret_cast<int*>(C.rec)[c_h.step_cost] = (int)min(busy, (long long)0x7fffffff);
  if (io.qpos || io.vel) {  // results into dense staging or straight into the caller's pinned buffers (posted PCIe writes)
    const int nQ = CM(nQ), nD = CM(nD);
    if (G < 32 && !order) {
      // Small groups in launch order: the 32 / G environments of a warp are neighbours in the dense (B, n) buffers, so the warp
      // writes ONE contiguous run per buffer (full sectors) instead of 32 / G short rows -- posted writes over PCIe are paid
      // per transaction (tools/e2e_breakdown.py: 42 us of a 332 us Franka call went into them).
      constexpr int EPW = 32 / G;
      __syncwarp();
      const int lane32 = threadIdx.x & 31, warp = threadIdx.x >> 5;
      const int env0 = (blockIdx.x * (blockDim.x >> 5) + warp) * EPW;
      const int n_live = min(EPW, B - env0);
      if (n_live > 0) {
        const float* s0 = smem + (size_t)(warp * EPW) * c_s.env_words - c_h.in_start;  // workspace of the warp's first environment
        if (io.qpos) for (int t = lane32; t < n_live * nQ; t += 32) { const int e = t / nQ, q = t - e * nQ; __stcs(io.qpos + (size_t)env0 * nQ + t, s0[(size_t)e * c_s.env_words + c_h.qpos + q]); }
        if (io.vel) for (int t = lane32; t < n_live * nD; t += 32) { const int e = t / nD, d = t - e * nD; __stcs(io.vel + (size_t)env0 * nD + t, s0[(size_t)e * c_s.env_words + c_h.vel + d]); }
      }
    } else if (live) {
      if (io.qpos) for (int q = C.lane; q < nQ; q += G) __stcs(io.qpos + (size_t)C.env * nQ + q, IN(qpos)[q]);
      if (io.vel) for (int d = C.lane; d < nD; d += G) __stcs(io.vel + (size_t)C.env * nD + d, IN(vel)[d]);
    }
  }
  __syncwarp();
  if (warp_live) bulk_store_in<G>(state, B, smem, order, end_fk ? c_h.n_in : c_h.n_state);
  if (trace && live && C.lane == 0) {
    unsigned smid;
    TRACE_MARK(5);
    asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
#pragma unroll
    for (int k = 0; k < 6; k++) trace[8 * C.env + k] = tr[k];
    trace[8 * C.env + 6] = smid;
  }
#undef TRACE_MARK
#undef PHASE_END
#undef PHASE_BEGIN
}

// Diagnostic outputs (cfg_diagnostics = 1: every step, with the reference's timing -- the mass matrix and the smooth forces of
// the configuration the step STARTED from; cfg_diagnostics = 0, the default: only on demand by k_diag for the current state).
template <int G>
__device__ __forceinline__ void write_dynamics_diagnostics(const Ctx<G>& C) {
  const int nD = CM(nD), MS = c_s.MS, lane = C.lane;
  float* rec = C.rec;
  for (int d = lane; d < nD; d += G) {
    rec[c_h.acc_smooth + d] = WK(acc_smooth)[d];
    rec[c_h.qf_smooth + d] = WK(force)[d];
    rec[c_h.qf_passive + d] = WK(qf_passive)[d];
    rec[c_h.qf_bias + d] = WK(qf_bias)[d];
    rec[c_h.mass_Dinv + d] = WK(Dinv)[d];
  }
  for (int i = 0; i < nD; i++)
    for (int j = lane; j < nD; j += G) rec[c_h.mass_mat + i * nD + j] = sym_get(WK(M), i, j);
}
template <int G>
__device__ __forceinline__ void write_geom_poses(const Ctx<G>& C) {
  const int nG = CM(nG);
  float* rec = C.rec;
  for (int g = C.lane; g < nG; g += G) {
    const int l = GM(g_link[g]);
    const f3 lp = ld3(IN(links_pos), l);
    const f4 lq = ld4(IN(links_quat), l);
    st3(rec + c_h.geoms_pos, g, lp + rotate(f3(GM(g_pos[g][0]), GM(g_pos[g][1]), GM(g_pos[g][2])), lq));
    st4(rec + c_h.geoms_quat, g, quat_compose(f4(GM(g_quat[g][0]), GM(g_quat[g][1]), GM(g_quat[g][2]), GM(g_quat[g][3])), lq));
  }
}

template <int G>
__device__ __forceinline__ void step_epilogue(const Ctx<G>& C, int clear_ext, bool end_fk) {
  // ---- outputs that live in scratch recycled by the kinematics
  const int nD = CM(nD), nL = CM(nL), MS = c_s.MS, lane = C.lane;
  float* rec = C.rec;
  for (int d = lane; d < nD; d += G) {
    rec[c_h.acc + d] = WK(acc)[d];
    rec[c_h.force + d] = WK(force)[d] + WK(qfrc_c)[d];
    rec[c_h.qf_constraint + d] = WK(qfrc_c)[d];
  }
  for (int l = lane; l < nL; l += G) {
    st3(rec + c_h.cacc_lin, l, ld3(WK(cacc), 2 * l));
    st3(rec + c_h.cacc_ang, l, ld3(WK(cacc), 2 * l + 1));
  }
  if (CM(diagnostics)) write_dynamics_diagnostics(C);
  C.sync();
  if (end_fk) phase_kinematics(C);
  if (CM(diagnostics)) write_geom_poses(C);  // geoms_state.pos / quat of the post-integration configuration (FK:828-864)
  if (clear_ext && C.SI()[c_h.has_ext_force]) {  // RS:2603-2612 clear_external_force
    for (int i = lane; i < 3 * nL; i += G) { rec[c_h.cfrc_applied_vel + i] = 0.f; rec[c_h.cfrc_applied_ang + i] = 0.f; }
    C.sync();
    if (lane == 0) C.SI()[c_h.has_ext_force] = 0;
  }
  C.sync();
}

template <int G>
__global__ void __launch_bounds__(GSR_MAX_WARPS_PER_CTA * 32) k_forward(float* state, int B, const DevModel* gm, const uint8_t* mask, int dirty_only) {
  extern __shared__ __align__(16) float smem[];
  Ctx<G> C;
  if (!make_ctx(C, state, B, gm, smem)) return;
  if (mask && !mask[C.env]) return;
  // deferred kinematics (set_qpos(..., skip_forward=True), set_state(..., partial=True)): only the environments flagged in their record
  if (dirty_only && !reinterpret_cast<const int*>(C.rec)[c_h.fk_dirty]) return;
  {  // masked environments are few: plain vector copies
    const float4* src = reinterpret_cast<const float4*>(C.rec);
    float4* dst = reinterpret_cast<float4*>(C.S);
    for (int i = c_h.in_start / 4 + C.lane; i < c_h.n_in / 4; i += G) dst[i] = src[i];
    C.sync();
  }
  phase_kinematics(C);
  phase_geoms(C);
  if (C.lane == 0) reinterpret_cast<int*>(C.rec)[c_h.fk_dirty] = 0;  // PRE block: lives in the record only
  C.sync();
  {
    float4* dst = reinterpret_cast<float4*>(C.rec);
    const float4* src = reinterpret_cast<const float4*>(C.S);
    for (int i = c_h.in_start / 4 + C.lane; i < c_h.n_in / 4; i += G) dst[i] = src[i];
  }
}

// RigidSolver.set_state (rigid_solver.py:1458-1579) for the masked environments in ONE launch: collider.reset / clear and
// constraint_solver.reset / clear, errno = 0, state from dense (B, n) buffers (qpos, dofs_vel, dofs_acc, links_pos / quat of the
// fixed root links, i_pos_shift, mass_shift, friction_ratio), ctrl_force = 0 / ctrl_mode = FORCE, external wrenches cleared,
// then -- unless `partial` -- the forward kinematics of the new configuration; with `partial` the kinematics are left to the
// next step (fk_dirty), as the reference does.
struct StateIO {
  const float *qpos, *vel, *acc, *links_pos, *links_quat, *i_pos_shift, *mass_shift, *friction_ratio;
};
template <int G>
__global__ void __launch_bounds__(GSR_MAX_WARPS_PER_CTA * 32) k_set_state(float* state, int B, const DevModel* gm, const uint8_t* mask, StateIO io, int partial) {
  extern __shared__ __align__(16) float smem[];
  Ctx<G> C;
  if (!make_ctx(C, state, B, gm, smem)) return;
  if (mask && !mask[C.env]) return;
  const int nD = CM(nD), nQ = CM(nQ), nL = CM(nL), nG = CM(nG), lane = C.lane, env = C.env;
  float* rec = C.rec;
  int* reci = reinterpret_cast<int*>(rec);
  // collider.reset (warm-start cache) / constraint_solver.reset (warm start); clear additionally drops contacts and row counts
  for (int d = lane; d < nD; d += G) rec[c_h.qacc_ws + d] = 0.f;
  int* cv = reinterpret_cast<int*>(C.cache) + c_h.cache_valid;
  for (int w = lane; w < CM(n_cache_words); w += G) cv[w] = 0;
  if (!partial) {
    const int nc = reci[c_h.n_contacts];
    for (int i = lane; i < nc * GSR_CONTACT_WORDS; i += G) {
      const int k = i % GSR_CONTACT_WORDS;
      if (k < 4) reinterpret_cast<int*>(C.contacts)[i] = -1; else C.contacts[i] = 0.f;
    }
  }
  bool any_shift = false;
  for (int q = lane; q < nQ; q += G) if (io.qpos) rec[c_h.qpos + q] = io.qpos[(size_t)env * nQ + q];
  for (int d = lane; d < nD; d += G) {
    if (io.vel) rec[c_h.vel + d] = io.vel[(size_t)env * nD + d];
    if (io.acc) rec[c_h.acc + d] = io.acc[(size_t)env * nD + d];
    rec[c_h.ctrl_force + d] = 0.f;
    reci[c_h.ctrl_mode + d] = CTRL_FORCE;
  }
  for (int l = lane; l < nL; l += G) {
    if (GM(l_parent[l]) == -1 && GM(l_is_fixed[l])) {  // only fixed roots keep a user-set pose; every other pose follows from qpos
      if (io.links_pos) for (int k = 0; k < 3; k++) rec[c_h.links_pos + 3 * l + k] = io.links_pos[((size_t)env * nL + l) * 3 + k];
      if (io.links_quat) for (int k = 0; k < 4; k++) rec[c_h.links_quat + 4 * l + k] = io.links_quat[((size_t)env * nL + l) * 4 + k];
    }
    for (int k = 0; k < 3; k++) {
      if (io.i_pos_shift) { const float v = io.i_pos_shift[((size_t)env * nL + l) * 3 + k]; rec[c_h.com_shift + 3 * l + k] = v; any_shift |= v != 0.f; }
      rec[c_h.cfrc_applied_vel + 3 * l + k] = 0.f; rec[c_h.cfrc_applied_ang + 3 * l + k] = 0.f;
    }
    if (io.mass_shift) { const float v = io.mass_shift[(size_t)env * nL + l]; rec[c_h.mass_shift + l] = v; any_shift |= v != 0.f; }
  }
  for (int g = lane; g < nG; g += G) if (io.friction_ratio) rec[c_h.friction_ratio + g] = io.friction_ratio[(size_t)env * nG + g];
  any_shift = C.any(any_shift);
  if (lane == 0) {
    reci[c_h.is_warmstart] = 0; reci[c_h.errno_] = 0; reci[c_h.has_ext_force] = 0;
    if (io.i_pos_shift || io.mass_shift) reci[c_h.has_shift] = any_shift ? 1 : 0;
    if (!partial) {
      reci[c_h.n_contacts] = 0; reci[c_h.n_broad] = 0;
      reci[c_h.n_constraints] = 0; reci[c_h.n_constraints_equality] = 0; reci[c_h.n_constraints_frictionloss] = 0;
    }
    reci[c_h.fk_dirty] = partial ? 1 : 0;
  }
  __threadfence_block();
  C.sync();
  if (partial) return;
  {
    const float4* src = reinterpret_cast<const float4*>(rec);
    float4* dst = reinterpret_cast<float4*>(C.S);
    for (int i = c_h.in_start / 4 + lane; i < c_h.n_in / 4; i += G) dst[i] = src[i];
    C.sync();
  }
  phase_kinematics(C);
  phase_geoms(C);
  C.sync();
  {
    float4* dst = reinterpret_cast<float4*>(rec);
    const float4* src = reinterpret_cast<const float4*>(C.S);
    for (int i = c_h.in_start / 4 + lane; i < c_h.n_in / 4; i += G) dst[i] = src[i];
  }
}

// Several fields of every environment's HOT record into dense (B, n) buffers in ONE launch: kernel_get_state
// (abd/accessor.py) and the observation getters of an RL loop (qpos, dofs position / velocity, link poses and velocities ...).
#define GSR_MAX_GATHER 16
struct GatherIO {
  int n;                        // number of segments
  int off[GSR_MAX_GATHER];      // word offset inside the HOT record
  int width[GSR_MAX_GATHER];    // words per environment
  int start[GSR_MAX_GATHER + 1];  // prefix sum of the widths
  float* dst[GSR_MAX_GATHER];   // dense (B, width) destinations
};
__global__ void k_gather_fields(const float* state, int B, GatherIO io) {
  const int total = io.start[io.n];
  const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= (long long)B * total) return;
  const int env = (int)(t / total), w = (int)(t - (long long)env * total);
  int s = 0;
#pragma unroll
  for (int k = 1; k < GSR_MAX_GATHER; k++) if (k < io.n && w >= io.start[k]) s = k;
  const int i = w - io.start[s];
  io.dst[s][(size_t)env * io.width[s] + i] = state[(size_t)env * c_h.hot_stride + io.off[s] + i];
}

// On-demand diagnostics of the CURRENT state (get_mass_mat, geoms_state.pos / quat, qf_* with cfg_diagnostics = 0): smooth
// dynamics from the kinematic cache of the record, outputs into the diagnostic fields; the state itself is not touched.
template <int G>
__global__ void __launch_bounds__(GSR_MAX_WARPS_PER_CTA * 32) k_diag(float* state, int B, const DevModel* gm) {
  extern __shared__ __align__(16) float smem[];
  Ctx<G> C;
  if (!make_ctx(C, state, B, gm, smem)) return;
  {
    const float4* src = reinterpret_cast<const float4*>(C.rec);
    float4* dst = reinterpret_cast<float4*>(C.S);
    for (int i = c_h.in_start / 4 + C.lane; i < c_h.n_in / 4; i += G) dst[i] = src[i];
    C.sync();
  }
  phase_dynamics(C);
  C.sync();
  write_dynamics_diagnostics(C);
  write_geom_poses(C);
}

// collider.clear + constraint_solver.reset for masked envs (collider.py:563-607, solver.py:114-179)
__global__ void k_reset(float* state, int B, const uint8_t* mask) {
  const int env = blockIdx.x * blockDim.x + threadIdx.x;
  if (env >= B) return;
  if (mask && !mask[env]) return;
  float* rec = state + (size_t)env * c_h.hot_stride;
  int* reci = reinterpret_cast<int*>(rec);
  for (int d = 0; d < c_m.nD; d++) rec[c_h.qacc_ws + d] = 0.f;
  reci[c_h.is_warmstart] = 0;
  int* cv = reinterpret_cast<int*>(state + c_hb.cache_base + (size_t)env * c_h.cache_stride) + c_h.cache_valid;
  for (int w = 0; w < c_m.n_cache_words; w++) cv[w] = 0;
  float* ct = state + c_hb.contact_base + (size_t)env * c_h.contact_stride;
  const int nc = reci[c_h.n_contacts];
  for (int c = 0; c < nc; c++) {
    float* r = ct + (size_t)c * GSR_CONTACT_WORDS;
    for (int k = 0; k < 4; k++) reinterpret_cast<int*>(r)[k] = -1;
    for (int k = 4; k < GSR_CONTACT_WORDS; k++) r[k] = 0.f;
  }
  reci[c_h.n_contacts] = 0; reci[c_h.n_broad] = 0;
  reci[c_h.n_constraints] = 0; reci[c_h.n_constraints_equality] = 0; reci[c_h.n_constraints_frictionloss] = 0;
  reci[c_h.errno_] = 0;
}

// RigidSolver.build's initial state: qpos0, FORCE control mode (abd/misc.py:170-173), unit friction ratios, model link poses
__global__ void k_init(float* state, int B) {
  const int env = blockIdx.x * blockDim.x + threadIdx.x;
  if (env >= B) return;
  float* rec = state + (size_t)env * c_h.hot_stride;
  int* reci = reinterpret_cast<int*>(rec);
  for (int q = 0; q < c_m.nQ; q++) rec[c_h.qpos + q] = c_m.qpos0[q];
  for (int d = 0; d < c_m.nD; d++) reci[c_h.ctrl_mode + d] = CTRL_FORCE;
  for (int g = 0; g < c_m.nG; g++) rec[c_h.friction_ratio + g] = 1.0f;
  if (c_m.batch_dofs)  // RigidOptions.batch_dofs_info: every environment starts from the model's dof parameters
    for (int d = 0; d < c_m.nD; d++) {
      rec[c_h.bd_kp + d] = c_m.d_kp[d]; rec[c_h.bd_armature + d] = c_m.d_armature[d]; rec[c_h.bd_damping + d] = c_m.d_damping[d];
      rec[c_h.bd_stiffness + d] = c_m.d_stiffness[d]; rec[c_h.bd_frictionloss + d] = c_m.d_frictionloss[d]; rec[c_h.bd_invweight + d] = c_m.d_invweight[d];
      for (int k = 0; k < 3; k++) rec[c_h.bd_bias + 3 * d + k] = c_m.d_act_bias[d][k];
      for (int k = 0; k < 2; k++) { rec[c_h.bd_limit + 2 * d + k] = c_m.d_limit[d][k]; rec[c_h.bd_force_range + 2 * d + k] = c_m.d_force_range[d][k]; }
    }
  for (int l = 0; l < c_m.nL; l++) {
    for (int k = 0; k < 3; k++) rec[c_h.links_pos + 3 * l + k] = c_m.l_pos[l][k];
    for (int k = 0; k < 4; k++) rec[c_h.links_quat + 4 * l + k] = c_m.l_quat[l][k];
  }
}

// gsr_step_host staging: (B, n) env-major host-order array <-> a field of the records
__global__ void k_scatter(float* state, int B, int field_off, int n, const float* src) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= B * n) return;
  const int env = t / n, i = t - env * n;
  state[(size_t)env * c_h.hot_stride + field_off + i] = src[t];
}
__global__ void k_gather(const float* state, int B, int field_off, int n, float* dst) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= B * n) return;
  const int env = t / n, i = t - env * n;
  dst[t] = state[(size_t)env * c_h.hot_stride + field_off + i];
}
