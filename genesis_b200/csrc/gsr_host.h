// gsr_host.h -- host-only part of libgsrigid: the named model arrays, their translation into the device-side model
// (DevModel), and the state / workspace layouts.  Plain C++ (no CUDA runtime calls) so that the test-only host emulation of
// the lane kernel (tests/lane_emu) shares the exact model filling and record layout of the product library.
#pragma once
#include <cstdio>
#include <cstring>
#include <map>
#include <string>
#include <vector>

#include "../../include/gsrigid.h"
#include "gsr_types.h"

struct GsrFieldInfo {
  std::string name;
  int is_int;
  size_t seg_base, env_stride, offset, items, comps, item_stride;
};

// what gsr_model_set_* collects
struct GsrSpec {
  std::map<std::string, std::vector<float>> f;
  std::map<std::string, std::vector<int>> i;
  std::map<std::string, double> s;
  std::string err;
  DevModel hm;     // host copy of the model
  HotLayout lay;
  SmemLayout sl;
  LaneLayout ll;
  std::vector<GsrFieldInfo> fields;
  std::vector<unsigned short> fac_tasks;  // DevModel::fac_tasks (host copy)
  size_t n_words = 0;
  int group = 32;  // lanes per environment (1 = lane-per-environment kernel)
  int warps_per_cta = 0;
  double S(const char* k, double d) const { auto it = s.find(k); return it == s.end() ? d : it->second; }
};

static inline int gsr_fail(GsrSpec* h, int code, const std::string& msg) {
  if (h) h->err = msg;
  return code;
}
static inline size_t gsr_round4(size_t x) { return (x + 3) & ~(size_t)3; }

static inline int gsr_fill_model(GsrSpec* h) {
  DevModel& m = h->hm;
  const float* keep_verts = m.verts; const float* keep_support = m.support_v; const int* keep_svid = m.support_vid; const int* keep_pairs = m.pairs; const int* keep_pidx = m.pair_idx;
  const unsigned short* keep_fac = m.fac_tasks;
  const float* keep_terrain = m.terrain_hf;
  memset(&m, 0, sizeof(m));
  m.terrain_hf = keep_terrain;
  m.verts = keep_verts; m.support_v = keep_support; m.support_vid = keep_svid; m.pairs = keep_pairs; m.pair_idx = keep_pidx; m.fac_tasks = keep_fac;
  auto S = [&](const char* k, double d) { auto it = h->s.find(k); return it == h->s.end() ? d : it->second; };
  m.nL = (int)S("n_links", 0); m.nJ = (int)S("n_joints", 0); m.nD = (int)S("n_dofs", 0); m.nQ = (int)S("n_qs", 0);
  m.nG = (int)S("n_geoms", 0); m.nE = (int)S("n_entities", 0); m.nEq = (int)S("n_equalities", 0); m.nP = (int)S("n_pairs", 0);
  m.nV = (int)S("n_verts", 0);
  if (m.nL > GSR_MAX_LINKS || m.nJ > GSR_MAX_JOINTS || m.nD > GSR_MAX_DOFS || m.nQ > GSR_MAX_QS || m.nG > GSR_MAX_GEOMS ||
      m.nE > GSR_MAX_ENTS || m.nEq > GSR_MAX_EQ) {
    char buf[256];
    snprintf(buf, sizeof(buf), "model exceeds compiled capacities (links %d/%d joints %d/%d dofs %d/%d qs %d/%d geoms %d/%d entities %d/%d eq %d/%d)",
             m.nL, GSR_MAX_LINKS, m.nJ, GSR_MAX_JOINTS, m.nD, GSR_MAX_DOFS, m.nQ, GSR_MAX_QS, m.nG, GSR_MAX_GEOMS, m.nE, GSR_MAX_ENTS, m.nEq, GSR_MAX_EQ);
    return gsr_fail(h, GSR_ELIMIT, buf);
  }
  m.max_contact_pairs = (int)S("max_contact_pairs", 0); m.max_broad_pairs = (int)S("max_broad_pairs", 0);
  m.n_contacts_per_pair = (int)S("n_contacts_per_pair", 5); m.len_constraints = (int)S("len_constraints", 0);
  m.iterations = (int)S("cfg_iterations", 50); m.ls_iterations = (int)S("cfg_ls_iterations", 50);
  m.integrator = (int)S("cfg_integrator", INTEGRATOR_APPROX_IMPLICITFAST); m.solver_type = (int)S("cfg_solver_type", SOLVER_NEWTON);
  m.gjk_fallback = (int)S("cfg_gjk_fallback", 1);
  m.phase_barriers = (int)S("cfg_phase_barriers", 31);
  m.noslip_iterations = (int)S("cfg_noslip_iterations", 0); m.noslip_tolerance = (float)S("cfg_noslip_tolerance", 1e-6);
  m.enable_collision = (int)S("cfg_enable_collision", 1); m.enable_joint_limit = (int)S("cfg_enable_joint_limit", 1);
  m.enable_multi_contact = (int)S("cfg_enable_multi_contact", 1); m.sort_contacts = (int)S("cfg_sort_contacts", 1);
  m.meaninertia = (float)S("meaninertia", 1.0); m.dt = (float)S("dt", 0.01); m.eps = (float)S("cfg_eps", 1.1920929e-07);
  m.tolerance = (float)S("cfg_tolerance", 1e-6); m.ls_tolerance = (float)S("cfg_ls_tolerance", 1e-2);
  m.mc_perturbation = 1e-2f; m.mc_tolerance = 1e-2f;  // collider.py:85-86
  m.gravity[0] = (float)S("gravity_x", 0.0); m.gravity[1] = (float)S("gravity_y", 0.0); m.gravity[2] = (float)S("gravity_z", -9.81);
  m.ccd_eps = 1e-9f; m.ccd_tolerance = 1e-6f; m.ccd_iterations = 50;  // mpr.py:13-19
  m.n_cache_words = (m.nP + 31) / 32 > 0 ? (m.nP + 31) / 32 : 1;
  m.diagnostics = (int)S("cfg_diagnostics", 1);
  m.batch_dofs = (int)S("cfg_batch_dofs_info", 0);
  m.disable_constraint = (int)S("cfg_disable_constraint", 0);

  bool missing = false;
  std::string missing_name;
  auto FV = [&](const char* k, size_t n) -> const float* {
    auto it = h->f.find(k);
    if (it == h->f.end() || it->second.size() < n) { if (n) { missing = true; missing_name = k; } static float z[1] = {0}; return z; }
    return it->second.data();
  };
  auto IV = [&](const char* k, size_t n) -> const int* {
    auto it = h->i.find(k);
    if (it == h->i.end() || it->second.size() < n) { if (n) { missing = true; missing_name = k; } static int z[1] = {0}; return z; }
    return it->second.data();
  };
#define CPI(dst, key, n) { const int* p_ = IV(key, n); if (!missing) memcpy(dst, p_, sizeof(int) * (n)); }
#define CPF(dst, key, n) { const float* p_ = FV(key, n); if (!missing) memcpy(dst, p_, sizeof(float) * (n)); }
  const int nL = m.nL, nJ = m.nJ, nD = m.nD, nQ = m.nQ, nG = m.nG, nE = m.nE, nEq = m.nEq;
  CPI(m.l_parent, "links_parent", nL) CPI(m.l_root, "links_root", nL) CPI(m.l_entity, "links_entity", nL)
  CPI(m.l_is_fixed, "links_is_fixed", nL) CPI(m.l_joint_start, "links_joint_start", nL) CPI(m.l_joint_end, "links_joint_end", nL)
  CPI(m.l_dof_start, "links_dof_start", nL) CPI(m.l_dof_end, "links_dof_end", nL) CPI(m.l_q_start, "links_q_start", nL)
  CPI(m.l_q_end, "links_q_end", nL)
  CPF(m.l_pos, "links_pos", 3 * nL) CPF(m.l_quat, "links_quat", 4 * nL) CPF(m.l_ipos, "links_inertial_pos", 3 * nL)
  CPF(m.l_iquat, "links_inertial_quat", 4 * nL) CPF(m.l_inertia, "links_inertial_i", 9 * nL) CPF(m.l_mass, "links_inertial_mass", nL)
  { const float* p = FV("links_invweight", 2 * nL); if (!missing) for (int l = 0; l < nL; l++) { m.l_invweight[l] = p[2 * l]; m.l_invweight_rot[l] = p[2 * l + 1]; } }
  CPI(m.j_type, "joints_type", nJ) CPI(m.j_q_start, "joints_q_start", nJ) CPI(m.j_dof_start, "joints_dof_start", nJ)
  CPI(m.j_n_dofs, "joints_n_dofs", nJ) CPF(m.j_pos, "joints_pos", 3 * nJ) CPF(m.j_sol, "joints_sol_params", 7 * nJ)
  CPI(m.d_link, "dofs_link", nD) CPI(m.d_entity, "dofs_entity", nD) CPI(m.d_joint, "dofs_joint", nD)
  CPF(m.d_motion_ang, "dofs_motion_ang", 3 * nD) CPF(m.d_motion_vel, "dofs_motion_vel", 3 * nD) CPF(m.d_limit, "dofs_limit", 2 * nD)
  CPF(m.d_stiffness, "dofs_stiffness", nD) CPF(m.d_damping, "dofs_damping", nD) CPF(m.d_armature, "dofs_armature", nD)
  CPF(m.d_frictionloss, "dofs_frictionloss", nD) CPF(m.d_kp, "dofs_kp", nD) CPF(m.d_act_bias, "dofs_act_bias", 3 * nD)
  CPF(m.d_force_range, "dofs_force_range", 2 * nD) CPF(m.d_invweight, "dofs_invweight", nD) CPF(m.qpos0, "qpos0", nQ)
  CPI(m.e_link_start, "entities_link_start", nE) CPI(m.e_link_end, "entities_link_end", nE) CPI(m.e_dof_start, "entities_dof_start", nE)
  CPI(m.e_dof_end, "entities_dof_end", nE) CPI(m.e_geom_start, "entities_geom_start", nE) CPI(m.e_geom_end, "entities_geom_end", nE)
  CPF(m.e_gravcomp, "entities_gravity_compensation", nE)
  CPI(m.g_link, "geoms_link", nG) CPI(m.g_type, "geoms_type", nG) CPI(m.g_is_fixed, "geoms_is_fixed", nG)
  CPI(m.g_vert_start, "geoms_vert_start", nG) CPI(m.g_vert_end, "geoms_vert_end", nG) CPI(m.g_support_start, "geoms_support_start", nG)
  CPF(m.g_pos, "geoms_pos", 3 * nG) CPF(m.g_quat, "geoms_quat", 4 * nG) CPF(m.g_friction, "geoms_friction", nG)
  CPF(m.g_sol, "geoms_sol_params", 7 * nG) CPF(m.g_center, "geoms_center", 3 * nG)
  {
    const float* p = FV("geoms_data", 7 * nG);
    const float* a = FV("geoms_init_aabb", 24 * nG);
    if (!missing)
      for (int g = 0; g < nG; g++)
        for (int k = 0; k < 3; k++) { m.g_data[g][k] = p[7 * g + k]; m.g_aabb_lo[g][k] = a[24 * g + k]; m.g_aabb_hi[g][k] = a[24 * g + 21 + k]; }
  }
  CPI(m.eq_type, "eq_type", nEq) CPI(m.eq_obj1, "eq_obj1", nEq) CPI(m.eq_obj2, "eq_obj2", nEq)
  CPF(m.eq_data, "eq_data", 11 * nEq) CPF(m.eq_sol, "eq_sol_params", 7 * nEq)
#undef CPI
#undef CPF
  if (missing) return gsr_fail(h, GSR_EINVAL, "missing or short model array: " + missing_name);

  for (int i = 0, t = 0; i < GSR_MAX_DOFS; i++)
    for (int j = 0; j <= i; j++) m.tri[t++] = (unsigned short)((i << 8) | j);
  // derived topology tables: ancestor masks (links and dofs), subtree mass per root, tree depth
  m.max_depth = 1;
  for (int l = 0; l < nL; l++) {
    unsigned long long am = 0ull, dm = 0ull;
    int a = l, depth = 0;
    while (a != -1) {
      am |= 1ull << a;
      for (int d = m.l_dof_start[a]; d < m.l_dof_end[a]; d++) dm |= 1ull << d;
      a = m.l_parent[a];
      depth++;
    }
    m.l_ancmask[l] = am; m.l_dofmask[l] = dm; m.l_depth[l] = depth;
    if (depth > m.max_depth) m.max_depth = depth;
    m.l_root_mass[l] = 0.f;
  }
  for (int l = 0; l < nL; l++) m.l_root_mass[m.l_root[l]] += m.l_mass[l];
  // sparse L^T D L elimination: for pivot i only the pairs (j, k), k <= j < i, of its ancestor dofs
  h->fac_tasks.clear();
  for (int i = 0; i < nD; i++) {
    m.fac_start[i] = (int)h->fac_tasks.size();
    const unsigned long long anc = m.l_dofmask[m.d_link[i]];
    for (int j = 0; j < i; j++) {
      if (!((anc >> j) & 1ull)) continue;
      for (int k = 0; k <= j; k++)
        if ((anc >> k) & 1ull) h->fac_tasks.push_back((unsigned short)((j << 8) | k));
    }
  }
  m.fac_start[nD] = (int)h->fac_tasks.size();
  // lane kernel tables: compact slot of every tree root, composite (subtree) mass accumulated leaf-to-root like FD:382-388
  m.n_roots = 0;
  for (int l = 0; l < nL; l++)
    if (m.l_parent[l] < 0) m.l_rootslot[l] = m.n_roots++;
  for (int l = 0; l < nL; l++) { m.l_rootslot[l] = m.l_rootslot[m.l_root[l]]; m.l_submass[l] = m.l_mass[l]; }
  for (int l = nL - 1; l >= 0; l--)
    if (m.l_parent[l] >= 0) m.l_submass[m.l_parent[l]] += m.l_submass[l];
  for (int l = 0; l < nL; l++) {
    m.l_descmask[l] = 0ull;
    for (int d = 0; d < nL; d++)
      if ((m.l_ancmask[d] >> l) & 1ull) m.l_descmask[l] |= 1ull << d;
  }

  // has_cc: any valid pair that is not plane-box (collider.py:449-459) -> contacts are x-sorted (collider.py:804-808)
  {  // height-field terrain tables (model arrays terrain_rc / terrain_scale / terrain_xyz_maxmin; the heights go to the device in gsr_model_commit)
    auto rc = h->i.find("terrain_rc");
    auto sc = h->f.find("terrain_scale");
    auto mm = h->f.find("terrain_xyz_maxmin");
    if (rc != h->i.end() && rc->second.size() == 2 && sc != h->f.end() && sc->second.size() == 2 && mm != h->f.end() && mm->second.size() == 6) {
      m.terrain_rows = rc->second[0]; m.terrain_cols = rc->second[1]; m.terrain_hs = sc->second[0];
      for (int k = 0; k < 6; k++) m.terrain_maxmin[k] = mm->second[k];
    }
  }
  m.has_cc = 0;
  m.has_mpr = 0;
  m.has_box_box = 0;
  const bool box_box_detection = h->S("cfg_box_box_detection", 0) != 0;  // options/solvers.py:511
  m.box_box_detection = box_box_detection ? 1 : 0;
  m.use_gjk = h->S("cfg_use_gjk_collision", 0) != 0 ? 1 : 0;
  {
    const int* pr = IV("pairs", 2 * (size_t)m.nP);
    for (int p = 0; p < m.nP; p++) {
      int ta = m.g_type[pr[2 * p]], tb = m.g_type[pr[2 * p + 1]];
      if (ta > tb) { int t = ta; ta = tb; tb = t; }
      const bool box_box = box_box_detection && ta == GEOM_BOX && tb == GEOM_BOX;  // specialised like plane-box (collider.py:456-459)
      if (box_box || tb == GEOM_TERRAIN) { m.has_box_box = 1; continue; }  // terrain: not convex, its own narrow-phase pass
      if (!(ta == GEOM_PLANE && tb == GEOM_BOX)) m.has_cc = 1;
      const bool analytic = (ta == GEOM_CAPSULE && tb == GEOM_CAPSULE) || (ta == GEOM_SPHERE && tb == GEOM_CAPSULE) || (ta == GEOM_CAPSULE && tb == GEOM_SPHERE);
      if (ta != GEOM_PLANE && tb != GEOM_PLANE && !analytic) { m.has_mpr = 1; if (m.use_gjk) m.has_box_box = 1; }  // GJK-first lives in the full variant
    }
  }

  return GSR_OK;
}


// ---- state record layout (HBM) and per-environment workspace layouts (shared memory) -------------------------------------
static inline int gsr_compute_layouts(GsrSpec* h, int n_envs) {
  DevModel& m = h->hm;
  // ---- lanes per environment: a full warp unless the batch is large enough to fill the machine with smaller groups
  auto S = [&](const char* k, double d) { return h->S(k, d); };
  int G = (int)S("cfg_group_size", 0);
  if (G != 1 && G != 8 && G != 16 && G != 32) G = (m.nD <= 12 && m.nL <= 16) ? 8 : 32;  // measured on B200: profiles/README.md
  h->group = G;
  h->warps_per_cta = (int)S("cfg_warps_per_cta", 0);

  // ---- HOT record layout
  const int nL = m.nL, nD = m.nD, nQ = m.nQ, nG = m.nG;
  const int nBD = m.batch_dofs ? nD : 0;
  HotLayout& L = h->lay;
  memset(&L, 0, sizeof(L));
  h->fields.clear();
  size_t off = 0;
#define X(name, is_int, items, comps)                                                                   \
  {                                                                                                     \
    size_t it_ = (size_t)(items), cp_ = (size_t)(comps);                                                \
    L.name = (int)off;                                                                                  \
    h->fields.push_back(GsrFieldInfo{#name, is_int, 0, 0, off, it_, cp_, cp_});                            \
    off += (it_ * cp_ > 0 ? it_ * cp_ : 1);                                                             \
  }
  GSR_HOT_PRE_FIELDS(X)
  off = gsr_round4(off);
  L.in_start = (int)off;
  GSR_HOT_STATE_FIELDS(X)
  off = gsr_round4(off);
  L.n_state = (int)off;
  GSR_HOT_CACHE_FIELDS(X)
  off = gsr_round4(off);
  L.n_in = (int)off;
  L.kin_dead = L.cinr_I;
  GSR_HOT_OUT_FIELDS(X)
#undef X
  L.hot_stride = (int)gsr_round4(off);
  const int nCP = m.max_contact_pairs > 0 ? m.max_contact_pairs : 1;
  const int lenC = m.len_constraints > 0 ? m.len_constraints : 1;
  const int MS = nD | 1, JS = nD | 1;
  L.contact_stride = nCP * GSR_CONTACT_WORDS;
  L.cache_valid = (int)gsr_round4(3 * (size_t)(m.nP > 0 ? m.nP : 1));
  L.cache_stride = (int)gsr_round4(L.cache_valid + m.n_cache_words);
  size_t spill = (size_t)lenC * JS + 8 * (size_t)lenC;
  L.spill_minvjt = (int)spill;
  if (m.noslip_iterations > 0) spill += (size_t)lenC * JS;  // M^-1 J^T rows of the noslip post-solver
  // contact staging records + the results of the perturbed multi-contact detections of one batch (32 lanes x 4 x 8 words)
  if (spill < (size_t)nCP * GSR_CONTACT_WORDS + 1024) spill = (size_t)nCP * GSR_CONTACT_WORDS + 1024;
  L.spill_stride = (int)gsr_round4(spill);
  const size_t B = (size_t)n_envs;
  L.contact_base = (long long)(B * L.hot_stride);
  L.cache_base = L.contact_base + (long long)(B * L.contact_stride);
  L.spill_base = L.cache_base + (long long)(B * L.cache_stride);
  h->n_words = (size_t)L.spill_base + B * L.spill_stride;
  for (auto& f : h->fields) { f.seg_base = 0; f.env_stride = L.hot_stride; }
  {  // contact record fields (AoS) and the warm-start cache
    auto add = [&](const char* name, int is_int, size_t base, size_t stride, size_t o, size_t items, size_t comps, size_t istride) {
      h->fields.push_back(GsrFieldInfo{name, is_int, base, stride, o, items, comps, istride});
    };
    const size_t cb = (size_t)L.contact_base, cs = L.contact_stride, W = GSR_CONTACT_WORDS;
    add("contact_geom_a", 1, cb, cs, CT_GA, nCP, 1, W); add("contact_geom_b", 1, cb, cs, CT_GB, nCP, 1, W);
    add("contact_link_a", 1, cb, cs, CT_LA, nCP, 1, W); add("contact_link_b", 1, cb, cs, CT_LB, nCP, 1, W);
    add("contact_penetration", 0, cb, cs, CT_PEN, nCP, 1, W); add("contact_friction", 0, cb, cs, CT_MU, nCP, 1, W);
    add("contact_pos", 0, cb, cs, CT_POS, nCP, 3, W); add("contact_normal", 0, cb, cs, CT_NRM, nCP, 3, W);
    add("contact_force", 0, cb, cs, CT_FRC, nCP, 3, W);
    add("cache_normal", 0, (size_t)L.cache_base, L.cache_stride, 0, m.nP > 0 ? m.nP : 1, 3, 3);
    add("cache_valid", 1, (size_t)L.cache_base, L.cache_stride, L.cache_valid, m.n_cache_words, 1, 1);
  }

  // ---- shared-memory workspace layout (see gsr_types.h): [IN mirror | overlay pool | always-live block]
  SmemLayout& s = h->sl;
  memset(&s, 0, sizeof(s));
  s.MS = MS; s.JS = JS;
  int NCS = (int)S("cfg_smem_rows", nD > 12 ? 32 : 12);  // rows kept on chip (more => SPILL record): 12 lets 17 Franka CTAs fit an SM,
  // 32 covers 99.9 % of the Go2 env-steps and leaves twice the L1 for the model tables (profiles/README.md, r1c sweeps)
  if (NCS > lenC) NCS = lenC;
  if (NCS < 4) NCS = 4;
  s.NCS = NCS;
  s.NCC = NCS / 4 > 0 ? NCS / 4 : 1;
  // scratch of the dynamics / collision / integrate+FK phases starts after the IN mirror.  Dynamics: the composite inertias and
  // f_d = crb * cdof are dead once the mass matrix is assembled, so the bias-force scratch (cdd, cfrc) overlays them; the smooth
  // force terms that only the diagnostics read after the phase (qf_passive, qf_bias) and D^-1 (needed after the phase only by the
  // CG solver, noslip, the implicit integrators and the diagnostics) follow them inside the pool when nothing reads them later.
  const bool newton = m.solver_type == SOLVER_NEWTON;
  const bool dyn_terms_live = m.diagnostics != 0;
  const bool dinv_live = m.diagnostics != 0 || !newton || m.noslip_iterations > 0 || m.integrator != INTEGRATOR_APPROX_IMPLICITFAST;
  int x = L.n_in, xmax = x;
  const bool fdof_on_L = 6 * nD <= nD * MS;  // f_d = crb * cdof waits in L's slot: the factorisation copies M there afterwards
  s.crb = x; x += 10 * nL;
  if (!fdof_on_L) { s.fdof = x; x += 6 * nD; }
  if (x > xmax) xmax = x;
  x = L.n_in;
  s.cdd = x; x += 6 * nL; s.cfrc = x; x += 6 * nL;
  if (!dyn_terms_live) { s.qf_passive = x; x += nD; s.qf_bias = x; x += nD; }
  if (!dinv_live) { s.Dinv = x; x += nD; }
  if (x > xmax) xmax = x;
  x = L.n_in;
  // the broad-phase work queue holds at most 2 G - 1 candidates (a full batch is flushed as soon as G are queued)
  // geom poses are kept for full-warp models; small-group models recompute them per pair (gsr_dynamics.cuh geom_pose)
  if (h->group == 32 || h->group == 1) { s.gpos = x; x += 3 * nG; s.gquat = x; x += 4 * nG; } else { s.gpos = -1; s.gquat = -1; }
  s.aabb = x; x += 6 * nG; s.queue = x; x += 2 * (h->group == 1 ? 32 : h->group);
  if (x > xmax) xmax = x;
  // integration scratch (link accelerations, next velocity / position): written out by the epilogue before the kinematics
  // of the new configuration start, whose scratch (joint anchors / axes, the two scan buffers) therefore overlays it.  The second
  // scan buffer goes on top of M / L (dead whenever the kinematics run: k_step's epilogue, k_forward, k_set_state) when it fits there.
  x = L.n_in;
  s.cacc = x; x += 6 * nL; s.vel_next = x; x += nD; s.qpos_next = x; x += nQ;
  if (x > xmax) xmax = x;
  x = L.n_in;
  s.xanchor = x; x += 3 * m.nJ; s.xaxis = x; x += 3 * m.nJ; s.pose_a = x; x += 8 * nL;
  const bool pose_b_on_M = 8 * nL <= nD * (nD + 1) / 2 + nD * MS;
  if (!pose_b_on_M) { s.pose_b = x; x += 8 * nL; }
  if (x > xmax) xmax = x;
  // solver arrays overlay the dead part of the kinematic cache and the phase scratch
  x = L.kin_dead;
  s.contacts = x; x += s.NCC * GSR_CONTACT_WORDS;
  const bool h_own = !newton || m.noslip_iterations > 0;  // noslip solves with the L^T D L factor after the Newton solve
  if (h_own) { s.H = x; x += nD * MS; }  // otherwise H overlays the L^T D L factor, which is dead once acc_smooth is known
  s.Hinv = x; x += nD; s.tmpv = x; x += nD;
  s.qacc = x; x += nD; s.Ma = x; x += nD; s.grad = x; x += nD; s.Mgrad = x; x += nD; s.search = x; x += nD; s.mv = x; x += nD;
  s.prev_grad = x; x += nD; s.prev_Mgrad = x; x += nD;
  s.r_aref = x; x += NCS; s.r_D = x; x += NCS; s.r_R = x; x += NCS; s.r_floss = x; x += NCS; s.r_Jaref = x; x += NCS;
  s.r_jv = x; x += NCS; s.r_force = x; x += NCS; s.r_active = x; x += NCS;
  s.J = x; x += NCS * JS;
  if (x > xmax) xmax = x;
  // always-live block
  x = (int)gsr_round4(xmax);
  s.M = x; x += nD * (nD + 1) / 2;  // packed lower triangle (gsr_group.cuh sym_row_dot)
  s.L = x; x += nD * MS;
  if (fdof_on_L) s.fdof = s.L;
  if (pose_b_on_M) s.pose_b = s.M;
  if (dinv_live) { s.Dinv = x; x += nD; }
  if (!h_own) s.H = s.L;
  s.force = x; x += nD; s.acc_smooth = x; x += nD;
  // acc is written when the solver ends, after the last read of acc_smooth (cost evaluations), element by element: one slot, unless
  // the diagnostics read acc_smooth in the epilogue.  qf_applied waits in acc's slot from the actuation loop to the end of the phase.
  // (noslip reads acc_smooth after the solver has committed acc: no sharing then)
  if (dyn_terms_live || m.noslip_iterations > 0) { s.acc = x; x += nD; } else s.acc = s.acc_smooth;
  s.qf_applied = s.acc;
  if (dyn_terms_live) { s.qf_passive = x; x += nD; s.qf_bias = x; x += nD; }
  s.qfrc_c = x; x += nD;
  s.env_words = (int)gsr_round4(x) - L.in_start;  // the PRE block of the record is not mirrored


  // ---- lane-per-environment kernel (group == 1): interleaved workspace, word w of lane t at S[w * 33 + t] (gsr_lane.cuh)
  {
    LaneLayout& q = h->ll;
    memset(&q, 0, sizeof(q));
    const int nR = m.n_roots > 0 ? m.n_roots : 1;
    const int ntri = nD * (nD + 1) / 2;
    int NCSl = (int)S("cfg_lane_rows", 8);
    if (NCSl > lenC) NCSl = lenC;
    if (NCSl < 4) NCSl = 4;
    q.NCS = NCSl;
    q.NCC = NCSl / 4 > 0 ? NCSl / 4 : 1;
    int y = 0;
    q.n_head = L.root_COM; y = q.n_head;  // record words [0, root_COM): qpos .. has_ext_force, dofs_pos (same offsets)
    q.lpos = y; y += 3 * nL; q.lquat = y; y += 4 * nL;
    q.rcom = y; y += 3 * nR;
    q.cdof = y; y += 12 * nD;
    q.M = y; y += ntri; q.L = y; y += ntri; q.Dinv = y; y += nD; q.Hinv = y; y += nD;
    q.force = y; y += nD; q.acc_smooth = y; y += nD; q.acc = y; y += nD; q.qf_applied = y; y += nD; q.qfrc_c = y; y += nD;
    q.pool = y;
    int ymax = y;
    // dynamics scratch
    q.crb = y; y += 9 * nL; q.cfrc = y; y += 6 * nL; q.cvel = y; y += 6 * nL; q.cdd = y; y += 6 * nL;
    if (y > ymax) ymax = y;
    // collision scratch; the on-chip contacts and the per-link contact forces live until the end of the constraint phase
    y = q.pool;
    q.ct = y; y += q.NCC * GSR_CONTACT_WORDS; q.lcf = y; y += 3 * nL;
    const int after_lcf = y;
    q.aabb = y; y += 6 * nG;
    q.nqueue = m.nP < m.max_broad_pairs ? m.nP : m.max_broad_pairs;
    if (q.nqueue < 1) q.nqueue = 1;
    q.queue = y; y += q.nqueue;
    if (y > ymax) ymax = y;
    // solver
    y = after_lcf;
    q.qacc = y; y += nD; q.Ma = y; y += nD; q.grad = y; y += nD; q.Mgrad = y; y += nD; q.search = y; y += nD; q.mv = y; y += nD;
    q.prev_grad = y; y += nD; q.prev_Mgrad = y; y += nD; q.tmpv = y; y += nD;
    if (m.solver_type != SOLVER_NEWTON || m.noslip_iterations > 0) { q.H = y; y += ntri; } else q.H = q.L;  // Newton: H overlays the dead L^T D L factor
    q.r_aref = y; y += NCSl; q.r_D = y; y += NCSl; q.r_R = y; y += NCSl; q.r_floss = y; y += NCSl; q.r_Jaref = y; y += NCSl;
    q.r_jv = y; y += NCSl; q.r_force = y; y += NCSl; q.r_active = y; y += NCSl;
    q.J = y; y += NCSl * nD;
    if (y > ymax) ymax = y;
    // outputs + integration + final kinematics
    y = after_lcf;
    q.cacc = y; y += 6 * nL; q.out4 = y; y += 4 * nD; q.cnt = y; y += 4; q.vn = y; y += nD; q.qn = y; y += nQ;
    if (y > ymax) ymax = y;
    y = after_lcf;
    q.fcvel = y; y += 6 * nL; q.rcom_out = y; y += 3 * nL; q.cinr = y; y += 9 * nL;
    if (y > ymax) ymax = y;
    q.env_words = ymax;
    // a warp's interleaved workspace must fit one SM's shared memory; larger models run the lane-group kernel
    if (h->group == 1 && (size_t)q.env_words * 33 * sizeof(float) > 227 * 1024) h->group = (m.nD <= 12 && m.nL <= 16) ? 8 : 32;
  }
  return GSR_OK;
}
