// gsr_api.cu -- C ABI of libgsrigid.so (see include/gsrigid.h for the contract and the reference interfaces replaced).
#include <cuda_runtime.h>
#include <cstdio>
#include <cstring>
#include <map>
#include <string>
#include <vector>

#include "../../include/gsrigid.h"
#include <cstdint>
#include "gsr_kernels.cuh"

namespace {

struct FieldInfo {
  std::string name;
  int is_int;
  size_t seg_base, env_stride, offset, items, comps, item_stride;
};

struct KernelStat {
  const char* name;
  long long launches;
};

enum { K_STEP = 0, K_FORWARD = 1, K_RESET = 2, K_INIT = 3, K_SCATTER = 4, K_GATHER = 5, K_COUNT = 6 };

}  // namespace

struct gsr_handle_s {
  std::map<std::string, std::vector<float>> f;
  std::map<std::string, std::vector<int>> i;
  std::map<std::string, double> s;
  std::string err;
  bool committed = false;
  int device = 0, B = 0;
  int group = 32;  // lanes per environment
  int warps_per_cta = 0;
  DevModel hm;     // host copy of the model
  HotLayout lay;
  SmemLayout sl;
  std::vector<FieldInfo> fields;
  size_t n_words = 0;
  float* state = nullptr;
  DevModel* d_model = nullptr;  // global-memory copy for lane-varying table reads
  float *d_verts = nullptr, *d_support = nullptr;
  int *d_pairs = nullptr, *d_pair_idx = nullptr;
  // staging for gsr_step_host
  float *d_stage = nullptr, *h_stage = nullptr;
  size_t stage_elems = 0;
  KernelStat stats[K_COUNT] = {{"k_step", 0}, {"k_forward", 0}, {"k_reset", 0}, {"k_init", 0}, {"k_scatter", 0}, {"k_gather", 0}};
  long long total_launches = 0;
  // optional per-kernel event timing (gsr_set_profiling)
  bool profiling = false;
  struct Span { int k; cudaEvent_t a, b; };
  std::vector<Span> spans;
  std::vector<cudaEvent_t> event_pool;
  double time_ms[K_COUNT] = {0, 0, 0, 0, 0, 0};
  long long timed[K_COUNT] = {0, 0, 0, 0, 0, 0};
  cudaStream_t prof_stream = nullptr;
};

// event bracket around one kernel launch; no-op unless profiling is enabled
struct Bracket {
  gsr_handle_s* h; int k; cudaStream_t st; cudaEvent_t a = nullptr, b = nullptr;
  static cudaEvent_t get(gsr_handle_s* h) {
    if (!h->event_pool.empty()) { cudaEvent_t e = h->event_pool.back(); h->event_pool.pop_back(); return e; }
    cudaEvent_t e; cudaEventCreate(&e); return e;
  }
  Bracket(gsr_handle_s* h_, int k_, cudaStream_t st_) : h(h_), k(k_), st(st_) {
    h->stats[k].launches++; h->total_launches++;
    if (h->profiling) { a = get(h); b = get(h); cudaEventRecord(a, st); }
  }
  ~Bracket() {
    if (a) { cudaEventRecord(b, st); h->spans.push_back({k, a, b}); h->prof_stream = st; }
  }
};

// which handle's model currently sits in __constant__ memory of each device
static gsr_handle_s* g_active[64] = {nullptr};

static int fail(gsr_handle_t h, int code, const std::string& msg) {
  if (h) h->err = msg;
  return code;
}
#define CUDA_OK(call)                                                                    \
  do {                                                                                   \
    cudaError_t e_ = (call);                                                             \
    if (e_ != cudaSuccess) return fail(h, GSR_ECUDA, std::string(#call) + ": " + cudaGetErrorString(e_)); \
  } while (0)

static int activate(gsr_handle_t h, cudaStream_t st) {
  CUDA_OK(cudaSetDevice(h->device));
  if (g_active[h->device] != h) {
    // constant memory is per device: re-upload when another model (e.g. a second scene on this GPU) ran last
    CUDA_OK(cudaMemcpyToSymbolAsync(c_m, &h->hm, sizeof(DevModel), 0, cudaMemcpyHostToDevice, st));
    CUDA_OK(cudaMemcpyToSymbolAsync(c_h, &h->lay, sizeof(HotLayout), 0, cudaMemcpyHostToDevice, st));
    CUDA_OK(cudaMemcpyToSymbolAsync(c_s, &h->sl, sizeof(SmemLayout), 0, cudaMemcpyHostToDevice, st));
    CUDA_OK(cudaMemcpyAsync(h->d_model, &h->hm, sizeof(DevModel), cudaMemcpyHostToDevice, st));
    g_active[h->device] = h;
  }
  return GSR_OK;
}

static int fill_model(gsr_handle_t h) {
  DevModel& m = h->hm;
  const float* keep_verts = m.verts; const float* keep_support = m.support_v; const int* keep_pairs = m.pairs; const int* keep_pidx = m.pair_idx;
  memset(&m, 0, sizeof(m));
  m.verts = keep_verts; m.support_v = keep_support; m.pairs = keep_pairs; m.pair_idx = keep_pidx;
  auto S = [&](const char* k, double d) { auto it = h->s.find(k); return it == h->s.end() ? d : it->second; };
  m.nL = (int)S("n_links", 0); m.nJ = (int)S("n_joints", 0); m.nD = (int)S("n_dofs", 0); m.nQ = (int)S("n_qs", 0);
  m.nG = (int)S("n_geoms", 0); m.nE = (int)S("n_entities", 0); m.nEq = (int)S("n_equalities", 0); m.nP = (int)S("n_pairs", 0);
  m.nV = (int)S("n_verts", 0);
  if (m.nL > GSR_MAX_LINKS || m.nJ > GSR_MAX_JOINTS || m.nD > GSR_MAX_DOFS || m.nQ > GSR_MAX_QS || m.nG > GSR_MAX_GEOMS ||
      m.nE > GSR_MAX_ENTS || m.nEq > GSR_MAX_EQ) {
    char buf[256];
    snprintf(buf, sizeof(buf), "model exceeds compiled capacities (links %d/%d joints %d/%d dofs %d/%d qs %d/%d geoms %d/%d entities %d/%d eq %d/%d)",
             m.nL, GSR_MAX_LINKS, m.nJ, GSR_MAX_JOINTS, m.nD, GSR_MAX_DOFS, m.nQ, GSR_MAX_QS, m.nG, GSR_MAX_GEOMS, m.nE, GSR_MAX_ENTS, m.nEq, GSR_MAX_EQ);
    return fail(h, GSR_ELIMIT, buf);
  }
  m.max_contact_pairs = (int)S("max_contact_pairs", 0); m.max_broad_pairs = (int)S("max_broad_pairs", 0);
  m.n_contacts_per_pair = (int)S("n_contacts_per_pair", 5); m.len_constraints = (int)S("len_constraints", 0);
  m.iterations = (int)S("cfg_iterations", 50); m.ls_iterations = (int)S("cfg_ls_iterations", 50);
  m.integrator = (int)S("cfg_integrator", INTEGRATOR_APPROX_IMPLICITFAST); m.solver_type = (int)S("cfg_solver_type", SOLVER_NEWTON);
  m.enable_collision = (int)S("cfg_enable_collision", 1); m.enable_joint_limit = (int)S("cfg_enable_joint_limit", 1);
  m.enable_multi_contact = (int)S("cfg_enable_multi_contact", 1); m.sort_contacts = (int)S("cfg_sort_contacts", 1);
  m.meaninertia = (float)S("meaninertia", 1.0); m.dt = (float)S("dt", 0.01); m.eps = (float)S("cfg_eps", 1.1920929e-07);
  m.tolerance = (float)S("cfg_tolerance", 1e-6); m.ls_tolerance = (float)S("cfg_ls_tolerance", 1e-2);
  m.mc_perturbation = 1e-2f; m.mc_tolerance = 1e-2f;  // collider.py:85-86
  m.gravity[0] = (float)S("gravity_x", 0.0); m.gravity[1] = (float)S("gravity_y", 0.0); m.gravity[2] = (float)S("gravity_z", -9.81);
  m.ccd_eps = 1e-9f; m.ccd_tolerance = 1e-6f; m.ccd_iterations = 50;  // mpr.py:13-19
  m.n_cache_words = (m.nP + 31) / 32 > 0 ? (m.nP + 31) / 32 : 1;
  m.diagnostics = (int)S("cfg_diagnostics", 1);

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
  { const float* p = FV("links_invweight", 2 * nL); if (!missing) for (int l = 0; l < nL; l++) m.l_invweight[l] = p[2 * l]; }
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
  if (missing) return fail(h, GSR_EINVAL, "missing or short model array: " + missing_name);

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
  for (int l = 0; l < nL; l++) {
    m.l_descmask[l] = 0ull;
    for (int d = 0; d < nL; d++)
      if ((m.l_ancmask[d] >> l) & 1ull) m.l_descmask[l] |= 1ull << d;
  }

  // has_cc: any valid pair that is not plane-box (collider.py:449-459) -> contacts are x-sorted (collider.py:804-808)
  m.has_cc = 0;
  {
    const int* pr = IV("pairs", 2 * (size_t)m.nP);
    for (int p = 0; p < m.nP; p++) {
      int ta = m.g_type[pr[2 * p]], tb = m.g_type[pr[2 * p + 1]];
      if (ta > tb) { int t = ta; ta = tb; tb = t; }
      if (!(ta == GEOM_PLANE && tb == GEOM_BOX)) m.has_cc = 1;
    }
  }

  return GSR_OK;
}

extern "C" {

const char* gsr_version(void) { return "gsrigid 0.1 (sm_100a)"; }
const char* gsr_last_error(gsr_handle_t h) { return h ? h->err.c_str() : "null handle"; }

int gsr_model_begin(gsr_handle_t* out) {
  if (!out) return GSR_EINVAL;
  *out = new gsr_handle_s();
  return GSR_OK;
}
int gsr_model_set_f32(gsr_handle_t h, const char* name, const float* data, size_t count) {
  if (!h || !name || (!data && count)) return GSR_EINVAL;
  h->f[name] = std::vector<float>(data, data + count);
  return GSR_OK;
}
int gsr_model_set_i32(gsr_handle_t h, const char* name, const int32_t* data, size_t count) {
  if (!h || !name || (!data && count)) return GSR_EINVAL;
  h->i[name] = std::vector<int>(data, data + count);
  return GSR_OK;
}
int gsr_model_set_scalar(gsr_handle_t h, const char* name, double value) {
  if (!h || !name) return GSR_EINVAL;
  h->s[name] = value;
  return GSR_OK;
}

static size_t round4(size_t x) { return (x + 3) & ~(size_t)3; }

int gsr_model_commit(gsr_handle_t h, int device, int n_envs) {
  if (!h || n_envs <= 0) return fail(h, GSR_EINVAL, "bad handle or n_envs");
  if (h->committed) return fail(h, GSR_ESTATE, "model already committed (use gsr_model_refresh)");
  memset(&h->hm, 0, sizeof(h->hm));
  int rc_fill = fill_model(h);
  if (rc_fill) return rc_fill;
  DevModel& m = h->hm;
  if (m.nL > 64 || m.nD > 64) return fail(h, GSR_ELIMIT, "ancestor masks hold at most 64 links / dofs");
  h->device = device;
  h->B = n_envs;
  CUDA_OK(cudaSetDevice(device));
  auto upload = [&](const void* src, size_t bytes, void** dst) -> cudaError_t {
    if (bytes == 0) bytes = 4;
    cudaError_t e = cudaMalloc(dst, bytes);
    if (e != cudaSuccess) return e;
    return src ? cudaMemcpy(*dst, src, bytes, cudaMemcpyHostToDevice) : cudaMemset(*dst, 0, bytes);
  };
  {
    auto& v = h->f["verts"]; CUDA_OK(upload(v.empty() ? nullptr : v.data(), v.size() * 4, (void**)&h->d_verts));
    auto& sv = h->f["support_v"]; CUDA_OK(upload(sv.empty() ? nullptr : sv.data(), sv.size() * 4, (void**)&h->d_support));
    auto& pr = h->i["pairs"]; CUDA_OK(upload(pr.empty() ? nullptr : pr.data(), pr.size() * 4, (void**)&h->d_pairs));
    auto& pi = h->i["pair_idx"]; CUDA_OK(upload(pi.empty() ? nullptr : pi.data(), pi.size() * 4, (void**)&h->d_pair_idx));
    CUDA_OK(cudaMalloc((void**)&h->d_model, sizeof(DevModel)));
  }
  m.verts = h->d_verts; m.support_v = h->d_support; m.pairs = h->d_pairs; m.pair_idx = h->d_pair_idx;

  // ---- lanes per environment: a full warp unless the batch is large enough to fill the machine with smaller groups
  auto S = [&](const char* k, double d) { auto it = h->s.find(k); return it == h->s.end() ? d : it->second; };
  int G = (int)S("cfg_group_size", 0);
  if (G != 8 && G != 16 && G != 32) G = (m.nD <= 12 && m.nL <= 16) ? 8 : 32;  // measured on B200: profiles/README.md
  h->group = G;
  h->warps_per_cta = (int)S("cfg_warps_per_cta", 0);

  // ---- HOT record layout
  const int nL = m.nL, nD = m.nD, nQ = m.nQ, nG = m.nG;
  HotLayout& L = h->lay;
  memset(&L, 0, sizeof(L));
  h->fields.clear();
  size_t off = 0;
#define X(name, is_int, items, comps)                                                                   \
  {                                                                                                     \
    size_t it_ = (size_t)(items), cp_ = (size_t)(comps);                                                \
    L.name = (int)off;                                                                                  \
    h->fields.push_back(FieldInfo{#name, is_int, 0, 0, off, it_, cp_, cp_});                            \
    off += (it_ * cp_ > 0 ? it_ * cp_ : 1);                                                             \
  }
  GSR_HOT_IN_FIELDS(X)
  off = round4(off);
  L.n_in = (int)off;
  L.kin_dead = L.cinr_I;
  GSR_HOT_OUT_FIELDS(X)
#undef X
  L.hot_stride = (int)round4(off);
  const int nCP = m.max_contact_pairs > 0 ? m.max_contact_pairs : 1;
  const int lenC = m.len_constraints > 0 ? m.len_constraints : 1;
  const int MS = nD | 1, JS = nD | 1;
  L.contact_stride = nCP * GSR_CONTACT_WORDS;
  L.cache_valid = (int)round4(3 * (size_t)(m.nP > 0 ? m.nP : 1));
  L.cache_stride = (int)round4(L.cache_valid + m.n_cache_words);
  size_t spill = (size_t)lenC * JS + 8 * (size_t)lenC;
  if (spill < (size_t)nCP * GSR_CONTACT_WORDS) spill = (size_t)nCP * GSR_CONTACT_WORDS;
  L.spill_stride = (int)round4(spill);
  const size_t B = (size_t)n_envs;
  L.contact_base = (long long)(B * L.hot_stride);
  L.cache_base = L.contact_base + (long long)(B * L.contact_stride);
  L.spill_base = L.cache_base + (long long)(B * L.cache_stride);
  h->n_words = (size_t)L.spill_base + B * L.spill_stride;
  for (auto& f : h->fields) { f.seg_base = 0; f.env_stride = L.hot_stride; }
  {  // contact record fields (AoS) and the warm-start cache
    auto add = [&](const char* name, int is_int, size_t base, size_t stride, size_t o, size_t items, size_t comps, size_t istride) {
      h->fields.push_back(FieldInfo{name, is_int, base, stride, o, items, comps, istride});
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
  int NCS = (int)S("cfg_smem_rows", nD > 12 ? 64 : 24);
  if (NCS > lenC) NCS = lenC;
  if (NCS < 4) NCS = 4;
  s.NCS = NCS;
  s.NCC = NCS / 4 > 0 ? NCS / 4 : 1;
  // scratch of the dynamics / collision / integrate+FK phases starts after the IN mirror
  int x = L.n_in, xmax = x;
  s.crb = x; x += 10 * nL; s.fdof = x; x += 6 * nD; s.cdd = x; x += 6 * nL; s.cfrc = x; x += 6 * nL;
  if (x > xmax) xmax = x;
  x = L.n_in;
  s.gpos = x; x += 3 * nG; s.gquat = x; x += 4 * nG; s.aabb = x; x += 6 * nG; s.queue = x; x += 2 * 32;
  if (x > xmax) xmax = x;
  x = L.n_in;
  s.xanchor = x; x += 3 * m.nJ; s.xaxis = x; x += 3 * m.nJ; s.cacc = x; x += 6 * nL; s.vel_next = x; x += nD; s.qpos_next = x; x += nQ;
  s.pose_a = x; x += 8 * nL; s.pose_b = x; x += 8 * nL;
  if (x > xmax) xmax = x;
  // solver arrays overlay the dead part of the kinematic cache and the phase scratch
  x = L.kin_dead;
  s.contacts = x; x += s.NCC * GSR_CONTACT_WORDS;
  const bool newton = m.solver_type == SOLVER_NEWTON;
  if (!newton) { s.H = x; x += nD * MS; }  // Newton: H overlays the L^T D L factor, which is dead once acc_smooth is known
  s.Hinv = x; x += nD; s.tmpv = x; x += nD;
  s.qacc = x; x += nD; s.Ma = x; x += nD; s.grad = x; x += nD; s.Mgrad = x; x += nD; s.search = x; x += nD; s.mv = x; x += nD;
  s.prev_grad = x; x += nD; s.prev_Mgrad = x; x += nD;
  s.r_aref = x; x += NCS; s.r_D = x; x += NCS; s.r_R = x; x += NCS; s.r_floss = x; x += NCS; s.r_Jaref = x; x += NCS;
  s.r_jv = x; x += NCS; s.r_force = x; x += NCS; s.r_active = x; x += NCS;
  s.J = x; x += NCS * JS;
  if (x > xmax) xmax = x;
  // always-live block
  x = (int)round4(xmax);
  s.M = x; x += nD * MS; s.L = x; x += nD * MS; s.Dinv = x; x += nD;
  if (newton) s.H = s.L; s.force = x; x += nD; s.acc_smooth = x; x += nD;
  s.acc = x; x += nD; s.qf_applied = x; x += nD; s.qf_passive = x; x += nD; s.qf_bias = x; x += nD; s.qfrc_c = x; x += nD;
  s.env_words = (int)round4(x);

  h->committed = true;
  g_active[device] = nullptr;
  return GSR_OK;
}

int gsr_model_refresh(gsr_handle_t h) {
  if (!h || !h->committed) return fail(h, GSR_ESTATE, "model not committed");
  int rc = fill_model(h);
  if (rc) return rc;
  g_active[h->device] = nullptr;  // forces a re-upload on the next call
  return GSR_OK;
}

int gsr_state_words(gsr_handle_t h, size_t* n_words) {
  if (!h || !h->committed) return fail(h, GSR_ESTATE, "model not committed");
  if (n_words) *n_words = h->n_words;
  return GSR_OK;
}
int gsr_field_count(gsr_handle_t h) { return (h && h->committed) ? (int)h->fields.size() : GSR_ESTATE; }
int gsr_field_info(gsr_handle_t h, int index, const char** name, int* is_int, size_t* segment_base, size_t* env_stride,
                   size_t* offset, size_t* items, size_t* comps, size_t* item_stride) {
  if (!h || !h->committed) return fail(h, GSR_ESTATE, "model not committed");
  if (index < 0 || index >= (int)h->fields.size()) return fail(h, GSR_EINVAL, "field index out of range");
  const FieldInfo& f = h->fields[index];
  if (name) *name = f.name.c_str();
  if (is_int) *is_int = f.is_int;
  if (segment_base) *segment_base = f.seg_base;
  if (env_stride) *env_stride = f.env_stride;
  if (offset) *offset = f.offset;
  if (items) *items = f.items;
  if (comps) *comps = f.comps;
  if (item_stride) *item_stride = f.item_stride;
  return GSR_OK;
}
int gsr_bind(gsr_handle_t h, void* state_words) {
  if (!h || !h->committed) return fail(h, GSR_ESTATE, "model not committed");
  if (!state_words) return fail(h, GSR_EINVAL, "null state buffer");
  if (((uintptr_t)state_words & 15) != 0) return fail(h, GSR_EINVAL, "state buffer must be 16-byte aligned");
  h->state = (float*)state_words;
  return GSR_OK;
}
int gsr_launch_info(gsr_handle_t h, int* group_size, int* smem_bytes_per_env, int* smem_rows, int* smem_contacts) {
  if (!h || !h->committed) return fail(h, GSR_ESTATE, "model not committed");
  if (group_size) *group_size = h->group;
  if (smem_bytes_per_env) *smem_bytes_per_env = h->sl.env_words * 4;
  if (smem_rows) *smem_rows = h->sl.NCS;
  if (smem_contacts) *smem_contacts = h->sl.NCC;
  return GSR_OK;
}

}  // extern "C"

// ---- launches ---------------------------------------------------------------------------------------------------------
template <typename K>
static cudaError_t set_smem(K kernel, size_t bytes) {
  return cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
}
static int group_launch(gsr_handle_t h, int which, cudaStream_t st, const uint8_t* mask, int clear_ext) {
  const int G = h->group, EPW = 32 / G;
  const size_t warp_bytes = (size_t)EPW * h->sl.env_words * sizeof(float);
  int warps = h->warps_per_cta > 0 ? h->warps_per_cta : (G == 32 ? 8 : 4);  // measured on B200 (profiles/README.md)
  if (warps > GSR_MAX_WARPS_PER_CTA) warps = GSR_MAX_WARPS_PER_CTA;
  while (warps > 1 && warps * (warp_bytes + 8) > 224 * 1024) warps--;  // one big CTA per SM, its warps in phase lockstep
  const int envs_per_cta = warps * EPW;
  const size_t smem = (size_t)warps * warp_bytes + (size_t)warps * 8;  // + one mbarrier per warp (TMA staging)
  if (smem > 227 * 1024) return fail(h, GSR_ELIMIT, "per-warp shared-memory workspace exceeds 227 KB (lower cfg_smem_rows or raise cfg_group_size)");
  dim3 grid((h->B + envs_per_cta - 1) / envs_per_cta), block(warps * 32);
#define LAUNCH(GG)                                                                                       \
  if (which == K_STEP) {                                                                                 \
    CUDA_OK(set_smem(k_step<GG>, smem));                                                                 \
    Bracket br_(h, K_STEP, st);                                                                          \
    k_step<GG><<<grid, block, smem, st>>>(h->state, h->B, h->d_model, clear_ext);                        \
  } else {                                                                                               \
    CUDA_OK(set_smem(k_forward<GG>, smem));                                                              \
    Bracket br_(h, K_FORWARD, st);                                                                       \
    k_forward<GG><<<grid, block, smem, st>>>(h->state, h->B, h->d_model, mask);                          \
  }
  if (G == 32) { LAUNCH(32) } else if (G == 16) { LAUNCH(16) } else { LAUNCH(8) }
#undef LAUNCH
  CUDA_OK(cudaGetLastError());
  return GSR_OK;
}

extern "C" {

int gsr_init_state(gsr_handle_t h, void* cuda_stream) {
  if (!h || !h->state) return fail(h, GSR_ESTATE, "state not bound");
  cudaStream_t st = (cudaStream_t)cuda_stream;
  int rc = activate(h, st);
  if (rc) return rc;
  CUDA_OK(cudaMemsetAsync(h->state, 0, h->n_words * sizeof(float), st));
  {
    Bracket br_(h, K_INIT, st);
    k_init<<<(h->B + 127) / 128, 128, 0, st>>>(h->state, h->B);
  }
  CUDA_OK(cudaGetLastError());
  return group_launch(h, K_FORWARD, st, nullptr, 0);
}

int gsr_step(gsr_handle_t h, int n_substeps, void* cuda_stream) {
  if (!h || !h->state) return fail(h, GSR_ESTATE, "state not bound");
  cudaStream_t st = (cudaStream_t)cuda_stream;
  int rc = activate(h, st);
  if (rc) return rc;
  for (int s = 0; s < n_substeps; s++) {
    rc = group_launch(h, K_STEP, st, nullptr, s == n_substeps - 1);
    if (rc) return rc;
  }
  return GSR_OK;
}

int gsr_forward_kinematics(gsr_handle_t h, const uint8_t* env_mask_dev, void* cuda_stream) {
  if (!h || !h->state) return fail(h, GSR_ESTATE, "state not bound");
  cudaStream_t st = (cudaStream_t)cuda_stream;
  int rc = activate(h, st);
  if (rc) return rc;
  return group_launch(h, K_FORWARD, st, env_mask_dev, 0);
}

int gsr_reset_envs(gsr_handle_t h, const uint8_t* env_mask_dev, void* cuda_stream) {
  if (!h || !h->state) return fail(h, GSR_ESTATE, "state not bound");
  cudaStream_t st = (cudaStream_t)cuda_stream;
  int rc = activate(h, st);
  if (rc) return rc;
  {
    Bracket br_(h, K_RESET, st);
    k_reset<<<(h->B + 127) / 128, 128, 0, st>>>(h->state, h->B, env_mask_dev);
  }
  CUDA_OK(cudaGetLastError());
  return GSR_OK;
}

int gsr_step_host(gsr_handle_t h, int n_substeps, const float* ctrl_pos_host, float* qpos_host, float* vel_host, void* cuda_stream) {
  if (!h || !h->state) return fail(h, GSR_ESTATE, "state not bound");
  cudaStream_t st = (cudaStream_t)cuda_stream;
  int rc = activate(h, st);
  if (rc) return rc;
  const DevModel& m = h->hm;
  const size_t B = h->B;
  const size_t need = B * (size_t)(m.nQ + m.nD);
  if (h->stage_elems < need) {
    if (h->d_stage) cudaFree(h->d_stage);
    CUDA_OK(cudaMalloc((void**)&h->d_stage, need * 4));
    h->stage_elems = need;
  }
  // host arrays are (B, n) env-major like the public API; one contiguous DMA each way plus a scatter / gather kernel
  if (ctrl_pos_host) {
    const int n = (int)(B * m.nD);
    CUDA_OK(cudaMemcpyAsync(h->d_stage, ctrl_pos_host, (size_t)n * 4, cudaMemcpyHostToDevice, st));
    Bracket br_(h, K_SCATTER, st);
    k_scatter<<<(n + 255) / 256, 256, 0, st>>>(h->state, h->B, h->lay.ctrl_pos, m.nD, h->d_stage);
  }
  rc = gsr_step(h, n_substeps, cuda_stream);
  if (rc) return rc;
  float* d_q = h->d_stage;
  float* d_v = h->d_stage + B * m.nQ;
  if (qpos_host) {
    const int n = (int)(B * m.nQ);
    { Bracket br_(h, K_GATHER, st); k_gather<<<(n + 255) / 256, 256, 0, st>>>(h->state, h->B, h->lay.qpos, m.nQ, d_q); }
    CUDA_OK(cudaMemcpyAsync(qpos_host, d_q, (size_t)n * 4, cudaMemcpyDeviceToHost, st));
  }
  if (vel_host) {
    const int n = (int)(B * m.nD);
    { Bracket br_(h, K_GATHER, st); k_gather<<<(n + 255) / 256, 256, 0, st>>>(h->state, h->B, h->lay.vel, m.nD, d_v); }
    CUDA_OK(cudaMemcpyAsync(vel_host, d_v, (size_t)n * 4, cudaMemcpyDeviceToHost, st));
  }
  CUDA_OK(cudaGetLastError());
  CUDA_OK(cudaStreamSynchronize(st));
  return GSR_OK;
}

int gsr_set_profiling(gsr_handle_t h, int enabled) {
  if (!h) return GSR_EINVAL;
  if (enabled) { for (int k = 0; k < K_COUNT; k++) { h->time_ms[k] = 0; h->timed[k] = 0; } }
  h->profiling = enabled != 0;
  return GSR_OK;
}
int gsr_kernel_time_ms(gsr_handle_t h, int index, double* total_ms, long long* timed_launches) {
  if (!h || index < 0 || index >= K_COUNT) return GSR_EINVAL;
  if (!h->spans.empty()) {
    CUDA_OK(cudaSetDevice(h->device));
    CUDA_OK(cudaEventSynchronize(h->spans.back().b));
    for (auto& sp : h->spans) {
      float ms = 0.f;
      CUDA_OK(cudaEventElapsedTime(&ms, sp.a, sp.b));
      h->time_ms[sp.k] += ms; h->timed[sp.k]++;
      h->event_pool.push_back(sp.a); h->event_pool.push_back(sp.b);
    }
    h->spans.clear();
  }
  if (total_ms) *total_ms = h->time_ms[index];
  if (timed_launches) *timed_launches = h->timed[index];
  return GSR_OK;
}

long long gsr_launch_count(gsr_handle_t h) { return h ? h->total_launches : 0; }
int gsr_kernel_info(gsr_handle_t h, int index, const char** name, long long* launches) {
  if (!h || index < 0 || index >= K_COUNT) return GSR_EINVAL;
  if (name) *name = h->stats[index].name;
  if (launches) *launches = h->stats[index].launches;
  return GSR_OK;
}

void gsr_destroy(gsr_handle_t h) {
  if (!h) return;
  if (h->committed) {
    cudaSetDevice(h->device);
    if (g_active[h->device] == h) g_active[h->device] = nullptr;
    cudaFree(h->d_verts); cudaFree(h->d_support); cudaFree(h->d_pairs); cudaFree(h->d_pair_idx); cudaFree(h->d_model);
    if (h->d_stage) cudaFree(h->d_stage);
    for (auto& sp : h->spans) { cudaEventDestroy(sp.a); cudaEventDestroy(sp.b); }
    for (auto e : h->event_pool) cudaEventDestroy(e);
  }
  delete h;
}

}  // extern "C"
