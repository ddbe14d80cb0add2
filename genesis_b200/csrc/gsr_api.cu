// gsr_api.cu -- C ABI of libgsrigid.so (see include/gsrigid.h for the contract and the reference interfaces replaced).
#include <cuda_runtime.h>
#include <cstdio>
#include <cstring>
#include <map>
#include <string>
#include <vector>

#include "../../include/gsrigid.h"
#include "gsr_kernels.cuh"

namespace {

struct FieldInfo {
  std::string name;
  int is_int;
  size_t offset, items;
};

struct KernelStat {
  const char* name;
  long long launches;
};

}  // namespace

struct gsr_handle_s {
  std::map<std::string, std::vector<float>> f;
  std::map<std::string, std::vector<int>> i;
  std::map<std::string, double> s;
  std::string err;
  bool committed = false;
  int device = 0, B = 0;
  DevModel hm;  // host copy of the constant-memory model
  Layout lay;
  std::vector<FieldInfo> fields;
  size_t n_f32 = 0, n_i32 = 0;
  float* F = nullptr;
  int* I = nullptr;
  float *d_verts = nullptr, *d_support = nullptr;
  int *d_pairs = nullptr, *d_pair_idx = nullptr;
  // staging for gsr_step_host
  float *d_stage = nullptr, *h_stage = nullptr;
  size_t stage_elems = 0;
  KernelStat stats[6] = {{"k_dynamics", 0}, {"k_collide", 0}, {"k_constraint", 0}, {"k_integrate", 0}, {"k_forward", 0}, {"k_reset", 0}};
  long long total_launches = 0;
  // optional per-kernel event timing (gsr_set_profiling)
  bool profiling = false;
  struct Span { int k; cudaEvent_t a, b; };
  std::vector<Span> spans;
  std::vector<cudaEvent_t> event_pool;
  double time_ms[6] = {0, 0, 0, 0, 0, 0};
  long long timed[6] = {0, 0, 0, 0, 0, 0};
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
    CUDA_OK(cudaMemcpyToSymbolAsync(c_lay, &h->lay, sizeof(Layout), 0, cudaMemcpyHostToDevice, st));
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
  CPI(m.d_link, "dofs_link", nD) CPI(m.d_entity, "dofs_entity", nD)
  CPF(m.d_motion_ang, "dofs_motion_ang", 3 * nD) CPF(m.d_motion_vel, "dofs_motion_vel", 3 * nD) CPF(m.d_limit, "dofs_limit", 2 * nD)
  CPF(m.d_stiffness, "dofs_stiffness", nD) CPF(m.d_damping, "dofs_damping", nD) CPF(m.d_armature, "dofs_armature", nD)
  CPF(m.d_frictionloss, "dofs_frictionloss", nD) CPF(m.d_kp, "dofs_kp", nD) CPF(m.d_act_bias, "dofs_act_bias", 3 * nD)
  CPF(m.d_force_range, "dofs_force_range", 2 * nD) CPF(m.d_invweight, "dofs_invweight", nD) CPF(m.qpos0, "qpos0", nQ)
  { const float* p = FV("mass_parent_mask", (size_t)nD * nD); if (!missing) for (int a = 0; a < nD; a++) for (int b = 0; b < nD; b++) m.mass_mask[a][b] = p[a * nD + b] != 0.f; }
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

int gsr_model_commit(gsr_handle_t h, int device, int n_envs) {
  if (!h || n_envs <= 0) return fail(h, GSR_EINVAL, "bad handle or n_envs");
  if (h->committed) return fail(h, GSR_ESTATE, "model already committed (use gsr_model_refresh)");
  memset(&h->hm, 0, sizeof(h->hm));
  int rc_fill = fill_model(h);
  if (rc_fill) return rc_fill;
  DevModel& m = h->hm;
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
  }
  m.verts = h->d_verts; m.support_v = h->d_support; m.pairs = h->d_pairs; m.pair_idx = h->d_pair_idx;

  // ---- state layout
  const int nL = m.nL, nJ = m.nJ, nD = m.nD, nQ = m.nQ, nG = m.nG;
  const int nC = m.len_constraints > 0 ? m.len_constraints : 1, nCP = m.max_contact_pairs > 0 ? m.max_contact_pairs : 1;
  const int nP = m.nP > 0 ? m.nP : 1, nW = m.n_cache_words;
  size_t off_f = 0, off_i = 0;
  h->fields.clear();
#define X(name, is_int, rows)                                                     \
  {                                                                               \
    size_t r_ = (size_t)(rows);                                                   \
    if (r_ == 0) r_ = 1;                                                          \
    size_t& off_ = (is_int) ? off_i : off_f;                                      \
    h->lay.name = (int)off_;                                                      \
    h->fields.push_back(FieldInfo{#name, is_int, off_, r_});                      \
    off_ += r_;                                                                   \
  }
  GSR_FIELDS(X)
#undef X
  (void)nP;
  h->n_f32 = off_f * (size_t)n_envs;
  h->n_i32 = off_i * (size_t)n_envs;
  if (off_f >= (size_t)1 << 31 || off_i >= (size_t)1 << 31) return fail(h, GSR_ELIMIT, "state layout exceeds int32 row index");
  h->committed = true;
  g_active[device] = nullptr;
  return GSR_OK;
}

int gsr_model_refresh(gsr_handle_t h) {
  if (!h || !h->committed) return fail(h, GSR_ESTATE, "model not committed");
  int rc = fill_model(h);
  if (rc) return rc;
  g_active[h->device] = nullptr;  // forces a constant-memory re-upload on the next call
  return GSR_OK;
}

int gsr_state_sizes(gsr_handle_t h, size_t* n_float32, size_t* n_int32) {
  if (!h || !h->committed) return fail(h, GSR_ESTATE, "model not committed");
  if (n_float32) *n_float32 = h->n_f32;
  if (n_int32) *n_int32 = h->n_i32;
  return GSR_OK;
}
int gsr_field_count(gsr_handle_t h) { return (h && h->committed) ? (int)h->fields.size() : GSR_ESTATE; }
int gsr_field_info(gsr_handle_t h, int index, const char** name, int* is_int, size_t* offset, size_t* items) {
  if (!h || !h->committed) return fail(h, GSR_ESTATE, "model not committed");
  if (index < 0 || index >= (int)h->fields.size()) return fail(h, GSR_EINVAL, "field index out of range");
  const FieldInfo& f = h->fields[index];
  if (name) *name = f.name.c_str();
  if (is_int) *is_int = f.is_int;
  if (offset) *offset = f.offset;
  if (items) *items = f.items;
  return GSR_OK;
}
int gsr_bind(gsr_handle_t h, float* state_f32, int32_t* state_i32) {
  if (!h || !h->committed) return fail(h, GSR_ESTATE, "model not committed");
  if (!state_f32 || !state_i32) return fail(h, GSR_EINVAL, "null state buffer");
  h->F = state_f32;
  h->I = state_i32;
  return GSR_OK;
}

static int launch_cfg(gsr_handle_t h, dim3& grid, dim3& block) {
  block = dim3(GSR_BLOCK);
  grid = dim3((h->B + GSR_BLOCK - 1) / GSR_BLOCK);
  return 0;
}

int gsr_init_state(gsr_handle_t h, void* cuda_stream) {
  if (!h || !h->F) return fail(h, GSR_ESTATE, "state not bound");
  cudaStream_t st = (cudaStream_t)cuda_stream;
  int rc = activate(h, st);
  if (rc) return rc;
  const DevModel& m = h->hm;
  const size_t B = h->B;
  CUDA_OK(cudaMemsetAsync(h->F, 0, h->n_f32 * sizeof(float), st));
  CUDA_OK(cudaMemsetAsync(h->I, 0, h->n_i32 * sizeof(int), st));
  // host-built initial rows: qpos0, FORCE control mode, unit friction ratio, model pose for every link
  std::vector<float> rows;
  auto put_f = [&](int off, int n, const float* vals) -> cudaError_t {
    rows.assign((size_t)n * B, 0.f);
    for (int r = 0; r < n; r++) for (size_t b = 0; b < B; b++) rows[r * B + b] = vals[r];
    return cudaMemcpyAsync(h->F + (size_t)off * B, rows.data(), rows.size() * 4, cudaMemcpyHostToDevice, st);
  };
  CUDA_OK(put_f(h->lay.qpos, m.nQ, m.qpos0));
  CUDA_OK(cudaStreamSynchronize(st));
  std::vector<float> ones(m.nG > 0 ? m.nG : 1, 1.f);
  CUDA_OK(put_f(h->lay.friction_ratio, m.nG, ones.data()));
  CUDA_OK(cudaStreamSynchronize(st));
  CUDA_OK(put_f(h->lay.links_pos, 3 * m.nL, &m.l_pos[0][0]));
  CUDA_OK(cudaStreamSynchronize(st));
  CUDA_OK(put_f(h->lay.links_quat, 4 * m.nL, &m.l_quat[0][0]));
  CUDA_OK(cudaStreamSynchronize(st));
  {
    std::vector<int> modes((size_t)(m.nD > 0 ? m.nD : 1) * B, CTRL_FORCE);  // abd/misc.py:170-173
    CUDA_OK(cudaMemcpyAsync(h->I + (size_t)h->lay.ctrl_mode * B, modes.data(), (size_t)m.nD * B * 4, cudaMemcpyHostToDevice, st));
    CUDA_OK(cudaStreamSynchronize(st));
  }
  dim3 grid, block;
  launch_cfg(h, grid, block);
  { Bracket br_(h, 4, st); k_forward<<<grid, block, 0, st>>>(h->F, h->I, h->B, nullptr, 1); }
  CUDA_OK(cudaGetLastError());
  return GSR_OK;
}

int gsr_step(gsr_handle_t h, int n_substeps, void* cuda_stream) {
  if (!h || !h->F) return fail(h, GSR_ESTATE, "state not bound");
  cudaStream_t st = (cudaStream_t)cuda_stream;
  int rc = activate(h, st);
  if (rc) return rc;
  dim3 grid, block;
  launch_cfg(h, grid, block);
  for (int s = 0; s < n_substeps; s++) {
    { Bracket br_(h, 0, st); k_dynamics<<<grid, block, 0, st>>>(h->F, h->I, h->B); }
    if (h->hm.enable_collision && h->hm.nG > 0) {
      { Bracket br_(h, 1, st); k_collide<<<grid, block, 0, st>>>(h->F, h->I, h->B); }
    }
    { Bracket br_(h, 2, st); k_constraint<<<grid, block, 0, st>>>(h->F, h->I, h->B); }
    { Bracket br_(h, 3, st); k_integrate<<<grid, block, 0, st>>>(h->F, h->I, h->B, s == n_substeps - 1); }
  }
  CUDA_OK(cudaGetLastError());
  return GSR_OK;
}

int gsr_forward_kinematics(gsr_handle_t h, const uint8_t* env_mask_dev, void* cuda_stream) {
  if (!h || !h->F) return fail(h, GSR_ESTATE, "state not bound");
  cudaStream_t st = (cudaStream_t)cuda_stream;
  int rc = activate(h, st);
  if (rc) return rc;
  dim3 grid, block;
  launch_cfg(h, grid, block);
  { Bracket br_(h, 4, st); k_forward<<<grid, block, 0, st>>>(h->F, h->I, h->B, env_mask_dev, 0); }
  CUDA_OK(cudaGetLastError());
  return GSR_OK;
}

int gsr_reset_envs(gsr_handle_t h, const uint8_t* env_mask_dev, void* cuda_stream) {
  if (!h || !h->F) return fail(h, GSR_ESTATE, "state not bound");
  cudaStream_t st = (cudaStream_t)cuda_stream;
  int rc = activate(h, st);
  if (rc) return rc;
  dim3 grid, block;
  launch_cfg(h, grid, block);
  { Bracket br_(h, 5, st); k_reset<<<grid, block, 0, st>>>(h->F, h->I, h->B, env_mask_dev); }
  CUDA_OK(cudaGetLastError());
  return GSR_OK;
}

int gsr_step_host(gsr_handle_t h, int n_substeps, const float* ctrl_pos_host, float* qpos_host, float* vel_host, void* cuda_stream) {
  if (!h || !h->F) return fail(h, GSR_ESTATE, "state not bound");
  cudaStream_t st = (cudaStream_t)cuda_stream;
  CUDA_OK(cudaSetDevice(h->device));
  const DevModel& m = h->hm;
  const size_t B = h->B;
  const size_t need = B * (size_t)(m.nQ > m.nD ? m.nQ : m.nD);
  if (h->stage_elems < need) {
    if (h->d_stage) cudaFree(h->d_stage);
    if (h->h_stage) cudaFreeHost(h->h_stage);
    CUDA_OK(cudaMalloc((void**)&h->d_stage, need * 4));
    CUDA_OK(cudaMallocHost((void**)&h->h_stage, need * 4));
    h->stage_elems = need;
  }
  // The public API is env-major (B, n) like the reference's transposed torch views; the device rows are item-major.
  // The transposes are done on the host through a pinned staging buffer.
  if (ctrl_pos_host) {
    for (size_t b = 0; b < B; b++)
      for (int d = 0; d < m.nD; d++) h->h_stage[(size_t)d * B + b] = ctrl_pos_host[b * m.nD + d];
    CUDA_OK(cudaMemcpyAsync(h->F + (size_t)h->lay.ctrl_pos * B, h->h_stage, (size_t)m.nD * B * 4, cudaMemcpyHostToDevice, st));
    CUDA_OK(cudaStreamSynchronize(st));
  }
  int rc = gsr_step(h, n_substeps, cuda_stream);
  if (rc) return rc;
  if (qpos_host) {
    CUDA_OK(cudaMemcpyAsync(h->h_stage, h->F + (size_t)h->lay.qpos * B, (size_t)m.nQ * B * 4, cudaMemcpyDeviceToHost, st));
    CUDA_OK(cudaStreamSynchronize(st));
    for (size_t b = 0; b < B; b++)
      for (int q = 0; q < m.nQ; q++) qpos_host[b * m.nQ + q] = h->h_stage[(size_t)q * B + b];
  }
  if (vel_host) {
    CUDA_OK(cudaMemcpyAsync(h->h_stage, h->F + (size_t)h->lay.vel * B, (size_t)m.nD * B * 4, cudaMemcpyDeviceToHost, st));
    CUDA_OK(cudaStreamSynchronize(st));
    for (size_t b = 0; b < B; b++)
      for (int d = 0; d < m.nD; d++) vel_host[b * m.nD + d] = h->h_stage[(size_t)d * B + b];
  }
  CUDA_OK(cudaStreamSynchronize(st));
  return GSR_OK;
}

int gsr_set_profiling(gsr_handle_t h, int enabled) {
  if (!h) return GSR_EINVAL;
  if (enabled) { for (int k = 0; k < 6; k++) { h->time_ms[k] = 0; h->timed[k] = 0; } }
  h->profiling = enabled != 0;
  return GSR_OK;
}
int gsr_kernel_time_ms(gsr_handle_t h, int index, double* total_ms, long long* timed_launches) {
  if (!h || index < 0 || index >= 6) return GSR_EINVAL;
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
  if (!h || index < 0 || index >= 6) return GSR_EINVAL;
  if (name) *name = h->stats[index].name;
  if (launches) *launches = h->stats[index].launches;
  return GSR_OK;
}

void gsr_destroy(gsr_handle_t h) {
  if (!h) return;
  if (h->committed) {
    cudaSetDevice(h->device);
    if (g_active[h->device] == h) g_active[h->device] = nullptr;
    cudaFree(h->d_verts); cudaFree(h->d_support); cudaFree(h->d_pairs); cudaFree(h->d_pair_idx);
    if (h->d_stage) cudaFree(h->d_stage);
    if (h->h_stage) cudaFreeHost(h->h_stage);
    for (auto& sp : h->spans) { cudaEventDestroy(sp.a); cudaEventDestroy(sp.b); }
    for (auto e : h->event_pool) cudaEventDestroy(e);
  }
  delete h;
}

}  // extern "C"
