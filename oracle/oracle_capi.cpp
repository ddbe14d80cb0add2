// TEST INFRASTRUCTURE ONLY: C entry points of the CPU oracle (see rigid_oracle.hpp header).  Loaded through ctypes by
// tests/, __graft_entry__.smoke() and bench.py (cpu_baseline / --impl reference) -- never by the product path.
#include <cstdio>
#include <string>
#include <chrono>
#include "rigid_oracle.hpp"

using namespace orc;

struct Handle {
  int use_double = 0;
  Oracle<float> f;
  Oracle<double> d;
};

template <typename T>
static bool get_field(Oracle<T>& o, int b, const std::string& n, double* out) {
  Env<T>& e = o.envs[b];
  auto cp = [&](const T* p, size_t cnt) { for (size_t i = 0; i < cnt; i++) out[i] = (double)p[i]; return true; };
  auto cpv = [&](const std::vector<V3<T>>& v) { return cp(reinterpret_cast<const T*>(v.data()), v.size() * 3); };
  auto cpq = [&](const std::vector<Q4<T>>& v) { return cp(reinterpret_cast<const T*>(v.data()), v.size() * 4); };
#define VF(name) if (n == #name) return cp(e.name.data(), e.name.size());
  VF(qpos) VF(vel) VF(acc) VF(acc_smooth) VF(force) VF(qf_bias) VF(qf_passive) VF(qf_applied) VF(qf_smooth) VF(qf_constraint)
  VF(dofs_pos) VF(ctrl_pos) VF(ctrl_vel) VF(ctrl_force) VF(M) VF(L) VF(Dinv) VF(qacc) VF(qacc_ws) VF(cinr_mass) VF(diag) VF(aref)
  VF(efc_D) VF(efc_force) VF(Jaref) VF(grad) VF(search) VF(Ma)
#undef VF
  if (n == "links_pos") return cpv(e.l_pos);
  if (n == "links_quat") return cpq(e.l_quat);
  if (n == "root_COM") return cpv(e.root_COM);
  if (n == "i_pos") return cpv(e.i_pos);
  if (n == "cd_vel") return cpv(e.cd_vel);
  if (n == "cd_ang") return cpv(e.cd_ang);
  if (n == "cacc_lin") return cpv(e.cacc_lin);
  if (n == "cacc_ang") return cpv(e.cacc_ang);
  if (n == "cdof_ang") return cpv(e.cdof_ang);
  if (n == "cdof_vel") return cpv(e.cdof_vel);
  if (n == "cdofd_ang") return cpv(e.cdofd_ang);
  if (n == "cdofd_vel") return cpv(e.cdofd_vel);
  if (n == "geoms_pos") return cpv(e.g_pos);
  if (n == "geoms_quat") return cpq(e.g_quat);
  if (n == "aabb_min") return cpv(e.aabb_min);
  if (n == "aabb_max") return cpv(e.aabb_max);
  if (n == "links_contact_force") return cpv(e.contact_force);
  if (n == "contact_cache_normal") return cpv(e.contact_cache_normal);
  if (n == "jac") return cp(e.jac.data(), (size_t)e.n_con * o.m.nD);
  if (n == "ctrl_mode") { for (size_t i = 0; i < e.ctrl_mode.size(); i++) out[i] = e.ctrl_mode[i]; return true; }
  if (n == "n_contacts") { out[0] = e.n_contacts; return true; }
  if (n == "n_broad") { out[0] = e.n_broad; return true; }
  if (n == "n_constraints") { out[0] = e.n_con; out[1] = e.n_eq; out[2] = e.n_floss; return true; }
  if (n == "n_iter") { out[0] = e.n_iter; return true; }
  if (n == "errno") { out[0] = e.errno_; return true; }
  if (n == "is_warmstart") { out[0] = e.is_warmstart; return true; }
  if (n == "contacts") {  // 16 doubles per contact: ga gb la lb | pen | pos3 | normal3 | force3 | friction | pair
    for (int c = 0; c < e.n_contacts; c++) {
      const Contact<T>& ct = e.contacts[c];
      double* r = out + 16 * c;
      r[0] = ct.geom_a; r[1] = ct.geom_b; r[2] = ct.link_a; r[3] = ct.link_b; r[4] = ct.penetration;
      for (int k = 0; k < 3; k++) { r[5 + k] = ct.pos[k]; r[8 + k] = ct.normal[k]; r[11 + k] = ct.force[k]; }
      r[14] = ct.friction; r[15] = ct.pair_idx;
    }
    return true;
  }
  return false;
}

template <typename T>
static bool set_field(Oracle<T>& o, int b, const std::string& n, const double* in) {
  Env<T>& e = o.envs[b];
  auto cp = [&](std::vector<T>& v) { for (size_t i = 0; i < v.size(); i++) v[i] = (T)in[i]; return true; };
  if (n == "qpos") { cp(e.qpos); o.forward_all(e); return true; }
  if (n == "vel") { cp(e.vel); o.forward_all(e); return true; }
  if (n == "ctrl_pos") return cp(e.ctrl_pos);
  if (n == "ctrl_vel") return cp(e.ctrl_vel);
  if (n == "ctrl_force") return cp(e.ctrl_force);
  if (n == "friction_ratio") return cp(e.friction_ratio);
  if (n == "ctrl_mode") { for (size_t i = 0; i < e.ctrl_mode.size(); i++) e.ctrl_mode[i] = (int)in[i]; return true; }
  if (n == "cfrc_applied_vel") { for (size_t i = 0; i < e.cfrc_applied_vel.size(); i++) for (int k = 0; k < 3; k++) e.cfrc_applied_vel[i][k] = (T)in[3 * i + k]; return true; }
  if (n == "cfrc_applied_ang") { for (size_t i = 0; i < e.cfrc_applied_ang.size(); i++) for (int k = 0; k < 3; k++) e.cfrc_applied_ang[i][k] = (T)in[3 * i + k]; return true; }
  if (n == "mass_shift") { cp(e.mass_shift); o.forward_all(e); return true; }
  if (n == "com_shift") { for (size_t i = 0; i < e.i_pos_shift.size(); i++) for (int k = 0; k < 3; k++) e.i_pos_shift[i][k] = (T)in[3 * i + k]; o.forward_all(e); return true; }
  if (n == "add_weld") {  // [link1, link2]: the reference's kernel_add_weld_constraint at the current poses
    typename Env<T>::DynWeld w;
    w.l1 = (int)in[0]; w.l2 = (int)in[1];
    auto shared = e.l_pos[w.l1];
    auto p1 = inv_transform_by_quat(shared - e.l_pos[w.l1], e.l_quat[w.l1]);
    auto p2 = inv_transform_by_quat(shared - e.l_pos[w.l2], e.l_quat[w.l2]);
    for (int k = 0; k < 3; k++) { w.data[k] = p2[k]; w.data[3 + k] = p1[k]; }
    auto rel = quat_mul(inv_quat(e.l_quat[w.l1]), e.l_quat[w.l2]);
    w.data[6] = rel.w; w.data[7] = rel.x; w.data[8] = rel.y; w.data[9] = rel.z; w.data[10] = 1;
    const T sol[7] = {T(2) * (T)o.cfg.dt, 1, T(0.9), T(0.95), T(0.001), T(0.5), 2};
    for (int k = 0; k < 7; k++) w.sol[k] = sol[k];
    e.dyn_welds.push_back(w);
    return true;
  }
  if (n == "clear_welds") { e.dyn_welds.clear(); return true; }
  if (n == "reset") { o.reset_env(e); return true; }
  return false;
}

static void load_cfg(const RawModel& r, Config& c) {
  auto S = [&](const char* k, double d) { auto it = r.s.find(k); return it == r.s.end() ? d : it->second; };
  c.dt = S("dt", c.dt);
  c.gravity[0] = S("gravity_x", 0.0); c.gravity[1] = S("gravity_y", 0.0); c.gravity[2] = S("gravity_z", -9.81);
  c.iterations = (int)S("cfg_iterations", c.iterations); c.ls_iterations = (int)S("cfg_ls_iterations", c.ls_iterations);
  c.tolerance = S("cfg_tolerance", c.tolerance); c.ls_tolerance = S("cfg_ls_tolerance", c.ls_tolerance);
  c.integrator = (int)S("cfg_integrator", c.integrator); c.solver_type = (int)S("cfg_solver_type", c.solver_type);
  c.enable_collision = (int)S("cfg_enable_collision", 1); c.enable_joint_limit = (int)S("cfg_enable_joint_limit", 1);
  c.enable_multi_contact = (int)S("cfg_enable_multi_contact", 1); c.sort_contacts = (int)S("cfg_sort_contacts", 1);
  c.eps = S("cfg_eps", c.eps);
  c.gjk_fallback = (int)S("cfg_gjk_fallback", 1);
  c.broadphase_sap = (int)S("cfg_broadphase_sap", 0);
  c.box_box_detection = (int)S("cfg_box_box_detection", 0);
  c.use_gjk_collision = (int)S("cfg_use_gjk_collision", 0);
  c.noslip_iterations = (int)S("cfg_noslip_iterations", 0); c.noslip_tolerance = S("cfg_noslip_tolerance", c.noslip_tolerance);
}

extern "C" {

void* orc_model_new() { return new RawModel(); }
void orc_model_free(void* p) { delete (RawModel*)p; }
void orc_model_set_f(void* p, const char* name, const double* data, int n) { ((RawModel*)p)->f[name] = std::vector<double>(data, data + n); }
void orc_model_set_i(void* p, const char* name, const int* data, int n) { ((RawModel*)p)->i[name] = std::vector<int>(data, data + n); }
void orc_model_set_s(void* p, const char* name, double v) { ((RawModel*)p)->s[name] = v; }

void* orc_create(void* raw, int B, int use_double) {
  Handle* h = new Handle();
  h->use_double = use_double;
  const RawModel& r = *(RawModel*)raw;
  if (use_double) { h->d.m.load(r); load_cfg(r, h->d.cfg); h->d.init(B); }
  else { h->f.m.load(r); load_cfg(r, h->f.cfg); h->f.init(B); }
  return h;
}
void orc_destroy(void* p) { delete (Handle*)p; }

int orc_get(void* p, int b, const char* name, double* out) {
  Handle* h = (Handle*)p;
  return (h->use_double ? get_field(h->d, b, name, out) : get_field(h->f, b, name, out)) ? 0 : -1;
}
int orc_set(void* p, int b, const char* name, const double* in) {
  Handle* h = (Handle*)p;
  return (h->use_double ? set_field(h->d, b, name, in) : set_field(h->f, b, name, in)) ? 0 : -1;
}
void orc_set_threads(void* p, int n) {
  Handle* h = (Handle*)p;
  h->d.n_threads = n; h->f.n_threads = n;
}
void orc_step(void* p, int n) {
  Handle* h = (Handle*)p;
  if (h->use_double) h->d.step(n); else h->f.step(n);
}
// run n steps and return elapsed wall seconds (used by bench.py's cpu_baseline leg)
double orc_step_timed(void* p, int n) {
  auto t0 = std::chrono::steady_clock::now();
  orc_step(p, n);
  return std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
}
// single pieces, for unit tests of the smooth dynamics
void orc_forward_dynamics(void* p, int b) {
  Handle* h = (Handle*)p;
  if (h->use_double) h->d.forward_dynamics(h->d.envs[b]); else h->f.forward_dynamics(h->f.envs[b]);
}
void orc_detection(void* p, int b) {
  Handle* h = (Handle*)p;
  if (h->use_double) h->d.detection(h->d.envs[b]); else h->f.detection(h->f.envs[b]);
}
}
