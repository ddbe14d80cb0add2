// TEST INFRASTRUCTURE ONLY -- host emulation of the lane-per-environment step kernel (genesis_b200/csrc/gsr_lane.cuh).
//
// The lane kernel maps one thread to one environment and its device functions use no warp intrinsics, so the same source
// compiles for the host (GSR_LANE_EMU: stride-1 workspace, one "lane").  This library runs those functions environment by
// environment on the CPU so that the kernel's arithmetic can be checked against the oracle in a container without a GPU
// (tests/test_lane_emu.py).  It shares gsr_host.h (model filling + record layout) with libgsrigid.so and mirrors the subset
// of the C ABI the tests need.  It is never loaded by the genesis_b200 package: the product path stays CUDA-only.
#define GSR_LANE_EMU 1
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "../../genesis_b200/csrc/gsr_host.h"
#include "../../genesis_b200/csrc/gsr_lane.cuh"

struct gsr_handle_s : GsrSpec {
  bool committed = false;
  int B = 0;
  float* state = nullptr;
  std::vector<float> ws;
};

static void activate(gsr_handle_s* h) {
  // host shadows of the __constant__ symbols: the host versions of the lane functions read these
  memcpy((void*)&c_m, &h->hm, sizeof(DevModel));
  memcpy((void*)&c_h, &h->lay, sizeof(HotLayout));
  memcpy((void*)&c_l, &h->ll, sizeof(LaneLayout));
}

extern "C" {

int gsr_model_begin(gsr_handle_t* out) { *out = new gsr_handle_s(); return GSR_OK; }
int gsr_model_set_f32(gsr_handle_t h, const char* name, const float* data, size_t count) { h->f[name] = std::vector<float>(data, data + count); return GSR_OK; }
int gsr_model_set_i32(gsr_handle_t h, const char* name, const int32_t* data, size_t count) { h->i[name] = std::vector<int>(data, data + count); return GSR_OK; }
int gsr_model_set_scalar(gsr_handle_t h, const char* name, double value) { h->s[name] = value; return GSR_OK; }
const char* gsr_last_error(gsr_handle_t h) { return h ? h->err.c_str() : "null handle"; }

int gsr_model_commit(gsr_handle_t h, int, int n_envs) {
  memset(&h->hm, 0, sizeof(h->hm));
  int rc = gsr_fill_model(h);
  if (rc) return rc;
  DevModel& m = h->hm;
  m.verts = h->f["verts"].data(); m.support_v = h->f["support_v"].data();
  m.pairs = h->i["pairs"].data(); m.pair_idx = h->i["pair_idx"].data(); m.fac_tasks = h->fac_tasks.data();
  h->B = n_envs;
  h->s["cfg_group_size"] = 1;
  rc = gsr_compute_layouts(h, n_envs);
  if (rc) return rc;
  h->ws.assign((size_t)h->ll.env_words + 16, 0.f);
  h->committed = true;
  return GSR_OK;
}
int gsr_model_refresh(gsr_handle_t h) {
  const float* v = h->hm.verts; const float* sv = h->hm.support_v; const int* p = h->hm.pairs; const int* pi = h->hm.pair_idx;
  int rc = gsr_fill_model(h);
  h->hm.verts = v; h->hm.support_v = sv; h->hm.pairs = p; h->hm.pair_idx = pi;
  return rc;
}
int gsr_state_words(gsr_handle_t h, size_t* n) { *n = h->n_words; return GSR_OK; }
int gsr_field_count(gsr_handle_t h) { return (int)h->fields.size(); }
int gsr_field_info(gsr_handle_t h, int index, const char** name, int* is_int, size_t* segment_base, size_t* env_stride,
                   size_t* offset, size_t* items, size_t* comps, size_t* item_stride) {
  if (index < 0 || index >= (int)h->fields.size()) return GSR_EINVAL;
  const GsrFieldInfo& f = h->fields[index];
  *name = f.name.c_str(); *is_int = f.is_int; *segment_base = f.seg_base; *env_stride = f.env_stride; *offset = f.offset;
  *items = f.items; *comps = f.comps; *item_stride = f.item_stride;
  return GSR_OK;
}
int gsr_bind(gsr_handle_t h, void* state_words) { h->state = (float*)state_words; return GSR_OK; }
int gsr_launch_info(gsr_handle_t h, int* g, int* bytes, int* rows, int* contacts) {
  *g = 1; *bytes = h->ll.env_words * 4; *rows = h->ll.NCS; *contacts = h->ll.NCC;
  return GSR_OK;
}

// k_init of gsr_kernels.cuh, on the host
int gsr_init_state(gsr_handle_t h, void*) {
  activate(h);
  memset(h->state, 0, h->n_words * sizeof(float));
  const DevModel& m = h->hm;
  for (int env = 0; env < h->B; env++) {
    float* rec = h->state + (size_t)env * h->lay.hot_stride;
    int* reci = reinterpret_cast<int*>(rec);
    for (int q = 0; q < m.nQ; q++) rec[h->lay.qpos + q] = m.qpos0[q];
    for (int d = 0; d < m.nD; d++) reci[h->lay.ctrl_mode + d] = CTRL_FORCE;
    for (int g = 0; g < m.nG; g++) rec[h->lay.friction_ratio + g] = 1.0f;
    for (int l = 0; l < m.nL; l++) {
      for (int k = 0; k < 3; k++) rec[h->lay.links_pos + 3 * l + k] = m.l_pos[l][k];
      for (int k = 0; k < 4; k++) rec[h->lay.links_quat + 4 * l + k] = m.l_quat[l][k];
    }
  }
  return GSR_OK;
}

int gsr_step(gsr_handle_t h, int n_substeps, void*) {
  activate(h);
  for (int s = 0; s < n_substeps; s++) {
    for (int env = 0; env < h->B; env++) {
      LaneCtx C;
      C.S = h->ws.data(); C.Sw = C.S;
      C.rec = h->state + (size_t)env * h->lay.hot_stride; C.rec0 = C.rec;
      C.contacts = h->state + h->lay.contact_base + (size_t)env * h->lay.contact_stride;
      C.cache = h->state + h->lay.cache_base + (size_t)env * h->lay.cache_stride;
      C.spill = h->state + h->lay.spill_base + (size_t)env * h->lay.spill_stride;
      C.gm = &h->hm; C.n_live = 1; C.lane = 0; C.live = true;
      lane_step(C, s == n_substeps - 1);
    }
  }
  return GSR_OK;
}

void gsr_destroy(gsr_handle_t h) { delete h; }
}
