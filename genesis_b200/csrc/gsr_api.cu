// gsr_api.cu -- C ABI of libgsrigid.so (see include/gsrigid.h for the contract and the reference interfaces replaced).
#include <cuda_runtime.h>
#include <cstdio>
#include <cstring>
#include <map>
#include <string>
#include <vector>

#include "../../include/gsrigid.h"
#include <cstdint>
#include <cub/cub.cuh>

#include "gsr_host.h"
#include "gsr_kernels.cuh"
#ifndef GSR_DEV_ONLY_G  // developer builds (-DGSR_DEV_ONLY_G=32): one group size, no lane kernel -- a third of the compile time
#include "gsr_lane.cuh"
#endif

namespace {

struct KernelStat {
  const char* name;
  long long launches;
};

enum { K_STEP = 0, K_FORWARD = 1, K_RESET = 2, K_INIT = 3, K_SCATTER = 4, K_GATHER = 5, K_GATHER_COST = 6, K_ORDER_FLAG = 7, K_DIAG = 8, K_SET_STATE = 9, K_GATHER_FIELDS = 10, K_COUNT = 11 };

}  // namespace

struct gsr_handle_s : GsrSpec {
  bool committed = false;
  int device = 0, B = 0;
  float* state = nullptr;
  DevModel* d_model = nullptr;  // global-memory copy for lane-varying table reads
  float *d_verts = nullptr, *d_support = nullptr, *d_terrain = nullptr;
  int* d_support_vid = nullptr;
  unsigned short* d_fac = nullptr;
  unsigned long long* d_trace = nullptr;  // gsr_debug_trace
  // cost-ordered scheduling of the lane-group step kernel: every `reorder_period` steps the environments are sorted by the
  // busy time they measured in their last step (descending), so the expensive ones start first and the tail of a launch
  // overlaps the bulk of the batch instead of running alone on a few SMs (profiles/README.md "tail")
  int *d_order = nullptr, *d_iota = nullptr, *d_cost = nullptr, *d_cost_sorted = nullptr;
  void* d_sort_tmp = nullptr;
  size_t sort_tmp_bytes = 0;
  int reorder_period = 16, steps_since_reorder = 0;
  int auto_warps = 0;  // warps per CTA chosen by auto_warps_per_cta (0 = not chosen yet)
  int *d_pairs = nullptr, *d_pair_idx = nullptr;
  // staging for gsr_step_host
  float *d_stage = nullptr, *h_stage = nullptr;
  size_t stage_elems = 0;
  // adaptive on-chip row capacity (reorder_envs): histogram of rows per environment, read back one period later
  int *d_rows_hist = nullptr, *h_rows_hist = nullptr;
  cudaEvent_t rows_event = nullptr;
  bool rows_pending = false, adaptive_rows = true;
  int rows_default = 0, rows_changes = 0;
  // kinematic-cache policy (gsr_set_kinematics_policy): eager = every step leaves a valid cache (reference behaviour, default);
  // lazy = steps recompute it on chip and never store it; cache_valid tracks what the records currently hold
  bool eager_fk = true, cache_valid = false;
  int host_io_mode = 2;  // cfg_host_io: 0 = staged copies + k_scatter / k_gather, 1 = k_step reads / writes dense device staging, 2 = 1 + results posted straight into page-locked caller buffers
  KernelStat stats[K_COUNT] = {{"k_step", 0}, {"k_forward", 0}, {"k_reset", 0}, {"k_init", 0}, {"k_scatter", 0}, {"k_gather", 0},
                              {"k_gather_cost", 0}, {"k_order_flag", 0}, {"k_diag", 0}, {"k_set_state", 0}, {"k_gather_fields", 0}};
  long long total_launches = 0;
  // optional per-kernel event timing (gsr_set_profiling)
  bool profiling = false;
  struct Span { int k; cudaEvent_t a, b; };
  std::vector<Span> spans;
  std::vector<cudaEvent_t> event_pool;
  double time_ms[K_COUNT] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0};
  long long timed[K_COUNT] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0};
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
#ifdef GSR_SPEC_HEADER
    CUDA_OK(cudaMemcpyToSymbolAsync(c_h_rt, &h->lay, sizeof(HotLayout), 0, cudaMemcpyHostToDevice, st));
    CUDA_OK(cudaMemcpyToSymbolAsync(c_s_rt, &h->sl, sizeof(SmemLayout), 0, cudaMemcpyHostToDevice, st));
#else
    CUDA_OK(cudaMemcpyToSymbolAsync(c_h, &h->lay, sizeof(HotLayout), 0, cudaMemcpyHostToDevice, st));
    CUDA_OK(cudaMemcpyToSymbolAsync(c_s, &h->sl, sizeof(SmemLayout), 0, cudaMemcpyHostToDevice, st));
#endif
#ifndef GSR_DEV_ONLY_G
    CUDA_OK(cudaMemcpyToSymbolAsync(c_l, &h->ll, sizeof(LaneLayout), 0, cudaMemcpyHostToDevice, st));
#endif
    CUDA_OK(cudaMemcpyAsync(h->d_model, &h->hm, sizeof(DevModel), cudaMemcpyHostToDevice, st));
    g_active[h->device] = h;
  }
  return GSR_OK;
}

#define GSR_ROWS_BINS 32  // rows per environment in buckets of 8 (the last bucket collects everything above 240)
#define GSR_ROWS_GROW_MAX 96  // adapt_rows never raises the on-chip row capacity beyond this (see there)
__global__ void k_iota(int* v, int n) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t < n) v[t] = t;
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
  int rc_fill = gsr_fill_model(h);
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
    auto& svid = h->i["support_vid"]; CUDA_OK(upload(svid.empty() ? nullptr : svid.data(), svid.size() * 4, (void**)&h->d_support_vid));
    auto& thf = h->f["terrain_hf"]; CUDA_OK(upload(thf.empty() ? nullptr : thf.data(), thf.size() * 4, (void**)&h->d_terrain));
    auto& pr = h->i["pairs"]; CUDA_OK(upload(pr.empty() ? nullptr : pr.data(), pr.size() * 4, (void**)&h->d_pairs));
    auto& pi = h->i["pair_idx"]; CUDA_OK(upload(pi.empty() ? nullptr : pi.data(), pi.size() * 4, (void**)&h->d_pair_idx));
    CUDA_OK(upload(h->fac_tasks.empty() ? nullptr : h->fac_tasks.data(), h->fac_tasks.size() * sizeof(unsigned short), (void**)&h->d_fac));
    CUDA_OK(cudaMalloc((void**)&h->d_model, sizeof(DevModel)));
  }
  m.fac_tasks = h->d_fac;
  m.verts = h->d_verts; m.support_v = h->d_support; m.support_vid = h->d_support_vid; m.pairs = h->d_pairs; m.pair_idx = h->d_pair_idx;
  m.terrain_hf = h->d_terrain;

  { int rc_lay = gsr_compute_layouts(h, n_envs); if (rc_lay) return rc_lay; }
#ifdef GSR_SPEC_HEADER
  if (!gsr_spec_matches(h->lay, h->sl, h->hm) || h->group != GSR_SPEC_GROUP)
    return fail(h, GSR_EINVAL, "this copy of the library was specialised for another model / option set (tools/specialize.py)");
#endif
  h->reorder_period = (int)h->S("cfg_reorder_period", 16);
  h->host_io_mode = (int)h->S("cfg_host_io", 2);
  // the caller fixing cfg_smem_rows switches the adaptation off (sweeps); the lane kernel has its own layout
  h->adaptive_rows = h->S("cfg_adaptive_rows", h->s.count("cfg_smem_rows") ? 0 : 1) != 0 && h->group != 1;
#ifdef GSR_SPEC_HEADER
  h->adaptive_rows = false;  // the workspace layout is part of the build
#endif
  // every buffer of the cost-ordered launch, the adaptive row capacity and gsr_step_host is allocated here: nothing on the
  // step path calls cudaMalloc / cudaHostAlloc / cudaEventCreate (those serialise the device and used to land inside timed loops)
  if (h->group != 1) {
    const int B = n_envs;
    CUDA_OK(cudaMalloc((void**)&h->d_order, sizeof(int) * (B + 1))); CUDA_OK(cudaMalloc((void**)&h->d_iota, sizeof(int) * B));
    CUDA_OK(cudaMalloc((void**)&h->d_cost, sizeof(int) * B)); CUDA_OK(cudaMalloc((void**)&h->d_cost_sorted, sizeof(int) * B));
    CUDA_OK(cub::DeviceRadixSort::SortPairsDescending(nullptr, h->sort_tmp_bytes, h->d_cost, h->d_cost_sorted, h->d_iota, h->d_order, B, 0, 32, (cudaStream_t)0));
    CUDA_OK(cudaMalloc(&h->d_sort_tmp, h->sort_tmp_bytes));
    k_iota<<<(B + 255) / 256, 256>>>(h->d_iota, B);
    k_iota<<<(B + 255) / 256, 256>>>(h->d_order, B);
    CUDA_OK(cudaMemset(h->d_order + B, 0, sizeof(int)));  // flag 0 = identity order until the first sort
    CUDA_OK(cudaMalloc((void**)&h->d_rows_hist, sizeof(int) * GSR_ROWS_BINS));
    CUDA_OK(cudaHostAlloc((void**)&h->h_rows_hist, sizeof(int) * GSR_ROWS_BINS, cudaHostAllocDefault));
    CUDA_OK(cudaEventCreateWithFlags(&h->rows_event, cudaEventDisableTiming));
    h->rows_default = h->sl.NCS;
  }
  {
    const size_t need = (size_t)n_envs * (size_t)(m.nQ + 2 * m.nD);
    CUDA_OK(cudaMalloc((void**)&h->d_stage, need * 4));
    h->stage_elems = need;
  }
  CUDA_OK(cudaDeviceSynchronize());
  h->committed = true;
  g_active[device] = nullptr;
  return GSR_OK;
}

// ---- model-specialised builds (tools/specialize.py) ------------------------------------------------------------------
// Host-only: fills the model and the layouts from the arrays / scalars set so far (no CUDA call) and writes the header a
// specialised build includes: every layout offset and every build-time count / option of the model as a literal.
long long gsr_model_spec_header(gsr_handle_t h, int n_envs, char* buf, size_t cap) {
  if (!h) return GSR_EINVAL;
  int rc = gsr_fill_model(h);
  if (rc) return rc;
  rc = gsr_compute_layouts(h, n_envs);
  if (rc) return rc;
  std::string o = "// GENERATED by gsr_model_spec_header (tools/specialize.py): layouts and build-time scalars of ONE model as literals\n#pragma once\n";
  o += "struct SpecH {\n";
#define X(name, is_int, items, comps) o += "  static constexpr int " #name " = " + std::to_string(h->lay.name) + ";\n";
  GSR_HOT_PRE_FIELDS(X) GSR_HOT_IN_FIELDS(X) GSR_HOT_OUT_FIELDS(X)
#undef X
#define X(name) o += "  static constexpr int " #name " = " + std::to_string(h->lay.name) + ";\n";
  GSR_HOT_SCALAR_FIELDS(X)
#undef X
  o += "};\nstruct SpecS {\n";
#define X(name) o += "  static constexpr int " #name " = " + std::to_string(h->sl.name) + ";\n";
  GSR_SMEM_FIELDS(X)
#undef X
  o += "};\n";
#define X(name) o += "#define CMX_" #name " " + std::to_string(h->hm.name) + "\n";
  GSR_MODEL_SPEC_SCALARS(X)
#undef X
  o += "#define GSR_SPEC_GROUP " + std::to_string(h->group) + "\n";
  o += "static inline bool gsr_spec_matches(const HotLayout& L, const SmemLayout& S, const DevModel& m) {\n  return true";
#define X(name, is_int, items, comps) o += " && L." #name " == SpecH::" #name;
  GSR_HOT_PRE_FIELDS(X) GSR_HOT_IN_FIELDS(X) GSR_HOT_OUT_FIELDS(X)
#undef X
#define X(name) o += " && L." #name " == SpecH::" #name;
  GSR_HOT_SCALAR_FIELDS(X)
#undef X
#define X(name) o += " && S." #name " == SpecS::" #name;
  GSR_SMEM_FIELDS(X)
#undef X
#define X(name) o += " && m." #name " == " + std::to_string(h->hm.name);
  GSR_MODEL_SPEC_SCALARS(X)
#undef X
  o += ";\n}\n";
  if (buf && cap > o.size()) memcpy(buf, o.c_str(), o.size() + 1);
  return (long long)o.size() + 1;
}
/* 1 when this library was built for one model (GSR_SPEC_HEADER): gsr_model_commit then refuses any other */
int gsr_is_specialised(void) {
#ifdef GSR_SPEC_HEADER
  return 1;
#else
  return 0;
#endif
}

int gsr_model_refresh(gsr_handle_t h) {
  if (!h || !h->committed) return fail(h, GSR_ESTATE, "model not committed");
  int rc = gsr_fill_model(h);
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
  const GsrFieldInfo& f = h->fields[index];
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
  const bool lane = h->group == 1;
  if (group_size) *group_size = h->group;
  if (smem_bytes_per_env) *smem_bytes_per_env = lane ? h->ll.env_words * 4 : h->sl.env_words * 4;
  if (smem_rows) *smem_rows = lane ? h->ll.NCS : h->sl.NCS;
  if (smem_contacts) *smem_contacts = lane ? h->ll.NCC : h->sl.NCC;
  return GSR_OK;
}

}  // extern "C"

// ---- launches ---------------------------------------------------------------------------------------------------------
template <typename K>
static cudaError_t set_smem(K kernel, size_t bytes) {
  return cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
}
// lane-per-environment step kernel: one warp = 32 environments, interleaved workspace (gsr_lane.cuh)
static int lane_launch(gsr_handle_t h, cudaStream_t st, int clear_ext) {
#ifdef GSR_DEV_ONLY_G
  return fail(h, GSR_ESTATE, "developer build without the lane kernel");
#else
  const size_t warp_bytes = (size_t)h->ll.env_words * GSR_LANE_STRIDE_DEV * sizeof(float);
  if (warp_bytes > 227 * 1024) return fail(h, GSR_ELIMIT, "lane kernel: per-warp workspace exceeds 227 KB (use cfg_group_size 8/16/32)");
  int warps = h->warps_per_cta > 0 ? h->warps_per_cta : 1;  // independent warps: small CTAs balance best
  if (warps > 4) warps = 4;
  while (warps > 1 && warps * warp_bytes > 227 * 1024) warps--;
  const size_t smem = warps * warp_bytes;
  const int n_warps = (h->B + 31) / 32;
  dim3 grid((n_warps + warps - 1) / warps), block(warps * 32);
  CUDA_OK(set_smem(k_step_lane, smem));
  {
    Bracket br_(h, K_STEP, st);
    k_step_lane<<<grid, block, smem, st>>>(h->state, h->B, h->d_model, clear_ext);
  }
  CUDA_OK(cudaGetLastError());
  return GSR_OK;
#endif
}

// order[B] = 1 when the most expensive environment costs at least 1.5x the median one
__global__ void k_order_flag(const int* cost_sorted, int* order, int B) {
  if (threadIdx.x == 0 && blockIdx.x == 0) order[B] = (2 * (long long)cost_sorted[0] > 3 * (long long)cost_sorted[B / 2]) ? 1 : 0;
}
__global__ void k_gather_cost(const float* state, int B, int field_off, int* dst, int* rows_hist) {
  const int env = blockIdx.x * blockDim.x + threadIdx.x;
  if (env >= B) return;
  const int* rec = reinterpret_cast<const int*>(state + (size_t)env * c_h.hot_stride);
  dst[env] = rec[field_off];
  if (rows_hist) atomicAdd(rows_hist + min((rec[c_h.n_constraints] + 7) >> 3, GSR_ROWS_BINS - 1), 1);
}
// sort the environments by their last measured step cost, most expensive first (two tiny kernels + one radix sort)
static int activate(gsr_handle_t h, cudaStream_t st);
// Row capacity of the shared-memory workspace follows the workload: when a quarter of the environments carry more
// constraint rows than fit on chip (their rows then live in the SPILL record and every solver access is an L2 round trip:
// grasp cluster, 65 rows against 32, measured 4.7 M -> 5.5 M env-steps/s with 72 rows on chip) the capacity is raised to
// cover 97 % of them, and lowered again when the batch calms down.  Only the shared-memory layout changes (the records
// in HBM do not), between two launches, from a histogram that is one reorder period old -- no host synchronisation.
static int adapt_rows(gsr_handle_t h, cudaStream_t st) {
  const int B = h->B, cur = h->sl.NCS, lenC = h->hm.len_constraints;
  int cum = 0, b97 = GSR_ROWS_BINS - 1, spilled = 0;
  for (int b = 0; b < GSR_ROWS_BINS; b++) if (8 * b > cur) spilled += h->h_rows_hist[b];
  for (int b = 0; b < GSR_ROWS_BINS; b++) { cum += h->h_rows_hist[b]; if (100LL * cum >= 97LL * B) { b97 = b; break; } }
  int want = 8 * b97;
  if (b97 == GSR_ROWS_BINS - 1) want = lenC;  // beyond the histogram: out of reach (below)
  if (want < h->rows_default) want = h->rows_default;
  if (want > lenC) want = lenC;
  const size_t env_bytes_now = (size_t)h->sl.env_words * 4;
  const size_t env_bytes_want = env_bytes_now + (size_t)(want - cur) * (h->sl.JS + 8 + GSR_CONTACT_WORDS / 4) * 4;
  const size_t per_warp = env_bytes_want * (32 / h->group);
  // Raising the capacity trades residency for on-chip rows; it pays while the rows stay moderate (72 rows, above) and loses
  // badly when they do not: grasp with three boxes, 200 rows per environment -- 232 rows on chip 1.48 M env-steps/s against
  // 2.44 M with the default 32 (every environment in the SPILL record, but four times the residency; 64 / 96 / 128 rows:
  // 2.16 / 1.97 / 1.49 M).  So: never beyond GSR_ROWS_GROW_MAX rows, and never below four warps per SM.
  const bool grow = want > cur && want <= GSR_ROWS_GROW_MAX && 4LL * spilled >= B && per_warp * 4 <= 220 * 1024;
  bool shrink = want + 16 <= cur;
  if (want > GSR_ROWS_GROW_MAX && cur > h->rows_default) { want = h->rows_default; shrink = true; }  // out of reach: back to full residency
  if (!grow && !shrink) return GSR_OK;
  h->s["cfg_smem_rows"] = want;
  int rc = gsr_compute_layouts(h, B);
  if (rc) return rc;
  h->auto_warps = 0;
  h->rows_changes++;
  g_active[h->device] = nullptr;  // re-upload the layouts (stream-ordered after the launches that used the old ones)
  return activate(h, st);
}

static int reorder_envs(gsr_handle_t h, cudaStream_t st) {
  const int B = h->B;
  if (h->adaptive_rows) {
    const bool hist_ready = h->rows_pending && cudaEventQuery(h->rows_event) == cudaSuccess;
    cudaGetLastError();  // "not ready" is not an error
    if (hist_ready) {
      h->rows_pending = false;
      int rc = adapt_rows(h, st);
      if (rc) return rc;
    }
    if (!h->rows_pending) CUDA_OK(cudaMemsetAsync(h->d_rows_hist, 0, sizeof(int) * GSR_ROWS_BINS, st));
  }
  const bool take_hist = h->adaptive_rows && !h->rows_pending;
  { Bracket br_(h, K_GATHER_COST, st); k_gather_cost<<<(B + 255) / 256, 256, 0, st>>>(h->state, B, h->lay.step_cost, h->d_cost, take_hist ? h->d_rows_hist : nullptr); }
  if (take_hist) {
    CUDA_OK(cudaMemcpyAsync(h->h_rows_hist, h->d_rows_hist, sizeof(int) * GSR_ROWS_BINS, cudaMemcpyDeviceToHost, st));
    CUDA_OK(cudaEventRecord(h->rows_event, st));
    h->rows_pending = true;
  }
  CUDA_OK(cub::DeviceRadixSort::SortPairsDescending(h->d_sort_tmp, h->sort_tmp_bytes, h->d_cost, h->d_cost_sorted, h->d_iota, h->d_order, B, 0, 32, st));
  { Bracket br_(h, K_ORDER_FLAG, st); k_order_flag<<<1, 32, 0, st>>>(h->d_cost_sorted, h->d_order, B); }
  CUDA_OK(cudaGetLastError());
  return GSR_OK;
}

// Warps per CTA of the full-warp (G = 32) step kernel when the caller does not fix it.  All warps of a CTA walk the phases
// together, a CTA needs w x (workspace + mbarrier) + 1 KB of the SM's 228 KB and the 128-register kernel allows 16 warps per
// SM, so w decides how many environments are resident (r = CTAs x w) and into how many waves the batch falls.  Measured on
// B200 (profiles/README.md: Go2 w = 7 vs 8, grasp cluster w = 7 vs 11): a wave costs about 1 + 0.35 r / 16 (more resident
// warps hide more latency but share the issue slots), and partial waves cost a whole one.
static int auto_warps_per_cta(const gsr_handle_t h, size_t warp_bytes) {
  int dev = 0, n_sm = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, dev);
  int best_w = 1;
  double best = 1e30;
  for (int w = 1; w <= GSR_MAX_WARPS_PER_CTA; w++) {
    const size_t cta = (size_t)w * (warp_bytes + 8);
    if (cta > 227 * 1024) break;
    int ctas = (int)((size_t)233472 / (cta + 1024));
    if (ctas > 16 / w) ctas = 16 / w;
    if (ctas < 1) continue;
    const int r = ctas * w;
    const long long per_wave = (long long)n_sm * r;
    const double waves = (double)((h->B + per_wave - 1) / per_wave);
    // ties go to mid-sized CTAs: 1-2 warp CTAs drift apart in the 20 k-instruction kernel (instruction cache, Go2: 11.5 M with
    // 1 warp against 12.4 M with 8), 8+ warps wait for each other at the phase barriers (grasp cluster, 72 rows on chip: 5.49 M
    // with 4 warps against 5.34 M with 8)
    const double cost = waves * (1.0 + 0.35 * r / 16.0) * (1.0 + 0.002 * (w > 6 ? w - 6 : 6 - w));
    if (cost < best - 1e-9) { best = cost; best_w = w; }
  }
  return best_w;
}

static int group_launch(gsr_handle_t h, int which, cudaStream_t st, const uint8_t* mask, int clear_ext, HostIO io = HostIO{nullptr, nullptr, nullptr},
                        StateIO sio = StateIO{nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr}, int fk_flags = 2) {
  if (h->group == 1 && which == K_STEP) return lane_launch(h, st, clear_ext);
  const int G = h->group == 1 ? 8 : h->group, EPW = 32 / G;  // the (cold) forward-kinematics paths of a lane model use the group kernel
  const size_t warp_bytes = (size_t)EPW * h->sl.env_words * sizeof(float);
  if (h->auto_warps == 0) h->auto_warps = G == 32 ? auto_warps_per_cta(h, warp_bytes) : 1;  // G < 32: 1-warp CTAs measured best
  int warps = h->warps_per_cta > 0 ? h->warps_per_cta : h->auto_warps;
  if (warps > GSR_MAX_WARPS_PER_CTA) warps = GSR_MAX_WARPS_PER_CTA;
  if (G < 32 && which == K_STEP && warps > GSR_DEV_SMALL_WARPS) warps = GSR_DEV_SMALL_WARPS;  // the small-group step kernels are compiled for one-warp CTAs (launch bounds, gsr_kernels.cuh)
  while (warps > 1 && warps * (warp_bytes + 8) > 224 * 1024) warps--;  // one big CTA per SM, its warps in phase lockstep
  const int envs_per_cta = warps * EPW;
  const size_t smem = (size_t)warps * warp_bytes + (size_t)warps * 8;  // + one mbarrier per warp (TMA staging)
  if (smem > 227 * 1024) return fail(h, GSR_ELIMIT, "per-warp shared-memory workspace exceeds 227 KB (lower cfg_smem_rows or raise cfg_group_size)");
  dim3 grid((h->B + envs_per_cta - 1) / envs_per_cta), block(warps * 32);
#define LAUNCH(GG)                                                                                       \
  if (which == K_STEP) {                                                                                 \
    if (h->hm.has_box_box) {                                                                             \
      CUDA_OK((set_smem(k_step<GG, true, true>, smem)));                                                 \
      Bracket br_(h, K_STEP, st);                                                                        \
      k_step<GG, true, true><<<grid, block, smem, st>>>(h->state, h->B, h->d_model, clear_ext, h->d_order,  h->d_trace, io, fk_flags);      \
    } else if (h->hm.has_mpr) {                                                                          \
      CUDA_OK((set_smem(k_step<GG, true>, smem)));                                                       \
      Bracket br_(h, K_STEP, st);                                                                        \
      k_step<GG, true><<<grid, block, smem, st>>>(h->state, h->B, h->d_model, clear_ext, h->d_order,  h->d_trace, io, fk_flags);            \
    } else {                                                                                             \
      CUDA_OK((set_smem(k_step<GG, false>, smem)));                                                      \
      Bracket br_(h, K_STEP, st);                                                                        \
      k_step<GG, false><<<grid, block, smem, st>>>(h->state, h->B, h->d_model, clear_ext, h->d_order,  h->d_trace, io, fk_flags);           \
    }                                                                                                    \
  } else if (which == K_SET_STATE) {                                                                     \
    CUDA_OK(set_smem(k_set_state<GG>, smem));                                                            \
    Bracket br_(h, K_SET_STATE, st);                                                                     \
    k_set_state<GG><<<grid, block, smem, st>>>(h->state, h->B, h->d_model, mask, sio, clear_ext);        \
  } else if (which == K_DIAG) {                                                                          \
    CUDA_OK(set_smem(k_diag<GG>, smem));                                                                 \
    Bracket br_(h, K_DIAG, st);                                                                          \
    k_diag<GG><<<grid, block, smem, st>>>(h->state, h->B, h->d_model);                                   \
  } else {                                                                                               \
    CUDA_OK(set_smem(k_forward<GG>, smem));                                                              \
    Bracket br_(h, K_FORWARD, st);                                                                       \
    k_forward<GG><<<grid, block, smem, st>>>(h->state, h->B, h->d_model, mask, clear_ext);               \
  }
#ifdef GSR_DEV_ONLY_G
  if (G != GSR_DEV_ONLY_G) return fail(h, GSR_ESTATE, "developer build for one group size only");
  LAUNCH(GSR_DEV_ONLY_G)
#else
  if (G == 32) { LAUNCH(32) } else if (G == 16) { LAUNCH(16) } else { LAUNCH(8) }
#endif
#undef LAUNCH
  CUDA_OK(cudaGetLastError());
  return GSR_OK;
}

// n substeps; `io` (mapped host buffers of gsr_step_host) feeds the first launch and is filled by the last one
static int step_impl(gsr_handle_t h, int n_substeps, void* cuda_stream, HostIO io) {
  if (!h || !h->state) return fail(h, GSR_ESTATE, "state not bound");
  cudaStream_t st = (cudaStream_t)cuda_stream;
  int rc = activate(h, st);
  if (rc) return rc;
  for (int s = 0; s < n_substeps; s++) {
    if (h->group != 1 && h->reorder_period > 0 && ++h->steps_since_reorder >= h->reorder_period) {
      h->steps_since_reorder = 0;
      rc = reorder_envs(h, st);
      if (rc) return rc;
    }
    const bool last = s == n_substeps - 1;
    // the lane kernel recomputes and writes the kinematics itself; diagnostics (geom poses of the new configuration) need them too
    const bool end_fk = h->eager_fk || h->group == 1 || h->hm.diagnostics != 0;
    const int fk_flags = (h->cache_valid ? 0 : 1) | (end_fk ? 2 : 0);
    rc = group_launch(h, K_STEP, st, nullptr, last, HostIO{s == 0 ? io.ctrl_pos : nullptr, last ? io.qpos : nullptr, last ? io.vel : nullptr},
                      StateIO{nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr}, fk_flags);
    if (rc) return rc;
    h->cache_valid = end_fk;
  }
  return GSR_OK;
}

// device-side alias of a page-locked (pinned / registered) host pointer, or null for pageable memory
static void* mapped_alias(const void* p) {
  cudaPointerAttributes a;
  if (cudaPointerGetAttributes(&a, p) != cudaSuccess) { cudaGetLastError(); return nullptr; }
  return a.type == cudaMemoryTypeHost ? a.devicePointer : nullptr;
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
  rc = group_launch(h, K_FORWARD, st, nullptr, 0);
  h->cache_valid = rc == GSR_OK;
  return rc;
}

int gsr_set_kinematics_policy(gsr_handle_t h, int eager) {
  if (!h || !h->committed) return fail(h, GSR_ESTATE, "model not committed");
  h->eager_fk = eager != 0;
  return GSR_OK;
}
int gsr_kinematics_valid(gsr_handle_t h) { return (h && h->cache_valid) ? 1 : 0; }

int gsr_step(gsr_handle_t h, int n_substeps, void* cuda_stream) { return step_impl(h, n_substeps, cuda_stream, HostIO{nullptr, nullptr, nullptr}); }

int gsr_forward_kinematics(gsr_handle_t h, const uint8_t* env_mask_dev, void* cuda_stream) {
  if (!h || !h->state) return fail(h, GSR_ESTATE, "state not bound");
  cudaStream_t st = (cudaStream_t)cuda_stream;
  int rc = activate(h, st);
  if (rc) return rc;
  rc = group_launch(h, K_FORWARD, st, env_mask_dev, 0);
  if (rc == GSR_OK && !env_mask_dev) h->cache_valid = true;  // a masked pass refreshes its environments only
  return rc;
}

int gsr_set_state(gsr_handle_t h, const uint8_t* env_mask_dev, const float* qpos, const float* dofs_vel, const float* dofs_acc,
                  const float* links_pos, const float* links_quat, const float* i_pos_shift, const float* mass_shift,
                  const float* friction_ratio, int partial, void* cuda_stream) {
  if (!h || !h->state) return fail(h, GSR_ESTATE, "state not bound");
  cudaStream_t st = (cudaStream_t)cuda_stream;
  int rc = activate(h, st);
  if (rc) return rc;
  return group_launch(h, K_SET_STATE, st, env_mask_dev, partial ? 1 : 0, HostIO{nullptr, nullptr, nullptr},
                      StateIO{qpos, dofs_vel, dofs_acc, links_pos, links_quat, i_pos_shift, mass_shift, friction_ratio});
}

int gsr_gather_fields(gsr_handle_t h, int n, const char* const* names, void* const* dst_dev, void* cuda_stream) {
  if (!h || !h->state) return fail(h, GSR_ESTATE, "state not bound");
  if (n <= 0 || n > GSR_MAX_GATHER || !names || !dst_dev) return fail(h, GSR_EINVAL, "gsr_gather_fields: 1..16 fields");
  cudaStream_t st = (cudaStream_t)cuda_stream;
  int rc = activate(h, st);
  if (rc) return rc;
  GatherIO io;
  memset(&io, 0, sizeof(io));
  io.n = n;
  for (int k = 0; k < n; k++) {
    const GsrFieldInfo* f = nullptr;
    for (const auto& c : h->fields) if (c.name == names[k]) { f = &c; break; }
    if (!f || f->seg_base != 0 || f->item_stride != f->comps) return fail(h, GSR_EINVAL, std::string("gsr_gather_fields: not a HOT-record field: ") + names[k]);
    io.off[k] = (int)f->offset; io.width[k] = (int)(f->items * f->comps); io.start[k + 1] = io.start[k] + io.width[k];
    io.dst[k] = (float*)dst_dev[k];
    if (!io.dst[k]) return fail(h, GSR_EINVAL, "gsr_gather_fields: null destination");
  }
  const long long total = (long long)h->B * io.start[n];
  { Bracket br_(h, K_GATHER_FIELDS, st); k_gather_fields<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(h->state, h->B, io); }
  CUDA_OK(cudaGetLastError());
  return GSR_OK;
}

int gsr_compute_diagnostics(gsr_handle_t h, void* cuda_stream) {
  if (!h || !h->state) return fail(h, GSR_ESTATE, "state not bound");
  cudaStream_t st = (cudaStream_t)cuda_stream;
  int rc = activate(h, st);
  if (rc) return rc;
  return group_launch(h, K_DIAG, st, nullptr, 0);
}

int gsr_forward_kinematics_dirty(gsr_handle_t h, void* cuda_stream) {
  if (!h || !h->state) return fail(h, GSR_ESTATE, "state not bound");
  cudaStream_t st = (cudaStream_t)cuda_stream;
  int rc = activate(h, st);
  if (rc) return rc;
  return group_launch(h, K_FORWARD, st, nullptr, 1);
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
  if (h->group != 1 && n_substeps > 0 && h->host_io_mode > 0) {
    // Fused host I/O: k_step itself reads the targets from / writes the results to dense (B, n) buffers (HostIO), so no
    // scatter / gather launch is left.  Inputs always come through one DMA into a device staging buffer: reads of mapped
    // host memory from the SMs are bounded by the PCIe read tags (measured: Franka 30000 e2e 0.583 ms against 0.531 ms
    // staged).  Outputs: mode 1 = device staging + one DMA per buffer; mode 2 = posted writes straight into page-locked
    // caller buffers, overlapping the transfer with the compute of the same launch (pageable buffers fall back to mode 1).
    const size_t need1 = B * (size_t)(m.nQ + 2 * m.nD);
    if (h->stage_elems < need1) {
      if (h->d_stage) cudaFree(h->d_stage);
      CUDA_OK(cudaMalloc((void**)&h->d_stage, need1 * 4));
      h->stage_elems = need1;
    }
    float *d_c = h->d_stage, *d_q = d_c + B * m.nD, *d_v = d_q + B * m.nQ;
    HostIO io{nullptr, nullptr, nullptr};
    if (ctrl_pos_host) {
      CUDA_OK(cudaMemcpyAsync(d_c, ctrl_pos_host, B * m.nD * 4, cudaMemcpyHostToDevice, st));
      io.ctrl_pos = d_c;
    }
    float* a_q = h->host_io_mode == 2 && qpos_host ? (float*)mapped_alias(qpos_host) : nullptr;
    float* a_v = h->host_io_mode == 2 && vel_host ? (float*)mapped_alias(vel_host) : nullptr;
    if (qpos_host) io.qpos = a_q ? a_q : d_q;
    if (vel_host) io.vel = a_v ? a_v : d_v;
    rc = step_impl(h, n_substeps, cuda_stream, io);
    if (rc) return rc;
    if (qpos_host && !a_q) CUDA_OK(cudaMemcpyAsync(qpos_host, d_q, B * m.nQ * 4, cudaMemcpyDeviceToHost, st));
    if (vel_host && !a_v) CUDA_OK(cudaMemcpyAsync(vel_host, d_v, B * m.nD * 4, cudaMemcpyDeviceToHost, st));
    CUDA_OK(cudaGetLastError());
    CUDA_OK(cudaStreamSynchronize(st));
    return GSR_OK;
  }
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

#ifdef GSR_NARROW_TRACE
extern "C" int gsr_debug_narrow(unsigned long long* out16) {  // developer build only: read and clear the counters of gsr_gjk.cuh
  if (cudaMemcpyFromSymbol(out16, g_narrow, sizeof(unsigned long long) * 16) != cudaSuccess) return GSR_ECUDA;
  unsigned long long z[16] = {0};
  return cudaMemcpyToSymbol(g_narrow, z, sizeof(z)) == cudaSuccess ? GSR_OK : GSR_ECUDA;
}
#endif
#ifdef GSR_SOLVER_TRACE
int gsr_debug_solver_read(gsr_handle_t h, unsigned long long* out, int n_envs) {  // developer build only (make trace)
  CUDA_OK(cudaDeviceSynchronize());
  CUDA_OK(cudaMemcpyFromSymbol(out, g_solver_dbg, (size_t)n_envs * 16 * sizeof(unsigned long long)));
  return GSR_OK;
}
#endif

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

// Developer diagnostics (tools/step_trace.py): while enabled, the lane-group step kernel records for every environment the
// start / end time (globaltimer, ns) and the SM it ran on; gsr_debug_trace_read copies 8 * n_envs words to the host.
int gsr_debug_trace(gsr_handle_t h, int enabled) {
  if (!h || !h->committed) return fail(h, GSR_ESTATE, "model not committed");
  CUDA_OK(cudaSetDevice(h->device));
  if (enabled && !h->d_trace) { CUDA_OK(cudaMalloc((void**)&h->d_trace, (size_t)h->B * 8 * sizeof(unsigned long long))); CUDA_OK(cudaMemset(h->d_trace, 0, (size_t)h->B * 8 * sizeof(unsigned long long))); }
  if (!enabled && h->d_trace) { cudaFree(h->d_trace); h->d_trace = nullptr; }
  return GSR_OK;
}
int gsr_debug_trace_read(gsr_handle_t h, unsigned long long* host_out) {
  if (!h || !h->d_trace || !host_out) return fail(h, GSR_ESTATE, "tracing is not enabled");
  CUDA_OK(cudaMemcpy(host_out, h->d_trace, (size_t)h->B * 8 * sizeof(unsigned long long), cudaMemcpyDeviceToHost));
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
    cudaFree(h->d_terrain);
    cudaFree(h->d_verts); cudaFree(h->d_support); cudaFree(h->d_support_vid); cudaFree(h->d_pairs); cudaFree(h->d_pair_idx); cudaFree(h->d_model); cudaFree(h->d_fac); if (h->d_trace) cudaFree(h->d_trace);
    if (h->d_order) { cudaFree(h->d_order); cudaFree(h->d_iota); cudaFree(h->d_cost); cudaFree(h->d_cost_sorted); cudaFree(h->d_sort_tmp); h->d_order = nullptr; }
    if (h->d_rows_hist) { cudaFree(h->d_rows_hist); cudaFreeHost(h->h_rows_hist); cudaEventDestroy(h->rows_event); h->d_rows_hist = nullptr; h->h_rows_hist = nullptr; h->rows_event = nullptr; }
    if (h->d_stage) { cudaFree(h->d_stage); h->d_stage = nullptr; }
    for (auto& sp : h->spans) { cudaEventDestroy(sp.a); cudaEventDestroy(sp.b); }
    for (auto e : h->event_pool) cudaEventDestroy(e);
  }
  delete h;
}

}  // extern "C"
