// api.cu -- the C ABI of libaicp_b200.so (include/aicp_b200.h).  Thin: argument checks, host<->device staging on the
// handle's stream, and calls into index.cu / normals.cu / icp.cu / overlap.cu / crop.cu / prefilter.cu / alignability.cu / svm.cu /
// ingest.cu.  No per-point work happens on the host; what does run there is per-call or per-cluster control logic (the 6-digit ratio
// round trip, ~100 3x3 eigen-decompositions and the reference's greedy cluster matching in alignability.cu, file parsing in
// ingest.cu / svm.cu / config_yaml.cpp).
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <exception>
#include <mutex>
#include <atomic>
#include <string>
#include <thread>
#include <vector>

#include "handle.cuh"

namespace aicp {

static std::string g_create_error;
static std::mutex g_mutex;

int fail(Handle* h, int code, const char* fmt, ...) {
  char buf[512];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof(buf), fmt, ap);
  va_end(ap);
  if (h) h->last_error = buf;
  else { std::lock_guard<std::mutex> lk(g_mutex); g_create_error = buf; }
  return code;
}

int fail_cuda(Handle* h, cudaError_t e, const char* expr, int line) {
  cudaGetLastError();   // clear the sticky flag where possible
  return fail(h, AICP_B200_ERR_CUDA, "CUDA error %s (%s) at line %d: %s", cudaGetErrorName(e), cudaGetErrorString(e), line, expr);
}

bool is_device_ptr(const void* p) {
  cudaPointerAttributes a;
  if (cudaPointerGetAttributes(&a, p) != cudaSuccess) { cudaGetLastError(); return false; }
  return a.type == cudaMemoryTypeDevice || a.type == cudaMemoryTypeManaged;
}

// Host-to-device copy of a cloud in pieces of 512 KB (AICP_B200_H2D_CHUNK_KB, 0 = one copy).  While a copy engine has a
// long read of host memory in flight, the GPU's own fetches of launch commands wait behind it on the same PCIe link; with
// eight registrations launching kernels beside the uploads of the next pairs that costs throughput (measured: a saturating
// background upload halves the batched rate; tools/e2e_probe.py).  Pieces leave gaps for the command fetches: +1.4 % on the
// end-to-end rate of a 64-pair batch; 64 KB pieces cost more in commands than they give (profiles/round2_k_e2e_probe.txt).
static int h2d_copy(Handle* h, void* dst, const void* src, size_t bytes) {
  static const size_t chunk = [] { const char* e = getenv("AICP_B200_H2D_CHUNK_KB"); return (size_t)(e ? atoi(e) : 512) * 1024; }();
  if (!chunk || bytes <= chunk) { CUDA_TRY(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, h->stream)); return AICP_B200_OK; }
  for (size_t off = 0; off < bytes; off += chunk)
    CUDA_TRY(cudaMemcpyAsync((char*)dst + off, (const char*)src + off, bytes - off < chunk ? bytes - off : chunk, cudaMemcpyHostToDevice, h->stream));
  return AICP_B200_OK;
}

// make `n` points available on the device: device pointers are used in place, host pointers are staged into `buf`
int upload_points(Handle* h, DevBuf<float4>& buf, const float* xyzw, int64_t n, const float4** out_dev) {
  if (is_device_ptr(xyzw)) { *out_dev = reinterpret_cast<const float4*>(xyzw); return AICP_B200_OK; }
  CUDA_TRY(buf.reserve((size_t)n));
  int rc = h2d_copy(h, buf.p, xyzw, sizeof(float4) * (size_t)n);
  if (rc) return rc;
  *out_dev = buf.p;
  return AICP_B200_OK;
}

// copy `bytes` from a device buffer to a host-or-device destination
static int download(Handle* h, void* dst, const void* src_dev, size_t bytes) {
  CUDA_TRY(cudaMemcpyAsync(dst, src_dev, bytes, is_device_ptr(dst) ? cudaMemcpyDeviceToDevice : cudaMemcpyDeviceToHost, h->stream));
  CUDA_TRY(cudaStreamSynchronize(h->stream));
  return AICP_B200_OK;
}

static int stage_owned(Handle* h, DevBuf<float4>& buf, const float* xyzw, int64_t n) {
  // registration keeps its own copy of both clouds (the reference copies into its DP members too,
  // pointmatcher_registration.cpp:16-20), so device inputs are copied device-to-device
  CUDA_TRY(buf.reserve((size_t)n));
  if (!is_device_ptr(xyzw)) return h2d_copy(h, buf.p, xyzw, sizeof(float4) * (size_t)n);
  CUDA_TRY(cudaMemcpyAsync(buf.p, xyzw, sizeof(float4) * (size_t)n, cudaMemcpyDeviceToDevice, h->stream));
  return AICP_B200_OK;
}

static int load_config(Handle* h) {
  if (!h->cfg_from_file) return AICP_B200_OK;
  std::string err;
  aicp_b200_icp_config cfg;
  int rc = parse_icp_yaml(h->cfg_path.c_str(), &cfg, &err);
  if (rc) return fail(h, rc, "%s", err.c_str());
  h->cfg = cfg;
  return AICP_B200_OK;
}

static int bind_device(Handle* h) {
  CUDA_TRY(cudaSetDevice(h->device));
  return AICP_B200_OK;
}

}  // namespace aicp

using namespace aicp;

#define H_CHECK(h) do { if (!(h)) return AICP_B200_ERR_BAD_ARG; int rc__ = bind_device(h); if (rc__) return rc__; } while (0)

extern "C" {

const char* aicp_b200_version(void) { return "aicp_b200 0.1 (sm_100a)"; }

int aicp_b200_create(const char* icp_yaml_path, int device, aicp_b200_handle** out) {
  if (!out) return AICP_B200_ERR_BAD_ARG;
  *out = nullptr;
  Handle* h = nullptr;
  int count = 0;
  cudaError_t e = cudaGetDeviceCount(&count);
  if (e != cudaSuccess || count == 0) {
    cudaGetLastError();
    return fail(nullptr, AICP_B200_ERR_CUDA, "no CUDA device available (%s); libaicp_b200 has no CPU fallback",
                e != cudaSuccess ? cudaGetErrorString(e) : "device count 0");
  }
  if (device < 0) { if (cudaGetDevice(&device) != cudaSuccess) device = 0; }
  if (device >= count) return fail(nullptr, AICP_B200_ERR_BAD_ARG, "device %d out of range (%d devices)", device, count);
  Handle* nh = new Handle();
  nh->device = device;
  default_icp_config(&nh->cfg);
  if (icp_yaml_path && *icp_yaml_path) { nh->cfg_path = icp_yaml_path; nh->cfg_from_file = true; }
  h = nh;
  if (cudaSetDevice(device) != cudaSuccess || cudaStreamCreateWithFlags(&nh->stream, cudaStreamNonBlocking) != cudaSuccess) {
    delete nh;
    return fail(nullptr, AICP_B200_ERR_CUDA, "cannot initialise CUDA device %d", device);
  }
  for (int i = 0; i < 4; ++i) cudaEventCreate(&nh->ev[i]);
  if (const char* e = getenv("AICP_B200_CROP")) nh->crop_legacy = strcmp(e, "legacy") == 0;
  if (const char* e = getenv("AICP_B200_FUSED_TAIL")) nh->fused_tail = atoi(e) != 0;
  if (const char* e = getenv("AICP_B200_SPREAD")) { int v = atoi(e); if (v == 1 || v == 2 || v == 4 || v == 8) nh->loop_spread = v; }
  (void)h;
  *out = reinterpret_cast<aicp_b200_handle*>(nh);
  return AICP_B200_OK;
}

int aicp_b200_comm_destroy(aicp_b200_handle* hh);

int aicp_b200_destroy(aicp_b200_handle* hh) {
  Handle* h = reinterpret_cast<Handle*>(hh);
  if (!h) return AICP_B200_OK;
  cudaSetDevice(h->device);
  if (h->comm) aicp_b200_comm_destroy(hh);
  if (h->al_child) { aicp_b200_destroy(reinterpret_cast<aicp_b200_handle*>(h->al_child)); h->al_child = nullptr; }
  for (Handle* w : h->workers) aicp_b200_destroy(reinterpret_cast<aicp_b200_handle*>(w));
  h->workers.clear();
  for (Handle* c : h->device_children) aicp_b200_destroy(reinterpret_cast<aicp_b200_handle*>(c));
  h->device_children.clear();
  cudaSetDevice(h->device);
  if (h->done_ev) cudaEventDestroy(h->done_ev);
  if (h->setup_exec) cudaGraphExecDestroy(h->setup_exec);
  for (int i = 0; i < 2; ++i) if (h->batch_ev[i]) cudaEventDestroy(h->batch_ev[i]);
  if (h->stream) cudaStreamSynchronize(h->stream);
  h->ref_in.release(); h->ref_ix.release(); h->refc_pts.release(); h->refc_rec.release(); h->refc_cell.release(); h->normals.release(); h->knn_pos.release();
  h->ref_rk2.release(); h->app_pts.release(); h->app_normals.release(); h->app_keys.release(); h->app_vals.release(); h->app_new.release(); h->app_flag.release();
  h->app_scan.release(); h->app_tiles.release(); h->app_rk2.release(); h->app_rmax.release(); h->app_list.release();
  if (h->app_meta) cudaFree(h->app_meta);
  h->read_in.release(); h->read_ix.release(); h->read0.release(); h->read_out.release(); h->read_init.release();
  h->match_pos.release(); h->d2.release(); h->hist.release(); h->cand.release(); h->acc_slots.release(); h->trace_idx.release();
  if (h->progress_host) cudaFreeHost((void*)h->progress_host);
  if (h->crop_total_host) cudaFreeHost((void*)h->crop_total_host);
  h->crop_stash.release();
  h->tmp_ix.release(); h->tmp_a.release(); h->tmp_b.release(); h->tmp_i.release(); h->tmp_f.release();
  h->ovl_bits_a.release(); h->ovl_bits_b.release(); h->ovl_counts.release(); h->crop_status.release(); h->crop_out.release(); h->map.release();
  h->pf_ix.release(); h->pf_sampled.release(); h->pf_normals.release(); h->pf_normals_orig.release(); h->pf_out.release();
  h->pf_keys.release(); h->pf_keys_alt.release(); h->pf_vals.release(); h->pf_vals_alt.release(); h->pf_sort_tmp.release();
  h->pf_flag.release(); h->pf_slot.release(); h->pf_tiles.release(); h->pf_mask.release(); h->pf_count.release();
  h->pf_label.release(); h->pf_seed_pos.release(); h->pf_labels_out.release();
  h->pf_status.release(); h->pf_mutual.release(); h->pf_parent.release(); h->pf_root.release(); h->pf_clabel.release();
  svm_release(h);
  h->acc.release(); h->acc_tmp.release();
  h->al_moved.release(); h->al_mean.release(); h->al_ext.release(); h->al_sums.release(); h->al_cnt.release(); h->al_counts.release(); h->al_axes.release();
  for (int i = 0; i < 2; ++i) { h->al_fov[i].release(); h->al_pts[i].release(); h->al_lab[i].release(); h->al_boxes[i].release(); }
  if (h->pf_meta) cudaFree(h->pf_meta);
  if (h->pf_meta_host) cudaFreeHost(h->pf_meta_host);
  for (int i = 0; i < 2; ++i) if (h->pf_ev[i]) cudaEventDestroy(h->pf_ev[i]);
  if (h->st) cudaFree(h->st);
  if (h->st_host) cudaFreeHost(h->st_host);
  for (int i = 0; i < 4; ++i) if (h->ev[i]) cudaEventDestroy(h->ev[i]);
  if (h->ev_fork) cudaEventDestroy(h->ev_fork);
  if (h->ev_init) cudaEventDestroy(h->ev_init);
  if (h->ev_join) cudaEventDestroy(h->ev_join);
  if (h->side) cudaStreamDestroy(h->side);
  for (cudaEvent_t e : h->prof_ev) cudaEventDestroy(e);
  if (h->stream) cudaStreamDestroy(h->stream);
  delete h;
  return AICP_B200_OK;
}

const char* aicp_b200_last_error(const aicp_b200_handle* hh) {
  const Handle* h = reinterpret_cast<const Handle*>(hh);
  if (h) return h->last_error.c_str();
  std::lock_guard<std::mutex> lk(g_mutex);
  return g_create_error.c_str();
}

int aicp_b200_set_config(aicp_b200_handle* hh, const char* icp_yaml_path) {
  Handle* h = reinterpret_cast<Handle*>(hh);
  if (!h) return AICP_B200_ERR_BAD_ARG;
  if (icp_yaml_path && *icp_yaml_path) { h->cfg_path = icp_yaml_path; h->cfg_from_file = true; }
  else { h->cfg_path.clear(); h->cfg_from_file = false; default_icp_config(&h->cfg); }
  return AICP_B200_OK;
}

int aicp_b200_set_config_struct(aicp_b200_handle* hh, const aicp_b200_icp_config* cfg) {
  Handle* h = reinterpret_cast<Handle*>(hh);
  if (!h || !cfg) return AICP_B200_ERR_BAD_ARG;
  h->cfg = *cfg;
  h->cfg_path.clear();
  h->cfg_from_file = false;
  return AICP_B200_OK;
}

int aicp_b200_get_config(aicp_b200_handle* hh, aicp_b200_icp_config* cfg) {
  Handle* h = reinterpret_cast<Handle*>(hh);
  if (!h || !cfg) return AICP_B200_ERR_BAD_ARG;
  int rc = load_config(h);
  if (rc) return rc;
  *cfg = h->cfg;
  return AICP_B200_OK;
}

int aicp_b200_parse_icp_yaml(const char* icp_yaml_path, aicp_b200_icp_config* cfg, char* err, int err_len) {
  if (!cfg) return AICP_B200_ERR_BAD_ARG;
  std::string e;
  int rc = parse_icp_yaml(icp_yaml_path, cfg, &e);
  if (err && err_len > 0) { strncpy(err, e.c_str(), (size_t)err_len - 1); err[err_len - 1] = 0; }
  return rc;
}

int aicp_b200_register(aicp_b200_handle* hh, const float* ref_xyzw, int64_t n_ref, const float* read_xyzw, int64_t n_read,
                       const float* init_T, float* out_T, aicp_b200_stats* stats) {
  Handle* h = reinterpret_cast<Handle*>(hh);
  H_CHECK(h);
  if (!ref_xyzw || !read_xyzw || !out_T || n_ref < 1 || n_read < 1 || n_ref > (1ll << 30) || n_read > (1ll << 30))
    return fail(h, AICP_B200_ERR_BAD_ARG, "registerClouds: null cloud or empty cloud (n_ref %lld, n_read %lld)", (long long)n_ref, (long long)n_read);
  int rc = load_config(h);     // the reference re-reads the YAML on every call (pointmatcher_registration.cpp:103)
  if (rc) return rc;
  if ((rc = stage_owned(h, h->ref_in, ref_xyzw, n_ref))) return rc;
  if ((rc = stage_owned(h, h->read_in, read_xyzw, n_read))) return rc;
  h->n_ref = n_ref; h->n_read = n_read;
  float init_host[16];
  const float* init = nullptr;
  if (init_T) {
    if (is_device_ptr(init_T)) { CUDA_TRY(cudaMemcpy(init_host, init_T, sizeof(init_host), cudaMemcpyDeviceToHost)); init = init_host; }
    else init = init_T;
  }
  return run_registration(h, init, true, stats, out_T);
}

int aicp_b200_set_reference(aicp_b200_handle* hh, const float* ref_xyzw, int64_t n_ref) {
  Handle* h = reinterpret_cast<Handle*>(hh);
  H_CHECK(h);
  if (!ref_xyzw || n_ref < 1 || n_ref > (1ll << 30)) return fail(h, AICP_B200_ERR_BAD_ARG, "set_reference: null or empty cloud");
  int rc = load_config(h);
  if (rc) return rc;
  if ((rc = stage_owned(h, h->ref_in, ref_xyzw, n_ref))) return rc;
  h->n_ref = n_ref;
  h->ref_ready = false;
  return AICP_B200_OK;
}

int aicp_b200_register_to_reference(aicp_b200_handle* hh, const float* read_xyzw, int64_t n_read, const float* init_T,
                                    float* out_T, aicp_b200_stats* stats) {
  Handle* h = reinterpret_cast<Handle*>(hh);
  H_CHECK(h);
  if (h->n_ref < 1) return fail(h, AICP_B200_ERR_BAD_ARG, "register_to_reference: call aicp_b200_set_reference first");
  if (!read_xyzw || !out_T || n_read < 1 || n_read > (1ll << 30)) return fail(h, AICP_B200_ERR_BAD_ARG, "register_to_reference: null or empty reading");
  int rc = load_config(h);
  if (rc) return rc;
  if ((rc = stage_owned(h, h->read_in, read_xyzw, n_read))) return rc;
  h->n_read = n_read;
  bool rebuild = !h->ref_ready || h->ref_knn != h->cfg.knn_normals;
  float init_host[16];
  const float* init = nullptr;
  if (init_T) {
    if (is_device_ptr(init_T)) { CUDA_TRY(cudaMemcpy(init_host, init_T, sizeof(init_host), cudaMemcpyDeviceToHost)); init = init_host; }
    else init = init_T;
  }
  return run_registration(h, init, rebuild, stats, out_T);
}

int aicp_b200_reference_append(aicp_b200_handle* hh, const float* xyzw, int64_t n, aicp_b200_append_info* info) {
  Handle* h = reinterpret_cast<Handle*>(hh);
  H_CHECK(h);
  if (n < 0 || (n > 0 && !xyzw) || h->n_ref + n > (1ll << 28)) return fail(h, AICP_B200_ERR_BAD_ARG, "reference_append: bad arguments");
  if (h->comm) return fail(h, AICP_B200_ERR_BAD_ARG, "reference_append: not available on a handle with a communicator");
  if (h->n_ref < 1) return fail(h, AICP_B200_ERR_BAD_ARG, "reference_append: call aicp_b200_set_reference first");
  if (info) { memset(info, 0, sizeof(*info)); info->n_total = h->n_ref; }
  if (n == 0) return AICP_B200_OK;
  int rc = load_config(h);
  if (rc) return rc;
  if (!h->ref_ready || h->ref_knn != h->cfg.knn_normals) {
    // no live index to update yet (or the chain asks for other neighbourhoods): grow the stored cloud, the next registration builds
    const int64_t total = h->n_ref + n;
    if ((size_t)total > h->ref_in.cap) {
      DevBuf<float4> bigger;
      CUDA_TRY(bigger.reserve((size_t)total + (size_t)total / 4));
      CUDA_TRY(cudaMemcpyAsync(bigger.p, h->ref_in.p, sizeof(float4) * (size_t)h->n_ref, cudaMemcpyDeviceToDevice, h->stream));
      CUDA_TRY(cudaStreamSynchronize(h->stream));
      h->ref_in.release();
      h->ref_in = bigger;
    }
    CUDA_TRY(cudaMemcpyAsync(h->ref_in.p + h->n_ref, xyzw, sizeof(float4) * (size_t)n,
                             is_device_ptr(xyzw) ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice, h->stream));
    CUDA_TRY(cudaStreamSynchronize(h->stream));
    h->n_ref = total; h->ref_ready = false;
    if (info) info->n_total = total;
    return AICP_B200_OK;
  }
  const float4* pts;
  if ((rc = upload_points(h, h->tmp_a, xyzw, n, &pts))) return rc;
  return run_reference_append(h, pts, n, info);
}

int aicp_b200_get_output_reading(aicp_b200_handle* hh, float* xyzw, int64_t n) {
  Handle* h = reinterpret_cast<Handle*>(hh);
  H_CHECK(h);
  if (!xyzw || n != h->n_read || !h->read_out.p) return fail(h, AICP_B200_ERR_BAD_ARG, "get_output_reading: no registration of %lld points has run", (long long)n);
  return download(h, xyzw, h->read_out.p, sizeof(float4) * (size_t)n);
}

int aicp_b200_get_initialized_reading(aicp_b200_handle* hh, float* xyzw, int64_t n) {
  Handle* h = reinterpret_cast<Handle*>(hh);
  H_CHECK(h);
  if (!xyzw || n != h->n_read || !h->read_in.p) return fail(h, AICP_B200_ERR_BAD_ARG, "get_initialized_reading: no registration of %lld points has run", (long long)n);
  // pointmatcher_registration.hpp:37-46: falls back to the raw reading when no initial transform was given
  return download(h, xyzw, h->has_init_reading ? h->read_init.p : h->read_in.p, sizeof(float4) * (size_t)n);
}

int aicp_b200_get_reference_normals(aicp_b200_handle* hh, float* normals_xyzd, int64_t n) {
  Handle* h = reinterpret_cast<Handle*>(hh);
  H_CHECK(h);
  if (!normals_xyzd || !h->ref_ready || n != h->n_ref) return fail(h, AICP_B200_ERR_BAD_ARG, "get_reference_normals: no reference of %lld points", (long long)n);
  CUDA_TRY(h->tmp_b.reserve((size_t)n));
  int rc = scatter_normals(h, h->ref_ix.pts.p, h->normals.p, h->ref_ix.n, h->tmp_b.p);
  if (rc) return rc;
  return download(h, normals_xyzd, h->tmp_b.p, sizeof(float4) * (size_t)n);
}

int aicp_b200_wait_stream(aicp_b200_handle* hh, void* producer_stream) {
  Handle* h = reinterpret_cast<Handle*>(hh);
  H_CHECK(h);
  // The library reads device-pointer inputs on its own non-blocking stream(s); this orders everything it enqueues from now
  // on after the work already enqueued on the caller's stream (the producer of those inputs).  Batch workers and child
  // handles start behind an event of h->stream, so they inherit the dependency.
  if (!h->wait_ev) CUDA_TRY(cudaEventCreateWithFlags(&h->wait_ev, cudaEventDisableTiming));
  CUDA_TRY(cudaEventRecord(h->wait_ev, reinterpret_cast<cudaStream_t>(producer_stream)));
  CUDA_TRY(cudaStreamWaitEvent(h->stream, h->wait_ev, 0));
  return AICP_B200_OK;
}

int aicp_b200_enable_match_trace(aicp_b200_handle* hh, int enable) {
  Handle* h = reinterpret_cast<Handle*>(hh);
  if (!h) return AICP_B200_ERR_BAD_ARG;
  h->trace_matches = enable != 0;
  return AICP_B200_OK;
}

int aicp_b200_set_profiling(aicp_b200_handle* hh, int enable) {
  Handle* h = reinterpret_cast<Handle*>(hh);
  if (!h) return AICP_B200_ERR_BAD_ARG;
  h->profiling = enable < 0 ? 0 : (enable > 2 ? 2 : enable);
  return AICP_B200_OK;
}

int aicp_b200_set_knn_schedule(aicp_b200_handle* hh, int schedule) {
  Handle* h = reinterpret_cast<Handle*>(hh);
  if (!h || schedule < 0 || schedule > 2) return AICP_B200_ERR_BAD_ARG;
  h->knn_schedule = schedule;
  return AICP_B200_OK;
}

int aicp_b200_set_match_schedule(aicp_b200_handle* hh, int schedule) {
  Handle* h = reinterpret_cast<Handle*>(hh);
  if (!h || schedule < 0 || schedule > 2) return AICP_B200_ERR_BAD_ARG;
  h->match_schedule = schedule;
  return AICP_B200_OK;
}

int aicp_b200_set_loop_schedule(aicp_b200_handle* hh, int schedule) {
  Handle* h = reinterpret_cast<Handle*>(hh);
  if (!h || schedule < 0 || schedule > 2) return AICP_B200_ERR_BAD_ARG;
  h->loop_schedule = schedule;
  return AICP_B200_OK;
}

int aicp_b200_get_trace_matches(aicp_b200_handle* hh, int32_t* idx, int64_t iters, int64_t n_read) {
  Handle* h = reinterpret_cast<Handle*>(hh);
  H_CHECK(h);
  if (!idx || !h->trace_idx.p || n_read != h->trace_n || iters > h->trace_iters || iters < 0)
    return fail(h, AICP_B200_ERR_BAD_ARG, "get_trace_matches: trace holds %lld x %lld", (long long)h->trace_iters, (long long)h->trace_n);
  if (iters == 0) return AICP_B200_OK;
  return download(h, idx, h->trace_idx.p, sizeof(int32_t) * (size_t)(iters * n_read));
}

int aicp_b200_surface_normals(aicp_b200_handle* hh, const float* xyzw, int64_t n, int32_t knn, float* out_normals, int32_t* out_knn) {
  Handle* h = reinterpret_cast<Handle*>(hh);
  H_CHECK(h);
  if (!xyzw || !out_normals || n < 1) return fail(h, AICP_B200_ERR_BAD_ARG, "surface_normals: null or empty cloud");
  const float4* pts;
  int rc = upload_points(h, h->tmp_a, xyzw, n, &pts);
  if (rc) return rc;
  if ((rc = build_index(h, h->tmp_ix, pts, n))) return rc;
  CUDA_TRY(h->tmp_b.reserve((size_t)h->tmp_ix.n * 2));
  int* knn_dev = nullptr;
  if (out_knn) { CUDA_TRY(h->tmp_i.reserve((size_t)n * knn)); knn_dev = h->tmp_i.p; }
  float4* nm = h->tmp_b.p;              // Morton order
  float4* no = h->tmp_b.p + h->tmp_ix.n;   // original order
  if ((rc = run_surface_normals(h, h->tmp_ix, knn, nm, knn_dev))) return rc;
  if ((rc = scatter_normals(h, h->tmp_ix.pts.p, nm, h->tmp_ix.n, no))) return rc;
  CUDA_TRY(cudaStreamSynchronize(h->stream));
  if ((rc = download(h, out_normals, no, sizeof(float4) * (size_t)n))) return rc;
  if (out_knn) return download(h, out_knn, knn_dev, sizeof(int32_t) * (size_t)n * knn);
  return AICP_B200_OK;
}

int aicp_b200_match(aicp_b200_handle* hh, const float* ref_xyzw, int64_t n_ref, const float* qry_xyzw, int64_t n_qry,
                    int32_t* out_idx, float* out_d2) {
  Handle* h = reinterpret_cast<Handle*>(hh);
  H_CHECK(h);
  if (!ref_xyzw || !qry_xyzw || !out_idx || !out_d2 || n_ref < 1 || n_qry < 0) return fail(h, AICP_B200_ERR_BAD_ARG, "match: bad arguments");
  if (n_qry == 0) return AICP_B200_OK;
  const float4 *ref, *qry;
  int rc = upload_points(h, h->tmp_a, ref_xyzw, n_ref, &ref);
  if (rc) return rc;
  if ((rc = upload_points(h, h->tmp_b, qry_xyzw, n_qry, &qry))) return rc;
  if ((rc = build_index(h, h->tmp_ix, ref, n_ref))) return rc;
  CUDA_TRY(h->tmp_i.reserve((size_t)n_qry));
  CUDA_TRY(h->tmp_f.reserve((size_t)n_qry));
  if ((rc = run_match_stage(h, h->tmp_ix, qry, n_qry, h->tmp_i.p, h->tmp_f.p))) return rc;
  if ((rc = download(h, out_idx, h->tmp_i.p, sizeof(int32_t) * (size_t)n_qry))) return rc;
  return download(h, out_d2, h->tmp_f.p, sizeof(float) * (size_t)n_qry);
}

int aicp_b200_trim_threshold(aicp_b200_handle* hh, const float* d2, int64_t n, float ratio, float* out_limit, int64_t* out_n_valid) {
  Handle* h = reinterpret_cast<Handle*>(hh);
  H_CHECK(h);
  if (!d2 || !out_limit || n < 0 || n > (1ll << 30)) return fail(h, AICP_B200_ERR_BAD_ARG, "trim_threshold: bad arguments");
  const float* dev = d2;
  if (!is_device_ptr(d2)) {
    CUDA_TRY(h->tmp_f.reserve((size_t)(n > 0 ? n : 1)));
    if (n > 0) CUDA_TRY(cudaMemcpyAsync(h->tmp_f.p, d2, sizeof(float) * (size_t)n, cudaMemcpyHostToDevice, h->stream));
    dev = h->tmp_f.p;
  }
  return run_trim_stage(h, dev, n, ratio, out_limit, out_n_valid);
}

int aicp_b200_overlap(aicp_b200_handle* hh, const float* ref_xyzw, int64_t n_ref, const double ref_origin[3],
                      const float* read_xyzw, int64_t n_read, const double read_origin[3], double resolution,
                      float* overlap_pct, int64_t counts[3]) {
  Handle* h = reinterpret_cast<Handle*>(hh);
  H_CHECK(h);
  if ((n_ref > 0 && !ref_xyzw) || (n_read > 0 && !read_xyzw) || !ref_origin || !read_origin || !overlap_pct || n_ref < 0 || n_read < 0 ||
      n_ref > (1ll << 30) || n_read > (1ll << 30))
    return fail(h, AICP_B200_ERR_BAD_ARG, "overlap: bad arguments");
  const float4 *ref = nullptr, *read = nullptr;
  int rc;
  if (n_ref > 0 && (rc = upload_points(h, h->tmp_a, ref_xyzw, n_ref, &ref))) return rc;
  if (n_read > 0 && (rc = upload_points(h, h->tmp_b, read_xyzw, n_read, &read))) return rc;
  return run_overlap(h, ref, n_ref, ref_origin, read, n_read, read_origin, resolution, overlap_pct, counts);
}

int aicp_b200_crop_box(aicp_b200_handle* hh, const float* xyzw, int64_t n, float box_min, float box_max, const float rotation_rpy[3],
                       const float translation[3], float* out_xyzw, int64_t* n_out) {
  Handle* h = reinterpret_cast<Handle*>(hh);
  H_CHECK(h);
  if (!xyzw || !rotation_rpy || !translation || !n_out || n < 0 || n > (1ll << 31))
    return fail(h, AICP_B200_ERR_BAD_ARG, "crop_box: bad arguments");
  *n_out = 0;
  if (n == 0) return AICP_B200_OK;
  const float4* pts;
  int rc = upload_points(h, h->tmp_a, xyzw, n, &pts);
  if (rc) return rc;
  // the result stays on the device (h->crop_out) when out_xyzw is NULL: aicp_b200_get_cropped() returns its address
  const bool out_is_dev = out_xyzw && is_device_ptr(out_xyzw);
  float4* dst = out_is_dev ? reinterpret_cast<float4*>(out_xyzw) : nullptr;
  if (!dst) { CUDA_TRY(h->crop_out.reserve((size_t)n)); dst = h->crop_out.p; }
  if ((rc = run_crop_box(h, pts, n, box_min, box_max, rotation_rpy, translation, dst, n_out))) return rc;
  h->crop_n = *n_out;
  if (out_xyzw && !out_is_dev && *n_out > 0) return download(h, out_xyzw, dst, sizeof(float4) * (size_t)*n_out);
  return AICP_B200_OK;
}

int aicp_b200_map_append(aicp_b200_handle* hh, const float* xyzw, int64_t n, int replace) {
  Handle* h = reinterpret_cast<Handle*>(hh);
  H_CHECK(h);
  if (n < 0 || (n > 0 && !xyzw) || n > (1ll << 31)) return fail(h, AICP_B200_ERR_BAD_ARG, "map_append: bad arguments");
  if (replace) h->map_n = 0;
  if (n == 0) return AICP_B200_OK;
  const int64_t total = h->map_n + n;
  if (total > (1ll << 31)) return fail(h, AICP_B200_ERR_BAD_ARG, "map_append: map would exceed 2^31 points");
  if ((size_t)total > h->map.cap) {
    // grow geometrically and carry the existing points over (DevBuf::reserve alone would drop them)
    DevBuf<float4> bigger;
    CUDA_TRY(bigger.reserve((size_t)total + (size_t)total / 2));
    if (h->map_n > 0) CUDA_TRY(cudaMemcpyAsync(bigger.p, h->map.p, sizeof(float4) * (size_t)h->map_n, cudaMemcpyDeviceToDevice, h->stream));
    CUDA_TRY(cudaStreamSynchronize(h->stream));
    h->map.release();
    h->map = bigger;
  }
  CUDA_TRY(cudaMemcpyAsync(h->map.p + h->map_n, xyzw, sizeof(float4) * (size_t)n,
                           is_device_ptr(xyzw) ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice, h->stream));
  CUDA_TRY(cudaStreamSynchronize(h->stream));
  h->map_n = total;
  return AICP_B200_OK;
}

int64_t aicp_b200_map_size(const aicp_b200_handle* hh) {
  const Handle* h = reinterpret_cast<const Handle*>(hh);
  return h ? h->map_n : 0;
}

int aicp_b200_map_crop(aicp_b200_handle* hh, float box_min, float box_max, const float rotation_rpy[3], const float translation[3],
                       int64_t* n_out) {
  Handle* h = reinterpret_cast<Handle*>(hh);
  H_CHECK(h);
  if (!rotation_rpy || !translation || !n_out) return fail(h, AICP_B200_ERR_BAD_ARG, "map_crop: bad arguments");
  *n_out = 0; h->crop_n = 0;
  if (h->map_n == 0) return AICP_B200_OK;
  CUDA_TRY(h->crop_out.reserve((size_t)h->map_n));
  int rc = run_crop_box(h, h->map.p, h->map_n, box_min, box_max, rotation_rpy, translation, h->crop_out.p, n_out);
  if (rc) return rc;
  h->crop_n = *n_out;
  return AICP_B200_OK;
}

int aicp_b200_download_cropped(aicp_b200_handle* hh, float* xyzw, int64_t n) {
  Handle* h = reinterpret_cast<Handle*>(hh);
  H_CHECK(h);
  if (n != h->crop_n || (n > 0 && !xyzw)) return fail(h, AICP_B200_ERR_BAD_ARG, "download_cropped: the last crop holds %lld points", (long long)h->crop_n);
  if (n == 0) return AICP_B200_OK;
  return download(h, xyzw, h->crop_out.p, sizeof(float4) * (size_t)n);
}

const float* aicp_b200_get_cropped(aicp_b200_handle* hh, int64_t* n_out) {
  Handle* h = reinterpret_cast<Handle*>(hh);
  if (!h) return nullptr;
  if (n_out) *n_out = h->crop_n;
  return reinterpret_cast<const float*>(h->crop_out.p);
}

int aicp_b200_prefilter_default_config(aicp_b200_prefilter_config* cfg) {
  if (!cfg) return AICP_B200_ERR_BAD_ARG;
  cfg->leaf_size = 0.08f;
  cfg->knn_normals = 30;
  cfg->n_neighbours = 15;
  cfg->min_cluster_size = 50;
  cfg->max_cluster_size = 1000000;
  cfg->smoothness_threshold = (float)(3.0 / 180.0 * M_PI);
  cfg->curvature_threshold = 1.0f;
  return AICP_B200_OK;
}

int aicp_b200_prefilter(aicp_b200_handle* hh, const float* xyzw, int64_t n, const aicp_b200_prefilter_config* cfg,
                        const float viewpoint[3], float* out_xyzw, int64_t* n_out, aicp_b200_prefilter_info* info) {
  Handle* h = reinterpret_cast<Handle*>(hh);
  H_CHECK(h);
  if (!n_out || n < 0 || (n > 0 && !xyzw)) return fail(h, AICP_B200_ERR_BAD_ARG, "prefilter: bad arguments");
  *n_out = 0;
  if (info) memset(info, 0, sizeof(*info));
  h->pf_n_out = 0; h->pf_n_sampled = 0; h->pf_n_clusters = 0; h->pf_has_segments = false;
  if (n == 0) return AICP_B200_OK;
  aicp_b200_prefilter_config def;
  if (!cfg) { aicp_b200_prefilter_default_config(&def); cfg = &def; }
  const float4* pts;
  int rc = upload_points(h, h->tmp_a, xyzw, n, &pts);
  if (rc) return rc;
  if ((rc = run_prefilter(h, pts, n, cfg, viewpoint, info))) return rc;
  *n_out = h->pf_n_out;
  if (out_xyzw && h->pf_n_out > 0) return download(h, out_xyzw, h->pf_out.p, sizeof(float4) * (size_t)h->pf_n_out);
  return AICP_B200_OK;
}

const float* aicp_b200_get_prefiltered(aicp_b200_handle* hh, int64_t* n_out) {
  Handle* h = reinterpret_cast<Handle*>(hh);
  if (!h) return nullptr;
  if (n_out) *n_out = h->pf_n_out;
  return reinterpret_cast<const float*>(h->pf_out.p);
}

int aicp_b200_prefilter_get_sampled(aicp_b200_handle* hh, float* xyzw, int64_t n) {
  Handle* h = reinterpret_cast<Handle*>(hh);
  H_CHECK(h);
  if (n != h->pf_n_sampled || (n > 0 && !xyzw)) return fail(h, AICP_B200_ERR_BAD_ARG, "prefilter_get_sampled: the last call sampled %lld points", (long long)h->pf_n_sampled);
  if (n == 0) return AICP_B200_OK;
  return download(h, xyzw, h->pf_sampled.p, sizeof(float4) * (size_t)n);
}

int aicp_b200_prefilter_get_normals(aicp_b200_handle* hh, float* normals_xyzc, int64_t n) {
  Handle* h = reinterpret_cast<Handle*>(hh);
  H_CHECK(h);
  if (n != h->pf_n_sampled || (n > 0 && !normals_xyzc)) return fail(h, AICP_B200_ERR_BAD_ARG, "prefilter_get_normals: the last call sampled %lld points", (long long)h->pf_n_sampled);
  if (n == 0) return AICP_B200_OK;
  if (!h->pf_has_segments) return fail(h, AICP_B200_ERR_BAD_ARG, "prefilter_get_normals: the last call did not reach the normals stage");
  return download(h, normals_xyzc, h->pf_normals_orig.p, sizeof(float4) * (size_t)n);
}

int aicp_b200_prefilter_get_labels(aicp_b200_handle* hh, int32_t* labels, int64_t n) {
  Handle* h = reinterpret_cast<Handle*>(hh);
  H_CHECK(h);
  if (n != h->pf_n_sampled || (n > 0 && !labels)) return fail(h, AICP_B200_ERR_BAD_ARG, "prefilter_get_labels: the last call sampled %lld points", (long long)h->pf_n_sampled);
  if (n == 0) return AICP_B200_OK;
  if (!h->pf_has_segments) {          // fewer sampled points than neighbours: no region can be kept
    if (is_device_ptr(labels)) { CUDA_TRY(cudaMemsetAsync(labels, 0xFF, sizeof(int32_t) * (size_t)n, h->stream)); CUDA_TRY(cudaStreamSynchronize(h->stream)); }
    else for (int64_t i = 0; i < n; ++i) labels[i] = -1;
    return AICP_B200_OK;
  }
  return download(h, labels, h->pf_labels_out.p, sizeof(int32_t) * (size_t)n);
}

int aicp_b200_voxel_grid(aicp_b200_handle* hh, const float* xyzw, int64_t n, float leaf_size, float* out_xyzw, int64_t* n_out) {
  Handle* h = reinterpret_cast<Handle*>(hh);
  H_CHECK(h);
  if (!n_out || n < 0 || (n > 0 && (!xyzw || !out_xyzw))) return fail(h, AICP_B200_ERR_BAD_ARG, "voxel_grid: bad arguments");
  *n_out = 0;
  if (n == 0) return AICP_B200_OK;
  const float4* pts;
  int rc = upload_points(h, h->tmp_a, xyzw, n, &pts);
  if (rc) return rc;
  if ((rc = run_voxel_grid(h, pts, n, leaf_size, n_out))) return rc;
  if (*n_out > 0) return download(h, out_xyzw, h->pf_sampled.p, sizeof(float4) * (size_t)*n_out);
  return AICP_B200_OK;
}

int aicp_b200_map_prefilter(aicp_b200_handle* hh, const aicp_b200_prefilter_config* cfg, int64_t* n_out, aicp_b200_prefilter_info* info) {
  Handle* h = reinterpret_cast<Handle*>(hh);
  H_CHECK(h);
  if (n_out) *n_out = 0;
  if (info) memset(info, 0, sizeof(*info));
  if (h->map_n == 0) return AICP_B200_OK;
  aicp_b200_prefilter_config def;
  if (!cfg) { aicp_b200_prefilter_default_config(&def); cfg = &def; }
  int rc = run_prefilter(h, h->map.p, h->map_n, cfg, nullptr, info);
  if (rc) return rc;
  // prior_map_->updateCloud(map_prefiltered) (app.cpp:491-492): the output never exceeds the input, so it fits in place
  if (h->pf_n_out > 0) CUDA_TRY(cudaMemcpyAsync(h->map.p, h->pf_out.p, sizeof(float4) * (size_t)h->pf_n_out, cudaMemcpyDeviceToDevice, h->stream));
  CUDA_TRY(cudaStreamSynchronize(h->stream));
  h->map_n = h->pf_n_out;
  if (n_out) *n_out = h->pf_n_out;
  return AICP_B200_OK;
}

int aicp_b200_accumulate_sweep(aicp_b200_handle* hh, const float* sweep_xyzw, int64_t n, float box_half, const double body_pose[16],
                               int clear_first, int64_t* n_added) {
  Handle* h = reinterpret_cast<Handle*>(hh);
  H_CHECK(h);
  if (n < 0 || (n > 0 && !sweep_xyzw) || !body_pose || !(box_half > 0.f) || n > (1ll << 28))
    return fail(h, AICP_B200_ERR_BAD_ARG, "accumulate_sweep: bad arguments");
  const float4* pts = nullptr;
  int rc;
  if (n > 0 && (rc = upload_points(h, h->tmp_a, sweep_xyzw, n, &pts))) return rc;
  int64_t added = 0;
  rc = run_accumulate_sweep(h, pts, n, box_half, body_pose, clear_first, &added);
  if (n_added) *n_added = added;
  return rc;
}

const float* aicp_b200_get_accumulated(aicp_b200_handle* hh, int64_t* n_out) {
  Handle* h = reinterpret_cast<Handle*>(hh);
  if (!h) return nullptr;
  if (n_out) *n_out = h->acc_n;
  return reinterpret_cast<const float*>(h->acc.p);
}

int aicp_b200_download_accumulated(aicp_b200_handle* hh, float* xyzw, int64_t n) {
  Handle* h = reinterpret_cast<Handle*>(hh);
  H_CHECK(h);
  if (n != h->acc_n || (n > 0 && !xyzw)) return fail(h, AICP_B200_ERR_BAD_ARG, "download_accumulated: %lld points are accumulated", (long long)h->acc_n);
  if (n == 0) return AICP_B200_OK;
  return download(h, xyzw, h->acc.p, sizeof(float4) * (size_t)n);
}

int aicp_b200_read_pcd(const char* path, float* out_xyzw, int64_t capacity, int64_t* n_out, char* err, int err_len) {
  if (!path || !n_out) return AICP_B200_ERR_BAD_ARG;
  std::string e;
  int rc;
  try { rc = read_pcd(path, out_xyzw, capacity, n_out, &e); }
  catch (const std::exception& ex) { e = std::string("aicp_b200_read_pcd: ") + ex.what(); rc = AICP_B200_ERR_CONFIG; }      // no exception may cross the C ABI
  if (err && err_len > 0) snprintf(err, (size_t)err_len, "%s", e.c_str());
  return rc;
}

int aicp_b200_read_ply(const char* path, float* out_xyzw, int64_t capacity, int64_t* n_out, char* err, int err_len) {
  if (!path || !n_out) return AICP_B200_ERR_BAD_ARG;
  std::string e;
  int rc;
  try { rc = read_ply(path, out_xyzw, capacity, n_out, &e); }
  catch (const std::exception& ex) { e = std::string("aicp_b200_read_ply: ") + ex.what(); rc = AICP_B200_ERR_CONFIG; }      // no exception may cross the C ABI
  if (err && err_len > 0) snprintf(err, (size_t)err_len, "%s", e.c_str());
  return rc;
}

int aicp_b200_write_pcd(const char* path, const float* xyzw, int64_t n, char* err, int err_len) {
  if (!path || n < 0 || (n > 0 && !xyzw)) return AICP_B200_ERR_BAD_ARG;
  std::string e;
  int rc;
  try { rc = write_pcd_binary(path, xyzw, n, &e); }
  catch (const std::exception& ex) { e = std::string("aicp_b200_write_pcd: ") + ex.what(); rc = AICP_B200_ERR_CONFIG; }
  if (err && err_len > 0) snprintf(err, (size_t)err_len, "%s", e.c_str());
  return rc;
}

int aicp_b200_read_pose_file(const char* path, int64_t* rows, double* poses, int64_t capacity, int64_t* n_out, char* err, int err_len) {
  if (!path || !n_out) return AICP_B200_ERR_BAD_ARG;
  std::string e;
  int rc;
  try { rc = read_pose_file(path, rows, poses, capacity, n_out, &e); }
  catch (const std::exception& ex) { e = std::string("aicp_b200_read_pose_file: ") + ex.what(); rc = AICP_B200_ERR_CONFIG; }      // no exception may cross the C ABI
  if (err && err_len > 0) snprintf(err, (size_t)err_len, "%s", e.c_str());
  return rc;
}

int aicp_b200_fov_overlap(aicp_b200_handle* hh, const float* a_xyzw, int64_t n_a, const float* b_xyzw, int64_t n_b, const double pose_a[16],
                          const double pose_b[16], float range, float angular_view, float* out_a, float* out_b, int64_t counts[2],
                          float* overlap_pct) {
  Handle* h = reinterpret_cast<Handle*>(hh);
  H_CHECK(h);
  if (n_a < 0 || n_b < 0 || (n_a > 0 && !a_xyzw) || (n_b > 0 && !b_xyzw) || !pose_a || !pose_b || !overlap_pct)
    return fail(h, AICP_B200_ERR_BAD_ARG, "fov_overlap: bad arguments");
  const float4 *A = nullptr, *B = nullptr;
  int rc;
  if (n_a > 0 && (rc = upload_points(h, h->tmp_a, a_xyzw, n_a, &A))) return rc;
  if (n_b > 0 && (rc = upload_points(h, h->tmp_b, b_xyzw, n_b, &B))) return rc;
  if ((rc = run_fov_overlap(h, A, n_a, B, n_b, pose_a, pose_b, range, angular_view, overlap_pct, counts))) return rc;
  if (out_a && h->al_fov_n[0] > 0 && (rc = download(h, out_a, h->al_fov[0].p, sizeof(float4) * (size_t)h->al_fov_n[0]))) return rc;
  if (out_b && h->al_fov_n[1] > 0 && (rc = download(h, out_b, h->al_fov[1].p, sizeof(float4) * (size_t)h->al_fov_n[1]))) return rc;
  return AICP_B200_OK;
}

const float* aicp_b200_get_fov_filtered(aicp_b200_handle* hh, int which, int64_t* n_out) {
  Handle* h = reinterpret_cast<Handle*>(hh);
  if (!h || which < 0 || which > 1) return nullptr;
  if (n_out) *n_out = h->al_fov_n[which];
  return reinterpret_cast<const float*>(h->al_fov[which].p);
}

int aicp_b200_alignability(aicp_b200_handle* hh, const float* a_xyzw, int64_t n_a, const float* b_xyzw, int64_t n_b, const double pose_a[16],
                           const double pose_b[16], const aicp_b200_prefilter_config* cfg, float* alignability_pct, int32_t* matching,
                           int64_t matching_cap, int64_t info[3]) {
  Handle* h = reinterpret_cast<Handle*>(hh);
  H_CHECK(h);
  if (n_a < 0 || n_b < 0 || (n_a > 0 && !a_xyzw) || (n_b > 0 && !b_xyzw) || !pose_a || !pose_b || !alignability_pct)
    return fail(h, AICP_B200_ERR_BAD_ARG, "alignability: bad arguments");
  aicp_b200_prefilter_config def;
  if (!cfg) { aicp_b200_prefilter_default_config(&def); cfg = &def; }
  const float4 *A = nullptr, *B = nullptr;
  int rc;
  // the pre-filter stages host input in tmp_a itself; keep two separate staging buffers for the two clouds
  if (n_a > 0 && (rc = upload_points(h, h->tmp_a, a_xyzw, n_a, &A))) return rc;
  if (n_b > 0 && (rc = upload_points(h, h->tmp_b, b_xyzw, n_b, &B))) return rc;
  int64_t inf[3] = {0, 0, 0};
  std::vector<int32_t> m((size_t)(n_b > 0 ? n_b : 1), -1);       // one entry per kept cluster of B: never more than points
  if ((rc = run_alignability(h, A, n_a, B, n_b, pose_a, pose_b, cfg, alignability_pct, m.data(), inf))) return rc;
  for (int64_t j = 0; matching && j < inf[1] && j < matching_cap; ++j) matching[j] = m[(size_t)j];
  if (info) { info[0] = inf[0]; info[1] = inf[1]; info[2] = inf[2]; }
  return AICP_B200_OK;
}

int aicp_b200_alignment_risk(aicp_b200_handle* hh, const float* ref_xyzw, int64_t n_ref, const float* read_xyzw, int64_t n_read,
                             const double ref_pose[16], const double read_pose[16], float range, float angular_view, float octree_overlap_pct,
                             float* fov_overlap_pct, float* alignability_pct, double* risk) {
  Handle* h = reinterpret_cast<Handle*>(hh);
  H_CHECK(h);
  if (n_ref < 1 || n_read < 1 || !ref_xyzw || !read_xyzw || !ref_pose || !read_pose || !risk)
    return fail(h, AICP_B200_ERR_BAD_ARG, "alignment_risk: bad arguments");
  const float4 *A, *B;
  int rc;
  if ((rc = upload_points(h, h->tmp_a, ref_xyzw, n_ref, &A))) return rc;
  if ((rc = upload_points(h, h->tmp_b, read_xyzw, n_read, &B))) return rc;
  float fov = 0.f, al = 0.f;
  if ((rc = run_fov_overlap(h, A, n_ref, B, n_read, ref_pose, read_pose, range, angular_view, &fov, nullptr))) return rc;
  aicp_b200_prefilter_config cfg;
  aicp_b200_prefilter_default_config(&cfg);
  if ((rc = run_alignability(h, h->al_fov[0].p, h->al_fov_n[0], h->al_fov[1].p, h->al_fov_n[1], ref_pose, read_pose, &cfg, &al, nullptr, nullptr))) return rc;
  const double feat[2] = {(double)(float)octree_overlap_pct, (double)(float)al};      // app.cpp:175-176
  if ((rc = svm_predict(h, feat, 1, 2, risk, nullptr))) return rc;
  if (fov_overlap_pct) *fov_overlap_pct = fov;
  if (alignability_pct) *alignability_pct = al;
  return AICP_B200_OK;
}

int aicp_b200_svm_parse(const char* model_xml_path, aicp_b200_svm_summary* out, char* err, int err_len) {
  if (!model_xml_path || !out) return AICP_B200_ERR_BAD_ARG;
  std::string e;
  int rc = svm_parse_summary(model_xml_path, out, &e);
  if (err && err_len > 0) { snprintf(err, (size_t)err_len, "%s", e.c_str()); }
  return rc;
}

int aicp_b200_svm_load(aicp_b200_handle* hh, const char* model_xml_path) {
  Handle* h = reinterpret_cast<Handle*>(hh);
  H_CHECK(h);
  if (!model_xml_path || !*model_xml_path) return fail(h, AICP_B200_ERR_BAD_ARG, "svm_load: empty path");
  return svm_load(h, model_xml_path);
}

int aicp_b200_svm_info(aicp_b200_handle* hh, int32_t* dim, int32_t* sv_total) {
  Handle* h = reinterpret_cast<Handle*>(hh);
  H_CHECK(h);
  return svm_info(h, dim, sv_total);
}

int aicp_b200_svm_predict(aicp_b200_handle* hh, const double* features, int64_t n, int32_t dim, double* probabilities, float* raw) {
  Handle* h = reinterpret_cast<Handle*>(hh);
  H_CHECK(h);
  if (n < 0 || (n > 0 && (!features || !probabilities))) return fail(h, AICP_B200_ERR_BAD_ARG, "svm_predict: bad arguments");
  return svm_predict(h, features, n, dim, probabilities, raw);
}

float aicp_b200_autotune_ratio(float overlap_pct) {
  // app.cpp:198-202
  float current_ratio = overlap_pct / 100.0;
  if (current_ratio < 0.25) current_ratio = 0.25;
  else if (current_ratio > 0.70) current_ratio = 0.70;
  // fileIO.cpp:194-198: std::stringstream << float (precision 6, general format), parsed back by libpointmatcher
  char buf[64];
  snprintf(buf, sizeof(buf), "%g", (double)current_ratio);
  return strtof(buf, nullptr);
}

// common implementation of the two batch entry points; origins != nullptr: one AICP step per pair (overlap -> clamp and
// 6-digit round trip -> registration with that ratio), else registration only with the given / configured ratios
// optional alignment-risk stage of the batched pipeline (App::runAicpPipeline with failure_prediction_mode, app.cpp:232-246)
struct RiskArgs {
  const double* ref_poses;       // n_pairs x 16, column-major
  const double* read_poses;
  float range, angular_view;
  const char* model_path;
  double threshold;
  float* out_alignability;
  double* out_risk;
  int prefilter_first;           // the inputs are raw clouds: App::setAndFilterReading / filterCloud first (app.cpp:77-110)
  int64_t* out_n_filtered;       // nullable, n_pairs x 2: points left after the pre-filter
};

static int batch_impl(aicp_b200_handle* hh, int64_t n_pairs, const float* const* ref_xyzw, const int64_t* n_ref,
                      const float* const* read_xyzw, const int64_t* n_read, const float* ratios, const double* ref_origins,
                      const double* read_origins, double resolution, int streams, float* out_T, float* out_overlap,
                      aicp_b200_stats* stats, int32_t* status, float* batch_ms, const RiskArgs* risk = nullptr) {
  Handle* h = reinterpret_cast<Handle*>(hh);
  H_CHECK(h);
  if (n_pairs < 0 || (n_pairs > 0 && (!ref_xyzw || !n_ref || !read_xyzw || !n_read || !out_T)))
    return fail(h, AICP_B200_ERR_BAD_ARG, "register_batch: bad arguments");
  if (batch_ms) *batch_ms = 0.f;
  if (n_pairs == 0) return AICP_B200_OK;
  int rc = load_config(h);     // one parse for the whole batch
  if (rc) return rc;
  if (streams <= 0) streams = 4;
  if (streams > 32) streams = 32;
  if ((int64_t)streams > n_pairs) streams = (int)n_pairs;
  while ((int)h->workers.size() < streams) {
    aicp_b200_handle* w = nullptr;
    rc = aicp_b200_create(nullptr, h->device, &w);
    if (rc) return fail(h, rc, "register_batch: cannot create worker: %s", aicp_b200_last_error(nullptr));
    Handle* wh = reinterpret_cast<Handle*>(w);
    if (cudaEventCreate(&wh->done_ev) != cudaSuccess) return fail(h, AICP_B200_ERR_CUDA, "register_batch: event creation failed");
    h->workers.push_back(wh);
  }
  if (!h->batch_ev[0]) { CUDA_TRY(cudaEventCreate(&h->batch_ev[0])); CUDA_TRY(cudaEventCreate(&h->batch_ev[1])); }
  // device-side span: every worker stream starts after batch_ev[0]; the master stream ends after every worker
  CUDA_TRY(cudaEventRecord(h->batch_ev[0], h->stream));
  for (int w = 0; w < streams; ++w) {
    Handle* wh = h->workers[w];
    wh->cfg = h->cfg; wh->cfg_from_file = false;
    wh->profiling = h->profiling; wh->trace_matches = false;
    wh->batch_worker = streams > 1; wh->batch_streams = streams; wh->knn_schedule = h->knn_schedule; wh->match_schedule = h->match_schedule;
    wh->loop_schedule = h->loop_schedule;
    if (risk && wh->svm_path != risk->model_path) {
      if ((rc = svm_load(wh, risk->model_path))) return fail(h, rc, "pipeline_batch: %s", wh->last_error.c_str());
      wh->svm_path = risk->model_path;
    }
    CUDA_TRY(cudaStreamWaitEvent(wh->stream, h->batch_ev[0], 0));
  }
  std::atomic<int64_t> next(0);
  std::atomic<int> first_err(0);
  std::vector<std::string> errs((size_t)streams);
  auto work = [&](int w) {
    Handle* wh = h->workers[w];
    cudaSetDevice(wh->device);
    for (;;) {
      int64_t i = next.fetch_add(1);
      if (i >= n_pairs) break;
      if (ratios) wh->cfg.ratio = ratios[i];
      int r = AICP_B200_OK;
      if (ref_origins || risk) {
        // App::runAicpPipeline (app.cpp:218-247): computeOverlap, [computeAlignmentRisk,] then computeRegistration with the
        // auto-tuned ratio; both clouds are staged once into the worker's own buffers and used by every stage
        float ov = 0.f;
        double origin_a[3], origin_b[3];
        const double* oa = ref_origins ? ref_origins + 3 * i : origin_a;
        const double* ob = read_origins ? read_origins + 3 * i : origin_b;
        if (risk) for (int d = 0; d < 3; ++d) { origin_a[d] = risk->ref_poses[16 * i + 12 + d]; origin_b[d] = risk->read_poses[16 * i + 12 + d]; }
        if (!ref_xyzw[i] || !read_xyzw[i] || n_ref[i] < 1 || n_read[i] < 1 || n_ref[i] > (1ll << 30) || n_read[i] > (1ll << 30))
          r = fail(wh, AICP_B200_ERR_BAD_ARG, "aicp_batch: null or empty cloud in pair %lld", (long long)i);
        int64_t nr = n_ref[i], nq = n_read[i];
        if (!r && risk && risk->prefilter_first) {
          // both raw clouds through regionGrowingUniformPlaneSegmentationFilter; the filtered clouds never leave the device
          aicp_b200_prefilter_config pcfg;
          aicp_b200_prefilter_default_config(&pcfg);
          const float* raw[2] = {ref_xyzw[i], read_xyzw[i]};
          const int64_t raw_n[2] = {n_ref[i], n_read[i]};
          DevBuf<float4>* dst[2] = {&wh->ref_in, &wh->read_in};
          int64_t* cnt[2] = {&nr, &nq};
          for (int c = 0; c < 2 && !r; ++c) {
            const float4* pts;
            r = upload_points(wh, wh->tmp_a, raw[c], raw_n[c], &pts);
            if (!r) r = run_prefilter(wh, pts, raw_n[c], &pcfg, nullptr, nullptr);
            if (!r && wh->pf_n_out < 1) r = fail(wh, AICP_B200_ERR_BAD_ARG, "pipeline_batch: the pre-filter left no points of a cloud of pair %lld", (long long)i);
            if (!r) r = stage_owned(wh, *dst[c], reinterpret_cast<const float*>(wh->pf_out.p), wh->pf_n_out);
            *cnt[c] = wh->pf_n_out;
          }
          if (risk->out_n_filtered) { risk->out_n_filtered[2 * i] = nr; risk->out_n_filtered[2 * i + 1] = nq; }
        } else {
          if (!r) r = stage_owned(wh, wh->ref_in, ref_xyzw[i], n_ref[i]);
          if (!r) r = stage_owned(wh, wh->read_in, read_xyzw[i], n_read[i]);
        }
        if (!r) r = run_overlap(wh, wh->ref_in.p, nr, oa, wh->read_in.p, nq, ob, resolution, &ov, nullptr);
        if (out_overlap) out_overlap[i] = ov;
        bool skip = false;
        if (!r && risk) {
          // computeAlignmentRisk (app.cpp:143-185); the registration runs only when the risk is at most the threshold (:241-243)
          float fov = 0.f, al = 0.f;
          double rk = 0.0;
          aicp_b200_prefilter_config pcfg;
          aicp_b200_prefilter_default_config(&pcfg);
          r = run_fov_overlap(wh, wh->ref_in.p, nr, wh->read_in.p, nq, risk->ref_poses + 16 * i, risk->read_poses + 16 * i,
                              risk->range, risk->angular_view, &fov, nullptr);
          if (!r) r = run_alignability(wh, wh->al_fov[0].p, wh->al_fov_n[0], wh->al_fov[1].p, wh->al_fov_n[1], risk->ref_poses + 16 * i,
                                       risk->read_poses + 16 * i, &pcfg, &al, nullptr, nullptr);
          const double feat[2] = {(double)ov, (double)al};
          if (!r) r = svm_predict(wh, feat, 1, 2, &rk, nullptr);
          if (risk->out_alignability) risk->out_alignability[i] = al;
          if (risk->out_risk) risk->out_risk[i] = rk;
          skip = !r && rk > risk->threshold;
          if (skip) {
            for (int q = 0; q < 16; ++q) out_T[16 * i + q] = (q % 5 == 0) ? 1.f : 0.f;       // T stays the identity it was created as (app.cpp:356)
            if (stats) memset(stats + i, 0, sizeof(aicp_b200_stats));
          }
        }
        if (!r && !skip) {
          wh->cfg.ratio = aicp_b200_autotune_ratio(ov);
          wh->n_ref = nr; wh->n_read = nq;
          r = run_registration(wh, nullptr, true, stats ? stats + i : nullptr, out_T + 16 * i);
        }
      } else {
        r = aicp_b200_register(reinterpret_cast<aicp_b200_handle*>(wh), ref_xyzw[i], n_ref[i], read_xyzw[i], n_read[i], nullptr,
                               out_T + 16 * i, stats ? stats + i : nullptr);
      }
      if (status) status[i] = r;
      if (r) {
        int expected = 0;
        if (first_err.compare_exchange_strong(expected, r)) errs[(size_t)w] = wh->last_error;
      }
    }
    cudaEventRecord(wh->done_ev, wh->stream);
  };
  std::vector<std::thread> threads;
  for (int w = 1; w < streams; ++w) threads.emplace_back(work, w);
  work(0);
  for (auto& t : threads) t.join();
  for (int w = 0; w < streams; ++w) CUDA_TRY(cudaStreamWaitEvent(h->stream, h->workers[w]->done_ev, 0));
  CUDA_TRY(cudaEventRecord(h->batch_ev[1], h->stream));
  CUDA_TRY(cudaStreamSynchronize(h->stream));
  if (batch_ms) CUDA_TRY(cudaEventElapsedTime(batch_ms, h->batch_ev[0], h->batch_ev[1]));
  if (first_err.load()) {
    for (const std::string& e : errs) if (!e.empty()) { h->last_error = "register_batch: " + e; break; }
    return first_err.load();
  }
  return AICP_B200_OK;
}

int aicp_b200_register_batch(aicp_b200_handle* hh, int64_t n_pairs, const float* const* ref_xyzw, const int64_t* n_ref,
                             const float* const* read_xyzw, const int64_t* n_read, const float* ratios, int streams,
                             float* out_T, aicp_b200_stats* stats, int32_t* status, float* batch_ms) {
  return batch_impl(hh, n_pairs, ref_xyzw, n_ref, read_xyzw, n_read, ratios, nullptr, nullptr, 0.0, streams, out_T, nullptr, stats,
                    status, batch_ms);
}

int aicp_b200_register_batch_devices(aicp_b200_handle* hh, const int32_t* devices, int32_t n_devices, int64_t n_pairs,
                                     const float* const* ref_xyzw, const int64_t* n_ref, const float* const* read_xyzw,
                                     const int64_t* n_read, const float* ratios, int streams, float* out_T, aicp_b200_stats* stats,
                                     int32_t* status, float* batch_ms) {
  Handle* h = reinterpret_cast<Handle*>(hh);
  H_CHECK(h);
  if (!devices || n_devices < 1 || n_devices > 64 || n_pairs < 0 || (n_pairs > 0 && (!ref_xyzw || !n_ref || !read_xyzw || !n_read || !out_T)))
    return fail(h, AICP_B200_ERR_BAD_ARG, "register_batch_devices: bad arguments");
  if (batch_ms) *batch_ms = 0.f;
  if (n_pairs == 0) return AICP_B200_OK;
  int rc = load_config(h);
  if (rc) return rc;
  // one child handle per device, kept across calls (each grows its own worker pool inside batch_impl)
  std::vector<Handle*> child((size_t)n_devices, nullptr);
  for (int d = 0; d < n_devices; ++d) {
    for (Handle* c : h->device_children) if (c->device == devices[d]) child[(size_t)d] = c;
    for (int e = 0; e < d; ++e) if (devices[e] == devices[d]) return fail(h, AICP_B200_ERR_BAD_ARG, "register_batch_devices: device %d listed twice", devices[d]);
    if (!child[(size_t)d]) {
      aicp_b200_handle* c = nullptr;
      rc = aicp_b200_create(nullptr, devices[d], &c);
      if (rc) return fail(h, rc, "register_batch_devices: %s", aicp_b200_last_error(nullptr));
      child[(size_t)d] = reinterpret_cast<Handle*>(c);
      h->device_children.push_back(child[(size_t)d]);
    }
    Handle* c = child[(size_t)d];
    c->cfg = h->cfg; c->cfg_from_file = false;
    c->profiling = h->profiling; c->knn_schedule = h->knn_schedule; c->match_schedule = h->match_schedule; c->loop_schedule = h->loop_schedule;
  }
  std::vector<int> rcs((size_t)n_devices, 0);
  std::vector<float> ms((size_t)n_devices, 0.f);
  auto run = [&](int d) {
    // the pairs of this device, i = d, d + G, d + 2G, ...: contiguous copies of the argument arrays, results scattered back
    std::vector<const float*> refs, reads;
    std::vector<int64_t> nr, nq, idx;
    std::vector<float> rat;
    for (int64_t i = d; i < n_pairs; i += n_devices) {
      idx.push_back(i); refs.push_back(ref_xyzw[i]); reads.push_back(read_xyzw[i]); nr.push_back(n_ref[i]); nq.push_back(n_read[i]);
      if (ratios) rat.push_back(ratios[i]);
    }
    const int64_t m = (int64_t)idx.size();
    if (m == 0) return;
    std::vector<float> T((size_t)m * 16);
    std::vector<int32_t> stt((size_t)m, 0);
    aicp_b200_stats* sub = stats ? static_cast<aicp_b200_stats*>(malloc(sizeof(aicp_b200_stats) * (size_t)m)) : nullptr;
    rcs[(size_t)d] = batch_impl(reinterpret_cast<aicp_b200_handle*>(child[(size_t)d]), m, refs.data(), nr.data(), reads.data(), nq.data(),
                                ratios ? rat.data() : nullptr, nullptr, nullptr, 0.0, streams, T.data(), nullptr, sub, stt.data(), &ms[(size_t)d]);
    for (int64_t j = 0; j < m; ++j) {
      memcpy(out_T + 16 * idx[(size_t)j], T.data() + 16 * j, 16 * sizeof(float));
      if (status) status[idx[(size_t)j]] = stt[(size_t)j];
      if (stats) stats[idx[(size_t)j]] = sub[j];
    }
    free(sub);
  };
  std::vector<std::thread> threads;
  for (int d = 1; d < n_devices; ++d) threads.emplace_back(run, d);
  run(0);
  for (auto& t : threads) t.join();
  cudaSetDevice(h->device);
  int first = 0;
  for (int d = 0; d < n_devices; ++d) {
    if (batch_ms && ms[(size_t)d] > *batch_ms) *batch_ms = ms[(size_t)d];
    if (rcs[(size_t)d] && !first) { first = rcs[(size_t)d]; h->last_error = "device " + std::to_string(devices[d]) + ": " + child[(size_t)d]->last_error; }
  }
  return first;
}

int aicp_b200_aicp_batch(aicp_b200_handle* hh, int64_t n_pairs, const float* const* ref_xyzw, const int64_t* n_ref,
                         const double* ref_origins, const float* const* read_xyzw, const int64_t* n_read,
                         const double* read_origins, double resolution, int streams, float* out_T, float* out_overlap,
                         aicp_b200_stats* stats, int32_t* status, float* batch_ms) {
  if (!ref_origins || !read_origins) return AICP_B200_ERR_BAD_ARG;
  return batch_impl(hh, n_pairs, ref_xyzw, n_ref, read_xyzw, n_read, nullptr, ref_origins, read_origins, resolution, streams, out_T,
                    out_overlap, stats, status, batch_ms);
}

int aicp_b200_pipeline_batch(aicp_b200_handle* hh, int64_t n_pairs, const float* const* ref_xyzw, const int64_t* n_ref,
                             const double* ref_poses, const float* const* read_xyzw, const int64_t* n_read, const double* read_poses,
                             double resolution, float sensor_range, float angular_view, const char* svm_model_path, double risk_threshold,
                             int prefilter_first, int streams, float* out_T, float* out_overlap, float* out_alignability, double* out_risk,
                             int64_t* out_n_filtered, aicp_b200_stats* stats, int32_t* status, float* batch_ms) {
  if (!ref_poses || !read_poses || !svm_model_path || !*svm_model_path) return AICP_B200_ERR_BAD_ARG;
  RiskArgs ra{ref_poses, read_poses, sensor_range, angular_view, svm_model_path, risk_threshold, out_alignability, out_risk,
              prefilter_first, out_n_filtered};
  return batch_impl(hh, n_pairs, ref_xyzw, n_ref, read_xyzw, n_read, nullptr, nullptr, nullptr, resolution, streams, out_T, out_overlap,
                    stats, status, batch_ms, &ra);
}

}  // extern "C"
