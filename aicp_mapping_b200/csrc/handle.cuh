// handle.cuh -- host-side state behind an aicp_b200_handle: device buffers that persist across calls (no per-call
// cudaMalloc on the steady-state path), the stream, the parsed libpointmatcher chain and the last error text.
#pragma once

#include <atomic>
#include <string>
#include <vector>

#include "common.cuh"
#include "search.cuh"

namespace aicp {

// bumped by every (re)allocation of a device buffer: a captured CUDA graph holds raw pointers, so a graph is only replayed
// while this number is what it was when the graph was captured (icp.cu, setup graph of the batch workers)
inline std::atomic<unsigned long long> g_alloc_generation{0};

// grow-only device buffer
template <typename T>
struct DevBuf {
  T* p = nullptr;
  size_t cap = 0;
  cudaError_t reserve(size_t n) {
    if (n <= cap) return cudaSuccess;
    g_alloc_generation.fetch_add(1, std::memory_order_relaxed);
    if (p) cudaFree(p);
    p = nullptr; cap = 0;
    size_t want = n + n / 8 + 64;
    cudaError_t e = cudaMalloc((void**)&p, want * sizeof(T));
    if (e == cudaSuccess) cap = want;
    return e;
  }
  void release() { if (p) { cudaFree(p); g_alloc_generation.fetch_add(1, std::memory_order_relaxed); } p = nullptr; cap = 0; }
};

// what the captured setup of a registration depends on besides the data (icp.cu)
struct SetupKey {
  int n_ref = -1, n_read = -1, knn = 0, has_init = 0, knn_schedule = 0;
  unsigned long long generation = 0;
  bool operator==(const SetupKey& o) const {
    return n_ref == o.n_ref && n_read == o.n_read && knn == o.knn && has_init == o.has_init && knn_schedule == o.knn_schedule &&
           generation == o.generation;
  }
};

// Morton-ordered spatial index over one cloud
struct SpatialIndex {
  DevBuf<float4> pts;            // Morton order, .w = original index
  DevBuf<float4> rec;            // 4 float4 per internal node of the radix tree (see search.cuh)
  DevBuf<int4> node_meta;        // (first, split, end, -) per internal node, build-time scratch
  DevBuf<unsigned int> keys, keys_alt, vals, vals_alt;
  DevBuf<int> flags;             // parent links (internal nodes | points)
  DevBuf<float4> chunkbox;       // boxes of every 32 / 1024 / 32768 consecutive sorted points (lo, hi pairs)
  DevBuf<int> owner;             // per point: lowest node with > 8 (first n) resp. > 32 (next n) points above it
  DevBuf<float4> cellbox;        // per internal node: shrunk float box of the node's Morton cell (lo, hi), see search.cuh
  DevBuf<unsigned int> sort_tmp;  // radix sort scratch: digit histograms, tickets, per-tile status words
  IndexMeta* meta = nullptr;     // device
  int n = 0;
  IndexView view() const { return IndexView{pts.p, rec.p, owner.p, owner.p + n, cellbox.p, n}; }
  void release() {
    pts.release(); rec.release(); node_meta.release(); keys.release(); keys_alt.release(); vals.release(); vals_alt.release();
    flags.release(); chunkbox.release(); owner.release(); cellbox.release(); sort_tmp.release();
    if (meta) cudaFree(meta);
    meta = nullptr;
  }
};

struct Comm;   // multi-GPU state (comm.cu)
struct SvmModel;   // alignment-risk classifier (svm.cu)

struct Handle {
  int device = 0;
  cudaStream_t stream = nullptr;
  cudaEvent_t ev[4] = {nullptr, nullptr, nullptr, nullptr};
  cudaStream_t side = nullptr;   // reading-side setup of a single registration runs here, beside the reference's normals
  cudaEvent_t ev_fork = nullptr, ev_init = nullptr, ev_join = nullptr;
  aicp_b200_icp_config cfg;
  std::string cfg_path;
  bool cfg_from_file = false;
  std::string last_error;
  int launches = 0;

  // reference side
  DevBuf<float4> ref_in;
  SpatialIndex ref_ix;           // original frame (normals search)
  DevBuf<float4> refc_pts;       // centred points, Morton order
  DevBuf<float4> refc_rec;       // centred tree records
  DevBuf<float4> refc_cell;      // centred cell boxes
  DevBuf<float4> normals;        // Morton order (nx,ny,nz,density)
  DevBuf<int> knn_pos;           // n x knn neighbour positions (Morton order), scratch of the normals filter
  DevBuf<float> ref_rk2;         // per reference point (Morton order): squared distance to its last k-NN list entry
  bool ref_recentre = false;     // the reference changed by an append: centred copies and mean must be refreshed
  DevBuf<float4> app_pts, app_normals;                       // merge targets of an append (swapped with the live arrays)
  DevBuf<unsigned int> app_keys, app_vals, app_new, app_flag, app_scan, app_tiles;
  DevBuf<float> app_rk2, app_rmax;
  DevBuf<int> app_list;
  IndexMeta* app_meta = nullptr;
  int64_t n_ref = 0;
  bool ref_ready = false;
  int ref_knn = 0;

  // reading side
  DevBuf<float4> read_in, read0, read_out, read_init;
  SpatialIndex read_ix;          // reading' in Morton order (.w = original reading index): coherent tree walks per warp
  DevBuf<int> match_pos;
  DevBuf<float> d2;
  DevBuf<unsigned int> hist;
  DevBuf<unsigned long long> acc_slots;   // normal-equation partial sums: ACC_SLOTS x 32 x (lo, hi), all zero between launches
  DevBuf<unsigned int> cand;     // d2 keys of the quantile's first-digit bin (k_select23)
  int64_t n_read = 0;
  bool has_init_reading = false;

  // loop state
  DeviceState* st = nullptr;     // device
  DeviceState* st_host = nullptr;  // pinned
  volatile int* progress_host = nullptr;   // mapped pinned: [0] iterations completed, [1] loop finished (written by the device)
  int* progress_dev = nullptr;             // the same two words, device address
  DevBuf<int> trace_idx;
  bool trace_matches = false;
  int knn_schedule = 0;          // SurfaceNormal k-NN kernel: 0 auto, 1 warp per query (latency), 2 tile per warp (throughput)
  int match_schedule = 0;        // correspondence search kernel: 0 auto, 1 one query per thread (k_match), 2 one tile per warp (k_match_tile)
  bool batch_worker = false;     // this handle is one of the concurrent workers of aicp_b200_register_batch
  int batch_streams = 1;         // how many such workers share the GPU (the persistent loop kernel takes 1 / batch_streams of its blocks)
  int loop_schedule = 0;         // ICP loop: 0 auto, 1 three launches per iteration + host look-ahead, 2 one persistent cooperative kernel
  bool fused_tail = true;        // multi-launch loop: k_quantile_accumulate instead of k_select23 + k_accumulate (AICP_B200_FUSED_TAIL=0: off)
  int loop_spread = 0;           // experiments (AICP_B200_SPREAD): lanes per query in the search phase of the loop kernel, 0 = automatic
  int loop_occ[2] = {0, 0};      // co-resident blocks per SM of k_icp_loop<false / true>
  int n_sm = 0;
  int profiling = 0;             // 0 off, 1 CUDA events around k_match only, 2 around every stage
  std::vector<cudaEvent_t> prof_ev;   // 3 setup + 4 per iteration
  int64_t trace_iters = 0, trace_n = 0;

  // scratch for stage entry points and overlap
  SpatialIndex tmp_ix;
  DevBuf<float4> tmp_a, tmp_b;
  DevBuf<int> tmp_i;
  DevBuf<float> tmp_f;
  DevBuf<unsigned int> ovl_bits_a, ovl_bits_b;
  DevBuf<unsigned long long> ovl_counts;
  DevBuf<unsigned long long> crop_status;   // chained-scan tile status of the crop box + total + ticket
  int64_t crop_n = 0;
  volatile unsigned long long* crop_total_host = nullptr;   // mapped pinned: the number of kept points, written by the crop kernels
  unsigned long long* crop_total_dev = nullptr;
  DevBuf<float4> crop_stash;                 // kept points per CTA run, before k_crop_gather puts them in input order
  bool crop_legacy = false;                  // AICP_B200_CROP=legacy: the round-1 kernel (A/B measurements)
  DevBuf<float4> map;                       // persistent device-resident map (aicp_b200_map_*), original point order
  int64_t map_n = 0;
  DevBuf<float4> crop_out;                  // cropped cloud when the caller asks for a device-resident result

  // pre-filter (prefilter.cu): VoxelGrid -> normals -> region growing
  SpatialIndex pf_ix;                       // Morton index over the voxel-grid output
  DevBuf<float4> pf_sampled;                // voxel centroids, ascending voxel index ("sampled order")
  DevBuf<float4> pf_normals, pf_normals_orig;   // (nx, ny, nz, curvature): Morton order / sampled order
  DevBuf<float4> pf_out;                    // kept clusters, concatenated
  DevBuf<unsigned int> pf_keys, pf_keys_alt, pf_vals, pf_vals_alt, pf_sort_tmp, pf_flag, pf_slot, pf_tiles, pf_mask, pf_mutual, pf_count;
  DevBuf<int> pf_label, pf_seed_pos, pf_labels_out, pf_parent, pf_root, pf_clabel;
  DevBuf<unsigned long long> pf_status;     // chained-scan tile status of k_vg_fused + its ticket
  void* pf_meta = nullptr;                  // PfMeta, device
  void* pf_meta_host = nullptr;             // pinned
  int64_t pf_n_sampled = 0, pf_n_out = 0, pf_n_clusters = 0;
  bool pf_has_segments = false;
  cudaEvent_t pf_ev[2] = {nullptr, nullptr};

  // FOV overlap filter + alignability (alignability.cu)
  DevBuf<float4> al_moved, al_fov[2], al_pts[2], al_mean;
  int64_t al_fov_n[2] = {0, 0};
  DevBuf<int> al_lab[2], al_ext;
  DevBuf<unsigned long long> al_sums;
  DevBuf<unsigned int> al_cnt, al_counts;
  DevBuf<float> al_axes, al_boxes[2];
  Handle* al_child = nullptr;    // second stream + buffers for the other cloud of an alignability call (not inside batches)

  // sweep accumulation (ingest.cu)
  DevBuf<float4> acc, acc_tmp;
  int64_t acc_n = 0;

  Comm* comm = nullptr;
  SvmModel* svm = nullptr;
  std::string svm_path;          // model loaded into a batch worker

  // batch workers: one child handle (own stream + buffers) per concurrent registration
  std::vector<Handle*> workers;
  std::vector<Handle*> device_children;   // aicp_b200_register_batch_devices: one child handle (with its own workers) per GPU
  cudaEvent_t batch_ev[2] = {nullptr, nullptr};
  cudaEvent_t wait_ev = nullptr;     // aicp_b200_wait_stream: recorded on the caller's producer stream, waited on by h->stream
  cudaEvent_t done_ev = nullptr;
  // batch workers: the ~30 launches of a registration's setup as one CUDA graph, replayed while sizes and buffers stay put
  cudaGraphExec_t setup_exec = nullptr;
  SetupKey setup_key, setup_seen;
  int setup_launches = 0;
  bool setup_graph_off = false;      // a capture failed on this handle: stay with plain launches
};

// ---- index.cu
// with_tree = false: Morton order only (pts), no radix tree -- enough to make the queries of a warp spatially coherent
int build_index(Handle* h, SpatialIndex& ix, const float4* pts_dev, int64_t n, bool with_tree = true, cudaEvent_t after_stats = nullptr);
int build_tree(Handle* h, SpatialIndex& ix, int n);     // radix tree + boxes over ix.keys / ix.pts (already Morton-ordered)
void launch_index_stats(Handle* h, const float4* pts, int n, IndexMeta* m);
void launch_morton_keys(Handle* h, const SpatialIndex& ix, const float4* pts, int n, unsigned int* keys, unsigned int* vals, unsigned int first_index);
// ---- append.cu
int run_reference_append(Handle* h, const float4* new_pts, int64_t m, aicp_b200_append_info* info);
// ---- sort.cu
int radix_sort_pairs(Handle* h, unsigned int* keys, unsigned int* vals, unsigned int* keys_alt, unsigned int* vals_alt, int n,
                     DevBuf<unsigned int>& scratch);
// ---- normals.cu
int run_surface_normals(Handle* h, const SpatialIndex& ix, int knn, float4* normals_morton, int* knn_out_orig, int q0 = 0, int q1 = -1,
                        const int* qlist = nullptr, float* rk2 = nullptr);
int run_knn(Handle* h, const SpatialIndex& ix, int knn, int* knn_out_orig, int q0 = 0, int q1 = -1, const int* qlist = nullptr);   // lists -> h->knn_pos (Morton positions)
// ---- icp.cu
int run_registration(Handle* h, const float* init_T_host, bool rebuild_reference, aicp_b200_stats* stats, float* out_T);
int run_match_stage(Handle* h, const SpatialIndex& ix, const float4* qry, int64_t n_qry, int* out_idx, float* out_d2);
int scatter_normals(Handle* h, const float4* pts_morton, const float4* normals_morton, int n, float4* out_dev);
int run_trim_stage(Handle* h, const float* d2_dev, int64_t n, float ratio, float* out_limit, int64_t* out_n_valid);
// ---- overlap.cu
int run_overlap(Handle* h, const float4* ref, int64_t n_ref, const double* ref_origin, const float4* read, int64_t n_read,
                const double* read_origin, double resolution, float* overlap_pct, int64_t* counts);
// ---- crop.cu
int run_crop_box(Handle* h, const float4* pts, int64_t n, float bmin, float bmax, const float* rpy, const float* translation,
                 float4* out_dev, int64_t* n_out);
// ---- prefilter.cu
int run_voxel_grid(Handle* h, const float4* pts, int64_t n, float leaf, int64_t* n_out);
int run_prefilter(Handle* h, const float4* pts, int64_t n, const aicp_b200_prefilter_config* cfg, const float* viewpoint,
                  aicp_b200_prefilter_info* info);
int exclusive_scan_u32(Handle* h, const unsigned int* in, unsigned int* out, int n, unsigned int* tile_scratch, unsigned int* total);
// ---- alignability.cu
int run_fov_overlap(Handle* h, const float4* A, int64_t nA, const float4* B, int64_t nB, const double* poseA, const double* poseB,
                    float range, float angular_view, float* overlap_pct, int64_t* counts);
int run_alignability(Handle* h, const float4* A, int64_t nA, const float4* B, int64_t nB, const double* poseA, const double* poseB,
                     const aicp_b200_prefilter_config* cfg, float* out_alignability, int32_t* matching, int64_t* info);
// ---- ingest.cu
int run_accumulate_sweep(Handle* h, const float4* sweep, int64_t n, float half, const double* body_pose, int clear_first, int64_t* n_added);
void pose_to_float_transform(const double* pose, float* T);
int read_pcd(const char* path, float* out, int64_t cap, int64_t* n_out, std::string* err);
int read_ply(const char* path, float* out, int64_t cap, int64_t* n_out, std::string* err);
int write_pcd_binary(const char* path, const float* xyzw, int64_t n, std::string* err);
int read_pose_file(const char* path, int64_t* rows_out, double* poses_out, int64_t cap, int64_t* n_out, std::string* err);
// ---- svm.cu
int svm_load(Handle* h, const char* path);
int svm_parse_summary(const char* path, aicp_b200_svm_summary* out, std::string* err);
int svm_predict(Handle* h, const double* features, int64_t n, int32_t dim, double* probabilities, float* raw);
int svm_info(Handle* h, int32_t* dim, int32_t* sv_total);
void svm_release(Handle* h);
// ---- comm.cu (sharded registration; NCCL is loaded at run time, the library has no link-time dependency on it)
int comm_allreduce_u32(Handle* h, unsigned int* buf, size_t count);
int comm_allreduce_u64(Handle* h, unsigned long long* buf, size_t count);
unsigned long long* comm_limbs(Handle* h);
long long comm_total_reading(Handle* h);
int comm_begin_registration(Handle* h, long long n_read_local, bool want_peer);
int comm_ranks(Handle* h);
int comm_slice(Handle* h, int n, int* q0, int* q1, int* per);          // this rank's slice of n replicated items (multiples of 32)
int comm_allgather_bytes(Handle* h, void* buf, size_t bytes_per_rank);   // in place: rank r's bytes_per_rank at offset r * bytes_per_rank
bool comm_peer_view(Handle* h, PeerView* pv);      // true when the exchange runs over peer-mapped memory inside the loop kernel
// ---- config_yaml.cpp
int parse_icp_yaml(const char* path, aicp_b200_icp_config* cfg, std::string* err);
void default_icp_config(aicp_b200_icp_config* cfg);
// ---- api.cu
bool is_device_ptr(const void* p);
int upload_points(Handle* h, DevBuf<float4>& buf, const float* xyzw, int64_t n, const float4** out_dev);

}  // namespace aicp
