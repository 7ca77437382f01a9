// svm.cu -- the alignment-risk classifier on the GPU (classifier part).
//
// replaces aicp::SVM::load (aicp_core/src/classification/svm.cpp:103-107: cv::ml::SVM::load(file)) and aicp::SVM::test
// (svm.cpp:53-101: per sample, svm_->predict(sample, output, 1) -> raw decision value, then
// probability = 1.0 - 1.0 / (1.0 + exp(-output)), :82), which App::computeAlignmentRisk calls with the two features
// (octree overlap, alignability) of every cloud pair (app.cpp:175-181).  Training (SVM::train = cv::ml::SVM::trainAuto,
// svm.cpp:18-51) is an offline tool of the reference and is not part of this library.
//
// The arithmetic lives in OpenCV (find_package(OpenCV), aicp_core/CMakeLists.txt:13; >= 3.0, unpinned).  Restated from its
// published algorithm [UPSTREAM modules/ml/src/svm.cpp: SVMKernelImpl::calc_non_rbf_base / calc_poly, SVMImpl::PredictBody]:
//   samples are float32; per support vector  s = sum_k (float)(sv[k] * x[k]) accumulated in double;
//   K = (float)(s * gamma + coef0), POLY: K = pow(K, degree);   sum = -rho + sum_i alpha[i] * K[index[i]] in double;
//   two classes + RAW_OUTPUT: the result is (float)sum.
// cv::pow: an integer degree is binary exponentiation in float32 (iPow), restated exactly; a fractional degree (the shipped
// cross-validated models: 3.43) goes through OpenCV's own float32 exp/log -- here pow runs in float64 and is rounded to
// float32; the difference is below the six digits the reference prints (probs_opencv3.txt) and is covered by the golden
// vectors the reference ships (aicp_core/data/classification/probs_opencv3.txt for data/labels/testing_labelled_27Aug.txt).
//
// k_svm_predict: one thread per sample, model staged in shared memory (<= 2048 support vectors x <= 8 features).
#include <math.h>

#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <sstream>
#include <string>
#include <vector>

#include "handle.cuh"

namespace aicp {

struct SvmModel {
  int kernel = 0;                 // 0 LINEAR, 1 POLY
  double degree = 0, gamma = 1, coef0 = 0, rho = 0;
  int dim = 0, sv_total = 0, sv_count = 0;
  std::vector<float> sv;          // sv_total x dim
  std::vector<double> alpha;      // sv_count
  std::vector<int> index;         // sv_count
  // device copies
  DevBuf<float> d_sv;
  DevBuf<double> d_alpha;
  DevBuf<int> d_index;
  DevBuf<float> d_x;
  DevBuf<double> d_prob;
  DevBuf<float> d_raw;
  bool loaded = false;
};

namespace {

// text between <tag> / <tag attr...> and </tag> inside s[from, to); returns false if absent
bool xml_block(const std::string& s, const char* tag, size_t from, size_t to, size_t* b, size_t* e) {
  const std::string open = std::string("<") + tag;
  size_t p = from;
  while (true) {
    p = s.find(open, p);
    if (p == std::string::npos || p >= to) return false;
    const char c = s[p + open.size()];
    if (c == '>' || c == ' ' || c == '\t' || c == '\n' || c == '\r') break;     // not a prefix of a longer tag name
    p += open.size();
  }
  const size_t gt = s.find('>', p);
  if (gt == std::string::npos || gt >= to) return false;
  const std::string close = std::string("</") + tag + ">";
  const size_t q = s.find(close, gt);
  if (q == std::string::npos || q > to) return false;
  *b = gt + 1; *e = q;
  return true;
}

bool xml_number(const std::string& s, const char* tag, size_t from, size_t to, double* out) {
  size_t b, e;
  if (!xml_block(s, tag, from, to, &b, &e)) return false;
  char* end = nullptr;
  const std::string t = s.substr(b, e - b);
  *out = strtod(t.c_str(), &end);
  return end != t.c_str();
}

template <typename T>
void parse_numbers(const std::string& s, size_t b, size_t e, std::vector<T>* out) {
  const std::string t = s.substr(b, e - b);
  const char* p = t.c_str();
  char* end = nullptr;
  while (true) {
    const double v = strtod(p, &end);
    if (end == p) break;
    out->push_back((T)v);
    p = end;
  }
}

}  // namespace

// cv::ml::SVM::load for the files AICP ships: OpenCV 3 format (<opencv_ml_svm>, svmType) and the legacy 2.4 format
// (<my_svm type_id="opencv-ml-svm">, svm_type), C_SVC with two classes, LINEAR or POLY kernel
static int svm_parse(const char* path, SvmModel* m, std::string* err) {
  std::ifstream f(path);
  if (!f.good()) { *err = std::string("cannot open SVM model file ") + path; return AICP_B200_ERR_CONFIG; }
  std::stringstream ss; ss << f.rdbuf();
  const std::string s = ss.str();
  const size_t N = s.size();
  size_t b, e;
  if (!xml_block(s, "svmType", 0, N, &b, &e) && !xml_block(s, "svm_type", 0, N, &b, &e)) { *err = "SVM model: no svmType"; return AICP_B200_ERR_CONFIG; }
  if (s.substr(b, e - b).find("C_SVC") == std::string::npos) { *err = "SVM model: only C_SVC is supported (the reference trains C_SVC, svm.cpp:9)"; return AICP_B200_ERR_CONFIG; }
  size_t kb, ke;
  if (!xml_block(s, "kernel", 0, N, &kb, &ke)) { *err = "SVM model: no kernel"; return AICP_B200_ERR_CONFIG; }
  if (!xml_block(s, "type", kb, ke, &b, &e)) { *err = "SVM model: no kernel type"; return AICP_B200_ERR_CONFIG; }
  const std::string kt = s.substr(b, e - b);
  m->degree = 0; m->gamma = 1; m->coef0 = 0;
  if (kt.find("POLY") != std::string::npos) {
    m->kernel = 1;
    if (!xml_number(s, "degree", kb, ke, &m->degree) || !xml_number(s, "gamma", kb, ke, &m->gamma) || !xml_number(s, "coef0", kb, ke, &m->coef0)) {
      *err = "SVM model: POLY kernel needs degree, gamma and coef0"; return AICP_B200_ERR_CONFIG;
    }
  } else if (kt.find("LINEAR") != std::string::npos) {
    m->kernel = 0;
  } else {
    *err = "SVM model: kernel " + kt + " is not supported (LINEAR, POLY; the reference uses POLY, svm.cpp:10)"; return AICP_B200_ERR_CONFIG;
  }
  double v;
  if (!xml_number(s, "var_count", 0, N, &v)) { *err = "SVM model: no var_count"; return AICP_B200_ERR_CONFIG; }
  m->dim = (int)v;
  if (!xml_number(s, "class_count", 0, N, &v) || (int)v != 2) { *err = "SVM model: exactly two classes are supported"; return AICP_B200_ERR_CONFIG; }
  if (!xml_number(s, "sv_total", 0, N, &v)) { *err = "SVM model: no sv_total"; return AICP_B200_ERR_CONFIG; }
  m->sv_total = (int)v;
  if (m->dim < 1 || m->dim > 8 || m->sv_total < 1 || m->sv_total > 2048) { *err = "SVM model: var_count outside [1,8] or sv_total outside [1,2048]"; return AICP_B200_ERR_CONFIG; }
  size_t sb, se;
  if (!xml_block(s, "support_vectors", 0, N, &sb, &se)) { *err = "SVM model: no support_vectors"; return AICP_B200_ERR_CONFIG; }
  m->sv.clear();
  size_t p = sb;
  while (xml_block(s, "_", p, se, &b, &e)) { parse_numbers(s, b, e, &m->sv); p = e + 4; }
  if ((int)m->sv.size() != m->sv_total * m->dim) { *err = "SVM model: support_vectors do not hold sv_total x var_count numbers"; return AICP_B200_ERR_CONFIG; }
  size_t db, de;
  if (!xml_block(s, "decision_functions", 0, N, &db, &de) || !xml_block(s, "_", db, de, &db, &de)) { *err = "SVM model: no decision function"; return AICP_B200_ERR_CONFIG; }
  if (!xml_number(s, "sv_count", db, de, &v) || !xml_number(s, "rho", db, de, &m->rho)) { *err = "SVM model: decision function without sv_count / rho"; return AICP_B200_ERR_CONFIG; }
  m->sv_count = (int)v;
  m->alpha.clear(); m->index.clear();
  if (!xml_block(s, "alpha", db, de, &b, &e)) { *err = "SVM model: no alpha"; return AICP_B200_ERR_CONFIG; }
  parse_numbers(s, b, e, &m->alpha);
  if (xml_block(s, "index", db, de, &b, &e)) parse_numbers(s, b, e, &m->index);
  else for (int i = 0; i < m->sv_count; ++i) m->index.push_back(i);      // OpenCV: absent index = identity
  if ((int)m->alpha.size() != m->sv_count || (int)m->index.size() != m->sv_count || m->sv_count > m->sv_total) { *err = "SVM model: alpha / index do not match sv_count"; return AICP_B200_ERR_CONFIG; }
  for (int i : m->index) if (i < 0 || i >= m->sv_total) { *err = "SVM model: support-vector index out of range"; return AICP_B200_ERR_CONFIG; }
  return AICP_B200_OK;
}

__global__ void __launch_bounds__(128) k_svm_predict(const float* __restrict__ sv, const double* __restrict__ alpha, const int* __restrict__ index,
                                                     int dim, int sv_total, int sv_count, int kernel, double degree, int ipower, double gamma, double coef0,
                                                     double rho, const float* __restrict__ x, long long n, float* __restrict__ raw,
                                                     double* __restrict__ prob) {
  extern __shared__ unsigned char smem[];
  double* s_alpha = reinterpret_cast<double*>(smem);
  float* s_sv = reinterpret_cast<float*>(s_alpha + sv_count);
  int* s_index = reinterpret_cast<int*>(s_sv + sv_total * dim);
  for (int i = threadIdx.x; i < sv_count; i += blockDim.x) { s_alpha[i] = alpha[i]; s_index[i] = index[i]; }
  for (int i = threadIdx.x; i < sv_total * dim; i += blockDim.x) s_sv[i] = sv[i];
  __syncthreads();
  const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= n) return;
  float xs[8];
  for (int d = 0; d < dim; ++d) xs[d] = x[t * dim + d];
  double sum = -rho;
  for (int i = 0; i < sv_count; ++i) {
    const float* v = s_sv + s_index[i] * dim;
    double s = 0.0;
    for (int d = 0; d < dim; ++d) s = s + (double)__fmul_rn(v[d], xs[d]);
    float kv = (float)(s * gamma + coef0);
    if (kernel == 1) {
      if (ipower > 0) {                      // cv::pow, integer power on float32: iPow's binary exponentiation
        float a = 1.f, b = kv;
        int p = ipower;
        while (p > 1) { if (p & 1) a = __fmul_rn(a, b); b = __fmul_rn(b, b); p >>= 1; }
        kv = __fmul_rn(a, b);
      } else {
        kv = (float)pow((double)kv, degree);
      }
    }
    sum = sum + s_alpha[i] * (double)kv;
  }
  const float out = (float)sum;
  if (raw) raw[t] = out;
  prob[t] = 1.0 - 1.0 / (1.0 + exp(-(double)out));       // svm.cpp:82
}

int svm_parse_summary(const char* path, aicp_b200_svm_summary* out, std::string* err) {
  SvmModel m;
  int rc = svm_parse(path, &m, err);
  if (rc) return rc;
  out->kernel = m.kernel; out->dim = m.dim; out->sv_total = m.sv_total; out->sv_count = m.sv_count;
  out->degree = m.degree; out->gamma = m.gamma; out->coef0 = m.coef0; out->rho = m.rho;
  out->alpha_sum = 0; out->sv_sum = 0;
  for (double a : m.alpha) out->alpha_sum += a;
  for (float v : m.sv) out->sv_sum += (double)v;
  out->index_first = m.index.front(); out->index_last = m.index.back();
  return AICP_B200_OK;
}

int svm_load(Handle* h, const char* path) {
  if (!h->svm) h->svm = new SvmModel();
  SvmModel* m = h->svm;
  m->loaded = false;
  std::string err;
  int rc = svm_parse(path, m, &err);
  if (rc) return fail(h, rc, "%s", err.c_str());
  CUDA_TRY(m->d_sv.reserve(m->sv.size())); CUDA_TRY(m->d_alpha.reserve(m->alpha.size())); CUDA_TRY(m->d_index.reserve(m->index.size()));
  CUDA_TRY(cudaMemcpyAsync(m->d_sv.p, m->sv.data(), sizeof(float) * m->sv.size(), cudaMemcpyHostToDevice, h->stream));
  CUDA_TRY(cudaMemcpyAsync(m->d_alpha.p, m->alpha.data(), sizeof(double) * m->alpha.size(), cudaMemcpyHostToDevice, h->stream));
  CUDA_TRY(cudaMemcpyAsync(m->d_index.p, m->index.data(), sizeof(int) * m->index.size(), cudaMemcpyHostToDevice, h->stream));
  CUDA_TRY(cudaStreamSynchronize(h->stream));
  m->loaded = true;
  return AICP_B200_OK;
}

void svm_release(Handle* h) {
  if (!h->svm) return;
  SvmModel* m = h->svm;
  m->d_sv.release(); m->d_alpha.release(); m->d_index.release(); m->d_x.release(); m->d_prob.release(); m->d_raw.release();
  delete m;
  h->svm = nullptr;
}

int svm_info(Handle* h, int32_t* dim, int32_t* sv_total) {
  if (!h->svm || !h->svm->loaded) return fail(h, AICP_B200_ERR_CONFIG, "no SVM model loaded (aicp_b200_svm_load)");
  if (dim) *dim = h->svm->dim;
  if (sv_total) *sv_total = h->svm->sv_total;
  return AICP_B200_OK;
}

// features: n x dim doubles, row-major (cast to float32 like svm.cpp:72-74); probabilities: n doubles; raw: nullable n floats
int svm_predict(Handle* h, const double* features, int64_t n, int32_t dim, double* probabilities, float* raw) {
  if (!h->svm || !h->svm->loaded) return fail(h, AICP_B200_ERR_CONFIG, "no SVM model loaded (aicp_b200_svm_load)");
  SvmModel* m = h->svm;
  if (dim != m->dim) return fail(h, AICP_B200_ERR_BAD_ARG, "svm_predict: %d features per sample, the model has %d", dim, m->dim);
  if (n == 0) return AICP_B200_OK;
  std::vector<float> xf((size_t)n * dim);
  for (size_t i = 0; i < xf.size(); ++i) xf[i] = (float)features[i];
  CUDA_TRY(m->d_x.reserve(xf.size())); CUDA_TRY(m->d_prob.reserve((size_t)n)); CUDA_TRY(m->d_raw.reserve((size_t)n));
  CUDA_TRY(cudaMemcpyAsync(m->d_x.p, xf.data(), sizeof(float) * xf.size(), cudaMemcpyHostToDevice, h->stream));
  const double rdeg = nearbyint(m->degree);
  const int ipower = (fabs(rdeg - m->degree) < 2.220446049250313e-16 && m->degree >= 1.0 && m->degree <= 64.0) ? (int)rdeg : 0;
  const size_t smem = sizeof(double) * m->sv_count + sizeof(float) * m->sv_total * m->dim + sizeof(int) * m->sv_count;
  if (smem > 48 * 1024) CUDA_TRY(cudaFuncSetAttribute(k_svm_predict, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  k_svm_predict<<<(unsigned)((n + 127) / 128), 128, smem, h->stream>>>(m->d_sv.p, m->d_alpha.p, m->d_index.p, m->dim, m->sv_total, m->sv_count,
                                                                     m->kernel, m->degree, ipower, m->gamma, m->coef0, m->rho, m->d_x.p, (long long)n,
                                                                     m->d_raw.p, m->d_prob.p);
  CUDA_TRY(cudaGetLastError());
  h->launches += 1;
  CUDA_TRY(cudaMemcpyAsync(probabilities, m->d_prob.p, sizeof(double) * (size_t)n, cudaMemcpyDeviceToHost, h->stream));
  if (raw) CUDA_TRY(cudaMemcpyAsync(raw, m->d_raw.p, sizeof(float) * (size_t)n, cudaMemcpyDeviceToHost, h->stream));
  CUDA_TRY(cudaStreamSynchronize(h->stream));
  return AICP_B200_OK;
}

}  // namespace aicp
