// config_yaml.cpp -- reader for the libpointmatcher ICP-chain YAML files that AICP ships and rewrites.
//
// replaces PointmatcherRegistration::applyConfig() -> icp_.loadFromYaml(ifs)
//          (aicp_core/src/registration/pointmatcher_registration.cpp:48-68) for the module chain the B200 path
// implements.  The file must be accepted VERBATIM because App::computeRegistration regenerates it before every call
// (aicp_core/src/registration/app.cpp:204-205, aicp_core/src/utils/fileIO.cpp:179-214): see
// aicp_core/config/icp/icp_autotuned.yaml:9-58 and icp_autotuned_default.yaml.
//
// Only the YAML subset those files use is understood: top-level sections, "- Module:" list items or a bare
// "Module" / "Module:" scalar, and "param: value" maps one level below; '#' comments.  Anything that would change
// the numerical result and is not implemented on the GPU (another matcher, outlier filter, minimiser or data-points
// filter, maxDist, force2D, knn != 1 ...) is rejected with a message instead of being silently ignored.
// Module defaults are libpointmatcher's ([UPSTREAM] parameter docs: SurfaceNormal knn 5, TrimmedDist ratio 0.85,
// Counter maxIterationCount 40, Differential 0.001 / 0.001 / 3).
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <map>
#include <sstream>
#include <string>
#include <vector>

#include "../../include/aicp_b200.h"

namespace aicp {

void default_icp_config(aicp_b200_icp_config* cfg) {
  // values of aicp_core/config/icp/icp_autotuned_default.yaml (the template AICP starts every run from)
  cfg->knn_normals = 20;
  cfg->reading_normals = 0;
  cfg->ratio = 0.70f;
  cfg->max_iterations = 20;
  cfg->min_diff_rot = 0.001f;
  cfg->min_diff_trans = 0.01f;
  cfg->smooth_length = 4;
  cfg->matcher_epsilon = 0.f;
}

namespace {

struct Module {
  std::string name;
  std::map<std::string, std::string> params;
  int line;
};
struct Section {
  std::string name;
  std::vector<Module> modules;
};

std::string trim(const std::string& s) {
  size_t a = s.find_first_not_of(" \t\r\n");
  if (a == std::string::npos) return "";
  size_t b = s.find_last_not_of(" \t\r\n");
  return s.substr(a, b - a + 1);
}

std::string strip_comment(const std::string& s) {
  bool in_s = false, in_d = false;
  for (size_t i = 0; i < s.size(); ++i) {
    char c = s[i];
    if (c == '\'' && !in_d) in_s = !in_s;
    else if (c == '"' && !in_s) in_d = !in_d;
    else if (c == '#' && !in_s && !in_d && (i == 0 || s[i - 1] == ' ' || s[i - 1] == '\t')) return s.substr(0, i);
  }
  return s;
}

std::string unquote(const std::string& s) {
  if (s.size() >= 2 && ((s.front() == '"' && s.back() == '"') || (s.front() == '\'' && s.back() == '\''))) return s.substr(1, s.size() - 2);
  return s;
}

bool parse_sections(std::istream& in, std::vector<Section>* out, std::string* err) {
  std::string raw;
  int line_no = 0;
  int module_indent = -1;
  while (std::getline(in, raw)) {
    ++line_no;
    std::string line = strip_comment(raw);
    if (trim(line).empty()) continue;
    if (line.find('\t') != std::string::npos) {
      std::string lead = line.substr(0, line.find_first_not_of(" \t"));
      if (lead.find('\t') != std::string::npos) { *err = "line " + std::to_string(line_no) + ": tab indentation"; return false; }
    }
    int indent = (int)line.find_first_not_of(' ');
    std::string body = trim(line);
    if (indent == 0) {
      size_t colon = body.find(':');
      if (colon == std::string::npos) { *err = "line " + std::to_string(line_no) + ": expected 'section:'"; return false; }
      Section s;
      s.name = trim(body.substr(0, colon));
      out->push_back(s);
      module_indent = -1;
      std::string rest = trim(body.substr(colon + 1));
      if (!rest.empty()) {            // "logger: NullLogger" on one line
        Module m; m.name = unquote(rest); m.line = line_no;
        out->back().modules.push_back(m);
      }
      continue;
    }
    if (out->empty()) { *err = "line " + std::to_string(line_no) + ": content before the first section"; return false; }
    Section& sec = out->back();
    bool is_item = body.size() >= 2 && body[0] == '-' && body[1] == ' ';
    if (is_item) { body = trim(body.substr(2)); indent += 2; }
    if (module_indent < 0 || is_item || indent <= module_indent) {
      // a module line: "Name:" / "Name" / "- Name:"
      if (!is_item && module_indent >= 0 && indent > module_indent) { /* unreachable */ }
      size_t colon = body.find(':');
      Module m; m.line = line_no;
      if (colon == std::string::npos) m.name = unquote(body);
      else {
        m.name = trim(body.substr(0, colon));
        std::string rest = trim(body.substr(colon + 1));
        if (!rest.empty()) { *err = "line " + std::to_string(line_no) + ": unexpected value after module name '" + m.name + "'"; return false; }
      }
      sec.modules.push_back(m);
      module_indent = indent;
      continue;
    }
    // a parameter of the current module
    size_t colon = body.find(':');
    if (colon == std::string::npos || sec.modules.empty()) { *err = "line " + std::to_string(line_no) + ": expected 'param: value'"; return false; }
    sec.modules.back().params[trim(body.substr(0, colon))] = unquote(trim(body.substr(colon + 1)));
  }
  return true;
}

bool to_float(const std::string& s, float* v) {
  if (s.empty()) return false;
  if (s == "inf" || s == ".inf" || s == "+inf") { *v = HUGE_VALF; return true; }
  char* end = nullptr;
  float f = strtof(s.c_str(), &end);
  if (end == s.c_str() || *end != 0) return false;
  *v = f;
  return true;
}
bool to_int(const std::string& s, int* v) {
  if (s.empty()) return false;
  char* end = nullptr;
  long l = strtol(s.c_str(), &end, 10);
  if (end == s.c_str() || *end != 0) return false;
  *v = (int)l;
  return true;
}

#define CFG_FAIL(msg) do { *err = (msg); return AICP_B200_ERR_CONFIG; } while (0)

int surface_normal_filter(const Module& m, const char* where, int* knn, std::string* err) {
  *knn = 5;
  for (const auto& kv : m.params) {
    if (kv.first == "knn") { if (!to_int(kv.second, knn) || *knn < 3) CFG_FAIL(std::string(where) + ": bad knn '" + kv.second + "'"); }
    else if (kv.first == "epsilon") { float e; if (!to_float(kv.second, &e)) CFG_FAIL(std::string(where) + ": bad epsilon"); /* search is exact */ }
    else if (kv.first == "keepNormals") { if (kv.second != "1") CFG_FAIL(std::string(where) + ": keepNormals must be 1 for PointToPlaneErrorMinimizer"); }
    else if (kv.first == "keepDensities" || kv.first == "keepEigenValues" || kv.first == "keepEigenVectors" || kv.first == "keepMatchedIds") { /* extra descriptors never read by the chain */ }
    else CFG_FAIL(std::string(where) + ": unsupported SurfaceNormalDataPointsFilter parameter '" + kv.first + "'");
  }
  return AICP_B200_OK;
}

}  // namespace

int parse_icp_yaml(const char* path, aicp_b200_icp_config* cfg, std::string* err) {
  default_icp_config(cfg);
  if (!path || !*path) return AICP_B200_OK;
  std::ifstream ifs(path);
  if (!ifs.good()) CFG_FAIL(std::string("[Pointmatcher] Cannot open config file ") + path);   // pointmatcher_registration.cpp:60-64
  std::vector<Section> sections;
  if (!parse_sections(ifs, &sections, err)) { *err = std::string(path) + ": " + *err; return AICP_B200_ERR_CONFIG; }

  bool have_ref_normals = false, have_trim = false, have_counter = false, have_diff = false, have_minimizer = false, have_matcher = false;
  cfg->reading_normals = 0;
  for (const Section& s : sections) {
    if (s.name == "readingDataPointsFilters") {
      for (const Module& m : s.modules) {
        if (m.name != "SurfaceNormalDataPointsFilter")
          CFG_FAIL("readingDataPointsFilters: '" + m.name + "' is not implemented on the B200 path (only SurfaceNormalDataPointsFilter)");
        int knn; int rc = surface_normal_filter(m, "readingDataPointsFilters", &knn, err);
        if (rc) return rc;
        cfg->reading_normals = 0;   // parsed and validated; PointToPlane never reads reading normals, so the GPU path skips them
      }
    } else if (s.name == "referenceDataPointsFilters") {
      for (const Module& m : s.modules) {
        if (m.name != "SurfaceNormalDataPointsFilter")
          CFG_FAIL("referenceDataPointsFilters: '" + m.name + "' is not implemented on the B200 path (only SurfaceNormalDataPointsFilter)");
        if (have_ref_normals) CFG_FAIL("referenceDataPointsFilters: more than one SurfaceNormalDataPointsFilter");
        int rc = surface_normal_filter(m, "referenceDataPointsFilters", &cfg->knn_normals, err);
        if (rc) return rc;
        have_ref_normals = true;
      }
    } else if (s.name == "matcher") {
      if (s.modules.size() != 1 || s.modules[0].name != "KDTreeMatcher") CFG_FAIL("matcher: only KDTreeMatcher is implemented");
      have_matcher = true;
      for (const auto& kv : s.modules[0].params) {
        if (kv.first == "knn") { int k; if (!to_int(kv.second, &k) || k != 1) CFG_FAIL("KDTreeMatcher: only knn: 1 is implemented"); }
        else if (kv.first == "epsilon") { if (!to_float(kv.second, &cfg->matcher_epsilon) || cfg->matcher_epsilon < 0) CFG_FAIL("KDTreeMatcher: bad epsilon"); }
        else if (kv.first == "maxDist") { float d; if (!to_float(kv.second, &d) || d != HUGE_VALF) CFG_FAIL("KDTreeMatcher: finite maxDist is not implemented"); }
        else if (kv.first == "searchType") { /* all libnabo search types return the same neighbours at epsilon 0 */ }
        else CFG_FAIL("KDTreeMatcher: unsupported parameter '" + kv.first + "'");
      }
    } else if (s.name == "outlierFilters") {
      for (const Module& m : s.modules) {
        if (m.name != "TrimmedDistOutlierFilter") CFG_FAIL("outlierFilters: '" + m.name + "' is not implemented (only TrimmedDistOutlierFilter)");
        if (have_trim) CFG_FAIL("outlierFilters: more than one TrimmedDistOutlierFilter");
        have_trim = true;
        cfg->ratio = 0.85f;
        for (const auto& kv : m.params) {
          if (kv.first == "ratio") { if (!to_float(kv.second, &cfg->ratio) || !(cfg->ratio > 0.f) || cfg->ratio > 1.f) CFG_FAIL("TrimmedDistOutlierFilter: ratio '" + kv.second + "' outside (0,1]"); }
          else CFG_FAIL("TrimmedDistOutlierFilter: unsupported parameter '" + kv.first + "'");
        }
      }
    } else if (s.name == "errorMinimizer") {
      if (s.modules.size() != 1 || s.modules[0].name != "PointToPlaneErrorMinimizer") CFG_FAIL("errorMinimizer: only PointToPlaneErrorMinimizer is implemented");
      have_minimizer = true;
      for (const auto& kv : s.modules[0].params) {
        if (kv.first == "force2D") { if (kv.second != "0") CFG_FAIL("PointToPlaneErrorMinimizer: force2D is not implemented"); }
        else CFG_FAIL("PointToPlaneErrorMinimizer: unsupported parameter '" + kv.first + "'");
      }
    } else if (s.name == "transformationCheckers") {
      for (const Module& m : s.modules) {
        if (m.name == "CounterTransformationChecker") {
          have_counter = true;
          cfg->max_iterations = 40;
          for (const auto& kv : m.params) {
            if (kv.first == "maxIterationCount") { if (!to_int(kv.second, &cfg->max_iterations) || cfg->max_iterations < 1) CFG_FAIL("CounterTransformationChecker: bad maxIterationCount"); }
            else CFG_FAIL("CounterTransformationChecker: unsupported parameter '" + kv.first + "'");
          }
        } else if (m.name == "DifferentialTransformationChecker") {
          have_diff = true;
          cfg->min_diff_rot = 0.001f; cfg->min_diff_trans = 0.001f; cfg->smooth_length = 3;
          for (const auto& kv : m.params) {
            if (kv.first == "minDiffRotErr") { if (!to_float(kv.second, &cfg->min_diff_rot)) CFG_FAIL("DifferentialTransformationChecker: bad minDiffRotErr"); }
            else if (kv.first == "minDiffTransErr") { if (!to_float(kv.second, &cfg->min_diff_trans)) CFG_FAIL("DifferentialTransformationChecker: bad minDiffTransErr"); }
            else if (kv.first == "smoothLength") { if (!to_int(kv.second, &cfg->smooth_length) || cfg->smooth_length < 1) CFG_FAIL("DifferentialTransformationChecker: bad smoothLength"); }
            else CFG_FAIL("DifferentialTransformationChecker: unsupported parameter '" + kv.first + "'");
          }
        } else CFG_FAIL("transformationCheckers: '" + m.name + "' is not implemented");
      }
    } else if (s.name == "inspector" || s.name == "logger") {
      /* no effect on the result */
    } else {
      CFG_FAIL("unknown section '" + s.name + "'");
    }
  }
  if (!have_matcher) { /* libpointmatcher default matcher is KDTreeMatcher */ }
  if (!have_minimizer) CFG_FAIL("errorMinimizer section missing (libpointmatcher would default to PointToPoint, not implemented)");
  if (!have_ref_normals) CFG_FAIL("PointToPlaneErrorMinimizer needs a SurfaceNormalDataPointsFilter in referenceDataPointsFilters");
  if (!have_trim) CFG_FAIL("outlierFilters must hold a TrimmedDistOutlierFilter");
  if (!have_counter) CFG_FAIL("transformationCheckers must hold a CounterTransformationChecker (the device loop needs a bound)");
  if (cfg->max_iterations > AICP_B200_MAX_ITERS) CFG_FAIL("maxIterationCount above " + std::to_string(AICP_B200_MAX_ITERS));
  if (!have_diff) { cfg->min_diff_rot = -1.f; cfg->min_diff_trans = -1.f; cfg->smooth_length = 1; }   // never satisfied
  return AICP_B200_OK;
}

}  // namespace aicp
