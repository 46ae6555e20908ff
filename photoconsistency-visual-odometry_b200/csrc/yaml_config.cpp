// yaml_config.cpp -- reader for the OpenCV-FileStorage YAML-1.0 subset the reference's
// configuration files use (config_files/*.yml): a "%YAML:1.0" header followed by
// `key: scalar` or `key: [a, b, ...]` lines, where keys may contain spaces and parentheses
// ("max_num_iterations (at each level)").  Replaces ReadConfigurationFile
// (CPhotoconsistencyOdometryAnalytic.h:581-607, CPhotoconsistencyOdometryCeres.h:526-576).
// Host-only: usable (and tested) without a GPU.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <map>
#include <sstream>
#include <string>
#include <vector>

#include "phovo_internal.h"

namespace {

std::string trim(const std::string& s) {
  size_t a = s.find_first_not_of(" \t\r\n");
  if (a == std::string::npos) return "";
  size_t b = s.find_last_not_of(" \t\r\n");
  return s.substr(a, b - a + 1);
}

bool parse_number(const std::string& tok, double* out) {
  std::string t = trim(tok);
  if (t.empty()) return false;
  char* end = nullptr;
  double v = std::strtod(t.c_str(), &end);
  if (end == t.c_str()) return false;
  while (*end == ' ' || *end == '\t') ++end;
  if (*end != '\0') return false;
  *out = v;
  return true;
}

typedef std::map<std::string, std::vector<double> > KeyValues;

bool read_file(const char* path, KeyValues* kv, std::string* err) {
  std::ifstream f(path);
  if (!f.is_open()) { *err = std::string("cannot open configuration file '") + path + "'"; return false; }
  std::string line, pending_key, pending_val;
  bool in_seq = false;
  int lineno = 0;
  while (std::getline(f, line)) {
    ++lineno;
    size_t hash = line.find('#');
    if (hash != std::string::npos) line = line.substr(0, hash);
    std::string t = trim(line);
    if (t.empty() || t[0] == '%' || t == "---" || t == "...") continue;
    std::string key, val;
    if (in_seq) {
      pending_val += " " + t;
      if (t.find(']') == std::string::npos) continue;
      key = pending_key; val = pending_val; in_seq = false;
    } else {
      // the key ends at the LAST ':' that is followed by a space or end of line and precedes any '['
      size_t lim = t.find('[');
      size_t colon = std::string::npos;
      for (size_t i = 0; i < t.size() && i < lim; ++i)
        if (t[i] == ':' && (i + 1 == t.size() || t[i + 1] == ' ' || t[i + 1] == '\t' || t[i + 1] == '[')) colon = i;
      if (colon == std::string::npos) {
        std::ostringstream o; o << path << ":" << lineno << ": expected 'key: value'";
        *err = o.str(); return false;
      }
      key = trim(t.substr(0, colon));
      val = trim(t.substr(colon + 1));
      if (!val.empty() && val[0] == '[' && val.find(']') == std::string::npos) {
        in_seq = true; pending_key = key; pending_val = val; continue;
      }
    }
    if (key.size() >= 2 && (key[0] == '"' || key[0] == '\'')) key = key.substr(1, key.size() - 2);
    std::vector<double> nums;
    if (!val.empty() && val[0] == '[') {
      size_t close = val.find(']');
      std::string body = val.substr(1, close - 1);
      std::stringstream ss(body);
      std::string tok;
      while (std::getline(ss, tok, ',')) {
        if (trim(tok).empty()) continue;
        double v;
        if (!parse_number(tok, &v)) {
          std::ostringstream o; o << path << ":" << lineno << ": '" << trim(tok) << "' is not a number";
          *err = o.str(); return false;
        }
        nums.push_back(v);
      }
    } else {
      double v;
      if (parse_number(val, &v)) nums.push_back(v);
      // non-numeric scalars (strings) are ignored: the reference reads none
    }
    (*kv)[key] = nums;
  }
  if (in_seq) { *err = std::string(path) + ": unterminated '[' sequence"; return false; }
  return true;
}

// Missing trailing entries repeat the last given one (reference: reads past the std::vector,
// e.g. config_5_level_optimization_ceres.yml:11 has 4 radii for 5 levels).  Extra entries are ignored.
template <class T>
void fill_levels(const KeyValues& kv, const char* key, T* dst) {
  KeyValues::const_iterator it = kv.find(key);
  if (it == kv.end() || it->second.empty()) return;
  const std::vector<double>& v = it->second;
  for (int l = 0; l < PHOVO_MAX_LEVELS; ++l) dst[l] = (T)(l < (int)v.size() ? v[l] : v.back());
}

void read_scalar(const KeyValues& kv, const char* key, int32_t* dst) {
  KeyValues::const_iterator it = kv.find(key);
  if (it != kv.end() && !it->second.empty()) *dst = (int32_t)it->second[0];
}

}  // namespace

int phovo_internal_parse_yaml(const char* path, phovo_config* cfg, std::string* err) {
  KeyValues kv;
  if (!read_file(path, &kv, err)) return PHOVO_E_CONFIG;
  KeyValues::const_iterator it = kv.find("numOptimizationLevels");
  if (it == kv.end() || it->second.empty()) {
    *err = std::string(path) + ": missing key 'numOptimizationLevels'";
    return PHOVO_E_CONFIG;
  }
  int n = (int)it->second[0];
  if (n < 1 || n > PHOVO_MAX_LEVELS) {
    std::ostringstream o; o << path << ": numOptimizationLevels=" << n << " outside [1," << PHOVO_MAX_LEVELS << "]";
    *err = o.str(); return PHOVO_E_CONFIG;
  }
  cfg->num_levels = n;
  fill_levels(kv, "blurFilterSize (at each level)", cfg->blur_filter_size);
  fill_levels(kv, "imageGradientsScalingFactor (at each level)", cfg->grad_scale);
  fill_levels(kv, "lambda_optimization_step (at each level)", cfg->lambda_step);
  fill_levels(kv, "max_num_iterations (at each level)", cfg->max_num_iterations);
  fill_levels(kv, "min_gradient_norm (at each level)", cfg->min_gradient_norm);
  fill_levels(kv, "function_tolerance (at each level)", cfg->function_tolerance);
  fill_levels(kv, "gradient_tolerance (at each level)", cfg->gradient_tolerance);
  fill_levels(kv, "parameter_tolerance (at each level)", cfg->parameter_tolerance);
  fill_levels(kv, "initial_trust_region_radius (at each level)", cfg->initial_trust_region_radius);
  fill_levels(kv, "max_trust_region_radius (at each level)", cfg->max_trust_region_radius);
  fill_levels(kv, "min_trust_region_radius (at each level)", cfg->min_trust_region_radius);
  fill_levels(kv, "min_relative_decrease (at each level)", cfg->min_relative_decrease);
  read_scalar(kv, "num_threads", &cfg->num_threads);
  read_scalar(kv, "num_linear_solver_threads", &cfg->num_linear_solver_threads);
  read_scalar(kv, "minimizer_progress_to_stdout", &cfg->minimizer_progress_to_stdout);
  read_scalar(kv, "visualizeIterations", &cfg->visualize_iterations);
  // levels beyond num_levels never run
  for (int l = n; l < PHOVO_MAX_LEVELS; ++l) cfg->max_num_iterations[l] = 0;
  for (int l = 0; l < n; ++l) {
    int k = cfg->blur_filter_size[l];
    if (k < 0 || (k > 0 && (k % 2) == 0)) {
      std::ostringstream o; o << path << ": blurFilterSize[" << l << "]=" << k << " must be 0 or odd (cv::GaussianBlur)";
      *err = o.str(); return PHOVO_E_CONFIG;
    }
  }
  return PHOVO_OK;
}

void phovo_internal_default_config(phovo_config* cfg) {
  // CPhotoconsistencyOdometryAnalytic.h:430-443
  std::memset(cfg, 0, sizeof(*cfg));
  cfg->mode = PHOVO_MODE_ANALYTIC_REF;
  cfg->num_levels = 5;
  for (int l = 0; l < PHOVO_MAX_LEVELS; ++l) {
    cfg->blur_filter_size[l] = 0;
    cfg->grad_scale[l] = 0.0625;
    cfg->lambda_step[l] = 1.;
    cfg->max_num_iterations[l] = 0;
    cfg->min_gradient_norm[l] = 300.;
    // Ceres-mode: ceres::Solver::Options defaults
    cfg->function_tolerance[l] = 1e-6;
    cfg->gradient_tolerance[l] = 1e-10;
    cfg->parameter_tolerance[l] = 1e-8;
    cfg->initial_trust_region_radius[l] = 1e4;
    cfg->max_trust_region_radius[l] = 1e16;
    cfg->min_trust_region_radius[l] = 1e-32;
    cfg->min_relative_decrease[l] = 1e-3;
  }
  cfg->max_num_iterations[2] = 5;
  cfg->max_num_iterations[3] = 20;
  cfg->max_num_iterations[4] = 50;
  cfg->min_depth = 0.3;
  cfg->max_depth = 5.0;
  cfg->num_threads = 1;
  cfg->num_linear_solver_threads = 1;
}
