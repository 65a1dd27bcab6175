// map_eval_b200.hpp — C++ host side of the drop-in: MapEval's config / CLI / metric-output surface re-hosted on the
// C-ABI of libmapeval_b200.so.  Names, argument meaning and error behaviour follow the reference
// (map_eval/src/map_eval.h, map_eval.cpp, map_eval_main.cpp) for the path this repository replaces:
//
//   reference                                           here
//   ------------------------------------------------    ---------------------------------------------------------
//   struct Param                       map_eval.h:60     struct Param (same fields and defaults)
//   loadParametersFromYAML   map_eval_main.cpp:120-208    loadParametersFromYAML (same keys, same required/optional split)
//   MapEval::MapEval(Param&)        map_eval.h:123-189    MapEvalB200::MapEvalB200 (results folder + header lines)
//   MapEval::process()               map_eval.cpp:4-102    MapEvalB200::process (same sequence, same return codes)
//   computeMME(cloud, gt)           map_eval.cpp:149-189   computeMME            -> me_eval_mme
//   calculateMetricsWithInitialMatrix    :1204-1260       calculateMetricsWithInitialMatrix -> me_transform + me_eval_nn
//   calculateVMD                          :240-390        calculateVMD          -> me_eval_awd (+ the two text files)
//   saveMmeResults / saveRegistrationResults :392-482     same line formats in map_results.txt
//
//   VoxelDownSample                  map_eval.cpp:38-39    me_voxel_downsample (§8f N1)
//   performRegistration / performICPRegistration :191-237, 1366-1394   performRegistration -> me_icp (§8f N2: point-to-point,
//                                                         point-to-plane and generalized ICP) + calculateMetrics(reg)
//   rendered clouds                  :485-499, 568-736    render.hpp (§8f N3): entropy maps, raw / inlier NN-distance maps
// Limits of this driver: point-to-plane ICP (registration_methods: 1) needs normals in the ground-truth FILE, as in the
// reference (Open3D raises without them), and no down-sampling (VoxelDownSample's normal averaging is not built);
// validate() rejects such a configuration before anything is written.
#pragma once
#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <ctime>
#include <filesystem>
#include <fstream>
#include <iomanip>
#include <iostream>
#include <sstream>
#include <stdexcept>
#include <string>
#include <vector>

#include "../../include/mapeval_b200.h"
#include "cloud_io.hpp"
#include "gpu_group.hpp"
#include "render.hpp"
#include "yaml_lite.hpp"

namespace fs = std::filesystem;

struct Param {   // map_eval.h:60-116
  std::string evaluation_map_pcd_path_ = "/data/map_evaluation/canteen/";
  std::string map_gt_path_ = "/data/map_evaluation/canteen/merged_scan.pcd";
  std::string result_path_ = "/home/hts/workspace/dataset/eva_results/";
  std::string pcd_file_name_ = "map.pcd";
  std::string name_;
  int evaluation_method_ = 2;
  double voxel_size_ = 1.0;
  double icp_max_distance_ = 2.5;
  double nn_radius_ = 0.2;
  bool save_immediate_result_ = false;
  bool evaluate_mme_ = true;
  bool evaluate_gt_mme_ = true;
  bool evaluate_using_initial_ = true;
  bool evaluate_noised_gt_ = false;
  bool use_visualization = false;
  bool enable_debug = false;
  bool use_tbb_mme = true;
  double trunc_dist_[5] = {0, 0, 0, 0, 0};   // the reference leaves it uninitialised when accuracy_level is absent
  bool trunc_dist_set_ = false;
  double initial_matrix_[16] = {1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1};
  double noise_std_dev_ = 0.1;
  double vmd_voxel_size_ = 3.0;
  double downsample_size = 0.01;
  // additions of this implementation (absent keys keep the reference behaviour)
  int gpu_device_ = 0;
  int n_gpus_ = 1;                      // GPUs gpu_device_ .. gpu_device_ + n_gpus_ - 1 share the sweeps (NCCL all-reduce)
  bool geometric_gt_pairing_ = false;   // false = reproduce map_eval.cpp:1233/:1241 verbatim
};

// map_eval_main.cpp:120-208
inline Param loadParametersFromYAML(const std::string &yaml_file_path) {
  try {
    yaml_lite::Document config = yaml_lite::Document::LoadFile(yaml_file_path);
    Param param;
    param.evaluation_method_ = config["registration_methods"].as<int>();
    param.icp_max_distance_ = config["icp_max_distance"].as<double>();
    if (config["accuracy_level"] && config["accuracy_level"].size() >= 5) {
      for (int i = 0; i < 5; ++i) param.trunc_dist_[i] = config["accuracy_level"][i].as<double>();
      param.trunc_dist_set_ = true;
    }
    if (config["initial_matrix"] && config["initial_matrix"].size() >= 4) {
      for (int i = 0; i < 4; ++i)
        for (int j = 0; j < 4; ++j) param.initial_matrix_[i * 4 + j] = config["initial_matrix"][i][j].as<double>();
    }
    param.save_immediate_result_ = config["save_immediate_result"].as<bool>();
    param.evaluate_mme_ = config["evaluate_mme"].as<bool>();
    param.evaluate_gt_mme_ = config["evaluate_gt_mme"].as<bool>();
    param.evaluate_using_initial_ = config["evaluate_using_initial"].as<bool>();
    param.nn_radius_ = config["nn_radius"].as<double>();
    param.vmd_voxel_size_ = config["vmd_voxel_size"].as<double>();
    param.downsample_size = config["downsample_size"].as<double>();
    param.evaluation_map_pcd_path_ = config["estimate_map_path"].as<std::string>();
    param.map_gt_path_ = config["gt_map_path"].as<std::string>();
    param.name_ = config["scene_name"].as<std::string>();
    if (!param.evaluation_map_pcd_path_.empty() && param.evaluation_map_pcd_path_.back() != '/')
      param.evaluation_map_pcd_path_ += '/';
    param.result_path_ = param.evaluation_map_pcd_path_ + "map_results/";
    if (config["pcd_file_name"]) param.pcd_file_name_ = config["pcd_file_name"].as<std::string>();
    if (config["evaluate_noised_gt"]) param.evaluate_noised_gt_ = config["evaluate_noised_gt"].as<bool>();
    if (config["noise_std_dev"]) param.noise_std_dev_ = config["noise_std_dev"].as<double>();
    if (config["voxel_size"]) param.voxel_size_ = config["voxel_size"].as<double>();
    if (config["use_visualization"]) param.use_visualization = config["use_visualization"].as<bool>();
    param.enable_debug = config["enable_debug"].as<bool>();
    if (config["use_tbb_mme"]) param.use_tbb_mme = config["use_tbb_mme"].as<bool>();
    if (config["gpu_device"]) param.gpu_device_ = config["gpu_device"].as<int>();
    if (config["n_gpus"]) param.n_gpus_ = config["n_gpus"].as<int>();
    if (config["geometric_gt_pairing"]) param.geometric_gt_pairing_ = config["geometric_gt_pairing"].as<bool>();
    return param;
  } catch (const std::exception &e) {
    std::cerr << "Error loading parameters: " << e.what() << std::endl;
    throw std::runtime_error("Failed to parse YAML file: " + yaml_file_path);
  }
}

inline void dumpParam(const Param &p, std::ostream &os) {
  os << std::setprecision(17);
  os << "registration_methods=" << p.evaluation_method_ << "\nicp_max_distance=" << p.icp_max_distance_ << "\naccuracy_level=";
  for (int i = 0; i < 5; ++i) os << (i ? "," : "") << p.trunc_dist_[i];
  os << "\naccuracy_level_set=" << p.trunc_dist_set_ << "\ninitial_matrix=";
  for (int i = 0; i < 16; ++i) os << (i ? "," : "") << p.initial_matrix_[i];
  os << "\nsave_immediate_result=" << p.save_immediate_result_ << "\nevaluate_mme=" << p.evaluate_mme_
     << "\nevaluate_gt_mme=" << p.evaluate_gt_mme_ << "\nevaluate_using_initial=" << p.evaluate_using_initial_
     << "\nnn_radius=" << p.nn_radius_ << "\nvmd_voxel_size=" << p.vmd_voxel_size_ << "\ndownsample_size=" << p.downsample_size
     << "\nestimate_map_path=" << p.evaluation_map_pcd_path_ << "\ngt_map_path=" << p.map_gt_path_ << "\nscene_name=" << p.name_
     << "\nresult_path=" << p.result_path_ << "\npcd_file_name=" << p.pcd_file_name_ << "\nevaluate_noised_gt=" << p.evaluate_noised_gt_
     << "\nnoise_std_dev=" << p.noise_std_dev_ << "\nvoxel_size=" << p.voxel_size_ << "\nuse_visualization=" << p.use_visualization
     << "\nenable_debug=" << p.enable_debug << "\nuse_tbb_mme=" << p.use_tbb_mme << "\n";
}

// Eigen's default IOFormat for `vec.transpose()`: coefficients right-aligned to a common width, one space apart
// Eigen's operator<< for a 4x4 matrix: rows on separate lines, every coefficient right-aligned to the widest one
inline std::string eigenMatrix4(const double *T, int precision) {
  std::string cell[16];
  size_t w = 0;
  for (int i = 0; i < 16; ++i) {
    std::ostringstream os;
    os << std::fixed << std::setprecision(precision) << T[i];
    cell[i] = os.str();
    w = std::max(w, cell[i].size());
  }
  std::string out;
  for (int r = 0; r < 4; ++r) {
    if (r) out += "\n";
    for (int c = 0; c < 4; ++c) {
      if (c) out += " ";
      out += std::string(w - cell[r * 4 + c].size(), ' ') + cell[r * 4 + c];
    }
  }
  return out;
}

inline std::string eigenRow(const double *v, int n, int precision) {
  std::vector<std::string> s(n);
  size_t w = 0;
  for (int i = 0; i < n; ++i) {
    std::ostringstream os;
    os << std::fixed << std::setprecision(precision) << v[i];
    s[i] = os.str();
    w = std::max(w, s[i].size());
  }
  std::string out;
  for (int i = 0; i < n; ++i) {
    if (i) out += " ";
    out += std::string(w - s[i].size(), ' ') + s[i];
  }
  return out;
}

class TicToc {   // include/tic_toc.h:10-25 (milliseconds, system_clock)
 public:
  TicToc() { tic(); }
  void tic() { start = std::chrono::system_clock::now(); }
  double toc() { return std::chrono::duration<double>(std::chrono::system_clock::now() - start).count() * 1000; }
 private:
  std::chrono::time_point<std::chrono::system_clock> start;
};

class MapEvalB200 {
 public:
  // Configurations this driver cannot run, reported BEFORE map_results.txt is opened (no partial record is left behind).
  static bool validate(const Param &p, std::string *why) {
    if (p.evaluate_using_initial_) return true;
    if (p.evaluation_method_ < 0 || p.evaluation_method_ > 2) {
      *why = "Invalid registration type specified (registration_methods: " + std::to_string(p.evaluation_method_) +
             "; 0 = point-to-point, 1 = point-to-plane, 2 = generalized ICP, map_eval.cpp:1368-1390)";
      return false;
    }
    if (p.evaluation_method_ == 1 && p.downsample_size > 0) {
      *why = "registration_methods: 1 (point-to-plane ICP) takes the target normals from the ground-truth file; with "
             "downsample_size > 0 they would have to be averaged per voxel (Open3D VoxelDownSample), which this driver does "
             "not do: set downsample_size: 0, or use registration_methods: 0 / 2";
      return false;
    }
    return true;
  }

  explicit MapEvalB200(Param &param) : param_(param) {   // map_eval.h:123-189
    t1 = t2 = t3 = t4 = t5 = t6 = t7 = 0.0;
    if (param_.pcd_file_name_ == "merged_maps_all_trans.pcd") subfolder = "merged_maps_all_results/";
    else if (param_.pcd_file_name_ == "merged_maps_s0_trans.pcd") subfolder = "merged_maps_s0_results/";
    else if (param_.pcd_file_name_ == "merged_maps_s1_trans.pcd") subfolder = "merged_maps_s1_results/";
    else if (param_.pcd_file_name_ == "global_pcd_lidar.pcd") subfolder = "map_results/";
    else {
      subfolder = "map_results/";
      std::cerr << "ERROR: Invalid PCD file name: " << param_.pcd_file_name_ << std::endl;   // as the reference (map.pcd lands here too)
    }
    results_subfolder = param_.evaluation_map_pcd_path_ + subfolder;
    std::cout << "INFO: Saving results to: " << results_subfolder << std::endl;
    if (!fs::exists(results_subfolder)) fs::create_directory(results_subfolder);
    results_file_path = results_subfolder + "map_results.txt";
    file_result.open(results_file_path, std::ios::app);
    if (!file_result.is_open()) std::cerr << "ERROR: Failed to open results file at " << results_file_path << std::endl;
    auto now = std::chrono::system_clock::now();
    std::time_t now_c = std::chrono::system_clock::to_time_t(now);
    std::stringstream time_stream;
    time_stream << std::put_time(std::localtime(&now_c), "%Y-%m-%d %X");
    file_result << param_.name_ << " ===================== " << time_stream.str() << " ===================== " << std::endl;
    file_result << "Ground Truth Path: " << param_.map_gt_path_ << std::endl;
    file_result << "Evaluation Map Path: " << param_.evaluation_map_pcd_path_ + param_.pcd_file_name_ << std::endl;
    std::cout << "INFO: Evaluation details saved to " << results_file_path << std::endl;
  }
  ~MapEvalB200() {
    file_result.close();
    gpus_.destroy();
  }

  int process();                                   // map_eval.cpp:4-102
  int computeMME();                                // map_eval.cpp:149-189
  int calculateMetricsWithInitialMatrix();         // map_eval.cpp:1204-1260
  int performRegistration();                       // map_eval.cpp:191-237, 1366-1394 (all three registration methods), 1147-1202
  int calculateVMD();                              // map_eval.cpp:240-390
  void saveMmeResults();                           // map_eval.cpp:392-421 (result line + the rendered entropy clouds)
  void saveRegistrationResults();                  // map_eval.cpp:424-482 (text lines only)

  Param param_;
  // results, named as the reference's members (map_eval.h:319-361)
  me_nn_result nn_{};
  me_mme_result mme_est_res_{}, mme_gt_res_{};
  me_awd_result awd_{};
  double mme_est = 0.0, mme_gt = 0.0, max_abs_entropy = 0.0, min_abs_entropy = 0.0;
  double vmd = 0.0, scs_overall = 0.0, full_chamfer_dist = 0.0;

 private:
  int fail(const char *what) {
    std::cerr << "ERROR: " << what << ": "
              << (!gpus_.error().empty() ? gpus_.error().c_str() : (ctx_ ? me_last_error(ctx_) : me_last_error(nullptr))) << std::endl;
    return -1;
  }
  render::ColoredCloud map_3d_entropy_, gt_3d_entropy_, map_3d_render_raw_, map_3d_render_inlier_;   // map_eval.h:341-347
  int renderEntropy(int which, int64_t n, render::ColoredCloud *out);
  int renderDistances(bool cutoff_strict);
  GpuGroup gpus_;            // one context per GPU; rank r evaluates the r-th shard of every sweep's query range
  me_ctx *ctx_ = nullptr;    // = gpus_.ctx(0): the voxel stage and the single-GPU calls
  std::vector<double> map_3d_, gt_3d_;   // N x 3 fp64, the layout of open3d PointCloud::points_ (as loaded)
  std::vector<double> gt_normals_;       // PointCloud::normals_ of the ground truth (point-to-plane ICP only)
  int64_t n_est_ = 0, n_gt_ = 0;         // point counts after VoxelDownSample (the clouds the metrics see)
  double t1, t2, t3, t4, t5, t6, t7, t_fcd = 0.0, t_acc = 0.0;
  double t_vmd = 0.0, t_v = 0.0, t_cdf = 0.0, t_scs = 0.0;
  std::string subfolder, results_subfolder, results_file_path;
  std::ofstream file_result;
};

inline int MapEvalB200::process() {
  TicToc tic_toc;
  std::string err;
  if (!validate(param_, &err)) {
    std::cerr << "ERROR: " << err << std::endl;
    return -1;
  }
  // point-to-plane ICP: Open3D needs normals on the target cloud; they can only come from the ground-truth file
  std::vector<double> *want_normals = (!param_.evaluate_using_initial_ && param_.evaluation_method_ == 1) ? &gt_normals_ : nullptr;
  std::string file_extension = param_.map_gt_path_.substr(param_.map_gt_path_.find_last_of(".") + 1);
  bool gt_ok = false;
  if (file_extension == "pcd") gt_ok = cloud_io::read_pcd(param_.map_gt_path_, gt_3d_, &err, want_normals);
  else if (file_extension == "ply") gt_ok = cloud_io::read_ply(param_.map_gt_path_, gt_3d_, &err, want_normals);
  else {
    std::cerr << "ERROR: Unsupported ground truth file format: " << param_.map_gt_path_ << std::endl;
    return -1;
  }
  if (!gt_ok && param_.enable_debug) std::cerr << "WARNING: " << err << std::endl;   // the reference ignores this return value
  bool success = cloud_io::read_pcd(param_.evaluation_map_pcd_path_ + param_.pcd_file_name_, map_3d_, &err);
  if (param_.enable_debug)
    std::cout << "INFO: Loading map point cloud from: " << param_.evaluation_map_pcd_path_ + param_.pcd_file_name_ << std::endl;
  if (!success) {
    std::cerr << "ERROR: Failed to load point cloud from the specified path." << std::endl;
    return -1;
  }
  if (map_3d_.empty() || gt_3d_.empty()) {
    std::cerr << "ERROR: One or both point clouds are empty!" << std::endl;
    return -1;
  }
  if (want_normals && gt_normals_.size() != gt_3d_.size()) {
    std::cerr << "ERROR: TransformationEstimationPointToPlane requires pre-computed normal vectors for the target PointCloud: "
              << param_.map_gt_path_ << " carries none (registration_methods: 1)." << std::endl;
    return -1;
  }
  if (!gpus_.create(std::max(1, param_.n_gpus_), param_.gpu_device_, param_.vmd_voxel_size_)) return fail("cannot create the B200 context(s)");
  ctx_ = gpus_.ctx(0);
  if (gpus_.size() > 1) std::cout << "INFO: sharding the sweeps over " << gpus_.size() << " GPUs (NCCL all-reduce of the accumulators)" << std::endl;
  // every GPU holds both clouds (the lattices are replicated); VoxelDownSample (map_eval.cpp:38-39) runs on each of them.
  // The reference calls it unconditionally (Open3D raises for voxel_size <= 0); a non-positive value skips the step here.
  std::vector<int64_t> ne(gpus_.size(), (int64_t)(map_3d_.size() / 3)), ng(gpus_.size(), (int64_t)(gt_3d_.size() / 3));
  if (gpus_.for_each([&](int r, me_ctx *c) {
        int rc = me_set_cloud(c, ME_CLOUD_EST, map_3d_.data(), (int64_t)(map_3d_.size() / 3));
        if (rc == ME_OK) rc = me_set_cloud(c, ME_CLOUD_GT, gt_3d_.data(), (int64_t)(gt_3d_.size() / 3));
        if (rc == ME_OK && param_.downsample_size > 0) {
          rc = me_voxel_downsample(c, ME_CLOUD_EST, param_.downsample_size, &ne[r]);
          if (rc == ME_OK) rc = me_voxel_downsample(c, ME_CLOUD_GT, param_.downsample_size, &ng[r]);
        }
        return rc;
      }) != ME_OK)
    return fail("uploading / down-sampling the clouds");
  if (!(param_.downsample_size > 0)) std::cout << "INFO: downsample_size <= 0: clouds are evaluated as loaded." << std::endl;
  const int64_t n_est = ne[0], n_gt = ng[0];
  n_est_ = n_est; n_gt_ = n_gt;

  file_result << std::fixed << std::setprecision(15) << "Estimated-Ground Truth point count: " << n_est << " / " << n_gt
              << std::endl;
  if (param_.enable_debug)
    std::cout << "INFO: Loaded point clouds: " << n_est << " points (Map), " << n_gt << " points (Ground Truth)."
              << std::endl;

  if (param_.evaluate_mme_) {
    if (param_.enable_debug) std::cout << "INFO: Starting MME calculation..." << std::endl;
    if (computeMME() != 0) return -1;
    if (param_.enable_debug) std::cout << "INFO: MME calculation completed. Saving results." << std::endl;
    if (param_.save_immediate_result_) saveMmeResults();
    t2 = tic_toc.toc();
    if (param_.enable_debug) std::cout << "INFO: MME calculation completed in: " << (t2 - t1) / 1000.0 << " seconds." << std::endl;
  }
  if (param_.enable_debug) std::cout << "INFO: Starting registration..." << std::endl;
  if (param_.evaluate_using_initial_) {
    if (param_.enable_debug) std::cout << "INFO: Using initial matrix without registration." << std::endl;
    if (calculateMetricsWithInitialMatrix() != 0) return -1;
  } else {
    // map_eval.cpp:81 -> performRegistration() -> performICPRegistration() (:1366-1394) -> calculateMetrics(reg) (:1147-1202)
    if (performRegistration() != 0) return -1;
  }
  if (calculateVMD() != 0) return -1;
  if (param_.enable_debug) std::cout << "INFO: VMD calculation completed." << std::endl;
  if (param_.save_immediate_result_) {
    if (param_.enable_debug) std::cout << "INFO: Saving registration results..." << std::endl;
    saveRegistrationResults();
  }
  if (param_.enable_debug) std::cout << "INFO: Results saved successfully." << std::endl;
  return 0;
}

// ColorPointCloudByMME (map_eval.cpp:686-736) on the entropies of the sweep that just ran (each GPU holds its shard)
inline int MapEvalB200::renderEntropy(int which, int64_t n, render::ColoredCloud *out) {
  std::vector<double> ent((size_t)n, 0.0), part((size_t)n), xyz((size_t)n * 3);
  for (int r = 0; r < gpus_.size(); ++r) {
    if (me_get_entropies(gpus_.ctx(r), which, part.data()) != ME_OK) { ctx_ = gpus_.ctx(r); const int rc = fail("me_get_entropies"); ctx_ = gpus_.ctx(0); return rc; }
    for (int64_t i = 0; i < n; ++i) ent[i] += part[i];          // disjoint shards: the other ranks hold zeros
  }
  int64_t got = 0;
  if (me_get_cloud(ctx_, which, xyz.data(), n, &got) != ME_OK || got != n) return fail("me_get_cloud");
  *out = render::color_by_entropy(xyz, ent);
  return 0;
}

// renderDistanceOnPointCloud (map_eval.cpp:586-606) for the raw map (every estimated point against the ground truth) and
// the inlier map (the corresponding points against each other), both clipped at accuracy_level[0] (:485-487).
inline int MapEvalB200::renderDistances(bool cutoff_strict) {
  const int64_t n = n_est_;
  std::vector<int32_t> idx((size_t)n, -1), pidx((size_t)n);
  std::vector<double> d2((size_t)n, std::nan("")), pd2((size_t)n), est((size_t)n * 3), gt((size_t)n_gt_ * 3);
  for (int r = 0; r < gpus_.size(); ++r) {
    if (me_get_nn(gpus_.ctx(r), ME_CLOUD_EST, pidx.data(), pd2.data()) != ME_OK) return fail("me_get_nn");
    for (int64_t i = 0; i < n; ++i)
      if (pidx[i] >= 0) { idx[i] = pidx[i]; d2[i] = pd2[i]; }
  }
  int64_t got = 0;
  if (me_get_cloud(ctx_, ME_CLOUD_EST, est.data(), n, &got) != ME_OK || got != n) return fail("me_get_cloud(est)");
  if (me_get_cloud(ctx_, ME_CLOUD_GT, gt.data(), n_gt_, &got) != ME_OK || got != n_gt_) return fail("me_get_cloud(gt)");
  const double dis = param_.trunc_dist_[0];
  map_3d_render_raw_ = render::color_by_distance(est, d2, dis);
  // corresponding_cloud_est / corresponding_cloud_gt: the kept est->gt pairs (:1083-1089); the reference's second call of
  // getDiffRegResultWithCorrespondence on path A overwrites them with index-swapped points (:1241) — not reproduced
  std::vector<double> ce, cg;
  const double R = param_.icp_max_distance_;
  for (int64_t i = 0; i < n; ++i) {
    if (idx[i] < 0) continue;
    if (!(cutoff_strict ? d2[i] < R * R : d2[i] <= R)) continue;
    for (int a = 0; a < 3; ++a) { ce.push_back(est[3 * i + a]); cg.push_back(gt[3 * (size_t)idx[i] + a]); }
  }
  map_3d_render_inlier_ = render::ColoredCloud();
  if (ce.empty()) return 0;
  me_options opt{};
  opt.abi_version = ME_ABI_VERSION; opt.device = param_.gpu_device_; opt.rank = 0; opt.world = 1;
  me_ctx *tmp = nullptr;
  if (me_create(&opt, &tmp) != ME_OK) return fail("me_create (inlier map)");
  const int64_t nc = (int64_t)(ce.size() / 3);
  me_nn_params p{};
  p.icp_max_distance = R; p.cutoff_mode = ME_CUTOFF_DIST_LT_R; p.pairing = ME_PAIRING_GEOMETRIC; p.want_full_cd = 1; p.directions = 1;
  me_nn_accum acc;
  std::vector<int32_t> ci((size_t)nc);
  std::vector<double> cd2((size_t)nc);
  int rc = me_set_cloud(tmp, ME_CLOUD_EST, ce.data(), nc);
  if (rc == ME_OK) rc = me_set_cloud(tmp, ME_CLOUD_GT, cg.data(), nc);
  if (rc == ME_OK) rc = me_eval_nn_accum(tmp, &p, &acc, nullptr);
  if (rc == ME_OK) rc = me_get_nn(tmp, ME_CLOUD_EST, ci.data(), cd2.data());
  if (rc != ME_OK) std::cerr << "ERROR: inlier distance map: " << me_last_error(tmp) << std::endl;
  me_destroy(tmp);
  if (rc != ME_OK) return -1;
  map_3d_render_inlier_ = render::color_by_distance(ce, cd2, dis);
  return 0;
}

inline int MapEvalB200::computeMME() {
  if (!param_.evaluate_mme_) return 0;
  // use_tbb_mme only selects between two CPU threadings of the same arithmetic in the reference (:153-157)
  auto eval_mme = [&](int which, int min_neighbors, int64_t n_total, me_mme_result *out) -> int {
    std::vector<me_mme_accum> acc(gpus_.size());
    if (gpus_.for_each([&](int r, me_ctx *c) { return me_eval_mme_accum(c, which, param_.nn_radius_, min_neighbors, &acc[r]); }) != ME_OK)
      return -1;
    if (!gpus_.reduce_mme(acc)) return -1;
    return me_mme_finalize(&acc[0], n_total, out) == ME_OK ? 0 : -1;
  };
  if (eval_mme(ME_CLOUD_EST, 10, n_est_, &mme_est_res_) != 0) return fail("me_eval_mme(est)");
  if (param_.save_immediate_result_ && renderEntropy(ME_CLOUD_EST, n_est_, &map_3d_entropy_) != 0) return -1;   // :179
  mme_est = mme_est_res_.mme;
  if (mme_est_res_.n_valid * 100.0 / (double)mme_est_res_.n_total < 0.6)
    std::cerr << "valid points is too small, please check the input point cloud" << std::endl;   // :1732
  min_abs_entropy = mme_est_res_.min_abs_entropy; max_abs_entropy = mme_est_res_.max_abs_entropy;   // :179 -> :700-701
  if (param_.evaluate_gt_mme_) {
    if (eval_mme(ME_CLOUD_GT, 5, n_gt_, &mme_gt_res_) != 0) return fail("me_eval_mme(gt)");
    if (param_.save_immediate_result_ && renderEntropy(ME_CLOUD_GT, n_gt_, &gt_3d_entropy_) != 0) return -1;      // :181
    mme_gt = mme_gt_res_.mme;
    std::cout << "GT MME Valid_points " << mme_gt_res_.n_valid * 100.0 / (double)mme_gt_res_.n_total << "% " << mme_gt_res_.n_valid
              << " " << mme_gt_res_.n_total << std::endl;
    std::cout << "MME EST-GT: " << mme_est << " " << mme_gt << std::endl;
    min_abs_entropy = mme_gt_res_.min_abs_entropy; max_abs_entropy = mme_gt_res_.max_abs_entropy;   // :181, last call wins
  } else {
    std::cout << "MME EST: " << mme_est << std::endl;
  }
  return 0;
}

inline int MapEvalB200::calculateMetricsWithInitialMatrix() {
  TicToc tic;
  if (!param_.trunc_dist_set_)
    std::cerr << "WARNING: accuracy_level missing: the reference would read an uninitialised trunc_dist_ (map_eval.h:85); using zeros."
              << std::endl;
  me_nn_params p{};
  for (int i = 0; i < 5; ++i) p.tau[i] = param_.trunc_dist_[i];
  p.icp_max_distance = param_.icp_max_distance_;
  p.cutoff_mode = ME_CUTOFF_SQDIST_LE_R;
  p.pairing = param_.geometric_gt_pairing_ ? ME_PAIRING_GEOMETRIC : ME_PAIRING_AS_WRITTEN;
  // path A never calls computeChamferDistance (full_chamfer_dist stays 0, map_eval.h:334); the sums are still cheap,
  // so they are produced and reported on stdout, while the results file keeps the reference's value
  p.want_full_cd = 1;
  p.directions = 3;
  std::vector<me_nn_accum> e2g(gpus_.size()), g2e(gpus_.size());
  if (gpus_.for_each([&](int r, me_ctx *c) {
        int rc = me_transform(c, ME_CLOUD_EST, param_.initial_matrix_);      // map_3d_->Transform(initial_matrix), :1206
        return rc != ME_OK ? rc : me_eval_nn_accum(c, &p, &e2g[r], &g2e[r]);
      }) != ME_OK)
    return fail("me_transform / me_eval_nn_accum");
  if (!gpus_.reduce_nn(e2g, g2e)) return fail("all-reduce of the NN accumulators");
  if (me_nn_finalize(&p, &e2g[0], &g2e[0], n_est_, n_gt_, &nn_) != ME_OK) return fail("me_nn_finalize");
  t_acc = tic.toc() / 1000.0;
  if (nn_.gt_to_est.n_ub > 0)
    std::cerr << "WARNING: " << nn_.gt_to_est.n_ub << " gt->est pairs index past the end of a cloud (the reference reads out of "
                 "bounds there, map_eval.cpp:1233 vs :1093-1094: undefined behaviour); they are left out of the gt->est statistics, "
                 "so the CD vector / gt->est lines are not comparable with a reference run.  geometric_gt_pairing: true pairs "
                 "the points geometrically instead." << std::endl;
  if (param_.save_immediate_result_ && renderDistances(false) != 0) return -1;
  std::cout << "INFO: Chamfer Distance: " << eigenRow(nn_.cd, 5, 6) << std::endl;
  std::cout << "INFO: F1 Score: " << eigenRow(nn_.f1, 5, 6) << std::endl;
  std::cout << "INFO: est-gt MME: " << mme_est << " " << mme_gt << std::endl;
  std::cout << "INFO: IoU: " << eigenRow(nn_.iou, 5, 6) << std::endl;
  std::cout << "INFO: Full Chamfer distance (not computed by the reference on this path): " << nn_.full_cd << std::endl;
  return 0;
}

inline int MapEvalB200::performRegistration() {
  TicToc tic_toc;
  if (gpus_.size() > 1) std::cout << "INFO: ICP runs on GPU " << param_.gpu_device_ << " only; the metric sweeps are sharded." << std::endl;
  // ICP needs all correspondences on one GPU: run it on a single-shard context, then hand every GPU the aligned cloud
  me_icp_result reg{};
  me_set_shard(ctx_, 0, 1);
  int rc = ME_OK;
  if (param_.evaluation_method_ == ME_ICP_POINT_TO_PLANE)
    rc = me_set_normals(ctx_, ME_CLOUD_GT, gt_normals_.data(), (int64_t)(gt_normals_.size() / 3));
  // performICPRegistration (:1366-1394): ICPConvergenceCriteria() = {1e-6, 1e-6, 30}
  if (rc == ME_OK) rc = me_icp(ctx_, param_.evaluation_method_, param_.icp_max_distance_, 30, 1e-6, 1e-6, param_.initial_matrix_, &reg);
  me_set_shard(ctx_, 0, gpus_.size());
  if (rc != ME_OK) return fail("me_icp");
  t4 = tic_toc.toc();
  std::cout << "INFO: ICP registration time: " << (t4 - t3) / 1000.0 << " [s]" << std::endl;
  std::cout << "INFO: Aligned transformation: \n" << eigenMatrix4(reg.transformation, 6) << std::endl;
  std::cout << "INFO: ICP overlap ratio: " << reg.fitness << std::endl;
  std::cout << "INFO: ICP correspondences RMSE: " << reg.inlier_rmse << std::endl;
  std::cout << "INFO: ICP correspondences size: " << reg.n_corr << std::endl;
  file_result << std::fixed << std::setprecision(5) << "Aligned cloud: " << eigenMatrix4(reg.transformation, 5) << std::endl;
  file_result << std::fixed << std::setprecision(5) << "Aligned results: " << reg.fitness << " " << reg.n_corr << std::endl;
  // the other GPUs apply the same transformation to their copy of the (down-sampled) cloud
  for (int r = 1; r < gpus_.size(); ++r)
    if (me_transform(gpus_.ctx(r), ME_CLOUD_EST, reg.transformation) != ME_OK) return fail("me_transform (aligned cloud)");
  // calculateMetrics(reg) (:1147-1202): est->gt on the ICP correspondences (NN within R, strict), gt->est through
  // EvaluateRegistration(gt, est, R) (:1168), cd_vec (:1171), full Chamfer distance (:1194)
  TicToc tic;
  me_nn_params p{};
  for (int i = 0; i < 5; ++i) p.tau[i] = param_.trunc_dist_[i];
  p.icp_max_distance = param_.icp_max_distance_;
  p.cutoff_mode = ME_CUTOFF_DIST_LT_R;
  p.pairing = ME_PAIRING_GEOMETRIC;
  p.want_full_cd = 1;
  p.directions = 3;
  std::vector<me_nn_accum> e2g(gpus_.size()), g2e(gpus_.size());
  if (gpus_.for_each([&](int r, me_ctx *c) { return me_eval_nn_accum(c, &p, &e2g[r], &g2e[r]); }) != ME_OK) return fail("me_eval_nn_accum");
  if (!gpus_.reduce_nn(e2g, g2e)) return fail("all-reduce of the NN accumulators");
  if (me_nn_finalize(&p, &e2g[0], &g2e[0], n_est_, n_gt_, &nn_) != ME_OK) return fail("me_nn_finalize");
  t_acc = tic.toc() / 1000.0;
  if (param_.save_immediate_result_ && renderDistances(true) != 0) return -1;
  full_chamfer_dist = nn_.full_cd;
  t_fcd = t_acc;
  t5 = tic_toc.toc();
  std::cout << "INFO: Chamfer Distance: " << eigenRow(nn_.cd, 5, 6) << std::endl;
  std::cout << "INFO: Full Chamfer Distance: " << full_chamfer_dist << std::endl;
  return 0;
}

inline int MapEvalB200::calculateVMD() {
  TicToc ticToc;
  int64_t n_rows = 0;
  double *rows = nullptr;
  if (me_eval_awd(ctx_, param_.vmd_voxel_size_, 100, 5, &awd_, &n_rows, &rows) != ME_OK) return fail("me_eval_awd");
  std::cout << "Build voxel map: " << awd_.n_voxels_gt << std::endl;
  std::cout << "Build voxel map: " << awd_.n_voxels_est << std::endl;
  std::cout << "Update active/old/new voxel num: " << awd_.n_active << " " << awd_.n_old << " " << awd_.n_new << std::endl;
  t_v = ticToc.toc();
  std::ofstream output_file(results_subfolder + "voxel_errors.txt");
  if (!output_file.is_open()) {
    std::cerr << "ERROR: Failed to open voxel error output file." << std::endl;
    me_free(rows);
    return 0;   // the reference returns early here (map_eval.cpp:257-260)
  }
  for (int64_t r = 0; r < n_rows; ++r) {   // map_eval.cpp:292-302: default ostream formatting, one space apart
    const double *v = rows + r * 27;
    for (int c = 0; c < 27; ++c) {
      if (c == 10 || c == 11) output_file << (long long)v[c];
      else output_file << v[c];
      output_file << (c == 26 ? "" : " ");
    }
    output_file << std::endl;
  }
  output_file.close();
  if (param_.enable_debug) std::cout << "INFO: Voxel errors results saved to " << results_subfolder + "voxel_errors.txt" << std::endl;
  vmd = awd_.awd;
  t_vmd = ticToc.toc();
  std::cout << "INFO: Calculated VMD: " << vmd << std::endl;
  std::vector<double> ws(n_rows);
  for (int64_t r = 0; r < n_rows; ++r) ws[r] = rows[r * 27 + 9];
  me_free(rows);
  std::sort(ws.begin(), ws.end());   // map_eval.cpp:330-341
  std::ofstream cdf_file(results_subfolder + "voxel_wasserstein_cdf.txt");
  if (!cdf_file.is_open()) {
    std::cerr << "ERROR: Failed to open CDF output file." << std::endl;
    return 0;
  }
  for (size_t i = 0; i < ws.size(); ++i) cdf_file << ws[i] << " " << static_cast<double>(i + 1) / ws.size() << std::endl;
  cdf_file.close();
  t_cdf = ticToc.toc();
  if (param_.enable_debug) std::cout << "INFO: CDF results saved to " << results_subfolder + "voxel_wasserstein_cdf.txt" << std::endl;
  scs_overall = awd_.scs;
  t_scs = ticToc.toc();
  std::cout << "INFO: Spatial Consistency Score (SCS): " << scs_overall << std::endl;
  return 0;
}

inline void MapEvalB200::saveMmeResults() {   // map_eval.cpp:392-421
  if (!param_.evaluate_mme_) return;
  file_result << std::fixed << std::setprecision(5) << "MME: " << mme_est << " " << mme_gt << " " << min_abs_entropy << " "
              << max_abs_entropy << std::endl;
  if (param_.enable_debug) std::cout << "INFO: MME results saved to " << results_file_path << std::endl;
  // map_eval.cpp:404-412
  if (!render::write_pcd(results_subfolder + "map_entropy.pcd", map_3d_entropy_)) std::cerr << "ERROR: cannot write map_entropy.pcd" << std::endl;
  else if (param_.enable_debug) std::cout << "INFO: Saved rendered entropy map to " << results_subfolder + "map_entropy.pcd" << std::endl;
  if (param_.evaluate_gt_mme_) {
    if (!render::write_pcd(results_subfolder + "gt_entropy.pcd", gt_3d_entropy_)) std::cerr << "ERROR: cannot write gt_entropy.pcd" << std::endl;
    else if (param_.enable_debug)
      std::cout << "INFO: Saved rendered entropy ground truth map to " << results_subfolder + "gt_entropy.pcd" << std::endl;
  }
}

inline void MapEvalB200::saveRegistrationResults() {   // map_eval.cpp:424-482
  if (param_.enable_debug) {
    std::cout << "INFO: AC+MME Time: " << t_acc + (t2 - t1) / 1000.0 << std::endl;
    std::cout << "INFO: CD+MME Time: " << t_fcd + (t2 - t1) / 1000.0 << std::endl;
    std::cout << "INFO: AWD+SCS Time: " << t_v / 1000.0 + (t_vmd - t_v) / 1000.0 + (t_scs - t_cdf) / 1000.0 << std::endl;
  }
  file_result << std::fixed << std::setprecision(15) << "RMSE/AC: " << eigenRow(nn_.est_to_gt.rmse, 5, 15) << std::endl;
  file_result << std::fixed << std::setprecision(15) << "Comp: " << eigenRow(nn_.est_to_gt.fitness, 5, 15) << std::endl;
  file_result << std::fixed << std::setprecision(5) << "FULL CD: " << full_chamfer_dist << std::endl;   // 0 on path A, as the reference
  file_result << std::fixed << std::setprecision(5) << "VMD: " << vmd << std::endl;
  file_result << std::fixed << std::setprecision(5) << "SCS: " << scs_overall << std::endl;
  file_result << "Time load-MME-mesh-ICP-Metric-AC-FCD: " << t1 / 1000.0 << " " << (t2 - t1) / 1000.0 << " " << (t3) / 1000.0 << " "
              << (t4 - t3) / 1000.0 << " " << (t5 - t4) / 1000.0 << " " << t_acc << " " << t_fcd << std::endl;
  file_result << "VMD Time voxelization-WD-CDF-SCS: " << t_v / 1000.0 << " " << (t_vmd - t_v) / 1000.0 << " "
              << (t_cdf - t_vmd) / 1000.0 << " " << (t_scs - t_cdf) / 1000.0 << std::endl;
  file_result << "AC+MME Time: " << t_acc + (t2 - t1) / 1000.0 << std::endl;
  file_result << "CD+MME Time: " << t_fcd + (t2 - t1) / 1000.0 << std::endl;
  file_result << "AWD+SCS Time: " << t_v / 1000.0 + (t_vmd - t_v) / 1000.0 + (t_scs - t_cdf) / 1000.0 << std::endl;
  file_result.close();
  if (param_.enable_debug) std::cout << "INFO: Results saved to " << results_subfolder + "map_results.txt" << std::endl;
  // map_eval.cpp:485-499
  if (!render::write_pcd(results_subfolder + "raw_rendered_dis_map.pcd", map_3d_render_raw_)) std::cerr << "ERROR: cannot write raw_rendered_dis_map.pcd" << std::endl;
  else if (param_.enable_debug) std::cout << "INFO: Saved raw distance error map to " << results_subfolder + "raw_rendered_dis_map.pcd" << std::endl;
  if (!render::write_pcd(results_subfolder + "inlier_rendered_dis_map.pcd", map_3d_render_inlier_)) std::cerr << "ERROR: cannot write inlier_rendered_dis_map.pcd" << std::endl;
  else if (param_.enable_debug) std::cout << "INFO: Saved inlier distance error map to " << results_subfolder + "inlier_rendered_dis_map.pcd" << std::endl;
}
