// map_eval — the re-hosted MapEval executable (reference: map_eval/src/map_eval_main.cpp:211-244).
//
//   map_eval [config.yaml] [--gpus N] [--dump-config] [--read-cloud file.pcd|file.ply]
//
// Without an argument the configuration is read from ../config/config.yaml relative to the working directory, the
// reference's hard-coded path (map_eval_main.cpp:213; its argv handling is commented out at :214-216).
#include "map_eval_b200.hpp"

static void displayProgramInformation(const Param &param) {
  std::cout << "\n================================================================================\n"
            << "  map_eval (B200 hot path) — MapEval metric pipeline on sm_100a\n"
            << "================================================================================\n"
            << "  scene                 : " << param.name_ << "\n"
            << "  ground truth map      : " << param.map_gt_path_ << "\n"
            << "  estimated map         : " << param.evaluation_map_pcd_path_ + param.pcd_file_name_ << "\n"
            << "  results               : " << param.result_path_ << "\n"
            << "  icp_max_distance      : " << param.icp_max_distance_ << "\n"
            << "  accuracy_level        : " << param.trunc_dist_[0] << " " << param.trunc_dist_[1] << " " << param.trunc_dist_[2] << " "
            << param.trunc_dist_[3] << " " << param.trunc_dist_[4] << "\n"
            << "  evaluate_using_initial: " << (param.evaluate_using_initial_ ? "true" : "false") << "\n"
            << "  evaluate_mme / gt     : " << (param.evaluate_mme_ ? "true" : "false") << " / " << (param.evaluate_gt_mme_ ? "true" : "false")
            << "  (nn_radius " << param.nn_radius_ << ")\n"
            << "  vmd_voxel_size        : " << param.vmd_voxel_size_ << "\n"
            << "  downsample_size       : " << param.downsample_size << "\n"
            << "================================================================================\n\n";
}

int main(int argc, char **argv) {
  std::string config_file = "../config/config.yaml";
  bool dump = false;
  int gpus_override = 0;
  for (int i = 1; i < argc; ++i) {
    std::string a = argv[i];
    if (a == "--dump-config") dump = true;
    else if (a == "--read-cloud" && i + 1 < argc) {   // reader self-check: point count and coordinate sums
      std::string path = argv[++i], err;
      std::vector<double> xyz;
      const bool is_ply = path.size() > 4 && path.substr(path.size() - 4) == ".ply";
      if (!(is_ply ? cloud_io::read_ply(path, xyz, &err) : cloud_io::read_pcd(path, xyz, &err))) {
        std::cerr << "ERROR: " << err << "\n";
        return EXIT_FAILURE;
      }
      double sx = 0, sy = 0, sz = 0;
      for (size_t k = 0; k + 2 < xyz.size(); k += 3) { sx += xyz[k]; sy += xyz[k + 1]; sz += xyz[k + 2]; }
      std::cout << std::setprecision(17) << "points " << xyz.size() / 3 << " sum " << sx << " " << sy << " " << sz << "\n";
      return EXIT_SUCCESS;
    } else if (a == "--gpus" && i + 1 < argc) gpus_override = std::atoi(argv[++i]);
    else config_file = a;
  }
  std::cout << "Loading configuration from: " << config_file << "\n";
  Param param;
  try {
    param = loadParametersFromYAML(config_file);
  } catch (const std::exception &e) {
    std::cerr << "\n[ERROR] Failed to load configuration: " << e.what() << "\n";
    return EXIT_FAILURE;
  }
  if (gpus_override > 0) param.n_gpus_ = gpus_override;
  if (dump) {
    dumpParam(param, std::cout);
    return EXIT_SUCCESS;
  }
  {
    std::string why;
    if (!MapEvalB200::validate(param, &why)) {      // before the constructor opens map_results.txt
      std::cerr << "\n[ERROR] " << why << "\n";
      return EXIT_FAILURE;
    }
  }
  displayProgramInformation(param);
  std::cout << "Starting evaluation...\n================================================================================\n\n";
  MapEvalB200 map_eval(param);
  const int rc = map_eval.process();   // the reference ignores this value (map_eval_main.cpp:237); we report it
  std::cout << "\n================================================================================\n"
            << (rc == 0 ? "Evaluation completed successfully!" : "Evaluation FAILED (see the messages above).")
            << "\n================================================================================\n\n";
  return rc == 0 ? EXIT_SUCCESS : EXIT_FAILURE;
}
