// multi-rmsds -- all-to-all RMSD over a concatenation of trajectories.  Command line
// and output as Tools/multi-rmsds.cpp (:119-125,367-420) with MultiTrajOptions
// (src/OptionsFramework.cpp:417-502): per-sub-trajectory skip/stride, composite
// --range, trajectory table between header and matrix.
#include "tool_common.hpp"

using namespace lbfe;

int main(int argc, char* argv[]) {
  const std::string header = invocationHeader(argc, argv);
  Options o;
  addBasicOptions(o);
  o.add("selection", 's', "name == 'CA'", "Which atoms to use");
  o.add("modeltype", 0, "", "Model types: pdb");
  o.add("skip", 'k', "0", "Number of frames to skip in sub-trajectories");
  o.add("stride", 'i', "1", "Step through sub-trajectories by this amount");
  o.add("range", 'r', "", "Which frames to use in composite trajectory");
  o.add("noout", 'N', "0", "Do not output the matrix (i.e. only calc pair-wise RMSD stats)", true);
  o.add("threads", 0, "1", "Number of threads to use (0=all available); accepted, unused on the GPU path");
  o.add("stats", 0, "0", "Show some statistics for matrix", true);
  o.add("binary", 0, "", "Also write the matrix as float32 NumPy .npy to this file (extension; written even with --noout 1)");
  o.add("precision", 'p', "2", "Write out matrix coefficients with this many digits.");
  o.addPositional("model", 1); o.addPositional("traj", -1);
  try {
    if (!o.parse(argc, argv)) return 255;
    if (o.has("help") || o.has("fullhelp") || !o.has("model") || !o.has("traj")) {
      std::cerr << o.usage(argv[0], "model traj [traj ...]");
      return 255;
    }
    const long verbosity = std::stol(o.str("verbosity"));
    const bool noop = o.flag("noout"), stats = o.flag("stats");
    const int device = (int)o.uns("gpu"), engine = engineOf(o), ngpus = gpusOf(o);
    const std::string range = o.str("range");

    Model model = readPDB(o.str("model"));
    MultiTraj mt;
    mt.skip = (uint32_t)o.uns("skip"); mt.stride = (uint32_t)o.uns("stride");
    if (mt.stride == 0) throw Error("Error- stride must be at least 1");
    for (const std::string& name : o.multi("traj")) {
      mt.trajs.push_back(openTrajectory(name, (uint32_t)model.size()));
      if (mt.trajs.back()->natoms() != model.size()) throw Error("Error- '" + name + "' does not match the model size");
    }
    const std::vector<uint32_t> picked = selectAtoms(model, o.str("selection"));
    if (picked.empty()) throw Error("Error- selection matched no atoms");
    const std::vector<uint32_t> sel = selectionToFrameIndices(model, picked);
    // frameList(): sorted + unique (src/OptionsFramework.cpp:460-464)
    const std::vector<uint32_t> indices = uniquify(assignTrajectoryFrames(mt.nframes(), range, 0, 1));
    const size_t F = indices.size(), na = model.size();
    for (uint32_t k : indices)
      if (k >= mt.nframes()) throw Error("Cannot seek past end of MultiTraj");
    Frames A;
    stageStreaming({&A}, {&sel}, F, na, device, [&](size_t j, float* dst) {
      const std::pair<uint32_t, uint32_t> loc = mt.location(indices[j]);
      mt.trajs[loc.first]->readFramePlanes(loc.second, dst);
    });
    Replicas RA;
    RA.build(A, device, ngpus);
    if (verbosity > 1) std::cerr << "Calculating RMSD...\n";
    HostMatrix M;
    const std::string binary = o.str("binary");
    const unsigned precision = (unsigned)o.uns("precision");
    const bool gpu_text = !noop && gpuText(!binary.empty(), precision);  // text formatted on the GPU
    const bool want_m = (!noop || !binary.empty()) && !gpu_text;
    auto preamble = [&]() {
      std::cout << "# " << header << std::endl;
      // trajectoryTable(), src/OptionsFramework.cpp:479-502
      if (!range.empty()) std::cout << "# Note- composite frame range used was '" << range << "'\n";
      std::cout << "# traj\tstart\tend\tfilename\n";
      uint32_t start_cnt = 0, j = 0;
      for (size_t i = 0; i < mt.trajs.size(); ++i) {
        const uint32_t n = mt.nframes(i);
        if (n == 0) std::cout << "# Warning- '" << mt.trajs[i]->filename() << "' was skipped due to insufficient frames\n";
        else { std::cout << "# " << j << "\t" << start_cnt << "\t" << (start_cnt + n - 1) << "\t" << mt.trajs[i]->filename() << "\n"; ++j; }
        start_cnt += n;
      }
      std::cout.flush();
    };
    warnHostMemory(want_m ? F * F * sizeof(float) : 0, verbosity);
    if (want_m) M.allocate(F * F);
    loosb200_stats st;
    if (gpu_text) {
      preamble();
      check(loosb200_rmsds_self_multi_text(RA.h.data(), ngpus, precision, stdoutSink, nullptr, &st, engine), "multi-rmsds");
      fflush(stdout);
    } else {
      check(loosb200_rmsds_self_multi(RA.h.data(), ngpus, want_m ? M.data() : nullptr, F, &st, engine), "multi-rmsds");
    }
    if (!binary.empty()) writeNpy(binary, M.data(), F, F, F);
    if (verbosity || noop || stats) {
      char line[128];
      snprintf(line, sizeof(line), "Max rmsd = %.4f, avg rmsd = %.4f\n", st.max, st.count ? st.sum / (double)st.count : 0.0);
      std::cerr << line;
    }
    if (!noop && !gpu_text) {
      preamble();
      writeMatrix(std::cout, M.data(), F, F, F, precision);
    }
  } catch (const std::exception& e) {
    std::cerr << e.what() << std::endl;
    return 255;
  }
  return 0;
}
