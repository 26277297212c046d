// rmsds -- all-to-all (or trajectory x trajectory) RMSD matrix.  Same command
// line, selection strings, stderr lines and stdout format as Tools/rmsds.cpp
// (options :134-161, flow :490-570); the thread pool over centeredRMSD is replaced
// by loosb200_rmsds_self / loosb200_rmsds_cross.
#include "tool_common.hpp"

using namespace lbfe;

int main(int argc, char* argv[]) {
  const std::string header = invocationHeader(argc, argv);
  Options o;
  addBasicOptions(o);
  o.add("noout", 'N', "0", "Do not output the matrix (i.e. only calc pair-wise RMSD stats)", true);
  o.add("threads", 0, "1", "Number of threads to use (0=all available); accepted, unused on the GPU path");
  o.add("sel1", 0, "name == 'CA'", "Atom selection for first system");
  o.add("skip1", 0, "0", "Skip n-frames of first trajectory");
  o.add("range1", 0, "", "Matlab-style range of frames to use from first trajectory");
  o.add("sel2", 0, "name == 'CA'", "Atom selection for second system");
  o.add("skip2", 0, "0", "Skip n-frames of second trajectory");
  o.add("range2", 0, "", "Matlab-style range of frames to use from second trajectory");
  o.add("stats", 0, "0", "Show some statistics for matrix", true);
  o.add("binary", 0, "", "Also write the matrix as float32 NumPy .npy to this file (extension; written even with --noout 1)");
  o.add("precision", 'p', "2", "Write out matrix coefficients with this many digits.");
  o.addPositional("model1", 1); o.addPositional("traj1", 1); o.addPositional("model2", 1); o.addPositional("traj2", 1);
  try {
    if (!o.parse(argc, argv)) return 255;  // exit(-1)
    const bool bad = !((o.has("model1") && o.has("traj1")) && !(o.has("model2") ^ o.has("traj2")));  // check(), :164-166
    if (o.has("help") || o.has("fullhelp") || bad) {
      std::cerr << o.usage(argv[0], "model-1 trajectory-1 [model-2 trajectory-2]");
      return 255;
    }
    const long verbosity = std::stol(o.str("verbosity"));
    const bool noop = o.flag("noout"), stats = o.flag("stats");
    const std::string binary = o.str("binary");
    const unsigned precision = (unsigned)o.uns("precision");
    const bool gpu_text = !noop && gpuText(!binary.empty(), precision);  // text formatted on the GPU
    const bool want_m = (!noop || !binary.empty()) && !gpu_text;
    const int device = (int)o.uns("gpu"), engine = engineOf(o), ngpus = gpusOf(o);

    Model model = readPDB(o.str("model1"));
    const std::shared_ptr<Trajectory> trajp = openTrajectory(o.str("traj1"), (uint32_t)model.size());
    Trajectory& traj = *trajp;
    if (traj.natoms() != model.size()) throw Error("Error- trajectory has " + std::to_string(traj.natoms()) + " atoms but the model has " + std::to_string(model.size()));
    const std::vector<uint32_t> picked = selectAtoms(model, o.str("sel1"));
    if (picked.empty()) throw Error("Error- selection '" + o.str("sel1") + "' matched no atoms");
    const std::vector<uint32_t> sel = selectionToFrameIndices(model, picked);
    const std::vector<uint32_t> indices = assignTrajectoryFrames(traj.nframes(), o.str("range1"), (uint32_t)o.uns("skip1"));
    if (verbosity > 1) std::cerr << "Reading trajectory - " << o.str("traj1") << std::endl;
    Frames A;
    stageTrajectory(A, traj, indices, model.size(), sel, device);

    Replicas RA;
    RA.build(A, device, ngpus);

    HostMatrix M;
    size_t rows = indices.size(), cols = indices.size();
    loosb200_stats st;
    if (!o.has("model2")) {
      if (verbosity > 1) std::cerr << "Calculating RMSD...\n";
      warnHostMemory(want_m ? rows * cols * sizeof(float) : 0, verbosity);
      if (want_m) M.allocate(rows * cols);
      if (gpu_text) {
        std::cout << "# " << header << std::endl;
        check(loosb200_rmsds_self_multi_text(RA.h.data(), ngpus, precision, stdoutSink, nullptr, &st, engine), "rmsds");
      }
      else check(loosb200_rmsds_self_multi(RA.h.data(), ngpus, want_m ? M.data() : nullptr, rows, &st, engine), "rmsds");
    } else {
      Model model2 = readPDB(o.str("model2"));
      const std::shared_ptr<Trajectory> traj2p = openTrajectory(o.str("traj2"), (uint32_t)model2.size());
      Trajectory& traj2 = *traj2p;
      if (traj2.natoms() != model2.size()) throw Error("Error- second trajectory and model differ in size");
      const std::vector<uint32_t> picked2 = selectAtoms(model2, o.str("sel2"));
      const std::vector<uint32_t> sel2 = selectionToFrameIndices(model2, picked2);
      const std::vector<uint32_t> indices2 = assignTrajectoryFrames(traj2.nframes(), o.str("range2"), (uint32_t)o.uns("skip2"));
      if (verbosity > 1) std::cerr << "Reading trajectory - " << o.str("traj2") << std::endl;
      Frames B;
      stageTrajectory(B, traj2, indices2, model2.size(), sel2, device);
      cols = indices2.size();
      Replicas RB;
      RB.build(B, device, ngpus);
      if (verbosity > 1) std::cerr << "Calculating RMSD...\n";
      warnHostMemory(want_m ? rows * cols * sizeof(float) : 0, verbosity);
      if (want_m) M.allocate(rows * cols);
      if (gpu_text) {
        std::cout << "# " << header << std::endl;
        check(loosb200_rmsds_cross_multi_text(RA.h.data(), RB.h.data(), ngpus, precision, stdoutSink, nullptr, &st, engine), "rmsds");
      }
      else check(loosb200_rmsds_cross_multi(RA.h.data(), RB.h.data(), ngpus, want_m ? M.data() : nullptr, rows, &st, engine), "rmsds");
    }
    if (verbosity || noop || stats) {  // showStatsHalf / showStatsWhole, :425-456
      char line[128];
      snprintf(line, sizeof(line), "Max rmsd = %.4f, avg rmsd = %.4f\n", st.max, st.count ? st.sum / (double)st.count : 0.0);
      std::cerr << line;
    }
    if (!binary.empty()) writeNpy(binary, M.data(), rows, cols, rows);
    if (gpu_text) fflush(stdout);
    if (!noop && !gpu_text) {
      std::cout << "# " << header << std::endl;
      writeMatrix(std::cout, M.data(), rows, cols, rows, precision);
    }
  } catch (const std::exception& e) {
    std::cerr << e.what() << std::endl;
    return 255;
  }
  return 0;
}
