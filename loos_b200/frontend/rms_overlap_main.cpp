// rms-overlap -- RMSD between every frame of trajectory set A and every frame of set B.
// Command line and output as Tools/rms-overlap.cpp (:163-215,580-648): the same
// centeredRMSD loop as rmsds' two-trajectory mode, over two MultiTrajectories.
// Reference quirks kept for drop-in parity and documented here:
//   * the default selection "name == 'backbone' && ! hydrogen" (:584) selects nothing on
//     ordinary models -- pass --selection;
//   * showFractionalStats (:458-507) lacks braces after `if (R(i,j) > max)`, so its "Max
//     rmsd ... between frames" is always the LAST off-diagonal element, and it skips i == j
//     although the matrix is rectangular.  The same numbers are printed here.
#include <sstream>

#include "tool_common.hpp"

using namespace lbfe;

static std::vector<std::string> splitSpaces(const std::string& s) {  // boost::split(is_any_of(" "))
  std::vector<std::string> out;
  std::string cur;
  for (char c : s) { if (c == ' ') { out.push_back(cur); cur.clear(); } else cur += c; }
  out.push_back(cur);
  return out;
}

static std::string trajectoryTable(const MultiTraj& mt, const std::string& range) {  // :223-246
  std::ostringstream oss;
  if (!range.empty()) oss << "# Note- composite frame range used was '" << range << "'\n";
  oss << "# traj\tstart\tend\tfilename\n";
  uint32_t start_cnt = 0, j = 0;
  for (size_t i = 0; i < mt.trajs.size(); ++i) {
    const uint32_t n = mt.nframes(i);
    if (n == 0) oss << "# Warning- '" << mt.trajs[i]->filename() << "' was skipped due to insufficient frames\n";
    else { oss << "# " << j << "\t" << start_cnt << "\t" << (start_cnt + n - 1) << "\t" << mt.trajs[i]->filename() << "\n"; ++j; }
    start_cnt += n;
  }
  return oss.str();
}

int main(int argc, char* argv[]) {
  const std::string header = invocationHeader(argc, argv);
  Options o;
  addBasicOptions(o);
  o.add("selection", 's', "name == 'backbone' && ! hydrogen", "Which atoms to use");
  o.add("modeltype", 0, "", "Model types: pdb");
  o.add("set-A", 'A', "", "Space separated set of trajectories to compare pair-wise to B");
  o.add("set-B", 'B', "", "Space separated set of trajectories to compare pair-wise to A");
  o.add("skip", 'k', "0", "Number of frames to skip in sub-trajectories");
  o.add("stride", 'i', "1", "Step through sub-trajectories by this amount");
  o.add("range", 'r', "", "Which frames to use in composite trajectory");
  o.add("noout", 'N', "0", "Do not output the matrix (i.e. only calc pair-wise RMSD stats)", true);
  o.add("threads", 0, "1", "Number of threads to use (0=all available); accepted, unused on the GPU path");
  o.add("cutoff", 'c', "-1", "Outputs fraction of frame-pairs below cutoff.");
  o.add("stats", 0, "0", "Show some statistics for matrix", true);
  o.add("binary", 0, "", "Also write the matrix as float32 NumPy .npy to this file (extension; written even with --noout 1)");
  o.add("precision", 'p', "2", "Write out matrix coefficients with this many digits.");
  o.addPositional("model", 1);
  try {
    if (!o.parse(argc, argv)) return 255;
    if (o.has("help") || o.has("fullhelp") || !o.has("model") || o.str("set-A").empty() || o.str("set-B").empty()) {
      std::cerr << o.usage(argv[0], "model");
      return 255;
    }
    const long verbosity = std::stol(o.str("verbosity"));
    const bool noop = o.flag("noout"), stats = o.flag("stats");
    const float cutoff = (float)o.dbl("cutoff");
    const int device = (int)o.uns("gpu"), engine = engineOf(o), ngpus = gpusOf(o);
    const std::string range = o.str("range");
    Model model = readPDB(o.str("model"));
    MultiTraj mt[2];
    const std::string sets[2] = {o.str("set-A"), o.str("set-B")};
    for (int s = 0; s < 2; ++s) {
      mt[s].skip = (uint32_t)o.uns("skip"); mt[s].stride = (uint32_t)o.uns("stride");
      if (mt[s].stride == 0) throw Error("Error- stride must be at least 1");
      for (const std::string& name : splitSpaces(sets[s])) {
        mt[s].trajs.push_back(openTrajectory(name, (uint32_t)model.size()));
        if (mt[s].trajs.back()->natoms() != model.size()) throw Error("Error- '" + name + "' does not match the model size");
      }
    }
    const std::vector<uint32_t> picked = selectAtoms(model, o.str("selection"));
    if (picked.empty()) throw Error("Error- selection '" + o.str("selection") + "' matched no atoms");
    const std::vector<uint32_t> sel = selectionToFrameIndices(model, picked);
    Frames fs[2];
    size_t F[2];
    for (int s = 0; s < 2; ++s) {
      const std::vector<uint32_t> indices = uniquify(assignTrajectoryFrames(mt[s].nframes(), range, 0, 1));  // frameList(), :215-218
      for (uint32_t k : indices) if (k >= mt[s].nframes()) throw Error("Cannot seek past end of MultiTraj");
      F[s] = indices.size();
      stageStreaming({&fs[s]}, {&sel}, F[s], model.size(), device, [&](size_t j, float* dst) {
        const std::pair<uint32_t, uint32_t> loc = mt[s].location(indices[j]);
        mt[s].trajs[loc.first]->readFramePlanes(loc.second, dst);
      });
    }
    if (verbosity > 1) std::cerr << "Calculating RMSD...\n";
    const bool want_fraction = cutoff > 0;
    const std::string binary = o.str("binary");
    const unsigned precision = (unsigned)o.uns("precision");
    const bool gpu_text = !noop && gpuText(want_fraction || !binary.empty(), precision);  // text formatted on the GPU
    const bool need_matrix = (!noop || want_fraction || !binary.empty()) && !gpu_text;
    Replicas R0, R1;
    R0.build(fs[0], device, ngpus);
    R1.build(fs[1], device, ngpus);
    HostMatrix Mh;
    warnHostMemory(need_matrix ? F[0] * F[1] * sizeof(float) : 0, verbosity);
    if (need_matrix) Mh.allocate(F[0] * F[1]);
    const float* M = Mh.data();
    loosb200_stats st;
    if (gpu_text) {
      std::cout << "# " << header << std::endl;
      std::cout << trajectoryTable(mt[0], range) << trajectoryTable(mt[1], range);
      std::cout.flush();
      check(loosb200_rmsds_cross_multi_text(R0.h.data(), R1.h.data(), ngpus, precision, stdoutSink, nullptr, &st, engine), "rms-overlap");
      fflush(stdout);
    } else {
      check(loosb200_rmsds_cross_multi(R0.h.data(), R1.h.data(), ngpus, need_matrix ? Mh.data() : nullptr, F[0], &st, engine), "rms-overlap");
    }
    if (!binary.empty()) writeNpy(binary, M, F[0], F[1], F[0]);
    if (verbosity || noop || stats || want_fraction) {
      char line[512];
      if (want_fraction) {  // showFractionalStats, bug-for-bug (see the header comment)
        const size_t total = F[0] * F[1];
        double avg = 0.0, var = 0.0, mval = 0.0;
        unsigned mi = 0, mj = 0;
        unsigned long below = 0;
        for (size_t i = 0; i < F[0]; ++i)
          for (size_t j = 0; j < F[1]; ++j) {
            if (i == j) continue;
            const float v = M[j * F[0] + i];
            var += (double)v * (double)v; avg += v;
            if ((double)v > mval) mi = (unsigned)i;
            mj = (unsigned)j; mval = v;
            if (v < cutoff) ++below;
          }
        avg /= (double)total;
        var = var / (double)total - avg * avg;
        snprintf(line, sizeof(line), "Max rmsd = %.4f between frames %d, %d, avg rmsd = %.4f, variance = %.4f, frames below %.4f = %lu, total = %lu\n",
                 mval, mi, mj, avg, var, (double)cutoff, below, (unsigned long)total);
        (noop ? std::cout : std::cerr) << line;
      } else {  // showStats, :441-456
        snprintf(line, sizeof(line), "Max rmsd = %.4f, avg rmsd = %.4f\n", st.max, st.count ? st.sum / (double)st.count : 0.0);
        std::cerr << line;
      }
    }
    if (!noop && !gpu_text) {
      std::cout << "# " << header << std::endl;
      std::cout << trajectoryTable(mt[0], range) << trajectoryTable(mt[1], range);
      writeMatrix(std::cout, M, F[0], F[1], F[0], precision);
    }
  } catch (const std::exception& e) {
    std::cerr << e.what() << std::endl;
    return 255;
  }
  return 0;
}
