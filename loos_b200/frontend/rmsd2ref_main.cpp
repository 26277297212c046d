// rmsd2ref -- per-frame RMSD to a reference structure or to the iteratively aligned
// average.  Command line, stderr lines and stdout format as Tools/rmsd2ref.cpp
// (:97-137,152-260) with TrajectoryWithFrameIndices (src/OptionsFramework.cpp:339-413).
#include <cmath>

#include "tool_common.hpp"

using namespace lbfe;

static std::vector<double> coordsOf(const Model& m, const std::vector<uint32_t>& picked) {
  std::vector<double> v(3 * picked.size());
  for (size_t i = 0; i < picked.size(); ++i) { v[3 * i] = m[picked[i]].x; v[3 * i + 1] = m[picked[i]].y; v[3 * i + 2] = m[picked[i]].z; }
  return v;
}

int main(int argc, char* argv[]) {
  const std::string hdr = invocationHeader(argc, argv);
  Options o;
  addBasicOptions(o);
  o.add("skip", 'k', "0", "Number of frames to skip");
  o.add("modeltype", 0, "", "Model types: pdb");
  o.add("trajtype", 0, "", "Trajectory types: dcd");
  o.add("stride", 'i', "1", "Take every ith frame");
  o.add("range", 'r', "", "Which frames to use (matlab style range, overrides stride and skip)");
  o.add("align", 0, "name == 'CA'", "Align using this selection");
  o.add("rmsd", 0, "!(hydrogen || segid =~ 'SOLV|BULK')", "Compute the RMSD over this selection");
  o.add("target", 0, "", "Compute RMSD against this reference target (must have coordinates)");
  o.add("talign", 0, "", "Selection for target to use to align (default is to use --align)");
  o.add("trmsd", 0, "", "Compute the RMSD over this selection for the target (default is to use --rmsd)");
  o.add("tolerance", 0, "1e-06", "Tolerance to use for iterative alignment");
  o.addPositional("model", 1); o.addPositional("traj", 1);
  try {
    if (!o.parse(argc, argv)) return 255;
    if (o.has("help") || o.has("fullhelp") || !o.has("model") || !o.has("traj")) {
      std::cerr << o.usage(argv[0], "model trajectory");
      return 255;
    }
    const uint32_t skip = (uint32_t)o.uns("skip"), stride = (uint32_t)o.uns("stride");
    const std::string range = o.str("range");
    if (skip > 0 && !range.empty()) {  // src/OptionsFramework.cpp:375-379
      std::cerr << "Error- you cannot specify both a skip and a frame range...I might get confused!\n";
      return 255;
    }
    if (stride == 0) throw Error("Error- stride must be at least 1");
    const int device = (int)o.uns("gpu");
    const std::string alignment = o.str("align"), selection = o.str("rmsd"), target_name = o.str("target");
    std::string target_align = o.str("talign"), target_selection = o.str("trmsd");
    if (!target_name.empty()) {  // postConditions, Tools/rmsd2ref.cpp:122-137
      if (target_align.empty()) { target_align = alignment; std::cerr << "Warning: Using --align selection for target\n"; }
      if (target_selection.empty()) { target_selection = selection; std::cerr << "Warning: Using --rmsd selection for target\n"; }
    }
    std::cout << "# " << hdr << std::endl;

    Model molecule = readPDB(o.str("model"));
    const std::shared_ptr<Trajectory> trajp = openTrajectory(o.str("traj"), (uint32_t)molecule.size());
    Trajectory& traj = *trajp;
    if (traj.natoms() != molecule.size()) throw Error("Error- trajectory and model differ in size");
    const std::vector<uint32_t> subset = selectAtoms(molecule, selection);
    if (subset.empty()) throw Error("Error- --rmsd selection matched no atoms");
    const std::vector<uint32_t> indices = uniquify(assignTrajectoryFrames(traj.nframes(), range, skip, stride));
    const size_t F = indices.size();
    char line[512];

    Model target;
    std::vector<uint32_t> target_subset;
    if (target_name.empty()) {
      snprintf(line, sizeof(line), "Computing RMSD vs avg conformation using %d atoms from \"%s\".\n", (int)subset.size(), selection.c_str());
      std::cerr << line;
    } else {
      target = readPDB(target_name);
      target_subset = selectAtoms(target, target_selection);
      if (target_subset.size() != subset.size()) {
        snprintf(line, sizeof(line), "Error- target selection has %u atoms while trajectory selection has %u.\n", (unsigned)target_subset.size(), (unsigned)subset.size());
        std::cerr << line;
        return 255;
      }
      snprintf(line, sizeof(line), "Computing RMSD vs target %s using %d atoms from \"%s\".\n", target_name.c_str(), (int)subset.size(), selection.c_str());
      std::cerr << line;
    }

    Frames fr, fa;
    std::vector<double> rmsds(F, 0.0);
    const bool do_align = !alignment.empty();
    std::vector<uint32_t> align_subset;
    const std::vector<uint32_t> sel_r = selectionToFrameIndices(molecule, subset);
    std::vector<uint32_t> sel_a;
    std::vector<float> buf;  // only the --align '' / no-target corner needs the frames on the host
    auto reader = [&](size_t j, float* dst) { traj.readFramePlanes(indices[j], dst); };
    if (do_align) {
      align_subset = selectAtoms(molecule, alignment);
      if (align_subset.empty()) throw Error("Error- --align selection matched no atoms");
      sel_a = selectionToFrameIndices(molecule, align_subset);
      stageStreaming({&fr, &fa}, {&sel_r, &sel_a}, F, molecule.size(), device, reader);
    } else {
      stageStreaming({&fr}, {&sel_r}, F, molecule.size(), device, reader, target_name.empty() ? &buf : nullptr);
    }
    if (target_name.empty()) {
      if (do_align) {
        snprintf(line, sizeof(line), "Aligning using %d atoms from \"%s\".\n", (int)align_subset.size(), alignment.c_str());
        std::cerr << line;
        std::cerr << "Computing using average structure...\n";
        int iters = 0; double rms = 0.0;
        check(loosb200_rmsd2ref_iterative(fa.h, fr.h, o.dbl("tolerance"), 1000, rmsds.data(), nullptr, nullptr, &iters, &rms),
              "iterative alignment");
      } else {
        // --align '': the reference indexes an empty transform list here (undefined
        // behaviour); identity transforms are the evident intent.  The
        // average of the untransformed frames is the reference structure.
        std::cerr << "Computing using average structure...\n";
        std::vector<double> avg(3 * subset.size(), 0.0);
        const size_t na = molecule.size();
        for (size_t f = 0; f < F; ++f)
          for (size_t i = 0; i < subset.size(); ++i)
            for (int k = 0; k < 3; ++k) avg[3 * i + k] += (double)buf[(f * 3 + k) * na + molecule[subset[i]].index];
        for (double& v : avg) v /= (double)F;
        check(loosb200_rmsd2ref(fr.h, nullptr, nullptr, avg.data(), rmsds.data(), nullptr), "rmsd2ref");
      }
    } else {
      const std::vector<double> tr = coordsOf(target, target_subset);
      if (do_align) {
        const std::vector<uint32_t> ta_sel = selectAtoms(target, target_align);
        snprintf(line, sizeof(line), "Aligning using %d atoms from \"%s\".\n", (int)ta_sel.size(), alignment.c_str());
        std::cerr << line;
        if (ta_sel.size() != align_subset.size()) throw Error("Cannot superimpose AtomicGroups with differing numbers of atoms");
        const std::vector<double> ta = coordsOf(target, ta_sel);
        check(loosb200_rmsd2ref(fa.h, fr.h, ta.data(), tr.data(), rmsds.data(), nullptr), "rmsd2ref");
      } else {
        check(loosb200_rmsd2ref(fr.h, nullptr, nullptr, tr.data(), rmsds.data(), nullptr), "rmsd2ref");
      }
    }

    double avg_rmsd = 0.0;  // Tools/rmsd2ref.cpp:246-258
    for (double d : rmsds) avg_rmsd += d;
    avg_rmsd /= (double)F;
    double std_rmsd = 0.0;
    for (double d : rmsds) std_rmsd += (d - avg_rmsd) * (d - avg_rmsd);
    std_rmsd = sqrt(std_rmsd / (double)(F - 1));
    snprintf(line, sizeof(line), "Average RMSD was %.3lf, std RMSD was %.3lf\n", avg_rmsd, std_rmsd);
    std::cerr << line;
    for (size_t i = 0; i < F; ++i) std::cout << i << "\t" << formatG(rmsds[i], 6) << "\n";
    std::cout.flush();
  } catch (const std::exception& e) {
    std::cerr << e.what() << std::endl;
    return 255;
  }
  return 0;
}
