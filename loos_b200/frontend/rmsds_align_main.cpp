// rmsds-align -- pairwise RMSD over one selection after aligning on another, for one or two
// trajectories.  Command line and output as Packages/Clustering/rmsds-align.py (:118-256):
//   stdout: "# <command line>", then one line per frame of the first trajectory with
//   `round(rmsd, precision)` as Python prints a float, each value followed by a space.
// The script's quirks are kept, because a drop-in has to print what the script prints:
//   * a two-part range "skip:stride" stops at len(args.traj) -- the length of the trajectory's file
//     NAME, not its frame count (:188,202); give "skip:stride:stop" to choose the frames.  A warning
//     says so on stderr.
//   * the form of --range2 is decided by the number of parts of --range1 (:201-206).
//   * "Error in range" goes to stdout and the exit status is 0 (:193-195).
// argparse's prefix abbreviations of long options (--prec for --precision) are not accepted.
#include "tool_common.hpp"

using namespace lbfe;

namespace {

std::vector<std::string> splitColon(const std::string& s) {
  std::vector<std::string> parts;
  size_t a = 0;
  for (;;) {
    const size_t b = s.find(':', a);
    parts.push_back(s.substr(a, b == std::string::npos ? std::string::npos : b - a));
    if (b == std::string::npos) break;
    a = b + 1;
  }
  return parts;
}

long pyInt(const std::string& s) {  // int(): optional sign, digits, surrounding blanks
  size_t pos = 0;
  long v = 0;
  try { v = std::stol(s, &pos); } catch (...) { throw Error("ValueError: invalid literal for int() with base 10: '" + s + "'"); }
  while (pos < s.size() && isspace((unsigned char)s[pos])) ++pos;
  if (pos != s.size()) throw Error("ValueError: invalid literal for int() with base 10: '" + s + "'");
  return v;
}

std::vector<long> pyRange(long start, long stop, long step) {
  if (step == 0) throw Error("ValueError: range() arg 3 must not be zero");
  std::vector<long> r;
  if (step > 0) for (long i = start; i < stop; i += step) r.push_back(i);
  else for (long i = start; i > stop; i += step) r.push_back(i);
  return r;
}

// select_1 / select_2 (:186-206).  `nparts_decides` is the part count of --range1 in both calls.
bool frameList(const std::vector<std::string>& parts, size_t nparts_decides, const std::string& traj_name,
               std::vector<long>& out, bool* open_ended) {
  *open_ended = false;
  if (nparts_decides == 2) {
    if (parts.size() < 2) throw Error("IndexError: list index out of range");
    out = pyRange(pyInt(parts[0]), (long)traj_name.size(), pyInt(parts[1]));
    *open_ended = true;
  } else if (nparts_decides == 3) {
    if (parts.size() < 3) throw Error("IndexError: list index out of range");
    out = pyRange(pyInt(parts[0]), pyInt(parts[2]), pyInt(parts[1]));
  } else {
    return false;
  }
  return true;
}

std::vector<uint32_t> checkedFrames(const std::vector<long>& list, const Trajectory& traj, const std::string& name) {
  std::vector<uint32_t> r;
  for (long i : list) {
    // Trajectory::readFrame -> seekFrameImpl throws beyond the end (src/dcd.cpp:329-336)
    if (i < 0 || (unsigned long)i >= traj.nframes()) throw Error("Error- " + name + ": requested frame " + std::to_string(i) + " is out of range (the trajectory has " + std::to_string(traj.nframes()) + " frames)");
    r.push_back((uint32_t)i);
  }
  return r;
}

}  // namespace

int main(int argc, char* argv[]) {
  std::string header;  // " ".join(sys.argv)
  for (int i = 0; i < argc; ++i) header += (i ? " " : "") + std::string(argv[i]);
  Options o;
  o.add("help", 'h', "", "show this help message and exit", true);
  o.add("fullhelp", 0, "", "Print detailed description of all options", true);
  o.add("model2", 0, "", "Model for 2nd trajectory (default = model)");
  o.add("traj2", 0, "", "2nd trajectory (default = traj)");
  o.add("align", 0, "name == \"CA\"", "Align selection for 1st trajectory (default = CA)");
  o.add("align2", 0, "", "Align selection 2nd trajectory (default = align)");
  o.add("rmsd", 0, "name =~ \"^(C|O|N|CA)$\"", "RMSD selection for 1st trajectory(default = (C|O|N|CA))");
  o.add("rmsd2", 0, "", "RMSD selection for 2nd trajectory (default = rmsd)");
  o.add("precision", 0, "2", "Write out matrix coefficients with this many digits (default = 2)");
  o.add("range1", 0, "0:1", "Range of frames to use from the first trajectory [FORMAT - skip:stride:stop, stop is optional] (default = 0:1:last)");
  o.add("range2", 0, "", "Range of frames to use from the second trajectory [FORMAT - skip:stride:stop, stop is optional] (default = range1)");
  o.add("gpu", 0, "0", "CUDA device to use (extension; the script has no such option)");
  o.add("engine", 0, "auto", "auto|tensor|fp64: tensor = int8 tensor cores on a 24-bit fixed-point grid (values within ~1e-6 A of the script's), "
                             "fp64 = double-precision contractions (within 1e-8 A); auto = tensor for selections of up to 100,000 atoms (extension)");
  o.addPositional("model", 1); o.addPositional("traj", 1);
  try {
    if (!o.parse(argc, argv)) return 2;  // argparse exits with status 2 on a usage error
    if (o.has("help") || o.has("fullhelp")) { std::cout << o.usage(argv[0], "model traj"); return 0; }
    if (!o.has("model") || !o.has("traj")) {
      std::cerr << o.usage(argv[0], "model traj") << argv[0] << ": error: the following arguments are required: model, traj\n";
      return 2;
    }
    const std::string model_name = o.str("model"), traj_name = o.str("traj");
    const std::string model2_name = o.has("model2") ? o.str("model2") : model_name;
    const std::string traj2_name = o.has("traj2") ? o.str("traj2") : traj_name;
    const std::string align_sel = o.str("align"), rmsd_sel = o.str("rmsd");
    const std::string align2_sel = o.has("align2") ? o.str("align2") : align_sel;
    const std::string rmsd2_sel = o.has("rmsd2") ? o.str("rmsd2") : rmsd_sel;
    const std::string range1 = o.str("range1"), range2 = o.has("range2") ? o.str("range2") : range1;
    const int precision = (int)pyInt(o.str("precision"));
    const int device = (int)o.uns("gpu"), engine = engineOf(o);

    const std::vector<std::string> parts1 = splitColon(range1), parts2 = splitColon(range2);
    std::vector<long> list1, list2;
    bool open1 = false, open2 = false;
    if (!frameList(parts1, parts1.size(), traj_name, list1, &open1) || !frameList(parts2, parts1.size(), traj2_name, list2, &open2)) {
      std::cout << "Error in range" << std::endl;
      return 0;
    }

    Model model = readPDB(model_name);
    const std::shared_ptr<Trajectory> traj = openTrajectory(traj_name, (uint32_t)model.size());
    if (traj->natoms() != model.size()) throw Error("Error- trajectory '" + traj_name + "' does not match the model size");
    Model model2 = readPDB(model2_name);
    const std::shared_ptr<Trajectory> traj2 = openTrajectory(traj2_name, (uint32_t)model2.size());
    if (traj2->natoms() != model2.size()) throw Error("Error- trajectory '" + traj2_name + "' does not match the model size");
    if (open1 && traj_name.size() != traj->nframes())
      std::cerr << "Warning: as in rmsds-align.py, the open-ended --range1 '" << range1 << "' stops at frame " << traj_name.size()
                << " (the length of the trajectory's file name); the trajectory has " << traj->nframes()
                << " frames -- give skip:stride:stop to choose them\n";
    if (open2 && traj2_name.size() != traj2->nframes() && (traj2_name != traj_name || !open1))
      std::cerr << "Warning: as in rmsds-align.py, the open-ended --range2 stops at frame " << traj2_name.size()
                << " (the length of the trajectory's file name); the trajectory has " << traj2->nframes() << " frames\n";
    const std::vector<uint32_t> frames1 = checkedFrames(list1, *traj, traj_name), frames2 = checkedFrames(list2, *traj2, traj2_name);

    const std::vector<uint32_t> a1 = selectAtoms(model, align_sel), r1 = selectAtoms(model, rmsd_sel);
    const std::vector<uint32_t> a2 = selectAtoms(model2, align2_sel), r2 = selectAtoms(model2, rmsd2_sel);

    // the script prints the header before its loop; here it goes out once the result exists, so
    // that a failure leaves stdout empty (same bytes on success)
    const size_t F1 = frames1.size(), F2 = frames2.size();
    if (F1 == 0 || F2 == 0) {  // the inner loop never runs: one empty line per frame of trajectory 1
      std::cout << "# " << header << std::endl;
      for (size_t i = 0; i < F1; ++i) std::cout << "\n";
      return 0;
    }
    // AtomicGroup::superposition / ::rmsd (src/AG_linalg.cpp:189-192, src/AG_numerical.cpp:303-304)
    if (a1.size() != a2.size()) throw Error("Cannot superimpose groups with differing numbers of atoms");
    if (r1.size() != r2.size()) throw Error("Cannot compute RMSD between groups with different sizes");
    if (a1.empty() || r1.empty()) throw Error("Error- a selection matched no atoms");

    const std::vector<uint32_t> ia1 = selectionToFrameIndices(model, a1), ir1 = selectionToFrameIndices(model, r1);
    const std::vector<uint32_t> ia2 = selectionToFrameIndices(model2, a2), ir2 = selectionToFrameIndices(model2, r2);
    Frames fa1, fr1, fa2, fr2;
    stageStreaming({&fa1, &fr1}, {&ia1, &ir1}, F1, model.size(), device, [&](size_t j, float* dst) { traj->readFramePlanes(frames1[j], dst); });
    stageStreaming({&fa2, &fr2}, {&ia2, &ir2}, F2, model2.size(), device, [&](size_t j, float* dst) { traj2->readFramePlanes(frames2[j], dst); });

    std::vector<double> M(F1 * F2);
    check(loosb200_rmsds_align_engine(fa1.h, fr1.h, fa2.h, fr2.h, M.data(), F1, engine), "rmsds-align");

    std::cout << "# " << header << std::endl;
    writePyRoundMatrix(M.data(), F1, F2, F1, precision, readerThreads(),
                       [](const char* data, size_t n) { if (fwrite(data, 1, n, stdout) != n) throw Error("Error- cannot write to stdout"); });
    fflush(stdout);
  } catch (const std::exception& e) {
    std::cerr << e.what() << std::endl;
    return 1;  // an uncaught Python exception exits with status 1
  }
  return 0;
}
