// Native front-end for the three RMSD tools: just enough of LOOS's L1/L2/L4 layers
// (PDB model, DCD trajectory, selection language, frame ranges, MultiTrajectory
// indexing, Matrix text output, option parsing) to drive the C ABI in
// include/loosb200.h from the same command lines, with bit-exact integer/text
// semantics.  LOOS itself cannot be linked here (Boost/NetCDF/HDF5 are absent), see
// DESIGN.md; in a LOOS build these pieces are LOOS's own.
// Every function cites the reference code whose behaviour it reproduces.
#pragma once
#include <stdint.h>

#include <fstream>
#include <functional>
#include <iosfwd>
#include <map>
#include <memory>
#include <stdexcept>
#include <string>
#include <vector>

namespace lbfe {

struct Error : std::runtime_error {
  explicit Error(const std::string& s) : std::runtime_error(s) {}
};

// ---- model (src/Atom.hpp, src/pdb.cpp:66-141) ----
struct Atom {
  uint32_t index = 0;  // position among the atom records of the file (pdb.cpp:73)
  long id = 0, resid = 0;
  std::string record, name, altloc, resname, chain, icode, element;
  std::string segid = "    ";  // Atom::init (src/Atom.cpp:189-207): what a line too short for the field keeps
  double x = 0, y = 0, z = 0;
  double occupancy = 1.0, bfactor = 0.0;  // defaults of Atom::init as well
};
typedef std::vector<Atom> Model;

int parseHybrid36(const std::string& src, unsigned pos, unsigned n);  // src/utils.cpp:259-316
Model readPDB(const std::string& path);                               // src/pdb.cpp:342-380
Model readPDB(std::istream& is);

// ---- selection language (src/grammar.yy, src/scanner.ll, src/KernelActions.cpp) ----
// Returns the indices INTO `model` (ascending: AtomicGroup::select keeps model
// order, src/AtomicGroup.cpp:341-351) of the atoms the expression accepts.
std::vector<uint32_t> selectAtoms(const Model& model, const std::string& selection);  // src/utils.cpp:183-201
// Atom::index() of the selected atoms (what updateGroupCoords uses as the frame offset).
std::vector<uint32_t> selectionToFrameIndices(const Model& model, const std::vector<uint32_t>& picked);

// ---- frame lists (src/index_range_parser.cpp:60-173, src/utils_structural.cpp:113-123) ----
std::vector<uint32_t> parseIndexRange(const std::string& spec, uint32_t maxsize);
std::vector<uint32_t> assignTrajectoryFrames(uint32_t nframes, const std::string& spec, uint32_t skip,
                                             uint32_t stride = 1);
std::vector<uint32_t> uniquify(const std::vector<uint32_t>& v);  // src/utils.hpp:427-436

// ---- trajectories ----
// What the tools need from a trajectory (src/Trajectory.hpp: natoms, nframes, readFrame +
// updateGroupCoords): frame `i` as three planes X[natoms] Y[natoms] Z[natoms].
class Trajectory {
 public:
  virtual ~Trajectory() {}
  virtual uint32_t natoms() const = 0;
  virtual uint32_t nframes() const = 0;
  virtual const std::string& filename() const = 0;
  // out holds 3*natoms floats.  Safe to call from several threads at once (frames are fetched with
  // positional reads into per-thread scratch): the staging loop decodes a batch of frames in parallel.
  virtual void readFramePlanes(uint32_t i, float* out) = 0;
};

// A file opened for positional reads (pread): no shared file position, no shared buffer.
class PosFile {
 public:
  PosFile() {}
  ~PosFile();
  PosFile(const PosFile&) = delete;
  PosFile& operator=(const PosFile&) = delete;
  void open(const std::string& path);  // throws Error
  uint64_t size() const { return size_; }
  // exactly n bytes from offset `off`, or throws Error(what)
  void readAt(uint64_t off, void* dst, size_t n, const std::string& what) const;

 private:
  int fd_ = -1;
  uint64_t size_ = 0;
};
// createTrajectory (src/sfactories.cpp:133-148): by suffix -- dcd; nc / netcdf (Amber NetCDF);
// crd / mdcrd when the file is NetCDF (the reference's text Amber format is not read here).
// `natoms` is the model's size; Amber NetCDF insists on a match (src/amber_netcdf.cpp:44-45).
std::shared_ptr<Trajectory> openTrajectory(const std::string& path, uint32_t natoms);

// ---- DCD (src/dcd.cpp:109-261,329-371) ----
class DCD : public Trajectory {
 public:
  explicit DCD(const std::string& path);
  uint32_t natoms() const override { return natoms_; }
  uint32_t nframes() const override { return nframes_; }
  bool hasCrystal() const { return icntrl_[10] == 1; }
  float timestep() const { return delta_; }
  const std::string& filename() const override { return path_; }
  // Frame `i` as three planes X[natoms] Y[natoms] Z[natoms] (the record order of the
  // file): out must hold 3*natoms floats.
  void readFramePlanes(uint32_t i, float* out) override;

 private:
  uint32_t recordLen();
  std::string path_;
  std::ifstream ifs_;  // header only; frames go through pf_
  PosFile pf_;
  bool swab_ = false;
  int icntrl_[20];
  float delta_ = 0;
  uint32_t natoms_ = 0, nframes_ = 0;
  uint64_t first_frame_pos_ = 0, frame_size_ = 0;
};

// ---- Amber NetCDF trajectories (src/amber_netcdf.cpp:20-130) ----
// The reference links libnetcdf; here the NetCDF-3 container (classic and 64-bit-offset: magic
// "CDF\1" / "CDF\2", big-endian header of dimensions, attributes and variables, fixed-size records)
// is parsed directly.  Same checks as AmberNetcdf::init: global attributes Conventions (must contain
// "AMBER") and ConventionVersion ("1.0"), dimension "atom" equal to the model size, dimension
// "frame", variable "coordinates" (frame, atom, spatial) as float or double.  NetCDF-4 (HDF5) files
// are refused.
class AmberNetcdf : public Trajectory {
 public:
  AmberNetcdf(const std::string& path, uint32_t natoms_expected);
  uint32_t natoms() const override { return natoms_; }
  uint32_t nframes() const override { return nframes_; }
  const std::string& filename() const override { return path_; }
  bool hasPeriodicBox() const { return periodic_; }
  void readFramePlanes(uint32_t i, float* out) override;
  static bool isNetcdf(const std::string& path);  // isFileNetCDF, src/amber_netcdf.hpp:22

 private:
  std::string path_;
  std::ifstream ifs_;  // header only; frames go through pf_
  PosFile pf_;
  uint32_t natoms_ = 0, nframes_ = 0;
  bool periodic_ = false, coords_double_ = false;
  uint64_t coord_begin_ = 0, rec_size_ = 0, coord_vsize_ = 0;
};

// ---- GROMACS XTC trajectories (src/xtc.cpp:165-486; format of the xdrfile library) ----
// XDR (big-endian) frames: magic 1995, natoms, step, time, box[9], natoms again; up to 9 atoms as plain
// floats, otherwise precision, integer bounding box, the "small" table index and a bit stream in which
// every atom is either three integers packed against the bounding box or a short run of small
// differences to its predecessor (with the first two atoms of a run swapped: water molecules).
// Coordinates come out as the reference computes them: (float)int * (1 / precision) in float, times
// 10.0 in double (nm -> Angstrom).  The frame index is built by one scan over the file (scanFrames).
class XTC : public Trajectory {
 public:
  explicit XTC(const std::string& path);
  uint32_t natoms() const override { return natoms_; }
  uint32_t nframes() const override { return (uint32_t)frame_pos_.size() - 1; }
  const std::string& filename() const override { return path_; }
  void readFramePlanes(uint32_t i, float* out) override;

 private:
  std::string path_;
  std::ifstream ifs_;  // frame scan only; frames go through pf_
  PosFile pf_;
  uint32_t natoms_ = 0;
  std::vector<uint64_t> frame_pos_;  // start of every frame, plus the end of the last one
};

// ---- MDTraj HDF5 trajectories (src/mdtrajtraj.cpp:32-142) ----
// Root-level dataset "coordinates" (frames x atoms x 3 floats, nm -> x 10.0), optional "cell_lengths".
// The reference links libhdf5; here the subset of the HDF5 file format that PyTables / MDTraj / h5py
// write for such files is parsed directly (hdf5_mini.cpp lists it).  Parity unpinned -- see there.
class MDTrajH5 : public Trajectory {
 public:
  MDTrajH5(const std::string& path, uint32_t natoms_expected);
  ~MDTrajH5() override;
  uint32_t natoms() const override { return natoms_; }
  uint32_t nframes() const override { return nframes_; }
  const std::string& filename() const override { return path_; }
  bool hasPeriodicBox() const { return periodic_; }
  void readFramePlanes(uint32_t i, float* out) override;

 private:
  struct Impl;
  std::string path_;
  std::unique_ptr<Impl> impl_;
  uint32_t natoms_ = 0, nframes_ = 0;
  bool periodic_ = false;
};

// ---- MultiTrajectory indexing (src/MultiTraj.hpp:98-105, src/MultiTraj.cpp:48-58) ----
struct MultiTraj {
  std::vector<std::shared_ptr<Trajectory>> trajs;
  uint32_t skip = 0, stride = 1;
  uint32_t nframes(size_t i) const;
  uint32_t nframes() const;
  // composite frame -> (trajectory, raw frame)
  std::pair<uint32_t, uint32_t> location(uint32_t k) const;
};

// ---- output (src/MatrixImpl.hpp:210-222, src/utils.cpp:119-164) ----
// M: float32 column-major, ld >= m.  Same bytes as `os << setprecision(p) << M`.
void writeMatrix(std::ostream& os, const float* M, size_t m, size_t n, size_t ld, unsigned precision);
// readAsciiMatrix (src/MatrixRead.hpp:102-143): skip lines until one scans as "# m n", then m*n
// whitespace-separated values, row by row.  Returns column-major float32 (ld = m), i.e. exactly
// what writeMatrix was given.  Errors carry the reference's messages.
std::vector<float> readMatrix(std::istream& is, size_t* m, size_t* n);
// Binary side channel (extension): the same column-major float32 matrix as a NumPy
// .npy file (format 1.0, fortran_order True) -- 4 bytes per value instead of ~5 of text and no
// formatting cost; numpy.load gives M[i, j].
void writeNpy(const std::string& path, const float* M, size_t m, size_t n, size_t ld);
std::string invocationHeader(int argc, char* argv[]);
// default-floatfield formatting of a double with `precision` significant digits
// (std::ostream default: 6), as rmsd2ref prints its values.
std::string formatG(double v, unsigned precision);
// "%.{precision}g" of a float from exact integer arithmetic (same bytes as printf; test hook)
int formatFloatG(float v, unsigned precision, char* out);

// ---- options (src/OptionsFramework.cpp:742-795: boost::program_options semantics used by
// the tools: long names without prefix guessing, short aliases, every option takes a
// value -- booleans too --, positionals in declaration order) ----
// The text Python prints for `round(v, ndigits)` of a float (float.__round__ = correctly rounded
// decimal at that position, converted back to double; repr = shortest digits that round-trip,
// exponent form when the decimal exponent is < -4 or >= 16).  Packages/Clustering/rmsds-align.py:254
// writes its matrix that way.  Appends to `out`.
void pyRoundRepr(double v, int ndigits, std::string& out);
// The body of rmsds-align.py's output (:254-256): row i = `round(M(i,j), precision)` + " " for every j, then a
// newline; M column-major with leading dimension ld.  Rows are formatted in blocks on `threads` threads and
// handed to `write` in order.
void writePyRoundMatrix(const double* M, size_t rows, size_t cols, size_t ld, int precision, unsigned threads,
                        const std::function<void(const char*, size_t)>& write);

class Options {
 public:
  void add(const std::string& longname, char shortname, const std::string& defval, const std::string& help,
           bool is_bool = false);
  void addPositional(const std::string& name, int max_count);  // -1: unlimited
  // returns false (after printing to stderr) on a syntax error
  bool parse(int argc, char* argv[]);
  bool has(const std::string& name) const;
  std::string str(const std::string& name) const;
  unsigned long uns(const std::string& name) const;
  double dbl(const std::string& name) const;
  bool flag(const std::string& name) const;
  const std::vector<std::string>& multi(const std::string& name) const;
  std::string usage(const std::string& prog, const std::string& positional_help) const;

 private:
  struct Opt { std::string name, defval, help; char shortname; bool is_bool, set; std::vector<std::string> values; };
  std::vector<Opt> opts_;
  std::vector<std::pair<std::string, int>> positional_;
  Opt* find(const std::string& name);
  const Opt* find(const std::string& name) const;
};

}  // namespace lbfe
