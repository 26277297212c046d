// TEST INFRASTRUCTURE ONLY (oracle/).  Never linked into, imported by or called from the product path.
//
// The REFERENCE'S OWN PDB atom-record parser and MultiTrajectory indexing, compiled from the sources where they lie:
//   * src/pdb.cpp:65-156          PDB::parseAtomRecord
//   * src/utils.hpp:323-359       template parseStringAs<T>;  src/utils.cpp:228-242 its std::string specialisation
//   * src/utils.cpp:253-317       parseStringAsHybrid36 (with its tables)
//   * src/utils.hpp:426-436       uniquifyVector
//   * src/MultiTraj.hpp:97-105    MultiTrajectory::nframes(i);  src/MultiTraj.cpp:48-58 frameIndexToLocation
//   * src/Selectors.cpp:37-132,156-164  BackboneSelector (its two name tables) and HydrogenSelector
//   * src/index_range_parser.cpp:63-102 struct RangeItem: validate() and generate() of one range of the frame-range
//                                 grammar (the Boost.Spirit grammar that builds RangeItems cannot be compiled here)
// oracle/Makefile extracts the line ranges into a throw-away directory; nothing of them stays in the repo.
// Below: shells for Atom / PDB / MultiTrajectory with just the members those functions touch (the real classes,
// src/Atom.hpp, src/pdb.hpp, src/MultiTraj.hpp, stand on AtomicGroup / Trajectory / Boost), and a C interface.
#include <loos_defs.hpp>  // oracle/shim
#include <Coord.hpp>      // /root/reference/src

#include <algorithm>
#include <cstring>
#include <iterator>
#include <map>
#include <memory>
#include <set>
#include <sstream>
#include <stdexcept>
#include <string>
#include <vector>

namespace loos {

struct LOOSError : public std::runtime_error { explicit LOOSError(const std::string& s) : std::runtime_error(s) {} };
struct ParseError : public LOOSError { explicit ParseError(const std::string& s) : LOOSError(s) {} };

#include "_ref/parse_string_as.inc"      // template <typename T> T parseStringAs(...)
template <> std::string parseStringAs<std::string>(const std::string& source, const uint pos, const uint nelem);
int parseStringAsHybrid36(const std::string& source, const uint pos = 0, const uint nelem = 0);
#include "_ref/parse_string_spec.inc"    // the std::string specialisation
#include "_ref/hybrid36.inc"             // tables + parseStringAsHybrid36
#include "_ref/uniquify.inc"             // uniquifyVector

class Atom {  // the setters parseAtomRecord calls (src/Atom.hpp)
 public:
  void index(uint i) { index_ = i; }
  void recordName(const std::string& s) { record = s; }
  void id(int i) { id_ = i; }
  int id() const { return id_; }
  void name(const std::string& s) { name_ = s; }
  void altLoc(const std::string& s) { altloc = s; }
  void resname(const std::string& s) { resname_ = s; }
  void chainId(const std::string& s) { chain = s; }
  void resid(int i) { resid_ = i; }
  void iCode(const std::string& s) { icode = s; }
  void coords(const GCoord& c) { xyz = c; }
  void occupancy(greal r) { occ = r; }
  void bfactor(greal r) { b = r; }
  void segid(const std::string& s) { segid_ = s; }
  void PDBelement(const std::string& s) { element = s; }
  // what the selectors read (src/Atom.hpp)
  enum bits { nullbit = 0, massbit = 1 };
  std::string name() const { return name_; }
  std::string resname() const { return resname_; }
  bool checkProperty(bits) const { return has_mass; }
  double mass() const { return mass_; }
  bool has_mass = false;
  double mass_ = 1.0;
  // defaults: Atom::init, src/Atom.cpp:189-207
  uint index_ = 0;
  int id_ = 1, resid_ = 1;
  std::string record = "ATOM", name_ = "    ", altloc = " ", resname_ = "   ", chain = " ", icode, segid_ = "    ", element;
  GCoord xyz;
  greal occ = 1.0, b = 0.0;
};
typedef std::shared_ptr<Atom> pAtom;

class PDB {
 public:
  PDB() : _max_index(0), strictness_policy(false), _missing_q(false), _missing_b(false), _missing_segid(false) {}
  void parseAtomRecord(const std::string& s);
  void append(pAtom pa) { atoms.push_back(pa); }
  uint _max_index;
  bool strictness_policy, _missing_q, _missing_b, _missing_segid;
  std::vector<pAtom> atoms;
  std::map<int, pAtom> _atomid_to_patom;
};

#include "_ref/pdb_atom_record.inc"      // PDB::parseAtomRecord

struct AtomSelector {};
class BackboneSelector : public AtomSelector {  // src/Selectors.hpp:53-62
  static const uint nresnames = 48;
  static std::string residue_names[nresnames];
  static const uint natomnames = 33;
  static std::string atom_names[natomnames];
 public:
  bool operator()(const pAtom&) const;
};
struct HydrogenSelector : public AtomSelector { bool operator()(const pAtom&) const; };
#include "_ref/selectors_backbone.inc"   // the tables and BackboneSelector::operator()
#include "_ref/selectors_hydrogen.inc"   // HydrogenSelector::operator()

struct FakeTraj { uint n; uint nframes() const { return n; } };
class MultiTrajectory {
 public:
  typedef std::pair<uint, uint> Location;  // src/MultiTraj.hpp
  std::vector<std::shared_ptr<FakeTraj>> _trajectories;
  uint _skip, _stride;
#include "_ref/multitraj_nframes.inc"      // uint nframes(const uint i) const
  Location frameIndexToLocation(const uint i);
};
#include "_ref/multitraj_location.inc"   // MultiTrajectory::frameIndexToLocation

}  // namespace loos

namespace loos { namespace internal {
using namespace std;
#include "_ref/range_item.inc"           // struct RangeItem
} }

extern "C" {

// One ATOM/HETATM line -> the fields parseAtomRecord stores.  ints: index? no -- id, resid; strs (each up to 8 chars,
// NUL-terminated, 16 bytes apart): record, name, altloc, resname, chain, icode, segid, element; reals: x, y, z, occ, b.
int ref_pdb_atom(const char* line, int* ints2, char* strs8x16, double* reals5, char* err, int errcap) {
  try {
    loos::PDB p;
    p.parseAtomRecord(std::string(line));
    const loos::Atom& a = *p.atoms.at(0);
    ints2[0] = a.id_; ints2[1] = a.resid_;
    const std::string* f[8] = {&a.record, &a.name_, &a.altloc, &a.resname_, &a.chain, &a.icode, &a.segid_, &a.element};
    for (int k = 0; k < 8; ++k) { memset(strs8x16 + 16 * k, 0, 16); strncpy(strs8x16 + 16 * k, f[k]->c_str(), 15); }
    reals5[0] = a.xyz[0]; reals5[1] = a.xyz[1]; reals5[2] = a.xyz[2]; reals5[3] = a.occ; reals5[4] = a.b;
    return 0;
  } catch (const std::exception& e) {
    if (err && errcap > 0) { strncpy(err, e.what(), (size_t)errcap - 1); err[errcap - 1] = 0; }
    return -1;
  }
}

int ref_hybrid36(const char* field, int n) { return loos::parseStringAsHybrid36(std::string(field), 0, (unsigned)n); }

// MultiTrajectory: composite frame count and the (trajectory, frame) of composite frame k for k < cap
long ref_multitraj(const unsigned* sizes, int ntraj, unsigned skip, unsigned stride, unsigned* loc_traj, unsigned* loc_frame, long cap) {
  loos::MultiTrajectory mt;
  mt._skip = skip; mt._stride = stride;
  for (int i = 0; i < ntraj; ++i) mt._trajectories.push_back(std::make_shared<loos::FakeTraj>(loos::FakeTraj{sizes[i]}));
  long total = 0;
  for (int i = 0; i < ntraj; ++i) total += mt.nframes((unsigned)i);
  for (long k = 0; k < total && k < cap; ++k) {
    const loos::MultiTrajectory::Location loc = mt.frameIndexToLocation((unsigned)k);
    loc_traj[k] = loc.first; loc_frame[k] = loc.second;
  }
  return total;
}

int ref_backbone(const char* resname, const char* name) {
  loos::pAtom a(new loos::Atom);
  a->resname(resname); a->name(name);
  return loos::BackboneSelector()(a) ? 1 : 0;
}
int ref_hydrogen(const char* name, int has_mass, double mass) {
  loos::pAtom a(new loos::Atom);
  a->name(name); a->has_mass = has_mass != 0; a->mass_ = mass;
  return loos::HydrogenSelector()(a) ? 1 : 0;
}

// RangeItem(a) / RangeItem(a, b) / RangeItem(a, b, c) -> generate(): nargs picks the constructor.  Returns the number of
// values (written up to cap), or -1 when generate() throws ("Invalid range").
long ref_range_item(int nargs, unsigned a, unsigned b, int c, unsigned* out, long cap) {
  try {
    const loos::internal::RangeItem item = nargs == 1 ? loos::internal::RangeItem(a)
                                           : nargs == 2 ? loos::internal::RangeItem(a, b) : loos::internal::RangeItem(a, b, c);
    const std::vector<unsigned> v = item.generate();
    for (long k = 0; k < (long)v.size() && k < cap; ++k) out[k] = v[(size_t)k];
    return (long)v.size();
  } catch (const loos::ParseError&) {
    return -1;
  }
}

// uniquifyVector<uint>: returns the number of unique values written to out
long ref_uniquify(const unsigned* v, long n, unsigned* out) {
  const std::vector<unsigned> u = loos::uniquifyVector(std::vector<unsigned>(v, v + n));
  std::copy(u.begin(), u.end(), out);
  return (long)u.size();
}

}  // extern "C"
