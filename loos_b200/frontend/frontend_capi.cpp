// Test hooks: a C view of the front-end pieces so that tests/ can check them
// bit-exactly against independent restatements (not part of include/loosb200.h).
#include <string.h>

#include <sstream>
#include <thread>

#include "frontend.hpp"

using namespace lbfe;

static thread_local std::string g_err;
static int fail(const std::exception& e) { g_err = e.what(); return -1; }

extern "C" {
const char* lbfe_last_error(void) { return g_err.c_str(); }

// indices into the model (and Atom::index, identical for PDB) of the selected atoms
long lbfe_select(const char* pdb_path, const char* selection, uint32_t* out, long cap) {
  try {
    Model m = readPDB(pdb_path);
    std::vector<uint32_t> s = selectionToFrameIndices(m, selectAtoms(m, selection));
    for (long i = 0; i < (long)s.size() && i < cap; ++i) out[i] = s[i];
    return (long)s.size();
  } catch (const std::exception& e) { return fail(e); }
}

long lbfe_pdb_atoms(const char* pdb_path, long* ids, long* resids, double* xyz, char* names /* 8 bytes each: name|resname */, long cap) {
  try {
    Model m = readPDB(pdb_path);
    for (long i = 0; i < (long)m.size() && i < cap; ++i) {
      ids[i] = m[i].id; resids[i] = m[i].resid;
      xyz[3 * i] = m[i].x; xyz[3 * i + 1] = m[i].y; xyz[3 * i + 2] = m[i].z;
      memset(names + 16 * i, 0, 16);
      strncpy(names + 16 * i, m[i].name.c_str(), 7);
      strncpy(names + 16 * i + 8, m[i].segid.c_str(), 7);
    }
    return (long)m.size();
  } catch (const std::exception& e) { return fail(e); }
}

// One ATOM / HETATM line through the PDB parser; same interface as the reference-side checker's ref_pdb_atom:
// ints: id, resid; strs (16 bytes apart): record, name, altloc, resname, chain, icode, segid, element; reals: x y z occ b
int lbfe_pdb_atom_line(const char* line, int* ints2, char* strs8x16, double* reals5) {
  try {
    std::istringstream is(std::string(line) + "\n");
    const Model m = readPDB(is);
    if (m.size() != 1) throw Error("not an ATOM/HETATM record");
    const Atom& a = m[0];
    ints2[0] = a.id; ints2[1] = a.resid;
    const std::string* f[8] = {&a.record, &a.name, &a.altloc, &a.resname, &a.chain, &a.icode, &a.segid, &a.element};
    for (int k = 0; k < 8; ++k) { memset(strs8x16 + 16 * k, 0, 16); strncpy(strs8x16 + 16 * k, f[k]->c_str(), 15); }
    reals5[0] = a.x; reals5[1] = a.y; reals5[2] = a.z; reals5[3] = a.occupancy; reals5[4] = a.bfactor;
    return 0;
  } catch (const std::exception& e) { return fail(e); }
}

int lbfe_hybrid36(const char* field, int n) { try { return parseHybrid36(std::string(field), 0, (unsigned)n); } catch (const std::exception& e) { fail(e); return -999999; } }

long lbfe_parse_range(const char* spec, uint32_t maxsize, uint32_t* out, long cap) {
  try {
    std::vector<uint32_t> r = parseIndexRange(spec, maxsize);
    for (long i = 0; i < (long)r.size() && i < cap; ++i) out[i] = r[i];
    return (long)r.size();
  } catch (const std::exception& e) { return fail(e); }
}

long lbfe_assign_frames(uint32_t nframes, const char* spec, uint32_t skip, uint32_t stride, int uniq, uint32_t* out, long cap) {
  try {
    std::vector<uint32_t> r = assignTrajectoryFrames(nframes, spec, skip, stride);
    if (uniq) r = uniquify(r);
    for (long i = 0; i < (long)r.size() && i < cap; ++i) out[i] = r[i];
    return (long)r.size();
  } catch (const std::exception& e) { return fail(e); }
}

// uniquifyVector (src/utils.hpp:426-436): sorted, duplicates removed
long lbfe_uniquify(const uint32_t* v, long n, uint32_t* out) {
  const std::vector<uint32_t> u = uniquify(std::vector<uint32_t>(v, v + n));
  for (size_t i = 0; i < u.size(); ++i) out[i] = u[i];
  return (long)u.size();
}

int lbfe_dcd_info(const char* path, uint32_t* natoms, uint32_t* nframes, int* crystal) {
  try { DCD d(path); *natoms = d.natoms(); *nframes = d.nframes(); *crystal = d.hasCrystal(); return 0; }
  catch (const std::exception& e) { return fail(e); }
}

int lbfe_dcd_read(const char* path, uint32_t frame, float* planes) {
  try { DCD d(path); d.readFramePlanes(frame, planes); return 0; }
  catch (const std::exception& e) { return fail(e); }
}

// any supported trajectory through openTrajectory (suffix dispatch)
// Python's repr(round(v, ndigits)) (the text of rmsds-align.py's matrix); returns the length or -1
int lbfe_py_round_repr(double v, int ndigits, char* out, int cap) {
  std::string s;
  pyRoundRepr(v, ndigits, s);
  if ((int)s.size() + 1 > cap) return -1;
  memcpy(out, s.c_str(), s.size() + 1);
  return (int)s.size();
}

// rmsds-align's matrix text for a column-major double matrix; returns the length or -1 when `cap` is too small
long lbfe_py_round_matrix(const double* M, long rows, long cols, long ld, int precision, int threads, char* out, long cap) {
  try {
    std::string all;
    writePyRoundMatrix(M, (size_t)rows, (size_t)cols, (size_t)ld, precision, (unsigned)threads,
                       [&](const char* d, size_t n) { all.append(d, n); });
    if ((long)all.size() > cap) return -1;
    memcpy(out, all.data(), all.size());
    return (long)all.size();
  } catch (const std::exception& e) { return fail(e); }
}

int lbfe_traj_info(const char* path, uint32_t natoms_expected, uint32_t* natoms, uint32_t* nframes) {
  try { auto t = openTrajectory(path, natoms_expected); *natoms = t->natoms(); *nframes = t->nframes(); return 0; }
  catch (const std::exception& e) { return fail(e); }
}

int lbfe_traj_read(const char* path, uint32_t natoms_expected, uint32_t frame, float* planes) {
  try { auto t = openTrajectory(path, natoms_expected); t->readFramePlanes(frame, planes); return 0; }
  catch (const std::exception& e) { return fail(e); }
}

// Frames frames[0..n) of one open trajectory decoded by `threads` threads at once (frame k by thread
// k % threads) into out[k][3][natoms]: what the staging loop of the tools does with a batch.
int lbfe_traj_read_many(const char* path, uint32_t natoms_expected, const uint32_t* frames, int n, int threads, float* out) {
  try {
    auto t = openTrajectory(path, natoms_expected);
    const size_t stride = 3 * (size_t)t->natoms();
    std::vector<std::thread> pool;
    std::vector<std::string> errs((size_t)threads);
    for (int w = 0; w < threads; ++w)
      pool.emplace_back([&, w]() {
        try {
          for (int k = w; k < n; k += threads) t->readFramePlanes(frames[k], out + (size_t)k * stride);
        } catch (const std::exception& e) { errs[(size_t)w] = e.what(); }
      });
    for (std::thread& th : pool) th.join();
    for (const std::string& e : errs) if (!e.empty()) throw Error(e);
    return 0;
  } catch (const std::exception& e) { return fail(e); }
}

// MultiTraj: sizes[] = frames in each sub-trajectory; returns composite count and the
// (traj, frame) location of composite frame k for every k < cap
long lbfe_multitraj(const uint32_t* sizes, int ntraj, uint32_t skip, uint32_t stride, uint32_t* loc_traj, uint32_t* loc_frame, long cap) {
  long total = 0;
  for (int i = 0; i < ntraj; ++i) total += sizes[i] <= skip ? 0 : (sizes[i] - skip + stride - 1) / stride;
  for (long k = 0; k < total && k < cap; ++k) {
    uint32_t t, j = 0;
    for (t = 0; t < (uint32_t)ntraj; ++t) {
      const uint32_t n = sizes[t] <= skip ? 0 : (sizes[t] - skip + stride - 1) / stride;
      if (j + n > k) break;
      j += n;
    }
    loc_traj[k] = t; loc_frame[k] = skip + ((uint32_t)k - j) * stride;
  }
  return total;
}

// text -> column-major floats (ld = m); returns 0 and the shape, or -1 (lbfe_last_error)
int lbfe_read_matrix(const char* text, long len, float* out, long cap, long* m, long* n) {
  try {
    std::istringstream is(std::string(text, (size_t)len));
    size_t mm = 0, nn = 0;
    const std::vector<float> M = readMatrix(is, &mm, &nn);
    *m = (long)mm; *n = (long)nn;
    if ((long)M.size() > cap) return -2;
    memcpy(out, M.data(), M.size() * sizeof(float));
    return 0;
  } catch (const std::exception& e) { return fail(e); }
}

// n floats -> their "%.{p}g" texts, '\n'-separated, into out (cap bytes); returns the length
long lbfe_format_floats(const float* v, size_t n, unsigned precision, char* out, long cap) {
  long len = 0;
  char buf[48];
  for (size_t i = 0; i < n; ++i) {
    const int l = formatFloatG(v[i], precision, buf);
    if (len + l + 1 > cap) return -1;
    memcpy(out + len, buf, (size_t)l);
    len += l;
    out[len++] = '\n';
  }
  return len;
}

long lbfe_matrix_text(const float* M, size_t m, size_t n, size_t ld, unsigned precision, char* out, long cap) {
  std::ostringstream os;
  writeMatrix(os, M, m, n, ld, precision);
  const std::string s = os.str();
  if ((long)s.size() < cap) memcpy(out, s.data(), s.size());
  return (long)s.size();
}
}
