// TEST INFRASTRUCTURE ONLY (oracle/).  Never linked into, imported by or called from the product path.
//
// The REFERENCE'S OWN XTC code, compiled from the sources where they lie under /root/reference:
//   * src/xtc.cpp:36-357,395-416   XTC::magicints ..., sizeofint(s), decodebits, decodeints,
//                                  readCompressedCoords, readUncompressedCoords, readFrameHeader
//   * src/xtcwriter.cpp:206-655,680-702  XTCWriter::magicints ..., encodebits, encodeints,
//                                  writeCompressedCoordsFloat, writeHeader, writeBox
//   * src/xdr.hpp (XDRReader / XDRWriter) and src/Coord.hpp, read in place via -I/root/reference/src;
//   * src/utils.hpp:301-318 (swab), through oracle/shim/utils.hpp.
// oracle/Makefile extracts the line ranges into a throw-away directory; nothing of them stays in the repo.
// Below: minimal class shells with the members those functions use (src/xtc.hpp:60-140,
// src/xtcwriter.hpp:48-175 declare the real ones, which drag in Trajectory / AtomicGroup / Boost), the
// frame loops around them restated from src/xtc.cpp:375-392 (parseFrame) and src/xtcwriter.cpp:661-677,
// 705-734 (allocateBuffers, writeFrame), and a C interface for tests.
#include <loos_defs.hpp>  // oracle/shim
#include <utils.hpp>      // oracle/shim
#include <Coord.hpp>      // /root/reference/src
#include <xdr.hpp>        // /root/reference/src

#include <climits>
#include <sstream>
#include <vector>

namespace loos {

class XTC {
 public:
  struct Header { uint magic; uint natoms, step; float time, box[9]; };  // src/xtc.hpp:68-72
  static const int magicints[];
  static const int firstidx, lastidx;
  static const int magic;
  static const uint min_compressed_system_size;
  typedef float xtc_t;  // src/xtc.hpp:82

  explicit XTC(std::istream* s) : xdr_file(s), natoms_(0), precision_(0), _filename("(memory)") {}
  static int sizeofint(int);
  static int sizeofints(uint*, const uint);
  static int decodebits(int*, uint);
  static void decodeints(int*, const int, int, uint*, int*);
  bool readCompressedCoords(void);
  bool readUncompressedCoords(void);
  bool readFrameHeader(Header&);

  internal::XDRReader xdr_file;
  uint natoms_;
  xtc_t precision_;
  std::string _filename;
  std::vector<GCoord> coords_;
};

class XTCWriter {
 public:
  static const int magicints[];
  static const int firstidx, lastidx, DIM;
  explicit XTCWriter(std::ostream* s) : buf1size(0), buf2size(0), buf1(0), buf2(0), _filename("(memory)") { xdr.setStream(s); }
  ~XTCWriter() { delete[] buf1; delete[] buf2; }
  int sizeofint(const int size) const;
  int sizeofints(const int num_of_bits, const unsigned int sizes[]) const;
  void encodebits(int* buf, int num_of_bits, const int num) const;
  void encodeints(int* buf, const int num_of_ints, const int num_of_bits, const unsigned int* sizes, const unsigned int* nums) const;
  void writeCompressedCoordsFloat(float* ptr, int size, float precision);
  void writeHeader(const int natoms, const int step, const float time);
  void writeBox(const GCoord& box);
  void allocateBuffers(const size_t size) {  // src/xtcwriter.cpp:661-677
    size_t size3 = size * 3;
    if (size3 > buf1size) {
      delete[] buf1;
      delete[] buf2;
      buf1 = new int[size3];
      size_t size3plus = size3 * 1.2;
      buf2 = new int[size3plus];
      buf1size = size3;
      buf2size = size3 * 1.2;
    }
  }
  size_t buf1size, buf2size;
  int *buf1, *buf2;
  internal::XDRWriter xdr;
  std::string _filename;
};

#include "_ref/xtc_reader_core.inc"  // verbatim reference lines, built by Makefile
#include "_ref/xtc_writer_core.inc"  // verbatim reference lines, built by Makefile

}  // namespace loos

extern "C" {

// frames: nframes x natoms x 3 doubles in Angstrom (GCoord is double); box in Angstrom.  The loop is
// writeFrame (src/xtcwriter.cpp:705-734): header, box, coordinates / 10.0 narrowed to float, compressed.
long ref_xtc_write(const double* xyz, int nframes, int natoms, float precision, const double* box3, double dt,
                   unsigned char* out, long cap) {
  try {
    std::ostringstream os(std::ios::binary);
    loos::XTCWriter w(&os);
    std::vector<float> crds((size_t)natoms * 3);
    for (int f = 0; f < nframes; ++f) {
      const unsigned step = (unsigned)f;  // step_ += steps_per_frame_ (1)
      w.writeHeader(natoms, (int)step, (float)(dt * step));
      w.writeBox(loos::GCoord(box3[0], box3[1], box3[2]));
      for (int i = 0, k = 0; i < natoms; ++i)
        for (int a = 0; a < 3; ++a) crds[k++] = xyz[((size_t)f * natoms + i) * 3 + a] / 10.0;  // Convert to nm
      w.writeCompressedCoordsFloat(crds.data(), natoms, precision);
    }
    const std::string s = os.str();
    if ((long)s.size() > cap) return -2;
    memcpy(out, s.data(), s.size());
    return (long)s.size();
  } catch (const std::exception&) { return -1; }
}

// Decodes every frame of an in-memory XTC file: out = nframes x natoms x 3 doubles (Angstrom), the GCoords
// parseFrame (src/xtc.cpp:375-392) leaves in coords_.  Returns the number of frames or a negative error.
int ref_xtc_read(const unsigned char* data, long len, double* out, int max_frames, int* natoms_out) {
  try {
    std::istringstream is(std::string((const char*)data, (size_t)len), std::ios::binary);
    loos::XTC x(&is);
    int nf = 0;
    for (;;) {
      loos::XTC::Header h;
      x.coords_.clear();
      if (!x.readFrameHeader(h)) break;
      if (x.natoms_ == 0) x.natoms_ = h.natoms;
      const bool ok = x.natoms_ <= loos::XTC::min_compressed_system_size ? x.readUncompressedCoords() : x.readCompressedCoords();
      if (!ok) break;
      if (nf >= max_frames) return -2;
      if (x.coords_.size() != x.natoms_) return -3;
      for (size_t i = 0; i < x.coords_.size(); ++i)
        for (int a = 0; a < 3; ++a) out[((size_t)nf * x.natoms_ + i) * 3 + a] = x.coords_[i][a];
      ++nf;
    }
    *natoms_out = (int)x.natoms_;
    return nf;
  } catch (const std::exception&) { return -1; }
}

}  // extern "C"
