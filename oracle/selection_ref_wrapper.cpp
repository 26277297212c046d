// TEST INFRASTRUCTURE ONLY (oracle/).  Never linked into, imported by or called from the product path.
//
// The REFERENCE'S OWN selection language, compiled from the sources where they lie under /root/reference/src:
// the flex scanner (scanner.cc, pre-generated from scanner.ll), the bison grammar (grammar.cc / grammar.hh /
// stack.hh / location.hh / position.hh, pre-generated from grammar.yy), Kernel.cpp, KernelActions.cpp,
// KernelValue.cpp, KernelStack.cpp, ParserDriver.hpp, LoosLexer.hpp, Atom.{hpp,cpp}, Coord.hpp, exceptions.hpp --
// all unmodified, passed to the compiler by oracle/Makefile -- plus Selectors.cpp:37-132,156-164 (backbone /
// hydrogen) extracted into a throw-away directory.  <FlexLexer.h> is the reference's own src/FlexLexer.h (on the
// include path where it lies).  Stand-ins (oracle/shim_sel): boost::regex over std::regex, loos_defs.hpp without
// Boost, Selectors.hpp without AtomicGroup.
// Restated here: selectAtoms (src/utils.cpp:183-201: Parser::parse, error prefix), AtomicGroup::select
// (src/AtomicGroup.cpp:341-351: atoms in order, those the predicate accepts) and KernelSelector::operator()
// (src/Selectors.cpp:189-200), because Parser.hpp / AtomicGroup stand on the rest of LOOS.
#include <loos_defs.hpp>     // oracle/shim_sel
#include <Atom.hpp>          // /root/reference/src
#include <Kernel.hpp>        // /root/reference/src
#include <ParserDriver.hpp>  // /root/reference/src
#include <Selectors.hpp>     // oracle/shim_sel

#include <algorithm>
#include <cstring>

namespace loos {
#include "_ref/selectors_backbone.inc"
#include "_ref/selectors_hydrogen.inc"
}  // namespace loos

extern "C" {

// Atoms: ids / resids [n]; strs: n x 4 fields of 8 bytes (name, resname, segid, chain id), NUL-terminated.
// mask[i] = 1 when atom i is selected.  Returns 0, or -1 with the exception text in err.
int ref_select(const char* selection, int n, const int* ids, const int* resids, const char* strs, unsigned char* mask,
               char* err, int errcap) {
  try {
    loos::Kernel krnl;
    loos::ParserDriver driver(krnl);
    try {
      krnl.clearActions();
      driver.parse(std::string(selection));
    } catch (loos::ParseError& e) {
      throw loos::ParseError("Error in parsing '" + std::string(selection) + "' ... " + e.what());
    }
    for (int i = 0; i < n; ++i) {
      loos::pAtom pa(new loos::Atom);
      pa->index((unsigned)i);
      pa->id(ids[i]);
      pa->resid(resids[i]);
      pa->name(std::string(strs + 32 * (size_t)i));
      pa->resname(std::string(strs + 32 * (size_t)i + 8));
      pa->segid(std::string(strs + 32 * (size_t)i + 16));
      pa->chainId(std::string(strs + 32 * (size_t)i + 24));
      krnl.execute(pa);
      if (krnl.stack().size() != 1) throw loos::LOOSError("Execution error - unexpected values on stack");
      loos::internal::Value results = krnl.stack().pop();
      if (results.type != loos::internal::Value::INT) throw loos::LOOSError("Execution error - unexpected value on top of stack");
      mask[i] = results.itg ? 1 : 0;
    }
    return 0;
  } catch (const std::exception& e) {
    if (err && errcap > 0) { strncpy(err, e.what(), (size_t)errcap - 1); err[errcap - 1] = 0; }
    return -1;
  }
}

}  // extern "C"
