// TEST INFRASTRUCTURE ONLY (oracle/): the two predicates src/KernelActions.cpp takes from src/Selectors.hpp
// (BackboneSelector :53-62, HydrogenSelector :121-123); their bodies are the reference's, extracted from
// src/Selectors.cpp:37-132,156-164 by oracle/Makefile.  The real header stands on AtomicGroup.
#if !defined(LOOS_SELECTORS_HPP)
#define LOOS_SELECTORS_HPP
#include <loos_defs.hpp>
#include <Atom.hpp>
namespace loos {
struct AtomSelector { virtual ~AtomSelector() {} };
class BackboneSelector : public AtomSelector {
  static const uint nresnames = 48;
  static std::string residue_names[nresnames];
  static const uint natomnames = 33;
  static std::string atom_names[natomnames];
 public:
  bool operator()(const pAtom&) const;
};
struct HydrogenSelector : public AtomSelector { bool operator()(const pAtom&) const; };
}  // namespace loos
#endif
