// TEST INFRASTRUCTURE (oracle shim) -- not product code.
// Stand-in for dune-common's version macros (third party, absent).  Only the macro the reference's
// SimulatorUtilities.hpp:264 tests is provided; it selects between two spellings of the leaf grid view
// inside writeVtkOutput, which the oracle never instantiates.
#ifndef ORACLE_SHIM_DUNE_VERSION_HH
#define ORACLE_SHIM_DUNE_VERSION_HH
#define DUNE_VERSION_NEWER(module, major, minor) 1
#endif
