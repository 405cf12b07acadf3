// TEST INFRASTRUCTURE (oracle shim) -- not product code.
// Stand-in for dune-grid's VTKWriter (third party, absent).  The reference's SimulatorUtilities.hpp names it
// only inside the template writeVtkOutput (:233-285), which the oracle never instantiates; the declarations
// below merely let the header parse.  The arithmetic the oracle uses from that header is
// estimateCellVelocity (:59-86), computePhaseVelocities (:153-170) and computeCapPressure (:219-230).
#ifndef ORACLE_SHIM_DUNE_VTKWRITER_HH
#define ORACLE_SHIM_DUNE_VTKWRITER_HH
#include <string>
#include <vector>
namespace Dune {
    namespace VTK { enum OutputType { ascii }; }
    template <class GridView>
    class VTKWriter {
    public:
        explicit VTKWriter(const GridView&) {}
        template <class V> void addCellData(const V&, const std::string&, int = 1) {}
        void write(const std::string&, VTK::OutputType) {}
    };
}
#endif
