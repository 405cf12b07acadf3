// TEST INFRASTRUCTURE (oracle shim) -- not product code.
// Stand-in for opm-core's EclipseGridInspector (third party, absent); only gridSize().
#ifndef ORACLE_SHIM_ECLIPSEGRIDINSPECTOR_HPP
#define ORACLE_SHIM_ECLIPSEGRIDINSPECTOR_HPP
#include <opm/parser/eclipse/Deck/Deck.hpp>
namespace Opm {
    class EclipseGridInspector {
    public:
        explicit EclipseGridInspector(DeckConstPtr deck) : deck_(deck) {}
        std::array<int, 3> gridSize() const { return deck_->dims_; }
    private:
        DeckConstPtr deck_;
    };
}
#endif
