// TEST INFRASTRUCTURE (oracle shim) -- see Deck.hpp in this directory.
#include <opm/parser/eclipse/Deck/Deck.hpp>
