// TEST INFRASTRUCTURE (oracle shim) -- not product code.
// Miniature stand-in for opm-parser's Deck/DeckKeyword (third party, absent): a
// keyword -> data map with exactly the calls ReservoirPropertyCommon_impl.hpp:70-115,
// 195-242,592-779 makes (hasKeyword, getKeyword().getSIDoubleData()/getIntData()).
#ifndef ORACLE_SHIM_DECK_HPP
#define ORACLE_SHIM_DECK_HPP
#include <map>
#include <memory>
#include <string>
#include <vector>
#include <array>
#include <stdexcept>
namespace Opm {
    class DeckKeyword {
    public:
        const std::vector<double>& getSIDoubleData() const { return d_; }
        const std::vector<int>& getIntData() const { return i_; }
        std::vector<double> d_;
        std::vector<int> i_;
    };
    class Deck {
    public:
        bool hasKeyword(const std::string& k) const { return kw_.count(k) != 0; }
        const DeckKeyword& getKeyword(const std::string& k) const
        {
            std::map<std::string, DeckKeyword>::const_iterator it = kw_.find(k);
            if (it == kw_.end()) { throw std::runtime_error("Deck: no keyword " + k); }
            return it->second;
        }
        void setDouble(const std::string& k, const std::vector<double>& v) { kw_[k].d_ = v; }
        void setInt(const std::string& k, const std::vector<int>& v) { kw_[k].i_ = v; }
        std::array<int, 3> dims_;
    private:
        std::map<std::string, DeckKeyword> kw_;
    };
    typedef std::shared_ptr<const Deck> DeckConstPtr;
    typedef std::shared_ptr<Deck> DeckPtr;
}
#endif
