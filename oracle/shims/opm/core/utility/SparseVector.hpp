// TEST INFRASTRUCTURE (oracle shim) -- not product code.
// Stand-in for opm-core's <opm/core/utility/SparseVector.hpp> (third party, absent).
// Call site on the hot path: EulerUpstreamResidual_impl.hpp:293 (element(i): value or 0).
#ifndef ORACLE_SHIM_SPARSEVECTOR_HPP
#define ORACLE_SHIM_SPARSEVECTOR_HPP
#include <vector>
#include <algorithm>
#include <cassert>
namespace Opm {
    template <typename T>
    class SparseVector {
    public:
        SparseVector() : size_(0), default_elem_() {}
        explicit SparseVector(int sz) : size_(sz), default_elem_() {}
        // Elements must be added in order of increasing index.
        void addElement(const T& elem, int index)
        {
            assert(indices_.empty() || index > indices_.back());
            assert(index < size_);
            data_.push_back(elem);
            indices_.push_back(index);
        }
        bool empty() const { return size_ == 0; }
        int size() const { return size_; }
        int nonzeroSize() const { return int(data_.size()); }
        void clear() { data_.clear(); indices_.clear(); size_ = 0; }
        const T& element(int index) const
        {
            std::vector<int>::const_iterator lb = std::lower_bound(indices_.begin(), indices_.end(), index);
            if (lb != indices_.end() && *lb == index) {
                return data_[lb - indices_.begin()];
            }
            return default_elem_;
        }
        const T& nonzeroElement(int i) const { return data_[i]; }
        int nonzeroIndex(int i) const { return indices_[i]; }
    private:
        std::vector<T> data_;
        std::vector<int> indices_;
        int size_;
        T default_elem_;
    };
}
#endif
