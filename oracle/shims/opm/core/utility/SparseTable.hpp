// TEST INFRASTRUCTURE (oracle shim) -- not product code.
// Minimal stand-in for opm-core's SparseTable (third party, absent): CSR rows of T.
#ifndef ORACLE_SHIM_SPARSETABLE_HPP
#define ORACLE_SHIM_SPARSETABLE_HPP
#include <vector>
namespace Opm {
    template <typename T>
    class SparseTable {
    public:
        SparseTable() : row_start_(1, 0) {}
        template <class It> void appendRow(It b, It e) { data_.insert(data_.end(), b, e); row_start_.push_back(int(data_.size())); }
        int size() const { return int(row_start_.size()) - 1; }
        int dataSize() const { return int(data_.size()); }
        const T* operator[](int r) const { return &data_[row_start_[r]]; }
        T* operator[](int r) { return &data_[row_start_[r]]; }
        int rowSize(int r) const { return row_start_[r+1] - row_start_[r]; }
        const T& data(int i) const { return data_[i]; }
        void clear() { data_.clear(); row_start_.assign(1, 0); }
    private:
        std::vector<T> data_;
        std::vector<int> row_start_;
    };
}
#endif
