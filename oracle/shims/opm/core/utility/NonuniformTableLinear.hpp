// TEST INFRASTRUCTURE (oracle shim) -- not product code.
// Stand-in for opm-core 2016.04's <opm/core/utility/NonuniformTableLinear.hpp>
// and <opm/core/utility/linearInterpolation.hpp> (third party, absent from
// /root/reference; dune.module:13 names opm-core without a version pin beyond the
// module's own 2016.04-pre). Restated from the published algorithm:
//   tableIndex(): binary search for the interval j with table[j] <= x < table[j+1];
//                 first/last interval when x is out of range;
//   value:        (y[j+1]-y[j])/(x[j+1]-x[j])*(x-x[j]) + y[j]   (slope form,
//                 therefore linear EXTRAPOLATION outside the table);
//   derivative:   (y[j+1]-y[j])/(x[j+1]-x[j]);
//   inverse:      same interpolation with the roles of x and y swapped
//                 (reversed copies when y is decreasing).
// The reference pins nothing at this boundary ("parity unpinned"): this
// shim IS the contract shared by the oracle and the device tables.
// Call sites: RockJfunc.hpp:70-88,107,110,220; RockAnisotropicRelperm.hpp:74-76,82,88.
#ifndef ORACLE_SHIM_NONUNIFORMTABLELINEAR_HPP
#define ORACLE_SHIM_NONUNIFORMTABLELINEAR_HPP
#include <vector>
#include <algorithm>
#include <cassert>
namespace Opm {
    inline int tableIndex(const std::vector<double>& table, double x)
    {
        int n = int(table.size()) - 1;
        if (n < 2) {
            return 0;
        }
        int jl = 0;
        int ju = n;
        bool ascend = (table[n] > table[0]);
        while (ju - jl > 1) {
            int jm = (ju + jl)/2;
            if ((x >= table[jm]) == ascend) {
                jl = jm;
            } else {
                ju = jm;
            }
        }
        return jl;
    }
    inline double linearInterpolationDerivative(const std::vector<double>& xv,
                                                const std::vector<double>& yv, double x)
    {
        int ix1 = tableIndex(xv, x);
        int ix2 = ix1 + 1;
        return (yv[ix2] - yv[ix1])/(xv[ix2] - xv[ix1]);
    }
    inline double linearInterpolation(const std::vector<double>& xv,
                                      const std::vector<double>& yv, double x)
    {
        int ix1 = tableIndex(xv, x);
        int ix2 = ix1 + 1;
        return (yv[ix2] - yv[ix1])/(xv[ix2] - xv[ix1])*(x - xv[ix1]) + yv[ix1];
    }

    template <typename T>
    class NonuniformTableLinear {
    public:
        NonuniformTableLinear() {}
        template <class XC, class YC>
        NonuniformTableLinear(const XC& x, const YC& y)
            : x_values_(x.begin(), x.end()), y_values_(y.begin(), y.end())
        {
            assert(x_values_.size() == y_values_.size());
        }
        std::pair<double, double> domain() { return std::make_pair(x_values_.front(), x_values_.back()); }
        double operator()(const double x) const { return linearInterpolation(x_values_, y_values_, x); }
        double derivative(const double x) const { return linearInterpolationDerivative(x_values_, y_values_, x); }
        double inverse(const double y) const
        {
            if (y_values_.front() < y_values_.back()) {
                return linearInterpolation(y_values_, x_values_, y);
            } else {
                if (y_values_reversed_.empty()) {
                    y_values_reversed_ = y_values_;
                    std::reverse(y_values_reversed_.begin(), y_values_reversed_.end());
                    x_values_reversed_ = x_values_;
                    std::reverse(x_values_reversed_.begin(), x_values_reversed_.end());
                }
                return linearInterpolation(y_values_reversed_, x_values_reversed_, y);
            }
        }
        bool operator==(const NonuniformTableLinear<T>& o) const
        {
            return x_values_ == o.x_values_ && y_values_ == o.y_values_;
        }
        // Oracle-only accessors (used to hand the very same nodes to the device tables).
        const std::vector<double>& xValues() const { return x_values_; }
        const std::vector<T>& yValues() const { return y_values_; }
    protected:
        std::vector<double> x_values_;
        std::vector<T> y_values_;
        mutable std::vector<double> x_values_reversed_;
        mutable std::vector<T> y_values_reversed_;
    };
}
#endif
