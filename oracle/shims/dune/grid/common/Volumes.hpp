// TEST INFRASTRUCTURE (oracle shim) -- not product code.
// Stand-in for dune-cornerpoint's <dune/grid/common/Volumes.hpp> (third party, absent):
// inner(a,b) = std::inner_product(a.begin(), a.end(), b.begin(), T()) -- a left fold.
// Call sites: EulerUpstreamResidual_impl.hpp:204,215,249,260,273.
#ifndef ORACLE_SHIM_VOLUMES_HPP
#define ORACLE_SHIM_VOLUMES_HPP
#include <dune/common/fvector.hh>
#include <numeric>
namespace Dune {
    template <typename T, int dim>
    inline T inner(const FieldVector<T, dim>& a, const FieldVector<T, dim>& b)
    {
        return std::inner_product(a.begin(), a.end(), b.begin(), T());
    }
}
#endif
