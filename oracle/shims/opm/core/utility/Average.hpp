// TEST INFRASTRUCTURE (oracle shim) -- not product code.
// Stand-in for opm-core's <opm/core/utility/Average.hpp> (third party, absent).
// Published semantics: copy first operand into the result type, add the second,
// scale by one half. Call sites: EulerUpstreamResidual_impl.hpp:73,229;
// CflCalculator.hpp:110,160.
#ifndef ORACLE_SHIM_AVERAGE_HPP
#define ORACLE_SHIM_AVERAGE_HPP
#include <type_traits>
#include <cmath>
namespace Opm { namespace utils {
    template <typename T, typename Tresult>
    Tresult arithmeticAverage(const T& t1, const T& t2)
    {
        static_assert(!std::is_integral<T>::value, "no integral averages");
        Tresult retval(t1);
        retval += t2;
        retval *= 0.5;
        return retval;
    }
    template <typename T>
    T geometricAverage(const T& t1, const T& t2) { return std::sqrt(t1*t2); }
    template <typename T>
    T harmonicAverage(const T& t1, const T& t2) { return (2*t1*t2)/(t1 + t2); }
}}
#endif
