// TEST INFRASTRUCTURE (oracle shim): boost::mpl::if_c == std::conditional (BoundaryConditions.hpp:343-350).
#ifndef ORACLE_SHIM_BOOST_MPL_IF_HPP
#define ORACLE_SHIM_BOOST_MPL_IF_HPP
#include <type_traits>
namespace boost { namespace mpl {
    template <bool C, class A, class B> struct if_c { typedef typename std::conditional<C, A, B>::type type; };
}}
#endif
