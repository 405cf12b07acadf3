// TEST INFRASTRUCTURE (oracle shim) -- not product code.
// Boost is absent. ReservoirPropertyCapillaryAnisotropicRelperm_impl.hpp:95 uses
// boost::lambda::_1/visc as a unary functor x -> x/visc.
#ifndef ORACLE_SHIM_BOOST_LAMBDA_HPP
#define ORACLE_SHIM_BOOST_LAMBDA_HPP
namespace boost { namespace lambda {
    struct DivideBy { double d; double operator()(double x) const { return x/d; } };
    struct Placeholder1 {};
    inline DivideBy operator/(const Placeholder1&, double d) { DivideBy f = { d }; return f; }
    static const Placeholder1 _1 = Placeholder1();
}}
#endif
