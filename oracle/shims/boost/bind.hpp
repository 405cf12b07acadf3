// TEST INFRASTRUCTURE (oracle shim) -- not product code.
// Boost is absent. Matrix.hpp:519-523 uses boost::bind(std::multiplies<T>(), _1, scalar);
// std::bind with std::placeholders::_1 is the standardised form of the same thing.
#ifndef ORACLE_SHIM_BOOST_BIND_HPP
#define ORACLE_SHIM_BOOST_BIND_HPP
#include <functional>
namespace boost { using std::bind; }
using std::placeholders::_1;
using std::placeholders::_2;
#endif
