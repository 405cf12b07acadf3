// TEST INFRASTRUCTURE (oracle shim) -- not product code.
// Stand-in for opm-common's <opm/common/ErrorMacros.hpp>, which is a third-party
// header absent from /root/reference. Call sites: EulerUpstream_impl.hpp:208,345,
// CflCalculator.hpp:76, RockJfunc.hpp:156, BoundaryConditions.hpp:148.
#ifndef ORACLE_SHIM_ERRORMACROS_HPP
#define ORACLE_SHIM_ERRORMACROS_HPP
#include <sstream>
#include <stdexcept>
#include <iostream>
#include <cassert>

#define OPM_THROW(Exception, message)                                        \
    do {                                                                     \
        std::ostringstream oss__;                                            \
        oss__ << message;                                                    \
        throw Exception(oss__.str());                                        \
    } while (false)

#ifdef ORACLE_SHIM_VERBOSE_MESSAGES
#define OPM_MESSAGE(x) do { std::cerr << x << std::endl; } while (false)
#else
#define OPM_MESSAGE(x) do { } while (false)
#endif

#define OPM_ERROR_IF(cond, message) do { if (cond) { OPM_THROW(std::logic_error, message); } } while (false)
#define OPM_MESSAGE_IF(cond, m) do { if (cond) OPM_MESSAGE(m); } while (false)
#endif
