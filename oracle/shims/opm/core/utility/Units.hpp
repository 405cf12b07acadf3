// TEST INFRASTRUCTURE (oracle shim) -- not product code.
// Stand-in for opm-core's <opm/core/utility/Units.hpp> (third party, absent).
// Only the constants touched by ReservoirPropertyCommon*.hpp / EulerUpstream_impl.hpp.
#ifndef ORACLE_SHIM_UNITS_HPP
#define ORACLE_SHIM_UNITS_HPP
namespace Opm {
    namespace prefix {
        const double micro = 1.0e-6;
        const double milli = 1.0e-3;
        const double centi = 1.0e-2;
        const double deci  = 1.0e-1;
        const double kilo  = 1.0e3;
        const double mega  = 1.0e6;
        const double giga  = 1.0e9;
    }
    namespace unit {
        inline double square(double v) { return v*v; }
        inline double cubic(double v)  { return v*v*v; }
        const double meter  = 1;
        const double inch   = 2.54 * prefix::centi*meter;
        const double feet   = 12 * inch;
        const double second = 1;
        const double minute = 60 * second;
        const double hour   = 60 * minute;
        const double day    = 24 * hour;
        const double year   = 365 * day;
        const double kilogram = 1;
        const double gravity = 9.80665 * meter/square(second);
        const double Newton = kilogram*meter / square(second);
        const double Pascal = Newton / square(meter);
        const double barsa  = 100000 * Pascal;
        const double atm    = 101325 * Pascal;
        const double Pas    = Pascal * second;
        const double Poise  = prefix::deci*Pas;
        namespace perm_details {
            const double p_grad   = atm / (prefix::centi*meter);
            const double area     = square(prefix::centi*meter);
            const double flux     = cubic (prefix::centi*meter) / second;
            const double velocity = flux / area;
            const double visc     = prefix::centi*Poise;
            const double darcy    = (velocity * visc) / p_grad;
        }
        const double darcy = perm_details::darcy;
        namespace convert {
            inline double from(const double q, const double unit) { return q * unit; }
            inline double to  (const double q, const double unit) { return q / unit; }
        }
    }
}
#endif
