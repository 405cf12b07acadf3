// TEST INFRASTRUCTURE (oracle shim) -- not product code.
// Stand-in for dune-common's FieldVector (third party, absent). Operation order follows
// the published DenseVector implementation: two_norm() = sqrt of a left-fold sum of
// squares from 0; operator* (dot) = left fold from 0; operator/=(k) divides each entry
// (no reciprocal); binary +,- copy the left operand and apply +=,-=.
// Call sites: EulerUpstreamResidual_impl.hpp:532-546, CflCalculator.hpp:168-170,
// Matrix.hpp:667-682, RockJfunc.hpp:186-189.
#ifndef ORACLE_SHIM_FVECTOR_HH
#define ORACLE_SHIM_FVECTOR_HH
#include <cmath>
#include <cstddef>
#include <iterator>
#include <iostream>
#include <numeric>
namespace Dune {
    template <class K, int SIZE>
    class FieldVector {
    public:
        enum { dimension = SIZE, size_ = SIZE };
        typedef K value_type;
        typedef K field_type;
        typedef K* iterator;
        typedef const K* const_iterator;
        typedef std::size_t size_type;
        FieldVector() { for (int i = 0; i < SIZE; ++i) d_[i] = K(); }
        FieldVector(const K& k) { for (int i = 0; i < SIZE; ++i) d_[i] = k; }
        FieldVector& operator=(const K& k) { for (int i = 0; i < SIZE; ++i) d_[i] = k; return *this; }
        K& operator[](size_type i) { return d_[i]; }
        const K& operator[](size_type i) const { return d_[i]; }
        iterator begin() { return d_; }
        iterator end() { return d_ + SIZE; }
        const_iterator begin() const { return d_; }
        const_iterator end() const { return d_ + SIZE; }
        size_type size() const { return SIZE; }
        FieldVector& operator+=(const FieldVector& y) { for (int i = 0; i < SIZE; ++i) d_[i] += y.d_[i]; return *this; }
        FieldVector& operator-=(const FieldVector& y) { for (int i = 0; i < SIZE; ++i) d_[i] -= y.d_[i]; return *this; }
        FieldVector& operator*=(const K& k) { for (int i = 0; i < SIZE; ++i) d_[i] *= k; return *this; }
        FieldVector& operator/=(const K& k) { for (int i = 0; i < SIZE; ++i) d_[i] /= k; return *this; }
        FieldVector operator+(const FieldVector& b) const { FieldVector z = *this; return (z += b); }
        FieldVector operator-(const FieldVector& b) const { FieldVector z = *this; return (z -= b); }
        K operator*(const FieldVector& y) const
        {
            K result(0);
            for (int i = 0; i < SIZE; ++i) { result += d_[i]*y.d_[i]; }
            return result;
        }
        K two_norm2() const
        {
            K result(0);
            for (int i = 0; i < SIZE; ++i) { result += d_[i]*d_[i]; }
            return result;
        }
        K two_norm() const { return std::sqrt(two_norm2()); }
        bool operator==(const FieldVector& o) const { for (int i = 0; i < SIZE; ++i) if (d_[i] != o.d_[i]) return false; return true; }
        bool operator!=(const FieldVector& o) const { return !(*this == o); }
    private:
        K d_[SIZE];
    };
    template <class K, int SIZE>
    inline std::istream& operator>>(std::istream& in, FieldVector<K, SIZE>& v)
    {
        FieldVector<K, SIZE> w;
        for (int i = 0; i < SIZE; ++i) { in >> w[i]; }
        if (in) { v = w; }
        return in;
    }
    template <class K, int SIZE>
    inline std::ostream& operator<<(std::ostream& s, const FieldVector<K, SIZE>& v)
    {
        for (int i = 0; i < SIZE; ++i) { s << ((i > 0) ? " " : "") << v[i]; }
        return s;
    }
}
#endif
