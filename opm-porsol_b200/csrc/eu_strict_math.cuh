// Device functions that reproduce the reference's arithmetic operation by operation.
// MUST be compiled with -fmad=false (no FMA contraction): the order of every add, multiply
// and divide below is the reference's, so the results are bit-identical to the CPU code.
// Citations are relative to the reference tree (opm/porsol/...).
#ifndef EU_STRICT_MATH_CUH
#define EU_STRICT_MATH_CUH

#include "eu_internal.h"

// opm-core 2016.04 linearInterpolation.hpp tableIndex(): binary search, first/last interval
// outside the table (third party; call sites common/RockJfunc.hpp:70-88,107,110).
__device__ __forceinline__ int sm_table_index(int size, const double* __restrict__ table, double x)
{
    int n = size - 1;
    if (n < 2) return 0;
    int jl = 0, ju = n;
    const bool ascend = (table[n] > table[0]);
    while (ju - jl > 1) {
        int jm = (ju + jl)/2;
        if ((x >= table[jm]) == ascend) jl = jm; else ju = jm;
    }
    return jl;
}

// linearInterpolation(): slope form, extrapolates outside the table.
__device__ __forceinline__ double sm_table_eval(int n, const double* __restrict__ x, const double* __restrict__ y, double xv)
{
    int i1 = sm_table_index(n, x, xv);
    int i2 = i1 + 1;
    return (y[i2] - y[i1])/(x[i2] - x[i1])*(xv - x[i1]) + y[i1];
}

__device__ __forceinline__ double sm_rock_col(const EuTablesDev& t, int rock, int col, double s)
{
    int b = t.offset[rock];
    int n = t.offset[rock + 1] - b;
    return sm_table_eval(n, t.s + b, t.cols[col] + b, s);
}

// kind 0: common/ReservoirPropertyCapillary_impl.hpp:44-70,118-131,154-168
// kind 1: common/ReservoirPropertyCapillaryAnisotropicRelperm_impl.hpp:75-101,
//         common/RockAnisotropicRelperm.hpp:70-77
template <int KIND>
__device__ __forceinline__ void sm_mobility(const EuTablesDev& t, int phase, int rock, double s, double* mob)
{
    const double visc = t.visc[phase];
    if (KIND == 0) {
        double kr;
        if (t.n_rocks > 0) {
            kr = sm_rock_col(t, rock, phase, s);
        } else {
            kr = phase == 0 ? s*s : (1 - s)*(1 - s);
        }
        mob[0] = kr/visc;
    } else {
        for (int i = 0; i < 9; ++i) mob[i] = 0.0;
        if (t.n_rocks > 0) {
            mob[0] = sm_rock_col(t, rock, 1 + 3*phase + 0, s);
            mob[4] = sm_rock_col(t, rock, 1 + 3*phase + 1, s);
            mob[8] = sm_rock_col(t, rock, 1 + 3*phase + 2, s);
            for (int i = 0; i < 9; ++i) mob[i] = mob[i]/visc;
        } else {
            double kr = phase == 0 ? s*s : (1.0 - s)*(1.0 - s);
            double m = kr/visc;
            mob[0] = mob[4] = mob[8] = m;
        }
    }
}

// common/ReservoirPropertyCommon_impl.hpp:465-476, common/RockJfunc.hpp:99-112,
// common/RockAnisotropicRelperm.hpp:79-83
__device__ __forceinline__ double sm_cap_pressure(const EuTablesDev& t, int rock, const double* __restrict__ K,
                                                  double poro, double s)
{
    if (t.n_rocks > 0) {
        if (t.kind == EU_MOB_DIAGONAL) return sm_rock_col(t, rock, 0, s);
        double J = sm_rock_col(t, rock, 2, s);
        if (t.use_j) {
            double tr = 0;                // common/Matrix.hpp:637-647
            tr += K[0]; tr += K[4]; tr += K[8];
            double sqrt_k_phi = sqrt(tr/(3*poro));
            return J*t.sigma_cos_theta/sqrt_k_phi;
        }
        return J;
    }
    return 1e5*(1 - s);
}

// common/Matrix.hpp:667-682
__device__ __forceinline__ void sm_prod3(const double* A, const double* x, double* res)
{
    res[0] = res[1] = res[2] = 0.0;
#pragma unroll
    for (int col = 0; col < 3; ++col) {
#pragma unroll
        for (int row = 0; row < 3; ++row) res[row] += A[3*row + col]*x[col];
    }
}
// dune-cornerpoint Volumes.hpp inner(): std::inner_product from 0
__device__ __forceinline__ double sm_inner3(const double* a, const double* b)
{
    double r = 0.0;
    r = r + a[0]*b[0];
    r = r + a[1]*b[1];
    r = r + a[2]*b[2];
    return r;
}
__device__ __forceinline__ double sm_two_norm3(const double* a)
{
    double r = 0.0;
    r += a[0]*a[0];
    r += a[1]*a[1];
    r += a[2]*a[2];
    return sqrt(r);
}
// opm-core Average.hpp: r = a; r += b; r *= 0.5
__device__ __forceinline__ void sm_aver9(const double* __restrict__ a, const double* __restrict__ b, double* r)
{
#pragma unroll
    for (int i = 0; i < 9; ++i) { double v = a[i]; v += b[i]; v *= 0.5; r[i] = v; }
}

// LAPACK dgetrf+dgetri on 3x3 (common/Matrix.hpp:780-802); same unblocked algorithm as
// oracle/ref_lapack3.cpp.  Used by TensorMobility::setToInverse only.
__device__ inline int sm_invert3(double* A)
{
    const int n = 3, ld = 3;
    int ipiv[3];
    int info = 0;
    for (int j = 0; j < n; ++j) {
        int p = j;
        double best = fabs(A[j + j*ld]);
        for (int i = j + 1; i < n; ++i) {
            if (fabs(A[i + j*ld]) > best) { best = fabs(A[i + j*ld]); p = i; }
        }
        ipiv[j] = p;
        if (A[p + j*ld] != 0.0) {
            if (p != j) {
                for (int k = 0; k < n; ++k) { double v = A[j + k*ld]; A[j + k*ld] = A[p + k*ld]; A[p + k*ld] = v; }
            }
            if (j < n - 1) {
                const double r = 1.0/A[j + j*ld];
                for (int i = j + 1; i < n; ++i) A[i + j*ld] *= r;
            }
        } else if (info == 0) {
            info = j + 1;
        }
        if (j < n - 1) {
            for (int k = j + 1; k < n; ++k) {
                const double akj = A[j + k*ld];
                for (int i = j + 1; i < n; ++i) A[i + k*ld] -= A[i + j*ld]*akj;
            }
        }
    }
    if (info != 0) return info;
    for (int j = 0; j < n; ++j) {
        A[j + j*ld] = 1.0/A[j + j*ld];
        const double ajj = -A[j + j*ld];
        for (int k = 0; k < j; ++k) {
            const double v = A[k + j*ld];
            if (v != 0.0) {
                for (int i = 0; i < k; ++i) A[i + j*ld] += v*A[i + k*ld];
                A[k + j*ld] = v*A[k + k*ld];
            }
        }
        for (int i = 0; i < j; ++i) A[i + j*ld] *= ajj;
    }
    double work[3];
    for (int j = n - 2; j >= 0; --j) {
        for (int i = j + 1; i < n; ++i) { work[i] = A[i + j*ld]; A[i + j*ld] = 0.0; }
        for (int k = j + 1; k < n; ++k) {
            const double v = -work[k];
            for (int i = 0; i < n; ++i) A[i + j*ld] += v*A[i + k*ld];
        }
    }
    for (int j = n - 2; j >= 0; --j) {
        const int jp = ipiv[j];
        if (jp != j) {
            for (int i = 0; i < n; ++i) { double v = A[i + j*ld]; A[i + j*ld] = A[i + jp*ld]; A[i + jp*ld] = v; }
        }
    }
    return 0;
}

// ScalarMobility (common/ReservoirPropertyCapillary.hpp:47-74) /
// TensorMobility<3> (common/ReservoirPropertyCapillaryAnisotropicRelperm.hpp:49-100)
template <int KIND>
__device__ __forceinline__ void sm_mob_multiply(const double* m, const double* v, double* out)
{
    if (KIND == 0) {
        out[0] = v[0]*m[0]; out[1] = v[1]*m[0]; out[2] = v[2]*m[0];
    } else {
        double tmp[3];
        sm_prod3(m, v, tmp);
        out[0] = tmp[0]; out[1] = tmp[1]; out[2] = tmp[2];
    }
}
template <int KIND>
__device__ __forceinline__ void sm_mob_sum(const double* a, const double* b, double* r)
{
    const int n = KIND == 0 ? 1 : 9;
#pragma unroll
    for (int i = 0; i < n; ++i) r[i] = a[i] + b[i];
}
template <int KIND>
__device__ __forceinline__ void sm_mob_average(const double* a, const double* b, double* r)
{
    const int n = KIND == 0 ? 1 : 9;
#pragma unroll
    for (int i = 0; i < n; ++i) r[i] = 0.5*(a[i] + b[i]);
}
template <int KIND>
__device__ __forceinline__ void sm_mob_inverse(const double* a, double* r)
{
    if (KIND == 0) {
        r[0] = 1.0/a[0];
    } else {
        for (int i = 0; i < 9; ++i) r[i] = a[i];
        sm_invert3(r);
    }
}

// euler/EulerUpstreamResidual_impl.hpp:510-547: direction and spacing of the two-point
// capillary-pressure gradient through the face centroid(s).
__device__ __forceinline__ void sm_cap_direction(const double* cell_c, const double* nb_c, const double* f_c,
                                                 const double* nbf_c, double* dirhat, double* d0d1)
{
    double a[3], b[3];
#pragma unroll
    for (int i = 0; i < 3; ++i) { a[i] = cell_c[i] - f_c[i]; b[i] = nb_c[i] - nbf_c[i]; }
    double d0 = sm_two_norm3(a);
    double d1 = sm_two_norm3(b);
    *d0d1 = d0 + d1;
#pragma unroll
    for (int i = 0; i < 3; ++i) {
        double v = nb_c[i] - nbf_c[i];
        v = v + f_c[i];
        v = v - cell_c[i];
        dirhat[i] = v;
    }
    double nrm = sm_two_norm3(dirhat);
#pragma unroll
    for (int i = 0; i < 3; ++i) dirhat[i] /= nrm;
}

#endif
