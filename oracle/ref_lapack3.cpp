// TEST INFRASTRUCTURE -- not product code.
//
// The reference's TensorMobility::setToInverse (ReservoirPropertyCapillaryAnisotropicRelperm.hpp:
// 79-83) calls invert() (Matrix.hpp:780-802) = LAPACK dgetrf + dgetri.  No BLAS/LAPACK is
// installed in this image, so the two routines are restated here from the published
// unblocked algorithms (dgetf2: partial-pivot right-looking LU, column scaling by the
// reciprocal pivot; dgetri: invert U in place, then solve inv(A)*L = inv(U) column by
// column from the right, then undo the column interchanges).  Column-major, as LAPACK.
// On the diagonal matrices the anisotropic-relperm path actually produces
// (RockAnisotropicRelperm.hpp:71-77) every LAPACK variant yields exactly diag(1/d_i),
// or info>0 with the matrix left as its own LU factors when some d_i == 0.
// The other BLAS/LAPACK entry points declared in blas_lapack.hpp are never reached from
// the transport path; they are defined as traps so the library links.
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <vector>

extern "C" {

void dgetrf_(const int* m_, const int* n_, double* A, const int* ld_, int* ipiv, int* info)
{
    const int m = *m_, n = *n_, ld = *ld_;
    *info = 0;
    const int mn = m < n ? m : n;
    for (int j = 0; j < mn; ++j) {
        // idamax over A(j:m-1, j)
        int p = j;
        double best = std::fabs(A[j + j*ld]);
        for (int i = j + 1; i < m; ++i) {
            if (std::fabs(A[i + j*ld]) > best) { best = std::fabs(A[i + j*ld]); p = i; }
        }
        ipiv[j] = p + 1;
        if (A[p + j*ld] != 0.0) {
            if (p != j) {
                for (int k = 0; k < n; ++k) { double t = A[j + k*ld]; A[j + k*ld] = A[p + k*ld]; A[p + k*ld] = t; }
            }
            if (j < m - 1) {
                const double r = 1.0 / A[j + j*ld];
                for (int i = j + 1; i < m; ++i) { A[i + j*ld] *= r; }
            }
        } else if (*info == 0) {
            *info = j + 1;
        }
        if (j < mn - 1) {
            for (int k = j + 1; k < n; ++k) {
                const double akj = A[j + k*ld];
                for (int i = j + 1; i < m; ++i) { A[i + k*ld] -= A[i + j*ld]*akj; }
            }
        }
    }
}

void dgetri_(const int* n_, double* A, const int* ld_, const int* ipiv, double* work, int* /*lwork*/, int* info)
{
    const int n = *n_, ld = *ld_;
    *info = 0;
    // dtrti2 (upper, non-unit)
    for (int j = 0; j < n; ++j) {
        if (A[j + j*ld] == 0.0) { *info = j + 1; return; }
    }
    for (int j = 0; j < n; ++j) {
        A[j + j*ld] = 1.0 / A[j + j*ld];
        const double ajj = -A[j + j*ld];
        // x := U(0:j-1,0:j-1) * A(0:j-1, j)   (dtrmv upper, no-trans, non-unit)
        for (int k = 0; k < j; ++k) {
            const double t = A[k + j*ld];
            if (t != 0.0) {
                for (int i = 0; i < k; ++i) { A[i + j*ld] += t*A[i + k*ld]; }
                A[k + j*ld] = t*A[k + k*ld];
            }
        }
        for (int i = 0; i < j; ++i) { A[i + j*ld] *= ajj; }
    }
    // solve inv(A)*L = inv(U)
    for (int j = n - 2; j >= 0; --j) {
        for (int i = j + 1; i < n; ++i) { work[i] = A[i + j*ld]; A[i + j*ld] = 0.0; }
        for (int k = j + 1; k < n; ++k) {
            const double t = -work[k];
            for (int i = 0; i < n; ++i) { A[i + j*ld] += t*A[i + k*ld]; }
        }
    }
    for (int j = n - 2; j >= 0; --j) {
        const int jp = ipiv[j] - 1;
        if (jp != j) {
            for (int i = 0; i < n; ++i) { double t = A[i + j*ld]; A[i + j*ld] = A[i + jp*ld]; A[i + jp*ld] = t; }
        }
    }
}

#define TRAP(name) void name() { std::fprintf(stderr, "oracle: unexpected call to " #name "\n"); std::abort(); }
TRAP(dgemv_)
TRAP(dgemm_)
TRAP(dsyrk_)
TRAP(dtrmm_)
TRAP(dgeqrf_)
TRAP(dorgqr_)

} // extern "C"
