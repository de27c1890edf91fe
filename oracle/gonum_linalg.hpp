// oracle/gonum_linalg.hpp — TEST INFRASTRUCTURE ONLY (CPU oracle). Never linked into, imported by
// or executed from the product path (libgomilp_b200.so / gomilp_b200/); see oracle/README.md.
//
// Restatement, at arithmetic-order level, of the dense linear algebra that Gonum's lp.Simplex
// reaches (all paths relative to /root/reference/vendor/gonum.org/v1/gonum/):
//   mat/solve.go:21-140, mat/lu.go:28-134,293-325, mat/matrix.go:284-322, mat/qr.go:23-69,
//   lapack/gonum/{dgetrf,dgetf2,dgetrs,dlaswp,dgecon,dlacn2,dlatrs,drscl,dlange,dlantr,dtrcon,
//   dgeqrf,dgeqr2,dlarfg,dlarf,dlapy2}.go, blas/gonum/{level1double,level2double,level3double,dgemm}.go,
//   internal/asm/f64/{dot_amd64.s,axpyunitaryto_amd64.s}, floats/floats.go.
//
// Rounding model: Gonum's amd64 kernels are SSE2 (MULPD/ADDPD, no FMA), so every a*x+y here is a
// rounded multiply followed by a rounded add — compile with -ffp-contract=off (oracle/Makefile).
// All matrices are row-major (gonum/mat/dense.go:44-61).
//
// Deliberate, documented simplifications (neither can change a value the simplex reads, only the
// last bits of a condition ESTIMATE that is compared with 1e12 / 1e16):
//   * Dgeqrf's blocked path (k > 128 columns, nb = 32; dgeqrf.go:46-81, ilaenv.go:52-56,308-312) is
//     replaced by the unblocked Dgeqr2 recurrence for every size.
//   * Dgetrf's blocked path (n > 64, nb = 64; dgetrf.go:39-68) is executed as the unblocked
//     right-looking elimination. This is bit-identical, not an approximation: with a row-major
//     Dtrsm (level3double.go:92-113) and Gonum's Dgemm (dgemm.go:188-199) every trailing element
//     receives its rank-1 corrections as separate mul+add in ascending pivot order in both forms.
#pragma once
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <limits>
#include <vector>

namespace orc {

using vec = std::vector<double>;
using ivec = std::vector<int>;

// lapack/gonum/general.go:130-142
constexpr double kEps = 1.0 / 9007199254740992.0;  // dlamchE = 2^-53
constexpr double kPrec = 2.0 * kEps;               // dlamchP
constexpr double kSafeMin = 2.2250738585072014e-308;  // dlamchS = 2^-1022

enum class Norm { One, Inf };  // lapack.MaxColumnSum / lapack.MaxRowSum

// ---------------------------------------------------------------- level 1
// internal/asm/f64/dot_amd64.s:43-92: two 2-lane accumulators over i mod 4, tail into lane 0,
// combined as (s0+s2)+(s1+s3).
inline double dot_unitary(const double* x, const double* y, int n) {
    double s0 = 0, s1 = 0, s2 = 0, s3 = 0;
    int i = 0;
    for (; i + 4 <= n; i += 4) {
        s0 += x[i] * y[i];
        s1 += x[i + 1] * y[i + 1];
        s2 += x[i + 2] * y[i + 2];
        s3 += x[i + 3] * y[i + 3];
    }
    for (; i < n; ++i) s0 += x[i] * y[i];
    return (s0 + s2) + (s1 + s3);
}

// blas/gonum/level1double.go:121-160 (first index of the largest |x_i|; NaN never wins)
inline int idamax(int n, const double* x, int inc) {
    if (n < 1) return -1;
    int best = 0;
    double mx = std::fabs(x[0]);
    for (int i = 1; i < n; ++i) {
        double a = std::fabs(x[(size_t)i * inc]);
        if (a > mx) { mx = a; best = i; }
    }
    return best;
}

inline double dasum(int n, const double* x, int inc) {  // level1double.go:91-116
    double s = 0;
    for (int i = 0; i < n; ++i) s += std::fabs(x[(size_t)i * inc]);
    return s;
}

inline void dscal(int n, double a, double* x, int inc) {
    for (int i = 0; i < n; ++i) x[(size_t)i * inc] *= a;
}

// y += a*x, element-wise mul then add (axpyunitaryto_amd64.s:51-140; no cross-element reduction)
inline void axpy(int n, double a, const double* x, int incx, double* y, int incy) {
    for (int i = 0; i < n; ++i) y[(size_t)i * incy] = a * x[(size_t)i * incx] + y[(size_t)i * incy];
}

// level1double.go:19-88 scaled sum of squares
inline double dnrm2(int n, const double* x, int inc) {
    if (n < 1) return 0;
    if (n == 1) return std::fabs(x[0]);
    double scale = 0, ssq = 1;
    for (int i = 0; i < n; ++i) {
        double v = x[(size_t)i * inc];
        if (v == 0) continue;
        double a = std::fabs(v);
        if (std::isnan(a)) return a;
        if (scale < a) {
            ssq = 1 + ssq * (scale / a) * (scale / a);
            scale = a;
        } else {
            ssq = ssq + (a / scale) * (a / scale);
        }
    }
    if (std::isinf(scale)) return scale;
    return scale * std::sqrt(ssq);
}

// floats/floats.go:458-474 — first minimum, NaN skipped, index 0 if everything is NaN
inline int min_idx(const double* s, int n) {
    double mn = std::numeric_limits<double>::quiet_NaN();
    int ind = 0;
    for (int i = 0; i < n; ++i) {
        double v = s[i];
        if (std::isnan(v)) continue;
        if (v < mn || std::isnan(mn)) { mn = v; ind = i; }
    }
    return ind;
}

// ---------------------------------------------------------------- norms
inline double lange(Norm norm, int m, int n, const double* a, int lda) {  // dlange.go:43-71
    if (m == 0 && n == 0) return 0;
    if (norm == Norm::One) {
        vec w(n, 0.0);
        for (int i = 0; i < m; ++i)
            for (int j = 0; j < n; ++j) w[j] += std::fabs(a[(size_t)i * lda + j]);
        double v = 0;
        for (int j = 0; j < n; ++j) v = std::fmax(v, w[j]);  // math.Max: NaN propagates
        for (int j = 0; j < n; ++j) if (std::isnan(w[j])) return w[j];
        return v;
    }
    double v = 0;
    bool nan = false;
    for (int i = 0; i < m; ++i) {
        double s = 0;
        for (int j = 0; j < n; ++j) s += std::fabs(a[(size_t)i * lda + j]);
        if (std::isnan(s)) nan = true;
        v = std::fmax(v, s);
    }
    return nan ? std::numeric_limits<double>::quiet_NaN() : v;
}

// dlantr.go, upper / non-unit, square n×n only (the only form reached: qr.go:32-34)
inline double lantr_upper_nonunit(Norm norm, int n, const double* a, int lda) {
    if (n == 0) return 0;
    if (norm == Norm::One) {
        vec w(n, 0.0);
        for (int i = 0; i < n; ++i)
            for (int j = i; j < n; ++j) w[j] += std::fabs(a[(size_t)i * lda + j]);
        double mx = 0;
        for (int j = 0; j < n; ++j) {
            if (std::isnan(w[j])) return w[j];
            if (w[j] > mx) mx = w[j];
        }
        return mx;
    }
    double mx = 0;
    for (int i = 0; i < n; ++i) {
        double s = 0;
        for (int j = i; j < n; ++j) s += std::fabs(a[(size_t)i * lda + j]);
        if (std::isnan(s)) return s;
        if (s > mx) mx = s;
    }
    return mx;
}

// ---------------------------------------------------------------- LU
// dgetf2.go:30-69 executed over the whole matrix (see header note on blocking). Returns ok.
inline bool getrf(int n, double* a, int lda, int* ipiv) {
    bool ok = true;
    for (int j = 0; j < n; ++j) {
        int jp = j + idamax(n - j, a + (size_t)j * lda + j, lda);
        ipiv[j] = jp;
        if (a[(size_t)jp * lda + j] == 0) {
            ok = false;
        } else {
            if (jp != j) {
                double* r0 = a + (size_t)j * lda;
                double* r1 = a + (size_t)jp * lda;
                for (int k = 0; k < n; ++k) std::swap(r0[k], r1[k]);
            }
            if (j < n - 1) {
                double ajj = a[(size_t)j * lda + j];
                if (std::fabs(ajj) >= kSafeMin) {
                    dscal(n - j - 1, 1 / ajj, a + (size_t)(j + 1) * lda + j, lda);
                } else {
                    // dgetf2.go:57-60 divides the SAME element n-j-1 times
                    for (int i = 0; i < n - j - 1; ++i)
                        a[(size_t)(j + 1) * lda + j] = a[(size_t)(j + 1) * lda + j] / a[(size_t)j * lda + j];
                }
            }
        }
        if (j < n - 1) {
            // Dger(alpha=-1, x = column below the pivot (inc lda), y = pivot row tail)
            const double* y = a + (size_t)j * lda + j + 1;
            for (int i = j + 1; i < n; ++i) {
                double t = -1.0 * a[(size_t)i * lda + j];
                double* row = a + (size_t)i * lda + j + 1;
                for (int k = 0; k < n - j - 1; ++k) row[k] = t * y[k] + row[k];
            }
        }
    }
    return ok;
}

// dgetrs.go:41-47 NoTrans, one right-hand side; Dtrsm forms level3double.go:66-113
inline void getrs(int n, const double* lu, int lda, const int* ipiv, double* b) {
    for (int k = 0; k < n; ++k) std::swap(b[k], b[ipiv[k]]);
    for (int i = 0; i < n; ++i) {  // L, unit diagonal
        const double* row = lu + (size_t)i * lda;
        for (int k = 0; k < i; ++k) {
            double va = row[k];
            if (va != 0) b[i] = (-va) * b[k] + b[i];
        }
    }
    for (int i = n - 1; i >= 0; --i) {  // U, multiply by the reciprocal of the diagonal
        const double* row = lu + (size_t)i * lda;
        for (int k = i + 1; k < n; ++k) {
            double va = row[k];
            if (va != 0) b[i] = (-va) * b[k] + b[i];
        }
        double t = 1 / row[i];
        b[i] *= t;
    }
}

// ---------------------------------------------------------------- Dtrsv (only inside Dlatrs)
inline void trsv(bool upper, bool trans, bool nonunit, int n, const double* a, int lda, double* x) {
    if (n == 0) return;
    if (n == 1) { if (nonunit) x[0] /= a[0]; return; }
    if (!trans) {
        if (upper) {
            for (int i = n - 1; i >= 0; --i) {
                double s = 0;
                for (int j = i + 1; j < n; ++j) s += x[j] * a[(size_t)i * lda + j];
                x[i] -= s;
                if (nonunit) x[i] /= a[(size_t)i * lda + i];
            }
        } else {
            for (int i = 0; i < n; ++i) {
                double s = 0;
                for (int j = 0; j < i; ++j) s += x[j] * a[(size_t)i * lda + j];
                x[i] -= s;
                if (nonunit) x[i] /= a[(size_t)i * lda + i];
            }
        }
        return;
    }
    if (upper) {
        for (int i = 0; i < n; ++i) {
            if (nonunit) x[i] /= a[(size_t)i * lda + i];
            double xi = x[i];
            for (int j = i + 1; j < n; ++j) x[j] -= a[(size_t)i * lda + j] * xi;
        }
    } else {
        for (int i = n - 1; i >= 0; --i) {
            if (nonunit) x[i] /= a[(size_t)i * lda + i];
            double xi = x[i];
            for (int j = 0; j < i; ++j) x[j] -= a[(size_t)i * lda + j] * xi;
        }
    }
}

// ---------------------------------------------------------------- Dlatrs (dlatrs.go:30-350)
// Triangular solve with overflow-guarding scale factor. cnorm holds the off-diagonal column
// 1-norms (computed when !normin). Returns `scale` with T*x = scale*b.
inline double latrs(bool upper, bool trans, bool nonunit, bool normin, int n, const double* a, int lda,
                    double* x, double* cnorm) {
    if (n == 0) return 0;
    const double smlnum = kSafeMin / kPrec;
    const double bignum = 1 / smlnum;
    double scale = 1;
    if (!normin) {
        if (upper) {
            cnorm[0] = 0;
            for (int j = 1; j < n; ++j) cnorm[j] = dasum(j, a + j, lda);
        } else {
            for (int j = 0; j < n - 1; ++j) cnorm[j] = dasum(n - j - 1, a + (size_t)(j + 1) * lda + j, lda);
            cnorm[n - 1] = 0;
        }
    }
    double tmax = cnorm[idamax(n, cnorm, 1)];
    double tscal;
    if (tmax <= bignum) {
        tscal = 1;
    } else {
        tscal = 1 / (smlnum * tmax);
        dscal(n, tscal, cnorm, 1);
    }
    double xmax = std::fabs(x[idamax(n, x, 1)]);
    double xbnd = xmax;
    double grow = 0;
    int jfirst, jlast, jinc;
    const bool forward = (trans == upper);  // Upper+Trans and Lower+NoTrans sweep 0..n-1
    if (forward) { jfirst = 0; jlast = n; jinc = 1; } else { jfirst = n - 1; jlast = -1; jinc = -1; }

    // growth bound (decides whether the plain Dtrsv is safe)
    if (tscal != 1) {
        grow = 0;
    } else if (!trans) {
        if (nonunit) {
            grow = 1 / std::fmax(xbnd, smlnum);
            xbnd = grow;
            bool bail = false;
            for (int j = jfirst; j != jlast; j += jinc) {
                if (grow <= smlnum) { bail = true; break; }
                double tjj = std::fabs(a[(size_t)j * lda + j]);
                xbnd = std::fmin(xbnd, std::fmin(1.0, tjj) * grow);
                if (tjj + cnorm[j] >= smlnum) grow *= tjj / (tjj + cnorm[j]);
                else grow = 0;
            }
            if (!bail) grow = xbnd;
        } else {
            grow = std::fmin(1.0, 1 / std::fmax(xbnd, smlnum));
            for (int j = jfirst; j != jlast; j += jinc) {
                if (grow <= smlnum) break;
                grow *= 1 / (1 + cnorm[j]);
            }
        }
    } else {
        if (nonunit) {
            grow = 1 / std::fmax(xbnd, smlnum);
            xbnd = grow;
            bool bail = false;
            for (int j = jfirst; j != jlast; j += jinc) {
                if (grow <= smlnum) { bail = true; break; }
                double xj = 1 + cnorm[j];
                grow = std::fmin(grow, xbnd / xj);
                double tjj = std::fabs(a[(size_t)j * lda + j]);
                if (xj > tjj) xbnd *= tjj / xj;
            }
            if (!bail) grow = std::fmin(grow, xbnd);
        } else {
            grow = std::fmin(1.0, 1 / std::fmax(xbnd, smlnum));
            for (int j = jfirst; j != jlast; j += jinc) {
                if (grow <= smlnum) break;
                grow /= (1 + cnorm[j]);
            }
        }
    }

    if (grow * tscal > smlnum) {
        trsv(upper, trans, nonunit, n, a, lda, x);
        if (tscal != 1) dscal(n, 1 / tscal, cnorm, 1);
        return scale;
    }

    // careful path
    if (xmax > bignum) {
        scale = bignum / xmax;
        dscal(n, scale, x, 1);
        xmax = bignum;
    }
    auto singular_column = [&](int j) {
        for (int i = 0; i < n; ++i) x[i] = 0;
        x[j] = 1;
        scale = 0;
        xmax = 0;
    };
    if (!trans) {
        for (int j = jfirst; j != jlast; j += jinc) {
            double xj = std::fabs(x[j]);
            bool skip_div = false;
            double tjjs = 0;
            if (nonunit) {
                tjjs = a[(size_t)j * lda + j] * tscal;
            } else {
                tjjs = tscal;
                if (tscal == 1) skip_div = true;
            }
            if (!skip_div) {
                double tjj = std::fabs(tjjs);
                if (tjj > smlnum) {
                    if (tjj < 1 && xj > tjj * bignum) {
                        double rec = 1 / xj;
                        dscal(n, rec, x, 1);
                        scale *= rec;
                        xmax *= rec;
                    }
                    x[j] /= tjjs;
                    xj = std::fabs(x[j]);
                } else if (tjj > 0) {
                    if (xj > tjj * bignum) {
                        double rec = (tjj * bignum) / xj;
                        if (cnorm[j] > 1) rec /= cnorm[j];
                        dscal(n, rec, x, 1);
                        scale *= rec;
                        xmax *= rec;
                    }
                    x[j] /= tjjs;
                    xj = std::fabs(x[j]);
                } else {
                    singular_column(j);
                    xj = 1;
                }
            }
            if (xj > 1) {
                double rec = 1 / xj;
                if (cnorm[j] > (bignum - xmax) * rec) {
                    rec *= 0.5;
                    dscal(n, rec, x, 1);
                    scale *= rec;
                }
            } else if (xj * cnorm[j] > bignum - xmax) {
                dscal(n, 0.5, x, 1);
                scale *= 0.5;
            }
            if (upper) {
                if (j > 0) {
                    axpy(j, -x[j] * tscal, a + j, lda, x, 1);
                    xmax = std::fabs(x[idamax(j, x, 1)]);
                }
            } else if (j < n - 1) {
                axpy(n - j - 1, -x[j] * tscal, a + (size_t)(j + 1) * lda + j, lda, x + j + 1, 1);
                xmax = std::fabs(x[j + 1 + idamax(n - j - 1, x + j + 1, 1)]);
            }
        }
    } else {
        for (int j = jfirst; j != jlast; j += jinc) {
            double xj = std::fabs(x[j]);
            double uscal = tscal;
            double rec = 1 / std::fmax(xmax, 1.0);
            double tjjs = 0;
            if (cnorm[j] > (bignum - xj) * rec) {
                rec *= 0.5;
                tjjs = nonunit ? a[(size_t)j * lda + j] * tscal : tscal;
                double tjj = std::fabs(tjjs);
                if (tjj > 1) {
                    rec = std::fmin(1.0, rec * tjj);
                    uscal /= tjjs;
                }
                if (rec < 1) {
                    dscal(n, rec, x, 1);
                    scale *= rec;
                    xmax *= rec;
                }
            }
            double sumj = 0;
            if (uscal == 1) {
                // Ddot with a strided column: plain ascending accumulation (DotInc)
                if (upper) {
                    for (int i = 0; i < j; ++i) sumj += a[(size_t)i * lda + j] * x[i];
                } else if (j < n - 1) {
                    for (int i = j + 1; i < n; ++i) sumj += a[(size_t)i * lda + j] * x[i];
                }
            } else {
                if (upper) {
                    for (int i = 0; i < j; ++i) sumj += (a[(size_t)i * lda + j] * uscal) * x[i];
                } else {
                    for (int i = j + 1; i < n; ++i) sumj += (a[(size_t)i * lda + j] * uscal) * x[i];
                }
            }
            if (uscal == tscal) {
                x[j] -= sumj;
                double xjj = std::fabs(x[j]);
                bool skip_div = false;
                double t2 = 0;
                if (nonunit) {
                    t2 = a[(size_t)j * lda + j] * tscal;
                } else {
                    t2 = tscal;
                    if (tscal == 1) skip_div = true;
                }
                if (!skip_div) {
                    double tjj = std::fabs(t2);
                    if (tjj > smlnum) {
                        if (tjj < 1 && xjj > tjj * bignum) {
                            rec = 1 / xjj;
                            dscal(n, rec, x, 1);
                            scale *= rec;
                            xmax *= rec;
                        }
                        x[j] /= t2;
                    } else if (tjj > 0) {
                        if (xjj > tjj * bignum) {
                            rec = (tjj * bignum) / xjj;
                            dscal(n, rec, x, 1);
                            scale *= rec;
                            xmax *= rec;
                        }
                        x[j] /= t2;
                    } else {
                        singular_column(j);
                    }
                }
            } else {
                x[j] = x[j] / tjjs - sumj;
            }
            xmax = std::fmax(xmax, std::fabs(x[j]));
        }
    }
    scale /= tscal;
    if (tscal != 1) dscal(n, 1 / tscal, cnorm, 1);
    return scale;
}

// ---------------------------------------------------------------- Dlacn2 / Drscl
struct Lacn2State { int isave[3] = {0, 0, 0}; };

// dlacn2.go:30-134 reverse-communication 1-norm estimator (Higham). Returns (est, kase).
inline void lacn2(int n, double* v, double* x, int* isgn, double& est, int& kase, Lacn2State& st) {
    const int itmax = 5;
    auto copysign1 = [](double t) { return std::copysign(1.0, t); };
    if (kase == 0) {
        for (int i = 0; i < n; ++i) x[i] = 1 / (double)n;
        kase = 1;
        st.isave[0] = 1;
        return;
    }
    bool alt = false;  // fall through to the alternating-sign safeguard
    switch (st.isave[0]) {
    case 1:
        if (n == 1) {
            v[0] = x[0];
            est = std::fabs(v[0]);
            kase = 0;
            return;
        }
        est = dasum(n, x, 1);
        for (int i = 0; i < n; ++i) { x[i] = copysign1(x[i]); isgn[i] = (int)x[i]; }
        kase = 2;
        st.isave[0] = 2;
        return;
    case 2:
        st.isave[1] = idamax(n, x, 1);
        st.isave[2] = 2;
        for (int i = 0; i < n; ++i) x[i] = 0;
        x[st.isave[1]] = 1;
        kase = 1;
        st.isave[0] = 3;
        return;
    case 3: {
        std::memcpy(v, x, sizeof(double) * n);
        double estold = est;
        est = dasum(n, v, 1);
        bool same = true;
        for (int i = 0; i < n; ++i)
            if ((int)copysign1(x[i]) != isgn[i]) { same = false; break; }
        if (!same && est > estold) {
            for (int i = 0; i < n; ++i) { x[i] = copysign1(x[i]); isgn[i] = (int)x[i]; }
            kase = 2;
            st.isave[0] = 4;
            return;
        }
        alt = true;
        break;
    }
    case 4: {
        int jlast = st.isave[1];
        st.isave[1] = idamax(n, x, 1);
        if (x[jlast] != std::fabs(x[st.isave[1]]) && st.isave[2] < itmax) {
            st.isave[2] += 1;
            for (int i = 0; i < n; ++i) x[i] = 0;
            x[st.isave[1]] = 1;
            kase = 1;
            st.isave[0] = 3;
            return;
        }
        alt = true;
        break;
    }
    case 5: {
        double tmp = 2 * (dasum(n, x, 1)) / (double)(3 * n);
        if (tmp > est) {
            std::memcpy(v, x, sizeof(double) * n);
            est = tmp;
        }
        kase = 0;
        return;
    }
    default:
        kase = 0;
        return;
    }
    if (alt) {
        double altsgn = 1.0;
        for (int i = 0; i < n; ++i) {
            x[i] = altsgn * (1 + (double)i / (double)(n - 1));
            altsgn *= -1;
        }
        kase = 1;
        st.isave[0] = 5;
    }
}

inline void rscl(int n, double a, double* x) {  // drscl.go:15-47 : x /= a without over/underflow
    double cden = a, cnum = 1.0;
    const double smlnum = kSafeMin, bignum = 1 / smlnum;
    for (;;) {
        double cden1 = cden * smlnum, cnum1 = cnum / bignum, mul;
        bool done;
        if (cnum != 0 && std::fabs(cden1) > std::fabs(cnum)) { mul = smlnum; done = false; cden = cden1; }
        else if (std::fabs(cnum1) > std::fabs(cden)) { mul = bignum; done = false; cnum = cnum1; }
        else { mul = cnum / cden; done = true; }
        dscal(n, mul, x, 1);
        if (done) break;
    }
}

// dgecon.go:26-81 : reciprocal condition number of an LU-factored matrix
inline double gecon(Norm norm, int n, const double* lu, int lda, double anorm) {
    if (n == 0) return 1;
    if (anorm == 0) return 0;
    vec work(4 * (size_t)n, 0.0);
    ivec iwork(n, 0);
    double rcond = 0, ainvnm = 0;
    int kase = 0;
    bool normin = false;
    Lacn2State st;
    const int kase1 = (norm == Norm::One) ? 1 : 2;
    double* x = work.data();
    double* v = work.data() + n;
    double* cl = work.data() + 2 * (size_t)n;
    double* cu = work.data() + 3 * (size_t)n;
    for (;;) {
        lacn2(n, v, x, iwork.data(), ainvnm, kase, st);
        if (kase == 0) {
            if (ainvnm != 0) rcond = (1 / ainvnm) / anorm;
            return rcond;
        }
        double sl, su;
        if (kase == kase1) {
            sl = latrs(false, false, false, normin, n, lu, lda, x, cl);
            su = latrs(true, false, true, normin, n, lu, lda, x, cu);
        } else {
            su = latrs(true, true, true, normin, n, lu, lda, x, cu);
            sl = latrs(false, true, false, normin, n, lu, lda, x, cl);
        }
        double scale = sl * su;
        normin = true;
        if (scale != 1) {
            int ix = idamax(n, x, 1);
            if (scale == 0 || scale < std::fabs(x[ix]) * kSafeMin) return rcond;
            rscl(n, scale, x);
        }
    }
}

// dtrcon.go:21-82, upper / non-unit
inline double trcon_upper_nonunit(Norm norm, int n, const double* a, int lda) {
    if (n == 0) return 1;
    double rcond = 0;
    const double smlnum = kSafeMin * (double)n;
    double anorm = lantr_upper_nonunit(norm, n, a, lda);
    if (anorm <= 0) return rcond;
    vec work(3 * (size_t)n, 0.0);
    ivec iwork(n, 0);
    double* x = work.data();
    double* v = work.data() + n;
    double* cn = work.data() + 2 * (size_t)n;
    double ainvnm = 0;
    bool normin = false;
    const int kase1 = (norm == Norm::One) ? 1 : 2;
    int kase = 0;
    Lacn2State st;
    for (;;) {
        lacn2(n, v, x, iwork.data(), ainvnm, kase, st);
        if (kase == 0) {
            if (ainvnm != 0) rcond = (1 / anorm) / ainvnm;
            return rcond;
        }
        double scale = latrs(true, kase != kase1, true, normin, n, a, lda, x, cn);
        normin = true;
        if (scale != 1) {
            int ix = idamax(n, x, 1);
            double xnorm = std::fabs(x[ix]);
            if (scale == 0 || scale < xnorm * smlnum) return rcond;
            rscl(n, scale, x);
        }
    }
}

// ---------------------------------------------------------------- Householder QR (unblocked)
// dlarfg.go:27-62
inline void larfg(int n, double& alpha, double* x, int incx, double& tau) {
    if (n <= 1) { tau = 0; return; }
    double xnorm = dnrm2(n - 1, x, incx);
    if (xnorm == 0) { tau = 0; return; }
    double beta = -std::copysign(std::hypot(alpha, xnorm), alpha);
    const double safmin = kSafeMin / kEps;
    int knt = 0;
    if (std::fabs(beta) < safmin) {
        const double rsafmn = 1 / safmin;
        for (;;) {
            ++knt;
            dscal(n - 1, rsafmn, x, incx);
            beta *= rsafmn;
            alpha *= rsafmn;
            if (std::fabs(beta) >= safmin) break;
        }
        xnorm = dnrm2(n - 1, x, incx);
        beta = -std::copysign(std::hypot(alpha, xnorm), alpha);
    }
    tau = (beta - alpha) / beta;
    dscal(n - 1, 1 / (alpha - beta), x, incx);
    for (int j = 0; j < knt; ++j) beta *= safmin;
    alpha = beta;
}

// dgeqr2.go:34-59 + dlarf.go (left application, v stored down a column with stride lda).
// a is m×n row-major with m >= n; on return the upper triangle holds R.
inline void geqr2(int m, int n, double* a, int lda) {
    vec work(n, 0.0);
    const int k = std::min(m, n);
    for (int i = 0; i < k; ++i) {
        double tau;
        double& aii = a[(size_t)i * lda + i];
        larfg(m - i, aii, a + (size_t)std::min(i + 1, m - 1) * lda + i, lda, tau);
        if (i < n - 1 && tau != 0) {
            double saved = aii;
            aii = 1;
            const int rows = m - i, cols = n - i - 1;
            double* c = a + (size_t)i * lda + i + 1;
            const double* vcol = a + (size_t)i * lda + i;
            // work = C^T v (Dgemv Trans: row-wise axpy, skipping v_r == 0), then C -= tau v work^T.
            // Dlarf trims trailing zero rows of v / zero columns of C first; skipping them here is
            // value-identical because their contributions are exact zeros.
            for (int j = 0; j < cols; ++j) work[j] = 0;
            for (int r = 0; r < rows; ++r) {
                double t = vcol[(size_t)r * lda];
                if (t != 0) axpy(cols, t, c + (size_t)r * lda, 1, work.data(), 1);
            }
            for (int r = 0; r < rows; ++r) {
                double t = -tau * vcol[(size_t)r * lda];
                axpy(cols, t, work.data(), 1, c + (size_t)r * lda, 1);
            }
            aii = saved;
        }
    }
}

// ---------------------------------------------------------------- mat-level objects
struct LU {
    int n = 0;
    vec lu;
    ivec piv;
    double cond = 0;
    // mat/lu.go:63-84 (+ updateCond :28-50 with anorm >= 0)
    void factorize(const double* a, int lda, int n_, Norm norm, bool transpose_input = false) {
        n = n_;
        lu.assign((size_t)n * n, 0.0);
        piv.assign(n, 0);
        if (!transpose_input) {
            for (int i = 0; i < n; ++i) std::memcpy(&lu[(size_t)i * n], a + (size_t)i * lda, sizeof(double) * n);
        } else {
            for (int i = 0; i < n; ++i)
                for (int j = 0; j < n; ++j) lu[(size_t)i * n + j] = a[(size_t)j * lda + i];
        }
        double anorm = lange(norm, n, n, lu.data(), n);
        getrf(n, lu.data(), n, piv.data());
        cond = 1 / gecon(norm, n, lu.data(), n, anorm);
    }
    // mat/lu.go:110-134 : exp(sum log|u_ii|) * sign — underflow counts as singular
    double det() const {
        double s = 0, sign = 1;
        for (int i = 0; i < n; ++i) {
            double v = lu[(size_t)i * n + i];
            if (v < 0) sign *= -1;
            if (piv[i] != i) sign *= -1;
            s += std::log(std::fabs(v));
        }
        return std::exp(s) * sign;
    }
};

enum class SolveErr { None, DetZero, Cond };

// VecDense.SolveVec for a square system (mat/solve.go:110-140 -> :79-96 -> lu.go:293-325).
// x is written unless det()==0. `transpose` models SolveVec(a.T(), b): the transpose is
// materialised by lu.lu.Copy(a) and solved NoTrans.
inline SolveErr solve_vec(const double* a, int lda, int n, bool transpose, const double* b, double* x,
                          double* cond_out = nullptr) {
    LU f;
    f.factorize(a, lda, n, Norm::Inf, transpose);
    if (cond_out) *cond_out = f.cond;
    if (f.det() == 0) return SolveErr::DetZero;
    vec tmp(b, b + n);
    getrs(n, f.lu.data(), n, f.piv.data(), tmp.data());
    std::memcpy(x, tmp.data(), sizeof(double) * n);
    if (f.cond > 1e16) return SolveErr::Cond;
    return SolveErr::None;
}

// mat.Cond(a, 1) for an m×k matrix with m >= k (mat/matrix.go:284-322)
inline double cond1(const double* a, int lda, int m, int k) {
    if (m == k) {
        LU f;
        f.factorize(a, lda, m, Norm::One);
        return f.cond;
    }
    vec q((size_t)m * k);
    for (int i = 0; i < m; ++i) std::memcpy(&q[(size_t)i * k], a + (size_t)i * lda, sizeof(double) * k);
    geqr2(m, k, q.data(), k);
    return 1 / trcon_upper_nonunit(Norm::One, k, q.data(), k);
}

}  // namespace orc
