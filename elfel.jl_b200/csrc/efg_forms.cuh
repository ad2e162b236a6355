// efg_forms.cuh -- per-element arithmetic of the hot path, FP64, one thread per element.
//
// Everything a QPIterator precomputes (src/QPIterators.jl:14-50: N, dN/dxi per quadrature
// point, weights) lives in __constant__ memory (c_tab); the element matrix is produced
// column by column in registers.  The expressions follow the reference's operation order
// (cited per function) so that the STRICT instantiation (no FMA contraction) is bit-identical
// to the CPU oracle; the default instantiation lets nvcc contract a*b+c into DFMA.
#pragma once
#include <cstdint>
#include <utility>
#include <type_traits>

#define EFG_MAXQ 25     // largest rule: Gauss order 5 on the square (25 points); triangles go up to 13 points
// Default (non-strict) FP mode only -- the strict mode always performs the reference's operations one by one:
#ifndef TL_DIV_CORR
#define TL_DIV_CORR 0     // 1: quotients by the Jacobian determinant get a Markstein residual correction (correctly rounded in almost all cases);
                          // 0 (measured: T6 heat 3.16 -> 3.00 ms, Q4 1.73 -> 1.62 ms, full-size parity unchanged): q = n * (1/d) with a ~1-ulp reciprocal
#endif
#ifndef TL_FAST_ACC
#define TL_FAST_ACC 1     // 1: element-matrix entries are accumulated with pre-scaled factors and FMAs (2 DFMA per entry and quadrature
                          // point instead of 4 operations); differences to the reference's rounding sequence are O(1e-16) relative
#endif

#ifndef TL_HEAT_NUM
#define TL_HEAT_NUM 1     // default FP mode of the heat forms: undivided gradient numerators, one factor kappa*w/det per quadrature point
#endif

#ifndef TL_Q4_TENSOR
#define TL_Q4_TENSOR 1    // FEH1_Q4 heat, default FP mode: Jacobian columns shared between the points of the tensor-product rule
#endif

struct QTab {
    double w[EFG_MAXQ];
    double N[EFG_MAXQ][6];
    double gp[EFG_MAXQ][6][2];
    int npts, pad_;           // points of the active rule (read by the run-time-NQ instantiations, NQ_ = 0)
};

// slot 0: T3, 1: Q4, 2: T6 -- tables of the ACTIVE quadrature rule for each element kind
// (the T3 table at the T6 rule's points is the pressure basis of the Taylor-Hood pair);
// slot 3: FEH1_T3_BUBBLE (4 functions), slot 4: FEL2_T3 / FEL2_Q4 (one constant function) -- SURVEY 8f row f5.
// slot 5: FEH1_T4 (3-D): weights and N of the tetrahedron rule; its parametric gradients are the constants of
// src/FElements.jl:380-386 and live in the kernel.
#define EFG_NTAB 6
#define EFG_TAB_T4 5
__constant__ QTab c_tab[EFG_NTAB];
__constant__ double c_prm[16];

__host__ __device__ constexpr int kind_slot(int kind) { return kind == 3 ? 0 : (kind == 4 ? 1 : 2); }

// ---- arithmetic with or without contraction ---------------------------------------------------
template <bool S> __device__ __forceinline__ double fmul(double a, double b) {
    if constexpr (S) return __dmul_rn(a, b); else return a * b;
}
template <bool S> __device__ __forceinline__ double fadd(double a, double b) {
    if constexpr (S) return __dadd_rn(a, b); else return a + b;
}
template <bool S> __device__ __forceinline__ double fsub(double a, double b) {
    if constexpr (S) return __dsub_rn(a, b); else return a - b;
}
template <bool S> __device__ __forceinline__ double fdiv(double a, double b) {
    if constexpr (S) return __ddiv_rn(a, b); else return a / b;
}

// Quotients n/d with a shared divisor.  STRICT: the IEEE division the reference performs.
// Default: one ~1-ulp reciprocal per divisor, then q = n*inv refined by one FMA residual step
// (Markstein: q' = q + inv*(n - q*d) with the residual exact in the FMA) -- 3 DFMA-pipe ops per quotient
// instead of a ~25-instruction IEEE division sequence; q' is the correctly rounded quotient except in
// rare near-tie cases (then 1 ulp off), far inside the 1e-12 parity bar.
template <bool S> struct SharedDivisor {
    double d, inv;
    __device__ __forceinline__ explicit SharedDivisor(double d_) : d(d_), inv(0.0) {
#if defined(TL_FAST_RCP) && !TL_FAST_RCP
        if constexpr (!S) inv = 1.0 / d_;
#else
        if constexpr (!S) {   // MUFU.RCP64H seed (>= 20 good bits) + two Newton steps -> ~1 ulp reciprocal
            double r;
            asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(d_));
            double e = fma(-d_, r, 1.0);
            r = fma(r, e, r);
            e = fma(-d_, r, 1.0);
            inv = fma(r, e, r);
        }
#endif
    }
    __device__ __forceinline__ double operator()(double n) const {
        if constexpr (S) return __ddiv_rn(n, d);
#if TL_DIV_CORR
        else { const double q = n * inv; const double r = fma(-q, d, n); return fma(r, inv, q); }
#else
        else return n * inv;      // <= 1.5 ulp: the reciprocal is good to ~1 ulp (two Newton steps), far inside the 1e-12 parity bar
#endif
    }
};

// ---- geometry: Jacobian, JxW and spatial gradients at one quadrature point ---------------------
// GK = kind of the geometry carrier (whose nodes X,Y are given), BK = kind whose basis gradients
// are wanted (BK != GK only for the Reddy/veclap Stokes forms: T3 Jacobian applied to T6 gradients,
// examples/stokes/colliding_flow/ht_p2_p1.jl:72-76).
// BS = c_tab slot of the BK basis functions (default: the H1 element with BK nodes).
template <bool S, int GK, int BK, int BS = kind_slot(BK)>
__device__ __forceinline__ void geo_qp(const double (&X)[GK], const double (&Y)[GK], int q,
                                       double (&gx)[BK], double (&gy)[BK], double &JxW)
{
    const QTab &tg = c_tab[kind_slot(GK)];
    const QTab &tb = c_tab[BS];
    // _jac: src/FElements.jl:148-156 -- J = sum_n x_n (outer) dN_n/dxi, node order, first term assigned.
    // ALWAYS evaluated without FMA contraction: the sum cancels from O(1) to O(h), so any other
    // rounding sequence differs from the reference by O(eps/h) relative -- more than the 1e-12 /
    // 1e-14 parity bar at h = 1/4000.  With identical J the rest is well conditioned.
    double J00 = __dmul_rn(X[0], tg.gp[q][0][0]), J01 = __dmul_rn(X[0], tg.gp[q][0][1]);
    double J10 = __dmul_rn(Y[0], tg.gp[q][0][0]), J11 = __dmul_rn(Y[0], tg.gp[q][0][1]);
#pragma unroll
    for (int n = 1; n < GK; n++) {
        J00 = __dadd_rn(J00, __dmul_rn(X[n], tg.gp[q][n][0])); J01 = __dadd_rn(J01, __dmul_rn(X[n], tg.gp[q][n][1]));
        J10 = __dadd_rn(J10, __dmul_rn(Y[n], tg.gp[q][n][0])); J11 = __dadd_rn(J11, __dmul_rn(Y[n], tg.gp[q][n][1]));
    }
    // Jacobian(Val{2}): src/FElements.jl:120-129
    const double d = __dsub_rn(__dmul_rn(J00, J11), __dmul_rn(J10, J01));
    JxW = fmul<S>(d, tg.w[q]);                                // JxW = J * weight(qp)
    // bfungrad: src/QPIterators.jl:132-140 -- gradpar / Jac == (Jac' \ g)', two quotients by det per basis function
    const SharedDivisor<S> div(d);
#pragma unroll
    for (int n = 0; n < BK; n++) {
        gx[n] = div(fsub<S>(fmul<S>(J11, tb.gp[q][n][0]), fmul<S>(J10, tb.gp[q][n][1])));
        gy[n] = div(fsub<S>(fmul<S>(J00, tb.gp[q][n][1]), fmul<S>(J01, tb.gp[q][n][0])));
    }
}

// Default FP mode of the heat forms: the UNDIVIDED gradient numerators (J11*g0 - J10*g1, J00*g1 - J01*g0) and the determinant.
// With c = kappa * w / det the entry sum_q dot(gradN_i, gradN_j) * kappa * JxW becomes sum_q (nx_i*c)*nx_j + (ny_i*c)*ny_j:
// the per-gradient quotients by det disappear (8-12 DMUL per quadrature point).  The Jacobian itself is still the
// uncontracted node-order sum (see geo_qp); differences to the reference's rounding sequence stay O(1e-16) relative.
template <int GK, int BK, int BS = kind_slot(BK)>
__device__ __forceinline__ void geo_qp_num(const double (&X)[GK], const double (&Y)[GK], int q,
                                           double (&nx)[BK], double (&ny)[BK], double &det)
{
    const QTab &tg = c_tab[kind_slot(GK)];
    const QTab &tb = c_tab[BS];
    double J00 = __dmul_rn(X[0], tg.gp[q][0][0]), J01 = __dmul_rn(X[0], tg.gp[q][0][1]);
    double J10 = __dmul_rn(Y[0], tg.gp[q][0][0]), J11 = __dmul_rn(Y[0], tg.gp[q][0][1]);
#pragma unroll
    for (int n = 1; n < GK; n++) {
        J00 = __dadd_rn(J00, __dmul_rn(X[n], tg.gp[q][n][0])); J01 = __dadd_rn(J01, __dmul_rn(X[n], tg.gp[q][n][1]));
        J10 = __dadd_rn(J10, __dmul_rn(Y[n], tg.gp[q][n][0])); J11 = __dadd_rn(J11, __dmul_rn(Y[n], tg.gp[q][n][1]));
    }
    det = __dsub_rn(__dmul_rn(J00, J11), __dmul_rn(J10, J01));
#pragma unroll
    for (int n = 0; n < BK; n++) {
        nx[n] = J11 * tb.gp[q][n][0] - J10 * tb.gp[q][n][1];
        ny[n] = J00 * tb.gp[q][n][1] - J01 * tb.gp[q][n][0];
    }
}

template <int BK, int NQ> struct Geo {
    double gx[NQ][BK], gy[NQ][BK];
    double JxW[NQ];
};

template <bool S, int GK, int BK, int NQ, int BS = kind_slot(BK)>
__device__ __forceinline__ void geo_compute(const double (&X)[GK], const double (&Y)[GK], Geo<BK, NQ> &G)
{
#pragma unroll
    for (int q = 0; q < NQ; q++) geo_qp<S, GK, BK, BS>(X, Y, q, G.gx[q], G.gy[q], G.JxW[q]);
}
// table slot of a form's basis functions: F::BSLOT if the form names one, else the H1 element with F::BK nodes
template <class F, class = void> struct form_bslot { static constexpr int value = kind_slot(F::BK); };
template <class F> struct form_bslot<F, std::void_t<decltype(F::BSLOT)>> { static constexpr int value = F::BSLOT; };

// does the sink of an element worker also take the element load vector (StageEmit<F, true>)?
template <class E, class = void> struct emit_with_f { static constexpr bool value = false; };
template <class E> struct emit_with_f<E, std::void_t<decltype(E::WITH_F)>> { static constexpr bool value = E::WITH_F; };

// Generic element driver: all quadrature points' gradients in registers, then the owned columns one
// by one.  emit.template col<J>(out) receives column J of the element matrix.
template <class F, bool S, class Emit, int J>
__device__ __forceinline__ void element_column(const Geo<F::BK, F::NQ> &G, uint32_t m, Emit &emit)
{
    if (m & (1u << J)) {
        double out[F::ND];
        F::template column<S, J>(G, out);
        emit.template col<J>(out);
    }
}
template <class F, bool S, class Emit, int... Js>
__device__ __forceinline__ void element_columns(std::integer_sequence<int, Js...>, const Geo<F::BK, F::NQ> &G, uint32_t m, Emit &emit)
{
    (element_column<F, S, Emit, Js>(G, m, emit), ...);
}
template <class F, bool S, class Emit>
__device__ __forceinline__ void element_generic(const double (&X)[F::GK], const double (&Y)[F::GK], uint32_t m, Emit &emit)
{
    Geo<F::BK, F::NQ> G;
    geo_compute<S, F::GK, F::BK, F::NQ, form_bslot<F>::value>(X, Y, G);
    element_columns<F, S, Emit>(std::make_integer_sequence<int, F::ND>{}, G, m, emit);
}

// where the dof numbers of an element come from (symbolic phase only)
struct DofSrc {
    const int32_t *conn0, *conn1;          // 0-based node ids of mesh 0 / mesh 1
    const int32_t *dof0, *dof1, *dof2;     // 0-based dof numbers (ncomp x nnodes), -1 = dof number 0
    const int32_t *cdof0, *cdof1, *cdof2;  // 0-based dof numbers of the spaces' cell fields (ncomp x nel), null = no cell field
};

// A form provides:
//   ND      local dofs of the combined element matrix (rows == columns)
//   NT      COO triplets the reference appends per element
//   GK,BK   geometry carrier kind / basis kind, GMESH = mesh slot of the geometry carrier
//   mask(i,j)   is local entry (i,j) appended?       kidx(i,j)  its position in append order
//   edofs()     combined element dof vector (0-based)
//   column<S,J>()  column J of the element matrix, all quadrature points summed in order
//   element<S>(X, Y, mask, emit)  whole element: emits the columns whose mask bit is set

// B(g,k) and D*B of the elasticity / Stokes-gen kernels (examples/elasticity/stretch/t6.jl:42-58):
// B(g,1) = (g1, 0, g2), B(g,2) = (0, g2, g1); the literal zero products are dropped (x + 0*D == x).
template <bool S> __device__ __forceinline__ void DB(const double *D, int comp, double gx, double gy, double (&o)[3])
{
#pragma unroll
    for (int r = 0; r < 3; r++)
        o[r] = comp == 0 ? fadd<S>(fmul<S>(D[0 * 3 + r], gx), fmul<S>(D[2 * 3 + r], gy))
                         : fadd<S>(fmul<S>(D[1 * 3 + r], gy), fmul<S>(D[2 * 3 + r], gx));
}
template <bool S> __device__ __forceinline__ double dotB(const double (&db)[3], int comp, double gx, double gy)
{
    return comp == 0 ? fadd<S>(fmul<S>(db[0], gx), fmul<S>(db[2], gy))
                     : fadd<S>(fmul<S>(db[1], gy), fmul<S>(db[2], gx));
}


// One entry of a B'DB element matrix, all quadrature points: row dof of component ci with gradients (ax, ay), column given
// by db[q] = D*B_j (fast mode: already scaled by JxW).  Strict mode = the operation sequence of column_rt.
template <bool S, int NQ>
__device__ __forceinline__ double bdb_entry(const double (&db)[NQ][3], int ci, const double (&ax)[NQ], const double (&ay)[NQ], const double (&jw)[NQ])
{
    double acc = 0.0;
    if constexpr (!S && TL_FAST_ACC) {
#pragma unroll
        for (int q = 0; q < NQ; q++) {
            const double a = ci == 0 ? db[q][0] : db[q][1];
            const double g1 = ci == 0 ? ax[q] : ay[q], g2 = ci == 0 ? ay[q] : ax[q];
            acc = q == 0 ? fma(a, g1, db[q][2] * g2) : fma(a, g1, fma(db[q][2], g2, acc));
        }
    } else {
#pragma unroll
        for (int q = 0; q < NQ; q++) {
            const double t = fmul<S>(dotB<S>(db[q], ci, ax[q], ay[q]), jw[q]);
            acc = q == 0 ? t : fadd<S>(acc, t);
        }
    }
    return acc;
}

// ------------------------------------------------------------------------------------------------
// K1 heat: ke[i,j] += dot(gradN[i], gradN[j]) * (kappa * JxW)   examples/heat/poisson/t3.jl:53-58
template <int VK, int NQ_> struct HeatForm {
    static constexpr int ND = VK, NT = VK * VK, GK = VK, BK = VK, NQ = NQ_, GMESH = 0, NSPACES = 1;
    static constexpr bool SPLIT = false;
    static constexpr bool SYM = true;       // the element matrix is bitwise symmetric: a sink may take the upper triangle only (emit.tri)
    __host__ __device__ static constexpr bool mask(int, int) { return true; }
    __host__ __device__ static constexpr int kidx(int i, int j) { return j * ND + i; }
    __device__ static void edofs(const DofSrc &s, int64_t e, int32_t (&d)[ND]) {
#pragma unroll
        for (int a = 0; a < VK; a++) d[a] = s.dof0[s.conn0[e * VK + a]];
    }
    // The element matrix is bitwise symmetric (g_i.g_j: the two products commute, same sum order), so
    // only the upper triangle is accumulated, quadrature point by quadrature point like the reference's
    // loop; gradients of one point at a time keep the register footprint small.
    template <class Emit, int J> __device__ __forceinline__ static void emit_col(const double (&K)[ND][ND], uint32_t m, Emit &emit) {
        if (m & (1u << J)) {
            double out[ND];
#pragma unroll
            for (int i = 0; i < ND; i++) out[i] = (i <= J) ? K[i][J] : K[J][i];
            emit.template col<J>(out);
        }
    }
    template <class Emit, int... Js> __device__ __forceinline__ static void emit_cols(std::integer_sequence<int, Js...>, const double (&K)[ND][ND], uint32_t m, Emit &emit) {
        (emit_col<Emit, Js>(K, m, emit), ...);
    }
    template <bool S, class Emit>
    __device__ __forceinline__ static void element(const double (&X)[GK], const double (&Y)[GK], uint32_t m, Emit &emit) {
        const double kappa = c_prm[0];
        double K[ND][ND];
        // FEH1_Q4 on the tensor-product Gauss rule (points i outer / j inner, src/RefShapes.jl:350-362), default FP mode: dN/dxi
        // depends on eta only and dN/deta on xi only, so the first Jacobian column takes NP distinct values (one per j) and the
        // second NP (one per i) instead of NP^2 each -- the same individually rounded node-order sums, computed once
        // NQ_ = 0: the number of quadrature points is read from the table at run time (the less common rules: triangles with
        // 4 / 6 / 7 / 9 / 12 / 13 points, Gauss orders 4 and 5 on the square -- src/RefShapes.jl:120-230, 85-110); same arithmetic,
        // the loop is not unrolled
        const int nq = NQ ? NQ : c_tab[kind_slot(GK)].npts;
        constexpr bool Q4T = !S && TL_FAST_ACC && TL_HEAT_NUM && TL_Q4_TENSOR && VK == 4 && (NQ == 4 || NQ == 9);
        constexpr int NP = NQ == 9 ? 3 : 2;
        double A00[NP], A10[NP], B01[NP], B11[NP];
        if constexpr (Q4T) {
            const QTab &tg = c_tab[kind_slot(GK)];
            {
                {
#pragma unroll
                    for (int t = 0; t < NP; t++) {
                        const int qa = t, qb = t * NP;          // (i = 0, j = t) and (i = t, j = 0)
                        double a0 = __dmul_rn(X[0], tg.gp[qa][0][0]), a1 = __dmul_rn(Y[0], tg.gp[qa][0][0]);
                        double b0 = __dmul_rn(X[0], tg.gp[qb][0][1]), b1 = __dmul_rn(Y[0], tg.gp[qb][0][1]);
#pragma unroll
                        for (int n = 1; n < GK; n++) {
                            a0 = __dadd_rn(a0, __dmul_rn(X[n], tg.gp[qa][n][0])); a1 = __dadd_rn(a1, __dmul_rn(Y[n], tg.gp[qa][n][0]));
                            b0 = __dadd_rn(b0, __dmul_rn(X[n], tg.gp[qb][n][1])); b1 = __dadd_rn(b1, __dmul_rn(Y[n], tg.gp[qb][n][1]));
                        }
                        A00[t] = a0; A10[t] = a1; B01[t] = b0; B11[t] = b1;
                    }
                }
            }
        }
#pragma unroll
        for (int q = 0; q < nq; q++) {
            if constexpr (Q4T) {
                const QTab &tg = c_tab[kind_slot(GK)];
                const double J00 = A00[q % NP], J10 = A10[q % NP], J01 = B01[q / NP], J11 = B11[q / NP];
                const double det = __dsub_rn(__dmul_rn(J00, J11), __dmul_rn(J10, J01));
                double nx[BK], ny[BK];
#pragma unroll
                for (int n = 0; n < BK; n++) {
                    nx[n] = J11 * tg.gp[q][n][0] - J10 * tg.gp[q][n][1];
                    ny[n] = J00 * tg.gp[q][n][1] - J01 * tg.gp[q][n][0];
                }
                const SharedDivisor<false> div(det);
                const double c = div(kappa * tg.w[q]);
                double sx[BK], sy[BK];
#pragma unroll
                for (int i = 0; i < ND; i++) { sx[i] = nx[i] * c; sy[i] = ny[i] * c; }
#pragma unroll
                for (int j = 0; j < ND; j++)
#pragma unroll
                    for (int i = 0; i <= j; i++)
                        K[i][j] = q == 0 ? fma(sx[i], nx[j], sy[i] * ny[j]) : fma(sx[i], nx[j], fma(sy[i], ny[j], K[i][j]));
                continue;
            }
            if constexpr (!S && TL_FAST_ACC && TL_HEAT_NUM) {
                double nx[BK], ny[BK], det;
                geo_qp_num<GK, BK>(X, Y, q, nx, ny, det);
                const SharedDivisor<false> div(det);
                const double c = div(kappa * c_tab[kind_slot(GK)].w[q]);          // kappa * JxW / det^2
                double sx[BK], sy[BK];
#pragma unroll
                for (int i = 0; i < ND; i++) { sx[i] = nx[i] * c; sy[i] = ny[i] * c; }
#pragma unroll
                for (int j = 0; j < ND; j++)
#pragma unroll
                    for (int i = 0; i <= j; i++)
                        K[i][j] = q == 0 ? fma(sx[i], nx[j], sy[i] * ny[j]) : fma(sx[i], nx[j], fma(sy[i], ny[j], K[i][j]));
                continue;
            }
            double gx[BK], gy[BK], JxW;
            geo_qp<S, GK, BK>(X, Y, q, gx, gy, JxW);
            const double kJ = fmul<S>(kappa, JxW);
            if constexpr (!S && TL_FAST_ACC) {
                double sx[BK], sy[BK];
#pragma unroll
                for (int i = 0; i < ND; i++) { sx[i] = gx[i] * kJ; sy[i] = gy[i] * kJ; }
#pragma unroll
                for (int j = 0; j < ND; j++)
#pragma unroll
                    for (int i = 0; i <= j; i++)
                        K[i][j] = q == 0 ? fma(sx[i], gx[j], sy[i] * gy[j]) : fma(sx[i], gx[j], fma(sy[i], gy[j], K[i][j]));
            } else {
#pragma unroll
                for (int j = 0; j < ND; j++)
#pragma unroll
                    for (int i = 0; i <= j; i++) {
                        const double t = fmul<S>(fadd<S>(fmul<S>(gx[i], gx[j]), fmul<S>(gy[i], gy[j])), kJ);
                        K[i][j] = q == 0 ? t : fadd<S>(K[i][j], t);
                    }
            }
        }
        if constexpr (Emit::TRI) emit.tri(K, m);
        else emit_cols<Emit>(std::make_integer_sequence<int, ND>{}, K, m, emit);
        // fused load vector (examples/heat/poisson/t3.jl:57: fe[j] += N[j]*Q*JxW in the same quadrature loop; Q = c_prm[1]):
        // a second, cheap pass over the quadrature points (Jacobian determinant only), every operation individually rounded
        // in the reference's order in BOTH FP modes, so the vector is bit-identical to the CPU loop like efg_vec_assemble's
        if constexpr (emit_with_f<Emit>::value) {
            const QTab &tg = c_tab[kind_slot(GK)];
            const double Q = c_prm[1];
            double f[ND];
#pragma unroll
            for (int j = 0; j < ND; j++) f[j] = 0.0;
#pragma unroll
            for (int q = 0; q < nq; q++) {
                double J00 = __dmul_rn(X[0], tg.gp[q][0][0]), J01 = __dmul_rn(X[0], tg.gp[q][0][1]);
                double J10 = __dmul_rn(Y[0], tg.gp[q][0][0]), J11 = __dmul_rn(Y[0], tg.gp[q][0][1]);
#pragma unroll
                for (int n = 1; n < GK; n++) {
                    J00 = __dadd_rn(J00, __dmul_rn(X[n], tg.gp[q][n][0])); J01 = __dadd_rn(J01, __dmul_rn(X[n], tg.gp[q][n][1]));
                    J10 = __dadd_rn(J10, __dmul_rn(Y[n], tg.gp[q][n][0])); J11 = __dadd_rn(J11, __dmul_rn(Y[n], tg.gp[q][n][1]));
                }
                const double JxW = __dmul_rn(__dsub_rn(__dmul_rn(J00, J11), __dmul_rn(J10, J01)), tg.w[q]);
#pragma unroll
                for (int j = 0; j < ND; j++) f[j] = __dadd_rn(f[j], __dmul_rn(__dmul_rn(tg.N[q][j], Q), JxW));
            }
            emit.fvec(f, m);
        }
    }
};

// K2 elasticity: ke[i,j] += dot(D*B_j, B_i) * JxW               examples/elasticity/stretch/t6.jl:52-58
template <int VK, int NQ_> struct ElasticityForm {
    static constexpr bool SYM = false;
    static constexpr int ND = 2 * VK, NT = ND * ND, GK = VK, BK = VK, NQ = NQ_, GMESH = 0, NSPACES = 1;
    template <bool S, class Emit>
    __device__ __forceinline__ static void element(const double (&X)[GK], const double (&Y)[GK], uint32_t m, Emit &emit) {
        element_generic<ElasticityForm<VK, NQ_>, S, Emit>(X, Y, m, emit);
    }
    __host__ __device__ static constexpr bool mask(int, int) { return true; }
    __host__ __device__ static constexpr int kidx(int i, int j) { return j * ND + i; }
    __device__ static void edofs(const DofSrc &s, int64_t e, int32_t (&d)[ND]) {
#pragma unroll
        for (int a = 0; a < VK; a++) {
            const int64_t n = s.conn0[e * VK + a];
            d[2 * a] = s.dof0[2 * n]; d[2 * a + 1] = s.dof0[2 * n + 1];
        }
    }
    // SPLIT forms: phase 1 runs as (a) one thread per element: geometry -> shared memory, (b) one thread per
    // staged column: column_rt() with the column index known only at run time (gjx/gjy = gradients of the
    // column's node).  column<S,J>() is the same arithmetic with J known at compile time.
    static constexpr bool SPLIT = true;
    __device__ __forceinline__ static int colnode(int J) { return J >> 1; }
    template <bool S> __device__ __forceinline__ static void column_rt(const Geo<BK, NQ> &G, int J, const double (&gjx)[NQ],
                                                                        const double (&gjy)[NQ], double (&out)[ND]) {
        const int cj = J & 1;
        double db[NQ][3];
#pragma unroll
        for (int q = 0; q < NQ; q++) DB<S>(c_prm, cj, gjx[q], gjy[q], db[q]);
        if constexpr (!S && TL_FAST_ACC) {      // D*B_j scaled by JxW once, then two FMAs per entry and quadrature point
#pragma unroll
            for (int q = 0; q < NQ; q++)
#pragma unroll
                for (int r = 0; r < 3; r++) db[q][r] *= G.JxW[q];
#pragma unroll
            for (int i = 0; i < ND; i++) {
                double acc = 0.0;
#pragma unroll
                for (int q = 0; q < NQ; q++) {
                    const double a = (i % 2 == 0) ? db[q][0] : db[q][1];        // B_i = (g1, 0, g2) or (0, g2, g1)
                    const double g1 = (i % 2 == 0) ? G.gx[q][i / 2] : G.gy[q][i / 2], g2 = (i % 2 == 0) ? G.gy[q][i / 2] : G.gx[q][i / 2];
                    acc = q == 0 ? fma(a, g1, db[q][2] * g2) : fma(a, g1, fma(db[q][2], g2, acc));
                }
                out[i] = acc;
            }
            return;
        }
#pragma unroll
        for (int i = 0; i < ND; i++) {
            double acc = 0.0;
#pragma unroll
            for (int q = 0; q < NQ; q++) {
                const double t = fmul<S>(dotB<S>(db[q], i % 2, G.gx[q][i / 2], G.gy[q][i / 2]), G.JxW[q]);
                acc = q == 0 ? t : fadd<S>(acc, t);
            }
            out[i] = acc;
        }
    }
    template <bool S, int J> __device__ __forceinline__ static void column(const Geo<BK, NQ> &G, double (&out)[ND]) {
        double gjx[NQ], gjy[NQ];
#pragma unroll
        for (int q = 0; q < NQ; q++) { gjx[q] = G.gx[q][J / 2]; gjy[q] = G.gy[q][J / 2]; }
        column_rt<S>(G, J, gjx, gjy, out);
    }
    // BOTH columns of node nb (local dofs 2nb, 2nb+1) in one sweep over the row nodes: the row gradients are read once for
    // two columns.  g(k) reads value k of the element's geometry record (gx[q][n] at q*BK+n, gy at NQ*BK + q*BK+n, JxW at
    // 2*NQ*BK + q), put(i, v0, v1) receives row i of the two columns.  Same arithmetic as column_rt, entry by entry.
    __device__ __forceinline__ static bool pairable(int J0) { return (J0 & 1) == 0; }
    // one column, same streaming sweep (put(i, v))
    template <bool S, class Load, class Put>
    __device__ __forceinline__ static void column_single_rt(Load &&g, int J, Put &&put) {
        const int nb = J >> 1, cj = J & 1;
        double db[NQ][3], jw[NQ];
#pragma unroll
        for (int q = 0; q < NQ; q++) {
            jw[q] = g(2 * NQ * BK + q);
            DB<S>(c_prm, cj, g(q * BK + nb), g(NQ * BK + q * BK + nb), db[q]);
            if constexpr (!S && TL_FAST_ACC) {
#pragma unroll
                for (int r = 0; r < 3; r++) db[q][r] *= jw[q];
            }
        }
#pragma unroll
        for (int a = 0; a < VK; a++) {
            double ax[NQ], ay[NQ];
#pragma unroll
            for (int q = 0; q < NQ; q++) { ax[q] = g(q * BK + a); ay[q] = g(NQ * BK + q * BK + a); }
            put(2 * a, bdb_entry<S, NQ>(db, 0, ax, ay, jw));
            put(2 * a + 1, bdb_entry<S, NQ>(db, 1, ax, ay, jw));
        }
    }
    template <bool S, class Load, class Put>
    __device__ __forceinline__ static void column_pair_rt(Load &&g, int nb, Put &&put) {
        double db0[NQ][3], db1[NQ][3], jw[NQ];
#pragma unroll
        for (int q = 0; q < NQ; q++) {
            jw[q] = g(2 * NQ * BK + q);
            const double gjx = g(q * BK + nb), gjy = g(NQ * BK + q * BK + nb);
            DB<S>(c_prm, 0, gjx, gjy, db0[q]);
            DB<S>(c_prm, 1, gjx, gjy, db1[q]);
            if constexpr (!S && TL_FAST_ACC) {
#pragma unroll
                for (int r = 0; r < 3; r++) { db0[q][r] *= jw[q]; db1[q][r] *= jw[q]; }
            }
        }
#pragma unroll
        for (int a = 0; a < VK; a++) {
            double ax[NQ], ay[NQ];
#pragma unroll
            for (int q = 0; q < NQ; q++) { ax[q] = g(q * BK + a); ay[q] = g(NQ * BK + q * BK + a); }
            put(2 * a, bdb_entry<S, NQ>(db0, 0, ax, ay, jw), bdb_entry<S, NQ>(db1, 0, ax, ay, jw));
            put(2 * a + 1, bdb_entry<S, NQ>(db0, 1, ax, ay, jw), bdb_entry<S, NQ>(db1, 1, ax, ay, jw));
        }
    }
};

// Taylor-Hood T6/T3 Stokes forms.  Combined local dofs of the "2-space" forms (gen, veclap_alt):
// 0..11 = velocity (node-major, component-minor), 12..14 = pressure.  Of the "3-space" forms
// (Reddy, veclap): 0..5 = ux, 6..11 = uy, 12..14 = p.  The p-p block is never appended.
//
// K3 gen: kuu += dot(B_i, D*B_j)*JxW; kup[i,j] += (-JxW*Np[j]) * gradNu[i][c[i]]; assemble kuu, kup, kup'
//        examples/stokes/colliding_flow/ht_p2_p1_gen.jl:58-78
// K5 veclap_alt: kuu[i,j] += (mu*JxW)*dot(g_i,g_j) only where c[i]==c[j] (others stay explicit zeros)
//        examples/stokes/colliding_flow/ht_p2_p1_veclap_alt.jl:71-87
template <bool VECLAP_ALT> struct Stokes2Form {
    static constexpr bool SYM = false;
    static constexpr int VK = 6, PK = 3, NQ = 3;
    static constexpr int ND = 15, NT = 144 + 36 + 36, GK = 6, BK = 6, GMESH = 0, NSPACES = 2;
    template <bool S, class Emit>
    __device__ __forceinline__ static void element(const double (&X)[GK], const double (&Y)[GK], uint32_t m, Emit &emit) {
        element_generic<Stokes2Form<VECLAP_ALT>, S, Emit>(X, Y, m, emit);
    }
    __host__ __device__ static constexpr bool mask(int i, int j) { return !(i >= 12 && j >= 12); }
    __host__ __device__ static constexpr int kidx(int i, int j) {
        return (i < 12 && j < 12) ? j * 12 + i                       // assemble!(ass, kuu)
             : (i < 12)           ? 144 + (j - 12) * 12 + i          // assemble!(ass, kup)
                                  : 180 + j * 3 + (i - 12);          // assemble!(ass, transpose(kup))
    }
    __device__ static void edofs(const DofSrc &s, int64_t e, int32_t (&d)[ND]) {
#pragma unroll
        for (int a = 0; a < 6; a++) {
            const int64_t n = s.conn0[e * 6 + a];
            d[2 * a] = s.dof0[2 * n]; d[2 * a + 1] = s.dof0[2 * n + 1];
        }
#pragma unroll
        for (int m = 0; m < 3; m++) d[12 + m] = s.dof1[s.conn1[e * 3 + m]];
    }
    static constexpr bool SPLIT = true;
    static constexpr int PAIR_COLS = 12;     // columns 0..11 pair up by node, 12..14 (pressure) are single
    __device__ __forceinline__ static int colnode(int J) { return J < 12 ? (J >> 1) : 0; }
    template <bool S> __device__ __forceinline__ static void column_rt(const Geo<6, 3> &G, int J, const double (&gjx)[3],
                                                                        const double (&gjy)[3], double (&out)[ND]) {
        const QTab &tp = c_tab[kind_slot(3)];
        if (J < 12) {
            const int cj = J & 1;
            if constexpr (!VECLAP_ALT) {
                double db[3][3];
#pragma unroll
                for (int q = 0; q < 3; q++) DB<S>(c_prm, cj, gjx[q], gjy[q], db[q]);
                if constexpr (!S && TL_FAST_ACC) {      // (see ElasticityForm::column_rt)
#pragma unroll
                    for (int q = 0; q < 3; q++)
#pragma unroll
                        for (int r = 0; r < 3; r++) db[q][r] *= G.JxW[q];
#pragma unroll
                    for (int i = 0; i < 12; i++) {
                        double acc = 0.0;
#pragma unroll
                        for (int q = 0; q < 3; q++) {
                            const double a = (i % 2 == 0) ? db[q][0] : db[q][1];
                            const double g1 = (i % 2 == 0) ? G.gx[q][i / 2] : G.gy[q][i / 2], g2 = (i % 2 == 0) ? G.gy[q][i / 2] : G.gx[q][i / 2];
                            acc = q == 0 ? fma(a, g1, db[q][2] * g2) : fma(a, g1, fma(db[q][2], g2, acc));
                        }
                        out[i] = acc;
                    }
                } else {
#pragma unroll
                    for (int i = 0; i < 12; i++) {
                        double acc = 0.0;
#pragma unroll
                        for (int q = 0; q < 3; q++) {
                            const double t = fmul<S>(dotB<S>(db[q], i % 2, G.gx[q][i / 2], G.gy[q][i / 2]), G.JxW[q]);
                            acc = q == 0 ? t : fadd<S>(acc, t);
                        }
                        out[i] = acc;
                    }
                }
            } else {
                const double mu = c_prm[0];
#pragma unroll
                for (int i = 0; i < 12; i++) {
                    double acc = 0.0;
                    if (i % 2 == cj) {
#pragma unroll
                        for (int q = 0; q < 3; q++) {
                            const double t = fmul<S>(fmul<S>(mu, G.JxW[q]),
                                                     fadd<S>(fmul<S>(G.gx[q][i / 2], gjx[q]), fmul<S>(G.gy[q][i / 2], gjy[q])));
                            acc = q == 0 ? t : fadd<S>(acc, t);
                        }
                    }
                    out[i] = acc;
                }
            }
            // rows 12..14 = transpose(kup): value kup[J, m]
#pragma unroll
            for (int m = 0; m < 3; m++) {
                double acc = 0.0;
#pragma unroll
                for (int q = 0; q < 3; q++) {
                    const double gj = (cj == 0) ? gjx[q] : gjy[q];
                    const double t = fmul<S>(fmul<S>(-G.JxW[q], tp.N[q][m]), gj);
                    acc = q == 0 ? t : fadd<S>(acc, t);
                }
                out[12 + m] = acc;
            }
        } else {
            const int m = J - 12;
#pragma unroll
            for (int i = 0; i < 12; i++) {
                double acc = 0.0;
#pragma unroll
                for (int q = 0; q < 3; q++) {
                    const double gi = (i % 2 == 0) ? G.gx[q][i / 2] : G.gy[q][i / 2];
                    const double t = fmul<S>(fmul<S>(-G.JxW[q], tp.N[q][m]), gi);
                    acc = q == 0 ? t : fadd<S>(acc, t);
                }
                out[i] = acc;
            }
            out[12] = out[13] = out[14] = 0.0;
        }
    }
    template <bool S, int J> __device__ __forceinline__ static void column(const Geo<6, 3> &G, double (&out)[ND]) {
        double gjx[3], gjy[3];
#pragma unroll
        for (int q = 0; q < 3; q++) { gjx[q] = G.gx[q][J < 12 ? J / 2 : 0]; gjy[q] = G.gy[q][J < 12 ? J / 2 : 0]; }
        column_rt<S>(G, J, gjx, gjy, out);
    }
    // both velocity columns of node nb in one sweep (see ElasticityForm::column_pair_rt); pressure columns stay single
    __device__ __forceinline__ static bool pairable(int J0) { return (J0 & 1) == 0 && J0 < 12; }
    // one column, streaming sweep over the row nodes (put(i, v)); rows 12..14 of a pressure column are never appended
    template <bool S, class Load, class Put>
    __device__ __forceinline__ static void column_single_rt(Load &&g, int J, Put &&put) {
        constexpr int NQ_ = 3, BK_ = 6;
        const QTab &tp = c_tab[kind_slot(3)];
        double jw[NQ_];
#pragma unroll
        for (int q = 0; q < NQ_; q++) jw[q] = g(2 * NQ_ * BK_ + q);
        if (J >= 12) {                          // column of pressure dof m: kup[i, m] = sum_q (-JxW*Np[m]) * gradNu[i][c[i]]
            const int m = J - 12;
            double w[NQ_];
#pragma unroll
            for (int q = 0; q < NQ_; q++) w[q] = fmul<S>(-jw[q], tp.N[q][m]);
#pragma unroll
            for (int a = 0; a < 6; a++) {
                double a0 = 0.0, a1 = 0.0;
#pragma unroll
                for (int q = 0; q < NQ_; q++) {
                    const double t0 = fmul<S>(w[q], g(q * BK_ + a)), t1 = fmul<S>(w[q], g(NQ_ * BK_ + q * BK_ + a));
                    a0 = q == 0 ? t0 : fadd<S>(a0, t0);
                    a1 = q == 0 ? t1 : fadd<S>(a1, t1);
                }
                put(2 * a, a0);
                put(2 * a + 1, a1);
            }
            return;
        }
        const int nb = J >> 1, cj = J & 1;
        double gjx[NQ_], gjy[NQ_];
#pragma unroll
        for (int q = 0; q < NQ_; q++) { gjx[q] = g(q * BK_ + nb); gjy[q] = g(NQ_ * BK_ + q * BK_ + nb); }
        if constexpr (!VECLAP_ALT) {
            double db[NQ_][3];
#pragma unroll
            for (int q = 0; q < NQ_; q++) {
                DB<S>(c_prm, cj, gjx[q], gjy[q], db[q]);
                if constexpr (!S && TL_FAST_ACC) {
#pragma unroll
                    for (int r = 0; r < 3; r++) db[q][r] *= jw[q];
                }
            }
#pragma unroll
            for (int a = 0; a < 6; a++) {
                double ax[NQ_], ay[NQ_];
#pragma unroll
                for (int q = 0; q < NQ_; q++) { ax[q] = g(q * BK_ + a); ay[q] = g(NQ_ * BK_ + q * BK_ + a); }
                put(2 * a, bdb_entry<S, NQ_>(db, 0, ax, ay, jw));
                put(2 * a + 1, bdb_entry<S, NQ_>(db, 1, ax, ay, jw));
            }
        } else {
            const double mu = c_prm[0];
#pragma unroll
            for (int a = 0; a < 6; a++) {
                double acc = 0.0;
#pragma unroll
                for (int q = 0; q < NQ_; q++) {
                    const double t = fmul<S>(fmul<S>(mu, jw[q]), fadd<S>(fmul<S>(g(q * BK_ + a), gjx[q]), fmul<S>(g(NQ_ * BK_ + q * BK_ + a), gjy[q])));
                    acc = q == 0 ? t : fadd<S>(acc, t);
                }
                put(2 * a, cj == 0 ? acc : 0.0);
                put(2 * a + 1, cj == 1 ? acc : 0.0);
            }
        }
#pragma unroll
        for (int m = 0; m < 3; m++) {
            double acc = 0.0;
#pragma unroll
            for (int q = 0; q < NQ_; q++) {
                const double t = fmul<S>(fmul<S>(-jw[q], tp.N[q][m]), cj == 0 ? gjx[q] : gjy[q]);
                acc = q == 0 ? t : fadd<S>(acc, t);
            }
            put(12 + m, acc);
        }
    }
    template <bool S, class Load, class Put>
    __device__ __forceinline__ static void column_pair_rt(Load &&g, int nb, Put &&put) {
        constexpr int NQ_ = 3, BK_ = 6;
        const QTab &tp = c_tab[kind_slot(3)];
        double jw[NQ_], gjx[NQ_], gjy[NQ_];
#pragma unroll
        for (int q = 0; q < NQ_; q++) { jw[q] = g(2 * NQ_ * BK_ + q); gjx[q] = g(q * BK_ + nb); gjy[q] = g(NQ_ * BK_ + q * BK_ + nb); }
        if constexpr (!VECLAP_ALT) {
            double db0[NQ_][3], db1[NQ_][3];
#pragma unroll
            for (int q = 0; q < NQ_; q++) {
                DB<S>(c_prm, 0, gjx[q], gjy[q], db0[q]);
                DB<S>(c_prm, 1, gjx[q], gjy[q], db1[q]);
                if constexpr (!S && TL_FAST_ACC) {
#pragma unroll
                    for (int r = 0; r < 3; r++) { db0[q][r] *= jw[q]; db1[q][r] *= jw[q]; }
                }
            }
#pragma unroll
            for (int a = 0; a < 6; a++) {
                double ax[NQ_], ay[NQ_];
#pragma unroll
                for (int q = 0; q < NQ_; q++) { ax[q] = g(q * BK_ + a); ay[q] = g(NQ_ * BK_ + q * BK_ + a); }
                put(2 * a, bdb_entry<S, NQ_>(db0, 0, ax, ay, jw), bdb_entry<S, NQ_>(db1, 0, ax, ay, jw));
                put(2 * a + 1, bdb_entry<S, NQ_>(db0, 1, ax, ay, jw), bdb_entry<S, NQ_>(db1, 1, ax, ay, jw));
            }
        } else {
            const double mu = c_prm[0];
#pragma unroll
            for (int a = 0; a < 6; a++) {
                double acc = 0.0;
#pragma unroll
                for (int q = 0; q < NQ_; q++) {
                    const double t = fmul<S>(fmul<S>(mu, jw[q]), fadd<S>(fmul<S>(g(q * BK_ + a), gjx[q]), fmul<S>(g(NQ_ * BK_ + q * BK_ + a), gjy[q])));
                    acc = q == 0 ? t : fadd<S>(acc, t);
                }
                put(2 * a, acc, 0.0);          // entries with c[i] != c[j] stay explicit zeros
                put(2 * a + 1, 0.0, acc);
            }
        }
#pragma unroll
        for (int m = 0; m < 3; m++) {          // rows 12..14 = transpose(kup): kup[J, m]
            double a0 = 0.0, a1 = 0.0;
#pragma unroll
            for (int q = 0; q < NQ_; q++) {
                const double w = fmul<S>(-jw[q], tp.N[q][m]);
                const double t0 = fmul<S>(w, gjx[q]), t1 = fmul<S>(w, gjy[q]);
                a0 = q == 0 ? t0 : fadd<S>(a0, t0);
                a1 = q == 0 ? t1 : fadd<S>(a1, t1);
            }
            put(12 + m, a0, a1);
        }
    }
};

// K4 Reddy: examples/stokes/colliding_flow/ht_p2_p1.jl:77-101 (Jacobian of the PRESSURE element);
// veclap: examples/stokes/colliding_flow/ht_p2_p1_veclap.jl:75-94 (no ux-uy coupling blocks).
// The same loop on the other velocity / pressure pairs of the examples, PAIR =
//   EFG_PAIR_T6_T3 (6)   FEH1_T6 / FEH1_T3 on the T6toT3 mesh (slot 1), Jacobian of the pressure element
//   EFG_PAIR_T3B_T3 (7)  FEH1_T3_BUBBLE / FEH1_T3, one T3 mesh: examples/stokes/colliding_flow/p1b_p1.jl:53-109,
//                        test/test_stokes.jl:190-247; the 4th velocity dof sits on the cell (src/FElements.jl:339)
//   EFG_PAIR_Q4_L2 (14)  FEH1_Q4 / FEL2_Q4, one Q4 mesh: examples/stokes/colliding_flow/q1_q0.jl:52-108 -- Jacobian of
//                        the velocity element (:70), the single pressure dof sits on the cell (src/FElements.jl:410)
//   EFG_PAIR_T3_L2 (13)  FEH1_T3 / FEL2_T3 (src/FElements.jl:422-448), same structure on triangles
// Local dofs: [ux: NV][uy: NV][p: NP]; eldofs() of a space = its vertex dofs, then its cell dof (src/FEIterators.jl:185-194).
#define EFG_PAIR_T6_T3 6
#define EFG_PAIR_T3B_T3 7
#define EFG_PAIR_T3_L2 13
#define EFG_PAIR_Q4_L2 14
template <int PAIR> struct ReddyPair;
template <> struct ReddyPair<EFG_PAIR_T6_T3>  { static constexpr int NV = 6, NP = 3, VN = 6, PN = 3, VC = 0, PC = 0, GK = 3, GMESH = 1, PMESH = 1, VS = 2, PS = 0; };
template <> struct ReddyPair<EFG_PAIR_T3B_T3> { static constexpr int NV = 4, NP = 3, VN = 3, PN = 3, VC = 1, PC = 0, GK = 3, GMESH = 0, PMESH = 0, VS = 3, PS = 0; };
template <> struct ReddyPair<EFG_PAIR_T3_L2>  { static constexpr int NV = 3, NP = 1, VN = 3, PN = 0, VC = 0, PC = 1, GK = 3, GMESH = 0, PMESH = 0, VS = 0, PS = 4; };
template <> struct ReddyPair<EFG_PAIR_Q4_L2>  { static constexpr int NV = 4, NP = 1, VN = 4, PN = 0, VC = 0, PC = 1, GK = 4, GMESH = 0, PMESH = 0, VS = 1, PS = 4; };

template <bool VECLAP, int PAIR = EFG_PAIR_T6_T3, int NQ_ = 3> struct Stokes3Form {
    using P = ReddyPair<PAIR>;
    static constexpr bool SYM = false;
    static constexpr int NV = P::NV, NP = P::NP, NQ = NQ_;
    // BK = number of velocity basis functions (sizes the gradient arrays), BSLOT = their table in c_tab
    static constexpr int ND = 2 * NV + NP, NT = (VECLAP ? 2 : 4) * NV * NV + 4 * NV * NP, GK = P::GK, BK = NV, BSLOT = P::VS, GMESH = P::GMESH, NSPACES = 3;
    static constexpr bool SPLIT = false;
    template <bool S, class Emit>
    __device__ __forceinline__ static void element(const double (&X)[GK], const double (&Y)[GK], uint32_t m, Emit &emit) {
        element_generic<Stokes3Form<VECLAP, PAIR, NQ_>, S, Emit>(X, Y, m, emit);
    }
    __host__ __device__ static constexpr bool mask(int i, int j) {
        if (i >= 2 * NV && j >= 2 * NV) return false;
        if (VECLAP && ((i < NV && j >= NV && j < 2 * NV) || (j < NV && i >= NV && i < 2 * NV))) return false;
        return true;
    }
    __host__ __device__ static constexpr int kidx(int i, int j) {
        constexpr int VV = NV * NV, VP = NV * NP, U2 = 2 * NV;
        if (!VECLAP) {
            if (i < NV && j < NV) return j * NV + i;                                   // kuxux
            if (i < NV && j < U2) return VV + (j - NV) * NV + i;                       // kuxuy
            if (i < U2 && j < NV) return 2 * VV + j * NV + (i - NV);                   // transpose(kuxuy)
            if (i < U2 && j < U2) return 3 * VV + (j - NV) * NV + (i - NV);            // kuyuy
            if (i < NV) return 4 * VV + (j - U2) * NV + i;                             // kuxp
            if (j < NV) return 4 * VV + VP + j * NP + (i - U2);                        // transpose(kuxp)
            if (i < U2) return 4 * VV + 2 * VP + (j - U2) * NV + (i - NV);             // kuyp
            return 4 * VV + 3 * VP + (j - NV) * NP + (i - U2);                         // transpose(kuyp)
        } else {
            if (i < NV && j < NV) return j * NV + i;                                   // kuxux
            if (i >= NV && i < U2 && j >= NV && j < U2) return VV + (j - NV) * NV + (i - NV); // kuyuy
            if (i < NV) return 2 * VV + (j - U2) * NV + i;                             // kuxp
            if (j < NV) return 2 * VV + VP + j * NP + (i - U2);                        // transpose(kuxp)
            if (i < U2) return 2 * VV + 2 * VP + (j - U2) * NV + (i - NV);             // kuyp
            return 2 * VV + 3 * VP + (j - NV) * NP + (i - U2);                         // transpose(kuyp)
        }
    }
    __device__ static void edofs(const DofSrc &s, int64_t e, int32_t (&d)[ND]) {
#pragma unroll
        for (int a = 0; a < P::VN; a++) {
            const int64_t n = s.conn0[e * P::VN + a];
            d[a] = s.dof0[n]; d[NV + a] = s.dof1[n];
        }
        if constexpr (P::VC) { d[P::VN] = s.cdof0[e]; d[NV + P::VN] = s.cdof1[e]; }
        const int32_t *pconn = P::PMESH ? s.conn1 : s.conn0;
#pragma unroll
        for (int m = 0; m < P::PN; m++) d[2 * NV + m] = s.dof2[pconn[e * P::PN + m]];
        if constexpr (P::PC) d[2 * NV + P::PN] = s.cdof2[e];
    }
    template <bool S, int J> __device__ __forceinline__ static void column(const Geo<NV, NQ> &G, double (&out)[ND]) {
        const QTab &tp = c_tab[P::PS];
        const double mu = c_prm[0];
#pragma unroll
        for (int i = 0; i < ND; i++) out[i] = 0.0;
        if constexpr (J < NV) {          // column of ux dof J
#pragma unroll
            for (int i = 0; i < NV; i++) {
                double a = 0.0, b = 0.0;
#pragma unroll
                for (int q = 0; q < NQ; q++) {
                    const double mJ = fmul<S>(mu, G.JxW[q]);
                    double t;
                    if constexpr (!VECLAP)   // (mu*JxW) * (2*gx_i*gx_j + gy_i*gy_j)
                        t = fmul<S>(mJ, fadd<S>(fmul<S>(fmul<S>(2.0, G.gx[q][i]), G.gx[q][J]), fmul<S>(G.gy[q][i], G.gy[q][J])));
                    else
                        t = fmul<S>(mJ, fadd<S>(fmul<S>(G.gx[q][i], G.gx[q][J]), fmul<S>(G.gy[q][i], G.gy[q][J])));
                    a = q == 0 ? t : fadd<S>(a, t);
                    if constexpr (!VECLAP) { // transpose(kuxuy): value kuxuy[J, i] = (mu*JxW) * (gx_J * gy_i)
                        const double u = fmul<S>(mJ, fmul<S>(G.gx[q][J], G.gy[q][i]));
                        b = q == 0 ? u : fadd<S>(b, u);
                    }
                }
                out[i] = a; out[NV + i] = b;
            }
#pragma unroll
            for (int m = 0; m < NP; m++) {  // transpose(kuxp): kuxp[J, m] = (-JxW) * (gx_J * Np_m)
                double acc = 0.0;
#pragma unroll
                for (int q = 0; q < NQ; q++) {
                    const double t = fmul<S>(-G.JxW[q], fmul<S>(G.gx[q][J], tp.N[q][m]));
                    acc = q == 0 ? t : fadd<S>(acc, t);
                }
                out[2 * NV + m] = acc;
            }
        } else if constexpr (J < 2 * NV) {  // column of uy dof b
            constexpr int b_ = J - NV;
#pragma unroll
            for (int i = 0; i < NV; i++) {
                double a = 0.0, c = 0.0;
#pragma unroll
                for (int q = 0; q < NQ; q++) {
                    const double mJ = fmul<S>(mu, G.JxW[q]);
                    if constexpr (!VECLAP) { // kuxuy[i, b] = (mu*JxW) * (gx_i * gy_b)
                        const double u = fmul<S>(mJ, fmul<S>(G.gx[q][i], G.gy[q][b_]));
                        a = q == 0 ? u : fadd<S>(a, u);
                    }
                    double t;
                    if constexpr (!VECLAP)   // (mu*JxW) * (gx_i*gx_b + 2*gy_i*gy_b)
                        t = fmul<S>(mJ, fadd<S>(fmul<S>(G.gx[q][i], G.gx[q][b_]), fmul<S>(fmul<S>(2.0, G.gy[q][i]), G.gy[q][b_])));
                    else
                        t = fmul<S>(mJ, fadd<S>(fmul<S>(G.gx[q][i], G.gx[q][b_]), fmul<S>(G.gy[q][i], G.gy[q][b_])));
                    c = q == 0 ? t : fadd<S>(c, t);
                }
                out[i] = a; out[NV + i] = c;
            }
#pragma unroll
            for (int m = 0; m < NP; m++) {  // transpose(kuyp): kuyp[b, m] = (-JxW) * (gy_b * Np_m)
                double acc = 0.0;
#pragma unroll
                for (int q = 0; q < NQ; q++) {
                    const double t = fmul<S>(-G.JxW[q], fmul<S>(G.gy[q][b_], tp.N[q][m]));
                    acc = q == 0 ? t : fadd<S>(acc, t);
                }
                out[2 * NV + m] = acc;
            }
        } else {                        // column of p dof m: kuxp[i, m], kuyp[i, m]
            constexpr int m = J - 2 * NV;
#pragma unroll
            for (int i = 0; i < NV; i++) {
                double a = 0.0, c = 0.0;
#pragma unroll
                for (int q = 0; q < NQ; q++) {
                    const double t = fmul<S>(-G.JxW[q], fmul<S>(G.gx[q][i], tp.N[q][m]));
                    const double u = fmul<S>(-G.JxW[q], fmul<S>(G.gy[q][i], tp.N[q][m]));
                    a = q == 0 ? t : fadd<S>(a, t);
                    c = q == 0 ? u : fadd<S>(c, u);
                }
                out[i] = a; out[NV + i] = c;
            }
        }
    }
};

// ------------------------------------------------------------------------------------------------
// K1 heat on FEH1_T4: examples/heat/poisson/t4.jl:31-57 -- the loop of t3.jl in 3-D.
//   _jac: J = sum_n x_n (outer) g_n, 3x3, node order, first term assigned (src/FElements.jl:148-156), never contracted
//   Jacobian(Val{3}): src/FElements.jl:138-146 (literal grouping)
//   bfungrad: g / Jac = (Jac' \ g)' -- StaticArrays' closed-form 3x3 solve: det by dot(col1, cross(col2, col3)), then the
//   cofactor rows times g, three divisions by the determinant per basis function
// One thread per element; assembled by the two-pass path (DIM3: the tiled kernel's geometry blocks are 2-D).
// ------------------------------------------------------------------------------------------------
template <class F, class = void> struct form_dim3 { static constexpr bool value = false; };
template <class F> struct form_dim3<F, std::void_t<decltype(F::DIM3)>> { static constexpr bool value = F::DIM3; };

template <int NQ_> struct HeatFormT4 {
    static constexpr bool DIM3 = true, SYM = true, SPLIT = false;
    static constexpr int ND = 4, NT = 16, GK = 4, BK = 4, NQ = NQ_, GMESH = 0, NSPACES = 1;
    __host__ __device__ static constexpr bool mask(int, int) { return true; }
    __host__ __device__ static constexpr int kidx(int i, int j) { return j * ND + i; }
    __device__ static void edofs(const DofSrc &s, int64_t e, int32_t (&d)[ND]) {
#pragma unroll
        for (int a = 0; a < 4; a++) d[a] = s.dof0[s.conn0[e * 4 + a]];
    }
    template <bool S>
    __device__ __forceinline__ static void geometry(const double (&X)[4], const double (&Y)[4], const double (&Z)[4],
                                                    double (&g)[4][3], double &det) {
        const double GP[4][3] = {{-1.0, -1.0, -1.0}, {+1.0, 0.0, 0.0}, {0.0, +1.0, 0.0}, {0.0, 0.0, +1.0}};
        double J[3][3];
#pragma unroll
        for (int k = 0; k < 3; k++) { J[0][k] = __dmul_rn(X[0], GP[0][k]); J[1][k] = __dmul_rn(Y[0], GP[0][k]); J[2][k] = __dmul_rn(Z[0], GP[0][k]); }
#pragma unroll
        for (int n = 1; n < 4; n++)
#pragma unroll
            for (int k = 0; k < 3; k++) {
                J[0][k] = __dadd_rn(J[0][k], __dmul_rn(X[n], GP[n][k]));
                J[1][k] = __dadd_rn(J[1][k], __dmul_rn(Y[n], GP[n][k]));
                J[2][k] = __dadd_rn(J[2][k], __dmul_rn(Z[n], GP[n][k]));
            }
        // the determinant and the cofactors cancel like the 2-D Jacobian: individually rounded in every mode
        auto m2 = [](double a, double b, double c, double d) { return __dsub_rn(__dmul_rn(a, b), __dmul_rn(c, d)); };
        det = __dadd_rn(__dsub_rn(__dmul_rn(J[0][0], m2(J[1][1], J[2][2], J[2][1], J[1][2])),
                                  __dmul_rn(J[0][1], m2(J[1][0], J[2][2], J[1][2], J[2][0]))),
                        __dmul_rn(J[0][2], m2(J[1][0], J[2][1], J[1][1], J[2][0])));
        double a[3][3];
#pragma unroll
        for (int r = 0; r < 3; r++)
#pragma unroll
            for (int c = 0; c < 3; c++) a[r][c] = J[c][r];
        const double c0 = m2(a[1][1], a[2][2], a[2][1], a[1][2]), c1 = m2(a[2][1], a[0][2], a[0][1], a[2][2]), c2 = m2(a[0][1], a[1][2], a[1][1], a[0][2]);
        const double d = __dadd_rn(__dadd_rn(__dmul_rn(a[0][0], c0), __dmul_rn(a[1][0], c1)), __dmul_rn(a[2][0], c2));
        const double M[3][3] = {{m2(a[1][1], a[2][2], a[1][2], a[2][1]), m2(a[0][2], a[2][1], a[0][1], a[2][2]), m2(a[0][1], a[1][2], a[0][2], a[1][1])},
                                {m2(a[1][2], a[2][0], a[1][0], a[2][2]), m2(a[0][0], a[2][2], a[0][2], a[2][0]), m2(a[0][2], a[1][0], a[0][0], a[1][2])},
                                {m2(a[1][0], a[2][1], a[1][1], a[2][0]), m2(a[0][1], a[2][0], a[0][0], a[2][1]), m2(a[0][0], a[1][1], a[0][1], a[1][0])}};
        const SharedDivisor<S> div(d);
#pragma unroll
        for (int j = 0; j < 4; j++)
#pragma unroll
            for (int r = 0; r < 3; r++)
                g[j][r] = div(__dadd_rn(__dadd_rn(__dmul_rn(M[r][0], GP[j][0]), __dmul_rn(M[r][1], GP[j][1])), __dmul_rn(M[r][2], GP[j][2])));
    }
    template <class Emit, int J> __device__ __forceinline__ static void emit_col(const double (&K)[ND][ND], uint32_t m, Emit &emit) {
        if (m & (1u << J)) {
            double out[ND];
#pragma unroll
            for (int i = 0; i < ND; i++) out[i] = (i <= J) ? K[i][J] : K[J][i];
            emit.template col<J>(out);
        }
    }
    template <bool S, class Emit>
    __device__ __forceinline__ static void element(const double (&)[4], const double (&)[4], uint32_t, Emit &) {}      // (2-D entry point: unused)
    template <bool S, class Emit>
    __device__ __forceinline__ static void element3(const double (&X)[4], const double (&Y)[4], const double (&Z)[4], uint32_t m, Emit &emit) {
        const QTab &t = c_tab[EFG_TAB_T4];
        const double kappa = c_prm[0];
        double g[4][3], det;
        geometry<S>(X, Y, Z, g, det);          // affine element: the same Jacobian at every quadrature point
        double K[ND][ND];
#pragma unroll
        for (int q = 0; q < NQ; q++) {
            const double kJ = fmul<S>(kappa, fmul<S>(det, t.w[q]));
#pragma unroll
            for (int j = 0; j < ND; j++)
#pragma unroll
                for (int i = 0; i <= j; i++) {
                    const double dt = fadd<S>(fadd<S>(fmul<S>(g[i][0], g[j][0]), fmul<S>(g[i][1], g[j][1])), fmul<S>(g[i][2], g[j][2]));
                    const double v = fmul<S>(dt, kJ);
                    K[i][j] = q == 0 ? v : fadd<S>(K[i][j], v);
                }
        }
        if constexpr (Emit::TRI) emit.tri(K, m);
        else { emit_col<Emit, 0>(K, m, emit); emit_col<Emit, 1>(K, m, emit); emit_col<Emit, 2>(K, m, emit); emit_col<Emit, 3>(K, m, emit); }
    }
};
