/*
 * elfel_oracle.c -- CPU ORACLE (TEST INFRASTRUCTURE, NOT PRODUCT CODE).
 *
 * A plain-C, single-threaded restatement of the algorithm of Elfel.jl's assembly
 * hot path: per-element quadrature loop -> LocalMatrixAssembler -> COO append in
 * SysmatAssemblerSparse -> SparseArrays.sparse().  Only tests/, bench.py's
 * cpu_baseline / --impl reference leg and __graft_entry__.smoke() may load it, and
 * only as the checker / CPU baseline.  The product (libelfelgpu.so) never links it.
 *
 * Build: gcc -O2 -ffp-contract=off (no FMA contraction, so this file is the fixed
 * point of the operation order written in the reference).
 *
 * Parity pin: validated against the reference's own golden values by
 * tests/test_oracle_golden.py (test/test_assemblers.jl:27-34, test/test_heat.jl:110,
 * test/test_stokes.jl:130,551-553,779, test/test_qpiterators.jl:21-61,
 * test/test_refshapes.jl:26-29,45-48, test/test_feiterators.jl:58-64,154-172).
 * Julia itself cannot run here (no julia binary, MeshCore/MeshSteward/StaticArrays are
 * not vendored), so ULP-level operation order inside StaticArrays' 2x2 solve and
 * SparseArrays.sparse is restated from their published algorithms (StaticArrays 1.0.1
 * src/solve.jl, SparseArrays sparse!) and is "parity unpinned" below 1e-12.
 *
 * Every function cites the reference file:line it follows (paths relative to the
 * Elfel.jl repository root).
 *
 * Array layouts are the reference's memory layouts, 1-based Int64 indices:
 *   conn    : nen x nel   Int64 (node ids of element e at conn[e*nen + k])  -- MeshCore IncRel
 *   xy      : 2 x nnodes  Float64                                          -- VecAttrib "geom"
 *   dofnums : ncomp x nnodes Int64 (FEField.dofnums, src/FEFields.jl:15)
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <math.h>

#define EFO_T3 3
#define EFO_Q4 4
#define EFO_T6 6
/* finite elements with a dof on the cell itself (FEData ndofperfeat = [1,0,1,0] / [0,0,1,0]) */
#define EFO_T3B 7   /* FEH1_T3_BUBBLE: src/FElements.jl:324-355 -- 3 vertex functions + the cubic bubble of the cell      */
#define EFO_L2 1    /* FEL2_T3 / FEL2_Q4: src/FElements.jl:394-445 -- one constant function, dof on the cell; the geometry */
                    /* carrier is the H1 element of the mesh (_geometrycarrier, :403,:432)                                */

/* Spaces whose element is not the plain H1 element of their mesh (row f5).  NULL = the classic case.
 *   vfe : element of the velocity space(s) (spaces 0, 1 of the Reddy forms): 0 = H1 element of mesh 0, or EFO_T3B
 *   pfe : element of the pressure space (space 2):                           0 = H1 element of mesh 1, or EFO_L2
 *   cdof0..2 : FEField.dofnums of the dim-2 (cell) field of spaces 0..2, ncomp x nel, NULL = the space has none */
typedef struct {
    int vfe, pfe;
    const int64_t *cdof0, *cdof1, *cdof2;
} efo_ext;

enum {
    EFO_FORM_HEAT = 1,            /* examples/heat/poisson/t3.jl:53-58, q4.jl:43-48 */
    EFO_FORM_ELASTICITY = 2,      /* examples/elasticity/stretch/t6.jl:42-58 */
    EFO_FORM_STOKES_GEN = 3,      /* examples/stokes/colliding_flow/ht_p2_p1_gen.jl:48-79 */
    EFO_FORM_STOKES_REDDY = 4,    /* examples/stokes/colliding_flow/ht_p2_p1.jl:56-104 */
    EFO_FORM_STOKES_VECLAP_ALT = 5, /* examples/stokes/colliding_flow/ht_p2_p1_veclap_alt.jl:61-90 */
    EFO_FORM_STOKES_VECLAP = 6    /* examples/stokes/colliding_flow/ht_p2_p1_veclap.jl:55-97 */
};

/* ------------------------------------------------------------------------------------------
 * Quadrature rules.  src/RefShapes.jl:85-110 (_gauss1), :113-119 (_triangle),
 * :301-323 (triangle rule), :333-366 (square tensor product, i outer / j inner).
 * pc is npts x 2 row-major here (point p: pc[2p], pc[2p+1]).  Returns npts or -1.
 * For triangles `rule` is npts (1 or 3); for squares `rule` is the Gauss order (1..5).
 * ------------------------------------------------------------------------------------------ */
static int gauss1(int order, double *pc, double *w) /* src/RefShapes.jl:91-106 */
{
    switch (order) {
    case 1: pc[0] = 0.0; w[0] = 2.0; return 1;
    case 2: pc[0] = -0.577350269189626; pc[1] = 0.577350269189626; w[0] = 1.0; w[1] = 1.0; return 2;
    case 3: pc[0] = -0.774596669241483; pc[1] = 0.0; pc[2] = 0.774596669241483;
            w[0] = 0.5555555555555556; w[1] = 0.8888888888888889; w[2] = 0.5555555555555556; return 3;
    case 4: pc[0] = -0.86113631159405; pc[1] = -0.33998104358486; pc[2] = 0.33998104358486; pc[3] = 0.86113631159405;
            w[0] = 0.34785484513745; w[1] = 0.65214515486255; w[2] = 0.65214515486255; w[3] = 0.34785484513745; return 4;
    case 5: pc[0] = -0.906179845938664; pc[1] = -0.538469310105683; pc[2] = 0.000000000000000;
            pc[3] = 0.538469310105683; pc[4] = 0.906179845938664;
            w[0] = 0.236926885056189; w[1] = 0.478628670499367; w[2] = 0.568888888888889;
            w[3] = 0.478628670499367; w[4] = 0.236926885056189; return 5;
    default: return -1; /* order > 5 is Golub-Welsch in the reference: out of scope */
    }
}

int efo_quadrature(int elemkind, int rule, double *pc, double *w)
{
    if (elemkind == EFO_T3 || elemkind == EFO_T6 || elemkind == EFO_T3B) {
        if (rule == 1) { /* src/RefShapes.jl:114-116 */
            pc[0] = 1.0 / 3.; pc[1] = 1.0 / 3.;
            w[0] = 1.0 / 2.0;
            return 1;
        } else if (rule == 3) { /* src/RefShapes.jl:117-119 */
            pc[0] = 2.0 / 3; pc[1] = 1.0 / 6;
            pc[2] = 1.0 / 6; pc[3] = 2.0 / 3;
            pc[4] = 1.0 / 6; pc[5] = 1.0 / 6;
            w[0] = (1.0 / 3) / 2; w[1] = (1.0 / 3) / 2; w[2] = (1.0 / 3) / 2;
            return 3;
        }
        /* the higher rules of _triangle, src/RefShapes.jl:120-230: literal tables (weights divided by 2 where the source does) */
        {
            static const double P4[4][2] = {{0.333333333333333, 0.333333333333333}, {0.200000000000000, 0.200000000000000},
                                            {0.600000000000000, 0.200000000000000}, {0.200000000000000, 0.600000000000000}};
            static const double W4[4] = {-0.281250000000000, 0.260416666666667, 0.260416666666667, 0.260416666666667};
            static const double P6[6][2] = {{0.816847572980459, 0.091576213509771}, {0.091576213509771, 0.816847572980459},
                                            {0.091576213509771, 0.091576213509771}, {0.108103018168070, 0.445948490915965},
                                            {0.445948490915965, 0.108103018168070}, {0.445948490915965, 0.445948490915965}};
            static const double W6[6] = {0.109951743655322, 0.109951743655322, 0.109951743655322, 0.223381589678011, 0.223381589678011,
                                         0.223381589678011};                                  /* ... / 2 */
            static const double P7[7][2] = {{0.101286507323456, 0.101286507323456}, {0.797426958353087, 0.101286507323456},
                                            {0.101286507323456, 0.797426958353087}, {0.470142064105115, 0.470142064105115},
                                            {0.059715871789770, 0.470142064105115}, {0.470142064105115, 0.059715871789770},
                                            {0.333333333333333, 0.333333333333333}};
            static const double W7[7] = {0.062969590272414, 0.062969590272414, 0.062969590272414, 0.066197076394253, 0.066197076394253,
                                         0.066197076394253, 0.112500000000000};
            static const double P9[9][2] = {{0.437525248383384, 0.437525248383384}, {0.124949503233232, 0.437525248383384},
                                            {0.437525248383384, 0.124949503233232}, {0.165409927389841, 0.037477420750088},
                                            {0.037477420750088, 0.165409927389841}, {0.797112651860071, 0.165409927389841},
                                            {0.165409927389841, 0.797112651860071}, {0.037477420750088, 0.797112651860071},
                                            {0.797112651860071, 0.037477420750088}};
            static const double W9[9] = {0.205950504760887, 0.205950504760887, 0.205950504760887, 0.063691414286223, 0.063691414286223,
                                         0.063691414286223, 0.063691414286223, 0.063691414286223, 0.063691414286223};   /* ... ./ 2 */
            static const double P12[12][2] = {{0.063089014491502, 0.063089014491502}, {0.873821971016996, 0.063089014491502},
                                              {0.063089014491502, 0.873821971016996}, {0.249286745170910, 0.249286745170910},
                                              {0.501426509658179, 0.249286745170910}, {0.249286745170910, 0.501426509658179},
                                              {0.310352451033785, 0.053145049844816}, {0.053145049844816, 0.310352451033785},
                                              {0.636502499121399, 0.310352451033785}, {0.310352451033785, 0.636502499121399},
                                              {0.053145049844816, 0.636502499121399}, {0.636502499121399, 0.053145049844816}};
            static const double W12[12] = {0.050844906370207, 0.050844906370207, 0.050844906370207, 0.116786275726379, 0.116786275726379,
                                           0.116786275726379, 0.082851075618374, 0.082851075618374, 0.082851075618374, 0.082851075618374,
                                           0.082851075618374, 0.082851075618374};               /* ... ./ 2 */
            static const double P13[13][2] = {{0.333333333333333, 0.333333333333333}, {0.479308067841923, 0.260345966079038},
                                              {0.260345966079038, 0.479308067841923}, {0.260345966079038, 0.260345966079038},
                                              {0.869739794195568, 0.065130102902216}, {0.065130102902216, 0.869739794195568},
                                              {0.065130102902216, 0.065130102902216}, {0.638444188569809, 0.312865496004875},
                                              {0.638444188569809, 0.048690315425316}, {0.312865496004875, 0.638444188569809},
                                              {0.312865496004875, 0.048690315425316}, {0.048690315425316, 0.638444188569809},
                                              {0.048690315425316, 0.312865496004875}};
            static const double W13[13] = {-0.149570044467670, 0.175615257433204, 0.175615257433204, 0.175615257433204, 0.053347235608839,
                                           0.053347235608839, 0.053347235608839, 0.077113760890257, 0.077113760890257, 0.077113760890257,
                                           0.077113760890257, 0.077113760890257, 0.077113760890257};   /* ... ' / 2 */
            const double (*P)[2] = NULL; const double *W = NULL; int halve = 0;
            switch (rule) {
            case 4: P = P4; W = W4; break;
            case 6: P = P6; W = W6; halve = 1; break;
            case 7: P = P7; W = W7; break;
            case 9: P = P9; W = W9; halve = 1; break;
            case 12: P = P12; W = W12; halve = 1; break;
            case 13: P = P13; W = W13; halve = 1; break;
            default: return -1;
            }
            for (int q = 0; q < rule; q++) { pc[2 * q] = P[q][0]; pc[2 * q + 1] = P[q][1]; w[q] = halve ? W[q] / 2 : W[q]; }
            return rule;
        }
    } else if (elemkind == EFO_Q4) { /* src/RefShapes.jl:350-362 */
        double p1[5], w1[5];
        int np = gauss1(rule, p1, w1);
        if (np < 0) return -1;
        int r = 0;
        for (int i = 0; i < np; i++)
            for (int j = 0; j < np; j++) {
                pc[2 * r] = p1[i]; pc[2 * r + 1] = p1[j];
                w[r] = w1[i] * w1[j];
                r++;
            }
        return r;
    }
    return -1;
}

/* ------------------------------------------------------------------------------------------
 * Basis functions and parametric gradients.
 * T3: src/FElements.jl:239-246; T6: :264-288; Q4: :306-320.  g is nbf x 2 row-major.
 * Julia evaluates `-3+4*r+4*s` as (-3 + 4r) + 4s, `4-8*r-4*s` as (4 - 8r) - 4s.
 * ------------------------------------------------------------------------------------------ */
int efo_nbf(int elemkind) { return (elemkind == EFO_T3B || elemkind == 40 /* EFO_T4 */) ? 4 : elemkind; }
/* dofs of an element: on its nodes (dim 0 field) / on the cell (dim 2 field); _storedofs! visits dim 0 first
 * (src/FEIterators.jl:185-194), _number_edofs numbers the basis functions in the same order (src/FESpaces.jl:87-105) */
static int fe_nodedofs(int fe) { return fe == EFO_T3B ? 3 : (fe == EFO_L2 ? 0 : fe); }
static int fe_celldofs(int fe) { return (fe == EFO_T3B || fe == EFO_L2) ? 1 : 0; }

void efo_bfun(int elemkind, double r, double s, double *N)
{
    if (elemkind == EFO_T3) {
        N[0] = (1 - r - s); N[1] = r; N[2] = s;
    } else if (elemkind == EFO_T6) {
        double t = 1. - r - s;
        N[0] = t * (t + t - 1);
        N[1] = r * (r + r - 1);
        N[2] = s * (s + s - 1);
        N[3] = 4 * r * t;
        N[4] = 4 * r * s;
        N[5] = 4 * s * t;
    } else if (elemkind == EFO_Q4) {
        N[0] = 0.25 * (1. - r) * (1. - s);
        N[1] = 0.25 * (1. + r) * (1. - s);
        N[2] = 0.25 * (1. + r) * (1. + s);
        N[3] = 0.25 * (1. - r) * (1. + s);
    } else if (elemkind == EFO_T3B) {   /* src/FElements.jl:341-347: ((1 - xi - eta) * xi) * eta */
        N[0] = (1 - r - s); N[1] = r; N[2] = s;
        N[3] = (1 - r - s) * r * s;
    } else if (elemkind == EFO_L2) {    /* src/FElements.jl:412-414, 441-443 */
        N[0] = 1.0;
    }
}

void efo_bfungradpar(int elemkind, double r, double s, double *g)
{
    if (elemkind == EFO_T3) {
        g[0] = -1.; g[1] = -1.;
        g[2] = +1.; g[3] = 0.;
        g[4] = 0.;  g[5] = +1.;
    } else if (elemkind == EFO_T6) {
        g[0] = -3 + 4 * r + 4 * s; g[1] = -3 + 4 * r + 4 * s;
        g[2] = 4 * r - 1;          g[3] = 0.0;
        g[4] = 0.0;                g[5] = 4 * s - 1;
        g[6] = 4 - 8 * r - 4 * s;  g[7] = -4 * r;
        g[8] = 4 * s;              g[9] = 4 * r;
        g[10] = -4 * s;            g[11] = 4 - 4 * r - 8 * s;
    } else if (elemkind == EFO_Q4) {
        g[0] = -(1. - s) * 0.25; g[1] = -(1. - r) * 0.25;
        g[2] = (1. - s) * 0.25;  g[3] = -(1. + r) * 0.25;
        g[4] = (1. + s) * 0.25;  g[5] = (1. + r) * 0.25;
        g[6] = -(1. + s) * 0.25; g[7] = (1. - r) * 0.25;
    } else if (elemkind == EFO_T3B) {   /* src/FElements.jl:349-356 */
        g[0] = -1.; g[1] = -1.;
        g[2] = +1.; g[3] = 0.;
        g[4] = 0.;  g[5] = +1.;
        g[6] = (-r * s + (1 - r - s) * s); g[7] = (-r * s + (1 - r - s) * r);
    } else if (elemkind == EFO_L2) {    /* src/FElements.jl:416-419, 445-448 */
        g[0] = 0.0; g[1] = 0.0;
    }
}

/* ------------------------------------------------------------------------------------------
 * Per-quadrature-point tables: QPIterator ctor / __bfundata, src/QPIterators.jl:14-50,79-84.
 * ------------------------------------------------------------------------------------------ */
#define MAXQP 25
#define MAXBF 6
typedef struct {
    int kind, nbf, npts;
    double w[MAXQP];
    double N[MAXQP][MAXBF];        /* scalar basis functions       */
    double gp[MAXQP][MAXBF][2];    /* scalar parametric gradients  */
} qptab;

/* shape: the element kind whose reference shape supplies the rule (an L2 element lives on its mesh's T3 / Q4 cell) */
static int qptab_init_shape(qptab *t, int elemkind, int shape, int rule)
{
    double pc[2 * MAXQP];
    t->kind = elemkind; t->nbf = efo_nbf(elemkind);
    t->npts = efo_quadrature(shape, rule, pc, t->w);
    if (t->npts < 0) return -1;
    for (int q = 0; q < t->npts; q++) {
        efo_bfun(elemkind, pc[2 * q], pc[2 * q + 1], t->N[q]);
        efo_bfungradpar(elemkind, pc[2 * q], pc[2 * q + 1], &t->gp[q][0][0]);
    }
    return 0;
}
static int qptab_init(qptab *t, int elemkind, int rule) { return qptab_init_shape(t, elemkind, elemkind, rule); }

/* _jac: src/FElements.jl:148-156.  J = sum_n x_n (outer) gradNpar_n, summed in node order,
 * first term assigned.  J[i][k] = sum_n x_n[i] * g_n[k].
 * Jacobian(Val{2}): src/FElements.jl:120-129: J11*J22 - J21*J12. */
static double jacjac(const double *xy, const int64_t *nodes, int nen, const double gp[][2], double J[2][2])
{
    const double *x = xy + 2 * (nodes[0] - 1);
    J[0][0] = x[0] * gp[0][0]; J[0][1] = x[0] * gp[0][1];
    J[1][0] = x[1] * gp[0][0]; J[1][1] = x[1] * gp[0][1];
    for (int n = 1; n < nen; n++) {
        x = xy + 2 * (nodes[n] - 1);
        J[0][0] = J[0][0] + x[0] * gp[n][0]; J[0][1] = J[0][1] + x[0] * gp[n][1];
        J[1][0] = J[1][0] + x[1] * gp[n][0]; J[1][1] = J[1][1] + x[1] * gp[n][1];
    }
    return J[0][0] * J[1][1] - J[1][0] * J[0][1];
}

/* bfungrad: src/QPIterators.jl:132-140.  gradpar[j] / Jac with Adjoint{SVector{2}} / SMatrix{2,2}
 * = (Jac' \ g)' ; StaticArrays 1.0.1 2x2 solve: d = det(Jac'), x = ((a22 b1 - a12 b2)/d, (a11 b2 - a21 b1)/d)
 * with a = Jac'.  Two true divisions by d. */
static void bfungrad(int nbf, const double gp[][2], double J[2][2], double g[][2])
{
    double d = J[0][0] * J[1][1] - J[1][0] * J[0][1];
    for (int j = 0; j < nbf; j++) {
        g[j][0] = (J[1][1] * gp[j][0] - J[1][0] * gp[j][1]) / d;
        g[j][1] = (J[0][0] * gp[j][1] - J[0][1] * gp[j][0]) / d;
    }
}

/* ------------------------------------------------------------------------------------------
 * FEH1_T4 -- the 3-D member of row f5: examples/heat/poisson/t4.jl (the same integrate! loop as t3.jl on tetrahedra).
 *   basis / parametric gradients   src/FElements.jl:359-386 (N = (1-r-s-t, r, s, t), constant gradient rows)
 *   rules                          src/RefShapes.jl:232-259 (_tetrahedron npts 1, 4, 5), weights as literally written
 *   _jac                           src/FElements.jl:148-156: J = sum_n x_n (outer) g_n, 3x3, node order, first term assigned
 *   Jacobian(Val{3})               src/FElements.jl:138-146 (the unrolled determinant, literal grouping)
 *   bfungrad                       src/QPIterators.jl:132-140: g / Jac = (Jac' \ g)' with StaticArrays 1.0.1's closed-form
 *                                  3x3 solve (src/solve.jl): d = det(a) = dot(col1, cross(col2, col3)), then the cofactor
 *                                  rows times b, each divided by d (three true divisions).  Not vendored: restated from the
 *                                  published source, "parity unpinned" below 1e-12 like the 2x2 solve above.
 * xyz is 3 x nnodes (VecAttrib of SVector{3}).
 * ------------------------------------------------------------------------------------------ */
#define EFO_T4 40
static const double T4_GP[4][3] = {{-1.0, -1.0, -1.0}, {+1.0, 0.0, 0.0}, {0.0, +1.0, 0.0}, {0.0, 0.0, +1.0}};

int efo_quadrature_t4(int npts, double *pc /* npts x 3 */, double *w)
{
    if (npts == 1) {
        pc[0] = 0.25; pc[1] = 0.25; pc[2] = 0.25;
        w[0] = 1.0 / 6.0;
        return 1;
    } else if (npts == 4) {
        const double a = 0.13819660, b = 0.58541020;
        const double P[4][3] = {{a, a, a}, {b, a, a}, {a, b, a}, {a, a, b}};
        for (int q = 0; q < 4; q++) { pc[3 * q] = P[q][0]; pc[3 * q + 1] = P[q][1]; pc[3 * q + 2] = P[q][2]; w[q] = 0.041666666666666666667; }
        return 4;
    } else if (npts == 5) { /* Zienkiewicz #3 */
        const double a = 1.0 / 6.0, b = 0.25, c = 0.5, d = -0.8, e = 0.45;
        const double P[5][3] = {{b, b, b}, {c, a, a}, {a, c, a}, {a, a, c}, {a, a, a}};
        const double W[5] = {d, e, e, e, e};
        for (int q = 0; q < 5; q++) { pc[3 * q] = P[q][0]; pc[3 * q + 1] = P[q][1]; pc[3 * q + 2] = P[q][2]; w[q] = W[q] / 6; }
        return 5;
    }
    return -1;
}
void efo_bfun_t4(double r, double s, double t, double *N) { N[0] = (1 - r - s - t); N[1] = r; N[2] = s; N[3] = t; }

static double jacjac3(const double *xyz, const int64_t *nodes, double J[3][3])
{
    const double *x = xyz + 3 * (nodes[0] - 1);
    for (int i = 0; i < 3; i++) for (int k = 0; k < 3; k++) J[i][k] = x[i] * T4_GP[0][k];
    for (int n = 1; n < 4; n++) {
        x = xyz + 3 * (nodes[n] - 1);
        for (int i = 0; i < 3; i++) for (int k = 0; k < 3; k++) J[i][k] = J[i][k] + x[i] * T4_GP[n][k];
    }
    return (+J[0][0] * (J[1][1] * J[2][2] - J[2][1] * J[1][2])
            - J[0][1] * (J[1][0] * J[2][2] - J[1][2] * J[2][0])
            + J[0][2] * (J[1][0] * J[2][1] - J[1][1] * J[2][0]));
}
static void bfungrad3(double J[3][3], double g[4][3])
{
    double a[3][3];                                   /* a = Jac' */
    for (int r = 0; r < 3; r++) for (int c = 0; c < 3; c++) a[r][c] = J[c][r];
    /* det(a): x0, x1, x2 = columns of a; dot(x0, cross(x1, x2)) */
    const double c0 = a[1][1] * a[2][2] - a[2][1] * a[1][2];
    const double c1 = a[2][1] * a[0][2] - a[0][1] * a[2][2];
    const double c2 = a[0][1] * a[1][2] - a[1][1] * a[0][2];
    const double d = (a[0][0] * c0 + a[1][0] * c1) + a[2][0] * c2;
    for (int j = 0; j < 4; j++) {
        const double *b = T4_GP[j];
        g[j][0] = (((a[1][1] * a[2][2] - a[1][2] * a[2][1]) * b[0] + (a[0][2] * a[2][1] - a[0][1] * a[2][2]) * b[1]) + (a[0][1] * a[1][2] - a[0][2] * a[1][1]) * b[2]) / d;
        g[j][1] = (((a[1][2] * a[2][0] - a[1][0] * a[2][2]) * b[0] + (a[0][0] * a[2][2] - a[0][2] * a[2][0]) * b[1]) + (a[0][2] * a[1][0] - a[0][0] * a[1][2]) * b[2]) / d;
        g[j][2] = (((a[1][0] * a[2][1] - a[1][1] * a[2][0]) * b[0] + (a[0][1] * a[2][0] - a[0][0] * a[2][1]) * b[1]) + (a[0][0] * a[1][1] - a[0][1] * a[1][0]) * b[2]) / d;
    }
}

/* ------------------------------------------------------------------------------------------
 * COO buffer = SysmatAssemblerSparse (src/Assemblers.jl:19-24), caller-allocated.
 * ------------------------------------------------------------------------------------------ */
/* The sink of assemble!: mode 0 is the reference's literal COO buffer (three growing vectors, 24 B per triplet).
 * Modes 1-3 are the DIRECT-ACCUMULATE oracle: the same element traversal and the same left-to-right
 * sums, but into a pre-built CSC pattern of a column block [c0, c1] instead of a COO list, so that matrices of the
 * BASELINE sizes can be checked without 24 B/triplet + sparse() scratch:
 *   1 = count the appended triplets of every column of the block, 2 = record their row indices (-> pattern),
 *   3 = nzval[slot(row, col)] = nzval[...] + v in append order.  nzval starts as -0.0, and -0.0 + v == v exactly
 *       (also for v = +-0.0), so the first contribution is "assigned" and later ones are folded left to right --
 *       what sparse() does with duplicates (efo_sparse below).  Tested == mode 0 + efo_sparse bit for bit. */
typedef struct {
    int mode;
    int64_t *row, *col; double *val; int64_t n;          /* mode 0 */
    int64_t nrow, ncol, c0, c1;                           /* modes 1-3: column block, 1-based inclusive */
    int64_t *cnt, *cand;                                  /* modes 1-2: per-column counts / fill cursors, candidate rows */
    const int64_t *colptr, *rowval; double *nzval;        /* mode 3: pattern of the block (colptr 1-based, rebased) */
    int bad;                                              /* an index was < 1 or > nrow/ncol (sparse(): ArgumentError) */
} coo;

static inline void coo_put(coo *a, int64_t r, int64_t c, double v)
{
    if (a->mode == 0) { a->row[a->n] = r; a->col[a->n] = c; a->val[a->n] = v; a->n++; return; }
    a->n++;
    if (r < 1 || r > a->nrow || c < 1 || c > a->ncol) { a->bad = 1; return; }
    if (c < a->c0 || c > a->c1) return;
    const int64_t lc = c - a->c0;
    if (a->mode == 1) { a->cnt[lc]++; return; }
    if (a->mode == 2) { a->cand[a->cnt[lc]++] = r; return; }
    int64_t lo = a->colptr[lc] - 1, hi = a->colptr[lc + 1] - 1;
    while (lo < hi) { int64_t mid = (lo + hi) >> 1; if (a->rowval[mid] < r) lo = mid + 1; else hi = mid; }
    a->nzval[lo] = a->nzval[lo] + v;
}

/* assemble!(self, lma): src/Assemblers.jl:97-102 with init! ordering src/LocalAssemblers.jl:68-83:
 * column-major, k = j*nr + i, row = rdofs[i], col = cdofs[j]. M is column-major nr x nc. */
static void coo_append(coo *a, int nr, int nc, const int64_t *rdofs, const int64_t *cdofs, const double *M)
{
    for (int j = 0; j < nc; j++)
        for (int i = 0; i < nr; i++) coo_put(a, rdofs[i], cdofs[j], M[j * nr + i]);
}
/* assemble!(self, transpose(lma)): src/Assemblers.jl:109-114.  Iterating transpose(parent) in
 * column-major order of the transposed view: outer = parent row i, inner = parent column j. */
static void coo_append_T(coo *a, int nr, int nc, const int64_t *rdofs, const int64_t *cdofs, const double *M)
{
    for (int i = 0; i < nr; i++)
        for (int j = 0; j < nc; j++) coo_put(a, cdofs[j], rdofs[i], M[j * nr + i]);
}

/* _storedofs!: src/FEIterators.jl:161-171 -- node-major, component-minor. */
static void eldofs(const int64_t *nodes, int nen, const int64_t *dofnums, int ncomp, int64_t *d)
{
    int p = 0;
    for (int k = 0; k < nen; k++)
        for (int i = 0; i < ncomp; i++)
            d[p++] = dofnums[(nodes[k] - 1) * ncomp + i];
}

/* the same for an element with a cell field: the dofs of the dim-0 entities first, then those of the cell e (0-based) */
static void eldofs_fe(int fe, const int64_t *nodes, const int64_t *dofnums, const int64_t *celldofnums, int64_t e, int64_t *d)
{
    const int nn = fe_nodedofs(fe);
    eldofs(nodes, nn, dofnums, 1, d);
    if (fe_celldofs(fe)) d[nn] = celldofnums[e];
}

/* B(g,k): examples/elasticity/stretch/t6.jl:42 */
static void Bmat(const double g[2], int k, double b[3])
{
    if (k == 1) { b[0] = g[0]; b[1] = 0; b[2] = g[1]; }
    else        { b[0] = 0; b[1] = g[1]; b[2] = g[0]; }
}
/* D*b for SMatrix{3,3} (column-major storage Dcm[c*3+r]) times SVector{3}: StaticArrays unrolled
 * row sums ((D[r,1] b1 + D[r,2] b2) + D[r,3] b3). */
static void Dmul(const double *Dcm, const double b[3], double o[3])
{
    for (int r = 0; r < 3; r++)
        o[r] = (Dcm[0 * 3 + r] * b[0] + Dcm[1 * 3 + r] * b[1]) + Dcm[2 * 3 + r] * b[2];
}
static double dot3(const double a[3], const double b[3]) { return (a[0] * b[0] + a[1] * b[1]) + a[2] * b[2]; }
static double dot2(const double a[2], const double b[2]) { return a[0] * b[0] + a[1] * b[1]; }

/* Number of COO triplets each element appends for a form. */
int64_t efo_triplets_per_element(int form, int vkind, int pkind)
{
    int64_t nu = efo_nbf(vkind), np = efo_nbf(pkind);   /* (element kinds incl. EFO_T3B / EFO_L2) */
    switch (form) {
    case EFO_FORM_HEAT: return nu * nu;
    case EFO_FORM_ELASTICITY: return 4 * nu * nu;
    case EFO_FORM_STOKES_GEN:
    case EFO_FORM_STOKES_VECLAP_ALT: return 4 * nu * nu + 2 * (2 * nu * np);
    case EFO_FORM_STOKES_REDDY: return 4 * nu * nu + 4 * nu * np;
    case EFO_FORM_STOKES_VECLAP: return 2 * nu * nu + 4 * nu * np;
    }
    return -1;
}

/* ------------------------------------------------------------------------------------------
 * The element loop (the user-written integrate! closures), appending COO triplets.
 *
 * mesh 0 (vconn/vkind/vxy): the mesh of space 0 (heat/elasticity: the only mesh; Stokes: velocity mesh).
 * mesh 1 (pconn/pkind/pxy): Stokes pressure mesh (T6toT3), else NULL.
 * dof0/dof1/dof2: FEField.dofnums of spaces 0..2; ncomp = components per node.
 *   HEAT:        space0 scalar.          params = [kappa]
 *   ELASTICITY:  space0 2 comps.         params = D (3x3 column-major, 9 doubles)
 *   STOKES_GEN:  space0 = Uh (2 comps), space1 = Ph.             params = D (9)
 *   STOKES_VECLAP_ALT: same spaces.                               params = [mu]
 *   STOKES_REDDY / STOKES_VECLAP: space0 = ux, space1 = uy, space2 = p.  params = [mu]
 * quad: triangle npts or square Gauss order (the QPIterator settings).
 * Elements [e0, e1) are processed (0-based range) so the baseline can time a bounded sample.
 * Returns number of triplets appended, or <0 on error.
 * ------------------------------------------------------------------------------------------ */
static int64_t element_loop(coo *ap, int form, int quad, int64_t e0, int64_t e1, const int64_t *elist,
                            const int64_t *vconn, int vkind, const double *vxy,
                            const int64_t *pconn, int pkind, const double *pxy,
                            const int64_t *dof0, const int64_t *dof1, const int64_t *dof2,
                            const double *params, const efo_ext *x)
{
#define a (*ap)
    const int values = (a.mode == 0 || a.mode == 3);   /* pattern passes skip the quadrature loop (ke stays 0) */
    /* elements of the velocity / pressure spaces: the H1 element of the mesh unless row f5's ext block says otherwise */
    const int vfe = (x && x->vfe) ? x->vfe : vkind, pfe = (x && x->pfe) ? x->pfe : pkind;
    if ((vfe != vkind || pfe != pkind) && form != EFO_FORM_STOKES_REDDY && form != EFO_FORM_STOKES_VECLAP) return -2;
    qptab vq, pq;
    if (vkind == EFO_T4) { if (form != EFO_FORM_HEAT) return -2; vq.nbf = 4; vq.npts = 0; }
    else if (qptab_init_shape(&vq, vfe, vkind, quad) < 0) return -1;
    if (pconn && qptab_init_shape(&pq, pfe, pkind, quad) < 0) return -1;
    const int nu = vq.nbf;
    double J[2][2];
    double g[MAXBF][2];

    if (form == EFO_FORM_HEAT && vkind == EFO_T4) { /* examples/heat/poisson/t4.jl:31-57; quad = npts of the tetrahedron rule */
        const double kappa = params[0];
        double pc3[3 * 5], w3[5], J3[3][3], g3[4][3];
        const int np3 = efo_quadrature_t4(quad, pc3, w3);
        if (np3 < 0) return -1;
        int64_t d[4]; double ke[16];
        for (int64_t ee = e0; ee < e1; ee++) {
            const int64_t e = elist ? elist[ee] : ee;
            const int64_t *nodes = vconn + e * 4;
            eldofs(nodes, 4, dof0, 1, d);
            memset(ke, 0, sizeof ke);
            for (int q = 0; values && q < np3; q++) {
                double Jd = jacjac3(vxy, nodes, J3);
                bfungrad3(J3, g3);
                double JxW = Jd * w3[q];
                for (int j = 0; j < 4; j++)
                    for (int i = 0; i < 4; i++)
                        ke[j * 4 + i] = ke[j * 4 + i] + ((g3[i][0] * g3[j][0] + g3[i][1] * g3[j][1]) + g3[i][2] * g3[j][2]) * (kappa * JxW);
            }
            coo_append(&a, 4, 4, d, d, ke);
        }
        return a.n;
    }

    if (form == EFO_FORM_HEAT) { /* examples/heat/poisson/t3.jl:41-64 */
        const double kappa = params[0];
        int64_t d[MAXBF]; double ke[MAXBF * MAXBF];
        for (int64_t ee = e0; ee < e1; ee++) {
            const int64_t e = elist ? elist[ee] : ee;
            const int64_t *nodes = vconn + e * nu;
            eldofs(nodes, nu, dof0, 1, d);
            memset(ke, 0, sizeof ke);
            for (int q = 0; values && q < vq.npts; q++) {
                double Jd = jacjac(vxy, nodes, nu, vq.gp[q], J);
                bfungrad(nu, vq.gp[q], J, g);
                double JxW = Jd * vq.w[q];
                for (int j = 0; j < nu; j++)
                    for (int i = 0; i < nu; i++)
                        ke[j * nu + i] = ke[j * nu + i] + dot2(g[i], g[j]) * (kappa * JxW);
            }
            coo_append(&a, nu, nu, d, d, ke);
        }
        return a.n;
    }

    if (form == EFO_FORM_ELASTICITY) { /* examples/elasticity/stretch/t6.jl:40-63 */
        const int nd = 2 * nu;
        int64_t d[2 * MAXBF]; double ke[4 * MAXBF * MAXBF];
        for (int64_t ee = e0; ee < e1; ee++) {
            const int64_t e = elist ? elist[ee] : ee;
            const int64_t *nodes = vconn + e * nu;
            eldofs(nodes, nu, dof0, 2, d);
            memset(ke, 0, sizeof ke);
            for (int q = 0; values && q < vq.npts; q++) {
                double Jd = jacjac(vxy, nodes, nu, vq.gp[q], J);
                bfungrad(nu, vq.gp[q], J, g);
                double JxW = Jd * vq.w[q];
                for (int j = 0; j < nd; j++) {
                    double Bj[3], DBj[3];
                    Bmat(g[j / 2], j % 2 + 1, Bj); /* edofbfnum/edofcompnt: src/FESpaces.jl:87-105 */
                    Dmul(params, Bj, DBj);
                    for (int i = 0; i < nd; i++) {
                        double Bi[3];
                        Bmat(g[i / 2], i % 2 + 1, Bi);
                        ke[j * nd + i] = ke[j * nd + i] + dot3(DBj, Bi) * JxW;
                    }
                }
            }
            coo_append(&a, nd, nd, d, d, ke);
        }
        return a.n;
    }

    if (form == EFO_FORM_STOKES_GEN || form == EFO_FORM_STOKES_VECLAP_ALT) {
        /* gen: examples/stokes/colliding_flow/ht_p2_p1_gen.jl:46-90
         * veclap_alt: ht_p2_p1_veclap_alt.jl:60-98, test/test_stokes.jl:629-668 */
        const int nd = 2 * nu, np = pkind;
        int64_t du[2 * MAXBF], dp[MAXBF];
        double kuu[4 * MAXBF * MAXBF], kup[2 * MAXBF * MAXBF];
        for (int64_t ee = e0; ee < e1; ee++) {
            const int64_t e = elist ? elist[ee] : ee;
            const int64_t *unodes = vconn + e * nu, *pnodes = pconn + e * np;
            eldofs(unodes, nu, dof0, 2, du);
            eldofs(pnodes, np, dof1, 1, dp);
            memset(kuu, 0, sizeof kuu); memset(kup, 0, sizeof kup);
            for (int q = 0; values && q < vq.npts; q++) {
                double Jd = jacjac(vxy, unodes, nu, vq.gp[q], J); /* velocity element Jacobian (:61) */
                double JxW = Jd * vq.w[q];
                bfungrad(nu, vq.gp[q], J, g);
                const double *Np = pq.N[q];
                if (form == EFO_FORM_STOKES_GEN) {
                    for (int j = 0; j < nd; j++) {
                        double Bj[3], DBj[3];
                        Bmat(g[j / 2], j % 2 + 1, Bj);
                        Dmul(params, Bj, DBj);
                        for (int i = 0; i < nd; i++) {
                            double Bi[3];
                            Bmat(g[i / 2], i % 2 + 1, Bi);
                            kuu[j * nd + i] = kuu[j * nd + i] + dot3(Bi, DBj) * (JxW);
                        }
                    }
                } else {
                    const double mu = params[0];
                    for (int j = 0; j < nd; j++)
                        for (int i = 0; i < nd; i++)
                            if (i % 2 == j % 2)
                                kuu[j * nd + i] = kuu[j * nd + i] + (mu * JxW) * (dot2(g[i / 2], g[j / 2]));
                }
                for (int j = 0; j < np; j++)
                    for (int i = 0; i < nd; i++)
                        kup[j * nd + i] = kup[j * nd + i] + (-JxW * Np[j]) * g[i / 2][i % 2];
            }
            coo_append(&a, nd, nd, du, du, kuu);
            coo_append(&a, nd, np, du, dp, kup);
            coo_append_T(&a, nd, np, du, dp, kup);
        }
        return a.n;
    }

    if (form == EFO_FORM_STOKES_REDDY || form == EFO_FORM_STOKES_VECLAP) {
        /* Reddy: examples/stokes/colliding_flow/ht_p2_p1.jl:55-113, test/test_stokes.jl:374-422
         * veclap: examples/stokes/colliding_flow/ht_p2_p1_veclap.jl:55-106
         * row f5, the same loop on other element pairs (one mesh: pconn == vconn):
         *   FEH1_T3_BUBBLE / FEH1_T3   examples/stokes/colliding_flow/p1b_p1.jl:53-109, test/test_stokes.jl:190-247
         *   FEH1_Q4 / FEL2_Q4           examples/stokes/colliding_flow/q1_q0.jl:52-108 (Jacobian of the VELOCITY element :70) */
        const int np = pq.nbf;
        if (!pconn) return -2;
        if (pfe == EFO_L2 && vfe != vkind) return -2;     /* (no example pairs a bubble velocity with an L2 pressure) */
        const double mu = params[0];
        int64_t dx[MAXBF], dy[MAXBF], dp[MAXBF];
        double kxx[MAXBF * MAXBF], kyy[MAXBF * MAXBF], kxy[MAXBF * MAXBF], kxp[MAXBF * MAXBF], kyp[MAXBF * MAXBF];
        double gp_[MAXBF][2];
        for (int64_t ee = e0; ee < e1; ee++) {
            const int64_t e = elist ? elist[ee] : ee;
            const int64_t *unodes = vconn + e * vkind, *pnodes = pconn + e * pkind;
            eldofs_fe(vfe, unodes, dof0, x ? x->cdof0 : NULL, e, dx);
            eldofs_fe(vfe, unodes, dof1, x ? x->cdof1 : NULL, e, dy);
            eldofs_fe(pfe, pnodes, dof2, x ? x->cdof2 : NULL, e, dp);
            memset(kxx, 0, sizeof kxx); memset(kyy, 0, sizeof kyy); memset(kxy, 0, sizeof kxy);
            memset(kxp, 0, sizeof kxp); memset(kyp, 0, sizeof kyp);
            for (int q = 0; values && q < vq.npts; q++) {
                double Jd = pfe == EFO_L2 ? jacjac(vxy, unodes, vkind, vq.gp[q], J)    /* q1_q0.jl:70: jacjac(uxel, uxqp) */
                                          : jacjac(pxy, pnodes, pkind, pq.gp[q], J);   /* PRESSURE element Jacobian (:72) */
                double JxW = Jd * pq.w[q];
                bfungrad(np, pq.gp[q], J, gp_); /* gradNp: computed, unused (:74) */
                bfungrad(nu, vq.gp[q], J, g);   /* gradNux == gradNuy numerically */
                const double *Np = pq.N[q];
                if (form == EFO_FORM_STOKES_REDDY) {
                    for (int j = 0; j < nu; j++)
                        for (int i = 0; i < nu; i++)
                            kxx[j * nu + i] = kxx[j * nu + i] + (mu * JxW) * (2 * g[i][0] * g[j][0] + g[i][1] * g[j][1]);
                    for (int j = 0; j < nu; j++)
                        for (int i = 0; i < nu; i++)
                            kyy[j * nu + i] = kyy[j * nu + i] + (mu * JxW) * (g[i][0] * g[j][0] + 2 * g[i][1] * g[j][1]);
                    for (int j = 0; j < nu; j++)
                        for (int i = 0; i < nu; i++)
                            kxy[j * nu + i] = kxy[j * nu + i] + (mu * JxW) * (g[i][0] * g[j][1]);
                } else {
                    for (int j = 0; j < nu; j++)
                        for (int i = 0; i < nu; i++)
                            kxx[j * nu + i] = kxx[j * nu + i] + (mu * JxW) * dot2(g[i], g[j]);
                    for (int j = 0; j < nu; j++)
                        for (int i = 0; i < nu; i++)
                            kyy[j * nu + i] = kyy[j * nu + i] + (mu * JxW) * dot2(g[i], g[j]);
                }
                for (int j = 0; j < np; j++)
                    for (int i = 0; i < nu; i++)
                        kxp[j * nu + i] = kxp[j * nu + i] + (-JxW) * (g[i][0] * Np[j]);
                for (int j = 0; j < np; j++)
                    for (int i = 0; i < nu; i++)
                        kyp[j * nu + i] = kyp[j * nu + i] + (-JxW) * (g[i][1] * Np[j]);
            }
            if (form == EFO_FORM_STOKES_REDDY) { /* ht_p2_p1.jl:94-101 */
                coo_append(&a, nu, nu, dx, dx, kxx);
                coo_append(&a, nu, nu, dx, dy, kxy);
                coo_append_T(&a, nu, nu, dx, dy, kxy);
                coo_append(&a, nu, nu, dy, dy, kyy);
            } else {                             /* ht_p2_p1_veclap.jl:89-94 */
                coo_append(&a, nu, nu, dx, dx, kxx);
                coo_append(&a, nu, nu, dy, dy, kyy);
            }
            coo_append(&a, nu, np, dx, dp, kxp);
            coo_append_T(&a, nu, np, dx, dp, kxp);
            coo_append(&a, nu, np, dy, dp, kyp);
            coo_append_T(&a, nu, np, dy, dp, kyp);
        }
        return a.n;
    }
    return -2;
#undef a
}

int64_t efo_assemble_coo(int form, int quad, int64_t e0, int64_t e1,
                         const int64_t *vconn, int vkind, const double *vxy,
                         const int64_t *pconn, int pkind, const double *pxy,
                         const int64_t *dof0, const int64_t *dof1, const int64_t *dof2,
                         const double *params,
                         int64_t *row, int64_t *col, double *val, const efo_ext *ext)
{
    coo a;
    memset(&a, 0, sizeof a);
    a.row = row; a.col = col; a.val = val;
    return element_loop(&a, form, quad, e0, e1, NULL, vconn, vkind, vxy, pconn, pkind, pxy, dof0, dof1, dof2, params, ext);
}

/* ------------------------------------------------------------------------------------------
 * DIRECT-ACCUMULATE oracle for a column block [c0, c1] (1-based inclusive) of the nrow x ncol matrix.
 * Elements visited: elist[0..nsel) (ascending element numbers, e.g. the elements touching the block) or all of
 * [0, nel) when elist is NULL -- elements without a column in the block contribute nothing either way.
 *   efo_direct_pattern: colptr (c1-c0+2 entries, 1-based, rebased to the block) and rowval (capacity = the count the
 *     first call returns when rowval is NULL): the rows of every column ascending, duplicates removed, explicit zeros
 *     kept -- the pattern sparse() produces (efo_sparse).  Returns nnz of the block, -1 ArgumentError, -3 no memory.
 *   efo_direct_values: nzval of that pattern, every nonzero = left-to-right sum of its contributions in append order
 *     (ascending element, the reference's assemble! order inside an element).
 * ------------------------------------------------------------------------------------------ */
static int cmp_i64(const void *x, const void *y)
{
    const int64_t p = *(const int64_t *)x, q = *(const int64_t *)y;
    return (p > q) - (p < q);
}

int64_t efo_direct_pattern(int form, int quad, int64_t nel, const int64_t *elist, int64_t nsel,
                           const int64_t *vconn, int vkind, const double *vxy,
                           const int64_t *pconn, int pkind, const double *pxy,
                           const int64_t *dof0, const int64_t *dof1, const int64_t *dof2,
                           int64_t nrow, int64_t ncol, int64_t c0, int64_t c1,
                           int64_t *colptr, int64_t *rowval, const efo_ext *ext)
{
    const int64_t nc = c1 - c0 + 1;
    const double dummy[9] = {0};
    coo a;
    memset(&a, 0, sizeof a);
    a.nrow = nrow; a.ncol = ncol; a.c0 = c0; a.c1 = c1;
    int64_t *cnt = (int64_t *)calloc((size_t)nc + 1, sizeof(int64_t));
    if (!cnt) return -3;
    a.mode = 1; a.cnt = cnt;
    const int64_t n_it = elist ? nsel : nel;
    if (element_loop(&a, form, quad, 0, n_it, elist, vconn, vkind, vxy, pconn, pkind, pxy, dof0, dof1, dof2, dummy, ext) < 0 || a.bad) {
        free(cnt);
        return -1;
    }
    /* counts -> start offsets (cursor array for the fill pass) */
    int64_t tot = 0;
    for (int64_t j = 0; j < nc; j++) { const int64_t c = cnt[j]; cnt[j] = tot; tot += c; }
    cnt[nc] = tot;
    int64_t *start = (int64_t *)malloc(((size_t)nc + 1) * sizeof(int64_t));
    int64_t *cand = (int64_t *)malloc((size_t)(tot > 0 ? tot : 1) * sizeof(int64_t));
    if (!start || !cand) { free(cnt); free(start); free(cand); return -3; }
    memcpy(start, cnt, ((size_t)nc + 1) * sizeof(int64_t));
    a.mode = 2; a.cand = cand; a.n = 0;
    element_loop(&a, form, quad, 0, n_it, elist, vconn, vkind, vxy, pconn, pkind, pxy, dof0, dof1, dof2, dummy, ext);
    int64_t nnz = 0;
    colptr[0] = 1;
    for (int64_t j = 0; j < nc; j++) {
        int64_t *r = cand + start[j];
        const int64_t m = start[j + 1] - start[j];
        qsort(r, (size_t)m, sizeof(int64_t), cmp_i64);
        for (int64_t k = 0; k < m; k++)
            if (k == 0 || r[k] != r[k - 1]) { if (rowval) rowval[nnz] = r[k]; nnz++; }
        colptr[j + 1] = nnz + 1;
    }
    free(cnt); free(start); free(cand);
    return nnz;
}

int64_t efo_direct_values(int form, int quad, int64_t nel, const int64_t *elist, int64_t nsel,
                          const int64_t *vconn, int vkind, const double *vxy,
                          const int64_t *pconn, int pkind, const double *pxy,
                          const int64_t *dof0, const int64_t *dof1, const int64_t *dof2,
                          const double *params, int64_t nrow, int64_t ncol, int64_t c0, int64_t c1,
                          const int64_t *colptr, const int64_t *rowval, double *nzval, const efo_ext *ext)
{
    coo a;
    memset(&a, 0, sizeof a);
    a.mode = 3; a.nrow = nrow; a.ncol = ncol; a.c0 = c0; a.c1 = c1;
    a.colptr = colptr; a.rowval = rowval; a.nzval = nzval;
    const int64_t nnz = colptr[c1 - c0 + 1] - 1;
    for (int64_t k = 0; k < nnz; k++) nzval[k] = -0.0;
    const int64_t n_it = elist ? nsel : nel;
    const int64_t r = element_loop(&a, form, quad, 0, n_it, elist, vconn, vkind, vxy, pconn, pkind, pxy, dof0, dof1, dof2, params, ext);
    if (r < 0) return r;
    return a.bad ? -1 : nnz;
}

/* ------------------------------------------------------------------------------------------
 * finish!: src/Assemblers.jl:121-123 -> SparseArrays.sparse(I, J, V, m, n) (Julia stdlib sparse!):
 *  (1) counting sort of the triplets by row into CSR, preserving input order inside a row;
 *  (2) per row, a later duplicate column is folded into the FIRST occurrence: v = v + dup
 *      (left to right in input order); numerical zeros are NOT dropped;
 *  (3) CSR -> CSC by counting sort on columns => row indices ascending within a column;
 *  (4) colptr is Int64, 1-based, length n+1.
 * Any index < 1 or > m / n is an ArgumentError in Julia: return -1 here.
 * colptr: ncol+1; rowval/nzval: capacity ntrip.  Returns nnz.
 * ------------------------------------------------------------------------------------------ */
int64_t efo_sparse(int64_t nrow, int64_t ncol, int64_t ntrip,
                   const int64_t *I, const int64_t *Jc, const double *V,
                   int64_t *colptr, int64_t *rowval, double *nzval)
{
    int64_t *rowptr = (int64_t *)calloc((size_t)nrow + 2, sizeof(int64_t));
    int64_t *ccol = (int64_t *)malloc((size_t)(ntrip > 0 ? ntrip : 1) * sizeof(int64_t));
    double *cval = (double *)malloc((size_t)(ntrip > 0 ? ntrip : 1) * sizeof(double));
    int64_t *last = (int64_t *)malloc((size_t)(ncol > 0 ? ncol : 1) * sizeof(int64_t));
    if (!rowptr || !ccol || !cval || !last) { free(rowptr); free(ccol); free(cval); free(last); return -3; }
    for (int64_t k = 0; k < ntrip; k++) {
        if (I[k] < 1 || I[k] > nrow || Jc[k] < 1 || Jc[k] > ncol) { free(rowptr); free(ccol); free(cval); free(last); return -1; }
        rowptr[I[k] + 1]++;
    }
    /* rowptr[i+1] = start of row i (1-based i) while filling */
    for (int64_t i = 1; i <= nrow; i++) rowptr[i + 1] += rowptr[i];
    /* now rowptr[i] = number of entries in rows < i ... shift: start(i) = rowptr[i] */
    for (int64_t k = 0; k < ntrip; k++) {
        int64_t p = rowptr[I[k]]++;
        ccol[p] = Jc[k]; cval[p] = V[k];
    }
    /* after fill rowptr[i] = end of row i = start of row i+1; start of row 1 = 0 */
    for (int64_t j = 0; j < ncol; j++) last[j] = -1;
    for (int64_t j = 0; j <= ncol; j++) colptr[j] = 0;
    int64_t w = 0, start = 0;
    for (int64_t i = 1; i <= nrow; i++) {
        int64_t end = rowptr[i];
        int64_t wrow = w;
        for (int64_t p = start; p < end; p++) {
            int64_t j = ccol[p] - 1;
            if (last[j] >= wrow) {
                cval[last[j]] = cval[last[j]] + cval[p];
            } else {
                last[j] = w; ccol[w] = ccol[p]; cval[w] = cval[p]; w++;
                colptr[j + 1]++;
            }
        }
        start = end;
        rowptr[i] = w; /* compacted end of row i */
    }
    int64_t nnz = w;
    /* colptr: counts -> 1-based pointers */
    colptr[0] = 1;
    for (int64_t j = 0; j < ncol; j++) colptr[j + 1] += colptr[j];
    /* transpose compacted CSR into CSC */
    int64_t *next = last; /* reuse */
    for (int64_t j = 0; j < ncol; j++) next[j] = colptr[j] - 1;
    start = 0;
    for (int64_t i = 1; i <= nrow; i++) {
        int64_t end = rowptr[i];
        for (int64_t p = start; p < end; p++) {
            int64_t q = next[ccol[p] - 1]++;
            rowval[q] = i; nzval[q] = cval[p];
        }
        start = end;
    }
    free(rowptr); free(ccol); free(cval); free(last);
    return nnz;
}

/* ------------------------------------------------------------------------------------------
 * System VECTOR assembly: SysvecAssembler + LocalVectorAssembler.
 *   start!(av, nrow)         src/Assemblers.jl:196-200   val = zeros(nrow)
 *   init!(fe, eldofs(el))    src/LocalAssemblers.jl:147-151   fe.V zeroed per element
 *   fe[j] += N[j] * Q * JxW  examples/heat/poisson/t3.jl:57, q4.jl:47, test/test_heat.jl:53,167
 *                            (Julia parses N[j]*Q*JxW as (N[j]*Q)*JxW; JxW = J * weight(qp))
 *   assemble!(av, fe)        src/Assemblers.jl:217-223   val[gi] += V[i], i ascending, element by element
 *   finish!(av)              src/Assemblers.jl:230-232   copy of val
 * Returns 0, or -1 for an unavailable rule, -2 for a dof number outside 1..nrow (Julia: BoundsError).
 * ------------------------------------------------------------------------------------------ */
int64_t efo_assemble_vec_heat(int quad, int64_t e0, int64_t e1, const int64_t *conn, int kind, const double *xy,
                              const int64_t *dofnums, double Q, int64_t nrow, double *val)
{
    if (kind == EFO_T4) { /* examples/heat/poisson/t4.jl:41-55; xy is 3 x nnodes */
        double pc3[3 * 5], w3[5], J3[3][3], N4[5][4];
        const int np3 = efo_quadrature_t4(quad, pc3, w3);
        if (np3 < 0) return -1;
        for (int q = 0; q < np3; q++) efo_bfun_t4(pc3[3 * q], pc3[3 * q + 1], pc3[3 * q + 2], N4[q]);
        int64_t d4[4]; double f4[4];
        for (int64_t i = 0; i < nrow; i++) val[i] = 0.0;
        for (int64_t e = e0; e < e1; e++) {
            const int64_t *nodes = conn + e * 4;
            eldofs(nodes, 4, dofnums, 1, d4);
            for (int j = 0; j < 4; j++) f4[j] = 0.0;
            for (int q = 0; q < np3; q++) {
                double JxW = jacjac3(xy, nodes, J3) * w3[q];
                for (int j = 0; j < 4; j++) f4[j] = f4[j] + (N4[q][j] * Q) * JxW;
            }
            for (int i = 0; i < 4; i++) {
                if (d4[i] < 1 || d4[i] > nrow) return -2;
                val[d4[i] - 1] = val[d4[i] - 1] + f4[i];
            }
        }
        return 0;
    }
    qptab vq;
    if (qptab_init(&vq, kind, quad) < 0) return -1;
    const int nu = kind;
    double J[2][2];
    int64_t d[MAXBF]; double fe[MAXBF];
    for (int64_t i = 0; i < nrow; i++) val[i] = 0.0;
    for (int64_t e = e0; e < e1; e++) {
        const int64_t *nodes = conn + e * nu;
        eldofs(nodes, nu, dofnums, 1, d);
        for (int j = 0; j < nu; j++) fe[j] = 0.0;
        for (int q = 0; q < vq.npts; q++) {
            double Jd = jacjac(xy, nodes, nu, vq.gp[q], J);
            double JxW = Jd * vq.w[q];
            for (int j = 0; j < nu; j++) fe[j] = fe[j] + (vq.N[q][j] * Q) * JxW;
        }
        for (int i = 0; i < nu; i++) {
            if (d[i] < 1 || d[i] > nrow) return -2;
            val[d[i] - 1] = val[d[i] - 1] + fe[i];
        }
    }
    return 0;
}

/* ------------------------------------------------------------------------------------------
 * y = K * x for a SparseMatrixCSC (call site examples/heat/poisson/t3.jl:78
 * `KT = K * T`).  SparseArrays' mul!(C, A, B, 1, 0): C zeroed, then for every column k in order,
 * for every stored entry of the column: C[rowval] += nzval * x[k] -- a rounded product followed
 * by a rounded sum, so y[r] accumulates its terms in ascending column order starting from 0.0.
 * (restated from SparseArrays' published algorithm; the stdlib is not vendored: "parity unpinned"
 * at ULP level, pinned end to end through the heat goldens.)
 * ------------------------------------------------------------------------------------------ */
void efo_spmv_csc(int64_t nrow, int64_t ncol, const int64_t *colptr, const int64_t *rowval, const double *nzval,
                  const double *x, double *y)
{
    for (int64_t i = 0; i < nrow; i++) y[i] = 0.0;
    for (int64_t k = 0; k < ncol; k++) {
        const double xk = x[k];
        for (int64_t p = colptr[k] - 1; p < colptr[k + 1] - 1; p++) y[rowval[p] - 1] = y[rowval[p] - 1] + nzval[p] * xk;
    }
}

/* ------------------------------------------------------------------------------------------
 * Post-processing integrators: location() and the evaluate_*_error loops.
 *   location(el, qp)   src/FEIterators.jl:227-235   loc = x_1*N_1, then loc += x_i*N_i in node order
 *   evaluate_pressure_error / evaluate_velocity_error
 *                      examples/stokes/colliding_flow/ht_p2_p1.jl:120-178, ht_p2_p1_gen.jl:124-153,
 *                      test/test_stokes.jl:438-496:
 *       for el, for qp:  JxW = J*weight;  a_c = 0.0; for j: a_c += dofvals_c[j]*N[j];
 *                        E += JxW * ((a_1 - t_1)^2 [+ (a_2 - t_2)^2]);     return sqrt(E)
 * The true-solution closures are evaluated by the caller at the locations: `truth` is ncomp x npts x nel
 * (component c of point q of element e at truth[(e*npts + q)*ncomp + c]).  eldofvals(el) are read
 * from the system vector U through the dof numbers (what scattersysvec! put into the fields).
 * ------------------------------------------------------------------------------------------ */
int64_t efo_qp_locations(int quad, int64_t nel, const int64_t *conn, int kind, const double *xy, double *out)
{
    qptab vq;
    if (qptab_init(&vq, kind, quad) < 0) return -1;
    for (int64_t e = 0; e < nel; e++) {
        const int64_t *nodes = conn + e * kind;
        for (int q = 0; q < vq.npts; q++) {
            const double *x = xy + 2 * (nodes[0] - 1);
            double lx = x[0] * vq.N[q][0], ly = x[1] * vq.N[q][0];
            for (int i = 1; i < kind; i++) {
                x = xy + 2 * (nodes[i] - 1);
                lx = lx + x[0] * vq.N[q][i]; ly = ly + x[1] * vq.N[q][i];
            }
            out[(e * vq.npts + q) * 2] = lx; out[(e * vq.npts + q) * 2 + 1] = ly;
        }
    }
    return vq.npts;
}

/* dof[c] : dofnums array (ncomp_c x nnodes) of the space of component c, comp[c] its component (0-based),
 * ncs[c] the number of components of that space.  Returns sqrt(E), or -1.0 for an unavailable rule. */
double efo_l2_error(int quad, int64_t nel, const int64_t *conn, int kind, const double *xy, int ncomp,
                    const int64_t *dof0, int ncs0, int comp0, const int64_t *dof1, int ncs1, int comp1,
                    const double *U, const double *truth)
{
    qptab vq;
    if (qptab_init(&vq, kind, quad) < 0) return -1.0;
    double J[2][2];
    double E = 0.0;
    for (int64_t e = 0; e < nel; e++) {
        const int64_t *nodes = conn + e * kind;
        for (int q = 0; q < vq.npts; q++) {
            double Jd = jacjac(xy, nodes, kind, vq.gp[q], J);
            double JxW = Jd * vq.w[q];
            double a0 = 0.0, a1 = 0.0;
            for (int j = 0; j < kind; j++) {
                a0 = a0 + U[dof0[(nodes[j] - 1) * ncs0 + comp0] - 1] * vq.N[q][j];
                if (ncomp > 1) a1 = a1 + U[dof1[(nodes[j] - 1) * ncs1 + comp1] - 1] * vq.N[q][j];
            }
            const double *t = truth + (e * vq.npts + q) * ncomp;
            double d0 = a0 - t[0];
            if (ncomp > 1) { double d1 = a1 - t[1]; E = E + JxW * (d0 * d0 + d1 * d1); }
            else E = E + JxW * (d0 * d0);
        }
    }
    return sqrt(E);
}

/* The same integrator for scalar spaces whose element has a cell dof (row f5): examples/stokes/colliding_flow/
 * p1b_p1.jl:117-172 -- evaluate_velocity_error sums all FOUR basis functions of FEH1_T3_BUBBLE (eldofvals: vertex values
 * then the bubble's, src/FEIterators.jl:201-208); jacjac / location of such an element use the first `kind` (vertex)
 * gradients / functions (src/FElements.jl:148-156 loops over the nodes).  fe = EFO_T3B, or a plain H1 kind. */
double efo_l2_error_fe(int quad, int64_t nel, const int64_t *conn, int kind, const double *xy, int fe, int ncomp,
                       const int64_t *dof0, const int64_t *cdof0, const int64_t *dof1, const int64_t *cdof1,
                       const double *U, const double *truth)
{
    qptab vq;
    if (qptab_init_shape(&vq, fe, kind, quad) < 0) return -1.0;
    double J[2][2];
    double E = 0.0;
    int64_t d0[MAXBF], d1[MAXBF];
    for (int64_t e = 0; e < nel; e++) {
        const int64_t *nodes = conn + e * kind;
        eldofs_fe(fe, nodes, dof0, cdof0, e, d0);
        if (ncomp > 1) eldofs_fe(fe, nodes, dof1, cdof1, e, d1);
        for (int q = 0; q < vq.npts; q++) {
            double Jd = jacjac(xy, nodes, kind, vq.gp[q], J);
            double JxW = Jd * vq.w[q];
            double a0 = 0.0, a1 = 0.0;
            for (int j = 0; j < vq.nbf; j++) {
                a0 = a0 + U[d0[j] - 1] * vq.N[q][j];
                if (ncomp > 1) a1 = a1 + U[d1[j] - 1] * vq.N[q][j];
            }
            const double *t = truth + (e * vq.npts + q) * ncomp;
            double e0 = a0 - t[0];
            if (ncomp > 1) { double e1 = a1 - t[1]; E = E + JxW * (e0 * e0 + e1 * e1); }
            else E = E + JxW * (e0 * e0);
        }
    }
    return sqrt(E);
}
