// ca_device.cuh -- per-trajectory device code of the one-atom path (samplers, beams, Lindblad RHS, DP5).
//
// Everything here is `__host__ __device__` and free of warp intrinsics, so the very same source is run
// (a) by the sm_100a kernels in ca_kernels.cu and (b) on the CPU by tests/hostemu (test-only harness that
// lets the build container, which has no GPU, check this code against the oracle).
//
// Reference map (file:line relative to the reference repository):
//   Philox sampler      <- samples_generate, harmonic branch      src/atom_sampler.jl:84-96
//   phase-noise tables  <- ϕ                                       src/lasernoise_sampler.jl:31-37
//   beams               <- A, A_phase (src/utilities.jl:27-48), Ω (src/rydberg_model.jl:51-74),
//                          simple_flattopHG/LG_field (src/arbitrary_beams.jl:50-81)
//   coefficients        <- GenerateHamiltonian closures            src/rydberg_model.jl:152-188
//   RHS                 <- QuantumOptics dmaster_h! on operators/J src/rydberg_model.jl:25,126-130
//   DP5 + controller    <- OrdinaryDiffEq DP5 / PIController (see oracle/coldatoms_oracle.c header)
#pragma once
#include <cstdint>
#include <cmath>
#include <cstring>

#include "../../include/coldatoms_b200.h"

#if defined(__CUDACC__)
#define CA_HD __host__ __device__ __forceinline__
#else
#define CA_HD inline
#endif

namespace ca {

constexpr double kPi = 3.14159265358979323846264338327950288;
constexpr double kTwoPi = 6.28318530717958647692528676655900577;
constexpr int kMaxModes = 17;  // flat-top orders up to 32

// ---------------------------------------------------------------------------------------------------
// Philox4x32-10: counter = (traj_lo, traj_hi, draw, stream), key = (seed_lo, seed_hi)
// ---------------------------------------------------------------------------------------------------
// CA_NORM_SHORT: precision of the reciprocals / square roots inside the step-size controller's error norm (device code only).
//   2 (default) raw MUFU seeds (relative error ~1e-6 in EEst², i.e. ~1e-7 in the next dt, the size the controller's single-precision
//     log2/exp2 already has), 1 one refinement (~1e-11), 0 the full-precision rcp_d / sqrt_d.  The norm only steers the controller
//     (accept iff EEst <= 1, next dt ~ EEst^-0.17); its serial sqrt -> rcp chains sat on every step's critical path.
#ifndef CA_NORM_SHORT
#define CA_NORM_SHORT 2
#endif
struct U4 { uint32_t x, y, z, w; };

CA_HD uint32_t mulhi32(uint32_t a, uint32_t b) {
#if defined(__CUDA_ARCH__)
    return __umulhi(a, b);
#else
    return (uint32_t)(((uint64_t)a * b) >> 32);
#endif
}

CA_HD U4 philox(uint64_t seed, uint64_t index, uint32_t draw, uint32_t stream) {
    uint32_t c0 = (uint32_t)index, c1 = (uint32_t)(index >> 32), c2 = draw, c3 = stream;
    uint32_t k0 = (uint32_t)seed, k1 = (uint32_t)(seed >> 32);
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        uint32_t h0 = mulhi32(0xD2511F53u, c0), l0 = 0xD2511F53u * c0;
        uint32_t h1 = mulhi32(0xCD9E8D57u, c2), l1 = 0xCD9E8D57u * c2;
        uint32_t n0 = h1 ^ c1 ^ k0, n2 = h0 ^ c3 ^ k1;
        c0 = n0; c1 = l1; c2 = n2; c3 = l0;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
    return U4{c0, c1, c2, c3};
}
CA_HD double u53(uint32_t a, uint32_t b) {
    return ((double)(a >> 5) * 67108864.0 + (double)(b >> 6)) * (1.0 / 9007199254740992.0);
}
CA_HD double u53_open(uint32_t a, uint32_t b) {
    return ((double)(a >> 5) * 67108864.0 + (double)(b >> 6) + 0.5) * (1.0 / 9007199254740992.0);
}
#ifndef CA_TRACE_LL
#define CA_TRACE_LL 1   // ρ_ll from the conserved trace (Runge-Kutta steps keep linear invariants) instead of a second population quadrature
#endif
#ifndef CA_ANCH_V2
#define CA_ANCH_V2 1   // Gaussian-pair anchored coefficients as functions of τ with the noise slopes folded in (see eval_anchored)
#endif
CA_HD int abs_hi32(double x) {   // high word of |x|: orders non-negative doubles like integers
#if defined(__CUDA_ARCH__)
    return __double2hiint(x) & 0x7fffffff;
#else
    long long b;
    memcpy(&b, &x, 8);
    return (int)((b >> 32) & 0x7fffffff);
#endif
}
CA_HD void sincos_d(double a, double* s, double* c) {
#if defined(__CUDA_ARCH__)
    sincos(a, s, c);
#else
    *s = sin(a); *c = cos(a);
#endif
}

// ---------------------------------------------------------------------------------------------------
// Lean fp64 math for the hot loop.  All literals live in ONE constant-memory struct so that they reach the
// FP64 pipe as c[bank][offset] operands (the CUDA math library materialises each 64-bit literal with two UMOVs:
// 15% of all issued instructions in the first profile).  Accuracy ~1 ulp; arguments outside the fast range fall
// back to the library.
// ---------------------------------------------------------------------------------------------------
struct MathConsts {
    double log2e, ln2_hi, ln2_lo, magic;                 // exp
    double ex[14];                                       // 1/k!
    double two_over_pi, pio2_hi, pio2_mid, pio2_lo;      // sincos
    double sn[6], cs[6];                                 // fdlibm kernel_sin / kernel_cos minimax coefficients
    // Dormand-Prince 5(4)
    double a21, a31, a32, a41, a42, a43, a51, a52, a53, a54, a61, a62, a63, a64, a65, b1, b3, b4, b5, b6;
    double c2, c3, c4, c5, e1, e3, e4, e5, e6, e7, d1, d3, d4, d5, d6, d7;
    // the same tableau as rows indexed by stage (rolled stage loop): ra[s][j] = a_sj, rc[s] = c_s, rb/re/rd[j] = b_j, e_j, d_j
    double ra[8][8], rc[8], rb[8], re[8], rd[8];
};
#define CA_MATH_CONSTS_INIT                                                                                                      \
    {1.4426950408889634, 0.6931471805599453, 2.3190468138462996e-17, 6755399441055744.0,                                         \
     {1.0, 1.0, 0.5, 1.0 / 6, 1.0 / 24, 1.0 / 120, 1.0 / 720, 1.0 / 5040, 1.0 / 40320, 1.0 / 362880, 1.0 / 3628800,              \
      1.0 / 39916800, 1.0 / 479001600, 1.0 / 6227020800.0},                                                                      \
     0.6366197723675814, 1.5707963267948966, 6.123233995736766e-17, -1.4973849048591698e-33,                                     \
     {-1.66666666666666324348e-01, 8.33333333332248946124e-03, -1.98412698298579493134e-04, 2.75573137070700676789e-06,          \
      -2.50507602534068634195e-08, 1.58969099521155010221e-10},                                                                  \
     {4.16666666666666019037e-02, -1.38888888888741095749e-03, 2.48015872894767294178e-05, -2.75573143513906633035e-07,          \
      2.08757232129817482790e-09, -1.13596475577881948265e-11},                                                                  \
     1.0 / 5.0, 3.0 / 40.0, 9.0 / 40.0, 44.0 / 45.0, -56.0 / 15.0, 32.0 / 9.0, 19372.0 / 6561.0, -25360.0 / 2187.0,              \
     64448.0 / 6561.0, -212.0 / 729.0, 9017.0 / 3168.0, -355.0 / 33.0, 46732.0 / 5247.0, 49.0 / 176.0, -5103.0 / 18656.0,        \
     35.0 / 384.0, 500.0 / 1113.0, 125.0 / 192.0, -2187.0 / 6784.0, 11.0 / 84.0, 0.2, 0.3, 0.8, 8.0 / 9.0, 71.0 / 57600.0,       \
     -71.0 / 16695.0, 71.0 / 1920.0, -17253.0 / 339200.0, 22.0 / 525.0, -1.0 / 40.0, -12715105075.0 / 11282082432.0,             \
     87487479700.0 / 32700410799.0, -10690763975.0 / 1880347072.0, 701980252875.0 / 199316789632.0,                              \
     -1453857185.0 / 822651844.0, 69997945.0 / 29380423.0,                                                                       \
     {{0}, {0}, {0, 1.0 / 5.0}, {0, 3.0 / 40.0, 9.0 / 40.0}, {0, 44.0 / 45.0, -56.0 / 15.0, 32.0 / 9.0},                        \
      {0, 19372.0 / 6561.0, -25360.0 / 2187.0, 64448.0 / 6561.0, -212.0 / 729.0},                                               \
      {0, 9017.0 / 3168.0, -355.0 / 33.0, 46732.0 / 5247.0, 49.0 / 176.0, -5103.0 / 18656.0}, {0}},                             \
     {0, 0, 0.2, 0.3, 0.8, 8.0 / 9.0, 1.0, 1.0},                                                                                 \
     {0, 35.0 / 384.0, 0, 500.0 / 1113.0, 125.0 / 192.0, -2187.0 / 6784.0, 11.0 / 84.0, 0},                                      \
     {0, 71.0 / 57600.0, 0, -71.0 / 16695.0, 71.0 / 1920.0, -17253.0 / 339200.0, 22.0 / 525.0, -1.0 / 40.0},                     \
     {0, -12715105075.0 / 11282082432.0, 0, 87487479700.0 / 32700410799.0, -10690763975.0 / 1880347072.0,                        \
      701980252875.0 / 199316789632.0, -1453857185.0 / 822651844.0, 69997945.0 / 29380423.0}}
#if defined(__CUDACC__)
static __constant__ MathConsts c_mc = CA_MATH_CONSTS_INIT;
#endif
static const MathConsts h_mc = CA_MATH_CONSTS_INIT;
#if defined(__CUDA_ARCH__)
#define MC c_mc
#else
#define MC h_mc
#endif

CA_HD double rcp_d(double x) {  // 1/x for normal x (no special cases): MUFU seed + 2 Newton steps
#if defined(__CUDA_ARCH__)
    double r;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x));
    double e = fma(-x, r, 1.0);
    r = fma(r, e, r);
    e = fma(-x, r, 1.0);
    return fma(r, e, r);
#else
    return 1.0 / x;
#endif
}
// Short forms for the step-size controller's error norm only (relative error ~1e-11: one refinement of the MUFU seed).
CA_HD double rcp_norm(double x) {
#if defined(__CUDA_ARCH__) && CA_NORM_SHORT
    double r;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x));
#if CA_NORM_SHORT >= 2
    return r;
#else
    return fma(r, fma(-x, r, 1.0), r);
#endif
#else
    return rcp_d(x);
#endif
}
// odd series Σ_{k=0..9} c_k u^{2k+1}: atanh-type (c_k = 1/(2k+1)) and atan-type (alternating), for |u| < 0.12 (u^21/21 < 3e-21)
CA_HD double odd_series(double u, bool alternating) {
    const double z = alternating ? -(u * u) : u * u;
    double p = 1.0 / 19.0;
    p = fma(p, z, 1.0 / 17.0); p = fma(p, z, 1.0 / 15.0); p = fma(p, z, 1.0 / 13.0); p = fma(p, z, 1.0 / 11.0);
    p = fma(p, z, 1.0 / 9.0); p = fma(p, z, 1.0 / 7.0); p = fma(p, z, 1.0 / 5.0); p = fma(p, z, 1.0 / 3.0);
    return fma(u * z, p, u);
}
CA_HD double log1p_small(double x) { return 2.0 * odd_series(x * rcp_d(2.0 + x), false); }   // ln(1+x) = 2 atanh(x / (2 + x)),  |x| < 0.25
CA_HD double atan_small(double t) { return odd_series(t, true); }                             // |t| < 0.12
CA_HD double sqrt_d(double x);
CA_HD double sqrt_norm(double x) {
#if defined(__CUDA_ARCH__) && CA_NORM_SHORT
    if (!(x > 0.0)) return 0.0;
    double y;
    asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
#if CA_NORM_SHORT >= 2
    return x * y;
#else
    const double g = x * y, h = 0.5 * y;
    return fma(g, fma(-h, g, 0.5), g);
#endif
#else
    return sqrt_d(x);
#endif
}
CA_HD double sqrt_d(double x) {  // sqrt for x >= 0 (0 -> 0)
#if defined(__CUDA_ARCH__)
    if (!(x > 0.0)) return 0.0;
    double y;
    asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
    double g = x * y, h = 0.5 * y;
    double r = fma(-h, g, 0.5);
    g = fma(g, r, g); h = fma(h, r, h);
    r = fma(-h, g, 0.5);
    g = fma(g, r, g); h = fma(h, r, h);
    double d = fma(-g, g, x);
    return fma(d, h, g);
#else
    return sqrt(x);
#endif
}
CA_HD double exp_d(double x) {
    if (!(x > -700.0 && x < 700.0)) return exp(x);
    const double t = fma(x, MC.log2e, MC.magic);
    const double n = t - MC.magic;
    double r = fma(-n, MC.ln2_hi, x);
    r = fma(-n, MC.ln2_lo, r);
    // degree-13 Taylor polynomial in Estrin form (|r| <= ln2/2: truncation 4e-18)
    const double r2 = r * r, r4 = r2 * r2, r8 = r4 * r4;
    const double p01 = fma(MC.ex[1], r, MC.ex[0]), p23 = fma(MC.ex[3], r, MC.ex[2]), p45 = fma(MC.ex[5], r, MC.ex[4]),
                 p67 = fma(MC.ex[7], r, MC.ex[6]), p89 = fma(MC.ex[9], r, MC.ex[8]), pab = fma(MC.ex[11], r, MC.ex[10]),
                 pcd = fma(MC.ex[13], r, MC.ex[12]);
    const double q0 = fma(p23, r2, p01), q1 = fma(p67, r2, p45), q2 = fma(pab, r2, p89);
    const double s0 = fma(q1, r4, q0), s1 = fma(pcd, r4, q2);
    const double p = fma(s1, r8, s0);
#if defined(__CUDA_ARCH__)
    const int k = __double2loint(t);
    return p * __hiloint2double((k + 1023) << 20, 0);
#else
    return ldexp(p, (int)n);
#endif
}
// MUFU seed + ONE refinement (relative error ~1e-11): scale factors of the two-atom kernels' error norm
CA_HD double rcp_r1(double x) {
#if defined(__CUDA_ARCH__)
    double r;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x));
    return fma(r, fma(-x, r, 1.0), r);
#else
    return 1.0 / x;
#endif
}
CA_HD double sqrt_r1(double x) {
#if defined(__CUDA_ARCH__)
    if (!(x > 0.0)) return 0.0;
    double y;
    asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
    const double g = x * y, h = 0.5 * y;
    return fma(g, fma(-h, g, 0.5), g);
#else
    return sqrt(x);
#endif
}
// ln x for normal x > 0 (~1 ulp): x = m 2^e with m in [sqrt(1/2), sqrt(2)), ln m = 2 atanh((m - 1)/(m + 1)) by the odd series
// (|u| <= 0.1716: u^21/21 < 4e-18), e ln2 added in two pieces.  No special cases: callers pass finite positive arguments.
CA_HD double log_d(double x) {
    int e;
    double m;
#if defined(__CUDA_ARCH__)
    const int hi = __double2hiint(x);
    e = ((hi >> 20) & 0x7ff) - 1022;
    m = __hiloint2double((hi & 0x800fffff) | 0x3fe00000, __double2loint(x));   // m in [0.5, 1)
#else
    m = frexp(x, &e);
#endif
    if (m < 0.70710678118654752440) { m *= 2.0; e -= 1; }
    const double u = (m - 1.0) * rcp_d(m + 1.0);
    return fma((double)e, MC.ln2_hi, fma((double)e, MC.ln2_lo, 2.0 * odd_series(u, false)));
}
CA_HD void sincos_fast(double a, double* sn, double* cs) {
    if (!(fabs(a) < 1.0e6)) { sincos_d(a, sn, cs); return; }
    const double t = fma(a, MC.two_over_pi, MC.magic);
    const double n = t - MC.magic;
    double r = fma(-n, MC.pio2_hi, a);
    r = fma(-n, MC.pio2_mid, r);
    r = fma(-n, MC.pio2_lo, r);
    const double z = r * r;
    double ps = fma(MC.sn[5], z, MC.sn[4]);
    double pc = fma(MC.cs[5], z, MC.cs[4]);
    ps = fma(ps, z, MC.sn[3]); pc = fma(pc, z, MC.cs[3]);
    ps = fma(ps, z, MC.sn[2]); pc = fma(pc, z, MC.cs[2]);
    ps = fma(ps, z, MC.sn[1]); pc = fma(pc, z, MC.cs[1]);
    ps = fma(ps, z, MC.sn[0]); pc = fma(pc, z, MC.cs[0]);
    const double s = fma(ps * z, r, r);
    const double c = fma(pc * z, z, fma(-0.5, z, 1.0));
#if defined(__CUDA_ARCH__)
    const int q = __double2loint(t);
#else
    const int q = (int)(long long)n;
#endif
    const double ss = (q & 1) ? c : s, cc = (q & 1) ? s : c;
    *sn = (q & 2) ? -ss : ss;
    *cs = ((q + 1) & 2) ? -cc : cc;
}

// ---------------------------------------------------------------------------------------------------
// Derived (device) model: everything that is constant over the ensemble, precomputed on the host.
// ---------------------------------------------------------------------------------------------------
struct DevBeam {
    double hO;        // Ω0/2
    double cs, sn;    // cosθ, sinθ
    double w0, z0, inv_z0, inv_w0sq, k;  // k = 2 z0 / w0^2
    double sqz_inv;   // 1/sqz
    int kind, nx, ny, nl;                // HG: nx = n/2, ny = m/2 ; LG: nl = l/2
    double cx[kMaxModes], cy[kMaxModes]; // HG_coeff(i, n/2), HG_coeff(j, m/2) or LG_coeff(p, l/2) in cx
};

struct DevModel {
    DevBeam red, blue;
    double aDx, aDz, bDx, bDz;  // Δ_D = aDx Vx + aDz Vz ; δ_D = bDx Vx + bDz Vz
    double dsign;               // -1: H_pp = -Δ_D-Δ0 (1-atom) ; +1: CZ convention
    double Delta0, delta0;
    double G0, G1, Gl, Gr;      // gated decay rates
    double wr, wz;              // trap frequencies
    double sig[6];              // sampler standard deviations
    double U0, w0, z0, mfac;    // mfac = m / vconst^2
    double ecut;                // U0 (1 - 1e-2)
    double center[2][3];
    double c6, xi, tau;
    double dtn, inv_dtn;        // phase-noise node spacing
    int n_nodes;
    int atom_motion, free_motion, laser_noise;
    int vz_from_vy, doppler_z_legacy, rho2_mode;
    int rows;                   // bit0: ψ0 has a |0> component, bit1: |l> component
    int noise_coarse;           // phase-noise tables: interpolation plan (noise_c / noise_p): every c-th node is summed exactly (see gen_noise_trace_c)
    double rho0[25];            // Hermitian-packed initial state (see pack order below)
};

// Hermitian packing of a 5x5 matrix into 25 reals: diag (0,1,r,p,l) then upper off-diagonals
// (0,1),(0,r),(0,p),(0,l),(1,r),(1,p),(1,l),(r,p),(r,l),(p,l) as (re,im).
constexpr int kNV = 50;  // packed ρ (25) + packed second moment (25)

// ---------------------------------------------------------------------------------------------------
// Beams.  Returns Ω(x,y,z) e^{iφ} / 2 (the Hamiltonian coefficient) for noise phase φ.
// Gaussian: with s = z'/z0, q = 1+s², g = x'²/(w0² q):  A·A_phase = e^{-g(1+is)} (1+is)/q  exactly
// (w0/w = q^-½, e^{i atan s} = (1+is) q^-½, curvature phase = -s g): one exp + one sincos, no atan.
// ---------------------------------------------------------------------------------------------------
CA_HD void beam_half_omega(const DevBeam& b, double x, double y, double z, double phi, double* re, double* im) {
    if (b.kind == CA_BEAM_GAUSS) {
        double r2 = x * x + y * y;
        double xr2, zr;
        if (b.sn == 0.0 && b.cs == 1.0) { xr2 = r2; zr = z; }
        else { double rp = sqrt_d(r2); double xr = rp * b.cs - z * b.sn; zr = rp * b.sn + z * b.cs; xr2 = xr * xr; }
        double s = zr * b.inv_z0;
        double iq = rcp_d(1.0 + s * s);
        double g = xr2 * b.inv_w0sq * iq;
        double e = b.hO * iq * exp_d(-g);
        double sn, cs;
        sincos_fast(phi - s * g, &sn, &cs);
        *re = e * (cs - s * sn);
        *im = e * (sn + s * cs);
    } else if (b.kind == CA_BEAM_UNIFORM) {
        double sn, cs;
        sincos_fast(phi, &sn, &cs);
        *re = b.hO * cs; *im = b.hO * sn;
    } else {
        // flat-top HG / LG (arbitrary_beams.jl:50-81): E = Ω (w0/w) e^{-i phase0} Σ c HG HG e^{-i(i+j+1)at}
        double s = z * b.inv_z0;
        double q = 1.0 + s * s;
        double iq = rcp_d(q);                // (in-house rcp / sqrt / exp / sincos throughout: the library versions drag their slow paths
        double rq = sqrt_d(iq);              //  through the instruction cache at every anchor refresh)   w0/w
        double iw2 = b.inv_w0sq * iq;        // 1/w^2
        double r2 = x * x + y * y;
        double phase0 = b.k * z + s * r2 * iw2;  // k z + (k/2) z r²/(z²+zr²) = k z + s r²/w²
        // g1 = e^{-i atan s} = (1 - i s) rq ; g2 = g1²
        double g1r = rq, g1i = -s * rq;
        double Sr, Si;
        if (b.kind == CA_BEAM_FLATTOP_HG) {
            double g2r = g1r * g1r - g1i * g1i, g2i = 2.0 * g1r * g1i;
            double sums[2][2];
#pragma unroll
            for (int ax = 0; ax < 2; ++ax) {
                double u = ax == 0 ? x : y * b.sqz_inv;   // HG(y, w*sqz, j)
                int nmax = ax == 0 ? b.nx : b.ny;
                const double* cf = ax == 0 ? b.cx : b.cy;
                double xi = sqrt_d(2.0 * iw2) * u;        // sqrt(2) u / w
                double h0 = 1.0, h1 = 2.0 * xi;           // H_0, H_1
                double pr = 1.0, pi = 0.0;                // g2^a
                double ar = cf[0], ai = 0.0;
                for (int a = 1; a <= nmax; ++a) {
                    // advance Hermite two orders: H_{2a}
                    double h2 = 2.0 * xi * h1 - 2.0 * (2 * a - 1) * h0;       // H_{2a}
                    double h3 = 2.0 * xi * h2 - 2.0 * (2 * a) * h1;           // H_{2a+1}
                    h0 = h2; h1 = h3;
                    double t = pr * g2r - pi * g2i; pi = pr * g2i + pi * g2r; pr = t;
                    ar += cf[a] * h2 * pr; ai += cf[a] * h2 * pi;
                }
                double ex = exp_d(-u * u * iw2);
                sums[ax][0] = ar * ex; sums[ax][1] = ai * ex;
            }
            Sr = sums[0][0] * sums[1][0] - sums[0][1] * sums[1][1];
            Si = sums[0][0] * sums[1][1] + sums[0][1] * sums[1][0];
        } else {
            double u = 2.0 * r2 * iw2;
            double l0 = 1.0, l1 = 1.0 - u;
            double pr = 1.0, pi = 0.0;
            double ar = b.cx[0], ai = 0.0;
            for (int p = 1; p <= b.nl; ++p) {
                double lp = l1;  // L_p
                double t = pr * g1r - pi * g1i; pi = pr * g1i + pi * g1r; pr = t;
                ar += b.cx[p] * lp * pr; ai += b.cx[p] * lp * pi;
                double l2 = ((2.0 * p + 1.0 - u) * l1 - p * l0) * rcp_d(p + 1.0);
                l0 = l1; l1 = l2;
            }
            double ex = exp_d(-r2 * iw2);
            Sr = ar * ex; Si = ai * ex;
        }
        // E/2 = hO rq g1 S e^{i(phi - phase0)}
        double tr = g1r * Sr - g1i * Si, ti = g1r * Si + g1i * Sr;
        double sn, cs;
        sincos_fast(phi - phase0, &sn, &cs);
        double f = b.hO * rq;
        *re = f * (tr * cs - ti * sn);
        *im = f * (tr * sn + ti * cs);
    }
}

// Gaussian beams of NB evaluation points in ONE straight-line block (no branches, no library calls): the 2*NB
// exp / sincos dependency chains are independent, so the scheduler interleaves them and the FP64 pipe stays fed
// (with branches between them -- kind dispatch, range checks -- each chain ran alone and the pipe idled on latency).
// o[n] = {Re Ωr/2, Im Ωr/2, Re Ωb/2, Im Ωb/2}.  Phases are assumed |φ| < 1e8.
template <int NB>
CA_HD void gauss_batch(const DevBeam& br, const DevBeam& bb, const double* x, const double* y, const double* z, const double* phr,
                       const double* phb, double (*o)[4]) {
    double rp[NB];
#pragma unroll
    for (int n = 0; n < NB; ++n) rp[n] = sqrt_d(fmax(x[n] * x[n] + y[n] * y[n], 1e-300));
    // (rolling this beam loop to shrink the code was measured slower: 228 vs 209 ms; the 2*NB chains stay unrolled)
#pragma unroll
    for (int q = 0; q < 2; ++q) {
        const DevBeam* bp = q == 0 ? &br : &bb;
        const double bcs = bp->cs, bsn = bp->sn, biz0 = bp->inv_z0, biw = bp->inv_w0sq, bhO = bp->hO;
        double re[NB], im[NB];
#pragma unroll
        for (int n = 0; n < NB; ++n) {
            const double phi = q == 0 ? phr[n] : phb[n];
            const double xr = rp[n] * bcs - z[n] * bsn, zr = rp[n] * bsn + z[n] * bcs;
            const double s = zr * biz0;
            const double iq = rcp_d(fma(s, s, 1.0));
            const double g = xr * xr * biw * iq;
            // exp(-g), g >= 0 (clamped instead of branching)
            const double xx = fmax(-g, -700.0);
            const double t = fma(xx, MC.log2e, MC.magic);
            const double nn = t - MC.magic;
            double r = fma(-nn, MC.ln2_hi, xx);
            r = fma(-nn, MC.ln2_lo, r);
            const double q2 = r * r, q4 = q2 * q2, q8 = q4 * q4;
            const double p01 = fma(MC.ex[1], r, MC.ex[0]), p23 = fma(MC.ex[3], r, MC.ex[2]), p45 = fma(MC.ex[5], r, MC.ex[4]),
                         p67 = fma(MC.ex[7], r, MC.ex[6]), p89 = fma(MC.ex[9], r, MC.ex[8]), pab = fma(MC.ex[11], r, MC.ex[10]),
                         pcd = fma(MC.ex[13], r, MC.ex[12]);
            const double e0 = fma(fma(pcd, q4, fma(pab, q2, p89)), q8, fma(fma(p67, q2, p45), q4, fma(p23, q2, p01)));
#if defined(__CUDA_ARCH__)
            const double e = bhO * iq * e0 * __hiloint2double((__double2loint(t) + 1023) << 20, 0);
#else
            const double e = bhO * iq * ldexp(e0, (int)nn);
#endif
            // sincos(phi - s g)
            const double a = fma(-s, g, phi);
            const double t2 = fma(a, MC.two_over_pi, MC.magic);
            const double m2 = t2 - MC.magic;
            double w = fma(-m2, MC.pio2_hi, a);
            w = fma(-m2, MC.pio2_mid, w);
            w = fma(-m2, MC.pio2_lo, w);
            const double zz = w * w;
            double ps = fma(MC.sn[5], zz, MC.sn[4]), pc = fma(MC.cs[5], zz, MC.cs[4]);
            ps = fma(ps, zz, MC.sn[3]); pc = fma(pc, zz, MC.cs[3]);
            ps = fma(ps, zz, MC.sn[2]); pc = fma(pc, zz, MC.cs[2]);
            ps = fma(ps, zz, MC.sn[1]); pc = fma(pc, zz, MC.cs[1]);
            ps = fma(ps, zz, MC.sn[0]); pc = fma(pc, zz, MC.cs[0]);
            const double sv = fma(ps * zz, w, w);
            const double cv = fma(pc * zz, zz, fma(-0.5, zz, 1.0));
#if defined(__CUDA_ARCH__)
            const int qd = __double2loint(t2);
#else
            const int qd = (int)(long long)m2;
#endif
            const double ss = (qd & 1) ? cv : sv, cc = (qd & 1) ? sv : cv;
            const double sn = (qd & 2) ? -ss : ss, cs = ((qd + 1) & 2) ? -cc : cc;
            re[n] = e * (cs - s * sn);
            im[n] = e * (sn + s * cs);
        }
        if (q == 0) {
#pragma unroll
            for (int n = 0; n < NB; ++n) { o[n][0] = re[n]; o[n][1] = im[n]; }
        } else {
#pragma unroll
            for (int n = 0; n < NB; ++n) { o[n][2] = re[n]; o[n][3] = im[n]; }
        }
    }
}

// ---- anchored evaluation of a Gaussian beam (hot path of the adaptive integrator) -----------------------------------
// A DP5 step needs the coefficient c(t) = (Ω/2) B(r(t)) e^{iφ(t)} at five times inside [t, t+h] for both beams.  φ(t) is
// piecewise linear on the phase-noise nodes (rydberg_model.jl:166-167; the lane keeps a three-node window) and ln B along
// the ballistic trajectory is smooth on the scale w0/v >> h.  Every kAnchorEvery attempts ALL lanes of a warp re-anchor together
// (no divergence): one exact exp + sincos at the current time t_a and a quadratic model of ln B from its values at
// t_a, t_a + H, t_a + 2H (2H = (kAnchorEvery + 2) h; exact up to the third difference, ~1e-14 for µm-scale beams, guarded through |b2|).
// In between
//     c(t) = c(t_a) exp(δ),  δ = τ (b1 + τ b2) + i (φ(t) - φ(t_a)),  τ = t - t_a,
// costs three short real series per stage time (e^{Re δ}, cos Im δ, sin Im δ; |Re δ| < 3e-4, |Im δ| < 6e-3 guarded:
// truncation < 4e-16).  Nothing accumulates:
// every value hangs off an exactly evaluated anchor; a lane whose step leaves [t_a, t_a + 2H] re-anchors on the spot.
constexpr int kAnchorEvery = 32;   // attempts between warp-wide re-anchorings (power of two)
#ifndef CA_ANCHOR_EVERY_GENERIC
#define CA_ANCHOR_EVERY_GENERIC 64
#endif
constexpr int kAnchorEveryGeneric = CA_ANCHOR_EVERY_GENERIC;   // the same for the flat-top / uniform beam class (BEAMS = 2): no phase noise there, and the
                                                               // refresh is big code that evicts the step loop from the instruction cache (32: -16 %, 128: guard failures)
struct BeamGeom { double s, iq, g; };
CA_HD BeamGeom beam_geom(const DevBeam& b, double rp, double z) {
    const double xr = rp * b.cs - z * b.sn, zr = rp * b.sn + z * b.cs;
    BeamGeom o;
    o.s = zr * b.inv_z0;
    o.iq = rcp_d(fma(o.s, o.s, 1.0));
    o.g = xr * xr * b.inv_w0sq * o.iq;
    return o;
}
// exact: returns Ω e^{iφ}/2 (and {e^{-g}, g, cos a, sin a, a} with a = φ - s g)
CA_HD void gauss_exact(const DevBeam& b, double rp, double z, double phi, double* re, double* im, double* anc) {
    const BeamGeom q = beam_geom(b, rp, z);
    const double eg = exp_d(fmax(-q.g, -700.0));
    const double a = fma(-q.s, q.g, phi);
    double sn, cs;
    sincos_fast(a, &sn, &cs);
    const double e = b.hO * q.iq * eg;
    *re = e * (cs - q.s * sn);
    *im = e * (sn + q.s * cs);
    anc[0] = eg; anc[1] = q.g; anc[2] = cs; anc[3] = sn; anc[4] = a;
}
// log of the noise-free Gaussian beam factor up to the constant ln(Ω0/2):  ln[ iq e^{-g} (1 + i s) e^{-i s g} ]
CA_HD void gauss_log(const DevBeam& b, double rp, double z, double* lre, double* lim) {
    const BeamGeom q = beam_geom(b, rp, z);
    const double s2 = q.s * q.s;
    double l1p, at;
    if (s2 < 2.5e-3) {   // |s| = |z'|/z0 is tiny for every realistic beam: alternating series, truncation < 1e-19
        l1p = s2 * fma(-s2, fma(-s2, fma(-s2, fma(-s2, fma(-s2, 1.0 / 6.0, 1.0 / 5.0), 1.0 / 4.0), 1.0 / 3.0), 0.5), 1.0);
        at = q.s * fma(-s2, fma(-s2, fma(-s2, fma(-s2, fma(-s2, fma(-s2, 1.0 / 13.0, 1.0 / 11.0), 1.0 / 9.0), 1.0 / 7.0), 1.0 / 5.0), 1.0 / 3.0), 1.0);
    } else { l1p = log1p(s2); at = atan(q.s); }
    *lre = -0.5 * l1p - q.g;
    *lim = at - q.s * q.g;
}

// IEEE operations without FMA contraction: the sampler's rejection test and the recapture indicator are
// evaluated in the reference's operation order so that they round exactly like the oracle (-ffp-contract=off).
#if defined(__CUDA_ARCH__)
#define CA_MUL(a, b) __dmul_rn((a), (b))
#define CA_ADD(a, b) __dadd_rn((a), (b))
#define CA_DIV(a, b) __ddiv_rn((a), (b))
#else
#define CA_MUL(a, b) ((a) * (b))
#define CA_ADD(a, b) ((a) + (b))
#define CA_DIV(a, b) ((a) / (b))
#endif

// A(x,y,z,w0,z0) at θ=0, n=1: utilities.jl:27-31, literal operation order
CA_HD double trap_A_literal(double x, double y, double z, double w0, double z0) {
    double xt = sqrt(CA_ADD(CA_MUL(x, x), CA_MUL(y, y)));
    double zz = CA_DIV(z, z0);
    double w = CA_MUL(w0, sqrt(CA_ADD(1.0, CA_MUL(zz, zz))));
    return CA_MUL(CA_DIV(w0, w), exp(-CA_DIV(CA_MUL(xt, xt), CA_MUL(w, w))));
}
// K = m/vconst^2 * (vx^2 + vy^2 + vz^2) / 2 : atom_sampler.jl:35-39 (mfac = m/vconst^2)
CA_HD double kinetic_literal(double mfac, double vx, double vy, double vz) {
    double s = CA_ADD(CA_ADD(CA_MUL(vx, vx), CA_MUL(vy, vy)), CA_MUL(vz, vz));
    return CA_DIV(CA_MUL(mfac, s), 2.0);
}

struct SamplerConsts { double sig[6]; double U0, w0, z0, mfac, ecut; };

// One atom sample (atom_sampler.jl:84-96): attempts a = 0,1,..; 3 Philox calls per attempt, Box-Muller pairs
// (x,y),(z,vx),(vy,vz); accept iff K + U0(1 - A²) < ecut.  Returns the number of attempts (or -1).
CA_HD int sample_atom_literal(const SamplerConsts& sc, uint64_t seed, uint64_t index, uint32_t stream, double* out) {
    for (uint32_t a = 0; a < 100000u; ++a) {
#pragma unroll
        for (uint32_t c = 0; c < 3; ++c) {
            U4 w = philox(seed, index, a * 3u + c, stream);
            double u1 = u53_open(w.x, w.y), u2 = u53(w.z, w.w);
            double r = sqrt(CA_MUL(-2.0, log(u1)));
            double sn, cs;
            sincos_d(CA_MUL(kTwoPi, u2), &sn, &cs);
            out[2 * c] = CA_MUL(sc.sig[2 * c], CA_MUL(r, cs));
            out[2 * c + 1] = CA_MUL(sc.sig[2 * c + 1], CA_MUL(r, sn));
        }
        double A = trap_A_literal(out[0], out[1], out[2], sc.w0, sc.z0);
        double en = CA_ADD(CA_MUL(sc.U0, CA_ADD(1.0, -CA_MUL(A, A))), kinetic_literal(sc.mfac, out[3], out[4], out[5]));
        if (en < sc.ecut) return (int)a + 1;
    }
    return -1;
}

// Thermal averages of scalar functionals of one sampled atom / atom pair.
//   kind 0 (calibrate_two_photon, rydberg_model.jl:285-306): with Δ = the red beam's Doppler shift Δ(vx, vz, red) ALONE -- the
//     reference passes it as the detuning of Ω_twophoton / δ_twophoton (:293, :300) --, |Ωr|, |Ωb| at the sampled position:
//     o[0] = |Ωb Ωr / (2Δ)| (:267-269), o[1] = (|Ωr|² − |Ωb|²)/(4Δ) (:275-277).
//   kind 1 (get_blockade_stark_shift_factor, cz_model.jl:77-104): o[0] = 1/(1e-18 + |r1 + c1 − r2 − c2|²·³) (:101).
struct AvgConsts { DevBeam red, blue; double aDx, aDz; double center[2][3]; int kind; };
CA_HD void thermal_terms(const AvgConsts& ac, const double* s1, const double* s2, double* o) {
    if (ac.kind == 0) {
        double rr, ri, br, bi;
        beam_half_omega(ac.red, s1[0], s1[1], s1[2], 0.0, &rr, &ri);
        beam_half_omega(ac.blue, s1[0], s1[1], s1[2], 0.0, &br, &bi);
        const double r2 = 4.0 * (rr * rr + ri * ri), b2 = 4.0 * (br * br + bi * bi);   // |Ωr|², |Ωb|²
        const double D = ac.aDx * s1[3] + ac.aDz * s1[5];
        o[0] = sqrt(r2 * b2) / fabs(2.0 * D);
        o[1] = (r2 - b2) / (4.0 * D);
    } else {
        const double dx = (s1[0] + ac.center[0][0]) - (s2[0] + ac.center[1][0]);
        const double dy = (s1[1] + ac.center[0][1]) - (s2[1] + ac.center[1][1]);
        const double dz = (s1[2] + ac.center[0][2]) - (s2[2] + ac.center[1][2]);
        const double q = dx * dx + dy * dy + dz * dz;
        o[0] = 1.0 / (1e-18 + q * q * q);
        o[1] = 0.0;
    }
}

// Metropolis branch of samples_generate (atom_sampler.jl:97-120) as MANY independent chains: chain c starts at the
// reference's initial state, burns `skip` transitions in and then emits every freq-th state; its k-th sample is global
// sample c + k*n_chains.  One chain (n_chains = 1) is exactly the reference algorithm.  RNG: Philox(seed; index =
// chain<<32 | transition, draw 0..2 -> Box-Muller proposal pairs, draw 3 -> acceptance uniform; stream).
struct MetroConsts { double sd[6]; double U0, w0, z0, mfac, T, ecut; int harmonic; };
CA_HD double metro_energy(const MetroConsts& mc, const double* c) {
    double pot;
    if (mc.harmonic) {   // Π_Harmonic: atom_sampler.jl:24-30
        const double r2 = CA_ADD(CA_MUL(c[0], c[0]), CA_MUL(c[1], c[1])), zz = CA_DIV(c[2], mc.z0);
        pot = CA_MUL(mc.U0, CA_ADD(CA_DIV(CA_MUL(2.0, r2), CA_MUL(mc.w0, mc.w0)), CA_MUL(zz, zz)));
    } else {
        const double A = trap_A_literal(c[0], c[1], c[2], mc.w0, mc.z0);
        pot = CA_MUL(mc.U0, CA_ADD(1.0, -CA_MUL(A, A)));
    }
    return CA_ADD(pot, kinetic_literal(mc.mfac, c[3], c[4], c[5]));
}
// Runs chain `chain`; writes its samples to out[(chain + k*n_chains)*6 ..]; returns accepted transitions, *steps = transitions made.
CA_HD long long metropolis_chain(const MetroConsts& mc, uint64_t seed, uint32_t stream, long long chain, long long n_chains, long long N,
                                 long long freq, long long skip, double* out, long long* steps) {
    const long long M = chain < N ? (N - chain + n_chains - 1) / n_chains : 0;   // samples of this chain
    *steps = 0;
    if (M == 0) return 0;
    const double v0 = CA_DIV(mc.sd[3], sqrt(3.0));
    double cur[6] = {0.0, 0.0, 0.0, v0, v0, v0};   // atom_sampler.jl:104
    double Ec = metro_energy(mc, cur);
    const long long L = skip + (M - 1) * freq + 1;   // states visited up to the last emitted one
    long long acc = 0, k = 0;
    for (long long i = 0;; ++i) {
        if (i >= skip && (i - skip) % freq == 0) {
            double* o = out + (size_t)(chain + k * n_chains) * 6;
#pragma unroll
            for (int q = 0; q < 6; ++q) o[q] = cur[q];
            if (++k == M) break;
        }
        if (i + 1 >= L) break;
        const uint64_t idx = ((uint64_t)chain << 32) | (uint64_t)(i + 1);
        double nw[6];
#pragma unroll
        for (uint32_t c = 0; c < 3; ++c) {
            U4 w = philox(seed, idx, c, stream);
            const double u1 = u53_open(w.x, w.y), u2 = u53(w.z, w.w);
            const double r = sqrt(CA_MUL(-2.0, log(u1)));
            double sn, cs;
            sincos_d(CA_MUL(kTwoPi, u2), &sn, &cs);
            nw[2 * c] = CA_ADD(cur[2 * c], CA_MUL(mc.sd[2 * c], CA_MUL(r, cs)));
            nw[2 * c + 1] = CA_ADD(cur[2 * c + 1], CA_MUL(mc.sd[2 * c + 1], CA_MUL(r, sn)));
        }
        U4 w = philox(seed, idx, 3u, stream);
        const double u = u53(w.x, w.y);
        const double En = metro_energy(mc, nw);
        const double p_acc = CA_DIV(exp(-CA_DIV(En, mc.T)), exp(-CA_DIV(Ec, mc.T)));   // prob_boltzmann ratio, :108
        if (p_acc > u && En < mc.ecut) {
#pragma unroll
            for (int q = 0; q < 6; ++q) cur[q] = nw[q];
            Ec = En; ++acc;
        }
        ++*steps;
    }
    return acc;
}

// Phase-noise trace (lasernoise_sampler.jl:31-37 on the uniform node grid of rydberg_model.jl:216):
// out[j*stride] = Σ_k a_k cos(2π f_k j dtn + φ_k), j = 0..n_nodes-1, φ_k ~ U[0,2π) from Philox (two per call) or
// host-supplied.  Nodes are produced CH at a time with the Reinsch-stabilised cosine recurrence
// (c += d; d += σ c; σ = -4 sin²(δ/2)): 3 flop per (node, frequency) plus one sincos per (chunk, frequency).
// Per-frequency constants of the node recurrence, computed once per call on the host (NoiseK[nf]).
struct NoiseK { double om, amp, sh, ch, sig; };  // ω = 2π f, a, sin(ω dtn/2), cos(ω dtn/2), -4 sin²(ω dtn/2)

template <int CH>
CA_HD void gen_noise_trace(const NoiseK* __restrict__ nk, int nf, int n_nodes, double dtn, double t_first, const double* __restrict__ phases,
                           uint64_t seed, uint64_t gi, uint32_t stream, double* __restrict__ out, int stride) {
    constexpr int KB = 8;  // frequencies advanced together: 8 independent c/d chains hide the FP64 latency
    for (int c0 = 0; c0 < n_nodes; c0 += CH) {
        double acc[CH];
#pragma unroll
        for (int j = 0; j < CH; ++j) acc[j] = 0.0;
        const double t0 = fma((double)c0, dtn, t_first);
        for (int k0 = 0; k0 < nf; k0 += KB) {
            double c[KB], d[KB], sig[KB];
#pragma unroll
            for (int q = 0; q < KB; q += 2) {
                // phases of frequencies k0+q, k0+q+1 come from one Philox call (k0 and KB are even)
                double ph[2];
                if (phases) { ph[0] = k0 + q < nf ? phases[k0 + q] : 0.0; ph[1] = k0 + q + 1 < nf ? phases[k0 + q + 1] : 0.0; }
                else { U4 w = philox(seed, gi, (uint32_t)((k0 + q) >> 1), stream); ph[0] = kTwoPi * u53(w.x, w.y); ph[1] = kTwoPi * u53(w.z, w.w); }
#pragma unroll
                for (int e = 0; e < 2; ++e) {
                    const int k = k0 + q + e < nf ? k0 + q + e : nf - 1;
                    const NoiseK kk = nk[k];
                    const double a = k0 + q + e < nf ? kk.amp : 0.0;
                    double s0, c0v;
                    sincos_fast(fma(kk.om, t0, ph[e]), &s0, &c0v);
                    sig[q + e] = kk.sig;
                    c[q + e] = a * c0v;
                    d[q + e] = -2.0 * a * (s0 * kk.ch + c0v * kk.sh) * kk.sh;  // c_1 - c_0 = -2 a sin(θ0 + δ/2) sin(δ/2)
                }
            }
            acc[0] += ((c[0] + c[1]) + (c[2] + c[3])) + ((c[4] + c[5]) + (c[6] + c[7]));
#pragma unroll
            for (int j = 1; j < CH; ++j) {
                if ((j & 7) == 0 && c0 + j >= n_nodes) break;   // (uniform) the last chunk stops at the end of the table, to a multiple of 8 nodes
#pragma unroll
                for (int q = 0; q < KB; ++q) { c[q] += d[q]; d[q] = fma(sig[q], c[q], d[q]); }
                acc[j] += ((c[0] + c[1]) + (c[2] + c[3])) + ((c[4] + c[5]) + (c[6] + c[7]));
            }
        }
#pragma unroll
        for (int j = 0; j < CH; ++j)
            if (c0 + j < n_nodes) out[(size_t)(c0 + j) * stride] = acc[j];
    }
}

// host helper: fills NoiseK[nf]
inline void fill_noise_k(const double* f, const double* amp, int nf, double dtn, NoiseK* out) {
    for (int k = 0; k < nf; ++k) {
        const double om = kTwoPi * f[k];
        out[k].om = om; out[k].amp = amp[k];
        out[k].sh = sin(0.5 * om * dtn); out[k].ch = cos(0.5 * om * dtn);
        out[k].sig = -4.0 * out[k].sh * out[k].sh;
    }
}

// ---- band-limited shortcut -----------------------------------------------------------------------------------------
// The trace is a sum of cosines with f <= f_max, and the node grid (rydberg_model.jl:216) oversamples it hundreds of
// times.  Only every c-th node (c = 32, 16, 8, 4 or 2) is summed exactly; the nodes in between follow from 8-point (degree 7)
// Lagrange interpolation of the exact ones, whose error Σ_k a_k (ω_k c dtn)^8 · 43.07/8! is kept below 1e-14 rad by the
// choice of c (observed: 6e-16 for the R1 workload at c = 4) -- the table equals the directly summed one to rounding
// while the O(nodes x frequencies) work drops by c.
struct LagrangeW { double w[57][8]; };   // rows: c=2 (r=1), c=4 (r=1..3), c=8, c=16, c=32 (r=1..c-1); columns: nodes at offsets -3..4
#define CA_LAGRANGE_INIT {{ \
    {-5.0 / 2048.0, 49.0 / 2048.0, -245.0 / 2048.0, 1225.0 / 2048.0, 1225.0 / 2048.0, -245.0 / 2048.0, 49.0 / 2048.0, -5.0 / 2048.0}, \
    {-495.0 / 262144.0, 5005.0 / 262144.0, -27027.0 / 262144.0, 225225.0 / 262144.0, 75075.0 / 262144.0, -19305.0 / 262144.0, 4095.0 / 262144.0, -429.0 / 262144.0}, \
    {-5.0 / 2048.0, 49.0 / 2048.0, -245.0 / 2048.0, 1225.0 / 2048.0, 1225.0 / 2048.0, -245.0 / 2048.0, 49.0 / 2048.0, -5.0 / 2048.0}, \
    {-429.0 / 262144.0, 4095.0 / 262144.0, -19305.0 / 262144.0, 75075.0 / 262144.0, 225225.0 / 262144.0, -27027.0 / 262144.0, 5005.0 / 262144.0, -495.0 / 262144.0}, \
    {-36363.0 / 33554432.0, 374325.0 / 33554432.0, -2121175.0 / 33554432.0, 31817625.0 / 33554432.0, 4545375.0 / 33554432.0, -1272705.0 / 33554432.0, 276675.0 / 33554432.0, -29325.0 / 33554432.0}, \
    {-495.0 / 262144.0, 5005.0 / 262144.0, -27027.0 / 262144.0, 225225.0 / 262144.0, 75075.0 / 262144.0, -19305.0 / 262144.0, 4095.0 / 262144.0, -429.0 / 262144.0}, \
    {-78793.0 / 33554432.0, 783783.0 / 33554432.0, -4061421.0 / 33554432.0, 24819795.0 / 33554432.0, 14891877.0 / 33554432.0, -3436587.0 / 33554432.0, 709137.0 / 33554432.0, -73359.0 / 33554432.0}, \
    {-5.0 / 2048.0, 49.0 / 2048.0, -245.0 / 2048.0, 1225.0 / 2048.0, 1225.0 / 2048.0, -245.0 / 2048.0, 49.0 / 2048.0, -5.0 / 2048.0}, \
    {-73359.0 / 33554432.0, 709137.0 / 33554432.0, -3436587.0 / 33554432.0, 14891877.0 / 33554432.0, 24819795.0 / 33554432.0, -4061421.0 / 33554432.0, 783783.0 / 33554432.0, -78793.0 / 33554432.0}, \
    {-429.0 / 262144.0, 4095.0 / 262144.0, -19305.0 / 262144.0, 75075.0 / 262144.0, 225225.0 / 262144.0, -27027.0 / 262144.0, 5005.0 / 262144.0, -495.0 / 262144.0}, \
    {-29325.0 / 33554432.0, 276675.0 / 33554432.0, -1272705.0 / 33554432.0, 4545375.0 / 33554432.0, 31817625.0 / 33554432.0, -2121175.0 / 33554432.0, 374325.0 / 33554432.0, -36363.0 / 33554432.0}, \
    {-2452131.0 / 4294967296.0, 25487301.0 / 4294967296.0, -148426047.0 / 4294967296.0, 4205404665.0 / 4294967296.0, 280360311.0 / 4294967296.0, -81394929.0 / 4294967296.0, 17895339.0 / 4294967296.0, -1907213.0 / 4294967296.0}, \
    {-36363.0 / 33554432.0, 374325.0 / 33554432.0, -2121175.0 / 33554432.0, 31817625.0 / 33554432.0, 4545375.0 / 33554432.0, -1272705.0 / 33554432.0, 276675.0 / 33554432.0, -29325.0 / 33554432.0}, \
    {-6554145.0 / 4294967296.0, 66852279.0 / 4294967296.0, -369446805.0 / 4294967296.0, 3899716275.0 / 4294967296.0, 899934525.0 / 4294967296.0, -242051355.0 / 4294967296.0, 51996217.0 / 4294967296.0, -5479695.0 / 4294967296.0}, \
    {-495.0 / 262144.0, 5005.0 / 262144.0, -27027.0 / 262144.0, 225225.0 / 262144.0, 75075.0 / 262144.0, -19305.0 / 262144.0, 4095.0 / 262144.0, -429.0 / 262144.0}, \
    {-9293031.0 / 4294967296.0, 93181473.0 / 4294967296.0, -492530643.0 / 4294967296.0, 3447714501.0 / 4294967296.0, 1567142955.0 / 4294967296.0, -383079389.0 / 4294967296.0, 80179407.0 / 4294967296.0, -8347977.0 / 4294967296.0}, \
    {-78793.0 / 33554432.0, 783783.0 / 33554432.0, -4061421.0 / 33554432.0, 24819795.0 / 33554432.0, 14891877.0 / 33554432.0, -3436587.0 / 33554432.0, 709137.0 / 33554432.0, -73359.0 / 33554432.0}, \
    {-10481445.0 / 4294967296.0, 103470675.0 / 4294967296.0, -526350825.0 / 4294967296.0, 2882397375.0 / 4294967296.0, 2241864625.0 / 4294967296.0, -484242759.0 / 4294967296.0, 98423325.0 / 4294967296.0, -10113675.0 / 4294967296.0}, \
    {-5.0 / 2048.0, 49.0 / 2048.0, -245.0 / 2048.0, 1225.0 / 2048.0, 1225.0 / 2048.0, -245.0 / 2048.0, 49.0 / 2048.0, -5.0 / 2048.0}, \
    {-10113675.0 / 4294967296.0, 98423325.0 / 4294967296.0, -484242759.0 / 4294967296.0, 2241864625.0 / 4294967296.0, 2882397375.0 / 4294967296.0, -526350825.0 / 4294967296.0, 103470675.0 / 4294967296.0, -10481445.0 / 4294967296.0}, \
    {-73359.0 / 33554432.0, 709137.0 / 33554432.0, -3436587.0 / 33554432.0, 14891877.0 / 33554432.0, 24819795.0 / 33554432.0, -4061421.0 / 33554432.0, 783783.0 / 33554432.0, -78793.0 / 33554432.0}, \
    {-8347977.0 / 4294967296.0, 80179407.0 / 4294967296.0, -383079389.0 / 4294967296.0, 1567142955.0 / 4294967296.0, 3447714501.0 / 4294967296.0, -492530643.0 / 4294967296.0, 93181473.0 / 4294967296.0, -9293031.0 / 4294967296.0}, \
    {-429.0 / 262144.0, 4095.0 / 262144.0, -19305.0 / 262144.0, 75075.0 / 262144.0, 225225.0 / 262144.0, -27027.0 / 262144.0, 5005.0 / 262144.0, -495.0 / 262144.0}, \
    {-5479695.0 / 4294967296.0, 51996217.0 / 4294967296.0, -242051355.0 / 4294967296.0, 899934525.0 / 4294967296.0, 3899716275.0 / 4294967296.0, -369446805.0 / 4294967296.0, 66852279.0 / 4294967296.0, -6554145.0 / 4294967296.0}, \
    {-29325.0 / 33554432.0, 276675.0 / 33554432.0, -1272705.0 / 33554432.0, 4545375.0 / 33554432.0, 31817625.0 / 33554432.0, -2121175.0 / 33554432.0, 374325.0 / 33554432.0, -36363.0 / 33554432.0}, \
    {-1907213.0 / 4294967296.0, 17895339.0 / 4294967296.0, -81394929.0 / 4294967296.0, 280360311.0 / 4294967296.0, 4205404665.0 / 4294967296.0, -148426047.0 / 4294967296.0, 25487301.0 / 4294967296.0, -2452131.0 / 4294967296.0}, \
    {-160452435.0 / 549755813888.0, 1676110821.0 / 549755813888.0, -9904291215.0 / 549755813888.0, 544736016825.0 / 549755813888.0, 17572129575.0 / 549755813888.0, -5187962065.0 / 549755813888.0, 1146812667.0 / 549755813888.0, -122550285.0 / 549755813888.0}, \
    {-2452131.0 / 4294967296.0, 25487301.0 / 4294967296.0, -148426047.0 / 4294967296.0, 4205404665.0 / 4294967296.0, 280360311.0 / 4294967296.0, -81394929.0 / 4294967296.0, 17895339.0 / 4294967296.0, -1907213.0 / 4294967296.0}, \
    {-459276625.0 / 549755813888.0, 4750428375.0 / 549755813888.0, -27281031525.0 / 549755813888.0, 530464501875.0 / 549755813888.0, 54875638125.0 / 549755813888.0, -15653050875.0 / 549755813888.0, 3422351625.0 / 549755813888.0, -363747087.0 / 549755813888.0}, \
    {-36363.0 / 33554432.0, 374325.0 / 33554432.0, -2121175.0 / 33554432.0, 31817625.0 / 33554432.0, 4545375.0 / 33554432.0, -1272705.0 / 33554432.0, 276675.0 / 33554432.0, -29325.0 / 33554432.0}, \
    {-722557719.0 / 549755813888.0, 7403598657.0 / 549755813888.0, -41420133027.0 / 549755813888.0, 510848307333.0 / 549755813888.0, 94601538395.0 / 549755813888.0, -25975337661.0 / 549755813888.0, 5613717663.0 / 549755813888.0, -593319753.0 / 549755813888.0}, \
    {-6554145.0 / 4294967296.0, 66852279.0 / 4294967296.0, -369446805.0 / 4294967296.0, 3899716275.0 / 4294967296.0, 899934525.0 / 4294967296.0, -242051355.0 / 4294967296.0, 51996217.0 / 4294967296.0, -5479695.0 / 4294967296.0}, \
    {-944279765.0 / 549755813888.0, 9589094515.0 / 549755813888.0, -52371208505.0 / 549755813888.0, 486304078975.0 / 549755813888.0, 136165142113.0 / 549755813888.0, -35832932135.0 / 549755813888.0, 7649727085.0 / 549755813888.0, -803808395.0 / 549755813888.0}, \
    {-495.0 / 262144.0, 5005.0 / 262144.0, -27027.0 / 262144.0, 225225.0 / 262144.0, 75075.0 / 262144.0, -19305.0 / 262144.0, 4095.0 / 262144.0, -429.0 / 262144.0}, \
    {-1119941691.0 / 549755813888.0, 11276125245.0 / 549755813888.0, -60231010455.0 / 549755813888.0, 457309523825.0 / 549755813888.0, 178947204975.0 / 549755813888.0, -44899480521.0 / 549755813888.0, 9461576355.0 / 549755813888.0, -988183845.0 / 549755813888.0}, \
    {-9293031.0 / 4294967296.0, 93181473.0 / 4294967296.0, -492530643.0 / 4294967296.0, 3447714501.0 / 4294967296.0, 1567142955.0 / 4294967296.0, -383079389.0 / 4294967296.0, 80179407.0 / 4294967296.0, -8347977.0 / 4294967296.0}, \
    {-1246556025.0 / 549755813888.0, 12448939503.0 / 549755813888.0, -65139799725.0 / 549755813888.0, 424395664875.0 / 549755813888.0, 222302491125.0 / 549755813888.0, -52849271475.0 / 549755813888.0, 10984358385.0 / 549755813888.0, -1140012775.0 / 549755813888.0}, \
    {-78793.0 / 33554432.0, 783783.0 / 33554432.0, -4061421.0 / 33554432.0, 24819795.0 / 33554432.0, 14891877.0 / 33554432.0, -3436587.0 / 33554432.0, 709137.0 / 33554432.0, -73359.0 / 33554432.0}, \
    {-1322622015.0 / 549755813888.0, 13105981785.0 / 549755813888.0, -67277373163.0 / 549755813888.0, 388138691325.0 / 549755813888.0, 265568578275.0 / 549755813888.0, -59362388085.0 / 549755813888.0, 12158561415.0 / 549755813888.0, -1253615649.0 / 549755813888.0}, \
    {-10481445.0 / 4294967296.0, 103470675.0 / 4294967296.0, -526350825.0 / 4294967296.0, 2882397375.0 / 4294967296.0, 2241864625.0 / 4294967296.0, -484242759.0 / 4294967296.0, 98423325.0 / 4294967296.0, -10113675.0 / 4294967296.0}, \
    {-1348075197.0 / 549755813888.0, 13258916811.0 / 549755813888.0, -66858793281.0 / 549755813888.0, 349151476023.0 / 549755813888.0, 308074831785.0 / 549755813888.0, -64129862943.0 / 549755813888.0, 12931536149.0 / 549755813888.0, -1324215459.0 / 549755813888.0}, \
    {-5.0 / 2048.0, 49.0 / 2048.0, -245.0 / 2048.0, 1225.0 / 2048.0, 1225.0 / 2048.0, -245.0 / 2048.0, 49.0 / 2048.0, -5.0 / 2048.0}, \
    {-1324215459.0 / 549755813888.0, 12931536149.0 / 549755813888.0, -64129862943.0 / 549755813888.0, 308074831785.0 / 549755813888.0, 349151476023.0 / 549755813888.0, -66858793281.0 / 549755813888.0, 13258916811.0 / 549755813888.0, -1348075197.0 / 549755813888.0}, \
    {-10113675.0 / 4294967296.0, 98423325.0 / 4294967296.0, -484242759.0 / 4294967296.0, 2241864625.0 / 4294967296.0, 2882397375.0 / 4294967296.0, -526350825.0 / 4294967296.0, 103470675.0 / 4294967296.0, -10481445.0 / 4294967296.0}, \
    {-1253615649.0 / 549755813888.0, 12158561415.0 / 549755813888.0, -59362388085.0 / 549755813888.0, 265568578275.0 / 549755813888.0, 388138691325.0 / 549755813888.0, -67277373163.0 / 549755813888.0, 13105981785.0 / 549755813888.0, -1322622015.0 / 549755813888.0}, \
    {-73359.0 / 33554432.0, 709137.0 / 33554432.0, -3436587.0 / 33554432.0, 14891877.0 / 33554432.0, 24819795.0 / 33554432.0, -4061421.0 / 33554432.0, 783783.0 / 33554432.0, -78793.0 / 33554432.0}, \
    {-1140012775.0 / 549755813888.0, 10984358385.0 / 549755813888.0, -52849271475.0 / 549755813888.0, 222302491125.0 / 549755813888.0, 424395664875.0 / 549755813888.0, -65139799725.0 / 549755813888.0, 12448939503.0 / 549755813888.0, -1246556025.0 / 549755813888.0}, \
    {-8347977.0 / 4294967296.0, 80179407.0 / 4294967296.0, -383079389.0 / 4294967296.0, 1567142955.0 / 4294967296.0, 3447714501.0 / 4294967296.0, -492530643.0 / 4294967296.0, 93181473.0 / 4294967296.0, -9293031.0 / 4294967296.0}, \
    {-988183845.0 / 549755813888.0, 9461576355.0 / 549755813888.0, -44899480521.0 / 549755813888.0, 178947204975.0 / 549755813888.0, 457309523825.0 / 549755813888.0, -60231010455.0 / 549755813888.0, 11276125245.0 / 549755813888.0, -1119941691.0 / 549755813888.0}, \
    {-429.0 / 262144.0, 4095.0 / 262144.0, -19305.0 / 262144.0, 75075.0 / 262144.0, 225225.0 / 262144.0, -27027.0 / 262144.0, 5005.0 / 262144.0, -495.0 / 262144.0}, \
    {-803808395.0 / 549755813888.0, 7649727085.0 / 549755813888.0, -35832932135.0 / 549755813888.0, 136165142113.0 / 549755813888.0, 486304078975.0 / 549755813888.0, -52371208505.0 / 549755813888.0, 9589094515.0 / 549755813888.0, -944279765.0 / 549755813888.0}, \
    {-5479695.0 / 4294967296.0, 51996217.0 / 4294967296.0, -242051355.0 / 4294967296.0, 899934525.0 / 4294967296.0, 3899716275.0 / 4294967296.0, -369446805.0 / 4294967296.0, 66852279.0 / 4294967296.0, -6554145.0 / 4294967296.0}, \
    {-593319753.0 / 549755813888.0, 5613717663.0 / 549755813888.0, -25975337661.0 / 549755813888.0, 94601538395.0 / 549755813888.0, 510848307333.0 / 549755813888.0, -41420133027.0 / 549755813888.0, 7403598657.0 / 549755813888.0, -722557719.0 / 549755813888.0}, \
    {-29325.0 / 33554432.0, 276675.0 / 33554432.0, -1272705.0 / 33554432.0, 4545375.0 / 33554432.0, 31817625.0 / 33554432.0, -2121175.0 / 33554432.0, 374325.0 / 33554432.0, -36363.0 / 33554432.0}, \
    {-363747087.0 / 549755813888.0, 3422351625.0 / 549755813888.0, -15653050875.0 / 549755813888.0, 54875638125.0 / 549755813888.0, 530464501875.0 / 549755813888.0, -27281031525.0 / 549755813888.0, 4750428375.0 / 549755813888.0, -459276625.0 / 549755813888.0}, \
    {-1907213.0 / 4294967296.0, 17895339.0 / 4294967296.0, -81394929.0 / 4294967296.0, 280360311.0 / 4294967296.0, 4205404665.0 / 4294967296.0, -148426047.0 / 4294967296.0, 25487301.0 / 4294967296.0, -2452131.0 / 4294967296.0}, \
    {-122550285.0 / 549755813888.0, 1146812667.0 / 549755813888.0, -5187962065.0 / 549755813888.0, 17572129575.0 / 549755813888.0, 544736016825.0 / 549755813888.0, -9904291215.0 / 549755813888.0, 1676110821.0 / 549755813888.0, -160452435.0 / 549755813888.0} }}
#if defined(__CUDACC__)
static __constant__ LagrangeW c_lw = CA_LAGRANGE_INIT;
#endif
static const LagrangeW h_lw = CA_LAGRANGE_INIT;
#if defined(__CUDA_ARCH__)
#define LW c_lw
#else
#define LW h_lw
#endif

// 16-point (degree 15) weights: the same construction with twice the window lets the coarse grid be 4x sparser at the same error bound
// (R1: every 16th node instead of every 4th, 79 summed nodes per trace instead of 258).
#include "ca_lagrange16.h"
struct LagrangeW16 { double w[53][16]; };
#if defined(__CUDACC__)
static __constant__ LagrangeW16 c_lw16 = CA_LAGRANGE16_INIT;
#endif
static const LagrangeW16 h_lw16 = CA_LAGRANGE16_INIT;
#if defined(__CUDA_ARCH__)
#define LW16 c_lw16
#else
#define LW16 h_lw16
#endif

// Interpolation plan, encoded in one int (DevModel::noise_coarse): low byte = coarsening factor c (1: every node is summed), bit 8 set =
// 16-point window (else 8-point).
CA_HD int noise_c(int plan) { return plan & 0xff; }
CA_HD int noise_p(int plan) { return (plan & 0x100) ? 16 : 8; }
CA_HD int noise_coarse_count(int n_nodes, int plan) {   // coarse nodes m = -(p/2 - 1) .. ceil((n-1)/c) + p/2
    const int c = noise_c(plan);
    return (n_nodes - 1 + c - 1) / c + noise_p(plan);
}
#ifndef CA_NOISE_CHUNK
#define CA_NOISE_CHUNK 64
#endif
constexpr int kNoiseChunk = CA_NOISE_CHUNK;   // nodes summed per pass over the frequencies (accumulators in registers)

// host: the cheapest plan whose interpolation error bound stays below 1e-14 rad (1 = sum every node).  p-point central Lagrange
// interpolation of cos(ω t + φ) on nodes of spacing H errs by at most (ω H)^p / p! · max_s Π|s - j| = (ω H)^p · K_p / p!, with
// K_8 = (0.5·1.5·2.5·3.5)² = 43.07 and K_16 = (0.5·1.5· … ·7.5)² = 6.2696e7.  Cost model: one sincos set-up per (pass, frequency) ~45
// flop-equivalents, 3 per (summed node, frequency), p per interpolated node.
inline int choose_noise_coarse(const double* f, const double* amp_a, const double* amp_b, int nf, double dtn, int n_nodes) {
    if (n_nodes < 64) return 1;
    int best = 1;
    double best_cost = 1e300;
    for (int p = 8; p <= 16; p += 8) {
        for (int c = 32; c >= (p == 16 ? 8 : 2); c >>= 1) {
            const int plan = c | (p == 16 ? 0x100 : 0), n_c = noise_coarse_count(n_nodes, plan);
            if (p == 8 && n_c * c > 2 * n_nodes) continue;   // the padding nodes must not dominate
            double ba = 0.0, bb = 0.0;
            for (int k = 0; k < nf; ++k) {
                const double x = kTwoPi * fabs(f[k]) * c * dtn, x2 = x * x, x8 = (x2 * x2) * (x2 * x2), xp = p == 16 ? x8 * x8 : x8;
                ba += fabs(amp_a[k]) * xp; bb += fabs(amp_b ? amp_b[k] : 0.0) * xp;
            }
            const double kp = p == 16 ? 2.9966e-6 : 43.07 / 40320.0;   // K_p / p!
            if (!(fmax(ba, bb) * kp <= 1e-14)) continue;
            const double cost = (double)((n_c + kNoiseChunk - 1) / kNoiseChunk) * nf * 45.0 + (double)n_c * nf * 3.0 + (double)n_nodes * p;
            if (cost < best_cost) { best_cost = cost; best = plan; }
        }
    }
    return best;
}

// nk must hold the recurrence constants of spacing c*dtn (fill_noise_k(f, amp, nf, noise_c(plan)*dtn)); tmp: noise_coarse_count(n_nodes, plan) slots.
// weights of the plan: (c - 1) rows (fine offsets r = 1..c-1) of p doubles, row-major, in the constant tables
CA_HD const double* noise_weight_rows(int plan) {
    const int c = noise_c(plan);
    if (noise_p(plan) == 8) return LW.w[c == 2 ? 0 : (c == 4 ? 1 : (c == 8 ? 4 : (c == 16 ? 11 : 26)))];
    return LW16.w[c == 8 ? 0 : (c == 16 ? 7 : 22)];
}
// `wrow`: the plan's weight rows.  The one-atom kernel stages them in shared memory first: read through the constant cache with a
// run-time row index, the 31 x 16 rows of (c = 32, p = 16) thrashed it (S1 sweep: +7 % kernel time).
template <int PW>
CA_HD void interpolate_noise(int n_nodes, int c, const double* __restrict__ wrow, const double* __restrict__ tmp, int tstride, double* __restrict__ out,
                             int stride) {
    double win[PW];
#pragma unroll
    for (int i = 0; i < PW; ++i) win[i] = tmp[(size_t)i * tstride];
    for (int mc = 0; mc * c < n_nodes; ++mc) {
        if (mc > 0) {
#pragma unroll
            for (int i = 0; i < PW - 1; ++i) win[i] = win[i + 1];
            win[PW - 1] = tmp[(size_t)(mc + PW - 1) * tstride];
        }
        out[(size_t)(mc * c) * stride] = win[PW / 2 - 1];
        for (int r = 1; r < c; ++r) {
            const int j = mc * c + r;
            if (j >= n_nodes) break;
            const double* wv = wrow + (r - 1) * PW;
            double a = wv[0] * win[0], b = wv[1] * win[1];   // two chains
#pragma unroll
            for (int i = 2; i < PW; i += 2) { a = fma(wv[i], win[i], a); b = fma(wv[i + 1], win[i + 1], b); }
            out[(size_t)j * stride] = a + b;
        }
    }
}
// wrow: noise_weight_rows(plan), or a copy of those (c - 1) * p doubles in faster memory
template <int CH>
CA_HD void gen_noise_trace_c(const NoiseK* __restrict__ nk, int nf, int n_nodes, double dtn, int plan, const double* __restrict__ phases, uint64_t seed,
                             uint64_t gi, uint32_t stream, double* __restrict__ out, int stride, double* __restrict__ tmp, int tstride,
                             const double* __restrict__ wrow = nullptr) {
    const int c = noise_c(plan);
    if (c <= 1) { gen_noise_trace<CH>(nk, nf, n_nodes, dtn, 0.0, phases, seed, gi, stream, out, stride); return; }
    const int n_c = noise_coarse_count(n_nodes, plan), pw = noise_p(plan);
    gen_noise_trace<CH>(nk, nf, n_c, c * dtn, -(pw / 2 - 1.0) * c * dtn, phases, seed, gi, stream, tmp, tstride);
    if (!wrow) wrow = noise_weight_rows(plan);
    if (pw == 8) interpolate_noise<8>(n_nodes, c, wrow, tmp, tstride, out, stride);
    else interpolate_noise<16>(n_nodes, c, wrow, tmp, tstride, out, stride);
}

// ---------------------------------------------------------------------------------------------------
// Lindblad right-hand side on the reduced state (exact):
//   core block {1,r,p}: u[0..8] = n1, nr, np, Re/Im ρ1r, Re/Im ρ1p, Re/Im ρrp
//   spectator rows x ∈ {0, l}: v = (ρx1, ρxr, ρxp) as 6 reals each (only if ψ0 has that component)
//   populations ρ00, ρll are pure quadratures of (np, nr); ρ0l is constant.
// H (basis 1,r,p) = [[0,0,A],[0,-dr,B*],[A*,B,-dp]], A = Ωr/2 (σ1p), B = Ωb/2 (σpr).
// ---------------------------------------------------------------------------------------------------
struct Coef { double Ar, Ai, Br, Bi, dp, dr; };
struct Rates { double G0, G1, Gl, Gr, Gp, gp2, gr2; };

template <int NS>
CA_HD void lindblad_rhs(const Coef& c, const Rates& g, const double* u, double* du) {
    const double n1 = u[0], nr = u[1], np = u[2], xr = u[3], xi = u[4], yr = u[5], yi = u[6], zr = u[7], zi = u[8];
    const double imAy = c.Ai * yr - c.Ar * yi;  // Im(A conj(y))
    const double imBz = c.Br * zi + c.Bi * zr;  // Im(B z)
    du[0] = 2.0 * imAy + g.G1 * np;
    du[1] = -2.0 * imBz - g.Gr * nr;
    du[2] = 2.0 * (imBz - imAy) - g.Gp * np;
    {   // ρ1r : W = A z* - y B + dr x ; dρ = -iW - Γr/2 ρ
        double Wr = c.Ar * zr + c.Ai * zi - yr * c.Br + yi * c.Bi + c.dr * xr;
        double Wi = c.Ai * zr - c.Ar * zi - yr * c.Bi - yi * c.Br + c.dr * xi;
        du[3] = Wi - g.gr2 * xr; du[4] = -Wr - g.gr2 * xi;
    }
    {   // ρ1p : W = A (np - n1) - x B* + dp y
        double dn = np - n1;
        double Wr = c.Ar * dn - xr * c.Br - xi * c.Bi + c.dp * yr;
        double Wi = c.Ai * dn - xi * c.Br + xr * c.Bi + c.dp * yi;
        du[5] = Wi - g.gp2 * yr; du[6] = -Wr - g.gp2 * yi;
    }
    {   // ρrp : W = B* (np - nr) - x* A + (dp - dr) z
        double dm = np - nr, dpr = c.dp - c.dr, gz = g.gp2 + g.gr2;
        double Wr = c.Br * dm - xr * c.Ar - xi * c.Ai + dpr * zr;
        double Wi = -c.Bi * dm - xr * c.Ai + xi * c.Ar + dpr * zi;
        du[7] = Wi - gz * zr; du[8] = -Wr - gz * zi;
    }
#pragma unroll
    for (int o = 9; o + 6 <= NS; o += 6) {  // spectator rows: dv_j = i Σ_k v_k H_kj - γ_j/2 v_j
        const double v1r = u[o], v1i = u[o + 1], vrr = u[o + 2], vri = u[o + 3], vpr = u[o + 4], vpi = u[o + 5];
        // dv1 = i vp A*
        double t1r = vpr * c.Ar + vpi * c.Ai, t1i = vpi * c.Ar - vpr * c.Ai;  // vp A*
        du[o] = -t1i; du[o + 1] = t1r;
        // dvr = i(-dr vr + vp B) - Γr/2 vr
        double t2r = -c.dr * vrr + vpr * c.Br - vpi * c.Bi, t2i = -c.dr * vri + vpr * c.Bi + vpi * c.Br;
        du[o + 2] = -t2i - g.gr2 * vrr; du[o + 3] = t2r - g.gr2 * vri;
        // dvp = i(v1 A + vr B* - dp vp) - Γp/2 vp
        double t3r = v1r * c.Ar - v1i * c.Ai + vrr * c.Br + vri * c.Bi - c.dp * vpr;
        double t3i = v1r * c.Ai + v1i * c.Ar + vri * c.Br - vrr * c.Bi - c.dp * vpi;
        du[o + 4] = -t3i - g.gp2 * vpr; du[o + 5] = t3r - g.gp2 * vpi;
    }
}

// Dormand-Prince 5(4) tableau and Hairer's dense-output weights (see oracle).
namespace dp {
constexpr double a21 = 1.0 / 5.0, a31 = 3.0 / 40.0, a32 = 9.0 / 40.0, a41 = 44.0 / 45.0, a42 = -56.0 / 15.0, a43 = 32.0 / 9.0,
                 a51 = 19372.0 / 6561.0, a52 = -25360.0 / 2187.0, a53 = 64448.0 / 6561.0, a54 = -212.0 / 729.0,
                 a61 = 9017.0 / 3168.0, a62 = -355.0 / 33.0, a63 = 46732.0 / 5247.0, a64 = 49.0 / 176.0, a65 = -5103.0 / 18656.0,
                 b1 = 35.0 / 384.0, b3 = 500.0 / 1113.0, b4 = 125.0 / 192.0, b5 = -2187.0 / 6784.0, b6 = 11.0 / 84.0;
constexpr double c2 = 0.2, c3 = 0.3, c4 = 0.8, c5 = 8.0 / 9.0;
constexpr double e1 = 71.0 / 57600.0, e3 = -71.0 / 16695.0, e4 = 71.0 / 1920.0, e5 = -17253.0 / 339200.0, e6 = 22.0 / 525.0, e7 = -1.0 / 40.0;
constexpr double d1 = -12715105075.0 / 11282082432.0, d3 = 87487479700.0 / 32700410799.0, d4 = -10690763975.0 / 1880347072.0,
                 d5 = 701980252875.0 / 199316789632.0, d6 = -1453857185.0 / 822651844.0, d7 = 69997945.0 / 29380423.0;
}  // namespace dp

struct OdeOpts {
    double rtol, atol, dt0, dtmax;
    long long maxiters;
    int adaptive, dense;
};

// PI step-size controller of OrdinaryDiffEq's DP5 (β1 = 0.17, β2 = 0.04, γ = 0.9, q in [1/qmax, 1/qmin] = [0.1, 5], qoldinit = 1e-4):
// accept: q = EEst^β1 qold^-β2 / γ, dt <- h/q, qold <- max(EEst, 1e-4);  reject: dt <- h / min(1/qmin, EEst^β1/γ).
// The accept/reject DECISION (EEst² <= 1) is taken in double by the caller.  The fractional powers only size the next step, so
// they are evaluated in single precision through log2/exp2 (two MUFU ops instead of a ~60-deep double log/exp chain on the
// critical path of every step; relative error ~1e-7 in dt).  DiffEqBase itself evaluates them with an approximate `fastpower`.
struct StepCtrl {
    float l2q_old;   // log2(qold)
    CA_HD void reset() { l2q_old = -13.287712f; }   // log2(1e-4)
    static CA_HD float lg2(float x) {
#if defined(__CUDA_ARCH__)
        float r;
        asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));   // EEst² below 1e-38 flushes to log2(0) = -inf, which the clamps saturate like any tiny value
        return r;
#else
        return log2f(x);
#endif
    }
    static CA_HD float ex2(float x) {   // |x| <= 3.33 here: no denormal range to handle
#if defined(__CUDA_ARCH__)
        float r;
        asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
        return r;
#else
        return exp2f(x);
#endif
    }
    // returns the factor dt/h
    CA_HD double accept(double EEst2) {
        // EEst² = 0 needs no special case: log2(0) = -inf saturates both clamps (dt x 10, qold <- 1e-4), as any tiny value does
        const float l2E = 0.5f * lg2((float)EEst2);                                // log2(EEst); -inf / +inf saturate in the clamp
        float l2q = fmaf(0.17f, l2E, fmaf(-0.04f, l2q_old, 0.15200309f));             // - log2(0.9)
        l2q = fmaxf(-3.3219281f, fminf(2.3219281f, l2q));                             // q in [0.1, 5]
        l2q_old = fmaxf(l2E, -13.287712f);
        return (double)ex2(-l2q);
    }
    CA_HD double reject(double EEst2) const {
        const float l2E = 0.5f * lg2((float)EEst2);
        const float l2q = fminf(2.3219281f, fmaf(0.17f, l2E, 0.15200309f));
        return (double)ex2(-l2q);
    }
    // Both outcomes from one log2, selected by `acc`, with no branch between the error norm and the new dt (the caller decides
    // `acc` in double while the conversion and the two MUFU operations are already in flight).
    CA_HD double decide(double EEst2, bool acc) {
        const float l2E = 0.5f * lg2((float)EEst2);
        const float base = fmaf(0.17f, l2E, 0.15200309f);
        const float qa = fmaxf(-3.3219281f, fminf(2.3219281f, fmaf(-0.04f, l2q_old, base)));
        const float qr = fminf(2.3219281f, base);
        l2q_old = acc ? fmaxf(l2E, -13.287712f) : l2q_old;
        return (double)ex2(-(acc ? qa : qr));
    }
};

// Per-trajectory integrator state. NS = 9 (+6 per spectator row).
//
// Storage plan (from the first ncu profile: at 255 registers/thread only 8 warps/SM were resident and spill reloads
// + register shuffling dominated the stalls): the DP5 stage derivatives k1..k7 live in SHARED memory, laid out
// [slot][component][thread] (conflict-free), six slots per thread (k7 reuses k2's slot; FSAL k1 <- k7 is a toggle of
// the slot index, not a copy).  Registers hold only y, the stage input, the stage output and y_new.  The stage loop is
// rolled (dynamic slot / tableau indexing works in shared and constant memory), so the whole step is a few KB of code.
template <int NS, int BEAMS = 0>   // BEAMS: 0 decide at run time (all coefficient paths compiled), 1 Gaussian pair + free flight only,
                                     // 2 other beam kinds (free flight), 3 Gaussian pair + trap oscillation
struct Lane1 {
    static constexpr int kEvery = BEAMS == 2 ? kAnchorEveryGeneric : kAnchorEvery;   // attempts between re-anchorings of this variant
    static constexpr int kAB = BEAMS == 1 ? 7 : 9;   // anchor slots per beam: the free-flight Gaussian pair has no cubic term (slots 7, 8) and no trap rotation
    static constexpr int kCB = (BEAMS == 1 || BEAMS == 2) ? 4 : 6;   // doubles per stage in cb: Ar, Ai, Br, Bi (+ dp, dr unless the free-flight variants read them from registers)
    // trajectory
    double x0, y0, z0, vx, vy, vz, vzq;  // vzq: velocity used for "Vz" (vy when the B1 quirk is on)
    double dpc, drc;                      // constant detunings (free motion) incl. Δ0, δ0
    // noise cache
    const double* tab;                    // table base for this lane: tab[(trace*n_nodes + node)*stride]
    int stride, jc;
    double nr1, nr2, nr3, sr0, sr1, nb1, nb2, nb3, sb0, sb1;  // node values jc+1, jc+2 and the PREFETCHED jc+3, slopes of the intervals [jc, jc+1] and [jc+1, jc+2] (red / blue trace)
    double tmid, tnx, twend;              // time of node jc+1; the same, but +inf in the last interval (or without a table): the time at which the window
                                          // advances; end of the two cached intervals (+inf without a table)
    // integrator
    double* kb;                           // k storage: kb[(slot*NS + i)*kstride]
    double* cb;                           // stage coefficients cb[(n*6 + c)*kstride], n = stage 2..6, c = Ar,Ai,Br,Bi,dp,dr
    double* ab;                           // per beam q: ab[(q*kAB + i)*kstride], i = c(t_a) re,im | b1 re,im | b2 re,im | φ(t_a) | b3 re,im (kAB = 9) or without b3 (kAB = 7)
    int kstride, p;                       // p: slot of k1 (0/1); k2 and k7 live in slot 1-p; k3..k6 in slots 2..5
    bool win_ok;                          // the anchor's quadratic model of ln B passed its guard
    double t, h, dt;                      // committed time, size of the last accepted step (0: none yet), proposed next step size
    StepCtrl ctrl;
    double y[NS];                         // committed state (the end of the last accepted step)
    double tr0;                           // trace of ρ0
    double P[2], SB[2], SD[2];            // ρ00, ρll quadratures; SB / SD = Σ_j b_j / d_j (ρpp, ρrr)(stage j) of the last accepted step (dense output)
    unsigned iters_left;                  // attempts left before CA_ERR_MAXITERS (maxiters clamped to 2^32-1)
    int n_rhs, n_acc, n_rej, n_exact;     // n_exact: steps whose coefficients took the exact (generic / guard-failed) path
    int status;

    // anchor time and end of the anchor's validity window (registers: the check at the top of every step must not wait for shared memory)
    double anc_t0, anc_t1;
    CA_HD void set_anchor_window(double a, double b) { anc_t0 = a; anc_t1 = b; }
    CA_HD double anchor_t0() const { return anc_t0; }
    CA_HD double anchor_t1() const { return anc_t1; }
    // a failed trajectory parks at t = +inf, so that the step loop's only test is t < T
    CA_HD void fail(int code) { status = code; t = 1e300; }
    CA_HD double* kptr(int j) const {     // j = 1..7
        const int slot = (j == 1) ? p : ((j == 2 || j == 7) ? 1 - p : j - 1);
        return kb + (size_t)slot * NS * kstride;
    }

    CA_HD void coefficients(const DevModel& m, double tt, Coef& c) {
        double X, Y, Z, Vx, Vz;
        if (m.free_motion) {
            X = x0 + vx * tt; Y = y0 + vy * tt; Z = z0 + vz * tt;
            c.dp = dpc; c.dr = drc;
        } else {  // R, V: atom_sampler.jl:125-143
            double sr, cr, sz, cz;
            sincos_fast(m.wr * tt, &sr, &cr);
            sincos_fast(m.wz * tt, &sz, &cz);
            X = x0 * cr + vx / m.wr * sr; Y = y0 * cr + vy / m.wr * sr; Z = z0 * cz + vz / m.wz * sz;
            Vx = vx * cr - x0 * m.wr * sr;
            Vz = vzq * cz - z0 * m.wz * sz;
            double Dd = m.aDx * Vx + m.aDz * Vz, dd = m.bDx * Vx + m.bDz * Vz;
            if (m.doppler_z_legacy) { double Vzt = vz * cz - z0 * m.wz * sz; Dd = m.aDz * Vzt; dd = m.bDz * Vzt; }
            c.dp = m.Delta0 - m.dsign * Dd; c.dr = m.delta0 - m.dsign * dd;
        }
        double pr, pb;
        noise_at(m, tt, pr, pb);
        if (m.red.kind == CA_BEAM_GAUSS && m.blue.kind == CA_BEAM_GAUSS) {
            double o[1][4];
            gauss_batch<1>(m.red, m.blue, &X, &Y, &Z, &pr, &pb, o);
            c.Ar = o[0][0]; c.Ai = o[0][1]; c.Br = o[0][2]; c.Bi = o[0][3];
        } else {
            beam_half_omega(m.red, X, Y, Z, pr, &c.Ar, &c.Ai);
            beam_half_omega(m.blue, X, Y, Z, pb, &c.Br, &c.Bi);
        }
    }

    // ---- phase noise: Gridded(Linear) on the node table (rydberg_model.jl:166-167) --------------------------------
    // The lane caches THREE consecutive nodes (jc, jc+1 and jc+2), i.e. two intervals: a step shorter than the node
    // spacing never leaves the window, so the window is advanced at most once per step (advance_window below).
    CA_HD void window_at(const DevModel& m, double tt) {   // make t_jc <= tt < tnx
        int j = (int)(tt * m.inv_dtn);
        j = j < 0 ? 0 : (j > m.n_nodes - 2 ? m.n_nodes - 2 : j);
        if (j != jc) {
            const double* q0 = tab + (size_t)j * stride;
            const double* q1 = q0 + (size_t)m.n_nodes * stride;
            const int last = m.n_nodes - 1;
            if (j == jc + 1) { sr0 = sr1; nr1 = nr2; nr2 = nr3; sb0 = sb1; nb1 = nb2; nb2 = nb3; }   // nr3 / nb3 were requested one advance ago
            else {
                const double a0 = q0[0], b0 = q1[0];
                const int o2 = (j + 2 < last ? 2 : last - j) * stride;
                nr1 = q0[stride]; nb1 = q1[stride]; nr2 = q0[o2]; nb2 = q1[o2];
                sr0 = (nr1 - a0) * m.inv_dtn; sb0 = (nb1 - b0) * m.inv_dtn;
            }
            sr1 = (nr2 - nr1) * m.inv_dtn; sb1 = (nb2 - nb1) * m.inv_dtn;   // (zero slope past the last node)
            const int o3 = (j + 3 < last ? 3 : last - j) * stride;
            nr3 = q0[o3]; nb3 = q1[o3];   // prefetch: first needed at the next advance, so nothing waits for this load
            jc = j;
        }
        tmid = (j + 1) * m.dtn;
        tnx = (j >= m.n_nodes - 2) ? 1e300 : tmid;
        twend = (j >= m.n_nodes - 2) ? 1e300 : tmid + m.dtn;
    }
    // Sequential advance by one node, written with selects so that it is PREDICATED, not branched: the lanes of a warp cross node
    // boundaries at different steps (their step sizes differ), and as a branch this block ran, mostly empty, in every fifth step.
    CA_HD void advance_window_if(const DevModel& m, bool adv) {
        const int last = m.n_nodes - 1;
        const int jn = jc + 1;
        const double s1r = (nr3 - nr2) * m.inv_dtn, s1b = (nb3 - nb2) * m.inv_dtn;   // slope of the interval that becomes the right one
        sr0 = adv ? sr1 : sr0; sb0 = adv ? sb1 : sb0;
        sr1 = adv ? s1r : sr1; sb1 = adv ? s1b : sb1;
        nr1 = adv ? nr2 : nr1; nb1 = adv ? nb2 : nb1;
        nr2 = adv ? nr3 : nr2; nb2 = adv ? nb3 : nb2;
        if (adv) {   // the only memory operation: the next prefetch (nothing waits for it before the next advance)
            const double* q0 = tab + (size_t)jn * stride;
            const int o3 = (jn + 3 < last ? 3 : last - jn) * stride;
            nr3 = q0[o3]; nb3 = q0[(size_t)m.n_nodes * stride + o3];
        }
        const bool tail = jn >= m.n_nodes - 2;
        const double tm = tmid + m.dtn;   // (re-synchronised to (jc+1) dtn whenever window_at runs)
        tnx = adv ? (tail ? 1e300 : tm) : tnx;
        twend = adv ? (tail ? 1e300 : tm + m.dtn) : twend;
        tmid = adv ? tm : tmid;
        jc = adv ? jn : jc;
    }
    // Both cached intervals are anchored at their common node jc+1: φ(t) = φ_{jc+1} + (t - t_{jc+1}) · slope(left | right interval).
    // Branch-free: without a table the node values and slopes stay zero (init), so the result is 0 for any finite tt.
    CA_HD void phases_lin(const DevModel& m, double tt, double& pr, double& pb) const {  // needs t_jc <= tt < twend
        const double dtm = tt - tmid;
        const bool hi = !(tt < tnx);
        pr = fma(dtm, hi ? sr1 : sr0, nr1); pb = fma(dtm, hi ? sb1 : sb0, nb1);
    }
    CA_HD void phases_win(const DevModel& m, double tt, double& pr, double& pb) const {
        phases_lin(m, tt, pr, pb);
        if (m.xi != 0.0 && tt >= m.tau) pr += m.xi;
    }
    CA_HD void noise_at(const DevModel& m, double tt, double& pr, double& pb) const {   // stateless generic lookup
        pr = 0.0; pb = 0.0;
        if (tab) {
            int j = (int)(tt * m.inv_dtn);
            j = j < 0 ? 0 : (j > m.n_nodes - 2 ? m.n_nodes - 2 : j);
            const double* q0 = tab + (size_t)j * stride;
            const double* q1 = q0 + (size_t)m.n_nodes * stride;
            const double w = (tt - j * m.dtn) * m.inv_dtn;
            const double a0 = q0[0], a1 = q0[stride], b0 = q1[0], b1 = q1[stride];
            pr = fma(w, a1 - a0, a0); pb = fma(w, b1 - b0, b0);
        }
        if (m.xi != 0.0 && tt >= m.tau) pr += m.xi;
    }

    CA_HD bool fast_model(const DevModel& m) const {   // anchored coefficients are available
        if (!(NS <= 15 && ab != nullptr && m.xi == 0.0 && m.dtn > 0.0)) return false;   // (NS = 21 is register-bound: the anchored code costs it more in spills than it saves)
        if (m.free_motion) return true;                       // any beam kind in free flight
        return (BEAMS == 0 || BEAMS == 3) && gauss_pair(m);   // trap oscillation: Gaussian pairs only
    }
    // position at time tn_ (free flight or oscillation about the trap centre: atom_sampler.jl:125-143)
    CA_HD void position_at(const DevModel& m, double tn_, double& X, double& Y, double& Z) const {
        if (BEAMS == 1 || BEAMS == 2 || m.free_motion) { X = fma(vx, tn_, x0); Y = fma(vy, tn_, y0); Z = fma(vz, tn_, z0); }
        else {
            double sr, cr, sz, cz;
            sincos_fast(m.wr * tn_, &sr, &cr); sincos_fast(m.wz * tn_, &sz, &cz);
            X = x0 * cr + vx / m.wr * sr; Y = y0 * cr + vy / m.wr * sr; Z = z0 * cz + vz / m.wz * sz;
        }
    }
    CA_HD bool gauss_pair(const DevModel& m) const { return m.red.kind == CA_BEAM_GAUSS && m.blue.kind == CA_BEAM_GAUSS; }
    CA_HD double stage_time(int n, double tt0, double hh, double tn) const { return n == 4 ? tn : fma(MC.rc[n + 2], hh, tt0); }

    // Anchor for ANY beam kind (flat-top HG / LG, uniform): the exact coefficient at tt and a CUBIC model of ln c on [tt, tt + 3H]
    // from the exact values at tt, tt+H, tt+2H, tt+3H through complex logarithms of their ratios to the anchor.  µm-scale flat-top
    // beams (and their e^{-ikz} factor) vary faster than the wide Gaussian ones, hence one order more here and in the series below.
    CA_HD void refresh_anchor_generic(const DevModel& m, double tt, double H) {
        double pr = 0.0, pb = 0.0;
        if (tab) {
            if (!(tt < tnx) || tt < tmid - m.dtn) window_at(m, tt);
            phases_win(m, tt, pr, pb);
        }
        win_ok = true;
        const double iH = rcp_d(H);
#pragma unroll 1
        for (int q = 0; q < 2; ++q) {
            const DevBeam& bm = q == 0 ? m.red : m.blue;
            double* a = ab + (size_t)q * kAB * kstride;
            double c0r = 0.0, c0i = 0.0, inv0 = 0.0, l[3][2] = {{0.0, 0.0}, {0.0, 0.0}, {0.0, 0.0}};
#pragma unroll 1
            for (int i = 0; i < 4; ++i) {
                const double tn_ = fma((double)i, H, tt);
                double re, im;
                beam_half_omega(bm, fma(vx, tn_, x0), fma(vy, tn_, y0), fma(vz, tn_, z0), 0.0, &re, &im);
                if (i == 0) {
                    c0r = re; c0i = im;
                    const double n2 = re * re + im * im;
                    win_ok = win_ok && n2 > 1e-280;
                    inv0 = n2 > 1e-280 ? 1.0 / n2 : 0.0;
                } else {   // ln(c_i / c_0) = ln|r| + i arg r  (the flat-top fields carry e^{-ikz}: the phase moves by mrad over the window)
                    const double rr = (re * c0r + im * c0i) * inv0, ri = (im * c0r - re * c0i) * inv0, wr = rr - 1.0;
                    win_ok = win_ok && (wr * wr + ri * ri < 0.01);
                    // ln|r| and arg r for |r - 1| < 0.1 (guarded above) by short odd series (truncation < 1e-19): the library log1p / atan2
                    // are several hundred instructions of rarely executed code that the refresh would pull through the instruction cache
                    const double lr = 0.5 * log1p_small(fma(wr, 2.0 + wr, ri * ri)), li = atan_small(ri * rcp_d(rr));
                    if (i == 1) { l[0][0] = lr; l[0][1] = li; } else if (i == 2) { l[1][0] = lr; l[1][1] = li; } else { l[2][0] = lr; l[2][1] = li; }
                }
            }
            double sn = 0.0, cs = 1.0;
            const double ph = q == 0 ? pr : pb;
            if (tab) sincos_fast(ph, &sn, &cs);
            a[0] = c0r * cs - c0i * sn; a[kstride] = c0r * sn + c0i * cs;
#pragma unroll
            for (int c = 0; c < 2; ++c) {   // Newton forward differences -> monomials in τ
                const double d1 = l[0][c], d2 = l[1][c] - 2.0 * l[0][c], d3 = l[2][c] - 3.0 * l[1][c] + 3.0 * l[0][c];
                a[(2 + c) * kstride] = (d1 - 0.5 * d2 + (1.0 / 3.0) * d3) * iH;
                a[(4 + c) * kstride] = (0.5 * d2 - 0.5 * d3) * iH * iH;
                a[(7 + c) * kstride] = (1.0 / 6.0) * d3 * iH * iH * iH;
            }
            a[6 * kstride] = ph;
        }
        set_anchor_window(tt, fma(3.0, H, tt));
    }

    // Exact anchor at time tt: c(tt) with its noise phase and the quadratic model of ln B on [tt, tt + 2H].
    CA_HD void refresh_anchor(const DevModel& m, double tt, double H) {
        double pr = 0.0, pb = 0.0;
        if (tab) {
            if (!(tt < tnx) || tt < tmid - m.dtn) window_at(m, tt);
            phases_win(m, tt, pr, pb);
        }
        win_ok = true;
        const double iH = rcp_d(H);
#pragma unroll 1
        for (int q = 0; q < 2; ++q) {
            const DevBeam& bm = q == 0 ? m.red : m.blue;
            double* a = ab + (size_t)q * kAB * kstride;
            // rolled on purpose: this code runs once per 16 steps and must not evict the step loop from the instruction cache
            double L0r = 0.0, L0i = 0.0, L1r = 0.0, L1i = 0.0, L2r = 0.0, L2i = 0.0;
#pragma unroll 1
            for (int i = 0; i < 3; ++i) {
                const double tn_ = fma((double)i, H, tt);
                double X, Y, Z;
                position_at(m, tn_, X, Y, Z);
                const double rp = sqrt_d(fmax(X * X + Y * Y, 1e-300));
                L0r = L1r; L0i = L1i; L1r = L2r; L1i = L2i;
                gauss_log(bm, rp, Z, &L2r, &L2i);
                if (i == 0) {
                    double re, im, anc[5];
                    gauss_exact(bm, rp, Z, q == 0 ? pr : pb, &re, &im, anc);
                    a[0] = re; a[kstride] = im;
                }
            }
            const double L[3][2] = {{L0r, L0i}, {L1r, L1i}, {L2r, L2i}};
#pragma unroll
            for (int c = 0; c < 2; ++c) {   // ln B(tt + τ) - ln B(tt) = τ (b1 + τ b2)
                const double d1 = L[1][c] - L[0][c], d2 = L[2][c] - L[1][c];
                a[(2 + c) * kstride] = (1.5 * d1 - 0.5 * d2) * iH;
                const double b2 = 0.5 * (d2 - d1) * iH * iH;
                a[(4 + c) * kstride] = b2;
                win_ok = win_ok && fabs(d2 - d1) < 1e-6;
                // Re δ(τ) = τ(b1 + τ b2) on the whole validity window 0 <= τ <= 2H: bounded here so that eval_anchored does not re-check it
                if (c == 0) win_ok = win_ok && 2.0 * H * (fabs(a[2 * kstride]) + 2.0 * H * fabs(b2)) < 3e-4;
            }
            a[6 * kstride] = q == 0 ? pr : pb;
        }
        set_anchor_window(tt, fma(2.0, H, tt));
        if ((BEAMS == 0 || BEAMS == 3) && !m.free_motion) {   // cos/sin(ω t_a) for the time-dependent Doppler shifts (the Gaussian layout leaves slots 7,8,16,17 free)
            double sr, cr, sz, cz;
            sincos_fast(m.wr * tt, &sr, &cr); sincos_fast(m.wz * tt, &sz, &cz);
            ab[7 * kstride] = cr; ab[8 * kstride] = sr; ab[16 * kstride] = cz; ab[17 * kstride] = sz;
        }
    }

    // c(t) = c(t_a) exp(δ) at the five stage times.  GP (both beams Gaussian): quadratic ln B, guards |Re δ| < 3e-4, |Im δ| < 6e-3.
    // Otherwise (flat-top HG / LG, uniform): cubic model and three more series orders, guards |Re δ| < 5e-3, |Im δ| < 0.1 (the
    // e^{-ikz} factor of those fields turns the phase by ~1e-2 rad over an anchor window).  One compiled copy runs per launch.
    template <bool GP>
    CA_HD bool eval_anchored(const DevModel& m, double tt0, double hh, double tn) {
        double ca[2][2], b1[2][2], b2[2][2], b3[2][2], pa[2];
#pragma unroll
        for (int q = 0; q < 2; ++q) {
            const double* a = ab + (size_t)q * kAB * kstride;
            ca[q][0] = a[0]; ca[q][1] = a[kstride];
            b1[q][0] = a[2 * kstride]; b1[q][1] = a[3 * kstride]; b2[q][0] = a[4 * kstride]; b2[q][1] = a[5 * kstride];
            pa[q] = a[6 * kstride];
            b3[q][0] = GP ? 0.0 : a[7 * kstride]; b3[q][1] = GP ? 0.0 : a[8 * kstride];
        }
        const double ta = anchor_t0();
        double worst[2] = {0.0, 0.0};   // one accumulator per beam: two short chains instead of one long one
#if CA_ANCH_V2
        int worst_hi = 0;
#endif
#if CA_ANCH_V2
        // Gaussian pair, second form: everything is a function of τ = t - t_a alone.  The piecewise-linear noise phase of the two cached
        // intervals (anchored at their common node, phases_lin) is folded into the imaginary part of the quadratic model once per step,
        //     Im δ(τ) = τ (b1i + s + τ b2i) + [φ_mid - φ_a + (t_a - t_mid) s],      s = slope of the interval τ falls into,
        // so a stage costs one FMA for τ, one comparison and two FMAs per beam for Im δ (it was 12 FP64 operations, now 6), and the
        // per-evaluation guard is the largest |Im δ| compared as an integer (high word of the double) instead of a running sum.
        if (GP) {
            const double tau0 = tt0 - ta, thr = tnx - ta, tam = ta - tmid;
            const double g0[2] = {b1[0][1] + sr0, b1[1][1] + sb0}, g1[2] = {b1[0][1] + sr1, b1[1][1] + sb1};
            const double pm[2] = {nr1 - pa[0], nb1 - pa[1]};
            const double c0[2] = {fma(tam, sr0, pm[0]), fma(tam, sb0, pm[1])}, c1[2] = {fma(tam, sr1, pm[0]), fma(tam, sb1, pm[1])};
#pragma unroll
            for (int n = 0; n < 5; ++n) {
                const double tau = n == 4 ? tn - ta : fma(MC.rc[n + 2], hh, tau0);
                const bool hi = !(tau < thr);
#pragma unroll
                for (int q = 0; q < 2; ++q) {
                    const double dr = tau * fma(tau, b2[q][0], b1[q][0]);
                    const double di = fma(tau, fma(tau, b2[q][1], hi ? g1[q] : g0[q]), hi ? c1[q] : c0[q]);
                    const double d2 = di * di;
                    { const int ah = abs_hi32(di); worst_hi = ah > worst_hi ? ah : worst_hi; }
                    const double ed = fma(dr, fma(dr, fma(dr, 1.0 / 6.0, 0.5), 1.0), 1.0);
                    const double cs = fma(d2, fma(d2, 1.0 / 24.0, -0.5), 1.0);
                    const double sn = di * fma(d2, fma(d2, 1.0 / 120.0, -1.0 / 6.0), 1.0);
                    const double er = ed * cs, ei = ed * sn;
                    cb[(n * kCB + 2 * q) * kstride] = ca[q][0] * er - ca[q][1] * ei;
                    cb[(n * kCB + 2 * q + 1) * kstride] = ca[q][0] * ei + ca[q][1] * er;
                }
            }
        } else
#endif
#pragma unroll
        for (int n = 0; n < 5; ++n) {
            const double tt = n == 4 ? tn : fma(MC.rc[n + 2], hh, tt0);
            const double tau = tt - ta;
            double ph[2];
            phases_lin(m, tt, ph[0], ph[1]);   // fast_model() excludes the CZ phase jump (ξ = 0); no table -> zeros, no branch
#pragma unroll
            for (int q = 0; q < 2; ++q) {
                // δ = ln B(t) - ln B(t_a) + i (φ(t) - φ(t_a)) = dr + i di ;  e^δ = e^{dr} (cos di + i sin di) by short real series
                double dr, di, ed, cs, sn;
                if (GP) {
                    dr = tau * fma(tau, b2[q][0], b1[q][0]);
                    di = fma(tau, fma(tau, b2[q][1], b1[q][1]), ph[q] - pa[q]);
                } else {
                    dr = tau * fma(tau, fma(tau, b3[q][0], b2[q][0]), b1[q][0]);
                    di = fma(tau, fma(tau, fma(tau, b3[q][1], b2[q][1]), b1[q][1]), ph[q] - pa[q]);
                }
                const double d2 = di * di;
                if (GP) worst[q] += d2;   // |Re δ| is bounded once per anchor (refresh_anchor); Σ di² >= max di² is the per-step guard
                else worst[q] = fma(400.0 * dr, dr, worst[q] + d2);
                if (GP) {
                    ed = fma(dr, fma(dr, fma(dr, 1.0 / 6.0, 0.5), 1.0), 1.0);
                    cs = fma(d2, fma(d2, 1.0 / 24.0, -0.5), 1.0);
                    sn = di * fma(d2, fma(d2, 1.0 / 120.0, -1.0 / 6.0), 1.0);
                } else {
                    ed = fma(dr, fma(dr, fma(dr, fma(dr, fma(dr, fma(dr, 1.0 / 720.0, 1.0 / 120.0), 1.0 / 24.0), 1.0 / 6.0), 0.5), 1.0), 1.0);
                    cs = fma(d2, fma(d2, fma(d2, fma(d2, 1.0 / 40320.0, -1.0 / 720.0), 1.0 / 24.0), -0.5), 1.0);
                    sn = di * fma(d2, fma(d2, fma(d2, fma(d2, 1.0 / 362880.0, -1.0 / 5040.0), 1.0 / 120.0), -1.0 / 6.0), 1.0);
                }
                const double er = ed * cs, ei = ed * sn;
                cb[(n * kCB + 2 * q) * kstride] = ca[q][0] * er - ca[q][1] * ei;
                cb[(n * kCB + 2 * q + 1) * kstride] = ca[q][0] * ei + ca[q][1] * er;
            }
        }
        if (GP && (BEAMS == 0 || BEAMS == 3) && !m.free_motion) {
            // trap oscillation: V(t) from the anchored cos/sin(ω t_a) by a rotation through ω τ (|ω τ| < 1e-2: series exact to 1e-18),
            // then Δ_D, δ_D as in `coefficients`
            const double cra = ab[7 * kstride], sra = ab[8 * kstride], cza = ab[16 * kstride], sza = ab[17 * kstride];
#pragma unroll
            for (int n = 0; n < 5; ++n) {
                const double tt = n == 4 ? tn : fma(MC.rc[n + 2], hh, tt0);
                const double xr = m.wr * (tt - ta), xz = m.wz * (tt - ta), xr2 = xr * xr, xz2 = xz * xz;
                const double cr_ = fma(xr2, fma(xr2, 1.0 / 24.0, -0.5), 1.0), sr_ = xr * fma(xr2, fma(xr2, 1.0 / 120.0, -1.0 / 6.0), 1.0);
                const double cz_ = fma(xz2, fma(xz2, 1.0 / 24.0, -0.5), 1.0), sz_ = xz * fma(xz2, fma(xz2, 1.0 / 120.0, -1.0 / 6.0), 1.0);
                const double cr = cra * cr_ - sra * sr_, sr = sra * cr_ + cra * sr_, cz = cza * cz_ - sza * sz_, sz = sza * cz_ + cza * sz_;
                const double Vx = vx * cr - x0 * m.wr * sr, Vz = vzq * cz - z0 * m.wz * sz;
                double Dd = m.aDx * Vx + m.aDz * Vz, dd = m.bDx * Vx + m.bDz * Vz;
                if (m.doppler_z_legacy) { const double Vzt = vz * cz - z0 * m.wz * sz; Dd = m.aDz * Vzt; dd = m.bDz * Vzt; }
                cb[(n * kCB + 4) * kstride] = m.Delta0 - m.dsign * Dd; cb[(n * kCB + 5) * kstride] = m.delta0 - m.dsign * dd;
            }
        } else if (!(BEAMS == 1 || BEAMS == 2)) {   // (the free-flight variants read dpc / drc directly)
#pragma unroll
            for (int n = 0; n < 5; ++n) { cb[(n * kCB + 4) * kstride] = dpc; cb[(n * kCB + 5) * kstride] = drc; }
        }
        // Gaussian pair: every |di| < 6e-3 (di⁶/720 < 7e-17), |dr| < 3e-4 (dr⁴/24 < 4e-16); else |di| < 0.1 (di¹⁰/3.6e6 < 3e-17), |dr| < 5e-3 (dr⁷/5040 < 2e-20)
#if CA_ANCH_V2
        if (GP) return worst_hi < 0x3f789374;   // every |Im δ| < 6e-3 (high word of 6.0e-3 = 0x3f789374bc6a7efa)
#endif
        return worst[0] + worst[1] < (GP ? 3.6e-5 : 1e-2);
    }

    // Coefficients of the five distinct stage times of one step (c2..c5 and t+h; stage 7 and the next stage 1 share the
    // last) into the shared-memory buffer cb.  `refresh`: warp-uniform request to re-anchor now.
    CA_HD void stage_coefficients(const DevModel& m, double tt0, double hh, double tn, bool refresh) {
        bool done = false;
        const bool gp = (BEAMS == 1 || BEAMS == 3) ? true : (BEAMS == 2 ? false : gauss_pair(m));   // uniform over the launch
        if (tab) advance_window_if(m, !(tt0 < tnx) && tt0 < twend);   // the common sequential advance (never in the last interval: tnx = +inf)
        // ONE decision on the common path: no re-anchoring requested, the anchor model still covers the step (win_ok is only ever set by
        // an anchor refresh, i.e. it implies fast_model), the step lies inside the two cached noise intervals (time only moves forward,
        // so only the upper ends are checked; a fresh trajectory starts with an empty window and an empty anchor range and therefore
        // takes the careful path below)
        bool ok = !refresh && win_ok && tn <= anchor_t1() && tt0 < tnx && tn < twend;
        if (!ok && fast_model(m)) {
            ok = true;
            if (tab) {
                if (!(tt0 < tnx) || tt0 < tmid - m.dtn) window_at(m, tt0);
                ok = tn < twend;
            }
            if (refresh || !(tn <= anchor_t1()) || tt0 < anchor_t0()) {   // two steps of headroom over the re-anchoring cadence
                if (gp) refresh_anchor(m, tt0, (0.5 * kEvery + 1.0) * hh);
                else refresh_anchor_generic(m, tt0, ((kEvery + 2.0) / 3.0) * hh);
            }
            ok = ok && win_ok;
        }
        if (ok) done = gp ? eval_anchored<true>(m, tt0, hh, tn) : eval_anchored<false>(m, tt0, hh, tn);
        if (!done) {   // generic / guard-failed: exact evaluation stage by stage (one rolled copy of the code)
            ++n_exact;
#pragma unroll 1
            for (int n = 0; n < 5; ++n) {
                Coef c;
                coefficients(m, stage_time(n, tt0, hh, tn), c);
                double* o = cb + (size_t)n * kCB * kstride;
                o[0] = c.Ar; o[kstride] = c.Ai; o[2 * kstride] = c.Br; o[3 * kstride] = c.Bi;
                if (kCB == 6) { o[4 * kstride] = c.dp; o[5 * kstride] = c.dr; }
            }
        }
    }

    // error-norm contribution helpers (DiffEqBase: |err| / (atol + rtol max(|u0|,|u1|)), complex modulus)
    static CA_HD double nrm_real(double e, double a, double b, double atol, double rtol) {
        double sc = atol + rtol * fmax(fabs(a), fabs(b));
        double r = e * rcp_norm(sc);
        return r * r;
    }
    static CA_HD double nrm_cplx(double er, double ei, double ar, double ai, double br, double bi, double atol, double rtol) {
        double sc = atol + rtol * sqrt_norm(fmax(ar * ar + ai * ai, br * br + bi * bi));
        double isc = rcp_norm(sc);
        return 2.0 * (er * er + ei * ei) * (isc * isc);  // both (i,j) and (j,i) entries of the full matrix
    }

    CA_HD void init(const DevModel& m, const Rates& g, const OdeOpts& oo, const double* smp, const double* table, int tstride,
                    double t0, double* kbase, int kstr, double* cbase, double* abase) {
        kb = kbase; kstride = kstr; p = 0; cb = cbase; ab = abase; win_ok = false;
        if (ab) set_anchor_window(1e300, -1e300);
        if (m.atom_motion) { x0 = smp[0]; y0 = smp[1]; z0 = smp[2]; vx = smp[3]; vy = smp[4]; vz = smp[5]; }
        else { x0 = y0 = z0 = vx = vy = vz = 0.0; }
        vzq = m.vz_from_vy ? vy : vz;
        {
            double Dd = m.aDx * vx + m.aDz * vzq, dd = m.bDx * vx + m.bDz * vzq;
            if (m.doppler_z_legacy) { Dd = m.aDz * vz; dd = m.bDz * vz; }
            dpc = m.Delta0 - m.dsign * Dd; drc = m.delta0 - m.dsign * dd;
        }
        tab = table; stride = tstride; jc = -5; nr1 = nr2 = nr3 = sr0 = sr1 = nb1 = nb2 = nb3 = sb0 = sb1 = 0.0; tmid = 0.0; tnx = table ? -1e300 : 1e300; twend = tnx;   // with a table: an empty window, so the first step loads it (window_at)
        // initial state
        y[0] = m.rho0[1]; y[1] = m.rho0[2]; y[2] = m.rho0[3];
        y[3] = m.rho0[5 + 8]; y[4] = m.rho0[5 + 9];      // (1,r)
        y[5] = m.rho0[5 + 10]; y[6] = m.rho0[5 + 11];    // (1,p)
        y[7] = m.rho0[5 + 14]; y[8] = m.rho0[5 + 15];    // (r,p)
        if (NS >= 15) {
            if (m.rows & 1) {  // row 0: (0,1),(0,r),(0,p)
#pragma unroll
                for (int q = 0; q < 6; ++q) y[9 + q] = m.rho0[5 + q];
            } else {           // only row l present: ρ_l1 = conj ρ_1l etc.
                y[9] = m.rho0[5 + 12]; y[10] = -m.rho0[5 + 13]; y[11] = m.rho0[5 + 16]; y[12] = -m.rho0[5 + 17];
                y[13] = m.rho0[5 + 18]; y[14] = -m.rho0[5 + 19];
            }
        }
        if (NS >= 21) {
            y[15] = m.rho0[5 + 12]; y[16] = -m.rho0[5 + 13]; y[17] = m.rho0[5 + 16]; y[18] = -m.rho0[5 + 17];
            y[19] = m.rho0[5 + 18]; y[20] = -m.rho0[5 + 19];
        }
        P[0] = m.rho0[0]; P[1] = m.rho0[4];
        tr0 = ((m.rho0[0] + m.rho0[1]) + (m.rho0[2] + m.rho0[3])) + m.rho0[4];
        t = t0; h = 0.0; iters_left = oo.maxiters > 0xffffffffLL ? 0xffffffffu : (unsigned)oo.maxiters; n_rhs = 0; n_acc = 0; n_rej = 0; n_exact = 0; status = 0;
        ctrl.reset();
        double k1[NS];
        Coef cf0;
        coefficients(m, t, cf0);
        lindblad_rhs<NS>(cf0, g, y, k1);
        {
            double* k1p = kptr(1);
#pragma unroll
            for (int i = 0; i < NS; ++i) k1p[i * kstride] = k1[i];
        }
        const double g1[2] = {g.G0 * y[2], g.Gl * y[2] + g.Gr * y[1]};
        n_rhs = 1;
        SB[0] = SB[1] = SD[0] = SD[1] = 0.0;
        dt = oo.dt0;
        if (dt <= 0.0) {  // OrdinaryDiffEq ode_determine_initdt (Hairer hinit), norm over all 25 matrix entries
            double d0 = 0.0, d1 = 0.0;
            auto acc_r = [&](double a, double fa) { double sk = oo.atol + fabs(a) * oo.rtol; d0 += (a / sk) * (a / sk); d1 += (fa / sk) * (fa / sk); };
            auto acc_c = [&](double ar, double ai, double fr, double fi) {
                double sk = oo.atol + sqrt(ar * ar + ai * ai) * oo.rtol;
                d0 += 2.0 * (ar * ar + ai * ai) / (sk * sk); d1 += 2.0 * (fr * fr + fi * fi) / (sk * sk);
            };
            acc_r(y[0], k1[0]); acc_r(y[1], k1[1]); acc_r(y[2], k1[2]); acc_r(P[0], g1[0]); acc_r(P[1], g1[1]);
#pragma unroll
            for (int i = 3; i < NS; i += 2) acc_c(y[i], y[i + 1], k1[i], k1[i + 1]);
            {   // the constant ρ0l coherence contributes to d0 only
                double ar = m.rho0[5 + 6], ai = m.rho0[5 + 7];
                double sk = oo.atol + sqrt(ar * ar + ai * ai) * oo.rtol;
                d0 += 2.0 * (ar * ar + ai * ai) / (sk * sk);
            }
            d0 = sqrt(d0 / 25.0); d1 = sqrt(d1 / 25.0);
            double dt0 = (d0 < 1e-5 || d1 < 1e-5) ? 1e-6 : 0.01 * d0 / d1;
            dt0 = fmin(dt0, oo.dtmax);
            double us[NS], f1[NS];
#pragma unroll
            for (int i = 0; i < NS; ++i) us[i] = y[i] + dt0 * k1[i];
            Coef c2;
            coefficients(m, t + dt0, c2);
            lindblad_rhs<NS>(c2, g, us, f1);
            n_rhs++;
            double d2 = 0.0;
            auto dif_r = [&](double a, double df) { double sk = oo.atol + fabs(a) * oo.rtol; d2 += (df / sk) * (df / sk); };
            dif_r(y[0], f1[0] - k1[0]); dif_r(y[1], f1[1] - k1[1]); dif_r(y[2], f1[2] - k1[2]);
            dif_r(P[0], g.G0 * us[2] - g1[0]); dif_r(P[1], g.Gl * us[2] + g.Gr * us[1] - g1[1]);
#pragma unroll
            for (int i = 3; i < NS; i += 2) {
                double sk = oo.atol + sqrt(y[i] * y[i] + y[i + 1] * y[i + 1]) * oo.rtol;
                double dr = f1[i] - k1[i], di = f1[i + 1] - k1[i + 1];
                d2 += 2.0 * (dr * dr + di * di) / (sk * sk);
            }
            d2 = sqrt(d2 / 25.0) / dt0;
            double mx = fmax(d1, d2);
            double dt1 = (mx <= 1e-15) ? fmax(1e-6, dt0 * 1e-3) : pow(10.0, -(2.0 + log10(mx)) / 5.0);
            dt = fmin(100.0 * dt0, fmin(dt1, oo.dtmax));
        }
    }

    // Attempt one step towards `tstop`.  An accepted step is committed at once on the straight-line path (y <- y_new, t <- t+h, k1 <- k7 by
    // a toggle of the slot index); a rejection, the rare case, leaves early.  The state at the START of the last accepted step, which only
    // the dense output needs, is not kept: state_at() re-forms it from the stage derivatives that are still in shared memory.
    CA_HD void attempt(const DevModel& m, const Rates& g, const OdeOpts& oo, double tstop, bool refresh = false) {
        if (iters_left-- == 0u) { fail(CA_ERR_MAXITERS); return; }
        double hh = dt < oo.dtmax ? dt : oo.dtmax;   // (dt is finite here: a NaN norm has already failed the trajectory)
        bool clipped = false;
        if (t + hh >= tstop || (tstop - (t + hh)) < 1e-13 * fabs(tstop)) { hh = tstop - t; clipped = true; }
        const double tn = clipped ? tstop : t + hh;
        constexpr int NC = NS;   // (a two-phase form that integrated the spectator rows after the core block was 2-6 % slower: DESIGN.md §10)
        double us[NC], ko[NC];
        stage_coefficients(m, t, hh, tn, refresh);
        Coef cf;
        // population quadratures ρ00' = Γ0 ρpp, ρll' = Γl ρpp + Γr ρrr: the b/e/d-weighted sums of the STAGE VALUES of ρpp and ρrr are
        // accumulated (the rates are applied once per step to the sums, not to every stage)
        // (with CA_TRACE_LL only the ρpp sums are needed: ρll' = Γl ρpp + Γr ρrr follows from the conserved trace)
        double SBp = MC.rb[1] * y[2], SEp = MC.re[1] * y[2], SDp = MC.rd[1] * y[2];
#if !CA_TRACE_LL
        double SBr = MC.rb[1] * y[1], SEr = MC.re[1] * y[1], SDr = MC.rd[1] * y[1];
#endif
#pragma unroll
        for (int s = 2; s <= 6; ++s) {   // fully unrolled: tableau weights become constant-bank operands, slots immediates
            // y + Σ_j (h a_sj) k_j with the weights scaled once per stage: s-1 multiplications instead of NS
#pragma unroll
            for (int i = 0; i < NC; ++i) us[i] = y[i];
#pragma unroll
            for (int j = 1; j < s; ++j) {
                const double w = hh * MC.ra[s][j];
                const double* kp = kptr(j);
#pragma unroll
                for (int i = 0; i < NC; ++i) us[i] = fma(w, kp[i * kstride], us[i]);
            }
            {
                const double* o = cb + (size_t)(s - 2) * kCB * kstride;
                cf.Ar = o[0]; cf.Ai = o[kstride]; cf.Br = o[2 * kstride]; cf.Bi = o[3 * kstride];
                if (BEAMS == 1 || BEAMS == 2) { cf.dp = dpc; cf.dr = drc; }   // free flight: the detunings are constants of the trajectory
                else { cf.dp = o[4 * kstride]; cf.dr = o[5 * kstride]; }
            }
            lindblad_rhs<NC>(cf, g, us, ko);
            {
                const double wb = MC.rb[s], we = MC.re[s], wd = MC.rd[s];
                SBp = fma(wb, us[2], SBp); SEp = fma(we, us[2], SEp); SDp = fma(wd, us[2], SDp);
#if !CA_TRACE_LL
                SBr = fma(wb, us[1], SBr); SEr = fma(we, us[1], SEr); SDr = fma(wd, us[1], SDr);
#endif
            }
            double* ks = kptr(s);
#pragma unroll
            for (int i = 0; i < NC; ++i) ks[i * kstride] = ko[i];
        }
        // y_new and the error combination share one pass over k1,k3..k6 (ko still holds k6)
        double er[NC], un[NS], Pn[2];
#pragma unroll
        for (int i = 0; i < NC; ++i) { un[i] = MC.rb[6] * ko[i]; er[i] = MC.re[6] * ko[i]; }
#pragma unroll
        for (int j = 1; j <= 5; ++j) {
            if (j == 2) continue;
            const double wb = MC.rb[j], we = MC.re[j];
            const double* kp = kptr(j);
            double kv[NC];
#pragma unroll
            for (int i = 0; i < NC; ++i) kv[i] = kp[i * kstride];
#pragma unroll
            for (int i = 0; i < NC; ++i) { un[i] = fma(wb, kv[i], un[i]); er[i] = fma(we, kv[i], er[i]); }
        }
#pragma unroll
        for (int i = 0; i < NC; ++i) un[i] = fma(hh, un[i], y[i]);
        lindblad_rhs<NC>(cf, g, un, ko);  // stage 7 = next step's stage 1; same coefficients as stage 6 (c6 = c7 = 1)
        {
            double* k7p = kptr(7);
#pragma unroll
            for (int i = 0; i < NC; ++i) { k7p[i * kstride] = ko[i]; er[i] = fma(MC.re[7], ko[i], er[i]); }   // error / h (h² goes into EEst² once)
        }
        SEp = fma(MC.re[7], un[2], SEp); SDp = fma(MC.rd[7], un[2], SDp);
        Pn[0] = fma(hh * g.G0, SBp, P[0]);
#if CA_TRACE_LL
        Pn[1] = tr0 - ((Pn[0] + un[0]) + (un[1] + un[2]));
#else
        SEr = fma(MC.re[7], un[1], SEr); SDr = fma(MC.rd[7], un[1], SDr);
        Pn[1] = fma(hh, fma(g.Gl, SBp, g.Gr * SBr), P[1]);
#endif
        n_rhs += 6;
        if (oo.adaptive) {
#if CA_TRACE_LL
            const double e00 = g.G0 * SEp;   // the error of ρll is minus the sum of the other populations' errors
            double s = nrm_real(e00, P[0], Pn[0], oo.atol, oo.rtol) + nrm_real((e00 + er[0]) + (er[1] + er[2]), P[1], Pn[1], oo.atol, oo.rtol);
#else
            double s = nrm_real(g.G0 * SEp, P[0], Pn[0], oo.atol, oo.rtol) + nrm_real(fma(g.Gl, SEp, g.Gr * SEr), P[1], Pn[1], oo.atol, oo.rtol);
#endif
#pragma unroll
            for (int i = 0; i < 3; ++i) s += nrm_real(er[i], y[i], un[i], oo.atol, oo.rtol);
#pragma unroll
            for (int i = 3; i < NC; i += 2) s += nrm_cplx(er[i], er[i + 1], y[i], y[i + 1], un[i], un[i + 1], oo.atol, oo.rtol);
            double EEst2 = s * (hh * hh * (1.0 / 25.0));
            const bool acc = EEst2 <= 1.0;                 // NaN and overflow fail it and are sorted out on the rare side
            const double prop = hh * ctrl.decide(EEst2, acc);
            const bool keep = acc && clipped && dt > prop;   // a step clipped at tstop does not shrink the running step size
            dt = keep ? dt : prop;
            if (!acc) {   // rejected (rare): nothing is committed, the next attempt restarts from the same y, k1 with the smaller dt
                if (!(EEst2 <= 1e300)) { fail(CA_ERR_NONFINITE); return; }
                n_rej++;
                if (dt < 1e-14 * fmax(1.0, fabs(t))) fail(CA_ERR_NONFINITE);
                return;
            }
        } else {
            dt = oo.dt0;
            if (!(fabs(un[0] + un[1] + un[2]) < 1e300)) { fail(CA_ERR_NONFINITE); return; }   // fixed step beyond the stability limit
        }
        // accepted: commit on the straight-line path
        n_acc++;
#pragma unroll
        for (int i = 0; i < NS; ++i) y[i] = un[i];
        P[0] = Pn[0]; P[1] = Pn[1];
        t = tn; h = hh; p ^= 1;
        SB[0] = SBp; SD[0] = SDp;
#if !CA_TRACE_LL
        SB[1] = SBr; SD[1] = SDr;
#endif
    }

    // State at time T inside the last accepted step [t-h, t] (Hairer contd5), Hermitian-packed ρ into out[0..24].  The increment
    // r2 = y(t) - y(t-h) = h Σ_j b_j k_j is re-formed from the stage derivatives (k1 of that step now sits in the slot kptr(2) names, its
    // k7 is the current k1), and the state at the start of the step is y - r2.
    CA_HD void state_at(const DevModel& m, double T, double* out) const {
        double s[NS], p0, p1;
        if (h == 0.0 || T >= t) {
#pragma unroll
            for (int i = 0; i < NS; ++i) s[i] = y[i];
            p0 = P[0]; p1 = P[1];
        } else {
            double th = (T - (t - h)) / h, th1 = 1.0 - th;
            const double *k1 = kptr(2), *k3 = kptr(3), *k4 = kptr(4), *k5 = kptr(5), *k6 = kptr(6), *k7 = kptr(1);
            double yo1 = 0.0, yo2 = 0.0;
#pragma unroll
            for (int i = 0; i < NS; ++i) {
                const int o = i * kstride;
                const double r2 = h * (MC.rb[1] * k1[o] + MC.rb[3] * k3[o] + MC.rb[4] * k4[o] + MC.rb[5] * k5[o] + MC.rb[6] * k6[o]);
                const double yo = y[i] - r2;
                if (i == 1) yo1 = yo;
                if (i == 2) yo2 = yo;
                double r3 = h * k1[o] - r2;
                double r4 = r2 - h * k7[o] - r3;
                double r5 = h * (MC.rd[1] * k1[o] + MC.rd[3] * k3[o] + MC.rd[4] * k4[o] + MC.rd[5] * k5[o] + MC.rd[6] * k6[o] + MC.rd[7] * k7[o]);
                s[i] = yo + th * (r2 + th1 * (r3 + th * (r4 + th1 * r5)));
            }
            {
                // stage-1 / stage-7 derivatives of the quadratures from ρpp, ρrr at the step ends; increments and d-sums from SB, SD
                const double g10 = m.G0 * yo2, g11 = m.Gl * yo2 + m.Gr * yo1, g70 = m.G0 * y[2], g71 = m.Gl * y[2] + m.Gr * y[1];
                double r2 = h * (m.G0 * SB[0]), r3 = h * g10 - r2, r4 = r2 - h * g70 - r3, r5 = h * (m.G0 * SD[0]);
                p0 = (P[0] - r2) + th * (r2 + th1 * (r3 + th * (r4 + th1 * r5)));
#if CA_TRACE_LL
                (void)g11; (void)g71;
                p1 = tr0 - ((p0 + s[0]) + (s[1] + s[2]));   // the dense output is a linear combination of stage derivatives: it keeps the trace too
#else
                r2 = h * (m.Gl * SB[0] + m.Gr * SB[1]); r3 = h * g11 - r2; r4 = r2 - h * g71 - r3; r5 = h * (m.Gl * SD[0] + m.Gr * SD[1]);
                p1 = (P[1] - r2) + th * (r2 + th1 * (r3 + th * (r4 + th1 * r5)));
#endif
            }
        }
#pragma unroll
        for (int i = 0; i < 25; ++i) out[i] = 0.0;
        out[0] = p0; out[1] = s[0]; out[2] = s[1]; out[3] = s[2]; out[4] = p1;
        out[5 + 8] = s[3]; out[5 + 9] = s[4]; out[5 + 10] = s[5]; out[5 + 11] = s[6]; out[5 + 14] = s[7]; out[5 + 15] = s[8];
        out[5 + 6] = m.rho0[5 + 6]; out[5 + 7] = m.rho0[5 + 7];  // ρ0l is constant
        if (NS >= 15) {
            if (m.rows & 1) {
#pragma unroll
                for (int q = 0; q < 6; ++q) out[5 + q] = s[9 + q];
            } else {
                out[5 + 12] = s[9]; out[5 + 13] = -s[10]; out[5 + 16] = s[11]; out[5 + 17] = -s[12]; out[5 + 18] = s[13]; out[5 + 19] = -s[14];
            }
        }
        if (NS >= 21) {
            out[5 + 12] = s[15]; out[5 + 13] = -s[16]; out[5 + 16] = s[17]; out[5 + 17] = -s[18]; out[5 + 18] = s[19]; out[5 + 19] = -s[20];
        }
    }
};

// Second moment of a Hermitian-packed 5x5 matrix: packed ρ·ρ (CA_RHO2_MATSQ) or |ρ_ij|² (CA_RHO2_ABS2).
// FULL=false: only the block structure of the rows-free case (ρ00, ρll scalars + 3x3 core) is non-zero.
template <bool FULL>
CA_HD void second_moment(const double* p, int mode, double* o) {
    // unpack upper triangle: idx(a,b), a<b
    constexpr int off[5][5] = {{-1, 0, 1, 2, 3}, {-1, -1, 4, 5, 6}, {-1, -1, -1, 7, 8}, {-1, -1, -1, -1, 9}, {-1, -1, -1, -1, -1}};
    if (mode == CA_RHO2_ABS2) {
#pragma unroll
        for (int a = 0; a < 5; ++a) o[a] = p[a] * p[a];
#pragma unroll
        for (int q = 0; q < 10; ++q) { o[5 + 2 * q] = p[5 + 2 * q] * p[5 + 2 * q] + p[6 + 2 * q] * p[6 + 2 * q]; o[6 + 2 * q] = 0.0; }
        return;
    }
    double re[5][5], im[5][5];
#pragma unroll
    for (int a = 0; a < 5; ++a)
#pragma unroll
        for (int b = 0; b < 5; ++b) {
            if (a == b) { re[a][b] = p[a]; im[a][b] = 0.0; }
            else if (a < b) { re[a][b] = p[5 + 2 * off[a][b]]; im[a][b] = p[6 + 2 * off[a][b]]; }
            else { re[a][b] = p[5 + 2 * off[b][a]]; im[a][b] = -p[6 + 2 * off[b][a]]; }
        }
#pragma unroll
    for (int a = 0; a < 5; ++a)
#pragma unroll
        for (int b = a; b < 5; ++b) {
            bool core_ab = (a >= 1 && a <= 3 && b >= 1 && b <= 3);
            bool live = FULL || core_ab || (a == b);
            double sr = 0.0, si = 0.0;
            if (live) {
#pragma unroll
                for (int k = 0; k < 5; ++k) {
                    bool kin = FULL || (core_ab ? (k >= 1 && k <= 3) : (k == a));
                    if (kin) { sr += re[a][k] * re[k][b] - im[a][k] * im[k][b]; si += re[a][k] * im[k][b] + im[a][k] * re[k][b]; }
                }
            }
            if (a == b) o[a] = sr;
            else { o[5 + 2 * off[a][b]] = sr; o[6 + 2 * off[a][b]] = si; }
        }
}

}  // namespace ca
