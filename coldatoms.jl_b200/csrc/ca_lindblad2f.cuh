// ca_lindblad2f.cuh -- per-lane code of the FULL-FORM two-atom kernel (k_lindblad2_full): the whole 25x25 density matrix, any initial
// state, every parameterisation of the reference's two-atom drivers.
//
// The blocked kernel (ca_lindblad2.cuh) covers simulation_czlp (src/cz_model.jl:107-171) for ψ0 inside the {0,1,r,p}² block.  This one is
// the general path SURVEY App. A6 asks to fall back to, and the home of the legacy drivers of src/two_atom_model.jl:
//   * ψ0 with |l> components, or a density matrix ρ0 as the initial state;
//   * two_atom_simulation (:89-190): per-atom lasers and detunings, constant V_rr, beam-frame offsets;
//   * direct_CZ_simulation (:192-319): direct |1> <-> |r> coupling, two consecutive time segments with the state carried over and a
//     phase on the couplings of the second.
//
// Same representation as the blocked kernel: the Hermitian ρ is stored as the REAL matrix M[I][J] = Re ρ_IJ + Im ρ_IJ (625 reals), lane J
// (0..24; lanes 25..31 idle) owns column J, I = a1 + 5 a2.  G = Hρ needs only column J of ρ; with the transposed reals M[J][I] (one
// exchange through shared memory) a coupling c = H_II' maps the source reals (m, mt) = (M[I'][J], M[J][I']) onto S = Im G + Re G and
// D = Im G - Re G (four FMAs), and dM[I][J] = D_IJ + S_JI - (γ_I + γ_J)/2 M[I][J] + feeding needs one real from the transposed owner
// (second exchange).  Row indices are compile-time constants, the coefficients are uniform over the lanes: straight-line code.
// All Runge-Kutta vectors live in shared memory (25 reals per lane do not fit the register file next to the RHS).
#pragma once
#include <cstring>

#include "ca_lindblad2.cuh"
#include "ca_model.h"

namespace ca {

constexpr int kF = 25;                 // reals per lane
constexpr int kFRS = 26;               // row stride of the 25x25 exchange matrices (even: rows stay 16-byte aligned)
constexpr int kFXN = kF * kFRS;        // doubles per exchange matrix
enum { FB_Y = 0, FB_1, FB_2, FB_3, FB_4, FB_5, FB_6, FB_YT, kFB };   // per-lane buffers fb[(buf*kF + row)*32]
constexpr int kCoefF = 24;             // doubles per stage in the coefficient area
constexpr int kWarpDoublesF = 2 * kFXN + kFB * kF * 32 + kStages2 * kCoefF + 12;

// extension of the device model for the legacy parameterisations (ca_model.two_atom_flags)
struct DevModel2x {
    DevBeam red2, blue2;
    double Delta0_2, delta0_2, V_rr, seg2_phase;
    int flags;
    int _pad;
};

// host-side derivation of the legacy extension (ca_model.two_atom_flags)
inline int derive_model2x(const ca_model& mm, DevModel2x* x) {
    std::memset(x, 0, sizeof *x);
    x->flags = mm.two_atom_flags;
    if (mm.two_atom_flags & CA_2A_PER_ATOM_LASERS) {
        int rc;
        if ((rc = derive_beam(mm.red2, &x->red2))) return rc;
        if ((rc = derive_beam(mm.blue2, &x->blue2))) return rc;
        x->Delta0_2 = mm.Delta0_2; x->delta0_2 = mm.delta0_2;
    }
    x->V_rr = mm.V_rr; x->seg2_phase = mm.seg2_phase;
    return CA_OK;
}

// coefficients of one stage time: A = Ωr/2 on |1><p|, B = Ωb/2 on |p><r|, C = direct Ω/2 on |1><r|, detunings (H_pp = -dp, H_rr = -dr), V
struct CoefF { double ar1, ai1, br1, bi1, cr1, ci1, ar2, ai2, br2, bi2, cr2, ci2, dp1, dr1, dp2, dr2, V, pad[kCoefF - 17]; };

// host: ρ0 (625 complex, column-major) or |ψ0><ψ0| -> per-lane columns of M, out[lane * kF + row]
inline void rho0_columns(const double* psi0, const double* rho0, double* out) {
    for (int J = 0; J < 25; ++J)
        for (int I = 0; I < 25; ++I) {
            double re, im;
            if (rho0) { re = rho0[2 * (I + 25 * J)]; im = rho0[2 * (I + 25 * J) + 1]; }
            else {
                const double ar = psi0[2 * I], ai = psi0[2 * I + 1], br = psi0[2 * J], bi = psi0[2 * J + 1];   // ρ_IJ = ψ_I conj(ψ_J)
                re = ar * br + ai * bi; im = ai * br - ar * bi;
            }
            out[J * kF + I] = re + im;
        }
    for (int l = 25; l < 32; ++l)
        for (int I = 0; I < 25; ++I) out[l * kF + I] = 0.0;
}

// Exact coefficients of the five stage times, one (quantity, stage) pair per lane (q = 0..3: red1, blue1, red2, blue2; q = 4: detunings, V)
template <class W>
CA_HD void eval_coefficients2f(W& w, const DevModel& m, const DevModel2x& x, const double* tab, int seg, double t0, double hh, double tn) {
    const int q = w.lane % 5, si = w.lane / 5;
    const bool shift_after = (x.flags & CA_2A_SHIFT_AFTER_MOTION) != 0, per_atom = (x.flags & CA_2A_PER_ATOM_LASERS) != 0;
    const bool direct = (x.flags & CA_2A_DIRECT) != 0;
    if (si < kStages2) {
        const double tt = si == kStages2 - 1 ? tn : t0 + MC.rc[si + 2] * hh;
        double* o = w.coef + si * kCoefF;
        if (q < 4) {
            const int a = q >> 1;
            double pos[3];
            atom_position2(m, w.smp + 6 * a, tt, pos);
            if (shift_after) { pos[0] += m.center[a][0]; pos[1] += m.center[a][1]; pos[2] += m.center[a][2]; }
            double ph = 0.0;
            if (tab) {
                int j = (int)(tt * m.inv_dtn);
                j = j < 0 ? 0 : (j > m.n_nodes - 2 ? m.n_nodes - 2 : j);
                const double* p = tab + (size_t)q * m.n_nodes + j;
                const double wgt = (tt - j * m.dtn) * m.inv_dtn;
                ph = p[0] + wgt * (p[1] - p[0]);
            }
            if (!(q & 1) && tt >= m.tau) ph += m.xi;  // ϕr(t) = t < τ ? 0 : ξ on both red beams (cz_model.jl:136-137,48)
            if (direct && seg == 1) ph += x.seg2_phase;   // exp(1.0im * phase) * Ω / 2 (two_atom_model.jl:292-296)
            const DevBeam& bm = (q & 1) ? ((a == 1 && per_atom) ? x.blue2 : m.blue) : ((a == 1 && per_atom) ? x.red2 : m.red);
            double re = 0.0, im = 0.0;
            if (!(direct && !(q & 1))) {   // (the direct drive has no red beam)
                if (bm.kind == CA_BEAM_GAUSS) {
                    double anc[5];
                    const double rp = sqrt_d(fmax(pos[0] * pos[0] + pos[1] * pos[1], 1e-300));
                    gauss_exact(bm, rp, pos[2], ph, &re, &im, anc);
                } else beam_half_omega(bm, pos[0], pos[1], pos[2], ph, &re, &im);
            }
            // slots per atom: A (0,1), B (2,3), C (4,5)
            const int slot = 6 * a + (direct ? ((q & 1) ? 4 : 0) : ((q & 1) ? 2 : 0));
            o[slot] = re; o[slot + 1] = im;
            if (direct && (q & 1)) { o[6 * a + 2] = 0.0; o[6 * a + 3] = 0.0; }
            if (!direct && (q & 1)) { o[6 * a + 4] = 0.0; o[6 * a + 5] = 0.0; }
        } else {
            double pos[2][3], vel[2][2];
#pragma unroll
            for (int a = 0; a < 2; ++a) {
                const double* s = w.smp + 6 * a;
                atom_position2(m, s, tt, pos[a]);
                if (shift_after) { pos[a][0] += m.center[a][0]; pos[a][1] += m.center[a][1]; pos[a][2] += m.center[a][2]; }
                if (m.free_motion) { vel[a][0] = s[3]; vel[a][1] = m.vz_from_vy ? s[4] : s[5]; }
                else {  // V: atom_sampler.jl:135-143
                    double sr, cr, sz, cz;
                    sincos_fast(m.wr * tt, &sr, &cr); sincos_fast(m.wz * tt, &sz, &cz);
                    vel[a][0] = s[3] * cr - s[0] * m.wr * sr;
                    vel[a][1] = (m.vz_from_vy ? s[4] : s[5]) * cz - s[2] * m.wz * sz;
                    if (m.doppler_z_legacy) vel[a][1] = s[5] * cz - s[2] * m.wz * sz;   // the legacy 2-arg Δ(Vz, ...) uses the true Vz
                }
                if (m.doppler_z_legacy && m.free_motion) vel[a][1] = s[5];
            }
#pragma unroll
            for (int a = 0; a < 2; ++a) {
                const double D0 = (a == 1 && per_atom) ? x.Delta0_2 : m.Delta0, d0 = (a == 1 && per_atom) ? x.delta0_2 : m.delta0;
                const double Dd = m.aDx * vel[a][0] + m.aDz * vel[a][1], dd = m.bDx * vel[a][0] + m.bDz * vel[a][1];
                o[12 + 2 * a] = direct ? 0.0 : D0 - m.dsign * Dd;      // dp: H_pp = -dp
                o[13 + 2 * a] = direct ? -D0 : d0 - m.dsign * dd;      // dr: H_rr = -dr  (direct drive: H_rr = +Δ0, two_atom_model.jl:266-267)
            }
            if (x.flags & CA_2A_CONST_VRR) o[16] = x.V_rr;   // t -> V_rr (two_atom_model.jl:165,283)
            else {
                const double dx = pos[0][0] - pos[1][0], dy = pos[0][1] - pos[1][1], dz = pos[0][2] - pos[1][2];
                const double r2 = dx * dx + dy * dy + dz * dz;
                o[16] = m.c6 * rcp_d(1e-18 + r2 * r2 * r2);      // get_V: cz_model.jl:12-17
            }
        }
    }
    w.sync();
}

// dM/dt of this lane's column.  tp[I] returns the transposed partner M[J][I] of u[I].
template <class W>
CA_HD void rhs2f(W& w, const DevModel& m, const CoefF& cf, const double (&u)[kF], double (&du)[kF], double (&tp)[kF]) {
    double* xm = w.xm;
    double* xp = w.xp;
    const int J = w.lane;
    const bool act = J < kF;
    const int b1c = J % 5, b2c = J / 5;
    if (act) {
#pragma unroll
        for (int I = 0; I < kF; ++I) xm[I * kFRS + J] = u[I];
    }
    w.sync();
#pragma unroll
    for (int I = 0; I < kF; ++I) tp[I] = act ? xm[J * kFRS + I] : 0.0;
    const double Gp = m.G0 + m.G1 + m.Gl;
    const double gam[5] = {0.0, 0.0, m.Gr, Gp, 0.0};
    const double gJ = act ? 0.5 * (gam[b1c] + gam[b2c]) : 0.0;
    const double E1[5] = {0.0, 0.0, -cf.dr1, -cf.dp1, 0.0}, E2[5] = {0.0, 0.0, -cf.dr2, -cf.dp2, 0.0};
#pragma unroll
    for (int a2 = 0; a2 < 5; ++a2) {
#pragma unroll
        for (int a1 = 0; a1 < 5; ++a1) {
            const int I = a1 + 5 * a2;
            double e = E1[a1] + E2[a2];
            if (a1 == 2 && a2 == 2) e += cf.V;   // |rr><rr| (get_V, cz_model.jl:12-17,63)
            double S = e * u[I], D = -e * tp[I];
            // H1 ⊗ 1: level order 0,1,r,p,l; couplings 1<->p (A), p<->r (B), 1<->r (C)
            if (a1 == 1) { CA_MMAC(S, D, cf.ar1, cf.ai1, u[3 + 5 * a2], tp[3 + 5 * a2]); CA_MMAC(S, D, cf.cr1, cf.ci1, u[2 + 5 * a2], tp[2 + 5 * a2]); }
            if (a1 == 2) { CA_MMAC(S, D, cf.br1, -cf.bi1, u[3 + 5 * a2], tp[3 + 5 * a2]); CA_MMAC(S, D, cf.cr1, -cf.ci1, u[1 + 5 * a2], tp[1 + 5 * a2]); }
            if (a1 == 3) { CA_MMAC(S, D, cf.ar1, -cf.ai1, u[1 + 5 * a2], tp[1 + 5 * a2]); CA_MMAC(S, D, cf.br1, cf.bi1, u[2 + 5 * a2], tp[2 + 5 * a2]); }
            // 1 ⊗ H2
            if (a2 == 1) { CA_MMAC(S, D, cf.ar2, cf.ai2, u[a1 + 15], tp[a1 + 15]); CA_MMAC(S, D, cf.cr2, cf.ci2, u[a1 + 10], tp[a1 + 10]); }
            if (a2 == 2) { CA_MMAC(S, D, cf.br2, -cf.bi2, u[a1 + 15], tp[a1 + 15]); CA_MMAC(S, D, cf.cr2, -cf.ci2, u[a1 + 5], tp[a1 + 5]); }
            if (a2 == 3) { CA_MMAC(S, D, cf.ar2, -cf.ai2, u[a1 + 5], tp[a1 + 5]); CA_MMAC(S, D, cf.br2, cf.bi2, u[a1 + 10], tp[a1 + 10]); }
            if (act) xp[I * kFRS + J] = S;
            du[I] = fma(-(0.5 * (gam[a1] + gam[a2]) + gJ), u[I], D);
        }
    }
    // population feeding of the jump operators (JumpOperatorsTwo, cz_model.jl:2-9): |a><p| with Γ_a (a = 0, 1, l) and |l><r| with Γr per atom:
    // dρ_(a,a2),(a,b2) += Γ ρ_(s,a2),(s,b2) for atom 1 (s = p, r), likewise for atom 2
    {
        const double rate[5] = {m.G0, m.G1, 0.0, 0.0, m.Gl};
        const double w1 = act ? rate[b1c] : 0.0, w1r = (act && b1c == 4) ? m.Gr : 0.0;
        const double w2 = act ? rate[b2c < 5 ? b2c : 0] : 0.0, w2r = (act && b2c == 4) ? m.Gr : 0.0;
        const int cb2 = act ? b2c : 0, cb1 = act ? b1c : 0;
#pragma unroll
        for (int a2 = 0; a2 < 5; ++a2) {   // atom 1: rows with a1 == b1c
            const double vp = xm[(3 + 5 * a2) * kFRS + 3 + 5 * cb2], vr = xm[(2 + 5 * a2) * kFRS + 2 + 5 * cb2];
            const double add = fma(w1, vp, w1r * vr);
#pragma unroll
            for (int a1 = 0; a1 < 5; ++a1)
                if (a1 == 0 || a1 == 1 || a1 == 4) du[a1 + 5 * a2] += (b1c == a1) ? add : 0.0;
        }
#pragma unroll
        for (int a1 = 0; a1 < 5; ++a1) {   // atom 2: rows with a2 == b2c
            const double vp = xm[(a1 + 15) * kFRS + cb1 + 15], vr = xm[(a1 + 10) * kFRS + cb1 + 10];
            const double add = fma(w2, vp, w2r * vr);
#pragma unroll
            for (int a2 = 0; a2 < 5; ++a2)
                if (a2 == 0 || a2 == 1 || a2 == 4) du[a1 + 5 * a2] += (b2c == a2) ? add : 0.0;
        }
    }
    w.sync();
#pragma unroll
    for (int I = 0; I < kF; ++I) du[I] = act ? du[I] + xp[J * kFRS + I] : 0.0;
}

// Adds the state o (this lane's column of M at one output time) and its second moment to the per-warp accumulators a0[(which*kF + row)*32].
template <class W>
CA_HD void accumulate_output2f(W& w, const DevModel& m, const double (&o)[kF], double* a0) {
    double* xm = w.xm;   // xm and xp are contiguous: 1300 doubles hold the 625 complex entries of ρ
    const int J = w.lane;
    const bool act = J < kF;
    if (act) {
#pragma unroll
        for (int I = 0; I < kF; ++I) xm[I * kFRS + J] = o[I];
    }
    w.sync();
    double cr[kF], ci[kF];   // column J of ρ: ρ_IJ = ((M_IJ + M_JI) + i (M_IJ - M_JI)) / 2
#pragma unroll
    for (int I = 0; I < kF; ++I) { const double t = act ? xm[J * kFRS + I] : 0.0; cr[I] = 0.5 * (o[I] + t); ci[I] = 0.5 * (o[I] - t); }
    w.sync();
#pragma unroll
    for (int I = 0; I < kF; ++I) a0[I * 32] += o[I];
    if (m.rho2_mode == CA_RHO2_MATSQ) {   // (ρ·ρ)_IJ = Σ_K ρ_IK ρ_KJ: row I of ρ from shared memory, column J from registers
        if (act) {
#pragma unroll
            for (int I = 0; I < kF; ++I) { xm[2 * (I * kF + J)] = cr[I]; xm[2 * (I * kF + J) + 1] = ci[I]; }
        }
        w.sync();
#pragma unroll 1
        for (int I = 0; I < kF; ++I) {
            double ar = 0.0, ai = 0.0;
            const double* row = xm + 2 * I * kF;
#pragma unroll
            for (int K = 0; K < kF; ++K) CA_CMAC(ar, ai, row[2 * K], row[2 * K + 1], cr[K], ci[K]);
            a0[(kF + I) * 32] += ar + ai;
        }
        w.sync();
    } else {
#pragma unroll
        for (int I = 0; I < kF; ++I) a0[(kF + I) * 32] += cr[I] * cr[I] + ci[I] * ci[I];   // |ρ_IJ|² (real: Im part 0)
    }
}

// One time segment of one trajectory: DP5 (OrdinaryDiffEq semantics, see oracle) from the state in fb[FB_Y] over tspan[0..nt); outputs are
// ADDED to acc[(j*2 + which)*kF*32 ...] (pointer offset by the lane); the final state is left in fb[FB_Y] (next segment's initial state).
template <class W>
CA_HD void integrate2f(W& w, const DevModel& m, const DevModel2x& x, const OdeOpts& oo, const double* tspan, int nt, int seg, const double* tab,
                       double* acc, Stats2& st) {
    const double rtol = oo.rtol, atol = oo.atol;
    double* fb = w.kb;   // fb[(buf*kF + row)*32], lane offset included
#define CA_FB(buf, s) fb[((buf) * kF + (s)) * 32]
    const bool act = w.lane < kF;
    double un[kF], k7[kF], unt[kF], ut[kF];
#pragma unroll
    for (int s = 0; s < kF; ++s) un[s] = CA_FB(FB_Y, s);
    const double t0 = tspan[0], tend = tspan[nt - 1];
    double t = t0, h = 0.0, tcommit = t0, dt = oo.dt0;
    bool pending = false;
    StepCtrl2 ctrl;
    ctrl.reset();
    long long iters = 0;
    int status = 0, n_rhs = 0, n_acc = 0, n_rej = 0;
    const CoefF* coef = reinterpret_cast<const CoefF*>(w.coef);
    eval_coefficients2f(w, m, x, tab, seg, t0, 0.0, t0);
    rhs2f(w, m, coef[kStages2 - 1], un, k7, unt);
#pragma unroll
    for (int s = 0; s < kF; ++s) { CA_FB(FB_YT, s) = unt[s]; CA_FB(FB_1, s) = k7[s]; }
    n_rhs = 1;
    if (dt <= 0.0) {  // OrdinaryDiffEq initial step (Hairer), norms over all 625 entries
        double d0 = 0.0, d1 = 0.0, isk2[kF];
#pragma unroll
        for (int s = 0; s < kF; ++s) {
            const double a2 = 0.5 * (un[s] * un[s] + unt[s] * unt[s]);   // |ρ_IJ|²
            const double sk = atol + sqrt(a2) * rtol;
            isk2[s] = act ? 1.0 / (sk * sk) : 0.0;   // (I,J) and (J,I) share |err|² = (e² + e_t²)/2: weight 1 per real
            d0 += isk2[s] * un[s] * un[s];
            d1 += isk2[s] * k7[s] * k7[s];
        }
        d0 = sqrt(w.sum(d0) / 625.0);
        d1 = sqrt(w.sum(d1) / 625.0);
        double dt0 = (d0 < 1e-5 || d1 < 1e-5) ? 1e-6 : 0.01 * d0 / d1;
        dt0 = fmin(dt0, oo.dtmax);
#pragma unroll
        for (int s = 0; s < kF; ++s) ut[s] = un[s] + dt0 * k7[s];
        eval_coefficients2f(w, m, x, tab, seg, t0, 0.0, t0 + dt0);
        double k2[kF], dum[kF];
        rhs2f(w, m, coef[kStages2 - 1], ut, k2, dum);
        n_rhs++;
        double d2 = 0.0;
#pragma unroll
        for (int s = 0; s < kF; ++s) { const double a = k2[s] - k7[s]; d2 += isk2[s] * a * a; }
        d2 = sqrt(w.sum(d2) / 625.0) / dt0;
        const double mx = fmax(d1, d2);
        const double dt1 = (mx <= 1e-15) ? fmax(1e-6, dt0 * 1e-3) : pow(10.0, -(2.0 + log10(mx)) / 5.0);
        dt = fmin(100.0 * dt0, fmin(dt1, oo.dtmax));
    }
    for (int j = 0; j < nt; ++j) {
        const double T = tspan[j];
        const double tstop = oo.dense ? tend : T;
        while (status == 0 && tcommit < T) {
            if (pending) {  // commit the accepted step (FSAL)
#pragma unroll
                for (int s = 0; s < kF; ++s) { CA_FB(FB_Y, s) = un[s]; CA_FB(FB_1, s) = k7[s]; }
                t = tcommit; h = 0.0; pending = false;
            }
            if (++iters > oo.maxiters) { status = CA_ERR_MAXITERS; break; }
            double hh = fmin(dt, oo.dtmax);
            bool clipped = false;
            if (t + hh >= tstop || (tstop - (t + hh)) < 1e-13 * fabs(tstop)) { hh = tstop - t; clipped = true; }
            const double tn = clipped ? tstop : t + hh;
            eval_coefficients2f(w, m, x, tab, seg, t, hh, tn);
            double kk[kF], dum[kF];
#pragma unroll
            for (int s = 0; s < kF; ++s) ut[s] = fma(hh * MC.a21, CA_FB(FB_1, s), CA_FB(FB_Y, s));
            rhs2f(w, m, coef[0], ut, kk, dum);   // k2
#pragma unroll
            for (int s = 0; s < kF; ++s) {
                CA_FB(FB_2, s) = kk[s];
                ut[s] = fma(hh, fma(MC.a32, kk[s], MC.a31 * CA_FB(FB_1, s)), CA_FB(FB_Y, s));
            }
            rhs2f(w, m, coef[1], ut, kk, dum);   // k3
#pragma unroll
            for (int s = 0; s < kF; ++s) {
                CA_FB(FB_3, s) = kk[s];
                ut[s] = fma(hh, fma(MC.a43, kk[s], fma(MC.a42, CA_FB(FB_2, s), MC.a41 * CA_FB(FB_1, s))), CA_FB(FB_Y, s));
            }
            rhs2f(w, m, coef[2], ut, kk, dum);   // k4
#pragma unroll
            for (int s = 0; s < kF; ++s) {
                CA_FB(FB_4, s) = kk[s];
                ut[s] = fma(hh, fma(MC.a54, kk[s], fma(MC.a53, CA_FB(FB_3, s), fma(MC.a52, CA_FB(FB_2, s), MC.a51 * CA_FB(FB_1, s)))), CA_FB(FB_Y, s));
            }
            rhs2f(w, m, coef[3], ut, kk, dum);   // k5
#pragma unroll
            for (int s = 0; s < kF; ++s) {
                CA_FB(FB_5, s) = kk[s];
                ut[s] = fma(hh, fma(MC.a65, kk[s], fma(MC.a64, CA_FB(FB_4, s), fma(MC.a63, CA_FB(FB_3, s), fma(MC.a62, CA_FB(FB_2, s), MC.a61 * CA_FB(FB_1, s))))),
                            CA_FB(FB_Y, s));
            }
            rhs2f(w, m, coef[4], ut, kk, dum);   // k6
            // error weights are parked in ut (free now): er = Σ e_j k_j without the stage-7 term
#pragma unroll
            for (int s = 0; s < kF; ++s) {
                CA_FB(FB_6, s) = kk[s];
                const double k1v = CA_FB(FB_1, s), k3v = CA_FB(FB_3, s), k4v = CA_FB(FB_4, s), k5v = CA_FB(FB_5, s);
                un[s] = fma(hh, fma(MC.b6, kk[s], fma(MC.b5, k5v, fma(MC.b4, k4v, fma(MC.b3, k3v, MC.b1 * k1v)))), CA_FB(FB_Y, s));
                ut[s] = fma(MC.e6, kk[s], fma(MC.e5, k5v, fma(MC.e4, k4v, fma(MC.e3, k3v, MC.e1 * k1v))));
            }
            rhs2f(w, m, coef[4], un, k7, unt);   // stage 7 = next step's stage 1 (c6 = c7 = 1)
            n_rhs += 6;
            bool accept = true;
            if (oo.adaptive) {
                double sacc = 0.0;
#pragma unroll
                for (int s = 0; s < kF; ++s) {
                    const double e = hh * fma(MC.e7, k7[s], ut[s]);
                    const double yv = CA_FB(FB_Y, s), yt = CA_FB(FB_YT, s);
                    const double isc = StepCtrl2::rcpn(atol + rtol * StepCtrl2::sqrtn(0.5 * fmax(yv * yv + yt * yt, un[s] * un[s] + unt[s] * unt[s])));
                    sacc += act ? (e * e) * (isc * isc) : 0.0;
                }
                sacc = w.sum(sacc);
                const double EEst2 = sacc * (1.0 / 625.0);
                const bool acc_ = EEst2 <= 1.0;                // NaN and overflow fail it and are sorted out on the rare side
                const double prop = hh * ctrl.decide(EEst2, acc_);
                const bool keep = acc_ && clipped && dt > prop;   // a step clipped at tstop does not shrink the running step size
                dt = keep ? dt : prop;
                if (!acc_) {
                    if (!(EEst2 <= 1e300)) { status = CA_ERR_NONFINITE; break; }
                    accept = false;
                    n_rej++;
                    if (dt < 1e-14 * fmax(1.0, fabs(t))) { status = CA_ERR_NONFINITE; break; }
                }
            } else {
                dt = oo.dt0;
                double tr = 0.0;
#pragma unroll
                for (int s = 0; s < kF; ++s) tr += fabs(un[s]);
                if (!(w.sum(act ? tr : 0.0) < 1e300)) { status = CA_ERR_NONFINITE; break; }   // a fixed step beyond the stability limit blows up
            }
            if (accept) {
                n_acc++; h = hh; tcommit = tn; pending = true;
#pragma unroll
                for (int s = 0; s < kF; ++s) CA_FB(FB_YT, s) = unt[s];   // becomes the partner of y at the commit
            }
        }
        // ---- output at T (Hairer contd5 inside the last accepted step [t, tcommit]; FB_Y / FB_1 still hold its start)
        double o[kF];
        if (status != 0) {
#pragma unroll
            for (int s = 0; s < kF; ++s) o[s] = 0.0;
        } else if (!pending || T >= tcommit) {
#pragma unroll
            for (int s = 0; s < kF; ++s) o[s] = pending ? un[s] : CA_FB(FB_Y, s);
        } else {
            const double th = (T - t) / h, th1 = 1.0 - th;
#pragma unroll
            for (int s = 0; s < kF; ++s) {
                const double yv = CA_FB(FB_Y, s), k1v = CA_FB(FB_1, s);
                const double k3 = CA_FB(FB_3, s), k4 = CA_FB(FB_4, s), k5 = CA_FB(FB_5, s), k6 = CA_FB(FB_6, s);
                const double r2 = un[s] - yv, r3 = h * k1v - r2, r4 = r2 - h * k7[s] - r3;
                const double r5 = h * (MC.d1 * k1v + MC.d3 * k3 + MC.d4 * k4 + MC.d5 * k5 + MC.d6 * k6 + MC.d7 * k7[s]);
                o[s] = yv + th * (r2 + th1 * (r3 + th * (r4 + th1 * r5)));
            }
        }
        accumulate_output2f(w, m, o, acc + (size_t)j * 2 * kF * 32);
    }
    if (pending) {   // leave the final state in FB_Y for a following segment
#pragma unroll
        for (int s = 0; s < kF; ++s) CA_FB(FB_Y, s) = un[s];
    }
#undef CA_FB
    st.n_rhs += n_rhs; st.n_acc += n_acc; st.n_rej += n_rej; st.n_exact = 0; st.status = status;
}

}  // namespace ca
