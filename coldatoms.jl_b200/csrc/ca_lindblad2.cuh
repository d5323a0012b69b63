// ca_lindblad2.cuh -- per-lane code of the two-atom CZ kernel K4 (simulation_czlp, src/cz_model.jl:107-171).
//
// One trajectory per WARP, written as a template over a "warp" type W that supplies the lane id, the warp's
// shared-memory areas, sync() and warp sums: the sm_100a kernel (ca_lindblad2.cu) instantiates it with
// __syncwarp / shuffles, tests/hostemu instantiates the SAME source with 32 host threads and a barrier.
//
// State: for ψ0 without |l> components ρ is 16x16 (levels 0,1,r,p of both atoms) ⊕ 4x4 (atom 1
// lost, L1) ⊕ 4x4 (atom 2 lost, L2) ⊕ ρ_ll,ll.  Every Hermitian block is stored as a REAL matrix M of the same shape,
//     M[A][B] = Re ρ_AB + Im ρ_AB      (so  2 ρ_AB = (M[A][B] + M[B][A]) + i (M[A][B] - M[B][A])  for every A, B:
//                                       no case distinction between the triangles and the diagonal),
// i.e. 256 + 16 + 16 = 288 reals = 9 per lane; ρ_ll,ll follows from the conserved trace.  Lane (B, h) owns rows
// A = 8h .. 8h+7 of column B = (b1, b2) of the 16x16 block (A = a1 + 4 a2) and one element of L1 (lanes 0-15) or L2.
//
// RHS: with G = Hρ (column B of G only needs column B of ρ because H acts on the row index),
//     dρ = -i (G - G†) + dissipator,      H = H1 ⊗ 1 + 1 ⊗ H2 + V |rr><rr|   (cz_model.jl:31-65).
// A lane fetches the transposed reals M[B][A'] of its half column (one exchange through shared memory, conflict-free: rows
// are padded to 18 doubles) and applies H with the SAME register-resident coefficients in every lane (H1 couples rows inside
// a group of four, H2 couples the two groups and the sibling lane).  The complex product is never formed: a coupling
// c = H_AA' maps the source reals (m, mt) = (M[A'][B], M[B][A']) onto S = Im G + Re G += Re c m + Im c mt and
// D = Im G - Re G += Im c m - Re c mt (four FMAs), and the lane needs exactly one real, S_BA, from the transposed owner
// (second exchange):
//     dM[A][B] = D_AB + S_BA - (γ_A + γ_B)/2 M[A][B] + feeding.   No gather tables, no per-lane
// coefficient loads; the only irregular accesses are the few population-feeding terms of the jump operators
// (JumpOperatorsTwo, cz_model.jl:2-9).
#pragma once
#include "ca_device.cuh"

namespace ca {

constexpr int kS2 = 9;             // reals per lane
constexpr int kRS = 18;            // row stride of the 16x16 exchange matrices (144 B: 8 lanes x LDS.128 hit 8 bank groups)
#ifndef CA_XSKEW
#define CA_XSKEW 8                 // rows 8..15 are shifted by 8 doubles: the two half-column lanes (B,0) and (B,1) of a column, which publish rows
#endif                             // j and 8+j at once, then fall into different banks (8*18 = 144 = 0 mod 16 collided 2-way on every publish)
constexpr int kXL = 16 * kRS + CA_XSKEW;   // offset of the 32 l-block reals in an exchange buffer
constexpr int kXN = kXL + 32;      // doubles per exchange buffer
#ifndef CA_RK_SMEM
#define CA_RK_SMEM 0   // 1: y and every stage derivative live in shared memory between the RHS evaluations (registers are left to the RHS)
#endif
#if CA_RK_SMEM
enum { KB_3 = 0, KB_4 = 1, KB_5 = 2, KB_6 = 3, KB_YT = 4, KB_Y = 5, KB_1 = 6, KB_2 = 7, kKB = 8 };   // per-lane buffers in shared memory
#else
#ifndef CA_Y_SMEM
#define CA_Y_SMEM 0    // 1: y (start of the step) lives in shared memory between its uses (stage combinations, norm, dense output)
#endif
#ifndef CA_K1_SMEM
#define CA_K1_SMEM 0   // 1: the same for k1
#endif
enum { KB_3 = 0, KB_4 = 1, KB_5 = 2, KB_6 = 3, KB_YT = 4, KB_Y = 5, KB_1 = 6, kKB = 5 + 2 * (CA_Y_SMEM || CA_K1_SMEM) };   // per-lane buffers in shared memory
#endif
constexpr int kStages2 = 5;        // distinct stage times c2,c3,c4,c5,c6(=c7)
constexpr int kAccSlots = 10;      // accumulator slots per lane: 9 + ρ_ll,ll (lane 0)
// per-lane anchor state of the coefficient evaluation, in shared memory as an[slot * 32] (see eval_coefficients2)
enum { AN_CR = 0, AN_CI, AN_B1R, AN_B1I, AN_B2R, AN_B2I, AN_PHI, AN_TA, AN_TEND, AN_V0, AN_V1, AN_V2, AN_J, kAN };
#ifndef CA_ANCHOR2
#define CA_ANCHOR2 1   // 0: exact exp + sincos on every step (the round-1 behaviour)
#endif
constexpr int kAnchorEvery2 = 32;  // attempts between warp-wide re-anchorings (power of two)
constexpr int kWarpDoubles2 = 2 * kXN + kKB * kS2 * 32 + kStages2 * 16 + 12 + kAN * 32;

// Hamiltonian coefficients at one stage time: Ωr/2 (σ1p), Ωb/2 (σpr) of both atoms (with phase noise and the phase jump),
// detunings (H_pp = -dp, H_rr = -dr) and the blockade shift V.
struct Coef2 { double ar1, ai1, br1, bi1, ar2, ai2, br2, bi2, dp1, dr1, dp2, dr2, V, pad0, pad1, pad2; };

CA_HD int idx5(int A) { return (A & 3) + 5 * (A >> 2); }   // 16-block index a1 + 4 a2 -> reference index a1 + 5 a2
CA_HD int xoff(int A, int B) { return A * kRS + B + (A >= 8 ? CA_XSKEW : 0); }   // element (A, B) of a 16x16 exchange matrix

// (lane, slot) holds M[I][J] = Re ρ_IJ + Im ρ_IJ of the 25x25 matrix; returns the composite indices (reference order a1 + 5 a2)
// and the (lane, slot) of the transposed entry.
CA_HD void slot_entry(int lane, int slot, int* I, int* J, int* tlane, int* tslot) {
    if (slot < 8) {
        const int B = lane >> 1, A = 8 * (lane & 1) + slot;
        *I = idx5(A); *J = idx5(B);
        *tlane = (A << 1) | (B >> 3); *tslot = B & 7;
    } else {
        const int blk = lane >> 4, a = (lane & 15) >> 2, b = lane & 3;
        *I = blk == 0 ? 4 + 5 * a : a + 20;
        *J = blk == 0 ? 4 + 5 * b : b + 20;
        *tlane = blk * 16 + 4 * b + a; *tslot = 8;
    }
}

// host: ρ0 = |ψ0><ψ0| in the per-lane real layout [32][kS2]; false if ψ0 has |l> components (outside the block structure)
inline bool rho0_lanes(const double* psi0, double* rho0m, double* tr0) {
    for (int i = 0; i < 25; ++i)
        if ((i % 5 == 4 || i / 5 == 4) && (psi0[2 * i] != 0.0 || psi0[2 * i + 1] != 0.0)) return false;
    double tr = 0.0;
    for (int i = 0; i < 25; ++i) tr += psi0[2 * i] * psi0[2 * i] + psi0[2 * i + 1] * psi0[2 * i + 1];
    *tr0 = tr;
    for (int l = 0; l < 32; ++l)
        for (int s = 0; s < kS2; ++s) {
            int I, J, tl, ts;
            slot_entry(l, s, &I, &J, &tl, &ts);
            const double ar = psi0[2 * I], ai = psi0[2 * I + 1], br = psi0[2 * J], bi = psi0[2 * J + 1];   // ρ_IJ = ψ_I conj(ψ_J)
            rho0m[l * kS2 + s] = (ar * br + ai * bi) + (ai * br - ar * bi);
        }
    return true;
}

// full complex entry from the sums of a slot and of its transposed slot: out[I + 25 J] (column-major 25x25)
CA_HD void scatter_entry(int lane, int slot, double v, double vt, double* out) {
    if (slot == 9) { if (lane == 0) out[2 * (24 + 25 * 24)] = v; return; }
    int I, J, tl, ts;
    slot_entry(lane, slot, &I, &J, &tl, &ts);
    out[2 * (I + 25 * J)] = 0.5 * (v + vt);
    out[2 * (I + 25 * J) + 1] = I == J ? 0.0 : 0.5 * (v - vt);
}

// Step-size control of the two-atom kernels.  The one-atom kernel sizes its steps with single-precision log2 / exp2 and MUFU-seeded
// reciprocals in the error norm (a step there is ~3 k cycles and the controller sat on its critical path).  A two-atom step is ~14 k
// cycles, and its ensembles reject steps 30 times per trajectory, so EEst sits next to 1 often: with the approximate controller one
// trajectory in six took a different accept / reject sequence than the reference algorithm (a different, equally valid DP5 solution,
// but 1e-5 away in the coherences of an undamped thermal ensemble).  Here the norm uses once-refined reciprocals / square roots (1e-11)
// and the controller the double-precision powers of OrdinaryDiffEq's PI controller (see oracle): the step sequence is the oracle's to
// rounding (~+3 % run time with in-house double log / exp; the library pow / sqrt / division cost 25 %).
#ifndef CA_CTRL_PRECISE2
#define CA_CTRL_PRECISE2 1
#endif
struct StepCtrl2 {
#if CA_CTRL_PRECISE2
    double l_qold;   // ln(qold)
    CA_HD void reset() { l_qold = -9.2103403719761827361; }   // ln(1e-4)
    // accept: q = EEst^0.17 qold^-0.04 / 0.9 in [0.1, 5], dt = h / q, qold <- max(EEst, 1e-4); reject: dt = h / min(5, EEst^0.17 / 0.9)
    // (DP5's PIController in OrdinaryDiffEq, see oracle) with in-house double log / exp (~1 ulp, no slow paths)
    CA_HD double decide(double EEst2, bool acc) {   // returns dt / h
        const double lE = EEst2 > 1e-300 ? 0.5 * log_d(EEst2) : -350.0;                       // ln(EEst); EEst = 0 saturates the clamps
        const double base = fma(0.17, lE, 0.10536051565782630123);                            // - ln(0.9)
        const double la = fmax(-2.3025850929940456840, fmin(1.6094379124341003746, fma(-0.04, l_qold, base)));   // ln q in [ln 0.1, ln 5]
        const double lr = fmin(1.6094379124341003746, base);
        l_qold = acc ? fmax(lE, -9.2103403719761827361) : l_qold;
        return exp_d(-(acc ? la : lr));
    }
    static CA_HD double rcpn(double x) { return rcp_r1(x); }    // (1e-11: a decision could only flip for |EEst² - 1| < 1e-11)
    static CA_HD double sqrtn(double x) { return sqrt_r1(x); }
#else
    StepCtrl c;
    CA_HD void reset() { c.reset(); }
    CA_HD double decide(double EEst2, bool acc) { return c.decide(EEst2, acc); }
    static CA_HD double rcpn(double x) { return rcp_norm(x); }
    static CA_HD double sqrtn(double x) { return sqrt_norm(x); }
#endif
};

// Per-lane static data.
struct Lane2 {
    int lane, h;
    unsigned dg;                // bit j (0..8): own element j is a diagonal entry
    int a8, blk;
    int o_pub, o_row, o_sib, o_sibrow, o_f1lo, o_f1hi, o_f2;
    int o8_t, o8_n1, o8_n1t, o8_n2, o8_n2t, o8_fp, o8_fr, o8_pp;
    double dlo, dhi, dec8;      // (γ_A + γ_B)/2 without the a1 part
    double w0, w1, u0, u1, w8;  // population-feeding weights

    CA_HD void init(int ln, const DevModel& m) {
        lane = ln; h = ln & 1;
        const int B = ln >> 1, b1 = B & 3, b2 = B >> 2;
        const double Gp = m.G0 + m.G1 + m.Gl;
        const double gam[4] = {0.0, 0.0, m.Gr, Gp};
        dg = 0;
        for (int j = 0; j < 8; ++j) if (8 * h + j == B) dg |= 1u << j;
        o_pub = xoff(8 * h, B);                  // own rows 8h + j: o_pub + j * kRS
        o_row = xoff(B, 8 * h);                  // row B, columns 8h .. 8h+7 (the transposed partners)
        o_sib = xoff(8 * (1 - h) + 4, B);        // the sibling lane's rows 8(1-h)+4 + i
        o_sibrow = xoff(B, 8 * (1 - h) + 4);
        // jumps of atom 1 (σ0p, σ1p): rows with a1 = b1 in {0,1} are fed from M[(p,a2)][(p,b2)]
        o_f1lo = xoff(3 + 4 * (2 * h), 3 + 4 * b2);
        o_f1hi = xoff(3 + 4 * (2 * h + 1), 3 + 4 * b2);
        w0 = b1 == 0 ? m.G0 : 0.0; w1 = b1 == 1 ? m.G1 : 0.0;
        // jumps of atom 2: rows with a2 = b2 in {0,1} (only the h = 0 half has them) are fed from M[(a1,p)][(b1,p)]
        o_f2 = xoff(12, b1 + 12);
        u0 = (h == 0 && b2 == 0) ? m.G0 : 0.0; u1 = (h == 0 && b2 == 1) ? m.G1 : 0.0;
        dlo = 0.5 * (gam[b1] + gam[b2] + gam[2 * h]);
        dhi = 0.5 * (gam[b1] + gam[b2] + gam[2 * h + 1]);
        // l-blocks: lane = blk*16 + 4 a + b holds element (a, b) of L1 (atom 1 lost; indices of atom 2) or L2
        blk = ln >> 4; a8 = (ln & 15) >> 2;
        const int b8 = ln & 3, base = kXL + blk * 16;
        if (a8 == b8) dg |= 1u << 8;
        o8_t = base + 4 * b8 + a8;
        const int n1 = a8 == 3 ? 1 : 3, n2 = 2;
        o8_n1 = base + 4 * n1 + b8; o8_n1t = base + 4 * b8 + n1;
        o8_n2 = base + 4 * n2 + b8; o8_n2t = base + 4 * b8 + n2;
        if (blk == 0) { o8_fp = xoff(3 + 4 * a8, 3 + 4 * b8); o8_fr = xoff(2 + 4 * a8, 2 + 4 * b8); }
        else { o8_fp = xoff(a8 + 12, b8 + 12); o8_fr = xoff(a8 + 8, b8 + 8); }
        o8_pp = base + 15;
        w8 = (a8 == b8 && a8 == 0) ? m.G0 : ((a8 == b8 && a8 == 1) ? m.G1 : 0.0);
        dec8 = 0.5 * (gam[a8] + gam[b8]);
    }
    CA_HD bool is_eq(int j) const { return (dg >> j) & 1u; }
};

CA_HD void ld2(const double* p, double& a, double& b) {
#if defined(__CUDA_ARCH__)
    const double2 v = *reinterpret_cast<const double2*>(p);
    a = v.x; b = v.y;
#else
    a = p[0]; b = p[1];
#endif
}

// G += c x (complex)
#define CA_CMAC(gr, gi, cr, ci, xr_, xi_)        \
    do {                                         \
        gr = fma((cr), (xr_), gr);               \
        gr = fma(-(ci), (xi_), gr);              \
        gi = fma((cr), (xi_), gi);               \
        gi = fma((ci), (xr_), gi);               \
    } while (0)

// TWICE the complex entry of a row from its own real `mo` and the transposed real `mt`
CA_HD void cplx2(double mo, double mt, double& re, double& im) { re = mo + mt; im = mo - mt; }

// One complex coupling c = cr + i ci acting on a source row given by its two reals (m = M[A'][B], mt = M[B][A']):
// with G = H ρ and 2ρ_A'B = (m + mt) + i (m - mt),   S = Im G + Re G  gets  cr m + ci mt,   D = Im G - Re G  gets  ci m - cr mt.
#define CA_MMAC(S_, D_, cr, ci, m_, mt_)         \
    do {                                         \
        S_ = fma((cr), (m_), S_);                \
        S_ = fma((ci), (mt_), S_);               \
        D_ = fma((ci), (m_), D_);                \
        D_ = fma(-(cr), (mt_), D_);              \
    } while (0)

// dM/dt of this lane's nine reals.  The complex product G = Hρ is never formed: in the M representation a coupling
// H_AA' = c maps the source reals (M[A'][B], M[B][A']) straight onto S = Im G + Re G (published for the transposed owner)
// and D = Im G - Re G (kept), four FMAs per coupling:  dM[A][B] = D_AB + S_BA - (γ_A + γ_B)/2 M[A][B] + feeding.
// tp[j] returns the transposed partner value: |ρ_AB|² = (u² + tp²)/2 (also on the diagonal, where tp = u).
template <class W>
CA_HD void rhs2(W& w, const Lane2& L, const DevModel& m, const Coef2& cf, const double (&u)[kS2], double (&du)[kS2], double (&tp)[kS2]) {
    double* xm = w.xm;
    double* xp = w.xp;
    // ---- A: publish M
#pragma unroll
    for (int j = 0; j < 8; ++j) xm[L.o_pub + j * kRS] = u[j];
    xm[kXL + L.lane] = u[8];
    w.sync();
    // ---- B: transposed reals of the own half column and of the sibling rows the 1 <-> p coupling of atom 2 needs
    double sm_[4], st[4];
    {
#pragma unroll
        for (int q = 0; q < 4; ++q) ld2(xm + L.o_row + 2 * q, tp[2 * q], tp[2 * q + 1]);
        ld2(xm + L.o_sibrow, st[0], st[1]);
        ld2(xm + L.o_sibrow + 2, st[2], st[3]);
#pragma unroll
        for (int i = 0; i < 4; ++i) sm_[i] = xm[L.o_sib + i * kRS];
    }
    double S[8], D[8];
    {
        const double e2lo = L.h ? -cf.dr2 : 0.0, e2hi = L.h ? -cf.dp2 : 0.0, vrr = L.h ? cf.V : 0.0;
        const double e1[4] = {0.0, 0.0, -cf.dr1, -cf.dp1};
#pragma unroll
        for (int j = 0; j < 8; ++j) {   // real diagonal of H: S += e m, D -= e mt
            double e = e1[j & 3] + (j < 4 ? e2lo : e2hi);
            if (j == 2) e += vrr;   // |rr><rr| (get_V, cz_model.jl:12-17,63)
            S[j] = e * u[j]; D[j] = -e * tp[j];
        }
        // H1 ⊗ 1: rows (1, r, p) of each group of four
#pragma unroll
        for (int g = 0; g < 2; ++g) {
            const int j0 = 4 * g;
            CA_MMAC(S[j0 + 1], D[j0 + 1], cf.ar1, cf.ai1, u[j0 + 3], tp[j0 + 3]);     // H[1,p] = Ωr/2
            CA_MMAC(S[j0 + 2], D[j0 + 2], cf.br1, -cf.bi1, u[j0 + 3], tp[j0 + 3]);    // H[r,p] = (Ωb/2)*
            CA_MMAC(S[j0 + 3], D[j0 + 3], cf.ar1, -cf.ai1, u[j0 + 1], tp[j0 + 1]);    // H[p,1] = (Ωr/2)*
            CA_MMAC(S[j0 + 3], D[j0 + 3], cf.br1, cf.bi1, u[j0 + 2], tp[j0 + 2]);     // H[p,r] = Ωb/2
        }
        // 1 ⊗ H2: h = 0 holds a2 = (0, 1), h = 1 holds a2 = (r, p); the 1 <-> p coupling crosses to the sibling lane
        const double k01r = L.h ? cf.br2 : 0.0, k01i = L.h ? -cf.bi2 : 0.0;   // row r <- (Ωb/2)* p
        const double k10r = L.h ? cf.br2 : 0.0, k10i = L.h ? cf.bi2 : 0.0;    // row p <- Ωb/2 r
        const double ksr = cf.ar2, ksi = L.h ? -cf.ai2 : cf.ai2;              // row p <- (Ωr/2)* 1 ; row 1 <- Ωr/2 p
#pragma unroll
        for (int a = 0; a < 4; ++a) {
            CA_MMAC(S[a], D[a], k01r, k01i, u[4 + a], tp[4 + a]);
            CA_MMAC(S[4 + a], D[4 + a], ksr, ksi, sm_[a], st[a]);
            CA_MMAC(S[4 + a], D[4 + a], k10r, k10i, u[a], tp[a]);
        }
    }
    // l-block element (a8, b): H_s L with H_s the Hamiltonian of the surviving atom
    double S8, D8;
    {
        tp[8] = xm[L.o8_t];
        const double alr = L.blk ? cf.ar1 : cf.ar2, ali = L.blk ? cf.ai1 : cf.ai2, ber = L.blk ? cf.br1 : cf.br2, bei = L.blk ? cf.bi1 : cf.bi2;
        const double dpx = L.blk ? cf.dp1 : cf.dp2, drx = L.blk ? cf.dr1 : cf.dr2;
        // a = 1: Ωr/2 L_p ; a = r: (Ωb/2)* L_p - dr L_r ; a = p: (Ωr/2)* L_1 + Ωb/2 L_r - dp L_p
        const double c1r = L.a8 == 0 ? 0.0 : (L.a8 == 2 ? ber : alr);
        const double c1i = L.a8 == 0 ? 0.0 : (L.a8 == 1 ? ali : (L.a8 == 2 ? -bei : -ali));
        const double c2r = L.a8 == 3 ? ber : 0.0, c2i = L.a8 == 3 ? bei : 0.0;
        const double e = L.a8 == 2 ? -drx : (L.a8 == 3 ? -dpx : 0.0);
        S8 = e * u[8]; D8 = -e * tp[8];
        CA_MMAC(S8, D8, c1r, c1i, xm[L.o8_n1], xm[L.o8_n1t]);
        CA_MMAC(S8, D8, c2r, c2i, xm[L.o8_n2], xm[L.o8_n2t]);
    }
    // publish S for the transposed owner, keep D; decay and population feeding
    {
        const double ga[4] = {0.0, 0.0, 0.5 * m.Gr, 0.5 * (m.G0 + m.G1 + m.Gl)};
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            xp[L.o_pub + j * kRS] = S[j];
            du[j] = fma(-((j < 4 ? L.dlo : L.dhi) + ga[j & 3]), u[j], D[j]);
        }
        xp[kXL + L.lane] = S8;
        const double vlo = xm[L.o_f1lo], vhi = xm[L.o_f1hi];
        du[0] = fma(L.w0, vlo, du[0]); du[1] = fma(L.w1, vlo, du[1]);
        du[4] = fma(L.w0, vhi, du[4]); du[5] = fma(L.w1, vhi, du[5]);
#pragma unroll
        for (int a = 0; a < 4; ++a) {
            const double v = xm[L.o_f2 + a * kRS];
            du[a] = fma(L.u0, v, du[a]); du[4 + a] = fma(L.u1, v, du[4 + a]);
        }
        double d8 = fma(-L.dec8, u[8], D8);
        d8 = fma(m.Gl, xm[L.o8_fp], d8);
        d8 = fma(m.Gr, xm[L.o8_fr], d8);
        du[8] = fma(L.w8, xm[L.o8_pp], d8);
    }
    w.sync();
    // ---- C: add the transposed owner's S
    {
        double q[8];
#pragma unroll
        for (int k = 0; k < 4; ++k) ld2(xp + L.o_row + 2 * k, q[2 * k], q[2 * k + 1]);
#pragma unroll
        for (int j = 0; j < 8; ++j) du[j] += q[j];
        du[8] += xp[L.o8_t];
    }
}

// Coefficients of the five distinct stage times of one step, one (quantity, stage) pair per lane:
// q = 0..3 -> Ω/2 of red1, blue1, red2, blue2 (own phase-noise traces, phase jump ξ on the red beams: cz_model.jl:37-65,
// 136-137), q = 4 -> detunings and V(t) = c6 / (1e-18 + R^6) (cz_model.jl:12-17).
//
// Gaussian beams are ANCHORED like in the one-atom kernel (ca_device.cuh, "anchored evaluation"): every kAnchorEvery2 attempts all
// lanes rebuild, at the step's start time t_a, the exact coefficient c(t_a) = (Ω/2) B(r(t_a)) e^{iφ_a} (φ_a = the lane's own noise
// phase of that step) and a quadratic model of ln B along the atom's path on [t_a, t_a + 2H]; in between
//     c(t) = c(t_a) exp(τ (b1 + τ b2) + i (φ(t) - φ_a)),   τ = t - t_a,
// costs three short real series instead of a full exp + sincos (a ~100-deep dependent chain that every step of the warp waited
// for).  Guards as in the one-atom kernel (|Re δ| < 3e-4 bounded per anchor, |Im δ| < 6e-3 per evaluation); a lane whose guard
// fails evaluates exactly and re-anchors at the next step.  The lane keeps a three-node window of its phase-noise trace in shared
// memory (the table itself lives in global scratch): the common case touches no global memory.
CA_HD void atom_position2(const DevModel& m, const double* s, double tt, double* pos) {
    if (m.free_motion) { pos[0] = s[0] + s[3] * tt; pos[1] = s[1] + s[4] * tt; pos[2] = s[2] + s[5] * tt; }
    else {  // R: atom_sampler.jl:125-133 (oscillation about the origin although centres are shifted, App. B6)
        double sr, cr, sz, cz;
        sincos_fast(m.wr * tt, &sr, &cr); sincos_fast(m.wz * tt, &sz, &cz);
        pos[0] = s[0] * cr + s[3] / m.wr * sr; pos[1] = s[1] * cr + s[4] / m.wr * sr; pos[2] = s[2] * cz + s[5] / m.wz * sz;
    }
}

// Per-lane registers of the coefficient evaluation that persist over the steps of a trajectory.
struct CoefLane2 {
    double pf;    // prefetched table value of node J + 3 (requested one window advance ago: nothing waits for the load)
    double cst;   // c_s of this lane's stage (Dormand-Prince nodes 0.2, 0.3, 0.8, 8/9; the fifth stage sits at t + h itself)
    int J;        // first node of the cached window (J, J+1, J+2 in shared memory: AN_V0..2), -10: empty
    CA_HD void reset(int lane) { pf = 0.0; J = -10; const int si = lane / 5; cst = si < 4 ? MC.rc[si + 2] : 1.0; }
};

// Gridded(Linear) phase of trace `tr` at time tt (cz_model.jl:45-46) through the lane's cached window of three nodes J, J+1, J+2.
// The common cases -- tt still inside [t_J, t_J+1), or one interval further -- touch no global memory and convert no integers: the window
// start time t_J is cached (AN_J), node J+2 is in the window and node J+3 was prefetched into a register when the window last advanced
// (the first version converted tt to an index and back on every call, and a lane that advanced stored a freshly loaded value at
// once: the whole warp waited for that load at the barrier, 1 % of the kernel's samples on one STS).
CA_HD double noise_phase2(const DevModel& m, const double* tr, double* an, CoefLane2& cl, double tt) {
    double tJ = an[AN_J * 32], v0 = an[AN_V0 * 32], v1 = an[AN_V1 * 32];
    const double rel = tt - tJ;
    if (!(cl.J >= 0 && rel >= 0.0 && rel < m.dtn)) {
        const int last = m.n_nodes - 1;
        if (cl.J >= 0 && rel >= m.dtn && rel < 2.0 * m.dtn && cl.J + 1 <= m.n_nodes - 2) {   // the next interval
            cl.J += 1;
            v0 = v1; v1 = an[AN_V2 * 32];
            an[AN_V2 * 32] = cl.pf;
        } else {   // first use, a jump (rejected step), or the end of the table
            int j = (int)(tt * m.inv_dtn);
            j = j < 0 ? 0 : (j > m.n_nodes - 2 ? m.n_nodes - 2 : j);
            cl.J = j;
            v0 = tr[j]; v1 = tr[j + 1];
            an[AN_V2 * 32] = tr[j + 2 < last ? j + 2 : last];
        }
        cl.pf = tr[cl.J + 3 < last ? cl.J + 3 : last];
        tJ = cl.J * m.dtn;
        an[AN_V0 * 32] = v0; an[AN_V1 * 32] = v1; an[AN_J * 32] = tJ;
    }
    const double wgt = (tt - tJ) * m.inv_dtn;
    return v0 + wgt * (v1 - v0);
}

template <class W>
CA_HD void eval_coefficients2(W& w, const DevModel& m, const double* tab, CoefLane2& cl, double t0, double hh, double tn, bool refresh, int& n_exact) {
    const int q = w.lane % 5, si = w.lane / 5;
    if (si < kStages2) {
        const double tt = si == kStages2 - 1 ? tn : t0 + cl.cst * hh;
        double* o = reinterpret_cast<double*>(w.coef + si);
        if (q < 4) {
            const int a = q >> 1;
            const double* s = w.smp + 6 * a;
            double* an = w.anc;
            double ph = 0.0;
            if (tab) ph = noise_phase2(m, tab + (size_t)q * m.n_nodes, an, cl, tt);
            if (!(q & 1) && tt >= m.tau) ph += m.xi;  // ϕr(t) = t < τ ? 0 : ξ on both red beams (cz_model.jl:136-137,48)
            double re, im;
            const DevBeam& bm = (q & 1) ? m.blue : m.red;
            bool done = false;
            if (CA_ANCHOR2 && bm.kind == CA_BEAM_GAUSS && hh > 0.0) {
                double t_a = an[AN_TA * 32], t_end = an[AN_TEND * 32];
                if (refresh || (tt > t_end && t_end > -1e299) || t0 < t_a) {   // rebuild the anchor at the start of this step
                    const double H = (0.5 * kAnchorEvery2 + 1.0) * hh, iH = rcp_d(H);
                    double L[3][2], pos[3];
#pragma unroll 1
                    for (int i = 0; i < 3; ++i) {
                        atom_position2(m, s, fma((double)i, H, t0), pos);
                        const double rp = sqrt_d(fmax(pos[0] * pos[0] + pos[1] * pos[1], 1e-300));
                        gauss_log(bm, rp, pos[2], &L[i][0], &L[i][1]);
                        if (i == 0) {
                            double anc[5], cr, ci;
                            gauss_exact(bm, rp, pos[2], ph, &cr, &ci, anc);
                            an[AN_CR * 32] = cr; an[AN_CI * 32] = ci;
                        }
                    }
                    bool ok = true;
#pragma unroll
                    for (int c = 0; c < 2; ++c) {   // ln B(t_a + τ) - ln B(t_a) = τ (b1 + τ b2)
                        const double d1 = L[1][c] - L[0][c], d2 = L[2][c] - L[1][c];
                        const double b1 = (1.5 * d1 - 0.5 * d2) * iH, b2 = 0.5 * (d2 - d1) * iH * iH;
                        an[(AN_B1R + c) * 32] = b1; an[(AN_B2R + c) * 32] = b2;
                        ok = ok && fabs(d2 - d1) < 1e-6;
                        if (c == 0) ok = ok && 2.0 * H * (fabs(b1) + 2.0 * H * fabs(b2)) < 3e-4;
                    }
                    t_a = t0; t_end = ok ? fma(2.0, H, t0) : -1e300;   // (-1e300: no model until the next warp-wide refresh)
                    an[AN_PHI * 32] = ph; an[AN_TA * 32] = t_a; an[AN_TEND * 32] = t_end;
                }
                if (tt <= t_end) {
                    const double tau = tt - t_a;
                    const double dr = tau * fma(tau, an[AN_B2R * 32], an[AN_B1R * 32]);
                    const double di = fma(tau, fma(tau, an[AN_B2I * 32], an[AN_B1I * 32]), ph - an[AN_PHI * 32]);
                    const double d2 = di * di;
                    if (d2 < 3.6e-5) {   // |di| < 6e-3: di^6/720 < 7e-17; |dr| < 3e-4 (bounded at the anchor): dr^4/24 < 4e-16
                        const double ed = fma(dr, fma(dr, fma(dr, 1.0 / 6.0, 0.5), 1.0), 1.0);
                        const double cs = fma(d2, fma(d2, 1.0 / 24.0, -0.5), 1.0);
                        const double sn = di * fma(d2, fma(d2, 1.0 / 120.0, -1.0 / 6.0), 1.0);
                        const double er = ed * cs, ei = ed * sn, car = an[AN_CR * 32], cai = an[AN_CI * 32];
                        re = car * er - cai * ei; im = car * ei + cai * er;
                        done = true;
                    } else an[AN_TEND * 32] = t0;   // the phase ran away (noise, or the ξ jump): re-anchor at the next step
                }
            }
            if (!done) {
                double pos[3];
                atom_position2(m, s, tt, pos);
                if (bm.kind == CA_BEAM_GAUSS) {
                    double anc[5];
                    const double rp = sqrt_d(fmax(pos[0] * pos[0] + pos[1] * pos[1], 1e-300));
                    gauss_exact(bm, rp, pos[2], ph, &re, &im, anc);
                } else beam_half_omega(bm, pos[0], pos[1], pos[2], ph, &re, &im);
                if (hh > 0.0) ++n_exact;
            }
            o[2 * q] = re; o[2 * q + 1] = im;
        } else {
            double pos[2][3], vel[2][2];
#pragma unroll
            for (int a = 0; a < 2; ++a) {
                const double* s = w.smp + 6 * a;
                if (m.free_motion) {
                    pos[a][0] = s[0] + s[3] * tt; pos[a][1] = s[1] + s[4] * tt; pos[a][2] = s[2] + s[5] * tt;
                    vel[a][0] = s[3]; vel[a][1] = m.vz_from_vy ? s[4] : s[5];
                } else {  // R, V: atom_sampler.jl:125-143 (oscillation about the origin although centres are shifted, App. B6)
                    double sr, cr, sz, cz;
                    sincos_fast(m.wr * tt, &sr, &cr); sincos_fast(m.wz * tt, &sz, &cz);
                    pos[a][0] = s[0] * cr + s[3] / m.wr * sr; pos[a][1] = s[1] * cr + s[4] / m.wr * sr; pos[a][2] = s[2] * cz + s[5] / m.wz * sz;
                    vel[a][0] = s[3] * cr - s[0] * m.wr * sr;
                    vel[a][1] = (m.vz_from_vy ? s[4] : s[5]) * cz - s[2] * m.wz * sz;
                }
            }
#pragma unroll
            for (int a = 0; a < 2; ++a) {
                const double Dd = m.aDx * vel[a][0] + m.aDz * vel[a][1], dd = m.bDx * vel[a][0] + m.bDz * vel[a][1];
                o[8 + 2 * a] = m.Delta0 - m.dsign * Dd;      // dp: H_pp = -dp
                o[9 + 2 * a] = m.delta0 - m.dsign * dd;      // dr: H_rr = -dr
            }
            const double dx = pos[0][0] - pos[1][0], dy = pos[0][1] - pos[1][1], dz = pos[0][2] - pos[1][2];
            const double r2 = dx * dx + dy * dy + dz * dz;
            o[12] = m.c6 * rcp_d(1e-18 + r2 * r2 * r2);      // get_V: cz_model.jl:12-17
        }
    }
    w.sync();
}

// The kernel is bound by shared-memory wavefronts, so k3 and k5 stay in registers for their later uses (switches kept for
// experiments); their shared-memory copies (and k6's) exist only for Hairer's dense output and are written only by a step that
// overshoots the next output time.
#ifndef CA_DENSE_ALWAYS
#define CA_DENSE_ALWAYS 0   // 1: k3, k5, k6 are stored for the dense output on every step (no warp-uniform branches around the stores)
#endif
#ifndef CA_K3_SMEM
#define CA_K3_SMEM 0
#endif
#ifndef CA_K5_SMEM
#define CA_K5_SMEM 0
#endif
#if CA_K3_SMEM
#define CA_K3L(s) kb[(KB_3 * kS2 + (s)) * 32]
#else
#define CA_K3L(s) k3[s]
#endif
#if CA_K5_SMEM
#define CA_K5L(s) kb[(KB_5 * kS2 + (s)) * 32]
#else
#define CA_K5L(s) k5[s]
#endif

#ifdef CA_TIMERS
#define CA_TICK(w, i) (w).tick(i)
#else
#define CA_TICK(w, i)
#endif

struct Stats2 { long long n_rhs, n_acc, n_rej, n_exact; int status; };

// Adds the state o (this lane's nine reals at one output time) and its second moment to the per-warp accumulators a0.
template <class W>
CA_HD void accumulate_output2(W& w, const Lane2& L, const DevModel& m, const double (&o)[kS2], double tr0, int status, double* a0) {
    auto diag_sum = [&](const double (&v)[kS2]) {
        double s = 0.0;
#pragma unroll
        for (int j = 0; j < kS2; ++j) s += L.is_eq(j) ? v[j] : 0.0;
        return s;
    };
    // second moment: complex entries through the exchange buffers (xm and xp are contiguous: 16x16 + 2x4x4 complex)
    {
        double* xm = w.xm;
#pragma unroll
        for (int s = 0; s < 8; ++s) xm[L.o_pub + s * kRS] = o[s];
        xm[kXL + L.lane] = o[8];
        w.sync();
        double xr[kS2], xi[kS2];
#pragma unroll
        for (int s = 0; s < 8; ++s) cplx2(o[s], xm[L.o_row + s], xr[s], xi[s]);   // 2ρ
        cplx2(o[8], xm[L.o8_t], xr[8], xi[8]);
        w.sync();
        if (m.rho2_mode == CA_RHO2_MATSQ) {   // (ρ·ρ)_AB = Σ_K ρ_AK ρ_KB, block by block
            const int B = L.lane >> 1;
#pragma unroll
            for (int s = 0; s < 8; ++s) { double* p = xm + ((8 * L.h + s) * 16 + B) * 2; p[0] = xr[s]; p[1] = xi[s]; }
            { double* p = xm + 512 + L.lane * 2; p[0] = xr[8]; p[1] = xi[8]; }
            w.sync();
#pragma unroll 1
            for (int s = 0; s < 8; ++s) {
                double ar = 0.0, ai = 0.0;
                const double* row = xm + (8 * L.h + s) * 32;
                for (int K = 0; K < 16; ++K) {
                    const double cr = row[2 * K], ci = row[2 * K + 1], dr = xm[(K * 16 + B) * 2], di = xm[(K * 16 + B) * 2 + 1];
                    CA_CMAC(ar, ai, cr, ci, dr, di);
                }
                a0[(kAccSlots + s) * 32] += 0.25 * (ar + ai);
            }
            {
                double ar = 0.0, ai = 0.0;
                const int b8 = L.lane & 3;
                const double* blk = xm + 512 + L.blk * 32;
                for (int K = 0; K < 4; ++K) {
                    const double cr = blk[(4 * L.a8 + K) * 2], ci = blk[(4 * L.a8 + K) * 2 + 1], dr = blk[(4 * K + b8) * 2], di = blk[(4 * K + b8) * 2 + 1];
                    CA_CMAC(ar, ai, cr, ci, dr, di);
                }
                a0[(kAccSlots + 8) * 32] += 0.25 * (ar + ai);
            }
            w.sync();
        } else {
#pragma unroll
            for (int s = 0; s < kS2; ++s) a0[(kAccSlots + s) * 32] += 0.25 * (xr[s] * xr[s] + xi[s] * xi[s]);   // |ρ_AB|² (real: Im part 0)
        }
    }
    double o_ll = tr0 - w.sum(diag_sum(o));
    if (status != 0) o_ll = 0.0;
#pragma unroll
    for (int s = 0; s < kS2; ++s) a0[s * 32] += o[s];
    if (L.lane == 0) { a0[9 * 32] += o_ll; a0[(kAccSlots + 9) * 32] += o_ll * o_ll; }
}

#if CA_RK_SMEM
// One trajectory: DP5 (OrdinaryDiffEq semantics, see oracle) over tspan; ρ and ρ·ρ at every tspan point are ADDED to
// acc[((j*2 + which)*kAccSlots + slot)*32] (pointer already offset by the lane).
//
// Register plan: the RHS (two exchanges, ~130 FMAs in 18 accumulator chains) is the part that needs registers to keep enough
// independent FMAs in flight with only two warps per scheduler, so NOTHING of the Runge-Kutta bookkeeping stays live across it:
// y (start of the step), k1..k6 and the transposed partner of y sit in per-lane shared-memory buffers kb[(buf*kS2 + s)*32]
// (conflict-free) and are re-read where a stage combination needs them; only the newest stage derivative is used from registers.
// An accepted step leaves (un, k7, unt) in registers until the top of the next attempt, which commits them to the buffers.
template <class W>
CA_HD void integrate2(W& w, const Lane2& L, const DevModel& m, const OdeOpts& oo, const double* tspan, int nt, const double* tab,
                      const double* rho0, double tr0, double* acc, Stats2& st) {
    const double rtol = oo.rtol, atol = oo.atol;
    double* kb = w.kb;   // kb[(buf*kS2 + s)*32], lane offset included
#define CA_KB(buf, s) kb[((buf) * kS2 + (s)) * 32]
    double un[kS2], k7[kS2], unt[kS2], ut[kS2];
#pragma unroll
    for (int s = 0; s < kS2; ++s) { un[s] = rho0[s]; CA_KB(KB_Y, s) = rho0[s]; }
    const double t0 = tspan[0], tend = tspan[nt - 1];
    double t = t0, h = 0.0, tcommit = t0, dt = oo.dt0;
    bool pending = false;   // (un, k7, unt) hold an accepted step that is not yet in the buffers
    StepCtrl2 ctrl;
    ctrl.reset();
    long long iters = 0;
    int status = 0, n_rhs = 0, n_acc = 0, n_rej = 0, n_exact = 0;
    w.anc[AN_TA * 32] = 1e300; w.anc[AN_TEND * 32] = -1e300; w.anc[AN_J * 32] = 0.0;   // no anchor
    CoefLane2 cl;
    cl.reset(w.lane);   // empty noise window
    auto diag_sum = [&](const double (&v)[kS2]) {
        double s = 0.0;
#pragma unroll
        for (int j = 0; j < kS2; ++j) s += L.is_eq(j) ? v[j] : 0.0;
        return s;
    };
    // k1 = f(t0, y)
    eval_coefficients2(w, m, tab, cl, t0, 0.0, t0, false, n_exact);
    rhs2(w, L, m, w.coef[kStages2 - 1], un, k7, unt);
#pragma unroll
    for (int s = 0; s < kS2; ++s) { CA_KB(KB_YT, s) = unt[s]; CA_KB(KB_1, s) = k7[s]; }
    n_rhs = 1;
    double y_ll = tr0 - w.sum(diag_sum(un)), un_ll = y_ll;
    if (dt <= 0.0) {  // OrdinaryDiffEq initial step (Hairer), norms over all 625 entries
        double d0 = 0.0, d1 = 0.0, isk2[kS2];
#pragma unroll
        for (int s = 0; s < kS2; ++s) {
            const double a2 = 0.5 * (un[s] * un[s] + unt[s] * unt[s]);   // |ρ_AB|²
            const double sk = atol + sqrt(a2) * rtol;
            isk2[s] = 1.0 / (sk * sk);   // the entries (A,B) and (B,A) of the full matrix share |err|² = (e² + e_t²)/2: weight 1 per real
            d0 += isk2[s] * un[s] * un[s];
            d1 += isk2[s] * k7[s] * k7[s];
        }
        const double f0_ll = -w.sum(diag_sum(k7));
        const double sk_ll = atol + fabs(y_ll) * rtol;
        d0 = sqrt((w.sum(d0) + y_ll * y_ll / (sk_ll * sk_ll)) / 625.0);
        d1 = sqrt((w.sum(d1) + f0_ll * f0_ll / (sk_ll * sk_ll)) / 625.0);
        double dt0 = (d0 < 1e-5 || d1 < 1e-5) ? 1e-6 : 0.01 * d0 / d1;
        dt0 = fmin(dt0, oo.dtmax);
#pragma unroll
        for (int s = 0; s < kS2; ++s) ut[s] = un[s] + dt0 * k7[s];
        eval_coefficients2(w, m, tab, cl, t0, 0.0, t0 + dt0, false, n_exact);
        double k2[kS2], dum[kS2];
        rhs2(w, L, m, w.coef[kStages2 - 1], ut, k2, dum);
        n_rhs++;
        double d2 = 0.0;
#pragma unroll
        for (int s = 0; s < kS2; ++s) { const double a = k2[s] - k7[s]; d2 += isk2[s] * a * a; }
        const double f1_ll = -w.sum(diag_sum(k2));
        d2 = sqrt((w.sum(d2) + (f1_ll - f0_ll) * (f1_ll - f0_ll) / (sk_ll * sk_ll)) / 625.0) / dt0;
        const double mx = fmax(d1, d2);
        const double dt1 = (mx <= 1e-15) ? fmax(1e-6, dt0 * 1e-3) : pow(10.0, -(2.0 + log10(mx)) / 5.0);
        dt = fmin(100.0 * dt0, fmin(dt1, oo.dtmax));
    }
    for (int j = 0; j < nt; ++j) {
        const double T = tspan[j];
        const double tstop = oo.dense ? tend : T;
        while (status == 0 && tcommit < T) {
            if (pending) {  // commit the accepted step (FSAL): y <- un, k1 <- k7 (the partner of y was stored at the accept)
#pragma unroll
                for (int s = 0; s < kS2; ++s) { CA_KB(KB_Y, s) = un[s]; CA_KB(KB_1, s) = k7[s]; }
                y_ll = un_ll;
                t = tcommit; h = 0.0; pending = false;
            }
            if (++iters > oo.maxiters) { status = CA_ERR_MAXITERS; break; }
            double hh = fmin(dt, oo.dtmax);
            bool clipped = false;
            if (t + hh >= tstop || (tstop - (t + hh)) < 1e-13 * fabs(tstop)) { hh = tstop - t; clipped = true; }
            const double tn = clipped ? tstop : t + hh;
            CA_TICK(w, 3);
            eval_coefficients2(w, m, tab, cl, t, hh, tn, ((iters - 1) & (kAnchorEvery2 - 1)) == 0, n_exact);
            CA_TICK(w, 1);
            double kk[kS2], dum[kS2];
            // stage 2
#pragma unroll
            for (int s = 0; s < kS2; ++s) ut[s] = fma(hh * MC.a21, CA_KB(KB_1, s), CA_KB(KB_Y, s));
            rhs2(w, L, m, w.coef[0], ut, kk, dum);   // k2
#pragma unroll
            for (int s = 0; s < kS2; ++s) {
                CA_KB(KB_2, s) = kk[s];
                ut[s] = fma(hh, fma(MC.a32, kk[s], MC.a31 * CA_KB(KB_1, s)), CA_KB(KB_Y, s));
            }
            rhs2(w, L, m, w.coef[1], ut, kk, dum);   // k3
#pragma unroll
            for (int s = 0; s < kS2; ++s) {
                CA_KB(KB_3, s) = kk[s];
                ut[s] = fma(hh, fma(MC.a43, kk[s], fma(MC.a42, CA_KB(KB_2, s), MC.a41 * CA_KB(KB_1, s))), CA_KB(KB_Y, s));
            }
            rhs2(w, L, m, w.coef[2], ut, kk, dum);   // k4
#pragma unroll
            for (int s = 0; s < kS2; ++s) {
                CA_KB(KB_4, s) = kk[s];
                ut[s] = fma(hh, fma(MC.a54, kk[s], fma(MC.a53, CA_KB(KB_3, s), fma(MC.a52, CA_KB(KB_2, s), MC.a51 * CA_KB(KB_1, s)))), CA_KB(KB_Y, s));
            }
            rhs2(w, L, m, w.coef[3], ut, kk, dum);   // k5
#pragma unroll
            for (int s = 0; s < kS2; ++s) {
                CA_KB(KB_5, s) = kk[s];
                ut[s] = fma(hh, fma(MC.a65, kk[s], fma(MC.a64, CA_KB(KB_4, s), fma(MC.a63, CA_KB(KB_3, s), fma(MC.a62, CA_KB(KB_2, s), MC.a61 * CA_KB(KB_1, s))))),
                            CA_KB(KB_Y, s));
            }
            rhs2(w, L, m, w.coef[4], ut, kk, dum);   // k6
            double er[kS2];
#pragma unroll
            for (int s = 0; s < kS2; ++s) {
                CA_KB(KB_6, s) = kk[s];
                const double k1v = CA_KB(KB_1, s), k3v = CA_KB(KB_3, s), k4v = CA_KB(KB_4, s), k5v = CA_KB(KB_5, s);
                un[s] = fma(hh, fma(MC.b6, kk[s], fma(MC.b5, k5v, fma(MC.b4, k4v, fma(MC.b3, k3v, MC.b1 * k1v)))), CA_KB(KB_Y, s));
                er[s] = fma(MC.e6, kk[s], fma(MC.e5, k5v, fma(MC.e4, k4v, fma(MC.e3, k3v, MC.e1 * k1v))));
            }
            CA_TICK(w, 3);
            rhs2(w, L, m, w.coef[4], un, k7, unt);   // stage 7 = next step's stage 1 (c6 = c7 = 1)
            CA_TICK(w, 2);
            n_rhs += 6;
            bool accept = true;
            double un_ll_new = un_ll;
            if (oo.adaptive) {
                double sacc = 0.0, sdu = 0.0, sde = 0.0;
#pragma unroll
                for (int s = 0; s < kS2; ++s) {
                    const double e = hh * fma(MC.e7, k7[s], er[s]);
                    const double yv = CA_KB(KB_Y, s), yt = CA_KB(KB_YT, s);
                    const double isc = StepCtrl2::rcpn(atol + rtol * StepCtrl2::sqrtn(0.5 * fmax(yv * yv + yt * yt, un[s] * un[s] + unt[s] * unt[s])));
                    const bool eq = L.is_eq(s);
                    sacc += (e * e) * (isc * isc);
                    sdu += eq ? un[s] : 0.0; sde += eq ? e : 0.0;
                }
                w.sum3(sacc, sdu, sde);
                un_ll_new = tr0 - sdu;
                {
                    const double isc = StepCtrl2::rcpn(atol + rtol * fmax(fabs(y_ll), fabs(un_ll_new)));
                    sacc += (sde * sde) * (isc * isc);   // err of ρ_ll,ll = -Σ err of the other diagonal entries
                }
                const double EEst2 = sacc * (1.0 / 625.0);
                const bool acc = EEst2 <= 1.0;                 // NaN and overflow fail it and are sorted out on the rare side
                const double prop = hh * ctrl.decide(EEst2, acc);
                const bool keep = acc && clipped && dt > prop;   // a step clipped at tstop does not shrink the running step size
                dt = keep ? dt : prop;
                if (!acc) {
                    if (!(EEst2 <= 1e300)) { status = CA_ERR_NONFINITE; break; }
                    accept = false;
                    n_rej++;
                    if (dt < 1e-14 * fmax(1.0, fabs(t))) { status = CA_ERR_NONFINITE; break; }
                }
            } else {
                dt = oo.dt0;
                un_ll_new = tr0 - w.sum(diag_sum(un));
                if (!(fabs(un_ll_new) < 1e300)) { status = CA_ERR_NONFINITE; break; }   // a fixed step beyond the stability limit blows up: report it
            }
            if (accept) {
                n_acc++; h = hh; tcommit = tn; un_ll = un_ll_new; pending = true;
#pragma unroll
                for (int s = 0; s < kS2; ++s) CA_KB(KB_YT, s) = unt[s];   // becomes the partner of y at the commit
            }
        }
        // ---- output at T (Hairer contd5 inside the last accepted step [t, tcommit]; KB_Y / KB_1 still hold its start)
        double o[kS2];
        if (status != 0) {
#pragma unroll
            for (int s = 0; s < kS2; ++s) o[s] = 0.0;
        } else if (!pending || T >= tcommit) {
#pragma unroll
            for (int s = 0; s < kS2; ++s) o[s] = pending ? un[s] : CA_KB(KB_Y, s);
        } else {
            const double th = (T - t) / h, th1 = 1.0 - th;
#pragma unroll
            for (int s = 0; s < kS2; ++s) {
                const double yv = CA_KB(KB_Y, s), k1v = CA_KB(KB_1, s);
                const double k3 = CA_KB(KB_3, s), k4 = CA_KB(KB_4, s), k5 = CA_KB(KB_5, s), k6 = CA_KB(KB_6, s);
                const double r2 = un[s] - yv, r3 = h * k1v - r2, r4 = r2 - h * k7[s] - r3;
                const double r5 = h * (MC.d1 * k1v + MC.d3 * k3 + MC.d4 * k4 + MC.d5 * k5 + MC.d6 * k6 + MC.d7 * k7[s]);
                o[s] = yv + th * (r2 + th1 * (r3 + th * (r4 + th1 * r5)));
            }
        }
        accumulate_output2(w, L, m, o, tr0, status, acc + (size_t)j * 2 * kAccSlots * 32);
    }
#undef CA_KB
    st.n_rhs = n_rhs; st.n_acc = n_acc; st.n_rej = n_rej; st.n_exact = n_exact; st.status = status;
}
#else
// One trajectory: DP5 (OrdinaryDiffEq semantics, see oracle) over tspan; ρ and ρ·ρ at every tspan point are ADDED to
// acc[((j*2 + which)*kAccSlots + slot)*32] (pointer already offset by the lane).
template <class W>
CA_HD void integrate2(W& w, const Lane2& L, const DevModel& m, const OdeOpts& oo, const double* tspan, int nt, const double* tab,
                      const double* rho0, double tr0, double* acc, Stats2& st) {
    const double rtol = oo.rtol, atol = oo.atol;
    double* kb = w.kb;   // kb[(buf*kS2 + s)*32], lane offset included
#if CA_Y_SMEM
#define CA_YL(s) kb[(KB_Y * kS2 + (s)) * 32]
#else
    double y[kS2];
#define CA_YL(s) y[s]
#endif
#if CA_K1_SMEM
#define CA_K1L(s) kb[(KB_1 * kS2 + (s)) * 32]
#else
    double k1[kS2];
#define CA_K1L(s) k1[s]
#endif
    double un[kS2], k7[kS2], unt[kS2], ut[kS2];
#pragma unroll
    for (int s = 0; s < kS2; ++s) { CA_YL(s) = rho0[s]; un[s] = rho0[s]; }
    const double t0 = tspan[0], tend = tspan[nt - 1];
    double t = t0, h = 0.0, tcommit = t0, dt = oo.dt0;
    StepCtrl2 ctrl;
    ctrl.reset();
    long long iters = 0;
    int status = 0, n_rhs = 0, n_acc = 0, n_rej = 0, n_exact = 0;
    w.anc[AN_TA * 32] = 1e300; w.anc[AN_TEND * 32] = -1e300; w.anc[AN_J * 32] = 0.0;   // no anchor
    CoefLane2 cl;
    cl.reset(w.lane);   // empty noise window
    auto diag_sum = [&](const double (&v)[kS2]) {
        double s = 0.0;
#pragma unroll
        for (int j = 0; j < kS2; ++j) s += L.is_eq(j) ? v[j] : 0.0;
        return s;
    };
    // k1 = f(t0, y)
    eval_coefficients2(w, m, tab, cl, t0, 0.0, t0, false, n_exact);
    rhs2(w, L, m, w.coef[kStages2 - 1], un, k7, unt);
#pragma unroll
    for (int s = 0; s < kS2; ++s) { kb[(KB_YT * kS2 + s) * 32] = unt[s]; CA_K1L(s) = k7[s]; }
    n_rhs = 1;
    double y_ll = tr0 - w.sum(diag_sum(un)), un_ll = y_ll;
    if (dt <= 0.0) {  // OrdinaryDiffEq initial step (Hairer), norms over all 625 entries
        double d0 = 0.0, d1 = 0.0, isk2[kS2];
#pragma unroll
        for (int s = 0; s < kS2; ++s) {
            const double a2 = 0.5 * (CA_YL(s) * CA_YL(s) + unt[s] * unt[s]);   // |ρ_AB|²
            const double sk = atol + sqrt(a2) * rtol;
            isk2[s] = 1.0 / (sk * sk);   // the entries (A,B) and (B,A) of the full matrix share |err|² = (e² + e_t²)/2: weight 1 per real
            d0 += isk2[s] * CA_YL(s) * CA_YL(s);
            d1 += isk2[s] * CA_K1L(s) * CA_K1L(s);
        }
        const double f0_ll = -w.sum(diag_sum(k7));
        const double sk_ll = atol + fabs(y_ll) * rtol;
        d0 = sqrt((w.sum(d0) + y_ll * y_ll / (sk_ll * sk_ll)) / 625.0);
        d1 = sqrt((w.sum(d1) + f0_ll * f0_ll / (sk_ll * sk_ll)) / 625.0);
        double dt0 = (d0 < 1e-5 || d1 < 1e-5) ? 1e-6 : 0.01 * d0 / d1;
        dt0 = fmin(dt0, oo.dtmax);
#pragma unroll
        for (int s = 0; s < kS2; ++s) ut[s] = CA_YL(s) + dt0 * CA_K1L(s);
        eval_coefficients2(w, m, tab, cl, t0, 0.0, t0 + dt0, false, n_exact);
        double k2[kS2], dum[kS2];
        rhs2(w, L, m, w.coef[kStages2 - 1], ut, k2, dum);
        n_rhs++;
        double d2 = 0.0;
#pragma unroll
        for (int s = 0; s < kS2; ++s) { const double a = k2[s] - CA_K1L(s); d2 += isk2[s] * a * a; }
        const double f1_ll = -w.sum(diag_sum(k2));
        d2 = sqrt((w.sum(d2) + (f1_ll - f0_ll) * (f1_ll - f0_ll) / (sk_ll * sk_ll)) / 625.0) / dt0;
        const double mx = fmax(d1, d2);
        const double dt1 = (mx <= 1e-15) ? fmax(1e-6, dt0 * 1e-3) : pow(10.0, -(2.0 + log10(mx)) / 5.0);
        dt = fmin(100.0 * dt0, fmin(dt1, oo.dtmax));
    }
    for (int j = 0; j < nt; ++j) {
        const double T = tspan[j];
        const double tstop = oo.dense ? tend : T;
        while (status == 0 && tcommit < T) {
            if (h != 0.0) {  // commit the pending accepted step (FSAL)
#pragma unroll
                for (int s = 0; s < kS2; ++s) { CA_YL(s) = un[s]; CA_K1L(s) = k7[s]; }
                y_ll = un_ll;
                t = tcommit; h = 0.0;
            }
            if (++iters > oo.maxiters) { status = CA_ERR_MAXITERS; break; }
            double hh = fmin(dt, oo.dtmax);
            bool clipped = false;
            if (t + hh >= tstop || (tstop - (t + hh)) < 1e-13 * fabs(tstop)) { hh = tstop - t; clipped = true; }
            const double tn = clipped ? tstop : t + hh;
            CA_TICK(w, 3);
            eval_coefficients2(w, m, tab, cl, t, hh, tn, ((iters - 1) & (kAnchorEvery2 - 1)) == 0, n_exact);
            const bool for_dense = tn > T || CA_K3_SMEM || CA_K5_SMEM || CA_DENSE_ALWAYS;   // this step may be interpolated at T (warp-uniform)
            CA_TICK(w, 1);
            double k2[kS2], k3[kS2], k5[kS2], kk[kS2], er[kS2], dum[kS2];
#pragma unroll
            for (int s = 0; s < kS2; ++s) ut[s] = fma(hh * MC.a21, CA_K1L(s), CA_YL(s));
            rhs2(w, L, m, w.coef[0], ut, k2, dum);
#pragma unroll
            for (int s = 0; s < kS2; ++s) ut[s] = fma(hh, fma(MC.a32, k2[s], MC.a31 * CA_K1L(s)), CA_YL(s));
            rhs2(w, L, m, w.coef[1], ut, k3, dum);
#pragma unroll
            for (int s = 0; s < kS2; ++s) {
                if (for_dense) kb[(KB_3 * kS2 + s) * 32] = k3[s];
                ut[s] = fma(hh, fma(MC.a43, k3[s], fma(MC.a42, k2[s], MC.a41 * CA_K1L(s))), CA_YL(s));
            }
            rhs2(w, L, m, w.coef[2], ut, kk, dum);   // k4
#pragma unroll
            for (int s = 0; s < kS2; ++s) {
                kb[(KB_4 * kS2 + s) * 32] = kk[s];
                ut[s] = fma(hh, fma(MC.a54, kk[s], fma(MC.a53, CA_K3L(s), fma(MC.a52, k2[s], MC.a51 * CA_K1L(s)))), CA_YL(s));
            }
            rhs2(w, L, m, w.coef[3], ut, k5, dum);
#pragma unroll
            for (int s = 0; s < kS2; ++s) {
                if (for_dense) kb[(KB_5 * kS2 + s) * 32] = k5[s];
                const double k4 = kb[(KB_4 * kS2 + s) * 32];
                ut[s] = fma(hh, fma(MC.a65, k5[s], fma(MC.a64, k4, fma(MC.a63, CA_K3L(s), fma(MC.a62, k2[s], MC.a61 * CA_K1L(s))))), CA_YL(s));
            }
            rhs2(w, L, m, w.coef[4], ut, kk, dum);   // k6
#pragma unroll
            for (int s = 0; s < kS2; ++s) {
                if (for_dense) kb[(KB_6 * kS2 + s) * 32] = kk[s];
                const double k4 = kb[(KB_4 * kS2 + s) * 32];
                const double k3v = CA_K3L(s), k5v = CA_K5L(s);
                un[s] = fma(hh, fma(MC.b6, kk[s], fma(MC.b5, k5v, fma(MC.b4, k4, fma(MC.b3, k3v, MC.b1 * CA_K1L(s))))), CA_YL(s));
                er[s] = fma(MC.e6, kk[s], fma(MC.e5, k5v, fma(MC.e4, k4, fma(MC.e3, k3v, MC.e1 * CA_K1L(s)))));
            }
            CA_TICK(w, 3);
            rhs2(w, L, m, w.coef[4], un, k7, unt);   // stage 7 = next step's stage 1 (c6 = c7 = 1)
            CA_TICK(w, 2);
            n_rhs += 6;
            bool accept = true;
            double un_ll_new = un_ll;
            if (oo.adaptive) {
                double sacc = 0.0, sdu = 0.0, sde = 0.0;
#pragma unroll
                for (int s = 0; s < kS2; ++s) {
                    const double e = hh * fma(MC.e7, k7[s], er[s]);
                    const double yt = kb[(KB_YT * kS2 + s) * 32];
                    const double isc = StepCtrl2::rcpn(atol + rtol * StepCtrl2::sqrtn(0.5 * fmax(CA_YL(s) * CA_YL(s) + yt * yt, un[s] * un[s] + unt[s] * unt[s])));
                    const bool eq = L.is_eq(s);
                    sacc += (e * e) * (isc * isc);
                    sdu += eq ? un[s] : 0.0; sde += eq ? e : 0.0;
                }
                w.sum3(sacc, sdu, sde);
                un_ll_new = tr0 - sdu;
                {
                    const double isc = StepCtrl2::rcpn(atol + rtol * fmax(fabs(y_ll), fabs(un_ll_new)));
                    sacc += (sde * sde) * (isc * isc);   // err of ρ_ll,ll = -Σ err of the other diagonal entries
                }
                const double EEst2 = sacc * (1.0 / 625.0);
                const bool acc = EEst2 <= 1.0;                 // NaN and overflow fail it and are sorted out on the rare side
                const double prop = hh * ctrl.decide(EEst2, acc);
                const bool keep = acc && clipped && dt > prop;   // a step clipped at tstop does not shrink the running step size
                dt = keep ? dt : prop;
                if (!acc) {
                    if (!(EEst2 <= 1e300)) { status = CA_ERR_NONFINITE; break; }
                    accept = false;
                    n_rej++;
                    if (dt < 1e-14 * fmax(1.0, fabs(t))) { status = CA_ERR_NONFINITE; break; }
                }
            } else {
                dt = oo.dt0;
                un_ll_new = tr0 - w.sum(diag_sum(un));
                if (!(fabs(un_ll_new) < 1e300)) { status = CA_ERR_NONFINITE; break; }   // a fixed step beyond the stability limit blows up: report it
            }
            if (accept) {
                n_acc++; h = hh; tcommit = tn; un_ll = un_ll_new;
#pragma unroll
                for (int s = 0; s < kS2; ++s) kb[(KB_YT * kS2 + s) * 32] = unt[s];   // becomes the partner of y at the commit
            }
        }
        // ---- output at T (Hairer contd5 inside the last accepted step)
        double o[kS2];
        if (h == 0.0 || T >= tcommit) {
#pragma unroll
            for (int s = 0; s < kS2; ++s) o[s] = un[s];
        } else {
            const double th = (T - t) / h, th1 = 1.0 - th;
#pragma unroll
            for (int s = 0; s < kS2; ++s) {
                const double k3 = kb[(KB_3 * kS2 + s) * 32], k4 = kb[(KB_4 * kS2 + s) * 32], k5 = kb[(KB_5 * kS2 + s) * 32], k6 = kb[(KB_6 * kS2 + s) * 32];
                const double r2 = un[s] - CA_YL(s), r3 = h * CA_K1L(s) - r2, r4 = r2 - h * k7[s] - r3;
                const double r5 = h * (MC.d1 * CA_K1L(s) + MC.d3 * k3 + MC.d4 * k4 + MC.d5 * k5 + MC.d6 * k6 + MC.d7 * k7[s]);
                o[s] = CA_YL(s) + th * (r2 + th1 * (r3 + th * (r4 + th1 * r5)));
            }
        }
        if (status != 0) {
#pragma unroll
            for (int s = 0; s < kS2; ++s) o[s] = 0.0;
        }
        accumulate_output2(w, L, m, o, tr0, status, acc + (size_t)j * 2 * kAccSlots * 32);
    }
#undef CA_YL
#undef CA_K1L
    st.n_rhs = n_rhs; st.n_acc = n_acc; st.n_rej = n_rej; st.n_exact = n_exact; st.status = status;
}
#endif

}  // namespace ca
