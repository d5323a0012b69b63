// ca_lindblad2.cu -- K4: two-atom CZ ensemble (simulation_czlp, src/cz_model.jl:107-171) on sm_100a.
//
// One trajectory per WARP.  The 25x25 density matrix is reduced to its reachable, Hermitian-packed support
// (for ψ0 without |l> components 16x16 ⊕ 4x4 ⊕ 4x4 ⊕ 1 = 157 packed elements instead of 625) which
// is dealt across the 32 lanes (5 complex slots per lane) and kept in REGISTERS for all DP5 buffers.  Only the
// current stage vector lives in shared memory (full, un-packed, 4.6 KB per warp): every lane gathers the few
// entries its elements couple to through a term table that the host derives symbolically from the reference's
// operator lists (operators ⊗ Id, Id ⊗ operators, nr ⊗ nr: cz_model.jl:31; JumpOperatorsTwo: cz_model.jl:2-9).
// The time-dependent coefficients (4 beams with their own phase-noise traces, Doppler shifts, V(t) = c6/R^6:
// cz_model.jl:12-17,37-65) of all five distinct DP5 stage times are evaluated in ONE pass per step, one
// (quantity, stage) pair per lane.  Error norm and observables are warp-shuffle reductions; ρ and ρ·ρ are
// accumulated per warp in global memory without atomics and reduced by k_finalize2.
#include <cuda_runtime.h>

#include <algorithm>
#include <cstring>
#include <map>
#include <vector>

#include "ca_host.h"
#include "ca_device.cuh"

namespace ca {

cudaStream_t ctx_stream(ca_ctx* c);
int ctx_sm_count(ca_ctx* c);
int ctx_device(ca_ctx* c);
void* ctx_buf(ca_ctx* c, int which, size_t bytes);  // grow-only device buffers owned by the context
cudaEvent_t ctx_event(ca_ctx* c, int i);

constexpr int kSlots = 5;          // complex slots per lane
constexpr int kMaxTerms = 12;      // gather terms per slot (max over lanes), excluding the self term
constexpr int kMaxSq = 16;         // terms of (ρ·ρ)_IJ
constexpr int kWarps2 = 8;         // warps (trajectories) per CTA
constexpr int kMaxFull = 320;      // stage vector entries per warp (289 used for |++>)
constexpr int kNCoef = 24;         // per-stage coefficient table: 16 dynamic (∓i c_k) + 4 rates + zero
constexpr int kStages = 5;         // distinct stage times c2,c3,c4,c5,c6(=c7)
// Gather terms per slot position after dealing the elements by decreasing term count.  These are the order statistics
// of the full 16x16+4x4+4x4+1 support with all decay channels on; any sub-support / fewer channels has smaller ones.
// The gather loops are unrolled to these counts and the per-lane descriptors live in registers.
__host__ __device__ constexpr int slot_terms(int s) { return s == 0 ? 8 : (s == 1 ? 5 : (s == 2 ? 5 : (s == 3 ? 4 : 3))); }
__host__ __device__ constexpr int slot_offset(int s) { return s == 0 ? 0 : (s == 1 ? 8 : (s == 2 ? 13 : (s == 3 ? 18 : 22))); }
constexpr int kTermRegs = 25;

// Static structure shared by all trajectories (device copy lives in global memory, staged into smem).
struct Struct2 {
    int n_full;                                   // entries of the stage vector
    int n_terms[kSlots];                          // padded gather-term count per slot position
    int n_sq[kSlots];
    unsigned terms[kSlots][kMaxTerms][32];        // (idx << 6) | cid
    unsigned sq[kSlots][kMaxSq][32];              // (idxA << 16) | idxB
    unsigned short own_idx[kSlots][32];           // index of (I,J) in the stage vector (n_full = unused slot)
    unsigned short own_tidx[kSlots][32];          // index of (J,I)
    short row[kSlots][32], col[kSlots][32];       // composite indices I, J (0..24), -1 = unused
    signed char nd[kSlots][32][5];                // E_I - E_J = -(nd0 dp1 + nd1 dr1 + nd2 dp2 + nd3 dr2) + nd4 V
    float weight[kSlots][32];                     // error-norm multiplicity: 1 diagonal, 2 off-diagonal, 0 unused
    double decay[kSlots][32];                     // (γ_I + γ_J)/2
    double rho0[kSlots][32][2];
};

// ----------------------------------------------------------------------------------------------------
// host: symbolic derivation of the gather structure from the reference's operator lists
// ----------------------------------------------------------------------------------------------------
namespace {
enum { L0 = 0, L1 = 1, LR = 2, LP = 3, LL = 4 };
struct Ent { int r, c; };

// cid layout: 0..7  left  terms  -i c_k for k = σ1p(1), σp1(1), σpr(1), σrp(1), σ1p(2), σp1(2), σpr(2), σrp(2)
//             8..15 right terms  +i c_k
//             16..19 rates Γ0, Γ1, Γl, Γr ; 20 zero
constexpr int CID_RATE = 16, CID_ZERO = 20;

int build_struct2(const double* psi0, const DevModel& m, Struct2* S) {
    std::memset(S, 0, sizeof *S);
    const int d = 25;
    // off-diagonal Hamiltonian operators (k = 0..7) and their sparse entries on the 25-dim space (index a1 + 5 a2)
    const int lv[4][2] = {{L1, LP}, {LP, L1}, {LP, LR}, {LR, LP}};  // σ1p, σp1, σpr, σrp (rydberg_model.jl:11-16,25)
    std::vector<std::vector<Ent>> Hop(8);
    for (int a = 0; a < 2; ++a)
        for (int k = 0; k < 4; ++k)
            for (int o = 0; o < 5; ++o) {
                int r = a == 0 ? lv[k][0] + 5 * o : o + 5 * lv[k][0];
                int c = a == 0 ? lv[k][1] + 5 * o : o + 5 * lv[k][1];
                Hop[4 * a + k].push_back({r, c});
            }
    // jumps σ0p, σ1p, σlp, σlr on each atom (cz_model.jl:2-9); rate index 0..3
    const int jl[4][2] = {{L0, LP}, {L1, LP}, {LL, LP}, {LL, LR}};
    struct Jmp { std::vector<Ent> e; int rate; };
    std::vector<Jmp> J;
    for (int a = 0; a < 2; ++a)
        for (int k = 0; k < 4; ++k) {
            Jmp j; j.rate = k;
            for (int o = 0; o < 5; ++o) j.e.push_back({a == 0 ? jl[k][0] + 5 * o : o + 5 * jl[k][0], a == 0 ? jl[k][1] + 5 * o : o + 5 * jl[k][1]});
            J.push_back(j);
        }
    const double rates[4] = {m.G0, m.G1, m.Gl, m.Gr};
    // reachable support of ρ
    std::vector<char> act(d * d, 0);
    for (int i = 0; i < d; ++i)
        for (int j = 0; j < d; ++j)
            if ((psi0[2 * i] != 0 || psi0[2 * i + 1] != 0) && (psi0[2 * j] != 0 || psi0[2 * j + 1] != 0)) act[i + d * j] = 1;
    for (bool grew = true; grew;) {
        grew = false;
        auto mark = [&](int i, int j) { if (!act[i + d * j]) { act[i + d * j] = 1; act[j + d * i] = 1; grew = true; } };
        for (int i = 0; i < d; ++i)
            for (int j = 0; j < d; ++j) {
                if (!act[i + d * j]) continue;
                for (auto& op : Hop)
                    for (auto& e : op) { if (e.c == i) mark(e.r, j); if (e.r == j) mark(i, e.c); }
                for (auto& jp : J) {
                    if (rates[jp.rate] == 0.0) continue;
                    for (auto& e1 : jp.e) for (auto& e2 : jp.e) if (e1.c == i && e2.c == j) mark(e1.r, e2.r);
                }
            }
    }
    // stage-vector index of every active full entry
    std::vector<int> uidx(d * d, -1);
    int n_full = 0;
    for (int j = 0; j < d; ++j) for (int i = 0; i < d; ++i) if (act[i + d * j]) uidx[i + d * j] = n_full++;
    if (n_full + 1 > kMaxFull) { set_error("two-atom kernel: reachable support too large (ψ0 with |l> components is not supported on the GPU path)"); return CA_ERR_UNSUPPORTED; }
    S->n_full = n_full;  // entry n_full is a permanent zero
    // owned (packed) elements with their gather terms
    struct Elem { int I, J; std::vector<unsigned> terms; std::vector<unsigned> sq; };
    std::vector<Elem> el;
    for (int Jc = 0; Jc < d; ++Jc)
        for (int I = 0; I <= Jc; ++I) {
            if (!act[I + d * Jc]) continue;
            Elem e; e.I = I; e.J = Jc;
            for (int k = 0; k < 8; ++k)
                for (auto& en : Hop[k]) {
                    if (en.r == I && act[en.c + d * Jc]) e.terms.push_back(((unsigned)uidx[en.c + d * Jc] << 6) | (unsigned)k);          // -i c_k ρ[c,J]
                    if (en.c == Jc && act[I + d * en.r]) e.terms.push_back(((unsigned)uidx[I + d * en.r] << 6) | (unsigned)(8 + k));   // +i ρ[I,r] c_k
                }
            for (auto& jp : J) {
                if (rates[jp.rate] == 0.0) continue;
                for (auto& e1 : jp.e) for (auto& e2 : jp.e)
                    if (e1.r == I && e2.r == Jc && act[e1.c + d * e2.c]) e.terms.push_back(((unsigned)uidx[e1.c + d * e2.c] << 6) | (unsigned)(CID_RATE + jp.rate));
            }
            for (int K = 0; K < d; ++K)
                if (act[I + d * K] && act[K + d * Jc]) e.sq.push_back(((unsigned)uidx[I + d * K] << 16) | (unsigned)uidx[K + d * Jc]);
            if ((int)e.terms.size() > kMaxTerms || (int)e.sq.size() > kMaxSq) { set_error("two-atom kernel: term table overflow"); return CA_ERR_UNSUPPORTED; }
            el.push_back(e);
        }
    if ((int)el.size() > kSlots * 32) { set_error("two-atom kernel: reachable support too large for 5 slots per lane (ψ0 with |l> components is not supported on the GPU path)"); return CA_ERR_UNSUPPORTED; }
    // deal elements so that every slot position holds elements of similar term count (little padding)
    std::stable_sort(el.begin(), el.end(), [](const Elem& a, const Elem& b) { return a.terms.size() > b.terms.size(); });
    const double gam[5] = {0.0, 0.0, m.Gr, m.G0 + m.G1 + m.Gl, 0.0};
    for (int s = 0; s < kSlots; ++s)
        for (int l = 0; l < 32; ++l) {
            S->own_idx[s][l] = S->own_tidx[s][l] = (unsigned short)n_full;
            S->row[s][l] = S->col[s][l] = -1;
            for (int t = 0; t < kMaxTerms; ++t) S->terms[s][t][l] = ((unsigned)n_full << 6) | CID_ZERO;
            for (int t = 0; t < kMaxSq; ++t) S->sq[s][t][l] = ((unsigned)n_full << 16) | (unsigned)n_full;
        }
    for (size_t q = 0; q < el.size(); ++q) {
        const int s = (int)(q / 32), l = (int)(q % 32);
        const Elem& e = el[q];
        S->own_idx[s][l] = (unsigned short)uidx[e.I + d * e.J];
        S->own_tidx[s][l] = (unsigned short)uidx[e.J + d * e.I];
        S->row[s][l] = (short)e.I; S->col[s][l] = (short)e.J;
        S->n_terms[s] = std::max(S->n_terms[s], (int)e.terms.size());
        S->n_sq[s] = std::max(S->n_sq[s], (int)e.sq.size());
        for (size_t t = 0; t < e.terms.size(); ++t) S->terms[s][t][l] = e.terms[t];
        for (size_t t = 0; t < e.sq.size(); ++t) S->sq[s][t][l] = e.sq[t];
        const int a1 = e.I % 5, a2 = e.I / 5, b1 = e.J % 5, b2 = e.J / 5;
        S->nd[s][l][0] = (signed char)((a1 == LP) - (b1 == LP)); S->nd[s][l][1] = (signed char)((a1 == LR) - (b1 == LR));
        S->nd[s][l][2] = (signed char)((a2 == LP) - (b2 == LP)); S->nd[s][l][3] = (signed char)((a2 == LR) - (b2 == LR));
        S->nd[s][l][4] = (signed char)((a1 == LR && a2 == LR) - (b1 == LR && b2 == LR));
        S->weight[s][l] = e.I == e.J ? 1.0f : 2.0f;
        S->decay[s][l] = 0.5 * (gam[a1] + gam[a2] + gam[b1] + gam[b2]);
        const double ar = psi0[2 * e.I], ai = psi0[2 * e.I + 1], br = psi0[2 * e.J], bi = psi0[2 * e.J + 1];
        S->rho0[s][l][0] = ar * br + ai * bi; S->rho0[s][l][1] = ai * br - ar * bi;
    }
    return CA_OK;
}
}  // namespace

// ----------------------------------------------------------------------------------------------------
// device
// ----------------------------------------------------------------------------------------------------
struct R2Params {
    DevModel model;
    const Struct2* st;
    const double* tspan;
    const double* samples1;  // NULL or [n][6] (before the centre shift)
    const double* samples2;
    const double* phases[4]; // NULL or [n][nf]: red1, blue1, red2, blue2
    const double* f;
    const double* amp_red;
    const double* amp_blue;
    double* scratch;         // per warp: [4][n_nodes] phase-noise tables
    double* accum;           // per warp: [nt][2 (ρ, ρ²)][kSlots][32][2]
    unsigned long long* counters;
    OdeOpts ode;
    long long n_traj, traj_offset;
    unsigned long long seed;
    int nt, nf;
};

struct Smem2 {
    Struct2 st;
    double us[kWarps2][kMaxFull][2];
    double ctab[kWarps2][kStages][kNCoef][2];
    double dtab[kWarps2][kStages][5];   // dp1, dr1, dp2, dr2, V at each stage time
    double smp[kWarps2][2][6];
};

__device__ __forceinline__ double warp_sum2(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// One (quantity, stage time) pair per lane: q = 0..3 -> Ω/2 of red1, blue1, red2, blue2 (with phase noise and the
// Levine-Pichler phase jump on the red beams), q = 4 -> Doppler detunings and V(t).
__device__ void eval_coefficients(const DevModel& m, Smem2& sm, int w, int lane, const double* tab, double t0, double hh, double tn) {
    const int q = lane % 5, si = lane / 5;
    if (si < kStages) {
        const double cs[kStages] = {dp::c2, dp::c3, dp::c4, dp::c5, 1.0};
        const double tt = si == kStages - 1 ? tn : t0 + cs[si] * hh;
        double pos[2][3], vel[2][2];
#pragma unroll
        for (int a = 0; a < 2; ++a) {
            const double* s = sm.smp[w][a];
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
        if (q < 4) {
            const int a = q >> 1;
            double ph = 0.0;
            if (tab) {
                int j = (int)(tt * m.inv_dtn);
                j = j < 0 ? 0 : (j > m.n_nodes - 2 ? m.n_nodes - 2 : j);
                const double* p = tab + (size_t)q * m.n_nodes + j;
                const double wgt = (tt - j * m.dtn) * m.inv_dtn;
                ph = p[0] + wgt * (p[1] - p[0]);
            }
            if (!(q & 1) && tt >= m.tau) ph += m.xi;  // ϕr(t) = t < τ ? 0 : ξ on both red beams (cz_model.jl:136-137,48)
            double re, im;
            const DevBeam& bm = (q & 1) ? m.blue : m.red;
            if (bm.kind == CA_BEAM_GAUSS) {
                double anc[5];
                const double rp = sqrt_d(fmax(pos[a][0] * pos[a][0] + pos[a][1] * pos[a][1], 1e-300));
                gauss_exact(bm, rp, pos[a][2], ph, &re, &im, anc);
            } else beam_half_omega(bm, pos[a][0], pos[a][1], pos[a][2], ph, &re, &im);
            // operators of atom a: k = 4a + {σ1p: Ωr/2, σp1: conj, σpr: Ωb/2, σrp: conj}; left coefficient -i c, right +i c
            double (*ct)[2] = sm.ctab[w][si];
            const int k0 = 4 * a + ((q & 1) ? 2 : 0);
            ct[k0][0] = im; ct[k0][1] = -re;            // -i (re + i im)
            ct[k0 + 1][0] = -im; ct[k0 + 1][1] = -re;   // -i (re - i im)
            ct[8 + k0][0] = -im; ct[8 + k0][1] = re;    // +i (re + i im)
            ct[8 + k0 + 1][0] = im; ct[8 + k0 + 1][1] = re;
        } else {
            double* dt_ = sm.dtab[w][si];
#pragma unroll
            for (int a = 0; a < 2; ++a) {
                const double Dd = m.aDx * vel[a][0] + m.aDz * vel[a][1], dd = m.bDx * vel[a][0] + m.bDz * vel[a][1];
                dt_[2 * a] = m.Delta0 - m.dsign * Dd;      // dp: H_pp = -dp
                dt_[2 * a + 1] = m.delta0 - m.dsign * dd;  // dr: H_rr = -dr
            }
            const double dx = pos[0][0] - pos[1][0], dy = pos[0][1] - pos[1][1], dz = pos[0][2] - pos[1][2];
            const double r2 = dx * dx + dy * dy + dz * dz;
            dt_[4] = m.c6 * rcp_d(1e-18 + r2 * r2 * r2);   // get_V: cz_model.jl:12-17
        }
    }
    __syncwarp();
}

template <int S>
struct Regs2 { double v[S][2]; };

__global__ void __launch_bounds__(kWarps2 * 32, 1) k_lindblad2(const __grid_constant__ R2Params P) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    Smem2& sm = *reinterpret_cast<Smem2*>(smem_raw);
    const DevModel& m = P.model;
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    {   // stage the static structure
        const unsigned* src = reinterpret_cast<const unsigned*>(P.st);
        unsigned* dst = reinterpret_cast<unsigned*>(&sm.st);
        for (int i = threadIdx.x; i < (int)(sizeof(Struct2) / 4); i += blockDim.x) dst[i] = src[i];
    }
    __syncthreads();
    const Struct2& st = sm.st;
    const long long wg = (long long)blockIdx.x * kWarps2 + w, W = (long long)gridDim.x * kWarps2;
    // per-lane static data
    unsigned short oidx[kSlots], otidx[kSlots];
    double decay[kSlots], wgt[kSlots];
    int ndp[kSlots];
#pragma unroll
    for (int s = 0; s < kSlots; ++s) {
        oidx[s] = st.own_idx[s][lane]; otidx[s] = st.own_tidx[s][lane];
        decay[s] = st.decay[s][lane]; wgt[s] = st.weight[s][lane];
        ndp[s] = (st.nd[s][lane][0] & 0xff) | ((st.nd[s][lane][1] & 0xff) << 8) | ((st.nd[s][lane][2] & 0xff) << 16) | ((st.nd[s][lane][3] & 0xff) << 24);
    }
    unsigned dreg[kTermRegs];
    signed char ndv[kSlots];
#pragma unroll
    for (int s = 0; s < kSlots; ++s) {
        ndv[s] = st.nd[s][lane][4];
#pragma unroll
        for (int t = 0; t < slot_terms(s); ++t) dreg[slot_offset(s) + t] = st.terms[s][t][lane];
    }
    double (*us)[2] = sm.us[w];
    for (int i = lane; i < kMaxFull; i += 32) { us[i][0] = 0.0; us[i][1] = 0.0; }
    for (int si = 0; si < kStages; ++si)
        for (int i = lane; i < kNCoef; i += 32) {
            double v = 0.0;
            if (i >= CID_RATE && i < CID_RATE + 4) v = i == CID_RATE ? m.G0 : (i == CID_RATE + 1 ? m.G1 : (i == CID_RATE + 2 ? m.Gl : m.Gr));
            sm.ctab[w][si][i][0] = v; sm.ctab[w][si][i][1] = 0.0;
        }
    __syncwarp();
    long long tot_rhs = 0, tot_acc = 0, tot_rej = 0, tot_att = 0;
    int worst = 0;
    double* acc_w = P.accum + (size_t)wg * P.nt * 2 * kSlots * 32 * 2;
    double* tab_w = (m.laser_noise && P.nf > 0) ? P.scratch + (size_t)wg * 4 * m.n_nodes : nullptr;
    const double rtol = P.ode.rtol, atol = P.ode.atol;

    // RHS of this lane's elements at stage table `si`, stage vector in smem, own values in `u`.  All loads of a slot are
    // independent (descriptors are registers, the loops are unrolled), so their latencies overlap.
    auto rhs = [&](int si, const double (&u)[kSlots][2], double (&du)[kSlots][2]) {
        const double (*ct)[2] = sm.ctab[w][si];
        const double* dtb = sm.dtab[w][si];
        const double d0 = dtb[0], d1 = dtb[1], d2 = dtb[2], d3 = dtb[3], V = dtb[4];
#pragma unroll
        for (int s = 0; s < kSlots; ++s) {
            // self term: (-i (E_I - E_J) - (γ_I+γ_J)/2) ρ_IJ with E = -dp, -dr on the p, r levels and +V on |rr>
            const double dE = -((double)(signed char)(ndp[s] & 0xff) * d0 + (double)(signed char)((ndp[s] >> 8) & 0xff) * d1 +
                                (double)(signed char)((ndp[s] >> 16) & 0xff) * d2 + (double)(signed char)((ndp[s] >> 24) & 0xff) * d3) +
                              (double)ndv[s] * V;
            double ar = -decay[s] * u[s][0] + dE * u[s][1];
            double ai = -decay[s] * u[s][1] - dE * u[s][0];
            double br = 0.0, bi = 0.0;
#pragma unroll
            for (int t = 0; t < slot_terms(s); ++t) {
                const unsigned dsc = dreg[slot_offset(s) + t];
                const double2 c = *reinterpret_cast<const double2*>(ct[dsc & 63u]);
                const double2 r = *reinterpret_cast<const double2*>(us[dsc >> 6]);
                if (t & 1) { br = fma(c.x, r.x, br); br = fma(-c.y, r.y, br); bi = fma(c.x, r.y, bi); bi = fma(c.y, r.x, bi); }
                else { ar = fma(c.x, r.x, ar); ar = fma(-c.y, r.y, ar); ai = fma(c.x, r.y, ai); ai = fma(c.y, r.x, ai); }
            }
            du[s][0] = ar + br; du[s][1] = wgt[s] == 1.0 ? 0.0 : ai + bi;  // diagonal elements are real
        }
    };
    auto publish = [&](const double (&u)[kSlots][2]) {  // owners write (I,J) and (J,I) = conj
        __syncwarp();
#pragma unroll
        for (int s = 0; s < kSlots; ++s) {
            if (wgt[s] != 0.0) {
                us[oidx[s]][0] = u[s][0]; us[oidx[s]][1] = u[s][1];
                us[otidx[s]][0] = u[s][0]; us[otidx[s]][1] = -u[s][1];
            }
        }
        __syncwarp();
    };

    for (long long tr = wg; tr < P.n_traj; tr += W) {
        const uint64_t gi = (uint64_t)(P.traj_offset + tr);
        // ---- samples (cz_model.jl:111-125): lanes 0,1 draw / load atom 1,2 and add the centre shift
        if (lane < 2) {
            double s6[6] = {0, 0, 0, 0, 0, 0};
            const double* src = lane == 0 ? P.samples1 : P.samples2;
            if (src) { for (int q = 0; q < 6; ++q) s6[q] = src[tr * 6 + q]; }
            else {
                SamplerConsts sc;
                for (int q = 0; q < 6; ++q) sc.sig[q] = m.sig[q];
                sc.U0 = m.U0; sc.w0 = m.w0; sc.z0 = m.z0; sc.mfac = m.mfac; sc.ecut = m.ecut;
                int att = sample_atom_literal(sc, P.seed, gi, (uint32_t)lane, s6);
                if (att < 0) worst = max(worst, -CA_ERR_NONFINITE); else tot_att += att;
            }
            for (int q = 0; q < 3; ++q) s6[q] += m.center[lane][q];
            for (int q = 0; q < 6; ++q) sm.smp[w][lane][q] = m.atom_motion ? s6[q] : 0.0;
        }
        // ---- phase-noise tables: lane = chunk of 32 nodes, all four traces (cz_model.jl:41-46)
        if (tab_w) {
            for (int q = 0; q < 4; ++q) {
                const double* ph = P.phases[q] ? P.phases[q] + tr * P.nf : nullptr;
                for (int c0 = lane * 32; c0 < m.n_nodes; c0 += 32 * 32) {
                    double acc[32];
#pragma unroll
                    for (int j = 0; j < 32; ++j) acc[j] = 0.0;
                    const double t0 = c0 * m.dtn;
                    const double* amp = (q & 1) ? P.amp_blue : P.amp_red;
                    double ph_next = 0.0;
                    for (int k = 0; k < P.nf; ++k) {
                        double phk;
                        if (ph) phk = ph[k];
                        else if (k & 1) phk = ph_next;
                        else { U4 r4 = philox(P.seed, gi, (uint32_t)(k >> 1), (uint32_t)(2 + q)); phk = kTwoPi * u53(r4.x, r4.y); ph_next = kTwoPi * u53(r4.z, r4.w); }
                        const double om = kTwoPi * P.f[k];
                        double s0, c0v, sh, ch;
                        sincos(om * t0 + phk, &s0, &c0v);
                        sincos(0.5 * om * m.dtn, &sh, &ch);
                        const double a = amp[k], sig = -4.0 * sh * sh;
                        double c = a * c0v, dlt = -2.0 * a * (s0 * ch + c0v * sh) * sh;
                        acc[0] += c;
#pragma unroll
                        for (int j = 1; j < 32; ++j) { c += dlt; dlt += sig * c; acc[j] += c; }
                    }
#pragma unroll
                    for (int j = 0; j < 32; ++j)
                        if (c0 + j < m.n_nodes) tab_w[(size_t)q * m.n_nodes + c0 + j] = acc[j];
                }
            }
        }
        __syncwarp();
        __threadfence_block();
        // ---- DP5
        double y[kSlots][2], k1[kSlots][2], k2[kSlots][2], k3[kSlots][2], k4[kSlots][2], k5[kSlots][2], k6[kSlots][2], k7[kSlots][2], un[kSlots][2],
            ut[kSlots][2];
#pragma unroll
        for (int s = 0; s < kSlots; ++s) { y[s][0] = st.rho0[s][lane][0]; y[s][1] = st.rho0[s][lane][1]; un[s][0] = y[s][0]; un[s][1] = y[s][1]; }
        const double t0 = P.tspan[0], tend = P.tspan[P.nt - 1];
        double t = t0, h = 0.0, tcommit = t0, dt = P.ode.dt0, lnq_old = log(1e-4);
        long long iters = 0;
        int status = 0;
        int n_rhs = 0, n_acc = 0, n_rej = 0;
        // k1 = f(t0, y)
        eval_coefficients(m, sm, w, lane, tab_w, t0, 0.0, t0);
        publish(y);
        rhs(kStages - 1, y, k1);
        n_rhs = 1;
        if (dt <= 0.0) {  // OrdinaryDiffEq initial step (Hairer), norms over all 625 entries
            double d0 = 0.0, d1 = 0.0;
#pragma unroll
            for (int s = 0; s < kSlots; ++s) {
                const double sk = atol + sqrt(y[s][0] * y[s][0] + y[s][1] * y[s][1]) * rtol;
                d0 += wgt[s] * (y[s][0] * y[s][0] + y[s][1] * y[s][1]) / (sk * sk);
                d1 += wgt[s] * (k1[s][0] * k1[s][0] + k1[s][1] * k1[s][1]) / (sk * sk);
            }
            d0 = sqrt(warp_sum2(d0) / 625.0); d1 = sqrt(warp_sum2(d1) / 625.0);
            double dt0 = (d0 < 1e-5 || d1 < 1e-5) ? 1e-6 : 0.01 * d0 / d1;
            dt0 = fmin(dt0, P.ode.dtmax);
#pragma unroll
            for (int s = 0; s < kSlots; ++s) { ut[s][0] = y[s][0] + dt0 * k1[s][0]; ut[s][1] = y[s][1] + dt0 * k1[s][1]; }
            eval_coefficients(m, sm, w, lane, tab_w, t0, 0.0, t0 + dt0);
            publish(ut);
            rhs(kStages - 1, ut, k2);
            n_rhs++;
            double d2 = 0.0;
#pragma unroll
            for (int s = 0; s < kSlots; ++s) {
                const double sk = atol + sqrt(y[s][0] * y[s][0] + y[s][1] * y[s][1]) * rtol;
                const double a = k2[s][0] - k1[s][0], b = k2[s][1] - k1[s][1];
                d2 += wgt[s] * (a * a + b * b) / (sk * sk);
            }
            d2 = sqrt(warp_sum2(d2) / 625.0) / dt0;
            const double mx = fmax(d1, d2);
            const double dt1 = (mx <= 1e-15) ? fmax(1e-6, dt0 * 1e-3) : pow(10.0, -(2.0 + log10(mx)) / 5.0);
            dt = fmin(100.0 * dt0, fmin(dt1, P.ode.dtmax));
        }
        for (int j = 0; j < P.nt; ++j) {
            const double T = P.tspan[j];
            const double tstop = P.ode.dense ? tend : T;
            while (status == 0 && tcommit < T) {
                if (h != 0.0) {
#pragma unroll
                    for (int s = 0; s < kSlots; ++s) { y[s][0] = un[s][0]; y[s][1] = un[s][1]; k1[s][0] = k7[s][0]; k1[s][1] = k7[s][1]; }
                    t = tcommit; h = 0.0;
                }
                if (++iters > P.ode.maxiters) { status = CA_ERR_MAXITERS; break; }
                double hh = fmin(dt, P.ode.dtmax);
                bool clipped = false;
                if (t + hh >= tstop || (tstop - (t + hh)) < 1e-13 * fabs(tstop)) { hh = tstop - t; clipped = true; }
                const double tn = clipped ? tstop : t + hh;
                eval_coefficients(m, sm, w, lane, tab_w, t, hh, tn);
#pragma unroll
                for (int s = 0; s < kSlots; ++s)
#pragma unroll
                    for (int c = 0; c < 2; ++c) ut[s][c] = y[s][c] + hh * (dp::a21 * k1[s][c]);
                publish(ut); rhs(0, ut, k2);
#pragma unroll
                for (int s = 0; s < kSlots; ++s)
#pragma unroll
                    for (int c = 0; c < 2; ++c) ut[s][c] = y[s][c] + hh * (dp::a31 * k1[s][c] + dp::a32 * k2[s][c]);
                publish(ut); rhs(1, ut, k3);
#pragma unroll
                for (int s = 0; s < kSlots; ++s)
#pragma unroll
                    for (int c = 0; c < 2; ++c) ut[s][c] = y[s][c] + hh * (dp::a41 * k1[s][c] + dp::a42 * k2[s][c] + dp::a43 * k3[s][c]);
                publish(ut); rhs(2, ut, k4);
#pragma unroll
                for (int s = 0; s < kSlots; ++s)
#pragma unroll
                    for (int c = 0; c < 2; ++c) ut[s][c] = y[s][c] + hh * (dp::a51 * k1[s][c] + dp::a52 * k2[s][c] + dp::a53 * k3[s][c] + dp::a54 * k4[s][c]);
                publish(ut); rhs(3, ut, k5);
#pragma unroll
                for (int s = 0; s < kSlots; ++s)
#pragma unroll
                    for (int c = 0; c < 2; ++c)
                        ut[s][c] = y[s][c] + hh * (dp::a61 * k1[s][c] + dp::a62 * k2[s][c] + dp::a63 * k3[s][c] + dp::a64 * k4[s][c] + dp::a65 * k5[s][c]);
                publish(ut); rhs(4, ut, k6);
#pragma unroll
                for (int s = 0; s < kSlots; ++s)
#pragma unroll
                    for (int c = 0; c < 2; ++c)
                        un[s][c] = y[s][c] + hh * (dp::b1 * k1[s][c] + dp::b3 * k3[s][c] + dp::b4 * k4[s][c] + dp::b5 * k5[s][c] + dp::b6 * k6[s][c]);
                publish(un); rhs(4, un, k7);
                n_rhs += 6;
                bool accept = true;
                if (P.ode.adaptive) {
                    double sacc = 0.0;
#pragma unroll
                    for (int s = 0; s < kSlots; ++s) {
                        const double er = hh * (dp::e1 * k1[s][0] + dp::e3 * k3[s][0] + dp::e4 * k4[s][0] + dp::e5 * k5[s][0] + dp::e6 * k6[s][0] + dp::e7 * k7[s][0]);
                        const double ei = hh * (dp::e1 * k1[s][1] + dp::e3 * k3[s][1] + dp::e4 * k4[s][1] + dp::e5 * k5[s][1] + dp::e6 * k6[s][1] + dp::e7 * k7[s][1]);
                        const double isc = rcp_d(atol + rtol * sqrt_d(fmax(y[s][0] * y[s][0] + y[s][1] * y[s][1], un[s][0] * un[s][0] + un[s][1] * un[s][1])));
                        sacc += wgt[s] * (er * er + ei * ei) * (isc * isc);
                    }
                    const double EEst2 = warp_sum2(sacc) * (1.0 / 625.0);
                    if (!(EEst2 == EEst2) || EEst2 > 1e300) { status = CA_ERR_NONFINITE; break; }
                    const double beta1 = 0.17, beta2 = 0.04, gam = 0.9, qmin = 0.2, qmax = 10.0, ln_qoldinit = -9.210340371976182;
                    double q, lnE = 0.0;
                    if (EEst2 == 0.0) q = 1.0 / qmax;
                    else { lnE = 0.5 * log(EEst2); q = exp_d(beta1 * lnE - beta2 * lnq_old) * (1.0 / gam); q = fmax(1.0 / qmax, fmin(1.0 / qmin, q)); }
                    if (EEst2 <= 1.0) {
                        lnq_old = (EEst2 == 0.0) ? ln_qoldinit : fmax(lnE, ln_qoldinit);
                        const double prop = hh * rcp_d(q);
                        dt = clipped ? fmax(dt, prop) : prop;
                    } else {
                        accept = false;
                        dt = hh * rcp_d(fmin(1.0 / qmin, exp_d(beta1 * lnE) * (1.0 / gam)));
                        n_rej++;
                        if (dt < 1e-14 * fmax(1.0, fabs(t))) { status = CA_ERR_NONFINITE; break; }
                    }
                } else dt = P.ode.dt0;
                if (accept) { n_acc++; h = hh; tcommit = tn; }
            }
            // ---- output at T (Hairer contd5 inside the last accepted step)
            double o[kSlots][2];
            if (h == 0.0 || T >= tcommit) {
#pragma unroll
                for (int s = 0; s < kSlots; ++s) { o[s][0] = un[s][0]; o[s][1] = un[s][1]; }
            } else {
                const double th = (T - t) / h, th1 = 1.0 - th;
#pragma unroll
                for (int s = 0; s < kSlots; ++s)
#pragma unroll
                    for (int c = 0; c < 2; ++c) {
                        const double r2 = un[s][c] - y[s][c], r3 = h * k1[s][c] - r2, r4 = r2 - h * k7[s][c] - r3;
                        const double r5 = h * (dp::d1 * k1[s][c] + dp::d3 * k3[s][c] + dp::d4 * k4[s][c] + dp::d5 * k5[s][c] + dp::d6 * k6[s][c] + dp::d7 * k7[s][c]);
                        o[s][c] = y[s][c] + th * (r2 + th1 * (r3 + th * (r4 + th1 * r5)));
                    }
            }
            if (status != 0) {
#pragma unroll
                for (int s = 0; s < kSlots; ++s) { o[s][0] = 0.0; o[s][1] = 0.0; }
            }
            publish(o);
            double* a0 = acc_w + (size_t)j * 2 * kSlots * 32 * 2;
#pragma unroll
            for (int s = 0; s < kSlots; ++s) {
                double qr = 0.0, qi = 0.0;
                if (m.rho2_mode == CA_RHO2_MATSQ) {  // (ρ·ρ)_IJ = Σ_K ρ_IK ρ_KJ
                    const int nq = st.n_sq[s];
                    for (int tq = 0; tq < nq; ++tq) {
                        const unsigned dsc = st.sq[s][tq][lane];
                        const double2 a = *reinterpret_cast<const double2*>(us[dsc >> 16]);
                        const double2 b = *reinterpret_cast<const double2*>(us[dsc & 0xffffu]);
                        qr += a.x * b.x - a.y * b.y; qi += a.x * b.y + a.y * b.x;
                    }
                } else qr = o[s][0] * o[s][0] + o[s][1] * o[s][1];
                double* p = a0 + ((size_t)s * 32 + lane) * 2;
                p[0] += o[s][0]; p[1] += o[s][1];
                double* p2 = p + kSlots * 32 * 2;
                p2[0] += qr; p2[1] += qi;
            }
            __syncwarp();
        }
        tot_rhs += n_rhs; tot_acc += n_acc; tot_rej += n_rej;
        if (status) worst = max(worst, -status);
    }
    tot_att = (long long)warp_sum2((double)tot_att);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) worst = max(worst, __shfl_xor_sync(0xffffffffu, worst, o));
    if (lane == 0) {
        atomicAdd(P.counters + 0, (unsigned long long)tot_rhs);
        atomicAdd(P.counters + 1, (unsigned long long)tot_acc);
        atomicAdd(P.counters + 2, (unsigned long long)tot_rej);
        atomicAdd(P.counters + 3, (unsigned long long)tot_att);
        if (worst) atomicMax(P.counters + 4, (unsigned long long)worst);
    }
}

// K6 for the two-atom path: reduce the per-warp accumulators in a fixed order and scatter the packed elements
// into full 25x25 column-major complex matrices (both triangles), scaled.
__global__ void k_finalize2(const double* __restrict__ accum, int n_warps, int nt, const Struct2* __restrict__ st, double scale,
                            double* __restrict__ rho, double* __restrict__ rho2) {
    const int e = blockIdx.x * blockDim.x + threadIdx.x;  // (j, which, slot, lane)
    const int per_t = 2 * kSlots * 32;
    if (e >= nt * per_t) return;
    const int j = e / per_t, r = e % per_t, which = r / (kSlots * 32), sl = r % (kSlots * 32), s = sl / 32, l = sl % 32;
    const int I = st->row[s][l], J = st->col[s][l];
    if (I < 0) return;
    double sr = 0.0, si = 0.0;
    for (int wv = 0; wv < n_warps; ++wv) {
        const double* p = accum + ((size_t)wv * nt + j) * per_t * 2 + (size_t)r * 2;
        sr += p[0]; si += p[1];
    }
    double* out = (which ? rho2 : rho) + (size_t)j * 625 * 2;
    out[2 * (I + 25 * J)] = sr * scale; out[2 * (I + 25 * J) + 1] = si * scale;
    if (I != J) { out[2 * (J + 25 * I)] = sr * scale; out[2 * (J + 25 * I) + 1] = -si * scale; }
}

}  // namespace ca

using namespace ca;

#define CA_CUDA2(expr)                                                                                      \
    do {                                                                                                    \
        cudaError_t _e = (expr);                                                                            \
        if (_e != cudaSuccess) { set_error(std::string(#expr) + ": " + cudaGetErrorString(_e)); return CA_ERR_CUDA; } \
    } while (0)

extern "C" int ca_rydberg2_run(ca_ctx* c, const ca_model* model, const double* tspan, int32_t nt, const double* psi0, int64_t n_traj,
                               int64_t traj_offset, int64_t n_total, uint64_t seed, const double* samples1, const double* samples2,
                               const double* const* phases4, const double* f, const double* amp_red, const double* amp_blue, int32_t nf,
                               const ca_ode_opts* ode, int32_t out_flags, double* rho, double* rho2, ca_stats* stats) {
    if (!c || !model || !tspan || nt < 1 || !psi0 || n_traj < 0 || !rho || !rho2) { set_error("bad argument"); return CA_ERR_ARG; }
    for (int j = 1; j < nt; ++j)
        if (!(tspan[j] > tspan[j - 1])) { set_error("tspan must be strictly increasing"); return CA_ERR_ARG; }
    CA_CUDA2(cudaSetDevice(ctx_device(c)));
    cudaStream_t s = ctx_stream(c);
    R2Params P;
    std::memset(&P, 0, sizeof P);
    int rc = derive_model(*model, nullptr, 0, tspan[0], tspan[nt - 1], &P.model);
    if (rc) return rc;
    if (P.model.red.kind != CA_BEAM_GAUSS && P.model.red.kind != CA_BEAM_UNIFORM) { /* any beam kind works; nothing to check */ }
    const bool noise = P.model.laser_noise != 0;
    if (noise) {
        if (!f || !amp_red || !amp_blue || nf < 1) { set_error("laser_noise=true needs f and amplitudes"); return CA_ERR_ARG; }
        if (tspan[0] < 0) { set_error("laser noise is tabulated on [0, tspan[end]]: tspan[0] must be >= 0"); return CA_ERR_ARG; }
    }
    fill_ode(ode, tspan[0], tspan[nt - 1], &P.ode);
    if (!P.ode.adaptive && !(P.ode.dt0 > 0)) { set_error("adaptive=false needs dt > 0"); return CA_ERR_ARG; }
    static thread_local Struct2 hs;
    if ((rc = build_struct2(psi0, P.model, &hs))) return rc;
    for (int sl = 0; sl < kSlots; ++sl)
        if (hs.n_terms[sl] > slot_terms(sl)) { set_error("two-atom kernel: gather table exceeds the compiled per-slot term counts"); return CA_ERR_UNSUPPORTED; }
    cudaEvent_t e0 = ctx_event(c, 0), e1 = ctx_event(c, 1), e2 = ctx_event(c, 2), e3 = ctx_event(c, 3);
    CA_CUDA2(cudaEventRecord(e0, s));
    // device buffers (ids 20..): struct, tspan, samples x2, phases x4, f/amps, scratch, accum, counters, out
    auto up = [&](int id, const void* host, size_t bytes, const void** dev) -> int {
        void* p = ctx_buf(c, id, bytes);
        if (!p) return CA_ERR_NOMEM;
        cudaError_t e = cudaMemcpyAsync(p, host, bytes, cudaMemcpyHostToDevice, s);
        if (e != cudaSuccess) { set_error(std::string("cudaMemcpyAsync: ") + cudaGetErrorString(e)); return CA_ERR_CUDA; }
        *dev = p;
        return CA_OK;
    };
    const void* d;
    if ((rc = up(20, &hs, sizeof hs, &d))) return rc; P.st = (const Struct2*)d;
    if ((rc = up(21, tspan, sizeof(double) * nt, &d))) return rc; P.tspan = (const double*)d;
    if (samples1 && n_traj > 0) { if ((rc = up(22, samples1, sizeof(double) * 6 * (size_t)n_traj, &d))) return rc; P.samples1 = (const double*)d; }
    if (samples2 && n_traj > 0) { if ((rc = up(23, samples2, sizeof(double) * 6 * (size_t)n_traj, &d))) return rc; P.samples2 = (const double*)d; }
    P.nf = 0;
    if (noise) {
        P.nf = nf;
        std::vector<double> fa((size_t)3 * nf);
        std::memcpy(fa.data(), f, sizeof(double) * nf);
        std::memcpy(fa.data() + nf, amp_red, sizeof(double) * nf);
        std::memcpy(fa.data() + 2 * (size_t)nf, amp_blue, sizeof(double) * nf);
        if ((rc = up(24, fa.data(), sizeof(double) * 3 * (size_t)nf, &d))) return rc;
        CA_CUDA2(cudaStreamSynchronize(s));
        P.f = (const double*)d; P.amp_red = P.f + nf; P.amp_blue = P.f + 2 * (size_t)nf;
        if (phases4 && n_traj > 0)
            for (int q = 0; q < 4; ++q)
                if (phases4[q]) { if ((rc = up(25 + q, phases4[q], sizeof(double) * (size_t)nf * n_traj, &d))) return rc; P.phases[q] = (const double*)d; }
    }
    CA_CUDA2(cudaStreamSynchronize(s));  // `hs` and host vectors are reused after return
    const int sms = ctx_sm_count(c);
    int grid = (int)std::max<long long>(1, std::min<long long>((n_traj + kWarps2 - 1) / kWarps2, sms));
    const int n_warps = grid * kWarps2;
    if (noise) { P.scratch = (double*)ctx_buf(c, 29, sizeof(double) * (size_t)n_warps * 4 * P.model.n_nodes); if (!P.scratch) return CA_ERR_NOMEM; }
    const size_t acc_bytes = sizeof(double) * (size_t)n_warps * nt * 2 * kSlots * 32 * 2;
    P.accum = (double*)ctx_buf(c, 30, acc_bytes);
    P.counters = (unsigned long long*)ctx_buf(c, 31, 64);
    if (!P.accum || !P.counters) return CA_ERR_NOMEM;
    CA_CUDA2(cudaMemsetAsync(P.accum, 0, acc_bytes, s));
    CA_CUDA2(cudaMemsetAsync(P.counters, 0, 64, s));
    P.n_traj = n_traj; P.traj_offset = traj_offset; P.seed = seed; P.nt = nt;
    CA_CUDA2(cudaFuncSetAttribute(k_lindblad2, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(Smem2)));
    CA_CUDA2(cudaEventRecord(e1, s));
    int launches = 0;
    if (n_traj > 0) {
        k_lindblad2<<<grid, kWarps2 * 32, sizeof(Smem2), s>>>(P);
        CA_CUDA2(cudaGetLastError());
        ++launches;
    }
    CA_CUDA2(cudaEventRecord(e2, s));
    const double div = (out_flags & CA_OUT_SUMS) ? 1.0 : 1.0 / (double)(n_total > 0 ? n_total : std::max<long long>(1, n_traj));
    const size_t out_doubles = (size_t)nt * 625 * 2;
    double *d_rho, *d_rho2;
    if (out_flags & CA_OUT_DEVICE) { d_rho = rho; d_rho2 = rho2; }
    else { d_rho = (double*)ctx_buf(c, 32, sizeof(double) * 2 * out_doubles); if (!d_rho) return CA_ERR_NOMEM; d_rho2 = d_rho + out_doubles; }
    CA_CUDA2(cudaMemsetAsync(d_rho, 0, sizeof(double) * out_doubles, s));
    CA_CUDA2(cudaMemsetAsync(d_rho2, 0, sizeof(double) * out_doubles, s));
    {
        const int tot = nt * 2 * kSlots * 32;
        k_finalize2<<<(tot + 127) / 128, 128, 0, s>>>(P.accum, n_traj > 0 ? n_warps : 0, nt, P.st, div, d_rho, d_rho2);
        CA_CUDA2(cudaGetLastError());
        ++launches;
    }
    CA_CUDA2(cudaEventRecord(e3, s));
    if (out_flags & CA_OUT_DEVICE) {
        if (stats) { std::memset(stats, 0, sizeof *stats); stats->n_traj = n_traj; stats->n_launches = launches; }
        return CA_OK;
    }
    unsigned long long hc[8];
    CA_CUDA2(cudaMemcpyAsync(rho, d_rho, sizeof(double) * out_doubles, cudaMemcpyDeviceToHost, s));
    CA_CUDA2(cudaMemcpyAsync(rho2, d_rho2, sizeof(double) * out_doubles, cudaMemcpyDeviceToHost, s));
    CA_CUDA2(cudaMemcpyAsync(hc, P.counters, 64, cudaMemcpyDeviceToHost, s));
    CA_CUDA2(cudaStreamSynchronize(s));
    const int status = hc[4] ? -(int)hc[4] : 0;
    if (stats) {
        std::memset(stats, 0, sizeof *stats);
        stats->n_traj = n_traj; stats->n_rhs = (int64_t)hc[0]; stats->n_accept = (int64_t)hc[1]; stats->n_reject = (int64_t)hc[2];
        stats->n_sample_attempts = (int64_t)hc[3]; stats->n_launches = launches; stats->status = status;
        float ms = 0;
        cudaEventElapsedTime(&ms, e1, e2); stats->kernel_ms = ms;
        cudaEventElapsedTime(&ms, e0, e3); stats->total_ms = ms;
    }
    if (status == CA_ERR_MAXITERS) { set_error("a trajectory hit maxiters"); return status; }
    if (status) { set_error("a trajectory produced a non-finite state or the step size underflowed"); return status; }
    return CA_OK;
}
