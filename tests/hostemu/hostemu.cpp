// hostemu.cpp -- TEST-ONLY harness: runs the per-trajectory device code of coldatoms.jl_b200/csrc/ca_device.cuh on
// the CPU (the header is __host__ __device__), one "lane" at a time, so that the build container (no GPU) can
// check the exact source the sm_100a kernels execute against the oracle.  Never used by the product.
#include <cmath>
#include <cstring>
#include <vector>

#include <pthread.h>

#include "../../coldatoms.jl_b200/csrc/ca_device.cuh"
#include "../../coldatoms.jl_b200/csrc/ca_lindblad2.cuh"
#include "../../coldatoms.jl_b200/csrc/ca_lindblad2f.cuh"
#include "../../coldatoms.jl_b200/csrc/ca_model.h"

using namespace ca;

template <int NS>
static int run_lane(const DevModel& m, const OdeOpts& oo, const double* tspan, int nt, const double* smp, const double* tab, double* out,
                    long long* stats) {
    Rates rt;
    rt.G0 = m.G0; rt.G1 = m.G1; rt.Gl = m.Gl; rt.Gr = m.Gr; rt.Gp = m.G0 + m.G1 + m.Gl; rt.gp2 = 0.5 * rt.Gp; rt.gr2 = 0.5 * m.Gr;
    Lane1<NS> L;
    double kstore[6 * NS], cstore[30], astore[20];
    L.init(m, rt, oo, smp, tab, 1, tspan[0], kstore, 1, cstore, NS <= 15 ? astore : nullptr);
    const double tend = tspan[nt - 1];
    unsigned n_att = 0;
    for (int j = 0; j < nt; ++j) {
        const double T = tspan[j];
        const double tstop = oo.dense ? tend : T;
        while (L.status == 0 && L.t < T) { L.attempt(m, rt, oo, tstop, (n_att & (kAnchorEvery - 1)) == 0); ++n_att; }
        if (L.status) return L.status;
        double v[kNV];
        L.state_at(m, T, v);
        second_moment<(NS > 9)>(v, m.rho2_mode, v + 25);
        std::memcpy(out + (size_t)j * kNV, v, sizeof v);
    }
    stats[0] += L.n_rhs; stats[1] += L.n_acc; stats[2] += L.n_rej; stats[3] += L.n_exact;
    return 0;
}

// ---- two-atom kernel: the per-lane template of ca_lindblad2.cuh run by 32 host threads in lock-step ----
struct HostShared {
    pthread_barrier_t bar;
    double red[32][3];
    std::vector<double> mem;
};
struct HostWarp {
    int lane;
    double *xm, *xp, *kb;
    Coef2* coef;
    double* smp;
    double* anc;
    HostShared* sh;
    void sync() { pthread_barrier_wait(&sh->bar); }
    double sum(double v) {
        sh->red[lane][0] = v;
        sync();
        double s = 0.0;
        for (int l = 0; l < 32; ++l) s += sh->red[l][0];
        sync();
        return s;
    }
    void sum3(double& a, double& b, double& c) {
        sh->red[lane][0] = a; sh->red[lane][1] = b; sh->red[lane][2] = c;
        sync();
        double s0 = 0.0, s1 = 0.0, s2 = 0.0;
        for (int l = 0; l < 32; ++l) { s0 += sh->red[l][0]; s1 += sh->red[l][1]; s2 += sh->red[l][2]; }
        sync();
        a = s0; b = s1; c = s2;
    }
};
struct LaneArgs {
    HostWarp w;
    const DevModel* m; const OdeOpts* oo; const double* tspan; int nt; const double* tab; const double* rho0; double tr0; double* acc;
    Stats2 st;
};
static void* lane_main(void* p) {
    LaneArgs* a = static_cast<LaneArgs*>(p);
    Lane2 L;
    L.init(a->w.lane, *a->m);
    integrate2(a->w, L, *a->m, *a->oo, a->tspan, a->nt, a->tab, a->rho0, a->tr0, a->acc, a->st);
    return nullptr;
}

// ---- full-form two-atom kernel (ca_lindblad2f.cuh) run by 32 host threads in lock-step --------------------------------------------
struct HostWarpF {
    int lane;
    double *xm, *xp, *kb;
    double* coef;
    double* smp;
    HostShared* sh;
    void sync() { pthread_barrier_wait(&sh->bar); }
    double sum(double v) {
        sh->red[lane][0] = v;
        sync();
        double s = 0.0;
        for (int l = 0; l < 32; ++l) s += sh->red[l][0];
        sync();
        return s;
    }
};
struct LaneArgsF {
    HostWarpF w;
    const DevModel* m; const DevModel2x* x; const OdeOpts* oo; const double* tspan; int nt, nt_seg1; const double* tab; double* acc;
    Stats2 st;
};
static void* lane_main_f(void* p) {
    LaneArgsF* a = static_cast<LaneArgsF*>(p);
    a->st.n_rhs = a->st.n_acc = a->st.n_rej = a->st.n_exact = 0; a->st.status = 0;
    if (a->nt_seg1 > 0 && a->nt_seg1 < a->nt) {
        integrate2f(a->w, *a->m, *a->x, *a->oo, a->tspan, a->nt_seg1, 0, a->tab, a->acc, a->st);
        if (a->st.status == 0)
            integrate2f(a->w, *a->m, *a->x, *a->oo, a->tspan + a->nt_seg1, a->nt - a->nt_seg1, 1, a->tab, a->acc + (size_t)a->nt_seg1 * 2 * kF * 32, a->st);
    } else integrate2f(a->w, *a->m, *a->x, *a->oo, a->tspan, a->nt, 0, a->tab, a->acc, a->st);
    return nullptr;
}

extern "C" int he_rydberg2f_traj(const ca_model* model, const double* tspan, int nt, int nt_seg1, const double* psi0, const double* rho0,
                                 const double* sample1, const double* sample2, const double* const* phases4, const double* f,
                                 const double* amp_red, const double* amp_blue, int nf, const ca_ode_opts* ode, unsigned long long seed,
                                 unsigned long long gi, double* rho, double* rho2, long long* stats) {
    DevModel m;
    double tmax = tspan[nt - 1];
    if (nt_seg1 > 0 && tspan[nt_seg1 - 1] > tmax) tmax = tspan[nt_seg1 - 1];
    int rc = derive_model(*model, nullptr, 0, tspan[0], tmax, &m);
    if (rc) return rc;
    DevModel2x x;
    if ((rc = derive_model2x(*model, &x))) return rc;
    OdeOpts oo;
    const int n1 = nt_seg1 > 0 ? nt_seg1 : nt;
    fill_ode(ode, tspan[0], tspan[n1 - 1], &oo);
    if (nt_seg1 > 0 && !(ode && ode->dtmax > 0)) oo.dtmax = std::fmax(tspan[n1 - 1] - tspan[0], tspan[nt - 1] - tspan[nt_seg1]);
    std::vector<double> tab;
    const double* tp = nullptr;
    if (m.laser_noise && nf > 0) {
        tab.resize((size_t)4 * m.n_nodes);
        std::vector<NoiseK> kr(nf), kb(nf);
        fill_noise_k(f, amp_red, nf, m.dtn, kr.data());
        fill_noise_k(f, amp_blue, nf, m.dtn, kb.data());
        for (int q = 0; q < 4; ++q)
            gen_noise_trace<64>((q & 1) ? kb.data() : kr.data(), nf, m.n_nodes, m.dtn, 0.0, phases4 ? phases4[q] : nullptr, seed, gi, 2u + q,
                                tab.data() + (size_t)q * m.n_nodes, 1);
        tp = tab.data();
    }
    std::vector<double> rho0f(32 * kF);
    rho0_columns(psi0, rho0, rho0f.data());
    HostShared sh;
    pthread_barrier_init(&sh.bar, nullptr, 32);
    sh.mem.assign(kWarpDoublesF + 2, 0.0);
    double* base = sh.mem.data();
    double* smp = base + 2 * kFXN + kFB * kF * 32 + kStages2 * kCoefF;
    for (int a = 0; a < 2; ++a) {
        double s6[6] = {0, 0, 0, 0, 0, 0};
        const double* src = a == 0 ? sample1 : sample2;
        if (src) std::memcpy(s6, src, sizeof s6);
        else {
            SamplerConsts sc;
            for (int q = 0; q < 6; ++q) sc.sig[q] = m.sig[q];
            sc.U0 = m.U0; sc.w0 = m.w0; sc.z0 = m.z0; sc.mfac = m.mfac; sc.ecut = m.ecut;
            sample_atom_literal(sc, seed, gi, (uint32_t)a, s6);
        }
        if (!(x.flags & CA_2A_SHIFT_AFTER_MOTION)) for (int q = 0; q < 3; ++q) s6[q] += m.center[a][q];
        for (int q = 0; q < 6; ++q) smp[6 * a + q] = m.atom_motion ? s6[q] : 0.0;
    }
    std::vector<double> acc((size_t)nt * 2 * kF * 32, 0.0);
    std::vector<LaneArgsF> la(32);
    std::vector<pthread_t> th(32);
    for (int l = 0; l < 32; ++l) {
        HostWarpF& w = la[l].w;
        w.lane = l; w.xm = base; w.xp = base + kFXN; w.kb = base + 2 * kFXN + l;
        w.coef = base + 2 * kFXN + kFB * kF * 32;
        w.smp = smp;
        w.sh = &sh;
        for (int s = 0; s < kF; ++s) w.kb[(FB_Y * kF + s) * 32] = rho0f[l * kF + s];
        la[l].m = &m; la[l].x = &x; la[l].oo = &oo; la[l].tspan = tspan; la[l].nt = nt; la[l].nt_seg1 = nt_seg1; la[l].tab = tp;
        la[l].acc = acc.data() + l;
    }
    for (int l = 0; l < 32; ++l) pthread_create(&th[l], nullptr, lane_main_f, &la[l]);
    for (int l = 0; l < 32; ++l) pthread_join(th[l], nullptr);
    pthread_barrier_destroy(&sh.bar);
    for (int j = 0; j < nt; ++j)
        for (int which = 0; which < 2; ++which) {
            const double* a = acc.data() + ((size_t)j * 2 + which) * kF * 32;
            double* out = (which ? rho2 : rho) + (size_t)j * 1250;
            for (int J = 0; J < 25; ++J)
                for (int I = 0; I < 25; ++I) {
                    const double s = a[I * 32 + J], st = a[J * 32 + I];
                    out[2 * (I + 25 * J)] = 0.5 * (s + st);
                    out[2 * (I + 25 * J) + 1] = I == J ? 0.0 : 0.5 * (s - st);
                }
        }
    stats[0] = la[0].st.n_rhs; stats[1] = la[0].st.n_acc; stats[2] = la[0].st.n_rej; stats[3] = 0;
    return la[0].st.status;
}

extern "C" {

// One two-atom trajectory through the device source: full 25x25 column-major complex ρ(t_j), (ρ·ρ)(t_j).
int he_rydberg2_traj(const ca_model* model, const double* tspan, int nt, const double* psi0, const double* sample1, const double* sample2,
                     const double* const* phases4, const double* f, const double* amp_red, const double* amp_blue, int nf,
                     const ca_ode_opts* ode, unsigned long long seed, unsigned long long gi, double* rho, double* rho2, long long* stats) {
    DevModel m;
    int rc = derive_model(*model, nullptr, 0, tspan[0], tspan[nt - 1], &m);
    if (rc) return rc;
    OdeOpts oo;
    fill_ode(ode, tspan[0], tspan[nt - 1], &oo);
    std::vector<double> tab;
    const double* tp = nullptr;
    if (m.laser_noise && nf > 0) {
        tab.resize((size_t)4 * m.n_nodes);
        std::vector<NoiseK> kr(nf), kb(nf);
        fill_noise_k(f, amp_red, nf, m.dtn, kr.data());
        fill_noise_k(f, amp_blue, nf, m.dtn, kb.data());
        for (int q = 0; q < 4; ++q)
            gen_noise_trace<64>((q & 1) ? kb.data() : kr.data(), nf, m.n_nodes, m.dtn, 0.0, phases4 ? phases4[q] : nullptr, seed, gi, 2u + q,
                                tab.data() + (size_t)q * m.n_nodes, 1);
        tp = tab.data();
    }
    double rho0m[32 * kS2], tr0;
    if (!rho0_lanes(psi0, rho0m, &tr0)) return CA_ERR_UNSUPPORTED;
    HostShared sh;
    pthread_barrier_init(&sh.bar, nullptr, 32);
    sh.mem.assign(kWarpDoubles2 + 2, 0.0);
    double* base = sh.mem.data();
    for (int a = 0; a < 2; ++a) {
        double s6[6] = {0, 0, 0, 0, 0, 0};
        const double* src = a == 0 ? sample1 : sample2;
        if (src) std::memcpy(s6, src, sizeof s6);
        else {
            SamplerConsts sc;
            for (int q = 0; q < 6; ++q) sc.sig[q] = m.sig[q];
            sc.U0 = m.U0; sc.w0 = m.w0; sc.z0 = m.z0; sc.mfac = m.mfac; sc.ecut = m.ecut;
            sample_atom_literal(sc, seed, gi, (uint32_t)a, s6);
        }
        for (int q = 0; q < 3; ++q) s6[q] += m.center[a][q];
        for (int q = 0; q < 6; ++q) base[2 * kXN + kKB * kS2 * 32 + kStages2 * 16 + 6 * a + q] = m.atom_motion ? s6[q] : 0.0;
    }
    std::vector<double> acc((size_t)nt * 2 * kAccSlots * 32, 0.0);
    std::vector<LaneArgs> la(32);
    std::vector<pthread_t> th(32);
    for (int l = 0; l < 32; ++l) {
        HostWarp& w = la[l].w;
        w.lane = l; w.xm = base; w.xp = base + kXN; w.kb = base + 2 * kXN + l;
        w.coef = reinterpret_cast<Coef2*>(base + 2 * kXN + kKB * kS2 * 32);
        w.smp = base + 2 * kXN + kKB * kS2 * 32 + kStages2 * 16;
        w.anc = w.smp + 12 + l;
        w.sh = &sh;
        la[l].m = &m; la[l].oo = &oo; la[l].tspan = tspan; la[l].nt = nt; la[l].tab = tp; la[l].rho0 = rho0m + l * kS2; la[l].tr0 = tr0;
        la[l].acc = acc.data() + l;
        pthread_create(&th[l], nullptr, lane_main, &la[l]);
    }
    for (int l = 0; l < 32; ++l) pthread_join(th[l], nullptr);
    pthread_barrier_destroy(&sh.bar);
    std::memset(rho, 0, sizeof(double) * nt * 1250);
    std::memset(rho2, 0, sizeof(double) * nt * 1250);
    for (int j = 0; j < nt; ++j)
        for (int which = 0; which < 2; ++which)
            for (int s = 0; s < kAccSlots; ++s)
                for (int l = 0; l < 32; ++l) {
                    int I, J, tl = 0, ts = 9;
                    if (s < 9) slot_entry(l, s, &I, &J, &tl, &ts);
                    const double* a = acc.data() + (size_t)(j * 2 + which) * kAccSlots * 32;
                    scatter_entry(l, s, a[s * 32 + l], a[ts * 32 + tl], (which ? rho2 : rho) + (size_t)j * 1250);
                }
    stats[0] = la[0].st.n_rhs; stats[1] = la[0].st.n_acc; stats[2] = la[0].st.n_rej;
    stats[3] = 0;
    for (int l = 0; l < 32; ++l) stats[3] += la[l].st.n_exact;
    return la[0].st.status;
}

// One trajectory: packed (ρ, ρ²) per time point into out[nt][50].
int he_rydberg1_traj(const ca_model* model, const double* tspan, int nt, const double* psi0, const double* sample, const double* phases_red,
                     const double* phases_blue, const double* f, const double* amp_red, const double* amp_blue, int nf, const ca_ode_opts* ode,
                     unsigned long long seed, unsigned long long gi, double* out, long long* stats) {
    DevModel m;
    int rc = derive_model(*model, psi0, 5, tspan[0], tspan[nt - 1], &m);
    if (rc) return rc;
    OdeOpts oo;
    fill_ode(ode, tspan[0], tspan[nt - 1], &oo);
    std::vector<double> tab;
    const double* tp = nullptr;
    if (m.laser_noise && nf > 0) {
        tab.resize((size_t)2 * m.n_nodes);
        std::vector<NoiseK> kr(nf), kb(nf);
        const int cz = choose_noise_coarse(f, amp_red, amp_blue, nf, m.dtn, m.n_nodes);   // same plan as run_r1
        std::vector<double> tmp(noise_coarse_count(m.n_nodes, cz));
        fill_noise_k(f, amp_red, nf, noise_c(cz) * m.dtn, kr.data());
        fill_noise_k(f, amp_blue, nf, noise_c(cz) * m.dtn, kb.data());
        gen_noise_trace_c<64>(kr.data(), nf, m.n_nodes, m.dtn, cz, phases_red, seed, gi, 2u, tab.data(), 1, tmp.data(), 1);
        gen_noise_trace_c<64>(kb.data(), nf, m.n_nodes, m.dtn, cz, phases_blue, seed, gi, 3u, tab.data() + m.n_nodes, 1, tmp.data(), 1);
        tp = tab.data();
    }
    double smp[6] = {0, 0, 0, 0, 0, 0};
    if (sample) std::memcpy(smp, sample, sizeof smp);
    else {
        SamplerConsts sc;
        for (int q = 0; q < 6; ++q) sc.sig[q] = m.sig[q];
        sc.U0 = m.U0; sc.w0 = m.w0; sc.z0 = m.z0; sc.mfac = m.mfac; sc.ecut = m.ecut;
        sample_atom_literal(sc, seed, gi, 0u, smp);
    }
    stats[0] = stats[1] = stats[2] = stats[3] = 0;
    if (m.rows == 0) return run_lane<9>(m, oo, tspan, nt, smp, tp, out, stats);
    if (m.rows == 3) return run_lane<21>(m, oo, tspan, nt, smp, tp, out, stats);
    return run_lane<15>(m, oo, tspan, nt, smp, tp, out, stats);
}

int he_noise_trace(const double* f, const double* amp, int nf, int n_nodes, double dtn, const double* phases, unsigned long long seed,
                   unsigned long long gi, unsigned stream, double* out) {
    std::vector<NoiseK> kk(nf);
    const int cz = choose_noise_coarse(f, amp, nullptr, nf, dtn, n_nodes);
    std::vector<double> tmp(noise_coarse_count(n_nodes, cz));
    fill_noise_k(f, amp, nf, noise_c(cz) * dtn, kk.data());
    gen_noise_trace_c<64>(kk.data(), nf, n_nodes, dtn, cz, phases, seed, gi, stream, out, 1, tmp.data(), 1);
    return 0;
}

int he_sample_atom(const double* trap, const double* atom, unsigned long long seed, unsigned long long gi, unsigned stream, double eps, double* out) {
    ca_model mm;
    std::memset(&mm, 0, sizeof mm);
    mm.atom_m = atom[0]; mm.atom_T = atom[1]; mm.trap_U0 = trap[0]; mm.trap_w0 = trap[1]; mm.trap_z0 = trap[2];
    mm.red.kind = mm.blue.kind = CA_BEAM_UNIFORM; mm.blue.w0 = mm.blue.z0 = 1.0;
    DevModel m;
    int rc = derive_model(mm, nullptr, 0, 0.0, 1.0, &m);
    if (rc) return rc;
    SamplerConsts sc;
    for (int q = 0; q < 6; ++q) sc.sig[q] = m.sig[q];
    sc.U0 = m.U0; sc.w0 = m.w0; sc.z0 = m.z0; sc.mfac = m.mfac; sc.ecut = trap[0] * (1.0 - eps);
    return sample_atom_literal(sc, seed, gi, stream, out);
}

int he_eval_beam(const ca_beam* beam, const double* xyz, long long n, double* out) {
    DevBeam b;
    int rc = derive_beam(*beam, &b);
    if (rc) return rc;
    for (long long i = 0; i < n; ++i) {
        double re, im;
        beam_half_omega(b, xyz[3 * i], xyz[3 * i + 1], xyz[3 * i + 2], 0.0, &re, &im);
        out[2 * i] = 2.0 * re; out[2 * i + 1] = 2.0 * im;
    }
    return 0;
}

// multi-chain Metropolis sampler through the device source (all chains run serially here)
double he_metropolis(const double* trap, const double* atom, long long N, long long freq, long long skip, int harmonic, double eps,
                     long long n_chains, unsigned long long seed, unsigned stream, double* out) {
    MetroConsts mc;
    const double U0 = trap[0], w0 = trap[1], z0 = trap[2], m = atom[0], T = atom[1];
    const double vc = std::sqrt(1.380649e-23 * 1e-6 / 1.66053906660e-27);
    const double vstep = vc * std::sqrt(T / m), rstep = std::sqrt(T / U0) / 2.0;
    mc.sd[0] = mc.sd[1] = w0 * rstep; mc.sd[2] = z0 * std::sqrt(2.0) * rstep; mc.sd[3] = mc.sd[4] = mc.sd[5] = vstep;
    mc.U0 = U0; mc.w0 = w0; mc.z0 = z0; mc.mfac = m / (vc * vc); mc.T = T; mc.ecut = U0 * (1.0 - eps); mc.harmonic = harmonic;
    long long acc = 0, visited = 0;
    for (long long c = 0; c < n_chains; ++c) {
        long long steps = 0;
        acc += metropolis_chain(mc, seed, stream, c, n_chains, N, freq, skip, out, &steps);
        if (c < N) visited += steps + 1;
    }
    return visited ? (double)acc / (double)visited : 0.0;
}

// thermal-average functionals of ca_device.cuh (thermal_terms) for n host-supplied samples: out[n][2]
int he_thermal_terms(int kind, const ca_model* model, const double* s1, const double* s2, long long n, double* out) {
    DevModel d;
    int rc = derive_model(*model, nullptr, 0, 0.0, 1.0, &d);
    if (rc) return rc;
    AvgConsts ac;
    ac.red = d.red; ac.blue = d.blue; ac.aDx = d.aDx; ac.aDz = d.aDz; ac.kind = kind;
    std::memcpy(ac.center, d.center, sizeof ac.center);
    const double zero[6] = {0, 0, 0, 0, 0, 0};
    for (long long i = 0; i < n; ++i) thermal_terms(ac, s1 + 6 * i, s2 ? s2 + 6 * i : zero, out + 2 * i);
    return 0;
}

const char* he_last_error() { return get_error(); }
}
