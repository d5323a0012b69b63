// hostemu.cpp -- TEST-ONLY harness: runs the per-trajectory device code of coldatoms.jl_b200/csrc/ca_device.cuh on
// the CPU (the header is __host__ __device__), one "lane" at a time, so that the build container (no GPU) can
// check the exact source the sm_100a kernels execute against the oracle.  Never used by the product.
#include <cmath>
#include <cstring>
#include <vector>

#include "../../coldatoms.jl_b200/csrc/ca_device.cuh"
#include "../../coldatoms.jl_b200/csrc/ca_model.h"

using namespace ca;

template <int NS>
static int run_lane(const DevModel& m, const OdeOpts& oo, const double* tspan, int nt, const double* smp, const double* tab, double* out,
                    long long* stats) {
    Rates rt;
    rt.G0 = m.G0; rt.G1 = m.G1; rt.Gl = m.Gl; rt.Gr = m.Gr; rt.Gp = m.G0 + m.G1 + m.Gl; rt.gp2 = 0.5 * rt.Gp; rt.gr2 = 0.5 * m.Gr;
    Lane1<NS> L;
    double kstore[6 * NS], cstore[30], astore[20];
    L.init(m, rt, oo, smp, tab, 1, tspan[0], kstore, 1, cstore, NS == 9 ? astore : nullptr);
    const double tend = tspan[nt - 1];
    for (int j = 0; j < nt; ++j) {
        const double T = tspan[j];
        const double tstop = oo.dense ? tend : T;
        while (L.status == 0 && L.tcommit < T) L.attempt(m, rt, oo, tstop);
        if (L.status) return L.status;
        double v[kNV];
        L.state_at(m, T, v);
        second_moment<(NS > 9)>(v, m.rho2_mode, v + 25);
        std::memcpy(out + (size_t)j * kNV, v, sizeof v);
    }
    stats[0] += L.n_rhs; stats[1] += L.n_acc; stats[2] += L.n_rej;
    return 0;
}

extern "C" {

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
        fill_noise_k(f, amp_red, nf, m.dtn, kr.data());
        fill_noise_k(f, amp_blue, nf, m.dtn, kb.data());
        gen_noise_trace<64>(kr.data(), nf, m.n_nodes, m.dtn, phases_red, seed, gi, 2u, tab.data(), 1);
        gen_noise_trace<64>(kb.data(), nf, m.n_nodes, m.dtn, phases_blue, seed, gi, 3u, tab.data() + m.n_nodes, 1);
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
    stats[0] = stats[1] = stats[2] = 0;
    if (m.rows == 0) return run_lane<9>(m, oo, tspan, nt, smp, tp, out, stats);
    if (m.rows == 3) return run_lane<21>(m, oo, tspan, nt, smp, tp, out, stats);
    return run_lane<15>(m, oo, tspan, nt, smp, tp, out, stats);
}

int he_noise_trace(const double* f, const double* amp, int nf, int n_nodes, double dtn, const double* phases, unsigned long long seed,
                   unsigned long long gi, unsigned stream, double* out) {
    std::vector<NoiseK> kk(nf);
    fill_noise_k(f, amp, nf, dtn, kk.data());
    gen_noise_trace<64>(kk.data(), nf, n_nodes, dtn, phases, seed, gi, stream, out, 1);
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

const char* he_last_error() { return get_error(); }
}
