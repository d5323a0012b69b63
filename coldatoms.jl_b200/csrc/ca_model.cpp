// ca_model.cpp -- host-side derivation of the device model (pure C++; also compiled into the test-only
// host-emulation harness under tests/hostemu, which is why it has no CUDA dependency).
#include <cmath>
#include <cstring>
#include <string>

#include "ca_device.cuh"
#include "ca_model.h"

namespace ca {

// ----------------------------------------------------------------------------------------------------
// error plumbing
// ----------------------------------------------------------------------------------------------------
static thread_local std::string g_err;
void set_error(const std::string& s) { g_err = s; }
const char* get_error() { return g_err.c_str(); }

// ----------------------------------------------------------------------------------------------------
// host-side model derivation
// ----------------------------------------------------------------------------------------------------
static double factorial(int n) { double r = 1.0; for (int i = 2; i <= n; ++i) r *= i; return r; }
static double LG_coeff(int k, int n) {  // arbitrary_beams.jl:24-30
    double res = 0.0;
    for (int i = k; i <= n; ++i) res += factorial(i) / (std::pow(2.0, i) * factorial(i - k));
    return ((k % 2) ? -1.0 : 1.0) * res / factorial(k);
}
static double HG_coeff(int k, int n) {  // arbitrary_beams.jl:32-38
    double res = 0.0;
    for (int i = k; i <= n; ++i) res += factorial(2 * i) / (factorial(i) * std::pow(2.0, 3 * i) * factorial(i - k) * factorial(2 * k));
    return res;
}

int derive_beam(const ca_beam& b, DevBeam* d) {
    std::memset(d, 0, sizeof *d);
    d->hO = 0.5 * b.Omega0;
    d->kind = b.kind;
    d->w0 = b.w0; d->z0 = b.z0;
    if (b.kind != CA_BEAM_UNIFORM) {
        if (!(b.w0 > 0.0) || !(b.z0 > 0.0)) { set_error("beam w0 and z0 must be positive"); return CA_ERR_ARG; }
        d->inv_z0 = 1.0 / b.z0; d->inv_w0sq = 1.0 / (b.w0 * b.w0); d->k = 2.0 * b.z0 / (b.w0 * b.w0);
    }
    // exact 0 / ±1 for the common angles so that the θ=0 fast path triggers and θ=π/2 has no 6e-17 residue
    // beyond what cos(π/2) gives in the reference (keep the reference's values: cos(pi/2) = 6.1e-17)
    d->cs = std::cos(b.theta); d->sn = std::sin(b.theta);
    d->sqz_inv = 1.0;
    if (b.kind == CA_BEAM_FLATTOP_HG) {
        if (b.hg_n < 0 || b.hg_m < 0 || b.hg_n / 2 >= kMaxModes || b.hg_m / 2 >= kMaxModes) { set_error("HG order out of range"); return CA_ERR_ARG; }
        if (!(b.sqz > 0.0)) { set_error("HG sqz must be positive"); return CA_ERR_ARG; }
        d->nx = b.hg_n / 2; d->ny = b.hg_m / 2; d->sqz_inv = 1.0 / b.sqz;
        for (int i = 0; i <= d->nx; ++i) d->cx[i] = HG_coeff(i, d->nx);
        for (int j = 0; j <= d->ny; ++j) d->cy[j] = HG_coeff(j, d->ny);
    } else if (b.kind == CA_BEAM_FLATTOP_LG) {
        if (b.lg_l < 0 || b.lg_l / 2 >= kMaxModes) { set_error("LG order out of range"); return CA_ERR_ARG; }
        d->nl = b.lg_l / 2;
        for (int p = 0; p <= d->nl; ++p) d->cx[p] = LG_coeff(p, d->nl);
    } else if (b.kind != CA_BEAM_GAUSS && b.kind != CA_BEAM_UNIFORM) {
        set_error("unknown beam kind"); return CA_ERR_ARG;
    }
    return CA_OK;
}

static const double kVconst = std::sqrt(1.380649e-23 * 1e-6 / 1.66053906660e-27);  // atom_sampler.jl:3-9

int derive_model(const ca_model& m, const double* psi0, int npsi, double t0, double tend, DevModel* d) {
    std::memset(d, 0, sizeof *d);
    int rc;
    if ((rc = derive_beam(m.red, &d->red))) return rc;
    if ((rc = derive_beam(m.blue, &d->blue))) return rc;
    if (!(m.atom_m > 0) || !(m.trap_U0 > 0) || !(m.trap_w0 > 0) || !(m.trap_z0 > 0) || m.atom_T < 0) { set_error("bad atom/trap params"); return CA_ERR_ARG; }
    double kr = m.red.kind == CA_BEAM_UNIFORM && !(m.red.w0 > 0) ? 0.0 : 2.0 * m.red.z0 / (m.red.w0 * m.red.w0);
    double kb = 2.0 * m.blue.z0 / (m.blue.w0 * m.blue.w0);
    if (m.doppler_kind == CA_DOPPLER_XZ) {  // rydberg_model.jl:77-98
        d->aDx = kr * std::sin(m.red.theta); d->aDz = kr * std::cos(m.red.theta);
        d->bDx = kr * std::sin(m.red.theta) + kb * std::sin(m.blue.theta);
        d->bDz = kr * std::cos(m.red.theta) + kb * std::cos(m.blue.theta);
        d->doppler_z_legacy = 0;
    } else {  // legacy arb_bms
        d->aDx = 0.0; d->bDx = 0.0; d->aDz = kr; d->bDz = m.parallel ? kr + kb : kr - kb;
        d->doppler_z_legacy = 1;
    }
    d->dsign = (m.detuning_sign == CA_SIGN_CZ) ? 1.0 : -1.0;
    d->Delta0 = m.Delta0; d->delta0 = m.delta0;
    d->G0 = m.decay_intermediate ? m.Gamma0 : 0.0; d->G1 = m.decay_intermediate ? m.Gamma1 : 0.0;
    d->Gl = m.decay_intermediate ? m.Gammal : 0.0; d->Gr = m.decay_rydberg ? m.Gammar : 0.0;
    if (d->G0 < 0 || d->G1 < 0 || d->Gl < 0 || d->Gr < 0) { set_error("negative decay rate"); return CA_ERR_ARG; }
    double om = kVconst / m.trap_w0 * std::sqrt(m.trap_U0 / m.atom_m);  // utilities.jl:73-79
    d->wr = 2.0 * om; d->wz = std::sqrt(2.0) * om;
    double T = m.atom_T, U0 = m.trap_U0, w0 = m.trap_w0, z0 = m.trap_z0;
    d->sig[0] = d->sig[1] = std::sqrt(T * w0 * w0 / (4.0 * U0));  // atom_sampler.jl:88
    d->sig[2] = std::sqrt(T * z0 * z0 / (2.0 * U0));
    d->sig[3] = d->sig[4] = d->sig[5] = std::sqrt(kVconst * kVconst * T / m.atom_m);
    d->U0 = U0; d->w0 = w0; d->z0 = z0; d->mfac = m.atom_m / (kVconst * kVconst);
    d->ecut = U0 * (1.0 - 1e-2);
    for (int k = 0; k < 3; ++k) { d->center[0][k] = m.center1[k]; d->center[1][k] = m.center2[k]; }
    d->c6 = m.c6; d->xi = m.xi; d->tau = tend / 2.0;
    d->n_nodes = m.n_noise_nodes > 0 ? m.n_noise_nodes : 1001;
    if (d->n_nodes < 2) { set_error("n_noise_nodes must be >= 2"); return CA_ERR_ARG; }
    d->dtn = tend / (d->n_nodes - 1); d->inv_dtn = d->dtn > 0 ? 1.0 / d->dtn : 0.0;
    d->atom_motion = m.atom_motion; d->free_motion = m.free_motion; d->laser_noise = m.laser_noise;
    d->vz_from_vy = (m.compat & CA_COMPAT_VZ_FROM_VY) ? 1 : 0;
    d->rho2_mode = m.rho2_mode;
    d->noise_coarse = 1;
    (void)t0;
    if (psi0 && npsi == 5) {
        double re[5], im[5];
        for (int a = 0; a < 5; ++a) { re[a] = psi0[2 * a]; im[a] = psi0[2 * a + 1]; }
        for (int a = 0; a < 5; ++a) d->rho0[a] = re[a] * re[a] + im[a] * im[a];
        int q = 0;
        for (int a = 0; a < 5; ++a)
            for (int b = a + 1; b < 5; ++b, ++q) {  // ρ_ab = ψ_a conj(ψ_b)
                d->rho0[5 + 2 * q] = re[a] * re[b] + im[a] * im[b];
                d->rho0[6 + 2 * q] = im[a] * re[b] - re[a] * im[b];
            }
        d->rows = ((re[0] != 0.0 || im[0] != 0.0) ? 1 : 0) | ((re[4] != 0.0 || im[4] != 0.0) ? 2 : 0);
    }
    return CA_OK;
}

void fill_ode(const ca_ode_opts* o, double t0, double tend, OdeOpts* d) {
    d->rtol = (o && o->reltol > 0) ? o->reltol : 1e-6;
    d->atol = (o && o->abstol > 0) ? o->abstol : 1e-8;
    d->dt0 = o ? o->dt : 0.0;
    d->dtmax = (o && o->dtmax > 0) ? o->dtmax : (tend - t0);
    d->maxiters = (o && o->maxiters > 0) ? o->maxiters : 100000;
    d->adaptive = o ? (o->adaptive != 0) : 1;
    d->dense = o ? (o->dense != 0) : 1;
}

}  // namespace ca
