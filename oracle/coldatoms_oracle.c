/*
 * coldatoms_oracle.c -- CPU ORACLE (test infrastructure, NOT the product).
 *
 * A plain-C restatement of the Monte-Carlo master-equation hot path of mgoloshchapov/ColdAtoms.jl
 * (module NeutralAtoms).  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * `--impl reference` legs may call it; the product library (coldatoms.jl_b200/csrc) never does.
 *
 * PARITY UNPINNED at the master_dynamic boundary: the reference ships no tests (test/runtests.jl:1-6 is
 * commented out) and Julia is not available in the build container, so the only reference-held known
 * answers are the flat-top beam profiles of notebooks/modeling.ipynb cell 13 (tests/golden/).  Everything
 * else is pinned against scipy (DOP853 at rtol 1e-12) and analytic limits in tests/.
 *
 * The arithmetic of the path lives in third-party Julia packages that are not under /root/reference:
 *   QuantumOptics 1.2.2      timeevolution.master_dynamic -> dmaster_h!:
 *                            dρ = -i(Hρ - ρH) + Σ_j (J_j ρ J_j† - ½ J_j†J_j ρ - ½ ρ J_j†J_j),
 *                            defaults alg=DP5(), reltol=1e-6, abstol=1e-8, output at tspan through
 *                            FunctionCallingCallback(funcat=tspan) = dense interpolation, no tstops.
 *   OrdinaryDiffEq 6.93      DP5 (Dormand-Prince 5(4), FSAL), Hairer's dense output (contd5),
 *                            PI controller with DP5's beta2=0.04, beta1=0.17, gamma=0.9, qmin=0.2, qmax=10,
 *                            qoldinit=1e-4, automatic initial step (Hairer's hinit), maxiters=1e5.
 *   DiffEqBase 6.167         error norm sqrt(sum(abs2, err ./ (abstol + max(|u0|,|u1|) reltol))/length).
 *                            Its PI controller uses an approximate `fastpower`; we use exact pow, so step
 *                            sequences agree only to ~1e-4 relative -- far inside the solver tolerance.
 *   Interpolations 0.15      Gridded(Linear) on the 1001 noise nodes.
 * Each function below cites the reference file:line it follows.
 */
#include "coldatoms_oracle.h"

#include <complex.h>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

typedef double complex cplx;

#define CO_PI 3.14159265358979323846264338327950288
/* src/atom_sampler.jl:3-10 with PhysicalConstants.CODATA2018 k_B, m_u */
#define CO_KB 1.380649e-23
#define CO_MU 1.66053906660e-27

double co_vconst(void) { return sqrt(CO_KB * 1e-6 / CO_MU); }

/* ------------------------------------------------------------------------------------------------
 * Philox4x32-10 (Salmon et al. 2011) -- the counter-based RNG shared with the device sampler.
 * counter = (traj_lo, traj_hi, draw, stream), key = (seed_lo, seed_hi).
 * ---------------------------------------------------------------------------------------------- */
static void philox4x32_10(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4]) {
    uint32_t c0 = ctr[0], c1 = ctr[1], c2 = ctr[2], c3 = ctr[3];
    uint32_t k0 = key[0], k1 = key[1];
    for (int r = 0; r < 10; ++r) {
        uint64_t p0 = (uint64_t)0xD2511F53u * c0;
        uint64_t p1 = (uint64_t)0xCD9E8D57u * c2;
        uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0;
        uint32_t n1 = (uint32_t)p1;
        uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1;
        uint32_t n3 = (uint32_t)p0;
        c0 = n0; c1 = n1; c2 = n2; c3 = n3;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
    out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

void co_philox(uint64_t seed, uint64_t index, uint32_t draw, uint32_t stream, uint32_t out[4]) {
    uint32_t ctr[4] = {(uint32_t)index, (uint32_t)(index >> 32), draw, stream};
    uint32_t key[2] = {(uint32_t)seed, (uint32_t)(seed >> 32)};
    philox4x32_10(ctr, key, out);
}

static inline double u53(uint32_t a, uint32_t b) { /* [0,1) with 53 random bits */
    return ((double)(a >> 5) * 67108864.0 + (double)(b >> 6)) * (1.0 / 9007199254740992.0);
}
static inline double u53_open(uint32_t a, uint32_t b) { /* (0,1) */
    return ((double)(a >> 5) * 67108864.0 + (double)(b >> 6) + 0.5) * (1.0 / 9007199254740992.0);
}

/* ------------------------------------------------------------------------------------------------
 * Gaussian-beam primitives: src/utilities.jl:2-55
 * ---------------------------------------------------------------------------------------------- */
static double beam_w(double z, double w0, double z0) { return w0 * sqrt(1.0 + (z / z0) * (z / z0)); } /* :2-4 */

/* A(x,y,z,w0,z0; n=1, θ): utilities.jl:27-31 */
static double beam_A(double x, double y, double z, double w0, double z0, double theta) {
    double xt = sqrt(x * x + y * y), zt = z;
    double xr = xt * cos(theta) - zt * sin(theta);
    double zr = xt * sin(theta) + zt * cos(theta);
    double w = beam_w(zr, w0, z0);
    return (w0 / w) * exp(-((xr * xr) / (w * w)));
}
/* A_phase: utilities.jl:43-48 */
static cplx beam_A_phase(double x, double y, double z, double w0, double z0, double theta) {
    double xt = sqrt(x * x + y * y), zt = z;
    double xr = xt * cos(theta) - zt * sin(theta);
    double zr = xt * sin(theta) + zt * cos(theta);
    double k = 2.0 * z0 / (w0 * w0);
    return cexp(-1.0 * I * k * zr * (0.5 * (xr * xr) / (zr * zr + z0 * z0)) + 1.0 * I * atan(zr / z0));
}

/* src/arbitrary_beams.jl:16-22 : HG(x,w,n) = exp(-x^2/w^2) H_n(sqrt(2) x / w), physicists' Hermite */
static double hermite_phys(int n, double x) {
    double h0 = 1.0, h1 = 2.0 * x;
    if (n == 0) return h0;
    for (int k = 1; k < n; ++k) { double h2 = 2.0 * x * h1 - 2.0 * k * h0; h0 = h1; h1 = h2; }
    return h1;
}
static double HG(double x, double w, int n) { return exp(-x * x / (w * w)) * hermite_phys(n, sqrt(2.0) * x / w); }
/* :9-14 : LG(r2,w,k) = exp(-r2/w^2) L_k(2 r2/w^2)  (binomial(k,k) M(-k,1,x) = Laguerre L_k) */
static double laguerre(int n, double x) {
    double l0 = 1.0, l1 = 1.0 - x;
    if (n == 0) return l0;
    for (int k = 1; k < n; ++k) { double l2 = ((2.0 * k + 1.0 - x) * l1 - k * l0) / (k + 1.0); l0 = l1; l1 = l2; }
    return l1;
}
static double LG(double r2, double w, int k) { return exp(-r2 / (w * w)) * laguerre(k, 2.0 * r2 / (w * w)); }
static double factorial(int n) { double r = 1.0; for (int i = 2; i <= n; ++i) r *= i; return r; }
/* :24-30 */
double co_LG_coeff(int k, int n) {
    double res = 0.0;
    for (int i = k; i <= n; ++i) res += factorial(i) / (pow(2.0, i) * factorial(i - k));
    return ((k % 2) ? -1.0 : 1.0) * res / factorial(k);
}
/* :32-38 */
double co_HG_coeff(int k, int n) {
    double res = 0.0;
    for (int i = k; i <= n; ++i)
        res += factorial(2 * i) / (factorial(i) * pow(2.0, 3 * i) * factorial(i - k) * factorial(2 * k));
    return res;
}
/* :50-67 */
static cplx flattopHG_field(double x, double y, double z, const ca_beam* b) {
    double Om = b->Omega0, w0 = b->w0, zr = b->z0, sqz = b->sqz;
    int n = b->hg_n, m = b->hg_m;
    double w = w0 * sqrt(1.0 + (z / zr) * (z / zr));
    double k = 2.0 * zr / (w0 * w0);
    double phase0 = k * z + k / 2.0 * z * (x * x + y * y) / (z * z + zr * zr);
    cplx E = 0.0;
    for (int i = 0; i <= n; i += 2)
        for (int j = 0; j <= m; j += 2) {
            double gouy = (i + j + 1) * atan(z / zr);
            double cij = co_HG_coeff(i / 2, n / 2) * co_HG_coeff(j / 2, m / 2);
            E += cij * HG(x, w, i) * HG(y, w * sqz, j) * cexp(-1.0 * I * gouy);
        }
    return Om * w0 * E / w * cexp(-1.0 * I * phase0);
}
/* :69-81 */
static cplx flattopLG_field(double x, double y, double z, const ca_beam* b) {
    double Om = b->Omega0, w0 = b->w0, zr = b->z0;
    double w = w0 * sqrt(1.0 + (z / zr) * (z / zr));
    double k = 2.0 * zr / (w0 * w0);
    double phase0 = k * z + k / 2.0 * z * (x * x + y * y) / (z * z + zr * zr);
    cplx E = 0.0;
    int Nd2 = b->lg_l / 2;
    for (int i = 0; i <= Nd2; ++i)
        E += co_LG_coeff(i, Nd2) * LG(x * x + y * y, w, i) * cexp(-1.0 * I * (i + 1) * atan(z / zr));
    return Om * w0 / w * cexp(-1.0 * I * phase0) * E;
}

/* Ω dispatch: src/rydberg_model.jl:51-74 (5/6/7-param forms), Ω_red/Ω_blue: rydberg_model_arb_bms.jl:1-13 */
static cplx Omega_field(double x, double y, double z, const ca_beam* b) {
    switch (b->kind) {
        case CA_BEAM_GAUSS:
            return b->Omega0 * beam_A(x, y, z, b->w0, b->z0, b->theta) * beam_A_phase(x, y, z, b->w0, b->z0, b->theta);
        case CA_BEAM_FLATTOP_HG:
            return flattopHG_field(x, y, z, b);
        case CA_BEAM_FLATTOP_LG:
            return flattopLG_field(x, y, z, b);
        case CA_BEAM_UNIFORM:
        default:
            return b->Omega0;
    }
}

void co_eval_beam(const ca_beam* beam, const double* xyz, int64_t n, double* out) {
    for (int64_t i = 0; i < n; ++i) {
        cplx v = Omega_field(xyz[3 * i], xyz[3 * i + 1], xyz[3 * i + 2], beam);
        out[2 * i] = creal(v); out[2 * i + 1] = cimag(v);
    }
}

/* ------------------------------------------------------------------------------------------------
 * Atom sampler: src/atom_sampler.jl:16-50 (energies), :84-96 (harmonic branch, MvNormal + rejection)
 * ---------------------------------------------------------------------------------------------- */
static double energy_H(const double c[6], double U0, double w0, double z0, double m) { /* Π + K, :16-50 */
    double A = beam_A(c[0], c[1], c[2], w0, z0, 0.0);
    double vc = co_vconst();
    return U0 * (1.0 - A * A) + m / (vc * vc) * (c[3] * c[3] + c[4] * c[4] + c[5] * c[5]) / 2.0;
}
static double energy_H_harmonic(const double c[6], double U0, double w0, double z0, double m) { /* :24-31 */
    double vc = co_vconst();
    double r2 = c[0] * c[0] + c[1] * c[1];
    return U0 * (2.0 * r2 / (w0 * w0) + (c[2] / z0) * (c[2] / z0)) +
           m / (vc * vc) * (c[3] * c[3] + c[4] * c[4] + c[5] * c[5]) / 2.0;
}

void co_sampler_sigmas(const double trap[3], const double atom[2], double sig[6]) { /* :88 */
    double U0 = trap[0], w0 = trap[1], z0 = trap[2], m = atom[0], T = atom[1];
    double vc = co_vconst();
    sig[0] = sig[1] = sqrt(T * w0 * w0 / (4.0 * U0));
    sig[2] = sqrt(T * z0 * z0 / (2.0 * U0));
    sig[3] = sig[4] = sig[5] = sqrt(vc * vc * T / m);
}

/* One sample for global index `index`: attempts a = 0,1,..., three Philox calls per attempt
 * (Box-Muller pairs (x,y), (z,vx), (vy,vz)); accept iff H(cord) < U0(1-eps) with the Gaussian potential. */
int64_t co_sample_one(const double trap[3], const double atom[2], uint64_t seed, uint64_t index,
                      uint32_t stream, double eps, double out[6]) {
    double sig[6];
    co_sampler_sigmas(trap, atom, sig);
    for (uint32_t a = 0;; ++a) {
        double nrm[6];
        for (uint32_t c = 0; c < 3; ++c) {
            uint32_t w[4];
            co_philox(seed, index, a * 3u + c, stream, w);
            double u1 = u53_open(w[0], w[1]), u2 = u53(w[2], w[3]);
            double r = sqrt(-2.0 * log(u1));
            nrm[2 * c] = r * cos(2.0 * CO_PI * u2);
            nrm[2 * c + 1] = r * sin(2.0 * CO_PI * u2);
        }
        for (int i = 0; i < 6; ++i) out[i] = sig[i] * nrm[i];
        if (energy_H(out, trap[0], trap[1], trap[2], atom[0]) < trap[0] * (1.0 - eps)) return (int64_t)a + 1;
        if (a > 100000u) return -1;
    }
}

int64_t co_sample_atoms(const double* trap, const double* atom, int64_t n, int64_t offset, uint64_t seed,
                        int32_t stream, double eps, double* out) {
    int64_t attempts = 0;
    for (int64_t i = 0; i < n; ++i) attempts += co_sample_one(trap, atom, seed, (uint64_t)(offset + i), (uint32_t)stream, eps, out + 6 * i);
    return attempts;
}

/* Metropolis branch, src/atom_sampler.jl:97-120 (CPU-reference only).  One chain of N*freq+skip states,
 * thinned [1+skip:freq:end].  RNG: Philox stream `stream`, index = step. */
double co_sample_atoms_metropolis(const double* trap, const double* atom, int64_t N, int64_t freq, int64_t skip,
                                  int harmonic, double eps, uint64_t seed, int32_t stream, double* out) {
    double U0 = trap[0], w0 = trap[1], z0 = trap[2], m = atom[0], T = atom[1];
    double vc = co_vconst();
    double vstep = vc * sqrt(T / m), rstep = sqrt(T / U0) / 2.0;
    double sd[6] = {w0 * rstep, w0 * rstep, z0 * sqrt(2.0) * rstep, vstep, vstep, vstep};
    int64_t L = N * freq + skip;
    double* chain = (double*)malloc(sizeof(double) * 6 * (size_t)L);
    double cur[6] = {0, 0, 0, vstep / sqrt(3.0), vstep / sqrt(3.0), vstep / sqrt(3.0)};
    memcpy(chain, cur, sizeof cur);
    int64_t acc = 0;
    for (int64_t i = 1; i < L; ++i) {
        double nw[6];
        for (uint32_t c = 0; c < 3; ++c) {
            uint32_t w[4];
            co_philox(seed, (uint64_t)i, c, (uint32_t)stream, w);
            double u1 = u53_open(w[0], w[1]), u2 = u53(w[2], w[3]);
            double r = sqrt(-2.0 * log(u1));
            nw[2 * c] = cur[2 * c] + sd[2 * c] * r * cos(2.0 * CO_PI * u2);
            nw[2 * c + 1] = cur[2 * c + 1] + sd[2 * c + 1] * r * sin(2.0 * CO_PI * u2);
        }
        uint32_t w[4];
        co_philox(seed, (uint64_t)i, 3u, (uint32_t)stream, w);
        double u = u53(w[0], w[1]);
        double En = harmonic ? energy_H_harmonic(nw, U0, w0, z0, m) : energy_H(nw, U0, w0, z0, m);
        double Ec = harmonic ? energy_H_harmonic(cur, U0, w0, z0, m) : energy_H(cur, U0, w0, z0, m);
        double p_acc = exp(-En / T) / exp(-Ec / T);
        if (p_acc > u && En < U0 * (1.0 - eps)) { memcpy(cur, nw, sizeof cur); ++acc; }
        memcpy(chain + 6 * i, cur, sizeof cur);
    }
    int64_t k = 0;
    for (int64_t i = skip; i < L && k < N; i += freq, ++k) memcpy(out + 6 * k, chain + 6 * i, 6 * sizeof(double));
    free(chain);
    return (double)acc / (double)L;
}

/* The same Metropolis kernel run as n_chains independent chains (chain c emits samples c, c+K, ...): the parallel
 * decomposition the GPU path uses (ca_sample_atoms_metropolis); n_chains = 1 visits exactly the states of the single
 * reference chain.  RNG: Philox index = chain<<32 | transition.  Returns accepted / states visited. */
double co_sample_atoms_metropolis_chains(const double* trap, const double* atom, int64_t N, int64_t freq, int64_t skip,
                                         int harmonic, double eps, int64_t n_chains, uint64_t seed, int32_t stream, double* out) {
    double U0 = trap[0], w0 = trap[1], z0 = trap[2], m = atom[0], T = atom[1];
    double vc = co_vconst();
    double vstep = vc * sqrt(T / m), rstep = sqrt(T / U0) / 2.0;
    double sd[6] = {w0 * rstep, w0 * rstep, z0 * sqrt(2.0) * rstep, vstep, vstep, vstep};
    int64_t acc = 0, visited = 0;
    for (int64_t c = 0; c < n_chains; ++c) {
        int64_t M = c < N ? (N - c + n_chains - 1) / n_chains : 0;
        if (M == 0) continue;
        double cur[6] = {0, 0, 0, vstep / sqrt(3.0), vstep / sqrt(3.0), vstep / sqrt(3.0)};
        int64_t L = skip + (M - 1) * freq + 1, k = 0;
        ++visited;
        for (int64_t i = 0;; ++i) {
            if (i >= skip && (i - skip) % freq == 0) { memcpy(out + 6 * (c + k * n_chains), cur, sizeof cur); if (++k == M) break; }
            if (i + 1 >= L) break;
            uint64_t idx = ((uint64_t)c << 32) | (uint64_t)(i + 1);
            double nw[6];
            for (uint32_t q = 0; q < 3; ++q) {
                uint32_t w[4];
                co_philox(seed, idx, q, (uint32_t)stream, w);
                double u1 = u53_open(w[0], w[1]), u2 = u53(w[2], w[3]);
                double r = sqrt(-2.0 * log(u1));
                nw[2 * q] = cur[2 * q] + sd[2 * q] * (r * cos(2.0 * CO_PI * u2));
                nw[2 * q + 1] = cur[2 * q + 1] + sd[2 * q + 1] * (r * sin(2.0 * CO_PI * u2));
            }
            uint32_t w[4];
            co_philox(seed, idx, 3u, (uint32_t)stream, w);
            double u = u53(w[0], w[1]);
            double En = harmonic ? energy_H_harmonic(nw, U0, w0, z0, m) : energy_H(nw, U0, w0, z0, m);
            double Ec = harmonic ? energy_H_harmonic(cur, U0, w0, z0, m) : energy_H(cur, U0, w0, z0, m);
            double p_acc = exp(-En / T) / exp(-Ec / T);
            if (p_acc > u && En < U0 * (1.0 - eps)) { memcpy(cur, nw, sizeof cur); ++acc; }
            ++visited;
        }
    }
    return visited ? (double)acc / (double)visited : 0.0;
}

/* trap_frequencies: src/utilities.jl:73-79 */
void co_trap_frequencies(const double atom[2], const double trap[3], double* wr, double* wz) {
    double om = co_vconst() / trap[1] * sqrt(trap[0] / atom[0]);
    *wr = 2.0 * om; *wz = sqrt(2.0) * om;
}

/* R, V: src/atom_sampler.jl:125-143 */
static double traj_R(double t, double ri, double vi, double om, int free) {
    return free ? ri + vi * t : ri * cos(om * t) + vi / om * sin(om * t);
}
static double traj_V(double t, double ri, double vi, double om, int free) {
    return free ? vi : vi * cos(om * t) - ri * om * sin(om * t);
}

/* ------------------------------------------------------------------------------------------------
 * Laser phase noise: src/lasernoise_sampler.jl:16-37
 * ---------------------------------------------------------------------------------------------- */
void co_phi_amplitudes(const double* f, int nf, double h0, const double* hg, const double* sg, const double* fg, int ng, double* out) { /* :16-28 */
    double df = f[1] - f[0];
    for (int i = 0; i < nf; ++i) {
        double res = 2.0 * h0;
        for (int g = 0; g < ng; ++g) res += 2.0 * hg[g] * exp(-(f[i] - fg[g]) * (f[i] - fg[g]) / (2.0 * sg[g] * sg[g]));
        out[i] = 2.0 * sqrt(df * res) / f[i];
    }
}
void co_phase_draw(uint64_t seed, uint64_t index, uint32_t stream, int nf, double* phases) { /* :33: U[0,2π) */
    for (int k = 0; k < nf; k += 2) {
        uint32_t w[4];
        co_philox(seed, index, (uint32_t)(k >> 1), stream, w);
        phases[k] = 2.0 * CO_PI * u53(w[0], w[1]);
        if (k + 1 < nf) phases[k + 1] = 2.0 * CO_PI * u53(w[2], w[3]);
    }
}
void co_phi(const double* tn, int n_nodes, const double* f, const double* amp, const double* phases, int nf, double* out) { /* :34 */
    for (int j = 0; j < n_nodes; ++j) {
        double s = 0.0;
        for (int k = 0; k < nf; ++k) s += amp[k] * cos(2.0 * CO_PI * f[k] * tn[j] + phases[k]);
        out[j] = s;
    }
}
/* Gridded(Linear) on uniform knots t_j = j*dtn: src/rydberg_model.jl:166-167 */
static double interp_lin(const double* tab, int n_nodes, double dtn, double t) {
    if (!tab) return 0.0;
    int j = (int)floor(t / dtn);
    if (j < 0) j = 0;
    if (j > n_nodes - 2) j = n_nodes - 2;
    double w = (t - j * dtn) / dtn;
    return tab[j] * (1.0 - w) + tab[j + 1] * w;
}

/* ------------------------------------------------------------------------------------------------
 * Sparse operators on an N-level (or N x N-level) basis; column-major dense ρ.
 * ---------------------------------------------------------------------------------------------- */
#define MAXE 64
typedef struct { int n; int r[MAXE], c[MAXE]; cplx v[MAXE]; } spop;

static void sp_clear(spop* o) { o->n = 0; }
static void sp_add(spop* o, int r, int c, cplx v) { o->r[o->n] = r; o->c[o->n] = c; o->v[o->n] = v; o->n++; }
/* |a><b| on 5 levels */
static void sp_sigma(spop* o, int a, int b) { sp_clear(o); sp_add(o, a, b, 1.0); }
/* o1 ⊗ Id5 (first factor fastest: index = a1 + 5 a2) */
static void sp_kron_left(spop* out, const spop* o1) {
    sp_clear(out);
    for (int e = 0; e < o1->n; ++e) for (int k = 0; k < 5; ++k) sp_add(out, o1->r[e] + 5 * k, o1->c[e] + 5 * k, o1->v[e]);
}
static void sp_kron_right(spop* out, const spop* o2) {
    sp_clear(out);
    for (int e = 0; e < o2->n; ++e) for (int k = 0; k < 5; ++k) sp_add(out, k + 5 * o2->r[e], k + 5 * o2->c[e], o2->v[e]);
}

/* level indices, src/rydberg_model.jl:2-7 */
enum { L0 = 0, L1 = 1, LR = 2, LP = 3, LL = 4 };

/* ------------------------------------------------------------------------------------------------
 * Lindblad right-hand side in QuantumOptics' dmaster_h! form.
 * ---------------------------------------------------------------------------------------------- */
#define MAXH 17
#define MAXJ 8
typedef struct {
    int d;
    int nH; spop Hop[MAXH];            /* operators of the TimeDependentSum */
    int nJ; spop J[MAXJ];              /* jump operators (with sqrt(Γ) folded in) */
    double gam[25];                    /* diagonal of Σ J†J (all jumps are |a><b| (⊗Id): diagonal) */
    /* per-trajectory closures */
    const ca_model* m;
    int natoms;
    double s[2][6];                    /* initial conditions after atom_motion gating (+centre shift) */
    double wr, wz;
    const double* phi_tab[4];          /* red1, blue1, red2, blue2 noise tables or NULL */
    int n_nodes; double dtn;
    double tau;                        /* CZ phase-jump time */
    int64_t n_rhs;
    double shift[2][3];                /* CA_2A_SHIFT_AFTER_MOTION: beam-frame offsets added after R(t) (two_atom_model.jl:160-166) */
    int segment;                       /* 0 / 1: time segment of direct_CZ_simulation (two_atom_model.jl:304-305) */
} lind_sys;

static void coefficients(lind_sys* S, double t, cplx* c) {
    const ca_model* m = S->m;
    int fr = m->free_motion;
    const int fl = S->natoms == 2 ? m->two_atom_flags : 0;
    for (int k = 0; k < MAXH; ++k) c[k] = 0.0;
    for (int a = 0; a < S->natoms; ++a) {
        const double* s = S->s[a];
        /* legacy two_atom_simulation: atom 2 has its own beams and detunings (two_atom_model.jl:168-173) */
        const ca_beam* red = (a == 1 && (fl & CA_2A_PER_ATOM_LASERS)) ? &m->red2 : &m->red;
        const ca_beam* blue = (a == 1 && (fl & CA_2A_PER_ATOM_LASERS)) ? &m->blue2 : &m->blue;
        const double Delta0 = (a == 1 && (fl & CA_2A_PER_ATOM_LASERS)) ? m->Delta0_2 : m->Delta0;
        const double delta0 = (a == 1 && (fl & CA_2A_PER_ATOM_LASERS)) ? m->delta0_2 : m->delta0;
        /* get_atom_trajectories: src/rydberg_model.jl:28-43 (note Vz built from vy: line 41) */
        double X = traj_R(t, s[0], s[3], S->wr, fr), Y = traj_R(t, s[1], s[4], S->wr, fr), Z = traj_R(t, s[2], s[5], S->wz, fr);
        if (fl & CA_2A_SHIFT_AFTER_MOTION) { X += S->shift[a][0]; Y += S->shift[a][1]; Z += S->shift[a][2]; }
        double Vx = traj_V(t, s[0], s[3], S->wr, fr);
        double vz_src = (m->compat & CA_COMPAT_VZ_FROM_VY) ? s[4] : s[5];
        double Vz = traj_V(t, s[2], vz_src, S->wz, fr);
        /* the Doppler shifts of BOTH atoms are built from the first atom's laser parameters (two_atom_model.jl:158-159, :168-169) */
        double kr = 2.0 * m->red.z0 / (m->red.w0 * m->red.w0);
        double kb = 2.0 * m->blue.z0 / (m->blue.w0 * m->blue.w0);
        double Dd, dd;
        if (m->doppler_kind == CA_DOPPLER_XZ) { /* src/rydberg_model.jl:77-98 */
            Dd = kr * sin(m->red.theta) * Vx + kr * cos(m->red.theta) * Vz;
            dd = (kr * sin(m->red.theta) + kb * sin(m->blue.theta)) * Vx + (kr * cos(m->red.theta) + kb * cos(m->blue.theta)) * Vz;
        } else { /* legacy arb_bms: true vz, no angles */
            double Vzt = traj_V(t, s[2], s[5], S->wz, fr);
            Dd = kr * Vzt;
            dd = (m->parallel ? (kr + kb) : (kr - kb)) * Vzt;
        }
        double ph_r = interp_lin(S->phi_tab[2 * a], S->n_nodes, S->dtn, t);
        double ph_b = interp_lin(S->phi_tab[2 * a + 1], S->n_nodes, S->dtn, t);
        if (S->natoms == 2) ph_r += (t < S->tau) ? 0.0 : m->xi; /* cz_model.jl:136-137 */
        cplx* ca = c + 6 * a;
        if (fl & CA_2A_DIRECT) {   /* direct_CZ_simulation: Δ0 on nr, Ω(blue beam)/2 on σ1r / σr1, phase factor in segment 2 (:262-299) */
            cplx Od = cexp(1.0 * I * (ph_b + (S->segment ? m->seg2_phase : 0.0))) * Omega_field(X, Y, Z, blue);
            ca[1] = Delta0;
            c[13 + 2 * a] = Od / 2.0; c[14 + 2 * a] = conj(Od) / 2.0;
            continue;
        }
        cplx Or = cexp(1.0 * I * ph_r) * Omega_field(X, Y, Z, red);   /* rydberg_model.jl:175 */
        cplx Ob = cexp(1.0 * I * ph_b) * Omega_field(X, Y, Z, blue);  /* :176 */
        if (m->detuning_sign == CA_SIGN_RYDBERG1) { ca[0] = -Dd - Delta0; ca[1] = -dd - delta0; }   /* :180-181 */
        else { ca[0] = Dd - Delta0; ca[1] = dd - delta0; }                                        /* cz_model.jl:55-56 */
        ca[2] = Or / 2.0; ca[3] = conj(Or) / 2.0; ca[4] = Ob / 2.0; ca[5] = conj(Ob) / 2.0;             /* :182-185 */
    }
    if (S->natoms == 2) { /* get_V: cz_model.jl:12-17 */
        if (fl & CA_2A_CONST_VRR) c[12] = m->V_rr;   /* t -> V_rr: two_atom_model.jl:165 */
        else {
            double p[2][3];
            for (int a = 0; a < 2; ++a) {
                p[a][0] = traj_R(t, S->s[a][0], S->s[a][3], S->wr, fr); p[a][1] = traj_R(t, S->s[a][1], S->s[a][4], S->wr, fr);
                p[a][2] = traj_R(t, S->s[a][2], S->s[a][5], S->wz, fr);
                if (fl & CA_2A_SHIFT_AFTER_MOTION) for (int k = 0; k < 3; ++k) p[a][k] += S->shift[a][k];
            }
            double dx = p[0][0] - p[1][0], dy = p[0][1] - p[1][1], dz = p[0][2] - p[1][2];
            double r2 = dx * dx + dy * dy + dz * dz;
            c[12] = m->c6 / (1e-18 + r2 * r2 * r2);
        }
    }
}

static void lind_rhs_dense(lind_sys* S, double t, const cplx* rho, cplx* out);
static int g_dense_emulation = 0;
static void lind_rhs(lind_sys* S, double t, const cplx* rho, cplx* out) {
    if (g_dense_emulation) { lind_rhs_dense(S, t, rho, out); return; }
    int d = S->d, n = d * d;
    cplx c[MAXH];
    coefficients(S, t, c);
    S->n_rhs++;
    cplx tmp[625];
    for (int i = 0; i < n; ++i) tmp[i] = 0.0;
    for (int k = 0; k < S->nH; ++k) {
        const spop* o = &S->Hop[k];
        for (int e = 0; e < o->n; ++e) {
            cplx v = c[k] * o->v[e];
            int r = o->r[e], cc = o->c[e];
            for (int j = 0; j < d; ++j) tmp[r + d * j] += v * rho[cc + d * j];   /* Hρ */
            for (int i = 0; i < d; ++i) tmp[i + d * cc] -= rho[i + d * r] * v;   /* -ρH */
        }
    }
    for (int i = 0; i < n; ++i) out[i] = -1.0 * I * tmp[i];
    for (int j = 0; j < S->nJ; ++j) { /* J ρ J† */
        const spop* o = &S->J[j];
        for (int e1 = 0; e1 < o->n; ++e1)
            for (int e2 = 0; e2 < o->n; ++e2)
                out[o->r[e1] + d * o->r[e2]] += o->v[e1] * conj(o->v[e2]) * rho[o->c[e1] + d * o->c[e2]];
    }
    for (int j = 0; j < d; ++j) for (int i = 0; i < d; ++i) out[i + d * j] -= 0.5 * (S->gam[i] + S->gam[j]) * rho[i + d * j];
}

/* ------------------------------------------------------------------------------------------------
 * "dense emulation" of the reference's own arithmetic (timing only): the operators the
 * reference builds with `ket ⊗ dagger(ket)` are DENSE d x d matrices and QuantumOptics' dmaster_h! applies them with dense
 * `mul!` calls -- per RHS two per term of the TimeDependentSum (6 / 13 terms) and five per jump operator (4 / 8 jumps):
 * 32 products of 5x5 (one atom) or 66 products of 25x25 complex matrices (two atoms).  Same numbers as lind_rhs up to rounding.
 * ---------------------------------------------------------------------------------------------- */
void co_set_dense_emulation(int on) { g_dense_emulation = on; }

static void sp_to_dense(const spop* o, int d, cplx* M) {
    for (int i = 0; i < d * d; ++i) M[i] = 0.0;
    for (int e = 0; e < o->n; ++e) M[o->r[e] + d * o->c[e]] += o->v[e];
}
/* C = alpha A B + beta C, all d x d column-major (the generic mul! the dense Operators dispatch to) */
static void dense_mul(int d, cplx alpha, const cplx* A, const cplx* B, cplx beta, cplx* C) {
    for (int j = 0; j < d; ++j)
        for (int i = 0; i < d; ++i) {
            cplx acc = 0.0;
            for (int k = 0; k < d; ++k) acc += A[i + d * k] * B[k + d * j];
            C[i + d * j] = alpha * acc + (beta == 0.0 ? 0.0 : beta * C[i + d * j]);
        }
}
static void lind_rhs_dense(lind_sys* S, double t, const cplx* rho, cplx* out) {
    const int d = S->d, n = d * d;
    cplx c[MAXH];
    coefficients(S, t, c);
    S->n_rhs++;
    cplx *Hd = (cplx*)malloc(sizeof(cplx) * n * 3), *Jd = Hd + n, *tmp = Hd + 2 * n;
    for (int i = 0; i < n; ++i) out[i] = 0.0;
    for (int k = 0; k < S->nH; ++k) {   /* LazySum: every term applied on its own, from the left and from the right */
        sp_to_dense(&S->Hop[k], d, Hd);
        dense_mul(d, -1.0 * I * c[k], Hd, rho, 1.0, out);
        dense_mul(d, 1.0 * I * c[k], rho, Hd, 1.0, out);
    }
    for (int j = 0; j < S->nJ; ++j) {   /* dmaster_h!: J ρ J† - ½ J†J ρ - ½ ρ J†J as five products */
        sp_to_dense(&S->J[j], d, Jd);
        for (int a = 0; a < d; ++a) for (int b = 0; b < d; ++b) Hd[a + d * b] = conj(Jd[b + d * a]);   /* J† */
        dense_mul(d, 1.0, Jd, rho, 0.0, tmp);
        dense_mul(d, 1.0, tmp, Hd, 1.0, out);
        dense_mul(d, -0.5, Hd, tmp, 1.0, out);
        dense_mul(d, 1.0, rho, Hd, 0.0, tmp);
        dense_mul(d, -0.5, tmp, Jd, 1.0, out);
    }
    free(Hd);
}

static void build_operators(lind_sys* S, const ca_model* m, int natoms) {
    spop ops[6]; /* operators = [np, nr, σ1p, σp1, σpr, σrp]: src/rydberg_model.jl:25 */
    sp_sigma(&ops[0], LP, LP); sp_sigma(&ops[1], LR, LR); sp_sigma(&ops[2], L1, LP);
    sp_sigma(&ops[3], LP, L1); sp_sigma(&ops[4], LP, LR); sp_sigma(&ops[5], LR, LP);
    double G0 = m->decay_intermediate ? m->Gamma0 : 0.0, G1 = m->decay_intermediate ? m->Gamma1 : 0.0;
    double Gl = m->decay_intermediate ? m->Gammal : 0.0, Gr = m->decay_rydberg ? m->Gammar : 0.0; /* :221-223 */
    spop js[4]; /* JumpOperators: src/rydberg_model.jl:126-130 */
    sp_clear(&js[0]); sp_add(&js[0], L0, LP, sqrt(G0));
    sp_clear(&js[1]); sp_add(&js[1], L1, LP, sqrt(G1));
    sp_clear(&js[2]); sp_add(&js[2], LL, LP, sqrt(Gl));
    sp_clear(&js[3]); sp_add(&js[3], LL, LR, sqrt(Gr));
    S->m = m; S->natoms = natoms;
    if (natoms == 1) {
        S->d = 5; S->nH = 6; S->nJ = 4;
        for (int k = 0; k < 6; ++k) S->Hop[k] = ops[k];
        for (int k = 0; k < 4; ++k) S->J[k] = js[k];
    } else { /* cz_model.jl:31 and :2-9 */
        S->d = 25; S->nH = 17; S->nJ = 8;
        for (int k = 0; k < 6; ++k) { sp_kron_left(&S->Hop[k], &ops[k]); sp_kron_right(&S->Hop[6 + k], &ops[k]); }
        sp_clear(&S->Hop[12]); sp_add(&S->Hop[12], LR + 5 * LR, LR + 5 * LR, 1.0);
        /* direct_operators σgr, σrg per atom (two_atom_model.jl:38-44; g -> level 1); their coefficients are zero unless CA_2A_DIRECT */
        spop d1r, dr1;
        sp_sigma(&d1r, L1, LR); sp_sigma(&dr1, LR, L1);
        sp_kron_left(&S->Hop[13], &d1r); sp_kron_left(&S->Hop[14], &dr1); sp_kron_right(&S->Hop[15], &d1r); sp_kron_right(&S->Hop[16], &dr1);
        for (int k = 0; k < 4; ++k) { sp_kron_left(&S->J[k], &js[k]); sp_kron_right(&S->J[4 + k], &js[k]); }
    }
    for (int i = 0; i < S->d; ++i) S->gam[i] = 0.0;
    for (int j = 0; j < S->nJ; ++j)
        for (int e = 0; e < S->J[j].n; ++e) S->gam[S->J[j].c[e]] += creal(S->J[j].v[e] * conj(S->J[j].v[e]));
}

/* ------------------------------------------------------------------------------------------------
 * Dormand-Prince 5(4) as OrdinaryDiffEq.DP5 + DiffEqBase PI controller.
 * ---------------------------------------------------------------------------------------------- */
static const double A21 = 1.0 / 5.0, A31 = 3.0 / 40.0, A32 = 9.0 / 40.0, A41 = 44.0 / 45.0, A42 = -56.0 / 15.0, A43 = 32.0 / 9.0,
                    A51 = 19372.0 / 6561.0, A52 = -25360.0 / 2187.0, A53 = 64448.0 / 6561.0, A54 = -212.0 / 729.0,
                    A61 = 9017.0 / 3168.0, A62 = -355.0 / 33.0, A63 = 46732.0 / 5247.0, A64 = 49.0 / 176.0, A65 = -5103.0 / 18656.0,
                    A71 = 35.0 / 384.0, A73 = 500.0 / 1113.0, A74 = 125.0 / 192.0, A75 = -2187.0 / 6784.0, A76 = 11.0 / 84.0;
static const double C2 = 0.2, C3 = 0.3, C4 = 0.8, C5 = 8.0 / 9.0;
static const double E1 = 71.0 / 57600.0, E3 = -71.0 / 16695.0, E4 = 71.0 / 1920.0, E5 = -17253.0 / 339200.0, E6 = 22.0 / 525.0, E7 = -1.0 / 40.0;
static const double D1 = -12715105075.0 / 11282082432.0, D3 = 87487479700.0 / 32700410799.0, D4 = -10690763975.0 / 1880347072.0,
                    D5 = 701980252875.0 / 199316789632.0, D6 = -1453857185.0 / 822651844.0, D7 = 69997945.0 / 29380423.0;

static double rms_scaled(const cplx* e, const cplx* u0, const cplx* u1, int n, double atol, double rtol) {
    double s = 0.0;
    for (int i = 0; i < n; ++i) {
        double sc = atol + fmax(cabs(u0[i]), cabs(u1[i])) * rtol;
        double r = cabs(e[i]) / sc;
        s += r * r;
    }
    return sqrt(s / n);
}

/* integrate one trajectory; out_rho[nt*n] receives ρ(tspan[j]). returns 0 / CA_ERR_* */
static int dp5_solve(lind_sys* S, const cplx* u0, const double* tspan, int nt, const ca_ode_opts* oo, cplx* out_rho,
                     int64_t* n_acc, int64_t* n_rej) {
    int n = S->d * S->d;
    double rtol = oo->reltol > 0 ? oo->reltol : 1e-6, atol = oo->abstol > 0 ? oo->abstol : 1e-8;
    int64_t maxiters = oo->maxiters > 0 ? oo->maxiters : 100000;
    int adaptive = oo->adaptive, dense = oo->dense;
    double t0 = tspan[0], tend = tspan[nt - 1];
    double dtmax = oo->dtmax > 0 ? oo->dtmax : (tend - t0);
    cplx u[625], un[625], us[625], k1[625], k2[625], k3[625], k4[625], k5[625], k6[625], k7[625], er[625];
    memcpy(u, u0, sizeof(cplx) * n);
    memcpy(out_rho, u0, sizeof(cplx) * n);
    if (nt == 1) return 0;
    double t = t0;
    lind_rhs(S, t, u, k1);
    double dt = oo->dt;
    if (!adaptive && dt <= 0) return CA_ERR_ARG;
    if (dt <= 0) { /* OrdinaryDiffEq ode_determine_initdt (Hairer's hinit) */
        double d0 = 0, d1 = 0;
        for (int i = 0; i < n; ++i) { double sk = atol + cabs(u[i]) * rtol; d0 += pow(cabs(u[i]) / sk, 2); d1 += pow(cabs(k1[i]) / sk, 2); }
        d0 = sqrt(d0 / n); d1 = sqrt(d1 / n);
        double dt0 = (d0 < 1e-5 || d1 < 1e-5) ? 1e-6 : 0.01 * d0 / d1;
        dt0 = fmin(dt0, dtmax);
        for (int i = 0; i < n; ++i) us[i] = u[i] + dt0 * k1[i];
        lind_rhs(S, t + dt0, us, k2);
        double d2 = 0;
        for (int i = 0; i < n; ++i) { double sk = atol + cabs(u[i]) * rtol; d2 += pow(cabs(k2[i] - k1[i]) / sk, 2); }
        d2 = sqrt(d2 / n) / dt0;
        double mx = fmax(d1, d2);
        double dt1 = (mx <= 1e-15) ? fmax(1e-6, dt0 * 1e-3) : pow(10.0, -(2.0 + log10(mx)) / 5.0);
        dt = fmin(100.0 * dt0, fmin(dt1, dtmax));
    }
    const double beta1 = 0.17, beta2 = 0.04, gamma = 0.9, qmin = 0.2, qmax = 10.0;
    double qold = 1e-4;
    int jout = 1;
    int64_t iters = 0;
    while (jout < nt) {
        if (++iters > maxiters) return CA_ERR_MAXITERS;
        double tstop = dense ? tend : tspan[jout];
        double h = fmin(dt, dtmax);
        int clipped = 0;
        if (t + h >= tstop || (tstop - (t + h)) < 1e-13 * fabs(tstop)) { h = tstop - t; clipped = 1; }
        for (int i = 0; i < n; ++i) us[i] = u[i] + h * (A21 * k1[i]);
        lind_rhs(S, t + C2 * h, us, k2);
        for (int i = 0; i < n; ++i) us[i] = u[i] + h * (A31 * k1[i] + A32 * k2[i]);
        lind_rhs(S, t + C3 * h, us, k3);
        for (int i = 0; i < n; ++i) us[i] = u[i] + h * (A41 * k1[i] + A42 * k2[i] + A43 * k3[i]);
        lind_rhs(S, t + C4 * h, us, k4);
        for (int i = 0; i < n; ++i) us[i] = u[i] + h * (A51 * k1[i] + A52 * k2[i] + A53 * k3[i] + A54 * k4[i]);
        lind_rhs(S, t + C5 * h, us, k5);
        for (int i = 0; i < n; ++i) us[i] = u[i] + h * (A61 * k1[i] + A62 * k2[i] + A63 * k3[i] + A64 * k4[i] + A65 * k5[i]);
        lind_rhs(S, t + h, us, k6);
        for (int i = 0; i < n; ++i) un[i] = u[i] + h * (A71 * k1[i] + A73 * k3[i] + A74 * k4[i] + A75 * k5[i] + A76 * k6[i]);
        double tn = clipped ? tstop : t + h;
        lind_rhs(S, tn, un, k7);
        int accept = 1;
        if (adaptive) {
            for (int i = 0; i < n; ++i) er[i] = h * (E1 * k1[i] + E3 * k3[i] + E4 * k4[i] + E5 * k5[i] + E6 * k6[i] + E7 * k7[i]);
            double EEst = rms_scaled(er, u, un, n, atol, rtol);
            if (!(EEst == EEst) || isinf(EEst)) return CA_ERR_NONFINITE;
            double q, q11 = 0;
            if (EEst == 0.0) q = 1.0 / qmax;
            else { q11 = pow(EEst, beta1); q = q11 / pow(qold, beta2); q = fmax(1.0 / qmax, fmin(1.0 / qmin, q / gamma)); }
            if (EEst <= 1.0) { qold = fmax(EEst, 1e-4); dt = clipped ? fmax(dt, h / q) : h / q; }
            else { accept = 0; dt = h / fmin(1.0 / qmin, q11 / gamma); (*n_rej)++; }
            if (dt < 1e-14 * fmax(1.0, fabs(t))) return CA_ERR_NONFINITE;
        } else dt = oo->dt;
        if (!accept) continue;
        (*n_acc)++;
        /* outputs inside (t, tn] */
        while (jout < nt && tspan[jout] <= tn) {
            double T = tspan[jout];
            cplx* o = out_rho + (size_t)jout * n;
            if (T == tn) memcpy(o, un, sizeof(cplx) * n);
            else { /* Hairer contd5 = OrdinaryDiffEq DP5 dense output */
                double th = (T - t) / h, th1 = 1.0 - th;
                for (int i = 0; i < n; ++i) {
                    cplx r2 = un[i] - u[i];
                    cplx r3 = h * k1[i] - r2;
                    cplx r4 = r2 - h * k7[i] - r3;
                    cplx r5 = h * (D1 * k1[i] + D3 * k3[i] + D4 * k4[i] + D5 * k5[i] + D6 * k6[i] + D7 * k7[i]);
                    o[i] = u[i] + th * (r2 + th1 * (r3 + th * (r4 + th1 * r5)));
                }
            }
            ++jout;
        }
        t = tn;
        memcpy(u, un, sizeof(cplx) * n);
        memcpy(k1, k7, sizeof(cplx) * n);
    }
    return 0;
}

/* ------------------------------------------------------------------------------------------------
 * Ensemble drivers
 * ---------------------------------------------------------------------------------------------- */
static void accumulate(int d, int nt, const cplx* traj, int rho2_mode, double* acc, double* acc2) {
    int n = d * d;
    cplx sq[625];
    for (int j = 0; j < nt; ++j) {
        const cplx* r = traj + (size_t)j * n;
        if (rho2_mode == CA_RHO2_MATSQ) { /* `ρ2 .+= ρt .^ 2` = matrix square per time point (rydberg_model.jl:257) */
            for (int b = 0; b < d; ++b) for (int a = 0; a < d; ++a) {
                cplx s = 0.0;
                for (int k = 0; k < d; ++k) s += r[a + d * k] * r[k + d * b];
                sq[a + d * b] = s;
            }
        } else for (int i = 0; i < n; ++i) sq[i] = creal(r[i] * conj(r[i]));
        for (int i = 0; i < n; ++i) {
            acc[2 * ((size_t)j * n + i)] += creal(r[i]); acc[2 * ((size_t)j * n + i) + 1] += cimag(r[i]);
            acc2[2 * ((size_t)j * n + i)] += creal(sq[i]); acc2[2 * ((size_t)j * n + i) + 1] += cimag(sq[i]);
        }
    }
}

static int run_ensemble(int natoms, const ca_model* model, const double* tspan, int nt, int nt_seg1, const double* psi0, const double* rho0_in, int64_t n_traj,
                        int64_t traj_offset, int64_t n_total, uint64_t seed, const double* samples1, const double* samples2,
                        const double* const* phases, const double* f, const double* amp_red, const double* amp_blue, int nf,
                        const ca_ode_opts* ode, int out_flags, double* rho, double* rho2, ca_stats* stats, int nthreads) {
    int d = natoms == 1 ? 5 : 25, n = d * d;
    size_t outn = (size_t)2 * nt * n;
    ca_ode_opts oo;
    if (ode) oo = *ode; else { memset(&oo, 0, sizeof oo); oo.adaptive = 1; oo.dense = 1; }
    int n_nodes = model->n_noise_nodes > 0 ? model->n_noise_nodes : 1001;
    double tend = tspan[nt - 1];
    if (nt_seg1 > 0 && tspan[nt_seg1 - 1] > tend) tend = tspan[nt_seg1 - 1];   /* two segments: tables span the longer time axis */
    double dtn = tend / (n_nodes - 1); /* tspan_noise = 0:tend/1000:tend (rydberg_model.jl:216) */
    double* tn = (double*)malloc(sizeof(double) * n_nodes);
    for (int j = 0; j < n_nodes; ++j) tn[j] = j * dtn;
    /* 1-atom: noise only if laser_noise (rydberg_model.jl:161-171); CZ: amplitudes zeroed (cz_model.jl:134-135) */
    int use_noise = model->laser_noise && f && nf > 0;
    int ntr = 2 * natoms;
    cplx rho0[625];
    if (rho0_in) { for (int i = 0; i < n; ++i) rho0[i] = rho0_in[2 * i] + I * rho0_in[2 * i + 1]; }   /* ρ0 = copy(ρ_1): two_atom_model.jl:226 */
    else for (int b = 0; b < d; ++b) for (int a = 0; a < d; ++a)
        rho0[a + d * b] = (psi0[2 * a] + I * psi0[2 * a + 1]) * conj(psi0[2 * b] + I * psi0[2 * b + 1]);
    const int shift_after = natoms == 2 && (model->two_atom_flags & CA_2A_SHIFT_AFTER_MOTION);
#ifdef _OPENMP
    int T = nthreads > 0 ? nthreads : omp_get_max_threads();
#else
    int T = 1;
#endif
    double* accs = (double*)calloc((size_t)T * outn * 2, sizeof(double));
    int64_t tot_rhs = 0, tot_acc = 0, tot_rej = 0, tot_att = 0;
    int status = 0;
#pragma omp parallel num_threads(T) reduction(+ : tot_rhs, tot_acc, tot_rej, tot_att)
    {
#ifdef _OPENMP
        int tid = omp_get_thread_num();
#else
        int tid = 0;
#endif
        double* acc = accs + (size_t)tid * outn * 2;
        double* acc2 = acc + outn;
        lind_sys S;
        memset(&S, 0, sizeof S);
        build_operators(&S, model, natoms);
        double atom[2] = {model->atom_m, model->atom_T}, trap[3] = {model->trap_U0, model->trap_w0, model->trap_z0};
        co_trap_frequencies(atom, trap, &S.wr, &S.wz);
        S.n_nodes = n_nodes; S.dtn = dtn; S.tau = tend / 2.0;
        double* tabs = (double*)malloc(sizeof(double) * 4 * n_nodes);
        double* ph = (double*)malloc(sizeof(double) * (nf > 0 ? nf : 1));
        cplx* traj = (cplx*)malloc(sizeof(cplx) * (size_t)nt * n);
#pragma omp for schedule(static)
        for (int64_t i = 0; i < n_traj; ++i) {
            uint64_t gi = (uint64_t)(traj_offset + i);
            for (int a = 0; a < natoms; ++a) {
                const double* src = a == 0 ? samples1 : samples2;
                double smp[6];
                if (src) memcpy(smp, src + 6 * i, sizeof smp);
                else {
                    int64_t att = co_sample_one(trap, atom, seed, gi, (uint32_t)a, 1e-2, smp);
                    tot_att += att;
                }
                if (natoms == 2) {
                    const double* c = a == 0 ? model->center1 : model->center2;
                    if (shift_after) { S.shift[a][0] = c[0]; S.shift[a][1] = c[1]; S.shift[a][2] = c[2]; }
                    else { smp[0] += c[0]; smp[1] += c[1]; smp[2] += c[2]; }
                }
                for (int k = 0; k < 6; ++k) S.s[a][k] = model->atom_motion ? smp[k] : 0.0; /* rydberg_model.jl:35 */
            }
            for (int tr = 0; tr < 4; ++tr) S.phi_tab[tr] = NULL;
            if (use_noise)
                for (int tr = 0; tr < ntr; ++tr) {
                    if (phases && phases[tr]) memcpy(ph, phases[tr] + (size_t)nf * i, sizeof(double) * nf);
                    else co_phase_draw(seed, gi, (uint32_t)(2 + tr), nf, ph);
                    co_phi(tn, n_nodes, f, (tr & 1) ? amp_blue : amp_red, ph, nf, tabs + (size_t)tr * n_nodes);
                    S.phi_tab[tr] = tabs + (size_t)tr * n_nodes;
                }
            S.n_rhs = 0;
            int64_t na = 0, nr = 0;
            int rc;
            S.segment = 0;
            if (nt_seg1 > 0 && nt_seg1 < nt) {   /* two consecutive solves, the second from the final state of the first (two_atom_model.jl:304-305) */
                rc = dp5_solve(&S, rho0, tspan, nt_seg1, &oo, traj, &na, &nr);
                if (rc == 0) {
                    S.segment = 1;
                    cplx mid[625];
                    memcpy(mid, traj + (size_t)(nt_seg1 - 1) * n, sizeof(cplx) * n);
                    rc = dp5_solve(&S, mid, tspan + nt_seg1, nt - nt_seg1, &oo, traj + (size_t)nt_seg1 * n, &na, &nr);
                }
            } else rc = dp5_solve(&S, rho0, tspan, nt, &oo, traj, &na, &nr);
            if (rc != 0) {
#pragma omp critical
                status = rc;
            }
            tot_rhs += S.n_rhs; tot_acc += na; tot_rej += nr;
            accumulate(d, nt, traj, model->rho2_mode, acc, acc2);
        }
        free(tabs); free(ph); free(traj);
    }
    double div = (out_flags & CA_OUT_SUMS) ? 1.0 : (double)(n_total > 0 ? n_total : n_traj);
    for (size_t i = 0; i < outn; ++i) {
        double s = 0, s2 = 0;
        for (int t = 0; t < T; ++t) { s += accs[(size_t)t * outn * 2 + i]; s2 += accs[(size_t)t * outn * 2 + outn + i]; }
        rho[i] = s / div; rho2[i] = s2 / div;
    }
    free(accs); free(tn);
    if (stats) {
        memset(stats, 0, sizeof *stats);
        stats->n_traj = n_traj; stats->n_rhs = tot_rhs; stats->n_accept = tot_acc; stats->n_reject = tot_rej;
        stats->n_sample_attempts = tot_att; stats->status = status;
    }
    return status;
}

/* simulation / simulation_blue_intens: src/rydberg_model.jl:194-265, src/rydberg_model_arb_bms.jl:15-86 */
int co_rydberg1_run(const ca_model* model, const double* tspan, int32_t nt, const double* psi0, int64_t n_traj,
                    int64_t traj_offset, int64_t n_total, uint64_t seed, const double* samples, const double* phases_red,
                    const double* phases_blue, const double* f, const double* amp_red, const double* amp_blue, int32_t nf,
                    const ca_ode_opts* ode, int32_t out_flags, double* rho, double* rho2, ca_stats* stats, int32_t nthreads) {
    const double* ph[4] = {phases_red, phases_blue, NULL, NULL};
    return run_ensemble(1, model, tspan, nt, 0, psi0, NULL, n_traj, traj_offset, n_total, seed, samples, NULL, ph, f, amp_red, amp_blue, nf,
                        ode, out_flags, rho, rho2, stats, nthreads);
}

/* simulation_czlp: src/cz_model.jl:107-171 */
int co_rydberg2_run(const ca_model* model, const double* tspan, int32_t nt, const double* psi0, int64_t n_traj,
                    int64_t traj_offset, int64_t n_total, uint64_t seed, const double* samples1, const double* samples2,
                    const double* const* phases4, const double* f, const double* amp_red, const double* amp_blue, int32_t nf,
                    const ca_ode_opts* ode, int32_t out_flags, double* rho, double* rho2, ca_stats* stats, int32_t nthreads) {
    const double* ph[4] = {NULL, NULL, NULL, NULL};
    if (phases4) for (int k = 0; k < 4; ++k) ph[k] = phases4[k];
    /* CZ draws noise even when laser_noise=false but with zero amplitudes (cz_model.jl:134-135): same as no noise */
    return run_ensemble(2, model, tspan, nt, 0, psi0, NULL, n_traj, traj_offset, n_total, seed, samples1, samples2, ph, f, amp_red, amp_blue,
                        nf, ode, out_flags, rho, rho2, stats, nthreads);
}

/* release_evolve + release_recapture: src/basic_experiments.jl:25-47, :72-81 */
/* two_atom_simulation / direct_CZ_simulation (src/two_atom_model.jl:89-190, :192-319) and simulation_czlp with a general initial state:
 * rho0 (625 complex, column-major) replaces |psi0><psi0| when given; tspan[0..nt_seg1) and tspan[nt_seg1..nt) are two consecutive solves */
int co_rydberg2_run_ex(const ca_model* model, const double* tspan, int32_t nt, int32_t nt_seg1, const double* psi0, const double* rho0,
                       int64_t n_traj, int64_t traj_offset, int64_t n_total, uint64_t seed, const double* samples1, const double* samples2,
                       const double* const* phases4, const double* f, const double* amp_red, const double* amp_blue, int32_t nf,
                       const ca_ode_opts* ode, int32_t out_flags, double* rho, double* rho2, ca_stats* stats, int32_t nthreads) {
    const double* ph[4] = {NULL, NULL, NULL, NULL};
    if (phases4) for (int k = 0; k < 4; ++k) ph[k] = phases4[k];
    return run_ensemble(2, model, tspan, nt, nt_seg1, psi0, rho0, n_traj, traj_offset, n_total, seed, samples1, samples2, ph, f, amp_red, amp_blue,
                        nf, ode, out_flags, rho, rho2, stats, nthreads);
}

int co_release_recapture(const double* trap, const double* atom, const double* tspan, int32_t nt, int64_t n, int64_t traj_offset,
                         int64_t n_total, double eps, uint64_t seed, const double* samples, int32_t out_flags, double* probs,
                         uint8_t* indicators) {
    double U0 = trap[0], w0 = trap[1], z0 = trap[2], m = atom[0];
    double vc = co_vconst();
    for (int j = 0; j < nt; ++j) probs[j] = 0.0;
    for (int64_t i = 0; i < n; ++i) {
        double c[6];
        if (samples) memcpy(c, samples + 6 * i, sizeof c);
        else co_sample_one(trap, atom, seed, (uint64_t)(traj_offset + i), 0u, 1e-2, c); /* sampler eps stays 1e-2 (:73) */
        double kin = m / (vc * vc) * (c[3] * c[3] + c[4] * c[4] + c[5] * c[5]) / 2.0; /* K: atom_sampler.jl:35-39 */
        int alive = 1;
        for (int j = 0; j < nt; ++j) {
            double x = c[0] + c[3] * tspan[j], y = c[1] + c[4] * tspan[j], z = c[2] + c[5] * tspan[j]; /* :30-33 */
            double A = beam_A(x, y, z, w0, z0, 0.0);
            double pot = U0 * (1.0 - A * A);                                                          /* :36 */
            int rec = (kin + pot) < U0 * (1.0 - eps);                                                  /* :37 */
            alive = alive && rec;                                                                      /* :39-44 */
            if (indicators) indicators[(size_t)i * nt + j] = (uint8_t)alive;
            probs[j] += alive;
        }
    }
    if (!(out_flags & CA_OUT_SUMS)) { double div = (double)(n_total > 0 ? n_total : n); for (int j = 0; j < nt; ++j) probs[j] /= div; }
    return 0;
}

int co_num_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}
