/* coldatoms_oracle.h -- CPU oracle (test infrastructure only; see coldatoms_oracle.c). */
#ifndef COLDATOMS_ORACLE_H
#define COLDATOMS_ORACLE_H
#include "../include/coldatoms_b200.h"
#ifdef __cplusplus
extern "C" {
#endif
double co_vconst(void);
void co_philox(uint64_t seed, uint64_t index, uint32_t draw, uint32_t stream, uint32_t out[4]);
double co_LG_coeff(int k, int n);
double co_HG_coeff(int k, int n);
void co_eval_beam(const ca_beam* beam, const double* xyz, int64_t n, double* out);
void co_sampler_sigmas(const double trap[3], const double atom[2], double sig[6]);
int64_t co_sample_one(const double trap[3], const double atom[2], uint64_t seed, uint64_t index, uint32_t stream, double eps, double out[6]);
int64_t co_sample_atoms(const double* trap, const double* atom, int64_t n, int64_t offset, uint64_t seed, int32_t stream, double eps, double* out);
double co_sample_atoms_metropolis(const double* trap, const double* atom, int64_t N, int64_t freq, int64_t skip, int harmonic, double eps, uint64_t seed, int32_t stream, double* out);
double co_sample_atoms_metropolis_chains(const double* trap, const double* atom, int64_t N, int64_t freq, int64_t skip, int harmonic, double eps, int64_t n_chains, uint64_t seed, int32_t stream, double* out);
void co_trap_frequencies(const double atom[2], const double trap[3], double* wr, double* wz);
void co_phi_amplitudes(const double* f, int nf, double h0, const double* hg, const double* sg, const double* fg, int ng, double* out);
void co_phase_draw(uint64_t seed, uint64_t index, uint32_t stream, int nf, double* phases);
void co_phi(const double* tn, int n_nodes, const double* f, const double* amp, const double* phases, int nf, double* out);
int co_rydberg1_run(const ca_model* model, const double* tspan, int32_t nt, const double* psi0, int64_t n_traj, int64_t traj_offset, int64_t n_total, uint64_t seed, const double* samples, const double* phases_red, const double* phases_blue, const double* f, const double* amp_red, const double* amp_blue, int32_t nf, const ca_ode_opts* ode, int32_t out_flags, double* rho, double* rho2, ca_stats* stats, int32_t nthreads);
int co_rydberg2_run(const ca_model* model, const double* tspan, int32_t nt, const double* psi0, int64_t n_traj, int64_t traj_offset, int64_t n_total, uint64_t seed, const double* samples1, const double* samples2, const double* const* phases4, const double* f, const double* amp_red, const double* amp_blue, int32_t nf, const ca_ode_opts* ode, int32_t out_flags, double* rho, double* rho2, ca_stats* stats, int32_t nthreads);
int co_rydberg2_run_ex(const ca_model* model, const double* tspan, int32_t nt, int32_t nt_seg1, const double* psi0, const double* rho0, int64_t n_traj, int64_t traj_offset, int64_t n_total, uint64_t seed, const double* samples1, const double* samples2, const double* const* phases4, const double* f, const double* amp_red, const double* amp_blue, int32_t nf, const ca_ode_opts* ode, int32_t out_flags, double* rho, double* rho2, ca_stats* stats, int32_t nthreads);
int co_release_recapture(const double* trap, const double* atom, const double* tspan, int32_t nt, int64_t n, int64_t traj_offset, int64_t n_total, double eps, uint64_t seed, const double* samples, int32_t out_flags, double* probs, uint8_t* indicators);
int co_num_threads(void);
void co_set_dense_emulation(int on); /* time the reference's dense mul! arithmetic (32 / 66 dense products per RHS) */
#ifdef __cplusplus
}
#endif
#endif
