/*
 * coldatoms_b200.h -- C ABI of the B200-native Monte-Carlo master-equation hot path.
 *
 * Drop-in boundary for the ensemble loops of mgoloshchapov/ColdAtoms.jl (module NeutralAtoms).
 * The reference has no FFI; its boundary is the Julia function API.  Each entry point below
 * replaces the body of one reference ensemble driver (file:line relative to the reference):
 *
 *   ca_rydberg1_run      <- simulation                src/rydberg_model.jl:194-265  (lines 206-264)
 *                           simulation_blue_intens    src/rydberg_model_arb_bms.jl:15-86
 *   ca_rydberg2_run      <- simulation_czlp           src/cz_model.jl:107-171       (lines 111-170)
 *   ca_rydberg2_run_ex   <- two_atom_simulation       src/two_atom_model.jl:89-190  (legacy parameterisation, model->two_atom_flags)
 *                           direct_CZ_simulation      src/two_atom_model.jl:192-319 (two segments, rho0 input, direct coupling)
 *   ca_release_recapture <- release_recapture         src/basic_experiments.jl:72-81
 *   ca_sample_atoms      <- samples_generate (harmonic=true branch) src/atom_sampler.jl:84-96
 *   ca_sample_atoms_metropolis <- samples_generate (harmonic=false branch) src/atom_sampler.jl:97-120
 *   ca_phase_noise       <- ϕ                         src/lasernoise_sampler.jl:31-37
 *   ca_rydberg1_sweep    <- the (config, ψ0) loops of src/fidelity.jl:78-88 and BASELINE config 5
 *   ca_rydberg2_sweep    <- the seven simulation_czlp ensembles of get_cz_infidelity src/fidelity.jl:193-240
 *   ca_thermal_average   <- the sample loops of calibrate_two_photon src/rydberg_model.jl:285-306 and
 *                           get_blockade_stark_shift_factor src/cz_model.jl:77-104
 *
 * Conventions
 *   - plain C types only; the caller owns every host buffer, the library owns device memory
 *     inside ca_ctx; calls on one context are serialised by the caller.
 *   - return 0 on success, a negative CA_ERR_* otherwise; ca_last_error() gives the text.
 *   - units: length um, time us, angular frequency rad/us, energy/temperature uK, mass a.m.u.
 *   - all floating point is IEEE fp64 / complex128 (complex stored as (re, im) pairs).
 *   - density matrices are returned per time point as d x d complex128, column-major
 *     (= Julia Matrix{ComplexF64}); level order 0,1,r,p,l (src/rydberg_model.jl:2-7); two-atom index
 *     = a1 + 5*a2 (QuantumOptics tensor convention, first factor fastest).
 *   - there is NO CPU fallback in this library: without a CUDA device ca_init fails.
 */
#ifndef COLDATOMS_B200_H
#define COLDATOMS_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define CA_VERSION 200 /* 0.2.0 */

/* error codes */
#define CA_OK 0
#define CA_ERR_ARG (-1)       /* bad argument */
#define CA_ERR_CUDA (-2)      /* CUDA runtime error */
#define CA_ERR_NODEVICE (-3)  /* no CUDA device / wrong architecture */
#define CA_ERR_MAXITERS (-4)  /* some trajectory hit ode.maxiters */
#define CA_ERR_NONFINITE (-5) /* a trajectory produced a non-finite state or dt underflow */
#define CA_ERR_UNSUPPORTED (-6)
#define CA_ERR_NOMEM (-7)
#define CA_ERR_NCCL (-8)      /* NCCL missing (libnccl.so.2 could not be loaded) or a NCCL call failed */

/* beam kinds: Ω dispatch of src/rydberg_model.jl:51-74 and Ω_red/Ω_blue of src/rydberg_model_arb_bms.jl:1-13 */
#define CA_BEAM_GAUSS 0      /* 5-param TEM00 in the θ-rotated frame: Ω0*A*A_phase (utilities.jl:27-48) */
#define CA_BEAM_FLATTOP_HG 1 /* simple_flattopHG_field (arbitrary_beams.jl:50-67) */
#define CA_BEAM_FLATTOP_LG 2 /* simple_flattopLG_field (arbitrary_beams.jl:69-81) */
#define CA_BEAM_UNIFORM 3    /* Ω_red: spatially uniform Ω0 (rydberg_model_arb_bms.jl:1-4) */

/* Doppler forms */
#define CA_DOPPLER_XZ 0 /* Δ(vx,vz,red), δ(vx,vz,red,blue) with θ (rydberg_model.jl:77-98) */
#define CA_DOPPLER_Z 1  /* legacy arb_bms: Δ=k_r vz, δ=(k_r ± k_b) vz   */

/* sign convention of the diagonal */
#define CA_SIGN_RYDBERG1 0 /* H_pp = -Δ_D-Δ0, H_rr = -δ_D-δ0  (rydberg_model.jl:180-181) */
#define CA_SIGN_CZ 1       /* H_pp = +Δ_D-Δ0, H_rr = +δ_D-δ0  (cz_model.jl:55-56)        */

/* compat bits: reference quirks that change numbers; default = all on */
#define CA_COMPAT_VZ_FROM_VY 1 /* B1: Vz(t) closure is built from vy (rydberg_model.jl:41) */
#define CA_COMPAT_DEFAULT (CA_COMPAT_VZ_FROM_VY)

/* second-moment output */
#define CA_RHO2_MATSQ 0 /* mean of ρ·ρ (what `ρt .^ 2` does on a Vector{Operator}) */
#define CA_RHO2_ABS2 1  /* mean of |ρ_ij|^2 element-wise */

/* output flags */
#define CA_OUT_MEAN 0  /* outputs divided by n_total (reference behaviour) */
#define CA_OUT_SUMS 1  /* un-normalised sums over the trajectories of THIS call (multi-GPU shards) */
#define CA_OUT_DEVICE 2 /* rho/rho2 are device pointers on the context's device; call returns without
                           synchronising (work is ordered on the context stream) */

typedef struct ca_ctx ca_ctx;

typedef struct ca_beam {
    double Omega0; /* rad/us */
    double w0;     /* waist, um */
    double z0;     /* Rayleigh length, um */
    double theta;  /* rotation of the beam axis (GAUSS only) */
    double n;      /* 5th laser param; ignored by Ω like in the reference (rydberg_model.jl:53-54) */
    double sqz;    /* HG: y-waist squeeze factor (arbitrary_beams.jl:51,62) */
    int32_t kind;  /* CA_BEAM_* */
    int32_t hg_n, hg_m; /* flat-top HG orders (even) */
    int32_t lg_l;  /* flat-top LG order: sum over p = 0..lg_l/2 */
    int32_t _pad;
} ca_beam;

/* One-atom model parameters = RydbergConfig (src/utilities.jl:177-200) minus arrays. */
typedef struct ca_model {
    double atom_m, atom_T;              /* atom_params */
    double trap_U0, trap_w0, trap_z0;   /* trap_params */
    ca_beam red, blue;                  /* red/blue_laser_params */
    double Delta0, delta0;              /* detuning_params */
    double Gamma0, Gamma1, Gammal, Gammar; /* decay_params (gated by the two decay flags) */
    int32_t atom_motion, free_motion, laser_noise;
    int32_t decay_intermediate, decay_rydberg;
    int32_t doppler_kind; /* CA_DOPPLER_* */
    int32_t parallel;     /* CA_DOPPLER_Z only: δ=(k_r+k_b)vz if 1 else (k_r-k_b)vz */
    int32_t detuning_sign;/* CA_SIGN_* */
    int32_t compat;       /* CA_COMPAT_* bits */
    int32_t rho2_mode;    /* CA_RHO2_* */
    int32_t n_noise_nodes;/* 0 -> 1001 = reference tspan_noise (rydberg_model.jl:216) */
    int32_t _pad;
    /* two-atom only (CZLPConfig, src/utilities.jl:203-233; read at cz_model.jl:111,137,161) */
    double center1[3], center2[3];
    double c6;
    double xi; /* red phase jump at t >= tspan[end]/2 (cz_model.jl:130,137) */
    /* legacy two-atom parameterisations (src/two_atom_model.jl:89-190 two_atom_simulation, :192-319 direct_CZ_simulation); all
     * zero = the cz_model.jl behaviour above */
    int32_t two_atom_flags; /* CA_2A_* bits */
    int32_t _pad2;
    ca_beam red2, blue2;    /* atom 2's beams for Ω (CA_2A_PER_ATOM_LASERS); the Doppler shifts of BOTH atoms use red/blue, as the
                               reference passes red/blue_laser_params_first to Δ and δ of atom 2 (two_atom_model.jl:158-159) */
    double Delta0_2, delta0_2; /* detuning_params_second (CA_2A_PER_ATOM_LASERS) */
    double V_rr;            /* constant blockade shift replacing c6/R^6 (CA_2A_CONST_VRR; two_atom_model.jl:165,283) */
    double seg2_phase;      /* direct CZ: the couplings of the SECOND time segment carry exp(i*seg2_phase) (two_atom_model.jl:287-299) */
} ca_model;

/* two_atom_flags */
#define CA_2A_PER_ATOM_LASERS 1   /* atom 2 uses red2/blue2 and Delta0_2/delta0_2 */
#define CA_2A_CONST_VRR 2         /* V = V_rr instead of c6/(1e-18 + R^6) */
#define CA_2A_SHIFT_AFTER_MOTION 4 /* the centres are beam-frame offsets added AFTER the motion R(t) (legacy X1(t)-X0, dist+X2(t)-X0,
                                      two_atom_model.jl:160-166), not shifts of the initial positions (cz_model.jl:124-125) */
#define CA_2A_DIRECT 8            /* direct one-photon coupling |1><r| with the BLUE beam's Ω(x,y,z)/2 and H_rr = +Delta0 (no Doppler shift,
                                      no intermediate level: direct_CZ_simulation, two_atom_model.jl:262-299); the red beam is unused.
                                      The reference's first-segment Hamiltonian couples σgr and σrg with the SAME coefficient (not
                                      conjugated, :270-274): non-Hermitian for a complex Ω; the Hermitian form is implemented. */

/* ode_kwargs subset (forwarded verbatim to OrdinaryDiffEq by the reference, rydberg_model.jl:255) */
typedef struct ca_ode_opts {
    double reltol;   /* <=0 -> 1e-6 (QuantumOptics default) */
    double abstol;   /* <=0 -> 1e-8 */
    double dt;       /* initial (adaptive) or fixed step; <=0 -> automatic initial step */
    double dtmax;    /* <=0 -> tspan[end]-tspan[0] */
    int64_t maxiters;/* <=0 -> 100000 (OrdinaryDiffEq default) */
    int32_t adaptive;/* 1 (default) embedded DP5(4) with PI control; 0 fixed step dt */
    int32_t dense;   /* 1 (default) outputs by the DP5 dense interpolant like the reference;
                        0: steps are clipped to hit every tspan point */
} ca_ode_opts;

typedef struct ca_stats {
    int64_t n_traj;
    int64_t n_rhs;      /* RHS evaluations actually executed, summed over trajectories */
    int64_t n_accept;
    int64_t n_reject;
    int64_t n_sample_attempts; /* sampler draws incl. rejected */
    double kernel_ms;   /* device time of the integration kernel(s) (CUDA events) */
    double total_ms;    /* device time of the whole call on the context stream */
    double noise_ms;    /* device time of the phase-noise table kernel */
    int32_t n_launches; /* kernels launched by this call */
    int32_t status;     /* 0 ok, or the CA_ERR_* observed on device */
    int64_t n_exact_steps; /* one-atom path: attempted steps whose Hamiltonian coefficients took the exact evaluation instead
                              of the anchored model (guard failed / generic variant); 0 on the other paths */
} ca_stats;

int ca_version(void);
const char* ca_last_error(void);

/* Create a context on one CUDA device (sm_100 required). One context per process and GPU. */
int ca_init(int device, ca_ctx** out);
int ca_finalize(ca_ctx* ctx);
/* Run on a caller-owned CUDA stream (e.g. torch's current stream; NULL is the legacy default stream, as in CUDA).
 * ca_reset_stream restores the context's own non-blocking stream. */
int ca_set_stream(ca_ctx* ctx, void* cuda_stream);
int ca_reset_stream(ca_ctx* ctx);
int ca_synchronize(ca_ctx* ctx);
/*
 * Deferred statistics of the most recent ensemble call on this context (ca_rydberg1_run / _sweep / ca_rydberg2_run).
 * A CA_OUT_DEVICE call returns without synchronising, so its device-side status (a trajectory that hit maxiters or went
 * non-finite is LEFT OUT of the sums) cannot be part of its return value:
 *   ca_collect_stats     synchronises the context stream, fills *stats and returns the status the synchronous call would
 *                        have returned (CA_OK / CA_ERR_MAXITERS / CA_ERR_NONFINITE);
 *   ca_stats_to_device   enqueues (no synchronisation) the conversion of the counters into 8 doubles at dev_out8 (device
 *                        pointer): n_rhs, n_accept, n_reject, n_sample_attempts, #(status == MAXITERS) in {0,1},
 *                        #(status == NONFINITE) in {0,1}, n_exact_steps, n_traj.  All eight ADD across shards, so a
 *                        multi-GPU caller appends them to the buffer it all-reduces and every rank sees the same status.
 */
int ca_collect_stats(ca_ctx* ctx, ca_stats* stats);
int ca_stats_to_device(ca_ctx* ctx, double* dev_out8);
/* Measured FP64 FMA throughput of the device in TFLOP/s (2 flop per DFMA); roofline denominator. */
int ca_measure_fp64_peak(ca_ctx* ctx, double* tflops, double* sm_mhz_est);

/*
 * One-atom ensemble: mean over trajectories of ρ(t) and ρ(t)^2 at every tspan point.
 *   psi0        5 x (re,im)
 *   n_traj      trajectories integrated by THIS call; global indices traj_offset .. traj_offset+n_traj-1
 *   n_total     divisor for CA_OUT_MEAN (0 -> n_traj)
 *   samples     NULL -> on-device Philox sampler keyed by (seed, global index); else n_traj x 6 row-major
 *               (x,y,z,vx,vy,vz) host array, used as is (parity protocol)
 *   phases_red/blue  NULL -> on-device Philox U[0,2π); else n_traj x nf host arrays
 *   f, amp_red, amp_blue   nf-vectors (cfg.f, cfg.*_laser_phase_amplitudes); may be NULL iff !laser_noise
 *   rho, rho2   nt*25 complex128 each (host, or device with CA_OUT_DEVICE)
 */
int ca_rydberg1_run(ca_ctx* ctx, const ca_model* model, const double* tspan, int32_t nt,
                    const double* psi0, int64_t n_traj, int64_t traj_offset, int64_t n_total,
                    uint64_t seed, const double* samples, const double* phases_red,
                    const double* phases_blue, const double* f, const double* amp_red,
                    const double* amp_blue, int32_t nf, const ca_ode_opts* ode, int32_t out_flags,
                    double* rho, double* rho2, ca_stats* stats);

/*
 * Two-atom CZ ensemble (simulation_czlp). psi0: 25 x (re,im), index a1+5*a2.
 * samples1/2: NULL or n_traj x 6 host arrays BEFORE the centre shift (cz_model.jl:124-125 adds it);
 * phases: NULL or 4 host arrays n_traj x nf in the order red1, blue1, red2, blue2.
 * rho, rho2: nt*625 complex128 each.
 */
int ca_rydberg2_run(ca_ctx* ctx, const ca_model* model, const double* tspan, int32_t nt,
                    const double* psi0, int64_t n_traj, int64_t traj_offset, int64_t n_total,
                    uint64_t seed, const double* samples1, const double* samples2,
                    const double* const* phases4, const double* f, const double* amp_red,
                    const double* amp_blue, int32_t nf, const ca_ode_opts* ode, int32_t out_flags,
                    double* rho, double* rho2, ca_stats* stats);

/*
 * General two-atom entry: everything ca_rydberg2_run does, plus
 *   rho0     NULL, or 625 complex128 (column-major 25x25) initial density matrix replacing |psi0><psi0| (direct_CZ_simulation starts
 *            from a density matrix, two_atom_model.jl:192,226); psi0 may then be NULL;
 *   nt_seg1  0, or the number of leading tspan points that form a FIRST time segment: tspan[0..nt_seg1) and tspan[nt_seg1..nt) are
 *            then two consecutive solves (each a fresh DP5 solve on its own time axis, the second starting from the per-trajectory
 *            final state of the first: two_atom_model.jl:304-305); outputs are concatenated (nt matrices).
 * Initial states with |l> components, rho0, two segments and any two_atom_flags run on the full 25x25 kernel (k_lindblad2_full);
 * the cz_model.jl case with psi0 inside the {0,1,r,p}^2 block runs on the blocked kernel (k_lindblad2).  ca_rydberg2_run is
 * ca_rydberg2_run_ex with rho0 = NULL, nt_seg1 = 0.
 */
int ca_rydberg2_run_ex(ca_ctx* ctx, const ca_model* model, const double* tspan, int32_t nt, int32_t nt_seg1,
                       const double* psi0, const double* rho0, int64_t n_traj, int64_t traj_offset, int64_t n_total,
                       uint64_t seed, const double* samples1, const double* samples2,
                       const double* const* phases4, const double* f, const double* amp_red,
                       const double* amp_blue, int32_t nf, const ca_ode_opts* ode, int32_t out_flags,
                       double* rho, double* rho2, ca_stats* stats);

/*
 * Batched one-atom sweep: n_points parameter points in one launch (fidelity.jl:78-88 loops; BASELINE
 * config 5).  Each point p has its own model, ψ0 (5 complex), end time t_end[p] (outputs only ρ(t_end),
 * ρ²(t_end): the `[end]` the fidelity callers read) and n_traj_per_point trajectories.
 * rho/rho2: n_points*25 complex128.
 */
int ca_rydberg1_sweep(ca_ctx* ctx, const ca_model* models, const double* psi0s, const double* t_end,
                      int32_t n_points, int64_t n_traj_per_point, int64_t traj_offset, int64_t n_total,
                      uint64_t seed, const double* f, const double* amp_red, const double* amp_blue,
                      int32_t nf, const ca_ode_opts* ode, int32_t out_flags, double* rho, double* rho2,
                      ca_stats* stats);

/*
 * Batched two-atom sweep: n_points CZ configurations in one launch (the seven ensembles of get_cz_infidelity, src/fidelity.jl:193-240;
 * SURVEY 8f-1).  Point p has its own model, ψ0 (25 complex, inside the {0,1,r,p}^2 block), end time t_end[p] (tspan = [0, t_end[p]];
 * only ρ(t_end), ρ²(t_end) are returned) and trajectory count n_traj_per_point[p]; its trajectories carry the global indices
 * traj_offset + (Σ_{q<p} n_traj_per_point[q]) + i, and all trajectories are handed out longest first across the points.
 * rho/rho2: n_points*625 complex128, each point divided by its own count (CA_OUT_MEAN) or left as sums (CA_OUT_SUMS).
 */
int ca_rydberg2_sweep(ca_ctx* ctx, const ca_model* models, const double* psi0s, const double* t_end, int32_t n_points,
                      const int64_t* n_traj_per_point, int64_t traj_offset, uint64_t seed, const double* f,
                      const double* amp_red, const double* amp_blue, int32_t nf, const ca_ode_opts* ode,
                      int32_t out_flags, double* rho, double* rho2, ca_stats* stats);

/*
 * Release and recapture (basic_experiments.jl:25-81).  trap = (U0,w0,z0), atom = (m,T).
 * samples NULL -> Philox harmonic sampler (sampler eps stays 1e-2: basic_experiments.jl:73).
 * probs[nt]: mean indicator (CA_OUT_MEAN) or counts as doubles (CA_OUT_SUMS); with CA_OUT_DEVICE a device pointer, written on
 * the context stream without synchronising (indicators must be NULL then; stats carry no sampler count).
 * indicators: NULL or n x nt bytes (host) for bit-exact parity.
 */
int ca_release_recapture(ca_ctx* ctx, const double* trap, const double* atom, const double* tspan,
                         int32_t nt, int64_t n, int64_t traj_offset, int64_t n_total, double eps,
                         uint64_t seed, const double* samples, int32_t out_flags, double* probs,
                         uint8_t* indicators, ca_stats* stats);

/* Philox samplers exposed for statistical tests / exported helpers (samples_generate, ϕ). */
int ca_sample_atoms(ca_ctx* ctx, const double* trap, const double* atom, int64_t n, int64_t traj_offset,
                    uint64_t seed, int32_t stream_id, double eps, double* samples_out /* n x 6 host */,
                    int64_t* attempts_out);

/*
 * Metropolis branch of samples_generate (src/atom_sampler.jl:97-120) run as n_chains independent chains (one per thread):
 * every chain starts at the reference's initial state, discards `skip` states and emits every freq-th one; chain c's
 * k-th sample is samples_out[c + k*n_chains].  n_chains = 1 is the reference's single chain; n_chains <= 0 picks
 * min(n, 256 per SM).  harmonic != 0 uses Π_Harmonic in the Boltzmann weight (the `harmonic` keyword forwarded by
 * prob_boltzmann / H).  acc_rate_out: accepted transitions / states visited (= acc_rate/(N*freq+skip) for one chain).
 */
int ca_sample_atoms_metropolis(ca_ctx* ctx, const double* trap, const double* atom, int64_t n, int64_t freq,
                               int64_t skip, int32_t harmonic, double eps, int64_t n_chains, uint64_t seed,
                               int32_t stream_id, double* samples_out, double* acc_rate_out);
int ca_phase_noise(ca_ctx* ctx, const double* tnodes, int32_t n_nodes, const double* f,
                   const double* amp, int32_t nf, int64_t n, int64_t traj_offset, uint64_t seed,
                   int32_t stream_id, const double* phases /* NULL or n x nf */,
                   double* phi_out /* n x n_nodes host */);

/* Flat-top / Gaussian field evaluation on device (golden-vector check of the inlined beam code):
 * out[i] = Ω(x[i],y[i],z[i]) as (re,im). */
int ca_eval_beam(ca_ctx* ctx, const ca_beam* beam, const double* xyz /* n x 3 host */, int64_t n,
                 double* out /* n x 2 host */);

/*
 * Thermal averages of scalar functionals of the sampled atoms, reduced on the device (un-normalised SUMS over
 * [traj_offset, traj_offset+n), so shards add and the caller divides by the ensemble size):
 *   CA_AVG_TWOPHOTON  the two sample loops of calibrate_two_photon (src/rydberg_model.jl:293 and :300) for model->red/blue:
 *       sums[0] = Σ Ω_twophoton(Ωr(x), Ωb(x), Δ(vx,vz,red)), sums[1] = Σ δ_twophoton(Ωr(x), Ωb(x), Δ(vx,vz,red))
 *       (Δ is the red Doppler shift alone, as the reference passes it);
 *   CA_AVG_RM6        the map/mean of get_blockade_stark_shift_factor (src/cz_model.jl:101) for model->center1/center2:
 *       sums[0] = Σ 1/(1e-18 + |r1+c1-r2-c2|^6), sums[1] = 0.
 * samples1/samples2: NULL -> Philox harmonic sampler (streams atom1/atom2, the ensembles' own), or host n x 6 arrays
 * (samples2 only for CA_AVG_RM6).  eps <= 0 -> the sampler's default 1e-2 (atom_sampler.jl:80).
 */
#define CA_AVG_TWOPHOTON 0
#define CA_AVG_RM6 1
int ca_thermal_average(ca_ctx* ctx, int32_t kind, const ca_model* model, int64_t n, int64_t traj_offset,
                       uint64_t seed, double eps, const double* samples1, const double* samples2,
                       double* sums_out /* 2 host */, int64_t* attempts_out);

/*
 * ---- several GPUs in ONE process (the dead simulation_mpi of src/rydberg_model.jl:309-383 meant to MPI.Reduce the
 * same payload) -----------------------------------------------------------------------------------------------------------
 * ca_init_multi creates one context per listed device (devices == NULL: 0..ndev-1; ndev <= 0: every visible device) and, for
 * ndev > 1, one NCCL clique (ncclCommInitAll; libnccl.so.2 is loaded with dlopen on first use, so single-GPU users never need it).
 * The *_multi ensemble calls take exactly the arguments of their single-GPU counterparts (host output only: CA_OUT_MEAN or
 * CA_OUT_SUMS).  Device g integrates the contiguous slice [g*n/G, (g+1)*n/G) of the trajectory index space -- the Philox streams
 * are keyed by the GLOBAL index, so the ensemble is the one a single GPU would draw --, every device leaves its un-normalised
 * sums in its own HBM, and ONE ncclAllReduce(sum, fp64) per call combines [sum rho | sum rho*rho | 8 counters] over NVLink.
 * The counters carry the device-side status: a trajectory that hit maxiters / went non-finite on ANY device fails the call.
 * ca_rydberg1_sweep_multi shards by grid point instead and needs no collective (results are gathered on the host).
 */
typedef struct ca_multi ca_multi;
int ca_init_multi(const int* devices, int32_t ndev, ca_multi** out);
int ca_finalize_multi(ca_multi* m);
int ca_multi_size(const ca_multi* m);            /* number of devices */
ca_ctx* ca_multi_context(ca_multi* m, int32_t i); /* the i-th single-GPU context (owned by m) */
int ca_rydberg1_run_multi(ca_multi* m, const ca_model* model, const double* tspan, int32_t nt,
                          const double* psi0, int64_t n_traj, int64_t traj_offset, int64_t n_total,
                          uint64_t seed, const double* samples, const double* phases_red,
                          const double* phases_blue, const double* f, const double* amp_red,
                          const double* amp_blue, int32_t nf, const ca_ode_opts* ode, int32_t out_flags,
                          double* rho, double* rho2, ca_stats* stats);
int ca_rydberg2_run_multi(ca_multi* m, const ca_model* model, const double* tspan, int32_t nt,
                          const double* psi0, int64_t n_traj, int64_t traj_offset, int64_t n_total,
                          uint64_t seed, const double* samples1, const double* samples2,
                          const double* const* phases4, const double* f, const double* amp_red,
                          const double* amp_blue, int32_t nf, const ca_ode_opts* ode, int32_t out_flags,
                          double* rho, double* rho2, ca_stats* stats);
int ca_rydberg1_sweep_multi(ca_multi* m, const ca_model* models, const double* psi0s, const double* t_end,
                            int32_t n_points, int64_t n_traj_per_point, int64_t traj_offset, int64_t n_total,
                            uint64_t seed, const double* f, const double* amp_red, const double* amp_blue,
                            int32_t nf, const ca_ode_opts* ode, int32_t out_flags, double* rho, double* rho2,
                            ca_stats* stats);
int ca_release_recapture_multi(ca_multi* m, const double* trap, const double* atom, const double* tspan,
                               int32_t nt, int64_t n, int64_t traj_offset, int64_t n_total, double eps,
                               uint64_t seed, const double* samples, int32_t out_flags, double* probs,
                               ca_stats* stats);

#ifdef __cplusplus
}
#endif
#endif /* COLDATOMS_B200_H */
