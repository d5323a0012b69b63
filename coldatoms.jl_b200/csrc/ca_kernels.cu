// ca_kernels.cu -- sm_100a kernels and the C ABI (include/coldatoms_b200.h) of the one-atom path,
// the samplers and release-recapture.  The two-atom kernel lives in ca_lindblad2.cu.
//
// Kernels (SURVEY §8a "Kernel <-> reference mapping"):
//   k_lindblad1<NS,SWEEP,BEAMS>  K1+K2+K3 fused (ca_lindblad1.cuh; one translation unit per variant: ca_l1_inst.cu)
//   k_finalize             K6: packed sums -> full column-major complex matrices, scaled by 1/N.
//   k_release_recapture    K5.
//   k_sample_atoms, k_phase_noise, k_eval_beam   exposed samplers / beam code for tests and helpers.
//   k_dfma_peak            FP64 FMA-chain microbenchmark (roofline denominator).
#include <cuda_runtime.h>

#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <string>
#include <vector>

#include "ca_host.h"
#include "ca_lindblad1.cuh"

namespace ca {

// K6: packed Hermitian sums -> column-major complex d x d matrices scaled by `scale`.  Output matrix i is accumulator matrix
// i*stride + first (sweeps keep only the last time point of every grid point: stride = nt, first = nt - 1).
__global__ void k_finalize5(const double* __restrict__ accum, long long n_mats, long long stride, long long first, double scale,
                            double* __restrict__ rho, double* __restrict__ rho2) {
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_mats * 2) return;
    const long long mat = i >> 1;
    const int which = (int)(i & 1);
    const double* p = accum + (mat * stride + first) * kNV + which * 25;
    double* o = (which ? rho2 : rho) + mat * 50;
    const int off[5][5] = {{-1, 0, 1, 2, 3}, {-1, -1, 4, 5, 6}, {-1, -1, -1, 7, 8}, {-1, -1, -1, -1, 9}, {-1, -1, -1, -1, -1}};
    for (int b = 0; b < 5; ++b)
        for (int a = 0; a < 5; ++a) {
            double re, im;
            if (a == b) { re = p[a]; im = 0.0; }
            else if (a < b) { re = p[5 + 2 * off[a][b]]; im = p[6 + 2 * off[a][b]]; }
            else { re = p[5 + 2 * off[b][a]]; im = -p[6 + 2 * off[b][a]]; }
            o[2 * (a + 5 * b)] = re * scale; o[2 * (a + 5 * b) + 1] = im * scale;
        }
}

__global__ void k_counts_to_double(const unsigned long long* __restrict__ counts, int nt, double scale, double* __restrict__ out) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j < nt) out[j] = (double)counts[j] * scale;
}

// ----------------------------------------------------------------------------------------------------
// K5: release and recapture (basic_experiments.jl:25-47, :72-81)
// ----------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_release_recapture(SamplerConsts sc, const double* __restrict__ tspan, int nt, long long n,
                                                           long long traj_offset, unsigned long long seed, double thresh,
                                                           const double* __restrict__ samples, unsigned long long* __restrict__ counts,
                                                           unsigned char* __restrict__ indicators, unsigned long long* __restrict__ attempts) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const int lane = threadIdx.x & 31;
    const bool valid = i < n;
    double c[6] = {0, 0, 0, 0, 0, 0};
    int att = 0;
    if (valid) {
        if (samples) {
#pragma unroll
            for (int q = 0; q < 6; ++q) c[q] = samples[i * 6 + q];
        } else att = sample_atom_literal(sc, seed, (uint64_t)(traj_offset + i), 0u, c);
    }
    const double kin = kinetic_literal(sc.mfac, c[3], c[4], c[5]);
    bool alive = valid;
    for (int j = 0; j < nt; ++j) {
        if (alive) {
            const double t = tspan[j];
            const double x = __dadd_rn(c[0], __dmul_rn(c[3], t)), y = __dadd_rn(c[1], __dmul_rn(c[4], t)), z = __dadd_rn(c[2], __dmul_rn(c[5], t));
            const double A = trap_A_literal(x, y, z, sc.w0, sc.z0);
            const double pot = __dmul_rn(sc.U0, __dadd_rn(1.0, -__dmul_rn(A, A)));
            alive = __dadd_rn(kin, pot) < thresh;
        }
        const unsigned mask = __ballot_sync(0xffffffffu, alive);
        if (indicators && valid) indicators[i * nt + j] = alive ? 1 : 0;
        if (lane == 0 && mask) atomicAdd(counts + j, (unsigned long long)__popc(mask));
        if (!mask && !indicators) break;
    }
    if (attempts) {
        double a = warp_sum((double)att);
        if (lane == 0 && a > 0) atomicAdd(attempts, (unsigned long long)a);
    }
}

__global__ void k_sample_atoms(SamplerConsts sc, long long n, long long traj_offset, unsigned long long seed, unsigned stream,
                               double* __restrict__ out, unsigned long long* __restrict__ attempts) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    double c[6];
    int att = sample_atom_literal(sc, seed, (uint64_t)(traj_offset + i), stream, c);
#pragma unroll
    for (int q = 0; q < 6; ++q) out[i * 6 + q] = c[q];
    if (att > 0) atomicAdd(attempts, (unsigned long long)att);
}

// Metropolis sampler: one chain per thread (atom_sampler.jl:97-120 run as n_chains independent chains)
__global__ void __launch_bounds__(128) k_metropolis(MetroConsts mc, unsigned long long seed, unsigned stream, long long n_chains, long long N,
                                                    long long freq, long long skip, double* __restrict__ out, unsigned long long* __restrict__ counters) {
    const long long c = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= n_chains) return;
    long long steps = 0;
    const long long acc = metropolis_chain(mc, seed, stream, c, n_chains, N, freq, skip, out, &steps);
    atomicAdd(counters, (unsigned long long)acc);
    atomicAdd(counters + 1, (unsigned long long)steps);
}

// ϕ tables: one thread per trajectory, output phi[i][node] (lasernoise_sampler.jl:31-37 on a uniform grid)
__global__ void __launch_bounds__(128) k_phase_noise(const NoiseK* __restrict__ nk, int nf, int n_nodes, double dtn, int coarse,
                                                     const double* __restrict__ phases, long long n, long long traj_offset,
                                                     unsigned long long seed, unsigned stream, double* __restrict__ out, double* __restrict__ tmp) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    gen_noise_trace_c<kNoiseChunk>(nk, nf, n_nodes, dtn, coarse, phases ? phases + i * nf : nullptr, seed, (uint64_t)(traj_offset + i), stream,
                                   out + (size_t)i * n_nodes, 1, tmp + (size_t)i * noise_coarse_count(n_nodes, coarse), 1);
}

__global__ void k_eval_beam(DevBeam b, const double* __restrict__ xyz, long long n, double* __restrict__ out) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    double re, im;
    beam_half_omega(b, xyz[3 * i], xyz[3 * i + 1], xyz[3 * i + 2], 0.0, &re, &im);
    out[2 * i] = 2.0 * re; out[2 * i + 1] = 2.0 * im;
}

// Thermal averages (calibrate_two_photon rydberg_model.jl:285-306, get_blockade_stark_shift_factor cz_model.jl:77-104): grid-stride
// over samples (device-drawn with the same Philox streams as the ensembles, or host-supplied), per-thread partial sums in index
// order, fixed-order block reduction -> partials[block][2]; k_sum_partials adds the blocks in order, so a given (n, grid) is
// bit-reproducible and shards of one ensemble add up.
__global__ void __launch_bounds__(128) k_thermal_avg(AvgConsts ac, SamplerConsts sc, long long n, long long traj_offset, unsigned long long seed,
                                                     const double* __restrict__ s1_in, const double* __restrict__ s2_in,
                                                     double* __restrict__ partials, unsigned long long* __restrict__ attempts) {
    __shared__ double red[4][2];
    double a0 = 0.0, a1 = 0.0;
    long long att = 0;
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        double s1[6], s2[6] = {0, 0, 0, 0, 0, 0}, o[2];
        if (s1_in) {
#pragma unroll
            for (int q = 0; q < 6; ++q) s1[q] = s1_in[i * 6 + q];
        } else att += sample_atom_literal(sc, seed, (uint64_t)(traj_offset + i), 0u, s1);
        if (ac.kind == 1) {
            if (s2_in) {
#pragma unroll
                for (int q = 0; q < 6; ++q) s2[q] = s2_in[i * 6 + q];
            } else att += sample_atom_literal(sc, seed, (uint64_t)(traj_offset + i), 1u, s2);
        }
        thermal_terms(ac, s1, s2, o);
        a0 += o[0]; a1 += o[1];
    }
    a0 = warp_sum(a0); a1 = warp_sum(a1);
    const double fa = warp_sum((double)att);
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    if (lane == 0) { red[w][0] = a0; red[w][1] = a1; if (fa > 0) atomicAdd(attempts, (unsigned long long)fa); }
    __syncthreads();
    if (threadIdx.x == 0) {
        partials[2 * blockIdx.x] = (red[0][0] + red[1][0]) + (red[2][0] + red[3][0]);
        partials[2 * blockIdx.x + 1] = (red[0][1] + red[1][1]) + (red[2][1] + red[3][1]);
    }
}
__global__ void k_sum_partials(const double* __restrict__ partials, int nb, double* __restrict__ out) {
    if (threadIdx.x < 2) {
        double s = 0.0;
        for (int b = 0; b < nb; ++b) s += partials[2 * b + threadIdx.x];
        out[threadIdx.x] = s;
    }
}

// FP64 FMA-chain microbenchmark (roofline denominator): 8 independent chains per thread, 8 warps per SM (two per scheduler keep the
// half-rate FP64 pipe saturated: DFMA latency 8.4 cycles, issue interval 2 -- profiles/micro/lat.cu; more warps only add
// scheduling noise), long unrolled bodies so that the loop overhead is < 1 %.
__global__ void __launch_bounds__(256) k_dfma_peak(double* out, int iters, double a, double b) {
    double x0 = threadIdx.x, x1 = x0 + 1, x2 = x0 + 2, x3 = x0 + 3, x4 = x0 + 4, x5 = x0 + 5, x6 = x0 + 6, x7 = x0 + 7;
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int r = 0; r < 32; ++r) {
            x0 = fma(x0, a, b); x1 = fma(x1, a, b); x2 = fma(x2, a, b); x3 = fma(x3, a, b);
            x4 = fma(x4, a, b); x5 = fma(x5, a, b); x6 = fma(x6, a, b); x7 = fma(x7, a, b);
        }
    }
    double s = x0 + x1 + x2 + x3 + x4 + x5 + x6 + x7;
    if (s == 12345.678) out[0] = s;
}

// ----------------------------------------------------------------------------------------------------
// context
// ----------------------------------------------------------------------------------------------------
#define CA_CUDA(expr)                                                                                       \
    do {                                                                                                    \
        cudaError_t _e = (expr);                                                                            \
        if (_e != cudaSuccess) {                                                                            \
            set_error(std::string(#expr) + ": " + cudaGetErrorString(_e));                                  \
            return CA_ERR_CUDA;                                                                             \
        }                                                                                                   \
    } while (0)

struct DevBuf {
    void* p = nullptr;
    size_t cap = 0;
    int ensure(size_t bytes) {
        if (bytes <= cap) return CA_OK;
        if (p) cudaFree(p);
        p = nullptr; cap = 0;
        size_t want = bytes + bytes / 4;
        cudaError_t e = cudaMalloc(&p, want);
        if (e != cudaSuccess) { set_error(std::string("cudaMalloc: ") + cudaGetErrorString(e)); return CA_ERR_NOMEM; }
        cap = want;
        return CA_OK;
    }
    void release() { if (p) cudaFree(p); p = nullptr; cap = 0; }
};

}  // namespace ca

struct ca_ctx {
    int device = 0;
    int sm_count = 0;
    cudaStream_t own_stream = nullptr, stream = nullptr;
    cudaEvent_t ev[4] = {nullptr, nullptr, nullptr, nullptr};
    ca::DevBuf scratch, accum, out, counters, in_samples, in_ph[4], in_f, in_models, in_tspan, misc;
    std::map<int, ca::DevBuf> extra;  // buffers of other translation units (ca_lindblad2.cu), by id
    void* pinned = nullptr;
    size_t pinned_cap = 0;
    // most recent ensemble call (ca_collect_stats / ca_stats_to_device): its 8 device counters, size and launch count
    unsigned long long* last_counters = nullptr;
    long long last_n = 0;
    int last_launches = 0;
    int last_kind = 0;   // 1: one-atom (counters[5..7] = noise cycles, total cycles, exact-coefficient steps), 2: two-atom
};

namespace ca {

static int ensure_pinned(ca_ctx* c, size_t bytes) {
    if (bytes <= c->pinned_cap) return CA_OK;
    if (c->pinned) cudaFreeHost(c->pinned);
    c->pinned = nullptr; c->pinned_cap = 0;
    CA_CUDA(cudaMallocHost(&c->pinned, bytes + bytes / 4));
    c->pinned_cap = bytes + bytes / 4;
    return CA_OK;
}

static int upload(ca_ctx* c, DevBuf& b, const void* host, size_t bytes) {
    int rc = b.ensure(bytes);
    if (rc) return rc;
    CA_CUDA(cudaMemcpyAsync(b.p, host, bytes, cudaMemcpyHostToDevice, c->stream));
    return CA_OK;
}

cudaStream_t ctx_stream(ca_ctx* c) { return c->stream; }
int ctx_sm_count(ca_ctx* c) { return c->sm_count; }
int ctx_device(ca_ctx* c) { return c->device; }
cudaEvent_t ctx_event(ca_ctx* c, int i) { return c->ev[i]; }
void* ctx_buf(ca_ctx* c, int which, size_t bytes) {
    DevBuf& b = c->extra[which];
    if (b.ensure(bytes ? bytes : 8)) return nullptr;
    return b.p;
}

void ctx_set_last(ca_ctx* c, unsigned long long* counters, long long n, int launches, int kind) {
    c->last_counters = counters; c->last_n = n; c->last_launches = launches; c->last_kind = kind;
}

// counters -> 8 doubles that add across shards (ca_stats_to_device)
__global__ void k_export_counters(const unsigned long long* __restrict__ cnt, long long n, int kind, double* __restrict__ out) {
    const int i = threadIdx.x;
    if (i >= 8) return;
    double v;
    if (i < 4) v = (double)cnt[i];
    else if (i == 4) v = cnt[4] == (unsigned long long)(-CA_ERR_MAXITERS) ? 1.0 : 0.0;
    else if (i == 5) v = cnt[4] == (unsigned long long)(-CA_ERR_NONFINITE) ? 1.0 : 0.0;
    else if (i == 6) v = (double)cnt[kind == 1 ? 7 : 5];
    else v = (double)n;
    out[i] = v;
}

// Synchronise, read the counters of the most recent ensemble call and report them (the tail of every synchronous call; the whole of
// ca_collect_stats after a CA_OUT_DEVICE call).
int collect_last(ca_ctx* c, ca_stats* stats) {
    if (stats) std::memset(stats, 0, sizeof *stats);
    if (!c->last_counters) { set_error("no ensemble call has run on this context"); return CA_ERR_ARG; }
    unsigned long long hc[8];
    CA_CUDA(cudaMemcpyAsync(hc, c->last_counters, sizeof hc, cudaMemcpyDeviceToHost, c->stream));
    CA_CUDA(cudaStreamSynchronize(c->stream));
    const int status = hc[4] ? -(int)hc[4] : 0;
    if (stats) {
        stats->n_traj = c->last_n; stats->n_rhs = (int64_t)hc[0]; stats->n_accept = (int64_t)hc[1]; stats->n_reject = (int64_t)hc[2];
        stats->n_sample_attempts = (int64_t)hc[3]; stats->n_launches = c->last_launches; stats->status = status;
        float ms = 0;
        cudaEventElapsedTime(&ms, c->ev[1], c->ev[2]); stats->kernel_ms = ms;
        if (c->last_kind == 1) {
            stats->noise_ms = hc[6] ? ms * (double)hc[5] / (double)hc[6] : 0.0;  // share of warp-cycles spent generating noise tables
            stats->n_exact_steps = (int64_t)hc[7];
        } else stats->n_exact_steps = (int64_t)hc[5];
        cudaEventElapsedTime(&ms, c->ev[0], c->ev[3]); stats->total_ms = ms;
    }
    if (status == CA_ERR_MAXITERS) { set_error("a trajectory hit maxiters (it is left out of the sums)"); return status; }
    if (status) { set_error("a trajectory produced a non-finite state or the step size underflowed (it is left out of the sums)"); return status; }
    return CA_OK;
}

// Shared driver of ca_rydberg1_run (n_points = 1) and ca_rydberg1_sweep.
static int run_r1(ca_ctx* c, const ca_model* models, const double* psi0s, const double* tspans, int nt, int n_points,
                  long long n_per_point, long long traj_offset, long long n_total, uint64_t seed, const double* samples,
                  const double* phases_red, const double* phases_blue, const double* f, const double* amp_red,
                  const double* amp_blue, int nf, const ca_ode_opts* ode, int out_flags, bool last_only, double* rho, double* rho2,
                  ca_stats* stats) {
    if (!c) { set_error("null context"); return CA_ERR_ARG; }
    if (!models || !psi0s || !tspans || nt < 1 || n_points < 1 || n_per_point < 0 || !rho || !rho2) { set_error("bad argument"); return CA_ERR_ARG; }
    CA_CUDA(cudaSetDevice(c->device));
    std::vector<DevModel> dm(n_points);
    int rows = 0;
    bool any_noise = false;
    int n_nodes_max = 2, n_coarse_max = 0;
    for (int p = 0; p < n_points; ++p) {
        const double* ts = tspans + (size_t)p * nt;
        for (int j = 1; j < nt; ++j)
            if (!(ts[j] > ts[j - 1])) { set_error("tspan must be strictly increasing"); return CA_ERR_ARG; }
        int rc = derive_model(models[p], psi0s + 10 * (size_t)p, 5, ts[0], ts[nt - 1], &dm[p]);
        if (rc) return rc;
        rows |= dm[p].rows;
        if (dm[p].laser_noise) {
            if (!f || !amp_red || !amp_blue || nf < 1) { set_error("laser_noise=true needs f and amplitudes"); return CA_ERR_ARG; }
            any_noise = true;
            if (ts[0] < 0) { set_error("laser noise is tabulated on [0, tspan[end]]: tspan[0] must be >= 0"); return CA_ERR_ARG; }
            dm[p].noise_coarse = choose_noise_coarse(f, amp_red, amp_blue, nf, dm[p].dtn, dm[p].n_nodes);
            n_coarse_max = std::max(n_coarse_max, noise_coarse_count(dm[p].n_nodes, dm[p].noise_coarse));
        }
        n_nodes_max = std::max(n_nodes_max, dm[p].n_nodes);
    }
    OdeOpts oo;
    fill_ode(ode, tspans[0], tspans[nt - 1], &oo);
    if (n_points > 1 && !(ode && ode->dtmax > 0)) oo.dtmax = 1e300;  // per-point horizons differ
    if (!oo.adaptive && !(oo.dt0 > 0)) { set_error("adaptive=false needs dt > 0"); return CA_ERR_ARG; }

    R1Params P;
    std::memset(&P, 0, sizeof P);
    P.model = dm[0];
    P.ode = oo;
    P.n_points = n_points; P.nt = nt; P.nf = any_noise ? nf : 0;
    P.n_per_point = n_per_point; P.traj_offset = traj_offset; P.seed = seed;
    P.groups_per_point = (n_per_point + 31) / 32;
    P.n_groups = P.groups_per_point * n_points;
    P.n_nodes_max = n_nodes_max; P.n_coarse_max = n_coarse_max;
    const bool sweep = n_points > 1;

    int rc;
    CA_CUDA(cudaEventRecord(c->ev[0], c->stream));
    if ((rc = upload(c, c->in_tspan, tspans, sizeof(double) * (size_t)nt * n_points))) return rc;
    P.tspans = (const double*)c->in_tspan.p;
    if (sweep) {
        if ((rc = upload(c, c->in_models, dm.data(), sizeof(DevModel) * (size_t)n_points))) return rc;
        P.models = (const DevModel*)c->in_models.p;
    }
    const long long n_all = n_per_point * n_points;
    if (samples) { if ((rc = upload(c, c->in_samples, samples, sizeof(double) * 6 * (size_t)n_all))) return rc; P.samples = (const double*)c->in_samples.p; }
    if (any_noise) {
        const size_t nkb = sizeof(NoiseK) * 2 * (size_t)nf * n_points;
        if ((rc = ensure_pinned(c, nkb))) return rc;
        NoiseK* hk = (NoiseK*)c->pinned;
        for (int p = 0; p < n_points; ++p) {
            fill_noise_k(f, amp_red, nf, noise_c(dm[p].noise_coarse) * dm[p].dtn, hk + (size_t)p * nf);
            fill_noise_k(f, amp_blue, nf, noise_c(dm[p].noise_coarse) * dm[p].dtn, hk + (size_t)(n_points + p) * nf);
        }
        if ((rc = c->in_f.ensure(nkb))) return rc;
        CA_CUDA(cudaMemcpyAsync(c->in_f.p, c->pinned, nkb, cudaMemcpyHostToDevice, c->stream));
        CA_CUDA(cudaStreamSynchronize(c->stream));  // pinned staging buffer is reused below
        P.nk_red = (const NoiseK*)c->in_f.p; P.nk_blue = P.nk_red + (size_t)n_points * nf;
        if (phases_red) { if ((rc = upload(c, c->in_ph[0], phases_red, sizeof(double) * (size_t)nf * n_all))) return rc; P.phases_red = (const double*)c->in_ph[0].p; }
        if (phases_blue) { if ((rc = upload(c, c->in_ph[1], phases_blue, sizeof(double) * (size_t)nf * n_all))) return rc; P.phases_blue = (const double*)c->in_ph[1].p; }
    }
    // variant: NS from the rows of ψ0; the anchored-coefficient code is compiled per beam class (all points free-flight Gaussian pairs = 1 /
    // free flight with other beams = 2 / Gaussian pairs oscillating in the trap = 3 / anything else or a mixed sweep = 0)
    const int ns_sel = rows == 0 ? 9 : (rows == 3 ? 21 : 15);
    int n_gp = 0, n_free = 0;
    for (int p = 0; p < n_points; ++p) {
        n_gp += dm[p].red.kind == CA_BEAM_GAUSS && dm[p].blue.kind == CA_BEAM_GAUSS;
        n_free += dm[p].free_motion != 0;
    }
    int beams_sel = 0;
    if (n_gp == n_points && n_free == n_points) beams_sel = 1;
    else if (ns_sel == 9 && n_gp == 0 && n_free == n_points) beams_sel = 2;
    else if (ns_sel == 9 && n_gp == n_points && n_free == 0) beams_sel = 3;
    if (ns_sel == 21) beams_sel = 0;
    // grid: persistent warps, one CTA per SM
    const int wpb = threads_for_rt(ns_sel, beams_sel) / 32;
    const int bps = 1;
    long long blocks_needed = (P.n_groups + wpb - 1) / wpb;
    int grid = (int)std::max<long long>(1, std::min<long long>(blocks_needed, (long long)bps * c->sm_count));
    if (any_noise) {
        if ((rc = c->scratch.ensure(sizeof(double) * (size_t)grid * wpb * (2 * n_nodes_max + n_coarse_max) * 32))) return rc;
        P.scratch = (double*)c->scratch.p;
    }
    const size_t n_mats = (size_t)n_points * nt;
    if ((rc = c->accum.ensure(sizeof(double) * n_mats * kNV))) return rc;
    if ((rc = c->counters.ensure(sizeof(unsigned long long) * 8))) return rc;
    CA_CUDA(cudaMemsetAsync(c->accum.p, 0, sizeof(double) * n_mats * kNV, c->stream));
    CA_CUDA(cudaMemsetAsync(c->counters.p, 0, sizeof(unsigned long long) * 8, c->stream));
    P.accum = (double*)c->accum.p;
    P.counters = (unsigned long long*)c->counters.p;

    int launches = 0;
    CA_CUDA(cudaEventRecord(c->ev[1], c->stream));
    if (n_all > 0) {
        cudaError_t le;
        if (ns_sel == 9) {
            le = beams_sel == 1 ? launch_l1_9_1(sweep, c->stream, P, grid) : (beams_sel == 2 ? launch_l1_9_2(sweep, c->stream, P, grid)
               : (beams_sel == 3 ? launch_l1_9_3(sweep, c->stream, P, grid) : launch_l1_9_0(sweep, c->stream, P, grid)));
        } else if (ns_sel == 15) {
            le = beams_sel == 1 ? launch_l1_15_1(sweep, c->stream, P, grid) : launch_l1_15_0(sweep, c->stream, P, grid);
        } else le = launch_l1_21_0(sweep, c->stream, P, grid);
        if (le != cudaSuccess) { set_error(std::string("k_lindblad1 launch: ") + cudaGetErrorString(le)); return CA_ERR_CUDA; }
        ++launches;
    }
    CA_CUDA(cudaEventRecord(c->ev[2], c->stream));
    // K6 finalize
    const double div = (out_flags & CA_OUT_SUMS) ? 1.0 : 1.0 / (double)(n_total > 0 ? n_total : std::max<long long>(1, n_per_point));
    const size_t out_mats = last_only ? (size_t)n_points : n_mats;
    double *d_rho, *d_rho2;
    if (out_flags & CA_OUT_DEVICE) { d_rho = rho; d_rho2 = rho2; }
    else {
        if ((rc = c->out.ensure(sizeof(double) * 100 * out_mats))) return rc;
        d_rho = (double*)c->out.p; d_rho2 = d_rho + 50 * out_mats;
    }
    k_finalize5<<<(unsigned)((out_mats * 2 + 127) / 128), 128, 0, c->stream>>>(P.accum, (long long)out_mats, last_only ? nt : 1, last_only ? nt - 1 : 0,
                                                                             div, d_rho, d_rho2);
    CA_CUDA(cudaGetLastError());
    ++launches;
    CA_CUDA(cudaEventRecord(c->ev[3], c->stream));
    ctx_set_last(c, P.counters, n_all, launches, 1);
    if (out_flags & CA_OUT_DEVICE) {   // asynchronous: the device-side status is read later (ca_collect_stats / ca_stats_to_device)
        if (stats) { std::memset(stats, 0, sizeof *stats); stats->n_traj = n_all; stats->n_launches = launches; }
        return CA_OK;
    }
    if ((rc = ensure_pinned(c, sizeof(double) * 100 * out_mats + 64))) return rc;
    double* hp = (double*)c->pinned;
    CA_CUDA(cudaMemcpyAsync(hp, d_rho, sizeof(double) * 100 * out_mats, cudaMemcpyDeviceToHost, c->stream));
    rc = collect_last(c, stats);   // synchronises
    if (rc == CA_ERR_CUDA || rc == CA_ERR_ARG) return rc;
    std::memcpy(rho, hp, sizeof(double) * 50 * out_mats);
    std::memcpy(rho2, hp + 50 * out_mats, sizeof(double) * 50 * out_mats);
    return rc;
}

}  // namespace ca

// ----------------------------------------------------------------------------------------------------
// C ABI
// ----------------------------------------------------------------------------------------------------
using namespace ca;

extern "C" {

int ca_version(void) { return CA_VERSION; }
const char* ca_last_error(void) { return get_error(); }

int ca_init(int device, ca_ctx** out) {
    if (!out) { set_error("null out pointer"); return CA_ERR_ARG; }
    *out = nullptr;
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n == 0) { set_error(std::string("no CUDA device: ") + cudaGetErrorString(e)); return CA_ERR_NODEVICE; }
    if (device < 0 || device >= n) { set_error("device index out of range"); return CA_ERR_ARG; }
    cudaDeviceProp prop;
    CA_CUDA(cudaGetDeviceProperties(&prop, device));
    if (prop.major != 10) {
        set_error(std::string("device ") + prop.name + " is not sm_100 (this library ships sm_100a code only)");
        return CA_ERR_NODEVICE;
    }
    CA_CUDA(cudaSetDevice(device));
    ca_ctx* c = new ca_ctx();
    c->device = device; c->sm_count = prop.multiProcessorCount;
    CA_CUDA(cudaStreamCreateWithFlags(&c->own_stream, cudaStreamNonBlocking));
    c->stream = c->own_stream;
    for (auto& ev : c->ev) CA_CUDA(cudaEventCreate(&ev));
    *out = c;
    return CA_OK;
}

int ca_finalize(ca_ctx* c) {
    if (!c) return CA_OK;
    cudaSetDevice(c->device);
    cudaStreamSynchronize(c->stream);
    DevBuf* bufs[] = {&c->scratch, &c->accum, &c->out, &c->counters, &c->in_samples, &c->in_ph[0], &c->in_ph[1], &c->in_ph[2], &c->in_ph[3],
                      &c->in_f, &c->in_models, &c->in_tspan, &c->misc};
    for (auto* b : bufs) b->release();
    for (auto& kv : c->extra) kv.second.release();
    if (c->pinned) cudaFreeHost(c->pinned);
    for (auto& ev : c->ev) if (ev) cudaEventDestroy(ev);
    if (c->own_stream) cudaStreamDestroy(c->own_stream);
    delete c;
    return CA_OK;
}

int ca_set_stream(ca_ctx* c, void* s) {
    if (!c) { set_error("null context"); return CA_ERR_ARG; }
    c->stream = (cudaStream_t)s;
    return CA_OK;
}

int ca_reset_stream(ca_ctx* c) {
    if (!c) { set_error("null context"); return CA_ERR_ARG; }
    c->stream = c->own_stream;
    return CA_OK;
}

int ca_synchronize(ca_ctx* c) {
    if (!c) { set_error("null context"); return CA_ERR_ARG; }
    CA_CUDA(cudaSetDevice(c->device));
    CA_CUDA(cudaStreamSynchronize(c->stream));
    return CA_OK;
}

int ca_collect_stats(ca_ctx* c, ca_stats* stats) {
    if (!c) { set_error("null context"); return CA_ERR_ARG; }
    CA_CUDA(cudaSetDevice(c->device));
    return collect_last(c, stats);
}

int ca_stats_to_device(ca_ctx* c, double* dev_out8) {
    if (!c || !dev_out8) { set_error("bad argument"); return CA_ERR_ARG; }
    if (!c->last_counters) { set_error("no ensemble call has run on this context"); return CA_ERR_ARG; }
    CA_CUDA(cudaSetDevice(c->device));
    k_export_counters<<<1, 32, 0, c->stream>>>(c->last_counters, c->last_n, c->last_kind, dev_out8);
    CA_CUDA(cudaGetLastError());
    return CA_OK;
}

int ca_measure_fp64_peak(ca_ctx* c, double* tflops, double* sm_mhz_est) {
    if (!c || !tflops) { set_error("bad argument"); return CA_ERR_ARG; }
    CA_CUDA(cudaSetDevice(c->device));
    int rc = c->misc.ensure(64);
    if (rc) return rc;
    const int iters = 2048, blocks = c->sm_count, threads = 256;   // one 8-warp CTA per SM
    float best = 1e30f;
    for (int rep = 0; rep < 6; ++rep) {
        CA_CUDA(cudaEventRecord(c->ev[0], c->stream));
        k_dfma_peak<<<blocks, threads, 0, c->stream>>>((double*)c->misc.p, iters, 1.0000001, 1e-9);
        CA_CUDA(cudaEventRecord(c->ev[1], c->stream));
        CA_CUDA(cudaStreamSynchronize(c->stream));
        float ms = 0;
        CA_CUDA(cudaEventElapsedTime(&ms, c->ev[0], c->ev[1]));
        if (rep > 0 && ms < best) best = ms;
    }
    const double flops = 2.0 * 8.0 * 32.0 * iters * (double)blocks * threads;
    *tflops = flops / (best * 1e-3) / 1e12;
    if (sm_mhz_est) *sm_mhz_est = (*tflops * 1e12) / (2.0 * 64.0 * c->sm_count) / 1e6;  // 64 DFMA/clk/SM
    return CA_OK;
}

int ca_rydberg1_run(ca_ctx* ctx, const ca_model* model, const double* tspan, int32_t nt, const double* psi0, int64_t n_traj,
                    int64_t traj_offset, int64_t n_total, uint64_t seed, const double* samples, const double* phases_red,
                    const double* phases_blue, const double* f, const double* amp_red, const double* amp_blue, int32_t nf,
                    const ca_ode_opts* ode, int32_t out_flags, double* rho, double* rho2, ca_stats* stats) {
    return run_r1(ctx, model, psi0, tspan, nt, 1, n_traj, traj_offset, n_total, seed, samples, phases_red, phases_blue, f, amp_red,
                  amp_blue, nf, ode, out_flags, false, rho, rho2, stats);
}

int ca_rydberg1_sweep(ca_ctx* ctx, const ca_model* models, const double* psi0s, const double* t_end, int32_t n_points,
                      int64_t n_traj_per_point, int64_t traj_offset, int64_t n_total, uint64_t seed, const double* f,
                      const double* amp_red, const double* amp_blue, int32_t nf, const ca_ode_opts* ode, int32_t out_flags, double* rho,
                      double* rho2, ca_stats* stats) {
    if (!t_end || n_points < 1) { set_error("bad argument"); return CA_ERR_ARG; }
    std::vector<double> ts((size_t)2 * n_points);
    for (int p = 0; p < n_points; ++p) { ts[2 * p] = 0.0; ts[2 * p + 1] = t_end[p]; }
    return run_r1(ctx, models, psi0s, ts.data(), 2, n_points, n_traj_per_point, traj_offset, n_total, seed, nullptr, nullptr, nullptr, f,
                  amp_red, amp_blue, nf, ode, out_flags, true, rho, rho2, stats);
}

int ca_release_recapture(ca_ctx* c, const double* trap, const double* atom, const double* tspan, int32_t nt, int64_t n,
                         int64_t traj_offset, int64_t n_total, double eps, uint64_t seed, const double* samples, int32_t out_flags,
                         double* probs, uint8_t* indicators, ca_stats* stats) {
    if (!c || !trap || !atom || !tspan || nt < 1 || n < 0 || !probs) { set_error("bad argument"); return CA_ERR_ARG; }
    if ((out_flags & CA_OUT_DEVICE) && indicators) { set_error("CA_OUT_DEVICE: indicators must be NULL"); return CA_ERR_ARG; }
    CA_CUDA(cudaSetDevice(c->device));
    ca_model m;
    std::memset(&m, 0, sizeof m);
    m.atom_m = atom[0]; m.atom_T = atom[1]; m.trap_U0 = trap[0]; m.trap_w0 = trap[1]; m.trap_z0 = trap[2];
    m.red.kind = m.blue.kind = CA_BEAM_UNIFORM; m.blue.w0 = m.blue.z0 = 1.0;
    DevModel d;
    int rc = derive_model(m, nullptr, 0, 0.0, 1.0, &d);
    if (rc) return rc;
    SamplerConsts sc;
    for (int q = 0; q < 6; ++q) sc.sig[q] = d.sig[q];
    sc.U0 = d.U0; sc.w0 = d.w0; sc.z0 = d.z0; sc.mfac = d.mfac; sc.ecut = d.ecut;
    CA_CUDA(cudaEventRecord(c->ev[0], c->stream));
    if ((rc = upload(c, c->in_tspan, tspan, sizeof(double) * nt))) return rc;
    const double* d_samples = nullptr;
    if (samples && n > 0) { if ((rc = upload(c, c->in_samples, samples, sizeof(double) * 6 * (size_t)n))) return rc; d_samples = (const double*)c->in_samples.p; }
    if ((rc = c->counters.ensure(sizeof(unsigned long long) * ((size_t)nt + 8)))) return rc;
    CA_CUDA(cudaMemsetAsync(c->counters.p, 0, sizeof(unsigned long long) * ((size_t)nt + 8), c->stream));
    unsigned long long* d_counts = (unsigned long long*)c->counters.p;
    unsigned char* d_ind = nullptr;
    if (indicators && n > 0) { if ((rc = c->out.ensure((size_t)n * nt))) return rc; d_ind = (unsigned char*)c->out.p; }
    const double thresh = trap[0] * (1.0 - eps);
    CA_CUDA(cudaEventRecord(c->ev[1], c->stream));
    if (n > 0) {
        k_release_recapture<<<(unsigned)((n + 255) / 256), 256, 0, c->stream>>>(sc, (const double*)c->in_tspan.p, nt, n, traj_offset, seed, thresh,
                                                                                d_samples, d_counts, d_ind, d_counts + nt);
        CA_CUDA(cudaGetLastError());
    }
    CA_CUDA(cudaEventRecord(c->ev[2], c->stream));
    if (out_flags & CA_OUT_DEVICE) {   // counts (or means) as doubles straight into the caller's device buffer; no synchronisation
        const double sc_ = (out_flags & CA_OUT_SUMS) ? 1.0 : 1.0 / (double)(n_total > 0 ? n_total : n);
        k_counts_to_double<<<(nt + 127) / 128, 128, 0, c->stream>>>(d_counts, nt, sc_, probs);
        CA_CUDA(cudaGetLastError());
        CA_CUDA(cudaEventRecord(c->ev[3], c->stream));
        if (stats) { std::memset(stats, 0, sizeof *stats); stats->n_traj = n; stats->n_launches = (n > 0 ? 1 : 0) + 1; }
        return CA_OK;
    }
    if ((rc = ensure_pinned(c, sizeof(unsigned long long) * ((size_t)nt + 8)))) return rc;
    CA_CUDA(cudaMemcpyAsync(c->pinned, d_counts, sizeof(unsigned long long) * ((size_t)nt + 8), cudaMemcpyDeviceToHost, c->stream));
    if (d_ind) CA_CUDA(cudaMemcpyAsync(indicators, d_ind, (size_t)n * nt, cudaMemcpyDeviceToHost, c->stream));
    CA_CUDA(cudaEventRecord(c->ev[3], c->stream));
    CA_CUDA(cudaStreamSynchronize(c->stream));
    const unsigned long long* hc = (const unsigned long long*)c->pinned;
    const double div = (out_flags & CA_OUT_SUMS) ? 1.0 : (double)(n_total > 0 ? n_total : n);  // N = 0 gives 0/0 = NaN like `recapture ./ N` (basic_experiments.jl:80)
    for (int j = 0; j < nt; ++j) probs[j] = (double)hc[j] / div;
    if (stats) {
        std::memset(stats, 0, sizeof *stats);
        stats->n_traj = n; stats->n_sample_attempts = (int64_t)hc[nt]; stats->n_launches = n > 0 ? 1 : 0;
        float ms = 0;
        cudaEventElapsedTime(&ms, c->ev[1], c->ev[2]); stats->kernel_ms = ms;
        cudaEventElapsedTime(&ms, c->ev[0], c->ev[3]); stats->total_ms = ms;
    }
    return CA_OK;
}

int ca_sample_atoms(ca_ctx* c, const double* trap, const double* atom, int64_t n, int64_t traj_offset, uint64_t seed, int32_t stream_id,
                    double eps, double* samples_out, int64_t* attempts_out) {
    if (!c || !trap || !atom || n < 0 || !samples_out) { set_error("bad argument"); return CA_ERR_ARG; }
    CA_CUDA(cudaSetDevice(c->device));
    ca_model m;
    std::memset(&m, 0, sizeof m);
    m.atom_m = atom[0]; m.atom_T = atom[1]; m.trap_U0 = trap[0]; m.trap_w0 = trap[1]; m.trap_z0 = trap[2];
    m.red.kind = m.blue.kind = CA_BEAM_UNIFORM; m.blue.w0 = m.blue.z0 = 1.0;
    DevModel d;
    int rc = derive_model(m, nullptr, 0, 0.0, 1.0, &d);
    if (rc) return rc;
    SamplerConsts sc;
    for (int q = 0; q < 6; ++q) sc.sig[q] = d.sig[q];
    sc.U0 = d.U0; sc.w0 = d.w0; sc.z0 = d.z0; sc.mfac = d.mfac; sc.ecut = trap[0] * (1.0 - eps);
    if (n == 0) { if (attempts_out) *attempts_out = 0; return CA_OK; }
    if ((rc = c->out.ensure(sizeof(double) * 6 * (size_t)n))) return rc;
    if ((rc = c->counters.ensure(64))) return rc;
    CA_CUDA(cudaMemsetAsync(c->counters.p, 0, 64, c->stream));
    k_sample_atoms<<<(unsigned)((n + 127) / 128), 128, 0, c->stream>>>(sc, n, traj_offset, seed, (unsigned)stream_id, (double*)c->out.p,
                                                                       (unsigned long long*)c->counters.p);
    CA_CUDA(cudaGetLastError());
    CA_CUDA(cudaMemcpyAsync(samples_out, c->out.p, sizeof(double) * 6 * (size_t)n, cudaMemcpyDeviceToHost, c->stream));
    unsigned long long att = 0;
    CA_CUDA(cudaMemcpyAsync(&att, c->counters.p, 8, cudaMemcpyDeviceToHost, c->stream));
    CA_CUDA(cudaStreamSynchronize(c->stream));
    if (attempts_out) *attempts_out = (int64_t)att;
    return CA_OK;
}

int ca_sample_atoms_metropolis(ca_ctx* c, const double* trap, const double* atom, int64_t n, int64_t freq, int64_t skip, int32_t harmonic,
                               double eps, int64_t n_chains, uint64_t seed, int32_t stream_id, double* samples_out, double* acc_rate_out) {
    if (!c || !trap || !atom || n < 0 || freq < 1 || skip < 0 || !samples_out) { set_error("bad argument"); return CA_ERR_ARG; }
    CA_CUDA(cudaSetDevice(c->device));
    if (n_chains <= 0) n_chains = std::min<int64_t>(n, (int64_t)c->sm_count * 256);
    n_chains = std::max<int64_t>(1, std::min<int64_t>(n_chains, std::max<int64_t>(1, n)));
    if (n_chains > 0xffffffffLL || n * freq + skip > 0xffffffffLL) { set_error("chain / step index exceeds 32 bits"); return CA_ERR_ARG; }
    MetroConsts mc;
    const double U0 = trap[0], w0 = trap[1], z0 = trap[2], m = atom[0], T = atom[1];
    const double vc = std::sqrt(1.380649e-23 * 1e-6 / 1.66053906660e-27);   // vconst (atom_sampler.jl:3-10), as in ca_model.cpp
    const double vstep = vc * std::sqrt(T / m), rstep = std::sqrt(T / U0) / 2.0;   // atom_sampler.jl:99-101
    mc.sd[0] = mc.sd[1] = w0 * rstep; mc.sd[2] = z0 * std::sqrt(2.0) * rstep; mc.sd[3] = mc.sd[4] = mc.sd[5] = vstep;
    mc.U0 = U0; mc.w0 = w0; mc.z0 = z0; mc.mfac = m / (vc * vc); mc.T = T; mc.ecut = U0 * (1.0 - eps); mc.harmonic = harmonic ? 1 : 0;
    if (acc_rate_out) *acc_rate_out = 0.0;
    if (n == 0) return CA_OK;
    int rc;
    if ((rc = c->out.ensure(sizeof(double) * 6 * (size_t)n))) return rc;
    if ((rc = c->counters.ensure(64))) return rc;
    CA_CUDA(cudaMemsetAsync(c->counters.p, 0, 64, c->stream));
    k_metropolis<<<(unsigned)((n_chains + 127) / 128), 128, 0, c->stream>>>(mc, seed, (unsigned)stream_id, n_chains, n, freq, skip, (double*)c->out.p,
                                                                            (unsigned long long*)c->counters.p);
    CA_CUDA(cudaGetLastError());
    CA_CUDA(cudaMemcpyAsync(samples_out, c->out.p, sizeof(double) * 6 * (size_t)n, cudaMemcpyDeviceToHost, c->stream));
    unsigned long long hc[2] = {0, 0};
    CA_CUDA(cudaMemcpyAsync(hc, c->counters.p, 16, cudaMemcpyDeviceToHost, c->stream));
    CA_CUDA(cudaStreamSynchronize(c->stream));
    // acc_rate / (N*freq + skip) with one chain (atom_sampler.jl:119); with K chains: accepted / (transitions + K initial states)
    if (acc_rate_out) *acc_rate_out = (double)hc[0] / (double)(hc[1] + (unsigned long long)n_chains);
    return CA_OK;
}

int ca_phase_noise(ca_ctx* c, const double* tnodes, int32_t n_nodes, const double* f, const double* amp, int32_t nf, int64_t n,
                   int64_t traj_offset, uint64_t seed, int32_t stream_id, const double* phases, double* phi_out) {
    if (!c || !tnodes || n_nodes < 2 || !f || !amp || nf < 1 || n < 0 || !phi_out) { set_error("bad argument"); return CA_ERR_ARG; }
    CA_CUDA(cudaSetDevice(c->device));
    // the device generator works on the uniform grid the reference uses (rydberg_model.jl:216): check it
    const double dtn = (tnodes[n_nodes - 1] - tnodes[0]) / (n_nodes - 1);
    if (tnodes[0] != 0.0) { set_error("tnodes must start at 0"); return CA_ERR_UNSUPPORTED; }
    for (int j = 0; j < n_nodes; ++j)
        if (std::fabs(tnodes[j] - j * dtn) > 1e-9 * std::fabs(tnodes[n_nodes - 1])) { set_error("tnodes must be uniformly spaced"); return CA_ERR_UNSUPPORTED; }
    if (n == 0) return CA_OK;
    int rc;
    std::vector<NoiseK> hk((size_t)nf);
    const int coarse = choose_noise_coarse(f, amp, nullptr, nf, dtn, n_nodes);
    fill_noise_k(f, amp, nf, noise_c(coarse) * dtn, hk.data());
    if ((rc = upload(c, c->in_f, hk.data(), sizeof(NoiseK) * (size_t)nf))) return rc;
    CA_CUDA(cudaStreamSynchronize(c->stream));
    const double* d_ph = nullptr;
    if (phases) { if ((rc = upload(c, c->in_ph[0], phases, sizeof(double) * (size_t)nf * n))) return rc; d_ph = (const double*)c->in_ph[0].p; }
    if ((rc = c->out.ensure(sizeof(double) * (size_t)n * n_nodes))) return rc;
    if ((rc = c->scratch.ensure(sizeof(double) * (size_t)n * noise_coarse_count(n_nodes, coarse)))) return rc;
    k_phase_noise<<<(unsigned)((n + 127) / 128), 128, 0, c->stream>>>((const NoiseK*)c->in_f.p, nf, n_nodes, dtn, coarse, d_ph, n, traj_offset, seed,
                                                                     (unsigned)stream_id, (double*)c->out.p, (double*)c->scratch.p);
    CA_CUDA(cudaGetLastError());
    CA_CUDA(cudaMemcpyAsync(phi_out, c->out.p, sizeof(double) * (size_t)n * n_nodes, cudaMemcpyDeviceToHost, c->stream));
    CA_CUDA(cudaStreamSynchronize(c->stream));
    return CA_OK;
}

int ca_eval_beam(ca_ctx* c, const ca_beam* beam, const double* xyz, int64_t n, double* out) {
    if (!c || !beam || !xyz || n < 0 || !out) { set_error("bad argument"); return CA_ERR_ARG; }
    CA_CUDA(cudaSetDevice(c->device));
    DevBeam b;
    int rc = derive_beam(*beam, &b);
    if (rc) return rc;
    if (n == 0) return CA_OK;
    if ((rc = upload(c, c->in_samples, xyz, sizeof(double) * 3 * (size_t)n))) return rc;
    if ((rc = c->out.ensure(sizeof(double) * 2 * (size_t)n))) return rc;
    k_eval_beam<<<(unsigned)((n + 127) / 128), 128, 0, c->stream>>>(b, (const double*)c->in_samples.p, n, (double*)c->out.p);
    CA_CUDA(cudaGetLastError());
    CA_CUDA(cudaMemcpyAsync(out, c->out.p, sizeof(double) * 2 * (size_t)n, cudaMemcpyDeviceToHost, c->stream));
    CA_CUDA(cudaStreamSynchronize(c->stream));
    return CA_OK;
}

int ca_thermal_average(ca_ctx* c, int32_t kind, const ca_model* model, int64_t n, int64_t traj_offset, uint64_t seed, double eps,
                       const double* samples1, const double* samples2, double* sums_out, int64_t* attempts_out) {
    if (!c || !model || n < 0 || !sums_out || (kind != CA_AVG_TWOPHOTON && kind != CA_AVG_RM6)) { set_error("bad argument"); return CA_ERR_ARG; }
    if (kind == CA_AVG_RM6 && ((samples1 != nullptr) != (samples2 != nullptr))) { set_error("supply both sample sets or neither"); return CA_ERR_ARG; }
    CA_CUDA(cudaSetDevice(c->device));
    DevModel d;
    int rc = derive_model(*model, nullptr, 0, 0.0, 1.0, &d);
    if (rc) return rc;
    AvgConsts ac;
    ac.red = d.red; ac.blue = d.blue; ac.aDx = d.aDx; ac.aDz = d.aDz; ac.kind = kind;
    std::memcpy(ac.center, d.center, sizeof ac.center);
    SamplerConsts sc;
    for (int q = 0; q < 6; ++q) sc.sig[q] = d.sig[q];
    sc.U0 = d.U0; sc.w0 = d.w0; sc.z0 = d.z0; sc.mfac = d.mfac; sc.ecut = d.U0 * (1.0 - (eps > 0 ? eps : 1e-2));
    sums_out[0] = sums_out[1] = 0.0;
    if (attempts_out) *attempts_out = 0;
    if (n == 0) return CA_OK;
    const double *d1 = nullptr, *d2 = nullptr;
    if (samples1) { if ((rc = upload(c, c->in_samples, samples1, sizeof(double) * 6 * (size_t)n))) return rc; d1 = (const double*)c->in_samples.p; }
    if (samples2 && kind == CA_AVG_RM6) { if ((rc = upload(c, c->in_ph[0], samples2, sizeof(double) * 6 * (size_t)n))) return rc; d2 = (const double*)c->in_ph[0].p; }
    const int nb = (int)std::min<int64_t>((n + 127) / 128, (int64_t)c->sm_count * 8);
    if ((rc = c->misc.ensure(sizeof(double) * 2 * (size_t)(nb + 1)))) return rc;
    if ((rc = c->counters.ensure(64))) return rc;
    CA_CUDA(cudaMemsetAsync(c->counters.p, 0, 64, c->stream));
    double* part = (double*)c->misc.p;
    k_thermal_avg<<<nb, 128, 0, c->stream>>>(ac, sc, n, traj_offset, seed, d1, d2, part, (unsigned long long*)c->counters.p);
    CA_CUDA(cudaGetLastError());
    k_sum_partials<<<1, 32, 0, c->stream>>>(part, nb, part + 2 * (size_t)nb);
    CA_CUDA(cudaGetLastError());
    unsigned long long att = 0;
    CA_CUDA(cudaMemcpyAsync(sums_out, part + 2 * (size_t)nb, 2 * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
    CA_CUDA(cudaMemcpyAsync(&att, c->counters.p, 8, cudaMemcpyDeviceToHost, c->stream));
    CA_CUDA(cudaStreamSynchronize(c->stream));
    if (attempts_out) *attempts_out = (int64_t)att;
    return CA_OK;
}

}  // extern "C"
