// ca_lindblad1.cuh -- K1+K2+K3: the fused one-atom ensemble kernel (Philox atom sampler, phase-noise table generation, DP5
// Lindblad integrator; one trajectory per thread, state in registers; observables reduced by warp shuffles and accumulated
// with fp64 RED atomics).  Replaces the ensemble loop of simulation (src/rydberg_model.jl:206-264) and of
// simulation_blue_intens (src/rydberg_model_arb_bms.jl:15-86).  Every (NS, BEAMS) variant is compiled in its own
// translation unit (ca_l1_inst.cu) so that the library builds in parallel.
#pragma once
#include <cuda_runtime.h>

#include "ca_device.cuh"

namespace ca {

// ----------------------------------------------------------------------------------------------------
// device helpers
// ----------------------------------------------------------------------------------------------------
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// ----------------------------------------------------------------------------------------------------
// K1+K2+K3: fused one-atom ensemble kernel
// ----------------------------------------------------------------------------------------------------
// Threads per CTA (one CTA per SM).  NS = 9: 256 (the register file's limit at 255 registers per thread).  NS = 15 (a |0> or |l> component in
// ψ0: four of the six fidelity basis states) keeps 6 x 15 stage derivatives per thread in shared memory, so shared memory sets the limit: the
// free-flight Gaussian-pair variant (124 doubles per thread with the compact coefficient / anchor areas) runs 224 threads = 7 warps (222 KB);
// its ~0.4 KB of register spills per thread then live in L2, and still 7 warps beat 4 warps with the spills in L1 (489k -> 536k traj/s;
// 6 warps: 510k; profiles/r2/ab_ns15_threads.txt).  The generic NS = 15 variant (140 doubles) and NS = 21 (126, no anchors) stay at 128.
#ifndef CA_NS15_THREADS
#define CA_NS15_THREADS 224
#endif
template <int NS, int BEAMS> __host__ __device__ constexpr int threads_for() { return NS == 9 ? 256 : ((NS == 15 && BEAMS == 1) ? CA_NS15_THREADS : 128); }
inline int threads_for_rt(int ns, int beams) { return ns == 9 ? 256 : ((ns == 15 && beams == 1) ? CA_NS15_THREADS : 128); }
template <int BEAMS> __host__ __device__ constexpr int cb_doubles() { return 5 * ((BEAMS == 1 || BEAMS == 2) ? 4 : 6); }   // Lane1::kCB per stage
template <int NS, int BEAMS> __host__ __device__ constexpr int smem_doubles_per_thread() {   // k slots + stage coefficients + anchors (NS <= 15)
    return 6 * NS + cb_doubles<BEAMS>() + (NS <= 15 ? (BEAMS == 1 ? 14 : 20) : 0);   // (Lane1::kAB slots per beam)
}

struct R1Params {
    DevModel model;            // single-point runs: lives in the constant bank with the other params
    const DevModel* models;    // sweeps: one per point
    const double* tspans;      // [n_points][nt]
    const double* samples;     // NULL or [n_points*n_per_point][6]
    const double* phases_red;  // NULL or [.][nf]
    const double* phases_blue;
    const NoiseK* nk_red;      // [n_points][nf] per-frequency recurrence constants (dtn depends on the point's t_end)
    const NoiseK* nk_blue;
    double* scratch;           // per-warp phase-noise tables [2][n_nodes_max][32] + coarse staging [n_coarse_max][32]
    double* accum;             // [n_points][nt][kNV]
    unsigned long long* counters;  // n_rhs, n_accept, n_reject, n_attempts, status(max |err|)
    OdeOpts ode;
    long long n_per_point, traj_offset, groups_per_point, n_groups;
    unsigned long long seed;
    int n_points, nt, nf, n_nodes_max, n_coarse_max;
};

template <int NS, bool SWEEP, int BEAMS>
__global__ void __launch_bounds__(threads_for<NS, BEAMS>(), 1) k_lindblad1(const __grid_constant__ R1Params P) {
    constexpr int T = threads_for<NS, BEAMS>();
    extern __shared__ __align__(16) unsigned char smem_raw[];
    double* sk = reinterpret_cast<double*>(smem_raw);                                   // [6][NS][T] stage derivatives
    double* scb = sk + 6 * NS * T;                                                      // [5][6][T] stage coefficients
    double* sab = scb + cb_doubles<BEAMS>() * T;                                                         // [2][2][5][T] beam anchors (NS <= 15)
    DevModel* sm_models = reinterpret_cast<DevModel*>(sk + smem_doubles_per_thread<NS, BEAMS>() * T);  // [warps] (sweeps only)
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#ifndef CA_WARP_MAJOR
#define CA_WARP_MAJOR 1
#endif
    // Warp-major numbering: the groups of a round go to warp 0 of every CTA first, then warp 1, ...  A PARTIAL last round (strong
    // scaling: 10^6 / 8 GPUs = 3.3 rounds) is then spread over all SMs with fewer resident warps each, which finish sooner than a
    // few fully loaded SMs would.
    // Sweeps keep the block-major numbering: consecutive groups belong to the same grid point, so the warps of a CTA, which walk the list in
    // lock-step, integrate windows of the same length (warp-major mixed pulse times 10x apart in one CTA: S1 -7 %).
    const long long wg = (CA_WARP_MAJOR && !SWEEP) ? (long long)warp * gridDim.x + blockIdx.x : (long long)blockIdx.x * (T / 32) + warp;
    const long long W = (long long)gridDim.x * (T / 32);
    long long tot_rhs = 0, tot_acc = 0, tot_rej = 0, tot_att = 0, tot_exact = 0;
    long long cyc_noise = 0, cyc_all = -clock64();
    int worst = 0;
    // All warps of a CTA walk the group list in lock-step (one barrier per 32-trajectory group): measured, the warps
    // otherwise drift apart over a long ensemble and the instruction cache (the step is ~36 KB of code) thrashes:
    // 85.8 -> 96.8 ms per wave between 3 and 27 waves.
    const long long n_iter = (P.n_groups + W - 1) / W;
    for (long long it = 0; it < n_iter; ++it) {
        const long long g_raw = wg + it * W;
        __syncthreads();
        // (a warp without a group in the last round still walks through K1 / K2 with valid = false: both barriers of the round are
        // executed by every thread of the CTA)
        const bool active = g_raw < P.n_groups;
        const long long g = active ? g_raw : 0;
        const int pt = (int)(g / P.groups_per_point);
        const long long il = (g - (long long)pt * P.groups_per_point) * 32 + lane;
        const bool valid = active && il < P.n_per_point;
        const long long li = (long long)pt * P.n_per_point + il;  // index into host-supplied arrays
        const uint64_t gi = (uint64_t)(P.traj_offset + li);
        if (SWEEP) {
            __syncwarp();
            const int nw = (int)(sizeof(DevModel) / sizeof(double));
            const double* src = reinterpret_cast<const double*>(P.models + pt);
            double* dst = reinterpret_cast<double*>(&sm_models[warp]);
            for (int q = lane; q < nw; q += 32) dst[q] = src[q];
            __syncwarp();
        }
        const DevModel* mp;
        if constexpr (SWEEP) mp = &sm_models[warp]; else mp = &P.model;
        const DevModel& m = *mp;
        Rates rt;
        rt.G0 = m.G0; rt.G1 = m.G1; rt.Gl = m.Gl; rt.Gr = m.Gr; rt.Gp = m.G0 + m.G1 + m.Gl; rt.gp2 = 0.5 * rt.Gp; rt.gr2 = 0.5 * m.Gr;
        const double* tsp = P.tspans + (size_t)pt * P.nt;
        // ---- K1: atom sample
        double smp[6] = {0, 0, 0, 0, 0, 0};
        if (valid && m.atom_motion) {
            if (P.samples) {
#pragma unroll
                for (int q = 0; q < 6; ++q) smp[q] = P.samples[li * 6 + q];
            } else {
                SamplerConsts sc;
#pragma unroll
                for (int q = 0; q < 6; ++q) sc.sig[q] = m.sig[q];
                sc.U0 = m.U0; sc.w0 = m.w0; sc.z0 = m.z0; sc.mfac = m.mfac; sc.ecut = m.ecut;
                int att = sample_atom_literal(sc, P.seed, gi, 0u, smp);
                if (att < 0) worst = max(worst, -CA_ERR_NONFINITE); else tot_att += att;
            }
        }
        // ---- K2: phase-noise tables for this warp's 32 trajectories
        double* tab = nullptr;
        if (m.laser_noise && P.nf > 0) {
            cyc_noise -= clock64();
            tab = P.scratch + (size_t)wg * (2 * P.n_nodes_max + P.n_coarse_max) * 32 + lane;
            // the plan's Lagrange weights ((c - 1) x p <= 496 doubles) staged in the stage-derivative area, which is idle during K2: the
            // interpolation reads them as uniform shared-memory loads instead of run-time-indexed constant loads
            double* wsm = sk + warp * 512;
            const int nwt = (noise_c(m.noise_coarse) - 1) * noise_p(m.noise_coarse);
            if (nwt > 0) {
                const double* wr = noise_weight_rows(m.noise_coarse);
                for (int q = lane; q < nwt; q += 32) wsm[q] = wr[q];
            }
            __syncwarp();
            if (valid) {
                double* tmp = tab + (size_t)2 * P.n_nodes_max * 32;
                gen_noise_trace_c<kNoiseChunk>(P.nk_red + (size_t)pt * P.nf, P.nf, m.n_nodes, m.dtn, m.noise_coarse,
                                               P.phases_red ? P.phases_red + li * P.nf : nullptr, P.seed, gi, 2u, tab, 32, tmp, 32, wsm);
                gen_noise_trace_c<kNoiseChunk>(P.nk_blue + (size_t)pt * P.nf, P.nf, m.n_nodes, m.dtn, m.noise_coarse,
                                               P.phases_blue ? P.phases_blue + li * P.nf : nullptr, P.seed, gi, 3u, tab + (size_t)m.n_nodes * 32, 32, tmp, 32, wsm);
            }
            __syncwarp();
            cyc_noise += clock64();
        }
        __syncthreads();   // every warp's K2 is over before any warp's K3 writes stage derivatives over the staged weights
        if (!active) continue;
        // ---- K3: integrate
        Lane1<NS, BEAMS> L;
        L.init(m, rt, P.ode, smp, tab, 32, tsp[0], sk + threadIdx.x, T, scb + threadIdx.x, NS <= 15 ? sab + threadIdx.x : nullptr);
        if (!valid) L.t = 1e300;
        const double tend = tsp[P.nt - 1];
        unsigned n_att = 0;
        double* acc_pt = P.accum + (size_t)pt * P.nt * kNV;
        for (int j = 0; j < P.nt; ++j) {
            const double T = tsp[j];
            const double tstop = P.ode.dense ? tend : T;
            while (true) {
                const bool need = L.t < T;   // invalid and failed lanes are parked at t = +inf
                if (!__any_sync(0xffffffffu, need)) break;
                if (need) L.attempt(m, rt, P.ode, tstop, (n_att & (Lane1<NS, BEAMS>::kEvery - 1)) == 0);   // all lanes re-anchor their beam model together
                ++n_att;
            }
            double v[kNV];
            L.state_at(m, T, v);
            second_moment<(NS > 9)>(v, m.rho2_mode, v + 25);
            const bool ok = valid && L.status == 0;
            double mine0 = 0.0, mine1 = 0.0;
#pragma unroll
            for (int q = 0; q < kNV; ++q) {
                // NS = 9 (ψ0 inside the core block): only the five populations and the three core coherences of ρ and of ρ·ρ are non-zero
                // (packed indices 0-4, 13-16, 19-20 of each half), the other 28 sums are skipped at compile time
                if (NS == 9) { const int r = q % 25; if (!(r < 5 || (r >= 13 && r <= 16) || r == 19 || r == 20)) continue; }
                double s = warp_sum(ok ? v[q] : 0.0);
                if (q < 32) { if (lane == q) mine0 = s; }
                else { if (lane == q - 32) mine1 = s; }
            }
            atomicAdd(acc_pt + (size_t)j * kNV + lane, mine0);
            if (lane < kNV - 32) atomicAdd(acc_pt + (size_t)j * kNV + 32 + lane, mine1);
        }
        if (valid) { tot_rhs += L.n_rhs; tot_acc += L.n_acc; tot_rej += L.n_rej; tot_exact += L.n_exact; if (L.status) worst = max(worst, -L.status); }
    }
    // ---- statistics
    tot_rhs = (long long)warp_sum((double)tot_rhs); tot_acc = (long long)warp_sum((double)tot_acc);
    tot_rej = (long long)warp_sum((double)tot_rej); tot_att = (long long)warp_sum((double)tot_att);
    tot_exact = (long long)warp_sum((double)tot_exact);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) worst = max(worst, __shfl_xor_sync(0xffffffffu, worst, o));
    if (lane == 0) {
        atomicAdd(P.counters + 0, (unsigned long long)tot_rhs);
        atomicAdd(P.counters + 1, (unsigned long long)tot_acc);
        atomicAdd(P.counters + 2, (unsigned long long)tot_rej);
        atomicAdd(P.counters + 3, (unsigned long long)tot_att);
        if (worst) atomicMax(P.counters + 4, (unsigned long long)worst);
        atomicAdd(P.counters + 5, (unsigned long long)cyc_noise);               // warp-cycles in phase-noise generation
        atomicAdd(P.counters + 6, (unsigned long long)(cyc_all + clock64()));   // warp-cycles in the whole kernel
        atomicAdd(P.counters + 7, (unsigned long long)tot_exact);               // steps on the exact coefficient path
    }
}


template <int NS, bool SWEEP, int BEAMS>
static cudaError_t launch_lindblad1(cudaStream_t stream, const R1Params& P, int grid) {
    constexpr int T = threads_for<NS, BEAMS>();
    const size_t smem = sizeof(double) * smem_doubles_per_thread<NS, BEAMS>() * T + (SWEEP ? sizeof(DevModel) * (T / 32) : 0);
    cudaError_t e = cudaFuncSetAttribute(k_lindblad1<NS, SWEEP, BEAMS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    k_lindblad1<NS, SWEEP, BEAMS><<<grid, T, smem, stream>>>(P);
    return cudaGetLastError();
}

// one function per (NS, BEAMS) translation unit: ca_l1_inst.cu compiled with -DCA_INST_NS=.. -DCA_INST_BEAMS=..
#define CA_L1_DECL(NS, BEAMS) cudaError_t launch_l1_##NS##_##BEAMS(bool sweep, cudaStream_t stream, const R1Params& P, int grid)
CA_L1_DECL(9, 0); CA_L1_DECL(9, 1); CA_L1_DECL(9, 2); CA_L1_DECL(9, 3); CA_L1_DECL(15, 0); CA_L1_DECL(15, 1); CA_L1_DECL(21, 0);

}  // namespace ca
