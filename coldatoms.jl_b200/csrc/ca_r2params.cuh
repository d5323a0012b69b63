// ca_r2params.cuh -- kernel parameters, context accessors and the phase-noise table generator shared by the two two-atom kernels
// (ca_lindblad2.cu: blocked, ca_lindblad2f.cu: full form).
#pragma once
#include <cuda_runtime.h>

#include "ca_host.h"
#include "ca_lindblad2f.cuh"

namespace ca {

cudaStream_t ctx_stream(ca_ctx* c);
int ctx_sm_count(ca_ctx* c);
int ctx_device(ca_ctx* c);
void* ctx_buf(ca_ctx* c, int which, size_t bytes);  // grow-only device buffers owned by the context
cudaEvent_t ctx_event(ca_ctx* c, int i);
void ctx_set_last(ca_ctx* c, unsigned long long* counters, long long n, int launches, int kind);
int collect_last(ca_ctx* c, ca_stats* stats);

struct R2Params {
    DevModel model;
    const double* rho0m;     // [32][kS2] initial state in the per-lane real layout
    double tr0;              // trace of ρ0
    const double* tspan;
    const double* samples1;  // NULL or [n][6] (before the centre shift)
    const double* samples2;
    const double* phases[4]; // NULL or [n][nf]: red1, blue1, red2, blue2
    const double* f;
    const double* amp_red;
    const double* amp_blue;
    double* scratch;         // per warp: [4][n_nodes] phase-noise tables
    double* accum;           // per warp: [nt][2 (ρ, ρ²)][kAccSlots][32]
    unsigned long long* counters;
    const unsigned* order;   // NULL or [n_traj]: trajectory indices, most expensive first (k_order2_*)
    OdeOpts ode;
    long long n_traj, traj_offset;
    unsigned long long seed;
    int nt, nf;
    // full-form kernel only
    DevModel2x x;            // legacy parameterisations (ca_model.two_atom_flags)
    const double* rho0f;     // [32][kF] columns of M(ρ0)
    int nt_seg1;             // 0, or the number of leading tspan points that form the first time segment
    // sweeps (ca_rydberg2_sweep; blocked kernel only): rho0m is [n_points][32][kS2], tspan is [n_points][nt]
    const DevModel* models;  // NULL, or one model per point
    const double* tr0s;      // [n_points]
    const long long* first;  // [n_points + 1] prefix sums of the per-point trajectory counts
    double* accum_pt;        // [n_points][nt][2][kAccSlots][32] sums per point (atomics)
    int n_points, n_nodes_max;
};

int run_r2_full(ca_ctx* c, const ca_model* model, const double* tspan, int32_t nt, int32_t nt_seg1, const double* psi0, const double* rho0,
                int64_t n_traj, int64_t traj_offset, int64_t n_total, uint64_t seed, const double* samples1, const double* samples2,
                const double* const* phases4, const double* f, const double* amp_red, const double* amp_blue, int32_t nf,
                const ca_ode_opts* ode, int32_t out_flags, double* rho, double* rho2, ca_stats* stats);

#if defined(__CUDACC__)
// ---- phase-noise tables of one trajectory: lane = chunk of 32 nodes, all four traces (cz_model.jl:41-46) ----
__device__ __forceinline__ void gen_noise_tables2(const R2Params& P, const DevModel& m, long long tr, uint64_t gi, int lane, double* tab_w) {
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

#endif

}  // namespace ca
