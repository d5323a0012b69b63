// ca_lindblad2.cu -- K4: two-atom CZ ensemble (simulation_czlp, src/cz_model.jl:107-171) on sm_100a.
//
// One trajectory per WARP; the per-lane code (state layout, RHS, DP5, observables) lives in ca_lindblad2.cuh and is
// shared with the CPU test harness.  This file holds the kernel shell (Philox samples, phase-noise tables, the
// trajectory loop, statistics), the K6 finalize kernel and the C ABI entry ca_rydberg2_run.
#include <cuda_runtime.h>

#include <algorithm>
#include <cstdio>
#include <cstring>
#include <vector>

#include "ca_host.h"
#include "ca_lindblad2.cuh"
#include "ca_r2params.cuh"

namespace ca {


#ifndef CA_WARPS2
#define CA_WARPS2 8
#endif
constexpr int kWarps2 = CA_WARPS2;   // warps (trajectories) per CTA, one CTA per SM


struct DevWarp {
    int lane;
    double *xm, *xp, *kb;
    Coef2* coef;
    double* smp;
    double* anc;   // per-lane anchor slots anc[slot * 32]
#ifdef CA_TIMERS
    long long tm[4] = {0, 0, 0, 0}, tl = 0;   // cycles: 0 noise+sampling, 1 coefficients, 2 stage-7 RHS, 3 everything else
    __device__ __forceinline__ void tick(int i) { const long long c = clock64(); tm[i] += c - tl; tl = c; }
#endif
    __device__ __forceinline__ void sync() { __syncwarp(); }
    __device__ __forceinline__ double sum(double v) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
        return v;
    }
    __device__ __forceinline__ void sum3(double& a, double& b, double& c) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            a += __shfl_xor_sync(0xffffffffu, a, o); b += __shfl_xor_sync(0xffffffffu, b, o); c += __shfl_xor_sync(0xffffffffu, c, o);
        }
    }
};

// point of a sweep that trajectory tr belongs to: first[pt] <= tr < first[pt + 1]
__device__ __forceinline__ int sweep_point2(const R2Params& P, long long tr) {
    int lo = 0, hi = P.n_points - 1;
    while (lo < hi) { const int mid = (lo + hi + 1) >> 1; if (P.first[mid] <= tr) lo = mid; else hi = mid - 1; }
    return lo;
}

// SWEEP: the trajectories belong to n_points parameter points (ca_rydberg2_sweep): model, initial state and tspan are per point (the
// model is staged in shared memory per warp), the warp's accumulator is flushed into the point's sums after every trajectory.
template <bool SWEEP>
__global__ void __launch_bounds__(kWarps2 * 32, 1) k_lindblad2(const __grid_constant__ R2Params P) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int lane = threadIdx.x & 31, wi = threadIdx.x >> 5;
    DevModel* sm_model = reinterpret_cast<DevModel*>(reinterpret_cast<double*>(smem_raw) + (size_t)kWarps2 * kWarpDoubles2) + wi;   // (SWEEP only)
    const DevModel* mp = &P.model;
    double* base = reinterpret_cast<double*>(smem_raw) + (size_t)wi * kWarpDoubles2;
    DevWarp W;
    W.lane = lane; W.xm = base; W.xp = base + kXN; W.kb = base + 2 * kXN + lane;
    W.coef = reinterpret_cast<Coef2*>(base + 2 * kXN + kKB * kS2 * 32);
    W.smp = base + 2 * kXN + kKB * kS2 * 32 + kStages2 * 16;
    W.anc = W.smp + 12 + lane;
#ifdef CA_TIMERS
    W.tl = clock64();
#endif
    Lane2 L;
    if (!SWEEP) L.init(lane, P.model);
    double rho0[kS2];
#pragma unroll
    for (int s = 0; s < kS2; ++s) rho0[s] = P.rho0m[lane * kS2 + s];
    double tr0 = P.tr0;
    const double* tspan = P.tspan;
    const long long wg = (long long)blockIdx.x * kWarps2 + wi, Wn = (long long)gridDim.x * kWarps2;
    long long tot_rhs = 0, tot_acc = 0, tot_rej = 0, tot_att = 0, tot_exact = 0;
    int worst = 0;
    double* acc_w = P.accum + (size_t)wg * P.nt * 2 * kAccSlots * 32 + lane;
    // Trajectories are handed out dynamically (step counts vary several-fold with the pair distance: V ~ R^-6 sets the step
    // size), first come first served: counters[7] is the next unassigned index.
    for (long long slot = wg;;) {
        if (slot >= P.n_traj) break;
        const long long tr = P.order ? (long long)P.order[slot] : slot;
        const uint64_t gi = (uint64_t)(P.traj_offset + tr);
        int pt = 0;
        if (SWEEP) {
            pt = sweep_point2(P, tr);
            __syncwarp();
            const int nw = (int)(sizeof(DevModel) / sizeof(double));
            const double* src = reinterpret_cast<const double*>(P.models + pt);
            double* dst = reinterpret_cast<double*>(sm_model);
            for (int q = lane; q < nw; q += 32) dst[q] = src[q];
            __syncwarp();
            mp = sm_model;
            L.init(lane, *mp);
#pragma unroll
            for (int s = 0; s < kS2; ++s) rho0[s] = P.rho0m[((size_t)pt * 32 + lane) * kS2 + s];
            tr0 = P.tr0s[pt];
            tspan = P.tspan + (size_t)pt * P.nt;
        }
        const DevModel& m = *mp;
        double* tab_w = (m.laser_noise && P.nf > 0) ? P.scratch + (size_t)wg * 4 * P.n_nodes_max : nullptr;
        CA_TICK(W, 3);
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
            for (int q = 0; q < 6; ++q) W.smp[6 * lane + q] = m.atom_motion ? s6[q] : 0.0;
        }
        if (tab_w) gen_noise_tables2(P, m, tr, gi, lane, tab_w);
        __syncwarp();
        __threadfence_block();
        CA_TICK(W, 0);
        Stats2 st;
        integrate2(W, L, m, P.ode, tspan, P.nt, tab_w, rho0, tr0, acc_w, st);
        tot_rhs += st.n_rhs; tot_acc += st.n_acc; tot_rej += st.n_rej; tot_exact += st.n_exact;
        if (st.status) worst = max(worst, -st.status);
        if (SWEEP) {   // flush this trajectory's contribution into the sums of its point
            double* dst = P.accum_pt + (size_t)pt * P.nt * 2 * kAccSlots * 32 + lane;
            for (int i = 0; i < P.nt * 2 * kAccSlots; ++i) {
                const double v = acc_w[(size_t)i * 32];
                if (v != 0.0) atomicAdd(dst + (size_t)i * 32, v);
                acc_w[(size_t)i * 32] = 0.0;
            }
        }
        unsigned long long nx = 0;
        if (lane == 0) nx = atomicAdd(P.counters + 7, 1ull);
        slot = Wn + (long long)__shfl_sync(0xffffffffu, nx, 0);
        __syncwarp();
    }
    tot_att = (long long)W.sum((double)tot_att);
    tot_exact = (long long)W.sum((double)tot_exact);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) worst = max(worst, __shfl_xor_sync(0xffffffffu, worst, o));
    if (lane == 0) {
        atomicAdd(P.counters + 0, (unsigned long long)tot_rhs);
        atomicAdd(P.counters + 1, (unsigned long long)tot_acc);
        atomicAdd(P.counters + 2, (unsigned long long)tot_rej);
        atomicAdd(P.counters + 3, (unsigned long long)tot_att);
        if (worst) atomicMax(P.counters + 4, (unsigned long long)worst);
        atomicAdd(P.counters + 5, (unsigned long long)tot_exact);   // (lane, step) beam evaluations that took the exact path
#ifdef CA_TIMERS
        if (wg == 0 || wg == 777) printf("timers (warp %lld, cycles): noise %lld coef %lld rhs7 %lld other %lld | rhs %lld steps %lld\n", wg, W.tm[0], W.tm[1], W.tm[2], W.tm[3], tot_rhs, tot_acc + tot_rej);
#endif
    }
}

// ---- longest-first hand-out --------------------------------------------------------------------------------------
// The step count of a trajectory scales with the blockade shift V = c6/R^6 of ITS atom pair (a pair 30 % closer needs 8x the
// steps), so a few trajectories are an order of magnitude longer than the mean.  Handing the expensive ones out first keeps
// the tail of the persistent grid short.  Cost key: 8 log2 of V at the closest approach of the sampled pair, 256 buckets,
// counting sort (order inside a bucket is arbitrary: the ensemble sums do not depend on it beyond rounding).
__device__ __forceinline__ int cost_bucket2(const R2Params& P, long long tr) {
    const DevModel& m = P.models ? P.models[sweep_point2(P, tr)] : P.model;
    double s[2][6];
#pragma unroll
    for (int a = 0; a < 2; ++a) {
        const double* src = a == 0 ? P.samples1 : P.samples2;
        if (src) { for (int q = 0; q < 6; ++q) s[a][q] = src[tr * 6 + q]; }
        else {
            SamplerConsts sc;
            for (int q = 0; q < 6; ++q) sc.sig[q] = m.sig[q];
            sc.U0 = m.U0; sc.w0 = m.w0; sc.z0 = m.z0; sc.mfac = m.mfac; sc.ecut = m.ecut;
            sample_atom_literal(sc, P.seed, (uint64_t)(P.traj_offset + tr), (uint32_t)a, s[a]);
        }
        for (int q = 0; q < 3; ++q) s[a][q] += m.center[a][q];
        if (!m.atom_motion) for (int q = 0; q < 6; ++q) s[a][q] = 0.0;
    }
    const double tend = P.models ? P.tspan[(size_t)sweep_point2(P, tr) * P.nt + P.nt - 1] : P.tspan[P.nt - 1];
    double r2min = 1e300;
#pragma unroll
    for (int k = 0; k < 3; ++k) {   // start, middle and end of the window
        const double t = 0.5 * k * tend;
        double d2 = 0.0;
        for (int q = 0; q < 3; ++q) {
            const double d = (s[0][q] - s[1][q]) + (m.free_motion ? (s[0][3 + q] - s[1][3 + q]) * t : 0.0);
            d2 += d * d;
        }
        r2min = fmin(r2min, d2);
    }
    double c2 = 0.0;
    for (int q = 0; q < 3; ++q) { const double d = m.center[0][q] - m.center[1][q]; c2 += d * d; }
    const double ratio = (1e-18 + c2 * c2 * c2) / (1e-18 + r2min * r2min * r2min);   // V / V_nominal
    const int b = 128 + (int)rint(8.0 * log2(fmax(ratio, 1e-30)));
    return b < 0 ? 0 : (b > 255 ? 255 : b);
}
__global__ void k_order2_hist(const __grid_constant__ R2Params P, unsigned char* __restrict__ bucket, unsigned* __restrict__ hist) {
    const long long tr = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (tr >= P.n_traj) return;
    const int b = cost_bucket2(P, tr);
    bucket[tr] = (unsigned char)b;
    atomicAdd(hist + b, 1u);
}
__global__ void k_order2_scan(unsigned* __restrict__ hist) {   // hist -> start offsets, most expensive bucket first
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    unsigned run = 0;
    for (int b = 255; b >= 0; --b) { const unsigned c = hist[b]; hist[b] = run; run += c; }
}
__global__ void k_order2_scatter(long long n, const unsigned char* __restrict__ bucket, unsigned* __restrict__ hist, unsigned* __restrict__ order) {
    const long long tr = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (tr >= n) return;
    order[atomicAdd(hist + bucket[tr], 1u)] = (unsigned)tr;
}

// K6 for the two-atom path: reduce the per-warp accumulators in a fixed order and combine every real-packed entry with its
// transposed partner into the full 25x25 column-major complex matrices, scaled.  rho/rho2 must be zero-filled.
__global__ void k_finalize2(const double* __restrict__ accum, int n_warps, int nt, double scale, double* __restrict__ rho, double* __restrict__ rho2) {
    const int e = blockIdx.x * blockDim.x + threadIdx.x;  // (j, which, slot, lane)
    const int per_t = 2 * kAccSlots * 32;
    if (e >= nt * per_t) return;
    const int j = e / per_t, r = e % per_t, which = r / (kAccSlots * 32), sl = r % (kAccSlots * 32), s = sl / 32, l = sl % 32;
    if (s == 9 && l != 0) return;
    int I, J, tl = 0, ts = 9;
    if (s < 9) slot_entry(l, s, &I, &J, &tl, &ts);
    const int rt = which * (kAccSlots * 32) + ts * 32 + tl;   // transposed entry (itself for slot 9)
    double sum = 0.0, sumt = 0.0;
    for (int wv = 0; wv < n_warps; ++wv) { const double* a = accum + ((size_t)wv * nt + j) * per_t; sum += a[r]; sumt += a[rt]; }
    double* out = (which ? rho2 : rho) + (size_t)j * 625 * 2;
    scatter_entry(l, s, sum * scale, sumt * scale, out);
}

}  // namespace ca

using namespace ca;

static constexpr size_t kSmem2 = sizeof(double) * (size_t)kWarps2 * kWarpDoubles2;

#define CA_CUDA2(expr)                                                                                      \
    do {                                                                                                    \
        cudaError_t _e = (expr);                                                                            \
        if (_e != cudaSuccess) { set_error(std::string(#expr) + ": " + cudaGetErrorString(_e)); return CA_ERR_CUDA; } \
    } while (0)


// The blocked kernel (k_lindblad2): simulation_czlp with psi0 inside the {0,1,r,p}^2 block.
static int run_r2_blocked(ca_ctx* c, const ca_model* model, const double* tspan, int32_t nt, const double* psi0, int64_t n_traj,
                          int64_t traj_offset, int64_t n_total, uint64_t seed, const double* samples1, const double* samples2,
                          const double* const* phases4, const double* f, const double* amp_red, const double* amp_blue, int32_t nf,
                          const ca_ode_opts* ode, int32_t out_flags, double* rho, double* rho2, ca_stats* stats) {
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
    double h_rho0[32 * kS2];
    if (!rho0_lanes(psi0, h_rho0, &P.tr0)) { set_error("internal: blocked two-atom kernel called with |l> components"); return CA_ERR_ARG; }
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
    if ((rc = up(20, h_rho0, sizeof h_rho0, &d))) return rc; P.rho0m = (const double*)d;
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
    CA_CUDA2(cudaStreamSynchronize(s));  // host staging buffers go out of scope after return
    const int sms = ctx_sm_count(c);
    int grid = (int)std::max<long long>(1, std::min<long long>((n_traj + kWarps2 - 1) / kWarps2, sms));
    const int n_warps = grid * kWarps2;
    P.n_nodes_max = P.model.n_nodes;
    if (noise) { P.scratch = (double*)ctx_buf(c, 29, sizeof(double) * (size_t)n_warps * 4 * P.model.n_nodes); if (!P.scratch) return CA_ERR_NOMEM; }
    const size_t acc_bytes = sizeof(double) * (size_t)n_warps * nt * 2 * kAccSlots * 32;
    P.accum = (double*)ctx_buf(c, 30, acc_bytes);
    P.counters = (unsigned long long*)ctx_buf(c, 31, 64);
    if (!P.accum || !P.counters) return CA_ERR_NOMEM;
    CA_CUDA2(cudaMemsetAsync(P.accum, 0, acc_bytes, s));
    CA_CUDA2(cudaMemsetAsync(P.counters, 0, 64, s));
    P.n_traj = n_traj; P.traj_offset = traj_offset; P.seed = seed; P.nt = nt;
    CA_CUDA2(cudaFuncSetAttribute(k_lindblad2<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmem2));
    CA_CUDA2(cudaEventRecord(e1, s));
    int launches = 0;
    if (n_traj > 2 * (long long)n_warps && n_traj < 0xffffffffLL && P.model.atom_motion) {   // more than two rounds: order longest-first
        unsigned char* d_bucket = (unsigned char*)ctx_buf(c, 33, (size_t)n_traj);
        unsigned* d_hist = (unsigned*)ctx_buf(c, 34, 256 * sizeof(unsigned));
        unsigned* d_order = (unsigned*)ctx_buf(c, 35, sizeof(unsigned) * (size_t)n_traj);
        if (!d_bucket || !d_hist || !d_order) return CA_ERR_NOMEM;
        CA_CUDA2(cudaMemsetAsync(d_hist, 0, 256 * sizeof(unsigned), s));
        const unsigned nb = (unsigned)((n_traj + 127) / 128);
        k_order2_hist<<<nb, 128, 0, s>>>(P, d_bucket, d_hist);
        k_order2_scan<<<1, 32, 0, s>>>(d_hist);
        k_order2_scatter<<<nb, 128, 0, s>>>(n_traj, d_bucket, d_hist, d_order);
        CA_CUDA2(cudaGetLastError());
        launches += 3;
        P.order = d_order;
    }
    if (n_traj > 0) {
        k_lindblad2<false><<<grid, kWarps2 * 32, kSmem2, s>>>(P);
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
        const int tot = nt * 2 * kAccSlots * 32;
        k_finalize2<<<(tot + 127) / 128, 128, 0, s>>>(P.accum, n_traj > 0 ? n_warps : 0, nt, div, d_rho, d_rho2);
        CA_CUDA2(cudaGetLastError());
        ++launches;
    }
    CA_CUDA2(cudaEventRecord(e3, s));
    ctx_set_last(c, P.counters, n_traj, launches, 2);
    if (out_flags & CA_OUT_DEVICE) {   // asynchronous: the device-side status is read later (ca_collect_stats / ca_stats_to_device)
        if (stats) { std::memset(stats, 0, sizeof *stats); stats->n_traj = n_traj; stats->n_launches = launches; }
        return CA_OK;
    }
    CA_CUDA2(cudaMemcpyAsync(rho, d_rho, sizeof(double) * out_doubles, cudaMemcpyDeviceToHost, s));
    CA_CUDA2(cudaMemcpyAsync(rho2, d_rho2, sizeof(double) * out_doubles, cudaMemcpyDeviceToHost, s));
    return collect_last(c, stats);   // synchronises; reports a trajectory that hit maxiters / went non-finite
}


/*
 * Batched two-atom sweep (get_cz_infidelity's seven ensembles, src/fidelity.jl:193-240, in ONE launch): point p has its own model,
 * ψ0 (inside the {0,1,r,p}² block), end time and trajectory count; trajectories of all points are handed out longest first.
 */
extern "C" int ca_rydberg2_sweep(ca_ctx* c, const ca_model* models, const double* psi0s, const double* t_end, int32_t n_points,
                                 const int64_t* n_traj_per_point, int64_t traj_offset, uint64_t seed, const double* f, const double* amp_red,
                                 const double* amp_blue, int32_t nf, const ca_ode_opts* ode, int32_t out_flags, double* rho, double* rho2,
                                 ca_stats* stats) {
    if (!c || !models || !psi0s || !t_end || n_points < 1 || !n_traj_per_point || !rho || !rho2) { set_error("bad argument"); return CA_ERR_ARG; }
    CA_CUDA2(cudaSetDevice(ctx_device(c)));
    cudaStream_t s = ctx_stream(c);
    constexpr int nt = 2;
    R2Params P;
    std::memset(&P, 0, sizeof P);
    std::vector<DevModel> dm(n_points);
    std::vector<double> h_rho0((size_t)n_points * 32 * kS2), h_tr0(n_points), h_ts((size_t)n_points * nt);
    std::vector<long long> first(n_points + 1, 0);
    bool noise = false;
    int n_nodes_max = 2;
    double tmax = 0.0;
    for (int p = 0; p < n_points; ++p) {
        if (models[p].two_atom_flags != 0) { set_error("ca_rydberg2_sweep: legacy two_atom_flags run through ca_rydberg2_run_ex"); return CA_ERR_UNSUPPORTED; }
        if (!(t_end[p] > 0) || n_traj_per_point[p] < 0) { set_error("bad t_end / trajectory count"); return CA_ERR_ARG; }
        int rc = derive_model(models[p], nullptr, 0, 0.0, t_end[p], &dm[p]);
        if (rc) return rc;
        if (!rho0_lanes(psi0s + 50 * (size_t)p, h_rho0.data() + (size_t)p * 32 * kS2, &h_tr0[p])) {
            set_error("ca_rydberg2_sweep: psi0 with |l> components runs through ca_rydberg2_run (full-form kernel)"); return CA_ERR_UNSUPPORTED;
        }
        h_ts[2 * p] = 0.0; h_ts[2 * p + 1] = t_end[p];
        first[p + 1] = first[p] + n_traj_per_point[p];
        noise = noise || dm[p].laser_noise;
        n_nodes_max = std::max(n_nodes_max, dm[p].n_nodes);
        tmax = std::max(tmax, t_end[p]);
    }
    const long long n_traj = first[n_points];
    if (noise && (!f || !amp_red || !amp_blue || nf < 1)) { set_error("laser_noise=true needs f and amplitudes"); return CA_ERR_ARG; }
    fill_ode(ode, 0.0, tmax, &P.ode);
    if (!(ode && ode->dtmax > 0)) P.ode.dtmax = 1e300;   // per-point horizons differ
    if (!P.ode.adaptive && !(P.ode.dt0 > 0)) { set_error("adaptive=false needs dt > 0"); return CA_ERR_ARG; }
    cudaEvent_t e0 = ctx_event(c, 0), e1 = ctx_event(c, 1), e2 = ctx_event(c, 2), e3 = ctx_event(c, 3);
    CA_CUDA2(cudaEventRecord(e0, s));
    auto up = [&](int id, const void* host, size_t bytes, const void** dev) -> int {
        void* p = ctx_buf(c, id, bytes);
        if (!p) return CA_ERR_NOMEM;
        cudaError_t e = cudaMemcpyAsync(p, host, bytes, cudaMemcpyHostToDevice, s);
        if (e != cudaSuccess) { set_error(std::string("cudaMemcpyAsync: ") + cudaGetErrorString(e)); return CA_ERR_CUDA; }
        *dev = p;
        return CA_OK;
    };
    const void* d;
    int rc;
    P.model = dm[0];
    if ((rc = up(50, dm.data(), sizeof(DevModel) * (size_t)n_points, &d))) return rc; P.models = (const DevModel*)d;
    if ((rc = up(20, h_rho0.data(), sizeof(double) * h_rho0.size(), &d))) return rc; P.rho0m = (const double*)d;
    if ((rc = up(51, h_tr0.data(), sizeof(double) * n_points, &d))) return rc; P.tr0s = (const double*)d;
    if ((rc = up(21, h_ts.data(), sizeof(double) * h_ts.size(), &d))) return rc; P.tspan = (const double*)d;
    if ((rc = up(52, first.data(), sizeof(long long) * (n_points + 1), &d))) return rc; P.first = (const long long*)d;
    std::vector<double> fa;
    if (noise) {
        P.nf = nf;
        fa.resize((size_t)3 * nf);
        std::memcpy(fa.data(), f, sizeof(double) * nf);
        std::memcpy(fa.data() + nf, amp_red, sizeof(double) * nf);
        std::memcpy(fa.data() + 2 * (size_t)nf, amp_blue, sizeof(double) * nf);
        if ((rc = up(24, fa.data(), sizeof(double) * 3 * (size_t)nf, &d))) return rc;
        P.f = (const double*)d; P.amp_red = P.f + nf; P.amp_blue = P.f + 2 * (size_t)nf;
    }
    CA_CUDA2(cudaStreamSynchronize(s));  // host staging buffers go out of scope after return
    const int sms = ctx_sm_count(c);
    const int grid = (int)std::max<long long>(1, std::min<long long>((n_traj + kWarps2 - 1) / kWarps2, sms));
    const int n_warps = grid * kWarps2;
    P.n_points = n_points; P.n_nodes_max = n_nodes_max;
    if (noise) { P.scratch = (double*)ctx_buf(c, 29, sizeof(double) * (size_t)n_warps * 4 * n_nodes_max); if (!P.scratch) return CA_ERR_NOMEM; }
    const size_t per_pt = (size_t)nt * 2 * kAccSlots * 32;
    P.accum = (double*)ctx_buf(c, 30, sizeof(double) * (size_t)n_warps * per_pt);
    P.accum_pt = (double*)ctx_buf(c, 53, sizeof(double) * (size_t)n_points * per_pt);
    P.counters = (unsigned long long*)ctx_buf(c, 31, 64);
    if (!P.accum || !P.accum_pt || !P.counters) return CA_ERR_NOMEM;
    CA_CUDA2(cudaMemsetAsync(P.accum, 0, sizeof(double) * (size_t)n_warps * per_pt, s));
    CA_CUDA2(cudaMemsetAsync(P.accum_pt, 0, sizeof(double) * (size_t)n_points * per_pt, s));
    CA_CUDA2(cudaMemsetAsync(P.counters, 0, 64, s));
    P.n_traj = n_traj; P.traj_offset = traj_offset; P.seed = seed; P.nt = nt;
    const size_t smem = kSmem2 + sizeof(DevModel) * kWarps2;
    CA_CUDA2(cudaFuncSetAttribute(k_lindblad2<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    CA_CUDA2(cudaEventRecord(e1, s));
    int launches = 0;
    if (n_traj > 2 * (long long)n_warps && n_traj < 0xffffffffLL) {   // more than two rounds: order longest-first across all points
        unsigned char* d_bucket = (unsigned char*)ctx_buf(c, 33, (size_t)n_traj);
        unsigned* d_hist = (unsigned*)ctx_buf(c, 34, 256 * sizeof(unsigned));
        unsigned* d_order = (unsigned*)ctx_buf(c, 35, sizeof(unsigned) * (size_t)n_traj);
        if (!d_bucket || !d_hist || !d_order) return CA_ERR_NOMEM;
        CA_CUDA2(cudaMemsetAsync(d_hist, 0, 256 * sizeof(unsigned), s));
        const unsigned nb = (unsigned)((n_traj + 127) / 128);
        k_order2_hist<<<nb, 128, 0, s>>>(P, d_bucket, d_hist);
        k_order2_scan<<<1, 32, 0, s>>>(d_hist);
        k_order2_scatter<<<nb, 128, 0, s>>>(n_traj, d_bucket, d_hist, d_order);
        CA_CUDA2(cudaGetLastError());
        launches += 3;
        P.order = d_order;
    }
    if (n_traj > 0) {
        k_lindblad2<true><<<grid, kWarps2 * 32, smem, s>>>(P);
        CA_CUDA2(cudaGetLastError());
        ++launches;
    }
    CA_CUDA2(cudaEventRecord(e2, s));
    // finalize: the last time point of every point, each divided by its own trajectory count (CA_OUT_MEAN)
    const size_t out_doubles = (size_t)n_points * 625 * 2;
    double *d_rho, *d_rho2;
    if (out_flags & CA_OUT_DEVICE) { d_rho = rho; d_rho2 = rho2; }
    else { d_rho = (double*)ctx_buf(c, 32, sizeof(double) * 2 * out_doubles); if (!d_rho) return CA_ERR_NOMEM; d_rho2 = d_rho + out_doubles; }
    CA_CUDA2(cudaMemsetAsync(d_rho, 0, sizeof(double) * out_doubles, s));
    CA_CUDA2(cudaMemsetAsync(d_rho2, 0, sizeof(double) * out_doubles, s));
    for (int p = 0; p < n_points; ++p) {
        const double div = (out_flags & CA_OUT_SUMS) ? 1.0 : 1.0 / (double)std::max<long long>(1, n_traj_per_point[p]);
        const int tot = 2 * kAccSlots * 32;
        k_finalize2<<<(tot + 127) / 128, 128, 0, s>>>(P.accum_pt + (size_t)p * per_pt + (size_t)(nt - 1) * 2 * kAccSlots * 32, 1, 1, div,
                                                     d_rho + (size_t)p * 1250, d_rho2 + (size_t)p * 1250);
        ++launches;
    }
    CA_CUDA2(cudaGetLastError());
    CA_CUDA2(cudaEventRecord(e3, s));
    ctx_set_last(c, P.counters, n_traj, launches, 2);
    if (out_flags & CA_OUT_DEVICE) {
        if (stats) { std::memset(stats, 0, sizeof *stats); stats->n_traj = n_traj; stats->n_launches = launches; }
        return CA_OK;
    }
    CA_CUDA2(cudaMemcpyAsync(rho, d_rho, sizeof(double) * out_doubles, cudaMemcpyDeviceToHost, s));
    CA_CUDA2(cudaMemcpyAsync(rho2, d_rho2, sizeof(double) * out_doubles, cudaMemcpyDeviceToHost, s));
    return collect_last(c, stats);
}

extern "C" int ca_rydberg2_run_ex(ca_ctx* c, const ca_model* model, const double* tspan, int32_t nt, int32_t nt_seg1, const double* psi0,
                                  const double* rho0, int64_t n_traj, int64_t traj_offset, int64_t n_total, uint64_t seed,
                                  const double* samples1, const double* samples2, const double* const* phases4, const double* f,
                                  const double* amp_red, const double* amp_blue, int32_t nf, const ca_ode_opts* ode, int32_t out_flags,
                                  double* rho, double* rho2, ca_stats* stats) {
    if (!c || !model || !tspan || nt < 1 || (!psi0 && !rho0) || n_traj < 0 || !rho || !rho2 || nt_seg1 < 0 || nt_seg1 >= nt + (nt_seg1 == 0)) {
        set_error("bad argument"); return CA_ERR_ARG;
    }
    for (int j = 1; j < nt; ++j)
        if (j != nt_seg1 && !(tspan[j] > tspan[j - 1])) { set_error("tspan must be strictly increasing (within each time segment)"); return CA_ERR_ARG; }
    bool blocked = !rho0 && nt_seg1 == 0 && model->two_atom_flags == 0;
    if (blocked)
        for (int i = 0; i < 25; ++i)
            if ((i % 5 == 4 || i / 5 == 4) && (psi0[2 * i] != 0.0 || psi0[2 * i + 1] != 0.0)) blocked = false;   // |l> components: SURVEY App. A6 fall-back
    if (blocked)
        return run_r2_blocked(c, model, tspan, nt, psi0, n_traj, traj_offset, n_total, seed, samples1, samples2, phases4, f, amp_red, amp_blue, nf,
                              ode, out_flags, rho, rho2, stats);
    return run_r2_full(c, model, tspan, nt, nt_seg1, psi0, rho0, n_traj, traj_offset, n_total, seed, samples1, samples2, phases4, f, amp_red,
                       amp_blue, nf, ode, out_flags, rho, rho2, stats);
}

extern "C" int ca_rydberg2_run(ca_ctx* c, const ca_model* model, const double* tspan, int32_t nt, const double* psi0, int64_t n_traj,
                               int64_t traj_offset, int64_t n_total, uint64_t seed, const double* samples1, const double* samples2,
                               const double* const* phases4, const double* f, const double* amp_red, const double* amp_blue, int32_t nf,
                               const ca_ode_opts* ode, int32_t out_flags, double* rho, double* rho2, ca_stats* stats) {
    return ca_rydberg2_run_ex(c, model, tspan, nt, 0, psi0, nullptr, n_traj, traj_offset, n_total, seed, samples1, samples2, phases4, f, amp_red,
                              amp_blue, nf, ode, out_flags, rho, rho2, stats);
}
