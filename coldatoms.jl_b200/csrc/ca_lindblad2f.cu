// ca_lindblad2f.cu -- the full-form two-atom kernel (per-lane code: ca_lindblad2f.cuh), its finalize kernel and host driver.
// Own translation unit: the 25-row straight-line RHS is the slowest thing to compile in the library.
#include <cuda_runtime.h>

#include <algorithm>
#include <cstring>
#include <string>
#include <vector>

#include "ca_r2params.cuh"

namespace ca {

// ---- full-form kernel (ca_lindblad2f.cuh): any initial state, legacy parameterisations, two time segments ---------------------
constexpr int kWarpsF = 3;   // 62.7 KB of shared memory per warp
struct DevWarpF {
    int lane;
    double *xm, *xp, *kb;
    double* coef;
    double* smp;
    __device__ __forceinline__ void sync() { __syncwarp(); }
    __device__ __forceinline__ double sum(double v) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
        return v;
    }
};

__global__ void __launch_bounds__(kWarpsF * 32, 1) k_lindblad2_full(const __grid_constant__ R2Params P) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const DevModel& m = P.model;
    const int lane = threadIdx.x & 31, wi = threadIdx.x >> 5;
    double* base = reinterpret_cast<double*>(smem_raw) + (size_t)wi * kWarpDoublesF;
    DevWarpF W;
    W.lane = lane; W.xm = base; W.xp = base + kFXN; W.kb = base + 2 * kFXN + lane;
    W.coef = base + 2 * kFXN + kFB * kF * 32;
    W.smp = W.coef + kStages2 * kCoefF;
    const long long wg = (long long)blockIdx.x * kWarpsF + wi, Wn = (long long)gridDim.x * kWarpsF;
    long long tot_rhs = 0, tot_acc = 0, tot_rej = 0, tot_att = 0;
    int worst = 0;
    double* acc_w = P.accum + (size_t)wg * P.nt * 2 * kF * 32 + lane;
    double* tab_w = (m.laser_noise && P.nf > 0) ? P.scratch + (size_t)wg * 4 * m.n_nodes : nullptr;
    const bool shift_after = (P.x.flags & CA_2A_SHIFT_AFTER_MOTION) != 0;
    for (long long slot = wg;;) {
        if (slot >= P.n_traj) break;
        const long long tr = slot;
        const uint64_t gi = (uint64_t)(P.traj_offset + tr);
        if (lane < 2) {   // samples (cz_model.jl:111-125; two_atom_model.jl:128-135)
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
            if (!shift_after) for (int q = 0; q < 3; ++q) s6[q] += m.center[lane][q];
            for (int q = 0; q < 6; ++q) W.smp[6 * lane + q] = m.atom_motion ? s6[q] : 0.0;
        }
        if (tab_w) gen_noise_tables2(P, m, tr, gi, lane, tab_w);
#pragma unroll
        for (int s = 0; s < kF; ++s) W.kb[(FB_Y * kF + s) * 32] = P.rho0f[lane * kF + s];
        __syncwarp();
        __threadfence_block();
        Stats2 st;
        st.n_rhs = st.n_acc = st.n_rej = st.n_exact = 0; st.status = 0;
        if (P.nt_seg1 > 0 && P.nt_seg1 < P.nt) {   // two consecutive solves (two_atom_model.jl:304-305)
            integrate2f(W, m, P.x, P.ode, P.tspan, P.nt_seg1, 0, tab_w, acc_w, st);
            if (st.status == 0) integrate2f(W, m, P.x, P.ode, P.tspan + P.nt_seg1, P.nt - P.nt_seg1, 1, tab_w, acc_w + (size_t)P.nt_seg1 * 2 * kF * 32, st);
        } else integrate2f(W, m, P.x, P.ode, P.tspan, P.nt, 0, tab_w, acc_w, st);
        tot_rhs += st.n_rhs; tot_acc += st.n_acc; tot_rej += st.n_rej;
        if (st.status) worst = max(worst, -st.status);
        unsigned long long nx = 0;
        if (lane == 0) nx = atomicAdd(P.counters + 7, 1ull);
        slot = Wn + (long long)__shfl_sync(0xffffffffu, nx, 0);
        __syncwarp();
    }
    tot_att = (long long)W.sum((double)tot_att);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) worst = max(worst, __shfl_xor_sync(0xffffffffu, worst, o));
    if (lane == 0) {
        atomicAdd(P.counters + 0, (unsigned long long)tot_rhs);
        atomicAdd(P.counters + 1, (unsigned long long)tot_acc);
        atomicAdd(P.counters + 2, (unsigned long long)tot_rej);
        atomicAdd(P.counters + 3, (unsigned long long)tot_att);
        if (worst) atomicMax(P.counters + 4, (unsigned long long)worst);
    }
}

// K6 for the full-form path: per-warp accumulators [warp][nt][2][kF][32] -> 25x25 column-major complex matrices:
// ρ_IJ = ((M_IJ + M_JI) + i (M_IJ - M_JI)) / 2 with M_IJ = row I of lane J.
__global__ void k_finalize2f(const double* __restrict__ accum, int n_warps, int nt, double scale, double* __restrict__ rho, double* __restrict__ rho2) {
    const int e = blockIdx.x * blockDim.x + threadIdx.x;   // (j, which, I, J)
    if (e >= nt * 2 * 625) return;
    const int j = e / 1250, r = e % 1250, which = r / 625, IJ = r % 625, I = IJ % 25, J = IJ / 25;
    double s = 0.0, st = 0.0;
    for (int wv = 0; wv < n_warps; ++wv) {
        const double* a = accum + (((size_t)wv * nt + j) * 2 + which) * kF * 32;
        s += a[I * 32 + J]; st += a[J * 32 + I];
    }
    double* out = (which ? rho2 : rho) + (size_t)j * 1250 + 2 * (I + 25 * J);
    out[0] = 0.5 * (s + st) * scale;
    out[1] = I == J ? 0.0 : 0.5 * (s - st) * scale;
}


#define CA_CUDA2(expr)                                                                                      \
    do {                                                                                                    \
        cudaError_t _e = (expr);                                                                            \
        if (_e != cudaSuccess) { set_error(std::string(#expr) + ": " + cudaGetErrorString(_e)); return CA_ERR_CUDA; } \
    } while (0)

static constexpr size_t kSmemF = sizeof(double) * (size_t)kWarpsF * kWarpDoublesF;

// The full-form kernel (k_lindblad2_full): any initial state, legacy parameterisations, two time segments.
int run_r2_full(ca_ctx* c, const ca_model* model, const double* tspan, int32_t nt, int32_t nt_seg1, const double* psi0, const double* rho0,
                       int64_t n_traj, int64_t traj_offset, int64_t n_total, uint64_t seed, const double* samples1, const double* samples2,
                       const double* const* phases4, const double* f, const double* amp_red, const double* amp_blue, int32_t nf,
                       const ca_ode_opts* ode, int32_t out_flags, double* rho, double* rho2, ca_stats* stats) {
    CA_CUDA2(cudaSetDevice(ctx_device(c)));
    cudaStream_t s = ctx_stream(c);
    R2Params P;
    std::memset(&P, 0, sizeof P);
    double tmax = tspan[nt - 1];
    if (nt_seg1 > 0) tmax = std::max(tmax, tspan[nt_seg1 - 1]);   // the noise grid / phase-jump time refer to the longer time axis
    int rc = derive_model(*model, nullptr, 0, tspan[0], tmax, &P.model);
    if (rc) return rc;
    if ((rc = derive_model2x(*model, &P.x))) return rc;
    const bool noise = P.model.laser_noise != 0;
    if (noise) {
        if (!f || !amp_red || !amp_blue || nf < 1) { set_error("laser_noise=true needs f and amplitudes"); return CA_ERR_ARG; }
        for (int j = 0; j < nt; ++j) if (tspan[j] < 0) { set_error("laser noise is tabulated on [0, max(tspan)]: tspan must be >= 0"); return CA_ERR_ARG; }
    }
    const int n1 = nt_seg1 > 0 ? nt_seg1 : nt;
    fill_ode(ode, tspan[0], tspan[n1 - 1], &P.ode);
    if (nt_seg1 > 0) P.ode.dtmax = (ode && ode->dtmax > 0) ? ode->dtmax : std::max(tspan[n1 - 1] - tspan[0], tspan[nt - 1] - tspan[nt_seg1]);
    if (!P.ode.adaptive && !(P.ode.dt0 > 0)) { set_error("adaptive=false needs dt > 0"); return CA_ERR_ARG; }
    std::vector<double> h_rho0(32 * kF);
    rho0_columns(psi0, rho0, h_rho0.data());
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
    if ((rc = up(40, h_rho0.data(), sizeof(double) * 32 * kF, &d))) return rc; P.rho0f = (const double*)d;
    if ((rc = up(21, tspan, sizeof(double) * nt, &d))) return rc; P.tspan = (const double*)d;
    if (samples1 && n_traj > 0) { if ((rc = up(22, samples1, sizeof(double) * 6 * (size_t)n_traj, &d))) return rc; P.samples1 = (const double*)d; }
    if (samples2 && n_traj > 0) { if ((rc = up(23, samples2, sizeof(double) * 6 * (size_t)n_traj, &d))) return rc; P.samples2 = (const double*)d; }
    std::vector<double> fa;
    if (noise) {
        P.nf = nf;
        fa.resize((size_t)3 * nf);
        std::memcpy(fa.data(), f, sizeof(double) * nf);
        std::memcpy(fa.data() + nf, amp_red, sizeof(double) * nf);
        std::memcpy(fa.data() + 2 * (size_t)nf, amp_blue, sizeof(double) * nf);
        if ((rc = up(24, fa.data(), sizeof(double) * 3 * (size_t)nf, &d))) return rc;
        P.f = (const double*)d; P.amp_red = P.f + nf; P.amp_blue = P.f + 2 * (size_t)nf;
        if (phases4 && n_traj > 0)
            for (int q = 0; q < 4; ++q)
                if (phases4[q]) { if ((rc = up(25 + q, phases4[q], sizeof(double) * (size_t)nf * n_traj, &d))) return rc; P.phases[q] = (const double*)d; }
    }
    CA_CUDA2(cudaStreamSynchronize(s));  // host staging buffers go out of scope after return
    const int sms = ctx_sm_count(c);
    const int grid = (int)std::max<long long>(1, std::min<long long>((n_traj + kWarpsF - 1) / kWarpsF, sms));
    const int n_warps = grid * kWarpsF;
    if (noise) { P.scratch = (double*)ctx_buf(c, 29, sizeof(double) * (size_t)n_warps * 4 * P.model.n_nodes); if (!P.scratch) return CA_ERR_NOMEM; }
    const size_t acc_bytes = sizeof(double) * (size_t)n_warps * nt * 2 * kF * 32;
    P.accum = (double*)ctx_buf(c, 41, acc_bytes);
    P.counters = (unsigned long long*)ctx_buf(c, 31, 64);
    if (!P.accum || !P.counters) return CA_ERR_NOMEM;
    CA_CUDA2(cudaMemsetAsync(P.accum, 0, acc_bytes, s));
    CA_CUDA2(cudaMemsetAsync(P.counters, 0, 64, s));
    P.n_traj = n_traj; P.traj_offset = traj_offset; P.seed = seed; P.nt = nt; P.nt_seg1 = nt_seg1;
    CA_CUDA2(cudaFuncSetAttribute(k_lindblad2_full, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmemF));
    CA_CUDA2(cudaEventRecord(e1, s));
    int launches = 0;
    if (n_traj > 0) {
        k_lindblad2_full<<<grid, kWarpsF * 32, kSmemF, s>>>(P);
        CA_CUDA2(cudaGetLastError());
        ++launches;
    }
    CA_CUDA2(cudaEventRecord(e2, s));
    const double div = (out_flags & CA_OUT_SUMS) ? 1.0 : 1.0 / (double)(n_total > 0 ? n_total : std::max<long long>(1, n_traj));
    const size_t out_doubles = (size_t)nt * 625 * 2;
    double *d_rho, *d_rho2;
    if (out_flags & CA_OUT_DEVICE) { d_rho = rho; d_rho2 = rho2; }
    else { d_rho = (double*)ctx_buf(c, 32, sizeof(double) * 2 * out_doubles); if (!d_rho) return CA_ERR_NOMEM; d_rho2 = d_rho + out_doubles; }
    {
        const int tot = nt * 2 * 625;
        k_finalize2f<<<(tot + 127) / 128, 128, 0, s>>>(P.accum, n_traj > 0 ? n_warps : 0, nt, div, d_rho, d_rho2);
        CA_CUDA2(cudaGetLastError());
        ++launches;
    }
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


}  // namespace ca
