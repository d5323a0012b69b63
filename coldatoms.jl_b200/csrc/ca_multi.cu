// ca_multi.cu -- several GPUs driven from ONE process through the C ABI: one ca_ctx per device, trajectories sharded by
// contiguous ranges of the GLOBAL index, one ncclAllReduce of [Σρ | Σρρ | counters] per ensemble call.
//
// NCCL is loaded with dlopen("libnccl.so.2") on the first multi-device init: a Julia caller picks it up from the system (or
// LD_LIBRARY_PATH), a Python process that already imported torch gets torch's copy; single-GPU users never load it.
#include <cuda_runtime.h>
#include <dlfcn.h>

#include <algorithm>
#include <cstring>
#include <string>
#include <vector>

#include "ca_host.h"

namespace ca {
cudaStream_t ctx_stream(ca_ctx* c);
int ctx_device(ca_ctx* c);
void* ctx_buf(ca_ctx* c, int which, size_t bytes);

// the handful of NCCL entry points used, with NCCL's own C signatures (nccl.h: ncclComm_t is an opaque pointer, ncclResult_t /
// ncclDataType_t / ncclRedOp_t are int-sized enums; ncclSuccess = 0, ncclFloat64 = 8, ncclSum = 0)
struct Nccl {
    void* h = nullptr;
    int (*CommInitAll)(void** comms, int ndev, const int* devlist) = nullptr;
    int (*CommDestroy)(void* comm) = nullptr;
    int (*AllReduce)(const void* send, void* recv, size_t count, int dtype, int op, void* comm, cudaStream_t s) = nullptr;
    int (*GroupStart)() = nullptr;
    int (*GroupEnd)() = nullptr;
    const char* (*GetErrorString)(int) = nullptr;
};
static Nccl g_nccl;
static int load_nccl() {
    if (g_nccl.h) return CA_OK;
    void* h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_LOCAL);
    if (!h) h = dlopen("libnccl.so", RTLD_NOW | RTLD_LOCAL);
    if (!h) { set_error(std::string("multi-GPU needs NCCL: ") + dlerror()); return CA_ERR_NCCL; }
    Nccl n;
    n.h = h;
    n.CommInitAll = (decltype(n.CommInitAll))dlsym(h, "ncclCommInitAll");
    n.CommDestroy = (decltype(n.CommDestroy))dlsym(h, "ncclCommDestroy");
    n.AllReduce = (decltype(n.AllReduce))dlsym(h, "ncclAllReduce");
    n.GroupStart = (decltype(n.GroupStart))dlsym(h, "ncclGroupStart");
    n.GroupEnd = (decltype(n.GroupEnd))dlsym(h, "ncclGroupEnd");
    n.GetErrorString = (decltype(n.GetErrorString))dlsym(h, "ncclGetErrorString");
    if (!n.CommInitAll || !n.CommDestroy || !n.AllReduce || !n.GroupStart || !n.GroupEnd || !n.GetErrorString) {
        set_error("libnccl.so.2 lacks an expected symbol"); dlclose(h); return CA_ERR_NCCL;
    }
    g_nccl = n;
    return CA_OK;
}
}  // namespace ca

using namespace ca;

struct ca_multi {
    std::vector<ca_ctx*> ctx;
    std::vector<void*> comm;   // ncclComm_t per device (empty for one device)
};

#define CA_NCCL(expr)                                                                                              \
    do {                                                                                                           \
        int _r = (expr);                                                                                           \
        if (_r != 0) { set_error(std::string(#expr) + ": " + g_nccl.GetErrorString(_r)); return CA_ERR_NCCL; }     \
    } while (0)
#define CA_CUDAM(expr)                                                                                             \
    do {                                                                                                           \
        cudaError_t _e = (expr);                                                                                   \
        if (_e != cudaSuccess) { set_error(std::string(#expr) + ": " + cudaGetErrorString(_e)); return CA_ERR_CUDA; } \
    } while (0)

namespace {
constexpr int kBufMulti = 60;   // ctx_buf id of the per-device [Σρ | Σρρ | 8 counters] buffer

// shard g of n items over G devices: [lo, lo + cnt)
inline void shard(long long n, int g, int G, long long* lo, long long* cnt) {
    const long long a = n * g / G, b = n * (g + 1) / G;
    *lo = a; *cnt = b - a;
}

// All-reduce every device's buffer in place (one grouped call), bring device 0's copy to the host, check the status counters.
int reduce_and_fetch(ca_multi* m, size_t n_out, std::vector<double>& host, ca_stats* stats, int launches) {
    const int G = (int)m->ctx.size();
    if (G > 1) {
        CA_NCCL(g_nccl.GroupStart());
        for (int g = 0; g < G; ++g) {
            double* buf = (double*)ctx_buf(m->ctx[g], kBufMulti, 0);
            CA_NCCL(g_nccl.AllReduce(buf, buf, n_out + 8, /*ncclFloat64*/ 8, /*ncclSum*/ 0, m->comm[g], ctx_stream(m->ctx[g])));
        }
        CA_NCCL(g_nccl.GroupEnd());
    }
    host.resize(n_out + 8);
    CA_CUDAM(cudaSetDevice(ctx_device(m->ctx[0])));
    CA_CUDAM(cudaMemcpyAsync(host.data(), ctx_buf(m->ctx[0], kBufMulti, 0), sizeof(double) * (n_out + 8), cudaMemcpyDeviceToHost, ctx_stream(m->ctx[0])));
    for (int g = 0; g < G; ++g) {
        CA_CUDAM(cudaSetDevice(ctx_device(m->ctx[g])));
        CA_CUDAM(cudaStreamSynchronize(ctx_stream(m->ctx[g])));
    }
    const double* cnt = host.data() + n_out;
    const int status = cnt[5] > 0 ? CA_ERR_NONFINITE : (cnt[4] > 0 ? CA_ERR_MAXITERS : CA_OK);
    if (stats) {
        std::memset(stats, 0, sizeof *stats);
        stats->n_rhs = (int64_t)cnt[0]; stats->n_accept = (int64_t)cnt[1]; stats->n_reject = (int64_t)cnt[2];
        stats->n_sample_attempts = (int64_t)cnt[3]; stats->n_exact_steps = (int64_t)cnt[6]; stats->n_traj = (int64_t)cnt[7];
        stats->n_launches = launches; stats->status = status;
        double kmax = 0.0, tmax = 0.0;   // device times: the slowest device (they run concurrently)
        for (int g = 0; g < G; ++g) {
            ca_stats s1;
            ca_collect_stats(m->ctx[g], &s1);
            kmax = std::max(kmax, s1.kernel_ms); tmax = std::max(tmax, s1.total_ms);
        }
        stats->kernel_ms = kmax; stats->total_ms = tmax;
    }
    if (status == CA_ERR_MAXITERS) set_error("a trajectory hit maxiters on some device (it is left out of the sums)");
    else if (status) set_error("a trajectory produced a non-finite state or the step size underflowed on some device");
    return status;
}
}  // namespace

extern "C" {

int ca_init_multi(const int* devices, int32_t ndev, ca_multi** out) {
    if (!out) { set_error("null out pointer"); return CA_ERR_ARG; }
    *out = nullptr;
    int n_vis = 0;
    cudaError_t e = cudaGetDeviceCount(&n_vis);
    if (e != cudaSuccess || n_vis == 0) { set_error(std::string("no CUDA device: ") + cudaGetErrorString(e)); return CA_ERR_NODEVICE; }
    std::vector<int> devs;
    if (ndev <= 0) { for (int i = 0; i < n_vis; ++i) devs.push_back(i); }
    else for (int i = 0; i < ndev; ++i) devs.push_back(devices ? devices[i] : i);
    for (size_t i = 0; i < devs.size(); ++i)
        for (size_t j = 0; j < i; ++j)
            if (devs[i] == devs[j]) { set_error("duplicate device in the list"); return CA_ERR_ARG; }
    ca_multi* m = new ca_multi();
    for (int d : devs) {
        ca_ctx* c = nullptr;
        int rc = ca_init(d, &c);
        if (rc) { ca_finalize_multi(m); return rc; }
        m->ctx.push_back(c);
    }
    if (devs.size() > 1) {
        int rc = load_nccl();
        if (rc) { ca_finalize_multi(m); return rc; }
        m->comm.assign(devs.size(), nullptr);
        int r = g_nccl.CommInitAll(m->comm.data(), (int)devs.size(), devs.data());
        if (r != 0) { set_error(std::string("ncclCommInitAll: ") + g_nccl.GetErrorString(r)); m->comm.clear(); ca_finalize_multi(m); return CA_ERR_NCCL; }
    }
    *out = m;
    return CA_OK;
}

int ca_finalize_multi(ca_multi* m) {
    if (!m) return CA_OK;
    for (void* c : m->comm) if (c) g_nccl.CommDestroy(c);
    for (ca_ctx* c : m->ctx) ca_finalize(c);
    delete m;
    return CA_OK;
}

int ca_multi_size(const ca_multi* m) { return m ? (int)m->ctx.size() : 0; }
ca_ctx* ca_multi_context(ca_multi* m, int32_t i) { return (m && i >= 0 && i < (int)m->ctx.size()) ? m->ctx[i] : nullptr; }

int ca_rydberg1_run_multi(ca_multi* m, const ca_model* model, const double* tspan, int32_t nt, const double* psi0, int64_t n_traj,
                          int64_t traj_offset, int64_t n_total, uint64_t seed, const double* samples, const double* phases_red,
                          const double* phases_blue, const double* f, const double* amp_red, const double* amp_blue, int32_t nf,
                          const ca_ode_opts* ode, int32_t out_flags, double* rho, double* rho2, ca_stats* stats) {
    if (!m || m->ctx.empty() || nt < 1 || n_traj < 0 || !rho || !rho2) { set_error("bad argument"); return CA_ERR_ARG; }
    if (out_flags & CA_OUT_DEVICE) { set_error("the multi-GPU calls return host arrays"); return CA_ERR_ARG; }
    const int G = (int)m->ctx.size();
    const size_t half = (size_t)nt * 50, n_out = 2 * half;
    int launches = 0;
    for (int g = 0; g < G; ++g) {
        long long lo, cnt;
        shard(n_traj, g, G, &lo, &cnt);
        CA_CUDAM(cudaSetDevice(ctx_device(m->ctx[g])));
        double* buf = (double*)ctx_buf(m->ctx[g], kBufMulti, sizeof(double) * (n_out + 8));
        if (!buf) return CA_ERR_NOMEM;
        ca_stats s1;
        int rc = ca_rydberg1_run(m->ctx[g], model, tspan, nt, psi0, cnt, traj_offset + lo, n_total, seed, samples ? samples + 6 * lo : nullptr,
                                 phases_red ? phases_red + (size_t)nf * lo : nullptr, phases_blue ? phases_blue + (size_t)nf * lo : nullptr, f,
                                 amp_red, amp_blue, nf, ode, CA_OUT_SUMS | CA_OUT_DEVICE, buf, buf + half, &s1);
        if (rc) return rc;
        launches += s1.n_launches;
        if ((rc = ca_stats_to_device(m->ctx[g], buf + n_out))) return rc;
    }
    std::vector<double> host;
    const int status = reduce_and_fetch(m, n_out, host, stats, launches);
    if (status == CA_ERR_CUDA || status == CA_ERR_NCCL) return status;
    const double div = (out_flags & CA_OUT_SUMS) ? 1.0 : 1.0 / (double)(n_total > 0 ? n_total : std::max<long long>(1, n_traj));
    for (size_t i = 0; i < half; ++i) { rho[i] = host[i] * div; rho2[i] = host[half + i] * div; }
    return status;
}

int ca_rydberg2_run_multi(ca_multi* m, const ca_model* model, const double* tspan, int32_t nt, const double* psi0, int64_t n_traj,
                          int64_t traj_offset, int64_t n_total, uint64_t seed, const double* samples1, const double* samples2,
                          const double* const* phases4, const double* f, const double* amp_red, const double* amp_blue, int32_t nf,
                          const ca_ode_opts* ode, int32_t out_flags, double* rho, double* rho2, ca_stats* stats) {
    if (!m || m->ctx.empty() || nt < 1 || n_traj < 0 || !rho || !rho2) { set_error("bad argument"); return CA_ERR_ARG; }
    if (out_flags & CA_OUT_DEVICE) { set_error("the multi-GPU calls return host arrays"); return CA_ERR_ARG; }
    const int G = (int)m->ctx.size();
    const size_t half = (size_t)nt * 1250, n_out = 2 * half;
    int launches = 0;
    for (int g = 0; g < G; ++g) {
        long long lo, cnt;
        shard(n_traj, g, G, &lo, &cnt);
        CA_CUDAM(cudaSetDevice(ctx_device(m->ctx[g])));
        double* buf = (double*)ctx_buf(m->ctx[g], kBufMulti, sizeof(double) * (n_out + 8));
        if (!buf) return CA_ERR_NOMEM;
        const double* ph[4] = {nullptr, nullptr, nullptr, nullptr};
        if (phases4) for (int q = 0; q < 4; ++q) ph[q] = phases4[q] ? phases4[q] + (size_t)nf * lo : nullptr;
        ca_stats s1;
        int rc = ca_rydberg2_run(m->ctx[g], model, tspan, nt, psi0, cnt, traj_offset + lo, n_total, seed, samples1 ? samples1 + 6 * lo : nullptr,
                                 samples2 ? samples2 + 6 * lo : nullptr, phases4 ? ph : nullptr, f, amp_red, amp_blue, nf, ode,
                                 CA_OUT_SUMS | CA_OUT_DEVICE, buf, buf + half, &s1);
        if (rc) return rc;
        launches += s1.n_launches;
        if ((rc = ca_stats_to_device(m->ctx[g], buf + n_out))) return rc;
    }
    std::vector<double> host;
    const int status = reduce_and_fetch(m, n_out, host, stats, launches);
    if (status == CA_ERR_CUDA || status == CA_ERR_NCCL) return status;
    const double div = (out_flags & CA_OUT_SUMS) ? 1.0 : 1.0 / (double)(n_total > 0 ? n_total : std::max<long long>(1, n_traj));
    for (size_t i = 0; i < half; ++i) { rho[i] = host[i] * div; rho2[i] = host[half + i] * div; }
    return status;
}

// Sweeps shard by GRID POINT (every point keeps all its trajectories on one device: no data-path collective; BASELINE config 5).
// Points are dealt round-robin so that every device sees every region of the grid (pulse times set the cost of a point).
int ca_rydberg1_sweep_multi(ca_multi* m, const ca_model* models, const double* psi0s, const double* t_end, int32_t n_points,
                            int64_t n_traj_per_point, int64_t traj_offset, int64_t n_total, uint64_t seed, const double* f,
                            const double* amp_red, const double* amp_blue, int32_t nf, const ca_ode_opts* ode, int32_t out_flags, double* rho,
                            double* rho2, ca_stats* stats) {
    if (!m || m->ctx.empty() || !models || !psi0s || !t_end || n_points < 1 || !rho || !rho2) { set_error("bad argument"); return CA_ERR_ARG; }
    if (out_flags & CA_OUT_DEVICE) { set_error("the multi-GPU calls return host arrays"); return CA_ERR_ARG; }
    const int G = (int)m->ctx.size();
    // The trajectory index of (point p, local i) is traj_offset + p*n_per_point + i, whichever device integrates it.  A device takes a
    // CONTIGUOUS block of points (so that one launch with one offset covers it), blocks of near-equal total pulse time.
    std::vector<int> first(G + 1, 0);
    {
        double tot = 0.0;
        for (int p = 0; p < n_points; ++p) tot += t_end[p];
        double run = 0.0;
        int g = 1;
        for (int p = 0; p < n_points && g < G; ++p) {
            run += t_end[p];
            while (g < G && run >= tot * g / G) first[g++] = p + 1;
        }
        for (; g < G; ++g) first[g] = n_points;
        first[G] = n_points;
    }
    std::vector<std::vector<double>> dev_out(G);
    ca_stats tot_st;
    std::memset(&tot_st, 0, sizeof tot_st);
    // asynchronous launches on every device first ...
    for (int g = 0; g < G; ++g) {
        const int p0 = first[g], np = first[g + 1] - p0;
        if (np <= 0) continue;
        CA_CUDAM(cudaSetDevice(ctx_device(m->ctx[g])));
        double* buf = (double*)ctx_buf(m->ctx[g], kBufMulti, sizeof(double) * ((size_t)np * 100 + 8));
        if (!buf) return CA_ERR_NOMEM;
        ca_stats s1;
        int rc = ca_rydberg1_sweep(m->ctx[g], models + p0, psi0s + 10 * (size_t)p0, t_end + p0, np, n_traj_per_point,
                                   traj_offset + (int64_t)p0 * n_traj_per_point, n_total, seed, f, amp_red, amp_blue, nf, ode,
                                   (out_flags & CA_OUT_SUMS) | CA_OUT_DEVICE, buf, buf + (size_t)np * 50, &s1);
        if (rc) return rc;
        tot_st.n_launches += s1.n_launches;
    }
    // ... then gather
    int status = CA_OK;
    for (int g = 0; g < G; ++g) {
        const int p0 = first[g], np = first[g + 1] - p0;
        if (np <= 0) continue;
        CA_CUDAM(cudaSetDevice(ctx_device(m->ctx[g])));
        double* buf = (double*)ctx_buf(m->ctx[g], kBufMulti, 0);
        CA_CUDAM(cudaMemcpyAsync(rho + 50 * (size_t)p0, buf, sizeof(double) * 50 * np, cudaMemcpyDeviceToHost, ctx_stream(m->ctx[g])));
        CA_CUDAM(cudaMemcpyAsync(rho2 + 50 * (size_t)p0, buf + (size_t)np * 50, sizeof(double) * 50 * np, cudaMemcpyDeviceToHost, ctx_stream(m->ctx[g])));
        ca_stats s1;
        const int rc = ca_collect_stats(m->ctx[g], &s1);   // synchronises this device
        if (rc == CA_ERR_CUDA || rc == CA_ERR_ARG) return rc;
        if (rc) status = std::min(status, rc);
        tot_st.n_traj += s1.n_traj; tot_st.n_rhs += s1.n_rhs; tot_st.n_accept += s1.n_accept; tot_st.n_reject += s1.n_reject;
        tot_st.n_sample_attempts += s1.n_sample_attempts; tot_st.n_exact_steps += s1.n_exact_steps;
        tot_st.kernel_ms = std::max(tot_st.kernel_ms, s1.kernel_ms); tot_st.total_ms = std::max(tot_st.total_ms, s1.total_ms);
    }
    tot_st.status = status;
    if (stats) *stats = tot_st;
    return status;
}

int ca_release_recapture_multi(ca_multi* m, const double* trap, const double* atom, const double* tspan, int32_t nt, int64_t n,
                               int64_t traj_offset, int64_t n_total, double eps, uint64_t seed, const double* samples, int32_t out_flags,
                               double* probs, ca_stats* stats) {
    if (!m || m->ctx.empty() || nt < 1 || n < 0 || !probs) { set_error("bad argument"); return CA_ERR_ARG; }
    if (out_flags & CA_OUT_DEVICE) { set_error("the multi-GPU calls return host arrays"); return CA_ERR_ARG; }
    const int G = (int)m->ctx.size();
    int launches = 0;
    for (int g = 0; g < G; ++g) {
        long long lo, cnt;
        shard(n, g, G, &lo, &cnt);
        CA_CUDAM(cudaSetDevice(ctx_device(m->ctx[g])));
        double* buf = (double*)ctx_buf(m->ctx[g], kBufMulti, sizeof(double) * ((size_t)nt + 8));
        if (!buf) return CA_ERR_NOMEM;
        CA_CUDAM(cudaMemsetAsync(buf + nt, 0, sizeof(double) * 8, ctx_stream(m->ctx[g])));
        ca_stats s1;
        int rc = ca_release_recapture(m->ctx[g], trap, atom, tspan, nt, cnt, traj_offset + lo, n_total, eps, seed, samples ? samples + 6 * lo : nullptr,
                                      CA_OUT_SUMS | CA_OUT_DEVICE, buf, nullptr, &s1);
        if (rc) return rc;
        launches += s1.n_launches;
    }
    std::vector<double> host;
    if (G > 1) {
        CA_NCCL(g_nccl.GroupStart());
        for (int g = 0; g < G; ++g) {
            double* buf = (double*)ctx_buf(m->ctx[g], kBufMulti, 0);
            CA_NCCL(g_nccl.AllReduce(buf, buf, (size_t)nt, 8, 0, m->comm[g], ctx_stream(m->ctx[g])));
        }
        CA_NCCL(g_nccl.GroupEnd());
    }
    host.resize(nt);
    CA_CUDAM(cudaSetDevice(ctx_device(m->ctx[0])));
    CA_CUDAM(cudaMemcpyAsync(host.data(), ctx_buf(m->ctx[0], kBufMulti, 0), sizeof(double) * nt, cudaMemcpyDeviceToHost, ctx_stream(m->ctx[0])));
    for (int g = 0; g < G; ++g) {
        CA_CUDAM(cudaSetDevice(ctx_device(m->ctx[g])));
        CA_CUDAM(cudaStreamSynchronize(ctx_stream(m->ctx[g])));
    }
    const double div = (out_flags & CA_OUT_SUMS) ? 1.0 : (double)(n_total > 0 ? n_total : n);
    for (int j = 0; j < nt; ++j) probs[j] = host[j] / div;
    if (stats) { std::memset(stats, 0, sizeof *stats); stats->n_traj = n; stats->n_launches = launches; }
    return CA_OK;
}

}  // extern "C"
