// ca_lindblad2.cu -- two-atom CZ ensemble (simulation_czlp, src/cz_model.jl:107-171): placeholder until K4 lands.
#include "ca_host.h"

extern "C" int ca_rydberg2_run(ca_ctx*, const ca_model*, const double*, int32_t, const double*, int64_t, int64_t, int64_t, uint64_t,
                               const double*, const double*, const double* const*, const double*, const double*, const double*, int32_t,
                               const ca_ode_opts*, int32_t, double*, double*, ca_stats*) {
    ca::set_error("ca_rydberg2_run: the two-atom kernel is not built into this library yet");
    return CA_ERR_UNSUPPORTED;
}
