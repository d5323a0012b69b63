// ca_model.h -- host-side model derivation and error plumbing (no CUDA dependency)
#pragma once
#include <string>

#include "../../include/coldatoms_b200.h"

namespace ca {
struct DevBeam;
struct DevModel;
struct OdeOpts;
void set_error(const std::string& s);
const char* get_error();
int derive_beam(const ca_beam& b, DevBeam* d);
int derive_model(const ca_model& m, const double* psi0, int npsi, double t0, double tend, DevModel* d);
void fill_ode(const ca_ode_opts* o, double t0, double tend, OdeOpts* d);
}  // namespace ca
