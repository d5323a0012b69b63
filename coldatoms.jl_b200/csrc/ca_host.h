// ca_host.h -- host-side declarations shared by the translation units of libcoldatoms_b200.so
#pragma once
#include <cuda_runtime.h>

#include "ca_model.h"

