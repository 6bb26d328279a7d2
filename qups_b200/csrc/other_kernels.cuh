// other_kernels.cuh — launchers for the non-DAS entry points of the C ABI.
#pragma once
#include <cuda_runtime.h>
#include "../../include/qups_b200.h"

namespace qups {
// return 0, a cudaError_t (> 0), -3 (unsupported) or -4 (allocation)
int launch_wsinterpd2(const qups_ws2_params &p, void *y, const void *w, const void *x, const void *t1, const void *t2,
                      cudaStream_t st);
int launch_greens(const qups_greens_params &p, void *y, const void *Pi, const void *a, const void *Pr, const void *Pv,
                  const void *kern, cudaStream_t st);
int launch_convd(const qups_convd_params &p, void *z, const void *x, const void *y, cudaStream_t st);
} // namespace qups
