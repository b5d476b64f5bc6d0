// scan kernel instantiations: stored type __nv_bfloat16, metric PSX_METRIC_L2
#include "psx_scan_inst.cuh"

namespace psx {
template cudaError_t launch_scan_shape<__nv_bfloat16, PSX_METRIC_L2>(int, int, bool, int, const ScanParams&, const ScanLaunch&, cudaStream_t);
}
