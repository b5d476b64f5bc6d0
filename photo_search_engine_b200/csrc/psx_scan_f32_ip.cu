// scan kernel instantiations: stored type float, metric PSX_METRIC_IP
#include "psx_scan_inst.cuh"

namespace psx {
template cudaError_t launch_scan_shape<float, PSX_METRIC_IP>(int, int, bool, int, const ScanParams&, const ScanLaunch&, cudaStream_t);
}
