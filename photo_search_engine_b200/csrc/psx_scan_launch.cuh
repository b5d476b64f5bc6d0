// psx_scan_launch.cuh -- host-side entry to the scan kernel's instantiations.  The kernel template
// (psx_scan.cuh) is instantiated per (stored type, metric) in its own translation unit
// (psx_scan_<type>_<metric>.cu) so that the library builds in parallel.
#pragma once
#include "psx_scan.cuh"

namespace psx {

struct ScanLaunch {
    int grid, block;
    size_t smem;
    bool pdl;  // launch with programmatic stream serialization (may start before the previous kernel has drained)
};

// ppl: 16-byte pieces per lane of a row/chunk when the row shape allows the unrolled dot (0 = generic loop);
// qreg: the query block lives in registers; mode: PSX_SCAN_DEAL or PSX_SCAN_GROUPS.
template <typename T, int METRIC>
cudaError_t launch_scan_shape(int device, int ppl, bool qreg, int mode, const ScanParams& p, const ScanLaunch& l, cudaStream_t st);

extern template cudaError_t launch_scan_shape<float, PSX_METRIC_IP>(int, int, bool, int, const ScanParams&, const ScanLaunch&, cudaStream_t);
extern template cudaError_t launch_scan_shape<float, PSX_METRIC_L2>(int, int, bool, int, const ScanParams&, const ScanLaunch&, cudaStream_t);
extern template cudaError_t launch_scan_shape<__nv_bfloat16, PSX_METRIC_IP>(int, int, bool, int, const ScanParams&, const ScanLaunch&, cudaStream_t);
extern template cudaError_t launch_scan_shape<__nv_bfloat16, PSX_METRIC_L2>(int, int, bool, int, const ScanParams&, const ScanLaunch&, cudaStream_t);

}  // namespace psx
