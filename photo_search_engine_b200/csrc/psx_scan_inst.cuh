// psx_scan_inst.cuh -- definitions behind psx_scan_launch.cuh; included only by the
// psx_scan_<type>_<metric>.cu translation units, which instantiate them explicitly.
#pragma once
#include <atomic>

#include "psx_scan_launch.cuh"

namespace psx {

template <typename T, int METRIC, int PPL, bool QREG, int MODE>
static cudaError_t launch_scan_one(int device, const ScanParams& p, const ScanLaunch& l, cudaStream_t st) {
    static std::atomic<bool> ready[64];
    if (device >= 0 && device < 64 && !ready[device].load()) {
        cudaError_t e = cudaFuncSetAttribute(scan_topk_kernel<T, METRIC, PPL, QREG, MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                             PSX_SMEM_LIMIT);
        if (e != cudaSuccess) return e;
        ready[device].store(true);
    }
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)l.grid);
    cfg.blockDim = dim3((unsigned)l.block);
    cfg.dynamicSmemBytes = l.smem;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = l.pdl ? 1 : 0;
    return cudaLaunchKernelEx(&cfg, scan_topk_kernel<T, METRIC, PPL, QREG, MODE>, p);
}

template <typename T, int METRIC, int MODE>
static cudaError_t launch_scan_mode(int device, int ppl, bool qreg, const ScanParams& p, const ScanLaunch& l, cudaStream_t st) {
    if (qreg) {
        switch (ppl) {
            case 1: return launch_scan_one<T, METRIC, 1, true, MODE>(device, p, l, st);
            case 2: return launch_scan_one<T, METRIC, 2, true, MODE>(device, p, l, st);
            case 3: return launch_scan_one<T, METRIC, 3, true, MODE>(device, p, l, st);
            case 4: return launch_scan_one<T, METRIC, 4, true, MODE>(device, p, l, st);
            case 6: return launch_scan_one<T, METRIC, 6, true, MODE>(device, p, l, st);
            case 8: return launch_scan_one<T, METRIC, 8, true, MODE>(device, p, l, st);
            default: break;
        }
    } else if (ppl == 8) {
        return launch_scan_one<T, METRIC, 8, false, MODE>(device, p, l, st);
    }
    return launch_scan_one<T, METRIC, 0, false, MODE>(device, p, l, st);
}

template <typename T, int METRIC>
cudaError_t launch_scan_shape(int device, int ppl, bool qreg, int mode, const ScanParams& p, const ScanLaunch& l, cudaStream_t st) {
    return mode == PSX_SCAN_GROUPS ? launch_scan_mode<T, METRIC, PSX_SCAN_GROUPS>(device, ppl, qreg, p, l, st)
                                   : launch_scan_mode<T, METRIC, PSX_SCAN_DEAL>(device, ppl, qreg, p, l, st);
}

}  // namespace psx
