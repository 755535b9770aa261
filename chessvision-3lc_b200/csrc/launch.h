// Host-side kernel launch helper (internal): every kernel that calls griddep_wait() (common.cuh) is launched with
// programmatic stream serialization, so its prologue overlaps the tail of the kernel in front of it.
#pragma once
#include <cuda_runtime.h>
#include <stdlib.h>

namespace cvb {

// CVB_PDL=0 turns programmatic dependent launch off (A/B measurements).
inline bool pdl_enabled() {
    static const bool on = [] { const char* e = getenv("CVB_PDL"); return !(e && e[0] == '0'); }();
    return on;
}

template <class... KArgs, class... Args>
inline cudaError_t launch_k(void (*kernel)(KArgs...), int grid, int block, size_t smem, cudaStream_t s, bool pdl, Args&&... args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(grid);
    cfg.blockDim = dim3(block);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = s;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = pdl && pdl_enabled() ? 1 : 0;
    return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}

}  // namespace cvb
