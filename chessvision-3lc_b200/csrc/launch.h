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

// cluster > 1: thread-block clusters of that many consecutive CTAs (the CTA pairs of cta_group::2 kernels); grid % cluster == 0.
template <class... KArgs, class... Args>
inline cudaError_t launch_kc(void (*kernel)(KArgs...), int grid, int block, size_t smem, cudaStream_t s, bool pdl, int cluster, Args&&... args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(grid);
    cfg.blockDim = dim3(block);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = s;
    cudaLaunchAttribute attr[2];
    int n = 0;
    if (pdl && pdl_enabled()) {
        attr[n].id = cudaLaunchAttributeProgrammaticStreamSerialization;
        attr[n].val.programmaticStreamSerializationAllowed = 1;
        ++n;
    }
    if (cluster > 1) {
        attr[n].id = cudaLaunchAttributeClusterDimension;
        attr[n].val.clusterDim.x = static_cast<unsigned>(cluster);
        attr[n].val.clusterDim.y = 1;
        attr[n].val.clusterDim.z = 1;
        ++n;
    }
    cfg.attrs = attr;
    cfg.numAttrs = n;
    return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}

template <class... KArgs, class... Args>
inline cudaError_t launch_k(void (*kernel)(KArgs...), int grid, int block, size_t smem, cudaStream_t s, bool pdl, Args&&... args) {
    return launch_kc(kernel, grid, block, smem, s, pdl, 1, static_cast<Args&&>(args)...);
}

}  // namespace cvb
