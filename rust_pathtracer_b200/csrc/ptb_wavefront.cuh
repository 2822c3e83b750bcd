// ptb_wavefront.cuh — wavefront integrator (SoA queues, one kernel per stage).  Placeholder until
// the stage kernels land; the fused integrator is the default.
#pragma once
#include <string>
#include "ptb_kernels.cuh"

namespace ptb {
struct WavefrontState {
    void release() {}
};
inline int wavefront_render(WavefrontState&, const DScene<float>&, void*, uint32_t, uint32_t, uint32_t, uint64_t, const ptb_config&, cudaStream_t,
                            int, DeviceCounters*, cudaEvent_t, cudaEvent_t, uint64_t*, std::string& err) {
    err = "wavefront integrator not built yet";
    return PTB_E_UNSUPPORTED;
}
}  // namespace ptb
