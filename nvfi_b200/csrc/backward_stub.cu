// Temporary: backward entry points (replaced by backward.cu).
#include "nvfi_common.cuh"
extern "C" int64_t nvfi_backward_partials_bytes(void) { return 0; }
extern "C" int nvfi_render_backward(const NvfiField*, const NvfiRenderArgs*, const NvfiRenderBuffers*,
                                    const NvfiRenderGrads*, void*) {
  return NVFI_EUNSUPPORTED;
}
