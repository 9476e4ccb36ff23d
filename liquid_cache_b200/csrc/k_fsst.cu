// k_fsst.cu — FSST compression of the dictionary values at insert time (device side).
// Placeholder translation unit: the compress kernel lands here (see DESIGN.md "insert path").
#include "kernels.h"
namespace lc {
cudaError_t launch_fsst_compress(const FsstCompressWork*, uint32_t, cudaStream_t) { return cudaErrorNotSupported; }
}  // namespace lc
