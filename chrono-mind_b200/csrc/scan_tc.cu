// scan_tc.cu — K1/K2: tcgen05 bf16 scan + fp32 rescore. (placeholder until the kernel lands)
#include "kernels.h"
namespace cm {
cudaError_t scan_tc_supported(const DeviceIndex&) { return cudaErrorNotSupported; }
}  // namespace cm
