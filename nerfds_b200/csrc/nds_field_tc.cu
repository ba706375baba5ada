// tcgen05 tensor-core field engine (placeholder until the kernel lands).
#include "nds_host.h"

namespace nds {
std::string tc_engine_supports(const ndsr_config&, int, int) { return "tensor-core engine not built yet"; }
int tc_engine_load(ndsr_handle*) { return NDSR_ERR_UNSUPPORTED; }
int tc_engine_field(ndsr_handle*, const CallParams&, const FieldArgs&, cudaStream_t) { return NDSR_ERR_UNSUPPORTED; }
void tc_engine_free(ndsr_handle*) {}
}  // namespace nds
