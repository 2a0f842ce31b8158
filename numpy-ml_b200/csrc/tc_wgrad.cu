// placeholder until the tcgen05 kernel lands
#include "npml_common.cuh"
namespace npml {
bool tc_wgrad_supported(const GatherWgrad&, int) { return false; }
size_t tc_wgrad_ws_bytes(const GatherWgrad&, int) { return 0; }
int tc_wgrad(const GatherWgrad&, int, const void*, const void*, float*, void*, size_t, cudaStream_t) { set_error("tc path not built"); return 9; }
}
