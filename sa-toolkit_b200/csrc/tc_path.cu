// placeholder until the tcgen05 path lands
#include "tc_path.cuh"
namespace sa {
bool tc_layer_supported(bool, int, int, int) { return false; }
const char* tc_pack_weights(tc_weights&, const float*, bool, int, int, int, int, int, bool) { return "tensor-core path not built"; }
void tc_free_weights(tc_weights& w) { if (w.d_w) cudaFree(w.d_w); w = tc_weights(); }
const char* tc_init(tc_context&, int) { return "tensor-core path not built"; }
size_t tc_workspace_bytes(const sa_hifigan_cfg&, int, int) { return 0; }
const char* tc_forward(tc_context&, const tc_forward_args&, int64_t*) { return "tensor-core path not built"; }
}
