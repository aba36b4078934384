// umma.cu - tcgen05 path (placeholder until the tensor-core kernels land: every layer reports "simt").
#include "umma.h"

namespace wun {

void umma_init(UmmaState* st, const Plan& plan) {
    st->fwd_ok.assign(2 * plan.cfg.num_layers + 1, 0);
    st->enabled = false;
}
void umma_destroy(UmmaState*) {}
int64_t umma_workspace_floats(const UmmaState&, const Plan&, int64_t, bool) { return 0; }
bool umma_try_forward(UmmaState&, const Plan&, const ConvOp&, int, const float*, const float*, float*, const int64_t*,
                      int, cudaStream_t, bool, int64_t*) {
    return false;
}
std::string umma_describe(const UmmaState&, const Plan&) { return "  tensor-core path: disabled\n"; }
const char* umma_layer_kernel(const UmmaState& st, int layer, int pass) {
    if (pass == 0 && layer >= 0 && layer < (int)st.fwd_ok.size() && st.fwd_ok[layer]) return "umma";
    return "simt";
}

}  // namespace wun
