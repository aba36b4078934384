// umma.h - interface of the tcgen05 (5th-gen tensor core) path to the engine.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include <string>
#include <vector>

#include "plan.h"

namespace wun {

struct UmmaState {
    std::vector<int> fwd_ok;     // per conv layer (0..2L): 1 if the forward runs on tcgen05
    bool enabled = false;
};

void umma_init(UmmaState* st, const Plan& plan);
void umma_destroy(UmmaState* st);
int64_t umma_workspace_floats(const UmmaState& st, const Plan& plan, int64_t batch, bool training);
// returns true if it enqueued (or, when dry, counted) the layer's forward
bool umma_try_forward(UmmaState& st, const Plan& plan, const ConvOp& op, int layer_index, const float* params,
                      const float* mix, float* ws, const int64_t* tensor_off, int batch, cudaStream_t stream, bool dry,
                      int64_t* launches);
std::string umma_describe(const UmmaState& st, const Plan& plan);
const char* umma_layer_kernel(const UmmaState& st, int layer, int pass);

}  // namespace wun
