// tcgen05 / TMA path (DI_MATH_TF32) -- placeholder until the kernels land: creating a TF32 engine fails loudly.
#include "engine.h"
#include "tc_common.cuh"

namespace di {
bool tc_available() { return false; }
bool tc_init(Engine& e) { e.err = "DI_MATH_TF32 kernels are not built in this revision"; return false; }
void tc_destroy(Engine&) {}
bool tc_rebind(Engine&) { return true; }
void tc_train_step(Engine&, const StepArgs&, int) {}
void tc_forward(Engine&, int, int64_t, int64_t, int64_t, bool, float*, int64_t) {}
}  // namespace di
