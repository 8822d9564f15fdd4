// Translation unit of the sparse (inducing-point) model; see sgpr_abi.cuh (one model per handle) and sgpr_batch.cuh (the
// per-column models of one fit, batched and trained on the device).
#include "sgpr_batch.cuh"
