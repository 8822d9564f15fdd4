// Translation unit of the sparse (inducing-point) model; see sgpr_abi.cuh.
#include "sgpr_abi.cuh"
