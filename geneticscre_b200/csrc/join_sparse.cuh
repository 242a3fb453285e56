// Sparse (carrier-list) join kernels -- placeholder until the dense path is parity-green on the GPU.
#pragma once
#include "common.cuh"

namespace gcre {

static inline bool sparse_supported(int /*n*/, long long /*t_needed*/) { return false; }

static inline int launch_join_sparse(cudaStream_t, const JoinParams&, int, bool, int, int*) { return -3; }

}  // namespace gcre
