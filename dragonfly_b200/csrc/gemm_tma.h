// Host-visible declarations of the TMA scoring kernel (kernel body: gemm_tma.cuh, built in kernels.cu).
#pragma once
#include <cuda.h>
#include "common.cuh"

namespace dfb {

struct ScoreTmaArgs {
  int n_rb, n_cb, K;
  double* partial;
  int64_t ld_partial;
};

int make_tensor_map_2d_f64(CUtensorMap* out, const double* base, int64_t rows, int64_t cols_ld,
                           int64_t cols);
int launch_score_tma(dfb_handle* h, const CUtensorMap& tmW, const CUtensorMap& tmK,
                     const ScoreTmaArgs& g);

}  // namespace dfb
