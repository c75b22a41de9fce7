// Row f1 of SURVEY 8(f): the sparse k-NN graph that feeds the denoise path.
// Reference: co_datasets/tsp_graph_dataset.py:52-62 - sklearn KDTree(leaf_size=30, euclidean).query(points, k=K) on
// float64 coordinates; neighbours in ascending distance (self first), edge_index = [repeat_interleave(arange(N), K);
// knn.flatten()].  Here: brute force in fp64 (N <= 10 000 in every reference recipe: 80 KB of squared distances per
// query point in shared memory), one block per query point, K rounds of block-wide arg-min.  Squared distances are
// formed with separate multiplies and adds (no FMA contraction) so the ordering is the one the CPU computes; exact ties
// resolve to the smaller index.
#pragma once
#include "common.cuh"

namespace dfb {

__global__ void __launch_bounds__(256) k_knn_bruteforce(const double* __restrict__ pts, int N, int K,
                                                        long long* __restrict__ edge_index /* [2][N*K] */,
                                                        long long node_offset) {
  extern __shared__ double d2[];
  __shared__ double w_val[8];
  __shared__ int w_idx[8];
  __shared__ int s_pick;
  const int q = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const double qx = pts[2 * q], qy = pts[2 * q + 1];
  for (int j = tid; j < N; j += 256) {
    const double dx = __dsub_rn(pts[2 * j], qx), dy = __dsub_rn(pts[2 * j + 1], qy);
    d2[j] = __dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy));
  }
  __syncthreads();
  for (int k = 0; k < K; ++k) {
    double best = INFINITY;
    int bi = 0x7fffffff;
    for (int j = tid; j < N; j += 256) {
      const double v = d2[j];
      if (v < best || (v == best && j < bi)) { best = v; bi = j; }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const double ov = __shfl_xor_sync(0xffffffffu, best, o);
      const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
      if (ov < best || (ov == best && oi < bi)) { best = ov; bi = oi; }
    }
    if (lane == 0) { w_val[warp] = best; w_idx[warp] = bi; }
    __syncthreads();
    if (tid == 0) {
      double b = w_val[0];
      int i = w_idx[0];
      for (int w = 1; w < 8; ++w)
        if (w_val[w] < b || (w_val[w] == b && w_idx[w] < i)) { b = w_val[w]; i = w_idx[w]; }
      s_pick = i;
      d2[i] = INFINITY;
      edge_index[(size_t)q * K + k] = node_offset + q;                       // row: owner node
      edge_index[(size_t)N * K + (size_t)q * K + k] = node_offset + i;       // col: k-th nearest neighbour
    }
    __syncthreads();
  }
}

}  // namespace dfb
