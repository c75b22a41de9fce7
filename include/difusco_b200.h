/* difusco_b200 - C-ABI of the B200-native DIFUSCO denoising-inference hot path.
 *
 * The reference (Edward-Sun/DIFUSCO) has no FFI / plugin interface of its own: its seams are Python
 * methods.  This header is the boundary the Python host mirror (difusco_b200/*.py) binds with
 * ctypes; every entry point names the reference interface whose device work it replaces
 * (paths relative to /root/reference/difusco).  Plain C types only; pointers are raw host or
 * device pointers as stated; `stream` is a cudaStream_t passed as void*.  Every function returns
 * 0 on success or a negative DFB_E_* code; dfb_last_error() gives the message.  Nothing throws
 * across the boundary.  A context is bound to one device and is not thread-safe (one per GPU /
 * per process, as in the reference's one-process-per-GPU DDP launch, train.py:106-115).
 *
 * Buffer ownership: every input / output buffer named in a signature is the caller's.  Scratch memory is ONE
 * arena owned by the context: dfb_load_weights and dfb_prepare_graph size it (device malloc happens only there, and only
 * when a graph is larger than anything prepared before); dfb_set_points, dfb_encoder_forward, dfb_denoise_step and
 * dfb_denoise never allocate device or host memory, never synchronise the host with the stream and never read the
 * environment: per-call scalars travel through two pinned staging slots, per-step tables live in device memory, so
 * the step path is CUDA-graph capturable - and dfb_denoise itself replays a captured graph.
 */
#ifndef DIFUSCO_B200_H_
#define DIFUSCO_B200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define DFB_ABI_VERSION 2

enum {
  DFB_OK = 0,
  DFB_E_INVALID = -1,   /* bad argument / state (maps to ValueError on the Python side)   */
  DFB_E_CUDA = -2,      /* a CUDA runtime / driver call failed                              */
  DFB_E_UNSUPPORTED = -3, /* reachable reference flag this build does not implement (NotImplementedError) */
  DFB_E_NOMEM = -4
};

enum { DFB_TASK_TSP = 0, DFB_TASK_MIS = 1 };               /* pl_tsp_model.py / pl_mis_model.py          */
enum { DFB_DIFFUSION_CATEGORICAL = 0, DFB_DIFFUSION_GAUSSIAN = 1 }; /* pl_meta_model.py:27-36           */
enum { DFB_EDGE_IMPL_TC = 0, DFB_EDGE_IMPL_FP32 = 1, DFB_EDGE_IMPL_TC1 = 2 }; /* tcgen05 product path (CTA-pair kernel on the
  middle layers) / fp32 validation kernel / the single-CTA tcgen05 kernel on every layer (A/B and validation) */

typedef struct dfb_ctx dfb_ctx;

int dfb_abi_version(void);

/* Create a context on CUDA device `device`.  Fails (DFB_E_CUDA) when no device is present:
 * there is no CPU fallback.  The environment is read here and nowhere else (A/B switches, all default to the product
 * path): DFB_GRAPH_CAPTURE=0 plain launches instead of the captured loop, DFB_PAIR_KERNEL=0 single-CTA edge kernel for
 * every layer, DFB_SERPENTINE=0 every layer sweeps the edge stream upwards, DFB_TC_PROBE tuning-build counters. */
int dfb_create(dfb_ctx** out, int device);
int dfb_destroy(dfb_ctx* ctx);
/* Message of the last failure on `ctx` (or of the last failed dfb_create when ctx == NULL). */
const char* dfb_last_error(const dfb_ctx* ctx);

/* --aggregation flag (train.py:63; gnn_encoder.py:184-191): 0 sum (default), 1 mean, 2 max. */
int dfb_set_aggregation(dfb_ctx* ctx, int mode);

/* Select the fused edge-layer implementation (tests only; default DFB_EDGE_IMPL_TC). */
int dfb_set_edge_impl(dfb_ctx* ctx, int impl);

/* GNNEncoder.__init__ + load_state_dict (models/gnn_encoder.py:294-348).
 * `names[i]` are GNNEncoder.state_dict() keys (an optional leading "model." - the Lightning
 * checkpoint prefix, pl_meta_model.py:38 - is accepted), `tensors[i]` HOST fp32 pointers with
 * `numels[i]` elements.  hidden_dim must be 256; out_channels 1 (gaussian) or 2 (categorical);
 * node_feature_only 0 (TSP) / 1 (MIS).  Missing or mis-sized tensors -> DFB_E_INVALID. */
int dfb_load_weights(dfb_ctx* ctx, int n_layers, int hidden_dim, int out_channels,
                     int node_feature_only, int n_tensors, const char* const* names,
                     const float* const* tensors, const int64_t* numels);

/* The graph of one forward call: edge_index (2,E) int64, row = edge_index[0] (owner node),
 * col = edge_index[1] (gnn_encoder.py:110,417-423).  HOST or DEVICE pointer (detected).
 * Need not be row-sorted (MIS is not, mis_dataset.py:43-48): a stable row sort is kept
 * internally and all edge-valued I/O stays in the caller's edge order.
 * gn_segments: number of equal consecutive row blocks the head GroupNorm normalises separately:
 * 1 for every sparse call (gnn_encoder.py:400-401: batch dim 1 over ALL edges), B for the dense
 * API with B samples (gnn_encoder.py:380). */
int dfb_prepare_graph(dfb_ctx* ctx, const int64_t* edge_index, int64_t num_nodes, int64_t num_edges,
                      int gn_segments, void* stream);

/* TSP only: node coordinates (V,2) fp32, HOST or DEVICE.  Computes the step-invariant
 * h0 = node_embed(pos_embed(x)) (gnn_encoder.py:394, :211-227) and layer 0's node linears. */
int dfb_set_points(dfb_ctx* ctx, const float* points, void* stream);

/* GNNEncoder.forward (gnn_encoder.py:452-462) on the prepared graph.  DEVICE pointers.
 *   TSP: xt (E,) fp32 edge values in caller edge order, out (E, out_channels)
 *   MIS: xt (V,) fp32 node values,                         out (V, out_channels)
 * `t` is the (single) timestep, as at inference (pl_tsp_model.py:124-130). */
int dfb_encoder_forward(dfb_ctx* ctx, const float* xt, float t, float* out, void* stream);

/* One reverse-diffusion step = *_denoise_step (pl_tsp_model.py:122-151, pl_mis_model.py:118-140)
 * = forward + softmax + categorical_posterior (pl_meta_model.py:102-146) or gaussian_posterior
 * (:148-175), fused on the device.
 *   t            source timestep t1 fed to the network
 *   consts       HOST, 4 floats computed by the caller from the float64 schedule tables:
 *                categorical: {c[0][0], c[0][1], c[1][0], c[1][1]} with
 *                   p = c[xt][0]*p0[0] + c[xt][1]*p0[1]   (closed form of :113-137)
 *                gaussian: {a, b1, b2, noise} with xt' = a*(xt - b1*pred) + b2*pred + noise*z
 *   last         1 when target_t == 0: categorical returns clamp(p, min=0) (the heatmap)
 *                instead of a Bernoulli sample (:139-142)
 *   uniforms     DEVICE (N,) injected U[0,1) (categorical) / N(0,1) (gaussian ddpm) draws or
 *                NULL -> in-kernel Philox4x32-10 keyed by (seed, step_index, element)
 *   xt_in/xt_out DEVICE (N,), N = E (TSP) or V (MIS); may alias
 *   p_out        DEVICE (N,) optional: pre-sampling probability p (categorical)
 *   net_out      DEVICE (N,out_channels) optional: raw network output */
int dfb_denoise_step(dfb_ctx* ctx, int diffusion_type, const float* xt_in, float t,
                     const float* consts, int last, const float* uniforms, uint64_t seed,
                     int step_index, float* xt_out, float* p_out, float* net_out, void* stream);

/* The whole loop of test_step (pl_tsp_model.py:207-217 / pl_mis_model.py:176-186): `steps`
 * denoise steps on DEVICE buffers, no host synchronisation inside.
 *   t1          HOST (steps,) source timesteps; consts HOST (steps,4); last_flags HOST (steps,)
 *   uniforms    DEVICE (steps,N) or NULL (Philox)
 * xt is updated in place; after the call it holds the raw heatmap (clamp(p,min=0)) or the
 * gaussian xt (the caller applies +1e-6 / *0.5+0.5, pl_tsp_model.py:219-222). */
int dfb_denoise(dfb_ctx* ctx, int diffusion_type, float* xt, int steps, const int32_t* t1,
                const float* consts, const int32_t* last_flags, const float* uniforms,
                uint64_t seed, void* stream);

/* dfb_denoise replays the whole loop as ONE captured CUDA graph (on a stream of the library, fenced to `stream` by
 * events; re-captured only when the prepared graph, the buffers, the implementation switches or `steps` change).
 * dfb_set_graph_capture(ctx, 0) turns that off (plain launches).  Environment DFB_GRAPH_CAPTURE=0 does the same. */
int dfb_set_graph_capture(dfb_ctx* ctx, int enabled);

/* End-to-end with HOST buffers (the call bench.py times as `e2e`): H2D of points / edge_index /
 * xt0, graph preparation, `steps` denoise steps, D2H of the final xt into heatmap_out (N,).
 * points may be NULL for MIS. */
int dfb_denoise_host(dfb_ctx* ctx, int diffusion_type, const float* points,
                     const int64_t* edge_index, int64_t num_nodes, int64_t num_edges,
                     int gn_segments, const float* xt0, int steps, const int32_t* t1,
                     const float* consts, const int32_t* last_flags, uint64_t seed,
                     float* heatmap_out, void* stream);

/* Row f1 (the step BEFORE the path): sparse k-NN graph of one TSP instance on the GPU.  Replaces the KDTree query +
 * edge_index assembly of TSPGraphDataset.__getitem__ (co_datasets/tsp_graph_dataset.py:52-62): float64 coordinates
 * (HOST or DEVICE), neighbours in ascending euclidean distance with self first, edge_index (2, N*k) int64 DEVICE with
 * row = arange(N).repeat_interleave(k) and both rows shifted by node_offset (block-diagonal batching). */
int dfb_knn_graph(dfb_ctx* ctx, const double* points, int64_t num_nodes, int k, int64_t node_offset,
                  int64_t* edge_index, void* stream);

/* Row f2 (the step AFTER the path): greedy edge-insertion tour merge, utils/tsp_utils.py:89-145 with
 * utils/cython_merge/cython_merge.pyx:19-120.  HOST code, HOST pointers, no context.
 *   points (n,2) float64; heat (E,) float32 over edge_index (2,E) int64 (the dense case passes the row-major complete
 *   graph with heat = adj.flatten()); tour (n+1,) int64 out; merge_iterations out (the reference's counter).
 * Only the non-zero heat entries are sorted instead of the reference's dense n*n argsort.
 *   mode 0: returns 0 when the tour completes inside them (identical to the reference), 1 when it does not (the rest
 *           of the reference's order is a tie at key 0 whose order is numpy's argsort artefact: the caller then runs
 *           dfb_tsp_merge_order on that argsort), 2 when two different pairs tie exactly (same fallback).
 *   mode 1: never falls back; leftover fragment ends are joined by increasing distance (NOT the reference's result
 *           once the non-zero entries run out; opt-in for large n).
 * Negative return: DFB_E_INVALID. */
int dfb_tsp_merge_sparse(const double* points, int64_t n, const float* heat, const int64_t* edge_index, int64_t E,
                         int mode, int64_t* tour, int64_t* merge_iterations);
/* The reference loop over an explicit visiting order of the flattened n*n entries (cython_merge.pyx:44-98). */
int dfb_tsp_merge_order(int64_t n, const int64_t* order, int64_t count, int64_t* tour, int64_t* merge_iterations);

/* Row f3: batched 2-opt, utils/tsp_utils.py:12-49 (batched_two_opt_torch).  points (n,2) float64 HOST, tours
 * (batch, n+1) int64 HOST, updated in place; iterations_out = the reference's `iterator`.  Same moves in the same
 * order as the reference on its CPU device (float64, first-occurrence arg-min, batch-wide stopping rule). */
int dfb_two_opt(dfb_ctx* ctx, const double* points, int64_t n, int64_t* tours, int64_t batch, int64_t max_iterations,
                int64_t* iterations_out, void* stream);

/* Row f4: the MCTS solver's text heat map (tsp_mcts/convert_numpy_to_txt.py:57-73; parsed by tsp_mcts/code/TSP_IO.h:461-492):
 * "<n>\n" then n lines of n values "%.6f" separated by one blank.  matrix (n,n) float64 HOST.  HOST code, no context.
 * Returns DFB_E_INVALID when the file cannot be written. */
int dfb_write_heatmap_txt(const char* path, int64_t n, const double* matrix);

/* Number of kernels this context has launched since creation (bench.py's gpu_launches). */
int64_t dfb_launch_count(const dfb_ctx* ctx);

/* Device-side duration in ms and launch count of the fused edge-layer kernel accumulated between
 * dfb_profile_begin / dfb_profile_end (CUDA events on the launching stream; roofline.achieved). */
int dfb_profile_begin(dfb_ctx* ctx);
int dfb_profile_end(dfb_ctx* ctx, double* edge_kernel_ms, int64_t* edge_kernel_launches);

/* Test hook: GEMM1 only (acc = e_in * C_layer^T on the tensor-core path), accumulator dumped to
 * acc_out (E,256).  DEVICE pointers.  Used by the parity tests to localise failures. */
int dfb_debug_edge_gemm(dfb_ctx* ctx, int layer, const float* e_in, float* acc_out, void* stream);

/* Tuning hook: per-phase cycle counters of the tcgen05 edge kernels (DFB_TC_PROBE bit 7, --prof build); out must hold
 * 32 unsigned 64-bit values (host): [0..7] phases and [8..15] E1 sub-phases of the single-CTA kernel, [16..23] phases of
 * the CTA-pair kernel.  Read-and-reset. */
int dfb_debug_phase_cycles(dfb_ctx* ctx, unsigned long long* out);

/* Diagnostic: watchdog record of the tcgen05 kernel's bounded barrier waits (host-mapped memory, readable after a
 * launch failure): out[4] = {wait-site code or 0, blockIdx.x, parity, threadIdx.x}. */
int dfb_debug_watchdog(dfb_ctx* ctx, int* out);

#ifdef __cplusplus
}
#endif
#endif /* DIFUSCO_B200_H_ */
