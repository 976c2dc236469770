/*
 * scanerf_b200.h -- C ABI of libscanerf_b200.so, the drop-in boundary of the
 * B200-native (sm_100a) ScaNeRF hot path.
 *
 * Conventions
 *   - every pointer is a DEVICE pointer unless its name ends in _host;
 *   - row-major, dense (contiguous) arrays; shapes are given in the comments;
 *   - outputs are written in place into caller-allocated memory, elements the
 *     reference leaves untouched (sentinel fills) are left untouched here too;
 *   - `stream` is a cudaStream_t (NULL = legacy default stream, which is what the
 *     reference launches on);
 *   - return value 0 = launched; otherwise a cudaError_t, and snrf_last_error()
 *     returns a thread-local description.  Asynchronous faults surface at the next
 *     synchronisation, as with the reference's AT_CUDA_CHECK(cudaGetLastError()).
 * Each entry point cites the reference interface (file:line under the reference
 * checkout) it replaces.
 */
#ifndef SCANERF_B200_H
#define SCANERF_B200_H

#ifdef __cplusplus
extern "C" {
#endif

/* ---- library ---------------------------------------------------------------- */
const char* snrf_last_error(void);
int snrf_version(void);
int snrf_device_sm_count(void);

/* ---- hash-grid encode ------------------------------------------------------- */
/* hashgrid/include/hashgrid.h:19-25 (embedding_forward_cuda, corner/size != NULL) and
 * :38-42 (embedding_bg_forward_cuda, corner = size = NULL).
 * points[B,3] f32, table[L,T,2] f32 (T power of two), res[L,3] i32, corner[3], size[3]
 * -> out[B,L,2] f32 (out_bf16=0) or out[B,2L] bf16 (out_bf16=1, tensor-core MLP input).
 * idx_out (optional) [B,L,8] u32: the hashed table indices of the 8 cell corners. */
int snrf_hash_fwd(const float* points, const float* table, const int* res, const float* corner,
                  const float* size, void* out, unsigned* idx_out, int B, int L, int T, int out_bf16,
                  void* stream);
/* hashgrid/include/hashgrid.h:27-35 / :45-51 (embedding_backward_cuda, embedding_bg_backward_cuda).
 * grad_in[B,L,2]; ACCUMULATES into grad_points[B,3] (may be NULL: skip d/dx) and
 * grad_table[L,T,2].  aggregate_levels: levels [0,n) use the warp-aggregated scatter
 * (-1 = default L/2). */
int snrf_hash_bwd(const float* points, const float* grad_in, const float* table, const int* res,
                  const float* corner, const float* size, float* grad_points, float* grad_table,
                  int B, int L, int T, int aggregate_levels, void* stream);
/* tuning hook: force the number of levels walked per CTA row (0 = automatic) */
void snrf_hash_set_levels_per_block(int lpb);

/* ---- rays, boxes, samplers -------------------------------------------------- */
/* cuda/include/compute_ray.h (compute_ray_forward): Ks[N,9], C2Ws[N,12], locs[B,3] i32
 * = (view, px, py) -> rays_o, rays_d [B,3]. */
int snrf_compute_ray_fwd(float* rays_o, float* rays_d, const float* Ks, const float* C2Ws,
                         const int* locs, int B, void* stream);
/* cuda/include/compute_ray.h (compute_ray_backward): accumulates grad_C2Ws[N,12].
 * ref_index_bug=1 reproduces cuda/compute_ray_kernel.cu:71-72 (gradients read at view_idx). */
int snrf_compute_ray_bwd(const float* grad_o, const float* grad_d, const float* Ks, float* grad_C2Ws,
                         const int* locs, int B, int ref_index_bug, void* stream);
/* cuda/include/helper.h (ray_aabb_intersection K=1, ray_aabb_intersection_v2):
 * centers[K,3], sizes[K,3] (full extents) -> bounds[B,K,2] = (near,far) or (-1,-1). */
int snrf_ray_aabb(const float* rays_o, const float* rays_d, const float* centers, const float* sizes,
                  float* bounds, int B, int K, void* stream);
/* cuda/include/helper.h (sample_points_grid): occupied[2^lx,2^ly,2^lz] bytes, log2dim[3] i32
 * (device) -> z_vals, dists [B,S]; counts[B] i32 optional (occupied segments per ray). */
int snrf_sample_grid(const float* rays_o, const float* rays_d, float* z_vals, float* dists,
                     const float* corner, const float* size, const unsigned char* occupied,
                     const int* log2dim, int* counts, int B, int S, void* stream);
/* cuda/include/sample.h (background_sampling_cuda) */
int snrf_bg_sampling(const float* starts, const float* bg_depth, float* z_vals, int B, int S,
                     float sample_range, void* stream);
/* cuda/include/sample.h (sample_insideout_block); *miss_flag set to 1 if any ray misses the box */
int snrf_sample_insideout(const float* rays_o, const float* rays_d, int S, int Sbg, const float* center,
                          const float* size, float far, float* z_vals, float* z_vals_bg, int* miss_flag,
                          int B, void* stream);

/* ---- front-to-back compositing ------------------------------------------------ */
/* hashgrid/__init__.py:344-366,564-596 (HashGrid.cal_integrate_weight / accumulate / tail of
 * render_batch_rays), which the reference runs as a chain of torch ops.
 * Per-sample heads are addressed base + n*stride (strides in floats): sigma[R*S], tint,
 * diffuse, specular [R*S,3]; z_vals, dists [R,S]; rays_d [R,3] (|d| scales the step).
 * -> weights[R,S], trans[R,S] (transmittance before each sample; may be NULL in fwd),
 *    out[R,16] = depth, tint3, diffuse3, specular3 (= sum w tint*spec), l2_3 (= sum w spec^2),
 *    T_left, 2 pad.  infinity != 0: last step is 1e10. */
int snrf_composite_fwd(const float* sigma, const float* tint, const float* diffuse, const float* specular,
                       int s_sigma, int s_tint, int s_diffuse, int s_specular,
                       const float* z_vals, const float* dists, const float* rays_d, int R, int S,
                       int infinity, float* weights, float* trans, float* out, void* stream);
/* backward of the above: g_out[R,16] (same row layout; l2 columns act on specular only, the
 * weights inside l2 are detached as in the reference), g_weights[R,S] optional.  WRITES the
 * per-sample head gradients (strides gs_*) and grad_rays_d[R,3] (optional). */
int snrf_composite_bwd(const float* sigma, const float* tint, const float* diffuse, const float* specular,
                       int s_sigma, int s_tint, int s_diffuse, int s_specular,
                       const float* z_vals, const float* dists, const float* rays_d, const float* trans,
                       const float* g_out, const float* g_weights, int R, int S, int infinity,
                       float* g_sigma, float* g_tint, float* g_diffuse, float* g_specular,
                       int gs_sigma, int gs_tint, int gs_diffuse, int gs_specular,
                       float* grad_rays_d, void* stream);

/* ---- decoder MLP on the tensor cores (tcgen05 / TMEM) ------------------------- */
/* Self-test of the tensor-core conventions (no reference counterpart): X[128,64], W[64,64],
 * G[128,64] f32, rounded to bf16 inside -> Y = X W^T [128,64], DX = G W [128,64],
 * DW = 2 G^T X [64,64], DWo = G^T X[:,32:48] [64,16], YS = X[:,32:64] W[0:16,0:32]^T [128,16] (f32). */
int snrf_umma_selftest(const float* X, const float* W, const float* G, float* Y, float* DX, float* DW,
                       float* DWo, float* YS, void* stream);

/* network.py:151-190 (ShallowMLP.forward): feats[N,32] f32 (hash-encode output), mask32[32] f32
 * (level mask, NULL = ones), rays_d[R,3] (sample n belongs to ray n / S; normalised inside with the
 * reference's +1e-8), params = HOST array of 16 DEVICE pointers in network.ShallowMLP state_dict
 * order (weight, bias per Linear: Spatial_MLP.mlp.0 [64,32], .mlp.2 [64,64], sigma_layer [1,32],
 * diffuse_layer [3,32], tint_layer [3,32], Directional_MLP.mlp.0 [64,48], .2 [64,64], .4 [3,64]).
 * -> heads[N,10] f32 = (sigma, tint3, diffuse3, specular3).  bf16(x3) operands, f32 accumulation. */
int snrf_decoder_fwd(const float* feats, const float* mask32, const float* rays_d, const float* const* params,
                     float* heads, int N, int S, void* stream);

/* Operand precision of the decoder GEMMs: 1 (default) = error-compensated bf16x3 split operands in
 * every forward GEMM (~fp32 accuracy), 0 = plain bf16 operands (fastest). */
void snrf_decoder_set_precision(int split);
/* Backward of snrf_decoder_fwd (autograd of network.py:151-190).  grad_heads[N,10] (column order
 * of heads) -> grad_feats[N,32] WRITTEN; grad_rays_d[R,3] ACCUMULATED (may be NULL; the view
 * direction enters through the SH encoding only); grad_params = HOST array of 16 DEVICE pointers,
 * shapes/order of params, ACCUMULATED.  The forward is recomputed per 128-sample tile. */
int snrf_decoder_bwd(const float* feats, const float* mask32, const float* rays_d, const float* const* params,
                     const float* grad_heads, float* grad_feats, float* grad_rays_d, float* const* grad_params,
                     int N, int S, void* stream);

/* ---- sparse Adam -------------------------------------------------------------- */
/* cuda/include/adam.h (adam_step_cuda: half_state=0, adam_step_cuda_fp16: half_state=1; kernels
 * cuda/adam_kernel.cu:23-69, 97-144).  Element (k,d), k<rows, d<dim, lives at k*row_stride+d
 * (the reference hard-codes row_stride 8).  Elements whose gradient is exactly 0 are skipped.
 * exp_avg / exp_avg_sq are f32 (half_state=0) or f16 scaled by 128 / 128^2 (half_state=1).
 * `step` is the 1-based step used in the bias corrections (the reference passes step+1).
 * zero_grad != 0 clears every consumed gradient in the same pass (extension). */
int snrf_adam_step(float* params, float* grads, void* exp_avg, void* exp_avg_sq,
                   long long rows, int dim, int row_stride, int half_state,
                   float lr, float beta1, float beta2, float eps, int step, int zero_grad,
                   void* stream);

/* ---- ray <-> proxy mesh (fastMesh) --------------------------------------------- */
/* fastMesh/include/fastMesh.h:11-57 (class fastMesh).  The handle owns one device allocation on
 * the current device; queries may run on any stream of that device. */
int snrf_mesh_create(const char* ply_path, void** handle);                 /* build(path): fastMesh.h:22-26 */
int snrf_mesh_create_from_arrays(const float* verts_host, int nv, const int* faces_host, int nf, void** handle);
int snrf_mesh_destroy(void* handle);                                        /* destroy(): fastMesh.h:52-55 */
int snrf_mesh_bounds(void* handle, float* bound6_host);                     /* getSceneBound(): fastMesh.h:28-38 */
int snrf_mesh_stats(void* handle, long long* stats3_host);                  /* occupied cells, list entries, faces */
/* fisrtHit (sic) fastMesh_kernel.cu:230-329: z_depth[B] = nearest t>0 within the first cell that has a
 * hit, 0 = miss; hit_face[B] i32 optional (extension): that face's index or -1. */
int snrf_mesh_first_hit(void* handle, const float* rays_o, const float* rays_d, float* z_depth, int* hit_face,
                        int B, void* stream);
/* firstEnter fastMesh_kernel.cu:125-227 */
int snrf_mesh_first_enter(void* handle, const float* rays_o, const float* rays_d, float* z_depth, int B, void* stream);
/* sample_points fastMesh_kernel.cu:23-122: t_start[B] (-1 = leave the row untouched), z_vals[B,S] */
int snrf_mesh_sample(void* handle, const float* rays_o, const float* rays_d, const float* t_start, float* z_vals,
                     int B, int S, void* stream);

/* ---- mesh ingest (host) ------------------------------------------------------- */
/* cuda/include/voxelize.h:12-119 (voxelize_mesh): ALL pointers are host pointers.
 * log2dim[3], corner[3], size[3]; vis / outside: bytes [2^lx * 2^ly * 2^lz]. */
int snrf_voxelize_mesh_host(const int* log2dim, const float* corner, const float* size,
                            const char* model_path, unsigned char* vis, int init_out,
                            unsigned char* outside);

#ifdef __cplusplus
}
#endif
#endif /* SCANERF_B200_H */
