/* pea_b200.h -- C ABI of libpea_b200.so (sm_100a only).
 *
 * The reference (OPPO-Mente-Lab/PEA-Diffusion) is 100 % Python and defines no FFI of its own:
 * its hot path is reached through two Python protocols (the adapter nn.Module and diffusers' attention
 * processor).  This header is therefore the boundary the *rebuild* defines; each entry point cites the reference
 * statement(s) whose arithmetic it replaces.  All file:line citations are relative to the reference tree.
 *
 * Conventions (every function):
 *   - extern "C", plain C types.  Pointers named d_* / inside descriptors are DEVICE pointers
 *     (tensor.data_ptr()); descriptor structs themselves live in HOST memory.
 *   - `stream` is a cudaStream_t / CUstream passed as void* (torch.cuda.current_stream().cuda_stream).
 *   - Stream-ordered, no host synchronisation, never allocates or frees caller-visible memory: outputs and
 *     workspaces are allocated by the caller; `*_workspace_bytes` report sizes.
 *   - Returns 0 on success, a negative PEA_E* code otherwise; pea_last_error() gives a thread-local message.
 *   - There is no CPU fallback: without an sm_100 device every compute entry point fails with PEA_ENODEV / a
 *     CUDA launch error.
 */
#ifndef PEA_B200_H_
#define PEA_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PEA_B200_VERSION 201 /* 0.2.1: flags on the row helpers; 0.2.0: flags on the gemm / attn2 descriptors, inv_prescale on pea_kd_scale_grads */

enum { PEA_OK = 0, PEA_EINVAL = -1, PEA_ECUDA = -2, PEA_ENODEV = -3, PEA_EUNSUPPORTED = -4 };
/* `flags` of pea_gemm_desc / pea_attn2_desc.
 * PEA_LAUNCH_PDL: launch the kernel as a PROGRAMMATIC DEPENDENT of the kernel that precedes it in the stream
 * (cudaLaunchAttributeProgrammaticStreamSerialization): its prologue -- barrier init, TMEM allocation, descriptor
 * prefetch -- overlaps that kernel's tail, and every global-memory access is ordered after that kernel's completion by
 * griddepcontrol.wait.  Every tcgen05 kernel of the library also issues griddepcontrol.launch_dependents right after
 * its own wait, so a chain  to_q GEMM -> attn2 core -> to_out GEMM  (diffusers AttnProcessor2_0's q = to_q(h);
 * o = SDPA(...); out = to_out[0](o)) runs without the launch gaps and prologues between its links.  Without the flag
 * both instructions are no-ops and the launch is an ordinary stream-ordered one. */
enum { PEA_LAUNCH_PDL = 1 };
typedef enum { PEA_F32 = 0, PEA_F16 = 1, PEA_BF16 = 2 } pea_dtype;

int pea_version(void);
const char* pea_last_error(void);
/* Number of kernels this library has launched in the calling process (bench.py's `gpu_launches`). */
long long pea_launch_count(void);

/* ----------------------------------------------------------------------------------------------------------
 * KD loss: base + output-level + feature-level masked MSE and its gradient, one pass.
 * Replaces train_sdxl_zh.py:400-405 (L0), :416-421 (L1), :423-441 (L2) and train_sd_zh.py:218-272.
 *
 * A tap contributes   term = group_scale * mean_b( mean_chw( (s - t)^2 * w_b ) )   [+ the same with t2, w2]
 * and, when d_grad != NULL,  d_grad = grad_scale * d(term)/ds  (teacher tensors get no gradient).
 * `out` (device, 4 + 2*ntaps floats): [0..2] = L0, L1, L2 as the reference logs them (L2 unweighted),
 * [3] = L0 + L1 + feature_weight * L2, [4+i] = tap i's unweighted term, [4+ntaps+i] = 1.0 if tap i was finite
 * and counted, 0.0 if it was skipped (train_sd_zh.py:246-268 semantics, only when skip_nonfinite != 0).
 * -------------------------------------------------------------------------------------------------------- */
#define PEA_KD_MAX_TAPS 16
typedef struct {
  const void* d_student;   /* [B, numel_per_sample], contiguous */
  const void* d_teacher;   /* same shape */
  const void* d_teacher2;  /* optional second target for the same student tensor (noise_pred: eps and eps_T) */
  void* d_grad;            /* optional, same shape/dtype as student */
  const float* d_weight;   /* [B] per-sample weight for (s - t)^2, e.g. zh_b or 1 - zh_b */
  const float* d_weight2;  /* [B] per-sample weight for (s - t2)^2 (with d_teacher2) */
  long long numel_per_sample;
  int group;               /* 0: L0 (base), 1: L1 (output KD), 2: L2 (feature KD) -- for the (s,t) pair */
  int group2;              /* group of the (s,t2) pair */
} pea_kd_tap;

size_t pea_kd_workspace_bytes(const pea_kd_tap* taps, int ntaps, int B);
int pea_kd_mse_fwd_bwd(const pea_kd_tap* taps, int ntaps, int B, pea_dtype dtype, float feature_weight,
                       float grad_scale, int skip_nonfinite, void* d_workspace, float* d_out, void* stream);
/* Backward fix-up: d_grad *= (*d_upstream) * inv_prescale * counted_flag; exits immediately on device when the factor
 * is 1.  `inv_prescale` = 1 / the `grad_scale` the fused pass was given.  Under fp16 loss scaling (the reference trains
 * with --precision 16 + DeepSpeed's loss scaler, train_sdxl_zh.sh:89) the caller passes its loss scale as `grad_scale`
 * so that the gradient is scaled BEFORE it is rounded to fp16 (unscaled feature-KD coefficients, ~1e-8, flush to
 * zero in fp16); autograd's upstream gradient then equals the loss scale and this pass has nothing to do. */
int pea_kd_scale_grads(const pea_kd_tap* taps, int ntaps, int B, pea_dtype dtype, const float* d_upstream,
                       float inv_prescale, const float* d_out, void* stream);

/* ----------------------------------------------------------------------------------------------------------
 * attn2 core: softmax(q k^T * scale) v per head, keys resident in shared memory.
 * Replaces diffusers AttnProcessor2_0's F.scaled_dot_product_attention call as reached from
 * train_sdxl_zh.py:397,415 / train_sd_zh.py:215,231 / tests/test_sdxl_zh.py:384-391.
 *   q, o, dq, do : [B, N, H*d] rows of stride ld_q elements (head h occupies columns [h*d, (h+1)*d))
 *   k, v         : [B, L, H*d] rows of stride ld_kv elements, batch stride bs_kv elements (K|V may interleave)
 *   dk, dv       : fp32, same indexing as k, v with ld_dkv / bs_dkv; ACCUMULATED into (caller zero-fills)
 * L <= 80 (77 for CLIP; 52/64/32 for the other encoders), d in {64} (+{40,48,80,160}: see DESIGN.md).
 * Alignment: q / k / v / o / do / dq 16-byte aligned with ld_* and bs_kv multiples of 8 elements (128-bit row accesses,
 * TMA boxes); dk / dv 8-byte aligned with even ld_dkv / bs_dkv.  d = 64 runs the tcgen05/TMA kernels, the other head
 * sizes the warp-MMA kernels.  Violations return PEA_EINVAL with the reason in pea_last_error().
 * -------------------------------------------------------------------------------------------------------- */
typedef struct {
  int B, N, L, H, d;
  float scale;
  pea_dtype dtype;         /* PEA_BF16 or PEA_F16 */
  const void* d_q; long long ld_q;
  const void* d_k; const void* d_v; long long ld_kv, bs_kv;
  void* d_o; long long ld_o;
  /* backward only */
  const void* d_do; long long ld_do;
  void* d_dq; long long ld_dq;
  float* d_dk; float* d_dv; long long ld_dkv, bs_dkv;
  unsigned flags;          /* PEA_LAUNCH_* */
} pea_attn2_desc;

int pea_attn2_fwd(const pea_attn2_desc* desc, void* stream);
int pea_attn2_bwd(const pea_attn2_desc* desc, void* stream);

/* ----------------------------------------------------------------------------------------------------------
 * Tensor-core GEMM (tcgen05 + TMEM accumulators, TMA-fed), the building block of the adapter
 * (train_sdxl_zh.py:58-67 forward and its autograd backward) and of the attn2 K/V projections
 * (AttnProcessor2_0: attn.to_k / attn.to_v) and their context gradient.
 *
 *   C[m, n] = sum_k A(m, k) * B(n, k)                      m < M, n < N, k < K
 *   A stored [M, K] row-major (a_kmajor = 1, lda = row stride) or [K, M] row-major (a_kmajor = 0);
 *   B stored [N, K] row-major (b_kmajor = 1)               or [K, N] row-major (b_kmajor = 0).
 * Operands are 16-bit planes; with nsplit_a = Sa > 1 the true operand is A[0] + A[1] + ... (a bf16/fp16 split of an
 * fp32 value, see pea_split_planes) and likewise for B; the kernel accumulates the cross products A[i] * B[j] with
 * i + j < max(Sa, Sb), i.e. the most significant ones -- the fp32-accurate mode (Sa = Sb = 2: ~2^-16 relative).
 * Epilogue:  v = alpha * acc [+ bias[n]] [* gelu'(aux[m, n])] [+ rowadd[(m / rowadd_group), n] * rowadd_scale]
 *            out0[m, n] (=|+=) v          (raw, any dtype)                        -- optional
 *            out1 planes   = split_S1( act ? gelu(v) : v )                         -- optional
 * split_k > 1 partitions the K loop over that many CTAs per tile which add their partial sums atomically; it needs
 * out0 fp32 with out0_accumulate = 1 (the caller zero-fills or really accumulates) and no out1 / bias / aux.
 * Leading dimensions are in elements and must be multiples of 8; base pointers 16-byte aligned.
 * -------------------------------------------------------------------------------------------------------- */
typedef struct {
  int M, N, K;
  pea_dtype in_dtype;       /* PEA_BF16 or PEA_F16 */
  int nsplit_a, nsplit_b;   /* 1..3 planes per operand */
  const void* d_A[3]; int a_kmajor; long long lda;
  const void* d_B[3]; int b_kmajor; long long ldb;
  float alpha;
  const float* d_bias;
  const void* d_aux; pea_dtype aux_dtype; long long ldaux;
  const float* d_rowadd; int rowadd_group; float rowadd_scale;
  int act;                  /* 0: identity, 1: exact-erf GELU applied to out1 */
  void* d_out0; pea_dtype out0_dtype; long long ldo0; int out0_accumulate;
  void* d_out1[3]; int out1_nsplit; pea_dtype out1_dtype; long long ldo1;
  int split_k;              /* 0 or 1: none */
  unsigned flags;           /* PEA_LAUNCH_* */
} pea_gemm_desc;

int pea_gemm(const pea_gemm_desc* desc, void* stream);

/* ----------------------------------------------------------------------------------------------------------
 * Row-wise helpers of the adapter (LayerNorm of train_sdxl_zh.py:47,60; torch.mean(x, 1) of :66; the
 * bias gradient; operand splitting) and the CFG-dropout select of train_sdxl_zh.py:392-395.  `flags`: PEA_LAUNCH_*
 * (the adapter chains its launches with PEA_LAUNCH_PDL; every kernel orders its accesses with griddepcontrol.wait).
 * -------------------------------------------------------------------------------------------------------- */
int pea_layernorm_fwd(const void* d_x, pea_dtype x_dtype, const float* d_gamma, const float* d_beta, int M, int D,
                      float eps, void* const d_y_planes[3], int nsplit, pea_dtype plane_dtype, float* d_mean,
                      float* d_rstd, unsigned flags, void* stream);
/* dgamma/dbeta are ACCUMULATED (+=) so two calls (cond + uncond) can share them; dx optional. */
int pea_layernorm_bwd(const void* d_x, pea_dtype x_dtype, const float* d_dy, const float* d_gamma,
                      const float* d_mean, const float* d_rstd, int M, int D, float* d_dgamma, float* d_dbeta,
                      float* d_dx, unsigned flags, void* stream);
/* planes = split_S(x): x fp32/bf16/fp16 [rows, cols] with row stride ld -> S dense planes of plane_dtype. */
int pea_split_planes(const void* d_x, pea_dtype x_dtype, long long rows, long long cols, long long ld,
                     void* const d_planes[3], int nsplit, pea_dtype plane_dtype, unsigned flags, void* stream);
/* The same for `count` tensors of one dtype in ONE launch when each is dense fp32 with 8 | rows*cols (the adapter's weight
 * matrices, re-split every step: train_sdxl_zh.py:58-67 under fp16/bf16 autocast); otherwise one launch per tensor. */
typedef struct pea_split_item {
  const void* d_x; long long rows, cols, ld;
  void* d_planes[3];
} pea_split_item;
int pea_split_planes_multi(const pea_split_item* items, int count, pea_dtype x_dtype, int nsplit, pea_dtype plane_dtype,
                           unsigned flags, void* stream);
/* out[b, n] = mean_l (x[b*L + l, n] (+ res[b*L + l, n])), fp32 in/out (x1 of train_sdxl_zh.py:64-66). */
int pea_mean_tokens(const float* d_x, const void* d_res, pea_dtype res_dtype, int B, int L, int D, float* d_out,
                    unsigned flags, void* stream);
/* out[n] += sum_m x[m, n]  (bias gradient). */
int pea_colsum_accum(const void* d_x, pea_dtype x_dtype, int M, int N, long long ld, float* d_out, unsigned flags,
                     void* stream);
/* out[b] = mask[b] ? a[b] : c[b] over rows of `row_elems` elements (CFG dropout), and its backward routing. */
int pea_cfg_select(const unsigned char* d_mask, const void* d_uncond, const void* d_cond, void* d_out, int B,
                   long long row_elems, pea_dtype dtype, void* stream);

/* ----------------------------------------------------------------------------------------------------------
 * Input side of the step: offset noise + DDPM add_noise + cast in one pass.
 * Replaces train_sdxl_zh.py:311-323 (train_sd_zh.py:157-169):
 *     noise  = noise_raw + noise_offset * offs[b, c]                     (offs ~ N(0,1) of shape [B, C, 1, 1])
 *     noisy  = sqrt(acp[t_b]) * latents + sqrt(1 - acp[t_b]) * noise     (DDPMScheduler.add_noise)
 * latents [B, C, HW] any float dtype; noise_raw / offs fp32; acp = alphas_cumprod fp32 [T]; t int64 [B];
 * outputs d_noise, d_noisy in out_dtype (the UNet dtype).
 * -------------------------------------------------------------------------------------------------------- */
int pea_add_noise(const void* d_latents, pea_dtype lat_dtype, const float* d_noise_raw, const float* d_offs, float noise_offset,
                  const float* d_acp, int T, const long long* d_t, int B, int C, long long HW, void* d_noise, void* d_noisy,
                  pea_dtype out_dtype, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* PEA_B200_H_ */
