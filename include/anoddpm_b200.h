/*
 * include/anoddpm_b200.h -- C ABI of libanoddpm_b200.so (sm_100a kernels for the AnoDDPM hot path).
 *
 * The reference (Julian-Wyatt/AnoDDPM) is pure Python and has no FFI of its own; its drop-in
 * boundary is the Python class API (GaussianDiffusionModel / UNetModel / Simplex_CLASS).  This
 * header is the boundary BELOW that API: each entry point names the reference code it replaces
 * (paths relative to the reference tree).  INTEGRATION.md shows the ctypes binding.
 *
 * Conventions
 *   - every function returns 0 on success, <0 on error (ANODDPM_ERR_*); the message is in
 *     anoddpm_last_error() (thread-local).  Nothing throws, exits or allocates persistently.
 *   - all tensor arguments are DEVICE pointers owned by the caller unless marked [host];
 *     `stream` is a cudaStream_t passed as void*.  Entry points are asynchronous, re-entrant
 *     and CUDA-graph capturable (no allocation, no synchronisation, no host reads of device data).
 *   - activations inside the UNet are NHWC fp32 ("pixels x channels"); images at the API edge
 *     are NCHW fp32 like the reference.
 */
#ifndef ANODDPM_B200_H
#define ANODDPM_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define ANODDPM_ABI_VERSION 1
/* Per-channel statistics buffers hold ANODDPM_STAT_REPLICAS partial copies, [replica][B][C][2] fp64: producers that
 * run many CTAs on the same image (the conv epilogue) spread their atomics over the replicas (contention / 8); consumers
 * (anoddpm_gn_apply) add the replicas up.  anoddpm_gn_stats writes replica 0. */
#define ANODDPM_STAT_REPLICAS 8

int anoddpm_abi_version(void);
const char* anoddpm_last_error(void);
/* name of the tensor-core conv path compiled in ("tcgen05-tf32") -- for bench/DESIGN bookkeeping */
const char* anoddpm_build_info(void);

/* ------------------------------------------------------------------ simplex noise ------------- */

/* [host] Permutation tables of one seed.  tables[0..255] = perm, tables[256..511] = perm_grad_index3.
 * Replaces simplex.py:166-192 (_init / overflow); integer, bit-exact. */
int anoddpm_simplex_init(int64_t seed, int16_t* tables /* [host] 512 */);
int anoddpm_simplex_init_batch(const int64_t* seeds /* [host] n */, int n, int16_t* tables /* [host] n*512 */);

/* Fractal simplex noise written straight into an fp32 (B,C,H,W) tensor:
 *   out[b,c,y,x] = fp32( sum_o persistence^o * noise3(x*2^o/f, y*2^o/f, t[b]*2^o/f ; tables[b*C+c]) )   (fp64 inside)
 * = GaussianDiffusion.py:96-137 (generate_simplex_noise) -> simplex.py:75-93 (rand_3d_fixed_T_octaves)
 *   -> :833-840 (_noise3a) -> :321-830 (_noise3), evaluated per sample (batched semantic of DESIGN.md).
 * shared_seed != 0: tables holds C table sets (one per channel) shared by all samples. */
int anoddpm_simplex_noise(float* out, int B, int C, int H, int W, const int16_t* tables, int shared_seed,
                          const int64_t* t /* B */, int octaves, double persistence, double frequency, void* stream);

/* Same, for a sampling loop captured in a CUDA graph: tables_all holds the table sets of EVERY step back to back
 * ((B*C)*512 int16 per step, or C*512 when shared_seed) and the kernel uses set number *step (a device scalar the graph
 * increments), so a replay needs no host work.  GaussianDiffusion.py:351-357 calls the noise generator once per step. */
int anoddpm_simplex_noise_steps(float* out, int B, int C, int H, int W, const int16_t* tables_all, const int64_t* step,
                                int shared_seed, const int64_t* t /* B */, int octaves, double persistence,
                                double frequency, void* stream);

/* Volume / slice fields for the Simplex_CLASS numpy API: out[d,y,x] (fp64 if out_f64 else fp32),
 * z coordinate of slice d = z[d] (NULL -> d).  simplex.py:37-54 (rand_3d_octaves), :75-93. */
int anoddpm_simplex_octaves(void* out, int out_f64, int D, int H, int W, const int16_t* tables /* 512 */,
                            const double* z /* D or NULL */, int octaves, double persistence, double frequency,
                            void* stream);

/* out[z][y][x] = noise3(X[x], Y[y], Z[z])   -- simplex.py:833-840 (_noise3a / noise3array) */
int anoddpm_simplex_grid3(double* out, const double* X, int nx, const double* Y, int ny, const double* Z, int nz,
                          const int16_t* tables, void* stream);
/* out[i][j] = noise2(X[j], Y[i])            -- simplex.py:211-318 (_noise2 / _noise2a, square case) */
int anoddpm_simplex_grid2(double* out, const double* X, int nx, const double* Y, int ny, const int16_t* tables,
                          void* stream);

/* ------------------------------------------------------------------ diffusion elementwise ----- */

/* out = coef_a[t_b]*x + coef_b[t_b]*noise, per sample b.  GaussianDiffusion.py:361-371 (sample_q, with
 * sqrt_alphas_cumprod / sqrt_one_minus_alphas_cumprod) and :373-382 (sample_q_gradual, sqrt_alphas / sqrt_betas).
 * coef_* are fp32 tables of length T (= extract(), :32-36, done once). */
int anoddpm_q_sample(float* out, const float* x, const float* noise, const int64_t* t, const float* coef_a,
                     const float* coef_b, int B, int64_t per_sample, void* stream);

/* One reverse step given eps = model(x_t, t):  GaussianDiffusion.py:298-318 (sample_p) with :269-296
 * (p_mean_variance), :228-230, :253-267 fused.  tables = 5 rows of length T:
 *   [sqrt_recip_alphas_cumprod, sqrt_recipm1_alphas_cumprod, posterior_mean_coef1, posterior_mean_coef2,
 *    sigma0 = (t!=0)*exp(0.5*log(model_var))].   pred_x0 may be NULL. */
int anoddpm_p_sample(float* sample, float* pred_x0, const float* x_t, const float* eps, const float* noise,
                     const int64_t* t, const float* tables, int T, int B, int64_t per_sample, void* stream);

/* The same reverse step with its simplex noise generated in place (single-channel images): per pixel the `octaves`-octave
 * noise3 value of (x, y, t_b) under seed table tables_all[*step][b] (the layout of anoddpm_simplex_noise_steps; step may be
 * NULL = 0), cast to fp32, then the posterior arithmetic of anoddpm_p_sample -- bit-identical to anoddpm_simplex_noise_steps
 * followed by anoddpm_p_sample, without writing or reading the noise tensor (12 B/px instead of 16 + 4).
 * GaussianDiffusion.py:96-137 + 298-318. */
int anoddpm_p_sample_simplex(float* sample, float* pred_x0, const float* x_t, const float* eps,
                             const int16_t* tables_all /* [steps][B][512] */, const int64_t* step /* [device], may be NULL */,
                             const int64_t* t, const float* tables, int T, int B, int H, int W, int octaves,
                             double persistence, double frequency, void* stream);

/* Detection epilogue in one pass (reference detection.py:229-250 + evaluation.py:26-76 do ~10 torch/numpy passes with
 * D2H copies): counts[b] += {n11, n10, n01, n00, sum((image-recon)^2)} and counts[b][5] = max(counts[b][5], max(image)),
 * where "1/0" pairs are (mask == 1, (image-recon)^2 > threshold).  Caller initialises counts to {0,0,0,0,0,-inf}. */
int anoddpm_detection_counts(double* counts /* B x 6 */, const float* image, const float* recon, const float* mask,
                             float threshold, int B, int64_t per_sample, void* stream);

/* Variational-bound term of one (x_0, x_t, eps, t) batch (GaussianDiffusion.py:384-397 calc_vlb_xt with :43-93):
 * acc[b] += {sum over the sample of KL(q(x_{t-1}|x_t,x_0) || p), sum of the discretised-Gaussian decoder NLL} (fp64, caller
 * zeroes); vlb[b] = (t[b] == 0 ? acc[b][1] : acc[b][0]) / (per_sample * ln 2).  tables = the (sqrt_recip_alphas_cumprod,
 * sqrt_recipm1_alphas_cumprod, posterior_mean_coef1, posterior_mean_coef2) rows as for anoddpm_p_sample; logvars =
 * [2][T] fp32: posterior_log_variance_clipped, log(model variance).  pred_x0 (optional) receives clamp(x_0 estimate). */
int anoddpm_vlb_terms(double* acc /* B x 2 */, float* pred_x0, const float* x_0, const float* x_t, const float* eps,
                      const int64_t* t, const float* tables, const float* logvars, int T, int B, int64_t per_sample,
                      void* stream);

/* ------------------------------------------------------------------ UNet denoiser ops --------- */

/* emb[b, :] = cat(sin(t_b*w), cos(t_b*w)), w = freq[0..dim/2) = exp(-i*ln(1e4)/half) (built once by the caller
 * exactly as UNet.py:50-55 builds it, so the argument t*w is bit-identical to the reference).  UNet.py:38-57. */
int anoddpm_timestep_embedding(float* emb /* B x dim */, const int64_t* t, const float* freq /* dim/2 */, int B,
                               int dim, void* stream);

/* out[b][:] = table[idx[b]][:] for rows of S floats (S % 4 == 0): the per-timestep embedding table of the sampling loop --
 * the timestep MLP and every ResBlock's embedding projection (UNet.py:206-207, 396-397) depend on t only, so the loop computes
 * them once for all T timesteps and gathers a row per sample per step.  An index outside [0, T) fills the row with NaN. */
int anoddpm_gather_rows(float* out, const float* table, const int64_t* idx, int B, int S, int T, void* stream);

/* y[b, n] = g(sum_k f(x[b,k]) * w[n,k] + bias[n]); act bit 0: f = SiLU (else identity), bit 1: g = SiLU.
 * UNet.py:271-276, :185-188 (nn.Linear between nn.SiLU).  ldx / ldy: row strides of x / y (elements). */
int anoddpm_linear(float* y, int ldy, const float* x, int ldx, const float* w, const float* bias, int B, int N, int K,
                   int act, void* stream);

/* Per-channel statistics of an NHWC tensor: stats[b][c] += (sum, sum of squares) over the HW pixels, fp64 (caller
 * zeroes).  Per-channel sums are grouping-agnostic: a GroupNorm over this tensor, or over its concatenation with another
 * one (torch.cat, UNet.py:402), adds up the channels of each group in anoddpm_gn_apply.  The tensor-core conv
 * (anoddpm_conv impl 3) can emit the same numbers for its output from its epilogue (gn_stats field).  UNet.py:409-411. */
int anoddpm_gn_stats(double* stats /* REPLICAS x B x C x 2 (replica 0 written) */, const float* x, int C, int B, int HW,
                     void* stream);

/* out = resample(act(GroupNorm(cat(x1, x2)))) on NHWC: act = SiLU if silu; mode 0 same, 1 avg-pool 2x2, 2 nearest x2
 * (UNet.py:202-210: GN -> SiLU -> h_upd, in that order).  stats1 / stats2: per-channel statistics of x1 / x2
 * ([REPLICAS][B][C1][2], [REPLICAS][B][C2][2]); x2, stats2 may be NULL (C2 = 0).
 * round_tf32 = output mode: 0 fp32; 1 fp32 rounded to TF32 (rna) for a kind::tf32 conv; 2 IEEE fp16 (rn, saturating;
 * `out` / `raw_out` then point to half elements) for a kind::f16 conv.
 * raw_out (optional): mode 0: also writes the rounded, channel-concatenated copy of the INPUT, which the fused 1x1
 * skip convolution of the same ResBlock reads as its K-major operand (one pass over x for both); mode 1 (single
 * source): also writes the fp32 avg-pooled INPUT, the residual of a downsampling ResBlock (UNet.py:178,216). */
int anoddpm_gn_apply(void* out, void* raw_out, const float* x1, int C1, const float* x2, int C2,
                     const double* stats1, const double* stats2, const float* gamma, const float* beta, int B, int H,
                     int W, int groups, float eps, int silu, int mode, int round_tf32, void* stream);

/* The same GroupNorm (+SiLU) for an fp16 input produced by anoddpm_conv with out_f16 (same resolution, one source,
 * C % 8 == 0): y = act((x + shift - mean) * rstd * gamma + beta) with shift[b][c] = shift_c[c] + shift_bc[b][c]
 * (either may be NULL) -- the conv bias and timestep-embedding row the producer left out; stats are those of x + shift. */
int anoddpm_gn_apply_f16in(void* out, const void* x_f16, int C, const double* stats, const float* gamma,
                           const float* beta, const float* shift_c, const float* shift_bc, int shift_stride, int B, int H,
                           int W, int groups, float eps, int silu, int out_mode, void* stream);

/* Per-(sample, channel) affine form of GroupNorm for the fused operand transform of anoddpm_conv (gn_table):
 * table[b][c] = (scale, shift) with scale = rstd[b, g(c)] * gamma[c] and shift = beta[c] + (off[b][c] - mean[b, g(c)]) * scale,
 * so that GroupNorm(x + off)[b, c] = x * scale + shift.  Statistics as in anoddpm_gn_apply (per-channel fp64 sums in
 * REPLICAS copies, of one tensor or of the concatenation of two); off = shift_c[c] + shift_bc[b][c] (either may be NULL) is
 * the bias + timestep-embedding row an fp16 conv output left out (see out_f16).  HW = pixels behind each statistic. */
int anoddpm_gn_table(float* table /* [B][C1+C2][2] */, const double* stats1, int C1, const double* stats2, int C2,
                     const float* gamma, const float* beta, const float* shift_c, const float* shift_bc, int shift_stride,
                     int B, int HW, int groups, float eps, void* stream);

/* Convolution as implicit GEMM over NHWC pixels.  out[b,y,x,:] =
 *     conv_k(a)[..] (k = ksize 3 pad 1, or 1)  +  r1 . W_r1  +  r2 . W_r2      (1x1 "skip" sources, optional)
 *   + bias  +  emb[b, :] (optional)  +  residual (optional)
 * w is packed [Cout][Ktot] with Ktot = ksize^2*Ca + Cr1 + Cr2 ordered (tap-major, channel-minor), then r1, r2.
 * Replaces nn.Conv2d 3x3/1x1 + skip_connection + the adds of ResBlock.forward (UNet.py:202-217) and the
 * Conv1d 1x1 qkv/proj of AttentionBlock (UNet.py:115-125).
 * res_mode: 0 none, 1 same resolution, 2 res is 2H x 2W and is avg-pooled 2x2, 3 res is H/2 x W/2 nearest-upsampled
 * (x_upd of UNet.py:206-207).  impl: 0 = exact fp32 CUDA-core kernel; 1 = tcgen05 kind::tf32, one CTA per 128x128 tile;
 * 2 = tcgen05 kind::tf32 cta_group::2, persistent CTA pairs (256 x BN tiles, double-buffered TMEM);
 * 3 = as 2 with halo reuse (16x16 spatial regions, vertical taps share one TMA box, two M-subtiles share each weight
 *     tile) -- 3x3 convs with H, W % 16 == 0; cuts L2->SM operand traffic ~2.5x.  The fast path. */
typedef struct anoddpm_conv_desc {
    void* out;         /* fp32 NHWC; fp16 when out_f16 */
    const void* a;     /* fp32, or fp16 when operand_f16 (same for r1, r2, w) */
    const void* r1;
    const void* r2;
    const void* w;
    const float* bias;
    const float* emb;
    const float* res;
    int B, H, W;
    int Ca, Cr1, Cr2, Cout;
    int ksize;
    int emb_stride;
    int res_mode;
    int impl;
    int operand_f16;    /* impl 2/3 only: a, r1, r2, w hold IEEE fp16 (tcgen05 kind::f16, fp32 accumulate): the same 10-bit
                           mantissa as TF32 at twice the tensor rate and half the operand bytes; channels % 64 == 0 */
    double* gn_stats;   /* optional, impl 3 (or impl 2 with H*W % 32 == 0): += (sum, sum of squares) of the OUTPUT per (sample, channel), layout
                           [ANODDPM_STAT_REPLICAS][B][Cout][2] fp64 (sum over replicas = what anoddpm_gn_stats computes) */
    int out_f16;        /* impl 3 only, no residual: `out` holds fp16 and receives the convolution sum WITHOUT bias and
                           embedding (the half then carries no per-channel offset); gn_stats still describes
                           h = out + bias + emb, and anoddpm_gn_apply_f16in adds the offset back in fp32.  For an output
                           that is consumed by a GroupNorm only (conv1 of a ResBlock, UNet.py:190-200). */
    /* ---- impl 3, fp16 operands: GroupNorm + SiLU applied to the A operand inside the convolution (UNet.py:170-172,
     * 190-193: the GN -> SiLU that precedes every 3x3 conv), so no normalised copy of the activation is ever written ---- */
    const void* a2;         /* optional second main 3x3 source, channel-concatenated after `a` (torch.cat of the up path,
                               UNet.py:402); Ca2 channels, % 64 == 0 */
    int Ca2;
    const float* gn_table;  /* optional [B][Ca + Ca2][2] fp32 (scale, shift) from anoddpm_gn_table: `a` / `a2` hold the RAW
                               fp16 tensor(s) and the loader rewrites every landed element to SiLU(v * scale + shift) (fp32
                               math, rounded to nearest fp16) before the tensor cores read it; zero padding stays zero */
    void* out16;            /* optional fp16 copy of the fp32 output (same NHWC shape): the raw operand of the next
                               convolution with gn_table, and of a fused 1x1 skip convolution; not with out_f16 */
    /* ---- impl 2, small maps: split-K.  When the problem has fewer tiles than CTA pairs (8x8 maps), each tile's K loop is
     * cut into up to 8 ranges that run as separate CTA-pair jobs; partial accumulators meet in `workspace` and the last job
     * to arrive finishes the tile (sums in split order: deterministic).  The first anoddpm_conv_workspace_bytes() bytes
     * must be ZERO on first use (arrival counters; the kernel leaves them zero) and the buffer must not be shared by
     * convolutions running concurrently.  NULL / too small: no split. ---- */
    void* workspace;
    long long workspace_bytes;
} anoddpm_conv_desc;
int anoddpm_conv(const anoddpm_conv_desc* d /* [host] */, void* stream);
/* bytes of `workspace` that enable split-K for this problem (0 if it would not be split) */
long long anoddpm_conv_workspace_bytes(const anoddpm_conv_desc* d /* [host] */);
/* 1 if impl=1/2 (tcgen05 tiles: channels % 32, Cout % 16, power-of-two W) covers the problem, else 0. */
int anoddpm_conv_tc_supported(const anoddpm_conv_desc* d /* [host] */);
/* 1 if impl=3 covers the problem (ksize 3, H % 16 == W % 16 == 0, channels % 32 == 0, Cout % 16 == 0). */
int anoddpm_conv_tc3_supported(const anoddpm_conv_desc* d /* [host] */);

/* Stem: NCHW image (Cin small) -> NHWC features, 3x3 pad 1, w is OIHW.  UNet.py:279-281. */
int anoddpm_conv_stem(float* out_nhwc, const float* x_nchw, const float* w_oihw, const float* bias, int B, int Cin,
                      int H, int W, int Cout, void* stream);
/* Head: NHWC activated features -> NCHW image (Cout small), 3x3 pad 1, w packed [Cout][9*Cin] tap-major. UNet.py:384-388 */
int anoddpm_conv_head(float* out_nchw, const float* a_nhwc, const float* w_packed, const float* bias, int B, int Cin,
                      int H, int W, int Cout, void* stream);
/* The whole output head in one pass: out = conv3x3(SiLU(GroupNorm(x))) (UNet.py:384-388).  x is the raw last feature map
 * (NHWC fp32), stats its per-channel statistics ([REPLICAS][B][Cin][2]); the normalisation is applied while the tile is
 * staged, so no normalised copy is written.  Cin <= 512. */
int anoddpm_conv_head_gn(float* out_nchw, const float* x_nhwc, const double* stats, const float* gamma, const float* beta,
                         const float* w_packed, const float* bias, int B, int Cin, int H, int W, int Cout, int groups,
                         float eps, void* stream);

/* The single-output head in two streaming passes (Cout == 1 makes the 3x3 convolution separable by tap): pass 1 reads the raw
 * last feature map once, without halo, applies GroupNorm + SiLU from gn_table ([B][C] (scale, shift), anoddpm_gn_table) and
 * reduces the C channels of every pixel against the nine tap weight vectors into nine fp32 maps in `workspace`
 * (anoddpm_conv_head_sep_workspace_bytes, L2-sized); pass 2 adds the nine shifted maps + bias into out_nchw.  C == 128 or 256,
 * H*W % 64 == 0.  UNet.py:384-388. */
int anoddpm_conv_head_sep_supported(int C, int Cout /* [host] */);
long long anoddpm_conv_head_sep_workspace_bytes(int B, int H, int W /* [host] */);
int anoddpm_conv_head_sep(float* out_nchw, const float* x_nhwc, const float* gn_table, const float* w_packed,
                          const float* bias, float* workspace, int B, int C, int H, int W, void* stream);

/* Stem for single-channel images in one pass that also emits what the first ResBlock needs: the per-channel GroupNorm
 * statistics of the output (stats, [REPLICAS][B][Cout][2] fp64, accumulated; may be NULL) and an fp16 copy of the output
 * (out16, may be NULL) -- the raw operand of a convolution with gn_table.  Cin == 1, W % 8 == 0.  UNet.py:279-281. */
int anoddpm_conv_stem_fused(float* out_nhwc, void* out16_nhwc, double* stats, const float* x_nchw, const float* w_oihw,
                            const float* bias, int B, int Cin, int H, int W, int Cout, void* stream);

/* Input pipeline: the deterministic part of the reference's MRI slice transform on the device
 * (dataset.py:585-593, 631-640: image[:, s, :] -> ToPILImage -> CenterCrop(crop) -> Resize(bilinear) -> ToTensor ->
 * Normalize(0.5, 0.5)).  Sample b reads its slice at vols + vol_offset[b] with rows row_stride elements apart (so a slice
 * of a cached [Hs][D][Ws] volume needs no host-side gather); out is (B, 1, Ho, Wo) fp32.  The random affine jitter of the
 * training transform is NOT applied here. */
int anoddpm_slice_resize_norm(float* out, const float* vols, const int64_t* vol_offset /* [device] */, int B, int Hs, int Ws,
                              int64_t row_stride, int crop, int Ho, int Wo, void* stream);

/* QKV attention on qkv[B][T][3*C] (channel layout [head][q|k|v][ch], UNet.py:146-147):
 * out[b][t][h*ch + c] = sum_s softmax_s( (q_t . k_s) / sqrt(ch) ) v_s[c].   UNet.py:137-153. */
int anoddpm_attention(void* out /* B x T x C; fp32, or fp16 when out_mode == 2 */, const float* qkv, int B, int T, int C,
                      int heads, int out_mode /* as anoddpm_gn_apply */, void* stream);

/* The same attention on tcgen05 tensor cores (fp16 operands rounded to nearest, fp32 accumulation and softmax):
 * a pack pass (q, k scaled by ch^-1/4 as UNet.py:143, v transposed) and one CTA per (128 queries, sample x head).
 * Needs T and C/heads in {64,128,192,256} (anoddpm_attention_tc_supported) and a caller-owned, 128-byte aligned
 * workspace of anoddpm_attention_tc_workspace_bytes(B, T, C) bytes. */
int anoddpm_attention_tc_supported(int T, int C, int heads);
long long anoddpm_attention_tc_workspace_bytes(int B, int T, int C);
int anoddpm_attention_tc(void* out, const float* qkv, void* workspace, int B, int T, int C, int heads, int out_mode,
                         void* stream);

/* FP64 FMA throughput probe: `blocks` CTAs of 256 threads run iters x 8 dependent fma per thread (2 * 8 * iters * 256 *
 * blocks flop).  bench.py times it to obtain the FP64 roofline denominator of the simplex generator. */
int anoddpm_fp64_fma_probe(double* out, int blocks, int iters, void* stream);

/* ------------------------------------------------------------------ training: backward kernels ------------- */

/* Weight gradient of a 3x3 (padding 1) or 1x1 convolution on tcgen05 tensor cores (what autograd computes for
 * nn.Conv2d / nn.Conv1d weights in loss.backward(), diffusion_training.py:101; UNet.py:170-200, 115-118):
 *     dw[co][tap][ci] += alpha * sum_{b,y,x} dy[b,y,x,co] * a[b, y+ty-1, x+tx-1, ci]       (tap = 3*ty + tx; 1x1: tap 0)
 * dy, a: NHWC fp16 (the operands stay channel-contiguous: MN-major UMMA descriptors, no transposed copies); dw: fp32,
 * packed like the forward weights ([Cout][ksize^2][Cin]), ACCUMULATED (split-K partial sums are added with atomics, so
 * the caller zeroes it or passes a running gradient).  alpha = 1 / loss scale.
 * Needs Cin, Cout % 128 == 0, H % 8 == 0 and W == 8 or W % 16 == 0 (anoddpm_conv_wgrad_supported). */
int anoddpm_conv_wgrad_supported(int B, int H, int W, int Cin, int Cout, int ksize);
int anoddpm_conv_wgrad(float* dw, const void* dy_f16, const void* a_f16, int B, int H, int W, int Cin, int Cout, int ksize,
                       float alpha, void* stream);

/* Backward of out = resample(act(GroupNorm(cat(x1, x2) [+ shift]))) (the forward of anoddpm_gn_apply /
 * anoddpm_gn_apply_f16in; UNet.py:170-172, 190-193, 202-210), two passes over x and da:
 *   pass 1: sums[b][c] = (sum_p g, sum_p g * xhat), g = gradient wrt the GroupNorm output (da through the resample and
 *           SiLU'), pass 2: dx = rstd * (g * gamma - mean_group(g * gamma) - xhat * mean_group(g * gamma * xhat)) (+ addend).
 * x1: fp32 NHWC, or fp16 (x1_f16: the conv sum stored by out_f16; shift_c / shift_bc as in anoddpm_gn_apply_f16in; dx1 is
 * then fp16 and dsum[b][c] += sum_p dx -- the gradient of the per-(sample, channel) shift, i.e. of bias + embedding row).
 * da: [B][Ho][Wo][C1+C2] fp16 or fp32 at the OUTPUT resolution of `mode`.  add: gradient reaching x on the skip path
 * (add_mode 1 same resolution, 2 half resolution: x was avg-pooled there, 3 double resolution: x was upsampled there).
 * dgamma / dbeta (+= alpha * ...) may be NULL.  sums_ws: caller-owned B * C * 2 floats.  All activation gradients carry
 * the loss scale; alpha = 1 / scale un-scales the parameter gradients. */
int anoddpm_gn_silu_bwd(void* dx1, float* dx2, float* dgamma, float* dbeta, float* dsum, float* sums_ws, const void* x1,
                        int x1_f16, int C1, const float* x2, int C2, const double* stats1, const double* stats2,
                        const float* gamma, const float* beta, const float* shift_c, const float* shift_bc, int shift_stride,
                        const void* da, int da_f16, const void* add, int add_mode, int add_f16, int B, int H, int W,
                        int groups, float eps, int silu, int mode, float alpha, void* stream);

/* y = half(x * scale) (y may be NULL) and colsum[c] += alpha * sum_n x[n][c] (colsum may be NULL) for x fp32 [N][C]: the
 * fp16 operand copy of an output gradient and the bias gradient of the convolution that produced the output, one pass. */
int anoddpm_cast_colsum(void* y_f16, const float* x, float* colsum, int64_t N, int C, float scale, float alpha, void* stream);
int anoddpm_colsum_f16(float* colsum, const void* x_f16, int64_t N, int C, float alpha, void* stream);

/* Strided batched fp32 GEMM on CUDA cores: C[z](m, n) (+)= alpha * sum_k A[z](m, k) * B[z](k, n), element strides
 * (a_ms, a_ks), (b_ks, b_ns), ldc; batch z = z1 * n2 + z2 with per-matrix strides.  The attention core backward
 * (UNet.py:137-153: dV = P^T dO, dP = dO V^T, dQ = dL K, dK = dL^T Q) and the small timestep-MLP gradients. */
int anoddpm_gemm_batched(float* c, const float* a, const float* b, int M, int N, int K, int64_t a_ms, int64_t a_ks,
                         int64_t b_ks, int64_t b_ns, int64_t ldc, int n1, int n2, int64_t a_s1, int64_t a_s2, int64_t b_s1,
                         int64_t b_s2, int64_t c_s1, int64_t c_s2, float alpha, int accumulate, void* stream);
/* rows of length L: softmax in place; dl = p * (dp - <dp, p>) written over dp */
int anoddpm_softmax_rows(float* s, int64_t rows, int L, void* stream);
int anoddpm_softmax_bwd_rows(float* dp, const float* p, int64_t rows, int L, void* stream);
/* dx = dy * SiLU'(x), elementwise */
int anoddpm_silu_bwd(float* dx, const float* dy, const float* x, int64_t n, void* stream);
/* Stem conv (Cin = 1, UNet.py:279-281): dw[co][tap] += alpha * sum dy[b,p,co] * x[b,p+tap]; dy NHWC fp16, x NCHW fp32. */
int anoddpm_stem_wgrad(float* dw, const void* dy_f16, const float* x, int B, int H, int W, int Cout, float alpha, void* stream);
/* Head conv (Cout = 1, UNet.py:384-388): da (NHWC fp16) = dy (*) w^T, dw[tap][ci] += alpha * sum dy * a, dbias += alpha * sum dy. */
int anoddpm_head_bwd(void* da_f16, float* dw, float* dbias, const float* dy, const void* a_f16, const float* w, int B, int H,
                     int W, int C, float alpha, void* stream);

/* fp16 operand pack of a convolution weight (OIHW / OI1 fp32, the nn.Parameter): transposed = 0 the forward pack
 * dst[co][col_offset + tap * Cin + ci] (rows of row_stride halves; a fused 1x1 skip weight goes behind the 9 taps with
 * col_offset = 9 * Cin); transposed = 1 the data-gradient pack dst[ci][col_offset + (T-1-tap) * Cout + co] (flipped,
 * in/out swapped), so that d_input = anoddpm_conv(d_output, pack).  Parameters change every optimiser step; one pass each. */
int anoddpm_pack_conv_w(void* dst_f16, const float* w, int Cout, int Cin, int ksize, int64_t row_stride, int col_offset,
                        int transposed, void* stream);

/* ------------------------------------------------------------------ training: fused optimiser step -------- */

/* *out += sum(g[i]^2) in fp64 (caller zeroes).  First half of torch.nn.utils.clip_grad_norm_ (diffusion_training.py:104). */
int anoddpm_sumsq(double* out, const float* g, int64_t n, void* stream);

/* One AdamW step with gradient clipping and EMA over flat fp32 buffers of n elements (diffusion_training.py:75,104-107:
 * clip_grad_norm_(params, max_norm); AdamW(lr, betas, eps, weight_decay).step(); UNet.py:423-427 update_ema_params).
 * sumsq: device scalar from anoddpm_sumsq (NULL or max_norm <= 0: no clipping); ema may be NULL; step counts from 1. */
int anoddpm_adamw_ema_step(float* p, const float* g, float* m, float* v, float* ema, const double* sumsq, int64_t n,
                           float lr, float beta1, float beta2, float eps, float weight_decay, int step, float max_norm,
                           float ema_decay, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* ANODDPM_B200_H */
