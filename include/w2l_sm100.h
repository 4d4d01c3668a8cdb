/* libw2l_sm100 -- C ABI of the B200 (sm_100a) hot path for wav2letter_pytorch.
 *
 * The reference (assafmu/wav2letter_pytorch) has no FFI of its own: its hot path is a chain of torch
 * library calls made from Python.  Each entry point below replaces one of those call sites; the
 * reference file:line it stands in for is cited at the declaration.  The Python host package
 * (wav2letter_pytorch_b200/) binds these with ctypes and mirrors the reference's module API.
 *
 * Conventions
 *   - every pointer is a DEVICE pointer unless the name ends in _host; the caller owns all memory
 *     (including workspaces); the library never allocates, frees or synchronises; work is enqueued on
 *     `stream` (a cudaStream_t passed as void*)
 *   - return value: 0 = W2L_OK, otherwise an error code; w2l_last_error() gives thread-local text
 *   - activations are TIME-MAJOR: [B, T, C] with C contiguous ("rows" = frames); bf16 unless stated
 *   - conv weights are packed [k, Cout_pad, Cin] bf16 (tap-major; each tap is a K-major GEMM operand)
 *   - there is no CPU fallback: without a CUDA device every compute entry point returns W2L_ERR_CUDA
 */
#ifndef W2L_SM100_H_
#define W2L_SM100_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define W2L_OK 0
#define W2L_ERR_INVALID_ARGUMENT 1
#define W2L_ERR_CUDA 2
#define W2L_ERR_UNSUPPORTED 3

#define W2L_ACT_NONE 0
#define W2L_ACT_RELU 1      /* jasper.py:448  activation=torch.nn.ReLU()            */
#define W2L_ACT_CLAMP20 2   /* wav2letter.py:46 torch.clamp(output, min=0, max=20) */

#define W2L_PAD_ZERO 0      /* jasper.py:96-105 nn.Conv1d(padding=p)                */
#define W2L_PAD_REFLECT 1   /* wav2letter.py:28-34 nn.ReflectionPad1d               */

#define W2L_DTYPE_BF16 0
#define W2L_DTYPE_F32 1
/* OR-ed into the `act` argument of the BatchNorm / activation passes (w2l_bn_act_pad, w2l_bn_finalize_act_pad,
 * w2l_bn_act_bwd_reduce, w2l_bn_act_bwd_apply): the activation buffers of the call (z, res, y / dyp, dz, g_out) are fp32
 * instead of bf16 -- the fp32-faithful mode, whose GEMMs run with w2l_conv_desc::x_dtype = W2L_DTYPE_F32 */
#define W2L_STORE_F32 0x100
/* OR-ed into the `act` argument of w2l_bn_act_bwd_apply: red[C:2C] holds the RAW sums of g * (z - mean) that
 * w2l_conv1d_dgrad_wt_bnred accumulated; the pass multiplies them by invstd itself */
#define W2L_RED_RAW 0x200

int w2l_version(void);
const char* w2l_last_error(void);
/* number of kernels this library has launched in the calling process (bench.py's gpu_launches) */
int64_t w2l_launch_count(void);
/* hash of the sources the library was built from (csrc/ + this header); wav2letter_pytorch_b200/_lib.py compares it with the tree */
const char* w2l_source_hash(void);
/* Cap on the SMs the persistent conv GEMM grids occupy (0 = all).  The data-parallel host sets it to
 * (#SM - collective CTAs) so that the gradient all-reduce running beside backward gets SMs of its own
 * (Lightning DDP's bucketed all-reduce overlapped with backward, README.md:40 / config.yaml:21). */
int w2l_set_sm_budget(int32_t sms);
int32_t w2l_get_sm_budget(void);

/* Host-side Levenshtein distance between two int32 symbol sequences (HOST pointers).  Replaces the
 * python-Levenshtein calls behind Decoder.wer / Decoder.cer (decoder.py:31-60).  Returns -1 on bad input. */
int64_t w2l_edit_distance_host(const int32_t* a_host, int64_t n, const int32_t* b_host, int64_t m);
/* Batched form (one call per training step): pair p compares a[a_off[p]:a_off[p+1]] with b[b_off[p]:b_off[p+1]];
 * bit-parallel (Myers/Hyyro) and spread over `threads` host threads (0 = hardware concurrency, capped at 16). */
int w2l_edit_distance_batch_host(const int32_t* a_host, const int64_t* a_off_host, const int32_t* b_host,
                                 const int64_t* b_off_host, int64_t n_pairs, int64_t* out_host, int32_t threads);

/* ---------------------------------------------------------------------------------------------
 * Greedy CTC decoding.  Replaces GreedyDecoder.decode -> torch.max(probs, 2) + the per-frame Python
 * loop of process_string (decoder.py:104-119, 121-145).
 *   scores  [N, T, C] fp32, arbitrary N/T strides (elements), class stride 1
 *   sizes   [N] int32 or NULL (=> T for every utterance); values are clamped to [0, T]
 *   argmax  [N, T] int32 out : first maximal index, NaN counts as maximal (torch.max semantics)
 *   tokens  [N, T] int32 out : kept symbols, compacted to the front of each row
 *   offsets [N, T] int32 out : frame index of each kept symbol
 *   counts  [N]    int32 out : number of kept symbols
 * keep(t) = t < size && a[t] != blank && (t == 0 || a[t] != a[t-1])  (previous frame's RAW argmax).
 * workspace: w2l_greedy_decode_workspace_bytes(N, T) bytes.
 */
size_t w2l_greedy_decode_workspace_bytes(int64_t N, int64_t T);
int w2l_greedy_decode(const float* scores, int64_t N, int64_t T, int64_t C, int64_t stride_n, int64_t stride_t,
                      const int32_t* sizes, int32_t blank, int32_t* argmax, int32_t* tokens, int32_t* offsets,
                      int32_t* counts, void* workspace, size_t workspace_bytes, void* stream);

/* ---------------------------------------------------------------------------------------------
 * Device-side CER / WER / length-ratio of a decoded batch (replaces the per-utterance host loop of
 * ConvCTCASR.add_string_metrics, base_asr_models.py:53-69, and its device sync).
 *   tokens/counts      output of w2l_greedy_decode (collapsed label ids per utterance)
 *   ref_ids            [N, ref_stride] int32: the reference texts encoded as label ids (characters outside the label
 *                      set: n_labels + code point), ref_lens [N]; ref_stride <= 1023
 *   space_index        id of ' ': removed for CER (decoder.py:58 replace(' ', '')), word separator for WER (split())
 *   *_den              the host-known denominators: sum len(ref without spaces), sum #words, sum len(text)
 *   ratios [3] fp32 out: cer, wer, len_ratio.  Words are compared through 64-bit FNV-1a hashes of their ids.
 */
size_t w2l_string_metrics_workspace_bytes(int64_t N, int64_t T, int64_t ref_stride);
int w2l_string_metrics(const int32_t* tokens, const int32_t* counts, int64_t N, int64_t T, int32_t space_index,
                       const int32_t* ref_ids, const int32_t* ref_lens, int64_t ref_stride, float cer_den, float wer_den,
                       float len_den, float* ratios, void* workspace, size_t workspace_bytes, void* stream);

/* ---------------------------------------------------------------------------------------------
 * CTC loss + gradient.  Replaces self.criterion = nn.CTCLoss(blank=0, reduction='mean',
 * zero_infinity=True) and its backward (base_asr_models.py:23, 81, 90).
 *   x        [N, T, C] fp32 with N/T strides in elements (the reference passes out.transpose(0,1), a
 *            [T,N,C] view of this very layout).  from_logits=0: x holds log-probabilities and `grad`
 *            is d nll/d log_probs in the ATen convention (exp(lp) - occupancy).  from_logits=1: x
 *            holds raw logits, log_softmax is computed on the fly and `grad` is d nll/d logits
 *            (log_softmax backward fused).
 *   targets  [N, target_stride] int32, zero padded (data_loader.py:151-158)
 *   nll      [N] fp32 out: per-utterance negative log likelihood (inf -> 0 when zero_infinity)
 *   grad     [N, T, C] fp32 contiguous out, multiplied by grad_scale_n:
 *              reduction_mean=1 -> 1 / (N * max(S_n, 1))   (torch 'mean');  0 -> 1
 *            rows t >= input_length_n and infeasible utterances get exact zeros.  NULL skips it.
 *   loss     [1] fp32 out (optional): mean_n(nll_n / max(S_n,1)) if reduction_mean else sum_n nll_n
 * workspace: w2l_ctc_loss_workspace_bytes(N, T, S_max).
 */
size_t w2l_ctc_loss_workspace_bytes(int64_t N, int64_t T, int64_t S_max);
int w2l_ctc_loss(const float* x, int32_t from_logits, int64_t N, int64_t T, int64_t C, int64_t stride_n,
                 int64_t stride_t, const int32_t* targets, int64_t target_stride, const int32_t* input_lengths,
                 const int32_t* target_lengths, int32_t blank, int32_t zero_infinity, int32_t reduction_mean,
                 float* nll, float* grad, float* loss, void* workspace, size_t workspace_bytes, void* stream);

/* ---------------------------------------------------------------------------------------------
 * Conv1d as a tcgen05/TMEM implicit GEMM fed by TMA.  Replaces nn.Conv1d forward / backward-data /
 * backward-weight at wav2letter.py:35-36,42 and jasper.py:96-105,127,433,468 (stride 1; the stride-2
 * first layer goes through w2l_im2col_ncw first and then runs here with k=1).
 *
 * w2l_conv1d_fwd:  y[b, t, co] = bias[co] + sum_j sum_ci x[b, t + x_row_offset + j*dilation, ci] * w[j, co, ci]
 *   x  [B, x_rows, Cin] bf16 (already padded when x_row_offset = 0; rows outside [0, x_rows) read as 0,
 *      which is exactly Jasper's zero padding with x_row_offset = -p)
 *   w  [k, Cout_pad, Cin] bf16, Cout_pad % 16 == 0, Cin % 8 == 0
 *   y  [B, y_rows, ldy] (bf16 or fp32): rows t in [0, T_out) are written at y row (t + y_row_offset),
 *      columns [0, Cout)
 *   epilogue: v = acc + bias;  if scale: v = v*scale[co] + shift[co];  act(v)
 *   bn_stats (nullable) [2*Cout] fp32, zeroed by the caller: the epilogue adds the per-channel sum and sum of
 *      squares of the values it stores (after bf16 rounding) -- BatchNorm1d's training statistics
 *      (wav2letter.py:37, jasper.py:363) without a second pass over y; w2l_bn_finalize consumes it
 * w2l_conv1d_dgrad: dx[b, u, ci] = sum_j sum_co dy[b, u + dy_row_offset - j*dilation, co] * w[j, co, ci]
 *   (same packed weights, read as an MN-major operand; no transposed copy is kept)
 * w2l_conv1d_wgrad: dw[j, co, ci] (+)= sum_b sum_t dy[b, t, co] * x[b, t + x_row_offset + j*dilation, ci]
 *   dw [k, Cout, Cin] fp32; accumulate=0 requires dw to be zero-filled by the caller iff the kernel
 *   reports split-K > 1 through w2l_conv1d_wgrad_splits().
 */
typedef struct {
  int32_t B;            /* utterances                                  */
  int32_t T_out;        /* output rows per utterance                   */
  int32_t Cin, Cout;    /* logical channel counts                      */
  int32_t Cout_pad;     /* rows per tap in the packed weight tensor    */
  int32_t k, dilation;
  int32_t x_rows;       /* rows per utterance in the input buffer      */
  int32_t x_row_offset; /* input row read by tap 0 of output row 0     */
  int32_t y_rows;       /* rows per utterance in the output buffer     */
  int32_t y_row_offset;
  int32_t ldy;          /* output row pitch in elements                */
  int32_t y_dtype;      /* W2L_DTYPE_*                                 */
  int32_t act;          /* W2L_ACT_*                                   */
  int32_t x_dtype;      /* operand type of activations AND weights: W2L_DTYPE_BF16 (0), or W2L_DTYPE_F32 = fp32 storage
                           multiplied as tf32 with fp32 accumulation (the fp32-faithful mode; reference arithmetic is
                           nn.Conv1d in fp32, wav2letter.py:35-36) */
} w2l_conv_desc;

int w2l_conv1d_fwd(const void* x, const void* w, const float* bias, const float* scale, const float* shift, float* bn_stats,
                   void* y, const w2l_conv_desc* d, void* stream);
/* Optional fp32 scratch for w2l_conv1d_fwd (caller-owned, ZERO-filled once, >= 4096 + 148*128*256*4 bytes, 256-byte aligned).
 * With it, a forward launch whose last wave of 128 x BN tiles is less than 3/4 full cuts those tiles along K over the idle SMs;
 * the pieces meet in the scratch and the last one to arrive runs the epilogue (fixed summation order: deterministic).  Forward
 * launches that use it must be stream-ordered with respect to each other.  NULL/0 unregisters. */
int w2l_set_gemm_scratch(void* scratch, size_t bytes);
/* K pieces per tile of the last wave that w2l_conv1d_fwd would use for this descriptor right now (0/1 = no split) */
int32_t w2l_conv1d_fwd_tail_parts(const w2l_conv_desc* d);
int w2l_conv1d_dgrad(const void* dy, const void* w, void* dx, const w2l_conv_desc* d, void* stream);
/* Backward-data with a second, transposed bf16 shadow wt [k, Cin_pad16, Cout_pad] (wt[k-1-j][ci][co] = w[j][co][ci],
 * written by w2l_pack_wt): both GEMM operands are then K-major, which the tensor pipe consumes ~30% faster than the
 * MN-major read of w2l_conv1d_dgrad (profiles/).  Passing B=1 with T_out = x_rows = y_rows = B*(T+pad) over a dy buffer
 * whose per-utterance tail rows are zero runs the whole batch as ONE flat row space (no per-utterance tile padding).
 * dy rows hold Cout_pad columns (ldy >= Cout_pad, zero padded) or just Cout columns (Cout <= ldy < Cout_pad: hidden widths
 * that are multiples of 8 but not of 16); in the latter case the tail of the last contraction chunk reads as zero. */
int w2l_conv1d_dgrad_wt(const void* dy, const void* wt, void* dx, const w2l_conv_desc* d, void* stream);
/* The same backward-data GEMM with the BatchNorm-backward REDUCTION of the block below folded into its epilogue.  dx is the
 * gradient with respect to that block's (halo-padded) output y [red->B, pad_left + T + pad_right, Cin]; the rows are in registers
 * when they are stored, so the epilogue also accumulates what w2l_bn_act_bwd_reduce would compute in a separate pass over (dy, z):
 *   red[0:C]  += sum_rows g,   red[C:2C] += sum_rows g * (z - mean)        g = gate(z*scale + shift) * keep-bit * dx/keep
 * (a reflect-halo row counts with the z / gate of the interior row it mirrors; rows t >= lens[b] carry none).  red[C:2C] is RAW --
 * not yet multiplied by invstd: the apply pass takes it with W2L_RED_RAW OR-ed into its `act`.  bf16 activations, blocks without
 * a residual branch; with dropout the stored keep-bits are required.  Replaces one pass of nn.BatchNorm1d / clamp backward
 * (wav2letter.py:43-46, jasper.py:363-376) per layer. */
typedef struct {
  const void* z;          /* the block's conv output, bf16 [B, T, C], C = the Cin of the GEMM's descriptor */
  const void* drop_mask;  /* keep-bits [B*T*C/8] written by the forward pass, or NULL (no dropout) */
  const float* scale;     /* the block's BatchNorm scale, shift, batch mean [C] (fin rows 0, 1, 2) */
  const float* shift;
  const float* mean;
  const int32_t* lens;    /* nullable */
  float* red;             /* fp32 [2C], accumulated into (zero before the launch) */
  int32_t B, T, pad_left, pad_right;
  int32_t act;            /* W2L_ACT_* of the block */
  float drop_p;
} w2l_bn_reduce;
int w2l_conv1d_dgrad_wt_bnred(const void* dy, const void* wt, void* dx, const w2l_conv_desc* d, const w2l_bn_reduce* red, void* stream);
int w2l_pack_wt(const float* w, void* wt, int32_t k, int32_t Cout, int32_t Cin, int32_t Cout_pad, int32_t Cin_pad, void* stream);
/* the same shadow in fp32 (operand of w2l_conv1d_dgrad_wt with x_dtype = F32) */
int w2l_pack_wt_f32(const float* w, float* wt, int32_t k, int32_t Cout, int32_t Cin, int32_t Cout_pad, int32_t Cin_pad, void* stream);
int32_t w2l_conv1d_wgrad_splits(const w2l_conv_desc* d);
int w2l_conv1d_wgrad(const void* dy, const void* x, float* dw, const w2l_conv_desc* d, void* stream);
/* fp32-faithful mode (x_dtype = F32): the weight gradient over TRANSPOSED fp32 operands dyT [B, Cout, dy_pitch] and
 * xT [B, Cin, x_pitch] (time contiguous, as w2l_tm_to_ct_f32 writes them; pitches in floats, multiples of 4), multiplied as tf32:
 *   dw[j, co, ci] += sum_b sum_t dyT[b, co, t] * x[b, t + x_row_offset + j*dilation, ci],  t < T_out, rows in [0, x_rows)
 * TMA needs the time coordinate of a load 16-byte aligned, so x comes as up to four copies DELAYED by s = 0..3 rows:
 * xT_shifted (HOST array of 4 device pointers), xT_shifted[s][b, ci, u] = x[b, u - s, ci] with s zeros in front
 * (x_pitch >= x_rows + 3); only the residues (-(x_row_offset + j*dilation)) mod 4 that occur must be non-null.
 * dw [k, Cout, Cin] fp32 must be ZERO on entry. */
int w2l_conv1d_wgrad_t(const float* dyT, int64_t dy_pitch, const float* const* xT_shifted, int64_t x_pitch, float* dw,
                       const w2l_conv_desc* d, void* stream);

/* ---------------------------------------------------------------------------------------------
 * Depthwise Conv1d (groups = channels): the first half of the separable sub-blocks of the shipped model/jasper.yaml
 * (jasper.py:318-341; the pointwise half is w2l_conv1d_* with k = 1).  Time-major bf16 activations, fp32 weights
 * stored [k, C]; zero padding `pad` on both sides.
 *   fwd:   y [B,T_out,C];  rows t >= out_lens[b] are written as 0 (the consumer MaskedConv1d's masked_fill)
 *   dgrad: dx [B,T,C] (stride 1; w2l_depthwise_dgrad_strided for stride > 1); dy rows t >= dy_lens[b] are read as 0
 *   wgrad: dw [k,C] fp32, ACCUMULATED with atomics (zero it first)
 */
int w2l_depthwise_fwd(const void* x, const float* w, void* y, int32_t B, int32_t T, int32_t C, int32_t T_out, int32_t k,
                      int32_t stride, int32_t dilation, int32_t pad, const int32_t* out_lens, void* stream);
int w2l_depthwise_dgrad(const void* dy, const float* w, void* dx, int32_t B, int32_t T, int32_t C, int32_t T_out, int32_t k,
                        int32_t dilation, int32_t pad, const int32_t* dy_lens, void* stream);
/* backward-data of a depthwise conv with stride >= 1 (gather over the taps that land on a multiple of the stride) */
int w2l_depthwise_dgrad_strided(const void* dy, const float* w, void* dx, int32_t B, int32_t T, int32_t C, int32_t T_out, int32_t k,
                                int32_t stride, int32_t dilation, int32_t pad, const int32_t* dy_lens, void* stream);
int w2l_depthwise_wgrad(const void* dy, const void* x, float* dw, int32_t B, int32_t T, int32_t C, int32_t T_out, int32_t k,
                        int32_t stride, int32_t dilation, int32_t pad, const int32_t* dy_lens, void* stream);
/* the four with fp32 activations (a separable Jasper block in the fp32-faithful mode; plain fp32 FMAs, as the reference's depthwise
 * nn.Conv1d runs them) */
int w2l_depthwise_fwd_f32(const float* x, const float* w, float* y, int32_t B, int32_t T, int32_t C, int32_t T_out, int32_t k,
                          int32_t stride, int32_t dilation, int32_t pad, const int32_t* out_lens, void* stream);
int w2l_depthwise_dgrad_f32(const float* dy, const float* w, float* dx, int32_t B, int32_t T, int32_t C, int32_t T_out, int32_t k,
                            int32_t dilation, int32_t pad, const int32_t* dy_lens, void* stream);
int w2l_depthwise_dgrad_strided_f32(const float* dy, const float* w, float* dx, int32_t B, int32_t T, int32_t C, int32_t T_out, int32_t k,
                                    int32_t stride, int32_t dilation, int32_t pad, const int32_t* dy_lens, void* stream);
int w2l_depthwise_wgrad_f32(const float* dy, const float* x, float* dw, int32_t B, int32_t T, int32_t C, int32_t T_out, int32_t k,
                            int32_t stride, int32_t dilation, int32_t pad, const int32_t* dy_lens, void* stream);

/* ---------------------------------------------------------------------------------------------
 * Memory-bound companions of the conv kernels (all time-major).
 */
/* [B, F, T] fp32 NCW (the collated batch, data_loader.py:149-158) -> [B, rows, k*F] bf16 with
 * out[b, r, j*F + f] = xpad[b, f, r*stride + j*dilation], xpad = x padded by (pad_left, ...) in `pad_mode`.
 * k=1, stride=1 gives the plain padded time-major copy.  wav2letter.py:41 + the layer-0 unfold. */
int w2l_im2col_ncw(const float* x, void* out, int32_t B, int32_t F, int32_t T, int32_t rows, int32_t k, int32_t stride,
                   int32_t dilation, int32_t pad_left, int32_t pad_mode, const int32_t* lens, void* stream);
/* the same with fp32 output (fp32-faithful mode) */
int w2l_im2col_ncw_f32(const float* x, float* out, int32_t B, int32_t F, int32_t T, int32_t rows, int32_t k, int32_t stride,
                       int32_t dilation, int32_t pad_left, int32_t pad_mode, const int32_t* lens, void* stream);
/* Strided layers beyond the first (Conv1dBlock with stride > 1, wav2letter.py:24-38; strided JasperBlock, jasper.py:289-298):
 * time-major bf16 x [B, x_rows, C] -> out [B, T_out, k*C] with out[b, t, j*C + c] = x[b, t*stride + j*dilation - pad_left, c]
 * (zero outside [0, x_rows)); the conv then runs as a k=1 GEMM over `out` with weights stored [Cout, k, Cin]. */
int w2l_im2col_tm(const void* x, void* out, int32_t B, int32_t x_rows, int32_t C, int32_t T_out, int32_t k, int32_t stride,
                  int32_t dilation, int32_t pad_left, void* stream);
/* adjoint of w2l_im2col_tm: dcol [B, T_out, k*C] bf16 -> dx [B, x_rows, C] bf16 (every row written; fp32 accumulation) */
int w2l_col2im_tm(const void* dcol, void* dx, int32_t B, int32_t x_rows, int32_t C, int32_t T_out, int32_t k, int32_t stride,
                  int32_t dilation, int32_t pad_left, void* stream);
/* time-major bf16/fp32 [B, T, C] (row pitch ld) -> NCW fp32 [B, C, T] */
int w2l_tm_to_ncw(const void* x, int32_t x_dtype, float* out, int32_t B, int32_t T, int32_t C, int32_t x_rows,
                  int32_t x_row_offset, int32_t ld, void* stream);
/* time-major fp32 rows [x_row_offset, x_row_offset + T) of x [B, x_rows, ld] -> channel-major out [B, C, out_pitch] fp32 (the pad
 * [T, out_pitch) is left untouched): the transposed operands of w2l_conv1d_wgrad_t */
int w2l_tm_to_ct_f32(const float* x, float* out, int32_t B, int32_t T, int32_t C, int32_t x_rows, int32_t x_row_offset, int32_t ld,
                     int32_t out_pitch, void* stream);
/* NCW fp32 [B, C, T] -> time-major bf16 [B, T, C] (dense) */
int w2l_ncw_to_tm(const float* x, void* out, int32_t B, int32_t C, int32_t T, void* stream);

/* per-channel sum / sum of squares over all B*T rows of z [B*T, C] bf16 -> stats[0:C], stats[C:2C]
 * (fp32, must be zeroed by the caller).  nn.BatchNorm1d training statistics, wav2letter.py:43, jasper.py:363 */
int w2l_bn_stats(const void* z, int64_t rows, int32_t C, float* stats, void* stream);
/* stats -> scale/shift (+ saved mean / invstd) and the running-stat update
 * r = (1-momentum)*r + momentum*batch (unbiased variance), as torch.nn.BatchNorm1d does. */
int w2l_bn_finalize(const float* stats, int64_t rows, int32_t C, const float* gamma, const float* beta,
                    const float* conv_bias /* nullable: added to the mean for running_mean only when z excludes it */,
                    float eps, float momentum, float* running_mean, float* running_var, float* scale, float* shift,
                    float* mean, float* invstd, int64_t* num_batches_tracked /* nullable: += 1 */, void* stream);
/* Jasper's sequence-length bookkeeping for a whole encoder in one launch (MaskedConv1d.get_seq_len / forward,
 * jasper.py:91-95,107-119): conv j truncates the incoming lengths to integers, masks with them, and hands on
 * (len + 2*pad - dil*(k-1) - 1) / stride + 1 in float32.  conv_params_host: HOST int32 [n_convs][4] =
 * (kernel, stride, dilation, padding); stride 0 = a conv built with use_mask=False (lengths pass through).
 * lens_out [n_convs+1][B] int32: row j = truncated lengths entering conv j, row n_convs = output lengths (also written
 * as int64 to final_out when given). */
int w2l_lens_chain(const void* lens_in, int32_t lens_is_int64, int32_t B, const int32_t* conv_params_host, int32_t n_convs,
                   int32_t* lens_out, int64_t* final_out, void* stream);
/* y = act(dropout(z*scale + shift [+ res*res_scale + res_shift])), written into a (possibly halo'd)
 * buffer: y[b, pad_left + t, c]; reflect halos of pad_left / pad_right rows are filled from the
 * interior (nn.ReflectionPad1d of the NEXT layer, wav2letter.py:28-34,41); rows t >= lens[b] are zeroed
 * when lens != NULL (the masked_fill of the consumer MaskedConv1d, jasper.py:116-119).
 * dropout (nn.Dropout, wav2letter.py:44 / jasper.py:372-376): keep-bits from a counter-based splitmix64 stream
 * (seed, element index / 8), 16 bits per element, p = drop_p (0 disables); drop_mask (nullable, B*T*C/8 bytes) receives
 * the keep-bits so that the backward passes read them back instead of re-deriving them. */
/* Optional device-resident dropout epoch (one uint64, caller-owned; NULL unregisters; per process like the GEMM scratch).  While
 * registered, every launch that draws keep-bits uses seed + *epoch * 0xA0761D6478BD642F instead of its by-value seed: a captured
 * CUDA graph replays its by-value arguments unchanged, so a step captured once (graph_step.py) bumps the epoch inside the graph
 * to draw a fresh mask per replay, the way the reference's nn.Dropout draws one per call (wav2letter.py:44). */
int w2l_set_dropout_epoch(const uint64_t* epoch);
int w2l_bn_act_pad(const void* z, const float* scale, const float* shift, const void* res, const float* res_scale,
                   const float* res_shift, void* y, int32_t B, int32_t T, int32_t C, int32_t pad_left,
                   int32_t pad_right, int32_t act, float drop_p, uint64_t seed, const int32_t* lens, void* drop_mask,
                   void* stream);
/* The training-mode pass: w2l_bn_finalize folded into w2l_bn_act_pad.  scale / shift come from the batch statistics
 * `stats` ([2C] sum, sum of squares over stat_rows rows, as the conv epilogue or w2l_bn_stats leaves them): every CTA derives
 * them for its channels, one of them writes fin [4][C] = (scale, shift, mean, invstd) for the backward passes and applies the
 * running-statistics / num_batches_tracked update of nn.BatchNorm1d (wav2letter.py:37, jasper.py:363; conv_bias is added to
 * the running mean because the conv here runs without its bias, which training-mode BatchNorm cancels).
 * zero_ptr / zero_count (nullable): a small fp32 buffer this launch clears for a LATER kernel -- the caller passes the layer's
 * backward reduction buffer, so that no memset launch is needed for it. */
int w2l_bn_finalize_act_pad(const void* z, const float* stats, int64_t stat_rows, const float* gamma, const float* beta,
                            const float* conv_bias, float eps, float momentum, float* running_mean, float* running_var,
                            int64_t* num_batches_tracked, float* fin, const void* res, const float* res_scale,
                            const float* res_shift, void* y, int32_t B, int32_t T, int32_t C, int32_t pad_left,
                            int32_t pad_right, int32_t act, float drop_p, uint64_t seed, const int32_t* lens, void* drop_mask,
                            float* zero_ptr, int32_t zero_count, void* stream);
/* In-place reflect halo for a buffer whose interior rows [pad_left, pad_left+T) were written by the fused
 * conv epilogue (inference: BatchNorm folded into scale/shift, wav2letter.py:41-46). */
int w2l_reflect_halo(void* y, int32_t B, int32_t T, int32_t C, int32_t pad_left, int32_t pad_right, void* stream);
int w2l_reflect_halo_f32(float* y, int32_t B, int32_t T, int32_t C, int32_t pad_left, int32_t pad_right, void* stream);
/* Backward of the above + BatchNorm backward, two passes over (dy_padded, z):
 *   g = fold_reflect(dyp)[b,t,c] * act'(.) * dropmask;  pass 1 reduces sum(g), sum(g*xhat) into
 *   red[0:C], red[C:2C] (zero on entry: one atomic per channel and CTA); pass 2 writes
 *   dz = gamma*invstd*(g - mean(g) - xhat*mean(g*xhat)), walking the rows in the reverse order of pass 1 (L2 reuse).
 *   dgamma = red[C:2C], dbeta = red[0:C]; pass 2 copies red to red_out (nullable) so that a persistent red buffer can be
 *   recycled, and clears zero_ptr[0:zero_count] (nullable; the caller passes the layer's forward statistics buffer). */
int w2l_bn_act_bwd_reduce(const void* dyp, const void* z, const void* res, const float* scale, const float* shift,
                          const float* res_scale, const float* res_shift, const float* mean, const float* invstd,
                          float* red, int32_t B, int32_t T, int32_t C, int32_t pad_left, int32_t pad_right, int32_t act,
                          float drop_p, uint64_t seed, const int32_t* lens, const void* drop_mask, void* stream);
/* dz [B, dz_rows, C]: rows >= T zero-filled; g_out (nullable): the masked g, bf16 [B,T,C] (Jasper's residual branch) */
int w2l_bn_act_bwd_apply(const void* dyp, const void* z, const void* res, const float* scale, const float* shift,
                         const float* res_scale, const float* res_shift, const float* mean, const float* invstd,
                         const float* gamma, const float* red, void* dz, int32_t dz_rows, void* g_out, int32_t B, int32_t T,
                         int32_t C, int32_t pad_left, int32_t pad_right, int32_t act, float drop_p, uint64_t seed,
                         const int32_t* lens, const void* drop_mask, float* red_out, float* zero_ptr, int32_t zero_count,
                         void* stream);

/* logits [rows, ld] fp32 -> log_softmax / softmax over the first C columns -> out [rows, C] fp32
 * (wav2letter.py:87, jasper.py:470-473).  mode 0 = log_softmax, 1 = softmax.  nan_flag (nullable, device int32, zeroed
 * by the caller) is OR-ed with 1 when an output is NaN: the device half of Jasper's `assert not NaN` (jasper.py:474). */
int w2l_log_softmax(const float* logits, int32_t ld, float* out, int64_t rows, int32_t C, int32_t mode, int32_t* nan_flag,
                    void* stream);
/* d logits = g - exp(lp) * sum_c g  (log_softmax backward), scaled by *gscale (device scalar, nullable),
 * written as bf16 into [rows, ld_out] with zero padding in columns >= C. */
int w2l_log_softmax_bwd(const float* g, const float* lp, const float* gscale, void* dlogits, int32_t ld_out,
                        int64_t rows, int32_t C, int32_t fused_identity, void* stream);
int w2l_log_softmax_bwd_f32(const float* g, const float* lp, const float* gscale, float* dlogits, int32_t ld_out,
                            int64_t rows, int32_t C, int32_t fused_identity, void* stream);
/* column sums of a bf16 [rows, ld] matrix -> out[C] fp32 (zeroed by the caller): bias gradient */
int w2l_colsum(const void* x, int64_t rows, int32_t C, int32_t ld, float* out, void* stream);
int w2l_colsum_f32(const float* x, int64_t rows, int32_t C, int32_t ld, float* out, void* stream);
/* fp32 -> bf16 cast (packed weight shadow refresh) */
int w2l_cast_bf16(const float* src, void* dst, int64_t n, void* stream);

/* ---------------------------------------------------------------------------------------------
 * Fused multi-tensor NovoGrad step.  Replaces the per-parameter Python loop of Novograd.step
 * (novograd.py:52-114): layer-wise second moment of ||g||^2, normalised gradient, decoupled weight
 * decay term, first moment, parameter update -- and refreshes the bf16 weight shadow in the same pass.
 * Arrays of device pointers / sizes live in device memory (n_tensors entries).  The work is cut into chunks of
 * w2l_novograd_chunk() elements, one CTA each: chunk_prefix[t] (device, int32 [n_tensors]) = index of tensor t's first
 * chunk, i.e. the exclusive prefix sum of ceil(numel[t] / chunk); n_chunks = the total.
 */
int32_t w2l_novograd_chunk(void);
int w2l_novograd_step(float* const* params, float* const* grads, float* const* exp_avg, float* exp_avg_sq /*[n]*/,
                      float* max_exp_avg_sq /*[n] or NULL: amsgrad=True*/, void* const* shadow_bf16 /* nullable entries */, const int64_t* numel, const int32_t* chunk_prefix,
                      int32_t n_tensors, int32_t n_chunks, float lr, float beta1, float beta2, float eps, float weight_decay,
                      int32_t grad_averaging, float* norms_ws /*[n] scratch*/, void* stream);

/* ---------------------------------------------------------------------------------------------
 * Data-parallel gradient averaging over NVLink peer memory.  Replaces the bucketed NCCL all-reduce that Lightning DDP
 * performs for the reference (README.md:40 `--gpus`, config.yaml:21) on the hot path's only exchange.
 * All ranks keep their gradients in a symmetric arena (same byte offsets everywhere).  HOST tables:
 *   peer_data_host[r]   this process's mapping of rank r's arena (r == rank: the local arena)
 *   peer_flags_host[r]  this process's mapping of rank r's flag block, ctas*world uint32, zero-initialised once
 *   multicast_base      NVLS multicast mapping of the arena (in-switch reduction), or NULL for the peer-to-peer path
 * Averages arena[offset : offset+numel] (floats, both multiples of 4) across ranks in place; every rank ends up with
 * bit-identical values.  seq: the same, strictly increasing EVEN number on every rank for successive calls (seq and
 * seq+1 tag the two barriers).  Every rank must enqueue the same calls in the same order.
 */
int w2l_grad_allreduce(void* const* peer_data_host, void* const* peer_flags_host, void* multicast_base, int64_t offset,
                       int64_t numel, int32_t rank, int32_t world, uint32_t seq, int32_t ctas, void* stream);

/* ---------------------------------------------------------------------------------------------
 * CTC prefix beam search on the HOST.  Replaces prefix_beam_search / PrefixBeamSearchLMDecoder.decode (decoder.py:147-267).
 *   probs_host   [n_utt, T_max, F] float64 probabilities (>= 0); frames_host[u] frames are used (NULL: T_max)
 *   blank        index of the blank label; space_id / end_id: index of ' ' / the end character ('>') or -1 if not a label
 *   is_word_host / is_term_host [F]: label matches the regex class \w / [\s|>] (the reference counts words with
 *                r'\w+[\s|>]' for the (words+1)**beta prior)
 *   k, alpha, beta, prune: beam width, LM weight, word-count exponent, per-frame emission threshold
 *   lm           optional language-model callback (label ids of the stripped prefix -> probability), or NULL (== 1)
 *   out_ids_host [n_utt, out_stride] int32 best prefix per utterance, out_len_host [n_utt] its length,
 *   out_score_host [n_utt] (optional) A_next[best] * (words+1)**beta of the last frame
 * Same float64 arithmetic, candidate order and tie-breaking as the reference: identical transcripts.
 */
typedef double (*w2l_lm_callback)(const int32_t* ids, int64_t n, void* user);
int w2l_prefix_beam_search_host(const double* probs_host, const int64_t* frames_host, int64_t n_utt, int64_t T_max, int64_t F,
                                int32_t blank, int32_t space_id, int32_t end_id, const uint8_t* is_word_host,
                                const uint8_t* is_term_host, int32_t k, double alpha, double beta, double prune,
                                w2l_lm_callback lm, void* lm_user, int32_t* out_ids_host, int64_t out_stride,
                                int64_t* out_len_host, double* out_score_host, int32_t threads);

/* ---------------------------------------------------------------------------------------------
 * Feature front-end for a collated batch.  Replaces SpectrogramExtractor._get_spect / extract (data/data_loader.py:33-88)
 * and the zero padding of _collator (data_loader.py:149-158):
 *   x = audio + dither * noise;  y[0] = x[0], y[i] = x[i] - preemph * x[i-1];  centred STFT (reflect padding n_fft/2, hop,
 *   `window` [win_length] centred inside n_fft), power = |X|^2, mel = mel_fb [n_mels, n_fft/2+1] @ power,
 *   log1p(mel + log_guard), then per utterance and feature: (v - mean_t) / (std_t(unbiased) + norm_eps).
 *   audio        [B, audio_stride] fp32, audio_lens[b] valid samples (>= 2); frames_b = 1 + audio_lens[b] / hop
 *   dither_noise [B, audio_stride] fp32 standard-normal noise or NULL (the reference draws torch.randn on the host)
 *   out          [B, n_mels, T_max] fp32; frames t >= frames_b are zero (the collator's padding)
 * workspace: w2l_logmel_workspace_bytes(B, T_max, n_mels).
 */
size_t w2l_logmel_workspace_bytes(int32_t B, int32_t T_max, int32_t n_mels);
int w2l_logmel_features(const float* audio, int64_t audio_stride, const float* dither_noise, const int32_t* audio_lens, int32_t B,
                        int32_t n_fft, int32_t win_length, int32_t hop, const float* window, const float* mel_fb, int32_t n_mels,
                        float dither, float preemph, float log_guard, float norm_eps, float* out, int32_t T_max, void* workspace,
                        size_t workspace_bytes, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* W2L_SM100_H_ */
