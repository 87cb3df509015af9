/*
 * fq3_codec.h — C ABI of the 12 Hz speech-tokenizer decoder kernels (libfq3codec.so).
 *
 * Replaces `speech_tokenizer.decode({"audio_codes": [1,T,16]})` of the un-vendored qwen_tts package, which the
 * reference calls eagerly (cuDNN/ATen conv1d, conv_transpose1d, cuBLAS; SURVEY.md §2c last row) at
 * faster_qwen3_tts/model.py:642, :782, :811, :884, :971, :988, :1054, :1136, :1153.
 *
 * The library is a kernel-launch interpreter: the host (Python, mirroring the reference's Python host) lowers the
 * decoder for a given frame count T to a flat list of ops over channels-last bf16 activations; every dense
 * contraction (pointwise / dilated / transposed convs as implicit GEMMs, transformer projections) goes through one
 * tensor-core GEMM kernel with fused epilogues, the rest (RVQ gather, norms, RoPE, sliding-window attention,
 * depthwise conv) are small CUDA-core kernels.  Plain pointers only; all pointers are device addresses.
 */
#ifndef FQ3_CODEC_H_
#define FQ3_CODEC_H_
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

enum fq3c_kind {
  FQ3C_GEMM = 0,      /* C[M,N] = epi(sum_t A[m + tap_off[t], :cin] . B[n, t*cin:(t+1)*cin])            */
  FQ3C_RVQ = 1,       /* A=codes i64 [M,Q]; B=codebooks bf16 [Q,cb,dim]; C[M, 2*dim] = [sum first | sum rest] */
  FQ3C_RMSNORM = 2,   /* C[M,N] = rmsnorm(A[M,N]) * scale(f32 [N]), eps=f0                                  */
  FQ3C_ROPE = 3,      /* in place on A[M, lda]: i0 heads of dim i1 starting at column i2, theta=f0; row m is position
                         m + *p0 when p0 (device int32) is given (stateful decode), else m                       */
  FQ3C_ATTN = 4,      /* A=qkv [M, lda] (q | k | v), i0 heads, i1 kv heads, i2 head_dim, window K; C[M, i0*i2].  Stateful
                         decode: A holds `taps` history rows in front of the M query rows (query m = row taps + m), of
                         which the last min(*p0, taps) are valid (p0 = device int32: positions decoded so far).
                         flags bit 0 (ABI 5): prompt-prefill form — head_dim 128, no history rows, M <= 2048: a two-phase
                         kernel (lanes over keys for the scores, lanes over dims for the output) with the same rounding
                         points; fp32 sums in another order                                                           */
  FQ3C_DWCONV = 5,    /* depthwise causal conv k=taps: C[m,c] = bias[c] + sum_j B(f32)[c,j] A[m-(taps-1)+j, c]; rows down
                         to -i0 in front of A are history rows (stateful decode), further back reads as zero        */
  FQ3C_LAYERNORM = 6, /* C = layernorm(A) * scale + bias (f32), eps=f0                                      */
  FQ3C_SNAKE = 7,     /* C = A + p1[c] * sin(A * p0[c])^2   (p0 = exp(alpha), p1 = 1/(exp(beta)+1e-9))      */
  FQ3C_COPY = 9,      /* C[M, N] = A[M, N] (bf16 rows; lda / ldc): history roll of the stateful decode               */
  FQ3C_ADVANCE = 10,  /* *(int32*)C += i0 (one thread): the stateful decode's position counter                        */
  FQ3C_ROLL = 11,     /* every history roll of a stateful chunk + the counter in ONE launch (ABI 5): A = device int64 [M, 5]
                         rows (buffer address, hist rows, new rows, cols, ld): rows [new, new + hist) of each bf16 buffer
                         move to [0, hist) (overlapping ranges in batches of `new` rows, a block barrier between them);
                         then *(int32*)C += i0.  Replaces 2 x FQ3C_COPY per buffer (through scratch) + FQ3C_ADVANCE    */
  FQ3C_QKNORM_ROPE_KV = 8 /* dense talker prefill: A = fused qkv rows [M, lda], i0 q heads, i1 kv heads of dim 128; q/k heads get the
                             per-head RMSNorm (bf16 gamma p0 / p1, eps f0) and the rotary embedding from the bf16 tables B (cos) /
                             bias (sin) at position row + i2, in place; finished k / v rows also go to the static KV cache
                             C (K) / C2 (V), each [kv_head][ldc][128] (talker_graph.py:153-170 without the copy)            */
};

enum fq3c_flags {
  FQ3C_BIAS = 1, FQ3C_GELU = 2, FQ3C_RESID = 4, FQ3C_SCALE = 8, FQ3C_SWIGLU = 16, FQ3C_CLAMP = 32,
  FQ3C_OUT_F32 = 64, FQ3C_SNAKE2 = 128 /* also write C2 = snake(C) */, FQ3C_SILU = 256 /* text_projection: Linear -> SiLU */
};

typedef struct fq3c_op {
  int32_t kind, flags;
  int32_t M, N, K;       /* GEMM: K = taps * cin */
  int32_t taps, cin;
  int32_t a_rows;        /* valid rows of A (rows outside [0, a_rows) read as zero) */
  int32_t col_mod;       /* bias / scale / snake parameter index = col % col_mod */
  int32_t tap_off[8];
  int32_t lda, ldc, ldr;
  int32_t i0, i1, i2;
  float f0, f1;
  const void* A;
  const void* B;
  const void* bias;      /* f32 */
  const void* res;       /* bf16 [M, ldr] */
  const void* scale;     /* f32 */
  const void* p0;        /* f32 */
  const void* p1;        /* f32 */
  void* C;
  void* C2;
  void* ws;              /* GEMM, optional: fp32 workspace for split-K partial tiles (>= splits * M * round8(N) * 4 bytes are used); */
  int64_t ws_bytes;      /* NULL / 0 = never split.  Ops of one stream-ordered list may share it.                                  */
  int32_t m_begin;       /* GEMM: only output rows [m_begin, M) are computed (tail-only streaming decode); pointers stay at row 0 */
  int32_t reserved;
  /* GEMM, optional (ABI 5): the RMSNorm that reads the finished output rows, fused behind the epilogue — a decoder layer's norm
   * reads exactly the residual stream its o / down projection just wrote.  norm_out bf16 [M, norm_ld] = FQ3C_RMSNORM(C rows)
   * with weights norm_w (f32 [N]) and eps norm_eps, same arithmetic and summation order as the stand-alone op.  With split-K it
   * runs inside the reduction kernel (one CTA per row), otherwise as one more launch.  Not with FQ3C_SWIGLU / FQ3C_OUT_F32. */
  void* norm_out;
  const void* norm_w;
  int32_t norm_ld;
  float norm_eps;
} fq3c_op;

int fq3c_abi_version(void);
const char* fq3c_last_error(void);
/* Launch every op in order on `stream` (no synchronisation). Returns 0 or a negative error code. */
int fq3c_run(const fq3c_op* ops, int n_ops, void* stream);
/* number of kernel launches issued so far by this process through fq3c_run / fq3c_graph_launch */
int64_t fq3c_launch_count(void);
/* A decode plan is static for a given frame count: capture its launches once (stream capture of fq3c_run) and replay
 * them as one CUDA graph.  The ops' buffers must stay alive and at the same addresses for the life of the graph. */
int fq3c_graph_create(const fq3c_op* ops, int n_ops, void** out_graph);
int fq3c_graph_launch(void* graph, void* stream);
int fq3c_graph_destroy(void* graph);

#ifdef __cplusplus
}
#endif
#endif
