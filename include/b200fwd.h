/* b200fwd -- C ABI of the B200-native forward-pass engine (libb200fwd.so).
 *
 * The reference (githubpradeep/llm_np_cp) has no FFI/plugin boundary: its boundary is the Python call
 * surface  model(input_ids, use_cache, kv_cache) -> (loss, logits, kv_cache, hidden_states, attentions)
 * (llama3.2_model_numpy.py:777-830) built from a config object + a {hf_name: tensor} dict
 * (llama3.2_model_numpy.py:1023-1047).  This header is what a binding for that surface needs; the
 * Python host in llm_np_cp_b200/ re-creates the reference's classes on top of it with ctypes.
 *
 * Conventions: plain pointers and sizes only; every function returns 0 on success or a negative
 * B2_ERR_* code and leaves a message retrievable with b2_last_error() (thread-local).  The engine owns
 * all device memory it allocates; caller-provided host buffers are never retained.  One engine per
 * device; an engine is not thread-safe, different engines may be driven from different host threads
 * (that is the data-parallel story -- no collective is involved).  "dev" pointers are CUDA device
 * pointers on the engine's device (e.g. torch tensors' data_ptr()).
 */
#ifndef B200FWD_H
#define B200FWD_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define B2_OK 0
#define B2_ERR_ARG (-1)      /* bad argument / unsupported shape (the reference raises ValueError) */
#define B2_ERR_CUDA (-2)     /* CUDA runtime / driver error */
#define B2_ERR_STATE (-3)    /* call order (e.g. forward before finalize, missing weight) */
#define B2_ERR_NO_GPU (-4)   /* no usable sm_100 device: there is NO CPU fallback */

#define B2_DTYPE_BF16 0
#define B2_DTYPE_F32 1
#define B2_DTYPE_F16 2

#define B2_ACT_SILU 0        /* llama3.2_model_numpy.py:101-105 */
#define B2_ACT_GELU_TANH 1   /* gemma2_model.py:168-169 */

/* Mirrors the config attributes the reference's constructors read
 * (llama3.2_model_numpy.py:350-357, :158-159, :523, :595-607; gemma2_model.py:434, :658, :867). */
typedef struct b2_config {
  int32_t vocab_size, hidden_size, intermediate_size, num_hidden_layers;
  int32_t num_attention_heads, num_key_value_heads, head_dim;
  int32_t max_positions;        /* length of the RoPE table the engine precomputes */
  float rope_theta, rms_norm_eps;
  int32_t hidden_act;           /* B2_ACT_* */
  int32_t gemma;                /* 1: (1+gamma) norms, 4-norm layers, sqrt(H) embedding scale (gemma2_model.py) */
  float final_logit_softcap;    /* 0 = off; Gemma-2: 30 (gemma2_model.py:867-870) */
  /* Options the reference does NOT implement; 0 = reference behaviour. */
  float attn_logit_softcap;     /* HF Gemma-2: 50 */
  int32_t sliding_window;       /* HF Gemma-2: 4096 on even layers */
  float attn_scale;             /* 0 = 1/sqrt(head_dim) (what the reference does) */
  int32_t always_causal;        /* 0 = reference quirk: no mask when q_len <= 2 (llama3.2_model_numpy.py:479) */
  int32_t precise;              /* 1: parity mode -- every bf16 activation is carried as hi+lo bf16 planes (exact weights,
                                   2x-K weight GEMMs, 3 MMAs per attention product): logits within 1e-2 of the
                                   reference at ~2-3x the cost */
  /* HF "llama3" RoPE frequency scaling (transformers modeling_rope_utils._compute_llama3_parameters); the reference
   * never applies it (llama3.2_model_numpy.py:47 uses plain base^(-2i/HD)).  rope_factor <= 1 = off (reference). */
  float rope_factor;                    /* Llama-3.2: 32 */
  float rope_low_freq_factor;           /* Llama-3.2: 1 */
  float rope_high_freq_factor;          /* Llama-3.2: 4 */
  int32_t rope_original_max_positions;  /* Llama-3.2: 8192 */
  /* Optional tensor parallelism for single long sequences / decode (the reference has no parallelism at all; BASELINE.json
   * north_star): this engine is rank tp_rank of tp_world engines, one per GPU.  Attention heads, FFN columns and
   * vocabulary rows are split evenly; o_proj / down_proj produce partial sums that are summed with ncclAllReduce over
   * NVLink, lm_head logits are all-gathered.  tp_world <= 1 = off.  b2_load_weight still takes the FULL tensors (each
   * engine keeps its slice); b2_tp_init must follow b2_engine_create on every rank. */
  int32_t tp_rank, tp_world;
} b2_config;

typedef struct b2_engine b2_engine;
typedef struct b2_cache b2_cache;

/* thread-local message of the last failing call */
const char* b2_last_error(void);
/* number of CUDA devices with compute capability 10.x (0 => every compute call fails with B2_ERR_NO_GPU) */
int b2_device_count(void);

/* -------- engine lifetime: replaces load_model()'s model construction (llama3.2_model_numpy.py:1030-1047) */
int b2_engine_create(const b2_config* cfg, int device, b2_engine** out);
void b2_engine_destroy(b2_engine* e);
/* Upload one tensor by its HF name (the reference's load_weights(name) seam, :1023-1028).  The engine
 * converts to bf16, concatenates q/k/v and interleaves gate/up into its fused layouts.  `lm_head.weight`
 * is aliased to `model.embed_tokens.weight` exactly like the reference (:1025-1026). */
int b2_load_weight(b2_engine* e, const char* hf_name, const void* host_ptr, const int64_t* shape, int ndim,
                   int dtype);
/* Checks that every tensor arrived and builds the fused layouts; must precede b2_forward. */
int b2_finalize(b2_engine* e);

/* -------- KV cache: replaces KVCache (llama3.2_model_numpy.py:311-340) with a preallocated
 * [L][2][max_batch][NKV][max_tokens][HD] bf16 buffer (post-RoPE K, un-repeated). */
int b2_cache_create(b2_engine* e, int max_batch, int max_tokens, b2_cache** out);
void b2_cache_destroy(b2_cache* c);
int b2_cache_len(const b2_cache* c);   /* KVCache.num_items() */
int b2_cache_reset(b2_cache* c);
/* Paged variant: every (layer, K|V) plane is a pool of n_pages pages of page_size positions ([n_pages][NKV][page][HD]) and a
 * block table [max_batch][ceil(max_tokens / page_size)] maps a sequence's logical pages to physical ones (the layout of
 * vLLM-style paged attention).  page_size: a multiple of 128.  n_pages: 0 = enough for max_tokens
 * of every row; smaller = the pool starts there and doubles on demand (allocation + plane-wise copy), i.e. memory follows
 * the tokens actually cached instead of max_tokens.  Pages are handed out row-interleaved as the sequences grow.
 * b2_cache_len / reset / read / destroy and b2_forward work on both kinds. */
int b2_cache_create_paged(b2_engine* e, int max_batch, int max_tokens, int page_size, int n_pages, b2_cache** out);
/* Replace the block table of an EMPTY paged cache (host int32 [max_batch * max_logical_pages], every entry a distinct page
 * in [0, n_pages)): lets a caller place pages itself (tests use a shuffled permutation). */
int b2_cache_set_block_table(b2_cache* c, const int32_t* table_host, int n_entries);
/* page_size (0 = dense cache), pool size in pages, pages currently handed out */
int b2_cache_info(const b2_cache* c, int32_t* page_size, int32_t* n_pages, int32_t* pages_in_use);
/* copy layer `layer`'s K or V ([batch][NKV][len][HD], bf16) to a device buffer -- KVCache.key_cache[i] */
int b2_cache_read(const b2_cache* c, int layer, int which /*0=K,1=V*/, int batch, void* dev_out);

#define B2_LOGITS_NONE 0
#define B2_LOGITS_ALL 1    /* [B,S,V] fp32 -- what the reference returns (:811) */
#define B2_LOGITS_LAST 2   /* [B,V] fp32, last position only -- what generate() consumes (:899) */
#define B2_LOGITS_ALL_LAST_TO_HOST 3 /* b2_forward_host only: [B,S,V] computed and kept on the device (as the
                                        reference's CuPy twin does), the last position [B,V] copied to the host */

/* -------- forward: LlamaForCausalLM_np.__call__ / Gemma2ForCausalLM_np.__call__
 * ids_dev     int64 [B,S] on the device
 * cache       NULL => use_cache=False (attention over this call's tokens only); else K/V are appended
 *             and cache length advances by S.  A multi-token call on a non-empty cache returns
 *             B2_ERR_ARG like the reference raises (App. B Q4) unless cfg.always_causal is set.
 * logits_dev  fp32 output per logits_mode (may be NULL with B2_LOGITS_NONE)
 * hidden_dev  optional fp32 [L+1,B,S,H]: the reference's hidden_states tuple (:693-720) for parity tests
 * stream      cudaStream_t; NULL is the CUDA default stream, as everywhere in CUDA              */
int b2_forward(b2_engine* e, b2_cache* cache, const int64_t* ids_dev, int B, int S, int logits_mode,
               float* logits_dev, float* hidden_dev, void* stream);
/* Extended form of b2_forward: everything LlamaModel.__call__ accepts on the path (llama3.2_model_numpy.py:624-732).
 * Exactly one of ids / inputs_embeds must be non-NULL (the reference raises ValueError otherwise, :645-652). */
typedef struct b2_forward_args {
  const int64_t* ids;           /* device int64 [B,S], or NULL */
  const float* inputs_embeds;   /* device fp32 [B,S,H] used instead of the embedding gather (:675-676), or NULL */
  const int32_t* position_ids;  /* device int32 [S]: RoPE positions (:657), shared by the batch like the reference's [1,S];
                                   NULL = cache length + s.  K/V are always appended at cache length + s. */
  int32_t B, S, logits_mode;
  float* logits;                /* device, per logits_mode */
  float* hidden;                /* optional device fp32 [L+1,B,S,H] */
  float* attn_probs;            /* optional device fp32 [L,B,NH,S,T], T = cached + S: opt-in debug materialisation of the
                                   reference's `attentions` tuple (:713-714); recomputed per layer by a separate kernel */
  void* stream;
} b2_forward_args;
int b2_forward_ex(b2_engine* e, b2_cache* cache, const b2_forward_args* args);
/* Tensor-parallel communicator.  Rank 0 calls b2_tp_unique_id (ncclGetUniqueId), the caller ships the B2_NCCL_ID_BYTES
 * bytes to the other ranks (e.g. torch.distributed broadcast), every rank then calls b2_tp_init (ncclCommInitRank). */
#define B2_NCCL_ID_BYTES 128
int b2_tp_unique_id(void* id_out);
int b2_tp_init(b2_engine* e, const void* id);
/* Peer-memory (NVLink P2P) path of the tensor-parallel decode step: each rank exports one buffer as a cudaIpc handle
 * (B2_IPC_HANDLE_BYTES bytes), the caller all-gathers the handles in rank order, each rank maps its peers.  Afterwards the
 * few-row Llama path replaces ncclAllReduce + RMSNorm by ONE kernel that barriers through flags in peer memory, sums the
 * ranks' partial planes straight over NVLink in rank order, adds the residual and normalises.  Optional: without these two
 * calls everything goes through NCCL. */
#define B2_IPC_HANDLE_BYTES 64
int b2_tp_ipc_handle(b2_engine* e, void* handle_out);
int b2_tp_set_peers(b2_engine* e, const void* handles_in_rank_order, int n_handles);
/* HF layer_types: is_sliding[i] != 0 => layer i applies cfg.sliding_window (default without this call: even layers,
 * transformers' Gemma-2 default pattern).  No effect while cfg.sliding_window == 0 (the reference never slides). */
int b2_set_layer_types(b2_engine* e, const int32_t* is_sliding, int n_layers);
/* Same call with HOST buffers (pageable or pinned): ids H2D, forward, logits D2H, synchronises. */
int b2_forward_host(b2_engine* e, b2_cache* cache, const int64_t* ids_host, int B, int S, int logits_mode,
                    float* logits_host, float* hidden_host);
/* Replays one-token steps through a captured CUDA graph (decode loop of generate(), :881-909).
 * 1 = enable (default 0). */
int b2_set_decode_graph(b2_engine* e, int enable);

/* -------- sampling: min_p_sampling (llama3.2_model_numpy.py:948-961) on the device.
 * logits_dev [B,ld] fp32; u_dev [B] uniforms in [0,1); out_ids_dev int64 [B];
 * probs_dev optional [B,V] fp32 receives the renormalised distribution (testable part). */
int b2_sample_min_p(b2_engine* e, const float* logits_dev, int B, int ld, float p_base, const float* u_dev,
                    int64_t* out_ids_dev, float* probs_dev, void* stream);

/* greedy pick on the device: out_ids_dev[b] = argmax(logits_dev[b, :V]) (lowest index on ties, like np.argmax) */
int b2_argmax(b2_engine* e, const float* logits_dev, int B, int ld, int64_t* out_ids_dev, void* stream);

/* -------- instrumentation */
typedef struct b2_stats {
  int64_t launches;        /* kernels launched by this library since the last reset */
  int64_t gemm_launches;   /* of which tcgen05 GEMMs */
  double gemm_ms;          /* device time of the GEMM launches (CUDA events), only while timing is on */
  double gemm_flops;       /* algorithmic FLOPs of those launches */
  double attn_ms;
  double attn_flops;
  int64_t attn_launches;
  int64_t timed_forwards;  /* forwards whose launches carried events (see b2_set_timing) */
  int64_t comm_calls;      /* tensor parallelism: NCCL collectives enqueued since the last reset */
  double comm_bytes;       /* payload bytes of those collectives on this rank */
  double comm_ms;          /* their device time, only while timing is on */
  int64_t p2p_fused;       /* tensor parallelism: fused all-reduce + RMSNorm kernels over peer memory launched (host-side count) */
} b2_stats;
/* on = 1: record CUDA events around every GEMM / attention launch; on = k > 1: only on every k-th forward (the other
 * forwards take the plain launch path, so a timed region mostly runs exactly what an untimed one runs); 0 = off */
int b2_set_timing(b2_engine* e, int on);
int b2_get_stats(b2_engine* e, b2_stats* out);    /* synchronises the engine stream when timing is on */
int b2_reset_stats(b2_engine* e);

/* -------- single-kernel entry points (per-kernel parity tests call these through the same ABI) */
int b2_op_rmsnorm(const float* x_dev, const float* g_dev, void* out_bf16_dev, float* out_f32_dev, int M, int H,
                  float eps, void* stream);
int b2_op_add_rmsnorm(float* x_dev, const float* y_dev, const float* g_dev, int M, int H, float eps, void* stream);
/* C = A . W^T with a fused epilogue.  epi: 0 bf16 out, 1 fp32 out (+softcap), 2 fp32 out += (residual),
 * 4 GLU on a gate/up-interleaved W (out bf16 [M,N/2]).  bn: N-tile (32/64/128/256), split_k >= 1 (epi 2 only). */
int b2_op_gemm(const void* a_bf16_dev, const void* w_bf16_dev, void* out_dev, int M, int N, int K, int epi,
               int bn, int split_k, float softcap, int act, void* stream);
/* few-row (M <= 64) cluster split-K GEMM of the decode step with the residual-add + half-RMSNorm epilogue:
 * x[M,N] fp32 += a[M,K] . w[N,K]^T;  a_out[M,N] bf16 = x * gamma;  ss[M,N/32] = sums of x^2 per 32-column block.
 * cluster = CTAs per 64-column output tile, reduced through distributed shared memory (0 = one wave of the device). */
int b2_op_gemm_sk(const void* a_bf16_dev, const void* w_bf16_dev, float* x_dev, void* a_out_bf16_dev, const float* gamma_dev,
                  float* ss_dev, int M, int N, int K, int cluster, void* stream);
/* fused QKV projection + RoPE + KV append: a [B*S,H] bf16, w [(NH+2NKV)*HD, H] bf16 ->
 * q [B*S, NH*HD] bf16, k/v caches [B][NKV][Tmax][HD] bf16 written at positions t0..t0+S-1 */
int b2_op_qkv_rope(const void* a_dev, const void* w_dev, void* q_dev, void* k_cache_dev, void* v_cache_dev, int B,
                   int S, int H, int NH, int NKV, int HD, int Tmax, int t0, float rope_theta, int bn, void* stream);
/* attention over caches holding t0+S tokens; q [B*S, NH*HD] bf16 -> out [B*S, NH*HD] bf16.
 * decode != 0 selects the one-token kernel (requires S == 1). */
int b2_op_attention(const void* q_dev, const void* k_cache_dev, const void* v_cache_dev, void* out_dev, int B, int S,
                    int NH, int NKV, int HD, int Tmax, int t0, float scale, float softcap, int causal, int window,
                    int decode, void* stream);
/* rows of gate/up interleaved the way b2_op_gemm(epi=4) expects: out[2I,H] from gate[I,H], up[I,H] (host) */
int b2_interleave_gate_up(const uint16_t* gate, const uint16_t* up, uint16_t* out, int I, int H);

#ifdef __cplusplus
}
#endif
#endif /* B200FWD_H */
