// Memory-bound kernels of the per-layer path: embedding gather, RMSNorm (warp-shuffle reduction),
// Gemma's post-norm + residual, last-row gather, min-p sampling.  All are HBM-bound: 16-byte
// vector accesses, one warp per row, fp32 statistics.
#pragma once
#include "common.cuh"

namespace b2 {

// x[m, :] = table[ids[m], :] * scale      (llama3.2_model_numpy.py:675; gemma2_model.py:738-739)
__global__ void __launch_bounds__(256)
embed_kernel(const long long* __restrict__ ids, const __nv_bfloat16* __restrict__ table, float* __restrict__ x,
             int M, int H, int V, float scale) {
  pdl_trigger();
  pdl_wait();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int row = blockIdx.x * 8 + warp;
  if (row >= M) return;
  long long id = ids[row];
  if (id < 0) id = 0;
  if (id >= V) id = V - 1;   // NumPy would raise IndexError; the host wrapper validates ids
  const uint4* src = reinterpret_cast<const uint4*>(table + static_cast<size_t>(id) * H);
  float4* dst = reinterpret_cast<float4*>(x + static_cast<size_t>(row) * H);
  for (int i = lane; i < H / 8; i += 32) {
    const uint4 u = __ldg(src + i);
    const __nv_bfloat162* h2 = reinterpret_cast<const __nv_bfloat162*>(&u);
    const float2 a = __bfloat1622float2(h2[0]), b = __bfloat1622float2(h2[1]);
    const float2 c = __bfloat1622float2(h2[2]), d = __bfloat1622float2(h2[3]);
    dst[2 * i] = make_float4(a.x * scale, a.y * scale, b.x * scale, b.y * scale);
    dst[2 * i + 1] = make_float4(c.x * scale, c.y * scale, d.x * scale, d.y * scale);
  }
}

// x = inputs_embeds * scale (caller-provided embeddings instead of token ids, llama3.2_model_numpy.py:645-676;
// Gemma-2 multiplies by sqrt(H) afterwards, gemma2_model.py:738-739).  n4 = number of float4.
__global__ void __launch_bounds__(256)
scale_copy_kernel(const float* __restrict__ src, float* __restrict__ dst, size_t n4, float scale) {
  pdl_trigger();
  pdl_wait();
  const float4* s4 = reinterpret_cast<const float4*>(src);
  float4* d4 = reinterpret_cast<float4*>(dst);
  for (size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < n4; i += static_cast<size_t>(gridDim.x) * blockDim.x) {
    const float4 v = s4[i];
    d4[i] = make_float4(v.x * scale, v.y * scale, v.z * scale, v.w * scale);
  }
}

// Opt-in debug path: materialises one layer's attention probabilities [B,NH,S,T] fp32 -- the `attentions` tuple the
// reference returns by default (llama3.2_model_numpy.py:501, :713-714) and the fused attention kernels never form.
// Recomputed from this layer's post-RoPE q and its K cache (hi planes): one warp per (batch, head, query) row, three
// passes over the row (scores, exp, normalise).  Not on the product path; parity tests compare it with the goldens.
__global__ void __launch_bounds__(256)
attn_probs_kernel(const __nv_bfloat16* __restrict__ q, int ldq, const __nv_bfloat16* __restrict__ kc, float* __restrict__ out,
                  int B, int S, int NH, int NKV, int HD, int Tmax, const int* __restrict__ t0_ptr, float scale, float softcap,
                  int causal, int window, const KvPages pages) {
  pdl_trigger();
  pdl_wait();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const size_t row = static_cast<size_t>(blockIdx.x) * 8 + warp;
  if (row >= static_cast<size_t>(B) * NH * S) return;
  const int s = static_cast<int>(row % S), h = static_cast<int>((row / S) % NH), b = static_cast<int>(row / (static_cast<size_t>(S) * NH));
  const int t0 = *t0_ptr, T = t0 + S, qpos = t0 + s;
  const int lo = window > 0 ? max(0, qpos - window + 1) : 0;
  const int hi = causal ? qpos : T - 1;
  const uint4* q4 = reinterpret_cast<const uint4*>(q + (static_cast<size_t>(b) * S + s) * ldq + h * HD);
  const int kvh = h / (NH / NKV);
  float* o = out + row * T;
  float m = -INFINITY;
  for (int t = lane; t < T; t += 32) {
    float v = -INFINITY;
    if (t >= lo && t <= hi) {
      const uint4* k4 = reinterpret_cast<const uint4*>(kc + kv_row(pages, b, kvh, NKV, Tmax, t) * HD);
      float acc = 0.f;
      for (int i = 0; i < HD / 8; ++i) {
        const uint4 a = __ldg(q4 + i), c = __ldg(k4 + i);
        const __nv_bfloat162* a2 = reinterpret_cast<const __nv_bfloat162*>(&a);
        const __nv_bfloat162* c2 = reinterpret_cast<const __nv_bfloat162*>(&c);
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const float2 x = __bfloat1622float2(a2[e]), y = __bfloat1622float2(c2[e]);
          acc = fmaf(x.x, y.x, acc);
          acc = fmaf(x.y, y.y, acc);
        }
      }
      v = acc * scale;
      if (softcap > 0.f) v = softcap * tanhf(v / softcap);
    }
    o[t] = v;
    m = fmaxf(m, v);
  }
  m = warp_max(m);
  float sum = 0.f;
  for (int t = lane; t < T; t += 32) {
    const float v = o[t];
    const float e = (v == -INFINITY) ? 0.f : expf(v - m);
    o[t] = e;
    sum += e;
  }
  sum = warp_sum(sum);
  const float inv = sum > 0.f ? 1.0f / sum : 0.f;
  for (int t = lane; t < T; t += 32) o[t] *= inv;
}

// Decode step with the RMSNorm split across the GEMMs (gemm_sk_tcgen05.cuh): x = table[id] * scale, a' = bf16(x * gamma)
// (the first GEMM's A operand), ss[row][0] = sum x^2 and ss[row][1..ss_n) = 0.  One CTA per row.
__global__ void __launch_bounds__(256)
embed_norm_row_kernel(const long long* __restrict__ ids, const __nv_bfloat16* __restrict__ table, float* __restrict__ x,
                      __nv_bfloat16* __restrict__ a_out, const float* __restrict__ gamma, float* __restrict__ ss, int ss_n, int H,
                      int V, float scale) {
  pdl_trigger();
  pdl_wait();
  __shared__ float red[8];
  const int row = blockIdx.x, tid = threadIdx.x;
  long long id = ids[row];
  if (id < 0) id = 0;
  if (id >= V) id = V - 1;
  const uint2* src = reinterpret_cast<const uint2*>(table + static_cast<size_t>(id) * H);
  float4* xr = reinterpret_cast<float4*>(x + static_cast<size_t>(row) * H);
  const float4* g4 = reinterpret_cast<const float4*>(gamma);
  float sq = 0.f;
  for (int i = tid; i < H / 4; i += 256) {
    const uint2 u = __ldg(src + i);
    const float2 a = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&u.x));
    const float2 b = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&u.y));
    const float4 v = make_float4(a.x * scale, a.y * scale, b.x * scale, b.y * scale);
    xr[i] = v;
    sq += (v.x * v.x + v.y * v.y) + (v.z * v.z + v.w * v.w);
    const float4 gg = __ldg(g4 + i);
    *reinterpret_cast<uint2*>(a_out + static_cast<size_t>(row) * H + i * 4) =
        make_uint2(pack_bf16(v.x * gg.x, v.y * gg.y), pack_bf16(v.z * gg.z, v.w * gg.w));
  }
  sq = warp_sum(sq);
  if ((tid & 31) == 0) red[tid >> 5] = sq;
  __syncthreads();
  float* srow = ss + static_cast<size_t>(row) * ss_n;
  if (tid == 0) {
    float t = 0.f;
#pragma unroll
    for (int w = 0; w < 8; ++w) t += red[w];
    srow[0] = t;
  }
  for (int i = 1 + tid; i < ss_n; i += 256) srow[i] = 0.f;
}

constexpr int NORM_MAXV = 24;  // float4 per lane: supports H <= 3072

// out = x * 1/sqrt(mean(x^2)+eps) * g   (LlamaRMSNorm_np, llama3.2_model_numpy.py:256-278; g already holds
// (1 + gamma) for Gemma-2, gemma2_model.py:334).  One warp per row; the row stays in registers between the
// reduction and the scaling pass.  Writes bf16 (the next GEMM's A operand; optional "lo" plane at
// +lo_off columns for the precise mode) and/or fp32 (hidden-state capture / final norm).
// `add` (optional): x[row] += add[row] first, written back -- the all-reduced partial sums of a row-parallel o_proj /
// down_proj under tensor parallelism (the residual add of llama3.2_model_numpy.py:569, :578).
__global__ void __launch_bounds__(256)
rmsnorm_kernel(float* __restrict__ x, const float* __restrict__ add, const float* __restrict__ g,
               __nv_bfloat16* __restrict__ out_bf16, int ld_out, int lo_off, float* __restrict__ out_f32, int M, int H, float eps) {
  pdl_trigger();
  pdl_wait();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  // rows are walked from the END of the residual stream: the GEMM in front of this kernel wrote x in ascending row
  // order and x (134 MB at 16384 tokens) is about the size of L2, so its last rows are the ones still cached
  const int row = M - 1 - (blockIdx.x * 8 + warp);
  if (row < 0) return;
  const int nv = H >> 7;  // float4 per lane
  float4* xr = reinterpret_cast<float4*>(x + static_cast<size_t>(row) * H);
  const float4* ar = add ? reinterpret_cast<const float4*>(add + static_cast<size_t>(row) * H) : nullptr;
  float4 v[NORM_MAXV];
  float ss = 0.f;
#pragma unroll
  for (int i = 0; i < NORM_MAXV; ++i) {
    if (i < nv) {
      v[i] = xr[i * 32 + lane];
      if (ar) {
        const float4 a4 = ar[i * 32 + lane];
        v[i].x += a4.x; v[i].y += a4.y; v[i].z += a4.z; v[i].w += a4.w;
        xr[i * 32 + lane] = v[i];
      }
      ss += v[i].x * v[i].x + v[i].y * v[i].y + v[i].z * v[i].z + v[i].w * v[i].w;
    }
  }
  ss = warp_sum(ss);
  const float rs = 1.0f / sqrtf(ss / static_cast<float>(H) + eps);
  const float4* gr = reinterpret_cast<const float4*>(g);
#pragma unroll
  for (int i = 0; i < NORM_MAXV; ++i) {
    if (i < nv) {
      const float4 gg = __ldg(gr + i * 32 + lane);
      const float4 o = make_float4(v[i].x * rs * gg.x, v[i].y * rs * gg.y, v[i].z * rs * gg.z, v[i].w * rs * gg.w);
      const int col = (i * 32 + lane) * 4;
      if (out_bf16) {
        __nv_bfloat16* d = out_bf16 + static_cast<size_t>(row) * ld_out + col;
        *reinterpret_cast<uint2*>(d) = make_uint2(pack_bf16(o.x, o.y), pack_bf16(o.z, o.w));
        if (lo_off) {
          const float lx = o.x - __bfloat162float(__float2bfloat16_rn(o.x));
          const float ly = o.y - __bfloat162float(__float2bfloat16_rn(o.y));
          const float lz = o.z - __bfloat162float(__float2bfloat16_rn(o.z));
          const float lw = o.w - __bfloat162float(__float2bfloat16_rn(o.w));
          *reinterpret_cast<uint2*>(d + lo_off) = make_uint2(pack_bf16(lx, ly), pack_bf16(lz, lw));
        }
      }
      if (out_f32) reinterpret_cast<float4*>(out_f32 + static_cast<size_t>(row) * H)[i * 32 + lane] = o;
    }
  }
}

// Few-row variant (decode, M <= 128): one 256-thread CTA per row so the row's loads (and the split-K partials of the
// preceding o_proj / down_proj GEMM: x[row] += sum_s partials[s][row], fixed order => deterministic)
// are all in flight at once instead of serialised inside a single warp.  Same arithmetic and summation order of
// the partials as rmsnorm_kernel.
__global__ void __launch_bounds__(256)
rmsnorm_row_kernel(float* __restrict__ x, const float* __restrict__ g, __nv_bfloat16* __restrict__ out_bf16,
                   int ld_out, int lo_off, float* __restrict__ out_f32, int H, float eps,
                   const float* __restrict__ partials, int nsplit, size_t split_stride) {
  pdl_trigger();
  pdl_wait();
  __shared__ float red[8];
  const int row = blockIdx.x, tid = threadIdx.x;
  const int nv4 = H >> 2;                       // float4 per row (<= 768)
  float4* xr = reinterpret_cast<float4*>(x + static_cast<size_t>(row) * H);
  float4 v[3];
  float ss = 0.f;
#pragma unroll
  for (int i = 0; i < 3; ++i) {
    const int idx = tid + i * 256;
    if (idx < nv4) {
      v[i] = xr[idx];
      if (nsplit > 0) {
        float4 pp[8];
#pragma unroll
        for (int s = 0; s < 8; ++s)
          if (s < nsplit) pp[s] = reinterpret_cast<const float4*>(partials + s * split_stride + static_cast<size_t>(row) * H)[idx];
#pragma unroll
        for (int s = 0; s < 8; ++s)
          if (s < nsplit) {
            v[i].x += pp[s].x;
            v[i].y += pp[s].y;
            v[i].z += pp[s].z;
            v[i].w += pp[s].w;
          }
        xr[idx] = v[i];
      }
      ss += v[i].x * v[i].x + v[i].y * v[i].y + v[i].z * v[i].z + v[i].w * v[i].w;
    }
  }
  ss = warp_sum(ss);
  if ((tid & 31) == 0) red[tid >> 5] = ss;
  __syncthreads();
  float tot = 0.f;
#pragma unroll
  for (int w = 0; w < 8; ++w) tot += red[w];     // same order in every thread
  const float rs = 1.0f / sqrtf(tot / static_cast<float>(H) + eps);
  const float4* gr = reinterpret_cast<const float4*>(g);
#pragma unroll
  for (int i = 0; i < 3; ++i) {
    const int idx = tid + i * 256;
    if (idx < nv4) {
      const float4 gg = __ldg(gr + idx);
      const float4 o = make_float4(v[i].x * rs * gg.x, v[i].y * rs * gg.y, v[i].z * rs * gg.z, v[i].w * rs * gg.w);
      if (out_bf16) {
        __nv_bfloat16* d = out_bf16 + static_cast<size_t>(row) * ld_out + idx * 4;
        *reinterpret_cast<uint2*>(d) = make_uint2(pack_bf16(o.x, o.y), pack_bf16(o.z, o.w));
        if (lo_off) {
          const float lx = o.x - __bfloat162float(__float2bfloat16_rn(o.x));
          const float ly = o.y - __bfloat162float(__float2bfloat16_rn(o.y));
          const float lz = o.z - __bfloat162float(__float2bfloat16_rn(o.z));
          const float lw = o.w - __bfloat162float(__float2bfloat16_rn(o.w));
          *reinterpret_cast<uint2*>(d + lo_off) = make_uint2(pack_bf16(lx, ly), pack_bf16(lz, lw));
        }
      }
      if (out_f32) reinterpret_cast<float4*>(out_f32 + static_cast<size_t>(row) * H)[idx] = o;
    }
  }
}

// Gemma-2 post-norms: x += rmsnorm(y) * g   (gemma2_model.py:633-637, :641-643)
__global__ void __launch_bounds__(256)
add_rmsnorm_kernel(float* __restrict__ x, const float* __restrict__ y, const float* __restrict__ g, int M, int H,
                   float eps, int nsplit, size_t split_stride) {
  pdl_trigger();
  pdl_wait();   // nsplit > 1: y = sum of split-K partials
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int row = blockIdx.x * 8 + warp;
  if (row >= M) return;
  const int nv = H >> 7;
  const float4* yr = reinterpret_cast<const float4*>(y + static_cast<size_t>(row) * H);
  float4* xr = reinterpret_cast<float4*>(x + static_cast<size_t>(row) * H);
  float4 v[NORM_MAXV];
  float ss = 0.f;
#pragma unroll
  for (int i = 0; i < NORM_MAXV; ++i) {
    if (i < nv) {
      v[i] = yr[i * 32 + lane];
      for (int s = 1; s < nsplit; ++s) {
        const float4 pp = reinterpret_cast<const float4*>(y + s * split_stride + static_cast<size_t>(row) * H)[i * 32 + lane];
        v[i].x += pp.x;
        v[i].y += pp.y;
        v[i].z += pp.z;
        v[i].w += pp.w;
      }
      ss += v[i].x * v[i].x + v[i].y * v[i].y + v[i].z * v[i].z + v[i].w * v[i].w;
    }
  }
  ss = warp_sum(ss);
  const float rs = 1.0f / sqrtf(ss / static_cast<float>(H) + eps);
  const float4* gr = reinterpret_cast<const float4*>(g);
#pragma unroll
  for (int i = 0; i < NORM_MAXV; ++i) {
    if (i < nv) {
      const float4 gg = __ldg(gr + i * 32 + lane);
      float4 o = xr[i * 32 + lane];
      o.x += v[i].x * rs * gg.x;
      o.y += v[i].y * rs * gg.y;
      o.z += v[i].z * rs * gg.z;
      o.w += v[i].w * rs * gg.w;
      xr[i * 32 + lane] = o;
    }
  }
}

// dst[b, :] = src[b*S + S-1, :]  (last position of every sequence; feeds last-token-only logits)
__global__ void __launch_bounds__(256)
gather_last_kernel(const float* __restrict__ src, float* __restrict__ dst, int B, int S, int H) {
  pdl_trigger();
  pdl_wait();
  const int b = blockIdx.x;
  const float4* s = reinterpret_cast<const float4*>(src + (static_cast<size_t>(b) * S + S - 1) * H);
  float4* d = reinterpret_cast<float4*>(dst + static_cast<size_t>(b) * H);
  for (int i = threadIdx.x; i < H / 4; i += blockDim.x) d[i] = s[i];
}

__global__ void add_scalar_kernel(int* p, int v) {
  pdl_trigger();
  pdl_wait();
  if (threadIdx.x == 0 && blockIdx.x == 0) *p += v;
}

// greedy pick: out[b] = argmax_v logits[b, v] (lowest index on ties, like np.argmax); one CTA per row
__global__ void __launch_bounds__(1024)
argmax_kernel(const float* __restrict__ logits, int ld, int V, long long* __restrict__ out) {
  pdl_trigger();
  pdl_wait();
  __shared__ float sv[32];
  __shared__ int si[32];
  const int b = blockIdx.x, tid = threadIdx.x;
  const float4* x4 = reinterpret_cast<const float4*>(logits + static_cast<size_t>(b) * ld);
  float best = -INFINITY;
  int bi = 0x7fffffff;
  for (int i = tid; i < V / 4; i += 1024) {
    const float4 v = x4[i];
    const float e[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
    for (int k = 0; k < 4; ++k)
      if (e[k] > best) {
        best = e[k];
        bi = i * 4 + k;
      }
  }
  for (int i = (V / 4) * 4 + tid; i < V; i += 1024) {
    const float v = logits[static_cast<size_t>(b) * ld + i];
    if (v > best) {
      best = v;
      bi = i;
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const float ov = __shfl_xor_sync(0xffffffffu, best, o);
    const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
    if (ov > best || (ov == best && oi < bi)) {
      best = ov;
      bi = oi;
    }
  }
  if ((tid & 31) == 0) {
    sv[tid >> 5] = best;
    si[tid >> 5] = bi;
  }
  __syncthreads();
  if (tid < 32) {
    best = sv[tid];
    bi = si[tid];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const float ov = __shfl_xor_sync(0xffffffffu, best, o);
      const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
      if (ov > best || (ov == best && oi < bi)) {
        best = ov;
        bi = oi;
      }
    }
    if (tid == 0) out[b] = bi;
  }
}

// ---------------------------------------------------------------------------------------------
// min-p sampling on the device (min_p_sampling, llama3.2_model_numpy.py:948-961):
//   p = softmax(logits); keep p >= p_base * max(p); renormalise; draw one index.
// The reference draws with an unseeded torch.multinomial, so parity is distributional; here the draw
// is an inverse-CDF lookup (in vocabulary order) of a caller-supplied uniform u[b].
// One CTA (1024 threads) per row.  probs_out (optional) receives the renormalised distribution.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ float block_reduce(float v, float* sh, bool is_max) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  v = is_max ? warp_max(v) : warp_sum(v);
  __syncthreads();
  if (lane == 0) sh[warp] = v;
  __syncthreads();
  float r = (threadIdx.x < (blockDim.x >> 5)) ? sh[threadIdx.x] : (is_max ? -INFINITY : 0.f);
  if (warp == 0) {
    r = is_max ? warp_max(r) : warp_sum(r);
    if (lane == 0) sh[32] = r;
  }
  __syncthreads();
  return sh[32];
}

__global__ void __launch_bounds__(1024)
min_p_sample_kernel(const float* __restrict__ logits, int ld, int V, float p_base, const float* __restrict__ u,
                    long long* __restrict__ out_ids, float* __restrict__ probs_out) {
  pdl_trigger();
  pdl_wait();
  __shared__ float sh[33];
  __shared__ float seg[1024];
  __shared__ int pick_seg;
  const int b = blockIdx.x, tid = threadIdx.x;
  const float* x = logits + static_cast<size_t>(b) * ld;
  // contiguous segment per thread so the CDF runs in vocabulary order
  const int per = (V + blockDim.x - 1) / blockDim.x;
  const int lo = min(V, tid * per), hi = min(V, lo + per);
  float mx = -INFINITY;
  for (int i = lo; i < hi; ++i) mx = fmaxf(mx, x[i]);
  mx = block_reduce(mx, sh, true);
  // p_i >= p_base * p_max  <=>  x_i >= mx + log(p_base)
  const float thr = mx + logf(p_base);
  float s = 0.f;
  for (int i = lo; i < hi; ++i) {
    const float xi = x[i];
    if (xi >= thr) s += expf(xi - mx);
  }
  seg[tid] = s;
  const float total = block_reduce(s, sh, false);
  if (probs_out) {
    float* po = probs_out + static_cast<size_t>(b) * V;
    for (int i = lo; i < hi; ++i) {
      const float xi = x[i];
      po[i] = (xi >= thr) ? expf(xi - mx) / total : 0.f;
    }
  }
  __syncthreads();
  const float target = u[b] * total;
  if (tid == 0) {
    float c = 0.f;
    int sgm = blockDim.x - 1;
    for (int t = 0; t < static_cast<int>(blockDim.x); ++t) {
      if (c + seg[t] > target) {
        sgm = t;
        break;
      }
      c += seg[t];
    }
    // fall back to the last non-empty segment on round-off
    while (sgm > 0 && seg[sgm] == 0.f) --sgm;
    pick_seg = sgm;
    seg[0] = c;  // cumulative mass before the picked segment (seg[0] no longer needed)
  }
  __syncthreads();
  if (tid == pick_seg) {
    float c = (tid == 0) ? 0.f : seg[0];
    if (tid == 0) c = 0.f;
    int pick = -1, last_kept = -1;
    for (int i = lo; i < hi; ++i) {
      const float xi = x[i];
      if (xi >= thr) {
        last_kept = i;
        c += expf(xi - mx);
        if (c > target) {
          pick = i;
          break;
        }
      }
    }
    out_ids[b] = pick >= 0 ? pick : last_kept;
  }
}

}  // namespace b2

namespace b2 {
// ---------------------------------------------------------------------------------------------
// Tensor-parallel decode: all-reduce + residual add + RMSNorm in ONE kernel over peer memory (NVLink P2P).
// Every rank's row-parallel o_proj / down_proj GEMM leaves its fp32 partial planes in a buffer that the other ranks have
// mapped (cudaIpc).  This kernel replaces ncclAllReduce + rmsnorm_row_kernel:
//   1. cross-GPU barrier through flags in peer memory: block 0 tells every peer "my partials of this round are written"
//      (st.release.sys into the peer's flag array), every block waits until all peers have said the same (ld.acquire.sys
//      on its own flags); rounds count up, nothing is ever reset;
//   2. x[row] += sum over split planes (plane order) of sum over ranks (rank order) -- peers' planes are read straight
//      over NVLink with volatile loads; every rank adds the same numbers in the same order, so the replicas of x stay
//      bit-identical;
//   3. the usual RMSNorm of the row.
// Two buffers ("sites": 0 after o_proj, 1 after down_proj) alternate, which is what makes overwriting a buffer safe: a
// rank re-writes site 0 only after it has passed the site-1 barrier of the same layer, i.e. after every peer has entered
// its own post-down_proj kernel and therefore finished reading site 0.
// ---------------------------------------------------------------------------------------------
constexpr int TP_MAX_WORLD = 8;
struct TpPeers {
  const float* part[TP_MAX_WORLD];   // this site's partial planes of rank p (own buffer for p == rank)
  unsigned* flags[TP_MAX_WORLD];     // rank p's flag array [2 sites][TP_MAX_WORLD]
  int rank, world;
};
__device__ __forceinline__ void st_release_sys(unsigned* p, unsigned v) {
  asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ unsigned ld_acquire_sys(const unsigned* p) {
  unsigned v;
  asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ float4 ld_volatile_f4(const float4* p) {
  float4 v;
  asm volatile("ld.volatile.global.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p) : "memory");
  return v;
}

__global__ void __launch_bounds__(256)
tp_allreduce_norm_row_kernel(float* __restrict__ x, const float* __restrict__ g, __nv_bfloat16* __restrict__ out_bf16, int ld_out,
                             float* __restrict__ out_f32, int H, float eps, const TpPeers peers, int nsplit, size_t split_stride,
                             unsigned* __restrict__ round_ctr, unsigned* __restrict__ done_ctr, int site) {
  pdl_trigger();
  pdl_wait();                                  // this rank's GEMM is complete: its partial planes are in this GPU's memory
  __shared__ float red[8];
  const int row = blockIdx.x, tid = threadIdx.x;
  if (tid == 0) {
    const unsigned round = *reinterpret_cast<volatile unsigned*>(round_ctr + site) + 1;   // same value in every block (bumped by the last block out)
    if (blockIdx.x == 0) {
      __threadfence_system();
      for (int p = 0; p < peers.world; ++p)
        if (p != peers.rank) st_release_sys(peers.flags[p] + site * TP_MAX_WORLD + peers.rank, round);
    }
    for (int p = 0; p < peers.world; ++p) {
      if (p == peers.rank) continue;
      unsigned spins = 0;
      while (ld_acquire_sys(peers.flags[peers.rank] + site * TP_MAX_WORLD + p) < round) {
        __nanosleep(64);
        if (++spins > (1u << 25)) {            // seconds: a peer never arrived -- fail loudly instead of hanging the GPU
          printf("b2: tensor-parallel barrier timeout (rank %d waiting for rank %d, site %d, round %u)\n", peers.rank, p, site, round);
          __trap();
        }
      }
    }
    red[0] = __uint_as_float(round);
  }
  __syncthreads();
  const unsigned round = __float_as_uint(red[0]);
  __syncthreads();
  const int nv4 = H >> 2;
  float4* xr = reinterpret_cast<float4*>(x + static_cast<size_t>(row) * H);
  float4 v[3];
  float ss = 0.f;
#pragma unroll
  for (int i = 0; i < 3; ++i) {
    const int idx = tid + i * 256;
    if (idx < nv4) {
      v[i] = xr[idx];
      for (int s = 0; s < nsplit; ++s) {
        // all ranks' loads of this plane are in flight together (one NVLink round trip, not `world` dependent ones);
        // the sum runs in rank order afterwards -- the same order on every rank
        float4 a[TP_MAX_WORLD];
#pragma unroll
        for (int p = 0; p < TP_MAX_WORLD; ++p)
          if (p < peers.world) {
            const float4* src = reinterpret_cast<const float4*>(peers.part[p] + s * split_stride + static_cast<size_t>(row) * H) + idx;
            a[p] = (p == peers.rank) ? *src : ld_volatile_f4(src);
          }
#pragma unroll
        for (int p = 0; p < TP_MAX_WORLD; ++p)
          if (p < peers.world) { v[i].x += a[p].x; v[i].y += a[p].y; v[i].z += a[p].z; v[i].w += a[p].w; }
      }
      xr[idx] = v[i];
      ss += v[i].x * v[i].x + v[i].y * v[i].y + v[i].z * v[i].z + v[i].w * v[i].w;
    }
  }
  ss = warp_sum(ss);
  if ((tid & 31) == 0) red[tid >> 5] = ss;
  __syncthreads();
  float tot = 0.f;
#pragma unroll
  for (int w = 0; w < 8; ++w) tot += red[w];
  const float rs = 1.0f / sqrtf(tot / static_cast<float>(H) + eps);
  const float4* gr = reinterpret_cast<const float4*>(g);
#pragma unroll
  for (int i = 0; i < 3; ++i) {
    const int idx = tid + i * 256;
    if (idx < nv4) {
      const float4 gg = __ldg(gr + idx);
      const float4 o = make_float4(v[i].x * rs * gg.x, v[i].y * rs * gg.y, v[i].z * rs * gg.z, v[i].w * rs * gg.w);
      if (out_bf16)
        *reinterpret_cast<uint2*>(out_bf16 + static_cast<size_t>(row) * ld_out + idx * 4) = make_uint2(pack_bf16(o.x, o.y), pack_bf16(o.z, o.w));
      if (out_f32) reinterpret_cast<float4*>(out_f32 + static_cast<size_t>(row) * H)[idx] = o;
    }
  }
  __syncthreads();
  if (tid == 0) {                              // last block out publishes the round for the next use of this site
    __threadfence();
    if (atomicAdd(done_ctr + site, 1u) == gridDim.x - 1) {
      done_ctr[site] = 0;
      __threadfence();
      *reinterpret_cast<volatile unsigned*>(round_ctr + site) = round;
    }
  }
}

// vocabulary-parallel lm_head (tensor parallelism): ncclAllGather leaves [world][rows][V_local]; the caller's logits are
// [rows][world * V_local].  One float4 per thread, coalesced on both sides.
__global__ void __launch_bounds__(256)
tp_unshard_logits_kernel(const float* __restrict__ src, float* __restrict__ dst, int rows, int v_local, int world) {
  pdl_trigger();
  pdl_wait();
  const size_t vl4 = static_cast<size_t>(v_local) / 4, n4 = static_cast<size_t>(rows) * world * vl4;
  const float4* s4 = reinterpret_cast<const float4*>(src);
  float4* d4 = reinterpret_cast<float4*>(dst);
  for (size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < n4; i += static_cast<size_t>(gridDim.x) * blockDim.x) {
    const size_t c = i % vl4, r = (i / vl4) % rows, w = i / (vl4 * rows);
    d4[(r * world + w) * vl4 + c] = s4[i];
  }
}
}  // namespace b2
