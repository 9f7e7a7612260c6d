// NCSN++ memory-bound and CUDA-core kernels (NHWC activations, fp32 math):
//   layout conversion, time embedding, GroupNorm(+SiLU), FIR resampling (upfirdn2d),
//   a generic fp32 implicit-GEMM convolution (reference-faithful path + odd shapes),
//   and a single-head attention core.
//
// Reference lines (mandt-lab/PSLD, main/models/score_fn/song_sde/):
//   GaussianFourierProjection ........ layerspp.py:32-41 ; temb MLP ncsnpp.py:292-311
//   Dense_0(SiLU(temb)) .............. layerspp.py:262-263
//   GroupNorm(min(C/4,32), eps=1e-6) . layerspp.py:219,231,67-68 ; ncsnpp.py:424-430
//   upfirdn2d ........................ op/upfirdn2d.py:159-200, op/upfirdn2d_kernel.cu:107-207
//   conv3x3 / conv1x1 / NIN .......... layers.py:85-109,531-540
//   AttnBlockpp core ................. layerspp.py:82-86

#include "common.cuh"

namespace psld {

// ======================================================================== layout
// NCHW fp32 -> NHWC T with the channel axis zero-padded to CP >= C (CP = 64 lets the 6-channel
// network input go through the tensor-core convolution, whose K chunks are 64 channels)
template <typename T>
__global__ void __launch_bounds__(256)
nchw_to_nhwc_kernel(const float* __restrict__ in, T* __restrict__ out, int N, int C, int HW,
                    int CP, int CW) {
  pdl_trigger_early();
  pdl_wait();
  // only channels [0, CW) of the CP-wide rows are written (CW < CP: the rest keeps the zeros the
  // buffer was allocated with)
  const int64_t total = (int64_t)N * CW * HW;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (int64_t)gridDim.x * blockDim.x) {
    const int c = (int)(i % CW);
    const int64_t r = i / CW;
    const int p = (int)(r % HW);
    const int n = (int)(r / HW);
    st_elt<T>(out + r * (CP * Elt<T>::kMul) + c, c < C ? in[((int64_t)n * C + c) * HW + p] : 0.f, CP);
  }
}

template <typename T>
__global__ void __launch_bounds__(256)
nhwc_to_nchw_kernel(const T* __restrict__ in, float* __restrict__ out, int N, int C, int HW) {
  pdl_trigger_early();
  pdl_wait();
  const int64_t total = (int64_t)N * C * HW;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (int64_t)gridDim.x * blockDim.x) {
    const int p = (int)(i % HW);
    const int64_t r = i / HW;
    const int c = (int)(r % C);
    const int n = (int)(r / C);
    out[i] = ld_elt<T>(in + ((int64_t)n * HW + p) * (C * Elt<T>::kMul) + c, C);
  }
}

static inline int ew_grid(int64_t total, int per_thread = 1) {
  int64_t g = ceil_div(total, 256LL * per_thread);
  const int64_t cap = 148 * 16;
  return (int)(g < 1 ? 1 : (g > cap ? cap : g));
}

int run_layout(const psld_op& op, cudaStream_t s) {
  const int N = op.i[PSLD_LAYOUT_N], C = op.i[PSLD_LAYOUT_C], HW = op.i[PSLD_LAYOUT_HW];
  const int dir = op.i[PSLD_LAYOUT_DIR], dt = op.i[PSLD_LAYOUT_DTYPE];
  PSLD_CHECK_ARG(N > 0 && C > 0 && HW > 0 && op.in[0] && op.out[0], "layout: bad arguments");
  int CP = op.i[PSLD_LAYOUT_CPAD];
  if (CP <= 0) CP = C;
  PSLD_CHECK_ARG(CP >= C && (dir == 0 || CP == C), "layout: bad channel padding");
  int CW = op.i[PSLD_LAYOUT_CWRITE];
  if (CW <= 0) CW = CP;
  PSLD_CHECK_ARG(CW >= C && CW <= CP, "layout: bad written-channel count");
  const int grid = ew_grid((int64_t)N * (dir == 0 ? CW : CP) * HW);
  if (dir == 0) {
    if (dt == PSLD_BF16)
      launch_pdl(nchw_to_nhwc_kernel<__nv_bfloat16>, dim3(grid), dim3(256), 0, s, 1, (const float*)op.in[0],
                                                            (__nv_bfloat16*)op.out[0], N, C, HW, CP, CW);
    else if (dt == PSLD_BF16S)
      launch_pdl(nchw_to_nhwc_kernel<bf16s>, dim3(grid), dim3(256), 0, s, 1, (const float*)op.in[0],
                 (bf16s*)op.out[0], N, C, HW, CP, CW);
    else
      launch_pdl(nchw_to_nhwc_kernel<float>, dim3(grid), dim3(256), 0, s, 1, (const float*)op.in[0], (float*)op.out[0],
                                                    N, C, HW, CP, CW);
  } else {
    if (dt == PSLD_BF16)
      launch_pdl(nhwc_to_nchw_kernel<__nv_bfloat16>, dim3(grid), dim3(256), 0, s, 1, (const __nv_bfloat16*)op.in[0],
                                                            (float*)op.out[0], N, C, HW);
    else if (dt == PSLD_BF16S)
      launch_pdl(nhwc_to_nchw_kernel<bf16s>, dim3(grid), dim3(256), 0, s, 1, (const bf16s*)op.in[0],
                 (float*)op.out[0], N, C, HW);
    else
      launch_pdl(nhwc_to_nchw_kernel<float>, dim3(grid), dim3(256), 0, s, 1, (const float*)op.in[0], (float*)op.out[0],
                                                    N, C, HW);
  }
  PSLD_CHECK_LAUNCH();
  return PSLD_OK;
}

// ======================================================================== a x + b y
// Classifier-free guidance: eps = (1 + w) eps_cond - w eps_uncond over the fp32 NCHW network
// outputs (and, with y == nullptr, the copy of the network input from the conditional to the
// unconditional program).  One float4 per thread, exact fp32 arithmetic in this order:
// fmaf is NOT used so that w = 0 returns eps_cond bit for bit (1 * x + (-0) * y).
__global__ void __launch_bounds__(256)
axpby_kernel(float* __restrict__ out, float a, const float* __restrict__ x, float b,
             const float* __restrict__ y, int64_t n) {
  pdl_trigger_early();
  pdl_wait();
  const int64_t n4 = n >> 2;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += stride) {
    const float4 xv = reinterpret_cast<const float4*>(x)[i];
    float4 o;
    if (y) {
      const float4 yv = reinterpret_cast<const float4*>(y)[i];
      o.x = __fadd_rn(__fmul_rn(a, xv.x), __fmul_rn(b, yv.x));
      o.y = __fadd_rn(__fmul_rn(a, xv.y), __fmul_rn(b, yv.y));
      o.z = __fadd_rn(__fmul_rn(a, xv.z), __fmul_rn(b, yv.z));
      o.w = __fadd_rn(__fmul_rn(a, xv.w), __fmul_rn(b, yv.w));
    } else {
      o = make_float4(a * xv.x, a * xv.y, a * xv.z, a * xv.w);
    }
    reinterpret_cast<float4*>(out)[i] = o;
  }
  for (int64_t i = (n4 << 2) + (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride)
    out[i] = y ? __fadd_rn(__fmul_rn(a, x[i]), __fmul_rn(b, y[i])) : a * x[i];
}

int launch_axpby(float* out, float a, const float* x, float b, const float* y, int64_t n,
                 cudaStream_t s) {
  PSLD_CHECK_ARG(out && x && n > 0, "axpby: bad arguments");
  PSLD_CHECK_ARG((((uintptr_t)out | (uintptr_t)x | (uintptr_t)y) & 15) == 0,
                 "axpby: buffers must be 16-byte aligned");
  launch_pdl(axpby_kernel, dim3(ew_grid(n, 4)), dim3(256), 0, s, 1, out, a, x, b, y, n);
  PSLD_CHECK_LAUNCH();
  return PSLD_OK;
}

int run_axpby(const psld_op& op, cudaStream_t s) {
  const int64_t n = (int64_t)op.i[0] | ((int64_t)op.i[1] << 31);
  return launch_axpby((float*)op.out[0], op.f[0], (const float*)op.in[0], op.f[1],
                      (const float*)op.in[1], n, s);
}

// ======================================================================== time embedding
// emb[r, j]: fourier: x = ((log t * W[j]) * 2) * pi in fp32, sin | cos  (layerspp.py:40-41; the
// multiplication order is kept: re-associating moves the embedding by 1e-4)
__global__ void temb_embed_kernel(const float* __restrict__ t, const float* __restrict__ W,
                                  float* __restrict__ emb, int nt, int nf, int emb_type,
                                  int logged, const int* __restrict__ step_ptr) {
  pdl_trigger_early();
  pdl_wait();
  if (step_ptr) t += *step_ptr;      // graph replay: t is a per-call table, the step lives on device
  const int E = emb_type == 0 ? 2 * nf : nf;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= nt * E) return;
  const int r = i / E, j = i % E;
  if (emb_type == 0) {
    const float lt = logged ? t[r] : logf(t[r]);
    const int jj = j < nf ? j : j - nf;
    float x = __fmul_rn(__fmul_rn(__fmul_rn(lt, W[jj]), 2.0f), 3.14159274101257324f);
    emb[i] = j < nf ? sinf(x) : cosf(x);
  } else {  // get_timestep_embedding (layers.py:500-514), embedding_dim = nf
    const int half = nf / 2;
    if (j >= 2 * half) { emb[i] = 0.f; return; }
    const int jj = j < half ? j : j - half;
    const float sc = logf(10000.0f) / (float)(half - 1);
    const float f = expf((float)jj * -sc);
    const float x = __fmul_rn(t[r], f);
    emb[i] = j < half ? sinf(x) : cosf(x);
  }
}

// out[r, o] = b[o] + sum_i act(in[r, i]) * W[o, i]; one warp per output element.
template <bool kSiluIn>
__global__ void __launch_bounds__(256)
linear_rows_kernel(const float* __restrict__ in, const float* __restrict__ W,
                   const float* __restrict__ b, float* __restrict__ out, int rows, int K, int O) {
  pdl_trigger_early();
  pdl_wait();
  const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (warp >= (int64_t)rows * O) return;
  const int r = (int)(warp / O), o = (int)(warp % O);
  const float* x = in + (int64_t)r * K;
  const float* w = W + (int64_t)o * K;
  float acc = 0.f;
  for (int i = lane; i < K; i += 32) {
    float v = x[i];
    if (kSiluIn) v = silu_f(v);
    acc = fmaf(v, w[i], acc);
  }
#pragma unroll
  for (int d = 16; d > 0; d >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, d);
  if (lane == 0) out[(int64_t)r * O + o] = acc + (b ? b[o] : 0.f);
}

int run_temb(const psld_op& op, cudaStream_t s) {
  const int nt = op.i[PSLD_TEMB_NT], nf = op.i[PSLD_TEMB_NF], et = op.i[PSLD_TEMB_EMB];
  const int totalC = op.i[PSLD_TEMB_TOTALC], logged = op.i[PSLD_TEMB_LOGGED];
  PSLD_CHECK_ARG(nt > 0 && nf > 0 && totalC > 0, "temb: bad sizes");
  PSLD_CHECK_ARG(op.in[0] && op.in[2] && op.in[4] && op.in[6] && op.out[0] && op.out[1],
                 "temb: null pointer");
  PSLD_CHECK_ARG(et == 1 || op.in[1], "temb: fourier embedding needs W");
  const int E = et == 0 ? 2 * nf : nf, D = 4 * nf;
  float* emb = (float*)op.out[1];
  float* h0 = emb + (int64_t)nt * E;
  float* h1 = h0 + (int64_t)nt * D;
  launch_pdl(temb_embed_kernel, dim3((unsigned)((int)ceil_div((int64_t)nt * E, 128))), dim3(128), 0, s, 1, 
      (const float*)op.in[0], (const float*)op.in[1], emb, nt, nf, et, logged, (const int*)op.out[2]);
  PSLD_CHECK_LAUNCH();
  launch_pdl(linear_rows_kernel<false>, dim3((unsigned)((int)ceil_div((int64_t)nt * D * 32, 256))), dim3(256), 0, s, 1, 
      emb, (const float*)op.in[2], (const float*)op.in[3], h0, nt, E, D);
  PSLD_CHECK_LAUNCH();
  launch_pdl(linear_rows_kernel<true>, dim3((unsigned)((int)ceil_div((int64_t)nt * D * 32, 256))), dim3(256), 0, s, 1, 
      h0, (const float*)op.in[4], (const float*)op.in[5], h1, nt, D, D);
  PSLD_CHECK_LAUNCH();
  launch_pdl(linear_rows_kernel<true>, dim3((unsigned)((int)ceil_div((int64_t)nt * totalC * 32, 256))), dim3(256), 0, s, 1, 
      h1, (const float*)op.in[6], (const float*)op.in[7], (float*)op.out[0], nt, D, totalC);
  PSLD_CHECK_LAUNCH();
  return PSLD_OK;
}

// ======================================================================== GroupNorm
// Pass 1 (stats): per-(sample, pixel-chunk) partial sums per group.  Each thread owns 8 channels
// (one 16-byte bf16 load / two float4 loads per pixel) and walks its pixels 4 at a time with the
// loads issued back to back (memory-level parallelism is what bounds this kernel, not bandwidth:
// the first version ran at 1.4 TB/s with one dependent load in flight per thread).  Sums are
// accumulated in fp32 over short runs and combined in fp64, so E[x^2]-E[x]^2 cancels in double.
// Pass 2 (apply): y = silu?((x - mean) * rstd * gamma + beta), written as ONE concatenated tensor.
template <typename T>
struct Vec8;
template <>
struct Vec8<float> {
  static __device__ __forceinline__ void load(const float* p, float (&v)[8], int = 0) {
    const float4 a = *reinterpret_cast<const float4*>(p);
    const float4 b = *reinterpret_cast<const float4*>(p + 4);
    v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
  }
  static __device__ __forceinline__ void store(float* p, const float (&v)[8], int = 0) {
    *reinterpret_cast<float4*>(p) = make_float4(v[0], v[1], v[2], v[3]);
    *reinterpret_cast<float4*>(p + 4) = make_float4(v[4], v[5], v[6], v[7]);
  }
};
template <>
struct Vec8<__nv_bfloat16> {
  static __device__ __forceinline__ void load(const __nv_bfloat16* p, float (&v)[8], int = 0) {
    const uint4 r = *reinterpret_cast<const uint4*>(p);
    const uint32_t w[4] = {r.x, r.y, r.z, r.w};
#pragma unroll
    for (int t = 0; t < 4; ++t) {
      const float2 f = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&w[t]));
      v[2 * t] = f.x;
      v[2 * t + 1] = f.y;
    }
  }
  static __device__ __forceinline__ void store(__nv_bfloat16* p, const float (&v)[8], int = 0) {
    uint32_t w[4];
#pragma unroll
    for (int t = 0; t < 4; ++t) {
      __nv_bfloat162 h = __floats2bfloat162_rn(v[2 * t], v[2 * t + 1]);
      w[t] = *reinterpret_cast<uint32_t*>(&h);
    }
    *reinterpret_cast<uint4*>(p) = make_uint4(w[0], w[1], w[2], w[3]);
  }
};

template <>
struct Vec8<bf16s> {
  static __device__ __forceinline__ void load(const bf16s* p, float (&v)[8], int lo) {
    const uint4 h = *reinterpret_cast<const uint4*>(p);
    const uint4 l = *reinterpret_cast<const uint4*>(p + lo);
    const uint32_t hw[4] = {h.x, h.y, h.z, h.w}, lw[4] = {l.x, l.y, l.z, l.w};
#pragma unroll
    for (int t = 0; t < 4; ++t) {
      const float2 a = bf2_to_f2(hw[t]), b = bf2_to_f2(lw[t]);
      v[2 * t] = a.x + b.x;
      v[2 * t + 1] = a.y + b.y;
    }
  }
  static __device__ __forceinline__ void store(bf16s* p, const float (&v)[8], int lo) {
    uint32_t hw[4], lw[4];
#pragma unroll
    for (int t = 0; t < 4; ++t) split_bf2(v[2 * t], v[2 * t + 1], hw[t], lw[t]);
    *reinterpret_cast<uint4*>(p) = make_uint4(hw[0], hw[1], hw[2], hw[3]);
    *reinterpret_cast<uint4*>(p + lo) = make_uint4(lw[0], lw[1], lw[2], lw[3]);
  }
};

// accurate SiLU for the fp32 path, fast-intrinsic SiLU when the result is rounded to bf16 anyway
// fast path: silu(x) = x * sigmoid(x) = h + h * tanh(h), h = x/2 : ONE MUFU op (tanh.approx.f32,
// |err| ~ 5e-4 absolute at most, well under the bf16 rounding of the result) instead of ex2 + rcp
// kFast = 2 (split-bf16 tier): x * rcp(1 + ex2(-x log2 e)) with the MUFU ex2 / rcp approximations
// (2 ulp each; the argument scaling adds |x| 2^-24): ~1e-6 relative for |x| < 16, an order of
// magnitude under the tier's 16-bit operand precision, at a third of the instructions of the
// IEEE expf + division form (the apply pass of this tier is instruction-bound, not HBM-bound).
template <int kFast>
__device__ __forceinline__ float silu_t(float x) {
  if (kFast == 1) {
    const float h = 0.5f * x;
    float t;
    asm("tanh.approx.f32 %0, %1;" : "=f"(t) : "f"(h));
    return fmaf(h, t, h);
  }
  if (kFast == 2) return __fdividef(x, 1.0f + __expf(-x));
  return x / (1.0f + expf(-x));
}

constexpr int GN_UNROLL = 4;

template <typename T, int VW>   // VW = channels per thread (8, or 4 when C % 8 != 0)
__global__ void __launch_bounds__(256)
gn_stats_kernel(const T* __restrict__ x1, const T* __restrict__ x2, double* __restrict__ part,
                int HW, int C1, int C2, int G, int nchunk) {
  pdl_trigger_early();
  pdl_wait();
  extern __shared__ double sh[];  // [2*G]
  const int n = blockIdx.y, chunk = blockIdx.x;
  const int C = C1 + C2, vpr = C / VW, cpg = C / G;
  const int rows_per_iter = blockDim.x / vpr;
  const int per = (HW + nchunk - 1) / nchunk;
  const int p0 = chunk * per, p1 = min(HW, p0 + per);
  for (int i = threadIdx.x; i < 2 * G; i += blockDim.x) sh[i] = 0.0;
  __syncthreads();
  const int tv = threadIdx.x % vpr, tr = threadIdx.x / vpr;
  if (tr < rows_per_iter) {
    const int c = tv * VW;
    const bool first = c < C1;
    const T* src = first ? x1 + c : x2 + (c - C1);
    const int lo = first ? C1 : C2;              // channels of the source = split-bf16 lo offset
    const int ld = lo * Elt<T>::kMul;            // pixel row stride
    float s[VW], q[VW];
    double ds[VW], dq[VW];
#pragma unroll
    for (int k = 0; k < VW; ++k) { s[k] = 0.f; q[k] = 0.f; ds[k] = 0.0; dq[k] = 0.0; }
    int cnt = 0;
    int p = p0 + tr;
    for (; p + (GN_UNROLL - 1) * rows_per_iter < p1; p += GN_UNROLL * rows_per_iter) {
      float v[GN_UNROLL][8];
#pragma unroll
      for (int u = 0; u < GN_UNROLL; ++u) {
        const T* ptr = src + ((int64_t)n * HW + p + u * rows_per_iter) * ld;
        if (VW == 8) Vec8<T>::load(ptr, v[u], lo);
        else { const float4 t = Vec4<T>::load(ptr, lo); v[u][0] = t.x; v[u][1] = t.y; v[u][2] = t.z; v[u][3] = t.w; }
      }
#pragma unroll
      for (int u = 0; u < GN_UNROLL; ++u)
#pragma unroll
        for (int k = 0; k < VW; ++k) { s[k] += v[u][k]; q[k] = fmaf(v[u][k], v[u][k], q[k]); }
      cnt += GN_UNROLL;
      if (cnt >= 32) {
#pragma unroll
        for (int k = 0; k < VW; ++k) { ds[k] += s[k]; dq[k] += q[k]; s[k] = 0.f; q[k] = 0.f; }
        cnt = 0;
      }
    }
    for (; p < p1; p += rows_per_iter) {
      float v[8];
      const T* ptr = src + ((int64_t)n * HW + p) * ld;
      if (VW == 8) Vec8<T>::load(ptr, v, lo);
      else { const float4 t = Vec4<T>::load(ptr, lo); v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w; }
#pragma unroll
      for (int k = 0; k < VW; ++k) { s[k] += v[k]; q[k] = fmaf(v[k], v[k], q[k]); }
    }
    // fold the thread's VW channels into their groups before touching shared memory
    int gprev = c / cpg;
    double as = 0.0, aq = 0.0;
#pragma unroll
    for (int k = 0; k < VW; ++k) {
      const int g = (c + k) / cpg;
      if (g != gprev) {
        atomicAdd(&sh[2 * gprev], as);
        atomicAdd(&sh[2 * gprev + 1], aq);
        as = 0.0; aq = 0.0; gprev = g;
      }
      as += ds[k] + (double)s[k];
      aq += dq[k] + (double)q[k];
    }
    atomicAdd(&sh[2 * gprev], as);
    atomicAdd(&sh[2 * gprev + 1], aq);
  }
  __syncthreads();
  double* dst = part + ((int64_t)n * nchunk + chunk) * 2 * G;
  for (int i = threadIdx.x; i < 2 * G; i += blockDim.x) dst[i] = sh[i];
}

// (sum, sumsq) of group g of sample n from the producers' per-sample micro-group accumulators
// (f64 [N, C/4, 2] per source, conv_tc epilogue): cpg/4 consecutive entries, which may straddle
// the two sources of a virtual concat.
__device__ __forceinline__ void gn_group_sums(const double* __restrict__ mg1,
                                              const double* __restrict__ mg2, int n, int C1, int C2,
                                              int cpg, int g, double& su, double& sq) {
  const int nmg1 = C1 >> 2, nmg2 = C2 >> 2, per = cpg >> 2;
  su = 0.0;
  sq = 0.0;
  for (int k = 0; k < per; ++k) {
    const int m = g * per + k;
    const double2 v = m < nmg1
        ? *reinterpret_cast<const double2*>(mg1 + ((int64_t)n * nmg1 + m) * 2)
        : *reinterpret_cast<const double2*>(mg2 + ((int64_t)n * nmg2 + (m - nmg1)) * 2);
    su += v.x;
    sq += v.y;
  }
}

// kMicro: group statistics come straight from the producers' accumulators (mg1, mg2); otherwise
// from the `part` chunks written by gn_stats_kernel.
template <typename TI, typename TO, int VW, int kFast, bool kMicro>
__global__ void __launch_bounds__(256)
gn_apply_kernel(const TI* __restrict__ x1, const TI* __restrict__ x2,
                const double* __restrict__ part, const double* __restrict__ mg1,
                const double* __restrict__ mg2, const float* __restrict__ gamma,
                const float* __restrict__ beta, TO* __restrict__ y, int HW, int C1, int C2, int G,
                int nchunk, int nchunk_apply, float eps, int silu) {
  pdl_trigger_early();
  pdl_wait();
  extern __shared__ float shf[];  // mean[G], rstd[G]
  const int n = blockIdx.y, chunk = blockIdx.x;
  const int C = C1 + C2, vpr = C / VW, cpg = C / G;
  for (int g = threadIdx.x; g < G; g += blockDim.x) {
    double su = 0, sq = 0;
    if (kMicro) {
      gn_group_sums(mg1, mg2, n, C1, C2, cpg, g, su, sq);
    } else {
      for (int k = 0; k < nchunk; ++k) {
        const double* src = part + ((int64_t)n * nchunk + k) * 2 * G;
        su += src[2 * g];
        sq += src[2 * g + 1];
      }
    }
    const double cntd = (double)HW * cpg;
    const double mean = su / cntd;
    double var = sq / cntd - mean * mean;
    if (var < 0) var = 0;
    shf[g] = (float)mean;
    shf[G + g] = (float)(1.0 / sqrt(var + (double)eps));
  }
  __syncthreads();
  const int rows_per_iter = blockDim.x / vpr;
  const int tv = threadIdx.x % vpr, tr = threadIdx.x / vpr;
  if (tr >= rows_per_iter) return;
  const int per = (HW + nchunk_apply - 1) / nchunk_apply;
  const int p0 = chunk * per, p1 = min(HW, p0 + per);
  const int c = tv * VW;
  const bool first = c < C1;
  const TI* src = first ? x1 + c : x2 + (c - C1);
  const int lo = first ? C1 : C2;                // split-bf16 lo offset of the source
  const int ld = lo * Elt<TI>::kMul;
  constexpr int MO = Elt<TO>::kMul;
  float sc[VW], bi[VW];
#pragma unroll
  for (int k = 0; k < VW; ++k) {
    const int g = (c + k) / cpg;
    sc[k] = shf[G + g] * gamma[c + k];
    bi[k] = beta[c + k] - sc[k] * shf[g];
  }
  int p = p0 + tr;
  for (; p + (GN_UNROLL - 1) * rows_per_iter < p1; p += GN_UNROLL * rows_per_iter) {
    float v[GN_UNROLL][8];
#pragma unroll
    for (int u = 0; u < GN_UNROLL; ++u) {
      const TI* ptr = src + ((int64_t)n * HW + p + u * rows_per_iter) * ld;
      if (VW == 8) Vec8<TI>::load(ptr, v[u], lo);
      else { const float4 t = Vec4<TI>::load(ptr, lo); v[u][0] = t.x; v[u][1] = t.y; v[u][2] = t.z; v[u][3] = t.w; }
    }
#pragma unroll
    for (int u = 0; u < GN_UNROLL; ++u) {
#pragma unroll
      for (int k = 0; k < VW; ++k) {
        float o = fmaf(v[u][k], sc[k], bi[k]);
        v[u][k] = silu ? silu_t<kFast>(o) : o;
      }
      TO* dst = y + ((int64_t)n * HW + p + u * rows_per_iter) * (C * MO) + c;
      if (VW == 8) Vec8<TO>::store(dst, v[u], C);
      else Vec4<TO>::store(dst, make_float4(v[u][0], v[u][1], v[u][2], v[u][3]), C);
    }
  }
  for (; p < p1; p += rows_per_iter) {
    float v[8];
    const TI* ptr = src + ((int64_t)n * HW + p) * ld;
    if (VW == 8) Vec8<TI>::load(ptr, v, lo);
    else { const float4 t = Vec4<TI>::load(ptr, lo); v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w; }
#pragma unroll
    for (int k = 0; k < VW; ++k) {
      float o = fmaf(v[k], sc[k], bi[k]);
      v[k] = silu ? silu_t<kFast>(o) : o;
    }
    TO* dst = y + ((int64_t)n * HW + p) * (C * MO) + c;
    if (VW == 8) Vec8<TO>::store(dst, v, C);
    else Vec4<TO>::store(dst, make_float4(v[0], v[1], v[2], v[3]), C);
  }
}

// Affine table for a GroupNorm-on-load convolution straight from the producers' accumulators:
// per (sample, channel) scale = rstd*gamma, shift = beta - mean*scale.  One thread per channel.
__global__ void __launch_bounds__(256)
gn_affine_micro_kernel(const double* __restrict__ mg1, const double* __restrict__ mg2,
                       const float* __restrict__ gamma, const float* __restrict__ beta,
                       float* __restrict__ affine, int HW, int C1, int C2, int G, float eps) {
  pdl_trigger_early();
  pdl_wait();
  const int n = blockIdx.x;
  const int C = C1 + C2, cpg = C / G;
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    double su, sq;
    gn_group_sums(mg1, mg2, n, C1, C2, cpg, c / cpg, su, sq);
    const double cntd = (double)HW * cpg;
    const double mean = su / cntd;
    double var = sq / cntd - mean * mean;
    if (var < 0) var = 0;
    const float rstd = (float)(1.0 / sqrt(var + (double)eps));
    const float sc = rstd * gamma[c];
    affine[((int64_t)n * C + c) * 2] = sc;
    affine[((int64_t)n * C + c) * 2 + 1] = beta[c] - sc * (float)mean;
  }
}

// Affine-only epilogue of the statistics pass: per (sample, channel) scale = rstd*gamma and
// shift = beta - mean*scale, consumed by the GroupNorm-on-load convolution (conv_gn_tc.cu).
__global__ void __launch_bounds__(256)
gn_affine_kernel(const double* __restrict__ part, const float* __restrict__ gamma,
                 const float* __restrict__ beta, float* __restrict__ affine, int HW, int C, int G,
                 int nchunk, float eps) {
  pdl_trigger_early();
  pdl_wait();
  extern __shared__ float shf[];  // mean[G], rstd[G]
  const int n = blockIdx.x;
  const int cpg = C / G;
  for (int g = threadIdx.x; g < G; g += blockDim.x) {
    double su = 0, sq = 0;
    for (int k = 0; k < nchunk; ++k) {
      const double* src = part + ((int64_t)n * nchunk + k) * 2 * G;
      su += src[2 * g];
      sq += src[2 * g + 1];
    }
    const double cntd = (double)HW * cpg;
    const double mean = su / cntd;
    double var = sq / cntd - mean * mean;
    if (var < 0) var = 0;
    shf[g] = (float)mean;
    shf[G + g] = (float)(1.0 / sqrt(var + (double)eps));
  }
  __syncthreads();
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    const int g = c / cpg;
    const float sc = shf[G + g] * gamma[c];
    affine[((int64_t)n * C + c) * 2] = sc;
    affine[((int64_t)n * C + c) * 2 + 1] = beta[c] - sc * shf[g];
  }
}

int run_gn(const psld_op& op, cudaStream_t s) {
  const int N = op.i[PSLD_GN_N], HW = op.i[PSLD_GN_HW], C1 = op.i[PSLD_GN_C1];
  const int C2 = op.i[PSLD_GN_C2], G = op.i[PSLD_GN_G], silu = op.i[PSLD_GN_SILU];
  const int idt = op.i[PSLD_GN_IN_DTYPE], odt = op.i[PSLD_GN_OUT_DTYPE];
  const int nchunk = op.i[PSLD_GN_NCHUNK];
  const int C = C1 + C2;
  PSLD_CHECK_ARG(N > 0 && HW > 0 && C1 > 0 && C2 >= 0 && G > 0 && nchunk > 0, "gn: bad sizes");
  PSLD_CHECK_ARG(C % G == 0 && C1 % 4 == 0 && C2 % 4 == 0 && C / 4 <= 256,
                 "gn: need C %% G == 0, C1,C2 %% 4 == 0, C <= 1024 (C1=%d C2=%d G=%d)", C1, C2, G);
  PSLD_CHECK_ARG(op.in[0] && (C2 == 0 || op.in[1]) && op.in[2] && op.in[3] && op.out[0] &&
                 op.out[1], "gn: null pointer");
  PSLD_CHECK_ARG(idt == odt, "gn: input and output dtypes must match (%d vs %d)", idt, odt);
  const bool v8 = (C1 % 8 == 0) && (C2 % 8 == 0);
  double* part = (double*)op.out[1];
  const size_t sh1 = 2 * G * sizeof(double), sh2 = 2 * G * sizeof(float);
  // producer-side statistics available for every source?  (conv_tc epilogue, PSLD_OP_CONV out[1])
  const bool fused = op.in[4] && (C2 == 0 || op.in[5]);
  int nchunk_eff = nchunk;
  if (fused) {
    PSLD_CHECK_ARG((C / G) % 4 == 0, "gn: fused statistics need (C/G) %% 4 == 0");
    if (op.i[PSLD_GN_AFFINE_ONLY]) {
      launch_pdl(gn_affine_micro_kernel, dim3((unsigned)N), dim3(256), 0, s, 1, (const double*)op.in[4],
                 (const double*)op.in[5], (const float*)op.in[2], (const float*)op.in[3],
                 (float*)op.out[0], HW, C1, C2, G, op.f[0]);
      PSLD_CHECK_LAUNCH();
      return PSLD_OK;
    }
    nchunk_eff = 1;      // gn_apply_kernel<.., kMicro = true> reads the accumulators itself
  } else {
    dim3 grid(nchunk, N);
#define GN_STATS(T, VW)                                                                        \
  launch_pdl(gn_stats_kernel<T, VW>, dim3(grid), dim3(256), sh1, s, 1, (const T*)op.in[0], (const T*)op.in[1], part, HW, \
                                                C1, C2, G, nchunk)
    if (idt == PSLD_BF16) { if (v8) GN_STATS(__nv_bfloat16, 8); else GN_STATS(__nv_bfloat16, 4); }
    else if (idt == PSLD_BF16S) { if (v8) GN_STATS(bf16s, 8); else GN_STATS(bf16s, 4); }
    else { if (v8) GN_STATS(float, 8); else GN_STATS(float, 4); }
#undef GN_STATS
  }
  PSLD_CHECK_LAUNCH();
  const float* ga0 = (const float*)op.in[2];
  const float* be0 = (const float*)op.in[3];
  if (op.i[PSLD_GN_AFFINE_ONLY]) {
    launch_pdl(gn_affine_kernel, dim3((unsigned)(N)), dim3(256), sh2, s, 1, part, ga0, be0, (float*)op.out[0], HW, C, G, nchunk_eff,
                                        op.f[0]);
    PSLD_CHECK_LAUNCH();
    return PSLD_OK;
  }
  // apply pass: pure streaming, ~16 KB of input per CTA
  const int vw = v8 ? 8 : 4;
  int nca = (int)ceil_div((int64_t)HW * (C / vw), 256 * 2 * GN_UNROLL);
  if (nca < 1) nca = 1;
  dim3 grid2(nca, N);
  const float eps = op.f[0];
  const float* ga = (const float*)op.in[2];
  const float* be = (const float*)op.in[3];
  const double* m1 = (const double*)op.in[4];
  const double* m2 = (const double*)op.in[5];
#define GN_APPLY(T, VW, FAST, MICRO)                                                            \
  launch_pdl(gn_apply_kernel<T, T, VW, FAST, MICRO>, grid2, dim3(256), sh2, s, 1, (const T*)op.in[0], \
             (const T*)op.in[1], part, m1, m2, ga, be, (T*)op.out[0], HW, C1, C2, G, nchunk_eff, nca, \
             eps, silu)
#define GN_APPLY2(T, VW, FAST) do { if (fused) GN_APPLY(T, VW, FAST, true); else GN_APPLY(T, VW, FAST, false); } while (0)
  if (idt == PSLD_BF16) { if (v8) GN_APPLY2(__nv_bfloat16, 8, 1); else GN_APPLY2(__nv_bfloat16, 4, 1); }
  else if (idt == PSLD_BF16S) { if (v8) GN_APPLY2(bf16s, 8, 2); else GN_APPLY2(bf16s, 4, 2); }
  else { if (v8) GN_APPLY2(float, 8, 0); else GN_APPLY2(float, 4, 0); }
#undef GN_APPLY2
#undef GN_APPLY
  PSLD_CHECK_LAUNCH();
  return PSLD_OK;
}

// ======================================================================== FIR (upfirdn2d)
// y[n,oy,ox,c] = sum_{ky,kx} kflip[ky][kx] * xu[oy*down + ky, ox*down + kx], where xu is the
// zero-stuffed (x up), (pad0,pad1)-padded / cropped input  (op/upfirdn2d.py:159-200).
struct FirTaps { float k[16]; };

// UP / DOWN are compile-time (the three cases the network uses: 2/1, 1/2, 1/1); 0 = runtime.
template <typename T, int VW, int UP, int DOWN>
__global__ void __launch_bounds__(256)
fir_kernel(const T* __restrict__ x, T* __restrict__ y, FirTaps taps, int N, int H, int W, int C,
           int Cact, int OH, int OW, int up_rt, int down_rt, int pad0, int KH) {
  pdl_trigger_early();
  pdl_wait();
  const int up = UP ? UP : up_rt, down = DOWN ? DOWN : down_rt;
  const int vpr = Cact / VW;      // only the first Cact channels are filtered (rest: left as is)
  const int64_t total = (int64_t)N * OH * OW * vpr;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (int64_t)gridDim.x * blockDim.x) {
    const int cv = (int)(i % vpr);
    int64_t r = i / vpr;
    const int ox = (int)(r % OW); r /= OW;
    const int oy = (int)(r % OH);
    const int n = (int)(r / OH);
    float acc[VW];
#pragma unroll
    for (int k = 0; k < VW; ++k) acc[k] = 0.f;
    // valid taps only: uy = oy*down + ky - pad0 must be a non-negative multiple of `up`
#pragma unroll 4
    for (int ky = 0; ky < KH; ++ky) {
      const int uy = oy * down + ky - pad0;          // coordinate in the zero-stuffed image
      if (uy < 0 || uy % up != 0) continue;
      const int iy = uy / up;
      if (iy >= H) continue;
#pragma unroll 4
      for (int kx = 0; kx < KH; ++kx) {
        const int ux = ox * down + kx - pad0;
        if (ux < 0 || ux % up != 0) continue;
        const int ix = ux / up;
        if (ix >= W) continue;
        const float w = taps.k[(KH - 1 - ky) * KH + (KH - 1 - kx)];   // true convolution: flipped
        const T* src = x + (((int64_t)n * H + iy) * W + ix) * (C * Elt<T>::kMul) + cv * VW;
        float v[8];
        if (VW == 8) Vec8<T>::load(src, v, C);
        else { const float4 t = Vec4<T>::load(src, C); v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w; }
#pragma unroll
        for (int k = 0; k < VW; ++k) acc[k] = fmaf(w, v[k], acc[k]);
      }
    }
    T* dst = y + (((int64_t)n * OH + oy) * OW + ox) * (C * Elt<T>::kMul) + cv * VW;
    if (VW == 8) {
      float o[8];
#pragma unroll
      for (int k = 0; k < 8; ++k) o[k] = acc[k < VW ? k : 0];
      Vec8<T>::store(dst, o, C);
    } else {
      Vec4<T>::store(dst, make_float4(acc[0], acc[1], acc[2], acc[3]), C);
    }
  }
}

// scalar-channel variant (C % 4 != 0, e.g. the 6-channel input pyramid) and NCHW planes
template <typename T>
__global__ void __launch_bounds__(256)
fir_scalar_kernel(const T* __restrict__ x, T* __restrict__ y, FirTaps taps, int N, int H, int W,
                  int C, int OH, int OW, int up_x, int up_y, int down_x, int down_y, int pad_x0,
                  int pad_y0, int KH, int KW, int64_t sn, int64_t sy, int64_t sx, int64_t sc,
                  int64_t on, int64_t oyS, int64_t oxS, int64_t oc, int lo) {
  pdl_trigger_early();
  pdl_wait();
  const int64_t total = (int64_t)N * OH * OW * C;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (int64_t)gridDim.x * blockDim.x) {
    // innermost index follows the contiguous output dimension chosen by the caller
    int64_t r = i;
    int c, ox, oy, n;
    if (oc == 1) { c = (int)(r % C); r /= C; ox = (int)(r % OW); r /= OW; oy = (int)(r % OH); n = (int)(r / OH); }
    else { ox = (int)(r % OW); r /= OW; oy = (int)(r % OH); r /= OH; c = (int)(r % C); n = (int)(r / C); }
    float acc = 0.f;
    for (int ky = 0; ky < KH; ++ky) {
      const int uy = oy * down_y + ky - pad_y0;
      if (uy < 0 || uy % up_y != 0) continue;
      const int iy = uy / up_y;
      if (iy >= H) continue;
      for (int kx = 0; kx < KW; ++kx) {
        const int ux = ox * down_x + kx - pad_x0;
        if (ux < 0 || ux % up_x != 0) continue;
        const int ix = ux / up_x;
        if (ix >= W) continue;
        const float w = taps.k[(KH - 1 - ky) * KW + (KW - 1 - kx)];
        acc = fmaf(w, ld_elt<T>(x + n * sn + iy * sy + ix * sx + c * sc, lo), acc);
      }
    }
    st_elt<T>(y + n * on + oy * oyS + ox * oxS + c * oc, acc, lo);
  }
}

// upsample_2d (up_or_down_sampling.py:195-224: up = 2, 4x4 taps, pad = (2, 1)) in polyphase form:
// one thread = one INPUT pixel (8 channels) -> its 2x2 output quad, from the 3x3 input
// neighbourhood.  Output (2qy+dy, 2qx+dx) sees the zero-stuffed image at uy = 2qy+dy+ky-2, which is
// an input row only for ky = 2*ry + 2 - dy (ry = iy - qy): 2x2 of the 16 taps per output, 9 loads
// per 4 outputs instead of 16.
template <typename T>
__global__ void __launch_bounds__(256)
fir_up2_kernel(const T* __restrict__ x, T* __restrict__ y, FirTaps taps, int N, int H, int W, int C) {
  pdl_trigger_early();
  pdl_wait();
  const int vpr = C / 8;
  const int64_t total = (int64_t)N * H * W * vpr;
  const int OW = 2 * W;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (int64_t)gridDim.x * blockDim.x) {
    const int cv = (int)(i % vpr);
    int64_t r = i / vpr;
    const int qx = (int)(r % W); r /= W;
    const int qy = (int)(r % H);
    const int n = (int)(r / H);
    float acc[4][8];
#pragma unroll
    for (int o = 0; o < 4; ++o)
#pragma unroll
      for (int k = 0; k < 8; ++k) acc[o][k] = 0.f;
#pragma unroll
    for (int ry = -1; ry <= 1; ++ry) {
      const int iy = qy + ry;
      if (iy < 0 || iy >= H) continue;
#pragma unroll
      for (int rx = -1; rx <= 1; ++rx) {
        const int ix = qx + rx;
        if (ix < 0 || ix >= W) continue;
        float v[8];
        Vec8<T>::load(x + (((int64_t)n * H + iy) * W + ix) * (C * Elt<T>::kMul) + cv * 8, v, C);
#pragma unroll
        for (int dy = 0; dy < 2; ++dy) {
          const int ky = 2 * ry + 2 - dy;
          if (ky < 0 || ky > 3) continue;
#pragma unroll
          for (int dx = 0; dx < 2; ++dx) {
            const int kx = 2 * rx + 2 - dx;
            if (kx < 0 || kx > 3) continue;
            const float w = taps.k[(3 - ky) * 4 + (3 - kx)];     // true convolution: flipped
#pragma unroll
            for (int k = 0; k < 8; ++k) acc[dy * 2 + dx][k] = fmaf(w, v[k], acc[dy * 2 + dx][k]);
          }
        }
      }
    }
#pragma unroll
    for (int dy = 0; dy < 2; ++dy)
#pragma unroll
      for (int dx = 0; dx < 2; ++dx)
        Vec8<T>::store(y + (((int64_t)n * 2 * H + 2 * qy + dy) * OW + 2 * qx + dx) * (C * Elt<T>::kMul) + cv * 8,
                       acc[dy * 2 + dx], C);
  }
}

// downsample_2d (up_or_down_sampling.py:227-257: down = 2, 4x4 taps, pad = (1, 1)): one thread = a
// 2x2 output quad (8 channels) from the 6x6 input patch it covers: 36 loads per 4 outputs instead
// of 64.  Output (2qy+dy, 2qx+dx) reads input row 4qy - 1 + r with tap ky = r - 2dy.
template <typename T>
__global__ void __launch_bounds__(256)
fir_down2_kernel(const T* __restrict__ x, T* __restrict__ y, FirTaps taps, int N, int H, int W,
                 int C) {
  pdl_trigger_early();
  pdl_wait();
  const int vpr = C / 8;
  const int OH = H / 2, OW = W / 2, QH = OH / 2, QW = OW / 2;
  const int64_t total = (int64_t)N * QH * QW * vpr;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (int64_t)gridDim.x * blockDim.x) {
    const int cv = (int)(i % vpr);
    int64_t q = i / vpr;
    const int qx = (int)(q % QW); q /= QW;
    const int qy = (int)(q % QH);
    const int n = (int)(q / QH);
    float acc[4][8];
#pragma unroll
    for (int o = 0; o < 4; ++o)
#pragma unroll
      for (int k = 0; k < 8; ++k) acc[o][k] = 0.f;
#pragma unroll
    for (int r = 0; r < 6; ++r) {
      const int iy = 4 * qy - 1 + r;
      if (iy < 0 || iy >= H) continue;
#pragma unroll
      for (int c = 0; c < 6; ++c) {
        const int ix = 4 * qx - 1 + c;
        if (ix < 0 || ix >= W) continue;
        float v[8];
        Vec8<T>::load(x + (((int64_t)n * H + iy) * W + ix) * (C * Elt<T>::kMul) + cv * 8, v, C);
#pragma unroll
        for (int dy = 0; dy < 2; ++dy) {
          const int ky = r - 2 * dy;
          if (ky < 0 || ky > 3) continue;
#pragma unroll
          for (int dx = 0; dx < 2; ++dx) {
            const int kx = c - 2 * dx;
            if (kx < 0 || kx > 3) continue;
            const float w = taps.k[(3 - ky) * 4 + (3 - kx)];     // true convolution: flipped
#pragma unroll
            for (int k = 0; k < 8; ++k) acc[dy * 2 + dx][k] = fmaf(w, v[k], acc[dy * 2 + dx][k]);
          }
        }
      }
    }
#pragma unroll
    for (int dy = 0; dy < 2; ++dy)
#pragma unroll
      for (int dx = 0; dx < 2; ++dx)
        Vec8<T>::store(y + (((int64_t)n * OH + 2 * qy + dy) * OW + 2 * qx + dx) * (C * Elt<T>::kMul) + cv * 8,
                       acc[dy * 2 + dx], C);
  }
}

int run_fir(const psld_op& op, cudaStream_t s) {
  const int N = op.i[PSLD_FIR_N], H = op.i[PSLD_FIR_H], W = op.i[PSLD_FIR_W], C = op.i[PSLD_FIR_C];
  const int up = op.i[PSLD_FIR_UP], down = op.i[PSLD_FIR_DOWN];
  const int pad0 = op.i[PSLD_FIR_PAD0], pad1 = op.i[PSLD_FIR_PAD1], KH = op.i[PSLD_FIR_KH];
  const int dt = op.i[PSLD_FIR_DTYPE];
  PSLD_CHECK_ARG(N > 0 && H > 0 && W > 0 && C > 0 && up >= 1 && down >= 1 && KH >= 1 && KH <= 4,
                 "fir: bad arguments");
  PSLD_CHECK_ARG(op.in[0] && op.out[0], "fir: null pointer");
  const int OH = (H * up + pad0 + pad1 - KH) / down + 1;   // op/upfirdn2d.py:103-104
  const int OW = (W * up + pad0 + pad1 - KH) / down + 1;
  PSLD_CHECK_ARG(OH > 0 && OW > 0, "fir: empty output");
  FirTaps taps;
  for (int i = 0; i < 16; ++i) taps.k[i] = i < KH * KH ? op.f[i] : 0.f;
  int Cact = op.i[PSLD_FIR_CACT] > 0 ? op.i[PSLD_FIR_CACT] : C;
  PSLD_CHECK_ARG(Cact <= C && (Cact == C || (C % 8 == 0 && Cact % 8 == 0)), "fir: bad active channel count");
  if (C % 8 == 0 && Cact == C && up == 2 && down == 1 && KH == 4 && pad0 == 2 && pad1 == 1) {
    const int grid = (int)ceil_div((int64_t)N * H * W * (C / 8), 256);
    if (dt == PSLD_BF16)
      launch_pdl(fir_up2_kernel<__nv_bfloat16>, dim3(grid), dim3(256), 0, s, 1,
                 (const __nv_bfloat16*)op.in[0], (__nv_bfloat16*)op.out[0], taps, N, H, W, C);
    else if (dt == PSLD_BF16S)
      launch_pdl(fir_up2_kernel<bf16s>, dim3(grid), dim3(256), 0, s, 1, (const bf16s*)op.in[0],
                 (bf16s*)op.out[0], taps, N, H, W, C);
    else
      launch_pdl(fir_up2_kernel<float>, dim3(grid), dim3(256), 0, s, 1, (const float*)op.in[0],
                 (float*)op.out[0], taps, N, H, W, C);
  } else if (C % 8 == 0 && Cact == C && up == 1 && down == 2 && KH == 4 && pad0 == 1 && pad1 == 1 &&
             H % 4 == 0 && W % 4 == 0) {
    const int grid = (int)ceil_div((int64_t)N * (H / 4) * (W / 4) * (C / 8), 256);
    if (dt == PSLD_BF16)
      launch_pdl(fir_down2_kernel<__nv_bfloat16>, dim3(grid), dim3(256), 0, s, 1,
                 (const __nv_bfloat16*)op.in[0], (__nv_bfloat16*)op.out[0], taps, N, H, W, C);
    else if (dt == PSLD_BF16S)
      launch_pdl(fir_down2_kernel<bf16s>, dim3(grid), dim3(256), 0, s, 1, (const bf16s*)op.in[0],
                 (bf16s*)op.out[0], taps, N, H, W, C);
    else
      launch_pdl(fir_down2_kernel<float>, dim3(grid), dim3(256), 0, s, 1, (const float*)op.in[0],
                 (float*)op.out[0], taps, N, H, W, C);
  } else if (C % 4 == 0) {
    const int vw = C % 8 == 0 ? 8 : 4;
    const int grid = (int)ceil_div((int64_t)N * OH * OW * (Cact / vw), 256);
#define FIR_LAUNCH2(T, VW, U, D)                                                              \
  launch_pdl(fir_kernel<T, VW, U, D>, dim3(grid), dim3(256), 0, s, 1, (const T*)op.in[0], (T*)op.out[0], taps, N, H, W, C, \
                                               Cact, OH, OW, up, down, pad0, KH)
#define FIR_LAUNCH(T, VW)                                          \
  do {                                                             \
    if (up == 2 && down == 1) FIR_LAUNCH2(T, VW, 2, 1);            \
    else if (up == 1 && down == 2) FIR_LAUNCH2(T, VW, 1, 2);       \
    else if (up == 1 && down == 1) FIR_LAUNCH2(T, VW, 1, 1);       \
    else FIR_LAUNCH2(T, VW, 0, 0);                                 \
  } while (0)
    if (dt == PSLD_BF16) { if (vw == 8) FIR_LAUNCH(__nv_bfloat16, 8); else FIR_LAUNCH(__nv_bfloat16, 4); }
    else if (dt == PSLD_BF16S) { if (vw == 8) FIR_LAUNCH(bf16s, 8); else FIR_LAUNCH(bf16s, 4); }
    else { if (vw == 8) FIR_LAUNCH(float, 8); else FIR_LAUNCH(float, 4); }
#undef FIR_LAUNCH
#undef FIR_LAUNCH2
  } else {
    const int grid = ew_grid((int64_t)N * OH * OW * C);
    const int64_t rs = dt == PSLD_BF16S ? 2 * C : C;      // pixel row stride in elements
    const int64_t sn = (int64_t)H * W * rs, sy = (int64_t)W * rs, sx = rs, sc = 1;
    const int64_t on = (int64_t)OH * OW * rs, oyS = (int64_t)OW * rs, oxS = rs, oc = 1;
    if (dt == PSLD_BF16)
      launch_pdl(fir_scalar_kernel<__nv_bfloat16>, dim3(grid), dim3(256), 0, s, 1, 
          (const __nv_bfloat16*)op.in[0], (__nv_bfloat16*)op.out[0], taps, N, H, W, C, OH, OW, up,
          up, down, down, pad0, pad0, KH, KH, sn, sy, sx, sc, on, oyS, oxS, oc, C);
    else if (dt == PSLD_BF16S)
      launch_pdl(fir_scalar_kernel<bf16s>, dim3(grid), dim3(256), 0, s, 1, (const bf16s*)op.in[0],
                 (bf16s*)op.out[0], taps, N, H, W, C, OH, OW, up, up, down, down, pad0, pad0, KH, KH,
                 sn, sy, sx, sc, on, oyS, oxS, oc, C);
    else
      launch_pdl(fir_scalar_kernel<float>, dim3(grid), dim3(256), 0, s, 1, (const float*)op.in[0], (float*)op.out[0], taps,
                                                  N, H, W, C, OH, OW, up, up, down, down, pad0,
                                                  pad0, KH, KH, sn, sy, sx, sc, on, oyS, oxS, oc, C);
  }
  PSLD_CHECK_LAUNCH();
  return PSLD_OK;
}

}  // namespace psld

// Standalone replacement for the reference's pybind op (op/upfirdn2d.cpp:12-23): planes of
// NCHW fp32, separate x/y factors and pads, taps on the host.
extern "C" int psld_upfirdn2d(const float* input, float* output, const float* taps_host, int kh,
                              int kw, int64_t planes, int in_h, int in_w, int up_x, int up_y,
                              int down_x, int down_y, int pad_x0, int pad_x1, int pad_y0,
                              int pad_y1, psld_stream_t stream) {
  using namespace psld;
  PSLD_CHECK_ARG(input && output && taps_host, "psld_upfirdn2d: null pointer");
  PSLD_CHECK_ARG(kh >= 1 && kw >= 1 && kh * kw <= 16, "psld_upfirdn2d: taps must fit 16 floats");
  PSLD_CHECK_ARG(planes > 0 && in_h > 0 && in_w > 0 && up_x >= 1 && up_y >= 1 && down_x >= 1 &&
                 down_y >= 1, "psld_upfirdn2d: bad sizes");
  const int OH = (in_h * up_y + pad_y0 + pad_y1 - kh) / down_y + 1;
  const int OW = (in_w * up_x + pad_x0 + pad_x1 - kw) / down_x + 1;
  PSLD_CHECK_ARG(OH > 0 && OW > 0, "psld_upfirdn2d: empty output");
  FirTaps taps;
  for (int i = 0; i < 16; ++i) taps.k[i] = i < kh * kw ? taps_host[i] : 0.f;
  const int grid = ew_grid(planes * OH * OW);
  // planes-as-batch, C = 1, contiguous W
  launch_pdl(fir_scalar_kernel<float>, dim3(grid), dim3(256), 0, (cudaStream_t)stream, 1, 
      input, output, taps, (int)planes, in_h, in_w, 1, OH, OW, up_x, up_y, down_x, down_y, pad_x0,
      pad_y0, kh, kw, (int64_t)in_h * in_w, in_w, 1, 0, (int64_t)OH * OW, OW, 1, 0, 0);
  PSLD_CHECK_LAUNCH();
  return PSLD_OK;
}
