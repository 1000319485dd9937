// Loss reductions, fused multi-tensor Adam, weight packing, dtype conversion and the sliding-window
// helpers.  All are HBM-bound elementwise / reduction kernels.
#pragma once
#include "common.cuh"

namespace mra {

// ---------------- block reduction to double + one atomic per block ----------------
template <int NV>
__device__ __forceinline__ void block_reduce_atomic(double (&v)[NV], double* acc) {
  __shared__ double sm[NV][8];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int k = 0; k < NV; ++k) {
    double s = warp_sum(v[k]);
    if (lane == 0) sm[k][warp] = s;
  }
  __syncthreads();
  if (threadIdx.x < NV) {
    double s = 0.0;
    const int nw = (blockDim.x + 31) >> 5;
    for (int w = 0; w < nw; ++w) s += sm[threadIdx.x][w];
    atomicAdd(acc + threadIdx.x, s);
  }
}

// ---------------- losses ----------------
// torch.nn.L1Loss / MSELoss / BCELoss (mean reduction is applied by the caller: acc / numel).
// BCE clamps log() at -100 like ATen (binary_cross_entropy).
template <typename T>
__global__ void __launch_bounds__(256) loss_fwd_kernel(int kind, const T* __restrict__ a, const T* __restrict__ b,
                                                        float target, long long n, double* acc) {
  double s[1] = {0.0};
  float part = 0.f;
  int cnt = 0;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const float x = to_f(a[i]);
    float v;
    if (kind == MRA_LOSS_L1) v = fabsf(x - to_f(b[i]));
    else if (kind == MRA_LOSS_MSE_CONST) { const float d = x - target; v = d * d; }
    else {
      const float l1 = fmaxf(logf(x), -100.f), l0 = fmaxf(logf(1.f - x), -100.f);
      v = -(target * l1 + (1.f - target) * l0);
    }
    part += v;
    if (++cnt == 64) { s[0] += (double)part; part = 0.f; cnt = 0; }
  }
  s[0] += (double)part;
  block_reduce_atomic<1>(s, acc);
}

template <typename T>
__global__ void __launch_bounds__(256) loss_bwd_kernel(int kind, const T* __restrict__ a, const T* __restrict__ b,
                                                        float target, long long n, const float* __restrict__ gout,
                                                        float scale, T* __restrict__ da) {
  const float g = gout[0] * scale;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const float x = to_f(a[i]);
    float v;
    if (kind == MRA_LOSS_L1) { const float d = x - to_f(b[i]); v = d > 0.f ? 1.f : (d < 0.f ? -1.f : 0.f); }
    else if (kind == MRA_LOSS_MSE_CONST) v = 2.f * (x - target);
    else v = (x - target) / fmaxf((1.f - x) * x, 1e-12f);
    da[i] = from_f<T>(v * g);
  }
}

template <typename T>
__global__ void __launch_bounds__(256) corr_sums_kernel(const T* __restrict__ x, const T* __restrict__ y,
                                                         long long n, double* acc) {
  double s[5] = {0, 0, 0, 0, 0};
  float p[5] = {0, 0, 0, 0, 0};
  int cnt = 0;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const float a = to_f(x[i]), b = to_f(y[i]);
    p[0] += a; p[1] += b; p[2] = fmaf(a, b, p[2]); p[3] = fmaf(a, a, p[3]); p[4] = fmaf(b, b, p[4]);
    if (++cnt == 64) {
#pragma unroll
      for (int k = 0; k < 5; ++k) { s[k] += (double)p[k]; p[k] = 0.f; }
      cnt = 0;
    }
  }
#pragma unroll
  for (int k = 0; k < 5; ++k) s[k] += (double)p[k];
  block_reduce_atomic<5>(s, acc);
}

// ---------------- Adam ----------------
// torch.optim.Adam (no amsgrad, no weight decay, maximize=False):
//   m = b1 m + (1-b1) g ; v = b2 v + (1-b2) g^2
//   p -= (lr / (1-b1^t)) * m / (sqrt(v)/sqrt(1-b2^t) + eps)
#define MRA_ADAM_MAX_TENSORS 48
struct AdamBatch {
  float* p[MRA_ADAM_MAX_TENSORS];
  const float* g[MRA_ADAM_MAX_TENSORS];
  float* m[MRA_ADAM_MAX_TENSORS];
  float* v[MRA_ADAM_MAX_TENSORS];
  bf16* shadow[MRA_ADAM_MAX_TENSORS];
  long long block_start[MRA_ADAM_MAX_TENSORS + 1];   // prefix sum of blocks per tensor
  long long numel[MRA_ADAM_MAX_TENSORS];
  int count;
  float lr, b1, b2, eps, step_size, bc2_sqrt;
  const float* hyper;      // optional device copy of {lr, b1, b2, eps, step_size, bc2_sqrt}: read instead of the by-value
                           // fields, so that a CUDA graph of the step can be replayed with new step counts / rates
};
#define MRA_ADAM_ELEMS_PER_BLOCK 4096

// one element of torch.optim.Adam: exp_avg.lerp_(grad, 1-b1) with ATen's two-branch lerp; exp_avg_sq.mul_(b2).addcmul_(g, g, 1-b2)
__device__ __forceinline__ void adam_elem(float gi, float& mi, float& vi, float& pi, float b1, float b2, float eps, float step_size,
                                          float bc2_sqrt) {
  const float wl = 1.f - b1, diff = gi - mi;
  mi = (wl < 0.5f) ? mi + wl * diff : gi - diff * (1.f - wl);
  vi = b2 * vi + (1.f - b2) * gi * gi;
  const float denom = sqrtf(vi) / bc2_sqrt + eps;
  pi = pi - step_size * (mi / denom);
}
__global__ void __launch_bounds__(256) adam_kernel(const AdamBatch B) {
  // locate this block's tensor (count <= 48: linear scan)
  int ti = 0;
  while (ti + 1 < B.count && (long long)blockIdx.x >= B.block_start[ti + 1]) ++ti;
  const long long base = ((long long)blockIdx.x - B.block_start[ti]) * MRA_ADAM_ELEMS_PER_BLOCK;
  float* __restrict__ p = B.p[ti];
  const float* __restrict__ g = B.g[ti];
  float* __restrict__ m = B.m[ti];
  float* __restrict__ v = B.v[ti];
  bf16* __restrict__ sh = B.shadow[ti];
  const long long n = B.numel[ti];
  float b1 = B.b1, b2 = B.b2, eps = B.eps, step_size = B.step_size, bc2_sqrt = B.bc2_sqrt;
  if (B.hyper) { b1 = B.hyper[1]; b2 = B.hyper[2]; eps = B.hyper[3]; step_size = B.hyper[4]; bc2_sqrt = B.hyper[5]; }
  // 16-byte accesses (4 elements per thread and iteration) when all arrays of the tensor are 16-byte aligned -- they are
  // for the packed parameters, their Adam moments and the gradient-bucket views (256-byte slots); the 4-byte path measured
  // 4.0 TB/s on the UNet's 334 M parameters (28 + 2 bytes each)
  const bool vec = (((uintptr_t)p | (uintptr_t)g | (uintptr_t)m | (uintptr_t)v) & 15) == 0 && (sh == nullptr || ((uintptr_t)sh & 7) == 0);
  if (vec) {
#pragma unroll
    for (int k = 0; k < MRA_ADAM_ELEMS_PER_BLOCK / 1024; ++k) {
      const long long i = base + ((long long)k * 256 + threadIdx.x) * 4;
      if (i + 3 < n) {
        const float4 g4 = *reinterpret_cast<const float4*>(g + i);
        float4 m4 = *reinterpret_cast<const float4*>(m + i), v4 = *reinterpret_cast<const float4*>(v + i),
               p4 = *reinterpret_cast<const float4*>(p + i);
        adam_elem(g4.x, m4.x, v4.x, p4.x, b1, b2, eps, step_size, bc2_sqrt);
        adam_elem(g4.y, m4.y, v4.y, p4.y, b1, b2, eps, step_size, bc2_sqrt);
        adam_elem(g4.z, m4.z, v4.z, p4.z, b1, b2, eps, step_size, bc2_sqrt);
        adam_elem(g4.w, m4.w, v4.w, p4.w, b1, b2, eps, step_size, bc2_sqrt);
        *reinterpret_cast<float4*>(m + i) = m4; *reinterpret_cast<float4*>(v + i) = v4; *reinterpret_cast<float4*>(p + i) = p4;
        if (sh) {
          __nv_bfloat162 lo = __floats2bfloat162_rn(p4.x, p4.y), hi = __floats2bfloat162_rn(p4.z, p4.w);
          uint2 o;
          o.x = *reinterpret_cast<uint32_t*>(&lo); o.y = *reinterpret_cast<uint32_t*>(&hi);
          *reinterpret_cast<uint2*>(sh + i) = o;
        }
      } else {
        for (long long j = i; j < n && j < i + 4; ++j) {
          float mi = m[j], vi = v[j], pi = p[j];
          adam_elem(g[j], mi, vi, pi, b1, b2, eps, step_size, bc2_sqrt);
          m[j] = mi; v[j] = vi; p[j] = pi;
          if (sh) sh[j] = __float2bfloat16_rn(pi);
        }
      }
    }
    return;
  }
#pragma unroll 4
  for (int k = 0; k < MRA_ADAM_ELEMS_PER_BLOCK / 256; ++k) {
    const long long i = base + k * 256 + threadIdx.x;
    if (i >= n) break;
    float mi = m[i], vi = v[i], pi = p[i];
    adam_elem(g[i], mi, vi, pi, b1, b2, eps, step_size, bc2_sqrt);
    m[i] = mi; v[i] = vi; p[i] = pi;
    if (sh) sh[i] = __float2bfloat16_rn(pi);
  }
}

// Device-resident step counter (CUDA-graph replay without a host <-> device race): state = {lr, beta1, beta2, eps, t}
// as doubles; one thread bumps t and derives the two bias corrections in double, exactly as torch.optim.Adam does on
// the host (bias_correction1 = 1 - beta1^t, step_size = lr / bias_correction1, bias_correction2_sqrt = sqrt(1 - beta2^t)),
// into the float block adam_kernel reads.  Part of the captured graph, so every replay advances by exactly one step
// no matter how far ahead the host runs.
__global__ void adam_advance_kernel(double* __restrict__ state, float* __restrict__ hyper) {
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  const double lr = state[0], b1 = state[1], b2 = state[2], eps = state[3];
  const double t = state[4] + 1.0;
  state[4] = t;
  hyper[0] = (float)lr; hyper[1] = (float)b1; hyper[2] = (float)b2; hyper[3] = (float)eps;
  hyper[4] = (float)(lr / (1.0 - pow(b1, t)));
  hyper[5] = (float)sqrt(1.0 - pow(b2, t));
}

// ---------------- conversions / packing ----------------
template <typename S, typename D>
__global__ void __launch_bounds__(256) convert_kernel(const S* __restrict__ s, D* __restrict__ d, long long n) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
    d[i] = from_f<D>(to_f(s[i]));
}

// wT[t][ci][co] = w[t][co][ci] via a 32x32 shared tile
template <typename S, typename D>
__global__ void pack_weight_t_kernel(const S* __restrict__ w, D* __restrict__ wT, int cout, int cin) {
  __shared__ float tile[32][33];
  const long long slab = (long long)blockIdx.z * cout * cin;
  const int ci0 = blockIdx.x * 32, co0 = blockIdx.y * 32;
  for (int r = threadIdx.y; r < 32; r += blockDim.y) {
    const int co = co0 + r, ci = ci0 + threadIdx.x;
    tile[r][threadIdx.x] = (co < cout && ci < cin) ? to_f(w[slab + (long long)co * cin + ci]) : 0.f;
  }
  __syncthreads();
  for (int r = threadIdx.y; r < 32; r += blockDim.y) {
    const int ci = ci0 + r, co = co0 + threadIdx.x;
    if (ci < cin && co < cout) wT[slab + (long long)ci * cout + co] = from_f<D>(tile[threadIdx.x][r]);
  }
}

// ---------------- sliding-window helpers (test.py:147-178) ----------------
template <typename T>
__global__ void __launch_bounds__(256) window_extract_kernel(const float* __restrict__ vol, int Y, int Z, int i0, int j0,
                                                              int k0, int px, int py, int pz, T* __restrict__ patch) {
  const long long n = (long long)px * py * pz;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(i % pz);
    const int b = (int)((i / pz) % py);
    const int a = (int)(i / ((long long)pz * py));
    const float v = vol[((long long)(i0 + a) * Y + (j0 + b)) * Z + (k0 + c)];
    patch[i] = from_f<T>((v - 127.5f) / 127.5f);
  }
}
template <typename T>
__global__ void __launch_bounds__(256) window_accumulate_kernel(const T* __restrict__ pred, float* __restrict__ label,
                                                                 float* __restrict__ weight, int Y, int Z, int i0, int j0,
                                                                 int k0, int px, int py, int pz) {
  const long long n = (long long)px * py * pz;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(i % pz);
    const int b = (int)((i / pz) % py);
    const int a = (int)(i / ((long long)pz * py));
    const long long o = ((long long)(i0 + a) * Y + (j0 + b)) * Z + (k0 + c);
    label[o] += to_f(pred[i]) * 127.5f + 127.5f;
    weight[o] += 1.0f;
  }
}
__global__ void __launch_bounds__(256) window_finalize_kernel(float* __restrict__ label, const float* __restrict__ weight,
                                                               long long n) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
    label[i] = label[i] / weight[i] + 0.01f;
}

inline unsigned ew_grid(long long n, int per_thread = 4) {
  long long b = (n + 256LL * per_thread - 1) / (256LL * per_thread);
  const long long cap = (long long)num_sms() * 16;
  if (b > cap) b = cap;
  if (b < 1) b = 1;
  return (unsigned)b;
}

}  // namespace mra
