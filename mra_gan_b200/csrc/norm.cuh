// InstanceNorm3d (+ReLU/LeakyReLU, +residual, +ReplicationPad3d) bandwidth kernels, channels-last.
//
// Layout: x is [N][V][C] (V = D*H*W); one thread owns VEC consecutive channels of one position, so a
// warp reads/writes full 128 B lines.  Statistics are accumulated in fp32 per thread (<= a few hundred
// terms), combined in fp64 (shared memory, then one atomicAdd(double) per block and channel).
//
// Algorithmic bytes: fwd = read x + write y(+halo) ; bwd = read gy + read x + write dx.
#pragma once
#include "common.cuh"

namespace mra {

template <typename T, int VEC> struct VecIO;
template <typename T> struct VecIO<T, 8> {
  static __device__ __forceinline__ void load(const T* p, float (&v)[8]) { Vec8<T>::load(p, v); }
  static __device__ __forceinline__ void store(T* p, const float (&v)[8]) { Vec8<T>::store(p, v); }
};
template <typename T> struct VecIO<T, 1> {
  static __device__ __forceinline__ void load(const T* p, float (&v)[1]) { v[0] = to_f(*p); }
  static __device__ __forceinline__ void store(T* p, const float (&v)[1]) { *p = from_f<T>(v[0]); }
};

struct NormP {
  int N, C, D, H, W;
  int pad, act; float slope;
  int res_pad;            // -1: none
  float eps, momentum;
  int use_running;        // 0: statistics from `stats`; 1: running statistics (eval); 2: mean / rstd given by the caller
  long long V;            // D*H*W
  int G;                  // channel groups = C / VEC
  int Gs;                 // groups per slice (<= 256)
  int rows_pb;            // rows per block in reduction kernels
};

__device__ __forceinline__ void mean_rstd_from_stats(const double* st, long long V, float eps,
                                                     float& mean, float& rstd) {
  const double m = st[0] / (double)V;
  double var = st[1] / (double)V - m * m;
  if (var < 0.0) var = 0.0;
  mean = (float)m;
  rstd = 1.0f / sqrtf((float)var + eps);
}

// ---- statistics: stats[n][c][2] += {sum x, sum x^2} ----
template <typename T, int VEC>
__global__ void __launch_bounds__(256) inorm_stats_kernel(const T* __restrict__ x, double* __restrict__ stats,
                                                           const NormP P) {
  const int n = blockIdx.y;
  const int t = threadIdx.x;
  const int cgl = t % P.Gs, rl = t / P.Gs;
  const int cg = blockIdx.z * 256 + cgl;
  const bool active = rl < P.rows_pb && cg < P.G;
  float s[VEC], ss[VEC];
#pragma unroll
  for (int j = 0; j < VEC; ++j) { s[j] = 0.f; ss[j] = 0.f; }
  if (active) {
    const T* base = x + (long long)n * P.V * P.C + (long long)cg * VEC;
    for (long long r = (long long)blockIdx.x * P.rows_pb + rl; r < P.V; r += (long long)gridDim.x * P.rows_pb) {
      float v[VEC];
      VecIO<T, VEC>::load(base + r * P.C, v);
#pragma unroll
      for (int j = 0; j < VEC; ++j) { s[j] += v[j]; ss[j] = fmaf(v[j], v[j], ss[j]); }
    }
  }
  __shared__ float sm[2][256 * VEC];
#pragma unroll
  for (int j = 0; j < VEC; ++j) { sm[0][t * VEC + j] = s[j]; sm[1][t * VEC + j] = ss[j]; }
  __syncthreads();
  // thread u < Gs*VEC reduces channel (u) of this slice over the row lanes
  for (int u = t; u < P.Gs * VEC; u += 256) {
    const int g = u / VEC, j = u % VEC;
    if (blockIdx.z * 256 + g >= P.G) continue;
    double a = 0.0, b = 0.0;
    for (int r = 0; r < P.rows_pb; ++r) {
      a += (double)sm[0][(r * P.Gs + g) * VEC + j];
      b += (double)sm[1][(r * P.Gs + g) * VEC + j];
    }
    const int c = (blockIdx.z * 256 + g) * VEC + j;
    atomicAdd(stats + ((long long)n * P.C + c) * 2 + 0, a);
    atomicAdd(stats + ((long long)n * P.C + c) * 2 + 1, b);
  }
}

// mean/rstd per (n,c) from stats (or running stats in eval mode) + running-stat EMA.
__global__ void inorm_finalize_kernel(const double* __restrict__ stats, float* __restrict__ mean,
                                      float* __restrict__ rstd, float* running_mean,
                                      float* running_var, const NormP P) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= P.C) return;
  if (P.use_running == 1) {
    const float m = running_mean[c], r = 1.0f / sqrtf(running_var[c] + P.eps);
    for (int n = 0; n < P.N; ++n) { mean[n * P.C + c] = m; rstd[n * P.C + c] = r; }
    return;
  }
  double msum = 0.0, vsum = 0.0;
  for (int n = 0; n < P.N; ++n) {
    const double* st = stats + ((long long)n * P.C + c) * 2;
    float m, r;
    mean_rstd_from_stats(st, P.V, P.eps, m, r);
    mean[n * P.C + c] = m;
    rstd[n * P.C + c] = r;
    const double mu = st[0] / (double)P.V;
    double var = st[1] / (double)P.V - mu * mu;
    if (var < 0.0) var = 0.0;
    msum += mu;
    vsum += var * ((double)P.V / (double)(P.V - 1));
  }
  if (running_mean) {
    // torch: instance_norm -> batch_norm on the (1, N*C, ...) view, buffers repeated then averaged over N
    running_mean[c] = (1.f - P.momentum) * running_mean[c] + P.momentum * (float)(msum / P.N);
    running_var[c] = (1.f - P.momentum) * running_var[c] + P.momentum * (float)(vsum / P.N);
  }
}

// ---- forward apply: y = pad(act((x-mean)*rstd) + residual) ----
template <typename T, int VEC>
__global__ void __launch_bounds__(256) inorm_fwd_kernel(const T* __restrict__ x, const float* __restrict__ mean,
                                                         const float* __restrict__ rstd,
                                                         const T* __restrict__ res, T* __restrict__ y,
                                                         const NormP P) {
  extern __shared__ float smf[];            // mean[C], rstd[C] of sample n
  const int n = blockIdx.y;
  for (int c = threadIdx.x; c < P.C; c += blockDim.x) {
    smf[c] = mean[n * P.C + c];
    smf[P.C + c] = rstd[n * P.C + c];
  }
  __syncthreads();
  const int p = P.pad;
  const int Dp = P.D + 2 * p, Hp = P.H + 2 * p, Wp = P.W + 2 * p;
  const long long items = (long long)Dp * Hp * Wp * P.G;
  const T* xn = x + (long long)n * P.V * P.C;
  T* yn = y + (long long)n * Dp * Hp * Wp * P.C;
  const int rp = P.res_pad;
  const int Hr = P.H + 2 * rp, Wr = P.W + 2 * rp, Dr = P.D + 2 * rp;
  const T* rn = (rp >= 0) ? res + (long long)n * Dr * Hr * Wr * P.C : nullptr;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < items;
       i += (long long)gridDim.x * blockDim.x) {
    const int cg = (int)(i % P.G);
    long long pos = i / P.G;
    const int pw = (int)(pos % Wp); pos /= Wp;
    const int ph = (int)(pos % Hp);
    const int pd = (int)(pos / Hp);
    const int sw = min(max(pw - p, 0), P.W - 1);
    const int sh = min(max(ph - p, 0), P.H - 1);
    const int sd = min(max(pd - p, 0), P.D - 1);
    const long long src = (((long long)sd * P.H + sh) * P.W + sw) * P.C + (long long)cg * VEC;
    float v[VEC];
    VecIO<T, VEC>::load(xn + src, v);
#pragma unroll
    for (int j = 0; j < VEC; ++j) {
      const int c = cg * VEC + j;
      v[j] = apply_act((v[j] - smf[c]) * smf[P.C + c], P.act, P.slope);
    }
    if (rn) {
      float r[VEC];
      const long long rsrc = ((((long long)sd + rp) * Hr + sh + rp) * Wr + sw + rp) * P.C + (long long)cg * VEC;
      VecIO<T, VEC>::load(rn + rsrc, r);
#pragma unroll
      for (int j = 0; j < VEC; ++j) v[j] += r[j];
    }
    VecIO<T, VEC>::store(yn + (((long long)pd * Hp + ph) * Wp + pw) * P.C + (long long)cg * VEC, v);
  }
}

// fold of a padded gradient onto interior voxel (d,h,w): sum over the padded coords that clamp to it
template <typename T, int VEC>
__device__ __forceinline__ void fold_gather(const T* __restrict__ gn, int d, int h, int w, int cg,
                                            const NormP& P, int Hp, int Wp, float (&g)[VEC]) {
  const int p = P.pad;
  const int d0 = (d == 0) ? 0 : d + p, d1 = (d == P.D - 1) ? d + 2 * p : d + p;
  const int h0 = (h == 0) ? 0 : h + p, h1 = (h == P.H - 1) ? h + 2 * p : h + p;
  const int w0 = (w == 0) ? 0 : w + p, w1 = (w == P.W - 1) ? w + 2 * p : w + p;
#pragma unroll
  for (int j = 0; j < VEC; ++j) g[j] = 0.f;
  for (int a = d0; a <= d1; ++a)
    for (int b = h0; b <= h1; ++b)
      for (int c = w0; c <= w1; ++c) {
        float t[VEC];
        VecIO<T, VEC>::load(gn + (((long long)a * Hp + b) * Wp + c) * P.C + (long long)cg * VEC, t);
#pragma unroll
        for (int j = 0; j < VEC; ++j) g[j] += t[j];
      }
}

// ---- backward statistics: sums[n][c][2] += {sum dy, sum dy*xhat}, dy = act'(xhat) * fold(gy) ----
template <typename T, int VEC>
__global__ void __launch_bounds__(256) inorm_bwd_stats_kernel(const T* __restrict__ gy, const T* __restrict__ x,
                                                               const float* __restrict__ mean,
                                                               const float* __restrict__ rstd,
                                                               double* __restrict__ sums, const NormP P) {
  const int n = blockIdx.y;
  const int t = threadIdx.x;
  const int cgl = t % P.Gs, rl = t / P.Gs;
  const int cg = blockIdx.z * 256 + cgl;
  const bool active = rl < P.rows_pb && cg < P.G;
  const int p = P.pad;
  const int Hp = P.H + 2 * p, Wp = P.W + 2 * p, Dp = P.D + 2 * p;
  float s[VEC], ss[VEC], mu[VEC], rs[VEC];
#pragma unroll
  for (int j = 0; j < VEC; ++j) { s[j] = 0.f; ss[j] = 0.f; mu[j] = 0.f; rs[j] = 0.f; }
  if (active) {
#pragma unroll
    for (int j = 0; j < VEC; ++j) {
      mu[j] = mean[n * P.C + cg * VEC + j];
      rs[j] = rstd[n * P.C + cg * VEC + j];
    }
    const T* xn = x + (long long)n * P.V * P.C + (long long)cg * VEC;
    const T* gn = gy + (long long)n * Dp * Hp * Wp * P.C;
    for (long long r = (long long)blockIdx.x * P.rows_pb + rl; r < P.V; r += (long long)gridDim.x * P.rows_pb) {
      const int w = (int)(r % P.W);
      const int h = (int)((r / P.W) % P.H);
      const int d = (int)(r / ((long long)P.W * P.H));
      float v[VEC], g[VEC];
      VecIO<T, VEC>::load(xn + r * P.C, v);
      fold_gather<T, VEC>(gn, d, h, w, cg, P, Hp, Wp, g);
#pragma unroll
      for (int j = 0; j < VEC; ++j) {
        const float xh = (v[j] - mu[j]) * rs[j];
        const float dy = g[j] * act_grad_from_output(xh, P.act, P.slope);
        s[j] += dy;
        ss[j] = fmaf(dy, xh, ss[j]);
      }
    }
  }
  __shared__ float sm[2][256 * VEC];
#pragma unroll
  for (int j = 0; j < VEC; ++j) { sm[0][t * VEC + j] = s[j]; sm[1][t * VEC + j] = ss[j]; }
  __syncthreads();
  for (int u = t; u < P.Gs * VEC; u += 256) {
    const int g = u / VEC, j = u % VEC;
    if (blockIdx.z * 256 + g >= P.G) continue;
    double a = 0.0, b = 0.0;
    for (int r = 0; r < P.rows_pb; ++r) {
      a += (double)sm[0][(r * P.Gs + g) * VEC + j];
      b += (double)sm[1][(r * P.Gs + g) * VEC + j];
    }
    const int c = (blockIdx.z * 256 + g) * VEC + j;
    atomicAdd(sums + ((long long)n * P.C + c) * 2 + 0, a);
    atomicAdd(sums + ((long long)n * P.C + c) * 2 + 1, b);
  }
}

// ---- backward apply ----
// Iterates the residual-padded position space when dres != nullptr (halo gets zeros), otherwise the
// interior.  dx = rstd * (dy - m1 - xhat * m2); dres(interior) = fold(gy).
template <typename T, int VEC>
__global__ void __launch_bounds__(256) inorm_bwd_kernel(const T* __restrict__ gy, const T* __restrict__ x,
                                                         const float* __restrict__ mean,
                                                         const float* __restrict__ rstd,
                                                         const double* __restrict__ sums, T* __restrict__ dx,
                                                         T* __restrict__ dres, const NormP P) {
  extern __shared__ float smf[];            // mean[C], rstd[C], m1[C], m2[C]
  const int n = blockIdx.y;
  for (int c = threadIdx.x; c < P.C; c += blockDim.x) {
    smf[c] = mean[n * P.C + c];
    smf[P.C + c] = rstd[n * P.C + c];
    if (P.use_running == 1) { smf[2 * P.C + c] = 0.f; smf[3 * P.C + c] = 0.f; }
    else {
      smf[2 * P.C + c] = (float)(sums[((long long)n * P.C + c) * 2 + 0] / (double)P.V);
      smf[3 * P.C + c] = (float)(sums[((long long)n * P.C + c) * 2 + 1] / (double)P.V);
    }
  }
  __syncthreads();
  const int p = P.pad;
  const int Hp = P.H + 2 * p, Wp = P.W + 2 * p, Dp = P.D + 2 * p;
  const int rp = dres ? P.res_pad : 0;
  const int Dr = P.D + 2 * rp, Hr = P.H + 2 * rp, Wr = P.W + 2 * rp;
  const long long items = (long long)Dr * Hr * Wr * P.G;
  const T* xn = x + (long long)n * P.V * P.C;
  const T* gn = gy + (long long)n * Dp * Hp * Wp * P.C;
  T* dxn = dx + (long long)n * P.V * P.C;
  T* drn = dres ? dres + (long long)n * Dr * Hr * Wr * P.C : nullptr;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < items;
       i += (long long)gridDim.x * blockDim.x) {
    const int cg = (int)(i % P.G);
    long long pos = i / P.G;
    const int rw = (int)(pos % Wr); pos /= Wr;
    const int rh = (int)(pos % Hr);
    const int rd = (int)(pos / Hr);
    const int w = rw - rp, h = rh - rp, d = rd - rp;
    const long long roff = (((long long)rd * Hr + rh) * Wr + rw) * P.C + (long long)cg * VEC;
    if (w < 0 || w >= P.W || h < 0 || h >= P.H || d < 0 || d >= P.D) {
      float z[VEC];
#pragma unroll
      for (int j = 0; j < VEC; ++j) z[j] = 0.f;
      VecIO<T, VEC>::store(drn + roff, z);
      continue;
    }
    const long long off = (((long long)d * P.H + h) * P.W + w) * P.C + (long long)cg * VEC;
    float v[VEC], g[VEC], o[VEC];
    VecIO<T, VEC>::load(xn + off, v);
    fold_gather<T, VEC>(gn, d, h, w, cg, P, Hp, Wp, g);
    if (drn) VecIO<T, VEC>::store(drn + roff, g);
#pragma unroll
    for (int j = 0; j < VEC; ++j) {
      const int c = cg * VEC + j;
      const float rs = smf[P.C + c];
      const float xh = (v[j] - smf[c]) * rs;
      const float dy = g[j] * act_grad_from_output(xh, P.act, P.slope);
      o[j] = rs * (dy - smf[2 * P.C + c] - xh * smf[3 * P.C + c]);
    }
    VecIO<T, VEC>::store(dxn + off, o);
  }
}


// the activations that follow a norm layer are piecewise linear: y = x > 0 ? x : x * nslope (none: 1, ReLU: 0)
__device__ __forceinline__ float norm_neg_slope(const NormP& P) {
  return P.act == MRA_ACT_NONE ? 1.f : (P.act == MRA_ACT_RELU ? 0.f : P.slope);
}

// eligibility of the row-streaming kernels (norm_stream.cuh): 16-byte items, power-of-two channel groups
inline int norm_fast_lg(const NormP& P, int vec) {
  if (vec != 8 || P.G > 256 || (P.G & (P.G - 1)) != 0) return -1;
  if (P.act != MRA_ACT_NONE && P.act != MRA_ACT_RELU && P.act != MRA_ACT_LRELU) return -1;
  int lg = 0;
  while ((1 << lg) < P.G) ++lg;
  return lg;
}
// ---- stand-alone replication pad fwd / bwd and activations ----
template <typename T, int VEC>
__global__ void __launch_bounds__(256) reppad_fwd_kernel(const T* __restrict__ x, T* __restrict__ y, const NormP P) {
  const int n = blockIdx.y, p = P.pad;
  const int Dp = P.D + 2 * p, Hp = P.H + 2 * p, Wp = P.W + 2 * p;
  const long long items = (long long)Dp * Hp * Wp * P.G;
  const T* xn = x + (long long)n * P.V * P.C;
  T* yn = y + (long long)n * Dp * Hp * Wp * P.C;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < items;
       i += (long long)gridDim.x * blockDim.x) {
    const int cg = (int)(i % P.G);
    long long pos = i / P.G;
    const int pw = (int)(pos % Wp); pos /= Wp;
    const int ph = (int)(pos % Hp);
    const int pd = (int)(pos / Hp);
    const int sw = min(max(pw - p, 0), P.W - 1), sh = min(max(ph - p, 0), P.H - 1), sd = min(max(pd - p, 0), P.D - 1);
    float v[VEC];
    VecIO<T, VEC>::load(xn + (((long long)sd * P.H + sh) * P.W + sw) * P.C + (long long)cg * VEC, v);
    VecIO<T, VEC>::store(yn + i * VEC, v);
  }
}

template <typename T, int VEC>
__global__ void __launch_bounds__(256) reppad_bwd_kernel(const T* __restrict__ gy, T* __restrict__ dx, const NormP P) {
  const int n = blockIdx.y, p = P.pad;
  const int Dp = P.D + 2 * p, Hp = P.H + 2 * p, Wp = P.W + 2 * p;
  const long long items = P.V * P.G;
  const T* gn = gy + (long long)n * Dp * Hp * Wp * P.C;
  T* dxn = dx + (long long)n * P.V * P.C;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < items;
       i += (long long)gridDim.x * blockDim.x) {
    const int cg = (int)(i % P.G);
    long long pos = i / P.G;
    const int w = (int)(pos % P.W); pos /= P.W;
    const int h = (int)(pos % P.H);
    const int d = (int)(pos / P.H);
    float g[VEC];
    fold_gather<T, VEC>(gn, d, h, w, cg, P, Hp, Wp, g);
    VecIO<T, VEC>::store(dxn + i * VEC, g);
  }
}

template <typename T>
__global__ void __launch_bounds__(256) act_fwd_kernel(const T* __restrict__ x, T* __restrict__ y, long long n,
                                                       int act, float slope) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
    y[i] = from_f<T>(apply_act(to_f(x[i]), act, slope));
}
template <typename T>
__global__ void __launch_bounds__(256) act_bwd_kernel(const T* __restrict__ dy, const T* __restrict__ y,
                                                       T* __restrict__ dx, long long n, int act, float slope) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
    dx[i] = from_f<T>(to_f(dy[i]) * act_grad_from_output(to_f(y[i]), act, slope));
}

// UNet skip connection (networks3D.py:339-343: torch.cat([x, self.model(x)], 1), followed in the parent block by the
// up-path's ReLU(True), :319,326): both halves are written straight into ONE channels-last buffer with the activation
// applied on the way -- out[pos][0:ca] = act(a[pos]), out[pos][ca:ca+cb] = act(b[pos]) -- instead of a concat copy plus
// an activation pass; the backward splits dout the same way with act'(.) taken from the stored output.
// One thread moves VEC consecutive channels (16 bytes for bf16 / VEC = 8); ca, cb are multiples of VEC.
template <typename T, int VEC>
__global__ void __launch_bounds__(256) cat2_act_fwd_kernel(const T* __restrict__ a, const T* __restrict__ b, T* __restrict__ out,
                                                            long long items, int ga, int gb, int act, float slope) {
  const int g = ga + gb;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < items; i += (long long)gridDim.x * blockDim.x) {
    const long long pos = i / g;
    const int cg = (int)(i - pos * g);
    const T* src = cg < ga ? a + (pos * ga + cg) * VEC : b + (pos * gb + (cg - ga)) * VEC;
    T v[VEC];
    if (VEC == 8 && sizeof(T) == 2) *reinterpret_cast<uint4*>(v) = *reinterpret_cast<const uint4*>(src);
    else {
#pragma unroll
      for (int j = 0; j < VEC; ++j) v[j] = src[j];
    }
#pragma unroll
    for (int j = 0; j < VEC; ++j) v[j] = from_f<T>(apply_act(to_f(v[j]), act, slope));
    T* dst = out + i * VEC;
    if (VEC == 8 && sizeof(T) == 2) *reinterpret_cast<uint4*>(dst) = *reinterpret_cast<const uint4*>(v);
    else {
#pragma unroll
      for (int j = 0; j < VEC; ++j) dst[j] = v[j];
    }
  }
}
template <typename T, int VEC>
__global__ void __launch_bounds__(256) cat2_act_bwd_kernel(const T* __restrict__ dout, const T* __restrict__ out, T* __restrict__ da,
                                                            T* __restrict__ db, long long items, int ga, int gb, int act, float slope) {
  const int g = ga + gb;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < items; i += (long long)gridDim.x * blockDim.x) {
    const long long pos = i / g;
    const int cg = (int)(i - pos * g);
    T gv[VEC], yv[VEC];
    if (VEC == 8 && sizeof(T) == 2) {
      *reinterpret_cast<uint4*>(gv) = *reinterpret_cast<const uint4*>(dout + i * VEC);
      *reinterpret_cast<uint4*>(yv) = *reinterpret_cast<const uint4*>(out + i * VEC);
    } else {
#pragma unroll
      for (int j = 0; j < VEC; ++j) { gv[j] = dout[i * VEC + j]; yv[j] = out[i * VEC + j]; }
    }
#pragma unroll
    for (int j = 0; j < VEC; ++j) gv[j] = from_f<T>(to_f(gv[j]) * act_grad_from_output(to_f(yv[j]), act, slope));
    T* dst = cg < ga ? (da ? da + (pos * ga + cg) * VEC : nullptr) : (db ? db + (pos * gb + (cg - ga)) * VEC : nullptr);
    if (!dst) continue;
    if (VEC == 8 && sizeof(T) == 2) *reinterpret_cast<uint4*>(dst) = *reinterpret_cast<const uint4*>(gv);
    else {
#pragma unroll
      for (int j = 0; j < VEC; ++j) dst[j] = gv[j];
    }
  }
}

// nn.Dropout (networks3D.py:244-245, 332-333): y = x * keep / (1 - p).  The keep mask (one byte per element) comes from
// the caller's RNG; the same kernel is its own backward (dx = gy * keep / (1 - p)).
template <typename T>
__global__ void __launch_bounds__(256) mask_scale_kernel(const T* __restrict__ x, const unsigned char* __restrict__ keep,
                                                          T* __restrict__ y, long long n, float scale) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
    y[i] = from_f<T>(keep[i] ? to_f(x[i]) * scale : 0.f);
}

// ---- host-side launch helpers ----
inline NormP make_norm_params(const mra_norm_desc& d, int vec) {
  NormP P;
  P.N = d.n; P.C = d.c; P.D = d.d; P.H = d.h; P.W = d.w;
  P.pad = d.pad; P.act = d.act; P.slope = d.slope; P.res_pad = d.res_pad;
  P.eps = d.eps; P.momentum = d.momentum; P.use_running = d.use_running;
  P.V = (long long)d.d * d.h * d.w;
  P.G = d.c / vec;
  P.Gs = P.G < 256 ? P.G : 256;
  P.rows_pb = 256 / P.Gs;
  return P;
}

inline unsigned grid_for(long long items, int per_block, int waves = 8) {
  long long b = (items + per_block - 1) / per_block;
  const long long cap = (long long)num_sms() * waves;
  if (b > cap) b = cap;
  if (b < 1) b = 1;
  return (unsigned)b;
}

template <typename T, int VEC>
int norm_stats_launch(const mra_norm_desc& d, const void* x, double* stats, cudaStream_t st) {
  NormP P = make_norm_params(d, VEC);
  MRA_CHECK_CUDA(cudaMemsetAsync(stats, 0, sizeof(double) * 2 * d.n * d.c, st));
  // blocks per sample: fill the machine ~4x across the batch
  long long bx = ((long long)num_sms() * 4 + d.n - 1) / d.n;
  const long long maxbx = (P.V + P.rows_pb - 1) / P.rows_pb;
  if (bx > maxbx) bx = maxbx;
  if (bx < 1) bx = 1;
  dim3 grid((unsigned)bx, d.n, (P.G + 255) / 256);
  inorm_stats_kernel<T, VEC><<<grid, 256, 0, st>>>(reinterpret_cast<const T*>(x), stats, P);
  MRA_LAUNCH_CHECK();
  return 0;
}

template <typename T, int VEC>
int norm_fwd_launch(const mra_norm_desc& d, const void* x, const double* stats, const void* res, void* y,
                    float* mean, float* rstd, float* rm, float* rv, cudaStream_t st) {
  NormP P = make_norm_params(d, VEC);
  if (d.use_running != 2) {                           // 2: the caller filled mean / rstd (batch norm with folded affine)
    inorm_finalize_kernel<<<(d.c + 127) / 128, 128, 0, st>>>(stats, mean, rstd, rm, rv, P);
    MRA_LAUNCH_CHECK();
  }
  const long long items = (long long)(d.d + 2 * d.pad) * (d.h + 2 * d.pad) * (d.w + 2 * d.pad) * P.G;
  dim3 grid(grid_for(items, 256 * 4, (16 + d.n - 1) / d.n), d.n);
  inorm_fwd_kernel<T, VEC><<<grid, 256, 2 * d.c * sizeof(float), st>>>(
      reinterpret_cast<const T*>(x), mean, rstd, reinterpret_cast<const T*>(res), reinterpret_cast<T*>(y), P);
  MRA_LAUNCH_CHECK();
  return 0;
}

template <typename T, int VEC>
int norm_bwd_launch(const mra_norm_desc& d, const void* gy, const void* x, const float* mean,
                    const float* rstd, void* dx, void* dres, double* sums, cudaStream_t st, int phases = 3) {
  NormP P = make_norm_params(d, VEC);
  if (d.use_running != 1 && (phases & 1)) {
    MRA_CHECK_CUDA(cudaMemsetAsync(sums, 0, sizeof(double) * 2 * d.n * d.c, st));
    long long bx = ((long long)num_sms() * 4 + d.n - 1) / d.n;
    const long long maxbx = (P.V + P.rows_pb - 1) / P.rows_pb;
    if (bx > maxbx) bx = maxbx;
    if (bx < 1) bx = 1;
    dim3 grid((unsigned)bx, d.n, (P.G + 255) / 256);
    inorm_bwd_stats_kernel<T, VEC><<<grid, 256, 0, st>>>(reinterpret_cast<const T*>(gy),
                                                         reinterpret_cast<const T*>(x), mean, rstd, sums, P);
    MRA_LAUNCH_CHECK();
  }
  if (!(phases & 2)) return 0;
  const int rp = dres ? d.res_pad : 0;
  const long long items = (long long)(d.d + 2 * rp) * (d.h + 2 * rp) * (d.w + 2 * rp) * P.G;
  dim3 grid(grid_for(items, 256 * 4, (16 + d.n - 1) / d.n), d.n);
  inorm_bwd_kernel<T, VEC><<<grid, 256, 4 * d.c * sizeof(float), st>>>(
      reinterpret_cast<const T*>(gy), reinterpret_cast<const T*>(x), mean, rstd, sums,
      reinterpret_cast<T*>(dx), reinterpret_cast<T*>(dres), P);
  MRA_LAUNCH_CHECK();
  return 0;
}

}  // namespace mra
