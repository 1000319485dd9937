// Row-streaming InstanceNorm kernels: the bandwidth kernels of norm.cuh rebuilt around a bulk-copy pipeline.
//
// One persistent CTA per SM.  A producer thread streams whole rows (W positions x C channels, contiguous in the
// channels-last layout) global -> shared with `cp.async.bulk` (the 1-D TMA path, completion counted on an
// mbarrier) into a ring of stages; 16 consumer warps read their 16-byte items from shared memory, do the
// arithmetic in fp32 and write results straight to global memory with 16-byte stores.  The ring keeps up to
// ~200 KB per SM in flight, so the kernels stay on the HBM / L2 roofline even when a CTA only sees a dozen rows
// (the 256-channel 32^3 tensors of the residual blocks), where the register-staged kernels of norm.cuh were
// latency bound (one dependent load round trip per row).
//
// Replication padding is a property of the row list, not of the inner loop:
//   forward : source row (d,h) is written to every padded row (a,b) that clamps onto it (1, p+1 or (p+1)^2 rows);
//   backward: interior row (d,h) receives the sum of the padded gradient rows that clamp onto it; the producer
//             streams those rows one stage each, the consumers accumulate them in registers (the w-direction
//             halo of a row is folded from shared memory).
// Algorithmic bytes: fwd = read x + write y(+halo) [+ read residual];
//                                bwd = read gy + read x (stats pass), read gy + read x + write dx [+ dres] (apply).
#pragma once
#include "conv_tc.cuh"
#include "norm.cuh"

namespace mra {
namespace ns {

using tc::mbar_arrive;
using tc::mbar_expect_tx;
using tc::mbar_init;
using tc::mbar_wait;
using tc::smem_u32;

constexpr int kMaxStages = 8;
constexpr uint32_t kSmemBudget = 200 * 1024;
constexpr uint32_t kHdrBytes = 128;                 // full[8] + empty[8] mbarriers

struct StreamP {
  NormP P;
  int lgG;
  int stages;
  uint32_t x_bytes;        // one interior row of x:   W * C * sizeof(T)
  uint32_t g_bytes;        // fwd: one interior row of the residual (0: none); bwd: one padded row of gy
  uint32_t stage_bytes;
  int* err;
};

__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(dst), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}

// The arithmetic runs on packed fp32 pairs (FADD2 / FMUL2 / FFMA2 of sm_100): these kernels sit close to the
// issue limit of the SM (one 16-byte item of 8 channels costs > 100 scalar instructions), the packed forms halve
// the fma-pipe share.  Rounding is identical to the scalar forms.
typedef float2 F8[4];
__device__ __forceinline__ float2 bc2(float a) { return make_float2(a, a); }

// 8 consecutive elements read from shared memory, kept raw until they are used
template <typename T> struct S8;
template <> struct S8<bf16> {
  uint32_t r[4];
  __device__ __forceinline__ void load(uint32_t a) {
    asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(a));
  }
  __device__ __forceinline__ void unpack(F8& v) const {
#pragma unroll
    for (int i = 0; i < 4; ++i) v[i] = make_float2(__uint_as_float(r[i] << 16), __uint_as_float(r[i] & 0xffff0000u));
  }
};
template <> struct S8<float> {
  float r[8];
  __device__ __forceinline__ void load(uint32_t a) {
    asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(r[0]), "=f"(r[1]), "=f"(r[2]), "=f"(r[3]) : "r"(a));
    asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(r[4]), "=f"(r[5]), "=f"(r[6]), "=f"(r[7]) : "r"(a + 16));
  }
  __device__ __forceinline__ void unpack(F8& v) const {
#pragma unroll
    for (int i = 0; i < 4; ++i) v[i] = make_float2(r[2 * i], r[2 * i + 1]);
  }
};
template <typename T> __device__ __forceinline__ void add8(const S8<T>& a, F8& g) {
  F8 t;
  a.unpack(t);
#pragma unroll
  for (int i = 0; i < 4; ++i) g[i] = __fadd2_rn(g[i], t[i]);
}
template <typename T> __device__ __forceinline__ void store8(T* p, const F8& v);
template <> __device__ __forceinline__ void store8<bf16>(bf16* p, const F8& v) {
  uint4 r;
  r.x = pack_bf16x2(v[0].x, v[0].y); r.y = pack_bf16x2(v[1].x, v[1].y);
  r.z = pack_bf16x2(v[2].x, v[2].y); r.w = pack_bf16x2(v[3].x, v[3].y);
  *reinterpret_cast<uint4*>(p) = r;
}
template <> __device__ __forceinline__ void store8<float>(float* p, const F8& v) {
  reinterpret_cast<float4*>(p)[0] = make_float4(v[0].x, v[0].y, v[1].x, v[1].y);
  reinterpret_cast<float4*>(p)[1] = make_float4(v[2].x, v[2].y, v[3].x, v[3].y);
}
// per-thread channel constants as packed pairs
__device__ __forceinline__ void load_pairs(const float* p, F8& v, float scale) {
  const float4 a = *reinterpret_cast<const float4*>(p), b = *reinterpret_cast<const float4*>(p + 4);
  v[0] = make_float2(a.x * scale, a.y * scale); v[1] = make_float2(a.z * scale, a.w * scale);
  v[2] = make_float2(b.x * scale, b.y * scale); v[3] = make_float2(b.z * scale, b.w * scale);
}
// dy = act'(xh) * g for the piecewise-linear activations (slope 1 for xh > 0, nslope otherwise)
__device__ __forceinline__ float2 act_grad2(float2 xh, float2 g, float nslope) {
  return __fmul2_rn(g, make_float2(xh.x > 0.f ? 1.f : nslope, xh.y > 0.f ? 1.f : nslope));
}

// the rows of a padded tensor that clamp onto interior index i (extent n, halo p): [lo, hi]
__device__ __forceinline__ void clamp_range(int i, int n, int p, int& lo, int& hi) {
  lo = (i == 0) ? 0 : i + p;
  hi = (i == n - 1) ? i + 2 * p : i + p;
}

struct Ring {
  uint64_t* full;
  uint64_t* empty;
  uint32_t base;           // shared address of stage 0
  uint32_t stage_bytes;
  int stages;
  int s;                   // current stage
  uint32_t ph;             // current phase
  __device__ __forceinline__ void advance() { if (++s == stages) { s = 0; ph ^= 1; } }
  __device__ __forceinline__ uint32_t addr() const { return base + (uint32_t)s * stage_bytes; }
};

template <int CONS>
__device__ __forceinline__ Ring ring_setup(unsigned char* smem, const StreamP& S) {
  Ring R;
  R.full = reinterpret_cast<uint64_t*>(smem);
  R.empty = R.full + kMaxStages;
  R.base = smem_u32(smem) + kHdrBytes;
  R.stage_bytes = S.stage_bytes;
  R.stages = S.stages;
  R.s = 0; R.ph = 0;
  if (threadIdx.x == 0) {
    for (int i = 0; i < S.stages; ++i) { mbar_init(R.full + i, 1); mbar_init(R.empty + i, CONS / 32); }
    tc::fence_barrier_init();
  }
  __syncthreads();
  return R;
}
// one lane polls the barrier, the warp follows
__device__ __forceinline__ bool ring_wait_full(Ring& R, int lane, int* err, int code) {
  bool ok = true;
  if (lane == 0) ok = mbar_wait(R.full + R.s, R.ph, err, code);
  return __shfl_sync(0xffffffffu, ok ? 1 : 0, 0) != 0;
}
__device__ __forceinline__ void ring_release(Ring& R, int lane) {
  __syncwarp();
  if (lane == 0) mbar_arrive(R.empty + R.s);
  R.advance();
}

// rows row0, row0 + step, ... of a [D][H] row space, walked without a division per row
struct RowWalk {
  int row, d, h, step, sd, sh, H;
  __device__ __forceinline__ RowWalk(int row0, int step_, int H_) : row(row0), step(step_), H(H_) {
    d = row0 / H; h = row0 - d * H; sd = step / H; sh = step - sd * H;
  }
  __device__ __forceinline__ void next() {
    row += step; d += sd; h += sh;
    if (h >= H) { h -= H; ++d; }
  }
};

// ------------------------------------------------------------------------------------------------ forward
template <typename T, int CONS>
__global__ void __launch_bounds__(CONS + 32, 1) inorm_fwd_stream_kernel(const T* __restrict__ x, const double* __restrict__ stats,
                                                                         float* __restrict__ mean, float* __restrict__ rstd,
                                                                         float* running_mean, float* running_var,
                                                                         const T* __restrict__ res, T* __restrict__ y,
                                                                         const StreamP S) {
  pdl_launch_dependents();
  extern __shared__ __align__(128) unsigned char smem[];
  const NormP& P = S.P;
  Ring R = ring_setup<CONS>(smem, S);
  pdl_wait();                                  // barriers are set up: from here on global memory is touched
  const int n = blockIdx.y, t = threadIdx.x;
  const int p = P.pad, rp = P.res_pad;
  const int Hp = P.H + 2 * p, Wp = P.W + 2 * p, Dp = P.D + 2 * p;
  const int rows = P.D * P.H;
  if (t >= CONS) {
    if (t == CONS) {
      const int Hr = P.H + 2 * rp, Wr = P.W + 2 * rp, Dr = P.D + 2 * rp;
      const T* xn = x + (long long)n * P.V * P.C;
      const T* rn = (rp >= 0) ? res + (long long)n * Dr * Hr * Wr * P.C : nullptr;
      for (RowWalk w(blockIdx.x, gridDim.x, P.H); w.row < rows; w.next()) {
        if (!mbar_wait(R.empty + R.s, R.ph ^ 1, S.err, 31)) break;
        mbar_expect_tx(R.full + R.s, S.x_bytes + S.g_bytes);
        bulk_g2s(R.addr(), xn + (long long)w.row * P.W * P.C, S.x_bytes, R.full + R.s);
        if (rn) bulk_g2s(R.addr() + S.x_bytes, rn + ((long long)((w.d + rp) * Hr + w.h + rp) * Wr + rp) * P.C, S.g_bytes, R.full + R.s);
        R.advance();
      }
    }
    return;
  }
  const int G = P.G, cg = t & (G - 1), wl = t >> S.lgG, wpp = CONS >> S.lgG, lane = t & 31;
  // mean / rstd of this thread's 8 channels straight from the conv epilogue's {sum, sum of squares} (the former
  // finalize kernel, same arithmetic); the first CTA of every sample also publishes them for the backward pass and
  // CTA (0, 0) applies the running-statistics EMA (torch: buffers averaged over the batch, unbiased variance)
  F8 nmu, rs;
  {
    float m8[8], r8[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int c = cg * 8 + j;
      if (P.use_running == 1) { m8[j] = running_mean[c]; r8[j] = 1.0f / sqrtf(running_var[c] + P.eps); }
      else if (P.use_running == 2) { m8[j] = mean[n * P.C + c]; r8[j] = rstd[n * P.C + c]; }     // given by the caller
      else mean_rstd_from_stats(stats + ((long long)n * P.C + c) * 2, P.V, P.eps, m8[j], r8[j]);
    }
    if (blockIdx.x == 0 && wl == 0 && P.use_running != 2) {
#pragma unroll
      for (int j = 0; j < 8; ++j) { mean[n * P.C + cg * 8 + j] = m8[j]; rstd[n * P.C + cg * 8 + j] = r8[j]; }
      if (n == 0 && P.use_running == 0 && running_mean) {
        for (int j = 0; j < 8; ++j) {
          const int c = cg * 8 + j;
          double msum = 0.0, vsum = 0.0;
          for (int nn = 0; nn < P.N; ++nn) {
            const double* st = stats + ((long long)nn * P.C + c) * 2;
            const double mu = st[0] / (double)P.V;
            double var = st[1] / (double)P.V - mu * mu;
            if (var < 0.0) var = 0.0;
            msum += mu;
            vsum += var * ((double)P.V / (double)(P.V - 1));
          }
          running_mean[c] = (1.f - P.momentum) * running_mean[c] + P.momentum * (float)(msum / P.N);
          running_var[c] = (1.f - P.momentum) * running_var[c] + P.momentum * (float)(vsum / P.N);
        }
      }
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) { nmu[i] = make_float2(-m8[2 * i], -m8[2 * i + 1]); rs[i] = make_float2(r8[2 * i], r8[2 * i + 1]); }
  }
  const float2 nsl = bc2(norm_neg_slope(P));
  T* yn = y + (long long)n * Dp * Hp * Wp * P.C + cg * 8;
  const uint32_t esz = sizeof(T), cb = (uint32_t)P.C * esz, cgoff = (uint32_t)cg * 8 * esz;
  const long long ystr_b = (long long)Wp * P.C, ystr_a = ystr_b * Hp;       // padded row / plane strides (elements)
  const bool has_res = rp >= 0;
  for (RowWalk w(blockIdx.x, gridDim.x, P.H); w.row < rows; w.next()) {
    int a0, a1, b0, b1;
    clamp_range(w.d, P.D, p, a0, a1);
    clamp_range(w.h, P.H, p, b0, b1);
    T* q0 = yn + a0 * ystr_a + b0 * ystr_b;
    const int na = a1 - a0, nb = b1 - b0;                                    // extra copies along d / h
    if (!ring_wait_full(R, lane, S.err, 32)) return;
    const uint32_t xs = R.addr() + cgoff;
    for (int pw = wl; pw < Wp; pw += wpp) {
      const int sw = min(max(pw - p, 0), P.W - 1);
      S8<T> xv, rv;
      xv.load(xs + (uint32_t)sw * cb);
      if (has_res) rv.load(xs + S.x_bytes + (uint32_t)sw * cb);
      F8 v;
      xv.unpack(v);
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const float2 xh = __fmul2_rn(__fadd2_rn(v[i], nmu[i]), rs[i]);
        const float2 lo = __fmul2_rn(xh, nsl);                       // 0 <= nslope <= 1: act(xh) = max(xh, nslope * xh)
        v[i] = make_float2(fmaxf(xh.x, lo.x), fmaxf(xh.y, lo.y));
      }
      if (has_res) add8<T>(rv, v);
      T* q = q0 + pw * P.C;
      store8<T>(q, v);
      if (na | nb) {                                                          // replication halo: border rows only
        T* qa = q;
        for (int a = 0; a <= na; ++a, qa += ystr_a) {
          T* qb = qa;
          for (int b = 0; b <= nb; ++b, qb += ystr_b)
            if (a | b) store8<T>(qb, v);
        }
      }
    }
    ring_release(R, lane);
  }
}

// ------------------------------------------------------------------------------------------------ backward
// Consumer side shared by the statistics and the apply kernels: fold the padded gradient rows of interior row
// (d,h) into g[u] (one entry per position lane of this thread) and fetch the matching x items.
// off[u] = byte offset of this thread's item u inside an interior row (position w = wl + u*wpp, channel group cg).
template <typename T, int U>
__device__ __forceinline__ bool bwd_gather_row(Ring& R, const StreamP& S, int d, int h, const uint32_t (&off)[U], int wl, int wpp,
                                               int lane, uint32_t cgoff, F8 (&g)[U], S8<T> (&xv)[U]) {
  const NormP& P = S.P;
  const int p = P.pad;
  const uint32_t cb = (uint32_t)P.C * (uint32_t)sizeof(T);
  int a0, a1, b0, b1;
  clamp_range(d, P.D, p, a0, a1);
  clamp_range(h, P.H, p, b0, b1);
  const int nsub = (a1 - a0 + 1) * (b1 - b0 + 1);
  for (int k = 0; k < nsub; ++k) {
    if (!ring_wait_full(R, lane, S.err, 33)) return false;
    const uint32_t gc = R.addr() + (uint32_t)p * cb, xs = R.addr() + S.g_bytes;      // centre of the padded row; x row
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int w = wl + u * wpp;
      if (w < P.W) {
        S8<T> c;
        c.load(gc + off[u]);
        if (k == 0) { xv[u].load(xs + off[u]); c.unpack(g[u]); }
        else add8<T>(c, g[u]);
        if (p > 0) {
          if (w == 0)
            for (int j = 0; j < p; ++j) { S8<T> e; e.load(R.addr() + cgoff + (uint32_t)j * cb); add8<T>(e, g[u]); }
          if (w == P.W - 1)
            for (int j = 1; j <= p; ++j) { S8<T> e; e.load(gc + off[u] + (uint32_t)j * cb); add8<T>(e, g[u]); }
        }
      }
    }
    ring_release(R, lane);
  }
  return true;
}

// producer of both backward kernels: for every interior row, the padded gradient rows that fold onto it (the
// first one travels with the x row)
template <typename T>
__device__ __forceinline__ void bwd_producer(Ring& R, const StreamP& S, const T* __restrict__ gy, const T* __restrict__ x, int n) {
  const NormP& P = S.P;
  const int p = P.pad;
  const int Hp = P.H + 2 * p, Wp = P.W + 2 * p, Dp = P.D + 2 * p;
  const T* xn = x + (long long)n * P.V * P.C;
  const T* gn = gy + (long long)n * Dp * Hp * Wp * P.C;
  const int rows = P.D * P.H;
  for (RowWalk w(blockIdx.x, gridDim.x, P.H); w.row < rows; w.next()) {
    int a0, a1, b0, b1;
    clamp_range(w.d, P.D, p, a0, a1);
    clamp_range(w.h, P.H, p, b0, b1);
    bool first = true;
    for (int a = a0; a <= a1; ++a)
      for (int b = b0; b <= b1; ++b) {
        if (!mbar_wait(R.empty + R.s, R.ph ^ 1, S.err, 34)) return;
        mbar_expect_tx(R.full + R.s, S.g_bytes + (first ? S.x_bytes : 0u));
        bulk_g2s(R.addr(), gn + (long long)(a * Hp + b) * Wp * P.C, S.g_bytes, R.full + R.s);
        if (first) bulk_g2s(R.addr() + S.g_bytes, xn + (long long)w.row * P.W * P.C, S.x_bytes, R.full + R.s);
        first = false;
        R.advance();
      }
  }
}

template <typename T, int CONS, int U>
__global__ void __launch_bounds__(CONS + 32, 1) inorm_bwd_stats_stream_kernel(const T* __restrict__ gy, const T* __restrict__ x,
                                                                               const float* __restrict__ mean,
                                                                               const float* __restrict__ rstd,
                                                                               double* __restrict__ sums, const StreamP S) {
  pdl_launch_dependents();
  extern __shared__ __align__(128) unsigned char smem[];
  const NormP& P = S.P;
  Ring R = ring_setup<CONS>(smem, S);
  pdl_wait();                                  // barriers are set up: from here on global memory is touched
  const int n = blockIdx.y, t = threadIdx.x;
  const int G = P.G, cg = t & (G - 1), wl = t >> S.lgG, wpp = CONS >> S.lgG, lane = t & 31;
  F8 s, ss;
#pragma unroll
  for (int i = 0; i < 4; ++i) { s[i] = bc2(0.f); ss[i] = bc2(0.f); }
  if (t >= CONS) {
    if (t == CONS) bwd_producer<T>(R, S, gy, x, n);
  } else {
    F8 nmu, rs;
    load_pairs(mean + n * P.C + cg * 8, nmu, -1.f);
    load_pairs(rstd + n * P.C + cg * 8, rs, 1.f);
    const float nslope = norm_neg_slope(P);
    const uint32_t esz = sizeof(T), cgoff = (uint32_t)cg * 8 * esz;
    uint32_t off[U];
#pragma unroll
    for (int u = 0; u < U; ++u) off[u] = (uint32_t)((wl + u * wpp) * P.C) * esz + cgoff;
    const int rows = P.D * P.H;
    for (RowWalk w(blockIdx.x, gridDim.x, P.H); w.row < rows; w.next()) {
      F8 g[U];
      S8<T> xv[U];
      if (!bwd_gather_row<T, U>(R, S, w.d, w.h, off, wl, wpp, lane, cgoff, g, xv)) break;
#pragma unroll
      for (int u = 0; u < U; ++u) {
        if (wl + u * wpp < P.W) {
          F8 v;
          xv[u].unpack(v);
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const float2 xh = __fmul2_rn(__fadd2_rn(v[i], nmu[i]), rs[i]);
            const float2 dy = act_grad2(xh, g[u][i], nslope);
            s[i] = __fadd2_rn(s[i], dy);
            ss[i] = __ffma2_rn(dy, xh, ss[i]);
          }
        }
      }
    }
  }
  // every copy that was issued has been consumed: the ring memory is free to serve as reduction scratch
  __syncthreads();
  float* sm = reinterpret_cast<float*>(smem + kHdrBytes);        // [2][CONS * 8]
  if (t < CONS) {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      *reinterpret_cast<float2*>(sm + t * 8 + 2 * i) = s[i];
      *reinterpret_cast<float2*>(sm + CONS * 8 + t * 8 + 2 * i) = ss[i];
    }
  }
  __syncthreads();
  for (int u = t; u < P.C; u += CONS + 32) {
    const int g = u >> 3, j = u & 7;
    double a = 0.0, b = 0.0;
    for (int r = 0; r < wpp; ++r) {
      a += (double)sm[(r * G + g) * 8 + j];
      b += (double)sm[CONS * 8 + (r * G + g) * 8 + j];
    }
    atomicAdd(sums + ((long long)n * P.C + u) * 2 + 0, a);
    atomicAdd(sums + ((long long)n * P.C + u) * 2 + 1, b);
  }
}

template <typename T, int CONS, int U>
__global__ void __launch_bounds__(CONS + 32, 1) inorm_bwd_stream_kernel(const T* __restrict__ gy, const T* __restrict__ x,
                                                                         const float* __restrict__ mean, const float* __restrict__ rstd,
                                                                         const double* __restrict__ sums, T* __restrict__ dx,
                                                                         T* __restrict__ dres, const StreamP S) {
  pdl_launch_dependents();
  extern __shared__ __align__(128) unsigned char smem[];
  const NormP& P = S.P;
  Ring R = ring_setup<CONS>(smem, S);
  pdl_wait();                                  // barriers are set up: from here on global memory is touched
  const int n = blockIdx.y, t = threadIdx.x;
  if (t >= CONS) {
    if (t == CONS) bwd_producer<T>(R, S, gy, x, n);
    return;
  }
  const int G = P.G, cg = t & (G - 1), wl = t >> S.lgG, wpp = CONS >> S.lgG, lane = t & 31;
  // dx = rs * (dy - m1 - xh * m2)
  F8 nmu, rs, nm1, nm2;
  load_pairs(mean + n * P.C + cg * 8, nmu, -1.f);
  load_pairs(rstd + n * P.C + cg * 8, rs, 1.f);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    float m[4] = {0.f, 0.f, 0.f, 0.f};
    if (P.use_running != 1) {
      const double* sp = sums + ((long long)n * P.C + cg * 8 + 2 * i) * 2;
      const double inv = 1.0 / (double)P.V;
#pragma unroll
      for (int q = 0; q < 4; ++q) m[q] = (float)(sp[q] * inv);
    }
    nm1[i] = make_float2(-m[0], -m[2]);
    nm2[i] = make_float2(-m[1], -m[3]);
  }
  const float nslope = norm_neg_slope(P);
  const uint32_t esz = sizeof(T), cgoff = (uint32_t)cg * 8 * esz;
  uint32_t off[U];
#pragma unroll
  for (int u = 0; u < U; ++u) off[u] = (uint32_t)((wl + u * wpp) * P.C) * esz + cgoff;
  const int rp = dres ? P.res_pad : 0;
  const int Dr = P.D + 2 * rp, Hr = P.H + 2 * rp, Wr = P.W + 2 * rp;
  char* dxn = reinterpret_cast<char*>(dx + (long long)n * P.V * P.C);
  T* drn = dres ? dres + (long long)n * Dr * Hr * Wr * P.C : nullptr;
  const long long xrow_bytes = (long long)P.W * P.C * esz;
  const int rows = P.D * P.H;
  for (RowWalk w(blockIdx.x, gridDim.x, P.H); w.row < rows; w.next()) {
    const int d = w.d, h = w.h;
    F8 g[U];
    S8<T> xv[U];
    if (!bwd_gather_row<T, U>(R, S, d, h, off, wl, wpp, lane, cgoff, g, xv)) return;
    char* dxr = dxn + w.row * xrow_bytes;
    char* drr = drn ? reinterpret_cast<char*>(drn + ((long long)((d + rp) * Hr + h + rp) * Wr + rp) * P.C) : nullptr;
#pragma unroll
    for (int u = 0; u < U; ++u) {
      if (wl + u * wpp < P.W) {
        if (drr) store8<T>(reinterpret_cast<T*>(drr + off[u]), g[u]);
        F8 v;
        xv[u].unpack(v);
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const float2 xh = __fmul2_rn(__fadd2_rn(v[i], nmu[i]), rs[i]);
          const float2 dy = act_grad2(xh, g[u][i], nslope);
          v[i] = __fmul2_rn(__fadd2_rn(__ffma2_rn(xh, nm2[i], dy), nm1[i]), rs[i]);
        }
        store8<T>(reinterpret_cast<T*>(dxr + off[u]), v);
      }
    }
    if (drn && rp > 0) {
      // zero halo of the residual gradient: the w-halo of this row and the halo rows that clamp onto it
      F8 z;
#pragma unroll
      for (int i = 0; i < 4; ++i) z[i] = bc2(0.f);
      T* drc = drn + cg * 8;
      int a0, a1, b0, b1;
      clamp_range(d, P.D, rp, a0, a1);
      clamp_range(h, P.H, rp, b0, b1);
      for (int a = a0; a <= a1; ++a)
        for (int b = b0; b <= b1; ++b) {
          T* zr = drc + (long long)(a * Hr + b) * Wr * P.C;
          if (a == d + rp && b == h + rp) {
            for (int rw = wl; rw < rp; rw += wpp) {
              store8<T>(zr + (long long)rw * P.C, z);
              store8<T>(zr + (long long)(Wr - 1 - rw) * P.C, z);
            }
          } else {
            for (int rw = wl; rw < Wr; rw += wpp) store8<T>(zr + (long long)rw * P.C, z);
          }
        }
    }
  }
}

// ------------------------------------------------------------------------------------------------ host side
inline bool stream_enabled() {
  static int on = -1;
  if (on < 0) { const char* e = getenv("MRA_NORM_STREAM"); on = (e && atoi(e) == 0) ? 0 : 1; }
  return on != 0;
}

// A tensor without halo is one contiguous run of positions per sample: re-cut it into rows of ~1024 items so that
// every stage is a decent bulk copy whatever the original W was (15^3 PatchGAN maps, 2^3 UNet bottlenecks, ...).
inline void recut_rows(NormP& P) {
  const long long target = 1024 / P.G > 0 ? 1024 / P.G : 1;
  long long w = 1;
  for (long long c = target; c >= 1; --c)
    if (P.V % c == 0) { w = c; break; }
  if (w <= P.W && (long long)P.W * P.G <= 2048) return;          // the natural row is at least as good
  P.D = 1; P.H = (int)(P.V / w); P.W = (int)w;
}

// 8 consumer warps x 4 items per thread and row beat 16 x 2 on B200 (fewer per-row instructions per item, no spills)
constexpr int kConsumers = 256;
inline int stream_consumers() { return kConsumers; }

template <typename T>
inline bool stream_plan(const mra_norm_desc& d, int vec, bool bwd, bool has_res, StreamP& S) {
  if (!stream_enabled()) return false;
  NormP P = make_norm_params(d, vec);
  const int lg = norm_fast_lg(P, vec);
  if (lg < 0) return false;
  if (d.act == MRA_ACT_LRELU && !(d.slope >= 0.f && d.slope <= 1.f)) return false;   // act(x) = max(x, slope x)
  if (d.pad == 0 && !has_res) recut_rows(P);
  if ((long long)P.D * P.H >= (1ll << 31)) return false;
  const int cons = stream_consumers();
  const long long xb = (long long)P.W * P.C * (long long)sizeof(T);
  long long gb = 0;
  if (bwd) gb = (long long)(P.W + 2 * P.pad) * P.C * (long long)sizeof(T);
  else if (has_res) gb = xb;
  long long stage = (xb + gb + 127) / 128 * 128;
  int stages = (int)((kSmemBudget - kHdrBytes) / stage);
  if (stages > kMaxStages) stages = kMaxStages;
  if (stages < 2) return false;
  if (bwd && (long long)P.W * P.G > 4 * cons) return false;                              // U <= 4 items per thread and row
  if (bwd && (long long)stages * stage < (long long)cons * 8 * 2 * 4) return false;      // reduction scratch
  S.P = P; S.lgG = lg; S.stages = stages;
  S.x_bytes = (uint32_t)xb; S.g_bytes = (uint32_t)gb; S.stage_bytes = (uint32_t)stage;
  S.err = tc::tc_err_flag();
  return true;
}
inline dim3 stream_grid(const StreamP& S) {
  long long bx = num_sms() / S.P.N;
  const long long rows = (long long)S.P.D * S.P.H;
  if (bx > rows) bx = rows;
  if (bx < 1) bx = 1;
  return dim3((unsigned)bx, S.P.N);
}
inline size_t stream_smem(const StreamP& S) { return kHdrBytes + (size_t)S.stages * S.stage_bytes; }

template <typename K> inline int stream_attr(K kernel) {
  MRA_CHECK_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(kSmemBudget + kHdrBytes)));
  return 0;
}

template <typename T, int CONS>
int norm_fwd_stream_launch(const StreamP& S, const void* x, const double* stats, const void* res, void* y, float* mean,
                           float* rstd, float* rm, float* rv, cudaStream_t st) {
  static bool attr = false;
  if (!attr) { if (int rc = stream_attr(inorm_fwd_stream_kernel<T, CONS>)) return rc; attr = true; }
  MRA_CHECK_CUDA(launch_pdl(inorm_fwd_stream_kernel<T, CONS>, stream_grid(S), dim3(CONS + 32), stream_smem(S), st, 1,
                            reinterpret_cast<const T*>(x), stats, mean, rstd, rm, rv, reinterpret_cast<const T*>(res),
                            reinterpret_cast<T*>(y), S));
  MRA_LAUNCH_CHECK();
  return 0;
}

template <typename T, int VEC>
int norm_fwd_launch_v2(const mra_norm_desc& d, const void* x, const double* stats, const void* res, void* y,
                       float* mean, float* rstd, float* rm, float* rv, cudaStream_t st) {
  StreamP S;
  if (VEC != 8 || !stream_plan<T>(d, VEC, false, res != nullptr, S))
    return norm_fwd_launch<T, VEC>(d, x, stats, res, y, mean, rstd, rm, rv, st);
  return norm_fwd_stream_launch<T, kConsumers>(S, x, stats, res, y, mean, rstd, rm, rv, st);
}

template <typename T, int CONS, int U>
int norm_bwd_stream_launch(const mra_norm_desc& d, const StreamP& S, const void* gy, const void* x, const float* mean,
                           const float* rstd, void* dx, void* dres, double* sums, cudaStream_t st, int phases) {
  static bool attr = false;
  if (!attr) {
    if (int rc = stream_attr(inorm_bwd_stats_stream_kernel<T, CONS, U>)) return rc;
    if (int rc = stream_attr(inorm_bwd_stream_kernel<T, CONS, U>)) return rc;
    attr = true;
  }
  if (d.use_running != 1 && (phases & 1)) {
    MRA_CHECK_CUDA(cudaMemsetAsync(sums, 0, sizeof(double) * 2 * d.n * d.c, st));
    MRA_CHECK_CUDA(launch_pdl(inorm_bwd_stats_stream_kernel<T, CONS, U>, stream_grid(S), dim3(CONS + 32), stream_smem(S), st, 1,
                              reinterpret_cast<const T*>(gy), reinterpret_cast<const T*>(x), mean, rstd, sums, S));
    MRA_LAUNCH_CHECK();
  }
  if (!(phases & 2)) return 0;
  MRA_CHECK_CUDA(launch_pdl(inorm_bwd_stream_kernel<T, CONS, U>, stream_grid(S), dim3(CONS + 32), stream_smem(S), st, 1,
                            reinterpret_cast<const T*>(gy), reinterpret_cast<const T*>(x), mean, rstd,
                            static_cast<const double*>(sums), reinterpret_cast<T*>(dx), reinterpret_cast<T*>(dres), S));
  MRA_LAUNCH_CHECK();
  return 0;
}

template <typename T, int CONS>
int norm_bwd_stream_pick(const mra_norm_desc& d, const StreamP& S, const void* gy, const void* x, const float* mean,
                         const float* rstd, void* dx, void* dres, double* sums, cudaStream_t st, int phases) {
  const int wpp = CONS >> S.lgG;
  const int u = (S.P.W + wpp - 1) / wpp;
  if (u <= 1) return norm_bwd_stream_launch<T, CONS, 1>(d, S, gy, x, mean, rstd, dx, dres, sums, st, phases);
  if (u <= 2) return norm_bwd_stream_launch<T, CONS, 2>(d, S, gy, x, mean, rstd, dx, dres, sums, st, phases);
  return norm_bwd_stream_launch<T, CONS, 4>(d, S, gy, x, mean, rstd, dx, dres, sums, st, phases);
}

template <typename T, int VEC>
int norm_bwd_launch_v2(const mra_norm_desc& d, const void* gy, const void* x, const float* mean, const float* rstd,
                       void* dx, void* dres, double* sums, cudaStream_t st, int phases = 3) {
  StreamP S;
  // a residual gradient keeps the tensor's own geometry (its halo is written row by row); `want_res` keeps both
  // phases of a split call on the same row geometry
  if (VEC != 8 || !stream_plan<T>(d, VEC, true, dres != nullptr || d.res_pad >= 0, S))
    return norm_bwd_launch<T, VEC>(d, gy, x, mean, rstd, dx, dres, sums, st, phases);
  return norm_bwd_stream_pick<T, kConsumers>(d, S, gy, x, mean, rstd, dx, dres, sums, st, phases);
}

}  // namespace ns
}  // namespace mra
