// Host-side planning of the implicit-GEMM convolution launches (pure C++, no CUDA calls).
//
// Every conv-family op (Conv3d / ConvTranspose3d, fprop / dgrad / wgrad) is lowered to one of two
// generic device computations:
//
//   GATHER   out[n, l*ostep + o0, cn] = sum_{tap, ck} A[n, l*astep + tap.d, ck] * B[tap.widx][cn][ck]
//            (l ranges over a "launch space" of positions; out-of-range A reads are zero)
//              conv  fprop : one launch, astep = stride, tap.d = t - pad,        B = w
//              convT dgrad : same with A = dy,                                   B = wT
//              conv  dgrad : A = dy, B = wT; stride 1: tap.d = pad - t; stride 2: one launch per
//              convT fprop : A = x,  B = w;  output parity phase r (o = 2l + r) holding the taps with
//                                            (r + pad - t) even, tap.d = (r + pad - t)/2, astep = 1
//
//   WGRAD    dW[tap][cm][cn] = sum_{n, q} Mop[n, pos_m, cm] * Nop[n, pos_n, cn]
//            where one operand is read at the dense position q and the other at q*sstep + tap.d:
//              conv  : M = dy (dense, q over y), N = x  (shifted, tap.d = t - pad)
//              convT : M = dy (shifted),          N = x  (dense, q over x)
//
// mra_conv_plan_describe() serialises a plan as int32 words (see describe()) so the CPU tests can
// emulate it against torch's conv3d / conv_transpose3d without a GPU.
#pragma once
#include <stdint.h>
#include <vector>

#include "../../include/mra_gan_b200.h"

namespace mra {

struct Tap {
  int dd, dh, dw;  // coordinate offset added to (l * astep)
  int widx;        // weight slab index (kd*k + kh)*k + kw
};

struct GatherLaunch {
  int o0[3];    // output coordinate offset (d, h, w)
  int ostep;    // output coordinate step
  int dims[3];  // launch-space extent (d, h, w)
  int astep;    // A coordinate step
  std::vector<Tap> taps;
  int box[3];   // tile box (d, h, w), product 128
};

struct GatherPlan {
  int n, ck, cn;      // batch, reduction channels, output channels
  int adims[3];       // A spatial dims
  int odims[3];       // output spatial dims
  std::vector<GatherLaunch> launches;
};

struct WgradPlan {
  int n, cm, cn;      // dW is [taps][cm][cn] with cm = cout, cn = cin
  int qdims[3];       // dense position space
  int mdims[3], ndims[3];
  int sstep;          // stride applied to q for the shifted operand
  int m_is_shifted;   // 1: dy (M side) is the shifted operand (ConvTranspose3d)
  std::vector<Tap> taps;
  int box[3];         // K-block box (d, h, w), product 64
};

// Pick power-of-two box extents (d,h,w) with product `total` minimising the padded volume of the
// launch space; ties prefer a longer w (contiguous) run.  `maxw` bounds the w extent (TMA boxDim<=256
// after multiplying by the element stride).
inline void choose_box(const int dims[3], int total, int maxw, int box[3]) {
  long long best = -1;
  int bb[3] = {1, 1, total};
  for (int bd = 1; bd <= total; bd <<= 1)
    for (int bh = 1; bd * bh <= total; bh <<= 1) {
      int bw = total / (bd * bh);
      if (bw > maxw) continue;
      long long pd = (long long)((dims[0] + bd - 1) / bd) * bd;
      long long ph = (long long)((dims[1] + bh - 1) / bh) * bh;
      long long pw = (long long)((dims[2] + bw - 1) / bw) * bw;
      long long vol = pd * ph * pw;
      bool better = best < 0 || vol < best ||
                    (vol == best && (bw > bb[2] || (bw == bb[2] && bh > bb[1])));
      if (better) { best = vol; bb[0] = bd; bb[1] = bh; bb[2] = bw; }
    }
  box[0] = bb[0]; box[1] = bb[1]; box[2] = bb[2];
}

inline bool build_gather_plan(const mra_conv_desc& d, int which, GatherPlan& P) {
  // which: 0 = fprop (A = x, out = y), 1 = dgrad (A = dy, out = dx)
  const int k = d.k, s = d.stride, p = d.pad;
  if (s != 1 && s != 2) return false;
  const bool fprop = which == 0;
  P.n = d.n;
  P.ck = fprop ? d.cin : d.cout;
  P.cn = fprop ? d.cout : d.cin;
  const int xin[3] = {d.din, d.hin, d.win}, yout[3] = {d.dout, d.hout, d.wout};
  for (int i = 0; i < 3; ++i) {
    P.adims[i] = fprop ? xin[i] : yout[i];
    P.odims[i] = fprop ? yout[i] : xin[i];
  }
  // direct addressing: A coord = o*s - p + t.  Used by conv fprop and convT dgrad.
  const bool direct = (fprop && !d.transposed) || (!fprop && d.transposed);
  P.launches.clear();
  if (direct) {
    GatherLaunch L;
    for (int i = 0; i < 3; ++i) { L.o0[i] = 0; L.dims[i] = P.odims[i]; }
    L.ostep = 1; L.astep = s;
    for (int kd = 0; kd < k; ++kd) for (int kh = 0; kh < k; ++kh) for (int kw = 0; kw < k; ++kw)
      L.taps.push_back(Tap{kd - p, kh - p, kw - p, (kd * k + kh) * k + kw});
    choose_box(L.dims, 128, 256 / s, L.box);
    P.launches.push_back(L);
  } else if (s == 1) {
    GatherLaunch L;
    for (int i = 0; i < 3; ++i) { L.o0[i] = 0; L.dims[i] = P.odims[i]; }
    L.ostep = 1; L.astep = 1;
    for (int kd = 0; kd < k; ++kd) for (int kh = 0; kh < k; ++kh) for (int kw = 0; kw < k; ++kw)
      L.taps.push_back(Tap{p - kd, p - kh, p - kw, (kd * k + kh) * k + kw});
    choose_box(L.dims, 128, 256, L.box);
    P.launches.push_back(L);
  } else {
    for (int rd = 0; rd < 2; ++rd) for (int rh = 0; rh < 2; ++rh) for (int rw = 0; rw < 2; ++rw) {
      const int r[3] = {rd, rh, rw};
      GatherLaunch L;
      bool empty = false;
      for (int i = 0; i < 3; ++i) {
        L.o0[i] = r[i];
        L.dims[i] = (P.odims[i] - r[i] + 1) / 2;
        if (L.dims[i] <= 0) empty = true;
      }
      if (empty) continue;
      L.ostep = 2; L.astep = 1;
      for (int kd = 0; kd < k; ++kd) {
        if ((rd + p - kd) & 1) continue;
        for (int kh = 0; kh < k; ++kh) {
          if ((rh + p - kh) & 1) continue;
          for (int kw = 0; kw < k; ++kw) {
            if ((rw + p - kw) & 1) continue;
            L.taps.push_back(Tap{(rd + p - kd) / 2, (rh + p - kh) / 2, (rw + p - kw) / 2,
                                 (kd * k + kh) * k + kw});
          }
        }
      }
      choose_box(L.dims, 128, 256, L.box);
      P.launches.push_back(L);
    }
  }
  return true;
}

inline bool build_wgrad_plan(const mra_conv_desc& d, WgradPlan& P) {
  const int k = d.k, s = d.stride, p = d.pad;
  if (s != 1 && s != 2) return false;
  P.n = d.n; P.cm = d.cout; P.cn = d.cin;
  const int xin[3] = {d.din, d.hin, d.win}, yout[3] = {d.dout, d.hout, d.wout};
  for (int i = 0; i < 3; ++i) {
    P.mdims[i] = yout[i];          // M operand = dy
    P.ndims[i] = xin[i];           // N operand = x
    P.qdims[i] = d.transposed ? xin[i] : yout[i];
  }
  P.sstep = s;
  P.m_is_shifted = d.transposed ? 1 : 0;
  P.taps.clear();
  for (int kd = 0; kd < k; ++kd) for (int kh = 0; kh < k; ++kh) for (int kw = 0; kw < k; ++kw)
    P.taps.push_back(Tap{kd - p, kh - p, kw - p, (kd * k + kh) * k + kw});
  choose_box(P.qdims, 64, 256 / s, P.box);
  return true;
}

// Serialisation for mra_conv_plan_describe():
//  gather: [0, n, ck, cn, adims[3], odims[3], nlaunch, then per launch:
//           o0[3], ostep, dims[3], astep, box[3], ntaps, ntaps x (dd, dh, dw, widx)]
//  wgrad : [1, n, cm, cn, qdims[3], mdims[3], ndims[3], sstep, m_is_shifted, box[3], ntaps,
//           ntaps x (dd, dh, dw, widx)]
inline int describe(const GatherPlan& P, int32_t* out, int cap) {
  std::vector<int32_t> v = {0, P.n, P.ck, P.cn};
  for (int i = 0; i < 3; ++i) v.push_back(P.adims[i]);
  for (int i = 0; i < 3; ++i) v.push_back(P.odims[i]);
  v.push_back((int)P.launches.size());
  for (const auto& L : P.launches) {
    for (int i = 0; i < 3; ++i) v.push_back(L.o0[i]);
    v.push_back(L.ostep);
    for (int i = 0; i < 3; ++i) v.push_back(L.dims[i]);
    v.push_back(L.astep);
    for (int i = 0; i < 3; ++i) v.push_back(L.box[i]);
    v.push_back((int)L.taps.size());
    for (const auto& t : L.taps) { v.push_back(t.dd); v.push_back(t.dh); v.push_back(t.dw); v.push_back(t.widx); }
  }
  if ((int)v.size() > cap) return -2;
  for (size_t i = 0; i < v.size(); ++i) out[i] = v[i];
  return (int)v.size();
}
inline int describe(const WgradPlan& P, int32_t* out, int cap) {
  std::vector<int32_t> v = {1, P.n, P.cm, P.cn};
  for (int i = 0; i < 3; ++i) v.push_back(P.qdims[i]);
  for (int i = 0; i < 3; ++i) v.push_back(P.mdims[i]);
  for (int i = 0; i < 3; ++i) v.push_back(P.ndims[i]);
  v.push_back(P.sstep); v.push_back(P.m_is_shifted);
  for (int i = 0; i < 3; ++i) v.push_back(P.box[i]);
  v.push_back((int)P.taps.size());
  for (const auto& t : P.taps) { v.push_back(t.dd); v.push_back(t.dh); v.push_back(t.dw); v.push_back(t.widx); }
  if ((int)v.size() > cap) return -2;
  for (size_t i = 0; i < v.size(); ++i) out[i] = v[i];
  return (int)v.size();
}

}  // namespace mra
