// Column gather kernel: stride-1 GATHER launches whose taps run along d only (kernel (kd, 1, 1)) with few channels
// -- the channel-expanded stem / head lowerings of conv_special.cuh (64 -> 64 channels, kd = 7).
//
// For these the per-tap kernel is hopeless: every output tile re-loads kd input tiles for 28 small (N = 64) MMAs and
// pays a barrier handshake per tap.  Here a CTA walks a COLUMN of output tiles along d and keeps
//   * all kd weight slabs resident in shared memory (loaded once per CTA), and
//   * a ring of input planes (one 16 x 8 position tile of one d-plane, 64 channels = 16 KB): stepping to the next
//     output plane loads ONE new plane instead of kd, and costs ONE barrier wait for 4 * kd back-to-back MMAs.
// Columns are cut into segments along d so that the persistent CTAs get a balanced number of work units.
//
// Dual-plane mode (64 output channels): tcgen05.mma runs N = 64 at 54.5 clk per 128x64x16 instruction where 32 would
// be ideal (tools/mma_probe.cu: the A-operand read bounds it), N = 128 at the ideal 64.  So a tile covers TWO output
// planes p, p + 1 side by side in a 128-column accumulator: input plane p + dmin + t (t = 0..kd) meets the weight rows
// [W[t] ; W[t-1]] -- with the resident slabs stored in REVERSE tap order these are simply two consecutive slabs, one
// N = 128 descriptor.  Only the first and the last plane of a pair (one tap each) issue N = 64 MMAs.
// Warp roles as in conv_tc.cuh: 0..7 = epilogue, 8 = TMA producer, 9 = TMEM allocator + MMA issuer.
#pragma once
#include "conv_tc.cuh"

namespace mra {
namespace tc {

struct ColP {
  int Dl, Hl, Wl, N;                // launch-space (output) dims
  int tiles_w, tiles_hw;            // 16 x 8 tiles per plane
  int kd, dmin;                     // taps along d: offsets dmin .. dmin + kd - 1
  int Cn, n_tile, kchunks;
  int seg_len, nseg;                // column segments
  int total_units;                  // N * tiles_hw * nseg
  int NPR;                          // plane ring depth (>= kd + 1)
  int ostep, od0, oh0, ow0;
  long long osn, osd, osh, osw;
  void* out;
  int out_bf16;
  const float* bias;
  int act;
  float slope;
  double* stats;
  const void* aux; float aux_nslope;   // norm-backward statistics (EpiArgs::aux)
  int* err;
  int nbuf;                         // accumulator buffers in TMEM (512 / accumulator width, at most 8)
  int dual;                         // 1: dual-plane tiles (accumulator 2 * n_tile = 128 columns)
  uint32_t tmem_cols;
  int16_t twi[kMaxTaps];            // weight slab of tap td
  int debug;
  unsigned long long* dbg;
};

template <int kMode>
__global__ void __launch_bounds__(kThreadsGather, 1)
gather_col_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                  const __grid_constant__ ColP P) {
  pdl_launch_dependents();
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const uint32_t slab_bytes = (uint32_t)P.n_tile * 128u;                 // one (tap, channel chunk) weight slab
  const uint32_t w_bytes = slab_bytes * (uint32_t)(P.kd * P.kchunks);
  const uint32_t slot_bytes = kABytes * (uint32_t)P.kchunks;             // one plane: 128 positions x 64 ch per chunk
  uint8_t* wres = smem;
  uint8_t* ring = smem + w_bytes;
  uint64_t* p_full = reinterpret_cast<uint64_t*>(ring + (size_t)P.NPR * slot_bytes);
  uint64_t* p_empty = p_full + P.NPR;
  uint64_t* w_bar = p_empty + P.NPR;
  uint64_t* acc_full = w_bar + 1;                // [nbuf]
  uint64_t* acc_empty = acc_full + P.nbuf;       // [nbuf]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_empty + P.nbuf);
  __shared__ EpiRed epi_red;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  if (threadIdx.x == 0) {
    prefetch_tmap(&tmA);
    prefetch_tmap(&tmB);
    for (int s = 0; s < P.NPR; ++s) { mbar_init(&p_full[s], 1); mbar_init(&p_empty[s], 1); }
    mbar_init(w_bar, 1);
    for (int b = 0; b < P.nbuf; ++b) { mbar_init(&acc_full[b], 1); mbar_init(&acc_empty[b], kEpiWarps); }
    fence_barrier_init();
  }
  if (warp == kMmaWarp) tmem_alloc(tmem_slot, P.tmem_cols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_wait();                                  // prologue done: from here on global memory is touched

  // unit -> (n, hw tile, first plane, number of planes)
  auto unit_coords = [&](int u, int& n, int& h0, int& w0, int& d0, int& len) {
    const int seg = u % P.nseg; u /= P.nseg;
    const int j = u % P.tiles_hw;
    n = u / P.tiles_hw;
    h0 = (j / P.tiles_w) * 16; w0 = (j % P.tiles_w) * 8;
    d0 = seg * P.seg_len;
    len = min(P.seg_len, P.Dl - d0);
  };

  if (warp >= kEpiWarps) {            // third warpgroup: producers, MMA issuer, idle warps -- one setmaxnreg for all of it
  regs_other();
  if (warp == kProdWarp) {
    if (elect_one()) {
      // resident weights: every (tap, chunk) slab once
      mbar_expect_tx(w_bar, w_bytes);
      for (int td = 0; td < P.kd; ++td)
        for (int kc = 0; kc < P.kchunks; ++kc)
          tma_load_2d(wres + (size_t)(P.dual ? kc * P.kd + (P.kd - 1 - td) : td * P.kchunks + kc) * slab_bytes, &tmB, w_bar,
                      kc * 64, (int)P.twi[td] * P.Cn);
      int s = 0;
      uint32_t ph = 0;
      bool ok = true;
      for (int u = blockIdx.x; u < P.total_units && ok; u += gridDim.x) {
        int n, h0, w0, d0, len;
        unit_coords(u, n, h0, w0, d0, len);
        const int nplanes = (P.dual ? 2 * ((len + 1) / 2) : len) + P.kd - 1;
        for (int p = 0; p < nplanes; ++p) {
          if (!mbar_wait(&p_empty[s], ph ^ 1u, P.err, 31)) { ok = false; break; }
          mbar_expect_tx(&p_full[s], slot_bytes);
          for (int kc = 0; kc < P.kchunks; ++kc)
            tma_load_5d(ring + (size_t)s * slot_bytes + (size_t)kc * kABytes, &tmA, &p_full[s], kc * 64, w0, h0,
                        d0 + P.dmin + p, n);
          if (++s == P.NPR) { s = 0; ph ^= 1u; }
        }
      }
    }
  } else if (warp == kMmaWarp) {
    if (elect_one()) {
      const uint32_t idesc = make_idesc(P.n_tile, 0, 0);
      const uint64_t desc0 = desc_kmajor_sw128(0);
      const uint32_t ring_u = smem_u32(ring) >> 4, wres_u = smem_u32(wres) >> 4;
      const uint32_t slot_u = slot_bytes >> 4, slab_u = slab_bytes >> 4, chunk_u = kABytes >> 4;
      const int kd = P.kd, kchunks = P.kchunks, NPR = P.NPR;
      bool ok = mbar_wait(w_bar, 0, P.err, 32);
      tc_fence_after();
      int s_old = 0;                 // ring slot of the oldest plane of the current tile
      int s_new = 0;                 // ring slot (and phase) of the next plane to wait for
      uint32_t ph_new = 0;
      int jt = 0;                    // tiles issued by this CTA
      int buf = 0;                   // accumulator buffer / phase of the next tile
      uint32_t aph = 0;
      const bool prof = (P.debug & 2) != 0;
      long long t_wacc = 0, t_wp = 0, t_begin = prof ? clock64() : 0;
      for (int u = blockIdx.x; u < P.total_units && ok; u += gridDim.x) {
        int n, h0, w0, d0, len;
        unit_coords(u, n, h0, w0, d0, len);
        int have = 0;                // planes of this unit already waited for
        if (P.dual) {
          const uint32_t idesc64 = make_idesc(64, 0, 0), idesc128 = make_idesc(128, 0, 0);
          const int npairs = (len + 1) / 2;
          for (int jp = 0; jp < npairs && ok; ++jp, ++jt) {
            const long long t0 = prof ? clock64() : 0;
            if (!mbar_wait(&acc_empty[buf], aph ^ 1u, P.err, 34)) { ok = false; break; }
            const long long t1 = prof ? clock64() : 0;
            while (have < 2 * jp + kd + 1) {    // a pair reads kd + 1 planes; two of them are new after the first pair
              if (!mbar_wait(&p_full[s_new], ph_new, P.err, 35)) { ok = false; break; }
              if (++s_new == NPR) { s_new = 0; ph_new ^= 1u; }
              ++have;
            }
            if (prof) { t_wacc += t1 - t0; t_wp += clock64() - t1; }
            if (!ok) break;
            tc_fence_after();
            const uint32_t d_tmem = tmem_base + (uint32_t)(buf * 128);
            int s_last = s_old + kd;
            if (s_last >= NPR) s_last -= NPR;
            // the two end planes first (one tap each, N = 64, they initialise their half of the accumulator) ...
            for (int kc = 0; kc < kchunks; ++kc) {
              const uint32_t a0 = ring_u + (uint32_t)s_old * slot_u + (uint32_t)kc * chunk_u;
              const uint32_t a1 = ring_u + (uint32_t)s_last * slot_u + (uint32_t)kc * chunk_u;
              const uint32_t b0 = wres_u + (uint32_t)(kc * kd + (kd - 1)) * slab_u;      // W[0]      -> plane p, columns [0, 64)
              const uint32_t b1 = wres_u + (uint32_t)(kc * kd) * slab_u;                 // W[kd - 1] -> plane p + 1, columns [64, 128)
              const uint64_t ad0 = desc0 | (uint64_t)(a0 & 0x3FFFu), bd0 = desc0 | (uint64_t)(b0 & 0x3FFFu);
              const uint64_t ad1 = desc0 | (uint64_t)(a1 & 0x3FFFu), bd1 = desc0 | (uint64_t)(b1 & 0x3FFFu);
              umma_f16(d_tmem, ad0, bd0, idesc64, (uint32_t)(kc != 0));
              umma_f16(d_tmem, ad0 + 2, bd0 + 2, idesc64, 1u);
              umma_f16(d_tmem, ad0 + 4, bd0 + 4, idesc64, 1u);
              umma_f16(d_tmem, ad0 + 6, bd0 + 6, idesc64, 1u);
              umma_f16(d_tmem + 64, ad1, bd1, idesc64, (uint32_t)(kc != 0));
              umma_f16(d_tmem + 64, ad1 + 2, bd1 + 2, idesc64, 1u);
              umma_f16(d_tmem + 64, ad1 + 4, bd1 + 4, idesc64, 1u);
              umma_f16(d_tmem + 64, ad1 + 6, bd1 + 6, idesc64, 1u);
            }
            // ... then the kd - 1 inner planes: rows [W[t] ; W[t - 1]] = two consecutive reversed slabs, N = 128
            int s = s_old;
            for (int t = 1; t < kd; ++t) {
              if (++s == NPR) s = 0;
              for (int kc = 0; kc < kchunks; ++kc) {
                const uint32_t a_u = ring_u + (uint32_t)s * slot_u + (uint32_t)kc * chunk_u;
                const uint32_t b_u = wres_u + (uint32_t)(kc * kd + (kd - 1 - t)) * slab_u;
                const uint64_t ad = desc0 | (uint64_t)(a_u & 0x3FFFu);
                const uint64_t bd = desc0 | (uint64_t)(b_u & 0x3FFFu);
                umma_f16(d_tmem, ad, bd, idesc128, 1u);
                umma_f16(d_tmem, ad + 2, bd + 2, idesc128, 1u);
                umma_f16(d_tmem, ad + 4, bd + 4, idesc128, 1u);
                umma_f16(d_tmem, ad + 6, bd + 6, idesc128, 1u);
              }
            }
            umma_commit(&acc_full[buf]);
            if (++buf == P.nbuf) { buf = 0; aph ^= 1u; }
            // the two oldest planes are dead after this pair; after the unit's last pair so are the other kd - 1
            const int nrel = (jp == npairs - 1) ? kd + 1 : 2;
            for (int r = 0; r < nrel; ++r) {
              umma_commit(&p_empty[s_old]);
              if (++s_old == NPR) s_old = 0;
            }
          }
          continue;
        }
        for (int j = 0; j < len && ok; ++j, ++jt) {
          const long long t0 = prof ? clock64() : 0;
          if (!mbar_wait(&acc_empty[buf], aph ^ 1u, P.err, 34)) { ok = false; break; }
          const long long t1 = prof ? clock64() : 0;
          while (have < j + kd) {    // first tile of a unit: kd planes; afterwards one new plane per tile
            if (!mbar_wait(&p_full[s_new], ph_new, P.err, 35)) { ok = false; break; }
            if (++s_new == NPR) { s_new = 0; ph_new ^= 1u; }
            ++have;
          }
          if (prof) { t_wacc += t1 - t0; t_wp += clock64() - t1; }
          if (!ok) break;
          tc_fence_after();
          const uint32_t d_tmem = tmem_base + (uint32_t)(buf * P.n_tile);
          uint32_t acc = 0;
          int s = s_old;
          uint32_t b_u = wres_u;
          for (int td = 0; td < kd; ++td) {
            uint32_t a_u = ring_u + (uint32_t)s * slot_u;
            for (int kc = 0; kc < kchunks; ++kc, a_u += chunk_u, b_u += slab_u) {
              const uint64_t ad = desc0 | (uint64_t)(a_u & 0x3FFFu);
              const uint64_t bd = desc0 | (uint64_t)(b_u & 0x3FFFu);
              umma_f16(d_tmem, ad, bd, idesc, acc);
              umma_f16(d_tmem, ad + 2, bd + 2, idesc, 1u);
              umma_f16(d_tmem, ad + 4, bd + 4, idesc, 1u);
              umma_f16(d_tmem, ad + 6, bd + 6, idesc, 1u);
              acc = 1u;
            }
            if (++s == NPR) s = 0;
          }
          umma_commit(&acc_full[buf]);
          if (++buf == P.nbuf) { buf = 0; aph ^= 1u; }
          // the oldest plane is dead after this tile; after the unit's last tile so are the other kd - 1
          const int nrel = (j == len - 1) ? kd : 1;
          for (int r = 0; r < nrel; ++r) {
            umma_commit(&p_empty[s_old]);
            if (++s_old == NPR) s_old = 0;
          }
        }
      }
      if (prof) {
        atomicAdd(P.dbg + 3, (unsigned long long)(clock64() - t_begin)); atomicAdd(P.dbg + 6, (unsigned long long)t_wacc);
        atomicAdd(P.dbg + 7, (unsigned long long)t_wp); atomicAdd(P.dbg + 5, 1ull); atomicAdd(P.dbg + 1, (unsigned long long)jt);
      }
    }
  }
  } else {
    regs_epilogue();
    const EpiWarp W(warp);
    const int q = W.q;
    const int row = q * 32 + lane;
    const int nchunks = P.n_tile / 32;                    // CHANNEL chunks (dual tiles hold each of them twice)
    const int acc_cols = P.dual ? 2 * P.n_tile : P.n_tile;
    const bool defer = kMode != 0 && nchunks <= 2;
    EpiStats st;
    st.clear();
    float d1[32], d2[32];
#pragma unroll
    for (int i = 0; i < 32; ++i) { d1[i] = 0.f; d2[i] = 0.f; }
    int st_n = -1;
    const EpiArgs E{P.out, P.out_bf16, P.bias, P.act, P.slope, P.stats != nullptr, P.aux, P.aux_nslope};
    int buf = 0;
    uint32_t aph = 0;
    bool ok = true;
    for (int u = blockIdx.x; u < P.total_units && ok; u += gridDim.x) {
      int n, h0, w0, d0, len;
      unit_coords(u, n, h0, w0, d0, len);
      const int lh = h0 + (row >> 3), lw = w0 + (row & 7);
      const bool valid = lw < P.Wl && lh < P.Hl;
      if (kMode != 0 && n != st_n) {
        epilogue_flush_stats(P.stats, st_n, P.Cn, 0, W.q, W.c_begin, 2, nchunks, lane, st, defer, d1, d2, epi_red);
        st_n = n;
      }
      const int ntiles = P.dual ? (len + 1) / 2 : len;
      for (int j = 0; j < ntiles && ok; ++j) {
        const int pl = P.dual ? 2 * j : j;                 // (first) output plane of the tile, relative to d0
        const long long obase = (long long)n * P.osn + (long long)((d0 + pl) * P.ostep + P.od0) * P.osd +
                                (long long)(lh * P.ostep + P.oh0) * P.osh + (long long)(lw * P.ostep + P.ow0) * P.osw;
        AuxRegs ax;
        epilogue_aux_first<kMode>(E, W.c_begin, acc_cols / 32, valid, obase, P.dual != 0, ax);
        ok = mbar_wait(&acc_full[buf], aph, P.err, 33);
        if (!ok) break;
        tc_fence_after();
        const uint32_t t_addr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(buf * acc_cols);
        uint64_t* rel_bar = &acc_empty[buf];
        epilogue_tile<kMode>(E, t_addr, W.c_begin, 2, acc_cols / 32, valid, obase, 0, lane, st, 0, defer, d1, d2, ax, [&]() {
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(rel_bar);
        }, P.dual != 0, (long long)P.ostep * P.osd, valid && pl + 1 < len);
        if (++buf == P.nbuf) { buf = 0; aph ^= 1u; }
      }
    }
    if (kMode != 0) epilogue_flush_stats(P.stats, st_n, P.Cn, 0, W.q, W.c_begin, 2, nchunks, lane, st, defer, d1, d2, epi_red);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == kMmaWarp) tmem_dealloc(tmem_base, P.tmem_cols);
}

// Eligibility + geometry: unit A step, taps along d only forming a contiguous range, one output-channel tile,
// weights + a plane ring that fit shared memory.
inline bool col_setup(const GatherLaunch& L, int n, int ck, int cn, ColP& P) {
  if (L.astep != 1 || L.taps.size() < 2 || (int)L.taps.size() > kMaxTaps) return false;
  int lo = 1 << 30, hi = -(1 << 30);
  for (const Tap& t : L.taps) {
    if (t.dh != 0 || t.dw != 0) return false;
    if (t.dd < lo) lo = t.dd;
    if (t.dd > hi) hi = t.dd;
  }
  const int kd = hi - lo + 1;
  if (kd != (int)L.taps.size()) return false;
  const int n_tile = pick_n_tile(cn);
  if (n_tile == 0 || n_tile != cn || ck % 64 != 0) return false;
  P.kd = kd; P.dmin = lo;
  for (int i = 0; i < kMaxTaps; ++i) P.twi[i] = -1;
  for (const Tap& t : L.taps) {
    if (P.twi[t.dd - lo] != -1) return false;
    P.twi[t.dd - lo] = (int16_t)t.widx;
  }
  P.Dl = L.dims[0]; P.Hl = L.dims[1]; P.Wl = L.dims[2]; P.N = n;
  P.Cn = cn; P.n_tile = n_tile; P.kchunks = ck / 64;
  P.tiles_w = (P.Wl + 7) / 8; P.tiles_hw = ((P.Hl + 15) / 16) * P.tiles_w;
  const size_t budget = kSmemLimit - 2048 - kEpiRedBytes;
  const size_t w_bytes = (size_t)kd * P.kchunks * n_tile * 128;
  const size_t slot = (size_t)kABytes * P.kchunks;
  if (w_bytes + (size_t)(kd + 1) * slot > budget) return false;
  int npr = (int)((budget - w_bytes) / slot);
  if (npr > kd + 6) npr = kd + 6;
  P.NPR = npr;
  // dual-plane tiles: two output planes per 128-column accumulator (needs kd + 1 resident planes + 2 in flight)
  P.dual = (n_tile == 64 && P.Dl >= 2 && npr >= kd + 3 && getenv("MRA_COL_NODUAL") == nullptr) ? 1 : 0;
  // segments: aim at >= 6 units per SM, but keep them long enough to amortise the kd - 1 warm-up planes
  const long long cols = (long long)n * P.tiles_hw;
  long long nseg = ((long long)num_sms() * 6 + cols - 1) / cols;
  const long long max_seg = P.Dl / (2 * kd) > 1 ? P.Dl / (2 * kd) : 1;
  if (nseg > max_seg) nseg = max_seg;
  if (nseg < 1) nseg = 1;
  P.seg_len = (int)((P.Dl + nseg - 1) / nseg);
  if (P.dual && (P.seg_len & 1)) ++P.seg_len;           // whole pairs per segment
  P.nseg = (P.Dl + P.seg_len - 1) / P.seg_len;
  const long long units = cols * P.nseg;
  if (units >= (1ll << 31)) return false;
  P.total_units = (int)units;
  return true;
}

inline int run_gather_col(const GatherPlan& plan, const GatherLaunch& L, ColP& P, const GatherRun& R, const CUtensorMap& tmB,
                          cudaStream_t st) {
  static bool attr_set = false;
  if (!attr_set) {
    MRA_CHECK_CUDA(cudaFuncSetAttribute(gather_col_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(kSmemLimit - kEpiRedBytes)));
    MRA_CHECK_CUDA(cudaFuncSetAttribute(gather_col_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(kSmemLimit - kEpiRedBytes)));
    MRA_CHECK_CUDA(cudaFuncSetAttribute(gather_col_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(kSmemLimit - kEpiRedBytes)));
    attr_set = true;
  }
  P.ostep = L.ostep; P.od0 = L.o0[0]; P.oh0 = L.o0[1]; P.ow0 = L.o0[2];
  P.osw = plan.cn;
  P.osh = (long long)plan.odims[2] * P.osw;
  P.osd = (long long)plan.odims[1] * P.osh;
  P.osn = (long long)plan.odims[0] * P.osd;
  P.out = R.out; P.out_bf16 = R.out_bf16;
  P.bias = R.bias; P.act = R.act; P.slope = R.slope;
  P.stats = R.stats; P.aux = R.aux; P.aux_nslope = R.aux_nslope; P.err = tc_err_flag();
  { const char* e = getenv("MRA_GATHER_DEBUG"); P.debug = e ? atoi(e) : 0; }
  P.dbg = tc_dbg_counters();
  const int acc_cols = P.dual ? 2 * P.n_tile : P.n_tile;
  P.nbuf = 512 / acc_cols > 8 ? 8 : 512 / acc_cols;
  P.tmem_cols = pow2_cols(P.nbuf * acc_cols);
  CUtensorMap tmA;
  if (int rc = make_act_map(&tmA, R.a, plan.n, plan.adims[0], plan.adims[1], plan.adims[2], plan.ck, 8, 16, 1, 1)) return rc;
  const size_t smem = (size_t)P.kd * P.kchunks * P.n_tile * 128 + (size_t)P.NPR * kABytes * P.kchunks + 1024 + 512;
  const int ctas = P.total_units < num_sms() ? P.total_units : num_sms();
  if (P.aux && P.stats) MRA_CHECK_CUDA(launch_pdl(gather_col_kernel<2>, dim3(ctas), dim3(kThreadsGather), smem, st, 1, tmA, tmB, P));
  else if (P.stats) MRA_CHECK_CUDA(launch_pdl(gather_col_kernel<1>, dim3(ctas), dim3(kThreadsGather), smem, st, 1, tmA, tmB, P));
  else MRA_CHECK_CUDA(launch_pdl(gather_col_kernel<0>, dim3(ctas), dim3(kThreadsGather), smem, st, 1, tmA, tmB, P));
  MRA_LAUNCH_CHECK();
  return 0;
}

}  // namespace tc
}  // namespace mra
