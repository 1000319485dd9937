// Phase-halo gather kernel: the stride-2 "dgrad-form" plans of conv_plan.h (Conv3d dgrad, ConvTranspose3d fprop with
// stride 2) -- 8 output parity phases over one launch space, every phase a small box of taps at input offsets
// {0, 1}^3 (k = 3) or {-1, 0, 1}^3 (k = 4).
//
// gather_tc_kernel runs them as (spatial tile, phase) work items and loads one 128-row A tile per tap entry: 18 A
// tiles of 16 KB per 64 input channels and spatial tile for k = 3, although the 8 phases read only 8 distinct shifts of
// ONE input neighbourhood.  Those layers (G.u1 / G.u2 fprop, G.d1 / G.d2 dgrad: 450 .. 800 TFLOP/s) are bound by exactly
// that L2 -> SM traffic (r01: 65 FLOP per byte).  Here a work item is a spatial tile (16 h x 8 w input positions of one
// d-plane) together with ALL the phases of a group:
//   * per 64-channel chunk the input halo planes of the tile (Hb x Wb rows of 128 B, one per d-offset) are loaded ONCE
//     into a ring and every tap of every phase reads them through a row-shifted UMMA descriptor (as gather_halo_kernel
//     does for stride-1 launches): A traffic per tile drops from 18 x 16 KB to 2 x 20 KB per chunk (k = 3);
//   * the accumulators of the group's 4 sub-items live side by side in TMEM (4 x 128 columns = all 512), the channel
//     chunks are the OUTER loop; a sub-item is a pair of w-adjacent phases of a 64-channel output (columns [0, 64) =
//     even w, [64, 128) = odd w, one N = 128 MMA where both phases use the same input offset) or one phase of a
//     128-channel output slice;
//   * the epilogue drains sub-item k while the MMAs of sub-items k + 1 .. run, and the first chunk of the next work
//     item starts on sub-item 0 again: the same overlap as a TMEM accumulator ring.
// Weights are still streamed per (chunk, tap) -- cta_group::2 sharing of the slabs is the next step.
//
// Warp roles: 0..7 epilogue, 8 weight producer, 9 plane producer, 10 TMEM allocator + MMA issuer.
#pragma once
#include "conv_tc_halo.cuh"

namespace mra {
namespace tc {

constexpr int kPhaseMaxEntries = 64;
constexpr int kPhaseMaxGroups = 2;
constexpr int kPhaseSubs = 4;                     // sub-items (accumulators of 128 columns) per work item

struct PhaseP {
  int Dl, Hl, Wl, N;                // launch space = the phases' common extent (input-side grid)
  int tiles_w, tiles_hw;            // 16 x 8 tiles per d-plane
  int Wb, Hb;                       // halo plane extents
  int hmin, wmin;                   // smallest tap offsets in h / w (box origin)
  int sbo, slot_bytes, plane_tx;
  int NP, NB;
  int kchunks;
  int Cn, n_tiles;                  // output channels; 128-channel slices per phase (1 when dual)
  int dual;                         // 1: Cn == 64, sub-items are pairs of w-adjacent phases
  int ngroups;                      // phase groups per spatial tile (1 when dual, else 2: one per d parity)
  int total_items;
  int ostep;
  long long osn, osd, osh, osw;
  void* out;
  int out_bf16;
  const float* bias;
  int act;
  float slope;
  double* stats;
  const void* aux; float aux_nslope;
  int* err;
  int debug;
  unsigned long long* dbg;
  // per group
  int8_t g_nd[kPhaseMaxGroups];                       // halo planes (d-offsets) of the group
  int8_t g_dd[kPhaseMaxGroups][4];
  // per sub-item s = group * 4 + k
  int16_t s_ent0[kPhaseMaxGroups * kPhaseSubs + 1];   // entries [s_ent0[s], s_ent0[s + 1])
  int8_t s_od[kPhaseMaxGroups * kPhaseSubs], s_oh[kPhaseMaxGroups * kPhaseSubs], s_ow[kPhaseMaxGroups * kPhaseSubs];
  // per entry
  int8_t e_pl[kPhaseMaxEntries];                      // plane index inside the group's list
  int16_t e_roff[kPhaseMaxEntries];                   // row shift (dh - hmin) * Wb + (dw - wmin)
  int16_t e_w1[kPhaseMaxEntries], e_w2[kPhaseMaxEntries];   // weight slabs: columns [0, ..) / [64, 128) (dual), -1 = none
  uint8_t e_acc[kPhaseMaxEntries];                    // 1: the entry's columns were written by an earlier entry (chunk 0)
};

struct PhaseItem { int n, nt, d, h0, w0, pg; };
// item order: group fastest, then the tiles of a plane, d, the 128-channel slice, n (statistics flush per (n, slice))
__device__ __forceinline__ PhaseItem phase_decode(const PhaseP& P, int item) {
  PhaseItem t;
  t.pg = item % P.ngroups; item /= P.ngroups;
  const int j = item % P.tiles_hw; item /= P.tiles_hw;
  t.d = item % P.Dl; item /= P.Dl;
  t.nt = item % P.n_tiles;
  t.n = item / P.n_tiles;
  t.h0 = (j / P.tiles_w) * 16; t.w0 = (j % P.tiles_w) * 8;
  return t;
}

template <int kMode>
__global__ void __launch_bounds__(kThreadsHalo, 1)
gather_phase_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                    const __grid_constant__ PhaseP P) {
  pdl_launch_dependents();
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  constexpr uint32_t kBStage = 128 * 128;                      // one weight stage: up to 128 rows x 64 channels
  uint8_t* planes = smem;
  uint8_t* bring = smem + (size_t)P.NP * P.slot_bytes;
  uint64_t* p_full = reinterpret_cast<uint64_t*>(bring + (size_t)P.NB * kBStage);
  uint64_t* p_empty = p_full + P.NP;
  uint64_t* b_full = p_empty + P.NP;
  uint64_t* b_empty = b_full + P.NB;
  uint64_t* acc_full = b_empty + P.NB;            // [kPhaseSubs]
  uint64_t* acc_empty = acc_full + kPhaseSubs;    // [kPhaseSubs]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_empty + kPhaseSubs);

  __shared__ EpiRed epi_red;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  if (threadIdx.x == 0) {
    prefetch_tmap(&tmA);
    prefetch_tmap(&tmB);
    for (int s = 0; s < P.NP; ++s) { mbar_init(&p_full[s], 1); mbar_init(&p_empty[s], 1); }
    for (int s = 0; s < P.NB; ++s) { mbar_init(&b_full[s], 1); mbar_init(&b_empty[s], 1); }
    for (int b = 0; b < kPhaseSubs; ++b) { mbar_init(&acc_full[b], 1); mbar_init(&acc_empty[b], kEpiWarps); }
    fence_barrier_init();
  }
  if (warp == kHaloMmaWarp) tmem_alloc(tmem_slot, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_wait();
  const bool prof = (P.debug & 2) != 0;
  // Every CTA walks the channel chunks and the sub-items of a work item in its own ROTATED order (sums are order-free,
  // the accumulate flags follow the iteration index): at any moment the SMs then stream different weight slabs
  // instead of all 148 asking one L2 slice for the same 16 KB.
  const int rot_kc = (int)(blockIdx.x % (unsigned)P.kchunks);
  const int rot_k = (int)((blockIdx.x / (unsigned)P.kchunks) % (unsigned)kPhaseSubs);

  if (warp >= kEpiWarps) {            // third warpgroup: producers, MMA issuer, idle warps -- one setmaxnreg for all of it
  regs_other();
  if (warp == kHaloPlaneWarp) {
    // ---- plane producer: per (work item, channel chunk) the group's halo planes
    if (elect_one()) {
      int s = 0;
      uint32_t ph = 0;
      bool ok = true;
      for (int it = blockIdx.x; it < P.total_items && ok; it += gridDim.x) {
        const PhaseItem t = phase_decode(P, it);
        const int nd = P.g_nd[t.pg];
        for (int kci = 0; kci < P.kchunks && ok; ++kci) {
          int kc = kci + rot_kc;
          if (kc >= P.kchunks) kc -= P.kchunks;
          for (int pl = 0; pl < nd; ++pl) {
            if (!mbar_wait(&p_empty[s], ph ^ 1u, P.err, 41)) { ok = false; break; }
            mbar_expect_tx(&p_full[s], (uint32_t)P.plane_tx);
            tma_load_5d(planes + (size_t)s * P.slot_bytes, &tmA, &p_full[s], kc * 64, t.w0 + P.wmin, t.h0 + P.hmin,
                        t.d + (int)P.g_dd[t.pg][pl], t.n);
            if (++s == P.NP) { s = 0; ph ^= 1u; }
          }
        }
      }
    }
  } else if (warp == kProdWarp) {
    // ---- weight producer: per (work item, channel chunk, sub-item, entry) the entry's slab(s)
    if (elect_one()) {
      int s = 0;
      uint32_t ph = 0;
      bool ok = true;
      long long t_wait = 0, t_begin = prof ? clock64() : 0;
      for (int it = blockIdx.x; it < P.total_items && ok; it += gridDim.x) {
        const PhaseItem t = phase_decode(P, it);
        const int n0 = t.nt * 128;
        for (int kci = 0; kci < P.kchunks && ok; ++kci) {
          int kc = kci + rot_kc;
          if (kc >= P.kchunks) kc -= P.kchunks;
         for (int ki = 0; ki < kPhaseSubs && ok; ++ki) {
          const int sub = t.pg * kPhaseSubs + ((ki + rot_k) & (kPhaseSubs - 1));
          for (int e = P.s_ent0[sub]; e < P.s_ent0[sub + 1]; ++e) {
            const long long tw0 = prof ? clock64() : 0;
            if (!mbar_wait(&b_empty[s], ph ^ 1u, P.err, 42)) { ok = false; break; }
            if (prof) t_wait += clock64() - tw0;
            uint8_t* dst = bring + (size_t)s * kBStage;
            const int w1 = P.e_w1[e], w2 = P.e_w2[e];
            if (P.dual) {
              mbar_expect_tx(&b_full[s], (w1 >= 0 && w2 >= 0) ? kBStage : kBStage / 2);
              tma_load_2d(dst, &tmB, &b_full[s], kc * 64, (w1 >= 0 ? w1 : w2) * P.Cn);
              if (w1 >= 0 && w2 >= 0) tma_load_2d(dst + kBStage / 2, &tmB, &b_full[s], kc * 64, w2 * P.Cn);
            } else {
              mbar_expect_tx(&b_full[s], kBStage);
              tma_load_2d(dst, &tmB, &b_full[s], kc * 64, w1 * P.Cn + n0);
            }
            if (++s == P.NB) { s = 0; ph ^= 1u; }
          }
         }
        }
      }
      if (prof) { atomicAdd(P.dbg + 0, (unsigned long long)t_wait); atomicAdd(P.dbg + 1, (unsigned long long)(clock64() - t_begin)); }
    }
  } else if (warp == kHaloMmaWarp) {
    if (elect_one()) {
      const uint32_t idesc128 = make_idesc_m(128, 128), idesc64 = make_idesc_m(128, 64);
      const uint64_t a_desc0 = desc_kmajor_sw128_sbo(0, (uint32_t)P.sbo);
      const uint64_t b_desc0 = desc_kmajor_sw128(0);
      const uint32_t planes_u = smem_u32(planes) >> 4, bring_u = smem_u32(bring) >> 4;
      const uint32_t slot_u = (uint32_t)P.slot_bytes >> 4, bst_u = kBStage >> 4, row_u = 128u >> 4;
      const int NP = P.NP, NB = P.NB, kchunks = P.kchunks, dual = P.dual;
      int ps = 0, bs = 0;
      uint32_t pph = 0, bph = 0;
      bool ok = true;
      int jt = 0;
      long long t_wait = 0, t_waitp = 0, t_wacc = 0, t_begin = prof ? clock64() : 0;
      for (int it = blockIdx.x; it < P.total_items && ok; it += gridDim.x, ++jt) {
        const uint32_t aph = (uint32_t)jt & 1u;
        const int pg = it % P.ngroups;
        const int nd = P.g_nd[pg];
        for (int kc = 0; kc < kchunks && ok; ++kc) {
          // the chunk's planes sit in nd consecutive ring slots starting at ps
          const long long tp0 = prof ? clock64() : 0;
          {
            int s = ps;
            uint32_t ph = pph;
            for (int pl = 0; pl < nd && ok; ++pl) {
              if (!mbar_wait(&p_full[s], ph, P.err, 45)) ok = false;
              if (++s == NP) { s = 0; ph ^= 1u; }
            }
          }
          if (prof) t_waitp += clock64() - tp0;
          if (!ok) break;
          tc_fence_after();
          for (int ki = 0; ki < kPhaseSubs && ok; ++ki) {
            const int k = (ki + rot_k) & (kPhaseSubs - 1);
            const int sub = pg * kPhaseSubs + k;
            if (kc == 0) {
              const long long ta0 = prof ? clock64() : 0;
              if (!mbar_wait(&acc_empty[k], aph ^ 1u, P.err, 44)) { ok = false; break; }
              if (prof) t_wacc += clock64() - ta0;
              tc_fence_after();
            }
            const uint32_t d_tmem0 = tmem_base + (uint32_t)(k * 128);
            for (int e = P.s_ent0[sub]; e < P.s_ent0[sub + 1]; ++e) {
              const long long tw0 = prof ? clock64() : 0;
              if (!mbar_wait(&b_full[bs], bph, P.err, 46)) { ok = false; break; }
              if (prof) t_wait += clock64() - tw0;
              tc_fence_after();
              int sl = ps + (int)P.e_pl[e];
              if (sl >= NP) sl -= NP;
              const uint32_t a_u = planes_u + (uint32_t)sl * slot_u + (uint32_t)P.e_roff[e] * row_u;
              const uint64_t ad = a_desc0 | (uint64_t)(a_u & 0x3FFFu);
              const uint64_t bd = b_desc0 | (uint64_t)((bring_u + (uint32_t)bs * bst_u) & 0x3FFFu);
              uint32_t idesc = idesc128, d_tmem = d_tmem0;
              if (dual) {
                const bool has1 = P.e_w1[e] >= 0, has2 = P.e_w2[e] >= 0;
                idesc = (has1 && has2) ? idesc128 : idesc64;
                if (!has1) d_tmem += 64u;
              }
              const uint32_t acc = (uint32_t)(kc != 0 || P.e_acc[e] != 0);
              umma_f16(d_tmem, ad, bd, idesc, acc);
              umma_f16(d_tmem, ad + 2, bd + 2, idesc, 1u);
              umma_f16(d_tmem, ad + 4, bd + 4, idesc, 1u);
              umma_f16(d_tmem, ad + 6, bd + 6, idesc, 1u);
              umma_commit(&b_empty[bs]);
              if (++bs == NB) { bs = 0; bph ^= 1u; }
            }
            if (ok && kc == kchunks - 1) umma_commit(&acc_full[k]);
          }
          // the chunk's planes are dead once every MMA issued so far has read them
          for (int pl = 0; pl < nd; ++pl) {
            if (ok) umma_commit(&p_empty[ps]);
            if (++ps == NP) { ps = 0; pph ^= 1u; }
          }
        }
      }
      if (prof) {
        atomicAdd(P.dbg + 2, (unsigned long long)t_wait); atomicAdd(P.dbg + 3, (unsigned long long)(clock64() - t_begin));
        atomicAdd(P.dbg + 6, (unsigned long long)t_wacc); atomicAdd(P.dbg + 7, (unsigned long long)t_waitp);
        atomicAdd(P.dbg + 5, 1ull);
      }
    }
  }
  } else {
    // ---- epilogue warps 0..7 -> TMEM lane quadrant (warp % 4), alternate 32-column chunks of the 128-column sub-item
    regs_epilogue();
    const EpiWarp W(warp);
    const int q = W.q;
    const int row = q * 32 + lane;
    const int nch = P.dual ? 2 : 4;                        // CHANNEL chunks of a sub-item (dual: each of them twice)
    const bool defer = kMode != 0 && nch <= 2;
    EpiStats st;
    st.clear();
    float d1[32], d2[32];
#pragma unroll
    for (int i = 0; i < 32; ++i) { d1[i] = 0.f; d2[i] = 0.f; }
    int st_n = -1, st_n0 = 0;
    const EpiArgs E{P.out, P.out_bf16, P.bias, P.act, P.slope, P.stats != nullptr, P.aux, P.aux_nslope};
    int jt = 0;
    bool ok = true;
    for (int it = blockIdx.x; it < P.total_items && ok; it += gridDim.x, ++jt) {
      const PhaseItem t = phase_decode(P, it);
      const uint32_t aph = (uint32_t)jt & 1u;
      const int lh = t.h0 + (row >> 3), lw = t.w0 + (row & 7);
      const bool valid = lw < P.Wl && lh < P.Hl;
      const int n0 = t.nt * 128;
      if (kMode != 0 && (t.n != st_n || n0 != st_n0)) {
        epilogue_flush_stats(P.stats, st_n, P.Cn, st_n0, W.q, W.c_begin, 2, nch, lane, st, defer, d1, d2, epi_red);
        st_n = t.n; st_n0 = n0;
      }
      for (int k = 0; k < kPhaseSubs && ok; ++k) {
        const int sub = t.pg * kPhaseSubs + k;
        const long long obase = (long long)t.n * P.osn + (long long)(t.d * P.ostep + P.s_od[sub]) * P.osd +
                                (long long)(lh * P.ostep + P.s_oh[sub]) * P.osh +
                                (long long)(lw * P.ostep + P.s_ow[sub]) * P.osw + n0;
        AuxRegs ax;
        epilogue_aux_first<kMode>(E, W.c_begin, 4, valid, obase, P.dual != 0, ax);
        ok = mbar_wait(&acc_full[k], aph, P.err, 43);
        if (!ok) break;
        tc_fence_after();
        const long long te0 = prof ? clock64() : 0;
        const uint32_t t_addr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(k * 128);
        uint64_t* rel_bar = &acc_empty[k];
        epilogue_tile<kMode>(E, t_addr, W.c_begin, 2, 4, valid, obase, n0, lane, st, 0, defer, d1, d2, ax, [&]() {
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(rel_bar);
        }, P.dual != 0, P.osw, valid);
        if (prof && threadIdx.x == 0) atomicAdd(P.dbg + 4, (unsigned long long)(clock64() - te0));
      }
    }
    if (kMode != 0) epilogue_flush_stats(P.stats, st_n, P.Cn, st_n0, W.q, W.c_begin, 2, nch, lane, st, defer, d1, d2, epi_red);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == kHaloMmaWarp) tmem_dealloc(tmem_base, 512);
}

// ------------------------------------------------------------------ host side
// Eligibility + tables: 8 parity phases over one launch space (gather_mergeable), 64 or a multiple of 128 output
// channels, the tap tables within their limits, planes + a weight ring within shared memory.
inline bool phase_setup(const GatherPlan& plan, const GatherRun& R, PhaseP& P) {
  // Opt-in (MRA_GATHER_PHASE=1): measured on B200 it loses to the per-tap merged kernel on every BASELINE layer
  // (profiles/r02_conv_layers_phase_vs_pertap.txt: G.u2 fprop 0.475 vs 0.369 ms, G.d1 dgrad 0.457 vs 0.333 ms) -- the
  // weight slabs, not the A tiles, dominate the L2 -> SM bytes of these layers and a single-CTA version streams
  // them just as often; it is kept as the base for a cta_group::2 version that shares the slabs.
  const char* on = getenv("MRA_GATHER_PHASE");
  if (!gather_mergeable(plan) || !on || atoi(on) == 0) return false;
  if (plan.ck % 64 != 0) return false;
  const bool dual = plan.cn == 64;
  if (!dual && plan.cn % 128 != 0) return false;
  const GatherLaunch* Ls = plan.launches.data();
  int lo[3] = {1 << 30, 1 << 30, 1 << 30}, hi[3] = {-(1 << 30), -(1 << 30), -(1 << 30)};
  for (int l = 0; l < 8; ++l) {
    if (Ls[l].ostep != 2) return false;
    for (const Tap& t : Ls[l].taps) {
      if (t.widx >= R.slabs) return false;
      const int o[3] = {t.dd, t.dh, t.dw};
      for (int i = 0; i < 3; ++i) { if (o[i] < lo[i]) lo[i] = o[i]; if (o[i] > hi[i]) hi[i] = o[i]; }
    }
  }
  memset(&P, 0, sizeof(P));
  P.Dl = Ls[0].dims[0]; P.Hl = Ls[0].dims[1]; P.Wl = Ls[0].dims[2]; P.N = plan.n;
  P.tiles_w = (P.Wl + 7) / 8; P.tiles_hw = ((P.Hl + 15) / 16) * P.tiles_w;
  P.hmin = lo[1]; P.wmin = lo[2];
  P.Hb = 16 + hi[1] - lo[1]; P.Wb = 8 + hi[2] - lo[2];
  if (P.Hb > 256 || P.Wb > 256) return false;
  P.sbo = P.Wb * 128;
  P.plane_tx = P.Wb * P.Hb * 128;
  P.slot_bytes = (P.plane_tx + 1023) / 1024 * 1024;
  P.kchunks = plan.ck / 64;
  P.Cn = plan.cn; P.dual = dual ? 1 : 0; P.n_tiles = dual ? 1 : plan.cn / 128;
  P.ngroups = dual ? 1 : 2;
  P.ostep = 2;
  // ---- sub-items and entries
  struct Ent { int dd, dh, dw, w1, w2; };
  std::vector<std::vector<Ent>> subs;                 // kPhaseSubs per group
  int sub_o0[kPhaseMaxGroups * kPhaseSubs][3];
  if (dual) {
    std::vector<std::vector<PairEntry>> pairs;
    int pair_o0[4][3];
    if (!build_phase_pairs(Ls, R.slabs, pairs, pair_o0)) return false;
    for (int p = 0; p < 4; ++p) {
      std::vector<Ent> v;
      for (const PairEntry& e : pairs[p]) v.push_back(Ent{e.dd, e.dh, e.dw, e.w1, e.w2});
      subs.push_back(v);
      for (int i = 0; i < 3; ++i) sub_o0[p][i] = pair_o0[p][i];
    }
  } else {
    for (int pd = 0; pd < 2; ++pd) {
      int cnt = 0;
      for (int l = 0; l < 8; ++l) {
        if (Ls[l].o0[0] != pd) continue;
        std::vector<Ent> v;
        for (const Tap& t : Ls[l].taps) v.push_back(Ent{t.dd, t.dh, t.dw, t.widx, -1});
        for (int i = 0; i < 3; ++i) sub_o0[(int)subs.size()][i] = Ls[l].o0[i];
        subs.push_back(v);
        ++cnt;
      }
      if (cnt != kPhaseSubs) return false;
    }
  }
  if ((int)subs.size() != P.ngroups * kPhaseSubs) return false;
  int ne = 0, max_nd = 0;
  for (int g = 0; g < P.ngroups; ++g) {
    // the group's planes: distinct d-offsets of its taps, ascending
    std::vector<int> dds;
    for (int k = 0; k < kPhaseSubs; ++k)
      for (const Ent& e : subs[g * kPhaseSubs + k]) {
        bool seen = false;
        for (int d : dds) if (d == e.dd) seen = true;
        if (!seen) dds.push_back(e.dd);
      }
    for (size_t a = 0; a < dds.size(); ++a)
      for (size_t b = a + 1; b < dds.size(); ++b) if (dds[b] < dds[a]) { int t = dds[a]; dds[a] = dds[b]; dds[b] = t; }
    if (dds.size() > 4) return false;
    P.g_nd[g] = (int8_t)dds.size();
    for (size_t a = 0; a < dds.size(); ++a) P.g_dd[g][a] = (int8_t)dds[a];
    if ((int)dds.size() > max_nd) max_nd = (int)dds.size();
    for (int k = 0; k < kPhaseSubs; ++k) {
      const int s = g * kPhaseSubs + k;
      P.s_ent0[s] = (int16_t)ne;
      P.s_od[s] = (int8_t)sub_o0[s][0]; P.s_oh[s] = (int8_t)sub_o0[s][1]; P.s_ow[s] = (int8_t)sub_o0[s][2];
      bool init1 = false, init2 = false;
      for (const Ent& e : subs[s]) {
        if (ne >= kPhaseMaxEntries) return false;
        int pl = 0;
        while (dds[pl] != e.dd) ++pl;
        P.e_pl[ne] = (int8_t)pl;
        P.e_roff[ne] = (int16_t)((e.dh - lo[1]) * P.Wb + (e.dw - lo[2]));
        P.e_w1[ne] = (int16_t)e.w1; P.e_w2[ne] = (int16_t)e.w2;
        if (e.w1 >= 0 && e.w2 >= 0) { if (init1 != init2) return false; P.e_acc[ne] = init1 ? 1 : 0; init1 = init2 = true; }
        else if (e.w1 >= 0) { P.e_acc[ne] = init1 ? 1 : 0; init1 = true; }
        else { P.e_acc[ne] = init2 ? 1 : 0; init2 = true; }
        ++ne;
      }
      if (subs[s].empty() || !init1 || (dual && !init2)) return false;     // every column must be written
    }
  }
  P.s_ent0[P.ngroups * kPhaseSubs] = (int16_t)ne;
  // ---- shared memory: a plane ring of up to 3 chunks' worth, the rest for the weight ring (>= 3 stages)
  const size_t budget = kSmemLimit - 2048 - kEpiRedBytes;
  int NP = 3 * max_nd;
  { const char* e = getenv("MRA_PHASE_NP"); if (e && atoi(e) >= max_nd && atoi(e) <= 4 * max_nd) NP = atoi(e); }
  while (NP > max_nd + 1 && (size_t)NP * P.slot_bytes + 3 * (size_t)(128 * 128) > budget) --NP;
  if ((size_t)NP * P.slot_bytes + 3 * (size_t)(128 * 128) > budget) return false;
  int NB = (int)((budget - (size_t)NP * P.slot_bytes) / (128 * 128));
  if (NB > 10) NB = 10;
  P.NP = NP; P.NB = NB;
  const long long total = (long long)plan.n * P.n_tiles * P.Dl * P.tiles_hw * P.ngroups;
  if (total >= (1ll << 31)) return false;
  P.total_items = (int)total;
  return true;
}

inline int run_gather_phase(const GatherPlan& plan, PhaseP& P, const GatherRun& R, cudaStream_t st) {
  static bool attr_set = false;
  if (!attr_set) {
    MRA_CHECK_CUDA(cudaFuncSetAttribute(gather_phase_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(kSmemLimit - kEpiRedBytes)));
    MRA_CHECK_CUDA(cudaFuncSetAttribute(gather_phase_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(kSmemLimit - kEpiRedBytes)));
    MRA_CHECK_CUDA(cudaFuncSetAttribute(gather_phase_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(kSmemLimit - kEpiRedBytes)));
    attr_set = true;
  }
  P.osw = plan.cn;
  P.osh = (long long)plan.odims[2] * P.osw;
  P.osd = (long long)plan.odims[1] * P.osh;
  P.osn = (long long)plan.odims[0] * P.osd;
  P.out = R.out; P.out_bf16 = R.out_bf16;
  P.bias = R.bias; P.act = R.act; P.slope = R.slope;
  P.stats = R.stats; P.aux = R.aux; P.aux_nslope = R.aux_nslope; P.err = tc_err_flag();
  { const char* e = getenv("MRA_GATHER_DEBUG"); P.debug = e ? atoi(e) : 0; }
  P.dbg = tc_dbg_counters();
  CUtensorMap tmA, tmB;
  if (int rc = make_act_map(&tmA, R.a, plan.n, plan.adims[0], plan.adims[1], plan.adims[2], plan.ck, P.Wb, P.Hb, 1, 1)) return rc;
  if (int rc = make_weight_map(&tmB, R.b, (long long)R.slabs * plan.cn, plan.ck, P.dual ? 64 : 128)) return rc;
  const size_t smem = (size_t)P.NP * P.slot_bytes + (size_t)P.NB * 128 * 128 + 1024 + 512;
  const int ctas = P.total_items < num_sms() ? P.total_items : num_sms();
  if (P.aux && P.stats) MRA_CHECK_CUDA(launch_pdl(gather_phase_kernel<2>, dim3(ctas), dim3(kThreadsHalo), smem, st, 1, tmA, tmB, P));
  else if (P.stats) MRA_CHECK_CUDA(launch_pdl(gather_phase_kernel<1>, dim3(ctas), dim3(kThreadsHalo), smem, st, 1, tmA, tmB, P));
  else MRA_CHECK_CUDA(launch_pdl(gather_phase_kernel<0>, dim3(ctas), dim3(kThreadsHalo), smem, st, 1, tmA, tmB, P));
  MRA_LAUNCH_CHECK();
  return 0;
}

}  // namespace tc
}  // namespace mra

namespace mra {
namespace tc {
// dispatch hook used by run_gather_tc (conv_tc_halo.cuh): true = the plan was run here (rc holds the result)
inline bool phase_try_run(const GatherPlan& plan, const GatherRun& R, cudaStream_t st, int* rc) {
  PhaseP P;
  if (!phase_setup(plan, R, P)) return false;
  *rc = run_gather_phase(plan, P, R, st);
  return true;
}
}  // namespace tc
}  // namespace mra
