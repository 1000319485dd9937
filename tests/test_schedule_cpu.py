"""CPU walk of the PERSISTENT SCHEDULES of the tensor-core kernels (mra_debug_schedule): the same halo_decode /
halo_stats_key / halo_plane_live / wseg_begin / wseg_next the kernels execute, compiled for the host.

What round 1's out-of-bounds statistics flush and this round's stream-K range-start rewrite have in common is that they are
pure index arithmetic which no value-level parity test can see (an atomicAdd of +0.0 past a buffer, a CTA starting one
work item late only for some launch shapes).  These tests check the arithmetic itself, for the BASELINE launches and for
grids / batch sizes no box of this pool has: every output element of a gather launch is produced exactly once, every
statistics flush stays inside [0, Cn) of its own sample, half-width tail items add into the partials of their tile, every
work item keeps a live input plane when dead planes are skipped, and the stream-K segments of all CTAs tile every work
item's K range exactly once, starting where a linear walk over the work items would start."""
import itertools

import pytest

from mra_gan_b200 import ops
from mra_gan_b200.ops import ConvGeom


def parse_gather(words):
    assert words[0] == 1
    nl, p, out = words[1], 2, []
    for _ in range(nl):
        keys = ("li pair mode N Dl Hl Wl Wb Cn n_tile n_tiles total_tiles split_from total_work units skip kd nsub "
                "nrec").split()
        h = dict(zip(keys, words[p:p + len(keys)]))
        p += len(keys)
        recs = [words[p + 15 * i: p + 15 * (i + 1)] for i in range(h["nrec"])]
        p += 15 * h["nrec"]
        h["recs"] = recs
        out.append(h)
    assert p == len(words)
    return out


HALO_CASES = [
    # (geom, in_dims, n, which, units)                                   units 0 = the launch's own grid (74 pairs / 148 CTAs)
    (ConvGeom(256, 256, 3, 1, 0), (34, 34, 34), 2, 0, 0),      # G.rb fprop at the bench's batch: 256 pair tiles, rem 34 -> split
    (ConvGeom(256, 256, 3, 1, 0), (34, 34, 34), 1, 0, 0),
    (ConvGeom(256, 256, 3, 1, 0), (34, 34, 34), 4, 0, 0),
    (ConvGeom(256, 256, 3, 1, 0), (34, 34, 34), 27, 0, 0),
    (ConvGeom(256, 256, 3, 1, 0), (34, 34, 34), 2, 1, 0),      # G.rb dgrad: flat tiles, 34^3 outputs, dead planes
    (ConvGeom(256, 256, 3, 1, 0), (34, 34, 34), 3, 1, 0),
    (ConvGeom(256, 256, 3, 1, 0), (34, 34, 34), 2, 0, 66),     # a 132-SM part
    (ConvGeom(256, 256, 3, 1, 0), (34, 34, 34), 2, 1, 7),
    (ConvGeom(256, 256, 3, 1, 0), (11, 10, 9), 1, 1, 0),       # odd plane count: the last pair's partner plane does not exist
    (ConvGeom(256, 512, 4, 1, 1), (16, 16, 16), 2, 0, 0),      # D.4 fprop: n_tile 256 x 2 tiles, zero padding
    (ConvGeom(256, 512, 4, 1, 1), (16, 16, 16), 2, 1, 0),      # D.4 dgrad
    (ConvGeom(256, 512, 4, 1, 1), (16, 16, 16), 5, 0, 0),
    (ConvGeom(64, 64, 3, 1, 1), (9, 7, 5), 1, 0, 0),
    (ConvGeom(128, 128, 3, 1, 1), (20, 20, 20), 3, 0, 0),
]


def _ids(c):
    g, dims, n, which, units = c
    return "c%d-%d_k%d_p%d_%s_n%d_%s_u%d" % (g.cin, g.cout, g.k, g.pad, "x".join(map(str, dims)), n, "fdw"[which], units)


@pytest.mark.parametrize("form", ["pair", "single", "dual"])
@pytest.mark.parametrize("case", HALO_CASES, ids=_ids)
def test_gather_halo_schedule_covers_every_output_once_and_flushes_inside_the_buffer(case, form, monkeypatch):
    g, dims, n, which, units = case
    single = form == "single"
    if form == "dual":
        monkeypatch.setenv("MRA_HALO_DUAL", "1")      # two position tiles x 128 channels per work item (conv_tc_halo.cuh)
    else:
        monkeypatch.delenv("MRA_HALO_DUAL", raising=False)
    launches = parse_gather(ops.schedule_describe(g, n, dims, which, units=units, single=single))
    assert launches, "case is expected to run on gather_halo_kernel"
    for L in launches:
        pair = bool(L["pair"])
        assert pair == (not single)
        assert L["nsub"] == (2 if form == "dual" and g.cout == 256 and g.cin == 256 and dims == (34, 34, 34) else L["nsub"])
        assert L["nsub"] == 1 or (form == "dual" and L["n_tile"] == 128)
        Cn, n_tile = L["Cn"], L["n_tile"]
        assert Cn % n_tile == 0 and L["n_tiles"] == Cn // n_tile
        cover = {}
        per_cta = {}
        for (unit, rank, work, nn, d, n0, width, h0, w0, f0, tb, coff, live, sub, rot) in L["recs"]:
            assert 0 <= work < L["total_work"] and 0 <= nn < L["N"] and 0 <= sub < L["nsub"]
            assert width in (n_tile, n_tile // 2) and (width == n_tile) == (work < L["split_from"])
            # channel range of the item and of the statistics flush it belongs to
            assert 0 <= n0 and n0 + width <= Cn
            assert tb % n_tile == 0 and tb <= n0 and n0 + width <= tb + n_tile <= Cn
            assert coff == (n0 - tb) // 32 and coff % 2 == 0 and coff + width // 32 <= n_tile // 32
            assert live != 0, "a work item must keep at least one live input plane"
            if not L["skip"]:
                assert live == (1 << L["kd"]) - 1
            if sub == 0:
                per_cta.setdefault((unit, rank), []).append((work, nn, tb))
                first = (nn, d, n0, width, tb, coff, live, rot)
            else:                                   # dual items: same sample, plane, channels, planes and weight order
                assert first == (nn, d, n0, width, tb, coff, live, rot)
            if d >= L["Dl"]:
                assert pair and rank == 1 and d == L["Dl"] and L["Dl"] % 2 == 1     # the odd plane's partner: nothing stored
                continue
            if L["mode"] == 0:
                pos = [(h0 + r // 8, w0 + r % 8) for r in range(128)]
            else:
                pos = [divmod(f0 + r, L["Wb"]) for r in range(128)]
            for (h, w) in pos:
                if h < L["Hl"] and w < L["Wl"]:
                    for c32 in range(n0 // 32, (n0 + width) // 32):
                        key = (nn, d, h, w, c32)
                        cover[key] = cover.get(key, 0) + 1
        want = L["N"] * L["Dl"] * L["Hl"] * L["Wl"] * (Cn // 32)
        assert len(cover) == want and set(cover.values()) == {1}, "every output element exactly once"
        # a CTA walks its items in increasing work order; the (sample, tile) key of its statistics partials changes
        # monotonically, so each key is flushed once per CTA and a half-width item never opens a key of its own
        for (unit, rank), items in per_cta.items():
            works = [w for (w, _, _) in items]
            assert works == sorted(works) and all(b - a == L["units"] for a, b in zip(works, works[1:]))
            keys = [(nn, tb) for (_, nn, tb) in items]
            if L["units"] % L["n_tiles"] == 0:      # (a grid that is no multiple of n_tiles alternates channel tiles: legal -- a
                seen = []                           # flush ADDS -- but one flush per item; not the case on 148 SMs)
                for k in keys:
                    if not seen or seen[-1] != k:
                        assert k not in seen, "a statistics key should not come back after it was flushed"
                        seen.append(k)
        # both CTAs of a pair see the same items (the same weight slabs), planes 2q and 2q + 1
        if pair:
            by_work = {}
            for (unit, rank, work, nn, d, n0, width, h0, w0, f0, tb, coff, live, sub, rot) in L["recs"]:
                by_work.setdefault((work, sub), {})[rank] = (unit, nn, d, n0, width, h0, w0, f0, live, rot)
            for work, rr in by_work.items():
                a, b = rr[0], rr[1]
                assert a[0] == b[0] and a[1] == b[1] and b[2] == a[2] + 1 and a[2] % 2 == 0 and a[3:] == b[3:]


def test_bench_launch_splits_its_tail_and_dgrad_skips_two_plane_steps():
    """The two launches the round-1 review was about, by the numbers: G.rb fprop at batch 2 has 256 pair tiles on 74 pairs
    (rem 34 -> 68 half-width items), and the dead-plane skip drops exactly 2 of the 51 (plane pair, td) steps of its dgrad."""
    (L,) = parse_gather(ops.schedule_describe(ConvGeom(256, 256, 3, 1, 0), 2, (34, 34, 34), 0))
    assert (L["total_tiles"], L["split_from"], L["total_work"], L["units"]) == (256, 222, 290, 74) and L["skip"] == 0
    (L,) = parse_gather(ops.schedule_describe(ConvGeom(256, 256, 3, 1, 0), 2, (34, 34, 34), 1))
    assert L["mode"] == 1 and L["Dl"] == 34 and L["skip"] == 1 and L["total_tiles"] == 340
    steps = {}
    for r in L["recs"]:
        if r[1] == 0 and r[3] == 0:
            steps[r[4]] = r[12]                     # first plane of the pair -> live mask
    assert len(steps) == 17 and sum(bin(m).count("1") for m in steps.values()) == 51 - 2
    assert steps[0] == 0b110 and steps[32] == 0b011 and all(m == 0b111 for d, m in steps.items() if 0 < d < 32)


def parse_wgrad(words):
    assert words[0] == 2
    keys = "pair m_tiles n_tiles n_groups n_items kblocks units cost_lo cost_hi".split()
    h = dict(zip(keys, words[1:1 + len(keys)]))
    h["total_cost"] = h["cost_lo"] + (h["cost_hi"] << 31)
    p = 1 + len(keys)
    h["groups"] = [dict(zip("tap0 ntaps gpi item0 n_items".split(), words[p + 5 * i: p + 5 * i + 5])) for i in range(h["n_groups"])]
    p += 5 * h["n_groups"]
    nseg = words[p]
    p += 1
    h["segs"] = [words[p + 9 * i: p + 9 * i + 9] for i in range(nseg)]
    assert p + 9 * nseg == len(words)
    return h


WGRAD_CASES = [
    (ConvGeom(256, 256, 3, 1, 0), (34, 34, 34), 2, 0),         # G.rb: CTA pairs, 14 items
    (ConvGeom(256, 256, 3, 1, 0), (34, 34, 34), 4, 0),
    (ConvGeom(64, 128, 3, 2, 1), (128, 128, 128), 2, 0),       # G.d1: 4 tap groups (8, 8, 8, 3)
    (ConvGeom(128, 64, 3, 2, 1, True, 1), (64, 64, 64), 2, 0),   # G.u2 (ConvTranspose3d: dense = x)
    (ConvGeom(256, 512, 4, 1, 1), (16, 16, 16), 2, 0),         # D.4
    (ConvGeom(512, 512, 4, 2, 1), (2, 2, 2), 4, 0),            # UNet-7 d7: 2048 items of a handful of K-blocks
    (ConvGeom(512, 512, 4, 2, 1), (8, 8, 8), 4, 0),
    (ConvGeom(1024, 512, 4, 2, 1, True, 0), (2, 2, 2), 4, 0),
    (ConvGeom(1024, 256, 4, 2, 1, True, 0), (16, 16, 16), 1, 0),
    (ConvGeom(64, 64, 4, 1, 1), (22, 22, 22), 1, 0),           # 64 taps, one 64-channel chunk each side
    (ConvGeom(256, 256, 3, 1, 0), (34, 34, 34), 2, 5),         # few units: every range crosses many items
    (ConvGeom(512, 512, 4, 2, 1), (4, 4, 4), 3, 131),          # a prime grid
    (ConvGeom(128, 256, 3, 2, 1), (15, 16, 17), 1, 0),
]


@pytest.mark.parametrize("case", WGRAD_CASES, ids=lambda c: "c%d-%d_k%d_s%d_%s_%s_n%d_u%d" % (
    c[0].cin, c[0].cout, c[0].k, c[0].stride, "T" if c[0].transposed else "C", "x".join(map(str, c[1])), c[2], c[3]))
def test_wgrad_stream_k_segments_tile_every_work_item_once(case):
    g, dims, n, units = case
    S = parse_wgrad(ops.schedule_describe(g, n, dims, 2, units=units))
    per_item = S["m_tiles"] * S["n_tiles"]
    kblocks, nunits = S["kblocks"], S["units"]
    # the work list as a linear walk sees it: (item, ntap) in order, every item per_item times
    works = []
    for G in S["groups"]:
        assert G["n_items"] == -(-G["ntaps"] // G["gpi"])
        for it in range(G["n_items"]):
            nt = min(G["gpi"], G["ntaps"] - it * G["gpi"])
            works += [(G["item0"] + it, nt, G["tap0"] + it * G["gpi"])] * per_item
    assert sum(nt for (_, nt, _) in works) * kblocks == S["total_cost"]
    share = -(-S["total_cost"] // nunits)
    # reference partition: cut the concatenated cost axis into `nunits` ranges of `share`, walk the work list linearly
    want = []
    off = 0
    for wi, (item, nt, tap0) in enumerate(works):
        span = nt * kblocks
        u_lo, u_hi = off // share, (off + span - 1) // share
        for u in range(u_lo, u_hi + 1):
            lo, hi = max(off, u * share), min(off + span, (u + 1) * share)
            kb0 = (lo - off) // nt
            kb1 = kblocks if hi == off + span else (hi - off) // nt
            if kb1 > kb0:
                rem = wi % per_item
                want.append([u, item, rem // S["n_tiles"], rem % S["n_tiles"], tap0, nt, kb0, kb1])
        off += span
    got = [[s[0], s[1], s[2], s[3], s[5], s[6], s[7], s[8]] for s in S["segs"]]
    assert got == sorted(want), "stream-K segments differ from the linear walk"
    # and, independently of the reference above: the K range of every (item, m tile, n tile) is tiled exactly once
    cover = {}
    for (u, item, mt, nt_, g_, tap0, ntap, kb0, kb1) in S["segs"]:
        assert 0 <= kb0 < kb1 <= kblocks and 0 <= u < nunits
        cover.setdefault((item, mt, nt_), []).append((kb0, kb1))
    assert len(cover) == S["n_items"] * per_item
    for key, rs in cover.items():
        rs.sort()
        assert rs[0][0] == 0 and rs[-1][1] == kblocks and all(a[1] == b[0] for a, b in zip(rs, rs[1:])), key
