"""CPU check of the implicit-GEMM lowering (csrc/conv_plan.h) through mra_conv_plan_describe():
the serialised plan is executed by a small torch emulator of the two generic device computations
(GATHER / WGRAD) and compared with torch's conv3d / conv_transpose3d and their autograd."""
import itertools

import pytest
import torch
import torch.nn.functional as F

from mra_gan_b200 import ops
from mra_gan_b200.ops import ConvGeom
from oracle.ops_ref import pack_weight

GEOMS = [
    ConvGeom(3, 5, 3, 1, 0), ConvGeom(4, 6, 3, 2, 1), ConvGeom(2, 3, 4, 2, 1), ConvGeom(3, 2, 4, 1, 1),
    ConvGeom(1, 4, 7, 1, 0), ConvGeom(4, 3, 3, 2, 1, True, 1), ConvGeom(3, 2, 4, 2, 1, True, 0),
]


def parse_gather(words):
    it = iter(words)
    nx = lambda k: [next(it) for _ in range(k)]
    kind, n, ck, cn = nx(4)
    assert kind == 0
    adims, odims = nx(3), nx(3)
    (nl,) = nx(1)
    launches = []
    for _ in range(nl):
        o0 = nx(3); (ostep,) = nx(1); dims = nx(3); (astep,) = nx(1); box = nx(3); (nt,) = nx(1)
        taps = [nx(4) for _ in range(nt)]
        launches.append(dict(o0=o0, ostep=ostep, dims=dims, astep=astep, box=box, taps=taps))
    return dict(n=n, ck=ck, cn=cn, adims=adims, odims=odims, launches=launches)


def run_gather(plan, a, b):
    """a: (N, D, H, W, Ck); b: [taps][Cn][Ck] -> out (N, Do, Ho, Wo, Cn)"""
    N = a.shape[0]
    out = torch.zeros((N,) + tuple(plan["odims"]) + (plan["cn"],), dtype=a.dtype)
    written = torch.zeros(tuple(plan["odims"]), dtype=torch.int32)
    AD = plan["adims"]
    for L in plan["launches"]:
        assert L["box"][0] * L["box"][1] * L["box"][2] == 128
        for l in itertools.product(*[range(x) for x in L["dims"]]):
            o = [l[i] * L["ostep"] + L["o0"][i] for i in range(3)]
            acc = torch.zeros(N, plan["cn"], dtype=a.dtype)
            for (dd, dh, dw, wi) in L["taps"]:
                p = [l[0] * L["astep"] + dd, l[1] * L["astep"] + dh, l[2] * L["astep"] + dw]
                if all(0 <= p[i] < AD[i] for i in range(3)):
                    acc += a[:, p[0], p[1], p[2], :] @ b[wi].T
            out[:, o[0], o[1], o[2], :] = acc
            written[o[0], o[1], o[2]] += 1
    assert int(written.min()) == 1 and int(written.max()) == 1, "every output written exactly once"
    return out


@pytest.mark.parametrize("g", GEOMS, ids=lambda g: "c%d-%d_k%d_s%d_p%d_%s" % (g.cin, g.cout, g.k, g.stride, g.pad, "T" if g.transposed else "C"))
def test_plan_matches_torch(g):
    torch.manual_seed(0)
    in_dims = (5, 6, 7) if g.k < 7 else (8, 7, 9)
    N = 2
    x = torch.randn((N,) + in_dims + (g.cin,), dtype=torch.float64)
    wshape = (g.cin, g.cout) if g.transposed else (g.cout, g.cin)
    w_ref = torch.randn(wshape + (g.k,) * 3, dtype=torch.float64)
    wp = pack_weight(w_ref, g.transposed)
    xr = x.permute(0, 4, 1, 2, 3).clone().requires_grad_(True)
    wr = w_ref.clone().requires_grad_(True)
    if g.transposed:
        y = F.conv_transpose3d(xr, wr, None, stride=g.stride, padding=g.pad, output_padding=g.output_padding)
    else:
        y = F.conv3d(xr, wr, None, stride=g.stride, padding=g.pad)
    assert tuple(y.shape[2:]) == g.out_dims(in_dims)
    dy = torch.randn_like(y)
    y.backward(dy)
    # fprop
    plan = parse_gather(ops.plan_describe(g, N, in_dims, 0))
    out = run_gather(plan, x, wp)
    assert torch.allclose(out, y.detach().permute(0, 2, 3, 4, 1), atol=1e-10)
    # dgrad
    plan = parse_gather(ops.plan_describe(g, N, in_dims, 1))
    dyl = dy.permute(0, 2, 3, 4, 1).contiguous()
    dx = run_gather(plan, dyl, wp.transpose(1, 2).contiguous())
    assert torch.allclose(dx, xr.grad.permute(0, 2, 3, 4, 1), atol=1e-10)
    # wgrad
    words = ops.plan_describe(g, N, in_dims, 2)
    kind, n, cm, cn = words[:4]
    assert kind == 1 and (cm, cn) == (g.cout, g.cin)
    qd, md, nd = words[4:7], words[7:10], words[10:13]
    sstep, m_shift = words[13:15]
    box = words[15:18]
    assert box[0] * box[1] * box[2] == 64
    nt = words[18]
    taps = [words[19 + 4 * i: 23 + 4 * i] for i in range(nt)]
    dw = torch.zeros(g.taps, cm, cn, dtype=torch.float64)
    for q in itertools.product(*[range(v) for v in qd]):
        for (dd, dh, dw_, wi) in taps:
            s = [q[0] * sstep + dd, q[1] * sstep + dh, q[2] * sstep + dw_]
            mp, np_ = (s, q) if m_shift else (q, s)
            if all(0 <= mp[i] < md[i] for i in range(3)) and all(0 <= np_[i] < nd[i] for i in range(3)):
                dw[wi] += dyl[:, mp[0], mp[1], mp[2], :].T @ x[:, np_[0], np_[1], np_[2], :]
    assert torch.allclose(dw, pack_weight(wr.grad, g.transposed), atol=1e-9)


def test_box_selection_baseline_shapes():
    # G.rb: 34^3 -> 32^3 : box (d,h,w) = (1,4,32), no padding waste
    w = ops.plan_describe(ConvGeom(256, 256, 3, 1, 0), 1, (34, 34, 34), 0)
    p = parse_gather(w)
    assert p["launches"][0]["box"] == [1, 4, 32] and len(p["launches"][0]["taps"]) == 27
    # ConvTranspose3d k3 s2 p1 op1: 8 parity phases with 1,2,2,4,2,4,4,8 taps (27 in total)
    p = parse_gather(ops.plan_describe(ConvGeom(256, 128, 3, 2, 1, True, 1), 1, (32, 32, 32), 0))
    assert sorted(len(L["taps"]) for L in p["launches"]) == [1, 2, 2, 2, 4, 4, 4, 8]
    # k4 s2 p1 transposed: 8 phases x 8 taps
    p = parse_gather(ops.plan_describe(ConvGeom(512, 256, 4, 2, 1, True, 0), 1, (8, 8, 8), 0))
    assert [len(L["taps"]) for L in p["launches"]] == [8] * 8


def parse_wgrad_launches(words):
    nt = words[18]
    taps = [words[19 + 4 * i: 23 + 4 * i] for i in range(nt)]
    p = 19 + 4 * nt
    ncc, nl = words[p], words[p + 1]
    p += 2
    launches = []
    for _ in range(nl):
        share, gpi = words[p], words[p + 1]
        ext = words[p + 2:p + 5]
        n = words[p + 5]
        p += 6
        ent = [words[p + 4 * i: p + 4 * i + 4] for i in range(n)]
        p += 4 * n
        launches.append(dict(share=share, gpi=gpi, ext=ext, entries=ent))
    assert p == len(words)
    return taps, ncc, launches, words[15:18]


@pytest.mark.parametrize("g,dims", [(ConvGeom(256, 256, 3, 1, 0), (34, 34, 34)), (ConvGeom(256, 512, 4, 1, 1), (16, 16, 16)),
                                    (ConvGeom(64, 128, 3, 2, 1), (64, 64, 64)), (ConvGeom(128, 64, 3, 2, 1, True, 1), (16, 16, 16)),
                                    (ConvGeom(64, 128, 4, 2, 1), (32, 32, 32)), (ConvGeom(128, 128, 4, 1, 1), (33, 33, 33))],
                         ids=["rb", "d4", "g_d1", "g_u2", "d2", "k4s1_w32"])
def test_wgrad_launch_structure(g, dims):
    """Every tap appears exactly once; taps that share a shared-memory box sit inside the extended box;
    the accumulators of one item fit the 512 TMEM columns."""
    taps, ncc, launches, box = parse_wgrad_launches(ops.plan_describe(g, 1, dims, 2))
    seen = []
    for L in launches:
        assert L["share"] in (1, 2) and L["gpi"] % L["share"] == 0 and L["gpi"] * ncc * 64 <= 512
        if L["share"] == 2:
            assert g.stride == 1 and box[2] % 16 == 0
        for i, (ti, odd, odh, odw) in enumerate(L["entries"]):
            seen.append(ti)
            dd, dh, dw, _ = taps[ti]
            assert 0 <= dd - odd <= L["ext"][0] and 0 <= dh - odh <= L["ext"][1] and 0 <= dw - odw <= L["ext"][2]
            if L["share"] == 2 and i % 2 == 1 and (i // L["gpi"]) == ((i - 1) // L["gpi"]):
                assert L["entries"][i - 1][1:] == [odd, odh, odw]
    assert sorted(seen) == list(range(len(taps)))
    if g == ConvGeom(256, 256, 3, 1, 0):
        assert [(L["share"], len(L["entries"])) for L in launches] == [(2, 18), (2, 6), (1, 3)]
