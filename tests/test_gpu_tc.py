"""GPU parity of the tcgen05/TMA tile engine (csrc/tc_gemm.cu) against plain fp32 PyTorch ops on the
same fp16-rounded operands. Tolerances: fp16 outputs 2e-3 relative to the row scale, fp32 outputs 1e-3."""
import ctypes as C

import pytest
import torch
import torch.nn.functional as F

from cellvit_b200 import _lib as L

pytestmark = pytest.mark.gpu


@pytest.fixture(autouse=True)
def _no_tf32():
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False
    yield


def _gemm(A, W, epi, block_n=0):
    M, K = A.shape
    N = W.shape[0]
    L.check(L.lib().cvb_op_gemm_f16(L.ptr(A), M, K, C.c_longlong(A.stride(0)), L.ptr(W), N, C.c_longlong(W.stride(0)),
                                    block_n, C.byref(epi), L.stream_ptr()), "gemm")
    torch.cuda.synchronize()


@pytest.mark.parametrize("M,K,N,bn,act", [(128, 64, 32, 32, 0), (300, 128, 64, 64, 1), (1000, 768, 1280, 256, 0),
                                          (4096, 1280, 3840, 256, 2), (4900, 1280, 1280, 128, 0), (513, 5120, 1280, 256, 0)])
def test_gemm_f16_epilogue(M, K, N, bn, act):
    g = torch.Generator(device="cuda").manual_seed(1)
    A = (torch.randn(M, K, device="cuda", generator=g) * 0.5).half()
    W = (torch.randn(N, K, device="cuda", generator=g) * 0.05).half()
    scale = torch.rand(N, device="cuda", generator=g) + 0.5
    shift = torch.randn(N, device="cuda", generator=g) * 0.1
    out = torch.full((M, N), float("nan"), device="cuda", dtype=torch.half)
    epi = L.TcEpilogue(kind=L.EPI_F16, act=act, scale=scale.data_ptr(), shift=shift.data_ptr(), out=out.data_ptr(), ldc=N)
    _gemm(A, W, epi, bn)
    ref = A.float() @ W.float().t() * scale + shift
    ref = [ref, F.relu(ref), F.gelu(ref)][act]
    err = (out.float() - ref).abs().max().item()
    assert torch.isfinite(out.float()).all()
    assert err <= 2e-3 * max(1.0, ref.abs().max().item()), err


def test_gemm_res_f32_table_and_seq_rows():
    g = torch.Generator(device="cuda").manual_seed(2)
    B, T, K, N = 2, 256, 768, 384
    A = (torch.randn(B * T, K, device="cuda", generator=g) * 0.5).half()
    W = (torch.randn(N, K, device="cuda", generator=g) * 0.05).half()
    bias = torch.randn(N, device="cuda", generator=g) * 0.1
    pos = torch.randn(T + 1, N, device="cuda", generator=g)
    out = torch.zeros(B * (T + 1), N, device="cuda")
    epi = L.TcEpilogue(kind=L.EPI_RES_F32, shift=bias.data_ptr(), out=out.data_ptr(), ldc=N, res=pos.data_ptr(), ldres=N,
                       res_mod=T, res_off=1, row_map=L.ROW_SEQ, row_seq=T, row_pad=1, row_off=1)
    _gemm(A, W, epi, 128)
    ref = (A.float() @ W.float().t() + bias).view(B, T, N) + pos[1:]
    got = out.view(B, T + 1, N)
    assert (got[:, 0] == 0).all()
    assert (got[:, 1:] - ref).abs().max().item() < 1e-3


def test_gemm_res_f32_window_unpartition_inplace():
    g = torch.Generator(device="cuda").manual_seed(3)
    B, h, ws, D = 2, 16, 14, 128
    gw = (h + ws - 1) // ws
    M = B * gw * gw * ws * ws
    A = (torch.randn(M, D, device="cuda", generator=g) * 0.5).half()
    W = (torch.randn(D, D, device="cuda", generator=g) * 0.05).half()
    bias = torch.randn(D, device="cuda", generator=g) * 0.1
    x = torch.randn(B, h, h, D, device="cuda", generator=g)
    x0 = x.clone()
    epi = L.TcEpilogue(kind=L.EPI_RES_F32, shift=bias.data_ptr(), out=x.data_ptr(), ldc=D, res=x.data_ptr(), ldres=D,
                       row_map=L.ROW_WINDOW, win_size=ws, win_grid=gw, tok_h=h, tok_w=h)
    _gemm(A, W, epi, 128)
    y = (A.float() @ W.float().t() + bias).view(B, gw, gw, ws, ws, D).permute(0, 1, 3, 2, 4, 5).reshape(B, gw * ws, gw * ws, D)
    ref = x0 + y[:, :h, :h]
    assert (x - ref).abs().max().item() < 1e-3


@pytest.mark.parametrize("Cin,Cout,hin", [(128, 64, 8), (1280, 512, 16)])
def test_convtranspose_scatter(Cin, Cout, hin):
    g = torch.Generator(device="cuda").manual_seed(4)
    B = 2
    x = (torch.randn(B, hin, hin, Cin, device="cuda", generator=g) * 0.5).half()       # NHWC
    w = (torch.randn(Cin, Cout, 2, 2, device="cuda", generator=g) * 0.05)              # reference layout
    bias = torch.randn(Cout, device="cuda", generator=g) * 0.1
    Wp = w.permute(2, 3, 1, 0).reshape(4 * Cout, Cin).contiguous().half()              # (dy,dx,co) x ci
    out = torch.full((B, 2 * hin, 2 * hin, Cout), float("nan"), device="cuda", dtype=torch.half)
    epi = L.TcEpilogue(kind=L.EPI_CONVT, shift=bias.data_ptr(), out=out.data_ptr(), ldc=Cout, ct_cout=Cout, ct_hin=hin, ct_win=hin)
    _gemm(x.view(-1, Cin), Wp, epi, 0)
    ref = F.conv_transpose2d(x.permute(0, 3, 1, 2).float(), Wp.view(2, 2, Cout, Cin).permute(3, 2, 0, 1).float(), bias, stride=2)
    err = (out.permute(0, 3, 1, 2).float() - ref).abs().max().item()
    assert err < 2e-3 * max(1.0, ref.abs().max().item()), err


def _pack_conv_w(w):  # [N, C, 3, 3] -> [N, 9*C] with k = tap*C + c
    return w.permute(0, 2, 3, 1).reshape(w.shape[0], -1).contiguous().half()


@pytest.mark.parametrize("C0,C1,N,H,W", [(64, 0, 64, 16, 16), (128, 64, 128, 32, 64), (64, 64, 64, 128, 128), (256, 256, 256, 64, 64),
                                         (64, 0, 64, 8, 256)])
def test_conv3x3_implicit_gemm(C0, C1, N, H, W):
    g = torch.Generator(device="cuda").manual_seed(5)
    B = 2
    s0 = (torch.randn(B, H, W, C0, device="cuda", generator=g) * 0.5).half()
    s1 = (torch.randn(B, H, W, C1, device="cuda", generator=g) * 0.5).half() if C1 else None
    w = torch.randn(N, C0 + C1, 3, 3, device="cuda", generator=g) * 0.03
    Wp = _pack_conv_w(w)
    scale = torch.rand(N, device="cuda", generator=g) + 0.5
    shift = torch.randn(N, device="cuda", generator=g) * 0.1
    out = torch.full((B, H, W, N), float("nan"), device="cuda", dtype=torch.half)
    epi = L.TcEpilogue(kind=L.EPI_F16, act=L.ACT_RELU, scale=scale.data_ptr(), shift=shift.data_ptr(), out=out.data_ptr(), ldc=N)
    L.check(L.lib().cvb_op_conv3x3_f16(L.ptr(s0), C0, L.ptr(s1), C1, B, H, W, L.ptr(Wp), N, 0, C.byref(epi), L.stream_ptr()), "conv")
    torch.cuda.synchronize()
    xin = torch.cat([s0, s1], -1) if C1 else s0
    ref = F.conv2d(xin.permute(0, 3, 1, 2).float(), Wp.view(N, 3, 3, C0 + C1).permute(0, 3, 1, 2).float(), padding=1)
    ref = F.relu(ref * scale[None, :, None, None] + shift[None, :, None, None])
    err = (out.permute(0, 3, 1, 2).float() - ref).abs().max().item()
    assert err < 2e-3 * max(1.0, ref.abs().max().item()), err


def test_conv3x3_fused_head():
    g = torch.Generator(device="cuda").manual_seed(6)
    B, H, W, Cc, nc = 2, 64, 128, 64, 6
    s0 = (torch.randn(B, H, W, Cc, device="cuda", generator=g) * 0.5).half()
    w = torch.randn(64, Cc, 3, 3, device="cuda", generator=g) * 0.03
    Wp = _pack_conv_w(w)
    scale = torch.rand(64, device="cuda", generator=g) + 0.5
    shift = torch.randn(64, device="cuda", generator=g) * 0.1
    hw = torch.randn(nc, 64, device="cuda", generator=g) * 0.1
    hb = torch.randn(nc, device="cuda", generator=g) * 0.1
    out = torch.full((B, nc, H, W), float("nan"), device="cuda")
    epi = L.TcEpilogue(kind=L.EPI_HEAD, scale=scale.data_ptr(), shift=shift.data_ptr(), head_w=hw.data_ptr(), head_b=hb.data_ptr(),
                       head_nc=nc, head_hw=H * W, head_out=out.data_ptr())
    L.check(L.lib().cvb_op_conv3x3_f16(L.ptr(s0), Cc, None, 0, B, H, W, L.ptr(Wp), 64, 64, C.byref(epi), L.stream_ptr()), "conv-head")
    torch.cuda.synchronize()
    ref = F.conv2d(s0.permute(0, 3, 1, 2).float(), Wp.view(64, 3, 3, Cc).permute(0, 3, 1, 2).float(), padding=1)
    ref = F.relu(ref * scale[None, :, None, None] + shift[None, :, None, None])
    ref = F.conv2d(ref, hw[:, :, None, None], hb)
    assert (out - ref).abs().max().item() < 1e-3


@pytest.mark.parametrize("M,K,N,bn", [(16384, 1280, 3840, 256), (19600, 1280, 1280, 256), (19000, 256, 512, 128), (37888, 128, 64, 64)])
def test_cta_pair_mode_is_bit_identical_to_single_cta(M, K, N, bn):
    g = torch.Generator(device="cuda").manual_seed(11)
    A = (torch.randn(M, K, device="cuda", generator=g) * 0.5).half()
    W = (torch.randn(N, K, device="cuda", generator=g) * 0.05).half()
    shift = torch.randn(N, device="cuda", generator=g) * 0.1
    outs = []
    try:
        for mode in (0, 1):
            L.lib().cvb_tc_set_pair_mode(mode)
            out = torch.full((M, N), float("nan"), device="cuda", dtype=torch.half)
            epi = L.TcEpilogue(kind=L.EPI_F16, act=L.ACT_GELU, shift=shift.data_ptr(), out=out.data_ptr(), ldc=N)
            _gemm(A, W, epi, bn)
            outs.append(out)
    finally:
        L.lib().cvb_tc_set_pair_mode(1)
    assert torch.isfinite(outs[1].float()).all()
    assert torch.equal(outs[0], outs[1])
    ref = F.gelu(A[:2048].float() @ W.float().t() + shift)
    assert (outs[1][:2048].float() - ref).abs().max().item() <= 2e-3 * max(1.0, ref.abs().max().item())


def test_cta_pair_mode_conv_and_head():
    g = torch.Generator(device="cuda").manual_seed(12)
    B, H, W, Cc, nc = 2, 256, 256, 64, 6
    s0 = (torch.randn(B, H, W, Cc, device="cuda", generator=g) * 0.5).half()
    s1 = (torch.randn(B, H, W, Cc, device="cuda", generator=g) * 0.5).half()
    w = torch.randn(64, 2 * Cc, 3, 3, device="cuda", generator=g) * 0.03
    Wp = _pack_conv_w(w)
    scale = torch.rand(64, device="cuda", generator=g) + 0.5
    shift = torch.randn(64, device="cuda", generator=g) * 0.1
    hw = torch.randn(nc, 64, device="cuda", generator=g) * 0.1
    hb = torch.randn(nc, device="cuda", generator=g) * 0.1
    res = []
    try:
        for mode in (0, 1):
            L.lib().cvb_tc_set_pair_mode(mode)
            out = torch.full((B, nc, H, W), float("nan"), device="cuda")
            epi = L.TcEpilogue(kind=L.EPI_HEAD, scale=scale.data_ptr(), shift=shift.data_ptr(), head_w=hw.data_ptr(), head_b=hb.data_ptr(),
                               head_nc=nc, head_hw=H * W, head_out=out.data_ptr())
            L.check(L.lib().cvb_op_conv3x3_f16(L.ptr(s0), Cc, L.ptr(s1), Cc, B, H, W, L.ptr(Wp), 64, 64, C.byref(epi), L.stream_ptr()), "conv-head")
            torch.cuda.synchronize()
            res.append(out)
    finally:
        L.lib().cvb_tc_set_pair_mode(1)
    assert torch.equal(res[0], res[1])
    ref = F.conv2d(torch.cat([s0, s1], -1).permute(0, 3, 1, 2).float(), Wp.view(64, 3, 3, 2 * Cc).permute(0, 3, 1, 2).float(), padding=1)
    ref = F.conv2d(F.relu(ref * scale[None, :, None, None] + shift[None, :, None, None]), hw[:, :, None, None], hb)
    assert (res[1] - ref).abs().max().item() < 1e-3


@pytest.mark.parametrize("C0,C1,N,H,W", [(64, 0, 64, 4, 128), (64, 64, 64, 64, 256), (128, 128, 128, 32, 128), (128, 0, 128, 16, 384)])
def test_patch_resident_conv_matches_kblock_conv_and_reference(C0, C1, N, H, W):
    """csrc/tc_gemm.cu conv_patch_kernel (UMMA descriptors pointing into a halo patch) vs the k-block conv and fp32."""
    g = torch.Generator(device="cuda").manual_seed(21)
    B = 2
    s0 = (torch.randn(B, H, W, C0, device="cuda", generator=g) * 0.5).half()
    s1 = (torch.randn(B, H, W, C1, device="cuda", generator=g) * 0.5).half() if C1 else None
    w = torch.randn(N, C0 + C1, 3, 3, device="cuda", generator=g) * 0.03
    Wp = _pack_conv_w(w)
    scale = torch.rand(N, device="cuda", generator=g) + 0.5
    shift = torch.randn(N, device="cuda", generator=g) * 0.1
    outs = []
    try:
        for mode in (0, 1):
            L.lib().cvb_tc_set_conv_patch_mode(mode)
            out = torch.full((B, H, W, N), float("nan"), device="cuda", dtype=torch.half)
            epi = L.TcEpilogue(kind=L.EPI_F16, act=L.ACT_RELU, scale=scale.data_ptr(), shift=shift.data_ptr(), out=out.data_ptr(), ldc=N)
            L.check(L.lib().cvb_op_conv3x3_f16(L.ptr(s0), C0, L.ptr(s1), C1, B, H, W, L.ptr(Wp), N, N, C.byref(epi), L.stream_ptr()), "conv")
            torch.cuda.synchronize()
            outs.append(out)
    finally:
        L.lib().cvb_tc_set_conv_patch_mode(1)
    xin = torch.cat([s0, s1], -1) if C1 else s0
    ref = F.conv2d(xin.permute(0, 3, 1, 2).float(), Wp.view(N, 3, 3, C0 + C1).permute(0, 3, 1, 2).float(), padding=1)
    ref = F.relu(ref * scale[None, :, None, None] + shift[None, :, None, None])
    tol = 2e-3 * max(1.0, ref.abs().max().item())
    for o in outs:
        assert (o.permute(0, 3, 1, 2).float() - ref).abs().max().item() < tol
    assert (outs[0].float() - outs[1].float()).abs().max().item() < 2 * tol


@pytest.mark.parametrize("M,K,N,bn", [(4096 * 4, 1280, 3840, 256), (19600, 1280, 1280, 256), (513, 5120, 1280, 256), (300, 128, 64, 64)])
def test_gemm_dynamic_tile_scheduler_is_bit_identical(M, K, N, bn):
    """Tiles claimed from a zeroed per-launch counter (SchedRing: single-CTA and CTA-pair mode) give exactly the result of the
    static tile lists, and the counter ends at tiles + one terminator claim per claiming warp."""
    g = torch.Generator(device="cuda").manual_seed(5)
    A = (torch.randn(M, K, device="cuda", generator=g) * 0.5).half()
    W = (torch.randn(N, K, device="cuda", generator=g) * 0.05).half()
    shift = torch.randn(N, device="cuda", generator=g) * 0.1
    outs = []
    for dynamic in (False, True, True):
        out = torch.full((M, N), float("nan"), device="cuda", dtype=torch.half)
        ctr = torch.zeros(1, dtype=torch.int32, device="cuda")
        epi = L.TcEpilogue(kind=L.EPI_F16, act=0, shift=shift.data_ptr(), out=out.data_ptr(), ldc=N,
                           sched_counter=ctr.data_ptr() if dynamic else None)
        _gemm(A, W, epi, bn)
        outs.append(out)
        if dynamic:
            assert int(ctr.item()) > 0
    assert torch.isfinite(outs[0].float()).all()
    assert torch.equal(outs[0], outs[1]) and torch.equal(outs[0], outs[2])
