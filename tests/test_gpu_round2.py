"""Round-2 GPU parity cases: every dispatch branch of the local term reached WITHOUT environment variables (VERDICT r1
weak #1), the forced tensor-core kernels through the run-time option table, the fused finish launch (one term and many
terms: csrc/finish.cu) against separate oracle calls, and the device-side loss combination of SURVEY 8a row A7.

Tolerances as in test_gpu_parity.py: loss rel. err <= 1e-5, gradients <= 1e-4 (max-norm relative) against the fp64 oracle.
"""
import contextlib
import os
import sys

import numpy as np
import pytest
import torch

from conftest import ROOT, relmax

sys.path.insert(0, os.path.join(ROOT, "oracle"))
import iic_oracle as O  # noqa: E402
from test_gpu_parity import GRAD_RTOL, _logit_views, _loss_close, views  # noqa: E402

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def iic(cuda_device):
    import iic_b200
    return iic_b200


@contextlib.contextmanager
def option(iic, name, value):
    old = iic._lib.set_option(name, value)
    try:
        yield
    finally:
        iic._lib.set_option(name, old)


def _check_local(iic, dev, B, K, H, W, pad, patch, seed):
    rng = np.random.default_rng(seed)
    x, y = views(rng, B, K, H, W)
    xd = torch.from_numpy(x).to(dev).requires_grad_(True)
    yd = torch.from_numpy(y).to(dev).requires_grad_(True)
    loss = iic.IIDSegmentationSmallPathLoss(padding=pad, patch_size=patch)(xd, yd)
    loss.backward()
    ol, ogx, ogy = O.iid_segmentation_small_path_loss(x, y, pad, patch, with_grads=True)
    ex, ey = relmax(xd.grad.cpu().numpy(), ogx), relmax(yd.grad.cpu().numpy(), ogy)
    print(f"loss err {abs(loss.item() - ol) / abs(ol):.2e}, grad err {max(ex, ey):.2e}")
    assert _loss_close(loss.item(), ol), (loss.item(), ol)
    assert ex <= GRAD_RTOL and ey <= GRAD_RTOL, (ex, ey)
    return loss.item(), xd.grad, yd.grad


# shapes that reach, by themselves, the branches the round-1 suite never launched
@pytest.mark.parametrize("B,K,H,W,pad,patch", [
    (16, 20, 224, 224, 1, 512),    # K = 20, padding 1, n_items >= 2 * SMs: local_bwd_tcrb_kernel<3> (the yaml head at 224^2)
    (8, 20, 112, 112, 1, 512),     # config 3's Up_conv3 layer per GPU: FFMA2 (too few row blocks for the tensor cores)
    (2, 20, 64, 512, 3, 1024),     # config-4 width: four 128-column panels in the row-block backward, packed joint
    (12, 10, 224, 224, 1, 512),    # config-2 shape, > 8 rows per SM twice over: local_bwd_tcrb10_kernel with two chunks
])
def test_dispatch_branches_without_env(iic, cuda_device, B, K, H, W, pad, patch):
    _check_local(iic, cuda_device, B, K, H, W, pad, patch, 777 + B + K + W)


@pytest.mark.parametrize("opt", ["tcp_p1", "tcrb_p1", "tc10_force", "tc10_tf32", "no_tc", "no_fast", "no_tma", "no_fused_epilogue",
                                 "fin_last_cta_epilogue"])
def test_forced_kernel_families(iic, cuda_device, opt):
    """The same small inputs through the kernel family a switch forces must agree with the oracle (and therefore with
    the default dispatch).  K = 20 / padding 1 for the packed joint and the row-block backward, K = 10 for the rest."""
    shape = (2, 20, 40, 64, 1, 512) if opt in ("tcp_p1", "tcrb_p1") else (2, 10, 40, 64, 1, 512)
    with contextlib.ExitStack() as stack:
        stack.enter_context(option(iic, opt, 1))
        if opt == "tc10_tf32":                          # small map: the K = 10 tensor-core backward must be forced as well
            stack.enter_context(option(iic, "tc10_force", 1))
        a = _check_local(iic, cuda_device, *shape, seed=4242)
    b = _check_local(iic, cuda_device, *shape, seed=4242)
    assert abs(a[0] - b[0]) <= 2e-6 * abs(b[0])
    assert relmax(a[1].cpu().numpy(), b[1].cpu().numpy()) <= 5e-5


def test_from_logits_wide_map_falls_back(iic, cuda_device):
    """W > 248 must take softmax + the probability path in BOTH directions (the fused forward used to accept
    it while the fused backward refused)."""
    rng = np.random.default_rng(9)
    l1, l2 = _logit_views(rng, 2, 10, 32, 256)
    a = torch.from_numpy(l1).to(cuda_device).requires_grad_(True)
    b = torch.from_numpy(l2).to(cuda_device).requires_grad_(True)
    loss = iic.IIDSegmentationSmallPathLoss(padding=1, patch_size=512).from_logits(a, b)
    loss.backward()
    p1, p2 = O.softmax(l1), O.softmax(l2)
    ol, g1, g2 = O.iid_segmentation_small_path_loss(p1, p2, 1, 512, with_grads=True)
    assert _loss_close(loss.item(), ol)
    assert relmax(a.grad.cpu().numpy(), O.softmax_backward(p1, g1)) <= GRAD_RTOL
    assert relmax(b.grad.cpu().numpy(), O.softmax_backward(p2, g2)) <= GRAD_RTOL


@pytest.mark.parametrize("B,H,W,T,force", [(6, 224, 224, 1.0, False), (3, 37, 44, 0.7, True), (2, 40, 248, 1.0, True)])
def test_from_logits_tensor_core_backward(iic, cuda_device, B, H, W, T, force):
    """The from-logits form of the fp16-split tensor-core backward (softmax in its transform warps, softmax adjoint in
    its drain): chosen by itself at config-2 size, forced on small / ragged maps; loss and LOGIT gradients against the
    fp64 oracle chained through its own softmax (contrastyou/trainer/_utils.py:15-23 feeding iic_loss.py:107-149)."""
    rng = np.random.default_rng(600 + H + W)
    l1, l2 = _logit_views(rng, B, 10, H, W)
    with contextlib.ExitStack() as stack:
        if force:
            stack.enter_context(option(iic, "tc10_force", 1))
        a = torch.from_numpy(l1).to(cuda_device).requires_grad_(True)
        b = torch.from_numpy(l2).to(cuda_device).requires_grad_(True)
        loss = iic.IIDSegmentationSmallPathLoss(padding=1, patch_size=512).from_logits(a, b, T=T)
        (0.5 * loss).backward()
    p1, p2 = O.softmax(l1, 1, T), O.softmax(l2, 1, T)
    ol, g1, g2 = O.iid_segmentation_small_path_loss(p1, p2, 1, 512, with_grads=True)
    ex = relmax(a.grad.cpu().numpy(), 0.5 * O.softmax_backward(p1, g1, 1, T))
    ey = relmax(b.grad.cpu().numpy(), 0.5 * O.softmax_backward(p2, g2, 1, T))
    print(f"loss err {abs(loss.item() - ol) / abs(ol):.2e}, grad err {max(ex, ey):.2e}")
    assert _loss_close(loss.item(), ol), (loss.item(), ol)
    assert ex <= GRAD_RTOL and ey <= GRAD_RTOL, (ex, ey)


# ---------------------------------------------------------------------------------------------------
# many terms, ONE finish launch: the (layer x sub-head) loop of semi_seg/epocher.py:249-277
# ---------------------------------------------------------------------------------------------------
def _heads(rng, S, B, K, H, W, dev):
    """S sub-heads as channel-block views of ONE (2B, S*K, H, W) tensor, like the batched cluster heads emit them."""
    xs, ys = zip(*[views(rng, B, K, H, W) for _ in range(S)])
    both = np.concatenate([np.concatenate(xs, 1), np.concatenate(ys, 1)], 0)            # (2B, S*K, H, W)
    t = torch.from_numpy(both).to(dev).requires_grad_(True)
    v = t.view(2 * B, S, K, H, W)
    pairs = [tuple(torch.chunk(v[:, s], 2, 0)) for s in range(S)]
    return t, pairs, xs, ys


@pytest.mark.parametrize("K,pad,H,W", [(10, 1, 40, 56), (20, 3, 24, 32), (20, 1, 32, 48)])
def test_five_heads_one_finish(iic, cuda_device, K, pad, H, W):
    """S = 5 sub-heads of one decoder layer through iic_losses vs five separate oracle calls + average_iter."""
    S, B = 5, 2
    rng = np.random.default_rng(31 + K + pad)
    t, pairs, xs, ys = _heads(rng, S, B, K, H, W, cuda_device)
    crit = iic.IIDSegmentationSmallPathLoss(padding=pad, patch_size=1024)
    losses = iic.iic_losses([(crit, a, b) for a, b in pairs])
    total = sum(losses) / float(S)
    total.backward()
    g = t.grad.cpu().numpy().reshape(2 * B, S, K, H, W)
    for s in range(S):
        ol, ogx, ogy = O.iid_segmentation_small_path_loss(xs[s], ys[s], pad, 1024, with_grads=True)
        assert _loss_close(losses[s].item(), ol), (s, losses[s].item(), ol)
        assert relmax(g[:B, s], ogx / S) <= GRAD_RTOL and relmax(g[B:, s], ogy / S) <= GRAD_RTOL
    # and the one-by-one public calls agree (same joint kernels; a single small term runs its epilogue inside the
    # finish launch, a batch in the multi-CTA epilogue kernel: same fp64 arithmetic, different summation order)
    t2 = t.detach().clone().requires_grad_(True)
    v2 = t2.view(2 * B, S, K, H, W)
    for s in range(S):
        a, b = torch.chunk(v2[:, s], 2, 0)
        assert abs(crit(a, b).item() - losses[s].item()) <= 1e-6 * abs(losses[s].item())


def test_config3_iteration_one_finish(iic, cuda_device):
    """BASELINE config 3 in miniature: 5 sub-heads x {Conv5 global, Up_conv3 padding 1, Up_conv2 padding 3}, K = 20,
    evaluated by iic_regularization (ONE finish launch for the 15 terms) against 15 oracle calls combined with the
    reference's average_iter / weighted_average_iter (contrastyou/helper/utils.py:46-56)."""
    S, B, K = 5, 2, 20
    rng = np.random.default_rng(2025)
    wrapper = iic.IICLossWrapper(["Conv5", "Up_conv3", "Up_conv2"], [1, 3], 1024)
    fi = [0.5, 0.25, 0.25]                                             # semi.yaml [1, .5, .5] normalised
    base = rng.standard_normal((S, B, K)) * 2
    gx = [O.softmax(base[s] + 0.7 * rng.standard_normal((B, K))).astype(np.float32) for s in range(S)]
    gy = [O.softmax(base[s] + 0.7 * rng.standard_normal((B, K))).astype(np.float32) for s in range(S)]
    g_in = [(torch.from_numpy(a).to(cuda_device).requires_grad_(True), torch.from_numpy(b).to(cuda_device).requires_grad_(True))
            for a, b in zip(gx, gy)]
    t3, p3, x3, y3 = _heads(rng, S, B, K, 16, 24, cuda_device)
    t2, p2, x2, y2 = _heads(rng, S, B, K, 32, 48, cuda_device)
    reg, per_layer = iic.iic_regularization([g_in, p3, p2], [wrapper["Conv5"], wrapper["Up_conv3"], wrapper["Up_conv2"]], fi)
    reg.backward()
    o_layers, o_grads = [], []
    o_layers.append(np.mean([O.iid_loss(a, b)[0] for a, b in zip(gx, gy)]))
    l3 = [O.iid_segmentation_small_path_loss(a, b, 1, 1024, with_grads=True) for a, b in zip(x3, y3)]
    l2 = [O.iid_segmentation_small_path_loss(a, b, 3, 1024, with_grads=True) for a, b in zip(x2, y2)]
    o_layers += [np.mean([r[0] for r in l3]), np.mean([r[0] for r in l2])]
    expect = sum(w * l for w, l in zip(fi, o_layers)) / (sum(fi) + 1e-16)
    assert abs(reg.item() - expect) <= 1e-5 * abs(expect), (reg.item(), expect)
    assert relmax(per_layer.detach().cpu().numpy(), np.asarray(o_layers)) <= 1e-5
    g2 = t2.grad.cpu().numpy().reshape(2 * B, S, K, 32, 48)
    for s in range(S):
        assert relmax(g2[:B, s], fi[2] / S * l2[s][1]) <= GRAD_RTOL
        assert relmax(g2[B:, s], fi[2] / S * l2[s][2]) <= GRAD_RTOL
    ox, _ = O.iid_loss_grads(gx[0], gy[0])
    assert relmax(g_in[0][0].grad.cpu().numpy(), fi[0] / S * ox) <= GRAD_RTOL


def test_fused_and_unfused_epilogue_agree(iic, cuda_device):
    """The finish launch's in-kernel epilogue (one warp per displacement) against the multi-CTA epilogue kernels."""
    rng = np.random.default_rng(8)
    x, y = views(rng, 3, 10, 48, 64)
    res = []
    for v in (0, 1):
        with option(iic, "no_fused_epilogue", v):
            xd = torch.from_numpy(x).to(cuda_device).requires_grad_(True)
            yd = torch.from_numpy(y).to(cuda_device).requires_grad_(True)
            loss = iic.IIDSegmentationSmallPathLoss(padding=1, patch_size=32)(xd, yd)      # several patches
            loss.backward()
            res.append((loss.item(), xd.grad.clone()))
    assert abs(res[0][0] - res[1][0]) <= 1e-6 * abs(res[1][0])
    assert relmax(res[0][1].cpu().numpy(), res[1][1].cpu().numpy()) <= 1e-5
    # a small batch: the default (finish launch + one multi-CTA epilogue launch) against the epilogue in the finish launch's last CTA
    x, y = views(rng, 2, 10, 40, 56)
    g0 = torch.rand(2, 8, 10, device=cuda_device).softmax(2)
    res = []
    for v in (0, 1):
        with option(iic, "fin_last_cta_epilogue", v):
            xd = torch.from_numpy(x).to(cuda_device).requires_grad_(True)
            yd = torch.from_numpy(y).to(cuda_device).requires_grad_(True)
            gx = g0[0].clone().requires_grad_(True)
            gy = g0[1].clone().requires_grad_(True)
            ll, lg = iic.iic_losses([(iic.IIDSegmentationSmallPathLoss(padding=1, patch_size=512), xd, yd), (iic.IIDLoss(), gx, gy)])
            (ll + lg[0]).backward()
            res.append((ll.item(), lg[0].item(), xd.grad.clone(), gx.grad.clone()))
    assert abs(res[0][0] - res[1][0]) <= 1e-6 * abs(res[1][0]) and abs(res[0][1] - res[1][1]) <= 1e-6 * abs(res[1][1])
    assert relmax(res[0][2].cpu().numpy(), res[1][2].cpu().numpy()) <= 1e-5
    assert relmax(res[0][3].cpu().numpy(), res[1][3].cpu().numpy()) <= 1e-5


def test_global_term_through_finish(iic, cuda_device):
    """IIDLoss for the udaiic sizes is ONE launch (joint from the rows + epilogue inside iic_finish): all three outputs
    and the non-contiguous-row case."""
    rng = np.random.default_rng(77)
    N, K = 64, 20
    both = O.softmax(rng.standard_normal((2 * N, 3, K)) * 2, axis=2).astype(np.float32)     # rows with stride 3*K
    t = torch.from_numpy(both).to(cuda_device)
    a, b = t[:N, 1].requires_grad_(True), t[N:, 1].requires_grad_(True)
    assert a.stride(0) == 3 * K
    loss, nl, P = iic.IIDLoss(lamb=1.5)(a, b)
    (loss + 0.5 * nl).backward()
    o1, o2, oP = O.iid_loss(both[:N, 1], both[N:, 1], 1.5)
    assert _loss_close(loss.item(), o1) and _loss_close(nl.item(), o2)
    assert relmax(P.detach().cpu().numpy(), oP) < 1e-6


def test_combine_iic_losses_on_device(iic, cuda_device):
    """SURVEY 8a row A7 on the GPU: combine_iic_losses == weighted_average_iter(average_iter(...)) of the reference's
    helper (contrastyou/helper/utils.py:46-56), values and gradients, on device tensors."""
    from iic_b200.semi_seg._utils import average_iter, combine_iic_losses, weighted_average_iter
    torch.manual_seed(3)
    raw = [torch.randn(5, device=cuda_device, requires_grad=True), torch.randn(5, device=cuda_device, requires_grad=True),
           torch.randn(3, device=cuda_device, requires_grad=True)]
    fi = [0.5, 0.25, 0.25]
    total, per_layer = combine_iic_losses([list(r.unbind(0)) for r in raw], fi)
    ref = weighted_average_iter([average_iter(list(r.unbind(0))) for r in raw], fi)
    assert abs(total.item() - ref.item()) <= 1e-6 * abs(ref.item())
    assert torch.allclose(per_layer, torch.stack([r.mean() for r in raw]), rtol=1e-6, atol=1e-7)
    g1 = torch.autograd.grad(total, raw, retain_graph=True)
    g2 = torch.autograd.grad(ref, raw)
    for a, b in zip(g1, g2):
        assert torch.allclose(a, b, rtol=1e-6, atol=1e-8)


# ---------------------------------------------------------------------------------------------------
# tensor-core joint for K <= 10, 3 x 3 window (csrc/local_fwd_tcj10.cu): fp16 hi/lo operands, accumulator sets drained every
# row pair.  Against the fp64 convolution of iic_loss.py:120-123 and against the FFMA2 joint it replaces on large maps.
# ---------------------------------------------------------------------------------------------------
def _joint64(x, y, pad):
    J = torch.nn.functional.conv2d(x.double().permute(1, 0, 2, 3), y.double().permute(1, 0, 2, 3), padding=pad)
    return J.permute(2, 3, 0, 1).contiguous()


@pytest.mark.parametrize("B,K,H,W,sharp,strided", [
    (32, 10, 224, 224, 1.0, False),   # BASELINE config 2
    (32, 10, 224, 224, 8.0, False),   # peaked maps: most products are tiny, a few are near 1
    (9, 10, 135, 224, 2.0, False),    # odd chunk lengths (row pairs cut by the image end), the widest map the kernel takes
    (20, 7, 64, 40, 1.0, False),      # fewer than 10 clusters, narrow map
    (11, 9, 112, 96, 1.0, True),      # channel block of a wider head output (sample and channel strides not dense)
])
def test_tensor_core_joint_k10(iic, cuda_device, B, K, H, W, sharp, strided):
    gen = torch.Generator(device=cuda_device).manual_seed(77 + B + K)
    if strided:
        full = (sharp * torch.randn(B, 3 * K, H, W, device=cuda_device, generator=gen))
        x = full[:, K:2 * K].softmax(1)
        big = torch.zeros(B, 3 * K, H, W, device=cuda_device)
        big[:, K:2 * K] = x
        x = big[:, K:2 * K]
        assert not x.is_contiguous()
    else:
        x = (sharp * torch.randn(B, K, H, W, device=cuda_device, generator=gen)).softmax(1)
    y = (sharp * torch.randn(B, K, H, W, device=cuda_device, generator=gen)).softmax(1)
    ref = _joint64(x, y, 1)
    J = iic.ops._local_joint(x, y, None, 1, H, W, H, W, check_simplex=True)[0]
    with option(iic, "no_tcj10", 1):
        Jf = iic.ops._local_joint(x, y, None, 1, H, W, H, W, check_simplex=True)[0]
    iic.raise_if_flagged(cuda_device)
    mass = ref.sum().item() / 9.0
    err_tc = (J - ref).abs().max().item() / mass
    err_ff = (Jf - ref).abs().max().item() / mass
    rel_tc = ((J - ref).abs() / ref.clamp_min(1e-300)).max().item()
    print(f"tensor-core joint: max err / mass {err_tc:.2e} (FFMA2 {err_ff:.2e}), max rel {rel_tc:.2e}")
    assert not torch.equal(J, Jf), "the option must select a different kernel"
    assert err_tc <= 5e-8 and rel_tc <= 4e-6
    # the assertion of iic_loss.py:113 rides along in the staging warps
    bad = x.clone()
    bad[B // 2, 0, H // 2, W // 3] += 0.01
    with iic.check_mode("deferred"):
        iic.ops._local_joint(bad, y, None, 1, H, W, H, W, check_simplex=True)
        with pytest.raises(AssertionError):
            iic.raise_if_flagged(cuda_device)


def test_tensor_core_joint_loss_and_gradients(iic, cuda_device):
    """The whole local term at config-2 size with the tensor-core joint against the same term with the FFMA2 joint (which the
    oracle tests pin at small batch): loss to 2e-6, gradients to 1e-4 of their maximum."""
    B, K, H, W = 32, 10, 224, 224
    gen = torch.Generator(device=cuda_device).manual_seed(5)
    base = torch.randn(B, K, H // 8, W // 8, device=cuda_device, generator=gen) * 3
    base = torch.nn.functional.interpolate(base, size=(H, W), mode="bilinear", align_corners=False)
    x = (base + 0.5 * torch.randn(B, K, H, W, device=cuda_device, generator=gen)).softmax(1)
    y = (base + 0.5 * torch.randn(B, K, H, W, device=cuda_device, generator=gen)).softmax(1)
    crit = iic.IIDSegmentationSmallPathLoss(padding=1, patch_size=512)
    out = {}
    for name, off in (("tc", 0), ("ffma2", 1)):
        with option(iic, "no_tcj10", off):
            xr, yr = x.clone().requires_grad_(True), y.clone().requires_grad_(True)
            loss = crit(xr, yr)
            loss.backward()
            out[name] = (loss.item(), xr.grad, yr.grad)
    lt, lf = out["tc"][0], out["ffma2"][0]
    gerr = max(((out["tc"][i] - out["ffma2"][i]).abs().max() / out["ffma2"][i].abs().max()).item() for i in (1, 2))
    print(f"loss tc {lt:.9f} ffma2 {lf:.9f} rel {abs(lt - lf) / abs(lf):.2e}; gradient max-norm rel {gerr:.2e}")
    assert abs(lt - lf) <= 2e-6 * abs(lf)
    assert gerr <= GRAD_RTOL
