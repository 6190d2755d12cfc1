"""Whole udaiic training iterations/s on synthetic ACDC-shaped data (north_star's third measurement).

One iteration = what semi_seg/epocher.py:143-187 does per batch: one UNet forward on
cat[labeled, unlabeled, flipped unlabeled], supervised KL on the labeled logits, the IIC regulariser
over (feature layer x sub-head) (:249-284), the UDA term (:215-226), backward, optimizer step.  The
backbone and the 1x1 / linear heads are stock torch/cuDNN (out of scope); what
changes between the arms is the loss path only:

  torch   the losses written in plain PyTorch eager ops, operator for operator what the reference does
          (outer-product joint, one F.conv2d whose filter is the whole map, simplex asserts with
          their host syncs) -- i.e. the reference's loss path on the same GPU, TF32 off
  b200    iic_b200 drop-in modules on softmax maps (checks deferred to one read per iteration)
  fused   iic_b200 with the head softmax fused (IIDSegmentationSmallPathLoss.from_logits, uda_from_logits)

    python tools/udaiic_iteration_bench.py [--K 10 --paddings 1 1 --unlabeled 10 --iters 20] [--per-sample-flips]

--per-sample-flips (added after the round's last GPU run -- not measured yet): the views are aligned with the
reference's seeded per-sample random flips instead of one fixed W flip.  The torch arm then runs the reference's
Python loops (clone/flip per sample + stack, semi_seg/epocher.py:148-149,160-161,264-266) and its supervised
KL_div on a `long` one-hot (epocher.py:165-166); the b200 / fused arms use draw_flip_flags + flip_stack, the UDA
term read through the flips and sup_kl_from_logits.

The UNet below has the reference's topology and channel widths (contrastyou/arch/unet.py:43-133:
5 levels, 16..256 channels, double 3x3 conv+BN+ReLU blocks, nearest-upsample + conv decoders) with
random initial weights; images are random.  Prints one JSON line.
"""
import argparse
import json
import os
import sys
import time

import torch
import torch.nn.functional as F
from torch import nn

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import iic_b200  # noqa: E402


def block(cin, cout):
    return nn.Sequential(nn.Conv2d(cin, cout, 3, padding=1, bias=False), nn.BatchNorm2d(cout), nn.ReLU(inplace=True),
                         nn.Conv2d(cout, cout, 3, padding=1, bias=False), nn.BatchNorm2d(cout), nn.ReLU(inplace=True))


def upblock(cin, cout):
    return nn.Sequential(nn.Upsample(scale_factor=2), nn.Conv2d(cin, cout, 3, padding=1, bias=False),
                         nn.BatchNorm2d(cout), nn.ReLU(inplace=True))


class UNet5(nn.Module):
    """Same topology / widths as the reference UNet; returns logits and the three hooked feature maps."""

    def __init__(self, cin=1, classes=4):
        super().__init__()
        w = [16, 32, 64, 128, 256]
        self.enc = nn.ModuleList([block(cin, w[0])] + [block(w[i], w[i + 1]) for i in range(4)])
        self.up = nn.ModuleList([upblock(w[i + 1], w[i]) for i in (3, 2, 1, 0)])
        self.dec = nn.ModuleList([block(2 * w[i], w[i]) for i in (3, 2, 1, 0)])
        self.head = nn.Conv2d(w[0], classes, 1)

    def forward(self, x):
        skips = []
        for i, e in enumerate(self.enc):
            x = e(x if i == 0 else F.max_pool2d(x, 2))
            skips.append(x)
        conv5 = x
        feats = {}
        for lvl, (u, d) in enumerate(zip(self.up, self.dec)):
            x = d(torch.cat((skips[3 - lvl], u(x)), 1))
            feats[lvl] = x
        return self.head(x), conv5, feats[2], feats[3]      # logits, Conv5, Up_conv3 (32ch,112^2), Up_conv2 (16ch,224^2)


# ---- the reference's loss path in plain eager PyTorch (for the "torch" arm) --------------------------
def _simplex(t):
    s = t.sum(1)
    return bool(torch.allclose(s, torch.ones_like(s), rtol=1e-4, atol=1e-4))


def torch_global_iic(x, y, lamb=1.0):
    assert _simplex(x) and _simplex(y)
    p = (x.unsqueeze(2) * y.unsqueeze(1)).sum(0)
    p = (p + p.t()) / 2.0
    p = p / p.sum()
    pi, pj = p.sum(1, keepdim=True).expand_as(p), p.sum(0, keepdim=True).expand_as(p)
    return (-p * (torch.log(p + 1e-10) - lamb * torch.log(pj + 1e-10) - lamb * torch.log(pi + 1e-10))).sum()


def torch_local_iic(x, y, padding, lamda=1.0):
    assert _simplex(x)
    T = 2 * padding + 1
    pij = F.conv2d(x.permute(1, 0, 2, 3).contiguous(), weight=y.permute(1, 0, 2, 3).contiguous(), padding=padding)
    pij = pij - pij.min().detach() + 1e-16
    pij = pij.permute(2, 3, 0, 1)
    pij = pij / pij.sum(dim=(2, 3), keepdim=True)
    pij = (pij + pij.permute(0, 1, 3, 2)) / 2.0
    pi, pj = pij.sum(2, keepdim=True), pij.sum(3, keepdim=True)
    loss = (-pij * (torch.log(pij + 1e-16) - lamda * torch.log(pi + 1e-16) - lamda * torch.log(pj + 1e-16))).sum() / (T * T)
    if torch.isnan(loss):
        raise RuntimeError(loss)
    return loss


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--K", type=int, default=10)
    ap.add_argument("--subheads", type=int, default=5)
    ap.add_argument("--paddings", type=int, nargs=2, default=[1, 1], help="Up_conv3, Up_conv2 (semi.yaml: 1 3)")
    ap.add_argument("--labeled", type=int, default=4)
    ap.add_argument("--unlabeled", type=int, default=10)
    ap.add_argument("--iters", type=int, default=20)
    ap.add_argument("--arms", nargs="+", default=["torch", "b200", "fused"])
    ap.add_argument("--per-sample-flips", action="store_true")
    a = ap.parse_args()
    dev = torch.device("cuda:0")
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.benchmark = True
    torch.manual_seed(0)
    K, S, nl, nu = a.K, a.subheads, a.labeled, a.unlabeled

    net = UNet5(1, 4).to(dev)
    enc_heads = nn.ModuleList([nn.Linear(256, K) for _ in range(S)]).to(dev)              # ClusterHead: GAP -> Linear -> softmax
    dec3_heads = nn.ModuleList([nn.Conv2d(32, K, 1) for _ in range(S)]).to(dev)           # LocalClusterHead: 1x1 conv -> softmax
    dec2_heads = nn.ModuleList([nn.Conv2d(16, K, 1) for _ in range(S)]).to(dev)
    params = list(net.parameters()) + list(enc_heads.parameters()) + list(dec3_heads.parameters()) + list(dec2_heads.parameters())
    opt = torch.optim.Adam(params, lr=1e-4)

    img_l = torch.rand(nl, 1, 224, 224, device=dev)
    tgt_l = torch.randint(0, 4, (nl, 224, 224), device=dev)
    img_u = torch.rand(nu, 1, 224, 224, device=dev)

    iic_b200.set_check_mode("deferred")
    g_loss = iic_b200.IIDLoss()
    l3 = iic_b200.IIDSegmentationSmallPathLoss(padding=a.paddings[0], patch_size=1024)
    l2 = iic_b200.IIDSegmentationSmallPathLoss(padding=a.paddings[1], patch_size=1024)
    mse = iic_b200.MSELoss()

    def ref_flip_stack(batch, flags):
        """The reference's loop: per sample a clone and up to two flips, then a stack."""
        out = []
        for x, f in zip(batch, flags):
            x = x.clone()
            if f & 1:
                x = x.flip(1)
            if f & 2:
                x = x.flip(2)
            out.append(x)
        return torch.stack(out, dim=0)

    step_no = [0]

    def iteration_flips(arm):
        """Same iteration with seeded per-sample flips and the fused supervised branch (see the module docstring)."""
        step_no[0] += 1
        flags = iic_b200.draw_flip_flags(step_no[0], nu)
        if arm == "torch":
            fl = flags.tolist()
            align = lambda t: ref_flip_stack(t, fl)                                  # noqa: E731
        else:
            fdev = flags.to(dev, non_blocking=True)
            align = lambda t: iic_b200.flip_stack(t, fdev)                           # noqa: E731
        img_u_tf = align(img_u)
        logits, conv5, up3, up2 = net(torch.cat((img_l, img_u, img_u_tf)))
        if arm == "torch":
            onehot = F.one_hot(tgt_l, 4).permute(0, 3, 1, 2)                          # class2one_hot: a `long` tensor
            p = logits[:nl].softmax(1)
            sup = (-onehot * torch.log((p + 1e-16) / (onehot + 1e-16))).sum(1).mean()
        else:
            sup = iic_b200.sup_kl_from_logits(logits[:nl], tgt_l)
        lu, lu_tf = logits[nl:nl + nu], logits[nl + nu:]
        f5 = conv5[nl:].mean((2, 3))
        f3a, f3b = align(up3[nl:nl + nu]), up3[nl + nu:]
        f2a, f2b = align(up2[nl:nl + nu]), up2[nl + nu:]
        iic_terms = []
        for s in range(S):
            e = enc_heads[s](f5).softmax(1)
            ea, eb = e[:nu], e[nu:]
            z3a, z3b = dec3_heads[s](f3a), dec3_heads[s](f3b)
            z2a, z2b = dec2_heads[s](f2a), dec2_heads[s](f2b)
            if arm == "torch":
                t = torch_global_iic(ea, eb) + 0.5 * torch_local_iic(z3a.softmax(1), z3b.softmax(1), a.paddings[0]) \
                    + 0.5 * torch_local_iic(z2a.softmax(1), z2b.softmax(1), a.paddings[1])
            elif arm == "b200":
                t = g_loss(ea, eb)[0] + 0.5 * l3(z3a.softmax(1), z3b.softmax(1)) + 0.5 * l2(z2a.softmax(1), z2b.softmax(1))
            else:
                t = g_loss(ea, eb)[0] + 0.5 * l3.from_logits(z3a, z3b) + 0.5 * l2.from_logits(z2a, z2b)
            iic_terms.append(t)
        iic = sum(iic_terms) / S / 2.0
        if arm == "torch":
            uda = F.mse_loss(lu_tf.softmax(1), align(lu).softmax(1).detach())
        else:
            uda = iic_b200.uda_from_logits(lu_tf, lu, "mse", teacher_flips=fdev)
        total = sup + 5.0 * uda + 0.1 * iic
        opt.zero_grad(set_to_none=True)
        total.backward()
        opt.step()
        if arm != "torch":
            iic_b200.raise_if_flagged(dev)
        return total

    def iteration(arm):
        if a.per_sample_flips:
            return iteration_flips(arm)
        img_u_tf = img_u.flip(3)
        logits, conv5, up3, up2 = net(torch.cat((img_l, img_u, img_u_tf)))
        sup = F.cross_entropy(logits[:nl], tgt_l)
        lu, lu_tf = logits[nl:nl + nu].flip(3), logits[nl + nu:]                 # align view 1 with view 2
        f5 = conv5[nl:].mean((2, 3))
        f3a, f3b = up3[nl:nl + nu].flip(3), up3[nl + nu:]
        f2a, f2b = up2[nl:nl + nu].flip(3), up2[nl + nu:]
        iic_terms = []
        for s in range(S):
            e = enc_heads[s](f5).softmax(1)
            ea, eb = e[:nu], e[nu:]
            z3a, z3b = dec3_heads[s](f3a), dec3_heads[s](f3b)
            z2a, z2b = dec2_heads[s](f2a), dec2_heads[s](f2b)
            if arm == "torch":
                t = torch_global_iic(ea, eb) + 0.5 * torch_local_iic(z3a.softmax(1), z3b.softmax(1), a.paddings[0]) \
                    + 0.5 * torch_local_iic(z2a.softmax(1), z2b.softmax(1), a.paddings[1])
            elif arm == "b200":
                t = g_loss(ea, eb)[0] + 0.5 * l3(z3a.softmax(1), z3b.softmax(1)) + 0.5 * l2(z2a.softmax(1), z2b.softmax(1))
            else:
                t = g_loss(ea, eb)[0] + 0.5 * l3.from_logits(z3a, z3b) + 0.5 * l2.from_logits(z2a, z2b)
            iic_terms.append(t)
        iic = sum(iic_terms) / S / 2.0
        if arm == "torch":
            uda = F.mse_loss(lu_tf.softmax(1), lu.softmax(1).detach())
        elif arm == "b200":
            uda = mse(lu_tf.softmax(1), lu.softmax(1).detach())
        else:
            uda = iic_b200.uda_from_logits(lu_tf, lu, "mse")
        total = sup + 5.0 * uda + 0.1 * iic
        opt.zero_grad(set_to_none=True)
        total.backward()
        opt.step()
        if arm != "torch":
            iic_b200.raise_if_flagged(dev)           # the one host read per iteration
        return total

    out = {"config": {"K": K, "subheads": S, "paddings": a.paddings, "labeled": nl, "unlabeled": nu, "size": 224,
                      "layers": ["Conv5 (global)", "Up_conv3 112^2", "Up_conv2 224^2"], "optimizer": "Adam",
                      "data": "synthetic", "iters": a.iters, "per_sample_flips": a.per_sample_flips}}
    for arm in a.arms:
        try:
            for _ in range(3):
                loss = iteration(arm)
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            for _ in range(a.iters):
                loss = iteration(arm)
            torch.cuda.synchronize()
            dt = (time.perf_counter() - t0) / a.iters
            out[arm] = {"it_per_s": round(1.0 / dt, 2), "ms_per_it": round(dt * 1e3, 2), "loss": round(float(loss), 5)}
        except Exception as e:  # noqa: BLE001
            out[arm] = {"error": f"{type(e).__name__}: {str(e)[:200]}"}
    print(json.dumps(out))


if __name__ == "__main__":
    main()
