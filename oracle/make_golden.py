"""Generate tests/golden/*.npz by RUNNING THE REFERENCE ITSELF -- TEST INFRASTRUCTURE ONLY.

Run in the build container (where /root/reference is mounted):

    python oracle/make_golden.py

Each fixture stores the float32 inputs a caller would pass, and what the UNMODIFIED reference
classes (loaded by oracle/ref_loader.py) return for them:
  * ``*_f64``: the reference run on ``.double()`` copies of the inputs -- the ground truth;
  * ``*_f32``: the reference run on the float32 inputs -- its own noise floor.
Gradients come from the reference's autograd.  The fixtures are small (a few hundred KB in total)
so they can be committed; the GPU box never needs /root/reference.
"""
from __future__ import annotations

import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import ref_loader  # noqa: E402

OUT = os.path.join(os.path.dirname(HERE), "tests", "golden")


def correlated_views(rng, B, K, H, W, noise=0.5, scale=3.0):
    """SURVEY 8(d) synthetic family: two noisy views of one smooth logit field -> MI is O(0.1-1)."""
    ch, cw = max(H // 4, 1), max(W // 4, 1)
    coarse = rng.standard_normal((B, K, ch, cw)) * scale
    ys = (np.arange(H) * ch // H)
    xs = (np.arange(W) * cw // W)
    base = coarse[:, :, ys][:, :, :, xs]
    l1 = base + noise * rng.standard_normal((B, K, H, W))
    l2 = base + noise * rng.standard_normal((B, K, H, W))
    return l1.astype(np.float32), l2.astype(np.float32)


def _relmax(a, ref):
    return float(np.abs(a.astype(np.float64) - ref).max() / max(np.abs(ref).max(), 1e-300))


def _softmax32(torch, logits, dim=1):
    return torch.from_numpy(logits).softmax(dim)


def run_global(ns, x32, y32, lamb):
    torch = ns.torch
    out = {}
    for tag, dt in (("f64", torch.float64), ("f32", torch.float32)):
        x = x32.to(dt).clone().requires_grad_(True)
        y = y32.to(dt).clone().requires_grad_(True)
        crit = ns.IIDLoss(lamb=lamb)
        loss, loss_nl, P = crit(x, y)
        gx, gy = torch.autograd.grad(loss, (x, y), retain_graph=True)
        # a second functional of all three outputs, to pin the full backward (g1, g2, gP)
        w = torch.linspace(-1.0, 1.0, P.numel(), dtype=dt).reshape(P.shape)
        full = 0.7 * loss - 0.3 * loss_nl + (w * P).sum()
        fx, fy = torch.autograd.grad(full, (x, y))
        out[f"loss_{tag}"] = loss.item()
        out[f"loss_no_lamb_{tag}"] = loss_nl.item()
        if tag == "f64":
            out.update({"P_f64": P.detach().numpy(), "gx_f64": gx.numpy(), "gy_f64": gy.numpy(),
                        "fullgx_f64": fx.numpy(), "fullgy_f64": fy.numpy()})
        else:  # the reference's own fp32 deviation from its fp64 run (max-norm relative)
            out["gerr_f32"] = max(_relmax(gx.numpy(), out["gx_f64"]), _relmax(gy.numpy(), out["gy_f64"]))
    out["joint_nosym_f64"] = ns.compute_joint(x32.double(), y32.double(), symmetric=False).numpy()
    return out


def run_local(ns, x32, y32, padding, lamda, mask32=None, patch_size=None):
    torch = ns.torch
    out = {}
    for tag, dt in (("f64", torch.float64), ("f32", torch.float32)):
        x = x32.to(dt).clone().requires_grad_(True)
        y = y32.to(dt).clone().requires_grad_(True)
        m = None if mask32 is None else mask32.to(dt)
        if patch_size is None:
            crit = ns.IIDSegmentationLoss(lamda=lamda, padding=padding)
        else:
            crit = ns.IIDSegmentationSmallPathLoss(lamda=lamda, padding=padding, patch_size=patch_size)
        loss = crit(x, y, m) if m is not None else crit(x, y)
        gx, gy = torch.autograd.grad(loss, (x, y))
        out[f"loss_{tag}"] = loss.item()
        if tag == "f64":
            out.update({"gx_f64": gx.numpy(), "gy_f64": gy.numpy()})
        else:
            out["gerr_f32"] = max(_relmax(gx.numpy(), out["gx_f64"]), _relmax(gy.numpy(), out["gy_f64"]))
    return out


def run_uda(ns, p32, t32, kind, weight=None):
    torch = ns.torch
    out = {}
    for tag, dt in (("f64", torch.float64), ("f32", torch.float32)):
        p = p32.to(dt).clone().requires_grad_(True)
        t = t32.to(dt)
        if kind == "mse":
            loss = torch.nn.MSELoss()(p, t)
        else:
            crit = ns.KL_div(weight=weight, verbose=False)
            if weight is not None:
                crit._weight = crit._weight.to(dt)
            loss = crit(p, t)
        (g,) = torch.autograd.grad(loss, (p,))
        out[f"loss_{tag}"] = loss.item()
        if tag == "f64":
            out["g_f64"] = g.numpy()
        else:
            out["gerr_f32"] = _relmax(g.numpy(), out["g_f64"])
    return out


def run_sup(ns, logits32, labels, weight=None):
    """The reference's supervised branch, verbatim (semi_seg/epocher.py:165-166,183-184): a `long` one-hot from
    class2one_hot, KL_div on the softmax, UniversalDice.add on the argmax."""
    torch = ns.torch
    C = logits32.shape[1]
    out = {}
    for tag, dt in (("f64", torch.float64), ("f32", torch.float32)):
        lg = logits32.to(dt).clone().requires_grad_(True)
        crit = ns.KL_div(weight=weight, verbose=False)
        if weight is not None:
            crit._weight = crit._weight.to(dt)
        onehot_target = ns.class2one_hot(labels.squeeze(1), C)
        loss = crit(lg.softmax(1), onehot_target)
        (g,) = torch.autograd.grad(loss, (lg,))
        out[f"loss_{tag}"] = loss.item()
        if tag == "f64":
            out["g_f64"] = g.numpy()
        else:
            out["gerr_f32"] = _relmax(g.numpy(), out["g_f64"])
    meter = ref_loader.load_dice()(C=C)
    meter.add(logits32.max(1)[1], labels.squeeze(1), group_name=["g"] * len(labels))
    out["intersection"] = meter._intersections[0].numpy()
    out["union"] = meter._unions[0].numpy()
    out["dice"] = meter.log.numpy()[0]
    return out


def make_sup(ns):
    """Supervised-branch fixtures; their own generator so the older fixtures stay byte-identical."""
    torch = ns.torch
    rng = np.random.default_rng(20260119)
    for name, shape, weight in (("s_2x4x6x5", (2, 4, 6, 5), None),            # ragged rows: scalar kernels
                                ("s_3x4x16x16", (3, 4, 16, 16), None),        # 16-byte rows: vector kernels
                                ("s_w_2x4x12x12", (2, 4, 12, 12), [1.0, 2.0, 0.5, 1.5]),   # square: kl_losses.py:118 transposes H and W
                                ("s_2x2x7x9", (2, 2, 7, 9), None),
                                ("s_1x8x8x8", (1, 8, 8, 8), None)):
        B, C = shape[:2]
        labels = torch.from_numpy(rng.integers(0, C, size=(B, 1) + shape[2:]).astype(np.int64))
        # logits that agree with the labels on roughly half of the pixels, so the Dice counts are not trivial
        lg = rng.standard_normal(shape) * 2
        agree = rng.random((B,) + shape[2:]) < 0.5
        onehot = np.stack([labels.squeeze(1).numpy() == c for c in range(C)], axis=1)
        lg = lg + 4.0 * onehot * agree[:, None]
        lg = torch.from_numpy(lg.astype(np.float32))
        res = run_sup(ns, lg, labels, weight)
        extra = {} if weight is None else {"weight": np.asarray(weight, dtype=np.float64)}
        np.savez_compressed(os.path.join(OUT, f"{name}.npz"), kind="sup", logits=lg.numpy(), labels=labels.numpy(),
                            **extra, **res)
        print(name, res["loss_f64"], res["loss_f32"], res["dice"])


def make_flip(ns):
    """Flip-alignment fixtures: the reference's seeded per-sample loop (semi_seg/epocher.py:148-149,160-161) and
    its UDA term on the flipped teacher logits (epocher.py:221-224)."""
    torch = ns.torch
    TensorRandomFlip, FixRandomSeed = ref_loader.load_flip()
    T = TensorRandomFlip(axis=[1, 2], threshold=0.8)          # semi_seg/epocher.py:121
    rng = np.random.default_rng(20260120)
    for name, shape, seed, kind in (("f_8x4x6x8_mse", (8, 4, 6, 8), 3, "mse"),        # 16-byte rows
                                    ("f_6x4x5x7_kl", (6, 4, 5, 7), 11, "kl"),         # ragged rows
                                    ("f_5x2x12x12_kl", (5, 2, 12, 12), 12345, "kl"),
                                    ("f_7x8x4x4_mse", (7, 8, 4, 4), 0, "mse")):
        student = torch.from_numpy((rng.standard_normal(shape) * 2).astype(np.float32))
        teacher = torch.from_numpy((rng.standard_normal(shape) * 2).astype(np.float32))
        with FixRandomSeed(seed):
            teacher_tf = torch.stack([T(x) for x in teacher], dim=0)
        with FixRandomSeed(seed):                              # the draws themselves, for the flag fixture
            import random as _random
            draws = [[_random.random() < 0.8 for _ in range(2)] for _ in range(shape[0])]
        flags = np.asarray([(1 if h else 0) | (2 if w else 0) for h, w in draws], dtype=np.uint8)
        out = {}
        for tag, dt in (("f64", torch.float64), ("f32", torch.float32)):
            s = student.to(dt).clone().requires_grad_(True)
            crit = torch.nn.MSELoss() if kind == "mse" else ns.KL_div(verbose=False)
            loss = crit(s.softmax(1), teacher_tf.to(dt).softmax(1).detach())
            (g,) = torch.autograd.grad(loss, (s,))
            out[f"loss_{tag}"] = loss.item()
            if tag == "f64":
                out["g_f64"] = g.numpy()
            else:
                out["gerr_f32"] = _relmax(g.numpy(), out["g_f64"])
        np.savez_compressed(os.path.join(OUT, f"{name}.npz"), kind=kind, seed=seed, student=student.numpy(),
                            teacher=teacher.numpy(), teacher_tf=teacher_tf.numpy(), flags=flags, **out)
        print(name, flags.tolist(), out["loss_f64"], out["loss_f32"])


def main():
    ns = ref_loader.load()
    torch = ns.torch
    torch.set_num_threads(4)
    os.makedirs(OUT, exist_ok=True)
    if "--sup-only" in sys.argv:
        make_sup(ns)
        return
    if "--flip-only" in sys.argv:
        make_flip(ns)
        return
    rng = np.random.default_rng(20260118)

    # ---- global IIDLoss ---------------------------------------------------------------
    for name, N, K, lamb, corr in (("g_n8_k5", 8, 5, 1.0, True), ("g_n32_k10", 32, 10, 1.0, True),
                                   ("g_n64_k20_l15", 64, 20, 1.5, True),
                                   ("g_n10_k20_indep", 10, 20, 1.0, False),
                                   ("g_n300_k7", 300, 7, 1.0, True)):
        if corr:
            base = rng.standard_normal((N, K)) * 2.0
            l1 = base + 0.7 * rng.standard_normal((N, K))
            l2 = base + 0.7 * rng.standard_normal((N, K))
        else:
            l1 = rng.standard_normal((N, K)) * 2.0
            l2 = rng.standard_normal((N, K)) * 2.0
        x = _softmax32(torch, l1.astype(np.float32))
        y = _softmax32(torch, l2.astype(np.float32))
        res = run_global(ns, x, y, lamb)
        np.savez_compressed(os.path.join(OUT, f"{name}.npz"), kind="global", x=x.numpy(), y=y.numpy(),
                            lamb=lamb, **res)
        print(name, res["loss_f64"], res["loss_f32"])

    # ---- local IIDSegmentationLoss / SmallPathLoss -------------------------------------
    local_cases = (
        # name, B, K, H, W, pad, lamda, mask, patch, correlated
        ("l_b2_k3_12x10_p1", 2, 3, 12, 10, 1, 1.0, False, None, True),
        ("l_b2_k4_9x11_p2", 2, 4, 9, 11, 2, 1.0, False, None, True),
        ("l_b1_k10_16x16_p0", 1, 10, 16, 16, 0, 1.0, False, None, True),
        ("l_b3_k5_20x24_p3", 3, 5, 20, 24, 3, 1.0, False, None, True),
        ("l_b2_k10_24x32_p1_l15", 2, 10, 24, 32, 1, 1.5, False, None, True),
        ("l_b2_k6_14x18_p1_mask", 2, 6, 14, 18, 1, 1.0, True, None, True),
        ("l_b2_k4_16x16_p1_indep", 2, 4, 16, 16, 1, 1.0, False, None, False),
        ("l_b1_k2_5x7_p7", 1, 2, 5, 7, 7, 1.0, False, None, True),       # window wider than the map
        ("sp_b2_k5_20x24_p1_patch8", 2, 5, 20, 24, 1, 1.0, False, 8, True),   # 4x5 ragged patches
        ("sp_b2_k4_24x24_p2_patch16_mask", 2, 4, 24, 24, 2, 1.0, True, 16, True),
        ("sp_b2_k10_28x28_p1_patch512", 2, 10, 28, 28, 1, 1.0, False, 512, True),  # one patch
    )
    for name, B, K, H, W, pad, lamda, use_mask, patch, corr in local_cases:
        if corr:
            l1, l2 = correlated_views(rng, B, K, H, W)
        else:
            l1 = (rng.standard_normal((B, K, H, W)) * 2).astype(np.float32)
            l2 = (rng.standard_normal((B, K, H, W)) * 2).astype(np.float32)
        x = _softmax32(torch, l1)
        y = _softmax32(torch, l2)
        mask = None
        if use_mask:
            mask = torch.from_numpy((rng.random((B, 1, H, W)) > 0.3).astype(np.float32))
        res = run_local(ns, x, y, pad, lamda, mask, patch)
        extra = {} if mask is None else {"mask": mask.numpy()}
        np.savez_compressed(os.path.join(OUT, f"{name}.npz"), kind="local", x=x.numpy(), y=y.numpy(),
                            padding=pad, lamda=lamda, patch_size=-1 if patch is None else patch,
                            **extra, **res)
        print(name, res["loss_f64"], res["loss_f32"])

    # ---- UDA ----------------------------------------------------------------------------
    for name, shape, kind, weight in (("u_mse_2x4x6x5", (2, 4, 6, 5), "mse", None),
                                      ("u_kl_2x4x6x5", (2, 4, 6, 5), "kl", None),
                                      ("u_kl_w_3x4x8x8", (3, 4, 8, 8), "kl", [1.0, 2.0, 0.5, 1.5]),
                                      ("u_mse_3x4x16x16", (3, 4, 16, 16), "mse", None),
                                      ("u_kl_4x4", (4, 4), "kl", None)):
        p = _softmax32(torch, (rng.standard_normal(shape) * 2).astype(np.float32))
        t = _softmax32(torch, (rng.standard_normal(shape) * 2).astype(np.float32))
        res = run_uda(ns, p, t, kind, weight)
        extra = {} if weight is None else {"weight": np.asarray(weight, dtype=np.float64)}
        np.savez_compressed(os.path.join(OUT, f"{name}.npz"), kind=kind, prob=p.numpy(), target=t.numpy(),
                            **extra, **res)
        print(name, res["loss_f64"], res["loss_f32"])

    # ---- patch_generator windows (iic_loss.py:152-160) -----------------------------------
    wins = {}
    for (h, w, ps) in ((100, 100, 32), (56, 56, 32), (224, 224, 512), (20, 24, 8), (33, 17, 16), (8, 8, 8)):
        fm = torch.zeros(1, 1, h, w)
        idx = torch.arange(h * w, dtype=torch.float32).reshape(1, 1, h, w)
        got = []
        for patch in ns.patch_generator(idx + fm, (ps, ps), (ps // 2, ps // 2)):
            first = int(patch[0, 0, 0, 0].item())
            got.append((first // w, first // w + patch.shape[2], first % w, first % w + patch.shape[3]))
        wins[f"{h}x{w}_p{ps}"] = np.asarray(got, dtype=np.int64)
    np.savez_compressed(os.path.join(OUT, "patch_windows.npz"), **wins)

    # ---- simplex boundary cases (dc2:utils/assertion.py:56-65) ---------------------------
    base = torch.full((2, 4, 3, 3), 0.25)
    cases, verdicts = [], []
    for delta in (0.0, 1.9e-4 / 4, 2.1e-4 / 4, -1.9e-4 / 4, -2.1e-4 / 4, 1e-3):
        t = base + delta
        cases.append(t.numpy())
        verdicts.append(bool(ns.simplex(t)))
    t = base.clone()
    t[0, 0, 0, 0] = float("nan")
    cases.append(t.numpy())
    verdicts.append(bool(ns.simplex(t)))
    np.savez_compressed(os.path.join(OUT, "simplex_cases.npz"), cases=np.stack(cases),
                        verdicts=np.asarray(verdicts))
    print("simplex verdicts", verdicts)

    make_sup(ns)
    make_flip(ns)


if __name__ == "__main__":
    main()
