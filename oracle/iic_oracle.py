"""CPU oracle for the IIC mutual-information hot path -- TEST INFRASTRUCTURE ONLY.

This file is a numpy/float64 restatement of the reference's algorithm.  It is the checker the CUDA
path is compared against.  Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s
``cpu_baseline`` / ``--impl reference`` legs may import it.  The product package never does.

Parity pin: the reference ships no tests and no golden vectors, so this
restatement is pinned against OUTPUTS OF THE REFERENCE ITSELF, run in this container by
``oracle/make_golden.py`` (which loads ``/root/reference/contrastyou/losses/iic_loss.py`` and the
``deepclustering2`` wheel by path) and committed under ``tests/golden/``.
``tests/test_oracle_golden.py`` checks every function below against those vectors.

Every function cites the reference lines it follows.  ``dc2:`` means a file inside
``/root/reference/deepclustering2-2.0.0-py3-none-any.whl``.

All functions take and return numpy arrays; everything is computed in float64 regardless of the
input dtype, which makes this the "fp64 run of the reference" that BASELINE.md section 2 names as
the ground truth for the ill-conditioned local term.
"""
from __future__ import annotations

import numpy as np

EPS_GLOBAL = 1e-10  # contrastyou/losses/iic_loss.py:64-68 (hard-coded, self.eps is unused)
EPS_LOCAL = 1e-16   # contrastyou/losses/iic_loss.py:124,141-143
EPS_KL = 1e-16      # dc2:deepclustering2/loss/kl_losses.py:89


def _f64(a):
    return np.asarray(a, dtype=np.float64)


# --------------------------------------------------------------------------------------------
# assertions
# --------------------------------------------------------------------------------------------
def simplex(t, axis: int = 1) -> bool:
    """dc2:deepclustering2/utils/assertion.py:56-65 -- allclose(sum over axis, 1, rtol=atol=1e-4).

    The reference casts the sum to float32 before comparing; ``allclose(a, b)`` is
    ``|a-b| <= atol + rtol*|b|`` with b == 1, i.e. ``|sum-1| <= 2e-4``, and NaN fails.
    """
    s = np.asarray(t).sum(axis=axis).astype(np.float32)
    return bool(np.all(np.abs(s - np.float32(1.0)) <= np.float32(1e-4) + np.float32(1e-4) * 1.0))


def softmax(logits, axis: int = 1, T: float = 1.0):
    """contrastyou/trainer/_utils.py:15-23 (SoftmaxWithT: divide by T, then softmax over dim)."""
    z = _f64(logits) / T
    z = z - z.max(axis=axis, keepdims=True)
    e = np.exp(z)
    return e / e.sum(axis=axis, keepdims=True)


def softmax_backward(prob, grad_prob, axis: int = 1, T: float = 1.0):
    """d(loss)/d(logits) for p = softmax(logits / T): p * (g - sum_k g_k p_k) / T."""
    p, g = _f64(prob), _f64(grad_prob)
    return p * (g - (g * p).sum(axis=axis, keepdims=True)) / T


# --------------------------------------------------------------------------------------------
# global IIC (IIDLoss / compute_joint)
# --------------------------------------------------------------------------------------------
def compute_joint(x_out, x_tf_out, symmetric: bool = True):
    """contrastyou/losses/iic_loss.py:74-94.

    J = sum_n x[n,:]^T y[n,:] (:88-89); optionally J <- (J + J^T)/2 (:91); P = J / sum(J) (:92).
    """
    x, y = _f64(x_out), _f64(x_tf_out)
    assert x.ndim == 2 and x.shape == y.shape
    J = x.T @ y
    if symmetric:
        J = (J + J.T) / 2.0
    return J / J.sum()


def _entropy_terms(P, lamb, eps):
    """-sum P (log(P+eps) - lamb log(pj+eps) - lamb log(pi+eps)) and d/dP of it.

    iic_loss.py:56-69 (global) and :135-146 (local, same expression with eps=1e-16).
    pi[i] = sum_j P[i,j] broadcast along j (:56-58 / :136); pj[j] = sum_i P[i,j] broadcast along i
    (:59 / :135).
    """
    pi = P.sum(axis=1, keepdims=True)
    pj = P.sum(axis=0, keepdims=True)
    loss = -(P * (np.log(P + eps) - lamb * np.log(pj + eps) - lamb * np.log(pi + eps))).sum()
    GP = (-np.log(P + eps) - P / (P + eps)
          + lamb * (np.log(pj + eps) + pj / (pj + eps))
          + lamb * (np.log(pi + eps) + pi / (pi + eps)))
    return loss, GP


def iid_loss(x_out, x_tf_out, lamb: float = 1.0):
    """contrastyou/losses/iic_loss.py:43-71 -- returns (loss, loss_no_lamb, p_i_j)."""
    P = compute_joint(x_out, x_tf_out)
    loss, _ = _entropy_terms(P, lamb, EPS_GLOBAL)
    loss_no_lamb, _ = _entropy_terms(P, 1.0, EPS_GLOBAL)
    return loss, loss_no_lamb, P


def iid_loss_grads(x_out, x_tf_out, lamb: float = 1.0, g_loss: float = 1.0,
                   g_loss_no_lamb: float = 0.0, g_P=None):
    """Analytic d/dx, d/dy of  g_loss*loss + g_loss_no_lamb*loss_no_lamb + <g_P, P>.

    Chain: P = Js/S, Js = (J+J^T)/2, J = x^T y (iic_loss.py:88-92).  With GP = dL/dP:
    GJs = (GP - sum(GP*P))/S ; GJ = (GJs + GJs^T)/2 ; dx = y GJ^T ; dy = x GJ  (SURVEY 8a, A3).
    """
    x, y = _f64(x_out), _f64(x_tf_out)
    J = x.T @ y
    Js = (J + J.T) / 2.0
    S = Js.sum()
    P = Js / S
    _, GP1 = _entropy_terms(P, lamb, EPS_GLOBAL)
    _, GP2 = _entropy_terms(P, 1.0, EPS_GLOBAL)
    GP = g_loss * GP1 + g_loss_no_lamb * GP2
    if g_P is not None:
        GP = GP + _f64(g_P)
    GJs = (GP - (GP * P).sum()) / S
    GJ = (GJs + GJs.T) / 2.0
    return y @ GJ.T, x @ GJ


# --------------------------------------------------------------------------------------------
# local IIC (IIDSegmentationLoss / IIDSegmentationSmallPathLoss)
# --------------------------------------------------------------------------------------------
def local_joint(x_out, x_tf_out, padding: int):
    """The F.conv2d at iic_loss.py:120-123, as explicit shifted-window sums.

    J[dy,dx,i,j] = sum_{n,u,v} x[n,i,u+dy-p,v+dx-p] * y[n,j,u,v], x zero outside the (patch) map.
    The reference's conv output is (K_i, K_j, T, T); :127 permutes it to (T, T, K_i, K_j), which is
    the layout returned here.
    """
    x, y = _f64(x_out), _f64(x_tf_out)
    B, K, H, W = x.shape
    p = int(padding)
    T = 2 * p + 1
    xp = np.zeros((B, K, H + 2 * p, W + 2 * p), dtype=np.float64)
    xp[:, :, p:p + H, p:p + W] = x
    J = np.empty((T, T, K, K), dtype=np.float64)
    for dy in range(T):
        for dx in range(T):
            J[dy, dx] = np.einsum("nihw,njhw->ij", xp[:, :, dy:dy + H, dx:dx + W], y, optimize=True)
    return J


def local_loss_from_joint(J, lamda: float = 1.0):
    """iic_loss.py:124-146 from the raw joint; returns (loss, GA = dL/dJ).

    m = min(J) detached (:124); A = J - m + 1e-16; per displacement Q = A / sum_ij A (:129);
    P = (Q + Q^T)/2 (:132); entropy expression / T^2 (:139-146).
    Backward (no gradient through m): GQ = (GP + GP^T)/2 ; GA = (GQ - sum(GQ*Q)) / s.
    """
    J = _f64(J)
    T = J.shape[0]
    m = J.min()
    A = J - m + 1e-16
    s = A.sum(axis=(2, 3), keepdims=True)
    Q = A / s
    P = (Q + np.swapaxes(Q, 2, 3)) / 2.0
    loss = 0.0
    GA = np.empty_like(J)
    for dy in range(T):
        for dx in range(T):
            l, GP = _entropy_terms(P[dy, dx], lamda, EPS_LOCAL)
            loss += l
            GQ = (GP + GP.T) / 2.0
            GA[dy, dx] = (GQ - (GQ * Q[dy, dx]).sum()) / s[dy, dx, 0, 0]
    return loss / (T * T), GA / (T * T)


def local_backward_from_GA(x_out, x_tf_out, GA, padding: int):
    """Adjoint of :func:`local_joint` (what autograd's convolution_backward computes for :123).

    dx[n,i,a,b] = sum_{d,j} GA[d,i,j] y[n,j,a-dy+p,b-dx+p];  dy[n,j,u,v] = sum_{d,i} GA[d,i,j] x[n,i,u+dy-p,v+dx-p].
    """
    x, y = _f64(x_out), _f64(x_tf_out)
    B, K, H, W = x.shape
    p = int(padding)
    T = 2 * p + 1
    xp = np.zeros((B, K, H + 2 * p, W + 2 * p), dtype=np.float64)
    xp[:, :, p:p + H, p:p + W] = x
    gxp = np.zeros_like(xp)
    gy = np.zeros_like(y)
    for dy in range(T):
        for dx in range(T):
            g = _f64(GA[dy, dx])
            gxp[:, :, dy:dy + H, dx:dx + W] += np.einsum("ij,njhw->nihw", g, y, optimize=True)
            gy += np.einsum("ij,nihw->njhw", g, xp[:, :, dy:dy + H, dx:dx + W], optimize=True)
    return gxp[:, :, p:p + H, p:p + W], gy


def iid_segmentation_loss(x_out, x_tf_out, padding: int, lamda: float = 1.0, mask=None,
                          with_grads: bool = False):
    """contrastyou/losses/iic_loss.py:107-149 (IIDSegmentationLoss.__call__).

    mask multiplies both maps first (:116-118); gradients then flow through the mask product.
    """
    x, y = _f64(x_out), _f64(x_tf_out)
    if mask is not None:
        m = _f64(mask)
        x, y = x * m, y * m
    J = local_joint(x, y, padding)
    loss, GA = local_loss_from_joint(J, lamda)
    if not with_grads:
        return loss
    gx, gy = local_backward_from_GA(x, y, GA, padding)
    if mask is not None:
        gx, gy = gx * m, gy * m
    return loss, gx, gy


def patch_windows(h: int, w: int, patch_size, step_size):
    """contrastyou/losses/iic_loss.py:152-160 -- the (h0, h1, w0, w1) windows patch_generator yields."""
    ph, pw = patch_size
    sh, sw = step_size
    hs = list(np.arange(0, h - ph, sh)) + [max(h - ph, 0)]
    ws = list(np.arange(0, w - pw, sw)) + [max(w - pw, 0)]
    return [(int(h0), int(min(h0 + ph, h)), int(w0), int(min(w0 + pw, w))) for h0 in hs for w0 in ws]


def iid_segmentation_small_path_loss(x_out, x_tf_out, padding: int, patch_size, lamda: float = 1.0,
                                     mask=None, with_grads: bool = False):
    """contrastyou/losses/iic_loss.py:164-186 -- mean over 50%-overlapping patches (step = patch//2,
    :169), each patch zero-padded at ITS OWN border, then ``average_iter`` (helper/utils.py:46-47).
    """
    x, y = _f64(x_out), _f64(x_tf_out)
    if np.isscalar(patch_size):
        patch_size = (int(patch_size), int(patch_size))
    step = (patch_size[0] // 2, patch_size[1] // 2)  # _pair(patch_size // 2) at :169
    wins = patch_windows(x.shape[2], x.shape[3], patch_size, step)
    total = 0.0
    gx = np.zeros_like(x)
    gy = np.zeros_like(y)
    for (h0, h1, w0, w1) in wins:
        mm = None if mask is None else _f64(mask)[:, :, h0:h1, w0:w1]
        r = iid_segmentation_loss(x[:, :, h0:h1, w0:w1], y[:, :, h0:h1, w0:w1], padding, lamda, mm,
                                  with_grads=with_grads)
        if with_grads:
            l, a, b = r
            gx[:, :, h0:h1, w0:w1] += a
            gy[:, :, h0:h1, w0:w1] += b
        else:
            l = r
        total += l
    n = float(len(wins))
    if with_grads:
        return total / n, gx / n, gy / n
    return total / n


# --------------------------------------------------------------------------------------------
# helper/utils.py averaging
# --------------------------------------------------------------------------------------------
def average_iter(a_list):
    """contrastyou/helper/utils.py:46-47."""
    return sum(a_list) / float(len(a_list))


def weighted_average_iter(a_list, weight_list):
    """contrastyou/helper/utils.py:54-56 -- sum(w*x) / (sum(w) + 1e-16)."""
    return sum(a * w for a, w in zip(a_list, weight_list)) / (sum(weight_list) + 1e-16)


# --------------------------------------------------------------------------------------------
# UDA consistency
# --------------------------------------------------------------------------------------------
def mse_loss(prob, target, with_grads: bool = False):
    """torch.nn.MSELoss() as built at semi_seg/trainer.py:137,194 -- mean over all elements."""
    p, t = _f64(prob), _f64(target)
    d = p - t
    loss = (d * d).mean()
    if with_grads:
        return loss, 2.0 * d / d.size
    return loss


def kl_div(prob, target, eps: float = EPS_KL, reduction: str = "mean", weight=None,
           with_grads: bool = False):
    """dc2:deepclustering2/loss/kl_losses.py:107-126 (KL_div.forward).

    kl = -target * log((prob+eps)/(target+eps)) (:115), optional per-class weight (:116-119; the
    ctor normalises it to sum to C, :100), sum over dim 1 (:120), then mean/sum/none (:121-126).
    Gradient flows to ``prob`` only (target is asserted grad-free, :112).
    """
    p, t = _f64(prob), _f64(target)
    kl = -t * np.log((p + eps) / (t + eps))
    wv = None
    if weight is not None:
        # the ctor normalises in float32: torch.Tensor(weight).float() / sum * len (:97-100)
        w32 = np.asarray(weight, dtype=np.float32)
        wv = _f64((w32 / w32.sum() * np.float32(len(w32))).astype(np.float32))
        shape = [1] * p.ndim
        shape[1] = -1
        wv = wv.reshape(shape)
        kl = kl * wv
    per = kl.sum(axis=1)
    g = -t / (p + eps)
    if wv is not None:
        g = g * wv
    if reduction == "mean":
        out, g = per.mean(), g / per.size
    elif reduction == "sum":
        out = per.sum()
    else:
        out = per
    if with_grads:
        return out, g
    return out


# --------------------------------------------------------------------------------------------
# supervised branch
# --------------------------------------------------------------------------------------------
def class2one_hot(seg, C: int):
    """dc2:deepclustering2/utils/assertion.py:101-116 -- integer one-hot over a new axis 1; labels
    outside range(C) fail the reference's ``assert sset(seg, list(range(C)))``."""
    seg = np.asarray(seg)
    if seg.ndim == 2:
        seg = seg[None]
    assert set(np.unique(seg).tolist()) <= set(range(C)), "a label lies outside [0, C)"
    return np.stack([seg == c for c in range(C)], axis=1).astype(np.int64)


def sup_kl_from_logits(logits, labels, eps: float = EPS_KL, weight=None, with_grads: bool = False):
    """semi_seg/epocher.py:165-166: ``KL_div()(label_logits.softmax(1), class2one_hot(target, C))``.

    Composition of :func:`softmax`, :func:`class2one_hot` and :func:`kl_div`; the gradient is taken back
    through the softmax to the logits."""
    p = softmax(logits, axis=1)
    t = class2one_hot(labels, p.shape[1])
    if not with_grads:
        return kl_div(p, t, eps=eps, weight=weight)
    loss, gp = kl_div(p, t, eps=eps, weight=weight, with_grads=True)
    return loss, softmax_backward(p, gp, axis=1)


def dice_counts(logits, labels):
    """What ``UniversalDice.add(logits.max(1)[1], labels)`` appends (semi_seg/epocher.py:183-184;
    dc2:meters2/individual_meters/general_dice_meter.py:41-95,142-172): per sample and class,
    intersection = sum(pred * target) and union = sum(pred + target) of the two one-hot maps."""
    z = np.asarray(logits)
    C = z.shape[1]
    pred = class2one_hot(z.argmax(axis=1), C)       # first maximal index, like torch.max
    tgt = class2one_hot(labels, C)
    axes = tuple(range(2, z.ndim))
    return (pred * tgt).sum(axis=axes), (pred + tgt).sum(axis=axes)


# --------------------------------------------------------------------------------------------
# per-sample flip alignment
# --------------------------------------------------------------------------------------------
FLIP_H, FLIP_W = 1, 2     # axis 1 / axis 2 of a (C, H, W) sample


def draw_flip_flags(seed: int, batch: int, axis=(1, 2), threshold: float = 0.8):
    """The draws of ``[TensorRandomFlip(axis, threshold)(x) for x in batch]`` under ``FixRandomSeed(seed)``
    (semi_seg/epocher.py:121,148-149; dc2:augment/tensor_augment.py:31-41; dc2:decorator/decorator.py:196-212):
    ``random.seed(seed)``, then per sample one ``random.random() < threshold`` per axis, in axis order."""
    import random
    rng = random.Random()
    rng.seed(seed)
    out = np.zeros(batch, dtype=np.uint8)
    for n in range(batch):
        for a in axis:
            if rng.random() < threshold:
                out[n] ^= FLIP_H if a == 1 else FLIP_W
    return out


def flip_stack(batch, flags):
    """``torch.stack([T(x) for x in batch])`` for the flips in ``flags`` (one (C, H, W) sample at a time)."""
    x = np.asarray(batch)
    out = np.empty_like(x)
    for n, f in enumerate(np.asarray(flags)):
        s = x[n]
        if f & FLIP_H:
            s = s[:, ::-1, :]
        if f & FLIP_W:
            s = s[:, :, ::-1]
        out[n] = s
    return out


def uda_from_logits_flipped(student_logits, teacher_logits, flags, kind: str = "mse", with_grads: bool = False):
    """semi_seg/epocher.py:160-161 + 221-224: criterion(softmax(student), softmax(flip_stack(teacher)).detach());
    the gradient goes back through the student's softmax."""
    p, t = softmax(student_logits, axis=1), softmax(flip_stack(teacher_logits, flags), axis=1)
    fn = mse_loss if kind == "mse" else kl_div
    if not with_grads:
        return fn(p, t)
    loss, gp = fn(p, t, with_grads=True)
    return loss, softmax_backward(p, gp, axis=1)
