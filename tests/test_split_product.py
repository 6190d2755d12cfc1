"""Host-side model of the arithmetic the tensor-core kernels use for an fp32-accurate product
(csrc/local_fwd_tc.cu, local_bwd_tc.cu, local_bwd_tcrb*.cu, local_fwd_tcp.cu):

    a*w = ah*wh + (al*wh + ah*wl) + al*wl,   ah = top 19 bits of a (what kind::tf32 reads), al = a - ah (exact)

with the main term as a TF32 product, the two correction terms on bf16 (round-to-nearest) copies and al*wl dropped.
The test bounds the per-product error of that scheme in float64 -- it documents why two MMAs per product are enough
for the loss <= 1e-5 / gradient <= 1e-4 tolerances; the GPU parity tests check the kernels themselves."""
import numpy as np


def trunc19(v):
    u = np.asarray(v, dtype=np.float32).view(np.uint32) & np.uint32(0xFFFFE000)
    return u.view(np.float32)


def bf16_rn(v):
    u = np.asarray(v, dtype=np.float32).view(np.uint32).astype(np.uint64)
    u = (u + 0x7FFF + ((u >> 16) & 1)) & 0xFFFF0000          # round to nearest even on the upper 16 bits
    return u.astype(np.uint32).view(np.float32)


def split_product(a, w):
    a = a.astype(np.float32)
    w = w.astype(np.float32)
    ah, wh = trunc19(a), trunc19(w)
    al, wl = a - ah, w - wh                                     # exact in fp32
    main = ah.astype(np.float64) * wh.astype(np.float64)
    corr = bf16_rn(al).astype(np.float64) * bf16_rn(w).astype(np.float64) \
        + bf16_rn(a).astype(np.float64) * bf16_rn(wl).astype(np.float64)
    return main + corr


def test_remainder_is_exact_and_small():
    rng = np.random.default_rng(0)
    a = (rng.random(100000) ** 3 * 0.05).astype(np.float32)
    ah = trunc19(a)
    al = a - ah
    assert np.all(ah.astype(np.float64) + al.astype(np.float64) == a.astype(np.float64))
    assert np.all(np.abs(al) <= np.abs(a) * 2.0 ** -10)


def test_split_product_error_bound():
    rng = np.random.default_rng(1)
    a = (rng.random(200000) ** 3 * 0.05 + 1e-8).astype(np.float32)
    w = ((rng.random(200000) - 0.4) * 2).astype(np.float32)
    got = split_product(a, w)
    ref = a.astype(np.float64) * w.astype(np.float64)
    rel = np.abs(got - ref) / np.abs(ref)
    # dropped al*wl: 2^-20; bf16 rounding of both operands of both correction terms (each <= 2^-10 of the product):
    # worst case 2 * 2 * 2^-9 * 2^-10 = 2^-17 = 7.6e-6, round-to-nearest, so sums average it down (next test)
    assert rel.max() < 8e-6, rel.max()
    assert rel.mean() < 1.5e-6, rel.mean()
    # a plain TF32 product would be three orders of magnitude worse
    plain = np.abs(trunc19(a).astype(np.float64) * trunc19(w).astype(np.float64) - ref) / np.abs(ref)
    assert plain.max() > 1e-3


def test_sum_of_split_products_is_unbiased_enough():
    """A long sum of split products (what one joint entry is) keeps a relative error far below the loss tolerance."""
    rng = np.random.default_rng(2)
    x = rng.random((64, 4096)).astype(np.float32) ** 3
    y = rng.random((64, 4096)).astype(np.float32) ** 3
    got = split_product(x, y).sum(axis=1)
    ref = (x.astype(np.float64) * y.astype(np.float64)).sum(axis=1)
    assert np.max(np.abs(got - ref) / ref) < 2e-6
