"""Executable specification (numpy, CPU) of what the second-generation fused decode stage
(onebit_b200/csrc/fused_gemv2.cuh) takes from its producer's records instead of from the vector itself:

  * item order: element 4*it + b  <->  column 32*(it/8) + it%8 + 8*b  (a bijection; the four columns of one quantiser item
    are the columns 8b + j of one 32-bit weight word);
  * RMSNorm denominator of r + LayerNorm(t) (modeling_bitllama.py:67-81 around bitnet.py:118) from per-CTA partial sums;
  * a bound on max |x'| from per-CTA (max, min) records: never below the true maximum, power-of-two scale from its
    exponent field, 23-bit integers that cannot saturate;
  * round-to-nearest-even through the 1.5 * 2^23 constant, exponent through the bit field instead of frexp.
"""
import numpy as np


def item_col(it):
    return ((it >> 3) << 5) + (it & 7)


def perm_index(c):
    return ((((c >> 5) << 3) + (c & 7)) << 2) + ((c >> 3) & 3)


def test_item_order_is_a_bijection_and_groups_the_columns_of_one_weight_word():
    for k in (256, 4096, 11008, 13824):
        cols = np.arange(k)
        idx = perm_index(cols)
        assert sorted(idx.tolist()) == list(range(k))
        it, b = idx >> 2, idx & 3
        assert (item_col(it) + 8 * b == cols).all()                     # the kernel's consumer-side view of the same map
        j = it & 7
        assert (cols % 8 == j).all()                                    # an item = plane j of one 32-bit weight word
        assert ((cols // 32) == (it >> 3)).all() and ((cols % 32) // 8 == b).all()


def _records(t, r, rows_per_cta):
    """what every producer CTA writes for its rows: base (sum t, sum t^2, max t, min t), resid (sum rt, sum r, sum r^2, max|r|)"""
    base, resid = [], []
    for lo in range(0, len(t), rows_per_cta):
        tt, rr = t[lo:lo + rows_per_cta].astype(np.float32), r[lo:lo + rows_per_cta].astype(np.float32)
        base.append((tt.sum(dtype=np.float32), (tt * tt).sum(dtype=np.float32), tt.max(), tt.min()))
        resid.append(((rr * tt).sum(dtype=np.float32), rr.sum(dtype=np.float32), (rr * rr).sum(dtype=np.float32), np.abs(rr).max()))
    return np.array(base, np.float32), np.array(resid, np.float32)


def test_rmsnorm_denominator_and_quantiser_bound_from_the_records():
    rng = np.random.default_rng(0)
    for n, rows, offset, outlier in [(4096, 32, 0.0, 1.0), (4096, 96, 0.3, 60.0), (11008, 160, -2.0, 1.0), (5120, 192, 0.0, 200.0)]:
        t = (rng.standard_normal(n) * 3.0 + offset).astype(np.float32)    # g * S @ x' of the producer (o / down projection)
        r = rng.standard_normal(n).astype(np.float32)                      # residual stream
        r[[3, n // 2]] *= outlier                                          # massive-activation channels
        wh = (rng.standard_normal(n) * 0.5).astype(np.float32)             # RMSNorm weight * input_factor of the consumer
        base, resid = _records(t, r, rows)
        # consumer: fp32 butterfly sums, fp64 only for the combination (finish_ln_fast and the `ss` expression)
        s_t, s_tt = base[:, 0].sum(dtype=np.float32), base[:, 1].sum(dtype=np.float32)
        s_rt, s_r, s_rr = (resid[:, i].sum(dtype=np.float32) for i in range(3))
        mu = float(s_t) / n
        var = float(s_tt) / n - mu * mu
        mean, rstd = np.float32(mu), np.float32(1.0 / np.sqrt(np.float32(max(var, 0.0)) + np.float32(1e-5)))
        ss = float(s_rr) + 2.0 * float(rstd) * (float(s_rt) - float(mean) * float(s_r)) + \
            float(rstd) ** 2 * (float(s_tt) - 2.0 * float(mean) * float(s_t) + n * float(mean) ** 2)
        rr = np.float32(1.0 / np.sqrt(np.float32(max(ss, 0.0) / n) + np.float32(1e-6)))
        # reference order of operations: LayerNorm, residual add, RMSNorm, * weight * input_factor
        ln = (t.astype(np.float64) - t.astype(np.float64).mean()) / np.sqrt(t.astype(np.float64).var() + 1e-5)
        x = r.astype(np.float64) + ln
        rms = 1.0 / np.sqrt((x * x).mean() + 1e-6)
        xp = x * rms * wh
        assert abs(float(rr) - rms) / rms < 2e-6, (n, rows, float(rr), rms)
        # bound from (max t, min t, max |r|, max |wh|)
        dev = max(base[:, 2].max() - mean, mean - base[:, 3].min()) * rstd
        bound = np.float32((resid[:, 3].max() + dev) * rr * np.abs(wh).max()) * np.float32(1.0001)
        assert bound >= np.abs(xp).max(), (bound, np.abs(xp).max())
        # exponent by bit field == frexp; scale; integers within 23 bits
        e = int((np.float32(bound).view(np.uint32) >> 23) & 0xFF) - 126
        assert e == np.frexp(bound)[1]
        scale = np.float32(2.0) ** (22 - e)
        q = np.rint(xp * float(scale))
        assert np.abs(q).max() <= 2 ** 22
        # the bound costs low bits, not parity: relative error of the dequantised vector far below fp16
        assert np.linalg.norm(q / float(scale) - xp) / np.linalg.norm(xp) < 2e-5


def test_silu_bound_from_the_gate_and_up_records():
    rng = np.random.default_rng(1)
    n = 11008
    g, u = (rng.standard_normal(n) * 2 + 0.1).astype(np.float32), (rng.standard_normal(n) * 0.7).astype(np.float32)
    h = rng.standard_normal(n).astype(np.float32)
    lg, lu = (g - g.mean()) / np.sqrt(g.var() + 1e-5), (u - u.mean()) / np.sqrt(u.var() + 1e-5)
    xp = lg / (1.0 + np.exp(-lg)) * lu * h
    gmax = (g.max() - g.mean()) / np.sqrt(g.var() + 1e-5)
    dev_u = max(u.max() - u.mean(), u.mean() - u.min()) / np.sqrt(u.var() + 1e-5)
    bound = max(gmax, 0.2785) * dev_u * np.abs(h).max()                  # |silu(x)| <= max(x_max, 0.2785)
    assert bound >= np.abs(xp).max()
    xs = np.linspace(-20, 0, 20001)
    assert np.abs(xs / (1 + np.exp(-xs))).max() < 0.2785                 # the constant: max |silu| on the negative axis


def test_magic_constant_rounding_is_round_to_nearest_even():
    rng = np.random.default_rng(2)
    y = np.concatenate([rng.uniform(-2 ** 22, 2 ** 22, 200000), np.arange(-8, 8) + 0.5, [2.0 ** 22, -(2.0 ** 22)]]).astype(np.float32)
    magic = np.float32(12582912.0)                                       # 1.5 * 2^23
    q = (y + magic).view(np.int32) - np.int32(0x4B400000)
    assert (q == np.rint(y).astype(np.int32)).all()
