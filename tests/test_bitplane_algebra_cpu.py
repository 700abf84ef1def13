"""Executable specification (numpy, CPU) of the integer arithmetic the bit-plane IMMA GEMV performs
(onebit_b200/csrc/imma_gemv.cuh): power-of-two quantisation of h*x, plane-scaled values, balanced base-256 digits as
plain bytes, int8 A values `bit << j` (-128 for plane 7), one accumulator per digit, exact recombination. Checked
against the pinned CPU oracle."""
import numpy as np

from oracle import oracle


def quantise(xp):
    amax = np.abs(xp).max(-1)
    e = np.where(amax > 0, np.frexp(amax)[1], 0)            # amax = f * 2^e, f in [0.5, 1)
    scale = np.ldexp(1.0, 22 - e)                            # |q| <= 2^22
    q = np.rint(xp * scale[:, None]).astype(np.int64)
    return q, scale


def digits_as_bytes(v):
    """bytes of (v + 0x00808080) ^ 0x00808080 read as int8 are the balanced digits of v"""
    u = ((v + 0x00808080) ^ 0x00808080).astype(np.int64) & 0xFFFFFFFF
    d = [((u >> (8 * i)) & 0xFF).astype(np.int64) for i in range(4)]
    return [np.where(x >= 128, x - 256, x) for x in d]


def test_bitplane_scheme_is_exact_and_matches_the_oracle():
    for seed, k, n, m in [(5, 4096, 64, 2), (6, 512, 33, 3), (7, 11008, 16, 1)]:
        case = oracle.synth_case(seed, k, n, m)
        x, packed, g, h = case["x"], case["packed"], case["g"], case["h"]
        xp = (x * h[None, :]).astype(np.float32)
        q, scale = quantise(xp)
        assert np.abs(q).max() <= 2 ** 22
        j = np.arange(k) % 8
        v = np.where(j == 7, -q, q << (7 - j)[None, :])
        ds = digits_as_bytes(v)
        assert all(np.abs(d).max() <= 128 for d in ds)
        assert (sum(ds[i] * 256 ** i for i in range(4)) == v).all()          # digits recombine to v
        bits = np.unpackbits(packed.view(np.uint8)[:, :, None], axis=-1, bitorder="little").reshape(n, -1).astype(np.int64)
        a = bits * np.where(j == 7, -128, 1 << j)[None, :]                    # int8 A fragments: w & (0x01010101 << j)
        acc = [a @ d.T for d in ds]                                           # one int32 accumulator per digit
        assert all(np.abs(c).max() < 2 ** 31 for c in acc)
        V = sum(acc[i] * 256 ** i for i in range(4))
        assert (V % 128 == 0).all() and (V // 128 == bits @ q.T).all()        # = 128 * sum_{bit=1} q, exactly
        t = (q.sum(-1)[None, :] - 2 * (V // 128)) / scale[None, :]            # sum_k s*q = sum q - 2 sum_{bit=1} q
        _, u = oracle.bitlinear_forward_c(x, packed, g, h, return_pre_ln=True)
        assert oracle.rel_l2((t * g[:, None]).T, u) < 2e-6                    # 23-bit quantisation: far below fp16


def test_prefill_sign_trick_puts_two_bits_on_two_fp16_sign_positions():
    # prefill_tc5.cu: (byte * (0x40008000 >> 2i)) & 0x80008000 -> bit 2i on bit 15, bit 2i+1 on bit 31, no carries
    for byte in range(256):
        for i in range(4):
            got = (byte * (0x40008000 >> (2 * i))) & 0x80008000
            want = (((byte >> (2 * i)) & 1) << 15) | (((byte >> (2 * i + 1)) & 1) << 31)
            assert got == want, (byte, i)
