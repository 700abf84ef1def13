"""Executable specification (numpy, CPU) of the integer arithmetic and data layout of the persistent decode step
(onebit_b200/csrc/persist_step.cu): static power-of-two quantisation, plane-scaled values, balanced base-255 digits
(bytes in [-127, 127], so 0x80 can mark "not written yet"), the B-fragment word layout the producers publish, the byte-sum
identity the consumers use for Q128 = 128 * sum q, the warp-level m16n8k32 fragment mapping, and the exact recombination.
Checked against the pinned CPU oracle."""
import numpy as np

from oracle import oracle

C255 = 127 * (1 + 255 + 255 ** 2 + 255 ** 3)


def digits255(v):
    """mirror of persist::digits255: u = v + C, divisions by 255 through the 0x80808081 multiply-high."""
    u = (v.astype(np.int64) + C255) & 0xFFFFFFFF
    assert (u == v.astype(np.int64) + C255).all()  # no 32-bit wrap for |v| <= 2^29
    q1 = ((u * 0x80808081) >> 32) >> 7
    assert (q1 == u // 255).all()
    e0 = u - q1 * 255
    q2 = q1 // 255
    e1 = q1 - q2 * 255
    q3 = q2 // 255
    e2 = q2 - q3 * 255
    return [e0 - 127, e1 - 127, e2 - 127, q3 - 127]


def frag_word(col0, j, d):
    u, W = col0 >> 8, (col0 & 255) >> 5
    return u * 256 + (j >> 1) * 64 + d * 16 + (W >> 1) * 4 + (j & 1) * 2 + (W & 1)


def publish(q, k):
    """digit vector [k] words (uint32) exactly as publish32 lays it out, from the integers q[k]"""
    cols = np.arange(k)
    j = cols & 7
    v = np.where(j == 7, -q, q << (7 - j))
    ds = digits255(v)
    assert all(np.abs(d).max() <= 127 for d in ds)
    assert (sum(ds[i] * 255 ** i for i in range(4)) == v).all()
    words = np.zeros(k, dtype=np.uint32)
    for col0 in range(0, k, 32):
        for jj in range(8):
            for d in range(4):
                w = 0
                for b in range(4):
                    w |= (int(ds[d][col0 + 8 * b + jj]) & 0xFF) << (8 * b)
                words[frag_word(col0, jj, d)] = w
    return words, ds


def test_digit_words_never_contain_the_sentinel_byte_and_q128_identity():
    rng = np.random.default_rng(0)
    for k in (256, 4096, 11008):
        q = rng.integers(-(1 << 22), (1 << 22) + 1, size=k)
        q[:4] = [1 << 22, -(1 << 22), 0, -1]
        words, _ = publish(q, k)
        b = words.view(np.uint8)
        assert (b != 0x80).all()                       # 0x80808080 stays a valid "not written" marker at byte granularity
        # consumer side: chunk i = 4 words; thread pattern (jp, d) = ((i >> 4) & 3, (i >> 2) & 3); x,y -> even plane, z,w -> odd
        sb = b.view(np.int8).astype(np.int64).reshape(-1, 4, 4).sum(-1)   # [chunk][word] byte sums (dp4a with 0x01010101)
        i = np.arange(k // 4)
        jp, d = (i >> 4) & 3, (i >> 2) & 3
        ce = 1 << (2 * jp)
        co = np.where(jp == 3, -128, 1 << (2 * jp + 1))
        q128 = (255 ** d * ((sb[:, 0] + sb[:, 1]) * ce + (sb[:, 2] + sb[:, 3]) * co)).sum()
        assert q128 == 128 * q.sum()


def test_fragment_mapping_and_exact_recombination_match_the_oracle():
    """Emulates mma.sync.m16n8k32.s8 with the register contents the kernel builds: A regs = weight words & plane mask,
    B regs = the published digit words, and checks sum_d 255^d acc_d == 128 * sum_{bit=1} q and the final t."""
    for seed, k, n in [(5, 512, 32), (7, 4096, 16)]:
        case = oracle.synth_case(seed, k, n, 1)
        x, packed, g, h = case["x"], case["packed"], case["g"], case["h"]
        xp = (x[0] * h).astype(np.float32)
        # static bound (persist_create): any e with |x'| < 2^e works; take a loose one like sqrt(K) * max|h| would give
        e = int(np.frexp(np.abs(xp).max() * 8.0)[1])
        q = np.clip(np.rint(xp * np.float32(2.0) ** (22 - e)), -(1 << 22), 1 << 22).astype(np.int64)
        words, _ = publish(q, k)
        wwords = packed.view(np.uint8).reshape(n, k // 32, 4).astype(np.uint32)
        wwords = (wwords[..., 0] | (wwords[..., 1] << 8) | (wwords[..., 2] << 16) | (wwords[..., 3] << 24)).astype(np.uint32)
        acc = np.zeros((n, 4), dtype=np.int64)  # [row][digit]
        for u in range(k // 256):
            for j in range(8):
                mask = np.uint32(0x01010101 << j)
                A = np.zeros((n, 32), dtype=np.int64)   # [row][k_mma]
                B = np.zeros((32, 4), dtype=np.int64)   # [k_mma][digit]
                for t in range(4):
                    for half in range(2):               # a0/a2 <-> word 2t + half of the unit; k_mma = 16*half + 4t + b
                        aw = wwords[:, u * 8 + 2 * t + half] & mask
                        for b in range(4):
                            A[:, 16 * half + 4 * t + b] = ((aw >> (8 * b)) & 0xFF).astype(np.uint8).view(np.int8)
                        for d in range(4):              # lane (g' = d, t) loads uint4 at u*256 + jp*64 + d*16 + t*4
                            wd = int(words[u * 256 + (j >> 1) * 64 + d * 16 + t * 4 + (j & 1) * 2 + half])
                            for b in range(4):
                                B[16 * half + 4 * t + b, d] = np.int8(np.uint8((wd >> (8 * b)) & 0xFF))
                acc += A @ B
        assert np.abs(acc).max() < 2 ** 31
        V = sum(acc[:, d] * 255 ** d for d in range(4))
        bits = np.unpackbits(packed.view(np.uint8)[:, :, None], axis=-1, bitorder="little").reshape(n, -1).astype(np.int64)
        assert (V == 128 * (bits @ q)).all()
        t = (128 * q.sum() - 2 * V).astype(np.float64) * 2.0 ** (e - 29)
        _, u_ref = oracle.bitlinear_forward_c(x, packed, g, h, return_pre_ln=True)
        assert oracle.rel_l2((t * g)[None, :], u_ref) < 2e-5   # static scale lost 3 of the 23 bits here: still ~1e-6


def test_residual_rmsnorm_sum_of_squares_identity():
    # residual_stage: sum (r + (u - mu) rs)^2 from the five exchanged sums
    rng = np.random.default_rng(3)
    r, u = rng.normal(size=4096), rng.normal(size=4096) * 3 + 0.5
    mu, var = u.mean(), u.var()
    rs = 1 / np.sqrt(var + 1e-5)
    direct = ((r + (u - mu) * rs) ** 2).sum()
    N = 4096
    su, suu, sr, srr, sru = u.sum(), (u * u).sum(), r.sum(), (r * r).sum(), (r * u).sum()
    alg = srr + 2 * rs * (sru - mu * sr) + rs * rs * (suu - 2 * mu * su + N * mu * mu)
    assert abs(alg - direct) / direct < 1e-12


def test_silu_bound_covers_the_true_maximum():
    # stage D1: |silu(g^) u^ h| <= rg ru (max|g u h| + |mg| max|u h| + |mu| max|g h| + |mg mu| max|h|)
    rng = np.random.default_rng(4)
    for _ in range(20):
        g, u = rng.normal(size=11008) * rng.uniform(0.1, 10) + rng.normal(), rng.normal(size=11008) * rng.uniform(0.1, 10) + rng.normal()
        h = rng.uniform(-1.5, 1.5, size=11008)
        mg, mu, rg, ru = g.mean(), u.mean(), 1 / np.sqrt(g.var() + 1e-5), 1 / np.sqrt(u.var() + 1e-5)
        gh, uh = (g - mg) * rg, (u - mu) * ru
        true = np.abs(gh / (1 + np.exp(-gh)) * uh * h).max()
        bound = rg * ru * (np.abs(g * u * h).max() + abs(mg) * np.abs(u * h).max() + abs(mu) * np.abs(g * h).max()
                           + abs(mg * mu) * np.abs(h).max())
        assert true <= bound
        assert bound / true < 64  # at most 6 of the 23 bits of headroom lost
