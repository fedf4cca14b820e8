"""The arithmetic identities behind the tensor-core epilogue (api.cu fast_requant levels),
brute-forced on the CPU against the oracle's literal restatement of pe.cl:185-203 (tf2o_requant, itself pinned
against the compiled pe.cl).  For a channel with accumulator bound S = |bias| + 128 * sum_k 2^shift_k:

  level 1 (no int32 intermediate wraps):  |S * alpha| / 2^20 + |beta| + 2^14 + 3 < 2^31
      =>  requant(acc) = clamp((acc * alpha + ((beta + 2^14) << 20)) >> 35)
  level 2 (folded):  S < 2^31 - 1 and |alpha| << nshift < 2^31 - 1, acc = tot * 2^nshift + bias
      =>  acc * alpha + ((beta + 2^14) << 20) = tot * (alpha << nshift) + [bias * alpha + ((beta + 2^14) << 20)]
  level 3 (hi32):  every nshift >= 3
      =>  (t >> 35) = high 32 bits of tot * (alpha << (nshift - 3)) + (B >> 3)

Python integers are exact, so the right-hand sides are evaluated without any machine-width assumption; the
test also checks that every right-hand side fits the 64-bit registers the kernel computes it in."""
import numpy as np

from oracle import oracle as O


def _clamp(y):
    return max(-128, min(127, y))


def _cases(rng, n):
    for _ in range(n):
        nshift = int(rng.integers(0, 16))
        alpha = int(rng.integers(1, 1 << int(rng.integers(1, 31)))) * int(rng.choice([-1, 1]))
        beta = int(rng.integers(-(1 << 29), 1 << 29))
        bias = int(rng.integers(-(1 << 24), 1 << 24))
        tot_bound = int(rng.integers(1, 1 << int(rng.integers(1, 24))))        # |tot| <= 127 * 64 * K planes-combined
        yield nshift, alpha, beta, bias, tot_bound


def test_folded_and_hi32_requant_equal_the_reference_form():
    L = O.lib()
    rng = np.random.default_rng(2024)
    seen = {1: 0, 2: 0, 3: 0}
    for nshift, alpha, beta, bias, tot_bound in _cases(rng, 4000):
        S = abs(bias) + (tot_bound << nshift)                                    # bound of |acc| (api.cu:291-295)
        if S > 2 ** 31:
            continue
        level1 = S * abs(alpha) / 2 ** 20 + 1 + abs(beta) + 16384 + 2 < 2 ** 31 - 1
        if not level1:
            continue
        level2 = S < 2 ** 31 - 1 and (abs(alpha) << nshift) < 2 ** 31 - 1
        level3 = level2 and nshift >= 3
        tots = [0, 1, -1, tot_bound, -tot_bound, tot_bound - 1] + [int(v) for v in rng.integers(-tot_bound, tot_bound + 1, 24)]
        for tot in tots:
            acc = tot * (1 << nshift) + bias
            assert -2 ** 31 <= acc < 2 ** 31                                     # the accumulator did not wrap
            ref = int(L.tf2o_requant(acc, alpha, beta))
            B = bias * alpha + ((beta + (1 << 14)) << 20)
            t1 = acc * alpha + ((beta + (1 << 14)) << 20)
            assert -2 ** 63 <= t1 < 2 ** 63
            assert _clamp(t1 >> 35) == ref
            seen[1] += 1
            if level2:
                A = alpha << nshift
                t2 = tot * A + B
                assert t2 == t1 and -2 ** 31 <= A < 2 ** 31 and -2 ** 63 <= B < 2 ** 63
                seen[2] += 1
            if level3:
                A3, B3 = alpha << (nshift - 3), B >> 3
                t3 = tot * A3 + B3
                assert -2 ** 63 <= t3 < 2 ** 63 and _clamp(t3 >> 32) == ref
                seen[3] += 1
    assert min(seen.values()) > 2000, seen


def test_level1_condition_is_needed():
    """Outside the range analysis the int32 wrap of `a + beta` (pe.cl:191-194) is real: the folded form differs,
    which is why such layers keep the literal epilogue."""
    L = O.lib()
    acc, alpha, beta = 2 ** 30, 2 ** 21 - 1, 2 ** 30                             # a = 2^31 - 2^10, a + beta wraps
    ref = int(L.tf2o_requant(acc, alpha, beta))
    folded = _clamp((acc * alpha + ((beta + (1 << 14)) << 20)) >> 35)
    assert ref != folded


def test_weight_plane_factorisation_is_exact_modulo_2_32():
    """acc = bias + sum_k (+-x_k) << s_k in wrap-around int32 (pe.cl:27-40) for ANY split s_k = base + 7p + e:
    the epilogue's recombination (sum_p tot_p << 7p) << base + bias is the same number modulo 2^32."""
    rng = np.random.default_rng(7)
    for _ in range(200):
        K = int(rng.integers(1, 300))
        x = rng.integers(-128, 128, K)
        s = rng.integers(0, 24, K)
        sign = rng.choice([-1, 1], K)
        bias = int(rng.integers(-2 ** 31, 2 ** 31))
        ref = bias
        for xi, si, sg in zip(x, s, sign):
            ref = (ref + ((int(sg) * int(xi)) << int(si))) & 0xFFFFFFFF
        base = int(s.min())
        planes = {}
        for xi, si, sg in zip(x, s, sign):
            p, e = divmod(int(si) - base, 7)
            planes[p] = planes.get(p, 0) + int(sg) * int(xi) * (1 << e)
        tot = sum(v << (7 * p) for p, v in planes.items())
        assert ((tot << base) + bias) & 0xFFFFFFFF == ref
