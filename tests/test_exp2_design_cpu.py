"""CPU checks of the arithmetic the f64 pair kernel's exp2 relies on (pybnesian_b200/csrc/pair_kernel.cuh), restated in
numpy with the constants read from the header: the accuracy of the table + polynomial split, the hoisted test-row norm of
the dot-product form (tile_f64_dot) and the exponent floor (pair_floor).  No GPU, no oracle: these pin the error bounds
DESIGN.md quotes, so that a change of PBN_EXP_BITS / PBN_EXP_DEG / PBN_F64_FLOOR_BITS that breaks them fails here."""
import os
import re

import numpy as np

HDR = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "pybnesian_b200", "csrc", "pair_kernel.cuh")


def _define(name):
    m = re.search(r"#ifndef %s\s*\n#define %s\s+(\S+)" % (name, name), open(HDR).read())
    assert m, name
    return int(m.group(1))


BITS, DEG, FLOOR = _define("PBN_EXP_BITS"), _define("PBN_EXP_DEG"), _define("PBN_F64_FLOOR_BITS")
K = 1 << BITS
A = np.log(2.0) / K
H2 = 0.25 * A * A
C0 = 1.0 - H2 * H2 / 192.0 if DEG == 3 else 1.0
C1 = A * (1.0 + H2 / 8.0) if DEG == 2 else A
C2 = A * A * (0.5 + H2 / 24.0) if DEG == 3 else 0.5 * A * A
C3 = A ** 3 / 6.0
SQ = _define("PBN_EXP_SQ")
# completed-square form (PBN_EXP_SQ): e^(a g) ~ c2 ((g + S)^2 + Cq) with S = round(1/a) an integer, c2 = a / (2 S)
S = float(np.floor(1.0 / A + 0.5))
SQ_C2 = A / (2.0 * S)
SQ_D = -(SQ_C2 - 0.5 * A * A) / 12.0  # zero-mean error over |g| <= 1/2
SQ_C = (1.0 + SQ_D) / SQ_C2 - S * S
NMIN = -(1022 - 27 if SQ else 1022) * K


def poly(g):
    if SQ:
        gs = g + S  # the kernel gets g + S straight from the argument split: t - (rint(t) - S)
        return SQ_C2 * (gs * gs + SQ_C)
    p = C3 * g + C2 if DEG == 3 else np.full_like(g, C2)
    return (p * g + C1) * g + C0


def exp2_tab(t, nshift=0, nmin=NMIN):
    """2^((t + nshift)/K) the way exp2_tab computes it: n = rint(t) + nshift clamped to nmin, g = t - rint(t)."""
    nd = np.rint(t)
    n = np.maximum(nd.astype(np.int64) + nshift, nmin)
    return np.ldexp(np.exp2((n % K) / K), (n // K).astype(np.int64)) * poly(t - nd)


def test_polynomial_error_is_what_the_docs_say():
    g = np.linspace(-0.5, 0.5, 200001)
    err = np.max(np.abs(poly(g) / np.exp(A * g) - 1.0))
    assert DEG in (2, 3)
    if SQ:
        # |S a - 1| (a^2 / 2) g^2 + a^3 |g|^3 / 6: 2.7e-13 at K = 4096 (pair_kernel.cuh), forty times inside the 1e-11 budget
        assert K == 4096 and S == 5909.0 and err < 2.3e-13
        assert abs(np.mean(poly(g) / np.exp(A * g) - 1.0)) < 2e-15  # no bias in sums of many terms
    else:
        assert err < (3e-14 if DEG == 2 else 1e-15) * (4096.0 / K) ** (DEG + 1) + 3e-16  # 2.5e-14 at K = 4096, degree 2


def test_table_split_matches_exp2():
    rng = np.random.default_rng(0)
    t = -rng.uniform(0, 900 * K, 200000)
    got, want = exp2_tab(t), np.exp2(t / K)
    assert np.max(np.abs(got / want - 1.0)) < (3e-13 if SQ else 4e-14)


def test_hoisted_row_norm_is_an_integer_shift_times_a_row_factor():
    """tile_f64_dot: exponent = at + (b + sum 2 yt p).  at = A_int + f; A_int joins the rounded exponent, 2^(f/K) scales
    the finished sum: same value as evaluating the full exponent, term by term."""
    rng = np.random.default_rng(1)
    at = -rng.uniform(0, 3e5)
    rest = rng.uniform(-2e5, 3e5, 100000)  # b + dot product; at + rest <= 0 for a real pair, keep those
    rest = rest[at + rest <= 0]
    ai = np.rint(at)
    hoisted = exp2_tab(rest, nshift=int(ai)) * np.exp((at - ai) * A)
    direct = np.exp2((at + rest) / K)
    assert np.max(np.abs(hoisted / direct - 1.0)) < (3e-13 if SQ else 5e-14)
    assert abs(hoisted.sum() / direct.sum() - 1.0) < 1e-14


def _floor(s):
    if FLOOR == 0 or s <= 0.0:
        return NMIN
    e = int(np.floor(np.log2(s)))  # exponent field of the running sum
    return max((e - FLOOR) * K, NMIN)


def _floored_sum(t, tile):
    """Sum of 2^(t_i/K) in tiles of `tile` terms with the floor refreshed from the running sum before each tile."""
    s = 0.0
    for i in range(0, len(t), tile):
        s += float(exp2_tab(t[i:i + tile], nmin=_floor(s)).sum())
    return s


def test_exponent_floor_changes_nothing_visible():
    rng = np.random.default_rng(2)
    n, tile = 200_000, 512
    near = -rng.uniform(0, 40 * K, n // 20)          # the terms that make up the sum
    far = -rng.uniform(60 * K, 1000 * K, n - len(near))  # the bulk: far pairs
    for order in ("near_first", "far_first", "shuffled"):
        t = np.concatenate([near, far] if order == "near_first" else [far, near])
        if order == "shuffled":
            t = rng.permutation(t)
        exact = float(np.exp2(t / K).sum())
        got = _floored_sum(t, tile)
        plain = _floored_sum(t, len(t))  # one tile: floor never above kNMin
        assert abs(got / exact - 1.0) < 1e-13
        assert abs(got - plain) <= n * 2.0 ** -FLOOR * exact + 4 * np.finfo(float).eps * exact
    # every floored lane reads table entry 0: the floor is a multiple of K
    assert _floor(3.7) % K == 0 and _floor(1e-200) % K == 0
    # a tiny row (all terms ~ 2^-600) keeps its relative accuracy: the floor follows the running sum down
    tiny = -rng.uniform(600 * K, 640 * K, 50_000)
    assert abs(_floored_sum(tiny, tile) / float(np.exp2(tiny / K).sum()) - 1.0) < 1e-13


def test_f32_mufu_offload_polynomial_and_exponent_insertion():
    """ex2_neg_soft2 (pair_kernel.cuh): 2^(-s) on the FP32 pipe for the share of the exponentials the MUFU offload takes - the
    argument split through the mantissa of -s + 1.5 2^23, the degree-4 polynomial in g = -f (coefficients read from the
    header) and the exponent added to the bit pattern, restated in float32 numpy.  3.5e-6 relative, mean ~0: 30 times inside
    the float32 bar of 1e-4, and sums of many terms carry no bias."""
    src = open(HDR).read()
    body = src[src.index("f32x2_t ex2_neg_soft2"):]
    body = body[:body.index("template <int D, bool CKDE, int R>")]
    coef = [np.float32(x) for x in re.findall(r"pack_f32x2\((-?[0-9.]+)f, -?[0-9.]+f\)", body)]
    # order in the source: clamp 125 is a fminf, then MAGIC, -MAGIC, -1, c4, -c3, c2, -c1, 1
    magic, nmagic, neg1, c4, c3, c2, c1, one = coef
    assert magic == np.float32(12582912.0) and nmagic == -magic and neg1 == -1 and one == 1
    rng = np.random.default_rng(5)
    s = np.concatenate([rng.uniform(0, 125, 400000), np.linspace(0, 4, 100001)]).astype(np.float32)
    s = np.minimum(s, np.float32(125.0))
    r = (s * neg1 + magic).astype(np.float32)
    g = (s + (r + nmagic).astype(np.float32)).astype(np.float32)
    assert np.all(np.abs(g) <= 0.5)
    p = (g * c4 + c3).astype(np.float32)
    for c in (c2, c1, one):
        p = (p * g + c).astype(np.float32)
    bits = p.view(np.int32) + (r.view(np.int32) << 23)          # int32 wrap-around is the point: only n << 23 survives
    got = bits.view(np.float32).astype(np.float64)
    want = np.exp2(-s.astype(np.float64))
    ok = want > 2.0 ** -124                                      # below: flushed / denormal, negligible next to any sum
    err = got[ok] / want[ok] - 1.0
    assert np.max(np.abs(err)) < 4e-6 and abs(err.mean()) < 2e-7


def test_group_skip_levels_bound_what_is_dropped():
    """tile_f64_dot_gskip / tile_f32_packed_gskip (pair_kernel.cuh): a group of training points is skipped when every rounded
    exponent is at or below K (exponent(sum) - kSkipBits) (f64) / every squared distance at or above kSkipBits32 -
    floor(log2 sum) (f32).  Restated with the constants of the header: every skipped term is then below 2^-bits of the
    running sum, so all skipped terms of a row stay below N 2^-bits of it - 6e-14 (f64) / 6e-8 (f32) at a million rows."""
    bits64, bits32 = _define("PBN_F64_GSKIP_BITS"), _define("PBN_F32_GSKIP_BITS")
    assert bits64 >= 60 and bits32 >= 40            # N = 2^20 rows: 2^-40 (1e-12) of the sum in f64, 2^-20 (1e-6) in f32
    rng = np.random.default_rng(7)
    for s in np.exp2(rng.uniform(-900, 60, 2000)):
        e = int(np.floor(np.log2(s)))                # exponent field of the running sum minus the bias
        level = max((e - bits64) * K, NMIN)          # pair_skip_level
        n = level                                    # the largest rounded exponent a skipped term can have
        term_max = np.exp2((n + 0.5) / K)            # g <= 1/2 on top of the rounded exponent
        if level > NMIN:
            assert term_max <= s * 2.0 ** -(bits64 - 1)
        thr32 = float(bits32 - e)                    # pair_skip_level_f32: bits + 1023 - biased exponent
        assert np.exp2(-thr32) <= s * 2.0 ** -bits32
    # a CKDE tests the marginal exponent against the LOWER of the two levels (f64) / the HIGHER threshold (f32): the joint
    # exponent is the marginal one minus a square, so it passes whenever the marginal one does
    sj, sm = 3.0e-7, 0.02
    lj, lm = (int(np.floor(np.log2(sj))) - bits64) * K, (int(np.floor(np.log2(sm))) - bits64) * K
    assert min(lj, lm) == lj and max(bits32 - np.floor(np.log2(sj)), bits32 - np.floor(np.log2(sm))) == bits32 - np.floor(np.log2(sj))
