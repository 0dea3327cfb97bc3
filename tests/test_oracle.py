"""CPU tests: pin the oracle to the reference's own known-answer tests and fixtures
(tests/golden/kat.json, transcribed from src/n_to_bits.rs:408-470 and benches/bench_n_to_bits.rs:68-78),
then check every restated variant against the scalar pair and against the independent numpy restatement.
"""
import json
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import cn_oracle_np as onp  # noqa: E402
from _oracle import DECODERS, ENCODERS  # noqa: E402

with open(os.path.join(ROOT, "tests", "golden", "kat.json")) as f:
    GOLD = json.load(f)

SIZES = [0, 1, 3, 4, 5, 15, 16, 17, 31, 32, 33, 63, 64, 65, 127, 128, 129, 1000, 4096, 40000, (1 << 16) + 17]


def rand_seq(rng, n, alphabet=b"ACGTUacgtu"):
    return np.frombuffer(alphabet, dtype=np.uint8)[rng.integers(0, len(alphabet), size=n)]


@pytest.mark.parametrize("variant", ENCODERS)
def test_reference_kat_encode(oracle, variant):
    if variant != "lut" and not oracle.simd_ok:
        pytest.skip("host CPU lacks AVX2/BMI2/PCLMUL")
    for case in GOLD["encode"]:
        assert variant in case["variants"]
        got = oracle.n_to_bits(case["input"].encode(), variant)
        assert [str(int(x)) for x in got] == case["words"], (variant, case["variants"][variant])


@pytest.mark.parametrize("variant", DECODERS)
def test_reference_kat_decode(oracle, variant):
    if variant != "lut" and not oracle.simd_ok:
        pytest.skip("host CPU lacks AVX2/BMI2/PCLMUL")
    for case in GOLD["decode"]:
        words = np.array([int(w) for w in case["words"]], dtype=np.uint64)
        assert oracle.bits_to_n(words, case["len"], variant) == case["output"].encode()


def test_reference_bench_fixture(oracle):
    fx = GOLD["bench_fixture"]
    n = (fx["unit"] * fx["repeat"]).encode()
    words = oracle.n_to_bits(n, "lut")
    assert words.size == fx["nwords"] and all(int(w) == int(fx["every_word"]) for w in words)
    assert oracle.bits_to_n(words, len(n), "lut") == n


def test_numpy_restatement_kats():
    for case in GOLD["encode"]:
        assert [str(int(x)) for x in onp.n_to_bits_lut(case["input"].encode())] == case["words"]
    for case in GOLD["decode"]:
        words = np.array([int(w) for w in case["words"]], dtype=np.uint64)
        assert onp.bits_to_n_lut(words, case["len"]) == case["output"].encode()
    for case in GOLD["encode2"]:
        assert [str(int(x)) for x in onp.n_to_bits2_lut(case["input"].encode())] == case["words"]
    for case in GOLD["decode2"]:
        words = np.array([int(w) for w in case["words"]], dtype=np.uint64)
        assert onp.bits_to_n2_lut(words, case["len"]) == case["output"].encode()


@pytest.mark.parametrize("size", SIZES)
def test_c_oracle_matches_numpy_restatement(oracle, size):
    rng = np.random.default_rng(size + 1)
    n = rand_seq(rng, size)
    words = oracle.n_to_bits(n, "lut")
    assert np.array_equal(words, onp.n_to_bits_lut(n))
    assert oracle.bits_to_n(words, size, "lut") == onp.bits_to_n_lut(words, size) == onp.canonical(n.tobytes())
    assert oracle.canonical(n) == onp.canonical(n.tobytes())


@pytest.mark.parametrize("size", SIZES)
def test_all_encoder_variants_agree_on_valid_alphabet(oracle, size):
    if not oracle.simd_ok:
        pytest.skip("host CPU lacks AVX2/BMI2/PCLMUL")
    rng = np.random.default_rng(size + 7)
    n = rand_seq(rng, size)
    ref = oracle.n_to_bits(n, "lut")
    for v in ENCODERS[1:]:
        assert np.array_equal(oracle.n_to_bits(n, v), ref), v
    # unaligned source slice (Rust &[u8] can start anywhere)
    if size > 3:
        for v in ENCODERS:
            assert np.array_equal(oracle.n_to_bits(n[3:].copy(), v), oracle.n_to_bits(n[3:], "lut")), v


@pytest.mark.parametrize("nwords", [0, 1, 2, 3, 17, 1250])
def test_all_decoder_variants_agree(oracle, nwords):
    if not oracle.simd_ok:
        pytest.skip("host CPU lacks AVX2/BMI2/PCLMUL")
    words = oracle.generate_words(nwords, seed=nwords)
    for length in sorted({0, max(0, 32 * nwords - 31), max(0, 32 * nwords - 1), 32 * nwords}):
        ref = oracle.bits_to_n(words, length, "lut")
        assert len(ref) == length
        for v in DECODERS[1:]:
            assert oracle.bits_to_n(words, length, v) == ref, (v, length)


@pytest.mark.parametrize("variant", DECODERS)
def test_length_check_is_the_reference_panic(oracle, variant):
    if variant != "lut" and not oracle.simd_ok:
        pytest.skip("host CPU lacks AVX2/BMI2/PCLMUL")
    words = np.zeros(2, dtype=np.uint64)
    with pytest.raises(ValueError, match=GOLD["panic_text"]):
        oracle.bits_to_n(words, 65, variant)
    assert oracle.bits_to_n(words, 64, variant) == b"A" * 64


def test_invalid_bytes_are_where_variants_disagree(oracle):
    # 'N' -> LUT says 0 ('A'), the bit trick says 3 ('G'); parity is defined on valid input only
    if not oracle.simd_ok:
        pytest.skip("host CPU lacks AVX2/BMI2/PCLMUL")
    n = b"N" * 32
    assert oracle.count_invalid(n) == 32 and oracle.count_invalid(b"ACGTUacgtu") == 0
    assert int(oracle.n_to_bits(n, "lut")[0]) == 0
    assert int(oracle.n_to_bits(n, "mul")[0]) == 0xFFFFFFFFFFFFFFFF


def test_multithreaded_drivers_match_single_thread(oracle):
    if not oracle.simd_ok:
        pytest.skip("host CPU lacks AVX2/BMI2/PCLMUL")
    for size in (0, 1, 33, 40000, (1 << 20) + 5):
        n = oracle.generate(size, seed=5, alphabet=10)
        ref = oracle.n_to_bits(n, "lut")
        for threads in (1, 2, 3, 8):
            assert np.array_equal(oracle.encode_mt(n, "movemask", threads), ref)
            assert np.array_equal(oracle.encode_mt(n, "lut", threads), ref)
            for v in ("shuffle", "lut"):
                assert oracle.decode_mt(ref, size, v, threads).tobytes() == oracle.canonical(n)


def test_generator_is_counter_based_and_shardable(oracle):
    a = oracle.generate(100003, seed=0xC0FFEE, alphabet=4)
    assert set(a.tobytes()) == set(b"ACGT")
    b = oracle.generate(50000, seed=0xC0FFEE, alphabet=4, offset=40000)
    assert np.array_equal(a[40000:90000], b)
    c = oracle.generate(100003, seed=0xC0FFEE, alphabet=10)
    assert set(c.tobytes()) == set(b"ACGTUacgtu")
    counts = np.bincount(a, minlength=128)[[65, 67, 71, 84]] / a.size
    assert np.all(np.abs(counts - 0.25) < 0.01)
    assert oracle.generate(64, seed=1).tobytes() != oracle.generate(64, seed=2).tobytes()
    w = oracle.generate_words(1000, seed=3)
    assert np.array_equal(w[100:200], oracle.generate_words(100, seed=3, first_word=100))


def test_config1_one_mib_roundtrip_on_cpu(oracle):
    """BASELINE config 1: 1 MiB random ACGT, n_to_bits_lut + bits_to_n_lut round trip, bit exact."""
    n = oracle.generate(1 << 20, seed=0xC0FFEE, alphabet=4)
    words = oracle.n_to_bits(n, "lut")
    assert words.size == 32768
    assert oracle.bits_to_n(words, n.size, "lut") == n.tobytes()


def test_property_roundtrip_and_shard_linearity(oracle):
    """Property checks with hypothesis: decode(encode(x)) == canonical(x); encoding is linear over concatenation at
    word boundaries (what makes sharding by offset exact); every restated variant agrees with the scalar pair."""
    from hypothesis import given, settings, strategies as st
    alphabet = st.sampled_from(list(b"ACGTUacgtu"))

    @settings(max_examples=150, deadline=None)
    @given(st.lists(alphabet, max_size=300), st.integers(min_value=0, max_value=9))
    def check(chars, cut_words):
        n = bytes(chars)
        words = oracle.n_to_bits(n, "lut")
        assert oracle.bits_to_n(words, len(n), "lut") == oracle.canonical(n)
        cut = min(len(n), 32 * cut_words)
        left, right = oracle.n_to_bits(n[:cut], "lut"), oracle.n_to_bits(n[cut:], "lut")
        assert np.array_equal(np.concatenate([left, right]), words)
        if oracle.simd_ok:
            for v in ("pext", "shift", "movemask", "mul"):
                assert np.array_equal(oracle.n_to_bits(n, v), words)
            for v in ("shuffle", "pdep", "clmul"):
                assert oracle.bits_to_n(words, len(n), v) == oracle.canonical(n)

    check()


# ---------------------------------------------------------------------------------------------------
# operations on packed words -- PARITY UNPINNED (no reference code): the oracle defines them through
# ASCII with the reference's codec; here that definition is cross-checked against the bit-level formulation
# (complement = code ^ 2, mismatch = non-zero 2-bit field of a ^ b) in numpy.
# ---------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("size", [0, 1, 31, 32, 33, 64, 1000, 4097])
def test_packed_ops_definition_vs_bit_formulation(oracle, size):
    n = oracle.generate(size, seed=size + 1, alphabet=10)
    m = oracle.generate(size, seed=size + 2, alphabet=10)
    a, b = oracle.n_to_bits(n, "lut"), oracle.n_to_bits(m, "lut")
    x = a ^ b
    fields = (x | (x >> np.uint64(1))) & np.uint64(0x5555555555555555)
    assert oracle.hamming(a, b, size) == sum(bin(int(f)).count("1") for f in fields)
    valid = np.full(a.size, 0xFFFFFFFFFFFFFFFF, dtype=np.uint64)
    if size % 32:
        valid[-1] = np.uint64((1 << (2 * (size % 32))) - 1)
    assert np.array_equal(oracle.complement(a, size), (a ^ np.uint64(0xAAAAAAAAAAAAAAAA)) & valid)
    codes = np.array([(int(a[i >> 5]) >> (2 * (i & 31))) & 3 for i in range(size)], dtype=np.uint64)
    rc_codes = (codes ^ np.uint64(2))[::-1]
    expect = np.zeros(a.size, dtype=np.uint64)
    for i, c in enumerate(rc_codes):
        expect[i >> 5] |= np.uint64(int(c) << (2 * (i & 31)))
    assert np.array_equal(oracle.complement(a, size, reverse=True), expect)
