"""CPU tests pinning the base-5 oracle (src/n_to_bits2.rs restatement) to the reference's known-answer tests
(src/n_to_bits2.rs:270-299, in tests/golden/kat.json) and cross-checking lut / pext / pdep / numpy."""
import json
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import cn_oracle_np as onp  # noqa: E402
from _oracle import DECODERS2, ENCODERS2  # noqa: E402

with open(os.path.join(ROOT, "tests", "golden", "kat.json")) as f:
    GOLD = json.load(f)

SIZES = [0, 1, 2, 3, 4, 5, 26, 27, 28, 31, 32, 33, 53, 54, 55, 80, 81, 82, 1000, 3455, 3456, 3457, 40000, 100003]


@pytest.mark.parametrize("variant", ENCODERS2)
def test_reference_kat_encode2(oracle, variant):
    if variant != "lut" and not oracle.simd_ok:
        pytest.skip("host CPU lacks AVX2/BMI2")
    for case in GOLD["encode2"]:
        got = oracle.n_to_bits2(case["input"].encode(), variant)
        assert [str(int(x)) for x in got] == case["words"], (variant, case["cite"])


@pytest.mark.parametrize("variant", DECODERS2)
def test_reference_kat_decode2(oracle, variant):
    if variant != "lut" and not oracle.simd_ok:
        pytest.skip("host CPU lacks AVX2/BMI2")
    for case in GOLD["decode2"]:
        words = np.array([int(w) for w in case["words"]], dtype=np.uint64)
        assert oracle.bits_to_n2(words, case["len"], variant) == case["output"].encode()


@pytest.mark.parametrize("size", SIZES)
def test_base5_variants_agree(oracle, size):
    n = oracle.generate2(size, seed=size + 3, alphabet=12)
    ref = oracle.n_to_bits2(n, "lut")
    assert ref.size == onp.words2_for_len(size) == oracle.words2_for_len(size)
    assert np.array_equal(ref, onp.n_to_bits2_lut(n))
    if oracle.simd_ok:
        assert np.array_equal(oracle.n_to_bits2(n, "pext"), ref)
    canon = oracle.canonical2(n)
    assert oracle.bits_to_n2(ref, size, "lut") == canon == onp.bits_to_n2_lut(ref, size)
    if oracle.simd_ok:
        assert oracle.bits_to_n2(ref, size, "pdep") == canon
    for shorter in {max(size - 1, 0), size // 2}:
        assert oracle.bits_to_n2(ref, shorter, "lut") == canon[:shorter]


def test_base5_length_panic(oracle):
    words = np.zeros(2, dtype=np.uint64)
    for v in DECODERS2:
        if v != "lut" and not oracle.simd_ok:
            continue
        with pytest.raises(ValueError, match=GOLD["panic_text"]):
            oracle.bits_to_n2(words, 55, v)
        assert oracle.bits_to_n2(words, 54, v) == b"A" * 54


def test_base5_generator(oracle):
    a = oracle.generate2(100003, seed=5, alphabet=5)
    assert set(a.tobytes()) == set(b"ACGTN")
    assert np.array_equal(a[40000:50000], oracle.generate2(10000, seed=5, alphabet=5, offset=40000))
    assert set(oracle.generate2(100003, seed=5, alphabet=12).tobytes()) == set(b"ACGTUNacgtun")
