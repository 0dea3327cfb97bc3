"""CPU tests of the drop-in boundary: the C-ABI library loads, exports every symbol the header declares
(and nothing else), and refuses to compute without a GPU instead of falling back."""
import os
import re
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "cute_nucleotides_cuda.h")


def declared_symbols():
    text = open(HEADER).read()
    return sorted(set(re.findall(r"^CN_API\s+[\w\s\*]+?\b(cn_\w+)\s*\(", text, flags=re.M)))


def test_header_declares_the_path():
    syms = declared_symbols()
    for must in ("cn_n_to_bits_host", "cn_bits_to_n_host", "cn_encode_device", "cn_decode_device", "cn_last_error"):
        assert must in syms


def test_library_exports_exactly_the_header(cn):
    from cute_nucleotides_b200 import _lib
    lib = _lib.load()
    syms = declared_symbols()
    assert sorted(_lib.PROTOTYPES) == syms          # the Python binding covers the whole header
    for s in syms:
        assert hasattr(lib, s), s
    out = subprocess.run(["nm", "-D", "--defined-only", _lib.LIB_PATH], capture_output=True, text=True, check=True).stdout
    exported = sorted(line.split()[-1] for line in out.splitlines() if " T " in line)
    assert exported == syms, "the library must export the header's symbols and nothing else (no leaked cudart)"


def test_pure_host_entry_points(cn):
    from cute_nucleotides_b200 import _lib
    lib = _lib.load()
    assert lib.cn_abi_version() == _lib.ABI_VERSION
    for length in (0, 1, 31, 32, 33, 64, 65, 10 * (1 << 30), 10 * (1 << 30) + 1):
        assert lib.cn_words_for_len(length) == (length + 31) // 32 == cn.words_for_len(length)
    assert lib.cn_length_panic_message() == b"The length is greater than the number of nucleotides!"
    assert lib.cn_set_tuning(0, 32, 1, 256) == 0
    assert lib.cn_set_tuning(0, 24, 4, 256) == _lib.CN_ERR_ARG and b"unsupported" in lib.cn_last_error()


def test_shard_planner_matches_python(cn):
    """cn_shard_bounds (what a Rust / C++ caller plans with) == sharded.shard_bounds (what the torch.distributed path uses)."""
    from cute_nucleotides_b200 import sharded
    for total in (0, 1, 31, 32, 33, 100003, (1 << 26) + 77, 10 * (1 << 30) + 21):
        for world in (1, 2, 3, 8):
            for granule in (32, 1 << 20):
                for r in range(world):
                    assert cn.shard_bounds_c(total, world, r, granule) == sharded.shard_bounds(total, world, r, granule)
    for total, world, granule in ((10 ** 6, 4, 27 * 128), (27 * 5 + 3, 3, 27)):
        for r in range(world):
            assert cn.shard_bounds_c(total, world, r, granule) == sharded.shard_bounds(total, world, r, granule, group=27)


def test_host_knobs_validate_their_arguments(cn):
    from cute_nucleotides_b200 import _lib
    lib = _lib.load()
    assert lib.cn_set_host_threads(-1) == _lib.CN_ERR_ARG and lib.cn_set_host_threads(65) == _lib.CN_ERR_ARG
    assert lib.cn_set_host_threads(0) == 0
    assert lib.cn_set_host_chunks(1000, 0) == _lib.CN_ERR_ARG and lib.cn_set_host_chunks(0, 4097) == _lib.CN_ERR_ARG
    assert lib.cn_set_host_chunks(0, 0) == 0 and lib.cn_set_host_chunks(16 << 20, 4 << 20) == 0
    assert cn.get_devices() == []                     # no device set unless cn_set_devices / CN_DEVICES says so


def test_length_check_needs_no_gpu(cn):
    # len > 32*nwords is rejected before any CUDA call, with the reference's panic text (src/n_to_bits.rs:52-54)
    with pytest.raises(cn.LengthError, match="The length is greater than the number of nucleotides!"):
        cn.bits_to_n_cuda(np.zeros(2, dtype=np.uint64), 65)


def test_empty_inputs_need_no_gpu(cn):
    assert cn.n_to_bits_cuda(b"").size == 0
    assert cn.bits_to_n_cuda(np.zeros(0, dtype=np.uint64), 0) == b""
    assert cn.bits_to_n_cuda(np.zeros(3, dtype=np.uint64), 0) == b""


def test_no_cpu_fallback_without_a_gpu(cn):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present: covered by the -m gpu tests")
    from cute_nucleotides_b200._lib import CuteNucleotidesError
    with pytest.raises(CuteNucleotidesError) as e:
        cn.n_to_bits_cuda(b"ATCG")
    assert e.value.code == 2      # CN_ERR_CUDA: loud failure, not a silent CPU result
    with pytest.raises(CuteNucleotidesError):
        cn.bits_to_n_cuda(np.array([0xD8], dtype=np.uint64), 4)


def test_product_never_touches_the_oracle():
    pkg = os.path.join(ROOT, "cute_nucleotides_b200")
    for dirpath, _, files in os.walk(pkg):
        for fn in files:
            if fn.endswith((".py", ".cu", ".cuh", ".cpp", ".hpp", ".h")):
                text = open(os.path.join(dirpath, fn)).read()
                assert "libcn_oracle" not in text and "cn_oracle" not in text and "oracle_" not in text, fn


def test_oracle_is_confined_to_its_allowed_callers():
    """Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may touch oracle/."""
    import ast
    offenders = []
    for dirpath, dirnames, files in os.walk(ROOT):
        dirnames[:] = [d for d in dirnames if d not in (".git", "tests", "oracle", "gpurun_out", "__pycache__", ".pytest_cache", ".hypothesis")]
        for fn in files:
            if not fn.endswith((".py", ".cu", ".cuh", ".cpp", ".hpp", ".h", ".rs")):
                continue
            path = os.path.join(dirpath, fn)
            rel = os.path.relpath(path, ROOT)
            text = open(path, errors="replace").read()
            if "_oracle" not in text and "cn_oracle" not in text and "libcn_oracle" not in text:
                continue
            if rel == "bench.py":
                tree = ast.parse(text)
                for node in ast.walk(tree):
                    if isinstance(node, ast.FunctionDef):
                        uses = any(isinstance(n, ast.Import) and any(a.name == "_oracle" for a in n.names) for n in ast.walk(node))
                        if uses and node.name not in ("cpu_roundtrip", "run_reference"):
                            offenders.append(f"bench.py:{node.name}")
            elif rel == "__graft_entry__.py":
                tree = ast.parse(text)
                for node in ast.walk(tree):
                    if isinstance(node, ast.FunctionDef):
                        uses = any(isinstance(n, ast.Import) and any(a.name == "_oracle" for a in n.names) for n in ast.walk(node))
                        if uses and node.name != "smoke":
                            offenders.append(f"__graft_entry__.py:{node.name}")
            else:
                offenders.append(rel)
    assert offenders == [], offenders
