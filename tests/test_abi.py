"""The C ABI: libfastq_gpu.so loads and exports every function include/fastq_gpu.h declares (no compute without a GPU)."""
import ctypes
import os
import re
import subprocess

import pytest

from _util import ROOT

SO = os.path.join(ROOT, "fastq_utils_b200", "libfastq_gpu.so")
HDR = os.path.join(ROOT, "include", "fastq_gpu.h")


@pytest.fixture(scope="module")
def built():
    if not os.path.exists(SO):
        subprocess.check_call(["make", "-C", os.path.join(ROOT, "fastq_utils_b200", "csrc"), "all"], stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
    return ctypes.CDLL(SO)


def declared_functions():
    text = open(HDR).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(fqg_[a-z_0-9]+)\s*\(", text)))


def test_header_declares_the_boundary():
    names = declared_functions()
    for must in ["fqg_create", "fqg_feed", "fqg_feed_device", "fqg_finish", "fqg_render", "fqg_fastq_info_mem", "fqg_index_records"]:
        assert must in names


def test_every_declared_symbol_is_exported(built):
    for name in declared_functions():
        assert hasattr(built, name), name


def test_no_torch_types_in_signatures():
    text = re.sub(r"/\*.*?\*/", "", open(HDR).read(), flags=re.S)  # declarations only
    assert "torch" not in text and "at::" not in text and "std::" not in text and "Tensor" not in text


def test_create_fails_loudly_without_device(built):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a CUDA device is present")
    import fastq_utils_b200 as fq
    with pytest.raises(RuntimeError):
        fq.FastqInfo(fq.MODE_SINGLE)
    with pytest.raises(RuntimeError):
        fq.fastq_info(["a.fq"], b"@r\nA\n+\nI\n")


def test_product_does_not_link_the_oracle_or_the_stand_in(built):
    out = subprocess.run(["nm", "-D", SO], capture_output=True, text=True).stdout
    assert "oracle_" not in out and "FqSimDevice" not in out
    ldd = subprocess.run(["ldd", SO], capture_output=True, text=True).stdout
    assert "liboracle" not in ldd and "libfastq_sim" not in ldd
