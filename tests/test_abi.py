"""The C-ABI library loads without a GPU and exports every symbol include/*.h declares."""
import ctypes as C
import importlib
import os
import re

from conftest import ROOT

pkg = importlib.import_module("rust-pseudoaligner_b200")
host = importlib.import_module("rust-pseudoaligner_b200.host")


def _declared(header):
    text = open(os.path.join(ROOT, "include", header)).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(psa_[a-z0-9_]+)\s*\(", text)))


def test_psa_h_symbols_exported():
    L = C.CDLL(pkg.lib_path())
    names = _declared("psa.h")
    assert len(names) >= 30
    for n in names:
        assert hasattr(L, n), "libpsa_b200.so lacks " + n
    assert set(names) == set(pkg.pseudoaligner.EXPORTS)


def test_psa_host_h_symbols_exported():
    L = C.CDLL(host.lib_path())
    for n in _declared("psa_host.h"):
        if n == "psa_process_reads":
            continue
        assert hasattr(L, n), "libpsa_host.so lacks " + n


def test_no_compute_without_gpu_is_an_error_not_a_fallback():
    """Without a device the library must fail loudly (PSA_ERR_CUDA), never compute on the CPU."""
    import numpy as np
    L = pkg.lib()
    assert L.psa_abi_version() == 2
    assert L.psa_strerror(-2) == b"CUDA error"
    try:
        import torch
        has_gpu = torch.cuda.is_available()
    except Exception:
        has_gpu = False
    if has_gpu:
        return
    flat = {"k": 5, "seq_words": np.zeros(1, np.uint64), "node_start": np.zeros(1, np.uint64),
            "node_len": np.array([5], np.uint32), "node_exts": np.zeros(1, np.uint8),
            "node_eq": np.zeros(1, np.uint32), "eq_offsets": np.array([0, 1], np.uint64),
            "eq_members": np.zeros(1, np.uint32)}
    try:
        pkg.Index(flat)
    except pkg.PsaError as e:
        assert e.code == -2
    else:
        raise AssertionError("index creation succeeded without a CUDA device")


def test_product_never_touches_the_oracle():
    """Nothing under the package may import, link or load oracle/ (or the test-only hostsim)."""
    pdir = os.path.join(ROOT, "rust-pseudoaligner_b200")
    for dp, _, files in os.walk(pdir):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".h")) or f == "Makefile":
                text = open(os.path.join(dp, f), errors="replace").read()
                for needle in ("libpsa_oracle", "import orc", "libhostsim", "orc_map", "orc_index"):
                    assert needle not in text, (f, needle)
                for line in text.splitlines():
                    if line.lstrip().startswith(("#include", "import ", "from ")):
                        assert "oracle" not in line and "hostsim" not in line, (f, line)
