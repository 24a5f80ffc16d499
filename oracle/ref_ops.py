"""Backend over the REFERENCE's own gsplat CUDA kernels (oracle/_ref/libgsplat_ref.so, built by
oracle/build_ref.py from /root/reference/gsplat with the GLM shim).

TEST INFRASTRUCTURE ONLY: lets tests and bench.py drive the reference's kernels through the very
same Python call sites as the product backend (OpsBackend of the package), to pin parity on the
rows the reference's tests do not cover and to time the reference's CUDA build on the same box.
"""
from __future__ import annotations

import os

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
REF_SO = os.path.join(_HERE, "_ref", "libgsplat_ref.so")
_loaded = False


def available() -> bool:
    return os.path.exists(REF_SO)


def _ns():
    global _loaded
    if not _loaded:
        if not available():
            raise RuntimeError(f"{REF_SO} not built (python oracle/build_ref.py in the build container)")
        torch.ops.load_library(REF_SO)
        _loaded = True
    return torch.ops.gsplat_ref


def backend(pkg):
    """An OpsBackend (class from the product package) bound to the reference library."""
    return pkg.OpsBackend(_ns)


PRECISE_SO = os.path.join(_HERE, "_ref", "libgsplat_ref_precise.so")
_loaded_precise = False


def precise_available() -> bool:
    return os.path.exists(PRECISE_SO)


def _ns_precise():
    global _loaded_precise
    if not _loaded_precise:
        if not precise_available():
            raise RuntimeError(f"{PRECISE_SO} not built (python oracle/build_ref.py in the build container)")
        torch.ops.load_library(PRECISE_SO)
        _loaded_precise = True
    return torch.ops.gsplat_ref_precise


def backend_precise(pkg):
    """The same reference sources compiled WITHOUT --use_fast_math (oracle/build_ref.py, VARIANTS): the oracle of the
    rolling-shutter tests, whose fast-math build is dominated by sin.approx noise."""
    return pkg.OpsBackend(_ns_precise)
