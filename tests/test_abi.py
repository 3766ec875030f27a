"""CPU tests of the drop-in boundary: libyolo2_b200.so builds in-tree for sm_100a, loads
without a GPU, and exports every symbol include/*.h declares (no compute calls here)."""
import ctypes as C
import re
import subprocess
from pathlib import Path

import pytest

from sr_object_detection_b200 import _lib, build

ROOT = Path(__file__).resolve().parents[1]
DECL = re.compile(r"^[A-Za-z_][\w\s\*]*?[\s\*]([A-Za-z_]\w*)\s*\(", re.M)
NOT_FUNCTIONS = {"if", "for", "while", "switch", "return", "sizeof", "defined"}


def _declared(header: Path):
    text = header.read_text()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    text = re.sub(r"//[^\n]*", "", text)
    text = re.sub(r"^\s*#.*$", "", text, flags=re.M)
    text = re.sub(r"typedef\s+struct\s*\w*\s*\{.*?\}\s*\w+\s*;", "", text, flags=re.S)   # struct bodies
    text = re.sub(r"struct\s+\w+\s*\{.*?\}\s*;", "", text, flags=re.S)
    text = re.sub(r"typedef\s+enum\s*\{.*?\}\s*\w+\s*;", "", text, flags=re.S)
    names = set()
    for stmt in text.split(";"):
        stmt = stmt.strip()
        if not stmt or stmt.startswith("typedef") or "(" not in stmt or "{" in stmt:
            continue
        m = re.match(r"^[\w\s\*]+?[\s\*]([A-Za-z_]\w*)\s*\(", stmt.replace("\n", " "))
        if m and m.group(1) not in NOT_FUNCTIONS:
            names.add(m.group(1))
    return names


@pytest.fixture(scope="module")
def exported():
    lib = build.build()
    out = subprocess.run(["nm", "-D", "--defined-only", str(lib)], capture_output=True, text=True, check=True).stdout
    return {line.split()[-1] for line in out.splitlines() if line.strip()}


def test_library_is_built_in_tree_for_sm100a():
    lib = build.build()
    assert lib.exists() and lib.parent == ROOT / "sr_object_detection_b200"
    sass = subprocess.run(["cuobjdump", "-lelf", str(lib)], capture_output=True, text=True).stdout
    assert "sm_100a" in sass, sass[:400]


def test_kernel_abi_header_symbols_are_exported(exported):
    names = _declared(ROOT / "include" / "yolo2_b200_kernels.h")
    assert len(names) >= 35, sorted(names)
    missing = sorted(n for n in names if n not in exported)
    assert not missing, f"declared in yolo2_b200_kernels.h but not exported: {missing}"


def test_darknet_api_header_symbols_are_exported(exported):
    names = _declared(ROOT / "include" / "darknet_b200.h")
    assert {"parse_network_cfg", "load_weights", "network_predict", "get_region_boxes", "do_nms_sort",
            "set_batch_network", "resize_network", "free_network", "network_detect_batch"} <= names
    missing = sorted(n for n in names if n not in exported)
    assert not missing, f"declared in darknet_b200.h but not exported: {missing}"
    assert "gpu_index" in exported


def test_detector_class_symbols_are_exported(exported):
    """yolo_v2_class.hpp's Detector (C++) and its C shim."""
    hpp = ROOT / "include" / "yolo_v2_class.hpp"
    if not hpp.exists():
        pytest.skip("Detector class not built yet")
    mangled = [s for s in exported if "Detector" in s]
    for member in ("detect", "load_image", "free_image", "tracking", "get_net_width", "get_net_height"):
        assert any(member in s for s in mangled), member


def test_library_loads_without_a_gpu_and_reports_identity():
    lib = _lib.load()
    assert b"yolo2-b200" in lib.y2_version()
    # struct mirrors used for by-value passing match the C side
    from sr_object_detection_b200 import darknet as dn
    dn.lib()
    assert lib.y2_abi_sizeof(4) == C.sizeof(dn.Box)


def test_no_cpu_fallback_symbols():
    """The product must not carry a CPU compute path: no im2col/gemm symbols, and nothing from
    oracle/ linked in."""
    lib = build.build()
    out = subprocess.run(["nm", "-D", "--defined-only", str(lib)], capture_output=True, text=True, check=True).stdout
    for forbidden in ("gemm_cpu", "im2col_cpu", "forward_convolutional_layer\n", "y2_oracle", "matmul_acc"):
        assert forbidden not in out
