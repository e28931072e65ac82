"""CPU: the C-ABI shared library loads and exports every symbol include/locohd_b200.h declares; without a GPU
the product path fails loudly (no CPU fallback, no oracle on the product path)."""
import ctypes
import re
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent


@pytest.fixture(scope="module")
def lib():
    from loco_hd_b200 import build, _capi
    build.build_cuda_lib()
    return _capi.load_library()


def test_header_symbols_are_exported(lib):
    from loco_hd_b200 import _capi
    header = re.sub(r"/\*.*?\*/", "", (ROOT / "include" / "locohd_b200.h").read_text(), flags=re.S)
    declared = sorted(set(re.findall(r"\b(locohd_[a-z0-9_]+)\s*\(", header)))
    assert declared, "no declarations found in the header"
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in include/locohd_b200.h but not exported"
    assert sorted(_capi.EXPORTED_SYMBOLS) == declared
    version = int(re.search(r"#define\s+LOCOHD_ABI_VERSION\s+(\d+)", header).group(1))
    assert lib.locohd_abi_version() == version == 2


def test_struct_layouts_match_header(tmp_path):
    """ctypes mirrors of the ABI structs have the layout gcc gives the header's structs."""
    import subprocess
    from loco_hd_b200 import _capi
    src = tmp_path / "layout.c"
    src.write_text(
        '#include <stdio.h>\n#include <stddef.h>\n#include "locohd_b200.h"\n'
        'int main(void){printf("%zu %zu %zu %zu %zu %zu %zu\\n", sizeof(locohd_weight_function), sizeof(locohd_job),'
        ' sizeof(locohd_params), offsetof(locohd_params, sd_params), offsetof(locohd_params, weight_functions),'
        ' offsetof(locohd_params, n_tag_pairs), offsetof(locohd_params, tag_pairs));return 0;}\n')
    exe = tmp_path / "layout"
    subprocess.run(["gcc", f"-I{ROOT / 'include'}", str(src), "-o", str(exe)], check=True)
    want = [int(v) for v in subprocess.run([str(exe)], check=True, capture_output=True, text=True).stdout.split()]
    P = _capi.ParamsC
    got = [ctypes.sizeof(_capi.WeightFunctionC), ctypes.sizeof(_capi.JobC), ctypes.sizeof(P), P.sd_params.offset,
           P.weight_functions.offset, P.n_tag_pairs.offset, P.tag_pairs.offset]
    assert got == want
    assert _capi.JOB_DTYPE.itemsize == ctypes.sizeof(_capi.JobC)


def test_no_gpu_fails_loudly(lib):
    """On a box without a CUDA device context creation must raise (status 101), never fall back."""
    from loco_hd_b200 import _capi
    if lib.locohd_device_count() > 0:
        pytest.skip("a CUDA device is present")
    with pytest.raises(_capi.LocoHDError) as e:
        _capi.Context(0)
    assert e.value.status == 101 and "no CPU fallback" in str(e.value)


def test_product_never_imports_the_oracle():
    for path in list((ROOT / "loco_hd_b200").rglob("*.py")) + list((ROOT / "loco_hd").rglob("*.py")) + \
            [q for q in (ROOT / "loco_hd_b200" / "csrc").glob("*") if q.is_file()]:
        text = path.read_text(errors="ignore")
        assert not re.search(r"^\s*(from|import)\s+oracle\b", text, re.M), f"{path} imports the oracle"
        assert "locohd_oracle" not in text, f"{path} references the oracle"
        # the host layer carries no PyTorch (north_star): torch is bench / test plumbing only
        if path.suffix == ".py":
            assert not re.search(r"^\s*(from|import)\s+torch\b", text, re.M), f"{path} imports torch"


def test_tile_unit_order_is_a_bijection():
    """locohd_tile_unit (host-side query of the tile kernel's unit order; the same inline function the kernel uses, in
    its 64-bit and 32-bit forms): every (tile, anchor) exactly once, slice by slice, for slices that divide the anchor
    count, that leave a short last slice, that exceed it, and for the tile-major order (slice 0)."""
    from loco_hd_b200 import _capi

    for n_tiles, n, s in [(5, 37, 8), (5, 40, 8), (3, 10, 0), (3, 10, 16), (4, 203, 24), (1, 9, 8), (6, 16, 16), (2, 1, 8)]:
        seen = [_capi.tile_unit(n_tiles, n, s, u) for u in range(n_tiles * n)]
        assert len(set(seen)) == n_tiles * n and all(0 <= t < n_tiles and 0 <= p < n for t, p in seen)
        if 0 < s < n:
            # slice-major: the slice index never decreases, and within a slice the tile index never decreases
            key = [(p // s, t) for t, p in seen]
            assert key == sorted(key)
            assert [p for t, p in seen[:s]] == list(range(s))   # a run of consecutive units = consecutive anchors of a tile
        else:
            assert seen == [(u // n, u % n) for u in range(n_tiles * n)]
    assert _capi.tile_unit(31219, 5000, 16, 31219 * 16 + 5) == (0, 21)
    assert _capi.tile_unit(2 ** 20, 5000, 16, 2 ** 20 * 16 * 7 + 16 * 3 + 2) == (3, 7 * 16 + 2)   # > 2^32 units: 64-bit form only
    with pytest.raises(_capi.LocoHDError):
        _capi.tile_unit(3, 10, 8, 30)
