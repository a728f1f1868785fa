"""CPU-side checks of the drop-in boundary: the library builds, loads, and exports every symbol that
include/rspnet_b200.h declares (no compute is launched here)."""
import re
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent


@pytest.fixture(scope="module")
def lib():
    from rspnet_b200 import build, _lib
    build.build()
    return _lib.load()


def _declared_symbols():
    text = (ROOT / "include" / "rspnet_b200.h").read_text()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(rsp_[a-z0-9_]+)\s*\(", text)))


def test_header_symbols_all_exported(lib):
    names = _declared_symbols()
    assert len(names) >= 30
    for n in names:
        assert hasattr(lib, n), f"{n} declared in the header but not exported by the library"


def test_binding_table_matches_header(lib):
    from rspnet_b200 import _lib
    assert sorted(_lib.SIGNATURES) == _declared_symbols()
    assert lib.rsp_abi_version() == 2


def test_no_cpu_path():
    import torch
    from rspnet_b200 import _lib
    if torch.cuda.is_available():
        pytest.skip("CUDA present")
    with pytest.raises(RuntimeError, match="no CUDA device"):
        _lib.call("rsp_init")


def test_sass_is_blackwell_native():
    """The conv kernels must contain tcgen05 MMA / TMEM loads (UTCHMMA / LDTM in SASS)."""
    import shutil
    import subprocess
    from rspnet_b200 import build
    if shutil.which("cuobjdump") is None:
        pytest.skip("cuobjdump not available")
    sass = subprocess.run(["cuobjdump", "-sass", str(build.build())], capture_output=True, text=True).stdout
    assert "UTCHMMA" in sass and "LDTM" in sass and "LDGSTS" in sass
