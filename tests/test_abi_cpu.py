"""CPU checks of the C-ABI boundary: the in-tree library builds for sm_100a, loads without a GPU, exports every symbol
include/unigen_b200.h declares (and nothing the ctypes layer does not know), refuses CPU tensors loudly, and reports
"no device" instead of falling back."""
import ctypes
import re
from pathlib import Path

import pytest
import torch

ROOT = Path(__file__).resolve().parent.parent
HEADER = ROOT / "include" / "unigen_b200.h"


def _declared_symbols():
    text = HEADER.read_text()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(ug_[a-z0-9_]+)\s*\(", text)))


def test_library_builds_and_exports_every_declared_symbol():
    import __graft_entry__ as g
    g.build()
    from unigen_b200 import _lib
    lib = ctypes.CDLL(str(_lib.LIB_PATH))
    declared = _declared_symbols()
    assert len(declared) >= 20
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in the header but not exported"
    assert set(declared) == set(_lib.SIGNATURES), set(declared) ^ set(_lib.SIGNATURES)
    assert lib.ug_abi_version() == int(re.search(r"#define UG_ABI_VERSION (\d+)", HEADER.read_text()).group(1))


def test_struct_layouts_match_header_field_order():
    from unigen_b200 import _lib
    text = HEADER.read_text()
    for struct, cls in (("ug_gemm_args", _lib.GemmArgs), ("ug_attn_args", _lib.AttnArgs), ("ug_gemv_job", _lib.GemvJob),
                        ("ug_flux_desc", _lib.FluxDesc), ("ug_flux_inputs", _lib.FluxInputs), ("ug_flux_outputs", _lib.FluxOutputs),
                        ("ug_conv2d_args", _lib.Conv2dArgs)):
        body = re.search(r"typedef struct %s \{(.*?)\} %s;" % (struct, struct), text, flags=re.S).group(1)
        body = re.sub(r"/\*.*?\*/", "", body, flags=re.S)
        names = []
        for decl in body.split(";"):
            decl = decl.strip()
            if not decl:
                continue
            decl = re.sub(r"\[[^\]]*\]", "", decl)
            for part in decl.split(","):
                names.append(part.strip().split()[-1].lstrip("*"))
        assert names == [f[0] for f in cls._fields_], (struct, names)


def test_no_device_is_an_error_not_a_fallback():
    from unigen_b200 import _lib, ops
    lib = _lib.load()
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    assert lib.ug_device_check() != 0
    assert lib.ug_last_error()
    with pytest.raises(ops.UgError):
        ops.gemm(torch.zeros(1, 8, 64, dtype=torch.bfloat16), torch.zeros(8, 64, dtype=torch.bfloat16))
    from unigen_b200.model import FluxArch, UniGenFlux
    with pytest.raises(ops.UgError):
        UniGenFlux(FluxArch.tiny(), device="cpu")


def test_product_package_never_imports_the_oracle():
    for f in (ROOT / "unigen_b200").glob("*.py"):
        src = f.read_text()
        assert "import oracle" not in src and "from oracle" not in src, f


def test_struct_offsets_match_a_compiled_probe(tmp_path):
    """sizeof / offsetof of every ABI struct as gcc lays them out == the ctypes mirrors the Python host side passes."""
    import ctypes as C
    import shutil
    import subprocess
    from unigen_b200 import _lib
    if shutil.which("gcc") is None:
        pytest.skip("gcc not available")
    structs = {"ug_gemm_args": _lib.GemmArgs, "ug_attn_args": _lib.AttnArgs, "ug_peer_table": _lib.PeerTable,
               "ug_qkv_scatter_args": _lib.QkvScatterArgs, "ug_conv2d_args": _lib.Conv2dArgs}
    lines = ['#include <stdio.h>', '#include <stddef.h>', f'#include "{HEADER}"', "int main(void) {"]
    for name, cls in structs.items():
        lines.append(f'  printf("{name} %zu\\n", sizeof({name}));')
        for field, _ in cls._fields_:
            lines.append(f'  printf("{name}.{field} %zu\\n", offsetof({name}, {field}));')
    lines += ["  return 0;", "}"]
    src = tmp_path / "probe.c"
    src.write_text("\n".join(lines))
    exe = tmp_path / "probe"
    subprocess.run(["gcc", "-std=c11", "-o", str(exe), str(src)], check=True)
    out = dict(ln.split() for ln in subprocess.run([str(exe)], capture_output=True, text=True, check=True).stdout.splitlines())
    for name, cls in structs.items():
        assert int(out[name]) == C.sizeof(cls), name
        for field, _ in cls._fields_:
            assert int(out[f"{name}.{field}"]) == getattr(cls, field).offset, (name, field)
