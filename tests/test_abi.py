"""CPU-side checks of the drop-in boundary: the C-ABI library loads, exports every symbol
include/hvb.h declares, the ctypes mirrors have the sizes the header's structs have, and the
product path fails loudly (no CPU fallback) when no B200 is present."""
import ctypes
import re
import subprocess
import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent


@pytest.fixture(scope="module")
def lib():
    from turingcodec_b200 import build, hvb
    build.build()
    return hvb.load_library()


def test_every_declared_symbol_is_exported(lib):
    from turingcodec_b200 import hvb
    names = hvb.exported_symbols()
    assert len(names) >= 30
    missing = [n for n in names if not hasattr(lib, n)]
    assert not missing, missing


def test_struct_sizes_match_header(tmp_path):
    """compile a C program against include/hvb.h printing sizeof of each struct; compare with numpy dtypes"""
    from turingcodec_b200 import hvb
    structs = {"hvb_block": hvb.block_t, "hvb_metric_task": hvb.metric_task_t, "hvb_sad4_task": hvb.sad4_task_t,
               "hvb_pred_task": hvb.pred_task_t, "hvb_subtract_bi_task": hvb.subtract_bi_task_t,
               "hvb_interp_satd_task": hvb.interp_satd_task_t, "hvb_intra_task": hvb.intra_task_t,
               "hvb_intra_sweep_task": hvb.intra_sweep_task_t, "hvb_transform_task": hvb.transform_task_t,
               "hvb_quant_task": hvb.quant_task_t, "hvb_ita_task": hvb.ita_task_t, "hvb_tu_task": hvb.tu_task_t,
               "hvb_tu_result": hvb.tu_result_t, "hvb_rdoq_ctx": hvb.rdoq_ctx_t, "hvb_rdoq_task": hvb.rdoq_task_t,
               "hvb_me_task": hvb.me_task_t, "hvb_me_result": hvb.me_result_t,
               "hvb_deblock_block": hvb.deblock_block_t, "hvb_deblock_ctu": hvb.deblock_ctu_t, "hvb_deblock_task": hvb.deblock_task_t,
               "hvb_sao_ctu": hvb.sao_ctu_t, "hvb_sao_task": hvb.sao_task_t,
               "hvb_sao_stats_task": hvb.sao_stats_task_t, "hvb_sao_stats": hvb.sao_stats_t,
               "hvb_coded_residual_task": hvb.coded_residual_task_t, "hvb_coded_residual": hvb.coded_residual_t,
               "hvb_intra_complexity_task": hvb.intra_complexity_task_t}
    src = tmp_path / "sizes.c"
    body = "\n".join(f'printf("{n} %zu\\n", sizeof({n}));' for n in structs)
    src.write_text(f'#include <stdio.h>\n#include "hvb.h"\nint main(void){{{body} return 0;}}\n')
    exe = tmp_path / "sizes"
    subprocess.run(["gcc", "-std=c11", "-I", str(ROOT / "include"), str(src), "-o", str(exe)], check=True)
    out = subprocess.run([str(exe)], check=True, capture_output=True, text=True).stdout
    for line in out.splitlines():
        name, size = line.split()
        assert structs[name].itemsize == int(size), name
    # field offsets of the two structs with 64-bit members
    assert hvb.me_task_t.fields["rateMvpFlag"][1] == 24 and hvb.me_task_t.fields["lambda"][1] == 40
    assert hvb.me_result_t.fields["cost"][1] == 16


def test_no_cpu_fallback_without_a_device(lib):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    from turingcodec_b200 import hvb
    assert lib.hvb_device_ok(0) == 0
    with pytest.raises(hvb.HvbError):
        hvb.Context(0, 1, 8)


def test_product_never_touches_the_oracle():
    """the oracle is test infrastructure: nothing under turingcodec_b200/ or include/ may reference it"""
    offenders = []
    for path in list((ROOT / "turingcodec_b200").rglob("*")) + list((ROOT / "include").rglob("*")):
        if path.suffix in {".py", ".cu", ".cuh", ".cpp", ".h"}:
            text = path.read_text(errors="ignore")
            if re.search(r"liboracle|oracle/|orc_[a-z]|libhavoc_ref", text):
                offenders.append(str(path))
    assert not offenders, offenders
