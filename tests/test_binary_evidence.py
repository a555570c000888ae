"""CPU-side guards on the BUILT library: the Blackwell instruction paths must stay in the binary (cuobjdump needs no GPU).

If a refactor silently turns a tcgen05 kernel into a legacy `mma.sync` one, or drops the block-scaled / TMEM-operand / peer paths,
these fail here, before any GPU time is spent."""

import csv
import io
import shutil
import subprocess
import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent
LIB = ROOT / "prime_b200" / "_C" / "libprime_b200.so"

needs_toolchain = pytest.mark.skipif(shutil.which("cuobjdump") is None or shutil.which("nvcc") is None, reason="CUDA toolkit not installed")


@pytest.fixture(scope="module")
def sass() -> str:
    from prime_b200.ops import _lib

    _lib.build()  # no-op when the sources are unchanged
    return subprocess.run(["cuobjdump", "-sass", str(LIB)], capture_output=True, text=True, check=True).stdout


def _kernels(sass: str) -> dict[str, str]:
    out, cur = {}, None
    for line in sass.splitlines():
        if "Function :" in line:
            cur = line.split("Function :")[1].strip()
            out[cur] = ""
        elif cur is not None:
            out[cur] += line + "\n"
    return out


@needs_toolchain
def test_no_legacy_tensor_core_instructions(sass):
    import re

    assert not re.search(r"\sHMMA\.", sass), "legacy mma.sync (HMMA) found: a kernel fell off the tcgen05 path"
    assert "arch = sm_100a" in sass


@needs_toolchain
def test_blackwell_paths_are_present(sass):
    k = _kernels(sass)

    def body(*needles):
        hits = [b for n, b in k.items() if all(x in n for x in needles)]
        assert hits, f"no kernel matching {needles}"
        return "\n".join(hits)

    gemm = body("gemm_bf16_kernel")
    assert "UTCHMMA.2CTA" in gemm and "UTMALDG" in gemm and "UTMAREDG" in gemm and "LDTM" in gemm  # cta_group::2 MMAs, TMA load / reduce-add
    mx = body("gemm_mxfp8_kernel")
    assert "UTCQMMA" in mx and "UTCCP" in mx  # block-scaled MMA with the scales copied into TMEM
    fwd2 = body("flash_fwd2_kernel")
    assert "UTCHMMA" in fwd2 and "STTM" in fwd2 and "MUFU.EX2" in fwd2  # P written to TMEM, consumed as an MMA operand
    assert "UTCHMMA" in body("bwd_dkdv_kernel") and "UTCHMMA" in body("bwd_dq_kernel")
    comm = body("adamw_push_kernel")
    assert "STG.E.128" in comm  # 128-bit peer stores
    assert "LD.E.STRONG.SYS" in body("outer_nesterov") or "STRONG.SYS" in body("outer_nesterov")
    # NVLS: the reduction is a multimem load (LDGMC), in f32 and in bf16 with fp32 accumulation in the switch
    assert "LDGMC.E.ADD.F32x4" in body("mc_grad_reduce_kernel")
    allred = body("mc_all_reduce_kernel")
    assert "LDGMC.E.ADD.F32x4" in allred and "BF16x8" in allred and "REDG.E.ADD.STRONG.SYS" in allred


def test_ncu_source_hotspots_condenses_runs(tmp_path, capsys):
    sys.path.insert(0, str(ROOT / "tools"))
    import ncu_source_hotspots as H

    buf = io.StringIO()
    w = csv.writer(buf)
    w.writerow(["Kernel Name", "void demo()"])
    w.writerow(["Address", "Source", "# Samples", "Instructions Executed"])
    rows = [("a", "      LDG.E R1, [R2]", 5, 100), ("b", "      FFMA R3, R1, R4, R5", 90, 100), ("c", "@P0   BRA 0x10", 4, 100),
            ("d", "      EXIT", 1, 1)]  # fmt: skip
    for a, s, n, e in rows:
        w.writerow([a, s, n, e])
    p = tmp_path / "src.csv"
    p.write_text(buf.getvalue())
    H.main(str(p), 2)
    out = capsys.readouterr().out
    assert "4 instructions, 100 stall samples" in out
    assert "exec       100" in out and "FFMA 90" in out and "(90.0 %)" in out


def test_step_composition_is_reproducible_from_the_committed_launch_list(capsys):
    """profiles/step_composition_r2_final.md is what tools/launch_breakdown.py prints for the committed ncu launch list."""
    import sys

    sys.path.insert(0, str(ROOT))
    from tools.launch_breakdown import family, main

    csv = ROOT / "profiles" / "launches_1gpu_step_r2_final.csv"
    if not csv.exists():
        pytest.skip("launch list not committed")
    main(str(csv))
    out = capsys.readouterr().out
    rows = {ln.split("|")[1].strip(): float(ln.split("|")[4].strip().rstrip(" %")) for ln in out.splitlines() if ln.startswith("| ") and ln.rstrip().endswith("% |") and "`" not in ln}
    assert rows["GEMM (tcgen05)"] > 70 and rows["attention"] < 16 and rows["torch / runtime kernels"] < 0.5, rows
    assert family("void <unnamed>::gemm_bf16_kernel<1, 1, 1, 0, 0>(CUtensorMap_st)") == "GEMM (tcgen05)"
    assert family("void at::native::vectorized_elementwise_kernel<4, at::native::FillFunctor<float>>") == "torch / runtime kernels"
