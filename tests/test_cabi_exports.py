"""CPU-side checks of the boundary: the C-ABI library builds for sm_100a, loads, and exports every symbol
include/reptext_rt.h declares.  No compute call is made (there is no GPU here)."""
import ctypes
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    src = open(os.path.join(ROOT, "include", "reptext_rt.h")).read()
    return sorted(set(re.findall(r"RT_API\s+[\w\s\*]+?\b(rt_\w+)\s*\(", src)))


def test_library_builds_loads_and_exports_the_header():
    from reptext_b200 import _lib, build
    path = build.build(force=False, verbose=False)
    assert os.path.exists(path)
    lib = ctypes.CDLL(path)
    names = _declared()
    assert len(names) >= 24
    missing = [n for n in names if not hasattr(lib, n)]
    assert not missing, missing
    assert sorted(_lib.EXPORTS) == names           # the ctypes binding covers exactly the header
    lib.rt_abi_version.restype = ctypes.c_int
    assert lib.rt_abi_version() == 2                # host-only call


def test_host_only_entry_points_validate_arguments():
    from reptext_b200 import _lib as L
    lib = L.lib()
    h = ctypes.c_void_p()
    cfg = L.ModelConfig()
    cfg.kind, cfg.dtype, cfg.attention_head_dim = 7, L.RT_BF16, 128
    assert lib.rt_model_create(ctypes.byref(cfg), ctypes.byref(h)) == L.RT_ERR_INVALID
    assert b"kind" in lib.rt_last_error()
    assert lib.rt_set_option(b"no_such_option", 1) == L.RT_ERR_INVALID
    v = ctypes.c_int(-1)
    assert lib.rt_get_option(b"profile", ctypes.byref(v)) == 0 and v.value == 0
    assert lib.rt_gemm(None, 0, None) == L.RT_ERR_INVALID
    assert lib.rt_euler_step(L.RT_BF16, None, None, None, 8, 0.0, 0.0, None) == L.RT_ERR_INVALID


def test_sass_contains_blackwell_tensor_and_tma_instructions():
    """tcgen05.mma -> UTC*MMA, tcgen05.ld/st -> LDTM/STTM, TMA -> UTMALDG (B200_PROFILING.md)."""
    import subprocess
    from reptext_b200 import build
    path = build.build(force=False, verbose=False)
    sass = subprocess.run(["cuobjdump", "-sass", path], capture_output=True, text=True).stdout
    for mnemonic in ("UTCHMMA", "LDTM", "STTM", "UTMALDG"):
        assert mnemonic in sass, mnemonic
    assert "HGMMA" not in sass


def test_gemm_issue_loops_stay_on_the_uniform_datapath():
    """Regression guard for a silent 8 % slowdown: when the TMA-producer / MMA-issuer loops of gemm_tc_kernel are compiled as
    single-lane (divergent) code, every tcgen05.mma / TMA operand goes through R2UR + a BRA.U.ANY waterfall loop (115 SASS
    instructions per k-block, ncu: issuing thread busy 82 %, tensor pipe 80 % active); warp-uniform loops with one elected
    lane have none (52 instructions, tensor pipe 94 %; profiles/r1_ncu_gemm_vs_cublas_probe.txt)."""
    import subprocess
    obj = os.path.join(ROOT, "reptext_b200", "csrc", "build", "gemm_sm100.o")
    from reptext_b200 import build
    build.build(force=False, verbose=False)
    sass = subprocess.run(["cuobjdump", "-sass", obj], capture_output=True, text=True).stdout
    funcs = [f for f in re.split(r"\n\s*Function : ", sass)[1:] if "gemm_tc_kernel" in f.split("\n")[0]]
    assert len(funcs) >= 5
    for f in funcs:
        assert f.count("UTCHMMA") == 4 and "UTMALDG" in f
        assert "BRA.U.ANY" not in f, f.split("\n")[0]
        assert f.count("R2UR") <= 72, (f.split("\n")[0], f.count("R2UR"))    # 32-49 today; the single-lane form had 138-210


def test_no_product_module_imports_the_oracle():
    pkg = os.path.join(ROOT, "reptext_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith(".py"):
                src = open(os.path.join(dirpath, f)).read()
                assert "oracle" not in re.sub(r'"""[\s\S]*?"""|#.*', "", src), f
