"""CPU-side checks of the drop-in boundary: the C-ABI library loads and exports every symbol that
include/vdbm_b200.h declares; no compute call is made (there is no GPU here)."""
import ctypes
import os
import re
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "vdbm_b200.h")


def declared_functions():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(vdbm_[a-z0-9_]+)\s*\(", src)))


@pytest.fixture(scope="module")
def so_path():
    from vdb_mapping_b200 import build
    return build.build_lib()


def test_header_declares_the_hot_path():
    names = declared_functions()
    for must in ["vdbm_create", "vdbm_destroy", "vdbm_set_config", "vdbm_source_add", "vdbm_accumulate", "vdbm_integrate",
                 "vdbm_insert", "vdbm_update_map", "vdbm_update_export", "vdbm_update_import", "vdbm_map_export",
                 "vdbm_section", "vdbm_probe", "vdbm_reset", "vdbm_stats", "vdbm_last_error"]:
        assert must in names


def test_library_exports_every_declared_symbol(so_path):
    lib = ctypes.CDLL(so_path)
    missing = [n for n in declared_functions() if not hasattr(lib, n)]
    assert not missing, missing
    assert lib.vdbm_abi_version() == 1


def test_header_is_plain_c(tmp_path):
    """The boundary is a C ABI: include/vdbm_b200.h compiles as C99 (-pedantic, warnings as errors) and a C program links it."""
    src = tmp_path / "c_abi.c"
    src.write_text('#include "vdbm_b200.h"\n'
                   'static int sink(void* u, uint64_t n, const uint32_t* i, const int32_t* o, const float* v, const uint64_t* a)\n'
                   '{ (void)u; (void)n; (void)i; (void)o; (void)v; (void)a; return 0; }\n'
                   'int main(void) { vdbm_mirror_sink s = sink; vdbm_params p; p.resolution = 0.1; return (s != 0 && p.resolution > 0 && vdbm_abi_version() == VDBM_ABI_VERSION) ? 0 : 1; }\n')
    subprocess.run(["gcc", "-std=c99", "-pedantic", "-Wall", "-Wextra", "-Werror", "-fsyntax-only", "-I", os.path.join(ROOT, "include"), str(src)], check=True)


def test_only_the_abi_is_exported(so_path):
    out = subprocess.run(["nm", "-D", "--defined-only", so_path], capture_output=True, text=True, check=True).stdout
    syms = [l.split()[-1] for l in out.splitlines() if " T " in l]
    assert syms and all(s.startswith("vdbm_") for s in syms), [s for s in syms if not s.startswith("vdbm_")]


def test_sass_is_sm100a_with_blackwell_256bit_accesses(so_path):
    out = subprocess.run(["cuobjdump", "-sass", so_path], capture_output=True, text=True).stdout
    assert "sm_100a" in out or "SM100a" in out.replace("_", "").upper() or "arch = sm_100" in out
    assert ".256" in out          # LDG/STG.E.ENL2.256 in the update / gather kernels
    assert "RED" in out           # red.global.or.b64 in the DDA kernel
    # -fmad=false: no fp64 contraction anywhere near the voxel paths. The ONLY DFMAs of the DDA kernel are its three
    # "next[axis] += delta[axis]" selects per step: fma(delta, 1.0 or 0.0, next), exact by construction
    dda = out.split("raycast_dda_kernelILi0E")[1].split("Function :")[0]
    steps = dda.count("DSETP.GEU")  # one MinIndex evaluation per unrolled step
    assert steps >= 2 and dda.count("DFMA") == 3 * steps and "DMUL" not in dda and "DADD" not in dda
    # (prep_rays / wall_dda contain DFMA too: inside the correctly rounded software division / sqrt / fmod sequences)


def test_create_fails_loudly_without_gpu(so_path):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from vdb_mapping_b200.mapping import OccupancyVDBMapping, VdbmError
    with pytest.raises(VdbmError):
        OccupancyVDBMapping(0.1)


def test_leaf_owner_is_a_pure_host_function(so_path):
    from vdb_mapping_b200.mapping import leaf_owner
    owners = [leaf_owner([8 * x, 8 * y, 0], 8) for x in range(-8, 8) for y in range(-8, 8)]
    assert set(owners) <= set(range(8)) and len(set(owners)) == 8
    # 2x2x2 bricks of leaves share an owner
    assert leaf_owner([0, 0, 0], 8) == leaf_owner([8, 8, 8], 8)
    assert leaf_owner([16, 0, 0], 4) == leaf_owner([24, 8, 8], 4)


def test_product_does_not_touch_the_oracle():
    """The product package must never import / link / execute anything under oracle/."""
    pkg = os.path.join(ROOT, "vdb_mapping_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".hpp", ".cpp")):
                txt = open(os.path.join(dirpath, f), errors="ignore").read()
                assert "oracle" not in txt.lower().replace("no cpu fallback", ""), os.path.join(dirpath, f)
