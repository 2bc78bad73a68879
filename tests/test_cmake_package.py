"""The CMake package (CMakeLists.txt + cmake/vdb_mappingConfig.cmake.in): builds the CUDA library with nvcc (cross-compile,
no GPU needed), installs it, and a consumer that only says find_package(vdb_mapping) + vdb_mapping::vdb_mapping — the
reference's own package and target names (/root/reference/CMakeLists.txt:25,49) — compiles and links against it."""
import os
import shutil
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.skipif(shutil.which("cmake") is None or shutil.which("ninja") is None or not os.path.exists("/usr/local/cuda/bin/nvcc"),
                    reason="needs cmake, ninja and nvcc")
def test_package_installs_and_a_reference_style_consumer_links(tmp_path):
    build, prefix, cbuild = tmp_path / "build", tmp_path / "prefix", tmp_path / "consumer"

    def run(*cmd):
        p = subprocess.run(cmd, capture_output=True, text=True, timeout=900)
        assert p.returncode == 0, " ".join(cmd) + "\n" + p.stdout[-3000:] + p.stderr[-3000:]
        return p.stdout

    run("cmake", "-S", ROOT, "-B", str(build), "-G", "Ninja", "-DCMAKE_CUDA_COMPILER=/usr/local/cuda/bin/nvcc", "-DVDBM_BUILD_TESTS=OFF")
    cmds = run("ninja", "-C", str(build), "-t", "commands", "vdbm_b200")
    assert "-fmad=false" in cmds and "sm_100a" in cmds and "-fvisibility=hidden" in cmds
    run("cmake", "--build", str(build))
    run("cmake", "--install", str(build), "--prefix", str(prefix))
    for f in ("include/vdbm_b200.h", "include/vdb_mapping/VDBMapping.hpp", "include/vdb_mapping/OccupancyVDBMapping.hpp",
              "lib/libvdbm_b200.so", "lib/cmake/vdb_mapping/vdb_mappingConfig.cmake", "lib/cmake/vdb_mapping/vdb_mappingTargets.cmake"):
        assert (prefix / f).exists(), f
    run("cmake", "-S", os.path.join(ROOT, "tests", "cmake_consumer"), "-B", str(cbuild), "-G", "Ninja", f"-DCMAKE_PREFIX_PATH={prefix}")
    run("cmake", "--build", str(cbuild))
    exe = cbuild / "consumer"
    assert exe.exists()
    needed = subprocess.run(["readelf", "-d", str(exe)], capture_output=True, text=True).stdout
    assert "libvdbm_b200.so" in needed
