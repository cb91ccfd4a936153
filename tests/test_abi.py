"""The C-ABI library loads, exports every symbol include/insta_b200.h declares, and refuses to run without a GPU."""
import os
import re
import subprocess

import pytest

from insta_b200 import capi

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")


def header_symbols():
    src = open(os.path.join(ROOT, "include", "insta_b200.h")).read()
    return sorted(set(re.findall(r"INSTA_API[^;(]*?\b(insta_\w+)\s*\(", src)))


def test_header_matches_exports():
    lib = capi.load_library()
    syms = header_symbols()
    assert len(syms) >= 30
    assert sorted(capi.EXPORTS) == syms
    for s in syms:
        assert hasattr(lib, s), s
    out = subprocess.check_output(["nm", "-D", "--defined-only", capi.LIB_PATH], text=True)
    exported = set(re.findall(r" T (insta_\w+)", out))
    assert set(syms) <= exported
    # nothing but the ABI leaks out of the shared object
    assert all(s.startswith("insta_") for s in re.findall(r" T (\w+)", out) if not s.startswith("_"))


def test_library_carries_sm100a_sass_only():
    out = subprocess.run(["cuobjdump", "-lelf", capi.LIB_PATH], capture_output=True, text=True).stdout
    archs = set(re.findall(r"sm_(\d+a?)", out))
    assert archs == {"100a"}, archs


def test_params_default():
    p = capi.default_params()
    assert p.struct_size == __import__("ctypes").sizeof(capi.Params)
    assert (p.n_levels, p.n_features_per_level, p.log2_hashmap_size, p.base_resolution) == (16, 8, 18, 16)
    assert p.max_cached_bvh == 4000 and p.exp_params == 16 and p.n_primitives_per_leaf == 2 and p.rng_seed == 42


def test_no_cpu_fallback():
    import torch

    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(capi.InstaError) as e:
        capi.Context(device=0)
    assert e.value.code == -4  # INSTA_ERR_NO_DEVICE


def test_product_never_imports_oracle():
    """The oracle is test infrastructure: no file of the product package may import, link or mention it."""
    pkg = os.path.join(ROOT, "insta_b200")
    for dp, _, fs in os.walk(pkg):
        for f in fs:
            if f.endswith((".py", ".cu", ".cpp", ".h")):
                assert "oracle" not in open(os.path.join(dp, f), errors="ignore").read().lower(), f


def test_hooks_adapter_compiles_and_links(tmp_path):
    """include/insta_b200_hooks.hpp (the C++ adapter a maintainer includes from src/core.cu) must compile as plain C++14 -- the reference's standard --
    and link against the library: a snapshot round trip through it needs no GPU."""
    import shutil
    import subprocess

    cxx = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else shutil.which("g++")
    if not cxx:
        pytest.skip("no host C++ compiler")
    src = tmp_path / "use_hooks.cpp"
    src.write_text(r"""
#include <insta_b200_hooks.hpp>
#include <cstdio>
#include <vector>
int main(int argc, char** argv) {
	std::vector<unsigned short> params(1024, 0x3c00);
	insta_snapshot s{};
	s.n_params = params.size();
	s.h_params_f16 = params.data();
	s.training_step = 42;
	s.aabb[3] = s.aabb[4] = s.aabb[5] = 1.f;
	s.render_aabb[3] = s.render_aabb[4] = s.render_aabb[5] = 1.f;
	insta::Hooks::save_snapshot(argv[1], nullptr, s);
	auto l = insta::Hooks::load_snapshot(argv[1]);
	if (l.data.n_params != 1024 || l.data.training_step != 42) return 1;
	try {
		insta::Hooks::load_snapshot("/nonexistent/file");
		return 2;
	} catch (const std::runtime_error& e) {
		std::printf("%s\n", e.what());
	}
	insta::Hooks h;  // no context: never created without a GPU
	return h.ctx() == nullptr ? 0 : 3;
}
""")
    exe = tmp_path / "use_hooks"
    lib_dir = os.path.dirname(capi.LIB_PATH)
    subprocess.check_call([cxx, "-std=c++14", "-Wall", "-Werror", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe), "-L", lib_dir, "-linsta_b200",
                           "-Wl,-rpath," + lib_dir])
    out = subprocess.run([str(exe), str(tmp_path / "s.msgpack")], capture_output=True, text=True)
    assert out.returncode == 0, out.stdout + out.stderr
    assert "cannot open" in out.stdout
