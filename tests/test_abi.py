"""The C-ABI shared library loads and exports every symbol include/bvh_cuda.h declares; struct layouts match;
the host mirror refuses to run without the CUDA path (no CPU fallback).  No compute calls here."""
import ctypes as C
import os
import re

import numpy as np
import pytest

import voidin_b200 as vb
from voidin_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_functions(header="bvh_cuda.h"):
    src = open(os.path.join(ROOT, "include", header)).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(bvh_cuda_[a-z_0-9]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    lib = _lib.load()
    names = _declared_functions()
    assert len(names) >= 20
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/bvh_cuda.h but not exported"
    assert sorted(_lib.SYMBOLS) == names
    assert lib.bvh_cuda_abi_version() == 5
    model_names = _declared_functions("bvh_cuda_models.h")
    assert len(model_names) == 10
    for n in model_names:
        assert hasattr(lib, n), f"{n} declared in include/bvh_cuda_models.h but not exported"


def test_struct_layouts_match_the_reference():
    assert vb.BVH_NODE.itemsize == 32 and vb.TLAS_NODE.itemsize == 32
    assert vb.INSTANCE.itemsize == 144 and vb.MESH_INFO.itemsize == 48
    assert vb.BVH_NODE.fields["left_first"][1] == 12 and vb.BVH_NODE.fields["count"][1] == 28
    assert vb.TLAS_NODE.fields["left_right"][1] == 12 and vb.TLAS_NODE.fields["instance_idx"][1] == 28
    assert vb.INSTANCE.fields["inv_transform"][1] == 64 and vb.INSTANCE.fields["mesh"][1] == 128
    assert vb.MESH_INFO.fields["vertex_offset"][1] == 32 and vb.MESH_INFO.fields["bvh_index"][1] == 36
    assert C.sizeof(_lib.BuildStats) == 104


def test_null_context_calls_are_rejected_not_crashing():
    lib = _lib.load()
    assert lib.bvh_cuda_blas_build(None, None, 0, None, 0, None, 0, None) == -1
    assert lib.bvh_cuda_tlas_build(None, None, 0, None, 0, None, None) == -1
    assert lib.bvh_cuda_launch_count(None) == 0
    lib.bvh_cuda_destroy(None)


def test_no_cpu_fallback_without_a_device():
    import torch

    if torch.cuda.is_available():
        pytest.skip("a CUDA device is present")
    with pytest.raises(vb.BvhCudaError):
        vb.Context(0)


def test_builder_argument_checks_mirror_the_reference():
    v = np.zeros((3, 3), np.float32)
    with pytest.raises(TypeError):
        vb.BvhBuilder(v, [0, 1, 2], ctx=object())  # not an in-place-permutable u32 buffer
    with pytest.raises(vb.BvhCudaError):
        vb.BvhBuilder(v, np.zeros(4, np.uint32), ctx=object())  # len % 3 != 0 (mesh/mod.rs:321)


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "voidin_b200")
    pat = re.compile(r"(^\s*(from|import)\s+oracle\b|libbvh_oracle|#include\s+\"[^\"]*oracle)", re.M)
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp")):
                assert not pat.search(open(os.path.join(dirpath, f)).read()), f


def test_headers_compile_as_c_and_cxx(tmp_path):
    """include/bvh_cuda.h is valid C11 and C++17; the header-only C++ mirror compiles and links against the library
    (no device calls: without a GPU Context() must throw, with one it must construct)."""
    import subprocess

    inc = os.path.join(ROOT, "include")
    c_src = tmp_path / "t.c"
    c_src.write_text('#include "bvh_cuda.h"\n#include "bvh_cuda_models.h"\nint main(void){ return sizeof(BvhNode)==32 && sizeof(TlasNode)==32 && sizeof(Instance)==144 && sizeof(MeshInfo)==48 ? 0 : 1; }\n')
    subprocess.check_call(["gcc", "-std=c11", "-Wall", "-Werror", "-I", inc, str(c_src), "-o", str(tmp_path / "t_c")])
    assert subprocess.call([str(tmp_path / "t_c")]) == 0
    cxx_src = tmp_path / "t.cpp"
    (tmp_path / "tri.obj").write_text("v 0 0 0\nv 1 0 0\nv 0 1 0\nv 1 1 0\nf 1 2 3 4\n")
    cxx_src.write_text(
        '#include "bvh_cuda.hpp"\n#include <cstdio>\n'
        "int main(int argc, char** argv){\n"
        "  bvh_cuda::Model m = bvh_cuda::ObjModel::import(argv[1]);  // host only: works without a GPU\n"
        "  if (m.meshes.size() != 1 || m.meshes[0].vertices.size() != 12 || m.meshes[0].indices.size() != 6) return 2;\n"
        "  try { bvh_cuda::GltfDocument::import(\"/nonexistent.glb\"); return 3; } catch (const bvh_cuda::Error&) {}\n"
        "  try { bvh_cuda::Context c(0); std::puts(\"ctx\");\n"
        "        bvh_cuda::Mesh& q = m.meshes[0];\n"
        "        bvh_cuda::Bvh b = bvh_cuda::BvhBuilder(c, q.vertices.data(), 4, q.indices.data(), 2).set_bin_number(8).build();\n"
        "        std::vector<bvh_cuda::Ray> rays(1); rays[0] = bvh_cuda::Ray{{0.25f, 0.25f, 1.0f}, {0.0f, 0.0f, -1.0f}};\n"
        "        (void)b.traverse_iter(c, q.vertices.data(), 4, q.indices.data(), 2, rays);  // parity is the GPU suite's job\n"
        "  } catch (const bvh_cuda::Error& e) { std::puts(\"nogpu\"); }\n"
        " return bvh_cuda_abi_version() == 5 ? 0 : 1; }\n")
    libdir = os.path.join(ROOT, "voidin_b200")
    subprocess.check_call(["g++", "-std=c++17", "-Wall", "-I", inc, str(cxx_src), "-o", str(tmp_path / "t_cxx"), "-L", libdir,
                           "-lbvh_cuda", f"-Wl,-rpath,{libdir}"])
    out = subprocess.run([str(tmp_path / "t_cxx"), str(tmp_path / "tri.obj")], capture_output=True, text=True)
    assert out.returncode == 0 and out.stdout.strip() in ("ctx", "nogpu")
