"""The C-ABI shared library loads and exports every symbol include/mppgpu.h declares (no compute calls: no GPU here)."""
import ctypes as C
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "mppgpu.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(mppgpu_[a-z_0-9]+)\s*\(", text)))


def test_header_declares_the_python_binding_exactly():
    from mpp_b200 import _lib
    assert sorted(_lib.EXPORTS) == declared_symbols()


def test_library_exports_every_declared_symbol():
    from mpp_b200 import _lib
    assert os.path.exists(_lib.LIB_PATH), "build libmppgpu.so first: python -c 'import __graft_entry__ as g; g.build()'"
    L = C.CDLL(_lib.LIB_PATH)
    for name in declared_symbols():
        assert hasattr(L, name), name
    L.mppgpu_version.restype = C.c_int
    assert L.mppgpu_version() >= 100


def test_no_cpu_fallback_without_a_device():
    """Without a CUDA device the library must fail loudly, never compute on the CPU."""
    import mpp_b200
    from mpp_b200._lib import lib
    if lib().mppgpu_device_count() > 0:
        pytest.skip("a GPU is present")
    with pytest.raises(mpp_b200.MPPError) as e:
        mpp_b200.VSFM(4, 15)
    assert "no CPU fallback" in str(e.value)


def test_product_never_imports_the_oracle():
    """oracle/ is test infrastructure: nothing under mpp_b200/ may reference it."""
    for dirpath, _, files in os.walk(os.path.join(ROOT, "mpp_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".inl", ".cpp", ".h")):
                text = open(os.path.join(dirpath, f)).read()
                assert "mpp_oracle" not in text and "import oracle" not in text and "from oracle" not in text, f


def _build_c_smoke(tmp_path):
    import subprocess
    exe = str(tmp_path / "cabi_smoke")
    libdir = os.path.join(ROOT, "mpp_b200")
    subprocess.check_call(["gcc", "-std=c99", "-Wall", "-Wextra", "-Werror", "-pedantic", "-I" + os.path.join(ROOT, "include"),
                           os.path.join(ROOT, "tests", "cabi_smoke.c"), "-o", exe, "-L" + libdir, "-lmppgpu", "-lm", "-Wl,-rpath," + libdir])
    return exe


def test_header_is_valid_c99_and_a_plain_c_program_links(tmp_path):
    """include/mppgpu.h compiled as C (gcc -std=c99 -pedantic -Werror) and linked against libmppgpu.so; without a GPU the program only
    loads the library and reports version / device count."""
    import subprocess
    out = subprocess.check_output([_build_c_smoke(tmp_path)]).decode()
    assert "mppgpu version" in out


@pytest.mark.gpu
def test_plain_c_program_steps_four_columns(tmp_path):
    """create -> set_mesh -> add_condition -> set_soils -> restart -> set_data -> PreStepDT -> StepDT -> get_data -> global mass balance
    from C99, 4 columns x 15 layers; the program checks convergence and the 1e-5 kg mass-balance gate itself."""
    import subprocess
    r = subprocess.run([_build_c_smoke(tmp_path), "run"], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    assert r.stdout.strip().endswith("ok")


def test_fortran_binding_covers_every_entry_point_and_is_current():
    """include/mppgpu_binding.F90 is generated from the C prototypes (tools/gen_fortran_binding.py): one bind(C) interface per entry point
    of include/mppgpu.h with as many dummy arguments as the C function has parameters, the three structs as bind(C) derived types with
    the header's member order, and the committed file is what the generator writes today."""
    import importlib.util
    import re
    spec = importlib.util.spec_from_file_location("gen_fortran_binding", os.path.join(ROOT, "tools", "gen_fortran_binding.py"))
    g = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(g)
    src, names = g.generate()
    assert open(os.path.join(ROOT, "include", "mppgpu_binding.F90")).read() == src, "run python tools/gen_fortran_binding.py"
    assert set(names) == set(declared_symbols())
    hdr = open(os.path.join(ROOT, "include", "mppgpu.h")).read()
    for ret, name, args in g.prototypes(hdr):
        nargs = 0 if args in ("", "void") else len(args.split(","))
        m = re.search(r"function %s\(([^)]*)\)\s*(?:&\s*)?bind\(C" % name, src.replace("&\n          ", ""))
        assert m, name
        dummies = [a for a in m.group(1).replace("&", "").split(",") if a.strip()]
        assert len(dummies) == nargs, (name, dummies, nargs)
    from mpp_b200 import _lib
    for sname, cls in (("mppgpu_xfer", _lib.Xfer), ("mppgpu_elm_columns", _lib.ElmColumns), ("mppgpu_elm_thermal_columns", _lib.ElmThermalColumns)):
        fields = dict(g.structs(hdr))[sname]
        assert [f[0] for f in fields] == [f[0] for f in cls._fields_], sname          # header order == ctypes mirror == Fortran type
        block = src[src.index("type, bind(C), public :: %s" % sname):src.index("end type %s" % sname)]
        assert [ln.split("::")[1].strip() for ln in block.splitlines()[1:] if "::" in ln] == [f[0] for f in fields]
