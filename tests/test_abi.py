"""The C ABI as a C compiler sees it: tests/abi/abi_smoke.c is built with gcc against include/tlab_gpu.h only (-Wall -Wextra
-Werror) and linked to libtlab_gpu.so, so a prototype that drifts from the library or a header that stops being plain C fails
here without Python in between.  Without a device the program must stop at tlab_gpu_init with the "no CPU fallback" error
(exit code 77); on a GPU box it runs one CFL-controlled RK4-5 step from host arrays."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _build(tmp_path):
    from tlab_b200 import build
    lib = build.build()
    exe = str(tmp_path / "abi_smoke")
    cmd = ["gcc", "-std=c11", "-Wall", "-Wextra", "-Werror", "-I", os.path.join(ROOT, "include"),
           os.path.join(ROOT, "tests", "abi", "abi_smoke.c"), "-o", exe, "-L", os.path.dirname(lib), "-ltlab_gpu", "-lm",
           "-Wl,-rpath," + os.path.dirname(lib)]
    r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    assert r.returncode == 0, r.stdout
    return exe


def test_c_host_compiles_links_and_refuses_to_run_without_a_device(tmp_path):
    import torch
    exe = _build(tmp_path)
    r = subprocess.run([exe], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=300)
    if torch.cuda.is_available():
        assert r.returncode == 0 and "ABI_SMOKE_OK" in r.stdout, r.stdout
    else:
        assert r.returncode == 77 and "no CPU fallback" in r.stdout, r.stdout


@pytest.mark.gpu
def test_c_host_runs_one_rk_step(tmp_path, cuda):
    exe = _build(tmp_path)
    r = subprocess.run([exe], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=300)
    assert r.returncode == 0 and "ABI_SMOKE_OK" in r.stdout, r.stdout


def test_fortran_module_binds_every_export_with_the_right_arity():
    """fortran/tlab_gpu_mod.f90 (module TLab_GPU_C) against include/tlab_gpu.h: every prototype has an interface with the same
    name and number of arguments, and nothing is bound that the header does not declare (no Fortran compiler in the image,
    so this is a textual check; INTEGRATION.md gives the gfortran -fsyntax-only recipe)."""
    import re
    from tlab_b200 import lib as tl
    protos = tl.parse_header()
    src = open(os.path.join(ROOT, "fortran", "tlab_gpu_mod.f90")).read()
    src = re.sub(r"&\s*\n\s*", " ", src)                     # join continuation lines
    bound = {}
    for m in re.finditer(r"function\s+(\w+)\s*\(([^)]*)\)\s*bind\(C,\s*name='(\w+)'\)", src, flags=re.I):
        fname, args, cname = m.group(1), m.group(2), m.group(3)
        assert fname == cname, (fname, cname)
        bound[cname] = len([a for a in args.split(",") if a.strip()])
    missing = sorted(set(protos) - set(bound))
    extra = sorted(set(bound) - set(protos))
    assert not missing, "exports without a Fortran interface: %s" % missing
    assert not extra, "Fortran interfaces without a C prototype: %s" % extra
    wrong = {n: (bound[n], len(protos[n][1])) for n in protos if bound[n] != len(protos[n][1])}
    assert not wrong, "argument counts differ (fortran, C): %s" % wrong
    # the signature-exact wrappers named in the module header exist
    for name in ["OPR_Partial_X_GPU", "OPR_Partial_Y_GPU", "OPR_Partial_Z_GPU", "OPR_Burgers_X_GPU", "OPR_Burgers_Y_GPU",
                 "OPR_Burgers_Z_GPU", "OPR_Poisson_GPU", "BOUNDARY_BCS_NEUMANN_Y_GPU", "FDM_Der1_Solve_GPU", "TRIDSS_GPU",
                 "TRIDPSS_GPU", "PENTADSS_GPU", "TLab_Transpose_GPU", "OPR_Fourier_X_Forward_GPU", "TIME_RUNGEKUTTA_GPU"]:
        assert re.search(r"subroutine\s+%s\s*\(" % name, src), name
    # OPR_Poisson_GPU keeps the dummy list of OPR_Poisson_interface (opr_elliptic.f90:34)
    assert re.search(r"subroutine OPR_Poisson_GPU\(nx, ny, nz, ibc, p, tmp1, tmp2, bcs_hb, bcs_ht, dpdy\)", src)
    assert re.search(r"subroutine OPR_Partial_X_GPU\(type, nx, ny, nz, bcs, g, u, result, tmp1\)", src)
    assert re.search(r"subroutine OPR_Burgers_X_GPU\(ivel, is, nx, ny, nz, bcs, s, u, result, tmp1, u_t\)", src)
