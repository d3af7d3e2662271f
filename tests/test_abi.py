"""CPU: the C-ABI library loads and exports every symbol include/rad_cuda.h declares; without a GPU the entry
point fails loudly (no CPU fallback); the product never touches oracle/."""
import ctypes
import os
import re
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    text = open(os.path.join(ROOT, "include", "rad_cuda.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(rad_[a-z0-9_]+)\s*\(", text)))


def test_header_symbols_exported(api):
    names = _declared_symbols()
    assert len(names) >= 30
    lib = ctypes.CDLL(os.path.join(ROOT, "radiosity_b200", "librad_cuda.so"))
    missing = [n for n in names if not hasattr(lib, n)]
    assert not missing, missing
    # and the Python binding table covers exactly the header
    assert sorted(api.CUDA_SIGNATURES) == names


def test_struct_layout(api):
    assert ctypes.sizeof(api.RadConfig) == 4 * 6 + 64 + 4
    assert ctypes.sizeof(api.RadStats) == 32


def test_no_cpu_fallback_without_gpu(api):
    try:
        import torch
        if torch.cuda.is_available():
            pytest.skip("GPU present")
    except ImportError:
        pass
    with pytest.raises(api.RadError) as e:
        api.Context(32, 1, 100)
    assert "no CUDA device" in str(e.value) or "CUDA" in str(e.value)


def test_bad_arguments_rejected_before_touching_cuda(api):
    lib = api.cuda_lib()
    cfg = api.RadConfig()
    h = ctypes.c_void_p()
    cfg.hemicube_side, cfg.hemicubes, cfg.max_patches = 30, 1, 10          # not a multiple of 16
    assert lib.rad_create(ctypes.byref(h), ctypes.byref(cfg)) == -1
    assert b"hemicube_side" in lib.rad_last_error(None)
    cfg.hemicube_side, cfg.hemicubes = 32, 65
    assert lib.rad_create(ctypes.byref(h), ctypes.byref(cfg)) == -1
    assert lib.rad_create(None, None) == -1
    assert lib.rad_version().startswith(b"radiosity_b200")


def test_product_does_not_reference_the_oracle():
    """oracle/ is test infrastructure: nothing under radiosity_b200/ or include/ may mention or link it."""
    bad = []
    for base in ("radiosity_b200", "include"):
        for dp, _, files in os.walk(os.path.join(ROOT, base)):
            for f in files:
                if f.endswith((".so", ".o", ".pyc")) or f == "radiosity":
                    continue
                txt = open(os.path.join(dp, f), errors="ignore").read()
                # comments may cite the oracle / the reference; code may not include, import, link or open them
                if re.search(r"#\s*include[^\n]*oracle|import\s+oracle|from\s+oracle|liboracle|libref_host|orc_[a-z_]+\s*\(|refp_[a-z_]+\s*\(", txt):
                    bad.append(os.path.join(dp, f))
    assert not bad, bad
    for so in ("librad_cuda.so", "libradiosity_host.so"):
        out = subprocess.run(["ldd", os.path.join(ROOT, "radiosity_b200", so)], capture_output=True, text=True).stdout
        assert "oracle" not in out and "ref_host" not in out


def test_headless_driver_fails_loudly_without_gpu():
    try:
        import torch
        if torch.cuda.is_available():
            pytest.skip("GPU present")
    except ImportError:
        pass
    exe = os.path.join(ROOT, "radiosity_b200", "radiosity")
    p = subprocess.run([exe, "area", "0.5", "hemicube", "32", "hemicubes", "1", "shots", "2"], capture_output=True, text=True)
    assert p.returncode != 0 and "no CUDA device" in p.stderr
    p = subprocess.run([exe, "area"], capture_output=True, text=True)       # odd argument count, as the reference rejects it
    assert p.returncode != 0


def test_header_is_plain_c99(tmp_path):
    """include/rad_cuda.h is the drop-in boundary: it must compile as C (no C++, no CUDA, no torch types) and every
    declared function must resolve against librad_cuda.so at link time."""
    import re
    import subprocess
    hdr = open(os.path.join(ROOT, "include", "rad_cuda.h")).read()
    names = sorted(set(re.findall(r"\b(rad_[a-z0-9_]+)\s*\(", hdr)))
    assert len(names) >= 30
    src = tmp_path / "use.c"
    body = "\n".join(f"\tp[{i}] = (void*){n};" for i, n in enumerate(names))
    src.write_text('#include "rad_cuda.h"\n#include <stdio.h>\nint main(void) {\n\tvoid* p[%d];\n%s\n\tprintf("%%p %%s\\n", p[0], rad_version());\n\treturn 0;\n}\n' % (len(names), body))
    exe = tmp_path / "use"
    lib_dir = os.path.join(ROOT, "radiosity_b200")
    r = subprocess.run(["gcc", "-std=c99", "-pedantic", "-Wall", "-Werror", "-Wno-pedantic", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe),
                        "-L", lib_dir, "-lrad_cuda", f"-Wl,-rpath,{lib_dir}"], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-3000:]
    r = subprocess.run([str(exe)], capture_output=True, text=True)
    assert r.returncode == 0 and "radiosity_b200" in r.stdout
