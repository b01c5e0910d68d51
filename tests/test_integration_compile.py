"""CPU: the reference-side bindings of INTEGRATION.md are real code -- host/external/b200_adapters.hpp (a
CBinaryEncoder + CBinarySoftDecoder subclass and a CKernProcLLR subclass over the C ABI) compiles against the reference's
OWN headers (headers/external/Codec.h, KernProc.h, Kernel.h through the oracle's MSVC-compat shim) and links with
libpkb200.so.  Needs /root/reference (build container only)."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = "/root/reference"


@pytest.mark.skipif(not os.path.isdir(os.path.join(REF, "headers", "external")), reason="reference headers not present")
def test_adapters_compile_against_reference_headers(tmp_path):
    pkg = os.path.join(ROOT, "polar-codes-with-bch-kernel_b200")
    src = tmp_path / "use.cpp"
    src.write_text('''
#include "b200_adapters.hpp"
#include <fstream>
int build_only(const char *spec_file, const CBinaryKernel &K) {
    std::ifstream f(spec_file);
    CB200ListDecoder dec(f, 8);            // drop-in for CMixedKernelListDecoder(std::istream&, unsigned)
    CBinarySoftDecoder &soft = dec;        // the simulator only sees the base classes (Simulator.cpp:139)
    CBinaryEncoder &enc = dec;
    CKanekoKernelProcessor proc(K);
    const CKernProcLLR &p = proc;
    return (int)(soft.GetMaxListSize() + enc.GetLength() + p.Size());
}
''')
    shim = os.path.join(ROOT, "oracle", "polar_shim")
    cmd = ["/usr/bin/g++", "-std=c++17", "-fpermissive", "-w", "-msse4.2", "-mpopcnt", "-fPIC", "-include", os.path.join(shim, "compat.h"),
           "-I", shim, "-I", os.path.join(REF, "headers", "external"), "-I", os.path.join(pkg, "host", "external"), "-I", os.path.join(ROOT, "include"),
           "-shared", "-o", str(tmp_path / "libuse.so"), str(src), "-L", pkg, "-lpkb200"]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr[-3000:]
    nm = subprocess.run(["nm", "-DC", str(tmp_path / "libuse.so")], capture_output=True, text=True).stdout
    assert "pk_polar_decode_batch" in nm and "pk_kproc_get_llrs" in nm and "build_only" in nm
