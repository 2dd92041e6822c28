"""Accuracy of the float64 power function used by the CUDA kernels (lisflood_code_b200/csrc/lf_math.cuh),
checked on the HOST: the header compiles as plain C++ (same arithmetic, fma() instead of __fma_rn)."""
import os
import re
import subprocess

from conftest import ROOT


def test_pw_accuracy_against_libm(tmp_path):
    exe = str(tmp_path / "lf_math_host_test")
    subprocess.check_call(["/usr/bin/g++", "-O2", "-ffp-contract=off", "-I", os.path.join(ROOT, "lisflood_code_b200", "csrc"),
                           "-o", exe, os.path.join(ROOT, "tests", "lf_math_host_test.cpp")])
    out = subprocess.check_output([exe], text=True)
    m = re.search(r"max rel err ([0-9.e+-]+) .*>1e-13: (\d+)", out)
    assert m, out
    # |y log2 x| reaches ~200 in the sampled range: 200 * 2^-52 = 4.4e-14
    assert float(m.group(1)) < 6e-14 and int(m.group(2)) == 0
    m = re.search(r"log2 max rel ([0-9.e+-]+) exp2 max rel ([0-9.e+-]+)", out)
    assert float(m.group(1)) < 1e-15 and float(m.group(2)) < 5e-16
    # special values follow pow() for x >= 0
    for line in out.splitlines():
        mm = re.match(r"pw\((\S+),(\S+)\)=(\S+) pow=(\S+)", line)
        if mm and mm.group(1) not in ("-1", "nan", "4.94066e-324", "1e-310"):   # denormal bases flush to 1e-300
            assert mm.group(3).lstrip("-") == mm.group(4).lstrip("-"), line
    # table-driven variants used by the soil column kernel: error grows with |y log2 x| like any 2^(y log2 x)
    m = re.search(r"pw_tab max rel err ([0-9.e+-]+) .* normalised by \(1\+\|y log2 x\|\) ([0-9.e+-]+) ; >1e-13: (\d+)", out)
    assert m, out
    assert float(m.group(2)) < 4e-16
    m = re.search(r"log2_tab max err/\(1\+\|log2 x\|\) ([0-9.e+-]+) ; exp_neg_tab max rel/\(1\+\|x\|\) ([0-9.e+-]+)", out)
    assert m and float(m.group(1)) < 4e-16 and float(m.group(2)) < 4e-16, out
    assert "pw_tab(0,0.3)=0 pw_tab(1,7.5)=1 pw_tab(1e-200,9)=0" in out and "exp_neg_tab(0)=1 exp_neg_tab(-1000)=0" in out
