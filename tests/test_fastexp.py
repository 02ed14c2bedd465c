"""The branch-free exponentials of tapir_b200/csrc/fastexp.cuh, compiled for the host and swept against
long double expl (2M arguments over [-760, 720] plus the special values)."""
import os
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_fastexp_accuracy(tmp_path):
    exe = str(tmp_path / "fastexp_check")
    subprocess.check_call(["g++", "-O2", "-mfma", "-std=c++17", "-o", exe,
                           os.path.join(ROOT, "tests", "native", "fastexp_check.cpp")])
    out = subprocess.run([exe], capture_output=True, text=True, check=True).stdout
    vals = dict(line.split() for line in out.strip().splitlines() if len(line.split()) == 2)
    assert float(vals["exp_bounded"]) <= 1.0          # ulp
    assert float(vals["exp_full"]) <= 1.1             # Estrin form
    assert float(vals["exp_pm_minus"]) <= 2.0 and float(vals["exp_pm_plus"]) <= 2.0   # even/odd split: 2 ulp
    assert int(vals["specials_bad"]) == 0
    assert int(vals["exp_pm_small_diff"]) == 0        # the reduction-free pair: bit-identical below ln(2)/2
