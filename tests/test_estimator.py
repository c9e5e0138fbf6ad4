"""CPU tests of the host-side CF estimator: fop_fo_estimate = the reference's
generateFOEstimate / getOptionSpline / monotone spline (OptionCalibration.h:26-184,
monotone_spline.h), against golden vectors generated from the reference's own headers and, where
the reference flavour of the oracle is built, against it directly.  Plus the reference's own unit
tests of the helpers (test.cpp:1214-1247, 1308-1316) through the C++ drop-in header."""
import json
import os
import subprocess

import numpy as np
import pytest

from fftoptionpricing_b200 import calibration

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def golden_est():
    return json.load(open(os.path.join(ROOT, "tests", "golden", "estimator_v1.json")))


def test_fo_estimate_matches_golden(golden_est):
    u = np.array(golden_est["u"])
    for c in golden_est["cases"]:
        est = calibration.generateFOEstimate(c["strikes"], c["options"], c["stock"], c["r"], c["T"],
                                             c["minStrike"], c["maxStrike"])
        got = est(c["N"], u)
        ref = np.array(c["phiHat_re"]) + 1j * np.array(c["phiHat_im"])
        assert np.max(np.abs(got - ref)) <= 1e-10 * (1 + np.max(np.abs(ref))), c["name"]


def test_fo_estimate_matches_reference_flavour(golden_est, oracles):
    ora = oracles["reference"]
    if ora is None:
        pytest.skip("reference flavour not built here")
    u = np.array(golden_est["u"])
    for c in golden_est["cases"]:
        ref = ora.fo_estimate(c["strikes"], c["options"], c["stock"], c["r"], c["T"], c["minStrike"],
                              c["maxStrike"], c["N"], u)
        got = calibration.generateFOEstimate(c["strikes"], c["options"], c["stock"], c["r"], c["T"],
                                             c["minStrike"], c["maxStrike"])(c["N"], u)
        assert np.max(np.abs(got - ref)) <= 1e-13 * (1 + np.max(np.abs(ref)))


def test_helper_unit_tests_of_the_reference(tmp_path):
    """transformPrices / filter / getKThatIsBelowOne with the reference's expectations."""
    src = tmp_path / "t.cpp"
    src.write_text(r'''
#include "fftop/OptionCalibration.h"
#include <cstdio>
int main() {
  int bad = 0;
  std::vector<double> arr = {3, 4, 5};
  // test.cpp:1214-1222
  bad += optioncal::transformPrices(arr, 4.0, 2.0, 6.0) != std::vector<double>({.5, .75, 1.0, 1.25, 1.5});
  auto id = [](double x, int) { return x; };
  // test.cpp:1224-1236
  auto r1 = optioncal::filter(arr, [](double, int) { return false; }, id, id);
  bad += !(std::get<0>(r1).empty() && std::get<1>(r1) == arr);
  // test.cpp:1238-1247
  auto r2 = optioncal::filter(arr, [](double, int i) { return i % 2 == 0; }, id, id);
  bad += !(std::get<0>(r2) == std::vector<double>({3, 5}) && std::get<1>(r2) == std::vector<double>({4}));
  // test.cpp:1308-1316
  bad += optioncal::getKThatIsBelowOne({.8, .9, 1, 1.1, 1.2}) != 2;
  bad += optioncal::getKThatIsBelowOne({.8, .9, 1.1, 1.2}) != 1;
  // monotone spline interpolates its knots and stays monotone on monotone data
  spline::MonotoneCubic s({0, 1, 2, 4}, {0, 1, 1.5, 1.6});
  bad += !(s(1) == 1 && s(4) == 1.6 && s(0.5) > 0 && s(0.5) < 1 && s(3) >= 1.5 && s(3) <= 1.6);
  std::printf(bad ? "FAIL %d\n" : "PASS\n", bad);
  return bad;
}
''')
    lib = os.path.join(ROOT, "fftoptionpricing_b200")
    exe = str(tmp_path / "t")
    r = subprocess.run(["g++", "-std=c++14", "-O1", "-I", os.path.join(ROOT, "include"), str(src), "-o", exe,
                        "-L", lib, "-lfftop_b200", f"-Wl,-rpath,{lib}", "-L/usr/local/cuda/lib64",
                        "-Wl,-rpath,/usr/local/cuda/lib64"], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-3000:]
    r = subprocess.run([exe], capture_output=True, text=True, timeout=60)
    assert r.returncode == 0 and "PASS" in r.stdout, r.stdout + r.stderr
