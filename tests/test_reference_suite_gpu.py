"""Runs the REFERENCE'S OWN test files, unmodified, against this build: the only
reference-held pin for the host logic of operators / hill climbing and an independent check of
the factors and scores (SciPy / lstsq oracles inside those files).

The files are not part of this repository and are never copied into it: they are read from
$PBN_REFERENCE_TESTS, from /root/reference/tests (the build container) or from tests/_reference_tests (a git-ignored
staging copy made by tools/stage_reference_tests.sh so that they travel to the GPU box).  Absent all three the test
is skipped.  `import pybnesian` resolves to the alias package at the repository root (re-export of pybnesian_b200).

Every test of the listed files must pass except the ones in EXPECTED_FAILURES, each with the out-of-scope class or
the reference behaviour it depends on (DESIGN.md section 8 repeats the list).
"""
import json
import os
import subprocess
import sys
import xml.etree.ElementTree as ET

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

FILES = [
    "factors/continuous/KDE_test.py", "factors/continuous/CKDE_test.py", "factors/continuous/ProductKDE_test.py",
    "factors/continuous/LinearGaussianCPD_test.py", "learning/scores/cvlikelihood_test.py",
    "learning/scores/holdoutlikelihood_test.py", "dataset/crossvalidation_test.py", "dataset/holdout_test.py",
    "learning/operators/operatorpool_test.py", "learning/operators/operators_test.py",
    "learning/operators/operatorset_test.py", "learning/operators/operatorstabuset_test.py",
    "learning/algorithms/hillclimbing_test.py",
    # the rows either side of the path and the classes their callers need
    "factors/factor_type_test.py", "factors/discrete/DiscreteFactor_test.py", "learning/parameters/mle_test.py",
    "learning/scores/bic_test.py", "models/BayesianNetwork_test.py", "models/SemiparametricBN_test.py",
    "models/HeterogeneousBN_test.py", "serialization/serialize_factor_test.py", "serialization/serialize_factor_type_test.py",
]

# test id (file::name) -> why it cannot pass on a build scoped to SURVEY.md section 8
EXPECTED_FAILURES = {
    "hillclimbing_test.py::test_hc_conditional_estimate":
        "needs ConditionalGaussianNetwork (conditional Bayesian networks: out of scope, SURVEY.md section 2)",
    "DiscreteFactor_test.py::test_data_type":
        "environment: the test pins the Arrow type pandas 1-2 produce for a Categorical (dictionary<int8, string>, int8 up to "
        "128 categories); this image's pandas 3 / pyarrow 24 produce dictionary<int8, large_string> and int16 codes from 128 "
        "categories on. DiscreteFactor.data_type() reports the type the data arrived with, as the reference does",
}


def reference_tests_dir():
    for d in (os.environ.get("PBN_REFERENCE_TESTS"), "/root/reference/tests", os.path.join(ROOT, "tests", "_reference_tests")):
        if d and os.path.isfile(os.path.join(d, "helpers", "util_test.py")):
            return d
    return None


def run_suite(ref, out_dir):
    os.makedirs(out_dir, exist_ok=True)
    xml = os.path.join(out_dir, "reference_suite.xml")
    env = dict(os.environ, PYTHONPATH=ROOT + os.pathsep + os.path.join(ref, "helpers") + os.pathsep + os.environ.get("PYTHONPATH", ""))
    cmd = [sys.executable, "-m", "pytest", "-q", "-p", "no:cacheprovider", "--rootdir", ref, "-c", os.devnull,
           "--junitxml", xml, "-x" if os.environ.get("PBN_REF_STOP_FIRST") else "-q"] + [os.path.join(ref, f) for f in FILES]
    proc = subprocess.run(cmd, env=env, cwd=out_dir, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=3000)
    results = {}
    for case in ET.parse(xml).getroot().iter("testcase"):
        cid = case.get("classname", "").split(".")[-1] + ".py::" + case.get("name")
        bad = [c for c in case if c.tag in ("failure", "error")]
        skipped = [c for c in case if c.tag == "skipped"]
        results[cid] = ("failed", (bad[0].get("message") or "")[:300]) if bad else (("skipped", "") if skipped else ("passed", ""))
    return results, proc.stdout


def test_reference_suite_unmodified(tmp_path):
    ref = reference_tests_dir()
    if ref is None:
        pytest.skip("the reference's tests/ directory is not available on this machine")
    out_dir = os.path.join(ROOT, "gpurun_out") if os.path.isdir(os.path.join(ROOT, "gpurun_out")) else str(tmp_path)
    results, log = run_suite(ref, out_dir)
    failed = {k: v[1] for k, v in results.items() if v[0] == "failed"}
    summary = {"files": FILES, "collected": len(results), "passed": sum(v[0] == "passed" for v in results.values()),
               "skipped": sum(v[0] == "skipped" for v in results.values()), "failed": failed,
               "expected_failures": EXPECTED_FAILURES}
    with open(os.path.join(out_dir, "reference_suite.json"), "w") as f:
        json.dump(summary, f, indent=1)
    with open(os.path.join(out_dir, "reference_suite.log"), "w") as f:
        f.write(log)
    assert len(results) >= 80, log[-3000:]
    unexpected = sorted(set(failed) - set(EXPECTED_FAILURES))
    assert not unexpected, "reference tests failing unexpectedly: %s\n%s" % (unexpected, log[-6000:])
    fixed = sorted(k for k in EXPECTED_FAILURES if results.get(k, ("", ""))[0] == "passed")
    assert not fixed, "listed as expected failures but passing: %s" % fixed
