"""CPU tier: the batched Nelder-Mead driver (ppmmodelpangenome_b200/csrc/nmkb.h) reproduces, bit for bit, the
evaluation sequence of the reference's optimiser header (source/nelder_mead_dfoptim.h) on analytic 8-D test
functions -- with and without speculative batching.  Golden: tests/golden/nm_trajectory.json, produced by
oracle/_ref/nm_ref_driver (the reference header compiled from /root/reference/source)."""
import hashlib
import json
import os
import subprocess

import pytest

from conftest import GOLD, ROOT

EMU = os.path.join(ROOT, "tests", "emu")


@pytest.fixture(scope="module")
def driver():
    subprocess.check_call(["make", "-s", "-C", EMU, "nm_driver"])
    return os.path.join(EMU, "nm_driver")


@pytest.mark.parametrize("which", [0, 1, 2])
@pytest.mark.parametrize("speculate", [1, 0])
def test_trajectory_identical_to_reference_optimiser(driver, which, speculate):
    gold = json.load(open(os.path.join(GOLD, "nm_trajectory.json")))[str(which)]
    r = subprocess.run([driver, str(which), str(speculate)], capture_output=True, text=True, check=True)
    lines = r.stdout.splitlines()
    assert len(lines) == gold["n_lines"]
    for i, ln in gold["sample"].items():
        assert lines[int(i)] == ln
    assert lines[-1] == gold["result"]
    assert hashlib.sha256(r.stdout.encode()).hexdigest() == gold["sha256"]
    # speculation evaluates more points but in fewer launches than logical evaluations
    ev, la = [int(x) for x in r.stderr.split() if x.isdigit()]
    if speculate:
        assert la < 0.8 * gold["n_lines"] and ev > gold["n_lines"]
    else:
        assert ev == gold["n_lines"] - 1


@pytest.mark.parametrize("which", [0, 1, 2])
@pytest.mark.parametrize("level", [1, 2])
@pytest.mark.parametrize("multi", [False, True])
def test_speculation_levels_and_interleaved_multistart_keep_the_trajectory(driver, which, level, multi):
    """Level 1 = the 4 trial points of an iteration per launch, level 2 = + its 8 shrink points (12 per launch); `multi`
    serves three trajectories round by round as the CLI's multi-start driver does.  Every logical evaluation sequence is
    the reference optimiser's, bit for bit."""
    gold = json.load(open(os.path.join(GOLD, "nm_trajectory.json")))[str(which)]
    r = subprocess.run([driver, str(which), "1", str(level)] + (["multi"] if multi else []), capture_output=True, text=True, check=True)
    assert hashlib.sha256(r.stdout.encode()).hexdigest() == gold["sha256"]
    ev, la = [int(x) for x in r.stderr.split() if x.isdigit()]
    assert la < gold["n_lines"]  # fewer launches than logical evaluations


@pytest.mark.skipif(not os.path.exists(os.path.join(ROOT, "oracle", "_ref", "nm_ref_driver")), reason="reference optimiser driver not built")
@pytest.mark.parametrize("which", [0, 1, 2])
def test_golden_is_what_the_reference_header_produces(which):
    gold = json.load(open(os.path.join(GOLD, "nm_trajectory.json")))[str(which)]
    r = subprocess.run([os.path.join(ROOT, "oracle", "_ref", "nm_ref_driver"), str(which)], capture_output=True, text=True, check=True)
    assert hashlib.sha256(r.stdout.encode()).hexdigest() == gold["sha256"]
