"""CPU checks of bench.py's bookkeeping (no GPU): the executed-work model behind roofline.frac against the opcode counts
ncu recorded for the same launch, the parity gate's arithmetic, the reference arm's loader, the sharded layout helpers."""
import glob
import os
import re
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402


def _opcodes(path, kernel):
    counts, inside = {}, False
    for line in open(path):
        if line.startswith("== "):
            inside = kernel in line and not counts
            continue
        m = re.match(r"([A-Z][A-Z0-9_.]*)\s+(\d+)", line)
        if inside and m:
            counts[m.group(1)] = int(m.group(2))
    return counts


def test_executed_work_model_matches_ncu_opcode_counts():
    """roofline.frac = executed FMA lane-operations / time / peak.  The DMMA count follows the kernel's loops exactly;
    the scalar FP64 count is per-group / per-item constants read off the source: both against what ncu counted for
    profile_mom_kernel<4> on config 2 (profiles/r02_ncu_opcodes_cfg2.txt, same synthetic batch)."""
    files = sorted(glob.glob(os.path.join(ROOT, "profiles", "r*_ncu_opcodes_cfg2.txt")))
    if not files:
        pytest.skip("no committed opcode capture")
    op = _opcodes(files[-1], "profile_mom_kernel<4")
    from tapir_b200 import synthetic
    off, rates, T = synthetic.build("cfg2")
    ex = bench.executed_fp64_ops(off, rates, T, 1024, 32, n_mom_epochs=4)
    assert abs(ex["dmma"] - op["DMMA"]) <= 0.005 * op["DMMA"], (ex["dmma"], op["DMMA"])
    scalar = sum(op.get(k, 0) for k in ("DMUL", "DFMA", "DADD", "DSETP", "DMNMX"))
    assert abs(ex["scalar"] - scalar) <= 0.10 * scalar, (ex["scalar"], scalar)
    lane = 256 * op["DMMA"] + 32 * scalar
    assert abs(ex["lane_ops"] - lane) <= 0.03 * lane


def test_site_classes_counts():
    rng = np.random.default_rng(3)
    loci = [rng.gamma(0.5, 0.004, n) for n in (0, 1, 40, 0, 700)]
    loci[2][:5] = [0.0, np.nan, np.inf, 0.5, 1e-200]
    off = np.concatenate([[0], np.cumsum([len(x) for x in loci])]).astype(np.int64)
    rates = np.concatenate(loci)
    n_slow, n_fast = bench.site_classes(off, rates, 32, chunk=64)
    for l, x in enumerate(loci):
        with np.errstate(over="ignore", invalid="ignore"):
            a = 16 * x * x
        act = np.isfinite(a) & (a != 0)
        slow = act & (x > 0) & (x < 0.01953125)
        assert (n_slow[l], n_fast[l]) == (int(slow.sum()), int((act & ~slow).sum())), l
    n0, n1 = bench.site_classes(off, rates, 0)
    assert n0.sum() == 0 and np.array_equal(n1, n_slow + n_fast)


def test_parity_gate_arithmetic():
    rng = np.random.default_rng(0)
    pi = rng.random((5, 40))
    si = rng.random((5, 4))
    se = si * 1e-14
    rows = {"integral": si * (1 + 1e-12), "abserr": se, "argmax_epoch": np.argmax(si, axis=1), "argmax_t": np.argmax(pi, axis=1)}
    good = bench.parity_check((pi, si, se), pi.copy(), rows)
    assert good["ok"] and good["n"] == 5 and good["max_rel_integral"] < 2e-12
    rows["integral"] = si * (1 + 1e-8)
    assert not bench.parity_check((pi, si, se), pi.copy(), rows)["ok"]
    rows["integral"] = si
    rows["argmax_epoch"] = (np.argmax(si, axis=1) + 1) % 4
    assert not bench.parity_check((pi, si, se), pi.copy(), rows)["ok"]


def test_reference_arm_uses_the_reference_file_when_installed():
    """build() copies the reference's unmodified tapir/compute.py to baseline/_ref (git-ignored); the CPU arm then runs
    THAT file's functions.  Its numbers are the port's, bit for bit."""
    ref = bench.load_reference_compute()
    if ref is None:
        assert bench.cpu_kind() == "port"
        pytest.skip("baseline/_ref not installed (no /root/reference at build time)")
    assert bench.cpu_kind() == "reference"
    from oracle import townsend_port as tp
    rng = np.random.default_rng(1)
    lam = rng.gamma(0.5, 0.004, 60)
    tcol = ref.get_time(0, 120)
    assert np.array_equal(ref.get_townsend_pi(tcol, lam), tp.pi_matrix(tp.time_column(0, 120), lam))
    a = ref.get_net_integral_for_epochs(lam, [[32, 42], [88, 98]])
    b = tp.epoch_integrals(lam, [[32, 42], [88, 98]])
    assert a == b
    bench._cpu_init(120, [37, 93], [[32, 42]])
    out = bench._cpu_worker(lam)
    assert np.array_equal(out[3], np.nansum(tp.pi_matrix(tp.time_column(0, 120), lam), axis=1))


def test_small_block_layout_roundtrip():
    from tapir_b200 import sharded
    from tapir_b200.engine import Engine
    L, K, E = 37, 3, 5
    offs, total = sharded.small_layout(L, K, E)
    assert all(o % 256 == 0 for o in offs) and total % 256 == 0 and len(offs) == len(Engine.SMALL_SECTIONS)
    raw = np.zeros(total, dtype=np.uint8)
    rng = np.random.default_rng(2)
    want = {}
    for name, o in zip(Engine.SMALL_SECTIONS, offs):
        dt = np.dtype(sharded._SECTION_DTYPE[name])
        w = {"pi_times": K, "integral": E, "abserr": E}.get(name, 1)
        v = (rng.random(L * w) * 100).astype(dt)
        raw[o:o + v.nbytes] = v.view(np.uint8)
        want[name] = v.reshape(L, w) if w > 1 or name in ("pi_times", "integral", "abserr") else v
    got = sharded.split_small_block(raw, L, K, E)
    for name in Engine.SMALL_SECTIONS:
        assert np.array_equal(got[name], want[name]), name


def test_workload_names_and_default():
    assert bench.DEFAULT_WORKLOAD == "cfg5"
    assert "500000 loci x 800 sites, 2048-point grid" in bench.workload_name("cfg5")
