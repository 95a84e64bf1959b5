"""The reference's own unit tests for the hot path (task/mod.rs:839-1254, time/mod.rs:257-305,
net/endpoint.rs:365-585, rand.rs:285-355), re-expressed against the oracle runtime: oracle/selftest.cpp."""
import os
import subprocess


def test_reference_outcome_assertions(oracle):
    exe = os.path.join(oracle.ORACLE_DIR, "oracle_selftest")
    if not os.path.exists(exe):
        subprocess.run(["make", "-C", oracle.ORACLE_DIR, "oracle_selftest"], check=True)
    p = subprocess.run([exe], capture_output=True, text=True, timeout=120)
    assert p.returncode == 0, p.stdout + p.stderr
    assert "SELFTEST PASSED" in p.stdout
    assert p.stdout.count("ok  ") >= 23


def test_check_determinism(oracle):
    import madsim_b200 as m

    assert oracle.check_determinism(m.Config(), m.Workload.pingpong(rounds=5), 3)
