"""What tools/ref_probe (the REAL madsim executor, Rust) must print, computed from the oracle.

    python tools/ref_probe/expected.py                 # (re)write tests/golden/ref_probe_expected.json
    python tools/ref_probe/expected.py --check probe.jsonl   # compare a run of the Rust probe

Per (workload, seed): the virtual time of the main future's LAST poll and the next output of the
seed's xoshiro256++ stream at that point (= the number of draws consumed so far), both read off the
oracle's canonical trace.  TEST INFRASTRUCTURE: imports the oracle."""
import json
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, os.path.join(ROOT, "tests"))
sys.path.insert(0, ROOT)
OUT = os.path.join(ROOT, "tests", "golden", "ref_probe_expected.json")
SEEDS = [0, 1, 2, 42, 2333, 4095]
M64 = (1 << 64) - 1


def xoshiro_outputs(seed: int, n: int):
    """rand_xoshiro 0.6 Xoshiro256PlusPlus::seed_from_u64 (SplitMix64 fill) — SURVEY Appendix A.1/A.2"""
    x, s = seed, []
    for _ in range(4):
        x = (x + 0x9E3779B97F4A7C15) & M64
        z = x
        z = ((z ^ (z >> 30)) * 0xBF58476D1CE4E5B9) & M64
        z = ((z ^ (z >> 27)) * 0x94D049BB133111EB) & M64
        s.append(z ^ (z >> 31))
    rotl = lambda v, k: ((v << k) | (v >> (64 - k))) & M64
    out = []
    for _ in range(n):
        out.append((rotl((s[0] + s[3]) & M64, 23) + s[0]) & M64)
        t = (s[1] << 17) & M64
        s[2] ^= s[0]
        s[3] ^= s[1]
        s[1] ^= s[2]
        s[0] ^= s[3]
        s[2] ^= t
        s[3] = rotl(s[3], 45)
    return out


def expected():
    import oracle_py as o
    import madsim_b200 as m

    programs = {"sleep": m.Workload.sleep(10**9), "pingpong": m.Workload.pingpong(rounds=1000),
                "chaos": m.Workload.chaos(rounds=8, period_ms=2000), "raft": m.Workload.raft(),
                "raft_faults": m.Workload.raft(faults=7), "election": m.Workload.election(), "etcd": m.Workload.etcd()}
    rows = []
    for seed in SEEDS:
        for name, wl in programs.items():
            res, tr, _, _ = o.run_seed(m.Config(), wl, seed, trace=True)
            assert res["status"] == m.ST_PASS
            main_polls = tr[tr["kind_id"] == (1 << 24)]  # MSIM_EV_POLL of spawn index 0 = the main future
            last = main_polls[-1]
            rows.append({"workload": name, "seed": seed, "elapsed_ns": int(last["now_ns"]),
                         "next_u64": xoshiro_outputs(seed, int(last["draws"]) + 1)[-1]})
    return rows


def main():
    if len(sys.argv) == 3 and sys.argv[1] == "--check":
        want = {(r["workload"], r["seed"]): r for r in json.load(open(OUT))["rows"]}
        bad = 0
        for line in open(sys.argv[2]):
            if not line.strip():
                continue
            got = json.loads(line)
            w = want[(got["workload"], got["seed"])]
            ok = (got["elapsed_ns"], got["next_u64"]) == (w["elapsed_ns"], w["next_u64"])
            bad += not ok
            print(("ok      " if ok else "MISMATCH"), got, "" if ok else f"expected {w}")
        print("PARITY PINNED against the reference" if not bad else f"{bad} mismatches")
        sys.exit(1 if bad else 0)
    json.dump({"format": 1, "rows": expected()}, open(OUT, "w"), indent=0)
    print("wrote", OUT)


if __name__ == "__main__":
    main()
