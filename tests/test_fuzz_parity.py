"""A short, fixed slice of tests/fuzz_parity.py: random configs x programs, host-compiled engine vs oracle."""
import fuzz_parity


def test_random_configs_and_programs_agree_with_the_oracle(oracle):
    cases, compared, skipped, escalated = fuzz_parity.run(rng_seed=20260923, seconds=8.0)
    assert cases > 50 and compared > 500
