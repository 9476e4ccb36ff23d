"""Dry run of the newest GPU tests on a machine without a GPU: the CPU oracle stands in for liblc_gpu.so
(tests/fake_native.py, installed by tests/conftest.py under LC_FAKE_NATIVE=1), so the Python half of those tests — the
mirror classes' marshalling and literal lowering, the tests' own fixtures and expectations — is known to hold before a GPU
is spent on them. It says nothing about the device code; tests that need an entry's HBM image stop there (skipped)."""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_new_gpu_tests_hold_against_the_oracle_backed_stand_in():
    # the files whose Python half is worth a dry run (marshalling-heavy mirrors); tests of device properties proper — budgets,
    # device-planned reads, concurrent callers — have nothing to say against a stand-in
    names = ["test_gpu_zy_fixed_len.py", "test_gpu_zy_ipc_strings.py", "test_gpu_zy_squeeze.py", "test_gpu_zz_multi_column_or.py",
             "test_gpu_ipc.py"]
    files = [os.path.join(ROOT, "tests", n) for n in names]
    env = dict(os.environ, LC_FAKE_NATIVE="1")
    r = subprocess.run([sys.executable, "-m", "pytest", *files, "-m", "gpu", "-q", "-p", "no:cacheprovider"], cwd=ROOT, env=env,
                       capture_output=True, text=True, timeout=1200)
    tail = "\n".join(r.stdout.strip().splitlines()[-15:])
    assert r.returncode == 0, tail
    assert " passed" in tail and "failed" not in tail, tail
