"""csrc/helper_pool.hpp (the parked copy threads of a handle) stressed on the CPU: compiled into a small harness with g++."""
import os
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_every_piece_of_every_run_is_executed_exactly_once(tmp_path):
    exe = tmp_path / "helper_pool_stress"
    subprocess.run(["g++", "-O2", "-std=c++17", "-pthread", "-Wall", "-I", os.path.join(ROOT, "php_aho_corasick_b200", "csrc"),
                    os.path.join(ROOT, "tests", "cpp", "helper_pool_stress.cpp"), "-o", str(exe)], check=True)
    r = subprocess.run([str(exe)], capture_output=True, text=True, timeout=120)
    assert r.returncode == 0 and r.stdout.startswith("ok "), r.stdout + r.stderr
