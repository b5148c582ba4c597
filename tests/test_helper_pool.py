"""The CUDA-free host pieces of a call, compiled into small harnesses with g++ and run on the CPU: csrc/helper_pool.hpp (the
parked copy threads of a handle) and csrc/staging_copy.hpp (the non-temporal copy into pinned staging)."""
import os
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_every_piece_of_every_run_is_executed_exactly_once(tmp_path):
    exe = tmp_path / "helper_pool_stress"
    subprocess.run(["g++", "-O2", "-std=c++17", "-pthread", "-Wall", "-I", os.path.join(ROOT, "php_aho_corasick_b200", "csrc"),
                    os.path.join(ROOT, "tests", "cpp", "helper_pool_stress.cpp"), "-o", str(exe)], check=True)
    r = subprocess.run([str(exe)], capture_output=True, text=True, timeout=120)
    assert r.returncode == 0 and r.stdout.startswith("ok "), r.stdout + r.stderr


def test_staging_copy_equals_memcpy_for_every_alignment_and_length_class(tmp_path):
    exe = tmp_path / "staging_copy_check"
    subprocess.run(["g++", "-O2", "-std=c++17", "-Wall", "-I", os.path.join(ROOT, "php_aho_corasick_b200", "csrc"),
                    os.path.join(ROOT, "tests", "cpp", "staging_copy_check.cpp"), "-o", str(exe)], check=True)
    r = subprocess.run([str(exe)], capture_output=True, text=True, timeout=120)
    assert r.returncode == 0 and r.stdout.startswith("ok "), r.stdout + r.stderr
