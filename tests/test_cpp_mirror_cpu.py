"""The C++ host side above the C-ABI (include/qmb200.hpp: qm::QMInterface, WbcBase / HierarchicalWbc, SqpMpc, QMController with the reference's
method names and error behaviour) compiles with -Wall -Wextra against the header, links against libqmb200.so and keeps the reference's error
conventions without a GPU: std::invalid_argument on a missing file (QMInterface.cpp:45), a loud failure - never a CPU fallback - on create."""
import os
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def build_demo():
    subprocess.check_call(["make", "-s", "-C", os.path.join(ROOT, "examples")])
    return os.path.join(ROOT, "examples", "_build", "plugin_demo")


def test_cpp_mirror_builds_and_keeps_the_error_conventions():
    import torch
    exe = build_demo()
    out = subprocess.run([exe, os.path.join(ROOT, "assets")], capture_output=True, text=True)
    assert out.returncode == 0, out.stdout + out.stderr
    lines = out.stdout.strip().splitlines()
    assert lines[0].startswith("invalid_argument: [QMInterface] Task file not found")
    if torch.cuda.is_available():
        assert lines[1] == "gpu"
    else:
        assert lines[1].startswith("no-gpu:") and "no CPU fallback" in lines[1]
