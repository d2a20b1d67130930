"""Static hygiene: no function in the package, bench.py, scripts or tests reads a global that nothing defines
(a NameError that would only fire when that line runs, e.g. on the GPU box)."""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_no_undefined_global_names():
    r = subprocess.run([sys.executable, os.path.join(ROOT, "scripts", "check_names.py")], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout


def test_checker_catches_an_undefined_name(tmp_path):
    p = tmp_path / "bad.py"
    p.write_text("import os\n\ndef f():\n    return os.getcwd() + missing_name\n")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "scripts", "check_names.py"), str(p)], capture_output=True, text=True)
    assert r.returncode == 1 and "missing_name" in r.stdout
