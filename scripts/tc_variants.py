"""Runs scripts/config_bench.py's tensor-path lines once per library in scripts/variants/libpinn_tc_*.so."""
import glob, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for lib in sorted(glob.glob(os.path.join(ROOT, "scripts", "variants", "libpinn_tc_*.so"))):
    print(os.path.basename(lib), flush=True)
    r = subprocess.run([sys.executable, os.path.join(ROOT, "scripts", "config_bench.py"), "tensor_short"], env=dict(os.environ, PINN_B200_LIB=lib),
                       capture_output=True, text=True)
    print(r.stdout + r.stderr[-500:], flush=True)
