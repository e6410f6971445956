"""Sweep register budget (SMCB_NUTS_MINB) / resident CTAs of the NUTS kernel (development aid)."""
import os
import subprocess
import sys

for v in sys.argv[2:]:
    env = dict(os.environ, SMCB_NUTS_MINB=v)
    print(f"--- SMCB_NUTS_MINB={v}", flush=True)
    subprocess.run([sys.executable, "tools/quick_time.py", sys.argv[1]], env=env)
