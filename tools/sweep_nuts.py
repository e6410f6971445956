"""Sweep resident CTAs/SM of the NUTS kernel (development aid)."""
import os
import subprocess
import sys

for cap in sys.argv[2:]:
    env = dict(os.environ, SMCB_NUTS_BLOCKS_PER_SM=cap)
    print(f"--- SMCB_NUTS_BLOCKS_PER_SM={cap}", flush=True)
    subprocess.run([sys.executable, "tools/quick_time.py", sys.argv[1]], env=env)
