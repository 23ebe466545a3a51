"""Dev aid: sweep the look-ahead block width (FAER_B200_NB) for the single-GPU LLT / LU entry points."""
import os
import subprocess
import sys

sizes = sys.argv[1:] or ["16384"]
for nb in ["512", "768", "1024", "1536", "2048"]:
    env = dict(os.environ, FAER_B200_NB=nb)
    print(f"--- FAER_B200_NB={nb}", flush=True)
    subprocess.run([sys.executable, "tools/time_factor.py", "all"] + sizes, env=env)
print("--- recursive drivers (FAER_B200_LOOKAHEAD_MIN_N=0)", flush=True)
subprocess.run([sys.executable, "tools/time_factor.py", "all"] + sizes, env=dict(os.environ, FAER_B200_LOOKAHEAD_MIN_N="0"))
