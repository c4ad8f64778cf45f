"""Device time of the two backward kernels separately (CUDA events around each through the profiler-free path:
NERFW_WGRAD_DEBUG bits: 1 wgrad skips MMAs, 2 wgrad skips reductions, 4 pass 1 skips scratch stores, 8 pass 1 only, 16 wgrad only)."""
import os, sys, subprocess
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for dbg in (sys.argv[1:] or ("0", "8", "12", "16", "19")):
    env = dict(os.environ, NERFW_WGRAD_DEBUG=dbg, NERFW_PROFILE_LIB="1")
    print(subprocess.run([sys.executable, os.path.join(ROOT, "scripts", "time_bwd.py")], env=env, capture_output=True, text=True).stdout.strip())
