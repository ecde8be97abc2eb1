"""Time the fused backward at one shape for a list of WHVI_BWD_STAGGER values (one process per value: the library
reads the variable once).    python tools/lab_stagger.py D v0 v1 ..."""
import os, subprocess, sys
D = sys.argv[1]
for v in sys.argv[2:]:
    env = dict(os.environ, WHVI_BWD_STAGGER=v)
    out = subprocess.run([sys.executable, "tools/bench_layer.py", "--dims", D, "--log2n", "27"], env=env, capture_output=True, text=True)
    print("stagger", v, (out.stdout.strip().splitlines() or [out.stderr[-300:]])[-1])
