#!/usr/bin/env python3
"""Build tuning variants of the library (extra -D flags) side by side and run a command against each.

    python tools/tune.py build  tag1=-DA=1,-DB=2  tag2=-DA=3        (here, no GPU needed)
    python tools/tune.py run    tag1,tag2  -- python tools/sweep_sizes.py 2048   (on the GPU box; '' = production lib)
"""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

if sys.argv[1] == "build":
    from kopek_b200 import build as B
    for spec in sys.argv[2:]:
        tag, _, defs = spec.partition("=")
        print(B.build(force=True, tag=tag, defs=[d for d in defs.split(",") if d]), flush=True)
else:
    tags = sys.argv[2].split(",")
    cmd = sys.argv[sys.argv.index("--") + 1:]
    for tag in tags:
        env = dict(os.environ)
        if tag:
            env["KOPEK_TUNE_TAG"] = tag
        print(f"==== {tag or 'production'}", flush=True)
        subprocess.run(cmd, env=env, cwd=ROOT)
