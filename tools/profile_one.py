#!/usr/bin/env python
"""Launch ONE product kernel a few times on bench-shaped inputs (target of an ncu capture).
  python tools/profile_one.py grid|holo|cost
See tools/profile_targets.py for the same workloads run back to back without a profiler."""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.exit(subprocess.call([sys.executable, os.path.join(HERE, "profile_targets.py"), sys.argv[1]]))
