#!/usr/bin/env python3
"""Same-named drop-in for src/pacybara_precluster.py (put this directory before the reference's on PATH)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pacybara_b200.precluster import main  # noqa: E402

sys.exit(main())
