#!/usr/bin/env python3
"""Same-named drop-in for src/pacybara_calcEdits.R (a Python program: pacybara.sh runs it by name)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pacybara_b200.calc_edits import main  # noqa: E402

sys.exit(main())
