#!/usr/bin/env python
"""frender (B200-native): drop-in for the reference's `frender.py scan|demux` command line."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from frender_b200.cli import main  # noqa: E402

if __name__ == "__main__":
    main()
