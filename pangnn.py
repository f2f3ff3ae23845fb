#!/usr/bin/env python3
"""Same command line as the reference's ``pangnn.py`` (flags: ``pangnn_b200/setup.py``), B200 hot path."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))

from pangnn_b200.train import main  # noqa: E402

if __name__ == "__main__":
    main()
