#!/usr/bin/env python3
"""Drop-in command line of crowsonkb/style_transfer on the B200 engine (see style_transfer_b200/cli.py)."""
import sys

from style_transfer_b200.cli import main

if __name__ == '__main__':
    sys.exit(main())
